"""TeaCache on the native engine — the algorithm of reference vllm_omni/diffusion/cache/teacache/hook.py:80-217 with the
Qwen-Image extractor (extractors.py:184-246), driven through the engine's staged forward instead of a Python
re-implementation of the blocks:

    PRE   (img_in, txt_in, temb, modulations, block 0's modulated image stream)      qimg_engine_forward_stages(1)
    rel = mean|mod - prev_mod| / (mean|prev_mod| + 1e-8)                              qimg_rel_l1_sums (one pass) + host
    accumulate |poly(rel)|; below the threshold -> reuse:  x_img += previous_residual qimg_bf16_add_inplace
                            else               -> compute: BLOCKS, residual = x - x0  qimg_engine_forward_stages(2), qimg_bf16_sub
    POST  (norm_out, proj_out)                                                        qimg_engine_forward_stages(4)

Like the reference (`.cpu().item()`, hook.py:204-205) the decision is taken on the host: one 8-byte read-back per forward.
The relative distance goes through the same bf16 roundings as the reference's tensor expression (means in fp32 rounded to
bf16, bf16 division).  Positive / negative CFG branches keep separate states (hook.py:115-122).
"""
from __future__ import annotations

import numpy as np
import torch

from vllm_omni_b200.diffusion.cache.teacache.config import TeaCacheConfig
from vllm_omni_b200.diffusion.cache.teacache.state import TeaCacheState


class TeaCacheHook:
    _HOOK_NAME = "teacache"

    def __init__(self, config: TeaCacheConfig):
        self.config = config
        self.rescale_func = np.poly1d(config.coefficients)
        self.states: dict[str, TeaCacheState] = {}
        self._forward_cnt = 0
        self._sums = None       # fp32 [2] device scratch of the reduction kernel
        self._ori = None        # image residual stream before the blocks
        self.decisions: list[tuple[str, bool, float]] = []  # (branch, computed?, rel distance) of the current run
        self._tp_group = None   # set in run() when the transformer is tensor parallel

    # ---- state -------------------------------------------------------------------------------------------------
    def reset_state(self) -> None:
        for st in self.states.values():
            st.reset()
        self._forward_cnt = 0
        self.decisions = []

    def _state(self, branch: str) -> TeaCacheState:
        if branch not in self.states:
            self.states[branch] = TeaCacheState()
        return self.states[branch]

    @staticmethod
    def _buf(t: torch.Tensor | None, like: torch.Tensor) -> torch.Tensor:
        if t is None or t.shape != like.shape or t.device != like.device:
            return torch.empty_like(like)
        return t

    # ---- decision (hook.py:170-217) ------------------------------------------------------------------------------
    def _should_compute(self, state: TeaCacheState, mod: torch.Tensor, qlib):
        if state.cnt == 0:
            state.accumulated_rel_l1_distance = 0.0
            return True, float("nan")
        if not state.has_mod:
            return True, float("nan")
        if self._sums is None or self._sums.device != mod.device:
            self._sums = torch.zeros(2, dtype=torch.float32, device=mod.device)
        qlib.rel_l1_sums(mod, state.previous_modulated_input, self._sums)
        if self._tp_group is not None:
            # tensor parallel: the sums come from fp32 atomics in a non-deterministic order, so two ranks could land on
            # different sides of the threshold and one would skip the blocks (and their collectives) while the other runs
            # them.  TP rank 0 decides for the group.
            import torch.distributed as dist
            dist.broadcast(self._sums, src=dist.get_global_rank(self._tp_group, 0), group=self._tp_group)
        s = self._sums.cpu()  # the reference's host sync (hook.py:204-205)
        n = float(mod.numel())
        num = (s[0] / n).to(torch.bfloat16)   # .abs().mean() of a bf16 tensor: fp32 accumulation, bf16 result
        den = (s[1] / n).to(torch.bfloat16)
        rel = (num / (den + 1e-8)).item()
        state.accumulated_rel_l1_distance += abs(float(self.rescale_func(rel)))
        if state.accumulated_rel_l1_distance < self.config.rel_l1_thresh:
            return False, rel
        state.accumulated_rel_l1_distance = 0.0
        return True, rel

    # ---- forward (hook.py:80-165) ------------------------------------------------------------------------------
    def run(self, module, run_stage, mod: torch.Tensor, x_img: torch.Tensor, qlib):
        """`run_stage(mask)` launches engine stages on the current inputs; `mod` / `x_img` are views of the workspace
        (block 0's modulated image stream and the image residual stream)."""
        self._tp_group = getattr(module, "tp_group", None) if getattr(module, "tp_size", 1) > 1 else None
        run_stage(qlib.STAGE_PRE)
        branch = "negative" if (module.do_true_cfg and self._forward_cnt % 2 == 1) else "positive"
        state = self._state(branch)
        should_compute, rel = self._should_compute(state, mod, qlib)
        # keep this step's modulated input now: the blocks reuse its workspace buffer (hook.py:160 does it afterwards)
        state.previous_modulated_input = self._buf(state.previous_modulated_input, mod)
        state.previous_modulated_input.copy_(mod)
        state.has_mod = True
        if not should_compute and state.has_residual:
            qlib.bf16_add_inplace(x_img, state.previous_residual)
            computed = False
        else:
            self._ori = self._buf(self._ori, x_img)
            self._ori.copy_(x_img)
            run_stage(qlib.STAGE_BLOCKS)
            state.previous_residual = self._buf(state.previous_residual, x_img)
            qlib.bf16_sub(state.previous_residual, x_img, self._ori)
            state.has_residual = True
            computed = True
        state.cnt += 1
        self._forward_cnt += 1
        self.decisions.append((branch, computed, rel))
        run_stage(qlib.STAGE_POST)


def apply_teacache_hook(module, config: TeaCacheConfig) -> None:
    """reference hook.py:232-257: attaches the hook to the transformer; its forward then runs through `TeaCacheHook.run`."""
    if not hasattr(module, "_teacache"):
        raise TypeError(f"{type(module).__name__} does not expose the staged forward TeaCache needs")
    module._teacache = TeaCacheHook(config)
