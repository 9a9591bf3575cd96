"""TeaCache on the native engine — the algorithm of reference vllm_omni/diffusion/cache/teacache/hook.py:80-217 with the
Qwen-Image extractor (extractors.py:184-246), driven through the engine's staged forward instead of a Python
re-implementation of the blocks:

    PRE   (img_in, txt_in, temb, modulations, block 0's modulated image stream)      qimg_engine_forward_stages(1)
    rel = mean|mod - prev_mod| / (mean|prev_mod| + 1e-8)                              qimg_rel_l1_sums (one pass) + host
    accumulate |poly(rel)|; below the threshold -> reuse:  x_img += previous_residual qimg_bf16_add_inplace
                            else               -> compute: BLOCKS, residual = x - x0  qimg_engine_forward_stages(2), qimg_bf16_sub
    POST  (norm_out, proj_out)                                                        qimg_engine_forward_stages(4)

The reference reads the distance back to the host every forward (`.cpu().item()`, hook.py:204-205).  Here the decision
is taken ON THE DEVICE (`qimg_tea_decide`: one thread reproducing the host arithmetic bit for bit — means in fp32 rounded
to bf16, bf16 division, the fp64 Horner polynomial and accumulator) and the engine's BLOCKS stage is launched predicated
on the resulting flag, so the denoise loop issues no host<->device synchronisation; the (flag, distance) history is read
once, when `decisions` is asked for.  Under tensor parallelism (collectives inside the blocks) the decision stays on the
host, taken by TP rank 0 for the whole group.  Positive / negative CFG branches keep separate states (hook.py:115-122).
"""
from __future__ import annotations

import numpy as np
import torch

from vllm_omni_b200.diffusion.cache.teacache.config import TeaCacheConfig
from vllm_omni_b200.diffusion.cache.teacache.state import TeaCacheState


class TeaCacheHook:
    _HOOK_NAME = "teacache"

    MAX_HISTORY = 4096  # forwards per run recorded in the device-side (flag, distance) history

    def __init__(self, config: TeaCacheConfig):
        self.config = config
        self.rescale_func = np.poly1d(config.coefficients)
        self.states: dict[str, TeaCacheState] = {}
        self._forward_cnt = 0
        self._sums = None       # fp32 [2] device scratch of the reduction kernel
        self._ori = None        # image residual stream before the blocks
        self._flag = None       # int32 [1]: 1 = reuse (device-side decision)
        self._hist = None       # fp32 [2 * MAX_HISTORY]: (flag, rel distance) per forward
        self._accum: dict[str, torch.Tensor] = {}  # branch -> fp64 [1] accumulated rescaled distance (device)
        self._log: list[tuple[str, int]] = []      # (branch, force) per forward of the current run
        self._host_decisions: list[tuple[str, bool, float]] = []
        self._tp_group = None   # set in run() when the transformer is tensor parallel

    # ---- state -------------------------------------------------------------------------------------------------
    def reset_state(self) -> None:
        for st in self.states.values():
            st.reset()
        self._forward_cnt = 0
        self._log = []
        self._host_decisions = []

    @property
    def decisions(self) -> list[tuple[str, bool, float]]:
        """(branch, computed?, rel distance) per forward of the current run.  Device mode: ONE read-back of the history."""
        if self._tp_group is not None or self._hist is None:
            return list(self._host_decisions)
        h = self._hist[: 2 * len(self._log)].cpu().view(-1, 2)
        return [(b, h[i, 0].item() == 0.0, float(h[i, 1])) for i, (b, _) in enumerate(self._log)]

    def _state(self, branch: str) -> TeaCacheState:
        if branch not in self.states:
            self.states[branch] = TeaCacheState()
        return self.states[branch]

    @staticmethod
    def _buf(t: torch.Tensor | None, like: torch.Tensor) -> torch.Tensor:
        if t is None or t.shape != like.shape or t.device != like.device:
            return torch.empty_like(like)
        return t

    # ---- host decision (tensor-parallel runs; hook.py:170-217) ---------------------------------------------------
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
        if self._tp_group is None:
            self._run_device(module, run_stage, mod, x_img, qlib, branch, state)
        else:
            self._run_host(run_stage, mod, x_img, qlib, branch, state)
        state.cnt += 1
        self._forward_cnt += 1
        run_stage(qlib.STAGE_POST)

    def _run_device(self, module, run_stage, mod, x_img, qlib, branch, state):
        dev = mod.device
        if self._flag is None or self._flag.device != dev:
            self._flag = torch.zeros(1, dtype=torch.int32, device=dev)
            self._hist = torch.zeros(2 * self.MAX_HISTORY, dtype=torch.float32, device=dev)
            self._sums = torch.zeros(2, dtype=torch.float32, device=dev)
            self._accum = {}
        if branch not in self._accum:
            self._accum[branch] = torch.zeros(1, dtype=torch.float64, device=dev)
        # which steps are forced is step-count logic, known on the host without looking at data (hook.py:183-190)
        force = 1 if state.cnt == 0 else (2 if not (state.has_mod and state.has_residual) else 0)
        if force == 0:
            qlib.rel_l1_sums(mod, state.previous_modulated_input, self._sums)
        idx = len(self._log)
        if idx >= self.MAX_HISTORY:
            raise RuntimeError("TeaCache history overflow: call refresh() between generations")
        qlib.tea_decide(self._sums, mod.numel(), self.config.coefficients, self.config.rel_l1_thresh, self._accum[branch],
                        self._flag, self._hist, idx, force)
        self._log.append((branch, force))
        state.previous_modulated_input = self._buf(state.previous_modulated_input, mod)
        state.previous_modulated_input.copy_(mod)   # before the blocks reuse the workspace buffer (hook.py:160)
        state.has_mod = True
        self._ori = self._buf(self._ori, x_img)
        self._ori.copy_(x_img)
        state.previous_residual = self._buf(state.previous_residual, x_img)
        lib = qlib.load()
        qlib.check(lib.qimg_engine_set_blocks_predicate(module._engine, self._flag.data_ptr()), "qimg_engine_set_blocks_predicate")
        try:
            run_stage(qlib.STAGE_BLOCKS)            # every kernel exits at once when the flag says "reuse"
        finally:
            qlib.check(lib.qimg_engine_set_blocks_predicate(module._engine, None), "qimg_engine_set_blocks_predicate")
        qlib.tea_residual(x_img, self._ori, state.previous_residual, self._flag)  # reuse: x += r;  compute: r = x - x0
        state.has_residual = True

    def _run_host(self, run_stage, mod, x_img, qlib, branch, state):
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
        self._host_decisions.append((branch, computed, rel))


def apply_teacache_hook(module, config: TeaCacheConfig) -> None:
    """reference hook.py:232-257: attaches the hook to the transformer; its forward then runs through `TeaCacheHook.run`."""
    if not hasattr(module, "_teacache"):
        raise TypeError(f"{type(module).__name__} does not expose the staged forward TeaCache needs")
    module._teacache = TeaCacheHook(config)
