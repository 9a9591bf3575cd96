"""Process-group bookkeeping for the native runner — the part of the reference's
`vllm_omni/diffusion/distributed/parallel_state.py` (:391-713) the Qwen-Image DiT path needs.

One process per GPU, `torch.distributed` (NCCL on GPUs, gloo in CPU tests).  Rank order follows the
reference's "tp-sp-pp-cfg-dp" string (:659): TP ranks are adjacent, DP is the outermost dimension.
  * DP (data parallel over images): independent units, NO collective on the data path; only the final
    gather of the [B/P, S_img, 64] latents to rank 0.  Not wired in the reference (groups only, :661-668).
  * TP group: kept for the tensor-parallel engine mode (all-reduce of row-parallel partial sums).
  * SP group (`ulysses_degree`, the reference's own multi-GPU mode for this model, attention/parallel/ulysses.py): ranks that
    split the ROWS of one forward between them (fused sequence parallelism: full weights per rank, the two all-to-alls of
    Ulysses are peer stores of the QKV-GEMM / attention epilogues; no collective call on the data path).
  * CFG group (size 1 or 2): with true-CFG on, the conditional and unconditional forwards of a step are independent;
    rank 0 of the group runs the positive branch, rank 1 the negative one, and one all-gather of the [B,S_img,64]
    noise predictions per step lets both apply the same fused combine + Euler step (SURVEY §8e "CFG parallel").
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass
class _State:
    world_size: int = 1
    rank: int = 0
    dp_size: int = 1
    tp_size: int = 1
    cfg_size: int = 1
    sp_size: int = 1
    sp_group: "dist.ProcessGroup | None" = None
    dp_group: "dist.ProcessGroup | None" = None
    tp_group: "dist.ProcessGroup | None" = None
    cfg_group: "dist.ProcessGroup | None" = None


_STATE = _State()


def get_torch_distributed_backend() -> str:
    """reference envs.py:113-115 -> "nccl" on CUDA; gloo when no GPU is visible (CPU tests)."""
    return "nccl" if torch.cuda.is_available() else "gloo"


def init_distributed_environment(world_size: int = -1, rank: int = -1, backend: str | None = None,
                                 distributed_init_method: str = "env://"):
    """reference parallel_state.py:391-430."""
    backend = backend or get_torch_distributed_backend()
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {}
        if world_size > 0:
            kw.update(world_size=world_size, rank=rank)
        dist.init_process_group(backend=backend, init_method=distributed_init_method, **kw)
    _STATE.world_size, _STATE.rank = dist.get_world_size(), dist.get_rank()


def initialize_model_parallel(data_parallel_size: int = 1, tensor_parallel_size: int = 1, backend: str | None = None,
                              cfg_parallel_size: int = 1, ulysses_degree: int = 1, **unused_reference_kwargs):
    """reference parallel_state.py:563-713 (DP, CFG, SP(ulysses) and TP groups are created; PP / ring groups are dead
    scaffolding for this model — SURVEY §2b).  rank = ((dp * cfg_size + cfg) * sp_size + sp) * tp_size + tp, the reference's
    "tp-sp-pp-cfg-dp" order."""
    ws = dist.get_world_size() if dist.is_initialized() else 1
    if cfg_parallel_size not in (1, 2):
        raise ValueError("cfg_parallel_size must be 1 or 2 (positive / negative branch)")
    sp = int(ulysses_degree or 1)
    if data_parallel_size * cfg_parallel_size * sp * tensor_parallel_size != ws:
        raise ValueError(f"dp({data_parallel_size}) * cfg({cfg_parallel_size}) * sp({sp}) * tp({tensor_parallel_size}) != world_size({ws})")
    _STATE.dp_size, _STATE.tp_size, _STATE.cfg_size, _STATE.sp_size = data_parallel_size, tensor_parallel_size, cfg_parallel_size, sp
    _STATE.tp_group = _STATE.cfg_group = _STATE.dp_group = _STATE.sp_group = None
    rank = dist.get_rank() if dist.is_initialized() else 0
    _STATE.rank = rank
    if ws == 1:
        return
    tp, cfg = tensor_parallel_size, cfg_parallel_size

    def rk(d, c, s_, t):
        return ((d * cfg + c) * sp + s_) * tp + t

    def make(groups):
        mine = None
        for ranks in groups:
            g = dist.new_group(ranks, backend=backend)
            if rank in ranks:
                mine = g
        return mine

    dp_r, cfg_r, sp_r, tp_r = range(data_parallel_size), range(cfg), range(sp), range(tp)
    _STATE.tp_group = make([[rk(d, c, s_, t) for t in tp_r] for d in dp_r for c in cfg_r for s_ in sp_r])
    _STATE.sp_group = make([[rk(d, c, s_, t) for s_ in sp_r] for d in dp_r for c in cfg_r for t in tp_r])
    _STATE.cfg_group = make([[rk(d, c, s_, t) for c in cfg_r] for d in dp_r for s_ in sp_r for t in tp_r])
    _STATE.dp_group = make([[rk(d, c, s_, t) for d in dp_r] for c in cfg_r for s_ in sp_r for t in tp_r])


def get_world_size() -> int:
    return _STATE.world_size


def get_data_parallel_world_size() -> int:
    return _STATE.dp_size


def get_data_parallel_rank() -> int:
    return _STATE.rank // (_STATE.tp_size * _STATE.sp_size * _STATE.cfg_size)


def get_sequence_parallel_world_size() -> int:
    return _STATE.sp_size


def get_sequence_parallel_rank() -> int:
    return (_STATE.rank // _STATE.tp_size) % _STATE.sp_size


def get_sp_group():
    return _STATE.sp_group


def get_cfg_parallel_world_size() -> int:
    return _STATE.cfg_size


def get_cfg_parallel_rank() -> int:
    return (_STATE.rank // (_STATE.tp_size * _STATE.sp_size)) % _STATE.cfg_size


def get_cfg_group():
    return _STATE.cfg_group


def cfg_all_gather(local: torch.Tensor) -> list[torch.Tensor]:
    """[positive, negative] noise predictions on both ranks of the CFG group (one exchange per denoise step)."""
    bufs = [torch.empty_like(local) for _ in range(_STATE.cfg_size)]
    dist.all_gather(bufs, local.contiguous(), group=_STATE.cfg_group)
    return bufs


def any_rank_in_model_group(flag: bool, device=None) -> bool:
    """Logical OR of `flag` over the ranks that compute ONE trajectory together (the TP, SP and CFG groups; DP
    replicas are independent and are not consulted).  Used for decisions every such rank must take identically."""
    if _STATE.tp_size == 1 and _STATE.cfg_size == 1 and _STATE.sp_size == 1:
        return bool(flag)
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device if dist.get_backend() == "nccl" else "cpu")
    if _STATE.tp_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=_STATE.tp_group)
    if _STATE.sp_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=_STATE.sp_group)
    if _STATE.cfg_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=_STATE.cfg_group)
    return bool(int(t.item()))


def get_tensor_model_parallel_world_size() -> int:
    return _STATE.tp_size


def get_tensor_model_parallel_rank() -> int:
    return _STATE.rank % _STATE.tp_size


def get_dp_group():
    return _STATE.dp_group


def get_tp_group():
    return _STATE.tp_group


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split of n_items units over `world` ranks (first n%world ranks get one extra)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_to_rank0(local: torch.Tensor, counts: list[int]) -> torch.Tensor | None:
    """Gather per-rank [n_i, ...] tensors along dim 0 on DP-rank 0 (the only exchange of the DP path)."""
    if _STATE.dp_size == 1:
        return local
    group = _STATE.dp_group
    dp_rank = get_data_parallel_rank()
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(_STATE.dp_size)] if dp_rank == 0 else None
    dst = dist.get_global_rank(group, 0) if group is not None else 0
    dist.gather(pad, bufs, dst=dst, group=group)
    if dp_rank != 0:
        return None
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)


def destroy_distributed_env():
    if dist.is_initialized():
        dist.destroy_process_group()
    global _STATE
    _STATE = _State()
