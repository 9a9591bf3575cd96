"""Tensor-parallel check (torchrun --nproc-per-node P, P GPUs): every rank holds the head / FFN shards of the same
synthetic checkpoint and runs the TP engine (TP_COMM=nccl: NCCL all-reduce of the row-parallel partial sums; TP_COMM=p2p: the fused
peer-memory reduce + epilogue + all-gather kernel); rank 0 compares the
output with the single-GPU engine on the same inputs (tolerance 1e-2, the reference's SP-vs-baseline convention,
tests/diffusion/attention/test_ulysses_sequence_parallel.py:332-343), and times both.
"""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vllm_omni_b200 import synthetic  # noqa: E402
from vllm_omni_b200.diffusion.distributed import parallel_state as ps  # noqa: E402
from vllm_omni_b200.diffusion.models.qwen_image.qwen_image_transformer import QwenImageTransformer2DModel  # noqa: E402


def build(L, dev, **kw):
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(dev):
            m = QwenImageTransformer2DModel(num_layers=L, **kw)
    finally:
        torch.set_default_dtype(torch.float32)
    m.load_weights(synthetic.synthetic_weights(L, seed=0, norm_jitter=0.1))
    return m


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    L = int(os.environ.get("TP_LAYERS", "4"))
    res = int(os.environ.get("TP_RES", "512"))
    B = int(os.environ.get("TP_BATCH", "1"))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    ps.init_distributed_environment(world_size=world, rank=rank, backend="nccl")
    ps.initialize_model_parallel(data_parallel_size=1, tensor_parallel_size=world, backend="nccl")
    comm = os.environ.get("TP_COMM", "nccl")  # "nccl": all-reduce callback; "p2p": fused peer-memory reduction kernel
    m = build(L, dev, tp_size=world, tp_rank=ps.get_tensor_model_parallel_rank(), tp_group=ps.get_tp_group(), tp_comm=comm)
    lat, txt = synthetic.synthetic_inputs(B, res, res, 64)
    t = torch.tensor([0.5], dtype=torch.bfloat16, device=dev)
    grid = [[(1, res // 16, res // 16)]] * B
    args = (lat.to(dev), txt.to(dev), None, t, grid, [64] * B)

    def run(model, n=3):
        out = model(*args, return_dict=False, uniform_timestep=True)[0]
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            out = model(*args, return_dict=False, uniform_timestep=True)[0]
        torch.cuda.synchronize()
        return out, (time.perf_counter() - t0) / n * 1e3

    out_tp, ms_tp = run(m)
    ok = True
    if comm == "p2p" and not m.p2p_healthy():
        print(f"rank {rank}: peer-memory barrier timed out", flush=True)
        ok = False
    if rank == 0:
        m1 = build(L, dev, tp_size=1)
        out_1 = m1(*args, return_dict=False, uniform_timestep=True)[0]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            out_1 = m1(*args, return_dict=False, uniform_timestep=True)[0]
        torch.cuda.synchronize()
        ms_1 = (time.perf_counter() - t0) / 3 * 1e3
        err = float((out_tp.float() - out_1.float()).norm() / out_1.float().norm())
        print(f"tp_check tp={world} comm={comm} L={L} {res}px B={B}: rel_fro(TP, single GPU) = {err:.3e}; "
              f"forward {ms_tp:.2f} ms (TP{world}) vs {ms_1:.2f} ms (1 GPU) -> speed-up {ms_1 / ms_tp:.2f}x")
        ok = ok and err <= 1e-2 and not torch.isnan(out_tp).any()
    dist.barrier()
    ps.destroy_distributed_env()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
