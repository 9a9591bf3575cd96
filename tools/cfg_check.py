"""CFG-parallel check (torchrun --nproc-per-node 2, 2 GPUs): with true-CFG on, rank 0 runs the positive branch and
rank 1 the negative branch of every denoise step and one all-gather of the noise predictions per step feeds the same
fused combine + Euler kernel on both.  Rank 0 compares the latents with the sequential two-forward path on
one GPU (must be bit-equal: same kernels, same operands) and times both.
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vllm_omni_b200.diffusion.data import OmniDiffusionConfig, TransformerConfig  # noqa: E402
from vllm_omni_b200.diffusion.distributed import parallel_state as ps  # noqa: E402
from vllm_omni_b200.diffusion.request import OmniDiffusionRequest  # noqa: E402
from vllm_omni_b200.diffusion.worker.gpu_worker import GPUWorker  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    assert world == 2
    L = int(os.environ.get("CFG_LAYERS", "4"))
    res = int(os.environ.get("CFG_RES", "512"))
    steps = int(os.environ.get("CFG_STEPS", "4"))
    od = OmniDiffusionConfig(model="synthetic", tf_model_config=TransformerConfig.from_dict({"num_layers": L}),
                             parallel_config={"cfg_parallel_size": 2}, num_gpus=world,
                             master_port=int(os.environ["MASTER_PORT"]), synthetic_weights_seed=0)
    w = GPUWorker(local_rank=int(os.environ["LOCAL_RANK"]), rank=rank, od_config=od)
    g = torch.Generator().manual_seed(1)
    pe = torch.randn(2, 64, 3584, generator=g).bfloat16()
    ne = torch.randn(2, 48, 3584, generator=g).bfloat16()  # a different text length on the negative branch
    lat = torch.cat([w.pipeline.prepare_latents(1, 16, res, res, torch.bfloat16, w.pipeline.device,
                                                torch.Generator().manual_seed(7 + u)) for u in range(2)])

    def req():
        return OmniDiffusionRequest(prompt_embeds=pe, negative_prompt_embeds=ne, latents=lat.clone(), height=res, width=res,
                                    num_inference_steps=steps, true_cfg_scale=4.0, output_type="latent")

    def run():
        out = w.execute_model([req()], od).output
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = w.execute_model([req()], od).output
        torch.cuda.synchronize()
        return out, (time.perf_counter() - t0) * 1e3

    out_par, ms_par = run()
    torch.distributed.barrier()
    ok = True
    if rank == 0:
        ps._STATE.cfg_size = 1  # sequential positive + negative forwards on this GPU alone
        out_seq, ms_seq = run()
        ps._STATE.cfg_size = 2
        same = torch.equal(out_par, out_seq)
        print(f"cfg_check L={L} {res}px {steps} steps: CFG-parallel latents bit-equal to sequential = {same}; "
              f"{ms_par:.1f} ms (2 GPUs) vs {ms_seq:.1f} ms (1 GPU) -> speed-up {ms_seq / ms_par:.2f}x")
        ok = same
    torch.distributed.barrier()
    w.shutdown()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
