"""Multi-GPU data-parallel check (run under torchrun --nproc-per-node N on a box with N GPUs):
every rank builds a GPUWorker (synthetic weights, reduced depth), the request's images are sharded over the DP
ranks, rank 0 gathers the latents and compares them with the same request executed on a single GPU.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tools/dp_check.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vllm_omni_b200.diffusion.data import OmniDiffusionConfig, TransformerConfig  # noqa: E402
from vllm_omni_b200.diffusion.distributed import parallel_state as ps  # noqa: E402
from vllm_omni_b200.diffusion.request import OmniDiffusionRequest  # noqa: E402
from vllm_omni_b200.diffusion.worker.gpu_worker import GPUWorker  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    L = int(os.environ.get("DP_LAYERS", "4"))
    od = OmniDiffusionConfig(model="synthetic", tf_model_config=TransformerConfig.from_dict({"num_layers": L}),
                             parallel_config={"data_parallel_size": world}, num_gpus=world,
                             master_port=int(os.environ["MASTER_PORT"]), synthetic_weights_seed=0)
    w = GPUWorker(local_rank=int(os.environ["LOCAL_RANK"]), rank=rank, od_config=od)
    g = torch.Generator().manual_seed(1)
    n_img = world + 1  # deliberately not divisible
    pe = torch.randn(n_img, 64, 3584, generator=g).bfloat16()
    req = OmniDiffusionRequest(prompt_embeds=pe, seed=123, height=512, width=512, num_inference_steps=4, true_cfg_scale=1.0,
                               output_type="latent")
    out = w.execute_model([req], od)
    ok = True
    if rank == 0:
        assert out.output.shape == (n_img, 1024, 64), out.output.shape
        # single-GPU reference on rank 0's pipeline with the same per-unit noise
        lat = torch.cat([w.pipeline.prepare_latents(1, 16, 512, 512, torch.bfloat16, w.pipeline.device,
                                                    torch.Generator().manual_seed(123 + u)) for u in range(n_img)])
        ref = w.pipeline.forward(OmniDiffusionRequest(prompt_embeds=pe, latents=lat, height=512, width=512, num_inference_steps=4,
                                                      true_cfg_scale=1.0, output_type="latent")).output
        err = float((out.output.float() - ref.float()).norm() / ref.float().norm())
        print(f"dp_check world={world}: gathered {tuple(out.output.shape)}, rel_fro vs single-GPU = {err:.3e}")
        ok = err <= 1e-2
    torch.distributed.barrier()
    w.shutdown()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
