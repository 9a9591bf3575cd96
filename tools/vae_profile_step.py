"""One warm native VAE decode (for `ncu --metrics gpu__time_duration.sum` launch lists; numbers under ncu are not bench values)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vllm_omni_b200 import synthetic  # noqa: E402
from vllm_omni_b200.diffusion.models.qwen_image.vae_decoder import B200VaeDecoder  # noqa: E402

B = int(os.environ.get("VP_B", "1"))
vae = B200VaeDecoder(synthetic.synthetic_vae_decoder_weights(seed=6), device="cuda")
z = torch.randn(B, 16, 1, 128, 128, generator=torch.Generator().manual_seed(1)).cuda()
for _ in range(int(os.environ.get("VP_ITERS", "2"))):
    vae.decode(z, return_dict=False)
torch.cuda.synchronize()
