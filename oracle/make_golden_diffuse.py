"""TEST INFRASTRUCTURE ONLY — executes the UNMODIFIED reference denoise loop `QwenImagePipeline.diffuse`
(vllm_omni/diffusion/models/qwen_image/pipeline_qwen_image.py:530-586: timestep broadcast + `/1000`, the two forwards, the
true-CFG combine + norm rescale, the scheduler call) on CPU with the shimmed reference transformer, checks
`oracle.qwen_image_oracle.diffuse` against it and writes tests/golden/diffuse_tiny.pt.

What is NOT the reference's code here: the scheduler object.  diffusers (`FlowMatchEulerDiscreteScheduler`) is an
un-vendored third-party dependency absent from this image, so `self.scheduler` is an adapter around the oracle's
restatement of its published `step` (x32 = x.float() + (sigma_next - sigma) * v -> v.dtype) — "parity unpinned" at that
boundary.  The pipeline module's other third-party imports (VAE, image processor, randn_tensor) are not
on this path and are stubbed by name only.      python -m oracle.make_golden_diffuse
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import qwen_image_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402
from vllm_omni_b200 import synthetic  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "diffuse_tiny.pt")
CASE = dict(L=2, H=2, joint=256, B=2, grid=(8, 6), T=20, T_neg=14, seed=31, steps=4, true_cfg_scale=4.0)


def import_reference_pipeline():
    st = ref_shim.init_reference()

    def mod(name, **attrs):
        m = sys.modules.get(name) or types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m

    class _NotOnThisPath:
        def __init__(self, *a, **k):
            pass

    mod("diffusers.image_processor", VaeImageProcessor=_NotOnThisPath)
    mod("diffusers.models.autoencoders")
    mod("diffusers.models.autoencoders.autoencoder_kl_qwenimage", AutoencoderKLQwenImage=_NotOnThisPath)
    mod("diffusers.schedulers")
    mod("diffusers.schedulers.scheduling_flow_match_euler_discrete", FlowMatchEulerDiscreteScheduler=_NotOnThisPath)
    mod("diffusers.utils")
    mod("diffusers.utils.torch_utils", randn_tensor=lambda *a, **k: None)
    base = os.path.join(ref_shim.REF_ROOT, "vllm_omni")
    for n, rel in (("vllm_omni.diffusion.model_loader", "diffusion/model_loader"), ("vllm_omni.model_executor", "model_executor"),
                   ("vllm_omni.model_executor.model_loader", "model_executor/model_loader"),
                   ("vllm_omni.diffusion.distributed", "diffusion/distributed")):
        if n not in sys.modules:
            ref_shim._ns(n, os.path.join(base, rel))
    import vllm_omni.diffusion.models.qwen_image.pipeline_qwen_image as RP
    return st, RP


class SchedulerAdapter:
    """Stands in for diffusers' FlowMatchEulerDiscreteScheduler (absent): explicit sigma table, `step` = the oracle's
    restatement of the published Euler update."""

    def __init__(self, sigmas: np.ndarray):
        self.sigmas = torch.from_numpy(np.asarray(sigmas, dtype=np.float32))
        self._step_index = 0

    def set_begin_index(self, i: int = 0):
        self._step_index = i

    def step(self, model_output, timestep, sample, return_dict=False):
        i = self._step_index
        self._step_index += 1
        return (O.euler_step(model_output, sample, self.sigmas[i], self.sigmas[i + 1]),)


def main():
    c = CASE
    st, RP = import_reference_pipeline()
    dims = O.DiTDims(num_layers=c["L"], num_heads=c["H"], joint_dim=c["joint"])
    w = dict(synthetic.synthetic_weights(c["L"], seed=c["seed"], dtype=torch.bfloat16, norm_jitter=0.1,
                                         num_heads=c["H"], joint_dim=c["joint"]))
    h, w_ = c["grid"]
    g = torch.Generator().manual_seed(400 + c["seed"])
    lat0 = torch.randn((c["B"], h * w_, 64), generator=g).bfloat16()
    pe = torch.randn((c["B"], c["T"], c["joint"]), generator=g).bfloat16()
    ne = torch.randn((c["B"], c["T_neg"], c["joint"]), generator=g).bfloat16()  # a different negative text length
    sig = O.flow_match_sigmas(c["steps"], h * w_)
    timesteps = torch.from_numpy(sig[:-1]) * 1000.0  # fp32, as scheduler.timesteps

    model, od = ref_shim.build_reference_model(c["L"], torch.bfloat16, num_attention_heads=c["H"], joint_attention_dim=c["joint"])
    with torch.no_grad():
        for k, p in model.named_parameters():
            p.copy_(w[k])
    out = {}
    for cfg in (False, True):
        pipe = object.__new__(RP.QwenImagePipeline)  # no __init__: no checkpoint, text encoder or VAE on this path
        torch.nn.Module.__init__(pipe)
        pipe.transformer = model
        pipe.scheduler = SchedulerAdapter(sig)
        pipe._interrupt = False
        pipe._attention_kwargs = None
        shapes = [[(1, h, w_)]] * c["B"]
        with torch.inference_mode(), st["set_forward_context"](vllm_config=st["vc"], omni_diffusion_config=od):
            ref = pipe.diffuse(pe, torch.ones(c["B"], c["T"], dtype=torch.long), ne if cfg else None,
                               torch.ones(c["B"], c["T_neg"], dtype=torch.long) if cfg else None, lat0.clone(), shapes,
                               [c["T"]] * c["B"], [c["T_neg"]] * c["B"] if cfg else None, timesteps, cfg, None, c["true_cfg_scale"])
        mine = O.diffuse(w, dims, lat0.clone(), pe, ne if cfg else None, sig, (1, h, w_), c["true_cfg_scale"])
        err = O.rel_fro(mine, ref)
        print(f"cfg={cfg}: oracle diffuse vs reference diffuse rel_fro = {err:.3e}, max|d| = {(mine.float() - ref.float()).abs().max():.3e}")
        assert err <= 1e-6, "oracle.diffuse deviates from the reference loop"
        out["cfg" if cfg else "nocfg"] = ref.clone()
    torch.save(dict(case=c, latents0=lat0, prompt_embeds=pe, negative_prompt_embeds=ne, sigmas=sig, **out), GOLDEN)
    print("saved", GOLDEN)


if __name__ == "__main__":
    main()
