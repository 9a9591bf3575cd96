"""TEST INFRASTRUCTURE ONLY — runs the UNMODIFIED reference TeaCache hook
(vllm_omni/diffusion/cache/teacache/hook.py, extractor for QwenImageTransformer2DModel) on the shimmed reference
transformer on CPU over a short denoise trajectory, checks oracle/teacache_oracle.py against it (decisions and outputs)
and writes tests/golden/teacache_tiny.pt.      python -m oracle.make_golden_teacache
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import qwen_image_oracle as O  # noqa: E402
from oracle import ref_shim, teacache_oracle as TO  # noqa: E402
from vllm_omni_b200 import synthetic  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "teacache_tiny.pt")
CASE = dict(L=2, H=2, joint=256, B=2, grid=(8, 6), T=20, seed=21, steps=8, thresh=0.12,
            coefficients=[0.0, 0.0, 0.0, 1.0, 0.0])  # identity rescale: random weights are far from the tuned polynomial's range


def main():
    c = CASE
    dims = O.DiTDims(num_layers=c["L"], num_heads=c["H"], joint_dim=c["joint"])
    w = dict(synthetic.synthetic_weights(c["L"], seed=c["seed"], dtype=torch.bfloat16, norm_jitter=0.1,
                                         num_heads=c["H"], joint_dim=c["joint"]))
    h, w_ = c["grid"]
    g = torch.Generator().manual_seed(300 + c["seed"])
    lat0 = torch.randn((c["B"], h * w_, 64), generator=g).bfloat16()
    pe = torch.randn((c["B"], c["T"], c["joint"]), generator=g).bfloat16()
    ne = torch.randn((c["B"], c["T"], c["joint"]), generator=g).bfloat16()
    sig = O.flow_match_sigmas(c["steps"], h * w_)

    ref_shim.init_reference()
    base = os.path.join(ref_shim.REF_ROOT, "vllm_omni", "diffusion", "cache")
    # bare namespace modules: the package __init__s pull in cache-dit and the CacheBackend selector, which are not on this path
    ref_shim._ns("vllm_omni.diffusion.cache", base)
    ref_shim._ns("vllm_omni.diffusion.cache.teacache", os.path.join(base, "teacache"))
    from vllm_omni.diffusion.cache.teacache.config import TeaCacheConfig
    from vllm_omni.diffusion.cache.teacache.hook import TeaCacheHook, apply_teacache_hook
    from vllm_omni.diffusion.hooks import HookRegistry

    out = {}
    for cfg in (False, True):
        model, od = ref_shim.build_reference_model(c["L"], torch.bfloat16, num_attention_heads=c["H"], joint_attention_dim=c["joint"])
        with torch.no_grad():
            for k, p in model.named_parameters():
                p.copy_(w[k])
        apply_teacache_hook(model, TeaCacheConfig(rel_l1_thresh=c["thresh"], coefficients=c["coefficients"],
                                                  transformer_type="QwenImageTransformer2DModel"))
        hook = HookRegistry.get_or_create(model).get_hook(TeaCacheHook._HOOK_NAME)
        computed = []
        orig = hook._should_compute_full_transformer

        def spy(state, mod, _orig=orig, _log=computed):
            r = _orig(state, mod)
            _log.append(bool(r))
            return r

        hook._should_compute_full_transformer = spy
        model.do_true_cfg = cfg
        lat = lat0.clone()
        sigt = torch.from_numpy(sig)
        mask = torch.ones(c["B"], c["T"], dtype=torch.long)
        for i in range(c["steps"]):
            t = (sigt[i] * 1000.0).expand(c["B"]).to(torch.bfloat16)
            kw = dict(hidden_states=lat, encoder_hidden_states_mask=mask, timestep=t / 1000,
                      img_shapes=[[(1, h, w_)]] * c["B"], txt_seq_lens=[c["T"]] * c["B"])
            noise = ref_shim.run_reference_model(model, od, encoder_hidden_states=pe, **kw)
            if cfg:
                neg = ref_shim.run_reference_model(model, od, encoder_hidden_states=ne, **kw)
                noise = O.cfg_combine(noise, neg, 4.0)
            lat = O.euler_step(noise, lat, sigt[i], sigt[i + 1])
        tc = TO.TeaCacheOracle(w, dims, c["thresh"], c["coefficients"])
        mine = TO.diffuse(tc, lat0.clone(), pe, ne if cfg else None, sig, (1, h, w_), 4.0)
        mine_dec = [d[1] for d in tc.decisions]
        err = O.rel_fro(mine, lat)
        print(f"cfg={cfg}: reference decisions {computed}\n          oracle    decisions {mine_dec}\n          latents rel_fro = {err:.3e}")
        assert computed == mine_dec, "TeaCache restatement takes different compute/reuse decisions than the reference hook"
        assert err <= 1e-6, "TeaCache restatement deviates from the reference hook"
        assert any(computed[1:]) and not all(computed), "choose a threshold that exercises both paths"
        out["cfg" if cfg else "nocfg"] = dict(latents=lat.clone(), decisions=computed)
        del model
    torch.save(dict(case=c, latents0=lat0, prompt_embeds=pe, negative_prompt_embeds=ne, sigmas=sig, **out), GOLDEN)
    print("saved", GOLDEN)


if __name__ == "__main__":
    main()
