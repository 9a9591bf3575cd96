"""RoPE table helper.  In the reference `RotaryEmbedding.forward_cuda`
(vllm_omni/diffusion/layers/rope.py:88-106, flash-attn Triton kernel) is a separate pass over
q and k; in the B200 engine the interleaved rotation is fused into the QKV GEMM epilogue
(csrc/qimg_gemm.cuh, EPI_QKV), so there is no stand-alone rotary op on the hot path.  This
module only documents the convention the epilogue implements (is_neox_style=False):
    out[2i]   = x[2i] * cos[i] - x[2i+1] * sin[i]
    out[2i+1] = x[2i+1] * cos[i] + x[2i] * sin[i]        with bf16 cos/sin (:403-406)
"""
INTERLEAVED = True
