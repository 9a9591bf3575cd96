"""CPU study (no GPU needed): how much attention accuracy do packed half-precision exponentials cost?

Emulates P = exp2((s - m) c) computed (a) in fp32 (MUFU.EX2 today), (b) with `ex2.approx.f16x2` (input and output rounded to
fp16), (c) with `ex2.approx.ftz.bf16x2`, and (d) f16x2 with a reference maximum that is stale by `delta` (the delayed-maximum
pipelines: P up to 2^delta).  P is rounded to bf16 for the P*V MMA and the row sum is taken from the unrounded P, as in the
kernels.  Output: relative Frobenius error of softmax(QK^T)V against float64.

Result on 2026-09 (S=4224, d=128): fp32 1.6e-3 / f16x2 1.7e-3 / bf16x2 3.6e-3 at |s|max~5; 4.4e-4 / 4.5e-4 / 6.6e-4 at
|s|max~46; f16x2 with delta=8: 1.8e-3 / 4.7e-4.  The error is dominated by the bf16 rounding of P; f16x2 exponentials
are essentially free in accuracy, but fp16's 2^16 range needs the clamp at 15 instead of 96 - and on sm_100a ptxas lowers
`ex2.approx.f16x2` to two MUFU.EX2.F16 per pair, so it saves XU work only if that form issues faster (unmeasured)."""
import math

import torch


def study(S=4224, d=128, qscale=1.0, kscale=1.0, rows=256):
    q = (torch.randn(rows, d) * qscale).bfloat16().float()
    k = (torch.randn(S, d) * kscale).bfloat16().float()
    v = torch.randn(S, d).bfloat16().float()
    s = (q @ k.T) / math.sqrt(d)
    ref = torch.softmax(s.double(), -1) @ v.double()
    x = (s - s.max(-1, keepdim=True).values) * 1.4426950408889634

    def err(p):
        o = (p.bfloat16().float() @ v) / p.sum(-1, keepdim=True)
        return ((o.double() - ref).norm() / ref.norm()).item()

    out = {"fp32": err(torch.exp2(x)), "f16x2": err(torch.exp2(x.half().float()).half().float()),
           "bf16x2": err(torch.exp2(x.bfloat16().float()).bfloat16().float())}
    for delta in (4, 8, 12):
        p = torch.nan_to_num(torch.exp2((x + delta).half().float()).half().float(), posinf=65504.0)
        out[f"f16x2+{delta}"] = err(p)
    return s.abs().max().item(), out


if __name__ == "__main__":
    torch.manual_seed(0)
    for sc in (1, 2, 3):
        mx, out = study(qscale=sc, kscale=sc)
        print(f"|s|max={mx:5.1f}: " + "  ".join(f"{k} {v:.2e}" for k, v in out.items()))
