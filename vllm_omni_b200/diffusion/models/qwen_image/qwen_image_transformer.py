"""B200-native `QwenImageTransformer2DModel` — host-side mirror of the reference class
(vllm_omni/diffusion/models/qwen_image/qwen_image_transformer.py:609-839).

Keeps the reference's constructor signature, attribute names (`transformer_blocks`,
`img_in`, `txt_in`, `txt_norm`, `time_text_embed`, `pos_embed`, `norm_out`, `proj_out`,
`do_true_cfg`, `in_channels`, `guidance_embeds`), parameter names (so the same
`(name, tensor)` checkpoint stream loads, incl. the q/k/v stacking of :805-815) and the
`forward(...)` contract (:692-802) — but the modules are thin parameter holders: the
arithmetic of the whole forward runs in the sm_100a C-ABI engine
(csrc/qimg_engine.cu -> tcgen05 GEMMs / FMHA / fused elementwise kernels).  There is no
PyTorch compute path and no fallback: without the built extension and an sm_100 device
`forward` raises.
"""
from __future__ import annotations

import os

import ctypes as C
from collections.abc import Iterable
from dataclasses import dataclass
from typing import Any

import torch
import torch.nn as nn

from vllm_omni_b200 import lib as qlib


@dataclass
class Transformer2DModelOutput:
    """Stand-in for diffusers' Transformer2DModelOutput; supports `out[0]` like BaseOutput
    (the pipeline indexes `[0]`, pipeline_qwen_image.py:556-566)."""

    sample: torch.Tensor

    def __getitem__(self, i):
        return (self.sample,)[i]


def _grids(img_shapes) -> tuple[tuple[int, int, int], ...]:
    """The reference's reading of `img_shapes` (QwenEmbedRope.forward :231-234): a list is indexed by sample and sample 0
    decides; its entry is a list of (frame, h, w) grids — `[[(1, h, w)]] * B` in the T2I pipeline
    (pipeline_qwen_image.py:688), `[[(1, h, w), (1, h2, w2), ...]] * B` in the edit pipelines (noisy latents first, then
    the condition image(s), pipeline_qwen_image_edit.py:602) — or a single grid tuple.  Unlike the reference, samples that
    disagree with sample 0 are rejected instead of silently sharing its table."""
    def is_grid(x):
        return isinstance(x, tuple) and len(x) == 3 and all(isinstance(v, int) for v in x)

    v = img_shapes
    if isinstance(v, list):
        if not v:
            raise ValueError("empty img_shapes")
        if any(e != v[0] for e in v):
            raise NotImplementedError("all samples of a batch must share the same image grids")
        v = v[0]
    if not isinstance(v, list):
        v = [v]
    grids = tuple(tuple(g) for g in v)
    if not grids or not all(is_grid(g) for g in grids):
        raise ValueError(f"bad img_shapes {img_shapes!r}")
    return grids


class _Linear(nn.Module):
    """Parameter holder with nn.Linear's parameter names (weight [out,in], bias [out])."""

    def __init__(self, in_features: int, out_features: int, weight: torch.Tensor | None = None,
                 bias: torch.Tensor | None = None):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features) if weight is None else weight, requires_grad=False)
        self.bias = nn.Parameter(torch.empty(out_features) if bias is None else bias, requires_grad=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # layer-level plug-in path (tcgen05 GEMM)
        shp = x.shape
        y = qlib.linear(x.reshape(-1, shp[-1]).contiguous(), self.weight, self.bias)
        return y.view(*shp[:-1], self.out_features)


class _RMSNormWeight(nn.Module):
    def __init__(self, dim: int, eps: float):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim), requires_grad=False)
        self.variance_epsilon = eps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        shp = x.shape
        return qlib.rms_norm(x.reshape(-1, shp[-1]).contiguous(), self.weight, self.variance_epsilon).view(shp)


class _GELUProj(nn.Module):
    def __init__(self, dim: int, inner: int):
        super().__init__()
        self.proj = _Linear(dim, inner)


class _FeedForward(nn.Module):
    """diffusers FeedForward parameter layout: net.0.proj, net.2 (qwen_image_transformer.py:491,501).
    Under tensor parallelism net.0.proj holds FF/P rows (column-parallel) and net.2 FF/P input columns
    (row-parallel; its bias is kept whole and applied after the all-reduce)."""

    def __init__(self, dim: int, tp: int = 1):
        super().__init__()
        self.net = nn.ModuleList([_GELUProj(dim, 4 * dim // tp), nn.Identity(), _Linear(4 * dim // tp, dim)])


class _Attn(nn.Module):
    """QwenImageCrossAttention parameter layout (:288-368)."""

    def __init__(self, dim: int, head_dim: int, eps: float, tp: int = 1):
        super().__init__()
        # TP: q|k|v rows of the local heads (column-parallel); out-projections take the local heads' columns
        self.to_qkv = _Linear(dim, 3 * dim // tp)
        self.add_kv_proj = _Linear(dim, 3 * dim // tp)
        self.norm_q = _RMSNormWeight(head_dim, eps)
        self.norm_k = _RMSNormWeight(head_dim, eps)
        self.norm_added_q = _RMSNormWeight(head_dim, eps)
        self.norm_added_k = _RMSNormWeight(head_dim, eps)
        self.to_out = nn.ModuleList([_Linear(dim // tp, dim)])
        self.to_add_out = _Linear(dim // tp, dim)


class _AdaLNHolder(nn.Module):
    """Stands where the reference has AdaLayerNorm modules (img_norm1, ...): no parameters."""

    def __init__(self, dim: int, eps: float):
        super().__init__()
        self.hidden_size, self.eps = dim, eps


class QwenImageTransformerBlock(nn.Module):
    """Parameter holder with the reference block's sub-module names (:461-505)."""

    def __init__(self, dim: int, num_attention_heads: int, attention_head_dim: int, eps: float,
                 mod_w: torch.Tensor, mod_b: torch.Tensor, tp: int = 1):
        super().__init__()
        self.dim = dim
        self.num_attention_heads = num_attention_heads
        self.attention_head_dim = attention_head_dim
        # img_mod / txt_mod = nn.Sequential(SiLU, Linear(dim, 6 dim)); the Linear weights are views into
        # one [L, 2, 6D, D] tensor so a single small-M launch computes every block's modulation.
        self.img_mod = nn.Sequential(nn.SiLU(), _Linear(dim, 6 * dim, mod_w[0], mod_b[0]))
        self.txt_mod = nn.Sequential(nn.SiLU(), _Linear(dim, 6 * dim, mod_w[1], mod_b[1]))
        self.img_norm1 = _AdaLNHolder(dim, eps)
        self.img_norm2 = _AdaLNHolder(dim, eps)
        self.txt_norm1 = _AdaLNHolder(dim, eps)
        self.txt_norm2 = _AdaLNHolder(dim, eps)
        self.attn = _Attn(dim, attention_head_dim, eps, tp)
        self.img_mlp = _FeedForward(dim, tp)
        self.txt_mlp = _FeedForward(dim, tp)


class QwenEmbedRope(nn.Module):
    """RoPE tables of QwenEmbedRope (reference :179-285) with scale_rope=True, returned as
    (cos, sin) fp32 pairs instead of complex64.  Cached per (frame,h,w,T)."""

    def __init__(self, theta: int, axes_dim: list[int], scale_rope: bool = True):
        super().__init__()
        self.theta, self.axes_dim, self.scale_rope = theta, list(axes_dim), scale_rope
        self._cache: dict = {}

    def _angles(self, index: torch.Tensor, dim: int) -> torch.Tensor:
        return torch.outer(index.float(), 1.0 / torch.pow(self.theta, torch.arange(0, dim, 2).to(torch.float32).div(dim)))

    def tables(self, frame, height: int = None, width: int = None, txt_len: int = None):
        """One (frame, height, width) grid, or a sequence of grids as the first argument (then `height` is txt_len):
        the idx-th grid takes its frame positions from idx (reference :267), tables are concatenated (:258)."""
        if isinstance(frame, (list, tuple)) and len(frame) and isinstance(frame[0], (list, tuple)):
            grids, txt_len = tuple(tuple(g) for g in frame), (height if txt_len is None else txt_len)
        else:
            grids = ((frame, height, width),)
        key = (grids, txt_len)
        if key in self._cache:
            return self._cache[key]
        pos_index = torch.arange(4096)
        neg_index = torch.arange(4096).flip(0) * -1 - 1
        pos = [self._angles(pos_index, d) for d in self.axes_dim]
        neg = [self._angles(neg_index, d) for d in self.axes_dim]
        angs, max_vid_index = [], 0
        for idx, (frame, height, width) in enumerate(grids):
            f_frame = pos[0][idx: idx + frame].view(frame, 1, 1, -1).expand(frame, height, width, -1)
            if self.scale_rope:
                f_h = torch.cat([neg[1][-(height - height // 2):], pos[1][: height // 2]], dim=0)
                f_w = torch.cat([neg[2][-(width - width // 2):], pos[2][: width // 2]], dim=0)
                max_vid_index = max(height // 2, width // 2, max_vid_index)
            else:
                f_h, f_w = pos[1][:height], pos[2][:width]
                max_vid_index = max(height, width, max_vid_index)
            f_h = f_h.view(1, height, 1, -1).expand(frame, height, width, -1)
            f_w = f_w.view(1, 1, width, -1).expand(frame, height, width, -1)
            angs.append(torch.cat([f_frame, f_h, f_w], dim=-1).reshape(frame * height * width, -1))
        ang = torch.cat(angs, dim=0)
        txt_ang = torch.cat(pos, dim=1)[max_vid_index: max_vid_index + txt_len]
        out = (torch.cos(ang), torch.sin(ang), torch.cos(txt_ang), torch.sin(txt_ang))
        self._cache[key] = out
        return out

    def forward(self, video_fhw, txt_seq_lens, device):
        return tuple(t.to(device) for t in self.tables(_grids(video_fhw), max(txt_seq_lens)))


class _TimestepEmbedder(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.linear_1 = _Linear(256, dim)
        self.linear_2 = _Linear(dim, dim)


class QwenTimestepProjEmbeddings(nn.Module):
    def __init__(self, embedding_dim: int):
        super().__init__()
        self.timestep_embedder = _TimestepEmbedder(embedding_dim)


class _NormOut(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.linear = _Linear(dim, 2 * dim)


class QwenImageTransformer2DModel(nn.Module):
    """Drop-in for the reference class of the same name (same ctor kwargs, :625-650)."""

    def __init__(
        self,
        od_config: Any = None,
        patch_size: int = 2,
        in_channels: int = 64,
        out_channels: int | None = 16,
        num_layers: int = 60,
        attention_head_dim: int = 128,
        num_attention_heads: int = 24,
        joint_attention_dim: int = 3584,
        guidance_embeds: bool = False,
        axes_dims_rope: tuple[int, int, int] = (16, 56, 56),
        zero_cond_t: bool = False,
        use_additional_t_cond: bool = False,
        use_layer3d_rope: bool = False,
        tp_size: int | None = None,
        tp_rank: int | None = None,
        tp_group=None,
        tp_comm: str | None = None,
        sp_size: int | None = None,
        sp_rank: int | None = None,
        sp_group=None,
    ):
        super().__init__()
        if od_config is not None and getattr(od_config, "tf_model_config", None) is not None:
            # only num_layers is read from transformer/config.json, as in the reference (:652-653)
            num_layers = od_config.tf_model_config.num_layers
        if zero_cond_t or use_additional_t_cond or use_layer3d_rope or guidance_embeds:
            raise NotImplementedError("edit/layered variants (zero_cond_t, additional_t_cond, layer3d rope) are §8f 'next' items")
        if attention_head_dim != 128:
            raise ValueError("the sm_100a kernels are specialised for head_dim 128 (Qwen-Image)")
        self.parallel_config = getattr(od_config, "parallel_config", None)
        # tensor parallelism over heads / FFN (new; the reference builds these linears with disable_tp=True)
        if tp_size is None:
            tp_size = getattr(self.parallel_config, "tensor_parallel_size", 1) if self.parallel_config is not None else 1
        if tp_size > 1 and tp_rank is None:
            from vllm_omni_b200.diffusion.distributed import parallel_state as _ps
            tp_rank, tp_group = _ps.get_tensor_model_parallel_rank(), _ps.get_tp_group()
        self.tp_size, self.tp_rank, self.tp_group = int(tp_size), int(tp_rank or 0), tp_group
        # how the row-parallel partial sums are reduced: "p2p" (default) = the GEMM epilogue pushes fp32 partial tiles to
        # the row owners over NVLink peer memory, one fused kernel reduces + applies bias/gate/residual + the next AdaLN and
        # all-gathers the modulated rows (csrc/qimg_tp_p2p.cu); "nccl" = bf16 partial sums + all-reduce callback + epilogue
        # kernel (comparison baseline)
        self.tp_comm = (tp_comm or os.environ.get("QIMG_TP_COMM", "p2p")).lower()
        if self.tp_comm not in ("nccl", "p2p"):
            raise ValueError(f"tp_comm must be 'nccl' or 'p2p', got {self.tp_comm!r}")
        if num_attention_heads % self.tp_size:
            raise ValueError(f"tensor_parallel_size {self.tp_size} must divide num_attention_heads {num_attention_heads}")
        # sequence parallelism (the reference's Ulysses mode, `ulysses_degree`), fused: full weights on every rank, own rows
        # through the linears, own heads through attention; the all-to-alls are peer stores of the GEMM / attention epilogues
        if sp_size is None:
            sp_size = getattr(self.parallel_config, "ulysses_degree", 1) if self.parallel_config is not None else 1
        if sp_size > 1 and sp_rank is None:
            from vllm_omni_b200.diffusion.distributed import parallel_state as _ps
            sp_rank, sp_group = _ps.get_sequence_parallel_rank(), _ps.get_sp_group()
        self.sp_size, self.sp_rank, self.sp_group = int(sp_size), int(sp_rank or 0), sp_group
        if self.sp_size > 1 and self.tp_size > 1:
            raise ValueError("tensor parallelism and sequence parallelism cannot be combined in the native engine")
        if num_attention_heads % self.sp_size:
            raise ValueError(f"ulysses_degree {self.sp_size} must divide num_attention_heads {num_attention_heads}")
        self.in_channels = in_channels
        self.out_channels = out_channels or in_channels
        self.inner_dim = num_attention_heads * attention_head_dim
        self.num_layers = num_layers
        self.num_attention_heads = num_attention_heads
        self.joint_attention_dim = joint_attention_dim
        self.guidance_embeds = guidance_embeds
        self.zero_cond_t = zero_cond_t
        self.do_true_cfg = False
        self.eps = 1e-6
        D = self.inner_dim

        self.pos_embed = QwenEmbedRope(theta=10000, axes_dim=list(axes_dims_rope), scale_rope=True)
        self.time_text_embed = QwenTimestepProjEmbeddings(embedding_dim=D)
        self.txt_norm = _RMSNormWeight(joint_attention_dim, 1e-6)
        self.img_in = _Linear(in_channels, D)
        self.txt_in = _Linear(joint_attention_dim, D)
        # all 2*L modulation projections in one allocation: [L, {img,txt}, 6D, D]
        self._mod_all_w = torch.empty(num_layers, 2, 6 * D, D)
        self._mod_all_b = torch.empty(num_layers, 2, 6 * D)
        self.transformer_blocks = nn.ModuleList(
            [QwenImageTransformerBlock(D, num_attention_heads, attention_head_dim, self.eps, self._mod_all_w[i], self._mod_all_b[i],
                                       self.tp_size)
             for i in range(num_layers)])
        self.norm_out = _NormOut(D)
        self.proj_out = _Linear(D, patch_size * patch_size * self.out_channels)

        self._engine = None
        self._engine_keepalive = None
        self._ws: dict = {}
        self._rope_dev: dict = {}
        self._p2p_flags = None   # (local ptr, [ptr per rank]) barrier flags, exchanged once
        self._p2p_ws: dict = {}  # shape key -> (local ptr, [ptr per rank], nbytes)
        self._p2p_key = None
        self._teacache = None  # set by cache.teacache.apply_teacache_hook

    # ------------------------------------------------------------------ weights
    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        """Same contract as the reference (:804-839): q/k/v checkpoint shards are stacked into
        to_qkv / add_kv_proj in (q, k, v) order; everything else is copied by name (already-stacked
        `to_qkv` / `add_kv_proj` names are accepted too)."""
        stacked = [
            (".to_qkv.", ".to_q.", 0), (".to_qkv.", ".to_k.", 1), (".to_qkv.", ".to_v.", 2),
            (".add_kv_proj.", ".add_q_proj.", 0), (".add_kv_proj.", ".add_k_proj.", 1), (".add_kv_proj.", ".add_v_proj.", 2),
        ]
        params = dict(self.named_parameters())
        loaded: set[str] = set()
        P, r = self.tp_size, self.tp_rank

        def rows(t):  # this rank's slice of an output-feature (column-parallel) dimension
            n = t.shape[0] // P
            return t[r * n:(r + 1) * n]

        def cols(t):  # this rank's slice of an input-feature (row-parallel) dimension
            n = t.shape[1] // P
            return t[:, r * n:(r + 1) * n]

        for name, w in weights:
            for param_name, weight_name, shard in stacked:
                if weight_name not in name:
                    continue
                name = name.replace(weight_name, param_name)
                p = params[name]
                n = p.shape[0] // 3
                p.data[shard * n:(shard + 1) * n].copy_(rows(w).to(p.dtype))  # local heads of q / k / v
                break
            else:
                p = params[name]
                if P > 1:
                    if ".attn.to_qkv." in name or ".attn.add_kv_proj." in name:   # already stacked [q;k;v]
                        w = torch.cat([rows(c) for c in w.chunk(3, dim=0)], dim=0)
                    elif name.endswith(".net.0.proj.weight") or name.endswith(".net.0.proj.bias"):
                        w = rows(w)
                    elif name.endswith(("attn.to_out.0.weight", "attn.to_add_out.weight", ".net.2.weight")):
                        w = cols(w)
                if tuple(p.shape) != tuple(w.shape):
                    raise ValueError(f"shape mismatch for {name}: {tuple(p.shape)} vs {tuple(w.shape)}")
                p.data.copy_(w.to(p.dtype))
            loaded.add(name)
        self._engine = None  # pointers may have been re-materialised
        self._p2p_key = None
        return loaded

    def _apply(self, fn, *args, **kwargs):
        # keep the per-block modulation parameters views of the big tensors across .to()/.cuda()
        super()._apply(fn, *args, **kwargs)
        new_w, new_b = fn(self._mod_all_w), fn(self._mod_all_b)
        with torch.no_grad():
            for i, blk in enumerate(self.transformer_blocks):
                new_w[i, 0].copy_(blk.img_mod[1].weight.data)
                new_w[i, 1].copy_(blk.txt_mod[1].weight.data)
                new_b[i, 0].copy_(blk.img_mod[1].bias.data)
                new_b[i, 1].copy_(blk.txt_mod[1].bias.data)
                blk.img_mod[1].weight.data = new_w[i, 0]
                blk.txt_mod[1].weight.data = new_w[i, 1]
                blk.img_mod[1].bias.data = new_b[i, 0]
                blk.txt_mod[1].bias.data = new_b[i, 1]
        self._mod_all_w, self._mod_all_b = new_w, new_b
        self._engine = None
        self._ws.clear()
        self._rope_dev.clear()
        self._p2p_release()
        return self

    # ------------------------------------------------------------------ engine
    def _build_engine(self):
        dev = self.img_in.weight.device
        if dev.type != "cuda":
            raise RuntimeError("QwenImageTransformer2DModel (B200) has no CPU path: move it to an sm_100 CUDA device")
        for n, p in self.named_parameters():
            if p.dtype != torch.bfloat16 or not p.is_contiguous():
                raise RuntimeError(f"parameter {n} must be contiguous bf16 (got {p.dtype})")
        qlib.device_check()
        dims = qlib.Dims(self.num_layers, self.num_attention_heads, 128, self.in_channels, self.proj_out.out_features,
                         self.joint_attention_dim, self.eps)
        te = self.time_text_embed.timestep_embedder
        g = qlib.GlobalWeights(
            t_lin1_w=te.linear_1.weight.data_ptr(), t_lin1_b=te.linear_1.bias.data_ptr(),
            t_lin2_w=te.linear_2.weight.data_ptr(), t_lin2_b=te.linear_2.bias.data_ptr(),
            txt_norm_w=self.txt_norm.weight.data_ptr(),
            img_in_w=self.img_in.weight.data_ptr(), img_in_b=self.img_in.bias.data_ptr(),
            txt_in_w=self.txt_in.weight.data_ptr(), txt_in_b=self.txt_in.bias.data_ptr(),
            norm_out_w=self.norm_out.linear.weight.data_ptr(), norm_out_b=self.norm_out.linear.bias.data_ptr(),
            proj_out_w=self.proj_out.weight.data_ptr(), proj_out_b=self.proj_out.bias.data_ptr(),
            mod_all_w=self._mod_all_w.data_ptr(), mod_all_b=self._mod_all_b.data_ptr())
        blocks = (qlib.BlockWeights * self.num_layers)()
        for i, b in enumerate(self.transformer_blocks):
            a = b.attn
            vals = dict(
                img_mod_w=b.img_mod[1].weight, img_mod_b=b.img_mod[1].bias, txt_mod_w=b.txt_mod[1].weight, txt_mod_b=b.txt_mod[1].bias,
                to_qkv_w=a.to_qkv.weight, to_qkv_b=a.to_qkv.bias, add_kv_w=a.add_kv_proj.weight, add_kv_b=a.add_kv_proj.bias,
                norm_q=a.norm_q.weight, norm_k=a.norm_k.weight, norm_added_q=a.norm_added_q.weight, norm_added_k=a.norm_added_k.weight,
                to_out_w=a.to_out[0].weight, to_out_b=a.to_out[0].bias, to_add_out_w=a.to_add_out.weight, to_add_out_b=a.to_add_out.bias,
                img_mlp_w1=b.img_mlp.net[0].proj.weight, img_mlp_b1=b.img_mlp.net[0].proj.bias,
                img_mlp_w2=b.img_mlp.net[2].weight, img_mlp_b2=b.img_mlp.net[2].bias,
                txt_mlp_w1=b.txt_mlp.net[0].proj.weight, txt_mlp_b1=b.txt_mlp.net[0].proj.bias,
                txt_mlp_w2=b.txt_mlp.net[2].weight, txt_mlp_b2=b.txt_mlp.net[2].bias)
            for k, v in vals.items():
                setattr(blocks[i], k, v.data_ptr())
        handle = C.c_void_p()
        qlib.check(qlib.load().qimg_engine_create(C.byref(dims), C.byref(g), blocks, C.byref(handle)), "qimg_engine_create")
        self._engine = handle
        self._engine_keepalive = (dims, g, blocks)
        if self.tp_size > 1 and self.tp_comm == "p2p":
            # declare the mode (workspace layout depends on it); the peer pointers follow in _p2p_workspace
            qlib.check(qlib.load().qimg_engine_set_tp_p2p(self._engine, self.tp_size, self.tp_rank, None, None),
                       "qimg_engine_set_tp_p2p")
        if self.sp_size > 1:
            qlib.check(qlib.load().qimg_engine_set_sp_p2p(self._engine, self.sp_size, self.sp_rank, None, None),
                       "qimg_engine_set_sp_p2p")
        if self.tp_size > 1 and self.tp_comm == "nccl":
            import torch.distributed as dist

            def _allreduce(buf, count, user, stream):  # called by the engine twice per block, on its stream
                try:
                    ws, off, _ = self._ws_current
                    o = buf - ws.data_ptr()
                    dist.all_reduce(ws[o:o + count * 2].view(torch.bfloat16), group=self.tp_group)
                    return 0
                except Exception as exc:  # surfaces as "TP all-reduce callback failed"
                    self._tp_error = exc
                    return 1

            self._allreduce_cb = qlib.ALLREDUCE_FN(_allreduce)
            qlib.check(qlib.load().qimg_engine_set_tp(self._engine, self.tp_size, self._allreduce_cb, None), "qimg_engine_set_tp")

    def __del__(self):
        try:
            if self._engine is not None:
                qlib.load().qimg_engine_destroy(self._engine)
        except Exception:
            pass

    @property
    def _peer(self):
        """(size, rank, group) of the ranks that share peer-memory workspaces: the SP group or the TP group."""
        return (self.sp_size, self.sp_rank, self.sp_group) if self.sp_size > 1 else (self.tp_size, self.tp_rank, self.tp_group)

    def enable_sequence_parallel(self, sp_size: int, sp_rank: int, sp_group) -> None:
        """Switch an (unsharded, tp_size == 1) model to fused sequence parallelism over `sp_group`, or back (sp_size = 1).
        Weights stay as they are — every rank holds the full model, exactly like a data-parallel replica."""
        if self.tp_size > 1:
            raise ValueError("a tensor-parallel model cannot switch to sequence parallelism")
        if self.num_attention_heads % max(int(sp_size), 1):
            raise ValueError(f"ulysses_degree {sp_size} must divide num_attention_heads {self.num_attention_heads}")
        self._p2p_release()
        if self._engine is not None:
            qlib.load().qimg_engine_destroy(self._engine)
        self.sp_size, self.sp_rank, self.sp_group = int(sp_size), int(sp_rank), sp_group
        self._engine = None  # rebuilt (and the mode declared) on the next forward

    def _p2p_release(self):
        """Unmap the peers' buffers and free the local ones (the next forward re-registers; collective like the set-up)."""
        groups = list(self._p2p_ws.values()) + ([(self._p2p_flags[0], self._p2p_flags[1], 0)] if self._p2p_flags else [])
        _, prank, pgroup = self._peer
        for local, peers, _ in groups:
            for r, ptr in enumerate(peers):
                if r != prank:
                    qlib.ipc_close_handle(ptr)
        if groups:
            torch.cuda.synchronize()
            import torch.distributed as dist
            dist.barrier(group=pgroup)  # nobody frees while a peer still has the buffer mapped
            for local, _, _ in groups:
                qlib.p2p_free(local)
        self._p2p_ws.clear()
        self._p2p_flags = None
        self._p2p_key = None

    def _p2p_exchange(self, ptr: int) -> list[int]:
        """All ranks of the TP group swap the CUDA-IPC handle of `ptr`; returns the pointer per rank (own = local)."""
        import torch.distributed as dist

        psize, prank, pgroup = self._peer
        handles = [None] * psize
        dist.all_gather_object(handles, qlib.ipc_get_handle(ptr), group=pgroup)
        return [ptr if r == prank else qlib.ipc_open_handle(h) for r, h in enumerate(handles)]

    def _p2p_workspace(self, B: int, S_img: int, T: int):
        """Peer-memory TP: the workspace (and once, the barrier flags) is cudaMalloc'ed by the library, exported over
        CUDA IPC and registered with the engine; switching between already-registered shapes is a pointer swap."""
        key = (B, S_img, T)
        if self._p2p_flags is None:
            fl = qlib.p2p_alloc(128)
            self._p2p_flags = (fl, self._p2p_exchange(fl))
        if key not in self._p2p_ws:
            nbytes = qlib.load().qimg_engine_workspace_bytes(self._engine, B, S_img, T)
            ws = qlib.p2p_alloc(nbytes)
            self._p2p_ws[key] = (ws, self._p2p_exchange(ws), nbytes)
        ws, peers, nbytes = self._p2p_ws[key]
        if self._p2p_key != key:
            P, prank, _ = self._peer
            arr_ws = (C.c_void_p * P)(*peers)
            arr_fl = (C.c_void_p * P)(*self._p2p_flags[1])
            setter = qlib.load().qimg_engine_set_sp_p2p if self.sp_size > 1 else qlib.load().qimg_engine_set_tp_p2p
            qlib.check(setter(self._engine, P, prank, arr_ws, arr_fl), "qimg_engine_set_{sp,tp}_p2p")
            self._p2p_key = key
        return ws, nbytes

    def p2p_healthy(self) -> bool:
        """Synchronising check of the cross-GPU barrier time-out flag (peer-memory TP only)."""
        err = C.c_int(0)
        qlib.check(qlib.load().qimg_engine_p2p_error(self._engine, C.byref(err)), "qimg_engine_p2p_error")
        return err.value == 0

    def _workspace(self, B: int, S_img: int, T: int, device):
        key = (B, S_img, T, str(device))
        if key not in self._ws:
            nbytes = qlib.load().qimg_engine_workspace_bytes(self._engine, B, S_img, T)
            buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
            off = (-buf.data_ptr()) % 1024
            self._ws[key] = (buf, off, nbytes)
        return self._ws[key]

    def _rope(self, img_shapes, txt_len: int, device):
        grids = _grids(img_shapes)
        key = (grids, txt_len, str(device))
        if key not in self._rope_dev:
            ic, isn, tc, tsn = self.pos_embed.tables(grids, txt_len)
            # cos/sin are cast to the activation dtype before use, as in the reference (:403-406)
            self._rope_dev[key] = tuple(t.to(torch.bfloat16).contiguous().to(device) for t in (ic, isn, tc, tsn))
        return self._rope_dev[key], sum(f * h * w for f, h, w in grids)

    # ------------------------------------------------------------------ forward
    def forward(
        self,
        hidden_states: torch.Tensor,
        encoder_hidden_states: torch.Tensor = None,
        encoder_hidden_states_mask: torch.Tensor = None,
        timestep: torch.Tensor = None,
        img_shapes: list | None = None,
        txt_seq_lens: list[int] | None = None,
        guidance: torch.Tensor = None,
        attention_kwargs: dict[str, Any] | None = None,
        additional_t_cond=None,
        return_dict: bool = True,
        uniform_timestep: bool = False,
    ):
        """Reference contract (:692-802).  `uniform_timestep=True` (set by the native denoise loop, where
        `timestep = t.expand(B)`, pipeline_qwen_image.py:552) lets the engine compute one modulation row
        for the whole batch; results are identical."""
        if guidance is not None or additional_t_cond is not None:
            raise NotImplementedError("guidance / additional_t_cond are not part of the Qwen-Image T2I hot path")
        if self._engine is None:
            self._build_engine()
        B, S_img, C_in = hidden_states.shape
        T = encoder_hidden_states.shape[1]
        if txt_seq_lens is not None and max(txt_seq_lens) != T:
            raise ValueError(f"max(txt_seq_lens)={max(txt_seq_lens)} must equal the text length {T} (reference RoPE broadcast)")
        dev = hidden_states.device
        hs = hidden_states.to(torch.bfloat16).contiguous()
        enc = encoder_hidden_states.to(torch.bfloat16).contiguous()
        ts = timestep.to(device=dev, dtype=torch.bfloat16).reshape(-1).contiguous()
        n_t = 1 if (uniform_timestep or ts.numel() == 1) else B
        if ts.numel() not in (1, B):
            raise ValueError("timestep must have 1 or batch_size entries")
        (ic, isn, tc, tsn), s_expected = self._rope(img_shapes, T, dev)
        if s_expected != S_img:
            raise ValueError(f"img_shapes implies {s_expected} image tokens, hidden_states has {S_img}")
        peer_mode = (self.tp_size > 1 and self.tp_comm == "p2p") or self.sp_size > 1
        if peer_mode:
            ws_ptr, nbytes = self._p2p_workspace(B, S_img, T)
        else:
            buf, off, nbytes = self._workspace(B, S_img, T, dev)
            self._ws_current = (buf, off, nbytes)
            ws_ptr = buf.data_ptr() + off
        out = torch.empty((B, S_img, self.proj_out.out_features), dtype=torch.bfloat16, device=dev)

        def run_stage(stages: int):
            rc = qlib.load().qimg_engine_forward_stages(
                self._engine, stages, hs.data_ptr(), enc.data_ptr(), ts.data_ptr(), n_t, ic.data_ptr(), isn.data_ptr(),
                tc.data_ptr(), tsn.data_ptr(), B, S_img, T, out.data_ptr(), ws_ptr, nbytes, qlib.stream_ptr())
            qlib.check(rc, "qimg_engine_forward")

        if self._teacache is None:
            run_stage(qlib.STAGE_ALL)
        else:
            # step cache: the hook decides between the blocks and the cached residual (cache/teacache/hook.py)
            if peer_mode:
                raise NotImplementedError("TeaCache with a peer-memory workspace (TP p2p / sequence parallel) is not wired")
            D = self.inner_dim
            lib = qlib.load()

            def view(o, rows):
                o += off
                return buf[o: o + rows * D * 2].view(torch.bfloat16).view(rows, D)

            mod = view(lib.qimg_engine_ws_offset_mod(self._engine, B, S_img, T), B * S_img)
            x_img = view(lib.qimg_engine_ws_offset_img(self._engine, B, S_img, T), B * S_img)
            self._teacache.run(self, run_stage, mod, x_img, qlib)
        return Transformer2DModelOutput(sample=out)

    def debug_streams(self, B: int, S_img: int, T: int, device):
        """(img [B,S_img,D], txt [B,T,D]) residual streams left in the workspace by the last forward."""
        buf, off, _ = self._workspace(B, S_img, T, device)
        D = self.inner_dim
        lib = qlib.load()
        oi = lib.qimg_engine_ws_offset_img(self._engine, B, S_img, T) + off
        ot = lib.qimg_engine_ws_offset_txt(self._engine, B, S_img, T) + off
        img = buf[oi: oi + B * S_img * D * 2].view(torch.bfloat16).view(B, S_img, D)
        txt = buf[ot: ot + B * T * D * 2].view(torch.bfloat16).view(B, T, D)
        return img, txt
