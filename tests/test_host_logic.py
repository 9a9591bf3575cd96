"""CPU: host-side logic of the drop-in boundary (no kernels run)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import qwen_image_oracle as O
from vllm_omni_b200 import lib as qlib
from vllm_omni_b200 import synthetic
from vllm_omni_b200.diffusion import registry
from vllm_omni_b200.diffusion.data import DiffusionParallelConfig, OmniDiffusionConfig, TransformerConfig
from vllm_omni_b200.diffusion.distributed import parallel_state as ps
from vllm_omni_b200.diffusion.models.qwen_image import pipeline_qwen_image as P
from vllm_omni_b200.diffusion.models.qwen_image.qwen_image_transformer import QwenEmbedRope, QwenImageTransformer2DModel
from vllm_omni_b200.diffusion.request import OmniDiffusionRequest
from vllm_omni_b200.diffusion.worker.gpu_worker import shard_request

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    hdr = open(os.path.join(ROOT, "include", "qimg_b200.h")).read()
    declared = set(re.findall(r"\b(qimg_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(qlib.EXPORTED_SYMBOLS), declared ^ set(qlib.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(qlib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert qlib.load().qimg_abi_version() == 1


def test_no_cpu_fallback():
    """Without a GPU the ops must fail loudly, not silently compute on the CPU."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    x = torch.zeros(4, 256, dtype=torch.bfloat16)
    with pytest.raises((AssertionError, RuntimeError)):
        qlib.rms_norm(x, torch.ones(256, dtype=torch.bfloat16))
    torch.set_default_dtype(torch.bfloat16)
    try:
        m = QwenImageTransformer2DModel(num_layers=1, num_attention_heads=1, joint_attention_dim=64)
    finally:
        torch.set_default_dtype(torch.float32)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 4, 64, dtype=torch.bfloat16), torch.zeros(1, 2, 64, dtype=torch.bfloat16), None,
          torch.zeros(1, dtype=torch.bfloat16), [[(1, 2, 2)]], [2])


def test_product_never_imports_oracle():
    for d, _, files in os.walk(os.path.join(ROOT, "vllm_omni_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(d, f)


def test_load_weights_stacks_qkv_and_keeps_reference_names():
    torch.set_default_dtype(torch.bfloat16)
    try:
        m = QwenImageTransformer2DModel(num_layers=2, num_attention_heads=2, joint_attention_dim=256)
    finally:
        torch.set_default_dtype(torch.float32)
    names = {n for n, _ in m.named_parameters()}
    assert names == set(synthetic.param_shapes(2, num_heads=2, joint_dim=256))
    ckpt = list(synthetic.split_qkv_checkpoint_names(synthetic.synthetic_weights(2, seed=1, num_heads=2, joint_dim=256)))
    assert any(".attn.to_q.weight" in n for n, _ in ckpt) and any(".attn.add_v_proj.bias" in n for n, _ in ckpt)
    loaded = m.load_weights(ckpt)
    assert loaded == names
    sd = dict(synthetic.synthetic_weights(2, seed=1, num_heads=2, joint_dim=256))
    params = dict(m.named_parameters())
    for k, v in sd.items():
        assert torch.equal(params[k].data, v), k
    # the per-block modulation parameters alias one [L,2,6D,D] tensor (single small-M launch per forward)
    assert torch.equal(m._mod_all_w[1, 1], sd["transformer_blocks.1.txt_mod.1.weight"])
    assert params["transformer_blocks.0.img_mod.1.weight"].data_ptr() == m._mod_all_w[0, 0].data_ptr()
    # attributes the reference's hooks / pipeline touch
    for attr in ("transformer_blocks", "img_in", "txt_in", "txt_norm", "time_text_embed", "pos_embed", "norm_out",
                 "proj_out", "do_true_cfg", "in_channels", "guidance_embeds"):
        assert hasattr(m, attr)


def test_rope_tables_match_oracle():
    pe = QwenEmbedRope(10000, [16, 56, 56], True)
    for (h, w, t) in ((8, 6, 24), (5, 7, 13), (64, 64, 128)):
        got = pe.tables(1, h, w, t)
        ref = O.rope_tables(1, h, w, t)
        for a, b in zip(got, ref):
            assert torch.equal(a, b)


def test_rope_tables_multi_grid_edit_layout():
    """Edit pipelines: noisy latents + condition image(s); grid idx takes frame position idx (reference :267), text
    positions start after the LARGEST grid (:251-257); img_shapes is read like the reference does (:231-234)."""
    from vllm_omni_b200.diffusion.models.qwen_image.qwen_image_transformer import _grids
    pe = QwenEmbedRope(10000, [16, 56, 56], True)
    grids = ((1, 8, 6), (1, 4, 10))
    got, ref = pe.tables(grids, 13), O.rope_tables(list(grids), 13)
    assert got[0].shape == (8 * 6 + 4 * 10, 64) and got[2].shape == (13, 64)
    for a, b in zip(got, ref):
        assert torch.equal(a, b)
    one = pe.tables(1, 8, 6, 13)
    assert torch.equal(got[0][:48, 8:], one[0][:, 8:])      # h/w axes of the first grid unchanged
    assert not torch.equal(got[2], one[2])                  # text offset moved: max(8//2, 6//2, 4//2, 10//2) = 5 vs 4
    assert _grids([[(1, 8, 6), (1, 4, 10)]] * 3) == grids
    assert _grids([[(1, 8, 6)]] * 2) == ((1, 8, 6),) and _grids((1, 8, 6)) == ((1, 8, 6),) and _grids([(1, 8, 6)] * 2) == ((1, 8, 6),)
    with pytest.raises(NotImplementedError):
        _grids([[(1, 8, 6)], [(1, 4, 6)]])


def test_scheduler_matches_oracle_tables():
    od = OmniDiffusionConfig(tf_model_config=TransformerConfig.from_dict({"num_layers": 1}))
    sch = P.FlowMatchEulerDiscreteScheduler()
    for n, s_img in ((50, 4096), (4, 256), (28, 1024)):
        mu = P.calculate_shift(s_img, 256, 8192, 0.5, 0.9)
        sch.set_timesteps(sigmas=np.linspace(1.0, 1 / n, n), mu=mu)
        ref = O.flow_match_sigmas(n, s_img)
        assert np.array_equal(sch.sigmas.numpy(), ref)
        assert torch.equal(sch.timesteps, torch.from_numpy(ref[:-1]) * 1000)


def test_registry_and_config():
    assert registry.DiffusionModelRegistry._try_load_model_cls("QwenImagePipeline") is P.QwenImagePipeline
    assert registry.DiffusionModelRegistry._try_load_model_cls("Nope") is None
    edit = registry.DiffusionModelRegistry._try_load_model_cls("QwenImageEditPipeline")
    assert issubclass(edit, P.QwenImagePipeline) and list(__import__("inspect").signature(edit.diffuse).parameters)[6] == "image_latents"
    assert registry.get_diffusion_post_process_func(OmniDiffusionConfig(model_class_name="QwenImageEditPipeline")) is not None
    od = OmniDiffusionConfig(parallel_config={"data_parallel_size": 4, "tensor_parallel_size": 2})
    assert od.num_gpus == 8 and od.parallel_config.world_size == 8
    with pytest.raises(ValueError):
        DiffusionParallelConfig(ulysses_degree=2, sequence_parallel_size=3)
    with pytest.raises(ValueError):
        registry.initialize_model(OmniDiffusionConfig(model_class_name="Nope"))
    assert registry.get_diffusion_post_process_func(od) is not None


def test_attention_selector_env(monkeypatch):
    from vllm_omni_b200.diffusion.attention import selector
    selector.get_attn_backend.cache_clear()
    monkeypatch.setenv("DIFFUSION_ATTENTION_BACKEND", "b200_fmha")
    assert selector.get_attn_backend(-1).get_name() == "B200_FMHA"
    selector.get_attn_backend.cache_clear()
    monkeypatch.setenv("DIFFUSION_ATTENTION_BACKEND", "FLASH_ATTN")
    with pytest.raises(ValueError):
        selector.get_attn_backend(-1)
    selector.get_attn_backend.cache_clear()


def test_shard_request_covers_every_unit_once():
    pe = torch.arange(3 * 5 * 8, dtype=torch.float32).view(3, 5, 8)
    req = OmniDiffusionRequest(prompt_embeds=pe, num_outputs_per_prompt=3, seed=7, height=256, width=256)
    for world in (1, 2, 4, 8, 16):
        seen, counts_ref = [], None
        for r in range(world):
            local, counts = shard_request(req, r, world)
            counts_ref = counts_ref or counts
            assert counts == counts_ref and sum(counts) == 9
            if local is None:
                assert counts[r] == 0
                continue
            assert local.prompt_embeds.shape[0] == counts[r] and local.num_outputs_per_prompt == 1
            lo, _ = ps.shard_range(9, r, world)
            for j in range(counts[r]):
                assert torch.equal(local.prompt_embeds[j], pe[(lo + j) // 3])
                seen.append(lo + j)
        assert seen == list(range(9))


def test_flops_formula_matches_survey():
    from vllm_omni_b200.flops import flops_per_forward
    assert abs(flops_per_forward(60, 4096, 128) / 7.056e13 - 1) < 2e-3
    assert abs(flops_per_forward(60, 16384, 128) / 4.254e14 - 1) < 2e-3
    assert abs(flops_per_forward(60, 4096, 128) - O.flops_per_forward(O.DiTDims(), 4096, 128)) < 1


def test_tensor_parallel_weight_sharding():
    """TP (new vs the reference, whose Qwen-Image linears are disable_tp=True): q|k|v rows of the local heads,
    out-projection / MLP-down columns, MLP-up rows; row-parallel biases stay whole."""
    H, joint, P = 4, 256, 2
    D = H * 128
    full = dict(synthetic.synthetic_weights(1, seed=1, num_heads=H, joint_dim=joint))
    torch.set_default_dtype(torch.bfloat16)
    try:
        ms = [QwenImageTransformer2DModel(num_layers=1, num_attention_heads=H, joint_attention_dim=joint, tp_size=P, tp_rank=r)
              for r in range(P)]
    finally:
        torch.set_default_dtype(torch.float32)
    ms[0].load_weights(synthetic.split_qkv_checkpoint_names(full.items()))   # q/k/v-separate checkpoint names
    ms[1].load_weights(full.items())                                         # already stacked names
    ps_ = [dict(m.named_parameters()) for m in ms]
    b = "transformer_blocks.0."
    qf, kf, vf = full[b + "attn.to_qkv.weight"].chunk(3, 0)
    for r in range(P):
        sl = slice(r * D // P, (r + 1) * D // P)
        assert torch.equal(ps_[r][b + "attn.to_qkv.weight"], torch.cat([qf[sl], kf[sl], vf[sl]]))
        assert ps_[r][b + "attn.to_out.0.bias"].shape == (D,)
        assert ps_[r][b + "img_mlp.net.2.bias"].shape == (D,)
    for k, dim in ((b + "attn.to_out.0.weight", 1), (b + "attn.to_add_out.weight", 1), (b + "img_mlp.net.0.proj.weight", 0),
                   (b + "txt_mlp.net.0.proj.bias", 0), (b + "txt_mlp.net.2.weight", 1)):
        assert torch.equal(torch.cat([ps_[0][k], ps_[1][k]], dim), full[k]), k
    assert torch.equal(ps_[0][b + "img_mod.1.weight"], full[b + "img_mod.1.weight"])  # replicated
    with pytest.raises(ValueError):
        QwenImageTransformer2DModel(num_layers=1, num_attention_heads=3, joint_attention_dim=64, tp_size=2, tp_rank=0)


def test_teacache_config_selector_and_decision_logic():
    """Step-cache host logic (reference cache/teacache/{config,hook,backend}.py, cache/selector.py): config validation,
    backend selection, and the accumulate / threshold / reset decision with the reference's bf16 roundings."""
    from vllm_omni_b200.diffusion.cache import get_cache_backend
    from vllm_omni_b200.diffusion.cache.teacache import TeaCacheBackend, TeaCacheConfig, TeaCacheHook, TeaCacheState
    from vllm_omni_b200.diffusion.data import DiffusionCacheConfig
    assert TeaCacheConfig().coefficients == [-450.0, 280.0, -45.0, 3.2, -0.02] and TeaCacheConfig().rel_l1_thresh == 0.2
    with pytest.raises(ValueError):
        TeaCacheConfig(rel_l1_thresh=0.0)
    with pytest.raises(ValueError):
        TeaCacheConfig(coefficients=[1.0, 2.0])
    with pytest.raises(KeyError):
        TeaCacheConfig(transformer_type="FluxTransformer2DModel")
    assert get_cache_backend(None, None) is None and get_cache_backend("none", {}) is None
    be = get_cache_backend("tea_cache", {"rel_l1_thresh": 0.3, "Fn_compute_blocks": 2})
    assert isinstance(be, TeaCacheBackend) and be.config.rel_l1_thresh == 0.3 and not be.is_enabled()
    with pytest.raises(ValueError):
        get_cache_backend("cache_dit", DiffusionCacheConfig())
    with pytest.raises(ValueError):
        get_cache_backend("deep_cache", {})

    class FakeTransformer:
        _teacache = None
        do_true_cfg = False

    class QwenImageTransformer2DModel(FakeTransformer):
        pass

    class Pipe:
        transformer = QwenImageTransformer2DModel()

    be.enable(Pipe)
    hook = Pipe.transformer._teacache
    assert be.is_enabled() and isinstance(hook, TeaCacheHook) and hook.config.rel_l1_thresh == 0.3

    class FakeLib:  # the reduction kernel's contract, on CPU tensors
        @staticmethod
        def rel_l1_sums(a, b, sums):
            sums[0] = (a - b).abs().float().sum()
            sums[1] = b.abs().float().sum()

    g = torch.Generator().manual_seed(0)
    hook = TeaCacheHook(TeaCacheConfig(rel_l1_thresh=0.05, coefficients=[0.0, 0.0, 0.0, 1.0, 0.0]))
    st = TeaCacheState()
    prev = torch.randn(4, 64, generator=g).bfloat16()
    assert hook._should_compute(st, prev, FakeLib) == (True, pytest.approx(float("nan"), nan_ok=True))  # first step
    st.cnt, st.has_mod, st.previous_modulated_input = 1, True, prev
    acc, seen = 0.0, []
    for k in range(6):
        cur = (prev.float() * (1 + 0.02 * (k + 1))).bfloat16()
        want = ((cur - prev).abs().mean() / (prev.abs().mean() + 1e-8)).item()  # reference expression, hook.py:198-205
        compute, rel = hook._should_compute(st, cur, FakeLib)
        assert rel == want
        acc += abs(want)
        assert compute == (acc >= 0.05)
        if compute:
            acc = 0.0
        seen.append(compute)
    assert any(seen) and not all(seen)
    hook.reset_state()
    assert hook._forward_cnt == 0 and hook.decisions == []


def test_edit_pipeline_condition_latents_host_logic():
    """QwenImageEditPipeline._condition_latents: grid bookkeeping and the batch-replication rule of the reference's
    prepare_latents (pipeline_qwen_image_edit.py:508-517), without touching the GPU."""
    from vllm_omni_b200.diffusion.models.qwen_image.pipeline_qwen_image_edit import QwenImageEditPipeline
    from vllm_omni_b200.diffusion.request import OmniDiffusionRequest
    f = QwenImageEditPipeline._condition_latents
    shapes = [[(1, 8, 6)]] * 4
    assert f(None, OmniDiffusionRequest(), 4, shapes) == (None, shapes)            # no image: text-to-image path
    il = torch.zeros(2, 24, 64, dtype=torch.bfloat16)
    out, sh = f(None, OmniDiffusionRequest(extra={"image_latents": il, "image_latent_grid": (4, 6)}), 4, shapes)
    assert out.shape == (4, 24, 64) and sh == [[(1, 8, 6), (1, 4, 6)]] * 4
    with pytest.raises(ValueError):   # 3 condition images cannot be spread over 4 prompts
        f(None, OmniDiffusionRequest(extra={"image_latents": torch.zeros(3, 24, 64), "image_latent_grid": (4, 6)}), 4, shapes)
    with pytest.raises(ValueError):   # grid does not describe the latents
        f(None, OmniDiffusionRequest(extra={"image_latents": il, "image_latent_grid": (5, 6)}), 4, shapes)
    with pytest.raises(ValueError):
        f(None, OmniDiffusionRequest(extra={"image_latents": il}), 4, shapes)


def test_c_abi_argument_checks_without_gpu():
    """Entry points validate their arguments before any CUDA call: exercised on the CPU-only build host (no compute)."""
    import ctypes as C
    lib = qlib.load()
    dims = qlib.Dims(2, 24, 128, 64, 64, 3584, 1e-6)
    g, blocks, h = qlib.GlobalWeights(), (qlib.BlockWeights * 2)(), C.c_void_p()
    assert lib.qimg_engine_create(C.byref(dims), C.byref(g), blocks, C.byref(h)) == 0
    try:
        two = (C.c_void_p * 2)(1024, 2048)
        assert lib.qimg_engine_set_tp_p2p(h, 3, 0, two, two) != 0 and b"tp_size" in lib.qimg_last_error()
        assert lib.qimg_engine_set_tp_p2p(h, 2, 2, two, two) != 0 and b"rank" in lib.qimg_last_error()
        assert lib.qimg_engine_set_tp_p2p(h, 2, 0, (C.c_void_p * 2)(1024, None), two) != 0
        assert lib.qimg_engine_set_tp(h, 5, qlib.ALLREDUCE_FN(lambda *a: 0), None) != 0   # 5 does not divide 24 heads
        assert lib.qimg_engine_set_tp(h, 2, C.cast(None, qlib.ALLREDUCE_FN), None) != 0   # TP needs a callback
        assert lib.qimg_engine_forward_stages(h, 0, *([None] * 3), 1, *([None] * 4), 1, 64, 8, None, None, 0, None) != 0
        assert b"stage" in lib.qimg_last_error()
        assert lib.qimg_engine_workspace_bytes(h, 1, 4096, 128) > 4096 * 3072 * 2 * 4
        assert lib.qimg_engine_ws_offset_mod(h, 1, 4096, 128) > lib.qimg_engine_ws_offset_txt(h, 1, 4096, 128) > 0
    finally:
        lib.qimg_engine_destroy(h)
    bad = qlib.Dims(2, 24, 64, 64, 64, 3584, 1e-6)
    assert lib.qimg_engine_create(C.byref(bad), C.byref(g), blocks, C.byref(h)) != 0 and b"head_dim" in lib.qimg_last_error()
    assert lib.qimg_set_fmha_mode(7) != 0 and lib.qimg_set_fmha_mode(0) != 0 and lib.qimg_set_gemm_mode(3) != 0
    assert lib.qimg_set_gemm_mode(2) == 0 and lib.qimg_get_gemm_mode() == 2 and lib.qimg_set_fmha_single_tile(2) != 0
    assert lib.qimg_engine_set_sp_p2p(None, 2, 0, None, None) != 0
    assert lib.qimg_rel_l1_sums(None, None, 12, None, None) != 0 and b"multiple of 8" in lib.qimg_last_error()


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the native one): one JSON line with the contract's
    keys; run here at a reduced resolution so that it takes seconds."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--res", "256", "--txt-len", "32"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["value"] > 0 and line["vs_baseline"] is None
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and 1 <= cb["cores"] <= (os.cpu_count() or 1) and cb["value"] == line["value"] and "probe" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_worker_proc_message_protocol(monkeypatch):
    """The runner protocol of the reference WorkerProc (gpu_worker.py:143-314): rpc dicts (output_rank /
    exec_all_ranks), request lists -> execute_model, exceptions -> DiffusionOutput(error=...), shutdown, ready handshake.
    The GPU worker itself is replaced by a recorder: this is host logic only."""
    from vllm_omni_b200.diffusion.data import DiffusionOutput
    from vllm_omni_b200.diffusion.worker import gpu_worker as gw

    class Recorder:
        def __init__(self):
            self.calls, self.shut = [], False

        def generate(self, reqs):
            self.calls.append(("generate", reqs))
            return DiffusionOutput(output=torch.zeros(1))

        def execute_model(self, reqs, od):
            self.calls.append(("execute_model", reqs))
            if reqs == ["boom"]:
                raise RuntimeError("kernel launch failed")
            return DiffusionOutput(output=torch.ones(1))

        def ping(self, x, y=0):
            return x + y

        def shutdown(self):
            self.shut = True

    class Queue:
        def __init__(self, items=()):
            self.items = list(items)
            self.out = []

        def dequeue(self, indefinite=True):
            return self.items.pop(0)

        def enqueue(self, x):
            self.out.append(x)

    class Pipe:
        sent = None

        def send(self, m):
            Pipe.sent = m

    monkeypatch.setattr(gw.WorkerProc, "_create_worker", lambda self, gpu_id, od: Recorder())
    od = OmniDiffusionConfig()
    msgs = [
        {"type": "rpc", "method": "ping", "args": (2,), "kwargs": {"y": 3}, "output_rank": 0},
        {"type": "rpc", "method": "ping", "args": (1,), "output_rank": 1},                  # not for this rank: ignored
        {"type": "rpc", "method": "ping", "args": (5,), "output_rank": 1, "exec_all_ranks": True},  # executed, no reply
        {"type": "rpc", "method": "nope", "output_rank": None},                              # error reply, loop survives
        [],                                                                                  # empty message: skipped
        ["req"],
        ["boom"],
        {"type": "shutdown"},
    ]
    res = Queue()
    gw.WorkerProc.worker_main(0, od, Pipe(), Queue(msgs), result_queue=res)
    assert Pipe.sent == {"status": "ready", "result_handle": None}
    assert res.out[0] == 5
    assert isinstance(res.out[1], dict) and res.out[1]["status"] == "error" and "nope" in res.out[1]["error"]
    assert isinstance(res.out[2], DiffusionOutput) and res.out[2].error is None
    assert isinstance(res.out[3], DiffusionOutput) and res.out[3].error == "kernel launch failed"
    assert len(res.out) == 4
    # a non-zero rank never owns the result queue (reference :165-170)
    p1 = gw.WorkerProc(od, gpu_id=1, broadcast_queue=Queue([{"type": "shutdown"}]), result_queue=Queue())
    assert p1.result_mq is None
    p1.worker_busy_loop()
    assert p1.worker.shut
    assert gw.get_diffusion_worker_class() is gw.WorkerProc


def test_prompt_encode_glue_matches_reference_golden(golden_dir):
    """`QwenImagePipeline.encode_prompt` / `_get_qwen_prompt_embeds` (chat template, 34-token preamble drop, masked
    extraction, right padding, per-image repeat, truncation) against the outputs of the UNMODIFIED reference functions
    (pipeline_qwen_image.py:348-434) on the same deterministic tokenizer / encoder stand-ins — bit-equal."""
    import os

    from oracle import prompt_stubs
    from vllm_omni_b200.diffusion.data import OmniDiffusionConfig, TransformerConfig
    from vllm_omni_b200.diffusion.models.qwen_image.pipeline_qwen_image import QwenImagePipeline
    fx = torch.load(os.path.join(golden_dir, "prompt_encode.pt"))
    od = OmniDiffusionConfig(tf_model_config=TransformerConfig.from_dict({"num_layers": 1}))
    pipe = QwenImagePipeline(od_config=od, text_encoder=prompt_stubs.StubTextEncoder(), tokenizer=prompt_stubs.StubTokenizer(),
                             transformer_kwargs=dict(num_attention_heads=2, joint_attention_dim=48))
    e, m = pipe.encode_prompt(fx["prompts"], num_images_per_prompt=2)
    assert torch.equal(e, fx["embeds_x2"]) and torch.equal(m, fx["mask_x2"])
    e, m = pipe.encode_prompt(fx["prompts"][1], num_images_per_prompt=1, max_sequence_length=40)
    assert torch.equal(e, fx["embeds_trunc40"]) and torch.equal(m, fx["mask_trunc40"])
    # pre-computed embeddings pass through untouched apart from the repeat (the stage-overlap path)
    e2, m2 = pipe.encode_prompt(None, num_images_per_prompt=2, prompt_embeds=fx["embeds_trunc40"], prompt_embeds_mask=fx["mask_trunc40"])
    assert e2.shape[0] == 2 and torch.equal(e2[0], fx["embeds_trunc40"][0]) and torch.equal(m2[1], fx["mask_trunc40"][0])


def test_two_stage_pipeline_overlaps_and_matches_serial():
    """Stage overlap (BASELINE configs[3]): the encoder stage works on request i+1 while the DiT stage denoises request i;
    results equal the serial order; the connector keeps the reference's put / get contract (adapter.py:15-170)."""
    import time

    from oracle import prompt_stubs
    from vllm_omni_b200.diffusion.data import DiffusionOutput
    from vllm_omni_b200.diffusion.request import OmniDiffusionRequest
    from vllm_omni_b200.distributed.device_connector import DeviceTensorConnector
    from vllm_omni_b200.entrypoints.stage_pipeline import TwoStagePipeline
    tok, enc = prompt_stubs.StubTokenizer(), prompt_stubs.StubTextEncoder()

    def encode(prompt):
        time.sleep(0.05)
        b = tok([prompt] if isinstance(prompt, str) else prompt)
        return enc(b.input_ids, b.attention_mask).hidden_states[-1], b.attention_mask

    def denoise(req):
        time.sleep(0.15)
        assert req.prompt_embeds is not None and req.prompt_attention_mask is not None
        return DiffusionOutput(output=req.prompt_embeds.float().mean(dim=(1, 2)))

    c = DeviceTensorConnector()
    ok, size, meta = c.put("0", "1", "x", {"prompt": (torch.ones(2, 3), torch.ones(2, 3))})
    assert ok and size == 48 and meta and c.get("0", "1", "x")["prompt"][0].shape == (2, 3)
    with pytest.raises(TimeoutError):
        c.get("0", "1", "missing", timeout=0.05)

    reqs = [OmniDiffusionRequest(prompt=p, negative_prompt="blurry" if i == 1 else None, true_cfg_scale=4.0 if i == 1 else 1.0)
            for i, p in enumerate(["a fox", "a city at night", "tea", "an old map"])]
    p = TwoStagePipeline(encode, denoise)
    t0 = time.perf_counter()
    serial = p.generate(reqs, overlap=False)
    t_serial = time.perf_counter() - t0
    p.timeline.clear()
    t0 = time.perf_counter()
    par = p.generate(reqs, overlap=True)
    t_par = time.perf_counter() - t0
    assert all(torch.equal(a.output, b.output) for a, b in zip(serial, par))
    enc_spans = {rid: (a, b) for s, rid, a, b in p.timeline if s == "encode"}
    den_spans = {rid: (a, b) for s, rid, a, b in p.timeline if s == "denoise"}
    assert enc_spans["req-2"][0] < den_spans["req-1"][1], "request 2 must be encoded while request 1 is being denoised"
    assert t_par < t_serial - 0.1   # 4 x (0.05 [+0.05] + 0.15) serial vs ~0.1 + 4 x 0.15 overlapped


def test_vae_weight_packing_is_the_implicit_gemm_layout():
    """The native convolution computes out[pixel, co] = sum_{tap, ci} x[pixel + (tap / 3 - 1, tap % 3 - 1), ci] * w[co, tap * Cp + ci]
    (include/qimg_b200.h).  Check on the CPU that `_pack3x3` of a reference Conv3d weight — last temporal tap, channels padded —
    gives exactly that matrix: an explicit im2col GEMM with it equals F.conv2d with weight[:, :, 2]."""
    import torch.nn.functional as F
    from vllm_omni_b200.diffusion.models.qwen_image.vae_decoder import B200VaeDecoder, _pack3x3
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 16, 5, 7, generator=g)
    w3d = torch.randn(8, 16, 3, 3, 3, generator=g)
    packed = _pack3x3(w3d, cin_pad=32)
    assert packed.shape == (8, 9 * 32)
    xp = F.pad(x, (1, 1, 1, 1)).permute(0, 2, 3, 1)[0]  # [H + 2, W + 2, C]
    cols = []
    for tap in range(9):
        dy, dx = tap // 3 - 1, tap % 3 - 1
        patch = xp[1 + dy:1 + dy + 5, 1 + dx:1 + dx + 7]  # x[pixel + (dy, dx)]
        cols.append(F.pad(patch, (0, 16)))                # channels padded to 32 with zeros
    A = torch.cat(cols, dim=-1).reshape(35, 9 * 32)
    want = F.conv2d(x, w3d[:, :, 2], padding=1)[0].permute(1, 2, 0).reshape(35, 8)
    assert (A @ packed.T - want).abs().max().item() < 1e-4
    # the decoder object: reference checkpoint keys in, packed fp32 matrices out; time_conv (never run for one frame) dropped
    from vllm_omni_b200 import synthetic
    dec = B200VaeDecoder(synthetic.synthetic_vae_decoder_weights(seed=1), device="cpu")
    assert dec.w["decoder.conv_in.weight"].shape == (384, 9 * 32) and dec.w["decoder.conv_out.weight"].shape == (3, 3, 3, 96)
    assert dec.w["decoder.up_blocks.1.resnets.0.conv_shortcut.weight"].shape == (384, 192)
    assert not any("time_conv" in k for k in dec.w) and dec.num_up_blocks == 4 and dec.num_res == 3
    assert dec.dtype == torch.float32 and dec.config.z_dim == 16 and len(dec.config.latents_mean) == 16
    with pytest.raises(RuntimeError, match="no CPU path"):
        dec.decode(torch.zeros(1, 16, 1, 8, 16))
    with pytest.raises(ValueError):
        B200VaeDecoder({"encoder.conv_in.weight": torch.zeros(1)}, device="cpu")


def test_post_process_func_is_the_reference_image_arithmetic():
    """(x / 2 + 0.5).clamp(0, 1) * 255, rounded half to even, NHWC uint8 -> PIL (VaeImageProcessor.postprocess behind the
    reference's get_qwen_image_post_process_func, pipeline_qwen_image.py:40-60); latents pass through untouched."""
    import numpy as np
    from vllm_omni_b200.diffusion.models.qwen_image.pipeline_qwen_image import get_qwen_image_post_process_func
    post = get_qwen_image_post_process_func(None)
    lat = torch.zeros(2, 16, 64)
    assert post(lat) is lat
    img = torch.tensor([-1.5, -1.0, -0.5, 0.0, 1 / 255, 0.5, 1.0, 2.0]).view(1, 1, 2, 4).expand(1, 3, 2, 4).contiguous()
    out = post(img)
    assert len(out) == 1 and out[0].size == (4, 2) and out[0].mode == "RGB"
    want = np.array([0, 0, 64, 128, 128, 191, 255, 255], dtype=np.uint8).reshape(2, 4)  # 63.75 -> 64, 127.5 -> 128 (even), 191.25
    assert (np.asarray(out[0])[..., 0] == want).all()
    u8 = torch.arange(24, dtype=torch.uint8).view(1, 2, 4, 3)
    assert (np.asarray(post(u8)[0]) == u8[0].numpy()).all()


def test_edit_pipeline_encodes_pixel_images_like_the_reference_glue():
    """`_encode_vae_image` (reference pipeline_qwen_image_edit.py:458-480: posterior mode, (x - mean) / std) followed by
    `_pack_latents` (:519-522), with a stand-in VAE on the CPU: the packed condition latents and their RoPE grid."""
    from types import SimpleNamespace
    from vllm_omni_b200.diffusion.models.qwen_image.pipeline_qwen_image_edit import QwenImageEditPipeline, QwenImageEditPlusPipeline
    from vllm_omni_b200.diffusion.models.qwen_image.vae_decoder import LATENTS_MEAN, LATENTS_STD, _DiagonalGaussian
    from vllm_omni_b200.diffusion.request import OmniDiffusionRequest

    class StubVae:
        config = SimpleNamespace(z_dim=16, latents_mean=LATENTS_MEAN, latents_std=LATENTS_STD)

        def encode(self, x):  # parameters = a fixed function of the image: [B, 32, 1, H/8, W/8]
            b, _, _, h, w = x.shape
            p = torch.nn.functional.avg_pool2d(x[:, :, 0], 8).mean(1, keepdim=True).repeat(1, 32, 1, 1)
            p = p * torch.linspace(0.5, 2.0, 32).view(1, 32, 1, 1)
            return SimpleNamespace(latent_dist=_DiagonalGaussian(p.unsqueeze(2)))

    pipe = SimpleNamespace(vae=StubVae(), device=torch.device("cpu"), _pack_latents=QwenImageEditPipeline._pack_latents)
    pipe._encode_vae_image = lambda im: QwenImageEditPipeline._encode_vae_image(pipe, im)
    img = torch.rand(2, 3, 64, 96) * 2 - 1
    il, grid = pipe._encode_vae_image(img)
    assert il.shape == (2, 4 * 6, 64) and il.dtype == torch.bfloat16 and grid == (4, 6)
    mode = StubVae().encode(img.unsqueeze(2)).latent_dist.mode()
    want = (mode - torch.tensor(LATENTS_MEAN).view(1, 16, 1, 1, 1)) / torch.tensor(LATENTS_STD).view(1, 16, 1, 1, 1)
    want = want.view(2, 16, 4, 2, 6, 2).permute(0, 2, 4, 1, 3, 5).reshape(2, 24, 64).bfloat16()
    assert torch.equal(il, want)
    shapes = [[(1, 8, 6)]] * 2
    out, sh = QwenImageEditPipeline._condition_latents(pipe, OmniDiffusionRequest(extra={"image": img}), 2, shapes)
    assert torch.equal(out, want) and sh == [[(1, 8, 6), (1, 4, 6)]] * 2
    out, sh = QwenImageEditPlusPipeline._condition_latents(pipe, OmniDiffusionRequest(extra={"image": [img, img[:, :, :32]]}), 2, shapes)
    assert out.shape == (2, 24 + 12, 64) and sh == [[(1, 8, 6), (1, 4, 6), (1, 2, 6)]] * 2 and torch.equal(out[:, :24], want)
    with pytest.raises(ValueError, match="can encode"):
        QwenImageEditPipeline._encode_vae_image(SimpleNamespace(vae=None), img)


def test_vae_c_abi_argument_checks_without_gpu():
    """The VAE entry points reject bad shapes / strides / pointers before any CUDA call (CPU-only build host, no compute)."""
    lib = qlib.load()
    conv = lib.qimg_conv2d_nhwc_tf32
    ok = dict(x=4096, ldx=96, w=8192, ldw=864, bias=None, res=None, ldr=0, out=12288, ldo=96, N=1, H=16, W=32, Cin=96, Cout=96, taps=9)

    def call(**kw):
        a = {**ok, **kw}
        return conv(a["x"], a["ldx"], a["w"], a["ldw"], a["bias"], a["res"], a["ldr"], a["out"], a["ldo"], a["N"], a["H"], a["W"],
                    a["Cin"], a["Cout"], a["taps"], None)

    assert call(taps=4) != 0 and b"taps" in lib.qimg_last_error()
    assert call(H=4) != 0 and b"H >= 8" in lib.qimg_last_error()
    assert call(Cin=80) != 0 and b"Cin" in lib.qimg_last_error()             # 3x3 needs whole 32-channel K blocks
    assert call(Cin=16, ldw=144) != 0 and b"Cin" in lib.qimg_last_error()
    assert call(Cout=98) != 0 and b"Cout" in lib.qimg_last_error()
    assert call(ldw=800) != 0 and b"leading dimensions" in lib.qimg_last_error()
    assert call(ldx=95) != 0 and call(res=4096, ldr=64) != 0
    assert call(x=4100) != 0 and b"aligned" in lib.qimg_last_error()
    assert call(x=None) != 0 and b"null" in lib.qimg_last_error()
    down = lib.qimg_conv2d_down2_nhwc_tf32
    assert down(4096, 96, 8192, 864, None, 12288, 96, 1, 17, 32, 96, 96, None) != 0 and b"even" in lib.qimg_last_error()
    assert down(4096, 96, 8192, 864, None, 12288, 96, 1, 16, 32, 48, 96, None) != 0 and b"Cin" in lib.qimg_last_error()
    assert lib.qimg_vae_rms_act(4096, 4096, 4096, 10, 100, 1, None) != 0 and b"96, 192 or 384" in lib.qimg_last_error()
    assert lib.qimg_vae_post_quant(4096, 4096, 4096, 4096, 1, 8, 16, 8, None) != 0 and b"z_dim" in lib.qimg_last_error()
    assert lib.qimg_vae_conv_out(4096, 4096, 4096, None, None, 1, 8, 16, 96, None) != 0        # neither output requested
    assert lib.qimg_vae_conv_out(4096, 4096, 4096, 4096, None, 1, 8, 16, 64, None) != 0 and b"96" in lib.qimg_last_error()
    assert lib.qimg_vae_upsample2x(4096, 4096, 1, 8, 16, 6, None) != 0
    assert lib.qimg_vae_softmax_rows(4096, 4, 100, 50, 1.0, None) != 0                         # ld < cols
    assert lib.qimg_vae_image_to_nhwc(4096, 4096, 1, 40, 8, 16, None) != 0
    assert lib.qimg_set_vae_conv_variant(2) != 0 and lib.qimg_set_vae_conv_variant(0) == 0
    assert lib.qimg_ln_modulate_indexed(4096, 4096, 4096, 4096, 8, 256, 4, 768, 1e-6, None, 2, None) != 0 and b"index" in lib.qimg_last_error()
    assert lib.qimg_select_rows(4096, 768, 4096, 4096, 8, 250, 4, 2, None) != 0                # D % 8
