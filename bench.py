#!/usr/bin/env python
"""bench.py — images/sec of the 50-step Qwen-Image DiT denoise (BASELINE.json metric) on N B200s.

A "step" is ONE pass of the hot path over one batch: the full `num_inference_steps`(=50)-step denoise (60-layer DiT
forward per timestep + fused scheduler/CFG step) of `--batch` (=4) synthetic 1024x1024 images per GPU (BASELINE.json
configs[1]).  N GPUs = data parallel over images (weak scaling, no collective on the data path).

  value    whole-job images/sec with inputs resident in HBM (CUDA events, max over ranks, per-launch profiling OFF)
  e2e      the same through the RUNNER: `GPUWorker.execute_model([request])` with ONE request of 4 N images in pinned host
           memory — shard over the DP ranks, H2D, denoise, gather on rank 0, D2H of the result — all inside the timed region
  roofline tcgen05 GEMM family (dominant kernel): algorithmic FLOPs / summed per-launch CUDA-event time, measured live in
           ONE extra profiled step after the timed region, against MEASURED_PEAKS.json (sustained bf16 figure); the
           attention kernel and the whole step are reported beside it
  gpu_eager_baseline (N = 1)  SURVEY §8d "GPU reference timing": the reference's eager op sequence (cuBLAS + ATen + the
           SDPA / flash-attn attention backend) on the same weights and inputs, >= 3 warm-up + >= 5 timed forwards
  sp, tp, cfg_parallel (N > 1)    BASELINE configs[2]: one forward split over N GPUs at B = 1 and B = 4 — fused sequence
           parallelism (Ulysses with the all-to-alls as peer stores of the GEMM / attention epilogues; bit-identical to
           one GPU), tensor parallelism over heads / FFN (fused GEMM + peer-memory reduce-scatter, csrc/qimg_tp_p2p.cu) —
           and CFG parallel, measured after the DP region
  cpu_baseline / --impl reference   the reference's CPU torch path (oracle port, see oracle/) on this host's cores
  --sweep  BASELINE configs[4]: batch x resolution grid (profiles/r02_sweep.json)

Usage: python bench.py --gpus N --steps K --warmup W   (under torchrun for N > 1)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec Qwen-Image 1024px 50-step DiT"
UNIT = "images/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=4, help="images per GPU")
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--txt-len", type=int, default=128)
    ap.add_argument("--num-inference-steps", type=int, default=50)
    ap.add_argument("--layers", type=int, default=60)
    ap.add_argument("--cfg", action="store_true", help="true-CFG on (2 forwards per timestep, reference default)")
    ap.add_argument("--cache", default="none", choices=["none", "tea_cache"], help="step cache (reference cache_backend)")
    ap.add_argument("--rel-l1-thresh", type=float, default=0.2)
    ap.add_argument("--graph", action="store_true", help="CUDA-graph the per-timestep launch sequence (B=1 latency)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the eager-GPU baseline (N=1) / TP + CFG legs (N>1)")
    ap.add_argument("--sweep", action="store_true", help="BASELINE configs[4] batch x resolution sweep (one JSON object)")
    ap.add_argument("--sweep-batches", default="1,2,4,8,16,32")
    ap.add_argument("--sweep-res", default="512,1024,2048")
    ap.add_argument("--sweep-timesteps", type=int, default=3, help="timed denoise timesteps per sweep cell (cost is uniform per timestep)")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        sm, pw, mx, reasons = [], [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1]); pw.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "power_w": statistics.median(pw) if pw else None, "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# CPU baseline: the reference's torch path (oracle port) on the host cores, bounded sample
# --------------------------------------------------------------------------------------------
_CPU_THREADS = None


def cpu_threads() -> int:
    """Thread count for the CPU arm: the fastest of a few candidates on a small probe (one full-width block at 512 px).
    All host cores is not always the reference's best case — on the 128-core GPU boxes torch's bf16 CPU GEMMs were slower
    with 128 threads than a quarter of them — and the baseline should be the reference at its best (both counts are
    printed in `sample`)."""
    global _CPU_THREADS
    if _CPU_THREADS is not None:
        return _CPU_THREADS
    from oracle import qwen_image_oracle as O
    from vllm_omni_b200 import synthetic
    cores = os.cpu_count() or 1
    cands = sorted({c for c in (cores, cores // 2, cores // 4, 32, 16, 8) if 1 <= c <= cores}, reverse=True)
    if len(cands) == 1:
        _CPU_THREADS = cands[0]
        return _CPU_THREADS
    dims = O.DiTDims(num_layers=1)
    w = dict(synthetic.synthetic_weights(1, seed=0))
    lat, txt = synthetic.synthetic_inputs(1, 512, 512, 64)
    t = torch.tensor([0.5], dtype=torch.bfloat16)
    best, best_t = cands[0], float("inf")
    with torch.inference_mode():
        for c in cands:
            torch.set_num_threads(c)
            O.model_forward(w, dims, lat, txt, t, (1, 32, 32))
            dt = min(_timed_cpu(lambda: O.model_forward(w, dims, lat, txt, t, (1, 32, 32))) for _ in range(2))
            if dt < best_t:
                best, best_t = c, dt
    _CPU_THREADS = best
    return best


def _timed_cpu(fn) -> float:
    t0 = time.perf_counter()
    fn()
    return time.perf_counter() - t0


def cpu_reference_sample(res: int, txt_len: int, num_steps: int, layers_full: int, sample_layers: int = 2, reps: int = 5,
                         cfg: bool = False):
    """Times `sample_layers` full-width DiT blocks (bf16, B=1) at the bench resolution on the host cores — `reps` timed
    forwards after a warm-up, MEDIAN taken (a single sample moved +-40 % between runs in round 1) — and extrapolates to
    layers_full x num_steps (per-layer cost is uniform).  Returns (images/s, seconds spent, description)."""
    from oracle import qwen_image_oracle as O
    from vllm_omni_b200 import synthetic
    threads = cpu_threads()
    torch.set_num_threads(threads)
    dims = O.DiTDims(num_layers=sample_layers)
    w = dict(synthetic.synthetic_weights(sample_layers, seed=0))
    lat, txt = synthetic.synthetic_inputs(1, res, res, txt_len)
    grid = (1, res // 16, res // 16)
    t = torch.tensor([0.5], dtype=torch.bfloat16)
    with torch.inference_mode():
        O.model_forward(w, dims, lat, txt, t, grid)  # warm-up
        ts = [_timed_cpu(lambda: O.model_forward(w, dims, lat, txt, t, grid)) for _ in range(reps)]
    dt = statistics.median(ts)
    per_image = dt / sample_layers * layers_full * num_steps * (2 if cfg else 1)
    desc = (f"median of {reps} B=1 {res}px T={txt_len} bf16 forwards of {sample_layers} full-width blocks through the oracle port "
            f"of the reference torch path ({dt:.2f}s each, min {min(ts):.2f} max {max(ts):.2f}; {threads} of {os.cpu_count()} host "
            f"threads: fastest of a probe), extrapolated x{layers_full // sample_layers} layers x{num_steps} steps")
    return 1.0 / per_image, sum(ts), desc


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the path.  /root/reference is a Python
    tree that does not travel to the GPU box, so this is the oracle port (bit-exact restatement, oracle/)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, desc = [], ""
    for i in range(min(args.warmup, 1) + args.steps):
        v, _, desc = cpu_reference_sample(args.res, args.txt_len, args.num_inference_steps, args.layers, cfg=args.cfg, reps=3)
        if i >= min(args.warmup, 1):
            vals.append(v)
    value = statistics.median(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * args.batch / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"Qwen-Image DiT {args.res}px, {args.num_inference_steps} steps, bf16, batch={args.batch} "
                               f"(reference torch path on CPU, bounded sample)", "layers": args.layers, "txt_len": args.txt_len,
                   "true_cfg": bool(args.cfg), "spread": [min(vals), max(vals)]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu_threads(), "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# native arm helpers
# --------------------------------------------------------------------------------------------
def make_worker(args, world, rank, local_rank, **parallel):
    """The product's device runner (vllm_omni_b200/diffusion/worker/gpu_worker.py) with synthetic weights."""
    from vllm_omni_b200.diffusion.data import DiffusionParallelConfig, OmniDiffusionConfig, TransformerConfig
    from vllm_omni_b200.diffusion.worker.gpu_worker import GPUWorker
    pc = DiffusionParallelConfig(**parallel)
    cache_cfg = {"rel_l1_thresh": args.rel_l1_thresh} if args.cache != "none" else None
    od = OmniDiffusionConfig(model="synthetic", tf_model_config=TransformerConfig.from_dict({"num_layers": args.layers}),
                             parallel_config=pc, num_gpus=world, synthetic_weights_seed=0, cache_backend=args.cache,
                             cache_config=cache_cfg)
    return GPUWorker(local_rank=local_rank, rank=rank, od_config=od)


def cuda_timed(fn, iters, barrier):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    barrier()
    return e0.elapsed_time(e1), out


def rel_fro(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    from vllm_omni_b200 import lib as qlib
    from vllm_omni_b200 import synthetic
    from vllm_omni_b200.diffusion.distributed import parallel_state as ps
    from vllm_omni_b200.diffusion.request import OmniDiffusionRequest
    from vllm_omni_b200.flops import flops_per_forward

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist

    worker = make_worker(args, world, rank, local_rank, data_parallel_size=world)  # also initialises torch.distributed
    pipe = worker.pipeline
    qlib.device_check()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.sweep:
        run_sweep(args, pipe, world, rank, dev, barrier)
        return

    B, res, T, NS, L = args.batch, args.res, args.txt_len, args.num_inference_steps, args.layers
    S_img = (res // 16) ** 2
    grid = (1, res // 16, res // 16)
    lat_h, txt_h, neg_h = synthetic.synthetic_inputs(B, res, res, T, neg=True)
    sig = None  # scheduler default: linspace(1, 1/N, N) + dynamic shift

    def device_step(lat_d, txt_d, neg_d, cfg=args.cfg, ns=NS):
        """hot path with inputs already in HBM"""
        if worker.cache_backend is not None:
            worker.cache_backend.refresh(pipe, ns)
        timesteps, _ = pipe.prepare_timesteps(ns, sig, lat_d.shape[1])
        b = lat_d.shape[0]
        mask = torch.ones(b, T, dtype=torch.long)
        return pipe.diffuse(txt_d, mask, neg_d if cfg else None, mask if cfg else None, lat_d, [[grid]] * b, [T] * b,
                            [T] * b if cfg else None, timesteps, cfg, None, 4.0)

    # ONE request for the whole job: 4 N images; every rank holds the same request (as the engine broadcasts it)
    lat_all, txt_all, neg_all = synthetic.synthetic_inputs(B * world, res, res, T, neg=True)
    lat_all, txt_all, neg_all = lat_all.pin_memory(), txt_all.pin_memory(), neg_all.pin_memory()
    out_h = torch.empty_like(lat_all).pin_memory()

    def e2e_step():
        """the runner's public call with host buffers: shard -> H2D -> denoise -> gather -> D2H"""
        req = OmniDiffusionRequest(prompt_embeds=txt_all, negative_prompt_embeds=neg_all if args.cfg else None, latents=lat_all,
                                   height=res, width=res, num_inference_steps=NS, true_cfg_scale=4.0 if args.cfg else 1.0,
                                   output_type="latent")
        out = worker.execute_model([req], worker.od_config)
        if out.error is not None:
            raise RuntimeError(out.error)
        if rank == 0:
            out_h.copy_(out.output, non_blocking=True)
        return out

    lat_d, txt_d, neg_d = lat_h.to(dev), txt_h.to(dev), neg_h.to(dev)
    if args.graph:
        pipe.enable_cuda_graph(True)
    # ---- warm-up (>= 3 steps by contract; also builds workspaces / TMA descriptors) ----
    for _ in range(max(args.warmup, 1)):
        device_step(lat_d, txt_d, neg_d)
    barrier()

    # ---- timed region 1: kernel path, inputs resident, no per-launch instrumentation ----
    clocks = ClockSampler(local_rank)
    clocks.start()
    qlib.reset_launch_count()
    ms_dev, _ = cuda_timed(lambda: device_step(lat_d, txt_d, neg_d), args.steps, barrier)
    launches = qlib.launch_count()
    clk = clocks.stop()

    # ---- one profiled step: per-launch CUDA events around the tensor-core kernels (roofline numbers) ----
    qlib.prof_enable(True)
    ms_prof, _ = cuda_timed(lambda: device_step(lat_d, txt_d, neg_d), 1, barrier)
    qlib.prof_enable(False)
    gemm_prof, fmha_prof = qlib.prof_collect(0), qlib.prof_collect(1)

    # ---- timed region 2: end to end through the runner with host buffers ----
    ms_e2e = None
    if not args.no_e2e:
        e2e_step()
        ms_e2e, _ = cuda_timed(e2e_step, args.steps, barrier)

    t = torch.tensor([ms_dev, ms_e2e if ms_e2e is not None else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e_max = float(t[0]), float(t[1])

    line = None
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        n_img = B * world * args.steps
        value = n_img / (ms_dev / 1e3)
        fwd_per_ts = 2 if args.cfg else 1
        flops_img = flops_per_forward(L, S_img, T) * NS * fwd_per_ts
        achieved_job = flops_img * B * args.steps / (ms_dev / 1e3) / 1e12  # per GPU, TFLOP/s
        peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
        traffic, traffic_src = None, None
        tj = os.path.join(ROOT, "profiles", "r02_roofline_traffic.json")
        if os.path.exists(tj):  # dram bytes per launch of the GEMM family from the committed `ncu --set full` captures
            with open(tj) as f:
                tjd = json.load(f)
            traffic, traffic_src = tjd.get("avg_bytes_per_launch"), tjd.get("source")
        gemm_tf = gemm_prof["flops"] / (gemm_prof["ms"] / 1e3) / 1e12 if gemm_prof["ms"] > 0 else None
        fmha_tf = fmha_prof["flops"] / (fmha_prof["ms"] / 1e3) / 1e12 if fmha_prof["ms"] > 0 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"Qwen-Image DiT {res}px, {NS} steps, bf16, batch={B} per GPU on {world}xB200 "
                                   f"(BASELINE.json configs[1]; data-parallel over images)",
                       "layers": L, "txt_len": T, "true_cfg": bool(args.cfg), "images_per_step": B * world,
                       "l2": "inputs larger than L2 (41 GB weights + >1 GB activations streamed per forward)",
                       "parallelism": f"dp{world}", "cache": args.cache, "cuda_graph": bool(args.graph),
                       "attention_mode": qlib.get_fmha_mode()},
            "roofline": {"bound": "tensor", "kernel": "gemm_umma2_kernel (tcgen05 cta_group::2, grouped img+txt, fused epilogues)",
                         "achieved": gemm_tf, "peak": peak, "unit": "TFLOP/s", "frac": (gemm_tf / peak) if gemm_tf else None,
                         "peak_source": f"{peak_kind} bf16_tflops_sustained (cuBLAS 8192^3 loop)", "traffic": traffic,
                         "traffic_source": traffic_src,
                         "flops_per_launch": gemm_prof["flops"] / max(gemm_prof["launches"], 1),
                         "launches": gemm_prof["launches"], "share_of_step": gemm_prof["ms"] / ms_prof,
                         "measured_in": "one extra profiled step (per-launch CUDA events) after the timed region",
                         "fmha": {"achieved": fmha_tf, "frac": (fmha_tf / peak) if fmha_tf else None,
                                  "launches": fmha_prof["launches"], "share_of_step": fmha_prof["ms"] / ms_prof},
                         "whole_step": {"achieved": achieved_job, "frac": achieved_job / peak,
                                        "flops_per_image": flops_img}},
            "clocks": clk, "gpu_launches": launches,
        }
        if worker.cache_backend is not None:
            dec = getattr(pipe.transformer, "_teacache", None)
            if dec is not None:
                line["config"]["cache_decisions"] = {"computed": sum(1 for d in dec.decisions if d[1]), "reused": sum(1 for d in dec.decisions if not d[1])}
        if ms_e2e is not None:
            line["e2e"] = {"value": n_img / (ms_e2e_max / 1e3), "unit": UNIT,
                           "h2d_bytes_per_step": (lat_all.numel() + txt_all.numel() * (2 if args.cfg else 1)) * 2,
                           "d2h_bytes_per_step": out_h.numel() * 2,
                           "through": "GPUWorker.execute_model: one request of %d images, sharded over %d DP rank(s), gathered on rank 0" % (B * world, world)}
        if not args.no_cpu_baseline and world == 1:  # the CPU leg is reported on rank 0 at N=1 only
            try:
                v, s, desc = cpu_reference_sample(res, T, NS, L, cfg=args.cfg)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cpu_threads(), "kind": "port", "sample": desc}
            except Exception as exc:  # never lose the measured line to a host-side failure
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": f"failed: {exc!r}"}

    # ---- extra legs AFTER the headline numbers exist; a watchdog thread guarantees that the line is printed even if a leg
    #      hangs (a multi-GPU leg whose ranks diverge blocks in a collective, out of reach of try / except) ----
    printed = threading.Event()

    def emit(extra: dict):
        if printed.is_set():
            return
        printed.set()
        if rank == 0:
            line.update(extra)
            print(json.dumps(line), flush=True)

    def watchdog(limit_s: float):
        if not printed.wait(limit_s):
            emit({"extras_error": f"extra legs did not finish within {limit_s:.0f} s; headline numbers above are complete"})
            os._exit(0)

    if not args.no_extras and args.cache == "none":
        threading.Thread(target=watchdog, args=(240.0 if world == 1 else 600.0,), daemon=True).start()
        extras = {}
        try:
            if world == 1:
                extras["gpu_eager_baseline"] = eager_gpu_baseline(pipe, lat_d, txt_d, grid, T, L, S_img, B)
                extras["vae_decode"] = vae_decode_leg(B, res, dev, ms_dev / args.steps)
            else:
                extras.update(multi_gpu_legs(args, pipe, world, rank, local_rank, dev, barrier))
        except Exception as exc:
            extras["extras_error"] = repr(exc)[:300]
        emit(extras)
    else:
        emit({})
    if world > 1:
        # leave together when possible, but never hang on a peer that a failed leg left behind
        threading.Thread(target=lambda: (time.sleep(90.0), os._exit(0)), daemon=True).start()
        dist.barrier()
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------
# N = 1: the reference's eager op sequence on the same GPU (SURVEY §8d "GPU reference timing")
# --------------------------------------------------------------------------------------------
def eager_gpu_baseline(pipe, lat_d, txt_d, grid, T, L, S_img, B) -> dict:
    from baseline import eager_torch
    from vllm_omni_b200.flops import flops_per_forward
    m = pipe.transformer
    t = torch.tensor([0.5], dtype=torch.bfloat16, device=lat_d.device)
    out = {"workload": f"one DiT forward, B={B}, S_img={S_img}, T={T}, L={L}, bf16; 3 warm-up + 5 timed forwards each, same weights"}
    fl = flops_per_forward(L, S_img, T) * B
    try:
        fn = lambda: m(lat_d, txt_d, None, t, [[grid]] * B, [T] * B, return_dict=False, uniform_timestep=True)[0]  # noqa: E731
        for _ in range(3):
            nat = fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            nat = fn()
        b.record()
        torch.cuda.synchronize()
        out["native_ms"] = a.elapsed_time(b) / 5
        out["native_tflops"] = fl / out["native_ms"] / 1e9
        for key, backend in (("sdpa_ms", "sdpa"), ("flash_attn_ms", "flash_attn")):
            ms, res_ = eager_torch.time_forward(m, lat_d, txt_d, t, grid, T, backend, warmup=3, iters=5)
            out[key] = ms
            if ms is None:
                out[key.replace("_ms", "_error")] = res_
            else:
                out[key.replace("_ms", "_tflops")] = fl / ms / 1e9
                out[key.replace("_ms", "_rel_fro_native")] = rel_fro(nat, res_)
        best = min(v for v in (out.get("sdpa_ms"), out.get("flash_attn_ms")) if v)
        out["speedup_vs_best_eager"] = best / out["native_ms"]
    except Exception as exc:  # never lose the measured line
        out["error"] = repr(exc)[:300]
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------------
# N = 1: the post-step of the same request — VAE decode of the B images, native tcgen05 TF32 convolutions
# next to the reference's eager CUDA op sequence (cuDNN TF32 + SDPA, baseline/eager_torch.py) on the same weights / latents
# --------------------------------------------------------------------------------------------
def vae_decode_leg(B, res, dev, denoise_ms) -> dict:
    out = {"workload": f"AutoencoderKLQwenImage decode of {B} latents {res // 8}x{res // 8} -> {res}x{res} px, fp32 NHWC, TF32 tensor "
                       "cores, synthetic weights; 2 warm-up + 5 timed decodes each"}
    try:
        from baseline import eager_torch
        from vllm_omni_b200 import lib as qlib
        from vllm_omni_b200 import synthetic
        from vllm_omni_b200.diffusion.models.qwen_image.vae_decoder import B200VaeDecoder
        W = synthetic.synthetic_vae_decoder_weights(seed=6)
        vae = B200VaeDecoder(W, device=dev)
        z = torch.randn(B, 16, 1, res // 8, res // 8, generator=torch.Generator().manual_seed(1)).to(dev)

        def timed(fn, iters=5):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / iters

        n0 = qlib.launch_count()
        img = vae.decode(z, return_dict=False)[0]
        out["launches_per_decode"] = qlib.launch_count() - n0
        out["native_ms"] = timed(lambda: vae.decode(z, return_dict=False))
        out["native_uint8_ms"] = timed(lambda: vae.decode_to_uint8(z))
        out["gflop_per_image"] = 4710.2 * (res / 1024) ** 2  # convolution + attention FLOPs (attention part scales faster; 1024 px value)
        out["native_tflops"] = B * out["gflop_per_image"] / out["native_ms"]
        out["share_of_request_time"] = out["native_ms"] / (out["native_ms"] + denoise_ms)
        Wd = {k: v.to(dev) for k, v in W.items()}
        with torch.no_grad():
            ref = eager_torch.vae_decode_eager(z, Wd)
            out["eager_cudnn_tf32_ms"] = timed(lambda: eager_torch.vae_decode_eager(z, Wd))
        out["speedup_vs_eager"] = out["eager_cudnn_tf32_ms"] / out["native_ms"]
        out["max_abs_vs_eager"] = float((img - ref).abs().max())
    except Exception as exc:  # never lose the measured line
        out["error"] = repr(exc)[:300]
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------------
# N > 1: tensor parallel (BASELINE configs[2]) and CFG parallel, measured after the DP region
# --------------------------------------------------------------------------------------------
def multi_gpu_legs(args, dp_pipe, world, rank, local_rank, dev, barrier) -> dict:
    import torch.distributed as dist

    from vllm_omni_b200 import synthetic
    from vllm_omni_b200.diffusion.data import DiffusionParallelConfig, OmniDiffusionConfig, TransformerConfig
    from vllm_omni_b200.diffusion.distributed import parallel_state as ps
    from vllm_omni_b200.diffusion.models.qwen_image.pipeline_qwen_image import QwenImagePipeline

    res, T, NS, L = args.res, args.txt_len, args.num_inference_steps, args.layers
    S_img, grid, D = (res // 16) ** 2, (1, res // 16, res // 16), 3072
    out: dict = {}

    def denoise(pipe, lat, txt, neg=None, ns=NS):
        timesteps, _ = pipe.prepare_timesteps(ns, None, lat.shape[1])
        b = lat.shape[0]
        mask = torch.ones(b, T, dtype=torch.long)
        cfg = neg is not None
        return pipe.diffuse(txt, mask, neg, mask if cfg else None, lat, [[grid]] * b, [T] * b, [T] * b if cfg else None,
                            timesteps, cfg, None, 4.0)

    def max_ms(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # ---- single-GPU reference for both legs: every rank runs its full replica on the same inputs ----
    t_half = torch.tensor([0.5], dtype=torch.bfloat16, device=dev)
    ref = {}
    for b in (1, 4):
        lat, txt = (t.to(dev) for t in synthetic.synthetic_inputs(b, res, res, T))
        denoise(dp_pipe, lat, txt, ns=2)
        ms_1, lat_1 = cuda_timed(lambda: denoise(dp_pipe, lat, txt), 1, barrier)
        f_1 = dp_pipe.transformer(lat, txt, None, t_half, [[grid]] * b, [T] * b, return_dict=False, uniform_timestep=True)[0]
        ref[b] = (lat, txt, max_ms(ms_1), lat_1, f_1.clone())

    def leg(pipe, b, bytes_per_forward):
        lat, txt, ms_1, lat_1, f_1 = ref[b]
        denoise(pipe, lat, txt, ns=2)  # warm-up: workspaces, IPC exchange, descriptors
        ms_p, lat_p = cuda_timed(lambda: denoise(pipe, lat, txt), 1, barrier)
        ms_p = max_ms(ms_p)
        f_p = pipe.transformer(lat, txt, None, t_half, [[grid]] * b, [T] * b, return_dict=False, uniform_timestep=True)[0]
        return {"value": b / (ms_p / 1e3), "unit": UNIT, "ms_per_forward": ms_p / NS, "ms_per_forward_n1": ms_1 / NS,
                "speedup_vs_n1": ms_1 / ms_p, "rel_fro_vs_single_gpu": rel_fro(f_p, f_1),
                "bit_identical_to_single_gpu": bool(torch.equal(f_p, f_1)) and bool(torch.equal(lat_p, lat_1)),
                "rel_fro_vs_single_gpu_50_steps": rel_fro(lat_p, lat_1), "nvlink_bytes_per_forward": bytes_per_forward}

    # ---- fused sequence parallelism (Ulysses), all N ranks in one SP group; the DP replica's own weights ----
    try:
        ps.initialize_model_parallel(data_parallel_size=1, tensor_parallel_size=1, ulysses_degree=world)
        dp_pipe.transformer.enable_sequence_parallel(world, ps.get_sequence_parallel_rank(), ps.get_sp_group())
        sp = {"sp_size": world,
              "kernel": "gemm_umma2_kernel<EPI_QKV> stores each head's q/k/v rows into the head owner's buffers, "
                        "fmha_joint_kernel stores each output row into the row owner's buffer (peer stores over NVLink); "
                        "all other kernels run on this rank's rows with the full weights"}
        for b in (1, 4):
            rows = b * (S_img + T)
            sp[f"b{b}"] = leg(dp_pipe, b, L * (world - 1) / world * (rows / world) * (3 * D + D) * 2)
        sp["healthy"] = bool(dp_pipe.transformer.p2p_healthy())
        out["sp"] = sp
    except Exception as exc:
        out["sp"] = {"error": repr(exc)[:400]}
    try:
        dp_pipe.transformer.enable_sequence_parallel(1, 0, None)
    except Exception as exc:
        out.setdefault("sp", {})["teardown_error"] = repr(exc)[:200]
    barrier()

    # ---- tensor parallel over heads / FFN, all N ranks in one TP group ----
    try:
        ps.initialize_model_parallel(data_parallel_size=1, tensor_parallel_size=world)
        od = OmniDiffusionConfig(model="synthetic", tf_model_config=TransformerConfig.from_dict({"num_layers": L}),
                                 parallel_config=DiffusionParallelConfig(tensor_parallel_size=world), num_gpus=world)
        torch.set_default_dtype(torch.bfloat16)
        try:
            with torch.device(dev):
                tp_pipe = QwenImagePipeline(od_config=od)
        finally:
            torch.set_default_dtype(torch.float32)
        tp_pipe.transformer.load_weights(synthetic.synthetic_weights(L, seed=0, device=dev, device_generate=True))
        tp = {"tp_size": world, "comm": tp_pipe.transformer.tp_comm,
              "kernel": "gemm_umma2_kernel<EPI_PARTIAL_F32> (fp32 partial tiles pushed to the row owners over NVLink) + "
                        "tp_reduce_ln_push_kernel (reduce + bias/gate/residual + next AdaLN + all-gather)"}
        for b in (1, 4):
            rows = b * (S_img + T)
            tp[f"b{b}"] = leg(tp_pipe, b, 2 * L * (world - 1) / world * rows * D * (4 + 2))
        tp["healthy"] = bool(tp_pipe.transformer.p2p_healthy()) if tp_pipe.transformer.tp_comm == "p2p" else True
        out["tp"] = tp
        del tp_pipe
        torch.cuda.empty_cache()
    except Exception as exc:
        out["tp"] = {"error": repr(exc)[:400]}
    barrier()

    # ---- CFG parallel: positive / negative branch of a true-CFG step on the two ranks of a CFG group ----
    if world % 2 == 0:
        try:
            lat, txt, neg = (t.to(dev) for t in synthetic.synthetic_inputs(1, res, res, T, neg=True))
            ps.initialize_model_parallel(data_parallel_size=world, tensor_parallel_size=1, cfg_parallel_size=1)
            denoise(dp_pipe, lat, txt, neg, ns=2)
            ms_seq, lat_seq = cuda_timed(lambda: denoise(dp_pipe, lat, txt, neg), 1, barrier)
            ps.initialize_model_parallel(data_parallel_size=world // 2, tensor_parallel_size=1, cfg_parallel_size=2)
            denoise(dp_pipe, lat, txt, neg, ns=2)
            ms_par, lat_par = cuda_timed(lambda: denoise(dp_pipe, lat, txt, neg), 1, barrier)
            ms_seq, ms_par = max_ms(ms_seq), max_ms(ms_par)
            out["cfg_parallel"] = {"workload": f"true-CFG denoise, B=1 per CFG group, {NS} steps, {world // 2} group(s) of 2 ranks",
                                   "value": (world // 2) / (ms_par / 1e3), "unit": UNIT, "ms_per_image_sequential": ms_seq,
                                   "ms_per_image_cfg_parallel": ms_par, "speedup": ms_seq / ms_par,
                                   "bit_equal_to_sequential": bool(torch.equal(lat_seq, lat_par))}
        except Exception as exc:
            out["cfg_parallel"] = {"error": repr(exc)[:400]}
        ps.initialize_model_parallel(data_parallel_size=world, tensor_parallel_size=1, cfg_parallel_size=1)
    barrier()
    return out


# --------------------------------------------------------------------------------------------
# BASELINE configs[4]: batch x resolution sweep
# --------------------------------------------------------------------------------------------
def run_sweep(args, pipe, world, rank, dev, barrier):
    import torch.distributed as dist

    from vllm_omni_b200 import synthetic
    from vllm_omni_b200.flops import flops_per_forward
    peaks, _ = measured_peaks()
    peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
    T, L, NS, nts = args.txt_len, args.layers, args.num_inference_steps, args.sweep_timesteps
    cells = []
    for res in [int(v) for v in args.sweep_res.split(",")]:
        S_img, grid = (res // 16) ** 2, (1, res // 16, res // 16)
        for B in [int(v) for v in args.sweep_batches.split(",")]:
            for cfg in ((False, True) if (res == 1024 and B == 4) else (False,)):
                lat, txt, neg = (t.to(dev) for t in synthetic.synthetic_inputs(B, res, res, T, neg=True))
                mask = torch.ones(B, T, dtype=torch.long)

                def run(n):
                    timesteps, _ = pipe.prepare_timesteps(NS, None, S_img)  # the real 50-step schedule, first n steps
                    return pipe.diffuse(txt, mask, neg if cfg else None, mask if cfg else None, lat, [[grid]] * B, [T] * B,
                                        [T] * B if cfg else None, timesteps[:n], cfg, None, 4.0)
                try:
                    run(1)
                    ms, _ = cuda_timed(lambda: run(nts), 1, barrier)
                    t = torch.tensor([ms], dtype=torch.float64, device=dev)
                    if world > 1:
                        dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms_ts = float(t) / nts  # per denoise timestep (1 or 2 forwards)
                    fl = flops_per_forward(L, S_img, T) * B * (2 if cfg else 1)
                    cells.append({"res": res, "batch_per_gpu": B, "true_cfg": cfg, "ms_per_timestep": ms_ts,
                                  "images_per_s_50_steps": B * world / (ms_ts * NS / 1e3), "tflops_per_gpu": fl / ms_ts / 1e9,
                                  "frac_of_sustained_peak": fl / ms_ts / 1e9 / peak})
                except Exception as exc:
                    cells.append({"res": res, "batch_per_gpu": B, "true_cfg": cfg, "error": repr(exc)[:200]})
                del lat, txt, neg
                torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps({"sweep": "BASELINE configs[4]: batch x resolution, data-parallel over images", "n_gpus": world,
                          "layers": L, "txt_len": T, "timed_timesteps_per_cell": nts,
                          "note": "images_per_s_50_steps = batch / (50 x measured per-timestep time); every timestep costs the same",
                          "peak_tflops": peak, "cells": cells}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
