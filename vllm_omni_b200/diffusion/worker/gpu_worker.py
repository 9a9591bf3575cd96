"""Device runner — drop-in for the reference `GPUWorker` / `WorkerProc`
(vllm_omni/diffusion/worker/gpu_worker.py:32-314): one process per GPU, same constructor, same
`generate / execute_model / shutdown`, same message protocol in `worker_busy_loop`
({"type":"rpc",...} / {"type":"shutdown"} / request list) and the same ready handshake in
`worker_main(rank, od_config, pipe_writer, broadcast_handle)`.

Differences, all behind that interface:
  * the pipeline is the native sm_100a `QwenImagePipeline` (registry.initialize_model);
  * data parallelism is REAL: the reference executes only reqs[0] on every rank (:129-130); here the
    images of a request (`num_outputs_per_prompt` x prompts) are sharded over the DP ranks, each rank
    denoises its share with its own full copy of the weights, and rank 0 gathers the latents;
  * queues are duck-typed (`dequeue(indefinite=True)` / `enqueue(obj)` / `export_handle()`): on the
    reference side they are vLLM shm `MessageQueue`s, in tests plain mp queues.
"""
from __future__ import annotations

import dataclasses
import logging
import os
import time

import torch

from vllm_omni_b200.diffusion import registry
from vllm_omni_b200.diffusion.data import DiffusionOutput, OmniDiffusionConfig, set_current_omni_diffusion_config
from vllm_omni_b200.diffusion.distributed import parallel_state as ps
from vllm_omni_b200.diffusion.request import OmniDiffusionRequest

logger = logging.getLogger(__name__)


def shard_request(req: OmniDiffusionRequest, dp_rank: int, dp_size: int):
    """Split the image units of one request over DP ranks.  Returns (local request or None, counts per rank).
    Units are (prompt, output index) pairs laid out prompt-major, exactly the batch order the reference
    pipeline builds with `repeat(1, num_images_per_prompt, 1)` (pipeline_qwen_image.py:427-431)."""
    import dataclasses

    pe = req.prompt_embeds
    n_prompts = pe.shape[0] if isinstance(pe, torch.Tensor) else (len(req.prompt) if isinstance(req.prompt, list) else 1)
    per = max(int(req.num_outputs_per_prompt or 1), 1)
    total = n_prompts * per
    counts = [ps.shard_range(total, r, dp_size)[1] - ps.shard_range(total, r, dp_size)[0] for r in range(dp_size)]
    lo, hi = ps.shard_range(total, dp_rank, dp_size)
    if hi == lo:
        return None, counts
    if not isinstance(pe, torch.Tensor):
        raise ValueError("DP sharding needs pre-computed prompt_embeds (text encoding is outside the native engine)")
    idx = torch.arange(lo, hi) // per  # prompt index of each local unit

    def take(t):
        return None if t is None else t[idx]

    def trim(e, m):
        """Drop the padding columns no prompt of THIS shard uses (a shard that lacks the longest prompt would otherwise
        carry a padded length T > max(txt_seq_lens), which the transformer rejects like the reference's RoPE broadcast)."""
        if e is None or m is None:
            return e, m
        t_loc = max(int(m.sum(dim=1).max()), 1)
        return e[:, :t_loc].contiguous(), m[:, :t_loc].contiguous()

    pe_l, pm_l = trim(take(pe), take(req.prompt_attention_mask))
    ne_l, nm_l = trim(take(req.negative_prompt_embeds), take(req.negative_attention_mask))
    local = dataclasses.replace(
        req, prompt_embeds=pe_l, negative_prompt_embeds=ne_l, prompt_attention_mask=pm_l, negative_attention_mask=nm_l,
        num_outputs_per_prompt=1, latents=None if req.latents is None else req.latents[lo:hi])
    if req.latents is None and req.seed is not None:
        # deterministic per-unit noise independent of the DP layout (also used at dp = 1): one generator per global unit
        g = [torch.Generator().manual_seed(req.seed + u) for u in range(lo, hi)]
        local.extra = dict(req.extra, unit_generators=g)
    return local, counts


def batch_key(req: OmniDiffusionRequest):
    """Requests that can share one denoise batch: same geometry, schedule, CFG setting and text length (the transformer's
    RoPE needs one text length per batch, and attention carries no mask — padding a shorter prompt would change its
    result, so only equal lengths are merged), pre-computed embeddings, no condition image, no caller-provided latents."""
    pe = req.prompt_embeds
    if not isinstance(pe, torch.Tensor) or req.extra or req.latents is not None or req.generator is not None:
        return None
    ne = req.negative_prompt_embeds
    return (req.height, req.width, req.num_inference_steps, req.true_cfg_scale, tuple(req.sigmas) if req.sigmas is not None else None,
            req.output_type, pe.shape[1], None if ne is None else ne.shape[1], req.num_outputs_per_prompt,
            req.prompt_attention_mask is None, req.negative_attention_mask is None, req.seed is None)


def merge_requests(reqs: list[OmniDiffusionRequest]):
    """Cross-request batching (the reference leaves it as a TODO and runs reqs[0] only, gpu_worker.py:129-130): requests
    with the same `batch_key` are concatenated along the prompt axis.  Returns [(merged request, [(index, n_units)])]."""
    groups: dict = {}
    order = []
    for i, r in enumerate(reqs):
        k = batch_key(r)
        k = ("solo", i) if k is None else k
        if k not in groups:
            groups[k] = []
            order.append(k)
        groups[k].append(i)
    out = []
    for k in order:
        idx = groups[k]
        per = max(int(reqs[idx[0]].num_outputs_per_prompt or 1), 1)
        if len(idx) == 1:
            r = reqs[idx[0]]
            n = (r.prompt_embeds.shape[0] if isinstance(r.prompt_embeds, torch.Tensor) else 1) * per
            out.append((r, [(idx[0], n)]))
            continue

        def cat(name):
            ts = [getattr(reqs[i], name) for i in idx]
            return None if ts[0] is None else torch.cat(ts, dim=0)

        # per-unit noise: every unit keeps the seed it would have had on its own (seed + local unit index)
        seeds = None
        if all(reqs[i].seed is not None for i in idx):
            seeds = [reqs[i].seed + u for i in idx for u in range(reqs[i].prompt_embeds.shape[0] * per)]
        merged = dataclasses.replace(reqs[idx[0]], prompt_embeds=cat("prompt_embeds"), negative_prompt_embeds=cat("negative_prompt_embeds"),
                                     prompt_attention_mask=cat("prompt_attention_mask"), negative_attention_mask=cat("negative_attention_mask"),
                                     seed=None, extra={"unit_seeds": seeds} if seeds else {})
        out.append((merged, [(i, reqs[i].prompt_embeds.shape[0] * per) for i in idx]))
    return out


class GPUWorker:
    def __init__(self, local_rank: int, rank: int, od_config: OmniDiffusionConfig):
        self.local_rank, self.rank, self.od_config = local_rank, rank, od_config
        self.pipeline = None
        self.cache_backend = None
        self.init_device_and_model()

    def init_device_and_model(self) -> None:
        world_size = self.od_config.num_gpus
        os.environ["MASTER_ADDR"] = os.environ.get("MASTER_ADDR", "127.0.0.1")
        if self.od_config.master_port is not None:
            os.environ["MASTER_PORT"] = str(self.od_config.master_port)
        os.environ["LOCAL_RANK"], os.environ["RANK"], os.environ["WORLD_SIZE"] = str(self.local_rank), str(self.rank), str(world_size)
        if not torch.cuda.is_available():
            raise RuntimeError("GPUWorker needs a CUDA (sm_100) device: the native engine has no CPU path")
        device = torch.device(f"cuda:{self.local_rank}")
        torch.cuda.set_device(device)
        with set_current_omni_diffusion_config(self.od_config):
            if world_size > 1:
                ps.init_distributed_environment(world_size=world_size, rank=self.rank)
            pc = self.od_config.parallel_config
            ps.initialize_model_parallel(data_parallel_size=pc.data_parallel_size, tensor_parallel_size=pc.tensor_parallel_size,
                                         cfg_parallel_size=pc.cfg_parallel_size, ulysses_degree=pc.ulysses_degree)
            t0 = time.perf_counter()
            prev = torch.get_default_dtype()
            torch.set_default_dtype(self.od_config.dtype)
            try:
                with torch.device(device):
                    self.pipeline = registry.initialize_model(self.od_config)
            finally:
                torch.set_default_dtype(prev)
            if self.od_config.synthetic_weights_seed is not None:
                from vllm_omni_b200 import synthetic
                self.pipeline.transformer.load_weights(synthetic.synthetic_weights(
                    self.pipeline.transformer.num_layers, seed=self.od_config.synthetic_weights_seed, device=device,
                    device_generate=True))
            logger.info("Worker %d: model ready in %.2fs", self.rank, time.perf_counter() - t0)
            # step cache (reference gpu_worker.py:103-107)
            from vllm_omni_b200.diffusion.cache.selector import get_cache_backend
            self.cache_backend = get_cache_backend(self.od_config.cache_backend, self.od_config.cache_config)
            if self.cache_backend is not None:
                self.cache_backend.enable(self.pipeline)

    def generate(self, requests: list[OmniDiffusionRequest]) -> DiffusionOutput:
        return self.execute_model(requests, self.od_config)

    @torch.inference_mode()
    def execute_model(self, reqs: list[OmniDiffusionRequest], od_config: OmniDiffusionConfig) -> DiffusionOutput:
        """One request (what the reference scheduler sends, scheduler.py:51-72) or a LIST of requests: compatible ones are
        merged into one denoise batch (`merge_requests`); the output then is the concatenation of the per-request results
        in request order (`DiffusionOutput.output` [sum units, S_img, 64]; `trajectory_timesteps` carries the unit counts)."""
        assert self.pipeline is not None
        if not reqs:
            raise ValueError("Cannot execute model with empty request list")
        if len(reqs) == 1:
            return self._execute_one(reqs[0])
        results: list = [None] * len(reqs)
        for merged, parts in merge_requests(list(reqs)):
            out = self._execute_one(merged)
            if out.error is not None:
                return out
            if out.output is None:  # non-zero DP ranks
                continue
            off = 0
            for i, n in parts:
                results[i] = out.output[off:off + n]
                off += n
        if any(r is None for r in results):
            return DiffusionOutput(output=None)
        return DiffusionOutput(output=torch.cat(results, dim=0), trajectory_timesteps=[int(r.shape[0]) for r in results])

    def _execute_one(self, req: OmniDiffusionRequest) -> DiffusionOutput:
        if self.cache_backend is not None and self.cache_backend.is_enabled():  # reference :132-134
            self.cache_backend.refresh(self.pipeline, req.num_inference_steps)
        if req.extra and req.extra.get("unit_seeds") and req.latents is None:
            # merged requests: one generator per unit, seeded as the unit's own request would have seeded it
            gens = [torch.Generator().manual_seed(s) for s in req.extra["unit_seeds"]]
            req = dataclasses.replace(req, extra={}, latents=self._unit_latents(dataclasses.replace(req, extra={"unit_generators": gens})))
        dp = ps.get_data_parallel_world_size()
        if dp == 1:
            if req.latents is None and req.seed is not None and isinstance(req.prompt_embeds, torch.Tensor):
                # the same per-unit noise a DP run draws, so dp=1 and dp>1 produce the same images for a seed
                local, _ = shard_request(req, 0, 1)
                req = dataclasses.replace(req, latents=self._unit_latents(local))
            return self.pipeline.forward(req)
        local, counts = shard_request(req, ps.get_data_parallel_rank(), dp)
        lat, err = None, None
        try:
            if local is not None:
                if local.latents is None and local.extra and local.extra.get("unit_generators") is not None:
                    local.latents = self._unit_latents(local)
                local.output_type = "latent"
                out = self.pipeline.forward(local)
                lat, err = out.output, out.error
        except Exception as e:  # a failing rank must not leave the others waiting in the gather
            logger.error("Worker %d: denoise failed: %s", self.rank, e, exc_info=True)
            err = str(e)
        flag = torch.tensor([1 if err else 0], dtype=torch.int32, device=self.pipeline.device)
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MAX, group=ps.get_dp_group())
        if int(flag.item()):
            return DiffusionOutput(error=err or "a data-parallel peer failed (see its log)")
        if lat is None:
            s_img = ((req.height or 1024) // 16) * ((req.width or 1024) // 16)
            lat = torch.zeros((0, s_img, self.pipeline.transformer.in_channels), dtype=torch.bfloat16, device=self.pipeline.device)
        full = ps.gather_to_rank0(lat, counts)
        return DiffusionOutput(output=full)

    def _unit_latents(self, local: OmniDiffusionRequest) -> torch.Tensor:
        h, w = local.height or 1024, local.width or 1024
        return torch.cat([self.pipeline.prepare_latents(1, self.pipeline.transformer.in_channels // 4, h, w, torch.bfloat16,
                                                        self.pipeline.device, g) for g in local.extra["unit_generators"]])

    def shutdown(self) -> None:
        ps.destroy_distributed_env()


class WorkerProc:
    """Runs one GPUWorker in its own process; message protocol of the reference WorkerProc (:143-314)."""

    def __init__(self, od_config: OmniDiffusionConfig, gpu_id: int, broadcast_queue, result_queue=None):
        self.od_config, self.gpu_id = od_config, gpu_id
        self.mq = broadcast_queue
        self.result_mq = result_queue if gpu_id == 0 else None
        self.worker = self._create_worker(gpu_id, od_config)
        self._running = True

    def _create_worker(self, gpu_id: int, od_config: OmniDiffusionConfig) -> GPUWorker:
        return GPUWorker(local_rank=gpu_id, rank=gpu_id, od_config=od_config)

    def return_result(self, output):
        if self.result_mq is not None:
            self.result_mq.enqueue(output)

    def recv_message(self):
        return self.mq.dequeue(indefinite=True)

    def execute_rpc(self, rpc_request: dict):
        method = rpc_request["method"]
        args, kwargs = rpc_request.get("args", ()), rpc_request.get("kwargs", {})
        output_rank, exec_all = rpc_request.get("output_rank"), rpc_request.get("exec_all_ranks", False)
        should_execute = exec_all or output_rank is None or output_rank == self.gpu_id
        should_reply = (output_rank is None or output_rank == self.gpu_id) and self.result_mq is not None
        if not should_execute:
            return None, False
        try:
            func = getattr(self.worker, method) if isinstance(method, str) else (lambda *a, **k: method(self.worker, *a, **k))
            return func(*args, **kwargs), should_reply
        except Exception as e:  # kernel / shape errors surface as Python exceptions -> error reply (reference :221-223)
            logger.error("Error executing RPC: %s", e, exc_info=True)
            return {"status": "error", "error": str(e)}, should_reply

    def worker_busy_loop(self) -> None:
        while self._running:
            try:
                msg = self.recv_message()
            except Exception as e:
                logger.error("Error receiving message in worker loop: %s", e, exc_info=True)
                continue
            if msg is None or (hasattr(msg, "__len__") and len(msg) == 0):
                continue
            if isinstance(msg, dict) and msg.get("type") == "rpc":
                result, should_reply = self.execute_rpc(msg)
                if should_reply:
                    self.return_result(result)
            elif isinstance(msg, dict) and msg.get("type") == "shutdown":
                self._running = False
            else:
                try:
                    output = self.worker.execute_model(msg, self.od_config)
                except Exception as e:
                    logger.error("Error executing forward in event loop: %s", e, exc_info=True)
                    output = DiffusionOutput(error=str(e))
                self.return_result(output)
        try:
            self.worker.shutdown()
        except Exception as exc:  # best effort
            logger.warning("Worker %s: shutdown error: %s", self.gpu_id, exc)

    @staticmethod
    def worker_main(rank: int, od_config: OmniDiffusionConfig, pipe_writer, broadcast_handle, result_queue=None) -> None:
        """`broadcast_handle` is the queue (or a handle with `.open(rank)`) the engine broadcasts on."""
        mq = broadcast_handle.open(rank) if hasattr(broadcast_handle, "open") else broadcast_handle
        proc = WorkerProc(od_config, gpu_id=rank, broadcast_queue=mq, result_queue=result_queue)
        handle = None
        if rank == 0 and proc.result_mq is not None and hasattr(proc.result_mq, "export_handle"):
            handle = proc.result_mq.export_handle()
        pipe_writer.send({"status": "ready", "result_handle": handle})
        proc.worker_busy_loop()


def get_diffusion_worker_class():
    """What the reference's `get_diffusion_worker_class()` (utils/platform_utils.py:39-58) returns on a B200 box."""
    return WorkerProc
