"""Two-stage overlap for text-to-image (BASELINE configs[3]): stage 0 encodes prompts (the Qwen2.5-VL text
encoder the reference runs INSIDE the DiT worker, pipeline_qwen_image.py:357-396), stage 1 denoises; while stage 1 works
on request i, stage 0 already encodes request i+1, and the embeddings travel through an OmniConnector-shaped hand-over
(`DeviceTensorConnector`) instead of blocking the DiT loop.  The reference builds such chains from `stage_args` YAML
(`stage_type: llm` -> `stage_type: diffusion`, entrypoints/omni_stage.py:403-, joined by
distributed/omni_connectors/adapter.py); its orchestrator, queues and process management are out of scope (§8 "unchanged")
— this is the in-process equivalent for the one edge the DiT hot path has.

Each stage runs in its own host thread on its own CUDA stream (and optionally its own device), so stage 0's kernels
overlap stage 1's; the only cross-stage dependency is the connector event."""
from __future__ import annotations

import dataclasses
import queue
import threading
import time
from collections.abc import Callable
from typing import Any

import torch

from vllm_omni_b200.diffusion.data import DiffusionOutput
from vllm_omni_b200.diffusion.request import OmniDiffusionRequest
from vllm_omni_b200.distributed.device_connector import DeviceTensorConnector


class TwoStagePipeline:
    """`encode(prompts) -> (embeds, mask)` feeds `denoise(request) -> DiffusionOutput`.

    encode   : e.g. `QwenImagePipeline._get_qwen_prompt_embeds` of a pipeline object that owns the text encoder, or any
               callable with that contract (an upstream AR stage);
    denoise  : e.g. `GPUWorker.execute_model([req], od_config)` / `QwenImagePipeline.forward`.
    """

    ENC, DIT = "0", "1"

    def __init__(self, encode: Callable[[Any], tuple[torch.Tensor, torch.Tensor]], denoise: Callable[[OmniDiffusionRequest], DiffusionOutput],
                 encoder_device: torch.device | None = None, dit_device: torch.device | None = None,
                 connector: DeviceTensorConnector | None = None):
        self.encode, self.denoise = encode, denoise
        self.encoder_device, self.dit_device = encoder_device, dit_device
        self.connector = connector or DeviceTensorConnector()
        self.timeline: list[tuple[str, str, float, float]] = []  # (stage, request id, start, end) host times, for tests / tracing
        self._lock = threading.Lock()

    def _log(self, stage, rid, t0, t1):
        with self._lock:
            self.timeline.append((stage, rid, t0, t1))

    def _stream_ctx(self, device):
        if device is None or not torch.cuda.is_available() or torch.device(device).type != "cuda":
            import contextlib
            return contextlib.nullcontext()
        return torch.cuda.stream(torch.cuda.Stream(device=device))

    def generate(self, requests: list[OmniDiffusionRequest], overlap: bool = True) -> list[DiffusionOutput]:
        """Requests carry `prompt` (and optionally `negative_prompt`); results come back in request order.
        overlap=False runs encode -> denoise strictly one request after the other (the reference's in-worker order)."""
        reqs = [r if r.request_id is not None else dataclasses.replace(r, request_id=f"req-{i}") for i, r in enumerate(requests)]
        results: dict[str, DiffusionOutput] = {}
        errors: list[BaseException] = []
        ready: "queue.Queue[str | None]" = queue.Queue()

        def stage0(r: OmniDiffusionRequest):
            t0 = time.perf_counter()
            payload = {"prompt": self.encode(r.prompt)}
            if r.negative_prompt is not None and (r.true_cfg_scale or 0) > 1:
                payload["negative"] = self.encode(r.negative_prompt)
            self.connector.put(self.ENC, self.DIT, r.request_id, payload)
            self._log("encode", r.request_id, t0, time.perf_counter())

        def stage1(r: OmniDiffusionRequest):
            t0 = time.perf_counter()
            payload = self.connector.get(self.ENC, self.DIT, r.request_id)
            dev = self.dit_device

            def mv(t):
                return t if dev is None else t.to(dev, non_blocking=True)
            pe, pm = payload["prompt"]
            kw = dict(prompt_embeds=mv(pe), prompt_attention_mask=pm.cpu() if pm is not None else None)
            if "negative" in payload:
                ne, nm = payload["negative"]
                kw.update(negative_prompt_embeds=mv(ne), negative_attention_mask=nm.cpu() if nm is not None else None)
            results[r.request_id] = self.denoise(dataclasses.replace(r, **kw))
            self._log("denoise", r.request_id, t0, time.perf_counter())

        if not overlap:
            for r in reqs:
                stage0(r)
                stage1(r)
            return [results[r.request_id] for r in reqs]

        def run0():
            try:
                with self._stream_ctx(self.encoder_device):
                    for r in reqs:
                        stage0(r)
                        ready.put(r.request_id)
            except BaseException as e:  # surfaced to the caller below
                errors.append(e)
            finally:
                ready.put(None)

        def run1():
            try:
                with self._stream_ctx(self.dit_device):
                    by_id = {r.request_id: r for r in reqs}
                    while True:
                        rid = ready.get()
                        if rid is None:
                            break
                        stage1(by_id[rid])
            except BaseException as e:
                errors.append(e)

        t0, t1 = threading.Thread(target=run0, name="stage0-encode"), threading.Thread(target=run1, name="stage1-dit")
        t0.start(); t1.start(); t0.join(); t1.join()
        if errors:
            raise errors[0]
        return [results[r.request_id] for r in reqs]
