"""DeviceTensorConnector — an OmniConnector for the edge "prompt-encoder stage -> DiT stage" (BASELINE
configs[3]).  Same call contract as the reference's connectors as used by `try_send_via_connector` /
`try_recv_via_connector` (vllm_omni/distributed/omni_connectors/adapter.py:15-90,90-170):

    put(from_stage, to_stage, request_id, payload) -> (success, serialized_size_bytes, metadata)
    get(from_stage, to_stage, request_id, metadata=None) -> payload

The reference's connectors serialise the payload (shared memory / Mooncake).  For this edge the payload is the prompt
embeddings ([B, T, 3584] bf16, ~0.9 MB per prompt): they stay DEVICE tensors, handed over by reference together with a
CUDA event recorded on the producer's stream; `get` makes the consumer's current stream wait on that event, so the two
stages overlap on their own streams (or GPUs — the tensor is then copied peer-to-peer by the consumer) with no host
synchronisation.  In-process (threads), like the reference's single-process orchestrator mode."""
from __future__ import annotations

import threading
from typing import Any

import torch


class DeviceTensorConnector:
    def __init__(self) -> None:
        self._slots: dict[tuple[str, str, str], tuple[Any, Any]] = {}
        self._cv = threading.Condition()

    @staticmethod
    def _nbytes(obj) -> int:
        if isinstance(obj, torch.Tensor):
            return obj.numel() * obj.element_size()
        if isinstance(obj, dict):
            return sum(DeviceTensorConnector._nbytes(v) for v in obj.values())
        if isinstance(obj, (list, tuple)):
            return sum(DeviceTensorConnector._nbytes(v) for v in obj)
        return 0

    def put(self, from_stage: str, to_stage: str, request_id: str, payload: Any):
        event = None
        if torch.cuda.is_available() and self._has_cuda(payload):
            event = torch.cuda.Event()
            event.record(torch.cuda.current_stream())  # everything the producer enqueued for this payload
        with self._cv:
            self._slots[(str(from_stage), str(to_stage), str(request_id))] = (payload, event)
            self._cv.notify_all()
        return True, self._nbytes(payload), {"in_process": True}

    def get(self, from_stage: str, to_stage: str, request_id: str, metadata: Any = None, timeout: float | None = 60.0):
        key = (str(from_stage), str(to_stage), str(request_id))
        with self._cv:
            if not self._cv.wait_for(lambda: key in self._slots, timeout=timeout):
                raise TimeoutError(f"connector: no payload for {key}")
            payload, event = self._slots.pop(key)
        if event is not None:
            torch.cuda.current_stream().wait_event(event)  # device-side ordering, no host block
        return payload

    @staticmethod
    def _has_cuda(obj) -> bool:
        if isinstance(obj, torch.Tensor):
            return obj.is_cuda
        if isinstance(obj, dict):
            return any(DeviceTensorConnector._has_cuda(v) for v in obj.values())
        if isinstance(obj, (list, tuple)):
            return any(DeviceTensorConnector._has_cuda(v) for v in obj)
        return False
