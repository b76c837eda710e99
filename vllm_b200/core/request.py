"""Request data model.

``EngineCoreRequest``/``EngineCoreOutput(s)`` are the wire structs between the
front-end process and the engine core (reference: vllm/v1/engine/__init__.py);
``Request`` is the scheduler-side mutable state (reference: vllm/v1/request.py).
"""
from __future__ import annotations

import enum
import time
from typing import Any, Optional, Union

import msgspec

from vllm_b200.sampling_params import PoolingParams, SamplingParams


class FinishReason(enum.IntEnum):
    STOP = 0
    LENGTH = 1
    ABORT = 2

    def __str__(self) -> str:
        return ("stop", "length", "abort")[int(self)]


class LoRARequest(msgspec.Struct, omit_defaults=True):
    lora_name: str
    lora_int_id: int
    lora_path: str = ""
    base_model_name: Optional[str] = None


class EngineCoreRequest(msgspec.Struct, omit_defaults=True):
    request_id: str
    prompt_token_ids: list[int]
    sampling_params: Optional[SamplingParams] = None
    pooling_params: Optional[PoolingParams] = None
    eos_token_id: Optional[int] = None
    arrival_time: float = 0.0
    lora_request: Optional[LoRARequest] = None
    cache_salt: Optional[str] = None
    priority: int = 0
    data_parallel_rank: Optional[int] = None
    kv_transfer_params: Optional[dict[str, Any]] = None
    mm_features: Optional[list[dict[str, Any]]] = None  # [{hash, offset, length, data?}]
    prompt_embeds: Optional[Any] = None
    client_index: int = 0
    current_wave: int = 0
    trace_headers: Optional[dict[str, str]] = None


class LogprobsLists(msgspec.Struct, omit_defaults=True):
    """Flat logprob payload: per position top-k ids/logprobs plus sampled rank."""
    token_ids: list[list[int]]
    logprobs: list[list[float]]
    ranks: list[int]


class EngineCoreOutput(msgspec.Struct, omit_defaults=True):
    request_id: str
    new_token_ids: list[int]
    new_logprobs: Optional[LogprobsLists] = None
    new_prompt_logprobs: Optional[LogprobsLists] = None
    pooling_output: Optional[Any] = None
    finish_reason: Optional[int] = None
    stop_reason: Union[int, str, None] = None
    num_cached_tokens: int = 0
    kv_transfer_params: Optional[dict[str, Any]] = None
    events: Optional[list[tuple[int, float]]] = None  # (EngineCoreEventType, ts)

    @property
    def finished(self) -> bool:
        return self.finish_reason is not None


class SchedulerStats(msgspec.Struct, omit_defaults=True):
    num_running_reqs: int = 0
    num_waiting_reqs: int = 0
    kv_cache_usage: float = 0.0
    prefix_cache_queries: int = 0
    prefix_cache_hits: int = 0
    num_preemptions: int = 0
    step_counter: int = 0
    spec_num_drafts: int = 0
    spec_num_draft_tokens: int = 0
    spec_num_accepted_tokens: int = 0
    num_scheduled_tokens: int = 0
    model_forward_ms: float = 0.0
    schedule_ms: float = 0.0
    kv_connector_stats: Optional[dict[str, Any]] = None


class EngineCoreOutputs(msgspec.Struct, omit_defaults=True):
    engine_index: int = 0
    outputs: list[EngineCoreOutput] = []
    scheduler_stats: Optional[SchedulerStats] = None
    timestamp: float = 0.0
    utility_output: Optional[Any] = None  # (call_id, failure_message|None, result)
    finished_requests: Optional[list[str]] = None
    wave_complete: Optional[int] = None
    start_wave: Optional[int] = None


class EngineCoreEventType(enum.IntEnum):
    QUEUED = 1
    SCHEDULED = 2
    PREEMPTED = 3


class RequestStatus(enum.IntEnum):
    WAITING = 0
    WAITING_FOR_FSM = 1
    WAITING_FOR_REMOTE_KVS = 2
    RUNNING = 3
    PREEMPTED = 4
    # anything above PREEMPTED is finished
    FINISHED_STOPPED = 5
    FINISHED_LENGTH_CAPPED = 6
    FINISHED_ABORTED = 7
    FINISHED_IGNORED = 8

    @staticmethod
    def is_finished(s: "RequestStatus") -> bool:
        return s > RequestStatus.PREEMPTED

    @staticmethod
    def get_finished_reason(s: "RequestStatus") -> Optional[FinishReason]:
        return _FINISHED_REASON.get(s)


_FINISHED_REASON = {
    RequestStatus.FINISHED_STOPPED: FinishReason.STOP,
    RequestStatus.FINISHED_LENGTH_CAPPED: FinishReason.LENGTH,
    RequestStatus.FINISHED_ABORTED: FinishReason.ABORT,
    RequestStatus.FINISHED_IGNORED: FinishReason.LENGTH,
}


class Request:
    """Scheduler-side request state."""

    __slots__ = (
        "request_id", "client_index", "priority", "sampling_params", "pooling_params", "eos_token_id",
        "lora_request", "arrival_time", "status", "events", "stop_reason", "kv_transfer_params",
        "max_tokens", "prompt_token_ids", "num_prompt_tokens", "_output_token_ids", "_all_token_ids",
        "spec_token_ids", "num_computed_tokens", "num_cached_tokens", "num_output_placeholders",
        "num_preemptions", "block_hashes", "cache_salt", "mm_features", "structured_output_request",
        "use_structured_output", "num_nans_in_logits", "trace_headers", "_hash_fn", "_block_size",
        "current_wave", "succeeded_first_schedule_ts", "spec_accept_ema",
    )

    def __init__(self, request_id: str, prompt_token_ids: list[int],
                 sampling_params: Optional[SamplingParams] = None,
                 pooling_params: Optional[PoolingParams] = None, eos_token_id: Optional[int] = None,
                 client_index: int = 0, arrival_time: Optional[float] = None,
                 lora_request: Optional[LoRARequest] = None, cache_salt: Optional[str] = None,
                 priority: int = 0, kv_transfer_params: Optional[dict] = None,
                 mm_features: Optional[list] = None, structured_output_request: Any = None,
                 block_hasher=None, trace_headers=None) -> None:
        self.request_id = request_id
        self.client_index = client_index
        self.priority = priority
        self.sampling_params = sampling_params
        self.pooling_params = pooling_params
        self.eos_token_id = eos_token_id
        self.lora_request = lora_request
        self.arrival_time = arrival_time if arrival_time is not None else time.time()
        self.status = RequestStatus.WAITING
        self.events: list[tuple[int, float]] = []
        self.stop_reason: Union[int, str, None] = None
        self.kv_transfer_params = kv_transfer_params
        self.structured_output_request = structured_output_request
        self.use_structured_output = structured_output_request is not None
        if self.use_structured_output:
            self.status = RequestStatus.WAITING_FOR_FSM
        if sampling_params is not None:
            assert sampling_params.max_tokens is not None
            self.max_tokens = sampling_params.max_tokens
        else:
            self.max_tokens = 1
        self.prompt_token_ids = prompt_token_ids
        self.num_prompt_tokens = len(prompt_token_ids)
        self._output_token_ids: list[int] = []
        self._all_token_ids: list[int] = list(prompt_token_ids)
        self.spec_token_ids: list[int] = []
        self.spec_accept_ema = 1.0        # dynamic-K speculation: running acceptance rate of this request's drafts
        self.num_computed_tokens = 0
        self.num_cached_tokens = -1
        self.num_output_placeholders = 0  # async scheduling
        self.num_preemptions = 0
        self.cache_salt = cache_salt
        self.mm_features = mm_features or []
        self.num_nans_in_logits = 0
        self.trace_headers = trace_headers
        self.current_wave = 0
        self.succeeded_first_schedule_ts = 0.0
        self.block_hashes: list[bytes] = []
        self._hash_fn = block_hasher
        if block_hasher is not None:
            self.block_hashes = block_hasher(self)

    @classmethod
    def from_engine_core_request(cls, r: EngineCoreRequest, block_hasher=None,
                                 structured_output_request=None) -> "Request":
        return cls(
            request_id=r.request_id, prompt_token_ids=r.prompt_token_ids,
            sampling_params=r.sampling_params, pooling_params=r.pooling_params,
            eos_token_id=r.eos_token_id, client_index=r.client_index, arrival_time=r.arrival_time,
            lora_request=r.lora_request, cache_salt=r.cache_salt, priority=r.priority,
            kv_transfer_params=r.kv_transfer_params, mm_features=r.mm_features,
            structured_output_request=structured_output_request, block_hasher=block_hasher,
            trace_headers=r.trace_headers)

    # -- token bookkeeping --------------------------------------------------
    def append_output_token_ids(self, token_ids: Union[int, list[int]]) -> None:
        if isinstance(token_ids, int):
            self._output_token_ids.append(token_ids)
            self._all_token_ids.append(token_ids)
        else:
            self._output_token_ids.extend(token_ids)
            self._all_token_ids.extend(token_ids)
        if self._hash_fn is not None:
            self.block_hashes.extend(self._hash_fn(self))

    @property
    def output_token_ids(self) -> list[int]:
        return self._output_token_ids

    @property
    def all_token_ids(self) -> list[int]:
        return self._all_token_ids

    @property
    def num_tokens(self) -> int:
        return len(self._all_token_ids)

    @property
    def num_tokens_with_spec(self) -> int:
        return len(self._all_token_ids) + len(self.spec_token_ids)

    @property
    def num_output_tokens(self) -> int:
        return len(self._output_token_ids)

    @property
    def is_output_processed(self) -> bool:
        return True

    def is_finished(self) -> bool:
        return RequestStatus.is_finished(self.status)

    def get_finished_reason(self) -> Optional[FinishReason]:
        return RequestStatus.get_finished_reason(self.status)

    def record_event(self, ev: EngineCoreEventType, ts: Optional[float] = None) -> None:
        self.events.append((int(ev), ts if ts is not None else time.monotonic()))

    def take_events(self) -> Optional[list[tuple[int, float]]]:
        if not self.events:
            return None
        ev, self.events = self.events, []
        return ev

    def __lt__(self, other: "Request") -> bool:  # priority queue ordering
        return (self.priority, self.arrival_time, self.request_id) < (
            other.priority, other.arrival_time, other.request_id)

    def __repr__(self) -> str:
        return (f"Request({self.request_id}, status={self.status.name}, prompt={self.num_prompt_tokens}, "
                f"out={self.num_output_tokens}, computed={self.num_computed_tokens})")
