"""Scheduler → worker step description and worker → scheduler result.

Reference: vllm/v1/core/sched/output.py (``SchedulerOutput`` :207,
``NewRequestData`` :36, ``CachedRequestData`` :130), vllm/v1/outputs.py
(``ModelRunnerOutput``).  Plain dataclasses (pickled over the shm ring to the
workers).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Optional


@dataclass
class NewRequestData:
    req_id: str
    prompt_token_ids: list[int]
    sampling_params: Any
    pooling_params: Any
    block_ids: tuple[list[int], ...]
    num_computed_tokens: int
    lora_request: Any = None
    mm_features: Optional[list] = None


@dataclass
class CachedRequestData:
    """Diffs for requests the workers already know."""
    req_ids: list[str] = field(default_factory=list)
    resumed_from_preemption: list[bool] = field(default_factory=list)
    # tokens the worker does not have yet (only when it can't derive them itself: PP, resumed)
    new_token_ids: list[list[int]] = field(default_factory=list)
    new_block_ids: list[Optional[tuple[list[int], ...]]] = field(default_factory=list)
    num_computed_tokens: list[int] = field(default_factory=list)
    num_output_tokens: list[int] = field(default_factory=list)

    @property
    def num_reqs(self) -> int:
        return len(self.req_ids)


@dataclass
class SchedulerOutput:
    scheduled_new_reqs: list[NewRequestData] = field(default_factory=list)
    scheduled_cached_reqs: CachedRequestData = field(default_factory=CachedRequestData)
    num_scheduled_tokens: dict[str, int] = field(default_factory=dict)
    total_num_scheduled_tokens: int = 0
    scheduled_spec_decode_tokens: dict[str, list[int]] = field(default_factory=dict)
    scheduled_encoder_inputs: dict[str, list[int]] = field(default_factory=dict)
    num_common_prefix_blocks: list[int] = field(default_factory=list)
    finished_req_ids: set[str] = field(default_factory=set)
    free_encoder_mm_hashes: list[str] = field(default_factory=list)
    structured_output_request_ids: list[str] = field(default_factory=list)
    grammar_bitmask: Any = None  # np.ndarray [n_structured, ceil(V/32)] int32
    kv_connector_metadata: Any = None
    preempted_req_ids: set[str] = field(default_factory=set)
    step_id: int = 0

    @property
    def is_empty(self) -> bool:
        return self.total_num_scheduled_tokens == 0 and not self.finished_req_ids


@dataclass
class LogprobsTensors:
    """Per sampled position: top-k ids / logprobs (col 0 = the sampled token) and rank."""
    token_ids: Any  # list[list[int]] or tensor [n, k+1]
    logprobs: Any
    ranks: Any

    def tolists(self):
        def tl(x):
            return x.tolist() if hasattr(x, "tolist") else x
        return tl(self.token_ids), tl(self.logprobs), tl(self.ranks)


@dataclass
class ModelRunnerOutput:
    req_ids: list[str] = field(default_factory=list)
    req_id_to_index: dict[str, int] = field(default_factory=dict)
    sampled_token_ids: list[list[int]] = field(default_factory=list)  # per request, ≥1 with spec decode
    logprobs: Optional[LogprobsTensors] = None
    prompt_logprobs_dict: dict[str, Optional[LogprobsTensors]] = field(default_factory=dict)
    pooler_output: list[Any] = field(default_factory=list)
    spec_token_ids: Optional[list[list[int]]] = None  # drafts for the next step
    kv_connector_output: Any = None
    num_nans_in_logits: Optional[dict[str, int]] = None
    forward_ms: float = 0.0


EMPTY_MODEL_RUNNER_OUTPUT = ModelRunnerOutput()
