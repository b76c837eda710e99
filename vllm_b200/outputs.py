"""User-facing output objects (reference: vllm/outputs.py)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Optional, Union


@dataclass
class Logprob:
    logprob: float
    rank: Optional[int] = None
    decoded_token: Optional[str] = None


SampleLogprobs = list[dict[int, Logprob]]
PromptLogprobs = list[Optional[dict[int, Logprob]]]


@dataclass
class CompletionOutput:
    index: int
    text: str
    token_ids: list[int]
    cumulative_logprob: Optional[float] = None
    logprobs: Optional[SampleLogprobs] = None
    finish_reason: Optional[str] = None  # stop|length|abort
    stop_reason: Union[int, str, None] = None
    lora_request: Any = None

    def finished(self) -> bool:
        return self.finish_reason is not None


@dataclass
class RequestMetrics:
    arrival_time: float = 0.0
    queued_ts: float = 0.0
    scheduled_ts: float = 0.0
    first_token_ts: float = 0.0
    last_token_ts: float = 0.0
    finished_time: Optional[float] = None
    num_generation_tokens: int = 0
    num_cached_tokens: int = 0

    @property
    def time_to_first_token(self) -> float:
        return self.first_token_ts - self.arrival_time if self.first_token_ts else 0.0


@dataclass
class RequestOutput:
    request_id: str
    prompt: Optional[str]
    prompt_token_ids: Optional[list[int]]
    prompt_logprobs: Optional[PromptLogprobs]
    outputs: list[CompletionOutput]
    finished: bool
    metrics: Optional[RequestMetrics] = None
    num_cached_tokens: Optional[int] = None
    kv_transfer_params: Optional[dict[str, Any]] = None

    def add(self, nxt: "RequestOutput", aggregate: bool) -> None:
        """Merge a later delta into this output (used when a consumer is slower
        than the producer; cf. RequestOutputCollector in the reference)."""
        self.finished |= nxt.finished
        self.kv_transfer_params = nxt.kv_transfer_params
        for n in nxt.outputs:
            for c in self.outputs:
                if c.index == n.index:
                    if aggregate:
                        c.text += n.text
                        c.token_ids = list(c.token_ids) + list(n.token_ids)
                        if n.logprobs:
                            c.logprobs = (c.logprobs or []) + n.logprobs
                        c.cumulative_logprob = n.cumulative_logprob
                    else:
                        c.text, c.token_ids, c.logprobs = n.text, n.token_ids, n.logprobs
                        c.cumulative_logprob = n.cumulative_logprob
                    c.finish_reason, c.stop_reason = n.finish_reason, n.stop_reason
                    break
            else:
                self.outputs.append(n)


@dataclass
class PoolingOutput:
    data: Any  # torch.Tensor


@dataclass
class PoolingRequestOutput:
    request_id: str
    outputs: PoolingOutput
    prompt_token_ids: list[int]
    finished: bool = True


@dataclass
class EmbeddingRequestOutput(PoolingRequestOutput):
    @property
    def embedding(self) -> list[float]:
        return self.outputs.data.tolist()


@dataclass
class BeamSearchSequence:
    tokens: list[int]
    cum_logprob: float = 0.0
    text: Optional[str] = None
    finish_reason: Optional[str] = None
    logprobs: list[dict[int, Logprob]] = field(default_factory=list)


@dataclass
class BeamSearchOutput:
    sequences: list[BeamSearchSequence]
