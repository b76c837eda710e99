"""Per-request sampling parameters (reference: vllm/sampling_params.py:215-1245).

A msgspec Struct so it serialises over the engine-core IPC channel with no
custom hooks.
"""
from __future__ import annotations

import copy
from enum import IntEnum
from typing import Any, Optional, Union

import msgspec

_SAMPLING_EPS = 1e-5
_MAX_TEMP = 1e-2


class SamplingType(IntEnum):
    GREEDY = 0
    RANDOM = 1
    RANDOM_SEED = 2


class RequestOutputKind(IntEnum):
    CUMULATIVE = 0  # every output carries the whole text so far
    DELTA = 1       # every output carries only the new text
    FINAL_ONLY = 2  # a single output at the end


class StructuredOutputsParams(msgspec.Struct, omit_defaults=True):
    json: Optional[Union[str, dict]] = None
    regex: Optional[str] = None
    choice: Optional[list[str]] = None
    grammar: Optional[str] = None
    json_object: Optional[bool] = None
    disable_fallback: bool = False
    whitespace_pattern: Optional[str] = None

    def kind(self) -> str:
        for k in ("json", "regex", "choice", "grammar"):
            if getattr(self, k) is not None:
                return k
        if self.json_object:
            return "json_object"
        raise ValueError("empty StructuredOutputsParams")


class SamplingParams(msgspec.Struct, omit_defaults=True, dict=True):
    n: int = 1
    best_of: Optional[int] = None
    presence_penalty: float = 0.0
    frequency_penalty: float = 0.0
    repetition_penalty: float = 1.0
    temperature: float = 1.0
    top_p: float = 1.0
    top_k: int = 0
    min_p: float = 0.0
    seed: Optional[int] = None
    stop: Optional[Union[str, list[str]]] = None
    stop_token_ids: Optional[list[int]] = None
    ignore_eos: bool = False
    max_tokens: Optional[int] = 16
    min_tokens: int = 0
    logprobs: Optional[int] = None
    prompt_logprobs: Optional[int] = None
    detokenize: bool = True
    skip_special_tokens: bool = True
    spaces_between_special_tokens: bool = True
    include_stop_str_in_output: bool = False
    output_kind: RequestOutputKind = RequestOutputKind.CUMULATIVE
    logit_bias: Optional[dict[int, float]] = None
    allowed_token_ids: Optional[list[int]] = None
    bad_words: Optional[list[str]] = None
    bad_words_token_ids: Optional[list[list[int]]] = None
    truncate_prompt_tokens: Optional[int] = None
    structured_outputs: Optional[StructuredOutputsParams] = None
    extra_args: Optional[dict[str, Any]] = None
    # filled by the input processor
    all_stop_token_ids: list[int] = []
    output_text_buffer_length: int = 0

    def __post_init__(self) -> None:
        if isinstance(self.stop, str):
            self.stop = [self.stop]
        elif self.stop is None:
            self.stop = []
        if self.stop_token_ids is None:
            self.stop_token_ids = []
        if 0 < self.temperature < _MAX_TEMP:
            self.temperature = _MAX_TEMP
        if self.top_k is None or self.top_k < 0:
            self.top_k = 0
        if self.temperature < _SAMPLING_EPS:
            # greedy: neutralise the filters so the sampler can skip them
            self.top_p, self.top_k, self.min_p = 1.0, 0, 0.0
        self._verify()
        if self.stop and not self.include_stop_str_in_output:
            self.output_text_buffer_length = max(len(s) for s in self.stop) - 1
        self.all_stop_token_ids = list(self.stop_token_ids)

    def _verify(self) -> None:
        if self.n < 1:
            raise ValueError(f"n must be at least 1, got {self.n}")
        if self.best_of is not None and self.best_of < self.n:
            raise ValueError("best_of must be >= n")
        for name in ("presence_penalty", "frequency_penalty"):
            if not -2.0 <= getattr(self, name) <= 2.0:
                raise ValueError(f"{name} must be in [-2, 2], got {getattr(self, name)}")
        if not 0.0 < self.repetition_penalty <= 2.0:
            raise ValueError(f"repetition_penalty must be in (0, 2], got {self.repetition_penalty}")
        if self.temperature < 0.0:
            raise ValueError(f"temperature must be non-negative, got {self.temperature}")
        if not 0.0 < self.top_p <= 1.0:
            raise ValueError(f"top_p must be in (0, 1], got {self.top_p}")
        if not 0.0 <= self.min_p <= 1.0:
            raise ValueError(f"min_p must be in [0, 1], got {self.min_p}")
        if self.max_tokens is not None and self.max_tokens < 1:
            raise ValueError(f"max_tokens must be at least 1, got {self.max_tokens}")
        if self.min_tokens < 0:
            raise ValueError("min_tokens must be >= 0")
        if self.max_tokens is not None and self.min_tokens > self.max_tokens:
            raise ValueError("min_tokens must be <= max_tokens")
        if self.logprobs is not None and self.logprobs < -1:
            raise ValueError("logprobs must be >= -1")
        if self.prompt_logprobs is not None and self.prompt_logprobs < -1:
            raise ValueError("prompt_logprobs must be >= -1")
        if self.stop and not self.detokenize:
            raise ValueError("stop strings require detokenize=True")
        if any(not s for s in self.stop or []):
            raise ValueError("stop cannot contain an empty string")
        if self.truncate_prompt_tokens is not None and self.truncate_prompt_tokens < -1:
            raise ValueError("truncate_prompt_tokens must be >= -1")

    @property
    def sampling_type(self) -> SamplingType:
        if self.temperature < _SAMPLING_EPS:
            return SamplingType.GREEDY
        if self.seed is not None:
            return SamplingType.RANDOM_SEED
        return SamplingType.RANDOM

    @property
    def needs_penalties(self) -> bool:
        return (abs(self.presence_penalty) >= _SAMPLING_EPS or abs(self.frequency_penalty) >= _SAMPLING_EPS
                or abs(self.repetition_penalty - 1.0) >= _SAMPLING_EPS)

    def update_from_generation_config(self, eos_token_id: Optional[int]) -> None:
        if eos_token_id is not None and not self.ignore_eos:
            if eos_token_id not in self.all_stop_token_ids:
                self.all_stop_token_ids.append(eos_token_id)

    def update_from_tokenizer(self, tokenizer) -> None:
        """Resolve ``bad_words`` strings to token-id sequences."""
        if not self.bad_words or tokenizer is None:
            return
        out: list[list[int]] = list(self.bad_words_token_ids or [])
        for w in self.bad_words:
            for prefix in ("", " "):
                ids = tokenizer.encode(prefix + w.lstrip(), add_special_tokens=False)
                if ids and ids not in out:
                    out.append(ids)
        self.bad_words_token_ids = out

    def clone(self) -> "SamplingParams":
        return copy.deepcopy(self)


class BeamSearchParams(msgspec.Struct, omit_defaults=True):
    beam_width: int = 4
    max_tokens: int = 16
    ignore_eos: bool = False
    temperature: float = 0.0
    length_penalty: float = 1.0
    include_stop_str_in_output: bool = False


class PoolingParams(msgspec.Struct, omit_defaults=True):
    """Reference: vllm/pooling_params.py."""
    task: str = "embed"  # embed|classify|score|token_embed
    dimensions: Optional[int] = None
    normalize: bool = True
    truncate_prompt_tokens: Optional[int] = None
    pooling_type: Optional[str] = None  # LAST|MEAN|CLS|ALL; None → the model's default (LAST for decoders)
    extra_args: Optional[dict[str, Any]] = None

    def clone(self) -> "PoolingParams":
        return copy.deepcopy(self)
