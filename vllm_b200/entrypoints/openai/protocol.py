"""OpenAI-compatible request / response schemas (reference: vllm/entrypoints/openai/*/protocol.py)."""
from __future__ import annotations

import time
import uuid
from typing import Any, Literal, Optional, Union

from pydantic import BaseModel, ConfigDict, Field

from vllm_b200.sampling_params import RequestOutputKind, SamplingParams, StructuredOutputsParams


def _id(prefix: str) -> str:
    return f"{prefix}-{uuid.uuid4().hex}"


class OpenAIBase(BaseModel):
    model_config = ConfigDict(extra="allow")


class ErrorInfo(OpenAIBase):
    message: str
    type: str = "BadRequestError"
    param: Optional[str] = None
    code: int = 400


class ErrorResponse(OpenAIBase):
    error: ErrorInfo


class ModelCard(OpenAIBase):
    id: str
    object: str = "model"
    created: int = Field(default_factory=lambda: int(time.time()))
    owned_by: str = "vllm_b200"
    root: Optional[str] = None
    parent: Optional[str] = None
    max_model_len: Optional[int] = None


class ModelList(OpenAIBase):
    object: str = "list"
    data: list[ModelCard] = Field(default_factory=list)


class UsageInfo(OpenAIBase):
    prompt_tokens: int = 0
    total_tokens: int = 0
    completion_tokens: Optional[int] = 0
    prompt_tokens_details: Optional[dict[str, int]] = None


class StreamOptions(OpenAIBase):
    include_usage: Optional[bool] = True
    continuous_usage_stats: Optional[bool] = False


class ResponseFormat(OpenAIBase):
    type: Literal["text", "json_object", "json_schema", "structural_tag"]
    json_schema: Optional[dict[str, Any]] = None


class _SamplingMixin(OpenAIBase):
    max_tokens: Optional[int] = None
    max_completion_tokens: Optional[int] = None
    n: int = 1
    temperature: Optional[float] = None
    top_p: Optional[float] = None
    top_k: Optional[int] = None
    min_p: Optional[float] = None
    presence_penalty: Optional[float] = 0.0
    frequency_penalty: Optional[float] = 0.0
    repetition_penalty: Optional[float] = None
    seed: Optional[int] = None
    stop: Optional[Union[str, list[str]]] = None
    stop_token_ids: Optional[list[int]] = None
    ignore_eos: bool = False
    min_tokens: int = 0
    skip_special_tokens: bool = True
    include_stop_str_in_output: bool = False
    logit_bias: Optional[dict[str, float]] = None
    allowed_token_ids: Optional[list[int]] = None
    bad_words: Optional[list[str]] = None
    response_format: Optional[ResponseFormat] = None
    structured_outputs: Optional[dict[str, Any]] = None
    guided_json: Optional[Union[str, dict]] = None
    guided_regex: Optional[str] = None
    guided_choice: Optional[list[str]] = None
    guided_grammar: Optional[str] = None
    priority: int = 0
    request_id: Optional[str] = None
    stream: bool = False
    stream_options: Optional[StreamOptions] = None
    truncate_prompt_tokens: Optional[int] = None
    kv_transfer_params: Optional[dict[str, Any]] = None
    vllm_xargs: Optional[dict[str, Any]] = None

    def _structured(self) -> Optional[StructuredOutputsParams]:
        if self.structured_outputs:
            return StructuredOutputsParams(**self.structured_outputs)
        if self.guided_json is not None:
            return StructuredOutputsParams(json=self.guided_json)
        if self.guided_regex is not None:
            return StructuredOutputsParams(regex=self.guided_regex)
        if self.guided_choice is not None:
            return StructuredOutputsParams(choice=self.guided_choice)
        if self.guided_grammar is not None:
            return StructuredOutputsParams(grammar=self.guided_grammar)
        rf = self.response_format
        if rf is not None:
            if rf.type == "json_object":
                return StructuredOutputsParams(json_object=True)
            if rf.type == "json_schema" and rf.json_schema:
                return StructuredOutputsParams(json=rf.json_schema.get("schema", rf.json_schema))
        return None

    def to_sampling_params(self, default_max_tokens: int, logprobs: Optional[int], prompt_logprobs: Optional[int],
                           defaults: Optional[dict] = None) -> SamplingParams:
        d = defaults or {}
        mt = self.max_completion_tokens if self.max_completion_tokens is not None else self.max_tokens
        if mt is None:
            mt = default_max_tokens
        extra = dict(self.vllm_xargs or {})
        if self.kv_transfer_params:
            extra["kv_transfer_params"] = self.kv_transfer_params
        return SamplingParams(
            n=self.n, presence_penalty=self.presence_penalty or 0.0, frequency_penalty=self.frequency_penalty or 0.0,
            repetition_penalty=self.repetition_penalty if self.repetition_penalty is not None else d.get("repetition_penalty", 1.0),
            temperature=self.temperature if self.temperature is not None else d.get("temperature", 1.0),
            top_p=self.top_p if self.top_p is not None else d.get("top_p", 1.0),
            top_k=self.top_k if self.top_k is not None else d.get("top_k", 0),
            min_p=self.min_p if self.min_p is not None else d.get("min_p", 0.0), seed=self.seed, stop=self.stop,
            stop_token_ids=self.stop_token_ids, ignore_eos=self.ignore_eos, max_tokens=mt, min_tokens=self.min_tokens,
            logprobs=logprobs, prompt_logprobs=prompt_logprobs, skip_special_tokens=self.skip_special_tokens,
            include_stop_str_in_output=self.include_stop_str_in_output,
            logit_bias={int(k): v for k, v in self.logit_bias.items()} if self.logit_bias else None,
            allowed_token_ids=self.allowed_token_ids, bad_words=self.bad_words,
            truncate_prompt_tokens=self.truncate_prompt_tokens, structured_outputs=self._structured(),
            output_kind=RequestOutputKind.DELTA if self.stream else RequestOutputKind.FINAL_ONLY,
            extra_args=extra or None)


class CompletionRequest(_SamplingMixin):
    model: Optional[str] = None
    prompt: Optional[Union[str, list[str], list[int], list[list[int]]]] = None
    echo: bool = False
    logprobs: Optional[int] = None
    prompt_logprobs: Optional[int] = None
    suffix: Optional[str] = None
    best_of: Optional[int] = None
    user: Optional[str] = None
    add_special_tokens: bool = True


class CompletionLogProbs(OpenAIBase):
    text_offset: list[int] = Field(default_factory=list)
    token_logprobs: list[Optional[float]] = Field(default_factory=list)
    tokens: list[str] = Field(default_factory=list)
    top_logprobs: list[Optional[dict[str, float]]] = Field(default_factory=list)


class CompletionChoice(OpenAIBase):
    index: int
    text: str
    logprobs: Optional[CompletionLogProbs] = None
    finish_reason: Optional[str] = None
    stop_reason: Optional[Union[int, str]] = None
    prompt_logprobs: Optional[list[Optional[dict[int, Any]]]] = None
    token_ids: Optional[list[int]] = None


class CompletionResponse(OpenAIBase):
    id: str = Field(default_factory=lambda: _id("cmpl"))
    object: str = "text_completion"
    created: int = Field(default_factory=lambda: int(time.time()))
    model: str
    choices: list[CompletionChoice]
    usage: Optional[UsageInfo] = None
    kv_transfer_params: Optional[dict[str, Any]] = None


class CompletionStreamResponse(OpenAIBase):
    id: str
    object: str = "text_completion"
    created: int = Field(default_factory=lambda: int(time.time()))
    model: str
    choices: list[CompletionChoice]
    usage: Optional[UsageInfo] = None


class FunctionCall(OpenAIBase):
    name: Optional[str] = None
    arguments: Optional[str] = None


class ToolCall(OpenAIBase):
    id: str = Field(default_factory=lambda: _id("chatcmpl-tool"))
    type: Literal["function"] = "function"
    function: FunctionCall
    index: Optional[int] = None


class ChatMessage(OpenAIBase):
    role: str
    content: Optional[str] = None
    reasoning_content: Optional[str] = None
    tool_calls: Optional[list[ToolCall]] = None


class ChatCompletionRequest(_SamplingMixin):
    model: Optional[str] = None
    messages: list[dict[str, Any]]
    logprobs: Optional[bool] = False
    top_logprobs: Optional[int] = 0
    prompt_logprobs: Optional[int] = None
    tools: Optional[list[dict[str, Any]]] = None
    tool_choice: Optional[Union[str, dict[str, Any]]] = "none"
    parallel_tool_calls: Optional[bool] = True
    user: Optional[str] = None
    chat_template: Optional[str] = None
    chat_template_kwargs: Optional[dict[str, Any]] = None
    add_generation_prompt: bool = True
    continue_final_message: bool = False
    echo: bool = False
    include_reasoning: bool = True


class ChatLogprob(OpenAIBase):
    token: str
    logprob: float = -9999.0
    bytes: Optional[list[int]] = None
    top_logprobs: list[dict[str, Any]] = Field(default_factory=list)


class ChatLogprobs(OpenAIBase):
    content: Optional[list[ChatLogprob]] = None


class ChatChoice(OpenAIBase):
    index: int
    message: ChatMessage
    logprobs: Optional[ChatLogprobs] = None
    finish_reason: Optional[str] = "stop"
    stop_reason: Optional[Union[int, str]] = None


class ChatCompletionResponse(OpenAIBase):
    id: str = Field(default_factory=lambda: _id("chatcmpl"))
    object: str = "chat.completion"
    created: int = Field(default_factory=lambda: int(time.time()))
    model: str
    choices: list[ChatChoice]
    usage: UsageInfo
    prompt_logprobs: Optional[list[Optional[dict[int, Any]]]] = None
    kv_transfer_params: Optional[dict[str, Any]] = None


class DeltaMessage(OpenAIBase):
    role: Optional[str] = None
    content: Optional[str] = None
    reasoning_content: Optional[str] = None
    tool_calls: Optional[list[ToolCall]] = None


class ChatStreamChoice(OpenAIBase):
    index: int
    delta: DeltaMessage
    logprobs: Optional[ChatLogprobs] = None
    finish_reason: Optional[str] = None
    stop_reason: Optional[Union[int, str]] = None


class ChatCompletionStreamResponse(OpenAIBase):
    id: str
    object: str = "chat.completion.chunk"
    created: int = Field(default_factory=lambda: int(time.time()))
    model: str
    choices: list[ChatStreamChoice]
    usage: Optional[UsageInfo] = None


class EmbeddingRequest(OpenAIBase):
    model: Optional[str] = None
    input: Union[str, list[str], list[int], list[list[int]]]
    encoding_format: Literal["float", "base64"] = "float"
    dimensions: Optional[int] = None
    user: Optional[str] = None
    truncate_prompt_tokens: Optional[int] = None
    normalize: bool = True


class EmbeddingData(OpenAIBase):
    index: int
    object: str = "embedding"
    embedding: Union[list[float], str]


class EmbeddingResponse(OpenAIBase):
    id: str = Field(default_factory=lambda: _id("embd"))
    object: str = "list"
    created: int = Field(default_factory=lambda: int(time.time()))
    model: str
    data: list[EmbeddingData]
    usage: UsageInfo


class TokenizeRequest(OpenAIBase):
    model: Optional[str] = None
    prompt: Optional[str] = None
    messages: Optional[list[dict[str, Any]]] = None
    add_special_tokens: bool = True
    add_generation_prompt: bool = True
    return_token_strs: bool = False


class TokenizeResponse(OpenAIBase):
    count: int
    max_model_len: int
    tokens: list[int]
    token_strs: Optional[list[str]] = None


class DetokenizeRequest(OpenAIBase):
    model: Optional[str] = None
    tokens: list[int]


class DetokenizeResponse(OpenAIBase):
    prompt: str


class ScoreRequest(OpenAIBase):
    model: Optional[str] = None
    text_1: Union[str, list[str]]
    text_2: Union[str, list[str]]


class ScoreResponse(OpenAIBase):
    id: str = Field(default_factory=lambda: _id("score"))
    object: str = "list"
    model: str
    data: list[dict[str, Any]]
    usage: UsageInfo


class ResponsesRequest(OpenAIBase):
    """Subset of the OpenAI Responses API (reference: entrypoints/openai/responses/)."""
    model: Optional[str] = None
    input: Union[str, list[dict[str, Any]]]
    instructions: Optional[str] = None
    max_output_tokens: Optional[int] = None
    temperature: Optional[float] = None
    top_p: Optional[float] = None
    stream: bool = False
    tools: Optional[list[dict[str, Any]]] = None
    tool_choice: Optional[Union[str, dict[str, Any]]] = None
    max_tool_calls: Optional[int] = None
    previous_response_id: Optional[str] = None
    store: bool = False


class AnthropicMessagesRequest(OpenAIBase):
    """Anthropic /v1/messages (reference: entrypoints/anthropic/)."""
    model: Optional[str] = None
    stream: bool = False
    messages: list[dict[str, Any]]
    system: Optional[Union[str, list[dict[str, Any]]]] = None
    max_tokens: int = 1024
    temperature: Optional[float] = None
    top_p: Optional[float] = None
    top_k: Optional[int] = None
    stop_sequences: Optional[list[str]] = None
    stream: bool = False
    tools: Optional[list[dict[str, Any]]] = None
