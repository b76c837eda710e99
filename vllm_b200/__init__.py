"""vllm_b200 — a B200-native (sm_100a) LLM serving engine with vLLM's capabilities."""
from vllm_b200.version import __version__

_LAZY = {
    "LLM": "vllm_b200.entrypoints.llm",
    "SamplingParams": "vllm_b200.sampling_params",
    "PoolingParams": "vllm_b200.sampling_params",
    "BeamSearchParams": "vllm_b200.sampling_params",
    "StructuredOutputsParams": "vllm_b200.sampling_params",
    "RequestOutput": "vllm_b200.outputs",
    "CompletionOutput": "vllm_b200.outputs",
    "PoolingRequestOutput": "vllm_b200.outputs",
    "EmbeddingRequestOutput": "vllm_b200.outputs",
    "EngineArgs": "vllm_b200.engine.arg_utils",
    "AsyncEngineArgs": "vllm_b200.engine.arg_utils",
    "LLMEngine": "vllm_b200.engine.llm_engine",
    "AsyncLLM": "vllm_b200.engine.async_llm",
    "AsyncLLMEngine": "vllm_b200.engine.async_llm",
    "TextPrompt": "vllm_b200.inputs",
    "TokensPrompt": "vllm_b200.inputs",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        return getattr(importlib.import_module(_LAZY[name]), name)
    raise AttributeError(name)


__all__ = ["__version__", *_LAZY]
