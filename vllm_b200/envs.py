"""Typed, lazily evaluated environment-variable registry.

Equivalent in role to the reference's ``vllm/envs.py`` (lazy ``__getattr__``
registry with an optional cache, cf. envs.py:560-2300): every variable is
declared once with a parser and a default; ``envs.NAME`` evaluates it on
access; ``enable_envs_cache()`` freezes values after engine start-up so the
hot path never touches ``os.environ``.
"""
from __future__ import annotations

import os
from typing import Any, Callable

_cache_enabled = False
_cache: dict[str, Any] = {}


def _bool(name: str, default: bool) -> Callable[[], bool]:
    return lambda: os.environ.get(name, "1" if default else "0").lower() in ("1", "true", "yes", "on")


def _int(name: str, default: int) -> Callable[[], int]:
    return lambda: int(os.environ.get(name, default))


def _float(name: str, default: float) -> Callable[[], float]:
    return lambda: float(os.environ.get(name, default))


def _str(name: str, default: str | None) -> Callable[[], str | None]:
    return lambda: os.environ.get(name, default)


environment_variables: dict[str, Callable[[], Any]] = {
    # --- logging / debug
    "VLLM_LOGGING_LEVEL": _str("VLLM_LOGGING_LEVEL", "INFO"),
    "VLLM_LOG_STATS_INTERVAL": _float("VLLM_LOG_STATS_INTERVAL", 10.0),
    "VLLM_TRACE_FUNCTION": _bool("VLLM_TRACE_FUNCTION", False),
    "VLLM_GPU_SYNC_CHECK": _str("VLLM_GPU_SYNC_CHECK", None),  # warn|error
    "VLLM_COMPUTE_NANS_IN_LOGITS": _bool("VLLM_COMPUTE_NANS_IN_LOGITS", False),
    "VLLM_RAISE_ON_LOGIT_NANS": _bool("VLLM_RAISE_ON_LOGIT_NANS", False),
    "VLLM_DEBUG_WORKSPACE": _bool("VLLM_DEBUG_WORKSPACE", False),
    # --- engine topology
    "VLLM_ENABLE_MULTIPROCESSING": _bool("VLLM_ENABLE_MULTIPROCESSING", False),
    "VLLM_WORKER_MULTIPROC_METHOD": _str("VLLM_WORKER_MULTIPROC_METHOD", "spawn"),
    "VLLM_EXECUTE_MODEL_TIMEOUT_SECONDS": _int("VLLM_EXECUTE_MODEL_TIMEOUT_SECONDS", 300),
    "VLLM_OUTPUT_PROC_CHUNK_SIZE": _int("VLLM_OUTPUT_PROC_CHUNK_SIZE", 128),
    "VLLM_RPC_BASE_PATH": _str("VLLM_RPC_BASE_PATH", "/tmp"),
    "VLLM_HOST_IP": _str("VLLM_HOST_IP", "127.0.0.1"),
    "VLLM_PORT": _int("VLLM_PORT", 0),
    "VLLM_API_KEY": _str("VLLM_API_KEY", None),
    # admin / debug HTTP routes (sleep, profile, cache reset, collective_rpc, checkpoint, elastic scale) are only
    # mounted in dev mode; runtime LoRA loading is its own opt-in (both as in the reference's envs.py)
    "VLLM_SERVER_DEV_MODE": _bool("VLLM_SERVER_DEV_MODE", False),
    "VLLM_SERVER_DEV_RPC_METHODS": _str("VLLM_SERVER_DEV_RPC_METHODS", ""),
    "VLLM_ALLOW_RUNTIME_LORA_UPDATING": _bool("VLLM_ALLOW_RUNTIME_LORA_UPDATING", False),
    # --- kernels / device
    "VLLM_B200_FORCE_REF_OPS": _bool("VLLM_B200_FORCE_REF_OPS", False),
    "VLLM_B200_GEMM": _str("VLLM_B200_GEMM", "auto"),  # auto|cublas|tcgen05
    # decode: pull the next few-tile GEMM's weights into L2 on a side stream while the HBM-bound attention kernel runs
    "VLLM_B200_L2_PREFETCH": _bool("VLLM_B200_L2_PREFETCH", False),
    # decode steps: pull up to this many MiB of the NEXT layer's weights into L2 on a side stream while the current layer
    # runs (0 = off).  Pays when a layer's kernels are latency-bound and HBM is mostly idle (high TP degree, small models)
    "VLLM_B200_LAYER_PREFETCH_MB": lambda: int(os.environ.get("VLLM_B200_LAYER_PREFETCH_MB", "0")),
    # device-resident token / block tables + on-device expansion of the step inputs (csrc/input_prep.cu)
    "VLLM_B200_GPU_INPUT_PREP": _bool("VLLM_B200_GPU_INPUT_PREP", True),
    "VLLM_B200_FUSED_MOE": _bool("VLLM_B200_FUSED_MOE", True),   # MoE combine + reduce in the w2 grouped-GEMM epilogue
    "VLLM_B200_FUSED_ACT": _bool("VLLM_B200_FUSED_ACT", True),   # gate|up as interleaved rows, SiLU·mul in the GEMM epilogue
    "VLLM_B200_GEMM_SK": _bool("VLLM_B200_GEMM_SK", False),  # decode GEMMs on the persistent stream-K kernel (measured slower than the
    # cluster split-K kernel on every decode shape: profiles/kernels_bench.md; kept for experiments)
    "VLLM_B200_ATTN_SPLITS": _int("VLLM_B200_ATTN_SPLITS", 0),
    "VLLM_B200_ATTN": _str("VLLM_B200_ATTN", "auto"),  # auto|simt (disable the tcgen05 attention kernel)
    "VLLM_B200_FUSED_AR": _bool("VLLM_B200_FUSED_AR", True),
    # fused GEMM+all-reduce algorithm: oneshot (every rank reduces everything), twoshot (row owners reduce, then
    # multicast; faster at every world size measured), auto = twoshot
    "VLLM_B200_FUSED_AR_MODE": lambda: os.environ.get("VLLM_B200_FUSED_AR_MODE", "auto"),
    "VLLM_B200_AR_BACKEND": _str("VLLM_B200_AR_BACKEND", "auto"),  # auto|nvls|p2p|nccl
    "VLLM_DISABLE_CUDA_GRAPH": _bool("VLLM_DISABLE_CUDA_GRAPH", False),
    "VLLM_BATCH_INVARIANT": _bool("VLLM_BATCH_INVARIANT", False),
    "VLLM_PP_LAYER_PARTITION": _str("VLLM_PP_LAYER_PARTITION", None),
    "VLLM_SLEEP_WHEN_IDLE": _bool("VLLM_SLEEP_WHEN_IDLE", False),
    # --- caches / paths
    "VLLM_CACHE_ROOT": _str("VLLM_CACHE_ROOT", os.path.expanduser("~/.cache/vllm_b200")),
    "VLLM_ALLOW_LONG_MAX_MODEL_LEN": _bool("VLLM_ALLOW_LONG_MAX_MODEL_LEN", False),
    "VLLM_NO_USAGE_STATS": _bool("VLLM_NO_USAGE_STATS", True),
    # --- profiling
    "VLLM_TORCH_PROFILER_DIR": _str("VLLM_TORCH_PROFILER_DIR", None),
    "VLLM_NVTX_SCOPES_FOR_PROFILING": _bool("VLLM_NVTX_SCOPES_FOR_PROFILING", False),
    "VLLM_DEBUG_MFU_METRICS": _bool("VLLM_DEBUG_MFU_METRICS", False),
}


def __getattr__(name: str) -> Any:
    if name in environment_variables:
        if _cache_enabled and name in _cache:
            return _cache[name]
        value = environment_variables[name]()
        if _cache_enabled:
            _cache[name] = value
        return value
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


def __dir__() -> list[str]:
    return list(environment_variables)


def enable_envs_cache() -> None:
    global _cache_enabled
    _cache_enabled = True
    for k, fn in environment_variables.items():
        _cache[k] = fn()


def disable_envs_cache() -> None:
    global _cache_enabled
    _cache_enabled = False
    _cache.clear()


def is_set(name: str) -> bool:
    return name in os.environ
