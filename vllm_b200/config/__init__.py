"""Typed configuration tree.

Role-equivalent of the reference's ``vllm/config/*`` (``VllmConfig`` aggregating
``ModelConfig``/``CacheConfig``/``ParallelConfig``/``SchedulerConfig``/... —
vllm/config/vllm.py, config/parallel.py:122-420, config/cache.py:111-138,
config/scheduler.py:74-99).  Plain dataclasses with ``__post_init__``
validation and a ``compute_hash`` used to key caches.  Defaults are sized for a
B200 (180 GB HBM3e): larger KV blocks, 16k batched tokens, 0.92 utilisation.
"""
from __future__ import annotations

import contextlib
import dataclasses
import hashlib
import json
from dataclasses import dataclass, field
from typing import Any, Literal, Optional

from vllm_b200.config.model import ModelConfig, get_preset_hf_config, PRESETS  # noqa: F401

# ---------------------------------------------------------------------------


def _hash_dict(d: dict) -> str:
    return hashlib.sha256(json.dumps(d, sort_keys=True, default=str).encode()).hexdigest()[:16]


class _ConfigBase:
    def to_dict(self) -> dict[str, Any]:
        return dataclasses.asdict(self)  # type: ignore[arg-type]

    def compute_hash(self) -> str:
        return _hash_dict(self.to_dict())

    def replace(self, **kw):
        return dataclasses.replace(self, **kw)  # type: ignore[type-var]


@dataclass
class CacheConfig(_ConfigBase):
    """Paged-KV configuration (reference: vllm/config/cache.py)."""

    block_size: int = 64  # tokens per KV page: a K page [64,128] / V^T page [128,64] is one 16 KB SW128 TMA box
    gpu_memory_utilization: float = 0.92
    swap_space_gb: float = 4.0
    cache_dtype: str = "auto"  # auto|bfloat16|fp8|fp8_e4m3|fp8_e5m2
    num_gpu_blocks_override: Optional[int] = None
    enable_prefix_caching: bool = True
    prefix_caching_hash_algo: str = "sha256"  # sha256|builtin
    sliding_window: Optional[int] = None
    kv_cache_memory_bytes: Optional[int] = None
    cpu_offload_gb: float = 0.0
    kv_sharing_fast_prefill: bool = False
    # filled in after profiling
    num_gpu_blocks: Optional[int] = None
    num_cpu_blocks: Optional[int] = None

    def __post_init__(self) -> None:
        if not (0.0 < self.gpu_memory_utilization <= 1.0):
            raise ValueError(f"gpu_memory_utilization must be in (0,1], got {self.gpu_memory_utilization}")
        if self.block_size <= 0 or self.block_size & (self.block_size - 1):
            raise ValueError(f"block_size must be a positive power of two, got {self.block_size}")
        if self.cache_dtype not in ("auto", "bfloat16", "float16", "fp8", "fp8_e4m3", "fp8_e5m2"):
            raise ValueError(f"unknown cache_dtype {self.cache_dtype}")


@dataclass
class ParallelConfig(_ConfigBase):
    """Parallel layout (reference: vllm/config/parallel.py:122-420).

    Rank layout is DP x PP x TP (TP fastest); EP reuses DP x TP ranks; DCP
    reuses TP ranks (parallel_state.py:1817-1955 in the reference).
    """

    tensor_parallel_size: int = 1
    pipeline_parallel_size: int = 1
    data_parallel_size: int = 1
    data_parallel_rank: int = 0
    data_parallel_size_local: Optional[int] = None
    data_parallel_master_ip: str = "127.0.0.1"
    data_parallel_master_port: int = 29550
    data_parallel_backend: str = "mp"
    data_parallel_external_lb: bool = False
    data_parallel_hybrid_lb: bool = False
    enable_expert_parallel: bool = False
    all2all_backend: str = "a2a"  # a2a (token all-to-all) | allgather_reducescatter | naive   (wide EP across DP replicas)
    enable_eplb: bool = False
    eplb_window_size: int = 1000
    eplb_step_interval: int = 3000
    num_redundant_experts: int = 0
    enable_dbo: bool = False
    decode_context_parallel_size: int = 1
    prefill_context_parallel_size: int = 1
    dcp_comm_backend: str = "ag_rs"  # ag_rs|a2a
    cp_kv_cache_interleave_size: int = 1
    disable_custom_all_reduce: bool = False
    distributed_executor_backend: Optional[str] = None  # uni|mp|external_launcher
    worker_cls: str = "auto"
    worker_extension_cls: str = ""
    max_parallel_loading_workers: Optional[int] = None
    enable_elastic_ep: bool = False
    rank: int = 0
    nnodes: int = 1
    node_rank: int = 0
    master_addr: str = "127.0.0.1"
    master_port: int = 29500

    def __post_init__(self) -> None:
        for name in ("tensor_parallel_size", "pipeline_parallel_size", "data_parallel_size",
                     "decode_context_parallel_size", "prefill_context_parallel_size"):
            if getattr(self, name) < 1:
                raise ValueError(f"{name} must be >= 1")
        if self.tensor_parallel_size % self.decode_context_parallel_size != 0:
            raise ValueError("decode_context_parallel_size must divide tensor_parallel_size")
        if self.distributed_executor_backend is None:
            self.distributed_executor_backend = "mp" if self.world_size > 1 else "uni"
        if self.data_parallel_size_local is None:
            self.data_parallel_size_local = self.data_parallel_size

    @property
    def world_size(self) -> int:
        return self.tensor_parallel_size * self.pipeline_parallel_size * self.prefill_context_parallel_size

    @property
    def world_size_across_dp(self) -> int:
        return self.world_size * self.data_parallel_size

    @property
    def use_ep(self) -> bool:
        return self.enable_expert_parallel and self.tensor_parallel_size * self.data_parallel_size > 1


@dataclass
class SchedulerConfig(_ConfigBase):
    """Reference: vllm/config/scheduler.py:74-99, arg_utils.py:2611-2633."""

    max_num_batched_tokens: int = 16384
    max_num_seqs: int = 1024
    max_model_len: int = 8192
    enable_chunked_prefill: bool = True
    long_prefill_token_threshold: int = 0
    max_num_partial_prefills: int = 1
    policy: Literal["fcfs", "priority"] = "fcfs"
    async_scheduling: bool = False
    num_scheduler_steps: int = 1
    scheduler_cls: Optional[str] = None
    disable_hybrid_kv_cache_manager: bool = False
    max_num_encoder_input_tokens: int = 16384
    encoder_cache_size: int = 16384

    def __post_init__(self) -> None:
        if self.max_num_seqs < 1:
            raise ValueError("max_num_seqs must be >= 1")
        if not self.enable_chunked_prefill and self.max_num_batched_tokens < self.max_model_len:
            raise ValueError(
                "max_num_batched_tokens must be >= max_model_len when chunked prefill is disabled")
        if self.policy not in ("fcfs", "priority"):
            raise ValueError(f"unknown scheduling policy {self.policy}")


@dataclass
class DeviceConfig(_ConfigBase):
    device: str = "auto"  # auto|cuda|cpu

    def resolve(self) -> str:
        if self.device != "auto":
            return self.device
        import torch
        return "cuda" if torch.cuda.is_available() else "cpu"


@dataclass
class LoadConfig(_ConfigBase):
    """Reference: model_executor/model_loader (dummy loader: weight_utils.py:1289-1349)."""

    load_format: str = "auto"  # auto|safetensors|pt|dummy|sharded_state|gguf|streamer (runai_streamer)
    model_loader_extra_config: Optional[dict] = None       # streamer: {"concurrency": 16, "memory_limit": bytes}
    download_dir: Optional[str] = None
    dummy_seed: int = 1234
    ignore_patterns: list[str] = field(default_factory=list)
    pt_load_map_location: str = "cpu"


@dataclass
class LoRAConfig(_ConfigBase):
    max_lora_rank: int = 16
    max_loras: int = 1
    max_cpu_loras: Optional[int] = None
    lora_dtype: str = "auto"
    fully_sharded_loras: bool = False


@dataclass
class SpeculativeConfig(_ConfigBase):
    """Reference: vllm/config/speculative.py."""

    method: str = "ngram"  # ngram|eagle|eagle3|mtp|medusa|draft_model|suffix|auto (from the draft checkpoint's architecture)
    num_speculative_tokens: int = 4
    model: Optional[str] = None
    prompt_lookup_max: int = 5
    prompt_lookup_min: int = 2
    draft_tensor_parallel_size: int = 1
    disable_by_batch_size: Optional[int] = None
    # dynamic K: per request, the number of drafts actually verified follows the acceptance rate it has seen
    # (EMA; shrink when drafts are mostly rejected, grow back when they are accepted) between 1 and num_speculative_tokens
    dynamic_k: bool = False
    dynamic_k_ema: float = 0.5

    def __post_init__(self) -> None:
        if self.method in ("auto", None, ""):
            from vllm_b200.spec_decode import infer_method
            self.method = infer_method(self.model)


@dataclass
class StructuredOutputsConfig(_ConfigBase):
    backend: str = "auto"  # auto|xgrammar|guidance|outlines|lm-format-enforcer|regex_fsm
    disable_fallback: bool = False
    reasoning_parser: str = ""


@dataclass
class ObservabilityConfig(_ConfigBase):
    otlp_traces_endpoint: Optional[str] = None
    collect_detailed_traces: Optional[list[str]] = None
    show_hidden_metrics: bool = False
    enable_mfu_metrics: bool = False
    kv_cache_metrics: bool = False


@dataclass
class CompilationConfig(_ConfigBase):
    """CUDA-graph policy. There is no tracing compiler in this framework: the
    decode step is a fixed kernel sequence captured directly.
    Modes mirror the reference's CUDAGraphMode (config/compilation.py:53-64)."""

    cudagraph_mode: str = "FULL_DECODE_ONLY"  # NONE|FULL_DECODE_ONLY|FULL  (FULL: mixed prefill+decode buckets too)
    cudagraph_capture_sizes: Optional[list[int]] = None
    max_cudagraph_capture_size: int = 512
    # token buckets of the mixed-batch graphs (mode FULL); the request bucket of each is min(tokens, max_num_seqs, 256)
    mixed_capture_sizes: Optional[list[int]] = None

    def mixed_sizes(self, max_num_batched_tokens: int) -> list[int]:
        if self.cudagraph_mode not in ("FULL", "FULL_AND_PIECEWISE"):
            return []
        sizes = self.mixed_capture_sizes if self.mixed_capture_sizes is not None else [64, 128, 256, 512]
        return sorted({s for s in sizes if s <= max_num_batched_tokens})

    def capture_sizes(self, max_num_seqs: int) -> list[int]:
        if self.cudagraph_capture_sizes is not None:
            return sorted(set(self.cudagraph_capture_sizes))
        sizes = [1, 2, 4] + list(range(8, 257, 8)) + list(range(272, 1025, 16))
        cap = min(self.max_cudagraph_capture_size, max_num_seqs)
        out = [s for s in sizes if s <= cap]
        if cap not in out:
            out.append(cap)
        return sorted(out)


@dataclass
class KVTransferConfig(_ConfigBase):
    kv_connector: Optional[str] = None
    kv_role: Optional[str] = None  # kv_producer|kv_consumer|kv_both
    kv_rank: int = 0
    kv_parallel_size: int = 1
    kv_buffer_size: float = 1e9
    kv_ip: str = "127.0.0.1"
    kv_port: int = 14579
    kv_connector_extra_config: dict[str, Any] = field(default_factory=dict)

    @property
    def is_kv_producer(self) -> bool:
        return self.kv_role in ("kv_producer", "kv_both")

    @property
    def is_kv_consumer(self) -> bool:
        return self.kv_role in ("kv_consumer", "kv_both")


@dataclass
class KVEventsConfig(_ConfigBase):
    enable_kv_cache_events: bool = False
    publisher: str = "null"  # null|zmq
    endpoint: str = "tcp://*:5557"
    topic: str = ""


@dataclass
class ProfilerConfig(_ConfigBase):
    profiler: Optional[str] = None  # torch|cuda
    torch_profiler_dir: Optional[str] = None
    delay_iterations: int = 0
    max_iterations: int = 0
    with_stack: bool = False
    record_shapes: bool = False


@dataclass
class FaultToleranceConfig(_ConfigBase):
    enable_fault_tolerance: bool = False
    engine_recovery_timeout_sec: float = 120.0


@dataclass
class OffloadConfig(_ConfigBase):
    kv_offload_backend: Optional[str] = None  # cpu|fs
    kv_offload_bytes: int = 0
    kv_offload_dir: Optional[str] = None
    eviction_policy: str = "lru"  # lru|arc


@dataclass
class VllmConfig(_ConfigBase):
    model_config: ModelConfig = field(default_factory=ModelConfig)
    cache_config: CacheConfig = field(default_factory=CacheConfig)
    parallel_config: ParallelConfig = field(default_factory=ParallelConfig)
    scheduler_config: SchedulerConfig = field(default_factory=SchedulerConfig)
    device_config: DeviceConfig = field(default_factory=DeviceConfig)
    load_config: LoadConfig = field(default_factory=LoadConfig)
    lora_config: Optional[LoRAConfig] = None
    speculative_config: Optional[SpeculativeConfig] = None
    structured_outputs_config: StructuredOutputsConfig = field(default_factory=StructuredOutputsConfig)
    observability_config: ObservabilityConfig = field(default_factory=ObservabilityConfig)
    compilation_config: CompilationConfig = field(default_factory=CompilationConfig)
    kv_transfer_config: Optional[KVTransferConfig] = None
    kv_events_config: Optional[KVEventsConfig] = None
    profiler_config: ProfilerConfig = field(default_factory=ProfilerConfig)
    fault_tolerance_config: FaultToleranceConfig = field(default_factory=FaultToleranceConfig)
    offload_config: OffloadConfig = field(default_factory=OffloadConfig)
    instance_id: str = ""

    def __post_init__(self) -> None:
        mc, sc, pc = self.model_config, self.scheduler_config, self.parallel_config
        # cross-field validation (reference: VllmConfig.__post_init__ in config/vllm.py)
        if sc.max_model_len > mc.max_model_len:
            sc.max_model_len = mc.max_model_len
        else:
            mc.max_model_len = sc.max_model_len
        if sc.enable_chunked_prefill:
            sc.max_num_batched_tokens = max(sc.max_num_batched_tokens, 1)
        mc.verify_with_parallel_config(pc)
        if self.cache_config.sliding_window is None:
            self.cache_config.sliding_window = mc.sliding_window
        if self.speculative_config is not None and sc.async_scheduling:
            if self.speculative_config.method not in ("ngram",):
                pass  # draft-model speculation is compatible with async scheduling
        if not self.instance_id:
            import uuid
            self.instance_id = uuid.uuid4().hex[:8]

    def compute_hash(self) -> str:
        d = self.to_dict()
        d.pop("instance_id", None)
        return _hash_dict(d)


_current: list[VllmConfig] = []


@contextlib.contextmanager
def set_current_vllm_config(cfg: VllmConfig):
    _current.append(cfg)
    try:
        yield
    finally:
        _current.pop()


def get_current_vllm_config() -> VllmConfig:
    if not _current:
        raise RuntimeError("no current VllmConfig set")
    return _current[-1]


def get_current_vllm_config_or_none() -> Optional[VllmConfig]:
    return _current[-1] if _current else None


__all__ = [
    "ModelConfig", "CacheConfig", "ParallelConfig", "SchedulerConfig", "DeviceConfig", "LoadConfig",
    "LoRAConfig", "SpeculativeConfig", "StructuredOutputsConfig", "ObservabilityConfig",
    "CompilationConfig", "KVTransferConfig", "KVEventsConfig", "ProfilerConfig", "FaultToleranceConfig",
    "OffloadConfig", "VllmConfig", "set_current_vllm_config", "get_current_vllm_config",
    "get_current_vllm_config_or_none", "PRESETS", "get_preset_hf_config",
]
