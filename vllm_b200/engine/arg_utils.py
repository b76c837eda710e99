"""EngineArgs: one flat dataclass mirrored as CLI flags and ``LLM(...)`` kwargs.

Behavioural reference: vllm/engine/arg_utils.py (``EngineArgs`` →
``create_engine_config``, argparse generated from field types; usage-context
defaults :2586-2640).
"""
from __future__ import annotations

import argparse
import dataclasses
import json
import os
from dataclasses import dataclass, field
from typing import Any, Optional, Union, get_args, get_origin

from vllm_b200.config import (CacheConfig, CompilationConfig, DeviceConfig, FaultToleranceConfig,
                              KVEventsConfig, KVTransferConfig, LoadConfig, LoRAConfig, ModelConfig,
                              ObservabilityConfig, OffloadConfig, ParallelConfig, ProfilerConfig,
                              SchedulerConfig, SpeculativeConfig, StructuredOutputsConfig, VllmConfig)


@dataclass
class EngineArgs:
    model: str = "opt-tiny"
    tokenizer: Optional[str] = None
    tokenizer_mode: str = "auto"
    skip_tokenizer_init: bool = False
    trust_remote_code: bool = False
    allowed_local_media_path: Optional[str] = None   # file:// media parts are refused unless set (then confined to it)
    dtype: str = "auto"
    seed: int = 0
    max_model_len: Optional[int] = None
    quantization: Optional[str] = None
    enforce_eager: bool = False
    max_logprobs: int = 20
    served_model_name: Optional[str] = None
    hf_overrides: Optional[dict[str, Any]] = None
    runner: str = "generate"
    model_impl: str = "auto"              # auto | vllm_b200 (native classes only) | transformers (force the HF back-end)
    load_format: str = "auto"
    model_loader_extra_config: Optional[dict[str, Any]] = None
    device: str = "auto"
    # cache
    block_size: Optional[int] = None      # default: 64 (tcgen05 attention page) when head_dim is 128, else 32
    gpu_memory_utilization: float = 0.92
    swap_space: float = 4.0
    kv_cache_dtype: str = "auto"
    num_gpu_blocks_override: Optional[int] = None
    enable_prefix_caching: bool = True
    prefix_caching_hash_algo: str = "sha256"
    kv_cache_memory_bytes: Optional[int] = None
    cpu_offload_gb: float = 0.0
    # parallel
    tensor_parallel_size: int = 1
    pipeline_parallel_size: int = 1
    data_parallel_size: int = 1
    data_parallel_rank: Optional[int] = None
    data_parallel_size_local: Optional[int] = None
    data_parallel_address: Optional[str] = None
    data_parallel_rpc_port: Optional[int] = None
    data_parallel_external_lb: bool = False
    data_parallel_hybrid_lb: bool = False
    enable_expert_parallel: bool = False
    all2all_backend: str = "a2a"
    logits_processors: Optional[list[str]] = None    # "pkg.module:Class" plug-ins (comma separated on the CLI)
    enable_eplb: bool = False
    num_redundant_experts: int = 0
    eplb_step_interval: int = 3000
    enable_dbo: bool = False
    decode_context_parallel_size: int = 1
    prefill_context_parallel_size: int = 1
    dcp_comm_backend: str = "ag_rs"
    disable_custom_all_reduce: bool = False
    distributed_executor_backend: Optional[str] = None
    worker_cls: str = "auto"                 # qualified class name replacing the built-in worker
    worker_extension_cls: str = ""           # qualified class whose methods are mixed into the worker (collective_rpc targets)
    enable_elastic_ep: bool = False
    # scheduler
    max_num_batched_tokens: Optional[int] = None
    max_num_seqs: Optional[int] = None
    enable_chunked_prefill: bool = True
    long_prefill_token_threshold: int = 0
    scheduling_policy: str = "fcfs"
    async_scheduling: bool = False
    scheduler_cls: Optional[str] = None
    # features
    enable_lora: bool = False
    max_lora_rank: int = 16
    max_loras: int = 1
    speculative_config: Optional[dict[str, Any]] = None
    structured_outputs_backend: str = "auto"
    reasoning_parser: str = ""
    kv_transfer_config: Optional[dict[str, Any]] = None
    kv_events_config: Optional[dict[str, Any]] = None
    kv_offload_backend: Optional[str] = None
    kv_offload_bytes: int = 0
    # compilation / graphs
    cudagraph_mode: str = "FULL_DECODE_ONLY"
    cudagraph_capture_sizes: Optional[list[int]] = None
    mixed_capture_sizes: Optional[list[int]] = None
    max_cudagraph_capture_size: int = 512
    # observability
    disable_log_stats: bool = False
    otlp_traces_endpoint: Optional[str] = None
    profiler: Optional[str] = None
    torch_profiler_dir: Optional[str] = None
    enable_fault_tolerance: bool = False
    # ---- reference flags with a direct meaning here ------------------------------------------------------------------
    spec_method: Optional[str] = None                 # shorthand for --speculative-config {"method": …, "model": …, "num_speculative_tokens": …}
    spec_model: Optional[str] = None
    spec_tokens: Optional[int] = None
    limit_mm_per_prompt: Optional[dict[str, Any]] = None     # {"image": 4, "audio": 0, …}: items allowed per request
    disable_sliding_window: bool = False              # run sliding-window models as full attention, capped at the window
    disable_cascade_attn: bool = False
    hf_config_path: Optional[str] = None              # directory whose config.json overrides the model's own
    nnodes: int = 1
    node_rank: int = 0
    master_addr: str = "127.0.0.1"
    master_port: int = 29501
    # ---- accepted for command-line compatibility; always-on or not applicable in an offline box ---------------------------
    revision: Optional[str] = None                    # (hub revisions: models are local directories or presets here)
    code_revision: Optional[str] = None
    tokenizer_revision: Optional[str] = None
    hf_token: Optional[str] = None
    download_dir: Optional[str] = None
    use_tqdm_on_load: bool = True
    enable_sleep_mode: bool = True                    # sleep / wake_up are always available
    enable_prompt_embeds: bool = True                 # prompt_embeds inputs are always accepted
    enable_mm_embeds: bool = True

    def create_engine_config(self) -> VllmConfig:
        if self.allowed_local_media_path:
            from vllm_b200.entrypoints.chat_utils import set_allowed_local_media_path
            set_allowed_local_media_path(self.allowed_local_media_path)
        hf_over = dict(self.hf_overrides or {})
        if self.hf_config_path:
            import json as _json
            with open(os.path.join(self.hf_config_path, "config.json")) as f:
                hf_over = {**_json.load(f), **hf_over}
        if self.limit_mm_per_prompt:
            hf_over["limit_mm_per_prompt"] = {str(k): (v.get("count") if isinstance(v, dict) else v)
                                              for k, v in dict(self.limit_mm_per_prompt).items()}
        self.hf_overrides = hf_over
        if self.spec_method or self.spec_model or self.spec_tokens:
            sc = dict(self.speculative_config or {})
            sc.setdefault("method", self.spec_method or ("auto" if self.spec_model else "ngram"))
            if self.spec_model:
                sc.setdefault("model", self.spec_model)
            if self.spec_tokens:
                sc.setdefault("num_speculative_tokens", int(self.spec_tokens))
            self.speculative_config = sc
        if self.disable_cascade_attn:
            os.environ["VLLM_B200_CASCADE_MIN_REQS"] = str(1 << 30)
        mc = ModelConfig(
            model=self.model, tokenizer=self.tokenizer, tokenizer_mode=self.tokenizer_mode,
            trust_remote_code=self.trust_remote_code, dtype=self.dtype, seed=self.seed,
            max_model_len=self.max_model_len or 0, quantization=self.quantization,
            enforce_eager=self.enforce_eager, max_logprobs=self.max_logprobs,
            served_model_name=self.served_model_name, hf_overrides=self.hf_overrides or {},
            runner=self.runner, model_impl=self.model_impl, skip_tokenizer_init=self.skip_tokenizer_init,
            logits_processors=[str(x) for x in self.logits_processors] if self.logits_processors else None)
        if self.disable_sliding_window and mc.hf_config.get("sliding_window"):
            # reference semantics (config/model.py ``disable_sliding_window``): full attention, context capped at the window
            w = int(mc.hf_config["sliding_window"])
            mc.hf_config["sliding_window"] = None
            mc.hf_config["use_sliding_window"] = False
            mc.max_model_len = min(mc.max_model_len, w) if mc.max_model_len > 0 else w
        block_size = self.block_size
        if block_size is None:
            # 64-token pages are the tcgen05 kernels' TMA box; models that run on the generic kernel (other head sizes,
            # ALiBi / sinks / soft-cap) use the 32-token page it is tuned for
            special = any(mc.hf_config.get(k) for k in ("attn_logit_softcapping", "alibi")) or \
                mc.hf_config.get("model_type") in ("bloom", "mpt", "gpt_oss")
            block_size = 64 if (self.device == "cpu" or mc.is_mla or (mc.head_dim == 128 and not special)) else 32
        cc = CacheConfig(
            block_size=block_size, gpu_memory_utilization=self.gpu_memory_utilization,
            swap_space_gb=self.swap_space, cache_dtype=self.kv_cache_dtype,
            num_gpu_blocks_override=self.num_gpu_blocks_override,
            enable_prefix_caching=self.enable_prefix_caching,
            prefix_caching_hash_algo=self.prefix_caching_hash_algo,
            kv_cache_memory_bytes=self.kv_cache_memory_bytes, cpu_offload_gb=self.cpu_offload_gb)
        pc = ParallelConfig(
            nnodes=self.nnodes, node_rank=self.node_rank, master_addr=self.master_addr,
            tensor_parallel_size=self.tensor_parallel_size, pipeline_parallel_size=self.pipeline_parallel_size,
            data_parallel_size=self.data_parallel_size, data_parallel_rank=self.data_parallel_rank or 0,
            data_parallel_size_local=self.data_parallel_size_local,
            data_parallel_master_ip=self.data_parallel_address or "127.0.0.1",
            data_parallel_master_port=self.data_parallel_rpc_port or 29550,
            data_parallel_external_lb=self.data_parallel_external_lb,
            data_parallel_hybrid_lb=self.data_parallel_hybrid_lb,
            enable_expert_parallel=self.enable_expert_parallel, all2all_backend=self.all2all_backend,
            enable_eplb=self.enable_eplb, num_redundant_experts=self.num_redundant_experts,
            eplb_step_interval=self.eplb_step_interval, enable_dbo=self.enable_dbo,
            decode_context_parallel_size=self.decode_context_parallel_size,
            prefill_context_parallel_size=self.prefill_context_parallel_size,
            dcp_comm_backend=self.dcp_comm_backend, disable_custom_all_reduce=self.disable_custom_all_reduce,
            distributed_executor_backend=self.distributed_executor_backend,
            worker_cls=self.worker_cls, worker_extension_cls=self.worker_extension_cls,
            enable_elastic_ep=self.enable_elastic_ep)
        # usage defaults sized for 180 GB HBM3e (reference picks 16384/8192 on big-HBM parts)
        mnbt = self.max_num_batched_tokens or 16384
        mns = self.max_num_seqs or 1024
        mml = mc.max_model_len
        chunked = self.enable_chunked_prefill
        if mc.is_encoder_only:
            # bidirectional attention needs the whole sequence in one forward pass; nothing to reuse across requests
            mc.runner = "pooling"
            chunked = False
            cc.enable_prefix_caching = False
        if mc.prefix_lm:
            # prefix-LM attention: image + prompt see each other in full, which only holds if they are prefilled together
            chunked = False
            cc.enable_prefix_caching = False
        if not chunked:
            mnbt = max(mnbt, mml)
        sc = SchedulerConfig(
            max_num_batched_tokens=mnbt, max_num_seqs=mns, max_model_len=mml,
            enable_chunked_prefill=chunked,
            long_prefill_token_threshold=self.long_prefill_token_threshold, policy=self.scheduling_policy,
            async_scheduling=self.async_scheduling, scheduler_cls=self.scheduler_cls)
        spec = SpeculativeConfig(**self.speculative_config) if self.speculative_config else None
        kvt = KVTransferConfig(**self.kv_transfer_config) if self.kv_transfer_config else None
        kve = KVEventsConfig(**self.kv_events_config) if self.kv_events_config else None
        return VllmConfig(
            model_config=mc, cache_config=cc, parallel_config=pc, scheduler_config=sc,
            device_config=DeviceConfig(self.device), load_config=LoadConfig(load_format=self.load_format,
                                   model_loader_extra_config=self.model_loader_extra_config),
            lora_config=LoRAConfig(self.max_lora_rank, self.max_loras) if self.enable_lora else None,
            speculative_config=spec,
            structured_outputs_config=StructuredOutputsConfig(self.structured_outputs_backend,
                                                              reasoning_parser=self.reasoning_parser),
            observability_config=ObservabilityConfig(otlp_traces_endpoint=self.otlp_traces_endpoint),
            compilation_config=CompilationConfig(self.cudagraph_mode, self.cudagraph_capture_sizes,
                                                 self.max_cudagraph_capture_size, self.mixed_capture_sizes),
            kv_transfer_config=kvt, kv_events_config=kve,
            profiler_config=ProfilerConfig(self.profiler, self.torch_profiler_dir),
            fault_tolerance_config=FaultToleranceConfig(self.enable_fault_tolerance),
            offload_config=OffloadConfig(self.kv_offload_backend, self.kv_offload_bytes))

    # ------------------------------------------------------------------ CLI
    @staticmethod
    def add_cli_args(parser: argparse.ArgumentParser) -> argparse.ArgumentParser:
        for f in dataclasses.fields(EngineArgs):
            flag = "--" + f.name.replace("_", "-")
            typ = f.type if not isinstance(f.type, str) else eval(f.type, globals(), {"Optional": Optional, "Any": Any, "Union": Union})
            origin = get_origin(typ)
            inner = typ
            if origin is Union:
                inner = next(a for a in get_args(typ) if a is not type(None))
                origin = get_origin(inner)
            default = f.default if f.default is not dataclasses.MISSING else None
            if inner is bool:
                parser.add_argument(flag, action=argparse.BooleanOptionalAction, default=default)
            elif origin in (dict,) or inner is dict:
                parser.add_argument(flag, type=json.loads, default=default)
            elif origin in (list,):
                elem = (get_args(inner) or (int,))[0]
                parser.add_argument(flag, type=(lambda s, e=elem: [e(x) for x in s.split(",")]), default=default)
            else:
                parser.add_argument(flag, type=inner, default=default)
        # reference-compatible aliases
        parser.add_argument("-tp", dest="tensor_parallel_size", type=int, default=argparse.SUPPRESS)
        parser.add_argument("-pp", dest="pipeline_parallel_size", type=int, default=argparse.SUPPRESS)
        parser.add_argument("-dp", dest="data_parallel_size", type=int, default=argparse.SUPPRESS)
        return parser

    @classmethod
    def from_cli_args(cls, args: argparse.Namespace) -> "EngineArgs":
        names = {f.name for f in dataclasses.fields(cls)}
        return cls(**{k: v for k, v in vars(args).items() if k in names})


    def to_llm_kwargs(self) -> dict:
        """Keyword arguments for ``LLM(**kwargs)`` equivalent to these engine args."""
        return {f.name: getattr(self, f.name) for f in dataclasses.fields(self)}


AsyncEngineArgs = EngineArgs
