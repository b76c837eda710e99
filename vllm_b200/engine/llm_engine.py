"""Synchronous engine front-end.

Behavioural reference: vllm/v1/engine/llm_engine.py (``LLMEngine`` :48,
``add_request`` :218 incl. n>1 fan-out :282-296, ``step`` :298).
"""
from __future__ import annotations

import time
from typing import Any, Optional, Union

from vllm_b200 import envs
from vllm_b200.config import VllmConfig
from vllm_b200.engine.arg_utils import EngineArgs
from vllm_b200.engine.core_client import EngineCoreClient
from vllm_b200.engine.processors import InputProcessor, IterationStats, OutputProcessor, ParentRequest
from vllm_b200.inputs import PromptType
from vllm_b200.logger import init_logger
from vllm_b200.outputs import PoolingRequestOutput, RequestOutput
from vllm_b200.sampling_params import PoolingParams, SamplingParams
from vllm_b200.tokenizers import init_tokenizer_from_config

logger = init_logger(__name__)


class LLMEngine:
    def __init__(self, vllm_config: VllmConfig, log_stats: bool = False, multiprocess_mode: Optional[bool] = None,
                 executor_class=None, stat_loggers=None) -> None:
        self.vllm_config = vllm_config
        self.model_config = vllm_config.model_config
        self.log_stats = log_stats
        self.tokenizer = init_tokenizer_from_config(self.model_config)
        self.input_processor = InputProcessor(vllm_config, self.tokenizer)
        from vllm_b200.tracing import maybe_init_tracer
        self.tracer = maybe_init_tracer(vllm_config.observability_config)
        self.output_processor = OutputProcessor(self.tokenizer, log_stats, self.tracer)
        if multiprocess_mode is None:
            multiprocess_mode = envs.VLLM_ENABLE_MULTIPROCESSING
        self.engine_core = EngineCoreClient.make_client(vllm_config, multiprocess_mode, False, log_stats,
                                                        executor_class)
        self.stat_logger = None
        if log_stats:
            from vllm_b200.metrics import StatLoggerManager
            self.stat_logger = StatLoggerManager(vllm_config, stat_loggers)
        from vllm_b200.usage import usage_message
        usage_message.report_usage(vllm_config, "LLM_CLASS")

    @classmethod
    def from_engine_args(cls, engine_args: EngineArgs, **kw) -> "LLMEngine":
        return cls(engine_args.create_engine_config(), log_stats=not engine_args.disable_log_stats, **kw)

    @classmethod
    def from_vllm_config(cls, vllm_config: VllmConfig, **kw) -> "LLMEngine":
        return cls(vllm_config, **kw)

    # ------------------------------------------------------------------ API
    def get_num_unfinished_requests(self) -> int:
        return self.output_processor.get_num_unfinished_requests()

    def has_unfinished_requests(self) -> bool:
        return self.output_processor.has_unfinished_requests()

    def add_request(self, request_id: str, prompt: PromptType, params: Union[SamplingParams, PoolingParams],
                    arrival_time: Optional[float] = None, lora_request=None, priority: int = 0,
                    trace_headers=None) -> None:
        text, req = self.input_processor.process_inputs(request_id, prompt, params, arrival_time, lora_request,
                                                        priority, trace_headers)
        n = params.n if isinstance(params, SamplingParams) else 1
        if n == 1:
            self.output_processor.add_request(req, text)
            self.engine_core.add_request(req)
            return
        parent = ParentRequest(request_id, req.sampling_params)
        import copy
        for i in range(n):
            cid, csp = parent.child(i)
            child = copy.copy(req)
            child.request_id, child.sampling_params = cid, csp
            self.output_processor.add_request(child, text, parent, i)
            self.engine_core.add_request(child)

    def abort_request(self, request_ids: Union[str, list[str]]) -> None:
        if isinstance(request_ids, str):
            request_ids = [request_ids]
        ids = self.output_processor.abort_requests(request_ids)
        self.engine_core.abort_requests(ids)

    def step(self) -> list[Union[RequestOutput, PoolingRequestOutput]]:
        outputs = self.engine_core.get_output()
        stats = IterationStats(iteration_timestamp=time.time()) if self.log_stats else None
        processed = self.output_processor.process_outputs(outputs.outputs, outputs.timestamp, stats)
        self.engine_core.abort_requests(processed.reqs_to_abort)
        if self.stat_logger is not None:
            self.stat_logger.record(outputs.scheduler_stats, stats)
        return processed.request_outputs

    # ---------------------------------------------------------------- misc
    def get_tokenizer(self):
        return self.tokenizer

    def start_profile(self):
        self.engine_core.profile(True)

    def stop_profile(self):
        self.engine_core.profile(False)

    def reset_prefix_cache(self):
        return self.engine_core.reset_prefix_cache()

    def sleep(self, level: int = 1):
        self.engine_core.sleep(level)

    def wake_up(self, tags=None):
        self.engine_core.wake_up(tags)

    def is_sleeping(self) -> bool:
        return self.engine_core.is_sleeping()

    def collective_rpc(self, method, args=(), kwargs=None):
        return self.engine_core.collective_rpc(method, args, kwargs)

    def get_metrics(self):
        from vllm_b200.metrics import get_metrics_snapshot
        return get_metrics_snapshot()

    def shutdown(self) -> None:
        ec = getattr(self, "engine_core", None)
        if ec is not None:
            ec.shutdown()
            self.engine_core = None

    def __del__(self):
        try:
            self.shutdown()
        except Exception:
            pass
