"""Asyncio engine front-end used by the API servers.

Behavioural reference: vllm/v1/engine/async_llm.py (``AsyncLLM`` :75, ``generate`` :550 async generator
draining a per-request ``RequestOutputCollector``, background ``output_handler`` :686-745, ``abort`` :749,
``encode`` :843, profile / sleep / health utilities :940-1104).
"""
from __future__ import annotations

import asyncio
import copy
import time
from typing import Any, AsyncGenerator, Optional, Union

from vllm_b200 import envs
from vllm_b200.config import VllmConfig
from vllm_b200.engine.arg_utils import EngineArgs
from vllm_b200.engine.core_client import EngineCoreClient
from vllm_b200.engine.exceptions import EngineDeadError, EngineGenerateError
from vllm_b200.engine.processors import (InputProcessor, IterationStats, OutputProcessor, ParentRequest,
                                         RequestOutputCollector)
from vllm_b200.inputs import PromptType
from vllm_b200.logger import init_logger
from vllm_b200.outputs import PoolingRequestOutput, RequestOutput
from vllm_b200.sampling_params import PoolingParams, RequestOutputKind, SamplingParams
from vllm_b200.tokenizers import init_tokenizer_from_config

logger = init_logger(__name__)


class AsyncLLM:
    def __init__(self, vllm_config: VllmConfig, log_stats: bool = True, stat_loggers=None,
                 start_engine_loop: bool = True) -> None:
        self.vllm_config = vllm_config
        self.model_config = vllm_config.model_config
        self.log_stats = log_stats
        self.tokenizer = init_tokenizer_from_config(self.model_config)
        self.input_processor = InputProcessor(vllm_config, self.tokenizer)
        from vllm_b200.tracing import maybe_init_tracer
        self.tracer = maybe_init_tracer(vllm_config.observability_config)
        self.output_processor = OutputProcessor(self.tokenizer, log_stats, self.tracer)
        self.engine_core = EngineCoreClient.make_client(vllm_config, True, True, log_stats)
        self.stat_logger = None
        if log_stats:
            from vllm_b200.metrics import StatLoggerManager
            self.stat_logger = StatLoggerManager(vllm_config, stat_loggers)
        self.output_handler: Optional[asyncio.Task] = None
        self._errored_with: Optional[BaseException] = None
        self._paused = False
        from vllm_b200.usage import usage_message
        usage_message.report_usage(vllm_config, "OPENAI_API_SERVER")

    @classmethod
    def from_engine_args(cls, engine_args: EngineArgs, **kw) -> "AsyncLLM":
        return cls(engine_args.create_engine_config(), log_stats=not engine_args.disable_log_stats, **kw)

    @classmethod
    def from_vllm_config(cls, vllm_config: VllmConfig, **kw) -> "AsyncLLM":
        return cls(vllm_config, **kw)

    # ------------------------------------------------------------ properties
    @property
    def is_running(self) -> bool:
        return self.output_handler is None or not self.output_handler.done()

    @property
    def errored(self) -> bool:
        return self._errored_with is not None or self.engine_core.res.engine_dead

    @property
    def dead_error(self) -> BaseException:
        return EngineDeadError()

    async def get_tokenizer(self):
        return self.tokenizer

    async def get_model_config(self):
        return self.model_config

    async def get_vllm_config(self):
        return self.vllm_config

    # ------------------------------------------------------------- requests
    async def add_request(self, request_id: str, prompt: PromptType, params: Union[SamplingParams, PoolingParams],
                          arrival_time: Optional[float] = None, lora_request=None, priority: int = 0,
                          trace_headers=None, data_parallel_rank: Optional[int] = None) -> RequestOutputCollector:
        if self.errored:
            raise EngineDeadError()
        self._run_output_handler()
        kind = params.output_kind if isinstance(params, SamplingParams) else RequestOutputKind.FINAL_ONLY
        queue = RequestOutputCollector(kind)
        text, req = self.input_processor.process_inputs(request_id, prompt, params, arrival_time, lora_request,
                                                        priority, trace_headers, data_parallel_rank)
        n = params.n if isinstance(params, SamplingParams) else 1
        if n == 1:
            self.output_processor.add_request(req, text, None, 0, queue)
            await self.engine_core.add_request_async(req)
            return queue
        parent = ParentRequest(request_id, req.sampling_params)
        for i in range(n):
            cid, csp = parent.child(i)
            child = copy.copy(req)
            child.request_id, child.sampling_params = cid, csp
            self.output_processor.add_request(child, text, parent, i, queue)
            await self.engine_core.add_request_async(child)
        return queue

    async def generate(self, prompt: PromptType, sampling_params: SamplingParams, request_id: str,
                       lora_request=None, trace_headers=None, priority: int = 0,
                       data_parallel_rank: Optional[int] = None) -> AsyncGenerator[RequestOutput, None]:
        try:
            q = await self.add_request(request_id, prompt, sampling_params, lora_request=lora_request,
                                       priority=priority, trace_headers=trace_headers,
                                       data_parallel_rank=data_parallel_rank)
            finished = False
            while not finished:
                out = q.get_nowait() or await q.get()
                finished = out.finished
                yield out
        except asyncio.CancelledError:
            await self.abort(request_id)
            raise
        except EngineDeadError:
            raise
        except ValueError:
            raise
        except Exception as e:
            await self.abort(request_id)
            raise EngineGenerateError() from e

    async def encode(self, prompt: PromptType, pooling_params: PoolingParams, request_id: str, lora_request=None,
                     trace_headers=None, priority: int = 0) -> AsyncGenerator[PoolingRequestOutput, None]:
        try:
            q = await self.add_request(request_id, prompt, pooling_params, lora_request=lora_request,
                                       priority=priority, trace_headers=trace_headers)
            out = q.get_nowait() or await q.get()
            yield out
        except asyncio.CancelledError:
            await self.abort(request_id)
            raise

    async def abort(self, request_id: Union[str, list[str]]) -> None:
        ids = [request_id] if isinstance(request_id, str) else list(request_id)
        ids = self.output_processor.abort_requests(ids)
        await self.engine_core.abort_requests_async(ids)

    # ------------------------------------------------------- output handler
    def _run_output_handler(self) -> None:
        if self.output_handler is not None:
            return
        chunk = envs.VLLM_OUTPUT_PROC_CHUNK_SIZE

        async def handler() -> None:
            try:
                while True:
                    outputs = await self.engine_core.get_output_async()
                    n = len(outputs.outputs)
                    stats = IterationStats(iteration_timestamp=time.time()) if self.log_stats else None
                    # chunked so one huge batch does not starve the event loop
                    for s in range(0, max(n, 1), chunk):
                        part = outputs.outputs[s:s + chunk]
                        processed = self.output_processor.process_outputs(part, outputs.timestamp, stats)
                        if processed.reqs_to_abort:
                            await self.engine_core.abort_requests_async(processed.reqs_to_abort)
                        if s + chunk < n:
                            await asyncio.sleep(0)
                    if self.stat_logger is not None and (outputs.scheduler_stats is not None or n):
                        self.stat_logger.record(outputs.scheduler_stats, stats)
            except asyncio.CancelledError:
                raise
            except Exception as e:
                logger.exception("AsyncLLM output handler failed")
                self._errored_with = e
                self.output_processor.propagate_error(e if isinstance(e, EngineDeadError) else EngineDeadError())

        self.output_handler = asyncio.get_running_loop().create_task(handler())

    # ------------------------------------------------------------ utilities
    async def check_health(self) -> None:
        if self.errored:
            raise self.dead_error

    async def start_profile(self) -> None:
        await self.engine_core.call_utility_async("profile", True)

    async def stop_profile(self) -> None:
        await self.engine_core.call_utility_async("profile", False)

    async def reset_prefix_cache(self) -> bool:
        return await self.engine_core.call_utility_async("reset_prefix_cache")

    async def sleep(self, level: int = 1) -> None:
        await self.reset_prefix_cache()
        await self.engine_core.call_utility_async("sleep", level)

    async def wake_up(self, tags: Optional[list[str]] = None) -> None:
        await self.engine_core.call_utility_async("wake_up", tags)

    async def is_sleeping(self) -> bool:
        return await self.engine_core.call_utility_async("is_sleeping")

    async def scale_elastic_ep(self, new_data_parallel_size: int) -> int:
        """Scale the number of engine replicas up or down without restarting the server."""
        if not hasattr(self.engine_core, "scale_elastic"):
            raise RuntimeError("elastic scaling needs --enable-elastic-ep (or data_parallel_size > 1)")
        return await self.engine_core.scale_elastic(new_data_parallel_size)

    async def get_fault_status(self) -> dict:
        return await self.engine_core.call_utility_async("get_fault_status")

    async def handle_fault(self, instruction: str) -> bool:
        return await self.engine_core.call_utility_async("handle_fault", instruction)

    async def collective_rpc(self, method: str, args: tuple = (), kwargs: Optional[dict] = None):
        return await self.engine_core.call_utility_async("collective_rpc", method, list(args), kwargs)

    async def save_sharded_state(self, path: str) -> None:
        await self.engine_core.call_utility_async("save_sharded_state", path)

    async def get_request_counts(self) -> tuple[int, int]:
        return tuple(await self.engine_core.call_utility_async("get_request_counts"))

    async def do_log_stats(self) -> None:
        if self.stat_logger is not None:
            self.stat_logger.log()

    def shutdown(self) -> None:
        if self.output_handler is not None:
            self.output_handler.cancel()
        ec = getattr(self, "engine_core", None)
        if ec is not None:
            ec.shutdown()
            self.engine_core = None


AsyncLLMEngine = AsyncLLM
