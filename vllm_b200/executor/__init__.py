"""Executors: run workers in-process (``uni``), one process per GPU (``mp``), or
attach to an external launcher such as torchrun (``external_launcher``).

Behavioural reference: vllm/v1/executor/abstract.py (``Executor.get_class`` :49,
``collective_rpc`` :163, ``execute_model`` :220), uniproc_executor.py:51,161,
multiproc_executor.py:111 (worker procs + shm broadcast ring + death pipe).
"""
from __future__ import annotations

import os
from concurrent.futures import Future
from typing import Any, Callable, Optional, Union

from vllm_b200.config import VllmConfig
from vllm_b200.core.kv_cache_manager import KVCacheConfig, KVCacheSpec
from vllm_b200.core.sched_output import ModelRunnerOutput, SchedulerOutput
from vllm_b200.logger import init_logger

logger = init_logger(__name__)


class Executor:
    uses_multiproc = False

    def __init__(self, vllm_config: VllmConfig) -> None:
        self.vllm_config = vllm_config
        self.is_failed = False
        self.failure_callback: Optional[Callable[[], None]] = None
        self._init_executor()

    @staticmethod
    def get_class(vllm_config: VllmConfig) -> type["Executor"]:
        b = vllm_config.parallel_config.distributed_executor_backend
        if b == "uni":
            return UniProcExecutor
        if b == "mp":
            from vllm_b200.executor.multiproc import MultiprocExecutor
            return MultiprocExecutor
        if b == "external_launcher":
            return ExternalLauncherExecutor
        if b == "ray":
            from vllm_b200.executor.ray_executor import RayExecutor
            return RayExecutor
        raise ValueError(f"unknown distributed_executor_backend {b!r}")

    def _init_executor(self) -> None:
        raise NotImplementedError

    def collective_rpc(self, method: Union[str, Callable], args: tuple = (), kwargs: Optional[dict] = None,
                       non_block: bool = False, unique_reply_rank: Optional[int] = None) -> Any:
        raise NotImplementedError

    # ---- convenience wrappers (same names as the reference's Executor API)
    def get_kv_cache_specs(self) -> list[dict[str, KVCacheSpec]]:
        return self.collective_rpc("get_kv_cache_specs")

    def determine_available_memory(self) -> list[int]:
        return self.collective_rpc("determine_available_memory")

    def initialize_from_config(self, cfg: KVCacheConfig) -> None:
        self.collective_rpc("initialize_from_config", (cfg,))
        self.collective_rpc("compile_or_warm_up_model")

    def execute_model(self, so: SchedulerOutput, non_block: bool = False):
        out = self.collective_rpc("execute_model", (so,), non_block=non_block,
                                  unique_reply_rank=self.output_rank)
        return out

    @property
    def output_rank(self) -> int:
        pc = self.vllm_config.parallel_config
        # first TP rank of the last PP stage owns the sampled tokens
        return pc.world_size - pc.tensor_parallel_size * pc.prefill_context_parallel_size

    @property
    def max_concurrent_batches(self) -> int:
        pc = self.vllm_config.parallel_config
        if self.vllm_config.scheduler_config.async_scheduling:
            return 2
        return pc.pipeline_parallel_size

    def profile(self, is_start: bool = True) -> None:
        self.collective_rpc("profile", (is_start,))

    def sleep(self, level: int = 1) -> None:
        self.collective_rpc("sleep", (level,))

    def wake_up(self, tags=None) -> None:
        self.collective_rpc("wake_up", (tags,))

    def save_sharded_state(self, path: str) -> None:
        self.collective_rpc("save_sharded_state", (path,))

    def check_health(self) -> None:
        if self.is_failed:
            raise RuntimeError("executor failed")

    def register_failure_callback(self, cb: Callable[[], None]) -> None:
        self.failure_callback = cb

    def shutdown(self) -> None:
        pass


def _run_method(worker, method: Union[str, Callable], args: tuple, kwargs: dict) -> Any:
    if isinstance(method, str):
        return getattr(worker, method)(*args, **kwargs)
    return method(worker, *args, **kwargs)


class _LazyFuture:
    """Future-like wrapper over a launched step: ``result()`` waits for the device → host copy."""

    def __init__(self, out) -> None:
        self._out = out

    def done(self) -> bool:
        d = getattr(self._out, "done", None)
        return d() if d is not None else True

    def result(self, timeout=None):
        g = getattr(self._out, "get", None)
        return g() if g is not None else self._out


class UniProcExecutor(Executor):
    """Single worker inside the engine-core process (TP=PP=1)."""

    def _init_executor(self) -> None:
        from vllm_b200.worker import resolve_worker_class
        self.worker = resolve_worker_class(self.vllm_config)(self.vllm_config, local_rank=0, rank=0)
        self.worker.init_device()
        self.worker.load_model()

    def collective_rpc(self, method, args=(), kwargs=None, non_block=False, unique_reply_rank=None):
        res = _run_method(self.worker, method, args, kwargs or {})
        out = res if unique_reply_rank is not None else [res]
        if non_block:
            f: Future = Future()
            f.set_result(out)
            return f
        return out

    def execute_model(self, so: SchedulerOutput, non_block: bool = False):
        out = self.worker.execute_model(so, non_block=non_block)
        return _LazyFuture(out) if non_block else out

    @property
    def output_rank(self) -> int:
        return 0

    def shutdown(self) -> None:
        if getattr(self, "worker", None) is not None:
            self.worker.shutdown()
            self.worker = None


class ExternalLauncherExecutor(UniProcExecutor):
    """torchrun-style: every rank runs the full engine; the scheduler is deterministic so all
    ranks take identical decisions (reference: uniproc_executor.py:161)."""

    def _init_executor(self) -> None:
        from vllm_b200.worker import resolve_worker_class
        rank = int(os.environ.get("RANK", 0))
        local_rank = int(os.environ.get("LOCAL_RANK", rank))
        self.worker = resolve_worker_class(self.vllm_config)(self.vllm_config, local_rank=local_rank, rank=rank,
                             distributed_init_method="env://")
        self.worker.init_device()
        self.worker.load_model()

    def determine_available_memory(self) -> list[int]:
        import torch
        import torch.distributed as dist
        mem = self.worker.determine_available_memory()
        if dist.is_initialized() and dist.get_world_size() > 1:
            from vllm_b200.parallel.state import get_world_group
            t = torch.tensor([mem], dtype=torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=get_world_group().cpu_group)
            mem = int(t.item())
        return [mem]
