"""Data-parallel engine cores.

Behavioural reference: vllm/v1/engine/core.py ``DPEngineCoreProc`` (:2000: per-rank scheduler, lock-step
dummy batches so expert-parallel collectives line up, "wave" bookkeeping with the front-end) and
coordinator.py ``DPCoordinator`` (:23).

Every DP rank is a full engine core (own scheduler, own KV cache, own workers on its own GPUs).  Dense models
need no coupling at all.  MoE models with expert parallelism across DP ranks execute their all-to-all in every
forward pass, so all ranks must run the same number of forward passes: an idle rank runs a dummy batch while
any other rank still has work.  The ranks agree on "does anyone have work?" through a tiny key-value
rendezvous (``torch.distributed.TCPStore`` hosted by DP rank 0) once per step.
"""
from __future__ import annotations

import queue
import time
from typing import Optional

from vllm_b200.config import VllmConfig
from vllm_b200.engine.core import REQ_SHUTDOWN, EngineCoreProc
from vllm_b200.logger import init_logger

logger = init_logger(__name__)


class DPCoordinator:
    """Step-synchronous agreement between DP ranks: ``any_unfinished(local)`` returns the OR over ranks.

    One monotonically increasing round counter per rank in the store; round r of rank i is published under
    ``dp/<r>/<i>``.  Rank 0 garbage-collects old rounds."""

    def __init__(self, host: str, port: int, rank: int, size: int, timeout_s: float = 300.0) -> None:
        from datetime import timedelta

        from torch.distributed import TCPStore
        self.rank, self.size = rank, size
        self.store = TCPStore(host, port, world_size=size, is_master=rank == 0,
                              timeout=timedelta(seconds=timeout_s), wait_for_workers=True)
        self.round = 0

    def any_unfinished(self, local: bool) -> bool:
        r = self.round
        self.store.set(f"dp/{r}/{self.rank}", b"1" if local else b"0")
        keys = [f"dp/{r}/{i}" for i in range(self.size)]
        self.store.wait(keys)
        res = any(self.store.get(k) == b"1" for k in keys)
        if self.rank == 0 and r >= 2:
            for i in range(self.size):
                try:
                    self.store.delete_key(f"dp/{r - 2}/{i}")
                except Exception:
                    pass
        self.round = r + 1
        return res


class DPEngineCoreProc(EngineCoreProc):
    def __init__(self, vllm_config: VllmConfig, input_addr: str, output_addr: str, engine_index: int = 0,
                 log_stats: bool = False, executor_class=None) -> None:
        pc = vllm_config.parallel_config
        self.dp_rank, self.dp_size = pc.data_parallel_rank, pc.data_parallel_size
        super().__init__(vllm_config, input_addr, output_addr, engine_index, log_stats, executor_class)
        # lock-step is only needed when experts are sharded across DP ranks
        cfg = vllm_config.model_config.hf_config
        is_moe = any(cfg.get(k) for k in ("num_local_experts", "num_experts", "n_routed_experts"))
        self.lockstep = bool(is_moe and pc.enable_expert_parallel and self.dp_size > 1)
        self.coord: Optional[DPCoordinator] = None
        if self.lockstep:
            self.coord = DPCoordinator(pc.data_parallel_master_ip, pc.data_parallel_master_port, self.dp_rank,
                                       self.dp_size)
        if self.lockstep and self.batch_queue is not None:
            # one forward pass (real or dummy) per agreement round, in the same order on every rank: no steps in flight
            logger.info("wide expert parallelism: running the engine step synchronously (no batch queue)")
            self.batch_queue, self.step_fn = None, self.step
        self.wave = 0
        self.steps_in_wave = 0
        logger.info("DP engine %d/%d up (lock-step=%s)", self.dp_rank, self.dp_size, self.lockstep)

    def run_busy_loop(self) -> None:
        if not self.lockstep:
            return super().run_busy_loop()
        running = False          # inside a wave: at least one rank has work
        while True:
            local = self.scheduler.has_requests()
            block = not local and not running
            while True:
                try:
                    kind, payload = self.input_queue.get(block=block, timeout=0.005 if block else 0)
                except queue.Empty:
                    break
                block = False
                if kind == REQ_SHUTDOWN:
                    return
                self._handle(kind, payload)
            local = self.scheduler.has_requests()
            # exactly one agreement and at most one forward pass (real or dummy) per iteration on EVERY rank: the
            # expert-parallel collectives inside the forward pass need all ranks, in the same order
            running = self.coord.any_unfinished(local)
            if local:
                outs, executed = self.step_fn()
                for ci, o in (outs or {}).items():
                    o.engine_index = self.engine_index
                    self.output_queue.put((ci, o))
                if not executed:
                    # bookkeeping-only step (e.g. the last request just finished): peers still run a forward pass
                    self.model_executor.collective_rpc("execute_dummy_batch")
            elif running:
                self.model_executor.collective_rpc("execute_dummy_batch")
            else:
                time.sleep(0.002)            # nobody has work: idle cadence of the agreement
                continue
            self.steps_in_wave += 1
            if not running:
                self.wave += 1
                self.steps_in_wave = 0
