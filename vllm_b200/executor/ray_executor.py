"""Ray executor: one Ray actor per GPU, across the nodes of a Ray cluster (``--distributed-executor-backend ray``).

Behavioural reference: vllm/v1/executor/ray_executor.py (``RayDistributedExecutor``: placement group with one GPU
bundle per worker, actors sorted so that the ranks of a node are contiguous, driver-less ``collective_rpc`` over
``ray.get``), ray_utils.py (``initialize_ray_cluster``, ``RayWorkerWrapper``).

Ray only places and supervises the processes.  The data path is the same as for the ``mp`` executor: the actors form
one ``torch.distributed`` NCCL world (rendezvous on the node of rank 0) and run the same ``Worker``; scheduler outputs
travel as one pickled object per step and only the output rank replies with a result.
"""
from __future__ import annotations

import threading
from concurrent.futures import Future
from typing import Any, Callable, Optional, Union

from vllm_b200.core.sched_output import SchedulerOutput
from vllm_b200.executor import Executor, _run_method
from vllm_b200.logger import init_logger

logger = init_logger(__name__)


def _import_ray():
    try:
        import ray
    except ImportError as e:                                    # pragma: no cover - exercised without ray installed
        raise ImportError("the ray executor needs the `ray` package; use --distributed-executor-backend mp on one node "
                          "or `serve --headless` engine nodes for multi-node data parallelism") from e
    return ray


class RayWorkerActor:
    """Body of the actor (wrapped with ``ray.remote`` at run time so importing this module does not need ray)."""

    def __init__(self, vllm_config) -> None:
        self.vllm_config = vllm_config
        self.worker = None

    def node_and_gpus(self) -> tuple[str, list[int]]:
        ray = _import_ray()
        ip = ray.util.get_node_ip_address()
        try:
            gpus = [int(g) for g in ray.get_gpu_ids()]
        except Exception:                                       # CPU-only actor
            gpus = []
        return ip, gpus

    def init_worker(self, rank: int, local_rank: int, init_method: str) -> None:
        from vllm_b200.worker import resolve_worker_class
        # Ray narrows CUDA_VISIBLE_DEVICES to the actor's GPU: the device index inside the actor is 0
        self.worker = resolve_worker_class(self.vllm_config)(self.vllm_config, local_rank=local_rank, rank=rank, distributed_init_method=init_method,
                             is_driver_worker=(rank == 0))
        self.worker.init_device()
        self.worker.load_model()

    def run(self, method: Union[str, Callable], args: tuple, kwargs: dict) -> Any:
        return _run_method(self.worker, method, args, kwargs)

    def execute_model(self, so: SchedulerOutput, reply: bool):
        out = self.worker.execute_model(so)
        return out if reply else None


class RayExecutor(Executor):
    uses_multiproc = True

    def _init_executor(self) -> None:
        ray = _import_ray()
        pc = self.vllm_config.parallel_config
        world = pc.world_size
        if not ray.is_initialized():
            ray.init(address=getattr(pc, "ray_address", None) or None, ignore_reinit_error=True)
        use_gpu = self.vllm_config.device_config.resolve() == "cuda"
        pg = getattr(pc, "placement_group", None)
        if pg is None:
            from ray.util.placement_group import placement_group
            bundles = [{"GPU": 1.0, "CPU": 1.0} if use_gpu else {"CPU": 1.0} for _ in range(world)]
            # PACK keeps a TP group on as few nodes as possible (NVSwitch domain first, network second)
            pg = placement_group(bundles, strategy="PACK")
            ray.get(pg.ready())
        self.placement_group = pg
        from ray.util.scheduling_strategies import PlacementGroupSchedulingStrategy
        actor_cls = ray.remote(RayWorkerActor)
        actors = [actor_cls.options(num_gpus=1 if use_gpu else 0, num_cpus=1,
                                    scheduling_strategy=PlacementGroupSchedulingStrategy(
                                        placement_group=pg, placement_group_bundle_index=i)).remote(self.vllm_config)
                  for i in range(world)]
        # ranks of one node must be contiguous (TP inside a node); the node that hosts this driver comes first
        where = ray.get([a.node_and_gpus.remote() for a in actors])
        driver_ip = ray.util.get_node_ip_address()
        order = sorted(range(world), key=lambda i: (where[i][0] != driver_ip, where[i][0], where[i][1][:1] or [i]))
        self.actors = [actors[i] for i in order]
        ips = [where[i][0] for i in order]
        from vllm_b200.utils import get_open_port
        init_method = f"tcp://{ips[0]}:{get_open_port()}"
        local_ranks = [ips[:r].count(ips[r]) for r in range(world)]
        self.node_ips = ips
        ray.get([a.init_worker.remote(r, 0 if use_gpu else local_ranks[r], init_method)
                 for r, a in enumerate(self.actors)])
        self._ray = ray
        self._out_rank = world - pc.tensor_parallel_size
        logger.info("ray executor: %d workers on %d node(s)", world, len(set(ips)))

    @property
    def output_rank(self) -> int:
        return self._out_rank

    def _resolve(self, refs, unique_reply_rank: Optional[int]):
        try:
            res = self._ray.get(refs)
        except Exception:
            self.is_failed = True
            if self.failure_callback is not None:
                self.failure_callback()
            raise
        return res[unique_reply_rank] if unique_reply_rank is not None else res

    def _maybe_async(self, refs, non_block: bool, unique_reply_rank: Optional[int]):
        if not non_block:
            return self._resolve(refs, unique_reply_rank)
        fut: Future = Future()

        def wait() -> None:
            try:
                fut.set_result(self._resolve(refs, unique_reply_rank))
            except Exception as e:                              # noqa: BLE001 - surfaced through the future
                fut.set_exception(e)
        threading.Thread(target=wait, daemon=True).start()
        return fut

    def collective_rpc(self, method, args=(), kwargs=None, non_block=False, unique_reply_rank=None):
        refs = [a.run.remote(method, tuple(args), kwargs or {}) for a in self.actors]
        return self._maybe_async(refs, non_block, unique_reply_rank)

    def execute_model(self, so: SchedulerOutput, non_block: bool = False):
        refs = [a.execute_model.remote(so, r == self._out_rank) for r, a in enumerate(self.actors)]
        return self._maybe_async(refs, non_block, self._out_rank)

    def check_health(self) -> None:
        self._ray.get([a.run.remote("check_health", (), {}) for a in self.actors])

    def shutdown(self) -> None:
        ray = getattr(self, "_ray", None)
        for a in getattr(self, "actors", []):
            try:
                ray.get(a.run.remote("shutdown", (), {}))
                ray.kill(a)
            except Exception:                                   # noqa: BLE001 - the actor may already be gone
                pass
        self.actors = []
