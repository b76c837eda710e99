"""One worker process per GPU of the node, driven from the engine-core process.

Behavioural reference: vllm/v1/executor/multiproc_executor.py (``MultiprocExecutor`` :111, worker
``worker_busy_loop`` :1029, death monitoring, ``collective_rpc`` with ``unique_reply_rank``).

Control plane: ONE shared-memory ring (``parallel/shm_broadcast.ShmRing``, native spin-wait) broadcasts every
RPC — including each step's ``SchedulerOutput`` — to all workers; each worker answers on its own
single-reader ring.  Workers keep the step pipeline full: ``execute_model`` is launched asynchronously
and a responder thread performs the device→host wait, so the worker can already accept step N+1.
"""
from __future__ import annotations

import multiprocessing
import os
import pickle
import queue
import threading
import time
import traceback
from concurrent.futures import Future
from typing import Any, Callable, Optional, Union

from vllm_b200 import envs
from vllm_b200.config import VllmConfig
from vllm_b200.core.sched_output import SchedulerOutput
from vllm_b200.executor import Executor, _run_method
from vllm_b200.logger import init_logger
from vllm_b200.parallel.shm_broadcast import ShmRing
from vllm_b200.utils import get_open_port

logger = init_logger(__name__)

_SLOT = 4 << 20


def _worker_main(vllm_config: VllmConfig, rank: int, local_rank: int, init_method: str, rpc_handle: dict,
                 resp_handle: dict, ready) -> None:
    from vllm_b200.logger import set_process_prefix
    set_process_prefix(f"Worker_{rank}")
    rpc = ShmRing.attach(rpc_handle, reader_rank=rank)
    resp = ShmRing.attach(resp_handle)
    out_q: "queue.Queue[Any]" = queue.Queue()

    def responder() -> None:
        while True:
            item = out_q.get()
            if item is None:
                return
            ok, val = item
            if isinstance(val, _Drain):
                try:
                    val.inner.get()
                except Exception:
                    pass
                continue
            try:
                if ok and hasattr(val, "get") and not isinstance(val, dict):
                    val = val.get()               # device → host wait of an async step
                payload = pickle.dumps((ok, val), protocol=pickle.HIGHEST_PROTOCOL)
            except Exception:
                payload = pickle.dumps((False, traceback.format_exc()))
            resp.put_bytes(payload)

    worker = None
    try:
        from vllm_b200.worker import resolve_worker_class
        worker = resolve_worker_class(vllm_config)(vllm_config, local_rank=local_rank, rank=rank, distributed_init_method=init_method,
                        is_driver_worker=rank == 0)
        worker.init_device()
        worker.load_model()
        ready.send(("ok", rank))
    except Exception:
        ready.send(("error", traceback.format_exc()))
        return
    t = threading.Thread(target=responder, daemon=True, name="responder")
    t.start()
    try:
        while True:
            method, args, kwargs, reply_rank = rpc.get()
            if method == "__shutdown__":
                break
            try:
                if isinstance(method, tuple):
                    import cloudpickle
                    method = cloudpickle.loads(method[1])
                if method == "execute_model":
                    res = worker.execute_model(*args, non_block=True)
                else:
                    res = _run_method(worker, method, args, kwargs)
                ok = True
            except Exception:
                res, ok = traceback.format_exc(), False
            if reply_rank is None or reply_rank == rank or not ok:
                out_q.put((ok, res))
            elif hasattr(res, "get"):
                # not the replying rank: still drain the async step so buffers are recycled in order
                out_q.put((ok, _Drain(res)))
    finally:
        out_q.put(None)
        t.join(timeout=5)
        try:
            worker.shutdown()
        except Exception:
            pass
        rpc.close()
        resp.close()


class _Drain:
    """Wraps an async step of a non-replying rank: waited for, never sent."""

    def __init__(self, inner) -> None:
        self.inner = inner


class MultiprocExecutor(Executor):
    uses_multiproc = True

    def _init_executor(self) -> None:
        pc = self.vllm_config.parallel_config
        self.world = pc.world_size
        ctx = multiprocessing.get_context(envs.VLLM_WORKER_MULTIPROC_METHOD)
        port = get_open_port()
        init_method = f"tcp://127.0.0.1:{port}"
        self.rpc = ShmRing(self.world, slot_bytes=_SLOT, n_slots=4)
        self.resp = [ShmRing(1, slot_bytes=_SLOT, n_slots=4) for _ in range(self.world)]
        for r in self.resp:
            r.reader_rank = 0
        self.procs = []
        pipes = []
        for rank in range(self.world):
            rx, tx = ctx.Pipe(duplex=False)
            p = ctx.Process(target=_worker_main, name=f"Worker_{rank}", daemon=True,
                            args=(self.vllm_config, rank, rank, init_method, self.rpc.handle(),
                                  self.resp[rank].handle(), tx))
            p.start()
            self.procs.append(p)
            pipes.append(rx)
        for rank, rx in enumerate(pipes):
            deadline = time.time() + 600
            while not rx.poll(0.5):
                if not self.procs[rank].is_alive():
                    self.shutdown()
                    raise RuntimeError(f"worker {rank} died during start-up")
                if time.time() > deadline:
                    self.shutdown()
                    raise TimeoutError(f"worker {rank} did not come up")
            status, info = rx.recv()
            if status != "ok":
                self.shutdown()
                raise RuntimeError(f"worker {rank} failed to initialise:\n{info}")
        self._pending: "queue.Queue[Optional[tuple[Future, Optional[int]]]]" = queue.Queue()
        self._lock = threading.Lock()
        self._collector = threading.Thread(target=self._collect, daemon=True, name="mp-executor-collector")
        self._collector.start()
        self._monitor = threading.Thread(target=self._watch, daemon=True, name="mp-executor-monitor")
        self._monitor.start()
        self._closed = False

    # ------------------------------------------------------------------ plumbing
    def _read(self, rank: int):
        while True:
            try:
                return pickle.loads(self.resp[rank].get_bytes(timeout=1.0))
            except TimeoutError:
                if self.is_failed or not self.procs[rank].is_alive():
                    raise RuntimeError(f"worker {rank} died")

    def _collect(self) -> None:
        while True:
            item = self._pending.get()
            if item is None:
                return
            fut, unique = item
            try:
                ranks = [unique] if unique is not None else list(range(self.world))
                vals = []
                for r in ranks:
                    ok, val = self._read(r)
                    if not ok:
                        raise RuntimeError(f"worker {r} raised:\n{val}")
                    vals.append(val)
                fut.set_result(vals[0] if unique is not None else vals)
            except Exception as e:
                self.is_failed = True
                fut.set_exception(e)

    def _watch(self) -> None:
        from multiprocessing.connection import wait
        sentinels = {p.sentinel: i for i, p in enumerate(self.procs)}
        while not getattr(self, "_closed", False):
            dead = wait(list(sentinels), timeout=1.0)
            if dead and not getattr(self, "_closed", False):
                self.is_failed = True
                logger.error("worker %s exited unexpectedly", [sentinels[d] for d in dead])
                if self.failure_callback is not None:
                    self.failure_callback()
                return

    def collective_rpc(self, method: Union[str, Callable], args: tuple = (), kwargs: Optional[dict] = None,
                       non_block: bool = False, unique_reply_rank: Optional[int] = None) -> Any:
        if self.is_failed:
            raise RuntimeError("executor failed")
        if not isinstance(method, str):
            import cloudpickle
            method = ("__callable__", cloudpickle.dumps(method))
        fut: Future = Future()
        with self._lock:      # one broadcast at a time; responses are consumed in the same order
            self.rpc.put((method, args, kwargs or {}, unique_reply_rank))
            self._pending.put((fut, unique_reply_rank))
        return fut if non_block else fut.result()

    def execute_model(self, so: SchedulerOutput, non_block: bool = False):
        return self.collective_rpc("execute_model", (so,), non_block=non_block, unique_reply_rank=self.output_rank)

    def shutdown(self) -> None:
        if getattr(self, "_closed", False):
            return
        self._closed = True
        try:
            self.rpc.put(("__shutdown__", (), {}, None), timeout=2.0)
        except Exception:
            pass
        if hasattr(self, "_pending"):
            self._pending.put(None)
        for p in getattr(self, "procs", []):
            p.join(timeout=5)
            if p.is_alive():
                p.terminate()
        for r in [getattr(self, "rpc", None), *getattr(self, "resp", [])]:
            if r is not None:
                r.close()
