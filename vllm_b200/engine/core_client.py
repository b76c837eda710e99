"""Clients of the engine core: in-process, sync multiprocess, asyncio multiprocess,
and data-parallel load-balancing variants.

Behavioural reference: vllm/v1/engine/core_client.py (``InprocClient`` :306,
``SyncMPClient`` :806, ``AsyncMPClient`` :978, ``DPLBAsyncMPClient`` :1435,
engine-death monitor :711).
"""
from __future__ import annotations

import asyncio
import multiprocessing
import os
import queue
import threading
import time
import weakref
from typing import Any, Optional

import msgspec

from vllm_b200 import envs
from vllm_b200.config import VllmConfig
from vllm_b200.core.request import EngineCoreOutputs, EngineCoreRequest
from vllm_b200.engine.core import (ENGINE_CORE_DEAD, REQ_ABORT, REQ_ADD, REQ_SHUTDOWN, REQ_UTILITY, EngineCore,
                                   EngineCoreProc)
from vllm_b200.engine.exceptions import EngineDeadError
from vllm_b200.engine.serial import decode_outputs_frames, encode_any, encode_request_frames
from vllm_b200.logger import init_logger
from vllm_b200.utils import get_open_zmq_ipc_path

logger = init_logger(__name__)


class EngineCoreClient:
    @staticmethod
    def make_client(vllm_config: VllmConfig, multiprocess: bool, asyncio_mode: bool, log_stats: bool = False,
                    executor_class=None) -> "EngineCoreClient":
        if asyncio_mode:
            if vllm_config.parallel_config.data_parallel_size > 1 or vllm_config.parallel_config.enable_elastic_ep:
                return DPLBAsyncMPClient(vllm_config, log_stats)
            return AsyncMPClient(vllm_config, log_stats)
        if multiprocess:
            return SyncMPClient(vllm_config, log_stats)
        return InprocClient(vllm_config, log_stats, executor_class)

    def shutdown(self) -> None: ...
    def get_output(self) -> EngineCoreOutputs: raise NotImplementedError
    def add_request(self, r: EngineCoreRequest) -> None: raise NotImplementedError
    def abort_requests(self, ids: list[str]) -> None: raise NotImplementedError
    def call_utility(self, method: str, *args) -> Any: raise NotImplementedError
    def dp_engines_running(self) -> bool: return False

    # convenience wrappers
    def profile(self, is_start=True): return self.call_utility("profile", is_start)
    def reset_prefix_cache(self): return self.call_utility("reset_prefix_cache")
    def sleep(self, level=1): return self.call_utility("sleep", level)
    def wake_up(self, tags=None): return self.call_utility("wake_up", tags)
    def is_sleeping(self): return self.call_utility("is_sleeping")
    def save_sharded_state(self, path): return self.call_utility("save_sharded_state", path)
    def collective_rpc(self, method, args=(), kwargs=None):
        return self.call_utility("collective_rpc", method, args, kwargs)


class InprocClient(EngineCoreClient):
    """Everything in one process (no busy loop): ``get_output`` runs one step."""

    def __init__(self, vllm_config: VllmConfig, log_stats: bool = False, executor_class=None) -> None:
        self.engine_core = EngineCore(vllm_config, executor_class, log_stats)

    def get_output(self) -> EngineCoreOutputs:
        outs, _ = self.engine_core.step_fn()
        return (outs or {}).get(0) or EngineCoreOutputs()

    def add_request(self, r: EngineCoreRequest) -> None:
        self.engine_core.add_request(r)

    def abort_requests(self, ids: list[str]) -> None:
        if ids:
            self.engine_core.abort_requests(ids)

    def call_utility(self, method: str, *args) -> Any:
        return getattr(self.engine_core, method)(*args)

    def shutdown(self) -> None:
        self.engine_core.shutdown()


class _Resources:
    """Things that must be torn down even if the client is garbage collected."""

    def __init__(self) -> None:
        self.procs: list[multiprocessing.Process] = []
        self.sockets: list[Any] = []
        self.ctx = None
        self.engine_dead = False

    def __call__(self) -> None:
        for s in self.sockets:
            try:
                s.close(linger=0)
            except Exception:
                pass
        for p in self.procs:
            if p.is_alive():
                p.terminate()
        for p in self.procs:
            p.join(timeout=5)
            if p.is_alive():
                p.kill()


def _launch_engines(vllm_config: VllmConfig, input_addr: str, output_addr: str, n: int, log_stats: bool,
                    start: int = 0) -> list[multiprocessing.Process]:
    ctx = multiprocessing.get_context(envs.VLLM_WORKER_MULTIPROC_METHOD)
    procs = []
    for i in range(start, start + n):
        p = ctx.Process(target=EngineCoreProc.run_engine_core, name=f"EngineCore_{i}",
                        kwargs=dict(vllm_config=vllm_config, input_addr=input_addr, output_addr=output_addr,
                                    engine_index=i, log_stats=log_stats, dp_rank=i, parent_pid=os.getpid()))
        p.start()
        procs.append(p)
    return procs


class MPClient(EngineCoreClient):
    """Shared plumbing: ROUTER (inputs to engines) + PULL (outputs from engines)."""

    def __init__(self, vllm_config: VllmConfig, log_stats: bool, asyncio_mode: bool) -> None:
        import zmq
        import zmq.asyncio
        self.vllm_config = vllm_config
        self.res = _Resources()
        self._finalizer = weakref.finalize(self, self.res)
        self.ctx = zmq.asyncio.Context() if asyncio_mode else zmq.Context()
        self.res.ctx = self.ctx
        # "k|in_addr|out_addr": this process is API server k of several in front of engines somebody else launched
        # (``serve --api-server-count N``; reference: v1/utils.py APIServerProcessManager + client_index routing)
        external = os.environ.get("VLLM_B200_EXTERNAL_ENGINE")
        self.external_engines = bool(external)
        if external:
            idx, in_addr, out_addr = external.split("|")
            self.client_index = int(idx)
        else:
            in_addr, out_addr = get_open_zmq_ipc_path(), get_open_zmq_ipc_path()
            self.client_index = 0
        self.input_socket = self.ctx.socket(zmq.ROUTER)
        self.input_socket.bind(in_addr)
        self.output_socket = self.ctx.socket(zmq.PULL)
        self.output_socket.bind(out_addr)
        self.res.sockets += [self.input_socket, self.output_socket]
        self.num_engines = vllm_config.parallel_config.data_parallel_size
        self.engine_ids = [i.to_bytes(2, "little") for i in range(self.num_engines)]
        self.in_addr, self.out_addr, self.log_stats = in_addr, out_addr, log_stats
        if not external:
            self.res.procs = _launch_engines(vllm_config, in_addr, out_addr, self.num_engines, log_stats)
        self._wait_ready(in_addr)
        self.utility_results: dict[int, Any] = {}
        self._call_id = 0

    def _wait_ready(self, in_addr: str) -> None:
        """Startup handshake with a liveness check on the child processes."""
        import zmq
        ready = 0
        # the handshake runs before any event loop exists: use a plain blocking view of the (possibly asyncio) socket
        s = zmq.Socket.shadow(self.input_socket.underlying) if hasattr(self.input_socket, "underlying") else self.input_socket
        deadline = time.time() + 3600
        while ready < self.num_engines:
            if not s.poll(timeout=1000):
                for p in self.res.procs:
                    if not p.is_alive():
                        raise RuntimeError(f"engine core {p.name} died during startup (exit {p.exitcode})")
                if time.time() > deadline:
                    raise TimeoutError("engine core startup timed out")
                continue
            ident, msg = s.recv_multipart()
            if msg == b"HELLO":
                continue
            info = msgspec.msgpack.decode(msg)
            if info.get("status") == "READY":
                ready += 1
                self.vllm_config.cache_config.num_gpu_blocks = info.get("num_gpu_blocks")

    def _next_call_id(self) -> int:
        self._call_id += 1
        return self._call_id

    def _check_alive(self) -> None:
        if self.res.engine_dead:
            raise EngineDeadError()
        for p in self.res.procs:
            if not p.is_alive():
                self.res.engine_dead = True
                raise EngineDeadError(f"engine core process {p.name} exited with code {p.exitcode}")

    def shutdown(self) -> None:
        try:
            for eid in ([] if self.external_engines else self.engine_ids):
                s = self.input_socket
                if hasattr(s, "underlying"):
                    import zmq
                    s = zmq.Socket.shadow(s.underlying)
                s.send_multipart([eid, REQ_SHUTDOWN], flags=1)  # NOBLOCK
        except Exception:
            pass
        time.sleep(0.05)
        self._finalizer()


class SyncMPClient(MPClient):
    def __init__(self, vllm_config: VllmConfig, log_stats: bool = False) -> None:
        super().__init__(vllm_config, log_stats, asyncio_mode=False)
        self.outputs_queue: "queue.Queue[Any]" = queue.Queue()
        self._stop = False
        t = threading.Thread(target=self._recv_loop, daemon=True, name="EngineCoreOutputThread")
        t.start()
        self._thread = t

    def _recv_loop(self) -> None:
        import zmq
        while not self._stop:
            try:
                if not self.output_socket.poll(timeout=500):
                    if any(not p.is_alive() for p in self.res.procs):
                        self.res.engine_dead = True
                        self.outputs_queue.put(EngineDeadError())
                        return
                    continue
                frames = self.output_socket.recv_multipart(copy=False)
            except zmq.ZMQError:
                return
            if len(frames) == 1 and frames[0].bytes == ENGINE_CORE_DEAD:
                self.res.engine_dead = True
                self.outputs_queue.put(EngineDeadError())
                return
            _, outs = decode_outputs_frames(frames)
            if outs.utility_output is not None:
                cid, err, result = outs.utility_output
                self.utility_results[cid] = (err, result)
                ev = self._utility_events.pop(cid, None)
                if ev:
                    ev.set()
                if not outs.outputs:
                    continue
            self.outputs_queue.put(outs)

    _utility_events: dict[int, threading.Event] = {}

    def get_output(self) -> EngineCoreOutputs:
        o = self.outputs_queue.get()
        if isinstance(o, Exception):
            raise o
        return o

    def add_request(self, r: EngineCoreRequest) -> None:
        self._check_alive()
        r.client_index = self.client_index
        self.input_socket.send_multipart([self.engine_ids[0], REQ_ADD, *encode_request_frames(r)], copy=False)

    def abort_requests(self, ids: list[str]) -> None:
        if ids and not self.res.engine_dead:
            self.input_socket.send_multipart([self.engine_ids[0], REQ_ABORT, encode_any(ids)])

    def call_utility(self, method: str, *args) -> Any:
        self._check_alive()
        cid = self._next_call_id()
        ev = threading.Event()
        self._utility_events[cid] = ev
        self.input_socket.send_multipart(
            [self.engine_ids[0], REQ_UTILITY, encode_any((self.client_index, cid, method, list(args)))])
        while not ev.wait(timeout=1.0):
            self._check_alive()
        err, result = self.utility_results.pop(cid)
        if err:
            raise RuntimeError(f"engine utility {method} failed: {err}")
        return result

    def shutdown(self) -> None:
        self._stop = True
        super().shutdown()


class AsyncMPClient(MPClient):
    def __init__(self, vllm_config: VllmConfig, log_stats: bool = False) -> None:
        super().__init__(vllm_config, log_stats, asyncio_mode=True)
        self.outputs_queue: Optional[asyncio.Queue] = None
        self._task: Optional[asyncio.Task] = None
        self._utility_futs: dict[int, asyncio.Future] = {}

    def _ensure_task(self) -> None:
        if self._task is None:
            self.outputs_queue = asyncio.Queue()
            self._task = asyncio.get_running_loop().create_task(self._recv_loop())

    async def _recv_loop(self) -> None:
        while True:
            try:
                ev = await self.output_socket.poll(timeout=1000)
            except Exception:
                return
            if not ev:
                if any(not p.is_alive() for p in self.res.procs):
                    self.res.engine_dead = True
                    await self.outputs_queue.put(EngineDeadError())
                    for f in self._utility_futs.values():
                        if not f.done():
                            f.set_exception(EngineDeadError())
                    return
                continue
            frames = await self.output_socket.recv_multipart(copy=False)
            if len(frames) == 1 and frames[0].bytes == ENGINE_CORE_DEAD:
                self.res.engine_dead = True
                await self.outputs_queue.put(EngineDeadError())
                return
            _, outs = decode_outputs_frames(frames)
            self._on_outputs(outs)
            if outs.utility_output is not None:
                cid, err, result = outs.utility_output
                f = self._utility_futs.pop(cid, None)
                if f is not None and not f.done():
                    if err:
                        f.set_exception(RuntimeError(f"engine utility failed: {err}"))
                    else:
                        f.set_result(result)
                if not outs.outputs:
                    continue
            await self.outputs_queue.put(outs)

    def _on_outputs(self, outs: EngineCoreOutputs) -> None:
        pass

    def _pick_engine(self, r: EngineCoreRequest) -> bytes:
        return self.engine_ids[0]

    async def get_output_async(self) -> EngineCoreOutputs:
        self._ensure_task()
        o = await self.outputs_queue.get()
        if isinstance(o, Exception):
            raise o
        return o

    async def add_request_async(self, r: EngineCoreRequest) -> None:
        self._ensure_task()
        self._check_alive()
        r.client_index = self.client_index
        eid = self._pick_engine(r)
        await self.input_socket.send_multipart([eid, REQ_ADD, *encode_request_frames(r)], copy=False)

    async def abort_requests_async(self, ids: list[str]) -> None:
        if ids and not self.res.engine_dead:
            for eid in self._engines_for(ids):
                await self.input_socket.send_multipart([eid, REQ_ABORT, encode_any(ids)])

    def _engines_for(self, ids: list[str]) -> list[bytes]:
        return self.engine_ids

    async def call_utility_async(self, method: str, *args, engine: Optional[bytes] = None) -> Any:
        self._ensure_task()
        self._check_alive()
        targets = [engine] if engine is not None else self.engine_ids
        futs = []
        for eid in targets:
            cid = self._next_call_id()
            f = asyncio.get_running_loop().create_future()
            self._utility_futs[cid] = f
            futs.append(f)
            await self.input_socket.send_multipart(
                [eid, REQ_UTILITY, encode_any((self.client_index, cid, method, list(args)))])
        res = await asyncio.gather(*futs)
        return res[0]

    def shutdown(self) -> None:
        if self._task is not None:
            self._task.cancel()
        super().shutdown()


class DPLBAsyncMPClient(AsyncMPClient):
    """Data-parallel front-end with internal load balancing: each request goes to the engine
    with the fewest in-flight requests (weighted towards waiting over running, as in the
    reference core_client.py:1435-1520); aborts are routed to the owner."""

    def __init__(self, vllm_config: VllmConfig, log_stats: bool = False) -> None:
        super().__init__(vllm_config, log_stats)
        self.inflight: list[int] = [0] * self.num_engines
        self.engines_seen: set[int] = set()
        self.active = self.num_engines          # engines [0, active) receive new requests (elastic scale-down drains)
        self.owner: dict[str, int] = {}
        self._rr = 0

    def _pick_engine(self, r: EngineCoreRequest) -> bytes:
        if r.data_parallel_rank is not None:
            i = r.data_parallel_rank
        else:
            m = min(self.inflight[:self.active])
            cands = [i for i, v in enumerate(self.inflight[:self.active]) if v == m]
            i = cands[self._rr % len(cands)]
            self._rr += 1
        self.inflight[i] += 1
        self.owner[r.request_id] = i
        self.engines_seen.add(i)
        return self.engine_ids[i]

    def _on_outputs(self, outs: EngineCoreOutputs) -> None:
        for o in outs.outputs:
            if o.finish_reason is not None:
                i = self.owner.pop(o.request_id, None)
                if i is not None:
                    self.inflight[i] = max(self.inflight[i] - 1, 0)

    def _engines_for(self, ids: list[str]) -> list[bytes]:
        return [self.engine_ids[i] for i in sorted({self.owner[r] for r in ids if r in self.owner})] \
            or self.engine_ids

    # ---------------------------------------------------------------- elastic scaling
    async def scale_elastic(self, new_size: int, timeout_s: float = 600.0) -> int:
        """Grow or shrink the number of engine-core replicas at runtime (reference: ``scale_elastic_ep``,
        async_llm.py:1051-1094 — drain → reconfigure).  Growing spawns new engine processes against the same
        sockets and waits for their READY handshake; shrinking stops routing to the highest replicas, waits until
        their in-flight requests finish, then shuts them down."""
        import copy
        if new_size < 1:
            raise ValueError("at least one engine replica is required")
        cur = self.num_engines
        if new_size > cur:
            cfg = copy.deepcopy(self.vllm_config)
            cfg.parallel_config.data_parallel_size = new_size
            procs = _launch_engines(cfg, self.in_addr, self.out_addr, new_size - cur, self.log_stats, start=cur)
            self.res.procs += procs
            ready, deadline = 0, time.time() + timeout_s
            while ready < new_size - cur:
                if not await self.input_socket.poll(timeout=500):
                    if any(not p.is_alive() for p in procs):
                        raise RuntimeError("a new engine replica died during start-up")
                    if time.time() > deadline:
                        raise TimeoutError("new engine replicas did not come up")
                    continue
                _ident, msg = await self.input_socket.recv_multipart()
                if msg != b"HELLO" and msgspec.msgpack.decode(msg).get("status") == "READY":
                    ready += 1
            self.engine_ids += [i.to_bytes(2, "little") for i in range(cur, new_size)]
            self.inflight += [0] * (new_size - cur)
            self.num_engines = self.active = new_size
        elif new_size < cur:
            self.active = new_size
            deadline = time.time() + timeout_s
            while any(self.inflight[new_size:]) and time.time() < deadline:
                await asyncio.sleep(0.05)
            for eid in self.engine_ids[new_size:]:
                await self.input_socket.send_multipart([eid, REQ_SHUTDOWN])
            for p in self.res.procs[new_size:]:
                p.join(timeout=10)
                if p.is_alive():
                    p.terminate()
            del self.res.procs[new_size:], self.engine_ids[new_size:], self.inflight[new_size:]
            self.num_engines = new_size
        return self.num_engines
