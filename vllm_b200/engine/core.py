"""EngineCore: scheduler + executor loop; optionally in its own process behind ZMQ.

Behavioural reference: vllm/v1/engine/core.py (``EngineCore`` :105,
``_initialize_kv_caches`` :254, ``step`` :597, ``step_with_batch_queue`` :638,
``EngineCoreProc`` :1021 with handshake :1143-1283, busy loop :1405, IO threads
:1688/:1791), DP variant :2000 (waves + dummy batches).
"""
from __future__ import annotations

import os
import queue
import signal
import threading
import time
from collections import deque
from concurrent.futures import Future
from typing import Any, Callable, Optional

import msgspec

from vllm_b200.config import VllmConfig
from vllm_b200.core.kv_cache_manager import get_kv_cache_config
from vllm_b200.core.request import (EngineCoreOutputs, EngineCoreRequest, Request, RequestStatus)
from vllm_b200.core.scheduler import Scheduler
from vllm_b200.executor import Executor
from vllm_b200.logger import init_logger

logger = init_logger(__name__)

# wire request types
REQ_ADD, REQ_ABORT, REQ_UTILITY, REQ_START_DP_WAVE, REQ_SHUTDOWN = b"A", b"B", b"U", b"W", b"S"
ENGINE_CORE_DEAD = b"ENGINE_CORE_DEAD"


class EngineCore:
    def __init__(self, vllm_config: VllmConfig, executor_class: Optional[type[Executor]] = None,
                 log_stats: bool = False) -> None:
        self.vllm_config = vllm_config
        self.log_stats = log_stats
        executor_class = executor_class or Executor.get_class(vllm_config)
        self.model_executor = executor_class(vllm_config)
        kv_cfg = self._initialize_kv_caches()
        from vllm_b200.structured_output import StructuredOutputManager
        self.structured_output_manager = StructuredOutputManager(vllm_config)
        connector = None
        if vllm_config.kv_transfer_config is not None and vllm_config.kv_transfer_config.kv_connector:
            from vllm_b200.kv_transfer import create_connector
            connector = create_connector(vllm_config, role="scheduler")
        sched_cls = Scheduler
        if vllm_config.scheduler_config.scheduler_cls:
            from vllm_b200.utils import resolve_obj_by_qualname
            sched_cls = resolve_obj_by_qualname(vllm_config.scheduler_config.scheduler_cls)
        self.scheduler: Scheduler = sched_cls(
            vllm_config, kv_cfg, structured_output_manager=self.structured_output_manager,
            kv_connector=connector, include_finished_set=vllm_config.parallel_config.data_parallel_size > 1,
            log_stats=log_stats)
        self.block_hasher = self.scheduler.get_block_hasher()
        n_batches = self.model_executor.max_concurrent_batches
        self.batch_queue: Optional[deque] = deque(maxlen=n_batches) if n_batches > 1 else None
        self.batch_queue_size = n_batches
        self.step_fn: Callable = self.step if self.batch_queue is None else self.step_with_batch_queue
        self._aborts_since_step: list[str] = []

    def _initialize_kv_caches(self):
        t0 = time.perf_counter()
        ex = self.model_executor
        specs = ex.get_kv_cache_specs()
        avail = ex.determine_available_memory()
        mm = min(avail)
        cc = self.vllm_config.cache_config
        max_len = self.vllm_config.scheduler_config.max_model_len
        # one config per worker (pipeline stages own different layers); every worker must end up with the same
        # number of blocks because block ids are global to the scheduler
        cfgs = [get_kv_cache_config(sp, mm, max_len, cc.num_gpu_blocks_override, auto_fit_max_model_len=False)[0]
                for sp in specs]
        nblk = min(c.num_blocks for c in cfgs)
        if any(c.num_blocks != nblk for c in cfgs):
            cfgs = [get_kv_cache_config(sp, mm, max_len, nblk, auto_fit_max_model_len=False)[0] for sp in specs]
        cfg = cfgs[0]
        cc.num_gpu_blocks = cfg.num_blocks
        ex.initialize_from_config(cfgs if len(cfgs) > 1 else cfg)
        logger.info("KV cache: %d blocks x %d tokens (%.2f GiB available), init %.2fs", cfg.num_blocks,
                    cc.block_size, mm / 2**30, time.perf_counter() - t0)
        return cfg

    # ----------------------------------------------------------------- API
    _mm_receiver = None

    def add_request(self, req: EngineCoreRequest) -> None:
        self._restore_mm(req)           # first: keeps the receiver cache in step with the sender even if we reject below
        so_req = None
        sp = req.sampling_params
        if sp is not None and sp.structured_outputs is not None:
            so_req = self.structured_output_manager.make_request(req.request_id, sp.structured_outputs)
        r = Request.from_engine_core_request(req, self.block_hasher, so_req)
        if len(r.prompt_token_ids) >= self.vllm_config.scheduler_config.max_model_len:
            raise ValueError(f"prompt of {len(r.prompt_token_ids)} tokens exceeds max_model_len")
        self.scheduler.add_request(r)

    def _restore_mm(self, req) -> None:
        if req.mm_features:
            if self._mm_receiver is None:
                from vllm_b200.multimodal.cache import ReceiverCache
                self._mm_receiver = ReceiverCache(int(os.environ.get("VLLM_B200_MM_CACHE_ITEMS", "256")))
            self._mm_receiver.restore(req.mm_features)

    def abort_requests(self, request_ids: list[str]) -> None:
        self.scheduler.finish_requests(request_ids, RequestStatus.FINISHED_ABORTED)

    def step(self) -> tuple[dict[int, EngineCoreOutputs], bool]:
        if not self.scheduler.has_requests():
            return {}, False
        so = self.scheduler.schedule()
        so.grammar_bitmask = self.scheduler.get_grammar_bitmask(so)
        try:
            mo = self.model_executor.execute_model(so)
        except Exception:
            self._dump_failed_step(so)
            raise
        out = self.scheduler.update_from_output(so, mo)
        if so.total_num_scheduled_tokens == 0 and not out:
            time.sleep(0.001)                    # only requests waiting on remote KV: do not spin the executor
        return out, so.total_num_scheduled_tokens > 0

    def _dump_failed_step(self, so) -> None:
        """Log what the failing forward pass was asked to do (reference: logging_utils/dump_input.py): request ids, token
        counts, positions in the sequence, speculative tokens and the engine configuration — not the prompts."""
        try:
            new = [(r.req_id, len(r.prompt_token_ids or ())) for r in so.scheduled_new_reqs]
            logger.error("engine step failed. config hash %s; %d tokens scheduled; new requests (id, prompt tokens): %s; "
                         "tokens per request: %s; spec tokens: %s; finished since last step: %s; scheduler: %d running, %d waiting",
                         self.vllm_config.compute_hash(), so.total_num_scheduled_tokens, new, dict(so.num_scheduled_tokens),
                         {k: len(v) for k, v in so.scheduled_spec_decode_tokens.items()}, sorted(so.finished_req_ids),
                         *self.scheduler.get_request_counts())
        except Exception:  # pragma: no cover - the dump must never mask the original error
            pass

    def step_with_batch_queue(self) -> tuple[Optional[dict[int, EngineCoreOutputs]], bool]:
        """Keep up to ``batch_queue_size`` steps in flight (PP stages / async scheduling)."""
        bq = self.batch_queue
        assert bq is not None
        scheduled = False
        if len(bq) < self.batch_queue_size and self.scheduler.has_requests():
            so = self.scheduler.schedule()
            if so.total_num_scheduled_tokens > 0 or so.finished_req_ids:
                so.grammar_bitmask = self.scheduler.get_grammar_bitmask(so)
                fut = self.model_executor.execute_model(so, non_block=True)
                bq.appendleft((fut, so))
                scheduled = so.total_num_scheduled_tokens > 0
                if scheduled and len(bq) < self.batch_queue_size and not bq[-1][0].done():
                    return None, True  # try to fill the pipeline before blocking
        if not bq:
            return None, False
        fut, so = bq.pop()
        mo = fut.result()
        return self.scheduler.update_from_output(so, mo), scheduled

    # -------------------------------------------------------------- utility
    def profile(self, is_start: bool = True) -> None:
        self.model_executor.profile(is_start)

    def reset_prefix_cache(self) -> bool:
        return self.scheduler.reset_prefix_cache()

    def sleep(self, level: int = 1) -> None:
        self.model_executor.sleep(level)

    def wake_up(self, tags=None) -> None:
        self.model_executor.wake_up(tags)

    def is_sleeping(self) -> bool:
        return all(self.model_executor.collective_rpc("is_sleeping"))

    def save_sharded_state(self, path: str) -> None:
        self.model_executor.save_sharded_state(path)

    def collective_rpc(self, method, args=(), kwargs=None):
        return self.model_executor.collective_rpc(method, args, kwargs)

    def get_request_counts(self) -> tuple[int, int]:
        return self.scheduler.get_request_counts()

    def echo(self, x: Any) -> Any:  # used by IPC tests, cf. test_engine_core_client.py in the reference
        return x

    def shutdown(self) -> None:
        if getattr(self, "model_executor", None) is not None:
            self.model_executor.shutdown()
        if getattr(self, "scheduler", None) is not None:
            self.scheduler.shutdown()


class EngineCoreProc(EngineCore):
    """EngineCore in a child process. Sockets: a DEALER for inputs (connects to the
    client's ROUTER) and a PUSH for outputs. Decode/encode runs on IO threads so the
    busy loop only touches in-process queues."""

    def __init__(self, vllm_config: VllmConfig, input_addr: str, output_addr: str, engine_index: int = 0,
                 log_stats: bool = False, executor_class=None) -> None:
        import zmq
        self.engine_index = engine_index
        self.input_queue: "queue.Queue[tuple[bytes, Any]]" = queue.Queue()
        self.output_queue: "queue.Queue[Any]" = queue.Queue()
        self._zctx = zmq.Context(io_threads=2)
        identity = engine_index.to_bytes(2, "little")
        # One (DEALER, PUSH) pair per front-end client: comma-separated address lists, position = client_index
        # (several API-server processes in front of the same engines; reference: --api-server-count).
        # handshake first so a slow model load does not look like a dead engine
        self._in_socks, self._out_socks = [], []
        for addr in input_addr.split(","):
            s = self._zctx.socket(zmq.DEALER)
            s.setsockopt(zmq.IDENTITY, identity)
            s.connect(addr)
            s.send(b"HELLO")
            self._in_socks.append(s)
        for addr in output_addr.split(","):
            s = self._zctx.socket(zmq.PUSH)
            s.connect(addr)
            self._out_socks.append(s)
        self._in_sock, self._out_sock = self._in_socks[0], self._out_socks[0]
        super().__init__(vllm_config, executor_class, log_stats)
        for s in self._in_socks:
            s.send(msgspec.msgpack.encode(
                {"status": "READY", "num_gpu_blocks": vllm_config.cache_config.num_gpu_blocks}))
        self._stop = False
        self._threads = [threading.Thread(target=self._input_loop, daemon=True),
                         threading.Thread(target=self._output_loop, daemon=True)]
        for t in self._threads:
            t.start()

    # ---- IO threads
    def _input_loop(self) -> None:
        from vllm_b200.engine.serial import decode_any, decode_request_frames
        import zmq
        poller = zmq.Poller()
        for s in self._in_socks:
            poller.register(s, zmq.POLLIN)
        while not self._stop:
            try:
                ready = poller.poll(500) if len(self._in_socks) > 1 else [(self._in_sock, zmq.POLLIN)]
                for sock, _ in ready:
                    frames = sock.recv_multipart(copy=False)
                    kind = frames[0].bytes
                    if kind == REQ_ADD:        # frame 1 = msgpack head, frames 2.. = out-of-band tensor storage (zero copy)
                        self.input_queue.put((kind, decode_request_frames(frames[1:])))
                    else:
                        self.input_queue.put((kind, decode_any(frames[1].bytes) if len(frames) > 1 else None))
            except Exception:
                break

    def _output_loop(self) -> None:
        from vllm_b200.engine.serial import encode_outputs_frames
        while True:
            item = self.output_queue.get()
            if item is None:
                break
            if isinstance(item, bytes):                      # ENGINE_CORE_DEAD: every front-end must hear it
                for s in self._out_socks:
                    s.send(item)
                continue
            ci = item[0] if 0 <= item[0] < len(self._out_socks) else 0
            self._out_socks[ci].send_multipart(encode_outputs_frames(item), copy=False)

    # ---- busy loop
    def run_busy_loop(self) -> None:
        while True:
            # drain inputs; block when idle
            block = not self.scheduler.has_requests() and not (self.batch_queue and len(self.batch_queue))
            while True:
                try:
                    kind, payload = self.input_queue.get(block=block, timeout=None if block else 0)
                except queue.Empty:
                    break
                block = False
                if kind == REQ_SHUTDOWN:
                    return
                self._handle(kind, payload)
            try:
                outs, _ = self.step_fn()
            except Exception as e:
                if not self._recover_from_fault(e):
                    raise
                continue
            for ci, o in (outs or {}).items():
                o.engine_index = self.engine_index
                self.output_queue.put((ci, o))

    # ---- fault tolerance (reference: v1/fault_tolerance/engine_core_sentinel.py:24-80)
    ft_status = "healthy"
    _ft_instruction: Optional[str] = None

    def get_fault_status(self) -> dict:
        return {"engine_index": self.engine_index, "status": self.ft_status}

    def handle_fault(self, instruction: str) -> bool:
        """Operator instruction for a parked engine: ``retry`` (resume serving) or ``abort`` (let it die)."""
        self._ft_instruction = instruction
        return True

    def _recover_from_fault(self, exc: Exception) -> bool:
        """With ``--enable-fault-tolerance``: fail the in-flight requests, park the busy loop (still answering
        utility RPCs) until an operator instruction arrives or the recovery window closes."""
        ft = self.vllm_config.fault_tolerance_config
        if ft is None or not ft.enable_fault_tolerance:
            return False
        logger.exception("engine step failed; parking for up to %.0fs (fault tolerance)", ft.engine_recovery_timeout_sec)
        self.ft_status = "unhealthy"
        from vllm_b200.core.request import EngineCoreOutput, FinishReason
        per_client: dict[int, list] = {}
        for rid, r in list(self.scheduler.requests.items()):
            per_client.setdefault(r.client_index, []).append(EngineCoreOutput(
                request_id=rid, new_token_ids=[], finish_reason=int(FinishReason.ABORT),
                stop_reason=f"engine fault: {type(exc).__name__}"))
        self.scheduler.finish_requests(list(self.scheduler.requests), RequestStatus.FINISHED_ABORTED)
        if self.batch_queue is not None:
            self.batch_queue.clear()
        for ci, outs in per_client.items():
            self.output_queue.put((ci, EngineCoreOutputs(engine_index=self.engine_index, outputs=outs)))
        deadline = time.monotonic() + ft.engine_recovery_timeout_sec
        self._ft_instruction = None
        while time.monotonic() < deadline:
            try:
                kind, payload = self.input_queue.get(timeout=0.2)
            except queue.Empty:
                continue
            if kind == REQ_SHUTDOWN:
                return False
            if kind == REQ_UTILITY:
                self._handle(kind, payload)
            if self._ft_instruction == "retry":
                try:
                    self.model_executor.collective_rpc("reset_after_fault")
                except Exception:
                    logger.exception("worker reset failed")
                    return False
                self.ft_status = "healthy"
                logger.info("fault cleared by operator: resuming")
                return True
            if self._ft_instruction == "abort":
                return False
        return False

    def _handle(self, kind: bytes, payload: Any) -> None:
        if kind == REQ_ADD:
            try:
                self.add_request(payload)
            except Exception as e:  # bad request must not kill the engine
                logger.warning("rejecting request %s: %s", payload.request_id, e)
                from vllm_b200.core.request import EngineCoreOutput, FinishReason
                self.output_queue.put((payload.client_index, EngineCoreOutputs(
                    engine_index=self.engine_index,
                    outputs=[EngineCoreOutput(request_id=payload.request_id, new_token_ids=[],
                                              finish_reason=int(FinishReason.ABORT), stop_reason=str(e))])))
        elif kind == REQ_ABORT:
            self.abort_requests(payload)
        elif kind == REQ_UTILITY:
            client_index, call_id, method, args = payload
            try:
                result = getattr(self, method)(*args)
                uo = (call_id, None, result)
            except Exception as e:
                logger.exception("utility %s failed", method)
                uo = (call_id, f"{type(e).__name__}: {e}", None)
            self.output_queue.put((client_index, EngineCoreOutputs(engine_index=self.engine_index,
                                                                   utility_output=uo)))

    def shutdown(self) -> None:
        self._stop = True
        self.output_queue.put(None)
        super().shutdown()
        try:
            for s in self._in_socks:
                s.close(linger=0)
            for s in self._out_socks:
                s.close(linger=200)
            self._zctx.term()
        except Exception:
            pass

    @staticmethod
    def run_engine_core(vllm_config: VllmConfig, input_addr: str, output_addr: str, engine_index: int = 0,
                        log_stats: bool = False, dp_rank: int = 0, parent_pid: Optional[int] = None) -> None:
        """Process entry point."""
        from vllm_b200.logger import set_process_prefix
        set_process_prefix(f"(EngineCore_{engine_index}) ")
        signal.signal(signal.SIGTERM, lambda *a: (_ for _ in ()).throw(SystemExit()))
        core = None
        try:
            if vllm_config.parallel_config.data_parallel_size > 1:
                from vllm_b200.engine.dp import DPEngineCoreProc
                vllm_config.parallel_config.data_parallel_rank = dp_rank
                core = DPEngineCoreProc(vllm_config, input_addr, output_addr, engine_index, log_stats)
            else:
                core = EngineCoreProc(vllm_config, input_addr, output_addr, engine_index, log_stats)
            core.run_busy_loop()
        except SystemExit:
            pass
        except Exception:
            logger.exception("EngineCore failed")
            if core is not None:
                core.output_queue.put(ENGINE_CORE_DEAD)
                time.sleep(0.2)
            raise
        finally:
            if core is not None:
                core.shutdown()
