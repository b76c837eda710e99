"""Point-to-point KV pull connector: disaggregated prefill → decode over the network.

Behavioural reference: vllm/distributed/kv_transfer/kv_connector/v1/nixl_connector.py (prefill instance keeps the
finished request's blocks alive and returns ``kv_transfer_params``; the decode instance allocates blocks, pulls the
remote blocks asynchronously while the request sits in WAITING_FOR_REMOTE_KVS, then notifies the producer so it can
free them).  Transport here is ZeroMQ over TCP between the worker processes (one REP server per producer worker);
the protocol is transport-agnostic, an NVLink / GPUDirect path only has to replace ``_serve`` / ``_pull``.

Flow:
  prefill instance (``kv_role=kv_producer``):  request with ``kv_transfer_params={"do_remote_decode": true}``,
      max_tokens=1  →  output carries ``kv_transfer_params={do_remote_prefill, remote_block_ids, remote_host,
      remote_port, remote_req_id}``; blocks stay allocated until the consumer says "done" (or a timeout).
  decode instance (``kv_role=kv_consumer``):  same prompt + those params  →  scheduler allocates blocks, the worker
      pulls in the background, the request resumes with the prompt already computed.
"""
from __future__ import annotations

import threading
import time
from typing import Any, Optional

import msgspec
import torch

from vllm_b200.kv_transfer import KVConnectorBase, KVConnectorOutput
from vllm_b200.logger import init_logger

logger = init_logger(__name__)


class P2pPullConnector(KVConnectorBase):
    def __init__(self, vllm_config, role: str) -> None:
        super().__init__(vllm_config, role)
        c = self.cfg
        self.producer = c.kv_role in ("kv_producer", "kv_both")
        self.consumer = c.kv_role in ("kv_consumer", "kv_both")
        self.host = c.kv_ip
        self.port = c.kv_port + c.kv_rank
        self.send_timeout = float(c.kv_connector_extra_config.get("send_timeout_s", 120.0))
        self.pull_timeout_ms = int(c.kv_connector_extra_config.get("pull_timeout_ms", 60000))
        # scheduler side
        self._pending: dict[str, tuple[int, int]] = {}
        self._jobs: list[dict] = []
        self._sending: dict[str, float] = {}
        # worker side
        self._kv: Optional[dict[str, torch.Tensor]] = None
        self._lock = threading.Lock()
        self._fin_send: set[str] = set()
        self._fin_recv: set[str] = set()
        self._fail_recv: set[str] = set()
        self._server: Optional[threading.Thread] = None
        self._stop = False
        self.num_pulled_blocks = 0
        self.num_served_blocks = 0

    # ------------------------------------------------------------------ scheduler side
    def get_num_new_matched_tokens(self, request, num_computed_tokens: int) -> tuple[int, bool]:
        p = request.kv_transfer_params or {}
        if not self.consumer or not p.get("do_remote_prefill"):
            return 0, False
        bs = self.block_size
        n_tok = min(len(p["remote_block_ids"]) * bs, request.num_tokens - 1) // bs * bs
        n_ext = max(n_tok - num_computed_tokens, 0)
        if n_ext == 0:
            p["do_remote_prefill"] = False
            return 0, False
        self._pending[request.request_id] = (num_computed_tokens // bs, n_tok // bs)
        return n_ext, True

    def update_state_after_alloc(self, request, blocks, num_external_tokens: int) -> None:
        rng = self._pending.pop(request.request_id, None)
        if rng is None or num_external_tokens == 0:
            return
        p = request.kv_transfer_params
        first, last = rng
        local = blocks.get_block_ids()[0]
        self._jobs.append({"req_id": request.request_id, "host": p["remote_host"], "port": p["remote_port"],
                           "remote_req_id": p["remote_req_id"], "remote_ids": list(p["remote_block_ids"][first:last]),
                           "local_ids": list(local[first:last])})
        p["do_remote_prefill"] = False           # consumed: the request is an ordinary one from here on

    def build_connector_meta(self, so) -> Any:
        now = time.monotonic()
        expired = [r for r, t in self._sending.items() if now - t > self.send_timeout]
        for r in expired:
            self._sending.pop(r, None)
        if not self._jobs and not expired:
            return None
        meta = {"pulls": self._jobs, "expire": expired}
        self._jobs = []
        return meta

    def request_finished(self, request, block_ids) -> tuple[bool, Optional[dict]]:
        p = request.kv_transfer_params or {}
        from vllm_b200.core.request import RequestStatus
        if not self.producer or not p.get("do_remote_decode") or request.status == RequestStatus.FINISHED_ABORTED:
            return False, None
        n_full = min(request.num_computed_tokens, request.num_tokens) // self.block_size
        if n_full == 0:
            return False, None
        self._sending[request.request_id] = time.monotonic()
        return True, {"do_remote_prefill": True, "do_remote_decode": False, "remote_block_ids": list(block_ids[0][:n_full]),
                      "remote_host": self.host, "remote_port": self.port, "remote_req_id": request.request_id}

    def mark_send_finished(self, request_id: str) -> None:
        self._sending.pop(request_id, None)

    # ------------------------------------------------------------------ worker side
    def register_kv_caches(self, kv_caches: dict[str, torch.Tensor]) -> None:
        self._kv = kv_caches
        if self.producer and self._server is None:
            self._server = threading.Thread(target=self._serve, daemon=True, name="kv-p2p-server")
            self._server.start()

    def _pack(self, block_ids: list[int]) -> bytes:
        idx = torch.tensor(block_ids, dtype=torch.long)
        parts = []
        for name in sorted(self._kv):
            t = self._kv[name]
            parts.append(t[idx.to(t.device)].contiguous().cpu().view(torch.uint8).reshape(-1))
        return torch.cat(parts).numpy().tobytes()

    def _serve(self) -> None:
        import zmq
        ctx = zmq.Context.instance()
        sock = ctx.socket(zmq.REP)
        sock.bind(f"tcp://{self.host}:{self.port}")
        sock.setsockopt(zmq.RCVTIMEO, 500)
        while not self._stop:
            try:
                msg = sock.recv()
            except zmq.Again:
                continue
            try:
                req = msgspec.msgpack.decode(msg)
                if req["op"] == "pull":
                    sock.send(self._pack(req["block_ids"]))
                    self.num_served_blocks += len(req["block_ids"])
                elif req["op"] == "done":
                    with self._lock:
                        self._fin_send.add(req["req_id"])
                    sock.send(b"ok")
                else:
                    sock.send(b"?")
            except Exception:  # pragma: no cover
                logger.exception("kv p2p server error")
                sock.send(b"")
        sock.close(linger=0)

    def _pull(self, job: dict) -> None:
        import zmq
        try:
            ctx = zmq.Context.instance()
            s = ctx.socket(zmq.REQ)
            s.setsockopt(zmq.RCVTIMEO, self.pull_timeout_ms)
            s.setsockopt(zmq.LINGER, 0)
            s.connect(f"tcp://{job['host']}:{job['port']}")
            s.send(msgspec.msgpack.encode({"op": "pull", "block_ids": job["remote_ids"]}))
            raw = torch.frombuffer(bytearray(s.recv()), dtype=torch.uint8)
            off = 0
            idx = torch.tensor(job["local_ids"], dtype=torch.long)
            for name in sorted(self._kv):
                t = self._kv[name]
                blk = t[0:1].numel() * t.element_size()
                n = blk * len(job["local_ids"])
                chunk = raw[off:off + n].view(t.dtype).view((len(job["local_ids"]),) + tuple(t.shape[1:]))
                t[idx.to(t.device)] = chunk.to(t.device)
                off += n
            s.send(msgspec.msgpack.encode({"op": "done", "req_id": job["remote_req_id"]}))
            s.recv()
            s.close(linger=0)
            self.num_pulled_blocks += len(job["local_ids"])
        except Exception:
            logger.exception("KV pull for %s failed", job["req_id"])
            with self._lock:
                self._fail_recv.add(job["req_id"])      # the blocks hold garbage: the scheduler recomputes, never caches them
        finally:
            with self._lock:
                self._fin_recv.add(job["req_id"])

    def start_load_kv(self, meta, kv_caches) -> None:
        if self._kv is None:
            self.register_kv_caches(kv_caches)
        if not meta:
            return
        for job in meta.get("pulls", ()):
            threading.Thread(target=self._pull, args=(job,), daemon=True, name="kv-p2p-pull").start()
        if meta.get("expire"):
            with self._lock:
                self._fin_send.update(meta["expire"])

    def get_finished(self) -> Optional[KVConnectorOutput]:
        with self._lock:
            if not self._fin_send and not self._fin_recv:
                return None
            out = KVConnectorOutput(set(self._fin_send), set(self._fin_recv), set(self._fail_recv))
            self._fin_send.clear()
            self._fin_recv.clear()
            self._fail_recv.clear()
        return out

    def shutdown(self) -> None:
        self._stop = True
