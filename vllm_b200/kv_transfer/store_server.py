"""Networked KV block store: a stand-alone key → block server (LRU-bounded host memory) and the client-side
``BlockStore`` that lets several engine instances on different machines share prefix KV without a shared file system.

Role of the reference's external-store connectors (LMCache / Mooncake-store back-ends of
vllm/distributed/kv_transfer/kv_connector/v1/): same connector surface (``RemoteStoreConnector`` is the ordinary block
connector over this store), transport = ZeroMQ REQ/REP with raw block payloads.

    python -m vllm_b200.kv_transfer.store_server --port 15000 --capacity-gb 64
    --kv-transfer-config '{"kv_connector": "RemoteStoreConnector", "kv_role": "kv_both",
                           "kv_connector_extra_config": {"store_url": "tcp://host:15000"}}'
"""
from __future__ import annotations

import argparse
import threading
from collections import OrderedDict
from typing import Optional

import msgspec
import torch

from vllm_b200.kv_transfer import BlockConnector, BlockStore


class StoreServer:
    def __init__(self, url: str = "tcp://127.0.0.1:15000", capacity_bytes: int = 8 << 30) -> None:
        import zmq
        self.ctx = zmq.Context.instance()
        self.sock = self.ctx.socket(zmq.REP)
        self.sock.bind(url)
        self.url = self.sock.getsockopt_string(zmq.LAST_ENDPOINT)
        self.capacity, self.used = capacity_bytes, 0
        self.od: "OrderedDict[str, bytes]" = OrderedDict()
        self.stats = {"puts": 0, "hits": 0, "misses": 0, "evictions": 0}
        self._stop = False
        self.thread: Optional[threading.Thread] = None

    def serve_forever(self) -> None:
        import zmq
        self.sock.setsockopt(zmq.RCVTIMEO, 200)
        while not self._stop:
            try:
                frames = self.sock.recv_multipart()
            except zmq.Again:
                continue
            op = frames[0]
            if op == b"PUT":
                key, val = frames[1].decode(), frames[2]
                if key not in self.od:
                    self.od[key] = val
                    self.used += len(val)
                    self.stats["puts"] += 1
                    while self.used > self.capacity and len(self.od) > 1:
                        _, old = self.od.popitem(last=False)
                        self.used -= len(old)
                        self.stats["evictions"] += 1
                self.sock.send(b"OK")
            elif op == b"GET":
                v = self.od.get(frames[1].decode())
                if v is None:
                    self.stats["misses"] += 1
                    self.sock.send_multipart([b"MISS", b""])
                else:
                    self.od.move_to_end(frames[1].decode())
                    self.stats["hits"] += 1
                    self.sock.send_multipart([b"HIT", v])
            elif op == b"HAS":
                keys = msgspec.msgpack.decode(frames[1])
                self.sock.send(msgspec.msgpack.encode([k in self.od for k in keys]))
            elif op == b"STATS":
                self.sock.send(msgspec.msgpack.encode(dict(self.stats, items=len(self.od), used=self.used)))
            else:
                self.sock.send(b"ERR")
        self.sock.close(linger=0)

    def start(self) -> "StoreServer":
        self.thread = threading.Thread(target=self.serve_forever, daemon=True, name="kv-store-server")
        self.thread.start()
        return self

    def stop(self) -> None:
        self._stop = True
        if self.thread is not None:
            self.thread.join(timeout=2)


class RemoteBlockStore(BlockStore):
    def __init__(self, url: str) -> None:
        import zmq
        self.url = url
        self.ctx = zmq.Context.instance()
        self._local = threading.local()
        self._zmq = zmq

    def _sock(self):
        s = getattr(self._local, "sock", None)
        if s is None:
            s = self.ctx.socket(self._zmq.REQ)
            s.setsockopt(self._zmq.RCVTIMEO, 30000)
            s.setsockopt(self._zmq.LINGER, 0)
            s.connect(self.url)
            self._local.sock = s
        return s

    def contains(self, key: str) -> bool:
        s = self._sock()
        s.send_multipart([b"HAS", msgspec.msgpack.encode([key])])
        return bool(msgspec.msgpack.decode(s.recv())[0])

    def put(self, key: str, t: torch.Tensor) -> None:
        s = self._sock()
        s.send_multipart([b"PUT", key.encode(), t.contiguous().cpu().view(torch.uint8).reshape(-1).numpy().tobytes()])
        s.recv()

    def get(self, key: str, out: torch.Tensor) -> bool:
        s = self._sock()
        s.send_multipart([b"GET", key.encode()])
        tag, val = s.recv_multipart()
        if tag != b"HIT":
            return False
        raw = out.view(torch.uint8).reshape(-1)
        if len(val) != raw.numel():
            return False
        raw.copy_(torch.frombuffer(bytearray(val), dtype=torch.uint8))
        return True

    def stats(self) -> dict:
        s = self._sock()
        s.send_multipart([b"STATS"])
        return msgspec.msgpack.decode(s.recv())


class RemoteStoreConnector(BlockConnector):
    def __init__(self, vllm_config, role: str) -> None:
        extra = vllm_config.kv_transfer_config.kv_connector_extra_config
        super().__init__(vllm_config, role, RemoteBlockStore(extra.get("store_url", "tcp://127.0.0.1:15000")))


def main(argv=None) -> None:
    ap = argparse.ArgumentParser("vllm_b200 KV store server")
    ap.add_argument("--host", default="0.0.0.0")
    ap.add_argument("--port", type=int, default=15000)
    ap.add_argument("--capacity-gb", type=float, default=8.0)
    a = ap.parse_args(argv)
    StoreServer(f"tcp://{a.host}:{a.port}", int(a.capacity_gb * (1 << 30))).serve_forever()


if __name__ == "__main__":
    main()
