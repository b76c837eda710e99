"""KV-cache event stream for external routers (reference: vllm/distributed/kv_events.py).

Events are msgspec structs; ``ZmqEventPublisher`` PUBlishes batches with a
monotonically increasing sequence number, ``NullPublisher`` drops them.
"""
from __future__ import annotations

import threading
import time
from typing import Any, Optional, Union

import msgspec


class BlockStored(msgspec.Struct, tag=True, omit_defaults=True):
    block_hashes: list[bytes]
    parent_block_hash: Optional[bytes]
    token_ids: list[int]
    block_size: int
    lora_id: Optional[int] = None
    medium: str = "GPU"


class BlockRemoved(msgspec.Struct, tag=True, omit_defaults=True):
    block_hashes: list[bytes]
    medium: str = "GPU"


class AllBlocksCleared(msgspec.Struct, tag=True):
    pass


KVCacheEvent = Union[BlockStored, BlockRemoved, AllBlocksCleared]


class KVEventBatch(msgspec.Struct):
    ts: float
    events: list[KVCacheEvent]
    data_parallel_rank: Optional[int] = None


class NullPublisher:
    def publish(self, batch: KVEventBatch) -> None:
        pass

    def shutdown(self) -> None:
        pass


class ZmqEventPublisher:
    def __init__(self, endpoint: str = "tcp://*:5557", topic: str = "", dp_rank: int = 0) -> None:
        import zmq
        self._ctx = zmq.Context.instance()
        self._sock = self._ctx.socket(zmq.PUB)
        if "*" in endpoint or endpoint.startswith("ipc://"):
            self._sock.bind(endpoint)
        else:
            self._sock.connect(endpoint)
        self._topic = topic.encode()
        self._seq = 0
        self._enc = msgspec.msgpack.Encoder()
        self._lock = threading.Lock()
        self._dp_rank = dp_rank

    def publish(self, batch: KVEventBatch) -> None:
        batch.data_parallel_rank = self._dp_rank
        with self._lock:
            payload = self._enc.encode(batch)
            self._sock.send_multipart([self._topic, self._seq.to_bytes(8, "big"), payload])
            self._seq += 1

    def shutdown(self) -> None:
        self._sock.close(linger=0)


def make_publisher(cfg: Any, dp_rank: int = 0):
    if cfg is None or not cfg.enable_kv_cache_events or cfg.publisher == "null":
        return NullPublisher()
    if cfg.publisher == "zmq":
        return ZmqEventPublisher(cfg.endpoint, cfg.topic, dp_rank)
    raise ValueError(f"unknown KV event publisher {cfg.publisher}")


def make_batch(events: list[KVCacheEvent]) -> KVEventBatch:
    return KVEventBatch(ts=time.time(), events=events)
