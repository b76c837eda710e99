"""Shared-memory message queues for the executor control plane.

``ShmRing``: lock-free single-writer / N-reader ring in POSIX shared memory.  The engine-core process
broadcasts each step's ``SchedulerOutput`` to all workers of the node through it; a small ring in the
other direction carries the reply of the output rank.  Waiting is done by the native spin-wait
(``native/spinloop.cpp``: pause → yield → nanosleep, GIL released) so an idle worker costs nothing and a
busy one reacts in well under a microsecond of the store.

Role of the reference's distributed/device_communicators/shm_broadcast.py:250-352 (ring + flags) and
:112 ``SpinCondition``; the protocol here is sequence-number based: slot i of the ring carries message
number ``seq`` where ``seq % n_slots == i``; the writer publishes ``written = seq+1`` after filling the slot,
each reader r publishes ``read[r] = seq+1`` after consuming it, and the writer may reuse a slot once every
reader's counter has passed ``seq - n_slots + 1``.  Messages larger than a slot are chunked.
"""
from __future__ import annotations

import ctypes
import pickle
import struct
import time
from multiprocessing import shared_memory
from typing import Any, Optional

from vllm_b200.native import lib as native_lib

_HDR = 64            # bytes reserved per counter (own cache line)


class ShmRing:
    def __init__(self, n_readers: int, slot_bytes: int = 1 << 20, n_slots: int = 8, name: Optional[str] = None,
                 create: bool = True) -> None:
        self.n_readers, self.slot_bytes, self.n_slots = n_readers, slot_bytes, n_slots
        self.ctrl_bytes = _HDR * (1 + n_readers)
        total = self.ctrl_bytes + n_slots * (8 + slot_bytes)
        if create:
            self.shm = shared_memory.SharedMemory(create=True, size=total, name=name)
            self.shm.buf[:self.ctrl_bytes] = bytes(self.ctrl_bytes)
        else:
            self.shm = shared_memory.SharedMemory(name=name)
        self.created = create
        self.name = self.shm.name
        self._base = ctypes.addressof(ctypes.c_char.from_buffer(self.shm.buf))
        self._lib = native_lib()
        self._seq = 0            # next message to write (writer) / read (reader)
        self.reader_rank = -1

    def handle(self) -> dict:
        return dict(name=self.name, n_readers=self.n_readers, slot_bytes=self.slot_bytes, n_slots=self.n_slots)

    @classmethod
    def attach(cls, handle: dict, reader_rank: int = -1) -> "ShmRing":
        r = cls(handle["n_readers"], handle["slot_bytes"], handle["n_slots"], handle["name"], create=False)
        r.reader_rank = reader_rank
        return r

    # ---------------------------------------------------------------- raw counters
    def _addr_written(self) -> int:
        return self._base

    def _addr_read(self, r: int) -> int:
        return self._base + _HDR * (1 + r)

    def _load(self, addr: int) -> int:
        if self._lib is not None:
            return self._lib.vbn_load_acquire_u64(addr)
        return struct.unpack_from("<Q", self.shm.buf, addr - self._base)[0]

    def _store(self, addr: int, v: int) -> None:
        if self._lib is not None:
            self._lib.vbn_store_release_u64(addr, v)
        else:
            struct.pack_into("<Q", self.shm.buf, addr - self._base, v)

    def _wait_ge(self, addr: int, target: int, timeout: Optional[float]) -> bool:
        us = -1 if timeout is None else int(timeout * 1e6)
        if self._lib is not None:
            return bool(self._lib.vbn_wait_ge_u64(addr, target, us))
        t0 = time.monotonic()
        while self._load(addr) < target:
            if timeout is not None and time.monotonic() - t0 > timeout:
                return False
            time.sleep(0.0001)
        return True

    def _slot(self, seq: int) -> int:
        return self.ctrl_bytes + (seq % self.n_slots) * (8 + self.slot_bytes)

    # ---------------------------------------------------------------- writer
    def _put_chunk(self, payload: memoryview, last: bool, timeout: Optional[float]) -> None:
        seq = self._seq
        need = seq - self.n_slots + 1           # every reader must have consumed message `need - 1`
        if need > 0:
            for r in range(self.n_readers):
                if not self._wait_ge(self._addr_read(r), need, timeout):
                    raise TimeoutError(f"shm ring: reader {r} did not release a slot")
        off = self._slot(seq)
        n = len(payload)
        struct.pack_into("<II", self.shm.buf, off, n, 1 if last else 0)
        self.shm.buf[off + 8: off + 8 + n] = payload
        self._store(self._addr_written(), seq + 1)
        self._seq = seq + 1

    def put_bytes(self, data: bytes, timeout: Optional[float] = None) -> None:
        mv = memoryview(data)
        if len(mv) == 0:
            self._put_chunk(mv, True, timeout)
            return
        for i in range(0, len(mv), self.slot_bytes):
            self._put_chunk(mv[i:i + self.slot_bytes], i + self.slot_bytes >= len(mv), timeout)

    def put(self, obj: Any, timeout: Optional[float] = None) -> None:
        self.put_bytes(pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL), timeout)

    # ---------------------------------------------------------------- reader
    def get_bytes(self, timeout: Optional[float] = None) -> bytes:
        assert self.reader_rank >= 0, "attach() with a reader rank first"
        parts = []
        while True:
            seq = self._seq
            if not self._wait_ge(self._addr_written(), seq + 1, timeout):
                raise TimeoutError("shm ring: no message")
            off = self._slot(seq)
            n, last = struct.unpack_from("<II", self.shm.buf, off)
            parts.append(bytes(self.shm.buf[off + 8: off + 8 + n]))
            self._store(self._addr_read(self.reader_rank), seq + 1)
            self._seq = seq + 1
            if last:
                break
        return parts[0] if len(parts) == 1 else b"".join(parts)

    def get(self, timeout: Optional[float] = None) -> Any:
        return pickle.loads(self.get_bytes(timeout))

    def close(self) -> None:
        try:
            self.shm.close()
            if self.created:
                self.shm.unlink()
        except Exception:  # pragma: no cover
            pass
