"""Concurrent streaming reader for safetensors checkpoints (``--load-format streamer``, alias ``runai_streamer``).

Behavioural reference: vllm/model_executor/model_loader/runai_streamer_loader.py + weight_utils.py
``runai_safetensors_weights_iterator`` (read the tensors of a checkpoint with many concurrent requests into a bounded
CPU buffer while the consumer copies earlier tensors to the device; knobs ``concurrency`` and ``memory_limit`` through
``--model-loader-extra-config``).

Here the reader is a thread pool of positional reads (``os.preadv`` releases the GIL) straight into the tensors'
storage: the safetensors header gives every tensor's byte range, tensors are requested in file order, at most
``memory_limit`` bytes are in flight, and the iterator hands out tensor i while i+1.. are still being read.  On a B200
node the point is to keep several NVMe queues busy; one mmap-ing thread reaches a fraction of the drives' bandwidth.
"""
from __future__ import annotations

import glob
import json
import os
import struct
from collections import deque
from concurrent.futures import ThreadPoolExecutor
from typing import Iterator, Optional

import torch

_DTYPES = {"F64": torch.float64, "F32": torch.float32, "F16": torch.float16, "BF16": torch.bfloat16,
           "I64": torch.int64, "I32": torch.int32, "I16": torch.int16, "I8": torch.int8, "U8": torch.uint8,
           "BOOL": torch.bool, "F8_E4M3": torch.float8_e4m3fn, "F8_E5M2": torch.float8_e5m2,
           "U16": torch.uint16, "U32": torch.uint32, "U64": torch.uint64}
_CHUNK = 16 << 20                         # one positional read; large tensors are split so all threads work on them


def read_header(path: str) -> tuple[dict, int]:
    with open(path, "rb") as f:
        (n,) = struct.unpack("<Q", f.read(8))
        hdr = json.loads(f.read(n))
    hdr.pop("__metadata__", None)
    return hdr, 8 + n


def checkpoint_files(path: str) -> list[str]:
    if os.path.isfile(path):
        return [path]
    idx = os.path.join(path, "model.safetensors.index.json")
    if os.path.exists(idx):
        with open(idx) as f:
            return sorted({os.path.join(path, v) for v in json.load(f)["weight_map"].values()})
    return sorted(glob.glob(os.path.join(path, "*.safetensors")))


def _read_into(fd: int, buf: memoryview, offset: int) -> int:
    done = 0
    while done < len(buf):
        n = os.preadv(fd, [buf[done:]], offset + done)
        if n <= 0:
            raise IOError(f"short read at offset {offset + done}")
        done += n
    return done


def stream_safetensors(path: str, concurrency: int = 16, memory_limit: Optional[int] = None,
                       pin_memory: bool = False, stats: Optional[dict] = None) -> Iterator[tuple[str, torch.Tensor]]:
    """Yields ``(name, cpu tensor)`` in file order.  ``memory_limit`` bounds the bytes requested but not yet consumed
    (default 2 GiB; a single larger tensor is still read on its own)."""
    limit = int(memory_limit) if memory_limit else 2 << 30
    concurrency = max(1, int(concurrency))
    with ThreadPoolExecutor(max_workers=concurrency, thread_name_prefix="wstream") as pool:
        for fn in checkpoint_files(path):
            hdr, base = read_header(fn)
            items = sorted(hdr.items(), key=lambda kv: kv[1]["data_offsets"][0])
            fd = os.open(fn, os.O_RDONLY)
            try:
                pending: deque = deque()                      # (name, tensor, futures, nbytes)
                in_flight = 0
                it = iter(items)
                nxt = next(it, None)
                while nxt is not None or pending:
                    while nxt is not None and (not pending or in_flight + (nxt[1]["data_offsets"][1] - nxt[1]["data_offsets"][0]) <= limit):
                        name, meta = nxt
                        lo, hi = meta["data_offsets"]
                        nbytes = hi - lo
                        raw = torch.empty(nbytes, dtype=torch.uint8, pin_memory=pin_memory and nbytes > 0)
                        mv = memoryview(raw.numpy()).cast("B") if nbytes else memoryview(b"")
                        futs = [pool.submit(_read_into, fd, mv[o:min(o + _CHUNK, nbytes)], base + lo + o)
                                for o in range(0, nbytes, _CHUNK)]
                        dt = _DTYPES[meta["dtype"]]
                        t = raw.view(dt).reshape(meta["shape"]) if nbytes else torch.empty(meta["shape"], dtype=dt)
                        pending.append((name, t, futs, nbytes))
                        in_flight += nbytes
                        nxt = next(it, None)
                    name, t, futs, nbytes = pending.popleft()
                    for f in futs:
                        f.result()
                    in_flight -= nbytes
                    if stats is not None:
                        stats["bytes"] = stats.get("bytes", 0) + nbytes
                        stats["tensors"] = stats.get("tensors", 0) + 1
                        stats["max_in_flight"] = max(stats.get("max_in_flight", 0), in_flight + nbytes)
                    yield name, t
            finally:
                os.close(fd)
