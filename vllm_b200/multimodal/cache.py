"""Sender / receiver halves of the multimodal input cache (reference: vllm/multimodal/cache.py — "P0 shadow" in the
front-end process, "P1 receiver" in the engine-core process): a processed item (pixel patches, log-mel window) crosses
the IPC channel once; repeats travel as their content hash only.

Both halves run the same LRU over the same sequence of hashes (requests reach the engine core in the order the
front-end processed them), so the sender always knows what the receiver still holds without any feedback channel.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Optional

CACHED_MODALITIES = ("image", "audio")


class SenderShadow:
    def __init__(self, capacity: int = 256) -> None:
        self.capacity = capacity
        self.od: "OrderedDict[str, None]" = OrderedDict()
        self.hits = self.misses = 0

    def strip(self, features: Optional[list[dict]]) -> None:
        """Drop the payload of every item the receiver is known to hold."""
        for f in features or ():
            if f.get("modality") not in CACHED_MODALITIES or f.get("data") is None:
                continue
            h = f["hash"]
            if h in self.od:
                self.od.move_to_end(h)
                f["data"] = None
                self.hits += 1
            else:
                self.od[h] = None
                self.misses += 1
                while len(self.od) > self.capacity:
                    self.od.popitem(last=False)


class ReceiverCache:
    def __init__(self, capacity: int = 256) -> None:
        self.capacity = capacity
        self.od: "OrderedDict[str, Any]" = OrderedDict()

    def restore(self, features: Optional[list[dict]]) -> None:
        for f in features or ():
            if f.get("modality") not in CACHED_MODALITIES:
                continue
            h = f["hash"]
            if f.get("data") is None:
                if h not in self.od:
                    raise RuntimeError(f"multimodal item {h[:12]} was sent by reference but is not cached "
                                       "(sender / receiver caches out of step)")
                self.od.move_to_end(h)
                f["data"] = self.od[h]
            else:
                self.od[h] = f["data"]
                self.od.move_to_end(h)
                while len(self.od) > self.capacity:
                    self.od.popitem(last=False)
