"""CPU-memory KV tier (reference: vllm/v1/kv_offload/* — CPU offload with LRU/ARC eviction — and
simple_kv_offload/).

``CpuBlockStore`` keeps evicted / finished prompt blocks in (pinned, when CUDA is present) host memory under
an LRU or ARC policy; ``OffloadingConnector`` plugs it into the connector protocol, so a request whose prefix
fell out of the HBM prefix cache is restored with a host→device copy instead of a recompute.  With 180 GB of
HBM the GPU tier is already large; the host tier matters for very long shared prefixes and many tenants.
An optional spill directory turns it into a two-level store (host → NVMe through the native IO).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional

import torch

from vllm_b200.kv_transfer import BlockConnector, BlockStore, FileBlockStore


class LRUPolicy:
    def __init__(self) -> None:
        self.od: "OrderedDict[str, None]" = OrderedDict()

    def touch(self, k: str) -> None:
        self.od.pop(k, None)
        self.od[k] = None

    def insert(self, k: str) -> None:
        self.touch(k)

    def evict(self) -> Optional[str]:
        return self.od.popitem(last=False)[0] if self.od else None

    def remove(self, k: str) -> None:
        self.od.pop(k, None)


class ARCPolicy:
    """Adaptive replacement cache (recency list T1, frequency list T2, ghost lists B1/B2, target p)."""

    def __init__(self, capacity: int) -> None:
        self.c = max(capacity, 1)
        self.p = 0
        self.t1: "OrderedDict[str, None]" = OrderedDict()
        self.t2: "OrderedDict[str, None]" = OrderedDict()
        self.b1: "OrderedDict[str, None]" = OrderedDict()
        self.b2: "OrderedDict[str, None]" = OrderedDict()

    def touch(self, k: str) -> None:
        if k in self.t1:
            del self.t1[k]
            self.t2[k] = None
        elif k in self.t2:
            self.t2.move_to_end(k)

    def insert(self, k: str) -> None:
        if k in self.t1 or k in self.t2:
            return self.touch(k)
        if k in self.b1:
            self.p = min(self.c, self.p + max(len(self.b2) // max(len(self.b1), 1), 1))
            del self.b1[k]
            self.t2[k] = None
        elif k in self.b2:
            self.p = max(0, self.p - max(len(self.b1) // max(len(self.b2), 1), 1))
            del self.b2[k]
            self.t2[k] = None
        else:
            self.t1[k] = None
        for ghost in (self.b1, self.b2):
            while len(ghost) > self.c:
                ghost.popitem(last=False)

    def evict(self) -> Optional[str]:
        if self.t1 and (len(self.t1) > self.p or not self.t2):
            k = self.t1.popitem(last=False)[0]
            self.b1[k] = None
            return k
        if self.t2:
            k = self.t2.popitem(last=False)[0]
            self.b2[k] = None
            return k
        return None

    def remove(self, k: str) -> None:
        self.t1.pop(k, None)
        self.t2.pop(k, None)


class CpuBlockStore(BlockStore):
    def __init__(self, capacity_bytes: int, policy: str = "lru", spill_dir: Optional[str] = None) -> None:
        self.capacity = capacity_bytes
        self.used = 0
        self.data: dict[str, torch.Tensor] = {}
        self._policy_name = policy
        self.policy = None
        self.spill = FileBlockStore(spill_dir) if spill_dir else None
        self.pin = torch.cuda.is_available()
        self.hits = self.misses = self.evictions = 0

    def _ensure_policy(self, block_bytes: int) -> None:
        if self.policy is None:
            n = max(self.capacity // max(block_bytes, 1), 1)
            self.policy = ARCPolicy(n) if self._policy_name == "arc" else LRUPolicy()

    def contains(self, key: str) -> bool:
        return key in self.data or (self.spill is not None and self.spill.contains(key))

    def put(self, key: str, t: torch.Tensor) -> None:
        if key in self.data:
            self.policy.touch(key)
            return
        raw = t.contiguous().view(torch.uint8).reshape(-1)
        self._ensure_policy(raw.numel())
        while self.used + raw.numel() > self.capacity and self.data:
            victim = self.policy.evict()
            if victim is None:
                break
            v = self.data.pop(victim, None)
            if v is not None:
                self.used -= v.numel()
                self.evictions += 1
                if self.spill is not None:
                    self.spill.put(victim, v)
        if raw.numel() > self.capacity:
            return
        buf = torch.empty(raw.numel(), dtype=torch.uint8, pin_memory=self.pin)
        buf.copy_(raw)
        self.data[key] = buf
        self.used += raw.numel()
        self.policy.insert(key)

    def get(self, key: str, out: torch.Tensor) -> bool:
        v = self.data.get(key)
        if v is not None:
            out.view(torch.uint8).reshape(-1).copy_(v)
            self.policy.touch(key)
            self.hits += 1
            return True
        if self.spill is not None and self.spill.get(key, out):
            self.hits += 1
            return True
        self.misses += 1
        return False


class OffloadingConnector(BlockConnector):
    """In-process host tier: the scheduler- and worker-side objects of one engine share the store."""

    _stores: dict[int, CpuBlockStore] = {}

    def __init__(self, vllm_config, role: str) -> None:
        extra = vllm_config.kv_transfer_config.kv_connector_extra_config
        key = id(vllm_config.kv_transfer_config)
        st = self._stores.get(key)
        if st is None:
            st = CpuBlockStore(int(extra.get("cpu_bytes_to_use", 4 << 30)), extra.get("eviction_policy", "lru"),
                               extra.get("spill_dir"))
            self._stores[key] = st
        super().__init__(vllm_config, role, st)
