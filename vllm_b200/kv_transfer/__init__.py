"""KV-cache connectors: move computed KV blocks between engine instances / storage tiers.

Behavioural reference: vllm/distributed/kv_transfer/kv_connector/v1/base.py (scheduler-side
``get_num_new_matched_tokens`` / ``update_state_after_alloc`` / ``build_connector_meta`` /
``request_finished``; worker-side ``start_load_kv`` / ``save_kv_layer`` / ``get_finished``) and factory.py
:152-228 (connector registry).

Connectors shipped here:

* ``SharedStorageConnector`` (alias ``ExampleConnector``): prefill instances write every full prompt block to
  a shared directory keyed by the prefix-cache hash chain; decode instances (or a restarted engine) find the
  longest stored prefix and load it instead of recomputing — disaggregated prefill / warm restart through a
  filesystem (NVMe, tmpfs, parallel FS).  Block files are moved with the GIL-free native IO.
* ``OffloadingConnector`` (see ``vllm_b200.kv_offload``): same protocol against a CPU-memory tier.
* ``MultiConnector``: fan-out over several connectors (first hit wins on load, all are saved to).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Any, Optional

import torch

from vllm_b200.logger import init_logger

logger = init_logger(__name__)


@dataclass
class KVConnectorMetadata:
    """Per-step instructions for the workers. ``loads`` / ``saves``: (block_id, key) pairs."""
    loads: list[tuple[int, str]] = field(default_factory=list)
    saves: list[tuple[int, str]] = field(default_factory=list)

    def __bool__(self) -> bool:
        return bool(self.loads or self.saves)


@dataclass
class KVConnectorOutput:
    finished_sending: set[str] = field(default_factory=set)
    finished_recving: set[str] = field(default_factory=set)
    failed_recving: set[str] = field(default_factory=set)     # transfers that ended without valid KV (⊆ finished_recving)


class KVConnectorBase:
    def __init__(self, vllm_config, role: str) -> None:
        self.vllm_config = vllm_config
        self.role = role
        self.cfg = vllm_config.kv_transfer_config
        self.block_size = vllm_config.cache_config.block_size
        self._recv_done: set[str] = set()
        self._recv_failed: set[str] = set()

    # ---- scheduler side
    def get_num_new_matched_tokens(self, request, num_computed_tokens: int) -> tuple[int, bool]:
        return 0, False

    def update_state_after_alloc(self, request, blocks, num_external_tokens: int) -> None:
        pass

    def build_connector_meta(self, scheduler_output) -> Any:
        return None

    def request_finished(self, request, block_ids) -> tuple[bool, Optional[dict]]:
        return False, None

    def is_recv_finished(self, request) -> bool:
        return request.request_id in self._recv_done

    def mark_recv_finished(self, request_id: str) -> None:
        self._recv_done.add(request_id)

    def mark_recv_failed(self, request_id: str) -> None:
        """The transfer ended but the destination blocks do not hold valid KV: the scheduler must recompute."""
        self._recv_failed.add(request_id)

    def take_recv_failed(self, request) -> bool:
        if request.request_id in self._recv_failed:
            self._recv_failed.discard(request.request_id)
            return True
        return False

    # ---- worker side
    def start_load_kv(self, meta: Any, kv_caches: dict[str, torch.Tensor]) -> None:
        pass

    def save_kv(self, meta: Any, kv_caches: dict[str, torch.Tensor]) -> None:
        pass

    def get_finished(self) -> Optional[KVConnectorOutput]:
        return None

    def shutdown(self) -> None:
        pass


class BlockStore:
    """Key → bytes store interface used by the block-granular connectors."""

    def contains(self, key: str) -> bool:
        raise NotImplementedError

    def put(self, key: str, t: torch.Tensor) -> None:
        raise NotImplementedError

    def get(self, key: str, out: torch.Tensor) -> bool:
        raise NotImplementedError


class FileBlockStore(BlockStore):
    def __init__(self, root: str) -> None:
        from vllm_b200.native import lib
        self.root = root
        os.makedirs(root, exist_ok=True)
        self.lib = lib()

    def _path(self, key: str) -> str:
        return os.path.join(self.root, key[:2], key + ".kv")

    def contains(self, key: str) -> bool:
        return os.path.exists(self._path(key))

    def put(self, key: str, t: torch.Tensor) -> None:
        p = self._path(key)
        if os.path.exists(p):
            return
        os.makedirs(os.path.dirname(p), exist_ok=True)
        raw = t.contiguous().view(torch.uint8).reshape(-1)
        if self.lib is not None:
            n = self.lib.vbn_write_file(p.encode(), raw.data_ptr(), raw.numel(), 0, 0)
            if n != raw.numel():
                raise OSError(f"short write to {p}: {n}")
        else:  # pragma: no cover
            with open(p + ".tmp", "wb") as f:
                f.write(raw.numpy().tobytes())
            os.replace(p + ".tmp", p)

    def get(self, key: str, out: torch.Tensor) -> bool:
        p = self._path(key)
        raw = out.view(torch.uint8).reshape(-1)
        if self.lib is not None:
            return self.lib.vbn_read_file(p.encode(), raw.data_ptr(), raw.numel(), 0, 0) == raw.numel()
        try:  # pragma: no cover
            with open(p, "rb") as f:
                data = f.read()
            raw.copy_(torch.frombuffer(bytearray(data), dtype=torch.uint8))
            return True
        except OSError:
            return False


class BlockConnector(KVConnectorBase):
    """Block-granular connector over a ``BlockStore``; keys are the prefix-cache block hashes, so any engine with
    the same model / block size / hash seed can reuse what another one stored."""

    def __init__(self, vllm_config, role: str, store: BlockStore) -> None:
        super().__init__(vllm_config, role)
        self.store = store
        self.producer = self.cfg.kv_role in (None, "kv_producer", "kv_both")
        self.consumer = self.cfg.kv_role in (None, "kv_consumer", "kv_both")
        # scheduler-side request tracking
        self._reqs: dict[str, Any] = {}
        self._blocks: dict[str, list[int]] = {}
        self._saved: dict[str, int] = {}
        self._pending_loads: dict[str, tuple[int, int]] = {}   # req → (first block index, n blocks)
        self.num_loaded_blocks = 0
        self.num_saved_blocks = 0
        # Keys carry the identity of WHAT is stored, not only the token prefix: a namespace hashed from the model,
        # dtypes, block size and parallel layout, plus the shard coordinates of the worker that owns the bytes —
        # two models sharing a store, or the TP ranks of one engine, can never read each other's blocks.
        import hashlib
        import json
        mc, cc, pc = vllm_config.model_config, vllm_config.cache_config, vllm_config.parallel_config
        ident = {"model": getattr(mc, "model", None), "dtype": str(getattr(mc, "torch_dtype", getattr(mc, "dtype", None))),
                 "kv_dtype": str(getattr(cc, "cache_dtype", None)), "block_size": self.block_size,
                 "tp": getattr(pc, "tensor_parallel_size", 1), "pp": getattr(pc, "pipeline_parallel_size", 1),
                 "dcp": getattr(pc, "decode_context_parallel_size", 1),
                 "hf": getattr(mc, "hf_config", None)}
        self._ns = hashlib.sha256(json.dumps(ident, sort_keys=True, default=str).encode()).hexdigest()[:12]
        self._tp, self._pp = ident["tp"] or 1, ident["pp"] or 1
        self._shard: Optional[str] = None      # resolved lazily on the worker (parallel groups exist by then)

    def _key(self, h) -> str:
        hx = h.hex() if isinstance(h, (bytes, bytearray)) else format(int(h) & ((1 << 128) - 1), "032x")
        return f"{self._ns}-{hx}"

    def _shards(self) -> list[str]:
        return [f".t{t}p{p}" for p in range(self._pp) for t in range(self._tp)]

    def _my_shard(self) -> str:
        if self._shard is None:
            t = p = 0
            try:
                from vllm_b200.parallel.state import get_pp_group, get_tp_group
                t, p = max(get_tp_group().rank_in_group, 0), max(get_pp_group().rank_in_group, 0)
            except Exception:
                pass
            self._shard = f".t{t}p{p}"
        return self._shard

    def _stored(self, key: str) -> bool:
        """Scheduler side: a block counts as stored only when every shard of it is."""
        return all(self.store.contains(key + sh) for sh in self._shards())

    # ---- scheduler side
    def get_num_new_matched_tokens(self, request, num_computed_tokens: int) -> tuple[int, bool]:
        if not self.consumer:
            return 0, False
        bs = self.block_size
        first = num_computed_tokens // bs
        # never cover the whole prompt: the last token must be computed to produce logits
        limit = min(len(request.block_hashes), (request.num_tokens - 1) // bs)
        n = first
        while n < limit and self._stored(self._key(request.block_hashes[n])):
            n += 1
        if n > first:
            self._pending_loads[request.request_id] = (first, n - first)
        return (n - first) * bs, False

    def update_state_after_alloc(self, request, blocks, num_external_tokens: int) -> None:
        self._reqs[request.request_id] = request
        if num_external_tokens == 0:
            self._pending_loads.pop(request.request_id, None)

    def build_connector_meta(self, so) -> Any:
        meta = KVConnectorMetadata()
        bs = self.block_size
        for nr in so.scheduled_new_reqs:
            self._blocks[nr.req_id] = list(nr.block_ids[0])
            pl = self._pending_loads.pop(nr.req_id, None)
            req = self._reqs.get(nr.req_id)
            if pl is not None and req is not None:
                first, n = pl
                for i in range(first, first + n):
                    meta.loads.append((self._blocks[nr.req_id][i], self._key(req.block_hashes[i])))
                self._saved[nr.req_id] = first + n       # those already exist in the store
        c = so.scheduled_cached_reqs
        for i, rid in enumerate(c.req_ids):
            nb = c.new_block_ids[i]
            if nb is not None:
                if c.resumed_from_preemption[i]:
                    self._blocks[rid] = list(nb[0])
                else:
                    self._blocks.setdefault(rid, []).extend(nb[0])
        if self.producer:
            for rid, n_sched in so.num_scheduled_tokens.items():
                req = self._reqs.get(rid)
                if req is None:
                    continue
                done = req.num_computed_tokens + n_sched          # tokens whose KV exists after this step
                full = min(done // bs, len(req.block_hashes), len(self._blocks.get(rid, ())))
                for i in range(self._saved.get(rid, 0), full):
                    meta.saves.append((self._blocks[rid][i], self._key(req.block_hashes[i])))
                self._saved[rid] = max(self._saved.get(rid, 0), full)
        for rid in so.finished_req_ids:
            for d in (self._reqs, self._blocks, self._saved, self._pending_loads):
                d.pop(rid, None)
        return meta if meta else None

    # ---- worker side
    def _stack_shape(self, kv_caches):
        names = sorted(kv_caches)
        return names, [kv_caches[n].shape[1:] for n in names]

    def start_load_kv(self, meta, kv_caches) -> None:
        if not meta or not meta.loads:
            return
        names = sorted(kv_caches)
        for bid, key in meta.loads:
            parts = [torch.empty(kv_caches[n].shape[1:], dtype=kv_caches[n].dtype) for n in names]
            flat = torch.empty(sum(p.numel() * p.element_size() for p in parts), dtype=torch.uint8)
            if not self.store.get(key + self._my_shard(), flat):
                raise RuntimeError(f"KV block {key} vanished from the store")
            off = 0
            for n, p in zip(names, parts):
                nbytes = p.numel() * p.element_size()
                kv_caches[n][bid].copy_(flat[off:off + nbytes].view(p.dtype).view(p.shape), non_blocking=True)
                off += nbytes
            self.num_loaded_blocks += 1

    def save_kv(self, meta, kv_caches) -> None:
        if not meta or not meta.saves:
            return
        names = sorted(kv_caches)
        for bid, key in meta.saves:
            key = key + self._my_shard()
            if self.store.contains(key):
                continue
            flat = torch.cat([kv_caches[n][bid].contiguous().view(torch.uint8).reshape(-1).cpu() for n in names])
            self.store.put(key, flat)
            self.num_saved_blocks += 1


class SharedStorageConnector(BlockConnector):
    def __init__(self, vllm_config, role: str) -> None:
        extra = vllm_config.kv_transfer_config.kv_connector_extra_config
        root = extra.get("shared_storage_path") or extra.get("path") or "/tmp/vllm_b200_kv"
        super().__init__(vllm_config, role, FileBlockStore(root))


class MultiConnector(KVConnectorBase):
    def __init__(self, vllm_config, role: str) -> None:
        super().__init__(vllm_config, role)
        import copy
        self.children = []
        for spec in self.cfg.kv_connector_extra_config.get("connectors", []):
            c = copy.copy(vllm_config)
            kc = copy.copy(self.cfg)
            kc.kv_connector = spec["kv_connector"]
            kc.kv_role = spec.get("kv_role", self.cfg.kv_role)
            kc.kv_connector_extra_config = spec.get("kv_connector_extra_config", {})
            c.kv_transfer_config = kc
            self.children.append(create_connector(c, role))
        self._choice: dict[str, int] = {}

    def get_num_new_matched_tokens(self, request, n):
        best, who = 0, -1
        for i, ch in enumerate(self.children):
            k, _ = ch.get_num_new_matched_tokens(request, n)
            if k > best:
                best, who = k, i
        for i, ch in enumerate(self.children):       # the losers must forget their tentative load
            if i != who and hasattr(ch, "_pending_loads"):
                ch._pending_loads.pop(request.request_id, None)
        return best, False

    def update_state_after_alloc(self, request, blocks, n_ext):
        for ch in self.children:
            ch.update_state_after_alloc(request, blocks, n_ext)

    def build_connector_meta(self, so):
        metas = [ch.build_connector_meta(so) for ch in self.children]
        return metas if any(metas) else None

    def start_load_kv(self, meta, kv_caches):
        for ch, m in zip(self.children, meta or []):
            ch.start_load_kv(m, kv_caches)

    def save_kv(self, meta, kv_caches):
        for ch, m in zip(self.children, meta or []):
            ch.save_kv(m, kv_caches)

    def shutdown(self):
        for ch in self.children:
            ch.shutdown()


_REGISTRY: dict[str, tuple[str, str]] = {
    "SharedStorageConnector": ("vllm_b200.kv_transfer", "SharedStorageConnector"),
    "ExampleConnector": ("vllm_b200.kv_transfer", "SharedStorageConnector"),
    "MultiConnector": ("vllm_b200.kv_transfer", "MultiConnector"),
    "P2pPullConnector": ("vllm_b200.kv_transfer.p2p", "P2pPullConnector"),
    "NixlConnector": ("vllm_b200.kv_transfer.p2p", "P2pPullConnector"),
    "P2pNcclConnector": ("vllm_b200.kv_transfer.p2p", "P2pPullConnector"),
    "OffloadingConnector": ("vllm_b200.kv_offload", "OffloadingConnector"),
    "RemoteStoreConnector": ("vllm_b200.kv_transfer.store_server", "RemoteStoreConnector"),
    "LMCacheConnectorV1": ("vllm_b200.kv_transfer.store_server", "RemoteStoreConnector"),
    "MooncakeStoreConnector": ("vllm_b200.kv_transfer.store_server", "RemoteStoreConnector"),
    "SimpleCPUOffloadConnector": ("vllm_b200.kv_offload", "OffloadingConnector"),
}


def register_connector(name: str, module: str, cls: str) -> None:
    _REGISTRY[name] = (module, cls)


class DecodeBenchConnector(KVConnectorBase):
    """Benchmark helper (reference: kv_connector/v1/decode_bench_connector.py): pretends every prompt's KV already
    exists remotely, so a decode instance can be load-tested at arbitrary context lengths without running prefill —
    the "loaded" blocks are filled with a constant instead of real keys / values."""

    def __init__(self, vllm_config, role: str) -> None:
        super().__init__(vllm_config, role)
        self.fill = float(self.cfg.kv_connector_extra_config.get("fill_value", 0.015))
        self._jobs: list[list[int]] = []
        self.num_filled_blocks = 0

    def get_num_new_matched_tokens(self, request, num_computed_tokens: int) -> tuple[int, bool]:
        n = (request.num_tokens - 1) // self.block_size * self.block_size        # all full blocks but keep the last token
        return max(n - num_computed_tokens, 0), False

    def update_state_after_alloc(self, request, blocks, num_external_tokens: int) -> None:
        if num_external_tokens > 0:
            ids = blocks.get_block_ids()[0]
            self._jobs.append(list(ids[: (request.num_computed_tokens + num_external_tokens) // self.block_size]))

    def build_connector_meta(self, scheduler_output) -> Any:
        jobs, self._jobs = self._jobs, []
        return {"fill": jobs} if jobs else None

    def start_load_kv(self, meta: Any, kv_caches: dict[str, torch.Tensor]) -> None:
        for ids in (meta or {}).get("fill", ()):
            idx = torch.tensor(ids, dtype=torch.long)
            for t in kv_caches.values():
                if t.dtype in (torch.float32, torch.float16, torch.bfloat16):
                    t[idx.to(t.device)] = self.fill
            self.num_filled_blocks += len(ids)


_REGISTRY["DecodeBenchConnector"] = ("vllm_b200.kv_transfer", "DecodeBenchConnector")


def create_connector(vllm_config, role: str = "worker"):
    import importlib
    name = vllm_config.kv_transfer_config.kv_connector
    if name not in _REGISTRY:
        raise ValueError(f"unknown KV connector {name!r}; known: {sorted(_REGISTRY)}")
    mod, cls = _REGISTRY[name]
    return getattr(importlib.import_module(mod), cls)(vllm_config, role)
