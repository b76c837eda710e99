"""KV-cache specs, sizing and the per-request block manager.

Reference behaviour: vllm/v1/kv_cache_interface.py:434-897 (spec taxonomy),
vllm/v1/core/kv_cache_utils.py (``get_kv_cache_configs``: sizing + auto-fit),
vllm/v1/core/kv_cache_manager.py:232-570 (``get_computed_blocks``,
``allocate_slots``, ``free``), single_type_kv_cache_manager.py (per-type
block reclamation for sliding-window layers), kv_cache_coordinator.py
(hybrid groups sharing one pool).

All groups share one ``BlockPool``; every group uses the same page size in
bytes so a physical block id is valid in every layer's tensor (hybrid models
get one block table per group).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence

from vllm_b200.core.block_pool import BlockPool, KVCacheBlock, make_block_hasher
from vllm_b200.utils import cdiv

# ---------------------------------------------------------------------------
# Specs
# ---------------------------------------------------------------------------


@dataclass(frozen=True)
class KVCacheSpec:
    block_size: int

    @property
    def page_size_bytes(self) -> int:
        raise NotImplementedError

    def max_memory_usage_bytes(self, max_model_len: int) -> int:
        return cdiv(max_model_len, self.block_size) * self.page_size_bytes


@dataclass(frozen=True)
class FullAttentionSpec(KVCacheSpec):
    num_kv_heads: int = 1
    head_size: int = 128
    dtype_size: int = 2
    sliding_window: Optional[int] = None  # informational when hybrid manager is disabled
    dcp: int = 1      # decode-context-parallel degree: ``block_size`` counts sequence positions, of which this rank
                      # stores every dcp-th (the ranks of a DCP group hold replicas of the same KV heads)

    @property
    def phys_block_size(self) -> int:
        return self.block_size // self.dcp

    @property
    def page_size_bytes(self) -> int:
        return 2 * self.phys_block_size * self.num_kv_heads * self.head_size * self.dtype_size


@dataclass(frozen=True)
class MLAAttentionSpec(KVCacheSpec):
    """One latent vector (kv_lora_rank + rope dims) per token, no K/V split."""
    head_size: int = 576
    dtype_size: int = 2
    dcp: int = 1      # decode-context-parallel degree: ``block_size`` counts sequence positions, of which this
                      # rank stores every dcp-th (block_size / dcp physical slots per block)

    @property
    def phys_block_size(self) -> int:
        return self.block_size // self.dcp

    @property
    def page_size_bytes(self) -> int:
        return self.phys_block_size * self.head_size * self.dtype_size


@dataclass(frozen=True)
class SlidingWindowSpec(FullAttentionSpec):
    def max_memory_usage_bytes(self, max_model_len: int) -> int:
        n = min(max_model_len, (self.sliding_window or max_model_len) - 1 + 16384)
        return (cdiv(n, self.block_size) + 1) * self.page_size_bytes


@dataclass(frozen=True)
class ChunkedLocalAttentionSpec(FullAttentionSpec):
    attention_chunk_size: int = 8192


@dataclass(frozen=True)
class MambaSpec(KVCacheSpec):
    """Constant-size recurrent state per request (one "block" per request)."""
    state_bytes: int = 0

    @property
    def page_size_bytes(self) -> int:
        return self.state_bytes

    def max_memory_usage_bytes(self, max_model_len: int) -> int:
        return self.state_bytes


@dataclass(frozen=True)
class CrossAttentionSpec(FullAttentionSpec):
    pass


@dataclass
class KVCacheGroupSpec:
    layer_names: list[str]
    spec: KVCacheSpec


@dataclass
class KVCacheTensor:
    size: int  # bytes
    shared_by: list[str]


@dataclass
class KVCacheConfig:
    num_blocks: int
    kv_cache_tensors: list[KVCacheTensor] = field(default_factory=list)
    kv_cache_groups: list[KVCacheGroupSpec] = field(default_factory=list)


def group_specs(layer_specs: dict[str, KVCacheSpec]) -> list[KVCacheGroupSpec]:
    """Layers with identical specs share a group (= one block table)."""
    groups: dict[KVCacheSpec, list[str]] = {}
    for name, spec in layer_specs.items():
        groups.setdefault(spec, []).append(name)
    return [KVCacheGroupSpec(names, spec) for spec, names in groups.items()]


def get_kv_cache_config(layer_specs: dict[str, KVCacheSpec], available_bytes: int, max_model_len: int,
                        num_gpu_blocks_override: Optional[int] = None,
                        auto_fit_max_model_len: bool = False) -> tuple[KVCacheConfig, int]:
    """Size the pool. Returns (config, possibly reduced max_model_len)."""
    if not layer_specs:
        return KVCacheConfig(num_blocks=2), max_model_len
    groups = group_specs(layer_specs)
    # every layer owns its own tensor of ``num_blocks`` pages (pages of different groups — attention K/V pages,
    # sliding-window pages, recurrent-state pages — may differ in size); block ids are shared across groups, so the
    # pool size is what fits when every layer gets one page per block
    per_block = sum(s.page_size_bytes for s in layer_specs.values())
    page = max(g.spec.page_size_bytes for g in groups)
    num_blocks = int(available_bytes // max(per_block, 1))
    if num_gpu_blocks_override is not None:
        num_blocks = num_gpu_blocks_override
    need_blocks = sum(cdiv(g.spec.max_memory_usage_bytes(max_model_len), g.spec.page_size_bytes) for g in groups) + 1
    if num_blocks < need_blocks:
        if auto_fit_max_model_len and num_blocks > 2:
            bs = groups[0].spec.block_size
            max_model_len = (num_blocks - 1) * bs
        else:
            raise ValueError(
                f"KV cache too small: {num_blocks} blocks available but one request of max_model_len="
                f"{max_model_len} needs {need_blocks}; lower max_model_len or raise "
                f"gpu_memory_utilization")
    tensors = [KVCacheTensor(size=sp.page_size_bytes * num_blocks, shared_by=[name]) for name, sp in layer_specs.items()]
    return KVCacheConfig(num_blocks=num_blocks, kv_cache_tensors=tensors, kv_cache_groups=groups), max_model_len


# ---------------------------------------------------------------------------
# Per-group managers
# ---------------------------------------------------------------------------


class SingleTypeManager:
    """Block bookkeeping for one KV group."""

    def __init__(self, spec: KVCacheSpec, pool: BlockPool, group_id: int) -> None:
        self.spec = spec
        self.block_size = spec.block_size
        self.pool = pool
        self.group_id = group_id
        self.req_to_blocks: dict[str, list[KVCacheBlock]] = {}
        self.num_cached: dict[str, int] = {}

    # number of *additional* blocks needed for ``num_tokens`` total tokens
    def blocks_needed(self, request_id: str, num_tokens: int, new_computed: Sequence[KVCacheBlock]) -> int:
        have = len(self.req_to_blocks.get(request_id, ())) + len(new_computed)
        need = cdiv(num_tokens, self.block_size) - have
        # cached blocks currently sitting in the free list are consumed by touching them
        evictable = sum(1 for b in new_computed if b.ref_cnt == 0 and not b.is_null)
        return max(need, 0) + evictable

    def save_new_computed(self, request_id: str, new_computed: Sequence[KVCacheBlock]) -> None:
        if request_id not in self.req_to_blocks:
            blocks = self.req_to_blocks[request_id] = []
            blocks.extend(new_computed)
            self.num_cached[request_id] = len(new_computed)
        else:
            assert not new_computed, "prefix hits only apply to requests without blocks"

    def allocate_new(self, request_id: str, num_tokens: int) -> list[KVCacheBlock]:
        blocks = self.req_to_blocks.setdefault(request_id, [])
        self.num_cached.setdefault(request_id, 0)
        n = cdiv(num_tokens, self.block_size) - len(blocks)
        if n <= 0:
            return []
        new = self.pool.get_new_blocks(n)
        blocks.extend(new)
        return new

    def cache_blocks(self, request, num_tokens: int) -> None:
        rid = request.request_id
        n_full = num_tokens // self.block_size
        n_cached = self.num_cached.get(rid, 0)
        if n_full <= n_cached:
            return
        self.pool.cache_full_blocks(request, self.req_to_blocks[rid], n_cached, n_full, self.group_id)
        self.num_cached[rid] = n_full

    def free(self, request_id: str) -> None:
        blocks = self.req_to_blocks.pop(request_id, [])
        self.num_cached.pop(request_id, None)
        self.pool.free_blocks(reversed(blocks))

    def remove_skipped_blocks(self, request_id: str, num_computed_tokens: int) -> None:
        """Release blocks no future token can attend to (no-op for full attention)."""

    def find_longest_hit(self, hashes: Sequence[bytes], max_blocks: int) -> list[KVCacheBlock]:
        out: list[KVCacheBlock] = []
        for h in hashes[:max_blocks]:
            got = self.pool.get_cached_block(h, (self.group_id,))
            if got is None:
                break
            out.append(got[0])
        return out

    def num_common_prefix_blocks(self, request_id: str, num_running: int) -> int:
        n = 0
        for b in self.req_to_blocks.get(request_id, ()):
            if b.ref_cnt == num_running:
                n += 1
            else:
                break
        return n


class FullAttentionManager(SingleTypeManager):
    pass


class SlidingWindowManager(SingleTypeManager):
    def _first_useful_block(self, num_computed_tokens: int) -> int:
        sw = self.spec.sliding_window
        first_token = max(0, num_computed_tokens - sw + 1)
        return first_token // self.block_size

    def remove_skipped_blocks(self, request_id: str, num_computed_tokens: int) -> None:
        blocks = self.req_to_blocks.get(request_id)
        if not blocks:
            return
        first = self._first_useful_block(num_computed_tokens)
        freed = []
        for i in range(min(first, len(blocks)) - 1, -1, -1):
            if blocks[i].is_null:
                break  # everything before was freed earlier
            freed.append(blocks[i])
            blocks[i] = self.pool.null_block
        self.pool.free_blocks(freed)

    def find_longest_hit(self, hashes: Sequence[bytes], max_blocks: int) -> list[KVCacheBlock]:
        """Longest prefix L such that the last ``w`` blocks before L are cached;
        earlier positions are filled with the null block."""
        w = cdiv(self.spec.sliding_window - 1, self.block_size)
        null = self.pool.null_block
        run: list[KVCacheBlock] = []
        best_len, best_run = 0, []
        for i in range(min(len(hashes), max_blocks) - 1, -1, -1):
            got = self.pool.get_cached_block(hashes[i], (self.group_id,))
            if got is None:
                run = []
                continue
            run.insert(0, got[0])
            if len(run) >= w or i == 0:
                best_len, best_run = i + len(run), list(run)
                break
        if best_len == 0:
            return []
        return [null] * (best_len - len(best_run)) + best_run


class MambaManager(SingleTypeManager):
    """One state block per request regardless of length."""

    def blocks_needed(self, request_id, num_tokens, new_computed) -> int:
        return 0 if self.req_to_blocks.get(request_id) else 1

    def allocate_new(self, request_id, num_tokens):
        if self.req_to_blocks.get(request_id):          # (an empty list is the "seen, nothing cached" marker)
            return []
        new = self.pool.get_new_blocks(1)
        self.req_to_blocks[request_id] = new
        self.num_cached[request_id] = 0
        return new

    def cache_blocks(self, request, num_tokens) -> None:
        return

    def find_longest_hit(self, hashes, max_blocks):
        return []


def _manager_for(spec: KVCacheSpec, pool: BlockPool, gid: int) -> SingleTypeManager:
    if isinstance(spec, MambaSpec):
        return MambaManager(spec, pool, gid)
    if isinstance(spec, SlidingWindowSpec):
        return SlidingWindowManager(spec, pool, gid)
    return FullAttentionManager(spec, pool, gid)


# ---------------------------------------------------------------------------


class KVCacheBlocks:
    """Per-group block lists handed to the scheduler."""

    __slots__ = ("blocks",)

    def __init__(self, blocks: tuple[list[KVCacheBlock], ...]) -> None:
        self.blocks = blocks

    def get_block_ids(self) -> tuple[list[int], ...]:
        return tuple([b.block_id for b in g] for g in self.blocks)

    def __add__(self, other: "KVCacheBlocks") -> "KVCacheBlocks":
        return KVCacheBlocks(tuple(a + b for a, b in zip(self.blocks, other.blocks)))

    def num_blocks(self) -> int:
        return len(self.blocks[0]) if self.blocks else 0


@dataclass
class PrefixCacheStats:
    requests: int = 0
    queries: int = 0
    hits: int = 0
    reset: bool = False


class KVCacheManager:
    def __init__(self, kv_cache_config: KVCacheConfig, max_model_len: int, enable_caching: bool = True,
                 hash_algo: str = "sha256", enable_kv_cache_events: bool = False, log_stats: bool = False,
                 use_eagle: bool = False) -> None:
        self.cfg = kv_cache_config
        self.max_model_len = max_model_len
        self.enable_caching = enable_caching
        self.log_stats = log_stats
        self.use_eagle = use_eagle
        groups = kv_cache_config.kv_cache_groups
        self.block_size = groups[0].spec.block_size if groups else 16
        self.pool = BlockPool(kv_cache_config.num_blocks, enable_caching, enable_kv_cache_events,
                              self.block_size)
        self.managers = [_manager_for(g.spec, self.pool, i) for i, g in enumerate(groups)]
        self.num_groups = len(self.managers)
        self.block_hasher = make_block_hasher(self.block_size, hash_algo) if enable_caching else None
        self.prefix_cache_stats = PrefixCacheStats()
        self._empty = KVCacheBlocks(tuple([] for _ in self.managers))
        # steady-state decode: one full-attention group → "does the last block still have room" is one comparison
        self._single_full = len(self.managers) == 1 and type(self.managers[0]) is FullAttentionManager

    @property
    def usage(self) -> float:
        return self.pool.get_usage()

    def make_prefix_cache_stats(self) -> Optional[PrefixCacheStats]:
        if not self.log_stats:
            return None
        s, self.prefix_cache_stats = self.prefix_cache_stats, PrefixCacheStats()
        return s

    def empty_blocks(self) -> KVCacheBlocks:
        return self._empty

    # -- prefix cache lookup -------------------------------------------------
    def get_computed_blocks(self, request) -> tuple[KVCacheBlocks, int]:
        if (not self.enable_caching
                or (request.sampling_params is not None and request.sampling_params.prompt_logprobs is not None)
                or request.pooling_params is not None and request.pooling_params.pooling_type == "ALL"):
            return self._empty, 0
        # never hit the whole prompt: the last token must be recomputed to get logits
        max_hit_tokens = request.num_tokens - 1
        max_blocks = max_hit_tokens // self.block_size
        hashes = request.block_hashes
        per_group = [m.find_longest_hit(hashes, max_blocks) for m in self.managers]
        n = min((len(g) for g in per_group), default=0)
        if any(len(g) != n for g in per_group):
            # hybrid: shrink to a length every group can serve
            while n > 0:
                per_group = [m.find_longest_hit(hashes, n) for m in self.managers]
                n2 = min(len(g) for g in per_group)
                if n2 == n:
                    break
                n = n2
            if n == 0:
                per_group = [[] for _ in self.managers]
        if self.use_eagle and n > 0:
            # the draft head needs the hidden state of the last cached token → drop one block
            n -= 1
            per_group = [g[:n] for g in per_group]
        if self.log_stats:
            s = self.prefix_cache_stats
            s.requests += 1
            s.queries += request.num_tokens
            s.hits += n * self.block_size
        return KVCacheBlocks(tuple(per_group)), n * self.block_size

    # -- allocation -------------------------------------------------------------
    def allocate_slots(self, request, num_new_tokens: int, num_new_computed_tokens: int = 0,
                       new_computed_blocks: Optional[KVCacheBlocks] = None, num_lookahead_tokens: int = 0,
                       delay_cache_blocks: bool = False, num_encoder_tokens: int = 0
                       ) -> Optional[KVCacheBlocks]:
        if num_new_tokens == 0 and num_new_computed_tokens == 0:
            raise ValueError("num_new_tokens must be > 0")
        rid = request.request_id
        if self._single_full and new_computed_blocks is None and num_new_computed_tokens == 0:
            # fast path (every decode step of every running request): nothing to allocate while the tokens fit in the
            # blocks the request already owns; a block is (hash-)cached only when a token completes it
            m = self.managers[0]
            blocks = m.req_to_blocks.get(rid)
            if blocks is not None:
                nc = request.num_computed_tokens
                total = nc + num_new_tokens + num_lookahead_tokens
                if total > self.max_model_len:
                    total = self.max_model_len
                bs = m.block_size
                if total <= len(blocks) * bs:
                    if self.enable_caching and not delay_cache_blocks:
                        final = nc + num_new_tokens
                        nt = request.num_tokens
                        if final > nt:
                            final = nt
                        if final // bs > m.num_cached.get(rid, 0):
                            m.cache_blocks(request, final)
                    return self._empty
        new_computed = new_computed_blocks.blocks if new_computed_blocks is not None else \
            tuple([] for _ in self.managers)
        # reclaim out-of-window blocks first so they can serve this very allocation
        for m in self.managers:
            m.remove_skipped_blocks(rid, request.num_computed_tokens)
        num_computed = request.num_computed_tokens + num_new_computed_tokens
        total = min(num_computed + num_new_tokens + num_lookahead_tokens, self.max_model_len)
        needed = sum(m.blocks_needed(rid, total, nc) for m, nc in zip(self.managers, new_computed))
        if needed > self.pool.get_num_free_blocks():
            return None
        if self.enable_caching:
            for nc in new_computed:
                self.pool.touch(nc)
        else:
            assert all(not nc for nc in new_computed)
        new_blocks = []
        for m, nc in zip(self.managers, new_computed):
            m.save_new_computed(rid, nc)
            new_blocks.append(m.allocate_new(rid, total))
        if self.enable_caching and not delay_cache_blocks:
            # only tokens that are final after this step (exclude speculative lookahead)
            final = min(num_computed + num_new_tokens, request.num_tokens)
            for m in self.managers:
                m.cache_blocks(request, final)
        return KVCacheBlocks(tuple(new_blocks))

    def cache_blocks(self, request, num_computed_tokens: int) -> None:
        if self.enable_caching:
            for m in self.managers:
                m.cache_blocks(request, min(num_computed_tokens, request.num_tokens))

    def free(self, request) -> None:
        for m in self.managers:
            m.free(request.request_id)

    def reset_prefix_cache(self) -> bool:
        ok = self.pool.reset_prefix_cache()
        if ok and self.log_stats:
            self.prefix_cache_stats.reset = True
        return ok

    def get_blocks(self, request_id: str) -> KVCacheBlocks:
        return KVCacheBlocks(tuple(list(m.req_to_blocks.get(request_id, [])) for m in self.managers))

    def get_block_ids(self, request_id: str) -> tuple[list[int], ...]:
        return self.get_blocks(request_id).get_block_ids()

    def get_num_common_prefix_blocks(self, request_id: str, num_running: int) -> list[int]:
        return [m.num_common_prefix_blocks(request_id, num_running) for m in self.managers]

    def take_events(self):
        return self.pool.take_events()
