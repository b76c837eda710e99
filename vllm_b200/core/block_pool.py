"""Paged-KV block pool with hash-chained prefix cache and LRU free list.

Reference behaviour: vllm/v1/core/block_pool.py (``cache_full_blocks`` :225,
``get_cached_block`` :198, ``get_new_blocks`` :647, ``free_blocks`` :719) and
kv_cache_utils.py (``FreeKVCacheBlockQueue``, ``hash_block_tokens``).

Design notes (our own):
* blocks are plain ``__slots__`` objects in a preallocated list; the free list
  is an intrusive doubly linked list so eviction candidates can be removed from
  the middle in O(1) when a cached-but-free block is re-used (prefix hit);
* freed blocks are appended in *reverse* order of the request's block table so
  the tail of a sequence is evicted before its head (deeper prefix = less
  likely to be shared);
* block 0 is a permanently reserved null block (padding target for CUDA-graph
  replays and for sliding-window holes).
"""
from __future__ import annotations

import hashlib
import struct
from typing import Callable, Iterable, Optional, Sequence

from vllm_b200.core.kv_events import AllBlocksCleared, BlockRemoved, BlockStored, KVCacheEvent

NONE_HASH = b"\x00" * 16


def _sha256_hash(parent: bytes, tokens: Sequence[int], extra: Optional[tuple]) -> bytes:
    h = hashlib.sha256()
    h.update(parent)
    h.update(struct.pack(f"<{len(tokens)}i", *tokens))
    if extra:
        h.update(repr(extra).encode())
    return h.digest()[:16]


def _builtin_hash(parent: bytes, tokens: Sequence[int], extra: Optional[tuple]) -> bytes:
    return hash((parent, tuple(tokens), extra)).to_bytes(8, "little", signed=True)


HASH_FNS: dict[str, Callable] = {"sha256": _sha256_hash, "builtin": _builtin_hash}


def make_block_hasher(block_size: int, algo: str = "sha256"):
    """Return ``fn(request) -> list[new block hashes]``; incremental: only
    hashes full blocks not hashed yet (called on creation and on each append)."""
    hash_fn = HASH_FNS[algo]

    def hasher(request) -> list[bytes]:
        done = len(request.block_hashes)
        toks = request.all_token_ids
        n_full = len(toks) // block_size
        if n_full <= done:
            return []
        parent = request.block_hashes[-1] if done else NONE_HASH
        out = []
        for i in range(done, n_full):
            extra = _extra_keys(request, i * block_size, (i + 1) * block_size)
            parent = hash_fn(parent, toks[i * block_size:(i + 1) * block_size], extra)
            out.append(parent)
        return out

    return hasher


def _extra_keys(request, start: int, end: int) -> Optional[tuple]:
    """Keys that make otherwise-equal token blocks distinct: LoRA adapter,
    cache salt (first block only), multimodal item hashes overlapping the block."""
    keys = []
    if request.lora_request is not None:
        keys.append(("lora", request.lora_request.lora_int_id))
    if request.cache_salt and start == 0:
        keys.append(("salt", request.cache_salt))
    for f in request.mm_features:
        off, ln = f["offset"], f["length"]
        if off < end and off + ln > start:
            keys.append(("mm", f["hash"]))
    return tuple(keys) if keys else None


class KVCacheBlock:
    __slots__ = ("block_id", "ref_cnt", "block_hash", "group_id", "prev_free", "next_free", "is_null")

    def __init__(self, block_id: int) -> None:
        self.block_id = block_id
        self.ref_cnt = 0
        self.block_hash: Optional[bytes] = None
        self.group_id = 0
        self.prev_free: Optional["KVCacheBlock"] = None
        self.next_free: Optional["KVCacheBlock"] = None
        self.is_null = False

    def __repr__(self) -> str:
        return f"Block({self.block_id}, ref={self.ref_cnt}, hashed={self.block_hash is not None})"


class FreeBlockQueue:
    """Intrusive doubly linked list with sentinels; head = next eviction victim."""

    def __init__(self, blocks: Iterable[KVCacheBlock]) -> None:
        self.head = KVCacheBlock(-1)
        self.tail = KVCacheBlock(-1)
        self.head.next_free = self.tail
        self.tail.prev_free = self.head
        self.num_free = 0
        for b in blocks:
            self.append(b)

    def append(self, b: KVCacheBlock) -> None:
        last = self.tail.prev_free
        last.next_free = b
        b.prev_free = last
        b.next_free = self.tail
        self.tail.prev_free = b
        self.num_free += 1

    def remove(self, b: KVCacheBlock) -> None:
        p, n = b.prev_free, b.next_free
        assert p is not None and n is not None, f"{b} is not in the free queue"
        p.next_free = n
        n.prev_free = p
        b.prev_free = b.next_free = None
        self.num_free -= 1

    def popleft(self) -> KVCacheBlock:
        b = self.head.next_free
        if b is self.tail:
            raise RuntimeError("no free KV blocks")
        self.remove(b)
        return b

    def popleft_n(self, n: int) -> list[KVCacheBlock]:
        return [self.popleft() for _ in range(n)]

    def all_free(self) -> list[KVCacheBlock]:
        out, b = [], self.head.next_free
        while b is not self.tail:
            out.append(b)
            b = b.next_free
        return out


class BlockPool:
    def __init__(self, num_blocks: int, enable_caching: bool = True, enable_kv_cache_events: bool = False,
                 block_size: int = 32) -> None:
        assert num_blocks >= 2, "need at least the null block and one usable block"
        self.num_blocks = num_blocks
        self.block_size = block_size
        self.enable_caching = enable_caching
        self.blocks = [KVCacheBlock(i) for i in range(num_blocks)]
        self.free_queue = FreeBlockQueue(self.blocks)
        # (hash, group) -> {block_id: block}; several blocks may carry the same hash
        self.cached: dict[tuple[bytes, int], dict[int, KVCacheBlock]] = {}
        self.null_block = self.free_queue.popleft()
        self.null_block.is_null = True
        self.enable_kv_cache_events = enable_kv_cache_events
        self.kv_event_queue: list[KVCacheEvent] = []

    # -- prefix cache -------------------------------------------------------
    def get_cached_block(self, block_hash: bytes, group_ids: Sequence[int] = (0,)
                         ) -> Optional[list[KVCacheBlock]]:
        out = []
        for g in group_ids:
            d = self.cached.get((block_hash, g))
            if not d:
                return None
            out.append(next(iter(d.values())))
        return out

    def cache_full_blocks(self, request, blocks: list[KVCacheBlock], num_cached_blocks: int,
                          num_full_blocks: int, group_id: int = 0) -> None:
        """Give hashes to ``blocks[num_cached_blocks:num_full_blocks]``."""
        if num_cached_blocks >= num_full_blocks:
            return
        hashes = request.block_hashes
        assert len(hashes) >= num_full_blocks, (len(hashes), num_full_blocks)
        new_hashes = []
        for i in range(num_cached_blocks, num_full_blocks):
            b = blocks[i]
            if b.is_null or b.block_hash is not None:
                continue
            b.block_hash = hashes[i]
            b.group_id = group_id
            self.cached.setdefault((hashes[i], group_id), {})[b.block_id] = b
            new_hashes.append(hashes[i])
        if self.enable_kv_cache_events and new_hashes:
            parent = hashes[num_cached_blocks - 1] if num_cached_blocks > 0 else None
            bs = self.block_size
            self.kv_event_queue.append(BlockStored(
                block_hashes=new_hashes, parent_block_hash=parent,
                token_ids=request.all_token_ids[num_cached_blocks * bs:num_full_blocks * bs],
                block_size=bs,
                lora_id=request.lora_request.lora_int_id if request.lora_request else None))

    def _maybe_evict(self, b: KVCacheBlock) -> None:
        if b.block_hash is None:
            return
        key = (b.block_hash, b.group_id)
        d = self.cached.get(key)
        if d is not None:
            d.pop(b.block_id, None)
            if not d:
                del self.cached[key]
        if self.enable_kv_cache_events:
            self.kv_event_queue.append(BlockRemoved(block_hashes=[b.block_hash]))
        b.block_hash = None

    # -- allocation -----------------------------------------------------------
    def get_num_free_blocks(self) -> int:
        return self.free_queue.num_free

    def get_new_blocks(self, n: int) -> list[KVCacheBlock]:
        if n > self.free_queue.num_free:
            raise ValueError(f"cannot get {n} free blocks (have {self.free_queue.num_free})")
        out = self.free_queue.popleft_n(n)
        for b in out:
            if self.enable_caching:
                self._maybe_evict(b)
            assert b.ref_cnt == 0
            b.ref_cnt = 1
        return out

    def touch(self, blocks: Iterable[KVCacheBlock]) -> None:
        """Take a reference on cached blocks (removing them from the free list
        if they were eviction candidates)."""
        for b in blocks:
            if b.ref_cnt == 0 and not b.is_null:
                self.free_queue.remove(b)
            b.ref_cnt += 1

    def free_blocks(self, ordered: Iterable[KVCacheBlock]) -> None:
        """``ordered``: eviction priority order (first = evict first)."""
        for b in ordered:
            if b.is_null:
                continue
            b.ref_cnt -= 1
            assert b.ref_cnt >= 0, b
            if b.ref_cnt == 0:
                self.free_queue.append(b)

    def reset_prefix_cache(self) -> bool:
        used = self.num_blocks - self.get_num_free_blocks()
        if used != 1:  # only the null block may be in use
            return False
        self.cached.clear()
        for b in self.blocks:
            b.block_hash = None
        if self.enable_kv_cache_events:
            self.kv_event_queue.append(AllBlocksCleared())
        return True

    def get_usage(self) -> float:
        total = self.num_blocks - 1
        return 1.0 - self.get_num_free_blocks() / total if total else 0.0

    def take_events(self) -> list[KVCacheEvent]:
        ev, self.kv_event_queue = self.kv_event_queue, []
        return ev
