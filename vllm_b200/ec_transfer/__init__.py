"""Encoder-cache transfer: share multimodal encoder outputs between instances (reference:
vllm/distributed/ec_transfer/ — encoder disaggregation: vision towers run on producer instances, LLM
instances consume the embeddings).

``ECSharedStorageConnector`` keeps projected embeddings in a shared directory keyed by the item's content hash
(the same hash that keys the prefix cache and the worker's ``EncoderCache``): a producer stores what it encodes,
a consumer looks the item up before running (or instead of owning) a vision tower.
"""
from __future__ import annotations

from typing import Optional

import torch

from vllm_b200.kv_transfer import FileBlockStore


class ECSharedStorageConnector:
    def __init__(self, path: str, role: str = "ec_both") -> None:
        self.store = FileBlockStore(path)
        self.producer = role in ("ec_producer", "ec_both")
        self.consumer = role in ("ec_consumer", "ec_both")
        self.meta: dict[str, tuple[tuple[int, ...], torch.dtype]] = {}

    def save(self, mm_hash: str, emb: torch.Tensor) -> None:
        if not self.producer or self.store.contains(mm_hash):
            return
        # layout header: rows, cols, dtype code → first 16 bytes of the blob
        code = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[emb.dtype]
        hdr = torch.tensor([emb.shape[0], emb.shape[1], code, 0], dtype=torch.int32).view(torch.uint8)
        self.store.put(mm_hash, torch.cat([hdr, emb.detach().cpu().contiguous().view(torch.uint8).reshape(-1)]))

    def load(self, mm_hash: str, device=None) -> Optional[torch.Tensor]:
        if not self.consumer or not self.store.contains(mm_hash):
            return None
        hdr = torch.empty(16, dtype=torch.uint8)
        import os
        size = os.path.getsize(self.store._path(mm_hash))
        blob = torch.empty(size, dtype=torch.uint8)
        if not self.store.get(mm_hash, blob):
            return None
        hdr = blob[:16].view(torch.int32)
        dt = (torch.float32, torch.bfloat16, torch.float16)[int(hdr[2])]
        t = blob[16:].view(dt).view(int(hdr[0]), int(hdr[1]))
        return t.to(device) if device is not None else t
