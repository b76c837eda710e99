"""In-place weight updates for RLHF / online fine-tuning loops.

Behavioural reference: vllm/distributed/weight_transfer/ (NCCL / CUDA-IPC engines, packed tensors) and the
``LLM.collective_rpc`` based update path (entrypoints/llm.py:855-905).

A trainer process joins a dedicated process group with every inference worker (``init_weight_update_group``),
then streams tensors: the metadata list goes through the engine RPC, the payload through one flat broadcast per
bucket (tensors are packed back-to-back into a bucket buffer so a 70B update is a few hundred large
broadcasts instead of thousands of small ones).  Workers feed the unpacked tensors to ``model.load_weights`` —
the same sharding / fusing logic as checkpoint loading — so trainer-side names are plain HF names.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Optional

import torch
import torch.distributed as dist


@dataclass
class TensorMeta:
    name: str
    shape: tuple[int, ...]
    dtype: torch.dtype
    offset: int            # element offset inside the bucket


def pack_buckets(named: Iterable[tuple[str, torch.Tensor]], bucket_bytes: int = 1 << 30
                 ) -> Iterable[tuple[list[TensorMeta], torch.Tensor]]:
    """Group tensors of one dtype into flat buckets of at most ``bucket_bytes``."""
    cur: dict[torch.dtype, tuple[list[TensorMeta], list[torch.Tensor], int]] = {}
    for name, t in named:
        metas, parts, n = cur.get(t.dtype, ([], [], 0))
        if n and (n + t.numel()) * t.element_size() > bucket_bytes:
            yield metas, torch.cat([p.reshape(-1) for p in parts])
            metas, parts, n = [], [], 0
        metas.append(TensorMeta(name, tuple(t.shape), t.dtype, n))
        parts.append(t)
        cur[t.dtype] = (metas, parts, n + t.numel())
    for metas, parts, _ in cur.values():
        if metas:
            yield metas, torch.cat([p.reshape(-1) for p in parts])


def unpack_bucket(metas: list[TensorMeta], flat: torch.Tensor) -> list[tuple[str, torch.Tensor]]:
    out = []
    for m in metas:
        n = 1
        for s in m.shape:
            n *= s
        out.append((m.name, flat[m.offset:m.offset + n].view(m.shape)))
    return out


class WeightTransferGroup:
    """One process group spanning the trainer (rank 0) and all inference workers (ranks 1..N)."""

    def __init__(self, master_addr: str, master_port: int, rank: int, world_size: int, backend: Optional[str] = None,
                 device: Optional[torch.device] = None) -> None:
        from datetime import timedelta
        self.device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
        backend = backend or ("nccl" if self.device.type == "cuda" else "gloo")
        store = dist.TCPStore(master_addr, master_port, world_size, is_master=rank == 0,
                              timeout=timedelta(seconds=600))
        # a stateless group: independent of the default group the workers already use for TP/PP
        if backend == "gloo":
            self.pg = dist.ProcessGroupGloo(store, rank, world_size, timedelta(seconds=600))
        else:
            self.pg = dist.ProcessGroupNCCL(store, rank, world_size)
        self.rank, self.world_size = rank, world_size

    def broadcast(self, t: torch.Tensor, src: int = 0) -> None:
        opts = dist.BroadcastOptions()
        opts.rootRank = src
        self.pg.broadcast([t], opts).wait()

    # trainer side
    def send_weights(self, named: Iterable[tuple[str, torch.Tensor]], notify, bucket_bytes: int = 1 << 30) -> int:
        """``notify(metas)`` must make every worker call ``recv_weights(metas)`` (e.g. through collective_rpc)."""
        n = 0
        for metas, flat in pack_buckets(named, bucket_bytes):
            notify(metas)
            self.broadcast(flat.to(self.device))
            n += len(metas)
        return n

    # worker side
    def recv_weights(self, metas: list[TensorMeta], model) -> int:
        n_last = 1
        for s in metas[-1].shape:
            n_last *= s
        flat = torch.empty(metas[-1].offset + n_last, dtype=metas[0].dtype, device=self.device)
        self.broadcast(flat)
        model.load_weights(iter(unpack_bucket(metas, flat)))
        return len(metas)
