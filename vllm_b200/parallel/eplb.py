"""Expert-parallel load balancing (EPLB).

Behavioural reference: vllm/distributed/eplb/ (``EplbState`` eplb_state.py:230 — sliding-window expert load
tracking, periodic rebalance; ``rebalance_experts`` policy — replicate hot experts into redundant slots and
pack physical experts onto ranks so that per-rank load is even; ``rebalance_execute.py`` — weight shuffling).

Pieces here:
* ``ExpertLoadTracker``: per-layer token counts per logical expert over a sliding window of steps
  (accumulated on the device, summed across the EP group when a rebalance is due).
* ``rebalance_experts``: hierarchical policy — (1) give the redundant slots to the experts with the highest
  load per replica, (2) pack the physical experts onto ranks with longest-processing-time-first so every rank
  gets the same number of slots and near-equal load; returns physical→logical, logical→physical maps.
* ``plan_transfers``: the minimal set of (src_rank, dst_rank, slot) weight copies between two placements.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch


class ExpertLoadTracker:
    def __init__(self, num_layers: int, num_experts: int, window: int = 1000, device="cpu") -> None:
        self.window = window
        self.buf = torch.zeros(window, num_layers, num_experts, dtype=torch.int32, device=device)
        self.step = 0

    def record(self, layer: int, topk_ids: torch.Tensor) -> None:
        """Count this step's routed tokens of one layer (call once per MoE layer per forward pass)."""
        E = self.buf.shape[2]
        slot = self.step % self.window
        self.buf[slot, layer] = torch.bincount(topk_ids.reshape(-1).long(), minlength=E)[:E].to(torch.int32)

    def end_step(self) -> None:
        self.step += 1
        self.buf[self.step % self.window].zero_()

    def load(self, group=None) -> torch.Tensor:
        """[layers, experts] token counts over the window, summed over the EP group."""
        tot = self.buf.sum(0).to(torch.int64)
        if group is not None and group.world_size > 1:
            tot = group.all_reduce(tot)
        return tot


@dataclass
class Placement:
    phy2log: torch.Tensor      # [layers, num_physical] logical expert held by each physical slot
    log2phy: torch.Tensor      # [layers, num_logical, max_replicas] physical slots of each logical expert (-1 pad)
    logcnt: torch.Tensor       # [layers, num_logical] replica count


def _balanced_pack(weights: list[float], n_bins: int, per_bin: int) -> list[int]:
    """LPT packing with a cardinality constraint: item i → bin index; every bin gets exactly ``per_bin`` items."""
    order = sorted(range(len(weights)), key=lambda i: -weights[i])
    load = [0.0] * n_bins
    cnt = [0] * n_bins
    out = [0] * len(weights)
    for i in order:
        b = min((j for j in range(n_bins) if cnt[j] < per_bin), key=lambda j: load[j])
        out[i] = b
        load[b] += weights[i]
        cnt[b] += 1
    return out


def rebalance_experts(load: torch.Tensor, num_physical: int, num_ranks: int) -> Placement:
    """load: [layers, num_logical]. ``num_physical`` = logical + redundant slots, divisible by ``num_ranks``."""
    L, E = load.shape
    assert num_physical >= E and num_physical % num_ranks == 0
    per_rank = num_physical // num_ranks
    phy2log = torch.empty(L, num_physical, dtype=torch.int64)
    logcnt = torch.ones(L, E, dtype=torch.int64)
    for l in range(L):
        w = load[l].double().clamp_min(0)
        cnt = [1] * E
        for _ in range(num_physical - E):                 # hottest per-replica load gets the next replica
            e = max(range(E), key=lambda i: float(w[i]) / cnt[i])
            cnt[e] += 1
        items, weights = [], []
        for e in range(E):
            for _ in range(cnt[e]):
                items.append(e)
                weights.append(float(w[e]) / cnt[e])
        bins = _balanced_pack(weights, num_ranks, per_rank)
        slots = [[] for _ in range(num_ranks)]
        for it, b in zip(items, bins):
            slots[b].append(it)
        phy2log[l] = torch.tensor([e for r in range(num_ranks) for e in sorted(slots[r])])
        logcnt[l] = torch.tensor(cnt)
    maxrep = int(logcnt.max())
    log2phy = torch.full((L, E, maxrep), -1, dtype=torch.int64)
    for l in range(L):
        fill = [0] * E
        for p in range(num_physical):
            e = int(phy2log[l, p])
            log2phy[l, e, fill[e]] = p
            fill[e] += 1
    return Placement(phy2log, log2phy, logcnt)


def route_to_physical(topk_ids: torch.Tensor, placement: Placement, layer: int, salt: Optional[torch.Tensor] = None
                      ) -> torch.Tensor:
    """Map logical expert ids to physical slots, spreading tokens over replicas round-robin by token index."""
    l2p, cnt = placement.log2phy[layer].to(topk_ids.device), placement.logcnt[layer].to(topk_ids.device)
    ids = topk_ids.long()
    T = ids.shape[0]
    tok = (torch.arange(T, device=ids.device)[:, None] if salt is None else salt[:, None]).expand_as(ids)
    rep = tok % cnt[ids]
    return l2p[ids, rep]


def plan_transfers(old: torch.Tensor, new: torch.Tensor, num_ranks: int) -> list[tuple[int, int, int, int]]:
    """old/new: [num_physical] logical expert per slot (one layer).  Returns (src_slot, dst_slot, src_rank, dst_rank)
    for every slot whose content changes, preferring a source on the destination's own rank."""
    P = old.numel()
    per = P // num_ranks
    holders: dict[int, list[int]] = {}
    for p in range(P):
        holders.setdefault(int(old[p]), []).append(p)
    moves = []
    for p in range(P):
        e = int(new[p])
        if int(old[p]) == e:
            continue
        srcs = holders[e]
        local = [s for s in srcs if s // per == p // per]
        s = (local or srcs)[0]
        moves.append((s, p, s // per, p // per))
    return moves
