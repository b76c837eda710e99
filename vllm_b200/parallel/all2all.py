"""Expert-parallel dispatch / combine across an EP group that spans data-parallel replicas ("wide EP").

Behavioural reference: vllm/distributed/device_communicators/all2all.py (``NaiveAll2AllManager``,
``AgRsAll2AllManager`` — all-gather + reduce-scatter, and the DeepEP / pplx style token all-to-all), the
``is_sequence_parallel`` token split inside a TP group (all2all.py:53-58), and the dummy-batch lock-step in
v1/engine/core.py:2184-2265 that keeps every rank inside these collectives.

Two algorithms (``--all2all-backend``):

``allgather_reducescatter`` (also ``naive``)
    every rank sees every replica's tokens (all-gather over the DP group; the TP ranks of a replica already hold the
    same tokens), runs its own experts over all of them, and one all-reduce over the EP group sums the partial
    outputs.  Traffic O(total tokens x H) per rank; no metadata exchange; best for small decode batches.

``--enable-dbo`` software-pipelines the token all-to-all over two micro-batches (see ``_a2a``).

``a2a`` (NCCL token all-to-all between replicas; the in-replica EP path does NOT go through this module: there the
    combine + reduce-scatter is fused into the w2 grouped-GEMM epilogue over peer memory, ``ops/moe.py::fused_moe_reduce_norm``)
    token rows travel only to the rank that owns the routed expert: counts exchange, one ``all_to_all`` of rows +
    expert ids, local grouped expert compute, the reverse ``all_to_all`` of results, weighted sum at the source.
    Inside a TP group each rank dispatches a 1/TP slice of the replica's tokens and the results are all-gathered.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist

from vllm_b200.parallel.state import GroupCoordinator


class All2AllManager:
    def __init__(self, ep_group: GroupCoordinator, dp_group: GroupCoordinator, tp_group: GroupCoordinator,
                 num_experts: int, backend: str = "allgather_reducescatter", dbo: bool = False) -> None:
        self.ep, self.dp, self.tp = ep_group, dp_group, tp_group
        self.W, self.rank = ep_group.world_size, max(ep_group.rank_in_group, 0)
        self.num_experts = num_experts
        self.per = -(-num_experts // self.W)
        self.backend = "agrs" if backend in ("allgather_reducescatter", "naive") else "a2a"
        # dual-batch overlap (reference: v1/worker/gpu_ubatch_wrapper.py + dbo/ubatching.py split the whole forward
        # into two micro-batches on two threads; here the overlap is local to the one place that has communication to
        # hide — the MoE dispatch / combine — and needs no second thread, CUDA-graph pair or yield points)
        self.dbo = bool(dbo) and self.backend == "a2a"
        self.stats = {"dispatched_rows": 0, "received_rows": 0, "calls": 0}

    # expert e lives on EP rank e // per (contiguous ranges, same rule as FusedMoE's expert map)
    def owner(self, ids: torch.Tensor) -> torch.Tensor:
        return torch.div(ids, self.per, rounding_mode="floor").clamp_(max=self.W - 1)

    def _dp_sizes(self, n: int, device) -> list[int]:
        if self.dp.world_size == 1:
            return [n]
        t = torch.tensor([n], dtype=torch.int64)
        out = [torch.zeros(1, dtype=torch.int64) for _ in range(self.dp.world_size)]
        dist.all_gather(out, t, group=self.dp.cpu_group)
        return [int(x) for x in out]

    def _all_gather_v(self, x: torch.Tensor, sizes: list[int]) -> torch.Tensor:
        if self.dp.world_size == 1:
            return x
        mx = max(max(sizes), 1)
        pad = torch.zeros((mx,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        pad[:x.shape[0]] = x
        outs = [torch.empty_like(pad) for _ in sizes]
        dist.all_gather(outs, pad, group=self.dp.device_group)
        return torch.cat([o[:s] for o, s in zip(outs, sizes)], dim=0)

    # ------------------------------------------------------------------------------------------------ AG + RS
    def _agrs(self, x, topk_w, topk_ids, experts_fn, reduce_tp):
        sizes = self._dp_sizes(x.shape[0], x.device)
        X = self._all_gather_v(x, sizes)
        Wt = self._all_gather_v(topk_w, sizes)
        Ids = self._all_gather_v(topk_ids, sizes)
        part = experts_fn(X, Wt, Ids)                                # this rank's experts over everybody's tokens
        # sum over the replicas' ranks with my TP index; the remaining sum over the TP group is either done here or
        # left to the caller's row-parallel reduction (which it then shares with the attention block's epilogue)
        if reduce_tp or self.tp.world_size == 1:
            dist.all_reduce(part, group=self.ep.device_group)
        elif self.dp.world_size > 1:
            dist.all_reduce(part, group=self.dp.device_group)
        me = max(self.dp.rank_in_group, 0)
        start = sum(sizes[:me])
        self.stats["dispatched_rows"] += X.shape[0]
        return part[start:start + sizes[me]]

    # ------------------------------------------------------------------------------------------------ all-to-all
    # The token all-to-all is written as three stages so that two micro-batches can be software-pipelined (DBO):
    #   _a2a_send   : sort (token, expert) pairs by owner rank, exchange counts, start the row / id all-to-all
    #   _a2a_compute: wait for the rows, run the local experts, start the return all-to-all
    #   _a2a_finish : wait for the results, apply routing weights and scatter-add at the source
    # Every collective is issued asynchronously (NCCL runs it on its own stream), so with two micro-batches in flight
    # the rows of one travel over NVLink while the experts of the other run on the SMs.
    def _a2a_send(self, x, topk_w, topk_ids, mine):
        k = topk_ids.shape[1]
        H = x.shape[1]
        ids = topk_ids[mine].reshape(-1).long()                      # [n*k]
        src_tok = mine.repeat_interleave(k)
        dest = self.owner(ids)
        order = torch.argsort(dest, stable=True)
        send_counts = torch.bincount(dest, minlength=self.W).tolist()
        rows = x[src_tok[order]].contiguous()
        eids = ids[order].contiguous()
        recv_counts_t = torch.zeros(self.W, dtype=torch.int64)
        dist.all_to_all_single(recv_counts_t, torch.tensor(send_counts, dtype=torch.int64), group=self.ep.cpu_group)
        recv_counts = recv_counts_t.tolist()
        n_recv = sum(recv_counts)
        recv_rows = torch.empty(n_recv, H, dtype=x.dtype, device=x.device)
        recv_eids = torch.empty(n_recv, dtype=torch.int64, device=x.device)
        h1 = dist.all_to_all_single(recv_rows, rows, recv_counts, send_counts, group=self.ep.device_group, async_op=True)
        h2 = dist.all_to_all_single(recv_eids, eids, recv_counts, send_counts, group=self.ep.device_group, async_op=True)
        self.stats["dispatched_rows"] += rows.shape[0]
        self.stats["received_rows"] += n_recv
        return dict(rows=rows, eids=eids, recv_rows=recv_rows, recv_eids=recv_eids, send=send_counts, recv=recv_counts,
                    wts=topk_w[mine].reshape(-1)[order].to(torch.float32), src=src_tok[order], wait=(h1, h2),
                    id_dtype=topk_ids.dtype)

    def _a2a_compute(self, st, experts_fn):
        for h in st["wait"]:
            h.wait()
        rr = st["recv_rows"]
        n_recv = rr.shape[0]
        # one routed expert per received row, unit weight: the source applies the routing weight after the return trip
        if n_recv:
            y = experts_fn(rr, torch.ones(n_recv, 1, dtype=torch.float32, device=rr.device),
                           st["recv_eids"].view(-1, 1).to(st["id_dtype"]))
        else:
            y = rr
        st["y"] = y.contiguous()                                     # keep alive until the return trip completed
        st["back"] = torch.empty(st["rows"].shape[0], rr.shape[1], dtype=rr.dtype, device=rr.device)
        st["wait"] = (dist.all_to_all_single(st["back"], st["y"], st["send"], st["recv"], group=self.ep.device_group,
                                             async_op=True),)

    def _a2a_finish(self, st, out):
        for h in st["wait"]:
            h.wait()
        out.index_add_(0, st["src"], st["back"].float() * st["wts"][:, None])

    def _a2a(self, x, topk_w, topk_ids, experts_fn, reduce_tp):
        T, H = x.shape
        tp, tpr = self.tp.world_size, max(self.tp.rank_in_group, 0)
        # sequence-parallel split inside the replica: this TP rank dispatches rows tpr, tpr+tp, ...
        mine = torch.arange(tpr, T, tp, device=x.device) if tp > 1 else torch.arange(T, device=x.device)
        out = torch.zeros(T, H, dtype=torch.float32, device=x.device)
        if self.dbo:
            # dual-batch overlap: every rank always runs exactly two micro-batches (possibly empty) so the collective
            # sequence stays identical everywhere:  send0 send1 | compute0 -> ret0 | compute1 -> ret1 | fin0 fin1
            half = (mine.shape[0] + 1) // 2
            a = self._a2a_send(x, topk_w, topk_ids, mine[:half])
            b = self._a2a_send(x, topk_w, topk_ids, mine[half:])      # travels while micro-batch 0 computes
            self._a2a_compute(a, experts_fn)
            self._a2a_compute(b, experts_fn)                         # micro-batch 0's results travel meanwhile
            self._a2a_finish(a, out)
            self._a2a_finish(b, out)
            self.stats["dbo_steps"] = self.stats.get("dbo_steps", 0) + 1
        else:
            a = self._a2a_send(x, topk_w, topk_ids, mine)
            self._a2a_compute(a, experts_fn)
            self._a2a_finish(a, out)
        out = out.to(x.dtype)
        if tp > 1 and reduce_tp:                                     # every TP rank filled a disjoint row set
            dist.all_reduce(out, group=self.tp.device_group)
        return out

    def dispatch_combine(self, x: torch.Tensor, topk_w: torch.Tensor, topk_ids: torch.Tensor,
                         experts_fn: Callable[[torch.Tensor, torch.Tensor, torch.Tensor], torch.Tensor],
                         reduce_tp: bool = True) -> torch.Tensor:
        """``experts_fn(rows, weights[n,k'], global_ids[n,k'])`` → the contribution of THIS rank's experts.
        ``reduce_tp=False`` returns a result that still has to be summed over the TP group (the caller fuses that
        sum with its residual-add / norm epilogue)."""
        self.stats["calls"] += 1
        fn = self._agrs if self.backend == "agrs" else self._a2a
        return fn(x, topk_w, topk_ids, experts_fn, reduce_tp)


_MANAGER: Optional[All2AllManager] = None
