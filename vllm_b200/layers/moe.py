"""Mixture-of-experts layer.

Behavioural reference: vllm/model_executor/layers/fused_moe/layer.py (``FusedMoE``: expert-parallel
expert map, ``weight_loader`` for w1/w2/w3 shards), modular_kernel.py:42-83 (router → dispatch →
experts → combine), router/fused_topk + grouped_topk.

Sharding (B200-first): with ``enable_expert_parallel`` every rank of the TP group owns E/W whole experts
(no intra-expert sharding → every expert GEMM keeps its full K, and the only collective is the combine
all-reduce, which under decode runs through the same fused GEMM-epilogue path as dense layers).  Without
EP the experts are TP-sharded along the intermediate dimension exactly like a dense MLP.
Activations are replicated across the TP group at the layer boundary (they already are after the
attention block's all-reduce), so "dispatch" is a local selection and "combine" is one all-reduce.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from vllm_b200 import ops
from vllm_b200.parallel.state import get_tp_group
from vllm_b200.parallel.utils import divide


class FusedMoE(nn.Module):
    def __init__(self, num_experts: int, top_k: int, hidden_size: int, intermediate_size: int, dtype=None,
                 renormalize: bool = True, use_grouped_topk: bool = False, n_group: int = 1, topk_group: int = 1,
                 scoring_func: str = "softmax", routed_scaling_factor: float = 1.0, e_score_bias: bool = False,
                 act: str = "silu", enable_ep: bool = False, reduce_results: bool = True, prefix: str = "",
                 quant_config=None) -> None:
        super().__init__()
        g = get_tp_group()
        self.tp_rank, self.tp_size = max(g.rank_in_group, 0), g.world_size
        self.num_experts, self.top_k = num_experts, top_k
        self.renormalize = renormalize
        self.use_grouped_topk, self.n_group, self.topk_group = use_grouped_topk, n_group, topk_group
        self.scoring_func, self.routed_scaling_factor = scoring_func, routed_scaling_factor
        self.custom_routing = None          # optional ``fn(router_logits, top_k) -> (weights fp32 [T,k], ids [T,k])``
        self.act = act
        self.reduce_results = reduce_results
        self.ep = enable_ep and self.tp_size > 1
        # wide EP: the expert-parallel group also spans the data-parallel replicas (EP = DP x TP); tokens differ per
        # replica, so the layer dispatches / combines through parallel/all2all.py instead of one TP all-reduce
        from vllm_b200.parallel.state import get_dp_group, get_ep_group
        epg = get_ep_group()
        self.wide_ep = bool(enable_ep and epg.world_size > self.tp_size)
        self.a2a = None
        ep_rank, ep_size = self.tp_rank, self.tp_size
        if self.wide_ep:
            from vllm_b200.config import get_current_vllm_config
            from vllm_b200.parallel.all2all import All2AllManager
            self.ep = True
            ep_rank, ep_size = max(epg.rank_in_group, 0), epg.world_size
            pcfg = get_current_vllm_config().parallel_config
            self.a2a = All2AllManager(epg, get_dp_group(), g, num_experts, pcfg.all2all_backend, pcfg.enable_dbo)
        dt = dtype or torch.get_default_dtype()
        # EPLB: ``num_redundant_experts`` extra physical slots hold replicas of hot experts; the router's logical ids
        # are mapped to physical slots per token, the slots are re-assigned (and their weights moved between ranks) by
        # ``eplb_apply`` whenever the runner decides to rebalance (reference: distributed/eplb/eplb_state.py)
        self.eplb = False
        self.num_physical = num_experts
        if self.ep and not self.wide_ep:
            from vllm_b200.config import get_current_vllm_config_or_none
            vc = get_current_vllm_config_or_none()
            pc = vc.parallel_config if vc is not None else None
            if pc is not None and pc.enable_eplb:
                R = pc.num_redundant_experts
                if (num_experts + R) % ep_size:
                    raise ValueError(f"num_experts + num_redundant_experts ({num_experts}+{R}) must divide by EP size {ep_size}")
                self.eplb, self.num_physical = True, num_experts + R
        if self.eplb:
            P, per = self.num_physical, self.num_physical // ep_size
            self.ep_rank, self.ep_size = ep_rank, ep_size
            self.e_start, self.e_end, self.local_experts = ep_rank * per, (ep_rank + 1) * per, per
            phy2log = torch.cat([torch.arange(num_experts), torch.arange(P - num_experts) % num_experts])
            self.register_buffer("phy2log", phy2log.long(), persistent=False)
            self.register_buffer("eplb_load", torch.zeros(num_experts, dtype=torch.int64), persistent=False)
            self._set_placement(phy2log)
            self.inter_per_rank = intermediate_size
        elif self.ep:
            # contiguous expert ranges; the map turns a global expert id into a local one (or -1)
            per = (num_experts + ep_size - 1) // ep_size
            self.e_start = min(num_experts, ep_rank * per)
            self.e_end = min(num_experts, self.e_start + per)
            self.local_experts = self.e_end - self.e_start
            emap = torch.full((num_experts,), -1, dtype=torch.int32)
            emap[self.e_start:self.e_end] = torch.arange(self.local_experts, dtype=torch.int32)
            self.register_buffer("expert_map", emap, persistent=False)
            self.inter_per_rank = intermediate_size
        else:
            self.e_start, self.e_end, self.local_experts = 0, num_experts, num_experts
            self.expert_map = None
            self.inter_per_rank = divide(intermediate_size, self.tp_size)
        I = self.inter_per_rank
        self.w13_weight = nn.Parameter(torch.empty(max(self.local_experts, 1), 2 * I, hidden_size, dtype=dt),
                                       requires_grad=False)
        self.w2_weight = nn.Parameter(torch.empty(max(self.local_experts, 1), hidden_size, I, dtype=dt),
                                      requires_grad=False)
        self.e_score_correction_bias = (nn.Parameter(torch.zeros(num_experts, dtype=torch.float32),
                                                     requires_grad=False) if e_score_bias else None)
        # block-scaled fp8 experts (DeepSeek-V3 checkpoints: ``weight`` e4m3 + ``weight_scale_inv`` per 128x128 block;
        # or online quantisation of a bf16 checkpoint).  Only the experts a batch routes to are expanded, so decode
        # traffic stays at the fp8 size.
        self.fp8_block = 0
        self.fp8_online = False
        self.compute_dtype = dt
        if quant_config is not None and getattr(quant_config, "name", "") == "fp8_block":
            blk = quant_config.block
            self.fp8_block = blk
            self.fp8_online = bool(quant_config.online)
            E = max(self.local_experts, 1)
            if not self.fp8_online:
                self.w13_weight = nn.Parameter(torch.empty(E, 2 * I, hidden_size, dtype=torch.float8_e4m3fn), requires_grad=False)
                self.w2_weight = nn.Parameter(torch.empty(E, hidden_size, I, dtype=torch.float8_e4m3fn), requires_grad=False)
            cb = lambda n: -(-n // blk)  # noqa: E731
            self.w13_scale = nn.Parameter(torch.ones(E, 2 * cb(I), cb(hidden_size), dtype=torch.float32), requires_grad=False)
            self.w2_scale = nn.Parameter(torch.ones(E, cb(hidden_size), cb(I), dtype=torch.float32), requires_grad=False)

    # ------------------------------------------------------------ routing
    def select_experts(self, router_logits: torch.Tensor):
        if self.custom_routing is not None:                      # model-specific router, e.g. ``sparsemixer_routing``
            return self.custom_routing(router_logits, self.top_k)
        if self.use_grouped_topk:
            return ops.grouped_topk(router_logits, self.top_k, self.n_group, self.topk_group, self.renormalize,
                                    self.scoring_func, self.e_score_correction_bias, self.routed_scaling_factor)
        w, ids = ops.topk_softmax(router_logits, self.top_k, self.renormalize)
        if self.routed_scaling_factor != 1.0:
            w = w * self.routed_scaling_factor
        return w, ids

    def forward(self, x: torch.Tensor, router_logits: torch.Tensor) -> torch.Tensor:
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        w, ids = self.select_experts(router_logits.reshape(-1, router_logits.shape[-1]))
        if self.wide_ep:
            out = self.a2a.dispatch_combine(x2, w.to(torch.float32), ids, self._local_experts,
                                            reduce_tp=self.reduce_results)
            return out.reshape(shape)
        if self.eplb:
            flat = ids.reshape(-1).long()
            self.eplb_load += torch.bincount(flat, minlength=self.num_experts)[:self.num_experts]
            tok = torch.arange(ids.shape[0], device=ids.device)[:, None].expand_as(ids)
            ids = self.log2phy[ids.long(), tok % self.logcnt[ids.long()]].to(ids.dtype)       # replica by token index
        out = self._local_experts(x2, w.to(torch.float32), ids)
        if self.reduce_results and self.tp_size > 1:
            out = get_tp_group().all_reduce(out)
        return out.reshape(shape)

    def forward_reduce_norm(self, x: torch.Tensor, router_logits: torch.Tensor, residual: torch.Tensor, norm):
        """(normed, residual) = norm(all_reduce(moe(x)) + residual) with the combine, the cross-rank reduction, the
        residual add and the RMSNorm fused around the grouped GEMMs (``ops/moe.py::fused_moe_reduce_norm``): the w2
        epilogue red.adds every weighted row into the accumulator of the rank that owns the token.  None = not covered
        (wide EP, EPLB, fp8 experts, prefill-sized batches, CPU) → the caller uses the unfused sequence."""
        if (not x.is_cuda or self.wide_ep or self.eplb or self.fp8_block or norm is None
                or x.dtype not in (torch.bfloat16, torch.float16) or self.w13_weight.dtype != x.dtype):
            return None
        from vllm_b200 import envs
        if envs.VLLM_B200_FORCE_REF_OPS or not envs.VLLM_B200_FUSED_MOE:
            return None
        comm = None
        if self.tp_size > 1:
            comm = get_tp_group().device_comm
            if comm is None:
                return None
        from vllm_b200.ops.moe import fused_moe_reduce_norm
        x2 = x.reshape(-1, x.shape[-1])
        w, ids = self.select_experts(router_logits.reshape(-1, router_logits.shape[-1]))
        out = fused_moe_reduce_norm(x2, self.w13_weight, self.w2_weight, w.to(torch.float32), ids, self.act,
                                    self.expert_map, residual.view(-1, x.shape[-1]), norm.weight, norm.eps,
                                    getattr(norm, "gemma", False), comm)
        if out is None:
            return None
        return out[0].view(x.shape), residual

    # ------------------------------------------------------------ EPLB
    def _set_placement(self, phy2log: torch.Tensor) -> None:
        """(Re)builds the routing tables IN PLACE — captured CUDA graphs keep pointing at the same buffers."""
        E, P = self.num_experts, self.num_physical
        phy2log = phy2log.long().cpu()
        cnt = torch.bincount(phy2log, minlength=E)
        l2p = torch.zeros(E, P - E + 1, dtype=torch.long)                 # a logical expert has at most R + 1 replicas
        fill = [0] * E
        for p_ in range(P):
            e = int(phy2log[p_])
            l2p[e, fill[e]] = p_
            fill[e] += 1
        if not hasattr(self, "log2phy"):
            emap = torch.full((P,), -1, dtype=torch.int32)
            emap[self.e_start:self.e_end] = torch.arange(self.local_experts, dtype=torch.int32)
            self.register_buffer("log2phy", l2p, persistent=False)
            self.register_buffer("logcnt", cnt, persistent=False)
            self.register_buffer("expert_map", emap, persistent=False)
        else:
            self.log2phy.copy_(l2p)
            self.logcnt.copy_(cnt)
        self.phy2log.copy_(phy2log)

    @torch.no_grad()
    def eplb_apply(self, new_phy2log: torch.Tensor, group) -> int:
        """Move expert weights so that physical slot p holds logical expert ``new_phy2log[p]`` (every EP rank calls this
        with the same argument).  Slots whose content is unchanged are untouched; a needed expert is copied from a slot
        on the same rank when one exists, otherwise received from a peer rank.  Returns the number of slots rewritten."""
        import torch.distributed as dist
        from vllm_b200.parallel.eplb import plan_transfers
        moves = plan_transfers(self.phy2log.cpu(), new_phy2log.cpu(), self.ep_size)
        per, me = self.local_experts, self.ep_rank
        tensors = [self.w13_weight, self.w2_weight] + ([self.w13_scale, self.w2_scale] if self.fp8_block else [])
        snap = [t.data.clone() for t in tensors]                       # reads see the OLD contents
        ops, recv_bufs = [], []
        for src, dst, sr, dr in moves:
            if sr == me and dr == me:
                for t, s0 in zip(tensors, snap):
                    t.data[dst - me * per].copy_(s0[src - me * per])
            elif sr == me:
                for s0 in snap:
                    ops.append(dist.P2POp(dist.isend, s0[src - me * per].contiguous(), group.ranks[dr], group.device_group))
            elif dr == me:
                for t in tensors:
                    buf = torch.empty_like(t.data[dst - me * per])
                    recv_bufs.append((t, dst - me * per, buf))
                    ops.append(dist.P2POp(dist.irecv, buf, group.ranks[sr], group.device_group))
        if ops:
            for w_ in dist.batch_isend_irecv(ops):
                w_.wait()
        for t, slot, buf in recv_bufs:
            t.data[slot].copy_(buf)
        self._set_placement(new_phy2log)
        self.eplb_load.zero_()
        return len(moves)

    def _local_experts(self, x2: torch.Tensor, w: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
        """Contribution of the experts this rank owns (ids are global; the expert map drops the others)."""
        if self.fp8_block and self.w13_weight.dtype == torch.float8_e4m3fn:
            return self._forward_fp8(x2, w, ids)
        return ops.fused_moe(x2, self.w13_weight, self.w2_weight, w, ids, self.act, self.expert_map)

    def _dequant(self, wq: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
        blk = self.fp8_block
        s = scale.repeat_interleave(blk, -2)[..., :wq.shape[-2], :].repeat_interleave(blk, -1)[..., :wq.shape[-1]]
        return (wq.float() * s).to(self.compute_dtype)

    def _forward_fp8(self, x: torch.Tensor, topk_w: torch.Tensor, topk_ids: torch.Tensor) -> torch.Tensor:
        """Expand only the routed experts to the compute dtype and run them through the regular fused path with a
        compacted expert numbering."""
        from vllm_b200.ops.moe import fp8_grouped_ok, fused_moe_fp8_grouped
        if self.act == "silu" and fp8_grouped_ok(x, self.w13_weight, self.w2_weight, topk_ids.numel(), self.fp8_block):
            return fused_moe_fp8_grouped(x, self.w13_weight, self.w13_scale, self.w2_weight, self.w2_scale, topk_w, topk_ids,
                                         self.expert_map)
        ids = topk_ids.long()
        if self.expert_map is not None:
            ids = self.expert_map[ids].long()
        active = torch.unique(ids[ids >= 0])
        if active.numel() == 0:
            return torch.zeros_like(x)
        A = active.numel()
        remap = torch.full((max(self.local_experts, 1) + 1,), A, dtype=torch.int32, device=x.device)
        remap[active] = torch.arange(A, dtype=torch.int32, device=x.device)
        w13 = self._dequant(self.w13_weight[active], self.w13_scale[active])
        w2 = self._dequant(self.w2_weight[active], self.w2_scale[active])
        compact = remap[torch.where(ids >= 0, ids, torch.full_like(ids, remap.numel() - 1))]   # A = "not on this rank"
        emap = torch.cat([torch.arange(A, dtype=torch.int32, device=x.device),
                          torch.tensor([-1], dtype=torch.int32, device=x.device)])
        return ops.fused_moe(x, w13, w2, topk_w, compact.int(), self.act, emap)

    def process_weights_after_loading(self) -> None:
        if self.fp8_block and self.fp8_online and self.w13_weight.dtype != torch.float8_e4m3fn:
            from vllm_b200.ops import ref
            qs13, ss13, qs2, ss2 = [], [], [], []
            for e in range(self.w13_weight.shape[0]):
                q, sc = ref.block_quant_weight_fp8(self.w13_weight.data[e], self.fp8_block)
                qs13.append(q); ss13.append(sc)
                q, sc = ref.block_quant_weight_fp8(self.w2_weight.data[e], self.fp8_block)
                qs2.append(q); ss2.append(sc)
            self.w13_weight = nn.Parameter(torch.stack(qs13), requires_grad=False)
            self.w2_weight = nn.Parameter(torch.stack(qs2), requires_grad=False)
            self.w13_scale = nn.Parameter(torch.stack(ss13).float(), requires_grad=False)
            self.w2_scale = nn.Parameter(torch.stack(ss2).float(), requires_grad=False)

    # ------------------------------------------------------------ weights
    def scale_loader(self, param: nn.Parameter, w: torch.Tensor, shard_id: str, expert_id: int) -> None:
        """``weight_scale_inv`` of one expert projection (block granularity; TP shards in units of blocks)."""
        if not (self.e_start <= expert_id < self.e_end):
            return
        le = expert_id - self.e_start
        blk = self.fp8_block
        Ib = -(-self.inter_per_rank // blk)
        r0 = 0 if self.ep else self.tp_rank * Ib
        if shard_id == "w2":
            param.data[le].copy_(w[:, r0:r0 + Ib])
        else:
            off = 0 if shard_id == "w1" else Ib
            param.data[le, off:off + Ib].copy_(w[r0:r0 + Ib])

    def weight_loader(self, param: nn.Parameter, w: torch.Tensor, shard_id: Optional[str] = None,
                      expert_id: Optional[int] = None) -> None:
        """shard_id: 'w1' (gate), 'w3' (up), 'w2' (down); ``w`` is the checkpoint's full tensor."""
        if expert_id is None:                          # router-side tensors (selection bias): replicated, whole
            param.data.copy_(w.reshape(param.shape))
            return
        if self.eplb:                                  # every local physical slot that holds this logical expert
            I = self.inter_per_rank
            for p_ in (self.phy2log[self.e_start:self.e_end] == expert_id).nonzero().flatten().tolist():
                if shard_id == "w2":
                    param.data[p_].copy_(w)
                else:
                    off = 0 if shard_id == "w1" else I
                    param.data[p_, off:off + I].copy_(w)
            return
        if not (self.e_start <= expert_id < self.e_end):
            return
        le = expert_id - self.e_start
        I = self.inter_per_rank
        if self.ep:
            r0 = 0
        else:
            r0 = self.tp_rank * I
        if shard_id == "w2":
            param.data[le].copy_(w[:, r0:r0 + I])
        else:
            off = 0 if shard_id == "w1" else I
            param.data[le, off:off + I].copy_(w[r0:r0 + I])


def make_expert_params_mapping(gate_name: str, down_name: str, up_name: str, num_experts: int,
                               prefix: str = "experts") -> list[tuple[str, str, int, str]]:
    """(param_name, checkpoint_weight_name, expert_id, shard_id) — reference: FusedMoE.make_expert_params_mapping."""
    out = []
    for e in range(num_experts):
        for shard, ck in (("w1", gate_name), ("w2", down_name), ("w3", up_name)):
            pname = f"{prefix}.w2_weight" if shard == "w2" else f"{prefix}.w13_weight"
            out.append((pname, f"{prefix}.{e}.{ck}.weight", e, shard))
    return out


def sparsemixer_routing(jitter_eps: float = 0.01):
    """Phi-3.5-MoE router at inference (reference: models/phimoe.py ``phimoe_routing_function`` / ``sparsemixer``):
    experts are picked one at a time; for each pick, logits further than ``2 * jitter_eps`` (relative) below the
    current maximum are masked before the softmax that yields the pick's weight, and the picked expert is removed
    from the next round.  Plugged into ``FusedMoE.custom_routing``."""
    def route(logits: torch.Tensor, top_k: int):
        scores = logits.float()
        work = scores
        ws, ids = [], []
        for _ in range(top_k):
            mx, idx = work.max(dim=-1, keepdim=True)
            factor = scores.abs().clamp(min=mx)
            masked = work.masked_fill(((mx - scores) / factor) > (2 * jitter_eps), float("-inf"))
            ws.append(torch.softmax(masked, dim=-1).gather(-1, idx))
            ids.append(idx)
            work = work.scatter(-1, idx, float("-inf"))
        return torch.cat(ws, dim=-1), torch.cat(ids, dim=-1).to(torch.int32)
    return route
