"""MoE expert computation on CUDA.

Current implementation: sort tokens by expert (torch), run one GEMM pair per active expert through
``ops.linear`` (tcgen05 kernel for small per-expert M, cuBLASLt otherwise) and scatter-add back.
The grouped tcgen05 kernel with fused dispatch/combine replaces the loop (see csrc/moe.cu).
"""
from __future__ import annotations

from typing import Optional

import torch


def fused_moe_cuda(x, w13, w2, topk_w, topk_ids, act="silu", expert_map: Optional[torch.Tensor] = None):
    from vllm_b200 import ops
    T, H = x.shape
    E = w13.shape[0]
    k = topk_ids.shape[1]
    ids = topk_ids.long()
    if expert_map is not None:
        ids = expert_map[ids].long()
    flat = ids.reshape(-1)
    order = torch.argsort(flat, stable=True)
    tok = order // k
    sorted_e = flat[order]
    n_invalid = int((sorted_e < 0).sum()) if expert_map is not None else 0
    counts_l = torch.bincount(sorted_e[n_invalid:], minlength=E).tolist()
    out = torch.zeros(T, H, dtype=torch.float32, device=x.device)
    xs = x[tok]
    ws = topk_w.reshape(-1)[order].float()
    off = n_invalid
    for e in range(E):
        n = counts_l[e]
        if n == 0:
            continue
        h = ops.linear(xs[off:off + n], w13[e])
        h = ops.silu_and_mul(h) if act == "silu" else ops.gelu_and_mul(h)
        y = ops.linear(h, w2[e]).float() * ws[off:off + n, None]
        out.index_add_(0, tok[off:off + n], y)
        off += n
    return out.to(x.dtype)
