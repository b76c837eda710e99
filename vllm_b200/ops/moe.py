"""MoE expert computation on CUDA.

Current implementation: sort tokens by expert (torch), run one GEMM pair per active expert through
``ops.linear`` (tcgen05 kernel for small per-expert M, cuBLASLt otherwise) and scatter-add back.
The grouped tcgen05 kernel with fused dispatch/combine replaces the loop (see csrc/moe.cu).
"""
from __future__ import annotations

from typing import Optional

import torch


GROUPED_MAX_ROWS = 1024      # (token, expert) pairs handled by the single-pass grouped kernels


def _grouped_ok(x, w13, w2, R: int) -> bool:
    from vllm_b200 import envs
    return (x.dtype in (torch.bfloat16, torch.float16) and w13.dtype == x.dtype and w2.dtype == x.dtype
            and x.shape[1] % 64 == 0 and w2.shape[2] % 64 == 0 and R <= GROUPED_MAX_ROWS and x.is_contiguous()
            and w13.is_contiguous() and w2.is_contiguous() and envs.VLLM_B200_GEMM != "cublas")


def fused_moe_grouped(x, w13, w2, topk_w, topk_ids, act, expert_map):
    """Decode-sized MoE, fully on the device (no host sync → CUDA-graph capturable):
    align (counting sort by expert) → gather → grouped tcgen05 GEMM (w13) → act·mul → grouped GEMM (w2) →
    weighted combine.  Empty (expert, chunk) CTAs exit before touching their weights, so the weight traffic is
    exactly that of the experts that were actually routed to."""
    from vllm_b200.ops import cuda as cu
    T, k = topk_ids.shape
    R = T * k
    E = w13.shape[0]
    ids = topk_ids if topk_ids.dtype == torch.int32 else topk_ids.to(torch.int32)
    sorted_tok, inv, offs = cu.moe_align(ids.contiguous(), E, expert_map)
    xs = cu.moe_gather(x, sorted_tok)
    # token-tile width: average rows per expert rounded up, so that most experts need one chunk
    avg = max(R // max(E, 1), 1)
    bn = 16
    while bn < 128 and bn < 2 * avg:
        bn += 16
    h = cu.gemm_tc_grouped(xs, w13, offs, bn)
    a = cu.act_and_mul(h, act) if hasattr(cu, "act_and_mul") else None
    if a is None:
        from vllm_b200 import ops
        a = ops.silu_and_mul(h) if act == "silu" else ops.gelu_and_mul(h)
    y = cu.gemm_tc_grouped(a, w2, offs, bn)
    return cu.moe_combine(y, inv, topk_w.to(torch.float32).contiguous(), T)


def fused_moe_cuda(x, w13, w2, topk_w, topk_ids, act="silu", expert_map: Optional[torch.Tensor] = None):
    from vllm_b200 import ops
    if _grouped_ok(x, w13, w2, topk_ids.numel()):
        return fused_moe_grouped(x, w13, w2, topk_w, topk_ids, act, expert_map)
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
