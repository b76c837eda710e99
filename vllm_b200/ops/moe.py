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


def fused_moe_fp8_grouped(x, w13q, w13_scale, w2q, w2_scale, topk_w, topk_ids, expert_map):
    """Block-scaled fp8 experts on the device (DeepSeek-V3 checkpoints): align → gather → per-(token,128) fp8 quant →
    grouped tcgen05 f8f6f4 GEMM (w13) → fused SiLU·mul + re-quant → grouped fp8 GEMM (w2) → weighted combine.  Weight
    traffic stays at the fp8 size (the previous path expanded the routed experts to bf16 on every call)."""
    from vllm_b200.ops import cuda as cu
    T, k = topk_ids.shape
    R = T * k
    E = w13q.shape[0]
    ids = topk_ids if topk_ids.dtype == torch.int32 else topk_ids.to(torch.int32)
    sorted_tok, inv, offs = cu.moe_align(ids.contiguous(), E, expert_map)
    xs = cu.moe_gather(x, sorted_tok)
    avg = max(R // max(E, 1), 1)
    bn = 16
    while bn < 128 and bn < 2 * avg:
        bn += 16
    if bn in (80, 112):
        bn += 16
    xq, xsc = cu.per_token_group_quant_fp8(xs, 128, col_major=True)
    h = cu.gemm_fp8_grouped(xq, xsc, w13q, w13_scale.contiguous(), offs, bn, x.dtype)
    aq, asc = cu.act_mul_quant(h, "silu", mode="group", group=128, col_major=True)
    y = cu.gemm_fp8_grouped(aq, asc, w2q, w2_scale.contiguous(), offs, bn, x.dtype)
    return cu.moe_combine(y, inv, topk_w.to(torch.float32).contiguous(), T)


def fp8_grouped_ok(x, w13q, w2q, R: int, block: int) -> bool:
    return (x.is_cuda and x.dtype in (torch.bfloat16, torch.float16) and block == 128 and R <= GROUPED_MAX_ROWS
            and w13q.dtype == torch.float8_e4m3fn and w2q.dtype == torch.float8_e4m3fn and x.shape[1] % 128 == 0
            and w2q.shape[2] % 128 == 0 and w13q.is_contiguous() and w2q.is_contiguous() and x.is_contiguous())


_local_acc: dict[int, torch.Tensor] = {}


def fused_moe_reduce_norm(x, w13, w2, topk_w, topk_ids, act, expert_map, residual, norm_w, eps, gemma=False, comm=None):
    """MoE block tail inside the GEMM: align → gather → grouped GEMM (w13) → act·mul → grouped GEMM (w2) whose epilogue
    scales every output row by its routing weight and ADDS it (red.add.f32) into the fp32 accumulator row of its token —
    on one GPU a local buffer; under TP / EP the accumulator of the rank that OWNS the token, reached straight over
    NVLink (peer pointers of the symmetric heap): top-k combine + reduce-scatter with no staging buffer and no separate
    combine / collective launch; a final kernel (barrier → owned rows + residual → RMSNorm → multicast → barrier, or
    the local finish kernel on one GPU) produces the next block's input and re-zeroes the accumulator.
    Returns (normed, residual) or None when the shape is not covered (callers fall back to combine + reduce kernels).
    Reference semantics: fused_moe/modular_kernel.py:42-83 (prepare → experts → finalize); baseline being replaced:
    device_communicators/all2all.py:101-150 (AgRs: all-gather + reduce-scatter around the experts)."""
    from vllm_b200.ops import cuda as cu
    from vllm_b200.parallel.symm import LAM_H_MAX, MOE_T_CAP
    T, k = topk_ids.shape
    H = x.shape[1]
    R = T * k
    multi = comm is not None and comm.world > 1
    if cu.batch_invariant():          # red.add epilogues sum a token's k expert rows in arrival order
        return None
    if (not _grouped_ok(x, w13, w2, R) or T > MOE_T_CAP or H > LAM_H_MAX or H % 8 or act not in ("silu", "gelu", "gelu_tanh")
            or (residual is not None and (not residual.is_contiguous() or residual.dtype != x.dtype))
            or (multi and (H // 8 > 1024 or norm_w is None))):
        return None
    E = w13.shape[0]
    ids = topk_ids if topk_ids.dtype == torch.int32 else topk_ids.to(torch.int32)
    sorted_tok, inv, offs, pairs = cu.moe_align(ids.contiguous(), E, expert_map, with_pairs=True)
    xs = cu.moe_gather(x, sorted_tok)
    avg = max(R // max(E, 1), 1)
    bn = 16
    while bn < 128 and bn < 2 * avg:
        bn += 16
    h = cu.gemm_tc_grouped(xs, w13, offs, bn)
    a = cu.act_and_mul(h, act)
    pw = topk_w.to(torch.float32).contiguous()
    if multi:
        acc_off = comm.next_moe_acc()
        per = (T + comm.world - 1) // comm.world
        cu.gemm_tc_grouped_combine(a, w2, offs, bn, pairs, pw, k, H, comm=comm, acc_off=acc_off, per=per)
        out = comm.moe_finish_rows(acc_off, T, H, x.dtype, residual, norm_w, eps, gemma)
        return out, residual
    dev = x.device.index or 0
    acc = _local_acc.get(dev)
    if acc is None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("MoE accumulator must be allocated before CUDA-graph capture")
        acc = _local_acc[dev] = torch.zeros(MOE_T_CAP * LAM_H_MAX, dtype=torch.float32, device=x.device)
    cu.gemm_tc_grouped_combine(a, w2, offs, bn, pairs, pw, k, H, acc_ptr=acc.data_ptr())
    out = torch.empty((T, H), dtype=x.dtype, device=x.device)
    cu.moe_finish(acc.data_ptr(), residual, norm_w, out, T, H, eps, gemma)
    return out, residual


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
