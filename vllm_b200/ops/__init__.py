"""Op dispatch: CUDA tensors → our sm_100a kernels (``ops/cuda.py`` → ``_C.so``),
CPU tensors → the PyTorch reference (``ops/ref.py``).

There is deliberately no multi-backend registry (the reference has
``CustomOp`` + IR providers: model_executor/custom_op.py:103-210, vllm/ir/op.py);
this framework targets exactly one device.  A CUDA tensor with the extension
missing raises — it never silently falls back to eager PyTorch.
"""
from __future__ import annotations

from typing import Optional

import torch

from vllm_b200 import envs
from vllm_b200.ops import ref

_cuda_mod = None


def cuda_ops():
    global _cuda_mod
    if _cuda_mod is None:
        from vllm_b200.ops import cuda as _m
        _cuda_mod = _m
    return _cuda_mod


def _use_cuda(x: torch.Tensor) -> bool:
    return x.is_cuda and not envs.VLLM_B200_FORCE_REF_OPS


def _vec_ok(x: torch.Tensor) -> bool:
    """The row kernels move 16-byte vectors: the row width must be a multiple of 16 bytes."""
    return (x.shape[-1] * x.element_size()) % 16 == 0


def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float, gemma: bool = False) -> torch.Tensor:
    if _use_cuda(x) and _vec_ok(x):
        return cuda_ops().rms_norm(x, weight, eps, gemma)
    return ref.rms_norm(x, weight, eps, gemma)


def fused_add_rms_norm(x: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor, eps: float,
                       gemma: bool = False) -> tuple[torch.Tensor, torch.Tensor]:
    """Returns (normed, new_residual). The CUDA path updates ``residual`` in place and
    writes the normed result into ``x``'s storage."""
    if _use_cuda(x) and _vec_ok(x):
        return cuda_ops().fused_add_rms_norm(x, residual, weight, eps, gemma)
    return ref.fused_add_rms_norm(x, residual, weight, eps, gemma)


def layer_norm(x, weight, bias, eps):
    if _use_cuda(x):
        return cuda_ops().layer_norm(x, weight, bias, eps)
    return ref.layer_norm(x, weight, bias, eps)


def silu_and_mul(x: torch.Tensor) -> torch.Tensor:
    if _use_cuda(x):
        return cuda_ops().act_and_mul(x, "silu")
    return ref.silu_and_mul(x)


def silu_and_mul_interleaved(x: torch.Tensor) -> torch.Tensor:
    """silu(x[..., 0::2]) * x[..., 1::2]: the activation for gate_up weights stored as interleaved (gate_i, up_i) rows."""
    if _use_cuda(x) and x.dtype in (torch.bfloat16, torch.float16) and (x.shape[-1] // 2) % 8 == 0:
        return cuda_ops().silu_and_mul_interleaved(x)
    return (torch.nn.functional.silu(x[..., 0::2].float()) * x[..., 1::2].float()).to(x.dtype)


def gated_silu_linear(x: torch.Tensor, weight_il: torch.Tensor) -> torch.Tensor:
    """silu(x Wg^T) * (x Wu^T) for a weight whose rows interleave (gate_i, up_i).  Decode batches on CUDA run ONE
    kernel: the stream-K tcgen05 GEMM applies SiLU·mul in its epilogue (csrc/gemm_sk.cu)."""
    K = x.shape[-1]
    x2 = x.reshape(-1, K)
    if _use_cuda(x):
        cu = cuda_ops()
        if 0 < x2.shape[0] <= 256 and cu._tc_ok(x2, weight_il) and weight_il.shape[0] % 4 == 0:
            fn = cu.gemm_sk if envs.VLLM_B200_GEMM_SK else cu.gemm_tc
            return fn(x2, weight_il, act=True).view(x.shape[:-1] + (weight_il.shape[0] // 2,))
    return silu_and_mul_interleaved(linear(x, weight_il))


def gelu_and_mul(x: torch.Tensor, approximate: str = "none") -> torch.Tensor:
    if _use_cuda(x):
        return cuda_ops().act_and_mul(x, "gelu_tanh" if approximate == "tanh" else "gelu")
    return ref.gelu_and_mul(x, approximate)


def activation(x: torch.Tensor, name: str) -> torch.Tensor:
    if _use_cuda(x):
        return cuda_ops().activation(x, name)
    return ref.ACT_FNS[name](x)


_HALF = (torch.bfloat16, torch.float16)
_SIMT_ATTN_SHAPES = {(128, 64), (64, 64), (32, 64), (128, 32), (128, 16), (64, 32), (64, 16), (256, 32), (32, 32), (32, 16)}


def _rope_ok(q, cos_sin_cache, head_dim) -> bool:
    """Shapes the vectorised RoPE / cache-write kernel covers (16-byte vectors over half-precision heads)."""
    rot = cos_sin_cache.shape[-1]
    return q.dtype in _HALF and head_dim % 8 == 0 and (rot // 2) % 8 == 0 and (head_dim - rot) % 8 == 0


def cuda_attention_supported(D: int, BS: int, dtype, vt: bool) -> bool:
    return dtype in _HALF and (vt or (D, BS) in _SIMT_ATTN_SHAPES)


def rotary_embedding(positions, q, k, cos_sin_cache, head_dim, is_neox=True):
    """In place on CUDA; returns (q, k)."""
    if _use_cuda(q) and _rope_ok(q, cos_sin_cache, head_dim):
        return cuda_ops().rotary_embedding(positions, q, k, cos_sin_cache, head_dim, is_neox)
    return ref.rotary_embedding(positions, q, k, cos_sin_cache, head_dim, is_neox)


def reshape_and_cache(key, value, kv_cache, slot_mapping, vt: bool = False) -> None:
    if _use_cuda(key) and (vt or (key.dtype in _HALF and key.shape[-1] % 8 == 0)):
        return cuda_ops().reshape_and_cache(key, value, kv_cache, slot_mapping, vt=vt)
    return ref.reshape_and_cache(key, value, kv_cache, slot_mapping)


def rope_and_cache(positions, q, k, v, cos_sin_cache, head_dim, is_neox, kv_cache, slot_mapping, vt: bool = False):
    """Fused RoPE(q,k) + paged KV write. q/k: [T, H*D] views; returns (q, k) rotated.
    ``vt``: the cache uses the tensor-core layout (V pages transposed), CUDA only."""
    if _use_cuda(q) and (vt or _rope_ok(q, cos_sin_cache, head_dim)):
        return cuda_ops().rope_and_cache(positions, q, k, v, cos_sin_cache, head_dim, is_neox, kv_cache,
                                         slot_mapping, vt=vt)
    q, k = ref.rotary_embedding(positions, q, k, cos_sin_cache, head_dim, is_neox)
    Hkv = kv_cache.shape[2]
    ref.reshape_and_cache(k.reshape(k.shape[0], Hkv, head_dim), v.reshape(v.shape[0], Hkv, head_dim),
                          kv_cache, slot_mapping)
    return q, k


def concat_and_cache_mla(kv_c, k_pe, kv_cache, slot_mapping) -> None:
    if _use_cuda(kv_c):
        return cuda_ops().concat_and_cache_mla(kv_c, k_pe, kv_cache, slot_mapping)
    return ref.concat_and_cache_mla(kv_c, k_pe, kv_cache, slot_mapping)


def paged_attention(q, kv_cache, md, scale, sliding_window=None, logits_soft_cap=0.0, causal=True,
                    out: Optional[torch.Tensor] = None, alibi_slopes: Optional[torch.Tensor] = None,
                    sinks: Optional[torch.Tensor] = None, chunk_size: int = 0, return_lse: bool = False):
    """``md``: AttentionMetadata (see vllm_b200.forward_context)."""
    if _use_cuda(q) and cuda_attention_supported(q.shape[-1], kv_cache.shape[3], q.dtype, md.kv_vt):
        return cuda_ops().paged_attention(q, kv_cache, md, scale, sliding_window, logits_soft_cap, causal, out,
                                          return_lse=return_lse, alibi_slopes=alibi_slopes, sinks=sinks,
                                          chunk_size=chunk_size)
    sl = md.seq_lens_cpu if md.seq_lens_cpu is not None else md.seq_lens.cpu()
    qsl = md.query_start_loc_cpu if md.query_start_loc_cpu is not None else md.query_start_loc.cpu()
    return ref.paged_attention(q, kv_cache, md.block_table, sl, qsl, scale,
                               causal=causal, sliding_window=sliding_window, logits_soft_cap=logits_soft_cap,
                               alibi_slopes=alibi_slopes, sinks=sinks, chunk_size=chunk_size, return_lse=return_lse)


def mla_attention(q_nope_abs, q_pe, kv_cache, md, scale, kv_lora_rank):
    if _use_cuda(q_pe):
        return cuda_ops().mla_attention(q_nope_abs, q_pe, kv_cache, md, scale, kv_lora_rank)
    sl = md.seq_lens_cpu if md.seq_lens_cpu is not None else md.seq_lens.cpu()
    qsl = md.query_start_loc_cpu if md.query_start_loc_cpu is not None else md.query_start_loc.cpu()
    return ref.mla_decode_attention(q_nope_abs, q_pe, kv_cache, md.block_table, sl, qsl, scale, kv_lora_rank)


def merge_attn_states(o1, lse1, o2, lse2):
    if _use_cuda(o1):
        return cuda_ops().merge_attn_states(o1, lse1, o2, lse2)
    return ref.merge_attn_states(o1, lse1, o2, lse2)


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = x @ weight.T + bias. CUDA: tcgen05 GEMM for the shapes it covers, else cuBLASLt."""
    if _use_cuda(x):
        return cuda_ops().linear(x, weight, bias)
    return torch.nn.functional.linear(x, weight, bias)


def topk_softmax(gating, k, renormalize):
    if _use_cuda(gating):
        return cuda_ops().topk_softmax(gating, k, renormalize)
    return ref.topk_softmax(gating, k, renormalize)


def grouped_topk(scores, k, n_group, topk_group, renormalize, scoring="sigmoid", bias=None,
                 routed_scaling_factor=1.0):
    if _use_cuda(scores):
        return cuda_ops().grouped_topk(scores, k, n_group, topk_group, renormalize, scoring, bias,
                                       routed_scaling_factor)
    return ref.grouped_topk(scores, k, n_group, topk_group, renormalize, scoring, bias, routed_scaling_factor)


def fused_moe(x, w13, w2, topk_w, topk_ids, act="silu", expert_map=None):
    if _use_cuda(x):
        return cuda_ops().fused_moe(x, w13, w2, topk_w, topk_ids, act, expert_map)
    return ref.fused_moe(x, w13, w2, topk_w, topk_ids, act, expert_map)


def block_fp8_linear(x, wq, w_scale, block=128, bias=None):
    if _use_cuda(x):
        return cuda_ops().block_fp8_linear(x, wq, w_scale, block, bias)
    return ref.block_fp8_linear(x, wq, w_scale, block, bias)


# -- quantisers, q/k-norm + RoPE, cache maintenance: CUDA kernels (csrc/quant.cu, qknorm_rope.cu, cache_ops.cu), fp32
#    PyTorch reference on the CPU
def scaled_fp8_quant(x, scale=None, per_token: bool = False):
    if _use_cuda(x) and x.dtype in (torch.bfloat16, torch.float16) and x.shape[-1] % 8 == 0:
        return cuda_ops().scaled_fp8_quant(x, scale, per_token)
    return ref.scaled_fp8_quant(x, scale, per_token)


def scaled_int8_quant(x, scale=None):
    if _use_cuda(x) and x.dtype in (torch.bfloat16, torch.float16) and x.shape[-1] % 8 == 0:
        return cuda_ops().scaled_int8_quant(x, scale)
    return ref.scaled_int8_quant(x, scale)


def per_token_group_quant_fp8(x, group: int = 128, col_major: bool = False, ue8m0: bool = False):
    if _use_cuda(x) and x.dtype in (torch.bfloat16, torch.float16):
        return cuda_ops().per_token_group_quant_fp8(x, group, col_major, ue8m0)
    return ref.per_token_group_quant_fp8(x.reshape(-1, x.shape[-1]), group)


def rms_norm_quant(x, weight, eps, *, residual=None, scale=None, mode: str = "static", group: int = 128,
                   gemma: bool = False):
    """(residual add +) RMSNorm + fp8 quant in one pass; returns (q, scale) and updates ``residual`` in place."""
    if _use_cuda(x) and x.dtype in (torch.bfloat16, torch.float16) and x.shape[-1] % 8 == 0:
        return cuda_ops().rms_norm_quant(x, weight, eps, residual=residual, scale=scale, mode=mode, group=group,
                                         gemma=gemma)
    if residual is not None:
        y, _ = ref.fused_add_rms_norm(x, residual, weight, eps, gemma)
    else:
        y = ref.rms_norm(x, weight, eps, gemma)
    if mode == "group":
        return ref.per_token_group_quant_fp8(y.reshape(-1, y.shape[-1]), group)
    return ref.scaled_fp8_quant(y, scale, per_token=(mode == "token"))


def act_mul_quant(x, act: str = "silu", *, scale=None, mode: str = "static", group: int = 128):
    if _use_cuda(x) and x.dtype in (torch.bfloat16, torch.float16) and x.shape[-1] % 16 == 0:
        return cuda_ops().act_mul_quant(x, act, scale=scale, mode=mode, group=group)
    y = ref.silu_and_mul(x) if act == "silu" else ref.gelu_and_mul(x, "tanh" if "tanh" in act else "none")
    if mode == "group":
        return ref.per_token_group_quant_fp8(y.reshape(-1, y.shape[-1]), group)
    return ref.scaled_fp8_quant(y, scale, per_token=(mode == "token"))


def fused_qk_norm_rope(positions, q, k, q_weight, k_weight, cos_sin_cache, head_dim, eps, is_neox=True, **kw):
    if _use_cuda(q) and q.dtype in (torch.bfloat16, torch.float16) and cuda_ops().qknorm_rope_ok(
            head_dim, cos_sin_cache.shape[-1] if cos_sin_cache is not None else 0, is_neox):
        return cuda_ops().fused_qk_norm_rope(positions, q, k, q_weight, k_weight, cos_sin_cache, head_dim, eps, is_neox,
                                             **kw)
    q, k = ref.fused_qk_norm_rope(positions, q, k, q_weight, k_weight, cos_sin_cache, head_dim, eps, is_neox)
    if kw.get("kv_cache") is not None:
        v = kw["v"]
        Hkv = k.shape[-1] // head_dim
        reshape_and_cache(k.reshape(-1, Hkv, head_dim), v.reshape(-1, Hkv, head_dim), kw["kv_cache"],
                          kw["slot_mapping"], vt=kw.get("vt", False))
    return q, k


def swap_blocks(src, dst, mapping) -> None:
    if (src.is_cuda or dst.is_cuda) and not envs.VLLM_B200_FORCE_REF_OPS and (src.is_cuda or src.is_pinned()) \
            and (dst.is_cuda or dst.is_pinned()) and (src[0].numel() * src.element_size()) % 16 == 0 \
            and src.is_contiguous() and dst.is_contiguous() and src.dtype == dst.dtype:
        return cuda_ops().swap_blocks(src, dst, mapping)
    return ref.swap_blocks(src, dst, mapping if not isinstance(mapping, torch.Tensor) else mapping.tolist())


def convert_fp8(x, scale: float = 1.0, to_fp8: bool = True):
    if _use_cuda(x) and (x.dtype in (torch.bfloat16, torch.float16) or not to_fp8):
        return cuda_ops().convert_fp8(x, scale, to_fp8)
    return ref.convert_fp8(x, scale, to_fp8)


def scaled_fp4_quant(x, global_scale=None, mx: bool = False):
    if _use_cuda(x) and x.dtype in (torch.bfloat16, torch.float16) and x.shape[-1] % 32 == 0:
        return cuda_ops().scaled_fp4_quant(x, global_scale, mx)
    raise NotImplementedError("fp4 activation quantisation runs on the device only")


def scaled_mm(a, b, scale_a, scale_b, out_dtype, bias=None):
    """W8A8: (a · bᵀ) * scale_a * scale_b (+bias).  fp8 operands run on the tcgen05 f8f6f4 kernel."""
    if _use_cuda(a):
        return cuda_ops().scaled_mm(a, b, scale_a, scale_b, out_dtype, bias)
    return ref.scaled_mm(a, b, scale_a, scale_b, out_dtype, bias)


# -- ops served by the PyTorch implementations on every device (elementwise / bookkeeping)
mul_and_silu = ref.mul_and_silu
fatrelu_and_mul = ref.fatrelu_and_mul
swigluoai_and_mul = ref.swigluoai_and_mul
fp4_dequant = ref.fp4_dequant
