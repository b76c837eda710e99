"""ctypes bindings to ``vllm_b200/_C.so`` (built by ``vllm_b200/csrc/build.py``).

Every wrapper passes raw device pointers + the current CUDA stream, so the calls are
CUDA-graph capturable.  A missing / stale library raises at import: on a GPU box the
kernels are the only path (no silent PyTorch fallback).
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

from vllm_b200 import envs

_LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "_C.so")


def _load() -> ctypes.CDLL:
    if not os.path.exists(_LIB_PATH):
        from vllm_b200.csrc.build import build
        build()
    return ctypes.CDLL(_LIB_PATH)


_lib = _load()
_vp, _i, _i64, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float

_DT = {torch.bfloat16: 0, torch.float16: 1, torch.float32: 2}
_ACT = {"silu": 0, "gelu": 1, "gelu_tanh": 2, "gelu_pytorch_tanh": 2, "relu": 3, "gelu_new": 4, "gelu_fast": 5,
        "quick_gelu": 6, "relu2": 7}


def _sig(name: str, *argtypes):
    fn = getattr(_lib, name)
    fn.argtypes = list(argtypes)
    fn.restype = ctypes.c_int
    return fn


_rms_norm = _sig("vb_rms_norm", _vp, _vp, _vp, _vp, _i, _i, _i64, _f, _i, _i, _vp)
_layer_norm = _sig("vb_layer_norm", _vp, _vp, _vp, _vp, _i, _i, _f, _i, _vp)
_act_and_mul = _sig("vb_act_and_mul", _vp, _vp, _i64, _i, _i, _i, _vp)
_activation = _sig("vb_activation", _vp, _vp, _i64, _i, _i, _vp)
_rope_and_cache = _sig("vb_rope_and_cache", _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i64, _i64, _i64, _vp, _vp,
                       _i, _i, _i, _f, _f, _i, _vp)
_concat_mla = _sig("vb_concat_and_cache_mla", _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i64, _i, _i, _vp)
_paged_attention = _sig("vb_paged_attention", _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i,
                        _i64, _i64, _i, _i, _i, _f, _f, _f, _i, _i, _i, _vp)
_paged_attention_ex = _sig("vb_paged_attention_ex", _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i,
                           _i64, _i64, _i, _i, _i, _f, _f, _f, _i, _i, _i, _vp, _vp, _i, _vp)
_fmha_tc = _sig("vb_fmha_tc2", _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i64, _i64, _i,
                _i, _i, _f, _i, _i, _i, _vp)
_fmha_tc_decode = _sig("vb_fmha_tc_decode", _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i64, _i,
                       _i, _i, _f, _i, _vp)
_merge_states = _sig("vb_merge_attn_states", _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp)
_argmax = _sig("vb_argmax", _vp, _i64, _i, _i, _vp, _i, _vp)
_fused_sample = _sig("vb_fused_sample", _vp, _i64, _i, _i, _vp, _vp, _vp, _vp, _i, _vp)
_topk_topp = _sig("vb_topk_topp", _vp, _i, _i, _vp, _vp, _vp, _vp)
_row_lse = _sig("vb_row_lse", _vp, _i64, _i, _i, _vp, _i, _vp)
_gemm_plan = _sig("vb_gemm_tc_plan", _i, _i, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i),
                  ctypes.POINTER(_i64), ctypes.POINTER(_i))
_gemm_tc = _sig("vb_gemm_tc2", _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i, _i, _i, _vp, _vp, _i, _i, _vp)


def _st() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


LAUNCHES = 0  # number of our kernels launched through these wrappers (bench.py reports the delta)


def _check(rc: int, name: str, n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n
    if rc != 0:
        raise RuntimeError(f"vllm_b200 CUDA op {name} failed (code {rc}"
                           f"{': ' + torch.cuda.get_device_name() if rc > 0 else ''})")


_NUM_SMS: dict[int, int] = {}


def num_sms() -> int:
    d = torch.cuda.current_device()
    if d not in _NUM_SMS:
        _NUM_SMS[d] = torch.cuda.get_device_properties(d).multi_processor_count
    return _NUM_SMS[d]


# ------------------------------------------------------------------ norms / act
def rms_norm(x, weight, eps, gemma=False):
    H = x.shape[-1]
    x2 = x.reshape(-1, H)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    out = torch.empty((x2.shape[0], H), dtype=x.dtype, device=x.device)
    _check(_rms_norm(out.data_ptr(), x2.data_ptr(), None, weight.data_ptr(), x2.shape[0], H, x2.stride(0), eps,
                     int(gemma), _DT[x.dtype], _st()), "rms_norm")
    return out.view(x.shape)


def fused_add_rms_norm(x, residual, weight, eps, gemma=False):
    """In place: residual += x; x = norm(residual). Returns (x, residual)."""
    H = x.shape[-1]
    assert x.is_contiguous() and residual.is_contiguous()
    T = x.numel() // H
    _check(_rms_norm(x.data_ptr(), x.data_ptr(), residual.data_ptr(), weight.data_ptr(), T, H, H, eps, int(gemma),
                     _DT[x.dtype], _st()), "fused_add_rms_norm")
    return x, residual


def layer_norm(x, weight, bias, eps):
    H = x.shape[-1]
    xc = x.contiguous()
    out = torch.empty_like(xc)
    _check(_layer_norm(out.data_ptr(), xc.data_ptr(), weight.data_ptr(), _p(bias), xc.numel() // H, H, eps,
                       _DT[x.dtype], _st()), "layer_norm")
    return out


def act_and_mul(x, act: str):
    d = x.shape[-1] // 2
    xc = x.contiguous()
    out = torch.empty(x.shape[:-1] + (d,), dtype=x.dtype, device=x.device)
    _check(_act_and_mul(out.data_ptr(), xc.data_ptr(), xc.numel() // (2 * d), d, _ACT[act], _DT[x.dtype], _st()),
           "act_and_mul")
    return out


_silu_mul_il = _sig("vb_silu_mul_interleaved", _vp, _vp, _i64, _i, _i, _vp)


def silu_and_mul_interleaved(x):
    """x[..., 2i] = gate_i, x[..., 2i+1] = up_i → silu(gate) * up."""
    d = x.shape[-1] // 2
    xc = x.contiguous()
    out = torch.empty(x.shape[:-1] + (d,), dtype=x.dtype, device=x.device)
    _check(_silu_mul_il(out.data_ptr(), xc.data_ptr(), xc.numel() // (2 * d), d, _DT[x.dtype], _st()), "silu_mul_interleaved")
    return out


def activation(x, name: str):
    xc = x.contiguous()
    out = torch.empty_like(xc)
    _check(_activation(out.data_ptr(), xc.data_ptr(), xc.numel(), _ACT[name], _DT[x.dtype], _st()), "activation")
    return out


# ------------------------------------------------------------------ rope / cache
def _kv_fp8(kv_cache) -> int:
    return int(kv_cache.dtype in (torch.float8_e4m3fn, torch.uint8))


def rope_and_cache(positions, q, k, v, cos_sin_cache, head_dim, is_neox, kv_cache, slot_mapping,
                   k_scale: float = 1.0, v_scale: float = 1.0, vt: bool = False):
    T = q.shape[0]
    Hq, Hkv = q.shape[-1] // head_dim, k.shape[-1] // head_dim
    assert q.stride(-1) == 1 and k.stride(-1) == 1 and v.stride(-1) == 1
    rot = cos_sin_cache.shape[-1]
    flags = _kv_fp8(kv_cache) | (2 if vt else 0)
    _check(_rope_and_cache(positions.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr(), cos_sin_cache.data_ptr(),
                           rot, T, Hq, Hkv, head_dim, q.stride(0), k.stride(0), v.stride(0), kv_cache.data_ptr(),
                           slot_mapping.data_ptr(), kv_cache.shape[3], int(is_neox), flags, k_scale,
                           v_scale, _DT[q.dtype], _st()), "rope_and_cache")
    return q, k


def rotary_embedding(positions, q, k, cos_sin_cache, head_dim, is_neox=True):
    T = q.shape[0]
    Hq = q.shape[-1] // head_dim
    Hkv = k.shape[-1] // head_dim if k is not None else 0
    if q.stride(-1) != 1:
        q = q.contiguous()
    rot = cos_sin_cache.shape[-1]
    _check(_rope_and_cache(positions.data_ptr(), q.data_ptr(), _p(k), None, cos_sin_cache.data_ptr(), rot, T, Hq,
                           Hkv, head_dim, q.stride(0), k.stride(0) if k is not None else 0, 0, None, None, 1,
                           int(is_neox), 0, 1.0, 1.0, _DT[q.dtype], _st()), "rotary_embedding")
    return q, k


def reshape_and_cache(key, value, kv_cache, slot_mapping, k_scale: float = 1.0, v_scale: float = 1.0,
                      vt: bool = False):
    T, Hkv, D = key.shape
    _check(_rope_and_cache(None, None, key.data_ptr(), value.data_ptr(), None, 0, T, 0, Hkv, D, 0, key.stride(0),
                           value.stride(0), kv_cache.data_ptr(), slot_mapping.data_ptr(), kv_cache.shape[3], 1,
                           _kv_fp8(kv_cache) | (2 if vt else 0), k_scale, v_scale, _DT[key.dtype], _st()),
           "reshape_and_cache")


def concat_and_cache_mla(kv_c, k_pe, kv_cache, slot_mapping):
    T, R = kv_c.shape
    P = k_pe.shape[-1]
    _check(_concat_mla(kv_c.data_ptr(), k_pe.data_ptr(), kv_cache.data_ptr(), slot_mapping.data_ptr(), T, R, P,
                       kv_c.stride(0), k_pe.stride(0), kv_cache.shape[1], _DT[kv_c.dtype], _st()),
           "concat_and_cache_mla")


# ------------------------------------------------------------------ attention
class AttnWorkspace:
    """Split-KV partials, sized once for (max_batch, Hq, max_splits, D)."""

    def __init__(self, max_tokens: int, Hq: int, D: int, max_splits: int, device) -> None:
        self.max_splits = max_splits
        self.part_out = torch.empty(max_tokens * Hq * max_splits * D, dtype=torch.float32, device=device)
        self.part_lse = torch.empty(max_tokens * Hq * max_splits, dtype=torch.float32, device=device)


_attn_ws: dict[tuple, AttnWorkspace] = {}


def _get_attn_ws(B: int, Hq: int, D: int, splits: int, device) -> AttnWorkspace:
    key = (device.index, Hq, D)
    ws = _attn_ws.get(key)
    need = B * Hq * splits * D
    if ws is None or ws.part_out.numel() < need:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("attention workspace must be allocated before CUDA-graph capture")
        ws = AttnWorkspace(max(B, 64), Hq, D, max(splits, 16), device)
        if ws.part_out.numel() < need:
            ws = AttnWorkspace(B, Hq, D, splits, device)
        _attn_ws[key] = ws
    return ws


def batch_invariant() -> bool:
    """``VLLM_BATCH_INVARIANT=1``: every dispatch decision that changes a row's floating-point reduction order is
    made from that row's own shape only — never from how many other rows share the launch.  Concretely: no split-K in
    the decode GEMMs (the split factor is otherwise chosen per (N, K) but switches off above 256 rows), no split-KV in
    decode attention (the split count depends on batch × heads), decode-only batches run the same varlen attention
    kernel as mixed batches, collectives reduce in rank order at every size (no in-switch multimem reduction, no
    Lamport fused GEMM + all-reduce), and the MoE combine does not use ``red.add`` epilogues (atomic arrival order).
    Reference: vllm/model_executor/layers/batch_invariant.py (``VLLM_BATCH_INVARIANT`` swaps in fixed-reduction-order
    matmul / softmax / rms-norm kernels, fixes the attention split size and pins NCCL to one algorithm :1105-1180),
    tests/v1/determinism/test_batch_invariance.py."""
    return bool(envs.VLLM_BATCH_INVARIANT)


def decode_splits(B: int, Hkv: int, max_seq_len: int, block_size: int) -> int:
    if batch_invariant():
        return 1
    forced = envs.VLLM_B200_ATTN_SPLITS
    if forced > 0:
        return forced
    ctas = max(B * Hkv, 1)
    target = 2 * num_sms()
    s = max(1, min(16, (target + ctas - 1) // ctas))
    pages = max(1, (max_seq_len + block_size - 1) // block_size)
    # a split costs ~10 us (workspace round trip + reduce kernel): measured on B200 (profiles/README.md §3) it only
    # pays once every split still streams >= ~2k tokens — at 1k-token contexts one CTA per (sequence, kv head) wins
    # even when that leaves SMs idle (14.5 us vs 25.8 us for 64 seqs x 1 kv head)
    s = min(s, max(1, max_seq_len // 2048))
    return max(1, min(s, pages))


def paged_attention(q, kv_cache, md, scale, sliding_window=None, logits_soft_cap=0.0, causal=True, out=None,
                    kv_scale: float = 1.0, return_lse: bool = False, alibi_slopes=None, sinks=None, chunk_size: int = 0):
    T, Hq, D = q.shape
    Hkv, BS = kv_cache.shape[2], kv_cache.shape[3]
    if out is None:
        # (bucket padding rows are written by no CTA: keep them finite for the row-wise layers that follow)
        out = (torch.zeros if md.flat else torch.empty)((T, Hq, D), dtype=q.dtype, device=q.device)
    lse = torch.empty((T, Hq), dtype=torch.float32, device=q.device) if return_lse else None
    bt = md.block_table
    if md.kv_vt:
        # tcgen05 kernel over the tensor-core cache layout
        assert D == 128 and BS == 64 and not logits_soft_cap and alibi_slopes is None and sinks is None and not chunk_size, \
            "tensor-core attention: D=128, 64-token pages, no soft-cap / ALiBi / sinks / chunked-local mask"
        # batch-invariant mode: a decode row takes the varlen kernel it would take inside a mixed batch
        decode = md.decode_only and not (batch_invariant() and md.query_start_loc is not None)
        splits = (md.num_splits or decode_splits(md.num_reqs, Hkv, md.max_seq_len, 128)) if decode else 1
        ws = _get_attn_ws(md.num_reqs, Hq, D, splits, q.device) if splits > 1 else None
        G = Hq // Hkv
        if decode and G <= 16 and (G & (G - 1)) == 0 and causal:
            # swap-AB decode kernel: KV tokens on the MMA M axis
            rc = _fmha_tc_decode(q.data_ptr(), kv_cache.data_ptr(), kv_cache.shape[0], bt.data_ptr(),
                                 md.seq_lens.data_ptr(), out.data_ptr(), _p(ws.part_out) if ws else None,
                                 _p(ws.part_lse) if ws else None, _p(lse), md.num_reqs, Hq, Hkv, q.stride(0),
                                 out.stride(0), bt.stride(0), splits, sliding_window or 0, scale, _DT[q.dtype],
                                 _st())
            _check(rc, f"fmha_tc_decode(G={G})", 2 if splits > 1 else 1)
            return (out, lse) if return_lse else out
        rc = _fmha_tc(q.data_ptr(), kv_cache.data_ptr(), kv_cache.shape[0], bt.data_ptr(), md.seq_lens.data_ptr(),
                      None if decode else md.query_start_loc.data_ptr(), out.data_ptr(),
                      _p(ws.part_out) if ws else None, _p(ws.part_lse) if ws else None, _p(lse), md.num_reqs, T,
                      md.max_query_len, Hq, Hkv, q.stride(0), out.stride(0), bt.stride(0), splits,
                      sliding_window or 0, scale, int(causal), int(md.flat and not decode), _DT[q.dtype], _st())
        _check(rc, f"fmha_tc(G={Hq // Hkv})", 2 if splits > 1 else 1)
        return (out, lse) if return_lse else out
    if alibi_slopes is not None:
        alibi_slopes = alibi_slopes.to(device=q.device, dtype=torch.float32).contiguous()
    if sinks is not None:
        sinks = sinks.to(device=q.device, dtype=torch.float32).contiguous()
    n_launch = 1
    if md.decode_only and not (batch_invariant() and md.query_start_loc is not None):
        splits = md.num_splits or decode_splits(md.num_reqs, Hkv, md.max_seq_len, BS)
        n_launch = 2 if splits > 1 else 1
        ws = _get_attn_ws(md.num_reqs, Hq, D, splits, q.device) if splits > 1 else None
        rc = _paged_attention_ex(q.data_ptr(), kv_cache.data_ptr(), bt.data_ptr(), md.seq_lens.data_ptr(), None,
                                 out.data_ptr(), _p(ws.part_out) if ws else None, _p(ws.part_lse) if ws else None,
                                 _p(lse), md.num_reqs, 1, Hq, Hkv, D, BS, q.stride(0), out.stride(0), bt.stride(0),
                                 splits, sliding_window or 0, scale, logits_soft_cap, kv_scale, int(causal),
                                 _kv_fp8(kv_cache), _DT[q.dtype], _p(alibi_slopes), _p(sinks), int(chunk_size or 0), _st())
    else:
        rc = _paged_attention_ex(q.data_ptr(), kv_cache.data_ptr(), bt.data_ptr(), md.seq_lens.data_ptr(),
                                 md.query_start_loc.data_ptr(), out.data_ptr(), None, None, _p(lse), md.num_reqs,
                                 md.max_query_len, Hq, Hkv, D, BS, q.stride(0), out.stride(0), bt.stride(0), 1,
                                 sliding_window or 0, scale, logits_soft_cap, kv_scale, int(causal),
                                 _kv_fp8(kv_cache), _DT[q.dtype], _p(alibi_slopes), _p(sinks), int(chunk_size or 0), _st())
    _check(rc, f"paged_attention(D={D}, BS={BS}, G={Hq // Hkv})", n_launch)
    return (out, lse) if return_lse else out


def merge_attn_states(o1, lse1, o2, lse2):
    T, H, D = o1.shape
    out = torch.empty_like(o1)
    lse = torch.empty_like(lse1)
    _check(_merge_states(out.data_ptr(), lse.data_ptr(), o1.data_ptr(), lse1.data_ptr(), o2.data_ptr(),
                         lse2.data_ptr(), T * H, D, _DT[o1.dtype], _st()), "merge_attn_states")
    return out, lse


_mla_decode = _sig("vb_mla_decode", _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i64, _i64, _i64, _i64,
                   _i, _i, _f, _i, _vp)
_mla_ws: dict[int, tuple] = {}


def mla_decode_splits(B: int, H: int, max_seq_len: int) -> int:
    if batch_invariant():
        return 1
    forced = envs.VLLM_B200_ATTN_SPLITS
    if forced > 0:
        return forced
    groups = (H + 31) // 32 if H > 16 else 1
    ctas = max(B * groups, 1)
    s = max(1, min(16, num_sms() // ctas))           # one CTA per SM (170 KB smem): stay inside ONE wave
    return max(1, min(s, max_seq_len // 512))       # every split should still stream >= 8 tiles


def mla_decode_tc(q_nope_abs, q_pe, kv_cache, block_table, seq_lens, scale, max_seq_len: int, num_splits: int = 0,
                  return_lse: bool = False):
    """tcgen05 MLA decode (csrc/mla_decode.cu): q_nope_abs [B, H, 512], q_pe [B, H, 64], cache [blocks, BS, 576]."""
    B, H, R = q_nope_abs.shape
    assert R == 512 and q_pe.shape[-1] == 64 and kv_cache.shape[-1] == 576 and q_nope_abs.stride(-1) == 1 and q_pe.stride(-1) == 1
    dev = q_nope_abs.device
    splits = num_splits or mla_decode_splits(B, H, max_seq_len)
    out = torch.empty((B, H, R), dtype=q_nope_abs.dtype, device=dev)
    lse = torch.empty((B, H), dtype=torch.float32, device=dev) if return_lse else None
    po = pl = None
    if splits > 1:
        key = dev.index or 0
        need = B * H * splits
        ws = _mla_ws.get(key)
        if ws is None or ws[1].numel() < need:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("MLA split-KV workspace must be allocated before CUDA-graph capture")
            n = max(need, 64 * 128 * 4)
            ws = _mla_ws[key] = (torch.empty(n * R, dtype=torch.float32, device=dev),
                                 torch.empty(n, dtype=torch.float32, device=dev))
        po, pl = ws
    _check(_mla_decode(q_nope_abs.data_ptr(), q_pe.data_ptr(), kv_cache.data_ptr(), kv_cache.shape[0],
                       block_table.data_ptr(), seq_lens.data_ptr(), out.data_ptr(), _p(po), _p(pl), _p(lse), B, H,
                       kv_cache.shape[1], q_nope_abs.stride(0), q_nope_abs.stride(1), q_pe.stride(0), q_pe.stride(1),
                       out.stride(0), block_table.stride(0), splits, scale, _DT[q_nope_abs.dtype], _st()),
           f"mla_decode(B={B},H={H})", 2 if splits > 1 else 1)
    return (out, lse) if return_lse else out


def _mla_tc_ok(q_nope_abs, q_pe, kv_cache) -> bool:
    return (q_nope_abs.dtype in (torch.bfloat16, torch.float16) and kv_cache.dtype == q_nope_abs.dtype
            and q_nope_abs.shape[-1] == 512 and q_pe.shape[-1] == 64 and kv_cache.dim() == 3 and kv_cache.shape[-1] == 576
            and kv_cache.shape[1] in (16, 32, 64, 128) and q_nope_abs.stride(-1) == 1 and q_pe.stride(-1) == 1
            and (q_nope_abs.stride(0) * 2) % 16 == 0 and (q_nope_abs.stride(1) * 2) % 16 == 0
            and (q_pe.stride(0) * 2) % 16 == 0 and (q_pe.stride(1) * 2) % 16 == 0 and envs.VLLM_B200_ATTN != "simt")


def mla_attention(q_nope_abs, q_pe, kv_cache, md, scale, kv_lora_rank):
    """MLA as MQA over the latent cache (K = 576-d latent+rope, V = first 512 dims of the same rows).

    Decode batches run on the tcgen05 MLA kernel (csrc/mla_decode.cu: the latent tile is shared by all heads of the
    CTA and by both matrix products, split-KV for long contexts; no host synchronisation, graph capturable).
    Shapes it does not cover (other latent widths) use dense batched GEMMs over the gathered pages; mixed / prefill
    batches use the variable-length reference loop (eager only)."""
    from vllm_b200.ops import ref
    if md.decode_only:
        B, H, R = q_nope_abs.shape
        if R == kv_lora_rank and _mla_tc_ok(q_nope_abs, q_pe, kv_cache):
            return mla_decode_tc(q_nope_abs, q_pe, kv_cache, md.block_table, md.seq_lens, scale, md.max_seq_len,
                                 md.num_splits or 0)
        bs = kv_cache.shape[1]
        nbk = max((md.max_seq_len + bs - 1) // bs, 1)
        lat = kv_cache[md.block_table[:B, :nbk].long()].reshape(B, nbk * bs, -1)        # [B, S, 576]
        kc, kp = lat[..., :kv_lora_rank], lat[..., kv_lora_rank:]
        att = (torch.bmm(q_nope_abs, kc.transpose(1, 2)) + torch.bmm(q_pe, kp.transpose(1, 2))).float() * scale
        pos = torch.arange(nbk * bs, device=att.device)[None, None, :]
        att = att.masked_fill(pos >= md.seq_lens[:B, None, None], float("-inf"))
        p = torch.softmax(att, dim=-1).to(q_nope_abs.dtype)
        return torch.bmm(p, kc)
    sl = md.seq_lens_cpu if md.seq_lens_cpu is not None else md.seq_lens.cpu()
    qsl = md.query_start_loc_cpu if md.query_start_loc_cpu is not None else md.query_start_loc.cpu()
    return ref.mla_decode_attention(q_nope_abs, q_pe, kv_cache, md.block_table, sl, qsl, scale, kv_lora_rank)


# ------------------------------------------------------------------ sampling
def argmax(logits: torch.Tensor) -> torch.Tensor:
    B, V = logits.shape
    out = torch.empty(B, dtype=torch.int64, device=logits.device)
    _check(_argmax(logits.data_ptr(), logits.stride(0), B, V, out.data_ptr(), _DT[logits.dtype], _st()), "argmax")
    return out


_step_buf: dict[int, torch.Tensor] = {}


def fused_sample(logits, temperature, seeds, step, out: Optional[torch.Tensor] = None):
    """``step``: per-row int32 [B] noise counters (a python int is broadcast)."""
    B, V = logits.shape
    if out is None:
        out = torch.empty(B, dtype=torch.int64, device=logits.device)
    if not isinstance(step, torch.Tensor):
        step = torch.full((B,), int(step) & 0x7FFFFFFF, dtype=torch.int32, device=logits.device)
    _check(_fused_sample(logits.data_ptr(), logits.stride(0), B, V, temperature.data_ptr(), seeds.data_ptr(),
                         step.data_ptr(), out.data_ptr(), _DT[logits.dtype], _st()), "fused_sample")
    return out


_fused_sample_cand = _sig("vb_fused_sample_cand", _vp, _i64, _i, _i, _vp, _vp, _vp, _i, _vp, _i, _vp)
_select_cand = _sig("vb_select_candidate", _vp, _i, _vp, _i, _vp)


def fused_sample_candidates(logits_local, temperature, seeds, step, idx_off: int, v_valid: int) -> torch.Tensor:
    """Per-row best (value, global id) of this rank's vocab shard as a [B, 8] bf16-viewable 16-byte record."""
    B = logits_local.shape[0]
    cand = torch.empty((B, 4), dtype=torch.float32, device=logits_local.device)
    _check(_fused_sample_cand(logits_local.data_ptr(), logits_local.stride(0), B, max(v_valid, 0), temperature.data_ptr(),
                              seeds.data_ptr(), step.data_ptr(), idx_off, cand.data_ptr(), _DT[logits_local.dtype], _st()),
           "fused_sample_cand")
    return cand


def select_candidates(cands: torch.Tensor, W: int, out: torch.Tensor) -> torch.Tensor:
    """cands: [B, W, 4] fp32 records gathered from the W ranks → winning token id per row."""
    B = out.shape[0]
    _check(_select_cand(cands.data_ptr(), W, out.data_ptr(), B, _st()), "select_candidate")
    return out


def topk_topp_mask_(logits: torch.Tensor, top_k, top_p, min_p=None) -> torch.Tensor:
    assert logits.dtype == torch.float32 and logits.is_contiguous()
    B, V = logits.shape
    _check(_topk_topp(logits.data_ptr(), B, V, _p(top_k), _p(top_p), _p(min_p), _st()), "topk_topp")
    return logits


def row_lse(logits: torch.Tensor) -> torch.Tensor:
    B, V = logits.shape
    out = torch.empty(B, dtype=torch.float32, device=logits.device)
    _check(_row_lse(logits.data_ptr(), logits.stride(0), B, V, out.data_ptr(), _DT[logits.dtype], _st()), "row_lse")
    return out


# ------------------------------------------------------------------ GEMM
class _GemmWorkspace:
    def __init__(self) -> None:
        self.ws: Optional[torch.Tensor] = None
        self.counters: Optional[torch.Tensor] = None


_gemm_ws: dict[int, _GemmWorkspace] = {}
_plan_cache: dict[tuple, tuple] = {}


def _plan(M: int, N: int, K: int):
    key = (M, N, K)
    p = _plan_cache.get(key)
    if p is None:
        swap, bn, splits, tiles = _i(), _i(), _i(), _i()
        wsb = _i64()
        _gemm_plan(M, N, K, num_sms(), ctypes.byref(swap), ctypes.byref(bn), ctypes.byref(splits), ctypes.byref(wsb),
                   ctypes.byref(tiles))
        p = (swap.value, bn.value, 1 if batch_invariant() else splits.value, wsb.value, tiles.value)
        _plan_cache[key] = p
    return p


def gemm_tc(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
            out: Optional[torch.Tensor] = None, act: bool = False) -> torch.Tensor:
    """Y = X W^T on our tcgen05 kernel. x: [M, K] (row stride free), w: [N, K] contiguous.  ``act`` (M <= 256): W's
    rows interleave (gate_i, up_i) and the epilogue emits silu(gate) * up → [M, N/2]."""
    M, K = x.shape
    N = w.shape[0]
    swap, bn, splits, wsb, tiles = _plan(M, N, K)
    dev = x.device.index or 0
    g = _gemm_ws.setdefault(dev, _GemmWorkspace())
    if splits > 1:
        if g.ws is None or g.ws.numel() * 4 < wsb or g.counters.numel() < tiles:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("GEMM split-K workspace must be allocated before CUDA-graph capture")
            g.ws = torch.empty(max(wsb // 4, 16 << 20), dtype=torch.float32, device=x.device)
            g.counters = torch.zeros(max(tiles, 4096), dtype=torch.int32, device=x.device)
    y = out if out is not None else torch.empty((M, N // 2 if act else N), dtype=x.dtype, device=x.device)
    _check(_gemm_tc(x.data_ptr(), w.data_ptr(), _p(bias), y.data_ptr(), M, N, K, x.stride(0), swap, bn, splits,
                    _p(g.ws), _p(g.counters), int(act), _DT[x.dtype], _st()), f"gemm_tc(M={M},N={N},K={K})")
    return y


_gemm_sk_plan = _sig("vb_gemm_sk_plan", _i, _i, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i),
                     ctypes.POINTER(_i64), ctypes.POINTER(_i))
_gemm_sk = _sig("vb_gemm_sk", _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i, _i, _i, _vp, _vp, _i, _i, _vp)
_sk_plan_cache: dict[tuple, tuple] = {}
_sk_ws: dict[int, tuple] = {}


def _sk_workspace(dev: torch.device) -> tuple:
    """Stream-K partial-tile workspace + tile counters, sized once for the largest plan (SMs x 2 slots x 256 x 128 fp32)."""
    key = dev.index or 0
    ws = _sk_ws.get(key)
    if ws is None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("stream-K GEMM workspace must be allocated before CUDA-graph capture")
        ws = _sk_ws[key] = (torch.empty(num_sms() * 2 * 256 * 128, dtype=torch.float32, device=dev),
                            torch.zeros(16384, dtype=torch.int32, device=dev))
    return ws


def gemm_sk(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
            act: bool = False) -> torch.Tensor:
    """Y = X W^T on the persistent stream-K tcgen05 kernel (csrc/gemm_sk.cu), M <= 256.  ``act``: W's rows are
    interleaved (gate_i, up_i) pairs and the epilogue emits silu(gate) * up ([M, N/2])."""
    M, K = x.shape
    N = w.shape[0]
    key = (M, N, K)
    plan = _sk_plan_cache.get(key)
    if plan is None:
        bn, grid, per, tiles = _i(), _i(), _i(), _i()
        wsb = _i64()
        rc = _gemm_sk_plan(M, N, K, num_sms(), ctypes.byref(bn), ctypes.byref(grid), ctypes.byref(per), ctypes.byref(wsb),
                           ctypes.byref(tiles))
        assert rc == 0 and tiles.value <= 16384
        plan = _sk_plan_cache[key] = (bn.value, grid.value, per.value)
    bn, grid, per = plan
    ws, counters = _sk_workspace(x.device)
    y = out if out is not None else torch.empty((M, N // 2 if act else N), dtype=x.dtype, device=x.device)
    _check(_gemm_sk(x.data_ptr(), w.data_ptr(), _p(bias), y.data_ptr(), M, N, K, x.stride(0), bn, grid, per, ws.data_ptr(),
                    counters.data_ptr(), int(act), _DT[x.dtype], _st()), f"gemm_sk(M={M},N={N},K={K})")
    return y


def _tc_ok(x2: torch.Tensor, w: torch.Tensor, force: bool = False) -> bool:
    mode = envs.VLLM_B200_GEMM
    if mode == "cublas" and not force:
        return False
    K = x2.shape[1]
    ok = (x2.dtype in (torch.bfloat16, torch.float16) and w.dtype == x2.dtype and K % 64 == 0
          and x2.stride(1) == 1 and (x2.stride(0) * 2) % 16 == 0 and x2.data_ptr() % 16 == 0 and w.is_contiguous())
    if mode == "tcgen05" or force:
        return ok
    # auto: our kernel for the weight-streaming decode shapes, cuBLASLt for the rest
    return ok and x2.shape[0] <= 256


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    K = x.shape[-1]
    x2 = x.reshape(-1, K)
    if x2.shape[0] > 0 and _tc_ok(x2, weight):
        # decode shapes: persistent stream-K kernel (every SM streams weights all the time); VLLM_B200_GEMM_SK=0 selects
        # the tile-per-CTA / cluster split-K kernel for A/B measurements
        if x2.shape[0] <= 256 and envs.VLLM_B200_GEMM_SK and weight.shape[0] * (K // 64) >= 64:
            y = gemm_sk(x2, weight, bias)
        else:
            y = gemm_tc(x2, weight, bias)
        return y.view(x.shape[:-1] + (weight.shape[0],))
    return torch.nn.functional.linear(x, weight, bias)


# ------------------------------------------------------------------ symmetric-heap collectives
_comm_ar_rows = _sig("vb_comm_allreduce_rows", _vp, _vp, _vp, _vp, _i, _i, _vp, _i64, _i64, _vp, _vp, _f, _i, _i, _i,
                     _i, _i, _vp)
_comm_ag = _sig("vb_comm_all_gather", _vp, _vp, _vp, _vp, _i, _i, _vp, _i64, _i, _i, _i, _i, _vp)
_comm_bar = _sig("vb_comm_barrier", _vp, _vp, _vp, _vp, _i, _i, _vp)


def comm_allreduce_rows(c, x, in_off, out_off, residual, norm_w, eps, T, H, gemma, dtype) -> None:
    _check(_comm_ar_rows(c.local_ptr, c.peer_arr, c._mc(), c.state.data_ptr(), c.rank, c.world, _p(x), in_off,
                         out_off, _p(residual), _p(norm_w), eps, T, H, int(gemma), num_sms(), _DT[dtype], _st()),
           "comm_allreduce_rows")


_comm_ar_oneshot = _sig("vb_comm_allreduce_oneshot", _vp, _vp, _vp, _vp, _i, _i, _vp, _i64, _vp, _i, _i, _i, _i, _vp)


def comm_allreduce_oneshot(c, x, in_off, dst, T, H) -> None:
    """Sum over ranks of [T,H] (``x`` or what already sits in the staging region) straight into the private ``dst``."""
    _check(_comm_ar_oneshot(c.local_ptr, c.peer_arr, c._mc(), c.state.data_ptr(), c.rank, c.world, _p(x), in_off,
                            dst.data_ptr(), T, H, num_sms(), _DT[dst.dtype], _st()), "comm_allreduce_oneshot")


def comm_all_gather(c, x, out_off, rows, cols) -> None:
    _check(_comm_ag(c.local_ptr, c.peer_arr, c._mc(), c.state.data_ptr(), c.rank, c.world, x.data_ptr(), out_off,
                    rows, cols, num_sms(), _DT[x.dtype], _st()), "comm_all_gather")


_gemm_ar_plan = _sig("vb_gemm_ar_plan", _i, _i, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i64),
                     ctypes.POINTER(_i))
_gemm_ar = _sig("vb_gemm_ar", _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i, _i, _vp, _vp,
                _i64, _i64, _i, _i, _i64, _i64, _f, _i, _i, _vp)
_ar_plan_cache: dict[tuple, Optional[tuple]] = {}


def gemm_ar(c, x, w, residual, norm_w, out, eps, gemma, mode=0, oring_off=0, oring_stride=0) -> bool:
    """Fused RowParallel GEMM + all-reduce + residual + RMSNorm (csrc/gemm_ar.cu). False = shape not covered."""
    from vllm_b200.parallel.symm import LAM_T_CAP
    M, K = x.shape
    N = w.shape[0]
    key = (M, N, K)
    plan = _ar_plan_cache.get(key, 0)
    if plan == 0:
        bn, splits, tiles = _i(), _i(), _i()
        wsb = _i64()
        rc = _gemm_ar_plan(M, N, K, num_sms(), ctypes.byref(bn), ctypes.byref(splits), ctypes.byref(wsb),
                           ctypes.byref(tiles))
        plan = (bn.value, splits.value, wsb.value, tiles.value) if rc == 0 else None
        _ar_plan_cache[key] = plan
    if plan is None:
        return False
    bn, splits, wsb, tiles = plan
    dev = x.device.index or 0
    g = _gemm_ws.setdefault(dev, _GemmWorkspace())
    if g.ws is None or g.ws.numel() * 4 < wsb or g.counters.numel() < tiles:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("GEMM split-K workspace must be allocated before CUDA-graph capture")
        g.ws = torch.empty(max(wsb // 4, 16 << 20), dtype=torch.float32, device=x.device)
        g.counters = torch.zeros(max(tiles, 4096), dtype=torch.int32, device=x.device)
    _check(_gemm_ar(c.local_ptr, c.peer_arr, c._mc(), c.state.data_ptr(), c.rank, c.world, x.data_ptr(), w.data_ptr(),
                    residual.data_ptr(), norm_w.data_ptr(), out.data_ptr() if out is not None else None, M, N, K,
                    x.stride(0), bn, splits, g.ws.data_ptr(), g.counters.data_ptr(), c.off_lam, c.lam_stride, LAM_T_CAP, mode, oring_off, oring_stride, eps,
                    int(gemma),
                    _DT[x.dtype], _st()), f"gemm_ar(M={M},N={N},K={K})")
    return True


def gemm_ar_set_trace(t: Optional[torch.Tensor]) -> None:
    """Point the fused kernel's phase stamps ([n_ctas, 8] int64 globaltimer values) at ``t`` (None = off)."""
    fn = _lib.vb_gemm_ar_set_trace
    fn.argtypes, fn.restype = [_vp], None
    fn(t.data_ptr() if t is not None else None)


def comm_barrier(c) -> None:
    _check(_comm_bar(c.local_ptr, c.peer_arr, c._mc(), c.state.data_ptr(), c.rank, c.world, _st()), "comm_barrier")


# ------------------------------------------------------------------ MoE (csrc/moe.cu + grouped mode of gemm_tc.cu)
_moe_topk = _sig("vb_moe_topk_softmax", _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp)
_moe_align = _sig("vb_moe_align2", _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp)
_moe_finish = _sig("vb_moe_finish", _vp, _vp, _vp, _vp, _i, _i, _f, _i, _i, _vp)
_gemm_grouped_red = _sig("vb_gemm_tc_grouped_red", _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i64, _i, _i, _vp, _vp, _i, _i,
                         _i, _vp)
_comm_ar_rows2 = _sig("vb_comm_allreduce_rows2", _vp, _vp, _vp, _vp, _i, _i, _vp, _i64, _i64, _vp, _vp, _f, _i, _i, _i, _i, _i64,
                      _i, _vp)
_moe_combine = _sig("vb_moe_combine", _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp)
_moe_gather = _sig("vb_moe_gather", _vp, _vp, _vp, _i, _i, _i, _vp)
_gemm_grouped = _sig("vb_gemm_tc_grouped", _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp)


def topk_softmax(gating, k, renormalize, scale: float = 1.0):
    T, E = gating.shape
    if E > 512 or k > 32 or gating.dtype not in _DT or not gating.is_contiguous():
        from vllm_b200.ops import ref
        w, ids = ref.topk_softmax(gating, k, renormalize)
        return w * scale, ids
    w = torch.empty((T, k), dtype=torch.float32, device=gating.device)
    ids = torch.empty((T, k), dtype=torch.int32, device=gating.device)
    _check(_moe_topk(gating.data_ptr(), w.data_ptr(), ids.data_ptr(), T, E, k, int(renormalize), float(scale),
                     _DT[gating.dtype], _st()), "moe_topk_softmax")
    return w, ids


def moe_align(topk_ids: torch.Tensor, num_local_experts: int, expert_map: Optional[torch.Tensor] = None,
              with_pairs: bool = False):
    """(sorted_tok [R], inv [R], offs [E+1]) for R = T*k (token, slot) pairs — all on the device.  ``with_pairs`` also
    returns sorted_pair [R]: the pair index (token * k + slot) of every sorted row."""
    T, k = topk_ids.shape
    R = T * k
    dev = topk_ids.device
    sorted_tok = torch.empty(R, dtype=torch.int32, device=dev)
    inv = torch.empty(R, dtype=torch.int32, device=dev)
    offs = torch.empty(num_local_experts + 1, dtype=torch.int32, device=dev)
    pairs = torch.empty(R, dtype=torch.int32, device=dev) if with_pairs else None
    _check(_moe_align(topk_ids.data_ptr(), _p(expert_map), sorted_tok.data_ptr(), inv.data_ptr(), offs.data_ptr(), R, k,
                      num_local_experts, _p(pairs), _st()), "moe_align")
    return (sorted_tok, inv, offs, pairs) if with_pairs else (sorted_tok, inv, offs)


def gemm_tc_grouped_combine(xs, w, offs, bn, sorted_pair, pair_w, k: int, H: int, *, acc_ptr: int = 0, comm=None,
                            acc_off: int = 0, per: int = 1) -> None:
    """Grouped GEMM whose epilogue adds w(row) * y[row] into the fp32 accumulator row of its token: a local buffer
    (``acc_ptr``) on one GPU; with ``comm`` the buffer at ``acc_off`` of the symmetric heap OF THE RANK THAT OWNS THE
    TOKEN (token // per), reached with red.add over NVLink → top-k combine + reduce-scatter inside the GEMM."""
    R, K = xs.shape
    E, N, _ = w.shape
    if comm is not None:
        rc = _gemm_grouped_red(xs.data_ptr(), w.data_ptr(), None, offs.data_ptr(), R, N, K, E, bn, comm.local_ptr + acc_off,
                               comm.peer_arr, acc_off, comm.world, per, sorted_pair.data_ptr(), pair_w.data_ptr(), k, H,
                               _DT[xs.dtype], _st())
    else:
        rc = _gemm_grouped_red(xs.data_ptr(), w.data_ptr(), None, offs.data_ptr(), R, N, K, E, bn, acc_ptr, None, 0, 1, 1,
                               sorted_pair.data_ptr(), pair_w.data_ptr(), k, H, _DT[xs.dtype], _st())
    _check(rc, f"gemm_tc_grouped_combine(R={R},N={N},K={K},E={E})")


def comm_finish_rows_acc32(c, acc_off: int, out_off: int, residual, norm_w, eps, T, H, gemma, dtype) -> None:
    """Second half of the fused MoE tail under TP/EP: barrier → every rank reads (and re-zeroes) the fp32 sums of the
    rows it owns, adds its residual rows, normalises and multicasts the rows → barrier (csrc/comm.cu ar_rows_kernel)."""
    _check(_comm_ar_rows2(c.local_ptr, c.peer_arr, c._mc(), c.state.data_ptr(), c.rank, c.world, None, c.off_in, out_off,
                          _p(residual), _p(norm_w), eps, T, H, int(gemma), num_sms(), acc_off, _DT[dtype], _st()),
           "comm_finish_rows_acc32")


def moe_finish(acc_ptr: int, residual, norm_w, out, T: int, H: int, eps: float, gemma: bool = False) -> None:
    _check(_moe_finish(acc_ptr, _p(residual), _p(norm_w), out.data_ptr(), T, H, eps, int(gemma), _DT[out.dtype], _st()),
           "moe_finish")


def moe_gather(x: torch.Tensor, sorted_tok: torch.Tensor) -> torch.Tensor:
    R, H = sorted_tok.numel(), x.shape[1]
    xs = torch.empty((R, H), dtype=x.dtype, device=x.device)
    _check(_moe_gather(x.data_ptr(), sorted_tok.data_ptr(), xs.data_ptr(), R, H, _DT[x.dtype], _st()), "moe_gather")
    return xs


def moe_combine(y: torch.Tensor, inv: torch.Tensor, w: torch.Tensor, T: int) -> torch.Tensor:
    H = y.shape[1]
    k = inv.numel() // T
    out = torch.empty((T, H), dtype=y.dtype, device=y.device)
    _check(_moe_combine(y.data_ptr(), inv.data_ptr(), w.data_ptr(), out.data_ptr(), T, H, k, _DT[y.dtype], _st()),
           "moe_combine")
    return out


def gemm_tc_grouped(xs: torch.Tensor, w: torch.Tensor, offs: torch.Tensor, bn: int) -> torch.Tensor:
    """Y[r] = xs[r] · w[e(r)]^T, rows sorted by expert (``offs``). w: [E, N, K]."""
    R, K = xs.shape
    E, N, _ = w.shape
    y = torch.empty((R, N), dtype=xs.dtype, device=xs.device)
    _check(_gemm_grouped(xs.data_ptr(), w.data_ptr(), y.data_ptr(), offs.data_ptr(), R, N, K, E, bn, _DT[xs.dtype],
                         _st()), f"gemm_tc_grouped(R={R},N={N},K={K},E={E})")
    return y


def grouped_topk(scores, k, n_group, topk_group, renormalize, scoring="sigmoid", bias=None, routed_scaling_factor=1.0):
    from vllm_b200.ops import ref
    return ref.grouped_topk(scores, k, n_group, topk_group, renormalize, scoring, bias, routed_scaling_factor)


def fused_moe(x, w13, w2, topk_w, topk_ids, act="silu", expert_map=None):
    from vllm_b200.ops import moe
    return moe.fused_moe_cuda(x, w13, w2, topk_w, topk_ids, act, expert_map)


_gemm_fp8_plan = _sig("vb_gemm_fp8_plan", _i, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i))
_gemm_fp8 = _sig("vb_gemm_fp8", _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i64, _i, _i, _i, _i, _i, _i, _i, _vp)
_gemm_fp8_grouped = _sig("vb_gemm_fp8_grouped", _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i64, _i, _i, _vp)
_fp8_plan_cache: dict[tuple, tuple] = {}


def _fp8_plan(M: int, N: int, K: int) -> tuple:
    key = (M, N, K)
    p = _fp8_plan_cache.get(key)
    if p is None:
        swap, bn, splits = _i(), _i(), _i()
        _gemm_fp8_plan(M, N, K, ctypes.byref(swap), ctypes.byref(bn), ctypes.byref(splits))
        p = _fp8_plan_cache[key] = (swap.value, bn.value, 1 if batch_invariant() else splits.value)
    return p


def gemm_fp8_block(xq, xs, wq, ws, out_dtype=torch.bfloat16, bias=None, out=None):
    """Block-scaled fp8 GEMM (tcgen05 kind::f8f6f4, csrc/gemm_fp8.cu).  xq [M, K] e4m3; xs [M, K/128] fp32 with
    stride(0) == 1 (the column-major output of ``per_token_group_quant_fp8(col_major=True)``); wq [N, K] e4m3;
    ws [ceil(N/128), K/128] fp32."""
    M, K = xq.shape
    N = wq.shape[0]
    assert xs.stride(0) == 1 and ws.is_contiguous() and wq.is_contiguous() and xq.stride(1) == 1
    swap, bn, splits = _fp8_plan(M, N, K)
    y = out if out is not None else torch.empty((M, N), dtype=out_dtype, device=xq.device)
    _check(_gemm_fp8(xq.data_ptr(), wq.data_ptr(), xs.data_ptr(), ws.data_ptr(), _p(bias), y.data_ptr(), M, N, K,
                     xq.stride(0), xs.stride(1), 0, 1, 1, swap, bn, splits, _DT[y.dtype], _st()),
           f"gemm_fp8_block(M={M},N={N},K={K})")
    return y


def gemm_fp8_scaled(xq, wq, scale_a, scale_b, out_dtype=torch.bfloat16, bias=None):
    """W8A8 fp8 GEMM with per-token / per-tensor activation scales and per-channel / per-tensor weight scales
    (the reference's ``cutlass_scaled_mm`` contract): y = (xq · wq^T) * scale_a[:, None] * scale_b[None, :]."""
    M, K = xq.shape
    N = wq.shape[0]
    swap, bn, splits = _fp8_plan(M, N, K)
    y = torch.empty((M, N), dtype=out_dtype, device=xq.device)
    sa, sb = scale_a.reshape(-1).float().contiguous(), scale_b.reshape(-1).float().contiguous()
    _check(_gemm_fp8(xq.data_ptr(), wq.data_ptr(), sa.data_ptr(), sb.data_ptr(), _p(bias), y.data_ptr(), M, N, K,
                     xq.stride(0), 0, 1, int(sa.numel() > 1), int(sb.numel() > 1), swap, bn, splits, _DT[y.dtype], _st()),
           f"gemm_fp8_scaled(M={M},N={N},K={K})")
    return y


def gemm_fp8_grouped(xq, xs, wq, ws, offs, bn: int, out_dtype=torch.bfloat16):
    """Grouped (MoE) block-scaled fp8 GEMM: rows of xq sorted by expert (``offs``), wq [E, N, K], ws [E, N/128, K/128]."""
    R, K = xq.shape
    E, N, _ = wq.shape
    y = torch.empty((R, N), dtype=out_dtype, device=xq.device)
    _check(_gemm_fp8_grouped(xq.data_ptr(), wq.data_ptr(), xs.data_ptr(), ws.data_ptr(), y.data_ptr(), offs.data_ptr(), R, N,
                             K, E, xs.stride(1), bn, _DT[y.dtype], _st()), f"gemm_fp8_grouped(R={R},N={N},K={K},E={E})")
    return y


def block_fp8_linear(x, wq, w_scale, block=128, bias=None):
    """Activations quantised per (token, 128-group) on the fly, weights per 128x128 block: fused quantiser kernel +
    tcgen05 block-scaled GEMM.  Shapes the kernel does not cover take the fp32 reference."""
    K = x.shape[-1]
    x2 = x.reshape(-1, K)
    if (block == 128 and K % 128 == 0 and x2.dtype in (torch.bfloat16, torch.float16) and x2.shape[0] > 0
            and wq.dtype == torch.float8_e4m3fn and wq.is_contiguous() and w_scale.dtype == torch.float32):
        xq, xs = per_token_group_quant_fp8(x2, 128, col_major=True)
        y = gemm_fp8_block(xq, xs, wq, w_scale.contiguous(), x.dtype, bias)
        return y.view(x.shape[:-1] + (wq.shape[0],))
    from vllm_b200.ops import ref
    return ref.block_fp8_linear(x, wq, w_scale, block, bias)


def scaled_mm(a, b, scale_a, scale_b, out_dtype, bias=None):
    if (a.dtype == torch.float8_e4m3fn and b.dtype == torch.float8_e4m3fn and a.shape[1] % 16 == 0 and a.stride(1) == 1
            and b.is_contiguous() and out_dtype in (torch.bfloat16, torch.float16) and (a.stride(0) % 16) == 0):
        return gemm_fp8_scaled(a, b, scale_a, scale_b, out_dtype, bias)
    from vllm_b200.ops import ref
    return ref.scaled_mm(a, b, scale_a, scale_b, out_dtype, bias)


# ------------------------------------------------------------------ LoRA (csrc/lora.cu)
_lora_shrink = _sig("vb_lora_shrink", _vp, _i64, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp)
_lora_expand = _sig("vb_lora_expand", _vp, _i64, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp)


def lora_bgmv(x, y, A, B, token_slots, offs) -> None:
    """y[t] += B[slot(t)] (A[slot(t)] x[t]) for every token with slot >= 0, in place.  A [S, n_sub, r, K],
    B [S, N, r] (scaling folded in), ``offs`` = (n_sub, o1, o2, o3) sub-projection boundaries of the N outputs."""
    T, K = x.shape
    N, r = B.shape[1], B.shape[2]
    n_sub, o1, o2, o3 = offs
    v = torch.empty((T, n_sub, r), dtype=torch.float32, device=x.device)
    _check(_lora_shrink(x.data_ptr(), x.stride(0), A.data_ptr(), token_slots.data_ptr(), v.data_ptr(), T, K, r, n_sub,
                        _DT[x.dtype], _st()), "lora_shrink")
    _check(_lora_expand(y.data_ptr(), y.stride(0), B.data_ptr(), token_slots.data_ptr(), v.data_ptr(), T, N, r, n_sub,
                        o1, o2, o3, _DT[y.dtype], _st()), "lora_expand")


# ------------------------------------------------------------------ activation quantisation (csrc/quant.cu)
_quant_rows = _sig("vb_quant_rows", _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _f, _i,
                   _vp)
_absmax_scale = _sig("vb_absmax_scale", _vp, _i64, _i64, _i64, _vp, _vp, _f, _i, _i, _vp)
_fp4_quant = _sig("vb_fp4_quant", _vp, _i64, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp)
_Q_FP8_STATIC, _Q_FP8_TOKEN, _Q_FP8_GROUP, _Q_INT8_TOKEN, _Q_INT8_STATIC = range(5)
_PRE_NONE, _PRE_RMS, _PRE_ACT = range(3)
_amax_scratch: dict[int, torch.Tensor] = {}


def _quant_call(x2, H, pre, qm, out_dtype, *, residual=None, weight=None, scale_in=None, group=128, col_major=False,
                ue8m0=False, gemma=False, act=0, eps=0.0, scale_ub=0.0):
    T = x2.shape[0]
    dev = x2.device
    q = torch.empty((T, H), dtype=out_dtype, device=dev)
    scale = None
    stride = 0
    if qm in (_Q_FP8_TOKEN, _Q_INT8_TOKEN):
        scale = torch.empty((T, 1), dtype=torch.float32, device=dev)
    elif qm == _Q_FP8_GROUP:
        G = H // group
        if col_major:
            stride = (T + 3) // 4 * 4            # TMA-friendly leading dimension
            scale = torch.empty((G, stride), dtype=torch.float32, device=dev)
        else:
            scale = torch.empty((T, G), dtype=torch.float32, device=dev)
    _check(_quant_rows(x2.data_ptr(), x2.stride(0), _p(residual), _p(weight), q.data_ptr(), _p(scale), _p(scale_in), T, H,
                       pre, qm, group, int(col_major), stride, int(ue8m0), int(gemma), act, eps, scale_ub, _DT[x2.dtype],
                       _st()), "quant_rows")
    if qm == _Q_FP8_GROUP and col_major:
        scale = scale[:, :T].t()                 # logical [T, G] view over the column-major storage
    return q, scale


def _rows(x):
    x2 = x.reshape(-1, x.shape[-1])
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    return x2


def dynamic_scale(x: torch.Tensor, qmax: float = 448.0) -> torch.Tensor:
    """scale[0] = max|x| / qmax on the device (pass 1 of dynamic per-tensor quantisation)."""
    x2 = _rows(x)
    dev = x2.device.index or 0
    scr = _amax_scratch.get(dev)
    if scr is None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("dynamic per-tensor quantisation scratch must exist before CUDA-graph capture")
        scr = _amax_scratch[dev] = torch.zeros(1, dtype=torch.int32, device=x2.device)
    scale = torch.empty(1, dtype=torch.float32, device=x2.device)
    _check(_absmax_scale(x2.data_ptr(), x2.shape[0], x2.shape[1], x2.stride(0), scr.data_ptr(), scale.data_ptr(), qmax,
                         num_sms(), _DT[x2.dtype], _st()), "absmax_scale", 2)
    return scale


def scaled_fp8_quant(x, scale=None, per_token: bool = False, scale_ub: float = 0.0):
    x2 = _rows(x)
    H = x2.shape[1]
    if per_token:
        q, s = _quant_call(x2, H, _PRE_NONE, _Q_FP8_TOKEN, torch.float8_e4m3fn, scale_ub=scale_ub)
        return q.view(x.shape), s.view(x.shape[:-1] + (1,))
    if scale is None:
        scale = dynamic_scale(x2)
    q, _ = _quant_call(x2, H, _PRE_NONE, _Q_FP8_STATIC, torch.float8_e4m3fn, scale_in=scale.reshape(-1))
    return q.view(x.shape), scale


def scaled_int8_quant(x, scale=None):
    x2 = _rows(x)
    H = x2.shape[1]
    if scale is None:
        q, s = _quant_call(x2, H, _PRE_NONE, _Q_INT8_TOKEN, torch.int8)
        return q.view(x.shape), s.view(x.shape[:-1] + (1,))
    q, _ = _quant_call(x2, H, _PRE_NONE, _Q_INT8_STATIC, torch.int8, scale_in=scale.reshape(-1))
    return q.view(x.shape), scale


def per_token_group_quant_fp8(x, group: int = 128, col_major: bool = False, ue8m0: bool = False):
    x2 = _rows(x)
    H = x2.shape[1]
    if H % group or group not in (32, 64, 128) or H % 8 or x2.dtype not in (torch.bfloat16, torch.float16):
        from vllm_b200.ops import ref
        return ref.per_token_group_quant_fp8(x2, group)
    return _quant_call(x2, H, _PRE_NONE, _Q_FP8_GROUP, torch.float8_e4m3fn, group=group, col_major=col_major, ue8m0=ue8m0)


def rms_norm_quant(x, weight, eps, *, residual=None, scale=None, mode: str = "static", group: int = 128,
                   gemma: bool = False, col_major: bool = False, ue8m0: bool = False):
    """Fused (residual add +) RMSNorm + fp8 quantisation.  mode: static | token | group.  ``residual`` is updated in
    place (residual += x) like fused_add_rms_norm.  Returns (q, scale)."""
    x2 = _rows(x)
    H = x2.shape[1]
    qm = {"static": _Q_FP8_STATIC, "token": _Q_FP8_TOKEN, "group": _Q_FP8_GROUP}[mode]
    if residual is not None:
        assert residual.is_contiguous()
    q, s = _quant_call(x2, H, _PRE_RMS, qm, torch.float8_e4m3fn, residual=residual, weight=weight, eps=eps, gemma=gemma,
                       scale_in=None if scale is None else scale.reshape(-1), group=group, col_major=col_major,
                       ue8m0=ue8m0)
    return q.view(x.shape), (scale if mode == "static" else s)


def act_mul_quant(x, act: str = "silu", *, scale=None, mode: str = "static", group: int = 128, col_major: bool = False,
                  ue8m0: bool = False):  # noqa: D401
    """Fused act(gate) * up + fp8 quantisation over x[..., 2H] → (q[..., H], scale)."""
    x2 = _rows(x)
    H = x2.shape[1] // 2
    qm = {"static": _Q_FP8_STATIC, "token": _Q_FP8_TOKEN, "group": _Q_FP8_GROUP}[mode]
    a = {"silu": 0, "gelu": 1, "gelu_tanh": 2, "gelu_pytorch_tanh": 2}[act]
    q, s = _quant_call(x2, H, _PRE_ACT, qm, torch.float8_e4m3fn, act=a, scale_in=None if scale is None else scale.reshape(-1),
                       group=group, col_major=col_major, ue8m0=ue8m0)
    return q.view(x.shape[:-1] + (H,)), (scale if mode == "static" else s)


def scaled_fp4_quant(x, global_scale=None, mx: bool = False, swizzled: bool = False):
    """NVFP4 (block 16, e4m3 block scales x fp32 ``global_scale``) or MXFP4 (block 32, e8m0 scales).
    Returns (packed uint8 [T, H/2], block scales uint8 [T, H/16 or H/32] — or the 128x4-tiled buffer if swizzled)."""
    x2 = _rows(x)
    T, H = x2.shape
    blk = 32 if mx else 16
    q = torch.empty((T, H // 2), dtype=torch.uint8, device=x2.device)
    G = H // blk
    if swizzled:
        sf = torch.zeros(((T + 127) // 128) * ((G + 3) // 4) * 512, dtype=torch.uint8, device=x2.device)
    else:
        sf = torch.empty((T, G), dtype=torch.uint8, device=x2.device)
    if global_scale is None and not mx:
        global_scale = torch.ones(1, dtype=torch.float32, device=x2.device)
    _check(_fp4_quant(x2.data_ptr(), x2.stride(0), q.data_ptr(), sf.data_ptr(), _p(global_scale), T, H, int(mx),
                      int(swizzled), _DT[x2.dtype], _st()), "fp4_quant")
    return q, sf


# ------------------------------------------------------------------ q/k norm + RoPE + cache (csrc/qknorm_rope.cu)
_qknorm_rope = _sig("vb_qknorm_rope_cache", _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i64, _i64, _i64, _vp,
                    _vp, _i, _i, _i, _i, _f, _f, _f, _i, _vp)


def qknorm_rope_ok(head_dim: int, rot: int, is_neox: bool) -> bool:
    E = head_dim // 32
    if head_dim % 32 or E not in (2, 4, 8, 16):
        return False
    if rot:
        if is_neox:
            half = rot // 2
            if half % E:
                return False
            span = half // E
            return span & (span - 1) == 0
        return E % 2 == 0
    return True


def fused_qk_norm_rope(positions, q, k, q_weight, k_weight, cos_sin_cache, head_dim, eps, is_neox=True, *, v=None,
                       kv_cache=None, slot_mapping=None, gemma: bool = False, k_scale: float = 1.0, v_scale: float = 1.0,
                       vt: bool = False):
    """In place: per-head RMSNorm(q), RMSNorm(k) → RoPE → (optionally) K/V insert into the paged cache; one launch."""
    T = q.shape[0]
    Hq, Hkv = q.shape[-1] // head_dim, k.shape[-1] // head_dim
    rot = cos_sin_cache.shape[-1] if cos_sin_cache is not None else 0
    assert q.stride(-1) == 1 and k.stride(-1) == 1
    flags = 0
    BS = 1
    if kv_cache is not None:
        flags = _kv_fp8(kv_cache) | (2 if vt else 0)
        BS = kv_cache.shape[3]
    _check(_qknorm_rope(_p(positions), q.data_ptr(), k.data_ptr(), _p(v), _p(q_weight), _p(k_weight), _p(cos_sin_cache),
                        rot, T, Hq, Hkv, head_dim, q.stride(0), k.stride(0), v.stride(0) if v is not None else 0,
                        _p(kv_cache), _p(slot_mapping), BS, int(is_neox), int(gemma), flags, eps, k_scale, v_scale,
                        _DT[q.dtype], _st()), "qknorm_rope_cache")
    return q, k


# ------------------------------------------------------------------ cache maintenance (csrc/cache_ops.cu)
_copy_blocks = _sig("vb_copy_blocks", _vp, _vp, _vp, _i, _i64, _vp)
_gather_cache = _sig("vb_gather_cache", _vp, _vp, _vp, _i64, _vp, _vp, _i, _i, _i, _i, _i64, _i64, _f, _i, _i, _vp)
_convert_fp8 = _sig("vb_convert_fp8", _vp, _vp, _i64, _f, _i, _i, _vp)


def swap_blocks(src: torch.Tensor, dst: torch.Tensor, mapping) -> None:
    """dst[j] = src[i] for (i, j) in mapping; whole blocks (dim 0).  Either side may be pinned host memory."""
    if isinstance(mapping, torch.Tensor):
        m = mapping.to(torch.int64)
    else:
        if not mapping:
            return
        m = torch.tensor(mapping, dtype=torch.int64)
    dev = src.device if src.is_cuda else dst.device
    m = m.to(dev, non_blocking=True).contiguous()
    block_bytes = src[0].numel() * src.element_size()
    assert dst[0].numel() * dst.element_size() == block_bytes and src.is_contiguous() and dst.is_contiguous()
    assert block_bytes % 16 == 0
    for t in (src, dst):
        assert t.is_cuda or t.is_pinned(), "host side of swap_blocks must be pinned (mapped) memory"
    _check(_copy_blocks(src.data_ptr(), dst.data_ptr(), m.data_ptr(), m.shape[0], block_bytes, _st()), "copy_blocks")


def copy_blocks(cache: torch.Tensor, mapping) -> None:
    swap_blocks(cache, cache, mapping)


def gather_cache(cache, block_table, cu_seq_lens, out, max_tokens: int, seq_starts=None, scale: float = 1.0):
    """Dense [total_tokens, row] copy of the cached rows of a batch (cache [blocks, BS, row], fp8 → out.dtype with
    ``scale`` when the cache is 8-bit)."""
    B = cu_seq_lens.numel() - 1
    BS, row = cache.shape[1], cache.shape[2]
    src8 = cache.element_size() == 1
    _check(_gather_cache(cache.data_ptr(), out.data_ptr(), block_table.data_ptr(), block_table.stride(0),
                         cu_seq_lens.data_ptr(), _p(seq_starts), B, max_tokens, BS, row, cache.stride(0), out.stride(0),
                         scale, int(src8), _DT[out.dtype], _st()), "gather_cache")
    return out


def convert_fp8(x: torch.Tensor, scale: float = 1.0, to_fp8: bool = True, out_dtype=torch.bfloat16):
    xc = x.contiguous()
    n = xc.numel()
    if n % 8:
        from vllm_b200.ops import ref
        return ref.convert_fp8(x, scale, to_fp8)
    if to_fp8:
        out = torch.empty(xc.shape, dtype=torch.float8_e4m3fn, device=xc.device)
        _check(_convert_fp8(xc.data_ptr(), out.data_ptr(), n, scale, 1, num_sms(), _DT[xc.dtype], _st()), "convert_fp8")
    else:
        out = torch.empty(xc.shape, dtype=out_dtype, device=xc.device)
        _check(_convert_fp8(xc.data_ptr(), out.data_ptr(), n, scale, 0, num_sms(), _DT[out_dtype], _st()), "convert_fp8")
    return out


# ------------------------------------------------------------------ SSM / linear attention (csrc/ssm.cu)
_conv1d = _sig("vb_causal_conv1d", _vp, _i64, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _vp)
_sel_scan = _sig("vb_selective_scan", _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _i, _i, _i,
                 _vp)
_ssd_scan = _sig("vb_ssd_scan", _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _i, _i, _i,
                 _i, _i, _f, _f, _vp)
_gdn_scan = _sig("vb_gdn_scan", _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _vp)


def _state_view(cache: torch.Tensor) -> torch.Tensor:
    """The recurrent-state pool as [rows, floats] fp32 (one row per request)."""
    c = cache.view(cache.shape[0], -1)
    return c if c.dtype == torch.float32 else c.view(torch.float32)


def causal_conv1d(x, weight, bias, cache, md, act: bool = True) -> torch.Tensor:
    """Depthwise causal conv over each request's tokens with the (k-1)-window kept at the head of its state row.
    x [T, C] (bf16/fp16, any row stride) → fp32 [T, C]; the window in the cache is advanced in place."""
    T, C = x.shape
    st = _state_view(cache)
    out = torch.zeros((T, C), dtype=torch.float32, device=x.device)
    bt = md.block_table
    _check(_conv1d(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), weight.data_ptr(), _p(bias), st.data_ptr(),
                   st.stride(0), md.query_start_loc.data_ptr(), md.seq_lens.data_ptr(), bt.data_ptr(), bt.stride(0),
                   md.num_reqs, C, weight.shape[-1], int(act), _DT[x.dtype], _st()), "causal_conv1d")
    return out


def selective_scan(x, dt, B, C, A, D, cache, state_off: int, md) -> torch.Tensor:
    """Mamba-1: fp32 x/dt [T, di], B/C [T, N], A [di, N], D [di] → y [T, di]; state rows updated in place."""
    T, di = x.shape
    N = B.shape[-1]
    st = _state_view(cache)
    y = torch.zeros((T, di), dtype=torch.float32, device=x.device)
    bt = md.block_table
    args = [t.contiguous() for t in (x, dt, B, C, A, D)]
    _check(_sel_scan(*[t.data_ptr() for t in args], y.data_ptr(), st.data_ptr(), st.stride(0), state_off,
                     md.query_start_loc.data_ptr(), md.seq_lens.data_ptr(), bt.data_ptr(), bt.stride(0), md.num_reqs, di, N,
                     _st()), "selective_scan")
    return y


def ssd_scan(xbc, dt_raw, dt_bias, A_log, D, cache, state_off: int, md, nh: int, hd: int, G: int, N: int,
             dt_min: float = 0.0, dt_max: float = float("inf")) -> torch.Tensor:
    """Mamba-2: conv output xbc fp32 [T, nh*hd + 2*G*N], raw dt [T, nh] fp32 → y [T, nh*hd] (D skip included)."""
    T = xbc.shape[0]
    st = _state_view(cache)
    y = torch.zeros((T, nh * hd), dtype=torch.float32, device=xbc.device)
    bt = md.block_table
    dt_max = min(dt_max, 3.0e38)
    _check(_ssd_scan(xbc.data_ptr(), xbc.stride(0), dt_raw.data_ptr(), dt_raw.stride(0), dt_bias.data_ptr(),
                     A_log.data_ptr(), D.data_ptr(), y.data_ptr(), st.data_ptr(), st.stride(0), state_off,
                     md.query_start_loc.data_ptr(), md.seq_lens.data_ptr(), bt.data_ptr(), bt.stride(0), md.num_reqs, nh, hd,
                     G, N, dt_min, dt_max, _st()), "ssd_scan")
    return y


def gdn_scan(q, k, v, g, beta, cache, state_off: int, md) -> torch.Tensor:
    """Gated delta rule: fp32 q/k [T, nv, dk], v [T, nv, dv], g/beta [T, nv] → o [T, nv, dv]; state [nv, dk, dv]."""
    T, nv, dk = q.shape
    dv = v.shape[-1]
    st = _state_view(cache)
    o = torch.zeros((T, nv, dv), dtype=torch.float32, device=q.device)
    bt = md.block_table
    a = [t.contiguous() for t in (q, k, v, g, beta)]
    _check(_gdn_scan(*[t.data_ptr() for t in a], o.data_ptr(), st.data_ptr(), st.stride(0), state_off,
                     md.query_start_loc.data_ptr(), md.seq_lens.data_ptr(), bt.data_ptr(), bt.stride(0), md.num_reqs, nv, dk,
                     dv, _st()), "gdn_scan")
    return o


# ------------------------------------------------------------------ GPU-side input preparation (csrc/input_prep.cu)
_prep_step = _sig("vb_prep_step3", _vp, _vp, _vp, _i, _i, _i, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp,
                  _i, _vp)
_bt_diff = _sig("vb_block_table_diff", _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i, _vp)
_append_sampled = _sig("vb_append_sampled", _vp, _i64, _vp, _vp, _vp, _i, _vp)


def prep_step(rows, num_computed, num_sched, token_table, block_rows, input_ids, positions, qsl, seq_lens, slot_mapping,
              block_table, nblk: int, block_size: int, pos_delta=None, pad_b: int = 0, write_tokens: bool = True,
              pad_t: int = 0) -> None:
    """Expand per-request (row, computed, scheduled) triples into the per-token step inputs on the device.  Entries
    [nb, pad_b) are CUDA-graph bucket padding: one-token requests on the null block (decode buckets, ``pad_t`` = 0) or
    empty requests plus padding tokens up to ``pad_t`` (mixed-batch buckets)."""
    nb = rows.numel()
    _check(_prep_step(rows.data_ptr(), num_computed.data_ptr(), num_sched.data_ptr(), nb, max(pad_b, nb), pad_t,
                      token_table.data_ptr(), token_table.stride(0), block_rows.data_ptr(), block_rows.stride(0),
                      input_ids.data_ptr(), positions.data_ptr(), qsl.data_ptr(), seq_lens.data_ptr(),
                      slot_mapping.data_ptr(), block_table.data_ptr(), block_table.stride(0), nblk, block_size,
                      _p(pos_delta), int(write_tokens), _st()), "prep_step")


def block_table_apply_diffs(block_rows, d_row, d_start, d_cnt, d_off, ids_flat) -> None:
    _check(_bt_diff(block_rows.data_ptr(), block_rows.stride(0), d_row.data_ptr(), d_start.data_ptr(), d_cnt.data_ptr(),
                    d_off.data_ptr(), ids_flat.data_ptr(), d_row.numel(), _st()), "block_table_diff")


def append_sampled(token_table, rows, pos, sampled) -> None:
    _check(_append_sampled(token_table.data_ptr(), token_table.stride(0), rows.data_ptr(), pos.data_ptr(),
                           sampled.data_ptr(), rows.numel(), _st()), "append_sampled")


# ------------------------------------------------------------------ weight-only FP4 decode GEMM (csrc/gemm_w4.cu)
_gemm_w4 = _sig("vb_gemm_w4", _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _f, _i, _i, _vp)
W4_MAX_TOKENS = 128


def pack_w4_tiles(codes: torch.Tensor, scales: torch.Tensor, fmt: str) -> torch.Tensor:
    """Tile-major operand layout of the W4A16 kernel.  ``codes``: uint8 [N, K/2] (two e2m1 codes per byte, low nibble
    first); ``scales``: one byte per 16 (``nvfp4``: e4m3) or 32 (``mxfp4``: e8m0) columns, [N, K/16] or [N, K/32].
    Returns uint8 [ceil(N/128), K/64, 4096 + 128 * {4, 2}]: every (128-row tile, 64-column block) is one contiguous
    record = the codes of its 128 rows followed by their scales."""
    N, K2 = codes.shape
    K = K2 * 2
    assert K % 64 == 0, "W4 kernel: K must be a multiple of 64"
    spb = 4 if fmt == "nvfp4" else 2                 # scale bytes per row per 64-column block
    sc = scales.view(torch.uint8) if scales.dtype != torch.uint8 else scales
    assert sc.shape == (N, K // 64 * spb), (sc.shape, N, K)
    Np = (N + 127) // 128 * 128
    if Np != N:
        codes = torch.nn.functional.pad(codes, (0, 0, 0, Np - N))
        sc = torch.nn.functional.pad(sc, (0, 0, 0, Np - N))
    c = codes.view(Np // 128, 128, K // 64, 32).permute(0, 2, 1, 3).reshape(Np // 128, K // 64, 4096)
    s = sc.reshape(Np // 128, 128, K // 64, spb).permute(0, 2, 1, 3).reshape(Np // 128, K // 64, 128 * spb)
    return torch.cat([c, s], dim=-1).contiguous()


def gemm_w4(x: torch.Tensor, tiled: torch.Tensor, N: int, gscale: float, fmt: str, bias: Optional[torch.Tensor] = None
            ) -> torch.Tensor:
    """Y = X · dequant(W)^T for FP4 weights in the tile-major layout of ``pack_w4_tiles`` (M <= 128 tokens)."""
    M, K = x.shape
    y = torch.empty((M, N), dtype=x.dtype, device=x.device)
    _check(_gemm_w4(x.data_ptr(), tiled.data_ptr(), _p(bias), y.data_ptr(), M, N, K, x.stride(0), float(gscale),
                    0 if fmt == "nvfp4" else 1, _DT[x.dtype], _st()), f"gemm_w4(M={M},N={N},K={K},{fmt})")
    return y


def gemm_w4_ok(x: torch.Tensor, K: int) -> bool:
    return (x.is_cuda and x.dtype in (torch.bfloat16, torch.float16) and x.dim() == 2 and 0 < x.shape[0] <= W4_MAX_TOKENS
            and K % 64 == 0 and x.stride(1) == 1 and (x.stride(0) * 2) % 16 == 0)


# ------------------------------------------------------------------ L2 weight prefetch on a side stream
_l2_prefetch = _sig("vb_l2_prefetch", _vp, _i64, _vp)
_side_streams: dict[int, torch.cuda.Stream] = {}
_side_dirty: dict[int, bool] = {}


def l2_prefetch_async(t: torch.Tensor) -> None:
    """Start pulling ``t`` into L2 on a forked stream, ordered after everything enqueued so far on the current stream
    (so it runs concurrently with the NEXT kernel launched here).  A hint only: nothing ever waits for it except
    ``l2_prefetch_join`` at the end of the forward pass (required to close a CUDA-graph capture)."""
    dev = t.device.index or 0
    side = _side_streams.get(dev)
    if side is None:
        if torch.cuda.is_current_stream_capturing():
            return
        side = _side_streams[dev] = torch.cuda.Stream(t.device)
    cur = torch.cuda.current_stream()
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        _check(_l2_prefetch(t.data_ptr(), t.numel() * t.element_size(), side.cuda_stream), "l2_prefetch", 0)
    _side_dirty[dev] = True


def l2_prefetch_join() -> None:
    dev = torch.cuda.current_device()
    if _side_dirty.get(dev):
        torch.cuda.current_stream().wait_stream(_side_streams[dev])
        _side_dirty[dev] = False
