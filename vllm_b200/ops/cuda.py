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
_fmha_tc = _sig("vb_fmha_tc", _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i64, _i64, _i,
                _i, _i, _f, _i, _i, _vp)
_fmha_tc_decode = _sig("vb_fmha_tc_decode", _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i64, _i,
                       _i, _i, _f, _i, _vp)
_merge_states = _sig("vb_merge_attn_states", _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp)
_argmax = _sig("vb_argmax", _vp, _i64, _i, _i, _vp, _i, _vp)
_fused_sample = _sig("vb_fused_sample", _vp, _i64, _i, _i, _vp, _vp, _vp, _vp, _i, _vp)
_topk_topp = _sig("vb_topk_topp", _vp, _i, _i, _vp, _vp, _vp, _vp)
_row_lse = _sig("vb_row_lse", _vp, _i64, _i, _i, _vp, _i, _vp)
_gemm_plan = _sig("vb_gemm_tc_plan", _i, _i, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i),
                  ctypes.POINTER(_i64), ctypes.POINTER(_i))
_gemm_tc = _sig("vb_gemm_tc", _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i, _i, _i, _vp, _vp, _i, _vp)


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


def decode_splits(B: int, Hkv: int, max_seq_len: int, block_size: int) -> int:
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
        out = torch.empty((T, Hq, D), dtype=q.dtype, device=q.device)
    lse = torch.empty((T, Hq), dtype=torch.float32, device=q.device) if return_lse else None
    bt = md.block_table
    if md.kv_vt:
        # tcgen05 kernel over the tensor-core cache layout
        assert D == 128 and BS == 64 and not logits_soft_cap and alibi_slopes is None and sinks is None and not chunk_size, \
            "tensor-core attention: D=128, 64-token pages, no soft-cap / ALiBi / sinks / chunked-local mask"
        decode = md.decode_only
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
                      sliding_window or 0, scale, int(causal), _DT[q.dtype], _st())
        _check(rc, f"fmha_tc(G={Hq // Hkv})", 2 if splits > 1 else 1)
        return (out, lse) if return_lse else out
    if alibi_slopes is not None:
        alibi_slopes = alibi_slopes.to(device=q.device, dtype=torch.float32).contiguous()
    if sinks is not None:
        sinks = sinks.to(device=q.device, dtype=torch.float32).contiguous()
    if md.decode_only:
        splits = md.num_splits or decode_splits(md.num_reqs, Hkv, md.max_seq_len, BS)
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
    _check(rc, f"paged_attention(D={D}, BS={BS}, G={Hq // Hkv})", 2 if (md.decode_only and splits > 1) else 1)
    return (out, lse) if return_lse else out


def merge_attn_states(o1, lse1, o2, lse2):
    T, H, D = o1.shape
    out = torch.empty_like(o1)
    lse = torch.empty_like(lse1)
    _check(_merge_states(out.data_ptr(), lse.data_ptr(), o1.data_ptr(), lse1.data_ptr(), o2.data_ptr(),
                         lse2.data_ptr(), T * H, D, _DT[o1.dtype], _st()), "merge_attn_states")
    return out, lse


def mla_attention(q_nope_abs, q_pe, kv_cache, md, scale, kv_lora_rank):
    """MLA as MQA over the latent cache (K = 576-d latent+rope, V = first 512 dims of the same rows).

    Decode batches run as dense batched GEMMs over the gathered latent pages with a device-side length mask:
    no host synchronisation, static shapes → CUDA-graph capturable.  Mixed / prefill batches use the
    variable-length loop (eager only).  A dedicated tcgen05 MLA kernel (latent tile shared by all 128 heads
    through TMEM) is the planned replacement for both."""
    from vllm_b200.ops import ref
    if md.decode_only:
        B, H, R = q_nope_abs.shape
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
        p = (swap.value, bn.value, splits.value, wsb.value, tiles.value)
        _plan_cache[key] = p
    return p


def gemm_tc(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Y = X W^T on our tcgen05 kernel. x: [M, K] (row stride free), w: [N, K] contiguous."""
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
    y = out if out is not None else torch.empty((M, N), dtype=x.dtype, device=x.device)
    _check(_gemm_tc(x.data_ptr(), w.data_ptr(), _p(bias), y.data_ptr(), M, N, K, x.stride(0), swap, bn, splits,
                    _p(g.ws), _p(g.counters), _DT[x.dtype], _st()), f"gemm_tc(M={M},N={N},K={K})")
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


def comm_barrier(c) -> None:
    _check(_comm_bar(c.local_ptr, c.peer_arr, c._mc(), c.state.data_ptr(), c.rank, c.world, _st()), "comm_barrier")


# ------------------------------------------------------------------ MoE (csrc/moe.cu + grouped mode of gemm_tc.cu)
_moe_topk = _sig("vb_moe_topk_softmax", _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp)
_moe_align = _sig("vb_moe_align", _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp)
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


def moe_align(topk_ids: torch.Tensor, num_local_experts: int, expert_map: Optional[torch.Tensor] = None):
    """(sorted_tok [R], inv [R], offs [E+1]) for R = T*k (token, slot) pairs — all on the device."""
    T, k = topk_ids.shape
    R = T * k
    dev = topk_ids.device
    sorted_tok = torch.empty(R, dtype=torch.int32, device=dev)
    inv = torch.empty(R, dtype=torch.int32, device=dev)
    offs = torch.empty(num_local_experts + 1, dtype=torch.int32, device=dev)
    _check(_moe_align(topk_ids.data_ptr(), _p(expert_map), sorted_tok.data_ptr(), inv.data_ptr(), offs.data_ptr(), R, k,
                      num_local_experts, _st()), "moe_align")
    return sorted_tok, inv, offs


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


def block_fp8_linear(x, wq, w_scale, block=128, bias=None):
    from vllm_b200.ops import ref
    return ref.block_fp8_linear(x, wq, w_scale, block, bias)
