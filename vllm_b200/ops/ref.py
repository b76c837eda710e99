"""Plain-PyTorch reference implementations of every op.

They are (a) the CPU execution path (BASELINE config 1: OPT-125m greedy decode
with no GPU) and (b) the fp32 oracles the CUDA kernels are tested against.
Semantics follow the reference's ``forward_native`` methods (e.g.
vllm/model_executor/layers/layernorm.py:96-116, activation.py,
rotary_embedding/base.py) — re-derived, not copied.

KV-cache layout (ours, chosen so one (page, head) K or V tile is a single
contiguous 1-D bulk-copy): ``[num_blocks, 2, num_kv_heads, block_size, head_dim]``.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F


# ---------------------------------------------------------------- norms
def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float, gemma: bool = False) -> torch.Tensor:
    xf = x.float()
    var = xf.pow(2).mean(-1, keepdim=True)
    w = weight.float()
    if gemma:
        w = w + 1.0
    return (xf * torch.rsqrt(var + eps) * w).to(x.dtype)


def fused_add_rms_norm(x: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor, eps: float,
                       gemma: bool = False) -> tuple[torch.Tensor, torch.Tensor]:
    """residual_out = x + residual (rounded to dtype); out = rms_norm(residual_out)."""
    r = (x.float() + residual.float()).to(x.dtype)
    return rms_norm(r, weight, eps, gemma), r


def layer_norm(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], eps: float) -> torch.Tensor:
    return F.layer_norm(x.float(), (x.shape[-1],), weight.float(), None if bias is None else bias.float(),
                        eps).to(x.dtype)


# ----------------------------------------------------------- activations
def silu_and_mul(x: torch.Tensor) -> torch.Tensor:
    d = x.shape[-1] // 2
    return (F.silu(x[..., :d].float()) * x[..., d:].float()).to(x.dtype)


def gelu_and_mul(x: torch.Tensor, approximate: str = "none") -> torch.Tensor:
    d = x.shape[-1] // 2
    return (F.gelu(x[..., :d].float(), approximate=approximate) * x[..., d:].float()).to(x.dtype)


def gelu_new(x: torch.Tensor) -> torch.Tensor:
    xf = x.float()
    return (0.5 * xf * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (xf + 0.044715 * xf.pow(3))))).to(x.dtype)


def gelu_fast(x: torch.Tensor) -> torch.Tensor:
    xf = x.float()
    return (0.5 * xf * (1.0 + torch.tanh(xf * 0.7978845608 * (1.0 + 0.044715 * xf * xf)))).to(x.dtype)


def quick_gelu(x: torch.Tensor) -> torch.Tensor:
    xf = x.float()
    return (xf * torch.sigmoid(1.702 * xf)).to(x.dtype)


def mul_and_silu(x: torch.Tensor) -> torch.Tensor:
    d = x.shape[-1] // 2
    return (x[..., :d].float() * F.silu(x[..., d:].float())).to(x.dtype)


def fatrelu_and_mul(x: torch.Tensor, threshold: float = 0.0) -> torch.Tensor:
    d = x.shape[-1] // 2
    g = x[..., :d].float()
    return (torch.where(g > threshold, g, torch.zeros_like(g)) * x[..., d:].float()).to(x.dtype)


def swigluoai_and_mul(x: torch.Tensor, alpha: float = 1.702, limit: float = 7.0) -> torch.Tensor:
    """gpt-oss SwiGLU: interleaved (gate, up) pairs, clamped, ``(up + 1) * gate * sigmoid(alpha * gate)``."""
    g, u = x[..., ::2].float().clamp(max=limit), x[..., 1::2].float().clamp(-limit, limit)
    return ((u + 1.0) * g * torch.sigmoid(alpha * g)).to(x.dtype)


def fused_qk_norm_rope(positions: torch.Tensor, q: torch.Tensor, k: torch.Tensor, q_weight: torch.Tensor,
                       k_weight: torch.Tensor, cos_sin_cache: torch.Tensor, head_dim: int, eps: float,
                       is_neox: bool = True) -> tuple[torch.Tensor, torch.Tensor]:
    """Per-head RMSNorm of q and k followed by RoPE (Qwen3 / Gemma-3 attention prologue)."""
    T = q.shape[0]
    qn = rms_norm(q.reshape(T, -1, head_dim), q_weight, eps).reshape(T, -1)
    kn = rms_norm(k.reshape(T, -1, head_dim), k_weight, eps).reshape(T, -1)
    return rotary_embedding(positions, qn, kn, cos_sin_cache, head_dim, is_neox)


ACT_FNS = {
    "relu": F.relu, "gelu": F.gelu, "gelu_new": gelu_new, "gelu_fast": gelu_fast,
    "gelu_pytorch_tanh": lambda x: F.gelu(x, approximate="tanh"), "quick_gelu": quick_gelu,
    "silu": F.silu, "relu2": lambda x: F.relu(x).square(),
}


# ----------------------------------------------------------------- RoPE
def rotary_embedding(positions: torch.Tensor, q: torch.Tensor, k: Optional[torch.Tensor],
                     cos_sin_cache: torch.Tensor, head_dim: int, is_neox: bool = True,
                     rot_dim: Optional[int] = None) -> tuple[torch.Tensor, Optional[torch.Tensor]]:
    """q: [T, Hq*D], k: [T, Hk*D]; cos_sin_cache: [max_pos, rot_dim] = cat(cos, sin) (fp32)."""
    rot = rot_dim or cos_sin_cache.shape[-1]
    cs = cos_sin_cache[positions.long()].float()  # [T, rot]
    cos, sin = cs[:, :rot // 2], cs[:, rot // 2:]

    def apply(x: torch.Tensor) -> torch.Tensor:
        T = x.shape[0]
        xv = x.view(T, -1, head_dim)
        xr, xp = xv[..., :rot].float(), xv[..., rot:]
        c, s = cos[:, None, :], sin[:, None, :]
        if is_neox:
            x1, x2 = xr[..., :rot // 2], xr[..., rot // 2:]
            o = torch.cat((x1 * c - x2 * s, x2 * c + x1 * s), dim=-1)
        else:
            x1, x2 = xr[..., 0::2], xr[..., 1::2]
            o = torch.stack((x1 * c - x2 * s, x2 * c + x1 * s), dim=-1).flatten(-2)
        return torch.cat((o.to(x.dtype), xp), dim=-1).reshape(x.shape)

    return apply(q), (apply(k) if k is not None else None)


# -------------------------------------------------------------- KV cache
def reshape_and_cache(key: torch.Tensor, value: torch.Tensor, kv_cache: torch.Tensor,
                      slot_mapping: torch.Tensor) -> None:
    """key/value: [T, Hkv, D]; kv_cache: [nb, 2, Hkv, bs, D]; slot = block*bs + off (<0 = skip)."""
    bs = kv_cache.shape[3]
    valid = slot_mapping >= 0
    slots = slot_mapping[valid].long()
    blk, off = slots // bs, slots % bs
    kv_cache[blk, 0, :, off, :] = key[valid].to(kv_cache.dtype)
    kv_cache[blk, 1, :, off, :] = value[valid].to(kv_cache.dtype)


def concat_and_cache_mla(kv_c: torch.Tensor, k_pe: torch.Tensor, kv_cache: torch.Tensor,
                         slot_mapping: torch.Tensor) -> None:
    """kv_c: [T, R], k_pe: [T, P]; kv_cache: [nb, bs, R+P]."""
    bs = kv_cache.shape[1]
    valid = slot_mapping >= 0
    slots = slot_mapping[valid].long()
    kv_cache[slots // bs, slots % bs] = torch.cat((kv_c[valid], k_pe[valid]), -1).to(kv_cache.dtype)


def gather_kv(kv_cache: torch.Tensor, block_table_row: torch.Tensor, seq_len: int
              ) -> tuple[torch.Tensor, torch.Tensor]:
    """Contiguous K, V [Hkv, seq_len, D] of one sequence."""
    bs = kv_cache.shape[3]
    nb = (seq_len + bs - 1) // bs
    blocks = kv_cache[block_table_row[:nb].long()]  # [nb, 2, H, bs, D]
    k = blocks[:, 0].permute(1, 0, 2, 3).reshape(kv_cache.shape[2], nb * bs, -1)[:, :seq_len]
    v = blocks[:, 1].permute(1, 0, 2, 3).reshape(kv_cache.shape[2], nb * bs, -1)[:, :seq_len]
    return k, v


# ------------------------------------------------------------- attention
def paged_attention(q: torch.Tensor, kv_cache: torch.Tensor, block_table: torch.Tensor,
                    seq_lens: torch.Tensor, query_start_loc: torch.Tensor, scale: float,
                    causal: bool = True, sliding_window: Optional[int] = None,
                    logits_soft_cap: float = 0.0, return_lse: bool = False,
                    kv_scale: float = 1.0, sinks: Optional[torch.Tensor] = None,
                    alibi_slopes: Optional[torch.Tensor] = None, chunk_size: int = 0):
    """Unified varlen paged attention.

    q: [T, Hq, D] (all requests' new tokens concatenated); request i owns rows
    ``query_start_loc[i]:query_start_loc[i+1]`` whose positions are the *last*
    q_len positions of its ``seq_lens[i]`` tokens.  KV for all positions
    (including the new tokens) must already be in the cache.
    """
    T, Hq, D = q.shape
    Hkv = kv_cache.shape[2]
    G = Hq // Hkv
    out = torch.empty_like(q)
    lse = torch.empty(T, Hq, dtype=torch.float32, device=q.device)
    qsl = query_start_loc.tolist()
    sls = seq_lens.tolist()
    for i in range(len(sls)):
        s, e = qsl[i], qsl[i + 1]
        ql, sl = e - s, sls[i]
        if ql == 0:
            continue
        k, v = gather_kv(kv_cache, block_table[i], sl)
        k = k.float() * kv_scale
        v = v.float() * kv_scale
        qi = q[s:e].float().permute(1, 0, 2)  # [Hq, ql, D]
        kk = k.repeat_interleave(G, dim=0)
        vv = v.repeat_interleave(G, dim=0)
        att = torch.einsum("hqd,hkd->hqk", qi, kk) * scale
        if logits_soft_cap > 0:
            att = logits_soft_cap * torch.tanh(att / logits_soft_cap)
        qpos = torch.arange(sl - ql, sl, device=q.device)[:, None]
        kpos = torch.arange(sl, device=q.device)[None, :]
        if alibi_slopes is not None:
            att = att + alibi_slopes.float().to(q.device)[:, None, None] * (kpos - qpos).float()[None]
        mask = torch.ones(ql, sl, dtype=torch.bool, device=q.device)
        if causal:
            mask &= kpos <= qpos
        if sliding_window:
            mask &= kpos > qpos - sliding_window
            if not causal:                   # bidirectional local attention (ModernBERT): the window is symmetric
                mask &= kpos < qpos + sliding_window
        if chunk_size:                       # chunked-local attention: keys of the query's own chunk only
            mask &= (kpos // chunk_size) == (qpos // chunk_size)
        att = att.masked_fill(~mask[None], float("-inf"))
        l = torch.logsumexp(att, dim=-1)
        if sinks is not None:            # a per-head logit that takes softmax mass but contributes no value
            l = torch.logaddexp(l, sinks.float().to(q.device)[:, None].expand_as(l))
        p = torch.exp(att - l[..., None])
        p = torch.nan_to_num(p)
        out[s:e] = torch.einsum("hqk,hkd->hqd", p, vv).permute(1, 0, 2).to(q.dtype)
        lse[s:e] = l.permute(1, 0)
    return (out, lse) if return_lse else out


def merge_attn_states(o1: torch.Tensor, lse1: torch.Tensor, o2: torch.Tensor, lse2: torch.Tensor
                      ) -> tuple[torch.Tensor, torch.Tensor]:
    """LSE-weighted merge of two partial attention results. o: [T,H,D], lse: [T,H]."""
    m = torch.maximum(lse1, lse2)
    m = torch.where(torch.isinf(m), torch.zeros_like(m), m)
    w1, w2 = torch.exp(lse1 - m), torch.exp(lse2 - m)
    den = (w1 + w2).clamp_min(1e-30)
    o = (o1.float() * (w1 / den)[..., None] + o2.float() * (w2 / den)[..., None]).to(o1.dtype)
    return o, m + torch.log(den)


def mla_decode_attention(q_nope_abs: torch.Tensor, q_pe: torch.Tensor, kv_cache: torch.Tensor,
                         block_table: torch.Tensor, seq_lens: torch.Tensor, query_start_loc: torch.Tensor,
                         scale: float, kv_lora_rank: int) -> torch.Tensor:
    """MLA as MQA over the latent cache. q_nope_abs: [T,H,R] (already multiplied by W_UK),
    q_pe: [T,H,P]; kv_cache: [nb, bs, R+P]; returns latent-space output [T,H,R]."""
    T, H, R = q_nope_abs.shape
    bs = kv_cache.shape[1]
    out = torch.empty(T, H, R, dtype=q_nope_abs.dtype, device=q_nope_abs.device)
    qsl, sls = query_start_loc.tolist(), seq_lens.tolist()
    for i in range(len(sls)):
        s, e = qsl[i], qsl[i + 1]
        ql, sl = e - s, sls[i]
        if ql == 0:
            continue
        nb = (sl + bs - 1) // bs
        lat = kv_cache[block_table[i, :nb].long()].reshape(nb * bs, -1)[:sl].float()
        kc, kp = lat[:, :kv_lora_rank], lat[:, kv_lora_rank:]
        att = (torch.einsum("qhr,kr->hqk", q_nope_abs[s:e].float(), kc)
               + torch.einsum("qhp,kp->hqk", q_pe[s:e].float(), kp)) * scale
        qpos = torch.arange(sl - ql, sl, device=att.device)[:, None]
        kpos = torch.arange(sl, device=att.device)[None, :]
        att = att.masked_fill(~(kpos <= qpos)[None], float("-inf"))
        p = torch.softmax(att, dim=-1)
        out[s:e] = torch.einsum("hqk,kr->qhr", p, kc).to(out.dtype)
    return out


def mla_partial_attention(q_abs: torch.Tensor, q_pe: torch.Tensor, lat: torch.Tensor, scale: float, kv_lora_rank: int,
                          causal_offset: Optional[int] = None) -> tuple[torch.Tensor, torch.Tensor]:
    """Attention of q [q, H, R]/[q, H, P] over latent rows ``lat`` [k, R+P]; returns (out [q,H,R] fp32, lse [q,H]).
    ``causal_offset``: query i may see keys 0..causal_offset+i (None = all keys).  k = 0 gives (0, -inf)."""
    Q, H, R = q_abs.shape
    if lat.shape[0] == 0:
        return (torch.zeros(Q, H, R, dtype=torch.float32, device=q_abs.device),
                torch.full((Q, H), float("-inf"), dtype=torch.float32, device=q_abs.device))
    lf = lat.float()
    kc, kp = lf[:, :kv_lora_rank], lf[:, kv_lora_rank:]
    att = (torch.einsum("qhr,kr->qhk", q_abs.float(), kc) + torch.einsum("qhp,kp->qhk", q_pe.float(), kp)) * scale
    if causal_offset is not None:
        qi = torch.arange(Q, device=att.device)[:, None, None]
        kj = torch.arange(lat.shape[0], device=att.device)[None, None, :]
        att = att.masked_fill(kj > qi + causal_offset, float("-inf"))
    lse = torch.logsumexp(att, dim=-1)
    p = torch.exp(att - lse[..., None])
    return torch.einsum("qhk,kr->qhr", p, kc), lse


def partial_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: float,
                      causal_offset: Optional[int] = None, soft_cap: float = 0.0) -> tuple[torch.Tensor, torch.Tensor]:
    """Attention of q [q, H, D] over k / v [n, Hkv, D] (H a multiple of Hkv, GQA) without normalising across calls:
    returns (out [q, H, D] fp32, lse [q, H]) so that partial results over disjoint key sets merge exactly.
    ``causal_offset``: query i may see keys 0..causal_offset+i (None = all keys).  n = 0 gives (0, -inf)."""
    Q, H, D = q.shape
    n, Hkv = k.shape[0], k.shape[1]
    if n == 0:
        return (torch.zeros(Q, H, D, dtype=torch.float32, device=q.device),
                torch.full((Q, H), float("-inf"), dtype=torch.float32, device=q.device))
    g = H // Hkv
    kf = k.float().repeat_interleave(g, dim=1)
    vf = v.float().repeat_interleave(g, dim=1)
    att = torch.einsum("qhd,khd->qhk", q.float(), kf) * scale
    if soft_cap:
        att = soft_cap * torch.tanh(att / soft_cap)
    if causal_offset is not None:
        qi = torch.arange(Q, device=att.device)[:, None, None]
        kj = torch.arange(n, device=att.device)[None, None, :]
        att = att.masked_fill(kj > qi + causal_offset, float("-inf"))
    lse = torch.logsumexp(att, dim=-1)
    p = torch.exp(att - lse[..., None])
    return torch.einsum("qhk,khd->qhd", p, vf), lse


# -------------------------------------------------------------- sampling
def apply_penalties(logits: torch.Tensor, prompt_mask: torch.Tensor, output_counts: torch.Tensor,
                    presence: torch.Tensor, frequency: torch.Tensor, repetition: torch.Tensor) -> torch.Tensor:
    """logits [B,V] fp32; prompt_mask [B,V] bool; output_counts [B,V] int."""
    seen = prompt_mask | (output_counts > 0)
    rep = repetition[:, None].expand_as(logits)
    logits = torch.where(seen & (logits > 0), logits / rep, logits)
    logits = torch.where(seen & (logits <= 0), logits * rep, logits)
    logits = logits - frequency[:, None] * output_counts.to(logits.dtype)
    logits = logits - presence[:, None] * (output_counts > 0).to(logits.dtype)
    return logits


def apply_top_k_top_p_min_p(logits: torch.Tensor, top_k: Optional[torch.Tensor], top_p: Optional[torch.Tensor],
                            min_p: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Mask filtered logits to -inf. top_k: int [B] (0 = off), top_p/min_p: float [B]."""
    if top_k is not None or top_p is not None:
        srt, idx = logits.sort(dim=-1, descending=False)
        V = logits.shape[-1]
        if top_k is not None:
            k = torch.where(top_k <= 0, torch.full_like(top_k, V), top_k).clamp(max=V).long()
            thresh = srt.gather(1, (V - k)[:, None])
            srt = srt.masked_fill(srt < thresh, float("-inf"))
        if top_p is not None:
            probs = srt.softmax(dim=-1)
            cum = probs.cumsum(dim=-1)
            mask = cum <= (1.0 - top_p)[:, None]
            mask[:, -1] = False
            srt = srt.masked_fill(mask, float("-inf"))
        logits = torch.empty_like(srt).scatter_(1, idx, srt)
    if min_p is not None:
        probs = logits.softmax(dim=-1)
        thresh = probs.max(dim=-1, keepdim=True).values * min_p[:, None]
        logits = logits.masked_fill(probs < thresh, float("-inf"))
    return logits


def gumbel_sample(logits: torch.Tensor, temperature: torch.Tensor, generators: Optional[dict] = None,
                  noise: Optional[torch.Tensor] = None) -> torch.Tensor:
    """argmax(logits/T + Gumbel). temperature [B] (0 → greedy). ``noise``: optional uniform(0,1) [B,V]."""
    greedy = temperature < 1e-5
    t = torch.where(greedy, torch.ones_like(temperature), temperature)
    scaled = logits / t[:, None]
    if noise is None:
        noise = torch.empty_like(scaled)
        if generators:
            noise.uniform_()
            for i, g in generators.items():
                noise[i].uniform_(generator=g)
        else:
            noise.uniform_()
    g = -torch.log(-torch.log(noise.clamp(1e-10, 1.0 - 1e-7)))
    sampled = (scaled + g).argmax(dim=-1)
    return torch.where(greedy, logits.argmax(dim=-1), sampled)


def compute_logprobs(logits: torch.Tensor, sampled: torch.Tensor, k: int
                     ) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Returns (ids [B,k+1], logprobs [B,k+1], rank [B]); column 0 is the sampled token."""
    lp = logits.float().log_softmax(dim=-1)
    s_lp = lp.gather(1, sampled[:, None].long())
    rank = (lp > s_lp).sum(dim=-1) + 1
    if k > 0:
        tv, ti = lp.topk(k, dim=-1)
        ids = torch.cat((sampled[:, None].long(), ti), dim=1)
        vals = torch.cat((s_lp, tv), dim=1)
    else:
        ids, vals = sampled[:, None].long(), s_lp
    return ids, vals, rank


# ------------------------------------------------------------------- MoE
def topk_softmax(gating: torch.Tensor, k: int, renormalize: bool) -> tuple[torch.Tensor, torch.Tensor]:
    p = gating.float().softmax(dim=-1)
    w, ids = p.topk(k, dim=-1)
    if renormalize:
        w = w / w.sum(dim=-1, keepdim=True)
    return w, ids.int()


def grouped_topk(scores_in: torch.Tensor, k: int, n_group: int, topk_group: int, renormalize: bool,
                 scoring: str = "sigmoid", bias: Optional[torch.Tensor] = None,
                 routed_scaling_factor: float = 1.0) -> tuple[torch.Tensor, torch.Tensor]:
    """DeepSeek group-limited routing."""
    s = scores_in.float().sigmoid() if scoring == "sigmoid" else scores_in.float().softmax(-1)
    T, E = s.shape
    sel = s + bias.float()[None] if bias is not None else s
    if n_group > 1:
        g = sel.view(T, n_group, E // n_group)
        gs = g.topk(2, dim=-1).values.sum(-1) if bias is not None else g.max(-1).values
        gi = gs.topk(topk_group, dim=-1).indices
        gm = torch.zeros_like(gs).scatter_(1, gi, 1.0)
        mask = gm[:, :, None].expand(T, n_group, E // n_group).reshape(T, E)
        sel = sel.masked_fill(mask == 0, float("-inf"))
    ids = sel.topk(k, dim=-1).indices
    w = s.gather(1, ids)
    if renormalize:
        w = w / (w.sum(-1, keepdim=True) + 1e-20)
    return w * routed_scaling_factor, ids.int()


def fused_moe(x: torch.Tensor, w13: torch.Tensor, w2: torch.Tensor, topk_w: torch.Tensor,
              topk_ids: torch.Tensor, act: str = "silu", expert_map: Optional[torch.Tensor] = None
              ) -> torch.Tensor:
    """x [T,H]; w13 [E,2I,H] (gate rows then up rows); w2 [E,H,I]. Returns [T,H]."""
    T, H = x.shape
    E = w13.shape[0]
    out = torch.zeros(T, H, dtype=torch.float32, device=x.device)
    ids = topk_ids.long()
    if expert_map is not None:
        ids = expert_map[ids].long()
    for e in range(E):
        tok, slot = (ids == e).nonzero(as_tuple=True)
        if tok.numel() == 0:
            continue
        h = x[tok].float() @ w13[e].float().t()
        I = h.shape[-1] // 2
        a = F.silu(h[:, :I]) if act == "silu" else F.gelu(h[:, :I])
        h = (a * h[:, I:]).to(x.dtype).float()
        y = h @ w2[e].float().t()
        out.index_add_(0, tok, y * topk_w[tok, slot].float()[:, None])
    return out.to(x.dtype)


# ---------------------------------------------------------- cache maintenance
def swap_blocks(src: torch.Tensor, dst: torch.Tensor, mapping: list[tuple[int, int]]) -> None:
    """Copy whole KV blocks src[i] → dst[j] (host tier ↔ device, or block migration)."""
    if not mapping:
        return
    si = torch.tensor([m[0] for m in mapping], dtype=torch.long, device=src.device)
    di = torch.tensor([m[1] for m in mapping], dtype=torch.long, device=dst.device)
    dst[di] = src[si].to(dst.device, dst.dtype)


def convert_fp8(x: torch.Tensor, scale: float = 1.0, to_fp8: bool = True) -> torch.Tensor:
    if to_fp8:
        return (x.float() / scale).clamp(-FP8_MAX, FP8_MAX).to(torch.float8_e4m3fn)
    return x.float() * scale


# ---------------------------------------------------------- quantisation
FP8_MAX = 448.0


def scaled_fp8_quant(x: torch.Tensor, scale: Optional[torch.Tensor] = None, per_token: bool = False
                     ) -> tuple[torch.Tensor, torch.Tensor]:
    """Static (given scale) or dynamic per-tensor / per-token fp8 e4m3 activation quantisation."""
    xf = x.float()
    if scale is None:
        amax = xf.abs().amax(dim=-1, keepdim=True) if per_token else xf.abs().amax()
        scale = (amax / FP8_MAX).clamp_min(1e-12)
    q = (xf / scale).clamp(-FP8_MAX, FP8_MAX).to(torch.float8_e4m3fn)
    return q, scale.float()


def scaled_int8_quant(x: torch.Tensor, scale: Optional[torch.Tensor] = None) -> tuple[torch.Tensor, torch.Tensor]:
    xf = x.float()
    if scale is None:
        scale = (xf.abs().amax(dim=-1, keepdim=True) / 127.0).clamp_min(1e-12)
    return (xf / scale).round().clamp(-128, 127).to(torch.int8), scale.float()


def scaled_mm(a: torch.Tensor, b: torch.Tensor, scale_a: torch.Tensor, scale_b: torch.Tensor, out_dtype,
              bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(a · bᵀ) with per-tensor / per-token (a) and per-tensor / per-channel (b) scales. a [M,K], b [N,K] (fp8/int8)."""
    y = (a.float() @ b.float().t()) * scale_a.reshape(-1, 1) * scale_b.reshape(1, -1)
    if bias is not None:
        y = y + bias.float()
    return y.to(out_dtype)


E2M1_VALUES = (0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0)


def fp4_dequant(packed: torch.Tensor, block_scale: torch.Tensor, global_scale, block: int = 16) -> torch.Tensor:
    """NVFP4 / MXFP4: packed uint8 [rows, cols/2] (low nibble first), one scale per ``block`` columns."""
    lut = torch.tensor(E2M1_VALUES + tuple(-v for v in E2M1_VALUES), dtype=torch.float32, device=packed.device)
    lo, hi = packed & 0xF, packed >> 4
    vals = torch.stack((lut[lo.long()], lut[hi.long()]), dim=-1).reshape(packed.shape[0], -1)
    sc = block_scale.float().repeat_interleave(block, dim=1)[:, :vals.shape[1]]
    return vals * sc * float(global_scale)


def mxfp4_quant(w: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """OCP MXFP4: e2m1 codes (two per byte, low nibble first) with one power-of-two e8m0 scale per 32 values
    (the GPT-OSS expert format).  Returns (packed uint8 [R, C/2], scales uint8 [R, C/32])."""
    wf = w.float()
    R, C = wf.shape
    g = wf.reshape(R, C // 32, 32)
    e = torch.ceil(torch.log2(g.abs().amax(-1).clamp_min(2.0 ** -120) / 6.0)).clamp(-126, 127)
    x = g / torch.exp2(e)[..., None]
    lut = torch.tensor(E2M1_VALUES, dtype=torch.float32, device=w.device)
    idx = (x.abs()[..., None] - lut).abs().argmin(-1)
    code = (idx + 8 * (x < 0).long()).reshape(R, C).to(torch.uint8)
    return (code[:, 0::2] | (code[:, 1::2] << 4)).contiguous(), (e + 127).to(torch.uint8)


def mxfp4_dequant(packed: torch.Tensor, scales: torch.Tensor) -> torch.Tensor:
    return fp4_dequant(packed, torch.exp2(scales.float() - 127.0), 1.0, 32)


def fp4_quant(w: torch.Tensor, block: int = 16) -> tuple[torch.Tensor, torch.Tensor, float]:
    """Reference NVFP4 quantiser: per-16 block scale (stored fp8 e4m3) + global fp32 scale → packed nibbles."""
    wf = w.float()
    R, C = wf.shape
    g = wf.reshape(R, C // block, block)
    amax = g.abs().amax(-1)
    gscale = float(amax.max() / (6.0 * FP8_MAX)) or 1.0
    bscale = (amax / 6.0 / gscale).clamp_min(1e-12).to(torch.float8_e4m3fn)
    x = g / (bscale.float()[..., None] * gscale)
    lut = torch.tensor(E2M1_VALUES, dtype=torch.float32, device=w.device)
    idx = (x.abs()[..., None] - lut).abs().argmin(-1)
    code = (idx + 8 * (x < 0).long()).reshape(R, C).to(torch.uint8)
    return (code[:, 0::2] | (code[:, 1::2] << 4)).contiguous(), bscale, gscale



def per_token_group_quant_fp8(x: torch.Tensor, group: int = 128) -> tuple[torch.Tensor, torch.Tensor]:
    """Returns (q [T, K], scale [T, ceil(K/group)]); a ragged last group is handled by zero padding."""
    T, K = x.shape
    G = (K + group - 1) // group
    xf = x.float()
    if G * group != K:
        xf = F.pad(xf, (0, G * group - K))
    xg = xf.view(T, G, group)
    amax = xg.abs().amax(-1).clamp_min(1e-10)
    scale = amax / FP8_MAX
    q = (xg / scale[..., None]).clamp(-FP8_MAX, FP8_MAX).to(torch.float8_e4m3fn).view(T, G * group)[:, :K]
    return q.contiguous(), scale


def block_quant_weight_fp8(w: torch.Tensor, block: int = 128) -> tuple[torch.Tensor, torch.Tensor]:
    N, K = w.shape
    nb, kb = (N + block - 1) // block, (K + block - 1) // block
    wp = torch.zeros(nb * block, kb * block, dtype=torch.float32, device=w.device)
    wp[:N, :K] = w.float()
    wb = wp.view(nb, block, kb, block)
    amax = wb.abs().amax(dim=(1, 3)).clamp_min(1e-10)
    scale = amax / FP8_MAX
    q = (wb / scale[:, None, :, None]).clamp(-FP8_MAX, FP8_MAX).view(nb * block, kb * block)[:N, :K]
    return q.to(torch.float8_e4m3fn).contiguous(), scale


def block_fp8_linear(x: torch.Tensor, wq: torch.Tensor, w_scale: torch.Tensor, block: int = 128,
                     bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Activation quantised per (token, 128-group), weight per 128x128 block; fp32 accumulate."""
    xq, xs = per_token_group_quant_fp8(x, block)
    N, K = wq.shape
    xd = xq.float() * xs.repeat_interleave(block, dim=1)[:, :K]
    ws = w_scale.repeat_interleave(block, 0)[:N].repeat_interleave(block, 1)[:, :K]
    y = xd @ (wq.float() * ws).t()
    if bias is not None:
        y = y + bias.float()
    return y.to(x.dtype)


# ---------------------------------------------------------------------------
# Mamba-2 state-space duality (reference: layers/mamba/ops/ssd_combined.py, ssd_chunk_scan.py, ssd_state_passing.py)
def _segsum(x: torch.Tensor) -> torch.Tensor:
    """[..., L] → [..., L, L] with out[i, j] = sum_{j < k <= i} x[k] for j <= i, -inf above the diagonal."""
    L = x.shape[-1]
    cs = torch.cumsum(x, dim=-1)
    d = cs[..., :, None] - cs[..., None, :]
    mask = torch.tril(torch.ones(L, L, dtype=torch.bool, device=x.device))
    return d.masked_fill(~mask, float("-inf"))


def ssd_chunk_scan(x: torch.Tensor, dt: torch.Tensor, A: torch.Tensor, B: torch.Tensor, C: torch.Tensor,
                   chunk: int, h0: Optional[torch.Tensor] = None) -> tuple[torch.Tensor, torch.Tensor]:
    """One sequence.  x [T, H, P], dt [T, H] (post-softplus), A [H] (negative), B / C [T, G, N] (G groups of heads),
    h0 [H, P, N] → (y [T, H, P] without the D skip, final state [H, P, N]).  Work is quadratic only inside a chunk:
    intra-chunk outputs through the masked decay matrix, inter-chunk through the carried state."""
    T, H, P = x.shape
    G, N = B.shape[1], B.shape[2]
    rep = H // G
    x, dt, B, C = x.float(), dt.float(), B.float(), C.float()
    pad = (-T) % chunk
    if pad:
        x = torch.cat([x, x.new_zeros(pad, H, P)])
        dt = torch.cat([dt, dt.new_zeros(pad, H)])
        B = torch.cat([B, B.new_zeros(pad, G, N)])
        C = torch.cat([C, C.new_zeros(pad, G, N)])
    nc = (T + pad) // chunk
    Bh = B.repeat_interleave(rep, dim=1).view(nc, chunk, H, N)
    Ch = C.repeat_interleave(rep, dim=1).view(nc, chunk, H, N)
    X = (x * dt[..., None]).view(nc, chunk, H, P)
    a = (dt * A.float()[None]).view(nc, chunk, H).permute(2, 0, 1)                # [H, nc, L] log-decay per step
    a_cs = torch.cumsum(a, dim=-1)
    Lm = torch.exp(_segsum(a))                                                     # [H, nc, L, L]
    y_diag = torch.einsum("clhn,cshn,hcls,cshp->clhp", Ch, Bh, Lm, X)
    decay_states = torch.exp(a_cs[:, :, -1:] - a_cs)                               # [H, nc, L]
    states = torch.einsum("clhn,hcl,clhp->chpn", Bh, decay_states, X)              # state each chunk adds
    init = h0.float()[None] if h0 is not None else states.new_zeros(1, H, P, N)
    states = torch.cat([init, states], dim=0)                                      # [nc+1, H, P, N]
    chunk_decay = torch.exp(_segsum(F.pad(a_cs[:, :, -1], (1, 0))))                # [H, nc+1, nc+1]
    new_states = torch.einsum("hzc,chpn->zhpn", chunk_decay, states)
    prev, final = new_states[:-1], new_states[-1]
    y_off = torch.einsum("clhn,chpn,hcl->clhp", Ch, prev, torch.exp(a_cs))
    y = (y_diag + y_off).reshape(nc * chunk, H, P)[:T]
    return y, final
