"""Paged attention layer.

Behavioural reference: vllm/model_executor/layers/attention/attention.py:218
(``Attention.forward`` :478-571 = KV-cache update + unified attention op,
``get_kv_cache_spec`` :597).  Here the layer owns a fused
RoPE + KV-write + attention call sequence over our own cache layout
``[num_blocks, 2, Hkv, block_size, D]``.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from vllm_b200 import envs, ops
from vllm_b200.core.kv_cache_manager import FullAttentionSpec, KVCacheSpec, MLAAttentionSpec, SlidingWindowSpec
from vllm_b200.forward_context import get_forward_context

_LAYER_REGISTRY: dict[str, "Attention"] = {}


def attention_layers() -> dict[str, "Attention"]:
    return _LAYER_REGISTRY


def clear_attention_layers() -> None:
    _LAYER_REGISTRY.clear()


class Attention(nn.Module):
    cascade_calls = 0          # how many layer invocations took the cascade path (tests / debugging)

    def __init__(self, num_heads: int, head_size: int, scale: float, num_kv_heads: Optional[int] = None,
                 sliding_window: Optional[int] = None, logits_soft_cap: float = 0.0, prefix: str = "",
                 causal: bool = True, rope: Optional[nn.Module] = None,
                 alibi_slopes: Optional[torch.Tensor] = None, sinks: Optional[torch.Tensor] = None,
                 chunk_size: int = 0) -> None:
        super().__init__()
        self.chunk_size = int(chunk_size or 0)        # chunked-local attention (Llama-4 rope layers)
        # ALiBi slopes / attention-sink logits of THIS rank's heads (buffers follow the module's device)
        self.register_buffer("alibi_slopes", None if alibi_slopes is None else alibi_slopes.float(), persistent=False)
        self.sinks = sinks                       # nn.Parameter owned by the model (loaded from the checkpoint) or None
        self.num_heads, self.head_size, self.scale = num_heads, head_size, scale
        self.num_kv_heads = num_kv_heads or num_heads
        self.sliding_window = sliding_window
        self.logits_soft_cap = logits_soft_cap
        self.causal = causal
        self.layer_name = prefix
        self.rope = rope
        assert prefix not in _LAYER_REGISTRY, f"duplicate attention layer name {prefix}"
        _LAYER_REGISTRY[prefix] = self

    def get_kv_cache_spec(self, block_size: int, dtype_size: int) -> KVCacheSpec:
        from vllm_b200.parallel.state import get_dcp_group
        d = get_dcp_group().world_size
        if d > 1:
            if self.sliding_window or self.alibi_slopes is not None or self.sinks is not None or self.chunk_size:
                raise NotImplementedError("decode context parallelism covers full causal attention (no sliding window / "
                                          "ALiBi / sinks / chunked-local layers)")
            return FullAttentionSpec(block_size=block_size * d, num_kv_heads=self.num_kv_heads, head_size=self.head_size,
                                     dtype_size=dtype_size, dcp=d)
        cls = SlidingWindowSpec if self.sliding_window else FullAttentionSpec
        return cls(block_size=block_size, num_kv_heads=self.num_kv_heads, head_size=self.head_size,
                   dtype_size=dtype_size, sliding_window=self.sliding_window)

    def _forward_dcp(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, positions, md, kv_cache) -> torch.Tensor:
        """Decode context parallelism for GQA / MQA caches (reference: config/parallel.py:342-388, v1/attention/ops/dcp.py:
        241-338).  It applies where tensor parallelism REPLICATES KV heads (tp > num_kv_heads): the d ranks that hold
        replicas of the same heads stop storing the same tokens d times and shard the cache by position instead (token p
        lives on rank p % d), so the group holds d x more context.  Per step: the ranks all-gather their query heads, each
        attends over its shard of the *previous* context for all of them, the partial results are merged by their
        log-sum-exp across the group; the tokens of the current step are attended locally (their K / V were just computed
        on every rank)."""
        from vllm_b200.ops import ref
        from vllm_b200.parallel.state import get_dcp_group
        g = get_dcp_group()
        d, r = g.world_size, g.rank_in_group
        T, D, h, hkv = q.shape[0], self.head_size, self.num_heads, self.num_kv_heads
        if self.rope is not None and positions is not None:
            q, k = self.rope(positions, q, k)
        k3, v3 = k.reshape(T, hkv, D), v.reshape(T, hkv, D)
        ops.reshape_and_cache(k3, v3, kv_cache, md.slot_mapping)               # positions owned elsewhere carry slot -1
        q3 = q.reshape(T, h, D)
        # gathered query heads: [T, d, hkv, h/hkv, D] → regrouped so that the heads of one KV head stay contiguous
        q_all = g.all_gather(q3.reshape(T, 1, hkv, h // hkv, D).contiguous(), dim=1)
        q_all = q_all.permute(0, 2, 1, 3, 4).reshape(T, d * h, D)
        qsl = (md.query_start_loc_cpu if md.query_start_loc_cpu is not None else md.query_start_loc.cpu()).tolist()
        sls = (md.seq_lens_cpu if md.seq_lens_cpu is not None else md.seq_lens.cpu()).tolist()
        bs = kv_cache.shape[3]
        o_ctx = torch.zeros(T, d * h, D, dtype=torch.float32, device=q.device)
        lse_ctx = torch.full((T, d * h), float("-inf"), dtype=torch.float32, device=q.device)
        o_new = torch.zeros(T, h, D, dtype=torch.float32, device=q.device)
        lse_new = torch.full((T, h), float("-inf"), dtype=torch.float32, device=q.device)
        for i in range(md.num_reqs):
            s, e = qsl[i], qsl[i + 1]
            if e <= s:
                continue
            ctx = sls[i] - (e - s)                                             # positions before this step's tokens
            n_loc = (ctx - r + d - 1) // d if ctx > r else 0                   # ... of which this rank stores n_loc
            if n_loc > 0:
                nb = (n_loc + bs - 1) // bs
                blk = kv_cache[md.block_table[i, :nb].long()]                  # [nb, 2, hkv, bs, D]
                kc = blk[:, 0].permute(0, 2, 1, 3).reshape(nb * bs, hkv, D)[:n_loc]
                vc = blk[:, 1].permute(0, 2, 1, 3).reshape(nb * bs, hkv, D)[:n_loc]
                o_ctx[s:e], lse_ctx[s:e] = ref.partial_attention(q_all[s:e], kc, vc, self.scale, None, self.logits_soft_cap)
            o_new[s:e], lse_new[s:e] = ref.partial_attention(q3[s:e], k3[s:e], v3[s:e], self.scale, 0, self.logits_soft_cap)
        # merge the context partials of the d ranks (softmax-weighted by their lse), keep this rank's heads
        lse_all = g.all_gather(lse_ctx[..., None].contiguous(), dim=-1)       # [T, d*h, d]
        lse_tot = torch.logsumexp(lse_all, dim=-1)
        w = torch.exp(lse_ctx - lse_tot).nan_to_num(0.0)                      # ranks without context contribute 0
        o_all = g.all_reduce((o_ctx * w[..., None]).contiguous())
        mine = o_all.reshape(T, hkv, d, h // hkv, D)[:, :, r].reshape(T, h, D)
        lse_c = lse_tot.reshape(T, hkv, d, h // hkv)[:, :, r].reshape(T, h)
        m = torch.maximum(lse_c, lse_new)
        wc, wn = torch.exp(lse_c - m).nan_to_num(0.0), torch.exp(lse_new - m)
        out = (mine * wc[..., None] + o_new * wn[..., None]) / (wc + wn)[..., None]
        return out.to(q.dtype).reshape(T, h * D)

    def forward(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor,
                positions: Optional[torch.Tensor] = None, qk_norm: Optional[tuple] = None) -> torch.Tensor:
        """q: [T, Hq*D], k/v: [T, Hkv*D] → [T, Hq*D]. If the layer has a ``rope``,
        rotation and the KV write happen in one fused kernel.  ``qk_norm`` = (q_norm, k_norm) per-head RMSNorm modules
        (Qwen3 / Gemma-3 / OLMo-2 style): on CUDA they are fused with the rotation and the cache insert into ONE launch
        (csrc/qknorm_rope.cu), otherwise applied here before the rotation."""
        ctx = get_forward_context()
        T = q.shape[0]
        D = self.head_size
        md = ctx.metadata_for(self.layer_name)
        kv_cache = ctx.kv_caches.get(self.layer_name)
        if md is not None and kv_cache is None:
            raise RuntimeError(f"attention layer {self.layer_name!r} has no KV cache bound (KV-cache config does not "
                               "cover this worker's layers)")
        if md is None or kv_cache is None:
            # memory-profiling run: no cache; keep the data flow (and the GEMM shapes) intact
            if self.rope is not None and positions is not None:
                q, k = self.rope(positions, q, k)
            return q.clone() if q.shape[-1] == self.num_heads * D else q
        from vllm_b200.parallel.state import get_dcp_group
        if get_dcp_group().world_size > 1:
            if qk_norm is not None:
                qn, kn = qk_norm
                q = qn(q.reshape(T, -1, D)).reshape(T, -1)
                k = kn(k.reshape(T, -1, D)).reshape(T, -1)
            return self._forward_dcp(q, k, v, positions, md, kv_cache)
        pcp = ctx.extra.get("pcp")
        mpos = ctx.extra.get("mrope_positions") if (self.rope is not None and self.rope.mrope_section) else None
        if qk_norm is not None:
            qn, kn = qk_norm
            fused = (md is not None and kv_cache is not None and q.is_cuda and pcp is None and mpos is None
                     and self.rope is not None and positions is not None and q.dtype in (torch.bfloat16, torch.float16)
                     and kv_cache.dtype == q.dtype and getattr(qn, "gemma", False) == getattr(kn, "gemma", False)
                     and qn.eps == kn.eps and not envs.VLLM_B200_FORCE_REF_OPS
                     and ops.cuda_ops().qknorm_rope_ok(D, self.rope.cos_sin_cache.shape[-1], self.rope.is_neox_style))
            if fused:
                ops.cuda_ops().fused_qk_norm_rope(positions, q, k, qn.weight, kn.weight, self.rope.cos_sin_cache, D, qn.eps,
                                                  self.rope.is_neox_style, v=v, kv_cache=kv_cache,
                                                  slot_mapping=md.slot_mapping, gemma=getattr(qn, "gemma", False),
                                                  vt=md.kv_vt)
                positions = None            # rotation and cache insert are done: skip them below
                skip_cache_write = True
            else:
                q = qn(q.reshape(T, -1, D)).reshape(T, -1)
                k = kn(k.reshape(T, -1, D)).reshape(T, -1)
                skip_cache_write = False
        else:
            skip_cache_write = False
        if skip_cache_write:
            pass
        elif mpos is not None:
            # multimodal rotary ids differ per channel section for image tokens: rotate un-fused, then the ordinary
            # cache write (text-only steps take the fused path below with 1-D positions = index + per-request delta)
            q, k = self.rope.forward_mrope(mpos, q, k)
            ops.reshape_and_cache(k.view(T, self.num_kv_heads, D), v.view(T, self.num_kv_heads, D), kv_cache,
                                  md.slot_mapping, vt=md.kv_vt)
        elif pcp is not None:
            # prefill context parallelism: this rank holds a share of the step's tokens.  Rotate locally, all-gather
            # K / V of the whole step (global token order) so that every rank's cache stays complete, attend for the
            # local queries only (metadata describes them as the tails of shorter sequences).
            if self.rope is not None and positions is not None:
                q, k = self.rope(positions, q, k)
            Tg = pcp["num_tokens"]
            k_all, v_all = pcp["gather"](k.contiguous()), pcp["gather"](v.contiguous())
            ops.reshape_and_cache(k_all.view(Tg, self.num_kv_heads, D), v_all.view(Tg, self.num_kv_heads, D), kv_cache,
                                  pcp["slot_mappings"][ctx.layer_to_group.get(self.layer_name, 0)], vt=md.kv_vt)
        elif self.rope is not None and positions is not None:
            q, k = ops.rope_and_cache(positions, q, k, v, self.rope.cos_sin_cache, D,
                                      self.rope.is_neox_style, kv_cache, md.slot_mapping, vt=md.kv_vt)
        else:
            ops.reshape_and_cache(k.view(T, self.num_kv_heads, D), v.view(T, self.num_kv_heads, D),
                                  kv_cache, md.slot_mapping, vt=md.kv_vt)
        if (md.cascade is not None and self.causal and not self.sliding_window and self.alibi_slopes is None
                and self.sinks is None and not self.chunk_size):
            # shared prefix: its K/V are read once for the whole batch (all queries sit after it → no causal mask),
            # the per-request remainders are attended separately, the two partial softmaxes are merged by their LSE
            md_p, md_s = md.cascade
            q3 = q.view(T, self.num_heads, D)
            o1, l1 = ops.paged_attention(q3, kv_cache, md_p, self.scale, None, self.logits_soft_cap, False, return_lse=True)
            o2, l2 = ops.paged_attention(q3, kv_cache, md_s, self.scale, None, self.logits_soft_cap, True, return_lse=True)
            out, _ = ops.merge_attn_states(o1, l1, o2, l2)
            Attention.cascade_calls += 1
            return out.view(T, self.num_heads * D)
        out = ops.paged_attention(q.view(T, self.num_heads, D), kv_cache, md, self.scale,
                                  self.sliding_window, self.logits_soft_cap, self.causal,
                                  alibi_slopes=self.alibi_slopes, sinks=self.sinks, chunk_size=self.chunk_size)
        return out.view(T, self.num_heads * D)


class MLAAttentionCore(nn.Module):
    """Latent-cache attention core for DeepSeek-style MLA (reference algorithm doc:
    layers/attention/mla_attention.py:1-200). Decode and prefill both run as MQA over
    the 576-d latent (absorbed W_UK / W_UV), which is also what the sm_100 CUTLASS-MLA
    decode kernel in the reference does."""

    def __init__(self, num_heads: int, kv_lora_rank: int, qk_rope_head_dim: int, scale: float,
                 prefix: str = "") -> None:
        super().__init__()
        self.num_heads, self.kv_lora_rank, self.qk_rope_head_dim = num_heads, kv_lora_rank, qk_rope_head_dim
        self.scale = scale
        self.layer_name = prefix
        self.head_size = kv_lora_rank + qk_rope_head_dim
        self.num_kv_heads = 1
        self.sliding_window = None
        _LAYER_REGISTRY[prefix] = self  # type: ignore[assignment]

    def get_kv_cache_spec(self, block_size: int, dtype_size: int) -> KVCacheSpec:
        from vllm_b200.parallel.state import get_dcp_group
        d = get_dcp_group().world_size
        return MLAAttentionSpec(block_size=block_size * d, head_size=self.head_size, dtype_size=dtype_size, dcp=d)

    def _forward_dcp(self, q_nope_abs, q_pe, kv_c, k_pe, md, kv_cache) -> torch.Tensor:
        """Decode context parallelism (reference: config/parallel.py:342-388, v1/attention/ops/dcp.py): the latent
        cache — identical on every TP rank without DCP — is sharded by position (token p lives on rank p % d), so
        the group holds d x more context.  Queries of all ranks are gathered, every rank attends over its shard of
        the *previous* context, partial results are merged with their log-sum-exp across the group; the tokens of
        the current step are attended locally (their latents were just computed on every rank)."""
        from vllm_b200.ops import ref
        from vllm_b200.parallel.state import get_dcp_group
        g = get_dcp_group()
        d, r = g.world_size, g.rank_in_group
        R = self.kv_lora_rank
        ops.concat_and_cache_mla(kv_c, k_pe, kv_cache, md.slot_mapping)       # non-owned positions carry slot -1
        T, h, _ = q_nope_abs.shape
        qn_all = g.all_gather(q_nope_abs.contiguous(), dim=1)                 # [T, d*h, R]
        qp_all = g.all_gather(q_pe.contiguous(), dim=1)
        qsl = (md.query_start_loc_cpu if md.query_start_loc_cpu is not None else md.query_start_loc.cpu()).tolist()
        sls = (md.seq_lens_cpu if md.seq_lens_cpu is not None else md.seq_lens.cpu()).tolist()
        bs = kv_cache.shape[1]
        o_ctx = torch.zeros(T, d * h, R, dtype=torch.float32, device=q_nope_abs.device)
        lse_ctx = torch.full((T, d * h), float("-inf"), dtype=torch.float32, device=q_nope_abs.device)
        o_new = torch.zeros(T, h, R, dtype=torch.float32, device=q_nope_abs.device)
        lse_new = torch.full((T, h), float("-inf"), dtype=torch.float32, device=q_nope_abs.device)
        new_lat = torch.cat((kv_c, k_pe), dim=-1)
        for i in range(md.num_reqs):
            s, e = qsl[i], qsl[i + 1]
            if e <= s:
                continue
            ctx = sls[i] - (e - s)                                             # positions before this step's tokens
            n_loc = (ctx - r + d - 1) // d if ctx > r else 0                   # ... of which this rank stores n_loc
            if n_loc > 0:
                nb = (n_loc + bs - 1) // bs
                lat = kv_cache[md.block_table[i, :nb].long()].reshape(nb * bs, -1)[:n_loc]
                o_ctx[s:e], lse_ctx[s:e] = ref.mla_partial_attention(qn_all[s:e], qp_all[s:e], lat, self.scale, R)
            o_new[s:e], lse_new[s:e] = ref.mla_partial_attention(q_nope_abs[s:e], q_pe[s:e], new_lat[s:e], self.scale, R,
                                                                 causal_offset=0)
        # merge the context partials of the d ranks: softmax-weighted by their lse
        lse_all = g.all_gather(lse_ctx[..., None].contiguous(), dim=-1)       # [T, d*h, d]
        lse_tot = torch.logsumexp(lse_all, dim=-1)
        w = torch.exp(lse_ctx - lse_tot).nan_to_num(0.0)                      # ranks without context contribute 0
        o_all = g.all_reduce((o_ctx * w[..., None]).contiguous())
        o_c, lse_c = o_all[:, r * h:(r + 1) * h], lse_tot[:, r * h:(r + 1) * h]
        # merge context and current-step parts for this rank's heads
        m = torch.maximum(lse_c, lse_new)
        wc, wn = torch.exp(lse_c - m).nan_to_num(0.0), torch.exp(lse_new - m)
        out = (o_c * wc[..., None] + o_new * wn[..., None]) / (wc + wn)[..., None]
        return out.to(q_nope_abs.dtype)

    def forward(self, q_nope_abs: torch.Tensor, q_pe: torch.Tensor, kv_c: torch.Tensor, k_pe: torch.Tensor
                ) -> torch.Tensor:
        ctx = get_forward_context()
        md = ctx.metadata_for(self.layer_name)
        kv_cache = ctx.kv_caches.get(self.layer_name)
        if md is not None and kv_cache is None:
            raise RuntimeError(f"MLA layer {self.layer_name!r} has no KV cache bound")
        if md is None or kv_cache is None:
            return torch.zeros_like(q_nope_abs)
        from vllm_b200.parallel.state import get_dcp_group
        if get_dcp_group().world_size > 1:
            return self._forward_dcp(q_nope_abs, q_pe, kv_c, k_pe, md, kv_cache)
        ops.concat_and_cache_mla(kv_c, k_pe, kv_cache, md.slot_mapping)
        return ops.mla_attention(q_nope_abs, q_pe, kv_cache, md, self.scale, self.kv_lora_rank)
