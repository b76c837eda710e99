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

from vllm_b200 import ops
from vllm_b200.core.kv_cache_manager import FullAttentionSpec, KVCacheSpec, MLAAttentionSpec, SlidingWindowSpec
from vllm_b200.forward_context import get_forward_context

_LAYER_REGISTRY: dict[str, "Attention"] = {}


def attention_layers() -> dict[str, "Attention"]:
    return _LAYER_REGISTRY


def clear_attention_layers() -> None:
    _LAYER_REGISTRY.clear()


class Attention(nn.Module):
    def __init__(self, num_heads: int, head_size: int, scale: float, num_kv_heads: Optional[int] = None,
                 sliding_window: Optional[int] = None, logits_soft_cap: float = 0.0, prefix: str = "",
                 causal: bool = True, rope: Optional[nn.Module] = None) -> None:
        super().__init__()
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
        cls = SlidingWindowSpec if self.sliding_window else FullAttentionSpec
        return cls(block_size=block_size, num_kv_heads=self.num_kv_heads, head_size=self.head_size,
                   dtype_size=dtype_size, sliding_window=self.sliding_window)

    def forward(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor,
                positions: Optional[torch.Tensor] = None) -> torch.Tensor:
        """q: [T, Hq*D], k/v: [T, Hkv*D] → [T, Hq*D]. If the layer has a ``rope``,
        rotation and the KV write happen in one fused kernel."""
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
        if self.rope is not None and positions is not None:
            q, k = ops.rope_and_cache(positions, q, k, v, self.rope.cos_sin_cache, D,
                                      self.rope.is_neox_style, kv_cache, md.slot_mapping, vt=md.kv_vt)
        else:
            ops.reshape_and_cache(k.view(T, self.num_kv_heads, D), v.view(T, self.num_kv_heads, D),
                                  kv_cache, md.slot_mapping, vt=md.kv_vt)
        out = ops.paged_attention(q.view(T, self.num_heads, D), kv_cache, md, self.scale,
                                  self.sliding_window, self.logits_soft_cap, self.causal)
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
        return MLAAttentionSpec(block_size=block_size, head_size=self.head_size, dtype_size=dtype_size)

    def forward(self, q_nope_abs: torch.Tensor, q_pe: torch.Tensor, kv_c: torch.Tensor, k_pe: torch.Tensor
                ) -> torch.Tensor:
        ctx = get_forward_context()
        md = ctx.metadata_for(self.layer_name)
        kv_cache = ctx.kv_caches.get(self.layer_name)
        if md is not None and kv_cache is None:
            raise RuntimeError(f"MLA layer {self.layer_name!r} has no KV cache bound")
        if md is None or kv_cache is None:
            return torch.zeros_like(q_nope_abs)
        ops.concat_and_cache_mla(kv_c, k_pe, kv_cache, md.slot_mapping)
        return ops.mla_attention(q_nope_abs, q_pe, kv_cache, md, self.scale, self.kv_lora_rank)
