"""Gemma-2: sandwich norms (pre + post around attention and MLP), attention / final logit soft-capping,
alternating sliding-window and global layers, ``query_pre_attn_scalar`` (reference: models/gemma2.py)."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from vllm_b200.config import VllmConfig
from vllm_b200.layers.attention import Attention
from vllm_b200.layers.common import RMSNorm
from vllm_b200.layers.linear import QKVParallelLinear, RowParallelLinear
from vllm_b200.layers.rotary import get_rope
from vllm_b200.models.llama import LlamaForCausalLM, LlamaMLP, LlamaModel


class Gemma2Attention(nn.Module):
    def __init__(self, cfg: dict, mc, dtype, quant_config, prefix: str, layer_idx: int) -> None:
        super().__init__()
        H, nh = cfg["hidden_size"], cfg["num_attention_heads"]
        nkv = cfg.get("num_key_value_heads") or nh
        hd = cfg.get("head_dim") or H // nh
        bias = bool(cfg.get("attention_bias", False))
        self.qkv_proj = QKVParallelLinear(H, hd, nh, nkv, bias, dtype, quant_config, f"{prefix}.qkv_proj")
        self.o_proj = RowParallelLinear(nh * hd, H, bias, dtype=dtype, quant_config=quant_config, prefix=f"{prefix}.o_proj")
        self.q_size, self.kv_size = self.qkv_proj.q_size, self.qkv_proj.kv_size
        types = cfg.get("layer_types")
        sliding = (types[layer_idx] == "sliding_attention") if types else (layer_idx % 2 == 0)
        max_pos = mc.max_model_len if mc.max_model_len > 0 else cfg.get("max_position_embeddings", 8192)
        rope = get_rope(hd, hd, max_pos, cfg.get("rope_theta", 10000.0), True, None)
        self.attn = Attention(self.qkv_proj.num_heads, hd, float(cfg.get("query_pre_attn_scalar", hd)) ** -0.5,
                              self.qkv_proj.num_kv_heads, sliding_window=cfg.get("sliding_window") if sliding else None,
                              logits_soft_cap=cfg.get("attn_logit_softcapping") or 0.0, prefix=f"{prefix}.attn", rope=rope,
                              causal=bool(cfg.get("is_causal", True)))

    def forward(self, positions, hidden):
        q, k, v = self.qkv_proj(hidden).split([self.q_size, self.kv_size, self.kv_size], dim=-1)
        return self.o_proj(self.attn(q, k, v, positions))


class Gemma2DecoderLayer(nn.Module):
    def __init__(self, cfg, mc, dtype, quant_config, prefix: str, layer_idx: int) -> None:
        super().__init__()
        H, eps = cfg["hidden_size"], cfg.get("rms_norm_eps", 1e-6)
        self.self_attn = Gemma2Attention(cfg, mc, dtype, quant_config, f"{prefix}.self_attn", layer_idx)
        self.mlp = LlamaMLP(H, cfg["intermediate_size"], cfg.get("hidden_activation", cfg.get("hidden_act", "gelu_pytorch_tanh")),
                            False, dtype, quant_config, f"{prefix}.mlp")
        self.input_layernorm = RMSNorm(H, eps, dtype, gemma=True)
        self.post_attention_layernorm = RMSNorm(H, eps, dtype, gemma=True)
        self.pre_feedforward_layernorm = RMSNorm(H, eps, dtype, gemma=True)
        self.post_feedforward_layernorm = RMSNorm(H, eps, dtype, gemma=True)

    def forward(self, positions, hidden, residual, next_norm: Optional[nn.Module]):
        """Same calling convention as the Llama layer: ``hidden`` = normalised block input, ``residual`` = stream."""
        a = self.post_attention_layernorm(self.self_attn(positions, hidden))
        stream = residual + a
        m = self.post_feedforward_layernorm(self.mlp(self.pre_feedforward_layernorm(stream)))
        stream = stream + m
        return (next_norm(stream) if next_norm is not None else stream), stream


class Gemma2Model(LlamaModel):
    def _make_layer(self, cfg, mc, dtype, quant_config, prefix, idx):
        return Gemma2DecoderLayer(cfg, mc, dtype, quant_config, prefix, idx)


class Gemma2ForCausalLM(LlamaForCausalLM):
    model_cls = Gemma2Model

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        cfg = vllm_config.model_config.hf_config
        cfg.setdefault("tie_word_embeddings", True)
        cfg["model_type"] = "gemma2"
        super().__init__(vllm_config, prefix)
