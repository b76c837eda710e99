"""Nomic-BERT (nomic-embed-text): a post-norm BERT encoder with rotary positions instead of learned ones, bias-free
projections and a SwiGLU MLP; mean-pooled for embeddings.

Reference parity: vllm/model_executor/models/bert_with_rope.py (``NomicBertModel``), HF ``modeling_nomic_bert``.
Bidirectional attention through the shared paged-attention layer (``causal=False``), like ModernBERT.
"""
from __future__ import annotations

from typing import Iterable

import torch
import torch.nn as nn
import torch.nn.functional as F

from vllm_b200.config import VllmConfig
from vllm_b200.layers.attention import Attention
from vllm_b200.layers.common import LayerNorm, VocabParallelEmbedding
from vllm_b200.layers.linear import QKVParallelLinear, RowParallelLinear
from vllm_b200.layers.rotary import get_rope
from vllm_b200.models.llama import LlamaMLP


class NomicBertLayer(nn.Module):
    def __init__(self, cfg: dict, mc, dtype, prefix: str) -> None:
        super().__init__()
        H, nh = cfg["hidden_size"], cfg["num_attention_heads"]
        hd = cfg.get("head_dim") or H // nh
        eps = cfg.get("layer_norm_eps", 1e-12)
        bias = bool(cfg.get("qkv_proj_bias", False))
        self.qkv_proj = QKVParallelLinear(H, hd, nh, nh, bias, dtype, None, f"{prefix}.self_attn.qkv_proj")
        self.o_proj = RowParallelLinear(nh * hd, H, bias, dtype=dtype, prefix=f"{prefix}.self_attn.o_proj")
        rp = cfg.get("rope_parameters") or {}
        max_pos = mc.max_model_len if mc.max_model_len > 0 else cfg.get("max_position_embeddings", 2048)
        scaling = {k: v for k, v in rp.items() if k != "rope_theta"} if rp.get("rope_type", "default") != "default" else None
        rope = get_rope(hd, int(hd * float(cfg.get("rotary_emb_fraction", 1.0))), max_pos,
                        rp.get("rope_theta", cfg.get("rotary_emb_base", 1000.0)), True, scaling)
        self.attn = Attention(self.qkv_proj.num_heads, hd, hd ** -0.5, self.qkv_proj.num_kv_heads, prefix=f"{prefix}.self_attn.attn",
                              causal=False, rope=rope)
        self.q_size = self.qkv_proj.q_size
        self.mlp = LlamaMLP(H, cfg["intermediate_size"], cfg.get("hidden_act", "silu"), bool(cfg.get("mlp_fc1_bias", False)), dtype,
                            None, f"{prefix}.mlp")
        self.post_attention_layernorm = LayerNorm(H, eps, True, dtype)
        self.post_mlp_layernorm = LayerNorm(H, eps, True, dtype)

    def forward(self, positions, x):
        q, k, v = self.qkv_proj(x).split([self.q_size] * 3, dim=-1)
        x = self.post_attention_layernorm(x + self.o_proj(self.attn(q, k, v, positions)))
        return self.post_mlp_layernorm(x + self.mlp(x))


class NomicBertModel(nn.Module):
    is_pooling_model = True
    default_pooling = "MEAN"
    stacked = [("qkv_proj", "q_proj", "q"), ("qkv_proj", "k_proj", "k"), ("qkv_proj", "v_proj", "v"),
               ("gate_up_proj", "gate_proj", 0), ("gate_up_proj", "up_proj", 1)]

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__()
        mc = vllm_config.model_config
        cfg, dtype = mc.hf_config, mc.torch_dtype
        self.cfg = cfg
        H = cfg["hidden_size"]
        self.word_embeddings = VocabParallelEmbedding(cfg["vocab_size"], H, dtype)
        self.token_type_embeddings = nn.Embedding(cfg.get("type_vocab_size", 2), H, dtype=dtype) if cfg.get("type_vocab_size", 2) else None
        self.emb_ln = LayerNorm(H, cfg.get("layer_norm_eps", 1e-12), True, dtype)
        self.layers = nn.ModuleList([NomicBertLayer(cfg, mc, dtype, f"{prefix}layers.{i}") for i in range(cfg["num_hidden_layers"])])

    def embed_input_ids(self, ids):
        x = self.word_embeddings(ids)
        if self.token_type_embeddings is not None:
            x = x + self.token_type_embeddings.weight[0]
        return self.emb_ln(x)

    def forward(self, input_ids, positions, intermediate_tensors=None, inputs_embeds=None):
        x = inputs_embeds if inputs_embeds is not None else self.embed_input_ids(input_ids)
        for layer in self.layers:
            x = layer(positions, x)
        return x

    def compute_logits(self, hidden):  # pragma: no cover - pooling-only
        raise RuntimeError("encoder-only model: use LLM.embed")

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        params = dict(self.named_parameters())
        modules = dict(self.named_modules())
        loaded: set[str] = set()
        for name, w in weights:
            for pre in ("model.", "bert.", "nomic_bert."):
                if name.startswith(pre):
                    name = name[len(pre):]
            n = (name.replace("embeddings.word_embeddings.", "word_embeddings.").replace("embeddings.token_type_embeddings.", "token_type_embeddings.")
                 .replace("embeddings.LayerNorm.", "emb_ln.").replace(".self_attn.o_proj.", ".o_proj."))
            target, shard = n, None
            for fused, part, sid in self.stacked:
                if f".{part}." in n:
                    target, shard = n.replace(part, fused).replace(".self_attn.qkv_proj.", ".qkv_proj."), sid
                    break
            if target not in params:
                continue
            mod = modules[target.rsplit(".", 1)[0]]
            loader = getattr(mod, "weight_loader", None)
            loader(params[target], w, shard) if loader is not None else params[target].data.copy_(w)
            loaded.add(target)
        return loaded
