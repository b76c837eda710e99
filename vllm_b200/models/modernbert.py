"""ModernBERT: encoder-only, rotary positions, alternating global / local (symmetric sliding window) attention, GeGLU
MLP, bias-free LayerNorms — the current default backbone of embedding / reranker checkpoints.

Reference parity: vllm/model_executor/models/modernbert.py.
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


class ModernBertLayer(nn.Module):
    def __init__(self, cfg: dict, mc, dtype, prefix: str, idx: int) -> None:
        super().__init__()
        H, nh = cfg["hidden_size"], cfg["num_attention_heads"]
        hd = H // nh
        eps, nb = cfg.get("norm_eps", 1e-5), bool(cfg.get("norm_bias", False))
        types = cfg.get("layer_types")
        is_global = (types[idx] == "full_attention") if types else (idx % cfg.get("global_attn_every_n_layers", 3) == 0)
        rp = cfg.get("rope_parameters") or {}
        theta = (rp.get("full_attention" if is_global else "sliding_attention") or {}).get(
            "rope_theta", cfg.get("global_rope_theta" if is_global else "local_rope_theta", 160000.0 if is_global else 10000.0))
        self.attn_norm = LayerNorm(H, eps, nb, dtype) if idx > 0 else None
        self.qkv_proj = QKVParallelLinear(H, hd, nh, nh, bool(cfg.get("attention_bias")), dtype, None, f"{prefix}.attn.qkv_proj")
        self.o_proj = RowParallelLinear(H, H, bool(cfg.get("attention_bias")), dtype=dtype, prefix=f"{prefix}.attn.o_proj")
        max_pos = mc.max_model_len if mc.max_model_len > 0 else cfg.get("max_position_embeddings", 8192)
        # |i - j| <= local_attention / 2  ⇔  the (exclusive) one-sided distance local_attention / 2 + 1 on both sides
        window = None if is_global else int(cfg.get("local_attention", 128)) // 2 + 1
        self.attn = Attention(self.qkv_proj.num_heads, hd, hd ** -0.5, self.qkv_proj.num_kv_heads, sliding_window=window,
                              prefix=f"{prefix}.attn.attn", causal=False, rope=get_rope(hd, hd, max_pos, theta, True, None))
        self.q_size = self.qkv_proj.q_size
        self.mlp_norm = LayerNorm(H, eps, nb, dtype)
        I = cfg["intermediate_size"]
        self.wi = nn.Linear(H, 2 * I, bias=bool(cfg.get("mlp_bias")), dtype=dtype)
        self.wo = nn.Linear(I, H, bias=bool(cfg.get("mlp_bias")), dtype=dtype)

    def get_kv_window(self):
        return self.attn.sliding_window

    def forward(self, positions, x):
        h = self.attn_norm(x) if self.attn_norm is not None else x
        q, k, v = self.qkv_proj(h).split([self.q_size] * 3, dim=-1)
        x = x + self.o_proj(self.attn(q, k, v, positions))
        a, g = self.wi(self.mlp_norm(x)).chunk(2, dim=-1)
        return x + self.wo(F.gelu(a) * g)


class ModernBertModel(nn.Module):
    is_pooling_model = True

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__()
        mc = vllm_config.model_config
        cfg, dtype = mc.hf_config, mc.torch_dtype
        self.cfg = cfg
        H = cfg["hidden_size"]
        eps, nb = cfg.get("norm_eps", 1e-5), bool(cfg.get("norm_bias", False))
        self.tok_embeddings = VocabParallelEmbedding(cfg["vocab_size"], H, dtype)
        self.emb_norm = LayerNorm(H, eps, nb, dtype)
        self.layers = nn.ModuleList([ModernBertLayer(cfg, mc, dtype, f"{prefix}layers.{i}", i) for i in range(cfg["num_hidden_layers"])])
        self.final_norm = LayerNorm(H, eps, nb, dtype)
        self.default_pooling = "CLS" if cfg.get("classifier_pooling", "cls") == "cls" else "MEAN"

    def embed_input_ids(self, ids):
        return self.emb_norm(self.tok_embeddings(ids))

    def forward(self, input_ids, positions, intermediate_tensors=None, inputs_embeds=None):
        x = inputs_embeds if inputs_embeds is not None else self.embed_input_ids(input_ids)
        for layer in self.layers:
            x = layer(positions, x)
        return self.final_norm(x)

    def compute_logits(self, hidden):  # pragma: no cover - pooling-only
        raise RuntimeError("encoder-only model: use LLM.embed / LLM.classify")

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        params = dict(self.named_parameters())
        modules = dict(self.named_modules())
        loaded: set[str] = set()
        ren = [("embeddings.tok_embeddings.", "tok_embeddings."), ("embeddings.norm.", "emb_norm."), (".attn.Wqkv.", ".qkv_proj."),
               (".attn.Wo.", ".o_proj."), (".mlp.Wi.", ".wi."), (".mlp.Wo.", ".wo.")]
        for name, w in weights:
            if name.startswith("model."):
                name = name[len("model."):]
            for a, b in ren:
                name = name.replace(a, b)
            if name not in params:
                continue
            mod = modules[name.rsplit(".", 1)[0]]
            loader = getattr(mod, "weight_loader", None)
            loader(params[name], w, None) if loader is not None else params[name].data.copy_(w)
            loaded.add(name)
        return loaded


class ModernBertForSequenceClassification(ModernBertModel):
    """Prediction head (dense → activation → LayerNorm) on the [CLS] / mean-pooled state, then a linear classifier
    (reference: models/modernbert.py)."""
    per_token = False

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__(vllm_config, prefix)
        cfg, dtype = self.cfg, vllm_config.model_config.torch_dtype
        H = cfg["hidden_size"]
        n = int(cfg.get("num_labels") or len(cfg.get("id2label") or {}) or 2)
        self.head_dense = nn.Linear(H, H, bias=bool(cfg.get("classifier_bias", False)), dtype=dtype)
        self.head_norm = LayerNorm(H, cfg.get("norm_eps", 1e-5), bool(cfg.get("norm_bias", False)), dtype)
        self.classifier = nn.Linear(H, n, dtype=dtype)
        from vllm_b200.layers.common import Activation
        self.head_act = Activation(cfg.get("classifier_activation", "gelu"))
        if self.per_token:
            self.default_pooling = "ALL"

    def score(self, pooled: torch.Tensor) -> torch.Tensor:
        x = self.head_norm(self.head_act(self.head_dense(pooled)))
        return self.classifier(x)

    def load_weights(self, weights):
        return super().load_weights((n.replace("head.dense.", "head_dense.").replace("head.norm.", "head_norm."), w) for n, w in weights)


class ModernBertForTokenClassification(ModernBertForSequenceClassification):
    per_token = True


class ColBERTModernBertModel(ModernBertModel):
    """ModernBERT late-interaction retriever (answerai-colbert, GTE-ModernColBERT): every token state → bias-free linear →
    L2 norm (reference: models/colbert.py)."""

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__(vllm_config, prefix)
        cfg = self.cfg
        self.default_pooling = "ALL"
        self.colbert_linear = nn.Linear(cfg["hidden_size"], int(cfg.get("colbert_dim") or cfg.get("dim") or 128), bias=False,
                                        dtype=vllm_config.model_config.torch_dtype)

    def project_embedding(self, x: torch.Tensor, per_token: bool) -> torch.Tensor:
        return self.colbert_linear(x) if per_token else x

    def load_weights(self, weights):
        return super().load_weights((("colbert_linear.weight" if n in ("linear.weight", "colbert_linear.weight", "1.linear.weight") else n), w)
                                    for n, w in weights)
