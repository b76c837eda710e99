"""Families that ARE the Llama block under other tensor names, config keys or fused checkpoint layouts.

Each class only normalises the config and re-keys / re-slices the checkpoint stream; the forward pass is the stock Llama
(or Qwen2 / Granite-style variant) path with its fused kernels.

  InternLM2   — ``wqkv`` packs each kv-group as [q·g | k | v] along the output dim; ``w1/w3/w2`` MLP; ``tok_embeddings`` /
                ``output``                                             (reference: models/internlm2.py)
  InternLM3   — Llama with ``qkv_bias`` / ``bias`` config keys          (reference registry: llama.LlamaForCausalLM)
  EXAONE-3    — GPT-style names (``transformer.h.N.attn.attention.q_proj``, ``mlp.c_fc_0/1``, ``c_proj``) (models/exaone.py)
  MiMo        — Qwen2 plus multi-token-prediction heads, which the base model ignores (models/mimo.py)
  MiniCPM     — Llama with muP scalars: scaled embeddings, residual scale depth/√L, logits ÷ (H / dim_model_base)
                (models/minicpm.py; dense checkpoints)
  TeleFLM     — Llama with muP input / output multipliers               (models/teleflm.py)
  Orion       — Llama with biased LayerNorms                            (models/orion.py)
  IQuest-Coder / TeleChat3 — plain aliases
"""
from __future__ import annotations

import math
from typing import Iterable

import torch

from vllm_b200.config import VllmConfig
from vllm_b200.layers.common import LogitsProcessor
from vllm_b200.models.llama import LlamaForCausalLM, Qwen2ForCausalLM
from vllm_b200.models.llama_variants import GraniteForCausalLM, VariantForCausalLM


class InternLM3ForCausalLM(LlamaForCausalLM):
    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        cfg = vllm_config.model_config.hf_config
        if "attention_bias" not in cfg and (cfg.get("qkv_bias") or cfg.get("bias")):
            if bool(cfg.get("qkv_bias")) != bool(cfg.get("bias")):
                raise NotImplementedError("InternLM3 with a bias on only one side of the attention block")
            cfg["attention_bias"] = True
        super().__init__(vllm_config, prefix)


class InternLM2ForCausalLM(LlamaForCausalLM):
    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        cfg = vllm_config.model_config.hf_config
        cfg.setdefault("attention_bias", bool(cfg.get("bias", False)))
        super().__init__(vllm_config, prefix)

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        c = self.cfg
        nh, nkv = c["num_attention_heads"], c.get("num_key_value_heads") or c["num_attention_heads"]
        hd = c.get("head_dim") or c["hidden_size"] // nh
        g = nh // nkv

        def mapped():
            for name, w in weights:
                name = (name.replace("model.tok_embeddings.", "model.embed_tokens.").replace("output.weight", "lm_head.weight")
                        .replace(".attention_norm.", ".input_layernorm.").replace(".ffn_norm.", ".post_attention_layernorm.")
                        .replace(".attention.wo.", ".self_attn.o_proj.").replace(".feed_forward.w1.", ".mlp.gate_proj.")
                        .replace(".feed_forward.w3.", ".mlp.up_proj.").replace(".feed_forward.w2.", ".mlp.down_proj."))
                if ".attention.wqkv." in name:
                    t = w.view(nkv, g + 2, hd, *w.shape[1:])
                    rest = w.shape[1:]
                    yield name.replace(".attention.wqkv.", ".self_attn.q_proj."), t[:, :g].reshape(nh * hd, *rest)
                    yield name.replace(".attention.wqkv.", ".self_attn.k_proj."), t[:, g].reshape(nkv * hd, *rest)
                    yield name.replace(".attention.wqkv.", ".self_attn.v_proj."), t[:, g + 1].reshape(nkv * hd, *rest)
                    continue
                if name.startswith("v_head."):                 # reward head: the pooling adapter owns it
                    continue
                yield name, w
        return super().load_weights(mapped())


class ExaoneForCausalLM(LlamaForCausalLM):
    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        cfg = vllm_config.model_config.hf_config
        cfg.setdefault("num_hidden_layers", cfg.get("num_layers"))
        cfg.setdefault("rms_norm_eps", cfg.get("layer_norm_epsilon", 1e-5))
        cfg.setdefault("hidden_act", cfg.get("activation_function", "silu"))
        super().__init__(vllm_config, prefix)

    def load_weights(self, weights):
        ren = [("transformer.wte.", "model.embed_tokens."), ("transformer.ln_f.", "model.norm."), ("transformer.h.", "model.layers."),
               (".attn.attention.out_proj.", ".self_attn.o_proj."), (".attn.attention.", ".self_attn."), (".ln_1.", ".input_layernorm."),
               (".ln_2.", ".post_attention_layernorm."), (".mlp.c_fc_0.", ".mlp.gate_proj."), (".mlp.c_fc_1.", ".mlp.up_proj."),
               (".mlp.c_proj.", ".mlp.down_proj.")]

        def mapped():
            for name, w in weights:
                for a, b in ren:
                    name = name.replace(a, b)
                yield name, w
        return super().load_weights(mapped())


class MiMoForCausalLM(Qwen2ForCausalLM):
    def load_weights(self, weights):
        return super().load_weights((n, w) for n, w in weights if "mtp_layers" not in n)


class MiniCPMForCausalLM(GraniteForCausalLM):
    def options(self, c):
        if c.get("num_experts"):
            raise NotImplementedError("MiniCPM-MoE checkpoints")
        c.setdefault("embedding_multiplier", float(c.get("scale_emb", 1.0)))
        c.setdefault("residual_multiplier", float(c.get("scale_depth", 1.0)) / math.sqrt(c["num_hidden_layers"]))
        c.setdefault("logits_scaling", c["hidden_size"] / float(c.get("dim_model_base", c["hidden_size"])))
        c.setdefault("attention_multiplier", (c.get("head_dim") or c["hidden_size"] // c["num_attention_heads"]) ** -0.5)
        return super().options(c)


class TeleFLMForCausalLM(LlamaForCausalLM):
    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__(vllm_config, prefix)
        c = self.cfg
        self._in_mult = float(c.get("input_mult", 1.0)) if c.get("use_mup") else 1.0
        if c.get("use_mup") and self.is_last:
            self.logits_processor = LogitsProcessor(c["vocab_size"], scale=float(c.get("output_mult", 1.0)) / float(c.get("mup_scale_factor", 1.0)))
        self.model.embed_scale = self._in_mult


class OrionForCausalLM(VariantForCausalLM):
    def options(self, c):
        return dict(norm="ln", eps=c.get("rms_norm_eps", 1e-5), qkv_bias=bool(c.get("attention_bias")), o_bias=bool(c.get("attention_bias")))
    renames = [(".post_attention_layernorm.", ".pre_mlp_norm.")]


# ---- decoders run as bidirectional encoders (embedding / reranking checkpoints trained without the causal mask) ----------
_POOL = {"avg": "MEAN", "mean": "MEAN", "cls": "CLS", "last": "LAST"}


def _bidirectional(base):
    from vllm_b200.models import adapters

    class Bidirectional(adapters.as_embedding_model(base)):           # type: ignore[misc, valid-type]
        """``is_causal = False`` on every attention layer; the whole sequence is one prefill pass (the engine disables
        chunked prefill and prefix reuse for these, like for the BERT family); pooling from ``config.pooling``
        (reference: models/config.py ``LlamaBidirectionalConfig``)."""

        def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
            cfg = vllm_config.model_config.hf_config
            cfg["is_causal"] = False
            super().__init__(vllm_config, prefix)
            self.default_pooling = _POOL.get(str(cfg.get("pooling", "avg")).lower(), "MEAN")
    return Bidirectional


LlamaBidirectionalModel = _bidirectional(LlamaForCausalLM)
LlamaBidirectionalModel.__name__ = "LlamaBidirectionalModel"


class LlamaBidirectionalForSequenceClassification(LlamaForCausalLM):
    pass


def _make_bidir_cls():
    from vllm_b200.models import adapters
    base = adapters.as_seq_cls_model(LlamaForCausalLM)

    class Cls(base):                                                   # type: ignore[misc, valid-type]
        def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
            cfg = vllm_config.model_config.hf_config
            cfg["is_causal"] = False
            super().__init__(vllm_config, prefix)
            self.default_pooling = _POOL.get(str(cfg.get("pooling", "avg")).lower(), "MEAN")
    Cls.__name__ = "LlamaBidirectionalForSequenceClassification"
    return Cls


LlamaBidirectionalForSequenceClassification = _make_bidir_cls()


class VoyageQwen3BidirectionalEmbedModel(_bidirectional(__import__("vllm_b200.models.llama", fromlist=["Qwen3ForCausalLM"]).Qwen3ForCausalLM)):
    """voyage-4 family: bidirectional Qwen3 with a bias-free linear embedding head (``linear.weight``: hidden → num_labels)
    applied to the pooled state (reference: models/voyage.py)."""

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__(vllm_config, prefix)
        cfg = self.cfg
        n = int(cfg.get("num_labels") or cfg["hidden_size"])
        self.voyage_linear = torch.nn.Linear(cfg["hidden_size"], n, bias=False, dtype=vllm_config.model_config.torch_dtype)
        self.default_pooling = _POOL.get(str(cfg.get("pooling", "avg")).lower(), "MEAN")

    def project_embedding(self, x, per_token: bool):
        return self.voyage_linear(x)

    def load_weights(self, weights):
        rest = []
        for n, w in weights:
            if n == "linear.weight":
                self.voyage_linear.weight.data.copy_(w)
            else:
                rest.append((n, w))
        return super().load_weights(rest) | {"voyage_linear.weight"}


# ---- Nemotron-NAS / DeciLM: a Llama whose layers were searched individually ----------------------------------------------
def _round_up(n: int, m: int) -> int:
    return n if n % m == 0 else n + m - n % m


class _NasLayer(torch.nn.Module):
    """One searched block: attention with its own KV-head count (or none at all), gated MLP with its own width (or none).
    The residual stream is explicit — the fused add-norm path assumes both halves exist."""

    def __init__(self, cfg, mc, dtype, quant_config, prefix, idx) -> None:
        super().__init__()
        from vllm_b200.layers.common import RMSNorm
        from vllm_b200.models.llama import LlamaAttention, LlamaMLP
        from vllm_b200.models.llama_variants import _Identity
        b = cfg["block_configs"][idx]
        a, f = b.get("attention") or {}, b.get("ffn") or {}
        if a.get("replace_with_linear") or f.get("replace_with_linear"):
            raise NotImplementedError("Nemotron-NAS blocks replaced by a single linear layer")
        H, eps = cfg["hidden_size"], cfg.get("rms_norm_eps", 1e-5)
        self.no_attn, self.no_ffn = bool(a.get("no_op")), bool(f.get("no_op"))
        if not self.no_ffn:
            I = f.get("intermediate_size") or _round_up(int(2 * float(f["ffn_mult"]) * H / 3), 256)
            self.post_attention_layernorm = RMSNorm(H, eps, dtype)
            self.mlp = LlamaMLP(H, int(I), cfg.get("hidden_act", "silu"), bool(cfg.get("mlp_bias", False)), dtype, quant_config, f"{prefix}.mlp")
        if not self.no_attn:
            sub = dict(cfg, num_key_value_heads=cfg["num_attention_heads"] // int(a["n_heads_in_group"]))
            self.self_attn = LlamaAttention(sub, mc, dtype, quant_config, f"{prefix}.self_attn", idx)
            self.input_layernorm = RMSNorm(H, eps, dtype)
        else:      # the first norm of the block is whatever comes first in it (LlamaModel applies it before calling the block)
            self.input_layernorm = self.post_attention_layernorm if not self.no_ffn else _Identity()

    def forward(self, positions, hidden, residual, next_norm):
        x, h = residual, hidden
        if not self.no_attn:
            x = x + self.self_attn.o_proj(self.self_attn(positions, h))
            h = self.post_attention_layernorm(x) if not self.no_ffn else None
        if not self.no_ffn:
            x = x + self.mlp(h)
        return (next_norm(x) if next_norm is not None else x), x


class _NasModel(__import__("vllm_b200.models.llama", fromlist=["LlamaModel"]).LlamaModel):
    def _make_layer(self, cfg, mc, dtype, quant_config, prefix, idx):
        return _NasLayer(cfg, mc, dtype, quant_config, prefix, idx)


class DeciLMForCausalLM(LlamaForCausalLM):
    """Llama-3.1-Nemotron-51B / Llama-3.3-Nemotron-Super-49B / Ultra-253B: ``block_configs`` gives every layer its own
    KV-head count and MLP width, and lets it drop the attention or the MLP altogether (reference: models/nemotron_nas.py)."""
    model_cls = _NasModel
