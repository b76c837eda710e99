"""Decoder families that are Llama with a twist, built from one configurable block:

  Granite   — embedding / attention / residual / logit multipliers          (reference: models/granite.py)
  OLMo-2    — no pre-norms: RMSNorm on each sub-layer OUTPUT, q/k RMSNorm over the whole projection (models/olmo2.py)
  Cohere    — LayerNorm, attention and MLP in parallel from one norm, interleaved rotary, logit scale (models/commandr.py)
  Phi-1/2   — LayerNorm, parallel block, plain GELU MLP with biases, partial rotary, biased LM head (models/phi.py)
  GLM-4     — sandwich norms, fused gate_up checkpoint tensor, half-rotary interleaved           (models/glm4.py)
  Nemotron  — LayerNorm with +1 weights, squared-ReLU plain MLP, partial rotary                  (models/nemotron.py)
  Gemma-3   — sandwich norms, per-head (1+w) q/k norms, 5 local : 1 global layers with their own rope base
              (models/gemma3.py; text tower)

The block keeps the residual stream explicit (no fused add+norm): these families differ exactly in where the norms
and multipliers sit, which is what the fused Llama path bakes in.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.nn as nn

from vllm_b200.config import VllmConfig
from vllm_b200.layers.attention import Attention
from vllm_b200.layers.common import Activation, LayerNorm, LogitsProcessor, ParallelLMHead, RMSNorm
from vllm_b200.layers.linear import ColumnParallelLinear, QKVParallelLinear, RowParallelLinear
from vllm_b200.layers.rotary import get_rope
from vllm_b200.models.llama import LlamaForCausalLM, LlamaMLP, LlamaModel
from vllm_b200.parallel.state import get_tp_group


class _LN1P(nn.Module):
    """LayerNorm whose stored weight is (gamma - 1) (Nemotron)."""

    def __init__(self, H, eps, dtype) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(H, dtype=dtype), requires_grad=False)
        self.bias = nn.Parameter(torch.zeros(H, dtype=dtype), requires_grad=False)
        self.eps = eps

    def forward(self, x):
        return torch.nn.functional.layer_norm(x.float(), (x.shape[-1],), self.weight.float() + 1.0, self.bias.float(),
                                              self.eps).to(x.dtype)


class _LNPlain(nn.Module):
    """LayerNorm without learnable scale or shift (OLMo-1)."""

    def __init__(self, eps) -> None:
        super().__init__()
        self.eps = eps

    def forward(self, x):
        return torch.nn.functional.layer_norm(x.float(), (x.shape[-1],), None, None, self.eps).to(x.dtype)


class _Identity(nn.Module):
    def forward(self, x, residual=None):
        return x if residual is None else (x + residual, x + residual)


def _norm(kind: str, H: int, eps: float, dtype) -> nn.Module:
    if kind == "rms":
        return RMSNorm(H, eps, dtype)
    if kind == "rms1p":
        return RMSNorm(H, eps, dtype, True)
    if kind == "ln":
        return LayerNorm(H, eps, True, dtype)
    if kind == "ln_nobias":
        return LayerNorm(H, eps, False, dtype)
    if kind == "ln1p":
        return _LN1P(H, eps, dtype)
    if kind == "ln_plain":
        return _LNPlain(eps)
    if kind == "none":
        return _Identity()
    raise ValueError(kind)


class _XIELU(nn.Module):
    """xIELU (arXiv:2411.13010; Apertus): x > 0 → a_p x² + β x, else a_n (expm1(min(x, ε)) − x) + β x, with
    a_p = softplus(alpha_p), a_n = β + softplus(alpha_n) learned per layer."""

    def __init__(self, dtype) -> None:
        super().__init__()
        self.alpha_p = nn.Parameter(torch.zeros(1, dtype=dtype), requires_grad=False)
        self.alpha_n = nn.Parameter(torch.zeros(1, dtype=dtype), requires_grad=False)
        self.beta = nn.Parameter(torch.full((), 0.5, dtype=dtype), requires_grad=False)
        self.eps = nn.Parameter(torch.full((), -1e-6, dtype=dtype), requires_grad=False)

    def forward(self, x):
        ap = torch.nn.functional.softplus(self.alpha_p)
        an = self.beta + torch.nn.functional.softplus(self.alpha_n)
        return torch.where(x > 0, ap * x * x + self.beta * x, (torch.expm1(torch.min(x, self.eps)) - x) * an + self.beta * x)


class _PlainMLP(nn.Module):
    def __init__(self, H, I, act, bias, dtype, prefix) -> None:
        super().__init__()
        # (names differ from the gated MLP's on purpose: the Llama loader fuses ``up_proj`` into ``gate_up_proj``)
        self.fc_in = ColumnParallelLinear(H, I, bias, dtype=dtype, prefix=f"{prefix}.fc_in")
        self.fc_out = RowParallelLinear(I, H, bias, dtype=dtype, prefix=f"{prefix}.fc_out")
        if act == "xielu":
            self.act_fn = _XIELU(dtype)
        self.act = Activation(act) if act not in ("relu2", "xielu") else None
        self.kind = act

    def forward(self, x):
        h = self.fc_in(x)
        if self.kind == "xielu":
            h = self.act_fn(h)
        else:
            h = torch.relu(h).square() if self.act is None else self.act(h)
        return self.fc_out(h)


class VariantAttention(nn.Module):
    def __init__(self, cfg, mc, dtype, prefix, layer_idx, o) -> None:
        super().__init__()
        H, nh = cfg["hidden_size"], cfg["num_attention_heads"]
        nkv = cfg.get("num_key_value_heads") or nh
        hd = cfg.get("head_dim") or H // nh
        self.hd = hd
        self.qkv_proj = QKVParallelLinear(H, hd, nh, nkv, o["qkv_bias"], dtype, None, f"{prefix}.qkv_proj")
        self.o_proj = RowParallelLinear(nh * hd, H, o["o_bias"], dtype=dtype, prefix=f"{prefix}.o_proj")
        self.q_size, self.kv_size = self.qkv_proj.q_size, self.qkv_proj.kv_size
        eps = o["eps"]
        self.clip_qkv = o.get("clip_qkv")                             # clamp the projections to ±clip (OLMo-1)
        self.q_pos_scale = o.get("q_pos_scale")                       # (beta, original_max_pos): Llama-4 style length scaling
        self.out_gate = None
        if o.get("attn_gate"):                                        # attention output · sigmoid(W_g x), element-wise
            self.out_gate = ColumnParallelLinear(H, nh * hd, False, dtype=dtype, prefix=f"{prefix}.out_gate")
        self.qk_norm = o["qk_norm"]                                   # None | "head" | "head1p" | "full"
        if self.qk_norm in ("head", "head1p", "head_post_rope"):
            self.q_norm, self.k_norm = RMSNorm(hd, eps, dtype, self.qk_norm == "head1p"), RMSNorm(hd, eps, dtype, self.qk_norm == "head1p")
        elif self.qk_norm == "full":
            assert get_tp_group().world_size == 1, "whole-projection q/k norms need the unsharded projection"
            self.q_norm, self.k_norm = RMSNorm(nh * hd, eps, dtype), RMSNorm(nkv * hd, eps, dtype)
        rp = cfg.get("rope_parameters") or {}
        sliding = o["sliding"](layer_idx)
        if "full_attention" in rp or "sliding_attention" in rp:             # per-layer-type rope (Gemma-3)
            rp = rp["sliding_attention" if sliding else "full_attention"]
        theta = rp.get("rope_theta", cfg.get("rope_theta", 10000.0))
        scaling = cfg.get("rope_scaling") or ({k: v for k, v in rp.items() if k not in ("rope_theta", "partial_rotary_factor")}
                                               if rp.get("rope_type", "default") != "default" else None)
        factor = rp.get("partial_rotary_factor", cfg.get("partial_rotary_factor", 1.0))
        max_pos = mc.max_model_len if mc.max_model_len > 0 else cfg.get("max_position_embeddings", 8192)
        rope = get_rope(hd, int(hd * factor), max_pos, theta, o["neox"], scaling) if o["use_rope"](layer_idx, sliding) else None
        self.attn = Attention(self.qkv_proj.num_heads, hd, o["scale"] if o["scale"] is not None else hd ** -0.5,
                              self.qkv_proj.num_kv_heads, sliding_window=cfg.get("sliding_window") if sliding else None,
                              logits_soft_cap=cfg.get("attn_logit_softcapping") or 0.0, prefix=f"{prefix}.attn", rope=rope)

    def forward(self, positions, x):
        a = self._attend(positions, x)
        if self.out_gate is not None:
            a = a * torch.sigmoid(self.out_gate(x))
        return self.o_proj(a)

    def _attend(self, positions, x):
        qkv = self.qkv_proj(x)
        if self.clip_qkv:
            qkv = qkv.clamp(-self.clip_qkv, self.clip_qkv)
        q, k, v = qkv.split([self.q_size, self.kv_size, self.kv_size], dim=-1)
        if self.q_pos_scale:
            # q · (1 + β log(1 + ⌊pos / L₀⌋)): a per-token scalar, so it commutes with the rotation applied later
            beta, L0 = self.q_pos_scale
            s = 1.0 + beta * torch.log1p(torch.floor(positions.to(torch.float32) / L0))
            q = (q * s.unsqueeze(-1).to(q.dtype))
        if self.qk_norm in ("head", "head1p"):
            # per-head norms ride along with RoPE + the KV insert (one fused launch on CUDA, layers/attention.py)
            return self.attn(q, k, v, positions, qk_norm=(self.q_norm, self.k_norm))
        elif self.qk_norm == "full":
            q, k = self.q_norm(q.contiguous()), self.k_norm(k.contiguous())
        elif self.qk_norm == "head_post_rope":
            # HunYuan: the per-head norms act on the ROTATED q / k — rotate here, norm, hand over without positions
            T = q.shape[0]
            if self.attn.rope is not None:
                q, k = self.attn.rope(positions, q.contiguous(), k.contiguous())
            q = self.q_norm(q.reshape(T, -1, self.hd)).reshape(T, -1)
            k = self.k_norm(k.reshape(T, -1, self.hd)).reshape(T, -1)
            return self.attn(q, k, v, None)
        return self.attn(q, k, v, positions)


class VariantLayer(nn.Module):
    def __init__(self, cfg, mc, dtype, quant_config, prefix, idx, o) -> None:
        super().__init__()
        H, eps = cfg["hidden_size"], o["eps"]
        self.o = o
        self.self_attn = VariantAttention(cfg, mc, dtype, f"{prefix}.self_attn", idx, o)
        self.moe_partial = False
        m = o.get("moe")
        if callable(m):                      # per-layer choice: leading dense layers, alternating sparse steps
            m = m(idx)
        if m:
            # sparse block (router + fused experts); returns this rank's PARTIAL sum, reduced in forward()
            from vllm_b200.models.mixtral import SparseMoeBlock
            self.mlp = SparseMoeBlock(cfg, dtype, quant_config, f"{prefix}.mlp", m.get("enable_ep", False),
                                      n_experts=m["n_experts"], top_k=m["top_k"], inter=m.get("inter") or cfg["intermediate_size"],
                                      renormalize=m.get("renormalize", False), shared_inter=m.get("shared_inter", 0),
                                      moe_kw=m.get("kw"))
            if m.get("routing") is not None:
                self.mlp.experts.custom_routing = m["routing"]
            self.moe_partial = True
        elif o["gated_mlp"]:
            self.mlp = LlamaMLP(H, cfg["intermediate_size"], o["act"], o["mlp_bias"], dtype, quant_config, f"{prefix}.mlp")
        else:
            self.mlp = _PlainMLP(H, cfg["intermediate_size"], o["act"], o["mlp_bias"], dtype, f"{prefix}.mlp")
        st = o["structure"]
        self.input_layernorm = _norm(o["norm"] if st != "post" else "none", H, eps, dtype)
        if st in ("pre", "sandwich"):
            self.pre_mlp_norm = _norm(o["norm"], H, eps, dtype)
        if st in ("post", "sandwich"):
            self.post_attn_norm, self.post_mlp_norm = _norm(o["norm"], H, eps, dtype), _norm(o["norm"], H, eps, dtype)
        self.rm = float(o["residual_multiplier"])

    def forward(self, positions, hidden, residual, next_norm: Optional[nn.Module]):
        st, x = self.o["structure"], residual
        if st == "parallel":
            x = x + (self.self_attn(positions, hidden) + self.mlp(hidden)) * self.rm
        else:
            a = self.self_attn(positions, hidden)
            if st in ("post", "sandwich"):
                a = self.post_attn_norm(a)
            x = x + a * self.rm
            m = self.mlp(self.pre_mlp_norm(x) if st in ("pre", "sandwich") else x)
            if self.moe_partial and get_tp_group().world_size > 1:
                m = get_tp_group().all_reduce(m)
            if st in ("post", "sandwich"):
                m = self.post_mlp_norm(m)
            x = x + m * self.rm
        return (next_norm(x) if next_norm is not None else x), x


_DEFAULTS = dict(norm="rms", structure="pre", eps=1e-6, qkv_bias=False, o_bias=False, mlp_bias=False, gated_mlp=True,
                 act="silu", qk_norm=None, neox=True, scale=None, residual_multiplier=1.0, sliding=lambda i: False,
                 embed_scale=1.0, logit_scale=1.0, lm_head_bias=False, final_norm_name="model.norm",
                 use_rope=lambda i, sliding: True, clip_qkv=None)


class VariantModel(LlamaModel):
    def __init__(self, vllm_config: VllmConfig, prefix: str = "model", quant_config=None) -> None:
        self._opts = vllm_config.model_config.hf_config["_variant_opts"]
        super().__init__(vllm_config, prefix, quant_config)
        o = self._opts
        self.embed_scale = o["embed_scale"]
        if self.norm is not None:
            self.norm = _norm(o["norm"], self.cfg["hidden_size"], o["eps"], vllm_config.model_config.torch_dtype)

    def _make_layer(self, cfg, mc, dtype, quant_config, prefix, idx):
        return VariantLayer(cfg, mc, dtype, quant_config, prefix, idx, self._opts)


class VariantForCausalLM(LlamaForCausalLM):
    model_cls = VariantModel
    renames: list[tuple[str, str]] = []

    def options(self, cfg: dict) -> dict:
        return {}

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        cfg = vllm_config.model_config.hf_config
        o = dict(_DEFAULTS)
        o.update(self.options(cfg))
        cfg["_variant_opts"] = o
        super().__init__(vllm_config, prefix)
        if self.is_last:
            if o["lm_head_bias"]:
                self.lm_head = ParallelLMHead(cfg["vocab_size"], cfg["hidden_size"], bias=True,
                                              dtype=vllm_config.model_config.torch_dtype, prefix="lm_head")
            self.logits_processor = LogitsProcessor(cfg["vocab_size"], scale=o["logit_scale"],
                                                    soft_cap=cfg.get("final_logit_softcapping"))

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        def mapped():
            for name, w in weights:
                for a, b in self.renames:
                    if a in name:
                        name = name.replace(a, b)
                if name.endswith(".mlp.gate_up_proj.weight") and self.cfg["_variant_opts"]["gated_mlp"]:
                    I = w.shape[0] // 2                     # checkpoint stores the fused tensor: feed both halves
                    yield name.replace("gate_up_proj", "gate_proj"), w[:I]
                    yield name.replace("gate_up_proj", "up_proj"), w[I:]
                    continue
                yield name, w
        return super().load_weights(mapped())


class GraniteForCausalLM(VariantForCausalLM):
    def options(self, c):
        types = c.get("layer_types")
        return dict(sliding=(lambda i: types[i] == "sliding_attention") if types and c.get("sliding_window") else (lambda i: False),
                    eps=c.get("rms_norm_eps", 1e-6), scale=float(c.get("attention_multiplier", 1.0)), qkv_bias=bool(c.get("attention_bias")),
                    o_bias=bool(c.get("attention_bias")), mlp_bias=bool(c.get("mlp_bias")), residual_multiplier=c.get("residual_multiplier", 1.0),
                    embed_scale=float(c.get("embedding_multiplier", 1.0)), logit_scale=1.0 / float(c.get("logits_scaling", 1.0)))
    renames = [(".post_attention_layernorm.", ".pre_mlp_norm.")]


class HyperCLOVAXForCausalLM(GraniteForCausalLM):
    """HyperCLOVA X (SEED / THINK text decoders): the Granite muP parameterisation (embedding, attention, residual multipliers;
    ``logits_scaling`` here MULTIPLIES the logits) with optional extra norms on the sub-layer outputs (``use_post_norm``)
    (reference: models/hyperclovax.py)."""
    def options(self, c):
        o = super().options(c)
        o["logit_scale"] = float(c.get("logit_scale", 1.0)) * float(c.get("logits_scaling", 1.0))
        if c.get("use_post_norm"):
            o["structure"] = "sandwich"
        return o
    renames = [(".post_attention_layernorm.", ".pre_mlp_norm."), (".post_norm1.", ".post_attn_norm."), (".post_norm2.", ".post_mlp_norm.")]


class Olmo2ForCausalLM(VariantForCausalLM):
    def options(self, c):
        return dict(structure="post", eps=c.get("rms_norm_eps", 1e-5), qk_norm="full", qkv_bias=bool(c.get("attention_bias")),
                    o_bias=bool(c.get("attention_bias")))
    renames = [(".post_attention_layernorm.", ".post_attn_norm."), (".post_feedforward_layernorm.", ".post_mlp_norm.")]


class CohereForCausalLM(VariantForCausalLM):
    def options(self, c):
        c.setdefault("tie_word_embeddings", True)
        return dict(norm="ln_nobias", structure="parallel", eps=c.get("layer_norm_eps", 1e-5), neox=False,
                    qkv_bias=bool(c.get("attention_bias")), o_bias=bool(c.get("attention_bias")), logit_scale=float(c.get("logit_scale", 1.0)),
                    qk_norm=None)


class PhiForCausalLM(VariantForCausalLM):
    def options(self, c):
        return dict(norm="ln", structure="parallel", eps=c.get("layer_norm_eps", 1e-5), qkv_bias=True, o_bias=True, mlp_bias=True,
                    gated_mlp=False, act=c.get("hidden_act", "gelu_new"), lm_head_bias=True)
    renames = [(".self_attn.dense.", ".self_attn.o_proj."), (".mlp.fc1.", ".mlp.fc_in."), (".mlp.fc2.", ".mlp.fc_out."),
               ("model.final_layernorm.", "model.norm.")]


class Glm4ForCausalLM(VariantForCausalLM):
    def options(self, c):
        return dict(structure="sandwich", eps=c.get("rms_norm_eps", 1e-5), qkv_bias=bool(c.get("attention_bias", True)), neox=False)
    renames = [(".post_attention_layernorm.", ".pre_mlp_norm."), (".post_self_attn_layernorm.", ".post_attn_norm."),
               (".post_mlp_layernorm.", ".post_mlp_norm.")]


class NemotronForCausalLM(VariantForCausalLM):
    def options(self, c):
        return dict(norm="ln1p", eps=c.get("norm_eps", 1e-5), gated_mlp=False, act="relu2", qkv_bias=bool(c.get("attention_bias")),
                    o_bias=bool(c.get("attention_bias")), mlp_bias=bool(c.get("mlp_bias")))
    renames = [(".post_attention_layernorm.", ".pre_mlp_norm."), (".mlp.up_proj.", ".mlp.fc_in."), (".mlp.down_proj.", ".mlp.fc_out.")]


class Gemma3ForCausalLM(VariantForCausalLM):
    def options(self, c):
        c.setdefault("tie_word_embeddings", True)
        types = c.get("layer_types")
        pat = c.get("sliding_window_pattern") or c.get("_sliding_window_pattern") or 6
        return dict(norm="rms1p", structure="sandwich", eps=c.get("rms_norm_eps", 1e-6), qk_norm="head1p",
                    act=c.get("hidden_activation", "gelu_pytorch_tanh"), scale=float(c.get("query_pre_attn_scalar", 256)) ** -0.5,
                    embed_scale=float(c["hidden_size"]) ** 0.5, qkv_bias=bool(c.get("attention_bias")), o_bias=bool(c.get("attention_bias")),
                    sliding=(lambda i: types[i] == "sliding_attention") if types else (lambda i: (i + 1) % pat != 0))
    renames = [(".post_attention_layernorm.", ".post_attn_norm."), (".pre_feedforward_layernorm.", ".pre_mlp_norm."),
               (".post_feedforward_layernorm.", ".post_mlp_norm.")]


def _types_sliding(c):
    types = c.get("layer_types")
    pat = c.get("sliding_window_pattern") or 4
    return (lambda i: types[i] == "sliding_attention") if types else (lambda i: (i + 1) % pat != 0)


class SmolLM3ForCausalLM(VariantForCausalLM):
    """Llama with rotary switched off on every n-th layer (reference: models/smollm3.py)."""
    def options(self, c):
        nope = c.get("no_rope_layers")
        return dict(eps=c.get("rms_norm_eps", 1e-6), qkv_bias=bool(c.get("attention_bias")), o_bias=bool(c.get("attention_bias")),
                    mlp_bias=bool(c.get("mlp_bias")),
                    use_rope=(lambda i, sl: bool(nope[i])) if nope else (lambda i, sl: (i + 1) % (c.get("no_rope_layer_interval") or 4) != 0))
    renames = [(".post_attention_layernorm.", ".pre_mlp_norm.")]


class Cohere2ForCausalLM(CohereForCausalLM):
    """Command-R7B / Command-A: Cohere block, 3 sliding-window layers (with rope) : 1 global layer (no rope)
    (reference: models/commandr.py Cohere2)."""
    def options(self, c):
        o = super().options(c)
        sl = _types_sliding(c)
        o.update(sliding=sl, use_rope=lambda i, sliding: sliding)
        return o


class Olmo3ForCausalLM(Olmo2ForCausalLM):
    """OLMo-2 block with sliding-window layers interleaved (reference: models/olmo2.py Olmo3)."""
    def options(self, c):
        o = super().options(c)
        o.update(sliding=_types_sliding(c))
        return o


class Exaone4ForCausalLM(VariantForCausalLM):
    """Post-norm block, per-head q/k RMSNorm, 3 local (rope) : 1 global (no rope) when a sliding window is configured
    (reference: models/exaone4.py)."""
    def options(self, c):
        hybrid = bool(c.get("sliding_window"))
        sl = _types_sliding(c) if hybrid else (lambda i: False)
        return dict(structure="post", eps=c.get("rms_norm_eps", 1e-5), qk_norm="head", sliding=sl,
                    use_rope=(lambda i, sliding: sliding) if hybrid else (lambda i, sliding: True))
    renames = [(".post_attention_layernorm.", ".post_attn_norm."), (".post_feedforward_layernorm.", ".post_mlp_norm.")]


class SeedOssForCausalLM(VariantForCausalLM):
    def options(self, c):
        return dict(eps=c.get("rms_norm_eps", 1e-6), qkv_bias=bool(c.get("attention_bias", True)),
                    o_bias=bool(c.get("attention_out_bias", False)), mlp_bias=bool(c.get("mlp_bias")))
    renames = [(".post_attention_layernorm.", ".pre_mlp_norm.")]


class ArceeForCausalLM(VariantForCausalLM):
    """Llama block with a plain squared-ReLU MLP (reference: models/arcee.py)."""
    def options(self, c):
        return dict(eps=c.get("rms_norm_eps", 1e-5), gated_mlp=False, act="relu2", qkv_bias=bool(c.get("attention_bias")),
                    o_bias=bool(c.get("attention_bias")), mlp_bias=bool(c.get("mlp_bias")))
    renames = [(".post_attention_layernorm.", ".pre_mlp_norm."), (".mlp.up_proj.", ".mlp.fc_in."), (".mlp.down_proj.", ".mlp.fc_out.")]


class HeliumForCausalLM(VariantForCausalLM):
    def options(self, c):
        return dict(eps=c.get("rms_norm_eps", 1e-8), neox=False, qkv_bias=bool(c.get("attention_bias")), o_bias=False,
                    mlp_bias=bool(c.get("mlp_bias")))
    renames = [(".post_attention_layernorm.", ".pre_mlp_norm.")]


class Ernie4_5ForCausalLM(VariantForCausalLM):
    def options(self, c):
        c.setdefault("tie_word_embeddings", True)
        return dict(eps=c.get("rms_norm_eps", 1e-5), neox=False, qkv_bias=bool(c.get("use_bias")), o_bias=bool(c.get("use_bias")),
                    mlp_bias=bool(c.get("use_bias")))
    renames = [(".post_attention_layernorm.", ".pre_mlp_norm.")]


class OlmoForCausalLM(VariantForCausalLM):
    """OLMo-1: Llama block with parameter-free LayerNorms and optional clamping of the q/k/v projections
    (reference: models/olmo.py)."""
    def options(self, c):
        return dict(norm="ln_plain", eps=1e-5, clip_qkv=c.get("clip_qkv"), qkv_bias=bool(c.get("attention_bias")),
                    o_bias=bool(c.get("attention_bias")))
    renames = [(".post_attention_layernorm.", ".pre_mlp_norm.")]


class StableLmForCausalLM(VariantForCausalLM):
    """StableLM-2 / StableLM-3B: LayerNorm (with bias), partial rotary, optional q/k/v biases, optional parallel residual
    (reference: models/stablelm.py).  Per-head q/k LayerNorm (stablelm-2-12b) is not implemented."""
    def options(self, c):
        if c.get("qk_layernorm"):
            raise NotImplementedError("StableLM with per-head q/k LayerNorm (stablelm-2-12b)")
        return dict(norm="ln", eps=c.get("layer_norm_eps", 1e-5), qkv_bias=bool(c.get("use_qkv_bias")),
                    structure="parallel" if c.get("use_parallel_residual") else "pre")
    renames = [(".post_attention_layernorm.", ".pre_mlp_norm.")]


class MinistralForCausalLM(VariantForCausalLM):
    """Mistral block with per-layer sliding-window / full attention (``layer_types``) (reference: models/mistral.py family)."""
    def options(self, c):
        types = c.get("layer_types")
        return dict(eps=c.get("rms_norm_eps", 1e-6), sliding=(lambda i: types[i] == "sliding_attention") if types else (lambda i: bool(c.get("sliding_window"))))
    renames = [(".post_attention_layernorm.", ".pre_mlp_norm.")]


class GlmForCausalLM(VariantForCausalLM):
    """GLM-4-9B (HF ``glm``): pre-norm Llama block, biased q/k/v, half-width interleaved rotary, fused gate_up tensor
    (reference: models/glm.py)."""
    def options(self, c):
        return dict(eps=c.get("rms_norm_eps", 1e-5), qkv_bias=bool(c.get("attention_bias", True)), neox=False)
    renames = [(".post_attention_layernorm.", ".pre_mlp_norm.")]


class VaultGemmaForCausalLM(VariantForCausalLM):
    """VaultGemma: the Gemma-2 recipe (scaled embeddings, (1+w) RMSNorm, tanh-GELU gated MLP, attention / logit soft caps,
    alternating sliding-window layers) with pre-norms only (HF ``modeling_vaultgemma``)."""
    def options(self, c):
        c.setdefault("tie_word_embeddings", True)
        types = c.get("layer_types")
        return dict(norm="rms1p", eps=c.get("rms_norm_eps", 1e-6), act=c.get("hidden_activation", "gelu_pytorch_tanh"),
                    scale=float(c.get("query_pre_attn_scalar", 256)) ** -0.5, embed_scale=float(c["hidden_size"]) ** 0.5,
                    qkv_bias=bool(c.get("attention_bias")), o_bias=bool(c.get("attention_bias")),
                    sliding=(lambda i: types[i] == "sliding_attention") if types else (lambda i: i % 2 == 0))
    renames = [(".pre_feedforward_layernorm.", ".pre_mlp_norm.")]


class HunYuanDenseV1ForCausalLM(VariantForCausalLM):
    """HunYuan dense: Llama block whose per-head q / k RMSNorms are applied AFTER the rotary embedding
    (reference: models/hunyuan_v1.py)."""
    def options(self, c):
        return dict(eps=c.get("rms_norm_eps", 1e-5), qk_norm="head_post_rope" if c.get("use_qk_norm", True) else None,
                    qkv_bias=bool(c.get("attention_bias")), o_bias=bool(c.get("attention_bias")), mlp_bias=bool(c.get("mlp_bias")))
    renames = [(".post_attention_layernorm.", ".pre_mlp_norm."), (".query_layernorm.", ".q_norm."), (".key_layernorm.", ".k_norm.")]


class PhimoeForCausalLM(VariantForCausalLM):
    """Phi-3.5-MoE: LayerNorm (with bias) block, biased attention projections, 16 experts / top-2 with the *sparsemixer*
    router (experts picked one at a time, weights from a softmax over the logits within a jitter band of the running
    maximum), biased LM head (reference: models/phimoe.py)."""
    expert_names = {"w1": ("w13_weight", "w1"), "w3": ("w13_weight", "w3"), "w2": ("w2_weight", "w2")}
    expert_container = "mlp.experts"
    renames = [(".post_attention_layernorm.", ".pre_mlp_norm."), (".block_sparse_moe.", ".mlp."), (".mlp.router.", ".mlp.gate.")]

    def options(self, c):
        from vllm_b200.layers.moe import sparsemixer_routing
        return dict(norm="ln", eps=c.get("rms_norm_eps", 1e-5), qkv_bias=bool(c.get("attention_bias", True)),
                    o_bias=bool(c.get("attention_bias", True)), lm_head_bias=bool(c.get("lm_head_bias", True)),
                    moe=dict(n_experts=c["num_local_experts"], top_k=c["num_experts_per_tok"], renormalize=False,
                             routing=sparsemixer_routing(float(c.get("router_jitter_noise", 0.01)))))


class HunYuanMoEV1ForCausalLM(VariantForCausalLM):
    """HunYuan MoE (A13B family): the HunYuan attention (per-head q / k norms after the rotation), softmax top-k router with
    renormalised weights, and an always-on shared MLP next to the routed experts (reference: models/hunyuan_v1.py)."""
    expert_names = {"w1": ("w13_weight", "w1"), "w3": ("w13_weight", "w3"), "w2": ("w2_weight", "w2")}
    expert_container = "mlp.experts"
    renames = [(".post_attention_layernorm.", ".pre_mlp_norm."), (".query_layernorm.", ".q_norm."), (".key_layernorm.", ".k_norm."),
               (".mlp.gate.wg.", ".mlp.gate."), (".mlp.shared_mlp.", ".mlp.shared_expert.")]

    def options(self, c):
        ne, tk = c["num_experts"], c.get("moe_topk", 1)
        if not isinstance(ne, int) or not isinstance(tk, int):
            raise NotImplementedError("HunYuan MoE with per-layer expert counts")
        return dict(eps=c.get("rms_norm_eps", 1e-5), qk_norm="head_post_rope" if c.get("use_qk_norm", True) else None,
                    qkv_bias=bool(c.get("attention_bias")), o_bias=bool(c.get("attention_bias")),
                    moe=dict(n_experts=ne, top_k=tk, renormalize=True, inter=c.get("moe_intermediate_size") or c["intermediate_size"],
                             shared_inter=c["intermediate_size"] if c.get("use_mixed_mlp_moe", True) else 0))


class ApertusForCausalLM(VariantForCausalLM):
    """Apertus: per-head q / k RMSNorm, plain (non-gated) MLP with the learnable xIELU activation
    (reference: models/apertus.py)."""
    def options(self, c):
        return dict(eps=c.get("rms_norm_eps", 1e-5), qk_norm="head", gated_mlp=False, act=c.get("hidden_act", "xielu"),
                    qkv_bias=bool(c.get("attention_bias")), o_bias=bool(c.get("attention_bias")), mlp_bias=bool(c.get("mlp_bias")))
    renames = [(".attention_layernorm.", ".input_layernorm."), (".feedforward_layernorm.", ".pre_mlp_norm."),
               (".mlp.up_proj.", ".mlp.fc_in."), (".mlp.down_proj.", ".mlp.fc_out.")]


class CwmForCausalLM(VariantForCausalLM):
    """Code World Model: the Llama-3 block with sliding-window and full-attention layers interleaved by ``layer_types``
    (reference: models/cwm.py)."""
    def options(self, c):
        types = c.get("layer_types")
        return dict(eps=c.get("rms_norm_eps", 1e-5), mlp_bias=bool(c.get("mlp_bias")),
                    sliding=(lambda i: types[i] == "sliding_attention") if types else (lambda i: False))
    renames = [(".post_attention_layernorm.", ".pre_mlp_norm.")]


class Jais2ForCausalLM(VariantForCausalLM):
    """Jais-2: LayerNorm (with bias), biased projections, plain squared-ReLU MLP (reference: models/jais2.py)."""
    def options(self, c):
        return dict(norm="ln", eps=c.get("layer_norm_eps", 1e-5), gated_mlp=False, act=c.get("hidden_act", "relu2"),
                    qkv_bias=bool(c.get("attention_bias", True)), o_bias=bool(c.get("attention_bias", True)),
                    mlp_bias=bool(c.get("mlp_bias", True)))
    renames = [(".post_attention_layernorm.", ".pre_mlp_norm."), (".mlp.up_proj.", ".mlp.fc_in."), (".mlp.down_proj.", ".mlp.fc_out.")]


class Ministral3ForCausalLM(VariantForCausalLM):
    """Ministral-3 / Mistral-Large-3 text block: Mistral with YaRN rotary and the Llama-4 length-dependent query scale
    ``1 + β log(1 + ⌊pos / L₀⌋)`` (reference: models/mistral.py ``Ministral3``, HF ``modeling_ministral3``)."""
    def options(self, c):
        rp = c.get("rope_parameters") or c.get("rope_scaling") or {}
        beta = rp.get("llama_4_scaling_beta")
        L0 = rp.get("original_max_position_embeddings") or c.get("max_position_embeddings", 8192)
        return dict(eps=c.get("rms_norm_eps", 1e-5), q_pos_scale=(float(beta), float(L0)) if beta else None,
                    sliding=lambda i: bool(c.get("sliding_window")))
    renames = [(".post_attention_layernorm.", ".pre_mlp_norm.")]


class FlexOlmoForCausalLM(VariantForCausalLM):
    """FlexOlmo: the OLMo-2 block (sub-layer output norms, whole-projection q / k norms) around a softmax top-k expert
    layer (reference: models/flex_olmo.py)."""
    expert_names = {"w1": ("w13_weight", "w1"), "w3": ("w13_weight", "w3"), "w2": ("w2_weight", "w2")}
    expert_container = "mlp.experts"
    renames = [(".post_attention_layernorm.", ".post_attn_norm."), (".post_feedforward_layernorm.", ".post_mlp_norm.")]

    def options(self, c):
        return dict(structure="post", eps=c.get("rms_norm_eps", 1e-6), qk_norm="full", qkv_bias=bool(c.get("attention_bias")),
                    o_bias=bool(c.get("attention_bias")),
                    moe=dict(n_experts=c["num_experts"], top_k=c["num_experts_per_tok"], renormalize=bool(c.get("norm_topk_prob", False))))


_EXPERTS_ANY = {"w1": ("w13_weight", "w1"), "w3": ("w13_weight", "w3"), "w2": ("w2_weight", "w2"),
                "gate_proj": ("w13_weight", "w1"), "up_proj": ("w13_weight", "w3"), "down_proj": ("w2_weight", "w2")}
_BIAS_RENAMES = [(".mlp.gate.e_score_correction_bias", ".mlp.experts.e_score_correction_bias"),
                 (".mlp.e_score_correction_bias", ".mlp.experts.e_score_correction_bias")]


class MiniMaxM2ForCausalLM(VariantForCausalLM):
    """MiniMax-M2: whole-projection q / k RMSNorm, partial rotary, every layer sparse — sigmoid scores, selection with a
    learned correction bias, mixing weights renormalised over the chosen experts (reference: models/minimax_m2.py)."""
    expert_names = _EXPERTS_ANY
    expert_container = "mlp.experts"
    renames = [(".post_attention_layernorm.", ".pre_mlp_norm."), (".block_sparse_moe.", ".mlp.")] + _BIAS_RENAMES

    def options(self, c):
        rp = c.get("rope_parameters") or {}
        if "partial_rotary_factor" not in c and "rotary_dim" in c:
            c["partial_rotary_factor"] = c["rotary_dim"] / (c.get("head_dim") or c["hidden_size"] // c["num_attention_heads"])
        c.setdefault("partial_rotary_factor", rp.get("partial_rotary_factor", 1.0))
        return dict(eps=c.get("rms_norm_eps", 1e-6), qk_norm="full" if c.get("use_qk_norm", True) else None,
                    moe=dict(n_experts=c["num_local_experts"], top_k=c["num_experts_per_tok"], renormalize=True,
                             kw=dict(use_grouped_topk=True, n_group=1, topk_group=1, scoring_func="sigmoid", e_score_bias=True)))


class ExaoneMoeForCausalLM(VariantForCausalLM):
    """EXAONE MoE (K-EXAONE): pre-norm block, per-head q / k RMSNorm, sliding-window layers with rotary and global layers
    without, leading dense layers then DeepSeek-V3 style experts (sigmoid scores + correction bias, grouped top-k, shared
    expert, routed scaling) (reference: models/exaone_moe.py)."""
    expert_names = _EXPERTS_ANY
    expert_container = "mlp.experts"
    renames = [(".post_attention_layernorm.", ".pre_mlp_norm."), (".mlp.shared_experts.", ".mlp.shared_expert.")] + _BIAS_RENAMES

    def options(self, c):
        hybrid = bool(c.get("sliding_window"))
        sl = _types_sliding(c) if hybrid else (lambda i: False)
        kinds = c.get("mlp_layer_types")
        dense_k = c.get("first_k_dense_replace", 1)
        moe = dict(n_experts=c["num_experts"], top_k=c["num_experts_per_tok"], renormalize=bool(c.get("norm_topk_prob", True)),
                   inter=c["moe_intermediate_size"], shared_inter=c["moe_intermediate_size"] * int(c.get("num_shared_experts", 0)),
                   kw=dict(use_grouped_topk=True, n_group=c.get("n_group") or 1, topk_group=c.get("topk_group") or 1,
                           scoring_func="sigmoid", e_score_bias=True, routed_scaling_factor=float(c.get("routed_scaling_factor", 1.0))))
        return dict(eps=c.get("rms_norm_eps", 1e-5), qk_norm="head", sliding=sl,
                    use_rope=(lambda i, sliding: sliding) if hybrid else (lambda i, sliding: True),
                    moe=lambda i: moe if (kinds[i] == "sparse" if kinds else i >= dense_k) else None)


class DbrxForCausalLM(VariantForCausalLM):
    """DBRX: bias-free LayerNorms, fused and clamped Wqkv, 16 experts / top-4 with the mixing weights normalised by their
    p-norm; experts stored as three [E·I, D] tensors (reference: models/dbrx.py)."""
    expert_names = _EXPERTS_ANY
    expert_container = "mlp.experts"

    def options(self, c):
        f = c.get("ffn_config") or {}
        p = f.get("moe_normalize_expert_weights", 1.0)
        if p not in (None, 1, 1.0):
            raise NotImplementedError(f"DBRX with moe_normalize_expert_weights={p} (only the 1-norm / none)")
        return dict(norm="ln_nobias", eps=1e-5, clip_qkv=c.get("clip_qkv"), act=c.get("hidden_act", "silu"),
                    moe=dict(n_experts=c["num_local_experts"], top_k=c["num_experts_per_tok"], renormalize=p is not None))

    def load_weights(self, weights):
        c = self.cfg
        nh, nkv = c["num_attention_heads"], c["num_key_value_heads"]
        hd = c["hidden_size"] // nh
        E, I, D = c["num_local_experts"], c["intermediate_size"], c["hidden_size"]

        def mapped():
            for name, w in weights:
                name = (name.replace("transformer.wte.", "model.embed_tokens.").replace("transformer.norm_f.", "model.norm.")
                        .replace("transformer.blocks.", "model.layers.").replace(".norm_attn_norm.norm_1.", ".input_layernorm.")
                        .replace(".norm_attn_norm.norm_2.", ".pre_mlp_norm.").replace(".norm_attn_norm.attn.", ".self_attn.")
                        .replace(".out_proj.", ".o_proj.").replace(".ffn.router.layer.", ".mlp.gate.").replace(".ffn.", ".mlp."))
                if ".self_attn.Wqkv." in name:
                    q, k, v = w.split([nh * hd, nkv * hd, nkv * hd], dim=0)
                    for part, t in (("q_proj", q), ("k_proj", k), ("v_proj", v)):
                        yield name.replace("Wqkv", part), t
                elif ".mlp.experts.mlp." in name:
                    kind = name.rsplit(".", 1)[1]                        # w1 (gate) | v1 (up) | w2 (down, stored [I, D])
                    base = name[: name.index(".mlp.experts.")] + ".mlp.experts."
                    w3 = w.reshape(E, I, D)
                    for e in range(E):
                        if kind == "w2":
                            yield f"{base}{e}.down_proj.weight", w3[e].t()
                        else:
                            yield f"{base}{e}.{'gate_proj' if kind == 'w1' else 'up_proj'}.weight", w3[e]
                else:
                    yield name, w
        return super().load_weights(mapped())


class AfmoeForCausalLM(VariantForCausalLM):
    """Arcee AFMoE (Trinity): sandwich norms, per-head q / k RMSNorm, sigmoid-gated attention output, sliding-window
    layers with rotary and global layers without, leading dense layers then sigmoid-routed experts with a selection bias,
    renormalised and scaled mixing weights and shared experts; optional muP embedding scale (reference: models/afmoe.py)."""
    expert_names = _EXPERTS_ANY
    expert_container = "mlp.experts"
    renames = [(".post_attention_layernorm.", ".post_attn_norm."), (".pre_mlp_layernorm.", ".pre_mlp_norm."),
               (".post_mlp_layernorm.", ".post_mlp_norm."), (".self_attn.gate_proj.", ".self_attn.out_gate."),
               (".mlp.router.gate.", ".mlp.gate."), (".mlp.expert_bias", ".mlp.experts.e_score_correction_bias"),
               (".mlp.shared_experts.", ".mlp.shared_expert.")]

    def options(self, c):
        types = c.get("layer_types")
        every = c.get("global_attn_every_n_layers") or 4
        sl = (lambda i: types[i] == "sliding_attention") if types else (lambda i: (i + 1) % every != 0)
        dense_k = c.get("num_dense_layers") or 0
        moe = dict(n_experts=c["num_experts"], top_k=c["num_experts_per_tok"], renormalize=True, inter=c["moe_intermediate_size"],
                   shared_inter=c["moe_intermediate_size"] * int(c.get("num_shared_experts") or 0),
                   kw=dict(use_grouped_topk=True, n_group=1, topk_group=1, scoring_func="sigmoid", e_score_bias=True,
                           routed_scaling_factor=float(c.get("route_scale") or 1.0)))
        return dict(structure="sandwich", eps=c.get("rms_norm_eps", 1e-5), qk_norm="head", attn_gate=True, sliding=sl,
                    use_rope=lambda i, sliding: sliding, qkv_bias=bool(c.get("attention_bias")), o_bias=bool(c.get("attention_bias")),
                    embed_scale=float(c["hidden_size"]) ** 0.5 if c.get("mup_enabled") else 1.0,
                    moe=lambda i: moe if i >= dense_k else None)


def _mix_in_expert_loader():
    # the per-expert / fused expert tensor loader lives with the Mixtral family; the variant block only renames
    from vllm_b200.models.mixtral import MixtralForCausalLM
    PhimoeForCausalLM.__bases__ = (VariantForCausalLM, MixtralForCausalLM)
    HunYuanMoEV1ForCausalLM.__bases__ = (VariantForCausalLM, MixtralForCausalLM)
    FlexOlmoForCausalLM.__bases__ = (VariantForCausalLM, MixtralForCausalLM)
    MiniMaxM2ForCausalLM.__bases__ = (VariantForCausalLM, MixtralForCausalLM)
    ExaoneMoeForCausalLM.__bases__ = (VariantForCausalLM, MixtralForCausalLM)
    AfmoeForCausalLM.__bases__ = (VariantForCausalLM, MixtralForCausalLM)
    DbrxForCausalLM.__bases__ = (VariantForCausalLM, MixtralForCausalLM)


_mix_in_expert_loader()
