"""Falcon-H1: every layer runs a Mamba-2 mixer and a GQA attention IN PARALLEL on the same normed input, sums them into
the residual, then a gated MLP.  The family is muP-parameterised: a dozen scalar multipliers (embedding, per-section
input-projection, SSM in / out, attention in / out, key, MLP gate / down, LM head) scale activations at fixed points.

All multipliers except the embedding and LM-head ones sit next to a linear layer, so the loader FOLDS them into that
layer's weight (and bias) once — the forward pass is the plain Bamba-style block with no extra elementwise kernels.
Reference parity: vllm/model_executor/models/falcon_h1.py, HF ``modeling_falcon_h1`` (``compute_mup_vector``).
"""
from __future__ import annotations

from typing import Iterable

import torch
import torch.nn as nn

from vllm_b200.config import VllmConfig
from vllm_b200.layers.common import LogitsProcessor, RMSNorm
from vllm_b200.layers.mamba import Mamba2Mixer
from vllm_b200.models.llama import LlamaAttention, LlamaMLP
from vllm_b200.models.mamba2 import BambaForCausalLM


def _ssm_dims(cfg: dict) -> tuple[int, int, int]:
    d_ssm = cfg.get("mamba_d_ssm") or int(cfg.get("mamba_expand", 2) * cfg["hidden_size"])
    nh = cfg["mamba_n_heads"]
    hd = cfg.get("mamba_d_head")
    hd = d_ssm // nh if hd in (None, "auto") else int(hd)
    return d_ssm, nh, hd


class FalconH1Layer(nn.Module):
    def __init__(self, cfg: dict, mc, dtype, prefix: str, idx: int) -> None:
        super().__init__()
        H, eps = cfg["hidden_size"], cfg.get("rms_norm_eps", 1e-5)
        self.self_attn = LlamaAttention(cfg, mc, dtype, None, f"{prefix}.self_attn", idx)
        _, nh, hd = _ssm_dims(cfg)
        G = cfg.get("mamba_n_groups", 1)
        lim = cfg.get("time_step_limit") or (0.0, float("inf"))
        gate_mode = bool(cfg.get("mamba_norm_before_gate", True)) if cfg.get("mamba_rms_norm", False) else None
        self.mamba = Mamba2Mixer(H, nh, hd, cfg.get("mamba_d_state", 256), cfg.get("mamba_d_conv", 4), G,
                                 bool(cfg.get("mamba_proj_bias", False)), bool(cfg.get("mamba_conv_bias", True)),
                                 cfg.get("mamba_chunk_size", 256), eps, dtype, f"{prefix}.mamba",
                                 time_step_limit=(float(lim[0]), float(lim[1])), norm_before_gate=gate_mode, norm_groups=G)
        self.feed_forward = LlamaMLP(H, cfg["intermediate_size"], cfg.get("hidden_act", "silu"), bool(cfg.get("mlp_bias", False)),
                                     dtype, None, f"{prefix}.feed_forward")
        self.input_layernorm = RMSNorm(H, eps, dtype)
        self.pre_ff_layernorm = RMSNorm(H, eps, dtype)

    def forward(self, positions, x):
        h = self.input_layernorm(x)
        x = x + self.mamba(h) + self.self_attn.o_proj(self.self_attn(positions, h))
        return x + self.feed_forward(self.pre_ff_layernorm(x))


class FalconH1ForCausalLM(BambaForCausalLM):
    layer_cls = FalconH1Layer

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        cfg = vllm_config.model_config.hf_config
        rp = cfg.get("rope_parameters") or {}
        cfg.setdefault("rope_theta", rp.get("rope_theta", 100000.0))
        super().__init__(vllm_config, prefix)
        self.embed_scale = float(cfg.get("embedding_multiplier") or 1.0)
        self.logits_processor = LogitsProcessor(cfg["vocab_size"], scale=float(cfg.get("lm_head_multiplier") or 1.0))

    def embed_input_ids(self, ids):
        x = self.embed_tokens(ids)
        return x * self.embed_scale if self.embed_scale != 1.0 else x

    def forward(self, input_ids, positions, intermediate_tensors=None, inputs_embeds=None):
        x = inputs_embeds if inputs_embeds is not None else self.embed_input_ids(input_ids)
        for layer in self.layers:
            x = layer(positions, x)
        return self.final_layernorm(x)

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        c = self.cfg
        d_ssm, nh, _ = _ssm_dims(c)
        gn = c.get("mamba_n_groups", 1) * c.get("mamba_d_state", 256)
        zx = [float(v) for v in (c.get("ssm_multipliers") or [1.0] * 5)]
        mup = torch.cat([torch.full((d_ssm,), zx[0]), torch.full((d_ssm,), zx[1]), torch.full((gn,), zx[2]),
                         torch.full((gn,), zx[3]), torch.full((nh,), zx[4])])
        ssm_in, ssm_out = float(c.get("ssm_in_multiplier") or 1.0), float(c.get("ssm_out_multiplier") or 1.0)
        a_in, a_out = float(c.get("attention_in_multiplier") or 1.0), float(c.get("attention_out_multiplier") or 1.0)
        key_m = float(c.get("key_multiplier") or 1.0)
        gate_m, down_m = [float(v) for v in (c.get("mlp_multipliers") or [1.0, 1.0])]

        def scale(name: str, w: torch.Tensor) -> torch.Tensor:
            """(y = f(s_in · x) · s_out with f linear)  ⇒  weight · s_in · s_out, bias · s_out."""
            bias = name.endswith(".bias")
            if ".mamba.in_proj." in name:
                v = mup.to(w.dtype).to(w.device)
                return w * v if bias else w * (v[:, None] * ssm_in)
            if ".mamba.out_proj." in name:
                return w * ssm_out
            if ".self_attn.k_proj." in name:
                return w * key_m if bias else w * (key_m * a_in)
            if ".self_attn.q_proj." in name or ".self_attn.v_proj." in name:
                return w if bias else w * a_in
            if ".self_attn.o_proj." in name:
                return w * a_out
            if ".feed_forward.gate_proj." in name:
                return w * gate_m
            if ".feed_forward.down_proj." in name:
                return w * down_m
            return w
        return super().load_weights((n, scale(n, w)) for n, w in weights)
