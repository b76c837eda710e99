"""Llama-4 (text tower): interleaved NoPE / RoPE layers — rope layers use chunked-local attention and a weight-free
L2 q/k norm applied AFTER the rotation, NoPE layers attend globally with a position-dependent query temperature —
and MoE layers whose top-1 router score scales the expert INPUT, plus an always-on shared expert.

Reference parity: vllm/model_executor/models/llama4.py (``Llama4Attention``, ``Llama4MoE``), layers/attention chunked
local attention backend wrapper.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.nn as nn

from vllm_b200.config import VllmConfig
from vllm_b200.layers.attention import Attention
from vllm_b200.layers.common import RMSNorm
from vllm_b200.layers.linear import QKVParallelLinear, ReplicatedLinear, RowParallelLinear
from vllm_b200.layers.moe import FusedMoE
from vllm_b200.layers.rotary import get_rope
from vllm_b200.models.llama import LlamaForCausalLM, LlamaMLP, LlamaModel
from vllm_b200.parallel.state import get_tp_group


class Llama4Attention(nn.Module):
    def __init__(self, cfg: dict, mc, dtype, quant_config, prefix: str, idx: int) -> None:
        super().__init__()
        H, nh = cfg["hidden_size"], cfg["num_attention_heads"]
        nkv = cfg.get("num_key_value_heads") or nh
        self.hd = cfg.get("head_dim") or H // nh
        bias = bool(cfg.get("attention_bias", False))
        self.qkv_proj = QKVParallelLinear(H, self.hd, nh, nkv, bias, dtype, quant_config, f"{prefix}.qkv_proj")
        self.o_proj = RowParallelLinear(nh * self.hd, H, bias, dtype=dtype, quant_config=quant_config, prefix=f"{prefix}.o_proj")
        self.q_size, self.kv_size = self.qkv_proj.q_size, self.qkv_proj.kv_size
        nope = cfg.get("no_rope_layers")
        self.use_rope = bool(nope[idx]) if nope else ((idx + 1) % (cfg.get("nope_layer_interval") or 4) != 0)
        self.l2_norm = bool(cfg.get("use_qk_norm", True)) and self.use_rope
        self.eps = cfg.get("rms_norm_eps", 1e-5)
        self.temp_tuning = bool(cfg.get("attn_temperature_tuning", True)) and not self.use_rope
        self.floor_scale, self.attn_scale = float(cfg.get("floor_scale", 8192)), float(cfg.get("attn_scale", 0.1))
        rp = cfg.get("rope_parameters") or {}
        scaling = cfg.get("rope_scaling") or ({k: v for k, v in rp.items() if k != "rope_theta"}
                                               if rp.get("rope_type", "default") != "default" else None)
        max_pos = mc.max_model_len if mc.max_model_len > 0 else cfg.get("max_position_embeddings", 8192)
        self.rope = get_rope(self.hd, self.hd, max_pos, rp.get("rope_theta", cfg.get("rope_theta", 500000.0)), False, scaling) \
            if self.use_rope else None
        # the rotation is applied here (the L2 norm has to follow it), so the attention layer itself gets no rope
        self.attn = Attention(self.qkv_proj.num_heads, self.hd, self.hd ** -0.5, self.qkv_proj.num_kv_heads,
                              prefix=f"{prefix}.attn", rope=None,
                              chunk_size=int(cfg.get("attention_chunk_size") or 0) if self.use_rope else 0)

    def _l2(self, x: torch.Tensor) -> torch.Tensor:
        T = x.shape[0]
        xf = x.view(T, -1, self.hd).float()
        return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.eps)).to(x.dtype).view(T, -1)

    def forward(self, positions, hidden):
        q, k, v = self.qkv_proj(hidden).split([self.q_size, self.kv_size, self.kv_size], dim=-1)
        if self.rope is not None:
            q, k = self.rope(positions, q.contiguous(), k.contiguous())
        if self.l2_norm:
            q, k = self._l2(q), self._l2(k)
        if self.temp_tuning:
            s = torch.log1p(torch.floor((positions.float() + 1.0) / self.floor_scale)) * self.attn_scale + 1.0
            q = (q.view(q.shape[0], -1, self.hd) * s[:, None, None]).to(q.dtype).view(q.shape[0], -1)
        return self.o_proj(self.attn(q, k, v))


class Llama4MoE(nn.Module):
    def __init__(self, cfg: dict, dtype, quant_config, prefix: str, enable_ep: bool) -> None:
        super().__init__()
        H, E = cfg["hidden_size"], cfg["num_local_experts"]
        self.top_k = cfg.get("num_experts_per_tok", 1)
        assert self.top_k == 1, "the router score is applied to the expert input, which is only well defined for top-1"
        self.router = ReplicatedLinear(H, E, False, dtype, None, f"{prefix}.router")
        self.experts = FusedMoE(E, 1, H, cfg["intermediate_size"], dtype, renormalize=False, act=cfg.get("hidden_act", "silu"),
                                enable_ep=enable_ep, reduce_results=False, prefix=f"{prefix}.experts")
        self.shared_expert = LlamaMLP(H, cfg["intermediate_size"], cfg.get("hidden_act", "silu"), False, dtype, quant_config,
                                      f"{prefix}.shared_expert")
        self.shared_expert.down_proj.reduce_results = False
        self.experts.select_experts = self._select      # type: ignore[method-assign]

    @staticmethod
    def _select(router_logits: torch.Tensor):
        top, ids = router_logits.float().topk(1, dim=-1)
        return torch.ones_like(top), ids.to(torch.int32)        # the score was already folded into the input

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        logits = self.router(x)
        score = torch.sigmoid(logits.float().max(dim=-1, keepdim=True).values).to(x.dtype)
        return self.experts(x * score, logits) + self.shared_expert(x)        # per-rank partial sums


class Llama4DecoderLayer(nn.Module):
    def __init__(self, cfg, mc, dtype, quant_config, prefix: str, idx: int, enable_ep: bool) -> None:
        super().__init__()
        H, eps = cfg["hidden_size"], cfg.get("rms_norm_eps", 1e-5)
        self.self_attn = Llama4Attention(cfg, mc, dtype, quant_config, f"{prefix}.self_attn", idx)
        moe_layers = cfg.get("moe_layers")
        step = cfg.get("interleave_moe_layer_step", 1)
        self.is_moe = (idx in moe_layers) if moe_layers is not None else (step > 0 and (idx + 1) % step == 0)
        if self.is_moe:
            self.feed_forward = Llama4MoE(cfg, dtype, quant_config, f"{prefix}.feed_forward", enable_ep)
        else:
            self.feed_forward = LlamaMLP(H, cfg.get("intermediate_size_mlp") or cfg["intermediate_size"],
                                         cfg.get("hidden_act", "silu"), False, dtype, quant_config, f"{prefix}.feed_forward")
        self.input_layernorm, self.post_attention_layernorm = RMSNorm(H, eps, dtype), RMSNorm(H, eps, dtype)

    def forward(self, positions, hidden, residual, next_norm: Optional[nn.Module]):
        stream = residual + self.self_attn(positions, hidden)
        m = self.feed_forward(self.post_attention_layernorm(stream))
        if self.is_moe and get_tp_group().world_size > 1:
            m = get_tp_group().all_reduce(m)
        stream = stream + m
        return (next_norm(stream) if next_norm is not None else stream), stream


class Llama4Model(LlamaModel):
    def __init__(self, vllm_config: VllmConfig, prefix: str = "model", quant_config=None) -> None:
        self._enable_ep = bool(vllm_config.parallel_config.enable_expert_parallel)
        super().__init__(vllm_config, prefix, quant_config)

    def _make_layer(self, cfg, mc, dtype, quant_config, prefix, idx):
        return Llama4DecoderLayer(cfg, mc, dtype, quant_config, prefix, idx, self._enable_ep)


class Llama4ForCausalLM(LlamaForCausalLM):
    model_cls = Llama4Model

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        params = dict(self.named_parameters())
        modules = dict(self.named_modules())
        loaded: set[str] = set()
        rest = []
        for name, w in weights:
            if name.startswith("language_model."):
                name = name[len("language_model."):]
            if ".feed_forward.experts." in name:
                n = self._localize(name)
                if n is None:
                    continue
                mod = modules[n.rsplit(".", 1)[0]]
                if n.endswith("gate_up_proj"):                       # [E, H, 2I] → rows [gate | up] per expert
                    I = w.shape[2] // 2
                    for e in range(w.shape[0]):
                        mod.weight_loader(params[n.replace("gate_up_proj", "w13_weight")], w[e, :, :I].t(), "w1", e)
                        mod.weight_loader(params[n.replace("gate_up_proj", "w13_weight")], w[e, :, I:].t(), "w3", e)
                    loaded.add(n.replace("gate_up_proj", "w13_weight"))
                elif n.endswith("down_proj"):                        # [E, I, H]
                    for e in range(w.shape[0]):
                        mod.weight_loader(params[n.replace("down_proj", "w2_weight")], w[e].t(), "w2", e)
                    loaded.add(n.replace("down_proj", "w2_weight"))
                continue
            rest.append((name, w))
        return loaded | super().load_weights(rest)
