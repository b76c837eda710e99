"""OLMo-hybrid: gated-delta-net layers (pre-norm) interleaved with OLMo-2 style attention layers (norms on the sub-layer
OUTPUTS, RMSNorm over the whole q / k projections).

The delta-net mixer is the Qwen3-Next one (``models/qwen3_next.py``: pooled conv + recurrent state, CUDA scan kernels) with
three differences handled here: separate q / k / v / gate / a / b projections and three depthwise convolutions in the
checkpoint (re-packed at load into the mixer's fused, key-head-interleaved tensors), write strength β ∈ (0, 2)
(``linear_allow_neg_eigval``), and a fixed 1e-5 epsilon in the gated output norm.
Reference parity: vllm/model_executor/models/olmo_hybrid.py, HF ``modeling_olmo_hybrid``.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.nn as nn

from vllm_b200.config import VllmConfig
from vllm_b200.layers.attention import Attention
from vllm_b200.layers.common import RMSNorm
from vllm_b200.layers.linear import QKVParallelLinear, RowParallelLinear
from vllm_b200.layers.rotary import get_rope
from vllm_b200.models.llama import LlamaForCausalLM, LlamaMLP, LlamaModel
from vllm_b200.models.llama_variants import _Identity
from vllm_b200.models.qwen3_next import GatedDeltaNet
from vllm_b200.parallel.state import get_tp_group


class _Attn(nn.Module):
    def __init__(self, cfg, mc, dtype, prefix) -> None:
        super().__init__()
        H, nh = cfg["hidden_size"], cfg["num_attention_heads"]
        nkv = cfg.get("num_key_value_heads") or nh
        hd = cfg.get("head_dim") or H // nh
        bias = bool(cfg.get("attention_bias", False))
        eps = cfg.get("rms_norm_eps", 1e-6)
        self.qkv_proj = QKVParallelLinear(H, hd, nh, nkv, bias, dtype, None, f"{prefix}.qkv_proj")
        self.o_proj = RowParallelLinear(nh * hd, H, bias, dtype=dtype, prefix=f"{prefix}.o_proj")
        self.sizes = [self.qkv_proj.q_size, self.qkv_proj.kv_size, self.qkv_proj.kv_size]
        self.q_norm, self.k_norm = RMSNorm(nh * hd, eps, dtype), RMSNorm(nkv * hd, eps, dtype)
        rp = cfg.get("rope_parameters") or {}
        rope = None
        if rp.get("rope_theta") is not None:
            max_pos = mc.max_model_len if mc.max_model_len > 0 else cfg.get("max_position_embeddings", 8192)
            scaling = {k: v for k, v in rp.items() if k != "rope_theta"} if rp.get("rope_type", "default") != "default" else None
            rope = get_rope(hd, hd, max_pos, rp["rope_theta"], True, scaling)
        self.attn = Attention(self.qkv_proj.num_heads, hd, hd ** -0.5, self.qkv_proj.num_kv_heads, prefix=f"{prefix}.attn", rope=rope)

    def forward(self, positions, x):
        q, k, v = self.qkv_proj(x).split(self.sizes, dim=-1)
        return self.o_proj(self.attn(self.q_norm(q.contiguous()), self.k_norm(k.contiguous()), v, positions))


class OlmoHybridLayer(nn.Module):
    def __init__(self, cfg, mc, dtype, quant_config, prefix, idx) -> None:
        super().__init__()
        H, eps = cfg["hidden_size"], cfg.get("rms_norm_eps", 1e-6)
        self.is_attn = cfg["layer_types"][idx] == "full_attention"
        self.mlp = LlamaMLP(H, cfg["intermediate_size"], cfg.get("hidden_act", "silu"), False, dtype, quant_config, f"{prefix}.mlp")
        self.post_attention_layernorm = RMSNorm(H, eps, dtype)
        if self.is_attn:
            self.self_attn = _Attn(cfg, mc, dtype, f"{prefix}.self_attn")
            self.input_layernorm = _Identity()
            self.post_feedforward_layernorm = RMSNorm(H, eps, dtype)
        else:
            self.linear_attn = GatedDeltaNet(cfg, dtype, f"{prefix}.linear_attn")
            self.linear_attn.eps = 1e-5
            self.linear_attn.beta_scale = 2.0 if cfg.get("linear_allow_neg_eigval", True) else 1.0
            self.input_layernorm = RMSNorm(H, eps, dtype)

    def forward(self, positions, hidden, residual, next_norm: Optional[nn.Module]):
        if self.is_attn:
            x = residual + self.post_attention_layernorm(self.self_attn(positions, hidden))
            x = x + self.post_feedforward_layernorm(self.mlp(x))
        else:
            x = residual + self.linear_attn(hidden)
            x = x + self.mlp(self.post_attention_layernorm(x))
        return (next_norm(x) if next_norm is not None else x), x


class OlmoHybridModel(LlamaModel):
    def __init__(self, vllm_config: VllmConfig, prefix: str = "model", quant_config=None) -> None:
        assert get_tp_group().world_size == 1, "delta-net hybrids run replicated (tensor_parallel_size=1) for now"
        super().__init__(vllm_config, prefix, quant_config)

    def _make_layer(self, cfg, mc, dtype, quant_config, prefix, idx):
        return OlmoHybridLayer(cfg, mc, dtype, quant_config, prefix, idx)


class OlmoHybridForCausalLM(LlamaForCausalLM):
    model_cls = OlmoHybridModel

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        c = self.cfg
        nk, nv, dk, dv = c["linear_num_key_heads"], c["linear_num_value_heads"], c["linear_key_head_dim"], c["linear_value_head_dim"]
        r = nv // nk
        params = dict(self.named_parameters())
        pending: dict[str, dict[str, torch.Tensor]] = {}
        loaded: set[str] = set()
        need = ("q_proj", "k_proj", "v_proj", "g_proj", "a_proj", "b_proj", "q_conv1d", "k_conv1d", "v_conv1d")

        def rest():
            for name, w in weights:
                if ".linear_attn." not in name:
                    yield name, w
                    continue
                n = self._localize(name)
                if n is None:
                    continue
                base, leaf = n.split(".linear_attn.", 1)
                base += ".linear_attn."
                key = leaf.split(".")[0]
                if key in need:
                    got = pending.setdefault(base, {})
                    got[key] = w
                    if len(got) == len(need):
                        H = got["q_proj"].shape[1]
                        qkvz = torch.cat([got["q_proj"].view(nk, dk, H), got["k_proj"].view(nk, dk, H),
                                          got["v_proj"].view(nk, r * dv, H), got["g_proj"].view(nk, r * dv, H)], dim=1).reshape(-1, H)
                        ba = torch.cat([got["b_proj"].view(nk, r, H), got["a_proj"].view(nk, r, H)], dim=1).reshape(-1, H)
                        conv = torch.cat([got[k].reshape(got[k].shape[0], -1) for k in ("q_conv1d", "k_conv1d", "v_conv1d")], dim=0)
                        for t, v in ((base + "in_proj_qkvz.weight", qkvz), (base + "in_proj_ba.weight", ba), (base + "conv_weight", conv)):
                            params[t].data.copy_(v)
                            loaded.add(t)
                        del pending[base]
                    continue
                t = base + {"o_norm.weight": "norm_weight", "o_proj.weight": "out_proj.weight"}.get(leaf, leaf)
                if t in params:
                    params[t].data.copy_(w)
                    loaded.add(t)
        out = super().load_weights(rest())
        assert not pending, f"incomplete delta-net projections for {sorted(pending)}"
        return out | loaded
