"""Granite MoE family: GraniteMoe (3.x MoE), GraniteMoeShared (+ always-on shared MLP) and GraniteMoeHybrid (Granite 4.0:
Mamba-2 and attention layers, optional MoE, shared MLP, NoPE or RoPE) — one layer class, Granite's four multipliers.

Reference parity: vllm/model_executor/models/granitemoe.py, granitemoeshared.py, granitemoehybrid.py.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from vllm_b200.config import VllmConfig
from vllm_b200.layers.attention import Attention
from vllm_b200.layers.common import LogitsProcessor, RMSNorm
from vllm_b200.layers.linear import MergedColumnParallelLinear, QKVParallelLinear, ReplicatedLinear, RowParallelLinear
from vllm_b200.layers.mamba import Mamba2Mixer
from vllm_b200.layers.moe import FusedMoE
from vllm_b200.layers.rotary import get_rope
from vllm_b200.models.llama import LlamaForCausalLM, LlamaModel
from vllm_b200.parallel.state import get_tp_group


class _Attn(nn.Module):
    def __init__(self, cfg, mc, dtype, prefix, sliding: bool = False) -> None:
        super().__init__()
        H, nh = cfg["hidden_size"], cfg["num_attention_heads"]
        nkv = cfg.get("num_key_value_heads") or nh
        hd = H // nh
        bias = bool(cfg.get("attention_bias", False))
        self.qkv_proj = QKVParallelLinear(H, hd, nh, nkv, bias, dtype, None, f"{prefix}.qkv_proj")
        self.o_proj = RowParallelLinear(H, H, bias, dtype=dtype, prefix=f"{prefix}.o_proj")
        self.sizes = [self.qkv_proj.q_size, self.qkv_proj.kv_size, self.qkv_proj.kv_size]
        rope = None
        if cfg.get("position_embedding_type", "rope") == "rope":
            rp = cfg.get("rope_parameters") or {}
            max_pos = mc.max_model_len if mc.max_model_len > 0 else cfg.get("max_position_embeddings", 8192)
            rope = get_rope(hd, hd, max_pos, rp.get("rope_theta", cfg.get("rope_theta", 10000.0)), True, cfg.get("rope_scaling"))
        self.attn = Attention(self.qkv_proj.num_heads, hd, float(cfg.get("attention_multiplier") or hd ** -0.5),
                              self.qkv_proj.num_kv_heads, prefix=f"{prefix}.attn", rope=rope,
                              sliding_window=cfg.get("sliding_window") if sliding else None)

    def forward(self, positions, x):
        q, k, v = self.qkv_proj(x).split(self.sizes, dim=-1)
        return self.o_proj(self.attn(q, k, v, positions))


class _SharedMLP(nn.Module):
    def __init__(self, H, I, dtype, prefix) -> None:
        super().__init__()
        self.input_linear = MergedColumnParallelLinear(H, [I, I], False, dtype, None, f"{prefix}.input_linear")
        self.output_linear = RowParallelLinear(I, H, False, dtype=dtype, prefix=f"{prefix}.output_linear")

    def forward(self, x):
        a, b = self.input_linear(x).chunk(2, dim=-1)
        return self.output_linear(F.silu(a) * b)


class GraniteMoeLayer(nn.Module):
    def __init__(self, cfg, mc, dtype, quant_config, prefix, idx, enable_ep) -> None:
        super().__init__()
        H, eps = cfg["hidden_size"], cfg.get("rms_norm_eps", 1e-6)
        types = cfg.get("layer_types")
        self.is_mamba = bool(types) and types[idx] == "mamba"
        if self.is_mamba:
            assert get_tp_group().world_size == 1, "hybrid Mamba-2 layers run replicated (tensor_parallel_size=1)"
            nh = cfg["mamba_n_heads"]
            hd = cfg.get("mamba_d_head") if cfg.get("mamba_d_head") not in (None, "auto") else cfg.get("mamba_expand", 2) * H // nh
            self.mamba = Mamba2Mixer(H, nh, hd, cfg.get("mamba_d_state", 256), cfg.get("mamba_d_conv", 4), cfg.get("mamba_n_groups", 1),
                                     bool(cfg.get("mamba_proj_bias", False)), bool(cfg.get("mamba_conv_bias", True)),
                                     cfg.get("mamba_chunk_size", 256), eps, dtype, f"{prefix}.mamba")
        else:
            # Granite-SWA checkpoints mark some attention layers as sliding-window in ``layer_types``
            self.self_attn = _Attn(cfg, mc, dtype, f"{prefix}.self_attn", sliding=bool(types) and types[idx] == "sliding_attention")
        E = cfg.get("num_local_experts") or 0
        self.router = self.experts = None
        if E > 0:
            self.router = ReplicatedLinear(H, E, False, dtype, None, f"{prefix}.block_sparse_moe.router")
            self.experts = FusedMoE(E, cfg["num_experts_per_tok"], H, cfg["intermediate_size"], dtype, renormalize=True,
                                    act=cfg.get("hidden_act", "silu"), enable_ep=enable_ep, reduce_results=True,
                                    prefix=f"{prefix}.block_sparse_moe.experts")
        sh = cfg.get("shared_intermediate_size") or 0
        self.shared_mlp = _SharedMLP(H, sh, dtype, f"{prefix}.shared_mlp") if sh > 0 else None
        self.input_layernorm, self.post_attention_layernorm = RMSNorm(H, eps, dtype), RMSNorm(H, eps, dtype)
        self.rm = float(cfg.get("residual_multiplier", 1.0))

    def forward(self, positions, hidden, residual, next_norm: Optional[nn.Module]):
        x = residual + (self.mamba(hidden) if self.is_mamba else self.self_attn(positions, hidden)) * self.rm
        h = self.post_attention_layernorm(x)
        m = self.experts(h, self.router(h)) if self.experts is not None else 0
        if self.shared_mlp is not None:
            m = m + self.shared_mlp(h)
        x = x + m * self.rm
        return (next_norm(x) if next_norm is not None else x), x


class GraniteMoeModel(LlamaModel):
    def __init__(self, vllm_config: VllmConfig, prefix: str = "model", quant_config=None) -> None:
        self._enable_ep = bool(vllm_config.parallel_config.enable_expert_parallel)
        super().__init__(vllm_config, prefix, quant_config)
        self.embed_scale = float(self.cfg.get("embedding_multiplier", 1.0))

    def _make_layer(self, cfg, mc, dtype, quant_config, prefix, idx):
        return GraniteMoeLayer(cfg, mc, dtype, quant_config, prefix, idx, self._enable_ep)


class GraniteMoeForCausalLM(LlamaForCausalLM):
    model_cls = GraniteMoeModel

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        cfg = vllm_config.model_config.hf_config
        if cfg.get("model_type") in ("granitemoe", "granitemoeshared"):
            cfg.setdefault("position_embedding_type", "rope")
        super().__init__(vllm_config, prefix)
        if self.is_last:
            self.logits_processor = LogitsProcessor(cfg["vocab_size"], scale=1.0 / float(cfg.get("logits_scaling", 1.0)))

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        from vllm_b200.models.mamba2 import _load_mixer_tensor
        params = dict(self.named_parameters())
        modules = dict(self.named_modules())
        loaded: set[str] = set()
        rest = []
        for name, w in weights:
            n = self._localize(name)
            if n is None:
                continue
            if ".mamba." in n:
                got = _load_mixer_tensor(params, n, w)
                if got:
                    loaded.add(got)
            elif ".block_sparse_moe.input_linear." in n:                    # [E, 2I, H]: rows [gate | up] per expert
                tgt = n.replace(".block_sparse_moe.input_linear.weight", ".experts.w13_weight")
                mod, I = modules[tgt.rsplit(".", 1)[0]], w.shape[1] // 2
                for e in range(w.shape[0]):
                    mod.weight_loader(params[tgt], w[e, :I], "w1", e)
                    mod.weight_loader(params[tgt], w[e, I:], "w3", e)
                loaded.add(tgt)
            elif ".block_sparse_moe.output_linear." in n:
                tgt = n.replace(".block_sparse_moe.output_linear.weight", ".experts.w2_weight")
                mod = modules[tgt.rsplit(".", 1)[0]]
                for e in range(w.shape[0]):
                    mod.weight_loader(params[tgt], w[e], "w2", e)
                loaded.add(tgt)
            elif ".block_sparse_moe.router.layer." in n:
                tgt = n.replace(".block_sparse_moe.router.layer.", ".router.")
                params[tgt].data.copy_(w)
                loaded.add(tgt)
            elif ".shared_mlp.input_linear." in n:
                modules[n.rsplit(".", 1)[0]].weight_loader(params[n], w, None)
                loaded.add(n)
            else:
                rest.append((name, w))
        return loaded | super().load_weights(rest)


GraniteMoeSharedForCausalLM = GraniteMoeForCausalLM
GraniteMoeHybridForCausalLM = GraniteMoeForCausalLM
