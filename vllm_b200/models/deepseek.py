"""DeepSeek-V2 / V3 (MLA attention + group-limited MoE routing + shared experts).

Reference parity: vllm/model_executor/models/deepseek_v2.py (``DeepseekV2MLAAttention``,
``DeepseekV2MoE``, ``load_weights`` expert mapping) and the MLA algorithm note in
layers/attention/mla_attention.py:1-200.

Attention runs in the *absorbed* form for both prefill and decode: the cache holds only the 512-d
latent + 64-d rope key per token (576 values, ~1/57 of an MHA cache for V3), queries are projected
into the latent space with W_UK, attention is MQA over the latent, and W_UV maps the result back.
That keeps one KV layout and one kernel family for every phase and is what lets 180 GB hold very
long batches.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.nn as nn

from vllm_b200.config import VllmConfig
from vllm_b200.layers.attention import MLAAttentionCore
from vllm_b200.layers.common import RMSNorm
from vllm_b200.layers.linear import (ColumnParallelLinear, ReplicatedLinear, RowParallelLinear,
                                     tp_reduce_residual_norm)
from vllm_b200.layers.moe import FusedMoE
from vllm_b200.layers.rotary import get_rope, yarn_mscale
from vllm_b200.models.llama import LlamaForCausalLM, LlamaMLP, LlamaModel
from vllm_b200.parallel.state import get_tp_group
from vllm_b200.parallel.utils import divide


class DeepseekMLA(nn.Module):
    def __init__(self, cfg: dict, mc, dtype, quant_config, prefix: str) -> None:
        super().__init__()
        H = cfg["hidden_size"]
        self.n_heads = cfg["num_attention_heads"]
        tp = get_tp_group().world_size
        self.local_heads = divide(self.n_heads, tp)
        self.nope, self.rope_dim, self.v_dim = cfg["qk_nope_head_dim"], cfg["qk_rope_head_dim"], cfg["v_head_dim"]
        self.q_head = self.nope + self.rope_dim
        self.q_lora, self.kv_lora = cfg.get("q_lora_rank"), cfg["kv_lora_rank"]
        eps = cfg.get("rms_norm_eps", 1e-6)
        if self.q_lora:
            self.q_a_proj = ReplicatedLinear(H, self.q_lora, False, dtype, quant_config, f"{prefix}.q_a_proj")
            self.q_a_layernorm = RMSNorm(self.q_lora, eps, dtype)
            self.q_b_proj = ColumnParallelLinear(self.q_lora, self.n_heads * self.q_head, False, dtype=dtype,
                                                 quant_config=quant_config, prefix=f"{prefix}.q_b_proj")
        else:
            self.q_proj = ColumnParallelLinear(H, self.n_heads * self.q_head, False, dtype=dtype,
                                               quant_config=quant_config, prefix=f"{prefix}.q_proj")
        self.kv_a_proj_with_mqa = ReplicatedLinear(H, self.kv_lora + self.rope_dim, False, dtype, quant_config,
                                                   f"{prefix}.kv_a_proj_with_mqa")
        self.kv_a_layernorm = RMSNorm(self.kv_lora, eps, dtype)
        # kept unquantised: it is consumed as two batched absorbed matrices, not as a GEMM weight
        self.kv_b_proj = ColumnParallelLinear(self.kv_lora, self.n_heads * (self.nope + self.v_dim), False,
                                              dtype=dtype, quant_config=None, prefix=f"{prefix}.kv_b_proj")
        self.o_proj = RowParallelLinear(self.n_heads * self.v_dim, H, False, dtype=dtype, quant_config=quant_config,
                                        prefix=f"{prefix}.o_proj")
        rs = cfg.get("rope_scaling")
        max_pos = mc.max_model_len if mc.max_model_len > 0 else cfg.get("max_position_embeddings", 4096)
        # DeepSeek checkpoints rotate (even, odd) pairs; GLM-5's MLA rotates the two halves (``rope_interleave: false``)
        self.rope = get_rope(self.rope_dim, self.rope_dim, max_pos, cfg.get("rope_theta", 10000.0),
                             not bool(cfg.get("rope_interleave", True)), rs)
        scale = self.q_head ** -0.5
        if rs and rs.get("mscale_all_dim"):
            m = yarn_mscale(rs["factor"], rs["mscale_all_dim"])
            scale = scale * m * m
        self.attn = MLAAttentionCore(self.local_heads, self.kv_lora, self.rope_dim, scale, prefix=f"{prefix}.attn")
        self._w_uk: Optional[torch.Tensor] = None
        self._w_uv: Optional[torch.Tensor] = None

    def _absorbed(self):
        if self._w_uk is None or self._w_uk.device != self.kv_b_proj.weight.device:
            w = self.kv_b_proj.weight.view(self.local_heads, self.nope + self.v_dim, self.kv_lora)
            self._w_uk = w[:, :self.nope].contiguous()                       # [h, nope, R]
            self._w_uv = w[:, self.nope:].transpose(1, 2).contiguous()       # [h, R, v]
        return self._w_uk, self._w_uv

    def process_weights_after_loading(self) -> None:
        self._w_uk = self._w_uv = None

    def forward(self, positions: torch.Tensor, hidden: torch.Tensor) -> torch.Tensor:
        T = hidden.shape[0]
        if self.q_lora:
            q = self.q_b_proj(self.q_a_layernorm(self.q_a_proj(hidden)))
        else:
            q = self.q_proj(hidden)
        q = q.view(T, self.local_heads, self.q_head)
        q_nope, q_pe = q[..., :self.nope], q[..., self.nope:]
        lat = self.kv_a_proj_with_mqa(hidden)
        kv_c = self.kv_a_layernorm(lat[:, :self.kv_lora].contiguous())
        k_pe = lat[:, self.kv_lora:].contiguous()
        q_pe, k_pe = self.rope(positions, q_pe.reshape(T, -1).contiguous(), k_pe)
        q_pe = q_pe.view(T, self.local_heads, self.rope_dim)
        w_uk, w_uv = self._absorbed()
        q_abs = torch.einsum("thn,hnr->thr", q_nope, w_uk)
        o_lat = self.attn(q_abs, q_pe, kv_c, k_pe)                           # [T, h, R]
        o = torch.einsum("thr,hrv->thv", o_lat, w_uv).reshape(T, self.local_heads * self.v_dim)
        return o


class DeepseekMoE(nn.Module):
    def __init__(self, cfg: dict, dtype, quant_config, prefix: str, enable_ep: bool) -> None:
        super().__init__()
        H = cfg["hidden_size"]
        E = cfg["n_routed_experts"]
        self.gate = ReplicatedLinear(H, E, False, dtype, None, f"{prefix}.gate")
        noaux = cfg.get("topk_method") == "noaux_tc"
        self.experts = FusedMoE(E, cfg["num_experts_per_tok"], H, cfg["moe_intermediate_size"], dtype,
                                renormalize=bool(cfg.get("norm_topk_prob", False)),
                                use_grouped_topk=True, n_group=cfg.get("n_group") or 1,
                                topk_group=cfg.get("topk_group") or 1,
                                scoring_func=cfg.get("scoring_func", "softmax"),
                                routed_scaling_factor=cfg.get("routed_scaling_factor", 1.0), e_score_bias=noaux,
                                act=cfg.get("hidden_act", "silu"), enable_ep=enable_ep, reduce_results=False,
                                prefix=f"{prefix}.experts", quant_config=quant_config)
        ns = cfg.get("n_shared_experts") or 0
        self.shared_experts = None
        if ns:
            self.shared_experts = LlamaMLP(H, cfg["moe_intermediate_size"] * ns, cfg.get("hidden_act", "silu"),
                                           False, dtype, quant_config, f"{prefix}.shared_experts")
            self.shared_experts.down_proj.reduce_results = False

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        out = self.experts(x, self.gate(x).float())
        if self.shared_experts is not None:
            out = out + self.shared_experts(x)
        return out          # per-rank partial sum


class DeepseekDecoderLayer(nn.Module):
    def __init__(self, cfg, mc, dtype, quant_config, prefix: str, layer_idx: int, enable_ep: bool) -> None:
        super().__init__()
        eps = cfg.get("rms_norm_eps", 1e-6)
        self.self_attn = DeepseekMLA(cfg, mc, dtype, quant_config, f"{prefix}.self_attn")
        self.input_layernorm = RMSNorm(cfg["hidden_size"], eps, dtype)
        self.post_attention_layernorm = RMSNorm(cfg["hidden_size"], eps, dtype)
        sparse = (cfg.get("n_routed_experts") and layer_idx >= cfg.get("first_k_dense_replace", 0)
                  and layer_idx % cfg.get("moe_layer_freq", 1) == 0)
        if sparse:
            self.mlp = DeepseekMoE(cfg, dtype, quant_config, f"{prefix}.mlp", enable_ep)
        else:
            self.mlp = LlamaMLP(cfg["hidden_size"], cfg["intermediate_size"], cfg.get("hidden_act", "silu"), False,
                                dtype, quant_config, f"{prefix}.mlp")
        self.sparse = bool(sparse)

    def forward(self, positions, hidden, residual, next_norm: Optional[nn.Module]):
        a = self.self_attn(positions, hidden)
        hidden, residual = self.self_attn.o_proj.forward_reduce_norm(a, residual, self.post_attention_layernorm)
        if not self.sparse:
            return self.mlp.forward_reduce_norm(hidden, residual, next_norm)
        partial = self.mlp(hidden)
        if next_norm is None:
            tp = get_tp_group()
            return (tp.all_reduce(partial) if tp.world_size > 1 else partial), residual
        return tp_reduce_residual_norm(partial, residual, next_norm)


class DeepseekModel(LlamaModel):
    def __init__(self, vllm_config: VllmConfig, prefix: str = "model", quant_config=None) -> None:
        self._enable_ep = bool(vllm_config.parallel_config.enable_expert_parallel)
        super().__init__(vllm_config, prefix, quant_config)

    def _make_layer(self, cfg, mc, dtype, quant_config, prefix, idx):
        return DeepseekDecoderLayer(cfg, mc, dtype, quant_config, prefix, idx, self._enable_ep)


class DeepseekV3ForCausalLM(LlamaForCausalLM):
    model_cls = DeepseekModel
    stacked = [("gate_up_proj", "gate_proj", 0), ("gate_up_proj", "up_proj", 1)]
    expert_names = {"gate_proj": ("w13_weight", "w1"), "up_proj": ("w13_weight", "w3"),
                    "down_proj": ("w2_weight", "w2")}

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        params = dict(self.named_parameters())
        modules = dict(self.named_modules())
        dense: list[tuple[str, torch.Tensor]] = []
        loaded: set[str] = set()
        marker = ".mlp.experts."
        for name, w in self._expand_fused(weights):
            if ".layers." in name:
                # multi-token-prediction heads (layer index >= num_hidden_layers) are not part of the base model
                li = int(name.split(".layers.")[1].split(".")[0])
                if li >= self.cfg["num_hidden_layers"]:
                    continue
            if name.endswith(".mlp.gate.e_score_correction_bias"):
                name = self._localize(name)
                if name is None:
                    continue
                name = name.replace(".mlp.gate.e_score_correction_bias", ".mlp.experts.e_score_correction_bias")
                if name in params:
                    params[name].data.copy_(w)
                    loaded.add(name)
                continue
            if marker not in name:
                dense.append((name, w))
                continue
            lname = self._localize(name)
            if lname is None:
                continue
            head, tail = lname.split(marker, 1)
            parts = tail.split(".")
            if not parts[0].isdigit() or parts[1] not in self.expert_names:
                continue
            suffix, shard = self.expert_names[parts[1]]
            if parts[-1] == "weight_scale_inv":                  # block scales of an fp8 expert projection
                pname = f"{head}{marker}{suffix.replace('_weight', '_scale')}"
                if pname in params:
                    modules[pname.rsplit(".", 1)[0]].scale_loader(params[pname], w, shard, int(parts[0]))
                    loaded.add(pname)
                continue
            pname = f"{head}{marker}{suffix}"
            if pname in params:
                modules[pname.rsplit(".", 1)[0]].weight_loader(params[pname], w, shard, int(parts[0]))
                loaded.add(pname)
        loaded |= super().load_weights(dense)
        for m in self.modules():
            if isinstance(m, DeepseekMLA):
                m.process_weights_after_loading()
        return loaded

    @staticmethod
    def _expand_fused(weights):
        for name, w in weights:
            if name.endswith(".experts.gate_up_proj"):
                base = name[: -len("gate_up_proj")]
                I = w.shape[1] // 2
                for e in range(w.shape[0]):
                    yield f"{base}{e}.gate_proj.weight", w[e, :I]
                    yield f"{base}{e}.up_proj.weight", w[e, I:]
            elif name.endswith(".experts.down_proj"):
                base = name[: -len("down_proj")]
                for e in range(w.shape[0]):
                    yield f"{base}{e}.down_proj.weight", w[e]
            else:
                yield name, w


# ------------------------------------------------------------------------------------------------ GLM-4.5 / 4.6 (MoE)
class _MiniCPM3Layer(nn.Module):
    """MLA attention + dense gated MLP with the muP residual scale ``scale_depth / sqrt(num_layers)`` on both branches
    (explicit residual stream: the fused add-norm path has no place for the multiplier)."""

    def __init__(self, cfg, mc, dtype, quant_config, prefix: str, layer_idx: int, enable_ep: bool) -> None:
        super().__init__()
        import math
        eps = cfg.get("rms_norm_eps", 1e-6)
        self.self_attn = DeepseekMLA(cfg, mc, dtype, quant_config, f"{prefix}.self_attn")
        self.input_layernorm = RMSNorm(cfg["hidden_size"], eps, dtype)
        self.post_attention_layernorm = RMSNorm(cfg["hidden_size"], eps, dtype)
        self.mlp = LlamaMLP(cfg["hidden_size"], cfg["intermediate_size"], cfg.get("hidden_act", "silu"), False, dtype, quant_config,
                            f"{prefix}.mlp")
        self.rm = float(cfg.get("scale_depth", 1.0)) / math.sqrt(cfg["num_hidden_layers"])

    def forward(self, positions, hidden, residual, next_norm: Optional[nn.Module]):
        x = residual + self.self_attn.o_proj(self.self_attn(positions, hidden)) * self.rm
        x = x + self.mlp(self.post_attention_layernorm(x)) * self.rm
        return (next_norm(x) if next_norm is not None else x), x


class _MiniCPM3Model(DeepseekModel):
    def __init__(self, vllm_config: VllmConfig, prefix: str = "model", quant_config=None) -> None:
        super().__init__(vllm_config, prefix, quant_config)
        self.embed_scale = float(self.cfg.get("scale_emb", 1.0))

    def _make_layer(self, cfg, mc, dtype, quant_config, prefix, idx):
        return _MiniCPM3Layer(cfg, mc, dtype, quant_config, prefix, idx, self._enable_ep)


class MiniCPM3ForCausalLM(DeepseekV3ForCausalLM):
    """MiniCPM3-4B: multi-head latent attention (rotate-half rotary, LongRoPE) in a dense decoder with the MiniCPM muP
    scalars — scaled embeddings, residual branches × ``scale_depth / sqrt(L)``, logits ÷ ``hidden / dim_model_base``
    (reference: models/minicpm3.py)."""
    model_cls = _MiniCPM3Model

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        cfg = vllm_config.model_config.hf_config
        cfg.setdefault("rope_interleave", False)
        cfg.pop("n_routed_experts", None)
        super().__init__(vllm_config, prefix)
        if self.is_last:
            from vllm_b200.layers.common import LogitsProcessor
            self.logits_processor = LogitsProcessor(cfg["vocab_size"], scale=float(cfg.get("dim_model_base", cfg["hidden_size"])) / cfg["hidden_size"])


class Glm4MoeLiteForCausalLM(DeepseekV3ForCausalLM):
    """GLM-4.7-Flash ("MoE lite"): the DeepSeek-V3 block — MLA with interleaved rotary, sigmoid router with selection bias,
    shared expert, leading dense layers given by ``mlp_layer_types`` (reference: models/glm4_moe_lite.py)."""

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        cfg = vllm_config.model_config.hf_config
        cfg.setdefault("topk_method", "noaux_tc")
        cfg.setdefault("scoring_func", "sigmoid")
        kinds = cfg.get("mlp_layer_types")
        if kinds and "first_k_dense_replace" not in cfg:
            cfg["first_k_dense_replace"] = next((i for i, k in enumerate(kinds) if k == "sparse"), len(kinds))
        cfg.setdefault("first_k_dense_replace", 1)
        rp = cfg.get("rope_parameters") or {}
        cfg.setdefault("rope_theta", rp.get("rope_theta", 10000.0))
        super().__init__(vllm_config, prefix)


class GlmMoeDsaForCausalLM(Glm4MoeLiteForCausalLM):
    """GLM-5 / DeepSeek-V3.2 family: the MLA + MoE block with a *lightning indexer* that restricts every query to its
    ``index_topk`` (2048) highest-scoring cached tokens.  PARTIAL: the indexer is not run — attention is dense over the whole
    context, which is exactly the sparse model's output while the context is at most ``index_topk`` tokens (the indexer then
    selects everything) and a superset of its attention pattern beyond; the indexer tensors of the checkpoint are skipped.
    Reference: models/deepseek_v2.py (``DeepseekV32ForCausalLM``, ``Indexer``), v1/attention/backends/mla/indexer.py."""

    rope_interleave_default = False

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        vllm_config.model_config.hf_config.setdefault("rope_interleave", self.rope_interleave_default)
        super().__init__(vllm_config, prefix)
        cfg = vllm_config.model_config
        topk = int(cfg.hf_config.get("index_topk") or 2048)
        if cfg.max_model_len > topk:
            from vllm_b200.logger import init_logger
            init_logger(__name__).warning(
                "sparse-attention indexer not applied: contexts beyond index_topk=%d tokens are attended densely "
                "(max_model_len=%d)", topk, cfg.max_model_len)

    def load_weights(self, weights):
        return super().load_weights((n, w) for n, w in weights if ".indexer." not in n)


class DeepseekV32ForCausalLM(GlmMoeDsaForCausalLM):
    """DeepSeek-V3.2-Exp: same partial support; its MLA keeps DeepSeek's interleaved rotary."""
    rope_interleave_default = True


class Glm4MoeDecoderLayer(DeepseekDecoderLayer):
    """DeepSeek-V3 style MoE block (sigmoid scores + correction bias, grouped top-k, shared experts, leading dense
    layers) around ordinary GQA attention with partial rotary and optional per-head q/k norms
    (reference: models/glm4_moe.py)."""

    def __init__(self, cfg, mc, dtype, quant_config, prefix: str, layer_idx: int, enable_ep: bool) -> None:
        nn.Module.__init__(self)
        from vllm_b200.models.llama import LlamaAttention
        eps = cfg.get("rms_norm_eps", 1e-5)
        self.self_attn = LlamaAttention(cfg, mc, dtype, quant_config, f"{prefix}.self_attn", layer_idx)
        self.input_layernorm = RMSNorm(cfg["hidden_size"], eps, dtype)
        self.post_attention_layernorm = RMSNorm(cfg["hidden_size"], eps, dtype)
        sparse = bool(cfg.get("n_routed_experts")) and layer_idx >= cfg.get("first_k_dense_replace", 0)
        if sparse:
            self.mlp = DeepseekMoE(cfg, dtype, quant_config, f"{prefix}.mlp", enable_ep)
        else:
            self.mlp = LlamaMLP(cfg["hidden_size"], cfg["intermediate_size"], cfg.get("hidden_act", "silu"), False,
                                dtype, quant_config, f"{prefix}.mlp")
        self.sparse = sparse


class Glm4MoeModel(DeepseekModel):
    def _make_layer(self, cfg, mc, dtype, quant_config, prefix, idx):
        return Glm4MoeDecoderLayer(cfg, mc, dtype, quant_config, prefix, idx, self._enable_ep)


class Glm4MoeForCausalLM(DeepseekV3ForCausalLM):
    model_cls = Glm4MoeModel
    stacked = [("qkv_proj", "q_proj", "q"), ("qkv_proj", "k_proj", "k"), ("qkv_proj", "v_proj", "v"),
               ("gate_up_proj", "gate_proj", 0), ("gate_up_proj", "up_proj", 1)]

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        cfg = vllm_config.model_config.hf_config
        cfg.setdefault("topk_method", "noaux_tc")
        cfg.setdefault("scoring_func", "sigmoid")
        cfg["qk_norm"] = bool(cfg.get("use_qk_norm", False))
        rp = cfg.get("rope_parameters") or {}
        cfg.setdefault("partial_rotary_factor", rp.get("partial_rotary_factor", 0.5))
        cfg.setdefault("rope_theta", rp.get("rope_theta", 10000.0))
        super().__init__(vllm_config, prefix)


class Ernie4_5_MoeForCausalLM(Glm4MoeForCausalLM):
    """ERNIE-4.5 MoE (text): GQA attention with interleaved rotary + softmax router whose top-k SELECTION adds a learned
    correction bias while the mixing weights stay the unbiased, renormalised probabilities; shared experts; leading dense
    layers (reference: models/ernie45_moe.py)."""

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        cfg = vllm_config.model_config.hf_config
        cfg.setdefault("n_routed_experts", cfg.get("moe_num_experts"))
        cfg.setdefault("num_experts_per_tok", cfg.get("moe_k", 2))
        cfg.setdefault("n_shared_experts", cfg.get("moe_num_shared_experts", 0))
        cfg.setdefault("first_k_dense_replace", cfg.get("moe_layer_start_index", 0))
        cfg.setdefault("scoring_func", "softmax")
        cfg.setdefault("norm_topk_prob", True)
        cfg.setdefault("partial_rotary_factor", 1.0)
        cfg.setdefault("rope_interleaved", True)
        cfg.setdefault("attention_bias", bool(cfg.get("use_bias", False)))
        cfg.setdefault("tie_word_embeddings", True)
        super().__init__(vllm_config, prefix)

    def load_weights(self, weights):
        def mapped():
            for name, w in weights:
                if name.endswith(".mlp.gate.moe_statics.e_score_correction_bias"):
                    yield name.replace(".moe_statics.", "."), w.reshape(-1)
                else:
                    yield name, w
        return super().load_weights(mapped())


class DeepseekForCausalLM(Glm4MoeForCausalLM):
    """DeepSeek-MoE (v1, 16B): Llama attention with full rotary, softmax router without bias or groups, shared experts,
    leading dense layer (reference: models/deepseek_v2.py ``DeepseekForCausalLM``)."""

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        cfg = vllm_config.model_config.hf_config
        cfg.setdefault("topk_method", "greedy")
        cfg.setdefault("scoring_func", "softmax")
        cfg.setdefault("partial_rotary_factor", 1.0)
        cfg.setdefault("norm_topk_prob", False)
        cfg.setdefault("first_k_dense_replace", 1)
        super().__init__(vllm_config, prefix)


class SolarOpenForCausalLM(Glm4MoeForCausalLM):
    """Solar-Open: the GLM-4.5 MoE composition with every layer sparse, no q / k norms and full-width rotary by default
    (reference: models/solar_open.py)."""

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        cfg = vllm_config.model_config.hf_config
        cfg.setdefault("first_k_dense_replace", 0)
        cfg.setdefault("intermediate_size", cfg["moe_intermediate_size"])
        cfg.setdefault("partial_rotary_factor", (cfg.get("rope_parameters") or {}).get("partial_rotary_factor", 1.0))
        super().__init__(vllm_config, prefix)


class Dots1ForCausalLM(Glm4MoeForCausalLM):
    """dots.llm1: the same composition (GQA attention with per-head q/k norms + DeepSeek-V3 style MoE), full rotary
    (reference: models/dots1.py)."""

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        cfg = vllm_config.model_config.hf_config
        cfg.setdefault("use_qk_norm", True)
        cfg.setdefault("partial_rotary_factor", 1.0)
        super().__init__(vllm_config, prefix)
