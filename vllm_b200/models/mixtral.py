"""Sparse-MoE decoders that reuse the Llama attention block: Mixtral, Qwen2-MoE, Qwen3-MoE.

Reference parity: vllm/model_executor/models/mixtral.py (``MixtralMoE`` gate + ``FusedMoE``,
``load_weights`` expert mapping), qwen2_moe.py (shared expert + sigmoid shared gate, ``mlp_only_layers``,
``decoder_sparse_step``), qwen3_moe.py (``norm_topk_prob``).
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.nn as nn

from vllm_b200.config import VllmConfig
from vllm_b200.layers.common import RMSNorm
from vllm_b200.layers.linear import ReplicatedLinear, tp_reduce_residual_norm
from vllm_b200.layers.moe import FusedMoE
from vllm_b200.models.llama import LlamaAttention, LlamaForCausalLM, LlamaMLP, LlamaModel
from vllm_b200.parallel.state import get_tp_group


class SparseMoeBlock(nn.Module):
    """router (replicated GEMM) → FusedMoE (+ optional always-on shared expert).  Returns the per-rank PARTIAL
    sum; the decoder layer reduces it together with the residual add + norm."""

    def __init__(self, cfg: dict, dtype, quant_config, prefix: str, enable_ep: bool, *, gate_name: str = "gate",
                 n_experts: int, top_k: int, inter: int, renormalize: bool, shared_inter: int = 0,
                 shared_gate: bool = False, moe_kw: Optional[dict] = None) -> None:
        super().__init__()
        hidden = cfg["hidden_size"]
        self.gate = ReplicatedLinear(hidden, n_experts, False, dtype, None, f"{prefix}.{gate_name}")
        # moe_kw: router flavour (sigmoid scores, selection bias, grouped top-k, routed scaling) — FusedMoE keywords
        self.experts = FusedMoE(n_experts, top_k, hidden, inter, dtype, renormalize=renormalize,
                                act=cfg.get("hidden_act", "silu"), enable_ep=enable_ep, reduce_results=False,
                                prefix=f"{prefix}.experts", **(moe_kw or {}))
        self.shared_expert = None
        self.shared_expert_gate = None
        if shared_inter > 0:
            self.shared_expert = LlamaMLP(hidden, shared_inter, cfg.get("hidden_act", "silu"), False, dtype,
                                          quant_config, f"{prefix}.shared_expert")
            self.shared_expert.down_proj.reduce_results = False
            if shared_gate:
                self.shared_expert_gate = ReplicatedLinear(hidden, 1, False, dtype, None,
                                                           f"{prefix}.shared_expert_gate")

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        out = self.experts(x, self.gate(x))
        if self.shared_expert is not None:
            s = self.shared_expert(x)
            if self.shared_expert_gate is not None:
                s = torch.sigmoid(self.shared_expert_gate(x)) * s
            out = out + s
        return out


class MoeDecoderLayer(nn.Module):
    def __init__(self, cfg, mc, dtype, quant_config, prefix: str, layer_idx: int, enable_ep: bool,
                 flavour: str) -> None:
        super().__init__()
        eps = cfg.get("rms_norm_eps", 1e-6)
        self.self_attn = LlamaAttention(cfg, mc, dtype, quant_config, f"{prefix}.self_attn", layer_idx)
        self.input_layernorm = RMSNorm(cfg["hidden_size"], eps, dtype)
        self.post_attention_layernorm = RMSNorm(cfg["hidden_size"], eps, dtype)
        if flavour == "mixtral":
            self.block_sparse_moe = SparseMoeBlock(
                cfg, dtype, quant_config, f"{prefix}.block_sparse_moe", enable_ep, n_experts=cfg["num_local_experts"],
                top_k=cfg["num_experts_per_tok"], inter=cfg["intermediate_size"], renormalize=True)
            self.moe, self.dense = self.block_sparse_moe, None
        else:
            sparse = (layer_idx not in cfg.get("mlp_only_layers", [])
                      and cfg.get("num_experts", 0) > 0
                      and (layer_idx + 1) % cfg.get("decoder_sparse_step", 1) == 0)
            if sparse:
                self.mlp = SparseMoeBlock(
                    cfg, dtype, quant_config, f"{prefix}.mlp", enable_ep, n_experts=cfg["num_experts"],
                    top_k=cfg["num_experts_per_tok"], inter=cfg["moe_intermediate_size"],
                    renormalize=bool(cfg.get("norm_topk_prob", flavour == "qwen3")),
                    shared_inter=cfg.get("shared_expert_intermediate_size", 0) if flavour == "qwen2" else 0,
                    shared_gate=flavour == "qwen2")
                self.moe, self.dense = self.mlp, None
            else:
                self.mlp = LlamaMLP(cfg["hidden_size"], cfg["intermediate_size"], cfg.get("hidden_act", "silu"),
                                    False, dtype, quant_config, f"{prefix}.mlp")
                self.moe, self.dense = None, self.mlp

    def forward(self, positions, hidden, residual, next_norm: Optional[nn.Module]):
        a = self.self_attn(positions, hidden)
        hidden, residual = self.self_attn.o_proj.forward_reduce_norm(a, residual, self.post_attention_layernorm)
        if self.dense is not None:
            return self.dense.forward_reduce_norm(hidden, residual, next_norm)
        if next_norm is not None and self.moe.shared_expert is None:
            fused = self.moe.experts.forward_reduce_norm(hidden, self.moe.gate(hidden), residual, next_norm)
            if fused is not None:
                return fused
        partial = self.moe(hidden)
        if next_norm is None:
            tp = get_tp_group()
            return (tp.all_reduce(partial) if tp.world_size > 1 else partial), residual
        return tp_reduce_residual_norm(partial, residual, next_norm)


class MoeModel(LlamaModel):
    flavour = "mixtral"

    def __init__(self, vllm_config: VllmConfig, prefix: str = "model", quant_config=None) -> None:
        self._enable_ep = bool(vllm_config.parallel_config.enable_expert_parallel)
        super().__init__(vllm_config, prefix, quant_config)

    def _make_layer(self, cfg, mc, dtype, quant_config, prefix, idx):
        return MoeDecoderLayer(cfg, mc, dtype, quant_config, prefix, idx, self._enable_ep, self.flavour)


class MixtralModel(MoeModel):
    flavour = "mixtral"


class Qwen2MoeModel(MoeModel):
    flavour = "qwen2"


class Qwen3MoeModel(MoeModel):
    flavour = "qwen3"


class MixtralForCausalLM(LlamaForCausalLM):
    model_cls = MixtralModel
    # checkpoint expert tensor names → (fused param suffix, shard id)
    expert_names = {"w1": ("w13_weight", "w1"), "w3": ("w13_weight", "w3"), "w2": ("w2_weight", "w2")}
    expert_container = "block_sparse_moe.experts"

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        params = dict(self.named_parameters())
        modules = dict(self.named_modules())
        dense: list[tuple[str, torch.Tensor]] = []
        loaded: set[str] = set()
        marker = f".{self.expert_container}."
        for name, w in self._expand_fused(weights):
            if marker not in name:
                dense.append((name, w))
                continue
            # <layer>.<container>.<expert id>.<w1|w2|w3 / gate_proj|...>.weight
            lname = self._localize(name)
            if lname is None:
                continue
            head, tail = lname.split(marker, 1)
            parts = tail.split(".")
            if not parts[0].isdigit():
                dense.append((name, w))
                continue
            eid, wname = int(parts[0]), parts[1]
            if wname not in self.expert_names:
                continue
            suffix, shard = self.expert_names[wname]
            pname = f"{head}{marker}{suffix}"
            if pname not in params:
                continue
            modules[pname.rsplit(".", 1)[0]].weight_loader(params[pname], w, shard, eid)
            loaded.add(pname)
        return loaded | super().load_weights(dense)


    def _expand_fused(self, weights):
        """Accept both checkpoint layouts: per-expert ``experts.<e>.<w1|w2|w3>.weight`` (hub checkpoints) and
        the fused ``experts.gate_up_proj [E,2I,H]`` / ``experts.down_proj [E,H,I]`` tensors newer HF
        releases keep in memory (there the Mixtral block is also called ``mlp``)."""
        container = self.expert_container
        block = container.split(".")[0]
        names = list(self.expert_names)          # [gate, up, down] in insertion order: w1, w3, w2
        gate_n, up_n, down_n = names[0], names[1], names[2]
        for name, w in weights:
            if block != "mlp" and ".mlp." in name and ".layers." in name:
                name = name.replace(".mlp.", f".{block}.")
            if name.endswith(".experts.gate_up_proj"):
                base = name[: -len("gate_up_proj")]
                I = w.shape[1] // 2
                for e in range(w.shape[0]):
                    yield f"{base}{e}.{gate_n}.weight", w[e, :I]
                    yield f"{base}{e}.{up_n}.weight", w[e, I:]
            elif name.endswith(".experts.down_proj"):
                base = name[: -len("down_proj")]
                for e in range(w.shape[0]):
                    yield f"{base}{e}.{down_n}.weight", w[e]
            else:
                yield name, w


class Qwen2MoeForCausalLM(MixtralForCausalLM):
    model_cls = Qwen2MoeModel
    expert_names = {"gate_proj": ("w13_weight", "w1"), "up_proj": ("w13_weight", "w3"),
                    "down_proj": ("w2_weight", "w2")}
    expert_container = "mlp.experts"

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        vllm_config.model_config.hf_config.setdefault("attention_bias", True)
        super().__init__(vllm_config, prefix)


class Qwen3MoeForCausalLM(MixtralForCausalLM):
    model_cls = Qwen3MoeModel
    expert_names = Qwen2MoeForCausalLM.expert_names
    expert_container = "mlp.experts"

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        vllm_config.model_config.hf_config["qk_norm"] = True
        super().__init__(vllm_config, prefix)


class OlmoeForCausalLM(MixtralForCausalLM):
    """OLMoE: Qwen-MoE style sparse block in every layer (softmax over all experts, top-k without renormalisation)
    behind attention with whole-projection q/k RMSNorm (reference: models/olmoe.py)."""
    model_cls = Qwen3MoeModel
    expert_names = Qwen2MoeForCausalLM.expert_names
    expert_container = "mlp.experts"

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        cfg = vllm_config.model_config.hf_config
        cfg["qk_norm"] = "full"
        cfg.setdefault("moe_intermediate_size", cfg["intermediate_size"])
        cfg.setdefault("norm_topk_prob", False)
        cfg.setdefault("decoder_sparse_step", 1)
        cfg.setdefault("mlp_only_layers", [])
        super().__init__(vllm_config, prefix)
