"""Mamba-2 language model and the Bamba hybrid (Mamba-2 + attention layers, gated MLP after every mixer).

Reference parity: vllm/model_executor/models/mamba2.py, bamba.py.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.nn as nn

from vllm_b200.config import VllmConfig
from vllm_b200.layers.common import LogitsProcessor, ParallelLMHead, RMSNorm, VocabParallelEmbedding
from vllm_b200.layers.mamba import Mamba2Mixer
from vllm_b200.models.llama import LlamaAttention, LlamaMLP
from vllm_b200.parallel.state import get_tp_group


def _load_mixer_tensor(params: dict, n: str, w: torch.Tensor) -> Optional[str]:
    n = (n.replace("mixer.conv1d.weight", "mixer.conv_weight").replace("mixer.conv1d.bias", "mixer.conv_bias")
          .replace("mixer.norm.weight", "mixer.norm_weight").replace("mamba.conv1d.weight", "mamba.conv_weight")
          .replace("mamba.conv1d.bias", "mamba.conv_bias").replace("mamba.norm.weight", "mamba.norm_weight"))
    if n not in params:
        return None
    p = params[n]
    p.data.copy_(w.reshape(p.shape) if n.endswith("conv_weight") else w)
    return n


class Mamba2Block(nn.Module):
    def __init__(self, cfg: dict, dtype, prefix: str) -> None:
        super().__init__()
        H = cfg["hidden_size"]
        eps = cfg.get("layer_norm_epsilon", 1e-5)
        self.norm = RMSNorm(H, eps, dtype)
        lim = cfg.get("time_step_limit") or (0.0, float("inf"))
        self.mixer = Mamba2Mixer(H, cfg["num_heads"], cfg["head_dim"], cfg.get("state_size", 128), cfg.get("conv_kernel", 4),
                                 cfg.get("n_groups", 1), bool(cfg.get("use_bias", False)), bool(cfg.get("use_conv_bias", True)),
                                 cfg.get("chunk_size", 256), eps, dtype, f"{prefix}.mixer",
                                 time_step_limit=(float(lim[0]), float(lim[1])))

    def forward(self, x):
        return x + self.mixer(self.norm(x))


class Mamba2ForCausalLM(nn.Module):
    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__()
        assert get_tp_group().world_size == 1, "Mamba-2 runs replicated (tensor_parallel_size=1)"
        mc = vllm_config.model_config
        cfg, dtype = mc.hf_config, mc.torch_dtype
        self.cfg = cfg
        H = cfg["hidden_size"]
        self.embeddings = VocabParallelEmbedding(cfg["vocab_size"], H, dtype)
        self.layers = nn.ModuleList([Mamba2Block(cfg, dtype, f"{prefix}backbone.layers.{i}")
                                     for i in range(cfg["num_hidden_layers"])])
        self.norm_f = RMSNorm(H, cfg.get("layer_norm_epsilon", 1e-5), dtype)
        self.lm_head = ParallelLMHead(cfg["vocab_size"], H, dtype=dtype)
        if cfg.get("tie_word_embeddings", False):
            self.lm_head.tie_weights(self.embeddings)
        self.logits_processor = LogitsProcessor(cfg["vocab_size"])

    def embed_input_ids(self, ids):
        return self.embeddings(ids)

    def forward(self, input_ids, positions, intermediate_tensors=None, inputs_embeds=None):
        x = inputs_embeds if inputs_embeds is not None else self.embeddings(input_ids)
        for layer in self.layers:
            x = layer(x)
        return self.norm_f(x)

    def compute_logits(self, hidden):
        return self.logits_processor(self.lm_head, hidden)

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        params = dict(self.named_parameters())
        loaded: set[str] = set()
        for name, w in weights:
            n = name[len("backbone."):] if name.startswith("backbone.") else name
            if n == "lm_head.weight" and self.cfg.get("tie_word_embeddings", False):
                continue
            got = _load_mixer_tensor(params, n, w)
            if got:
                loaded.add(got)
        return loaded


class BambaLayer(nn.Module):
    def __init__(self, cfg: dict, mc, dtype, prefix: str, idx: int) -> None:
        super().__init__()
        H, eps = cfg["hidden_size"], cfg.get("rms_norm_eps", 1e-5)
        self.is_attn = idx in set(cfg.get("attn_layer_indices") or [])
        if self.is_attn:
            acfg = dict(cfg)
            if cfg.get("partial_rotary_factor"):
                acfg["partial_rotary_factor"] = cfg["partial_rotary_factor"]
            self.self_attn = LlamaAttention(acfg, mc, dtype, None, f"{prefix}.self_attn", idx)
        else:
            nh = cfg["mamba_n_heads"]
            hd = cfg.get("mamba_d_head") if cfg.get("mamba_d_head") not in (None, "auto") else cfg.get("mamba_expand", 2) * H // nh
            self.mamba = Mamba2Mixer(H, nh, hd, cfg.get("mamba_d_state", 256), cfg.get("mamba_d_conv", 4),
                                     cfg.get("mamba_n_groups", 1), bool(cfg.get("mamba_proj_bias", False)),
                                     bool(cfg.get("mamba_conv_bias", True)), cfg.get("mamba_chunk_size", 256), eps, dtype,
                                     f"{prefix}.mamba")
        self.feed_forward = LlamaMLP(H, cfg["intermediate_size"], cfg.get("hidden_act", "silu"),
                                     bool(cfg.get("mlp_bias", False)), dtype, None, f"{prefix}.feed_forward")
        self.input_layernorm = RMSNorm(H, eps, dtype)
        self.pre_ff_layernorm = RMSNorm(H, eps, dtype)

    def forward(self, positions, x):
        h = self.input_layernorm(x)
        # (LlamaAttention returns the pre-projection context; its o_proj is applied by the caller)
        x = x + (self.self_attn.o_proj(self.self_attn(positions, h)) if self.is_attn else self.mamba(h))
        return x + self.feed_forward(self.pre_ff_layernorm(x))


class BambaForCausalLM(nn.Module):
    layer_cls = BambaLayer
    stacked = [("qkv_proj", "q_proj", "q"), ("qkv_proj", "k_proj", "k"), ("qkv_proj", "v_proj", "v"),
               ("gate_up_proj", "gate_proj", 0), ("gate_up_proj", "up_proj", 1)]

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__()
        assert get_tp_group().world_size == 1, "hybrid Mamba-2 models run replicated (tensor_parallel_size=1)"
        mc = vllm_config.model_config
        cfg, dtype = mc.hf_config, mc.torch_dtype
        self.cfg = cfg
        H = cfg["hidden_size"]
        self.embed_tokens = VocabParallelEmbedding(cfg["vocab_size"], H, dtype)
        self.layers = nn.ModuleList([self.layer_cls(cfg, mc, dtype, f"{prefix}model.layers.{i}", i)
                                     for i in range(cfg["num_hidden_layers"])])
        self.final_layernorm = RMSNorm(H, cfg.get("rms_norm_eps", 1e-5), dtype)
        self.lm_head = ParallelLMHead(cfg["vocab_size"], H, dtype=dtype)
        if cfg.get("tie_word_embeddings", False):
            self.lm_head.tie_weights(self.embed_tokens)
        self.logits_processor = LogitsProcessor(cfg["vocab_size"])

    def embed_input_ids(self, ids):
        return self.embed_tokens(ids)

    def forward(self, input_ids, positions, intermediate_tensors=None, inputs_embeds=None):
        x = inputs_embeds if inputs_embeds is not None else self.embed_tokens(input_ids)
        for layer in self.layers:
            x = layer(positions, x)
        return self.final_layernorm(x)

    def compute_logits(self, hidden):
        return self.logits_processor(self.lm_head, hidden)

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        params = dict(self.named_parameters())
        modules = dict(self.named_modules())
        loaded: set[str] = set()
        for name, w in weights:
            if "rotary_emb" in name:
                continue
            n = name[len("model."):] if name.startswith("model.") else name
            if n == "lm_head.weight" and self.cfg.get("tie_word_embeddings", False):
                continue
            if ".mamba." in n:
                got = _load_mixer_tensor(params, n, w)
                if got:
                    loaded.add(got)
                continue
            target, shard = n, None
            for fused, part, sid in self.stacked:
                if f".{part}." in n:
                    target, shard = n.replace(part, fused), sid
                    break
            if target not in params:
                continue
            mod = modules[target.rsplit(".", 1)[0]]
            loader = getattr(mod, "weight_loader", None)
            loader(params[target], w, shard) if loader is not None else params[target].data.copy_(w)
            loaded.add(target)
        return loaded
