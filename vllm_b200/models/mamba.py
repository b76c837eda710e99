"""Mamba (attention-free selective state-space LM).

Reference parity: vllm/model_executor/models/mamba.py (``MambaDecoderLayer`` = RMSNorm → mixer → residual,
``MambaModel``, tied LM head) over layers/mamba/mamba_mixer.py.
"""
from __future__ import annotations

import math
from typing import Iterable

import torch
import torch.nn as nn

from vllm_b200.config import VllmConfig
from vllm_b200.layers.common import LogitsProcessor, ParallelLMHead, RMSNorm, VocabParallelEmbedding
from vllm_b200.layers.mamba import MambaMixer
from vllm_b200.parallel.state import get_tp_group


class MambaBlock(nn.Module):
    def __init__(self, cfg: dict, dtype, prefix: str) -> None:
        super().__init__()
        H = cfg["hidden_size"]
        di = cfg.get("intermediate_size") or cfg.get("expand", 2) * H
        dtr = cfg.get("time_step_rank", "auto")
        dtr = math.ceil(H / 16) if dtr == "auto" else int(dtr)
        self.norm = RMSNorm(H, cfg.get("layer_norm_epsilon", 1e-5), dtype)
        # FalconMamba: weight-free RMS normalisation of dt / B / C (the norm modules keep their all-ones weights: the
        # checkpoint has none) — reference: models/falcon_mamba config ``mixer_rms_eps``
        bcdt = float(cfg["mixer_rms_eps"]) if cfg.get("model_type") == "falcon_mamba" and cfg.get("mixer_rms_eps") else None
        self.mixer = MambaMixer(H, di, cfg.get("state_size", 16), cfg.get("conv_kernel", 4), dtr,
                                bool(cfg.get("use_bias", False)), bool(cfg.get("use_conv_bias", True)), dtype,
                                f"{prefix}.mixer", bcdt_norm_eps=bcdt)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x + self.mixer(self.norm(x))


class MambaForCausalLM(nn.Module):
    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__()
        assert get_tp_group().world_size == 1, "Mamba runs replicated (tensor_parallel_size=1)"
        mc = vllm_config.model_config
        cfg, dtype = mc.hf_config, mc.torch_dtype
        self.cfg = cfg
        H = cfg["hidden_size"]
        self.embeddings = VocabParallelEmbedding(cfg["vocab_size"], H, dtype)
        self.layers = nn.ModuleList([MambaBlock(cfg, dtype, f"{prefix}backbone.layers.{i}")
                                     for i in range(cfg["num_hidden_layers"])])
        self.norm_f = RMSNorm(H, cfg.get("layer_norm_epsilon", 1e-5), dtype)
        self.lm_head = ParallelLMHead(cfg["vocab_size"], H, dtype=dtype)
        if cfg.get("tie_word_embeddings", True):
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
            n = n.replace("mixer.conv1d.weight", "mixer.conv_weight").replace("mixer.conv1d.bias", "mixer.conv_bias")
            if n == "lm_head.weight" and self.cfg.get("tie_word_embeddings", True):
                continue
            if n not in params:
                continue
            p = params[n]
            if n.endswith("conv_weight"):
                w = w.reshape(p.shape)                       # [di, 1, k] → [di, k]
            p.data.copy_(w)
            loaded.add(n)
        return loaded
