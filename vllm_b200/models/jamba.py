"""Jamba: hybrid Mamba / attention decoder with interleaved MoE feed-forward blocks.

Reference parity: vllm/model_executor/models/jamba.py (``JambaMambaDecoderLayer`` / ``JambaAttentionDecoderLayer``,
layer-type schedule from ``attn_layer_period/offset`` and ``expert_layer_period/offset``) — the canonical user of
hybrid KV-cache groups: attention layers hold paged K/V, Mamba layers hold one recurrent-state page per request,
both drawn from the same block pool.
"""
from __future__ import annotations

from typing import Iterable

import torch
import torch.nn as nn

from vllm_b200.config import VllmConfig
from vllm_b200.layers.attention import Attention
from vllm_b200.layers.common import LogitsProcessor, ParallelLMHead, RMSNorm, VocabParallelEmbedding
from vllm_b200.layers.linear import QKVParallelLinear, ReplicatedLinear, RowParallelLinear
from vllm_b200.layers.mamba import MambaMixer
from vllm_b200.layers.moe import FusedMoE
from vllm_b200.models.llama import LlamaMLP
from vllm_b200.parallel.state import get_tp_group


class JambaMoE(nn.Module):
    def __init__(self, cfg: dict, dtype, prefix: str) -> None:
        super().__init__()
        H = cfg["hidden_size"]
        self.router = ReplicatedLinear(H, cfg["num_experts"], False, dtype, None, f"{prefix}.router")
        self.experts = FusedMoE(cfg["num_experts"], cfg["num_experts_per_tok"], H, cfg["intermediate_size"], dtype,
                                renormalize=False, act=cfg.get("hidden_act", "silu"), reduce_results=True,
                                prefix=f"{prefix}.experts")

    def forward(self, x):
        return self.experts(x, self.router(x))


class JambaLayer(nn.Module):
    def __init__(self, cfg: dict, mc, dtype, prefix: str, idx: int) -> None:
        super().__init__()
        H, eps = cfg["hidden_size"], cfg.get("rms_norm_eps", 1e-6)
        is_attn = idx % cfg.get("attn_layer_period", 8) == cfg.get("attn_layer_offset", 4)
        is_moe = cfg.get("num_experts", 1) > 1 and idx % cfg.get("expert_layer_period", 2) == cfg.get("expert_layer_offset", 1)
        self.is_attn = is_attn
        if is_attn:
            nh, nkv = cfg["num_attention_heads"], cfg.get("num_key_value_heads") or cfg["num_attention_heads"]
            hd = H // nh
            self.qkv_proj = QKVParallelLinear(H, hd, nh, nkv, False, dtype, None, f"{prefix}.self_attn.qkv_proj")
            self.o_proj = RowParallelLinear(H, H, False, dtype=dtype, prefix=f"{prefix}.self_attn.o_proj")
            self.attn = Attention(self.qkv_proj.num_heads, hd, hd ** -0.5, self.qkv_proj.num_kv_heads,
                                  prefix=f"{prefix}.self_attn.attn")            # Jamba uses no positional encoding
            self.q_size, self.kv_size = self.qkv_proj.q_size, self.qkv_proj.kv_size
        else:
            di = cfg.get("mamba_expand", 2) * H
            dtr = cfg.get("mamba_dt_rank")
            dtr = -(-H // 16) if dtr in (None, "auto") else int(dtr)
            self.mamba = MambaMixer(H, di, cfg.get("mamba_d_state", 16), cfg.get("mamba_d_conv", 4), dtr,
                                    bool(cfg.get("mamba_proj_bias", False)), bool(cfg.get("mamba_conv_bias", True)), dtype,
                                    f"{prefix}.mamba", bcdt_norm_eps=eps)
        self.feed_forward = (JambaMoE(cfg, dtype, f"{prefix}.feed_forward") if is_moe else
                             LlamaMLP(H, cfg["intermediate_size"], cfg.get("hidden_act", "silu"), False, dtype, None,
                                      f"{prefix}.feed_forward"))
        self.input_layernorm = RMSNorm(H, eps, dtype)
        self.pre_ff_layernorm = RMSNorm(H, eps, dtype)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h = self.input_layernorm(x)
        if self.is_attn:
            q, k, v = self.qkv_proj(h).split([self.q_size, self.kv_size, self.kv_size], dim=-1)
            h = self.o_proj(self.attn(q, k, v))
        else:
            h = self.mamba(h)
        x = x + h
        return x + self.feed_forward(self.pre_ff_layernorm(x))


class JambaForCausalLM(nn.Module):
    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__()
        assert get_tp_group().world_size == 1, "Jamba's Mamba layers run replicated (tensor_parallel_size=1)"
        mc = vllm_config.model_config
        cfg, dtype = mc.hf_config, mc.torch_dtype
        self.cfg = cfg
        H = cfg["hidden_size"]
        self.embed_tokens = VocabParallelEmbedding(cfg["vocab_size"], H, dtype)
        self.layers = nn.ModuleList([JambaLayer(cfg, mc, dtype, f"{prefix}model.layers.{i}", i)
                                     for i in range(cfg["num_hidden_layers"])])
        self.final_layernorm = RMSNorm(H, cfg.get("rms_norm_eps", 1e-6), dtype)
        self.lm_head = ParallelLMHead(cfg["vocab_size"], H, dtype=dtype)
        if cfg.get("tie_word_embeddings", False):
            self.lm_head.tie_weights(self.embed_tokens)
        self.logits_processor = LogitsProcessor(cfg["vocab_size"])

    def embed_input_ids(self, ids):
        return self.embed_tokens(ids)

    def forward(self, input_ids, positions, intermediate_tensors=None, inputs_embeds=None):
        x = inputs_embeds if inputs_embeds is not None else self.embed_tokens(input_ids)
        for layer in self.layers:
            x = layer(x)
        return self.final_layernorm(x)

    def compute_logits(self, hidden):
        return self.logits_processor(self.lm_head, hidden)

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        params = dict(self.named_parameters())
        modules = dict(self.named_modules())
        loaded: set[str] = set()

        def put(name, w, shard=None, expert=None):
            if name not in params:
                return
            mod = modules[name.rsplit(".", 1)[0]]
            if expert is not None:
                mod.weight_loader(params[name], w, shard, expert)
            elif getattr(mod, "weight_loader", None) is not None:
                mod.weight_loader(params[name], w, shard)
            else:
                params[name].data.copy_(w.reshape(params[name].shape))
            loaded.add(name)
        for name, w in weights:
            n = name[len("model."):] if name.startswith("model.") else name
            n = n.replace("mamba.conv1d.weight", "mamba.conv_weight").replace("mamba.conv1d.bias", "mamba.conv_bias")
            if ".self_attn." in n:
                for part, sid in (("q_proj", "q"), ("k_proj", "k"), ("v_proj", "v")):
                    if f".{part}." in n:
                        put(n.replace(f"self_attn.{part}", "qkv_proj"), w, sid)
                        break
                else:
                    put(n.replace("self_attn.o_proj", "o_proj"), w)
            elif n.endswith("feed_forward.experts.gate_up_proj"):          # fused [E, 2I, H]
                I = w.shape[1] // 2
                base = n[: -len("gate_up_proj")] + "w13_weight"
                for e in range(w.shape[0]):
                    put(base, w[e, :I], "w1", e)
                    put(base, w[e, I:], "w3", e)
            elif n.endswith("feed_forward.experts.down_proj"):
                base = n[: -len("down_proj")] + "w2_weight"
                for e in range(w.shape[0]):
                    put(base, w[e], "w2", e)
            elif ".feed_forward.experts." in n and n.split(".feed_forward.experts.")[1].split(".")[0].isdigit():
                head, tail = n.split(".feed_forward.experts.")
                e, proj = tail.split(".")[0], tail.split(".")[1]
                shard = {"gate_proj": "w1", "up_proj": "w3", "down_proj": "w2"}[proj]
                put(f"{head}.feed_forward.experts.{'w2_weight' if shard == 'w2' else 'w13_weight'}", w, shard, int(e))
            elif ".feed_forward.gate_proj." in n:
                put(n.replace(".gate_proj.", ".gate_up_proj."), w, 0)
            elif ".feed_forward.up_proj." in n:
                put(n.replace(".up_proj.", ".gate_up_proj."), w, 1)
            else:
                put(n, w)
        return loaded
