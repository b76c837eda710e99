"""OPT and GPT-2 (learned positions, LayerNorm, non-gated MLP).

Architecture parity: vllm/model_executor/models/opt.py, gpt2.py.  OPT-125m greedy
decode on the CPU reference path is BASELINE config #1 (plumbing check).
"""
from __future__ import annotations

from typing import Iterable

import torch
import torch.nn as nn

from vllm_b200.config import VllmConfig
from vllm_b200.layers.attention import Attention
from vllm_b200.layers.common import Activation, LayerNorm, LogitsProcessor, ParallelLMHead, VocabParallelEmbedding
from vllm_b200.layers.linear import ColumnParallelLinear, QKVParallelLinear, RowParallelLinear


class OPTLayer(nn.Module):
    def __init__(self, cfg, dtype, prefix) -> None:
        super().__init__()
        H, nh = cfg["hidden_size"], cfg["num_attention_heads"]
        bias = cfg.get("enable_bias", True)
        self.head_dim = H // nh
        self.qkv_proj = QKVParallelLinear(H, self.head_dim, nh, nh, bias, dtype, prefix=f"{prefix}.self_attn.qkv_proj")
        self.out_proj = RowParallelLinear(H, H, bias, dtype=dtype, prefix=f"{prefix}.self_attn.out_proj")
        self.attn = Attention(self.qkv_proj.num_heads, self.head_dim, self.head_dim ** -0.5,
                              self.qkv_proj.num_kv_heads, prefix=f"{prefix}.self_attn.attn")
        self.do_ln_before = cfg.get("do_layer_norm_before", True)
        self.self_attn_layer_norm = LayerNorm(H, 1e-5, cfg.get("layer_norm_elementwise_affine", True), dtype)
        self.fc1 = ColumnParallelLinear(H, cfg["ffn_dim"], bias, dtype=dtype, prefix=f"{prefix}.fc1")
        self.fc2 = RowParallelLinear(cfg["ffn_dim"], H, bias, dtype=dtype, prefix=f"{prefix}.fc2")
        self.act = Activation(cfg.get("activation_function", "relu"))
        self.final_layer_norm = LayerNorm(H, 1e-5, cfg.get("layer_norm_elementwise_affine", True), dtype)

    def forward(self, h):
        r = h
        if self.do_ln_before:
            h = self.self_attn_layer_norm(h)
        q, k, v = self.qkv_proj(h).split([self.qkv_proj.q_size, self.qkv_proj.kv_size, self.qkv_proj.kv_size], -1)
        h = r + self.out_proj(self.attn(q, k, v))
        if not self.do_ln_before:
            h = self.self_attn_layer_norm(h)
        r = h
        if self.do_ln_before:
            h = self.final_layer_norm(h)
        h = r + self.fc2(self.act(self.fc1(h)))
        if not self.do_ln_before:
            h = self.final_layer_norm(h)
        return h


class OPTForCausalLM(nn.Module):
    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__()
        mc = vllm_config.model_config
        cfg = self.cfg = mc.hf_config
        dtype = mc.torch_dtype
        H, E = cfg["hidden_size"], cfg.get("word_embed_proj_dim", cfg["hidden_size"])
        self.embed_tokens = VocabParallelEmbedding(cfg["vocab_size"], E, dtype)
        self.embed_positions = nn.Embedding(cfg["max_position_embeddings"] + 2, H, dtype=dtype)
        self.embed_positions.weight.requires_grad_(False)
        self.project_in = nn.Linear(E, H, bias=False, dtype=dtype) if E != H else None
        self.project_out = nn.Linear(H, E, bias=False, dtype=dtype) if E != H else None
        self.layers = nn.ModuleList([OPTLayer(cfg, dtype, f"model.decoder.layers.{i}")
                                     for i in range(cfg["num_hidden_layers"])])
        self.final_layer_norm = LayerNorm(H, 1e-5, True, dtype) \
            if cfg.get("do_layer_norm_before", True) and not cfg.get("_remove_final_layer_norm") else None
        self.lm_head = ParallelLMHead(cfg["vocab_size"], E, dtype=dtype)
        if cfg.get("tie_word_embeddings", True):
            self.lm_head.tie_weights(self.embed_tokens)
        self.logits_processor = LogitsProcessor(cfg["vocab_size"])

    embed_scale = 1.0

    def embed_input_ids(self, ids):
        h = self.embed_tokens(ids)
        return h * self.embed_scale if self.embed_scale != 1.0 else h

    def forward(self, input_ids, positions, intermediate_tensors=None, inputs_embeds=None):
        h = inputs_embeds if inputs_embeds is not None else self.embed_input_ids(input_ids)
        if self.project_in is not None:
            h = self.project_in(h)
        h = h + self.embed_positions(positions + 2)
        for layer in self.layers:
            h = layer(h)
        if self.final_layer_norm is not None:
            h = self.final_layer_norm(h)
        if self.project_out is not None:
            h = self.project_out(h)
        return h

    def compute_logits(self, hidden):
        return self.logits_processor(self.lm_head, hidden)

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        params = dict(self.named_parameters())
        modules = dict(self.named_modules())
        loaded = set()
        for name, w in weights:
            name = self._rename(name)
            if name is None:
                continue
            if name.startswith("lm_head") and self.cfg.get("tie_word_embeddings", True):
                continue
            name = name.replace("self_attn.out_proj", "out_proj")
            shard = None
            for part, sid in (("q_proj", "q"), ("k_proj", "k"), ("v_proj", "v")):
                if f"self_attn.{part}" in name:
                    name, shard = name.replace(f"self_attn.{part}", "qkv_proj"), sid
            if name not in params:
                continue
            mod = modules[name.rsplit(".", 1)[0]]
            loader = getattr(mod, "weight_loader", None)
            if loader is not None:
                loader(params[name], w, shard)
            else:
                params[name].data.copy_(w)
            loaded.add(name)
        return loaded


    def _rename(self, name: str):
        return name.replace("model.decoder.", "").replace("decoder.", "")


def _as_opt_config(cfg: dict, *, ffn: str, act: str) -> dict:
    cfg.setdefault("ffn_dim", cfg.get(ffn))
    cfg.setdefault("activation_function", cfg.get(act, "gelu"))
    cfg.setdefault("do_layer_norm_before", True)
    cfg.setdefault("enable_bias", True)
    return cfg


class BioGptForCausalLM(OPTForCausalLM):
    """BioGPT: the OPT block (pre-LayerNorm, learned positions with offset 2) with sqrt(H)-scaled embeddings and GELU
    (reference: models via the Transformers back-end; HF ``modeling_biogpt``)."""

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        cfg = _as_opt_config(vllm_config.model_config.hf_config, ffn="intermediate_size", act="hidden_act")
        super().__init__(vllm_config, prefix)
        self.embed_scale = float(cfg["hidden_size"]) ** 0.5 if cfg.get("scale_embedding", True) else 1.0

    def _rename(self, name: str):
        name = name.replace("biogpt.", "")
        if name.startswith("layer_norm."):
            name = "final_layer_norm." + name[len("layer_norm."):]
        return name.replace("output_projection.", "lm_head.")


class _FixedTable(nn.Module):
    def __init__(self, table: torch.Tensor) -> None:
        super().__init__()
        self.register_buffer("table", table, persistent=False)

    def forward(self, idx: torch.Tensor) -> torch.Tensor:
        return self.table[idx]


class XGLMForCausalLM(OPTForCausalLM):
    """XGLM: the OPT block with sqrt(H)-scaled embeddings and fixed sinusoidal positions (offset 2)
    (HF ``modeling_xglm``; fairseq-style table: [sin | cos] halves, padding row zero)."""

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        import math
        cfg = _as_opt_config(vllm_config.model_config.hf_config, ffn="ffn_dim", act="activation_function")
        cfg.setdefault("hidden_size", cfg.get("d_model"))
        cfg.setdefault("num_attention_heads", cfg.get("attention_heads"))
        cfg.setdefault("num_hidden_layers", cfg.get("num_layers"))
        super().__init__(vllm_config, prefix)
        H = cfg["hidden_size"]
        self.embed_scale = math.sqrt(H) if cfg.get("scale_embedding", True) else 1.0
        n = cfg["max_position_embeddings"] + 2
        half = H // 2
        freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000.0) / (half - 1)))
        ang = torch.arange(n, dtype=torch.float32)[:, None] * freq[None]
        table = torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)
        if H % 2:
            table = torch.cat([table, torch.zeros(n, 1)], dim=1)
        pad = cfg.get("pad_token_id")
        if pad is not None and pad < n:
            table[pad] = 0
        self.embed_positions = _FixedTable(table.to(vllm_config.model_config.torch_dtype))   # a buffer: not a weight

    def _rename(self, name: str):
        if "embed_positions" in name:
            return None                      # the table is a function of the position, not a checkpoint tensor
        name = name.replace("model.", "", 1) if name.startswith("model.") else name
        if name.startswith("layer_norm."):
            name = "final_layer_norm." + name[len("layer_norm."):]
        return name


class GPT2Block(nn.Module):
    def __init__(self, cfg, dtype, prefix) -> None:
        super().__init__()
        H, nh = cfg["n_embd"], cfg["n_head"]
        inner = cfg.get("n_inner") or 4 * H
        eps = cfg.get("layer_norm_epsilon", 1e-5)
        self.head_dim = H // nh
        self.ln_1 = LayerNorm(H, eps, True, dtype)
        nkv = 1 if cfg.get("multi_query") else nh          # GPT-BigCode (StarCoder): one shared K/V head
        self.c_attn = QKVParallelLinear(H, self.head_dim, nh, nkv, True, dtype, prefix=f"{prefix}.attn.c_attn")
        self.c_proj = RowParallelLinear(H, H, True, dtype=dtype, prefix=f"{prefix}.attn.c_proj")
        self.attn = Attention(self.c_attn.num_heads, self.head_dim, self.head_dim ** -0.5,
                              self.c_attn.num_kv_heads, prefix=f"{prefix}.attn.attn")
        self.ln_2 = LayerNorm(H, eps, True, dtype)
        self.c_fc = ColumnParallelLinear(H, inner, True, dtype=dtype, prefix=f"{prefix}.mlp.c_fc")
        self.mlp_proj = RowParallelLinear(inner, H, True, dtype=dtype, prefix=f"{prefix}.mlp.c_proj")
        self.act = Activation(cfg.get("activation_function", "gelu_new"))

    def forward(self, h):
        x = self.ln_1(h)
        q, k, v = self.c_attn(x).split([self.c_attn.q_size, self.c_attn.kv_size, self.c_attn.kv_size], -1)
        h = h + self.c_proj(self.attn(q, k, v))
        return h + self.mlp_proj(self.act(self.c_fc(self.ln_2(h))))


class GPT2LMHeadModel(nn.Module):
    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__()
        mc = vllm_config.model_config
        cfg = self.cfg = mc.hf_config
        dtype = mc.torch_dtype
        H = cfg["n_embd"]
        self.wte = VocabParallelEmbedding(cfg["vocab_size"], H, dtype)
        self.wpe = nn.Embedding(cfg["n_positions"], H, dtype=dtype)
        self.wpe.weight.requires_grad_(False)
        self.h = nn.ModuleList([GPT2Block(cfg, dtype, f"transformer.h.{i}") for i in range(cfg["n_layer"])])
        self.ln_f = LayerNorm(H, cfg.get("layer_norm_epsilon", 1e-5), True, dtype)
        self.lm_head = ParallelLMHead(cfg["vocab_size"], H, dtype=dtype)
        self.lm_head.tie_weights(self.wte)
        self.logits_processor = LogitsProcessor(cfg["vocab_size"])

    def embed_input_ids(self, ids):
        return self.wte(ids)

    def forward(self, input_ids, positions, intermediate_tensors=None, inputs_embeds=None):
        h = (inputs_embeds if inputs_embeds is not None else self.wte(input_ids)) + self.wpe(positions)
        for blk in self.h:
            h = blk(h)
        return self.ln_f(h)

    def compute_logits(self, hidden):
        return self.logits_processor(self.lm_head, hidden)

    def load_weights(self, weights) -> set[str]:
        params = dict(self.named_parameters())
        modules = dict(self.named_modules())
        loaded = set()
        for name, w in weights:
            name = name.replace("transformer.", "")
            if name.startswith("lm_head") or ".attn.bias" in name or ".attn.masked_bias" in name:
                continue
            name = name.replace("attn.c_attn", "c_attn").replace("attn.c_proj", "c_proj")
            name = name.replace("mlp.c_fc", "c_fc").replace("mlp.c_proj", "mlp_proj")
            if name not in params:
                continue
            if any(s in name for s in ("c_attn", "c_proj", "c_fc", "mlp_proj")) and name.endswith(".weight"):
                w = w.t()  # HF GPT-2 uses Conv1D (in, out)
            mod = modules[name.rsplit(".", 1)[0]]
            loader = getattr(mod, "weight_loader", None)
            if loader is not None:
                loader(params[name], w.contiguous(), None)
            else:
                params[name].data.copy_(w)
            loaded.add(name)
        return loaded


class GPTBigCodeForCausalLM(GPT2LMHeadModel):
    """StarCoder / SantaCoder: GPT-2 skeleton with multi-query attention and ordinary ``nn.Linear`` weight layout
    (reference: models/gpt_bigcode.py)."""

    def load_weights(self, weights) -> set[str]:
        def fix():
            for name, w in weights:
                # undo the transpose the GPT-2 loader applies to Conv1D weights: these checkpoints store [out, in]
                if any(s in name for s in ("c_attn", "c_proj", "c_fc")) and name.endswith(".weight"):
                    w = w.t()
                yield name, w
        return super().load_weights(fix())
