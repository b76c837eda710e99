"""LayerNorm + rotary decoder families: GPT-NeoX / Pythia (parallel residual, per-head interleaved fused QKV),
StarCoder2 (GQA, bias everywhere, non-gated GELU MLP, sliding window), GPT-J (parallel block, interleaved rotary).

Reference parity: vllm/model_executor/models/gpt_neox.py, starcoder2.py, gpt_j.py.
"""
from __future__ import annotations

from typing import Iterable

import torch
import torch.nn as nn

from vllm_b200.config import VllmConfig
from vllm_b200.layers.attention import Attention
from vllm_b200.layers.common import Activation, LayerNorm, LogitsProcessor, ParallelLMHead, VocabParallelEmbedding
from vllm_b200.layers.linear import ColumnParallelLinear, QKVParallelLinear, RowParallelLinear
from vllm_b200.layers.rotary import get_rope


class _Block(nn.Module):
    """Pre-LN block; ``parallel``: x + attn(ln1 x) + mlp(ln2 x) (NeoX) or shared ln (GPT-J) instead of sequential."""

    def __init__(self, cfg: dict, mc, dtype, prefix: str, *, H, nh, nkv, inter, act, eps, rot_dim, neox_rope, bias,
                 parallel: bool, shared_ln: bool, sliding_window=None, attn_prefix="attention", mlp_prefix="mlp") -> None:
        super().__init__()
        hd = H // nh
        self.qkv_proj = QKVParallelLinear(H, hd, nh, nkv, bias, dtype, None, f"{prefix}.{attn_prefix}.qkv_proj")
        self.o_proj = RowParallelLinear(H, H, bias, dtype=dtype, prefix=f"{prefix}.{attn_prefix}.o_proj")
        max_pos = mc.max_model_len if mc.max_model_len > 0 else cfg.get("max_position_embeddings", 2048)
        rope = get_rope(hd, rot_dim, max_pos, cfg.get("rotary_emb_base", cfg.get("rope_theta", 10000.0)), neox_rope,
                        cfg.get("rope_scaling"))
        self.attn = Attention(self.qkv_proj.num_heads, hd, hd ** -0.5, self.qkv_proj.num_kv_heads,
                              sliding_window=sliding_window, prefix=f"{prefix}.{attn_prefix}.attn", rope=rope)
        self.fc_in = ColumnParallelLinear(H, inter, True if bias or True else False, dtype=dtype, prefix=f"{prefix}.{mlp_prefix}.fc_in")
        self.fc_out = RowParallelLinear(inter, H, True, dtype=dtype, prefix=f"{prefix}.{mlp_prefix}.fc_out")
        self.act = Activation(act)
        self.ln1 = LayerNorm(H, eps, dtype=dtype)
        self.ln2 = None if shared_ln else LayerNorm(H, eps, dtype=dtype)
        self.parallel = parallel
        self.q_size, self.kv_size = self.qkv_proj.q_size, self.qkv_proj.kv_size

    def _attn(self, positions, x):
        q, k, v = self.qkv_proj(x).split([self.q_size, self.kv_size, self.kv_size], dim=-1)
        return self.o_proj(self.attn(q, k, v, positions))

    def forward(self, positions, x):
        h1 = self.ln1(x)
        if self.parallel:
            h2 = h1 if self.ln2 is None else self.ln2(x)
            return x + self._attn(positions, h1) + self.fc_out(self.act(self.fc_in(h2)))
        x = x + self._attn(positions, h1)
        return x + self.fc_out(self.act(self.fc_in(self.ln2(x))))


class _DecoderLM(nn.Module):
    """Shared skeleton; subclasses provide the block arguments and the checkpoint-name mapping."""
    embed_name = "embed"
    block_kwargs: dict = {}

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__()
        mc = vllm_config.model_config
        cfg, dtype = mc.hf_config, mc.torch_dtype
        self.cfg = cfg
        H = cfg["hidden_size"] if "hidden_size" in cfg else cfg["n_embd"]
        self.embed = VocabParallelEmbedding(cfg["vocab_size"], H, dtype)
        self.layers = nn.ModuleList([_Block(cfg, mc, dtype, f"{prefix}layers.{i}", **self._block_args(cfg))
                                     for i in range(cfg.get("num_hidden_layers", cfg.get("n_layer", 0)))])
        self.final_ln = LayerNorm(H, self._block_args(cfg)["eps"], dtype=dtype)
        self.lm_head = ParallelLMHead(cfg["vocab_size"], H, bias=self.lm_head_bias, dtype=dtype)
        if cfg.get("tie_word_embeddings", False):
            self.lm_head.tie_weights(self.embed)
        self.logits_processor = LogitsProcessor(cfg["vocab_size"])

    lm_head_bias = False

    def embed_input_ids(self, ids):
        return self.embed(ids)

    def forward(self, input_ids, positions, intermediate_tensors=None, inputs_embeds=None):
        x = inputs_embeds if inputs_embeds is not None else self.embed(input_ids)
        for layer in self.layers:
            x = layer(positions, x)
        return self.final_ln(x)

    def compute_logits(self, hidden):
        return self.logits_processor(self.lm_head, hidden)

    renames: list[tuple[str, str]] = []

    def _map(self, name: str):
        """checkpoint name → (our name, shard id, tensor transform)"""
        return name, None, None

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        params = dict(self.named_parameters())
        modules = dict(self.named_modules())
        loaded: set[str] = set()
        for name, w in weights:
            name, shard, fn = self._map(name)
            if name is None or name not in params:
                continue
            if fn is not None:
                w = fn(w)
            mod = modules[name.rsplit(".", 1)[0]]
            loader = getattr(mod, "weight_loader", None)
            if loader is not None:
                loader(params[name], w, shard)
            else:
                params[name].data.copy_(w)
            loaded.add(name)
        return loaded


def _rename(name: str, table: list[tuple[str, str]]) -> str:
    for a, b in table:
        if a in name:
            name = name.replace(a, b)
    return name


class GPTNeoXForCausalLM(_DecoderLM):
    def _block_args(self, cfg):
        H, nh = cfg["hidden_size"], cfg["num_attention_heads"]
        hd = H // nh
        return dict(H=H, nh=nh, nkv=nh, inter=cfg["intermediate_size"], act=cfg.get("hidden_act", "gelu"),
                    eps=cfg.get("layer_norm_eps", 1e-5), rot_dim=int(hd * cfg.get("rotary_pct", 0.25)), neox_rope=True,
                    bias=True, parallel=bool(cfg.get("use_parallel_residual", True)), shared_ln=False)

    _T = [("gpt_neox.embed_in.", "embed."), ("gpt_neox.final_layer_norm.", "final_ln."), ("embed_out.", "lm_head."),
          ("gpt_neox.layers.", "layers."), (".input_layernorm.", ".ln1."), (".post_attention_layernorm.", ".ln2."),
          (".attention.query_key_value.", ".qkv_proj."), (".attention.dense.", ".o_proj."),
          (".mlp.dense_h_to_4h.", ".fc_in."), (".mlp.dense_4h_to_h.", ".fc_out.")]

    def _map(self, name):
        if "rotary_emb" in name or name.endswith(".attention.bias") or name.endswith("masked_bias"):
            return None, None, None
        n = _rename(name, self._T)
        fn = None
        if ".qkv_proj." in n:
            nh = self.cfg["num_attention_heads"]

            def fn(w, nh=nh):  # [nh, 3, hd, ...] interleaved per head → [q_all | k_all | v_all]
                hd = w.shape[0] // (3 * nh)
                v = w.view(nh, 3, hd, *w.shape[1:])
                return torch.cat([v[:, i].reshape(nh * hd, *w.shape[1:]) for i in range(3)], dim=0)
        return n, None, fn


class Starcoder2ForCausalLM(_DecoderLM):
    lm_head_bias = False

    def _block_args(self, cfg):
        H, nh = cfg["hidden_size"], cfg["num_attention_heads"]
        return dict(H=H, nh=nh, nkv=cfg.get("num_key_value_heads") or nh, inter=cfg["intermediate_size"],
                    act=cfg.get("hidden_act", "gelu_pytorch_tanh"), eps=cfg.get("norm_epsilon", 1e-5), rot_dim=H // nh,
                    neox_rope=True, bias=bool(cfg.get("use_bias", True)), parallel=False, shared_ln=False,
                    sliding_window=cfg.get("sliding_window"))

    _T = [("model.embed_tokens.", "embed."), ("model.norm.", "final_ln."), ("model.layers.", "layers."),
          (".input_layernorm.", ".ln1."), (".post_attention_layernorm.", ".ln2."), (".self_attn.o_proj.", ".o_proj."),
          (".mlp.c_fc.", ".fc_in."), (".mlp.c_proj.", ".fc_out.")]

    def _map(self, name):
        if "rotary_emb" in name:
            return None, None, None
        if name == "lm_head.weight" and self.cfg.get("tie_word_embeddings", True):
            return None, None, None
        for part, sid in (("q_proj", "q"), ("k_proj", "k"), ("v_proj", "v")):
            if f".self_attn.{part}." in name:
                return _rename(name, self._T).replace(f".self_attn.{part}.", ".qkv_proj."), sid, None
        return _rename(name, self._T), None, None

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        vllm_config.model_config.hf_config.setdefault("tie_word_embeddings", True)
        super().__init__(vllm_config, prefix)


class GPTJForCausalLM(_DecoderLM):
    lm_head_bias = True

    def _block_args(self, cfg):
        H, nh = cfg["n_embd"], cfg["n_head"]
        return dict(H=H, nh=nh, nkv=nh, inter=cfg.get("n_inner") or 4 * H, act=cfg.get("activation_function", "gelu_new"),
                    eps=cfg.get("layer_norm_epsilon", 1e-5), rot_dim=cfg.get("rotary_dim") or H // nh, neox_rope=False,
                    bias=False, parallel=True, shared_ln=True)

    _T = [("transformer.wte.", "embed."), ("transformer.ln_f.", "final_ln."), ("transformer.h.", "layers."),
          (".ln_1.", ".ln1."), (".attn.out_proj.", ".o_proj."), (".mlp.fc_in.", ".fc_in."), (".mlp.fc_out.", ".fc_out.")]

    def _map(self, name):
        if name.endswith("attn.bias") or name.endswith("masked_bias"):
            return None, None, None
        for part, sid in (("q_proj", "q"), ("k_proj", "k"), ("v_proj", "v")):
            if f".attn.{part}." in name:
                return _rename(name, self._T).replace(f".attn.{part}.", ".qkv_proj."), sid, None
        return _rename(name, self._T), None, None
