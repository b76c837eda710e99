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
from vllm_b200.parallel.state import get_tp_group


class _Block(nn.Module):
    """Pre-LN block; ``parallel``: x + attn(ln1 x) + mlp(ln2 x) (NeoX) or shared ln (GPT-J) instead of sequential."""

    def __init__(self, cfg: dict, mc, dtype, prefix: str, *, H, nh, nkv, inter, act, eps, rot_dim, neox_rope, bias,
                 parallel: bool, shared_ln: bool, sliding_window=None, attn_prefix="attention", mlp_prefix="mlp",
                 alibi_slopes=None, mlp_bias: bool = True) -> None:
        super().__init__()
        hd = H // nh
        self.qkv_proj = QKVParallelLinear(H, hd, nh, nkv, bias, dtype, None, f"{prefix}.{attn_prefix}.qkv_proj")
        self.o_proj = RowParallelLinear(H, H, bias, dtype=dtype, prefix=f"{prefix}.{attn_prefix}.o_proj")
        max_pos = mc.max_model_len if mc.max_model_len > 0 else cfg.get("max_position_embeddings", 2048)
        rope = None
        if alibi_slopes is None:
            rope = get_rope(hd, rot_dim, max_pos, cfg.get("rotary_emb_base", cfg.get("rope_theta", 10000.0)), neox_rope,
                            cfg.get("rope_scaling"))
        else:                                    # positional information comes from the attention bias: this rank's heads
            n_local = self.qkv_proj.num_heads
            r = get_tp_group().rank_in_group if get_tp_group().world_size > 1 else 0
            alibi_slopes = alibi_slopes[r * n_local:(r + 1) * n_local]
        self.attn = Attention(self.qkv_proj.num_heads, hd, hd ** -0.5, self.qkv_proj.num_kv_heads,
                              sliding_window=sliding_window, prefix=f"{prefix}.{attn_prefix}.attn", rope=rope,
                              alibi_slopes=alibi_slopes)
        self.fc_in = ColumnParallelLinear(H, inter, mlp_bias, dtype=dtype, prefix=f"{prefix}.{mlp_prefix}.fc_in")
        self.fc_out = RowParallelLinear(inter, H, mlp_bias, dtype=dtype, prefix=f"{prefix}.{mlp_prefix}.fc_out")
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
        H = cfg.get("hidden_size") or cfg.get("n_embd") or cfg["d_model"]
        self.embed = VocabParallelEmbedding(cfg["vocab_size"], H, dtype)
        self.layers = nn.ModuleList([_Block(cfg, mc, dtype, f"{prefix}layers.{i}", **self._block_args(cfg))
                                     for i in range(cfg.get("num_hidden_layers", cfg.get("n_layer", cfg.get("n_layers", 0))))])
        self.embed_ln = LayerNorm(H, self._block_args(cfg)["eps"], dtype=dtype) if self.has_embed_ln else None
        self.final_ln = LayerNorm(H, self._block_args(cfg)["eps"], dtype=dtype)
        self.lm_head = ParallelLMHead(cfg["vocab_size"], H, bias=self.lm_head_bias, dtype=dtype)
        if cfg.get("tie_word_embeddings", False):
            self.lm_head.tie_weights(self.embed)
        self.logits_processor = LogitsProcessor(cfg["vocab_size"])

    lm_head_bias = False
    has_embed_ln = False

    def embed_input_ids(self, ids):
        x = self.embed(ids)
        return self.embed_ln(x) if self.embed_ln is not None else x

    def forward(self, input_ids, positions, intermediate_tensors=None, inputs_embeds=None):
        x = inputs_embeds if inputs_embeds is not None else self.embed_input_ids(input_ids)
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


# ------------------------------------------------------------------------------------------------ ALiBi families
def alibi_slopes(n_heads: int) -> torch.Tensor:
    """Press et al. slopes as HF's ``build_alibi_tensor`` computes them (Bloom, Falcon-RW)."""
    import math
    cp2 = 2 ** math.floor(math.log2(n_heads))
    base = 2.0 ** (-(2.0 ** -(math.log2(cp2) - 3)))
    slopes = [base ** i for i in range(1, cp2 + 1)]
    if cp2 != n_heads:
        extra = 2.0 ** (-(2.0 ** -(math.log2(2 * cp2) - 3)))
        slopes += [extra ** i for i in range(1, 2 * (n_heads - cp2) + 1, 2)]
    return torch.tensor(slopes, dtype=torch.float32)


def mpt_alibi_slopes(n_heads: int, bias_max: float = 8.0) -> torch.Tensor:
    import math
    n2 = 2 ** math.ceil(math.log2(n_heads))
    m = torch.arange(1, n2 + 1, dtype=torch.float32) * (bias_max / n2)
    slopes = 1.0 / torch.pow(2.0, m)
    if n2 != n_heads:
        slopes = torch.cat([slopes[1::2], slopes[::2]])[:n_heads]
    return slopes


def _deinterleave_qkv(nh: int):
    def fn(w):                                   # [nh, 3, hd, ...] → [q_all | k_all | v_all]
        hd = w.shape[0] // (3 * nh)
        v = w.view(nh, 3, hd, *w.shape[1:])
        return torch.cat([v[:, i].reshape(nh * hd, *w.shape[1:]) for i in range(3)], dim=0)
    return fn


class BloomForCausalLM(_DecoderLM):
    """Reference parity: vllm/model_executor/models/bloom.py (ALiBi, embedding LayerNorm, per-head fused QKV)."""
    has_embed_ln = True

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        vllm_config.model_config.hf_config.setdefault("tie_word_embeddings", True)
        super().__init__(vllm_config, prefix)

    def _block_args(self, cfg):
        H, nh = cfg["hidden_size"], cfg.get("n_head") or cfg["num_attention_heads"]
        return dict(H=H, nh=nh, nkv=nh, inter=4 * H, act="gelu_pytorch_tanh", eps=cfg.get("layer_norm_epsilon", 1e-5),
                    rot_dim=0, neox_rope=True, bias=True, parallel=False, shared_ln=False, alibi_slopes=alibi_slopes(nh))

    _T = [("transformer.word_embeddings_layernorm.", "embed_ln."), ("transformer.word_embeddings.", "embed."),
          ("transformer.ln_f.", "final_ln."), ("transformer.h.", "layers."), (".input_layernorm.", ".ln1."),
          (".post_attention_layernorm.", ".ln2."), (".self_attention.query_key_value.", ".qkv_proj."),
          (".self_attention.dense.", ".o_proj."), (".mlp.dense_h_to_4h.", ".fc_in."), (".mlp.dense_4h_to_h.", ".fc_out.")]

    def _map(self, name):
        if name == "lm_head.weight":
            return None, None, None
        n = _rename(name, self._T)
        nh = self.cfg.get("n_head") or self.cfg["num_attention_heads"]
        return n, None, (_deinterleave_qkv(nh) if ".qkv_proj." in n else None)


class MPTForCausalLM(_DecoderLM):
    """Reference parity: vllm/model_executor/models/mpt.py (ALiBi with ``alibi_bias_max``, bias-free blocks)."""

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        vllm_config.model_config.hf_config.setdefault("tie_word_embeddings", True)
        super().__init__(vllm_config, prefix)

    def _block_args(self, cfg):
        H, nh = cfg["d_model"], cfg["n_heads"]
        ac = cfg.get("attn_config") or {}
        bias = not cfg.get("no_bias", True)
        return dict(H=H, nh=nh, nkv=nh, inter=int(cfg.get("expansion_ratio", 4) * H), act="gelu",
                    eps=cfg.get("layer_norm_epsilon", 1e-5), rot_dim=0, neox_rope=True, bias=bias, mlp_bias=bias,
                    parallel=False, shared_ln=False,
                    alibi_slopes=mpt_alibi_slopes(nh, ac.get("alibi_bias_max", 8)) if ac.get("alibi", True) else None)

    _T = [("transformer.wte.", "embed."), ("transformer.norm_f.", "final_ln."), ("transformer.blocks.", "layers."),
          (".norm_1.", ".ln1."), (".norm_2.", ".ln2."), (".attn.Wqkv.", ".qkv_proj."), (".attn.out_proj.", ".o_proj."),
          (".ffn.up_proj.", ".fc_in."), (".ffn.down_proj.", ".fc_out.")]

    def _map(self, name):
        if name == "lm_head.weight":
            return None, None, None
        return _rename(name, self._T), None, None


class FalconForCausalLM(_DecoderLM):
    """Reference parity: vllm/model_executor/models/falcon.py — the three generations share one block:
    falcon-rw (ALiBi, sequential, per-head fused QKV), falcon-7b (rotary, multi-query, parallel attention with one
    LayerNorm), falcon-40b/180b (``new_decoder_architecture``: GQA, separate ln_attn / ln_mlp)."""

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        vllm_config.model_config.hf_config.setdefault("tie_word_embeddings", True)
        super().__init__(vllm_config, prefix)

    def _nkv(self, cfg):
        nh = cfg["num_attention_heads"]
        if cfg.get("new_decoder_architecture"):
            return cfg.get("num_kv_heads") or nh
        return 1 if cfg.get("multi_query", True) else nh

    def _block_args(self, cfg):
        H, nh = cfg["hidden_size"], cfg["num_attention_heads"]
        new = bool(cfg.get("new_decoder_architecture"))
        par = bool(cfg.get("parallel_attn", True)) or new
        two_ln = new and cfg.get("num_ln_in_parallel_attn", 2) != 1
        slopes = alibi_slopes(nh) * (H // nh) ** -0.5 if cfg.get("alibi") else None     # HF scales bias with the scores
        return dict(H=H, nh=nh, nkv=self._nkv(cfg), inter=cfg.get("ffn_hidden_size") or 4 * H, act="gelu",
                    eps=cfg.get("layer_norm_epsilon", 1e-5), rot_dim=H // nh, neox_rope=True, bias=bool(cfg.get("bias")),
                    mlp_bias=bool(cfg.get("bias")), parallel=par, shared_ln=par and not two_ln, alibi_slopes=slopes)

    _T = [("transformer.word_embeddings.", "embed."), ("transformer.ln_f.", "final_ln."), ("transformer.h.", "layers."),
          (".ln_attn.", ".ln1."), (".ln_mlp.", ".ln2."), (".input_layernorm.", ".ln1."),
          (".post_attention_layernorm.", ".ln2."), (".self_attention.query_key_value.", ".qkv_proj."),
          (".self_attention.dense.", ".o_proj."), (".mlp.dense_h_to_4h.", ".fc_in."), (".mlp.dense_4h_to_h.", ".fc_out.")]

    def _map(self, name):
        if name == "lm_head.weight" and self.cfg.get("tie_word_embeddings", True):
            return None, None, None
        n = _rename(name, self._T)
        fn = None
        if ".qkv_proj." in n:
            cfg = self.cfg
            nh, nkv = cfg["num_attention_heads"], self._nkv(cfg)
            if cfg.get("new_decoder_architecture"):
                def fn(w, nh=nh, nkv=nkv):        # [nkv, nh/nkv + 2, hd, ...]
                    g = nh // nkv
                    hd = w.shape[0] // (nkv * (g + 2))
                    v = w.view(nkv, g + 2, hd, *w.shape[1:])
                    return torch.cat([v[:, :g].reshape(nh * hd, *w.shape[1:]), v[:, g].reshape(nkv * hd, *w.shape[1:]),
                                      v[:, g + 1].reshape(nkv * hd, *w.shape[1:])], dim=0)
            elif nkv == nh:
                fn = _deinterleave_qkv(nh)
            # multi-query: already [q_all | k | v]
        return n, None, fn
