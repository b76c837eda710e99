"""Llama-family decoder (Llama 2/3, Mistral, Qwen2, Qwen3, Gemma via flags).

Architecture parity with the reference's vllm/model_executor/models/llama.py
(``LlamaDecoderLayer.forward`` :310-327, ``LlamaModel.forward`` :400-438,
``load_weights`` stacked-param mapping :440-500), qwen2.py / qwen3.py / gemma.py.

The layer loop is written around ``RowParallelLinear.forward_reduce_norm``: each
row-parallel GEMM hands back the *already normalised* input of the next block,
so under TP>1 every layer is 2 fused GEMM+all-reduce+RMSNorm kernels with no
stand-alone collective or norm launch.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.nn as nn

from vllm_b200.config import VllmConfig
from vllm_b200.layers.attention import Attention
from vllm_b200.layers.common import (LogitsProcessor, ParallelLMHead, RMSNorm, VocabParallelEmbedding,
                                     get_act_and_mul)
from vllm_b200.layers.linear import MergedColumnParallelLinear, QKVParallelLinear, RowParallelLinear
from vllm_b200.layers.rotary import get_rope
from vllm_b200.parallel.state import get_pp_group, get_tp_group
from vllm_b200.parallel.utils import get_pp_indices


class LlamaMLP(nn.Module):
    def __init__(self, hidden, inter, act: str, bias: bool, dtype, quant_config, prefix) -> None:
        super().__init__()
        from vllm_b200 import envs
        self.gate_up_proj = MergedColumnParallelLinear(hidden, [inter, inter], bias, dtype, quant_config,
                                                       f"{prefix}.gate_up_proj",
                                                       interleave=(act == "silu" and envs.VLLM_B200_FUSED_ACT))
        self.down_proj = RowParallelLinear(inter, hidden, bias, dtype=dtype, quant_config=quant_config,
                                           prefix=f"{prefix}.down_proj")
        self.act = get_act_and_mul(act)

    def _gated(self, x):
        if self.gate_up_proj.interleave:
            return self.gate_up_proj.forward_gated_silu(x)
        return self.act(self.gate_up_proj(x))

    def forward(self, x):
        return self.down_proj(self._gated(x))

    def forward_reduce_norm(self, x, residual, norm):
        return self.down_proj.forward_reduce_norm(self._gated(x), residual, norm)


class LlamaAttention(nn.Module):
    def __init__(self, cfg: dict, mc, dtype, quant_config, prefix: str, layer_idx: int) -> None:
        super().__init__()
        hidden = cfg["hidden_size"]
        self.total_heads = cfg["num_attention_heads"]
        self.total_kv = cfg.get("num_key_value_heads") or self.total_heads
        self.head_dim = cfg.get("head_dim") or hidden // self.total_heads
        bias = bool(cfg.get("attention_bias", False))
        self.qkv_proj = QKVParallelLinear(hidden, self.head_dim, self.total_heads, self.total_kv, bias,
                                          dtype, quant_config, f"{prefix}.qkv_proj")
        self.o_proj = RowParallelLinear(self.total_heads * self.head_dim, hidden,
                                        bool(cfg.get("attention_out_bias", False)), dtype=dtype,
                                        quant_config=quant_config, prefix=f"{prefix}.o_proj")
        self.q_size, self.kv_size = self.qkv_proj.q_size, self.qkv_proj.kv_size
        from vllm_b200 import envs
        self._l2_prefetch = bool(envs.VLLM_B200_L2_PREFETCH)
        self.qk_norm = bool(cfg.get("qk_norm", False))
        self.qk_norm_full = cfg.get("qk_norm") == "full"          # OLMo-2 / OLMoE: one RMSNorm over the whole projection
        if self.qk_norm_full:
            assert get_tp_group().world_size == 1, "whole-projection q/k norms need the unsharded projection"
            eps = cfg.get("rms_norm_eps", 1e-6)
            self.q_norm = RMSNorm(self.total_heads * self.head_dim, eps, dtype)
            self.k_norm = RMSNorm(self.total_kv * self.head_dim, eps, dtype)
        elif self.qk_norm:
            eps = cfg.get("rms_norm_eps", 1e-6)
            self.q_norm = RMSNorm(self.head_dim, eps, dtype)
            self.k_norm = RMSNorm(self.head_dim, eps, dtype)
        rot = int(self.head_dim * cfg.get("partial_rotary_factor", 1.0))
        rope = get_rope(self.head_dim, rot, mc.max_model_len if mc.max_model_len > 0 else
                        cfg.get("max_position_embeddings", 8192), cfg.get("rope_theta", 10000.0),
                        not cfg.get("rope_interleaved", False), cfg.get("rope_scaling"))
        sw = cfg.get("sliding_window")
        if sw is not None and cfg.get("max_window_layers") is not None and layer_idx < cfg["max_window_layers"]:
            sw = None  # qwen2: only upper layers use SWA
        if cfg.get("use_sliding_window") is False:
            sw = None
        types = cfg.get("layer_types")
        if sw is not None and isinstance(types, list) and layer_idx < len(types) and types[layer_idx] in ("full_attention", "sliding_attention"):
            sw = sw if types[layer_idx] == "sliding_attention" else None       # per-layer windows (Mellum, Qwen3 SWA variants)
        self.attn = Attention(self.qkv_proj.num_heads, self.head_dim, self.head_dim ** -0.5,
                              self.qkv_proj.num_kv_heads, sliding_window=sw,
                              logits_soft_cap=cfg.get("attn_logit_softcapping") or 0.0,
                              prefix=f"{prefix}.attn", rope=rope, causal=bool(cfg.get("is_causal", True)))

    def forward(self, positions, hidden):
        qkv = self.qkv_proj(hidden)
        if hidden.is_cuda and hidden.shape[0] <= 256 and self._l2_prefetch:
            # decode: the o_proj weights stream into L2 while rope + attention run (csrc/cache_ops.cu::l2_prefetch_kernel)
            from vllm_b200.ops import cuda_ops
            cuda_ops().l2_prefetch_async(self.o_proj.weight)
        q, k, v = qkv.split([self.q_size, self.kv_size, self.kv_size], dim=-1)
        if self.qk_norm_full:
            q, k = self.q_norm(q.contiguous()), self.k_norm(k.contiguous())
        elif self.qk_norm:
            # per-head q/k RMSNorm: fused with RoPE and the KV insert on CUDA (one launch), applied by the layer otherwise
            return self.attn(q, k, v, positions, qk_norm=(self.q_norm, self.k_norm))
        return self.attn(q, k, v, positions)


class LlamaDecoderLayer(nn.Module):
    def __init__(self, cfg, mc, dtype, quant_config, prefix: str, layer_idx: int, gemma: bool = False) -> None:
        super().__init__()
        eps = cfg.get("rms_norm_eps", 1e-6)
        self.self_attn = LlamaAttention(cfg, mc, dtype, quant_config, f"{prefix}.self_attn", layer_idx)
        self.mlp = LlamaMLP(cfg["hidden_size"], cfg["intermediate_size"], cfg.get("hidden_act", "silu"),
                            bool(cfg.get("mlp_bias", False)), dtype, quant_config, f"{prefix}.mlp")
        self.input_layernorm = RMSNorm(cfg["hidden_size"], eps, dtype, gemma)
        self.post_attention_layernorm = RMSNorm(cfg["hidden_size"], eps, dtype, gemma)

    def forward(self, positions, hidden, residual, next_norm: Optional[nn.Module]):
        """``hidden`` is the normalised block input, ``residual`` the raw stream."""
        a = self.self_attn(positions, hidden)
        hidden, residual = self.self_attn.o_proj.forward_reduce_norm(a, residual, self.post_attention_layernorm)
        return self.mlp.forward_reduce_norm(hidden, residual, next_norm)


class LlamaModel(nn.Module):
    layer_cls = LlamaDecoderLayer

    def __init__(self, vllm_config: VllmConfig, prefix: str = "model", quant_config=None) -> None:
        super().__init__()
        mc = vllm_config.model_config
        cfg = mc.hf_config
        dtype = mc.torch_dtype
        self.cfg = cfg
        pp = get_pp_group()
        self.pp_rank, self.pp_size = max(pp.rank_in_group, 0), pp.world_size
        self.start, self.end = get_pp_indices(cfg["num_hidden_layers"], self.pp_rank, self.pp_size)
        self.gemma = cfg.get("model_type", "").startswith("gemma")
        self.embed_scale = cfg["hidden_size"] ** 0.5 if self.gemma else 1.0
        if self.pp_rank == 0 or (cfg.get("tie_word_embeddings") and self.pp_rank == self.pp_size - 1):
            rep = (not cfg.get("tie_word_embeddings")
                   and cfg["vocab_size"] * cfg["hidden_size"] * 2 <= (4 << 30))
            self.embed_tokens = VocabParallelEmbedding(cfg["vocab_size"], cfg["hidden_size"], dtype,
                                                       prefix=f"{prefix}.embed_tokens", replicate=rep)
        else:
            self.embed_tokens = None
        self.layers = nn.ModuleList(
            [self._make_layer(cfg, mc, dtype, quant_config, f"{prefix}.layers.{i}", i)
             for i in range(self.start, self.end)])
        self.norm = RMSNorm(cfg["hidden_size"], cfg.get("rms_norm_eps", 1e-6), dtype, self.gemma) \
            if self.pp_rank == self.pp_size - 1 else None

    def _make_layer(self, cfg, mc, dtype, quant_config, prefix, idx):
        return LlamaDecoderLayer(cfg, mc, dtype, quant_config, prefix, idx, self.gemma)

    @property
    def _layer_ahead(self):
        from vllm_b200 import envs
        mb = envs.VLLM_B200_LAYER_PREFETCH_MB
        return mb if mb > 0 else None

    def _prefetch_layer(self, i: int) -> None:
        """Start pulling layer ``i``'s weight matrices into L2 (side stream, csrc/cache_ops.cu::l2_prefetch_kernel) in the
        order its kernels will read them, up to the configured budget."""
        from vllm_b200 import envs
        from vllm_b200.ops import cuda_ops
        cu = cuda_ops()
        budget = envs.VLLM_B200_LAYER_PREFETCH_MB << 20
        ws = self._layer_weights.get(i) if hasattr(self, "_layer_weights") else None
        if ws is None:
            if not hasattr(self, "_layer_weights"):
                self._layer_weights = {}
            ws = self._layer_weights[i] = [p for _, p in self.layers[i].named_parameters() if p.dim() >= 2 and p.is_cuda]
        for w in ws:
            nbytes = w.numel() * w.element_size()
            if nbytes > budget:
                break
            cu.l2_prefetch_async(w)
            budget -= nbytes

    @property
    def _block_skips(self):
        got = self.__dict__.get("_block_skips_cache", 0)
        if got != 0:
            return got
        c = self.cfg
        out = None
        if c.get("bskcn_1") or c.get("bskcn_2"):
            tv = c.get("bskcn_tv") or [0.0, 0.0]
            out = {"save": [set(c.get("bskcn_1") or ()), set(c.get("bskcn_2") or ())],
                   "mix": [set(c.get("bskcn_3") or ()), set(c.get("bskcn_4") or ())], "tv": float(tv[1] if len(tv) > 1 else tv[0])}
        self.__dict__["_block_skips_cache"] = out
        return out

    def embed_input_ids(self, input_ids: torch.Tensor) -> torch.Tensor:
        h = self.embed_tokens(input_ids)
        return h * self.embed_scale if self.embed_scale != 1.0 else h

    def forward(self, input_ids, positions, intermediate_tensors=None, inputs_embeds=None):
        if self.pp_rank == 0:
            stream = inputs_embeds if inputs_embeds is not None else self.embed_input_ids(input_ids)
            residual = stream
            hidden = self.layers[0].input_layernorm(stream) if len(self.layers) else stream
        else:
            hidden, residual = self.layers[0].input_layernorm(intermediate_tensors["hidden_states"],
                                                              intermediate_tensors["residual"])
        n = len(self.layers)
        from vllm_b200.forward_context import get_forward_context_or_none
        fctx = get_forward_context_or_none()
        taps = fctx.extra.get("aux_layers") if fctx is not None else None     # EAGLE-3: residual streams of a few layers
        aux = [] if taps else None
        deep = fctx.extra.get("deepstack") if fctx is not None else None
        H = residual.shape[-1]
        ahead = self._layer_ahead if hidden.is_cuda and hidden.shape[0] <= 256 else None
        skip = self._block_skips
        saved: dict[int, torch.Tensor] = {}
        for i, layer in enumerate(self.layers):
            if skip is not None:
                # Solar "block skip connections": the stream entering a few early layers is blended (weight tv) into the stream
                # entering later ones; the fused add-norm already normalised the un-blended stream → redo this layer's input norm
                li = self.start + i
                for k, src in enumerate(skip["save"]):
                    if li in src:
                        saved[k] = residual
                for k, dst in enumerate(skip["mix"]):
                    if li in dst and k in saved:
                        residual = saved[k] * skip["tv"] + residual * (1.0 - skip["tv"])
                        hidden = layer.input_layernorm(residual)
            nxt = self.layers[i + 1].input_layernorm if i + 1 < n else self.norm
            if ahead is not None and i + 1 < n:
                self._prefetch_layer(i + 1)
            hidden, residual = layer(positions, hidden, residual, nxt)
            if deep is not None and (self.start + i) * H < deep.shape[1]:
                # DeepStack (Qwen3-VL): the i-th level of vision features is added to the stream after decoder layer i (zero
                # rows at text positions); the fused add-norm already produced ``hidden`` → redo the norm on the new stream
                residual = residual + deep[:, (self.start + i) * H:(self.start + i + 1) * H].to(residual.dtype)
                hidden = nxt(residual)
            if aux is not None and (self.start + i) in taps:
                aux.append(residual)
        if aux:
            fctx.extra["aux_hidden"] = torch.cat(aux, dim=-1)
        if hidden.is_cuda:
            from vllm_b200.ops import cuda_ops
            cuda_ops().l2_prefetch_join()
        if self.norm is None:  # not the last PP stage: ship the un-normalised pair
            return {"hidden_states": hidden, "residual": residual}
        if get_tp_group().world_size > 1 and hidden.is_cuda:
            # under TP the fused GEMM+all-reduce+norm returns a view of the symmetric heap, which later
            # collectives (logits all-gather, prompt-logprob passes) recycle: detach the final states from it
            hidden = hidden.clone()
        return hidden


class LlamaForCausalLM(nn.Module):
    packed_modules_mapping = {"qkv_proj": ["q_proj", "k_proj", "v_proj"], "gate_up_proj": ["gate_proj", "up_proj"]}
    model_cls = LlamaModel

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__()
        mc = vllm_config.model_config
        cfg = mc.hf_config
        self.cfg = cfg
        from vllm_b200.layers.quant import get_quant_config
        self.quant_config = get_quant_config(mc)
        self.model = self.model_cls(vllm_config, f"{prefix}model", self.quant_config)
        self.is_last = self.model.norm is not None
        if self.is_last:
            self.lm_head = ParallelLMHead(cfg["vocab_size"], cfg["hidden_size"], dtype=mc.torch_dtype,
                                          prefix="lm_head")
            if cfg.get("tie_word_embeddings"):
                self.lm_head.tie_weights(self.model.embed_tokens)
            self.logits_processor = LogitsProcessor(cfg["vocab_size"], soft_cap=cfg.get("final_logit_softcapping"))

    def forward(self, input_ids, positions, intermediate_tensors=None, inputs_embeds=None):
        return self.model(input_ids, positions, intermediate_tensors, inputs_embeds)

    def embed_input_ids(self, input_ids):
        return self.model.embed_input_ids(input_ids)

    def compute_logits(self, hidden: torch.Tensor) -> torch.Tensor:
        return self.logits_processor(self.lm_head, hidden)

    def sample_candidates(self, hidden, temperature, seeds, steps, out) -> bool:
        """Decode-graph fast path under TP: sample per vocab shard, gather one candidate per row (layers/common.py)."""
        return self.logits_processor.sample_candidates(self.lm_head, hidden, temperature, seeds, steps, out)

    # ------------------------------------------------------------ weights
    stacked = [("qkv_proj", "q_proj", "q"), ("qkv_proj", "k_proj", "k"), ("qkv_proj", "v_proj", "v"),
               ("gate_up_proj", "gate_proj", 0), ("gate_up_proj", "up_proj", 1)]

    def _localize(self, name: str) -> Optional[str]:
        """Checkpoint names carry global layer indices; this PP stage holds layers [start, end) as 0..n-1."""
        i = name.find(".layers.")
        if i < 0:
            return name
        j = i + len(".layers.")
        k = name.find(".", j)
        g = int(name[j:k])
        if not (self.model.start <= g < self.model.end):
            return None
        return name if self.model.start == 0 else f"{name[:j]}{g - self.model.start}{name[k:]}"

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        params = dict(self.named_parameters())
        modules = dict(self.named_modules())
        loaded: set[str] = set()
        for name, w in weights:
            if "rotary_emb.inv_freq" in name:
                continue
            name = self._localize(name)
            if name is None:
                continue
            if name == "lm_head.weight" and self.cfg.get("tie_word_embeddings"):
                continue
            target, shard = name, None
            for fused, part, sid in self.stacked:
                if f".{part}." in name:
                    target, shard = name.replace(part, fused), sid
                    break
            if target not in params:
                continue  # other PP stage / unused
            p = params[target]
            mod = modules[target.rsplit(".", 1)[0]]
            loader = getattr(mod, "weight_loader", None)
            if loader is not None:
                loader(p, w, shard)
            else:
                p.data.copy_(w)
            loaded.add(target)
        return loaded


class MistralForCausalLM(LlamaForCausalLM):
    pass


class Qwen2ForCausalLM(LlamaForCausalLM):
    """Qwen2: q/k/v bias, no o_proj bias."""

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        vllm_config.model_config.hf_config.setdefault("attention_bias", True)
        super().__init__(vllm_config, prefix)


class Qwen3ForCausalLM(LlamaForCausalLM):
    """Qwen3: per-head q/k RMSNorm before RoPE."""

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        vllm_config.model_config.hf_config["qk_norm"] = True
        super().__init__(vllm_config, prefix)


class GemmaForCausalLM(LlamaForCausalLM):
    """Gemma: (1+w) RMSNorm, sqrt(H)-scaled embeddings, GeGLU, tied head."""


class Phi3ForCausalLM(LlamaForCausalLM):
    """Phi-3 / Phi-3.5: Llama block with q|k|v and gate|up already fused in the checkpoint (``qkv_proj``,
    ``gate_up_proj``) and LongRoPE scaling (reference: models/phi3.py)."""
