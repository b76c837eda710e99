"""LFM2 / LFM2-MoE (Liquid): a stack of gated short-convolution blocks with a GQA attention block every few layers.

The short-conv operator is ``out_proj(C ⊙ conv_k(B ⊙ x))`` with a depthwise causal convolution of width ``conv_L_cache``
(3): its recurrent state is just the last k-1 gated inputs per channel, held in the same per-request state pool the SSM
mixers use (``MambaSpec``), so chunked prefill, prefix resume and CUDA-graph decode work as for Mamba.  On CUDA the
convolution is the varlen ``csrc/ssm.cu`` kernel (no activation); elsewhere a vectorised window product.
Reference parity: vllm/model_executor/models/lfm2.py, lfm2_moe.py, layers/mamba/short_conv.py.
"""
from __future__ import annotations

from typing import Iterable

import torch
import torch.nn as nn

from vllm_b200.config import VllmConfig
from vllm_b200.core.kv_cache_manager import KVCacheSpec, MambaSpec
from vllm_b200.forward_context import get_forward_context
from vllm_b200.layers import attention as _attn
from vllm_b200.layers.attention import Attention
from vllm_b200.layers.common import LogitsProcessor, ParallelLMHead, RMSNorm, VocabParallelEmbedding
from vllm_b200.layers.linear import QKVParallelLinear, ReplicatedLinear, RowParallelLinear
from vllm_b200.layers.rotary import get_rope
from vllm_b200.models.llama import LlamaMLP
from vllm_b200.parallel.state import get_tp_group


class ShortConv(nn.Module):
    def __init__(self, cfg: dict, dtype, prefix: str) -> None:
        super().__init__()
        H, self.k = cfg["hidden_size"], cfg.get("conv_L_cache", 3)
        bias = bool(cfg.get("conv_bias", False))
        self.H = H
        self.in_proj = ReplicatedLinear(H, 3 * H, bias, dtype, None, f"{prefix}.in_proj")
        self.out_proj = ReplicatedLinear(H, H, bias, dtype, None, f"{prefix}.out_proj")
        self.conv_weight = nn.Parameter(torch.empty(H, self.k, dtype=dtype), requires_grad=False)
        self.conv_bias = nn.Parameter(torch.zeros(H, dtype=dtype), requires_grad=False) if bias else None
        self.layer_name = prefix
        self.state_elems = H * (self.k - 1)
        self.num_kv_heads, self.head_size, self.sliding_window = 1, 0, None
        assert prefix not in _attn._LAYER_REGISTRY, f"duplicate layer name {prefix}"
        _attn._LAYER_REGISTRY[prefix] = self

    def get_kv_cache_spec(self, block_size: int, dtype_size: int) -> KVCacheSpec:
        return MambaSpec(block_size=block_size, state_bytes=self.state_elems * 4)

    def _conv(self, win: torch.Tensor) -> torch.Tensor:
        """win [..., H, k] (oldest → newest) → [..., H] fp32."""
        y = (win.float() * self.conv_weight.float()).sum(-1)
        return y + self.conv_bias.float() if self.conv_bias is not None else y

    def _segment(self, bx: torch.Tensor, state: torch.Tensor):
        xin = torch.cat([state.t().to(bx.dtype), bx], dim=0)                  # [k-1+n, H]
        return self._conv(xin.float().unfold(0, self.k, 1)), xin[-(self.k - 1):].t().float()

    def forward(self, hidden: torch.Tensor) -> torch.Tensor:
        ctx = get_forward_context()
        md = ctx.metadata_for(self.layer_name)
        cache = ctx.kv_caches.get(self.layer_name)
        B, C, x = self.in_proj(hidden).chunk(3, dim=-1)
        bx = B * x
        T, dev = hidden.shape[0], hidden.device
        if md is None or cache is None:                      # profiling run: one stateless sequence
            y, _ = self._segment(bx, torch.zeros(self.H, self.k - 1, device=dev))
            return self.out_proj((C.float() * y).to(hidden.dtype))
        from vllm_b200.layers.mamba import _cuda_ssm
        if _cuda_ssm(bx) and 2 <= self.k <= 4:
            from vllm_b200.ops import cuda_ops
            y = cuda_ops().causal_conv1d(bx.contiguous(), self.conv_weight, self.conv_bias, cache, md, act=False)
            return self.out_proj((C.float() * y).to(hidden.dtype))
        cache = cache.view(cache.shape[0], -1).view(torch.float32) if cache.dtype != torch.float32 else cache
        out = torch.zeros(T, self.H, dtype=torch.float32, device=dev)
        if md.decode_only:
            n = md.num_reqs
            slots = md.block_table[:n, 0].long()
            st = cache[slots][:, :self.state_elems].view(n, self.H, self.k - 1)
            win = torch.cat([st, bx[:n].float()[:, :, None]], dim=-1)
            out[:n] = self._conv(win)
            cache[slots, :self.state_elems] = win[:, :, 1:].reshape(n, -1)
        else:
            qsl = (md.query_start_loc_cpu if md.query_start_loc_cpu is not None else md.query_start_loc.cpu()).tolist()
            sls = (md.seq_lens_cpu if md.seq_lens_cpu is not None else md.seq_lens.cpu()).tolist()
            slots = md.block_table[:md.num_reqs, 0].tolist()
            for i in range(md.num_reqs):
                s, e = qsl[i], qsl[i + 1]
                if e <= s:
                    continue
                if sls[i] - (e - s) > 0:
                    st = cache[slots[i]][:self.state_elems].view(self.H, self.k - 1).clone()
                else:
                    st = torch.zeros(self.H, self.k - 1, dtype=torch.float32, device=dev)
                y, st = self._segment(bx[s:e], st)
                cache[slots[i]][:self.state_elems] = st.reshape(-1)
                out[s:e] = y
        return self.out_proj((C.float() * out).to(hidden.dtype))


class Lfm2Attention(nn.Module):
    def __init__(self, cfg: dict, mc, dtype, prefix: str) -> None:
        super().__init__()
        H, nh = cfg["hidden_size"], cfg["num_attention_heads"]
        nkv = cfg.get("num_key_value_heads") or nh
        hd = cfg.get("head_dim") or H // nh
        eps = cfg.get("norm_eps", 1e-5)
        self.qkv_proj = QKVParallelLinear(H, hd, nh, nkv, False, dtype, None, f"{prefix}.qkv_proj")
        self.out_proj = RowParallelLinear(nh * hd, H, False, dtype=dtype, prefix=f"{prefix}.out_proj")
        self.sizes = [self.qkv_proj.q_size, self.qkv_proj.kv_size, self.qkv_proj.kv_size]
        self.q_layernorm, self.k_layernorm = RMSNorm(hd, eps, dtype), RMSNorm(hd, eps, dtype)
        rp = cfg.get("rope_parameters") or {}
        max_pos = mc.max_model_len if mc.max_model_len > 0 else cfg.get("max_position_embeddings", 8192)
        scaling = ({k: v for k, v in rp.items() if k != "rope_theta"} if rp.get("rope_type", "default") != "default"
                   else cfg.get("rope_scaling"))
        rope = get_rope(hd, hd, max_pos, rp.get("rope_theta", cfg.get("rope_theta", 1000000.0)), True, scaling)
        self.attn = Attention(self.qkv_proj.num_heads, hd, hd ** -0.5, self.qkv_proj.num_kv_heads, prefix=f"{prefix}.attn", rope=rope)

    def forward(self, positions, x):
        q, k, v = self.qkv_proj(x).split(self.sizes, dim=-1)
        return self.out_proj(self.attn(q, k, v, positions, qk_norm=(self.q_layernorm, self.k_layernorm)))


def _ffn_width(cfg: dict) -> int:
    I = cfg["intermediate_size"]
    if cfg.get("block_auto_adjust_ff_dim", cfg.get("model_type") == "lfm2"):
        I = int(2 * I / 3)
        m = cfg.get("block_ffn_dim_multiplier")
        if m is not None:
            I = int(m * I)
            mult = cfg.get("block_multiple_of", 256)
            I = mult * ((I + mult - 1) // mult)
    return I


class Lfm2Layer(nn.Module):
    def __init__(self, cfg: dict, mc, dtype, prefix: str, idx: int, enable_ep: bool) -> None:
        super().__init__()
        H, eps = cfg["hidden_size"], cfg.get("norm_eps", 1e-5)
        types = cfg.get("layer_types")
        full = set(cfg.get("full_attn_idxs") or [])
        self.is_attn = (types[idx] == "full_attention") if types else (idx in full)
        if self.is_attn:
            self.self_attn = Lfm2Attention(cfg, mc, dtype, f"{prefix}.self_attn")
        else:
            self.conv = ShortConv(cfg, dtype, f"{prefix}.conv")
        self.sparse = bool(cfg.get("num_experts")) and idx >= cfg.get("num_dense_layers", 0)
        if self.sparse:
            from vllm_b200.models.mixtral import SparseMoeBlock
            kw = dict(use_grouped_topk=True, n_group=1, topk_group=1, scoring_func="sigmoid",
                      e_score_bias=bool(cfg.get("use_expert_bias", True)), routed_scaling_factor=float(cfg.get("routed_scaling_factor", 1.0)))
            self.feed_forward = SparseMoeBlock(cfg, dtype, None, f"{prefix}.feed_forward", enable_ep, n_experts=cfg["num_experts"],
                                               top_k=cfg["num_experts_per_tok"], inter=cfg["moe_intermediate_size"],
                                               renormalize=bool(cfg.get("norm_topk_prob", True)), moe_kw=kw)
        else:
            self.feed_forward = LlamaMLP(H, _ffn_width(cfg), "silu", False, dtype, None, f"{prefix}.feed_forward")
        self.operator_norm = RMSNorm(H, eps, dtype)
        self.ffn_norm = RMSNorm(H, eps, dtype)

    def forward(self, positions, x):
        h = self.operator_norm(x)
        x = x + (self.self_attn(positions, h) if self.is_attn else self.conv(h))
        m = self.feed_forward(self.ffn_norm(x))
        if self.sparse and get_tp_group().world_size > 1:
            m = get_tp_group().all_reduce(m)
        return x + m


class Lfm2ForCausalLM(nn.Module):
    stacked = [("qkv_proj", "q_proj", "q"), ("qkv_proj", "k_proj", "k"), ("qkv_proj", "v_proj", "v"),
               ("gate_up_proj", "gate_proj", 0), ("gate_up_proj", "up_proj", 1)]
    _expert = {"w1": "w1", "w3": "w3", "w2": "w2", "gate_proj": "w1", "up_proj": "w3", "down_proj": "w2"}

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        super().__init__()
        assert get_tp_group().world_size == 1, "short-conv hybrids run replicated (tensor_parallel_size=1)"
        mc = vllm_config.model_config
        cfg, dtype = mc.hf_config, mc.torch_dtype
        self.cfg = cfg
        H = cfg["hidden_size"]
        ep = bool(vllm_config.parallel_config.enable_expert_parallel)
        self.embed_tokens = VocabParallelEmbedding(cfg["vocab_size"], H, dtype)
        self.layers = nn.ModuleList([Lfm2Layer(cfg, mc, dtype, f"{prefix}model.layers.{i}", i, ep)
                                     for i in range(cfg["num_hidden_layers"])])
        self.embedding_norm = RMSNorm(H, cfg.get("norm_eps", 1e-5), dtype)
        self.lm_head = ParallelLMHead(cfg["vocab_size"], H, dtype=dtype)
        if cfg.get("tie_word_embeddings", True):
            self.lm_head.tie_weights(self.embed_tokens)
        self.logits_processor = LogitsProcessor(cfg["vocab_size"])

    def embed_input_ids(self, ids):
        return self.embed_tokens(ids)

    def forward(self, input_ids, positions, intermediate_tensors=None, inputs_embeds=None):
        x = inputs_embeds if inputs_embeds is not None else self.embed_tokens(input_ids)
        for layer in self.layers:
            x = layer(positions, x)
        return self.embedding_norm(x)

    def compute_logits(self, hidden):
        return self.logits_processor(self.lm_head, hidden)

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        params = dict(self.named_parameters())
        modules = dict(self.named_modules())
        loaded: set[str] = set()

        def expand():
            for name, w in weights:
                if name.endswith(".experts.gate_up_proj"):          # fused [E, 2I, H] slabs (in-memory HF layout)
                    I = w.shape[1] // 2
                    for e in range(w.shape[0]):
                        yield f"{name[:-len('gate_up_proj')]}{e}.w1.weight", w[e, :I]
                        yield f"{name[:-len('gate_up_proj')]}{e}.w3.weight", w[e, I:]
                elif name.endswith(".experts.down_proj"):
                    for e in range(w.shape[0]):
                        yield f"{name[:-len('down_proj')]}{e}.w2.weight", w[e]
                else:
                    yield name, w
        for name, w in expand():
            n = name[len("model."):] if name.startswith("model.") else name
            if n == "lm_head.weight" and self.cfg.get("tie_word_embeddings", True):
                continue
            if ".feed_forward.experts." in n:
                head, tail = n.split(".feed_forward.experts.", 1)
                parts = tail.split(".")
                t = f"{head}.feed_forward.experts." + ("w2_weight" if self._expert[parts[1]] == "w2" else "w13_weight")
                modules[f"{head}.feed_forward.experts"].weight_loader(params[t], w, self._expert[parts[1]], int(parts[0]))
                loaded.add(t)
                continue
            n = (n.replace(".conv.conv.weight", ".conv.conv_weight").replace(".conv.conv.bias", ".conv.conv_bias")
                 .replace(".feed_forward.expert_bias", ".feed_forward.experts.e_score_correction_bias"))
            if ".feed_forward.w" in n:
                n = n.replace(".w1.", ".gate_proj.").replace(".w3.", ".up_proj.").replace(".w2.", ".down_proj.")
            target, shard = n, None
            for fused, part, sid in self.stacked:
                if f".{part}." in n:
                    target, shard = n.replace(part, fused), sid
                    break
            if target not in params:
                continue
            p = params[target]
            if target.endswith("conv_weight"):
                p.data.copy_(w.reshape(p.shape))
            else:
                mod = modules[target.rsplit(".", 1)[0]]
                loader = getattr(mod, "weight_loader", None)
                loader(p, w, shard) if loader is not None else p.data.copy_(w)
            loaded.add(target)
        return loaded


class Lfm2MoeForCausalLM(Lfm2ForCausalLM):
    pass
