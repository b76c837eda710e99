"""Qwen3-Next: hybrid of gated-delta-net (linear attention with a recurrent [d_k x d_v] state per head) and gated
full attention (every 4th layer), zero-centred RMSNorm, MoE with a gated shared expert in every layer.

Reference parity: vllm/model_executor/models/qwen3_next.py, layers/mamba/ (GDN mixer, ``fused_gdn_decode`` /
``chunk_gated_delta_rule`` kernels, causal conv1d), csrc/gdn/fused_gdn_decode_kernel.cu.

State cache: one pool block per request and GDN layer holding ``[conv_state (conv_dim x (k-1)) | S (heads x d_k x d_v)]``
in fp32 (same ``MambaSpec`` mechanism as the SSM mixers).  Decode batches are one vectorised rank-1 state update per
head (graph-capturable); prefill walks the segment token by token through the same update (the chunked WY form is the
planned tcgen05 kernel — the recurrence is exact and is what the decode kernel computes anyway).
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from vllm_b200.config import VllmConfig
from vllm_b200.core.kv_cache_manager import KVCacheSpec, MambaSpec
from vllm_b200.forward_context import get_forward_context
from vllm_b200.layers import attention as _attn
from vllm_b200.layers.attention import Attention
from vllm_b200.layers.common import RMSNorm
from vllm_b200.layers.linear import ReplicatedLinear
from vllm_b200.layers.rotary import get_rope
from vllm_b200.models.llama import LlamaMLP
from vllm_b200.models.mixtral import Qwen2MoeForCausalLM, Qwen2MoeModel, SparseMoeBlock
from vllm_b200.parallel.state import get_tp_group


def _l2norm(x: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    return x * torch.rsqrt((x * x).sum(-1, keepdim=True) + eps)


class GatedDeltaNet(nn.Module):
    def __init__(self, cfg: dict, dtype, prefix: str) -> None:
        super().__init__()
        H = cfg["hidden_size"]
        self.nv, self.nk = cfg["linear_num_value_heads"], cfg["linear_num_key_heads"]
        self.dk, self.dv = cfg["linear_key_head_dim"], cfg["linear_value_head_dim"]
        self.key_dim, self.value_dim = self.dk * self.nk, self.dv * self.nv
        self.k = cfg.get("linear_conv_kernel_dim", 4)
        self.conv_dim = 2 * self.key_dim + self.value_dim
        self.eps = cfg.get("rms_norm_eps", 1e-6)
        self.beta_scale = 1.0                 # OLMo-hybrid lets the write strength reach 2 (negative eigenvalues of the update)
        self.in_proj_qkvz = ReplicatedLinear(H, 2 * self.key_dim + 2 * self.value_dim, False, dtype, None, f"{prefix}.in_proj_qkvz")
        self.in_proj_ba = ReplicatedLinear(H, 2 * self.nv, False, dtype, None, f"{prefix}.in_proj_ba")
        self.conv_weight = nn.Parameter(torch.empty(self.conv_dim, self.k, dtype=dtype), requires_grad=False)
        self.dt_bias = nn.Parameter(torch.ones(self.nv, dtype=torch.float32), requires_grad=False)
        self.A_log = nn.Parameter(torch.zeros(self.nv, dtype=torch.float32), requires_grad=False)
        self.norm_weight = nn.Parameter(torch.ones(self.dv, dtype=dtype), requires_grad=False)
        self.out_proj = ReplicatedLinear(self.value_dim, H, False, dtype, None, f"{prefix}.out_proj")
        self.layer_name = prefix
        self.state_elems = self.conv_dim * (self.k - 1) + self.nv * self.dk * self.dv
        self.num_kv_heads, self.head_size, self.sliding_window = 1, 0, None
        assert prefix not in _attn._LAYER_REGISTRY, f"duplicate layer name {prefix}"
        _attn._LAYER_REGISTRY[prefix] = self

    def get_kv_cache_spec(self, block_size: int, dtype_size: int) -> KVCacheSpec:
        return MambaSpec(block_size=block_size, state_bytes=self.state_elems * 4)

    # checkpoint layout: per key-head groups [q | k | v (r heads) | z (r heads)] and [b (r) | a (r)], r = nv / nk
    def _split(self, qkvz: torch.Tensor, ba: torch.Tensor):
        T, r = qkvz.shape[0], self.nv // self.nk
        g = qkvz.view(T, self.nk, 2 * self.dk + 2 * r * self.dv)
        q, k, v, z = g.split([self.dk, self.dk, r * self.dv, r * self.dv], dim=-1)
        b, a = ba.view(T, self.nk, 2 * r).split([r, r], dim=-1)
        return (q.reshape(T, self.key_dim), k.reshape(T, self.key_dim), v.reshape(T, self.value_dim),
                z.reshape(T, self.nv, self.dv), b.reshape(T, self.nv), a.reshape(T, self.nv))

    def _split_state(self, st: torch.Tensor):
        nconv = self.conv_dim * (self.k - 1)
        return (st[..., :nconv].reshape(*st.shape[:-1], self.conv_dim, self.k - 1),
                st[..., nconv:].reshape(*st.shape[:-1], self.nv, self.dk, self.dv))

    def _qkv_heads(self, conv: torch.Tensor):
        """conv output [n, conv_dim] (post-SiLU) → q, k [n, nv, dk] (l2-normalised, q scaled), v [n, nv, dv]."""
        n, r = conv.shape[0], self.nv // self.nk
        q, k, v = conv.split([self.key_dim, self.key_dim, self.value_dim], dim=-1)
        q = _l2norm(q.view(n, self.nk, self.dk)).repeat_interleave(r, dim=1) * self.dk ** -0.5
        k = _l2norm(k.view(n, self.nk, self.dk)).repeat_interleave(r, dim=1)
        return q, k, v.view(n, self.nv, self.dv)

    @staticmethod
    def _step(S, q, k, v, g, beta):
        """One token for a batch of states.  S [B, H, dk, dv]; q, k [B, H, dk]; v [B, H, dv]; g, beta [B, H]."""
        S = S * g.exp()[..., None, None]
        delta = (v - (S * k[..., None]).sum(-2)) * beta[..., None]
        S = S + k[..., None] * delta[..., None, :]
        return S, (S * q[..., None]).sum(-2)

    def _gated_norm(self, y: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
        y = y.float()
        y = y * torch.rsqrt(y.pow(2).mean(-1, keepdim=True) + self.eps)
        return self.norm_weight.float() * y * F.silu(z.float())

    def _segment(self, qkv, b, a, conv_state, S):
        n = qkv.shape[0]
        xin = torch.cat([conv_state.t().to(qkv.dtype), qkv], dim=0)
        conv = F.silu((xin.float().unfold(0, self.k, 1) * self.conv_weight.float()[None]).sum(-1))
        q, k, v = self._qkv_heads(conv)
        beta = b.float().sigmoid() * self.beta_scale
        g = -self.A_log.exp() * F.softplus(a.float() + self.dt_bias)
        out = torch.empty(n, self.nv, self.dv, dtype=torch.float32, device=qkv.device)
        S = S[None]
        for t in range(n):
            S, o = self._step(S, q[t:t + 1], k[t:t + 1], v[t:t + 1], g[t:t + 1], beta[t:t + 1])
            out[t] = o[0]
        return out, xin[-(self.k - 1):].t().float(), S[0]

    def forward(self, hidden: torch.Tensor) -> torch.Tensor:
        ctx = get_forward_context()
        md = ctx.metadata_for(self.layer_name)
        cache = ctx.kv_caches.get(self.layer_name)
        q, k, v, z, b, a = self._split(self.in_proj_qkvz(hidden), self.in_proj_ba(hidden))
        qkv = torch.cat([q, k, v], dim=-1)
        T, dev = hidden.shape[0], hidden.device
        out = torch.zeros(T, self.nv, self.dv, dtype=torch.float32, device=dev)
        if md is None or cache is None:
            out, _, _ = self._segment(qkv[:min(T, 8)], b[:min(T, 8)], a[:min(T, 8)],
                                      torch.zeros(self.conv_dim, self.k - 1, device=dev),
                                      torch.zeros(self.nv, self.dk, self.dv, device=dev))
            out = torch.cat([out, torch.zeros(T - out.shape[0], self.nv, self.dv, device=dev)]) if T > out.shape[0] else out
            return self.out_proj(self._gated_norm(out, z).reshape(T, -1).to(hidden.dtype))
        from vllm_b200.layers.mamba import _cuda_ssm
        if _cuda_ssm(qkv) and self.dk % 32 == 0 and self.dk <= 256 and 2 <= self.k <= 4:
            # device path (csrc/ssm.cu): causal conv + gated-delta-rule scan over the cache pool, any batch shape
            from vllm_b200.ops import cuda_ops
            cu = cuda_ops()
            conv = cu.causal_conv1d(qkv, self.conv_weight, None, cache, md)
            qh, kh, vh = self._qkv_heads(conv)
            g = -self.A_log.exp() * F.softplus(a.float() + self.dt_bias)
            o = cu.gdn_scan(qh, kh, vh, g, b.float().sigmoid() * self.beta_scale, cache, self.conv_dim * (self.k - 1), md)
            return self.out_proj(self._gated_norm(o, z).reshape(T, -1).to(hidden.dtype))
        cache = cache.view(cache.shape[0], -1).view(torch.float32) if cache.dtype != torch.float32 else cache
        if md.decode_only:
            Bn = md.num_reqs
            slots = md.block_table[:Bn, 0].long()
            conv_state, S = self._split_state(cache[slots])
            xin = torch.cat([conv_state, qkv[:Bn].float()[:, :, None]], dim=-1)
            conv = F.silu((xin * self.conv_weight.float()[None]).sum(-1))
            qh, kh, vh = self._qkv_heads(conv)
            g = -self.A_log.exp() * F.softplus(a[:Bn].float() + self.dt_bias)
            S, o = self._step(S, qh, kh, vh, g, b[:Bn].float().sigmoid() * self.beta_scale)
            cache[slots] = torch.cat([xin[:, :, 1:].reshape(Bn, -1), S.reshape(Bn, -1)], dim=-1)
            out[:Bn] = o
        else:
            qsl = (md.query_start_loc_cpu if md.query_start_loc_cpu is not None else md.query_start_loc.cpu()).tolist()
            sls = (md.seq_lens_cpu if md.seq_lens_cpu is not None else md.seq_lens.cpu()).tolist()
            slots = md.block_table[:md.num_reqs, 0].tolist()
            for i in range(md.num_reqs):
                s, e = qsl[i], qsl[i + 1]
                if e <= s:
                    continue
                if sls[i] - (e - s) > 0:
                    conv_state, S = self._split_state(cache[slots[i]])
                    conv_state, S = conv_state.clone(), S.clone()
                else:
                    conv_state = torch.zeros(self.conv_dim, self.k - 1, dtype=torch.float32, device=dev)
                    S = torch.zeros(self.nv, self.dk, self.dv, dtype=torch.float32, device=dev)
                o, conv_state, S = self._segment(qkv[s:e], b[s:e], a[s:e], conv_state, S)
                cache[slots[i]] = torch.cat([conv_state.reshape(-1), S.reshape(-1)])
                out[s:e] = o
        return self.out_proj(self._gated_norm(out, z).reshape(T, -1).to(hidden.dtype))


class GatedAttention(nn.Module):
    """Full attention with per-head zero-centred q/k RMSNorm, partial rotary and a sigmoid output gate that is
    produced by the query projection (``q_proj`` emits [query | gate] per head)."""

    def __init__(self, cfg: dict, mc, dtype, prefix: str) -> None:
        super().__init__()
        H, nh = cfg["hidden_size"], cfg["num_attention_heads"]
        self.nh, self.nkv = nh, cfg.get("num_key_value_heads") or nh
        self.hd = cfg.get("head_dim") or H // nh
        bias = bool(cfg.get("attention_bias", False))
        self.q_proj = ReplicatedLinear(H, nh * self.hd * 2, bias, dtype, None, f"{prefix}.q_proj")
        self.k_proj = ReplicatedLinear(H, self.nkv * self.hd, bias, dtype, None, f"{prefix}.k_proj")
        self.v_proj = ReplicatedLinear(H, self.nkv * self.hd, bias, dtype, None, f"{prefix}.v_proj")
        self.o_proj = ReplicatedLinear(nh * self.hd, H, bias, dtype, None, f"{prefix}.o_proj")
        eps = cfg.get("rms_norm_eps", 1e-6)
        self.q_norm, self.k_norm = RMSNorm(self.hd, eps, dtype, True), RMSNorm(self.hd, eps, dtype, True)
        rp = cfg.get("rope_parameters") or {}
        factor = rp.get("partial_rotary_factor", cfg.get("partial_rotary_factor", 1.0))
        max_pos = mc.max_model_len if mc.max_model_len > 0 else cfg.get("max_position_embeddings", 32768)
        rope = get_rope(self.hd, int(self.hd * factor), max_pos, rp.get("rope_theta", cfg.get("rope_theta", 10000.0)), True,
                        cfg.get("rope_scaling"))
        self.attn = Attention(nh, self.hd, self.hd ** -0.5, self.nkv, prefix=f"{prefix}.attn", rope=rope)

    def forward(self, positions, hidden):
        T = hidden.shape[0]
        qg = self.q_proj(hidden).view(T, self.nh, 2 * self.hd)
        q, gate = qg[..., :self.hd], qg[..., self.hd:].reshape(T, -1)
        q = self.q_norm(q.reshape(-1, self.hd)).view(T, -1)
        k = self.k_norm(self.k_proj(hidden).view(-1, self.hd)).view(T, -1)
        o = self.attn(q, k, self.v_proj(hidden), positions)
        return self.o_proj(o * torch.sigmoid(gate))


class Qwen3NextDecoderLayer(nn.Module):
    def __init__(self, cfg, mc, dtype, quant_config, prefix: str, idx: int, enable_ep: bool) -> None:
        super().__init__()
        H, eps = cfg["hidden_size"], cfg.get("rms_norm_eps", 1e-6)
        types = cfg.get("layer_types")
        interval = cfg.get("full_attention_interval", 4)
        self.is_attn = (types[idx] == "full_attention") if types else ((idx + 1) % interval == 0)
        if self.is_attn:
            self.self_attn = GatedAttention(cfg, mc, dtype, f"{prefix}.self_attn")
        else:
            self.linear_attn = GatedDeltaNet(cfg, dtype, f"{prefix}.linear_attn")
        sparse = (idx not in (cfg.get("mlp_only_layers") or []) and cfg.get("num_experts", 0) > 0
                  and (idx + 1) % cfg.get("decoder_sparse_step", 1) == 0)
        if sparse:
            self.mlp = SparseMoeBlock(cfg, dtype, quant_config, f"{prefix}.mlp", enable_ep, n_experts=cfg["num_experts"],
                                      top_k=cfg["num_experts_per_tok"], inter=cfg["moe_intermediate_size"],
                                      renormalize=bool(cfg.get("norm_topk_prob", True)),
                                      shared_inter=cfg.get("shared_expert_intermediate_size", 0), shared_gate=True)
        else:
            self.mlp = LlamaMLP(H, cfg["intermediate_size"], cfg.get("hidden_act", "silu"), False, dtype, quant_config,
                                f"{prefix}.mlp")
        self.input_layernorm = RMSNorm(H, eps, dtype, True)
        self.post_attention_layernorm = RMSNorm(H, eps, dtype, True)

    def forward(self, positions, hidden, residual, next_norm: Optional[nn.Module]):
        stream = residual + (self.self_attn(positions, hidden) if self.is_attn else self.linear_attn(hidden))
        stream = stream + self.mlp(self.post_attention_layernorm(stream))
        return (next_norm(stream) if next_norm is not None else stream), stream


class Qwen3NextModel(Qwen2MoeModel):
    def __init__(self, vllm_config: VllmConfig, prefix: str = "model", quant_config=None) -> None:
        assert get_tp_group().world_size == 1, "Qwen3-Next runs replicated (tensor_parallel_size=1) for now"
        super().__init__(vllm_config, prefix, quant_config)
        if self.norm is not None:
            cfg = self.cfg
            self.norm = RMSNorm(cfg["hidden_size"], cfg.get("rms_norm_eps", 1e-6), vllm_config.model_config.torch_dtype, True)

    def _make_layer(self, cfg, mc, dtype, quant_config, prefix, idx):
        return Qwen3NextDecoderLayer(cfg, mc, dtype, quant_config, prefix, idx, self._enable_ep)


class Qwen3NextForCausalLM(Qwen2MoeForCausalLM):
    model_cls = Qwen3NextModel

    def __init__(self, vllm_config: VllmConfig, prefix: str = "") -> None:
        vllm_config.model_config.hf_config.setdefault("attention_bias", False)
        super().__init__(vllm_config, prefix)

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        params = dict(self.named_parameters())
        loaded: set[str] = set()
        rest = []
        for name, w in weights:
            if ".self_attn." in name or ".linear_attn." in name:
                n = self._localize(name)
                if n is None:
                    continue
                n = n.replace("linear_attn.conv1d.weight", "linear_attn.conv_weight").replace(
                    "linear_attn.norm.weight", "linear_attn.norm_weight")
                if n in params:
                    p = params[n]
                    p.data.copy_(w.reshape(p.shape) if n.endswith("conv_weight") else w)
                    loaded.add(n)
                continue
            rest.append((name, w))
        return loaded | super().load_weights(rest)


class Qwen3_5ForCausalLM(Qwen3NextForCausalLM):
    """Qwen3.5 text decoder (dense and MoE): the Qwen3-Next hybrid with the gated-delta-net input projection stored as four
    plain tensors (q|k|v, z, b, a) instead of the per-key-head interleaved pair; the loader re-interleaves them into the
    layout the mixer splits (reference: models/qwen3_5.py).  Text positions only: the three M-RoPE axes coincide."""

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        c = self.cfg
        nk, nv = c["linear_num_key_heads"], c["linear_num_value_heads"]
        dk, dv, r = c["linear_key_head_dim"], c["linear_value_head_dim"], c["linear_num_value_heads"] // c["linear_num_key_heads"]
        pending: dict[str, dict[str, torch.Tensor]] = {}

        def mapped():
            for name, w in weights:
                name = name.replace("model.language_model.", "model.")
                if name.startswith(("model.visual.", "visual.", "mtp.")):
                    continue
                hit = next((k for k in ("in_proj_qkv", "in_proj_z", "in_proj_b", "in_proj_a") if f".linear_attn.{k}.weight" in name), None)
                if hit is None:
                    yield name, w
                    continue
                base = name[: name.index(".linear_attn.")] + ".linear_attn."
                got = pending.setdefault(base, {})
                got[hit] = w
                if len(got) < 4:
                    continue
                q, k, v = got["in_proj_qkv"].split([nk * dk, nk * dk, nv * dv], dim=0)
                z, b, a = got["in_proj_z"], got["in_proj_b"], got["in_proj_a"]
                H = q.shape[1]
                qkvz = torch.cat([q.view(nk, dk, H), k.view(nk, dk, H), v.view(nk, r * dv, H), z.view(nk, r * dv, H)], dim=1)
                ba = torch.cat([b.view(nk, r, H), a.view(nk, r, H)], dim=1)
                del pending[base]
                yield base + "in_proj_qkvz.weight", qkvz.reshape(-1, H)
                yield base + "in_proj_ba.weight", ba.reshape(-1, H)
        out = super().load_weights(mapped())
        assert not pending, f"incomplete gated-delta-net projections for {sorted(pending)}"
        return out


class Qwen3_5MoeForCausalLM(Qwen3_5ForCausalLM):
    pass
