"""Selective state-space (Mamba-1) mixer with a paged recurrent-state cache.

Behavioural reference: vllm/model_executor/layers/mamba/mamba_mixer.py (conv state + SSM state kept per request
in the KV-cache pool through ``MambaSpec``; causal-conv1d + selective-scan ops).

State layout: one pool block per request holds ``[conv_state (d_inner x (k-1)) | ssm_state (d_inner x d_state)]``
in fp32.  Decode batches (one token per request) are fully vectorised over the batch — gather states by block id,
one fused update, scatter back — with no host synchronisation, so they replay inside CUDA graphs.  Prefill /
mixed batches walk each request's segment (chunked prefill resumes from the cached state).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from vllm_b200.core.kv_cache_manager import KVCacheSpec, MambaSpec
from vllm_b200.forward_context import get_forward_context
from vllm_b200.layers import attention as _attn
from vllm_b200.layers.linear import ReplicatedLinear


def _cuda_ssm(x: torch.Tensor) -> bool:
    from vllm_b200 import envs
    return x.is_cuda and x.dtype in (torch.bfloat16, torch.float16) and not envs.VLLM_B200_FORCE_REF_OPS


class MambaMixer(nn.Module):
    def __init__(self, hidden: int, d_inner: int, d_state: int, d_conv: int, dt_rank: int, use_bias: bool,
                 use_conv_bias: bool, dtype, prefix: str, bcdt_norm_eps: Optional[float] = None) -> None:
        super().__init__()
        # Jamba normalises dt / B / C (RMSNorm) before the dt projection
        from vllm_b200.layers.common import RMSNorm
        self.dt_layernorm = RMSNorm(dt_rank, bcdt_norm_eps, dtype) if bcdt_norm_eps is not None else None
        self.b_layernorm = RMSNorm(d_state, bcdt_norm_eps, dtype) if bcdt_norm_eps is not None else None
        self.c_layernorm = RMSNorm(d_state, bcdt_norm_eps, dtype) if bcdt_norm_eps is not None else None
        self.d_inner, self.d_state, self.d_conv, self.dt_rank = d_inner, d_state, d_conv, dt_rank
        self.in_proj = ReplicatedLinear(hidden, 2 * d_inner, use_bias, dtype, None, f"{prefix}.in_proj")
        self.conv_weight = nn.Parameter(torch.empty(d_inner, d_conv, dtype=dtype), requires_grad=False)
        self.conv_bias = nn.Parameter(torch.zeros(d_inner, dtype=dtype), requires_grad=False) if use_conv_bias else None
        self.x_proj = ReplicatedLinear(d_inner, dt_rank + 2 * d_state, False, dtype, None, f"{prefix}.x_proj")
        self.dt_proj = ReplicatedLinear(dt_rank, d_inner, True, dtype, None, f"{prefix}.dt_proj")
        self.A_log = nn.Parameter(torch.empty(d_inner, d_state, dtype=torch.float32), requires_grad=False)
        self.D = nn.Parameter(torch.ones(d_inner, dtype=torch.float32), requires_grad=False)
        self.out_proj = ReplicatedLinear(d_inner, hidden, use_bias, dtype, None, f"{prefix}.out_proj")
        self.layer_name = prefix
        self.state_elems = d_inner * (d_conv - 1) + d_inner * d_state
        # registered next to the attention layers so the engine sizes a cache group for it
        self.num_kv_heads, self.head_size, self.sliding_window = 1, 0, None
        assert prefix not in _attn._LAYER_REGISTRY, f"duplicate layer name {prefix}"
        _attn._LAYER_REGISTRY[prefix] = self

    def get_kv_cache_spec(self, block_size: int, dtype_size: int) -> KVCacheSpec:
        return MambaSpec(block_size=block_size, state_bytes=self.state_elems * 4)

    # ------------------------------------------------------------------ math
    def _ssm_params(self, x_act: torch.Tensor):
        dbc = self.x_proj(x_act)
        dt, B, C = dbc.split([self.dt_rank, self.d_state, self.d_state], dim=-1)
        if self.dt_layernorm is not None:
            dt, B, C = self.dt_layernorm(dt.contiguous()), self.b_layernorm(B.contiguous()), self.c_layernorm(C.contiguous())
        dt = F.softplus(self.dt_proj(dt).float())
        return dt, B.float(), C.float()

    def _split_state(self, st: torch.Tensor):
        nconv = self.d_inner * (self.d_conv - 1)
        return (st[..., :nconv].reshape(*st.shape[:-1], self.d_inner, self.d_conv - 1),
                st[..., nconv:].reshape(*st.shape[:-1], self.d_inner, self.d_state))

    def _segment(self, x: torch.Tensor, conv_state: torch.Tensor, h: torch.Tensor):
        """x [q, d_inner] of ONE request → (y [q, d_inner] fp32, new conv_state, new h)."""
        k = self.d_conv
        xin = torch.cat([conv_state.t().to(x.dtype), x], dim=0)                 # [k-1+q, di]
        w = self.conv_weight.float()
        win = xin.float().unfold(0, k, 1)                                        # [q, di, k]
        conv = (win * w[None]).sum(-1)
        if self.conv_bias is not None:
            conv = conv + self.conv_bias.float()
        xa = F.silu(conv)
        dt, B, C = self._ssm_params(xa.to(x.dtype))
        A = -torch.exp(self.A_log)
        ys = torch.empty_like(xa)
        for t in range(x.shape[0]):
            dA = torch.exp(dt[t][:, None] * A)
            h = dA * h + (dt[t] * xa[t])[:, None] * B[t][None, :]
            ys[t] = (h * C[t][None, :]).sum(-1)
        ys = ys + self.D * xa
        return ys, xin[-(k - 1):].t().float(), h

    def forward(self, hidden: torch.Tensor) -> torch.Tensor:
        ctx = get_forward_context()
        md = ctx.metadata_for(self.layer_name)
        cache = ctx.kv_caches.get(self.layer_name)
        xz = self.in_proj(hidden)
        x, z = xz.split([self.d_inner, self.d_inner], dim=-1)
        T = x.shape[0]
        if md is None or cache is None:          # profiling run: stateless pass over the whole buffer as one sequence
            y, _, _ = self._segment(x, x.new_zeros(self.d_inner, self.d_conv - 1, dtype=torch.float32),
                                    x.new_zeros(self.d_inner, self.d_state, dtype=torch.float32))
            return self.out_proj((y * F.silu(z.float())).to(hidden.dtype))
        if _cuda_ssm(x) and self.d_state in (8, 16, 32) and 2 <= self.d_conv <= 4:
            # device path (csrc/ssm.cu): conv window and SSM state are advanced in place in the cache pool by two varlen
            # kernels — same code for decode, prefill and chunked prefill; no host loop, graph capturable
            from vllm_b200.ops import cuda_ops
            cu = cuda_ops()
            xa = cu.causal_conv1d(x, self.conv_weight, self.conv_bias, cache, md)
            dt, Bm, Cm = self._ssm_params(xa.to(x.dtype))
            y = cu.selective_scan(xa, dt, Bm, Cm, -torch.exp(self.A_log), self.D, cache,
                                  self.d_inner * (self.d_conv - 1), md)
            return self.out_proj((y * F.silu(z.float())).to(hidden.dtype))
        cache = cache.view(cache.shape[0], -1).view(torch.float32) if cache.dtype != torch.float32 else cache
        if md.decode_only:
            B_ = md.num_reqs
            slots = md.block_table[:B_, 0].long()
            st = cache[slots]                                                     # [B, state]
            conv_state, h = self._split_state(st)
            xin = torch.cat([conv_state, x[:B_].float()[:, :, None]], dim=-1)      # [B, di, k]
            conv = (xin * self.conv_weight.float()[None]).sum(-1)
            if self.conv_bias is not None:
                conv = conv + self.conv_bias.float()
            xa = F.silu(conv)
            dt, Bm, Cm = self._ssm_params(xa.to(x.dtype))
            A = -torch.exp(self.A_log)
            h = torch.exp(dt[:, :, None] * A[None]) * h + (dt * xa)[:, :, None] * Bm[:, None, :]
            y = (h * Cm[:, None, :]).sum(-1) + self.D * xa
            cache[slots] = torch.cat([xin[:, :, 1:].reshape(B_, -1), h.reshape(B_, -1)], dim=-1)
            out = torch.zeros(T, self.d_inner, dtype=torch.float32, device=x.device)
            out[:B_] = y
        else:
            qsl = (md.query_start_loc_cpu if md.query_start_loc_cpu is not None else md.query_start_loc.cpu()).tolist()
            sls = (md.seq_lens_cpu if md.seq_lens_cpu is not None else md.seq_lens.cpu()).tolist()
            slots = md.block_table[:md.num_reqs, 0].tolist()
            out = torch.zeros(T, self.d_inner, dtype=torch.float32, device=x.device)
            for i in range(md.num_reqs):
                s, e = qsl[i], qsl[i + 1]
                if e <= s:
                    continue
                if sls[i] - (e - s) > 0:
                    conv_state, h = self._split_state(cache[slots[i]])
                    conv_state, h = conv_state.clone(), h.clone()
                else:
                    conv_state = torch.zeros(self.d_inner, self.d_conv - 1, dtype=torch.float32, device=x.device)
                    h = torch.zeros(self.d_inner, self.d_state, dtype=torch.float32, device=x.device)
                y, conv_state, h = self._segment(x[s:e], conv_state, h)
                cache[slots[i]] = torch.cat([conv_state.reshape(-1), h.reshape(-1)])
                out[s:e] = y
        return self.out_proj((out * F.silu(z.float())).to(hidden.dtype))


class Mamba2Mixer(nn.Module):
    """Mamba-2 (SSD) mixer: scalar decay per head, B/C shared by head groups, gated RMSNorm before the output
    projection.  Reference: vllm/model_executor/layers/mamba/mamba_mixer2.py (+ ops/ssd_*.py, causal_conv1d.py,
    mamba_ssm.py ``selective_state_update``).

    Cache block per request (fp32): ``[conv_state (conv_dim x (k-1)) | ssm_state (heads x head_dim x d_state)]``.
    Decode batches are one vectorised state update (graph-capturable); prefill / mixed batches run the chunked SSD
    scan per request and resume from the cached state (chunked prefill, prefix resume)."""

    def __init__(self, hidden: int, num_heads: int, head_dim: int, d_state: int, d_conv: int, n_groups: int,
                 use_bias: bool, use_conv_bias: bool, chunk_size: int, eps: float, dtype, prefix: str,
                 time_step_limit: tuple[float, float] = (0.0, float("inf")), norm_before_gate: bool = False,
                 norm_groups: int = 1) -> None:
        super().__init__()
        self.nh, self.hd, self.N, self.k, self.G = num_heads, head_dim, d_state, d_conv, n_groups
        self.norm_groups = norm_groups        # gated RMSNorm statistics per group of channels (Nemotron-H, Zamba-2, Falcon-H1)
        # gate placement: False → norm(y · silu(z)) (Mamba-2 default), True → norm(y) · silu(z), None → y · silu(z), no norm
        self.norm_before_gate = norm_before_gate
        self.d_inner = num_heads * head_dim
        self.conv_dim = self.d_inner + 2 * n_groups * d_state
        self.chunk = chunk_size
        self.dt_min, self.dt_max = time_step_limit
        self.in_proj = ReplicatedLinear(hidden, 2 * self.d_inner + 2 * n_groups * d_state + num_heads, use_bias, dtype,
                                        None, f"{prefix}.in_proj")
        self.conv_weight = nn.Parameter(torch.empty(self.conv_dim, d_conv, dtype=dtype), requires_grad=False)
        self.conv_bias = nn.Parameter(torch.zeros(self.conv_dim, dtype=dtype), requires_grad=False) if use_conv_bias else None
        self.dt_bias = nn.Parameter(torch.ones(num_heads, dtype=torch.float32), requires_grad=False)
        self.A_log = nn.Parameter(torch.zeros(num_heads, dtype=torch.float32), requires_grad=False)
        self.D = nn.Parameter(torch.ones(num_heads, dtype=torch.float32), requires_grad=False)
        self.norm_weight = nn.Parameter(torch.ones(self.d_inner, dtype=dtype), requires_grad=False)
        self.norm_eps = eps
        self.out_proj = ReplicatedLinear(self.d_inner, hidden, use_bias, dtype, None, f"{prefix}.out_proj")
        self.layer_name = prefix
        self.state_elems = self.conv_dim * (d_conv - 1) + num_heads * head_dim * d_state
        self.num_kv_heads, self.head_size, self.sliding_window = 1, 0, None
        assert prefix not in _attn._LAYER_REGISTRY, f"duplicate layer name {prefix}"
        _attn._LAYER_REGISTRY[prefix] = self

    def get_kv_cache_spec(self, block_size: int, dtype_size: int) -> KVCacheSpec:
        return MambaSpec(block_size=block_size, state_bytes=self.state_elems * 4)

    def _split_state(self, st: torch.Tensor):
        nconv = self.conv_dim * (self.k - 1)
        return (st[..., :nconv].reshape(*st.shape[:-1], self.conv_dim, self.k - 1),
                st[..., nconv:].reshape(*st.shape[:-1], self.nh, self.hd, self.N))

    def _dt(self, dt: torch.Tensor) -> torch.Tensor:
        dt = F.softplus(dt.float() + self.dt_bias)
        return dt.clamp(self.dt_min, self.dt_max) if (self.dt_min > 0 or self.dt_max != float("inf")) else dt

    def _gated_norm(self, y: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
        if self.norm_before_gate is None:
            return y.float() * F.silu(z.float())
        y = y.float() if self.norm_before_gate else y.float() * F.silu(z.float())
        if self.norm_groups > 1:
            g = y.view(*y.shape[:-1], self.norm_groups, -1)
            y = (g * torch.rsqrt(g.pow(2).mean(-1, keepdim=True) + self.norm_eps)).view(y.shape)
        else:
            y = y * torch.rsqrt(y.pow(2).mean(-1, keepdim=True) + self.norm_eps)
        y = self.norm_weight.float() * y
        return y * F.silu(z.float()) if self.norm_before_gate else y

    def _segment(self, xbc: torch.Tensor, dt: torch.Tensor, conv_state: torch.Tensor, h: torch.Tensor):
        """ONE request: xbc [q, conv_dim], dt [q, nh] raw → (y [q, d_inner] fp32, conv_state, h)."""
        from vllm_b200.ops import ref
        q = xbc.shape[0]
        xin = torch.cat([conv_state.t().to(xbc.dtype), xbc], dim=0)                      # [k-1+q, conv_dim]
        conv = (xin.float().unfold(0, self.k, 1) * self.conv_weight.float()[None]).sum(-1)
        if self.conv_bias is not None:
            conv = conv + self.conv_bias.float()
        xa = F.silu(conv)
        x, B, C = xa.split([self.d_inner, self.G * self.N, self.G * self.N], dim=-1)
        y, h = ref.ssd_chunk_scan(x.view(q, self.nh, self.hd), self._dt(dt), -torch.exp(self.A_log),
                                  B.view(q, self.G, self.N), C.view(q, self.G, self.N), self.chunk, h)
        y = y + self.D[None, :, None] * x.view(q, self.nh, self.hd)
        return y.reshape(q, self.d_inner), xin[-(self.k - 1):].t().float(), h

    def forward(self, hidden: torch.Tensor) -> torch.Tensor:
        ctx = get_forward_context()
        md = ctx.metadata_for(self.layer_name)
        cache = ctx.kv_caches.get(self.layer_name)
        z, xbc, dt = self.in_proj(hidden).split([self.d_inner, self.conv_dim, self.nh], dim=-1)
        T = hidden.shape[0]
        dev = hidden.device
        if md is None or cache is None:                      # profiling run: one stateless sequence
            y, _, _ = self._segment(xbc, dt, torch.zeros(self.conv_dim, self.k - 1, device=dev),
                                    torch.zeros(self.nh, self.hd, self.N, device=dev))
            return self.out_proj(self._gated_norm(y, z).to(hidden.dtype))
        if _cuda_ssm(xbc) and self.N % 32 == 0 and self.N <= 256 and 2 <= self.k <= 4:
            from vllm_b200.ops import cuda_ops
            cu = cuda_ops()
            xa = cu.causal_conv1d(xbc, self.conv_weight, self.conv_bias, cache, md)
            y = cu.ssd_scan(xa, dt.float(), self.dt_bias, self.A_log, self.D, cache, self.conv_dim * (self.k - 1), md,
                            self.nh, self.hd, self.G, self.N, self.dt_min, self.dt_max)
            return self.out_proj(self._gated_norm(y, z).to(hidden.dtype))
        cache = cache.view(cache.shape[0], -1).view(torch.float32) if cache.dtype != torch.float32 else cache
        out = torch.zeros(T, self.d_inner, dtype=torch.float32, device=dev)
        if md.decode_only:
            Bn = md.num_reqs
            slots = md.block_table[:Bn, 0].long()
            conv_state, h = self._split_state(cache[slots])
            xin = torch.cat([conv_state, xbc[:Bn].float()[:, :, None]], dim=-1)            # [B, conv_dim, k]
            conv = (xin * self.conv_weight.float()[None]).sum(-1)
            if self.conv_bias is not None:
                conv = conv + self.conv_bias.float()
            xa = F.silu(conv)
            x, Bm, Cm = xa.split([self.d_inner, self.G * self.N, self.G * self.N], dim=-1)
            x = x.view(Bn, self.nh, self.hd)
            rep = self.nh // self.G
            Bm = Bm.view(Bn, self.G, self.N).repeat_interleave(rep, dim=1)                  # [B, nh, N]
            Cm = Cm.view(Bn, self.G, self.N).repeat_interleave(rep, dim=1)
            dtv = self._dt(dt[:Bn])                                                         # [B, nh]
            dA = torch.exp(dtv * (-torch.exp(self.A_log))[None])
            h = h * dA[:, :, None, None] + (dtv[:, :, None] * x)[..., None] * Bm[:, :, None, :]
            y = (h * Cm[:, :, None, :]).sum(-1) + self.D[None, :, None] * x
            cache[slots] = torch.cat([xin[:, :, 1:].reshape(Bn, -1), h.reshape(Bn, -1)], dim=-1)
            out[:Bn] = y.reshape(Bn, self.d_inner)
        else:
            qsl = (md.query_start_loc_cpu if md.query_start_loc_cpu is not None else md.query_start_loc.cpu()).tolist()
            sls = (md.seq_lens_cpu if md.seq_lens_cpu is not None else md.seq_lens.cpu()).tolist()
            slots = md.block_table[:md.num_reqs, 0].tolist()
            for i in range(md.num_reqs):
                s, e = qsl[i], qsl[i + 1]
                if e <= s:
                    continue
                if sls[i] - (e - s) > 0:
                    conv_state, h = self._split_state(cache[slots[i]])
                    conv_state, h = conv_state.clone(), h.clone()
                else:
                    conv_state = torch.zeros(self.conv_dim, self.k - 1, dtype=torch.float32, device=dev)
                    h = torch.zeros(self.nh, self.hd, self.N, dtype=torch.float32, device=dev)
                y, conv_state, h = self._segment(xbc[s:e], dt[s:e], conv_state, h)
                cache[slots[i]] = torch.cat([conv_state.reshape(-1), h.reshape(-1)])
                out[s:e] = y
        return self.out_proj(self._gated_norm(out, z).to(hidden.dtype))
