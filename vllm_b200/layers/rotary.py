"""Rotary position embeddings: cos/sin cache construction for the scaling
families the reference supports (vllm/model_executor/layers/rotary_embedding/:
base, linear, dynamic-NTK, YaRN, DeepSeek-YaRN, Llama-3, LongRoPE/Phi-3, M-RoPE).

The cache is a single fp32 tensor ``[max_pos, rot_dim]`` = cat(cos, sin); the
CUDA kernel (fused with the KV write) and the reference op both index it by
position.
"""
from __future__ import annotations

import math
from typing import Any, Optional

import torch
import torch.nn as nn

from vllm_b200 import ops


def _inv_freq(base: float, rot_dim: int) -> torch.Tensor:
    return 1.0 / (base ** (torch.arange(0, rot_dim, 2, dtype=torch.float32) / rot_dim))


def _yarn_find_dim(num_rot: float, dim: int, base: float, max_pos: int) -> float:
    return (dim * math.log(max_pos / (num_rot * 2 * math.pi))) / (2 * math.log(base))


def _yarn_range(low_rot: float, high_rot: float, dim: int, base: float, max_pos: int, truncate: bool = True):
    low, high = _yarn_find_dim(low_rot, dim, base, max_pos), _yarn_find_dim(high_rot, dim, base, max_pos)
    if truncate:                                  # GPT-OSS checkpoints set ``truncate: false`` (fractional ramp ends)
        low, high = math.floor(low), math.ceil(high)
    return max(low, 0), min(high, dim - 1)


def _yarn_ramp(low: float, high: float, n: int) -> torch.Tensor:
    if low == high:
        high += 0.001
    return ((torch.arange(n, dtype=torch.float32) - low) / (high - low)).clamp(0, 1)


def yarn_mscale(scale: float = 1.0, mscale: float = 1.0) -> float:
    return 1.0 if scale <= 1 else 0.1 * mscale * math.log(scale) + 1.0


class RotaryEmbedding(nn.Module):
    def __init__(self, head_size: int, rotary_dim: int, max_position: int, base: float,
                 is_neox_style: bool = True, rope_scaling: Optional[dict[str, Any]] = None) -> None:
        super().__init__()
        self.head_size, self.rotary_dim = head_size, rotary_dim
        self.max_position, self.base, self.is_neox_style = max_position, base, is_neox_style
        self.rope_scaling = rope_scaling or {}
        self.mscale = 1.0
        self.mrope_section = self.rope_scaling.get("mrope_section")
        cache = self._build()
        self.register_buffer("cos_sin_cache", cache, persistent=False)

    def _build(self) -> torch.Tensor:
        rs = self.rope_scaling
        kind = rs.get("rope_type", rs.get("type", "default"))
        d, base, max_pos = self.rotary_dim, self.base, self.max_position
        inv = _inv_freq(base, d)
        scale_pos = 1.0
        if kind in ("default", "mrope"):
            pass
        elif kind == "linear":
            scale_pos = 1.0 / rs["factor"]
        elif kind == "dynamic":
            f = rs["factor"]
            max_pos = int(max_pos * f)
            base = base * ((f * max_pos / self.max_position) - (f - 1)) ** (d / (d - 2))
            inv = _inv_freq(base, d)
        elif kind == "llama3":
            f, lo, hi = rs["factor"], rs.get("low_freq_factor", 1.0), rs.get("high_freq_factor", 4.0)
            orig = rs["original_max_position_embeddings"]
            wavelen = 2 * math.pi / inv
            smooth = ((orig / wavelen) - lo) / (hi - lo)
            inv = torch.where(wavelen < orig / hi, inv,
                              torch.where(wavelen > orig / lo, inv / f,
                                          (1 - smooth) * inv / f + smooth * inv))
        elif kind in ("yarn", "deepseek_yarn"):
            f = rs["factor"]
            orig = rs.get("original_max_position_embeddings", max_pos)
            lo, hi = _yarn_range(rs.get("beta_fast", 32), rs.get("beta_slow", 1), d, base, orig,
                                 bool(rs.get("truncate", True)))
            mask = (1 - _yarn_ramp(lo, hi, d // 2)) * rs.get("extrapolation_factor", 1.0)
            inv = inv / f * (1 - mask) + inv * mask
            max_pos = max(max_pos, int(orig * f))
            if kind == "deepseek_yarn" or "mscale_all_dim" in rs:
                # DeepSeek: cos/sin carry mscale(mscale)/mscale(mscale_all_dim); the attention layer multiplies its
                # softmax scale by mscale(mscale_all_dim)^2 (reference: rotary_embedding/deepseek_scaling_rope.py)
                self.mscale = (yarn_mscale(f, rs.get("mscale", 1.0)) / yarn_mscale(f, rs.get("mscale_all_dim", 0.0))
                               * rs.get("attn_factor", 1.0))
            else:
                self.mscale = yarn_mscale(f) * rs.get("attn_factor", 1.0)
        elif kind in ("longrope", "su"):
            orig = rs["original_max_position_embeddings"]
            factors = rs["long_factor"] if max_pos > orig else rs["short_factor"]
            inv = inv / torch.tensor(factors, dtype=torch.float32)
            s = max_pos / orig
            self.mscale = 1.0 if s <= 1.0 else math.sqrt(1 + math.log(s) / math.log(orig))
        else:
            raise ValueError(f"unsupported rope scaling type {kind!r}")
        self.max_position = max_pos
        t = torch.arange(max_pos, dtype=torch.float32) * scale_pos
        freqs = torch.outer(t, inv)
        return torch.cat((freqs.cos() * self.mscale, freqs.sin() * self.mscale), dim=-1).contiguous()

    def forward_mrope(self, positions: torch.Tensor, q: torch.Tensor, k: torch.Tensor):
        """positions [3, T] (temporal, height, width ids): channel section i of the rotary dims takes its angle from
        row i % 3 (reference: rotary_embedding/mrope.py; HF ``apply_multimodal_rotary_pos_emb``)."""
        half = self.rotary_dim // 2
        cs = self.cos_sin_cache[positions.long()]                              # [3, T, rot]
        cos3, sin3 = cs[..., :half], cs[..., half:]
        sec = list(self.mrope_section)
        if self.rope_scaling.get("mrope_interleaved"):
            # Qwen3-VL: frequency j belongs to axis j % 3 (t, h, w, t, h, w, …) while that axis still has channels left,
            # the tail falls back to the temporal axis
            j = torch.arange(half)
            idx = torch.zeros(half, dtype=torch.long)
            for ax in (1, 2):
                idx[(j % 3 == ax) & (j < 3 * sec[ax])] = ax
            idx = idx.to(q.device)
        else:
            idx = torch.cat([torch.full((n,), i % 3, dtype=torch.long) for i, n in enumerate(sec)]).to(q.device)[:half]
        ar = torch.arange(half, device=q.device)
        cos = cos3[idx, :, ar].t()                                              # [T, half]
        sin = sin3[idx, :, ar].t()
        T, D = positions.shape[1], self.head_size

        def rot(x):
            xh = x.view(T, -1, D).float()
            x1, x2, rest = xh[..., :half], xh[..., half:2 * half], xh[..., 2 * half:]
            c, s_ = cos[:, None, :], sin[:, None, :]
            return torch.cat((x1 * c - x2 * s_, x2 * c + x1 * s_, rest), dim=-1).to(x.dtype).view(T, -1)
        return rot(q), rot(k)

    def forward(self, positions: torch.Tensor, q: torch.Tensor, k: Optional[torch.Tensor]):
        if self.rotary_dim != self.head_size and q.is_cuda:
            # partial rotary handled by the generic kernel through rot_dim = cache width
            pass
        return ops.rotary_embedding(positions, q, k, self.cos_sin_cache, self.head_size, self.is_neox_style)


_CACHE: dict[tuple, RotaryEmbedding] = {}


def get_rope(head_size: int, rotary_dim: int, max_position: int, base: float, is_neox_style: bool = True,
             rope_scaling: Optional[dict[str, Any]] = None) -> RotaryEmbedding:
    key = (head_size, rotary_dim, max_position, base, is_neox_style,
           repr(sorted(rope_scaling.items())) if rope_scaling else None)
    if key not in _CACHE:
        _CACHE[key] = RotaryEmbedding(head_size, rotary_dim, max_position, base, is_neox_style, rope_scaling)
    return _CACHE[key]
