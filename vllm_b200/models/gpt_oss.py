"""GPT-OSS: attention sinks, alternating sliding-window / full layers, biased MoE with the clamped
``(up + 1) * gate * sigmoid(1.702 * gate)`` activation, YaRN rope; MXFP4 expert checkpoints are expanded at load.

Reference parity: vllm/model_executor/models/gpt_oss.py (``OAIAttention`` sinks, ``MLPBlock`` router with bias,
``swigluoai`` activation, mxfp4 ``*_blocks`` / ``*_scales`` expert tensors).
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from vllm_b200.config import VllmConfig
from vllm_b200.layers.attention import Attention
from vllm_b200.layers.common import RMSNorm
from vllm_b200.layers.linear import QKVParallelLinear, RowParallelLinear
from vllm_b200.layers.rotary import get_rope
from vllm_b200.models.llama import LlamaForCausalLM, LlamaModel
from vllm_b200.parallel.state import get_tp_group
from vllm_b200.parallel.utils import divide

_FP4 = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0, -0.0, -0.5, -1.0, -1.5, -2.0, -3.0, -4.0, -6.0])


def dequant_mxfp4(blocks: torch.Tensor, scales: torch.Tensor, dtype=torch.bfloat16) -> torch.Tensor:
    """[..., G, 16] uint8 (two e2m1 values per byte, low nibble first) + [..., G] e8m0 exponents → [..., G*32]."""
    lut = _FP4.to(blocks.device)
    lo, hi = (blocks & 0x0F).long(), (blocks >> 4).long()
    vals = torch.stack([lut[lo], lut[hi]], dim=-1).reshape(*blocks.shape[:-1], -1)            # [..., G, 32]
    vals = torch.ldexp(vals, (scales.to(torch.int32) - 127)[..., None])
    return vals.reshape(*blocks.shape[:-2], -1).to(dtype)


class GptOssAttention(nn.Module):
    def __init__(self, cfg: dict, mc, dtype, quant_config, prefix: str, layer_idx: int) -> None:
        super().__init__()
        H, nh = cfg["hidden_size"], cfg["num_attention_heads"]
        nkv = cfg.get("num_key_value_heads") or nh
        hd = cfg.get("head_dim") or H // nh
        bias = bool(cfg.get("attention_bias", True))
        self.qkv_proj = QKVParallelLinear(H, hd, nh, nkv, bias, dtype, quant_config, f"{prefix}.qkv_proj")
        self.o_proj = RowParallelLinear(nh * hd, H, bias, dtype=dtype, quant_config=quant_config, prefix=f"{prefix}.o_proj")
        self.q_size, self.kv_size = self.qkv_proj.q_size, self.qkv_proj.kv_size
        types = cfg.get("layer_types")
        sliding = (types[layer_idx] == "sliding_attention") if types else (layer_idx % 2 == 0)
        max_pos = mc.max_model_len if mc.max_model_len > 0 else cfg.get("max_position_embeddings", 131072)
        rp = cfg.get("rope_parameters") or {}
        scaling = cfg.get("rope_scaling") or ({k: v for k, v in rp.items() if k != "rope_theta"}
                                               if rp.get("rope_type", "default") != "default" else None)
        rope = get_rope(hd, hd, max_pos, rp.get("rope_theta", cfg.get("rope_theta", 150000.0)), True, scaling)
        self.total_heads = nh
        self.sinks = nn.Parameter(torch.zeros(self.qkv_proj.num_heads, dtype=torch.float32), requires_grad=False)
        self.attn = Attention(self.qkv_proj.num_heads, hd, hd ** -0.5, self.qkv_proj.num_kv_heads,
                              sliding_window=cfg.get("sliding_window") if sliding else None, prefix=f"{prefix}.attn",
                              rope=rope, sinks=self.sinks)

    def load_sinks(self, w: torch.Tensor) -> None:
        g = get_tp_group()
        n = self.sinks.shape[0]
        r = max(g.rank_in_group, 0) if g.world_size > 1 else 0
        self.sinks.data.copy_(w[r * n:(r + 1) * n].float())

    def forward(self, positions, hidden):
        q, k, v = self.qkv_proj(hidden).split([self.q_size, self.kv_size, self.kv_size], dim=-1)
        return self.o_proj(self.attn(q, k, v, positions))


class GptOssMoE(nn.Module):
    """Top-k over biased router logits, softmax over the selected k; experts carry biases; intermediate dim is
    TP-sharded (down-projection bias is applied on rank 0 so the all-reduce adds it once)."""

    ALPHA, LIMIT = 1.702, 7.0

    def __init__(self, cfg: dict, dtype, prefix: str) -> None:
        super().__init__()
        g = get_tp_group()
        self.tp_rank, self.tp_size = (max(g.rank_in_group, 0), g.world_size) if g.world_size > 1 else (0, 1)
        H, I, E = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_local_experts"]
        self.top_k, self.E, self.I = cfg["num_experts_per_tok"], E, divide(I, self.tp_size)
        self.limit = float(cfg.get("swiglu_limit", self.LIMIT))
        self.router = nn.Linear(H, E, dtype=dtype)
        P = lambda *s: nn.Parameter(torch.zeros(*s, dtype=dtype), requires_grad=False)   # noqa: E731
        self.w13, self.b13 = P(E, 2 * self.I, H), P(E, 2 * self.I)          # rows: [gate_local | up_local]
        self.w2, self.b2 = P(E, H, self.I), P(E, H)

    def _act(self, gu: torch.Tensor) -> torch.Tensor:
        gate, up = gu[..., :self.I], gu[..., self.I:]
        gate = gate.clamp(max=self.limit)
        up = up.clamp(min=-self.limit, max=self.limit)
        return (up + 1) * gate * torch.sigmoid(gate * self.ALPHA)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        T, H = x.shape
        logits = self.router(x)
        top, ids = torch.topk(logits, self.top_k, dim=-1)
        w = torch.softmax(top.float(), dim=-1).to(x.dtype)
        bias_on = 1.0 if self.tp_rank == 0 else 0.0
        if T * self.top_k <= 512:
            # decode-sized batches: gather the selected experts' matrices and run batched GEMVs — no host sync,
            # static shapes, CUDA-graph capturable
            flat = ids.reshape(-1)
            xk = x.repeat_interleave(self.top_k, dim=0).unsqueeze(1)                                     # [T*k, 1, H]
            gu = torch.baddbmm(self.b13[flat].unsqueeze(1), xk, self.w13[flat].transpose(1, 2))          # [T*k, 1, 2I]
            y = torch.bmm(self._act(gu), self.w2[flat].transpose(1, 2)).squeeze(1) + bias_on * self.b2[flat]
            out = (y.view(T, self.top_k, H) * w.unsqueeze(-1)).sum(1)
        else:
            out = torch.zeros_like(x)
            for e in torch.unique(ids).tolist():
                tok, slot = torch.where(ids == e)
                gu = F.linear(x[tok], self.w13[e], self.b13[e])
                y = F.linear(self._act(gu), self.w2[e]) + bias_on * self.b2[e]
                out.index_add_(0, tok, y * w[tok, slot, None])
        if self.tp_size > 1:
            out = get_tp_group().all_reduce(out)
        return out

    # ---- checkpoint tensors (HF layout: gate_up_proj [E, H, 2I] with gate/up interleaved on the last dim,
    #      down_proj [E, I, H]; mxfp4: *_blocks [E, rows, G, 16] + *_scales [E, rows, G] already [E, out, in])
    def load(self, name: str, w: torch.Tensor) -> bool:
        I, r = self.I, self.tp_rank
        if name == "gate_up_proj":
            w = w.transpose(1, 2)                                            # [E, 2I_total, H]
        if name in ("gate_up_proj", "gate_up_proj_dq"):
            gate, up = w[:, 0::2], w[:, 1::2]
            self.w13.data[:, :I].copy_(gate[:, r * I:(r + 1) * I])
            self.w13.data[:, I:].copy_(up[:, r * I:(r + 1) * I])
        elif name == "gate_up_proj_bias":
            gate, up = w[:, 0::2], w[:, 1::2]
            self.b13.data[:, :I].copy_(gate[:, r * I:(r + 1) * I])
            self.b13.data[:, I:].copy_(up[:, r * I:(r + 1) * I])
        elif name == "down_proj":
            self.w2.data.copy_(w.transpose(1, 2)[:, :, r * I:(r + 1) * I])  # [E, H, I]
        elif name == "down_proj_dq":
            self.w2.data.copy_(w[:, :, r * I:(r + 1) * I])
        elif name == "down_proj_bias":
            self.b2.data.copy_(w)
        else:
            return False
        return True


class GptOssDecoderLayer(nn.Module):
    def __init__(self, cfg, mc, dtype, quant_config, prefix: str, layer_idx: int) -> None:
        super().__init__()
        H, eps = cfg["hidden_size"], cfg.get("rms_norm_eps", 1e-5)
        self.self_attn = GptOssAttention(cfg, mc, dtype, quant_config, f"{prefix}.self_attn", layer_idx)
        self.mlp = GptOssMoE(cfg, dtype, f"{prefix}.mlp")
        self.input_layernorm = RMSNorm(H, eps, dtype)
        self.post_attention_layernorm = RMSNorm(H, eps, dtype)

    def forward(self, positions, hidden, residual, next_norm: Optional[nn.Module]):
        stream = residual + self.self_attn(positions, hidden)
        stream = stream + self.mlp(self.post_attention_layernorm(stream))
        return (next_norm(stream) if next_norm is not None else stream), stream


class GptOssModel(LlamaModel):
    def _make_layer(self, cfg, mc, dtype, quant_config, prefix, idx):
        return GptOssDecoderLayer(cfg, mc, dtype, quant_config, prefix, idx)


class GptOssForCausalLM(LlamaForCausalLM):
    model_cls = GptOssModel

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        modules = dict(self.named_modules())
        loaded: set[str] = set()
        pending: dict[str, torch.Tensor] = {}
        rest = []
        for name, w in weights:
            local = self._localize(name)
            if local is None:
                continue
            if ".mlp.experts." in local:
                mod_name, leaf = local.split(".mlp.experts.")
                moe = modules.get(mod_name + ".mlp")
                if moe is None:
                    continue
                if leaf.endswith("_blocks") or leaf.endswith("_scales"):
                    base = leaf.rsplit("_", 1)[0]
                    other = f"{mod_name}.{base}_{'scales' if leaf.endswith('_blocks') else 'blocks'}"
                    if other in pending:
                        o = pending.pop(other)
                        blocks, scales = (w, o) if leaf.endswith("_blocks") else (o, w)
                        moe.load(base + "_dq", dequant_mxfp4(blocks, scales, moe.w13.dtype))
                        loaded.add(f"{mod_name}.mlp.{'w13' if base.startswith('gate') else 'w2'}")
                    else:
                        pending[f"{mod_name}.{leaf}"] = w
                elif moe.load(leaf, w):
                    loaded.add(f"{mod_name}.mlp.{leaf}")
                continue
            if local.endswith(".self_attn.sinks"):
                modules[local.rsplit(".", 1)[0]].load_sinks(w)
                loaded.add(local)
                continue
            rest.append((name, w))
        return loaded | super().load_weights(rest)
