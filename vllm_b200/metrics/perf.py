"""Analytical per-step cost model: FLOPs and HBM bytes of one forward pass from the model config, and the
model-FLOPs / model-bandwidth utilisation of a measured step (reference: vllm/v1/metrics/perf.py:226-1265,
``VLLM_DEBUG_MFU_METRICS``).  Used by ``bench.py``-style reports and the periodic stats log.

Conventions: one multiply-accumulate = 2 FLOPs; weights are read once per step per GPU (decode regime), KV is
read once per attended token; activations are ignored (they stay in L2 / SMEM at decode batch sizes).
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Optional


@dataclass
class StepCost:
    flops: float          # whole model, all GPUs
    weight_bytes: float   # per step, all GPUs
    kv_bytes: float       # per step, all GPUs

    @property
    def bytes(self) -> float:
        return self.weight_bytes + self.kv_bytes


def _peaks() -> tuple[float, float]:
    """(bytes/s, flops/s) per GPU: MEASURED_PEAKS.json if present, else the B200 profiling-recipe fallbacks."""
    here = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    p = os.path.join(here, "MEASURED_PEAKS.json")
    bw, fl = 6.5e12, 1.6e15
    try:
        with open(p) as f:
            d = json.load(f)

        def find(o, keys):
            if isinstance(o, dict):
                for k, v in o.items():
                    if any(s in k.lower() for s in keys) and isinstance(v, (int, float)):
                        return float(v)
                    r = find(v, keys)
                    if r:
                        return r
            return None
        g = find(d, ("gb_s", "gbps", "bandwidth"))
        t = find(d, ("tflops", "tflop"))
        if g:
            bw = g * 1e9
        if t:
            fl = t * 1e12
    except Exception:
        pass
    return bw, fl


def model_step_cost(hf: dict, num_tokens: int, context_lens: list[int], dtype_bytes: int = 2,
                    kv_dtype_bytes: int = 2) -> StepCost:
    """``num_tokens`` new tokens this step; ``context_lens[i]`` = tokens attended by request i (incl. new ones)."""
    H = hf.get("hidden_size", hf.get("n_embd", 0))
    L = hf.get("num_hidden_layers", hf.get("n_layer", 0))
    V = hf["vocab_size"]
    nh = hf.get("num_attention_heads", hf.get("n_head", 1))
    nkv = hf.get("num_key_value_heads") or nh
    hd = hf.get("head_dim") or (H // nh if nh else 0)
    I = hf.get("intermediate_size") or 4 * H
    # ---- per-layer parameters
    if hf.get("kv_lora_rank"):                                     # MLA
        ql = hf.get("q_lora_rank") or 0
        qh = hf["qk_nope_head_dim"] + hf["qk_rope_head_dim"]
        attn_p = (H * ql + ql * nh * qh if ql else H * nh * qh) + H * (hf["kv_lora_rank"] + hf["qk_rope_head_dim"]) \
            + hf["kv_lora_rank"] * nh * (hf["qk_nope_head_dim"] + hf["v_head_dim"]) + nh * hf["v_head_dim"] * H
        kv_per_tok = hf["kv_lora_rank"] + hf["qk_rope_head_dim"]
        attn_dim = hf["kv_lora_rank"] + hf["qk_rope_head_dim"] + hf["kv_lora_rank"]   # QK over 576, PV over 512
    else:
        attn_p = H * (nh + 2 * nkv) * hd + nh * hd * H
        kv_per_tok = 2 * nkv * hd
        attn_dim = 2 * hd
    n_exp = hf.get("num_local_experts") or hf.get("n_routed_experts") or hf.get("num_experts") or 0
    topk = hf.get("num_experts_per_tok", 0)
    gated = 3 if hf.get("hidden_act", "silu") in ("silu", "gelu", "gelu_pytorch_tanh") and hf.get("model_type") not in ("opt", "gpt2", "bert") else 2
    if n_exp:
        Ie = hf.get("moe_intermediate_size") or I
        shared = (hf.get("n_shared_experts") or 0) * 3 * H * Ie + (3 * H * hf["shared_expert_intermediate_size"] if hf.get("shared_expert_intermediate_size") else 0)
        dense_layers = hf.get("first_k_dense_replace", 0)
        moe_layers = L - dense_layers
        mlp_p_total = n_exp * 3 * H * Ie + shared + H * n_exp
        mlp_p_active = topk * 3 * H * Ie + shared + H * n_exp
        # experts touched by a batch of T tokens (expected, uniform routing) bound the weight traffic
        frac = 1.0 - (1.0 - topk / n_exp) ** max(num_tokens, 1)
        mlp_bytes = (n_exp * 3 * H * Ie * frac + shared + H * n_exp) * dtype_bytes
        params_flops = L * attn_p + moe_layers * mlp_p_active + dense_layers * gated * H * I
        weight_bytes = (L * attn_p + dense_layers * gated * H * I) * dtype_bytes + moe_layers * mlp_bytes
        del mlp_p_total
    else:
        params_flops = L * (attn_p + gated * H * I)
        weight_bytes = params_flops * dtype_bytes
    params_flops += V * H                                          # LM head (embedding lookup is free)
    weight_bytes += V * H * dtype_bytes
    ctx = float(sum(context_lens))
    attn_flops = 2.0 * L * nh * attn_dim * ctx                     # QK^T + PV for one new token per request
    if num_tokens > len(context_lens):                             # prefill: causal triangle, approximate by T*ctx/2
        attn_flops *= (num_tokens / max(len(context_lens), 1)) / 2.0
    flops = 2.0 * params_flops * num_tokens + attn_flops
    kv_bytes = L * kv_per_tok * kv_dtype_bytes * ctx
    return StepCost(flops, weight_bytes, kv_bytes)


def utilisation(cost: StepCost, step_seconds: float, n_gpus: int = 1) -> dict[str, float]:
    bw, fl = _peaks()
    return {
        "mfu": cost.flops / step_seconds / (fl * n_gpus),
        "mbu": cost.bytes / step_seconds / (bw * n_gpus),
        "roofline_ms": max(cost.flops / (fl * n_gpus), cost.bytes / (bw * n_gpus)) * 1e3,
        "achieved_tb_s": cost.bytes / step_seconds / 1e12,
        "achieved_tflops": cost.flops / step_seconds / 1e12,
    }
