"""Quantisation configs and linear methods.

Reference registry: vllm/model_executor/layers/quantization/__init__.py:13-51.
Implemented here: ``fp8`` (per-tensor weight scale, dynamic per-token activation),
``fp8_block`` (DeepSeek-V3 128x128 block-scaled weights + per-token-group
activations; reference utils/fp8_utils.py, kernels/linear/scaled_mm/deep_gemm.py)
— on B200 these map to ``tcgen05.mma.kind::f8f6f4`` tiles with fp32 scale
application per K-block.  Online quantisation of bf16 checkpoints is supported
(``quantization="fp8_block"`` on an unquantised model).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from vllm_b200 import ops
from vllm_b200.layers.linear import LinearMethod
from vllm_b200.ops import ref


class QuantConfig:
    name = "none"

    def get_linear_method(self, layer: nn.Module, prefix: str) -> Optional[LinearMethod]:
        return None


class Fp8BlockLinearMethod(LinearMethod):
    def __init__(self, block: int = 128, online: bool = False) -> None:
        self.block = block
        self.online = online

    def create_weights(self, layer, in_per_rank, out_per_rank, dtype) -> None:
        b = self.block
        if self.online:
            layer.weight = nn.Parameter(torch.empty(out_per_rank, in_per_rank, dtype=dtype), requires_grad=False)
            return
        layer.weight = nn.Parameter(torch.empty(out_per_rank, in_per_rank, dtype=torch.float8_e4m3fn),
                                    requires_grad=False)
        sc = nn.Parameter(torch.ones(-(-out_per_rank // b), -(-in_per_rank // b), dtype=torch.float32),
                          requires_grad=False)
        sc.shard_block = b
        layer.weight_scale_inv = sc

    def process_weights_after_loading(self, layer) -> None:
        if self.online and layer.weight.dtype != torch.float8_e4m3fn:
            q, s = ref.block_quant_weight_fp8(layer.weight.data, self.block)
            layer.weight = nn.Parameter(q, requires_grad=False)
            layer.weight_scale_inv = nn.Parameter(s, requires_grad=False)

    def apply(self, layer, x, bias):
        return ops.block_fp8_linear(x, layer.weight, layer.weight_scale_inv, self.block, bias)


class Fp8BlockConfig(QuantConfig):
    name = "fp8_block"

    def __init__(self, block: int = 128, online: bool = False, ignored: tuple[str, ...] = ("lm_head", "gate.")) -> None:
        self.block, self.online, self.ignored = block, online, ignored

    def get_linear_method(self, layer, prefix):
        if any(s in prefix for s in self.ignored):
            return None
        return Fp8BlockLinearMethod(self.block, self.online)


class Fp8LinearMethod(LinearMethod):
    """Per-tensor (or per-output-channel) fp8 e4m3 weights + dynamic per-token activation scales
    (reference: quantization/fp8.py ``Fp8LinearMethod``, kernels/linear/scaled_mm).  Checkpoints in this format
    carry ``weight`` (fp8) and ``weight_scale`` (scalar or [out]); bf16 checkpoints are quantised after loading."""

    def __init__(self, online: bool, per_channel: bool = False) -> None:
        self.online, self.per_channel = online, per_channel

    def create_weights(self, layer, in_per_rank, out_per_rank, dtype) -> None:
        wdt = dtype if self.online else torch.float8_e4m3fn
        layer.weight = nn.Parameter(torch.empty(out_per_rank, in_per_rank, dtype=wdt), requires_grad=False)
        layer.weight_scale = nn.Parameter(torch.ones(out_per_rank if self.per_channel else 1, dtype=torch.float32),
                                          requires_grad=False)
        layer.out_dtype = dtype

    def process_weights_after_loading(self, layer) -> None:
        if layer.weight.dtype == torch.float8_e4m3fn:
            return
        w = layer.weight.data.float()
        amax = w.abs().amax(dim=1 if self.per_channel else None).clamp_min(1e-12)
        scale = (amax / ref.FP8_MAX).reshape(-1)
        q = (w / (scale[:, None] if self.per_channel else scale)).clamp(-ref.FP8_MAX, ref.FP8_MAX)
        layer.weight = nn.Parameter(q.to(torch.float8_e4m3fn), requires_grad=False)
        layer.weight_scale = nn.Parameter(scale.to(torch.float32), requires_grad=False)

    def apply(self, layer, x, bias):
        # dynamic per-token activation quantisation, fp32 accumulation, scales applied in the epilogue
        xf = x.float()
        a_scale = (xf.abs().amax(dim=-1, keepdim=True).clamp_min(1e-12) / ref.FP8_MAX)
        xq = (xf / a_scale).clamp(-ref.FP8_MAX, ref.FP8_MAX).to(torch.float8_e4m3fn)
        w_scale = layer.weight_scale
        if x.is_cuda and hasattr(torch, "_scaled_mm") and x.dim() == 2 and x.shape[-1] % 16 == 0 \
                and layer.weight.shape[0] % 16 == 0:
            # cuBLASLt fp8 GEMM (library) until the block-scaled tcgen05 kernel covers this case too
            y = torch._scaled_mm(xq, layer.weight.t(), scale_a=a_scale.float(),
                                 scale_b=(w_scale.reshape(1, -1) if self.per_channel
                                          else w_scale.expand(layer.weight.shape[0]).reshape(1, -1)).float(),
                                 out_dtype=layer.out_dtype)
        else:
            y = (xq.float() @ layer.weight.float().t()) * a_scale * (w_scale if self.per_channel else w_scale[0])
            y = y.to(layer.out_dtype)
        return y if bias is None else y + bias


class Fp8Config(QuantConfig):
    name = "fp8"

    def __init__(self, online: bool, per_channel: bool = False, ignored: tuple[str, ...] = ("lm_head", "gate.")) -> None:
        self.online, self.per_channel, self.ignored = online, per_channel, ignored

    def get_linear_method(self, layer, prefix):
        if any(s in prefix for s in self.ignored):
            return None
        return Fp8LinearMethod(self.online, self.per_channel)


QUANTIZATION_METHODS = ("fp8", "fp8_block", "fbgemm_fp8", "modelopt", "deepseek_v3_fp8", "awq", "gptq", "awq_marlin",
                        "gptq_marlin", "nvfp4", "modelopt_fp4")


def get_quant_config(model_config) -> Optional[QuantConfig]:
    q = model_config.quantization
    if q is None:
        return None
    if q == "fp8_block":
        online = not model_config.hf_config.get("quantization_config")
        return Fp8BlockConfig(online=online)
    hq = model_config.hf_config.get("quantization_config") or {}
    if q in ("awq", "gptq", "awq_marlin", "gptq_marlin", "nvfp4", "modelopt_fp4"):
        return None          # expanded to the model dtype at load time (model_loader/weight_only.py)
    if q in ("fp8", "fbgemm_fp8", "modelopt"):
        if hq.get("weight_block_size"):          # a "fp8" checkpoint with block scales (DeepSeek-V3 style)
            return Fp8BlockConfig(online=False)
        return Fp8Config(online=not hq, per_channel=(q == "fbgemm_fp8" or hq.get("strategy") == "channel"))
    if q == "deepseek_v3_fp8":
        return Fp8BlockConfig(online=not hq)
    raise ValueError(f"quantization method {q!r} is not supported yet")
