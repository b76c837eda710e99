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


def get_quant_config(model_config) -> Optional[QuantConfig]:
    q = model_config.quantization
    if q is None:
        return None
    if q == "fp8_block":
        online = not model_config.hf_config.get("quantization_config")
        return Fp8BlockConfig(online=online)
    if q == "fp8":
        return Fp8BlockConfig(online=True)
    raise ValueError(f"quantization method {q!r} is not supported yet")
