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

    def __init__(self, online: bool, per_channel: bool = False, static_act: bool = False) -> None:
        self.online, self.per_channel, self.static_act = online, per_channel, static_act and not online

    def create_weights(self, layer, in_per_rank, out_per_rank, dtype) -> None:
        wdt = dtype if self.online else torch.float8_e4m3fn
        layer.weight = nn.Parameter(torch.empty(out_per_rank, in_per_rank, dtype=wdt), requires_grad=False)
        layer.weight_scale = nn.Parameter(torch.ones(out_per_rank if self.per_channel else 1, dtype=torch.float32),
                                          requires_grad=False)
        if self.static_act:      # calibrated per-tensor activation scale shipped with the checkpoint
            layer.input_scale = nn.Parameter(torch.ones(1, dtype=torch.float32), requires_grad=False)
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
        # dynamic per-token (or static per-tensor) activation quantisation, fp32 accumulation, scales in the epilogue
        w_scale = layer.weight_scale
        static = getattr(layer, "input_scale", None)
        if (x.is_cuda and x.dtype in (torch.bfloat16, torch.float16) and x.shape[-1] % 16 == 0
                and layer.out_dtype in (torch.bfloat16, torch.float16)):
            # device path: fused quantiser kernel (csrc/quant.cu) + tcgen05 kind::f8f6f4 GEMM with per-token x
            # per-channel scales in its epilogue (csrc/gemm_fp8.cu) — no library GEMM, no fp32 round trip
            shape = x.shape
            x2 = x.reshape(-1, shape[-1])
            if static is not None:
                a_scale = static.float().max().reshape(1)
                xq, _ = ops.scaled_fp8_quant(x2, a_scale)
            else:
                xq, a_scale = ops.scaled_fp8_quant(x2, per_token=True)
            y = ops.scaled_mm(xq, layer.weight, a_scale, w_scale if self.per_channel else w_scale.reshape(-1)[:1],
                              layer.out_dtype, bias)
            return y.reshape(*shape[:-1], -1)
        xf = x.float()
        if static is not None:
            a_scale = static.float().max().expand(xf.shape[0], 1)
        else:
            a_scale = (xf.abs().amax(dim=-1, keepdim=True).clamp_min(1e-12) / ref.FP8_MAX)
        xq = (xf / a_scale).clamp(-ref.FP8_MAX, ref.FP8_MAX).to(torch.float8_e4m3fn)
        y = (xq.float() @ layer.weight.float().t()) * a_scale * (w_scale if self.per_channel else w_scale[0])
        y = y.to(layer.out_dtype)
        return y if bias is None else y + bias


class Fp8Config(QuantConfig):
    name = "fp8"

    def __init__(self, online: bool, per_channel: bool = False, ignored: tuple[str, ...] = ("lm_head", "gate."),
                 static_act: bool = False) -> None:
        self.online, self.per_channel, self.ignored, self.static_act = online, per_channel, ignored, static_act

    def get_linear_method(self, layer, prefix):
        if any(s in prefix for s in self.ignored):
            return None
        return Fp8LinearMethod(self.online, self.per_channel, self.static_act)


class Int8LinearMethod(LinearMethod):
    """W8A8 int8: symmetric per-output-channel (or per-tensor) int8 weights, dynamic per-token (or static per-tensor)
    int8 activations, int32 accumulation, scales in the epilogue (reference: quantization/compressed_tensors/schemes/
    compressed_tensors_w8a8_int8.py, kernels/linear/scaled_mm cutlass int8 path).  bf16 checkpoints are quantised
    after loading when ``online``."""

    def __init__(self, online: bool, per_channel: bool = True, static_act: bool = False) -> None:
        self.online, self.per_channel, self.static_act = online, per_channel, static_act and not online

    def create_weights(self, layer, in_per_rank, out_per_rank, dtype) -> None:
        layer.weight = nn.Parameter(torch.empty(out_per_rank, in_per_rank, dtype=dtype if self.online else torch.int8),
                                    requires_grad=False)
        layer.weight_scale = nn.Parameter(torch.ones(out_per_rank if self.per_channel else 1, 1, dtype=torch.float32),
                                          requires_grad=False)
        if self.static_act:
            layer.input_scale = nn.Parameter(torch.ones(1, dtype=torch.float32), requires_grad=False)
        layer.out_dtype = dtype

    def process_weights_after_loading(self, layer) -> None:
        if layer.weight.dtype == torch.int8:
            return
        w = layer.weight.data.float()
        amax = (w.abs().amax(dim=1, keepdim=True) if self.per_channel else w.abs().amax().reshape(1, 1)).clamp_min(1e-12)
        scale = amax / 127.0
        layer.weight = nn.Parameter((w / scale).round().clamp(-128, 127).to(torch.int8), requires_grad=False)
        layer.weight_scale = nn.Parameter(scale.float(), requires_grad=False)

    def apply(self, layer, x, bias):
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        xq, a_scale = ops.scaled_int8_quant(x2, layer.input_scale.float().max().reshape(1)
                                            if getattr(layer, "input_scale", None) is not None else None)
        if x2.is_cuda and hasattr(torch, "_int_mm") and x2.shape[0] > 16 and x2.shape[0] % 8 == 0 and x2.shape[1] % 8 == 0 \
                and layer.weight.shape[0] % 8 == 0:
            acc = torch._int_mm(xq, layer.weight.t()).float()                 # int8 tensor-core GEMM (library)
        else:
            acc = xq.float() @ layer.weight.float().t()
        y = (acc * a_scale.reshape(-1, 1) * layer.weight_scale.reshape(1, -1)).to(layer.out_dtype)
        y = y if bias is None else y + bias
        return y.reshape(*shape[:-1], -1)


class Int8Config(QuantConfig):
    name = "int8"

    def __init__(self, online: bool, per_channel: bool = True, ignored: tuple[str, ...] = ("lm_head", "gate."),
                 static_act: bool = False) -> None:
        self.online, self.per_channel, self.ignored, self.static_act = online, per_channel, ignored, static_act

    def get_linear_method(self, layer, prefix):
        if any(s in prefix for s in self.ignored):
            return None
        return Int8LinearMethod(self.online, self.per_channel, self.static_act)


class BlockFloatLinearMethod(LinearMethod):
    """Online micro-scaled weight formats: the bf16 checkpoint is quantised after loading and kept packed —
    ``nvfp4`` (e2m1 codes, fp8 scale per 16 values + one fp32 scale per tensor) or ``mxfp8`` (e4m3 codes, power-of-two
    e8m0 scale per 32 values).  The GEMM consumes the weights expanded per call for now; the packed layout is the one
    the tcgen05 ``kind::mxf4nvf4`` / ``kind::mxf8f6f4`` block-scaled MMAs take directly
    (reference: quantization/ online modes nvfp4 / mxfp8, modelopt.py, mxfp8 utils)."""

    def __init__(self, fmt: str) -> None:
        assert fmt in ("nvfp4", "mxfp8", "mxfp4")
        self.fmt = fmt

    def create_weights(self, layer, in_per_rank, out_per_rank, dtype) -> None:
        layer.weight = nn.Parameter(torch.empty(out_per_rank, in_per_rank, dtype=dtype), requires_grad=False)
        layer.out_dtype = dtype

    def process_weights_after_loading(self, layer) -> None:
        if getattr(layer, "weight_packed", None) is not None:
            return
        w = layer.weight.data.float()
        if self.fmt == "nvfp4":
            K = w.shape[1]
            pad = (-K) % 16
            if pad:
                w = torch.nn.functional.pad(w, (0, pad))
            packed, bscale, gscale = ref.fp4_quant(w, 16)
            layer.weight_packed = nn.Parameter(packed, requires_grad=False)
            layer.weight_scale = nn.Parameter(bscale, requires_grad=False)
            layer.weight_global_scale = float(gscale)
        elif self.fmt == "mxfp4":
            K = w.shape[1]
            pad = (-K) % 32
            if pad:
                w = torch.nn.functional.pad(w, (0, pad))
            packed, sc = ref.mxfp4_quant(w)
            layer.weight_packed = nn.Parameter(packed, requires_grad=False)
            layer.weight_scale = nn.Parameter(sc, requires_grad=False)
            layer.weight_global_scale = 1.0
        else:
            K = w.shape[1]
            pad = (-K) % 32
            if pad:
                w = torch.nn.functional.pad(w, (0, pad))
            g = w.reshape(w.shape[0], -1, 32)
            e = torch.ceil(torch.log2(g.abs().amax(-1).clamp_min(2.0 ** -120) / ref.FP8_MAX)).clamp(-127, 127)
            q = (g / torch.exp2(e)[..., None]).clamp(-ref.FP8_MAX, ref.FP8_MAX).to(torch.float8_e4m3fn)
            layer.weight_packed = nn.Parameter(q.reshape(w.shape[0], -1), requires_grad=False)
            layer.weight_scale = nn.Parameter((e + 127).to(torch.uint8), requires_grad=False)          # e8m0
        layer.in_features = K
        layer.out_features = w.shape[0]
        layer.weight = nn.Parameter(torch.empty(0, dtype=layer.out_dtype), requires_grad=False)       # drop the bf16 copy
        # decode path: tile-major copy of the codes + block scales for the W4A16 tcgen05 kernel (csrc/gemm_w4.cu), which
        # streams the 4-bit weights and expands them in shared memory (prefill keeps the expanded-weight GEMM)
        layer.weight_tiled = None
        if self.fmt in ("nvfp4", "mxfp4") and layer.weight_packed.is_cuda and layer.weight_packed.shape[1] * 2 % 64 == 0:
            from vllm_b200.ops import cuda_ops
            layer.weight_tiled = cuda_ops().pack_w4_tiles(layer.weight_packed, layer.weight_scale, self.fmt)

    def _expand(self, layer) -> torch.Tensor:
        if self.fmt == "nvfp4":
            w = ref.fp4_dequant(layer.weight_packed, layer.weight_scale, layer.weight_global_scale, 16)
        elif self.fmt == "mxfp4":
            w = ref.mxfp4_dequant(layer.weight_packed, layer.weight_scale)
        else:
            q = layer.weight_packed.float()
            sc = torch.exp2(layer.weight_scale.float() - 127.0).repeat_interleave(32, dim=1)
            w = q * sc
        return w[:, :layer.in_features].to(layer.out_dtype)

    def apply(self, layer, x, bias):
        if getattr(layer, "weight_packed", None) is None:
            return ops.linear(x, layer.weight, bias)
        tiled = getattr(layer, "weight_tiled", None)
        if tiled is not None and not ops.envs.VLLM_B200_FORCE_REF_OPS:
            x2 = x.reshape(-1, x.shape[-1])
            Kp = layer.weight_packed.shape[1] * 2
            cu = ops.cuda_ops()
            if x2.shape[1] == Kp and cu.gemm_w4_ok(x2, Kp):
                y = cu.gemm_w4(x2, tiled, layer.out_features, layer.weight_global_scale, self.fmt, bias)
                return y.view(*x.shape[:-1], layer.out_features)
        return ops.linear(x, self._expand(layer), bias)


class BlockFloatConfig(QuantConfig):
    def __init__(self, fmt: str, ignored: tuple[str, ...] = ("lm_head", "gate.")) -> None:
        self.name, self.fmt, self.ignored = fmt, fmt, ignored

    def get_linear_method(self, layer, prefix):
        if any(s in prefix for s in self.ignored):
            return None
        return BlockFloatLinearMethod(self.fmt)


def _compressed_tensors_scheme(hq: dict) -> dict:
    """Collapse a compressed-tensors ``quantization_config`` to the one scheme its Linear targets use
    (reference: quantization/compressed_tensors/compressed_tensors.py ``_get_scheme_from_parts``)."""
    groups = hq.get("config_groups") or {}
    for g in groups.values():
        w, a = g.get("weights") or {}, g.get("input_activations") or None
        return dict(w_bits=int(w.get("num_bits", 8)), w_type=w.get("type", "int"), w_strategy=w.get("strategy", "tensor"),
                    w_group=w.get("group_size"), w_symmetric=bool(w.get("symmetric", True)),
                    a_bits=int(a["num_bits"]) if a else None, a_dynamic=bool(a.get("dynamic", False)) if a else None,
                    fmt=g.get("format") or hq.get("format"))
    raise ValueError("compressed-tensors config has no config_groups")


QUANTIZATION_METHODS = ("fp8", "fp8_block", "fbgemm_fp8", "modelopt", "deepseek_v3_fp8", "awq", "gptq", "awq_marlin",
                        "gptq_marlin", "nvfp4", "modelopt_fp4", "int8", "w8a8", "compressed-tensors", "compressed_tensors",
                        "bitsandbytes", "mxfp8", "nvfp4_online", "mxfp4")


def get_quant_config(model_config) -> Optional[QuantConfig]:
    q = model_config.quantization
    if q is None:
        return None
    if q == "fp8_block":
        online = not model_config.hf_config.get("quantization_config")
        return Fp8BlockConfig(online=online)
    hq = model_config.hf_config.get("quantization_config") or {}
    if q in ("awq", "gptq", "awq_marlin", "gptq_marlin", "modelopt_fp4", "bitsandbytes") or (q == "nvfp4" and hq):
        return None          # expanded to the model dtype at load time (model_loader/weight_only.py)
    if q in ("fp8", "fbgemm_fp8", "modelopt"):
        if hq.get("weight_block_size"):          # a "fp8" checkpoint with block scales (DeepSeek-V3 style)
            return Fp8BlockConfig(online=False)
        return Fp8Config(online=not hq, per_channel=(q == "fbgemm_fp8" or hq.get("strategy") == "channel"))
    if q in ("int8", "w8a8"):
        return Int8Config(online=not hq)
    if q in ("mxfp8", "mxfp4") or (q in ("nvfp4", "nvfp4_online") and not hq):
        return BlockFloatConfig(q if q in ("mxfp8", "mxfp4") else "nvfp4")    # online: quantise the loaded bf16 weights
    if q in ("compressed-tensors", "compressed_tensors"):
        sc = _compressed_tensors_scheme(hq)
        ignored = tuple(i.replace("re:", "") for i in (hq.get("ignore") or ["lm_head"])) + ("gate.",)
        if sc["a_bits"] is None:
            return None          # weight-only (W4A16 / W8A16 pack-quantized): expanded at load (model_loader/weight_only.py)
        if sc["w_type"] == "float" and sc["w_bits"] == 8:
            if sc["w_strategy"] == "block":
                return Fp8BlockConfig(online=False)
            return Fp8Config(online=False, per_channel=sc["w_strategy"] == "channel", ignored=ignored,
                             static_act=not sc["a_dynamic"])
        if sc["w_type"] == "int" and sc["w_bits"] == 8 and sc["a_bits"] == 8:
            return Int8Config(online=False, per_channel=sc["w_strategy"] == "channel", ignored=ignored,
                              static_act=not sc["a_dynamic"])
        raise ValueError(f"unsupported compressed-tensors scheme {sc}")
    if q == "deepseek_v3_fp8":
        return Fp8BlockConfig(online=not hq)
    raise ValueError(f"quantization method {q!r} is not supported yet")
