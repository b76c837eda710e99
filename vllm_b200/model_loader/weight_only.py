"""Weight-only int4 checkpoints (AWQ, GPTQ).

Reference: vllm/model_executor/layers/quantization/{awq,gptq}.py (+ Marlin/Machete kernels that consume the
packed tensors directly).  Here the packed tensors are expanded while the checkpoint is streamed: every
``<module>.{qweight,qzeros,scales[,g_idx]}`` group becomes one ``<module>.weight`` in the model dtype and then flows
through the ordinary sharding / fusing weight loaders, so these checkpoints work with every parallelism mode and
every kernel.  (The model then occupies bf16 memory — 180 GB of HBM makes that viable for ≤70B models; a packed
int4 tcgen05 path is a later optimisation, not a functional requirement.)
"""
from __future__ import annotations

from typing import Iterable, Iterator, Optional

import torch

AWQ_ORDER = [0, 2, 4, 6, 1, 3, 5, 7]          # nibble i of a packed int32 holds column AWQ_ORDER[i]... (reverse map below)
_AWQ_REVERSE = [0, 4, 1, 5, 2, 6, 3, 7]


def _unpack_cols(x: torch.Tensor, bits: int = 4) -> torch.Tensor:
    """int32 [r, c] → [r, c * 32/bits] (nibble j of each word → column 8*i + j)."""
    n = 32 // bits
    shifts = torch.arange(n, device=x.device, dtype=torch.int32) * bits
    return ((x.unsqueeze(-1) >> shifts) & ((1 << bits) - 1)).reshape(x.shape[0], -1)


def _unpack_rows(x: torch.Tensor, bits: int = 4) -> torch.Tensor:
    """int32 [r, c] → [r * 32/bits, c] (nibble j of each word → row 8*i + j)."""
    n = 32 // bits
    shifts = torch.arange(n, device=x.device, dtype=torch.int32) * bits
    return ((x.unsqueeze(1) >> shifts[None, :, None]) & ((1 << bits) - 1)).reshape(-1, x.shape[1])


def dequant_awq(qweight: torch.Tensor, qzeros: torch.Tensor, scales: torch.Tensor, group_size: int) -> torch.Tensor:
    """AWQ GEMM layout: qweight [in, out/8], qzeros [in/g, out/8], scales [in/g, out] → weight [out, in] (fp32)."""
    q = _unpack_cols(qweight)
    z = _unpack_cols(qzeros)
    out = q.shape[1]
    perm = torch.tensor(_AWQ_REVERSE, device=q.device).repeat(out // 8) + torch.arange(0, out, 8, device=q.device).repeat_interleave(8)
    q, z = q[:, perm], z[:, perm]
    g = torch.arange(q.shape[0], device=q.device) // group_size
    w = (q.float() - z[g].float()) * scales[g].float()
    return w.t().contiguous()


def dequant_gptq(qweight: torch.Tensor, qzeros: torch.Tensor, scales: torch.Tensor, g_idx: Optional[torch.Tensor],
                 group_size: int, bits: int = 4, zero_offset: int = 1) -> torch.Tensor:
    """GPTQ layout: qweight [in/8, out], qzeros [in/g, out/8], scales [in/g, out], g_idx [in] → weight [out, in]."""
    q = _unpack_rows(qweight, bits)
    z = _unpack_cols(qzeros, bits) + zero_offset
    g = g_idx.long() if g_idx is not None else torch.arange(q.shape[0], device=q.device) // group_size
    w = (q.float() - z[g].float()) * scales[g].float()
    return w.t().contiguous()


def pack_awq(w: torch.Tensor, group_size: int):
    """Reference quantiser + packer (tests, synthetic checkpoints). w [out, in] → (qweight, qzeros, scales)."""
    wt = w.t().float()                                   # [in, out]
    I, O = wt.shape
    gr = wt.reshape(I // group_size, group_size, O)
    mn, mx = gr.amin(1), gr.amax(1)
    scale = ((mx - mn).clamp_min(1e-8) / 15.0)
    zero = (-mn / scale).round().clamp(0, 15)
    q = (gr / scale[:, None] + zero[:, None]).round().clamp(0, 15).reshape(I, O).to(torch.int32)

    def pack(x):
        x = x.reshape(x.shape[0], -1, 8)[:, :, AWQ_ORDER]
        out = torch.zeros(x.shape[0], x.shape[1], dtype=torch.int32)
        for j in range(8):
            out |= (x[:, :, j].to(torch.int32) & 0xF) << (4 * j)
        return out
    return pack(q), pack(zero.to(torch.int32)), scale.to(torch.float16)


def pack_gptq(w: torch.Tensor, group_size: int):
    wt = w.t().float()
    I, O = wt.shape
    gr = wt.reshape(I // group_size, group_size, O)
    mn, mx = gr.amin(1), gr.amax(1)
    scale = ((mx - mn).clamp_min(1e-8) / 15.0)
    zero = (-mn / scale).round().clamp(1, 15)            # stored as zero-1 ≥ 0
    q = (gr / scale[:, None] + zero[:, None]).round().clamp(0, 15).reshape(I, O).to(torch.int32)
    qw = torch.zeros(I // 8, O, dtype=torch.int32)
    for j in range(8):
        qw |= (q[j::8] & 0xF) << (4 * j)
    zz = (zero.to(torch.int32) - 1).reshape(zero.shape[0], -1, 8)
    qz = torch.zeros(zero.shape[0], O // 8, dtype=torch.int32)
    for j in range(8):
        qz |= (zz[:, :, j] & 0xF) << (4 * j)
    g_idx = torch.arange(I, dtype=torch.int32) // group_size
    return qw, qz, scale.to(torch.float16), g_idx


def dequant_nvfp4(weight: torch.Tensor, weight_scale: torch.Tensor, weight_scale_2) -> torch.Tensor:
    """ModelOpt NVFP4: ``weight`` uint8 [out, in/2], ``weight_scale`` fp8 [out, in/16], ``weight_scale_2`` fp32 scalar."""
    from vllm_b200.ops import ref
    return ref.fp4_dequant(weight, weight_scale, float(weight_scale_2), 16)


def dequant_compressed_tensors(packed: torch.Tensor, scale: torch.Tensor, zero_point: Optional[torch.Tensor],
                               shape, bits: int, group_size: Optional[int], symmetric: bool) -> torch.Tensor:
    """compressed-tensors ``pack-quantized``: ``weight_packed`` int32 [out, in*bits/32] (value j of a word → column
    8*i + j, stored with a +2^(bits-1) offset when symmetric), ``weight_scale`` [out, in/g] (or [out, 1]),
    optional ``weight_zero_point`` packed along the output dim → weight [out, in]."""
    out_f, in_f = int(shape[0]), int(shape[1])
    q = _unpack_cols(packed, bits)[:, :in_f].float()
    if symmetric or zero_point is None:
        q = q - float(1 << (bits - 1))
    else:
        zp = _unpack_rows(zero_point, bits)[:out_f].float() if zero_point.dtype == torch.int32 else zero_point.float()
        g = group_size or in_f
        q = q - zp.repeat_interleave(g, dim=1)[:, :in_f]
    sc = scale.float()
    if sc.shape[1] != in_f:
        sc = sc.repeat_interleave(group_size or in_f, dim=1)[:, :in_f]
    return q * sc


def pack_compressed_tensors(w: torch.Tensor, bits: int = 4, group_size: int = 32):
    """Reference symmetric group quantiser + packer (tests / synthetic checkpoints)."""
    O, I = w.shape
    g = w.float().reshape(O, I // group_size, group_size)
    qmax = float((1 << (bits - 1)) - 1)
    scale = (g.abs().amax(-1) / qmax).clamp_min(1e-8)
    q = (g / scale[..., None]).round().clamp(-qmax - 1, qmax).reshape(O, I).to(torch.int32) + (1 << (bits - 1))
    n = 32 // bits
    packed = torch.zeros(O, I // n, dtype=torch.int32)
    for j in range(n):
        packed |= (q[:, j::n] & ((1 << bits) - 1)) << (bits * j)
    return packed, scale.to(torch.float16), torch.tensor([O, I], dtype=torch.int64)


NF4_TABLE = [-1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635, -0.18477343022823334,
             -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725, 0.24611230194568634, 0.33791524171829224,
             0.44070982933044434, 0.5626170039176941, 0.7229568362236023, 1.0]


def dequant_bnb4(packed: torch.Tensor, absmax: torch.Tensor, quant_map: torch.Tensor, shape, blocksize: int = 64,
                 nested_absmax: Optional[torch.Tensor] = None, nested_quant_map: Optional[torch.Tensor] = None,
                 nested_blocksize: int = 256, offset: float = 0.0) -> torch.Tensor:
    """bitsandbytes 4-bit (NF4 / FP4) pre-quantised weights: two codes per byte (high nibble first), one absmax per
    ``blocksize`` values; with double quantisation the absmax vector is itself 8-bit block-quantised around ``offset``
    (reference: quantization/bitsandbytes.py + model_loader/bitsandbytes_loader.py)."""
    if nested_absmax is not None:
        am = nested_quant_map.float()[absmax.long().reshape(-1)]
        am = am * nested_absmax.float().repeat_interleave(nested_blocksize)[:am.numel()] + offset
    else:
        am = absmax.float().reshape(-1)
    b = packed.reshape(-1)
    codes = torch.stack([b >> 4, b & 0x0F], dim=-1).reshape(-1).long()
    n = int(shape[0]) * int(shape[1])
    vals = quant_map.float()[codes[:n]]
    return (vals * am.repeat_interleave(blocksize)[:n]).reshape(int(shape[0]), int(shape[1]))


def pack_bnb4(w: torch.Tensor, blocksize: int = 64, table=None):
    """Reference NF4 quantiser (tests / synthetic checkpoints) → (packed uint8 [n/2, 1], absmax, quant_map)."""
    qm = torch.tensor(table or NF4_TABLE, dtype=torch.float32)
    x = w.float().reshape(-1, blocksize)
    am = x.abs().amax(1).clamp_min(1e-12)
    codes = ((x / am[:, None])[..., None] - qm).abs().argmin(-1).reshape(-1).to(torch.uint8)
    packed = (codes[0::2] << 4) | codes[1::2]
    return packed.reshape(-1, 1), am, qm


def dequantize_stream(weights: Iterable[tuple[str, torch.Tensor]], quant_cfg: dict, dtype: torch.dtype
                      ) -> Iterator[tuple[str, torch.Tensor]]:
    """Wrap a checkpoint iterator: packed groups are replaced by ``<module>.weight`` tensors as soon as complete."""
    method = (quant_cfg.get("quant_method") or "").lower()
    gs = int(quant_cfg.get("group_size", quant_cfg.get("q_group_size", 128)))
    bits = int(quant_cfg.get("bits", quant_cfg.get("w_bit", 4)))
    need = {"qweight", "qzeros", "scales"} | ({"g_idx"} if method == "gptq" else set())
    pending: dict[str, dict[str, torch.Tensor]] = {}
    if method in ("compressed-tensors", "compressed_tensors"):
        grp0 = next(iter((quant_cfg.get("config_groups") or {}).values()))["weights"]
        cbits, cgs, csym = int(grp0.get("num_bits", 4)), grp0.get("group_size"), bool(grp0.get("symmetric", True))
        cneed = {"weight_packed", "weight_scale", "weight_shape"} | (set() if csym else {"weight_zero_point"})
        for name, t in weights:
            mod, _, leaf = name.rpartition(".")
            if leaf not in ("weight_packed", "weight_scale", "weight_shape", "weight_zero_point", "weight_g_idx"):
                yield name, t
                continue
            grp = pending.setdefault(mod, {})
            grp[leaf] = t
            if cneed <= set(grp):
                w = dequant_compressed_tensors(grp["weight_packed"], grp["weight_scale"], grp.get("weight_zero_point"),
                                               grp["weight_shape"].tolist(), cbits, cgs, csym)
                del pending[mod]
                yield f"{mod}.weight", w.to(dtype)
        if pending:
            raise ValueError(f"incomplete compressed-tensors groups for {sorted(pending)[:3]} ...")
        return
    if method == "bitsandbytes":
        import json as _json
        for name, t in weights:
            base, sep, leaf = name.partition(".weight")
            if not sep or (leaf == "" and t.dtype != torch.uint8):
                yield name, t
                continue
            grp = pending.setdefault(base, {})
            grp[leaf or "weight"] = t
            qs_key = next((k for k in grp if k.startswith(".quant_state.bitsandbytes__")), None)
            if "weight" in grp and ".absmax" in grp and ".quant_map" in grp and qs_key:
                st = _json.loads(bytes(grp[qs_key].reshape(-1).tolist()).decode())
                nested = ".nested_absmax" in grp
                if st.get("nested_blocksize") and not nested:
                    continue                                  # double-quantised: wait for the nested tensors
                w = dequant_bnb4(grp["weight"], grp[".absmax"], grp[".quant_map"], st["shape"], int(st.get("blocksize", 64)),
                                 grp.get(".nested_absmax"), grp.get(".nested_quant_map"), int(st.get("nested_blocksize") or 256),
                                 float(st.get("nested_offset") or 0.0))
                del pending[base]
                yield f"{base}.weight", w.to(dtype)
        if pending:
            raise ValueError(f"incomplete bitsandbytes groups for {sorted(pending)[:3]} ...")
        return
    if method in ("modelopt_fp4", "nvfp4", "modelopt"):
        # NVFP4: weight (uint8) + weight_scale (fp8) + weight_scale_2 (fp32); input_scale is an activation statistic
        for name, t in weights:
            mod, _, leaf = name.rpartition(".")
            if leaf == "input_scale":
                continue
            if leaf not in ("weight", "weight_scale", "weight_scale_2") or (leaf == "weight" and t.dtype != torch.uint8):
                yield name, t
                continue
            grp = pending.setdefault(mod, {})
            grp[leaf] = t
            if {"weight", "weight_scale", "weight_scale_2"} <= set(grp):
                yield f"{mod}.weight", dequant_nvfp4(grp["weight"], grp["weight_scale"], grp["weight_scale_2"]).to(dtype)
                del pending[mod]
        if pending:
            raise ValueError(f"incomplete NVFP4 tensors for {sorted(pending)[:3]} ...")
        return
    for name, t in weights:
        mod, _, leaf = name.rpartition(".")
        if leaf not in ("qweight", "qzeros", "scales", "g_idx"):
            yield name, t
            continue
        grp = pending.setdefault(mod, {})
        grp[leaf] = t
        if need <= set(grp):
            if method == "awq":
                w = dequant_awq(grp["qweight"], grp["qzeros"], grp["scales"], gs)
            else:
                gsz = gs if gs > 0 else grp["qweight"].shape[0] * (32 // bits)
                w = dequant_gptq(grp["qweight"], grp["qzeros"], grp["scales"], grp.get("g_idx"), gsz, bits,
                                 0 if quant_cfg.get("checkpoint_format") == "gptq_v2" else 1)
            del pending[mod]
            yield f"{mod}.weight", w.to(dtype)
    if pending:
        raise ValueError(f"incomplete quantised tensors for {sorted(pending)[:3]} ...")
