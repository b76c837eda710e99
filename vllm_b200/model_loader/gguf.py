"""GGUF checkpoints (llama.cpp): container parser, block de-quantisers, tensor-name mapping, config from metadata.

Reference: vllm/model_executor/model_loader/gguf_loader.py + layers/quantization/gguf.py (which keep the blocks packed
and call ggml CUDA kernels).  Here the blocks are expanded to the model dtype while the file is streamed — the same
policy as the other weight-only formats (``weight_only.py``): every parallelism mode and kernel applies unchanged.

Supported tensor types: F32, F16, BF16, Q8_0, Q4_0, Q4_1, Q5_0, Q4_K, Q6_K (what ``Q4_0`` / ``Q8_0`` / ``Q4_K_M`` files
contain).  Architectures: the llama family (llama, mistral, qwen2, qwen3, gemma) name scheme.
"""
from __future__ import annotations

import struct
from typing import Any, BinaryIO, Iterator, Optional

import numpy as np
import torch

MAGIC = 0x46554747
F32, F16, Q4_0, Q4_1, Q5_0, Q5_1, Q8_0, Q4_K, Q6_K, BF16 = 0, 1, 2, 3, 6, 7, 8, 12, 14, 30
# type → (elements per block, bytes per block)
BLOCK = {F32: (1, 4), F16: (1, 2), BF16: (1, 2), Q4_0: (32, 18), Q4_1: (32, 20), Q5_0: (32, 22), Q8_0: (32, 34),
         Q4_K: (256, 144), Q6_K: (256, 210)}
_SCALARS = {0: "<B", 1: "<b", 2: "<H", 3: "<h", 4: "<I", 5: "<i", 6: "<f", 7: "<?", 10: "<Q", 11: "<q", 12: "<d"}


def _read_str(f: BinaryIO) -> str:
    (n,) = struct.unpack("<Q", f.read(8))
    return f.read(n).decode("utf-8", errors="replace")


def _read_value(f: BinaryIO, t: int) -> Any:
    if t in _SCALARS:
        fmt = _SCALARS[t]
        return struct.unpack(fmt, f.read(struct.calcsize(fmt)))[0]
    if t == 8:
        return _read_str(f)
    if t == 9:
        (et,) = struct.unpack("<I", f.read(4))
        (n,) = struct.unpack("<Q", f.read(8))
        if et in _SCALARS and et != 7:
            fmt = _SCALARS[et]
            sz = struct.calcsize(fmt)
            return np.frombuffer(f.read(sz * n), dtype=np.dtype(fmt)).tolist() if n < (1 << 16) else f.seek(sz * n, 1) and None
        return [_read_value(f, et) for _ in range(n)]
    raise ValueError(f"unknown GGUF value type {t}")


class GGUFFile:
    def __init__(self, path: str) -> None:
        self.path = path
        with open(path, "rb") as f:
            magic, self.version = struct.unpack("<II", f.read(8))
            if magic != MAGIC:
                raise ValueError(f"{path} is not a GGUF file")
            n_tensors, n_kv = struct.unpack("<QQ", f.read(16))
            self.metadata: dict[str, Any] = {}
            for _ in range(n_kv):
                k = _read_str(f)
                (t,) = struct.unpack("<I", f.read(4))
                self.metadata[k] = _read_value(f, t)
            self.tensors: list[tuple[str, tuple[int, ...], int, int]] = []
            for _ in range(n_tensors):
                name = _read_str(f)
                (nd,) = struct.unpack("<I", f.read(4))
                dims = struct.unpack(f"<{nd}Q", f.read(8 * nd))
                t, off = struct.unpack("<IQ", f.read(12))
                self.tensors.append((name, tuple(reversed(dims)), t, off))      # ggml lists the fastest dim first
            align = int(self.metadata.get("general.alignment", 32))
            self.data_start = (f.tell() + align - 1) // align * align

    def read(self, shape: tuple[int, ...], t: int, off: int) -> torch.Tensor:
        n = int(np.prod(shape))
        per, nbytes = BLOCK[t]
        raw = np.fromfile(self.path, dtype=np.uint8, count=n // per * nbytes, offset=self.data_start + off)
        return dequantize(raw, t, n).reshape(shape)


def _f16(b: np.ndarray) -> np.ndarray:
    return b.copy().view(np.float16).astype(np.float32)


def dequantize(raw: np.ndarray, t: int, n: int) -> torch.Tensor:
    if t == F32:
        return torch.from_numpy(raw.view(np.float32).copy())
    if t == F16:
        return torch.from_numpy(raw.view(np.float16).astype(np.float32))
    if t == BF16:
        return torch.from_numpy(raw.view(np.int16).copy()).view(torch.bfloat16).float()
    per, nbytes = BLOCK[t]
    b = raw.reshape(-1, nbytes)
    if t == Q8_0:
        return torch.from_numpy(_f16(b[:, :2]) * b[:, 2:].view(np.int8).astype(np.float32)).reshape(-1)
    if t in (Q4_0, Q4_1):
        d = _f16(b[:, :2])
        qs = b[:, (2 if t == Q4_0 else 4):]
        q = np.concatenate([qs & 0x0F, qs >> 4], axis=1).astype(np.float32)          # low nibbles = first 16 values
        out = d * (q - 8.0) if t == Q4_0 else d * q + _f16(b[:, 2:4])
        return torch.from_numpy(out).reshape(-1)
    if t == Q5_0:
        d = _f16(b[:, :2])
        qh = b[:, 2:6].copy().view(np.uint32)                                        # [nb, 1]
        qs = b[:, 6:]
        hi = ((qh >> np.arange(32, dtype=np.uint32)[None, :]) & 1).astype(np.uint8)  # bit j → 5th bit of value j
        q = np.concatenate([qs & 0x0F, qs >> 4], axis=1) | (hi << 4)
        return torch.from_numpy(d * (q.astype(np.float32) - 16.0)).reshape(-1)
    if t == Q4_K:
        d, dmin = _f16(b[:, 0:2]), _f16(b[:, 2:4])
        sc_raw, qs = b[:, 4:16], b[:, 16:]
        sc, mn = np.empty((b.shape[0], 8), np.float32), np.empty((b.shape[0], 8), np.float32)
        for j in range(8):                                                           # 6-bit scales / mins (get_scale_min_k4)
            if j < 4:
                sc[:, j], mn[:, j] = sc_raw[:, j] & 63, sc_raw[:, j + 4] & 63
            else:
                sc[:, j] = (sc_raw[:, j + 4] & 0x0F) | ((sc_raw[:, j - 4] >> 6) << 4)
                mn[:, j] = (sc_raw[:, j + 4] >> 4) | ((sc_raw[:, j] >> 6) << 4)
        q = qs.reshape(-1, 4, 32)
        vals = np.stack([q & 0x0F, q >> 4], axis=2).reshape(-1, 8, 32).astype(np.float32)   # sub-blocks 2i, 2i+1 share bytes
        out = d[:, :, None] * sc[:, :, None] * vals - dmin[:, :, None] * mn[:, :, None]
        return torch.from_numpy(out).reshape(-1)
    if t == Q6_K:
        ql, qh, scales, d = b[:, :128], b[:, 128:192], b[:, 192:208].view(np.int8).astype(np.float32), _f16(b[:, 208:210])
        out = np.empty((b.shape[0], 256), np.float32)
        for half in range(2):
            l_, h_ = ql[:, half * 64:(half + 1) * 64], qh[:, half * 32:(half + 1) * 32]
            for k, (lsrc, shift_l, shift_h) in enumerate(((l_[:, :32], 0, 0), (l_[:, 32:], 0, 2), (l_[:, :32], 4, 4), (l_[:, 32:], 4, 6))):
                q = (((lsrc >> shift_l) & 0x0F) | (((h_ >> shift_h) & 3) << 4)).astype(np.int8) - 32
                base = half * 128 + k * 32
                s = scales[:, half * 8 + k * 2:half * 8 + k * 2 + 2]                 # one scale per 16 values
                out[:, base:base + 32] = d * np.repeat(s, 16, axis=1) * q
        return torch.from_numpy(out).reshape(-1)
    raise NotImplementedError(f"GGUF tensor type {t}")


# ------------------------------------------------------------------------------------------------ names / config
_LAYER_MAP = {"attn_norm": "input_layernorm", "ffn_norm": "post_attention_layernorm", "attn_q": "self_attn.q_proj",
              "attn_k": "self_attn.k_proj", "attn_v": "self_attn.v_proj", "attn_output": "self_attn.o_proj",
              "ffn_gate": "mlp.gate_proj", "ffn_up": "mlp.up_proj", "ffn_down": "mlp.down_proj",
              "attn_q_norm": "self_attn.q_norm", "attn_k_norm": "self_attn.k_norm"}
_ARCH = {"llama": "LlamaForCausalLM", "mistral": "MistralForCausalLM", "qwen2": "Qwen2ForCausalLM", "qwen3": "Qwen3ForCausalLM",
         "gemma": "GemmaForCausalLM"}


def hf_name(name: str) -> Optional[str]:
    if name == "token_embd.weight":
        return "model.embed_tokens.weight"
    if name == "output_norm.weight":
        return "model.norm.weight"
    if name == "output.weight":
        return "lm_head.weight"
    if name.startswith("blk."):
        _, i, rest = name.split(".", 2)
        stem, _, leaf = rest.rpartition(".")
        if stem in _LAYER_MAP:
            return f"model.layers.{i}.{_LAYER_MAP[stem]}.{leaf}"
    return None


def _unpermute(w: torch.Tensor, n_head: int) -> torch.Tensor:
    """Undo llama.cpp's q/k row permutation (it stores rotary pairs interleaved; HF keeps halves)."""
    out = w.shape[0]
    return w.view(n_head, out // n_head // 2, 2, *w.shape[1:]).transpose(1, 2).reshape(w.shape)


def gguf_config(path: str) -> dict[str, Any]:
    g = GGUFFile(path)
    md = g.metadata
    arch = md.get("general.architecture", "llama")
    if arch not in _ARCH:
        raise ValueError(f"GGUF architecture {arch!r} is not mapped (supported: {sorted(_ARCH)})")
    k = lambda s, d=None: md.get(f"{arch}.{s}", d)  # noqa: E731
    nh = int(k("attention.head_count"))
    emb = next((shape for n, shape, _, _ in g.tensors if n == "token_embd.weight"), None)
    cfg = dict(architectures=[_ARCH[arch]], model_type=arch, hidden_size=int(k("embedding_length")),
               intermediate_size=int(k("feed_forward_length")), num_hidden_layers=int(k("block_count")),
               num_attention_heads=nh, num_key_value_heads=int(k("attention.head_count_kv", nh)),
               rms_norm_eps=float(k("attention.layer_norm_rms_epsilon", 1e-5)), rope_theta=float(k("rope.freq_base", 10000.0)),
               max_position_embeddings=int(k("context_length", 4096)),
               vocab_size=int(k("vocab_size", emb[0] if emb else len(md.get("tokenizer.ggml.tokens") or []) or 32000)),
               hidden_act="silu", torch_dtype="bfloat16",
               tie_word_embeddings=not any(n == "output.weight" for n, *_ in g.tensors),
               bos_token_id=md.get("tokenizer.ggml.bos_token_id"), eos_token_id=md.get("tokenizer.ggml.eos_token_id"))
    if k("attention.key_length"):
        cfg["head_dim"] = int(k("attention.key_length"))
    return cfg


def iter_gguf_weights(path: str, dtype: torch.dtype) -> Iterator[tuple[str, torch.Tensor]]:
    g = GGUFFile(path)
    arch = g.metadata.get("general.architecture", "llama")
    nh = int(g.metadata.get(f"{arch}.attention.head_count", 1))
    nkv = int(g.metadata.get(f"{arch}.attention.head_count_kv", nh))
    for name, shape, t, off in g.tensors:
        hn = hf_name(name)
        if hn is None:
            continue
        w = g.read(shape, t, off)
        if arch in ("llama", "mistral") and hn.endswith("q_proj.weight"):
            w = _unpermute(w, nh)
        elif arch in ("llama", "mistral") and hn.endswith("k_proj.weight"):
            w = _unpermute(w, nkv)
        yield hn, w.to(dtype)


# ------------------------------------------------------------------------------------------------ writer (tests, export)
def quantize(w: torch.Tensor, t: int) -> np.ndarray:
    x = w.detach().float().numpy().reshape(-1)
    if t == F32:
        return x.astype(np.float32).view(np.uint8)
    if t == F16:
        return x.astype(np.float16).view(np.uint8)
    b = x.reshape(-1, 32)
    if t == Q8_0:
        d = np.abs(b).max(1, keepdims=True) / 127.0
        q = np.round(b / np.where(d == 0, 1, d)).clip(-128, 127).astype(np.int8)
        return np.concatenate([d.astype(np.float16).view(np.uint8), q.view(np.uint8)], axis=1).reshape(-1)
    if t == Q4_0:
        idx = np.abs(b).argmax(1)
        mx = b[np.arange(b.shape[0]), idx][:, None]
        d = mx / -8.0
        q = np.round(b / np.where(d == 0, 1, d) + 8.0).clip(0, 15).astype(np.uint8)
        packed = q[:, :16] | (q[:, 16:] << 4)
        return np.concatenate([d.astype(np.float16).view(np.uint8), packed], axis=1).reshape(-1)
    raise NotImplementedError(t)


def write_gguf(path: str, arch: str, metadata: dict[str, Any], tensors: list[tuple[str, torch.Tensor, int]]) -> None:
    def wstr(f, s):
        b = s.encode()
        f.write(struct.pack("<Q", len(b)) + b)
    with open(path, "wb") as f:
        md = {"general.architecture": arch, "general.alignment": 32, **metadata}
        f.write(struct.pack("<IIQQ", MAGIC, 3, len(tensors), len(md)))
        for k, v in md.items():
            wstr(f, k)
            if isinstance(v, str):
                f.write(struct.pack("<I", 8)); wstr(f, v)
            elif isinstance(v, float):
                f.write(struct.pack("<If", 6, v))
            else:
                f.write(struct.pack("<II", 4, int(v)))
        blobs, off = [], 0
        for name, w, t in tensors:
            raw = quantize(w, t)
            wstr(f, name)
            dims = tuple(reversed(w.shape))
            f.write(struct.pack("<I", len(dims)) + struct.pack(f"<{len(dims)}Q", *dims) + struct.pack("<IQ", t, off))
            blobs.append(raw)
            off += (raw.nbytes + 31) // 32 * 32
        pos = f.tell()
        f.write(b"\0" * ((pos + 31) // 32 * 32 - pos))
        for raw in blobs:
            f.write(raw.tobytes())
            f.write(b"\0" * ((raw.nbytes + 31) // 32 * 32 - raw.nbytes))
