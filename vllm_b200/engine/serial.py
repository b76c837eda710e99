"""msgpack (msgspec) wire codec for the engine-core IPC channel.

Reference: vllm/v1/serial_utils.py:136,313 (``MsgpackEncoder/Decoder`` with
out-of-band tensor buffers).  Tensors (pooling outputs, prompt embeds) are
shipped as (dtype, shape, raw bytes) ext payloads.
"""
from __future__ import annotations

from typing import Any

import msgspec
import numpy as np
import torch

from vllm_b200.core.request import EngineCoreOutputs, EngineCoreRequest

_TENSOR_EXT = 1


def _enc_hook(obj: Any) -> Any:
    if isinstance(obj, torch.Tensor):
        t = obj.detach().cpu().contiguous()
        if t.dtype == torch.bfloat16:
            raw, dt = t.view(torch.int16).numpy().tobytes(), "bfloat16"
        else:
            raw, dt = t.numpy().tobytes(), str(t.dtype).replace("torch.", "")
        return msgspec.msgpack.Ext(_TENSOR_EXT, msgspec.msgpack.encode((dt, list(t.shape), raw)))
    if isinstance(obj, np.ndarray):
        return obj.tolist()
    if isinstance(obj, (set, frozenset)):
        return sorted(obj)
    raise NotImplementedError(f"cannot serialise {type(obj)}")


def _ext_hook(code: int, data: memoryview) -> Any:
    if code == _TENSOR_EXT:
        dt, shape, raw = msgspec.msgpack.decode(bytes(data))
        if dt == "bfloat16":
            return torch.frombuffer(bytearray(raw), dtype=torch.int16).view(torch.bfloat16).reshape(shape)
        return torch.frombuffer(bytearray(raw), dtype=getattr(torch, dt)).reshape(shape)
    raise NotImplementedError(f"unknown ext code {code}")


_enc = msgspec.msgpack.Encoder(enc_hook=_enc_hook)
_req_dec = msgspec.msgpack.Decoder(EngineCoreRequest, ext_hook=_ext_hook)
_out_dec = msgspec.msgpack.Decoder(tuple[int, EngineCoreOutputs], ext_hook=_ext_hook)
_any_dec = msgspec.msgpack.Decoder(ext_hook=_ext_hook)


def encode_request(r: EngineCoreRequest) -> bytes:
    return _enc.encode(r)


def decode_request(b: bytes) -> EngineCoreRequest:
    return _req_dec.decode(b)


def encode_outputs(item: tuple[int, EngineCoreOutputs]) -> bytes:
    return _enc.encode(item)


def decode_outputs(b: bytes) -> tuple[int, EngineCoreOutputs]:
    return _out_dec.decode(b)


def encode_any(x: Any) -> bytes:
    return _enc.encode(x)


def decode_any(b: bytes) -> Any:
    return _any_dec.decode(b)
