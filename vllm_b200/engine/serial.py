"""msgpack (msgspec) wire codec for the engine-core IPC channel.

Reference: vllm/v1/serial_utils.py:136,313 (``MsgpackEncoder/Decoder`` with
out-of-band tensor buffers).  Two wire forms:

* single frame (``encode_* / decode_*``): tensors travel inline as (dtype, shape, raw bytes) ext payloads;
* multi frame (``OOBEncoder / OOBDecoder``): frame 0 is the msgpack message, every tensor / ndarray of at least
  ``OOB_THRESHOLD`` bytes is replaced by an (dtype, shape, frame index) ext and its storage is appended as an extra frame —
  a ``memoryview`` of the tensor itself, so zmq (``send_multipart(copy=False)``) ships prompt embeddings, multimodal
  features and pooling outputs without a single copy on the sending side, and the receiver wraps the received frame with
  ``torch.frombuffer`` (zero copy again).
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


# ------------------------------------------------------------------ out-of-band (zero-copy) tensor frames
OOB_THRESHOLD = 512          # bytes; smaller tensors stay inline (a frame costs more than the copy)
_OOB_EXT = 2


def _np_view(t: torch.Tensor) -> tuple[np.ndarray, str]:
    t = t.detach()
    if t.device.type != "cpu":
        t = t.cpu()
    t = t.contiguous()
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy(), "bfloat16"
    return t.numpy(), str(t.dtype).replace("torch.", "")


class OOBEncoder:
    """``encode(obj) -> [frame0, buf1, buf2, ...]``; the buffers alias the tensors (keep the tensors alive until sent)."""

    def __init__(self, threshold: int = OOB_THRESHOLD) -> None:
        self.threshold = threshold
        self._aux: list[Any] = []
        self._enc = msgspec.msgpack.Encoder(enc_hook=self._hook)

    def _hook(self, obj: Any) -> Any:
        if isinstance(obj, torch.Tensor):
            arr, dt = _np_view(obj)
            if arr.nbytes < self.threshold:
                return msgspec.msgpack.Ext(_TENSOR_EXT, msgspec.msgpack.encode((dt, list(obj.shape), arr.tobytes())))
            self._aux.append(memoryview(arr).cast("B") if arr.ndim else memoryview(arr.reshape(1)).cast("B"))
            return msgspec.msgpack.Ext(_OOB_EXT, msgspec.msgpack.encode((dt, list(obj.shape), len(self._aux))))
        if isinstance(obj, np.ndarray):
            if obj.nbytes < self.threshold or obj.dtype == object:
                return obj.tolist()
            a = np.ascontiguousarray(obj)
            self._aux.append(memoryview(a).cast("B"))
            return msgspec.msgpack.Ext(_OOB_EXT, msgspec.msgpack.encode(("np:" + a.dtype.str, list(a.shape), len(self._aux))))
        if isinstance(obj, (set, frozenset)):
            return sorted(obj)
        raise NotImplementedError(f"cannot serialise {type(obj)}")

    def encode(self, obj: Any) -> list[Any]:
        self._aux = []
        head = self._enc.encode(obj)
        frames, self._aux = [head, *self._aux], []
        return frames


class OOBDecoder:
    """``decode(frames)`` for what ``OOBEncoder.encode`` produced; ``frames`` may be bytes, memoryviews or zmq Frames."""

    def __init__(self, typ: Any = None) -> None:
        self._frames: list[Any] = []
        self._dec = (msgspec.msgpack.Decoder(typ, ext_hook=self._hook) if typ is not None
                     else msgspec.msgpack.Decoder(ext_hook=self._hook))

    @staticmethod
    def _buf(f: Any) -> Any:
        return f.buffer if hasattr(f, "buffer") else f

    def _hook(self, code: int, data: memoryview) -> Any:
        if code == _OOB_EXT:
            dt, shape, idx = msgspec.msgpack.decode(bytes(data))
            buf = self._buf(self._frames[idx])
            if dt.startswith("np:"):
                return np.frombuffer(buf, dtype=np.dtype(dt[3:])).reshape(shape)
            if len(buf) == 0:
                return torch.empty(shape, dtype=getattr(torch, dt))
            if dt == "bfloat16":
                return torch.frombuffer(buf, dtype=torch.int16).view(torch.bfloat16).reshape(shape)
            return torch.frombuffer(buf, dtype=getattr(torch, dt)).reshape(shape)
        return _ext_hook(code, data)

    def decode(self, frames: Any) -> Any:
        if isinstance(frames, (bytes, bytearray, memoryview)):
            frames = [frames]
        self._frames = list(frames)
        try:
            return self._dec.decode(self._buf(self._frames[0]))
        finally:
            self._frames = []


# multi-frame forms of the two hot messages (requests in, outputs out)
_req_oob_dec = OOBDecoder(EngineCoreRequest)
_out_oob_dec = OOBDecoder(tuple[int, EngineCoreOutputs])
_local = __import__("threading").local()


def _oob_enc() -> OOBEncoder:
    e = getattr(_local, "enc", None)
    if e is None:
        e = _local.enc = OOBEncoder()
    return e


def encode_request_frames(r: EngineCoreRequest) -> list[Any]:
    return _oob_enc().encode(r)


def decode_request_frames(frames: Any) -> EngineCoreRequest:
    return OOBDecoder(EngineCoreRequest).decode(frames) if len(frames) > 1 else _req_oob_dec.decode(frames)


def encode_outputs_frames(item: tuple[int, EngineCoreOutputs]) -> list[Any]:
    return _oob_enc().encode(item)


def decode_outputs_frames(frames: Any) -> tuple[int, EngineCoreOutputs]:
    return OOBDecoder(tuple[int, EngineCoreOutputs]).decode(frames) if len(frames) > 1 else _out_oob_dec.decode(frames)
