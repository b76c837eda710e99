"""Accidental-synchronisation detector (reference: vllm/utils/gpu_sync_debug.py, ``VLLM_GPU_SYNC_CHECK``).

After warm-up the hot step must never block the host on the device (no ``.item()``, ``.cpu()``, boolean masks on
device tensors ...): ``VLLM_GPU_SYNC_CHECK=warn|error`` arms ``torch.cuda.set_sync_debug_mode`` around
``execute_model`` so such a call warns / raises with a stack trace.  The sanctioned device→host read of the
step's sampled tokens happens outside the guarded region (``AsyncStepOutput.get``).
"""
from __future__ import annotations

import contextlib
import functools

import torch

from vllm_b200 import envs

_armed = False


def arm() -> None:
    """Called by the worker once warm-up / CUDA-graph capture is finished."""
    global _armed
    _armed = bool(envs.VLLM_GPU_SYNC_CHECK) and torch.cuda.is_available()


@contextlib.contextmanager
def sync_guard():
    mode = envs.VLLM_GPU_SYNC_CHECK
    if not _armed or mode not in ("warn", "error"):
        yield
        return
    prev = torch.cuda.get_sync_debug_mode()
    torch.cuda.set_sync_debug_mode(mode)
    try:
        yield
    finally:
        torch.cuda.set_sync_debug_mode(prev)


def guarded(fn):
    @functools.wraps(fn)
    def wrapper(*a, **kw):
        with sync_guard():
            return fn(*a, **kw)
    return wrapper


def check_logits_nans(logits: torch.Tensor, req_ids: list[str]) -> dict[str, int]:
    """``VLLM_COMPUTE_NANS_IN_LOGITS``: per-request NaN counts (host sync — debugging aid only);
    ``VLLM_RAISE_ON_LOGIT_NANS`` turns a hit into an error naming the requests."""
    if not (envs.VLLM_COMPUTE_NANS_IN_LOGITS or envs.VLLM_RAISE_ON_LOGIT_NANS):
        return {}
    counts = torch.isnan(logits).sum(dim=-1).tolist()
    bad = {r: int(c) for r, c in zip(req_ids, counts) if c}
    if bad and envs.VLLM_RAISE_ON_LOGIT_NANS:
        raise RuntimeError(f"NaNs in logits of requests {sorted(bad)}")
    return {r: int(c) for r, c in zip(req_ids, counts)}
