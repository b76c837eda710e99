"""Per-rank safetensors snapshots: save once, reload without re-sharding
(reference: model_loader/sharded_state_loader.py, Worker.save_sharded_state gpu_worker.py:1241)."""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from vllm_b200.parallel.state import get_world_group


def _fname(rank: int) -> str:
    return f"model-rank-{rank}.safetensors"


def save_sharded_state(model: nn.Module, path: str) -> str:
    from safetensors.torch import save_file
    os.makedirs(path, exist_ok=True)
    sd, seen = {}, set()
    for k, v in model.state_dict().items():
        if v.data_ptr() in seen:  # tied weights
            continue
        seen.add(v.data_ptr())
        sd[k] = v.detach().cpu().contiguous()
    fn = os.path.join(path, _fname(max(get_world_group().rank, 0)))
    save_file(sd, fn)
    return fn


def load_sharded_state(model: nn.Module, path: str) -> None:
    from safetensors import safe_open
    fn = os.path.join(path, _fname(max(get_world_group().rank, 0)))
    sd = model.state_dict()
    with safe_open(fn, framework="pt", device="cpu") as f:
        for k in f.keys():
            if k in sd:
                sd[k].copy_(f.get_tensor(k))
