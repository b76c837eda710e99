"""Multi-LoRA serving.

Behavioural reference: vllm/lora/ (``LoRAModelManager`` model_manager.py:71, per-layer LoRA wrappers
layers/*.py, Punica SGMV/BGMV ops, LRU adapter cache, packed-module handling for qkv / gate_up).

Design: adapters are attached to the *existing* parallel linear modules (no wrapper modules → checkpoint
names, quantisation methods and the fused TP kernels are untouched when no adapter is active).  Every
``LinearBase`` asks ``apply_lora`` for a delta after its base GEMM; the delta for a batch that mixes adapters is
computed per active slot on the rows that use it (shrink ``x·Aᵀ`` then expand ``·Bᵀ`` — the SGMV
decomposition), with TP sharding mirroring the base layer: column-parallel layers slice B by output rows,
row-parallel layers slice A by input columns and let the layer's all-reduce sum the partial deltas.
"""
from __future__ import annotations

import json
import os
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Any, Optional

import torch

from vllm_b200.core.request import LoRARequest  # noqa: F401  (public re-export)
from vllm_b200.logger import init_logger

logger = init_logger(__name__)

# sub-projection → (packed module suffix, index inside the packed output)
PACKED = {"q_proj": ("qkv_proj", 0), "k_proj": ("qkv_proj", 1), "v_proj": ("qkv_proj", 2),
          "gate_proj": ("gate_up_proj", 0), "up_proj": ("gate_up_proj", 1)}


@dataclass
class LoRAModel:
    """One adapter: target module path (HF naming, e.g. ``model.layers.0.self_attn.q_proj``) → (A [r,in], B [out,r])."""
    lora_id: int
    rank: int
    scaling: float
    weights: dict[str, tuple[torch.Tensor, torch.Tensor]] = field(default_factory=dict)

    @classmethod
    def from_tensors(cls, lora_id: int, tensors: dict[str, torch.Tensor], rank: int, alpha: float) -> "LoRAModel":
        w: dict[str, list] = {}
        for name, t in tensors.items():
            n = name
            for pre in ("base_model.model.",):
                if n.startswith(pre):
                    n = n[len(pre):]
            if ".lora_A" in n:
                w.setdefault(n.split(".lora_A")[0], [None, None])[0] = t
            elif ".lora_B" in n:
                w.setdefault(n.split(".lora_B")[0], [None, None])[1] = t
        ok = {k: (a, b) for k, (a, b) in w.items() if a is not None and b is not None}
        return cls(lora_id, rank, alpha / max(rank, 1), ok)

    @classmethod
    def from_path(cls, lora_id: int, path: str) -> "LoRAModel":
        """PEFT directory: adapter_config.json + adapter_model.safetensors (or .bin)."""
        with open(os.path.join(path, "adapter_config.json")) as f:
            cfg = json.load(f)
        st = os.path.join(path, "adapter_model.safetensors")
        if os.path.exists(st):
            from safetensors.torch import load_file
            tensors = load_file(st)
        else:
            tensors = torch.load(os.path.join(path, "adapter_model.bin"), map_location="cpu", weights_only=True)
        return cls.from_tensors(lora_id, tensors, int(cfg["r"]), float(cfg.get("lora_alpha", cfg["r"])))


@dataclass
class LoRABatch:
    """Per-step mapping installed by the model runner: token index → adapter slot (-1 = base model)."""
    token_slots: torch.Tensor                      # [T] int64 on the compute device
    active: list[int]                              # slots present in this batch
    rows: dict[int, torch.Tensor] = field(default_factory=dict)   # slot → token indices


_CURRENT: Optional[LoRABatch] = None


def set_batch(b: Optional[LoRABatch]) -> None:
    global _CURRENT
    _CURRENT = b


def lora_active() -> bool:
    return _CURRENT is not None


def apply_lora(layer, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """y += Σ_slots scaling · (x[rows] Aᵀ) Bᵀ on the output slice each sub-projection owns."""
    table = getattr(layer, "lora_table", None)
    b = _CURRENT
    if b is None or not table:
        return y
    x2, y2 = x.reshape(-1, x.shape[-1]), y.reshape(-1, y.shape[-1])
    stack = getattr(layer, "lora_stack", None)
    if stack is not None and x2.is_cuda:
        # device path: two BGMV launches over the token → slot table (csrc/lora.cu); no host-side grouping
        from vllm_b200.ops import cuda_ops
        if x2.stride(-1) != 1:
            x2 = x2.contiguous()
        cuda_ops().lora_bgmv(x2, y2, stack["A"], stack["B"], b.token_slots, stack["offs"])
        return y
    for slot in b.active:
        entries = table.get(slot)
        if not entries:
            continue
        rows = b.rows[slot]
        xs = x2.index_select(0, rows)
        for off, n, a, bb, scale in entries:
            d = (xs @ a.t()) @ bb.t()
            y2[:, off:off + n].index_add_(0, rows, (d * scale).to(y2.dtype))
    return y


class LoRAModelManager:
    """Owns the adapter slots of one worker's model: LRU cache of loaded adapters (CPU), up to ``max_loras``
    resident on the device, sliced for this TP rank at activation time."""

    def __init__(self, model: torch.nn.Module, lora_config, device: torch.device, dtype: torch.dtype) -> None:
        from vllm_b200.layers.linear import LinearBase
        self.cfg = lora_config
        self.device, self.dtype = device, dtype
        self.max_loras = lora_config.max_loras
        self.modules = {n: m for n, m in model.named_modules() if isinstance(m, LinearBase)}
        self.cpu_cache: "OrderedDict[int, LoRAModel]" = OrderedDict()
        self.max_cpu = lora_config.max_cpu_loras or max(4 * self.max_loras, 8)
        self.slot_of: "OrderedDict[int, int]" = OrderedDict()      # lora id → slot (LRU order)
        self.registered: dict[int, LoRAModel] = {}

    # ---------------------------------------------------------------- loading
    def add_adapter(self, lora: LoRAModel) -> None:
        if lora.rank > self.cfg.max_lora_rank:
            raise ValueError(f"LoRA rank {lora.rank} exceeds max_lora_rank {self.cfg.max_lora_rank}")
        self.cpu_cache[lora.lora_id] = lora
        self.cpu_cache.move_to_end(lora.lora_id)
        while len(self.cpu_cache) > self.max_cpu:
            old, _ = self.cpu_cache.popitem(last=False)
            if old in self.slot_of:
                self._evict_slot(old)

    def _ensure_loaded(self, req: Any) -> LoRAModel:
        lid = req.lora_int_id
        if lid not in self.cpu_cache:
            self.add_adapter(LoRAModel.from_path(lid, req.lora_path))
        self.cpu_cache.move_to_end(lid)
        return self.cpu_cache[lid]

    def remove_adapter(self, lora_id: int) -> bool:
        if lora_id in self.slot_of:
            self._evict_slot(lora_id)
        return self.cpu_cache.pop(lora_id, None) is not None

    def list_adapters(self) -> list[int]:
        return list(self.cpu_cache)

    # ---------------------------------------------------------------- slots
    def _evict_slot(self, lora_id: int) -> None:
        slot = self.slot_of.pop(lora_id)
        for m in self.modules.values():
            t = getattr(m, "lora_table", None)
            if t:
                t.pop(slot, None)
            st = getattr(m, "lora_stack", None)
            if st is not None:
                st["A"][slot].zero_()
                st["B"][slot].zero_()

    def _stack_for(self, mod) -> dict:
        """Per-module stacked adapter weights for the BGMV kernels: A [slots, n_sub, r_max, K], B [slots, N, r_max]
        (scaling folded into B; empty slots / sub-projections stay zero and contribute nothing)."""
        st = getattr(mod, "lora_stack", None)
        if st is None:
            from vllm_b200.layers.linear import MergedColumnParallelLinear, QKVParallelLinear
            w = mod.weight
            N, K = w.shape[0], w.shape[1]
            if isinstance(mod, QKVParallelLinear):
                sizes = [mod.q_size, mod.kv_size, mod.kv_size]
            elif isinstance(mod, MergedColumnParallelLinear):
                sizes = [s_ // mod.tp_size for s_ in mod.output_sizes]
            else:
                sizes = [N]
            r_max = (self.cfg.max_lora_rank + 7) // 8 * 8
            offs = [sum(sizes[:i + 1]) for i in range(len(sizes) - 1)] + [N] * (4 - len(sizes))
            st = {"A": torch.zeros(self.max_loras, len(sizes), r_max, K, dtype=self.dtype, device=self.device),
                  "B": torch.zeros(self.max_loras, N, r_max, dtype=self.dtype, device=self.device),
                  "offs": (len(sizes), offs[0], offs[1], offs[2]), "starts": [sum(sizes[:i]) for i in range(len(sizes))]}
            mod.lora_stack = st
        return st

    def _slice_for(self, mod, sub_idx: Optional[int], a: torch.Tensor, b: torch.Tensor):
        """(out_offset, out_len, A, B) for this TP rank."""
        from vllm_b200.layers.linear import (MergedColumnParallelLinear, QKVParallelLinear, RowParallelLinear,
                                             ColumnParallelLinear)
        tp_rank, tp_size = getattr(mod, "tp_rank", 0), getattr(mod, "tp_size", 1)
        if isinstance(mod, QKVParallelLinear):
            sizes = [mod.q_size, mod.kv_size, mod.kv_size]
            off = sum(sizes[:sub_idx])
            n = sizes[sub_idx]
            if sub_idx == 0:
                r0 = tp_rank * n
            else:                      # kv heads may be replicated when tp > num_kv_heads
                r0 = (tp_rank // max(mod.num_kv_replicas, 1)) * n
            return off, n, a, b[r0:r0 + n]
        if isinstance(mod, MergedColumnParallelLinear):
            sizes = [s // tp_size for s in mod.output_sizes]
            off, n = sum(sizes[:sub_idx]), sizes[sub_idx]
            return off, n, a, b[tp_rank * n:(tp_rank + 1) * n]
        if isinstance(mod, RowParallelLinear):
            k = mod.in_per_rank
            return 0, b.shape[0], a[:, tp_rank * k:(tp_rank + 1) * k], b
        if isinstance(mod, ColumnParallelLinear):
            n = mod.out_per_rank
            return 0, n, a, b[tp_rank * n:(tp_rank + 1) * n]
        return 0, b.shape[0], a, b

    def activate(self, req: Any) -> int:
        """Make the adapter resident; returns its slot."""
        lid = req.lora_int_id
        if lid in self.slot_of:
            self.slot_of.move_to_end(lid)
            return self.slot_of[lid]
        lora = self._ensure_loaded(req)
        if len(self.slot_of) >= self.max_loras:
            victim = next(iter(self.slot_of))
            self._evict_slot(victim)
        used = set(self.slot_of.values())
        slot = next(i for i in range(self.max_loras) if i not in used)
        self.slot_of[lid] = slot
        n_attached = 0
        for target, (a, b) in lora.weights.items():
            parent, _, leaf = target.rpartition(".")
            sub = None
            name = target
            if leaf in PACKED:
                packed, sub = PACKED[leaf]
                name = f"{parent}.{packed}"
            mod = self.modules.get(name)
            if mod is None:
                continue
            off, n, a2, b2 = self._slice_for(mod, sub, a, b)
            ent = (off, n, a2.to(self.device, self.dtype).contiguous(), b2.to(self.device, self.dtype).contiguous(),
                   lora.scaling)
            if not hasattr(mod, "lora_table") or mod.lora_table is None:
                mod.lora_table = {}
            mod.lora_table.setdefault(slot, []).append(ent)
            if self.device.type == "cuda" and getattr(mod, "weight", None) is not None and mod.weight.dim() == 2 \
                    and mod.weight.dtype == self.dtype and mod.weight.shape[1] % 8 == 0 \
                    and self.dtype in (torch.bfloat16, torch.float16):
                st = self._stack_for(mod)
                j = st["starts"].index(off) if off in st["starts"] else 0
                r = ent[2].shape[0]
                st["A"][slot, j, :r].copy_(ent[2])
                st["B"][slot, off:off + n, :r].copy_(ent[3] * lora.scaling)
            n_attached += 1
        logger.info("LoRA %d → slot %d (%d target modules)", lid, slot, n_attached)
        return slot

    def make_batch(self, lora_requests: list[Any], tokens_per_req: list[int]) -> Optional[LoRABatch]:
        """``lora_requests[i]`` (or None) applies to the next ``tokens_per_req[i]`` tokens of the flattened batch."""
        if not any(r is not None for r in lora_requests):
            return None
        slots = []
        for r, n in zip(lora_requests, tokens_per_req):
            s = self.activate(r) if r is not None else -1
            slots.extend([s] * n)
        ts = torch.tensor(slots, dtype=torch.int64, device=self.device)
        active = sorted({s for s in slots if s >= 0})
        return LoRABatch(ts, active, {s: (ts == s).nonzero(as_tuple=True)[0] for s in active})
