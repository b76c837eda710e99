"""Tensor-parallel linear layers.

Behavioural reference: vllm/model_executor/layers/linear.py
(``ColumnParallelLinear`` :407, ``QKVParallelLinear`` :971 with fused-shard
weight loading :1078-1325, ``MergedColumnParallelLinear`` :730-945,
``RowParallelLinear`` :1510, forward GEMM→all-reduce :1641-1667).

B200-first difference: ``RowParallelLinear.forward_reduce_norm`` is the unit the
decoder layers call — GEMM + TP all-reduce + residual add + RMSNorm as ONE
operation — so the fused sm_100a kernel (tcgen05 GEMM whose epilogue reduces
over NVSwitch and normalises) slots in with no graph rewriting.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from vllm_b200 import ops
from vllm_b200.parallel.state import get_tp_group
from vllm_b200.parallel.utils import divide


def _tp():
    g = get_tp_group()
    return max(g.rank_in_group, 0), g.world_size


class LinearMethod:
    """Unquantised bf16/fp16/fp32 weights."""

    def create_weights(self, layer: nn.Module, in_per_rank: int, out_per_rank: int, dtype: torch.dtype) -> None:
        layer.weight = nn.Parameter(torch.empty(out_per_rank, in_per_rank, dtype=dtype), requires_grad=False)

    def apply(self, layer: nn.Module, x: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
        return ops.linear(x, layer.weight, bias)

    def process_weights_after_loading(self, layer: nn.Module) -> None:
        pass


class LinearBase(nn.Module):
    def __init__(self, input_size: int, output_size: int, bias: bool, dtype: Optional[torch.dtype],
                 quant_config=None, prefix: str = "") -> None:
        super().__init__()
        self.input_size, self.output_size = input_size, output_size
        self.dtype = dtype or torch.get_default_dtype()
        self.prefix = prefix
        self.quant_method: LinearMethod = (quant_config.get_linear_method(self, prefix)
                                           if quant_config is not None else None) or LinearMethod()
        self.has_bias = bias

    def _make(self, in_per_rank: int, out_per_rank: int) -> None:
        self.quant_method.create_weights(self, in_per_rank, out_per_rank, self.dtype)
        if self.has_bias:
            self.bias = nn.Parameter(torch.zeros(out_per_rank, dtype=self.dtype), requires_grad=False)
        else:
            self.register_parameter("bias", None)

    def process_weights_after_loading(self) -> None:
        self.quant_method.process_weights_after_loading(self)

    lora_table = None   # slot → [(out_offset, out_len, A, B, scale)], attached by lora.LoRAModelManager

    def _apply(self, x: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
        y = self.quant_method.apply(self, x, bias)
        if self.lora_table:
            from vllm_b200.lora import apply_lora
            y = apply_lora(self, x, y)
        return y


class ReplicatedLinear(LinearBase):
    def __init__(self, input_size, output_size, bias=False, dtype=None, quant_config=None, prefix="") -> None:
        super().__init__(input_size, output_size, bias, dtype, quant_config, prefix)
        self._make(input_size, output_size)

    def forward(self, x):
        return self._apply(x, self.bias)

    def weight_loader(self, param: nn.Parameter, w: torch.Tensor, shard_id=None) -> None:
        assert param.shape == w.shape, (self.prefix, param.shape, w.shape)
        param.data.copy_(w)


class ColumnParallelLinear(LinearBase):
    """Y = X W^T with W split along the output dim."""

    def __init__(self, input_size, output_size, bias=False, gather_output=False, dtype=None,
                 quant_config=None, prefix="") -> None:
        super().__init__(input_size, output_size, bias, dtype, quant_config, prefix)
        self.tp_rank, self.tp_size = _tp()
        self.gather_output = gather_output
        self.out_per_rank = divide(output_size, self.tp_size)
        self._make(input_size, self.out_per_rank)

    def forward(self, x):
        y = self._apply(x, self.bias)
        if self.gather_output and self.tp_size > 1:
            y = get_tp_group().all_gather(y, -1)
        return y

    def weight_loader(self, param: nn.Parameter, w: torch.Tensor, shard_id=None) -> None:
        out_dim = getattr(param, "output_dim", 0)
        blk = getattr(param, "shard_block", 1)  # block-quant scales are sharded in units of blocks
        n = param.shape[out_dim]
        w = w.narrow(out_dim, self.tp_rank * n, n) if w.shape[out_dim] != n else w
        del blk
        param.data.copy_(w)


def _lora_enabled() -> bool:
    from vllm_b200.config import get_current_vllm_config_or_none
    cfg = get_current_vllm_config_or_none()
    return cfg is not None and getattr(cfg, "lora_config", None) is not None


class MergedColumnParallelLinear(ColumnParallelLinear):
    """Several column-parallel matrices fused along the output dim (gate|up)."""

    def __init__(self, input_size, output_sizes: list[int], bias=False, dtype=None, quant_config=None,
                 prefix="", interleave: bool = False) -> None:
        self.output_sizes = output_sizes
        super().__init__(input_size, sum(output_sizes), bias, False, dtype, quant_config, prefix)
        # gate|up stored as interleaved rows (gate_0, up_0, gate_1, up_1, ...): the gated activation then pairs
        # adjacent accumulator rows of ONE GEMM tile and fuses into the GEMM epilogue (``forward_gated_silu``)
        # (LoRA adapters address gate and up as contiguous output slices: keep the plain layout when they are enabled)
        self.interleave = bool(interleave and len(output_sizes) == 2 and output_sizes[0] == output_sizes[1]
                               and not bias and type(self.quant_method) is LinearMethod
                               and self.dtype in (torch.bfloat16, torch.float16) and not _lora_enabled())

    def forward_gated_silu(self, x: torch.Tensor) -> torch.Tensor:
        assert self.interleave
        if self.lora_table:
            return ops.silu_and_mul_interleaved(self._apply(x, None))
        return ops.gated_silu_linear(x, self.weight)

    def weight_loader(self, param: nn.Parameter, w: torch.Tensor, shard_id=None) -> None:
        if shard_id is not None and self.interleave and param is self.weight:
            n = self.output_sizes[shard_id] // self.tp_size
            param.data[shard_id::2].copy_(w.narrow(0, self.tp_rank * n, n))
            return
        if shard_id is None:  # already fused on disk
            off = 0
            for i, sz in enumerate(self.output_sizes):
                self.weight_loader(param, w.narrow(0, off, sz), i)
                off += sz
            return
        out_dim = getattr(param, "output_dim", 0)
        unit = getattr(param, "shard_block", 1)
        sizes = [-(-s // unit) for s in self.output_sizes]
        off = sum(sizes[:shard_id]) // self.tp_size
        n = sizes[shard_id] // self.tp_size
        src = w.narrow(out_dim, self.tp_rank * n, n)
        param.data.narrow(out_dim, off, n).copy_(src)


class QKVParallelLinear(ColumnParallelLinear):
    """Fused q|k|v projection; KV heads are replicated when ``num_kv_heads < tp``."""

    def __init__(self, hidden_size, head_size, total_num_heads, total_num_kv_heads=None, bias=False,
                 dtype=None, quant_config=None, prefix="") -> None:
        tp_rank, tp_size = _tp()
        self.head_size = head_size
        self.total_num_heads = total_num_heads
        self.total_num_kv_heads = total_num_kv_heads or total_num_heads
        self.num_heads = divide(total_num_heads, tp_size)
        if tp_size >= self.total_num_kv_heads:
            self.num_kv_heads, self.num_kv_replicas = 1, divide(tp_size, self.total_num_kv_heads)
        else:
            self.num_kv_heads, self.num_kv_replicas = divide(self.total_num_kv_heads, tp_size), 1
        out = (total_num_heads + 2 * self.num_kv_heads * tp_size) * head_size
        super().__init__(hidden_size, out, bias, False, dtype, quant_config, prefix)
        self.q_size = self.num_heads * head_size
        self.kv_size = self.num_kv_heads * head_size

    def weight_loader(self, param: nn.Parameter, w: torch.Tensor, shard_id=None) -> None:
        out_dim = getattr(param, "output_dim", 0)
        unit = getattr(param, "shard_block", 1)
        q, kv = self.q_size // unit, max(self.kv_size // unit, 1)
        if shard_id is None:  # fused checkpoint [q_total | k_total | v_total]
            tq = self.total_num_heads * self.head_size // unit
            tk = self.total_num_kv_heads * self.head_size // unit
            self.weight_loader(param, w.narrow(out_dim, 0, tq), "q")
            self.weight_loader(param, w.narrow(out_dim, tq, tk), "k")
            self.weight_loader(param, w.narrow(out_dim, tq + tk, tk), "v")
            return
        if shard_id == "q":
            off, n, src_rank = 0, q, self.tp_rank
        else:
            off = q + (0 if shard_id == "k" else kv)
            n, src_rank = kv, self.tp_rank // self.num_kv_replicas
        param.data.narrow(out_dim, off, n).copy_(w.narrow(out_dim, src_rank * n, n))


class RowParallelLinear(LinearBase):
    """Y = X W^T with W split along the input dim; partial sums all-reduced over TP."""

    def __init__(self, input_size, output_size, bias=False, input_is_parallel=True, reduce_results=True,
                 dtype=None, quant_config=None, prefix="") -> None:
        super().__init__(input_size, output_size, bias, dtype, quant_config, prefix)
        self.tp_rank, self.tp_size = _tp()
        self.in_per_rank = divide(input_size, self.tp_size)
        self.input_is_parallel = input_is_parallel
        self.reduce_results = reduce_results
        self._make(self.in_per_rank, output_size)

    def _local(self, x: torch.Tensor) -> torch.Tensor:
        if not self.input_is_parallel and self.tp_size > 1:
            x = x.chunk(self.tp_size, dim=-1)[self.tp_rank].contiguous()
        # bias is added once (rank 0) so the all-reduce sums it exactly once
        return self._apply(x, self.bias if self.tp_rank == 0 else None)

    def forward(self, x):
        y = self._local(x)
        if self.reduce_results and self.tp_size > 1:
            y = get_tp_group().all_reduce(y)
        return y

    def forward_reduce_norm(self, x: torch.Tensor, residual: torch.Tensor, norm: Optional[nn.Module]
                            ) -> tuple[torch.Tensor, torch.Tensor]:
        """(normed, new_residual) = norm(all_reduce(x W^T) + residual).

        TP>1 on GPU: one fused kernel (GEMM tiles → NVSwitch reduce → +residual →
        RMSNorm).  Otherwise the unfused sequence with identical numerics."""
        if norm is None:
            return self.forward(x), residual
        if self.tp_size > 1 and x.is_cuda and self.reduce_results:
            comm = get_tp_group().device_comm
            if (comm is not None and self.bias is None and type(self.quant_method) is LinearMethod
                    and not self.lora_table):
                out = comm.gemm_allreduce_rmsnorm(x, self.weight, residual, norm.weight, norm.eps,
                                                  getattr(norm, "gemma", False))
                if out is not None:
                    return out
            return tp_reduce_residual_norm(self._local(x), residual, norm)
        y = self.forward(x)
        return norm(y, residual)

    def weight_loader(self, param: nn.Parameter, w: torch.Tensor, shard_id=None) -> None:
        in_dim = getattr(param, "input_dim", 1)
        if param.dim() == 1:  # bias: replicated
            param.data.copy_(w)
            return
        n = param.shape[in_dim]
        w = w.narrow(in_dim, self.tp_rank * n, n) if w.shape[in_dim] != n else w
        param.data.copy_(w)


def tp_reduce_residual_norm(partial: torch.Tensor, residual: torch.Tensor, norm: nn.Module
                            ) -> tuple[torch.Tensor, torch.Tensor]:
    """(normed, residual) = norm(all_reduce(partial) + residual) for a per-rank partial sum that did not come
    out of a RowParallelLinear (MoE combine, shared-expert sums ...).

    On GPU under TP the residual stream is kept *sharded by row owner* by the device collectives
    (parallel/symm.py), so every reduce+norm of one forward pass must go through them — mixing in a plain
    all-reduce + full-residual norm would read rows this rank does not own."""
    g = get_tp_group()
    if g.world_size > 1 and partial.is_cuda and g.device_comm is not None:
        out = g.device_comm.reduce_residual_norm(partial, residual, norm.weight, norm.eps,
                                                 getattr(norm, "gemma", False))
        if out is not None:
            return out
    if g.world_size > 1:
        partial = g.all_reduce(partial)
    return norm(partial, residual)
