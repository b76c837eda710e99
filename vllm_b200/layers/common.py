"""Norm / activation / embedding / logits layers.

Behavioural reference: layers/layernorm.py:37 (RMSNorm with fused residual),
layers/activation.py, layers/vocab_parallel_embedding.py:198,487-507,521,
layers/logits_processor.py:57,118-187.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from vllm_b200 import ops
from vllm_b200.parallel.state import get_tp_group
from vllm_b200.utils import round_up


class RMSNorm(nn.Module):
    def __init__(self, hidden_size: int, eps: float = 1e-6, dtype: Optional[torch.dtype] = None,
                 gemma: bool = False) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size, dtype=dtype or torch.get_default_dtype()),
                                   requires_grad=False)
        if gemma:
            self.weight.data.zero_()
        self.eps = eps
        self.gemma = gemma

    def forward(self, x: torch.Tensor, residual: Optional[torch.Tensor] = None):
        if residual is None:
            return ops.rms_norm(x, self.weight, self.eps, self.gemma)
        return ops.fused_add_rms_norm(x, residual, self.weight, self.eps, self.gemma)


class LayerNorm(nn.Module):
    def __init__(self, hidden_size: int, eps: float = 1e-5, bias: bool = True, dtype=None) -> None:
        super().__init__()
        dt = dtype or torch.get_default_dtype()
        self.weight = nn.Parameter(torch.ones(hidden_size, dtype=dt), requires_grad=False)
        self.bias = nn.Parameter(torch.zeros(hidden_size, dtype=dt), requires_grad=False) if bias else None
        self.eps = eps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.layer_norm(x, self.weight, self.bias, self.eps)


class SiluAndMul(nn.Module):
    def forward(self, x):
        return ops.silu_and_mul(x)


class GeluAndMul(nn.Module):
    def __init__(self, approximate: str = "none") -> None:
        super().__init__()
        self.approximate = approximate

    def forward(self, x):
        return ops.gelu_and_mul(x, self.approximate)


class Activation(nn.Module):
    def __init__(self, name: str) -> None:
        super().__init__()
        self.name = name

    def forward(self, x):
        return ops.activation(x, self.name)


def get_act_and_mul(name: str) -> nn.Module:
    if name == "silu":
        return SiluAndMul()
    if name in ("gelu", "gelu_pytorch_tanh"):
        return GeluAndMul("tanh" if name == "gelu_pytorch_tanh" else "none")
    raise ValueError(f"unsupported gated activation {name}")


class VocabParallelEmbedding(nn.Module):
    """Embedding table sharded over the vocab dim; lookup of out-of-shard ids gives 0
    and the TP all-reduce restores the full embedding (collective C2 in SURVEY §2.4)."""

    def __init__(self, num_embeddings: int, embedding_dim: int, dtype=None, padding_size: int = 64,
                 prefix: str = "", replicate: bool = False) -> None:
        super().__init__()
        g = get_tp_group()
        self.tp_rank, self.tp_size = max(g.rank_in_group, 0), g.world_size
        if replicate:
            # B200-first: with 180 GB per GPU the input embedding (<= a few GB) is simply replicated, which
            # removes the embedding all-reduce (collective C2) from every step; the LM head stays sharded
            self.tp_rank, self.tp_size = 0, 1
        self.num_embeddings = num_embeddings
        self.padded = round_up(num_embeddings, padding_size * self.tp_size)
        self.per_rank = self.padded // self.tp_size
        self.start = self.tp_rank * self.per_rank
        self.end = self.start + self.per_rank
        self.embedding_dim = embedding_dim
        self.weight = nn.Parameter(
            torch.zeros(self.per_rank, embedding_dim, dtype=dtype or torch.get_default_dtype()),
            requires_grad=False)

    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        if self.tp_size == 1:
            return torch.nn.functional.embedding(ids, self.weight)
        comm = get_tp_group().device_comm
        if comm is not None and ids.is_cuda:
            out = comm.embedding_allreduce(ids, self.weight, self.start)
            if out is not None:
                return out
        mask = (ids < self.start) | (ids >= self.end)
        local = (ids - self.start).masked_fill(mask, 0)
        out = torch.nn.functional.embedding(local, self.weight)
        out.masked_fill_(mask.unsqueeze(-1), 0)
        return get_tp_group().all_reduce(out)

    def weight_loader(self, param: nn.Parameter, w: torch.Tensor, shard_id=None) -> None:
        n = min(self.end, w.shape[0]) - self.start
        if n > 0:
            param.data[:n].copy_(w[self.start:self.start + n])
        if n < self.per_rank:
            param.data[max(n, 0):].zero_()


class ParallelLMHead(VocabParallelEmbedding):
    def __init__(self, num_embeddings, embedding_dim, bias=False, dtype=None, prefix="") -> None:
        super().__init__(num_embeddings, embedding_dim, dtype, prefix=prefix)
        self.bias = nn.Parameter(torch.zeros(self.per_rank, dtype=self.weight.dtype),
                                 requires_grad=False) if bias else None

    def tie_weights(self, embed: VocabParallelEmbedding) -> None:
        self.weight = embed.weight

    def forward(self, ids):  # pragma: no cover
        raise RuntimeError("ParallelLMHead is used through LogitsProcessor")


class LogitsProcessor(nn.Module):
    """hidden → vocab logits. TP: each rank computes its vocab shard, shards are
    all-gathered (collective C3)."""

    def __init__(self, vocab_size: int, scale: float = 1.0, soft_cap: Optional[float] = None) -> None:
        super().__init__()
        self.vocab_size = vocab_size
        self.scale = scale
        self.soft_cap = soft_cap

    def sample_candidates(self, lm_head: ParallelLMHead, hidden: torch.Tensor, temperature, seeds, steps,
                          out: torch.Tensor) -> bool:
        """TP fast path for greedy / temperature-only sampling (no logprobs, penalties, filters): every rank runs the
        fused sampler over ITS vocab shard and only one 16-byte (value, token id) candidate per row crosses NVLink
        (our all-gather kernel), instead of the full logits (reference call site: logits_processor.py:118-130,171-187
        gathers the whole vocabulary).  Returns False when not applicable (the caller takes the full-logits path)."""
        g = get_tp_group()
        if (lm_head.tp_size == 1 or g.device_comm is None or not hidden.is_cuda or self.soft_cap or self.scale != 1.0
                or lm_head.bias is not None):
            return False
        logits = ops.linear(hidden, lm_head.weight, None)                     # [B, vocab / W] local shard
        cu = ops.cuda_ops()
        v_valid = min(lm_head.per_rank, self.vocab_size - lm_head.start)
        cand = cu.fused_sample_candidates(logits, temperature, seeds, steps, lm_head.start, v_valid)
        gathered = g.device_comm.all_gather(cand.view(torch.bfloat16), 1)     # [B, 8 * W] bf16 = [B, W, 4] fp32
        if gathered is None:
            return False
        cu.select_candidates(gathered.view(torch.float32), lm_head.tp_size, out)
        return True

    def forward(self, lm_head: ParallelLMHead, hidden: torch.Tensor) -> torch.Tensor:
        logits = ops.linear(hidden, lm_head.weight, lm_head.bias)
        if lm_head.tp_size > 1:
            logits = get_tp_group().all_gather(logits, -1)
        logits = logits[..., :self.vocab_size]
        if self.soft_cap:
            logits = torch.tanh(logits / self.soft_cap) * self.soft_cap
        if self.scale != 1.0:
            logits = logits * self.scale
        return logits
