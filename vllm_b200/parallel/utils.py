"""Partitioning helpers (reference: vllm/distributed/utils.py ``get_pp_indices``, ``divide``)."""
from __future__ import annotations

import os


def divide(a: int, b: int) -> int:
    assert a % b == 0, f"{a} is not divisible by {b}"
    return a // b


def get_pp_indices(num_layers: int, pp_rank: int, pp_size: int) -> tuple[int, int]:
    """Contiguous layer range of a pipeline stage; remainder layers go to the middle
    stages (first/last also hold embedding / LM head). ``VLLM_PP_LAYER_PARTITION=a,b,c`` overrides."""
    part = os.environ.get("VLLM_PP_LAYER_PARTITION")
    if part:
        sizes = [int(x) for x in part.split(",")]
        if len(sizes) != pp_size or sum(sizes) != num_layers:
            raise ValueError(f"invalid VLLM_PP_LAYER_PARTITION={part} for {num_layers} layers / pp={pp_size}")
    else:
        base, rem = divmod(num_layers, pp_size)
        sizes = [base] * pp_size
        # hand out the remainder starting from the second stage, ending before the last
        order = list(range(1, pp_size - 1)) + [0, pp_size - 1] if pp_size > 2 else list(range(pp_size))
        for i in range(rem):
            sizes[order[i % len(order)]] += 1
    start = sum(sizes[:pp_rank])
    return start, start + sizes[pp_rank]
