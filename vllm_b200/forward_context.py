"""Per-step ambient state for model code (reference: vllm/forward_context.py,
v1/attention/backend.py:369-452 ``CommonAttentionMetadata``).

The model runner builds one ``AttentionMetadata`` per KV-cache group per step
and installs it with ``set_forward_context``; ``Attention`` layers read it.
Under CUDA-graph replay the tensors are the runner's static buffers.
"""
from __future__ import annotations

import contextlib
from dataclasses import dataclass, field
from typing import Any, Optional

import torch


@dataclass
class AttentionMetadata:
    num_reqs: int
    num_actual_tokens: int
    max_query_len: int
    max_seq_len: int
    query_start_loc: torch.Tensor  # [B+1] int32, device
    seq_lens: torch.Tensor         # [B] int32, device
    block_table: torch.Tensor      # [B, max_blocks] int32, device
    slot_mapping: torch.Tensor     # [T] int64, device
    query_start_loc_cpu: Optional[torch.Tensor] = None
    seq_lens_cpu: Optional[torch.Tensor] = None
    decode_only: bool = False      # every request has exactly one new token
    kv_vt: bool = False            # tensor-core cache layout: V pages stored transposed [D, block_size]
    causal: bool = True
    # mixed-batch CUDA graphs: the varlen attention grid is sized from (tokens, requests) of the BUCKET and every CTA finds
    # its (request, query tile) from query_start_loc on the device (csrc/attention_tc.cu flat schedule)
    flat: bool = False
    # decode split-KV workspace (allocated once by the runner)
    workspace: Any = None
    num_splits: int = 0
    # cascade attention: (metadata of the shared-prefix pass, metadata of the per-request suffix pass) when every
    # request of the step shares its leading KV blocks (reference: v1/attention/backends use_cascade_attention)
    cascade: Any = None


@dataclass
class ForwardContext:
    attn_metadata: dict[int, AttentionMetadata] = field(default_factory=dict)  # KV group id → metadata
    kv_caches: dict[str, torch.Tensor] = field(default_factory=dict)           # layer name → cache
    layer_to_group: dict[str, int] = field(default_factory=dict)
    num_tokens: int = 0
    skip_attention: bool = False   # profiling run without a KV cache
    lora: Any = None
    extra: dict[str, Any] = field(default_factory=dict)

    def metadata_for(self, layer_name: str) -> Optional[AttentionMetadata]:
        return self.attn_metadata.get(self.layer_to_group.get(layer_name, 0))


_ctx: Optional[ForwardContext] = None


def get_forward_context() -> ForwardContext:
    assert _ctx is not None, "forward context is not set"
    return _ctx


def try_get_forward_context() -> Optional[ForwardContext]:
    return _ctx


@contextlib.contextmanager
def set_forward_context(ctx: ForwardContext):
    global _ctx
    prev, _ctx = _ctx, ctx
    try:
        yield ctx
    finally:
        _ctx = prev


def get_forward_context_or_none() -> Optional[ForwardContext]:
    return _ctx
