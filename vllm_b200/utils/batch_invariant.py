"""Batch-invariant execution (``VLLM_BATCH_INVARIANT=1``): a request's logits do not depend on what else is in the batch.

What changes (policy only — the kernels are the same, chosen so a row's floating-point reduction order is a function of
that row alone):

=====================  =========================================================  ==========================================
op                     default                                                    batch-invariant
=====================  =========================================================  ==========================================
decode GEMM (M ≤ 256)  cluster split-K chosen per (N, K); none above 256 rows      no split-K at any M (``ops/cuda.py _plan``)
fp8 GEMM               split-K per (N, K)                                          no split-K
decode attention       split-KV count from batch × kv heads × context              one CTA per (sequence, kv head), pages in order
decode-only batches    swap-AB decode kernel                                       the varlen kernel a mixed batch would use
MLA decode             split-KV from batch × heads                                 no split
TP all-reduce          one-shot ≤ 512 rows, NVLS two-shot above, Lamport fused     rank-ordered one-shot at every size
                       GEMM + all-reduce for decode
MoE combine            ``red.add.f32`` epilogue into per-token accumulators        sorted combine kernel (fixed k order)
NCCL fallbacks         library's size-dependent algorithm / protocol               tree + Simple, one channel, no NVLS / CollNet
=====================  =========================================================  ==========================================

Per-row kernels (RMSNorm / LayerNorm: one CTA per row with a thread count that depends on H only; RoPE; activation;
sampling: one row per CTA) are already invariant.  Not covered: MLA decode rows inside a mixed batch (they take the
prefill path), FP4 weight-only GEMMs, LoRA shrink/expand split over slots.

Reference: vllm/model_executor/layers/batch_invariant.py (``init_batch_invariance`` /
``override_envs_for_invariance``), vllm/distributed/device_communicators/cuda_communicator.py:61-64 (custom all-reduce
and symmetric-memory paths disabled in this mode), tests/v1/determinism/.
"""
from __future__ import annotations

import os

from vllm_b200 import envs

_NCCL = {"NCCL_LAUNCH_MODE": "GROUP", "NCCL_COLLNET_ENABLE": "0", "NCCL_NVLS_ENABLE": "0", "NCCL_P2P_NET_DISABLE": "1",
         "NCCL_MIN_NCHANNELS": "1", "NCCL_MAX_NCHANNELS": "1", "NCCL_PROTO": "Simple", "NCCL_ALGO": "allreduce:tree",
         "NCCL_NTHREADS": "1024", "NCCL_SOCKET_NTHREADS": "1"}


def enabled() -> bool:
    return bool(envs.VLLM_BATCH_INVARIANT)


def apply_env_overrides() -> bool:
    """Called by every worker before the process group exists: pins the library collectives to one algorithm and turns
    off TF32 / reduced-precision accumulation in the few cuBLAS calls left on the path."""
    if not enabled():
        return False
    for k, v in _NCCL.items():
        os.environ[k] = v
    os.environ.setdefault("CUBLAS_WORKSPACE_CONFIG", ":4096:8")
    import torch
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_bf16_reduced_precision_reduction = False
    torch.backends.cuda.matmul.allow_fp16_reduced_precision_reduction = False
    return True
