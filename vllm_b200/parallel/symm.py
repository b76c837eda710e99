"""Symmetric-heap communicator: the hot TP collectives as our own sm_100a kernels over NVSwitch.

One symmetric allocation per group (``torch.distributed._symmetric_memory`` does the VMM allocation, the
handle exchange and the NVLS multicast binding — that is *plumbing*; every byte that crosses NVLink on
the hot path is moved by the kernels in ``csrc/comm.cu`` / ``csrc/gemm_ar.cu``):

    [ flags 4 KiB | staging (GEMM partials) | out A | out B ]

* ``gemm_allreduce_rmsnorm``: RowParallel GEMM → cross-GPU sum → + residual → RMSNorm with no NCCL call,
  no intermediate copy: the tcgen05 GEMM epilogue writes its bf16 tile straight into the symmetric staging
  region, the reduction reads all replicas through ``multimem.ld_reduce`` (in-switch add), the residual
  stays sharded by row owner, and the normalised rows are broadcast with ``multimem.st``.
* ``all_reduce`` / ``all_gather`` / ``embedding_allreduce`` for the remaining collectives.

Reference call sites being replaced: SURVEY §2.4 C1 (RowParallelLinear all-reduce + fused_add_rms_norm),
C2 (embedding), C3 (logits gather); reference machinery: custom_all_reduce.py, symm_mem.py,
flashinfer_all_reduce.py, allreduce_rms_fusion.py.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from vllm_b200 import envs
from vllm_b200.logger import init_logger

logger = init_logger(__name__)

FLAGS_BYTES = 4096
LAM_T_CAP = 128      # the fused GEMM+all-reduce kernel handles up to 128 tokens (decode / small mixed batches)
LAM_H_MAX = 8192
ONESHOT_T_MAX = 512  # rows up to which a plain all-reduce is one-shot into private memory (no heap copy-out)
MOE_T_CAP = 256      # tokens covered by the fused MoE combine accumulators (decode / small mixed batches)
_ALIGN = 1 << 16


def _round(n: int) -> int:
    return (n + _ALIGN - 1) // _ALIGN * _ALIGN


class SymmComm:
    def __init__(self, group, region_bytes: int = 64 << 20) -> None:
        import torch.distributed._symmetric_memory as symm_mem
        from vllm_b200.ops import cuda_ops
        self._symm = symm_mem
        self.cu = cuda_ops()
        self.group = group
        self.rank, self.world = group.rank_in_group, group.world_size
        assert self.world <= 8, "one NVSwitch domain: up to 8 ranks"
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.region = 0
        self.buf = None
        self.hdl = None
        self._flip = 0
        self.state = torch.zeros(16, dtype=torch.int32, device=self.device)
        self._alloc(region_bytes)
        self.use_mc = bool(self.mc_ptr) and envs.VLLM_B200_AR_BACKEND in ("auto", "nvls")
        if envs.VLLM_B200_AR_BACKEND == "p2p":
            self.use_mc = False
        # VLLM_BATCH_INVARIANT: every reduction is the rank-ordered one-shot kernel, whatever the row count — the
        # in-switch multimem reduction and the Lamport fused GEMM + all-reduce pick their order / GEMM split by size
        self.batch_invariant = bool(envs.VLLM_BATCH_INVARIANT)
        if self.batch_invariant:
            self.use_mc = False
        m = envs.VLLM_B200_FUSED_AR_MODE
        self.fused_mode = 0 if m == "oneshot" else 1
        logger.info("SymmComm[%s] rank %d/%d: %d MiB regions, multicast=%s", group.name, self.rank, self.world,
                    self.region >> 20, self.use_mc)

    # ---------------------------------------------------------------- memory
    def _alloc(self, region_bytes: int) -> None:
        region = _round(region_bytes)
        # Lamport receive buffers for the fused decode kernel: 3 rotating buffers x world x T_CAP rows x H_MAX
        self.lam_stride = _round(self.world * LAM_T_CAP * LAM_H_MAX * 2)
        # two-shot mode: 3 rotating output buffers (normalised rows multicast by their owners), sentinel-armed too
        self.oring_stride = _round(LAM_T_CAP * LAM_H_MAX * 2)
        # fused MoE combine: 3 rotating fp32 accumulators [MOE_T_CAP, LAM_H_MAX] that the w2 grouped-GEMM epilogues of ALL
        # ranks add into through the multicast alias (multimem.red); zeroed again by the finish kernel that reads them
        self.moe_stride = _round(MOE_T_CAP * LAM_H_MAX * 4)
        total = FLAGS_BYTES + 3 * region + 3 * self.lam_stride + 3 * self.oring_stride + 3 * self.moe_stride
        self.buf = self._symm.empty(total, dtype=torch.uint8, device=self.device)
        self.buf.zero_()
        self.off_lam = FLAGS_BYTES + 3 * region
        self.off_oring = self.off_lam + 3 * self.lam_stride
        self.off_moe = self.off_oring + 3 * self.oring_stride
        self._moe_flip = 0
        self.buf[self.off_lam:self.off_moe].view(torch.int16).fill_(-32768)   # sentinel = bf16/fp16 negative zero
        self.lam_H = 0
        self._armed_mode = -1
        torch.cuda.synchronize()
        self.hdl = self._symm.rendezvous(self.buf, self.group.device_group)
        self.region = region
        self.off_in = FLAGS_BYTES
        self.off_out = [FLAGS_BYTES + region, FLAGS_BYTES + 2 * region]
        ptrs = list(self.hdl.buffer_ptrs)
        self.peer_arr = (ctypes.c_void_p * 8)(*[ptrs[i] if i < self.world else None for i in range(8)])
        self.local_ptr = ptrs[self.rank]
        mc = 0
        try:
            mc = int(self.hdl.multicast_ptr or 0)
        except Exception:
            mc = 0
        self.mc_ptr = mc
        self.state.zero_()
        self.group.barrier()

    def ensure_capacity(self, nbytes: int) -> None:
        if nbytes <= self.region:
            return
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("symmetric heap too small for a collective issued during CUDA-graph capture")
        torch.cuda.synchronize()
        self.group.barrier()
        self._alloc(max(nbytes, 2 * self.region))

    def _view(self, off: int, shape, dtype) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= s
        nbytes = n * dtype.itemsize
        return self.buf[off:off + nbytes].view(dtype).view(*shape)

    def staging(self, shape, dtype) -> torch.Tensor:
        """A tensor inside the symmetric staging region (GEMM epilogues write their partial sums here)."""
        return self._view(self.off_in, shape, dtype)

    def _next_out(self) -> int:
        self._flip ^= 1
        return self.off_out[self._flip]

    def _mc(self) -> Optional[int]:
        return self.mc_ptr if self.use_mc else None

    # ------------------------------------------------------------ collectives
    def _ok(self, x: torch.Tensor) -> bool:
        return (x.is_cuda and x.dtype in (torch.bfloat16, torch.float16) and x.dim() >= 1
                and x.shape[-1] % 8 == 0 and x.shape[-1] // 8 <= 1024)

    def allreduce_rows(self, x: Optional[torch.Tensor], T: int, H: int, dtype, residual=None, norm_w=None,
                       eps: float = 0.0, gemma: bool = False) -> torch.Tensor:
        """Sum over ranks of [T,H] partials (``x`` or, if None, what is already in the staging region),
        optional += residual (row-sharded, in place) and RMSNorm. Returns a view of the output region."""
        nbytes = T * H * dtype.itemsize
        self.ensure_capacity(nbytes)
        out_off = self._next_out()
        self.cu.comm_allreduce_rows(self, x, self.off_in, out_off, residual, norm_w, eps, T, H, gemma, dtype)
        return self._view(out_off, (T, H), dtype)

    def all_reduce(self, x: torch.Tensor) -> Optional[torch.Tensor]:
        if not self._ok(x) or not x.is_contiguous():
            return None
        H = x.shape[-1]
        T = x.numel() // H
        if T <= ONESHOT_T_MAX or self.batch_invariant:       # decode-size: every rank reduces all rows in the switch, straight into private memory
            self.ensure_capacity(T * H * x.dtype.itemsize)
            out = torch.empty_like(x)
            self.cu.comm_allreduce_oneshot(self, x.view(T, H), self.off_in, out, T, H)
            return out
        # large T: two-shot (1/W of the traffic per rank); the result lands in a rotating heap region that later
        # collectives reuse, so it is copied out once
        out = self.allreduce_rows(x.view(T, H), T, H, x.dtype)
        return out.view(x.shape).clone()

    def all_gather(self, x: torch.Tensor, dim: int) -> Optional[torch.Tensor]:
        if dim != x.dim() - 1 or not self._ok(x) or not x.is_contiguous():
            return None
        cols = x.shape[-1]
        rows = x.numel() // cols
        nbytes = rows * cols * self.world * x.dtype.itemsize
        self.ensure_capacity(nbytes)
        out_off = self._next_out()
        self.cu.comm_all_gather(self, x.view(rows, cols), out_off, rows, cols)
        return self._view(out_off, x.shape[:-1] + (cols * self.world,), x.dtype)

    def gemm_allreduce_rmsnorm(self, x: torch.Tensor, w: torch.Tensor, residual: torch.Tensor,
                               norm_w: torch.Tensor, eps: float, gemma: bool = False):
        """(normed, residual) = RMSNorm(all_reduce(x @ w.T) + residual)."""
        if not envs.VLLM_B200_FUSED_AR or x.dtype not in (torch.bfloat16, torch.float16) or self.batch_invariant:
            return None
        K = x.shape[-1]
        x2 = x.reshape(-1, K)
        T, H = x2.shape[0], w.shape[0]
        if H % 8 or H // 8 > 1024 or not residual.is_contiguous():
            return None
        if (T <= LAM_T_CAP and H <= LAM_H_MAX and envs.VLLM_B200_FUSED_AR and self.cu._tc_ok(x2, w, force=True)
                and residual.dtype == x.dtype):
            # decode: ONE kernel — tcgen05 GEMM whose epilogue multicasts the tile to every GPU (Lamport buffers),
            # then reduces + residual + RMSNorm in place; no barrier, no NCCL
            # default ("two-shot"): a one-shot push makes every rank receive and reduce W x the data; instead each
            # partial row goes only to the rank that owns the row, which reduces + normalises it and multicasts the
            # result — still ONE kernel and no barrier, inbound traffic independent of W.  The residual then stays
            # sharded by row owner (as in the prefill path).
            if H != self.lam_H or self._armed_mode != self.fused_mode:
                self._rearm_lamport(H)
            out = torch.empty((T, H), dtype=x.dtype, device=x.device)
            if self.cu.gemm_ar(self, x2, w, residual.view(T, H), norm_w, out, eps, gemma, self.fused_mode,
                               self.off_oring, self.oring_stride):
                return out.view(x.shape[:-1] + (H,)), residual
        self.ensure_capacity(T * H * x.dtype.itemsize)
        part = self.staging((T, H), x.dtype)
        if self.cu._tc_ok(x2, w):
            self.cu.gemm_tc(x2, w, None, out=part)       # epilogue stores straight into the symmetric heap
        else:
            torch.matmul(x2, w.t(), out=part)
        out = self.allreduce_rows(None, T, H, x.dtype, residual.view(T, H), norm_w, eps, gemma)
        return out.view(x.shape[:-1] + (H,)), residual

    def reduce_residual_norm(self, partial: torch.Tensor, residual: torch.Tensor, norm_w: torch.Tensor, eps: float,
                             gemma: bool = False):
        """Same contract as ``gemm_allreduce_rmsnorm`` for a partial sum that is already computed."""
        H = partial.shape[-1]
        if (self.batch_invariant or partial.dtype not in (torch.bfloat16, torch.float16) or H % 8 or H // 8 > 1024
                or not residual.is_contiguous() or residual.dtype != partial.dtype):
            return None
        p2 = partial.reshape(-1, H)
        if not p2.is_contiguous():
            p2 = p2.contiguous()
        T = p2.shape[0]
        out = self.allreduce_rows(p2, T, H, partial.dtype, residual.view(T, H), norm_w, eps, gemma)
        return out.view(partial.shape), residual

    def _rearm_lamport(self, H: int) -> None:
        """The Lamport slot layout depends on H: switching H re-arms all buffers (host side, collective)."""
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("Lamport buffers must be armed for this hidden size before CUDA-graph capture")
        torch.cuda.synchronize()
        self.group.barrier()
        self.buf[self.off_lam:self.off_moe].view(torch.int16).fill_(-32768)
        self.state[8:13].zero_()
        torch.cuda.synchronize()
        self.group.barrier()
        self.lam_H = H
        self._armed_mode = self.fused_mode

    def embedding_allreduce(self, ids: torch.Tensor, weight: torch.Tensor, start: int) -> Optional[torch.Tensor]:
        """Vocab-sharded embedding lookup + all-reduce: the masked lookup lands in the staging region,
        the reduce kernel sums the replicas."""
        H = weight.shape[1]
        if weight.dtype not in (torch.bfloat16, torch.float16) or H % 8 or H // 8 > 1024:
            return None
        T = ids.numel()
        self.ensure_capacity(T * H * weight.dtype.itemsize)
        part = self.staging((T, H), weight.dtype)
        per = weight.shape[0]
        mask = (ids < start) | (ids >= start + per)
        local = (ids - start).masked_fill(mask, 0)
        torch.index_select(weight, 0, local.reshape(-1).to(torch.int64), out=part)
        part.masked_fill_(mask.reshape(-1, 1), 0)
        if T <= ONESHOT_T_MAX or self.batch_invariant:
            out = torch.empty(ids.shape + (H,), dtype=weight.dtype, device=weight.device)
            self.cu.comm_allreduce_oneshot(self, None, self.off_in, out, T, H)
            return out
        out = self.allreduce_rows(None, T, H, weight.dtype)
        return out.view(ids.shape + (H,)).clone()

    def next_moe_acc(self) -> int:
        """Byte offset (same on every rank) of the next rotating fp32 MoE accumulator; its owned rows are all-zero."""
        self._moe_flip = (self._moe_flip + 1) % 3
        return self.off_moe + self._moe_flip * self.moe_stride

    def moe_finish_rows(self, acc_off: int, T: int, H: int, dtype, residual, norm_w, eps: float, gemma: bool = False):
        """Rows summed into the accumulator by the grouped-GEMM epilogues → + residual → RMSNorm → every rank."""
        self.ensure_capacity(T * H * dtype.itemsize)
        out_off = self._next_out()
        self.cu.comm_finish_rows_acc32(self, acc_off, out_off, residual, norm_w, eps, T, H, gemma, dtype)
        return self._view(out_off, (T, H), dtype)

    def barrier(self) -> None:
        self.cu.comm_barrier(self)

    def close(self) -> None:
        self.buf = None
        self.hdl = None
