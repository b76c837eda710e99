"""Process-group state: one process per GPU, groups for TP / PP / DP / EP / DCP / PCP.

Behavioural reference: vllm/distributed/parallel_state.py (``GroupCoordinator``
:380, ``initialize_model_parallel`` :1751, rank layout DP x PP x PCP x TP
:1817-1832, accessors :1389-1445).

B200-first design: ``torch.distributed`` (NCCL on GPU, gloo on CPU) is only the
*plumbing* (bootstrap, metadata, fallbacks, PP send/recv).  Hot collectives go
through ``vllm_b200.parallel.symm`` — a symmetric heap mapped into every peer
with an NVLS multicast alias — and are issued from inside our own sm_100a
kernels (see ``ops/comm.py``).  ``GroupCoordinator.all_reduce`` picks:
fused/NVLS kernel → P2P kernel → NCCL, by size and availability.
"""
from __future__ import annotations

import contextlib
import os
from dataclasses import dataclass
from datetime import timedelta
from typing import Any, Optional

import torch
import torch.distributed as dist

from vllm_b200.logger import init_logger

logger = init_logger(__name__)


class GroupCoordinator:
    def __init__(self, ranks: list[int], rank: int, backend: str, name: str,
                 use_device_comm: bool = False) -> None:
        self.name = name
        self.ranks = ranks
        self.rank = rank
        self.world_size = len(ranks)
        self.rank_in_group = ranks.index(rank) if rank in ranks else -1
        self.device_group = None
        self.cpu_group = None
        self.device_comm = None  # vllm_b200.parallel.symm.SymmComm
        self.backend = backend
        self.use_device_comm = use_device_comm

    # groups must be created collectively by all ranks: done in _make_groups
    @property
    def is_first_rank(self) -> bool:
        return self.rank_in_group == 0

    @property
    def is_last_rank(self) -> bool:
        return self.rank_in_group == self.world_size - 1

    @property
    def first_rank(self) -> int:
        return self.ranks[0]

    @property
    def last_rank(self) -> int:
        return self.ranks[-1]

    @property
    def next_rank(self) -> int:
        return self.ranks[(self.rank_in_group + 1) % self.world_size]

    @property
    def prev_rank(self) -> int:
        return self.ranks[(self.rank_in_group - 1) % self.world_size]

    # ---------------------------------------------------------- collectives
    def all_reduce(self, x: torch.Tensor) -> torch.Tensor:
        if self.world_size == 1:
            return x
        if self.device_comm is not None and x.is_cuda:
            out = self.device_comm.all_reduce(x)
            if out is not None:
                return out
        dist.all_reduce(x, group=self.device_group)
        return x

    def all_gather(self, x: torch.Tensor, dim: int = -1) -> torch.Tensor:
        if self.world_size == 1:
            return x
        if dim < 0:
            dim += x.dim()
        if self.device_comm is not None and x.is_cuda:
            out = self.device_comm.all_gather(x, dim)
            if out is not None:
                return out
        shape = list(x.shape)
        out = torch.empty([self.world_size * shape[0]] + shape[1:], dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x.contiguous(), group=self.device_group)
        out = out.view([self.world_size] + shape).movedim(0, dim)
        shape[dim] *= self.world_size
        return out.reshape(shape)

    def reduce_scatter(self, x: torch.Tensor, dim: int = 0) -> torch.Tensor:
        if self.world_size == 1:
            return x
        if dim < 0:
            dim += x.dim()
        xm = x.movedim(dim, 0).contiguous()
        chunk = xm.shape[0] // self.world_size
        out = torch.empty((chunk,) + tuple(xm.shape[1:]), dtype=x.dtype, device=x.device)
        dist.reduce_scatter_tensor(out, xm, group=self.device_group)
        return out.movedim(0, dim)

    def gather(self, x: torch.Tensor, dst: int = 0, dim: int = -1) -> Optional[torch.Tensor]:
        if self.world_size == 1:
            return x
        # NCCL gather == all_gather + select; sizes are small (logits)
        out = self.all_gather(x, dim)
        return out if self.rank_in_group == dst else None

    def broadcast(self, x: torch.Tensor, src: int = 0) -> torch.Tensor:
        if self.world_size == 1:
            return x
        dist.broadcast(x, src=self.ranks[src], group=self.device_group)
        return x

    def broadcast_object(self, obj: Any = None, src: int = 0) -> Any:
        if self.world_size == 1:
            return obj
        lst = [obj]
        dist.broadcast_object_list(lst, src=self.ranks[src], group=self.cpu_group)
        return lst[0]

    def all_gather_object(self, obj: Any) -> list[Any]:
        if self.world_size == 1:
            return [obj]
        out = [None] * self.world_size
        dist.all_gather_object(out, obj, group=self.cpu_group)
        return out

    def barrier(self) -> None:
        if self.world_size > 1:
            dist.barrier(group=self.cpu_group)

    def send(self, x: torch.Tensor, dst: Optional[int] = None) -> None:
        dst = self.next_rank if dst is None else self.ranks[dst]
        dist.send(x, dst, group=self.device_group)

    def recv(self, shape, dtype, src: Optional[int] = None, device=None) -> torch.Tensor:
        src = self.prev_rank if src is None else self.ranks[src]
        x = torch.empty(shape, dtype=dtype, device=device)
        dist.recv(x, src, group=self.device_group)
        return x

    def send_tensor_dict(self, d: dict[str, torch.Tensor], dst: Optional[int] = None) -> None:
        """PP boundary: metadata over the CPU group, tensors over the device group."""
        dst_g = self.next_rank if dst is None else self.ranks[dst]
        meta = [(k, tuple(v.shape), v.dtype) for k, v in d.items()]
        dist.send_object_list([meta], dst=dst_g, group=self.cpu_group)
        for v in d.values():
            dist.send(v.contiguous(), dst_g, group=self.device_group)

    def recv_tensor_dict(self, src: Optional[int] = None, device=None) -> dict[str, torch.Tensor]:
        src_g = self.prev_rank if src is None else self.ranks[src]
        box = [None]
        dist.recv_object_list(box, src=src_g, group=self.cpu_group)
        out = {}
        for k, shape, dtype in box[0]:
            t = torch.empty(shape, dtype=dtype, device=device)
            dist.recv(t, src_g, group=self.device_group)
            out[k] = t
        return out

    def destroy(self) -> None:
        if self.device_comm is not None:
            self.device_comm.close()
            self.device_comm = None


@dataclass
class _State:
    world: Optional[GroupCoordinator] = None
    tp: Optional[GroupCoordinator] = None
    pp: Optional[GroupCoordinator] = None
    dp: Optional[GroupCoordinator] = None
    ep: Optional[GroupCoordinator] = None
    dcp: Optional[GroupCoordinator] = None
    pcp: Optional[GroupCoordinator] = None
    initialized: bool = False
    we_inited_pg: bool = False


_S = _State()
_SINGLE = GroupCoordinator([0], 0, "none", "single")


def init_distributed_environment(world_size: int, rank: int, local_rank: int = 0,
                                 init_method: Optional[str] = None, backend: Optional[str] = None,
                                 timeout_s: int = 600) -> None:
    if world_size == 1 and not dist.is_initialized():
        return
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            backend = "cpu:gloo,cuda:nccl"
        if init_method is None:
            addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
            port = os.environ.get("MASTER_PORT", "29500")
            init_method = f"tcp://{addr}:{port}"
        kw = {}
        if torch.cuda.is_available():
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, init_method=init_method, world_size=world_size,
                                rank=rank, timeout=timedelta(seconds=timeout_s), **kw)
        _S.we_inited_pg = True


def _make_groups(all_rank_lists: list[list[int]], name: str, backend: str, use_device_comm: bool
                 ) -> GroupCoordinator:
    """Collectively create every subgroup; return the one this rank belongs to."""
    rank = dist.get_rank()
    mine = None
    for ranks in all_rank_lists:
        dev = dist.new_group(ranks, backend=backend)
        cpu = dist.new_group(ranks, backend="gloo")
        if rank in ranks:
            mine = GroupCoordinator(ranks, rank, backend, name, use_device_comm)
            mine.device_group, mine.cpu_group = dev, cpu
    assert mine is not None
    return mine


def initialize_model_parallel(tp: int = 1, pp: int = 1, dp: int = 1, dcp: int = 1, pcp: int = 1,
                              enable_device_comm: bool = True) -> None:
    """Build the groups. Layout: rank = ((dp_i * pp + pp_i) * pcp + pcp_i) * tp + tp_i."""
    if _S.initialized:
        return
    if not dist.is_initialized():
        assert tp * pp * dp * pcp == 1, "torch.distributed must be initialised for world_size > 1"
        _S.world = _S.tp = _S.pp = _S.dp = _S.ep = _S.dcp = _S.pcp = _SINGLE
        _S.initialized = True
        return
    world = dist.get_world_size()
    rank = dist.get_rank()
    assert world == tp * pp * dp * pcp, f"world_size {world} != tp{tp}*pp{pp}*dp{dp}*pcp{pcp}"
    backend = "nccl" if torch.cuda.is_available() and dist.get_backend() != "gloo" else "gloo"
    idx = torch.arange(world).reshape(dp, pp, pcp, tp)
    _S.world = GroupCoordinator(list(range(world)), rank, backend, "world")
    _S.world.device_group = dist.group.WORLD
    _S.world.cpu_group = dist.new_group(list(range(world)), backend="gloo")
    _S.tp = _make_groups(idx.reshape(-1, tp).tolist(), "tp", backend, True)
    _S.pp = _make_groups(idx.permute(0, 2, 3, 1).reshape(-1, pp).tolist(), "pp", backend, False)
    _S.dp = _make_groups(idx.permute(1, 2, 3, 0).reshape(-1, dp).tolist(), "dp", backend, False)
    _S.pcp = _make_groups(idx.permute(0, 1, 3, 2).reshape(-1, pcp).tolist(), "pcp", backend, False)
    # EP spans DP x PCP x TP for a fixed PP stage
    _S.ep = _make_groups(idx.permute(1, 0, 2, 3).reshape(pp, -1).tolist(), "ep", backend, True)
    # DCP re-uses TP ranks in chunks of dcp
    _S.dcp = _make_groups(idx.reshape(-1, dcp).tolist(), "dcp", backend, False) if dcp > 1 else _SINGLE
    if enable_device_comm and backend == "nccl":
        from vllm_b200.parallel.symm import SymmComm
        for g in (_S.tp, _S.ep):
            if g.world_size > 1 and g.device_comm is None:
                if g is _S.ep and _S.ep.ranks == _S.tp.ranks:
                    g.device_comm = _S.tp.device_comm
                    continue
                try:
                    g.device_comm = SymmComm(g)
                except Exception as e:  # pragma: no cover - hardware dependent
                    logger.warning("symmetric-memory communicator unavailable for %s (%s); using NCCL",
                                   g.name, e)
    _S.initialized = True


def model_parallel_is_initialized() -> bool:
    return _S.initialized


def ensure_model_parallel_initialized() -> None:
    if not _S.initialized:
        initialize_model_parallel()


def get_world_group() -> GroupCoordinator:
    return _S.world or _SINGLE


def get_tp_group() -> GroupCoordinator:
    return _S.tp or _SINGLE


def get_pp_group() -> GroupCoordinator:
    return _S.pp or _SINGLE


def get_dp_group() -> GroupCoordinator:
    return _S.dp or _SINGLE


def get_ep_group() -> GroupCoordinator:
    return _S.ep or _SINGLE


def get_dcp_group() -> GroupCoordinator:
    return _S.dcp or _SINGLE


def get_pcp_group() -> GroupCoordinator:
    return _S.pcp or _SINGLE


def get_tensor_model_parallel_world_size() -> int:
    return get_tp_group().world_size


def get_tensor_model_parallel_rank() -> int:
    return max(get_tp_group().rank_in_group, 0)


def tensor_model_parallel_all_reduce(x: torch.Tensor) -> torch.Tensor:
    return get_tp_group().all_reduce(x)


def tensor_model_parallel_all_gather(x: torch.Tensor, dim: int = -1) -> torch.Tensor:
    return get_tp_group().all_gather(x, dim)


def tensor_model_parallel_gather(x: torch.Tensor, dst: int = 0, dim: int = -1) -> Optional[torch.Tensor]:
    return get_tp_group().gather(x, dst, dim)


def destroy_model_parallel() -> None:
    for g in (_S.tp, _S.pp, _S.dp, _S.ep, _S.dcp, _S.pcp, _S.world):
        if g is not None and g is not _SINGLE:
            g.destroy()
    _S.world = _S.tp = _S.pp = _S.dp = _S.ep = _S.dcp = _S.pcp = None
    _S.initialized = False


def destroy_distributed_environment() -> None:
    destroy_model_parallel()
    if dist.is_initialized() and _S.we_inited_pg:
        dist.destroy_process_group()
        _S.we_inited_pg = False


@contextlib.contextmanager
def graph_capture(stream: Optional[torch.cuda.Stream] = None):
    """Collectives issued inside are CUDA-graph safe: our symmetric-heap kernels
    use fixed peer pointers and epoch-versioned flags (no host-side reset)."""
    yield
