"""Exposed-communication probe for the fused RowParallel kernel (BASELINE.json: "exposed (non-overlapped)
communication time per decode step", "each fused path reported as achieved fraction of its roofline").

For each RowParallel shape of the running model (attention output projection, MLP down projection) it times,
inside a CUDA graph of ``iters`` back-to-back calls on the live TP communicator,

* ``fused``: ``SymmComm.gemm_allreduce_rmsnorm``  (ONE kernel: tcgen05 GEMM → NVSwitch all-reduce → +residual →
  RMSNorm, ``csrc/gemm_ar.cu``), and
* ``local``: the same GEMM (``gemm_tc``) + ``fused_add_rms_norm`` with no peer traffic,

max over ranks.  exposed = fused − local per call; × layers = exposed communication per decode step.  The
roofline of the fused path is the slower of its weight bytes over the measured HBM bandwidth and the bytes it
moves over NVLink (two-shot: every partial row travels to its owner, the reduced row is multicast back:
2·T·H·2·(W−1)/W bytes per rank and direction) over the measured 770 GB/s per direction (B200_PROFILING.md).
"""
from __future__ import annotations

import json
import os

import torch

NVLINK_GBS = 770.0


def _hbm_gbs() -> float:
    try:
        root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        return float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def _time_graph(fn, iters: int, reps: int = 5) -> float:
    """µs per call of ``fn`` replayed ``iters`` times inside one CUDA graph (best of ``reps`` replays)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / iters)
    return best


PHASES = ("pdl_wait", "gemm", "splitk_push", "wait_reduce_bcast", "rearm", "collect_norm")


def trace_fused(fn, iters: int = 8) -> dict:
    """Per-phase time line of the LAST of ``iters`` back-to-back fused calls (graph replay): the kernel's epilogue
    stamps %globaltimer at its phase boundaries per CTA.  Returns median / max over CTAs of every phase (µs) and the
    span from the first CTA's start to the last CTA's end."""
    from vllm_b200.ops import cuda_ops
    cu = cuda_ops()
    tr = torch.zeros(160 * 8, dtype=torch.int64, device="cuda")
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    cu.gemm_ar_set_trace(tr)
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(iters):
                fn()
        g.replay()
        torch.cuda.synchronize()
        tr.zero_()
        g.replay()
        torch.cuda.synchronize()
    finally:
        cu.gemm_ar_set_trace(None)
    t = tr.view(160, 8).cpu()
    t = t[t[:, 0] > 0].double()
    if t.numel() == 0:
        return {}
    out = {"ctas": int(t.shape[0]), "span_us": round(float(t[:, 6].max() - t[:, 0].min()) / 1e3, 2)}
    for i, name in enumerate(PHASES):
        d = (t[:, i + 1] - t[:, i]) / 1e3
        out[name] = [round(float(d.median()), 2), round(float(d.max()), 2)]
    out["start_skew_us"] = round(float(t[:, 0].max() - t[:, 0].min()) / 1e3, 2)
    return out


def probe_row_parallel(runner, T: int, iters: int = 16) -> dict:
    import torch.distributed as dist

    from vllm_b200.layers.linear import RowParallelLinear
    from vllm_b200.ops import cuda_ops
    from vllm_b200.parallel.state import get_tp_group
    g = get_tp_group()
    comm = g.device_comm
    if comm is None or g.world_size == 1:
        return {"note": "no device communicator (TP=1)"}
    cu = cuda_ops()
    model = runner.model
    layers = getattr(getattr(model, "model", model), "layers", None)
    if not layers:
        return {"note": "model has no decoder layer list"}
    L = len(layers)
    shapes = []
    for name, mod in layers[0].named_modules():
        if isinstance(mod, RowParallelLinear) and getattr(mod, "weight", None) is not None and mod.weight.dim() == 2:
            shapes.append((name, mod.weight))
    W = g.world_size
    hbm = _hbm_gbs()
    out = {"tokens": T, "world": W, "layers": L, "paths": {}}
    exposed = 0.0
    for name, w in shapes:
        H, K = w.shape
        x = torch.randn(T, K, device=w.device, dtype=w.dtype) * 0.05
        res = torch.zeros(T, H, device=w.device, dtype=w.dtype)
        nw = torch.ones(H, device=w.device, dtype=w.dtype)
        y = torch.empty(T, H, device=w.device, dtype=w.dtype)

        def fused():
            comm.gemm_allreduce_rmsnorm(x, w, res, nw, 1e-5, False)

        def local():
            cu.gemm_tc(x, w, None, out=y)
            cu.fused_add_rms_norm(y, res, nw, 1e-5, False)
        if comm.gemm_allreduce_rmsnorm(x, w, res, nw, 1e-5, False) is None:
            out["paths"][name] = {"note": "shape not covered by the fused kernel"}
            continue
        torch.cuda.synchronize()
        g.barrier()
        t_f = _time_graph(fused, iters)
        g.barrier()
        t_l = _time_graph(local, iters)
        t = torch.tensor([t_f, t_l], device=w.device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=g.device_group)
        t_f, t_l = t.tolist()
        link_bytes = 2.0 * T * H * w.element_size() * (W - 1) / W
        w_bytes = float(H * K * w.element_size())
        roof_us = max(w_bytes / (hbm * 1e3), link_bytes / (NVLINK_GBS * 1e3))
        out["paths"][name] = {"H": H, "K_per_rank": K, "fused_us": round(t_f, 2), "gemm_plus_norm_us": round(t_l, 2),
                              "exposed_us": round(t_f - t_l, 2), "nvlink_bytes_per_rank": int(link_bytes),
                              "weight_bytes": int(w_bytes), "roofline_us": round(roof_us, 2),
                              "fraction_of_roofline": round(roof_us / t_f, 3)}
        try:
            out["paths"][name]["trace_us_median_max"] = trace_fused(fused)
        except Exception as e:      # diagnostic only
            out["paths"][name]["trace_us_median_max"] = {"error": str(e)[:200]}
        exposed += (t_f - t_l) * L
    out["exposed_comm_us_per_step"] = round(exposed, 1)
    return out
