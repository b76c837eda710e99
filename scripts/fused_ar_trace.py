"""Phase time line of the fused RowParallel kernel (csrc/gemm_ar.cu) on the live NVSwitch domain.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29655 \
        scripts/fused_ar_trace.py [--tokens 64] [--model llama-3-8b]

For the two RowParallel shapes of the model at TP=N (o_proj, down_proj) it prints, per rank-0 JSON line: fused µs per
call inside a graph of back-to-back calls (max over ranks), the local GEMM + add-norm time of the same shape, NCCL
all-reduce + add-norm of the same [T, H] for reference, and the per-phase medians/maxima of the kernel's own
%globaltimer stamps.  The summary that is kept lives in profiles/.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("VLLM_LOGGING_LEVEL", "WARNING")

import torch
import torch.distributed as dist

SHAPES = {  # model: (hidden, q_size, intermediate)
    "llama-3-8b": (4096, 4096, 14336),
    "llama-3-70b": (8192, 8192, 28672),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=64)
    ap.add_argument("--model", default="llama-3-8b")
    ap.add_argument("--iters", type=int, default=16)
    ap.add_argument("--flashinfer", action="store_true", help="also time cuBLAS GEMM + FlashInfer allreduce_fusion")
    args = ap.parse_args()
    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    from vllm_b200.benchmarks.comm_probe import _time_graph, trace_fused
    from vllm_b200.ops import cuda_ops
    from vllm_b200.parallel import state as ps
    ps.init_distributed_environment(world, rank, local_rank, "env://")
    ps.initialize_model_parallel(tp=world)
    tp = ps.get_tp_group()
    comm = tp.device_comm
    cu = cuda_ops()
    dev = torch.device("cuda", local_rank)
    H, Q, I = SHAPES[args.model]
    T = args.tokens
    res = {"world": world, "tokens": T, "model": args.model, "paths": {}}
    fi_ws = {}
    for name, K in (("o_proj", Q // world), ("down_proj", I // world)):
        x = (torch.randn(T, K, device=dev) * 0.05).bfloat16()
        w = (torch.randn(H, K, device=dev) * 0.02).bfloat16()
        r = torch.zeros(T, H, device=dev, dtype=torch.bfloat16)
        nw = torch.ones(H, device=dev, dtype=torch.bfloat16)
        y = torch.empty(T, H, device=dev, dtype=torch.bfloat16)

        def fused():
            comm.gemm_allreduce_rmsnorm(x, w, r, nw, 1e-5, False)

        def local():
            cu.gemm_tc(x, w, None, out=y)
            cu.fused_add_rms_norm(y, r, nw, 1e-5, False)

        def nccl():
            cu.gemm_tc(x, w, None, out=y)
            dist.all_reduce(y, group=tp.device_group)
            cu.fused_add_rms_norm(y, r, nw, 1e-5, False)

        # the reference's path: library GEMM → FlashInfer TRT-LLM allreduce_fusion (AR + residual + RMSNorm)
        fi_fn, fi_note = None, None
        try:
            if not args.flashinfer:
                raise RuntimeError("not requested (--flashinfer)")
            import flashinfer.comm as fic
            ws = fi_ws.get(H)
            if ws is None:
                ws = fi_ws[H] = fic.create_allreduce_fusion_workspace(
                    backend="trtllm", world_size=world, rank=rank, max_token_num=128, hidden_dim=H, dtype=torch.bfloat16,
                    group=tp.cpu_group if hasattr(tp, "cpu_group") else None)
            r_out = torch.empty_like(r)
            n_out = torch.empty_like(r)

            def fi_fn():
                torch.matmul(x, w.t(), out=y)
                fic.allreduce_fusion(input=y, workspace=ws, pattern=fic.AllReduceFusionPattern.kARResidualRMSNorm,
                                     launch_with_pdl=True, trigger_completion_at_end=True, residual_out=r_out,
                                     norm_out=n_out, residual_in=r, rms_gamma=nw, rms_eps=1e-5, use_oneshot=None,
                                     fp32_acc=True)
            fi_fn()
            torch.cuda.synchronize()
        except Exception as e:      # the comparison arm is optional
            fi_fn, fi_note = None, f"{type(e).__name__}: {str(e)[:160]}"

        if comm.gemm_allreduce_rmsnorm(x, w, r, nw, 1e-5, False) is None:
            res["paths"][name] = {"note": "not covered"}
            continue
        torch.cuda.synchronize()
        dist.barrier()
        ts = []
        for fn in (fused, local, nccl):
            dist.barrier()
            ts.append(_time_graph(fn, args.iters))
        t_fi = None
        try:
            ok = torch.tensor([1 if fi_fn is not None else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()):
                dist.barrier()
                t_fi = torch.tensor([_time_graph(fi_fn, args.iters)], device=dev, dtype=torch.float64)
                dist.all_reduce(t_fi, op=dist.ReduceOp.MAX)
                t_fi = round(t_fi.item(), 2)
        except Exception as e:
            fi_note = f"{type(e).__name__}: {str(e)[:160]}"
        t = torch.tensor(ts, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        tr = trace_fused(fused, iters=8)
        allt = [None] * world
        dist.all_gather_object(allt, tr)
        res["paths"][name] = {"K_per_rank": K, "H": H, "fused_us": round(t[0].item(), 2), "gemm_norm_local_us": round(t[1].item(), 2),
                              "gemm_nccl_ar_norm_us": round(t[2].item(), 2),
                              "cublas_gemm_flashinfer_ar_norm_us": t_fi, "flashinfer_note": fi_note, "trace_rank0": allt[0],
                              "trace_span_us_by_rank": [a.get("span_us") for a in allt],
                              "trace_wait_reduce_by_rank": [a.get("wait_reduce_bcast") for a in allt],
                              "trace_collect_by_rank": [a.get("collect_norm") for a in allt]}
    if rank == 0:
        print(json.dumps(res))
    dist.barrier()
    ps.destroy_distributed_environment()


if __name__ == "__main__":
    main()
