"""Multi-GPU checks, launched by tests/test_multigpu.py through torchrun (one process per GPU).

Mirrors the reference's tests/distributed/test_custom_all_reduce.py / test_symm_mem_allreduce.py strategy:
our collectives vs NCCL on the same data (integer-valued inputs so the sums are exact), eager and under
CUDA-graph capture, then an end-to-end TP run against the TP=1 result.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("VLLM_LOGGING_LEVEL", "WARNING")

import torch
import torch.distributed as dist


def check(cond, msg):
    if not cond:
        raise AssertionError(f"[rank {dist.get_rank()}] {msg}")


def test_collectives():
    from vllm_b200.ops import ref
    from vllm_b200.parallel import state as ps
    tp = ps.get_tp_group()
    comm = tp.device_comm
    check(comm is not None, "SymmComm was not created")
    rank, world = tp.rank_in_group, tp.world_size
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    for dtype in (torch.bfloat16, torch.float16):
        for T, H in ((1, 1024), (7, 4096), (64, 4096), (64, 8192), (300, 2048), (700, 1024)):
            x = torch.randint(-8, 9, (T, H), device=dev, generator=g).to(dtype)
            exp = x.clone()
            dist.all_reduce(exp, group=tp.device_group)
            for _ in range(3):  # repeated calls: epochs + double buffering
                got = comm.all_reduce(x)
                check(torch.equal(got, exp), f"all_reduce mismatch T={T} H={H} {dtype}")
            # fused residual + RMSNorm (residual rows are owned by rank = row // ceil(T/world))
            gr = torch.Generator(device=dev).manual_seed(99)
            res = torch.randint(-4, 5, (T, H), device=dev, generator=gr).to(dtype)
            w = (1 + 0.1 * torch.randn(H, device=dev, generator=gr)).to(dtype)
            res_mine = res.clone()
            normed = comm.allreduce_rows(x, T, H, dtype, res_mine, w, 1e-5).clone()
            e_norm, e_res = ref.fused_add_rms_norm(exp, res, w, 1e-5)
            torch.testing.assert_close(normed.float(), e_norm.float(), atol=2e-2, rtol=2e-2)
            per = (T + world - 1) // world
            r0, r1 = rank * per, min(T, rank * per + per)
            if r1 > r0:
                check(torch.equal(res_mine[r0:r1], e_res[r0:r1]), "owned residual rows mismatch")
        # all-gather (logits)
        x = torch.randn(5, 4096, device=dev, generator=g).to(dtype)
        exp = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(exp, x, group=tp.device_group)
        got = comm.all_gather(x, 1)
        check(torch.equal(got, torch.cat(exp, dim=1)), "all_gather mismatch")
    # CUDA-graph capture + replay
    x = torch.randint(-8, 9, (64, 4096), device=dev, generator=g).to(torch.bfloat16)
    exp = x.clone()
    dist.all_reduce(exp, group=tp.device_group)
    comm.all_reduce(x)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        comm.all_reduce(x)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(graph):
        y = comm.all_reduce(x)
    for _ in range(5):
        graph.replay()
    torch.cuda.synchronize()
    check(torch.equal(y, exp), "graph-replayed all_reduce mismatch")
    # fused GEMM + all-reduce + norm vs unfused
    K, H, T = 1024, 4096, 64
    gw = torch.Generator(device=dev).manual_seed(7 + rank)
    xk = (torch.randn(T, K, device=dev, generator=gw) * 0.5).to(torch.bfloat16)
    wk = (torch.randn(H, K, device=dev, generator=gw) * 0.05).to(torch.bfloat16)
    gr = torch.Generator(device=dev).manual_seed(5)
    res = torch.randn(T, H, device=dev, generator=gr).to(torch.bfloat16)
    nw = (1 + 0.1 * torch.randn(H, device=dev, generator=gr)).to(torch.bfloat16)
    y = torch.nn.functional.linear(xk, wk)
    dist.all_reduce(y, group=tp.device_group)
    e_norm, e_res = ref.fused_add_rms_norm(y, res, nw, 1e-5)
    res_mine = res.clone()
    out = comm.gemm_allreduce_rmsnorm(xk, wk, res_mine, nw, 1e-5)
    check(out is not None, "fused path declined")
    torch.testing.assert_close(out[0].float(), e_norm.float(), atol=6e-2, rtol=3e-2)
    # single-kernel paths (one-shot / two-shot Lamport): varying token counts, both hidden sizes, many calls
    modes = [0, 1]
    for mode, (H, K) in [(m, hk) for m in modes for hk in ((4096, 2048), (8192, 1024), (4096, 14336 // world // 8 * 8))]:
        comm.fused_mode = mode
        gw = torch.Generator(device=dev).manual_seed(11 + rank)
        wk = (torch.randn(H, K, device=dev, generator=gw) * 0.05).to(torch.bfloat16)
        nw = (1 + 0.1 * torch.randn(H, device=dev, generator=torch.Generator(device=dev).manual_seed(3))).to(torch.bfloat16)
        for it, T in enumerate((64, 7, 128, 1, 64, 33, 64, 64)):
            xk = (torch.randn(T, K, device=dev, generator=gw) * 0.5).to(torch.bfloat16)
            res = torch.randn(T, H, device=dev, generator=torch.Generator(device=dev).manual_seed(100 + it)).to(torch.bfloat16)
            y = torch.nn.functional.linear(xk, wk)
            dist.all_reduce(y, group=tp.device_group)
            e_norm, e_res = ref.fused_add_rms_norm(y, res, nw, 1e-5)
            res_mine = res.clone()
            out = comm.gemm_allreduce_rmsnorm(xk, wk, res_mine, nw, 1e-5)
            torch.testing.assert_close(out[0].float(), e_norm.float(), atol=6e-2, rtol=3e-2)
            per = (T + world - 1) // world if mode == 1 else T      # two-shot keeps the residual sharded by row owner
            r0, r1 = (rank * per, min(T, rank * per + per)) if mode == 1 else (0, T)
            if r1 > r0:
                torch.testing.assert_close(res_mine[r0:r1].float(), e_res[r0:r1].float(), atol=6e-2, rtol=3e-2)
        # CUDA-graph replay of the fused kernel (device-side buffer rotation)
        T = 64
        xk = (torch.randn(T, K, device=dev, generator=gw) * 0.5).to(torch.bfloat16)
        res0 = torch.randn(T, H, device=dev, generator=torch.Generator(device=dev).manual_seed(7)).to(torch.bfloat16)
        y = torch.nn.functional.linear(xk, wk)
        dist.all_reduce(y, group=tp.device_group)
        e_norm, _ = ref.fused_add_rms_norm(y, res0, nw, 1e-5)
        res_g = res0.clone()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            res_g.copy_(res0)
            og = comm.gemm_allreduce_rmsnorm(xk, wk, res_g, nw, 1e-5)[0]
        for _ in range(7):
            graph.replay()
        torch.cuda.synchronize()
        torch.testing.assert_close(og.float(), e_norm.float(), atol=6e-2, rtol=3e-2)
    comm.fused_mode = 1
    if rank == 0:
        print("collectives ok (multicast=%s)" % comm.use_mc, flush=True)


def test_fused_stress():
    """Protocol stress for csrc/gemm_ar.cu (reference strategy: tests/distributed/test_custom_all_reduce.py:33-100):
    integer-valued operands so that every partial and every sum is exact in bf16/fp16 → the owned residual rows must be
    BIT-identical to NCCL's; rows whose partial sums are all -0.0 (the Lamport sentinel bit pattern) and denormal-heavy
    rows; 1000 CUDA-graph replays interleaving four token counts so the three rotating buffers and the device-side call
    counter see every phase; both protocols (one-shot / two-shot)."""
    from vllm_b200.parallel import state as ps
    tp = ps.get_tp_group()
    comm = tp.device_comm
    rank, world = tp.rank_in_group, tp.world_size
    dev = torch.device("cuda", torch.cuda.current_device())
    H, K = 4096, 128
    for dtype in (torch.bfloat16, torch.float16):
        for mode in (1, 0):
            comm.fused_mode = mode
            g = torch.Generator(device=dev).manual_seed(1000 + rank)
            w = torch.randint(-1, 2, (H, K), device=dev, generator=g).to(dtype)
            tiny = 6e-8 if dtype == torch.float16 else 1e-21      # tiny * tiny underflows to (signed) zero in `dtype`
            w[:, 0] = tiny
            nw = torch.ones(H, device=dev, dtype=dtype)
            graphs = {}
            for T in (1, 7, 64, 128):
                x = torch.randint(-1, 2, (T, K), device=dev, generator=g).to(dtype)
                x[:, 0] = 0.0                 # only rows 0 / 1 use the tiny-scale column (keeps every other sum an exact integer)
                if T > 1:
                    x[0] = 0.0
                    x[0, 0] = -tiny       # every partial of row 0 rounds to -0.0 (the sentinel bit pattern): must not hang
                    x[1] = 0.0
                    x[1, 0] = tiny        # +0.0 / denormal-range row
                res0 = torch.randint(-3, 4, (T, H), device=dev,
                                     generator=torch.Generator(device=dev).manual_seed(50 + T)).to(dtype)
                y = torch.nn.functional.linear(x.float(), w.float())
                dist.all_reduce(y, group=tp.device_group)
                exp_res = (y + res0.float()).to(dtype)            # exact: small integers
                res = res0.clone()
                out = comm.gemm_allreduce_rmsnorm(x, w, res, nw, 1e-5)   # eager call (also arms the buffers for this H)
                check(out is not None, "fused path declined")
                torch.cuda.synchronize()
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph):
                    res.copy_(res0)
                    og = comm.gemm_allreduce_rmsnorm(x, w, res, nw, 1e-5)[0]
                graphs[T] = (gph, res, og, exp_res, x, res0)     # the graph reads x / res0: keep them alive
            order = torch.randint(0, 4, (1000,), generator=torch.Generator().manual_seed(77)).tolist()   # same on all ranks
            keys = [1, 7, 64, 128]
            for i, o in enumerate(order):
                T = keys[o]
                gph, res, og, exp_res = graphs[T][:4]
                gph.replay()
                if i % 50 == 0 or i > 990:
                    torch.cuda.synchronize()
                    per = (T + world - 1) // world if mode == 1 else T
                    r0, r1 = (rank * per, min(T, rank * per + per)) if mode == 1 else (0, T)
                    if r1 > r0:
                        check(torch.equal(res[r0:r1], exp_res[r0:r1]), f"replay {i}: residual rows differ (T={T}, mode={mode}, {dtype})")
                    e = exp_res.float()
                    e_norm = e * torch.rsqrt(e.pow(2).mean(-1, keepdim=True) + 1e-5)
                    torch.testing.assert_close(og.float(), e_norm, atol=3e-2, rtol=2e-2)
            torch.cuda.synchronize()
    comm.fused_mode = 1
    if rank == 0:
        print("fused stress ok", flush=True)


def test_fused_moe():
    """Fused MoE tail under EP inside the TP group: every rank owns E / W experts, the w2 grouped-GEMM epilogue red.adds
    weighted rows into the owner rank's accumulator over NVLink, the finish kernel normalises and multicasts."""
    from vllm_b200.ops import ref
    from vllm_b200.ops.moe import fused_moe_reduce_norm
    from vllm_b200.parallel import state as ps
    tp = ps.get_tp_group()
    comm = tp.device_comm
    rank, world = tp.rank_in_group, tp.world_size
    dev = torch.device("cuda", torch.cuda.current_device())
    dt = torch.bfloat16
    E, k, H, I = 8, 2, 2048, 1024
    g = torch.Generator(device=dev).manual_seed(4321)          # identical on every rank
    w13 = (torch.randn(E, 2 * I, H, device=dev, generator=g) * 0.03).to(dt)
    w2 = (torch.randn(E, H, I, device=dev, generator=g) * 0.03).to(dt)
    nw = (1 + 0.1 * torch.randn(H, device=dev, generator=g)).to(dt)
    per_e = E // world
    emap = torch.full((E,), -1, dtype=torch.int32, device=dev)
    emap[rank * per_e:(rank + 1) * per_e] = torch.arange(per_e, dtype=torch.int32, device=dev)
    w13_l, w2_l = w13[rank * per_e:(rank + 1) * per_e].contiguous(), w2[rank * per_e:(rank + 1) * per_e].contiguous()
    for it, T in enumerate((64, 7, 200, 1, 64, 64)):
        x = (torch.randn(T, H, device=dev, generator=g) * 0.5).to(dt)
        logits = torch.randn(T, E, device=dev, generator=g)
        res = torch.randn(T, H, device=dev, generator=g).to(dt)
        tw, ids = ref.topk_softmax(logits, k, True)
        exp = ref.fused_moe(x, w13, w2, tw, ids, "silu", None)
        e_norm, e_res = ref.fused_add_rms_norm(exp.clone(), res.clone(), nw, 1e-5)
        r2 = res.clone()
        out = fused_moe_reduce_norm(x, w13_l, w2_l, tw.float(), ids.int(), "silu", emap, r2, nw, 1e-5, False, comm)
        check(out is not None, "fused MoE path declined")
        torch.testing.assert_close(out[0].float(), e_norm.float(), atol=6e-2, rtol=4e-2)
        per = (T + world - 1) // world
        r0, r1 = rank * per, min(T, rank * per + per)
        if r1 > r0:
            torch.testing.assert_close(r2[r0:r1].float(), e_res[r0:r1].float(), atol=6e-2, rtol=4e-2)
    # CUDA-graph replay (rotating accumulators are re-zeroed by the finish kernel)
    T = 64
    x = (torch.randn(T, H, device=dev, generator=g) * 0.5).to(dt)
    logits = torch.randn(T, E, device=dev, generator=g)
    res0 = torch.randn(T, H, device=dev, generator=g).to(dt)
    tw, ids = ref.topk_softmax(logits, k, True)
    twf, idsi = tw.float().contiguous(), ids.int().contiguous()
    exp = ref.fused_moe(x, w13, w2, tw, ids, "silu", None)
    e_norm, _ = ref.fused_add_rms_norm(exp.clone(), res0.clone(), nw, 1e-5)
    rg = res0.clone()
    torch.cuda.synchronize()
    gph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gph):
        rg.copy_(res0)
        og = fused_moe_reduce_norm(x, w13_l, w2_l, twf, idsi, "silu", emap, rg, nw, 1e-5, False, comm)[0]
    for _ in range(10):
        gph.replay()
    torch.cuda.synchronize()
    torch.testing.assert_close(og.float(), e_norm.float(), atol=6e-2, rtol=4e-2)
    if rank == 0:
        print("fused moe ok", flush=True)


def test_e2e_tp(expected_path):
    import json
    from vllm_b200 import LLM, SamplingParams
    from vllm_b200.model_loader import initialize_dummy_weights
    world = dist.get_world_size() if dist.is_initialized() else int(os.environ["WORLD_SIZE"])
    llm = LLM("llama-small", tensor_parallel_size=world, distributed_executor_backend="external_launcher",
              max_model_len=1024, max_num_seqs=16, max_num_batched_tokens=1024, gpu_memory_utilization=0.3,
              load_format="dummy", cudagraph_capture_sizes=[1, 2, 4, 8])
    from tests.utils import llama_full_weights
    runner = llm.llm_engine.engine_core.engine_core.model_executor.worker.runner
    with torch.no_grad():
        runner.model.load_weights(llama_full_weights(llm.llm_engine.model_config.hf_config, seed=3, scale=0.05))
    prompts = [{"prompt_token_ids": [(13 * i + 7 * j) % 4000 + 5 for j in range(20 + 37 * i)]} for i in range(5)]
    outs = llm.generate(prompts, SamplingParams(max_tokens=12, temperature=0.0, ignore_eos=True, logprobs=0,
                                                prompt_logprobs=0))
    toks = [o.outputs[0].token_ids for o in outs]
    # teacher-forced quantities (no feedback): logprob of every prompt token given its prefix
    lps = [[list(d.values())[0].logprob for d in o.prompt_logprobs[1:]] for o in outs]
    if int(os.environ.get("RANK", 0)) == 0:
        with open(expected_path + f".tp{world}", "w") as f:
            json.dump({"tokens": toks, "logprobs": lps}, f)
        print("e2e tp done", flush=True)
    llm.shutdown()


if __name__ == "__main__":
    mode = sys.argv[1]
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local_rank)
    if mode in ("collectives", "stress", "moe"):
        from vllm_b200.parallel import state as ps
        ps.init_distributed_environment(world, rank, local_rank, "env://")
        ps.initialize_model_parallel(tp=world)
        {"collectives": test_collectives, "stress": test_fused_stress, "moe": test_fused_moe}[mode]()
        dist.barrier()
        ps.destroy_distributed_environment()
    elif mode == "e2e":
        test_e2e_tp(sys.argv[2])
