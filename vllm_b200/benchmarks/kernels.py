"""Head-to-head kernel micro-benchmarks on one B200 (tracked; the table it prints lands in profiles/kernels_bench.md).

    python -m vllm_b200.benchmarks.kernels [--out gpurun_out/kernels_bench.md] [--only gemm,fp8,attn,mla,quant]

Method (B200_PROFILING.md): CUDA events on the launching stream, >= 5 warm-up calls, every timed call preceded by a
write of a 256 MB buffer (flushes the 126 MB L2) on the same stream, median of the timed calls; library baselines
(cuBLASLt through torch, FlashInfer / the stock vLLM wheel's ops when importable) are timed the same way.  Achieved
bandwidth = algorithmic bytes / time against the measured copy bandwidth of MEASURED_PEAKS.json.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics

import torch

_FLUSH = None


def _flush():
    global _FLUSH
    if _FLUSH is None:
        _FLUSH = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")
    # READ a buffer twice the size of L2: the victims are replaced by clean lines.  (A write-based flush leaves 126 MB of
    # dirty lines whose write-back is then charged to the kernel under test: measured +40 % on TMA first-byte latency.)
    _FLUSH.view(torch.int64).sum()


def timeit(fn, iters: int = 20, warmup: int = 5) -> float:
    """median µs of fn() with the L2 flushed before every call"""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        _flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


def hbm_gbs() -> float:
    try:
        root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        return float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def bench_gemm(rows: list) -> None:
    from vllm_b200.ops import cuda as cu
    shapes = [("8B qkv", 6144, 4096), ("8B o", 4096, 4096), ("8B gate_up", 28672, 4096), ("8B down", 4096, 14336),
              ("lm_head", 128256, 4096), ("70B/TP8 qkv", 1280, 8192), ("70B/TP8 o", 8192, 1024),
              ("70B/TP8 gate_up", 7168, 8192), ("70B/TP8 down", 8192, 3584), ("8B/TP8 gate_up", 3584, 4096),
              ("8B/TP8 down", 4096, 1792)]
    peak = hbm_gbs()
    rows.append("\n## bf16 decode GEMM, M = 64 tokens (µs, L2 flushed) — stream-K vs tile/cluster split-K vs cuBLASLt\n")
    rows.append("| shape (N x K) | stream-K (ours) | tile + cluster split-K (ours, r1) | cuBLASLt | stream-K TB/s | of measured HBM |")
    rows.append("|---|---|---|---|---|---|")
    for name, N, K in shapes:
        x = torch.randn(64, K, device="cuda", dtype=torch.bfloat16) * 0.1
        w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.02
        y = torch.empty(64, N, device="cuda", dtype=torch.bfloat16)
        t_sk = timeit(lambda: cu.gemm_sk(x, w, out=y))
        t_tc = timeit(lambda: cu.gemm_tc(x, w, out=y))
        t_cb = timeit(lambda: torch.matmul(x, w.t(), out=y))
        tb = N * K * 2 / t_sk / 1e6
        rows.append(f"| {name} {N} x {K} | {t_sk:.1f} | {t_tc:.1f} | {t_cb:.1f} | {tb:.2f} | {tb * 1e3 / peak:.0%} |")
    for M in (1, 16, 128, 256):
        N, K = 6144, 4096
        x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16) * 0.1
        w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.02
        t_sk = timeit(lambda: cu.gemm_sk(x, w))
        t_tc = timeit(lambda: cu.gemm_tc(x, w))
        t_cb = timeit(lambda: torch.matmul(x, w.t()))
        tb = N * K * 2 / t_sk / 1e6
        rows.append(f"| qkv M={M} {N} x {K} | {t_sk:.1f} | {t_tc:.1f} | {t_cb:.1f} | {tb:.2f} | {tb * 1e3 / peak:.0%} |")


def bench_w4(rows: list) -> None:
    """Weight-only FP4 decode GEMM (csrc/gemm_w4.cu) against the bf16 kernel on the same shapes."""
    from vllm_b200.ops import cuda as cu
    from vllm_b200.ops import ref
    shapes = [("8B qkv", 6144, 4096), ("8B o", 4096, 4096), ("8B gate_up", 28672, 4096), ("8B down", 4096, 14336),
              ("lm_head", 128256, 4096), ("70B/TP8 gate_up", 7168, 8192), ("70B/TP8 down", 8192, 3584)]
    peak = hbm_gbs()
    rows.append("\n## weight-only FP4 (W4A16) decode GEMM, M = 64 tokens (µs, L2 flushed)\n")
    rows.append("| shape (N x K) | bf16 tile kernel | NVFP4 (e4m3 / 16) | MXFP4 (e8m0 / 32) | NVFP4 TB/s of packed bytes | of measured HBM | speed-up vs bf16 |")
    rows.append("|---|---|---|---|---|---|---|")
    for name, N, K in shapes:
        x = torch.randn(64, K, device="cuda", dtype=torch.bfloat16) * 0.1
        w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.02
        t_bf = timeit(lambda: cu.gemm_tc(x, w))
        # quantise in row chunks (the reference quantiser materialises an [R, C, 8] distance tensor)
        parts = [ref.fp4_quant(w[i:i + 4096].float(), 16) for i in range(0, N, 4096)]
        g = max(p_[2] for p_ in parts)
        codes = torch.cat([p_[0] for p_ in parts]); sc = torch.cat([p_[1] for p_ in parts])
        tiled = cu.pack_w4_tiles(codes, sc, "nvfp4")
        t_nv = timeit(lambda: cu.gemm_w4(x, tiled, N, g, "nvfp4"))
        parts = [ref.mxfp4_quant(w[i:i + 4096].float()) for i in range(0, N, 4096)]
        tiled_mx = cu.pack_w4_tiles(torch.cat([p_[0] for p_ in parts]), torch.cat([p_[1] for p_ in parts]), "mxfp4")
        t_mx = timeit(lambda: cu.gemm_w4(x, tiled_mx, N, 1.0, "mxfp4"))
        tb = tiled.numel() / t_nv / 1e6
        rows.append(f"| {name} {N} x {K} | {t_bf:.1f} | {t_nv:.1f} | {t_mx:.1f} | {tb:.2f} | {tb * 1e3 / peak:.0%} | {t_bf / t_nv:.2f}x |")
        del w, codes, sc, tiled, tiled_mx, parts
        torch.cuda.empty_cache()


def bench_fp8(rows: list) -> None:
    from vllm_b200.ops import cuda as cu
    from vllm_b200.ops import ref
    peak = hbm_gbs()
    rows.append("\n## block-scaled fp8 GEMM (DeepSeek 1x128 / 128x128 scales), tcgen05 kind::f8f6f4 (µs, L2 flushed)\n")
    rows.append("| shape (M, N x K) | ours: quant + GEMM | ours: GEMM only | bf16 stream-K same shape | cuBLASLt fp8 per-tensor (`torch._scaled_mm`) | GEMM-only TB/s (fp8 weight bytes) |")
    rows.append("|---|---|---|---|---|---|")
    for name, M, N, K in [("q_b_proj", 64, 24576, 1536), ("kv_a_proj", 64, 576, 7168), ("o_proj", 64, 7168, 16384),
                          ("dense gate_up", 64, 36864, 7168), ("dense down", 64, 7168, 18432),
                          ("expert gate_up", 64, 4096, 7168), ("prefill o_proj", 4096, 7168, 16384)]:
        x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16) * 0.1
        w = torch.randn(N, K, device="cuda") * 0.02
        wq, ws = ref.block_quant_weight_fp8(w, 128)
        wb = w.to(torch.bfloat16)
        del w
        ws = ws.contiguous()
        xq, xs = cu.per_token_group_quant_fp8(x, 128, col_major=True)
        t_all = timeit(lambda: cu.block_fp8_linear(x, wq, ws))
        t_g = timeit(lambda: cu.gemm_fp8_block(xq, xs, wq, ws))
        t_b = timeit(lambda: cu.gemm_sk(x, wb)) if M <= 256 else timeit(lambda: torch.matmul(x, wb.t()))
        try:
            one = torch.ones((), device="cuda")
            xq8 = x.to(torch.float8_e4m3fn)
            t_c = timeit(lambda: torch._scaled_mm(xq8, wq.t(), scale_a=one, scale_b=one, out_dtype=torch.bfloat16))
            t_c = f"{t_c:.1f}"
        except Exception as e:  # noqa: BLE001
            t_c = f"n/a ({type(e).__name__})"
        tb = N * K / t_g / 1e6
        rows.append(f"| {name} {M}, {N} x {K} | {t_all:.1f} | {t_g:.1f} | {t_b:.1f} | {t_c} | {tb:.2f} ({tb * 1e3 / peak:.0%}) |")
        del wq, wb


def _mk_paged(B, Hkv, D, BS, ctx, dtype=torch.bfloat16):
    nblk = (ctx + BS - 1) // BS + 1
    NB = B * nblk + 2
    bt = (torch.randperm(NB - 1, device="cuda")[:B * nblk].int() + 1).view(B, nblk).contiguous()
    sl = torch.full((B,), ctx, dtype=torch.int32, device="cuda")
    return NB, bt, sl


def bench_attn(rows: list) -> None:
    from vllm_b200.forward_context import AttentionMetadata
    from vllm_b200.ops import cuda as cu
    peak = hbm_gbs()
    rows.append("\n## paged decode attention, B = 64, D = 128, bf16 KV (µs, L2 flushed)\n")
    rows.append("| q heads / kv heads | context | ours (tcgen05 swap-AB) | FlashInfer (library: trtllm-gen / fa decode) | ours TB/s | of measured HBM |")
    rows.append("|---|---|---|---|---|---|")
    for Hq, Hkv in [(32, 8), (8, 1), (4, 1)]:
        for ctx in (1040, 4096):
            B, D, BS = 64, 128, 64
            NB, bt, sl = _mk_paged(B, Hkv, D, BS, ctx)
            cache = (torch.randn(NB, 2, Hkv, BS, D, device="cuda") * 0.3).to(torch.bfloat16)
            q = (torch.randn(B, Hq, D, device="cuda") * 0.3).to(torch.bfloat16)
            qsl = torch.arange(B + 1, dtype=torch.int32, device="cuda")
            md = AttentionMetadata(num_reqs=B, num_actual_tokens=B, max_query_len=1, max_seq_len=ctx, query_start_loc=qsl,
                                   seq_lens=sl, block_table=bt, slot_mapping=torch.empty(0), decode_only=True)
            md.kv_vt = True
            t_o = timeit(lambda: cu.paged_attention(q, cache, md, D ** -0.5))
            t_f = "n/a"
            try:
                import flashinfer
                ws = torch.empty(128 << 20, dtype=torch.uint8, device="cuda")
                kv = torch.randn(NB, 2, BS, Hkv, D, device="cuda", dtype=torch.bfloat16) * 0.3
                w = flashinfer.BatchDecodeWithPagedKVCacheWrapper(ws, "NHD", use_tensor_cores=True)
                nblk = (ctx + BS - 1) // BS
                indptr = torch.arange(B + 1, dtype=torch.int32, device="cuda") * nblk
                indices = bt[:, :nblk].reshape(-1).contiguous()
                last = torch.full((B,), (ctx - 1) % BS + 1, dtype=torch.int32, device="cuda")
                w.plan(indptr, indices, last, Hq, Hkv, D, BS, q_data_type=torch.bfloat16, kv_data_type=torch.bfloat16)
                t_f = f"{timeit(lambda: w.run(q, kv)):.1f}"
            except Exception as e:  # noqa: BLE001
                t_f = f"n/a ({type(e).__name__})"
            byt = B * ctx * Hkv * D * 2 * 2
            tb = byt / t_o / 1e6
            rows.append(f"| {Hq} / {Hkv} | {ctx} | {t_o:.1f} | {t_f} | {tb:.2f} | {tb * 1e3 / peak:.0%} |")


def bench_mla(rows: list) -> None:
    from vllm_b200.ops import cuda as cu
    peak = hbm_gbs()
    rows.append("\n## MLA decode over the 576-d latent cache, B = 64, bf16 (µs, L2 flushed)\n")
    rows.append("| heads | context | ours (tcgen05, latent tile shared by both products) | PyTorch bmm path it replaces | FlashInfer MLA (library) | ours TB/s (latent bytes) | of measured HBM |")
    rows.append("|---|---|---|---|---|---|---|")
    for H in (16, 128):
        for ctx in (1024, 4096):
            B, BS = 64, 64
            NB, bt, sl = _mk_paged(B, 1, 576, BS, ctx)
            cache = (torch.randn(NB, BS, 576, device="cuda") * 0.3).to(torch.bfloat16)
            qn = (torch.randn(B, H, 512, device="cuda") * 0.3).to(torch.bfloat16)
            qp = (torch.randn(B, H, 64, device="cuda") * 0.3).to(torch.bfloat16)
            t_o = timeit(lambda: cu.mla_decode_tc(qn, qp, cache, bt, sl, 0.07, ctx))

            def torch_path():
                nbk = (ctx + BS - 1) // BS
                lat = cache[bt[:, :nbk].long()].reshape(B, nbk * BS, -1)
                kc, kp = lat[..., :512], lat[..., 512:]
                att = (torch.bmm(qn, kc.transpose(1, 2)) + torch.bmm(qp, kp.transpose(1, 2))).float() * 0.07
                return torch.bmm(torch.softmax(att, -1).to(qn.dtype), kc)
            t_t = timeit(torch_path, iters=10)
            t_f = "n/a"
            try:
                import flashinfer
                ws = torch.empty(128 << 20, dtype=torch.uint8, device="cuda")
                w = flashinfer.mla.BatchMLAPagedAttentionWrapper(ws)
                nblk = (ctx + BS - 1) // BS
                qo_indptr = torch.arange(B + 1, dtype=torch.int32, device="cuda")
                kv_indptr = torch.arange(B + 1, dtype=torch.int32, device="cuda") * nblk
                kv_indices = bt[:, :nblk].reshape(-1).contiguous()
                w.plan(qo_indptr, kv_indptr, kv_indices, sl, H, 512, 64, BS, False, 0.07, torch.bfloat16, torch.bfloat16)
                ckv, kpe = cache[..., :512].contiguous(), cache[..., 512:].contiguous()
                t_f = f"{timeit(lambda: w.run(qn, qp, ckv, kpe)):.1f}"
            except Exception as e:  # noqa: BLE001
                t_f = f"n/a ({type(e).__name__})"
            byt = B * ctx * 576 * 2
            tb = byt / t_o / 1e6
            rows.append(f"| {H} | {ctx} | {t_o:.1f} | {t_t:.1f} | {t_f} | {tb:.2f} | {tb * 1e3 / peak:.0%} |")


def bench_quant(rows: list) -> None:
    from vllm_b200.ops import cuda as cu
    from vllm_b200.ops import ref
    peak = hbm_gbs()
    rows.append("\n## fused norm / activation + fp8 quantisation, T = 4096 tokens (µs, L2 flushed)\n")
    rows.append("| op | ours (one kernel) | unfused (our norm/act kernel + PyTorch quant) | ours GB/s | of measured HBM |")
    rows.append("|---|---|---|---|---|")
    T, H = 4096, 7168
    x = torch.randn(T, H, device="cuda", dtype=torch.bfloat16)
    r = torch.randn(T, H, device="cuda", dtype=torch.bfloat16)
    w = torch.ones(H, device="cuda", dtype=torch.bfloat16)
    t1 = timeit(lambda: cu.rms_norm_quant(x, w, 1e-6, residual=r, mode="group"))
    t2 = timeit(lambda: ref.per_token_group_quant_fp8(cu.fused_add_rms_norm(x.clone(), r, w, 1e-6)[0], 128), iters=10)
    gb = T * H * (2 + 2 + 2 + 1) / t1 / 1e3
    rows.append(f"| residual + RMSNorm + 1x128 fp8 quant [{T} x {H}] | {t1:.1f} | {t2:.1f} | {gb:.0f} | {gb / peak:.0%} |")
    g = torch.randn(T, 2 * 2048, device="cuda", dtype=torch.bfloat16)
    t1 = timeit(lambda: cu.act_mul_quant(g, "silu", mode="group"))
    t2 = timeit(lambda: ref.per_token_group_quant_fp8(cu.act_and_mul(g, "silu"), 128), iters=10)
    gb = T * 2048 * (4 + 1) / t1 / 1e3
    rows.append(f"| SiLU·mul + 1x128 fp8 quant [{T} x 2x2048] | {t1:.1f} | {t2:.1f} | {gb:.0f} | {gb / peak:.0%} |")
    t1 = timeit(lambda: cu.per_token_group_quant_fp8(x, 128, col_major=True))
    gb = T * H * 3 / t1 / 1e3
    rows.append(f"| 1x128 fp8 quant [{T} x {H}] | {t1:.1f} | — | {gb:.0f} | {gb / peak:.0%} |")


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/kernels_bench.md")
    ap.add_argument("--only", default="gemm,w4,fp8,attn,mla,quant")
    args = ap.parse_args()
    rows = [f"# Kernel micro-benchmarks on {torch.cuda.get_device_name()} (measured HBM copy {hbm_gbs():.0f} GB/s)"]
    for name, fn in (("gemm", bench_gemm), ("w4", bench_w4), ("fp8", bench_fp8), ("attn", bench_attn), ("mla", bench_mla),
                     ("quant", bench_quant)):
        if name in args.only.split(","):
            try:
                fn(rows)
            except Exception as e:  # noqa: BLE001 - keep the sections that work
                rows.append(f"\n(section {name} failed: {type(e).__name__}: {e})")
            torch.cuda.empty_cache()
    text = "\n".join(rows) + "\n"
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        f.write(text)
    print(text)


if __name__ == "__main__":
    main()
