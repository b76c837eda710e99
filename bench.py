#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): decode output tokens/s for a fixed batch of 64 synthetic requests
(1024-token prompts), Llama-3-8B at TP = N GPUs (the scaling line); ``--model llama-3-70b`` (TP8 headline),
``--model mixtral-8x7b`` and ``--model deepseek-small`` select the other BASELINE configs.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W [--impl reference]

Both arms print the SAME ``metric`` string and measure the SAME thing:

``value``  ours: 64 * K / (device time of K consecutive engine decode steps, CUDA events, max over ranks).
           reference: 64 * K / (decode wall time of K steps, see ``e2e``) — an external process cannot be
           event-timed, so for the reference ``value`` == ``e2e.value``.
``e2e``    BOTH arms, identical method (the reference's own latency harness, vllm/benchmarks/latency.py:34-178,
           times whole ``LLM.generate`` calls): the whole batch is prefilled in one engine step
           (max_num_batched_tokens = 64 * 1024), then
               t_decode = min_r wall(LLM.generate(max_tokens = K + 1)) - min_r wall(LLM.generate(max_tokens = 1))
           over R interleaved repetitions after a full warm-up generate; e2e = 64 * K / t_decode.  Every engine
           step inside copies its inputs host -> device from pinned memory and reads the sampled ids back.

``--impl reference`` runs the UNMODIFIED reference (baseline/_ref if the offline install worked, else the stock
vLLM wheel of the image — DESIGN.md §0) on its DEFAULT engine: multiprocess engine core, async scheduling,
torch.compile + CUDA graphs, its own custom all-reduce / NCCL.  Under torchrun only rank 0 drives it (it spawns
its own TP workers, one per GPU); the other torchrun ranks stay off the GPUs and wait.
"""
from __future__ import annotations

import argparse
import json
import os
os.environ.setdefault("VLLM_B200_NO_USAGE_STATS", "1")
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BASELINE_TOK_S = None  # BASELINE.json "published" is empty: no published number to divide by
METRIC = ("decode output tokens/s (whole box): batch of 64 synthetic requests x 1024-token prompts, "
          "K decode steps after one-step prefill, max over ranks")
DATA = "synthetic prompts (uniform random token ids), random-init weights (dummy loader)"
REPS = 3

# HF config.json of the benchmarked architectures (random-init weights).  Kept literal here so that the reference
# arm needs nothing from this repository.
_LLAMA = dict(architectures=["LlamaForCausalLM"], model_type="llama", hidden_act="silu", rms_norm_eps=1e-5,
              max_position_embeddings=8192, torch_dtype="bfloat16", tie_word_embeddings=False, bos_token_id=128000,
              eos_token_id=128001, vocab_size=128256, rope_theta=500000.0)
REF_CONFIGS = {
    "llama-3-8b": dict(_LLAMA, hidden_size=4096, intermediate_size=14336, num_hidden_layers=32,
                       num_attention_heads=32, num_key_value_heads=8),
    "llama-3-70b": dict(_LLAMA, hidden_size=8192, intermediate_size=28672, num_hidden_layers=80,
                        num_attention_heads=64, num_key_value_heads=8),
    "mixtral-8x7b": dict(architectures=["MixtralForCausalLM"], model_type="mixtral", hidden_act="silu",
                         rms_norm_eps=1e-5, max_position_embeddings=32768, torch_dtype="bfloat16",
                         tie_word_embeddings=False, bos_token_id=1, eos_token_id=2, vocab_size=32000,
                         rope_theta=1000000.0, hidden_size=4096, intermediate_size=14336, num_hidden_layers=32,
                         num_attention_heads=32, num_key_value_heads=8, num_local_experts=8, num_experts_per_tok=2,
                         sliding_window=None),
}


def _clock_sampler(stop: threading.Event, out: list) -> None:
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
        p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        return
    try:
        while not stop.is_set():
            line = p.stdout.readline()
            if not line:
                break
            out.append(line.strip())
    finally:
        p.terminate()


def _summarise_clocks(lines: list, local_rank: int) -> dict:
    """Median SM clock of the samples taken UNDER LOAD (power above idle) on this rank's GPU."""
    rows, mx, reasons = [], 0, set()
    for ln in lines:
        f = [x.strip() for x in ln.split(",")]
        if len(f) < 8 or not f[0].isdigit() or int(f[0]) != local_rank:
            continue
        try:
            rows.append((float(f[1]), float(f[3])))
            mx = max(mx, float(f[2]))
        except ValueError:
            continue
        for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
            if v.lower().startswith("active"):
                reasons.add(name)
    if rows:
        pmax = max(p for _, p in rows)
        load = [c for c, p in rows if p >= 0.6 * pmax] or [c for c, _ in rows]
    else:
        load = []
    return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
            "samples": len(rows)}


def _prompts(B: int, S_in: int, vocab: int) -> list:
    import torch
    g = torch.Generator().manual_seed(1234)
    return [{"prompt_token_ids": torch.randint(10, vocab - 10, (S_in,), generator=g).tolist()} for _ in range(B)]


def _timed_generates(gen, K: int, reps: int, sync=None):
    """min wall time of generate(max_tokens=1) and generate(max_tokens=K+1), interleaved; returns (t1, tK, all)."""
    t1s, tks = [], []
    for _ in range(reps):
        for n, acc in ((1, t1s), (K + 1, tks)):
            if sync:
                sync()
            # same GPU state at the start of both kinds of call: without the pause a generate(K+1) always starts right
            # after a 64 x 1024-token prefill burst (power-capped clocks) and a generate(1) right after light decode steps
            time.sleep(1.0)
            t0 = time.perf_counter()
            gen(n)
            if sync:
                sync()
            acc.append(time.perf_counter() - t0)
    return min(t1s), min(tks), {"generate_1_s": t1s, "generate_K1_s": tks}


def _prefill_tok_s(t1: float, B: int, S_in: int) -> float:
    """Second metric: prompt tokens per second of the whole call LLM.generate(max_tokens=1) — one-step prefill of B x S_in
    tokens through the public API, tokenised prompts in, first token out (includes scheduling, H2D, sampling, D2H)."""
    return B * S_in / t1


# ===================================================================================== reference arm
def _reference_unavailable(why: str) -> None:
    if int(os.environ.get("RANK", 0)) == 0:
        print(json.dumps({"impl": "reference", "unavailable": why}), flush=True)


def _sentinel_path() -> str:
    return os.path.join("/tmp", f"vb_ref_done_{os.environ.get('MASTER_PORT', '0')}_{os.environ.get('TORCHELASTIC_RUN_ID', 'x')}")


def run_reference(args) -> None:
    install_note = ("pip install of /root/reference into baseline/_ref is not possible offline: the metadata step needs "
                    "setuptools_rust (no Rust toolchain) and it pins torch==2.13 / flashinfer 0.6.17 vs torch 2.11 in "
                    "this image; see DESIGN.md")
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    sentinel = _sentinel_path()
    timeout_s = int(os.environ.get("VLLM_B200_REFERENCE_TIMEOUT", "820"))
    if rank != 0:
        # the reference spawns its own worker per GPU from rank 0; the other torchrun ranks must not hold a context
        t_end = time.time() + timeout_s + 30
        while not os.path.exists(sentinel) and time.time() < t_end:
            time.sleep(0.5)
        return
    try:
        os.unlink(sentinel)
    except OSError:
        pass
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != ROOT]      # nothing of ours importable from here on
    source = None
    if os.path.isdir(os.path.join(ref_dir, "vllm")):
        sys.path.insert(0, ref_dir)
        source = "baseline/_ref"
    if os.environ.get("VLLM_B200_REFERENCE_FROM_IMAGE", "1") == "0" and source is None:
        _reference_unavailable(install_note)
        open(sentinel, "w").close()
        return

    def finish(code: int = 0):
        try:
            open(sentinel, "w").close()
        finally:
            os._exit(code)

    def watchdog():
        time.sleep(timeout_s)
        _reference_unavailable(f"reference arm did not finish within {timeout_s}s ({install_note})")
        finish()
    threading.Thread(target=watchdog, daemon=True).start()
    try:
        # a clean, default environment for the stock engine: drop torchrun's rendezvous variables (the reference
        # would read RANK / LOCAL_RANK / WORLD_SIZE as its own) and do not override any of its engine defaults
        for k in list(os.environ):
            if k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "GROUP_RANK", "ROLE_RANK", "ROLE_NAME",
                     "ROLE_WORLD_SIZE", "GROUP_WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "OMP_NUM_THREADS") \
                    or k.startswith(("TORCHELASTIC_", "TORCH_NCCL_ASYNC", "VLLM_B200_")):
                os.environ.pop(k, None)
        os.environ.setdefault("HF_HUB_OFFLINE", "1")
        os.environ.setdefault("VLLM_LOGGING_LEVEL", "WARNING")
        os.environ.setdefault("VLLM_PLUGINS", "")                          # stock code path: no third-party plug-ins
        import tempfile

        import vllm
        from vllm import LLM, SamplingParams
        if source is None:
            source = f"vllm {vllm.__version__} wheel of the image ({install_note})"
        if args.model not in REF_CONFIGS:
            _reference_unavailable(f"reference arm has no config for {args.model}")
            finish()
        cfg = REF_CONFIGS[args.model]
        d = tempfile.mkdtemp(prefix="ref_cfg_")
        with open(os.path.join(d, "config.json"), "w") as f:
            json.dump(cfg, f)
        B, S_in, K, W = args.batch, args.input_len, args.steps, max(args.warmup, 3)
        llm = LLM(model=d, load_format="dummy", skip_tokenizer_init=True, dtype="bfloat16",
                  tensor_parallel_size=args.gpus, max_model_len=S_in + W + K + 64, max_num_seqs=max(B, 8),
                  max_num_batched_tokens=B * S_in, enable_prefix_caching=False, gpu_memory_utilization=0.85,
                  generation_config="vllm", enforce_eager=args.eager, disable_log_stats=False)
        prompts = _prompts(B, S_in, cfg["vocab_size"])
        last = {}

        def gen(n):
            sp = SamplingParams(max_tokens=n, temperature=0.0, ignore_eos=True, detokenize=False)
            last["outs"] = llm.generate(prompts, sp, use_tqdm=False)
        gen(W + K + 1)                           # warm-up: one prefill + (W + K) decode steps, graphs / JIT all hot
        gen(1)
        stop, lines = threading.Event(), []
        th = threading.Thread(target=_clock_sampler, args=(stop, lines), daemon=True)
        th.start()
        time.sleep(0.2)
        t1, tk, raw = _timed_generates(gen, K, REPS)
        stop.set()
        th.join(timeout=2)
        # the reference's own engine-core timestamps of the last K+1 run (first → last token of the batch)
        eng_ms = None
        try:
            st = [o.metrics for o in last["outs"] if getattr(o, "metrics", None) is not None]
            if st:
                eng_ms = (max(s.last_token_ts for s in st) - min(s.first_token_ts for s in st)) * 1e3 / K
        except Exception:
            eng_ms = None
        dec_s = tk - t1
        value = B * K / dec_s
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": dec_s * 1e3 / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": DATA,
            "config": {"model": args.model, "global_batch": B, "seq_len": S_in, "parallelism": f"tp{args.gpus}",
                       "cuda_graphs": not args.eager, "engine": "default (multiprocess engine core, async scheduling)",
                       "l2": "no flush needed: every step streams far more weights+KV than the 126 MB L2"},
            "clocks": _summarise_clocks(lines, 0),
            "e2e": {"value": value, "unit": "tokens/s", "ms_per_step": dec_s * 1e3 / K,
                    "method": "min wall(LLM.generate(K+1)) - min wall(LLM.generate(1)), %d reps" % REPS, **raw,
                    "engine_core_itl_ms": eng_ms, "prefill_tokens_per_s": _prefill_tok_s(t1, B, S_in)},
            "impl": "reference", "reference_source": source}), flush=True)
        try:
            del llm
        except Exception:
            pass
    except BaseException as e:  # noqa: BLE001 - whatever goes wrong, the arm reports instead of crashing the driver
        _reference_unavailable(f"reference arm failed: {type(e).__name__}: {str(e)[:300]} ({install_note})")
    finish()          # stock vLLM leaves non-daemon threads / process groups behind; the line above is the result


# ===================================================================================== our arm
def _exposed_comm(engine, runner, args, T: int) -> dict | None:
    """Per-call device time of the fused RowParallel kernel (GEMM + NVSwitch all-reduce + residual + RMSNorm) vs the
    same GEMM + a local norm without any peer traffic, for the two RowParallel shapes of the model → exposed
    communication per decode step and NVLink bytes.  Runs after the timed region."""
    try:
        from vllm_b200.benchmarks.comm_probe import probe_row_parallel
        return probe_row_parallel(runner, T)
    except Exception as e:  # noqa: BLE001 - diagnostic only
        return {"error": f"{type(e).__name__}: {e}"[:200]}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="llama-3-8b")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--input-len", type=int, default=1024)
    ap.add_argument("--eager", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sync", action="store_true", help="disable async scheduling (host waits for every step)")
    ap.add_argument("--parallelism", default="tp", choices=["tp", "ep"],
                    help="MoE models: ep = experts sharded over the TP group with the fused all-to-all kernels")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    assert world == args.gpus, f"launch with torchrun --nproc-per-node {args.gpus} (WORLD_SIZE={world})"
    torch.cuda.set_device(local_rank)

    from vllm_b200 import LLM, SamplingParams

    B, S_in, K, W = args.batch, args.input_len, args.steps, args.warmup
    SETTLE = 24                      # untimed decode steps between the prefill and the W warm-up steps (see phase A)
    out_len = W + K + 2 + SETTLE
    max_len = S_in + 2 * out_len + 64
    extra = {}
    if args.parallelism == "ep":
        extra["enable_expert_parallel"] = True
    llm = LLM(args.model, load_format="dummy", dtype="bfloat16", tensor_parallel_size=args.gpus,
              max_model_len=max_len, max_num_seqs=max(B, 8), max_num_batched_tokens=B * S_in,
              enable_prefix_caching=False, enforce_eager=args.eager,
              distributed_executor_backend="external_launcher" if world > 1 else "uni",
              cudagraph_capture_sizes=[B], disable_log_stats=True, gpu_memory_utilization=0.85,
              async_scheduling=not args.sync, multiprocess_mode=False, **extra)
    engine = llm.llm_engine
    runner = engine.engine_core.engine_core.model_executor.worker.runner
    vocab = engine.model_config.vocab_size
    prompts = _prompts(B, S_in, vocab)

    def add_batch(tag: str) -> None:
        sp = SamplingParams(max_tokens=out_len, temperature=0.0, ignore_eos=True, detokenize=False)
        for i, p in enumerate(prompts):
            engine.add_request(f"{tag}-{i}", p, sp)

    def barrier() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def prefill_until_decoding() -> None:
        seen: dict[str, int] = {}
        while len(seen) < B:
            for o in engine.step():
                seen[o.request_id] = len(o.outputs[0].token_ids)

    # ---------------- phase A: device-timed decode steps
    stop, lines = threading.Event(), []
    th = threading.Thread(target=_clock_sampler, args=(stop, lines), daemon=True)
    th.start()
    time.sleep(0.3)
    add_batch("a")
    prefill_until_decoding()
    # The one-step 64 x 1024-token prefill runs at the power cap and leaves the clocks capped for ~100 ms afterwards
    # (measured: ~15 decode steps at 6.7 instead of 5.4 ms on one GPU); an IDLE pause instead lets the clocks drop and the
    # timed steps pay the ramp (measured: +0.3 ms per step at TP4 / TP8).  So the GPU is kept busy: SETTLE untimed decode
    # steps (a fixed count: every rank must run the same steps, the collectives are inside them), then the W warm-up
    # steps, then — no gap — the K timed steps.
    for _ in range(SETTLE + W):
        engine.step()
    launches0 = runner.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(K):
        engine.step()
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = runner.launch_count - launches0
    while engine.has_unfinished_requests():
        engine.step()
    # device-only step time: K back-to-back replays of the captured decode graph with NO host work in between (the
    # buffers still hold the last step's inputs).  value vs this number = how much of a step is host-bound.
    gpu_only_ms = None
    g = getattr(runner, "graphs", {}).get(B)
    if g is not None:
        runner.finalize_pending()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            g.graph.replay()
        g0.record()
        for _ in range(K):
            g.graph.replay()
        g1.record()
        barrier()
        gpu_only_ms = g0.elapsed_time(g1) / K

    # ---------------- phase B: end to end through the public API, same method as the reference arm
    e2e, e2e_s, raw = None, 0.0, {}
    h2d_step = d2h_step = 0
    if not args.no_e2e:
        def gen(n):
            llm.generate(prompts, SamplingParams(max_tokens=n, temperature=0.0, ignore_eos=True, detokenize=False))
        gen(W + K + 1)
        gen(1)
        h2d0, d2h0 = runner.h2d_bytes, runner.d2h_bytes
        t1, tk, raw = _timed_generates(gen, K, REPS, sync=barrier)
        e2e_s = tk - t1
        # per-step wall clock of one more generate(K+1) (diagnostic: where the end-to-end time goes)
        stamps = []
        orig_step = engine.step

        def timed_step():
            r = orig_step()
            stamps.append((time.perf_counter(), runner.wait_s))
            return r
        engine.step = timed_step
        barrier()
        t_begin = (time.perf_counter(), runner.wait_s)
        gen(K + 1)
        engine.step = orig_step
        seq = [t_begin] + stamps
        raw["step_wall_ms"] = [round((b[0] - a[0]) * 1e3, 3) for a, b in zip(seq[:-1], seq[1:])]
        # host work per decode step = wall time of the step minus the time the host sat blocked on the device
        host = sorted((b[0] - a[0]) - (b[1] - a[1]) for a, b in zip(seq[3:-1], seq[4:]))
        raw["host_ms_per_step_p50"] = round(host[len(host) // 2] * 1e3, 3) if host else None
        n_steps = REPS * (K + 2)                          # engine steps executed by the timed generate calls
        h2d_step = (runner.h2d_bytes - h2d0) / n_steps
        d2h_step = (runner.d2h_bytes - d2h0) / n_steps
    stop.set()
    th.join(timeout=2)
    clocks = _summarise_clocks(lines, local_rank)
    comm = _exposed_comm(engine, runner, args, B) if world > 1 else None

    # ---------------- reduce over ranks (max time)
    t = torch.tensor([dev_ms, e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s = t.tolist()
    value = B * K / (dev_ms / 1e3)
    if not args.no_e2e:
        e2e = {"value": B * K / e2e_s, "unit": "tokens/s", "h2d_bytes_per_step": int(h2d_step),
               "d2h_bytes_per_step": int(d2h_step), "ms_per_step": e2e_s * 1e3 / K,
               "method": "min wall(LLM.generate(K+1)) - min wall(LLM.generate(1)), %d reps" % REPS, **raw,
               "prefill_tokens_per_s": _prefill_tok_s(t1, B, S_in)}
    if rank == 0:
        roof = None
        try:   # analytical roofline of the timed steps (weights + KV bytes, FLOPs) → achieved TB/s, MBU, MFU
            from vllm_b200.config import get_preset_hf_config
            from vllm_b200.metrics.perf import model_step_cost, utilisation
            hf = get_preset_hf_config(args.model)
            if hf is not None:
                ctx = S_in + SETTLE + W + K // 2
                roof = {k: round(v, 4) for k, v in utilisation(model_step_cost(hf, B, [ctx] * B), dev_ms / K / 1e3,
                                                                args.gpus).items()}
        except Exception:
            roof = None
        par = f"tp{args.gpus}" + ("+ep" if args.parallelism == "ep" else "")
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "itl_p50_ms": dev_ms / K, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": (value / BASELINE_TOK_S) if BASELINE_TOK_S else None, "dtype": "bf16", "data": DATA,
            "config": {"model": args.model, "global_batch": B, "seq_len": S_in, "parallelism": par,
                       "cuda_graphs": not args.eager, "engine": "in-process engine core, async scheduling",
                       "l2": "no flush needed: every step streams far more weights+KV than the 126 MB L2"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof,
            "gpu_only_ms_per_step": gpu_only_ms,
            "fused_comm": comm, "impl": "ours"}), flush=True)
    llm.shutdown() if hasattr(llm, "shutdown") else engine.shutdown()
    if world > 1 and dist.is_initialized():
        dist.barrier()


if __name__ == "__main__":
    main()
