"""Benchmark library behind ``vllm_b200 bench {latency,throughput,serve,startup}``.

Behavioural reference: vllm/benchmarks/{latency,throughput,serve,startup}.py and datasets.py (random /
sharegpt-style / sonnet datasets, request-rate Poisson arrivals, TTFT / TPOT / ITL / E2EL percentiles).
Everything is offline-safe: the default dataset is synthetic random token ids of a given shape.
"""
from __future__ import annotations

import asyncio
import json
import random
import time
from dataclasses import dataclass, field
from typing import Any, Optional

import numpy as np


# ------------------------------------------------------------------ datasets
@dataclass
class SampleRequest:
    prompt_token_ids: list[int]
    prompt_len: int
    expected_output_len: int
    prompt: Optional[str] = None


def random_dataset(num_requests: int, input_len: int, output_len: int, vocab_size: int, seed: int = 0,
                   range_ratio: float = 0.0, prefix_len: int = 0) -> list[SampleRequest]:
    """Uniform random token ids; lengths jittered by ``range_ratio``; optional shared prefix (prefix-cache tests)."""
    rng = np.random.default_rng(seed)
    prefix = rng.integers(5, vocab_size, prefix_len).tolist() if prefix_len else []
    out = []
    for _ in range(num_requests):
        il = int(round(input_len * (1 + rng.uniform(-range_ratio, range_ratio)))) if range_ratio else input_len
        ol = int(round(output_len * (1 + rng.uniform(-range_ratio, range_ratio)))) if range_ratio else output_len
        ids = prefix + rng.integers(5, vocab_size, max(il - prefix_len, 1)).tolist()
        out.append(SampleRequest(ids, len(ids), max(ol, 1)))
    return out


def sharegpt_dataset(path: str, tokenizer, num_requests: int, seed: int = 0, max_len: int = 4096) -> list[SampleRequest]:
    """ShareGPT-format JSON: first human turn = prompt, first reply length = expected output length."""
    with open(path) as f:
        data = json.load(f)
    rnd = random.Random(seed)
    rnd.shuffle(data)
    out = []
    for conv in data:
        turns = conv.get("conversations") or []
        if len(turns) < 2:
            continue
        p_ids = tokenizer.encode(turns[0]["value"])
        o_len = len(tokenizer.encode(turns[1]["value"]))
        if 4 <= len(p_ids) <= max_len and 4 <= o_len and len(p_ids) + o_len <= max_len:
            out.append(SampleRequest(p_ids, len(p_ids), o_len, turns[0]["value"]))
        if len(out) >= num_requests:
            break
    return out


def get_dataset(args, vocab_size: int, tokenizer=None) -> list[SampleRequest]:
    name = getattr(args, "dataset_name", "random")
    if name == "random":
        return random_dataset(args.num_prompts, args.input_len, args.output_len, vocab_size, args.seed,
                              getattr(args, "range_ratio", 0.0), getattr(args, "prefix_len", 0))
    if name == "sharegpt":
        return sharegpt_dataset(args.dataset_path, tokenizer, args.num_prompts, args.seed)
    raise ValueError(f"unknown dataset {name!r} (random|sharegpt)")


# ------------------------------------------------------------------ metrics
def percentiles(xs: list[float], ps=(50, 90, 99)) -> dict[str, float]:
    if not xs:
        return {f"p{p}": float("nan") for p in ps} | {"mean": float("nan")}
    a = np.asarray(xs, dtype=np.float64)
    return {"mean": float(a.mean())} | {f"p{p}": float(np.percentile(a, p)) for p in ps}


@dataclass
class RequestTiming:
    start: float = 0.0
    first_token: float = 0.0
    end: float = 0.0
    token_times: list[float] = field(default_factory=list)
    prompt_len: int = 0
    output_len: int = 0
    ok: bool = True


def summarize(timings: list[RequestTiming], wall: float) -> dict[str, Any]:
    good = [t for t in timings if t.ok and t.output_len > 0]
    ttft = [(t.first_token - t.start) * 1e3 for t in good]
    e2el = [(t.end - t.start) * 1e3 for t in good]
    tpot = [(t.end - t.first_token) * 1e3 / (t.output_len - 1) for t in good if t.output_len > 1]
    itl = [(b - a) * 1e3 for t in good for a, b in zip(t.token_times, t.token_times[1:])]
    out_toks = sum(t.output_len for t in good)
    in_toks = sum(t.prompt_len for t in good)
    return {
        "completed": len(good), "failed": len(timings) - len(good), "duration_s": wall,
        "request_throughput": len(good) / wall if wall > 0 else 0.0,
        "output_throughput": out_toks / wall if wall > 0 else 0.0,
        "total_token_throughput": (in_toks + out_toks) / wall if wall > 0 else 0.0,
        "ttft_ms": percentiles(ttft), "tpot_ms": percentiles(tpot), "itl_ms": percentiles(itl),
        "e2el_ms": percentiles(e2el),
    }


# ------------------------------------------------------------------ bench latency
def bench_latency(args) -> dict[str, Any]:
    """Fixed batch, fixed lengths, end-to-end ``LLM.generate`` latency (reference: benchmarks/latency.py)."""
    from vllm_b200 import LLM, SamplingParams
    from vllm_b200.engine.arg_utils import EngineArgs
    llm = LLM(**EngineArgs.from_cli_args(args).to_llm_kwargs())
    vocab = llm.llm_engine.model_config.hf_config["vocab_size"]
    reqs = random_dataset(args.batch_size, args.input_len, args.output_len, vocab, args.seed)
    prompts = [{"prompt_token_ids": r.prompt_token_ids} for r in reqs]
    sp = SamplingParams(max_tokens=args.output_len, temperature=1.0, top_p=1.0, ignore_eos=True, n=args.n)
    for _ in range(args.num_iters_warmup):
        llm.generate(prompts, sp, use_tqdm=False)
    lat = []
    for _ in range(args.num_iters):
        t0 = time.perf_counter()
        llm.generate(prompts, sp, use_tqdm=False)
        lat.append(time.perf_counter() - t0)
    llm.shutdown()
    res = {"avg_latency_s": float(np.mean(lat)), "latencies_s": lat,
           "percentiles_s": {str(p): float(np.percentile(lat, p)) for p in (10, 25, 50, 75, 90, 99)},
           "batch_size": args.batch_size, "input_len": args.input_len, "output_len": args.output_len}
    return res


# ------------------------------------------------------------------ bench throughput
def bench_throughput(args) -> dict[str, Any]:
    """All prompts submitted at once (reference: benchmarks/throughput.py)."""
    from vllm_b200 import LLM, SamplingParams
    from vllm_b200.engine.arg_utils import EngineArgs
    llm = LLM(**EngineArgs.from_cli_args(args).to_llm_kwargs())
    vocab = llm.llm_engine.model_config.hf_config["vocab_size"]
    reqs = get_dataset(args, vocab, llm.get_tokenizer())
    prompts = [{"prompt_token_ids": r.prompt_token_ids} for r in reqs]
    sps = [SamplingParams(max_tokens=r.expected_output_len, temperature=1.0, ignore_eos=True) for r in reqs]
    t0 = time.perf_counter()
    outs = llm.generate(prompts, sps, use_tqdm=False)
    dt = time.perf_counter() - t0
    llm.shutdown()
    n_in = sum(r.prompt_len for r in reqs)
    n_out = sum(len(o.outputs[0].token_ids) for o in outs)
    return {"elapsed_s": dt, "num_requests": len(reqs), "requests_per_s": len(reqs) / dt,
            "total_tokens_per_s": (n_in + n_out) / dt, "output_tokens_per_s": n_out / dt,
            "total_input_tokens": n_in, "total_output_tokens": n_out}


# ------------------------------------------------------------------ bench serve
async def _one_request(session, url: str, model: str, r: SampleRequest, api_key: Optional[str]) -> RequestTiming:
    t = RequestTiming(prompt_len=r.prompt_len)
    payload = {"model": model, "prompt": r.prompt_token_ids, "max_tokens": r.expected_output_len, "temperature": 0.0,
               "ignore_eos": True, "stream": True, "stream_options": {"include_usage": True}}
    headers = {"Authorization": f"Bearer {api_key}"} if api_key else {}
    t.start = time.perf_counter()
    try:
        async with session.post(url, json=payload, headers=headers) as resp:
            if resp.status != 200:
                t.ok = False
                return t
            async for raw in resp.content:
                line = raw.decode().strip()
                if not line.startswith("data:"):
                    continue
                body = line[5:].strip()
                if body == "[DONE]":
                    break
                ev = json.loads(body)
                now = time.perf_counter()
                if ev.get("choices") and (ev["choices"][0].get("text") or ev["choices"][0].get("token_ids")):
                    if not t.token_times:
                        t.first_token = now
                    t.token_times.append(now)
                if ev.get("usage"):
                    t.output_len = ev["usage"].get("completion_tokens", t.output_len)
    except Exception:
        t.ok = False
    t.end = time.perf_counter()
    if t.output_len == 0:
        t.output_len = len(t.token_times)
    return t


async def bench_serve_async(args) -> dict[str, Any]:
    """Open-loop load against a running OpenAI-compatible server (reference: benchmarks/serve.py): Poisson (or
    gamma-burstiness) arrivals at ``request_rate``, optional ``max_concurrency`` semaphore."""
    import aiohttp
    vocab = args.vocab_size
    reqs = get_dataset(args, vocab)
    url = f"{args.base_url.rstrip('/')}/v1/completions"
    rng = np.random.default_rng(args.seed)
    sem = asyncio.Semaphore(args.max_concurrency) if args.max_concurrency else None
    conn = aiohttp.TCPConnector(limit=0)
    async with aiohttp.ClientSession(connector=conn, timeout=aiohttp.ClientTimeout(total=6 * 3600)) as session:
        async def guarded(r):
            if sem is None:
                return await _one_request(session, url, args.model, r, args.api_key)
            async with sem:
                return await _one_request(session, url, args.model, r, args.api_key)
        tasks = []
        t0 = time.perf_counter()
        for r in reqs:
            tasks.append(asyncio.create_task(guarded(r)))
            if args.request_rate != float("inf"):
                theta = 1.0 / (args.request_rate * args.burstiness)
                await asyncio.sleep(float(rng.gamma(args.burstiness, theta)))
        timings = await asyncio.gather(*tasks)
        wall = time.perf_counter() - t0
    return summarize(list(timings), wall)


def bench_serve_native(args) -> dict[str, Any]:
    """Same measurement with the C++ epoll load generator (``native/tools/loadgen.cpp``): no Python in the timing
    path, so client-side overhead does not cap throughput at high request rates."""
    import subprocess
    from urllib.parse import urlparse
    from vllm_b200.native import build_loadgen
    u = urlparse(args.base_url)
    cmd = [build_loadgen(), "--host", u.hostname or "127.0.0.1", "--port", str(u.port or 80), "--model", args.model,
           "--num", str(args.num_prompts), "--input-len", str(args.input_len), "--output-len", str(args.output_len),
           "--vocab", str(args.vocab_size), "--seed", str(args.seed),
           "--rate", "0" if args.request_rate == float("inf") else str(args.request_rate),
           "--concurrency", str(args.max_concurrency or 0)]
    if args.api_key:
        cmd += ["--api-key", args.api_key]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if not r.stdout.strip():
        raise RuntimeError(f"loadgen failed: {r.stderr}")
    return json.loads(r.stdout)


def bench_serve(args) -> dict[str, Any]:
    if getattr(args, "native", False):
        return bench_serve_native(args)
    return asyncio.run(bench_serve_async(args))


# ------------------------------------------------------------------ bench mm_processor
def bench_mm_processor(args) -> dict[str, Any]:
    """Front-end multimodal preprocessing throughput with no engine: synthetic images (or audio clips) through the model's
    processor — resize / tiling / patching / log-mel, placeholder expansion, content hashing (reference:
    benchmarks/mm_processor.py)."""
    import numpy as np
    from vllm_b200.config.model import ModelConfig
    from vllm_b200.multimodal import MULTIMODAL_REGISTRY
    mc = ModelConfig(model=args.model, hf_overrides=getattr(args, "hf_overrides", None) or {})
    if not MULTIMODAL_REGISTRY.supports_multimodal(mc):
        raise SystemExit(f"{args.model} has no multimodal processor")
    proc = MULTIMODAL_REGISTRY.create_processor(mc)
    cfg = mc.hf_config
    audio = cfg.get("audio_config") is not None and cfg.get("vision_config") is None
    tok = cfg.get("audio_token_index", cfg.get("audio_token_id")) if audio else cfg.get("image_token_index", cfg.get("image_token_id"))
    rng = np.random.default_rng(0)
    n, uniq = args.num_items, (args.unique or args.num_items)
    if audio:
        items = [rng.standard_normal(int(16000 * args.audio_seconds)).astype(np.float32) * 0.1 for _ in range(uniq)]
        key = "audio"
    else:
        h, w = (int(x) for x in args.image_size.lower().split("x"))
        items = [rng.integers(0, 255, (h, w, 3), dtype=np.uint8) for _ in range(uniq)]
        key = "image"
    lat, toks = [], 0
    t0 = time.perf_counter()
    for i in range(n):
        t = time.perf_counter()
        ids, feats = proc.process([1, tok, 2], {key: items[i % uniq]})
        lat.append(time.perf_counter() - t)
        toks += len(ids) - 2
    total = time.perf_counter() - t0
    lat.sort()
    return {"model": args.model, "modality": key, "items": n, "unique_items": uniq, "items_per_s": n / total,
            "placeholder_tokens_per_item": toks / n, "latency_ms": {"mean": 1e3 * sum(lat) / n, "p50": 1e3 * lat[n // 2],
                                                                  "p99": 1e3 * lat[min(n - 1, int(n * 0.99))]}}


# ------------------------------------------------------------------ bench startup
def bench_startup(args) -> dict[str, Any]:
    """Cold engine start → first token (reference: benchmarks/startup.py)."""
    from vllm_b200 import LLM, SamplingParams
    from vllm_b200.engine.arg_utils import EngineArgs
    t0 = time.perf_counter()
    llm = LLM(**EngineArgs.from_cli_args(args).to_llm_kwargs())
    t1 = time.perf_counter()
    llm.generate([{"prompt_token_ids": [5, 6, 7, 8]}], SamplingParams(max_tokens=1), use_tqdm=False)
    t2 = time.perf_counter()
    llm.shutdown()
    return {"engine_init_s": t1 - t0, "first_token_s": t2 - t1, "total_s": t2 - t0}


# ------------------------------------------------------------------ bench sweep
def bench_sweep(args) -> dict[str, Any]:
    """Grid of ``bench serve`` runs against one running server (reference: benchmarks/sweep/: serve x parameter grid
    → table + the knee of the latency / throughput curve).  ``--request-rates`` and/or ``--max-concurrencies`` span
    the grid; every cell reports throughput and TTFT / TPOT / ITL percentiles, the summary names the highest-throughput
    cell that still meets the SLOs (``--slo-ttft-ms``, ``--slo-tpot-ms``)."""
    import copy
    rates = [float(x) for x in (args.request_rates or "inf").split(",")]
    concs = [int(x) if x not in ("", "none", "0") else None for x in (args.max_concurrencies or "none").split(",")]
    cells = []
    for rate in rates:
        for conc in concs:
            a = copy.copy(args)
            a.request_rate, a.max_concurrency = rate, conc
            res = bench_serve(a)
            cells.append({"request_rate": rate, "max_concurrency": conc, **res})
    def ok(c):
        return ((args.slo_ttft_ms is None or c["ttft_ms"]["p99"] <= args.slo_ttft_ms)
                and (args.slo_tpot_ms is None or c["tpot_ms"]["p99"] <= args.slo_tpot_ms))
    good = [c for c in cells if ok(c)]
    best = max(good, key=lambda c: c["output_throughput"]) if good else None
    return {"cells": cells, "best_within_slo": None if best is None else {
        "request_rate": best["request_rate"], "max_concurrency": best["max_concurrency"],
        "output_throughput": best["output_throughput"]}}
