"""``vllm_b200`` command line (reference: vllm/entrypoints/cli/*: serve, bench, chat, complete, run-batch,
collect-env).  ``python -m vllm_b200 <cmd>`` or the ``vllm_b200`` console script."""
from __future__ import annotations

import argparse
import json
import sys

from vllm_b200.engine.arg_utils import EngineArgs


def _cmd_serve(argv: list[str]) -> None:
    from vllm_b200.entrypoints.openai.api_server import main
    main(argv)


def _bench_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser("vllm_b200 bench")
    sub = ap.add_subparsers(dest="bench", required=True)
    common = dict(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    lat = sub.add_parser("latency", **common)
    lat.add_argument("--input-len", type=int, default=32)
    lat.add_argument("--output-len", type=int, default=128)
    lat.add_argument("--batch-size", type=int, default=8)
    lat.add_argument("--n", type=int, default=1)
    lat.add_argument("--num-iters-warmup", type=int, default=3)
    lat.add_argument("--num-iters", type=int, default=10)
    lat.add_argument("--output-json", default=None)
    EngineArgs.add_cli_args(lat)
    thr = sub.add_parser("throughput", **common)
    for p in (thr,):
        p.add_argument("--dataset-name", default="random", choices=["random", "sharegpt"])
        p.add_argument("--dataset-path", default=None)
        p.add_argument("--num-prompts", type=int, default=200)
        p.add_argument("--input-len", type=int, default=512)
        p.add_argument("--output-len", type=int, default=128)
        p.add_argument("--range-ratio", type=float, default=0.0)
        p.add_argument("--prefix-len", type=int, default=0)
        p.add_argument("--output-json", default=None)
    EngineArgs.add_cli_args(thr)
    srv = sub.add_parser("serve", **common)
    srv.add_argument("--base-url", default="http://127.0.0.1:8000")
    srv.add_argument("--model", required=True)
    srv.add_argument("--api-key", default=None)
    srv.add_argument("--dataset-name", default="random", choices=["random"])
    srv.add_argument("--num-prompts", type=int, default=200)
    srv.add_argument("--input-len", type=int, default=512)
    srv.add_argument("--output-len", type=int, default=128)
    srv.add_argument("--range-ratio", type=float, default=0.0)
    srv.add_argument("--prefix-len", type=int, default=0)
    srv.add_argument("--vocab-size", type=int, default=32000)
    srv.add_argument("--request-rate", type=float, default=float("inf"))
    srv.add_argument("--burstiness", type=float, default=1.0)
    srv.add_argument("--max-concurrency", type=int, default=None)
    srv.add_argument("--seed", type=int, default=0)
    srv.add_argument("--native", action="store_true", help="drive the load with the C++ epoll client")
    srv.add_argument("--output-json", default=None)
    sw = sub.add_parser("sweep", **common)
    for act in srv._actions:                       # same client options as ``bench serve`` ...
        if act.option_strings and act.dest not in ("help", "request_rate", "max_concurrency"):
            sw._add_action(act)
    sw.add_argument("--request-rates", default="inf", help="comma-separated request rates (inf = closed loop)")
    sw.add_argument("--max-concurrencies", default="none", help="comma-separated concurrency caps")
    sw.add_argument("--slo-ttft-ms", type=float, default=None)
    sw.add_argument("--slo-tpot-ms", type=float, default=None)
    mm = sub.add_parser("mm_processor", **common)
    mm.add_argument("--model", required=True)
    mm.add_argument("--num-items", type=int, default=64)
    mm.add_argument("--image-size", default="448x448", help="HxW of the synthetic images")
    mm.add_argument("--audio-seconds", type=float, default=10.0)
    mm.add_argument("--unique", type=int, default=0, help="number of distinct items (0 = all distinct; fewer exercises the cache)")
    mm.add_argument("--hf-overrides", type=json.loads, default=None)
    mm.add_argument("--output-json", default=None)
    st = sub.add_parser("startup", **common)
    st.add_argument("--output-json", default=None)
    EngineArgs.add_cli_args(st)
    return ap


def _cmd_bench(argv: list[str]) -> None:
    from vllm_b200 import benchmarks as B
    args = _bench_parser().parse_args(argv)
    fn = {"latency": B.bench_latency, "throughput": B.bench_throughput, "serve": B.bench_serve,
          "startup": B.bench_startup, "sweep": B.bench_sweep, "mm_processor": B.bench_mm_processor}[args.bench]
    res = fn(args)
    print(json.dumps(res, indent=2))
    if getattr(args, "output_json", None):
        with open(args.output_json, "w") as f:
            json.dump(res, f, indent=2)


def _client(argv: list[str], chat: bool) -> None:
    """Tiny interactive client against a running server (reference: cli/openai.py chat/complete)."""
    import requests
    ap = argparse.ArgumentParser("vllm_b200 chat" if chat else "vllm_b200 complete")
    ap.add_argument("--url", default="http://127.0.0.1:8000/v1")
    ap.add_argument("--model-name", default=None)
    ap.add_argument("--api-key", default=None)
    ap.add_argument("--max-tokens", type=int, default=256)
    ap.add_argument("-q", "--quick", default=None, help="send one message and exit")
    a = ap.parse_args(argv)
    hdr = {"Authorization": f"Bearer {a.api_key}"} if a.api_key else {}
    model = a.model_name or requests.get(f"{a.url}/models", headers=hdr, timeout=30).json()["data"][0]["id"]
    history: list[dict] = []
    while True:
        text = a.quick if a.quick is not None else input("> ")
        if chat:
            history.append({"role": "user", "content": text})
            r = requests.post(f"{a.url}/chat/completions", headers=hdr, timeout=600,
                              json={"model": model, "messages": history, "max_tokens": a.max_tokens}).json()
            out = r["choices"][0]["message"]["content"]
            history.append({"role": "assistant", "content": out})
        else:
            r = requests.post(f"{a.url}/completions", headers=hdr, timeout=600,
                              json={"model": model, "prompt": text, "max_tokens": a.max_tokens}).json()
            out = r["choices"][0]["text"]
        print(out)
        if a.quick is not None:
            return


def _cmd_run_batch(argv: list[str]) -> None:
    """OpenAI batch-file runner: JSONL of {custom_id, method, url, body} → JSONL of responses
    (reference: entrypoints/openai/run_batch.py)."""
    import asyncio
    ap = argparse.ArgumentParser("vllm_b200 run-batch")
    ap.add_argument("-i", "--input-file", required=True)
    ap.add_argument("-o", "--output-file", required=True)
    EngineArgs.add_cli_args(ap)
    args = ap.parse_args(argv)

    async def run() -> None:
        import httpx
        from vllm_b200.engine.async_llm import AsyncLLM
        from vllm_b200.entrypoints.openai.api_server import ServingContext, build_app
        ea = EngineArgs.from_cli_args(args)
        engine = AsyncLLM.from_engine_args(ea)
        app = build_app(ServingContext(engine, ea.served_model_name or ea.model, engine.model_config.max_model_len,
                                       engine.tokenizer))
        out = []
        async with httpx.AsyncClient(transport=httpx.ASGITransport(app=app), base_url="http://batch") as client:
            async def one(line: str):
                req = json.loads(line)
                r = await client.post(req["url"], json=req["body"], timeout=None)
                return {"id": f"batch-{req['custom_id']}", "custom_id": req["custom_id"],
                        "response": {"status_code": r.status_code, "body": r.json()},
                        "error": None if r.status_code == 200 else r.text}
            with open(args.input_file) as f:
                lines = [ln for ln in f if ln.strip()]
            out = await asyncio.gather(*[one(ln) for ln in lines])
        with open(args.output_file, "w") as f:
            for o in out:
                f.write(json.dumps(o) + "\n")
        engine.shutdown()
    asyncio.run(run())


def _cmd_collect_env(argv: list[str]) -> None:
    from vllm_b200.platforms import collect_env
    print(collect_env())


def _cmd_launch(argv: list[str]) -> None:
    """``launch <component>``: individual serving components (reference: cli/launch.py)."""
    if not argv or argv[0] in ("-h", "--help"):
        print("usage: vllm_b200 launch {render} ...")
        return
    comp = argv.pop(0)
    if comp != "render":
        raise SystemExit(f"unknown component {comp!r}; expected one of ['render']")
    from vllm_b200.entrypoints.render_server import main as render_main
    render_main(argv)


COMMANDS = {
    "launch": _cmd_launch,
    "serve": _cmd_serve, "bench": _cmd_bench, "chat": lambda a: _client(a, True),
    "complete": lambda a: _client(a, False), "run-batch": _cmd_run_batch, "collect-env": _cmd_collect_env,
}


def expand_config_file(argv: list[str]) -> list[str]:
    """``--config file.yaml``: the file's keys become options placed BEFORE the explicit command-line options, so the
    command line wins (reference: utils/argparse_utils.py FlexibleArgumentParser).  Booleans → ``--flag`` /
    ``--no-flag``; lists → comma joined; dicts → JSON; ``model`` / ``model_tag`` → the positional model."""
    if "--config" not in argv:
        return argv
    import yaml
    i = argv.index("--config")
    path = argv[i + 1]
    rest = argv[:i] + argv[i + 2:]
    with open(path) as f:
        cfg = yaml.safe_load(f) or {}
    opts: list[str] = []
    positional: list[str] = []
    for k, v in cfg.items():
        flag = "--" + str(k).replace("_", "-")
        if k in ("model", "model_tag") and rest and rest[0] == "serve":
            positional.append(str(v))
        elif isinstance(v, bool):
            opts.append(flag if v else "--no-" + flag[2:])
        elif isinstance(v, (list, tuple)):
            opts += [flag, ",".join(str(x) for x in v)]
        elif isinstance(v, dict):
            opts += [flag, json.dumps(v)]
        elif v is not None:
            opts += [flag, str(v)]
    if not rest:
        return opts
    # keep the sub-command first; a positional model given on the command line overrides the file's
    head, tail = rest[:1], rest[1:]
    if positional and (not tail or tail[0].startswith("-")):
        head += positional
    return head + opts + tail


def main(argv=None) -> None:
    argv = expand_config_file(list(sys.argv[1:] if argv is None else argv))
    if not argv or argv[0] in ("-h", "--help"):
        print("usage: vllm_b200 {" + ",".join(COMMANDS) + "} ...")
        return
    if argv[0] in ("-v", "--version"):
        from vllm_b200.version import __version__
        print(__version__)
        return
    cmd = argv.pop(0)
    if cmd not in COMMANDS:
        raise SystemExit(f"unknown command {cmd!r}; expected one of {sorted(COMMANDS)}")
    COMMANDS[cmd](argv)


if __name__ == "__main__":
    main()
