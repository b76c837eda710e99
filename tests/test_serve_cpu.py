"""Full serving stack in a subprocess (reference technique: tests/utils.py ``RemoteOpenAIServer``): ``vllm_b200 serve``
→ uvicorn → AsyncLLM → ZMQ → EngineCore process → worker, exercised over real HTTP with the benchmark client."""
import argparse
import json
import os
import subprocess
import sys
import time

import pytest
import requests

from vllm_b200.utils import get_open_port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def server():
    port = get_open_port()
    env = dict(os.environ, PYTHONPATH=ROOT, VLLM_LOGGING_LEVEL="WARNING")
    cmd = [sys.executable, "-m", "vllm_b200", "serve", "llama-tiny", "--device", "cpu", "--dtype", "float32",
           "--max-model-len", "256", "--max-num-seqs", "8", "--max-num-batched-tokens", "128",
           "--kv-cache-memory-bytes", str(16 << 20), "--host", "127.0.0.1", "--port", str(port)]
    proc = subprocess.Popen(cmd, env=env, cwd=ROOT, stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT)
    url = f"http://127.0.0.1:{port}"
    deadline = time.time() + 120
    while time.time() < deadline:
        if proc.poll() is not None:
            pytest.fail("server exited during start-up")
        try:
            if requests.get(url + "/health", timeout=1).status_code == 200:
                break
        except requests.RequestException:
            time.sleep(0.5)
    else:
        proc.terminate()
        pytest.fail("server did not become healthy")
    yield url
    proc.terminate()
    try:
        proc.wait(timeout=20)
    except subprocess.TimeoutExpired:
        proc.kill()


def test_http_completions_chat_and_stream(server):
    r = requests.post(server + "/v1/completions", json={"model": "llama-tiny", "prompt": [5, 6, 7, 8], "max_tokens": 5,
                                                         "temperature": 0}, timeout=60).json()
    assert r["usage"]["completion_tokens"] == 5
    r = requests.post(server + "/v1/chat/completions", json={"model": "llama-tiny", "max_tokens": 4, "temperature": 0,
                                                              "messages": [{"role": "user", "content": "hi"}]},
                      timeout=60).json()
    assert r["choices"][0]["message"]["role"] == "assistant" and r["usage"]["completion_tokens"] == 4
    chunks = 0
    with requests.post(server + "/v1/completions", json={"model": "llama-tiny", "prompt": "abc", "max_tokens": 6,
                                                          "temperature": 0, "stream": True}, stream=True, timeout=60) as s:
        for line in s.iter_lines():
            if line.startswith(b"data:") and b"[DONE]" not in line:
                chunks += 1
    assert chunks >= 6
    m = requests.get(server + "/metrics", timeout=10).text
    assert "vllm:generation_tokens" in m or "generation_tokens" in m


def test_bench_serve_against_live_server(server):
    from vllm_b200.benchmarks import bench_serve
    args = argparse.Namespace(base_url=server, model="llama-tiny", api_key=None, dataset_name="random", num_prompts=12,
                              input_len=24, output_len=6, range_ratio=0.0, prefix_len=0, vocab_size=500,
                              request_rate=float("inf"), burstiness=1.0, max_concurrency=4, seed=0)
    res = bench_serve(args)
    assert res["completed"] == 12 and res["failed"] == 0
    assert res["output_throughput"] > 0 and res["ttft_ms"]["mean"] > 0
    json.dumps(res)


def test_bench_sweep_grid(server):
    from vllm_b200.benchmarks import bench_sweep
    from vllm_b200.entrypoints.cli import _bench_parser
    args = _bench_parser().parse_args(["sweep", "--base-url", server, "--model", "llama-tiny", "--num-prompts", "6",
                                       "--input-len", "16", "--output-len", "4", "--vocab-size", "500",
                                       "--max-concurrencies", "1,4", "--slo-ttft-ms", "60000"])
    res = bench_sweep(args)
    assert [c["max_concurrency"] for c in res["cells"]] == [1, 4] and all(c["completed"] == 6 for c in res["cells"])
    assert res["best_within_slo"]["max_concurrency"] in (1, 4)


def test_token_in_token_out_and_render(server):
    r = requests.post(server + "/inference/v1/generate", json={"token_ids": [5, 6, 7, 8], "sampling_params": {"max_tokens": 4, "temperature": 0}}, timeout=60).json()
    assert len(r["token_ids"]) == 4 and r["finish_reason"] == "length"
    ref = requests.post(server + "/v1/completions", json={"model": "llama-tiny", "prompt": [5, 6, 7, 8], "max_tokens": 4, "temperature": 0, "return_token_ids": True}, timeout=60).json()
    assert ref["usage"]["completion_tokens"] == 4
    rd = requests.post(server + "/render", json={"messages": [{"role": "user", "content": "hi"}]}, timeout=30).json()
    assert rd["count"] == len(rd["token_ids"]) > 0 and "hi" in rd["prompt"]
    back = requests.post(server + "/derender", json={"token_ids": rd["token_ids"]}, timeout=30).json()
    assert "hi" in back["text"]


def test_native_loadgen_against_live_server(server):
    from vllm_b200.benchmarks import bench_serve
    args = argparse.Namespace(base_url=server, model="llama-tiny", api_key=None, dataset_name="random", num_prompts=10,
                              input_len=20, output_len=5, range_ratio=0.0, prefix_len=0, vocab_size=500,
                              request_rate=50.0, burstiness=1.0, max_concurrency=4, seed=1, native=True)
    res = bench_serve(args)
    assert res["completed"] == 10 and res["failed"] == 0
    assert res["output_throughput"] > 0 and res["ttft_ms"]["mean"] > 0 and res["e2el_ms"]["p99"] >= res["ttft_ms"]["p50"]


def _spawn(cmd, env):
    return subprocess.Popen(cmd, env=env, cwd=ROOT, stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT)


def _wait_healthy(procs_urls, deadline_s=180):
    deadline = time.time() + deadline_s
    for proc, url in procs_urls:
        while True:
            assert proc.poll() is None, f"{url} exited during start-up"
            assert time.time() < deadline, f"{url} never became healthy"
            try:
                if requests.get(url + "/health", timeout=1).status_code == 200:
                    break
            except requests.RequestException:
                time.sleep(0.5)


def test_disaggregated_serving_through_proxy(server):
    """prefill server + decode server + proxy == one ordinary server, token for token (greedy)."""
    env = dict(os.environ, PYTHONPATH=ROOT, VLLM_LOGGING_LEVEL="WARNING")
    kv_port = get_open_port()
    base = [sys.executable, "-m", "vllm_b200", "serve", "llama-tiny", "--device", "cpu", "--dtype", "float32",
            "--max-model-len", "256", "--max-num-seqs", "8", "--max-num-batched-tokens", "128", "--block-size", "16",
            "--kv-cache-memory-bytes", str(16 << 20), "--host", "127.0.0.1"]
    ports = [get_open_port() for _ in range(3)]
    ktc = lambda role: json.dumps({"kv_connector": "P2pPullConnector", "kv_role": role, "kv_port": kv_port})  # noqa: E731
    procs = [_spawn(base + ["--port", str(ports[0]), "--kv-transfer-config", ktc("kv_producer")], env),
             _spawn(base + ["--port", str(ports[1]), "--kv-transfer-config", ktc("kv_consumer")], env)]
    urls = [f"http://127.0.0.1:{p}" for p in ports]
    try:
        _wait_healthy(list(zip(procs, urls[:2])))
        procs.append(_spawn([sys.executable, "-m", "vllm_b200.entrypoints.disagg_proxy", "--prefill", urls[0],
                             "--decode", urls[1], "--host", "127.0.0.1", "--port", str(ports[2])], env))
        _wait_healthy([(procs[2], urls[2])], 60)
        body = {"model": "llama-tiny", "prompt": [(13 * i) % 300 + 7 for i in range(70)], "max_tokens": 12,
                "temperature": 0}
        want = requests.post(server + "/v1/completions", json=body, timeout=120).json()
        got = requests.post(urls[2] + "/v1/completions", json=body, timeout=120).json()
        assert got["choices"][0]["text"] == want["choices"][0]["text"]
        assert got["usage"]["completion_tokens"] == 12
        text = ""
        with requests.post(urls[2] + "/v1/completions", json=dict(body, stream=True), stream=True, timeout=120) as s:
            for line in s.iter_lines():
                if line.startswith(b"data:") and b"[DONE]" not in line:
                    text += json.loads(line[5:])["choices"][0]["text"]
        assert text == want["choices"][0]["text"]
        st = requests.get(urls[2] + "/proxy/stats", timeout=10).json()
        assert st["prefill"][urls[0]] == 2 and st["decode"][urls[1]] == 2
    finally:
        for p in procs:
            p.terminate()
        for p in procs:
            try:
                p.wait(timeout=20)
            except subprocess.TimeoutExpired:
                p.kill()


def test_native_frontend_matches_python_server(server):
    """C++ front-end daemon (HTTP → ZMQ/msgpack → engine core) returns the same tokens as the Python API server,
    streams SSE, detokenises incrementally and honours stop strings."""
    env = dict(os.environ, PYTHONPATH=ROOT, VLLM_LOGGING_LEVEL="WARNING")
    port = get_open_port()
    proc = _spawn([sys.executable, "-m", "vllm_b200", "serve", "llama-tiny", "--device", "cpu", "--dtype", "float32",
                   "--max-model-len", "256", "--max-num-seqs", "8", "--max-num-batched-tokens", "128",
                   "--kv-cache-memory-bytes", str(16 << 20), "--host", "127.0.0.1", "--port", str(port),
                   "--native-frontend"], env)
    url = f"http://127.0.0.1:{port}"
    try:
        _wait_healthy([(proc, url)])
        assert requests.get(url + "/v1/models", timeout=10).json()["data"][0]["id"] == "llama-tiny"
        prompt = [(17 * i) % 200 + 5 for i in range(40)]
        body = {"model": "llama-tiny", "prompt": prompt, "max_tokens": 12, "temperature": 0}
        want = requests.post(server + "/v1/completions", json=body, timeout=120).json()
        got = requests.post(url + "/v1/completions", json=body, timeout=120).json()
        assert got["usage"] == {"prompt_tokens": 40, "completion_tokens": 12, "total_tokens": 52}
        assert got["choices"][0]["finish_reason"] == "length"
        ids = got["choices"][0]["token_ids"]
        # same tokens as the Python stack (via its token-in/token-out route), same text as its detokeniser
        ref = requests.post(server + "/inference/v1/generate", json={"token_ids": prompt, "sampling_params": {
            "max_tokens": 12, "temperature": 0}}, timeout=120).json()
        assert ids == ref["token_ids"]
        assert got["choices"][0]["text"] == want["choices"][0]["text"]
        tok = requests.post(url + "/inference/v1/generate", json={"token_ids": prompt, "sampling_params": {
            "max_tokens": 12, "temperature": 0}}, timeout=120).json()
        assert tok["token_ids"] == ids and tok["num_prompt_tokens"] == 40
        # streaming: concatenated deltas == full text, [DONE] terminator, usage chunk on request
        text, saw_done, usage = "", False, None
        with requests.post(url + "/v1/completions", json=dict(body, stream=True, stream_options={"include_usage": True}),
                           stream=True, timeout=120) as s:
            for line in s.iter_lines():
                if not line.startswith(b"data:"):
                    continue
                if b"[DONE]" in line:
                    saw_done = True
                    continue
                d = json.loads(line[5:])
                if d["choices"]:
                    text += d["choices"][0]["text"]
                if d.get("usage"):
                    usage = d["usage"]
        assert saw_done and text == got["choices"][0]["text"] and usage["completion_tokens"] == 12
        # stop token id ends the request early with finish_reason=stop
        r = requests.post(url + "/v1/completions", json=dict(body, stop_token_ids=[ids[3]]), timeout=120).json()
        assert r["choices"][0]["finish_reason"] == "stop" and r["choices"][0]["token_ids"] == ids[:4]
        # stop string: cut at the match, engine request aborted
        full = got["choices"][0]["text"]
        if len(full) >= 4:
            stop = full[2:4]
            r = requests.post(url + "/v1/completions", json=dict(body, stop=[stop]), timeout=120).json()
            assert r["choices"][0]["finish_reason"] == "stop" and r["choices"][0]["text"] == full[:full.index(stop)]
        # errors + misc routes
        assert requests.post(url + "/v1/completions", json={"prompt": "text", "max_tokens": 2}, timeout=10).status_code == 400
        assert requests.post(url + "/v1/completions", data="{bad", timeout=10).status_code == 400
        assert requests.get(url + "/nope", timeout=10).status_code == 404
        assert requests.post(url + "/detokenize", json={"tokens": [3 + ord("h"), 3 + ord("i")]}, timeout=10).json()["prompt"] == "hi"
        m = requests.get(url + "/metrics", timeout=10).text
        assert "vllm:generation_tokens_total" in m and "vllm:request_success_total" in m
        # many concurrent keep-alive clients
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(8) as ex:
            outs = list(ex.map(lambda _: requests.post(url + "/v1/completions", json=body, timeout=120).json()["choices"][0]["token_ids"],
                               range(16)))
        assert all(o == ids for o in outs)
        # the C++ load generator (Connection: close + SSE) against the C++ front-end
        from vllm_b200.native import build_loadgen
        res = subprocess.run([build_loadgen(), "--host", "127.0.0.1", "--port", str(port), "--model", "llama-tiny",
                              "--num", "24", "--input-len", "16", "--output-len", "6", "--concurrency", "8",
                              "--vocab", "400"], capture_output=True, text=True, timeout=120)
        lg = json.loads(res.stdout.strip().splitlines()[-1])
        assert lg["completed"] == 24 and lg["failed"] == 0 and lg["output_throughput"] > 0
    finally:
        proc.terminate()
        try:
            proc.wait(timeout=30)
        except subprocess.TimeoutExpired:
            proc.kill()


def test_multiple_api_servers_share_one_engine(server):
    """--api-server-count 2: two HTTP processes accept on one socket and route to the same engine core by client index."""
    env = dict(os.environ, PYTHONPATH=ROOT, VLLM_LOGGING_LEVEL="WARNING")
    port = get_open_port()
    proc = _spawn([sys.executable, "-m", "vllm_b200", "serve", "llama-tiny", "--device", "cpu", "--dtype", "float32",
                   "--max-model-len", "256", "--max-num-seqs", "8", "--max-num-batched-tokens", "128",
                   "--kv-cache-memory-bytes", str(16 << 20), "--host", "127.0.0.1", "--port", str(port),
                   "--api-server-count", "2"], env)
    url = f"http://127.0.0.1:{port}"
    try:
        _wait_healthy([(proc, url)])
        body = {"model": "llama-tiny", "prompt": [9, 8, 7, 6, 5, 4], "max_tokens": 8, "temperature": 0}
        want = requests.post(server + "/v1/completions", json=body, timeout=120).json()["choices"][0]["text"]
        seen = set()
        deadline = time.time() + 60
        while len(seen) < 2 and time.time() < deadline:
            info = requests.get(url + "/server_info", timeout=30).json()        # new connection each time
            seen.add(info["api_server_index"])
            got = requests.post(url + "/v1/completions", json=body, timeout=120).json()
            assert got["choices"][0]["text"] == want
        assert seen == {0, 1}
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(8) as ex:
            outs = list(ex.map(lambda _: requests.post(url + "/v1/completions", json=body, timeout=120).json()["choices"][0]["text"],
                               range(16)))
        assert all(o == want for o in outs)
    finally:
        proc.terminate()
        try:
            proc.wait(timeout=40)
        except subprocess.TimeoutExpired:
            proc.kill()


def test_native_frontend_text_prompts_with_tiktoken_tokenizer(tmp_path):
    """Text prompts on the C++ tier: the launcher exports the BPE tables of a tiktoken-style tokenizer, the daemon
    tokenises (``/tokenize`` and ``prompt: "text"``) exactly like the Python tokenizer and detokenises the output."""
    import base64
    from vllm_b200.tokenizers.tiktoken_tok import TiktokenTokenizer
    ranks = {bytes([b]): b for b in range(256)}
    for i, t in enumerate([b"he", b"ll", b"hell", b"hello", b" w", b" wo", b" wor", b" world", b" t", b" th", b" the"]):
        ranks[t] = 256 + i
    with open(tmp_path / "model.tiktoken", "wb") as f:
        for t, r in ranks.items():
            f.write(base64.b64encode(t) + b" " + str(r).encode() + b"\n")
    with open(tmp_path / "tokenizer_config.json", "w") as f:
        json.dump({"added_tokens_decoder": {"267": {"content": "<|endoftext|>"}}, "eos_token": "<|endoftext|>"}, f)
    tok = TiktokenTokenizer.from_dir(str(tmp_path))
    env = dict(os.environ, PYTHONPATH=ROOT, VLLM_LOGGING_LEVEL="WARNING")
    port = get_open_port()
    proc = _spawn([sys.executable, "-m", "vllm_b200", "serve", "llama-tiny", "--tokenizer", str(tmp_path), "--device", "cpu",
                   "--dtype", "float32", "--max-model-len", "256", "--max-num-seqs", "8", "--max-num-batched-tokens", "128",
                   "--kv-cache-memory-bytes", str(16 << 20), "--host", "127.0.0.1", "--port", str(port), "--native-frontend"], env)
    url = f"http://127.0.0.1:{port}"
    try:
        _wait_healthy([(proc, url)])
        text = "hello world, the naïve 北京 test\n\n  1234 don't"
        want_ids = tok.encode(text)
        t = requests.post(url + "/tokenize", json={"prompt": text}, timeout=30).json()
        assert t["tokens"] == want_ids and t["count"] == len(want_ids)
        by_text = requests.post(url + "/v1/completions", json={"prompt": text, "max_tokens": 8, "temperature": 0}, timeout=120).json()
        by_ids = requests.post(url + "/v1/completions", json={"prompt": want_ids, "max_tokens": 8, "temperature": 0}, timeout=120).json()
        assert by_text["usage"]["prompt_tokens"] == len(want_ids)
        assert by_text["choices"][0]["token_ids"] == by_ids["choices"][0]["token_ids"]
        out_ids = [i for i in by_text["choices"][0]["token_ids"] if i < 267]
        assert by_text["choices"][0]["text"] == tok.enc.decode(out_ids, errors="replace")
        # chat completions on the C++ tier: messages rendered in the daemon's dialect (neutral role tags here), tokenised
        # natively, answered as chat.completion / chat.completion.chunk objects
        msgs = [{"role": "system", "content": "hello"}, {"role": "user", "content": [{"type": "text", "text": "the world"}]}]
        rendered = "<|system|>\nhello\n<|user|>\nthe world\n<|assistant|>\n"
        chat = requests.post(url + "/v1/chat/completions", json={"messages": msgs, "max_tokens": 6, "temperature": 0}, timeout=120).json()
        same = requests.post(url + "/v1/completions", json={"prompt": rendered, "max_tokens": 6, "temperature": 0,
                                                            "add_special_tokens": False}, timeout=120).json()
        assert chat["object"] == "chat.completion" and chat["id"].startswith("chatcmpl-")
        assert chat["choices"][0]["message"] == {"role": "assistant", "content": same["choices"][0]["text"]}
        assert chat["usage"]["prompt_tokens"] == same["usage"]["prompt_tokens"] == len(tok.encode(rendered))
        with requests.post(url + "/v1/chat/completions", json={"messages": msgs, "max_tokens": 6, "temperature": 0, "stream": True},
                           stream=True, timeout=120) as r:
            r.encoding = "utf-8"
            chunks = [json.loads(l[6:]) for l in r.iter_lines(decode_unicode=True) if l.startswith("data: {")]
        assert chunks and chunks[0]["object"] == "chat.completion.chunk" and chunks[0]["choices"][0]["delta"].get("role") == "assistant"
        assert "".join(c["choices"][0]["delta"].get("content", "") for c in chunks) == same["choices"][0]["text"]
        assert chunks[-1]["choices"][0]["finish_reason"] == "length"
        bad = requests.post(url + "/v1/chat/completions", json={"messages": [{"role": "user", "content": [{"type": "image_url"}]}]}, timeout=30)
        assert bad.status_code == 400 and "Python API server" in bad.json()["error"]["message"]
    finally:
        proc.terminate()
        try:
            proc.wait(timeout=30)
        except subprocess.TimeoutExpired:
            proc.kill()


def test_render_server_and_mm_processor_bench():
    """``launch render``: templating / tokenisation / detokenisation with no engine (TestClient over the same app the CLI
    serves); ``bench mm_processor``: the front-end processor benchmark runs for an image and an audio model."""
    from fastapi.testclient import TestClient
    from vllm_b200.entrypoints.render_server import build_render_app
    c = TestClient(build_render_app("llama-tiny", chat_template="chatml"))
    assert c.get("/health").status_code == 200
    r = c.post("/render", json={"messages": [{"role": "user", "content": "hi there"}]}).json()
    assert r["prompt"].startswith("<|im_start|>user\nhi there<|im_end|>") and r["prompt"].endswith("<|im_start|>assistant\n")
    assert r["count"] == len(r["token_ids"]) > 0
    back = c.post("/derender", json={"token_ids": r["token_ids"], "skip_special_tokens": False}).json()["text"]
    assert "hi there" in back
    t = c.post("/tokenize", json={"prompt": "abc def", "return_token_strs": True}).json()
    assert t["count"] == len(t["tokens"]) == len(t["token_strs"])
    assert c.post("/detokenize", json={"tokens": t["tokens"]}).json()["prompt"].endswith("abc def")
    assert c.get("/v1/models").json()["data"][0]["id"] == "llama-tiny"
    from vllm_b200.entrypoints.cli import _bench_parser
    from vllm_b200 import benchmarks as B
    a = _bench_parser().parse_args(["mm_processor", "--model", "llava-tiny", "--num-items", "6", "--image-size", "40x60", "--unique", "2"])
    res = B.bench_mm_processor(a)
    assert res["modality"] == "image" and res["items"] == 6 and res["items_per_s"] > 0 and res["placeholder_tokens_per_item"] == 4
    a = _bench_parser().parse_args(["mm_processor", "--model", "qwen2audio-tiny", "--num-items", "3", "--audio-seconds", "0.2"])
    res = B.bench_mm_processor(a)
    assert res["modality"] == "audio" and res["placeholder_tokens_per_item"] >= 1
