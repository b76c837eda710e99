"""OpenAI-compatible server tests against a real (tiny, CPU) engine running in its own process.
Mirrors tests/entrypoints/** of the reference (RemoteOpenAIServer + dummy weights), but in-process over ASGI."""
import asyncio
import json

import httpx
import pytest

from vllm_b200.engine.arg_utils import EngineArgs


@pytest.fixture(scope="module")
def server():
    from vllm_b200.engine.async_llm import AsyncLLM
    from vllm_b200.entrypoints.openai.api_server import ServingContext, build_app
    ea = EngineArgs(model="llama-tiny", device="cpu", dtype="float32", max_model_len=256, max_num_seqs=8,
                    max_num_batched_tokens=128, kv_cache_memory_bytes=8 << 20, disable_log_stats=False)
    loop = asyncio.new_event_loop()
    asyncio.set_event_loop(loop)
    engine = AsyncLLM.from_engine_args(ea)
    ctx = ServingContext(engine, "llama-tiny", 256, engine.tokenizer, tool_parser="hermes",
                         reasoning_parser=None)
    app = build_app(ctx)
    client = httpx.AsyncClient(transport=httpx.ASGITransport(app=app), base_url="http://test", timeout=120)

    def call(coro):
        return loop.run_until_complete(coro)

    yield client, call, engine
    call(client.aclose())
    engine.shutdown()
    loop.close()


def test_health_models_tokenize(server):
    client, call, _ = server
    assert call(client.get("/health")).status_code == 200
    models = call(client.get("/v1/models")).json()
    assert models["data"][0]["id"] == "llama-tiny"
    tok = call(client.post("/tokenize", json={"prompt": "hello"})).json()
    assert tok["count"] == len(tok["tokens"]) == 6
    detok = call(client.post("/detokenize", json={"tokens": tok["tokens"][1:]})).json()
    assert detok["prompt"] == "hello"


def test_completion_and_logprobs(server):
    client, call, _ = server
    r = call(client.post("/v1/completions", json={"model": "llama-tiny", "prompt": "hello", "max_tokens": 5,
                                                   "temperature": 0, "logprobs": 2})).json()
    ch = r["choices"][0]
    assert r["usage"]["completion_tokens"] == 5 and ch["finish_reason"] == "length"
    assert len(ch["logprobs"]["tokens"]) == 5 and len(ch["logprobs"]["top_logprobs"][0]) >= 2
    # deterministic
    r2 = call(client.post("/v1/completions", json={"prompt": "hello", "max_tokens": 5, "temperature": 0})).json()
    assert r2["choices"][0]["text"] == ch["text"]
    # batch of prompts + n
    r3 = call(client.post("/v1/completions", json={"prompt": ["a", "b"], "max_tokens": 3, "n": 2, "seed": 1})).json()
    assert sorted(c["index"] for c in r3["choices"]) == [0, 1, 2, 3]
    # unknown model → 404, bad params → 400
    assert call(client.post("/v1/completions", json={"model": "nope", "prompt": "x"})).status_code == 404
    assert call(client.post("/v1/completions", json={"prompt": "x", "temperature": -1})).status_code == 400
    assert call(client.post("/v1/completions", json={"prompt": "x" * 1000, "max_tokens": 1})).status_code == 400


def test_completion_streaming(server):
    client, call, _ = server

    async def run():
        text, done, usage = "", False, None
        async with client.stream("POST", "/v1/completions", json={
                "prompt": "hello", "max_tokens": 6, "temperature": 0, "stream": True,
                "stream_options": {"include_usage": True}}) as resp:
            assert resp.status_code == 200
            async for line in resp.aiter_lines():
                if not line.startswith("data: "):
                    continue
                payload = line[6:]
                if payload == "[DONE]":
                    done = True
                    break
                obj = json.loads(payload)
                if obj.get("usage"):
                    usage = obj["usage"]
                for c in obj["choices"]:
                    text += c["text"]
        return text, done, usage

    text, done, usage = call(run())
    full = call(client.post("/v1/completions", json={"prompt": "hello", "max_tokens": 6, "temperature": 0})).json()
    assert done and text == full["choices"][0]["text"] and usage["completion_tokens"] == 6


def test_chat_completion_and_stream(server):
    client, call, _ = server
    msgs = [{"role": "system", "content": "be brief"}, {"role": "user", "content": "hi"}]
    r = call(client.post("/v1/chat/completions", json={"messages": msgs, "max_tokens": 4, "temperature": 0,
                                                        "logprobs": True, "top_logprobs": 2})).json()
    ch = r["choices"][0]
    assert ch["message"]["role"] == "assistant" and r["usage"]["completion_tokens"] == 4
    assert len(ch["logprobs"]["content"]) == 4

    async def run():
        parts, roles = [], []
        async with client.stream("POST", "/v1/chat/completions", json={"messages": msgs, "max_tokens": 4,
                                                                       "temperature": 0, "stream": True}) as resp:
            async for line in resp.aiter_lines():
                if line.startswith("data: ") and line[6:] != "[DONE]":
                    d = json.loads(line[6:])["choices"]
                    if d:
                        roles.append(d[0]["delta"].get("role"))
                        parts.append(d[0]["delta"].get("content") or "")
        return "".join(parts), roles

    text, roles = call(run())
    assert roles[0] == "assistant" and text == (ch["message"]["content"] or "")


def test_responses_and_anthropic_and_metrics(server):
    client, call, _ = server
    r = call(client.post("/v1/responses", json={"input": "hi", "max_output_tokens": 3, "temperature": 0})).json()
    assert r["object"] == "response" and r["usage"]["output_tokens"] == 3
    a = call(client.post("/v1/messages", json={"messages": [{"role": "user", "content": "hi"}], "max_tokens": 3,
                                               "temperature": 0})).json()
    assert a["type"] == "message" and a["stop_reason"] == "max_tokens"
    # typed streaming events: Responses API
    async def sse(path, body):
        evs = []
        async with client.stream("POST", path, json=body) as s:
            cur = None
            async for line in s.aiter_lines():
                if line.startswith("event:"):
                    cur = line[6:].strip()
                elif line.startswith("data:"):
                    evs.append((cur, json.loads(line[5:])))
        return evs
    evs = call(sse("/v1/responses", {"input": "hi", "max_output_tokens": 4, "temperature": 0, "stream": True, "store": True}))
    kinds = [k for k, _ in evs]
    assert kinds[0] == "response.created" and kinds[-1] == "response.completed"
    assert "response.output_text.delta" in kinds and "response.output_item.done" in kinds
    done = evs[-1][1]["response"]
    assert "".join(d["delta"] for k, d in evs if k == "response.output_text.delta") == done["output_text"]
    assert done["usage"]["output_tokens"] == 4
    assert [d["sequence_number"] for _, d in evs] == list(range(1, len(evs) + 1))
    assert call(client.get(f"/v1/responses/{done['id']}")).json()["output_text"] == done["output_text"]
    assert call(client.get("/v1/responses/resp_nope")).status_code == 404
    # Anthropic streaming
    evs = call(sse("/v1/messages", {"messages": [{"role": "user", "content": "hi"}], "max_tokens": 3, "stream": True,
                                    "temperature": 0}))
    kinds = [k for k, _ in evs]
    assert kinds[:2] == ["message_start", "content_block_start"] and kinds[-2:] == ["message_delta", "message_stop"]
    assert evs[-2][1]["delta"]["stop_reason"] == "max_tokens" and evs[-2][1]["usage"]["output_tokens"] == 3
    assert "".join(d["delta"]["text"] for k, d in evs if k == "content_block_delta") == a["content"][0]["text"]
    # Cohere v2 chat
    c = call(client.post("/v2/chat", json={"messages": [{"role": "user", "content": "hi"}], "max_tokens": 3})).json()
    assert c["finish_reason"] == "MAX_TOKENS" and c["message"]["content"][0]["type"] == "text"
    m = call(client.get("/metrics")).text
    assert "vllm:generation_tokens_total" in m and "vllm:time_to_first_token_seconds" in m
    # admin routes are not mounted on a default (production) server
    assert call(client.post("/reset_prefix_cache")).status_code == 404
    assert call(client.post("/collective_rpc", json={"method": "save_sharded_state"})).status_code == 404
    assert call(client.post("/v1/load_lora_adapter", json={"lora_name": "x", "lora_path": "/tmp"})).status_code == 404


def test_dev_routes_need_dev_mode_and_api_key_covers_everything(server, monkeypatch):
    """Reference behaviour: VLLM_SERVER_DEV_MODE mounts the admin routes; an API key then guards every route but the
    liveness / scrape endpoints, and /collective_rpc only forwards allow-listed worker methods."""
    from vllm_b200.entrypoints.openai.api_server import ServingContext, build_app
    _, call, engine = server
    monkeypatch.setenv("VLLM_SERVER_DEV_MODE", "1")
    ctx = ServingContext(engine, "llama-tiny", 256, engine.tokenizer, api_key="sekret")
    client = httpx.AsyncClient(transport=httpx.ASGITransport(app=build_app(ctx)), base_url="http://test", timeout=60)
    auth = {"Authorization": "Bearer sekret"}
    assert call(client.get("/health")).status_code == 200
    assert call(client.post("/reset_prefix_cache")).status_code == 401
    assert call(client.post("/tokenize", json={"prompt": "hello"})).status_code == 401
    assert call(client.post("/reset_prefix_cache", headers=auth)).status_code in (200, 409)
    r = call(client.post("/collective_rpc", json={"method": "save_sharded_state", "args": ["/tmp/x"]}, headers=auth))
    assert r.status_code == 403
    assert call(client.get("/is_sleeping", headers=auth)).json() == {"is_sleeping": False}
    call(client.aclose())


def test_orca_load_headers(server):
    client, call, _ = server
    r = call(client.post("/v1/completions", json={"prompt": [5, 6, 7], "max_tokens": 2, "temperature": 0},
                         headers={"endpoint-load-metrics-format": "JSON"}))
    rep = r.headers["endpoint-load-metrics"]
    assert rep.startswith("JSON ") and "vllm:num_requests_running" in json.loads(rep[5:])["named_metrics"]
    r = call(client.post("/v1/completions", json={"prompt": [5, 6, 7], "max_tokens": 2, "temperature": 0},
                         headers={"endpoint-load-metrics-format": "TEXT"}))
    assert r.headers["endpoint-load-metrics"].startswith("TEXT named_metrics.vllm:num_requests_running=")


def test_tool_and_reasoning_parsers():
    from vllm_b200.reasoning import get_reasoning_parser
    from vllm_b200.tool_parsers import get_tool_parser
    p = get_tool_parser("hermes")()
    out = p.extract('ok <tool_call>\n{"name": "f", "arguments": {"x": 1}}\n</tool_call>')
    assert out["tool_calls"][0]["function"]["name"] == "f" and json.loads(out["tool_calls"][0]["function"]["arguments"]) == {"x": 1}
    assert out["content"] == "ok"
    m = get_tool_parser("mistral")().extract('[TOOL_CALLS][{"name": "g", "arguments": {"a": "b"}}]')
    assert m["tool_calls"][0]["function"]["name"] == "g"
    py = get_tool_parser("pythonic")().extract('[get_weather(city="Paris", days=2), f()]')
    assert [c["function"]["name"] for c in py["tool_calls"]] == ["get_weather", "f"]
    l3 = get_tool_parser("llama3_json")().extract('{"name": "h", "parameters": {"q": 2}}')
    assert l3["tool_calls"][0]["function"]["name"] == "h"
    # streaming: content passes through, the call is emitted once complete
    sp = get_tool_parser("hermes")()
    text, got = "", []
    for d in ["Sure", ". <tool", "_call>{\"name\": \"f\", ", "\"arguments\": {}}", "</tool_call>"]:
        prev, text = text, text + d
        r = sp.extract_streaming(prev, text, d)
        if r:
            got.append(r)
    assert "".join(g.get("content") or "" for g in got).strip() == "Sure."
    assert sum(len(g.get("tool_calls", [])) for g in got) == 1
    rp = get_reasoning_parser("deepseek_r1")()
    assert rp.extract("<think>hmm</think>\nanswer") == ("hmm", "answer")
    assert rp.extract("still thinking") == ("still thinking", None)
    q = get_reasoning_parser("qwen3")()
    assert q.extract("plain") == (None, "plain")


def test_audio_transcription_api():
    """/v1/audio/transcriptions + /translations against a tiny Whisper: WAV upload, formats, streaming, chunking."""
    import io
    import wave

    import numpy as np

    from vllm_b200.engine.async_llm import AsyncLLM
    from vllm_b200.entrypoints.openai.api_server import ServingContext, build_app
    ea = EngineArgs(model="whisper-tiny-test", device="cpu", dtype="float32", max_model_len=32, max_num_seqs=4,
                    max_num_batched_tokens=64, kv_cache_memory_bytes=4 << 20, enable_prefix_caching=False)
    loop = asyncio.new_event_loop()
    asyncio.set_event_loop(loop)
    engine = AsyncLLM.from_engine_args(ea)
    app = build_app(ServingContext(engine, "whisper-tiny-test", 32, engine.tokenizer))
    client = httpx.AsyncClient(transport=httpx.ASGITransport(app=app), base_url="http://test", timeout=300)
    call = loop.run_until_complete

    def wav_bytes(seconds: float, sr: int = 8000) -> bytes:
        t = np.arange(int(seconds * sr)) / sr
        pcm = (0.3 * np.sin(2 * np.pi * 330 * t) * 32767).astype("<i2")
        buf = io.BytesIO()
        with wave.open(buf, "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(sr)
            w.writeframes(pcm.tobytes())
        return buf.getvalue()
    try:
        short = wav_bytes(2.0)
        r = call(client.post("/v1/audio/transcriptions", files={"file": ("a.wav", short, "audio/wav")},
                             data={"model": "whisper-tiny-test"}))
        assert r.status_code == 200, r.text
        body = r.json()
        assert isinstance(body["text"], str) and body["usage"]["seconds"] == 2
        r2 = call(client.post("/v1/audio/transcriptions", files={"file": ("a.wav", short, "audio/wav")},
                              data={"response_format": "text"}))
        assert r2.text == body["text"]                                  # greedy → deterministic
        # 65 s of audio → three windows, verbose_json reports one segment per window with contiguous bounds
        r3 = call(client.post("/v1/audio/translations", files={"file": ("b.wav", wav_bytes(65.0), "audio/wav")},
                              data={"response_format": "verbose_json"})).json()
        assert r3["task"] == "translate" and len(r3["segments"]) == 3
        assert r3["segments"][0]["start"] == 0 and abs(r3["segments"][-1]["end"] - 65.0) < 1e-6
        assert all(abs(a["end"] - b["start"]) < 1e-9 for a, b in zip(r3["segments"], r3["segments"][1:]))
        srt = call(client.post("/v1/audio/transcriptions", files={"file": ("a.wav", short, "audio/wav")},
                               data={"response_format": "srt"})).text
        assert srt.startswith("1\n00:00:00,000 --> 00:00:02,000")
        # streaming: deltas concatenate to the non-streamed text
        async def stream():
            out = ""
            async with client.stream("POST", "/v1/audio/transcriptions", files={"file": ("a.wav", short, "audio/wav")},
                                     data={"stream": "true"}) as s:
                async for line in s.aiter_lines():
                    if line.startswith("data:") and "[DONE]" not in line:
                        out += json.loads(line[5:])["choices"][0]["delta"]["content"]
            return out
        assert call(stream()).strip() == body["text"]
        bad = call(client.post("/v1/audio/transcriptions", files={"file": ("x.wav", b"not audio", "audio/wav")}))
        assert bad.status_code == 400
    finally:
        call(client.aclose())
        engine.shutdown()
        loop.close()


def test_grpc_engine_service(server):
    """gRPC front-end next to the HTTP app: unary + server-streaming generation, health, model info, bad input."""
    import grpc
    from vllm_b200.entrypoints.grpc import EngineClient, start_grpc_server
    client, call, engine = server

    async def run():
        srv, port = await start_grpc_server(engine, "llama-tiny", "127.0.0.1", 0)
        c = EngineClient(f"127.0.0.1:{port}")
        try:
            h = await c.health()
            info = await c.model_info()
            full = [r async for r in c.generate(prompt_token_ids=[5, 6, 7, 8], max_tokens=5, temperature=0.0)]
            chunks = [r async for r in c.generate(prompt_token_ids=[5, 6, 7, 8], max_tokens=5, temperature=0.0, stream=True)]
            try:
                [r async for r in c.generate(max_tokens=2)]
                bad = None
            except grpc.aio.AioRpcError as e:
                bad = e.code()
            return h, info, full, chunks, bad
        finally:
            await c.close()
            await srv.stop(0)
    h, info, full, chunks, bad = call(run())
    assert h.healthy and info.model == "llama-tiny" and info.max_model_len == 256
    assert len(full) == 1 and full[0].finished and len(full[0].token_ids) == 5 and full[0].finish_reason == "length"
    assert [t for ch in chunks for t in ch.token_ids] == list(full[0].token_ids)
    assert "".join(ch.text for ch in chunks) == full[0].text and chunks[-1].finished and not chunks[0].finished
    assert bad == grpc.StatusCode.INVALID_ARGUMENT
    http = call(client.post("/inference/v1/generate", json={"token_ids": [5, 6, 7, 8], "sampling_params": {
        "max_tokens": 5, "temperature": 0}})).json()
    assert http["token_ids"] == list(full[0].token_ids)


def test_realtime_transcription_websocket():
    """/v1/realtime: append PCM16 chunks, commit, receive transcription deltas + the completed transcript, which
    equals the batch endpoint's answer for the same audio."""
    import base64
    import io
    import wave

    import numpy as np
    from starlette.testclient import TestClient

    from vllm_b200.engine.async_llm import AsyncLLM
    from vllm_b200.entrypoints.openai.api_server import ServingContext, build_app
    ea = EngineArgs(model="whisper-tiny-test", device="cpu", dtype="float32", max_model_len=32, max_num_seqs=4,
                    max_num_batched_tokens=64, kv_cache_memory_bytes=4 << 20, enable_prefix_caching=False)
    engine = AsyncLLM.from_engine_args(ea)
    app = build_app(ServingContext(engine, "whisper-tiny-test", 32, engine.tokenizer))
    t = np.arange(16000 * 2) / 16000
    pcm = (0.3 * np.sin(2 * np.pi * 330 * t) * 32767).astype("<i2")
    buf = io.BytesIO()
    with wave.open(buf, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
        w.writeframes(pcm.tobytes())
    try:
        with TestClient(app) as client:
            want = client.post("/v1/audio/transcriptions", files={"file": ("a.wav", buf.getvalue(), "audio/wav")}).json()["text"]
            with client.websocket_connect("/v1/realtime") as ws:
                assert ws.receive_json()["type"] == "transcription_session.created"
                ws.send_json({"type": "session.update", "session": {"input_audio_transcription": {"language": "en"}}})
                assert ws.receive_json()["type"] == "transcription_session.updated"
                raw = pcm.tobytes()
                for i in range(0, len(raw), 16000):
                    ws.send_json({"type": "input_audio_buffer.append", "audio": base64.b64encode(raw[i:i + 16000]).decode()})
                ws.send_json({"type": "input_audio_buffer.commit"})
                assert ws.receive_json()["type"] == "input_audio_buffer.committed"
                deltas, done = "", None
                while done is None:
                    ev = ws.receive_json()
                    if ev["type"].endswith(".delta"):
                        deltas += ev["delta"]
                    elif ev["type"].endswith(".completed"):
                        done = ev["transcript"]
                assert done == want and deltas.strip() == want
                ws.send_json({"type": "input_audio_buffer.commit"})
                assert ws.receive_json()["type"] == "error"
    finally:
        engine.shutdown()


@pytest.fixture()
def mcp_server():
    """A tiny MCP server (streamable HTTP transport): ``add`` answers as JSON, ``tools/list`` as an event stream."""
    import threading
    from aiohttp import web
    from vllm_b200.utils import get_open_port
    seen = []

    async def handle(request):
        msg = await request.json()
        seen.append((msg.get("method"), request.headers.get("Mcp-Session-Id")))
        m = msg.get("method")
        if "id" not in msg:
            return web.Response(status=202)
        if m == "initialize":
            res = {"protocolVersion": "2025-03-26", "capabilities": {"tools": {}}, "serverInfo": {"name": "calc", "version": "1"}}
            return web.json_response({"jsonrpc": "2.0", "id": msg["id"], "result": res}, headers={"Mcp-Session-Id": "s-1"})
        if m == "tools/list":
            res = {"tools": [{"name": "add", "description": "add two ints", "inputSchema": {
                "type": "object", "properties": {"a": {"type": "integer"}, "b": {"type": "integer"}}, "required": ["a", "b"]}},
                {"name": "boom", "description": "always fails", "inputSchema": {"type": "object", "properties": {}}}]}
            body = "event: message\ndata: " + json.dumps({"jsonrpc": "2.0", "id": msg["id"], "result": res}) + "\n\n"
            return web.Response(text=body, content_type="text/event-stream")
        if m == "tools/call":
            p = msg["params"]
            if p["name"] == "add":
                res = {"content": [{"type": "text", "text": str(p["arguments"]["a"] + p["arguments"]["b"])}]}
            else:
                res = {"content": [{"type": "text", "text": "kaput"}], "isError": True}
            return web.json_response({"jsonrpc": "2.0", "id": msg["id"], "result": res})
        return web.json_response({"jsonrpc": "2.0", "id": msg["id"], "error": {"code": -32601, "message": "no such method"}})

    port = get_open_port()
    loop = asyncio.new_event_loop()
    started = threading.Event()
    runner_box = []

    def run():
        asyncio.set_event_loop(loop)
        app = web.Application()
        app.router.add_post("/mcp", handle)
        runner = web.AppRunner(app)
        loop.run_until_complete(runner.setup())
        loop.run_until_complete(web.TCPSite(runner, "127.0.0.1", port).start())
        runner_box.append(runner)
        started.set()
        loop.run_forever()
    th = threading.Thread(target=run, daemon=True)
    th.start()
    assert started.wait(10)
    yield f"http://127.0.0.1:{port}/mcp", seen
    asyncio.run_coroutine_threadsafe(runner_box[0].cleanup(), loop).result(10)
    loop.call_soon_threadsafe(loop.stop)
    th.join(5)


def test_mcp_client_and_tool_loop(mcp_server):
    from vllm_b200.entrypoints.tool_server import MCPClient, MCPError, ToolServer, input_items_to_messages, run_tool_loop
    url, seen = mcp_server

    async def main():
        c = MCPClient(url, "calc")
        tools = await c.list_tools()
        assert [t["name"] for t in tools] == ["add", "boom"] and c.server_info["name"] == "calc"
        assert await c.call_tool("add", {"a": 2, "b": 40}) == "42"
        with pytest.raises(MCPError, match="kaput"):
            await c.call_tool("boom", {})
        with pytest.raises(MCPError, match="no such method"):
            await c.request("resources/list")
        # the session id handed out by initialize is echoed afterwards
        assert seen[0] == ("initialize", None) and all(sid == "s-1" for _, sid in seen[1:])

        ts = ToolServer()
        fns, route, items = await ts.resolve([
            {"type": "mcp", "server_url": url, "server_label": "calc", "allowed_tools": ["add"]},
            {"type": "function", "name": "get_weather", "parameters": {"type": "object", "properties": {}}}])
        assert [f["function"]["name"] for f in fns] == ["add", "get_weather"] and list(route) == ["add"]
        assert items == [{"type": "mcp_list_tools", "server_label": "calc", "tools": [
            {"name": "add", "description": "add two ints", "input_schema": tools[0]["inputSchema"]}]}]

        # a scripted model: asks for add(20, 22), then answers with the tool result
        turns = []

        async def chat(messages, tools):
            turns.append([m["role"] for m in messages])
            if messages[-1]["role"] != "tool":
                return {"message": {"content": "", "tool_calls": [{"id": "c1", "type": "function", "function": {
                    "name": "add", "arguments": json.dumps({"a": 20, "b": 22})}}]},
                    "usage": {"prompt_tokens": 5, "completion_tokens": 3, "total_tokens": 8}}
            return {"message": {"content": f"the sum is {messages[-1]['content']}"},
                    "usage": {"prompt_tokens": 9, "completion_tokens": 4, "total_tokens": 13}}
        msg, calls, usage = await run_tool_loop(chat, [{"role": "user", "content": "20+22?"}], fns, route)
        assert msg["content"] == "the sum is 42" and usage["total_tokens"] == 21
        assert turns == [["user"], ["user", "assistant", "tool"]]
        assert [(c["type"], c["name"], c["output"], c["error"]) for c in calls] == [("mcp_call", "add", "42", None)]

        # a call to a client-side function ends the loop with the call left for the client
        async def chat2(messages, tools):
            return {"message": {"content": "", "tool_calls": [{"id": "c2", "type": "function", "function": {
                "name": "get_weather", "arguments": "{}"}}]}, "usage": None}
        msg, calls, _ = await run_tool_loop(chat2, [{"role": "user", "content": "weather?"}], fns, route)
        assert calls == [] and msg["tool_calls"][0]["function"]["name"] == "get_weather"
    asyncio.run(main())
    msgs = input_items_to_messages([
        {"role": "user", "content": [{"type": "input_text", "text": "weather?"}]},
        {"type": "function_call", "call_id": "c2", "name": "get_weather", "arguments": "{}"},
        {"type": "function_call_output", "call_id": "c2", "output": "sunny"}])
    assert [m["role"] for m in msgs] == ["user", "assistant", "tool"]
    assert msgs[1]["tool_calls"][0]["id"] == "c2" and msgs[2] == {"role": "tool", "tool_call_id": "c2", "content": "sunny"}


def test_responses_api_with_mcp_tools(server, mcp_server):
    client, call, _ = server
    url, seen = mcp_server
    body = {"model": "llama-tiny", "input": "add 1 and 2", "max_output_tokens": 8, "temperature": 0.0, "store": True,
            "tools": [{"type": "mcp", "server_url": url, "server_label": "calc"}]}
    r = call(client.post("/v1/responses", json=body))
    assert r.status_code == 200, r.text
    d = r.json()
    kinds = [o["type"] for o in d["output"]]
    assert kinds[0] == "mcp_list_tools" and kinds[-1] == "message" and d["status"] == "completed"
    assert [t["name"] for t in d["output"][0]["tools"]] == ["add", "boom"]
    assert d["usage"]["output_tokens"] == 8 and ("tools/list", "s-1") in seen
    assert call(client.get(f"/v1/responses/{d['id']}")).json()["id"] == d["id"]
    # streaming variant: typed events, one added/done pair per output item
    r = call(client.post("/v1/responses", json=dict(body, stream=True)))
    evs = [l.split(": ", 1)[1] for l in r.text.splitlines() if l.startswith("event: ")]
    assert evs[0] == "response.created" and evs[-1] == "response.completed"
    assert evs.count("response.output_item.added") == evs.count("response.output_item.done") == 2
    # an unreachable tool server is a dependency failure, not a 500
    bad = dict(body, tools=[{"type": "mcp", "server_url": "http://127.0.0.1:9/mcp", "server_label": "x"}])
    assert call(client.post("/v1/responses", json=bad)).status_code == 424
