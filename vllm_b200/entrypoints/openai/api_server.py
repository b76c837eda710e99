"""OpenAI-compatible HTTP server (FastAPI + uvicorn).

Behavioural reference: vllm/entrypoints/launchers/ (``run_server`` api_server/entry.py:163, router registration
routers.py:12, ``serve_http`` + watchdog launcher.py:50,180), openai/{completion,chat_completion}/serving.py,
serve/ dev routes (sleep, profile, cache reset, tokenize), instrumentator/metrics.py, anthropic/, responses/.

Routes: /health /ping /version /v1/models /v1/completions /v1/chat/completions /v1/embeddings /v1/responses
/v1/messages (Anthropic) /tokenize /detokenize /score /metrics /start_profile /stop_profile /sleep /wake_up
/is_sleeping /reset_prefix_cache /server_info /load.
"""
from __future__ import annotations

import argparse
import asyncio
import base64
import json
import os
import signal
import socket
import struct
import time
from contextlib import asynccontextmanager
from typing import Any, AsyncGenerator, Optional

from fastapi import FastAPI, Request, WebSocket  # noqa: F401
from fastapi.responses import JSONResponse, Response, StreamingResponse

from vllm_b200 import envs
from vllm_b200.engine.arg_utils import EngineArgs
from vllm_b200.engine.exceptions import EngineDeadError
from vllm_b200.entrypoints.chat_utils import render_chat
from vllm_b200.entrypoints.openai import protocol as P
from vllm_b200.logger import init_logger
from vllm_b200.sampling_params import PoolingParams, RequestOutputKind, SamplingParams
from vllm_b200.utils import random_uuid
from vllm_b200.version import __version__

logger = init_logger(__name__)


class ServingContext:
    """Everything the route handlers need (engine client may be a real AsyncLLM or a test fake)."""

    def __init__(self, engine, model_name: str, max_model_len: int, tokenizer, api_key: Optional[str] = None,
                 tool_parser: Optional[str] = None, reasoning_parser: Optional[str] = None,
                 enable_auto_tools: bool = False, chat_template: Optional[str] = None) -> None:
        self.engine = engine
        self.model_name = model_name
        self.max_model_len = max_model_len
        self.tokenizer = tokenizer
        self.api_key = api_key
        self.tool_parser = tool_parser
        self.reasoning_parser = reasoning_parser
        self.enable_auto_tools = enable_auto_tools
        self.chat_template = chat_template
        self.inflight = 0
        self.responses_store: dict[str, dict] = {}
        self.tool_server = None                       # entrypoints/tool_server.ToolServer (``--tool-server``)
        self.lora_ids: dict[str, int] = {}
        self.lora_requests: dict[str, Any] = {}       # adapter name (usable as `model`) → LoRARequest


def _error(msg: str, code: int = 400, typ: str = "BadRequestError"):
    return JSONResponse(P.ErrorResponse(error=P.ErrorInfo(message=msg, type=typ, code=code)).model_dump(),
                        status_code=code)


def _usage(n_prompt: int, n_out: int, cached: Optional[int] = None) -> P.UsageInfo:
    return P.UsageInfo(prompt_tokens=n_prompt, completion_tokens=n_out, total_tokens=n_prompt + n_out,
                       prompt_tokens_details={"cached_tokens": cached} if cached else None)


def _completion_logprobs(ctx: ServingContext, token_ids, logprobs, k: int, offset0: int = 0) -> P.CompletionLogProbs:
    out = P.CompletionLogProbs()
    off = offset0
    for tid, d in zip(token_ids, logprobs):
        if d is None:
            out.tokens.append(ctx.tokenizer.decode([tid]))
            out.token_logprobs.append(None)
            out.top_logprobs.append(None)
            out.text_offset.append(off)
            continue
        tok = d[tid].decoded_token if d.get(tid) and d[tid].decoded_token is not None else ctx.tokenizer.decode([tid])
        out.tokens.append(tok)
        out.token_logprobs.append(max(d[tid].logprob, -9999.0) if tid in d else None)
        out.top_logprobs.append({(v.decoded_token if v.decoded_token is not None else ctx.tokenizer.decode([t])):
                                 max(v.logprob, -9999.0) for t, v in list(d.items())[:max(k, 1) + 1]})
        out.text_offset.append(off)
        off += len(tok)
    return out


def _chat_logprobs(ctx: ServingContext, token_ids, logprobs, k: int) -> P.ChatLogprobs:
    content = []
    for tid, d in zip(token_ids, logprobs):
        def tokstr(t, v):
            return v.decoded_token if v.decoded_token is not None else ctx.tokenizer.decode([t])
        main = d.get(tid)
        tok = tokstr(tid, main) if main else ctx.tokenizer.decode([tid])
        tops = [{"token": tokstr(t, v), "logprob": max(v.logprob, -9999.0), "bytes": list(tokstr(t, v).encode())}
                for t, v in list(d.items())[:k]] if k else []
        content.append(P.ChatLogprob(token=tok, logprob=max(main.logprob, -9999.0) if main else -9999.0,
                                     bytes=list(tok.encode()), top_logprobs=tops))
    return P.ChatLogprobs(content=content)


_OPEN_PATHS = frozenset({"/health", "/ping", "/version", "/metrics", "/load"})
# worker methods an operator may invoke over HTTP in dev mode (extend with VLLM_SERVER_DEV_RPC_METHODS=a,b)
_RPC_ALLOWED = frozenset({"checkpoint_prepare", "checkpoint_restore", "reset_after_fault", "get_memory_stats",
                          "report_device_memory", "is_sleeping", "list_loras", "update_weights", "init_weight_transfer"})


def build_app(ctx: ServingContext):
    @asynccontextmanager
    async def lifespan(app):
        yield

    app = FastAPI(lifespan=lifespan, title="vllm_b200")
    app.state.ctx = ctx

    @app.middleware("http")
    async def auth(request: Request, call_next):
        # every route except liveness / scrape endpoints is behind the key (the reference guards /v1 only and keeps
        # its admin routes out of production builds; here both protections apply)
        if ctx.api_key and request.url.path not in _OPEN_PATHS and request.method != "OPTIONS":
            if request.headers.get("Authorization") != f"Bearer {ctx.api_key}":
                return JSONResponse({"error": "Unauthorized"}, status_code=401)
        rid = request.headers.get("X-Request-Id")
        t0 = time.perf_counter()
        resp = await call_next(request)
        if rid:
            resp.headers["X-Request-Id"] = rid
        if getattr(ctx, "log_requests", False) and request.method == "POST":
            # request logger (reference: entrypoints/serve/utils/request_logger.py): who asked for what, never the payload
            logger.info("request %s %s -> %d in %.1f ms (client %s, id %s)", request.method, request.url.path,
                        resp.status_code, (time.perf_counter() - t0) * 1e3, request.client.host if request.client else "?", rid)
        fmt = request.headers.get("endpoint-load-metrics-format")
        if fmt:
            # ORCA load report for load balancers (reference: entrypoints/serve/utils/orca_metrics.py)
            try:
                running, waiting = await ctx.engine.get_request_counts()
                named = {"vllm:num_requests_running": float(running), "vllm:num_requests_waiting": float(waiting)}
                if fmt.lower() == "json":
                    resp.headers["endpoint-load-metrics"] = "JSON " + json.dumps({"named_metrics": named})
                else:
                    resp.headers["endpoint-load-metrics"] = "TEXT " + ", ".join(f"named_metrics.{k}={v}" for k, v in named.items())
            except Exception:
                pass
        return resp

    def check_model(name: Optional[str]):
        if name and name != ctx.model_name:
            return _error(f"The model `{name}` does not exist.", 404, "NotFoundError")
        return None

    # ------------------------------------------------------------- basics
    @app.get("/health")
    @app.get("/ping")
    async def health():
        try:
            await ctx.engine.check_health()
        except Exception:
            return Response(status_code=503)
        return Response(status_code=200)

    @app.get("/version")
    async def version():
        return {"version": __version__}

    @app.get("/v1/models")
    async def models():
        return P.ModelList(data=[P.ModelCard(id=ctx.model_name, root=ctx.model_name,
                                             max_model_len=ctx.max_model_len)]).model_dump()

    @app.get("/metrics")
    async def metrics():
        from vllm_b200.metrics import render_prometheus
        body, ctype = render_prometheus()
        return Response(content=body, media_type=ctype)

    @app.get("/load")
    async def load():
        return {"server_load": ctx.inflight}

    @app.get("/server_info")
    async def server_info():
        cfg = await ctx.engine.get_vllm_config()
        return {"version": __version__, "model": ctx.model_name, "config_hash": cfg.compute_hash(),
                "api_server_index": int(os.environ.get("VLLM_B200_API_SERVER_INDEX", "0")), "pid": os.getpid()}

    # ------------------------------------------------------------ tokenize
    @app.post("/tokenize")
    async def tokenize(req: P.TokenizeRequest):
        if req.messages is not None:
            text = render_chat(ctx.tokenizer, req.messages, ctx.chat_template, req.add_generation_prompt)
            ids = ctx.tokenizer.encode(text, add_special_tokens=False)
        else:
            ids = ctx.tokenizer.encode(req.prompt or "", add_special_tokens=req.add_special_tokens)
        strs = ctx.tokenizer.convert_ids_to_tokens(ids) if req.return_token_strs else None
        return P.TokenizeResponse(count=len(ids), max_model_len=ctx.max_model_len, tokens=ids, token_strs=strs)

    @app.post("/detokenize")
    async def detokenize(req: P.DetokenizeRequest):
        return P.DetokenizeResponse(prompt=ctx.tokenizer.decode(req.tokens, skip_special_tokens=False))

    # ---------------------------------------------------------- completions
    def _prompts(prompt) -> list[Any]:
        if isinstance(prompt, str):
            return [prompt]
        if isinstance(prompt, list) and prompt and isinstance(prompt[0], int):
            return [{"prompt_token_ids": prompt}]
        if isinstance(prompt, list) and prompt and isinstance(prompt[0], list):
            return [{"prompt_token_ids": p} for p in prompt]
        return list(prompt or [])

    @app.post("/v1/completions")
    async def completions(req: P.CompletionRequest, raw: Request):
        if (e := check_model(req.model)) is not None:
            return e
        prompts = _prompts(req.prompt)
        if not prompts:
            return _error("prompt must not be empty")
        rid_base = f"cmpl-{req.request_id or random_uuid()}"
        try:
            sp = req.to_sampling_params(16, req.logprobs, req.prompt_logprobs if req.prompt_logprobs is not None
                                        else (req.logprobs if req.echo else None))
        except (ValueError, TypeError) as ex:
            return _error(str(ex))
        gens = []
        try:
            for i, p in enumerate(prompts):
                gens.append(ctx.engine.generate(p, sp, f"{rid_base}-{i}", priority=req.priority,
                                                trace_headers=dict(raw.headers),
                                                lora_request=ctx.lora_requests.get(req.model)))
        except ValueError as ex:
            return _error(str(ex))
        created = int(time.time())

        def echo_text(out) -> str:
            return (out.prompt if out.prompt is not None else ctx.tokenizer.decode(out.prompt_token_ids)) if req.echo else ""

        if req.stream:
            async def stream() -> AsyncGenerator[str, None]:
                ctx.inflight += 1
                n_prompt = n_out = 0
                try:
                    first = {}
                    merged = _merge(gens)
                    async for pi, out in merged:
                        if isinstance(out, Exception):
                            yield f"data: {json.dumps(P.ErrorResponse(error=P.ErrorInfo(message=str(out), code=400)).model_dump())}\n\n"
                            break
                        for c in out.outputs:
                            idx = pi * req.n + c.index
                            text = c.text
                            if req.echo and idx not in first:
                                text = echo_text(out) + text
                            first[idx] = True
                            lp = _completion_logprobs(ctx, c.token_ids, c.logprobs, req.logprobs) if (
                                req.logprobs is not None and c.logprobs) else None
                            chunk = P.CompletionStreamResponse(id=rid_base, created=created, model=ctx.model_name, choices=[
                                P.CompletionChoice(index=idx, text=text, logprobs=lp, finish_reason=c.finish_reason,
                                                   stop_reason=c.stop_reason)])
                            n_out += len(c.token_ids)
                            yield f"data: {chunk.model_dump_json(exclude_unset=False, exclude_none=True)}\n\n"
                        if out.finished:
                            n_prompt += len(out.prompt_token_ids or [])
                    if req.stream_options and req.stream_options.include_usage:
                        final = P.CompletionStreamResponse(id=rid_base, created=created, model=ctx.model_name, choices=[],
                                                           usage=_usage(n_prompt, n_out))
                        yield f"data: {final.model_dump_json(exclude_none=True)}\n\n"
                    yield "data: [DONE]\n\n"
                finally:
                    ctx.inflight -= 1
            return StreamingResponse(stream(), media_type="text/event-stream")

        ctx.inflight += 1
        try:
            finals = await asyncio.gather(*[_last(g, raw) for g in gens])
        except ValueError as ex:
            return _error(str(ex))
        except EngineDeadError:
            return _error("engine is dead", 500, "InternalServerError")
        finally:
            ctx.inflight -= 1
        if any(f is None for f in finals):
            return _error("client disconnected", 499)
        choices, n_prompt, n_out = [], 0, 0
        for pi, out in enumerate(finals):
            n_prompt += len(out.prompt_token_ids or [])
            for c in out.outputs:
                lp = None
                if req.logprobs is not None and c.logprobs:
                    toks, lps = list(c.token_ids), list(c.logprobs)
                    if req.echo and out.prompt_logprobs:
                        toks, lps = list(out.prompt_token_ids) + toks, list(out.prompt_logprobs) + lps
                    lp = _completion_logprobs(ctx, toks, lps, req.logprobs)
                choices.append(P.CompletionChoice(index=pi * req.n + c.index, text=echo_text(out) + c.text, logprobs=lp,
                                                  finish_reason=c.finish_reason, stop_reason=c.stop_reason,
                                                  prompt_logprobs=_plp(out.prompt_logprobs) if req.prompt_logprobs is not None else None))
                n_out += len(c.token_ids)
        return P.CompletionResponse(id=rid_base, created=created, model=ctx.model_name, choices=choices,
                                    usage=_usage(n_prompt, n_out, finals[0].num_cached_tokens),
                                    kv_transfer_params=finals[0].kv_transfer_params).model_dump(exclude_none=True)

    # -------------------------------------------------------------- chat
    @app.post("/v1/chat/completions")
    async def chat(req: P.ChatCompletionRequest, raw: Request):
        if (e := check_model(req.model)) is not None:
            return e
        try:
            prompt = render_chat(ctx.tokenizer, req.messages, req.chat_template or ctx.chat_template,
                                 req.add_generation_prompt, req.continue_final_message, req.tools,
                                 **(req.chat_template_kwargs or {}))
        except Exception as ex:
            return _error(f"chat template error: {ex}")
        n_lp = req.top_logprobs if req.logprobs else None
        if req.logprobs and n_lp is None:
            n_lp = 0
        try:
            sp = req.to_sampling_params(ctx.max_model_len, n_lp if req.logprobs else None, req.prompt_logprobs)
            sp.max_tokens = req.max_completion_tokens or req.max_tokens  # None → fill the context
            if sp.max_tokens is None:
                sp.max_tokens = None
        except (ValueError, TypeError) as ex:
            return _error(str(ex))
        rid = f"chatcmpl-{req.request_id or random_uuid()}"
        tool_parser = reasoning = None
        if ctx.tool_parser and req.tools and req.tool_choice != "none":
            from vllm_b200.tool_parsers import get_tool_parser
            tool_parser = get_tool_parser(ctx.tool_parser)(ctx.tokenizer)
        if ctx.reasoning_parser and req.include_reasoning:
            from vllm_b200.reasoning import get_reasoning_parser
            reasoning = get_reasoning_parser(ctx.reasoning_parser)(ctx.tokenizer)
        try:
            if isinstance(prompt, str):
                from vllm_b200.entrypoints.chat_utils import chat_multi_modal_data
                mm = chat_multi_modal_data(req.messages)
                prompt = {"prompt": prompt, "multi_modal_data": mm} if mm else {"prompt": prompt}
            gen = ctx.engine.generate(prompt, sp, rid,
                                      priority=req.priority, trace_headers=dict(raw.headers),
                                      lora_request=ctx.lora_requests.get(req.model))
        except ValueError as ex:
            return _error(str(ex))
        created = int(time.time())

        if req.stream:
            async def stream() -> AsyncGenerator[str, None]:
                ctx.inflight += 1
                try:
                    sent_role: set[int] = set()
                    n_prompt = n_out = 0
                    states: dict[int, dict] = {}
                    async for out in gen:
                        for c in out.outputs:
                            if c.index not in sent_role:
                                sent_role.add(c.index)
                                ch = P.ChatCompletionStreamResponse(id=rid, created=created, model=ctx.model_name, choices=[
                                    P.ChatStreamChoice(index=c.index, delta=P.DeltaMessage(role="assistant", content=""))])
                                yield f"data: {ch.model_dump_json(exclude_none=True)}\n\n"
                            st = states.setdefault(c.index, {"text": "", "ids": []})
                            delta = P.DeltaMessage(content=c.text)
                            if reasoning is not None or tool_parser is not None:
                                prev = st["text"]
                                st["text"] += c.text
                                st["ids"] += list(c.token_ids)
                                dm = None
                                if reasoning is not None and not st.get("reasoning_done"):
                                    dm = reasoning.extract_streaming(prev, st["text"], c.text)
                                    if dm is not None and dm.get("done"):
                                        st["reasoning_done"] = True
                                    if dm is not None:
                                        delta = P.DeltaMessage(content=dm.get("content"),
                                                               reasoning_content=dm.get("reasoning_content"))
                                if tool_parser is not None and (reasoning is None or st.get("reasoning_done")):
                                    tm = tool_parser.extract_streaming(prev, st["text"], c.text)
                                    if tm is not None:
                                        delta = P.DeltaMessage(content=tm.get("content"), tool_calls=[
                                            P.ToolCall(**t) for t in tm.get("tool_calls", [])] or None)
                                    elif tool_parser.in_tool_call(st["text"]):
                                        delta = P.DeltaMessage()
                            lp = _chat_logprobs(ctx, c.token_ids, c.logprobs, req.top_logprobs or 0) if (
                                req.logprobs and c.logprobs) else None
                            fr = c.finish_reason
                            if fr is not None and tool_parser is not None and tool_parser.saw_tool_call(st["text"]):
                                fr = "tool_calls"
                            ch = P.ChatCompletionStreamResponse(id=rid, created=created, model=ctx.model_name, choices=[
                                P.ChatStreamChoice(index=c.index, delta=delta, logprobs=lp, finish_reason=fr,
                                                   stop_reason=c.stop_reason)])
                            n_out += len(c.token_ids)
                            yield f"data: {ch.model_dump_json(exclude_none=True)}\n\n"
                        if out.finished:
                            n_prompt = len(out.prompt_token_ids or [])
                    if req.stream_options is None or req.stream_options.include_usage:
                        if req.stream_options is not None:
                            fin = P.ChatCompletionStreamResponse(id=rid, created=created, model=ctx.model_name,
                                                                 choices=[], usage=_usage(n_prompt, n_out))
                            yield f"data: {fin.model_dump_json(exclude_none=True)}\n\n"
                    yield "data: [DONE]\n\n"
                except ValueError as ex:
                    yield f"data: {json.dumps(P.ErrorResponse(error=P.ErrorInfo(message=str(ex))).model_dump())}\n\n"
                    yield "data: [DONE]\n\n"
                finally:
                    ctx.inflight -= 1
            return StreamingResponse(stream(), media_type="text/event-stream")

        ctx.inflight += 1
        try:
            out = await _last(gen, raw)
        except ValueError as ex:
            return _error(str(ex))
        except EngineDeadError:
            return _error("engine is dead", 500, "InternalServerError")
        finally:
            ctx.inflight -= 1
        if out is None:
            return _error("client disconnected", 499)
        choices, n_out = [], 0
        for c in out.outputs:
            content, rc, tcs, fr = c.text, None, None, c.finish_reason
            if reasoning is not None:
                rc, content = reasoning.extract(content)
            if tool_parser is not None:
                parsed = tool_parser.extract(content or "")
                if parsed["tool_calls"]:
                    tcs = [P.ToolCall(**t) for t in parsed["tool_calls"]]
                    content = parsed["content"]
                    fr = "tool_calls"
            lp = _chat_logprobs(ctx, c.token_ids, c.logprobs, req.top_logprobs or 0) if (req.logprobs and c.logprobs) else None
            choices.append(P.ChatChoice(index=c.index, message=P.ChatMessage(role="assistant", content=content,
                                                                               reasoning_content=rc, tool_calls=tcs),
                                        logprobs=lp, finish_reason=fr, stop_reason=c.stop_reason))
            n_out += len(c.token_ids)
        return P.ChatCompletionResponse(id=rid, created=created, model=ctx.model_name, choices=choices,
                                        usage=_usage(len(out.prompt_token_ids or []), n_out, out.num_cached_tokens),
                                        prompt_logprobs=_plp(out.prompt_logprobs) if req.prompt_logprobs is not None else None,
                                        kv_transfer_params=out.kv_transfer_params).model_dump(exclude_none=True)

    # --------------------------------------------------------- embeddings
    @app.post("/v1/embeddings")
    async def embeddings(req: P.EmbeddingRequest):
        if (e := check_model(req.model)) is not None:
            return e
        prompts = _prompts(req.input)
        pp = PoolingParams(task="embed", dimensions=req.dimensions, normalize=req.normalize,
                           truncate_prompt_tokens=req.truncate_prompt_tokens)
        rid = f"embd-{random_uuid()}"

        async def one(i, p):
            async for o in ctx.engine.encode(p, pp, f"{rid}-{i}"):
                return o
        try:
            outs = await asyncio.gather(*[one(i, p) for i, p in enumerate(prompts)])
        except ValueError as ex:
            return _error(str(ex))
        data, n_tok = [], 0
        for i, o in enumerate(outs):
            vec = [float(x) for x in o.outputs.data.flatten().tolist()]
            emb: Any = vec
            if req.encoding_format == "base64":
                emb = base64.b64encode(struct.pack(f"<{len(vec)}f", *vec)).decode()
            data.append(P.EmbeddingData(index=i, embedding=emb))
            n_tok += len(o.prompt_token_ids)
        return P.EmbeddingResponse(id=rid, model=ctx.model_name, data=data, usage=_usage(n_tok, 0)).model_dump(exclude_none=True)

    @app.post("/score")
    @app.post("/v1/score")
    async def score(req: P.ScoreRequest):
        import torch
        a = [req.text_1] if isinstance(req.text_1, str) else list(req.text_1)
        b = [req.text_2] if isinstance(req.text_2, str) else list(req.text_2)
        if len(a) == 1:
            a = a * len(b)
        pp = PoolingParams(task="embed")
        rid = f"score-{random_uuid()}"

        async def one(tag, i, p):
            async for o in ctx.engine.encode(p, pp, f"{rid}-{tag}{i}"):
                return o
        ea = await asyncio.gather(*[one("a", i, p) for i, p in enumerate(a)])
        eb = await asyncio.gather(*[one("b", i, p) for i, p in enumerate(b)])
        data = [{"index": i, "object": "score",
                 "score": float(torch.nn.functional.cosine_similarity(x.outputs.data.float().flatten(),
                                                                      y.outputs.data.float().flatten(), dim=0))}
                for i, (x, y) in enumerate(zip(ea, eb))]
        return P.ScoreResponse(model=ctx.model_name, data=data,
                               usage=_usage(sum(len(o.prompt_token_ids) for o in ea + eb), 0)).model_dump()

    # ------------------------------------------- token-in / token-out and render-only fronts
    # ------------------------------------------------------------ speech to text
    # (reference: entrypoints/speech_to_text/{api_router,serving,speech_to_text}.py — multipart upload, 30-second
    #  chunking at quiet points, task tokens as the decoder prompt, json / text / verbose_json / srt / vtt, SSE)
    def _stt_prompt_ids(task: str, language: Optional[str], prompt_text: Optional[str]) -> list[int]:
        tok, cfg = ctx.tokenizer, ctx.engine.model_config.hf_config
        start = cfg.get("decoder_start_token_id", cfg.get("bos_token_id", 0))
        conv = getattr(tok, "convert_tokens_to_ids", None)
        unk = getattr(tok, "unk_token_id", None)

        def tid(t: str) -> Optional[int]:
            try:
                i = conv(t) if conv else None
            except Exception:
                return None
            return i if isinstance(i, int) and i >= 0 and i != unk else None
        sot = tid("<|startoftranscript|>")
        if sot is None:
            return [start]
        ids = []
        if prompt_text and tid("<|startofprev|>") is not None:
            ids += [tid("<|startofprev|>")] + tok.encode(" " + prompt_text.strip(), add_special_tokens=False)
        ids.append(sot)
        for t in (f"<|{language or 'en'}|>", f"<|{task}|>", "<|notimestamps|>"):
            if tid(t) is not None:
                ids.append(tid(t))
        return ids

    def _ts(sec: float, sep: str) -> str:
        ms = int(round(sec * 1000))
        return f"{ms // 3600000:02d}:{ms // 60000 % 60:02d}:{ms // 1000 % 60:02d}{sep}{ms % 1000:03d}"

    async def _speech_to_text(raw: Request, task: str):
        from vllm_b200.multimodal.audio import SAMPLE_RATE, load_audio, split_chunks
        if not ctx.engine.model_config.hf_config.get("is_encoder_decoder"):
            return _error("the served model does not support speech-to-text", 400)
        form = await raw.form()
        up = form.get("file")
        if up is None:
            return _error("multipart field `file` is required")
        err = check_model(form.get("model"))
        if err:
            return err
        try:
            wav = load_audio(await up.read())
        except Exception as ex:
            return _error(f"could not decode the audio file: {ex}")
        fmt = form.get("response_format", "json")
        if fmt not in ("json", "text", "verbose_json", "srt", "vtt"):
            return _error(f"unsupported response_format {fmt!r}")
        language = form.get("language")
        temperature = float(form.get("temperature", 0.0) or 0.0)
        stream = str(form.get("stream", "")).lower() in ("1", "true")
        max_tokens = max(1, ctx.max_model_len - 8)
        ids = _stt_prompt_ids(task, language, form.get("prompt"))
        chunks = split_chunks(wav)
        rid = f"{'transcribe' if task == 'transcribe' else 'translate'}-{random_uuid()}"
        sp = SamplingParams(temperature=temperature, max_tokens=min(max_tokens, ctx.max_model_len - len(ids) - 1),
                            seed=int(form["seed"]) if form.get("seed") else None)

        async def run_chunk(i: int, chunk, on_delta=None) -> tuple[str, int]:
            gen = ctx.engine.generate({"prompt_token_ids": list(ids), "multi_modal_data": {"audio": (chunk, SAMPLE_RATE)}},
                                      sp, f"{rid}-{i}")
            text, n, sent = "", 0, 0
            async for o in gen:
                text, n = o.outputs[0].text, len(o.outputs[0].token_ids)
                if on_delta is not None and len(text) > sent:
                    await on_delta(text[sent:])
                    sent = len(text)
            return text, n

        if stream:
            async def sse():
                q: asyncio.Queue = asyncio.Queue()

                async def push(d):
                    await q.put(d)

                async def producer():
                    for i, c in enumerate(chunks):
                        await run_chunk(i, c, push)
                    await q.put(None)
                task_ = asyncio.create_task(producer())
                try:
                    while True:
                        d = await q.get()
                        if d is None:
                            break
                        yield "data: " + json.dumps({"id": rid, "object": "transcription.chunk", "created": int(time.time()),
                                                      "model": ctx.model_name,
                                                      "choices": [{"index": 0, "delta": {"content": d}}]}) + "\n\n"
                finally:
                    task_.cancel()
                yield "data: [DONE]\n\n"
            return StreamingResponse(sse(), media_type="text/event-stream")

        results = await asyncio.gather(*[run_chunk(i, c) for i, c in enumerate(chunks)])
        text = "".join(t for t, _ in results).strip()
        if fmt == "text":
            return Response(content=text, media_type="text/plain")
        bounds, pos = [], 0.0
        for c in chunks:
            bounds.append((pos, pos + len(c) / SAMPLE_RATE))
            pos += len(c) / SAMPLE_RATE
        if fmt in ("srt", "vtt"):
            sep = "," if fmt == "srt" else "."
            lines = ["WEBVTT", ""] if fmt == "vtt" else []
            for i, ((a, b), (t, _)) in enumerate(zip(bounds, results)):
                if fmt == "srt":
                    lines.append(str(i + 1))
                lines += [f"{_ts(a, sep)} --> {_ts(b, sep)}", t.strip(), ""]
            return Response(content="\n".join(lines), media_type="text/plain")
        usage = {"type": "duration", "seconds": int(round(len(wav) / SAMPLE_RATE))}
        if fmt == "verbose_json":
            return {"task": task, "language": language or "en", "duration": len(wav) / SAMPLE_RATE, "text": text,
                    "segments": [{"id": i, "start": a, "end": b, "text": t.strip(), "tokens": n}
                                 for i, ((a, b), (t, n)) in enumerate(zip(bounds, results))], "usage": usage}
        return {"text": text, "usage": usage}

    @app.websocket("/v1/realtime")
    async def realtime(ws: WebSocket):
        """Streaming speech recognition over a WebSocket (reference: entrypoints/speech_to_text realtime API; event names
        follow the OpenAI Realtime transcription session): the client appends base64 PCM16 mono chunks
        (``input_audio_buffer.append``), commits (``input_audio_buffer.commit``) and receives
        ``conversation.item.input_audio_transcription.delta`` events followed by ``...completed``."""
        import numpy as np
        from vllm_b200.multimodal.audio import SAMPLE_RATE, resample, split_chunks
        await ws.accept()
        if not ctx.engine.model_config.hf_config.get("is_encoder_decoder"):
            await ws.send_json({"type": "error", "error": {"message": "the served model does not support speech-to-text"}})
            await ws.close()
            return
        sess = {"language": None, "prompt": None, "sample_rate": SAMPLE_RATE, "temperature": 0.0}
        buf: list[np.ndarray] = []
        n_items = 0
        await ws.send_json({"type": "transcription_session.created", "session": {"model": ctx.model_name, **sess}})
        try:
            while True:
                ev = await ws.receive_json()
                typ = ev.get("type")
                if typ in ("session.update", "transcription_session.update"):
                    s_ = ev.get("session") or {}
                    tr = s_.get("input_audio_transcription") or {}
                    sess["language"] = tr.get("language", sess["language"])
                    sess["prompt"] = tr.get("prompt", sess["prompt"])
                    sess["sample_rate"] = int(s_.get("input_audio_sample_rate", sess["sample_rate"]))
                    await ws.send_json({"type": "transcription_session.updated", "session": {"model": ctx.model_name, **sess}})
                elif typ == "input_audio_buffer.append":
                    pcm = np.frombuffer(base64.b64decode(ev.get("audio") or ""), dtype="<i2").astype(np.float32) / 32768.0
                    buf.append(pcm)
                elif typ == "input_audio_buffer.clear":
                    buf.clear()
                    await ws.send_json({"type": "input_audio_buffer.cleared"})
                elif typ == "input_audio_buffer.commit":
                    if not buf:
                        await ws.send_json({"type": "error", "error": {"message": "input audio buffer is empty"}})
                        continue
                    wav = np.concatenate(buf)
                    buf.clear()
                    if sess["sample_rate"] != SAMPLE_RATE:
                        wav = resample(wav, sess["sample_rate"], SAMPLE_RATE)
                    n_items += 1
                    item_id = f"item_{n_items}"
                    await ws.send_json({"type": "input_audio_buffer.committed", "item_id": item_id})
                    ids = _stt_prompt_ids("transcribe", sess["language"], sess["prompt"])
                    sp = SamplingParams(temperature=float(sess["temperature"]),
                                        max_tokens=max(1, min(ctx.max_model_len - 8, ctx.max_model_len - len(ids) - 1)))
                    full = ""
                    for ci, chunk in enumerate(split_chunks(wav)):
                        sent = 0
                        async for o in ctx.engine.generate({"prompt_token_ids": list(ids),
                                                            "multi_modal_data": {"audio": (chunk, SAMPLE_RATE)}},
                                                           sp, f"rt-{random_uuid()}-{ci}"):
                            text = o.outputs[0].text
                            if len(text) > sent:
                                await ws.send_json({"type": "conversation.item.input_audio_transcription.delta",
                                                    "item_id": item_id, "delta": text[sent:]})
                                sent = len(text)
                        full += text
                    await ws.send_json({"type": "conversation.item.input_audio_transcription.completed", "item_id": item_id,
                                        "transcript": full.strip()})
                else:
                    await ws.send_json({"type": "error", "error": {"message": f"unknown event type {typ!r}"}})
        except Exception as ex:  # WebSocketDisconnect or malformed frames end the session
            if type(ex).__name__ != "WebSocketDisconnect":
                logger.warning("realtime session ended: %s", ex)

    @app.post("/v1/audio/transcriptions")
    async def transcriptions(raw: Request):
        return await _speech_to_text(raw, "transcribe")

    @app.post("/v1/audio/translations")
    async def translations(raw: Request):
        return await _speech_to_text(raw, "translate")

    # (reference: entrypoints/generate "/inference/v1/generate", entrypoints/scale_out render/derender split — a
    #  tokenizer-less engine tier behind a separate rendering tier)
    @app.post("/inference/v1/generate")
    async def generate_tokens(body: dict, raw: Request):
        ids = body.get("token_ids") or body.get("prompt_token_ids")
        if not ids:
            return _error("token_ids is required")
        spd = dict(body.get("sampling_params") or {})
        spd.setdefault("max_tokens", 16)
        spd["detokenize"] = False
        try:
            sp = SamplingParams(**spd)
        except (TypeError, ValueError) as ex:
            return _error(str(ex))
        rid = f"gen-{body.get('request_id') or random_uuid()}"
        stream = bool(body.get("stream"))
        gen = ctx.engine.generate({"prompt_token_ids": list(ids)}, sp, rid, priority=int(body.get("priority", 0)))
        if stream:
            async def sse():
                sent = 0
                async for o in gen:
                    c = o.outputs[0]
                    new = c.token_ids[sent:] if len(c.token_ids) >= sent else c.token_ids
                    sent = len(c.token_ids)
                    yield "data: " + json.dumps({"request_id": rid, "token_ids": list(new),
                                                  "finish_reason": c.finish_reason}) + "\n\n"
                yield "data: [DONE]\n\n"
            return StreamingResponse(sse(), media_type="text/event-stream")
        last = None
        async for o in gen:
            last = o
        c = last.outputs[0]
        return {"request_id": rid, "token_ids": list(c.token_ids), "finish_reason": c.finish_reason,
                "num_prompt_tokens": len(ids), "num_cached_tokens": getattr(last, "num_cached_tokens", None)}

    @app.post("/v1/chat/completions/render")
    @app.post("/render")
    async def render(body: dict):
        """Chat messages (or a raw prompt) → the exact token ids the engine would see; no generation."""
        from vllm_b200.entrypoints.chat_utils import render_chat
        if body.get("messages") is not None:
            text = render_chat(ctx.tokenizer, body["messages"], ctx.chat_template,
                               add_generation_prompt=bool(body.get("add_generation_prompt", True)), tools=body.get("tools"))
        else:
            text = body.get("prompt", "")
        ids = ctx.tokenizer.encode(text)
        return {"prompt": text, "token_ids": list(ids), "count": len(ids), "max_model_len": ctx.max_model_len}

    @app.post("/derender")
    async def derender(body: dict):
        return {"text": ctx.tokenizer.decode(list(body.get("token_ids", [])),
                                             skip_special_tokens=bool(body.get("skip_special_tokens", True)))}

    # ------------------------------------------- pooling family (reference: entrypoints/pooling/*)
    async def _encode_many(inputs: list, pp: PoolingParams, tag: str):
        rid = f"{tag}-{random_uuid()}"

        async def one(i, p):
            async for o in ctx.engine.encode({"prompt": p} if isinstance(p, str) else {"prompt_token_ids": p}, pp,
                                              f"{rid}-{i}"):
                return o
        return await asyncio.gather(*[one(i, p) for i, p in enumerate(inputs)])

    @app.post("/pooling")
    async def pooling(body: dict):
        inp = body.get("input")
        inputs = [inp] if isinstance(inp, str) or (inp and isinstance(inp[0], int)) else list(inp)
        pp = PoolingParams(task=body.get("task", "embed"), normalize=bool(body.get("normalize", False)),
                           pooling_type=body.get("pooling_type"))
        outs = await _encode_many(inputs, pp, "pool")
        return {"id": f"pool-{random_uuid()}", "object": "list", "model": ctx.model_name,
                "data": [{"index": i, "object": "pooling", "data": o.outputs.data.float().tolist()} for i, o in enumerate(outs)],
                "usage": _usage(sum(len(o.prompt_token_ids) for o in outs), 0).model_dump()}

    @app.post("/classify")
    async def classify(body: dict):
        inp = body.get("input")
        inputs = [inp] if isinstance(inp, str) else list(inp)
        outs = await _encode_many(inputs, PoolingParams(task="classify", normalize=False), "cls")
        data = []
        for i, o in enumerate(outs):
            probs = o.outputs.data.float().flatten()
            data.append({"index": i, "label": str(int(probs.argmax())), "probs": probs.tolist(), "num_classes": probs.numel()})
        return {"id": f"classify-{random_uuid()}", "object": "list", "model": ctx.model_name, "data": data,
                "usage": _usage(sum(len(o.prompt_token_ids) for o in outs), 0).model_dump()}

    @app.post("/rerank")
    @app.post("/v1/rerank")
    @app.post("/v2/rerank")
    async def rerank(body: dict):
        """Jina / Cohere style: rank ``documents`` against ``query`` (bi-encoder cosine similarity)."""
        import torch
        docs = [d if isinstance(d, str) else d.get("text", "") for d in body["documents"]]
        outs = await _encode_many([body["query"]] + docs, PoolingParams(task="embed"), "rerank")
        q = outs[0].outputs.data.float().flatten()
        scored = sorted(((float(torch.nn.functional.cosine_similarity(q, o.outputs.data.float().flatten(), dim=0)), i)
                         for i, o in enumerate(outs[1:])), reverse=True)
        top = scored[: int(body.get("top_n") or len(scored))]
        return {"id": f"rerank-{random_uuid()}", "model": ctx.model_name,
                "results": [{"index": i, "relevance_score": sc, "document": {"text": docs[i]}} for sc, i in top],
                "usage": {"total_tokens": sum(len(o.prompt_token_ids) for o in outs)}}

    # ------------------------------------------- Responses API / Anthropic
    async def _chat_deltas(creq, raw):
        """Runs the chat handler in streaming mode and yields (delta_text, finish_reason, usage) tuples."""
        creq.stream = True
        creq.stream_options = P.StreamOptions(include_usage=True)
        res = await chat(creq, raw)
        if not isinstance(res, StreamingResponse):
            yield res
            return
        async for chunk in res.body_iterator:
            text = chunk.decode() if isinstance(chunk, (bytes, bytearray)) else chunk
            for line in text.splitlines():
                if not line.startswith("data:") or "[DONE]" in line:
                    continue
                d = json.loads(line[5:])
                ch = d["choices"][0] if d.get("choices") else {}
                yield (ch.get("delta", {}).get("content") or "", ch.get("finish_reason"), d.get("usage"))

    def _sse_event(typ: str, payload: dict) -> str:
        return f"event: {typ}\ndata: {json.dumps(dict(payload, type=typ))}\n\n"

    @app.get("/v1/responses/{response_id}")
    async def get_response(response_id: str):
        st = ctx.responses_store.get(response_id)
        if st is None or "response" not in st:
            return _error(f"response {response_id} not found", 404, "NotFoundError")
        return st["response"]

    @app.post("/v1/responses/{response_id}/cancel")
    async def cancel_response(response_id: str):
        st = ctx.responses_store.get(response_id)
        if st is None:
            return _error(f"response {response_id} not found", 404, "NotFoundError")
        if st.get("request_ids"):
            await ctx.engine.abort(st["request_ids"])
        if "response" in st:
            st["response"]["status"] = "cancelled"
        return st.get("response", {"id": response_id, "status": "cancelled"})

    async def _responses_with_tools(req, raw, msgs, rid):
        """Function tools are handed to the model and its calls returned as ``function_call`` items for the client to
        run; MCP tools (request ``{"type": "mcp"}`` entries and ``--tool-server``) are run here, their results fed
        back, until the model answers (reference: openai/responses/serving.py tool loop)."""
        from vllm_b200.entrypoints.tool_server import MCPError, ToolServer, run_tool_loop
        ts = ctx.tool_server or ToolServer()
        try:
            fns, route, items = await ts.resolve(req.tools)
        except (MCPError, OSError, KeyError) as e:
            return _error(f"tool server: {e}", 424, "FailedDependencyError")
        failure = []

        async def chat_once(messages, tools):
            c = P.ChatCompletionRequest(messages=messages, max_tokens=req.max_output_tokens, temperature=req.temperature,
                                        top_p=req.top_p, stream=False, tools=tools or None,
                                        tool_choice=(req.tool_choice or "auto") if tools else "none")
            res = await chat(c, raw)
            if not isinstance(res, dict):
                failure.append(res)
                return {"message": {"content": ""}, "usage": None}
            return {"message": res["choices"][0]["message"], "usage": res["usage"]}
        created = int(time.time())
        msg, calls, u = await run_tool_loop(chat_once, msgs, fns, route, max_turns=req.max_tool_calls or 8)
        if failure:
            return failure[0]
        output = items + calls
        for c in msg.get("tool_calls") or []:
            output.append({"type": "function_call", "id": f"fc_{random_uuid()}", "call_id": c["id"], "status": "completed",
                           "name": c["function"]["name"], "arguments": c["function"].get("arguments") or "{}"})
        text = msg.get("content") or ""
        if msg.get("reasoning_content"):
            output.append({"type": "reasoning", "id": f"rs_{random_uuid()}", "summary": [],
                           "content": [{"type": "reasoning_text", "text": msg["reasoning_content"]}]})
        if text or not msg.get("tool_calls"):
            output.append({"type": "message", "id": f"msg_{random_uuid()}", "role": "assistant", "status": "completed",
                           "content": [{"type": "output_text", "text": text, "annotations": []}]})
        out = {"id": rid, "object": "response", "created_at": created, "status": "completed", "model": ctx.model_name,
               "output": output, "output_text": text, "tools": req.tools or [],
               "usage": {"input_tokens": u["prompt_tokens"], "output_tokens": u["completion_tokens"],
                         "total_tokens": u["total_tokens"]}}
        if req.store:
            hist = msgs + [{"role": "assistant", "content": text, **({"tool_calls": msg["tool_calls"]}
                                                                       if msg.get("tool_calls") else {})}]
            ctx.responses_store[rid] = {"messages": hist, "response": out}
        if not req.stream:
            return out

        async def events():
            seq = 0

            def ev(typ, **kw):
                nonlocal seq
                seq += 1
                return _sse_event(typ, dict(kw, sequence_number=seq))
            shell = dict(out, status="in_progress", output=[], output_text="")
            yield ev("response.created", response=shell)
            yield ev("response.in_progress", response=shell)
            for i, item in enumerate(output):
                yield ev("response.output_item.added", output_index=i, item=item)
                if item["type"] == "mcp_call":
                    yield ev("response.mcp_call.completed" if not item["error"] else "response.mcp_call.failed",
                             output_index=i, item_id=item["id"])
                elif item["type"] == "function_call":
                    yield ev("response.function_call_arguments.done", output_index=i, item_id=item["id"],
                             arguments=item["arguments"])
                elif item["type"] == "message":
                    yield ev("response.output_text.delta", item_id=item["id"], output_index=i, content_index=0, delta=text)
                    yield ev("response.output_text.done", item_id=item["id"], output_index=i, content_index=0, text=text)
                yield ev("response.output_item.done", output_index=i, item=item)
            yield ev("response.completed", response=out)
        return StreamingResponse(events(), media_type="text/event-stream")

    @app.post("/v1/responses")
    async def responses(req: P.ResponsesRequest, raw: Request):
        msgs: list[dict] = []
        if req.previous_response_id and req.previous_response_id in ctx.responses_store:
            msgs += ctx.responses_store[req.previous_response_id]["messages"]
        if req.instructions:
            msgs.append({"role": "system", "content": req.instructions})
        if isinstance(req.input, str):
            msgs.append({"role": "user", "content": req.input})
        else:
            from vllm_b200.entrypoints.tool_server import input_items_to_messages
            msgs += input_items_to_messages(req.input)
        creq = P.ChatCompletionRequest(messages=msgs, max_tokens=req.max_output_tokens, temperature=req.temperature,
                                       top_p=req.top_p, stream=False)
        rid = f"resp_{random_uuid()}"
        if req.tools or ctx.tool_server is not None:
            return await _responses_with_tools(req, raw, msgs, rid)

        def build(text: str, usage: Optional[dict], status: str, created: int, mid: str) -> dict:
            u = usage or {"prompt_tokens": 0, "completion_tokens": 0, "total_tokens": 0}
            return {"id": rid, "object": "response", "created_at": created, "status": status, "model": ctx.model_name,
                    "output": [{"type": "message", "id": mid, "role": "assistant", "status": status,
                                "content": [{"type": "output_text", "text": text, "annotations": []}]}],
                    "output_text": text,
                    "usage": {"input_tokens": u["prompt_tokens"], "output_tokens": u["completion_tokens"],
                              "total_tokens": u["total_tokens"]}}

        if req.stream:
            # typed server-sent events (reference: openai/responses/streaming_events.py)
            async def events():
                created, mid, seq, text, usage = int(time.time()), f"msg_{random_uuid()}", 0, "", None

                def ev(typ, **kw):
                    nonlocal seq
                    seq += 1
                    return _sse_event(typ, dict(kw, sequence_number=seq))
                shell = build("", None, "in_progress", created, mid)
                shell["output"] = []
                yield ev("response.created", response=shell)
                yield ev("response.in_progress", response=shell)
                yield ev("response.output_item.added", output_index=0,
                         item={"type": "message", "id": mid, "role": "assistant", "status": "in_progress", "content": []})
                yield ev("response.content_part.added", item_id=mid, output_index=0, content_index=0,
                         part={"type": "output_text", "text": "", "annotations": []})
                async for item in _chat_deltas(creq, raw):
                    if not isinstance(item, tuple):
                        yield ev("error", message="generation failed")
                        return
                    delta, _fin, u = item
                    usage = u or usage
                    if delta:
                        text += delta
                        yield ev("response.output_text.delta", item_id=mid, output_index=0, content_index=0, delta=delta)
                yield ev("response.output_text.done", item_id=mid, output_index=0, content_index=0, text=text)
                yield ev("response.content_part.done", item_id=mid, output_index=0, content_index=0,
                         part={"type": "output_text", "text": text, "annotations": []})
                final = build(text, usage, "completed", created, mid)
                yield ev("response.output_item.done", output_index=0, item=final["output"][0])
                if req.store:
                    ctx.responses_store[rid] = {"messages": msgs + [{"role": "assistant", "content": text}],
                                                "response": final}
                yield ev("response.completed", response=final)
            return StreamingResponse(events(), media_type="text/event-stream")

        res = await chat(creq, raw)
        if not isinstance(res, dict):
            return res
        text = res["choices"][0]["message"].get("content") or ""
        out = build(text, res["usage"], "completed", res["created"], f"msg_{random_uuid()}")
        if req.store:
            ctx.responses_store[rid] = {"messages": msgs + [{"role": "assistant", "content": text}], "response": out}
        return out

    @app.post("/v1/messages")
    async def anthropic_messages(req: P.AnthropicMessagesRequest, raw: Request):
        msgs: list[dict] = []
        if req.system:
            sys_text = req.system if isinstance(req.system, str) else "".join(b.get("text", "") for b in req.system)
            msgs.append({"role": "system", "content": sys_text})
        for m in req.messages:
            c = m.get("content")
            if isinstance(c, list):
                c = "".join(b.get("text", "") for b in c if isinstance(b, dict) and b.get("type") == "text")
            msgs.append({"role": m["role"], "content": c})
        creq = P.ChatCompletionRequest(messages=msgs, max_tokens=req.max_tokens, temperature=req.temperature,
                                       top_p=req.top_p, top_k=req.top_k, stop=req.stop_sequences, stream=False)
        stop_map = {"stop": "end_turn", "length": "max_tokens", "tool_calls": "tool_use"}
        if getattr(req, "stream", False):
            async def events():
                mid = f"msg_{random_uuid()}"
                yield _sse_event("message_start", {"message": {
                    "id": mid, "type": "message", "role": "assistant", "model": ctx.model_name, "content": [],
                    "stop_reason": None, "stop_sequence": None, "usage": {"input_tokens": 0, "output_tokens": 0}}})
                yield _sse_event("content_block_start", {"index": 0, "content_block": {"type": "text", "text": ""}})
                fin, usage = None, None
                async for item in _chat_deltas(creq, raw):
                    if not isinstance(item, tuple):
                        yield _sse_event("error", {"error": {"type": "api_error", "message": "generation failed"}})
                        return
                    delta, f, u = item
                    fin, usage = f or fin, u or usage
                    if delta:
                        yield _sse_event("content_block_delta", {"index": 0, "delta": {"type": "text_delta", "text": delta}})
                yield _sse_event("content_block_stop", {"index": 0})
                yield _sse_event("message_delta", {"delta": {"stop_reason": stop_map.get(fin, "end_turn"), "stop_sequence": None},
                                                   "usage": {"input_tokens": (usage or {}).get("prompt_tokens", 0),
                                                             "output_tokens": (usage or {}).get("completion_tokens", 0)}})
                yield _sse_event("message_stop", {})
            return StreamingResponse(events(), media_type="text/event-stream")
        res = await chat(creq, raw)
        if not isinstance(res, dict):
            return res
        ch = res["choices"][0]
        return {"id": f"msg_{random_uuid()}", "type": "message", "role": "assistant", "model": ctx.model_name,
                "content": [{"type": "text", "text": ch["message"].get("content") or ""}],
                "stop_reason": stop_map.get(ch.get("finish_reason"), "end_turn"), "stop_sequence": None,
                "usage": {"input_tokens": res["usage"]["prompt_tokens"],
                          "output_tokens": res["usage"]["completion_tokens"]}}

    @app.post("/v2/chat")
    async def cohere_chat(body: dict, raw: Request):
        """Cohere v2 chat shape over the chat handler (reference: entrypoints/cohere/)."""
        msgs = []
        for m in body.get("messages") or []:
            c = m.get("content")
            if isinstance(c, list):
                c = "".join(p.get("text", "") for p in c if isinstance(p, dict))
            msgs.append({"role": m.get("role", "user"), "content": c})
        if not msgs:
            return _error("messages is required")
        creq = P.ChatCompletionRequest(messages=msgs, max_tokens=body.get("max_tokens"), temperature=body.get("temperature"),
                                       top_p=body.get("p"), top_k=body.get("k"), stop=body.get("stop_sequences"),
                                       seed=body.get("seed"), stream=False)
        res = await chat(creq, raw)
        if not isinstance(res, dict):
            return res
        ch = res["choices"][0]
        fin = {"stop": "COMPLETE", "length": "MAX_TOKENS", "tool_calls": "TOOL_CALL"}.get(ch.get("finish_reason"), "COMPLETE")
        u = res["usage"]
        return {"id": random_uuid(), "finish_reason": fin,
                "message": {"role": "assistant", "content": [{"type": "text", "text": ch["message"].get("content") or ""}]},
                "usage": {"billed_units": {"input_tokens": u["prompt_tokens"], "output_tokens": u["completion_tokens"]},
                          "tokens": {"input_tokens": u["prompt_tokens"], "output_tokens": u["completion_tokens"]}}}

    # ---------------------------------------- runtime LoRA management (opt-in: VLLM_ALLOW_RUNTIME_LORA_UPDATING=1)
    if envs.VLLM_ALLOW_RUNTIME_LORA_UPDATING:
        @app.post("/v1/load_lora_adapter")
        async def load_lora(body: dict):
            from vllm_b200.core.request import LoRARequest
            name, path = body["lora_name"], body["lora_path"]
            lid = ctx.lora_ids.setdefault(name, len(ctx.lora_ids) + 1)
            ctx.lora_requests[name] = LoRARequest(name, lid, path)
            await ctx.engine.collective_rpc("add_lora", (), {"lora_request": ctx.lora_requests[name]})
            return {"status": "ok", "lora_int_id": lid}

        @app.post("/v1/unload_lora_adapter")
        async def unload_lora(body: dict):
            req = ctx.lora_requests.pop(body["lora_name"], None)
            if req is None:
                return JSONResponse({"error": "unknown adapter"}, status_code=404)
            await ctx.engine.collective_rpc("remove_lora", (req.lora_int_id,))
            return {"status": "ok"}

    # -------------------------------------------------------- dev routes (opt-in: VLLM_SERVER_DEV_MODE=1, as in the
    # reference's entrypoints/serve/*: never mounted on a production server)
    if not envs.VLLM_SERVER_DEV_MODE:
        return app

    @app.post("/start_profile")
    async def start_profile():
        await ctx.engine.start_profile()
        return Response(status_code=200)

    @app.post("/stop_profile")
    async def stop_profile():
        await ctx.engine.stop_profile()
        return Response(status_code=200)

    @app.post("/sleep")
    async def sleep(level: int = 1):
        await ctx.engine.sleep(level)
        return Response(status_code=200)

    @app.post("/wake_up")
    async def wake_up():
        await ctx.engine.wake_up()
        return Response(status_code=200)

    @app.get("/is_sleeping")
    async def is_sleeping():
        return {"is_sleeping": await ctx.engine.is_sleeping()}

    @app.post("/reset_prefix_cache")
    async def reset_prefix_cache():
        ok = await ctx.engine.reset_prefix_cache()
        return Response(status_code=200 if ok else 409)

    # fault tolerance / LoRA management / checkpoint hooks (reference: entrypoints/serve/{fault_tolerance,lora,...})
    @app.get("/fault_tolerance/status")
    async def ft_status():
        return await ctx.engine.get_fault_status()

    @app.post("/fault_tolerance/apply")
    async def ft_apply(body: dict):
        ok = await ctx.engine.handle_fault(body.get("instruction", "retry"))
        return {"ok": bool(ok)}

    @app.post("/scale_elastic_ep")
    async def scale_elastic_ep(body: dict):
        try:
            n = await ctx.engine.scale_elastic_ep(int(body["new_data_parallel_size"]))
        except (RuntimeError, ValueError, KeyError) as ex:
            return _error(str(ex))
        return {"data_parallel_size": n}

    @app.post("/checkpoint/prepare")
    async def ckpt_prepare():
        await ctx.engine.collective_rpc("checkpoint_prepare")
        return Response(status_code=200)

    @app.post("/checkpoint/restore")
    async def ckpt_restore():
        await ctx.engine.collective_rpc("checkpoint_restore")
        return Response(status_code=200)

    @app.post("/collective_rpc")
    async def collective_rpc(body: dict):
        method = body.get("method")
        if method not in _RPC_ALLOWED and method not in envs.VLLM_SERVER_DEV_RPC_METHODS.split(","):
            return _error(f"collective_rpc method {method!r} is not on the allow-list", 403, "Forbidden")
        res = await ctx.engine.collective_rpc(body["method"], tuple(body.get("args", [])), body.get("kwargs"))
        return {"results": [r if isinstance(r, (int, float, str, bool, type(None))) else repr(r) for r in (res or [])]}

    return app


def _plp(plp):
    if plp is None:
        return None
    return [None if d is None else {t: {"logprob": max(v.logprob, -9999.0), "rank": v.rank,
                                        "decoded_token": v.decoded_token} for t, v in d.items()} for d in plp]


async def _last(gen, raw=None):
    """Drain a generator, aborting if the client goes away (reference: with_cancellation)."""
    out = None
    async for o in gen:
        out = o
        if raw is not None and getattr(raw, "is_disconnected", None) is not None and o is not None and not o.finished:
            try:
                if await raw.is_disconnected():
                    await gen.aclose()
                    return None
            except Exception:
                pass
    return out


async def _merge(gens):
    """Interleave several async generators, yielding (index, item)."""
    q: asyncio.Queue = asyncio.Queue()
    done = object()

    async def pump(i, g):
        try:
            async for item in g:
                await q.put((i, item))
        except Exception as e:  # surfaced to the consumer
            await q.put((i, e))
        finally:
            await q.put((i, done))

    tasks = [asyncio.create_task(pump(i, g)) for i, g in enumerate(gens)]
    remaining = len(tasks)
    try:
        while remaining:
            i, item = await q.get()
            if item is done:
                remaining -= 1
                continue
            yield i, item
    finally:
        for t in tasks:
            t.cancel()


# ---------------------------------------------------------------------------


def make_arg_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser("vllm_b200 serve")
    ap.add_argument("model_tag", nargs="?", default=None)
    ap.add_argument("--host", default="0.0.0.0")
    ap.add_argument("--port", type=int, default=8000)
    ap.add_argument("--api-key", default=None)
    ap.add_argument("--chat-template", default=None)
    ap.add_argument("--enable-auto-tool-choice", action="store_true")
    ap.add_argument("--tool-call-parser", default=None)
    ap.add_argument("--tool-server", default=None,
                    help="comma-separated MCP server URLs whose tools every Responses-API request may use")
    ap.add_argument("--uvicorn-log-level", default="warning")
    ap.add_argument("--api-server-count", type=int, default=1)
    ap.add_argument("--headless", action="store_true", help="engine cores only; connect to a remote front-end")
    ap.add_argument("--native-frontend", action="store_true", help="serve through the C++ front-end daemon")
    ap.add_argument("--grpc-port", type=int, default=None, help="also serve the gRPC engine service on this port")
    ap.add_argument("--enable-log-requests", action="store_true", help="log one line per API request (no payloads)")
    ap.add_argument("--engine-input-addr", default=None)
    ap.add_argument("--engine-output-addr", default=None)
    ap.add_argument("--data-parallel-start-rank", type=int, default=None)
    EngineArgs.add_cli_args(ap)
    return ap


async def run_server(args: argparse.Namespace) -> None:
    import uvicorn
    from vllm_b200.engine.async_llm import AsyncLLM
    if args.model_tag:
        args.model = args.model_tag
    ea = EngineArgs.from_cli_args(args)
    engine = AsyncLLM.from_engine_args(ea)
    from vllm_b200.entrypoints.chat_templates import resolve_chat_template
    template = resolve_chat_template(args.chat_template)        # built-in name, file, or jinja source
    try:
        engine.tokenizer.model_type = engine.model_config.hf_config.get("model_type")    # family fallback template
    except Exception:
        pass
    ctx = ServingContext(engine, ea.served_model_name or ea.model, engine.model_config.max_model_len, engine.tokenizer,
                         api_key=args.api_key, tool_parser=args.tool_call_parser,
                         reasoning_parser=ea.reasoning_parser or None, enable_auto_tools=args.enable_auto_tool_choice,
                         chat_template=template)
    ctx.log_requests = bool(getattr(args, "enable_log_requests", False))
    if getattr(args, "tool_server", None):
        from vllm_b200.entrypoints.tool_server import ToolServer
        ctx.tool_server = ToolServer([u for u in args.tool_server.split(",") if u])
    app = build_app(ctx)
    config = uvicorn.Config(app, host=args.host, port=args.port, log_level=args.uvicorn_log_level)
    server = uvicorn.Server(config)

    async def watchdog() -> None:  # terminate the HTTP server when the engine dies (reference launcher.py:180-203)
        while not server.should_exit:
            await asyncio.sleep(2)
            if engine.errored:
                logger.error("engine died; shutting the API server down")
                server.should_exit = True

    loop = asyncio.get_running_loop()
    for sig in (signal.SIGINT, signal.SIGTERM):
        try:
            loop.add_signal_handler(sig, lambda: setattr(server, "should_exit", True))
        except NotImplementedError:
            pass
    grpc_server = None
    if getattr(args, "grpc_port", None):
        from vllm_b200.entrypoints.grpc import start_grpc_server
        grpc_server, _ = await start_grpc_server(engine, ctx.model_name, args.host, args.grpc_port)
    wd = asyncio.create_task(watchdog())
    listen_fd = os.environ.get("VLLM_B200_LISTEN_FD")       # API server k of N: the parent owns the listening socket
    sockets = [socket.socket(fileno=int(listen_fd))] if listen_fd else None
    try:
        await server.serve(sockets=sockets)
    finally:
        wd.cancel()
        if grpc_server is not None:
            await grpc_server.stop(1)
        engine.shutdown()


def run_multi_api_server(args: argparse.Namespace, argv: list[str]) -> None:
    """``--api-server-count N``: N HTTP front-end processes accept on one shared listening socket and share the
    engine cores (each front-end is ``client_index`` k with its own ZMQ socket pair); the parent launches the
    engines and supervises (reference: entrypoints/cli/serve.py run_multi_api_server, v1/utils.py
    APIServerProcessManager)."""
    import subprocess
    import sys
    from vllm_b200.engine.core_client import _launch_engines
    from vllm_b200.utils import get_open_zmq_ipc_path
    if args.model_tag:
        args.model = args.model_tag
    ea = EngineArgs.from_cli_args(args)
    cfg = ea.create_engine_config()
    n = args.api_server_count
    addrs = [(get_open_zmq_ipc_path(), get_open_zmq_ipc_path()) for _ in range(n)]
    lsock = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
    lsock.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
    lsock.bind((args.host, args.port))
    lsock.listen(2048)
    lsock.set_inheritable(True)
    child_argv = [a for a in (argv if argv is not None else sys.argv[2:])]
    servers = []
    for k, (ia, oa) in enumerate(addrs):
        env = dict(os.environ, VLLM_B200_EXTERNAL_ENGINE=f"{k}|{ia}|{oa}", VLLM_B200_LISTEN_FD=str(lsock.fileno()),
                   VLLM_B200_API_SERVER_INDEX=str(k))
        servers.append(subprocess.Popen([sys.executable, "-m", "vllm_b200", "serve", *child_argv,
                                         "--api-server-count", "1"], env=env, pass_fds=[lsock.fileno()]))
    engines = _launch_engines(cfg, ",".join(a for a, _ in addrs), ",".join(b for _, b in addrs),
                              cfg.parallel_config.data_parallel_size, log_stats=not ea.disable_log_stats)
    stop = {"v": False}
    for sig in (signal.SIGINT, signal.SIGTERM):
        signal.signal(sig, lambda *a: stop.__setitem__("v", True))
    try:
        while not stop["v"]:
            if any(p.poll() is not None for p in servers) or any(not e.is_alive() for e in engines):
                logger.error("a front-end or engine process exited; shutting down")
                break
            time.sleep(0.5)
    finally:
        for p in servers:
            if p.poll() is None:
                p.terminate()
        for e in engines:
            if e.is_alive():
                e.terminate()
        for p in servers:
            try:
                p.wait(timeout=20)
            except subprocess.TimeoutExpired:
                p.kill()
        for e in engines:
            e.join(timeout=20)
        lsock.close()


def main(argv=None) -> None:
    args = make_arg_parser().parse_args(argv)
    if args.headless:
        from vllm_b200.entrypoints.native_frontend import run_headless
        return run_headless(args)
    if args.native_frontend:
        from vllm_b200.entrypoints.native_frontend import run_native_server
        return run_native_server(args)
    if args.api_server_count > 1:
        return run_multi_api_server(args, argv)
    try:
        import uvloop
        uvloop.run(run_server(args))
    except ImportError:
        asyncio.run(run_server(args))


if __name__ == "__main__":
    main()
