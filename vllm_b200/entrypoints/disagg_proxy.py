"""Disaggregated-serving proxy: one front door for a prefill pool and a decode pool.

Behavioural reference: tests/v1/kv_connector/nixl_integration/toy_proxy_server.py and
examples/online_serving/disaggregated_serving/ (round-robin over prefill / decode instances; the prefill call is the
user's request with ``max_tokens=1`` and ``kv_transfer_params={"do_remote_decode": true}``; its
``kv_transfer_params`` answer rides on the request sent to the decode instance, whose (streamed) response goes back
to the user untouched).

    python -m vllm_b200.entrypoints.disagg_proxy --prefill http://p0:8100 --decode http://d0:8200 --port 8000
"""
from __future__ import annotations

import argparse
import itertools
import json
from typing import Any

import aiohttp
from fastapi import FastAPI, Request
from fastapi.responses import JSONResponse, StreamingResponse


def build_proxy(prefill_urls: list[str], decode_urls: list[str]) -> FastAPI:
    app = FastAPI()
    p_iter, d_iter = itertools.cycle(prefill_urls), itertools.cycle(decode_urls)
    app.state.stats = {"prefill": {u: 0 for u in prefill_urls}, "decode": {u: 0 for u in decode_urls}}

    async def _handle(path: str, body: dict[str, Any], headers: dict[str, str]):
        p_url, d_url = next(p_iter), next(d_iter)
        app.state.stats["prefill"][p_url] += 1
        app.state.stats["decode"][d_url] += 1
        pre = dict(body)
        pre.update(max_tokens=1, stream=False, kv_transfer_params={"do_remote_decode": True})
        pre.pop("max_completion_tokens", None)
        pre.pop("stream_options", None)
        timeout = aiohttp.ClientTimeout(total=None)
        async with aiohttp.ClientSession(timeout=timeout) as s:
            async with s.post(p_url + path, json=pre, headers=headers) as r:
                if r.status != 200:
                    return JSONResponse(await r.json(), status_code=r.status)
                kv_params = (await r.json()).get("kv_transfer_params")
        dec = dict(body)
        if kv_params:
            dec["kv_transfer_params"] = kv_params
        if body.get("stream"):
            async def gen():
                async with aiohttp.ClientSession(timeout=timeout) as s:
                    async with s.post(d_url + path, json=dec, headers=headers) as r:
                        async for chunk in r.content.iter_any():
                            yield chunk
            return StreamingResponse(gen(), media_type="text/event-stream")
        async with aiohttp.ClientSession(timeout=timeout) as s:
            async with s.post(d_url + path, json=dec, headers=headers) as r:
                return JSONResponse(await r.json(), status_code=r.status)

    def _fwd_headers(req: Request) -> dict[str, str]:
        return {k: v for k, v in req.headers.items() if k.lower() in ("authorization", "x-request-id")}

    @app.post("/v1/completions")
    async def completions(req: Request):
        return await _handle("/v1/completions", await req.json(), _fwd_headers(req))

    @app.post("/v1/chat/completions")
    async def chat(req: Request):
        return await _handle("/v1/chat/completions", await req.json(), _fwd_headers(req))

    @app.get("/v1/models")
    async def models():
        async with aiohttp.ClientSession() as s:
            async with s.get(decode_urls[0] + "/v1/models") as r:
                return JSONResponse(await r.json(), status_code=r.status)

    @app.get("/health")
    async def health():
        return JSONResponse({"status": "ok"})

    @app.get("/proxy/stats")
    async def stats():
        return JSONResponse(app.state.stats)

    return app


def main(argv=None) -> None:
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--prefill", nargs="+", required=True, help="base URLs of the prefill instances")
    ap.add_argument("--decode", nargs="+", required=True, help="base URLs of the decode instances")
    ap.add_argument("--host", default="0.0.0.0")
    ap.add_argument("--port", type=int, default=8000)
    a = ap.parse_args(argv)
    import uvicorn
    uvicorn.run(build_proxy(a.prefill, a.decode), host=a.host, port=a.port, log_level="warning")


if __name__ == "__main__":
    main()
