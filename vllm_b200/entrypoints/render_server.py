"""``vllm_b200 launch render <model>``: a GPU-less pre / post-processing server — chat templating, tokenisation and
detokenisation only, no engine.  Disaggregated front-ends run it next to (or instead of) the in-process renderer so
CPU-heavy prompt preparation scales independently of the GPU replicas.

Routes: ``POST /render`` and ``/v1/chat/completions/render`` (messages or prompt → text + token ids), ``POST /derender``
(token ids → text), ``POST /tokenize`` / ``/detokenize`` (OpenAI-compatible server forms), ``GET /health``, ``/version``,
``/v1/models``.  Reference: vllm/entrypoints/cli/launch.py (``RenderSubcommand``), entrypoints/launchers/render/entry.py.
"""
from __future__ import annotations

import argparse
from typing import Optional


def build_render_app(model: str, tokenizer: Optional[str] = None, chat_template: Optional[str] = None,
                     tokenizer_mode: str = "auto", trust_remote_code: bool = False, max_model_len: Optional[int] = None,
                     served_model_name: Optional[str] = None):
    from fastapi import FastAPI
    from vllm_b200.config.model import ModelConfig
    from vllm_b200.entrypoints.chat_templates import resolve_chat_template
    from vllm_b200.entrypoints.chat_utils import render_chat
    from vllm_b200.tokenizers import init_tokenizer_from_config
    from vllm_b200.version import __version__
    mc = ModelConfig(model=model, tokenizer=tokenizer, tokenizer_mode=tokenizer_mode, trust_remote_code=trust_remote_code,
                     max_model_len=max_model_len or 0, served_model_name=served_model_name)
    tok = init_tokenizer_from_config(mc)
    template = resolve_chat_template(chat_template)
    app = FastAPI(title="vllm_b200 render")

    @app.get("/health")
    async def health():
        return {}

    @app.get("/version")
    async def version():
        return {"version": __version__}

    @app.get("/v1/models")
    async def models():
        return {"object": "list", "data": [{"id": mc.served_model_name, "object": "model", "max_model_len": mc.max_model_len}]}

    def _render(body: dict):
        if body.get("messages") is not None:
            text = render_chat(tok, body["messages"], body.get("chat_template") or template,
                               add_generation_prompt=bool(body.get("add_generation_prompt", True)),
                               continue_final_message=bool(body.get("continue_final_message", False)), tools=body.get("tools"),
                               model_type=mc.hf_config.get("model_type"))
            ids = tok.encode(text, add_special_tokens=False)
        else:
            text = body.get("prompt", "")
            ids = tok.encode(text, add_special_tokens=bool(body.get("add_special_tokens", True)))
        return text, list(ids)

    @app.post("/render")
    @app.post("/v1/chat/completions/render")
    async def render(body: dict):
        text, ids = _render(body)
        return {"prompt": text, "token_ids": ids, "count": len(ids), "max_model_len": mc.max_model_len}

    @app.post("/tokenize")
    async def tokenize(body: dict):
        _, ids = _render(body)
        out = {"count": len(ids), "max_model_len": mc.max_model_len, "tokens": ids}
        if body.get("return_token_strs"):
            out["token_strs"] = tok.convert_ids_to_tokens(ids)
        return out

    @app.post("/derender")
    async def derender(body: dict):
        return {"text": tok.decode(list(body.get("token_ids", [])), skip_special_tokens=bool(body.get("skip_special_tokens", True)))}

    @app.post("/detokenize")
    async def detokenize(body: dict):
        return {"prompt": tok.decode(list(body.get("tokens", [])), skip_special_tokens=False)}
    return app


def main(argv: list[str]) -> None:
    ap = argparse.ArgumentParser("vllm_b200 launch render")
    ap.add_argument("model")
    ap.add_argument("--tokenizer", default=None)
    ap.add_argument("--tokenizer-mode", default="auto")
    ap.add_argument("--chat-template", default=None)
    ap.add_argument("--trust-remote-code", action="store_true")
    ap.add_argument("--max-model-len", type=int, default=None)
    ap.add_argument("--served-model-name", default=None)
    ap.add_argument("--host", default="0.0.0.0")
    ap.add_argument("--port", type=int, default=8100)
    ap.add_argument("--uvicorn-log-level", default="info")
    a = ap.parse_args(argv)
    import uvicorn
    app = build_render_app(a.model, a.tokenizer, a.chat_template, a.tokenizer_mode, a.trust_remote_code, a.max_model_len,
                           a.served_model_name)
    uvicorn.run(app, host=a.host, port=a.port, log_level=a.uvicorn_log_level)
