"""gRPC front-end (reference: vllm/entrypoints/grpc_server.py + rust/proto/*.proto: token / text generation as a
server-streaming RPC next to the HTTP API).

The protobuf messages of ``vllm_b200.proto`` are built programmatically (descriptor pool + message factory), so the
server and the bundled client need only the ``protobuf`` and ``grpcio`` runtimes — no generated ``*_pb2.py``.

    python -m vllm_b200.entrypoints.grpc llama-3-8b --port 50051
    vllm_b200 serve MODEL --grpc-port 50051          # next to the HTTP server, same engine
"""
from __future__ import annotations

import asyncio
from typing import Any, AsyncIterator, Optional

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

from vllm_b200.logger import init_logger

logger = init_logger(__name__)
PKG = "vllm_b200.grpc"
_T = descriptor_pb2.FieldDescriptorProto


def _build_messages() -> dict[str, Any]:
    fd = descriptor_pb2.FileDescriptorProto(name="vllm_b200.proto", package=PKG, syntax="proto3")

    def msg(name, *fields):
        m = fd.message_type.add(name=name)
        for fname, num, typ, rep, tname in fields:
            f = m.field.add(name=fname, number=num, type=typ,
                            label=_T.LABEL_REPEATED if rep else _T.LABEL_OPTIONAL)
            if tname:
                f.type_name = f".{PKG}.{tname}"
    F, I32, I64, B, S, M = _T.TYPE_FLOAT, _T.TYPE_INT32, _T.TYPE_INT64, _T.TYPE_BOOL, _T.TYPE_STRING, _T.TYPE_MESSAGE
    msg("SamplingParams", ("temperature", 1, F, 0, 0), ("top_p", 2, F, 0, 0), ("top_k", 3, I32, 0, 0),
        ("max_tokens", 4, I32, 0, 0), ("seed", 5, I64, 0, 0), ("has_seed", 6, B, 0, 0), ("stop_token_ids", 7, I32, 1, 0),
        ("ignore_eos", 8, B, 0, 0), ("stop", 9, S, 1, 0), ("min_p", 10, F, 0, 0), ("repetition_penalty", 11, F, 0, 0),
        ("presence_penalty", 12, F, 0, 0), ("frequency_penalty", 13, F, 0, 0), ("has_temperature", 14, B, 0, 0))
    msg("GenerateRequest", ("request_id", 1, S, 0, 0), ("prompt", 2, S, 0, 0), ("prompt_token_ids", 3, I32, 1, 0),
        ("sampling_params", 4, M, 0, "SamplingParams"), ("stream", 5, B, 0, 0), ("priority", 6, I32, 0, 0))
    msg("GenerateResponse", ("request_id", 1, S, 0, 0), ("text", 2, S, 0, 0), ("token_ids", 3, I32, 1, 0),
        ("finished", 4, B, 0, 0), ("finish_reason", 5, S, 0, 0), ("num_prompt_tokens", 6, I32, 0, 0),
        ("num_cached_tokens", 7, I32, 0, 0))
    msg("HealthRequest")
    msg("HealthResponse", ("healthy", 1, B, 0, 0), ("version", 2, S, 0, 0))
    msg("AbortRequest", ("request_ids", 1, S, 1, 0))
    msg("AbortResponse")
    msg("ModelInfoRequest")
    msg("ModelInfoResponse", ("model", 1, S, 0, 0), ("max_model_len", 2, I32, 0, 0), ("vocab_size", 3, I32, 0, 0))
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return {m.name: message_factory.GetMessageClass(pool.FindMessageTypeByName(f"{PKG}.{m.name}")) for m in fd.message_type}


PB = _build_messages()
SERVICE = f"{PKG}.Engine"


def _sampling_params(sp_pb):
    from vllm_b200.sampling_params import SamplingParams
    kw: dict[str, Any] = {"max_tokens": sp_pb.max_tokens or 16}
    if sp_pb.has_temperature or sp_pb.temperature:
        kw["temperature"] = sp_pb.temperature
    if sp_pb.top_p:
        kw["top_p"] = sp_pb.top_p
    if sp_pb.top_k:
        kw["top_k"] = sp_pb.top_k
    if sp_pb.min_p:
        kw["min_p"] = sp_pb.min_p
    if sp_pb.has_seed:
        kw["seed"] = sp_pb.seed
    for k in ("repetition_penalty", "presence_penalty", "frequency_penalty"):
        if getattr(sp_pb, k):
            kw[k] = getattr(sp_pb, k)
    if sp_pb.stop_token_ids:
        kw["stop_token_ids"] = list(sp_pb.stop_token_ids)
    if sp_pb.stop:
        kw["stop"] = list(sp_pb.stop)
    if sp_pb.ignore_eos:
        kw["ignore_eos"] = True
    return SamplingParams(**kw)


class EngineServicer:
    def __init__(self, engine, model_name: str) -> None:
        self.engine, self.model_name = engine, model_name
        self._n = 0

    async def Generate(self, req, context) -> AsyncIterator[Any]:
        import grpc
        from vllm_b200.utils import random_uuid
        rid = req.request_id or f"grpc-{random_uuid()}"
        prompt: Any = {"prompt_token_ids": list(req.prompt_token_ids)} if req.prompt_token_ids else req.prompt
        if not prompt:
            await context.abort(grpc.StatusCode.INVALID_ARGUMENT, "prompt or prompt_token_ids is required")
        try:
            sp = _sampling_params(req.sampling_params)
        except (TypeError, ValueError) as e:
            await context.abort(grpc.StatusCode.INVALID_ARGUMENT, str(e))
        sent_t, sent_c, last = 0, 0, None
        try:
            async for o in self.engine.generate(prompt, sp, rid, priority=req.priority):
                last = o
                c = o.outputs[0]
                if req.stream:
                    yield PB["GenerateResponse"](
                        request_id=rid, text=c.text[sent_c:], token_ids=list(c.token_ids[sent_t:]), finished=o.finished,
                        finish_reason=c.finish_reason or "", num_prompt_tokens=len(o.prompt_token_ids or ()),
                        num_cached_tokens=o.num_cached_tokens or 0)
                    sent_t, sent_c = len(c.token_ids), len(c.text)
        except asyncio.CancelledError:           # client went away
            await self.engine.abort([rid])
            raise
        except ValueError as e:
            await context.abort(grpc.StatusCode.INVALID_ARGUMENT, str(e))
        if not req.stream and last is not None:
            c = last.outputs[0]
            yield PB["GenerateResponse"](request_id=rid, text=c.text, token_ids=list(c.token_ids), finished=True,
                                         finish_reason=c.finish_reason or "", num_prompt_tokens=len(last.prompt_token_ids or ()),
                                         num_cached_tokens=last.num_cached_tokens or 0)

    async def Health(self, req, context):
        from vllm_b200.version import __version__
        ok = True
        try:
            await self.engine.check_health()
        except Exception:
            ok = False
        return PB["HealthResponse"](healthy=ok, version=__version__)

    async def Abort(self, req, context):
        await self.engine.abort(list(req.request_ids))
        return PB["AbortResponse"]()

    async def GetModelInfo(self, req, context):
        mc = self.engine.model_config
        return PB["ModelInfoResponse"](model=self.model_name, max_model_len=mc.max_model_len, vocab_size=mc.vocab_size)


def _handlers(svc: EngineServicer):
    import grpc
    def ser(m):
        return m.SerializeToString()
    def de(name):
        return PB[name].FromString
    return grpc.method_handlers_generic_handler(SERVICE, {
        "Generate": grpc.unary_stream_rpc_method_handler(svc.Generate, de("GenerateRequest"), ser),
        "Health": grpc.unary_unary_rpc_method_handler(svc.Health, de("HealthRequest"), ser),
        "Abort": grpc.unary_unary_rpc_method_handler(svc.Abort, de("AbortRequest"), ser),
        "GetModelInfo": grpc.unary_unary_rpc_method_handler(svc.GetModelInfo, de("ModelInfoRequest"), ser),
    })


async def start_grpc_server(engine, model_name: str, host: str = "0.0.0.0", port: int = 50051):
    import grpc
    server = grpc.aio.server(options=[("grpc.max_send_message_length", -1), ("grpc.max_receive_message_length", -1)])
    server.add_generic_rpc_handlers((_handlers(EngineServicer(engine, model_name)),))
    bound = server.add_insecure_port(f"{host}:{port}")
    await server.start()
    logger.info("gRPC engine service on %s:%d", host, bound)
    return server, bound


class EngineClient:
    """Minimal asyncio client for the service (tests, examples)."""

    def __init__(self, target: str) -> None:
        import grpc
        self.channel = grpc.aio.insecure_channel(target)

        def ser(m):
            return m.SerializeToString()
        self._generate = self.channel.unary_stream(f"/{SERVICE}/Generate", ser, PB["GenerateResponse"].FromString)
        self._health = self.channel.unary_unary(f"/{SERVICE}/Health", ser, PB["HealthResponse"].FromString)
        self._abort = self.channel.unary_unary(f"/{SERVICE}/Abort", ser, PB["AbortResponse"].FromString)
        self._info = self.channel.unary_unary(f"/{SERVICE}/GetModelInfo", ser, PB["ModelInfoResponse"].FromString)

    def generate(self, prompt=None, prompt_token_ids=None, stream: bool = False, request_id: str = "", **sampling):
        sp = PB["SamplingParams"](**{k: v for k, v in sampling.items() if k not in ("seed",)})
        if "temperature" in sampling:
            sp.has_temperature = True
        if sampling.get("seed") is not None:
            sp.seed, sp.has_seed = int(sampling["seed"]), True
        req = PB["GenerateRequest"](request_id=request_id, prompt=prompt or "", prompt_token_ids=prompt_token_ids or [],
                                    sampling_params=sp, stream=stream)
        return self._generate(req)

    async def health(self):
        return await self._health(PB["HealthRequest"]())

    async def model_info(self):
        return await self._info(PB["ModelInfoRequest"]())

    async def abort(self, ids: list[str]):
        return await self._abort(PB["AbortRequest"](request_ids=ids))

    async def close(self) -> None:
        await self.channel.close()


def main(argv=None) -> None:
    import argparse
    from vllm_b200.engine.arg_utils import EngineArgs
    from vllm_b200.engine.async_llm import AsyncLLM
    ap = argparse.ArgumentParser("vllm_b200 grpc server")
    ap.add_argument("model_tag", nargs="?", default=None)
    ap.add_argument("--host", default="0.0.0.0")
    ap.add_argument("--port", type=int, default=50051)
    EngineArgs.add_cli_args(ap)
    args = ap.parse_args(argv)
    if args.model_tag:
        args.model = args.model_tag
    ea = EngineArgs.from_cli_args(args)

    async def run():
        engine = AsyncLLM.from_engine_args(ea)
        server, _ = await start_grpc_server(engine, ea.served_model_name or ea.model, args.host, args.port)
        try:
            await server.wait_for_termination()
        finally:
            engine.shutdown()
    asyncio.run(run())
