"""Request tracing (reference: vllm/tracing/__init__.py:66-148, otel.py:60-127).

OpenTelemetry spans are emitted when ``--otlp-traces-endpoint`` is set and the
SDK is importable; otherwise an in-memory recorder keeps the last N finished
request spans (also what the tests read).
"""
from __future__ import annotations

import functools
import time
from collections import deque
from typing import Any, Callable, Optional


class SpanAttributes:
    REQUEST_ID = "gen_ai.request.id"
    PROMPT_TOKENS = "gen_ai.usage.prompt_tokens"
    COMPLETION_TOKENS = "gen_ai.usage.completion_tokens"
    MAX_TOKENS = "gen_ai.request.max_tokens"
    N = "gen_ai.request.n"
    QUEUE_TIME = "gen_ai.latency.time_in_queue"
    TTFT = "gen_ai.latency.time_to_first_token"
    E2E = "gen_ai.latency.e2e"
    PREFILL_TIME = "gen_ai.latency.time_in_model_prefill"
    DECODE_TIME = "gen_ai.latency.time_in_model_decode"
    FINISH_REASON = "gen_ai.response.finish_reason"


def extract_trace_context(headers: Optional[dict[str, str]]):
    """W3C trace-context propagation from HTTP headers."""
    if not headers:
        return None
    try:
        from opentelemetry.trace.propagation.tracecontext import TraceContextTextMapPropagator
        return TraceContextTextMapPropagator().extract(headers)
    except Exception:
        return None


class RequestTracer:
    def __init__(self, endpoint: Optional[str] = None, keep: int = 1024) -> None:
        self.finished: deque[dict[str, Any]] = deque(maxlen=keep)
        self._otel = None
        self._provider = None
        if endpoint:
            try:
                from opentelemetry import trace
                from opentelemetry.sdk.resources import Resource
                from opentelemetry.sdk.trace import TracerProvider
                from opentelemetry.sdk.trace.export import BatchSpanProcessor
                if endpoint.startswith("http"):
                    from opentelemetry.exporter.otlp.proto.http.trace_exporter import OTLPSpanExporter
                else:
                    from opentelemetry.exporter.otlp.proto.grpc.trace_exporter import OTLPSpanExporter
                provider = TracerProvider(resource=Resource.create({"service.name": "vllm_b200"}))
                provider.add_span_processor(BatchSpanProcessor(OTLPSpanExporter(endpoint=endpoint)))
                self._provider = provider
                self._otel = provider.get_tracer("vllm_b200.llm_engine")
            except Exception:
                self._otel = None

    def request_finished(self, st, finish_reason) -> None:
        m = st.metrics
        now = time.time()
        attrs = {
            SpanAttributes.REQUEST_ID: st.request_id,
            SpanAttributes.PROMPT_TOKENS: len(st.prompt_token_ids),
            SpanAttributes.COMPLETION_TOKENS: len(st.output_token_ids),
            SpanAttributes.TTFT: (m.first_token_ts - m.arrival_time) if m.first_token_ts else 0.0,
            SpanAttributes.E2E: now - m.arrival_time,
            SpanAttributes.QUEUE_TIME: (m.scheduled_ts - m.queued_ts) if m.scheduled_ts and m.queued_ts else 0.0,
            SpanAttributes.DECODE_TIME: (m.last_token_ts - m.first_token_ts) if m.first_token_ts else 0.0,
            SpanAttributes.FINISH_REASON: str(finish_reason),
        }
        if st.max_tokens is not None:
            attrs[SpanAttributes.MAX_TOKENS] = st.max_tokens
        self.finished.append(attrs)
        if self._otel is not None:
            ctx = extract_trace_context(st.trace_headers)
            with self._otel.start_as_current_span("llm_request", context=ctx,
                                                  start_time=int(m.arrival_time * 1e9)) as span:
                for k, v in attrs.items():
                    span.set_attribute(k, v)


    def flush(self, timeout_ms: int = 5000) -> bool:
        """Push pending spans to the collector now (shutdown, tests)."""
        return bool(self._provider.force_flush(timeout_ms)) if self._provider is not None else True


def maybe_init_tracer(obs_config) -> Optional[RequestTracer]:
    if obs_config is None or not obs_config.otlp_traces_endpoint:
        return None
    return RequestTracer(obs_config.otlp_traces_endpoint)


_spans: deque = deque(maxlen=4096)


def instrument(span_name: Optional[str] = None) -> Callable:
    """Decorator recording wall-time of a function as a lightweight span."""
    def deco(fn):
        name = span_name or fn.__qualname__

        @functools.wraps(fn)
        def wrapper(*a, **kw):
            t0 = time.perf_counter()
            try:
                return fn(*a, **kw)
            finally:
                _spans.append((name, time.perf_counter() - t0))
        return wrapper
    return deco


def recorded_spans() -> list[tuple[str, float]]:
    return list(_spans)
