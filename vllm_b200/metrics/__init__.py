"""Stats → log line + Prometheus (reference: vllm/v1/metrics/loggers.py:99-330
``LoggingStatLogger``, :443-1282 ``PrometheusStatLogger`` with the ``vllm:*``
metric families; stats.py:186,425 ``SchedulerStats``/``IterationStats``;
reader.py in-process snapshot)."""
from __future__ import annotations

import time
from typing import Any, Optional

from vllm_b200 import envs
from vllm_b200.logger import init_logger

logger = init_logger(__name__)

_TTFT_BUCKETS = [0.001, 0.005, 0.01, 0.02, 0.04, 0.06, 0.08, 0.1, 0.25, 0.5, 0.75, 1.0, 2.5, 5.0, 7.5, 10.0,
                 20.0, 40.0, 80.0, 160.0, 640.0, 2560.0]
_ITL_BUCKETS = [0.001, 0.0025, 0.005, 0.0075, 0.01, 0.015, 0.02, 0.025, 0.03, 0.04, 0.05, 0.075, 0.1, 0.15, 0.2,
                0.3, 0.4, 0.5, 0.75, 1.0, 2.5, 5.0, 7.5, 10.0]
_REQ_BUCKETS = [0.3, 0.5, 0.8, 1.0, 1.5, 2.0, 2.5, 5.0, 10.0, 15.0, 20.0, 30.0, 40.0, 50.0, 60.0, 120.0, 240.0,
                480.0, 960.0, 1920.0, 7680.0]
_TOK_BUCKETS = [1, 2, 5, 10, 20, 50, 100, 200, 500, 1000, 2000, 5000, 10000, 20000, 50000, 100000]


class LoggingStatLogger:
    def __init__(self, vllm_config, engine_index: int = 0) -> None:
        self.interval = envs.VLLM_LOG_STATS_INTERVAL
        self.last = time.monotonic()
        self.prompt_tokens = self.gen_tokens = 0
        self.hits = self.queries = 0
        self.last_sched = None

    def record(self, sched, it) -> None:
        if it is not None:
            self.prompt_tokens += it.num_prompt_tokens
            self.gen_tokens += it.num_generation_tokens
        if sched is not None:
            self.last_sched = sched
            self.hits += sched.prefix_cache_hits
            self.queries += sched.prefix_cache_queries
        now = time.monotonic()
        if now - self.last >= self.interval:
            self.log(now)

    def log(self, now: Optional[float] = None) -> None:
        now = now or time.monotonic()
        dt = max(now - self.last, 1e-9)
        s = self.last_sched
        logger.info("Avg prompt throughput: %.1f tokens/s, Avg generation throughput: %.1f tokens/s, "
                    "Running: %d reqs, Waiting: %d reqs, GPU KV cache usage: %.1f%%, Prefix cache hit rate: %.1f%%",
                    self.prompt_tokens / dt, self.gen_tokens / dt, s.num_running_reqs if s else 0,
                    s.num_waiting_reqs if s else 0, (s.kv_cache_usage if s else 0) * 100,
                    100.0 * self.hits / self.queries if self.queries else 0.0)
        self.prompt_tokens = self.gen_tokens = 0
        self.last = now


_PROM: dict[str, Any] = {}


def _prom_metrics(model_name: str):
    """Create the metric families once per process (prometheus_client forbids duplicates)."""
    if _PROM:
        return _PROM
    from prometheus_client import Counter, Gauge, Histogram
    L = ["model_name", "engine"]
    g = lambda n, d: Gauge(n, d, L, multiprocess_mode="mostrecent")
    c = lambda n, d, extra=(): Counter(n, d, L + list(extra))
    h = lambda n, d, b: Histogram(n, d, L, buckets=b)
    _PROM.update(
        running=g("vllm:num_requests_running", "Number of requests in model execution batches."),
        waiting=g("vllm:num_requests_waiting", "Number of requests waiting to be processed."),
        kv_usage=g("vllm:kv_cache_usage_perc", "KV-cache usage. 1 means 100 percent usage."),
        prefix_q=c("vllm:prefix_cache_queries", "Prefix cache queries, in tokens."),
        prefix_h=c("vllm:prefix_cache_hits", "Prefix cache hits, in tokens."),
        preempt=c("vllm:num_preemptions", "Cumulative number of preemptions."),
        prompt=c("vllm:prompt_tokens", "Number of prefill tokens processed."),
        gen=c("vllm:generation_tokens", "Number of generation tokens processed."),
        success=c("vllm:request_success", "Count of successfully processed requests.", ("finished_reason",)),
        iter_tokens=h("vllm:iteration_tokens_total", "Histogram of number of tokens per engine step.", _TOK_BUCKETS),
        ttft=h("vllm:time_to_first_token_seconds", "Histogram of time to first token in seconds.", _TTFT_BUCKETS),
        itl=h("vllm:inter_token_latency_seconds", "Histogram of inter-token latency in seconds.", _ITL_BUCKETS),
        tpot=h("vllm:request_time_per_output_token_seconds", "Histogram of time per output token.", _ITL_BUCKETS),
        e2e=h("vllm:e2e_request_latency_seconds", "Histogram of e2e request latency in seconds.", _REQ_BUCKETS),
        queue=h("vllm:request_queue_time_seconds", "Time spent in WAITING phase.", _REQ_BUCKETS),
        prefill=h("vllm:request_prefill_time_seconds", "Time spent in PREFILL phase.", _REQ_BUCKETS),
        decode=h("vllm:request_decode_time_seconds", "Time spent in DECODE phase.", _REQ_BUCKETS),
        inference=h("vllm:request_inference_time_seconds", "Time spent in RUNNING phase.", _REQ_BUCKETS),
        n_prompt=h("vllm:request_prompt_tokens", "Number of prefill tokens per request.", _TOK_BUCKETS),
        n_gen=h("vllm:request_generation_tokens", "Number of generation tokens per request.", _TOK_BUCKETS),
        max_gen=h("vllm:request_params_max_tokens", "max_tokens request parameter.", _TOK_BUCKETS),
        spec_drafts=c("vllm:spec_decode_num_drafts", "Number of spec decoding drafts."),
        spec_draft_tokens=c("vllm:spec_decode_num_draft_tokens", "Number of draft tokens."),
        spec_accepted=c("vllm:spec_decode_num_accepted_tokens", "Number of accepted draft tokens."),
        fwd_ms=g("vllm:model_forward_time_milliseconds", "Last model step time in ms."),
        sched_ms=g("vllm:scheduler_time_milliseconds", "Last schedule() time in ms."),
    )
    return _PROM


class PrometheusStatLogger:
    def __init__(self, vllm_config, engine_index: int = 0) -> None:
        self.m = _prom_metrics(vllm_config.model_config.served_model_name)
        self.labels = dict(model_name=vllm_config.model_config.served_model_name, engine=str(engine_index))

    def _l(self, k, **extra):
        return self.m[k].labels(**self.labels, **extra)

    def record(self, sched, it) -> None:
        if sched is not None:
            self._l("running").set(sched.num_running_reqs)
            self._l("waiting").set(sched.num_waiting_reqs)
            self._l("kv_usage").set(sched.kv_cache_usage)
            self._l("prefix_q").inc(sched.prefix_cache_queries)
            self._l("prefix_h").inc(sched.prefix_cache_hits)
            self._l("preempt").inc(sched.num_preemptions)
            self._l("spec_drafts").inc(sched.spec_num_drafts)
            self._l("spec_draft_tokens").inc(sched.spec_num_draft_tokens)
            self._l("spec_accepted").inc(sched.spec_num_accepted_tokens)
            self._l("fwd_ms").set(sched.model_forward_ms)
            self._l("sched_ms").set(sched.schedule_ms)
            if sched.num_scheduled_tokens:
                self._l("iter_tokens").observe(sched.num_scheduled_tokens)
        if it is None:
            return
        self._l("prompt").inc(it.num_prompt_tokens)
        self._l("gen").inc(it.num_generation_tokens)
        for v in it.time_to_first_tokens:
            self._l("ttft").observe(v)
        for v in it.inter_token_latencies:
            self._l("itl").observe(v)
        for f in it.finished:
            self._l("success", finished_reason=f["finish_reason"]).inc()
            self._l("e2e").observe(f["e2e"])
            self._l("queue").observe(f["queue_time"])
            self._l("prefill").observe(f["prefill_time"])
            self._l("decode").observe(f["decode_time"])
            self._l("inference").observe(f["prefill_time"] + f["decode_time"])
            self._l("n_prompt").observe(f["num_prompt_tokens"])
            self._l("n_gen").observe(f["num_generation_tokens"])
            if f.get("max_tokens"):
                self._l("max_gen").observe(f["max_tokens"])
            if f["num_generation_tokens"] > 1:
                self._l("tpot").observe(f["decode_time"] / (f["num_generation_tokens"] - 1))

    def log(self) -> None:
        pass


class StatLoggerManager:
    def __init__(self, vllm_config, custom=None, engine_index: int = 0, prometheus: bool = True) -> None:
        self.loggers: list[Any] = [LoggingStatLogger(vllm_config, engine_index)]
        if prometheus:
            try:
                self.loggers.append(PrometheusStatLogger(vllm_config, engine_index))
            except Exception as e:  # pragma: no cover
                logger.warning("prometheus metrics disabled: %s", e)
        for f in custom or []:
            self.loggers.append(f(vllm_config, engine_index))

    def record(self, sched, it) -> None:
        for l in self.loggers:
            l.record(sched, it)

    def log(self) -> None:
        for l in self.loggers:
            l.log()


def get_metrics_snapshot() -> list[dict[str, Any]]:
    """In-process reader (reference: v1/metrics/reader.py)."""
    from prometheus_client import REGISTRY
    out = []
    for fam in REGISTRY.collect():
        if not fam.name.startswith("vllm:"):
            continue
        for s in fam.samples:
            out.append(dict(name=s.name, labels=dict(s.labels), value=s.value, type=fam.type))
    return out


def render_prometheus() -> tuple[bytes, str]:
    from prometheus_client import CONTENT_TYPE_LATEST, REGISTRY, generate_latest
    return generate_latest(REGISTRY), CONTENT_TYPE_LATEST
