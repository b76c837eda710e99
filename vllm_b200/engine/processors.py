"""Front-half processing: prompt → EngineCoreRequest, EngineCoreOutput → RequestOutput.

Behavioural reference: vllm/v1/engine/input_processor.py:38 (``process_inputs``),
output_processor.py:443 (``process_outputs``: detokenise, stop strings,
logprobs, RequestOutput assembly), detokenizer.py:168,251 (incremental
detokenisers), logprobs.py, parallel_sampling.py (``ParentRequest`` n>1 fan-out).
"""
from __future__ import annotations

import os

import asyncio
import time
from dataclasses import dataclass, field
from typing import Any, Optional, Union

from vllm_b200.config import VllmConfig
from vllm_b200.core.request import EngineCoreOutput, EngineCoreRequest, FinishReason, LogprobsLists
from vllm_b200.inputs import PromptType, parse_prompt
from vllm_b200.outputs import (CompletionOutput, Logprob, PoolingOutput, PoolingRequestOutput, RequestMetrics,
                               RequestOutput)
from vllm_b200.sampling_params import PoolingParams, RequestOutputKind, SamplingParams


class InputProcessor:
    def __init__(self, vllm_config: VllmConfig, tokenizer) -> None:
        self.vllm_config = vllm_config
        self.mc = vllm_config.model_config
        self.tokenizer = tokenizer
        self.max_model_len = vllm_config.scheduler_config.max_model_len
        self.mm_processor = None
        # sender half of the multimodal input cache (one engine core only: with DP the target engine is picked later)
        self.mm_sender = None
        from vllm_b200.multimodal import MULTIMODAL_REGISTRY
        has_mm = MULTIMODAL_REGISTRY.supports_multimodal(self.mc)
        if (has_mm or self.mc.hf_config.get("is_encoder_decoder")) and vllm_config.parallel_config.data_parallel_size == 1:
            from vllm_b200.multimodal.cache import SenderShadow
            self.mm_sender = SenderShadow(int(os.environ.get("VLLM_B200_MM_CACHE_ITEMS", "256")))
        if has_mm:
            self.mm_processor = MULTIMODAL_REGISTRY.create_processor(self.mc)

    def process_inputs(self, request_id: str, prompt: PromptType,
                       params: Union[SamplingParams, PoolingParams], arrival_time: Optional[float] = None,
                       lora_request=None, priority: int = 0, trace_headers=None,
                       data_parallel_rank: Optional[int] = None,
                       kv_transfer_params: Optional[dict] = None) -> tuple[Optional[str], EngineCoreRequest]:
        text, ids, extras = parse_prompt(prompt)
        embeds_feature = None
        if extras.get("encoder_prompt_token_ids") is not None or extras.get("encoder_prompt") is not None:
            # encoder-decoder request: the source sequence rides along as a feature record, the decoder prompt
            # defaults to the model's decoder start token
            from vllm_b200.multimodal import hash_item
            import numpy as np
            enc = extras.get("encoder_prompt_token_ids")
            if enc is None:
                enc = self.tokenizer.encode(extras["encoder_prompt"])
            enc = list(enc)
            if ids is None and text is None:
                start = self.mc.hf_config.get("decoder_start_token_id", self.mc.hf_config.get("bos_token_id", 0))
                ids = [start]
            embeds_feature = {"modality": "encoder_tokens", "hash": hash_item(np.asarray(enc, dtype=np.int64)),
                              "offset": 0, "length": 0, "data": enc}
        mmd = extras.get("multi_modal_data") or {}
        if self.mc.hf_config.get("is_encoder_decoder") and mmd.get("audio") is not None:
            # speech models: the 30-second log-mel window is the encoder input (reference: models/whisper.py
            # WhisperMultiModalProcessor); decoder prompt = the task tokens the caller built, or the start token
            from vllm_b200.multimodal import hash_item
            from vllm_b200.multimodal.audio import load_audio, log_mel_spectrogram
            audio = mmd["audio"]
            mel = audio if (hasattr(audio, "dim") and audio.dim() == 2) else log_mel_spectrogram(
                load_audio(audio), self.mc.hf_config.get("num_mel_bins", 80))
            if ids is None and text is None:
                ids = [self.mc.hf_config.get("decoder_start_token_id", self.mc.hf_config.get("bos_token_id", 0))]
            embeds_feature = {"modality": "audio", "hash": hash_item(mel.float().numpy()), "offset": 0, "length": 0,
                              "data": mel}
            extras = {k: v for k, v in extras.items() if k != "multi_modal_data"}
        if ids is None and text is None and extras.get("prompt_embeds") is not None:
            # prompt given directly as embeddings [n, hidden] (reference: ``EmbedsPrompt``): the positions carry a
            # dummy token id and the worker splices the rows in exactly like multimodal encoder outputs
            import torch
            from vllm_b200.multimodal import hash_item
            emb = torch.as_tensor(extras["prompt_embeds"])
            if emb.dim() != 2 or emb.shape[1] != self.mc.hidden_size:
                raise ValueError(f"prompt_embeds must be [num_tokens, {self.mc.hidden_size}]")
            ids = [0] * emb.shape[0]
            embeds_feature = {"modality": "embeds", "hash": hash_item(emb.float().numpy()), "offset": 0,
                              "length": emb.shape[0], "data": emb}
        if ids is None:
            if text is None:
                raise ValueError("prompt must contain text or token ids")
            if self.tokenizer is None:
                raise ValueError("tokenizer is disabled (skip_tokenizer_init); pass prompt_token_ids")
            ids = self.tokenizer.encode(text)
        if not ids:
            raise ValueError("the prompt is empty")
        trunc = getattr(params, "truncate_prompt_tokens", None)
        if trunc is not None and trunc > 0:
            ids = ids[-trunc:]
        V = self.mc.vocab_size
        mx = max(ids)
        if mx >= V or min(ids) < 0:
            raise ValueError(f"token id {mx} is out of vocabulary (size {V})")
        if len(ids) >= self.max_model_len:
            raise ValueError(
                f"the prompt ({len(ids)} tokens) is too long for max_model_len={self.max_model_len}; "
                "shorten it or raise --max-model-len")
        mm_features = None
        if extras.get("multi_modal_data") and self.mm_processor is not None:
            lim = self.mc.hf_config.get("limit_mm_per_prompt") or {}
            for kind, items in extras["multi_modal_data"].items():
                cap = lim.get(kind)
                n = len(items) if isinstance(items, (list, tuple)) else (0 if items is None else 1)
                if cap is not None and n > int(cap):
                    raise ValueError(f"at most {cap} {kind} item(s) per prompt are allowed (--limit-mm-per-prompt), got {n}")
            ids, mm_features = self.mm_processor.process(ids, extras["multi_modal_data"])
        if embeds_feature is not None:
            mm_features = (mm_features or []) + [embeds_feature]
        sp = pp = None
        if isinstance(params, SamplingParams):
            sp = params.clone()
            if sp.max_tokens is None:
                sp.max_tokens = self.max_model_len - len(ids)
            sp.max_tokens = max(1, min(sp.max_tokens, self.max_model_len - len(ids)))
            sp.update_from_generation_config(self.mc.eos_token_id)
            sp.update_from_tokenizer(self.tokenizer)
            if sp.logprobs is not None and sp.logprobs > self.mc.max_logprobs >= 0 and sp.logprobs != -1:
                raise ValueError(f"requested logprobs={sp.logprobs} > max_logprobs={self.mc.max_logprobs}")
            if sp.allowed_token_ids and max(sp.allowed_token_ids) >= V:
                raise ValueError("allowed_token_ids contains out-of-vocabulary ids")
            if sp.logit_bias:
                sp.logit_bias = {int(k): float(v) for k, v in sp.logit_bias.items()}
                if max(sp.logit_bias) >= V:
                    raise ValueError("logit_bias contains out-of-vocabulary ids")
            if kv_transfer_params is None and sp.extra_args:
                kv_transfer_params = sp.extra_args.get("kv_transfer_params")
        else:
            pp = params.clone()
        if self.mm_sender is not None:
            self.mm_sender.strip(mm_features)          # last step: nothing after this point may reject the request
        req = EngineCoreRequest(
            request_id=request_id, prompt_token_ids=ids, sampling_params=sp, pooling_params=pp,
            eos_token_id=self.mc.eos_token_id, arrival_time=arrival_time or time.time(),
            lora_request=lora_request, cache_salt=extras.get("cache_salt"), priority=priority,
            data_parallel_rank=data_parallel_rank, kv_transfer_params=kv_transfer_params,
            mm_features=mm_features, trace_headers=trace_headers)
        return text, req


# ---------------------------------------------------------------------------
# Incremental detokenisation
# ---------------------------------------------------------------------------


class IncrementalDetokenizer:
    """Prefix-offset incremental decode: re-decode a small window of trailing tokens and
    emit only text that cannot change any more (no dangling UTF-8 / merge effects)."""

    def __init__(self, tokenizer, prompt_ids: list[int], sp: SamplingParams) -> None:
        self.tok = tokenizer
        self.skip_special = sp.skip_special_tokens
        self.stop = list(sp.stop or [])
        self.include_stop = sp.include_stop_str_in_output
        self.buffer_len = sp.output_text_buffer_length
        self.min_tokens = sp.min_tokens
        self.token_ids: list[int] = list(prompt_ids[-6:])
        self.prefix_offset = 0
        self.read_offset = len(self.token_ids)
        self.n_prompt_ctx = len(self.token_ids)
        self.output_text = ""
        self._sent = 0
        self.output_token_count = 0

    def _decode(self, ids: list[int]) -> str:
        return self.tok.decode(ids, skip_special_tokens=self.skip_special)

    def update(self, new_ids: list[int], stop_terminated: bool) -> Optional[str]:
        """Append tokens; returns the matched stop string if one fired."""
        if not new_ids:
            return None
        skipped_last = None
        if stop_terminated and not self.include_stop:
            skipped_last = new_ids[-1]  # the EOS / stop token itself is not rendered
            new_ids = new_ids[:-1]
        for t in new_ids:
            self.token_ids.append(t)
            self.output_token_count += 1
            prefix = self._decode(self.token_ids[self.prefix_offset:self.read_offset])
            full = self._decode(self.token_ids[self.prefix_offset:])
            if len(full) > len(prefix) and not full.endswith("�"):
                self.output_text += full[len(prefix):]
                self.prefix_offset = self.read_offset
                self.read_offset = len(self.token_ids)
            if self.stop and self.output_token_count > self.min_tokens:
                hit = self._check_stop()
                if hit is not None:
                    return hit
        del skipped_last
        return None

    def _check_stop(self) -> Optional[str]:
        for s in self.stop:
            # only the tail that could contain a new match needs scanning
            start = max(0, len(self.output_text) - len(s) - 64)
            idx = self.output_text.find(s, start)
            if idx != -1:
                end = idx + len(s) if self.include_stop else idx
                self.output_text = self.output_text[:end]
                return s
        return None

    def get_next_text(self, finished: bool, delta: bool) -> str:
        hold = 0 if finished else self.buffer_len
        length = len(self.output_text) - hold
        if delta:
            if length <= self._sent:
                return ""
            out = self.output_text[self._sent:length]
            self._sent = length
            return out
        return self.output_text[:max(length, 0)]


class NoopDetokenizer:
    output_text = ""

    def update(self, new_ids, stop_terminated):
        return None

    def get_next_text(self, finished, delta):
        return ""


# ---------------------------------------------------------------------------


class LogprobsProcessor:
    def __init__(self, tokenizer, sp: SamplingParams) -> None:
        self.tok = tokenizer if sp.detokenize else None
        self.num_logprobs = sp.logprobs
        self.num_prompt_logprobs = sp.prompt_logprobs
        self.logprobs: Optional[list[dict[int, Logprob]]] = [] if sp.logprobs is not None else None
        self.prompt_logprobs: Optional[list[Optional[dict[int, Logprob]]]] = \
            [None] if sp.prompt_logprobs is not None else None
        self.cumulative_logprob: Optional[float] = 0.0 if sp.logprobs is not None else None

    def _entry(self, ids: list[int], vals: list[float], rank: int, k: Optional[int]) -> dict[int, Logprob]:
        d: dict[int, Logprob] = {}
        toks = self.tok.convert_ids_to_tokens(ids) if self.tok is not None else [None] * len(ids)
        for j, (i, v) in enumerate(zip(ids, vals)):
            if j > 0 and k is not None and k >= 0 and j > k:
                break
            r = rank if j == 0 else j
            if i in d:
                continue
            d[i] = Logprob(v, r, toks[j])
        return d

    def update(self, lp: Optional[LogprobsLists], plp: Optional[LogprobsLists]) -> None:
        if lp is not None and self.logprobs is not None:
            for ids, vals, rank in zip(lp.token_ids, lp.logprobs, lp.ranks):
                self.logprobs.append(self._entry(ids, vals, rank, self.num_logprobs))
                self.cumulative_logprob += vals[0]
        if plp is not None and self.prompt_logprobs is not None:
            for ids, vals, rank in zip(plp.token_ids, plp.logprobs, plp.ranks):
                self.prompt_logprobs.append(self._entry(ids, vals, rank, self.num_prompt_logprobs))

    def pop_prompt_logprobs(self):
        p = self.prompt_logprobs
        if p:
            self.prompt_logprobs = []
        return p or None


class RequestOutputCollector:
    """Single-slot mailbox between the output handler task and a per-request generator;
    merges outputs when the consumer is slower than the producer."""

    def __init__(self, kind: RequestOutputKind) -> None:
        self.aggregate = kind == RequestOutputKind.DELTA
        self.output: Any = None
        self.ready = asyncio.Event()

    def put(self, out: Any) -> None:
        if self.output is None or isinstance(out, Exception):
            self.output = out
            self.ready.set()
        elif isinstance(self.output, RequestOutput):
            self.output.add(out, self.aggregate)

    async def get(self) -> Any:
        while self.output is None:
            await self.ready.wait()
        out, self.output = self.output, None
        self.ready.clear()
        if isinstance(out, Exception):
            raise out
        return out

    def get_nowait(self) -> Any:
        out, self.output = self.output, None
        if out is not None:
            self.ready.clear()
        if isinstance(out, Exception):
            raise out
        return out


class ParentRequest:
    """Fan-out of one user request with n>1 into n child requests."""

    def __init__(self, request_id: str, sp: SamplingParams) -> None:
        self.request_id = request_id
        self.sampling_params = sp
        self.n = sp.n
        self.outputs: list[Optional[CompletionOutput]] = [None] * sp.n
        self.finished_children = 0
        self.kind = sp.output_kind

    def child(self, index: int) -> tuple[str, SamplingParams]:
        sp = self.sampling_params.clone()
        sp.n = 1
        if sp.seed is not None:
            sp.seed += index
        return f"{index}_{self.request_id}", sp

    def collect(self, index: int, out: CompletionOutput) -> tuple[list[CompletionOutput], bool]:
        if out.finished():
            self.finished_children += 1
        done = self.finished_children == self.n
        if self.kind != RequestOutputKind.FINAL_ONLY:
            return [out], done
        self.outputs[index] = out
        return ([o for o in self.outputs if o is not None] if done else []), done


@dataclass
class RequestState:
    request_id: str
    parent: Optional[ParentRequest]
    index: int
    prompt: Optional[str]
    prompt_token_ids: list[int]
    detok: Any
    logprobs: Optional[LogprobsProcessor]
    kind: RequestOutputKind
    queue: Optional[RequestOutputCollector]
    is_pooling: bool = False
    output_token_ids: list[int] = field(default_factory=list)
    sent_tokens: int = 0
    metrics: RequestMetrics = field(default_factory=RequestMetrics)
    num_cached_tokens: int = 0
    trace_headers: Optional[dict] = None
    max_tokens: Optional[int] = None


@dataclass
class OutputProcessorOutput:
    request_outputs: list[Any] = field(default_factory=list)
    reqs_to_abort: list[str] = field(default_factory=list)


@dataclass
class IterationStats:
    iteration_timestamp: float = 0.0
    num_generation_tokens: int = 0
    num_prompt_tokens: int = 0
    num_preempted_reqs: int = 0
    time_to_first_tokens: list[float] = field(default_factory=list)
    inter_token_latencies: list[float] = field(default_factory=list)
    finished: list[dict] = field(default_factory=list)  # per finished request timing dicts


class OutputProcessor:
    def __init__(self, tokenizer, log_stats: bool = False, tracer=None) -> None:
        self.tokenizer = tokenizer
        self.log_stats = log_stats
        self.tracer = tracer
        self.request_states: dict[str, RequestState] = {}
        self.parents: dict[str, ParentRequest] = {}

    def get_num_unfinished_requests(self) -> int:
        return len(self.request_states)

    def has_unfinished_requests(self) -> bool:
        return bool(self.request_states)

    def add_request(self, req: EngineCoreRequest, prompt: Optional[str], parent: Optional[ParentRequest] = None,
                    index: int = 0, queue: Optional[RequestOutputCollector] = None) -> None:
        if req.request_id in self.request_states:
            raise ValueError(f"request id {req.request_id} already running")
        sp = req.sampling_params
        if sp is not None:
            detok = IncrementalDetokenizer(self.tokenizer, req.prompt_token_ids, sp) \
                if (sp.detokenize and self.tokenizer is not None) else NoopDetokenizer()
            lp = LogprobsProcessor(self.tokenizer, sp) if (sp.logprobs is not None or sp.prompt_logprobs is not None) else None
            kind = sp.output_kind
        else:
            detok, lp, kind = NoopDetokenizer(), None, RequestOutputKind.FINAL_ONLY
        st = RequestState(req.request_id, parent, index, prompt, req.prompt_token_ids, detok, lp, kind, queue,
                          is_pooling=sp is None, trace_headers=req.trace_headers,
                          max_tokens=sp.max_tokens if sp else None)
        st.metrics.arrival_time = req.arrival_time
        self.request_states[req.request_id] = st
        if parent is not None:
            self.parents[parent.request_id] = parent

    def abort_requests(self, request_ids) -> list[str]:
        out = []
        for rid in request_ids:
            st = self.request_states.pop(rid, None)
            if st is not None:
                out.append(rid)
                if st.parent is not None:
                    pass
            elif rid in self.parents:  # abort all children of a parent
                p = self.parents.pop(rid)
                for i in range(p.n):
                    cid = f"{i}_{rid}"
                    if self.request_states.pop(cid, None) is not None:
                        out.append(cid)
        return out

    def propagate_error(self, e: Exception) -> None:
        for st in self.request_states.values():
            if st.queue is not None:
                st.queue.put(e)

    def process_outputs(self, outputs: list[EngineCoreOutput], engine_ts: Optional[float] = None,
                        stats: Optional[IterationStats] = None) -> OutputProcessorOutput:
        res = OutputProcessorOutput()
        now = time.time()
        for eco in outputs:
            st = self.request_states.get(eco.request_id)
            if st is None:
                continue  # aborted on the front-end while the step was in flight
            new_ids = eco.new_token_ids
            finish = FinishReason(eco.finish_reason) if eco.finish_reason is not None else None
            stop_reason = eco.stop_reason
            st.num_cached_tokens = eco.num_cached_tokens
            if stats is not None:
                self._update_stats(st, eco, stats, now)
            if eco.pooling_output is not None:
                st.request_states = None
                self.request_states.pop(eco.request_id, None)
                out = PoolingRequestOutput(st.request_id, PoolingOutput(eco.pooling_output), st.prompt_token_ids, True)
                (st.queue.put(out) if st.queue is not None else res.request_outputs.append(out))
                continue
            st.output_token_ids.extend(new_ids)
            stop_str = st.detok.update(new_ids, finish == FinishReason.STOP
                                       and not isinstance(stop_reason, str) and bool(new_ids))
            if stop_str is not None and finish is None:
                finish, stop_reason = FinishReason.STOP, stop_str
                res.reqs_to_abort.append(eco.request_id)  # engine does not know about stop strings
            if st.logprobs is not None:
                st.logprobs.update(eco.new_logprobs, eco.new_prompt_logprobs)
            finished = finish is not None
            ro = self._make_request_output(st, new_ids, finish, stop_reason, eco.kv_transfer_params)
            if ro is not None:
                if st.queue is not None:
                    st.queue.put(ro)
                else:
                    res.request_outputs.append(ro)
            if finished:
                self.request_states.pop(eco.request_id, None)
                if st.parent is not None and st.parent.finished_children == st.parent.n:
                    self.parents.pop(st.parent.request_id, None)
                if self.tracer is not None:
                    self.tracer.request_finished(st, finish)
        return res

    def _update_stats(self, st: RequestState, eco: EngineCoreOutput, stats: IterationStats, now: float) -> None:
        m = st.metrics
        n = len(eco.new_token_ids)
        stats.num_generation_tokens += n
        for ev, ts in eco.events or ():
            if ev == 1:
                m.queued_ts = ts
            elif ev == 2 and not m.scheduled_ts:
                m.scheduled_ts = ts
            elif ev == 3:
                stats.num_preempted_reqs += 1
        if n:
            if not m.first_token_ts:
                m.first_token_ts = now
                stats.time_to_first_tokens.append(now - m.arrival_time)
                stats.num_prompt_tokens += len(st.prompt_token_ids)
            else:
                stats.inter_token_latencies.append((now - m.last_token_ts) / n)
            m.last_token_ts = now
            m.num_generation_tokens += n
        if eco.finish_reason is not None:
            m.finished_time = now
            stats.finished.append(dict(
                e2e=now - m.arrival_time, num_prompt_tokens=len(st.prompt_token_ids),
                num_generation_tokens=m.num_generation_tokens, finish_reason=str(FinishReason(eco.finish_reason)),
                queue_time=(m.scheduled_ts - m.queued_ts) if m.scheduled_ts and m.queued_ts else 0.0,
                prefill_time=(m.first_token_ts - m.arrival_time) if m.first_token_ts else 0.0,
                decode_time=(m.last_token_ts - m.first_token_ts) if m.first_token_ts else 0.0,
                max_tokens=st.max_tokens, num_cached_tokens=eco.num_cached_tokens))

    def _make_request_output(self, st: RequestState, new_ids, finish, stop_reason, kv_params) -> Optional[RequestOutput]:
        finished = finish is not None
        if st.kind == RequestOutputKind.FINAL_ONLY and not finished:
            return None
        delta = st.kind == RequestOutputKind.DELTA
        text = st.detok.get_next_text(finished, delta)
        if delta:
            ids = list(new_ids)
            lps = None
            if st.logprobs is not None and st.logprobs.logprobs is not None:
                lps = st.logprobs.logprobs[st.sent_tokens:]
            st.sent_tokens = len(st.output_token_ids)
            if not ids and not text and not finished:
                return None
        else:
            ids = list(st.output_token_ids)
            lps = st.logprobs.logprobs if st.logprobs is not None else None
        co = CompletionOutput(
            index=st.index, text=text, token_ids=ids,
            cumulative_logprob=st.logprobs.cumulative_logprob if st.logprobs else None, logprobs=lps,
            finish_reason=str(finish) if finished else None, stop_reason=stop_reason if finished else None)
        plp = st.logprobs.pop_prompt_logprobs() if st.logprobs is not None else None
        rid, outs, all_done = st.request_id, [co], finished
        if st.parent is not None:
            rid = st.parent.request_id
            outs, all_done = st.parent.collect(st.index, co)
            if not outs:
                return None
        return RequestOutput(request_id=rid, prompt=st.prompt, prompt_token_ids=st.prompt_token_ids,
                             prompt_logprobs=plp, outputs=outs, finished=all_done, metrics=st.metrics,
                             num_cached_tokens=st.num_cached_tokens, kv_transfer_params=kv_params)
