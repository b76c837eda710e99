"""Unified token-budget continuous-batching scheduler.

Behavioural reference: vllm/v1/core/sched/scheduler.py — ``schedule`` :499
(running first :546-758, waiting :770-1190, output build :1296-1317,
``_update_after_schedule`` :1405), ``update_from_output`` :1759-2183,
preemption-by-recompute :1362-1403, stop checks (v1/core/sched/utils.py).

There is no separate prefill / decode phase: every request has
``num_computed_tokens`` and ``num_tokens_with_spec``; each step assigns tokens
until the budget is exhausted, so chunked prefill, prefix caching and
speculative decoding all fall out of the same loop.
"""
from __future__ import annotations

import heapq
import time
from collections import deque
from typing import Any, Iterable, Optional, Union

from vllm_b200.config import VllmConfig
from vllm_b200.core.kv_cache_manager import KVCacheConfig, KVCacheManager
from vllm_b200.core.request import (EngineCoreEventType, EngineCoreOutput, EngineCoreOutputs, LogprobsLists,
                                    Request, RequestStatus, SchedulerStats)
from vllm_b200.core.sched_output import CachedRequestData, ModelRunnerOutput, NewRequestData, SchedulerOutput
from vllm_b200.logger import init_logger

logger = init_logger(__name__)


# ---------------------------------------------------------------------------
# Waiting queues
# ---------------------------------------------------------------------------


class FCFSQueue:
    def __init__(self) -> None:
        self.q: deque[Request] = deque()

    def add(self, r: Request) -> None:
        self.q.append(r)

    def prepend(self, r: Request) -> None:
        self.q.appendleft(r)

    def prepend_many(self, rs: Iterable[Request]) -> None:
        self.q.extendleft(reversed(list(rs)))

    def peek(self) -> Request:
        return self.q[0]

    def pop(self) -> Request:
        return self.q.popleft()

    def remove(self, r: Request) -> None:
        self.q.remove(r)

    def __len__(self) -> int:
        return len(self.q)

    def __iter__(self):
        return iter(self.q)

    def __bool__(self) -> bool:
        return bool(self.q)


class PriorityQueue:
    """Min-heap on (priority, arrival_time)."""

    def __init__(self) -> None:
        self.h: list[Request] = []

    def add(self, r: Request) -> None:
        heapq.heappush(self.h, r)

    prepend = add

    def prepend_many(self, rs: Iterable[Request]) -> None:
        for r in rs:
            self.add(r)

    def peek(self) -> Request:
        return self.h[0]

    def pop(self) -> Request:
        return heapq.heappop(self.h)

    def remove(self, r: Request) -> None:
        self.h.remove(r)
        heapq.heapify(self.h)

    def __len__(self) -> int:
        return len(self.h)

    def __iter__(self):
        return iter(sorted(self.h))

    def __bool__(self) -> bool:
        return bool(self.h)


def check_stop(request: Request, max_model_len: int) -> bool:
    """Length / EOS / stop-token checks after appending one token."""
    sp = request.sampling_params
    if request.num_tokens >= max_model_len or request.num_output_tokens >= request.max_tokens:
        request.status = RequestStatus.FINISHED_LENGTH_CAPPED
        return True
    if request.pooling_params is not None:
        return False
    assert sp is not None
    if request.num_output_tokens < sp.min_tokens:
        return False
    last = request.output_token_ids[-1]
    if not sp.ignore_eos and last == request.eos_token_id:
        request.status = RequestStatus.FINISHED_STOPPED
        return True
    if last in sp.all_stop_token_ids and not (sp.ignore_eos and last == request.eos_token_id):
        request.status = RequestStatus.FINISHED_STOPPED
        request.stop_reason = last if last != request.eos_token_id else None
        return True
    return False


class Scheduler:
    def __init__(self, vllm_config: VllmConfig, kv_cache_config: KVCacheConfig,
                 structured_output_manager: Any = None, kv_connector: Any = None,
                 include_finished_set: bool = False, log_stats: bool = False) -> None:
        self.vllm_config = vllm_config
        sc = self.scheduler_config = vllm_config.scheduler_config
        cc = self.cache_config = vllm_config.cache_config
        self.max_num_running_reqs = sc.max_num_seqs
        self.max_num_scheduled_tokens = sc.max_num_batched_tokens
        self.max_model_len = sc.max_model_len
        self.log_stats = log_stats
        self.structured_output_manager = structured_output_manager
        self.connector = kv_connector
        self.include_finished_set = include_finished_set
        spec = vllm_config.speculative_config
        self.num_spec_tokens = spec.num_speculative_tokens if spec else 0
        self.spec_dynamic_k = bool(spec and getattr(spec, "dynamic_k", False))
        self.spec_ema = float(getattr(spec, "dynamic_k_ema", 0.5)) if spec else 0.5
        self.use_eagle = bool(spec and spec.method in ("eagle", "eagle3", "mtp"))
        self.num_lookahead_tokens = self.num_spec_tokens if self.use_eagle else 0
        ev = vllm_config.kv_events_config
        self.kv_cache_manager = KVCacheManager(
            kv_cache_config, max_model_len=self.max_model_len, enable_caching=cc.enable_prefix_caching,
            hash_algo=cc.prefix_caching_hash_algo,
            enable_kv_cache_events=bool(ev and ev.enable_kv_cache_events), log_stats=log_stats,
            use_eagle=self.use_eagle)
        self.block_size = self.kv_cache_manager.block_size
        from vllm_b200.core.kv_events import make_publisher
        self.kv_event_publisher = make_publisher(ev, vllm_config.parallel_config.data_parallel_rank)
        self.requests: dict[str, Request] = {}
        self.waiting: Union[FCFSQueue, PriorityQueue] = PriorityQueue() if sc.policy == "priority" else FCFSQueue()
        self.running: list[Request] = []
        self.finished_req_ids: set[str] = set()
        self.finished_req_ids_dict: Optional[dict[int, set[str]]] = {} if include_finished_set else None
        self.async_scheduling = sc.async_scheduling
        self.max_loras = vllm_config.lora_config.max_loras if vllm_config.lora_config else 0
        # stats
        self._num_preemptions = 0
        self._step = 0
        self._spec_drafts = self._spec_draft_tokens = self._spec_accepted = 0
        self._last_sched_ms = 0.0

    # ------------------------------------------------------------------ API
    def get_block_hasher(self):
        return self.kv_cache_manager.block_hasher

    def add_request(self, request: Request) -> None:
        self.waiting.add(request)
        self.requests[request.request_id] = request
        if self.log_stats:
            request.record_event(EngineCoreEventType.QUEUED)

    def has_unfinished_requests(self) -> bool:
        return bool(self.running) or bool(self.waiting)

    def has_finished_requests(self) -> bool:
        return bool(self.finished_req_ids)

    def has_requests(self) -> bool:
        return self.has_unfinished_requests() or self.has_finished_requests()

    def get_num_unfinished_requests(self) -> int:
        return len(self.running) + len(self.waiting)

    def get_request_counts(self) -> tuple[int, int]:
        return len(self.running), len(self.waiting)

    def reset_prefix_cache(self) -> bool:
        return self.kv_cache_manager.reset_prefix_cache()

    # ------------------------------------------------------------- schedule
    def schedule(self) -> SchedulerOutput:
        t0 = time.perf_counter()
        kv = self.kv_cache_manager
        budget = self.max_num_scheduled_tokens
        thr = self.scheduler_config.long_prefill_token_threshold
        scheduled_new: list[Request] = []
        scheduled_resumed: list[Request] = []
        scheduled_running: list[Request] = []
        preempted: list[Request] = []
        new_blocks_of: dict[str, Any] = {}
        num_sched: dict[str, int] = {}
        spec_sched: dict[str, list[int]] = {}
        structured_ids: list[str] = []

        # 1. running requests
        i = 0
        while i < len(self.running) and budget > 0:
            req = self.running[i]
            n_new = req.num_tokens_with_spec + req.num_output_placeholders - req.num_computed_tokens
            if 0 < thr < n_new:
                n_new = thr
            n_new = min(n_new, budget, self.max_model_len - 1 - req.num_computed_tokens)
            if self.async_scheduling and req.num_output_placeholders > 0 and (
                    req.num_output_tokens + req.num_output_placeholders >= req.max_tokens
                    or req.num_tokens + req.num_output_placeholders >= self.max_model_len):
                n_new = 0  # the in-flight tokens already reach the length cap: wait for them
            if n_new <= 0:
                i += 1
                continue
            can_schedule = True
            while True:
                blocks = kv.allocate_slots(req, n_new, num_lookahead_tokens=self.num_lookahead_tokens)
                if blocks is not None:
                    break
                # out of KV space: preempt (recompute later)
                if isinstance(self.waiting, PriorityQueue):
                    victim = max(self.running, key=lambda r: (r.priority, r.arrival_time))
                    self.running.remove(victim)
                    if victim in scheduled_running:
                        scheduled_running.remove(victim)
                        budget += num_sched.pop(victim.request_id)
                        new_blocks_of.pop(victim.request_id, None)
                        spec_sched.pop(victim.request_id, None)
                        i -= 1
                else:
                    victim = self.running.pop()
                self._preempt(victim)
                preempted.append(victim)
                if victim is req:
                    can_schedule = False
                    break
            if not can_schedule:
                break
            scheduled_running.append(req)
            new_blocks_of[req.request_id] = blocks
            num_sched[req.request_id] = n_new
            budget -= n_new
            i += 1
            if req.spec_token_ids:
                n_spec = n_new + req.num_computed_tokens - req.num_tokens - req.num_output_placeholders
                if n_spec > 0:
                    del req.spec_token_ids[n_spec:]
                    spec_sched[req.request_id] = list(req.spec_token_ids)
            if req.use_structured_output:
                structured_ids.append(req.request_id)

        # 2. waiting requests (only when nothing was preempted this step)
        skipped: list[Request] = []
        if not preempted:
            active_loras = {r.lora_request.lora_int_id for r in self.running if r.lora_request}
            while self.waiting and budget > 0 and len(self.running) < self.max_num_running_reqs:
                req = self.waiting.peek()
                if req.status == RequestStatus.WAITING_FOR_FSM:
                    so = req.structured_output_request
                    if so is not None and so.is_ready():
                        req.status = RequestStatus.WAITING
                    else:
                        skipped.append(self.waiting.pop())
                        continue
                if req.status == RequestStatus.WAITING_FOR_REMOTE_KVS:
                    if self.connector is not None and self.connector.is_recv_finished(req):
                        req.status = RequestStatus.WAITING
                        if self.connector.take_recv_failed(req):
                            # the transfer died: its destination blocks hold garbage.  Give them back uncached and
                            # recompute the prompt locally (reference: invalid-block handling in scheduler.py
                            # ``_handle_invalid_blocks``); the request must not ask for the remote copy again
                            kv.free(req)
                            req.num_computed_tokens = 0
                            if isinstance(req.kv_transfer_params, dict):
                                req.kv_transfer_params["do_remote_prefill"] = False
                        else:
                            kv.cache_blocks(req, req.num_computed_tokens)
                    else:
                        skipped.append(self.waiting.pop())
                        continue
                if (self.max_loras and req.lora_request
                        and req.lora_request.lora_int_id not in active_loras
                        and len(active_loras) >= self.max_loras):
                    skipped.append(self.waiting.pop())
                    continue

                n_ext, load_async = 0, False
                if req.num_computed_tokens == 0:
                    computed_blocks, n_local = kv.get_computed_blocks(req)
                    if self.connector is not None:
                        n_ext, load_async = self.connector.get_num_new_matched_tokens(req, n_local)
                    n_computed = n_local + n_ext
                else:  # resumed after an async KV load
                    computed_blocks, n_local = kv.empty_blocks(), 0
                    n_computed = req.num_computed_tokens
                if load_async:
                    n_new = 0
                else:
                    n_new = req.num_tokens - n_computed
                    if 0 < thr < n_new:
                        n_new = thr
                    if not self.scheduler_config.enable_chunked_prefill and n_new > budget:
                        break
                    n_new = min(n_new, budget)
                    assert n_new > 0
                blocks = kv.allocate_slots(req, n_new + n_ext, n_local, computed_blocks,
                                           num_lookahead_tokens=self.num_lookahead_tokens,
                                           delay_cache_blocks=load_async)
                if blocks is None:
                    break
                if self.connector is not None:
                    self.connector.update_state_after_alloc(req, computed_blocks + blocks, n_ext)
                self.waiting.pop()
                if load_async:
                    req.status = RequestStatus.WAITING_FOR_REMOTE_KVS
                    req.num_computed_tokens = n_computed
                    skipped.append(req)
                    continue
                self.running.append(req)
                if self.log_stats:
                    req.record_event(EngineCoreEventType.SCHEDULED)
                if req.status == RequestStatus.WAITING:
                    scheduled_new.append(req)
                else:
                    scheduled_resumed.append(req)
                if req.lora_request:
                    active_loras.add(req.lora_request.lora_int_id)
                if req.use_structured_output:
                    structured_ids.append(req.request_id)
                new_blocks_of[req.request_id] = kv.get_blocks(req.request_id)
                num_sched[req.request_id] = n_new
                budget -= n_new
                req.status = RequestStatus.RUNNING
                req.num_computed_tokens = n_computed
                if req.num_cached_tokens < 0:
                    req.num_cached_tokens = n_computed
        if skipped:
            self.waiting.prepend_many(skipped)

        total = sum(num_sched.values())
        assert total <= self.max_num_scheduled_tokens
        assert len(self.running) <= self.max_num_running_reqs

        common = [0] * kv.num_groups
        if self.running:
            common = kv.get_num_common_prefix_blocks(self.running[0].request_id, len(self.running))

        new_data = [NewRequestData(
            req_id=r.request_id, prompt_token_ids=r.prompt_token_ids, sampling_params=r.sampling_params,
            pooling_params=r.pooling_params, block_ids=new_blocks_of[r.request_id].get_block_ids(),
            num_computed_tokens=r.num_computed_tokens, lora_request=r.lora_request,
            mm_features=r.mm_features or None) for r in scheduled_new]
        cached = self._make_cached(scheduled_running, scheduled_resumed, num_sched, spec_sched, new_blocks_of)
        self._step += 1
        out = SchedulerOutput(
            scheduled_new_reqs=new_data, scheduled_cached_reqs=cached, num_scheduled_tokens=num_sched,
            total_num_scheduled_tokens=total, scheduled_spec_decode_tokens=spec_sched,
            num_common_prefix_blocks=common, finished_req_ids=self.finished_req_ids,
            structured_output_request_ids=structured_ids,
            preempted_req_ids={r.request_id for r in preempted}, step_id=self._step)
        if self.connector is not None:
            out.kv_connector_metadata = self.connector.build_connector_meta(out)
        events = kv.take_events()
        if events:
            from vllm_b200.core.kv_events import make_batch
            self.kv_event_publisher.publish(make_batch(events))
        # optimistic advance (lets the next step be scheduled before this one finishes)
        for rid, n in num_sched.items():
            self.requests[rid].num_computed_tokens += n
        if self.async_scheduling:
            for r in scheduled_new + scheduled_resumed + scheduled_running:
                # the token sampled by this step is not known yet: reserve its slot
                if r.num_computed_tokens == r.num_tokens + r.num_output_placeholders and \
                        r.pooling_params is None:
                    r.num_output_placeholders += 1
        self.finished_req_ids = set()
        self._last_sched_ms = (time.perf_counter() - t0) * 1e3
        return out

    def _make_cached(self, running, resumed, num_sched, spec_sched, new_blocks_of) -> CachedRequestData:
        c = CachedRequestData()
        pp = self.vllm_config.parallel_config.pipeline_parallel_size > 1
        empty = self.kv_cache_manager.empty_blocks()
        for is_resumed, reqs in ((False, running), (True, resumed)):
            for r in reqs:
                rid = r.request_id
                c.req_ids.append(rid)
                c.resumed_from_preemption.append(is_resumed)
                if is_resumed:
                    c.new_token_ids.append(list(r.all_token_ids))
                elif pp:
                    n = num_sched[rid] - len(spec_sched.get(rid, ()))
                    c.new_token_ids.append(r.all_token_ids[r.num_computed_tokens:r.num_computed_tokens + n])
                else:
                    c.new_token_ids.append([])
                blocks = new_blocks_of.get(rid)
                # (the shared empty result of the allocation fast path: nothing new to ship)
                c.new_block_ids.append(blocks.get_block_ids() if blocks is not None and blocks is not empty else None)
                c.num_computed_tokens.append(r.num_computed_tokens)
                c.num_output_tokens.append(r.num_output_tokens)
        return c

    def _preempt(self, r: Request) -> None:
        self.kv_cache_manager.free(r)
        r.status = RequestStatus.PREEMPTED
        r.num_computed_tokens = 0
        r.num_output_placeholders = 0
        r.spec_token_ids = []
        r.num_preemptions += 1
        self._num_preemptions += 1
        if self.log_stats:
            r.record_event(EngineCoreEventType.PREEMPTED)
        self.waiting.prepend(r)

    # ------------------------------------------------------------- grammar
    def get_grammar_bitmask(self, so: SchedulerOutput):
        if not so.structured_output_request_ids or self.structured_output_manager is None:
            return None
        return self.structured_output_manager.grammar_bitmask(
            self.requests, so.structured_output_request_ids, so.scheduled_spec_decode_tokens)

    # ---------------------------------------------------- update_from_output
    def update_from_output(self, so: SchedulerOutput, mo: ModelRunnerOutput) -> dict[int, EngineCoreOutputs]:
        sampled = mo.sampled_token_ids
        outputs: dict[int, list[EngineCoreOutput]] = {}
        lp_ids = lp_vals = lp_ranks = None
        if mo.logprobs is not None:
            lp_ids, lp_vals, lp_ranks = mo.logprobs.tolists()
        stopped_running: set[str] = set()
        stopped_preempted: list[Request] = []
        lp_offsets: list[int] = []
        if lp_ids is not None:
            acc = 0
            for g in sampled:
                lp_offsets.append(acc)
                acc += len(g)
        # plain generation step (no speculation / logprobs / pooling / prompt logprobs / NaN report): the per-request
        # work below shrinks to "append the token, check the stop conditions, emit the delta"
        plain = (not so.scheduled_spec_decode_tokens and lp_ids is None and not mo.pooler_output
                 and not mo.prompt_logprobs_dict and mo.spec_token_ids is None and not mo.num_nans_in_logits)
        requests, r2i, n_samp, mml = self.requests, mo.req_id_to_index, len(sampled), self.max_model_len
        for rid, n_sched in so.num_scheduled_tokens.items():
            req = requests.get(rid)
            if req is None or req.is_finished():
                # aborted while the step was in flight
                continue
            idx = r2i.get(rid)
            gen = sampled[idx] if idx is not None and idx < n_samp else []
            if plain and not req.use_structured_output:
                if not gen:
                    continue                      # prefill chunk: nothing to report
                if req.num_output_placeholders > 0:
                    req.num_output_placeholders = max(req.num_output_placeholders - len(gen), 0)
                was_preempted = req.status == RequestStatus.PREEMPTED
                stopped, new_ids = False, gen
                for k, tok in enumerate(gen):
                    req.append_output_token_ids(tok)
                    if check_stop(req, mml):
                        stopped, new_ids = True, gen[:k + 1]
                        break
                kv_params = None
                if stopped:
                    kv_params = self._free_request(req)
                    if was_preempted:
                        stopped_preempted.append(req)
                    else:
                        stopped_running.add(rid)
                outs = outputs.get(req.client_index)
                if outs is None:
                    outs = outputs[req.client_index] = []
                nct = req.num_cached_tokens
                outs.append(EngineCoreOutput(
                    request_id=rid, new_token_ids=list(new_ids),
                    finish_reason=int(req.get_finished_reason()) if stopped else None, stop_reason=req.stop_reason,
                    num_cached_tokens=nct if nct > 0 else 0, kv_transfer_params=kv_params,
                    events=req.take_events() if req.events else None))
                continue
            spec = so.scheduled_spec_decode_tokens.get(rid)
            if spec:
                n_draft = len(spec)
                n_acc = max(len(gen) - 1, 0)
                if gen and req.num_computed_tokens > 0:
                    # (a request preempted while this step was in flight was reset to 0: nothing to roll back)
                    req.num_computed_tokens -= n_draft - n_acc
                self._spec_drafts += 1
                self._spec_draft_tokens += n_draft
                self._spec_accepted += n_acc
                if self.spec_dynamic_k:
                    a = self.spec_ema
                    req.spec_accept_ema = (1 - a) * getattr(req, "spec_accept_ema", 1.0) + a * (n_acc / n_draft)
            if req.num_output_placeholders > 0 and gen:
                req.num_output_placeholders -= len(gen)
                req.num_output_placeholders = max(req.num_output_placeholders, 0)
            stopped = False
            new_ids = gen
            # batch-queue path (async scheduling / PP): schedule N+1 may have preempted this request before the
            # output of step N arrived — it then sits in the waiting queue, not in `running`
            was_preempted = req.status == RequestStatus.PREEMPTED
            if gen:
                for k, tok in enumerate(gen):
                    req.append_output_token_ids(tok)
                    if check_stop(req, self.max_model_len):
                        stopped = True
                        new_ids = gen[:k + 1]
                        break
                if req.use_structured_output and self.structured_output_manager is not None:
                    sor = req.structured_output_request
                    sor.accept_tokens(rid, new_ids)
                    if (not stopped and sor.grammar is not None and sor.grammar.is_terminated()
                            and not req.is_finished()):
                        # the constraint is fully satisfied and admits no continuation: finish here instead of
                        # waiting for the model to emit end-of-sequence
                        req.status = RequestStatus.FINISHED_STOPPED
                        req.stop_reason = "grammar_complete"     # (a string: the last token is content, not a stop token)
                        stopped = True
            pooled = None
            if mo.pooler_output and idx is not None and idx < len(mo.pooler_output):
                pooled = mo.pooler_output[idx]
                if pooled is not None:
                    req.status = RequestStatus.FINISHED_STOPPED
                    stopped = True
            new_lp = None
            if lp_ids is not None and req.sampling_params is not None and gen:
                sl = slice(lp_offsets[idx], lp_offsets[idx] + len(new_ids))
                if req.sampling_params.logprobs is not None:
                    new_lp = LogprobsLists(lp_ids[sl], lp_vals[sl], lp_ranks[sl])
            plp = None
            p = mo.prompt_logprobs_dict.get(rid)
            if p is not None:
                a, b, c = p.tolists()
                plp = LogprobsLists(a, b, c)
            if mo.spec_token_ids is not None and idx is not None and not stopped:
                drafts = mo.spec_token_ids[idx]
                if req.use_structured_output and self.structured_output_manager is not None:
                    drafts = req.structured_output_request.validate_tokens(drafts)
                if self.spec_dynamic_k and drafts:
                    # verify about as many drafts as are expected to survive (+1 to keep probing upwards)
                    k = max(1, min(self.num_spec_tokens, int(getattr(req, "spec_accept_ema", 1.0) * self.num_spec_tokens + 1.0)))
                    drafts = drafts[:k]
                req.spec_token_ids = list(drafts)
            if mo.num_nans_in_logits and rid in mo.num_nans_in_logits:
                req.num_nans_in_logits = mo.num_nans_in_logits[rid]
            kv_params = None
            if stopped:
                kv_params = self._free_request(req)
                if was_preempted:
                    stopped_preempted.append(req)
                else:
                    stopped_running.add(rid)
            if new_ids or pooled is not None or stopped or plp is not None:
                outputs.setdefault(req.client_index, []).append(EngineCoreOutput(
                    request_id=rid, new_token_ids=list(new_ids), new_logprobs=new_lp,
                    new_prompt_logprobs=plp, pooling_output=pooled,
                    finish_reason=int(req.get_finished_reason()) if stopped else None,
                    stop_reason=req.stop_reason, num_cached_tokens=max(req.num_cached_tokens, 0),
                    kv_transfer_params=kv_params, events=req.take_events()))
        if stopped_running:
            self.running = [r for r in self.running if r.request_id not in stopped_running]
        for r in stopped_preempted:
            # finished while parked: it must not be rescheduled (its entry in self.requests is gone)
            try:
                self.waiting.remove(r)
            except (ValueError, KeyError):
                pass
        if mo.kv_connector_output is not None and self.connector is not None:
            self._update_from_kv_xfer(mo.kv_connector_output)
        result = {ci: EngineCoreOutputs(outputs=outs, timestamp=time.monotonic())
                  for ci, outs in outputs.items()}
        if self.include_finished_set and self.finished_req_ids_dict:
            for ci, ids in self.finished_req_ids_dict.items():
                eco = result.setdefault(ci, EngineCoreOutputs(timestamp=time.monotonic()))
                eco.finished_requests = sorted(ids)
            self.finished_req_ids_dict.clear()
        stats = self.make_stats(so, mo)
        if stats is not None:
            eco = next(iter(result.values()), None)
            if eco is None:
                eco = result[0] = EngineCoreOutputs(timestamp=time.monotonic())
            eco.scheduler_stats = stats
        return result

    def _update_from_kv_xfer(self, kvo) -> None:
        for rid in getattr(kvo, "failed_recving", None) or ():
            self.connector.mark_recv_failed(rid)
        for rid in getattr(kvo, "finished_recving", None) or ():
            self.connector.mark_recv_finished(rid)
        for rid in getattr(kvo, "finished_sending", None) or ():
            if hasattr(self.connector, "mark_send_finished"):
                self.connector.mark_send_finished(rid)
            req = self.requests.get(rid)
            if req is not None:
                self._free_blocks(req)

    # -------------------------------------------------------------- finishing
    def finish_requests(self, request_ids: Union[str, Iterable[str]], status: RequestStatus) -> None:
        """Abort / externally finish requests (client disconnect, stop strings)."""
        assert RequestStatus.is_finished(status)
        if isinstance(request_ids, str):
            request_ids = (request_ids,)
        run_rm, wait_rm = set(), []
        for rid in set(request_ids):
            req = self.requests.get(rid)
            if req is None or req.is_finished():
                continue
            if req.status == RequestStatus.RUNNING:
                run_rm.add(rid)
            else:
                wait_rm.append(req)
            req.status = status
        if run_rm:
            self.running = [r for r in self.running if r.request_id not in run_rm]
        for r in wait_rm:
            self.waiting.remove(r)
        for rid in list(run_rm) + [r.request_id for r in wait_rm]:
            self._free_request(self.requests[rid])

    def _free_request(self, req: Request) -> Optional[dict]:
        assert req.is_finished()
        delay, kv_params = False, None
        if self.connector is not None:
            delay, kv_params = self.connector.request_finished(req, self.kv_cache_manager.get_block_ids(req.request_id))
        self.finished_req_ids.add(req.request_id)
        if self.finished_req_ids_dict is not None:
            self.finished_req_ids_dict.setdefault(req.client_index, set()).add(req.request_id)
        if not delay:
            self._free_blocks(req)
        return kv_params

    def _free_blocks(self, req: Request) -> None:
        self.kv_cache_manager.free(req)
        self.requests.pop(req.request_id, None)

    # ------------------------------------------------------------------ stats
    def make_stats(self, so: Optional[SchedulerOutput] = None, mo: Optional[ModelRunnerOutput] = None
                   ) -> Optional[SchedulerStats]:
        if not self.log_stats:
            return None
        pc = self.kv_cache_manager.make_prefix_cache_stats()
        s = SchedulerStats(
            num_running_reqs=len(self.running), num_waiting_reqs=len(self.waiting),
            kv_cache_usage=self.kv_cache_manager.usage,
            prefix_cache_queries=pc.queries if pc else 0, prefix_cache_hits=pc.hits if pc else 0,
            num_preemptions=self._num_preemptions, step_counter=self._step,
            spec_num_drafts=self._spec_drafts, spec_num_draft_tokens=self._spec_draft_tokens,
            spec_num_accepted_tokens=self._spec_accepted,
            num_scheduled_tokens=so.total_num_scheduled_tokens if so else 0,
            model_forward_ms=mo.forward_ms if mo else 0.0, schedule_ms=self._last_sched_ms)
        self._num_preemptions = 0
        self._spec_drafts = self._spec_draft_tokens = self._spec_accepted = 0
        return s

    def shutdown(self) -> None:
        self.kv_event_publisher.shutdown()
        if self.connector is not None and hasattr(self.connector, "shutdown"):
            self.connector.shutdown()
