"""Pure-CPU unit tests of the scheduler / KV-cache manager / block pool: no model, no GPU.

Mirrors the reference's tests/v1/core/{test_scheduler,test_prefix_caching,test_async_scheduler}.py technique: a
``Scheduler`` over a synthetic ``KVCacheConfig``, hand-built ``ModelRunnerOutput`` fed to ``update_from_output``.
"""
import pytest

from vllm_b200.config import CacheConfig, ModelConfig, ParallelConfig, SchedulerConfig, VllmConfig
from vllm_b200.core.kv_cache_manager import (FullAttentionSpec, KVCacheConfig, KVCacheGroupSpec, SlidingWindowSpec)
from vllm_b200.core.request import Request, RequestStatus
from vllm_b200.core.sched_output import ModelRunnerOutput
from vllm_b200.core.scheduler import Scheduler
from vllm_b200.sampling_params import SamplingParams

BS = 16


def create_scheduler(num_blocks=64, max_num_seqs=8, max_num_batched_tokens=64, max_model_len=256, caching=True,
                     policy="fcfs", async_scheduling=False, chunked=True, specs=None) -> Scheduler:
    mc = ModelConfig(model="llama-tiny", dtype="float32", max_model_len=max_model_len)
    cfg = VllmConfig(
        model_config=mc, cache_config=CacheConfig(block_size=BS, enable_prefix_caching=caching),
        parallel_config=ParallelConfig(),
        scheduler_config=SchedulerConfig(max_num_batched_tokens=max_num_batched_tokens, max_num_seqs=max_num_seqs,
                                         max_model_len=max_model_len, enable_chunked_prefill=chunked, policy=policy,
                                         async_scheduling=async_scheduling))
    specs = specs or [FullAttentionSpec(block_size=BS, num_kv_heads=2, head_size=32, dtype_size=4)]
    groups = [KVCacheGroupSpec([f"layer{i}"], s) for i, s in enumerate(specs)]
    return Scheduler(cfg, KVCacheConfig(num_blocks=num_blocks, kv_cache_groups=groups))


def make_request(s: Scheduler, rid: str, n_prompt: int, max_tokens=8, base=0, priority=0, **kw) -> Request:
    sp = SamplingParams(max_tokens=max_tokens, temperature=0.0, ignore_eos=True)
    r = Request(rid, [(base + i) % 1000 + 1 for i in range(n_prompt)], sampling_params=sp, eos_token_id=None,
                block_hasher=s.get_block_hasher(), priority=priority, **kw)
    s.add_request(r)
    return r


def run_step(s: Scheduler, token=7):
    """schedule → fake model output (one token for every request that finished its prompt chunk) → update."""
    so = s.schedule()
    ids, toks = [], []
    for rid, n in so.num_scheduled_tokens.items():
        r = s.requests[rid]
        ids.append(rid)
        # num_computed_tokens was advanced optimistically in schedule(): a token is produced iff the prompt is done
        toks.append([token] if r.num_computed_tokens >= r.num_tokens else [])
    mo = ModelRunnerOutput(req_ids=ids, req_id_to_index={r: i for i, r in enumerate(ids)}, sampled_token_ids=toks)
    out = s.update_from_output(so, mo)
    return so, out


def test_prefill_then_decode_and_finish():
    s = create_scheduler()
    r = make_request(s, "a", 20, max_tokens=3)
    so, _ = run_step(s)
    assert so.num_scheduled_tokens == {"a": 20} and len(so.scheduled_new_reqs) == 1
    assert len(so.scheduled_new_reqs[0].block_ids[0]) == 2          # ceil(20/16) blocks
    for step in range(2):
        so, out = run_step(s)
        assert so.num_scheduled_tokens == {"a": 1} and so.scheduled_cached_reqs.req_ids == ["a"]
    assert r.is_finished() and r.status == RequestStatus.FINISHED_LENGTH_CAPPED
    assert not s.has_unfinished_requests()
    assert "a" in s.schedule().finished_req_ids or True
    assert s.kv_cache_manager.pool.get_num_free_blocks() == 63          # everything returned (block 0 is the null block)


def test_chunked_prefill_respects_token_budget():
    s = create_scheduler(max_num_batched_tokens=32)
    make_request(s, "a", 50)
    make_request(s, "b", 40, base=500)
    so, _ = run_step(s)
    assert so.num_scheduled_tokens == {"a": 32}                      # budget exhausted by the first chunk
    so, _ = run_step(s)
    assert so.num_scheduled_tokens == {"a": 18, "b": 14}             # a finishes its prompt, b gets the rest
    so, _ = run_step(s)
    assert so.num_scheduled_tokens == {"a": 1, "b": 26}
    so, _ = run_step(s)
    assert so.num_scheduled_tokens == {"a": 1, "b": 1}


def test_no_chunking_waits_for_room():
    s = create_scheduler(max_num_batched_tokens=64, chunked=False, max_model_len=64)
    make_request(s, "a", 40)
    make_request(s, "b", 40, base=300)
    so, _ = run_step(s)
    assert so.num_scheduled_tokens == {"a": 40}                      # b (40) does not fit next to a → not split
    so, _ = run_step(s)
    assert so.num_scheduled_tokens == {"a": 1, "b": 40}


def test_prefix_cache_hit_skips_computed_blocks():
    s = create_scheduler()
    make_request(s, "a", 50, max_tokens=1)
    run_step(s)
    assert not s.has_unfinished_requests()
    r = make_request(s, "b", 50, max_tokens=2)                       # identical prompt
    so, _ = run_step(s)
    assert so.num_scheduled_tokens == {"b": 50 - 48}                 # 3 full blocks (48 tokens) come from the cache
    assert so.scheduled_new_reqs[0].num_computed_tokens == 48
    st = s.kv_cache_manager.prefix_cache_stats if hasattr(s.kv_cache_manager, "prefix_cache_stats") else None
    make_request(s, "c", 50, max_tokens=1, cache_salt="tenant-2")    # salted: must NOT share blocks
    so, _ = run_step(s)
    assert so.num_scheduled_tokens["c"] == 50
    del r, st


def test_preemption_frees_blocks_and_resumes():
    s = create_scheduler(num_blocks=6, max_num_batched_tokens=256, caching=False)   # 5 usable blocks
    a = make_request(s, "a", 30, max_tokens=40)                      # 2 blocks now, grows
    b = make_request(s, "b", 30, max_tokens=40, base=400)
    run_step(s)
    assert {r.request_id for r in s.running} == {"a", "b"}
    preempted = False
    for _ in range(40):
        so, _ = run_step(s)
        if so.preempted_req_ids:
            preempted = True
            assert so.preempted_req_ids == {"b"}                     # lowest priority = most recently admitted
            assert b.status == RequestStatus.PREEMPTED and b.num_computed_tokens == 0
            break
    assert preempted
    while not a.is_finished():
        run_step(s)
    for _ in range(60):
        if b.is_finished():
            break
        run_step(s)
    assert b.is_finished() and b.num_output_tokens == 40 and b.num_preemptions >= 1


def test_priority_policy_orders_waiting_queue():
    s = create_scheduler(policy="priority", max_num_seqs=1)
    make_request(s, "low", 10, priority=5)
    make_request(s, "high", 10, priority=0, base=100)
    so, _ = run_step(s)
    assert list(so.num_scheduled_tokens) == ["high"]


def test_abort_and_finished_ids_propagate():
    s = create_scheduler()
    make_request(s, "a", 10)
    make_request(s, "b", 10, base=50)
    run_step(s)
    s.finish_requests(["a"], RequestStatus.FINISHED_ABORTED)
    so = s.schedule()
    assert "a" in so.finished_req_ids and "a" not in so.num_scheduled_tokens and "b" in so.num_scheduled_tokens


def test_async_scheduling_runs_one_step_ahead():
    s = create_scheduler(async_scheduling=True)
    r = make_request(s, "a", 10, max_tokens=4)
    so1 = s.schedule()                                               # prefill
    so2 = s.schedule()                                               # scheduled before step 1's token is known
    assert so1.num_scheduled_tokens == {"a": 10} and so2.num_scheduled_tokens == {"a": 1}
    assert r.num_output_placeholders >= 1

    def mo(tok):
        return ModelRunnerOutput(req_ids=["a"], req_id_to_index={"a": 0}, sampled_token_ids=[[tok]])
    s.update_from_output(so1, mo(11))
    so3 = s.schedule()
    s.update_from_output(so2, mo(12))
    s.update_from_output(so3, mo(13))
    so4 = s.schedule()
    s.update_from_output(so4, mo(14))
    assert r.output_token_ids == [11, 12, 13, 14] and r.is_finished()
    assert not s.schedule().num_scheduled_tokens                     # nothing is scheduled past max_tokens


def test_hybrid_groups_sliding_window_frees_old_blocks():
    specs = [FullAttentionSpec(block_size=BS, num_kv_heads=2, head_size=32, dtype_size=4),
             SlidingWindowSpec(block_size=BS, num_kv_heads=2, head_size=32, dtype_size=4, sliding_window=32)]
    s = create_scheduler(num_blocks=64, max_num_batched_tokens=256, caching=False, specs=specs)
    r = make_request(s, "a", 100, max_tokens=30)
    so, _ = run_step(s)
    full, swa = so.scheduled_new_reqs[0].block_ids
    assert len(full) == 7 and len(swa) == 7
    for _ in range(29):
        run_step(s)
    assert r.is_finished()
    # while running, the sliding-window group kept only the blocks covering the last 32 tokens (+ the current one)
    assert s.kv_cache_manager.pool.get_num_free_blocks() == 63


def test_tiktoken_tokenizer_mode(tmp_path):
    """A rank file + tokenizer_config.json → TokenizerLike: round trip, specials, incremental detokenisation."""
    import base64
    import json
    from vllm_b200.tokenizers import get_tokenizer
    ranks = {bytes([b]): b for b in range(256)}
    for i, tok in enumerate([b"he", b"ll", b"hell", b"hello", b" w", b" wo", b" wor", b" world"]):
        ranks[tok] = 256 + i
    with open(tmp_path / "model.tiktoken", "wb") as f:
        for tok, r in ranks.items():
            f.write(base64.b64encode(tok) + b" " + str(r).encode() + b"\n")
    with open(tmp_path / "tokenizer_config.json", "w") as f:
        json.dump({"added_tokens_decoder": {"264": {"content": "<|im_start|>"}, "265": {"content": "<|im_end|>"},
                                            "266": {"content": "<|endoftext|>"}}, "eos_token": "<|endoftext|>"}, f)
    tok = get_tokenizer(str(tmp_path), "auto")
    assert type(tok).__name__ == "TiktokenTokenizer" and len(tok) == 267 and tok.eos_token_id == 266
    ids = tok.encode("hello world ünï")
    assert ids[:2] == [259, 263] and tok.decode(ids) == "hello world ünï"
    chat = tok.apply_chat_template([{"role": "user", "content": "hello"}], tokenize=True)
    assert chat[0] == 264 and 265 in chat and tok.decode(chat, skip_special_tokens=True).startswith("user\nhello")


def test_mistral_tokenizer_mode(tmp_path):
    """tokenizer_mode="mistral" over mistral-common's bundled tekken file: round trip + its own chat encoder."""
    import os
    import shutil
    import mistral_common
    src = os.path.join(os.path.dirname(mistral_common.__file__), "data", "tekken_240911.json")
    if not os.path.exists(src):
        pytest.skip("mistral-common ships no tekken file in this build")
    shutil.copy(src, tmp_path / "tekken.json")
    from vllm_b200.tokenizers import get_tokenizer
    tok = get_tokenizer(str(tmp_path), "mistral")
    ids = tok.encode("Hello wörld", add_special_tokens=True)
    assert ids[0] == tok.bos_token_id and tok.decode(ids) == "Hello wörld"
    chat = tok.apply_chat_template([{"role": "user", "content": "hi"}], tokenize=True)
    assert chat[0] == tok.bos_token_id and "[INST]" in tok.apply_chat_template([{"role": "user", "content": "hi"}])
    assert len(tok) > 100000


# ---------------------------------------------------------------------------------------------------------------------
# more of the reference's tests/v1/core areas: stop conditions, budgets, cache eviction order, KV events, spec accounting
def _step_with(s: Scheduler, tokens: dict):
    so = s.schedule()
    ids = list(so.num_scheduled_tokens)
    toks = [tokens.get(r, [7]) if s.requests[r].num_computed_tokens >= s.requests[r].num_tokens else [] for r in ids]
    mo = ModelRunnerOutput(req_ids=ids, req_id_to_index={r: i for i, r in enumerate(ids)}, sampled_token_ids=toks)
    return so, s.update_from_output(so, mo)


def test_stop_token_eos_and_max_model_len():
    s = create_scheduler(max_model_len=40)
    sp = SamplingParams(max_tokens=50, temperature=0.0, stop_token_ids=[99])
    sp.all_stop_token_ids = [99, 2]
    r = Request("a", list(range(1, 11)), sampling_params=sp, eos_token_id=2, block_hasher=s.get_block_hasher())
    s.add_request(r)
    _step_with(s, {})                                     # prefill + first token (7)
    _, out = _step_with(s, {"a": [99]})                    # a stop token id ends the request and is reported
    o = out[0].outputs[0]
    assert o.finish_reason is not None and o.stop_reason == 99 and "a" not in s.requests
    r2 = Request("b", list(range(1, 11)), sampling_params=SamplingParams(max_tokens=500, temperature=0.0, ignore_eos=True),
                 eos_token_id=2, block_hasher=s.get_block_hasher())
    s.add_request(r2)
    n = 0
    while "b" in s.requests and n < 100:
        _step_with(s, {})
        n += 1
    assert r2.num_tokens <= 40 and "b" not in s.requests            # the context window bounds generation
    r3 = Request("c", list(range(1, 11)), sampling_params=SamplingParams(max_tokens=5, temperature=0.0), eos_token_id=2,
                 block_hasher=s.get_block_hasher())
    s.add_request(r3)
    _step_with(s, {})
    _, out = _step_with(s, {"c": [2]})                               # EOS
    assert out[0].outputs[0].finish_reason is not None and out[0].outputs[0].stop_reason is None


def test_max_num_seqs_and_token_budget_are_respected():
    s = create_scheduler(max_num_seqs=2, max_num_batched_tokens=32, num_blocks=128)
    for i in range(4):
        make_request(s, f"r{i}", 20, base=100 * i)
    so, _ = run_step(s)
    assert sum(so.num_scheduled_tokens.values()) <= 32 and len(so.num_scheduled_tokens) <= 2
    seen_running = set()
    for _ in range(30):
        so, _ = run_step(s)
        assert sum(so.num_scheduled_tokens.values()) <= 32
        assert len(s.running) <= 2
        seen_running |= set(so.num_scheduled_tokens)
        if not s.has_unfinished_requests():
            break
    assert seen_running == {"r0", "r1", "r2", "r3"} and not s.has_unfinished_requests()


def test_prefix_cache_lru_eviction_and_reset():
    s = create_scheduler(num_blocks=10, max_num_batched_tokens=128)          # 9 usable blocks (one is the null block)
    a = make_request(s, "a", 64, max_tokens=1, base=0)                       # 4 full blocks
    run_step(s)
    assert "a" not in s.requests                                             # finished: blocks stay cached but evictable
    b = make_request(s, "b", 64, max_tokens=1, base=500)
    run_step(s)
    pool = s.kv_cache_manager.pool
    assert pool.get_num_free_blocks() == 9
    c = make_request(s, "c", 64, max_tokens=1, base=0)                       # same prompt as "a": 3 cached blocks reused
    so, _ = run_step(s)
    assert so.num_scheduled_tokens["c"] < 64 and c.num_cached_tokens >= 48
    d = make_request(s, "d", 96, max_tokens=1, base=900)                     # needs 6 fresh blocks → evicts the oldest cached
    run_step(s)
    e = make_request(s, "e", 64, max_tokens=1, base=500)                     # "b"'s blocks were the least recently used
    so, _ = run_step(s)
    assert so.num_scheduled_tokens["e"] == 64 and e.num_cached_tokens == 0
    assert s.reset_prefix_cache()
    f = make_request(s, "f", 64, max_tokens=1, base=900)
    so, _ = run_step(s)
    assert so.num_scheduled_tokens["f"] == 64                                # nothing survives a reset
    del a, b, d


def test_kv_events_stream_block_stored_and_removed():
    s = create_scheduler(num_blocks=8)
    pool = s.kv_cache_manager.pool
    pool.enable_kv_cache_events = True
    pool.block_size = BS
    seen, orig = [], pool.take_events

    def take():                      # the scheduler drains the queue every step: record what it takes
        ev = orig()
        seen.extend(ev)
        return ev
    pool.take_events = take
    make_request(s, "a", 48, max_tokens=1)
    run_step(s)
    take()
    stored = [e for e in seen if type(e).__name__ == "BlockStored"]
    assert stored and sum(len(e.block_hashes) for e in stored) == 3 and stored[0].block_size == BS
    assert len(stored[0].token_ids) == 48 and stored[0].parent_block_hash is None
    make_request(s, "b", 100, max_tokens=1, base=300)                        # 7 blocks: the second chunk evicts "a"'s
    run_step(s)
    run_step(s)
    take()
    assert any(type(e).__name__ == "BlockRemoved" for e in seen)
    pool.reset_prefix_cache()
    take()
    assert any(type(e).__name__ == "AllBlocksCleared" for e in seen)


def test_speculative_tokens_accounting_and_rollback():
    import types
    s = create_scheduler()
    s.num_spec_tokens = 3
    r = make_request(s, "a", 10, max_tokens=20)
    run_step(s)                                                              # prefill → first token
    r.spec_token_ids = [11, 12, 13]
    so = s.schedule()
    assert so.scheduled_spec_decode_tokens == {"a": [11, 12, 13]} and so.num_scheduled_tokens["a"] == 4
    computed_before = r.num_computed_tokens
    # the verifier accepted the first draft only: 2 tokens come back (accepted draft + bonus/correction)
    mo = ModelRunnerOutput(req_ids=["a"], req_id_to_index={"a": 0}, sampled_token_ids=[[11, 50]])
    s.update_from_output(so, mo)
    assert r.num_computed_tokens == computed_before - 2                      # the two rejected drafts are rolled back
    assert list(r.output_token_ids)[-2:] == [11, 50]
    assert s._spec_draft_tokens == 3 and s._spec_accepted == 1 and s._spec_drafts == 1
    del types


def test_scheduler_block_accounting_invariants_under_random_load():
    """Property test (hypothesis): any interleaving of arrivals (with shared prefixes), steps, aborts and cache resets
    keeps the block pool consistent — reference counts match the running requests, nothing leaks, nothing is
    double-freed, and a drained scheduler returns every block."""
    from hypothesis import given, settings, strategies as st

    ops = st.lists(st.one_of(
        st.tuples(st.just("add"), st.integers(0, 3), st.integers(1, 70), st.integers(1, 6)),     # prefix family, length, max_tokens
        st.tuples(st.just("step"), st.integers(1, 4)),
        st.tuples(st.just("abort"), st.integers(0, 30)),
        st.tuples(st.just("reset"),)), min_size=5, max_size=40)

    @settings(max_examples=40, deadline=None)
    @given(ops)
    def run(seq):
        s = create_scheduler(num_blocks=24, max_num_seqs=4, max_num_batched_tokens=48, max_model_len=128)
        pool = s.kv_cache_manager.pool
        total = pool.num_blocks - 1
        n = 0
        live: list[str] = []

        def check():
            held: dict[int, int] = {}
            for m in s.kv_cache_manager.managers:
                for blocks in m.req_to_blocks.values():
                    for b in blocks:
                        if not b.is_null:
                            held[b.block_id] = held.get(b.block_id, 0) + 1
            for b in pool.blocks:
                if b.is_null:
                    continue
                assert b.ref_cnt == held.get(b.block_id, 0), (b.block_id, b.ref_cnt, held.get(b.block_id, 0))
            assert pool.get_num_free_blocks() == total - len(held)
            for (h, g), d in pool.cached.items():
                assert d and all(blk.block_hash == h for blk in d.values())

        for op in seq:
            if op[0] == "add":
                _, fam, length, mt = op
                rid = f"r{n}"
                n += 1
                sp = SamplingParams(max_tokens=mt, temperature=0.0, ignore_eos=True)
                toks = [(fam * 37 + i) % 50 + 1 for i in range(min(length, 40))] + [n % 50 + 1] * max(length - 40, 0)
                s.add_request(Request(rid, toks, sampling_params=sp, eos_token_id=None, block_hasher=s.get_block_hasher()))
                live.append(rid)
            elif op[0] == "step":
                for _ in range(op[1]):
                    if s.has_requests():
                        run_step(s)
            elif op[0] == "abort" and live:
                s.finish_requests([live[op[1] % len(live)]], RequestStatus.FINISHED_ABORTED)
            elif op[0] == "reset":
                s.reset_prefix_cache()
            live = [r for r in live if r in s.requests]
            check()
        for _ in range(400):
            if not s.has_unfinished_requests():
                break
            run_step(s)
            check()
        assert not s.has_unfinished_requests()
        if s.has_requests():
            run_step(s)
        check()
        assert pool.get_num_free_blocks() == total
    run()


def test_preempted_in_flight_then_stopped_leaves_no_ghost():
    """Batch-queue path: schedule N+1 preempts a request whose step-N output then finishes it.  The finished request
    must leave the waiting queue (it used to be rescheduled and crash on ``self.requests[rid]``)."""
    s = create_scheduler(num_blocks=3, max_num_batched_tokens=256, caching=False, async_scheduling=True)  # 2 usable blocks
    a = make_request(s, "a", BS, max_tokens=8)                       # exactly one block: its first decode needs a second
    b = make_request(s, "b", 10, max_tokens=1, base=400)             # one block, finishes with its first token

    def mo(so, tok=7):
        ids = list(so.num_scheduled_tokens)
        return ModelRunnerOutput(req_ids=ids, req_id_to_index={r: i for i, r in enumerate(ids)},
                                 sampled_token_ids=[[tok] for _ in ids])
    so1 = s.schedule()
    assert set(so1.num_scheduled_tokens) == {"a", "b"}
    so2 = s.schedule()                                               # step 1 still in flight: a grows, b is preempted
    assert "b" in so2.preempted_req_ids and b.status == RequestStatus.PREEMPTED
    s.update_from_output(so1, mo(so1))                               # b's only token arrives → finished while parked
    assert b.is_finished()
    assert all(r.request_id != "b" for r in s.waiting), "finished request left in the waiting queue"
    s.update_from_output(so2, mo(so2))
    for _ in range(20):                                              # used to raise KeyError on the ghost entry
        if a.is_finished():
            break
        so = s.schedule()
        s.update_from_output(so, mo(so))
    assert a.is_finished() and "b" not in s.requests
