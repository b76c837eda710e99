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
