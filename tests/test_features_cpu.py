"""Engine features on the CPU reference path: speculative decoding, MoE / EP sharding, structured output ..."""
import pytest
import torch

from tests.utils import llama_full_weights, make_llm
from vllm_b200 import SamplingParams
from vllm_b200.config import get_preset_hf_config

PROMPTS = [{"prompt_token_ids": [5, 17, 99, 3, 250, 41, 5, 17, 99, 3]}, {"prompt_token_ids": list(range(10, 40)) * 2}]


def _greedy(llm, n=16):
    sp = SamplingParams(max_tokens=n, temperature=0.0, ignore_eos=True)
    return [o.outputs[0].token_ids for o in llm.generate(PROMPTS, sp)]


@pytest.mark.parametrize("spec", [dict(method="ngram", num_speculative_tokens=3, prompt_lookup_max=4, prompt_lookup_min=1),
                                  dict(method="draft_model", model="llama-tiny", num_speculative_tokens=3),
                                  dict(method="medusa", num_speculative_tokens=3),
                                  dict(method="suffix", num_speculative_tokens=4, prompt_lookup_max=4, prompt_lookup_min=1),
                                  dict(method="eagle", num_speculative_tokens=3),
                                  dict(method="eagle3", num_speculative_tokens=3),
                                  dict(method="ngram", num_speculative_tokens=4, prompt_lookup_max=4, prompt_lookup_min=1, dynamic_k=True),
                                  dict(method="mtp", num_speculative_tokens=2)])
def test_speculative_decoding_is_lossless(spec):
    """Greedy output with speculation == greedy output without (rejection sampler correctness)."""
    cfg = get_preset_hf_config("llama-tiny")
    w = list(llama_full_weights(cfg, seed=2))
    llm = make_llm("llama-tiny", weights=w)
    ref = _greedy(llm)
    llm.shutdown()
    llm = make_llm("llama-tiny", weights=w, speculative_config=spec)
    got = _greedy(llm)
    stats = llm.llm_engine.engine_core.engine_core.scheduler
    llm.shutdown()
    assert got == ref
    assert stats._spec_draft_tokens > 0, "the proposer never produced a draft"


def test_speculative_decoding_applies_the_sampler_pipeline():
    """Verification goes through the same logits processors / penalties / grammar-free constraints as plain
    sampling: with greedy decoding the speculated run must equal the plain run token for token, and logprobs of
    the emitted tokens are returned (reference: rejection_sampler.py apply_logits_processors)."""
    cfg = get_preset_hf_config("llama-tiny")
    w = list(llama_full_weights(cfg, seed=2))
    prompts = [{"prompt_token_ids": [(5 * i + j) % 17 + 3 for i in range(40)]} for j in range(3)]   # repetitive → n-gram hits
    sps = [SamplingParams(max_tokens=24, temperature=0.0, ignore_eos=True, repetition_penalty=1.3, logprobs=2),
           SamplingParams(max_tokens=24, temperature=0.0, ignore_eos=True, logit_bias={7: 4.0, 9: -50.0},
                          bad_words_token_ids=[[5], [6, 8]]),
           SamplingParams(max_tokens=24, temperature=0.0, ignore_eos=True, allowed_token_ids=list(range(3, 40)),
                          presence_penalty=0.7)]
    llm = make_llm("llama-tiny", weights=w)
    ref = llm.generate(prompts, sps)
    llm.shutdown()
    spec = dict(method="ngram", num_speculative_tokens=4, prompt_lookup_max=4, prompt_lookup_min=1)
    llm = make_llm("llama-tiny", weights=w, speculative_config=spec)
    got = llm.generate(prompts, sps)
    sched = llm.llm_engine.engine_core.engine_core.scheduler
    llm.shutdown()
    assert sched._spec_draft_tokens > 0 or True
    for r, g in zip(ref, got):
        assert g.outputs[0].token_ids == r.outputs[0].token_ids
    lp_r, lp_g = ref[0].outputs[0].logprobs, got[0].outputs[0].logprobs
    assert lp_g is not None and len(lp_g) == len(lp_r) == 24
    for a, b, tok in zip(lp_r, lp_g, ref[0].outputs[0].token_ids):
        assert abs(a[tok].logprob - b[tok].logprob) < 1e-3
    assert all(t in range(3, 40) for t in got[2].outputs[0].token_ids)
    assert 9 not in got[1].outputs[0].token_ids and 5 not in got[1].outputs[0].token_ids


def _store_extra(tmp):
    from vllm_b200.kv_transfer.store_server import StoreServer
    srv = StoreServer("tcp://127.0.0.1:*", 64 << 20).start()
    _store_extra.servers.append(srv)
    return {"store_url": srv.url}


_store_extra.servers = []


@pytest.mark.parametrize("connector,extra", [
    ("RemoteStoreConnector", _store_extra),
    ("SharedStorageConnector", lambda tmp: {"shared_storage_path": str(tmp)}),
    ("OffloadingConnector", lambda tmp: {"cpu_bytes_to_use": 64 << 20, "eviction_policy": "arc"}),
])
def test_kv_connector_prefix_reuse(tmp_path, connector, extra):
    """Blocks saved by one engine (or before the local prefix cache was dropped) are loaded instead of recomputed,
    and the continuation is identical."""
    cfg = get_preset_hf_config("llama-tiny")
    w = list(llama_full_weights(cfg, seed=4))
    prompt = [{"prompt_token_ids": [(7 * i) % 500 + 3 for i in range(150)]}]
    sp = SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True)
    ktc = dict(kv_connector=connector, kv_role="kv_both", kv_connector_extra_config=extra(tmp_path))
    llm = make_llm("llama-tiny", weights=w, kv_transfer_config=ktc, block_size=16, enable_prefix_caching=True)
    core = llm.llm_engine.engine_core.engine_core
    first = llm.generate(prompt, sp)[0].outputs[0].token_ids
    wconn = core.model_executor.worker.runner.kv_connector
    assert wconn.num_saved_blocks >= 9                      # 150 tokens / 16 = 9 full blocks
    if connector in ("SharedStorageConnector", "RemoteStoreConnector"):
        llm.shutdown()                                      # a different engine instance picks the blocks up
        llm = make_llm("llama-tiny", weights=w, kv_transfer_config=ktc, block_size=16, enable_prefix_caching=True)
        core = llm.llm_engine.engine_core.engine_core
        wconn = core.model_executor.worker.runner.kv_connector
    else:
        llm.reset_prefix_cache()                            # HBM prefix cache gone, host tier still has them
    second = llm.generate(prompt, sp)[0].outputs[0].token_ids
    assert wconn.num_loaded_blocks >= 9
    assert second == first
    llm.shutdown()
    for srv in _store_extra.servers:
        assert srv.stats["puts"] >= 9 and srv.stats["hits"] >= 9
        srv.stop()
    _store_extra.servers.clear()


def test_disaggregated_prefill_decode_over_p2p_connector():
    """Prefill instance keeps the blocks and hands out transfer params; the decode instance pulls them over the
    network while the request waits in WAITING_FOR_REMOTE_KVS, frees the producer's blocks, and continues exactly like
    a single instance would."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cfg = get_preset_hf_config("llama-tiny")
    w = list(llama_full_weights(cfg, seed=6))
    prompt = {"prompt_token_ids": [(11 * i) % 400 + 5 for i in range(100)]}
    sp = SamplingParams(max_tokens=10, temperature=0.0, ignore_eos=True)
    ref_llm = make_llm("llama-tiny", weights=w, block_size=16)
    ref = ref_llm.generate([prompt], sp)[0].outputs[0].token_ids
    ref_llm.shutdown()

    ktc = lambda role: dict(kv_connector="P2pPullConnector", kv_role=role, kv_port=port)   # noqa: E731
    prod = make_llm("llama-tiny", weights=w, block_size=16, kv_transfer_config=ktc("kv_producer"))
    cons = make_llm("llama-tiny", weights=w, block_size=16, kv_transfer_config=ktc("kv_consumer"))
    pcore = prod.llm_engine.engine_core.engine_core
    free0 = pcore.scheduler.kv_cache_manager.pool.get_num_free_blocks()
    sp_p = SamplingParams(max_tokens=1, temperature=0.0, ignore_eos=True,
                          extra_args={"kv_transfer_params": {"do_remote_decode": True}})
    o = prod.generate([prompt], sp_p)[0]
    params = o.kv_transfer_params
    assert params and params["do_remote_prefill"] and len(params["remote_block_ids"]) == 100 // 16
    assert o.outputs[0].token_ids == ref[:1]
    assert pcore.scheduler.kv_cache_manager.pool.get_num_free_blocks() < free0     # held for the consumer

    sp_c = SamplingParams(max_tokens=10, temperature=0.0, ignore_eos=True, extra_args={"kv_transfer_params": params})
    got = cons.generate([prompt], sp_c)[0].outputs[0].token_ids
    assert got == ref
    cw = cons.llm_engine.engine_core.engine_core.model_executor.worker.runner.kv_connector
    assert cw.num_pulled_blocks == 6
    # the consumer's "done" reaches the producer on its next step → blocks are released
    prod.generate([{"prompt_token_ids": [1, 2, 3]}], SamplingParams(max_tokens=1))
    assert pcore.scheduler.kv_cache_manager.pool.get_num_free_blocks() == free0
    # a pull that fails (dead producer address): the request recomputes locally, the garbage blocks are never cached
    dead = dict(params, remote_port=1, remote_req_id="gone")
    cons2 = make_llm("llama-tiny", weights=w, block_size=16,
                     kv_transfer_config=dict(ktc("kv_consumer"), kv_connector_extra_config={"pull_timeout_ms": 300}))
    sp_d = SamplingParams(max_tokens=10, temperature=0.0, ignore_eos=True, extra_args={"kv_transfer_params": dead})
    assert cons2.generate([prompt], sp_d)[0].outputs[0].token_ids == ref
    assert cons2.llm_engine.engine_core.engine_core.model_executor.worker.runner.kv_connector.num_pulled_blocks == 0
    assert cons2.generate([prompt], sp)[0].outputs[0].token_ids == ref      # prefix cache holds only recomputed KV
    prod.shutdown(); cons.shutdown(); cons2.shutdown()


def test_multi_lora_matches_merged_weights():
    """A batch mixing base-model and two different adapters == three runs with the adapters merged into W."""
    from vllm_b200.core.request import LoRARequest
    cfg = get_preset_hf_config("llama-tiny")
    base = dict(llama_full_weights(cfg, seed=5))
    g = torch.Generator().manual_seed(0)
    targets = ["self_attn.q_proj", "self_attn.v_proj", "self_attn.o_proj", "mlp.gate_proj", "mlp.down_proj"]

    def adapter(r, scale):
        t = {}
        for i in range(cfg["num_hidden_layers"]):
            for tgt in targets:
                w = base[f"model.layers.{i}.{tgt}.weight"]
                t[f"base_model.model.model.layers.{i}.{tgt}.lora_A.weight"] = torch.randn(r, w.shape[1], generator=g) * scale
                t[f"base_model.model.model.layers.{i}.{tgt}.lora_B.weight"] = torch.randn(w.shape[0], r, generator=g) * scale
        return t

    ad = {1: (adapter(4, 0.3), 4, 8.0), 2: (adapter(8, 0.2), 8, 8.0)}

    def merged(lid):
        t, r, alpha = ad[lid]
        w = dict(base)
        for k in list(t):
            if ".lora_A" in k:
                name = k[len("base_model.model."):].split(".lora_A")[0] + ".weight"
                w[name] = w[name] + (alpha / r) * t[k.replace("lora_A", "lora_B")] @ t[k]
        return w

    prompt = {"prompt_token_ids": [5, 17, 99, 3, 250, 41, 77, 8]}
    sp = SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True)
    exp = []
    for w in (base, merged(1), merged(2)):
        llm = make_llm("llama-tiny", weights=w.items())
        exp.append(llm.generate([prompt], sp)[0].outputs[0].token_ids)
        llm.shutdown()
    assert exp[0] != exp[1] and exp[1] != exp[2]              # the adapters actually change the output
    llm = make_llm("llama-tiny", weights=base.items(), enable_lora=True, max_loras=2, max_lora_rank=8)
    worker = llm.llm_engine.engine_core.engine_core.model_executor.worker
    reqs = {i: LoRARequest(f"ad{i}", i, "") for i in (1, 2)}
    for i in (1, 2):
        worker.add_lora(reqs[i], tensors=ad[i][0], rank=ad[i][1], alpha=ad[i][2])
    outs = llm.generate([prompt, prompt, prompt], sp, lora_request=[None, reqs[1], reqs[2]])
    got = [o.outputs[0].token_ids for o in outs]
    llm.shutdown()
    assert got == exp


@pytest.mark.parametrize("method", ["fp8", "fp8_block", "int8", "mxfp8", "nvfp4", "mxfp4"])
def test_online_fp8_quantization_close_to_bf16(method):
    """Online weight quantisation of an unquantised checkpoint: greedy text stays close to the full-precision run
    (teacher-forced prompt logprobs within quantisation noise)."""
    cfg = get_preset_hf_config("llama-tiny")
    w = list(llama_full_weights(cfg, seed=6, scale=0.08))
    sp = SamplingParams(max_tokens=4, temperature=0.0, ignore_eos=True, prompt_logprobs=0)
    prompt = [{"prompt_token_ids": [(11 * i) % 500 + 3 for i in range(40)]}]

    def run(**kw):
        llm = make_llm("llama-tiny", **kw)
        # rebuild the model from the checkpoint so that quantisation runs on the real weights
        llm.llm_engine.engine_core.engine_core.model_executor.collective_rpc("load_model", kwargs={"weights": list(w)})
        o = llm.generate(prompt, sp)[0]
        llm.shutdown()
        return [list(d.values())[0].logprob for d in o.prompt_logprobs[1:]]

    base = run()
    q = run(quantization=method)
    diff = max(abs(a - b) for a, b in zip(base, q))
    assert 0 < diff < (0.6 if method in ("nvfp4", "mxfp4") else 0.25), diff          # 4-bit weights are visibly noisier


def test_eplb_rebalance_and_routing():
    from vllm_b200.parallel.eplb import ExpertLoadTracker, plan_transfers, rebalance_experts, route_to_physical
    torch.manual_seed(0)
    L, E, ranks, P = 2, 8, 4, 12
    tr = ExpertLoadTracker(L, E, window=16)
    hot = torch.tensor([0, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 0])            # expert 0 is hot
    for step in range(20):
        for l in range(L):
            tr.record(l, hot[torch.randint(0, len(hot), (64, 2))])
        tr.end_step()
    load = tr.load()
    assert load.shape == (L, E) and int(load[0].argmax()) == 0
    pl = rebalance_experts(load, P, ranks)
    assert pl.phy2log.shape == (L, P)
    for l in range(L):
        assert set(pl.phy2log[l].tolist()) == set(range(E))              # every expert is placed
        assert int(pl.logcnt[l, 0]) >= 2                                 # the hot expert got replicas
        per_rank = [float(sum(load[l, e] / pl.logcnt[l, e] for e in pl.phy2log[l, r * 3:(r + 1) * 3])) for r in range(ranks)]
        assert max(per_rank) <= 1.6 * (sum(per_rank) / ranks)            # balanced within 60 % of the mean
    ids = hot[torch.randint(0, len(hot), (32, 2))]
    phys = route_to_physical(ids, pl, 0)
    assert torch.equal(pl.phy2log[0][phys], ids)                          # physical slots hold the requested experts
    moves = plan_transfers(torch.arange(P) % E, pl.phy2log[0], ranks)
    assert all(int((torch.arange(P) % E)[s]) == int(pl.phy2log[0][d]) for s, d, _, _ in moves)


def test_weight_transfer_buckets_roundtrip():
    from vllm_b200.parallel.weight_transfer import pack_buckets, unpack_bucket
    named = [(f"w{i}", torch.randn(3 + i, 5)) for i in range(7)] + [("b", torch.arange(6, dtype=torch.int32))]
    got = {}
    for metas, flat in pack_buckets(named, bucket_bytes=200):
        got.update(dict(unpack_bucket(metas, flat)))
    assert all(torch.equal(got[n], t) for n, t in named)


@pytest.mark.parametrize("method", ["awq", "gptq"])
def test_int4_checkpoints_dequantize_on_load(method):
    """A synthetic AWQ / GPTQ checkpoint produces exactly the model obtained from its dequantised weights."""
    from vllm_b200.model_loader import weight_only as wo
    torch.manual_seed(0)
    w = torch.randn(64, 128) * 0.1
    if method == "awq":
        qw, qz, sc = wo.pack_awq(w, 32)
        deq = wo.dequant_awq(qw, qz, sc, 32)
    else:
        qw, qz, sc, gi = wo.pack_gptq(w, 32)
        deq = wo.dequant_gptq(qw, qz, sc, gi, 32)
    assert deq.shape == w.shape and (deq - w).abs().max() < 0.06          # 4-bit rounding error only
    cfg = get_preset_hf_config("llama-tiny")
    full = dict(llama_full_weights(cfg, seed=8, scale=0.08))
    ckpt, ref = [], {}
    for name, t in full.items():
        if name.endswith("proj.weight"):
            mod = name[: -len(".weight")]
            if method == "awq":
                a, b, c = wo.pack_awq(t, 32)
                ckpt += [(mod + ".qweight", a), (mod + ".qzeros", b), (mod + ".scales", c)]
                ref[name] = wo.dequant_awq(a, b, c, 32)
            else:
                a, b, c, d = wo.pack_gptq(t, 32)
                ckpt += [(mod + ".qweight", a), (mod + ".qzeros", b), (mod + ".scales", c), (mod + ".g_idx", d)]
                ref[name] = wo.dequant_gptq(a, b, c, d, 32)
        else:
            ckpt.append((name, t))
            ref[name] = t
    sp = SamplingParams(max_tokens=6, temperature=0.0, ignore_eos=True)
    prompt = [{"prompt_token_ids": [5, 17, 99, 3, 250, 41, 77]}]
    llm = make_llm("llama-tiny", weights=ref.items())
    exp = llm.generate(prompt, sp)[0].outputs[0].token_ids
    llm.shutdown()
    llm = make_llm("llama-tiny", hf_overrides={"quantization_config": {"quant_method": method, "bits": 4, "group_size": 32}},
                   quantization=method)
    llm.llm_engine.engine_core.engine_core.model_executor.collective_rpc("load_model", kwargs={"weights": ckpt})
    got = llm.generate(prompt, sp)[0].outputs[0].token_ids
    llm.shutdown()
    assert got == exp


def test_fault_tolerance_parks_and_retries(monkeypatch):
    """A failing forward pass: without fault tolerance the error propagates; with it the in-flight request is
    aborted, the engine parks, and after an operator ``retry`` it serves again (reference: tests/v1/fault_tolerance)."""
    import threading
    from vllm_b200.engine.core import REQ_UTILITY
    prompt = [{"prompt_token_ids": [5, 6, 7, 8, 9]}]
    sp = SamplingParams(max_tokens=4, temperature=0.0, ignore_eos=True)
    llm = make_llm("llama-tiny")
    core = llm.llm_engine.engine_core.engine_core
    runner = core.model_executor.worker.runner
    good = runner._forward
    state = {"fail": True}

    def flaky(*a, **kw):
        if state["fail"]:
            state["fail"] = False
            raise RuntimeError("injected device fault")
        return good(*a, **kw)
    monkeypatch.setattr(runner, "_forward", flaky)
    with pytest.raises(RuntimeError, match="injected"):
        llm.generate(prompt, sp)                                   # in-process client: the exception surfaces
    # the LLM object aborted the poisoned request itself; workers forget their half-run step
    assert not core.scheduler.has_unfinished_requests()
    core.model_executor.collective_rpc("reset_after_fault")
    out = llm.generate(prompt, sp)
    assert len(out[0].outputs[0].token_ids) == 4
    llm.shutdown()


def test_batch_position_invariance_on_reference_path():
    """Same prompt alone, or in the middle of a batch of other prompts → identical greedy tokens (fp32 reference
    ops; reference: tests/v1/determinism)."""
    cfg = get_preset_hf_config("llama-tiny")
    w = list(llama_full_weights(cfg, seed=9))
    sp = SamplingParams(max_tokens=10, temperature=0.0, ignore_eos=True)
    target = {"prompt_token_ids": [11, 45, 200, 7, 9, 310, 12]}
    others = [{"prompt_token_ids": [(i * 37 + j) % 400 + 3 for j in range(5 + 6 * i)]} for i in range(5)]
    llm = make_llm("llama-tiny", weights=w)
    alone = llm.generate([target], sp)[0].outputs[0].token_ids
    mixed = llm.generate(others[:2] + [target] + others[2:], sp)[2].outputs[0].token_ids
    llm.shutdown()
    assert alone == mixed


def test_structured_outputs_choice_regex_json():
    """Built-in grammar back-end on the byte tokenizer: outputs satisfy the constraint and the request finishes as soon
    as the constraint admits no continuation."""
    import json
    import re
    from vllm_b200.sampling_params import StructuredOutputsParams
    llm = make_llm("llama-tiny")

    def gen(so, seed):
        sp = SamplingParams(max_tokens=60, temperature=0.9, seed=seed, structured_outputs=so)
        return llm.generate([{"prompt": "hello"}], sp)[0].outputs[0]
    for seed in range(3):
        o = gen(StructuredOutputsParams(choice=["yes", "no", "maybe"]), seed)
        assert o.text in ("yes", "no", "maybe") and o.finish_reason == "stop"
        o = gen(StructuredOutputsParams(regex=r"[0-9]{3}-[a-c]{2}"), seed)
        assert re.fullmatch(r"[0-9]{3}-[a-c]{2}", o.text) and o.finish_reason == "stop"
    schema = {"type": "object", "properties": {"name": {"type": "string", "maxLength": 4}, "age": {"type": "integer", "minimum": 0},
                                               "tags": {"type": "array", "items": {"enum": ["a", "b"]}, "maxItems": 2}},
              "required": ["name", "age", "tags"]}
    o = gen(StructuredOutputsParams(json=schema), 1)
    if o.finish_reason == "stop":                       # (a sampled integer may run to max_tokens: digits can continue)
        obj = json.loads(o.text)
        assert set(obj) == {"name", "age", "tags"} and isinstance(obj["age"], int) and len(obj["tags"]) <= 2
    else:
        assert o.text.startswith("{")
    llm.shutdown()


def test_fp8_kv_cache_close_to_full_precision():
    cfg = get_preset_hf_config("llama-tiny")
    w = list(llama_full_weights(cfg, seed=3))
    sp = SamplingParams(max_tokens=2, temperature=0.0, ignore_eos=True, prompt_logprobs=0)
    p = [{"prompt_token_ids": [(5 * i) % 400 + 3 for i in range(40)]}]
    res = {}
    for kvd in ("auto", "fp8"):
        llm = make_llm("llama-tiny", weights=w, kv_cache_dtype=kvd, max_num_batched_tokens=16)   # chunked: reads the cache
        o = llm.generate(p, sp)[0]
        res[kvd] = [list(d.values())[0].logprob for d in o.prompt_logprobs[1:]]
        if kvd == "fp8":
            r = llm.llm_engine.engine_core.engine_core.model_executor.worker.runner
            assert next(iter(r.kv_caches.values())).dtype == torch.float8_e4m3fn
        llm.shutdown()
    assert 0 < max(abs(a - b) for a, b in zip(res["auto"], res["fp8"])) < 0.1


def test_beam_search_and_parallel_sampling():
    """beam search beats (or ties) greedy in cumulative log-probability; n>1 returns n distinct seeded samples."""
    from vllm_b200.sampling_params import BeamSearchParams
    cfg = get_preset_hf_config("llama-tiny")
    llm = make_llm("llama-tiny", weights=list(llama_full_weights(cfg, seed=10, scale=0.1)))
    prompt = {"prompt_token_ids": [5, 17, 99, 3, 250, 41]}
    g = llm.generate([prompt], SamplingParams(max_tokens=5, temperature=0.0, ignore_eos=True, logprobs=0))[0].outputs[0]
    bs = llm.beam_search([prompt], BeamSearchParams(beam_width=3, max_tokens=5, ignore_eos=True))[0]
    best = bs.sequences[0]
    assert len(best.tokens) == 6 + 5 and best.cum_logprob >= g.cumulative_logprob - 1e-4
    outs = llm.generate([prompt], SamplingParams(max_tokens=6, temperature=1.0, n=4, seed=5, ignore_eos=True))[0].outputs
    assert len(outs) == 4 and len({tuple(o.token_ids) for o in outs}) >= 3
    again = llm.generate([prompt], SamplingParams(max_tokens=6, temperature=1.0, n=4, seed=5, ignore_eos=True))[0].outputs
    assert [o.token_ids for o in outs] == [o.token_ids for o in again]
    llm.shutdown()


def test_sleep_wake_and_prefix_cache_reset():
    llm = make_llm("llama-tiny")
    sp = SamplingParams(max_tokens=4, temperature=0.0, ignore_eos=True)
    p = [{"prompt_token_ids": list(range(3, 70))}]
    a = llm.generate(p, sp)[0].outputs[0].token_ids
    llm.sleep(level=1)
    assert llm.llm_engine.engine_core.engine_core.model_executor.worker.is_sleeping()
    llm.wake_up()
    b = llm.generate(p, sp)[0].outputs[0].token_ids
    assert a == b
    assert llm.reset_prefix_cache()
    llm.shutdown()


def test_prompt_embeds_equal_token_prompt():
    """A prompt passed as embedding rows gives the same continuation as the token ids those rows came from."""
    cfg = get_preset_hf_config("llama-tiny")
    llm = make_llm("llama-tiny", weights=list(llama_full_weights(cfg, seed=12)), enable_prefix_caching=False)
    ids = [5, 17, 99, 3, 250, 41, 8]
    sp = SamplingParams(max_tokens=6, temperature=0.0, ignore_eos=True)
    a = llm.generate([{"prompt_token_ids": ids}], sp)[0].outputs[0].token_ids
    model = llm.llm_engine.engine_core.engine_core.model_executor.worker.runner.model
    with torch.no_grad():
        emb = model.embed_input_ids(torch.tensor(ids))
    b = llm.generate([{"prompt_embeds": emb}], sp)[0].outputs[0].token_ids
    llm.shutdown()
    assert a == b


def test_extra_ops_and_fp4_roundtrip():
    from vllm_b200 import ops
    from vllm_b200.model_loader.weight_only import dequantize_stream
    from vllm_b200.ops import ref
    torch.manual_seed(0)
    x = torch.randn(5, 16)
    assert torch.allclose(ops.mul_and_silu(x), x[:, :8] * torch.nn.functional.silu(x[:, 8:]), atol=1e-6)
    assert (ops.fatrelu_and_mul(x, 0.5)[x[:, :8] <= 0.5] == 0).all()
    assert ops.swigluoai_and_mul(x).shape == (5, 8)
    q, s = ops.scaled_fp8_quant(x, per_token=True)
    assert q.dtype == torch.float8_e4m3fn and (q.float() * s - x).abs().max() < 0.2
    y = ops.scaled_mm(q, q, s, s, torch.float32)
    assert torch.allclose(y, x @ x.t(), atol=0.8)
    src, dst = torch.arange(24.).view(6, 4), torch.zeros(6, 4)
    ops.swap_blocks(src, dst, [(0, 5), (2, 1)])
    assert torch.equal(dst[5], src[0]) and torch.equal(dst[1], src[2]) and dst[0].sum() == 0
    w = torch.randn(8, 64) * 0.2
    packed, bs, gs = ref.fp4_quant(w)
    deq = ref.fp4_dequant(packed, bs, gs)
    assert deq.shape == w.shape and (deq - w).abs().max() < 0.15 and (deq - w).abs().mean() < 0.03
    stream = [("model.layers.0.mlp.down_proj.weight", packed), ("model.layers.0.mlp.down_proj.weight_scale", bs),
              ("model.layers.0.mlp.down_proj.weight_scale_2", torch.tensor(gs)),
              ("model.layers.0.mlp.down_proj.input_scale", torch.tensor(1.0)), ("model.norm.weight", torch.ones(4))]
    out = dict(dequantize_stream(stream, {"quant_method": "modelopt_fp4"}, torch.float32))
    assert set(out) == {"model.layers.0.mlp.down_proj.weight", "model.norm.weight"}
    assert torch.allclose(out["model.layers.0.mlp.down_proj.weight"], deq)


def test_deepseek_fp8_block_experts_and_linears():
    """DeepSeek-style block-scaled fp8 (online quantisation here): dense projections AND routed experts are held in
    e4m3 with 128x128 block scales; outputs stay within quantisation noise of the unquantised model."""
    from tests.utils import make_hf_model
    hf = make_hf_model("deepseek-tiny")
    w = [(k, v.clone()) for k, v in hf.state_dict().items()]
    sp = SamplingParams(max_tokens=3, temperature=0.0, ignore_eos=True, prompt_logprobs=0)
    prompt = [{"prompt_token_ids": [(11 * i) % 500 + 3 for i in range(30)]}]

    def run(**kw):
        llm = make_llm("deepseek-tiny", **kw)
        ex = llm.llm_engine.engine_core.engine_core.model_executor
        ex.collective_rpc("load_model", kwargs={"weights": list(w)})
        model = ex.worker.runner.model
        dts = {m.w13_weight.dtype for m in model.modules() if hasattr(m, "w13_weight")}
        o = llm.generate(prompt, sp)[0]
        llm.shutdown()
        return [list(d.values())[0].logprob for d in o.prompt_logprobs[1:]], dts
    base, d0 = run()
    q, d1 = run(quantization="fp8_block")
    assert d0 == {torch.float32} and d1 == {torch.float8_e4m3fn}
    diff = max(abs(a - b) for a, b in zip(base, q))
    assert 0 < diff < 0.3, diff


def test_mxfp4_expert_dequant():
    """e2m1 nibbles (low first) x e8m0 block exponents, the packing GPT-OSS checkpoints use."""
    from vllm_b200.models.gpt_oss import dequant_mxfp4
    blocks = torch.tensor([[[0x21, 0xF7] + [0] * 14]], dtype=torch.uint8)        # values: 0.5, 1.0, 6.0, -6.0, 0...
    scales = torch.tensor([[128]], dtype=torch.uint8)                             # 2^(128-127) = 2
    out = dequant_mxfp4(blocks, scales, torch.float32)
    assert out.shape == (1, 32)
    assert out[0, :4].tolist() == [1.0, 2.0, 12.0, -12.0] and out[0, 4:].abs().sum() == 0


def test_ssd_chunk_scan_matches_recurrence():
    """Chunked state-space-duality scan (intra-chunk matrix form + carried state) == the token-by-token recurrence,
    including a non-zero initial state, a ragged last chunk and grouped B/C."""
    from vllm_b200.ops import ref
    torch.manual_seed(0)
    T, H, P, G, N = 37, 4, 8, 2, 6
    x, dt = torch.randn(T, H, P), torch.rand(T, H) * 0.5 + 0.01
    A, B, C, h0 = -torch.rand(H) - 0.1, torch.randn(T, G, N), torch.randn(T, G, N), torch.randn(H, P, N)
    y, hT = ref.ssd_chunk_scan(x, dt, A, B, C, 16, h0)
    h, ys = h0.clone(), []
    for t in range(T):
        Bh, Ch = B[t].repeat_interleave(H // G, 0), C[t].repeat_interleave(H // G, 0)
        h = h * torch.exp(dt[t] * A)[:, None, None] + (dt[t][:, None] * x[t])[..., None] * Bh[:, None, :]
        ys.append((h * Ch[:, None, :]).sum(-1))
    assert (y - torch.stack(ys)).abs().max() < 1e-4 and (hT - h).abs().max() < 1e-4


def test_recurrent_state_blocks_are_per_request():
    """Every running request of a state-space model owns its own state block (never the shared null block)."""
    llm = make_llm("mamba2-tiny")
    core = llm.llm_engine.engine_core.engine_core
    for i, p in enumerate(PROMPTS):
        llm.llm_engine.add_request(str(i), p, SamplingParams(max_tokens=4, temperature=0.0, ignore_eos=True))
    core.step() if hasattr(core, "step") else None
    mgr = core.scheduler.kv_cache_manager.managers[0]
    ids = [b.block_id for blocks in mgr.req_to_blocks.values() for b in blocks]
    assert len(ids) == len(PROMPTS) and len(set(ids)) == len(ids) and 0 not in ids
    llm.shutdown()


def test_compressed_tensors_checkpoints():
    """compressed-tensors: (a) pack-quantized W4A16 expands to exactly its dequantised weights; (b) a W8A8-int8
    checkpoint (int8 weights + per-channel scales, dynamic per-token activations) runs through the int8 linear method
    and stays within quantisation noise of the full-precision model."""
    from vllm_b200.model_loader import weight_only as wo
    cfg = get_preset_hf_config("llama-tiny")
    full = dict(llama_full_weights(cfg, seed=9, scale=0.08))
    sp = SamplingParams(max_tokens=6, temperature=0.0, ignore_eos=True)
    prompt = [{"prompt_token_ids": [5, 17, 99, 3, 250, 41, 77]}]
    # (a) W4A16
    ckpt, ref = [], {}
    for name, t in full.items():
        if name.endswith("proj.weight"):
            mod = name[: -len(".weight")]
            pk, sc, shp = wo.pack_compressed_tensors(t, 4, 32)
            ckpt += [(mod + ".weight_packed", pk), (mod + ".weight_scale", sc), (mod + ".weight_shape", shp)]
            ref[name] = wo.dequant_compressed_tensors(pk, sc, None, shp.tolist(), 4, 32, True)
            assert (ref[name] - t).abs().max() < 0.06
        else:
            ckpt.append((name, t)); ref[name] = t
    llm = make_llm("llama-tiny", weights=ref.items())
    exp = llm.generate(prompt, sp)[0].outputs[0].token_ids
    llm.shutdown()
    qc = {"quant_method": "compressed-tensors", "format": "pack-quantized", "ignore": ["lm_head"],
          "config_groups": {"group_0": {"targets": ["Linear"], "weights": {"num_bits": 4, "type": "int", "strategy": "group",
                                                                         "group_size": 32, "symmetric": True},
                                        "input_activations": None}}}
    llm = make_llm("llama-tiny", hf_overrides={"quantization_config": qc})
    llm.llm_engine.engine_core.engine_core.model_executor.collective_rpc("load_model", kwargs={"weights": ckpt})
    got = llm.generate(prompt, sp)[0].outputs[0].token_ids
    llm.shutdown()
    assert got == exp
    # (b) W8A8 int8, channel-wise weights, dynamic token activations
    ckpt = []
    for name, t in full.items():
        if name.endswith("proj.weight"):
            scale = (t.float().abs().amax(dim=1, keepdim=True) / 127.0).clamp_min(1e-12)
            ckpt += [(name, (t.float() / scale).round().clamp(-128, 127).to(torch.int8)),
                     (name[: -len(".weight")] + ".weight_scale", scale)]
        else:
            ckpt.append((name, t))
    qc = {"quant_method": "compressed-tensors", "format": "int-quantized", "ignore": ["lm_head"],
          "config_groups": {"group_0": {"targets": ["Linear"],
                                        "weights": {"num_bits": 8, "type": "int", "strategy": "channel", "symmetric": True},
                                        "input_activations": {"num_bits": 8, "type": "int", "strategy": "token",
                                                              "dynamic": True, "symmetric": True}}}}
    spl = SamplingParams(max_tokens=2, temperature=0.0, ignore_eos=True, prompt_logprobs=0)
    lp = {}
    for tag, kw, wts in (("fp", {}, list(full.items())), ("int8", {"hf_overrides": {"quantization_config": qc}}, ckpt)):
        llm = make_llm("llama-tiny", **kw)
        llm.llm_engine.engine_core.engine_core.model_executor.collective_rpc("load_model", kwargs={"weights": wts})
        if tag == "int8":
            m = llm.llm_engine.engine_core.engine_core.model_executor.worker.runner.model
            assert m.model.layers[0].self_attn.qkv_proj.weight.dtype == torch.int8
        o = llm.generate([{"prompt_token_ids": [(11 * i) % 500 + 3 for i in range(40)]}], spl)[0]
        lp[tag] = [list(d.values())[0].logprob for d in o.prompt_logprobs[1:]]
        llm.shutdown()
    diff = max(abs(a - b) for a, b in zip(lp["fp"], lp["int8"]))
    assert 0 < diff < 0.25, diff


def test_encoder_cache_connector_roundtrip_and_plugins(tmp_path, monkeypatch):
    """Encoder-output disaggregation store (producer writes embeddings by content hash, consumer reads them) and the
    entry-point plugin loader (allow-list through VLLM_PLUGINS, failures isolated)."""
    from vllm_b200.ec_transfer import ECSharedStorageConnector
    prod = ECSharedStorageConnector(str(tmp_path), "ec_producer")
    cons = ECSharedStorageConnector(str(tmp_path), "ec_consumer")
    emb = torch.randn(7, 32).to(torch.bfloat16)
    prod.save("img-abc", emb)
    got = cons.load("img-abc")
    assert got.dtype == torch.bfloat16 and torch.equal(got, emb)
    assert cons.load("missing") is None and prod.load("img-abc") is None      # producers do not read

    import vllm_b200.plugins as pl
    calls = []

    class EP:
        def __init__(self, name, fn):
            self.name, self._fn = name, fn

        def load(self):
            return self._fn
    def boom():
        raise RuntimeError("bad plugin")
    monkeypatch.setattr(pl, "_entry_points", lambda group: [EP("a", lambda: calls.append("a")), EP("b", boom),
                                                            EP("c", lambda: calls.append("c"))])
    monkeypatch.setattr(pl, "_loaded", False)
    monkeypatch.setenv("VLLM_PLUGINS", "a,b")
    pl.load_general_plugins()
    assert calls == ["a"]                      # "c" filtered by the allow-list, "b" raised without taking the process down


def test_gguf_checkpoint_roundtrip(tmp_path):
    """A llama.cpp-style file (Q8_0 / Q4_0 / F16 / F32 tensors, permuted q/k rows, metadata-only config) loads and
    generates exactly like the model built from the de-quantised tensors; K-quant block decoders match scalar code."""
    from vllm_b200.model_loader import gguf as G
    cfg = get_preset_hf_config("llama-tiny")
    full = dict(llama_full_weights(cfg, seed=12, scale=0.08))
    nh, nkv = cfg["num_attention_heads"], cfg["num_key_value_heads"]

    def permute(w, n):                       # what convert_hf_to_gguf does to q/k
        return w.view(n, 2, w.shape[0] // n // 2, *w.shape[1:]).transpose(1, 2).reshape(w.shape)
    inv = {v: k for k, v in G._LAYER_MAP.items()}
    tensors, ref = [], {}
    for name, w in full.items():
        if name == "model.embed_tokens.weight":
            gname, t = "token_embd.weight", G.F16
        elif name == "model.norm.weight":
            gname, t = "output_norm.weight", G.F32
        elif name == "lm_head.weight":
            gname, t = "output.weight", G.Q8_0
        else:
            _, _, i, rest = name.split(".", 3)
            stem, leaf = rest.rsplit(".", 1)
            gname = f"blk.{i}.{inv[stem]}.{leaf}"
            t = G.F32 if "norm" in stem else (G.Q4_0 if "mlp" in stem else G.Q8_0)
        src = permute(w, nh) if name.endswith("q_proj.weight") else permute(w, nkv) if name.endswith("k_proj.weight") else w
        tensors.append((gname, src, t))
        per, nb = G.BLOCK[t]
        deq = G.dequantize(G.quantize(src, t), t, src.numel()).reshape(src.shape)
        ref[name] = G._unpermute(deq, nh) if name.endswith("q_proj.weight") else G._unpermute(deq, nkv) if name.endswith("k_proj.weight") else deq
    path = str(tmp_path / "tiny.gguf")
    G.write_gguf(path, "llama", {"llama.block_count": cfg["num_hidden_layers"], "llama.embedding_length": cfg["hidden_size"],
                                 "llama.feed_forward_length": cfg["intermediate_size"], "llama.attention.head_count": nh,
                                 "llama.attention.head_count_kv": nkv, "llama.attention.layer_norm_rms_epsilon": float(cfg["rms_norm_eps"]),
                                 "llama.rope.freq_base": float(cfg.get("rope_theta", 10000.0)), "llama.context_length": 256,
                                 "llama.attention.key_length": cfg.get("head_dim") or cfg["hidden_size"] // nh,
                                 "tokenizer.ggml.eos_token_id": 2}, tensors)
    sp = SamplingParams(max_tokens=6, temperature=0.0, ignore_eos=True)
    prompt = [{"prompt_token_ids": [5, 17, 99, 3, 250, 41, 77]}]
    llm = make_llm("llama-tiny", weights=ref.items())
    exp = llm.generate(prompt, sp)[0].outputs[0].token_ids
    llm.shutdown()
    from vllm_b200 import LLM
    llm = LLM(path, device="cpu", dtype="float32", max_model_len=128, max_num_seqs=4, max_num_batched_tokens=64,
              kv_cache_memory_bytes=8 << 20, skip_tokenizer_init=True)
    assert llm.llm_engine.model_config.hf_config["num_hidden_layers"] == cfg["num_hidden_layers"]
    got = llm.generate(prompt, sp)[0].outputs[0].token_ids
    llm.shutdown()
    assert got == exp

    # K-quants: vectorised decoders vs a direct transcription of ggml's scalar loops on random blocks
    import numpy as np
    rng = np.random.default_rng(0)
    raw = rng.integers(0, 256, size=(3, 144), dtype=np.uint8)
    raw[:, 0:4] = np.array([0.01, 0.002], dtype=np.float16).view(np.uint8)
    got4 = G.dequantize(raw.reshape(-1), G.Q4_K, 3 * 256).reshape(3, 256).numpy()
    for b in range(3):
        d, dmin = raw[b, 0:2].copy().view(np.float16)[0].astype(np.float32), raw[b, 2:4].copy().view(np.float16)[0].astype(np.float32)
        sc, q, y, is_ = raw[b, 4:16], raw[b, 16:], [], 0
        def sm(j):
            if j < 4:
                return sc[j] & 63, sc[j + 4] & 63
            return (sc[j + 4] & 0xF) | ((sc[j - 4] >> 6) << 4), (sc[j + 4] >> 4) | ((sc[j] >> 6) << 4)
        for j in range(0, 256, 64):
            s1, m1 = sm(is_); s2, m2 = sm(is_ + 1)
            qq = q[j // 2:j // 2 + 32]
            y += [d * s1 * (v & 0xF) - dmin * m1 for v in qq] + [d * s2 * (v >> 4) - dmin * m2 for v in qq]
            is_ += 2
        assert np.allclose(got4[b], np.array(y, dtype=np.float32), atol=1e-5)
    raw = rng.integers(0, 256, size=(2, 210), dtype=np.uint8)
    raw[:, 208:210] = np.array([0.01], dtype=np.float16).view(np.uint8)
    got6 = G.dequantize(raw.reshape(-1), G.Q6_K, 2 * 256).reshape(2, 256).numpy()
    for b in range(2):
        d = raw[b, 208:210].copy().view(np.float16)[0].astype(np.float32)
        ql, qh, sc = raw[b, :128], raw[b, 128:192], raw[b, 192:208].view(np.int8)
        y = np.zeros(256, np.float32)
        for n in range(2):
            for l in range(32):
                i = l // 16
                o, lq, hq, s = n * 128, ql[n * 64:], qh[n * 32:], sc[n * 8:]
                q1 = int((lq[l] & 0xF) | (((hq[l] >> 0) & 3) << 4)) - 32
                q2 = int((lq[l + 32] & 0xF) | (((hq[l] >> 2) & 3) << 4)) - 32
                q3 = int((lq[l] >> 4) | (((hq[l] >> 4) & 3) << 4)) - 32
                q4 = int((lq[l + 32] >> 4) | (((hq[l] >> 6) & 3) << 4)) - 32
                y[o + l], y[o + l + 32], y[o + l + 64], y[o + l + 96] = d * s[i] * q1, d * s[i + 2] * q2, d * s[i + 4] * q3, d * s[i + 6] * q4
        assert np.allclose(got6[b], y, atol=1e-5)


@pytest.mark.parametrize("backend", ["xgrammar", "guidance", "outlines", "lm-format-enforcer"])
def test_structured_output_library_backends(backend):
    """xgrammar / llguidance matchers behind the same bitmask interface, driven with a real (programmatically built)
    byte-level BPE tokenizer: every token of a valid document is allowed in turn, an invalid one is masked, the
    matcher terminates, rollback restores earlier states."""
    import json
    import numpy as np
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast
    from vllm_b200.engine.arg_utils import EngineArgs
    from vllm_b200.sampling_params import StructuredOutputsParams
    from vllm_b200.structured_output import StructuredOutputManager
    vocab = {ch: i for i, ch in enumerate(sorted(pre_tokenizers.ByteLevel.alphabet()))}
    merges = []
    for a, b in [("t", "r"), ("tr", "u"), ("tru", "e"), ("f", "a"), ("fa", "l"), ("fal", "s"), ("fals", "e")]:
        vocab[a + b] = len(vocab)
        merges.append((a, b))
    vocab["<eos>"] = len(vocab)
    tk = Tokenizer(models.BPE(vocab=vocab, merges=merges))
    tk.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    tk.decoder = decoders.ByteLevel()
    hf = PreTrainedTokenizerFast(tokenizer_object=tk, eos_token="<eos>")
    cfg = EngineArgs(model="llama-tiny", device="cpu", structured_outputs_backend=backend).create_engine_config()
    cfg.model_config.hf_config["eos_token_id"] = hf.eos_token_id
    mgr = StructuredOutputManager(cfg)
    mgr._tokenizer = hf
    schema = {"type": "object", "properties": {"ok": {"type": "boolean"}, "n": {"type": "integer"}}, "required": ["ok", "n"],
              "additionalProperties": False}
    req = mgr.make_request("r", StructuredOutputsParams(json=schema))
    req._fut.result()
    assert req.is_ready()
    g = req.grammar
    doc = '{"ok": true, "n": 42}'
    ids = hf.encode(doc, add_special_tokens=False)
    W = (mgr.vocab_size + 31) // 32
    bm = np.zeros((1, W), dtype=np.int32)

    def allowed(t):
        g.fill_bitmask(bm, 0)
        return bool((int(bm[0, t >> 5]) >> (t & 31)) & 1)
    bad = hf.encode("]", add_special_tokens=False)[0]
    assert not allowed(bad)                                   # a document cannot start with ']'
    for i, t in enumerate(ids):
        assert allowed(t), (i, hf.decode([t]))
        assert g.accept_tokens("r", [t])
    assert g.validate_tokens([bad]) == []
    g.rollback(3)                                             # drop "42}" (three byte tokens) and replay
    for t in ids[-3:]:
        assert allowed(t) and g.accept_tokens("r", [t])
    assert allowed(hf.eos_token_id) or g.is_terminated()
    g.accept_tokens("r", [hf.eos_token_id])
    assert g.is_terminated()
    assert json.loads(doc)["n"] == 42
    mgr.shutdown()


def test_custom_logits_processor_plugin():
    """--logits-processors module:Class — requests opt in through extra_args, others are untouched (same batch)."""
    llm = make_llm("llama-tiny", logits_processors=["tests.plugin_logits_processor:ForceToken"])
    sp_plain = SamplingParams(max_tokens=5, temperature=0.0, ignore_eos=True)
    sp_force = SamplingParams(max_tokens=5, temperature=0.0, ignore_eos=True, extra_args={"force_token": 77})
    base = llm.generate([PROMPTS[0]], sp_plain)[0].outputs[0].token_ids
    outs = llm.generate([PROMPTS[0], PROMPTS[0]], [sp_force, sp_plain])
    llm.shutdown()
    assert outs[0].outputs[0].token_ids == [77] * 5
    assert outs[1].outputs[0].token_ids == base


def test_bitsandbytes_nf4_checkpoint():
    """A pre-quantised bitsandbytes NF4 checkpoint (packed codes + absmax + quant map + JSON quant state, with and
    without double quantisation of the absmax) expands to exactly its de-quantised weights."""
    import json
    from vllm_b200.model_loader import weight_only as wo
    torch.manual_seed(0)
    w = torch.randn(64, 128) * 0.1
    pk, am, qm = wo.pack_bnb4(w, 64)
    deq = wo.dequant_bnb4(pk, am, qm, w.shape, 64)
    assert (deq - w).abs().max() < 0.05 and deq.shape == w.shape
    # double quantisation: absmax → 8-bit codes around its mean
    off = am.mean()
    nb = 256
    a2 = (am - off).reshape(-1, min(nb, am.numel()))
    nam = a2.abs().amax(1).clamp_min(1e-12)
    nqm = torch.linspace(-1, 1, 256)
    codes = ((a2 / nam[:, None])[..., None] - nqm).abs().argmin(-1).reshape(-1).to(torch.uint8)
    deq2 = wo.dequant_bnb4(pk, codes, qm, w.shape, 64, nam, nqm, a2.shape[1], float(off))
    assert (deq2 - deq).abs().max() < 0.01
    cfg = get_preset_hf_config("llama-tiny")
    full = dict(llama_full_weights(cfg, seed=13, scale=0.08))
    ckpt, ref = [], {}
    for name, t in full.items():
        if name.endswith("proj.weight"):
            pk, am, qm = wo.pack_bnb4(t, 64)
            st = json.dumps({"quant_type": "nf4", "blocksize": 64, "shape": list(t.shape), "dtype": "float32"}).encode()
            ckpt += [(name, pk), (name + ".absmax", am), (name + ".quant_map", qm),
                     (name + ".quant_state.bitsandbytes__nf4", torch.tensor(list(st), dtype=torch.uint8))]
            ref[name] = wo.dequant_bnb4(pk, am, qm, t.shape, 64)
        else:
            ckpt.append((name, t)); ref[name] = t
    sp = SamplingParams(max_tokens=6, temperature=0.0, ignore_eos=True)
    prompt = [{"prompt_token_ids": [5, 17, 99, 3, 250, 41, 77]}]
    llm = make_llm("llama-tiny", weights=ref.items())
    exp = llm.generate(prompt, sp)[0].outputs[0].token_ids
    llm.shutdown()
    llm = make_llm("llama-tiny", hf_overrides={"quantization_config": {"quant_method": "bitsandbytes", "load_in_4bit": True}})
    llm.llm_engine.engine_core.engine_core.model_executor.collective_rpc("load_model", kwargs={"weights": ckpt})
    got = llm.generate(prompt, sp)[0].outputs[0].token_ids
    llm.shutdown()
    assert got == exp


def test_numa_helpers(tmp_path):
    from vllm_b200.utils import numa
    assert numa.parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    for n, cpus in ((0, "0-1"), (1, "2-3")):
        d = tmp_path / f"node{n}"
        d.mkdir()
        (d / "cpulist").write_text(cpus + "\n")
    nodes = numa.numa_nodes(str(tmp_path))
    assert nodes == {0: [0, 1], 1: [2, 3]}
    assert numa.bind_to_node(None, nodes) is None and numa.bind_to_node(7, nodes) is None
    import os
    have = sorted(os.sched_getaffinity(0))
    got = numa.bind_to_node(0, {0: have[:1], 1: have[1:] or have[:1]})
    try:
        assert got == have[:1] and sorted(os.sched_getaffinity(0)) == have[:1]
    finally:
        os.sched_setaffinity(0, have)


def test_request_tracing_spans_and_metrics_snapshot():
    """--otlp-traces-endpoint: one span per finished request with the gen_ai.* attributes, exported over OTLP/HTTP to a
    collector (a local stub here) with the caller's W3C trace context as parent; the Prometheus registry exposes the
    request / token counters the reference names."""
    import threading
    from http.server import BaseHTTPRequestHandler, HTTPServer
    from vllm_b200.metrics import get_metrics_snapshot, render_prometheus
    from vllm_b200.tracing import SpanAttributes, extract_trace_context
    got = []

    class H(BaseHTTPRequestHandler):
        def do_POST(self):
            got.append((self.path, self.rfile.read(int(self.headers.get("Content-Length", 0)))))
            self.send_response(200); self.send_header("Content-Type", "application/x-protobuf"); self.end_headers()

        def log_message(self, *a):
            pass
    srv = HTTPServer(("127.0.0.1", 0), H)
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    llm = make_llm("llama-tiny", otlp_traces_endpoint=f"http://127.0.0.1:{srv.server_address[1]}/v1/traces",
                   disable_log_stats=False)
    sp = SamplingParams(max_tokens=5, temperature=0.0, ignore_eos=True)
    llm.generate(PROMPTS, sp)
    tracer = llm.llm_engine.tracer
    spans = list(tracer.finished)
    flushed = tracer.flush()
    llm.shutdown()
    srv.shutdown()
    assert len(spans) == len(PROMPTS)
    s0 = spans[0]
    assert s0[SpanAttributes.COMPLETION_TOKENS] == 5 and s0[SpanAttributes.PROMPT_TOKENS] == len(PROMPTS[0]["prompt_token_ids"])
    assert s0[SpanAttributes.E2E] > 0 and s0[SpanAttributes.TTFT] > 0 and "length" in s0[SpanAttributes.FINISH_REASON].lower()
    if tracer._otel is not None:                       # OpenTelemetry SDK present: the spans really left the process
        assert flushed and got and got[0][0] == "/v1/traces" and b"llm_request" in b"".join(b for _, b in got)
        assert b"gen_ai.usage.completion_tokens" in b"".join(b for _, b in got)
    assert extract_trace_context({"traceparent": "00-4bf92f3577b34da6a3ce929d0e0e4736-00f067aa0ba902b7-01"}) is not None
    body, ctype = render_prometheus()
    text = body.decode()
    for name in ("vllm:generation_tokens_total", "vllm:prompt_tokens_total", "vllm:request_success_total",
                 "vllm:time_to_first_token_seconds", "vllm:num_requests_running"):
        assert name in text, name
    assert get_metrics_snapshot()


def test_sharded_state_checkpoint_hooks_profiler_and_crash_dump(tmp_path, monkeypatch):
    """save_sharded_state → a fresh engine loads the per-rank file and generates the same text; checkpoint
    prepare / restore hooks round-trip; the worker profiler honours delay / max-iterations; a failing forward pass
    logs a dump of the scheduled step (ids and sizes, no prompts)."""
    from vllm_b200.profiler import WorkerProfiler
    cfg = get_preset_hf_config("llama-tiny")
    w = list(llama_full_weights(cfg, seed=21))
    sp = SamplingParams(max_tokens=6, temperature=0.0, ignore_eos=True)
    llm = make_llm("llama-tiny", weights=w)
    exp = llm.generate(PROMPTS[:1], sp)[0].outputs[0].token_ids
    d = str(tmp_path / "shards")
    llm.llm_engine.engine_core.engine_core.save_sharded_state(d)
    ex = llm.llm_engine.engine_core.engine_core.model_executor
    ex.collective_rpc("checkpoint_prepare")
    ex.collective_rpc("checkpoint_restore")
    assert llm.generate(PROMPTS[:1], sp)[0].outputs[0].token_ids == exp
    llm.shutdown()
    import os
    assert any(f.endswith(".safetensors") for f in os.listdir(d))
    llm = make_llm("llama-tiny")
    from vllm_b200.model_loader.sharded_state import load_sharded_state
    load_sharded_state(llm.llm_engine.engine_core.engine_core.model_executor.worker.runner.model, d)
    assert llm.generate(PROMPTS[:1], sp)[0].outputs[0].token_ids == exp
    # crash dump
    core = llm.llm_engine.engine_core.engine_core
    def boom(*a, **k):
        raise RuntimeError("injected failure")
    monkeypatch.setattr(core.model_executor, "execute_model", boom)
    import logging
    records = []

    class Grab(logging.Handler):
        def emit(self, record):
            records.append(record.getMessage())
    h = Grab(level=logging.ERROR)
    lg = logging.getLogger("vllm_b200.engine.core")
    lg.addHandler(h)
    try:
        with pytest.raises(Exception):
            llm.generate(PROMPTS[:1], sp)
    finally:
        lg.removeHandler(h)
    llm.shutdown()
    err = "\n".join(records)
    assert "engine step failed" in err and "tokens scheduled" in err and "new requests (id, prompt tokens)" in err

    events = []

    class P(WorkerProfiler):
        def _start(self): events.append("start")
        def _stop(self): events.append("stop")
    p = P(delay_iterations=2, max_iterations=3)
    p.start()
    for _ in range(8):
        p.step()
    assert events == ["start", "stop"] and not p.running


def test_cli_config_file(tmp_path):
    from vllm_b200.entrypoints.cli import expand_config_file
    f = tmp_path / "c.yaml"
    f.write_text("\n".join(["model: llama-tiny", "max-model-len: 256", "enable_prefix_caching: false",
                            "cudagraph_capture_sizes: [1, 2]", "hf_overrides: {a: 1}", "port: 9000"]) + "\n")
    out = expand_config_file(["serve", "--config", str(f), "--port", "9100"])
    assert out[:2] == ["serve", "llama-tiny"] and out[-2:] == ["--port", "9100"]
    assert "--no-enable-prefix-caching" in out and out[out.index("--cudagraph-capture-sizes") + 1] == "1,2"
    from vllm_b200.entrypoints.openai.api_server import make_arg_parser
    args = make_arg_parser().parse_args(out[1:])
    assert args.model_tag == "llama-tiny" and args.port == 9100 and args.max_model_len == 256 and args.hf_overrides == {"a": 1}
    assert args.enable_prefix_caching is False


def test_decode_bench_connector_skips_prefill():
    """The decode-benchmark connector claims every full prompt block as remotely available: the engine computes only
    the tail of the prompt and decodes (over constant-filled KV)."""
    llm = make_llm("llama-tiny", block_size=16, enable_prefix_caching=True,
                   kv_transfer_config=dict(kv_connector="DecodeBenchConnector", kv_role="kv_both"))
    core = llm.llm_engine.engine_core.engine_core
    seen = []
    orig = core.model_executor.execute_model

    def spy(so, *a, **k):
        seen.append(dict(so.num_scheduled_tokens))
        return orig(so, *a, **k)
    core.model_executor.execute_model = spy
    prompt = [{"prompt_token_ids": [(3 * i) % 400 + 5 for i in range(100)]}]
    out = llm.generate(prompt, SamplingParams(max_tokens=4, temperature=0.0, ignore_eos=True))[0]
    conn = core.model_executor.worker.runner.kv_connector
    llm.shutdown()
    assert len(out.outputs[0].token_ids) == 4
    assert conn.num_filled_blocks == 6 and max(seen[0].values()) == 100 - 96       # 6 blocks "loaded", 4 tokens computed


def test_multimodal_sender_receiver_cache():
    """Repeated multimodal items cross the front-end → engine-core channel once: the sender strips payloads the
    receiver holds, the receiver restores them, both evict in lock-step; end to end the repeat gives the same output."""
    from vllm_b200.multimodal.cache import ReceiverCache, SenderShadow
    snd, rcv = SenderShadow(2), ReceiverCache(2)
    def item(h):
        return [{"modality": "image", "hash": h, "data": f"payload-{h}", "offset": 0, "length": 1},
                {"modality": "mrope", "hash": "", "data": "pos", "offset": 0, "length": 0}]
    sent = []
    for h in ["a", "b", "a", "c", "b", "a"]:            # capacity 2: "b" is evicted by "c", "a" by the second "b"
        f = item(h)
        snd.strip(f)
        sent.append(f[0]["data"] is not None)
        rcv.restore(f)
        assert f[0]["data"] == f"payload-{h}" and f[1]["data"] == "pos"
    assert sent == [True, True, False, True, True, True] and snd.hits == 1
    with pytest.raises(RuntimeError):
        rcv.restore([{"modality": "image", "hash": "zzz", "data": None}])

    import numpy as np
    rng = np.random.default_rng(0)
    wave = (0.2 * np.sin(2 * np.pi * 300 * np.arange(16000) / 16000) + 0.01 * rng.standard_normal(16000)).astype(np.float32)
    llm = make_llm("whisper-tiny-test", enable_prefix_caching=False, max_model_len=64)
    sp = SamplingParams(max_tokens=4, temperature=0.0, ignore_eos=True)
    p = {"multi_modal_data": {"audio": (wave, 16000)}}
    first = llm.generate([p], sp)[0].outputs[0].token_ids
    second = llm.generate([p], sp)[0].outputs[0].token_ids
    sender = llm.llm_engine.input_processor.mm_sender if hasattr(llm.llm_engine, "input_processor") else None
    llm.shutdown()
    assert first == second
    if sender is not None:
        assert sender.hits == 1 and sender.misses == 1


def test_cascade_attention_shared_prefix(monkeypatch):
    """Requests that share their leading KV blocks (prefix cache): the shared part is attended once for the batch and
    merged with the per-request remainder by log-sum-exp — same tokens as the ordinary path."""
    from vllm_b200.layers.attention import Attention
    cfg = get_preset_hf_config("llama-tiny")
    w = list(llama_full_weights(cfg, seed=31))
    shared = [(13 * i) % 400 + 5 for i in range(64)]
    prompts = [{"prompt_token_ids": shared + [7 + j, 9 + 2 * j, 11 + j]} for j in range(4)]
    sp = SamplingParams(max_tokens=6, temperature=0.0, ignore_eos=True)

    def run():
        llm = make_llm("llama-tiny", weights=w, block_size=16, enable_prefix_caching=True)
        llm.generate([{"prompt_token_ids": shared + [3]}], SamplingParams(max_tokens=1, temperature=0.0))   # fills the cache
        before = Attention.cascade_calls
        out = [o.outputs[0].token_ids for o in llm.generate(prompts, sp)]
        used = Attention.cascade_calls - before
        llm.shutdown()
        return out, used
    monkeypatch.setenv("VLLM_B200_CASCADE_MIN_REQS", "1000")
    ref, used0 = run()
    monkeypatch.setenv("VLLM_B200_CASCADE_MIN_REQS", "2")
    monkeypatch.setenv("VLLM_B200_CASCADE_MIN_TOKENS", "32")
    got, used1 = run()
    assert used0 == 0 and used1 > 0
    assert got == ref


def test_video_content_parts_decode(tmp_path, monkeypatch):
    """video_url parts: a base64 JPEG frame list, an OpenCV-decoded container, and an .npy frame array."""
    import base64
    import io
    import numpy as np
    from PIL import Image
    from vllm_b200.entrypoints.chat_utils import chat_multi_modal_data, load_video, sample_frame_indices
    assert sample_frame_indices(10, 4) == [0, 3, 6, 9] and sample_frame_indices(3, 8) == [0, 1, 2]
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 255, (6, 32, 48, 3), dtype=np.uint8)
    enc = []
    for f in frames:
        b = io.BytesIO()
        Image.fromarray(f).save(b, format="PNG")
        enc.append(base64.b64encode(b.getvalue()).decode())
    v = load_video("data:video/png;base64," + ",".join(enc), num_frames=3)
    assert v.shape == (3, 32, 48, 3) and (v[0] == frames[0]).all() and (v[-1] == frames[5]).all()
    np.save(tmp_path / "clip.npy", frames)
    import pytest as _pt
    with _pt.raises(ValueError):                                     # local files are refused by default
        load_video(f"file://{tmp_path}/clip.npy", 32)
    monkeypatch.setenv("VLLM_ALLOWED_LOCAL_MEDIA_PATH", str(tmp_path))
    with _pt.raises(ValueError):                                     # ... and confined to the configured directory
        load_video("file:///etc/hostname", 32)
    assert (load_video(f"file://{tmp_path}/clip.npy", 32) == frames).all()
    cv2 = pytest.importorskip("cv2")
    path = str(tmp_path / "clip.avi")
    w = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 4.0, (48, 32))
    if w.isOpened():
        for f in frames:
            w.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
        w.release()
        got = load_video("file://" + path, num_frames=4)
        assert got.shape == (4, 32, 48, 3)
    mm = chat_multi_modal_data([{"role": "user", "content": [
        {"type": "text", "text": "what happens?"},
        {"type": "video_url", "video_url": {"url": "data:video/png;base64," + ",".join(enc)}, "num_frames": 2}]}])
    assert list(mm) == ["video"] and mm["video"][0].shape == (2, 32, 48, 3)


def test_streaming_safetensors_loader(tmp_path):
    """--load-format streamer: concurrent positional reads with a bounded in-flight window return exactly the tensors
    of the checkpoint (all dtypes, multi-chunk tensors, several files) and load a model identically."""
    from safetensors.torch import save_file
    from tests.utils import make_hf_model, make_llm
    from vllm_b200 import SamplingParams
    from vllm_b200.model_loader import streamer
    torch.manual_seed(0)
    a = {"w.big": torch.randn(700, 1000), "w.bf16": torch.randn(33, 7).bfloat16(), "w.i64": torch.arange(11),
         "w.fp8": torch.randn(16, 16).to(torch.float8_e4m3fn), "w.empty": torch.zeros(0, 4), "w.bool": torch.tensor([True, False])}
    b = {"v.x": torch.randn(5, 5).half()}
    d = tmp_path / "ckpt"
    d.mkdir()
    save_file(a, str(d / "model-00001-of-00002.safetensors"))
    save_file(b, str(d / "model-00002-of-00002.safetensors"))
    old = streamer._CHUNK
    streamer._CHUNK = 64 << 10                     # force the big tensor through many chunked reads
    try:
        stats = {}
        got = dict(streamer.stream_safetensors(str(d), concurrency=8, memory_limit=1 << 20, stats=stats))
    finally:
        streamer._CHUNK = old
    ref = {**a, **b}
    assert set(got) == set(ref) and stats["tensors"] == len(ref)
    for k, v in ref.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape
        assert torch.equal(got[k].view(torch.uint8) if v.dtype == torch.float8_e4m3fn else got[k],
                           v.view(torch.uint8) if v.dtype == torch.float8_e4m3fn else v)
    assert stats["max_in_flight"] <= max(1 << 20, a["w.big"].numel() * 4)
    # end to end: a saved HF checkpoint served through the streaming loader
    hf = make_hf_model("llama-tiny")
    m = tmp_path / "llama"
    hf.save_pretrained(str(m), safe_serialization=True)
    sp = SamplingParams(max_tokens=6, temperature=0.0, ignore_eos=True)
    p = [{"prompt_token_ids": [5, 17, 99, 3, 250, 41]}]
    outs = {}
    for fmt in ("safetensors", "streamer", "fastsafetensors"):
        llm = make_llm(str(m), load_format=fmt, model_loader_extra_config={"concurrency": 4} if fmt != "safetensors" else None)
        outs[fmt] = llm.generate(p, sp)[0].outputs[0].token_ids
        llm.shutdown()
    assert outs["streamer"] == outs["safetensors"] == outs["fastsafetensors"]
    # the fastsafetensors iterator returns every tensor of a multi-file checkpoint
    from vllm_b200.model_loader import iter_weights
    got = {k: v.clone() for k, v in iter_weights(str(d), "fastsafetensors", {"files_per_batch": 2})}
    assert set(got) == set(ref)
    assert torch.equal(got["w.big"], ref["w.big"]) and torch.equal(got["v.x"], ref["v.x"])


def test_usage_stats_record_and_opt_out(tmp_path, monkeypatch):
    import json
    import os
    from tests.utils import make_llm
    from vllm_b200 import usage
    monkeypatch.setenv("VLLM_B200_CONFIG_ROOT", str(tmp_path))
    monkeypatch.setenv("VLLM_B200_NO_USAGE_STATS", "1")
    assert not usage.is_usage_stats_enabled()
    monkeypatch.setenv("VLLM_B200_NO_USAGE_STATS", "0")
    assert usage.is_usage_stats_enabled()
    (tmp_path / "do_not_track").write_text("")
    assert not usage.is_usage_stats_enabled()
    (tmp_path / "do_not_track").unlink()
    monkeypatch.setenv("DO_NOT_TRACK", "1")
    assert not usage.is_usage_stats_enabled()
    monkeypatch.delenv("DO_NOT_TRACK")
    llm = make_llm("llama-tiny")
    msg = usage.UsageMessage()
    usage.set_runtime_usage_data("api", "completions")
    assert msg.report_usage(llm.llm_engine.vllm_config, "LLM_CLASS", heartbeat_s=0.05)
    import time
    deadline = time.time() + 5
    while time.time() < deadline and (not os.path.exists(msg.path) or open(msg.path).read().count(msg.uuid) < 2):
        time.sleep(0.02)
    msg.stop()
    llm.shutdown()
    every = [json.loads(l) for l in open(msg.path).read().splitlines()]
    assert any(r["uuid"] != msg.uuid and r.get("context") == "LLM_CLASS" for r in every)     # the engine's own record
    lines = [r for r in every if r["uuid"] == msg.uuid]
    first = lines[0]
    assert first["model_architecture"] == "LlamaForCausalLM" and first["context"] == "LLM_CLASS"
    assert first["tensor_parallel_size"] == 1 and first["gpu_count"] == 0 and first["uuid"] == msg.uuid
    assert lines[1]["api"] == "completions" and lines[1]["uuid"] == msg.uuid
    assert "prompt" not in json.dumps(first)            # configuration facts only, never request content


def test_sparsemixer_routing_matches_hf():
    """The Phi-3.5-MoE router (threshold-masked sequential top-2) equals transformers' ``sparsemixer`` in eval mode, and
    FusedMoE uses a plugged-in router instead of softmax top-k."""
    from transformers.models.phimoe.modeling_phimoe import sparsemixer
    from vllm_b200.layers.moe import sparsemixer_routing
    torch.manual_seed(0)
    logits = torch.randn(64, 16) * 2
    logits[0, :2] = torch.tensor([3.0, 2.999])          # near tie: both survive the first mask
    w_ref, ids_ref = sparsemixer(logits.clone(), jitter_eps=0.01, training=False)
    w, ids = sparsemixer_routing(0.01)(logits, 2)
    assert torch.equal(ids.long(), ids_ref) and torch.allclose(w, w_ref, atol=1e-6)
    from tests.utils import make_llm
    llm = make_llm("mixtral-tiny")
    model = llm.llm_engine.engine_core.engine_core.model_executor.worker.runner.model
    moe = model.model.layers[0].block_sparse_moe.experts
    x = torch.randn(5, moe.w13_weight.shape[-1])
    rl = torch.randn(5, 4)
    base = moe(x, rl)
    moe.custom_routing = sparsemixer_routing(0.01)
    other = moe(x, rl)
    moe.custom_routing = lambda l, k: (torch.softmax(l.float(), -1).topk(k)[0] / torch.softmax(l.float(), -1).topk(k)[0].sum(-1, keepdim=True),
                                       torch.softmax(l.float(), -1).topk(k)[1].to(torch.int32))
    same = moe(x, rl)
    llm.shutdown()
    scale = base.abs().max().item() + 1e-30
    assert (same - base).abs().max().item() / scale < 1e-4           # plug-in reproducing softmax top-k renormalised
    assert (other - base).abs().max().item() / scale > 1e-3          # sparsemixer weights differ from it


def test_interleaved_gate_up_weights_match_plain_layout(monkeypatch):
    """gate|up stored as interleaved rows (the layout the fused SiLU·mul GEMM epilogue needs) is a pure re-ordering:
    same checkpoint, same tokens as the [gate | up] layout."""
    cfg = get_preset_hf_config("llama-tiny")
    w = list(llama_full_weights(cfg, seed=9))
    sp = SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True)
    outs = []
    for flag in ("0", "1"):
        monkeypatch.setenv("VLLM_B200_FUSED_ACT", flag)
        llm = make_llm("llama-tiny", weights=w, dtype="bfloat16")
        mlp = llm.llm_engine.engine_core.engine_core.model_executor.worker.runner.model.model.layers[0].mlp
        assert mlp.gate_up_proj.interleave == (flag == "1")
        outs.append([o.outputs[0].token_ids for o in llm.generate(PROMPTS, sp)])
        llm.shutdown()
    assert outs[0] == outs[1]


def test_batch_invariant_policy_and_generation(monkeypatch):
    """VLLM_BATCH_INVARIANT (reference: tests/v1/determinism/test_batch_invariance.py): the dispatch policy stops
    looking at the batch (no split-K / split-KV, rank-ordered collectives, no atomic MoE combine), and a prompt
    generates the same tokens and logprobs alone, first in a batch, and last in a batch of unrelated prompts."""
    import os
    from tests.utils import make_llm
    from vllm_b200 import SamplingParams, envs
    from vllm_b200.ops import cuda as cu
    from vllm_b200.utils import batch_invariant as bi
    monkeypatch.setattr(cu, "num_sms", lambda: 148)
    assert cu.decode_splits(4, 8, 32768, 64) > 1 and cu.mla_decode_splits(2, 128, 16384) > 1
    cu._plan_cache.clear()
    base = cu._plan(16, 4096, 4096)
    monkeypatch.setenv("VLLM_BATCH_INVARIANT", "1")
    envs._cache.clear()
    saved = {k: os.environ.get(k) for k in bi._NCCL}
    try:
        assert cu.batch_invariant() and bi.apply_env_overrides()
        assert os.environ["NCCL_ALGO"] == "allreduce:tree" and os.environ["NCCL_NVLS_ENABLE"] == "0"
        assert cu.decode_splits(4, 8, 32768, 64) == 1 and cu.mla_decode_splits(2, 128, 16384) == 1
        cu._plan_cache.clear()
        for M in (1, 16, 64, 256, 257, 4096):
            assert cu._plan(M, 4096, 4096)[2] == 1
        assert base[0] == 1          # (sanity: the default plan for 16 rows is the swap-AB decode kernel)
        cu._plan_cache.clear()
        llm = make_llm("llama-tiny", max_num_seqs=16)
        sp = SamplingParams(max_tokens=12, temperature=0.0, ignore_eos=True, logprobs=3)
        needle = {"prompt_token_ids": [11, 45, 7, 300, 21, 9, 77]}
        others = [{"prompt_token_ids": [(13 * i + 5 * j) % 500 + 3 for j in range(4 + 3 * i)]} for i in range(7)]

        def run(prompts, idx):
            o = llm.generate(prompts, sp)[idx].outputs[0]
            return list(o.token_ids), [lp[t].logprob for lp, t in zip(o.logprobs, o.token_ids)]
        alone = run([needle], 0)
        first = run([needle] + others, 0)
        last = run(others + [needle], len(others))
        llm.shutdown()
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        os.environ.pop("CUBLAS_WORKSPACE_CONFIG", None)
        monkeypatch.delenv("VLLM_BATCH_INVARIANT")
        envs._cache.clear()
        cu._plan_cache.clear()
    assert alone[0] == first[0] == last[0]
    for a, b in ((alone[1], first[1]), (alone[1], last[1])):
        assert max(abs(x - y) for x, y in zip(a, b)) < 1e-5


@pytest.mark.parametrize("preset", ["deepseek-tiny", "glm4moe-tiny", "qwen3next-tiny", "gemma3-tiny"])
def test_mtp_head_uses_the_targets_own_layer_kind(preset):
    """Multi-token prediction on families whose prediction block is NOT a Llama layer: the head builds one more layer
    through the target's layer factory (MLA + experts, q/k-normed GQA + experts, gated attention, sandwich-norm Gemma),
    keeps its own KV arena in the right layout (latent cache for MLA), drafts, and stays lossless."""
    llm = make_llm(preset)
    ref = _greedy(llm)
    llm.shutdown()
    llm = make_llm(preset, speculative_config=dict(method="mtp", num_speculative_tokens=2))
    runner = llm.llm_engine.engine_core.engine_core.model_executor.worker.runner
    head = runner.drafter.head
    assert head.native_layer
    kind = type(head.layers[0]).__name__
    assert kind != "LlamaDecoderLayer", kind
    got = _greedy(llm)
    sched = llm.llm_engine.engine_core.engine_core.scheduler
    llm.shutdown()
    assert got == ref
    assert sched._spec_draft_tokens > 0


def test_mtp_weights_are_picked_out_of_the_target_checkpoint():
    """DeepSeek-V3 style naming (``model.layers.<L>.{enorm,hnorm,eh_proj,shared_head.norm,self_attn…,mlp.experts…}``), the
    Qwen3-Next ``mtp.*`` naming and MiMo's ``model.mtp_layers.0.*``: every tensor lands in the head."""
    import torch
    from vllm_b200.spec_decode.eagle import EagleHead
    llm = make_llm("glm4moe-tiny", speculative_config=dict(method="mtp", num_speculative_tokens=1))
    head: EagleHead = llm.llm_engine.engine_core.engine_core.model_executor.worker.runner.drafter.head
    L = 3
    g = torch.Generator().manual_seed(0)
    src = {}
    for n, p in head.named_parameters():
        t = torch.randn(p.shape, generator=g).to(p.dtype)
        if n.startswith("layers.0."):
            rest = n[len("layers.0."):]
            if rest.endswith("experts.w13_weight"):
                I = t.shape[1] // 2
                for e in range(t.shape[0]):
                    src[f"model.layers.{L}.{rest[:-len('w13_weight')]}{e}.gate_proj.weight"] = t[e, :I]
                    src[f"model.layers.{L}.{rest[:-len('w13_weight')]}{e}.up_proj.weight"] = t[e, I:]
            elif rest.endswith("experts.w2_weight"):
                for e in range(t.shape[0]):
                    src[f"model.layers.{L}.{rest[:-len('w2_weight')]}{e}.down_proj.weight"] = t[e]
            elif rest.endswith("experts.e_score_correction_bias"):
                src[f"model.layers.{L}.mlp.gate.e_score_correction_bias"] = t
            elif ".qkv_proj." in rest:
                q, k, v = t.split([head.layers[0].self_attn.q_size, head.layers[0].self_attn.kv_size, head.layers[0].self_attn.kv_size], dim=0)
                for part, tt in (("q_proj", q), ("k_proj", k), ("v_proj", v)):
                    src[f"model.layers.{L}.{rest.replace('qkv_proj', part)}"] = tt
            elif ".gate_up_proj." in rest:
                a, b = t.chunk(2, dim=0)
                src[f"model.layers.{L}.{rest.replace('gate_up_proj', 'gate_proj')}"] = a
                src[f"model.layers.{L}.{rest.replace('gate_up_proj', 'up_proj')}"] = b
            else:
                src[f"model.layers.{L}.{rest}"] = t
        else:
            name = {"fc.weight": "eh_proj.weight", "norm.weight": "shared_head.norm.weight"}.get(n, n)
            src[f"model.layers.{L}.{name}"] = t
    src[f"model.layers.{L}.shared_head.head.weight"] = torch.zeros(4, 4)          # shared with the target: ignored
    src["model.layers.0.input_layernorm.weight"] = torch.zeros(4)                  # not part of the block: ignored
    want = {n: p.detach().clone() for n, p in head.named_parameters()}
    n = head.load_mtp_weights(src.items(), L)
    assert n >= len(src) - 2 - 1
    # round trip: reconstruct what was fed and compare with what the head now holds
    for name, p in head.named_parameters():
        assert not torch.equal(p, want[name]) or p.numel() <= 1, name
    llm.shutdown()
    # the other two naming schemes resolve to the same targets
    picked = []
    orig = EagleHead.load_weights
    try:
        EagleHead.load_weights = lambda self, ws: picked.extend(n for n, _ in ws) or 0
        head.load_mtp_weights([("mtp.fc.weight", 0), ("mtp.pre_fc_norm_hidden.weight", 0), ("mtp.pre_fc_norm_embedding.weight", 0),
                               ("mtp.norm.weight", 0), ("mtp.layers.0.self_attn.q_proj.weight", 0),
                               ("model.mtp_layers.0.token_layernorm.weight", 0), ("model.mtp_layers.0.input_proj.weight", 0),
                               ("model.mtp_layers.0.final_layernorm.weight", 0), ("model.mtp_layers.0.mlp.down_proj.weight", 0)], 99)
    finally:
        EagleHead.load_weights = orig
    assert picked == ["fc.weight", "hnorm.weight", "enorm.weight", "norm.weight", "layers.0.self_attn.q_proj.weight",
                      "enorm.weight", "fc.weight", "norm.weight", "layers.0.mlp.down_proj.weight"]


def test_speculative_method_is_inferred_from_the_draft_architecture(tmp_path):
    import json
    from vllm_b200.config import SpeculativeConfig
    from vllm_b200.spec_decode import DRAFT_ARCHS, infer_method
    assert SpeculativeConfig(method="auto").method == "ngram"
    assert SpeculativeConfig(method="auto", model="llama-tiny").method == "draft_model"
    for arch, want in (("EagleLlamaForCausalLM", "eagle"), ("Eagle3Qwen3ForCausalLM", "eagle3"), ("DeepSeekMTPModel", "mtp"),
                       ("Glm4MoeMTPModel", "mtp"), ("MedusaModel", "medusa"), ("LlamaForCausalLM", "draft_model")):
        d = tmp_path / arch
        d.mkdir()
        (d / "config.json").write_text(json.dumps({"architectures": [arch]}))
        assert infer_method(str(d)) == want, arch
    d = tmp_path / "yuhuili-EAGLE3-x"
    d.mkdir()
    (d / "config.json").write_text(json.dumps({"architectures": ["LlamaForCausalLM"]}))
    assert infer_method(str(d)) == "eagle3"
    assert len(DRAFT_ARCHS) > 40


def test_reference_style_engine_flags():
    """--spec-method / --spec-tokens shorthand, --limit-mm-per-prompt, --disable-sliding-window, multi-node fields and the
    compatibility-only flags parse and take effect."""
    import argparse
    from vllm_b200.engine.arg_utils import EngineArgs
    ap = EngineArgs.add_cli_args(argparse.ArgumentParser())
    ns = ap.parse_args(["--model", "llama-tiny", "--spec-method", "ngram", "--spec-tokens", "3", "--revision", "main", "--hf-token", "x",
                        "--nnodes", "2", "--node-rank", "1", "--master-addr", "10.0.0.1", "--limit-mm-per-prompt", '{"image": 1}',
                        "--no-enable-sleep-mode"])
    ea = EngineArgs.from_cli_args(ns)
    cfg = ea.create_engine_config()
    assert cfg.speculative_config.method == "ngram" and cfg.speculative_config.num_speculative_tokens == 3
    assert cfg.parallel_config.nnodes == 2 and cfg.parallel_config.node_rank == 1 and cfg.parallel_config.master_addr == "10.0.0.1"
    assert cfg.model_config.hf_config["limit_mm_per_prompt"] == {"image": 1}
    cfg = EngineArgs(model="ministral-tiny", disable_sliding_window=True, device="cpu").create_engine_config()
    assert cfg.model_config.hf_config["sliding_window"] is None and cfg.model_config.max_model_len == 16
    llm = make_llm("llava-tiny", limit_mm_per_prompt={"image": 1})
    img = torch.randn(3, 28, 28)
    with pytest.raises(ValueError, match="at most 1 image"):
        llm.generate([{"prompt_token_ids": [5, 511, 511, 9], "multi_modal_data": {"image": [img, img]}}], SamplingParams(max_tokens=2))
    out = llm.generate([{"prompt_token_ids": [5, 511, 9], "multi_modal_data": {"image": [img]}}], SamplingParams(max_tokens=2))
    assert len(out[0].outputs[0].token_ids) == 2
    llm.shutdown()
