"""Engine features and model families executed on the B200 (CUDA kernels + CUDA graphs), each checked against the
plain-PyTorch reference ops or against its own lossless invariant.

Mirrors the reference's per-feature e2e suites (tests/v1/e2e/, tests/lora/, tests/v1/spec_decode/,
tests/models/language/generation/) on the small presets: what the CPU suite proves for the reference path, this file
proves for the device path.
"""
import os
import re

import pytest
import torch

pytestmark = pytest.mark.gpu

PROMPTS = [{"prompt_token_ids": [(13 * i + 7 * j) % 200 + 5 for j in range(20 + 23 * i)]} for i in range(4)]


def _wide(model):
    """hf_overrides that give a tiny preset production-like attention shapes (128-wide heads, 512-wide residual) so
    that the tcgen05 / vectorised kernels — not the small-shape fallbacks — are what runs."""
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config(model)
    if any(k in cfg for k in ("mamba_d_state", "state_size", "ssm_cfg", "linear_num_value_heads", "mamba_n_heads",
                              "attn_layer_indices", "layers_block_type", "attn_layer_period")):
        return {}                                  # SSM hybrids: widths are tied to the mixer geometry, keep the preset
    ov = {}
    heads = cfg.get("num_attention_heads") or cfg.get("n_head") or cfg.get("n_heads") or 4
    for key in ("hidden_size", "n_embd", "d_model"):
        if key in cfg:
            ov[key] = 128 * heads
    if "head_dim" in cfg:
        ov["head_dim"] = 128
    for key in ("intermediate_size", "ffn_dim", "n_inner"):
        if cfg.get(key):
            ov[key] = 1024
    if cfg.get("sliding_window"):
        ov["sliding_window"] = 96
    return ov


def _llm(model, weights=None, seed_dummy=7, **kw):
    from vllm_b200 import LLM
    d = dict(max_model_len=512, max_num_seqs=16, max_num_batched_tokens=512, gpu_memory_utilization=0.2,
             load_format="dummy", cudagraph_capture_sizes=[1, 2, 4, 8], dtype="bfloat16")
    d.update(kw)
    llm = LLM(model, **d)
    runner = llm.llm_engine.engine_core.engine_core.model_executor.worker.runner
    with torch.no_grad():
        if weights is not None:
            runner.model.load_weights(weights)
        else:
            from vllm_b200.model_loader import initialize_dummy_weights
            initialize_dummy_weights(runner.model, seed_dummy, -0.05, 0.05)
    return llm, runner


def _prompt_lps(outs):
    res = []
    for o in outs:
        for d in o.prompt_logprobs[1:]:
            (t, l), = [(t, l.logprob) for t, l in d.items() if l.rank is not None][:1]
            res.append((t, l, d))
    return res


FAMILIES = ["mamba2-tiny", "bamba-tiny", "jamba-tiny", "qwen3next-tiny", "granitemoehybrid-tiny",      # SSM / hybrid
            "gptoss-tiny", "gemma3-tiny", "gemma2-tiny", "cohere2-tiny",                               # sinks / sliding / softcap
            "qwen3moe-tiny", "llama4-tiny", "glm4moe-tiny", "olmoe-tiny",                               # MoE routers
            "phi3-tiny", "gptneox-tiny", "falcon-7b-tiny", "bloom-tiny", "mpt-tiny", "gpt2-tiny", "opt-tiny"]  # dense variants


@pytest.mark.parametrize("model", FAMILIES)
def test_model_family_device_path_matches_reference_ops(model):
    """Teacher-forced prompt logprobs: CUDA kernels + decode graphs vs ``VLLM_B200_FORCE_REF_OPS`` (plain PyTorch) on
    the same GPU and the same seeded weights."""
    from vllm_b200 import SamplingParams
    from vllm_b200.ops import cuda as cu
    sp = SamplingParams(max_tokens=6, temperature=0.0, ignore_eos=True, prompt_logprobs=0)
    ov = _wide(model)
    llm, runner = _llm(model, hf_overrides=ov)
    n0 = cu.LAUNCHES
    out = llm.generate(PROMPTS, sp)
    assert cu.LAUNCHES > n0, "no kernel of ours was launched"
    llm.shutdown()
    del runner
    os.environ["VLLM_B200_FORCE_REF_OPS"] = "1"
    try:
        llm_r, _ = _llm(model, enforce_eager=True, hf_overrides=ov)
        out_r = llm_r.generate(PROMPTS, sp)
        llm_r.shutdown()
    finally:
        os.environ.pop("VLLM_B200_FORCE_REF_OPS")
    diffs = sorted(abs(la - db[ta].logprob) for (ta, la, _), (_, _, db) in zip(_prompt_lps(out), _prompt_lps(out_r)))
    assert diffs[int(0.9 * len(diffs))] < 0.15 and diffs[len(diffs) // 2] < 0.05, (diffs[len(diffs) // 2], diffs[-5:])


@pytest.mark.parametrize("spec", [dict(method="ngram", num_speculative_tokens=3, prompt_lookup_max=4, prompt_lookup_min=1),
                                  dict(method="draft_model", model="llama-tiny", num_speculative_tokens=3),
                                  dict(method="eagle", num_speculative_tokens=3),
                                  dict(method="mtp", num_speculative_tokens=2)])
def test_speculative_decoding_lossless_on_device(spec):
    """Greedy with speculation == greedy without, with the verifier, rejection sampler and proposer on the GPU."""
    from tests.utils import llama_full_weights
    from vllm_b200 import SamplingParams
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config("llama-small")
    w = list(llama_full_weights(cfg, seed=2, scale=0.05))
    prompts = [{"prompt_token_ids": [(5 * i + j) % 17 + 3 for i in range(40)]} for j in range(3)]      # repetitive: n-gram hits
    sp = SamplingParams(max_tokens=24, temperature=0.0, ignore_eos=True, logprobs=0)
    llm, _ = _llm("llama-small", weights=w)
    ref = llm.generate(prompts, sp)
    llm.shutdown()
    if spec.get("method") == "draft_model":
        spec = dict(spec, model="llama-small")
    llm, _ = _llm("llama-small", weights=w, speculative_config=spec)
    got = llm.generate(prompts, sp)
    sched = llm.llm_engine.engine_core.engine_core.scheduler
    drafts = sched._spec_draft_tokens
    llm.shutdown()
    assert drafts > 0, "the proposer never produced a draft"
    # bf16 batched verification vs one-token decode may flip a near-tie: compare up to the first tie, and bound the
    # logprob of the emitted tokens
    for a, b in zip(got, ref):
        ta, tb = a.outputs[0].token_ids, b.outputs[0].token_ids
        n = next((i for i, (x, y) in enumerate(zip(ta, tb)) if x != y), len(ta))
        if n < len(ta):
            la = a.outputs[0].logprobs[n][ta[n]].logprob
            lb = b.outputs[0].logprobs[n][tb[n]].logprob
            assert abs(la - lb) < 0.05, (n, la, lb)      # a genuine near-tie, not a wrong token
        assert n >= 4


def test_multi_lora_on_device_matches_merged_weights():
    """Mixed batch (base + two adapters) through the BGMV kernels under CUDA graphs == runs with merged weights."""
    from tests.utils import llama_full_weights
    from vllm_b200 import SamplingParams
    from vllm_b200.config import get_preset_hf_config
    from vllm_b200.core.request import LoRARequest
    cfg = get_preset_hf_config("llama-small")
    base = dict(llama_full_weights(cfg, seed=5, scale=0.05))
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

    ad = {1: (adapter(8, 0.05), 8, 16.0), 2: (adapter(16, 0.04), 16, 16.0)}

    def merged(lid):
        t, r, alpha = ad[lid]
        w = dict(base)
        for k in list(t):
            if ".lora_A" in k:
                name = k[len("base_model.model."):].split(".lora_A")[0] + ".weight"
                w[name] = w[name] + (alpha / r) * t[k.replace("lora_A", "lora_B")] @ t[k]
        return w

    prompt = PROMPTS[1]
    sp = SamplingParams(max_tokens=4, temperature=0.0, ignore_eos=True, prompt_logprobs=0)
    exp = []
    for w in (base, merged(1), merged(2)):
        llm, _ = _llm("llama-small", weights=w.items())
        exp.append(llm.generate([prompt], sp)[0])
        llm.shutdown()
    llm, runner = _llm("llama-small", weights=base.items(), enable_lora=True, max_loras=2, max_lora_rank=16)
    worker = llm.llm_engine.engine_core.engine_core.model_executor.worker
    reqs = {i: LoRARequest(f"ad{i}", i, "") for i in (1, 2)}
    for i in (1, 2):
        worker.add_lora(reqs[i], tensors=ad[i][0], rank=ad[i][1], alpha=ad[i][2])
    outs = llm.generate([prompt, prompt, prompt], sp, lora_request=[None, reqs[1], reqs[2]])
    llm.shutdown()
    lp = [[l for _, l, _ in _prompt_lps([o])] for o in outs]
    le = [[l for _, l, _ in _prompt_lps([o])] for o in exp]
    for got, want in zip(lp, le):
        d = sorted(abs(a - b) for a, b in zip(got, want))
        assert d[len(d) // 2] < 0.03 and d[-1] < 0.25, d[-5:]
    # the adapters matter: adapter rows differ from the base rows by far more than the tolerance above
    assert max(abs(a - b) for a, b in zip(le[0], le[1])) > 0.3 and max(abs(a - b) for a, b in zip(le[1], le[2])) > 0.3


def test_sleep_wake_prefix_cache_and_kv_fp8_on_device():
    from vllm_b200 import SamplingParams
    sp = SamplingParams(max_tokens=6, temperature=0.0, ignore_eos=True)
    llm, runner = _llm("llama-small", enable_prefix_caching=True)
    p = [{"prompt_token_ids": [(7 * j) % 300 + 3 for j in range(200)]}]
    a = llm.generate(p, sp)[0]
    b = llm.generate(p, sp)[0]                                    # second run hits the prefix cache
    assert a.outputs[0].token_ids == b.outputs[0].token_ids
    assert (b.num_cached_tokens or 0) >= 128
    free0 = torch.cuda.mem_get_info()[0]
    llm.sleep(level=1)
    worker = llm.llm_engine.engine_core.engine_core.model_executor.worker
    assert worker.is_sleeping()
    assert torch.cuda.mem_get_info()[0] > free0 + (64 << 20), "sleep did not give device memory back"
    llm.wake_up()
    c = llm.generate(p, sp)[0]
    assert c.outputs[0].token_ids == a.outputs[0].token_ids
    llm.shutdown()
    del runner
    # fp8 KV cache: teacher-forced logprobs stay within fp8 noise of the bf16 cache (chunked prefill reads the cache)
    spl = SamplingParams(max_tokens=2, temperature=0.0, ignore_eos=True, prompt_logprobs=0)
    res = {}
    for kvd in ("auto", "fp8"):
        llm, r = _llm("llama-small", kv_cache_dtype=kvd, max_num_batched_tokens=64)
        res[kvd] = [l for _, l, _ in _prompt_lps(llm.generate(p, spl))]
        if kvd == "fp8":
            assert next(iter(r.kv_caches.values())).dtype == torch.float8_e4m3fn
        llm.shutdown()
        del r
    d = sorted(abs(x - y) for x, y in zip(res["auto"], res["fp8"]))
    assert 0 < d[-1] and d[len(d) // 2] < 0.08 and d[int(0.95 * len(d))] < 0.3, d[-5:]


def test_structured_output_and_penalties_on_device():
    from vllm_b200 import SamplingParams
    from vllm_b200.sampling_params import StructuredOutputsParams
    llm, _ = _llm("llama-small")
    for seed in range(2):
        sp = SamplingParams(max_tokens=40, temperature=0.9, seed=seed,
                            structured_outputs=StructuredOutputsParams(regex=r"[0-9]{3}-[a-c]{2}"))
        o = llm.generate([{"prompt": "hello"}], sp)[0].outputs[0]
        assert re.fullmatch(r"[0-9]{3}-[a-c]{2}", o.text) and o.finish_reason == "stop", o.text
        sp = SamplingParams(max_tokens=40, temperature=0.9, seed=seed,
                            structured_outputs=StructuredOutputsParams(choice=["yes", "no", "maybe"]))
        assert llm.generate([{"prompt": "hello"}], sp)[0].outputs[0].text in ("yes", "no", "maybe")
    # penalties / bias / allowed ids inside the fused sampler: allowed set is respected, banned token never appears
    sp = SamplingParams(max_tokens=24, temperature=0.8, seed=1, ignore_eos=True, allowed_token_ids=list(range(3, 40)),
                        presence_penalty=0.7, repetition_penalty=1.2, logit_bias={7: 4.0, 9: -100.0}, logprobs=2)
    o = llm.generate(PROMPTS[:3], sp)
    for r in o:
        ids = r.outputs[0].token_ids
        assert all(3 <= t < 40 and t != 9 for t in ids), ids
        assert len(r.outputs[0].logprobs) == len(ids)
    # n > 1 and min_p / top-k / top-p are reproducible under a seed
    sp = SamplingParams(max_tokens=8, temperature=1.0, n=3, seed=5, top_k=50, top_p=0.9, min_p=0.01, ignore_eos=True)
    a = llm.generate(PROMPTS[:1], sp)[0].outputs
    b = llm.generate(PROMPTS[:1], sp)[0].outputs
    assert [x.token_ids for x in a] == [x.token_ids for x in b] and len(a) == 3
    llm.shutdown()


def test_pooling_and_whisper_and_vlm_on_device():
    """Encoder-only pooling, encoder-decoder cross attention, and an M-RoPE VLM all run on the device path."""
    from vllm_b200 import LLM, SamplingParams
    llm = LLM("bert-tiny", max_model_len=128, gpu_memory_utilization=0.1, load_format="dummy", dtype="bfloat16")
    outs = llm.embed([{"prompt_token_ids": [3, 4, 5, 6, 7]}, {"prompt_token_ids": [9, 8, 7]}])
    e = torch.stack([torch.as_tensor(o.outputs.data).float() for o in outs])
    assert e.shape[0] == 2 and torch.isfinite(e).all() and (e.norm(dim=-1) > 0).all()
    llm.shutdown()
    import numpy as np
    llm, _ = _llm("whisper-tiny-test", max_model_len=64, max_num_batched_tokens=2048)
    audio = (np.sin(np.arange(16000) * 0.05).astype("float32"), 16000)
    o = llm.generate([{"prompt_token_ids": [1, 2, 3], "multi_modal_data": {"audio": audio}}],
                     SamplingParams(max_tokens=5, temperature=0.0, ignore_eos=True))
    assert len(o[0].outputs[0].token_ids) == 5
    llm.shutdown()
    llm, _ = _llm("qwen2vl-tiny", hf_overrides=_wide("qwen2vl-tiny"))
    o = llm.generate(PROMPTS[:2], SamplingParams(max_tokens=5, temperature=0.0, ignore_eos=True))
    assert all(len(x.outputs[0].token_ids) == 5 for x in o)
    llm.shutdown()
