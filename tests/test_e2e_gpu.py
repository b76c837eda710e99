"""End-to-end engine tests on the B200: CUDA kernels vs reference ops, CUDA-graph replay vs eager."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _llm(model="llama-small", **kw):
    from vllm_b200 import LLM
    d = dict(max_model_len=1024, max_num_seqs=16, max_num_batched_tokens=1024, gpu_memory_utilization=0.25,
             load_format="dummy")
    d.update(kw)
    llm = LLM(model, **d)
    runner = llm.llm_engine.engine_core.engine_core.model_executor.worker.runner
    from vllm_b200.model_loader import initialize_dummy_weights
    with torch.no_grad():
        initialize_dummy_weights(runner.model, 7, -0.05, 0.05)  # non-degenerate logits
    return llm, runner


PROMPTS = [{"prompt_token_ids": [(13 * i + 7 * j) % 4000 + 5 for j in range(20 + 37 * i)]} for i in range(5)]


def test_native_library_loaded():
    from vllm_b200.ops import cuda_ops
    assert os.path.basename(cuda_ops()._LIB_PATH) == "_C.so"
    maps = open("/proc/self/maps").read()
    assert "vllm_b200/_C.so" in maps


def test_graph_equals_eager_and_ref_logprobs():
    from vllm_b200 import SamplingParams
    sp = SamplingParams(max_tokens=12, temperature=0.0, ignore_eos=True, logprobs=0, prompt_logprobs=0)
    llm_g, runner = _llm(cudagraph_capture_sizes=[1, 2, 4, 8])
    assert runner.graphs, "CUDA graphs were not captured"
    out_g = llm_g.generate(PROMPTS, SamplingParams(max_tokens=12, temperature=0.0, ignore_eos=True))
    n0 = runner.launch_count
    assert n0 > 0
    out_g_lp = llm_g.generate(PROMPTS, sp)
    llm_g.shutdown()
    del runner
    llm_e, _ = _llm(enforce_eager=True)
    out_e = llm_e.generate(PROMPTS, SamplingParams(max_tokens=12, temperature=0.0, ignore_eos=True))
    llm_e.shutdown()
    for a, b in zip(out_g, out_e):
        assert a.outputs[0].token_ids == b.outputs[0].token_ids
    # reference ops (plain PyTorch on the same GPU): logprobs of the chosen tokens must agree
    os.environ["VLLM_B200_FORCE_REF_OPS"] = "1"
    try:
        llm_r, _ = _llm(enforce_eager=True)
        out_r = llm_r.generate(PROMPTS, sp)
        llm_r.shutdown()
    finally:
        os.environ.pop("VLLM_B200_FORCE_REF_OPS")
    agree = total = 0
    for a, b in zip(out_g_lp, out_r):
        for pa, pb in zip(a.prompt_logprobs[1:], b.prompt_logprobs[1:]):
            (ta, la), = [(t, l.logprob) for t, l in pa.items() if l.rank is not None][:1]
            lb = pb[ta].logprob
            assert abs(la - lb) < 0.08, (la, lb)
        ta, tb = a.outputs[0].token_ids, b.outputs[0].token_ids
        n = next((i for i, (x, y) in enumerate(zip(ta, tb)) if x != y), len(ta))
        agree += n
        total += len(ta)
    assert agree / total > 0.7  # bf16 rounding may flip a near-tie late in a sequence


def test_random_sampling_seeded_and_chunked_prefill():
    from vllm_b200 import SamplingParams
    llm, runner = _llm(max_num_batched_tokens=64, cudagraph_capture_sizes=[8])
    sp = SamplingParams(max_tokens=10, temperature=0.9, seed=11, ignore_eos=True)
    a = llm.generate(PROMPTS, sp)
    b = llm.generate(PROMPTS, sp)
    assert [o.outputs[0].token_ids for o in a] == [o.outputs[0].token_ids for o in b]
    g = llm.generate(PROMPTS, SamplingParams(max_tokens=10, temperature=0.0, ignore_eos=True))
    assert [o.outputs[0].token_ids for o in a] != [o.outputs[0].token_ids for o in g]
    # top-k / top-p filtered sampling runs and is reproducible under a seed (bf16 logits tie often, and ties
    # inside the kept set are broken by the noise, so top_k=1 is deliberately NOT compared with greedy)
    sp2 = SamplingParams(max_tokens=10, temperature=1.2, top_k=20, top_p=0.9, seed=3, ignore_eos=True)
    t1 = llm.generate(PROMPTS, sp2)
    t2 = llm.generate(PROMPTS, sp2)
    assert [o.outputs[0].token_ids for o in t1] == [o.outputs[0].token_ids for o in t2]
    llm.shutdown()


@pytest.mark.parametrize("model", ["mixtral-small", "deepseek-small"])
def test_moe_and_mla_models_cuda_vs_ref(model):
    """MoE (grouped tcgen05 GEMMs, device-side routing) and MLA models: CUDA path vs plain-PyTorch ops on the
    same weights — teacher-forced prompt logprobs must agree."""
    from vllm_b200 import SamplingParams
    sp = SamplingParams(max_tokens=6, temperature=0.0, ignore_eos=True, prompt_logprobs=0)
    llm, runner = _llm(model, cudagraph_capture_sizes=[1, 2, 4, 8])
    out = llm.generate(PROMPTS, sp)
    llm.shutdown()
    del runner
    os.environ["VLLM_B200_FORCE_REF_OPS"] = "1"
    try:
        llm_r, _ = _llm(model, enforce_eager=True)
        out_r = llm_r.generate(PROMPTS, sp)
        llm_r.shutdown()
    finally:
        os.environ.pop("VLLM_B200_FORCE_REF_OPS")
    # expert routing is discontinuous: a near-tie in a router logit may pick a different expert in the two
    # implementations and move that position a lot, so bound the bulk of the distribution, not the worst element
    diffs = []
    for a, b in zip(out, out_r):
        for pa, pb in zip(a.prompt_logprobs[1:], b.prompt_logprobs[1:]):
            (ta, la), = [(t, l.logprob) for t, l in pa.items() if l.rank is not None][:1]
            diffs.append(abs(la - pb[ta].logprob))
    diffs.sort()
    assert diffs[int(0.9 * len(diffs))] < 0.15 and diffs[len(diffs) // 2] < 0.05, (diffs[len(diffs) // 2], diffs[-5:])


@pytest.mark.parametrize("model,quant", [("deepseek-small", "fp8_block"), ("llama-small", "fp8"), ("mixtral-small", "fp8_block"),
                                         ("llama-small", "nvfp4"), ("llama-small", "mxfp4")])
def test_fp8_models_run_on_own_kernels_and_track_bf16(model, quant, monkeypatch):
    """Online-quantised fp8 runs (block-scaled tcgen05 f8f6f4 GEMMs, fused quantisers, grouped fp8 experts, MLA kernel)
    stay within fp8 accuracy of the bf16 run of the same weights: teacher-forced prompt logprobs."""
    from vllm_b200 import SamplingParams
    from vllm_b200 import LLM, envs
    # the bf16 run would store gate|up row-interleaved (fused GEMM epilogue) and the quantised run would not: the
    # dummy loader fills the stored tensor, so keep both runs on the plain layout to compare the SAME weights
    monkeypatch.setenv("VLLM_B200_FUSED_ACT", "0")
    envs._cache.pop("VLLM_B200_FUSED_ACT", None)
    sp = SamplingParams(max_tokens=4, temperature=0.0, ignore_eos=True, prompt_logprobs=0)
    kw = dict(max_model_len=1024, max_num_seqs=16, max_num_batched_tokens=1024, gpu_memory_utilization=0.25,
              load_format="dummy", model_loader_extra_config={"dummy_range": 0.05}, cudagraph_capture_sizes=[1, 2, 4, 8])
    llm = LLM(model, **kw)                    # same seeded random weights in both runs; the second quantises them online
    base = llm.generate(PROMPTS, sp)
    llm.shutdown()
    llm_q = LLM(model, quantization=quant, **kw)
    n0 = __import__("vllm_b200.ops.cuda", fromlist=["x"]).LAUNCHES
    out = llm_q.generate(PROMPTS, sp)
    llm_q.shutdown()
    assert __import__("vllm_b200.ops.cuda", fromlist=["x"]).LAUNCHES > n0
    diffs = []
    for a, b in zip(out, base):
        for pa, pb in zip(a.prompt_logprobs[1:], b.prompt_logprobs[1:]):
            (ta, la), = [(t, l.logprob) for t, l in pa.items() if l.rank is not None][:1]
            diffs.append(abs(la - pb[ta].logprob))
    diffs.sort()
    tol = (0.12, 0.6) if "fp8" in quant else (0.5, 1.5)          # 4-bit weights: coarser, still tracking
    assert diffs[len(diffs) // 2] < tol[0] and diffs[int(0.9 * len(diffs))] < tol[1], (diffs[len(diffs) // 2], diffs[-5:])


def test_smoke_entry():
    import __graft_entry__ as ge
    ge.smoke()


@pytest.mark.parametrize("model", ["llama-small", "bamba-tiny"])
def test_gpu_input_prep_equals_host_prep(model, monkeypatch):
    """Device-resident token / block tables + on-device expansion (csrc/input_prep.cu) produce exactly the tokens of the
    host-side numpy expansion: chunked prefill, more requests than rows (queueing, row moves on completion), prefix
    caching, CUDA-graph padding, and a hybrid model with several KV-cache groups."""
    from vllm_b200 import SamplingParams, envs
    prompts = [{"prompt_token_ids": [(11 * i + 5 * j) % 300 + 5 for j in range(30 + 41 * i)]} for i in range(9)]
    prompts += prompts[:3]                                     # repeated prompts: prefix-cache hits
    sps = [SamplingParams(max_tokens=5 + 3 * (i % 4), temperature=0.0, ignore_eos=True) for i in range(len(prompts))]
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("VLLM_B200_GPU_INPUT_PREP", mode)
        envs._cache.pop("VLLM_B200_GPU_INPUT_PREP", None)
        llm, runner = _llm(model, max_num_seqs=4, max_num_batched_tokens=64, enable_prefix_caching=(model == "llama-small"),
                           cudagraph_capture_sizes=[1, 2, 4])
        assert runner.gpu_prep == (mode == "1")
        res[mode] = [o.outputs[0].token_ids for o in llm.generate(prompts, sps)]
        llm.shutdown()
        del runner
    assert res["1"] == res["0"]


def test_mixed_batch_cuda_graphs_match_eager():
    """cudagraph_mode=FULL: steps that mix prefill chunks with decodes replay one captured graph per token bucket (device-
    side input expansion with bucket padding, flat varlen attention schedule, device-side last-token gather) and
    produce the tokens and logprobs of the eager run — chunked prefill, queueing (more requests than rows), prefix-cache
    hits, buckets that are far from full."""
    from vllm_b200 import SamplingParams
    prompts = [{"prompt_token_ids": [(11 * i + 5 * j) % 300 + 5 for j in range(17 + 29 * i)]} for i in range(10)]
    prompts += prompts[:2]
    sps = [SamplingParams(max_tokens=6 + 3 * (i % 4), temperature=0.0, ignore_eos=True, logprobs=0) for i in range(len(prompts))]
    kw = dict(max_num_seqs=6, max_num_batched_tokens=96, enable_prefix_caching=True, cudagraph_capture_sizes=[1, 2, 4, 8])
    llm, runner = _llm(cudagraph_mode="FULL", mixed_capture_sizes=[32, 64, 96], **kw)
    assert sorted(runner.mixed_graphs) == [32, 64, 96]
    n_eager0 = runner.launch_count
    plain = [SamplingParams(max_tokens=s.max_tokens, temperature=0.0, ignore_eos=True) for s in sps]
    got = llm.generate(prompts, plain)                           # logprob-free requests: eligible for the graphs
    replays = sum(getattr(g, "replays", 0) for g in runner.mixed_graphs.values())
    llm.shutdown()
    del runner
    llm_e, _ = _llm(enforce_eager=True, **kw)
    exp = llm_e.generate(prompts, sps)
    llm_e.shutdown()
    assert replays > 0, "no mixed step went through a captured graph"
    for a, b in zip(got, exp):
        ta, tb = a.outputs[0].token_ids, b.outputs[0].token_ids
        n = next((i for i, (x, y) in enumerate(zip(ta, tb)) if x != y), len(ta))
        if n < len(tb):       # a flip must be a genuine near-tie in the eager run's distribution
            lp = b.outputs[0].logprobs[n]
            assert ta[n] in lp or abs(list(lp.values())[0].logprob) > 0.5, (n, ta, tb)
        assert n >= min(4, len(tb))
