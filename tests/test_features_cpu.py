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
                                  dict(method="draft_model", model="llama-tiny", num_speculative_tokens=3)])
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


@pytest.mark.parametrize("connector,extra", [
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
    if connector == "SharedStorageConnector":
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
