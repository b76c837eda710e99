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
