"""Bitwise batch-invariance check on a GPU (VLLM_BATCH_INVARIANT=1).

    VLLM_BATCH_INVARIANT=1 python scripts/batch_invariance_check.py [--model llama-3-8b-4l] [--tp 1]

Generates from one prompt (a) alone, (b) first in a batch of 48 unrelated prompts of mixed length, (c) last in that batch,
(d) arriving while the others are mid-decode, and compares the sampled-token logprobs bit for bit.  Prints one JSON
line: {"invariant": bool, "max_abs_diff": {...}}.  Without the flag the same script reports how far the default
dispatch drifts (split-K / split-KV / in-switch reductions), which is the number the mode exists to remove.
Reference: tests/v1/determinism/test_batch_invariance.py.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("VLLM_LOGGING_LEVEL", "WARNING")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama-3-8b-4l")
    ap.add_argument("--tp", type=int, default=1)
    ap.add_argument("--tokens", type=int, default=32)
    a = ap.parse_args()
    from vllm_b200 import LLM, SamplingParams
    llm = LLM(a.model, load_format="dummy", tensor_parallel_size=a.tp, max_model_len=4096, max_num_seqs=64,
              gpu_memory_utilization=0.5)
    V = llm.llm_engine.vllm_config.model_config.get_vocab_size()
    sp = SamplingParams(max_tokens=a.tokens, temperature=0.0, ignore_eos=True, logprobs=1)
    needle = {"prompt_token_ids": [(17 * i + 3) % (V - 10) + 5 for i in range(301)]}
    others = [{"prompt_token_ids": [(31 * i + 7 * j) % (V - 10) + 5 for j in range(16 + 61 * (i % 9))]} for i in range(48)]

    def lps(out):
        o = out.outputs[0]
        return list(o.token_ids), [lp[t].logprob for lp, t in zip(o.logprobs, o.token_ids)]
    ref = lps(llm.generate([needle], sp)[0])
    runs = {"first": lps(llm.generate([needle] + others, sp)[0]),
            "last": lps(llm.generate(others + [needle], sp)[-1])}
    # (d) staggered arrival: the others are already decoding when the needle is admitted
    eng = llm.llm_engine
    for i, p in enumerate(others):
        eng.add_request(f"o{i}", p, SamplingParams(max_tokens=a.tokens * 2, temperature=0.0, ignore_eos=True))
    for _ in range(6):
        eng.step()
    eng.add_request("needle", needle, sp)
    got = None
    while eng.has_unfinished_requests():
        for o in eng.step():
            if o.request_id == "needle" and o.finished:
                got = lps(o)
    runs["staggered"] = got
    diff = {k: (max(abs(x - y) for x, y in zip(ref[1], v[1])) if v[0] == ref[0] else "tokens differ")
            for k, v in runs.items()}
    print(json.dumps({"batch_invariant_flag": bool(int(os.environ.get("VLLM_BATCH_INVARIANT", "0"))), "tp": a.tp,
                      "invariant": all(d == 0.0 for d in diff.values()), "max_abs_diff": diff}))
    llm.shutdown()


if __name__ == "__main__":
    main()
