"""Host cost of preparing one engine step: numpy expansion + per-token H2D (VLLM_B200_GPU_INPUT_PREP=0) vs the
device-resident path (four ints per request + block-table diffs, expansion by csrc/input_prep.cu).

    python scripts/host_prep_bench.py            # prints one JSON line per (batch, mode)

Measures `ModelRunner.prep_host_s` (wall time inside _prepare/_h2d or their GPU-path counterparts) per decode step
on llama-small at B = 64 / 256 / 1024 running requests with ~600-token contexts, and the H2D bytes per step.
"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(B: int) -> dict:
    import torch
    from vllm_b200 import LLM, SamplingParams
    llm = LLM("llama-small", max_model_len=1024, max_num_seqs=B, max_num_batched_tokens=16384, gpu_memory_utilization=0.5,
              load_format="dummy", cudagraph_capture_sizes=[B], async_scheduling=False)
    runner = llm.llm_engine.engine_core.engine_core.model_executor.worker.runner
    prompts = [{"prompt_token_ids": [(7 * i + 3 * j) % 3000 + 5 for j in range(560 + (i % 7) * 8)]} for i in range(B)]
    sp = SamplingParams(max_tokens=48, temperature=0.0, ignore_eos=True)
    llm.generate(prompts, SamplingParams(max_tokens=4, temperature=0.0, ignore_eos=True))      # warm-up
    t0, s0, h0 = runner.prep_host_s, runner._step, runner.h2d_bytes
    d0, n0 = runner.prep_decode_s, runner.prep_decode_n
    llm.generate(prompts, sp)
    torch.cuda.synchronize()
    steps = runner._step - s0
    res = {"batch": B, "gpu_input_prep": bool(runner.gpu_prep), "steps": steps,
           "host_prep_us_per_step": round((runner.prep_host_s - t0) / steps * 1e6, 1),
           "host_prep_us_per_decode_step": round((runner.prep_decode_s - d0) / max(runner.prep_decode_n - n0, 1) * 1e6, 1),
           "h2d_bytes_per_step": int((runner.h2d_bytes - h0) / steps)}
    llm.shutdown()
    return res


if __name__ == "__main__":
    if len(sys.argv) > 1:
        print(json.dumps(run(int(sys.argv[1]))), flush=True)
    else:
        for B in (64, 256, 1024):
            for mode in ("0", "1"):
                env = dict(os.environ, VLLM_B200_GPU_INPUT_PREP=mode, VLLM_LOGGING_LEVEL="WARNING")
                r = subprocess.run([sys.executable, __file__, str(B)], env=env, capture_output=True, text=True, timeout=300)
                line = [l for l in r.stdout.splitlines() if l.startswith("{")]
                print(line[-1] if line else json.dumps({"batch": B, "mode": mode, "error": r.stderr[-400:]}), flush=True)
