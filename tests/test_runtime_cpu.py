"""Host-side runtime pieces: shm ring + native spin-wait, GIL-free file IO, multiprocess executor (gloo)."""
import multiprocessing as mp
import os

import pytest


def _ring_reader(handle, rank, q):
    from vllm_b200.parallel.shm_broadcast import ShmRing
    ring = ShmRing.attach(handle, rank)
    tot = 0
    for _ in range(200):
        m = ring.get(timeout=20)
        tot += m["i"] + len(m["pad"])
    q.put(tot)
    ring.close()


def test_shm_ring_broadcast_chunked():
    from vllm_b200.native import lib
    from vllm_b200.parallel.shm_broadcast import ShmRing
    assert lib() is not None, "native runtime library failed to build"
    ctx = mp.get_context("spawn")
    ring = ShmRing(2, slot_bytes=4096, n_slots=4)
    q = ctx.Queue()
    ps = [ctx.Process(target=_ring_reader, args=(ring.handle(), r, q)) for r in range(2)]
    [p.start() for p in ps]
    exp = 0
    for i in range(200):
        pad = b"x" * (i * 97 % 9000)          # messages up to 2+ slots → chunking
        ring.put({"i": i, "pad": pad}, timeout=20)
        exp += i + len(pad)
    res = [q.get(timeout=30) for _ in ps]
    [p.join(10) for p in ps]
    ring.close()
    assert res == [exp, exp]


def test_native_file_io(tmp_path):
    import ctypes
    from vllm_b200.native import lib
    L = lib()
    data = os.urandom(1 << 20)
    p = str(tmp_path / "blk.bin").encode()
    assert L.vbn_write_file(p, data, len(data), 0, 1) == len(data)
    assert L.vbn_file_size(p) == len(data)
    buf = ctypes.create_string_buffer(len(data))
    assert L.vbn_read_file(p, buf, len(data), 0, 0) == len(data)
    assert buf.raw == data
    assert L.vbn_read_file(str(tmp_path / "missing").encode(), buf, 16, 0, 0) < 0


def test_multiproc_executor_tp2_matches_tp1():
    """2 worker processes (gloo) driven through the shm ring produce the TP=1 tokens."""
    from tests.utils import llama_full_weights, make_llm
    from vllm_b200 import SamplingParams
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config("llama-tiny")
    prompts = [{"prompt_token_ids": [5, 17, 99, 3, 250, 41]}, {"prompt_token_ids": list(range(10, 40))}]
    sp = SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True)
    llm = make_llm("llama-tiny", weights=llama_full_weights(cfg, seed=1))
    ref = [o.outputs[0].token_ids for o in llm.generate(prompts, sp)]
    llm.shutdown()
    llm = make_llm("llama-tiny", tensor_parallel_size=2, distributed_executor_backend="mp")
    ex = llm.llm_engine.engine_core.engine_core.model_executor
    ex.collective_rpc("load_model", kwargs={"weights": list(llama_full_weights(cfg, seed=1))})
    got = [o.outputs[0].token_ids for o in llm.generate(prompts, sp)]
    llm.shutdown()
    assert got == ref


def test_dp_coordinator_agreement():
    import threading
    from vllm_b200.engine.dp import DPCoordinator
    from vllm_b200.utils import get_open_port
    want = {r: [True, True, False, True, False] for r in range(3)}
    for attempt in range(3):                       # a stolen rendezvous port shows up as missing ranks: retry on a new one
        port = get_open_port()
        res = {}

        def run(rank):
            c = DPCoordinator("127.0.0.1", port, rank, 3)
            res[rank] = [c.any_unfinished((rank == 1 and step < 2) or (rank == 2 and step == 3)) for step in range(5)]

        ts = [threading.Thread(target=run, args=(r,), daemon=True) for r in range(3)]
        [t.start() for t in ts]
        [t.join(60) for t in ts]
        if len(res) == 3:
            break
    assert res == want


def test_multiproc_executor_pp2_matches_single():
    """Pipeline parallelism: 2 stages in 2 worker processes, hidden states travel over the PP group."""
    from tests.utils import llama_full_weights, make_llm
    from vllm_b200 import SamplingParams
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config("llama-tiny")
    prompts = [{"prompt_token_ids": [5, 17, 99, 3, 250, 41]}, {"prompt_token_ids": list(range(10, 40))}]
    sp = SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True)
    llm = make_llm("llama-tiny", weights=llama_full_weights(cfg, seed=1))
    ref = [o.outputs[0].token_ids for o in llm.generate(prompts, sp)]
    llm.shutdown()
    llm = make_llm("llama-tiny", pipeline_parallel_size=2, distributed_executor_backend="mp")
    ex = llm.llm_engine.engine_core.engine_core.model_executor
    ex.collective_rpc("load_model", kwargs={"weights": list(llama_full_weights(cfg, seed=1))})
    got = [o.outputs[0].token_ids for o in llm.generate(prompts, sp)]
    llm.shutdown()
    assert got == ref


def test_cli_bench_and_platform(capsys):
    import json
    from vllm_b200.entrypoints.cli import main
    from vllm_b200 import platforms
    assert platforms.current_platform.device_type in ("cpu", "cuda")
    assert "vllm_b200" in platforms.collect_env()
    main(["bench", "throughput", "--model", "llama-tiny", "--device", "cpu", "--dtype", "float32", "--num-prompts", "4",
          "--input-len", "16", "--output-len", "4", "--max-model-len", "128", "--kv-cache-memory-bytes", str(8 << 20)])
    out = capsys.readouterr().out
    res = json.loads(out[out.index("{"):])
    assert res["num_requests"] == 4 and res["total_output_tokens"] == 16


@pytest.mark.parametrize("dcp", [1, 2])
def test_deepseek_tp2_with_decode_context_parallel(dcp):
    """MLA + MoE model on 2 worker processes; with dcp=2 the latent cache is sharded by position across the two
    ranks (each stores every other token) and partial attention is merged by log-sum-exp — same tokens as TP=1."""
    import torch
    from tests.utils import make_hf_model, make_llm
    from vllm_b200 import SamplingParams
    hf = make_hf_model("deepseek-tiny")
    w = [(k, v.clone()) for k, v in hf.state_dict().items()]
    prompts = [{"prompt_token_ids": [5, 17, 99, 3, 250, 41, 8]}, {"prompt_token_ids": list(range(10, 47))}]
    sp = SamplingParams(max_tokens=10, temperature=0.0, ignore_eos=True)
    llm = make_llm("deepseek-tiny", weights=w)
    ref = [o.outputs[0].token_ids for o in llm.generate(prompts, sp)]
    llm.shutdown()
    # chunked prefill (16-token budget) exercises "previous context on other ranks + current chunk local"
    llm = make_llm("deepseek-tiny", tensor_parallel_size=2, decode_context_parallel_size=dcp,
                   distributed_executor_backend="mp", block_size=16, max_num_batched_tokens=16)
    ex = llm.llm_engine.engine_core.engine_core.model_executor
    ex.collective_rpc("load_model", kwargs={"weights": w})
    got = [o.outputs[0].token_ids for o in llm.generate(prompts, sp)]
    if dcp > 1:
        specs = ex.collective_rpc("get_kv_cache_specs")
        s0 = next(iter(specs[0].values()))
        assert s0.dcp == 2 and s0.block_size == 32 and s0.phys_block_size == 16
    llm.shutdown()
    assert got == ref


def test_gqa_decode_context_parallel_on_replicated_kv_heads():
    """DCP for an ordinary (non-MLA) cache: an MQA model at TP=2 replicates its single KV head on both ranks; with dcp=2
    the two replicas shard the cache by position instead (half the slots per block on each rank), gather each other's
    query heads, and merge partial attention by log-sum-exp — same tokens as TP=1, chunked prefill included."""
    from tests.utils import llama_full_weights, make_llm
    from vllm_b200 import SamplingParams
    from vllm_b200.config import get_preset_hf_config
    ov = {"num_key_value_heads": 1}
    cfg = dict(get_preset_hf_config("llama-tiny"), **ov)
    w = list(llama_full_weights(cfg, seed=4, scale=0.08))
    prompts = [{"prompt_token_ids": [5, 17, 99, 3, 250, 41, 8]}, {"prompt_token_ids": list(range(10, 47))}]
    sp = SamplingParams(max_tokens=10, temperature=0.0, ignore_eos=True)
    llm = make_llm("llama-tiny", weights=w, hf_overrides=ov)
    ref = [o.outputs[0].token_ids for o in llm.generate(prompts, sp)]
    llm.shutdown()
    llm = make_llm("llama-tiny", hf_overrides=ov, tensor_parallel_size=2, decode_context_parallel_size=2,
                   distributed_executor_backend="mp", block_size=16, max_num_batched_tokens=16)
    ex = llm.llm_engine.engine_core.engine_core.model_executor
    ex.collective_rpc("load_model", kwargs={"weights": w})
    got = [o.outputs[0].token_ids for o in llm.generate(prompts, sp)]
    specs = ex.collective_rpc("get_kv_cache_specs")
    s0 = next(iter(specs[0].values()))
    llm.shutdown()
    assert s0.dcp == 2 and s0.block_size == 32 and s0.phys_block_size == 16
    assert got == ref


def test_mixtral_expert_parallel_tp2():
    """--enable-expert-parallel: each of the 2 ranks owns 2 whole experts (expert map), combine = all-reduce."""
    from tests.utils import make_hf_model, make_llm
    from vllm_b200 import SamplingParams
    hf = make_hf_model("mixtral-tiny")
    w = [(k, v.clone()) for k, v in hf.state_dict().items()]
    prompts = [{"prompt_token_ids": [5, 17, 99, 3, 250, 41, 8]}, {"prompt_token_ids": list(range(10, 40))}]
    sp = SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True)
    llm = make_llm("mixtral-tiny", weights=w)
    ref = [o.outputs[0].token_ids for o in llm.generate(prompts, sp)]
    llm.shutdown()
    llm = make_llm("mixtral-tiny", tensor_parallel_size=2, enable_expert_parallel=True, distributed_executor_backend="mp")
    ex = llm.llm_engine.engine_core.engine_core.model_executor
    ex.collective_rpc("load_model", kwargs={"weights": w})

    def local_experts(worker):
        return [m.local_experts for m in worker.runner.model.modules() if hasattr(m, "local_experts")][0]
    assert ex.collective_rpc(local_experts) == [2, 2]
    got = [o.outputs[0].token_ids for o in llm.generate(prompts, sp)]
    llm.shutdown()
    assert got == ref


def test_data_parallel_engines_load_balanced():
    """2 engine-core processes behind the load-balancing async client: all requests finish, both engines serve."""
    import asyncio
    from vllm_b200 import SamplingParams
    from vllm_b200.engine.arg_utils import EngineArgs
    from vllm_b200.engine.async_llm import AsyncLLM

    async def main():
        ea = EngineArgs(model="llama-tiny", device="cpu", dtype="float32", max_model_len=128, max_num_seqs=8,
                        max_num_batched_tokens=128, kv_cache_memory_bytes=8 << 20, data_parallel_size=2,
                        disable_log_stats=True)
        eng = AsyncLLM.from_engine_args(ea)
        sp = SamplingParams(max_tokens=6, temperature=0.0, ignore_eos=True)
        used = set()

        async def one(i):
            last = None
            async for o in eng.generate({"prompt_token_ids": [3 + i, 9, 27, 81]}, sp, f"r{i}"):
                last = o
            return last
        tasks = [asyncio.create_task(one(i)) for i in range(8)]
        await asyncio.sleep(0)
        outs = await asyncio.gather(*tasks)
        used = set(eng.engine_core.engines_seen) if hasattr(eng.engine_core, "engines_seen") else set()
        eng.shutdown()
        return outs, used
    outs, used = asyncio.run(main())
    assert all(o is not None and len(o.outputs[0].token_ids) == 6 for o in outs)
    # same prompt → same tokens regardless of which replica served it
    assert outs[0].outputs[0].token_ids == outs[0].outputs[0].token_ids
    assert used == {0, 1}


def test_tp2_pp2_four_processes():
    """TP=2 x PP=2 on four worker processes (gloo): stage boundaries carry TP-sharded intermediates."""
    from tests.utils import llama_full_weights, make_llm
    from vllm_b200 import SamplingParams
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config("llama-tiny")
    prompts = [{"prompt_token_ids": [5, 17, 99, 3, 250, 41]}, {"prompt_token_ids": list(range(10, 40))}]
    sp = SamplingParams(max_tokens=6, temperature=0.0, ignore_eos=True)
    llm = make_llm("llama-tiny", weights=llama_full_weights(cfg, seed=1))
    ref = [o.outputs[0].token_ids for o in llm.generate(prompts, sp)]
    llm.shutdown()
    llm = make_llm("llama-tiny", tensor_parallel_size=2, pipeline_parallel_size=2, distributed_executor_backend="mp")
    ex = llm.llm_engine.engine_core.engine_core.model_executor
    ex.collective_rpc("load_model", kwargs={"weights": list(llama_full_weights(cfg, seed=1))})
    got = [o.outputs[0].token_ids for o in llm.generate(prompts, sp)]
    llm.shutdown()
    assert got == ref


def test_rlhf_weight_update_in_place():
    """collective_rpc('update_weights'): new tensors (HF names) replace the served weights without a restart."""
    from tests.utils import llama_full_weights, make_llm
    from vllm_b200 import SamplingParams
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config("llama-tiny")
    sp = SamplingParams(max_tokens=6, temperature=0.0, ignore_eos=True)
    p = [{"prompt_token_ids": [5, 17, 99, 3, 250, 41]}]
    llm = make_llm("llama-tiny", weights=llama_full_weights(cfg, seed=1), enable_prefix_caching=False)
    a = llm.generate(p, sp)[0].outputs[0].token_ids
    n = llm.collective_rpc("update_weights", args=(list(llama_full_weights(cfg, seed=2)),))[0]
    assert n > 0
    b = llm.generate(p, sp)[0].outputs[0].token_ids
    llm.shutdown()
    llm = make_llm("llama-tiny", weights=llama_full_weights(cfg, seed=2))
    c = llm.generate(p, sp)[0].outputs[0].token_ids
    llm.shutdown()
    assert b == c and a != b


def test_elastic_scale_up_and_down():
    """Start with one engine replica, grow to two under load-balancing, shrink back — requests keep being served."""
    import asyncio
    from vllm_b200 import SamplingParams
    from vllm_b200.engine.arg_utils import EngineArgs
    from vllm_b200.engine.async_llm import AsyncLLM

    async def main():
        ea = EngineArgs(model="llama-tiny", device="cpu", dtype="float32", max_model_len=128, max_num_seqs=8,
                        max_num_batched_tokens=128, kv_cache_memory_bytes=8 << 20, enable_elastic_ep=True,
                        disable_log_stats=True)
        eng = AsyncLLM.from_engine_args(ea)
        sp = SamplingParams(max_tokens=5, temperature=0.0, ignore_eos=True)

        async def one(tag, i):
            last = None
            async for o in eng.generate({"prompt_token_ids": [3 + i, 9, 27, 81]}, sp, f"{tag}{i}"):
                last = o
            return last.outputs[0].token_ids

        base = await asyncio.gather(*[one("a", i) for i in range(3)])
        assert await eng.scale_elastic_ep(2) == 2
        eng.engine_core.engines_seen.clear()
        grown = await asyncio.gather(*[one("b", i) for i in range(6)])
        seen = set(eng.engine_core.engines_seen)
        assert await eng.scale_elastic_ep(1) == 1
        shrunk = await asyncio.gather(*[one("c", i) for i in range(3)])
        eng.shutdown()
        return base, grown, seen, shrunk
    base, grown, seen, shrunk = asyncio.run(main())
    assert seen == {0, 1}
    assert grown[:3] == base and shrunk == base


@pytest.mark.parametrize("backend,tp,dbo", [("allgather_reducescatter", 1, False), ("a2a", 1, False), ("a2a", 2, False),
                                            ("a2a", 1, True)])
def test_wide_expert_parallel_across_dp_replicas(backend, tp, dbo):
    """DP=2 + --enable-expert-parallel: the experts are sharded over the two replicas (EP group = DP x TP), tokens are
    dispatched / combined across replicas, idle replicas run dummy batches in lock-step; outputs equal a single engine."""
    import asyncio
    from tests.utils import make_hf_model, make_llm
    from vllm_b200 import SamplingParams
    from vllm_b200.engine.arg_utils import EngineArgs
    from vllm_b200.engine.async_llm import AsyncLLM
    from vllm_b200.utils import get_open_port
    hf = make_hf_model("mixtral-tiny")
    w = [(k, v.clone()) for k, v in hf.state_dict().items()]
    prompts = [[5, 17, 99, 3, 250, 41, 8], list(range(10, 40)), [7, 7, 7, 9], [100, 200, 300, 400, 33]]
    sp = SamplingParams(max_tokens=6, temperature=0.0, ignore_eos=True)
    llm = make_llm("mixtral-tiny", weights=w)
    ref = [o.outputs[0].token_ids for o in llm.generate([{"prompt_token_ids": p} for p in prompts], sp)]
    llm.shutdown()

    async def main():
        ea = EngineArgs(model="mixtral-tiny", device="cpu", dtype="float32", max_model_len=128, max_num_seqs=8,
                        max_num_batched_tokens=128, kv_cache_memory_bytes=8 << 20, data_parallel_size=2,
                        enable_expert_parallel=True, all2all_backend=backend, data_parallel_rpc_port=get_open_port(),
                        tensor_parallel_size=tp, distributed_executor_backend="mp" if tp > 1 else None,
                        disable_log_stats=True, enable_dbo=dbo)
        eng = AsyncLLM.from_engine_args(ea)
        await eng.collective_rpc("load_model", kwargs={"weights": w})
        shards = [d["local_experts"] for d in await eng.collective_rpc("describe_model")]

        async def one(i, p):
            last = None
            async for o in eng.generate({"prompt_token_ids": p}, sp, f"r{i}"):
                last = o
            return last.outputs[0].token_ids
        # first a lone request (one replica works, the other runs dummy batches), then a burst on both replicas
        first = await one(99, prompts[0])
        outs = await asyncio.gather(*[one(i, p) for i, p in enumerate(prompts)])
        used = set(eng.engine_core.engines_seen)
        eng.shutdown()
        return first, outs, used, shards
    first, outs, used, shards = asyncio.run(main())
    assert shards == [4 // (2 * tp)] * tp     # 4 experts over EP = DP x TP ranks → whole experts per rank
    assert first == ref[0] and outs == ref
    assert used == {0, 1}


def test_pcp_plan_balances_causal_work():
    import numpy as np
    from vllm_b200.worker.model_runner import ModelRunner
    ns = np.array([64, 1, 1, 7, 1])
    cu = np.concatenate([[0], np.cumsum(ns)])
    pieces = ModelRunner.pcp_plan(cu, ns, 2)
    # every token exactly once
    seen = sorted((cu[i] + t) for pc in pieces for i, a, b in pc for t in range(a, b))
    assert seen == list(range(int(cu[-1])))
    # the long prefill: rank 0 gets chunks 0 and 3, rank 1 chunks 1 and 2 → equal causal cost (sum of positions)
    cost = [sum(t for i, a, b in pc if i == 0 for t in range(a, b)) for pc in pieces]
    assert abs(cost[0] - cost[1]) <= 64
    assert [sum(b - a for i, a, b in pc if i == 0) for pc in pieces] == [32, 32]


@pytest.mark.parametrize("tp,pcp", [(1, 2), (2, 2)])
def test_prefill_context_parallel_matches_single(tp, pcp):
    """PCP ranks each run the per-token layers on a share of the step's tokens, K/V are all-gathered per layer, the
    sampled tokens are identical to a single worker (long prefill split head+tail, decodes round-robin)."""
    from tests.utils import llama_full_weights, make_llm
    from vllm_b200 import SamplingParams
    from vllm_b200.config import get_preset_hf_config
    cfg = get_preset_hf_config("llama-tiny")
    w = list(llama_full_weights(cfg, seed=11))
    prompts = [{"prompt_token_ids": [(7 * i) % 400 + 3 for i in range(90)]}, {"prompt_token_ids": [5, 17, 99]},
               {"prompt_token_ids": list(range(10, 40))}]
    sp = SamplingParams(max_tokens=6, temperature=0.0, ignore_eos=True)
    llm = make_llm("llama-tiny", weights=w)
    ref = [o.outputs[0].token_ids for o in llm.generate(prompts, sp)]
    llm.shutdown()
    llm = make_llm("llama-tiny", tensor_parallel_size=tp, prefill_context_parallel_size=pcp, distributed_executor_backend="mp",
                   max_num_batched_tokens=64)          # 90-token prompt → chunked prefill across steps as well
    llm.llm_engine.engine_core.engine_core.model_executor.collective_rpc("load_model", kwargs={"weights": w})
    got = [o.outputs[0].token_ids for o in llm.generate(prompts, sp)]
    one = llm.generate(prompts[:1], sp)[0].outputs[0].token_ids          # a lone request: one rank idles on decode steps
    llm.shutdown()
    assert got == ref and one == ref[0]


def test_eplb_redundant_experts_and_runtime_rebalance():
    """EP=2 with 2 redundant expert slots: tokens are routed to replicas, a rebalance re-places the physical experts
    by observed load and ships the weights between the ranks — outputs stay those of the plain model throughout."""
    from tests.utils import make_hf_model, make_llm
    from vllm_b200 import SamplingParams
    hf = make_hf_model("mixtral-tiny")
    w = [(k, v.clone()) for k, v in hf.state_dict().items()]
    prompts = [{"prompt_token_ids": [5, 17, 99, 3, 250, 41, 8]}, {"prompt_token_ids": list(range(10, 40))}]
    sp = SamplingParams(max_tokens=8, temperature=0.0, ignore_eos=True)
    llm = make_llm("mixtral-tiny", weights=w)
    ref = [o.outputs[0].token_ids for o in llm.generate(prompts, sp)]
    llm.shutdown()
    llm = make_llm("mixtral-tiny", tensor_parallel_size=2, enable_expert_parallel=True, enable_eplb=True,
                   num_redundant_experts=2, eplb_step_interval=5, distributed_executor_backend="mp")
    ex = llm.llm_engine.engine_core.engine_core.model_executor
    ex.collective_rpc("load_model", kwargs={"weights": w})
    assert [d["local_experts"] for d in ex.collective_rpc("describe_model")] == [3, 3]      # (4 + 2) slots over 2 ranks
    got = [o.outputs[0].token_ids for o in llm.generate(prompts, sp)]       # > 5 steps: the periodic rebalance ran
    assert got == ref
    info = ex.collective_rpc("eplb_rebalance")                              # and an explicit one on top
    assert info[0]["layers"] == 2 and sum(info[0]["replicas"]) == 6 and max(info[0]["replicas"]) >= 2
    again = [o.outputs[0].token_ids for o in llm.generate(prompts, sp)]
    llm.shutdown()
    assert again == ref


def test_ray_executor_with_stub_cluster():
    """The ray executor drives Worker actors through placement-group scheduling; with an in-process stand-in for the
    ray API the whole engine runs through it and matches the in-process executor."""
    from tests import fake_ray
    from tests.utils import make_llm
    from vllm_b200 import SamplingParams
    prompts = [{"prompt_token_ids": [5, 17, 99, 3, 250, 41, 8]}, {"prompt_token_ids": list(range(10, 30))}]
    sp = SamplingParams(max_tokens=6, temperature=0.0, ignore_eos=True)
    llm = make_llm("llama-tiny")
    ref = [o.outputs[0].token_ids for o in llm.generate(prompts, sp)]
    llm.shutdown()
    names = fake_ray.install()
    try:
        llm = make_llm("llama-tiny", distributed_executor_backend="ray")
        ex = llm.llm_engine.engine_core.engine_core.model_executor
        assert type(ex).__name__ == "RayExecutor" and ex.node_ips == ["127.0.0.1"]
        pg = fake_ray.STATE["pgs"][0]
        assert pg.strategy == "PACK" and pg.bundles == [{"CPU": 1.0}]
        assert fake_ray.STATE["actors"][0][1]["scheduling_strategy"].index == 0
        got = [o.outputs[0].token_ids for o in llm.generate(prompts, sp)]
        assert ex.collective_rpc("check_health") == [True]
        llm.shutdown()
        assert got == ref and fake_ray.STATE["killed"] == 1
    finally:
        fake_ray.uninstall(names)
    # without ray installed the backend fails with an actionable message
    with pytest.raises(ImportError, match="ray executor needs"):
        make_llm("llama-tiny", distributed_executor_backend="ray")


def test_out_of_band_tensor_frames_roundtrip_and_alias():
    """engine/serial.py multi-frame codec: large tensors / arrays leave the msgpack body and travel as their own frames that
    ALIAS the source storage (zero copy on the sending side); small ones stay inline; typed decoding of engine messages."""
    import numpy as np
    import torch
    from vllm_b200.core.request import EngineCoreRequest
    from vllm_b200.engine.serial import (OOBDecoder, OOBEncoder, decode_request_frames, encode_request_frames)
    from vllm_b200.sampling_params import SamplingParams
    enc, dec = OOBEncoder(), OOBDecoder()
    big, small = torch.randn(64, 128).bfloat16(), torch.arange(5)
    msg = {"big": big, "small": small, "arr": np.arange(1000, dtype=np.float32).reshape(10, 100), "nest": [1, {"t": torch.randn(300)}]}
    frames = enc.encode(msg)
    assert len(frames) == 4 and len(frames[1]) == big.numel() * 2                # body + three out-of-band buffers
    big.view(torch.int16)[0, 0] = 1234                                          # the frame is a view of the tensor
    assert bytes(frames[1][:2]) == (1234).to_bytes(2, "little")
    out = dec.decode([bytes(f) for f in frames])
    assert torch.equal(out["big"], big) and torch.equal(out["small"], small) and np.array_equal(out["arr"], msg["arr"])
    assert torch.equal(out["nest"][1]["t"], msg["nest"][1]["t"])
    r = EngineCoreRequest(request_id="r1", prompt_token_ids=[1, 2, 3], sampling_params=SamplingParams(max_tokens=3),
                          prompt_embeds=torch.randn(3, 64))
    fr = encode_request_frames(r)
    assert len(fr) == 2
    back = decode_request_frames([bytes(f) for f in fr])
    assert back.request_id == "r1" and torch.equal(back.prompt_embeds, r.prompt_embeds)
