"""Failure propagation and teardown (reference: tests/v1/shutdown/test_forward_error.py, test_processor_error.py,
test_delete.py, test_startup_error.py): an error in the forward pass, a worker that dies, an engine core that is
killed, or a model that fails to load must surface as ``EngineDeadError`` / a start-up exception in the front-end —
promptly, for every waiting request — and leave no child process behind."""
import asyncio
import os
import time

import psutil
import pytest

EXT = "tests.worker_ext.FaultInjection"
SP_KW = dict(max_tokens=8, temperature=0.0, ignore_eos=True)


def _children():
    return [p for p in psutil.Process().children(recursive=True) if p.is_running() and p.status() != psutil.STATUS_ZOMBIE]


def _wait_no_new_children(before, timeout=15.0):
    before_pids = {p.pid for p in before}
    t0 = time.time()
    while time.time() - t0 < timeout:
        left = [p for p in _children() if p.pid not in before_pids and "resource_tracker" not in " ".join(p.cmdline())]
        if not left:
            return []
        time.sleep(0.2)
    return left


def _llm(**kw):
    from tests.utils import make_llm
    return make_llm("llama-tiny", worker_extension_cls=EXT, **kw)


def test_worker_extension_and_worker_cls_resolution():
    """--worker-extension-cls mixes methods into the worker (collective_rpc reaches them), clashes are rejected,
    --worker-cls swaps the class."""
    from vllm_b200.config import ParallelConfig
    from vllm_b200.worker import resolve_worker_class
    from vllm_b200.worker.worker import Worker

    class Cfg:
        parallel_config = ParallelConfig()
    assert resolve_worker_class(Cfg) is Worker
    Cfg.parallel_config = ParallelConfig(worker_extension_cls=EXT)
    cls = resolve_worker_class(Cfg)
    assert issubclass(cls, Worker) and hasattr(cls, "arm_failure") and not hasattr(Worker, "arm_failure")
    Cfg.parallel_config = ParallelConfig(worker_extension_cls="tests.worker_ext.Clashing")
    with pytest.raises(ValueError, match="execute_model"):
        resolve_worker_class(Cfg)
    Cfg.parallel_config = ParallelConfig(worker_cls="tests.worker_ext.FaultInjection")
    assert resolve_worker_class(Cfg).__name__ == "FaultInjection"
    with pytest.raises(ValueError):
        Cfg.parallel_config = ParallelConfig(worker_cls="NoModule")
        resolve_worker_class(Cfg)
    llm = _llm()
    info = llm.collective_rpc("whoami")
    assert info[0]["rank"] == 0 and info[0]["cls"] == "WorkerWithFaultInjection"
    llm.shutdown()


def test_forward_error_inprocess_raises_and_later_calls_fail_fast():
    from vllm_b200 import SamplingParams
    llm = _llm()
    assert llm.generate([{"prompt_token_ids": [1, 2, 3]}], SamplingParams(**SP_KW))[0].outputs[0].token_ids
    llm.collective_rpc("arm_failure", kwargs={"after_steps": 2})
    with pytest.raises(Exception, match="injected forward failure"):
        llm.generate([{"prompt_token_ids": [1, 2, 3]}], SamplingParams(**SP_KW))
    llm.shutdown()


@pytest.mark.parametrize("kind", ["raise", "exit"])
def test_forward_error_in_engine_core_process_sync(kind):
    """LLM over a separate engine-core process: the step raises (or the process dies outright) → generate() raises
    EngineDeadError, every later call fails fast, shutdown reaps the child."""
    from vllm_b200 import SamplingParams
    from vllm_b200.engine.exceptions import EngineDeadError
    before = _children()
    llm = _llm(multiprocess_mode=True)
    assert len(llm.generate([{"prompt_token_ids": [1, 2, 3]}], SamplingParams(**SP_KW))[0].outputs[0].token_ids) == 8
    assert llm.collective_rpc("arm_failure", kwargs={"after_steps": 2, "kind": kind}) == [True]
    t0 = time.time()
    with pytest.raises(EngineDeadError):
        llm.generate([{"prompt_token_ids": [4, 5, 6]}, {"prompt_token_ids": [7, 8]}], SamplingParams(**SP_KW))
    assert time.time() - t0 < 30
    t0 = time.time()
    with pytest.raises(EngineDeadError):
        llm.generate([{"prompt_token_ids": [4, 5, 6]}], SamplingParams(**SP_KW))
    assert time.time() - t0 < 10, "a dead engine must fail fast"
    llm.shutdown()
    assert _wait_no_new_children(before) == []


def test_engine_core_killed_async_all_streams_fail():
    """AsyncLLM: SIGKILL of the engine core mid-generation → every in-flight stream raises EngineDeadError, the
    engine reports errored / is_running False, new requests are refused."""
    from vllm_b200 import SamplingParams
    from vllm_b200.engine.arg_utils import EngineArgs
    from vllm_b200.engine.async_llm import AsyncLLM
    from vllm_b200.engine.exceptions import EngineDeadError
    before = _children()

    async def main():
        ea = EngineArgs(model="llama-tiny", device="cpu", dtype="float32", max_model_len=256, max_num_seqs=8,
                        max_num_batched_tokens=128, kv_cache_memory_bytes=8 << 20, disable_log_stats=True)
        eng = AsyncLLM.from_engine_args(ea)
        sp = SamplingParams(max_tokens=200, temperature=0.0, ignore_eos=True)
        started = asyncio.Event()

        async def one(i):
            n = 0
            try:
                async for o in eng.generate({"prompt_token_ids": [3 + i, 9, 27]}, sp, f"r{i}"):
                    n += 1
                    started.set()
            except Exception as e:
                return type(e), n
            return None, n
        tasks = [asyncio.create_task(one(i)) for i in range(4)]
        await asyncio.wait_for(started.wait(), 60)
        eng.engine_core.res.procs[0].kill()
        res = await asyncio.wait_for(asyncio.gather(*tasks), 60)
        errored = eng.errored
        with pytest.raises(EngineDeadError):
            async for _ in eng.generate({"prompt_token_ids": [1, 2]}, sp, "late"):
                pass
        eng.shutdown()
        return res, errored
    res, errored = asyncio.run(main())
    assert errored
    assert all(t is EngineDeadError for t, _ in res), res
    assert _wait_no_new_children(before) == []


def test_worker_process_death_in_mp_executor_propagates():
    """TP=2 multiproc executor: rank 1 exits during a step → the executor notices, the engine raises, shutdown
    terminates the surviving rank."""
    from vllm_b200 import SamplingParams
    before = _children()
    llm = _llm(tensor_parallel_size=2, distributed_executor_backend="mp")
    assert llm.generate([{"prompt_token_ids": [1, 2, 3]}], SamplingParams(**SP_KW))[0].outputs[0].token_ids
    armed = llm.collective_rpc("arm_failure", kwargs={"after_steps": 1, "kind": "exit", "only_rank": 1})
    assert armed == [False, True]
    t0 = time.time()
    with pytest.raises(Exception):
        llm.generate([{"prompt_token_ids": [4, 5, 6]}], SamplingParams(**SP_KW))
    assert time.time() - t0 < 60
    llm.shutdown()
    assert _wait_no_new_children(before) == []


def test_startup_error_in_engine_core_process_surfaces():
    """A model that cannot be built fails LLM(...) with an exception (no hang on the ready handshake) and leaves no
    child behind."""
    from vllm_b200 import LLM
    before = _children()
    t0 = time.time()
    with pytest.raises(Exception):
        LLM("llama-tiny", device="cpu", dtype="float32", max_model_len=64, kv_cache_memory_bytes=8 << 20,
            multiprocess_mode=True, worker_cls="tests.worker_ext.DoesNotExist")
    assert time.time() - t0 < 120
    assert _wait_no_new_children(before) == []


def test_delete_without_shutdown_reaps_children():
    """Dropping the last reference (no explicit shutdown) still tears the engine-core process down (weakref finalizer)."""
    import gc
    from vllm_b200 import SamplingParams
    before = _children()
    llm = _llm(multiprocess_mode=True)
    llm.generate([{"prompt_token_ids": [1, 2, 3]}], SamplingParams(**SP_KW))
    del llm
    gc.collect()
    assert _wait_no_new_children(before) == []
