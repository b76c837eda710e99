"""Worker: one process (or in-process object) per device.

Behavioural reference: vllm/v1/worker/gpu_worker.py (``init_device`` :316,
``load_model`` :456, ``determine_available_memory`` :481-634,
``compile_or_warm_up_model`` :705, ``execute_model`` :1064 with PP send/recv
:1112-1151, sleep/wake :204-260, profiler RPC :1157).
"""
from __future__ import annotations

import gc
import os
import time
from typing import Any, Optional

import torch

from vllm_b200.config import VllmConfig
from vllm_b200.core.kv_cache_manager import KVCacheConfig, KVCacheSpec
from vllm_b200.core.sched_output import ModelRunnerOutput, SchedulerOutput
from vllm_b200.logger import init_logger, set_process_prefix
from vllm_b200.parallel import state as pstate
from vllm_b200.worker.model_runner import ModelRunner

logger = init_logger(__name__)


class Worker:
    def __init__(self, vllm_config: VllmConfig, local_rank: int = 0, rank: int = 0,
                 distributed_init_method: Optional[str] = None, is_driver_worker: bool = True) -> None:
        self.vllm_config = vllm_config
        self.local_rank, self.rank = local_rank, rank
        self.init_method = distributed_init_method
        self.is_driver_worker = is_driver_worker
        self.device: Optional[torch.device] = None
        self.runner: Optional[ModelRunner] = None
        self.profiler = None
        self._sleep_saved: dict[str, torch.Tensor] = {}
        self._sleeping = False
        self.init_free_bytes = 0

    # ----------------------------------------------------------------- init
    def init_device(self) -> None:
        pc = self.vllm_config.parallel_config
        from vllm_b200.utils.batch_invariant import apply_env_overrides
        if apply_env_overrides():
            logger.info("batch-invariant mode: fixed-order reductions (no split-K / split-KV / in-switch reduce)")
        dev_type = self.vllm_config.device_config.resolve()
        if dev_type == "cuda":
            # DP rank d owns GPUs [d*world, (d+1)*world) of the node
            ndev = torch.cuda.device_count()
            self.local_rank = (pc.data_parallel_rank % max(pc.data_parallel_size_local or 1, 1) * pc.world_size
                               + self.local_rank) % max(ndev, 1)
            torch.cuda.set_device(self.local_rank)
            self.device = torch.device("cuda", self.local_rank)
            from vllm_b200.utils.numa import bind_worker_to_gpu_node
            cpus = bind_worker_to_gpu_node(self.local_rank)          # VLLM_B200_NUMA_BIND=1: CPUs of the GPU's NUMA node
            if cpus:
                logger.info("worker %d pinned to %d CPUs of its GPU's NUMA node", self.rank, len(cpus))
            gc.collect()
            torch.cuda.empty_cache()
            self.init_free_bytes, self.total_bytes = torch.cuda.mem_get_info(self.device)
            # a previous process on this GPU may still be tearing down (its memory is released
            # asynchronously): give it a moment before sizing the KV cache against free memory
            want = self.vllm_config.cache_config.gpu_memory_utilization * self.total_bytes
            t_end = time.time() + 20.0
            while self.init_free_bytes < min(want, 0.9 * self.total_bytes) and time.time() < t_end:
                time.sleep(0.5)
                self.init_free_bytes, _ = torch.cuda.mem_get_info(self.device)
        else:
            self.device = torch.device("cpu")
        world = pc.world_size
        if world > 1:
            set_process_prefix(f"(Worker rank={self.rank}) ")
        wide_ep = (pc.enable_expert_parallel and pc.data_parallel_size > 1 and self.vllm_config.model_config.is_moe)
        if wide_ep:
            # experts are sharded over DP x TP: one process group across every replica's workers
            # (rank = dp_rank * replica_world + local rank; rendezvous next to the DP coordinator's store)
            pstate.init_distributed_environment(
                world * pc.data_parallel_size, pc.data_parallel_rank * world + self.rank, self.local_rank,
                f"tcp://{pc.data_parallel_master_ip}:{pc.data_parallel_master_port + 1}")
        else:
            pstate.init_distributed_environment(world, self.rank, self.local_rank, self.init_method)
        pstate.initialize_model_parallel(pc.tensor_parallel_size, pc.pipeline_parallel_size,
                                         pc.data_parallel_size if wide_ep else 1,
                                         pc.decode_context_parallel_size, pc.prefill_context_parallel_size,
                                         enable_device_comm=not pc.disable_custom_all_reduce)
        torch.manual_seed(self.vllm_config.model_config.seed)
        self.runner = ModelRunner(self.vllm_config, self.device)
        self.runner.eplb_hook = self.eplb_rebalance

    def load_model(self, weights=None) -> None:
        self.runner.load_model(weights)

    def get_kv_cache_specs(self) -> dict[str, KVCacheSpec]:
        return self.runner.get_kv_cache_specs()

    def determine_available_memory(self) -> int:
        """Bytes available for the KV cache = util * total − (weights + peak activations + other)."""
        cc = self.vllm_config.cache_config
        if cc.kv_cache_memory_bytes is not None:
            return cc.kv_cache_memory_bytes
        if self.device.type != "cuda":
            return int(cc.swap_space_gb * (1 << 30)) or (1 << 30)
        torch.cuda.reset_peak_memory_stats(self.device)
        self.runner.profile_run()
        torch.cuda.synchronize()
        peak_torch = torch.cuda.max_memory_allocated(self.device)
        free, total = torch.cuda.mem_get_info(self.device)
        # memory this process holds outside the torch allocator (CUDA context, NCCL, symmetric heap);
        # measured against the free memory seen at start-up so other tenants are not billed to us
        non_torch = max((self.init_free_bytes - free) - torch.cuda.memory_reserved(self.device), 0)
        budget = min(int(total * cc.gpu_memory_utilization), self.init_free_bytes)
        graph_reserve = 0 if self.vllm_config.model_config.enforce_eager else (2 << 30)
        avail = budget - peak_torch - non_torch - graph_reserve
        gc.collect()
        torch.cuda.empty_cache()
        logger.info("memory: total %.1f GiB, budget %.1f, torch peak %.1f, non-torch %.1f → KV %.1f GiB",
                    total / 2**30, budget / 2**30, peak_torch / 2**30, non_torch / 2**30, avail / 2**30)
        return max(avail, 0)

    def initialize_from_config(self, kv_cache_config) -> None:
        if isinstance(kv_cache_config, (list, tuple)):      # one entry per worker rank
            kv_cache_config = kv_cache_config[self.rank]
        self.runner.initialize_kv_cache(kv_cache_config)

    def compile_or_warm_up_model(self) -> None:
        if self.device.type == "cuda":
            self.runner.capture_model()
        from vllm_b200.utils import gpu_sync_debug
        gpu_sync_debug.arm()      # VLLM_GPU_SYNC_CHECK: from here on the step may not block on the device

    # -------------------------------------------------------------- execute
    @torch.inference_mode()
    def execute_model(self, so: SchedulerOutput, non_block: bool = False):
        pp = pstate.get_pp_group()
        intermediate = None
        if pp.world_size > 1 and not pp.is_first_rank:
            if so.total_num_scheduled_tokens > 0:
                intermediate = pp.recv_tensor_dict(device=self.device)
        from vllm_b200.utils.gpu_sync_debug import sync_guard
        with sync_guard():
            out = self.runner.execute_model(so, intermediate, non_block=non_block)
        if pp.world_size > 1 and not pp.is_last_rank:
            if isinstance(getattr(out, "kv_connector_output", None), dict):
                pp.send_tensor_dict(out.kv_connector_output)
            return None
        return out

    # -------------------------------------------------------- sleep / wake
    @torch.inference_mode()
    def eplb_rebalance(self) -> dict:
        """Re-place the physical experts of every MoE layer from the routed-token counts since the last rebalance and
        move the weights (reference: eplb_state.py ``rearrange`` + rebalance_execute.py).  Collective over the EP group."""
        from vllm_b200.parallel.eplb import rebalance_experts
        layers = [m for m in self.runner.model.modules() if getattr(m, "eplb", False)]
        if not layers:
            return {"layers": 0, "moved": 0}
        g = pstate.get_tp_group()
        load = torch.stack([l.eplb_load for l in layers]).cpu()
        pl = rebalance_experts(load, layers[0].num_physical, layers[0].ep_size)
        moved = sum(l.eplb_apply(pl.phy2log[i], g) for i, l in enumerate(layers))
        return {"layers": len(layers), "moved": moved, "replicas": pl.logcnt[0].tolist(), "rank": self.rank}

    def describe_model(self) -> dict:
        """Shape of this worker's shard (debug / tests / ``collective_rpc`` from the API)."""
        m = self.runner.model
        experts = [int(x.local_experts) for x in m.modules() if hasattr(x, "local_experts")]
        return {"rank": self.rank, "param_bytes": sum(p.numel() * p.element_size() for p in m.parameters()),
                "local_experts": experts[0] if experts else 0, "class": type(m).__name__}

    def execute_dummy_batch(self) -> None:
        """One token through the model with no KV cache: keeps this rank's collectives (EP all-to-all, TP
        all-reduce) in step with peers that have real work (reference: gpu_worker.execute_dummy_batch)."""
        r = self.runner
        r.input_ids[:1].zero_()
        r.positions[:1].zero_()
        r._forward(1, {})

    def reset_after_fault(self) -> bool:
        """Fault-tolerance retry: forget every in-flight request and pending async step, keep weights and caches
        (reference: v1/worker/sentinel/gpu_worker_sentinel.py:30-89)."""
        if self.device.type == "cuda":
            torch.cuda.synchronize()
        r = self.runner
        r._pending = None
        for rid in list(r.req_index):
            r._remove(rid)
        r._parked.clear()
        return True

    # ---------------------------------------------------------------- process checkpoint / restore
    def checkpoint_prepare(self) -> None:
        """Before a CRIU / cuda-checkpoint style process snapshot: drop everything that cannot be snapshotted — the
        NVSwitch symmetric heaps (peer mappings, multicast objects) and captured CUDA graphs that reference them
        (reference: gpu_worker.py:262-266 → parallel_state.checkpoint_prepare_distributed_state)."""
        if self.device.type == "cuda":
            torch.cuda.synchronize()
        self.runner.graphs.clear()
        for g in (pstate.get_tp_group(), pstate.get_ep_group()):
            if g.device_comm is not None:
                g.device_comm.close()
                g.device_comm = None
        self._checkpointed = True

    def checkpoint_restore(self) -> None:
        """After restore: rebuild the symmetric heaps and re-capture the decode graphs."""
        if not getattr(self, "_checkpointed", False):
            return
        pc = self.vllm_config.parallel_config
        if self.device.type == "cuda" and pc.tensor_parallel_size > 1 and not pc.disable_custom_all_reduce:
            from vllm_b200.parallel.symm import SymmComm
            tp = pstate.get_tp_group()
            tp.device_comm = SymmComm(tp)
        self.compile_or_warm_up_model()
        self._checkpointed = False

    # ---------------------------------------------------------------- LoRA
    def add_lora(self, lora_request=None, tensors=None, rank: int = 0, alpha: float = 0.0) -> bool:
        """Load an adapter from ``lora_request.lora_path`` (PEFT dir) or from in-memory tensors."""
        from vllm_b200.lora import LoRAModel
        mgr = self.runner.lora_manager
        if mgr is None:
            raise RuntimeError("LoRA is not enabled (enable_lora=True)")
        if tensors is not None:
            mgr.add_adapter(LoRAModel.from_tensors(lora_request.lora_int_id, dict(tensors), rank, alpha))
        else:
            mgr.add_adapter(LoRAModel.from_path(lora_request.lora_int_id, lora_request.lora_path))
        return True

    def remove_lora(self, lora_id: int) -> bool:
        return self.runner.lora_manager.remove_adapter(lora_id)

    def list_loras(self) -> list[int]:
        return self.runner.lora_manager.list_adapters()

    def sleep(self, level: int = 1) -> None:
        """Level 1: weights → pinned CPU, KV discarded. Level 2: both discarded."""
        if self._sleeping:
            return
        model = self.runner.model
        self._param_shapes = {n: tuple(p.shape) for n, p in model.named_parameters()}
        if level == 1:
            for n, p in model.named_parameters():
                self._sleep_saved[n] = p.data.to("cpu", non_blocking=False).pin_memory() \
                    if self.device.type == "cuda" else p.data.clone()
        for p in model.parameters():
            p.data = torch.empty(0, dtype=p.dtype, device=self.device)
        self._kv_shapes = {k: (v.shape, v.dtype) for k, v in self.runner.kv_caches.items()}
        for k in list(self.runner.kv_caches):
            self.runner.kv_caches[k] = torch.empty(0, device=self.device)
        self.runner.graphs.clear()
        gc.collect()
        if self.device.type == "cuda":
            torch.cuda.empty_cache()
        self._sleeping = True
        self._sleep_level = level

    def wake_up(self, tags: Optional[list[str]] = None) -> None:
        if not self._sleeping:
            return
        model = self.runner.model
        if self._sleep_level == 1:
            for n, p in model.named_parameters():
                p.data = self._sleep_saved[n].to(self.device, non_blocking=True)
            self._sleep_saved.clear()
        else:
            # level 2 discarded the weights: re-materialise storage; the caller is expected to
            # push fresh weights (update_weights / reload_weights) before serving again
            for n, p in model.named_parameters():
                p.data = torch.zeros(self._param_shapes[n], dtype=p.dtype, device=self.device)
        for k, (shape, dt) in self._kv_shapes.items():
            self.runner.kv_caches[k] = torch.zeros(shape, dtype=dt, device=self.device)
        self._sleeping = False
        if self.device.type == "cuda":
            torch.cuda.synchronize()
            self.runner.capture_model()

    def is_sleeping(self) -> bool:
        return self._sleeping

    # ------------------------------------------------------------- profiler
    def profile(self, is_start: bool = True) -> None:
        from vllm_b200.profiler import make_profiler
        if self.profiler is None:
            self.profiler = make_profiler(self.vllm_config.profiler_config, self.rank)
        if is_start:
            self.profiler.start()
        else:
            self.profiler.stop()

    # -------------------------------------------------------------- weights
    def save_sharded_state(self, path: str) -> str:
        from vllm_b200.model_loader.sharded_state import save_sharded_state
        return save_sharded_state(self.runner.model, path)

    def update_weights(self, named: list[tuple[str, torch.Tensor]]) -> int:
        """RLHF-style in-place weight update (reference: gpu_worker.py:1329)."""
        loaded = self.runner.model.load_weights(iter(named))
        return len(loaded)

    def reload_weights(self) -> None:
        from vllm_b200.model_loader import initialize_dummy_weights, iter_weights
        mc = self.vllm_config.model_config
        with torch.no_grad():
            if os.path.isdir(mc.model):
                self.runner.model.load_weights(iter_weights(mc.model))
            else:
                initialize_dummy_weights(self.runner.model, self.vllm_config.load_config.dummy_seed)

    def check_health(self) -> bool:
        return True

    def get_model(self):
        return self.runner.model

    def shutdown(self) -> None:
        r = self.runner
        if r is not None:  # drop device memory even if someone still holds a reference to the runner
            r.graphs.clear()
            r.graph_pool = None
            r.kv_caches.clear()
            r.model = None
            r.block_tables.clear()
            r.slot_mappings.clear()
        self.runner = None
        pstate.destroy_model_parallel()
        gc.collect()
        if self.device is not None and self.device.type == "cuda":
            torch.cuda.empty_cache()
