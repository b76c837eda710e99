"""Worker profilers driven by start/stop RPC (reference: vllm/profiler/wrapper.py:28-489:
``WorkerProfiler`` state machine with torch / CUDA-range back-ends)."""
from __future__ import annotations

import os
from typing import Optional

import torch

from vllm_b200.logger import init_logger

logger = init_logger(__name__)


class WorkerProfiler:
    """delay → active → stopped, with an optional iteration cap."""

    def __init__(self, delay_iterations: int = 0, max_iterations: int = 0) -> None:
        self.delay, self.max_iters = delay_iterations, max_iterations
        self.requested = False
        self.running = False
        self._seen = 0
        self._active_iters = 0

    def start(self) -> None:
        self.requested = True
        self._seen = 0
        if self.delay == 0:
            self._begin()

    def step(self) -> None:
        """Call once per engine iteration."""
        if not self.requested:
            return
        self._seen += 1
        if not self.running and self._seen >= self.delay:
            self._begin()
        elif self.running:
            self._active_iters += 1
            if self.max_iters and self._active_iters >= self.max_iters:
                self.stop()

    def _begin(self) -> None:
        if not self.running:
            self.running = True
            self._active_iters = 0
            self._start()

    def stop(self) -> None:
        if self.running:
            self._stop()
        self.running = self.requested = False

    def _start(self) -> None: ...
    def _stop(self) -> None: ...


class TorchProfiler(WorkerProfiler):
    def __init__(self, out_dir: str, rank: int, with_stack=False, record_shapes=False, **kw) -> None:
        super().__init__(**kw)
        self.out_dir, self.rank = out_dir, rank
        self.with_stack, self.record_shapes = with_stack, record_shapes
        self.prof = None

    def _start(self) -> None:
        acts = [torch.profiler.ProfilerActivity.CPU]
        if torch.cuda.is_available():
            acts.append(torch.profiler.ProfilerActivity.CUDA)
        self.prof = torch.profiler.profile(activities=acts, with_stack=self.with_stack,
                                           record_shapes=self.record_shapes)
        self.prof.start()

    def _stop(self) -> None:
        self.prof.stop()
        os.makedirs(self.out_dir, exist_ok=True)
        path = os.path.join(self.out_dir, f"trace_rank{self.rank}.json.gz")
        self.prof.export_chrome_trace(path)
        logger.info("profiler trace written to %s", path)
        try:
            logger.info("\n%s", self.prof.key_averages().table(sort_by="self_cuda_time_total"
                                                               if torch.cuda.is_available() else "self_cpu_time_total",
                                                               row_limit=20))
        except Exception:
            pass
        self.prof = None


class CudaProfiler(WorkerProfiler):
    """cudaProfilerStart/Stop capture ranges for ncu / nsys."""

    def _start(self) -> None:
        if torch.cuda.is_available():
            torch.cuda.profiler.start()

    def _stop(self) -> None:
        if torch.cuda.is_available():
            torch.cuda.profiler.stop()


def make_profiler(cfg, rank: int) -> WorkerProfiler:
    kind = cfg.profiler or ("torch" if (cfg.torch_profiler_dir or os.environ.get("VLLM_TORCH_PROFILER_DIR")) else "cuda")
    kw = dict(delay_iterations=cfg.delay_iterations, max_iterations=cfg.max_iterations)
    if kind == "torch":
        d = cfg.torch_profiler_dir or os.environ.get("VLLM_TORCH_PROFILER_DIR") or "./vllm_b200_profile"
        return TorchProfiler(d, rank, cfg.with_stack, cfg.record_shapes, **kw)
    return CudaProfiler(**kw)
