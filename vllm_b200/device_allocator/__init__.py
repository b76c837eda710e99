"""Sleep-mode memory arena (reference: vllm/device_allocator/cumem.py:82 ``CuMemAllocator`` over
csrc/cumem_allocator.cpp).

Tensors allocated inside ``use_memory_pool(tag)`` live in CUDA virtual-memory-management blocks
(``native/cumem.cpp``).  ``sleep()`` optionally copies tagged blocks to pinned host memory, then unmaps and
releases their physical memory while keeping the virtual addresses reserved; ``wake_up()`` maps fresh physical
memory at the same addresses and restores the saved contents.  Because addresses never change, CUDA graphs
captured before the sleep remain valid afterwards.
"""
from __future__ import annotations

import contextlib
import ctypes
from typing import Optional

import torch

from vllm_b200.logger import init_logger
from vllm_b200.native import OUT as NATIVE_SO
from vllm_b200.native import lib as native_lib

logger = init_logger(__name__)


class CuMemAllocator:
    _instance: Optional["CuMemAllocator"] = None
    default_tag = "default"

    @classmethod
    def get_instance(cls) -> "CuMemAllocator":
        if cls._instance is None:
            cls._instance = cls()
        return cls._instance

    def __init__(self) -> None:
        self.lib = native_lib()
        if self.lib is None or not self.lib.vbn_cumem_available():
            raise RuntimeError("CUDA VMM arena unavailable (no driver / native library)")
        self.allocator = torch.cuda.memory.CUDAPluggableAllocator(NATIVE_SO, "vbn_torch_malloc", "vbn_torch_free")
        self.pool = None
        self.tags: dict[int, str] = {}          # block ptr → tag
        self.saved: dict[int, torch.Tensor] = {}
        self.current_tag = self.default_tag

    def _blocks(self) -> list[tuple[int, int, bool]]:
        cap = 1 << 16
        buf = (ctypes.c_uint64 * (3 * cap))()
        self.lib.vbn_cumem_list.argtypes = [ctypes.c_void_p, ctypes.c_int]
        self.lib.vbn_cumem_list.restype = ctypes.c_int
        n = self.lib.vbn_cumem_list(buf, cap)
        return [(int(buf[3 * i]), int(buf[3 * i + 1]), bool(buf[3 * i + 2])) for i in range(min(n, cap))]

    @contextlib.contextmanager
    def use_memory_pool(self, tag: Optional[str] = None):
        """Route allocations of the enclosed region into the VMM arena and label them ``tag``."""
        tag = tag or self.default_tag
        before = {p for p, _, _ in self._blocks()}
        if self.pool is None:
            self.pool = torch.cuda.MemPool(self.allocator.allocator())
        with torch.cuda.use_mem_pool(self.pool):
            yield
        for p, _, _ in self._blocks():
            if p not in before:
                self.tags[p] = tag

    def sleep(self, offload_tags: Optional[tuple[str, ...]] = None) -> int:
        """Unmap every arena block; blocks whose tag is in ``offload_tags`` are saved to pinned host memory first
        (weights), the others are discarded (KV cache).  Returns the number of bytes given back to the device."""
        offload_tags = (self.default_tag,) if offload_tags is None else offload_tags
        torch.cuda.synchronize()
        freed = 0
        for p, size, mapped in self._blocks():
            if not mapped:
                continue
            if self.tags.get(p, self.default_tag) in offload_tags:
                host = torch.empty(size, dtype=torch.uint8, pin_memory=True)
                ctypes.cdll.LoadLibrary("libcudart.so").cudaMemcpy(ctypes.c_void_p(host.data_ptr()), ctypes.c_void_p(p),
                                                                   ctypes.c_size_t(size), 2)
                self.saved[p] = host
            self.lib.vbn_cumem_unmap(p)
            freed += size
        logger.info("sleep: released %.2f GiB of device memory (%.2f GiB kept on the host)", freed / 2**30,
                    sum(t.numel() for t in self.saved.values()) / 2**30)
        return freed

    def wake_up(self, tags: Optional[list[str]] = None) -> None:
        for p, size, mapped in self._blocks():
            if mapped or (tags is not None and self.tags.get(p, self.default_tag) not in tags):
                continue
            rc = self.lib.vbn_cumem_remap(p)
            if rc != 0:
                raise RuntimeError(f"cuMemMap failed while waking up (rc={rc}): out of device memory?")
            host = self.saved.pop(p, None)
            if host is not None:
                ctypes.cdll.LoadLibrary("libcudart.so").cudaMemcpy(ctypes.c_void_p(p), ctypes.c_void_p(host.data_ptr()),
                                                                   ctypes.c_size_t(size), 1)
        torch.cuda.synchronize()

    def get_current_usage(self) -> int:
        return int(self.lib.vbn_cumem_mapped_bytes())
