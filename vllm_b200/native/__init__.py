"""Native (C++) runtime helpers built in-tree: spin-wait shm flags, GIL-free file IO, CUDA VMM allocator.
See ``vllm_b200/native/*.cpp``; ``build_native()`` compiles them with g++ into ``_native.so``."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "_native.so")


def build_native(force: bool = False) -> str:
    srcs = sorted(os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".cpp"))
    if not srcs:
        return ""
    newest = max(os.path.getmtime(s) for s in srcs)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= newest:
        return OUT
    cuda_inc = "/usr/local/cuda/include"
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", OUT, *srcs, f"-I{cuda_inc}", "-lpthread", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"native build failed:\n{r.stderr}")
    return OUT


_LIB = None


def lib():
    """ctypes handle of ``_native.so`` (built on first use; None if no compiler is available)."""
    global _LIB
    if _LIB is None:
        import ctypes
        try:
            path = build_native()
            if not path:
                _LIB = False
                return None
            L = ctypes.CDLL(path)
        except Exception:  # pragma: no cover - no toolchain
            _LIB = False
            return None
        u64p, i64, u64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int
        L.vbn_wait_ge_u64.argtypes = [u64p, u64, i64]; L.vbn_wait_ge_u64.restype = i32
        L.vbn_wait_all_ge_u64.argtypes = [u64p, i32, i64, u64, i64]; L.vbn_wait_all_ge_u64.restype = i32
        L.vbn_store_release_u64.argtypes = [u64p, u64]; L.vbn_store_release_u64.restype = None
        L.vbn_load_acquire_u64.argtypes = [u64p]; L.vbn_load_acquire_u64.restype = u64
        L.vbn_write_file.argtypes = [ctypes.c_char_p, ctypes.c_void_p, i64, i32, i32]; L.vbn_write_file.restype = i64
        L.vbn_read_file.argtypes = [ctypes.c_char_p, ctypes.c_void_p, i64, i64, i32]; L.vbn_read_file.restype = i64
        L.vbn_file_size.argtypes = [ctypes.c_char_p]; L.vbn_file_size.restype = i64
        L.vbn_cumem_available.restype = i32
        L.vbn_cumem_alloc.argtypes = [i32, u64, ctypes.POINTER(u64)]; L.vbn_cumem_alloc.restype = u64
        for f in ("vbn_cumem_unmap", "vbn_cumem_remap", "vbn_cumem_free"):
            getattr(L, f).argtypes = [u64]; getattr(L, f).restype = i32
        L.vbn_cumem_mapped_bytes.restype = u64
        _LIB = L
    return _LIB or None


LOADGEN = os.path.join(os.path.dirname(HERE), "_loadgen")


def build_loadgen(force: bool = False) -> str:
    """Compile ``native/tools/loadgen.cpp`` (epoll-based OpenAI load generator) into ``vllm_b200/_loadgen``."""
    src = os.path.join(HERE, "tools", "loadgen.cpp")
    if not force and os.path.exists(LOADGEN) and os.path.getmtime(LOADGEN) >= os.path.getmtime(src):
        return LOADGEN
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-o", LOADGEN, src], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"loadgen build failed:\n{r.stderr}")
    return LOADGEN


FRONTD = os.path.join(os.path.dirname(HERE), "_frontd")


def build_frontd(force: bool = False) -> str:
    """Compile ``native/frontend/frontd.cpp`` (C++ HTTP <-> engine-core front-end daemon) into ``vllm_b200/_frontd``."""
    srcs = [os.path.join(HERE, "frontend", f) for f in ("frontd.cpp", "wire.hpp", "tokenizer.hpp")]
    if not force and os.path.exists(FRONTD) and os.path.getmtime(FRONTD) >= max(os.path.getmtime(s) for s in srcs):
        return FRONTD
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-o", FRONTD, srcs[0], "-ldl", "-lpthread"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"frontd build failed:\n{r.stderr}")
    return FRONTD
