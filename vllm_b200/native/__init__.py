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
