"""Build every CUDA source in this directory into ``vllm_b200/_C.so`` for sm_100a.

Plain nvcc (cross-compiles without a GPU): one object per .cu in parallel, then one
shared library.  The .so is loaded with ctypes (ops/cuda.py) — no torch headers are
needed, so a full rebuild takes well under a minute.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "_C.so")
OBJ_DIR = os.path.join(HERE, "build")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-fvisibility=hidden", "-DNDEBUG"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _sig(path: str, deps: list[str]) -> str:
    h = hashlib.sha256()
    for p in [path] + deps:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(ARCH + FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = sorted(f for f in os.listdir(HERE) if f.endswith(".cu"))
    hdrs = sorted(os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h")))
    jobs = []
    objs = []
    for s in srcs:
        src = os.path.join(HERE, s)
        obj = os.path.join(OBJ_DIR, s[:-3] + ".o")
        sig_file = obj + ".sig"
        sig = _sig(src, hdrs)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(sig_file) and open(sig_file).read() == sig:
            continue
        jobs.append((src, obj, sig_file, sig))

    def compile_one(job):
        src, obj, sig_file, sig = job
        cmd = [nvcc, *ARCH, *FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        with open(sig_file, "w") as f:
            f.write(sig)
        return src

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for done in ex.map(compile_one, jobs):
                if verbose:
                    print("compiled", os.path.basename(done))
    if jobs or not os.path.exists(OUT):
        r = subprocess.run([nvcc, "-shared", *ARCH, "-o", OUT, *objs, "-lcudart"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
