"""Per-phase time line of the decode GEMM kernel (csrc/gemm_tc.cu) from its own clock64 stamps, L2 flushed.

    python scripts/gemm_trace.py            # prints one JSON line per shape

Stamps (per CTA): 0 kernel entry · 1 TMA producer starts (prologue done) · 3 first operand stage landed · 4 last MMA
issued · 5 accumulator complete (epilogue released) · 6 partial parked in shared memory · 7 past the cluster barrier ·
8 reduced slice stored · 10 exit.  Reported: median / max over CTAs of each interval in µs at the measured SM clock, plus
the event-timed duration of the launch.
"""
import ctypes
import json
import sys

import torch

sys.path.insert(0, ".")
from vllm_b200.ops import cuda as cu  # noqa: E402

SHAPES = [("8B/TP8 qkv", 768, 4096), ("8B/TP8 o", 4096, 512), ("8B/TP8 gate_up", 3584, 4096), ("8B/TP8 down", 4096, 1792),
          ("8B qkv", 6144, 4096), ("8B o", 4096, 4096), ("70B/TP8 qkv", 1280, 8192), ("70B/TP8 gate_up", 7168, 8192)]
PH = [("prologue", 0, 1), ("first_stage", 1, 3), ("mainloop", 3, 4), ("mma_drain", 4, 5), ("tmem_to_smem", 5, 6),
      ("cluster_barrier", 6, 7), ("dsmem_reduce_store", 7, 8), ("exit", 8, 10), ("total", 0, 10)]


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    lib = cu._lib
    lib.vb_gemm_tc_set_trace.argtypes, lib.vb_gemm_tc_set_trace.restype = [ctypes.c_void_p], None
    mhz = float(torch.cuda.clock_rate()) if hasattr(torch.cuda, "clock_rate") else 1965.0   # cycles / MHz = us
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, N, K in SHAPES:
        x = (torch.randn(M, K, device="cuda") * 0.05).bfloat16()
        w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
        for _ in range(3):
            cu.gemm_tc(x, w)
        torch.cuda.synchronize()
        tr = torch.zeros(1024 * 16, dtype=torch.int64, device="cuda")
        ts = []
        for rep in range(5):
            flush.zero_()
            tr.zero_()
            torch.cuda.synchronize()
            lib.vb_gemm_tc_set_trace(tr.data_ptr())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            cu.gemm_tc(x, w)
            e1.record()
            torch.cuda.synchronize()
            lib.vb_gemm_tc_set_trace(None)
            ts.append(e0.elapsed_time(e1) * 1e3)
        t = tr.view(-1, 16).cpu().double()
        t = t[t[:, 0] > 0]
        res = {"shape": name, "M": M, "N": N, "K": K, "ctas": int(t.shape[0]), "event_us": round(min(ts), 2)}
        for ph, a, b in PH:
            ok = (t[:, a] > 0) & (t[:, b] > 0)
            if ok.any():
                d = (t[ok, b] - t[ok, a]) / mhz
                res[ph] = [round(float(d.median()), 2), round(float(d.max()), 2)]
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
