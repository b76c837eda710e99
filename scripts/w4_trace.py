"""clock64 stamps of CTA 0 of the W4A16 GEMM: where a k-block's time goes (converter thread 0 + the MMA thread)."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from vllm_b200.ops import cuda as cu, ref
lib = cu._lib
lib.vb_gemm_w4_set_trace.argtypes, lib.vb_gemm_w4_set_trace.restype = [ctypes.c_void_p], None
N, K, M = 6144, 4096, 64
w = torch.randn(N, K, device="cuda") * 0.05
x = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
codes, sc, g = ref.fp4_quant(w, 16)
tiled = cu.pack_w4_tiles(codes, sc, "nvfp4")
for _ in range(3):
    cu.gemm_w4(x, tiled, N, g, "nvfp4")
tr = torch.zeros(256 * 8, dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
lib.vb_gemm_w4_set_trace(tr.data_ptr())
cu.gemm_w4(x, tiled, N, g, "nvfp4")
torch.cuda.synchronize()
lib.vb_gemm_w4_set_trace(None)
t = tr.view(256, 8)[: K // 64].cpu().double()
mhz = 1965.0
names = ["wait_full", "load+scale+wait_empty", "convert+store", "fence", "arrive→next_iter"]
d = [(t[:, 1] - t[:, 0]), (t[:, 2] - t[:, 1]), (t[:, 3] - t[:, 2]), (t[:, 4] - t[:, 3])]
nxt = t[1:, 0] - t[:-1, 4]
print("per k-block period (cycles):", float((t[1:, 0] - t[:-1, 0]).median()))
for n, v in zip(names, d + [nxt]):
    print(f"  {n:26s} median {float(v.median()):8.0f} cyc   mean {float(v.mean()):8.0f}")
print("  MMA issue lag after converter arrive (cycles):", float((t[:, 5] - t[:, 4]).median()))
print("  MMA period:", float((t[1:, 5] - t[:-1, 5]).median()))
