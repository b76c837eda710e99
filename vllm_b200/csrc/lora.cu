// Multi-LoRA batched matrix-vector kernels (BGMV): every token of the batch applies the adapter of ITS request.
//
// Reference: vllm/lora/ops/triton_ops/ (lora_shrink / lora_expand: SGMV + BGMV Triton kernels driven by per-token
// adapter indices; punica-style).  Ours: two sm_100a launches per LoRA-carrying linear layer, both indexed by a
// device-resident token → slot table (graph capturable, no host-side grouping, no index_select / index_add):
//
//   shrink  v[t, j, :] = x[t, :] · A[slot(t), j]^T          (A stacked  [slots, n_sub, r, K])
//   expand  y[t, n]   += v[t, sub(n), :] · B[slot(t), n, :]  (B stacked  [slots, N, r], scaling folded in)
//
// `n_sub` covers packed projections (q|k|v, gate|up) whose slices own separate A matrices.  Tokens with slot < 0
// (base model) exit immediately.  Work per token is O(r · (K + N)) bytes of adapter weights — bandwidth bound, so
// the kernels are plain vectorised SIMT: 16-byte loads, fp32 accumulation, warp-shuffle reductions.
#include "common.cuh"

namespace {

// grid (T, n_sub), block 128: warp w computes ranks w, w + 4, ...
template <typename T>
__global__ void __launch_bounds__(128) lora_shrink_kernel(const T* __restrict__ x, int64_t x_stride,
                                                           const T* __restrict__ A, const int64_t* __restrict__ slots,
                                                           float* __restrict__ v, int K, int r, int n_sub) {
  pdl_trigger();
  pdl_wait();
  const int t = blockIdx.x, j = blockIdx.y;
  const int64_t s = slots[t];
  if (s < 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const T* xr = x + t * x_stride;
  const T* Aj = A + ((s * n_sub + j) * (int64_t)r) * K;
  const int nvec = K / 8;
  for (int rr = warp; rr < r; rr += 4) {
    const T* ar = Aj + (int64_t)rr * K;
    float acc = 0.f;
    for (int i = lane; i < nvec; i += 32) {
      Vec16<T> a = ld16(ar + i * 8), b = ld16(xr + i * 8);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc += to_f32(a.e[e]) * to_f32(b.e[e]);
    }
    acc = warp_sum(acc);
    if (lane == 0) v[((int64_t)t * n_sub + j) * r + rr] = acc;
  }
}

// grid (T, ceil(N / 256)), block 256: one output feature per thread
template <typename T>
__global__ void __launch_bounds__(256) lora_expand_kernel(T* __restrict__ y, int64_t y_stride, const T* __restrict__ B,
                                                           const int64_t* __restrict__ slots, const float* __restrict__ v,
                                                           int N, int r, int n_sub, int o1, int o2, int o3) {
  extern __shared__ float vs[];            // [n_sub, r]
  pdl_trigger();
  pdl_wait();
  const int t = blockIdx.x;
  const int64_t s = slots[t];
  if (s < 0) return;
  for (int i = threadIdx.x; i < n_sub * r; i += blockDim.x) vs[i] = v[(int64_t)t * n_sub * r + i];
  __syncthreads();
  const int n = blockIdx.y * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int sub = (n_sub > 1 && n >= o1) + (n_sub > 2 && n >= o2) + (n_sub > 3 && n >= o3);
  const float* vv = vs + sub * r;
  const T* br = B + (s * (int64_t)N + n) * r;
  float acc = 0.f;
  for (int i = 0; i < r; i += 8) {
    Vec16<T> b = ld16(br + i);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc += to_f32(b.e[e]) * vv[i + e];
  }
  T* yp = y + t * y_stride + n;
  *yp = from_f32<T>(to_f32(*yp) + acc);
}

}  // namespace

EXPORT int vb_lora_shrink(const void* x, int64_t x_stride, const void* A, const int64_t* slots, float* v, int T, int K, int r,
                          int n_sub, int dtype, cudaStream_t st) {
  if (T == 0) return 0;
  if (dtype == DT_F32 || K % 8 || r < 1) return -1;
  DISPATCH_HALF(dtype, TT, {
    vb_launch(lora_shrink_kernel<TT>, dim3(T, n_sub), dim3(128), 0, st, reinterpret_cast<const TT*>(x), x_stride,
              reinterpret_cast<const TT*>(A), slots, v, K, r, n_sub);
  });
  return (int)cudaGetLastError();
}

// sub-projection boundaries o1 <= o2 <= o3 (unused ones = N)
EXPORT int vb_lora_expand(void* y, int64_t y_stride, const void* B, const int64_t* slots, const float* v, int T, int N, int r,
                          int n_sub, int o1, int o2, int o3, int dtype, cudaStream_t st) {
  if (T == 0) return 0;
  if (dtype == DT_F32 || r % 8 || n_sub < 1 || n_sub > 4) return -1;
  DISPATCH_HALF(dtype, TT, {
    vb_launch(lora_expand_kernel<TT>, dim3(T, (N + 255) / 256), dim3(256), (size_t)n_sub * r * 4, st,
              reinterpret_cast<TT*>(y), y_stride, reinterpret_cast<const TT*>(B), slots, v, N, r, n_sub, o1, o2, o3);
  });
  return (int)cudaGetLastError();
}
