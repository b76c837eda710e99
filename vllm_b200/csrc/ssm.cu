// State-space / linear-attention kernels over the paged recurrent-state cache (sm_100a):
//   causal depthwise conv1d with cached window     (Mamba-1/2, gated delta net; prefill, chunked prefill, decode)
//   Mamba-1 selective scan                          (per-channel A, state N <= 32)
//   Mamba-2 SSD scan / selective_state_update       (scalar decay per head, state [hd, N] per head)
//   gated delta rule scan                           (Qwen3-Next GDN: rank-1 corrected state [dk, dv] per head)
//
// Reference ops: csrc/libtorch_stable/mamba/selective_scan_fwd.cu, vllm/model_executor/layers/mamba/ops/
// causal_conv1d.py, mamba_ssm.py (selective_state_update), ssd_combined.py (+ ssd_chunk_scan / state passing),
// fla ops/fused_recurrent (GDN decode).  One design replaces that family here: every op is a varlen kernel driven
// by device-resident `query_start_loc` / `seq_lens` / cache-row ids, reads and writes the request's state row in
// the cache pool in place, and walks the request's tokens in order with the state held in registers — decode is
// the q = 1 case of the same kernel, so there is one code path, no host loop, no `.tolist()`, and every batch shape
// is CUDA-graph capturable.  A linear-time scan keeps all 148 SMs busy because the independent (request, head,
// state-row) warps number in the thousands even for one sequence.
//
// State row of request r (fp32): [ conv window (C x (k-1)) | recurrent state ... ]; `state_stride` floats per row.
#include "common.cuh"

namespace {

VB_DEVICE float silu_f(float x) { return x / (1.f + __expf(-x)); }
VB_DEVICE float softplus_f(float x) { return x > 20.f ? x : log1pf(__expf(x)); }

// ------------------------------------------------------------------ causal conv1d (+ SiLU), window cached per request
// grid (num_reqs, ceil(C / 256)); thread = one channel, walks the request's tokens with the k-1 window in registers.
template <typename T, int K>
__global__ void __launch_bounds__(256) causal_conv1d_kernel(const T* __restrict__ x, int64_t x_stride, float* __restrict__ out,
                                                             int64_t out_stride, const T* __restrict__ w,
                                                             const T* __restrict__ bias, float* __restrict__ state,
                                                             int64_t state_stride, const int* __restrict__ qsl,
                                                             const int* __restrict__ seq_lens,
                                                             const int* __restrict__ rows, int64_t rows_stride, int C,
                                                             int act) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  const int t0 = qsl[r], t1 = qsl[r + 1];
  if (c >= C || t1 <= t0) return;
  const bool has_init = seq_lens[r] - (t1 - t0) > 0;
  float* st = state + (int64_t)rows[r * rows_stride] * state_stride + (int64_t)c * (K - 1);
  float win[K];
#pragma unroll
  for (int j = 0; j < K - 1; ++j) win[j] = has_init ? st[j] : 0.f;
  float wr[K];
#pragma unroll
  for (int j = 0; j < K; ++j) wr[j] = to_f32(w[(int64_t)c * K + j]);
  const float b = bias ? to_f32(bias[c]) : 0.f;
  for (int t = t0; t < t1; ++t) {
    win[K - 1] = to_f32(x[t * x_stride + c]);
    float acc = b;
#pragma unroll
    for (int j = 0; j < K; ++j) acc += wr[j] * win[j];
    out[t * out_stride + c] = act ? silu_f(acc) : acc;
#pragma unroll
    for (int j = 0; j < K - 1; ++j) win[j] = win[j + 1];
  }
#pragma unroll
  for (int j = 0; j < K - 1; ++j) st[j] = win[j];
}

// ------------------------------------------------------------------ Mamba-1 selective scan
// thread = (request, channel); h[N] in registers.  x, dt: [T, di] fp32 (dt already softplus'ed); B, C: [T, N] fp32;
// A: [di, N] (negative); D: [di]; z: optional [T, 2*di]-strided gate (silu) — y = (scan + D x) * silu(z).
template <int N>
__global__ void __launch_bounds__(128) selective_scan_kernel(const float* __restrict__ x, const float* __restrict__ dt,
                                                              const float* __restrict__ Bm, const float* __restrict__ Cm,
                                                              const float* __restrict__ A, const float* __restrict__ D,
                                                              float* __restrict__ y, float* __restrict__ state,
                                                              int64_t state_stride, int64_t state_off,
                                                              const int* __restrict__ qsl, const int* __restrict__ seq_lens,
                                                              const int* __restrict__ rows, int64_t rows_stride, int di) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x;
  const int d = blockIdx.y * blockDim.x + threadIdx.x;
  const int t0 = qsl[r], t1 = qsl[r + 1];
  if (d >= di || t1 <= t0) return;
  const bool has_init = seq_lens[r] - (t1 - t0) > 0;
  float* st = state + (int64_t)rows[r * rows_stride] * state_stride + state_off + (int64_t)d * N;
  float h[N], a[N];
#pragma unroll
  for (int n = 0; n < N; ++n) { h[n] = has_init ? st[n] : 0.f; a[n] = A[(int64_t)d * N + n]; }
  const float dd = D[d];
  for (int t = t0; t < t1; ++t) {
    const float dtv = dt[(int64_t)t * di + d], xv = x[(int64_t)t * di + d];
    const float dx = dtv * xv;
    float acc = 0.f;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      h[n] = __expf(dtv * a[n]) * h[n] + dx * Bm[(int64_t)t * N + n];
      acc += h[n] * Cm[(int64_t)t * N + n];
    }
    y[(int64_t)t * di + d] = acc + dd * xv;
  }
#pragma unroll
  for (int n = 0; n < N; ++n) st[n] = h[n];
}

// ------------------------------------------------------------------ Mamba-2 SSD scan (decode == q = 1)
// warp = (request, head, p); lane holds NPL = N / 32 state elements of row p.
//   h[p, n] = exp(dt·A_h) h[p, n] + dt x[p] B[g, n];   y[p] = Σ_n h[p, n] C[g, n] + D_h x[p]
// xbc: [T, conv_dim] fp32 = x (nh*hd) | B (G*N) | C (G*N) (the conv1d output); dt_raw: [T, nh] (+dt_bias, softplus,
// clamp in-kernel).
template <int NPL>
__global__ void __launch_bounds__(256) ssd_scan_kernel(const float* __restrict__ xbc, int64_t xbc_stride,
                                                        const float* __restrict__ dt_raw, int64_t dt_stride,
                                                        const float* __restrict__ dt_bias, const float* __restrict__ A_log,
                                                        const float* __restrict__ D, float* __restrict__ y,
                                                        float* __restrict__ state, int64_t state_stride, int64_t state_off,
                                                        const int* __restrict__ qsl, const int* __restrict__ seq_lens,
                                                        const int* __restrict__ rows, int64_t rows_stride, int nh, int hd,
                                                        int G, float dt_min, float dt_max) {
  pdl_trigger();
  pdl_wait();
  constexpr int N = NPL * 32;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.y;
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // (head, p)
  const int t0 = qsl[r], t1 = qsl[r + 1];
  if (wid >= nh * hd || t1 <= t0) return;
  const int hh = wid / hd, p = wid % hd;
  const int g = hh / (nh / G);
  const bool has_init = seq_lens[r] - (t1 - t0) > 0;
  float* st = state + (int64_t)rows[r * rows_stride] * state_stride + state_off + ((int64_t)hh * hd + p) * N + lane * NPL;
  float h[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) h[i] = has_init ? st[i] : 0.f;
  const float a = -__expf(A_log[hh]), dd = D[hh], db = dt_bias[hh];
  const int d_inner = nh * hd;
  for (int t = t0; t < t1; ++t) {
    const float* row = xbc + (int64_t)t * xbc_stride;
    float dtv = softplus_f(dt_raw[(int64_t)t * dt_stride + hh] + db);
    dtv = fminf(fmaxf(dtv, dt_min), dt_max);
    const float xv = row[hh * hd + p];
    const float dA = __expf(dtv * a), dx = dtv * xv;
    const float* Bp = row + d_inner + g * N + lane * NPL;
    const float* Cp = row + d_inner + G * N + g * N + lane * NPL;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      h[i] = dA * h[i] + dx * Bp[i];
      acc += h[i] * Cp[i];
    }
    acc = warp_sum(acc);
    if (lane == 0) y[(int64_t)t * d_inner + hh * hd + p] = acc + dd * xv;
  }
#pragma unroll
  for (int i = 0; i < NPL; ++i) st[i] = h[i];
}

// ------------------------------------------------------------------ gated delta rule scan (GDN)
// warp = (request, value head, column j of the [dk, dv] state); lane holds KPL = dk / 32 rows of that column.
//   S *= exp(g);  delta_j = (v_j - Σ_i S[i,j] k_i) β;  S[i,j] += k_i delta_j;  o_j = Σ_i S[i,j] q_i
// q, k: [T, nv, dk] fp32 (already l2-normalised / scaled / repeated to value heads); v: [T, nv, dv]; g, beta: [T, nv].
template <int KPL>
__global__ void __launch_bounds__(256) gdn_scan_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                        const float* __restrict__ v, const float* __restrict__ g,
                                                        const float* __restrict__ beta, float* __restrict__ o,
                                                        float* __restrict__ state, int64_t state_stride, int64_t state_off,
                                                        const int* __restrict__ qsl, const int* __restrict__ seq_lens,
                                                        const int* __restrict__ rows, int64_t rows_stride, int nv, int dv) {
  pdl_trigger();
  pdl_wait();
  constexpr int DK = KPL * 32;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.y;
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // (head, j)
  const int t0 = qsl[r], t1 = qsl[r + 1];
  if (wid >= nv * dv || t1 <= t0) return;
  const int hh = wid / dv, j = wid % dv;
  const bool has_init = seq_lens[r] - (t1 - t0) > 0;
  // state layout [nv, dk, dv]: element (i, j) at ((hh * DK + i) * dv + j); lane owns rows i = lane + 32 * u
  float* st = state + (int64_t)rows[r * rows_stride] * state_stride + state_off + (int64_t)hh * DK * dv + j;
  float S[KPL];
#pragma unroll
  for (int u = 0; u < KPL; ++u) S[u] = has_init ? st[(int64_t)(lane + 32 * u) * dv] : 0.f;
  for (int t = t0; t < t1; ++t) {
    const float* kp = k + ((int64_t)t * nv + hh) * DK;
    const float* qp = q + ((int64_t)t * nv + hh) * DK;
    const float decay = __expf(g[(int64_t)t * nv + hh]);
    float kk[KPL], mem = 0.f;
#pragma unroll
    for (int u = 0; u < KPL; ++u) {
      kk[u] = kp[lane + 32 * u];
      S[u] *= decay;
      mem += S[u] * kk[u];
    }
    mem = warp_sum(mem);
    const float delta = (v[((int64_t)t * nv + hh) * dv + j] - mem) * beta[(int64_t)t * nv + hh];
    float acc = 0.f;
#pragma unroll
    for (int u = 0; u < KPL; ++u) {
      S[u] += kk[u] * delta;
      acc += S[u] * qp[lane + 32 * u];
    }
    acc = warp_sum(acc);
    if (lane == 0) o[((int64_t)t * nv + hh) * dv + j] = acc;
  }
#pragma unroll
  for (int u = 0; u < KPL; ++u) st[(int64_t)(lane + 32 * u) * dv] = S[u];
}

}  // namespace

// x: [T, C] (T-typed), out: [T, C] fp32; state rows hold the (k-1)-window at offset 0 ([C, k-1] fp32).
// rows: cache-row id of each request (block_table[:, 0], int32, stride rows_stride).
EXPORT int vb_causal_conv1d(const void* x, int64_t x_stride, float* out, int64_t out_stride, const void* w, const void* bias,
                            float* state, int64_t state_stride, const int* qsl, const int* seq_lens, const int* rows,
                            int64_t rows_stride, int num_reqs, int C, int K, int act, int dtype, cudaStream_t st) {
  if (num_reqs == 0) return 0;
  if (dtype == DT_F32 || K < 2 || K > 4) return -1;
  dim3 grid(num_reqs, (C + 255) / 256);
#define GO(KK) vb_launch(causal_conv1d_kernel<TT, KK>, grid, dim3(256), 0, st, reinterpret_cast<const TT*>(x), x_stride, out, \
                         out_stride, reinterpret_cast<const TT*>(w), reinterpret_cast<const TT*>(bias), state, state_stride,   \
                         qsl, seq_lens, rows, rows_stride, C, act)
  DISPATCH_HALF(dtype, TT, {
    if (K == 2) GO(2); else if (K == 3) GO(3); else GO(4);
  });
#undef GO
  return (int)cudaGetLastError();
}

EXPORT int vb_selective_scan(const float* x, const float* dt, const float* B, const float* C, const float* A, const float* D,
                             float* y, float* state, int64_t state_stride, int64_t state_off, const int* qsl,
                             const int* seq_lens, const int* rows, int64_t rows_stride, int num_reqs, int di, int N,
                             cudaStream_t st) {
  if (num_reqs == 0) return 0;
  dim3 grid(num_reqs, (di + 127) / 128);
#define GO(NN) vb_launch(selective_scan_kernel<NN>, grid, dim3(128), 0, st, x, dt, B, C, A, D, y, state, state_stride, \
                         state_off, qsl, seq_lens, rows, rows_stride, di)
  if (N == 16) GO(16); else if (N == 8) GO(8); else if (N == 32) GO(32); else return -1;
#undef GO
  return (int)cudaGetLastError();
}

EXPORT int vb_ssd_scan(const float* xbc, int64_t xbc_stride, const float* dt_raw, int64_t dt_stride, const float* dt_bias,
                       const float* A_log, const float* D, float* y, float* state, int64_t state_stride, int64_t state_off,
                       const int* qsl, const int* seq_lens, const int* rows, int64_t rows_stride, int num_reqs, int nh, int hd,
                       int G, int N, float dt_min, float dt_max, cudaStream_t st) {
  if (num_reqs == 0) return 0;
  if (N % 32 || N > 256 || nh % G) return -1;
  const int wpb = 8;
  dim3 grid((nh * hd + wpb - 1) / wpb, num_reqs);
#define GO(P) vb_launch(ssd_scan_kernel<P>, grid, dim3(wpb * 32), 0, st, xbc, xbc_stride, dt_raw, dt_stride, dt_bias, A_log, D, \
                        y, state, state_stride, state_off, qsl, seq_lens, rows, rows_stride, nh, hd, G, dt_min, dt_max)
  switch (N / 32) {
    case 1: GO(1); break;
    case 2: GO(2); break;
    case 4: GO(4); break;
    case 8: GO(8); break;
    default: return -1;
  }
#undef GO
  return (int)cudaGetLastError();
}

EXPORT int vb_gdn_scan(const float* q, const float* k, const float* v, const float* g, const float* beta, float* o,
                       float* state, int64_t state_stride, int64_t state_off, const int* qsl, const int* seq_lens,
                       const int* rows, int64_t rows_stride, int num_reqs, int nv, int dk, int dv, cudaStream_t st) {
  if (num_reqs == 0) return 0;
  if (dk % 32 || dk > 256) return -1;
  const int wpb = 8;
  dim3 grid((nv * dv + wpb - 1) / wpb, num_reqs);
#define GO(P) vb_launch(gdn_scan_kernel<P>, grid, dim3(wpb * 32), 0, st, q, k, v, g, beta, o, state, state_stride, state_off, \
                        qsl, seq_lens, rows, rows_stride, nv, dv)
  switch (dk / 32) {
    case 1: GO(1); break;
    case 2: GO(2); break;
    case 4: GO(4); break;
    case 8: GO(8); break;
    default: return -1;
  }
#undef GO
  return (int)cudaGetLastError();
}
