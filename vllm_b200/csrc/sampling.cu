// Sampling kernels: row argmax, fused temperature + Gumbel-max sampling, top-k/top-p/min-p
// threshold filtering, repetition / presence / frequency penalties.
//
// Reference: v1/sample/ops/topk_topp_sampler.py (FlashInfer sampling kernels or Triton),
// worker/gpu/sample/gumbel.py:182 (Triton Gumbel sampler), csrc sampler.cu (penalties).
// One CTA per row; logits are read once with 16-byte loads straight from the LM-head output
// dtype (no fp32 staging copy); the noise is a counter-based hash of (seed, step, token) so
// the result is independent of launch geometry and reproducible per request seed.
#include "common.cuh"

namespace {

VB_DEVICE uint32_t hash3(uint64_t seed, uint32_t step, uint32_t idx) {
  // 2 rounds of a 64-bit mix (splitmix64 finaliser) over (seed, step, idx)
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * ((uint64_t)step << 32 | idx);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (uint32_t)(z >> 32);
}
VB_DEVICE float gumbel(uint64_t seed, uint32_t step, uint32_t idx) {
  const float u = ((float)(hash3(seed, step, idx) >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
  return -__logf(-__logf(u));
}

struct ArgMax { float v; int i; };
VB_DEVICE ArgMax better(ArgMax a, ArgMax b) { return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a; }

template <typename T, bool SAMPLE>
__global__ void __launch_bounds__(1024) sample_kernel(const T* __restrict__ logits, int64_t row_stride, int V,
                                                       const float* __restrict__ temperature,
                                                       const int64_t* __restrict__ seeds, const int* __restrict__ step_ptr,
                                                       int64_t* __restrict__ out, int idx_off, float4* __restrict__ cand) {
  constexpr int N = Vec16<T>::N;
  __shared__ float sv[32];
  __shared__ int si[32];
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x;
  const T* lr = logits + (int64_t)row * row_stride;
  float temp = 0.f;
  uint64_t seed = 0;
  uint32_t step = 0;
  if (SAMPLE) {
    temp = temperature[row];
    seed = (uint64_t)seeds[row];
    step = (uint32_t)step_ptr[row];
  }
  const bool noisy = SAMPLE && temp >= 1e-5f;
  const float inv_t = noisy ? 1.f / temp : 1.f;
  ArgMax best{-INFINITY, 0x7fffffff};
  const int nvec = V / N;
  for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) {
    Vec16<T> v = ld16(lr + vi * N);
#pragma unroll
    for (int j = 0; j < N; ++j) {
      float x = to_f32(v.e[j]);
      const int idx = idx_off + vi * N + j;        // global vocabulary id: the noise is indexed by it (TP-invariant)
      if (noisy) x = x * inv_t + gumbel(seed, step, (uint32_t)idx);
      best = better(best, ArgMax{x, idx});
    }
  }
  for (int i = nvec * N + threadIdx.x; i < V; i += blockDim.x) {
    float x = to_f32(lr[i]);
    const int idx = idx_off + i;
    if (noisy) x = x * inv_t + gumbel(seed, step, (uint32_t)idx);
    best = better(best, ArgMax{x, idx});
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgMax other{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
    best = better(best, other);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sv[w] = best.v; si[w] = best.i; }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    best = lane < nw ? ArgMax{sv[lane], si[lane]} : ArgMax{-INFINITY, 0x7fffffff};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ArgMax other{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
      best = better(best, other);
    }
    if (lane == 0) {
      if (cand) cand[row] = make_float4(best.v, __int_as_float(best.i), 0.f, 0.f);   // this rank's candidate
      else out[row] = best.i == 0x7fffffff ? 0 : best.i;
    }
  }
}

// final pick over the per-rank candidates gathered as [B, W] x (value, index): same tie-break as the single-GPU kernel
__global__ void select_candidate_kernel(const float4* __restrict__ cands, int W, int64_t* __restrict__ out, int B) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= B) return;
  ArgMax best{-INFINITY, 0x7fffffff};
  for (int r = 0; r < W; ++r) {
    const float4 c = cands[(size_t)row * W + r];
    best = better(best, ArgMax{c.x, __float_as_int(c.y)});
  }
  out[row] = best.i == 0x7fffffff ? 0 : best.i;
}

// In-place top-k / top-p / min-p masking of fp32 logits (already divided by temperature).
// Per row: find the k-th largest value and the top-p probability threshold with a bisection on
// the value axis (each pass is one streaming read from L2/SMEM-resident row), then mask.
__global__ void __launch_bounds__(1024) topk_topp_kernel(float* __restrict__ logits, int V,
                                                          const int* __restrict__ top_k,
                                                          const float* __restrict__ top_p,
                                                          const float* __restrict__ min_p) {
  __shared__ float red[32];
  __shared__ float bc;
  const int row = blockIdx.x;
  float* lr = logits + (int64_t)row * V;
  const int k = top_k ? top_k[row] : 0;
  const float pth = top_p ? top_p[row] : 1.f;
  const float mp = min_p ? min_p[row] : 0.f;
  // row max / min
  float mx = -INFINITY, mn = INFINITY;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    const float x = lr[i];
    mx = fmaxf(mx, x);
    if (x > -INFINITY) mn = fminf(mn, x);
  }
  mx = warp_max(mx);
  mn = -warp_max(-mn);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) red[w] = mx;
  __syncthreads();
  if (w == 0) { float r = lane < nw ? red[lane] : -INFINITY; r = warp_max(r); if (lane == 0) bc = r; }
  __syncthreads();
  mx = bc;
  __syncthreads();
  if (lane == 0) red[w] = -mn;
  __syncthreads();
  if (w == 0) { float r = lane < nw ? red[lane] : -INFINITY; r = warp_max(r); if (lane == 0) bc = -r; }
  __syncthreads();
  mn = bc;
  __syncthreads();
  // Exact selection by bisection over the order-preserving integer image of the floats (32 passes build the
  // answer bit by bit): no sort, no scratch memory, ties at the boundary are kept like the sort-based reference.
  auto key_of = [](float x) -> uint32_t { const uint32_t u = __float_as_uint(x); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); };
  auto val_of = [](uint32_t k) -> float { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); };
  float cut = -INFINITY;  // keep x >= cut
  if (k > 0 && k < V) {
    // k-th largest key = max { c : count(key >= c) >= k }
    uint32_t prefix = 0;
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t cand = prefix | (1u << bit);
      float cnt = 0.f;
      for (int i = threadIdx.x; i < V; i += blockDim.x) cnt += key_of(lr[i]) >= cand ? 1.f : 0.f;
      cnt = block_sum(cnt, red);
      if (cnt >= (float)k) prefix = cand;
    }
    cut = val_of(prefix);
  }
  if (pth < 1.f || mp > 0.f) {
    float z = 0.f;
    for (int i = threadIdx.x; i < V; i += blockDim.x) { const float x = lr[i]; if (x >= cut) z += __expf(x - mx); }
    z = block_sum(z, red);
    if (mp > 0.f) cut = fmaxf(cut, mx + __logf(mp));  // p_i >= min_p * p_max
    if (pth < 1.f) {
      // nucleus: an element is dropped iff the probability mass of everything <= it is <= 1 - top_p.
      // smallest kept key = max { c : mass(key < c) <= (1 - top_p) * Z }
      const float thr = (1.f - pth) * z;
      const uint32_t kcut = key_of(cut);
      uint32_t prefix = 0;
      for (int bit = 31; bit >= 0; --bit) {
        const uint32_t cand = prefix | (1u << bit);
        float mass = 0.f;
        for (int i = threadIdx.x; i < V; i += blockDim.x) {
          const float x = lr[i];
          const uint32_t kx = key_of(x);
          if (kx >= kcut && kx < cand) mass += __expf(x - mx);
        }
        mass = block_sum(mass, red);
        if (mass <= thr) prefix = cand;
      }
      // never drop the arg-max (mass rounding): the largest element always survives
      cut = fmaxf(cut, fminf(val_of(prefix), mx));
    }
  }
  for (int i = threadIdx.x; i < V; i += blockDim.x)
    if (lr[i] < cut) lr[i] = -INFINITY;
}

// logits[b, v] adjusted by repetition / frequency / presence penalties given token histories.
__global__ void penalties_kernel(float* __restrict__ logits, int V, const int* __restrict__ prompt_mask_bits,
                                 const int* __restrict__ out_counts, const float* __restrict__ presence,
                                 const float* __restrict__ frequency, const float* __restrict__ repetition) {
  const int row = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V) return;
  const int cnt = out_counts[(int64_t)row * V + i];
  const bool in_prompt = prompt_mask_bits ? (prompt_mask_bits[(int64_t)row * ((V + 31) / 32) + (i >> 5)] >> (i & 31)) & 1 : false;
  float x = logits[(int64_t)row * V + i];
  if (in_prompt || cnt > 0) {
    const float rp = repetition[row];
    x = x > 0.f ? x / rp : x * rp;
  }
  x -= frequency[row] * (float)cnt;
  if (cnt > 0) x -= presence[row];
  logits[(int64_t)row * V + i] = x;
}

// per-row log-softmax statistics: out_lse[row] = logsumexp(logits[row])
template <typename T>
__global__ void __launch_bounds__(1024) row_lse_kernel(const T* __restrict__ logits, int64_t row_stride, int V,
                                                        float* __restrict__ out_lse) {
  __shared__ float red[32];
  __shared__ float bc;
  const int row = blockIdx.x;
  const T* lr = logits + (int64_t)row * row_stride;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < V; i += blockDim.x) mx = fmaxf(mx, to_f32(lr[i]));
  mx = warp_max(mx);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) red[w] = mx;
  __syncthreads();
  if (w == 0) { float r = lane < nw ? red[lane] : -INFINITY; r = warp_max(r); if (lane == 0) bc = r; }
  __syncthreads();
  mx = bc;
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < V; i += blockDim.x) s += __expf(to_f32(lr[i]) - mx);
  s = block_sum(s, red);
  if (threadIdx.x == 0) out_lse[row] = mx + __logf(s);
}

}  // namespace

EXPORT int vb_argmax(const void* logits, int64_t row_stride, int B, int V, int64_t* out, int dtype, cudaStream_t st) {
  if (B == 0) return 0;
  DISPATCH_DTYPE(dtype, TT, {
    sample_kernel<TT, false><<<B, 1024, 0, st>>>((const TT*)logits, row_stride, V, nullptr, nullptr, nullptr, out, 0, nullptr);
  });
  return (int)cudaGetLastError();
}

EXPORT int vb_fused_sample(const void* logits, int64_t row_stride, int B, int V, const float* temperature,
                           const int64_t* seeds, const int* step, int64_t* out, int dtype, cudaStream_t st) {
  if (B == 0) return 0;
  DISPATCH_DTYPE(dtype, TT, {
    vb_launch(sample_kernel<TT, true>, B, 1024, 0, st, (const TT*)logits, row_stride, V, temperature, seeds, step, out, 0, (float4*)nullptr);
  });
  return (int)cudaGetLastError();
}

// TP vocab shards: every rank samples over ITS shard (global ids = idx_off + local id, V = valid local rows) and emits one
// (value, id) candidate per row; after a 16-byte-per-row all-gather vb_select_candidate picks the winner — the 2 MiB of
// logits per rank that the full all-gather moved never leave the GPU.
EXPORT int vb_fused_sample_cand(const void* logits, int64_t row_stride, int B, int V, const float* temperature,
                                const int64_t* seeds, const int* step, int idx_off, void* cand, int dtype, cudaStream_t st) {
  if (B == 0) return 0;
  DISPATCH_DTYPE(dtype, TT, {
    vb_launch(sample_kernel<TT, true>, B, 1024, 0, st, (const TT*)logits, row_stride, V, temperature, seeds, step,
              (int64_t*)nullptr, idx_off, (float4*)cand);
  });
  return (int)cudaGetLastError();
}
EXPORT int vb_select_candidate(const void* cands, int W, int64_t* out, int B, cudaStream_t st) {
  if (B == 0) return 0;
  vb_launch(select_candidate_kernel, dim3((B + 127) / 128), dim3(128), 0, st, (const float4*)cands, W, out, B);
  return (int)cudaGetLastError();
}

EXPORT int vb_topk_topp(float* logits, int B, int V, const int* top_k, const float* top_p, const float* min_p,
                        cudaStream_t st) {
  if (B == 0) return 0;
  topk_topp_kernel<<<B, 1024, 0, st>>>(logits, V, top_k, top_p, min_p);
  return (int)cudaGetLastError();
}

EXPORT int vb_penalties(float* logits, int B, int V, const int* prompt_mask_bits, const int* out_counts,
                        const float* presence, const float* frequency, const float* repetition, cudaStream_t st) {
  if (B == 0) return 0;
  penalties_kernel<<<dim3((V + 255) / 256, B), 256, 0, st>>>(logits, V, prompt_mask_bits, out_counts, presence,
                                                              frequency, repetition);
  return (int)cudaGetLastError();
}

EXPORT int vb_row_lse(const void* logits, int64_t row_stride, int B, int V, float* out, int dtype, cudaStream_t st) {
  if (B == 0) return 0;
  DISPATCH_DTYPE(dtype, TT, {
    row_lse_kernel<TT><<<B, 1024, 0, st>>>((const TT*)logits, row_stride, V, out);
  });
  return (int)cudaGetLastError();
}
