// Mixture-of-experts routing / dispatch / combine kernels (decode-sized batches stay entirely on the device:
// no host sync, CUDA-graph capturable).  The expert GEMMs themselves are the grouped mode of gemm_tc.cu.
//
// Reference counterparts: csrc/moe/topk_softmax_kernels.cu, moe_align_sum_kernels.cu, moe_permute_unpermute_op.cu
// (here: routing probabilities + top-k in one warp per token; counting sort of (token, slot) pairs by expert with
// an inverse map; weighted un-permute that sums a token's k expert outputs in a fixed order → deterministic).
#include "common.cuh"

namespace moe {

constexpr int MAX_E_PER_LANE = 16;   // experts <= 512

// one warp per token: softmax over E logits, pick the k largest, optionally renormalise
template <typename T>
__global__ void topk_softmax_kernel(const T* __restrict__ logits, float* __restrict__ out_w, int* __restrict__ out_ids,
                                    int n_tokens, int E, int k, int renorm, float scale) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_tokens) return;
  const T* row = logits + (size_t)warp * E;
  float v[MAX_E_PER_LANE];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < MAX_E_PER_LANE; ++i) {
    const int e = lane + i * 32;
    v[i] = e < E ? to_f32(row[e]) : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_E_PER_LANE; ++i) {
    const int e = lane + i * 32;
    v[i] = e < E ? __expf(v[i] - mx) : 0.f;
    sum += v[i];
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  float picked_sum = 0.f;
  float my_w = 0.f; int my_id = 0;
  for (int j = 0; j < k; ++j) {
    float best = -1.f; int bi = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < MAX_E_PER_LANE; ++i) {
      const int e = lane + i * 32;
      if (e < E && (v[i] > best)) { best = v[i]; bi = e; }     // lowest index wins ties inside a lane
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if ((bi & 31) == lane) v[bi >> 5] = -1.f;                   // remove the winner
    if (lane == j) { my_w = best * inv; my_id = bi; }
    picked_sum += best * inv;
  }
  if (lane < k) {
    out_w[(size_t)warp * k + lane] = (renorm ? my_w / picked_sum : my_w) * scale;
    out_ids[(size_t)warp * k + lane] = my_id;
  }
}

// counting sort of R = T*k (token, slot) pairs by (mapped) expert id.  One CTA.
//   ids[i]  : global expert of pair i;  emap (optional): global → local expert or -1 (not on this rank)
//   sorted_tok[p] : token of the pair placed at sorted position p;  inv[i] : sorted position of pair i (or -1)
//   offs[e] : first sorted position of local expert e;  offs[E] = number of valid pairs
__global__ void align_kernel(const int* __restrict__ ids, const int* __restrict__ emap, int* __restrict__ sorted_tok,
                             int* __restrict__ inv, int* __restrict__ offs, int R, int k, int E,
                             int* __restrict__ sorted_pair) {
  extern __shared__ int sm[];
  int* cnt = sm;            // [E]
  int* start = sm + E;      // [E + 1]
  pdl_trigger();
  pdl_wait();
  for (int e = threadIdx.x; e < E; e += blockDim.x) cnt[e] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < R; i += blockDim.x) {
    int e = ids[i];
    if (emap) e = emap[e];
    if (e >= 0) atomicAdd(&cnt[e], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int e = 0; e < E; ++e) { start[e] = acc; offs[e] = acc; acc += cnt[e]; cnt[e] = 0; }
    start[E] = acc; offs[E] = acc;
  }
  __syncthreads();
  // placement: the order of pairs inside one expert's segment is arbitrary (atomic cursor) — every row of the
  // grouped GEMM is computed independently and the combine step goes through `inv`, so results do not depend on it
  for (int i = threadIdx.x; i < R; i += blockDim.x) {
    int e = ids[i];
    if (emap) e = emap[e];
    if (e < 0) { inv[i] = -1; continue; }
    const int p = start[e] + atomicAdd(&cnt[e], 1);
    sorted_tok[p] = i / k;
    if (sorted_pair) sorted_pair[p] = i;
    inv[i] = p;
  }
  // sorted positions past the valid pairs (padding rows of the activation buffer) point at token 0
  for (int p = start[E] + threadIdx.x; p < R; p += blockDim.x) { sorted_tok[p] = 0; if (sorted_pair) sorted_pair[p] = 0; }
}

// Tail of the fused MoE combine: acc (fp32, every rank holds the complete sum after the multimem.red epilogue of the
// w2 grouped GEMM) + residual → new residual, RMSNorm → out; acc is zeroed for its next use.  One CTA per token.
template <typename T, bool GEMMA>
__global__ void __launch_bounds__(1024) moe_finish_kernel(float* __restrict__ acc, T* __restrict__ residual,
                                                           const T* __restrict__ w, T* __restrict__ out, int H, float eps) {
  __shared__ float red[32];
  pdl_trigger();
  pdl_wait();
  const int64_t row = blockIdx.x;
  float* a = acc + row * H;
  T* r = residual ? residual + row * H : nullptr;
  constexpr int MAXV4 = 8;                     // float4 per thread kept in registers (H <= 1024 * 8 * 4)
  float4 v[MAXV4];
  float ss = 0.f;
  const int nv = H / 4;
#pragma unroll
  for (int i = 0; i < MAXV4; ++i) {
    const int vi = threadIdx.x + i * blockDim.x;
    if (vi < nv) {
      v[i] = *reinterpret_cast<const float4*>(a + vi * 4);
      *reinterpret_cast<float4*>(a + vi * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r) {
        const uint2 rv = *reinterpret_cast<const uint2*>(r + vi * 4);
        const float2 r0 = unpack2<T>(rv.x), r1 = unpack2<T>(rv.y);
        // the stream is kept in T: round the sum through T like the unfused path
        v[i].x = to_f32(from_f32<T>(v[i].x)) + r0.x; v[i].y = to_f32(from_f32<T>(v[i].y)) + r0.y;
        v[i].z = to_f32(from_f32<T>(v[i].z)) + r1.x; v[i].w = to_f32(from_f32<T>(v[i].w)) + r1.y;
        uint2 o;
        o.x = pack2<T>(v[i].x, v[i].y); o.y = pack2<T>(v[i].z, v[i].w);
        *reinterpret_cast<uint2*>(r + vi * 4) = o;
        const float2 q0 = unpack2<T>(o.x), q1 = unpack2<T>(o.y);
        v[i] = make_float4(q0.x, q0.y, q1.x, q1.y);
      }
      ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
  }
  if (w == nullptr) {                          // no norm: out = combined (+ residual) in T
#pragma unroll
    for (int i = 0; i < MAXV4; ++i) {
      const int vi = threadIdx.x + i * blockDim.x;
      if (vi < nv) {
        uint2 o;
        o.x = pack2<T>(v[i].x, v[i].y); o.y = pack2<T>(v[i].z, v[i].w);
        *reinterpret_cast<uint2*>(out + row * H + vi * 4) = o;
      }
    }
    return;
  }
  ss = block_sum(ss, red);
  const float inv = rsqrtf(ss / (float)H + eps);
#pragma unroll
  for (int i = 0; i < MAXV4; ++i) {
    const int vi = threadIdx.x + i * blockDim.x;
    if (vi < nv) {
      const uint2 wv = *reinterpret_cast<const uint2*>(w + vi * 4);
      float2 w0 = unpack2<T>(wv.x), w1 = unpack2<T>(wv.y);
      if (GEMMA) { w0.x += 1.f; w0.y += 1.f; w1.x += 1.f; w1.y += 1.f; }
      uint2 o;
      o.x = pack2<T>(v[i].x * inv * w0.x, v[i].y * inv * w0.y);
      o.y = pack2<T>(v[i].z * inv * w1.x, v[i].w * inv * w1.y);
      *reinterpret_cast<uint2*>(out + row * H + vi * 4) = o;
    }
  }
}

// out[t, :] = sum_j w[t, j] * y[inv[t*k + j], :]   (pairs not computed on this rank contribute 0)
template <typename T>
__global__ void combine_kernel(const T* __restrict__ y, const int* __restrict__ inv, const float* __restrict__ w,
                               T* __restrict__ out, int H, int k) {
  pdl_trigger();
  pdl_wait();
  const int t = blockIdx.x;
  for (int g = threadIdx.x; g < H / 8; g += blockDim.x) {
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int j = 0; j < k; ++j) {
      const int p = inv[t * k + j];
      if (p < 0) continue;
      const float wj = w[t * k + j];
      const Vec16<T> v = ld16(y + (size_t)p * H + g * 8);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += wj * to_f32(v.e[e]);
    }
    Vec16<T> o;
#pragma unroll
    for (int e = 0; e < 8; ++e) o.e[e] = from_f32<T>(acc[e]);
    st16(out + (size_t)t * H + g * 8, o);
  }
}

// xs[p, :] = x[sorted_tok[p], :]
template <typename T>
__global__ void gather_kernel(const T* __restrict__ x, const int* __restrict__ sorted_tok, T* __restrict__ xs, int H) {
  pdl_trigger();
  pdl_wait();
  const int p = blockIdx.x;
  const T* src = x + (size_t)sorted_tok[p] * H;
  for (int g = threadIdx.x; g < H / 8; g += blockDim.x) st16(xs + (size_t)p * H + g * 8, ld16(src + g * 8));
}

}  // namespace moe

EXPORT int vb_moe_topk_softmax(const void* logits, float* out_w, int* out_ids, int n_tokens, int E, int k, int renorm,
                               float scale, int dtype, cudaStream_t st) {
  if (E > 32 * moe::MAX_E_PER_LANE || k > 32) return -1;
  const int threads = 128, warps_per_block = threads / 32;
  dim3 grid((n_tokens + warps_per_block - 1) / warps_per_block);
  DISPATCH_DTYPE(dtype, T, {
    vb_launch(moe::topk_softmax_kernel<T>, grid, dim3(threads), 0, st, (const T*)logits, out_w, out_ids, n_tokens, E, k,
              renorm, scale);
  });
  return (int)cudaGetLastError();
}

EXPORT int vb_moe_align2(const int* ids, const int* emap, int* sorted_tok, int* inv, int* offs, int R, int k, int E,
                         int* sorted_pair, cudaStream_t st) {
  if (E > 4096) return -1;
  vb_launch(moe::align_kernel, dim3(1), dim3(256), (size_t)(2 * E + 1) * sizeof(int), st, ids, emap, sorted_tok, inv, offs,
            R, k, E, sorted_pair);
  return (int)cudaGetLastError();
}
EXPORT int vb_moe_align(const int* ids, const int* emap, int* sorted_tok, int* inv, int* offs, int R, int k, int E,
                        cudaStream_t st) {
  return vb_moe_align2(ids, emap, sorted_tok, inv, offs, R, k, E, nullptr, st);
}

// out = RMSNorm(acc + residual) (norm_w may be null: out = acc (+ residual)); residual updated in place; acc zeroed
EXPORT int vb_moe_finish(float* acc, void* residual, const void* norm_w, void* out, int T_, int H, float eps, int gemma,
                         int dtype, cudaStream_t st) {
  if (T_ == 0) return 0;
  if (dtype == DT_F32 || H % 4 || H > 1024 * 8 * 4) return -1;
  int thr = ((H / 4 + 7) / 8 + 31) / 32 * 32;
  thr = thr < 64 ? 64 : (thr > 1024 ? 1024 : thr);
  DISPATCH_HALF(dtype, T, {
    if (gemma) vb_launch(moe::moe_finish_kernel<T, true>, dim3(T_), dim3(thr), 0, st, acc, (T*)residual, (const T*)norm_w, (T*)out, H, eps);
    else vb_launch(moe::moe_finish_kernel<T, false>, dim3(T_), dim3(thr), 0, st, acc, (T*)residual, (const T*)norm_w, (T*)out, H, eps);
  });
  return (int)cudaGetLastError();
}

EXPORT int vb_moe_combine(const void* y, const int* inv, const float* w, void* out, int T_, int H, int k, int dtype,
                          cudaStream_t st) {
  if (H % 8) return -1;
  DISPATCH_HALF(dtype, T, {
    vb_launch(moe::combine_kernel<T>, dim3(T_), dim3(128), 0, st, (const T*)y, inv, w, (T*)out, H, k);
  });
  return (int)cudaGetLastError();
}

EXPORT int vb_moe_gather(const void* x, const int* sorted_tok, void* xs, int R, int H, int dtype, cudaStream_t st) {
  if (H % 8) return -1;
  DISPATCH_HALF(dtype, T, {
    vb_launch(moe::gather_kernel<T>, dim3(R), dim3(128), 0, st, (const T*)x, sorted_tok, (T*)xs, H);
  });
  return (int)cudaGetLastError();
}
