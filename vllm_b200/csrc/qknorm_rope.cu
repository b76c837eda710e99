// Fused per-head RMSNorm(q), RMSNorm(k) + RoPE + paged KV-cache insert (Qwen3 / Gemma-3 / OLMo-2 / Llama-4 style
// attention prologues).
//
// Reference ops: csrc/libtorch_stable/fused_qknorm_rope_kernel.cu:563-686 (q/k norm + rope over the fused qkv
// buffer) and the model-specific q-norm / rope / KV-insert fusions listed in SURVEY §2.3.  Ours goes one step
// further than the generic reference op: the rotated K head is written to its cache page from the registers that
// produced it and the V head is copied by a sibling warp of the same CTA, so a layer's attention prologue is ONE
// launch instead of norm(q), norm(k), rope, reshape_and_cache.
//
// Mapping: one warp per (token, head).  The head lives in registers (E = D / 32 consecutive elements per lane);
// sum of squares by warp shuffle; NeoX partner element (i, i + rot/2) sits in lane ^ (rot / 2 / E) at the same
// register slot, GPT-J pairs (2i, 2i+1) inside the lane.
#include "common.cuh"

namespace {

template <typename T, int E> struct LaneVec {
  alignas(16) T e[E];
  VB_DEVICE void load(const T* p) {
    if constexpr (E * sizeof(T) == 32) {
      reinterpret_cast<uint4*>(e)[0] = reinterpret_cast<const uint4*>(p)[0];
      reinterpret_cast<uint4*>(e)[1] = reinterpret_cast<const uint4*>(p)[1];
    } else if constexpr (E * sizeof(T) == 16) { *reinterpret_cast<uint4*>(e) = *reinterpret_cast<const uint4*>(p); }
    else if constexpr (E * sizeof(T) == 8) { *reinterpret_cast<uint2*>(e) = *reinterpret_cast<const uint2*>(p); }
    else if constexpr (E * sizeof(T) == 4) { *reinterpret_cast<uint32_t*>(e) = *reinterpret_cast<const uint32_t*>(p); }
    else {
#pragma unroll
      for (int j = 0; j < E; ++j) e[j] = p[j];
    }
  }
  VB_DEVICE void store(T* p) const {
    if constexpr (E * sizeof(T) == 32) {
      reinterpret_cast<uint4*>(p)[0] = reinterpret_cast<const uint4*>(e)[0];
      reinterpret_cast<uint4*>(p)[1] = reinterpret_cast<const uint4*>(e)[1];
    } else if constexpr (E * sizeof(T) == 16) { *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(e); }
    else if constexpr (E * sizeof(T) == 8) { *reinterpret_cast<uint2*>(p) = *reinterpret_cast<const uint2*>(e); }
    else if constexpr (E * sizeof(T) == 4) { *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<const uint32_t*>(e); }
    else {
#pragma unroll
      for (int j = 0; j < E; ++j) p[j] = e[j];
    }
  }
};

struct Params {
  const int64_t* positions;
  void* q; void* k; const void* v;
  const void* q_w; const void* k_w;        // [D] norm weights (null = no norm on that side)
  const float* cos_sin;                    // [max_pos, rot] = cos | sin
  void* kv_cache; const int64_t* slot_mapping;
  int64_t q_stride, k_stride, v_stride;
  int T, Hq, Hkv, D, rot, BS;
  int neox, gemma, kv_fp8, vt;
  float eps, k_inv_scale, v_inv_scale;
};

template <typename T, int E>
__global__ void __launch_bounds__(256) qknorm_rope_cache_kernel(const Params p) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int items = p.Hq + 2 * p.Hkv;
  const int64_t w_global = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5);
  if (w_global >= (int64_t)p.T * items) return;
  const int t = (int)(w_global / items), it = (int)(w_global % items);
  const int D = p.D;
  const int64_t slot = p.slot_mapping ? p.slot_mapping[t] : -1;
  const bool do_cache = p.kv_cache != nullptr && slot >= 0;
  const int64_t blk = do_cache ? slot / p.BS : 0;
  const int off = do_cache ? (int)(slot % p.BS) : 0;

  if (it >= p.Hq + p.Hkv) {
    // ---- V head → cache page
    if (!do_cache || p.v == nullptr) return;
    const int h = it - p.Hq - p.Hkv;
    LaneVec<T, E> a;
    a.load(reinterpret_cast<const T*>(p.v) + t * p.v_stride + (int64_t)h * D + lane * E);
    const size_t page = (((size_t)blk * 2 + 1) * p.Hkv + h) * p.BS * D;
    if (p.vt) {            // tensor-core layout: V pages transposed [D, BS]
      T* vc = reinterpret_cast<T*>(p.kv_cache) + page + off;
#pragma unroll
      for (int j = 0; j < E; ++j) vc[(size_t)(lane * E + j) * p.BS] = a.e[j];
    } else if (p.kv_fp8) {
      __nv_fp8_storage_t* vc = reinterpret_cast<__nv_fp8_storage_t*>(p.kv_cache) + page + (size_t)off * D + lane * E;
#pragma unroll
      for (int j = 0; j < E; ++j) vc[j] = __nv_cvt_float_to_fp8(to_f32(a.e[j]) * p.v_inv_scale, __NV_SATFINITE, __NV_E4M3);
    } else {
      a.store(reinterpret_cast<T*>(p.kv_cache) + page + (size_t)off * D + lane * E);
    }
    return;
  }
  // ---- q / k head: norm → rope → (k) cache
  const bool is_k = it >= p.Hq;
  const int h = is_k ? it - p.Hq : it;
  T* base = is_k ? reinterpret_cast<T*>(p.k) + t * p.k_stride + (int64_t)h * D
                 : reinterpret_cast<T*>(p.q) + t * p.q_stride + (int64_t)h * D;
  LaneVec<T, E> a;
  a.load(base + lane * E);
  float f[E];
#pragma unroll
  for (int j = 0; j < E; ++j) f[j] = to_f32(a.e[j]);
  const T* w = reinterpret_cast<const T*>(is_k ? p.k_w : p.q_w);
  if (w != nullptr) {
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < E; ++j) ss += f[j] * f[j];
    ss = warp_sum(ss);
    const float inv = rsqrtf(ss / (float)D + p.eps);
    LaneVec<T, E> wv;
    wv.load(w + lane * E);
#pragma unroll
    for (int j = 0; j < E; ++j) {
      float g = to_f32(wv.e[j]);
      if (p.gemma) g += 1.f;
      f[j] = to_f32(from_f32<T>(f[j] * inv * g));     // round through T like the unfused norm does
    }
  }
  if (p.rot > 0) {
    const int half = p.rot / 2;
    const float* cs = p.cos_sin + p.positions[t] * p.rot;
    if (p.neox) {
      const int span = half / E;                       // lanes per half (power of two, checked on the host)
      const bool rot_lane = lane < 2 * span;
      const bool first = (lane & span) == 0;
#pragma unroll
      for (int j = 0; j < E; ++j) {
        const float other = __shfl_xor_sync(0xffffffffu, f[j], span);
        if (rot_lane) {
          const int i = (lane & (span - 1)) * E + j;   // index inside the half
          const float co = cs[i], si = cs[half + i];
          f[j] = first ? f[j] * co - other * si : f[j] * co + other * si;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < E; j += 2) {
        const int e0 = lane * E + j;
        if (e0 < p.rot) {
          const int i = e0 >> 1;
          const float co = cs[i], si = cs[half + i];
          const float x1 = f[j], x2 = f[j + 1];
          f[j] = x1 * co - x2 * si;
          f[j + 1] = x2 * co + x1 * si;
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < E; ++j) a.e[j] = from_f32<T>(f[j]);
  a.store(base + lane * E);
  if (is_k && do_cache) {
    const size_t page = (((size_t)blk * 2 + 0) * p.Hkv + h) * p.BS * D + (size_t)off * D + lane * E;
    if (p.kv_fp8) {
      __nv_fp8_storage_t* kc = reinterpret_cast<__nv_fp8_storage_t*>(p.kv_cache) + page;
#pragma unroll
      for (int j = 0; j < E; ++j) kc[j] = __nv_cvt_float_to_fp8(to_f32(a.e[j]) * p.k_inv_scale, __NV_SATFINITE, __NV_E4M3);
    } else {
      a.store(reinterpret_cast<T*>(p.kv_cache) + page);
    }
  }
}

}  // namespace

// q, k: [T, H*D] views of the fused qkv buffer (rotated / normalised in place); v may be null (no cache write of V);
// kv_cache may be null (norm + rope only).  flags: bit0 fp8 cache, bit1 transposed V pages.
EXPORT int vb_qknorm_rope_cache(const int64_t* positions, void* q, void* k, const void* v, const void* q_w, const void* k_w,
                                const float* cos_sin, int rot, int T, int Hq, int Hkv, int D, int64_t q_stride,
                                int64_t k_stride, int64_t v_stride, void* kv_cache, const int64_t* slot_mapping, int BS,
                                int neox, int gemma, int flags, float eps, float k_scale, float v_scale, int dtype,
                                cudaStream_t st) {
  if (T == 0) return 0;
  if (dtype == DT_F32 || D % 32 != 0 || D > 512) return -1;
  const int E = D / 32;
  if (E != 2 && E != 4 && E != 8 && E != 16) return -1;
  if (rot > 0) {
    if (rot > D || rot % 2) return -1;
    if (neox) {
      const int half = rot / 2;
      if (half % E) return -1;
      const int span = half / E;
      if (span & (span - 1)) return -1;            // partner lane = lane ^ span
    } else if (E % 2) return -1;
  }
  Params p;
  p.positions = positions; p.q = q; p.k = k; p.v = v; p.q_w = q_w; p.k_w = k_w; p.cos_sin = cos_sin;
  p.kv_cache = kv_cache; p.slot_mapping = slot_mapping; p.q_stride = q_stride; p.k_stride = k_stride; p.v_stride = v_stride;
  p.T = T; p.Hq = Hq; p.Hkv = Hkv; p.D = D; p.rot = rot; p.BS = BS; p.neox = neox; p.gemma = gemma;
  p.kv_fp8 = flags & 1; p.vt = (flags >> 1) & 1; p.eps = eps; p.k_inv_scale = 1.f / k_scale; p.v_inv_scale = 1.f / v_scale;
  const int64_t warps = (int64_t)T * (Hq + 2 * Hkv);
  const int wpb = 8;
  const unsigned blocks = (unsigned)((warps + wpb - 1) / wpb);
  DISPATCH_HALF(dtype, TT, {
    switch (E) {
      case 2: vb_launch(qknorm_rope_cache_kernel<TT, 2>, dim3(blocks), dim3(wpb * 32), 0, st, p); break;
      case 4: vb_launch(qknorm_rope_cache_kernel<TT, 4>, dim3(blocks), dim3(wpb * 32), 0, st, p); break;
      case 8: vb_launch(qknorm_rope_cache_kernel<TT, 8>, dim3(blocks), dim3(wpb * 32), 0, st, p); break;
      default: vb_launch(qknorm_rope_cache_kernel<TT, 16>, dim3(blocks), dim3(wpb * 32), 0, st, p); break;
    }
  });
  return (int)cudaGetLastError();
}
