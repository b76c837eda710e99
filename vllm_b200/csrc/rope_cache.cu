// Fused RoPE(q,k) + paged KV-cache write, stand-alone RoPE, stand-alone cache write, MLA latent write.
//
// Reference ops: pos_encoding_kernels.cu (rotary_embedding), cache_kernels.cu
// (reshape_and_cache_flash :320-406, concat_and_cache_mla).  Ours fuses the three
// launches the reference issues per layer (rope, cache write; plus the q/k/v split
// copies) into one: q and k are rotated in place inside the fused qkv buffer and the
// rotated k / raw v go straight to the page.
//
// Cache layout: [num_blocks, 2, Hkv, block_size, D] — a (page, head) K or V tile is one
// contiguous block_size*D run, which is what the attention kernels bulk-copy.
#include "common.cuh"

namespace {

template <typename T> struct Fp8Store {
  static VB_DEVICE void store8(__nv_fp8_storage_t* dst, const Vec16<T>& v, float inv_scale) {
    union { uint2 u; __nv_fp8_storage_t b[8]; } o;
#pragma unroll
    for (int j = 0; j < Vec16<T>::N && j < 8; ++j)
      o.b[j] = __nv_cvt_float_to_fp8(to_f32(v.e[j]) * inv_scale, __NV_SATFINITE, __NV_E4M3);
    *reinterpret_cast<uint2*>(dst) = o.u;
  }
};

// One CTA per token. Work items: (head, 8-element chunk).
template <typename T, bool NEOX, bool FP8>
__global__ void rope_cache_kernel(const int64_t* __restrict__ positions, T* __restrict__ q, T* __restrict__ k,
                                  const T* __restrict__ v, const float* __restrict__ cos_sin, int rot, int Hq,
                                  int Hkv, int D, int64_t q_stride, int64_t k_stride, int64_t v_stride,
                                  void* __restrict__ kv_cache, const int64_t* __restrict__ slot_mapping, int BS,
                                  float k_inv_scale, float v_inv_scale, int vt) {
  constexpr int N = 8;  // 16-bit dtypes only
  pdl_trigger();
  pdl_wait();
  const int t = blockIdx.x;
  const int64_t slot = slot_mapping ? slot_mapping[t] : -1;
  const bool do_cache = kv_cache != nullptr && slot >= 0;
  const int64_t blk = do_cache ? slot / BS : 0;
  const int off = do_cache ? (int)(slot % BS) : 0;
  const int half = rot / 2;
  const float* cs = (rot > 0) ? cos_sin + positions[t] * rot : nullptr;
  const int chunks_rot = half / N;            // chunk pairs inside the rotary part
  const int chunks_pass = (D - rot) / N;      // untouched tail (partial rotary)
  const int per_head = chunks_rot + chunks_pass;
  const int n_qk = (Hq + Hkv) * per_head;
  const int n_v = (v != nullptr) ? Hkv * (D / N) : 0;
  for (int i = threadIdx.x; i < n_qk + n_v; i += blockDim.x) {
    if (i < n_qk) {
      const int h = i / per_head, c = i % per_head;
      const bool is_k = h >= Hq;
      if (!is_k && q == nullptr) continue;
      T* base = is_k ? k + t * k_stride + (int64_t)(h - Hq) * D : q + t * q_stride + (int64_t)h * D;
      size_t cache_off = 0;
      if (is_k && do_cache) cache_off = (((size_t)blk * 2 + 0) * Hkv + (h - Hq)) * BS * D + (size_t)off * D;
      if (c < chunks_rot) {
        Vec16<T> a, b, oa, ob;
        if (NEOX) {
          a = ld16(base + c * N);
          b = ld16(base + half + c * N);
#pragma unroll
          for (int j = 0; j < N; ++j) {
            const float co = cs[c * N + j], si = cs[half + c * N + j];
            const float x1 = to_f32(a.e[j]), x2 = to_f32(b.e[j]);
            oa.e[j] = from_f32<T>(x1 * co - x2 * si);
            ob.e[j] = from_f32<T>(x2 * co + x1 * si);
          }
          st16(base + c * N, oa);
          st16(base + half + c * N, ob);
          if (is_k && do_cache) {
            if (FP8) {
              auto* kc = (__nv_fp8_storage_t*)kv_cache + cache_off;
              Fp8Store<T>::store8(kc + c * N, oa, k_inv_scale);
              Fp8Store<T>::store8(kc + half + c * N, ob, k_inv_scale);
            } else {
              T* kc = (T*)kv_cache + cache_off;
              st16(kc + c * N, oa);
              st16(kc + half + c * N, ob);
            }
          }
        } else {  // GPT-J interleaved pairs (2i, 2i+1): two adjacent vectors cover 8 pairs
          a = ld16(base + 2 * c * N);
          b = ld16(base + 2 * c * N + N);
#pragma unroll
          for (int j = 0; j < N; j += 2) {
            const int p0 = c * N + j / 2, p1 = c * N + N / 2 + j / 2;
            float co = cs[p0], si = cs[half + p0];
            float x1 = to_f32(a.e[j]), x2 = to_f32(a.e[j + 1]);
            oa.e[j] = from_f32<T>(x1 * co - x2 * si);
            oa.e[j + 1] = from_f32<T>(x2 * co + x1 * si);
            co = cs[p1]; si = cs[half + p1];
            x1 = to_f32(b.e[j]); x2 = to_f32(b.e[j + 1]);
            ob.e[j] = from_f32<T>(x1 * co - x2 * si);
            ob.e[j + 1] = from_f32<T>(x2 * co + x1 * si);
          }
          st16(base + 2 * c * N, oa);
          st16(base + 2 * c * N + N, ob);
          if (is_k && do_cache) {
            if (FP8) {
              auto* kc = (__nv_fp8_storage_t*)kv_cache + cache_off;
              Fp8Store<T>::store8(kc + 2 * c * N, oa, k_inv_scale);
              Fp8Store<T>::store8(kc + 2 * c * N + N, ob, k_inv_scale);
            } else {
              T* kc = (T*)kv_cache + cache_off;
              st16(kc + 2 * c * N, oa);
              st16(kc + 2 * c * N + N, ob);
            }
          }
        }
      } else if (is_k && do_cache) {  // pass-through tail of k
        const int e = rot + (c - chunks_rot) * N;
        Vec16<T> a = ld16(base + e);
        if (FP8) Fp8Store<T>::store8((__nv_fp8_storage_t*)kv_cache + cache_off + e, a, k_inv_scale);
        else st16((T*)kv_cache + cache_off + e, a);
      }
    } else if (do_cache) {
      const int j = i - n_qk;
      const int h = j / (D / N), c = j % (D / N);
      Vec16<T> a = ld16(v + t * v_stride + (int64_t)h * D + c * N);
      if (vt) {
        // tensor-core layout: the V page is stored transposed [D, BS] (tokens contiguous) so that it is a
        // K-major B operand for O = P·V (see attention_tc.cu)
        T* vc = (T*)kv_cache + (((size_t)blk * 2 + 1) * Hkv + h) * BS * D + off;
#pragma unroll
        for (int e = 0; e < N; ++e) vc[(size_t)(c * N + e) * BS] = a.e[e];
      } else {
        const size_t cache_off = (((size_t)blk * 2 + 1) * Hkv + h) * BS * D + (size_t)off * D + c * N;
        if (FP8) Fp8Store<T>::store8((__nv_fp8_storage_t*)kv_cache + cache_off, a, v_inv_scale);
        else st16((T*)kv_cache + cache_off, a);
      }
    }
  }
}

// MLA: cache[blk, off, :R] = kv_c[t], cache[blk, off, R:R+P] = k_pe[t]
template <typename T>
__global__ void concat_cache_mla_kernel(const T* __restrict__ kv_c, const T* __restrict__ k_pe,
                                        T* __restrict__ cache, const int64_t* __restrict__ slot_mapping, int R, int P,
                                        int64_t kvc_stride, int64_t kpe_stride, int BS) {
  constexpr int N = Vec16<T>::N;
  const int t = blockIdx.x;
  const int64_t slot = slot_mapping[t];
  if (slot < 0) return;
  T* dst = cache + slot * (int64_t)(R + P);
  for (int i = threadIdx.x; i < (R + P) / N; i += blockDim.x) {
    const int e = i * N;
    Vec16<T> a = e < R ? ld16(kv_c + t * kvc_stride + e) : ld16(k_pe + t * kpe_stride + (e - R));
    st16(dst + e, a);
  }
}

}  // namespace

// q may be null (cache write only, rot = 0); kv_cache may be null (rope only).
EXPORT int vb_rope_and_cache(const int64_t* positions, void* q, void* k, const void* v, const float* cos_sin, int rot,
                             int T, int Hq, int Hkv, int D, int64_t q_stride, int64_t k_stride, int64_t v_stride,
                             void* kv_cache, const int64_t* slot_mapping, int BS, int neox, int kv_fp8,
                             float k_scale, float v_scale, int dtype, cudaStream_t st) {
  if (T == 0) return 0;
  if (dtype == DT_F32 || D % 8 != 0 || (rot / 2) % 8 != 0 || (D - rot) % 8 != 0) return -1;
  const int items = (Hq + Hkv) * ((rot / 2 + (D - rot)) / 8) + Hkv * (D / 8);
  int thr = ((items + 31) / 32) * 32;
  thr = thr > 512 ? 512 : (thr < 64 ? 64 : thr);
  if (q == nullptr) Hq = 0;
  const int vt = (kv_fp8 >> 1) & 1;   // bit 1 of the flags: V pages stored transposed
  kv_fp8 &= 1;
#define VB_ROPE_LAUNCH(NX, F8)                                                                                    \
  vb_launch(rope_cache_kernel<TT, NX, F8>, T, thr, 0, st, positions, (TT*)q, (TT*)k, (const TT*)v, cos_sin, rot, Hq, Hkv, \
                                                   D, q_stride, k_stride, v_stride, kv_cache, slot_mapping, BS,     \
                                                   1.f / k_scale, 1.f / v_scale, vt)
  DISPATCH_HALF(dtype, TT, {
    if (neox) { if (kv_fp8) VB_ROPE_LAUNCH(true, true); else VB_ROPE_LAUNCH(true, false); }
    else { if (kv_fp8) VB_ROPE_LAUNCH(false, true); else VB_ROPE_LAUNCH(false, false); }
  });
#undef VB_ROPE_LAUNCH
  return (int)cudaGetLastError();
}

EXPORT int vb_concat_and_cache_mla(const void* kv_c, const void* k_pe, void* cache, const int64_t* slot_mapping, int T,
                                   int R, int P, int64_t kvc_stride, int64_t kpe_stride, int BS, int dtype,
                                   cudaStream_t st) {
  if (T == 0) return 0;
  if (dtype == DT_F32 || R % 8 || P % 8) return -1;
  DISPATCH_HALF(dtype, TT, {
    concat_cache_mla_kernel<TT><<<T, 128, 0, st>>>((const TT*)kv_c, (const TT*)k_pe, (TT*)cache, slot_mapping, R, P,
                                                   kvc_stride, kpe_stride, BS);
  });
  return (int)cudaGetLastError();
}
