// Activation quantisation kernels and their norm / activation fusions (sm_100a).
//
// Reference ops being covered (none of the code is shared): csrc/libtorch_stable/quantization/w8a8/fp8/common.cu
// (static / dynamic per-tensor / per-token scaled_fp8_quant), w8a8/fp8/per_token_group_quant.cu (1x128 / 1x64 /
// 1x32 groups, column-major + UE8M0 scale options for DeepGEMM-style consumers), w8a8/int8/scaled_quant.cu,
// layernorm_quant_kernels.cu + quantization/fused_kernels/ (RMSNorm [+ residual] → fp8 static / per-token / per-group),
// activation quant fusions (silu_and_mul → fp8 / per-group fp8), quantization/fp4/nvfp4_quant_kernels.cu.
//
// Design: ONE kernel template.  One CTA per token; the row is read from HBM exactly once (16-byte vectors) and
// stays in registers as fp32 through  prologue (none | RMSNorm(+residual) | act(gate)·up)  →  amax reduction at
// the scale granularity (warp-shuffle inside a group, block reduction for per-token)  →  convert + 8-byte stores.
// The producer epilogues of the tcgen05 GEMMs consume the same scale layouts (gemm_fp8.cu).
#include "common.cuh"

namespace quant {

constexpr float FP8_MAX = 448.f;
constexpr int MAXV = 4;          // 8-element vectors per thread (blockDim 1024 → rows up to 32768)

enum : int { PRE_NONE = 0, PRE_RMS = 1, PRE_ACT_MUL = 2 };
enum : int { Q_FP8_STATIC = 0, Q_FP8_TOKEN = 1, Q_FP8_GROUP = 2, Q_INT8_TOKEN = 3, Q_INT8_STATIC = 4 };

struct Params {
  const void* x;          // [T, H] (PRE_ACT_MUL: [T, 2H] = gate | up)
  void* residual;         // PRE_RMS: optional [T, H], updated in place (residual += x)
  const void* weight;     // PRE_RMS: [H]
  void* q;                // [T, H] fp8 e4m3 / int8
  float* scale_out;       // per-token [T]; per-group [T, H/G] (row major) or [H/G, T_pad] (column major)
  const float* scale_in;  // static: [1]
  int64_t x_stride;
  int H, group;
  int scale_stride;       // column-major group scales: elements between consecutive groups (>= T)
  int col_major, ue8m0, gemma, act;
  float eps;
  float scale_ub;         // optional upper bound on dynamic per-token scales (0 = none)
};

VB_DEVICE float silu(float x) { return x / (1.f + __expf(-x)); }
VB_DEVICE float gelu_tanh(float x) { return 0.5f * x * (1.f + tanhf(0.7978845608028654f * (x + 0.044715f * x * x * x))); }
VB_DEVICE float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

VB_DEVICE uint32_t pack_fp8x4(float a, float b, float c, float d) {
  uint16_t lo, hi;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %2, %1;" : "=h"(lo) : "f"(a), "f"(b));
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %2, %1;" : "=h"(hi) : "f"(c), "f"(d));
  return (uint32_t)lo | ((uint32_t)hi << 16);
}
VB_DEVICE uint32_t pack_i8x4(float a, float b, float c, float d) {
  auto cv = [](float v) { int r = __float2int_rn(v); return (uint32_t)(max(-128, min(127, r)) & 0xff); };
  return cv(a) | (cv(b) << 8) | (cv(c) << 16) | (cv(d) << 24);
}
// power-of-two scale >= s (UE8M0 consumers: DeepSeek-V3.1 style scales, MXFP8)
VB_DEVICE float ceil_pow2(float s) {
  uint32_t b = __float_as_uint(s);
  uint32_t e = (b >> 23) & 0xff;
  if (b & 0x7fffff) ++e;
  return __uint_as_float(e << 23);
}

template <typename T, int PRE, int QM>
__global__ void __launch_bounds__(1024) quant_row_kernel(const Params p) {
  constexpr int N = 8;                       // elements per 16-byte vector (bf16 / fp16)
  __shared__ float red[32];
  pdl_trigger();
  pdl_wait();
  const int64_t row = blockIdx.x;
  const int H = p.H, nvec = H / N;
  const T* xr = reinterpret_cast<const T*>(p.x) + row * p.x_stride;
  float v[MAXV][N];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = threadIdx.x + i * blockDim.x;
    if (vi < nvec) {
      Vec16<T> a = ld16(xr + vi * N);
      if (PRE == PRE_ACT_MUL) {
        Vec16<T> u = ld16(xr + H + vi * N);
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const float g = to_f32(a.e[j]);
          const float ag = p.act == 0 ? silu(g) : (p.act == 1 ? gelu_erf(g) : gelu_tanh(g));
          v[i][j] = ag * to_f32(u.e[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < N; ++j) v[i][j] = to_f32(a.e[j]);
        if (PRE == PRE_RMS) {
          if (p.residual) {
            T* rr = reinterpret_cast<T*>(p.residual) + row * (int64_t)H + vi * N;
            Vec16<T> r = ld16(rr);
#pragma unroll
            for (int j = 0; j < N; ++j) {
              r.e[j] = from_f32<T>(v[i][j] + to_f32(r.e[j]));     // the stream is kept in T, as the unfused op does
              v[i][j] = to_f32(r.e[j]);
            }
            st16(rr, r);
          }
#pragma unroll
          for (int j = 0; j < N; ++j) ss += v[i][j] * v[i][j];
        }
      }
    }
  }
  if (PRE == PRE_RMS) {
    ss = block_sum(ss, red);
    const float inv = rsqrtf(ss / (float)H + p.eps);
    const T* w = reinterpret_cast<const T*>(p.weight);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = threadIdx.x + i * blockDim.x;
      if (vi < nvec) {
        Vec16<T> wv = ld16(w + vi * N);
#pragma unroll
        for (int j = 0; j < N; ++j) {
          float g = to_f32(wv.e[j]);
          if (p.gemma) g += 1.f;
          // round through T exactly like the unfused norm → quant sequence
          v[i][j] = to_f32(from_f32<T>(v[i][j] * inv * g));
        }
      }
    }
  }
  // ---- scales
  float inv_scale[MAXV];
  if (QM == Q_FP8_STATIC || QM == Q_INT8_STATIC) {
    const float s = 1.f / p.scale_in[0];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) inv_scale[i] = s;
  } else if (QM == Q_FP8_TOKEN || QM == Q_INT8_TOKEN) {
    float am = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
      if ((int)(threadIdx.x + i * blockDim.x) < nvec)
#pragma unroll
        for (int j = 0; j < N; ++j) am = fmaxf(am, fabsf(v[i][j]));
    am = warp_max(am);
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) red[wp] = am;
    __syncthreads();
    am = (lane < nw) ? red[lane] : 0.f;
    am = warp_max(am);
    const float qmax = QM == Q_FP8_TOKEN ? FP8_MAX : 127.f;
    float s = fmaxf(am / qmax, QM == Q_FP8_TOKEN ? 1e-12f : 1e-12f);
    if (p.scale_ub > 0.f) s = fminf(s, p.scale_ub);
    if (threadIdx.x == 0) p.scale_out[row] = s;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) inv_scale[i] = 1.f / s;
  } else {
    // per (token, group): `group / 8` consecutive threads hold one group → xor-shuffle inside that lane span
    const int tpg = p.group / N;               // 4 (G=32), 8 (G=64), 16 (G=128)
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = threadIdx.x + i * blockDim.x;
      float am = 0.f;
      if (vi < nvec)
#pragma unroll
        for (int j = 0; j < N; ++j) am = fmaxf(am, fabsf(v[i][j]));
      for (int o = tpg >> 1; o > 0; o >>= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o));
      float s = fmaxf(am, 1e-10f) / FP8_MAX;
      if (p.ue8m0) s = ceil_pow2(s);
      inv_scale[i] = 1.f / s;
      if (vi < nvec && (threadIdx.x & (tpg - 1)) == 0) {
        const int g = vi / tpg;
        if (p.col_major) p.scale_out[(int64_t)g * p.scale_stride + row] = s;
        else p.scale_out[row * (int64_t)(H / p.group) + g] = s;
      }
    }
  }
  // ---- convert + store (8 bytes per vector)
  uint2* qr = reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(p.q) + row * (int64_t)H);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = threadIdx.x + i * blockDim.x;
    if (vi < nvec) {
      const float s = inv_scale[i];
      uint2 o;
      if (QM == Q_INT8_TOKEN || QM == Q_INT8_STATIC) {
        o.x = pack_i8x4(v[i][0] * s, v[i][1] * s, v[i][2] * s, v[i][3] * s);
        o.y = pack_i8x4(v[i][4] * s, v[i][5] * s, v[i][6] * s, v[i][7] * s);
      } else {
        o.x = pack_fp8x4(v[i][0] * s, v[i][1] * s, v[i][2] * s, v[i][3] * s);
        o.y = pack_fp8x4(v[i][4] * s, v[i][5] * s, v[i][6] * s, v[i][7] * s);
      }
      qr[vi] = o;
    }
  }
}

// |x| max over a whole tensor → scale[0] = max(amax / 448, eps)  (dynamic per-tensor quantisation, pass 1 of 2)
template <typename T>
__global__ void __launch_bounds__(512) absmax_kernel(const T* __restrict__ x, int64_t n, int64_t row_len, int64_t stride,
                                                      unsigned* __restrict__ amax_bits) {
  __shared__ float red[32];
  float am = 0.f;
  const int64_t nv = n / 8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i * 8, r = e / row_len, c = e % row_len;
    Vec16<T> a = ld16(x + r * stride + c);
#pragma unroll
    for (int j = 0; j < 8; ++j) am = fmaxf(am, fabsf(to_f32(a.e[j])));
  }
  am = warp_max(am);
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  if (lane == 0) red[wp] = am;
  __syncthreads();
  if (wp == 0) {
    am = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.f;
    am = warp_max(am);
    if (lane == 0) atomicMax(amax_bits, __float_as_uint(am));     // non-negative floats order like their bit patterns
  }
}
__global__ void finish_scale_kernel(unsigned* amax_bits, float* scale, float qmax) {
  scale[0] = fmaxf(__uint_as_float(amax_bits[0]) / qmax, 1e-12f);
  amax_bits[0] = 0;                                               // self-resetting: graph replays start from zero
}

// ---------------------------------------------------------------- NVFP4 / MXFP4 (e2m1, two values per byte)
VB_DEVICE uint32_t pack_e2m1x8(const float (&f)[8]) {
  uint32_t out = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint16_t b;     // cvt packs (hi, lo) → one byte with `lo` in the low nibble
    asm("{\n\t.reg .b8 t;\n\tcvt.rn.satfinite.e2m1x2.f32 t, %2, %1;\n\tcvt.u16.u8 %0, t;\n\t}"
        : "=h"(b) : "f"(f[2 * k]), "f"(f[2 * k + 1]));
    out |= (uint32_t)(b & 0xff) << (8 * k);
  }
  return out;
}
// scale-factor offset in the 128x4 "swizzled" tile layout the block-scaled tcgen05 MMA reads (UTCCP 32x128b x4):
// tile (m / 128, k / 4) of 512 bytes; inside: [m % 32][ (m / 32) % 4 ][k % 4]
VB_DEVICE int64_t sf_swizzled_offset(int m, int k, int k_groups) {
  const int kt = (k_groups + 3) / 4;
  return ((int64_t)(m >> 7) * kt + (k >> 2)) * 512 + (m & 31) * 16 + ((m >> 5) & 3) * 4 + (k & 3);
}

// one thread = one 16-element block (NVFP4: e4m3 block scale x fp32 global scale; MXFP4: block 32 handled as two
// threads sharing an e8m0 scale).  x [T, H] → q [T, H/2] bytes, sf [T, H/16] (linear) or swizzled.
template <typename T, bool MX>
__global__ void __launch_bounds__(256) fp4_quant_kernel(const T* __restrict__ x, int64_t x_stride, uint8_t* __restrict__ q,
                                                         uint8_t* __restrict__ sf, const float* __restrict__ global_scale,
                                                         int Tn, int H, int swizzled) {
  pdl_trigger();
  pdl_wait();
  const int bpr = H / 16;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool live = idx < (int64_t)Tn * bpr;
  const int row = live ? (int)(idx / bpr) : 0, blk = live ? (int)(idx % bpr) : 0;
  float f[16];
  float am = 0.f;
  if (live) {
    const T* src = x + row * x_stride + blk * 16;
    Vec16<T> a = ld16(src), b = ld16(src + 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) { f[j] = to_f32(a.e[j]); f[8 + j] = to_f32(b.e[j]); }
#pragma unroll
    for (int j = 0; j < 16; ++j) am = fmaxf(am, fabsf(f[j]));
  }
  float out_scale;
  uint8_t sfb;
  if (MX) {
    am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 1));          // 32-element block = two neighbouring threads
    const float s = ceil_pow2(fmaxf(am, 1e-30f) / 6.f);
    sfb = (uint8_t)((__float_as_uint(s) >> 23) & 0xff);
    out_scale = 1.f / s;
  } else {
    const float gs = global_scale[0];
    const float sfv = gs * (am * (1.f / 6.f));
    __nv_fp8_e4m3 s8(sfv);
    sfb = *reinterpret_cast<uint8_t*>(&s8);
    const float sq = float(s8);
    out_scale = sq != 0.f ? gs / sq : 0.f;
  }
  if (!live) return;
#pragma unroll
  for (int j = 0; j < 16; ++j) f[j] *= out_scale;
  float lo[8], hi[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { lo[j] = f[j]; hi[j] = f[8 + j]; }
  uint2 o;
  o.x = pack_e2m1x8(lo);
  o.y = pack_e2m1x8(hi);
  *reinterpret_cast<uint2*>(q + row * (int64_t)(H / 2) + blk * 8) = o;
  if (MX) {
    if ((blk & 1) == 0) {
      const int kb = blk >> 1, kg = bpr / 2;
      sf[swizzled ? sf_swizzled_offset(row, kb, kg) : (int64_t)row * kg + kb] = sfb;
    }
  } else {
    sf[swizzled ? sf_swizzled_offset(row, blk, bpr) : (int64_t)row * bpr + blk] = sfb;
  }
}

template <typename T, int PRE>
static int launch_q(const Params& p, int T_, int qm, cudaStream_t st) {
  const int nvec = p.H / 8;
  int threads = ((nvec + MAXV - 1) / MAXV + 31) / 32 * 32;
  threads = max(32, min(1024, threads));
  if (qm == Q_FP8_GROUP) {
    const int tpg = p.group / 8;
    threads = max(threads, tpg);
    threads = (threads + tpg - 1) / tpg * tpg;
  }
  if ((int64_t)threads * MAXV < nvec) return -3;
  dim3 g(T_), b(threads);
  switch (qm) {
    case Q_FP8_STATIC: vb_launch(quant_row_kernel<T, PRE, Q_FP8_STATIC>, g, b, 0, st, p); break;
    case Q_FP8_TOKEN: vb_launch(quant_row_kernel<T, PRE, Q_FP8_TOKEN>, g, b, 0, st, p); break;
    case Q_FP8_GROUP: vb_launch(quant_row_kernel<T, PRE, Q_FP8_GROUP>, g, b, 0, st, p); break;
    case Q_INT8_TOKEN: vb_launch(quant_row_kernel<T, PRE, Q_INT8_TOKEN>, g, b, 0, st, p); break;
    case Q_INT8_STATIC: vb_launch(quant_row_kernel<T, PRE, Q_INT8_STATIC>, g, b, 0, st, p); break;
    default: return -2;
  }
  return (int)cudaGetLastError();
}

}  // namespace quant

// One entry point for the whole family.
//   pre : 0 none, 1 RMSNorm (+residual if non-null), 2 act(gate)*up over x[T, 2H]
//   qm  : 0 fp8 static, 1 fp8 dynamic per token, 2 fp8 per (token, group), 3 int8 dynamic per token, 4 int8 static
EXPORT int vb_quant_rows(const void* x, int64_t x_stride, void* residual, const void* weight, void* q, float* scale_out,
                         const float* scale_in, int T, int H, int pre, int qm, int group, int col_major,
                         int scale_stride, int ue8m0, int gemma, int act, float eps, float scale_ub, int dtype,
                         cudaStream_t st) {
  using namespace quant;
  if (T == 0) return 0;
  if (dtype == DT_F32 || H % 8 != 0) return -1;
  if (qm == Q_FP8_GROUP && (group != 32 && group != 64 && group != 128)) return -1;
  if (qm == Q_FP8_GROUP && H % group) return -1;
  Params p;
  p.x = x; p.residual = residual; p.weight = weight; p.q = q; p.scale_out = scale_out; p.scale_in = scale_in;
  p.x_stride = x_stride; p.H = H; p.group = group; p.scale_stride = scale_stride; p.col_major = col_major;
  p.ue8m0 = ue8m0; p.gemma = gemma; p.act = act; p.eps = eps; p.scale_ub = scale_ub;
  int rc = -2;
  DISPATCH_HALF(dtype, Tt, {
    if (pre == PRE_NONE) rc = launch_q<Tt, PRE_NONE>(p, T, qm, st);
    else if (pre == PRE_RMS) rc = launch_q<Tt, PRE_RMS>(p, T, qm, st);
    else if (pre == PRE_ACT_MUL) rc = launch_q<Tt, PRE_ACT_MUL>(p, T, qm, st);
  });
  return rc;
}

// dynamic per-tensor scale: scale[0] = max(|x|) / qmax.  `amax_bits` is one zero-initialised uint32 of scratch.
EXPORT int vb_absmax_scale(const void* x, int64_t rows, int64_t row_len, int64_t stride, unsigned* amax_bits, float* scale,
                           float qmax, int num_sms, int dtype, cudaStream_t st) {
  using namespace quant;
  if (dtype == DT_F32 || row_len % 8) return -1;
  const int64_t n = rows * row_len;
  int64_t nb = (n / 8 + 511) / 512;
  if (nb > (int64_t)num_sms * 4) nb = (int64_t)num_sms * 4;
  if (nb < 1) nb = 1;
  const int blocks = (int)nb;
  DISPATCH_HALF(dtype, Tt, {
    absmax_kernel<Tt><<<blocks, 512, 0, st>>>(reinterpret_cast<const Tt*>(x), n, row_len, stride, amax_bits);
  });
  finish_scale_kernel<<<1, 1, 0, st>>>(amax_bits, scale, qmax);
  return (int)cudaGetLastError();
}

// NVFP4 (mx = 0: block 16, e4m3 block scales, fp32 global scale) / MXFP4 (mx = 1: block 32, e8m0 scales)
EXPORT int vb_fp4_quant(const void* x, int64_t x_stride, void* q, void* sf, const float* global_scale, int T, int H, int mx,
                        int swizzled, int dtype, cudaStream_t st) {
  using namespace quant;
  if (T == 0) return 0;
  if (dtype == DT_F32 || H % 32 != 0) return -1;
  const int64_t total = (int64_t)T * (H / 16);
  const int blocks = (int)((total + 255) / 256);
  DISPATCH_HALF(dtype, Tt, {
    if (mx) vb_launch(fp4_quant_kernel<Tt, true>, dim3(blocks), dim3(256), 0, st, reinterpret_cast<const Tt*>(x), x_stride,
                      reinterpret_cast<uint8_t*>(q), reinterpret_cast<uint8_t*>(sf), global_scale, T, H, swizzled);
    else vb_launch(fp4_quant_kernel<Tt, false>, dim3(blocks), dim3(256), 0, st, reinterpret_cast<const Tt*>(x), x_stride,
                   reinterpret_cast<uint8_t*>(q), reinterpret_cast<uint8_t*>(sf), global_scale, T, H, swizzled);
  });
  return (int)cudaGetLastError();
}
