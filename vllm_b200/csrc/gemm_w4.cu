// Weight-only FP4 decode GEMM (W4A16): Y[M, N] = X[M, K] · dequant(W)[N, K]^T, M <= 128 tokens.
//
// W is NVFP4 (e2m1 codes, one e4m3 scale per 16 values, one fp32 scale per tensor) or MXFP4 (e2m1 codes, one e8m0
// power-of-two scale per 32 values — the GPT-OSS expert format).  Reference: csrc/libtorch_stable/quantization/fp4/
// nvfp4_scaled_mm_kernels.cu (W4A4 through CUTLASS block-scaled MMA), quantization/mxfp4.py / marlin fp4 (W4A16).
// Decode is bound by the bytes of W, so the point of FP4 here is to stream a quarter of them: the kernel keeps the
// activations in bf16/fp16 and expands the codes on the fly —
//   * the host packs W once into tile-major blocks [N/128][K/64][128 rows x 32 B codes | 128 x {4,2} B scales], so one
//     (tile, k-block) operand is ONE contiguous 4.5 KB bulk copy (cp.async.bulk → UBLKCP), no tensor map, full lines;
//   * eight converter warps (two threads per weight row) turn 64 codes into 64 bf16 with two byte-permutes per four codes
//     (PRMT as an 8-entry byte LUT: every e2m1 magnitude is exact in bf16), one packed multiply by the block scale
//     (exact: <= 5 significant bits), and write them in the 128-byte-swizzled K-major layout tcgen05 reads;
//   * a single thread issues tcgen05.mma kind::f16 (swap-AB: weight rows on the MMA M axis, tokens on N), fp32
//     accumulators in TMEM; the epilogue applies the tensor scale and the bias.
// Rings: 8 stages of (packed W block + activation tile, TMA) feed 3 stages of converted W; mbarriers both ways.
//
// Measured (profiles/README.md R2.8): correct on every shape, 5x faster than expanding the weight per call, but 0.5-0.65x
// of the bf16 kernel: the ALU expansion costs ~3 instructions per element and a k-block takes ~970 cycles per CTA
// (scripts/w4_trace.py: 476 convert+store, ~450 barrier / fence / load latency) against ~200 for its bytes.  The
// bandwidth-bound version is the block-scaled tensor-core path (tcgen05.mma kind::mxf8f6f4 / mxf4nvf4 with the scales in
// TMEM), which consumes the packed codes directly; this kernel is the portable fallback and the layout it will share.
#include "common.cuh"

namespace w4 {

constexpr int BK = 64, TM = 128, NUM_THREADS = 320, CONV_THREADS = 256;   // 1 TMA + 1 MMA + 8 converter / epilogue warps
constexpr int PSTAGES = 8, CSTAGES = 3;
constexpr int CODE_BYTES = TM * BK / 2;            // 4096
constexpr int CONV_BYTES = TM * BK * 2;            // 16384
constexpr int PK_STRIDE = 4608;                    // codes + up to 512 B of scales, 128-byte aligned

struct Params {
  void* y;
  const void* bias;
  const unsigned char* w;     // tile-major packed weights
  float gscale;               // tensor scale (1 for MXFP4)
  int M, N, K;
  int fmt;                    // 0 = NVFP4 (e4m3 scale / 16), 1 = MXFP4 (e8m0 scale / 32)
  int blk_bytes;              // bytes of one (tile, k-block) block in `w`
  long long* trace;           // debug: CTA 0 clock64 stamps [kb][8] (null in production)
};

VB_DEVICE uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

template <typename T> struct Lut;
// bytes of the 16-bit pattern of {0, .5, 1, 1.5, 2, 3, 4, 6}: HI0/HI1 = high bytes of entries 0-3 / 4-7, LO0/LO1 = low bytes
template <> struct Lut<__nv_bfloat16> {
  static constexpr uint32_t HI0 = 0x3F3F3F00u, HI1 = 0x40404040u, LO0 = 0xC0800000u, LO1 = 0xC0804000u;
};
template <> struct Lut<__half> {
  static constexpr uint32_t HI0 = 0x3E3C3800u, HI1 = 0x46444240u, LO0 = 0u, LO1 = 0u;
};

template <typename T> VB_DEVICE uint32_t mul2(uint32_t a, uint32_t s);
template <> VB_DEVICE uint32_t mul2<__nv_bfloat16>(uint32_t a, uint32_t s) {
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(s));
  return d;
}
template <> VB_DEVICE uint32_t mul2<__half>(uint32_t a, uint32_t s) {
  uint32_t d;
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(s));
  return d;
}

// block scale → both halves of a packed 16-bit pair
template <typename T> VB_DEVICE uint32_t scale_pair(uint32_t byte, int fmt) {
  float f;
  if (fmt == 0) {
    const __half_raw hr = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)byte, __NV_E4M3);
    f = __half2float(__half(hr));
  } else {
    f = byte == 0 ? 0.f : __uint_as_float(byte << 23);        // 2^(e-127)
  }
  return pack2<T>(f, f);
}

// four codes (the low 16 bits of `x`, code 0 in bits 0-3) → two packed 16-bit pairs, scaled
template <typename T>
VB_DEVICE void expand4(uint32_t x, uint32_t sc, uint32_t& o01, uint32_t& o23) {
  const uint32_t sel = x & 0x7777u;
  const uint32_t H = prmt(Lut<T>::HI0, Lut<T>::HI1, sel);
  const uint32_t L = prmt(Lut<T>::LO0, Lut<T>::LO1, sel);
  uint32_t w01 = prmt(L, H, 0x5140u), w23 = prmt(L, H, 0x7362u);
  w01 |= ((x & 0x8u) << 12) | ((x & 0x80u) << 24);
  w23 |= ((x & 0x800u) << 4) | ((x & 0x8000u) << 16);
  o01 = mul2<T>(w01, sc);
  o23 = mul2<T>(w23, sc);
}

template <typename T, int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_w4_kernel(const __grid_constant__ CUtensorMap tmap_b, const Params p) {
  constexpr int B_BYTES = BN * BK * 2;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* conv = smem;                                      // CSTAGES x 16 KB, 1024-aligned
  unsigned char* bsm = conv + CSTAGES * CONV_BYTES;                // PSTAGES x B_BYTES, 1024-aligned
  unsigned char* pk = bsm + PSTAGES * B_BYTES;                     // PSTAGES x PK_STRIDE
  uint64_t* full_p = reinterpret_cast<uint64_t*>(pk + PSTAGES * PK_STRIDE);
  uint64_t* empty_p = full_p + PSTAGES;
  uint64_t* conv_full = empty_p + PSTAGES;
  uint64_t* conv_empty = conv_full + CSTAGES;
  uint64_t* tmem_full = conv_empty + CSTAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a_tile = blockIdx.x, b_tile = blockIdx.y;
  const int nkb = p.K / BK;
  const int b_row0 = b_tile * BN;
  const int m_lim = p.M - b_row0;
  constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : 128);

  pdl_trigger();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_b);
#pragma unroll
    for (int s = 0; s < PSTAGES; ++s) { mbar_init(&full_p[s], 1); mbar_init(&empty_p[s], 1); }
#pragma unroll
    for (int c = 0; c < CSTAGES; ++c) { mbar_init(&conv_full[c], CONV_THREADS); mbar_init(&conv_empty[c], 1); }
    mbar_init(tmem_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // the packed weights do not depend on the previous kernel: request the first ring before griddepcontrol.wait
      const unsigned char* wt = p.w + (size_t)a_tile * nkb * p.blk_bytes;
      const int pre = min(nkb, PSTAGES);
      for (int kb = 0; kb < pre; ++kb) {
        mbar_expect_tx(&full_p[kb], (uint32_t)(p.blk_bytes + B_BYTES));
        bulk_g2s(pk + kb * PK_STRIDE, wt + (size_t)kb * p.blk_bytes, (uint32_t)p.blk_bytes, &full_p[kb]);
      }
      pdl_wait();
      for (int kb = 0; kb < pre; ++kb)
        tma_load_2d_hint(bsm + kb * B_BYTES, &tmap_b, kb * BK, b_row0, &full_p[kb], L2_EVICT_LAST);
      for (int kb = pre; kb < nkb; ++kb) {
        const int s = kb % PSTAGES;
        mbar_wait(&empty_p[s], ((kb / PSTAGES) & 1) ^ 1);
        mbar_expect_tx(&full_p[s], (uint32_t)(p.blk_bytes + B_BYTES));
        bulk_g2s(pk + s * PK_STRIDE, wt + (size_t)kb * p.blk_bytes, (uint32_t)p.blk_bytes, &full_p[s]);
        tma_load_2d_hint(bsm + s * B_BYTES, &tmap_b, kb * BK, b_row0, &full_p[s], L2_EVICT_LAST);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_f16(TM, BN, DTypeOf<T>::v == DT_BF16);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % PSTAGES, c = kb % CSTAGES;
      mbar_wait(&full_p[s], (kb / PSTAGES) & 1);             // activation tile (async proxy) landed
      mbar_wait(&conv_full[c], (kb / CSTAGES) & 1);          // converted weights written + fenced
      tc_fence_after();
      if (lane == 0) {
        if (p.trace && blockIdx.x == 0 && kb < 256) p.trace[kb * 8 + 5] = clock64();
        const uint64_t da = umma_desc_sw128(conv + c * CONV_BYTES), db = umma_desc_sw128(bsm + s * B_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_f16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        umma_commit(&conv_empty[c]);
        umma_commit(&empty_p[s]);
        if (kb == nkb - 1) umma_commit(tmem_full);
      }
      __syncwarp();
    }
  } else {
    // ===== converters, then epilogue.  Two threads per weight row (row = TMEM lane of the warp's quarter): warps 2-5 expand
    // the first 32 codes of a 64-code block, warps 6-9 the second 32 — two converter warps per scheduler hide each other's
    // shared-memory and barrier latency.
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int half = (warp - 2) >> 2;
    const uint32_t sw = (uint32_t)(row & 7);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % PSTAGES, c = kb % CSTAGES;
      const bool tr = p.trace && blockIdx.x == 0 && threadIdx.x == 64 && kb < 256;
      if (tr) p.trace[kb * 8 + 0] = clock64();
      mbar_wait(&full_p[s], (kb / PSTAGES) & 1);
      if (tr) p.trace[kb * 8 + 1] = clock64();
      const unsigned char* blk = pk + s * PK_STRIDE;
      const uint4 qv = *reinterpret_cast<const uint4*>(blk + row * 32 + half * 16);
      uint32_t sc[2];
      if (p.fmt == 0) {
        const uint32_t scw = *reinterpret_cast<const uint16_t*>(blk + CODE_BYTES + row * 4 + half * 2);
        sc[0] = scale_pair<T>(scw & 0xffu, 0);
        sc[1] = scale_pair<T>(scw >> 8, 0);
      } else {
        sc[0] = sc[1] = scale_pair<T>(blk[CODE_BYTES + row * 2 + half], 1);
      }
      mbar_wait(&conv_empty[c], ((kb / CSTAGES) & 1) ^ 1);
      if (tr) p.trace[kb * 8 + 2] = clock64();
      unsigned char* dst = conv + c * CONV_BYTES + (row >> 3) * 1024 + (row & 7) * 128;
      const uint32_t words[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {                 // word = 8 codes = one 16-byte chunk of the row
        uint4 o;
        expand4<T>(words[jj], sc[jj >> 1], o.x, o.y);
        expand4<T>(words[jj] >> 16, sc[jj >> 1], o.z, o.w);
        *reinterpret_cast<uint4*>(dst + (((uint32_t)(half * 4 + jj) ^ sw) << 4)) = o;
      }
      if (tr) p.trace[kb * 8 + 3] = clock64();
      fence_proxy_async();                             // generic-proxy writes → visible to the tensor core's reads
      if (tr) p.trace[kb * 8 + 4] = clock64();
      mbar_arrive(&conv_full[c]);
    }
    // ----- epilogue: accumulator is Y^T (row = output feature, column = token)
    constexpr int CH = (BN % 32 == 0) ? 32 : 16;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    pdl_wait();
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    T* y = reinterpret_cast<T*>(p.y);
    const int n = a_tile * TM + row;
    const float bias = (p.bias && n < p.N) ? to_f32(reinterpret_cast<const T*>(p.bias)[n]) : 0.f;
    for (int c0 = half * CH; c0 < BN; c0 += 2 * CH) {      // the two threads of a row alternate token chunks
      if (c0 >= m_lim) break;
      float v[CH];
      if constexpr (CH == 32) {
        uint32_t r[32];
        tmem_ld_32x32(t_row + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      } else {
        uint32_t r[16];
        tmem_ld_32x16(t_row + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
      }
      if (n < p.N) {
#pragma unroll
        for (int j = 0; j < CH; ++j)
          if (c0 + j < m_lim) y[(size_t)(b_row0 + c0 + j) * p.N + n] = from_f32<T>(v[j] * p.gscale + bias);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || !ptr) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

template <typename T, int BN>
static int launch(const CUtensorMap& tb, const Params& p, dim3 grid, cudaStream_t st) {
  const int smem = CSTAGES * CONV_BYTES + PSTAGES * (BN * BK * 2) + PSTAGES * PK_STRIDE + (2 * PSTAGES + 2 * CSTAGES + 1) * 8 +
                   16 + 1024;
  auto k = gemm_w4_kernel<T, BN>;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
  vb_launch(k, grid, dim3(NUM_THREADS), (size_t)smem, st, tb, p);
  return (int)cudaGetLastError();
}

}  // namespace w4

// x: [M, K] 16-bit (row stride x_stride), w: tile-major packed FP4 (see vllm_b200/ops/cuda.py::pack_w4_tiles),
// y: [M, N].  M <= 128, K % 64 == 0.  fmt: 0 NVFP4, 1 MXFP4.
static long long* g_w4_trace = nullptr;
EXPORT void vb_gemm_w4_set_trace(long long* t) { g_w4_trace = t; }

EXPORT int vb_gemm_w4(const void* x, const void* w, const void* bias, void* y, int M, int N, int K, int64_t x_stride,
                      float gscale, int fmt, int dtype, cudaStream_t st) {
  using namespace w4;
  if (dtype == DT_F32 || K % BK != 0 || M < 1 || M > 128 || (fmt != 0 && fmt != 1)) return -1;
  auto enc = get_encode();
  if (!enc) return -10;
  const bool bf16 = dtype == DT_BF16;
  const int bn = ((M + 15) / 16) * 16;
  CUtensorMap tb;
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)x_stride * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)bn};
    cuuint32_t es[2] = {1, 1};
    if (enc(&tb, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(x), dims,
            strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return -11;
  }
  Params p;
  p.y = y; p.bias = bias; p.w = reinterpret_cast<const unsigned char*>(w); p.gscale = gscale; p.M = M; p.N = N; p.K = K;
  p.fmt = fmt; p.blk_bytes = CODE_BYTES + (fmt == 0 ? 512 : 256); p.trace = g_w4_trace;
  dim3 grid((N + TM - 1) / TM, 1, 1);
#define GO(BNV)                                                              \
  if (bf16) return launch<__nv_bfloat16, BNV>(tb, p, grid, st);             \
  else return launch<__half, BNV>(tb, p, grid, st);
  switch (bn) {
    case 16: GO(16)
    case 32: GO(32)
    case 48: GO(48)
    case 64: GO(64)
    case 80: GO(80)
    case 96: GO(96)
    case 112: GO(112)
    case 128: GO(128)
    default: return -2;
  }
#undef GO
}
