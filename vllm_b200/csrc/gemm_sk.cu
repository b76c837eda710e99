// Persistent stream-K decode GEMM on tcgen05:  Y[M,N] = X[M,K] · W[N,K]^T (+bias) (optionally SiLU(gate)·up),  M <= 256.
//
// Why: decode GEMMs are weight-bandwidth bound, and one SM can keep only ~128 KB of TMA loads in flight (≈60 GB/s), so
// the weight stream reaches HBM speed only when ALL 148 SMs stream ALL the time.  Tile-per-CTA grids (gemm_tc.cu) leave
// SMs idle whenever the tile count is not a multiple of the SM count (gate_up: 224 tiles = 1.51 waves; qkv: 48 tiles
// x 2 splits = 96 CTAs) and pay a pipeline fill / drain per tile.  Here the grid is one CTA per SM and the flattened
// (feature tile, k-block) space is cut into equal contiguous chunks (stream-K): every CTA streams the same number of
// weight bytes through ONE TMA ring that never drains between tiles, accumulators ping-pong between two TMEM buffers
// so the epilogue of a tile runs under the MMAs of the next, and tiles cut by a chunk boundary are finished by their
// last-arriving contributor from fp32 partials in an L2-resident workspace (summed in a fixed order: results are
// bit-identical run to run; counters reset themselves: CUDA-graph safe).
//
// swap-AB operand roles as in gemm_tc.cu: A := weight tile (128 output features on the MMA M axis), B := all M tokens
// (MMA N = M padded to 16).  Reference op: cuBLASLt via layers/linear.py:210-218.
#include "common.cuh"

namespace gemmsk {

constexpr int BK = 64;
constexpr int TM = 128;
constexpr int NUM_THREADS = 192;
constexpr int EPI_THREADS = 128;

struct Params {
  void* y;
  const void* bias;
  float* ws;               // [grid][2][BN][128] fp32 partial tiles
  unsigned* counters;      // [n_tiles] zero-initialised, self-resetting
  int M, N, K;
  int kbt;                 // k-blocks per tile
  int per;                 // (tile, k-block) units per CTA
  int total;               // n_tiles * kbt
  int act;                 // 1: rows are interleaved (gate_i, up_i) pairs → y[:, i] = silu(gate_i) * up_i, N_out = N / 2
};

VB_DEVICE void bar_sync_epi() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
VB_DEVICE float4 ld_cg_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
VB_DEVICE float ld_cg_f(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

template <typename T, int BN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_sk_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const Params p) {
  constexpr int A_BYTES = TM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;      // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  volatile unsigned* s_flag = reinterpret_cast<volatile unsigned*>(tmem_slot + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x;
  const int u0 = min(cta * p.per, p.total), u1 = min(u0 + p.per, p.total);
  const int n_units = u1 - u0;
  constexpr uint32_t TMEM_COLS = 2 * BN <= 32 ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));

  pdl_trigger();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
    mbar_init(&acc_empty[0], 4); mbar_init(&acc_empty[1], 4);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0 && n_units > 0) {
      // ===== TMA producer: one ring over the whole chunk.  The first ring of weight tiles does not depend on the
      // previous kernel and is requested before griddepcontrol.wait (PDL): the stream starts under its tail.
      const int pre = min(n_units, STAGES);
      for (int it = 0; it < pre; ++it) {
        const int u = u0 + it, tile = u / p.kbt, kb = u - tile * p.kbt;
        mbar_expect_tx(&full[it], STAGE_BYTES);
        tma_load_2d_hint(smem + it * STAGE_BYTES, &tmap_a, kb * BK, tile * TM, &full[it], L2_EVICT_FIRST);
      }
      pdl_wait();
      for (int it = 0; it < pre; ++it) {
        const int u = u0 + it, tile = u / p.kbt, kb = u - tile * p.kbt;
        tma_load_2d_hint(smem + it * STAGE_BYTES + A_BYTES, &tmap_b, kb * BK, 0, &full[it], L2_EVICT_LAST);
      }
      for (int it = pre; it < n_units; ++it) {
        const int s = it % STAGES;
        const int u = u0 + it, tile = u / p.kbt, kb = u - tile * p.kbt;
        mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
        mbar_expect_tx(&full[s], STAGE_BYTES);
        tma_load_2d_hint(smem + s * STAGE_BYTES, &tmap_a, kb * BK, tile * TM, &full[s], L2_EVICT_FIRST);
        tma_load_2d_hint(smem + s * STAGE_BYTES + A_BYTES, &tmap_b, kb * BK, 0, &full[s], L2_EVICT_LAST);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: a segment = the k-blocks of one tile inside this chunk → one TMEM accumulator
    const uint32_t idesc = umma_idesc_f16(TM, BN, DTypeOf<T>::v == DT_BF16);
    int seg = 0;
    bool seg_start = true;
    for (int it = 0; it < n_units; ++it) {
      const int s = it % STAGES;
      const int u = u0 + it, kb = u % p.kbt;
      const int buf = seg & 1;
      if (seg_start) {
        mbar_wait(&acc_empty[buf], ((seg >> 1) & 1) ^ 1);
        tc_fence_after();
      }
      mbar_wait(&full[s], (it / STAGES) & 1);
      tc_fence_after();
      const bool seg_end = (kb == p.kbt - 1) || (it == n_units - 1);
      if (lane == 0) {
        const unsigned char* a_src = smem + s * STAGE_BYTES;
        const uint64_t da = umma_desc_sw128(a_src), db = umma_desc_sw128(a_src + A_BYTES);
        const uint32_t d = tmem_base + (uint32_t)(buf * BN);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_f16(d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (uint32_t)(!(seg_start && k == 0)));
        umma_commit(&empty[s]);
        if (seg_end) umma_commit(&acc_full[buf]);
      }
      __syncwarp();
      seg_start = seg_end;
      if (seg_end) ++seg;
    }
  } else {
    // ===== epilogue warps
    constexpr int CH = (BN % 32 == 0) ? 32 : 16;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    T* y = reinterpret_cast<T*>(p.y);
    const int n_out = p.act ? p.N / 2 : p.N;
    pdl_wait();
    int seg = 0;
    int u = u0;
    while (u < u1) {
      const int tile = u / p.kbt;
      const int kb_lo = u - tile * p.kbt;
      const int kb_hi = min(p.kbt, kb_lo + (u1 - u));          // exclusive
      const int buf = seg & 1;
      const bool whole = kb_lo == 0 && kb_hi == p.kbt;
      mbar_wait(&acc_full[buf], (seg >> 1) & 1);
      tc_fence_after();
      auto load_chunk = [&](int c, float (&v)[CH]) {
        if constexpr (CH == 32) {
          uint32_t r[32];
          tmem_ld_32x32(t_row + (uint32_t)(buf * BN + c), r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        } else {
          uint32_t r[16];
          tmem_ld_32x16(t_row + (uint32_t)(buf * BN + c), r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        }
      };
      // final store of `CH` columns starting at c for this thread's feature row
      auto store_chunk = [&](int c, float (&v)[CH]) {
        const int n = tile * TM + row;
        if (p.act) {
          // interleaved rows: even lane = gate_i, odd lane = up_i
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const float other = __shfl_xor_sync(0xffffffffu, v[j], 1);
            const float g = v[j];
            v[j] = (g / (1.f + __expf(-g))) * other;
          }
          if ((lane & 1) == 0 && n < p.N) {
            const int no = n >> 1;
#pragma unroll
            for (int j = 0; j < CH; ++j)
              if (c + j < p.M) y[(size_t)(c + j) * n_out + no] = from_f32<T>(v[j]);
          }
        } else if (n < p.N) {
          const float bias = p.bias ? to_f32(reinterpret_cast<const T*>(p.bias)[n]) : 0.f;
#pragma unroll
          for (int j = 0; j < CH; ++j)
            if (c + j < p.M) y[(size_t)(c + j) * p.N + n] = from_f32<T>(v[j] + bias);
        }
      };
      if (whole) {
#pragma unroll
        for (int c = 0; c < BN; c += CH) {
          float v[CH];
          load_chunk(c, v);
          if (c + CH >= BN) {                      // accumulator fully read: hand the buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
          }
          store_chunk(c, v);
        }
      } else {
        // ---- partial tile: park fp32 in this CTA's workspace slot; the last contributor to arrive finishes the tile
        const int which = ((cta * p.per) / p.kbt == tile) ? 0 : 1;
        float* slot = p.ws + ((size_t)(cta * 2 + which) * BN) * TM;
#pragma unroll
        for (int c = 0; c < BN; c += CH) {
          float v[CH];
          load_chunk(c, v);
          if (c + CH >= BN) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
          }
#pragma unroll
          for (int j = 0; j < CH; ++j)
            if (c + j < p.M) slot[(c + j) * TM + row] = v[j];
        }
        __threadfence();
        bar_sync_epi();
        const int first = (tile * p.kbt) / p.per, last = ((tile + 1) * p.kbt - 1) / p.per;
        if (et == 0) {
          const unsigned old = atomicAdd(&p.counters[tile], 1u);
          const bool fin = old == (unsigned)(last - first);
          if (fin) p.counters[tile] = 0;            // self-reset: every contributor has already arrived
          *s_flag = fin ? 1u : 0u;
        }
        bar_sync_epi();
        if (*s_flag) {
          __threadfence();
          // fixed summation order (contributor CTAs first..last) → bit-identical results whoever arrives last
          for (int c = 0; c < BN && c < p.M; c += CH) {
            float v[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] = 0.f;
            for (int cc = first; cc <= last; ++cc) {
              const int w2 = ((cc * p.per) / p.kbt == tile) ? 0 : 1;
              const float* src = p.ws + ((size_t)(cc * 2 + w2) * BN) * TM;
#pragma unroll
              for (int j = 0; j < CH; ++j)
                if (c + j < p.M) v[j] += ld_cg_f(src + (c + j) * TM + row);
            }
            store_chunk(c, v);
          }
        }
        bar_sync_epi();                              // s_flag is reused by the next partial segment
      }
      u += kb_hi - kb_lo;
      ++seg;
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
static int make_tmap(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                     uint32_t box_rows, bool bf16) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -10;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -11;
}

template <typename T, int BN, int STAGES>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, int grid, cudaStream_t st) {
  constexpr int STAGE_BYTES = TM * BK * 2 + BN * BK * 2;
  const int smem = STAGES * STAGE_BYTES + (2 * STAGES + 4) * 8 + 32 + 1024;
  auto k = gemm_sk_kernel<T, BN, STAGES>;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
  vb_launch(k, dim3(grid), dim3(NUM_THREADS), (size_t)smem, st, ta, tb, p);
  return (int)cudaGetLastError();
}

}  // namespace gemmsk

// plan: token tile width bn (M padded), grid (CTAs), units per CTA, workspace bytes and counters needed
EXPORT int vb_gemm_sk_plan(int M, int N, int K, int num_sms, int* bn, int* grid, int* per, int64_t* ws_bytes, int* n_tiles) {
  using namespace gemmsk;
  if (M > 256 || K % BK) return -1;
  int b = ((M + 15) / 16) * 16;
  if (b > 128 && b < 256) b = 256;
  const int nt = (N + TM - 1) / TM, kbt = K / BK;
  const long long total = (long long)nt * kbt;
  int g = num_sms;
  if (total / 8 < g) g = (int)(total / 8 > 1 ? total / 8 : 1);     // at least 8 k-blocks (128 KB of weights) per CTA
  int pr = (int)((total + g - 1) / g);
  g = (int)((total + pr - 1) / pr);
  *bn = b; *grid = g; *per = pr; *n_tiles = nt;
  *ws_bytes = (int64_t)g * 2 * b * TM * 4;
  return 0;
}

EXPORT int vb_gemm_sk(const void* x, const void* w, const void* bias, void* y, int M, int N, int K, int64_t x_stride, int bn,
                      int grid, int per, float* ws, unsigned* counters, int act, int dtype, cudaStream_t st) {
  using namespace gemmsk;
  if (dtype == DT_F32 || K % BK != 0 || M > 256) return -1;
  if (act && (N % 2 || bias)) return -1;
  const bool bf16 = dtype == DT_BF16;
  CUtensorMap ta, tb;
  int rc;
  if ((rc = make_tmap(&ta, w, N, K, K, TM, bf16))) return rc;
  if ((rc = make_tmap(&tb, x, M, K, x_stride, bn, bf16))) return rc;
  Params p;
  p.y = y; p.bias = bias; p.ws = ws; p.counters = counters; p.M = M; p.N = N; p.K = K; p.kbt = K / BK; p.per = per;
  p.total = ((N + TM - 1) / TM) * p.kbt; p.act = act;
#define GO(BNV, ST)                                                                    \
  return bf16 ? launch<__nv_bfloat16, BNV, ST>(ta, tb, p, grid, st) : launch<__half, BNV, ST>(ta, tb, p, grid, st)
  switch (bn) {
    case 16: GO(16, 10);
    case 32: GO(32, 10);
    case 48: GO(48, 9);
    case 64: GO(64, 8);
    case 80: GO(80, 8);
    case 96: GO(96, 7);
    case 112: GO(112, 7);
    case 128: GO(128, 6);
    case 256: GO(256, 4);
    default: return -2;
  }
#undef GO
}
