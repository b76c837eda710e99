// FP8 (e4m3) GEMM on the 5th-gen tensor cores, tcgen05 `kind::f8f6f4`:  Y[M,N] = (Xq · Wq^T) ∘ scales  → bf16/fp16
//
//   MODE_BLOCK  DeepSeek-V3 style block scaling: activations per (token, 128-wide K group), weights per 128x128
//               block (reference: csrc/libtorch_stable/quantization/w8a8/cutlass/c3x/
//               scaled_mm_blockwise_sm100_fp8_dispatch.cuh, DeepGEMM).  A K-block of 128 fp8 values is exactly one
//               128-byte swizzle row and one scale group, so the kernel accumulates EACH K-block into its own TMEM
//               buffer (two buffers, ping-pong), and the epilogue warps promote it into fp32 registers with
//               scale = w_scale[n_tile, kb] * x_scale[token, kb] while the tensor core already works on the next
//               K-block: full-precision promotion every 128 elements of K at no extra pass over the data.
//   MODE_CHANNEL per-token (or per-tensor) activation scale x per-channel (or per-tensor) weight scale (reference:
//               scaled_mm_sm100_fp8_dispatch.cuh): one TMEM accumulation over all of K, scales applied once.
//
// Operand roles follow gemm_tc.cu: SWAP=1 (decode, M <= 128 per tile) puts the 128 output features of a weight tile
// on the MMA M axis and the tokens on N — every CTA streams its [128 x K/splits] fp8 weight slab exactly once, at half
// the bytes of the bf16 kernel; few-tile GEMMs split K over a thread-block cluster and reduce through distributed
// shared memory.  SWAP=0 (prefill) is the 128-token x 128-feature tile, where both scales are per-thread scalars.
// Grouped (MoE experts) mode as in gemm_tc.cu, with per-expert weight scales.
//
// Pipeline: warp 0 TMA producer (SW128 boxes of 128 B x rows) → smem ring → warp 1 issues 4 x tcgen05.mma (K = 32 each)
// per K-block → tcgen05.commit → warps 2-5 tcgen05.ld + scale + accumulate.
#include "common.cuh"

namespace gemm8 {

constexpr int BK = 128;           // fp8 elements per K-block = 128 B = one swizzle row = one scale group
constexpr int TM = 128;
constexpr int NUM_THREADS = 192;
constexpr int EPI_THREADS = 128;
constexpr int SC_KB_MAX = 96;
// K-blocks of combined scales staged in shared memory per CTA (token tiles up to 64 wide: 96, wider: 64)
VB_HOST_DEVICE constexpr int sc_kb(int bn) { return bn <= 64 ? 96 : 64; }

struct Params {
  void* y;
  const void* bias;
  const float* xs;        // MODE_BLOCK: [KB, xs_stride] (column major: group-major, token contiguous); MODE_CHANNEL: [M] or [1]
  const float* ws;        // MODE_BLOCK: [(E *) ceil(N/128), KB]; MODE_CHANNEL: [N] or [1]
  int64_t xs_stride;
  int xs_per_token, ws_per_channel;   // MODE_CHANNEL: 0 = per-tensor scalar
  int M, N, K;
  int k_blocks_per_split, splits;
  const int* expert_offs;
  int n_experts;
};

template <int BN> struct Smem {
  static constexpr int A_BYTES = TM * BK;
  static constexpr int B_BYTES = BN * BK;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
};

template <typename T, int BN, int STAGES, bool SWAP, bool BLOCK>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_fp8_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const Params p) {
  using L = Smem<BN>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* stage_base = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * L::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;      // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  // combined scales ws[tile, kb] * xs[token, kb] of this CTA's K range, staged once (SWAP + BLOCK): the per-K-block
  // promotion then reads them with broadcast LDS instead of waiting on L2 (ncu: 70 % of the epilogue's issue slots
  // were long-scoreboard stalls on those loads)
  float* sc_s = reinterpret_cast<float*>(tmem_slot + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a_tile = blockIdx.x, b_tile = blockIdx.y, split = blockIdx.z;
  const int total_kb = (p.K + BK - 1) / BK;
  const int kb0 = split * p.k_blocks_per_split;
  const int nkb = max(0, min(p.k_blocks_per_split, total_kb - kb0));
  constexpr uint32_t TMEM_COLS = 2 * BN <= 32 ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : 256));

  pdl_trigger();
  int a_row0 = a_tile * TM;
  int b_row0 = b_tile * BN;
  int m_lim = SWAP ? p.M - b_row0 : 0;
  int ws_tile = a_tile;                       // row of the weight-scale table (MODE_BLOCK, SWAP)
  if constexpr (SWAP) {
    if (p.expert_offs) {
      pdl_wait();
      const int chunks = gridDim.y / p.n_experts;
      const int e = blockIdx.y / chunks, chunk = blockIdx.y % chunks;
      b_row0 = p.expert_offs[e] + chunk * BN;
      m_lim = p.expert_offs[e + 1] - b_row0;
      if (m_lim <= 0) return;
      a_row0 += e * p.N;
      ws_tile += e * ((p.N + TM - 1) / TM);
    }
  }
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
    if (lane == 0) {
      // ===== TMA producer.  Weights are read once (evict-first) and do not depend on the previous kernel: the first
      // ring of weight tiles is requested before griddepcontrol.wait (PDL), activations after it.
      const uint64_t pol_w = L2_EVICT_FIRST, pol_x = L2_EVICT_LAST;
      const void* tm_w = SWAP ? (const void*)&tmap_a : (const void*)&tmap_b;
      const void* tm_x = SWAP ? (const void*)&tmap_b : (const void*)&tmap_a;
      const int w_row = SWAP ? a_row0 : b_row0, x_row = SWAP ? b_row0 : a_row0;
      const int w_off = SWAP ? 0 : L::A_BYTES, x_off = SWAP ? L::A_BYTES : 0;
      const int pre = min(nkb, STAGES);
      for (int kb = 0; kb < pre; ++kb) {
        mbar_expect_tx(&full[kb], L::STAGE_BYTES);
        tma_load_2d_hint(stage_base + kb * L::STAGE_BYTES + w_off, tm_w, (kb0 + kb) * BK, w_row, &full[kb], pol_w);
      }
      pdl_wait();
      for (int kb = 0; kb < pre; ++kb)
        tma_load_2d_hint(stage_base + kb * L::STAGE_BYTES + x_off, tm_x, (kb0 + kb) * BK, x_row, &full[kb], pol_x);
      for (int kb = pre; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&empty[s], ((kb / STAGES) & 1) ^ 1);
        mbar_expect_tx(&full[s], L::STAGE_BYTES);
        const int kc = (kb0 + kb) * BK;
        tma_load_2d_hint(stage_base + s * L::STAGE_BYTES + w_off, tm_w, kc, w_row, &full[s], pol_w);
        tma_load_2d_hint(stage_base + s * L::STAGE_BYTES + x_off, tm_x, kc, x_row, &full[s], pol_x);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer
    const uint32_t idesc = umma_idesc_f8(TM, BN);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const int buf = BLOCK ? (kb & 1) : 0;
      if (BLOCK) {
        mbar_wait(&acc_empty[buf], ((kb >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
        tc_fence_after();
      }
      mbar_wait(&full[s], (kb / STAGES) & 1);
      tc_fence_after();
      if (lane == 0) {
        const unsigned char* a_src = stage_base + s * L::STAGE_BYTES;
        const uint64_t da = umma_desc_sw128(a_src), db = umma_desc_sw128(a_src + L::A_BYTES);
        const uint32_t d = tmem_base + (uint32_t)(buf * BN);
#pragma unroll
        for (int k = 0; k < BK / 32; ++k)     // UMMA_K = 32 fp8 = 32 B inside the swizzle span
          umma_f8(d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, BLOCK ? (uint32_t)(k != 0) : (uint32_t)((kb | k) != 0));
        umma_commit(&empty[s]);
        if (BLOCK || kb == nkb - 1) umma_commit(&acc_full[buf]);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue warps: TMEM lane quarter q, one accumulator row per thread
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    constexpr int CH = (BN % 32 == 0) ? 32 : 16;
    float acc[BN];
#pragma unroll
    for (int j = 0; j < BN; ++j) acc[j] = 0.f;
    pdl_wait();
    auto ld_chunk = [&](uint32_t taddr, float (&v)[CH]) {
      if constexpr (CH == 32) {
        uint32_t r[32];
        tmem_ld_32x32(taddr, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      } else {
        uint32_t r[16];
        tmem_ld_32x16(taddr, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
      }
    };
    if constexpr (BLOCK) {
      const int KBt = total_kb;
      // 16-byte scale loads need token offsets that are multiples of 4 (always true outside the grouped mode)
      const bool xs_vec = ((b_row0 & 3) == 0) && ((p.xs_stride & 3) == 0);
      const bool staged = SWAP && nkb <= sc_kb(BN);
      if (staged) {
        const int et = threadIdx.x - 64;
        for (int i = et; i < nkb * BN; i += EPI_THREADS) {
          const int kb = i / BN, j = i - kb * BN;
          const int kg = kb0 + kb;
          sc_s[i] = j < m_lim ? p.xs[(int64_t)kg * p.xs_stride + b_row0 + j] * p.ws[(int64_t)ws_tile * KBt + kg] : 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      // SWAP=0: both scales are per-thread scalars; the pair of the NEXT K-block is requested before waiting on this one
      float wsc_n = 0.f, xsc_n = 0.f;
      if (!SWAP && nkb > 0) {
        wsc_n = p.ws[(int64_t)b_tile * KBt + kb0];
        const int m = a_row0 + row;
        xsc_n = m < p.M ? p.xs[(int64_t)kb0 * p.xs_stride + m] : 0.f;
      }
      for (int kb = 0; kb < nkb; ++kb) {
        const int buf = kb & 1;
        const int kg = kb0 + kb;
        float wsc = 0.f, xsc_row = 0.f;
        if constexpr (SWAP) {
          if (!staged) wsc = p.ws[(int64_t)ws_tile * KBt + kg];
        } else {
          wsc = wsc_n; xsc_row = xsc_n;
          if (kb + 1 < nkb) {
            wsc_n = p.ws[(int64_t)b_tile * KBt + kg + 1];
            const int m = a_row0 + row;
            xsc_n = m < p.M ? p.xs[(int64_t)(kg + 1) * p.xs_stride + m] : 0.f;
          }
        }
        mbar_wait(&acc_full[buf], (kb >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < BN; c += CH) {
          float v[CH];
          ld_chunk(t_row + (uint32_t)(buf * BN + c), v);
          if (SWAP && staged) {
            const float4* sp = reinterpret_cast<const float4*>(sc_s + kb * BN + c);     // same address for every thread
#pragma unroll
            for (int j = 0; j < CH; j += 4) {
              const float4 s4 = sp[j >> 2];
              acc[c + j + 0] += v[j + 0] * s4.x; acc[c + j + 1] += v[j + 1] * s4.y;
              acc[c + j + 2] += v[j + 2] * s4.z; acc[c + j + 3] += v[j + 3] * s4.w;
            }
          } else if constexpr (SWAP) {
            // column = token: x scale varies along the chunk (same addresses for the whole warp → broadcast loads)
            const float* xp = p.xs + (int64_t)kg * p.xs_stride + b_row0 + c;
#pragma unroll
            for (int j = 0; j < CH; j += 4) {
              float4 s4;
              if (xs_vec && c + j + 4 <= m_lim) s4 = *reinterpret_cast<const float4*>(xp + j);
              else {
                s4.x = c + j + 0 < m_lim ? xp[j + 0] : 0.f; s4.y = c + j + 1 < m_lim ? xp[j + 1] : 0.f;
                s4.z = c + j + 2 < m_lim ? xp[j + 2] : 0.f; s4.w = c + j + 3 < m_lim ? xp[j + 3] : 0.f;
              }
              acc[c + j + 0] += v[j + 0] * (s4.x * wsc); acc[c + j + 1] += v[j + 1] * (s4.y * wsc);
              acc[c + j + 2] += v[j + 2] * (s4.z * wsc); acc[c + j + 3] += v[j + 3] * (s4.w * wsc);
            }
          } else {
            const float s = wsc * xsc_row;
#pragma unroll
            for (int j = 0; j < CH; ++j) acc[c + j] += v[j] * s;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
      }
    } else {
      if (nkb > 0) {
        mbar_wait(&acc_full[0], 0);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < BN; c += CH) {
          float v[CH];
          ld_chunk(t_row + (uint32_t)c, v);
#pragma unroll
          for (int j = 0; j < CH; ++j) acc[c + j] = v[j];
        }
      }
      // per-row / per-column scales are applied at the store (they commute with the split-K sum)
    }
    // ---- output
    T* y = reinterpret_cast<T*>(p.y);
    if (p.splits > 1) {
      float* red = reinterpret_cast<float*>(stage_base);          // ring is idle: all MMAs have retired
#pragma unroll
      for (int j = 0; j < BN; ++j) red[j * TM + row] = acc[j];
    } else if constexpr (SWAP) {
      const int n = a_tile * TM + row;
      if (n < p.N) {
        const float bias = p.bias ? to_f32(reinterpret_cast<const T*>(p.bias)[n]) : 0.f;
        const float wch = BLOCK ? 1.f : p.ws[p.ws_per_channel ? n : 0];
#pragma unroll
        for (int j = 0; j < BN; ++j) {
          if (j < m_lim) {
            float o = acc[j];
            if (!BLOCK) o *= wch * p.xs[p.xs_per_token ? b_row0 + j : 0];
            y[(size_t)(b_row0 + j) * p.N + n] = from_f32<T>(o + bias);
          }
        }
      }
    } else {
      const int m = a_row0 + row;
      if (m < p.M) {
        const int n0 = b_tile * BN;
        T* yr = y + (size_t)m * p.N + n0;
        const float xrow = BLOCK ? 1.f : p.xs[p.xs_per_token ? m : 0];
        const T* bp = p.bias ? reinterpret_cast<const T*>(p.bias) + n0 : nullptr;
#pragma unroll
        for (int j = 0; j < BN; j += 8) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int n = n0 + j + e;
            float s = xrow;
            if (!BLOCK) s *= p.ws[p.ws_per_channel ? min(n, p.N - 1) : 0];
            o[e] = acc[j + e] * (BLOCK ? 1.f : s) + (bp && n < p.N ? to_f32(bp[j + e]) : 0.f);
          }
          if (n0 + j + 8 <= p.N && (p.N % 8) == 0) {
            Vec16<T> ov;
#pragma unroll
            for (int e = 0; e < 8; ++e) ov.e[e] = from_f32<T>(o[e]);
            st16(yr + j, ov);
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) if (n0 + j + e < p.N) yr[j + e] = from_f32<T>(o[e]);
          }
        }
      }
    }
    tc_fence_before();
  }
  if constexpr (SWAP) {
    if (p.splits > 1) {
      // split-K over the cluster: every CTA reduces + stores 1/splits of the tile out of its peers' shared memory
      __syncwarp();
      cluster_sync();
      if (warp >= 2) {
        const int et = threadIdx.x - 64;
        T* y = reinterpret_cast<T*>(p.y);
        const int cols = min(BN, m_lim);
        const int units = cols * (TM / 4);
        const int per = (units + p.splits - 1) / p.splits;
        const int u_end = min(units, (split + 1) * per);
        const uint32_t red_s = smem_u32(stage_base);
        for (int u = split * per + et; u < u_end; u += EPI_THREADS) {
          const int col = u / (TM / 4), rq = u % (TM / 4);
          const uint32_t off = red_s + (uint32_t)(col * TM + rq * 4) * 4u;
          float4 a = ld_dsmem_f32x4(dsmem_addr(off, 0));
          for (int r = 1; r < p.splits; ++r) {
            const float4 b = ld_dsmem_f32x4(dsmem_addr(off, (uint32_t)r));
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
          }
          const int n = a_tile * TM + rq * 4;
          float o[4] = {a.x, a.y, a.z, a.w};
          const float xt = BLOCK ? 1.f : p.xs[p.xs_per_token ? b_row0 + col : 0];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (n + e < p.N) {
              if (!BLOCK) o[e] *= xt * p.ws[p.ws_per_channel ? n + e : 0];
              if (p.bias) o[e] += to_f32(reinterpret_cast<const T*>(p.bias)[n + e]);
            }
          }
          T* yr = y + (size_t)(b_row0 + col) * p.N + n;
          if (n + 4 <= p.N && (p.N & 3) == 0) {
            uint2 pk;
            pk.x = pack2<T>(o[0], o[1]); pk.y = pack2<T>(o[2], o[3]);
            *reinterpret_cast<uint2*>(yr) = pk;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) if (n + e < p.N) yr[e] = from_f32<T>(o[e]);
          }
        }
      }
      __syncwarp();
      cluster_sync();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------ host side
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
// row-major [rows, cols] byte matrix, box = [box_rows, 128 B], 128B swizzle
static int make_tmap(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t row_stride, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -10;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride};
  cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -11;
}

template <typename T, int BN, int STAGES, bool SWAP, bool BLOCK>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, dim3 grid, cudaStream_t st) {
  using L = Smem<BN>;
  const int smem = STAGES * L::STAGE_BYTES + (2 * STAGES + 4) * 8 + 32 + (SWAP && BLOCK ? sc_kb(BN) * BN * 4 : 0) + 1024;
  auto k = gemm_fp8_kernel<T, BN, STAGES, SWAP, BLOCK>;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
  if (grid.z > 1) vb_launch_cluster(k, grid, dim3(NUM_THREADS), (size_t)smem, st, grid.z, ta, tb, p);
  else vb_launch(k, grid, dim3(NUM_THREADS), (size_t)smem, st, ta, tb, p);
  return (int)cudaGetLastError();
}

template <typename T, bool BLOCK>
static int dispatch(const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, dim3 grid, int bn, bool swap, cudaStream_t st) {
  if (!swap) return launch<T, 128, 5, false, BLOCK>(ta, tb, p, grid, st);
  switch (bn) {
    case 16: return launch<T, 16, 8, true, BLOCK>(ta, tb, p, grid, st);
    case 32: return launch<T, 32, 8, true, BLOCK>(ta, tb, p, grid, st);
    case 48: return launch<T, 48, 8, true, BLOCK>(ta, tb, p, grid, st);
    case 64: return launch<T, 64, 8, true, BLOCK>(ta, tb, p, grid, st);
    case 96: return launch<T, 96, 6, true, BLOCK>(ta, tb, p, grid, st);
    case 128: return launch<T, 128, 5, true, BLOCK>(ta, tb, p, grid, st);
    default: return -2;
  }
}

}  // namespace gemm8

extern "C" int vb_gemm_max_active_clusters(int c);

// plan: swap (tokens on the MMA N axis) for M <= 128, token tile bn, cluster split-K factor
EXPORT int vb_gemm_fp8_plan(int M, int N, int K, int* swap, int* bn, int* splits) {
  using namespace gemm8;
  const int kb = (K + BK - 1) / BK;
  if (M <= 128) {
    int b = ((M + 15) / 16) * 16;
    if (b == 80) b = 96;
    if (b == 112) b = 128;
    const int nt = (N + TM - 1) / TM;
    int s = 1;
    for (int c = 8; c > 1; --c) {
      if (c * nt > 132 || kb / c < 4) continue;
      const int fit = vb_gemm_max_active_clusters(c);
      if (fit > 0 && nt > fit) continue;
      s = c;
      break;
    }
    if (nt >= 56 && kb < 16) s = 1;
    while ((kb + s - 1) / s > sc_kb(b) && s < 8 && (s + 1) * nt <= 132) ++s;     // staged scales cover sc_kb() K-blocks per CTA
    *swap = 1; *bn = b; *splits = s;
  } else {
    *swap = 0; *bn = 128; *splits = 1;
  }
  return 0;
}

// mode 0: block scales — xs [KB, xs_stride] fp32 (column major), ws [ceil(N/128), KB];
// mode 1: xs [M] (xs_per_token) or [1], ws [N] (ws_per_channel) or [1].
EXPORT int vb_gemm_fp8(const void* xq, const void* wq, const float* xs, const float* ws, const void* bias, void* y, int M, int N,
                       int K, int64_t x_stride, int64_t xs_stride, int mode, int xs_per_token, int ws_per_channel, int swap,
                       int bn, int splits, int dtype, cudaStream_t st) {
  using namespace gemm8;
  if (dtype == DT_F32 || K % 16 != 0) return -1;
  if (mode == 0 && (K % BK != 0)) return -1;
  CUtensorMap ta, tb;
  Params p;
  p.y = y; p.bias = bias; p.xs = xs; p.ws = ws; p.xs_stride = xs_stride; p.xs_per_token = xs_per_token;
  p.ws_per_channel = ws_per_channel; p.M = M; p.N = N; p.K = K; p.splits = splits; p.expert_offs = nullptr; p.n_experts = 0;
  const int kb = (K + BK - 1) / BK;
  p.k_blocks_per_split = (kb + splits - 1) / splits;
  int rc;
  dim3 grid;
  if (swap) {
    if ((rc = make_tmap(&ta, wq, N, K, K, TM))) return rc;
    if ((rc = make_tmap(&tb, xq, M, K, x_stride, bn))) return rc;
    grid = dim3((N + TM - 1) / TM, (M + bn - 1) / bn, splits);
  } else {
    if (splits != 1) return -2;
    if ((rc = make_tmap(&ta, xq, M, K, x_stride, TM))) return rc;
    if ((rc = make_tmap(&tb, wq, N, K, K, 128))) return rc;
    grid = dim3((M + TM - 1) / TM, (N + 127) / 128, 1);
  }
  if (dtype == DT_BF16)
    return mode == 0 ? dispatch<__nv_bfloat16, true>(ta, tb, p, grid, bn, swap, st)
                     : dispatch<__nv_bfloat16, false>(ta, tb, p, grid, bn, swap, st);
  return mode == 0 ? dispatch<__half, true>(ta, tb, p, grid, bn, swap, st) : dispatch<__half, false>(ta, tb, p, grid, bn, swap, st);
}

// Grouped (MoE) block-scaled GEMM: rows of xq sorted by expert (offs), wq [E, N, K] fp8, ws [E, ceil(N/128), KB].
EXPORT int vb_gemm_fp8_grouped(const void* xq, const void* wq, const float* xs, const float* ws, void* y, const int* offs,
                               int R, int N, int K, int E, int64_t xs_stride, int bn, int dtype, cudaStream_t st) {
  using namespace gemm8;
  if (dtype == DT_F32 || K % BK != 0 || bn > 128 || bn % 16 || bn == 80 || bn == 112) return -1;
  CUtensorMap ta, tb;
  Params p;
  p.y = y; p.bias = nullptr; p.xs = xs; p.ws = ws; p.xs_stride = xs_stride; p.xs_per_token = 1; p.ws_per_channel = 1;
  p.M = R; p.N = N; p.K = K; p.splits = 1; p.expert_offs = offs; p.n_experts = E;
  p.k_blocks_per_split = K / BK;
  int rc;
  if ((rc = make_tmap(&ta, wq, (uint64_t)E * N, K, K, TM))) return rc;
  if ((rc = make_tmap(&tb, xq, R, K, K, bn))) return rc;
  const int chunks = (R + bn - 1) / bn;
  dim3 grid((N + TM - 1) / TM, E * chunks, 1);
  if (dtype == DT_BF16) return dispatch<__nv_bfloat16, true>(ta, tb, p, grid, bn, true, st);
  return dispatch<__half, true>(ta, tb, p, grid, bn, true, st);
}
