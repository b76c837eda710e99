// bf16/fp16 GEMM on the 5th-gen tensor cores:  Y[M,N] = X[M,K] · W[N,K]^T (+bias)
//
// Hand-written tcgen05 pipeline (no CUTLASS): TMA (cp.async.bulk.tensor, 128B swizzle) →
// shared-memory ring → tcgen05.mma issued by one elected thread, fp32 accumulators in TMEM →
// tcgen05.ld epilogue.  Warp roles: 0 = TMA producer, 1 = MMA issuer + TMEM owner,
// 2..5 = epilogue (one per TMEM lane quarter).
//
// Two operand assignments share the kernel:
//   SWAP=1 (decode, M <= 256): A := W tile (128 output features), B := X (all M tokens, N_mma = M
//          padded to 16).  The accumulator is Y^T: TMEM lane = output feature, column = token.
//          Every CTA streams a [128 x K/splits] slab of the weight exactly once — the op is
//          weight-bandwidth bound, so the grid is (N/128) x splits CTAs sized to cover all 148 SMs,
//          the `splits` CTAs of a tile form a thread-block cluster: partials stay in shared memory and every
//          CTA reduces + stores 1/splits of the tile through DSMEM (no workspace, no atomics, graph safe).
//   SWAP=0 (prefill, large M): A := X tile (128 tokens), B := W tile (BN features).
//
// The reference uses cuBLASLt for this op (layers/linear.py:210-218); our kernel exists because it is
// the host for the fused epilogues (bias, SwiGLU, and in comm.cu the NVSwitch all-reduce + RMSNorm).
#include "common.cuh"

namespace gemm {

constexpr int BK = 64;            // K elements per stage (= 128 B swizzle span)
constexpr int TM = 128;           // UMMA M (rows of the A-operand tile)
constexpr int NUM_THREADS = 192;  // 6 warps
constexpr int EPI_THREADS = 128;

struct Params {
  void* y;                 // [M, N]
  const void* bias;        // [N] or null
  float* ws;               // split-K workspace [tiles][splits][BN_cols][128] fp32
  unsigned* counters;      // [tiles], zero-initialised, self-resetting
  int M, N, K;
  int k_blocks_per_split;  // BK blocks handled by one CTA
  int splits;
  int act;                 // 0 none, 1 silu-mul over interleaved (gate,up) feature pairs (SWAP only)
  int pf;                  // L2 prefetch ahead of the smem ring (SWAP)
  // grouped (MoE) mode, SWAP only: grid.y = n_experts * chunks; expert e multiplies the rows
  // [expert_offs[e], expert_offs[e+1]) of X (tokens sorted by expert) with W[e] = rows [e*N, (e+1)*N) of the weight
  const int* expert_offs;
  int n_experts;
  // grouped mode, fused MoE combine (+ cross-rank reduce-scatter): instead of storing Y, every output element is scaled
  // by its pair's routing weight and ADDED (red.add.f32, system scope, straight over NVLink) into the fp32 accumulator
  // row of the rank that OWNS the token (row owner = token / red_per, the same sharding the fused all-reduce kernels use
  // for the residual stream) — top-k combine + reduce-scatter inside the GEMM epilogue, one NVLink hop, no staging
  float* red_acc;          // this rank's accumulator [T, red_H] (also the single-GPU target)
  float* red_peer[8];      // the same buffer on every rank of the group (red_world > 1)
  const int* row_pair;     // sorted row → pair index i = token * k + slot
  const float* pair_w;     // [T * k] routing weights
  int red_k, red_H, red_world, red_per;
  long long* trace;        // debug: per-CTA clock64 stamps [cta][16] (null in production)
};

VB_DEVICE void red_add_f32(float* p, float v) { asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
VB_DEVICE void red_add_f32_sys(float* p, float v) {
  asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

template <int BN> struct SmemLayout {
  static constexpr int A_BYTES = TM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
};

// STAGES <= 4 is the two-CTAs-per-SM variant (smem <= 113 KB): used when the grid has more tiles than SMs, so the
// prologue / epilogue of one tile overlaps the weight stream of its neighbour instead of leaving the SM idle.
template <typename T, int BN, int STAGES, bool SWAP>
__global__ void __launch_bounds__(NUM_THREADS, STAGES <= 4 && BN <= 128 ? 2 : 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const Params p) {
  using L = SmemLayout<BN>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* stage_base = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * L::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile coordinates: A-operand tile index (x), B-operand tile index (y), K split (z)
  const int a_tile = blockIdx.x, b_tile = blockIdx.y, split = blockIdx.z;
  const int kb0 = split * p.k_blocks_per_split;
  const int total_kb = (p.K + BK - 1) / BK;
  const int nkb = max(0, min(p.k_blocks_per_split, total_kb - kb0));
  constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));

#define VB_TRACE(i) do { if (p.trace) p.trace[((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + (i)] = clock64(); } while (0)
  if (threadIdx.x == 0) VB_TRACE(0);
  pdl_trigger();
  int a_row0 = a_tile * TM;                 // first row of the A-operand tile in its tensor map
  int b_row0 = b_tile * BN;                 // first row of the B-operand tile
  int m_lim = SWAP ? p.M - b_row0 : 0;      // SWAP: number of real tokens (accumulator columns) of this tile
  if constexpr (SWAP) {
    if (p.expert_offs) {
      pdl_wait();                           // the offsets come from the routing kernels just before us
      const int chunks = gridDim.y / p.n_experts;
      const int e = blockIdx.y / chunks, chunk = blockIdx.y % chunks;
      b_row0 = p.expert_offs[e] + chunk * BN;
      m_lim = p.expert_offs[e + 1] - b_row0;
      if (m_lim <= 0) return;               // no token of this chunk routed to expert e: nothing to stream
      a_row0 += e * p.N;
    }
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
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
      VB_TRACE(1);
      // ===== TMA producer: weights are read exactly once → evict-first; activations are re-read by every CTA → evict-last
      // PDL: the weight operand does not depend on the preceding kernel, so the first ring of weight tiles is
      // requested BEFORE griddepcontrol.wait — the weight stream starts while the previous kernel is still
      // draining; only the activation operand waits for it.
      const int pre = SWAP ? min(nkb, STAGES) : 0;
      for (int kb = 0; kb < pre; ++kb) {
        unsigned char* a_dst = stage_base + kb * L::STAGE_BYTES;
        mbar_expect_tx(&full[kb], L::STAGE_BYTES);
        tma_load_2d_hint(a_dst, &tmap_a, (kb0 + kb) * BK, a_row0, &full[kb], L2_EVICT_FIRST);
      }
      // the smem ring alone (STAGES x 16 KB of weights) cannot cover the HBM latency-bandwidth product of one
      // SM's share when few CTAs stream a lot each: run an L2 prefetch PF k-blocks ahead of the ring
      constexpr int PF = SWAP ? 16 : 0;
      if (SWAP && p.pf)
        for (int kb = pre; kb < min(nkb, pre + PF); ++kb) tma_prefetch_l2_2d(&tmap_a, (kb0 + kb) * BK, a_row0);
      pdl_wait();
      VB_TRACE(2);
      for (int kb = 0; kb < pre; ++kb) {
        unsigned char* b_dst = stage_base + kb * L::STAGE_BYTES + L::A_BYTES;
        tma_load_2d_hint(b_dst, &tmap_b, (kb0 + kb) * BK, b_row0, &full[kb], L2_EVICT_LAST);
      }
      for (int kb = pre; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        if (SWAP && p.pf && kb + PF < nkb) tma_prefetch_l2_2d(&tmap_a, (kb0 + kb + PF) * BK, a_row0);
        mbar_wait(&empty[s], ((kb / STAGES) & 1) ^ 1);
        unsigned char* a_dst = stage_base + s * L::STAGE_BYTES;
        unsigned char* b_dst = a_dst + L::A_BYTES;
        mbar_expect_tx(&full[s], L::STAGE_BYTES);
        const int kc = (kb0 + kb) * BK;
        tma_load_2d_hint(a_dst, &tmap_a, kc, a_row0, &full[s], SWAP ? L2_EVICT_FIRST : L2_EVICT_LAST);
        tma_load_2d_hint(b_dst, &tmap_b, kc, b_row0, &full[s], SWAP ? L2_EVICT_LAST : L2_EVICT_FIRST);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer
    const uint32_t idesc = umma_idesc_f16(TM, BN, DTypeOf<T>::v == DT_BF16);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      mbar_wait(&full[s], (kb / STAGES) & 1);
      tc_fence_after();
      if (lane == 0) {
        if (kb == 0) VB_TRACE(3);
        const unsigned char* a_src = stage_base + s * L::STAGE_BYTES;
        const uint64_t da = umma_desc_sw128(a_src), db = umma_desc_sw128(a_src + L::A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)   // UMMA_K = 16 → advance 32 B inside the swizzle span
          umma_f16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        umma_commit(&empty[s]);                 // frees the smem slot once these MMAs retire
        if (kb == nkb - 1) { umma_commit(tmem_full); VB_TRACE(4); }
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue: warp w owns TMEM lanes 32*(w%4) .. +31; columns are processed in chunks of CH
    constexpr int CH = (BN % 32 == 0) ? 32 : 16;
    const int q = warp & 3;
    const int row = q * 32 + lane;            // index along the A-operand tile (0..127)
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    pdl_wait();   // outputs / workspace may alias buffers the previous kernel is still using
    if (nkb > 0) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
    }
    if (threadIdx.x == 64) VB_TRACE(5);
    auto load_chunk = [&](int c, float (&v)[CH]) {
      if (nkb == 0) {
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] = 0.f;
        return;
      }
      if constexpr (CH == 32) {
        uint32_t r[32];
        tmem_ld_32x32(t_row + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      } else {
        uint32_t r[16];
        tmem_ld_32x16(t_row + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
      }
    };
    const int tile_id = a_tile * gridDim.y + b_tile;
    bool do_store = true;
    if (p.splits > 1) {
      // split-K over a thread-block cluster: the `splits` CTAs of a tile are one cluster (grid z).  Each parks its
      // fp32 partial in its OWN shared memory (the operand ring is idle once tmem_full fired) — part 1 of 2, the
      // reduction itself follows the cluster barrier below.
      float* red = reinterpret_cast<float*>(stage_base);
      for (int c = 0; c < BN; c += CH) {
        float v[CH];
        load_chunk(c, v);
#pragma unroll
        for (int j = 0; j < CH; ++j) red[(c + j) * TM + row] = v[j];
      }
      if (threadIdx.x == 64) VB_TRACE(6);
      do_store = false;
    }
    if (do_store) {
      T* y = reinterpret_cast<T*>(p.y);
      for (int c = 0; c < BN; c += CH) {
        float v[CH];
        load_chunk(c, v);
        if (SWAP && p.red_acc) {
          // fused combine: acc[token(row), n] += w(row) * y[row, n]   (lanes = 32 consecutive n → 128-byte bursts)
          const int n = a_tile * TM + row;
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            if (c + j < m_lim) {
              const int pair = p.row_pair[b_row0 + c + j];
              const float val = v[j] * p.pair_w[pair];
              const int tok = pair / p.red_k;
              const size_t off = (size_t)tok * p.red_H + n;
              if (n < p.N) {
                if (p.red_world > 1) red_add_f32_sys(p.red_peer[tok / p.red_per] + off, val);
                else red_add_f32(p.red_acc + off, val);
              }
            }
          }
        } else if (SWAP && p.act) {
          // rows interleave (gate_i, up_i): even lane holds gate, odd lane up → y[tok, i] = silu(gate) * up
          const int n = a_tile * TM + row;
          const int m0 = b_row0 + c;
          const int n_out = p.N >> 1;
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const float other = __shfl_xor_sync(0xffffffffu, v[j], 1);
            const float g = v[j];
            const float r = (g / (1.f + __expf(-g))) * other;
            if ((lane & 1) == 0 && n < p.N && c + j < m_lim) y[(size_t)(m0 + j) * n_out + (n >> 1)] = from_f32<T>(r);
          }
        } else if (SWAP) {
          // accumulator is Y^T: row = output feature, column = token
          const int n = a_tile * TM + row;
          if (n < p.N) {
            const float bias = p.bias ? to_f32(reinterpret_cast<const T*>(p.bias)[n]) : 0.f;
            const int m0 = b_row0 + c;
#pragma unroll
            for (int j = 0; j < CH; ++j)
              if (c + j < m_lim) y[(size_t)(m0 + j) * p.N + n] = from_f32<T>(v[j] + bias);
          }
        } else {
          const int m = a_tile * TM + row;
          if (m < p.M) {
            const int n0 = b_tile * BN + c;
            T* yr = y + (size_t)m * p.N + n0;
            const T* bp = p.bias ? reinterpret_cast<const T*>(p.bias) + n0 : nullptr;
            if (n0 + CH <= p.N && (p.N % 8) == 0) {
#pragma unroll
              for (int j = 0; j < CH; j += 8) {
                Vec16<T> o;
#pragma unroll
                for (int e = 0; e < 8; ++e) o.e[e] = from_f32<T>(v[j + e] + (bp ? to_f32(bp[j + e]) : 0.f));
                st16(yr + j, o);
              }
            } else {
#pragma unroll
              for (int j = 0; j < CH; ++j)
                if (n0 + j < p.N) yr[j] = from_f32<T>(v[j] + (bp ? to_f32(bp[j]) : 0.f));
            }
          }
        }
      }
    }
    if (threadIdx.x == 64) VB_TRACE(9);
    tc_fence_before();
  }
  if constexpr (SWAP) {
    if (p.splits > 1) {
      // part 2: after the cluster barrier every CTA reduces 1/splits of the tile straight out of its peers'
      // shared memory (DSMEM) in split order (→ bit-identical run to run) and stores it.  No global workspace,
      // no atomics, no fences; the CTAs of a cluster are co-scheduled by hardware.
      __syncwarp();
      cluster_sync();
      if (warp >= 2) {
        const int et = threadIdx.x - 64;
        if (et == 0) VB_TRACE(7);
        T* y = reinterpret_cast<T*>(p.y);
        const int m_base = b_row0;
        const int cols = min(BN, m_lim);                           // real tokens in this tile
        const int units = cols * (TM / 4);                         // unit = 4 consecutive features of one token
        const int per = (units + p.splits - 1) / p.splits;
        const int u_end = min(units, (split + 1) * per);
        const uint32_t red_s = smem_u32(stage_base);
        // two units per pass: all DSMEM loads of both are in flight before the first add
        for (int u0 = split * per + et; u0 < u_end; u0 += 2 * EPI_THREADS) {
          float4 x[2][8];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int u = u0 + h * EPI_THREADS;
            if (u < u_end) {
              const uint32_t off = red_s + (uint32_t)((u / (TM / 4)) * TM + (u % (TM / 4)) * 4) * 4u;
#pragma unroll
              for (int r = 0; r < 8; ++r)
                if (r < p.splits) x[h][r] = ld_dsmem_f32x4(dsmem_addr(off, (uint32_t)r));
            }
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int u = u0 + h * EPI_THREADS;
            if (u >= u_end) break;
            const int col = u / (TM / 4), rq = u % (TM / 4);
            float4 acc = x[h][0];
#pragma unroll
            for (int r = 1; r < 8; ++r)
              if (r < p.splits) { acc.x += x[h][r].x; acc.y += x[h][r].y; acc.z += x[h][r].z; acc.w += x[h][r].w; }
            const int n = a_tile * TM + rq * 4;
            float o[4] = {acc.x, acc.y, acc.z, acc.w};
            if (p.act) {      // (gate, up) pairs are adjacent features: two outputs per unit
              if (n + 4 <= p.N) {
                T* ya = y + (size_t)(m_base + col) * (p.N >> 1) + (n >> 1);
                const uint32_t pk = pack2<T>((o[0] / (1.f + __expf(-o[0]))) * o[1], (o[2] / (1.f + __expf(-o[2]))) * o[3]);
                *reinterpret_cast<uint32_t*>(ya) = pk;
              }
              continue;
            }
            T* yr = y + (size_t)(m_base + col) * p.N + n;
            if (p.bias) {
#pragma unroll
              for (int e = 0; e < 4; ++e) if (n + e < p.N) o[e] += to_f32(reinterpret_cast<const T*>(p.bias)[n + e]);
            }
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
        if (et == 0) VB_TRACE(8);
      }
      __syncwarp();
      cluster_sync();      // nobody leaves (and frees its shared memory) while a peer may still be reading it
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
  if (threadIdx.x == 0) VB_TRACE(10);
#undef VB_TRACE
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

// row-major [rows, cols] 16-bit matrix, box = [box_rows, 64 cols], 128B swizzle
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

template <typename T, int BN, int STAGES, bool SWAP>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, dim3 grid, cudaStream_t st) {
  using L = SmemLayout<BN>;
  const int smem = STAGES * L::STAGE_BYTES + (2 * STAGES + 1) * 8 + 16 + 1024;
  auto k = gemm_kernel<T, BN, STAGES, SWAP>;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
  if (grid.z > 1) vb_launch_cluster(k, grid, dim3(NUM_THREADS), (size_t)smem, st, grid.z, ta, tb, p);
  else vb_launch(k, grid, dim3(NUM_THREADS), (size_t)smem, st, ta, tb, p);
  return (int)cudaGetLastError();
}

}  // namespace gemm

// how many clusters of `c` CTAs (1 CTA / SM, ~200 KB smem each) the device can hold at once
static int max_active_clusters(int c) {
  static int cache[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (c < 1 || c > 8) return 0;
  if (cache[c] == 0) {
    using namespace gemm;
    auto k = gemm_kernel<__nv_bfloat16, 64, 8, true>;
    const int smem = 8 * SmemLayout<64>::STAGE_BYTES + (2 * 8 + 1) * 8 + 16 + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1, 1, c); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = c;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, k, &cfg) != cudaSuccess) { cudaGetLastError(); n = -1; }
    cache[c] = n > 0 ? n : -1;
  }
  return cache[c];
}
EXPORT int vb_gemm_max_active_clusters(int c) { return max_active_clusters(c); }

static long long* g_trace = nullptr;
EXPORT void vb_gemm_tc_set_trace(long long* t) { g_trace = t; }

// Workspace requirement for vb_gemm_tc (bytes of fp32 partials, number of counters)
EXPORT int vb_gemm_tc_plan(int M, int N, int K, int num_sms, int* swap, int* bn, int* splits, int64_t* ws_bytes,
                           int* n_tiles) {
  using namespace gemm;
  const int kb = (K + BK - 1) / BK;
  if (M <= 256) {
    int b = ((M + 15) / 16) * 16;
    if (b > 128 && b < 256) b = 256;   // supported N_mma: 16..128 step 16, or 256
    const int mt = 1;
    const int nt = (N + TM - 1) / TM;
    // split-K factor = cluster size.  Measured on B200 (profiles/gemm_split_sweep.md): one CTA ingests ~57 GB/s
    // of weights, so few-tile GEMMs want as many CTAs as fit co-resident (clusters are GPC-local: the useful
    // ceiling is ~132 CTAs, and the occupancy query below is authoritative); a split costs ~2 us of epilogue,
    // which short-K GEMMs that already have >= 56 tiles do not win back.
    int s = 1;
    for (int c = 8; c > 1; --c) {
      if (c * nt * mt > 132 || kb / c < 4) continue;
      const int fit = max_active_clusters(c);
      if (fit > 0 && nt * mt > fit) continue;
      s = c;
      break;
    }
    if (nt * mt >= 56 && kb < 32) s = 1;
    *swap = 1; *bn = b; *splits = s; *n_tiles = nt * mt;
    *ws_bytes = 0;                        // split-K partials live in (distributed) shared memory
  } else {
    const int b = 128;
    const int mt = (M + TM - 1) / TM, nt = (N + b - 1) / b;
    *swap = 0; *bn = b; *splits = 1; *n_tiles = mt * nt; *ws_bytes = 0;
  }
  return 0;
}

EXPORT int vb_gemm_tc2(const void* x, const void* w, const void* bias, void* y, int M, int N, int K, int64_t x_stride,
                       int swap, int bn, int splits, float* ws, unsigned* counters, int act, int dtype, cudaStream_t st);
EXPORT int vb_gemm_tc(const void* x, const void* w, const void* bias, void* y, int M, int N, int K, int64_t x_stride,
                      int swap, int bn, int splits, float* ws, unsigned* counters, int dtype, cudaStream_t st) {
  return vb_gemm_tc2(x, w, bias, y, M, N, K, x_stride, swap, bn, splits, ws, counters, 0, dtype, st);
}
// act = 1: W rows interleave (gate_i, up_i); y is [M, N/2] = silu(gate) * up (swap-AB shapes only, N % 4 == 0, no bias)
EXPORT int vb_gemm_tc2(const void* x, const void* w, const void* bias, void* y, int M, int N, int K, int64_t x_stride,
                       int swap, int bn, int splits, float* ws, unsigned* counters, int act, int dtype, cudaStream_t st) {
  using namespace gemm;
  if (dtype == DT_F32 || K % 8 != 0) return -1;
  if (act && (!swap || bias || (N & 3))) return -1;
  const bool bf16 = dtype == DT_BF16;
  CUtensorMap ta, tb;
  Params p;
  p.trace = g_trace;
  p.expert_offs = nullptr; p.n_experts = 0; p.red_acc = nullptr; p.row_pair = nullptr; p.pair_w = nullptr; p.red_k = 1; p.red_H = 0; p.red_world = 1; p.red_per = 1;
  { static int pf = -1; if (pf < 0) { const char* e = std::getenv("VLLM_B200_GEMM_PF"); pf = (e && e[0] == '1') ? 1 : 0; } p.pf = pf; }
  p.y = y; p.bias = bias; p.ws = ws; p.counters = counters; p.M = M; p.N = N; p.K = K; p.splits = splits; p.act = act;
  const int kb = (K + BK - 1) / BK;
  p.k_blocks_per_split = (kb + splits - 1) / splits;
  int rc;
  if (swap) {
    if ((rc = make_tmap(&ta, w, N, K, K, TM, bf16))) return rc;
    if ((rc = make_tmap(&tb, x, M, K, x_stride, bn, bf16))) return rc;
    dim3 grid((N + TM - 1) / TM, (M + bn - 1) / bn, splits);
    static int n_sms = 0;
    if (!n_sms) { int d = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, d); }
    // multi-wave → 2 CTAs / SM.  VLLM_B200_GEMM_2CTA=1 forces the half-size ring for every decode GEMM: a CTA of the NEXT
    // kernel can then co-reside with a CTA of this one, so its launch + first-byte latency overlaps this kernel's epilogue
    static int force_two = -1;
    if (force_two < 0) { const char* e = std::getenv("VLLM_B200_GEMM_2CTA"); force_two = (e && e[0] == '1') ? 1 : 0; }
    const bool two = bn <= 128 && (force_two || (splits == 1 && (int)(grid.x * grid.y) > n_sms));
#define GO(BNV, ST, ST2)                                                                 \
    if (two) {                                                                           \
      if (bf16) return launch<__nv_bfloat16, BNV, ST2, true>(ta, tb, p, grid, st);       \
      else return launch<__half, BNV, ST2, true>(ta, tb, p, grid, st);                   \
    }                                                                                    \
    if (bf16) return launch<__nv_bfloat16, BNV, ST, true>(ta, tb, p, grid, st);          \
    else return launch<__half, BNV, ST, true>(ta, tb, p, grid, st);
    switch (bn) {
      case 16: GO(16, 8, 4)
      case 32: GO(32, 8, 4)
      case 48: GO(48, 8, 4)
      case 64: GO(64, 8, 4)
      case 80: GO(80, 6, 3)
      case 96: GO(96, 6, 3)
      case 112: GO(112, 6, 3)
      case 128: GO(128, 6, 3)
      case 256: GO(256, 4, 4)
      default: return -2;
    }
#undef GO
  } else {
    if ((rc = make_tmap(&ta, x, M, K, x_stride, TM, bf16))) return rc;
    if ((rc = make_tmap(&tb, w, N, K, K, bn, bf16))) return rc;
    dim3 grid((M + TM - 1) / TM, (N + bn - 1) / bn, 1);
    if (bn != 128) return -2;
    if (bf16) return launch<__nv_bfloat16, 128, 6, false>(ta, tb, p, grid, st);
    return launch<__half, 128, 6, false>(ta, tb, p, grid, st);
  }
}


// Grouped (MoE) GEMM: Y[r, :] = X[r, :] · W[e(r)]^T for rows sorted by expert; offs[e] .. offs[e+1] delimit expert e.
// X: [R, K], W: [E, N, K] contiguous, Y: [R, N].  bn = token-tile width (16..128), every expert gets
// ceil(R / bn) candidate chunks; chunks without tokens exit immediately (their weights are never read).
EXPORT int vb_gemm_tc_grouped_red(const void* x, const void* w, void* y, const int* offs, int R, int N, int K, int E, int bn,
                                  float* red_acc, void* const* red_peers, int64_t red_off, int red_world, int red_per,
                                  const int* row_pair, const float* pair_w, int red_k, int red_H, int dtype,
                                  cudaStream_t st);
EXPORT int vb_gemm_tc_grouped(const void* x, const void* w, void* y, const int* offs, int R, int N, int K, int E, int bn,
                              int dtype, cudaStream_t st) {
  return vb_gemm_tc_grouped_red(x, w, y, offs, R, N, K, E, bn, nullptr, nullptr, 0, 1, 1, nullptr, nullptr, 1, 0, dtype, st);
}
// red_acc != null: the epilogue adds w(row) * y[row, :] into acc[token(row), :] instead of storing y.  red_world > 1:
// the accumulator lives at byte offset red_off of every rank's symmetric buffer (red_peers) and a token's row is added
// on the rank that owns it (token / red_per).
EXPORT int vb_gemm_tc_grouped_red(const void* x, const void* w, void* y, const int* offs, int R, int N, int K, int E, int bn,
                                  float* red_acc, void* const* red_peers, int64_t red_off, int red_world, int red_per,
                                  const int* row_pair, const float* pair_w, int red_k, int red_H, int dtype,
                                  cudaStream_t st) {
  using namespace gemm;
  if (dtype == DT_F32 || K % 8 != 0 || bn > 128 || bn % 16) return -1;
  const bool bf16 = dtype == DT_BF16;
  CUtensorMap ta, tb;
  Params p;
  p.trace = nullptr; p.pf = 0;
  p.y = y; p.bias = nullptr; p.ws = nullptr; p.counters = nullptr; p.M = R; p.N = N; p.K = K; p.splits = 1; p.act = 0;
  p.expert_offs = offs; p.n_experts = E;
  p.red_acc = red_acc; p.row_pair = row_pair; p.pair_w = pair_w; p.red_k = red_k; p.red_H = red_H;
  p.red_world = red_world; p.red_per = red_per < 1 ? 1 : red_per;
  for (int i = 0; i < 8; ++i)
    p.red_peer[i] = (red_peers && i < red_world) ? reinterpret_cast<float*>((unsigned char*)red_peers[i] + red_off) : nullptr;
  const int kb = (K + BK - 1) / BK;
  p.k_blocks_per_split = kb;
  int rc;
  if ((rc = make_tmap(&ta, w, (uint64_t)E * N, K, K, TM, bf16))) return rc;
  if ((rc = make_tmap(&tb, x, R, K, K, bn, bf16))) return rc;
  const int chunks = (R + bn - 1) / bn;
  dim3 grid((N + TM - 1) / TM, E * chunks, 1);
#define GOG(BNV, ST2)                                                                    \
    if (bf16) return launch<__nv_bfloat16, BNV, ST2, true>(ta, tb, p, grid, st);         \
    else return launch<__half, BNV, ST2, true>(ta, tb, p, grid, st);
  switch (bn) {
    case 16: GOG(16, 4)
    case 32: GOG(32, 4)
    case 48: GOG(48, 4)
    case 64: GOG(64, 4)
    case 80: GOG(80, 3)
    case 96: GOG(96, 3)
    case 112: GOG(112, 3)
    case 128: GOG(128, 3)
    default: return -2;
  }
#undef GOG
}
