// Paged attention on the 5th-gen tensor cores (prefill + decode), head_dim 128, 64-token pages.
//
// Reference back-ends for this op: FlashInfer TRT-LLM-gen cubins / FA4 CuTe-DSL (v1/attention/backends/
// flashinfer.py:2239,2457).  This is our own kernel, designed around a cache layout chosen for tcgen05:
//   K page  [64 tokens, 128 dims]  (dims contiguous)      → K-major B operand of  S = Q·K^T
//   V page  [128 dims, 64 tokens]  (tokens contiguous)    → K-major B operand of  O = P·V
// so every operand is a plain 128B-swizzled K-major tile that one TMA box load drops into shared memory
// (page-table driven: the block id of each page becomes the TMA row coordinate), with no transposes.
//
// CTA = one (request, kv head) x 128 query rows (rows = q tokens x GQA group; decode uses G rows) x a KV
// range (split-KV for decode).  Warp roles: 0 TMA producer, 1 MMA issuer + TMEM owner, 2..5 softmax /
// correction / epilogue (one thread per query row = one TMEM lane, FA4 style: row max / sum are
// thread-local, no shuffles).  TMEM: S double-buffered (2 x 128 cols) + O (128 cols).
// Pipeline per 128-token KV tile j:  QK_j → softmax_j (two passes over S in TMEM, lazy O rescale) →
// P_j (bf16, swizzled smem) → PV_j accumulates into O;  QK_{j+1} is issued before PV_j.
#include "common.cuh"

namespace fmha {

constexpr int D = 128;
constexpr int BS = 64;              // tokens per page
constexpr int TILE_N = 128;         // kv tokens per iteration (2 pages)
constexpr int TILE_M = 128;         // query rows per CTA
constexpr int NUM_THREADS = 192;
constexpr int KV_STAGES = 2;
constexpr int MAX_TILES_SMEM = 1024;  // 2 block ids per tile cached in smem (covers 128k tokens)

constexpr int Q_BYTES = TILE_M * D * 2;            // 32 KB (2 k-blocks of [128 x 64])
constexpr int K_BYTES = TILE_N * D * 2;            // 32 KB per stage (2 k-blocks of [128 tokens x 64])
constexpr int V_BYTES = D * TILE_N * 2;            // 32 KB per stage (2 pages of [128 dims x 64 tokens])
constexpr int P_BYTES = TILE_M * TILE_N * 2;       // 32 KB (2 k-blocks of [128 rows x 64 tokens])

struct Params {
  const int* block_table;
  const int* seq_lens;
  const int* qsl;          // null → decode (1 token per request)
  void* out;               // [T, Hq, D]
  float* part_out;         // [T, Hq, S, D]
  float* part_lse;         // [T, Hq, S]  (log2 units)
  float* lse_out;          // optional [T, Hq] natural log
  int64_t out_stride;
  int bt_stride;
  int Hq, Hkv, G, TQ, num_splits;
  int window, causal;
  float scale_log2;
  int flat_B;              // > 0: grid.x enumerates the (request, query tile) pairs of `flat_B` requests (device-side
                           //      schedule from qsl: the launch geometry depends only on the bucket, not on the lengths)
};

VB_DEVICE void tma_load_3d(void* dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
VB_DEVICE void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
VB_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
VB_DEVICE void mbar_arrive_cnt(uint64_t* bar) { mbar_arrive(bar); }

template <typename T>
__global__ void __launch_bounds__(NUM_THREADS, 1)
fmha_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
            const __grid_constant__ CUtensorMap tmap_v, const Params p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* q_smem = smem;
  unsigned char* p_smem = q_smem + Q_BYTES;
  unsigned char* k_smem = p_smem + P_BYTES;                 // KV_STAGES x K_BYTES
  unsigned char* v_smem = k_smem + KV_STAGES * K_BYTES;     // KV_STAGES x V_BYTES
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_smem + KV_STAGES * V_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;                 // [KV_STAGES]
  uint64_t* v_full = k_full + KV_STAGES;       // [KV_STAGES]
  uint64_t* kv_empty = v_full + KV_STAGES;     // [KV_STAGES]
  uint64_t* s_full = kv_empty + KV_STAGES;     // [2]
  uint64_t* s_empty = s_full + 2;              // [2]
  uint64_t* p_full = s_empty + 2;
  uint64_t* pv_done = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);
  int* blk_ids = reinterpret_cast<int*>(tmem_slot + 2);   // [2 * MAX_TILES_SMEM]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int b = blockIdx.z;
  const int h = blockIdx.y;
  int split = 0, qt = 0;
  if (p.flat_B > 0) {
    // flat tile index → (request, tile): every warp walks the per-request tile counts 32 requests at a time (redundantly:
    // no shared memory or barrier before the early exit)
    int target = blockIdx.x, found_b = -1;
    for (int base = 0; base < p.flat_B && found_b < 0; base += 32) {
      const int r = base + lane;
      int nt = 0;
      if (r < p.flat_B) nt = (p.qsl[r + 1] - p.qsl[r] + p.TQ - 1) / p.TQ;
      int inc = nt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      const unsigned hit = __ballot_sync(0xffffffffu, target < inc);
      if (hit) {
        const int l = __ffs(hit) - 1;
        const int before = __shfl_sync(0xffffffffu, inc - nt, l);
        found_b = base + l;
        qt = target - before;
      } else {
        target -= __shfl_sync(0xffffffffu, inc, 31);
      }
    }
    if (found_b < 0) return;                       // more CTAs than tiles (bucket padding)
    b = found_b;
  } else if (p.qsl) {
    qt = blockIdx.x;
  } else {
    split = blockIdx.x;
  }
  const int seq_len = p.seq_lens[b];
  const int q0 = p.qsl ? p.qsl[b] : b;
  const int q_len = p.qsl ? p.qsl[b + 1] - q0 : 1;
  const int tok0 = qt * p.TQ;
  if (tok0 >= q_len) return;                       // uniform across the CTA
  const int n_tok = min(p.TQ, q_len - tok0);
  const int pos0 = seq_len - q_len + tok0;
  const int pos_last = pos0 + n_tok - 1;
  const int kv_end = p.causal ? pos_last + 1 : seq_len;
  const int kv_begin = p.window > 0 ? max(0, pos0 - p.window + 1) : 0;
  int tile_lo = kv_begin / TILE_N, tile_hi = (kv_end + TILE_N - 1) / TILE_N;
  {
    const int total = tile_hi - tile_lo;
    const int per = (total + p.num_splits - 1) / p.num_splits;
    tile_lo += split * per;
    tile_hi = min(tile_hi, tile_lo + per);
  }
  const int n_tiles = max(tile_hi - tile_lo, 0);
  const int n_pages_total = (kv_end + BS - 1) / BS;   // pages that hold valid tokens

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(&k_full[s], 1); mbar_init(&v_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 128); }
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  // zero the operand staging once: partially filled tiles must never expose NaN bit patterns to the MMA
  {
    uint4 z = make_uint4(0, 0, 0, 0);
    uint4* s4 = reinterpret_cast<uint4*>(smem);
    const int n16 = (Q_BYTES + P_BYTES + KV_STAGES * (K_BYTES + V_BYTES)) / 16;
    for (int i = threadIdx.x; i < n16; i += NUM_THREADS) s4[i] = z;
  }
  for (int i = threadIdx.x; i < 2 * n_tiles && i < 2 * MAX_TILES_SMEM; i += NUM_THREADS) {
    const int page = tile_lo * 2 + i;
    blk_ids[i] = page < n_pages_total ? p.block_table[(int64_t)b * p.bt_stride + page] : -1;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s0 = tmem_base, tmem_o = tmem_base + 256;

  if (n_tiles == 0) {
    // nothing to attend to (can happen for empty splits): emit a neutral partial
    if (warp >= 2 && p.num_splits > 1) {
      const int r = (warp & 3) * 32 + lane;
      const int tq = r / p.G, g = r % p.G;
      if (tq < n_tok) {
        const int64_t tok = q0 + tok0 + tq;
        const int64_t pi = (tok * p.Hq + (h * p.G + g)) * p.num_splits + split;
        for (int d = 0; d < D; ++d) p.part_out[pi * D + d] = 0.f;
        p.part_lse[pi] = -INFINITY;
      }
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
    return;
  }

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer
      mbar_expect_tx(q_full, 2 * p.G * p.TQ * 64 * 2);   // TMA always delivers the full box
      tma_load_3d(q_smem, &tmap_q, 0, h * p.G, q0 + tok0, q_full);
      tma_load_3d(q_smem + TILE_M * 128, &tmap_q, 64, h * p.G, q0 + tok0, q_full);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % KV_STAGES;
        mbar_wait(&kv_empty[s], ((j / KV_STAGES) & 1) ^ 1);
        int blk[2];
        blk[0] = j < MAX_TILES_SMEM ? blk_ids[2 * j] : -1;
        blk[1] = j < MAX_TILES_SMEM ? blk_ids[2 * j + 1] : -1;
        if (j >= MAX_TILES_SMEM) {
          for (int u = 0; u < 2; ++u) {
            const int page = (tile_lo + j) * 2 + u;
            blk[u] = page < n_pages_total ? p.block_table[(int64_t)b * p.bt_stride + page] : -1;
          }
        }
        const int n_valid = (blk[0] >= 0) + (blk[1] >= 0);
        unsigned char* ks = k_smem + s * K_BYTES;
        unsigned char* vs = v_smem + s * V_BYTES;
        mbar_expect_tx(&k_full[s], n_valid * BS * D * 2);
        mbar_expect_tx(&v_full[s], n_valid * BS * D * 2);
        for (int u = 0; u < 2; ++u) {
          if (blk[u] < 0) continue;
          const int chunk = (blk[u] * 2 + 0) * p.Hkv + h;       // K chunk index
          // K view: [chunks*BS rows, D cols]; k-block kb lands at kb*16KB + u*8KB
          tma_load_2d(ks + 0 * (TILE_N * 128) + u * (BS * 128), &tmap_k, 0, chunk * BS, &k_full[s]);
          tma_load_2d(ks + 1 * (TILE_N * 128) + u * (BS * 128), &tmap_k, 64, chunk * BS, &k_full[s]);
        }
        for (int u = 0; u < 2; ++u) {
          if (blk[u] < 0) continue;
          const int chunk = (blk[u] * 2 + 1) * p.Hkv + h;       // V chunk index
          // V view: [chunks*D rows, BS cols]; page u is k-block u of the PV product: [128 dims x 64 tokens]
          tma_load_2d(vs + u * (D * 128), &tmap_v, 0, chunk * D, &v_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer
    const uint32_t idesc = umma_idesc_f16(TILE_M, TILE_N, DTypeOf<T>::v == DT_BF16);
    mbar_wait(q_full, 0);
    for (int j = 0; j <= n_tiles; ++j) {
      if (j < n_tiles) {
        const int s = j % KV_STAGES, sb = j & 1;
        mbar_wait(&k_full[s], (j / KV_STAGES) & 1);
        mbar_wait(&s_empty[sb], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        if (lane == 0) {
          const unsigned char* ks = k_smem + s * K_BYTES;
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t da = umma_desc_sw128(q_smem + kb * (TILE_M * 128));
            const uint64_t db = umma_desc_sw128(ks + kb * (TILE_N * 128));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem_s0 + sb * 128, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          }
          umma_commit(&s_full[sb]);
        }
        __syncwarp();
      }
      if (j > 0) {
        const int jj = j - 1, s = jj % KV_STAGES;
        mbar_wait(&v_full[s], (jj / KV_STAGES) & 1);
        mbar_wait(p_full, jj & 1);
        tc_fence_after();
        if (lane == 0) {
          const unsigned char* vs = v_smem + s * V_BYTES;
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t da = umma_desc_sw128(p_smem + kb * (TILE_M * 128));
            const uint64_t db = umma_desc_sw128(vs + kb * (D * 128));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem_o, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (jj | kb | k) != 0);
          }
          umma_commit(&kv_empty[s]);
          umma_commit(pv_done);
        }
        __syncwarp();
      }
    }
  } else {
    // ===== softmax / correction / epilogue: thread = query row = TMEM lane
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t t_lane = (uint32_t)(qd * 32) << 16;
    const int tq = r / p.G, g = r % p.G;
    const bool row_valid = tq < n_tok && r < p.G * p.TQ;
    const int qpos = row_valid ? pos0 + tq : -1;
    float m_used = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const int sb = j & 1;
      const uint32_t t_s = tmem_s0 + sb * 128 + t_lane;
      const int tile_pos = (tile_lo + j) * TILE_N;
      mbar_wait(&s_full[sb], (j >> 1) & 1);
      tc_fence_after();
      // valid column range for this row: [lo, hi)
      int hi = min(kv_end, p.causal ? qpos + 1 : kv_end) - tile_pos;
      int lo = max(kv_begin, p.window > 0 ? qpos - p.window + 1 : 0) - tile_pos;
      if (!row_valid) hi = -1;
      // ---- pass 1: row max
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < TILE_N; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_s + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int col = c + i;
          const float sv = __uint_as_float(v[i]) * p.scale_log2;
          if (col < hi && col >= lo) mx = fmaxf(mx, sv);
        }
      }
      // lazy rescale (FA4): keep the old reference max unless the new one exceeds it by > 8 (2^8 headroom)
      float alpha = 1.f;
      bool rescale = false;
      if (mx > m_used + 8.f || (m_used == -INFINITY && mx > -INFINITY)) {
        alpha = (m_used == -INFINITY) ? 0.f : ex2(m_used - mx);
        rescale = (m_used != -INFINITY);
        m_used = mx;
      }
      const float m_ref = (m_used == -INFINITY) ? 0.f : m_used;
      // P (single smem buffer) and O may only be touched once PV_{j-1} has retired
      if (j > 0) { mbar_wait(pv_done, (j - 1) & 1); tc_fence_after(); }
      if (rescale) {
        l_run *= alpha;
#pragma unroll
        for (int c = 0; c < D; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_o + t_lane + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st_32x32(tmem_o + t_lane + c, v);
        }
        tmem_st_wait();
      }
      // ---- pass 2: P = exp2(s - m), row sum, bf16 → swizzled K-major smem
      float lsum = 0.f;
#pragma unroll
      for (int c = 0; c < TILE_N; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_s + c, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const int col = c + i;
          float p0 = (col < hi && col >= lo) ? ex2(__uint_as_float(v[i]) * p.scale_log2 - m_ref) : 0.f;
          float p1 = (col + 1 < hi && col + 1 >= lo) ? ex2(__uint_as_float(v[i + 1]) * p.scale_log2 - m_ref) : 0.f;
          lsum += p0 + p1;
          pk[i / 2] = pack2<T>(p0, p1);
        }
        // 32 columns = 4 chunks of 16 B inside k-block (c / 64)
        unsigned char* prow = p_smem + (c / 64) * (TILE_M * 128) + r * 128;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int chunk = ((c % 64) / 8 + ch) ^ (r & 7);
          *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
        }
      }
      l_run += lsum;
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(&s_empty[sb]);
      mbar_arrive(p_full);
    }
    // ---- epilogue
    mbar_wait(pv_done, (n_tiles - 1) & 1);
    tc_fence_after();
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    const int64_t tok = q0 + tok0 + tq;
    const int hq = h * p.G + g;
    if (p.num_splits == 1) {
      T* orow = reinterpret_cast<T*>(p.out) + tok * p.out_stride + (int64_t)hq * D;
#pragma unroll
      for (int c = 0; c < D; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_o + t_lane + c, v);
        tmem_ld_wait();
        if (row_valid) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 o;
            o.x = pack2<T>(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
            o.y = pack2<T>(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
            o.z = pack2<T>(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
            o.w = pack2<T>(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + c + i) = o;
          }
        }
      }
      if (row_valid && p.lse_out)
        p.lse_out[tok * p.Hq + hq] = l_run > 0.f ? (m_used + log2f(l_run)) * 0.6931471805599453f : -INFINITY;
    } else {
      const int64_t pi = (tok * p.Hq + hq) * p.num_splits + split;
      float* orow = p.part_out + pi * D;
#pragma unroll
      for (int c = 0; c < D; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_o + t_lane + c, v);
        tmem_ld_wait();
        if (row_valid) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(orow + c + i) =
                make_float4(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv,
                            __uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
        }
      }
      if (row_valid) p.part_lse[pi] = l_run > 0.f ? m_used + log2f(l_run) : -INFINITY;
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <typename T, int DD>
__global__ void __launch_bounds__(DD) split_reduce(const float* __restrict__ part_out, const float* __restrict__ part_lse,
                                                   T* __restrict__ out, float* __restrict__ lse_out, int Hq, int S,
                                                   int64_t out_stride) {
  const int tok = blockIdx.x, hq = blockIdx.y, d = threadIdx.x;
  const int64_t base = ((int64_t)tok * Hq + hq) * S;
  float m = -INFINITY;
  for (int s = 0; s < S; ++s) m = fmaxf(m, part_lse[base + s]);
  const float m_use = (m == -INFINITY) ? 0.f : m;
  float den = 0.f, o = 0.f;
  for (int s = 0; s < S; ++s) {
    const float wgt = ex2(part_lse[base + s] - m_use);
    den += wgt;
    o += wgt * part_out[(base + s) * DD + d];
  }
  out[(int64_t)tok * out_stride + (int64_t)hq * DD + d] = from_f32<T>(den > 0.f ? o / den : 0.f);
  if (lse_out && d == 0) lse_out[(int64_t)tok * Hq + hq] = den > 0.f ? (m + log2f(den)) * 0.6931471805599453f : -INFINITY;
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

}  // namespace fmha

// kv_cache: [nb, 2, Hkv, BS*D] with K pages [BS, D] and V pages stored transposed [D, BS].
EXPORT int vb_fmha_tc2(const void* q, const void* kv_cache, int64_t num_blocks, const int* block_table,
                       const int* seq_lens, const int* qsl, void* out, float* part_out, float* part_lse,
                       float* lse_out, int B, int T_total, int max_q_len, int Hq, int Hkv, int64_t q_stride,
                       int64_t out_stride, int bt_stride, int num_splits, int window, float scale, int causal,
                       int flat, int dtype, cudaStream_t st);
EXPORT int vb_fmha_tc(const void* q, const void* kv_cache, int64_t num_blocks, const int* block_table,
                      const int* seq_lens, const int* qsl, void* out, float* part_out, float* part_lse,
                      float* lse_out, int B, int T_total, int max_q_len, int Hq, int Hkv, int64_t q_stride,
                      int64_t out_stride, int bt_stride, int num_splits, int window, float scale, int causal,
                      int dtype, cudaStream_t st) {
  return vb_fmha_tc2(q, kv_cache, num_blocks, block_table, seq_lens, qsl, out, part_out, part_lse, lse_out, B, T_total,
                     max_q_len, Hq, Hkv, q_stride, out_stride, bt_stride, num_splits, window, scale, causal, 0, dtype, st);
}
// flat != 0 (varlen only): the grid is sized from (T_total, B) alone — T_total / TQ + B tiles bound every split of
// T_total tokens over B requests — so one captured launch serves any mixed batch of that bucket.
EXPORT int vb_fmha_tc2(const void* q, const void* kv_cache, int64_t num_blocks, const int* block_table,
                       const int* seq_lens, const int* qsl, void* out, float* part_out, float* part_lse,
                       float* lse_out, int B, int T_total, int max_q_len, int Hq, int Hkv, int64_t q_stride,
                       int64_t out_stride, int bt_stride, int num_splits, int window, float scale, int causal,
                       int flat, int dtype, cudaStream_t st) {
  using namespace fmha;
  if (B == 0) return 0;
  if (dtype == DT_F32) return -1;
  const int G = Hq / Hkv;
  if (G > 128 || 128 % G != 0) return -4;
  auto enc = get_encode();
  if (!enc) return -10;
  const bool bf16 = dtype == DT_BF16;
  const CUtensorMapDataType dt = bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  Params p;
  p.block_table = block_table; p.seq_lens = seq_lens; p.qsl = qsl; p.out = out; p.part_out = part_out;
  p.part_lse = part_lse; p.lse_out = lse_out; p.out_stride = out_stride; p.bt_stride = bt_stride; p.Hq = Hq;
  p.Hkv = Hkv; p.G = G; p.TQ = qsl ? TILE_M / G : 1; p.num_splits = qsl ? 1 : num_splits; p.window = window;
  p.causal = causal; p.scale_log2 = scale * 1.4426950408889634f;
  p.flat_B = (flat && qsl) ? B : 0;
  CUtensorMap tq, tk, tv;
  {  // q: [T, Hq, D] with token stride q_stride; box {64, G, TQ}
    cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)Hq, (cuuint64_t)T_total};
    cuuint64_t strides[2] = {(cuuint64_t)D * 2, (cuuint64_t)q_stride * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)G, (cuuint32_t)p.TQ};
    cuuint32_t es[3] = {1, 1, 1};
    if (enc(&tq, dt, 3, const_cast<void*>(q), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return -11;
  }
  {  // K view: [nb*2*Hkv*BS, D], box {64 cols, 64 rows}
    cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)num_blocks * 2 * Hkv * BS};
    cuuint64_t strides[1] = {(cuuint64_t)D * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)BS};
    cuuint32_t es[2] = {1, 1};
    if (enc(&tk, dt, 2, const_cast<void*>(kv_cache), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return -12;
  }
  {  // V view: [nb*2*Hkv*D, BS], box {64 cols (tokens), 128 rows (dims)}
    cuuint64_t dims[2] = {(cuuint64_t)BS, (cuuint64_t)num_blocks * 2 * Hkv * D};
    cuuint64_t strides[1] = {(cuuint64_t)BS * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)D};
    cuuint32_t es[2] = {1, 1};
    if (enc(&tv, dt, 2, const_cast<void*>(kv_cache), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return -13;
  }
  const int smem = Q_BYTES + P_BYTES + KV_STAGES * (K_BYTES + V_BYTES) + 32 * 8 + 16 + 2 * MAX_TILES_SMEM * 4 + 1024;
  dim3 grid = qsl ? dim3((max_q_len + p.TQ - 1) / p.TQ, Hkv, B) : dim3(p.num_splits, Hkv, B);
  if (p.flat_B) grid = dim3(T_total / p.TQ + B, Hkv, 1);
  if (bf16) {
    auto k = fmha_kernel<__nv_bfloat16>;
    static bool a = false;
    if (!a) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); a = true; }
    k<<<grid, NUM_THREADS, smem, st>>>(tq, tk, tv, p);
  } else {
    auto k = fmha_kernel<__half>;
    static bool a = false;
    if (!a) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); a = true; }
    k<<<grid, NUM_THREADS, smem, st>>>(tq, tk, tv, p);
  }
  int rc = (int)cudaGetLastError();
  if (rc) return rc;
  if (!qsl && p.num_splits > 1) {
    if (bf16) split_reduce<__nv_bfloat16, D><<<dim3(B, Hq), D, 0, st>>>(part_out, part_lse, (__nv_bfloat16*)out, lse_out, Hq, p.num_splits, out_stride);
    else split_reduce<__half, D><<<dim3(B, Hq), D, 0, st>>>(part_out, part_lse, (__half*)out, lse_out, Hq, p.num_splits, out_stride);
  }
  return (int)cudaGetLastError();
}
