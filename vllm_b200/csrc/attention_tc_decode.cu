// Decode attention on tcgen05, "swap-AB" form: KV tokens ride the MMA M axis.
//
//   S^T[tok, g] = K[tok, :] · Q[g, :]^T        A = K tile  [128 tokens x 128 dims]   (K page layout, K-major)
//                                              B = Q       [16 (GQA group, padded) x 128 dims]
//   O^T[d, g]  += V^T[d, tok] · P[g, tok]^T    A = V^T tile [128 dims x 128 tokens]  (tensor-core cache layout)
//                                              B = P       [16 x 128 tokens]
// Both MMAs run with M = 128 and N = 16, the accumulators need 48 TMEM columns, and the per-tile softmax is
// O(G) work per thread (thread = token): one 16-column tcgen05.ld, a warp-shuffle + one named barrier for the
// column max, an exp2 per group member and a 2-byte store of P.  The serial chain per 128-token tile is a few
// hundred cycles, so with a 3-stage (192 KB) TMA ring the kernel is bound by HBM, not by softmax latency —
// the M = 128-query-row kernel in attention_tc.cu (used for prefill) wastes 124 of 128 rows on decode and its
// 128-column softmax chain caps a CTA at ~20 GB/s.
//
// Reference for the op: FlashInfer trtllm_batch_decode_with_kv_cache (v1/attention/backends/flashinfer.py:2457).
#include "common.cuh"

namespace fmha_dec {

constexpr int D = 128, BS = 64, TILE = 128, NQ = 16;
constexpr int NUM_THREADS = 192;
constexpr int STAGES = 3;
constexpr int K_BYTES = TILE * D * 2;   // 32 KB: 2 k-blocks (dims 0-63 / 64-127) of [128 tokens x 64]
constexpr int V_BYTES = D * TILE * 2;   // 32 KB: 2 pages of [128 dims x 64 tokens]
constexpr int Q_BYTES = NQ * D * 2;     // 4 KB: 2 k-blocks of [16 x 64]
constexpr int P_BYTES = NQ * TILE * 2;  // 4 KB: 2 k-blocks of [16 x 64 tokens]
constexpr int MAX_TILES_SMEM = 1024;

struct Params {
  const int* block_table;
  const int* seq_lens;
  void* out;
  float* part_out;
  float* part_lse;
  float* lse_out;
  int64_t out_stride;
  int bt_stride;
  int Hq, Hkv, G, num_splits, window, B;
  float scale_log2;
};

VB_DEVICE void tma_load_3d(void* dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
VB_DEVICE void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
VB_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
VB_DEVICE void bar_sync_softmax() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// Persistent: grid = min(work items, SMs); a work item is (kv split, kv head, sequence).  Barriers, TMEM and the zeroed
// shared memory are set up once per CTA; the three roles walk the same item list independently, tied together only by
// the mbarrier phases (a CTA-global tile counter), so the TMA producer runs up to STAGES tiles ahead ACROSS item
// boundaries: the K/V stream of the next (sequence, head) is already in flight while the softmax / epilogue of the
// current one finish.  (The one-item-per-CTA version paid ~3.5 us of setup + pipeline fill per 11 us of streaming.)
template <typename T, int G>
__global__ void __launch_bounds__(NUM_THREADS, 1)
fmha_decode_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                   const __grid_constant__ CUtensorMap tmap_v, const Params p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* k_smem = smem;                                 // STAGES x 32 KB
  unsigned char* v_smem = k_smem + STAGES * K_BYTES;            // STAGES x 32 KB
  unsigned char* q_smem = v_smem + STAGES * V_BYTES;            // 2 x 4 KB (double-buffered across items)
  unsigned char* p_smem = q_smem + 2 * Q_BYTES;                 // 4 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_smem + P_BYTES);
  uint64_t* q_full = bars;                 // [2]
  uint64_t* q_empty = bars + 2;            // [2]
  uint64_t* k_full = q_empty + 2;
  uint64_t* v_full = k_full + STAGES;
  uint64_t* kv_empty = v_full + STAGES;
  uint64_t* s_full = kv_empty + STAGES;    // [2]
  uint64_t* s_empty = s_full + 2;          // [2]
  uint64_t* p_full = s_empty + 2;
  uint64_t* pv_done = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);
  float* wred = reinterpret_cast<float*>(tmem_slot + 2);        // [4 warps][NQ] column max / sum exchange
  int* blk_ids = reinterpret_cast<int*>(wred + 4 * NQ);         // [2 * MAX_TILES_SMEM] (producer warp only)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  pdl_wait();   // q, the new K/V rows and seq_lens come from the preceding kernels
  const int n_items = p.num_splits * p.Hkv * p.B;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
    for (int s = 0; s < 2; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); }
    for (int s = 0; s < STAGES; ++s) { mbar_init(&k_full[s], 1); mbar_init(&v_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 128); }
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 64);
  {
    // Q / P rows beyond the group and the partially filled last tile must be finite: zero once
    uint4 z = make_uint4(0, 0, 0, 0);
    uint4* s4 = reinterpret_cast<uint4*>(smem);
    const int n16 = (STAGES * (K_BYTES + V_BYTES) + 2 * Q_BYTES + P_BYTES) / 16;
    for (int i = threadIdx.x; i < n16; i += NUM_THREADS) s4[i] = z;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 32;   // S^T buffers at columns 0..15 and 16..31, O^T at 32..47

  // the KV tile range of a work item (identical arithmetic in every role)
  auto item_range = [&](int it, int& split, int& h, int& b, int& tile_lo, int& kv_begin, int& kv_end) -> int {
    split = it % p.num_splits;
    const int rest = it / p.num_splits;
    h = rest % p.Hkv;
    b = rest / p.Hkv;
    const int seq_len = p.seq_lens[b];
    kv_end = seq_len;
    kv_begin = p.window > 0 ? max(0, seq_len - p.window) : 0;
    tile_lo = kv_begin / TILE;
    int tile_hi = (kv_end + TILE - 1) / TILE;
    const int per = (tile_hi - tile_lo + p.num_splits - 1) / p.num_splits;
    tile_lo += split * per;
    tile_hi = min(tile_hi, tile_lo + per);
    return max(tile_hi - tile_lo, 0);
  };

  if (warp == 0) {
    // ===== TMA producer (the whole warp stages the item's block ids, lane 0 issues the copies)
    int gt = 0, qi = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      int split, h, b, tile_lo, kv_begin, kv_end;
      const int n_tiles = item_range(it, split, h, b, tile_lo, kv_begin, kv_end);
      if (n_tiles == 0) continue;
      const int n_pages_total = (kv_end + BS - 1) / BS;
      __syncwarp();
      for (int i = lane; i < 2 * n_tiles && i < 2 * MAX_TILES_SMEM; i += 32) {
        const int page = tile_lo * 2 + i;
        blk_ids[i] = page < n_pages_total ? p.block_table[(int64_t)b * p.bt_stride + page] : -1;
      }
      __syncwarp();
      if (lane == 0) {
        const int qb = qi & 1;
        mbar_wait(&q_empty[qb], ((qi >> 1) & 1) ^ 1);
        unsigned char* qs = q_smem + qb * Q_BYTES;
        mbar_expect_tx(&q_full[qb], 2 * G * 64 * 2);
        tma_load_3d(qs, &tmap_q, 0, h * G, b, &q_full[qb]);
        tma_load_3d(qs + NQ * 128, &tmap_q, 64, h * G, b, &q_full[qb]);
        for (int j = 0; j < n_tiles; ++j) {
          const int g = gt + j, s = g % STAGES;
          mbar_wait(&kv_empty[s], ((g / STAGES) & 1) ^ 1);
          int blk[2];
          for (int u = 0; u < 2; ++u) {
            if (j < MAX_TILES_SMEM) blk[u] = blk_ids[2 * j + u];
            else {
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
            const int chunk = (blk[u] * 2 + 0) * p.Hkv + h;
            tma_load_2d_hint(ks + 0 * (TILE * 128) + u * (BS * 128), &tmap_k, 0, chunk * BS, &k_full[s], L2_EVICT_FIRST);
            tma_load_2d_hint(ks + 1 * (TILE * 128) + u * (BS * 128), &tmap_k, 64, chunk * BS, &k_full[s], L2_EVICT_FIRST);
          }
          for (int u = 0; u < 2; ++u) {
            if (blk[u] < 0) continue;
            const int chunk = (blk[u] * 2 + 1) * p.Hkv + h;
            tma_load_2d_hint(vs + u * (D * 128), &tmap_v, 0, chunk * D, &v_full[s], L2_EVICT_FIRST);
          }
        }
      }
      gt += n_tiles;
      ++qi;
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_f16(128, NQ, DTypeOf<T>::v == DT_BF16);
    int gt = 0, qi = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      int split, h, b, tile_lo, kv_begin, kv_end;
      const int n_tiles = item_range(it, split, h, b, tile_lo, kv_begin, kv_end);
      if (n_tiles == 0) continue;
      const int qb = qi & 1;
      const unsigned char* qs = q_smem + qb * Q_BYTES;
      mbar_wait(&q_full[qb], (qi >> 1) & 1);
      for (int j = 0; j <= n_tiles; ++j) {
        if (j < n_tiles) {
          const int g = gt + j, s = g % STAGES, sb = g & 1;
          mbar_wait(&k_full[s], (g / STAGES) & 1);
          mbar_wait(&s_empty[sb], ((g >> 1) & 1) ^ 1);
          tc_fence_after();
          if (lane == 0) {
            const unsigned char* ks = k_smem + s * K_BYTES;
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
              const uint64_t da = umma_desc_sw128(ks + kb * (TILE * 128));
              const uint64_t db = umma_desc_sw128(qs + kb * (NQ * 128));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16(tmem_base + sb * NQ, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
            }
            umma_commit(&s_full[sb]);
            if (j == n_tiles - 1) umma_commit(&q_empty[qb]);      // every QK of this item has read Q
          }
          __syncwarp();
        }
        if (j > 0) {
          const int jj = j - 1, g = gt + jj, s = g % STAGES;
          mbar_wait(&v_full[s], (g / STAGES) & 1);
          mbar_wait(p_full, g & 1);
          tc_fence_after();
          if (lane == 0) {
            const unsigned char* vs = v_smem + s * V_BYTES;
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
              const uint64_t da = umma_desc_sw128(vs + kb * (D * 128));
              const uint64_t db = umma_desc_sw128(p_smem + kb * (NQ * 128));
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
      gt += n_tiles;
      ++qi;
    }
  } else {
    // ===== softmax: thread = kv token of the tile (TMEM lane); later thread = output dim
    const int qd = warp & 3;
    const int t = qd * 32 + lane;
    const uint32_t t_lane = (uint32_t)(qd * 32) << 16;
    int gt = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      int split, h, b, tile_lo, kv_begin, kv_end;
      const int n_tiles = item_range(it, split, h, b, tile_lo, kv_begin, kv_end);
      if (n_tiles == 0) {
        if (p.num_splits > 1) {
          for (int g = 0; g < G; ++g) {
            const int64_t pi = ((int64_t)b * p.Hq + (h * G + g)) * p.num_splits + split;
            p.part_out[pi * D + t] = 0.f;
            if (t == 0) p.part_lse[pi] = -INFINITY;
          }
        }
        continue;
      }
      float m_used[G], l_thr[G];
#pragma unroll
      for (int g = 0; g < G; ++g) { m_used[g] = -INFINITY; l_thr[g] = 0.f; }
      for (int j = 0; j < n_tiles; ++j) {
        const int gj = gt + j, sb = gj & 1;
        mbar_wait(&s_full[sb], (gj >> 1) & 1);
        tc_fence_after();
        uint32_t sv[16];
        tmem_ld_32x16(tmem_base + sb * NQ + t_lane, sv);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&s_empty[sb]);                 // S^T is in registers: the buffer can take QK_{j+2}
        const int kpos = (tile_lo + j) * TILE + t;
        const bool ok = kpos < kv_end && kpos >= kv_begin;
        float s[G], cmax[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
          s[g] = ok ? __uint_as_float(sv[g]) * p.scale_log2 : -INFINITY;
          cmax[g] = warp_max(s[g]);
        }
        if (lane == 0) {
#pragma unroll
          for (int g = 0; g < G; ++g) wred[qd * NQ + g] = cmax[g];
        }
        bar_sync_softmax();
        float alpha[G];
        bool rescale = false;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float mx = fmaxf(fmaxf(wred[g], wred[NQ + g]), fmaxf(wred[2 * NQ + g], wred[3 * NQ + g]));
          alpha[g] = 1.f;
          if (mx > m_used[g] + 8.f || (m_used[g] == -INFINITY && mx > -INFINITY)) {
            if (m_used[g] != -INFINITY) { alpha[g] = ex2(m_used[g] - mx); rescale = true; }
            m_used[g] = mx;
          }
        }
        // P and O^T may only be touched once the previous PV (of this or of the previous item) has retired
        if (gj > 0) { mbar_wait(pv_done, (gj - 1) & 1); tc_fence_after(); }
        if (rescale) {   // uniform across the CTA: every thread sees the same column maxima
          uint32_t ov[16];
          tmem_ld_32x16(tmem_o + t_lane, ov);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < G; ++g) {
            ov[g] = __float_as_uint(__uint_as_float(ov[g]) * alpha[g]);
            l_thr[g] *= alpha[g];
          }
          tmem_st_32x16(tmem_o + t_lane, ov);
          tmem_st_wait();
        }
        unsigned char* pcol = p_smem + (t >> 6) * (NQ * 128) + ((t & 63) & 7) * 2;
        const int chunk = (t & 63) >> 3;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float m_ref = (m_used[g] == -INFINITY) ? 0.f : m_used[g];
          const float pe = ok ? ex2(s[g] - m_ref) : 0.f;
          l_thr[g] += pe;
          *reinterpret_cast<T*>(pcol + g * 128 + ((chunk ^ (g & 7)) << 4)) = from_f32<T>(pe);
        }
        tc_fence_before();
        fence_proxy_async();
        mbar_arrive(p_full);
        bar_sync_softmax();                         // wred is reused by the next tile
      }
      // ---- epilogue of the item: column sums over the 128 threads, then thread = output dim d
      float l_tot[G];
#pragma unroll
      for (int g = 0; g < G; ++g) l_tot[g] = warp_sum(l_thr[g]);
      if (lane == 0) {
#pragma unroll
        for (int g = 0; g < G; ++g) wred[qd * NQ + g] = l_tot[g];
      }
      bar_sync_softmax();
      mbar_wait(pv_done, (gt + n_tiles - 1) & 1);
      tc_fence_after();
      uint32_t ov[16];
      tmem_ld_32x16(tmem_o + t_lane, ov);
      tmem_ld_wait();
      tc_fence_before();
      const int d = t;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float l = wred[g] + wred[NQ + g] + wred[2 * NQ + g] + wred[3 * NQ + g];
        const float inv = l > 0.f ? 1.f / l : 0.f;
        const float o = __uint_as_float(ov[g]) * inv;
        const int hq = h * G + g;
        if (p.num_splits == 1) {
          reinterpret_cast<T*>(p.out)[(int64_t)b * p.out_stride + (int64_t)hq * D + d] = from_f32<T>(o);
          if (p.lse_out && d == 0)
            p.lse_out[(int64_t)b * p.Hq + hq] = l > 0.f ? (m_used[g] + log2f(l)) * 0.6931471805599453f : -INFINITY;
        } else {
          const int64_t pi = ((int64_t)b * p.Hq + hq) * p.num_splits + split;
          p.part_out[pi * D + d] = o;
          if (d == 0) p.part_lse[pi] = l > 0.f ? m_used[g] + log2f(l) : -INFINITY;
        }
      }
      bar_sync_softmax();                           // wred is reused by the next item
      gt += n_tiles;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 64);
}

template <typename T>
__global__ void __launch_bounds__(128) split_reduce(const float* __restrict__ part_out, const float* __restrict__ part_lse,
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
    o += wgt * part_out[(base + s) * D + d];
  }
  out[(int64_t)tok * out_stride + (int64_t)hq * D + d] = from_f32<T>(den > 0.f ? o / den : 0.f);
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

template <typename T, int G>
static int launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const Params& p, dim3 grid,
                  cudaStream_t st) {
  const int smem = STAGES * (K_BYTES + V_BYTES) + 2 * Q_BYTES + P_BYTES + 32 * 8 + 16 + 4 * NQ * 4 +
                   2 * MAX_TILES_SMEM * 4 + 1024;
  auto k = fmha_decode_kernel<T, G>;
  static bool a = false;
  if (!a) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); a = true; }
  vb_launch(k, grid, dim3(NUM_THREADS), (size_t)smem, st, tq, tk, tv, p);
  return (int)cudaGetLastError();
}

}  // namespace fmha_dec

EXPORT int vb_fmha_tc_decode(const void* q, const void* kv_cache, int64_t num_blocks, const int* block_table,
                             const int* seq_lens, void* out, float* part_out, float* part_lse, float* lse_out, int B,
                             int Hq, int Hkv, int64_t q_stride, int64_t out_stride, int bt_stride, int num_splits,
                             int window, float scale, int dtype, cudaStream_t st) {
  using namespace fmha_dec;
  if (B == 0) return 0;
  if (dtype == DT_F32) return -1;
  const int G = Hq / Hkv;
  if (G > 16 || (G & (G - 1))) return -4;
  auto enc = get_encode();
  if (!enc) return -10;
  const bool bf16 = dtype == DT_BF16;
  const CUtensorMapDataType dt = bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  Params p;
  p.block_table = block_table; p.seq_lens = seq_lens; p.out = out; p.part_out = part_out; p.part_lse = part_lse;
  p.lse_out = lse_out; p.out_stride = out_stride; p.bt_stride = bt_stride; p.Hq = Hq; p.Hkv = Hkv; p.G = G;
  p.num_splits = num_splits; p.window = window; p.scale_log2 = scale * 1.4426950408889634f; p.B = B;
  CUtensorMap tq, tk, tv;
  {
    cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)Hq, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)D * 2, (cuuint64_t)q_stride * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)G, 1};
    cuuint32_t es[3] = {1, 1, 1};
    if (enc(&tq, dt, 3, const_cast<void*>(q), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return -11;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)num_blocks * 2 * Hkv * BS};
    cuuint64_t strides[1] = {(cuuint64_t)D * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)BS};
    cuuint32_t es[2] = {1, 1};
    if (enc(&tk, dt, 2, const_cast<void*>(kv_cache), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return -12;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)BS, (cuuint64_t)num_blocks * 2 * Hkv * D};
    cuuint64_t strides[1] = {(cuuint64_t)BS * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)D};
    cuuint32_t es[2] = {1, 1};
    if (enc(&tv, dt, 2, const_cast<void*>(kv_cache), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return -13;
  }
  static int n_sms = 0;
  if (!n_sms) { int dv = 0; cudaGetDevice(&dv); cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dv); }
  const int n_items = num_splits * Hkv * B;
  dim3 grid(n_items < n_sms ? n_items : n_sms);
  int rc = -5;
#define VB_GO(GV)                                                                          \
  rc = bf16 ? launch<__nv_bfloat16, GV>(tq, tk, tv, p, grid, st) : launch<__half, GV>(tq, tk, tv, p, grid, st)
  switch (G) {
    case 1: VB_GO(1); break;
    case 2: VB_GO(2); break;
    case 4: VB_GO(4); break;
    case 8: VB_GO(8); break;
    case 16: VB_GO(16); break;
  }
#undef VB_GO
  if (rc) return rc;
  if (num_splits > 1) {
    if (bf16) split_reduce<__nv_bfloat16><<<dim3(B, Hq), 128, 0, st>>>(part_out, part_lse, (__nv_bfloat16*)out, lse_out, Hq, num_splits, out_stride);
    else split_reduce<__half><<<dim3(B, Hq), 128, 0, st>>>(part_out, part_lse, (__half*)out, lse_out, Hq, num_splits, out_stride);
  }
  return (int)cudaGetLastError();
}
