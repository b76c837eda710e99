// Multi-head latent attention (MLA) decode on tcgen05: MQA over the paged latent cache, the latent tile shared by all
// heads of the CTA and by BOTH matrix products.
//
//   cache row (one per token, all heads share it):  [ c_kv : 512 | k_rope : 64 ]  (bf16 / fp16)
//   S^T[tok, h] = Kc[tok, 0:576] · Q[h, 0:576]^T      A = latent tile [128 tok x 64-dim k-blocks] K-major (TMA, SW128)
//                                                     B = Q           [N heads x 576]
//   O^T[d, h]  += Kc[tok, d]^T · P^T[tok, h]          A = THE SAME shared-memory bytes read as an MN-major operand
//                                                         (d contiguous inside a 128-byte swizzle row, tokens along K)
//                                                     B = P           [N heads x 128 tok]
// i.e. "swap-AB" as in attention_tc_decode.cu — the 128 KV tokens of a tile ride the MMA M axis for S and the MMA K axis
// for O, the heads are the MMA N (16 or 32 per CTA; 128-head models run 4 head groups per sequence) — and each latent
// byte is fetched from HBM once and from shared memory by two tcgen05.mma streams, never transposed, never copied.
//
// A tile is 64 KV tokens (9 k-blocks of [64 tok x 64 dims] = 72 KB); two tiles are resident.  Every k-block has its own
// full / empty mbarrier: QK starts on the first k-block that lands (rope first), the rope block is released when QK
// retires and each pair of latent blocks when the PV product of its 128 output dims retires — so the TMA stream of
// tile j+2 starts under the PV of tile j instead of after it (ncu of the whole-tile version: softmax warps spent 70 %
// of their time waiting for S, i.e. the load latency sat on the critical path once per tile).  The QK MMA is issued with M = 128 (the instruction shape the
// accumulator layout is simplest for): rows 64..127 read the following k-block's bytes, land in TMEM lanes the softmax
// never looks at, and cost nothing that matters — at 16-32 heads per CTA the kernel is bound by the latent stream
// (1152 B per token), not by tensor-core time.
//
// Reference kernels for the op: csrc/libtorch_stable/attention/mla/sm100_cutlass_mla_kernel.cu (+ cutlass_sm100_mla/),
// FlashMLA; algorithm: vllm/model_executor/layers/attention/mla_attention.py:1-200 (absorbed-weight decode).
#include "common.cuh"

namespace mla {

constexpr int DL = 512, DR = 64, DK = DL + DR;
constexpr int NKB = DK / 64;          // 9 k-blocks of 64 dims
constexpr int TILE = 64;
constexpr int KB_BYTES = TILE * 128;  // 8 KB: one k-block [64 tok x 64 dims]
constexpr int STAGE_BYTES = NKB * KB_BYTES;   // 72 KB
constexpr int STAGES = 2;
constexpr int NUM_THREADS = 192;
constexpr int MAX_IDS_SMEM = 1024;

struct Params {
  const int* block_table;
  const int* seq_lens;
  void* out;            // [B, H, 512]
  float* part_out;      // [B, H, splits, 512]
  float* part_lse;      // [B, H, splits]
  float* lse_out;       // optional [B, H] (natural log)
  int64_t out_stride;   // elements between batch rows of out
  int bt_stride;
  int H, BS, num_splits;
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

// MN-major SW128 operand: a 128-byte row holds 64 consecutive M (= latent dims) of ONE k (= token); 8 tokens per
// 1024-byte swizzle atom (SBO); the next 64 dims live `lbo_bytes` further (the next k-block buffer).
VB_DEVICE uint64_t umma_desc_sw128_mn(const void* smem_tile, uint32_t lbo_bytes) {
  const uint32_t addr = smem_u32(smem_tile);
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
VB_HOST_DEVICE uint32_t idesc_f16_a_mn(int M, int N, bool bf16) {
  return umma_idesc_f16(M, N, bf16) | (1u << 15);     // bit 15: A operand is MN-major
}

template <typename T, int N>
__global__ void __launch_bounds__(NUM_THREADS, 1)
mla_decode_kernel(const __grid_constant__ CUtensorMap tmap_qn, const __grid_constant__ CUtensorMap tmap_qr,
                  const __grid_constant__ CUtensorMap tmap_kv, const Params p) {
  constexpr int Q_BYTES = NKB * N * 128;     // 9 k-blocks of [N x 64]
  constexpr int P_BYTES = N * 128;           // [N x 64 tok]
  constexpr int CH = 16;                     // TMEM columns moved per tcgen05.ld/st
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* kv_smem = smem;                                // STAGES x 72 KB
  unsigned char* q_smem = kv_smem + STAGES * STAGE_BYTES;       // (also what QK's padding rows of the last k-block read)
  unsigned char* p_smem = q_smem + Q_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_smem + (P_BYTES < 8192 ? 8192 : P_BYTES));
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;              // [STAGES * NKB] one per k-block buffer
  uint64_t* kv_empty = kv_full + STAGES * NKB;
  uint64_t* s_full = kv_empty + STAGES * NKB;   // [2]
  uint64_t* s_empty = s_full + 2;            // [2]
  uint64_t* p_full = s_empty + 2;
  uint64_t* pv_done = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);
  float* wred = reinterpret_cast<float*>(tmem_slot + 2);        // [4][N]
  int* blk_ids = reinterpret_cast<int*>(wred + 4 * N);          // [MAX_IDS_SMEM]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, hg = blockIdx.y, b = blockIdx.z;
  const int h0 = hg * N;
  const int G = min(N, p.H - h0);            // real heads of this group
  const int BS = p.BS;
  const int ppt = BS >= TILE ? 1 : TILE / BS;           // cache pages (or page halves) per tile
  const int rows = BS >= TILE ? TILE : BS;              // token rows per TMA box
  pdl_trigger();
  pdl_wait();
  const int seq_len = p.seq_lens[b];
  int tile_lo = 0, tile_hi = (seq_len + TILE - 1) / TILE;
  {
    const int per = (tile_hi + p.num_splits - 1) / p.num_splits;
    tile_lo = split * per;
    tile_hi = min(tile_hi, tile_lo + per);
  }
  const int n_tiles = max(tile_hi - tile_lo, 0);
  constexpr uint32_t TMEM_COLS = 6 * N <= 128 ? 128 : 256;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qn); tma_prefetch_desc(&tmap_qr); tma_prefetch_desc(&tmap_kv);
    mbar_init(q_full, 1);
    for (int s = 0; s < STAGES * NKB; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 128); }
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  {
    // padded Q / P rows and the not-yet-loaded part of the last tile must be finite: zero everything once
    uint4 z = make_uint4(0, 0, 0, 0);
    uint4* s4 = reinterpret_cast<uint4*>(smem);
    const int n16 = (STAGES * STAGE_BYTES + Q_BYTES + (P_BYTES < 8192 ? 8192 : P_BYTES)) / 16;
    for (int i = threadIdx.x; i < n16; i += NUM_THREADS) s4[i] = z;
  }
  // token row (in the [num_blocks * BS, 576] view of the cache) of every TMA box of this split, -1 = past the sequence
  auto box_row = [&](int i) -> int {
    const int tok0 = tile_lo * TILE + i * rows;
    if (tok0 >= seq_len) return -1;
    const int blk = p.block_table[(int64_t)b * p.bt_stride + tok0 / BS];
    return blk * BS + tok0 % BS;
  };
  for (int i = threadIdx.x; i < ppt * n_tiles && i < MAX_IDS_SMEM; i += NUM_THREADS) blk_ids[i] = box_row(i);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 2 * N;           // S^T buffers at columns [0, N) and [N, 2N); O^T chunk c at 2N + cN

  if (n_tiles == 0) {
    if (warp >= 2 && p.num_splits > 1) {
      const int t = (warp & 3) * 32 + lane;
      for (int g = 0; g < G; ++g) {
        const int64_t pi = ((int64_t)b * p.H + (h0 + g)) * p.num_splits + split;
        for (int c = 0; c < 4; ++c) p.part_out[pi * DL + c * 128 + t] = 0.f;
        if (t == 0) p.part_lse[pi] = -INFINITY;
      }
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
    return;
  }

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer
      mbar_expect_tx(q_full, NKB * min(N, p.H) * 128);   // the full TMA box is always written (OOB heads as zeros)
      for (int kb = 0; kb < 8; ++kb) tma_load_3d(q_smem + kb * (N * 128), &tmap_qn, kb * 64, h0, b, q_full);
      tma_load_3d(q_smem + 8 * (N * 128), &tmap_qr, 0, h0, b, q_full);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % STAGES;
        const uint32_t par = ((j / STAGES) & 1) ^ 1;
        int row[8], n_valid = 0;
        for (int u = 0; u < ppt; ++u) {
          const int i = j * ppt + u;
          row[u] = i < MAX_IDS_SMEM ? blk_ids[i] : box_row(i);
          n_valid += row[u] >= 0;
        }
        const uint32_t bytes = (uint32_t)n_valid * rows * 128;
        unsigned char* stp = kv_smem + st * STAGE_BYTES;
#pragma unroll 1
        for (int i = 0; i < NKB; ++i) {
          const int kb = i == 0 ? NKB - 1 : i - 1;             // rope block first: QK releases it first
          const int slot = st * NKB + kb;
          mbar_wait(&kv_empty[slot], par);
          mbar_expect_tx(&kv_full[slot], bytes);
          for (int u = 0; u < ppt; ++u)
            if (row[u] >= 0)
              tma_load_2d_hint(stp + kb * KB_BYTES + u * rows * 128, &tmap_kv, kb * 64, row[u], &kv_full[slot], L2_EVICT_FIRST);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer
    const bool bf16 = DTypeOf<T>::v == DT_BF16;
    const uint32_t idesc_s = umma_idesc_f16(128, N, bf16);
    const uint32_t idesc_o = idesc_f16_a_mn(128, N, bf16);
    mbar_wait(q_full, 0);
    // descriptors are built once; every MMA adds a compile-time offset (in 16-byte units) to the address field
    const uint64_t d_kv = umma_desc_sw128(kv_smem);
    const uint64_t d_kv_mn = umma_desc_sw128_mn(kv_smem, KB_BYTES);
    const uint64_t d_q = umma_desc_sw128(q_smem);
    const uint64_t d_p = umma_desc_sw128(p_smem);
    for (int j = 0; j <= n_tiles; ++j) {
      if (j < n_tiles) {
        // ---- S^T_j = Kc · Q^T, k-block by k-block as they land (M = 128: rows 64..127 are padding, see the header)
        const int st = j % STAGES, sb = j & 1;
        const uint32_t par = (j / STAGES) & 1;
        mbar_wait(&s_empty[sb], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int i = 0; i < NKB; ++i) {
          const int kb = i == 0 ? NKB - 1 : i - 1;
          const int slot = st * NKB + kb;
          mbar_wait(&kv_full[slot], par);
          tc_fence_after();
          if (lane == 0) {
            const uint64_t da = d_kv + (uint64_t)((st * STAGE_BYTES + kb * KB_BYTES) >> 4);
            const uint64_t db = d_q + (uint64_t)((kb * (N * 128)) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem_base + sb * N, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc_s, (i | k) != 0);
          }
          __syncwarp();
        }
        if (lane == 0) {
          umma_commit(&s_full[sb]);
          umma_commit(&kv_empty[st * NKB + NKB - 1]);          // the rope block is only read by QK
        }
        __syncwarp();
      }
      if (j > 0) {
        // ---- O^T += V^T · P^T for tile j-1: four 128-dim chunks (k-blocks 2c, 2c+1 read MN-major), K = 64 tokens;
        // each chunk hands its two k-block buffers back to the producer as soon as it retires
        const int jj = j - 1, st = jj % STAGES;
        mbar_wait(p_full, jj & 1);
        tc_fence_after();
        if (lane == 0) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int kk = 0; kk < TILE / 16; ++kk) {
              const uint64_t da = d_kv_mn + (uint64_t)((st * STAGE_BYTES + (2 * c) * KB_BYTES + kk * 2048) >> 4);
              umma_f16(tmem_o + c * N, da, d_p + (uint64_t)(kk * 2), idesc_o, (jj | kk) != 0);
            }
            umma_commit(&kv_empty[st * NKB + 2 * c]);
            umma_commit(&kv_empty[st * NKB + 2 * c + 1]);
          }
          umma_commit(pv_done);
        }
        __syncwarp();
      }
    }
  } else {
    // ===== softmax (thread = kv token of the tile), then epilogue (thread = latent dim within a 128-chunk)
    const int qd = warp & 3;
    const int t = qd * 32 + lane;
    const uint32_t t_lane = (uint32_t)(qd * 32) << 16;
    float m_used[N], l_thr[N];
#pragma unroll
    for (int g = 0; g < N; ++g) { m_used[g] = -INFINITY; l_thr[g] = 0.f; }
    for (int j = 0; j < n_tiles; ++j) {
      const int sb = j & 1;
      mbar_wait(&s_full[sb], (j >> 1) & 1);
      tc_fence_after();
      float s[N];
#pragma unroll
      for (int c = 0; c < N; c += CH) {
        uint32_t sv[CH];
        tmem_ld_32x16(tmem_base + sb * N + c + t_lane, sv);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < CH; ++g) s[c + g] = __uint_as_float(sv[g]);
      }
      tc_fence_before();
      mbar_arrive(&s_empty[sb]);
      const int kpos = (tile_lo + j) * TILE + t;
      const bool ok = t < TILE && kpos < seq_len;        // TMEM lanes 64..127 are the QK padding rows
#pragma unroll
      for (int g = 0; g < N; ++g) {
        s[g] = ok ? s[g] * p.scale_log2 : -INFINITY;
        const float cm = warp_max(s[g]);
        if (lane == 0) wred[qd * N + g] = cm;
      }
      bar_sync_softmax();
      float alpha[N];
      bool rescale = false;
#pragma unroll
      for (int g = 0; g < N; ++g) {
        const float mx = fmaxf(fmaxf(wred[g], wred[N + g]), fmaxf(wred[2 * N + g], wred[3 * N + g]));
        alpha[g] = 1.f;
        if (mx > m_used[g] + 8.f || (m_used[g] == -INFINITY && mx > -INFINITY)) {
          if (m_used[g] != -INFINITY) { alpha[g] = ex2(m_used[g] - mx); rescale = true; }
          m_used[g] = mx;
        }
      }
      if (j > 0) { mbar_wait(pv_done, (j - 1) & 1); tc_fence_after(); }   // P and O^T are free again
      if (rescale) {       // uniform over the CTA (every thread sees the same column maxima)
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
#pragma unroll
          for (int c = 0; c < N; c += CH) {
            uint32_t ov[CH];
            tmem_ld_32x16(tmem_o + c4 * N + c + t_lane, ov);
            tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < CH; ++g) ov[g] = __float_as_uint(__uint_as_float(ov[g]) * alpha[c + g]);
            tmem_st_32x16(tmem_o + c4 * N + c + t_lane, ov);
          }
        }
        tmem_st_wait();
#pragma unroll
        for (int g = 0; g < N; ++g) l_thr[g] *= alpha[g];
      }
      unsigned char* pcol = p_smem + ((t & 63) & 7) * 2;
      const int chunk = (t & 63) >> 3;
#pragma unroll
      for (int g = 0; g < N; ++g) {
        const float m_ref = (m_used[g] == -INFINITY) ? 0.f : m_used[g];
        const float pe = ok ? ex2(s[g] - m_ref) : 0.f;
        l_thr[g] += pe;
        if (t < TILE) *reinterpret_cast<T*>(pcol + g * 128 + ((chunk ^ (g & 7)) << 4)) = from_f32<T>(pe);
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(p_full);
      bar_sync_softmax();                         // wred is reused by the next tile
    }
    // ---- epilogue
#pragma unroll
    for (int g = 0; g < N; ++g) {
      const float lt = warp_sum(l_thr[g]);
      if (lane == 0) wred[qd * N + g] = lt;
    }
    bar_sync_softmax();
    mbar_wait(pv_done, (n_tiles - 1) & 1);
    tc_fence_after();
    float inv[N], lsum[N];
#pragma unroll
    for (int g = 0; g < N; ++g) {
      lsum[g] = wred[g] + wred[N + g] + wred[2 * N + g] + wred[3 * N + g];
      inv[g] = lsum[g] > 0.f ? 1.f / lsum[g] : 0.f;
    }
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      const int d = c4 * 128 + t;
#pragma unroll
      for (int c = 0; c < N; c += CH) {
        uint32_t ov[CH];
        tmem_ld_32x16(tmem_o + c4 * N + c + t_lane, ov);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < CH; ++g) {
          if (c + g < G) {
            const float o = __uint_as_float(ov[g]) * inv[c + g];
            const int hq = h0 + c + g;
            if (p.num_splits == 1) {
              reinterpret_cast<T*>(p.out)[(int64_t)b * p.out_stride + (int64_t)hq * DL + d] = from_f32<T>(o);
            } else {
              const int64_t pi = ((int64_t)b * p.H + hq) * p.num_splits + split;
              p.part_out[pi * DL + d] = o;
            }
          }
        }
      }
    }
    if (t == 0) {
#pragma unroll
      for (int g = 0; g < N; ++g) {
        if (g < G) {
          const int hq = h0 + g;
          const float lse2 = lsum[g] > 0.f ? m_used[g] + log2f(lsum[g]) : -INFINITY;
          if (p.num_splits == 1) {
            if (p.lse_out) p.lse_out[(int64_t)b * p.H + hq] = lse2 * 0.6931471805599453f;
          } else {
            p.part_lse[((int64_t)b * p.H + hq) * p.num_splits + split] = lse2;
          }
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <typename T>
__global__ void __launch_bounds__(128) mla_split_reduce(const float* __restrict__ part_out, const float* __restrict__ part_lse,
                                                        T* __restrict__ out, float* __restrict__ lse_out, int H, int S,
                                                        int64_t out_stride) {
  const int tok = blockIdx.x, hq = blockIdx.y;
  const int64_t base = ((int64_t)tok * H + hq) * S;
  float m = -INFINITY;
  for (int s = 0; s < S; ++s) m = fmaxf(m, part_lse[base + s]);
  const float m_use = (m == -INFINITY) ? 0.f : m;
  float den = 0.f;
  for (int s = 0; s < S; ++s) den += ex2(part_lse[base + s] - m_use);
  for (int d = threadIdx.x; d < DL; d += blockDim.x) {
    float o = 0.f;
    for (int s = 0; s < S; ++s) o += ex2(part_lse[base + s] - m_use) * part_out[(base + s) * DL + d];
    out[(int64_t)tok * out_stride + (int64_t)hq * DL + d] = from_f32<T>(den > 0.f ? o / den : 0.f);
  }
  if (lse_out && threadIdx.x == 0) lse_out[(int64_t)tok * H + hq] = den > 0.f ? (m + log2f(den)) * 0.6931471805599453f : -INFINITY;
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

template <typename T, int N>
static int launch(const CUtensorMap& tqn, const CUtensorMap& tqr, const CUtensorMap& tkv, const Params& p, dim3 grid,
                  cudaStream_t st) {
  const int smem = STAGES * STAGE_BYTES + NKB * N * 128 + 8192 + (2 * STAGES * NKB + 8) * 8 + 16 + 4 * N * 4 + MAX_IDS_SMEM * 4 + 1024;
  auto k = mla_decode_kernel<T, N>;
  static bool a = false;
  if (!a) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); a = true; }
  vb_launch(k, grid, dim3(NUM_THREADS), (size_t)smem, st, tqn, tqr, tkv, p);
  return (int)cudaGetLastError();
}

}  // namespace mla

// q_nope: [B, H, 512] (absorbed W_UK), q_pe: [B, H, 64], cache: [num_blocks, BS, 576]; out: [B, H, 512].
EXPORT int vb_mla_decode(const void* q_nope, const void* q_pe, const void* cache, int64_t num_blocks, const int* block_table,
                         const int* seq_lens, void* out, float* part_out, float* part_lse, float* lse_out, int B, int H,
                         int BS, int64_t qn_stride_b, int64_t qn_stride_h, int64_t qr_stride_b, int64_t qr_stride_h,
                         int64_t out_stride, int bt_stride, int num_splits, float scale, int dtype, cudaStream_t st) {
  using namespace mla;
  if (B == 0) return 0;
  if (dtype == DT_F32) return -1;
  if (BS != 16 && BS != 32 && BS != 64 && BS != 128) return -4;
  auto enc = get_encode();
  if (!enc) return -10;
  const bool bf16 = dtype == DT_BF16;
  const CUtensorMapDataType dt = bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const int N = H > 16 ? 32 : 16;
  Params p;
  p.block_table = block_table; p.seq_lens = seq_lens; p.out = out; p.part_out = part_out; p.part_lse = part_lse;
  p.lse_out = lse_out; p.out_stride = out_stride; p.bt_stride = bt_stride; p.H = H; p.BS = BS; p.num_splits = num_splits;
  p.scale_log2 = scale * 1.4426950408889634f;
  CUtensorMap tqn, tqr, tkv;
  const int gbox = H < N ? H : N;
  {
    cuuint64_t dims[3] = {(cuuint64_t)DL, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)qn_stride_h * 2, (cuuint64_t)qn_stride_b * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)gbox, 1};
    cuuint32_t es[3] = {1, 1, 1};
    if (enc(&tqn, dt, 3, const_cast<void*>(q_nope), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return -11;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)DR, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)qr_stride_h * 2, (cuuint64_t)qr_stride_b * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)gbox, 1};
    cuuint32_t es[3] = {1, 1, 1};
    if (enc(&tqr, dt, 3, const_cast<void*>(q_pe), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return -12;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)DK, (cuuint64_t)num_blocks * BS};
    cuuint64_t strides[1] = {(cuuint64_t)DK * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)(BS >= TILE ? TILE : BS)};
    cuuint32_t es[2] = {1, 1};
    if (enc(&tkv, dt, 2, const_cast<void*>(cache), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return -13;
  }
  dim3 grid(num_splits, (H + N - 1) / N, B);
  int rc;
  if (N == 16) rc = bf16 ? launch<__nv_bfloat16, 16>(tqn, tqr, tkv, p, grid, st) : launch<__half, 16>(tqn, tqr, tkv, p, grid, st);
  else rc = bf16 ? launch<__nv_bfloat16, 32>(tqn, tqr, tkv, p, grid, st) : launch<__half, 32>(tqn, tqr, tkv, p, grid, st);
  if (rc) return rc;
  if (num_splits > 1) {
    if (bf16) mla_split_reduce<__nv_bfloat16><<<dim3(B, H), 128, 0, st>>>(part_out, part_lse, (__nv_bfloat16*)out, lse_out, H, num_splits, out_stride);
    else mla_split_reduce<__half><<<dim3(B, H), 128, 0, st>>>(part_out, part_lse, (__half*)out, lse_out, H, num_splits, out_stride);
  }
  return (int)cudaGetLastError();
}
