// Paged attention, CUDA-core path (decode with split-KV, and a varlen causal prefill fallback).
//
// What the reference runs here: FlashInfer / TRT-LLM-gen cubins (v1/attention/backends/flashinfer.py:2239,2457),
// Triton unified attention (ops/triton_unified_attention.py).  This file is our own memory-bound kernel:
//   * one CTA = (KV split, kv head, request / q tile); R <= 8 query rows that share the kv head (GQA group x
//     q tokens) are held in registers, so K/V pages are read once per group, not once per q head;
//   * K and V pages ([block_size, D] contiguous in our cache layout) are fetched with ONE 1-D bulk copy each
//     (cp.async.bulk → UBLKCP) into a 4-stage shared-memory ring, completion tracked by mbarriers;
//   * each warp owns block_size/4 tokens of every page and keeps its own online-softmax state; the four
//     states are merged once at the end (no per-page cross-warp traffic);
//   * split partials (fp32 out + LSE) are merged by a second tiny kernel; grid shapes depend only on
//     (batch bucket, splits) so the launch is CUDA-graph safe.
// The tcgen05 path for the compute-bound prefill case lives in attention_tc.cu.
#include "common.cuh"

namespace {

constexpr int WARPS = 4;
constexpr int THREADS = WARPS * 32;
constexpr int STAGES = 4;
constexpr int MAX_PAGES_SMEM = 2048;  // block ids cached per CTA (covers 64k tokens per split at BS=32)

struct AttnParams {
  const void* q;          // [T, Hq, D]
  const void* kv_cache;   // [nb, 2, Hkv, BS, D]
  const int* block_table; // [B, bt_stride]
  const int* seq_lens;    // [B]
  const int* qsl;         // [B+1] query start loc (nullptr → decode: one token per request)
  void* out;              // [T, Hq, D]
  float* part_out;        // [T, Hq, S, D]
  float* part_lse;        // [T, Hq, S]
  float* lse_out;         // optional [T, Hq]
  int64_t q_stride;       // elements between tokens in q
  int64_t out_stride;
  int bt_stride;
  int Hq, Hkv, num_splits;
  int window;             // sliding window (0 = off)
  float scale_log2;       // softmax scale * log2(e)
  float soft_cap;         // 0 = off
  float kv_scale;         // fp8 dequant scale
  int causal;
  const float* alibi;     // optional [Hq] slopes: score += slope * (kpos - qpos)   (Bloom / MPT / Falcon-RW)
  const float* sinks;     // optional [Hq] attention-sink logits: extra softmax mass that attends to nothing (GPT-OSS)
  int chunk;              // chunked-local attention (Llama-4): keys of the query's own chunk only (0 = off)
};

template <typename T, int D, int R, int BS, bool FP8>
__global__ void __launch_bounds__(THREADS) paged_attn_kernel(const AttnParams p) {
  constexpr int DPL = D / 32;                 // dims per lane
  constexpr int TPW = BS / WARPS;             // tokens per warp per page
  constexpr int KVB = FP8 ? 1 : sizeof(T);
  constexpr int PAGE_BYTES = BS * D * KVB;
  static_assert(TPW == 8 || TPW == 4 || TPW == 16, "block_size must be 16, 32 or 64");
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* kv_smem = smem;                                        // STAGES * 2 * PAGE_BYTES
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * 2 * PAGE_BYTES);
  int* blk_ids = reinterpret_cast<int*>(full + STAGES);                 // MAX_PAGES_SMEM
  float* p_smem = reinterpret_cast<float*>(blk_ids + MAX_PAGES_SMEM);   // WARPS * R * TPW
  float* merge = p_smem + WARPS * R * TPW;                              // WARPS * R * (D + 2)

  const int G = p.Hq / p.Hkv;
  const int TQ = R / G > 0 ? R / G : 1;       // q tokens per CTA
  const int b = blockIdx.z, h = blockIdx.y;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int split, qt;
  if (p.qsl) { split = 0; qt = blockIdx.x; } else { split = blockIdx.x; qt = 0; }
  const int seq_len = p.seq_lens[b];
  const int q0 = p.qsl ? p.qsl[b] : b;
  const int q_len = p.qsl ? p.qsl[b + 1] - q0 : 1;
  const int tok0 = qt * TQ;
  if (tok0 >= q_len) return;
  const int n_tok = min(TQ, q_len - tok0);
  const int pos0 = seq_len - q_len + tok0;     // absolute position of the tile's first query token
  const int pos_last = pos0 + n_tok - 1;

  // KV range for this CTA
  int kv_end = p.causal ? pos_last + 1 : seq_len;
  int kv_begin = 0;
  if (p.window > 0) kv_begin = max(0, pos0 - p.window + 1);
  int page_lo = kv_begin / BS, page_hi = (kv_end + BS - 1) / BS;
  const int n_pages_total = page_hi - page_lo;
  const int per_split = (n_pages_total + p.num_splits - 1) / p.num_splits;
  page_lo += split * per_split;
  page_hi = min(page_hi, page_lo + per_split);
  const int n_pages = max(page_hi - page_lo, 0);

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < n_pages && i < MAX_PAGES_SMEM; i += THREADS)
    blk_ids[i] = p.block_table[(int64_t)b * p.bt_stride + page_lo + i];
  __syncthreads();

  const unsigned char* cache = reinterpret_cast<const unsigned char*>(p.kv_cache);
  const size_t head_bytes = (size_t)PAGE_BYTES;
  auto issue = [&](int i) {
    const int s = i % STAGES;
    const size_t blk = (size_t)(i < MAX_PAGES_SMEM ? blk_ids[i]
                                                   : p.block_table[(int64_t)b * p.bt_stride + page_lo + i]);
    const unsigned char* kp = cache + ((blk * 2 + 0) * p.Hkv + h) * head_bytes;
    const unsigned char* vp = cache + ((blk * 2 + 1) * p.Hkv + h) * head_bytes;
    mbar_expect_tx(&full[s], 2 * PAGE_BYTES);
    bulk_g2s(kv_smem + (size_t)s * 2 * PAGE_BYTES, kp, PAGE_BYTES, &full[s]);
    bulk_g2s(kv_smem + (size_t)s * 2 * PAGE_BYTES + PAGE_BYTES, vp, PAGE_BYTES, &full[s]);
  };
  if (threadIdx.x == 0)
    for (int i = 0; i < STAGES - 1 && i < n_pages; ++i) issue(i);

  // query rows → registers (pre-scaled so that scores are in log2 units)
  float qreg[R][DPL];
  int row_pos[R];
  const T* qbase = reinterpret_cast<const T*>(p.q);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int tq = r / G, g = r % G;
    const bool valid = tq < n_tok && (r < G * TQ);
    row_pos[r] = valid ? pos0 + tq : -1;
    const T* qp = qbase + (int64_t)(q0 + tok0 + (valid ? tq : 0)) * p.q_stride + (int64_t)(h * G + g) * D + lane * DPL;
#pragma unroll
    for (int d = 0; d < DPL; ++d) qreg[r][d] = valid ? to_f32(qp[d]) * (p.soft_cap > 0.f ? 1.f : p.scale_log2) : 0.f;
  }
  float m_run[R], l_run[R], acc[R][DPL];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    m_run[r] = -INFINITY; l_run[r] = 0.f;
#pragma unroll
    for (int d = 0; d < DPL; ++d) acc[r][d] = 0.f;
  }
  // after the butterfly, lane l holds the score of token tsel(l) of the warp slice
  int tsel;
  if (TPW == 8) tsel = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  else if (TPW == 4) tsel = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
  else tsel = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
  float* my_p = p_smem + w * R * TPW;

  for (int i = 0; i < n_pages; ++i) {
    if (threadIdx.x == 0 && i + STAGES - 1 < n_pages) issue(i + STAGES - 1);
    mbar_wait(&full[i % STAGES], (i / STAGES) & 1);
    const unsigned char* kpage = kv_smem + (size_t)(i % STAGES) * 2 * PAGE_BYTES;
    const unsigned char* vpage = kpage + PAGE_BYTES;
    const int tok_base = (page_lo + i) * BS + w * TPW;   // absolute kv position of this warp's first token
    // ---- K slice → fp32 registers
    float kreg[TPW][DPL];
#pragma unroll
    for (int tt = 0; tt < TPW; ++tt) {
      const unsigned char* kp = kpage + ((size_t)(w * TPW + tt) * D + lane * DPL) * KVB;
      if (FP8) {
#pragma unroll
        for (int d = 0; d < DPL; ++d)
          kreg[tt][d] = __half2float(__nv_cvt_fp8_to_halfraw(kp[d], __NV_E4M3)) * p.kv_scale;
      } else {
        const T* kt = reinterpret_cast<const T*>(kp);
#pragma unroll
        for (int d = 0; d < DPL; ++d) kreg[tt][d] = to_f32(kt[d]);
      }
    }
    // ---- scores, one query row at a time
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float part[TPW];
#pragma unroll
      for (int tt = 0; tt < TPW; ++tt) {
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < DPL; ++d) s += qreg[r][d] * kreg[tt][d];
        part[tt] = s;
      }
      // multi-value butterfly: TPW partials x 32 lanes → one total per lane
#pragma unroll
      for (int width = TPW, off = 16; width > 1; width >>= 1, off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < width / 2; ++j) {
          const float send = upper ? part[j] : part[j + width / 2];
          const float keep = upper ? part[j + width / 2] : part[j];
          part[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      float s = part[0];
      if (TPW == 8) { s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 1); }
      else if (TPW == 4) { s += __shfl_xor_sync(0xffffffffu, s, 4); s += __shfl_xor_sync(0xffffffffu, s, 2);
                           s += __shfl_xor_sync(0xffffffffu, s, 1); }
      else { s += __shfl_xor_sync(0xffffffffu, s, 1); }
      if (p.soft_cap > 0.f) s = p.soft_cap * tanhf(s * (p.scale_log2 * 0.6931471805599453f) / p.soft_cap) * 1.4426950408889634f;
      const int kpos = tok_base + tsel;
      const int qpos = row_pos[r];
      if (p.alibi) s += p.alibi[h * G + (r % G)] * 1.4426950408889634f * (float)(kpos - qpos);
      bool ok = qpos >= 0 && kpos < kv_end && kpos >= kv_begin;
      if (p.causal) ok = ok && kpos <= qpos;
      if (p.window > 0) ok = ok && kpos > qpos - p.window && (p.causal || kpos < qpos + p.window);
      if (p.chunk > 0) ok = ok && (kpos / p.chunk == qpos / p.chunk);
      s = ok ? s : -INFINITY;
      const float pmax = warp_max(s);
      const float m_new = fmaxf(m_run[r], pmax);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = ex2(m_run[r] - m_use);
      const float pe = ex2(s - m_use);
      m_run[r] = m_new;
      l_run[r] = l_run[r] * alpha + pe;   // per-lane partial sum (4x/2x/... duplicated across the lanes of a token)
#pragma unroll
      for (int d = 0; d < DPL; ++d) acc[r][d] *= alpha;
      my_p[r * TPW + tsel] = pe;          // duplicates write the same value
    }
    __syncwarp();
    // ---- P @ V for this warp's tokens
#pragma unroll
    for (int tt = 0; tt < TPW; ++tt) {
      float vreg[DPL];
      const unsigned char* vp = vpage + ((size_t)(w * TPW + tt) * D + lane * DPL) * KVB;
      if (FP8) {
#pragma unroll
        for (int d = 0; d < DPL; ++d) vreg[d] = __half2float(__nv_cvt_fp8_to_halfraw(vp[d], __NV_E4M3)) * p.kv_scale;
      } else {
        const T* vt = reinterpret_cast<const T*>(vp);
#pragma unroll
        for (int d = 0; d < DPL; ++d) vreg[d] = to_f32(vt[d]);
      }
      const bool in_range = (tok_base + tt) < kv_end;   // never touch stale slots past the sequence
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float pe = my_p[r * TPW + tt];
        if (in_range) {
#pragma unroll
          for (int d = 0; d < DPL; ++d) acc[r][d] += pe * vreg[d];
        }
      }
    }
    __syncthreads();  // stage is free for the producer; also orders my_p reuse
  }

  // ---- merge the four warps
  constexpr int DUP = 32 / TPW;  // lanes holding the same token's p
#pragma unroll
  for (int r = 0; r < R; ++r) {
    float l = warp_sum(l_run[r]) / (float)DUP;
    float* mr = merge + (w * R + r) * (D + 2);
    if (lane == 0) { mr[D] = m_run[r]; mr[D + 1] = l; }
#pragma unroll
    for (int d = 0; d < DPL; ++d) mr[lane * DPL + d] = acc[r][d];
  }
  __syncthreads();
  const int T_total_rows = R;
  for (int idx = threadIdx.x; idx < T_total_rows * D; idx += THREADS) {
    const int r = idx / D, d = idx % D;
    const int tq = r / G, g = r % G;
    if (tq >= n_tok || r >= G * TQ) continue;
    float m = -INFINITY;
#pragma unroll
    for (int ww = 0; ww < WARPS; ++ww) m = fmaxf(m, merge[(ww * R + r) * (D + 2) + D]);
    const bool has_sink = p.sinks != nullptr && split == 0;      // counted once: split 0 carries the sink mass
    const float sink_l2 = has_sink ? p.sinks[h * G + g] * 1.4426950408889634f : -INFINITY;
    m = fmaxf(m, sink_l2);
    const float m_use = (m == -INFINITY) ? 0.f : m;
    float l = has_sink ? ex2(sink_l2 - m_use) : 0.f, o = 0.f;
#pragma unroll
    for (int ww = 0; ww < WARPS; ++ww) {
      const float* mr = merge + (ww * R + r) * (D + 2);
      const float sc = ex2(mr[D] - m_use);
      l += mr[D + 1] * sc;
      o += mr[d] * sc;
    }
    const float inv = l > 0.f ? 1.f / l : 0.f;
    const int64_t tok = q0 + tok0 + tq;
    const int hq = h * G + g;
    if (p.num_splits == 1) {
      reinterpret_cast<T*>(p.out)[tok * p.out_stride + (int64_t)hq * D + d] = from_f32<T>(o * inv);
      if (p.lse_out && d == 0) p.lse_out[tok * p.Hq + hq] = (l > 0.f) ? (m + log2f(l)) * 0.6931471805599453f : -INFINITY;
    } else {
      const int64_t pi = (tok * p.Hq + hq) * p.num_splits + split;
      p.part_out[pi * D + d] = o * inv;
      if (d == 0) p.part_lse[pi] = (l > 0.f) ? m + log2f(l) : -INFINITY;   // log2 units
    }
  }
}

// merge split partials: grid (T, Hq), D threads
template <typename T, int D>
__global__ void __launch_bounds__(D) split_reduce_kernel(const float* __restrict__ part_out,
                                                          const float* __restrict__ part_lse, T* __restrict__ out,
                                                          float* __restrict__ lse_out, int Hq, int S,
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

// LSE-weighted merge of two attention results (cascade / chunked context / DCP)
template <typename T>
__global__ void merge_states_kernel(T* __restrict__ out, float* __restrict__ lse_out, const T* __restrict__ o1,
                                    const float* __restrict__ l1, const T* __restrict__ o2,
                                    const float* __restrict__ l2, int D, int64_t n_rows) {
  const int64_t row = blockIdx.x;
  if (row >= n_rows) return;
  const float a = l1[row], b = l2[row];
  float m = fmaxf(a, b);
  if (m == -INFINITY) m = 0.f;
  const float w1 = __expf(a - m), w2 = __expf(b - m);
  const float den = fmaxf(w1 + w2, 1e-30f);
  for (int d = threadIdx.x; d < D; d += blockDim.x)
    out[row * D + d] = from_f32<T>((to_f32(o1[row * D + d]) * w1 + to_f32(o2[row * D + d]) * w2) / den);
  if (lse_out && threadIdx.x == 0) lse_out[row] = m + __logf(den);
}

template <typename T, int D, int R, int BS, bool FP8>
int launch_attn(const AttnParams& p, dim3 grid, cudaStream_t st) {
  constexpr int KVB = FP8 ? 1 : sizeof(T);
  constexpr int TPW = BS / WARPS;
  const size_t smem = (size_t)STAGES * 2 * BS * D * KVB + STAGES * 8 + MAX_PAGES_SMEM * 4 +
                      (size_t)WARPS * R * TPW * 4 + (size_t)WARPS * R * (D + 2) * 4;
  auto k = paged_attn_kernel<T, D, R, BS, FP8>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  k<<<grid, THREADS, smem, st>>>(p);
  return (int)cudaGetLastError();
}

template <typename T, int D, int BS, bool FP8>
int dispatch_rows(int R, const AttnParams& p, dim3 grid, cudaStream_t st) {
  switch (R) {
    case 1: return launch_attn<T, D, 1, BS, FP8>(p, grid, st);
    case 2: return launch_attn<T, D, 2, BS, FP8>(p, grid, st);
    case 4: return launch_attn<T, D, 4, BS, FP8>(p, grid, st);
    case 8: return launch_attn<T, D, 8, BS, FP8>(p, grid, st);
    default: return -2;
  }
}

template <typename T, bool FP8>
int dispatch_shape(int D, int BS, int R, const AttnParams& p, dim3 grid, cudaStream_t st) {
  if (D == 128 && BS == 64) return dispatch_rows<T, 128, 64, FP8>(R, p, grid, st);   // CUDA default page size: fp8 KV
  if (D == 64 && BS == 64) return dispatch_rows<T, 64, 64, FP8>(R, p, grid, st);     // caches, soft-cap / ALiBi / sink
  if (D == 32 && BS == 64) return dispatch_rows<T, 32, 64, FP8>(R, p, grid, st);     // layers, draft heads
  if (D == 128 && BS == 32) return dispatch_rows<T, 128, 32, FP8>(R, p, grid, st);
  if (D == 128 && BS == 16) return dispatch_rows<T, 128, 16, FP8>(R, p, grid, st);
  if (D == 64 && BS == 32) return dispatch_rows<T, 64, 32, FP8>(R, p, grid, st);
  if (D == 64 && BS == 16) return dispatch_rows<T, 64, 16, FP8>(R, p, grid, st);
  if (D == 256 && BS == 32) return dispatch_rows<T, 256, 32, FP8>(R, p, grid, st);
  if (D == 32 && BS == 32) return dispatch_rows<T, 32, 32, FP8>(R, p, grid, st);
  if (D == 32 && BS == 16) return dispatch_rows<T, 32, 16, FP8>(R, p, grid, st);
  return -3;
}

}  // namespace

// Unified entry. qsl == nullptr → decode (one query token per request, grid.x = num_splits).
// Returns 0 on success, <0 for unsupported shapes (caller raises).
EXPORT int vb_paged_attention_ex(const void* q, const void* kv_cache, const int* block_table, const int* seq_lens,
                                 const int* qsl, void* out, float* part_out, float* part_lse, float* lse_out, int B,
                                 int max_q_len, int Hq, int Hkv, int D, int BS, int64_t q_stride, int64_t out_stride,
                                 int bt_stride, int num_splits, int window, float scale, float soft_cap, float kv_scale,
                                 int causal, int kv_fp8, int dtype, const float* alibi, const float* sinks, int chunk,
                                 cudaStream_t st);

EXPORT int vb_paged_attention(const void* q, const void* kv_cache, const int* block_table, const int* seq_lens,
                              const int* qsl, void* out, float* part_out, float* part_lse, float* lse_out, int B,
                              int max_q_len, int Hq, int Hkv, int D, int BS, int64_t q_stride, int64_t out_stride,
                              int bt_stride, int num_splits, int window, float scale, float soft_cap, float kv_scale,
                              int causal, int kv_fp8, int dtype, cudaStream_t st) {
  return vb_paged_attention_ex(q, kv_cache, block_table, seq_lens, qsl, out, part_out, part_lse, lse_out, B, max_q_len, Hq,
                               Hkv, D, BS, q_stride, out_stride, bt_stride, num_splits, window, scale, soft_cap, kv_scale,
                               causal, kv_fp8, dtype, nullptr, nullptr, 0, st);
}

EXPORT int vb_paged_attention_ex(const void* q, const void* kv_cache, const int* block_table, const int* seq_lens,
                                 const int* qsl, void* out, float* part_out, float* part_lse, float* lse_out, int B,
                                 int max_q_len, int Hq, int Hkv, int D, int BS, int64_t q_stride, int64_t out_stride,
                                 int bt_stride, int num_splits, int window, float scale, float soft_cap, float kv_scale,
                                 int causal, int kv_fp8, int dtype, const float* alibi, const float* sinks, int chunk,
                                 cudaStream_t st) {
  if (B == 0) return 0;
  if (dtype == DT_F32) return -1;
  const int G = Hq / Hkv;
  AttnParams p;
  p.q = q; p.kv_cache = kv_cache; p.block_table = block_table; p.seq_lens = seq_lens; p.qsl = qsl; p.out = out;
  p.part_out = part_out; p.part_lse = part_lse; p.lse_out = lse_out; p.q_stride = q_stride; p.out_stride = out_stride;
  p.bt_stride = bt_stride; p.Hq = Hq; p.Hkv = Hkv; p.num_splits = qsl ? 1 : num_splits; p.window = window;
  p.scale_log2 = scale * 1.4426950408889634f; p.soft_cap = soft_cap; p.kv_scale = kv_scale; p.causal = causal;
  p.alibi = alibi; p.sinks = sinks; p.chunk = chunk;
  int R, TQ;
  if (G > 8 || (G & (G - 1))) return -4;     // group sizes 1,2,4,8
  if (qsl) { R = 8; TQ = 8 / G; } else { R = G; TQ = 1; }
  dim3 grid;
  if (qsl) grid = dim3((max_q_len + TQ - 1) / TQ, Hkv, B);
  else grid = dim3(p.num_splits, Hkv, B);
  int rc;
  DISPATCH_HALF(dtype, TT, {
    rc = kv_fp8 ? dispatch_shape<TT, true>(D, BS, R, p, grid, st) : dispatch_shape<TT, false>(D, BS, R, p, grid, st);
  });
  if (rc != 0) return rc;
  if (!qsl && p.num_splits > 1) {
    DISPATCH_HALF(dtype, TT, {
      if (D == 128) split_reduce_kernel<TT, 128><<<dim3(B, Hq), 128, 0, st>>>(part_out, part_lse, (TT*)out, lse_out, Hq, p.num_splits, out_stride);
      else if (D == 64) split_reduce_kernel<TT, 64><<<dim3(B, Hq), 64, 0, st>>>(part_out, part_lse, (TT*)out, lse_out, Hq, p.num_splits, out_stride);
      else if (D == 256) split_reduce_kernel<TT, 256><<<dim3(B, Hq), 256, 0, st>>>(part_out, part_lse, (TT*)out, lse_out, Hq, p.num_splits, out_stride);
      else split_reduce_kernel<TT, 32><<<dim3(B, Hq), 32, 0, st>>>(part_out, part_lse, (TT*)out, lse_out, Hq, p.num_splits, out_stride);
    });
  }
  return (int)cudaGetLastError();
}

EXPORT int vb_merge_attn_states(void* out, float* lse_out, const void* o1, const float* l1, const void* o2,
                                const float* l2, int64_t n_rows, int D, int dtype, cudaStream_t st) {
  if (n_rows == 0) return 0;
  DISPATCH_HALF(dtype, TT, {
    merge_states_kernel<TT><<<(unsigned)n_rows, 128, 0, st>>>((TT*)out, lse_out, (const TT*)o1, l1, (const TT*)o2, l2, D, n_rows);
  });
  return (int)cudaGetLastError();
}
