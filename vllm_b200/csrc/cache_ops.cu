// KV-cache maintenance kernels: block copy / swap, gather (optionally de-quantising) into contiguous workspaces,
// fp8 conversion.
//
// Reference ops: csrc/cache_kernels.cu — swap_blocks / copy_blocks(_mla), gather_and_maybe_dequant_cache,
// cp_gather_cache, convert_fp8 (SURVEY §2.3 "cache kernels").  Ours: every op is ONE launch driven by a device-
// resident (src, dst) index list — no per-block cudaMemcpy, graph capturable; blocks move as 16-byte vectors with
// one CTA (or several, for big pages) per block pair.  The host <-> device direction of swap_blocks uses pinned
// host memory mapped into the device address space, so the same kernel serves the offload tier.
#include "common.cuh"

namespace {

// dst[map[2i+1]] = src[map[2i]] for whole blocks of `block_bytes` (multiple of 16)
__global__ void __launch_bounds__(256) copy_blocks_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                                           const int64_t* __restrict__ map, int64_t vec_per_block,
                                                           int chunks) {
  pdl_trigger();
  pdl_wait();
  const int pair = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
  const int64_t s = map[2 * pair], d = map[2 * pair + 1];
  if (s < 0 || d < 0) return;
  const uint4* sp = src + s * vec_per_block;
  uint4* dp = dst + d * vec_per_block;
  const int64_t per = (vec_per_block + chunks - 1) / chunks;
  const int64_t lo = chunk * per, hi = min(vec_per_block, lo + per);
  // 4 independent 16-byte loads in flight per thread
  int64_t i = lo + threadIdx.x;
  for (; i + 3 * (int64_t)blockDim.x < hi; i += 4 * (int64_t)blockDim.x) {
    uint4 a = sp[i], b = sp[i + blockDim.x], c = sp[i + 2 * blockDim.x], e = sp[i + 3 * blockDim.x];
    dp[i] = a; dp[i + blockDim.x] = b; dp[i + 2 * blockDim.x] = c; dp[i + 3 * blockDim.x] = e;
  }
  for (; i < hi; i += blockDim.x) dp[i] = sp[i];
}

// Gather the KV rows of a batch of sequences out of the paged cache into dense [total_tokens, row_elems] workspaces
// (chunked-prefill context for MLA / cross-layer consumers).  cache: [num_blocks, BS, row] (one row per token: the
// MLA latent layout, or K / V planes addressed through `plane_stride`).  FP8 → T conversion with `scale` when SRC8.
template <typename T, bool SRC8>
__global__ void __launch_bounds__(256) gather_cache_kernel(const void* __restrict__ cache, T* __restrict__ out,
                                                            const int* __restrict__ block_table, int64_t bt_stride,
                                                            const int* __restrict__ cu_seq_lens,
                                                            const int* __restrict__ seq_starts, int BS, int row,
                                                            int64_t block_stride, int64_t out_stride, float scale,
                                                            int tokens_per_cta) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const int t0 = cu_seq_lens[b], n = cu_seq_lens[b + 1] - t0;
  const int start = seq_starts ? seq_starts[b] : 0;          // first cached position to gather (chunked context)
  const int first = blockIdx.x * tokens_per_cta;
  if (first >= n) return;
  const int last = min(n, first + tokens_per_cta);
  const int vec_per_row = row / 8;
  for (int i = first * vec_per_row + threadIdx.x; i < last * vec_per_row; i += blockDim.x) {
    const int tok = i / vec_per_row, c = i % vec_per_row;
    const int pos = start + tok;
    const int64_t blk = block_table[b * bt_stride + pos / BS];
    const int64_t src_off = blk * block_stride + (int64_t)(pos % BS) * row + c * 8;
    Vec16<T> o;
    if (SRC8) {
      const uint2 raw = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(cache) + src_off);
      const __nv_fp8_storage_t* b8 = reinterpret_cast<const __nv_fp8_storage_t*>(&raw);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const __half_raw hr = __nv_cvt_fp8_to_halfraw(b8[j], __NV_E4M3);
        o.e[j] = from_f32<T>(__half2float(__half(hr)) * scale);
      }
    } else {
      o = ld16(reinterpret_cast<const T*>(cache) + src_off);
    }
    st16(out + (int64_t)(t0 + tok) * out_stride + c * 8, o);
  }
}

template <typename T, bool TO8>
__global__ void __launch_bounds__(256) convert_fp8_kernel(const void* __restrict__ src, void* __restrict__ dst, int64_t n,
                                                           float scale) {
  const float inv = 1.f / scale;
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 8; i < n; i += (int64_t)gridDim.x * blockDim.x * 8) {
    if (TO8) {
      Vec16<T> a = ld16(reinterpret_cast<const T*>(src) + i);
      union { uint2 u; __nv_fp8_storage_t b[8]; } o;
#pragma unroll
      for (int j = 0; j < 8; ++j) o.b[j] = __nv_cvt_float_to_fp8(to_f32(a.e[j]) * inv, __NV_SATFINITE, __NV_E4M3);
      *reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(dst) + i) = o.u;
    } else {
      const uint2 raw = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(src) + i);
      const __nv_fp8_storage_t* b8 = reinterpret_cast<const __nv_fp8_storage_t*>(&raw);
      Vec16<T> o;
#pragma unroll
      for (int j = 0; j < 8; ++j) o.e[j] = from_f32<T>(__half2float(__half(__nv_cvt_fp8_to_halfraw(b8[j], __NV_E4M3))) * scale);
      st16(reinterpret_cast<T*>(dst) + i, o);
    }
  }
}

}  // namespace

// `map`: device int64 [n_pairs, 2] = (src block, dst block).  src / dst may be the same tensor (copy_blocks), two
// device tensors (migration) or a device tensor and a mapped pinned-host tensor (swap in / out).
EXPORT int vb_copy_blocks(const void* src, void* dst, const int64_t* map, int n_pairs, int64_t block_bytes, cudaStream_t st) {
  if (n_pairs == 0) return 0;
  if (block_bytes % 16) return -1;
  const int64_t vpb = block_bytes / 16;
  int chunks = (int)((vpb + 8191) / 8192);               // ~128 KB per CTA
  if (chunks < 1) chunks = 1;
  vb_launch(copy_blocks_kernel, dim3((unsigned)(n_pairs * chunks)), dim3(256), 0, st, reinterpret_cast<const uint4*>(src),
            reinterpret_cast<uint4*>(dst), map, vpb, chunks);
  return (int)cudaGetLastError();
}

EXPORT int vb_gather_cache(const void* cache, void* out, const int* block_table, int64_t bt_stride, const int* cu_seq_lens,
                           const int* seq_starts, int batch, int max_tokens, int BS, int row, int64_t block_stride,
                           int64_t out_stride, float scale, int src_fp8, int dtype, cudaStream_t st) {
  if (batch == 0 || max_tokens == 0) return 0;
  if (dtype == DT_F32 || row % 8) return -1;
  const int tpc = 16;
  dim3 grid((max_tokens + tpc - 1) / tpc, batch);
  DISPATCH_HALF(dtype, TT, {
    if (src_fp8) vb_launch(gather_cache_kernel<TT, true>, grid, dim3(256), 0, st, cache, reinterpret_cast<TT*>(out),
                           block_table, bt_stride, cu_seq_lens, seq_starts, BS, row, block_stride, out_stride, scale, tpc);
    else vb_launch(gather_cache_kernel<TT, false>, grid, dim3(256), 0, st, cache, reinterpret_cast<TT*>(out), block_table,
                   bt_stride, cu_seq_lens, seq_starts, BS, row, block_stride, out_stride, scale, tpc);
  });
  return (int)cudaGetLastError();
}

EXPORT int vb_convert_fp8(const void* src, void* dst, int64_t n, float scale, int to_fp8, int num_sms, int dtype,
                          cudaStream_t st) {
  if (n == 0) return 0;
  if (dtype == DT_F32 || n % 8) return -1;
  int64_t blocks = (n / 8 + 255) / 256;
  if (blocks > (int64_t)num_sms * 8) blocks = (int64_t)num_sms * 8;
  DISPATCH_HALF(dtype, TT, {
    if (to_fp8) convert_fp8_kernel<TT, true><<<(unsigned)blocks, 256, 0, st>>>(src, dst, n, scale);
    else convert_fp8_kernel<TT, false><<<(unsigned)blocks, 256, 0, st>>>(src, dst, n, scale);
  });
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------ L2 weight prefetch (side stream)
// Decode GEMMs with few tiles (qkv, o_proj) are dominated by first-byte latency, not bandwidth, while the HBM-bound
// attention kernel that precedes them leaves ~35 % of the HBM bandwidth idle.  This kernel, launched on a forked
// stream when attention starts, pulls the next GEMM's weights into the 126 MB L2 (cp.async.bulk.prefetch.L2, no SM
// resources beyond a few warps), so the GEMM's TMA stream then hits L2.
namespace {
__global__ void __launch_bounds__(256) l2_prefetch_kernel(const unsigned char* __restrict__ p, int64_t bytes, int chunk) {
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * chunk;
  if (i >= bytes) return;
  const int64_t n = min((int64_t)chunk, bytes - i) & ~(int64_t)15;
  if (n > 0)
    asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(p + i), "r"((uint32_t)n), "l"(L2_EVICT_LAST)
                 : "memory");
}
}  // namespace

EXPORT int vb_l2_prefetch(const void* p, int64_t bytes, cudaStream_t st) {
  if (bytes <= 0) return 0;
  const int chunk = 8192;
  const int64_t threads = (bytes + chunk - 1) / chunk;
  l2_prefetch_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(reinterpret_cast<const unsigned char*>(p), bytes, chunk);
  return (int)cudaGetLastError();
}
