// Stand-alone collectives over the symmetric heap: all-reduce (+ residual add + RMSNorm), all-gather,
// barrier.  They cover every TP collective that is not already the epilogue of a GEMM (see gemm_ar.cu).
//
// all-reduce with fused residual + RMSNorm (collective C1 of SURVEY §2.4 followed by the next layer's norm):
//   phase 0  (optional) copy the local partial [T,H] into the symmetric staging region `in_off`
//   barrier  every rank's partial is in place
//   phase 1  rank r owns rows r*ceil(T/W)..: sum over ranks (multimem.ld_reduce in the switch, or P2P loads),
//            += residual (residual stays SHARDED by row owner — nobody else ever reads those rows),
//            RMSNorm, broadcast the normalised row to every rank (multimem.st / P2P stores) at `out_off`
//   barrier  all rows have landed everywhere → kernel completion means the output is usable
// The grid is persistent (<= #SMs CTAs looping over rows) so the intra-kernel barriers cannot deadlock.
#include "comm.cuh"

namespace {

// acc32_off != 0: the rows this rank owns were already summed into an fp32 accumulator at that offset of the LOCAL
// buffer (by red.add epilogues of every rank's grouped GEMM: fused MoE combine + reduce-scatter); phase 1 reads and
// re-zeroes it instead of reducing the staged 16-bit partials.
template <typename T>
__global__ void __launch_bounds__(1024) ar_rows_kernel(CommCtx c, const T* __restrict__ x, size_t in_off, size_t out_off,
                                                        T* __restrict__ residual, const T* __restrict__ norm_w,
                                                        float eps, int T_rows, int H, int gemma, size_t acc32_off) {
  __shared__ float red[32];
  const uint32_t e0 = c.state[0];
  const int vec_per_row = H / 8;
  // ---- phase 0: stage the local partial
  if (x != nullptr) {
    const size_t total = (size_t)T_rows * vec_per_row;
    uint4* dst = reinterpret_cast<uint4*>(c.local + in_off);
    const uint4* src = reinterpret_cast<const uint4*>(x);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
      dst[i] = src[i];
  }
  grid_rank_barrier(c, &c.state[1], &c.state[2], e0 + 1, true);
  // ---- phase 1: my rows
  const int per = (T_rows + c.world - 1) / c.world;
  const int r0 = c.rank * per, r1 = min(T_rows, r0 + per);
  for (int row = r0 + blockIdx.x; row < r1; row += gridDim.x) {
    float acc[8];
    float ss = 0.f;
    const bool active = threadIdx.x < vec_per_row;
    const size_t off = ((size_t)row * H + threadIdx.x * 8) * sizeof(T);
    if (active) {
      if (acc32_off) {
        float4* a4 = reinterpret_cast<float4*>(c.local + acc32_off) + ((size_t)row * H + threadIdx.x * 8) / 4;
        const float4 lo = a4[0], hi = a4[1];
        a4[0] = make_float4(0.f, 0.f, 0.f, 0.f); a4[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        acc[0] = lo.x; acc[1] = lo.y; acc[2] = lo.z; acc[3] = lo.w; acc[4] = hi.x; acc[5] = hi.y; acc[6] = hi.z; acc[7] = hi.w;
      } else {
        reduce8<T>(c, in_off + off, acc);
      }
      if (residual) {
        Vec16<T> rv = ld16(residual + (size_t)row * H + threadIdx.x * 8), ro;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          // reproduce the unfused numerics: the all-reduced value is rounded to T before the add
          ro.e[i] = from_f32<T>(to_f32(from_f32<T>(acc[i])) + to_f32(rv.e[i]));
          acc[i] = to_f32(ro.e[i]);
        }
        st16(residual + (size_t)row * H + threadIdx.x * 8, ro);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) ss += acc[i] * acc[i];
    }
    uint32_t o[4];
    if (norm_w) {
      ss = block_sum(ss, red);
      const float inv = rsqrtf(ss / (float)H + eps);
      if (active) {
        Vec16<T> wv = ld16(norm_w + threadIdx.x * 8);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float g0 = to_f32(wv.e[2 * i]) + (gemma ? 1.f : 0.f), g1 = to_f32(wv.e[2 * i + 1]) + (gemma ? 1.f : 0.f);
          o[i] = pack2<T>(acc[2 * i] * inv * g0, acc[2 * i + 1] * inv * g1);
        }
      }
    } else if (active) {
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = pack2<T>(acc[2 * i], acc[2 * i + 1]);
    }
    if (active) bcast16(c, out_off + off, o);
  }
  // ---- everyone's rows have been written everywhere
  grid_rank_barrier(c, &c.state[3], &c.state[4], e0 + 2, false, true);
}

// One-shot all-reduce into a PRIVATE destination (small T: decode-size activations, embeddings): every rank pulls all
// W replicas of every row over NVLink (all loads in flight at once), sums them in RANK ORDER — so every GPU produces
// bit-identical rows, which a switch-side multimem.ld_reduce issued from different GPUs does not promise — and writes
// them straight to `dst`: no broadcast phase and no copy out of the symmetric heap afterwards.
// The trailing (non-waiting) barrier only keeps the staging region from being overwritten while a peer still reads.
template <typename T>
__global__ void __launch_bounds__(1024) ar_oneshot_kernel(CommCtx c, const T* __restrict__ x, size_t in_off,
                                                           T* __restrict__ dst, int T_rows, int H) {
  const uint32_t e0 = c.state[0];
  const int vec_per_row = H / 8;
  if (x != nullptr) {
    const size_t total = (size_t)T_rows * vec_per_row;
    uint4* st = reinterpret_cast<uint4*>(c.local + in_off);
    const uint4* src = reinterpret_cast<const uint4*>(x);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
      st[i] = src[i];
  }
  grid_rank_barrier(c, &c.state[1], &c.state[2], e0 + 1, true);
  if (threadIdx.x < vec_per_row) {
    for (int row = blockIdx.x; row < T_rows; row += gridDim.x) {
      float acc[8];
      const size_t el = (size_t)row * H + threadIdx.x * 8;
      uint4 w[VB_MAX_RANKS];
#pragma unroll
      for (int p = 0; p < VB_MAX_RANKS; ++p)
        if (p < c.world)   // .sys loads: peer lines may linger in this SM's L1 from the previous call otherwise
          asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(w[p].x), "=r"(w[p].y), "=r"(w[p].z), "=r"(w[p].w)
                       : "l"(c.peer[p] + in_off + el * sizeof(T)) : "memory");
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
      for (int p = 0; p < VB_MAX_RANKS; ++p)
        if (p < c.world) {
          const uint32_t u[4] = {w[p].x, w[p].y, w[p].z, w[p].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) { float2 f = unpack2<T>(u[i]); acc[2 * i] += f.x; acc[2 * i + 1] += f.y; }
        }
      Vec16<T> o;
#pragma unroll
      for (int i = 0; i < 8; ++i) o.e[i] = from_f32<T>(acc[i]);
      st16(dst + el, o);
    }
  }
  grid_rank_barrier(c, &c.state[3], &c.state[4], e0 + 2, false, true);
}

// all-gather along the last dim: out[row, rank*cols + j] = x[row, j] on every rank
template <typename T>
__global__ void all_gather_kernel(CommCtx c, const T* __restrict__ x, size_t out_off, int rows, int cols) {
  const uint32_t e0 = c.state[0];
  const int vec_per_row = cols / 8;
  const size_t total = (size_t)rows * vec_per_row;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / vec_per_row, v = i % vec_per_row;
    const uint4 u = *reinterpret_cast<const uint4*>(x + row * cols + v * 8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    bcast16(c, out_off + ((row * c.world + c.rank) * cols + v * 8) * sizeof(T), w);
  }
  grid_rank_barrier(c, &c.state[1], &c.state[2], e0 + 1, false, true);
}

__global__ void barrier_kernel(CommCtx c) {
  const uint32_t e0 = c.state[0];
  grid_rank_barrier(c, &c.state[1], &c.state[2], e0 + 1, false, true);
}

}  // namespace

static CommCtx make_ctx(void* local, void* const* peers, void* mc, uint32_t* state, int rank, int world) {
  CommCtx c;
  c.local = (unsigned char*)local;
  for (int i = 0; i < VB_MAX_RANKS; ++i) c.peer[i] = i < world ? (unsigned char*)peers[i] : nullptr;
  c.mc = (unsigned char*)mc;
  c.state = state;
  c.rank = rank;
  c.world = world;
  return c;
}

// x may be null when the partial already lives in the symmetric buffer at in_off (e.g. written by a GEMM epilogue)
EXPORT int vb_comm_allreduce_rows2(void* local, void* const* peers, void* mc, uint32_t* state, int rank, int world,
                                   const void* x, int64_t in_off, int64_t out_off, void* residual, const void* norm_w,
                                   float eps, int T_rows, int H, int gemma, int num_sms, int64_t acc32_off, int dtype,
                                   cudaStream_t st);
EXPORT int vb_comm_allreduce_rows(void* local, void* const* peers, void* mc, uint32_t* state, int rank, int world,
                                  const void* x, int64_t in_off, int64_t out_off, void* residual, const void* norm_w,
                                  float eps, int T_rows, int H, int gemma, int num_sms, int dtype, cudaStream_t st) {
  return vb_comm_allreduce_rows2(local, peers, mc, state, rank, world, x, in_off, out_off, residual, norm_w, eps, T_rows, H,
                                 gemma, num_sms, 0, dtype, st);
}
EXPORT int vb_comm_allreduce_rows2(void* local, void* const* peers, void* mc, uint32_t* state, int rank, int world,
                                   const void* x, int64_t in_off, int64_t out_off, void* residual, const void* norm_w,
                                   float eps, int T_rows, int H, int gemma, int num_sms, int64_t acc32_off, int dtype,
                                   cudaStream_t st) {
  if (T_rows == 0) return 0;
  if (dtype == DT_F32 || H % 8 != 0 || H / 8 > 1024) return -1;
  CommCtx c = make_ctx(local, peers, mc, state, rank, world);
  const int per = (T_rows + world - 1) / world;
  int grid = per < num_sms ? per : num_sms;
  if (grid < 1) grid = 1;
  int thr = ((H / 8 + 31) / 32) * 32;
  DISPATCH_HALF(dtype, TT, {
    ar_rows_kernel<TT><<<grid, thr, 0, st>>>(c, (const TT*)x, (size_t)in_off, (size_t)out_off, (TT*)residual,
                                              (const TT*)norm_w, eps, T_rows, H, gemma, (size_t)acc32_off);
  });
  return (int)cudaGetLastError();
}

EXPORT int vb_comm_allreduce_oneshot(void* local, void* const* peers, void* mc, uint32_t* state, int rank, int world,
                                     const void* x, int64_t in_off, void* dst, int T_rows, int H, int num_sms, int dtype,
                                     cudaStream_t st) {
  if (T_rows == 0) return 0;
  if (dtype == DT_F32 || H % 8 != 0 || H / 8 > 1024) return -1;
  CommCtx c = make_ctx(local, peers, mc, state, rank, world);
  int grid = T_rows < num_sms ? T_rows : num_sms;
  int thr = ((H / 8 + 31) / 32) * 32;
  DISPATCH_HALF(dtype, TT, {
    ar_oneshot_kernel<TT><<<grid, thr, 0, st>>>(c, (const TT*)x, (size_t)in_off, (TT*)dst, T_rows, H);
  });
  return (int)cudaGetLastError();
}

EXPORT int vb_comm_all_gather(void* local, void* const* peers, void* mc, uint32_t* state, int rank, int world,
                              const void* x, int64_t out_off, int rows, int cols, int num_sms, int dtype,
                              cudaStream_t st) {
  if (rows == 0) return 0;
  if (dtype == DT_F32 || cols % 8 != 0) return -1;
  CommCtx c = make_ctx(local, peers, mc, state, rank, world);
  const size_t total = (size_t)rows * (cols / 8);
  int grid = (int)((total + 255) / 256);
  if (grid > num_sms) grid = num_sms;
  DISPATCH_HALF(dtype, TT, {
    all_gather_kernel<TT><<<grid, 256, 0, st>>>(c, (const TT*)x, (size_t)out_off, rows, cols);
  });
  return (int)cudaGetLastError();
}

EXPORT int vb_comm_barrier(void* local, void* const* peers, void* mc, uint32_t* state, int rank, int world,
                           cudaStream_t st) {
  CommCtx c = make_ctx(local, peers, mc, state, rank, world);
  barrier_kernel<<<1, 32, 0, st>>>(c);
  return (int)cudaGetLastError();
}
