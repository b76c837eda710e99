// RowParallel GEMM + tensor-parallel all-reduce + residual add + RMSNorm in ONE kernel (decode shapes).
//
//   normed[T,H], residual[T,H] = RMSNorm( sum_ranks( X_r[T,K_r] · W_r[H,K_r]^T ) + residual )
//
// This is the hot communication-bound path of TP decode (o_proj / down_proj followed by the
// next layer's fused_add_rms_norm).  The reference runs cuBLASLt, then a separate FlashInfer/custom
// all-reduce(+norm) kernel; here the collective is the GEMM's epilogue:
//
//   1. tcgen05 swap-AB GEMM (weights stream through TMA, fp32 accumulators in TMEM, split-K across CTAs
//      so all SMs pull weights; deterministic last-arriver fix-up through an L2-resident workspace).
//   2. The CTA that completes an output tile converts it to bf16, transposes it through shared memory and
//      PUSHES it into every peer's receive buffer with `multimem.st` (one NVSwitch multicast store reaches
//      all GPUs) — or per-peer P2P stores when multicast is unavailable.  No barrier, no flag: the
//      receive buffers use the Lamport convention — they are pre-filled with a sentinel (bf16 -0.0) and a
//      16-byte granule is valid as soon as it contains no sentinel.
//   3. CTAs that are done with their tile turn into reducers: row t is owned by CTA t; it polls the W
//      replicas of that row, sums them in fp32, adds the residual, applies RMSNorm and writes the normalised
//      row (every rank computes every row — the data is already local, and it removes the second
//      communication phase entirely), then re-arms the buffer that will be used two calls later.
//
// Three rotating buffers indexed by a device-side call counter make the kernel CUDA-graph replayable with no
// host involvement.  Exposed latency is one NVLink traversal (~1.5 us) instead of two ~5 us barriers.
#include "comm.cuh"

namespace gemm_ar {

constexpr int BK = 64, TM = 128, NUM_THREADS = 192, EPI_THREADS = 128;
constexpr uint32_t SENT32 = 0x80008000u;   // two bf16/fp16 negative zeros

struct Params {
  void* out;               // [T, H] normalised output (local)
  void* residual;          // [T, H] in/out (local, replicated on every rank)
  const void* norm_w;      // [H]
  float* ws;               // split-K workspace
  unsigned* counters;      // per-tile arrival counters (self-resetting)
  size_t lam_off;          // byte offset of the 3 Lamport buffers inside the symmetric heap
  size_t lam_stride;       // bytes per buffer (= world * t_cap * H * 2)
  int t_cap;               // rows reserved per source rank in a buffer (layout is independent of M)
  int mode;                // 0 = one-shot (every rank receives and reduces everything), 1 = two-shot (row owners
                           //     reduce + normalise their rows, then multicast them): inbound traffic independent of W
  size_t oring_off;        // mode 1: byte offset of the 3 rotating [t_cap][N] output buffers (sentinel-armed)
  size_t oring_stride;
  int M, N, K;             // tokens, H, local K
  int k_blocks_per_split, splits;
  float eps;
  int gemma;
  unsigned long long* trace;   // optional [n_ctas][8] globaltimer stamps of the epilogue phases (benchmarks/comm_probe.py)
};

VB_DEVICE unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// A reducer that spins on data another CTA / GPU must produce cannot make progress if that producer never runs
// (a peer died, SMs held by another stream): after SPIN_LIMIT_NS it reports and traps instead of hanging the box.
constexpr unsigned long long SPIN_LIMIT_NS = 20ull * 1000 * 1000 * 1000;
VB_DEVICE void spin_guard(uint32_t& n, unsigned long long t0, int what) {
  if ((++n & 0xfffffu) == 0 && gtime() - t0 > SPIN_LIMIT_NS) {
    printf("gemm_ar: rank-visible spin limit exceeded (phase %d, block %d,%d)\n", what, blockIdx.x, blockIdx.z);
    __trap();
  }
}

template <int BN> struct SmemLayout {
  static constexpr int A_BYTES = TM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
};

VB_DEVICE void ld_relaxed_sys_v4(const void* p, uint32_t (&w)[4]) {
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(p) : "memory");
}
VB_DEVICE bool has_sentinel(const uint32_t (&w)[4]) {
  bool bad = false;
#pragma unroll
  for (int i = 0; i < 4; ++i) bad |= ((w[i] & 0xffffu) == 0x8000u) | ((w[i] >> 16) == 0x8000u);
  return bad;
}
VB_DEVICE uint32_t sanitize(uint32_t w) {   // -0.0 → +0.0 in both halves
  if ((w & 0xffffu) == 0x8000u) w &= 0xffff0000u;
  if ((w >> 16) == 0x8000u) w &= 0x0000ffffu;
  return w;
}

// 16-byte store into ONE peer's buffer (two-shot mode: a partial row goes only to the rank that owns the row)
VB_DEVICE void put16(const CommCtx& c, int dst, size_t off, const uint32_t (&v)[4]) {
  *reinterpret_cast<uint4*>(c.peer[dst] + off) = make_uint4(v[0], v[1], v[2], v[3]);
}

template <typename T, int BN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_ar_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const Params p,
               const CommCtx c) {
  using L = SmemLayout<BN>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* stage_base = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * L::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  int* is_last = reinterpret_cast<int*>(tmem_slot + 1);
  float* red = reinterpret_cast<float*>(is_last + 1);      // [8]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a_tile = blockIdx.x, split = blockIdx.z;
  const int kb0 = split * p.k_blocks_per_split;
  const int total_kb = (p.K + BK - 1) / BK;
  const int nkb = max(0, min(p.k_blocks_per_split, total_kb - kb0));
  constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));

  pdl_trigger();
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
      const int pre = min(nkb, STAGES);
      for (int kb = 0; kb < pre; ++kb) {   // weight tiles do not depend on the previous kernel (PDL prefetch)
        mbar_expect_tx(&full[kb], L::STAGE_BYTES);
        tma_load_2d_hint(stage_base + kb * L::STAGE_BYTES, &tmap_a, (kb0 + kb) * BK, a_tile * TM, &full[kb], L2_EVICT_FIRST);
      }
      pdl_wait();
      for (int kb = 0; kb < pre; ++kb)
        tma_load_2d_hint(stage_base + kb * L::STAGE_BYTES + L::A_BYTES, &tmap_b, (kb0 + kb) * BK, 0, &full[kb], L2_EVICT_LAST);
      for (int kb = pre; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&empty[s], ((kb / STAGES) & 1) ^ 1);
        unsigned char* a_dst = stage_base + s * L::STAGE_BYTES;
        mbar_expect_tx(&full[s], L::STAGE_BYTES);
        tma_load_2d_hint(a_dst, &tmap_a, (kb0 + kb) * BK, a_tile * TM, &full[s], L2_EVICT_FIRST);
        tma_load_2d_hint(a_dst + L::A_BYTES, &tmap_b, (kb0 + kb) * BK, 0, &full[s], L2_EVICT_LAST);
      }
    }
    if (p.splits > 1) { __syncwarp(); cluster_sync(); cluster_sync(); }   // the epilogue's two cluster barriers
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_f16(TM, BN, DTypeOf<T>::v == DT_BF16);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      mbar_wait(&full[s], (kb / STAGES) & 1);
      tc_fence_after();
      if (lane == 0) {
        const unsigned char* a_src = stage_base + s * L::STAGE_BYTES;
        const uint64_t da = umma_desc_sw128(a_src), db = umma_desc_sw128(a_src + L::A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_f16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        umma_commit(&empty[s]);
        if (kb == nkb - 1) umma_commit(tmem_full);
      }
      __syncwarp();
    }
    if (p.splits > 1) { __syncwarp(); cluster_sync(); cluster_sync(); }   // the epilogue's two cluster barriers
  } else {
    // =========================== epilogue + collective ===========================
    constexpr int CH = (BN % 32 == 0) ? 32 : 16;
    const int q = warp & 3;
    const int row = q * 32 + lane;                         // feature inside the tile
    const int et = threadIdx.x - 64;                       // 0..127 among the epilogue threads
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    const int n_ctas = gridDim.x * gridDim.z;
    const int cta = blockIdx.z * gridDim.x + blockIdx.x;
    const unsigned long long t_begin = gtime();
    uint32_t spins = 0;
#define VB_STAMP(i) do { if (p.trace && et == 0) p.trace[cta * 8 + (i)] = gtime(); } while (0)
    VB_STAMP(0);
    pdl_wait();
    VB_STAMP(1);
    // device-side state (slots 8..12 of c.state; 0..7 belong to comm.cu): [8] call counter → buffer rotation,
    // [9] CTA-done counter, [10+b] rows dirtied in buffer b by its last use
    const uint32_t call = c.state[8];
    const int dirty = (int)c.state[10 + (call + 2) % 3];   // rows to re-arm in the buffer used two calls from now
    unsigned char* lam_cur = c.local + p.lam_off + (size_t)(call % 3) * p.lam_stride;
    unsigned char* lam_clear = c.local + p.lam_off + (size_t)((call + 2) % 3) * p.lam_stride;
    const size_t lam_cur_off = p.lam_off + (size_t)(call % 3) * p.lam_stride;
    if (nkb > 0) { mbar_wait(tmem_full, 0); tc_fence_after(); }
    VB_STAMP(2);
    auto load_chunk = [&](int cc, float (&v)[CH]) {
      if (nkb == 0) {
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] = 0.f;
        return;
      }
      if constexpr (CH == 32) {
        uint32_t r[32];
        tmem_ld_32x32(t_row + cc, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      } else {
        uint32_t r[16];
        tmem_ld_32x16(t_row + cc, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
      }
    };
    const int tile_id = a_tile;
    const int n0 = a_tile * TM;
    bool do_push = true;
    if (p.splits > 1) {
      // split-K over a thread-block cluster (grid z): every CTA parks its fp32 partial in its own shared memory
      // (the operand ring is idle once tmem_full fired); after the cluster barrier EACH CTA reduces 1/splits of
      // the tile out of its peers' shared memory (DSMEM, split order → deterministic) and ships that slice.
      // No global workspace, no atomics; the fix-up and the push are spread over all SMs of the cluster.
      do_push = false;
      float* part_s = reinterpret_cast<float*>(stage_base);
      for (int cc = 0; cc < BN; cc += CH) {
        float v[CH];
        load_chunk(cc, v);
#pragma unroll
        for (int j = 0; j < CH; ++j) part_s[(cc + j) * TM + row] = v[j];
      }
      __syncwarp();
      cluster_sync();
      const int units = p.M * (TM / 8);                            // unit = 8 consecutive features of one token
      const int per = (units + p.splits - 1) / p.splits;
      const int u_end = min(units, (split + 1) * per);
      const uint32_t red_s = smem_u32(stage_base);
      const int own_per = (p.M + c.world - 1) / c.world;           // two-shot: rows owned by each rank
      for (int u = split * per + et; u < u_end; u += EPI_THREADS) {
        const int t = u / (TM / 8), g = u % (TM / 8);
        if (n0 + g * 8 >= p.N) continue;
        const uint32_t off = red_s + (uint32_t)(t * TM + g * 8) * 4u;
        float4 x0[8], x1[8];
#pragma unroll
        for (int r = 0; r < 8; ++r)
          if (r < p.splits) {
            const uint32_t ra = dsmem_addr(off, (uint32_t)r);
            x0[r] = ld_dsmem_f32x4(ra); x1[r] = ld_dsmem_f32x4(ra + 16);
          }
        float4 a0 = x0[0], a1 = x1[0];
#pragma unroll
        for (int r = 1; r < 8; ++r)
          if (r < p.splits) {
            a0.x += x0[r].x; a0.y += x0[r].y; a0.z += x0[r].z; a0.w += x0[r].w;
            a1.x += x1[r].x; a1.y += x1[r].y; a1.z += x1[r].z; a1.w += x1[r].w;
          }
        uint32_t w[4] = {pack2<T>(a0.x, a0.y), pack2<T>(a0.z, a0.w), pack2<T>(a1.x, a1.y), pack2<T>(a1.z, a1.w)};
#pragma unroll
        for (int e = 0; e < 4; ++e) w[e] = sanitize(w[e]);
        const size_t doff = lam_cur_off + (((size_t)c.rank * p.t_cap + t) * p.N + n0 + g * 8) * sizeof(T);
        if (p.mode == 1) put16(c, t / own_per, doff, w);
        else bcast16(c, doff, w);
      }
      __syncwarp();
      cluster_sync();                            // nobody's shared memory is read past this point
    }
    // the operand ring is idle now (all MMAs of this CTA retired before tmem_full fired): reuse it as the
    // [token][feature] transposition buffer
    T* tbuf = reinterpret_cast<T*>(stage_base);
    if (do_push) {
      for (int cc = 0; cc < BN; cc += CH) {
        float v[CH];
        load_chunk(cc, v);
#pragma unroll
        for (int j = 0; j < CH; ++j)
          if (cc + j < p.M) tbuf[(cc + j) * TM + row] = from_f32<T>(v[j]);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS));
      const int own_per = (p.M + c.world - 1) / c.world;
      // push: item = (token, group of 8 features); wire layout [src rank][token][H]
      for (int i = et; i < p.M * (TM / 8); i += EPI_THREADS) {
        const int t = i / (TM / 8), g = i % (TM / 8);
        if (n0 + g * 8 >= p.N) continue;
        uint4 u = *reinterpret_cast<const uint4*>(tbuf + t * TM + g * 8);
        const uint32_t w[4] = {sanitize(u.x), sanitize(u.y), sanitize(u.z), sanitize(u.w)};
        const size_t off = lam_cur_off + (((size_t)c.rank * p.t_cap + t) * p.N + n0 + g * 8) * sizeof(T);
        if (p.mode == 1) put16(c, t / own_per, off, w);
        else bcast16(c, off, w);
      }
    }
    tc_fence_before();
    VB_STAMP(3);
    // ---- reducer: CTA `cta` owns rows cta, cta + n_ctas, ...
    const int groups = p.N / 8;                            // 16-byte groups per row
    constexpr int MAXG = 8;                                // groups per thread kept in registers (H <= 8192)
    const bool two = p.mode == 1;
    const size_t oring_cur = p.oring_off + (size_t)(call % 3) * p.oring_stride;
    if (!two) {
    // One-shot: this CTA reduces rows cta, cta + n_ctas, ... of ALL tokens (every rank does the same work on the
    // same data → bit-identical rows everywhere).
    const int t_end = max(p.M, dirty);
    for (int t = cta; t < t_end; t += n_ctas) {
      if (t >= p.M) {   // row not used by this call but dirtied earlier: only re-arm it
        for (int g = et; g < groups * c.world; g += EPI_THREADS) {
          const int src = g / groups, gg = g % groups;
          *reinterpret_cast<uint4*>(lam_clear + (((size_t)src * p.t_cap + t) * p.N + gg * 8) * sizeof(T)) =
              make_uint4(SENT32, SENT32, SENT32, SENT32);
        }
        continue;
      }
      float acc[MAXG][8];
      float ss = 0.f;
#pragma unroll
      for (int gi = 0; gi < MAXG; ++gi) {
        const int g = et + gi * EPI_THREADS;
        if (g >= groups) break;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[gi][e] = 0.f;
        // issue the loads of all replicas first (memory-level parallelism), then re-poll only the late ones;
        // the sum is always taken in rank order so every GPU produces bit-identical rows
        uint32_t w[VB_MAX_RANKS][4];
#pragma unroll
        for (int src = 0; src < VB_MAX_RANKS; ++src)
          if (src < c.world)
            ld_relaxed_sys_v4(lam_cur + (((size_t)src * p.t_cap + t) * p.N + g * 8) * sizeof(T), w[src]);
#pragma unroll
        for (int src = 0; src < VB_MAX_RANKS; ++src) {
          if (src >= c.world) break;
          const unsigned char* ptr = lam_cur + (((size_t)src * p.t_cap + t) * p.N + g * 8) * sizeof(T);
          while (has_sentinel(w[src])) { spin_guard(spins, t_begin, 0); ld_relaxed_sys_v4(ptr, w[src]); }
#pragma unroll
          for (int e = 0; e < 4; ++e) { float2 f = unpack2<T>(w[src][e]); acc[gi][2 * e] += f.x; acc[gi][2 * e + 1] += f.y; }
          // re-arm the same slot of the buffer used two calls from now
          *reinterpret_cast<uint4*>(lam_clear + (((size_t)src * p.t_cap + t) * p.N + g * 8) * sizeof(T)) =
              make_uint4(SENT32, SENT32, SENT32, SENT32);
        }
        T* rrow = reinterpret_cast<T*>(p.residual) + (size_t)t * p.N + g * 8;
        Vec16<T> rv = ld16(rrow), ro;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          ro.e[e] = from_f32<T>(to_f32(from_f32<T>(acc[gi][e])) + to_f32(rv.e[e]));
          acc[gi][e] = to_f32(ro.e[e]);
          ss += acc[gi][e] * acc[gi][e];
        }
        st16(rrow, ro);
      }
      // row-wide sum of squares over the 128 epilogue threads
      ss = warp_sum(ss);
      if (lane == 0) red[q] = ss;
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS));
      const float tot = red[0] + red[1] + red[2] + red[3];
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS));
      const float inv = rsqrtf(tot / (float)p.N + p.eps);
#pragma unroll
      for (int gi = 0; gi < MAXG; ++gi) {
        const int g = et + gi * EPI_THREADS;
        if (g >= groups) break;
        Vec16<T> wv = ld16(reinterpret_cast<const T*>(p.norm_w) + g * 8), o;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float gam = to_f32(wv.e[e]) + (p.gemma ? 1.f : 0.f);
          o.e[e] = from_f32<T>(acc[gi][e] * inv * gam);
        }
        st16(reinterpret_cast<T*>(p.out) + (size_t)t * p.N + g * 8, o);
      }
    }
    } else {
      // Two-shot.  (A) the rows this rank owns, [r0, r1), are summed (+ residual) by ALL CTAs of the grid — one
      // 16-byte group of one row per thread — and the new residual-stream value h goes to every rank through the
      // output ring (NVLS multicast).  (B) every rank collects all rows of h and normalises them itself: the
      // RMSNorm is replicated (cheap) instead of serialising a row behind one CTA, and all ranks still produce
      // bit-identical rows because they normalise the same bf16 h.  The residual stays sharded by row owner.
      const int own_per = (p.M + c.world - 1) / c.world;
      const int r0 = min(p.M, c.rank * own_per), r1 = min(p.M, r0 + own_per);
      for (int u = cta * EPI_THREADS + et; u < (r1 - r0) * groups; u += n_ctas * EPI_THREADS) {
        const int t = r0 + u / groups, g = u % groups;
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        uint32_t w[VB_MAX_RANKS][4];
#pragma unroll
        for (int src = 0; src < VB_MAX_RANKS; ++src)
          if (src < c.world)
            ld_relaxed_sys_v4(lam_cur + (((size_t)src * p.t_cap + t) * p.N + g * 8) * sizeof(T), w[src]);
        T* rrow = reinterpret_cast<T*>(p.residual) + (size_t)t * p.N + g * 8;
        const Vec16<T> rv = ld16(rrow);
#pragma unroll
        for (int src = 0; src < VB_MAX_RANKS; ++src) {
          if (src >= c.world) break;
          const unsigned char* ptr = lam_cur + (((size_t)src * p.t_cap + t) * p.N + g * 8) * sizeof(T);
          while (has_sentinel(w[src])) { spin_guard(spins, t_begin, 1); ld_relaxed_sys_v4(ptr, w[src]); }
#pragma unroll
          for (int e = 0; e < 4; ++e) { float2 f = unpack2<T>(w[src][e]); acc[2 * e] += f.x; acc[2 * e + 1] += f.y; }
        }
        Vec16<T> ro;
#pragma unroll
        for (int e = 0; e < 8; ++e) ro.e[e] = from_f32<T>(to_f32(from_f32<T>(acc[e])) + to_f32(rv.e[e]));
        st16(rrow, ro);
        uint32_t o[4] = {sanitize(ro.u.x), sanitize(ro.u.y), sanitize(ro.u.z), sanitize(ro.u.w)};
        bcast16(c, oring_cur + ((size_t)t * p.N + g * 8) * sizeof(T), o);
      }
      VB_STAMP(4);
      // re-arm the buffers used two calls from now: the rows this rank owned under their previous token count
      // (input ring, all source ranks) and every row of the output ring
      {
        const int old_per = (dirty + c.world - 1) / c.world;
        const int o0 = min(dirty, c.rank * old_per), o1 = min(dirty, o0 + old_per), on = o1 - o0;
        unsigned char* oring_clear = c.local + p.oring_off + (size_t)((call + 2) % 3) * p.oring_stride;
        const int n_in = on * c.world * groups, n_all = n_in + dirty * groups;
        for (int u = cta * EPI_THREADS + et; u < n_all; u += n_ctas * EPI_THREADS) {
          unsigned char* dst;
          if (u < n_in) {
            const int rowi = u / groups, g = u % groups;
            dst = lam_clear + (((size_t)(rowi / on) * p.t_cap + o0 + rowi % on) * p.N + g * 8) * sizeof(T);
          } else {
            dst = oring_clear + (size_t)(u - n_in) * 16;
          }
          *reinterpret_cast<uint4*>(dst) = make_uint4(SENT32, SENT32, SENT32, SENT32);
        }
      }
      VB_STAMP(5);
      // (B) collect + normalise: rows cta, cta + n_ctas, ...
      const unsigned char* oc = c.local + oring_cur;
      for (int t = cta; t < p.M; t += n_ctas) {
        uint32_t w[MAXG][4];
#pragma unroll
        for (int gi = 0; gi < MAXG; ++gi) {
          const int g = et + gi * EPI_THREADS;
          if (g < groups) ld_relaxed_sys_v4(oc + ((size_t)t * p.N + g * 8) * sizeof(T), w[gi]);
        }
        float ss = 0.f;
#pragma unroll
        for (int gi = 0; gi < MAXG; ++gi) {
          const int g = et + gi * EPI_THREADS;
          if (g >= groups) break;
          const unsigned char* ptr = oc + ((size_t)t * p.N + g * 8) * sizeof(T);
          while (has_sentinel(w[gi])) { spin_guard(spins, t_begin, 2); ld_relaxed_sys_v4(ptr, w[gi]); }
#pragma unroll
          for (int e = 0; e < 4; ++e) { float2 f = unpack2<T>(w[gi][e]); ss += f.x * f.x + f.y * f.y; }
        }
        ss = warp_sum(ss);
        if (lane == 0) red[q] = ss;
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS));
        const float tot = red[0] + red[1] + red[2] + red[3];
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS));
        const float inv = rsqrtf(tot / (float)p.N + p.eps);
#pragma unroll
        for (int gi = 0; gi < MAXG; ++gi) {
          const int g = et + gi * EPI_THREADS;
          if (g >= groups) break;
          Vec16<T> wv = ld16(reinterpret_cast<const T*>(p.norm_w) + g * 8);
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = unpack2<T>(w[gi][e]);
            const float g0 = to_f32(wv.e[2 * e]) + (p.gemma ? 1.f : 0.f), g1 = to_f32(wv.e[2 * e + 1]) + (p.gemma ? 1.f : 0.f);
            o[e] = pack2<T>(f.x * inv * g0, f.y * inv * g1);
          }
          *reinterpret_cast<uint4*>(reinterpret_cast<T*>(p.out) + (size_t)t * p.N + g * 8) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    }
    VB_STAMP(6);
    // ---- last CTA out advances the call counter
    asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS));
    if (et == 0) {
      __threadfence();
      const unsigned prev = atomicAdd(&c.state[9], 1u);
      if (prev == (unsigned)(n_ctas - 1)) {
        c.state[9] = 0;
        c.state[10 + call % 3] = (uint32_t)p.M;
        c.state[8] = call + 1;
      }
    }
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
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || !ptr) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
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
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, const CommCtx& c, dim3 grid,
                  cudaStream_t st) {
  using L = SmemLayout<BN>;
  int smem = STAGES * L::STAGE_BYTES + (2 * STAGES + 1) * 8 + 64 + 1024;
  const int need_t = BN * TM * 2 + 1024;   // transposition buffer aliases the operand ring
  if (STAGES * L::STAGE_BYTES < need_t) return -6;
  auto k = gemm_ar_kernel<T, BN, STAGES>;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
  if (grid.z > 1) vb_launch_cluster(k, grid, dim3(NUM_THREADS), (size_t)smem, st, grid.z, ta, tb, p, c);
  else vb_launch(k, grid, dim3(NUM_THREADS), (size_t)smem, st, ta, tb, p, c);
  return (int)cudaGetLastError();
}

}  // namespace gemm_ar

static unsigned long long* g_trace = nullptr;
EXPORT void vb_gemm_ar_set_trace(void* p) { g_trace = reinterpret_cast<unsigned long long*>(p); }

// Plan: number of K splits so that tiles*splits <= num_sms (all CTAs must be co-resident: reducers spin on
// data produced by other CTAs / other GPUs).
extern "C" int vb_gemm_max_active_clusters(int c);

EXPORT int vb_gemm_ar_plan(int M, int N, int K, int num_sms, int* bn, int* splits, int64_t* ws_bytes, int* n_tiles) {
  using namespace gemm_ar;
  if (M > 128) return -1;
  const int kb = (K + BK - 1) / BK;
  const int nt = (N + TM - 1) / TM;
  if (nt > num_sms) return -2;
  int s = 1;                               // same policy as vb_gemm_tc_plan (cluster split-K)
  for (int c = 8; c > 1; --c) {
    if (c * nt > 132 || kb / c < 4) continue;
    const int fit = vb_gemm_max_active_clusters(c);
    if (fit > 0 && nt > fit) continue;
    s = c;
    break;
  }
  if (nt >= 56 && kb < 32) s = 1;
  const int b = ((M + 15) / 16) * 16;
  *bn = b; *splits = s; *n_tiles = nt;
  *ws_bytes = 0;                          // split-K partials live in (distributed) shared memory
  return 0;
}

EXPORT int vb_gemm_ar(void* local, void* const* peers, void* mc, uint32_t* state, int rank, int world, const void* x,
                      const void* w, void* residual, const void* norm_w, void* out, int M, int N, int K,
                      int64_t x_stride, int bn, int splits, float* ws, unsigned* counters, int64_t lam_off,
                      int64_t lam_stride, int t_cap, int mode, int64_t oring_off, int64_t oring_stride, float eps, int gemma, int dtype,
                      cudaStream_t st) {
  using namespace gemm_ar;
  if (dtype == DT_F32 || K % 8 != 0 || N % 8 != 0 || N / 8 > 8 * EPI_THREADS) return -1;
  const bool bf16 = dtype == DT_BF16;
  CommCtx c;
  c.local = (unsigned char*)local;
  for (int i = 0; i < VB_MAX_RANKS; ++i) c.peer[i] = i < world ? (unsigned char*)peers[i] : nullptr;
  c.mc = (unsigned char*)mc; c.state = state; c.rank = rank; c.world = world;
  CUtensorMap ta, tb;
  Params p;
  p.out = out; p.residual = residual; p.norm_w = norm_w; p.ws = ws; p.counters = counters;
  p.lam_off = (size_t)lam_off; p.lam_stride = (size_t)lam_stride; p.M = M; p.N = N; p.K = K; p.splits = splits;
  const int kb = (K + BK - 1) / BK;
  p.k_blocks_per_split = (kb + splits - 1) / splits;
  p.eps = eps; p.gemma = gemma; p.t_cap = t_cap; p.trace = g_trace;
  p.mode = mode; p.oring_off = (size_t)oring_off; p.oring_stride = (size_t)oring_stride;
  if (M > t_cap) return -7;
  int rc;
  if ((rc = make_tmap(&ta, w, N, K, K, TM, bf16))) return rc;
  if ((rc = make_tmap(&tb, x, M, K, x_stride, bn, bf16))) return rc;
  dim3 grid((N + TM - 1) / TM, 1, splits);
#define VB_GO(BNV, ST)                                                                  \
  if (bf16) return launch<__nv_bfloat16, BNV, ST>(ta, tb, p, c, grid, st);              \
  else return launch<__half, BNV, ST>(ta, tb, p, c, grid, st);
  switch (bn) {
    case 16: VB_GO(16, 8)
    case 32: VB_GO(32, 8)
    case 48: VB_GO(48, 8)
    case 64: VB_GO(64, 8)
    case 80: VB_GO(80, 6)
    case 96: VB_GO(96, 6)
    case 112: VB_GO(112, 6)
    case 128: VB_GO(128, 6)
    default: return -2;
  }
#undef VB_GO
}
