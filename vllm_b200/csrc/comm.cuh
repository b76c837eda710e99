// Cross-GPU primitives over the symmetric heap (NVSwitch P2P + NVLS multicast).
//
// Every rank maps the same allocation of every peer (`peer[r]`) plus one multicast alias (`mc`) that
// addresses all replicas at once: `multimem.ld_reduce` sums the replicas inside the switch,
// `multimem.st` writes to all of them.  Synchronisation uses monotonically increasing epochs kept in
// device memory, so the same kernels replay inside CUDA graphs with no host-side flag reset.
//
// Reference counterparts: csrc/custom_all_reduce.cuh (P2P flags), custom_all_gather_reduce_scatter.cuh,
// kimi_k3 gemm_rs_ar.py (NVLS reduce inside a GEMM), FlashInfer allreduce_fusion.
#pragma once
#include "common.cuh"

constexpr int VB_MAX_RANKS = 8;

// layout of the symmetric buffer (identical on every rank)
constexpr size_t VB_FLAGS_BYTES = 4096;   // [0,4096): uint32 flags[VB_MAX_RANKS] (+ padding)

struct CommCtx {
  unsigned char* local;                 // this rank's mapping of the symmetric buffer
  unsigned char* peer[VB_MAX_RANKS];    // every rank's buffer as seen from here (peer[rank] == local)
  unsigned char* mc;                    // multicast alias (nullptr if NVLS is unavailable)
  uint32_t* state;                      // local, non-symmetric: [0]=epoch [1]=arrive0 [2]=go [3]=arrive1
  int rank, world;
};


// Bounded spin: a barrier whose peer never arrives (dead rank, SMs held elsewhere) traps after ~20 s instead of hanging
// the GPU until the driver's watchdog (the reference's DCP kernels carry the same guard, dcp_direct_common.cuh:15).
VB_DEVICE unsigned long long vb_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
struct SpinGuard {
  unsigned long long t0;
  uint32_t n;
  VB_DEVICE SpinGuard() : t0(0), n(0) {}
  VB_DEVICE void tick(int what) {
    if ((++n & 0xfffffu) != 0) return;
    const unsigned long long now = vb_gtime();
    if (t0 == 0) { t0 = now; return; }
    if (now - t0 > 20ull * 1000 * 1000 * 1000) {
      printf("vllm_b200 comm: spin limit exceeded (site %d, block %d)\n", what, blockIdx.x);
      __trap();
    }
  }
};

// All ranks rendezvous: executed by (at least) `world` threads of ONE block per rank.
// Rank r publishes `target` into slot r of every peer's flag array, then waits for all slots of its own.
VB_DEVICE void rank_barrier(const CommCtx& c, uint32_t target) {
  const int t = threadIdx.x;
  if (t < c.world) {
    uint32_t* dst = reinterpret_cast<uint32_t*>(c.peer[t]) + c.rank;
    st_release_sys(dst, target);
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(c.local) + t;
    SpinGuard sg;
    while ((int32_t)(ld_acquire_sys(mine) - target) < 0) sg.tick(0);
  }
}

// Grid-wide "everyone arrived → one block talks to the peers → everyone continues".
// counter/go live in c.state; `target` must be unique per use (epoch + k).
// If `set_epoch` is true the leader also publishes `target` as the new epoch (state[0]) — safe because every
// block read the old epoch before it could arrive here.
VB_DEVICE void grid_rank_barrier(const CommCtx& c, uint32_t* counter, uint32_t* go, uint32_t target, bool wait_all,
                                 bool set_epoch = false) {
  __threadfence_system();
  __syncthreads();
  __shared__ int leader;
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(counter, 1u);
    leader = prev == gridDim.x - 1;
    if (leader) *counter = 0;
  }
  __syncthreads();
  if (leader) {
    rank_barrier(c, target);
    __syncthreads();
    if (threadIdx.x == 0) {
      if (set_epoch) c.state[0] = target;
      st_release_gpu(go, target);
    }
  } else if (wait_all) {
    if (threadIdx.x == 0) {
      SpinGuard sg;
      while ((int32_t)(ld_acquire_gpu(go) - target) < 0) sg.tick(1);
    }
  }
  if (wait_all) __syncthreads();
}

VB_DEVICE void multimem_ld_reduce_f16x8(const void* mc_ptr, uint32_t (&r)[4]) {
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "l"(mc_ptr) : "memory");
}
template <typename T> VB_DEVICE void mc_ld_reduce8(const void* mc_ptr, uint32_t (&r)[4]);
template <> VB_DEVICE void mc_ld_reduce8<__nv_bfloat16>(const void* p, uint32_t (&r)[4]) { multimem_ld_reduce_bf16x8(p, r); }
template <> VB_DEVICE void mc_ld_reduce8<__half>(const void* p, uint32_t (&r)[4]) { multimem_ld_reduce_f16x8(p, r); }

// sum 8 x 16-bit values held by every rank at byte offset `off` (fp32 accumulation)
template <typename T>
VB_DEVICE void reduce8(const CommCtx& c, size_t off, float (&acc)[8]) {
  if (c.mc) {
    uint32_t r[4];
    mc_ld_reduce8<T>(c.mc + off, r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = unpack2<T>(r[i]); acc[2 * i] = f.x; acc[2 * i + 1] = f.y; }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int p = 0; p < c.world; ++p) {
      uint32_t w[4];
      asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3])
                   : "l"(c.peer[(c.rank + p) % c.world] + off) : "memory");
#pragma unroll
      for (int i = 0; i < 4; ++i) { float2 f = unpack2<T>(w[i]); acc[2 * i] += f.x; acc[2 * i + 1] += f.y; }
    }
  }
}

// write 16 bytes to the same offset of every rank's buffer
VB_DEVICE void bcast16(const CommCtx& c, size_t off, const uint32_t (&v)[4]) {
  if (c.mc) {
    multimem_st_b32x4(c.mc + off, v);
  } else {
    const uint4 u = make_uint4(v[0], v[1], v[2], v[3]);
    for (int p = 0; p < c.world; ++p) *reinterpret_cast<uint4*>(c.peer[(c.rank + p) % c.world] + off) = u;
  }
}
