// Shared device helpers for the sm_100a kernels: dtype traits, vector access,
// warp/block reductions and thin PTX wrappers (mbarrier, 1-D bulk copy, TMA
// tensor copy, tcgen05 / TMEM).  No CUTLASS dependency.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstdlib>

#define VB_HOST_DEVICE __host__ __device__ __forceinline__
#define VB_DEVICE __device__ __forceinline__
#define EXPORT extern "C" __attribute__((visibility("default")))

// dtype codes shared with ops/cuda.py
enum : int { DT_BF16 = 0, DT_F16 = 1, DT_F32 = 2 };

template <typename T> struct DTypeOf;
template <> struct DTypeOf<__nv_bfloat16> { static constexpr int v = DT_BF16; };
template <> struct DTypeOf<__half> { static constexpr int v = DT_F16; };
template <> struct DTypeOf<float> { static constexpr int v = DT_F32; };

#define DISPATCH_DTYPE(dt, T, ...)                                   \
  do {                                                               \
    if ((dt) == DT_BF16) { using T = __nv_bfloat16; __VA_ARGS__; }   \
    else if ((dt) == DT_F16) { using T = __half; __VA_ARGS__; }      \
    else { using T = float; __VA_ARGS__; }                           \
  } while (0)

#define DISPATCH_HALF(dt, T, ...)                                    \
  do {                                                               \
    if ((dt) == DT_BF16) { using T = __nv_bfloat16; __VA_ARGS__; }   \
    else { using T = __half; __VA_ARGS__; }                          \
  } while (0)

VB_DEVICE float to_f32(float x) { return x; }
VB_DEVICE float to_f32(__half x) { return __half2float(x); }
VB_DEVICE float to_f32(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> VB_DEVICE T from_f32(float x);
template <> VB_DEVICE float from_f32<float>(float x) { return x; }
template <> VB_DEVICE __half from_f32<__half>(float x) { return __float2half_rn(x); }
template <> VB_DEVICE __nv_bfloat16 from_f32<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

// 16-byte vector of T
template <typename T> struct Vec16 {
  static constexpr int N = 16 / sizeof(T);
  union { uint4 u; T e[N]; };
};
template <typename T> VB_DEVICE Vec16<T> ld16(const T* p) {
  Vec16<T> v; v.u = *reinterpret_cast<const uint4*>(p); return v;
}
template <typename T> VB_DEVICE void st16(T* p, const Vec16<T>& v) { *reinterpret_cast<uint4*>(p) = v.u; }

// unpack a packed pair of 16-bit floats held in a u32
template <typename T> VB_DEVICE float2 unpack2(uint32_t x);
template <> VB_DEVICE float2 unpack2<__nv_bfloat16>(uint32_t x) {
  return make_float2(__uint_as_float(x << 16), __uint_as_float(x & 0xffff0000u));
}
template <> VB_DEVICE float2 unpack2<__half>(uint32_t x) {
  return __half22float2(*reinterpret_cast<__half2*>(&x));
}
template <typename T> VB_DEVICE uint32_t pack2(float a, float b);
template <> VB_DEVICE uint32_t pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&h);
}
template <> VB_DEVICE uint32_t pack2<__half>(float a, float b) {
  __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h);
}

VB_DEVICE uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
VB_DEVICE void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

VB_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
VB_DEVICE float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum; `red` is >= 32 floats of shared memory
VB_DEVICE float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  __syncthreads();
  return r;
}

// ------------------------------------------------------------------ PTX: misc
VB_DEVICE uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
VB_DEVICE bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
VB_DEVICE float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// -------------------------------------------------------------- PTX: mbarrier
VB_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
VB_DEVICE void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
VB_DEVICE void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
VB_DEVICE void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
VB_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
VB_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}

// ------------------------------------------------- PTX: bulk / tensor copies
// 1-D bulk copy global → shared, completion counted on an mbarrier (bytes % 16 == 0)
VB_DEVICE void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// 2-D TMA tile load (tensor map in param/global memory)
VB_DEVICE void tma_load_2d(void* dst_smem, const void* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
VB_DEVICE void tma_load_2d_hint(void* dst_smem, const void* tmap, int c0, int c1, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
               " [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
// HBM → L2 only (no shared-memory slot needed): lets a CTA keep far more bytes in flight than its smem ring holds
VB_DEVICE void tma_prefetch_l2_2d(const void* tmap, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tmap), "r"(c0), "r"(c1) : "memory");
}
VB_DEVICE void tma_store_2d(const void* tmap, const void* src_smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tmap), "r"(smem_u32(src_smem)), "r"(c0), "r"(c1) : "memory");
}
VB_DEVICE void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
VB_DEVICE void tma_store_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
VB_DEVICE void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
VB_DEVICE void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// L2 eviction policies for TMA cache hints
constexpr uint64_t L2_EVICT_FIRST = 0x12F0000000000000ull;
constexpr uint64_t L2_EVICT_LAST = 0x14F0000000000000ull;
constexpr uint64_t L2_EVICT_NORMAL = 0x1000000000000000ull;

// ------------------------------------------------------------ PTX: tcgen05
VB_DEVICE void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp, ncols pow2 >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
VB_DEVICE void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
VB_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
VB_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16/fp16 inputs, fp32 accumulate
VB_DEVICE void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
VB_DEVICE void umma_f8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
VB_DEVICE void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns → 32 registers per thread (warp w reads lanes 32*(w%4)..+31)
VB_DEVICE void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
VB_DEVICE void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
VB_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor: K-major operand tile written by TMA with 128-byte swizzle.
// Rows are 128 B (64 x 16-bit); 8-row groups are 1024 B apart (SBO); LBO is unused for
// swizzled K-major layouts; bits 46-47 = descriptor version 1 (sm_100); bits 61-63 = 2 (SWIZZLE_128B).
VB_DEVICE uint64_t umma_desc_sw128(const void* smem_tile) {
  const uint32_t addr = smem_u32(smem_tile);
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);        // start address
  d |= (uint64_t)1 << 16;                         // LBO (ignored)
  d |= (uint64_t)(1024 >> 4) << 32;               // SBO = 1024 B
  d |= (uint64_t)1 << 46;                         // version
  d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16 (bf16 or fp16 inputs, fp32 accumulator, both operands K-major)
VB_HOST_DEVICE uint32_t umma_idesc_f16(int M, int N, bool bf16) {
  uint32_t d = 0;
  d |= 1u << 4;                        // D format: f32
  d |= (bf16 ? 1u : 0u) << 7;          // A format
  d |= (bf16 ? 1u : 0u) << 10;         // B format
  d |= (uint32_t)(N >> 3) << 17;       // N / 8
  d |= (uint32_t)(M >> 4) << 24;       // M / 16
  return d;
}
// kind::f8f6f4 with e4m3 inputs
VB_HOST_DEVICE uint32_t umma_idesc_f8(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

// ------------------------------------------- multimem (NVLS) + system-scope flags
VB_DEVICE void multimem_ld_reduce_bf16x8(const void* mc_ptr, uint32_t (&r)[4]) {
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "l"(mc_ptr) : "memory");
}
VB_DEVICE void multimem_st_b32x4(void* mc_ptr, const uint32_t (&r)[4]) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               ::"l"(mc_ptr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
VB_DEVICE void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
VB_DEVICE uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
VB_DEVICE void red_add_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
VB_DEVICE uint32_t ld_volatile_u32(const uint32_t* p) {
  uint32_t v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}

// ------------------------------------------------ thread-block clusters / distributed shared memory
VB_DEVICE void cluster_sync() {   // every thread of every CTA in the cluster; warp-convergent
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
VB_DEVICE uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
// shared::cluster address of `smem_addr` (a shared::cta address of THIS CTA) in CTA `rank` of the cluster
VB_DEVICE uint32_t dsmem_addr(uint32_t smem_addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank)); return r;
}
VB_DEVICE float4 ld_dsmem_f32x4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// ------------------------------------------------ programmatic dependent launch (PDL)
// Every kernel calls pdl_trigger() first (lets the next kernel in the stream begin launching: its prologue,
// barrier init, TMEM alloc and — for GEMMs — the weight TMA stream overlap this kernel's tail) and
// pdl_wait() before the first access to memory a predecessor may still be reading or writing.
VB_DEVICE void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
VB_DEVICE void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool vb_pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = std::getenv("VLLM_B200_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t vb_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                             Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = vb_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Same, with the grid's z extent grouped into thread-block clusters of `cluster_z` CTAs (co-scheduled on one GPC,
// distributed shared memory between them).
template <typename... KArgs, typename... Args>
inline cudaError_t vb_launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     unsigned cluster_z, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = cluster_z;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = vb_pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

#define CUDA_CHECK_LAST() cudaGetLastError()
