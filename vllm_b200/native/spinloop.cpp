// Spin-wait primitives for the shared-memory control plane (one writer → N readers ring, see
// vllm_b200/parallel/shm_broadcast.py).  Called through ctypes, which drops the GIL for the duration of the
// call, so a worker blocked here costs the interpreter nothing.
//
// Role of the reference's csrc/spinloop.cpp (GIL-releasing flag wait with monitorx/mwaitx or a pause loop);
// this version waits on 64-bit sequence numbers instead of byte flags, escalates pause → yield → futex-less
// nanosleep, and reports a timeout so the Python side can check for a dead peer.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <thread>
#include <time.h>
#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
static inline void cpu_relax() { _mm_pause(); }
#elif defined(__aarch64__)
static inline void cpu_relax() { asm volatile("yield" ::: "memory"); }
#else
static inline void cpu_relax() {}
#endif

extern "C" {

// Wait until *addr >= target (acquire).  Returns 1 on success, 0 on timeout (timeout_us < 0: wait forever).
int vbn_wait_ge_u64(const volatile uint64_t* addr, uint64_t target, int64_t timeout_us) {
  const auto* a = reinterpret_cast<const std::atomic<uint64_t>*>(const_cast<const uint64_t*>(addr));
  const auto t0 = std::chrono::steady_clock::now();
  uint32_t spins = 0;
  for (;;) {
    if (a->load(std::memory_order_acquire) >= target) return 1;
    ++spins;
    if (spins < 4096) {
      cpu_relax();
    } else if (spins < 4096 + 2048) {
      std::this_thread::yield();
    } else {
      // long idle (engine waiting for requests): back off so an idle server does not burn a core per worker
      struct timespec ts = {0, 50 * 1000};
      nanosleep(&ts, nullptr);
      if (timeout_us >= 0) {
        const auto el = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0);
        if (el.count() > timeout_us) return 0;
      }
    }
    if (timeout_us >= 0 && (spins & 1023) == 0) {
      const auto el = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0);
      if (el.count() > timeout_us) return 0;
    }
  }
}

// Wait until every one of n counters is >= target (the writer waiting for all readers to release a slot).
int vbn_wait_all_ge_u64(const volatile uint64_t* addr, int n, int64_t stride_bytes, uint64_t target, int64_t timeout_us) {
  for (int i = 0; i < n; ++i) {
    const auto* p = reinterpret_cast<const volatile uint64_t*>(reinterpret_cast<const volatile char*>(addr) + i * stride_bytes);
    if (!vbn_wait_ge_u64(p, target, timeout_us)) return 0;
  }
  return 1;
}

void vbn_store_release_u64(volatile uint64_t* addr, uint64_t v) {
  reinterpret_cast<std::atomic<uint64_t>*>(const_cast<uint64_t*>(addr))->store(v, std::memory_order_release);
}

uint64_t vbn_load_acquire_u64(const volatile uint64_t* addr) {
  return reinterpret_cast<const std::atomic<uint64_t>*>(const_cast<const uint64_t*>(addr))->load(std::memory_order_acquire);
}

}  // extern "C"
