// CUDA virtual-memory-management arena for sleep mode (vllm_b200/device_allocator): every allocation is a
// reserved VA range backed by physical handles that can be unmapped + released while the VA stays reserved, so a
// sleeping engine gives its HBM back and wakes up with every tensor at the same address (CUDA graphs stay valid).
// Role of the reference's csrc/cumem_allocator.cpp; the driver API is resolved with dlopen so the library loads on
// machines without a GPU.
#include <cstdint>
#include <cstdio>
#include <dlfcn.h>
#include <mutex>
#include <unordered_map>
#include <cuda.h>

namespace {
struct Api {
  void* lib = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  bool ok = false;
};
Api& api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    a.lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!a.lib) return;
#define L(name, sym) a.name = reinterpret_cast<decltype(a.name)>(dlsym(a.lib, sym)); if (!a.name) return;
    L(MemAddressReserve, "cuMemAddressReserve") L(MemAddressFree, "cuMemAddressFree") L(MemCreate, "cuMemCreate")
    L(MemRelease, "cuMemRelease") L(MemMap, "cuMemMap") L(MemUnmap, "cuMemUnmap") L(MemSetAccess, "cuMemSetAccess")
    L(MemGetAllocationGranularity, "cuMemGetAllocationGranularity")
#undef L
    a.ok = true;
  });
  return a;
}
struct Block { size_t size; int device; CUmemGenericAllocationHandle handle; bool mapped; };
std::mutex g_mu;
std::unordered_map<uint64_t, Block> g_blocks;

CUmemAllocationProp prop_for(int device) {
  CUmemAllocationProp p = {};
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = device;
  return p;
}
int map_block(uint64_t ptr, Block& b) {
  Api& a = api();
  CUmemAllocationProp p = prop_for(b.device);
  if (a.MemCreate(&b.handle, b.size, &p, 0) != CUDA_SUCCESS) return -2;
  if (a.MemMap((CUdeviceptr)ptr, b.size, 0, b.handle, 0) != CUDA_SUCCESS) { a.MemRelease(b.handle); return -3; }
  CUmemAccessDesc d = {};
  d.location.type = CU_MEM_LOCATION_TYPE_DEVICE; d.location.id = b.device; d.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  if (a.MemSetAccess((CUdeviceptr)ptr, b.size, &d, 1) != CUDA_SUCCESS) return -4;
  b.mapped = true;
  return 0;
}
}  // namespace

extern "C" {

int vbn_cumem_available() { return api().ok ? 1 : 0; }

// Rounds `size` up to the allocation granularity; returns the device pointer (0 on failure).
uint64_t vbn_cumem_alloc(int device, uint64_t size, uint64_t* rounded) {
  Api& a = api();
  if (!a.ok) return 0;
  CUmemAllocationProp p = prop_for(device);
  size_t gran = 0;
  if (a.MemGetAllocationGranularity(&gran, &p, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0) return 0;
  const size_t sz = (size + gran - 1) / gran * gran;
  CUdeviceptr va = 0;
  if (a.MemAddressReserve(&va, sz, 0, 0, 0) != CUDA_SUCCESS) return 0;
  Block b{sz, device, 0, false};
  if (map_block((uint64_t)va, b) != 0) { a.MemAddressFree(va, sz); return 0; }
  std::lock_guard<std::mutex> g(g_mu);
  g_blocks[(uint64_t)va] = b;
  if (rounded) *rounded = sz;
  return (uint64_t)va;
}

// Give the physical memory back, keep the address range (sleep).
int vbn_cumem_unmap(uint64_t ptr) {
  Api& a = api();
  std::lock_guard<std::mutex> g(g_mu);
  auto it = g_blocks.find(ptr);
  if (it == g_blocks.end()) return -1;
  if (!it->second.mapped) return 0;
  a.MemUnmap((CUdeviceptr)ptr, it->second.size);
  a.MemRelease(it->second.handle);
  it->second.mapped = false;
  return 0;
}

// Back the same address range with fresh physical memory (wake up).
int vbn_cumem_remap(uint64_t ptr) {
  std::lock_guard<std::mutex> g(g_mu);
  auto it = g_blocks.find(ptr);
  if (it == g_blocks.end()) return -1;
  if (it->second.mapped) return 0;
  return map_block(ptr, it->second);
}

int vbn_cumem_free(uint64_t ptr) {
  Api& a = api();
  std::lock_guard<std::mutex> g(g_mu);
  auto it = g_blocks.find(ptr);
  if (it == g_blocks.end()) return -1;
  if (it->second.mapped) { a.MemUnmap((CUdeviceptr)ptr, it->second.size); a.MemRelease(it->second.handle); }
  a.MemAddressFree((CUdeviceptr)ptr, it->second.size);
  g_blocks.erase(it);
  return 0;
}

uint64_t vbn_cumem_mapped_bytes() {
  std::lock_guard<std::mutex> g(g_mu);
  uint64_t n = 0;
  for (auto& kv : g_blocks) if (kv.second.mapped) n += kv.second.size;
  return n;
}

}  // extern "C"

// torch.cuda.memory.CUDAPluggableAllocator entry points: allocations made inside
// ``CuMemAllocator.use_memory_pool`` come from the VMM arena above.
extern "C" {
void* vbn_torch_malloc(ssize_t size, int device, void* /*stream*/) {
  uint64_t rounded = 0;
  return reinterpret_cast<void*>(vbn_cumem_alloc(device, (uint64_t)size, &rounded));
}
void vbn_torch_free(void* ptr, ssize_t /*size*/, int /*device*/, void* /*stream*/) {
  vbn_cumem_free(reinterpret_cast<uint64_t>(ptr));
}
// enumerate live blocks: fills up to `cap` (ptr, size, mapped) triples, returns the count
int vbn_cumem_list(uint64_t* out, int cap) {
  std::lock_guard<std::mutex> g(g_mu);
  int n = 0;
  for (auto& kv : g_blocks) {
    if (n < cap) { out[3 * n] = kv.first; out[3 * n + 1] = kv.second.size; out[3 * n + 2] = kv.second.mapped ? 1 : 0; }
    ++n;
  }
  return n;
}
}
