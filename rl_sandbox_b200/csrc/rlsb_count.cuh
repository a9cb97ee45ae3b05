// rlsb_count.cuh — counts kernel launches issued by this library (bench.py's `gpu_launches`).
#pragma once
#include <atomic>
namespace rlsb {
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Programmatic dependent launch: a kernel launched through launch_pdl may be scheduled while its predecessor in the
// stream is still running; it executes pdl_wait() (griddepcontrol.wait) before it touches global memory, so what
// overlaps is the launch latency, block scheduling and the kernel's own prologue (barrier init, TMEM allocation) —
// the rollout at 800 start states is a chain of ~300 such launches of a few microseconds each.  RLSB_PDL=0 disables.
extern int g_pdl;
}  // namespace rlsb

#ifdef __CUDACC__
#include <cuda_runtime.h>
namespace rlsb {
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_pdl;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// "done once per device" flag for per-device function attributes (cudaFuncSetAttribute applies to the current device
// only: one process driving several GPUs must set it on each).  Setting an attribute twice is harmless, so a race between
// two first callers is benign; the flag itself is atomic.
struct PerDeviceOnce {
  std::atomic<unsigned long long> mask{0};
  // true if the caller still has to initialise the current device; `bit` receives the device's flag for done()
  bool need(unsigned long long& bit) const {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    bit = 1ull << (dev & 63);
    return (mask.load(std::memory_order_acquire) & bit) == 0;
  }
  void done(unsigned long long bit) { mask.fetch_or(bit, std::memory_order_release); }
};
}  // namespace rlsb
#endif
