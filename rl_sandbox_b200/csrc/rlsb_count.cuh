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
}  // namespace rlsb
#endif
