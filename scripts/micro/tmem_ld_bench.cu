// TMEM read throughput of 16 epilogue warps sweeping a 128 x 416 fp32 accumulator with tcgen05.ld.32x32b.x8 / x16 / x32
// (interleaved chunks per column quarter, as the GEMM epilogue reads it).  Prints cycles per full-tile pass.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../rl_sandbox_b200/csrc/rlsb_ptx.cuh"
using namespace rlsb;

template <int W>
__device__ __forceinline__ uint32_t ld_chunk(uint32_t addr) {
  uint32_t x = 0;
  if (W == 8) { uint32_t r[8]; tmem_ld8(addr, r); tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) x ^= r[i]; }
  if (W == 16) { uint32_t r[16]; tmem_ld16(addr, r); tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) x ^= r[i]; }
  if (W == 32) { uint32_t r[32]; tmem_ld32(addr, r); tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) x ^= r[i]; }
  return x;
}

template <int W, int WARPS>
__global__ void __launch_bounds__(WARPS * 32 + 32) bench(int reps, int cols, long long* cyc, uint32_t* sink) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (warp == WARPS) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < WARPS) {
    const int q = warp & 3, cq = warp >> 2, ncq = WARPS / 4;
    const uint32_t base = tbase + (static_cast<uint32_t>(q * 32) << 16);
    asm volatile("bar.sync 1, %0;" ::"n"(WARPS * 32));
    t0 = clock64();
    for (int r = 0; r < reps; ++r)
      for (int c = cq * W; c + W <= cols; c += ncq * W) acc ^= ld_chunk<W>(base + c);
    asm volatile("bar.sync 1, %0;" ::"n"(WARPS * 32));
    t1 = clock64();
  }
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == WARPS) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

template <int W, int WARPS>
void run(const char* name, long long* cyc, uint32_t* sink) {
  const int reps = 200, cols = 416;
  bench<W, WARPS><<<148, WARPS * 32 + 32>>>(reps, cols, cyc, sink);
  cudaDeviceSynchronize();
  bench<W, WARPS><<<148, WARPS * 32 + 32>>>(reps, cols, cyc, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < 148; ++i) s += h[i];
  printf("%-28s %8.0f cycles per 128x%d pass  (%.1f B/clk/SM)  %s\n", name, s / 148 / reps, cols,
         128.0 * cols * 4 / (s / 148 / reps), cudaGetErrorString(e));
}

int main() {
  long long* cyc; uint32_t* sink;
  cudaMalloc(&cyc, 148 * sizeof(long long)); cudaMalloc(&sink, 4);
  run<8, 16>("x8, 16 warps (ld+wait)", cyc, sink);
  run<16, 16>("x16, 16 warps (ld+wait)", cyc, sink);
  run<32, 16>("x32, 16 warps (ld+wait)", cyc, sink);
  run<8, 4>("x8, 4 warps (ld+wait)", cyc, sink);
  run<16, 4>("x16, 4 warps (ld+wait)", cyc, sink);
  run<32, 4>("x32, 4 warps (ld+wait)", cyc, sink);
  run<8, 8>("x8, 8 warps (ld+wait)", cyc, sink);
  run<32, 8>("x32, 8 warps (ld+wait)", cyc, sink);
  return 0;
}
