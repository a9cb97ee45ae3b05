// rlsb_rowops.cuh — device code shared by the stand-alone row kernels (rlsb_kernels.cu) and the persistent rollout
// kernel (rlsb_rollout.cu): MUFU-based sigmoid / tanh, the Philox noise readers and the 32 x 32 latent draw
// (reference: rssm.py:34-37, dists.py:177-179, common.py:76-80).  Everything here is `__device__ __forceinline__` or a
// template, so each translation unit gets its own copy.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "rlsb_detmath.h"
#include "rlsb_gemm.cuh"
#include "rlsb_kernels.cuh"
#include "rlsb_ptx.cuh"

namespace rlsb {
namespace rowops {

__device__ __forceinline__ uint32_t bf2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// sigmoid / tanh from MUFU.EX2 + MUFU.RCP (2 ulp each; no IEEE division sequence): |error| < 3e-7 absolute
__device__ __forceinline__ float fast_sigmoid(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float fast_tanh(float x) {
  // 1 - 2 / (1 + e^{2x}); e^{2x} -> inf gives 1, -> 0 gives -1
  return fmaf(-2.0f, rcp_approx(1.0f + ex2_approx(2.8853900817779268f * x)), 1.0f);
}

// ------------------------------------------------------------------------------------------
// noise
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float noise_uniform(const NoiseSpec& ns, int m, uint32_t stream,
                                               uint32_t e) {
  if (ns.explicit_noise) return __ldg(ns.explicit_noise + static_cast<size_t>(m) * ns.ld + e);
  return rlsb_noise_uniform(ns.seed_ptr ? __ldg(ns.seed_ptr) : ns.seed, ns.row_offset + static_cast<uint32_t>(m), ns.step, stream, e);
}

// Box-Muller on two Philox uniforms (device-only path; parity tests pass explicit normals)
__device__ __forceinline__ float noise_normal(const NoiseSpec& ns, int m, uint32_t stream, uint32_t e) {
  if (ns.explicit_noise) return __ldg(ns.explicit_noise + static_cast<size_t>(m) * ns.ld + e);
  const uint64_t key = ns.seed_ptr ? __ldg(ns.seed_ptr) : ns.seed;
  const float u1 = rlsb_noise_uniform(key, ns.row_offset + static_cast<uint32_t>(m), ns.step, stream, 2 * e);
  const float u2 = rlsb_noise_uniform(key, ns.row_offset + static_cast<uint32_t>(m), ns.step, stream, 2 * e + 1);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// ------------------------------------------------------------------------------------------
// latent sampling: one thread per (row, group); classes == 32
// ------------------------------------------------------------------------------------------
struct SampleLatentArgs {
  const float* logits;
  long long ld;
  int M, groups;
  NoiseSpec noise;
  uint8_t* idx_out;
  __nv_bfloat16* onehot_packed;
  int kpad;
  float* onehot_f32;
  long long ld_f32;
};

// Fast standard Gumbel for the screening pass of sample_latent_kernel.  |result - rlsb_gumbel(u)| < 2e-5 for every
// representable u: -log(u) by MUFU.LG2 (abs. error 2^-21.4 on [0.5, 1)) where it is >= 0.03, by its series in
// d = 1 - u (exact subtraction; truncation < 2e-10 relative) closer to 1; the outer log is MUFU.LG2 again
// (arguments are never subnormal: u >= 1e-20, -log(u) >= 5.9e-8).
__device__ __forceinline__ float fast_gumbel(float u) {
  u = rlsb_clamp_uniform(u);
  const float d = 1.0f - u;
  float t = -0.6931471805599453f * lg2_approx(u);
  const float ser = d * fmaf(d, fmaf(d, fmaf(d, fmaf(d, fmaf(d, 0.16666667f, 0.2f), 0.25f), 0.33333334f), 0.5f), 1.0f);
  t = u > 0.96875f ? ser : t;
  return -0.6931471805599453f * lg2_approx(t);
}

// order-preserving float -> uint32 key whose low 5 bits carry (31 - class): one unsigned max then finds the largest
// score AND, among scores equal in the upper 27 bits, the lowest class.  Dropping 5 mantissa bits moves a score by at
// most 32 ulp (3.8e-6 relative), which the screening margin covers.
__device__ __forceinline__ uint32_t score_key(float s, int k) {
  const uint32_t b = __float_as_uint(s);
  const uint32_t o = b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
  return (o & ~31u) | static_cast<uint32_t>(31 - k);
}
__device__ __forceinline__ float key_score(uint32_t key) {
  const uint32_t o = key & ~31u;
  return __uint_as_float(o ^ ((o >> 31) ? 0x80000000u : 0xffffffffu));
}

// the reference-order draw of one (row, group): argmax_k fl(logit_k + G(u_k)), lowest index on ties; G is the
// bit-reproducible Gumbel transform the C oracle restates (rlsb_detmath.h)
template <bool COHERENT>
__device__ __noinline__ int sample_group_exact(const SampleLatentArgs& a, int m, int g, uint64_t key) {
  const float4* lp = reinterpret_cast<const float4*>(a.logits + static_cast<size_t>(m) * a.ld + g * 32);
  float best = 0.f;
  int best_k = 0;
  for (int q = 0; q < 8; ++q) {
    const float4 l4 = COHERENT ? __ldcg(lp + q) : __ldg(lp + q);
    float u[4];
    if (a.noise.explicit_noise) {
      const float4 u4 = __ldg(reinterpret_cast<const float4*>(a.noise.explicit_noise +
                                                              static_cast<size_t>(m) * a.noise.ld + g * 32) + q);
      u[0] = u4.x; u[1] = u4.y; u[2] = u4.z; u[3] = u4.w;
    } else {
      uint32_t o[4];
      rlsb_philox4x32(a.noise.row_offset + static_cast<uint32_t>(m), a.noise.step, 0u,
                      static_cast<uint32_t>(g * 8 + q), static_cast<uint32_t>(key),
                      static_cast<uint32_t>(key >> 32), o);
#pragma unroll
      for (int t = 0; t < 4; ++t) u[t] = rlsb_u32_to_uniform(o[t]);
    }
    const float l[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float s = __fadd_rn(l[t], rlsb_gumbel(u[t]));
      const int k = q * 4 + t;
      if (k == 0 || s > best) {  // strict '>' keeps the lowest index on ties (== argmax)
        best = s;
        best_k = k;
      }
    }
  }
  return best_k;
}

// Latent sampling, classes == 32.  Eight lanes share one (row, group): a lane holds four classes (one float4 of
// logits, one Philox block of uniforms), so a warp instruction reads 512 contiguous bytes.  Screening pass: the
// Gumbel-max with fast_gumbel and a top-2 butterfly over the eight lanes; the winner is final when it leads the
// runner-up by more than the worst-case difference between the fast and the bit-reproducible scores (2 x 2.5e-5 +
// one rounding of the sum).  Otherwise (about 1 group in 10^4, and whenever a score is NaN / infinite) the group is
// redrawn in the reference order with the bit-reproducible transform — the indices are those of the oracle, always.
// The draw for the (row, group) items [item0, item1) of `a` (item = row * groups + group; item0 a multiple of 4): warp
// `warp0` of `nwarps` takes every nwarps-th batch of four items.  COHERENT: the logits were written earlier in the SAME
// kernel launch by other CTAs (the persistent rollout, rlsb_rollout.cu) — read them through L2 (ld.global.cg), never
// through the non-coherent path.
template <bool COHERENT>
__device__ __forceinline__ float4 sample_ld4(const float4* p) {
  return COHERENT ? __ldcg(p) : __ldg(p);
}
template <bool COHERENT>
__device__ __forceinline__ void sample_latent_items(const SampleLatentArgs& a, long long item0, long long item1,
                                                    long long warp0, long long nwarps) {
  const long long total = item1;
  const int lane = threadIdx.x & 31;
  const int q = lane & 7;
  const uint64_t key = a.noise.seed_ptr ? __ldg(a.noise.seed_ptr) : a.noise.seed;
  const int gshift = (a.groups & (a.groups - 1)) == 0 ? __ffs(a.groups) - 1 : -1;
  const long long iters = (total - item0 + 3) >> 2;
  for (long long it = warp0; it < iters; it += nwarps) {
    long long item = item0 + it * 4 + (lane >> 3);
    const bool live = item < total;
    if (!live) item = total - 1;
    int m, g;
    if (gshift >= 0) {
      m = static_cast<int>(item >> gshift);
      g = static_cast<int>(item) & (a.groups - 1);
    } else {
      m = static_cast<int>(item / a.groups);
      g = static_cast<int>(item - static_cast<long long>(m) * a.groups);
    }
    const float4 l4 = sample_ld4<COHERENT>(reinterpret_cast<const float4*>(a.logits + static_cast<size_t>(m) * a.ld + g * 32) + q);
    float u[4];
    if (a.noise.explicit_noise) {
      const float4 u4 = __ldg(reinterpret_cast<const float4*>(a.noise.explicit_noise +
                                                              static_cast<size_t>(m) * a.noise.ld + g * 32) + q);
      u[0] = u4.x; u[1] = u4.y; u[2] = u4.z; u[3] = u4.w;
    } else {
      uint32_t o[4];
      rlsb_philox4x32(a.noise.row_offset + static_cast<uint32_t>(m), a.noise.step, 0u,
                      static_cast<uint32_t>(g * 8 + q), static_cast<uint32_t>(key),
                      static_cast<uint32_t>(key >> 32), o);
#pragma unroll
      for (int t = 0; t < 4; ++t) u[t] = rlsb_u32_to_uniform(o[t]);
    }
    const float sc[4] = {l4.x + fast_gumbel(u[0]), l4.y + fast_gumbel(u[1]), l4.z + fast_gumbel(u[2]),
                         l4.w + fast_gumbel(u[3])};
    // a NaN / infinite score anywhere in the group: let the reference-order draw decide
    const bool odd = !(fabsf((sc[0] + sc[1]) + (sc[2] + sc[3])) < 3.0e38f);
    const uint32_t k0 = score_key(sc[0], q * 4), k1 = score_key(sc[1], q * 4 + 1), k2 = score_key(sc[2], q * 4 + 2),
                   k3 = score_key(sc[3], q * 4 + 3);
    const uint32_t hi01 = max(k0, k1), lo01 = min(k0, k1), hi23 = max(k2, k3), lo23 = min(k2, k3);
    uint32_t b1 = max(hi01, hi23);                                  // largest key of the lane
    uint32_t b2 = max(min(hi01, hi23), hi01 > hi23 ? lo01 : lo23);  // second largest
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const uint32_t ob1 = __shfl_xor_sync(0xffffffffu, b1, o);
      const uint32_t ob2 = __shfl_xor_sync(0xffffffffu, b2, o);
      b2 = max(min(b1, ob1), max(b2, ob2));
      b1 = max(b1, ob1);
    }
    const unsigned oddmask = __ballot_sync(0xffffffffu, odd);
    const bool group_odd = ((oddmask >> (lane & 24)) & 0xffu) != 0u;
    const float f1 = key_score(b1), f2 = key_score(b2);
    const float margin = 1.0e-4f + 8.0e-6f * fabsf(f1);
    const bool unsure = group_odd || !(f1 - f2 > margin);
    const int i1 = 31 - static_cast<int>(b1 & 31u);
    int best_k = i1;
    if (unsure && q == 0) best_k = sample_group_exact<COHERENT>(a, m, g, key);
    best_k = __shfl_sync(0xffffffffu, best_k, lane & 24);
    if (!live) continue;
    if (q == 0) a.idx_out[static_cast<size_t>(m) * a.groups + g] = static_cast<uint8_t>(best_k);
    if (a.onehot_packed && q < 4) {
      // group g occupies columns [32g, 32g+32): 4 chunks of 8 bf16, one per lane q = 0..3
      const int rel = best_k - q * 8;
      const uint32_t one = (rel & 1) ? 0x3F800000u : 0x00003F80u;   // bf16 1.0 in the high / low half
      const int word = (rel >= 0 && rel < 8) ? (rel >> 1) : -1;     // selects, not an indexed local array
      const size_t idx = packed_index(static_cast<size_t>(m), static_cast<size_t>(g * 32 + q * 8),
                                      static_cast<size_t>(a.kpad), kTileM);
      *reinterpret_cast<uint4*>(a.onehot_packed + idx) =
          make_uint4(word == 0 ? one : 0u, word == 1 ? one : 0u, word == 2 ? one : 0u, word == 3 ? one : 0u);
    }
    if (a.onehot_f32) {
      const int rel = best_k - q * 4;
      reinterpret_cast<float4*>(a.onehot_f32 + static_cast<size_t>(m) * a.ld_f32 + g * 32)[q] =
          make_float4(rel == 0 ? 1.f : 0.f, rel == 1 ? 1.f : 0.f, rel == 2 ? 1.f : 0.f, rel == 3 ? 1.f : 0.f);
    }
  }
}

}  // namespace rowops
}  // namespace rlsb
