/* rlsb_detmath.h — bit-reproducible scalar math shared by the CUDA kernels and the C oracle.
 *
 * The categorical draws of the imagination path (reference: utils/dists.py:177-179 ->
 * torch OneHotCategorical.sample -> aten::multinomial, i.e. an exponential race == Gumbel-max)
 * must give bit-identical indices on host and device for identical logits and uniforms.  A libm
 * logf differs between glibc and CUDA in the last ulp, so the noise transform is written here
 * with correctly-rounded IEEE operations only (add, mul, div, int<->float bit moves); the same
 * source compiles as CUDA device code (explicit *_rn intrinsics, never contracted to FMA) and as
 * plain C (build with -ffp-contract=off).
 *
 * Also: Philox4x32-10 counter-based generator and the uniform conversion.
 */
#ifndef RLSB_DETMATH_H
#define RLSB_DETMATH_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDA_ARCH__)
#define RLSB_HD __host__ __device__ __forceinline__
#define RLSB_ADD(a, b) __fadd_rn((a), (b))
#define RLSB_SUB(a, b) __fsub_rn((a), (b))
#define RLSB_MUL(a, b) __fmul_rn((a), (b))
#define RLSB_DIV(a, b) __fdiv_rn((a), (b))
#elif defined(__CUDACC__)
#define RLSB_HD __host__ __device__ __forceinline__
#define RLSB_ADD(a, b) ((a) + (b))
#define RLSB_SUB(a, b) ((a) - (b))
#define RLSB_MUL(a, b) ((a) * (b))
#define RLSB_DIV(a, b) ((a) / (b))
#else
#define RLSB_HD static inline
#define RLSB_ADD(a, b) ((a) + (b))
#define RLSB_SUB(a, b) ((a) - (b))
#define RLSB_MUL(a, b) ((a) * (b))
#define RLSB_DIV(a, b) ((a) / (b))
#endif

RLSB_HD uint32_t rlsb_f2u(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
#endif
}
RLSB_HD float rlsb_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

/* natural log for positive normal floats (fdlibm/musl logf scheme, every op rounded once). */
RLSB_HD float rlsb_logf(float x) {
  const float ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f;
  const float Lg1 = 0.66666662693f, Lg2 = 0.40000972152f, Lg3 = 0.28498786688f,
              Lg4 = 0.24279078841f;
  uint32_t ix = rlsb_f2u(x);
  ix += 0x3f800000u - 0x3f3504f3u;
  int k = (int)(ix >> 23) - 0x7f;
  ix = (ix & 0x007fffffu) + 0x3f3504f3u;
  float m = rlsb_u2f(ix);
  float f = RLSB_SUB(m, 1.0f);
  float s = RLSB_DIV(f, RLSB_ADD(2.0f, f));
  float z = RLSB_MUL(s, s);
  float w = RLSB_MUL(z, z);
  float t1 = RLSB_MUL(w, RLSB_ADD(Lg2, RLSB_MUL(w, Lg4)));
  float t2 = RLSB_MUL(z, RLSB_ADD(Lg1, RLSB_MUL(w, Lg3)));
  float R = RLSB_ADD(t2, t1);
  float hfsq = RLSB_MUL(RLSB_MUL(0.5f, f), f);
  float dk = (float)k;
  /* s*(hfsq+R) + dk*ln2_lo - hfsq + f + dk*ln2_hi */
  float r = RLSB_MUL(s, RLSB_ADD(hfsq, R));
  r = RLSB_ADD(r, RLSB_MUL(dk, ln2_lo));
  r = RLSB_SUB(r, hfsq);
  r = RLSB_ADD(r, f);
  r = RLSB_ADD(r, RLSB_MUL(dk, ln2_hi));
  return r;
}

/* clamp a uniform into the open interval the transform is defined on */
RLSB_HD float rlsb_clamp_uniform(float u) {
  const float lo = 1e-20f, hi = 0.99999994f; /* 1 - 2^-24 */
  u = u < lo ? lo : u;
  u = u > hi ? hi : u;
  return u;
}

/* standard Gumbel from a uniform: g = -log(-log(u)) */
RLSB_HD float rlsb_gumbel(float u) {
  float t = rlsb_logf(rlsb_clamp_uniform(u)); /* < 0 */
  return -rlsb_logf(-t);
}

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al. 2011).  counter = 4 x u32, key = 2 x u32.
 * ---------------------------------------------------------------------------------------- */
RLSB_HD void rlsb_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                             uint32_t k1, uint32_t out[4]) {
  for (int i = 0; i < 10; ++i) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* 24-bit uniform strictly inside (0,1): exact in fp32 */
RLSB_HD float rlsb_u32_to_uniform(uint32_t x) {
  return RLSB_ADD(RLSB_MUL((float)(x >> 8), 5.9604644775390625e-08f), 2.98023223876953125e-08f);
}

/* Noise addressing used by the imagination kernels (counters are GLOBAL start-state
 * indices so results do not depend on how start states are sharded over GPUs).
 *   stream 0: latent categorical noise, element e in [0,1024) of step t for start state n
 *   stream 1: action noise, element e in [0,A)
 * counter = (n, t, stream, e/4), component e%4. */
RLSB_HD float rlsb_noise_uniform(uint64_t seed, uint32_t n, uint32_t t, uint32_t stream, uint32_t e) {
  uint32_t o[4];
  rlsb_philox4x32(n, t, stream, e >> 2, (uint32_t)seed, (uint32_t)(seed >> 32), o);
  return rlsb_u32_to_uniform(o[e & 3]);
}

/* standard normal from two uniforms (Box-Muller, deterministic log; sin/cos are not needed
 * bit-exact across host/device because normals are only generated on the device in bench
 * mode — parity tests pass explicit normals). */

#endif /* RLSB_DETMATH_H */
