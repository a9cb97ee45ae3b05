/* rlsb_oracle.c — TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * Plain-C restatement of the integer / bit-exact parts of the DreamerV2 imagination path of
 * Midren/rl_sandbox, used as the checker for the CUDA kernels:
 *
 *   orc_lambda_return      ImaginativeCritic._lambda_return      agents/dreamer/ac.py:52-62
 *                          + discount shift/cumprod              agents/dreamer_v2.py:192-197
 *                          + advantage                           agents/dreamer/ac.py:118
 *   orc_sample_categorical OneHotCategorical(ST).sample()        utils/dists.py:177-179,
 *                          == aten::multinomial(probs, 1) == exponential race == Gumbel-max
 *                          (agents/dreamer/common.py:27-28, agents/dreamer/rssm.py:34-37)
 *   orc_philox_uniform     counter-based noise (no reference counterpart: the reference draws
 *                          from torch's global generator)
 *
 * The categorical draw is defined as idx = argmax_k(logit_k + g(u_k)), g(u) = -log(-log u), ties
 * to the lowest k (torch.argmax).  g is evaluated with a log built from correctly-rounded IEEE
 * single operations only, so the device can reproduce it bit for bit; orc_logf is written here
 * independently of the product's header and is itself checked against libm in
 * tests/test_oracle.py.  Build: gcc -O2 -ffp-contract=off -fPIC -shared (oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* fdlibm-style logf for positive normal x; one rounding per operation, no FMA. */
float orc_logf(float x) {
  const float ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f;
  const float Lg1 = 0.66666662693f, Lg2 = 0.40000972152f, Lg3 = 0.28498786688f, Lg4 = 0.24279078841f;
  uint32_t ix = f2u(x);
  ix += 0x3f800000u - 0x3f3504f3u;
  int k = (int)(ix >> 23) - 0x7f;
  ix = (ix & 0x007fffffu) + 0x3f3504f3u;
  volatile float f = u2f(ix) - 1.0f;
  volatile float s = f / (2.0f + f);
  volatile float z = s * s;
  volatile float w = z * z;
  volatile float a = w * Lg4;  a = Lg2 + a;  volatile float t1 = w * a;
  volatile float b = w * Lg3;  b = Lg1 + b;  volatile float t2 = z * b;
  volatile float R = t2 + t1;
  volatile float hf = 0.5f * f; volatile float hfsq = hf * f;
  volatile float dk = (float)k;
  volatile float r = hfsq + R; r = s * r;
  volatile float c = dk * ln2_lo; r = r + c;
  r = r - hfsq;
  r = r + f;
  c = dk * ln2_hi; r = r + c;
  return r;
}

float orc_gumbel(float u) {
  const float lo = 1e-20f, hi = 0.99999994f;
  if (u < lo) u = lo;
  if (u > hi) u = hi;
  float t = orc_logf(u);
  return -orc_logf(-t);
}

void orc_gumbel_array(const float* u, float* g, int64_t n) {
  for (int64_t i = 0; i < n; ++i) g[i] = orc_gumbel(u[i]);
}

void orc_logf_array(const float* x, float* y, int64_t n) {
  for (int64_t i = 0; i < n; ++i) y[i] = orc_logf(x[i]);
}

/* idx[r] = argmax_k(logits[r,k] + gumbel(u[r,k])), first maximum wins */
void orc_sample_categorical(const float* logits, const float* uniforms, int64_t rows, int classes,
                            int32_t* idx) {
  for (int64_t r = 0; r < rows; ++r) {
    float best = 0.f; int bk = 0;
    for (int k = 0; k < classes; ++k) {
      volatile float s = logits[r * classes + k] + orc_gumbel(uniforms[r * classes + k]);
      if (k == 0 || s > best) { best = s; bk = k; }
    }
    idx[r] = bk;
  }
}

/* Philox4x32-10, Salmon et al. SC'11 */
static void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t o[4]) {
  for (int i = 0; i < 10; ++i) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

void orc_philox_raw(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* o) {
  philox(c0, c1, c2, c3, k0, k1, o);
}

/* out[i] = uniform for (start state n0 + i/per_row, step t, stream, element i%per_row):
 * counter (n, t, stream, e/4), word e%4, u = (x>>8)*2^-24 + 2^-25 */
void orc_philox_uniform(uint64_t seed, uint32_t n0, uint32_t t, uint32_t stream, int per_row, int64_t count,
                        float* out) {
  for (int64_t i = 0; i < count; ++i) {
    uint32_t n = n0 + (uint32_t)(i / per_row), e = (uint32_t)(i % per_row), o[4];
    philox(n, t, stream, e >> 2, (uint32_t)seed, (uint32_t)(seed >> 32), o);
    volatile float a = (float)(o[e & 3] >> 8) * 5.9604644775390625e-08f;
    out[i] = a + 2.98023223876953125e-08f;
  }
}

/* time-major (T,N): V_H = v_H; V_t = r_t + d_t*((1-l) v_{t+1} + l V_{t+1}) in the reference's
 * operation order, one fp32 rounding per op (torch evaluates each op as a separate kernel). */
void orc_lambda_return(const float* r, const float* v, const float* d, int T, int64_t N, double lambda_,
                       float* vs, float* w, float* adv) {
  const int H = T - 1;
  const float c1 = (float)(1.0 - lambda_), c2 = (float)lambda_;
  for (int64_t n = 0; n < N; ++n) {
    float V = v[(int64_t)H * N + n];
    for (int t = H - 1; t >= 0; --t) {
      volatile float a = c1 * v[(int64_t)(t + 1) * N + n];
      volatile float b = c2 * V;
      volatile float mix = a + b;
      volatile float m = d[(int64_t)t * N + n] * mix;
      float out = r[(int64_t)t * N + n] + m;
      if (adv && t <= H - 2) adv[(int64_t)t * N + n] = V - v[(int64_t)t * N + n];
      vs[(int64_t)t * N + n] = out;
      V = out;
    }
    if (w) {
      float acc = 1.0f;
      for (int t = 0; t < T; ++t) {
        w[(int64_t)t * N + n] = acc;
        volatile float p = acc * d[(int64_t)t * N + n];
        acc = p;
      }
    }
  }
}
