// rlsb_ac.cu — K4: the actor-critic update of DreamerV2.train (agents/dreamer_v2.py:199-211) on the
// states K1 imagined: critic + actor MLP forward, both losses, and the full backward pass down to
// fp32 parameter gradients in nn.Linear layout — a chain of tcgen05 GEMMs with fused epilogues:
//
//   forward  (ac.py:68-72,113-116)  5 grouped launches (actor | critic), LayerNorm + ELU fused, x_hat / rstd kept
//   losses   (ac.py:70-81,117-146)  ac_loss_kernel: Normal(v,1) log-prob, categorical log-prob / entropy,
//                                   reinforce + entropy terms, d(loss)/d(head outputs), metrics
//   backward (optimizer.py:55-57)   per layer: wgrad_kernel (dW, db via a "ones" K tile) and
//                                   gemm_kernel<EPI_BWD> (dX with ELU' + LayerNorm backward + d_gamma/d_beta)
//
// Rows are the H per-step state images K1 kept (rlsb_imagine_out::determ_packed / stoch_packed): step
// t occupies rows [t*m_pad, t*m_pad + N); rows beyond N in a step are padding and carry zero weight.
#include <cstring>

#include "../../include/rlsb.h"
#include "rlsb_count.cuh"
#include "rlsb_detmath.h"
#include "rlsb_gemm.cuh"
#include "rlsb_kernels.cuh"
#include "rlsb_ptx.cuh"
#include "rlsb_wgrad.cuh"

namespace rlsb {

namespace {

inline int ru(int x, int m) { return (x + m - 1) / m * m; }
inline size_t rus(size_t x, size_t m) { return (x + m - 1) / m * m; }
size_t place(size_t& cursor, size_t bytes) {
  cursor = rus(cursor, 1024);
  size_t off = cursor;
  cursor += bytes;
  return off;
}

constexpr int kG = 2;  // group 0 = actor, group 1 = critic
constexpr int kScalars = 16;

struct AcLayer {
  int N = 0, RB = 0, kp = 0;  // forward: valid outputs, row block (one n-block), padded K
  size_t w_off = 0, bias_off = 0, g_off = 0, b_off = 0;
  int t_RB = 0, t_kp = 0;     // transposed image (dX GEMM): rows = in-features, K = padded out-features
  size_t wt_off = 0;
};

struct AcPlan {
  int D, S, A, Hd, Dp, Sp, Hp, Aout, H;
  AcLayer L[5];
  size_t ones_off;
  size_t packed_bytes;
};

int make_ac_plan(const rlsb_ac_cfg& c, AcPlan& P) {
  if (c.classes != 32 || c.groups <= 0 || c.groups > 64 || c.D <= 0 || c.A <= 0 || c.hidden <= 0 || c.H < 2) return -10;
  P.D = c.D; P.S = c.groups * c.classes; P.A = c.A; P.Hd = c.hidden; P.H = c.H;
  P.Dp = ru(P.D, 64); P.Sp = ru(P.S, 64); P.Hp = ru(P.Hd, 64);
  P.Aout = c.discrete ? c.A : 2 * c.A;   // TruncatedNormal head: (mean | std) pre-activations (dists.py:187-190)
  if (P.Aout > 32 || ru(P.Hd, 32) > 512) return -12;
  size_t cur = 0;
  for (int l = 0; l < 5; ++l) {
    AcLayer& L = P.L[l];
    L.N = (l == 4) ? P.Aout : P.Hd;
    L.RB = ru(L.N, 32);
    L.kp = (l == 0) ? (P.Dp + P.Sp) : P.Hp;
    L.w_off = place(cur, static_cast<size_t>(kG) * L.RB * L.kp * 2);
    L.bias_off = place(cur, static_cast<size_t>(kG) * L.RB * 4);
    L.g_off = place(cur, static_cast<size_t>(kG) * L.RB * 4);
    L.b_off = place(cur, static_cast<size_t>(kG) * L.RB * 4);
    if (l >= 1) {
      L.t_RB = ru(P.Hd, 32);
      L.t_kp = ru(L.N, 64);
      L.wt_off = place(cur, static_cast<size_t>(kG) * L.t_RB * L.t_kp * 2);
    }
  }
  P.ones_off = place(cur, 128 * 64 * 2);
  P.packed_bytes = rus(cur, 1024);
  return 0;
}

struct AcWorkspace {
  size_t x[4], pre[4], rstd[4], head_out, dy4, dp[2], col_part, partial, accum;
  int m_pad;       // rows per step image
  long long M;     // H * m_pad
  size_t bytes;
};

int make_ac_workspace(const AcPlan& P, long long N, AcWorkspace& W) {
  W.m_pad = ru(static_cast<int>(N), 128);
  W.M = static_cast<long long>(P.H) * W.m_pad;
  const size_t M = static_cast<size_t>(W.M);
  size_t cur = 0;
  for (int l = 0; l < 4; ++l) {
    W.x[l] = place(cur, kG * M * P.Hp * 2);
    W.pre[l] = place(cur, kG * M * P.Hp * 2);
    W.rstd[l] = place(cur, kG * M * 4);
  }
  W.head_out = place(cur, kG * M * 32 * 4);
  W.dy4 = place(cur, kG * M * 64 * 2);
  for (int i = 0; i < 2; ++i) W.dp[i] = place(cur, kG * M * P.Hp * 2);
  W.col_part = place(cur, static_cast<size_t>(256) * kG * 2 * 512 * 4);
  // weight-gradient partial tiles: the largest layer decides
  size_t pmax = 0;
  for (int l = 0; l < 5; ++l) {
    WgradParams wp{};
    wp.n_tiles = ru(P.L[l].N, 64) / 64;
    wp.G = kG;
    wp.m_tiles = static_cast<int>(M / 128);
    if (l == 0) {
      wp.n_seg = 3;
      wp.x_ktiles[0] = P.Dp / 64; wp.x_ktiles[1] = P.Sp / 64; wp.x_ktiles[2] = 1;
    } else {
      wp.n_seg = 2;
      wp.x_ktiles[0] = P.Hp / 64; wp.x_ktiles[1] = 1;
    }
    const int e = plan_wgrad(wp);
    if (e != 0) return e;
    const size_t b = wgrad_partial_bytes(wp);
    if (b > pmax) pmax = b;
  }
  W.partial = place(cur, pmax);
  W.accum = place(cur, kScalars * sizeof(double));
  W.bytes = rus(cur, 1024);
  return 0;
}

int copy_pad2(const float* src, int n, float* dst, int n_pad, float fill, cudaStream_t s) {
  return launch_copy_pad(src, n, dst, n_pad, fill, s);
}

// the "ones" K tile: element (row, 0) = 1, everything else 0 (bias gradient = dY^T * 1)
__global__ void ones_tile_kernel(__nv_bfloat16* dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 16-byte chunk: 128 rows x 8 chunks
  if (i >= 128 * 8) return;
  const int row = i >> 3, pos = i & 7;
  const int chunk = pos ^ (row & 7);   // logical chunk stored at this position
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (chunk == 0) v.x = 0x00003F80u;   // bf16 1.0 in element 0
  reinterpret_cast<uint4*>(dst)[i] = v;
}

// ------------------------------------------------------------------------------------------
// losses and d(loss)/d(head outputs)
// ------------------------------------------------------------------------------------------
enum AcAccum {
  ACC_LOSS_CRITIC = 0, ACC_REINFORCE, ACC_ENTROPY, ACC_PRED, ACC_TARGET, ACC_LAMBDA, ACC_AVG_SD, ACC_MEAN_VAL,
  ACC_DYNAMICS, ACC_AVG_VAL, ACC_COUNT
};
constexpr int kAccMinKey = 12, kAccMaxKey = 13;   // int-encoded float min / max live in accum[12], accum[13] (low words)

__device__ __forceinline__ int float_key(float f) {
  const int b = __float_as_int(f);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float key_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

struct AcLossArgs {
  const float* head_out;   // [2][M][32]
  long long group_stride;
  int m_pad, N, H, A;
  const float* vs;         // (H, N)
  const float* w;          // (H+1, N)
  const float* values;     // (H+1, N) target critic
  const float* actions;    // (H+1, N, A)
  const float* g_actions;  // (H, N, A) d loss / d a_t from rlsb_imagine_bwd (continuous actor) or nullptr
  int discrete;
  float rho, eta;
  int metrics_samples;
  uint64_t seed;
  const uint64_t* seed_ptr;   // device-resident key overriding `seed` (CUDA-graph replays)
  __nv_bfloat16* dy4;      // packed [2][M x 64]
  double* accum;
};

__global__ void __launch_bounds__(128) ac_loss_kernel(const AcLossArgs a) {
  const long long M = static_cast<long long>(a.H) * a.m_pad;
  const long long m = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  float acc[ACC_COUNT];
#pragma unroll
  for (int i = 0; i < ACC_COUNT; ++i) acc[i] = 0.f;
  const uint64_t key = a.seed_ptr ? __ldg(a.seed_ptr) + 0x9E3779B97F4A7C15ull : a.seed;
  if (m < M) {
    const int t = static_cast<int>(m / a.m_pad);
    const int i = static_cast<int>(m - static_cast<long long>(t) * a.m_pad);
    const bool valid = i < a.N;
    float dyc = 0.f;          // critic: d loss / d v
    float dya[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) dya[k] = 0.f;
    if (valid) {
      const size_t ti = static_cast<size_t>(t) * a.N + i;
      const float wt = __ldg(a.w + ti);
      // ---- critic (ac.py:68-72): -mean(Normal(v, 1).log_prob(vs) * w) over H*N elements ----------
      const float v = a.head_out[a.group_stride + m * 32];
      const float target = __ldg(a.vs + ti);
      const float diff = target - v;
      acc[ACC_LOSS_CRITIC] = (0.5f * diff * diff + 0.91893853320467274f) * wt;
      dyc = -diff * wt / (static_cast<float>(a.H) * static_cast<float>(a.N));
      acc[ACC_PRED] = v;
      acc[ACC_TARGET] = __ldg(a.values + ti);
      acc[ACC_LAMBDA] = target;
      // ---- actor (ac.py:113-146), states 0..H-2 ---------------------------------------------------
      if (!a.discrete) {
        // ---- TruncatedNormal actor (dists.py:108-129,187-190): mu = tanh(raw), sd = 2 sigmoid(raw_s / 2) + 0.1 ----
        // The loss terms cover states 0..H-2 (ac.py:113-135); the dynamics gradient d loss / d a_t reaches the
        // actor through EVERY sampled action a_0..a_{H-1} (a_{H-1} moves s_H and with it the bootstrap value).
        const bool in_loss = t < a.H - 1;
        const float* lg = a.head_out + m * 32;
        const float vnext = in_loss ? __ldg(a.vs + ti + a.N) : 0.f;
        const float adv = in_loss ? vnext - __ldg(a.values + ti) : 0.f;
        const float* act = a.actions + (ti + a.N) * a.A;
        const float inv_cnt = in_loss ? 1.0f / (static_cast<float>(a.H - 1) * static_cast<float>(a.N)) : 0.f;
        if (in_loss) acc[ACC_DYNAMICS] = -(1.0f - a.rho) * vnext * wt;
        float logp = 0.f, ent = 0.f, sum_mu = 0.f, sum_avg = 0.f, sum_sd = 0.f;
        float smin = 3.0e38f, smax = -3.0e38f;
        for (int k = 0; k < a.A; ++k) {
          const float mu = tanhf(lg[k]);
          const float sg = 1.0f / (1.0f + expf(-0.5f * lg[a.A + k]));
          const float sd = 2.0f * sg + 0.1f;
          const float e = (__ldg(act + k) - mu) / sd;
          const float lsd = logf(sd);
          logp += -0.5f * e * e - lsd - 0.91893853320467274f;
          ent += 1.4189385332046727f + lsd;
          const float ga = a.g_actions ? __ldg(a.g_actions + ti * a.A + k) : 0.f;
          const float g_mu = ga + (-a.rho * wt * adv * (e / sd)) * inv_cnt;
          const float g_sd = ga * e + (-a.rho * wt * adv * ((e * e - 1.0f) / sd) - a.eta * wt / sd) * inv_cnt;
          dya[k] = g_mu * (1.0f - mu * mu);
          dya[a.A + k] = g_sd * sg * (1.0f - sg);
          sum_mu += mu;
          if (a.metrics_samples > 0 && in_loss) {   // `metrics_samples` draws mu + sd * eps per element (ac.py:137-143)
            float s1 = 0.f, s2 = 0.f, emin = 3.0e38f, emax = -3.0e38f;
            for (int sidx = 0; sidx < a.metrics_samples; sidx += 4) {
              uint32_t o[4];
              rlsb_philox4x32(static_cast<uint32_t>(m), static_cast<uint32_t>(m >> 32) ^ (static_cast<uint32_t>(k) << 8), 7u,
                              static_cast<uint32_t>(sidx >> 2), static_cast<uint32_t>(key),
                              static_cast<uint32_t>(key >> 32), o);
              float z[4];
              const float r0 = sqrtf(-2.0f * __logf(rlsb_u32_to_uniform(o[0])));
              const float r1 = sqrtf(-2.0f * __logf(rlsb_u32_to_uniform(o[2])));
              __sincosf(6.2831853071795865f * rlsb_u32_to_uniform(o[1]), &z[1], &z[0]);
              __sincosf(6.2831853071795865f * rlsb_u32_to_uniform(o[3]), &z[3], &z[2]);
              z[0] *= r0; z[1] *= r0; z[2] *= r1; z[3] *= r1;
              for (int j = 0; j < 4 && sidx + j < a.metrics_samples; ++j) {
                s1 += z[j];
                s2 = fmaf(z[j], z[j], s2);
                emin = fminf(emin, z[j]);
                emax = fmaxf(emax, z[j]);
              }
            }
            const float inv_s = 1.0f / static_cast<float>(a.metrics_samples);
            const float mbar = s1 * inv_s;
            sum_avg += mu + sd * mbar;
            sum_sd += sd * sqrtf(fmaxf(s2 * inv_s - mbar * mbar, 0.f));
            smin = fminf(smin, mu + sd * emin);
            smax = fmaxf(smax, mu + sd * emax);
          }
        }
        if (in_loss) {
          acc[ACC_REINFORCE] = -a.rho * logp * wt * adv;
          acc[ACC_ENTROPY] = -a.eta * ent * wt;
          acc[ACC_MEAN_VAL] = sum_mu;
          acc[ACC_AVG_VAL] = sum_avg;
          acc[ACC_AVG_SD] = sum_sd;
        }
        if (a.metrics_samples > 0 && in_loss) {
          atomicMin(reinterpret_cast<int*>(a.accum + kAccMinKey), float_key(smin));
          atomicMax(reinterpret_cast<int*>(a.accum + kAccMaxKey), float_key(smax));
        }
      } else if (t < a.H - 1) {
        const float* lg = a.head_out + m * 32;
        const float adv = __ldg(a.vs + ti + a.N) - __ldg(a.values + ti);   // (vs[1:] - baseline[:-2])
        const float* act = a.actions + (ti + a.N) * a.A;                    // actions[1:-1]
        float mx = lg[0];
        for (int k = 1; k < a.A; ++k) mx = fmaxf(mx, lg[k]);
        float se = 0.f;
        for (int k = 0; k < a.A; ++k) se += expf(lg[k] - mx);
        const float lse = mx + logf(se);
        float ent = 0.f, logp_a = 0.f;
        int a_idx = 0;
        float best = __ldg(act);
        for (int k = 1; k < a.A; ++k) {
          const float x = __ldg(act + k);
          if (x > best) { best = x; a_idx = k; }
        }
        for (int k = 0; k < a.A; ++k) {
          const float lp = lg[k] - lse;
          const float pk = expf(lp);
          ent -= pk * lp;
          if (k == a_idx) logp_a = lp;
        }
        acc[ACC_REINFORCE] = -a.rho * logp_a * wt * adv;
        acc[ACC_ENTROPY] = -a.eta * ent * wt;
        const float inv_cnt = 1.0f / (static_cast<float>(a.H - 1) * static_cast<float>(a.N));
        // cumulative probabilities in registers: every loop over the classes below is fully unrolled over 32 with a
        // uniform (k < A) guard, so no array is indexed dynamically (no local memory)
        float cdf[32];
        float run = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          cdf[k] = 3.0e38f;
          if (k < a.A) {
            const float lp = lg[k] - lse;
            const float pk = expf(lp);
            dya[k] = (-a.rho * wt * adv * ((k == a_idx ? 1.0f : 0.f) - pk) + a.eta * wt * pk * (lp + ent)) * inv_cnt;
            run += pk;
            cdf[k] = run;
          }
        }
        acc[ACC_MEAN_VAL] = run;   // sum_k p_k (dist.mean summed over the action axis)
        // ---- statistics of `metrics_samples` draws per element (ac.py:137-143): the empirical class
        //      frequencies f_k decide avg_val (= sum f_k / A) and avg_sd (= mean sqrt(f_k (1 - f_k))).
        //      A draw u falls into class k iff cdf[k-1] <= u < cdf[k] (the last class takes the rest), so the class
        //      counts are differences of the cumulative counts below[k] = #{u < cdf[k]} — branch-free compare-and-add.
        if (a.metrics_samples > 0) {
          int below[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) below[k] = 0;
          for (int sidx = 0; sidx < a.metrics_samples; sidx += 4) {
            uint32_t o[4];
            rlsb_philox4x32(static_cast<uint32_t>(m), static_cast<uint32_t>(m >> 32), 7u, static_cast<uint32_t>(sidx >> 2),
                            static_cast<uint32_t>(key), static_cast<uint32_t>(key >> 32), o);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (sidx + j < a.metrics_samples) {
                const float u = rlsb_u32_to_uniform(o[j]) * run;
#pragma unroll
                for (int k = 0; k < 32; ++k)
                  if (k < a.A - 1) below[k] += (u < cdf[k]) ? 1 : 0;
              }
            }
          }
          float sd = 0.f;
          const float inv_s = 1.0f / static_cast<float>(a.metrics_samples);
          int prev = 0;
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            if (k < a.A) {
              const int upto = (k < a.A - 1) ? below[k] : a.metrics_samples;
              const float f = static_cast<float>(upto - prev) * inv_s;
              prev = upto;
              sd += sqrtf(f * (1.0f - f));
            }
          }
          acc[ACC_AVG_SD] = sd;
        }
      }
    }
    // ---- d(loss)/d(head outputs) as the packed bf16 operand of the backward GEMMs --------------------
    const size_t tile = static_cast<size_t>(m >> 7) * (kTileM * kTileK);
    const int row = static_cast<int>(m & 127);
    __nv_bfloat16* da = a.dy4 + tile + static_cast<size_t>(row) * kTileK;
    __nv_bfloat16* dc = da + static_cast<size_t>(M) * 64;
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      uint4 va = make_uint4(0u, 0u, 0u, 0u), vc = make_uint4(0u, 0u, 0u, 0u);
      if (ch < 4) {
        __nv_bfloat162 p0 = __floats2bfloat162_rn(dya[ch * 8 + 0], dya[ch * 8 + 1]);
        __nv_bfloat162 p1 = __floats2bfloat162_rn(dya[ch * 8 + 2], dya[ch * 8 + 3]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(dya[ch * 8 + 4], dya[ch * 8 + 5]);
        __nv_bfloat162 p3 = __floats2bfloat162_rn(dya[ch * 8 + 6], dya[ch * 8 + 7]);
        va = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                        *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
      }
      if (ch == 0) {
        __nv_bfloat162 p0 = __floats2bfloat162_rn(dyc, 0.f);
        vc.x = *reinterpret_cast<uint32_t*>(&p0);
      }
      const int pos = (ch ^ (row & 7)) << 3;
      *reinterpret_cast<uint4*>(da + pos) = va;
      *reinterpret_cast<uint4*>(dc + pos) = vc;
    }
  }
  // ---- block reduction of the scalar sums -> double accumulators ------------------------------------
  __shared__ float red[4][ACC_COUNT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < ACC_COUNT; ++i) {
    float v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < ACC_COUNT) {
    const double s = static_cast<double>(red[0][threadIdx.x]) + red[1][threadIdx.x] + red[2][threadIdx.x] +
                     red[3][threadIdx.x];
    atomicAdd(a.accum + threadIdx.x, s);
  }
}

__global__ void ac_finalize_kernel(const double* accum, int H, long long N, int A, int metrics_samples, int discrete,
                                   float* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double nc = static_cast<double>(H) * N, na = static_cast<double>(H - 1) * N;
  const float lc = static_cast<float>(accum[ACC_LOSS_CRITIC] / nc);
  const float lr = static_cast<float>(accum[ACC_REINFORCE] / na);
  const float le = static_cast<float>(accum[ACC_ENTROPY] / na);
  out[RLSB_AC_LOSS_CRITIC] = lc;
  out[RLSB_AC_LOSS_ACTOR_REINFORCE] = lr;
  const float ld = static_cast<float>(accum[ACC_DYNAMICS] / na);   // 0 for rho == 1 (ac.py:124-125)
  out[RLSB_AC_LOSS_ACTOR_DYNAMICS] = ld;
  out[RLSB_AC_LOSS_ACTOR_ENTROPY] = le;
  out[RLSB_AC_LOSS_ACTOR] = lr + ld + le;
  out[RLSB_AC_CRITIC_AVG_TARGET] = static_cast<float>(accum[ACC_TARGET] / nc);
  out[RLSB_AC_CRITIC_AVG_LAMBDA] = static_cast<float>(accum[ACC_LAMBDA] / nc);
  out[RLSB_AC_CRITIC_AVG_PRED] = static_cast<float>(accum[ACC_PRED] / nc);
  const float mean_val = static_cast<float>(accum[ACC_MEAN_VAL] / (na * A));
  out[RLSB_AC_ACTOR_MEAN_VAL] = mean_val;
  if (metrics_samples > 0 && !discrete) {
    out[RLSB_AC_ACTOR_AVG_VAL] = static_cast<float>(accum[ACC_AVG_VAL] / (na * A));
    out[RLSB_AC_ACTOR_AVG_SD] = static_cast<float>(accum[ACC_AVG_SD] / (na * A));
    out[RLSB_AC_ACTOR_MIN_VAL] = key_float(*reinterpret_cast<const int*>(accum + kAccMinKey));
    out[RLSB_AC_ACTOR_MAX_VAL] = key_float(*reinterpret_cast<const int*>(accum + kAccMaxKey));
  } else if (metrics_samples > 0) {
    out[RLSB_AC_ACTOR_AVG_VAL] = 1.0f / static_cast<float>(A);   // sum_k f_k == 1 for every element
    out[RLSB_AC_ACTOR_AVG_SD] = static_cast<float>(accum[ACC_AVG_SD] / (na * A));
    out[RLSB_AC_ACTOR_MIN_VAL] = A > 1 ? 0.f : 1.f;               // one-hot draws
    out[RLSB_AC_ACTOR_MAX_VAL] = 1.f;
  } else {
    out[RLSB_AC_ACTOR_AVG_VAL] = mean_val;
    out[RLSB_AC_ACTOR_AVG_SD] = 0.f;
    out[RLSB_AC_ACTOR_MIN_VAL] = 0.f;
    out[RLSB_AC_ACTOR_MAX_VAL] = 0.f;
  }
}

#define RLSB_TRY(expr)            \
  do {                            \
    int _e = (expr);              \
    if (_e != 0) return _e;       \
  } while (0)

}  // namespace

}  // namespace rlsb

using namespace rlsb;

extern "C" size_t rlsb_packed_rows(int64_t N) { return static_cast<size_t>((N + 127) / 128 * 128); }

extern "C" size_t rlsb_ac_packed_bytes(const rlsb_ac_cfg* cfg) {
  AcPlan P;
  if (!cfg || make_ac_plan(*cfg, P) != 0) return 0;
  return P.packed_bytes;
}

extern "C" size_t rlsb_ac_workspace_bytes(const rlsb_ac_cfg* cfg, int64_t N) {
  AcPlan P;
  if (!cfg || N <= 0 || make_ac_plan(*cfg, P) != 0) return 0;
  AcWorkspace W;
  if (make_ac_workspace(P, N, W) != 0) return 0;
  return W.bytes;
}

extern "C" int rlsb_ac_pack(const rlsb_ac_cfg* cfg, const rlsb_mlp_params* actor, const rlsb_mlp_params* critic,
                            void* packed, void* stream_) {
  if (!cfg || !actor || !critic || !packed) return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  AcPlan P;
  RLSB_TRY(make_ac_plan(*cfg, P));
  LaunchBatchScope batch(s);   // the small pack / pad launches below are queued and issued as multi-job kernels
  uint8_t* base = static_cast<uint8_t*>(packed);
  for (int l = 0; l < 5; ++l) {
    const AcLayer& L = P.L[l];
    for (int g = 0; g < kG; ++g) {
      const rlsb_mlp_params* hp = g == 0 ? actor : critic;
      if (!hp->w[l] || !hp->b[l]) return -20;
      const int n_out = (l == 4) ? (g == 0 ? P.Aout : 1) : P.Hd;
      const int k_in = (l == 0) ? (P.D + P.S) : P.Hd;
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(base + L.w_off) + static_cast<size_t>(g) * L.RB * L.kp;
      if (l == 0) {
        PackSeg segs[2] = {{0, 0, P.D}, {P.Dp, P.D, P.S}};
        RLSB_TRY(launch_pack(hp->w[l], k_in, n_out, dst, L.RB, L.RB, L.kp, 2, segs, s));
      } else {
        PackSeg segs[1] = {{0, 0, P.Hd}};
        RLSB_TRY(launch_pack(hp->w[l], k_in, n_out, dst, L.RB, L.RB, L.kp, 1, segs, s));
        // transposed image for dX = dY W: rows = in-features, K = out-features
        __nv_bfloat16* dt = reinterpret_cast<__nv_bfloat16*>(base + L.wt_off) + static_cast<size_t>(g) * L.t_RB * L.t_kp;
        RLSB_TRY(launch_pack_transposed(hp->w[l], k_in, n_out, k_in, dt, L.t_RB, L.t_RB, L.t_kp, s));
      }
      RLSB_TRY(copy_pad2(hp->b[l], n_out, reinterpret_cast<float*>(base + L.bias_off) + static_cast<size_t>(g) * L.RB,
                         L.RB, 0.f, s));
      if (l < 4) {
        RLSB_TRY(copy_pad2(hp->ln_g[l], L.N, reinterpret_cast<float*>(base + L.g_off) + static_cast<size_t>(g) * L.RB,
                           L.RB, 1.f, s));
        RLSB_TRY(copy_pad2(hp->ln_b[l], L.N, reinterpret_cast<float*>(base + L.b_off) + static_cast<size_t>(g) * L.RB,
                           L.RB, 0.f, s));
      }
    }
  }
  ones_tile_kernel<<<8, 128, 0, s>>>(reinterpret_cast<__nv_bfloat16*>(base + P.ones_off));
  count_launch();
  const int e = static_cast<int>(cudaGetLastError());
  if (e != 0) return e;
  return batch.end();
}

extern "C" int rlsb_ac_actor_slots(const rlsb_ac_cfg* cfg, int64_t N, void* workspace, rlsb_actor_slots* slots) {
  if (!cfg || !workspace || !slots || N <= 0) return -1;
  AcPlan P;
  RLSB_TRY(make_ac_plan(*cfg, P));
  AcWorkspace W;
  RLSB_TRY(make_ac_workspace(P, N, W));
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  for (int l = 0; l < 4; ++l) {   // group 0 (actor) comes first in every [actor | critic] image
    slots->x[l] = ws + W.x[l];
    slots->pre[l] = ws + W.pre[l];
    slots->rstd[l] = reinterpret_cast<float*>(ws + W.rstd[l]);
  }
  slots->m_pad = W.m_pad;
  slots->Hp = P.Hp;
  slots->steps = P.H;
  return 0;
}

extern "C" int rlsb_ac_update(const rlsb_ac_cfg* cfg, const void* packed, int64_t N, const void* determ_packed,
                              const void* stoch_packed, const float* vs, const float* w, const float* values,
                              const float* actions, const float* g_actions, uint64_t seed,
                              const uint64_t* seed_device, const rlsb_mlp_grads* actor_grads,
                              const rlsb_mlp_grads* critic_grads, float* scalars, void* workspace, void* stream_) {
  if (!cfg || !packed || !determ_packed || !stoch_packed || !vs || !w || !values || !actions || !actor_grads ||
      !critic_grads || !scalars || !workspace || N <= 0)
    return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  AcPlan P;
  RLSB_TRY(make_ac_plan(*cfg, P));
  AcWorkspace W;
  RLSB_TRY(make_ac_workspace(P, N, W));
  if (W.M > (1LL << 30)) return -3;
  const int M = static_cast<int>(W.M);
  const int m_tiles = M / 128;
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  auto bfw = [&](size_t off) { return reinterpret_cast<__nv_bfloat16*>(ws + off); };
  auto pbf = [&](size_t off) { return reinterpret_cast<const __nv_bfloat16*>(pk + off); };
  auto pf = [&](size_t off) { return reinterpret_cast<const float*>(pk + off); };
  const bool ln = cfg->layer_norm != 0;
  const float eps = 1e-5f;
  const long long act_gs = static_cast<long long>(M) * P.Hp;   // group stride of activation images
  float* head_out = reinterpret_cast<float*>(ws + W.head_out);
  double* accum = reinterpret_cast<double*>(ws + W.accum);
  const __nv_bfloat16* himg = static_cast<const __nv_bfloat16*>(determ_packed);
  const __nv_bfloat16* zimg = static_cast<const __nv_bfloat16*>(stoch_packed);
  const __nv_bfloat16* ones = pbf(P.ones_off);

  // ---- forward: actor | critic on every state of steps 0..H-1 -----------------------------------
  // (the hidden layers of the actor — group 0 — are already in place when the rollout filled its slots: only the
  //  critic — group 1 — runs them; the small last layer runs for both)
  const bool actor_done = cfg->actor_fwd_in_rollout != 0;
  for (int l = 0; l < 5; ++l) {
    const AcLayer& L = P.L[l];
    const int g0 = (actor_done && l < 4) ? 1 : 0;   // first group of this launch
    GemmParams g{};
    g.W = pbf(L.w_off) + static_cast<size_t>(g0) * L.RB * L.kp; g.RB = L.RB; g.NB = 1; g.G = kG - g0;
    g.M = M; g.m_tiles = m_tiles; g.N = L.N;
    g.bias = pf(L.bias_off) + static_cast<size_t>(g0) * L.RB;
    g.ln_eps = eps;
    g.row_period = W.m_pad; g.row_valid = static_cast<int>(N);
    if (l == 0) {
      g.n_seg = 2;
      g.A[0] = himg; g.a_ktiles[0] = P.Dp / 64; g.a_group_stride[0] = 0;
      g.A[1] = zimg; g.a_ktiles[1] = P.Sp / 64; g.a_group_stride[1] = 0;
    } else {
      g.n_seg = 1;
      g.A[0] = bfw(W.x[l - 1]) + static_cast<size_t>(g0) * act_gs; g.a_ktiles[0] = P.Hp / 64; g.a_group_stride[0] = act_gs;
    }
    if (l < 4) {
      const bool has_ln = (l == 0) || ln;
      g.ln_gamma = has_ln ? pf(L.g_off) + static_cast<size_t>(g0) * L.RB : nullptr;
      g.ln_beta = has_ln ? pf(L.b_off) + static_cast<size_t>(g0) * L.RB : nullptr;
      g.act = ACT_ELU;
      g.out_bf16 = bfw(W.x[l]) + static_cast<size_t>(g0) * act_gs; g.out_kpad = P.Hp; g.out_bf16_group_stride = act_gs;
      g.save_pre = bfw(W.pre[l]) + static_cast<size_t>(g0) * act_gs;
      g.save_rstd = has_ln ? reinterpret_cast<float*>(ws + W.rstd[l]) + static_cast<size_t>(g0) * M : nullptr;
      RLSB_TRY(launch_gemm(g, EPI_LN_ACT, s));
    } else {
      g.out_f32 = head_out; g.ldo = 32; g.out_group_stride = static_cast<long long>(M) * 32;
      RLSB_TRY(launch_gemm(g, EPI_PLAIN, s));
    }
  }

  // ---- losses, metrics, d(loss)/d(head outputs) ---------------------------------------------------
  cudaError_t ce = cudaMemsetAsync(accum, 0, kScalars * sizeof(double), s);
  if (ce != cudaSuccess) return static_cast<int>(ce);
  // running min / max of the metric draws as order-preserving int keys: +3.4e38 / very negative sentinels
  ce = cudaMemsetAsync(accum + kAccMinKey, 0x7f, sizeof(double), s);
  if (ce != cudaSuccess) return static_cast<int>(ce);
  ce = cudaMemsetAsync(accum + kAccMaxKey, 0x80, sizeof(double), s);
  if (ce != cudaSuccess) return static_cast<int>(ce);
  {
    AcLossArgs a{};
    a.head_out = head_out; a.group_stride = static_cast<long long>(M) * 32;
    a.m_pad = W.m_pad; a.N = static_cast<int>(N); a.H = P.H; a.A = P.A;
    a.vs = vs; a.w = w; a.values = values; a.actions = actions; a.g_actions = g_actions;
    a.discrete = cfg->discrete;
    a.rho = cfg->rho; a.eta = cfg->eta; a.metrics_samples = cfg->metrics_samples; a.seed = seed; a.seed_ptr = seed_device;
    a.dy4 = bfw(W.dy4); a.accum = accum;
    ac_loss_kernel<<<static_cast<unsigned>((W.M + 127) / 128), 128, 0, s>>>(a);
    count_launch();
    ce = cudaGetLastError();
    if (ce != cudaSuccess) return static_cast<int>(ce);
    ac_finalize_kernel<<<1, 32, 0, s>>>(accum, P.H, N, P.A, cfg->metrics_samples, cfg->discrete, scalars);
    count_launch();
  }

  // ---- backward ---------------------------------------------------------------------------------------
  const __nv_bfloat16* dy = bfw(W.dy4);      // d loss / d (layer l pre-activation output), packed
  long long dy_gs = static_cast<long long>(M) * 64;
  int dy_tiles = 1;
  for (int l = 4; l >= 0; --l) {
    const AcLayer& L = P.L[l];
    // (1) weight + bias gradient of layer l:  dW_l = dy^T x_{l-1},  db_l = dy^T 1
    WgradParams wp{};
    wp.dY = dy; wp.dy_group_stride = dy_gs; wp.n_tiles = dy_tiles;
    wp.G = kG; wp.m_tiles = m_tiles;
    int ones_col;
    if (l == 0) {
      wp.n_seg = 3;
      wp.X[0] = himg; wp.x_ktiles[0] = P.Dp / 64; wp.x_group_stride[0] = 0; wp.x_mtile_stride[0] = static_cast<long long>(P.Dp) * 128;
      wp.X[1] = zimg; wp.x_ktiles[1] = P.Sp / 64; wp.x_group_stride[1] = 0; wp.x_mtile_stride[1] = static_cast<long long>(P.Sp) * 128;
      wp.X[2] = ones; wp.x_ktiles[2] = 1; wp.x_group_stride[2] = 0; wp.x_mtile_stride[2] = 0;
      ones_col = P.Dp + P.Sp;
    } else {
      wp.n_seg = 2;
      wp.X[0] = bfw(W.x[l - 1]); wp.x_ktiles[0] = P.Hp / 64; wp.x_group_stride[0] = act_gs;
      wp.x_mtile_stride[0] = static_cast<long long>(P.Hp) * 128;
      wp.X[1] = ones; wp.x_ktiles[1] = 1; wp.x_group_stride[1] = 0; wp.x_mtile_stride[1] = 0;
      ones_col = P.Hp;
    }
    wp.partial = reinterpret_cast<float*>(ws + W.partial);
    RLSB_TRY(plan_wgrad(wp));
    RLSB_TRY(launch_wgrad(wp, s));
    WgradReduceParams rp{};
    rp.partial = wp.partial; rp.splits = wp.splits; rp.G = kG; rp.rows_pad = wp.n_slices * 128; rp.ld = wp.kt_total * 64;
    for (int g = 0; g < kG; ++g) {
      const rlsb_mlp_grads* gr = g == 0 ? actor_grads : critic_grads;
      rp.w_dst[g] = gr->w[l]; rp.b_dst[g] = gr->b[l];
      rp.n_out[g] = (l == 4) ? (g == 0 ? P.Aout : 1) : P.Hd;
    }
    if (l == 0) {
      rp.ld_dst = P.D + P.S; rp.n_seg = 2;
      rp.seg[0] = PackSeg{0, 0, P.D};
      rp.seg[1] = PackSeg{P.Dp, P.D, P.S};
    } else {
      rp.ld_dst = P.Hd; rp.n_seg = 1;
      rp.seg[0] = PackSeg{0, 0, P.Hd};
    }
    rp.ones_col = ones_col;
    RLSB_TRY(launch_wgrad_reduce(rp, s));
    if (l == 0) break;   // zs is detached (dreamer_v2.py:199-206): no gradient flows into the states

    // (2) d loss / d x_{l-1} = dy W_l, then ELU' and LayerNorm backward of layer l-1 in the epilogue
    const AcLayer& Lp = P.L[l - 1];
    const bool has_ln = (l - 1 == 0) || ln;
    GemmParams g{};
    g.n_seg = 1;
    g.A[0] = dy; g.a_ktiles[0] = L.t_kp / 64; g.a_group_stride[0] = dy_gs;
    g.W = pbf(L.wt_off); g.RB = L.t_RB; g.NB = 1; g.G = kG;
    g.M = M; g.m_tiles = m_tiles; g.N = P.Hd;
    g.row_period = W.m_pad; g.row_valid = static_cast<int>(N);
    g.ln_gamma = has_ln ? pf(Lp.g_off) : nullptr;
    g.ln_beta = has_ln ? pf(Lp.b_off) : nullptr;
    g.ln_eps = eps;
    g.act = ACT_ELU;
    g.bwd_pre = bfw(W.pre[l - 1]);
    g.bwd_rstd = has_ln ? reinterpret_cast<const float*>(ws + W.rstd[l - 1]) : nullptr;
    g.out_bf16 = bfw(W.dp[l & 1]); g.out_kpad = P.Hp; g.out_bf16_group_stride = act_gs;
    g.col_part = has_ln ? reinterpret_cast<float*>(ws + W.col_part) : nullptr;
    g.group_major = 1;
    RLSB_TRY(launch_gemm(g, EPI_BWD, s));
    if (has_ln) {
      ColsumReduceParams cp{};
      cp.col_part = g.col_part; cp.ctas = gemm_grid_size(g); cp.G = kG; cp.RB = g.RB; cp.N = P.Hd;
      cp.dgamma[0] = actor_grads->ln_g[l - 1]; cp.dbeta[0] = actor_grads->ln_b[l - 1];
      cp.dgamma[1] = critic_grads->ln_g[l - 1]; cp.dbeta[1] = critic_grads->ln_b[l - 1];
      RLSB_TRY(launch_colsum_reduce(cp, s));
    }
    dy = bfw(W.dp[l & 1]);
    dy_gs = act_gs;
    dy_tiles = P.Hp / 64;
  }
  return 0;
}

// Loss / metric scalars from caller-supplied head outputs: the loss kernel of rlsb_ac_update on its own.  Used with the
// head outputs of a rollout in the split-operand contraction mode (rlsb_imagine_cfg::parity) it evaluates
// ImaginativeCritic.calculate_loss / ImaginativeActor.calculate_loss (ac.py:68-81,113-146) at fp32-grade precision.
extern "C" int rlsb_ac_losses(const rlsb_ac_cfg* cfg, int64_t N, const float* head_out, const float* vs, const float* w,
                              const float* values, const float* actions, uint64_t seed, float* scalars, void* workspace,
                              void* stream_) {
  if (!cfg || !head_out || !vs || !w || !values || !actions || !scalars || !workspace || N <= 0) return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  AcPlan P;
  RLSB_TRY(make_ac_plan(*cfg, P));
  AcWorkspace W;
  RLSB_TRY(make_ac_workspace(P, N, W));
  if (W.M > (1LL << 30)) return -3;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  double* accum = reinterpret_cast<double*>(ws + W.accum);
  cudaError_t ce = cudaMemsetAsync(accum, 0, kScalars * sizeof(double), s);
  if (ce != cudaSuccess) return static_cast<int>(ce);
  ce = cudaMemsetAsync(accum + kAccMinKey, 0x7f, sizeof(double), s);
  if (ce != cudaSuccess) return static_cast<int>(ce);
  ce = cudaMemsetAsync(accum + kAccMaxKey, 0x80, sizeof(double), s);
  if (ce != cudaSuccess) return static_cast<int>(ce);
  AcLossArgs a{};
  a.head_out = head_out; a.group_stride = W.M * 32;
  a.m_pad = W.m_pad; a.N = static_cast<int>(N); a.H = P.H; a.A = P.A;
  a.vs = vs; a.w = w; a.values = values; a.actions = actions; a.g_actions = nullptr;
  a.discrete = cfg->discrete;
  a.rho = cfg->rho; a.eta = cfg->eta; a.metrics_samples = cfg->metrics_samples; a.seed = seed; a.seed_ptr = nullptr;
  a.dy4 = reinterpret_cast<__nv_bfloat16*>(ws + W.dy4); a.accum = accum;
  ac_loss_kernel<<<static_cast<unsigned>((W.M + 127) / 128), 128, 0, s>>>(a);
  count_launch();
  ce = cudaGetLastError();
  if (ce != cudaSuccess) return static_cast<int>(ce);
  ac_finalize_kernel<<<1, 32, 0, s>>>(accum, P.H, N, P.A, cfg->metrics_samples, cfg->discrete, scalars);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}
