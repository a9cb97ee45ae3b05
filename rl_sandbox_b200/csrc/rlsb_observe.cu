// rlsb_observe.cu — K5: the world-model observe scan, forward and backward.
//
// Replaces the T = 50 sequential RSSM.forward calls of WorldModel.calculate_loss
// (agents/dreamer/world_model.py:187-202 -> agents/dreamer/rssm.py:176-209) and their autograd:
//   x_t      = ELU(LN?(W_in [z_{t-1}, a_t]))                 rssm.py:179
//   h_t      = GRU(x_t, h_{t-1})                             common.py:69-81
//   prior_t  = W_p2 ELU(LN?(W_p1 h_t))                       rssm.py:192   (KL target / regulariser)
//   post_t   = W_q2 ELU(LN?(W_q1 [h_t, embed_t]))            rssm.py:195-199 (stoch_net)
//   z_t      ~ OneHotCategoricalST(post_t)                   rssm.py:34-37
// Only B = 16 sequences advance per step, so every kernel is latency bound; what the scan buys is
// ~9 launches per step instead of ~40 torch ops forward and a hand-rolled BPTT (≈ 14 launches per
// step, no autograd graph) backward.  Weight / bias gradients are formed ONCE at the end by the
// tcgen05 weight-gradient contraction over the images of all T steps stacked along the row axis
// (rlsb_wgrad.cu); LayerNorm gamma / beta gradients by fixed-order column reductions.
//
// Row layout: step t owns rows [t * m_pad, t * m_pad + B) of every stacked image (m_pad = B rounded
// up to the 128-row tile); padding rows are zero in every operand that enters a weight gradient.
#include <cmath>

#include "../../include/rlsb.h"
#include "rlsb_count.cuh"
#include "rlsb_gemm.cuh"
#include "rlsb_imagine_plan.cuh"
#include "rlsb_kernels.cuh"
#include "rlsb_ptx.cuh"
#include "rlsb_wgrad.cuh"

namespace rlsb {

using k1::LayerPlan;
using k1::TLayer;
using k1::place;
using k1::plan_nb;
using k1::ru;
using k1::rus;

namespace {

struct ObsPlan {
  int D, S, A, E, Dp, Sp, Ap, Ep, G3p, T;
  bool ln;
  LayerPlan img_in, gru, prior1, prior2, post1, post2;
  TLayer t_img_in, t_gru_x, t_gru_h, t_prior1, t_prior2, t_post1, t_post2;
  size_t ones_off;
  size_t packed_bytes;
};

int make_obs_plan(const rlsb_observe_cfg& c, ObsPlan& P) {
  if (c.classes != 32 || c.groups <= 0 || c.groups > 64 || c.D <= 0 || (c.D & 7) || c.A <= 0 || c.E <= 0 || c.T <= 0)
    return -10;
  P.D = c.D; P.S = c.groups * c.classes; P.A = c.A; P.E = c.E; P.T = c.T;
  P.Dp = ru(P.D, 64); P.Sp = ru(P.S, 64); P.Ap = ru(P.A, 64); P.Ep = ru(P.E, 64); P.G3p = ru(3 * P.D, 64);
  P.ln = c.layer_norm != 0;
  if (P.Ap > 64) return -12;
  size_t cur = 0;
  auto finish = [&](LayerPlan& L, int n, int kp, int ln_len) {
    L.N = n; L.kp = kp; L.G = 1;
    plan_nb(L);
    L.w_off = place(cur, static_cast<size_t>(L.NB) * L.RB * L.kp * 2);
    L.bias_off = place(cur, static_cast<size_t>(L.NB) * L.RB * 4);
    L.g_off = place(cur, static_cast<size_t>(ln_len) * 4);
    L.b_off = place(cur, static_cast<size_t>(ln_len) * 4);
  };
  finish(P.img_in, P.D, P.Sp + P.Ap, P.D);
  finish(P.gru, 3 * P.D, 2 * P.Dp, 3 * P.D);
  finish(P.prior1, P.D, P.Dp, P.D);
  finish(P.prior2, P.S, P.Dp, 32);
  finish(P.post1, P.D, P.Dp + P.Ep, P.D);
  finish(P.post2, P.S, P.Dp, 32);
  auto tplace = [&](TLayer& T, int rows, int kp) {
    LayerPlan tmp;
    tmp.N = rows;
    plan_nb(tmp);
    T.RB = tmp.RB; T.NB = tmp.NB; T.kp = kp;
    T.off = place(cur, static_cast<size_t>(T.NB) * T.RB * kp * 2);
  };
  tplace(P.t_img_in, P.Sp + P.Ap, P.Dp);
  tplace(P.t_gru_x, P.D, P.G3p);
  tplace(P.t_gru_h, P.D, P.G3p);
  tplace(P.t_prior1, P.D, P.Dp);
  tplace(P.t_prior2, P.D, P.Sp);
  tplace(P.t_post1, P.Dp + P.Ep, P.Dp);
  tplace(P.t_post2, P.D, P.Sp);
  P.ones_off = place(cur, 128 * 64 * 2);
  P.packed_bytes = rus(cur, 1024);
  return 0;
}

// everything the backward pass needs, stacked over the T steps (zero-initialised by the caller)
struct ObsTape {
  size_t z_img, h_img;                       // (T+1) slots: slot t = state entering step t
  size_t a_img, e_img, x_img, y_img, y2_img; // T slots, packed bf16
  size_t sc_x, st_x, sc_g, st_g, sc_y, st_y, sc_y2, st_y2;   // fp32 pre-activations + LayerNorm partial statistics
  int m_pad;
  long long ld3;                             // leading dimension of the GRU pre-activations
  size_t bytes;
};

void make_obs_tape(const ObsPlan& P, long long B, ObsTape& T) {
  const size_t m = static_cast<size_t>(ru(static_cast<int>(B), 128));
  const size_t n = static_cast<size_t>(P.T);
  T.m_pad = static_cast<int>(m);
  T.ld3 = ru(3 * P.D, 4);
  size_t cur = 0;
  T.z_img = place(cur, (n + 1) * m * P.Sp * 2);
  T.h_img = place(cur, (n + 1) * m * P.Dp * 2);
  T.a_img = place(cur, n * m * P.Ap * 2);
  T.e_img = place(cur, n * m * P.Ep * 2);
  T.x_img = place(cur, n * m * P.Dp * 2);
  T.y_img = place(cur, n * m * P.Dp * 2);
  T.y2_img = place(cur, n * m * P.Dp * 2);
  T.sc_x = place(cur, n * m * P.D * 4);
  T.st_x = place(cur, n * P.img_in.NB * m * 2 * 4);
  T.sc_g = place(cur, n * m * T.ld3 * 4);
  T.st_g = place(cur, n * P.gru.NB * m * 2 * 4);
  T.sc_y = place(cur, n * m * P.D * 4);
  T.st_y = place(cur, n * P.prior1.NB * m * 2 * 4);
  T.sc_y2 = place(cur, n * m * P.D * 4);
  T.st_y2 = place(cur, n * P.post1.NB * m * 2 * 4);
  T.bytes = rus(cur, 1024);
}

struct ObsBwdWs {
  // stacked dY images for the weight gradients (T slots, packed bf16, padding rows zero)
  size_t dp_in, g_pre, dp1, gl_prior, dp2, gl_post;
  // stacked d loss / d (LayerNorm affine output) for the gamma / beta reductions (T slots, fp32)
  size_t da_x, da_g, da_y, da_y2;
  // per-step scratch
  size_t g_y2, g_he, g_y, g_hprior, g_x, g_hgru, g_hdirect, g_za;
  size_t partial;
  long long ld_he, ld_za;
  size_t bytes;
};

int make_obs_bwd_ws(const ObsPlan& P, const ObsTape& T, ObsBwdWs& W) {
  const size_t m = static_cast<size_t>(T.m_pad), n = static_cast<size_t>(P.T);
  size_t cur = 0;
  W.dp_in = place(cur, n * m * P.Dp * 2);
  W.g_pre = place(cur, n * m * P.G3p * 2);
  W.dp1 = place(cur, n * m * P.Dp * 2);
  W.gl_prior = place(cur, n * m * P.Sp * 2);
  W.dp2 = place(cur, n * m * P.Dp * 2);
  W.gl_post = place(cur, n * m * P.Sp * 2);
  W.da_x = place(cur, n * m * P.D * 4);
  W.da_g = place(cur, n * m * T.ld3 * 4);
  W.da_y = place(cur, n * m * P.D * 4);
  W.da_y2 = place(cur, n * m * P.D * 4);
  W.ld_he = P.Dp + P.Ep;
  W.ld_za = P.Sp + P.Ap;
  W.g_y2 = place(cur, m * P.D * 4);
  W.g_he = place(cur, m * W.ld_he * 4);
  W.g_y = place(cur, m * P.D * 4);
  W.g_hprior = place(cur, m * P.D * 4);
  W.g_x = place(cur, m * P.D * 4);
  W.g_hgru = place(cur, m * P.D * 4);
  W.g_hdirect = place(cur, m * P.D * 4);
  W.g_za = place(cur, m * W.ld_za * 4);
  size_t pmax = 0;
  const int m_tiles = static_cast<int>(n * m / 128);
  const int shapes[6][2] = {{P.Dp, (P.Sp + P.Ap)}, {P.G3p, 2 * P.Dp}, {P.Dp, P.Dp}, {P.Sp, P.Dp}, {P.Dp, P.Dp + P.Ep},
                            {P.Sp, P.Dp}};
  for (int i = 0; i < 6; ++i) {
    WgradParams wp{};
    wp.n_tiles = shapes[i][0] / 64; wp.G = 1; wp.m_tiles = m_tiles;
    wp.n_seg = 2; wp.x_ktiles[0] = shapes[i][1] / 64; wp.x_ktiles[1] = 1;
    const int e = plan_wgrad(wp);
    if (e != 0) return e;
    const size_t b = wgrad_partial_bytes(wp);
    if (b > pmax) pmax = b;
  }
  W.partial = place(cur, pmax);
  W.bytes = rus(cur, 1024);
  return 0;
}

__device__ __forceinline__ uint32_t bfpair(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float fsig(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float ftanh(float x) {
  const float e = __expf(-2.0f * fabsf(x));
  return copysignf((1.0f - e) / (1.0f + e), x);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void row_stats(const float* stats, int NB, int m_pad, int m, int N, float eps, int lane,
                                          float& mean, float& rstd) {
  const float2* st = reinterpret_cast<const float2*>(stats);
  float s = 0.f, q = 0.f;
  for (int b = lane; b < NB; b += 32) {
    const float2 v = __ldg(&st[static_cast<size_t>(b) * m_pad + m]);
    s += v.x;
    q += v.y;
  }
  s = warp_sum(s);
  q = warp_sum(q);
  const float inv_n = 1.0f / static_cast<float>(N);
  mean = s * inv_n;
  rstd = 1.0f / sqrtf(fmaxf(q * inv_n - mean * mean, 0.f) + eps);
}

// (T, B, cols) fp32 -> T stacked packed images [m_pad x k_pad] (padding rows / columns zero)
__global__ void pack_steps_kernel(const float* __restrict__ src, int T, int B, int cols, __nv_bfloat16* __restrict__ dst,
                                  int m_pad, int k_pad) {
  const int chunks = k_pad >> 3;
  const long long total = static_cast<long long>(T) * m_pad * chunks;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % chunks);
    const int r = static_cast<int>((i / chunks) % m_pad);
    const int t = static_cast<int>(i / (static_cast<long long>(chunks) * m_pad));
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = ch * 8 + j;
      v[j] = (r < B && c < cols) ? __ldg(src + (static_cast<size_t>(t) * B + r) * cols + c) : 0.f;
    }
    *reinterpret_cast<uint4*>(dst + static_cast<size_t>(t) * m_pad * k_pad +
                              packed_index(static_cast<size_t>(r), static_cast<size_t>(ch) * 8, static_cast<size_t>(k_pad), kTileM)) =
        make_uint4(bfpair(v[0], v[1]), bfpair(v[2], v[3]), bfpair(v[4], v[5]), bfpair(v[6], v[7]));
  }
}

// z = onehot + p - p.detach():  g_logit = g_ext + p * (g_z - <g_z, p>) per group of 32 classes; packed bf16 out.
// g_z = gz_a (+ gz_b); any input may be nullptr.  One thread per (row of the padded image, group).
__global__ void st_softmax_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ g_ext,
                                      const float* __restrict__ gz_a, const float* __restrict__ gz_b, long long ld_b,
                                      int M, int m_pad, int groups, __nv_bfloat16* __restrict__ out, int kpad) {
  pdl_launch_dependents();
  pdl_wait();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(m_pad) * groups) return;
  const int m = static_cast<int>(i / groups);
  const int g = static_cast<int>(i - static_cast<long long>(m) * groups);
  const int S = groups * 32;
  float o[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) o[k] = 0.f;
  if (m < M) {
    const size_t base = static_cast<size_t>(m) * S + g * 32;
    if (g_ext) {
#pragma unroll
      for (int k = 0; k < 32; ++k) o[k] = g_ext[base + k];
    }
    if (gz_a || gz_b) {
      float p[32], gz[32];
      float mx = -3.0e38f;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        p[k] = logits[base + k];
        gz[k] = (gz_a ? gz_a[base + k] : 0.f) + (gz_b ? gz_b[static_cast<size_t>(m) * ld_b + g * 32 + k] : 0.f);
        mx = fmaxf(mx, p[k]);
      }
      float se = 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        p[k] = __expf(p[k] - mx);
        se += p[k];
      }
      const float inv = 1.0f / se;
      float dot = 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        p[k] *= inv;
        dot = fmaf(gz[k], p[k], dot);
      }
#pragma unroll
      for (int k = 0; k < 32; ++k) o[k] += p[k] * (gz[k] - dot);
    }
  }
#pragma unroll
  for (int ch = 0; ch < 4; ++ch)
    *reinterpret_cast<uint4*>(out + packed_index(static_cast<size_t>(m), static_cast<size_t>(g * 32 + ch * 8),
                                                 static_cast<size_t>(kpad), kTileM)) =
        make_uint4(bfpair(o[ch * 8], o[ch * 8 + 1]), bfpair(o[ch * 8 + 2], o[ch * 8 + 3]),
                   bfpair(o[ch * 8 + 4], o[ch * 8 + 5]), bfpair(o[ch * 8 + 6], o[ch * 8 + 7]));
}

// y = ELU(LN?(p)):  given g = d loss / d y (fp32) and the saved pre-activation p (+ LayerNorm partial statistics):
//   da = g * ELU'(a), a = gamma * x_hat + beta (or a = p);  d p = LayerNorm backward of da * gamma (or da)
// writes d p as a packed bf16 image (padding rows zero) and da (fp32) for the gamma / beta reduction.  Warp per row.
__global__ void ln_act_bwd_kernel(const float* __restrict__ g, long long ld_g, const float* __restrict__ pre,
                                  const float* __restrict__ stats, int NB, int M, int m_pad, int D,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                  __nv_bfloat16* __restrict__ dp, int kpad, float* __restrict__ da_out) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long row = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  if (row >= m_pad) return;
  const int m = static_cast<int>(row);
  const int chunks = D >> 3, kchunks = kpad >> 3;
  if (m >= M) {
    for (int c = lane; c < kchunks; c += 32)
      *reinterpret_cast<uint4*>(dp + packed_index(static_cast<size_t>(m), static_cast<size_t>(c) * 8,
                                                  static_cast<size_t>(kpad), kTileM)) = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  float mean = 0.f, rstd = 1.f;
  if (gamma) row_stats(stats, NB, m_pad, m, D, eps, lane, mean, rstd);
  const float* p = pre + static_cast<size_t>(m) * D;
  const float* gr = g + static_cast<size_t>(m) * ld_g;
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < chunks; c += 32) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = c * 8 + j;
      const float xh = (p[k] - mean) * rstd;
      const float a = gamma ? fmaf(xh, __ldg(gamma + k), __ldg(beta + k)) : p[k];
      const float da = gr[k] * (a > 0.f ? 1.0f : __expf(a));
      da_out[static_cast<size_t>(m) * D + k] = da;
      if (gamma) {
        const float dxh = da * __ldg(gamma + k);
        s1 += dxh;
        s2 = fmaf(dxh, xh, s2);
      }
    }
  }
  float m1 = 0.f, m2 = 0.f;
  if (gamma) {
    m1 = warp_sum(s1) / static_cast<float>(D);
    m2 = warp_sum(s2) / static_cast<float>(D);
  }
  for (int c = lane; c < kchunks; c += 32) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = c * 8 + j;
      float v = 0.f;
      if (k < D) {
        const float da = da_out[static_cast<size_t>(m) * D + k];
        if (gamma) {
          const float xh = (p[k] - mean) * rstd;
          v = rstd * (da * __ldg(gamma + k) - m1 - xh * m2);
        } else {
          v = da;
        }
      }
      o[j] = v;
    }
    *reinterpret_cast<uint4*>(dp + packed_index(static_cast<size_t>(m), static_cast<size_t>(c) * 8,
                                                static_cast<size_t>(kpad), kTileM)) =
        make_uint4(bfpair(o[0], o[1]), bfpair(o[2], o[3]), bfpair(o[4], o[5]), bfpair(o[6], o[7]));
  }
}

struct GruBwdArgs {
  const float* scratch;
  long long ld;
  const float* stats;
  int NB, M, m_pad, D;
  const float* gamma;
  const float* beta;
  float eps, update_bias;
  const float* h_prev;      // (M, D) fp32 or nullptr (= zeros: the initial state)
  const float* gh[5];
  long long ld_gh[5];
  int n_gh;
  __nv_bfloat16* g_pre;     // packed [m_pad x kpad]
  int kpad;
  float* g_hdirect;         // (M, D)
  float* da_out;            // (M, ld): d loss / d (gamma * x_hat + beta) for the three gates
};

__global__ void gru_gate_bwd_kernel(const GruBwdArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  if (warp >= a.m_pad) return;
  const int m = static_cast<int>(warp);
  const int D = a.D;
  const int chunks = D >> 3;
  if (m >= a.M) {
    for (int c = lane; c < (a.kpad >> 3); c += 32)
      *reinterpret_cast<uint4*>(a.g_pre + packed_index(static_cast<size_t>(m), static_cast<size_t>(c) * 8,
                                                       static_cast<size_t>(a.kpad), kTileM)) = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  float mean, rstd;
  row_stats(a.stats, a.NB, a.m_pad, m, 3 * D, a.eps, lane, mean, rstd);
  const float* src = a.scratch + static_cast<size_t>(m) * a.ld;
  float* da = a.da_out + static_cast<size_t>(m) * a.ld;
  float s1 = 0.f, s2 = 0.f;
  for (int ch = lane; ch < chunks; ch += 32) {
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int j = ch * 8 + jj;
      const float xr = (src[j] - mean) * rstd, xc = (src[D + j] - mean) * rstd, xu = (src[2 * D + j] - mean) * rstd;
      const float gr_ = __ldg(a.gamma + j), gc_ = __ldg(a.gamma + D + j), gu_ = __ldg(a.gamma + 2 * D + j);
      const float nr = fmaf(xr, gr_, __ldg(a.beta + j));
      const float nc = fmaf(xc, gc_, __ldg(a.beta + D + j));
      const float nu = fmaf(xu, gu_, __ldg(a.beta + 2 * D + j)) + a.update_bias;
      const float r = fsig(nr), c = ftanh(r * nc), u = fsig(nu);
      float gh = 0.f;
      for (int i = 0; i < a.n_gh; ++i) gh += a.gh[i][static_cast<size_t>(m) * a.ld_gh[i] + j];
      const float hp = a.h_prev ? a.h_prev[static_cast<size_t>(m) * D + j] : 0.f;
      const float g_t = gh * u * (1.0f - c * c);
      const float dar = g_t * nc * r * (1.0f - r);     // d / d (gamma x_hat + beta), reset gate
      const float dac = g_t * r;
      const float dau = gh * (c - hp) * u * (1.0f - u);
      da[j] = dar; da[D + j] = dac; da[2 * D + j] = dau;
      a.g_hdirect[static_cast<size_t>(m) * D + j] = gh * (1.0f - u);
      s1 += dar * gr_ + dac * gc_ + dau * gu_;
      s2 += dar * gr_ * xr + dac * gc_ * xc + dau * gu_ * xu;
    }
  }
  const float inv_n = 1.0f / static_cast<float>(3 * D);
  const float m1 = warp_sum(s1) * inv_n, m2 = warp_sum(s2) * inv_n;
  for (int c = lane; c < (a.kpad >> 3); c += 32) {
    float o[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int k = c * 8 + jj;
      float v = 0.f;
      if (k < 3 * D) {
        const float xh = (src[k] - mean) * rstd;
        v = rstd * (da[k] * __ldg(a.gamma + k) - m1 - xh * m2);
      }
      o[jj] = v;
    }
    *reinterpret_cast<uint4*>(a.g_pre + packed_index(static_cast<size_t>(m), static_cast<size_t>(c) * 8,
                                                     static_cast<size_t>(a.kpad), kTileM)) =
        make_uint4(bfpair(o[0], o[1]), bfpair(o[2], o[3]), bfpair(o[4], o[5]), bfpair(o[6], o[7]));
  }
}

// d gamma[k] = sum over (t, b) of da * x_hat, d beta[k] = sum of da; x_hat recomputed from the saved pre-activations
// and partial statistics.  One thread per column, rows in a fixed order (deterministic).
__global__ void ln_param_grad_kernel(const float* __restrict__ da, const float* __restrict__ pre, long long ld,
                                     const float* __restrict__ stats, int NB, int T, int B, int m_pad, int N,
                                     float eps, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  float ag = 0.f, ab = 0.f;
  const float inv_n = 1.0f / static_cast<float>(N);
  for (int t = 0; t < T; ++t) {
    const float2* st = reinterpret_cast<const float2*>(stats) + static_cast<size_t>(t) * NB * m_pad;
    for (int b = 0; b < B; ++b) {
      float s = 0.f, q = 0.f;
      for (int nb = 0; nb < NB; ++nb) {
        const float2 v = __ldg(&st[static_cast<size_t>(nb) * m_pad + b]);
        s += v.x;
        q += v.y;
      }
      const float mean = s * inv_n;
      const float rstd = 1.0f / sqrtf(fmaxf(q * inv_n - mean * mean, 0.f) + eps);
      const size_t off = (static_cast<size_t>(t) * m_pad + b) * ld + k;
      const float d = da[off];
      ag = fmaf(d, (pre[off] - mean) * rstd, ag);
      ab += d;
    }
  }
  dgamma[k] = ag;
  dbeta[k] = ab;
}

__global__ void extract_steps_kernel(const float* __restrict__ src, long long ld, int col0, int B, int n,
                                     float* __restrict__ dst) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * n) return;
  const int m = i / n, k = i % n;
  dst[i] = src[static_cast<size_t>(m) * ld + col0 + k];
}

__global__ void obs_ones_tile_kernel(__nv_bfloat16* dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128 * 8) return;
  const int row = i >> 3, pos = i & 7;
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if ((pos ^ (row & 7)) == 0) v.x = 0x00003F80u;
  reinterpret_cast<uint4*>(dst)[i] = v;
}


#define RLSB_TRY(expr)            \
  do {                            \
    int _e = (expr);              \
    if (_e != 0) return _e;       \
  } while (0)
#define RLSB_CUDA_OK()                                      \
  do {                                                      \
    cudaError_t _ce = cudaGetLastError();                   \
    if (_ce != cudaSuccess) return static_cast<int>(_ce);   \
  } while (0)

int copy_pad(const float* src, int n, float* dst, int n_pad, float fill, cudaStream_t s) {
  return launch_copy_pad(src, n, dst, n_pad, fill, s);
}

}  // namespace
}  // namespace rlsb

using namespace rlsb;

extern "C" size_t rlsb_observe_packed_bytes(const rlsb_observe_cfg* cfg) {
  ObsPlan P;
  if (!cfg || make_obs_plan(*cfg, P) != 0) return 0;
  return P.packed_bytes;
}
extern "C" size_t rlsb_observe_tape_bytes(const rlsb_observe_cfg* cfg, int64_t B) {
  ObsPlan P;
  if (!cfg || B <= 0 || make_obs_plan(*cfg, P) != 0) return 0;
  ObsTape T;
  make_obs_tape(P, B, T);
  return T.bytes;
}
extern "C" size_t rlsb_observe_bwd_workspace_bytes(const rlsb_observe_cfg* cfg, int64_t B) {
  ObsPlan P;
  if (!cfg || B <= 0 || make_obs_plan(*cfg, P) != 0) return 0;
  ObsTape T;
  make_obs_tape(P, B, T);
  ObsBwdWs W;
  if (make_obs_bwd_ws(P, T, W) != 0) return 0;
  return W.bytes;
}

extern "C" int rlsb_observe_pack(const rlsb_observe_cfg* cfg, const rlsb_observe_params* p, void* packed, void* stream_) {
  if (!cfg || !p || !packed) return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  ObsPlan P;
  RLSB_TRY(make_obs_plan(*cfg, P));
  LaunchBatchScope batch(s);   // the small pack / pad launches below are queued and issued as multi-job kernels
  uint8_t* base = static_cast<uint8_t*>(packed);
  auto wptr = [&](size_t off) { return reinterpret_cast<__nv_bfloat16*>(base + off); };
  auto fptr = [&](size_t off) { return reinterpret_cast<float*>(base + off); };
  auto fwd = [&](const LayerPlan& L, const float* w, const float* b, int ld, int n_seg, const PackSeg* segs,
                 const float* g, const float* be, int ln_len) -> int {
    if (!w) return -20;
    RLSB_TRY(launch_pack(w, ld, L.N, wptr(L.w_off), L.RB, L.NB * L.RB, L.kp, n_seg, segs, s));
    RLSB_TRY(copy_pad(b, L.N, fptr(L.bias_off), L.NB * L.RB, 0.f, s));
    if (g) {
      RLSB_TRY(copy_pad(g, ln_len, fptr(L.g_off), ln_len, 1.f, s));
      RLSB_TRY(copy_pad(be, ln_len, fptr(L.b_off), ln_len, 0.f, s));
    }
    return 0;
  };
  auto tr = [&](const TLayer& T, const float* w, int ld, int cols, int n_seg, const PackSeg* segs) -> int {
    return launch_pack_transposed_seg(w, ld, cols, wptr(T.off), T.RB, T.NB * T.RB, T.kp, 0, T.kp, n_seg, segs, s);
  };
  if (P.ln && (!p->img_in_ln_g || !p->prior1_ln_g || !p->post1_ln_g)) return -21;
  {
    PackSeg segs[2] = {{0, 0, P.S}, {P.Sp, P.S, P.A}};
    RLSB_TRY(fwd(P.img_in, p->img_in_w, p->img_in_b, P.S + P.A, 2, segs, P.ln ? p->img_in_ln_g : nullptr, p->img_in_ln_b, P.D));
    RLSB_TRY(tr(P.t_img_in, p->img_in_w, P.S + P.A, P.D, 2, segs));
  }
  {
    PackSeg segs[2] = {{0, 0, P.D}, {P.Dp, P.D, P.D}};
    RLSB_TRY(fwd(P.gru, p->gru_w, p->gru_b, 2 * P.D, 2, segs, p->gru_ln_g, p->gru_ln_b, 3 * P.D));
    PackSeg sx[1] = {{0, 0, P.D}}, sh[1] = {{0, P.D, P.D}};
    RLSB_TRY(tr(P.t_gru_x, p->gru_w, 2 * P.D, 3 * P.D, 1, sx));
    RLSB_TRY(tr(P.t_gru_h, p->gru_w, 2 * P.D, 3 * P.D, 1, sh));
  }
  {
    PackSeg seg[1] = {{0, 0, P.D}};
    RLSB_TRY(fwd(P.prior1, p->prior1_w, p->prior1_b, P.D, 1, seg, P.ln ? p->prior1_ln_g : nullptr, p->prior1_ln_b, P.D));
    RLSB_TRY(tr(P.t_prior1, p->prior1_w, P.D, P.D, 1, seg));
    RLSB_TRY(fwd(P.prior2, p->prior2_w, p->prior2_b, P.D, 1, seg, nullptr, nullptr, 0));
    RLSB_TRY(tr(P.t_prior2, p->prior2_w, P.D, P.S, 1, seg));
    RLSB_TRY(fwd(P.post2, p->post2_w, p->post2_b, P.D, 1, seg, nullptr, nullptr, 0));
    RLSB_TRY(tr(P.t_post2, p->post2_w, P.D, P.S, 1, seg));
  }
  {
    PackSeg segs[2] = {{0, 0, P.D}, {P.Dp, P.D, P.E}};
    RLSB_TRY(fwd(P.post1, p->post1_w, p->post1_b, P.D + P.E, 2, segs, P.ln ? p->post1_ln_g : nullptr, p->post1_ln_b, P.D));
    RLSB_TRY(tr(P.t_post1, p->post1_w, P.D + P.E, P.D, 2, segs));
  }
  obs_ones_tile_kernel<<<8, 128, 0, s>>>(wptr(P.ones_off));
  count_launch();
  RLSB_CUDA_OK();
  return batch.end();
}

extern "C" int rlsb_observe_fwd(const rlsb_observe_cfg* cfg, const void* packed, int64_t B_, const float* embed,
                                const float* actions, const rlsb_noise* noise, const rlsb_observe_out* out, void* tape_,
                                void* stream_) {
  if (!cfg || !packed || !embed || !actions || !noise || !out || !tape_ || B_ <= 0 || B_ > (1 << 20)) return -1;
  if (!out->prior_logits || !out->post_logits || !out->determ || !out->stoch_idx) return -2;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  ObsPlan P;
  RLSB_TRY(make_obs_plan(*cfg, P));
  ObsTape TP;
  make_obs_tape(P, B_, TP);
  const int B = static_cast<int>(B_), m_pad = TP.m_pad, m_tiles = m_pad / 128, T = P.T;
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  uint8_t* tape = static_cast<uint8_t*>(tape_);
  auto pbf = [&](size_t off) { return reinterpret_cast<const __nv_bfloat16*>(pk + off); };
  auto pf = [&](size_t off) { return reinterpret_cast<const float*>(pk + off); };
  auto img = [&](size_t off, int t, int kp) { return reinterpret_cast<__nv_bfloat16*>(tape + off) + static_cast<size_t>(t) * m_pad * kp; };
  auto fbuf = [&](size_t off, int t, size_t per_step) { return reinterpret_cast<float*>(tape + off) + static_cast<size_t>(t) * per_step; };
  const float eps = 1e-5f;
  const size_t BD = static_cast<size_t>(B) * P.D, BS = static_cast<size_t>(B) * P.S;

  {
    const long long tot_e = static_cast<long long>(T) * m_pad * (P.Ep >> 3), tot_a = static_cast<long long>(T) * m_pad * (P.Ap >> 3);
    pack_steps_kernel<<<static_cast<unsigned>((tot_e + 255) / 256 > 148 * 16 ? 148 * 16 : (tot_e + 255) / 256), 256, 0, s>>>(
        embed, T, B, P.E, img(TP.e_img, 0, P.Ep), m_pad, P.Ep);
    count_launch();
    pack_steps_kernel<<<static_cast<unsigned>((tot_a + 255) / 256 > 148 * 16 ? 148 * 16 : (tot_a + 255) / 256), 256, 0, s>>>(
        actions, T, B, P.A, img(TP.a_img, 0, P.Ap), m_pad, P.Ap);
    count_launch();
    RLSB_CUDA_OK();
  }
  auto base = [&](const LayerPlan& L) {
    GemmParams g{};
    g.W = pbf(L.w_off); g.RB = L.RB; g.NB = L.NB; g.G = 1;
    g.M = B; g.m_tiles = m_tiles; g.N = L.N;
    g.bias = pf(L.bias_off);
    g.ln_eps = eps;
    return g;
  };
  // Linear -> [LN] -> ELU through fp32 pre-activations kept on the tape (needed by the backward pass)
  auto layer = [&](const LayerPlan& L, GemmParams g, float* sc, float* st, __nv_bfloat16* outp) -> int {
    g.out_f32 = sc; g.ldo = P.D; g.stats = st;
    RLSB_TRY(launch_gemm(g, P.ln ? EPI_STATS : EPI_PLAIN, s));
    return launch_ln_act(sc, P.D, st, L.NB, L.RB, B, m_pad, L.N, P.ln ? pf(L.g_off) : nullptr,
                         P.ln ? pf(L.b_off) : nullptr, eps, ACT_ELU, outp, P.Dp, s);
  };
  for (int t = 0; t < T; ++t) {
    {
      GemmParams g = base(P.img_in);
      g.n_seg = 2;
      g.A[0] = img(TP.z_img, t, P.Sp); g.a_ktiles[0] = P.Sp / 64;
      g.A[1] = img(TP.a_img, t, P.Ap); g.a_ktiles[1] = P.Ap / 64;
      RLSB_TRY(layer(P.img_in, g, fbuf(TP.sc_x, t, static_cast<size_t>(m_pad) * P.D),
                     fbuf(TP.st_x, t, static_cast<size_t>(P.img_in.NB) * m_pad * 2), img(TP.x_img, t, P.Dp)));
    }
    {
      GemmParams g = base(P.gru);
      g.n_seg = 2;
      g.A[0] = img(TP.x_img, t, P.Dp); g.a_ktiles[0] = P.Dp / 64;
      g.A[1] = img(TP.h_img, t, P.Dp); g.a_ktiles[1] = P.Dp / 64;
      float* sc = fbuf(TP.sc_g, t, static_cast<size_t>(m_pad) * TP.ld3);
      float* st = fbuf(TP.st_g, t, static_cast<size_t>(P.gru.NB) * m_pad * 2);
      g.out_f32 = sc; g.ldo = TP.ld3; g.stats = st;
      RLSB_TRY(launch_gemm(g, EPI_STATS, s));
      // h_{t-1} in fp32: the previous step's determ (zeros for the initial state: a cleared slot of the output)
      const float* hprev = t > 0 ? out->determ + static_cast<size_t>(t - 1) * BD : nullptr;
      if (!hprev) {
        cudaError_t e = cudaMemsetAsync(out->determ, 0, BD * 4, s);
        if (e != cudaSuccess) return static_cast<int>(e);
        hprev = out->determ;   // zeros; overwritten in place (each element is read before it is written)
      }
      RLSB_TRY(launch_gru_gate(sc, TP.ld3, st, P.gru.NB, P.gru.RB, B, m_pad, P.D, pf(P.gru.g_off), pf(P.gru.b_off), eps,
                               -1.0f, hprev, P.D, out->determ + static_cast<size_t>(t) * BD, P.D,
                               img(TP.h_img, t + 1, P.Dp), P.Dp, s));
    }
    {
      GemmParams g = base(P.prior1);
      g.n_seg = 1;
      g.A[0] = img(TP.h_img, t + 1, P.Dp); g.a_ktiles[0] = P.Dp / 64;
      RLSB_TRY(layer(P.prior1, g, fbuf(TP.sc_y, t, static_cast<size_t>(m_pad) * P.D),
                     fbuf(TP.st_y, t, static_cast<size_t>(P.prior1.NB) * m_pad * 2), img(TP.y_img, t, P.Dp)));
      GemmParams g2 = base(P.prior2);
      g2.n_seg = 1;
      g2.A[0] = img(TP.y_img, t, P.Dp); g2.a_ktiles[0] = P.Dp / 64;
      g2.out_f32 = out->prior_logits + static_cast<size_t>(t) * BS; g2.ldo = P.S;
      RLSB_TRY(launch_gemm(g2, EPI_PLAIN, s));
    }
    {
      GemmParams g = base(P.post1);
      g.n_seg = 2;
      g.A[0] = img(TP.h_img, t + 1, P.Dp); g.a_ktiles[0] = P.Dp / 64;
      g.A[1] = img(TP.e_img, t, P.Ep); g.a_ktiles[1] = P.Ep / 64;
      RLSB_TRY(layer(P.post1, g, fbuf(TP.sc_y2, t, static_cast<size_t>(m_pad) * P.D),
                     fbuf(TP.st_y2, t, static_cast<size_t>(P.post1.NB) * m_pad * 2), img(TP.y2_img, t, P.Dp)));
      GemmParams g2 = base(P.post2);
      g2.n_seg = 1;
      g2.A[0] = img(TP.y2_img, t, P.Dp); g2.a_ktiles[0] = P.Dp / 64;
      g2.out_f32 = out->post_logits + static_cast<size_t>(t) * BS; g2.ldo = P.S;
      RLSB_TRY(launch_gemm(g2, EPI_PLAIN, s));
    }
    {
      NoiseSpec ns{};
      ns.explicit_noise = noise->latent_uniforms ? noise->latent_uniforms + static_cast<size_t>(t) * BS : nullptr;
      ns.ld = P.S; ns.seed = noise->seed; ns.seed_ptr = noise->seed_device; ns.step = static_cast<uint32_t>(t);
      ns.row_offset = noise->row_offset;
      RLSB_TRY(launch_sample_latent(out->post_logits + static_cast<size_t>(t) * BS, P.S, B, cfg->groups, cfg->classes, ns,
                                    out->stoch_idx + static_cast<size_t>(t) * B * cfg->groups, img(TP.z_img, t + 1, P.Sp),
                                    P.Sp, out->stoch ? out->stoch + static_cast<size_t>(t) * BS : nullptr, P.S, s));
    }
  }
  return 0;
}

extern "C" int rlsb_observe_bwd(const rlsb_observe_cfg* cfg, const void* packed, int64_t B_, const void* tape_,
                                const rlsb_observe_out* fwd, const float* g_prior_logits, const float* g_post_logits,
                                const float* g_determ, const float* g_stoch, const rlsb_observe_grads* grads,
                                float* g_embed, void* workspace, void* stream_) {
  if (!cfg || !packed || !tape_ || !fwd || !grads || !g_embed || !workspace || B_ <= 0) return -1;
  if (!fwd->post_logits || !fwd->determ) return -2;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  ObsPlan P;
  RLSB_TRY(make_obs_plan(*cfg, P));
  ObsTape TP;
  make_obs_tape(P, B_, TP);
  ObsBwdWs W;
  RLSB_TRY(make_obs_bwd_ws(P, TP, W));
  const int B = static_cast<int>(B_), m_pad = TP.m_pad, m_tiles = m_pad / 128, T = P.T;
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  const uint8_t* tape = static_cast<const uint8_t*>(tape_);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  auto pbf = [&](size_t off) { return reinterpret_cast<const __nv_bfloat16*>(pk + off); };
  auto pf = [&](size_t off) { return reinterpret_cast<const float*>(pk + off); };
  auto timg = [&](size_t off, int t, int kp) { return reinterpret_cast<const __nv_bfloat16*>(tape + off) + static_cast<size_t>(t) * m_pad * kp; };
  auto tfb = [&](size_t off, int t, size_t per_step) { return reinterpret_cast<const float*>(tape + off) + static_cast<size_t>(t) * per_step; };
  auto wimg = [&](size_t off, int t, int kp) { return reinterpret_cast<__nv_bfloat16*>(ws + off) + static_cast<size_t>(t) * m_pad * kp; };
  auto wfb = [&](size_t off, int t, size_t per_step) { return reinterpret_cast<float*>(ws + off) + static_cast<size_t>(t) * per_step; };
  auto f32 = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
  const float eps = 1e-5f;
  const size_t BD = static_cast<size_t>(B) * P.D, BS = static_cast<size_t>(B) * P.S, mD = static_cast<size_t>(m_pad) * P.D;
  const __nv_bfloat16* ones = pbf(P.ones_off);

  auto dx = [&](const TLayer& Tl, const __nv_bfloat16* A, int n_out, float* outp, long long ldo) -> int {
    GemmParams g{};
    g.A[0] = A; g.a_ktiles[0] = Tl.kp / 64; g.n_seg = 1;
    g.W = pbf(Tl.off); g.RB = Tl.RB; g.NB = Tl.NB; g.G = 1;
    g.M = B; g.m_tiles = m_tiles; g.N = n_out;
    g.out_f32 = outp; g.ldo = ldo;
    return launch_gemm(g, EPI_PLAIN, s);
  };
  auto ln_bwd = [&](const float* g, long long ld_g, const float* pre, const float* st, const LayerPlan& L,
                    __nv_bfloat16* dp, float* da) -> int {
    if (launch_pdl(ln_act_bwd_kernel, static_cast<unsigned>((m_pad * 32 + 255) / 256), 256, 0, s, g, ld_g, pre, st, L.NB, B, m_pad, P.D,
                                                                P.ln ? pf(L.g_off) : nullptr, P.ln ? pf(L.b_off) : nullptr,
                                                                eps, dp, P.Dp, da) != cudaSuccess) return static_cast<int>(cudaGetLastError());
    count_launch();
    return static_cast<int>(cudaGetLastError());
  };

  for (int t = T - 1; t >= 0; --t) {
    const bool last = t == T - 1;
    // ---- posterior logits: external gradient + the straight-through sample's gradient --------------------
    {
      const long long tot = static_cast<long long>(m_pad) * cfg->groups;
      if (launch_pdl(st_softmax_bwd_kernel, static_cast<unsigned>(static_cast<unsigned>((tot + 127) / 128)), 128, 0, s, 
          fwd->post_logits + static_cast<size_t>(t) * BS, g_post_logits ? g_post_logits + static_cast<size_t>(t) * BS : nullptr,
          g_stoch ? g_stoch + static_cast<size_t>(t) * BS : nullptr, last ? nullptr : f32(W.g_za), W.ld_za, B, m_pad,
          cfg->groups, wimg(W.gl_post, t, P.Sp), P.Sp) != cudaSuccess) return static_cast<int>(cudaGetLastError());
      count_launch();
      RLSB_CUDA_OK();
      RLSB_TRY(dx(P.t_post2, wimg(W.gl_post, t, P.Sp), P.D, f32(W.g_y2), P.D));
      RLSB_TRY(ln_bwd(f32(W.g_y2), P.D, tfb(TP.sc_y2, t, mD), tfb(TP.st_y2, t, static_cast<size_t>(P.post1.NB) * m_pad * 2),
                      P.post1, wimg(W.dp2, t, P.Dp), wfb(W.da_y2, t, mD)));
      RLSB_TRY(dx(P.t_post1, wimg(W.dp2, t, P.Dp), P.Dp + P.Ep, f32(W.g_he), W.ld_he));
      if (launch_pdl(extract_steps_kernel, static_cast<unsigned>((B * P.E + 255) / 256), 256, 0, s, f32(W.g_he), W.ld_he, P.Dp, B, P.E,
                                                                 g_embed + static_cast<size_t>(t) * B * P.E) != cudaSuccess) return static_cast<int>(cudaGetLastError());
      count_launch();
      RLSB_CUDA_OK();
    }
    // ---- prior logits -----------------------------------------------------------------------------------------
    {
      const long long tot = static_cast<long long>(m_pad) * cfg->groups;
      if (launch_pdl(st_softmax_bwd_kernel, static_cast<unsigned>(static_cast<unsigned>((tot + 127) / 128)), 128, 0, s, 
          nullptr, g_prior_logits ? g_prior_logits + static_cast<size_t>(t) * BS : nullptr, nullptr, nullptr, 0, B, m_pad,
          cfg->groups, wimg(W.gl_prior, t, P.Sp), P.Sp) != cudaSuccess) return static_cast<int>(cudaGetLastError());
      count_launch();
      RLSB_CUDA_OK();
      RLSB_TRY(dx(P.t_prior2, wimg(W.gl_prior, t, P.Sp), P.D, f32(W.g_y), P.D));
      RLSB_TRY(ln_bwd(f32(W.g_y), P.D, tfb(TP.sc_y, t, mD), tfb(TP.st_y, t, static_cast<size_t>(P.prior1.NB) * m_pad * 2),
                      P.prior1, wimg(W.dp1, t, P.Dp), wfb(W.da_y, t, mD)));
      RLSB_TRY(dx(P.t_prior1, wimg(W.dp1, t, P.Dp), P.D, f32(W.g_hprior), P.D));
    }
    // ---- h_t = GRU(x_t, h_{t-1}) ------------------------------------------------------------------------------
    {
      GruBwdArgs a{};
      a.scratch = tfb(TP.sc_g, t, static_cast<size_t>(m_pad) * TP.ld3); a.ld = TP.ld3;
      a.stats = tfb(TP.st_g, t, static_cast<size_t>(P.gru.NB) * m_pad * 2);
      a.NB = P.gru.NB; a.M = B; a.m_pad = m_pad; a.D = P.D;
      a.gamma = pf(P.gru.g_off); a.beta = pf(P.gru.b_off); a.eps = eps; a.update_bias = -1.0f;
      a.h_prev = t > 0 ? fwd->determ + static_cast<size_t>(t - 1) * BD : nullptr;
      int n = 0;
      a.gh[n] = f32(W.g_he); a.ld_gh[n++] = W.ld_he;
      a.gh[n] = f32(W.g_hprior); a.ld_gh[n++] = P.D;
      if (g_determ) { a.gh[n] = g_determ + static_cast<size_t>(t) * BD; a.ld_gh[n++] = P.D; }
      if (!last) {
        a.gh[n] = f32(W.g_hdirect); a.ld_gh[n++] = P.D;
        a.gh[n] = f32(W.g_hgru); a.ld_gh[n++] = P.D;
      }
      a.n_gh = n;
      a.g_pre = wimg(W.g_pre, t, P.G3p); a.kpad = P.G3p;
      a.g_hdirect = f32(W.g_hdirect);
      a.da_out = wfb(W.da_g, t, static_cast<size_t>(m_pad) * TP.ld3);
      if (launch_pdl(gru_gate_bwd_kernel, static_cast<unsigned>((m_pad * 32 + 255) / 256), 256, 0, s, a) != cudaSuccess) return static_cast<int>(cudaGetLastError());
      count_launch();
      RLSB_CUDA_OK();
      RLSB_TRY(dx(P.t_gru_x, wimg(W.g_pre, t, P.G3p), P.D, f32(W.g_x), P.D));
      RLSB_TRY(dx(P.t_gru_h, wimg(W.g_pre, t, P.G3p), P.D, f32(W.g_hgru), P.D));
      RLSB_TRY(ln_bwd(f32(W.g_x), P.D, tfb(TP.sc_x, t, mD), tfb(TP.st_x, t, static_cast<size_t>(P.img_in.NB) * m_pad * 2),
                      P.img_in, wimg(W.dp_in, t, P.Dp), wfb(W.da_x, t, mD)));
    }
    // ---- x_t = ELU(LN?(W_in [z_{t-1}, a_t])): d loss / d z_{t-1} feeds the previous step's sample gradient -------
    if (t > 0) RLSB_TRY(dx(P.t_img_in, wimg(W.dp_in, t, P.Dp), P.Sp + P.Ap, f32(W.g_za), W.ld_za));
  }

  // ---- parameter gradients over all T steps at once ----------------------------------------------------------------
  const int mt_all = T * m_tiles;
  auto wgrad = [&](const __nv_bfloat16* dY, int n_pad, int n_out, const __nv_bfloat16* X0, int k0, const __nv_bfloat16* X1,
                   int k1, float* dW, float* db, int ld_dst, int n_seg, const PackSeg* segs) -> int {
    WgradParams wp{};
    wp.dY = dY; wp.n_tiles = n_pad / 64; wp.G = 1; wp.m_tiles = mt_all;
    int ns = 0;
    wp.X[ns] = X0; wp.x_ktiles[ns] = k0 / 64; wp.x_mtile_stride[ns] = static_cast<long long>(k0) * 128; ++ns;
    if (X1) { wp.X[ns] = X1; wp.x_ktiles[ns] = k1 / 64; wp.x_mtile_stride[ns] = static_cast<long long>(k1) * 128; ++ns; }
    wp.X[ns] = ones; wp.x_ktiles[ns] = 1; wp.x_mtile_stride[ns] = 0; ++ns;
    wp.n_seg = ns;
    wp.partial = f32(W.partial);
    RLSB_TRY(plan_wgrad(wp));
    RLSB_TRY(launch_wgrad(wp, s));
    WgradReduceParams rp{};
    rp.partial = wp.partial; rp.splits = wp.splits; rp.G = 1; rp.rows_pad = wp.n_slices * 128; rp.ld = wp.kt_total * 64;
    rp.w_dst[0] = dW; rp.b_dst[0] = db; rp.n_out[0] = n_out; rp.ld_dst = ld_dst; rp.n_seg = n_seg;
    for (int i = 0; i < n_seg; ++i) rp.seg[i] = segs[i];
    rp.ones_col = k0 + (X1 ? k1 : 0);
    return launch_wgrad_reduce(rp, s);
  };
  {
    PackSeg segs[2] = {{0, 0, P.S}, {P.Sp, P.S, P.A}};
    RLSB_TRY(wgrad(wimg(W.dp_in, 0, P.Dp), P.Dp, P.D, timg(TP.z_img, 0, P.Sp), P.Sp, timg(TP.a_img, 0, P.Ap), P.Ap,
                   grads->img_in_w, grads->img_in_b, P.S + P.A, 2, segs));
  }
  {
    PackSeg segs[2] = {{0, 0, P.D}, {P.Dp, P.D, P.D}};
    RLSB_TRY(wgrad(wimg(W.g_pre, 0, P.G3p), P.G3p, 3 * P.D, timg(TP.x_img, 0, P.Dp), P.Dp, timg(TP.h_img, 0, P.Dp), P.Dp,
                   grads->gru_w, grads->gru_b, 2 * P.D, 2, segs));
  }
  {
    PackSeg seg[1] = {{0, 0, P.D}};
    RLSB_TRY(wgrad(wimg(W.dp1, 0, P.Dp), P.Dp, P.D, timg(TP.h_img, 1, P.Dp), P.Dp, nullptr, 0, grads->prior1_w,
                   grads->prior1_b, P.D, 1, seg));
    RLSB_TRY(wgrad(wimg(W.gl_prior, 0, P.Sp), P.Sp, P.S, timg(TP.y_img, 0, P.Dp), P.Dp, nullptr, 0, grads->prior2_w,
                   grads->prior2_b, P.D, 1, seg));
    RLSB_TRY(wgrad(wimg(W.gl_post, 0, P.Sp), P.Sp, P.S, timg(TP.y2_img, 0, P.Dp), P.Dp, nullptr, 0, grads->post2_w,
                   grads->post2_b, P.D, 1, seg));
  }
  {
    PackSeg segs[2] = {{0, 0, P.D}, {P.Dp, P.D, P.E}};
    RLSB_TRY(wgrad(wimg(W.dp2, 0, P.Dp), P.Dp, P.D, timg(TP.h_img, 1, P.Dp), P.Dp, timg(TP.e_img, 0, P.Ep), P.Ep,
                   grads->post1_w, grads->post1_b, P.D + P.E, 2, segs));
  }
  // LayerNorm gamma / beta
  auto lnp = [&](size_t da_off, size_t sc_off, size_t st_off, long long ld, int NB, int N, float* dg, float* db) -> int {
    if (!dg || !db) return 0;
    ln_param_grad_kernel<<<(N + 127) / 128, 128, 0, s>>>(f32(da_off), reinterpret_cast<const float*>(tape + sc_off), ld,
                                                          reinterpret_cast<const float*>(tape + st_off), NB, T, B, m_pad, N,
                                                          eps, dg, db);
    count_launch();
    return static_cast<int>(cudaGetLastError());
  };
  RLSB_TRY(lnp(W.da_g, TP.sc_g, TP.st_g, TP.ld3, P.gru.NB, 3 * P.D, grads->gru_ln_g, grads->gru_ln_b));
  if (P.ln) {
    RLSB_TRY(lnp(W.da_x, TP.sc_x, TP.st_x, P.D, P.img_in.NB, P.D, grads->img_in_ln_g, grads->img_in_ln_b));
    RLSB_TRY(lnp(W.da_y, TP.sc_y, TP.st_y, P.D, P.prior1.NB, P.D, grads->prior1_ln_g, grads->prior1_ln_b));
    RLSB_TRY(lnp(W.da_y2, TP.sc_y2, TP.st_y2, P.D, P.post1.NB, P.D, grads->post1_ln_g, grads->post1_ln_b));
  }
  return 0;
}
