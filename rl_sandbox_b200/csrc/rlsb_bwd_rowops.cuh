// rlsb_bwd_rowops.cuh — row-wise device code of the backward rollout (agents/dreamer_v2.py:199-207 through the chain
// rssm.py:176-193, common.py:69-81, rssm.py:34-37), shared by the stand-alone kernels of rlsb_imagine_bwd.cu and the
// persistent backward kernel (rlsb_rollout.cu).  COHERENT: gradient buffers written earlier in the SAME launch by other
// CTAs are read through L2 (ld.global.cg).
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "rlsb_gemm.cuh"
#include "rlsb_ptx.cuh"

namespace rlsb {
namespace bwdops {

__device__ __forceinline__ uint32_t bfpair(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float fsig(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float ftanh(float x) {
  const float e = __expf(-2.0f * fabsf(x));
  return copysignf((1.0f - e) / (1.0f + e), x);
}

template <bool COHERENT>
__device__ __forceinline__ float ldc(const float* p) {
  return COHERENT ? __ldcg(p) : *p;
}

// d loss / d (head outputs) of the reward head and the target critic for row m -> packed [Gb][m_pad x 64]
__device__ __forceinline__ void head_grad_row(int m, const float* __restrict__ g_r, const float* __restrict__ g_v, int M,
                                              int m_pad, int Gb, int gb_reward, int gb_critic, __nv_bfloat16* __restrict__ dy4) {
  const size_t tile = static_cast<size_t>(m >> 7) * (kTileM * kTileK);
  const int row = m & 127;
  for (int gb = 0; gb < Gb; ++gb) {
    float v = 0.f;
    if (m < M) {
      if (gb == gb_reward && g_r) v = g_r[m];
      if (gb == gb_critic && g_v) v = g_v[m];
    }
    __nv_bfloat16* dst = dy4 + static_cast<size_t>(gb) * m_pad * 64 + tile + static_cast<size_t>(row) * kTileK;
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      uint4 u = make_uint4(0u, 0u, 0u, 0u);
      if (ch == 0) u.x = bfpair(v, 0.f);
      *reinterpret_cast<uint4*>(dst + ((ch ^ (row & 7)) << 3)) = u;
    }
  }
}

// z = onehot + p - p.detach(), p = softmax(logits) over each group of 32 classes (rssm.py:34-37):
// g_logit_j = p_j (g_z_j - sum_k g_z_k p_k) for (row m, group g); g_z = ga (+ gb).
template <bool COHERENT>
__device__ __forceinline__ void st_softmax_bwd_item(int m, int g, const float* __restrict__ logits, long long ld_l,
                                                    const float* __restrict__ ga, long long ld_a, const float* __restrict__ gb,
                                                    long long ld_b, __nv_bfloat16* __restrict__ out, int kpad) {
  const float* lp = logits + static_cast<size_t>(m) * ld_l + g * 32;
  const float* pa = ga + static_cast<size_t>(m) * ld_a + g * 32;
  const float* pb = gb ? gb + static_cast<size_t>(m) * ld_b + g * 32 : nullptr;
  float l[32], gz[32];
  float mx = -3.0e38f;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    l[k] = lp[k];
    gz[k] = ldc<COHERENT>(pa + k) + (pb ? ldc<COHERENT>(pb + k) : 0.f);
    mx = fmaxf(mx, l[k]);
  }
  float se = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    l[k] = __expf(l[k] - mx);
    se += l[k];
  }
  const float inv = 1.0f / se;
  float dot = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    l[k] *= inv;
    dot = fmaf(gz[k], l[k], dot);
  }
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = l[ch * 8 + j] * (gz[ch * 8 + j] - dot);
    const size_t idx = packed_index(static_cast<size_t>(m), static_cast<size_t>(g * 32 + ch * 8),
                                    static_cast<size_t>(kpad), kTileM);
    *reinterpret_cast<uint4*>(out + idx) = make_uint4(bfpair(o[0], o[1]), bfpair(o[2], o[3]), bfpair(o[4], o[5]),
                                                      bfpair(o[6], o[7]));
  }
}

struct GruBwdArgs {
  const float* scratch;   // [m_pad x ld] pre-LayerNorm gate activations (reset | cand | update)
  long long ld;
  const float* stats;     // per-(row, n-block) (sum, sum of squares)
  int NB, M, m_pad, D;
  const float* gamma;
  const float* beta;
  float eps, update_bias;
  const float* h_prev;    // (M, D) fp32
  long long ld_h;
  const float* gh[4];     // up to four additive sources of d loss / d h_t
  long long ld_gh[4];
  int n_gh;
  __nv_bfloat16* g_pre;   // packed [m_pad x kpad]: d loss / d (W [x, h] + b)
  int kpad;
  float* g_hdirect;       // (M, D): g_h * (1 - u)
};

// GRU gates + joint LayerNorm backward of row m by one warp, lane -> chunks of 8 consecutive j (D % 8 == 0)
template <bool COHERENT>
__device__ __forceinline__ void gru_gate_bwd_row(const GruBwdArgs& a, int m, int lane) {
  const int D = a.D;
  const int chunks = D >> 3;
  if (m >= a.M) {   // padding rows of the operand image: zeros
    for (int c = lane; c < (a.kpad >> 3); c += 32)
      *reinterpret_cast<uint4*>(a.g_pre + packed_index(static_cast<size_t>(m), static_cast<size_t>(c) * 8,
                                                       static_cast<size_t>(a.kpad), kTileM)) = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  // row statistics of the joint LayerNorm over 3D
  float s = 0.f, q = 0.f;
  {
    const float2* st = reinterpret_cast<const float2*>(a.stats);
    for (int b = lane; b < a.NB; b += 32) {
      const float2 v = __ldg(&st[static_cast<size_t>(b) * a.m_pad + m]);
      s += v.x;
      q += v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
  }
  const float inv_n = 1.0f / static_cast<float>(3 * D);
  const float mean = s * inv_n;
  const float rstd = 1.0f / sqrtf(fmaxf(q * inv_n - mean * mean, 0.f) + a.eps);
  const float* src = a.scratch + static_cast<size_t>(m) * a.ld;

  // d loss / d x_hat for the three gates of one j, and x_hat (recomputed in both passes)
  auto gate_grads = [&](int j, float& dr, float& dc, float& du, float& xr, float& xc, float& xu, float& ghd) {
    xr = (src[j] - mean) * rstd;
    xc = (src[D + j] - mean) * rstd;
    xu = (src[2 * D + j] - mean) * rstd;
    const float gr_ = __ldg(a.gamma + j), gc_ = __ldg(a.gamma + D + j), gu_ = __ldg(a.gamma + 2 * D + j);
    const float nr = fmaf(xr, gr_, __ldg(a.beta + j));
    const float nc = fmaf(xc, gc_, __ldg(a.beta + D + j));
    const float nu = fmaf(xu, gu_, __ldg(a.beta + 2 * D + j)) + a.update_bias;
    const float r = fsig(nr);
    const float c = ftanh(r * nc);
    const float u = fsig(nu);
    float gh = 0.f;
    for (int i = 0; i < a.n_gh; ++i) gh += ldc<COHERENT>(a.gh[i] + static_cast<size_t>(m) * a.ld_gh[i] + j);
    const float hp = a.h_prev[static_cast<size_t>(m) * a.ld_h + j];
    const float g_u = gh * (c - hp);
    const float g_t = gh * u * (1.0f - c * c);      // d / d (r * nc)
    dr = g_t * nc * r * (1.0f - r) * gr_;
    dc = g_t * r * gc_;
    du = g_u * u * (1.0f - u) * gu_;
    ghd = gh * (1.0f - u);
  };

  float s1 = 0.f, s2 = 0.f;
  for (int ch = lane; ch < chunks; ch += 32) {
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      float dr, dc, du, xr, xc, xu, ghd;
      gate_grads(ch * 8 + jj, dr, dc, du, xr, xc, xu, ghd);
      s1 += dr + dc + du;
      s2 += dr * xr + dc * xc + du * xu;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  const float m1 = s1 * inv_n, m2 = s2 * inv_n;
  for (int ch = lane; ch < chunks; ch += 32) {
    float o_r[8], o_c[8], o_u[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      float dr, dc, du, xr, xc, xu, ghd;
      const int j = ch * 8 + jj;
      gate_grads(j, dr, dc, du, xr, xc, xu, ghd);
      o_r[jj] = rstd * (dr - m1 - xr * m2);
      o_c[jj] = rstd * (dc - m1 - xc * m2);
      o_u[jj] = rstd * (du - m1 - xu * m2);
      a.g_hdirect[static_cast<size_t>(m) * D + j] = ghd;
    }
    auto put = [&](int col, const float (&o)[8]) {
      *reinterpret_cast<uint4*>(a.g_pre + packed_index(static_cast<size_t>(m), static_cast<size_t>(col),
                                                       static_cast<size_t>(a.kpad), kTileM)) =
          make_uint4(bfpair(o[0], o[1]), bfpair(o[2], o[3]), bfpair(o[4], o[5]), bfpair(o[6], o[7]));
    };
    put(ch * 8, o_r);
    put(D + ch * 8, o_c);
    put(2 * D + ch * 8, o_u);
  }
  // padding columns [3D, kpad)
  for (int c = (3 * D >> 3) + lane; c < (a.kpad >> 3); c += 32)
    *reinterpret_cast<uint4*>(a.g_pre + packed_index(static_cast<size_t>(m), static_cast<size_t>(c) * 8,
                                                     static_cast<size_t>(a.kpad), kTileM)) = make_uint4(0u, 0u, 0u, 0u);
}

}  // namespace bwdops
}  // namespace rlsb
