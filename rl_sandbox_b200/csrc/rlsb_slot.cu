// rlsb_slot.cu — K3: slot attention (reference: rl_sandbox/vision/slot_attention.py:13-77).
//
// k/v projection and the per-slot GRU / MLP contractions run on the tcgen05 GEMM of rlsb_gemm.cu;
// the attention proper (softmax over slots, renormalisation over tokens, weighted mean) is the
// HBM-bound part: one CTA per frame streams that frame's bf16 k and v exactly once per iteration
// (2 * tokens * dim * 2 B = 301 KB at 196 x 384) from the packed operand image the GEMM wrote.
#include "../../include/rlsb.h"
#include "rlsb_count.cuh"
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

__device__ __forceinline__ uint32_t bf2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}

// ------------------------------------------------------------------------------------------
// rows of fp32 (optionally a + b, the residual of slot_attention.py:76) -> LayerNorm -> packed bf16
// one warp per row; C % 8 == 0, C <= 1024
// ------------------------------------------------------------------------------------------
__global__ void layernorm_pack_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                      float* __restrict__ sum_out, long long rows, int rows_pad, int C,
                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                      float eps, __nv_bfloat16* __restrict__ out, int kpad) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const int chunks = C >> 3, kchunks = kpad >> 3;
  for (long long r = warp0; r < rows_pad; r += nwarps) {
    float v[4][8];
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int ch = lane + 32 * u;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[u][j] = 0.f;
      if (r < rows && ch < chunks) {
        const float4* pa = reinterpret_cast<const float4*>(a + r * C + ch * 8);
        float4 x0 = pa[0], x1 = pa[1];
        if (b) {
          const float4* pb = reinterpret_cast<const float4*>(b + r * C + ch * 8);
          const float4 y0 = pb[0], y1 = pb[1];
          x0.x += y0.x; x0.y += y0.y; x0.z += y0.z; x0.w += y0.w;
          x1.x += y1.x; x1.y += y1.y; x1.z += y1.z; x1.w += y1.w;
          if (sum_out) {
            float4* po = reinterpret_cast<float4*>(sum_out + r * C + ch * 8);
            po[0] = x0; po[1] = x1;
          }
        }
        v[u][0] = x0.x; v[u][1] = x0.y; v[u][2] = x0.z; v[u][3] = x0.w;
        v[u][4] = x1.x; v[u][5] = x1.y; v[u][6] = x1.z; v[u][7] = x1.w;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[u][j];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / static_cast<float>(C);
    float m2 = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (lane + 32 * u < chunks) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[u][j] - mean;
          m2 = fmaf(d, d, m2);
        }
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, o);
    const float rstd = 1.0f / sqrtf(m2 / static_cast<float>(C) + eps);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int ch = lane + 32 * u;
      if (ch < kchunks) {
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float t = 0.f;
          if (r < rows && ch < chunks) {
            t = gamma ? (v[u][j] - mean) * rstd * __ldg(gamma + ch * 8 + j) + __ldg(beta + ch * 8 + j) : v[u][j];
          }
          y[j] = t;
        }
        const size_t idx = packed_index(static_cast<size_t>(r), static_cast<size_t>(ch * 8),
                                        static_cast<size_t>(kpad), kTileM);
        *reinterpret_cast<uint4*>(out + idx) =
            make_uint4(bf2(y[0], y[1]), bf2(y[2], y[3]), bf2(y[4], y[5]), bf2(y[6], y[7]));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// attention for one frame per CTA (slot_attention.py:69-74)
// ------------------------------------------------------------------------------------------
constexpr int kAttnThreads = 256;
constexpr int kAttnWarps = kAttnThreads / 32;
constexpr int kMaxSlots = 8;

struct AttnArgs {
  const __nv_bfloat16* kv;  // packed [BT_pad x 2*dim]
  const float* q;           // [B*K][dim] fp32
  int T, K, dim;
  float scale, eps;
  float* attn_out;          // [B][K][T] or nullptr
  float* upd_out;           // [B*K][dim] fp32
  __nv_bfloat16* upd_packed;  // packed [BK_pad x dim]
};

template <int K>
__global__ void __launch_bounds__(kAttnThreads, 2) slot_attn_kernel(const AttnArgs a) {
  // HBM-bound: k and v of the frame (2 x T x dim bf16) are streamed once each.  Two token rows per warp iteration keep four
  // 16-byte loads per lane in flight; q is read from shared memory as float4 vectors; the cross-warp reduction of the
  // updates is a three-round tree over a half-size buffer, so that six CTAs fit on an SM.
  extern __shared__ float sm[];
  const int dim = a.dim, T = a.T;
  const int chunks = dim >> 3;  // 16-byte chunks per k (or v) row
  float* sq = sm;                      // [K][dim]
  float* sattn = sq + K * dim;         // [K][T]
  float* sred = sattn + ((K * T + 3) & ~3);   // [warps / 2][K][dim], 16-byte aligned
  float* srs = sred + (kAttnWarps / 2) * K * dim;  // [K] row sums
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < K * dim; i += kAttnThreads) sq[i] = a.q[static_cast<size_t>(b) * K * dim + i];
  __syncthreads();
  const int kpad = 2 * dim;
  const int c0 = lane, c1 = lane + 32;
  const bool has1 = c1 < chunks;
  auto kv_chunk = [&](size_t row, int col) {
    return __ldg(reinterpret_cast<const uint4*>(
        a.kv + packed_index(row, static_cast<size_t>(col), static_cast<size_t>(kpad), kTileM)));
  };
  // ---- logits, softmax over slots -------------------------------------------------------------
  // the four 16-byte pieces of token rows j, j + 1 (column offset `col`: 0 = k, dim = v); zero beyond the frame
  auto fetch = [&](int j, int col, uint4 (&u)[2][2]) {
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    const size_t row = static_cast<size_t>(b) * T + j;
    const bool one = j < T, two = j + 1 < T;
    u[0][0] = (one && c0 < chunks) ? kv_chunk(row, col + c0 * 8) : zero;
    u[0][1] = (one && has1) ? kv_chunk(row, col + c1 * 8) : zero;
    u[1][0] = (two && c0 < chunks) ? kv_chunk(row + 1, col + c0 * 8) : zero;
    u[1][1] = (two && has1) ? kv_chunk(row + 1, col + c1 * 8) : zero;
  };
  uint4 nxt[2][2];
  fetch(2 * warp, 0, nxt);
  for (int j = 2 * warp; j < T; j += 2 * kAttnWarps) {
    const bool two = j + 1 < T;
    uint4 u[2][2];
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int h = 0; h < 2; ++h) u[t][h] = nxt[t][h];
    fetch(j + 2 * kAttnWarps, 0, nxt);   // the next pair of rows is in flight while this one is reduced
    float acc[2][K];
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int i = 0; i < K; ++i) acc[t][i] = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = h ? c1 : c0;
      if (c < chunks) {
        float kf[2][8];
        unpack8(u[0][h], kf[0]);
        unpack8(u[1][h], kf[1]);
#pragma unroll
        for (int i = 0; i < K; ++i) {
          const float4 q0 = *reinterpret_cast<const float4*>(sq + i * dim + c * 8);
          const float4 q1 = *reinterpret_cast<const float4*>(sq + i * dim + c * 8 + 4);
          const float qv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            acc[0][i] = fmaf(qv[e], kf[0][e], acc[0][i]);
            acc[1][i] = fmaf(qv[e], kf[1][e], acc[1][i]);
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int i = 0; i < K; ++i)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[t][i] += __shfl_xor_sync(0xffffffffu, acc[t][i], o);
    if (lane < 2 && (lane == 0 || two)) {
      const int t = lane;
      float lg[K];
#pragma unroll
      for (int i = 0; i < K; ++i) lg[i] = (t ? acc[1][i] : acc[0][i]) * a.scale;
      float mx = lg[0];
#pragma unroll
      for (int i = 1; i < K; ++i) mx = fmaxf(mx, lg[i]);
      float e[K], den = 0.f;
#pragma unroll
      for (int i = 0; i < K; ++i) {
        e[i] = __expf(lg[i] - mx);
        den += e[i];
      }
#pragma unroll
      for (int i = 0; i < K; ++i) sattn[i * T + j + t] = e[i] / den + a.eps;
    }
  }
  __syncthreads();
  // ---- renormalise over tokens ------------------------------------------------------------------
  if (warp < K) {
    float s = 0.f;
    for (int j = lane; j < T; j += 32) s += sattn[warp * T + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) srs[warp] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * T; i += kAttnThreads) {
    const float v = sattn[i] / srs[i / T];
    sattn[i] = v;
    if (a.attn_out) a.attn_out[static_cast<size_t>(b) * K * T + i] = v;
  }
  __syncthreads();
  // ---- updates = attn . v  (each warp accumulates a token subset, then a tree reduction over the warps) --------
  float acc[K][2][8];
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[i][h][e] = 0.f;
  fetch(2 * warp, dim, nxt);
  for (int j = 2 * warp; j < T; j += 2 * kAttnWarps) {
    const bool two = j + 1 < T;
    uint4 x[2][2];
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int h = 0; h < 2; ++h) x[t][h] = nxt[t][h];
    fetch(j + 2 * kAttnWarps, dim, nxt);
    float w[2][K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
      w[0][i] = sattn[i * T + j];
      w[1][i] = two ? sattn[i * T + j + 1] : 0.f;
    }
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float vf[8];
        unpack8(x[t][h], vf);
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[i][h][e] = fmaf(w[t][i], vf[e], acc[i][h][e]);
      }
  }
  // tree over the 8 warps: the upper half of the active warps hands its partial sums to the lower half
  for (int half = kAttnWarps / 2; half >= 1; half >>= 1) {
    if (warp >= half && warp < 2 * half) {
#pragma unroll
      for (int i = 0; i < K; ++i)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = h ? c1 : c0;
          if (c < chunks) {
            float4* dst = reinterpret_cast<float4*>(sred + ((warp - half) * K + i) * dim + c * 8);
            dst[0] = make_float4(acc[i][h][0], acc[i][h][1], acc[i][h][2], acc[i][h][3]);
            dst[1] = make_float4(acc[i][h][4], acc[i][h][5], acc[i][h][6], acc[i][h][7]);
          }
        }
    }
    __syncthreads();
    if (warp < half) {
#pragma unroll
      for (int i = 0; i < K; ++i)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = h ? c1 : c0;
          if (c < chunks) {
            const float4* src = reinterpret_cast<const float4*>(sred + (warp * K + i) * dim + c * 8);
            const float4 s0 = src[0], s1 = src[1];
            acc[i][h][0] += s0.x; acc[i][h][1] += s0.y; acc[i][h][2] += s0.z; acc[i][h][3] += s0.w;
            acc[i][h][4] += s1.x; acc[i][h][5] += s1.y; acc[i][h][6] += s1.z; acc[i][h][7] += s1.w;
          }
        }
    }
    __syncthreads();
  }
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = h ? c1 : c0;
        if (c < chunks) {
          const float* sv = acc[i][h];
          const size_t r = static_cast<size_t>(b) * K + i;
          float4* po = reinterpret_cast<float4*>(a.upd_out + r * dim + c * 8);
          po[0] = make_float4(sv[0], sv[1], sv[2], sv[3]);
          po[1] = make_float4(sv[4], sv[5], sv[6], sv[7]);
          *reinterpret_cast<uint4*>(a.upd_packed +
                                    packed_index(r, static_cast<size_t>(c * 8), static_cast<size_t>(dim), kTileM)) =
              make_uint4(bf2(sv[0], sv[1]), bf2(sv[2], sv[3]), bf2(sv[4], sv[5]), bf2(sv[6], sv[7]));
        }
      }
  }
}

// nn.GRUCell gates (torch order r, z, n): gi = W_ih u + b_ih, gh = W_hh s + b_hh
__global__ void slot_gru_kernel(const float* __restrict__ gi, const float* __restrict__ gh, long long ld,
                                const float* __restrict__ s_prev, long long rows, int dim,
                                float* __restrict__ s_new) {
  const long long total = rows * dim;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / dim;
    const int d = static_cast<int>(i - r * dim);
    const float* a = gi + r * ld;
    const float* b = gh + r * ld;
    const float rg = 1.0f / (1.0f + __expf(-(a[d] + b[d])));
    const float zg = 1.0f / (1.0f + __expf(-(a[dim + d] + b[dim + d])));
    const float n = tanhf(a[2 * dim + d] + rg * b[2 * dim + d]);
    s_new[i] = (1.0f - zg) * n + zg * s_prev[i];
  }
}

__global__ void add_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                                float* __restrict__ out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = a[i] + b[i];
}

__global__ void copy_pad_kernel2(const float* __restrict__ src, int n, float* __restrict__ dst, int n_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) dst[i] = (src && i < n) ? src[i] : 0.f;
}

// the "ones" K tile (element (row, 0) = 1): bias gradients come out of the weight-gradient contraction
__global__ void slot_ones_tile_kernel(__nv_bfloat16* dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128 * 8) return;
  const int row = i >> 3, pos = i & 7;
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if ((pos ^ (row & 7)) == 0) v.x = 0x00003F80u;
  reinterpret_cast<uint4*>(dst)[i] = v;
}

struct WPlan {
  int N, K, RB, NB;
  size_t w_off, b_off;
};
void plan_w(WPlan& L, size_t& cur) {
  if (ru(L.N, 32) <= 512) { L.NB = 1; L.RB = ru(L.N, 32); }
  else { L.NB = (L.N + 255) / 256; L.RB = ru((L.N + L.NB - 1) / L.NB, 32); }
  L.w_off = place(cur, static_cast<size_t>(L.NB) * L.RB * L.K * 2);
  L.b_off = place(cur, static_cast<size_t>(L.NB) * L.RB * 4);
}
struct SlotPlan {
  int dim, K, T, iters;
  WPlan kv, q, ih, hh, m1, m2;
  // transposed images for the backward dX GEMMs: rows = in-features, K = out-features
  WPlan t_kv, t_q, t_ih, t_hh, t_m1, t_m2;
  size_t ones_off;
  size_t ln_in_g, ln_in_b, ln_s_g, ln_s_b, ln_2_g, ln_2_b;
  size_t bytes;
};
int make_slot_plan(const rlsb_slot_cfg& c, SlotPlan& P) {
  if (c.dim <= 0 || (c.dim % 64) != 0 || c.dim > 512) return -30;
  if (c.slots <= 0 || c.slots > kMaxSlots || c.tokens <= 0 || c.iters <= 0) return -31;
  P.dim = c.dim; P.K = c.slots; P.T = c.tokens; P.iters = c.iters;
  size_t cur = 0;
  P.kv = {2 * c.dim, c.dim}; plan_w(P.kv, cur);
  P.q = {c.dim, c.dim}; plan_w(P.q, cur);
  P.ih = {3 * c.dim, c.dim}; plan_w(P.ih, cur);
  P.hh = {3 * c.dim, c.dim}; plan_w(P.hh, cur);
  P.m1 = {4 * c.dim, c.dim}; plan_w(P.m1, cur);
  P.m2 = {c.dim, 4 * c.dim}; plan_w(P.m2, cur);
  P.t_kv = {c.dim, 2 * c.dim}; plan_w(P.t_kv, cur);
  P.t_q = {c.dim, c.dim}; plan_w(P.t_q, cur);
  P.t_ih = {c.dim, 3 * c.dim}; plan_w(P.t_ih, cur);
  P.t_hh = {c.dim, 3 * c.dim}; plan_w(P.t_hh, cur);
  P.t_m1 = {c.dim, 4 * c.dim}; plan_w(P.t_m1, cur);
  P.t_m2 = {4 * c.dim, c.dim}; plan_w(P.t_m2, cur);
  P.ones_off = place(cur, 128 * 64 * 2);
  size_t* lns[6] = {&P.ln_in_g, &P.ln_in_b, &P.ln_s_g, &P.ln_s_b, &P.ln_2_g, &P.ln_2_b};
  for (auto* o : lns) *o = place(cur, static_cast<size_t>(c.dim) * 4);
  P.bytes = rus(cur, 1024);
  return 0;
}
struct SlotWs {
  size_t xn, kv, sn, q, upd, updp, sprevp, gi, gh, snew, sn2, hid, mlp, cur[2];
  int bt_pad, bk_pad;
  long long ld_g;
  size_t bytes;
};
void make_slot_ws(const SlotPlan& P, long long B, SlotWs& W) {
  const long long BT = B * P.T, BK = B * P.K;
  W.bt_pad = ru(static_cast<int>(BT), 128);
  W.bk_pad = ru(static_cast<int>(BK), 128);
  size_t cur = 0;
  W.xn = place(cur, static_cast<size_t>(W.bt_pad) * P.dim * 2);
  W.kv = place(cur, static_cast<size_t>(W.bt_pad) * 2 * P.dim * 2);
  W.sn = place(cur, static_cast<size_t>(W.bk_pad) * P.dim * 2);
  W.q = place(cur, static_cast<size_t>(W.bk_pad) * P.dim * 4);
  W.upd = place(cur, static_cast<size_t>(W.bk_pad) * P.dim * 4);
  W.updp = place(cur, static_cast<size_t>(W.bk_pad) * P.dim * 2);
  W.sprevp = place(cur, static_cast<size_t>(W.bk_pad) * P.dim * 2);
  W.ld_g = 3 * P.dim;
  W.gi = place(cur, static_cast<size_t>(W.bk_pad) * W.ld_g * 4);
  W.gh = place(cur, static_cast<size_t>(W.bk_pad) * W.ld_g * 4);
  W.snew = place(cur, static_cast<size_t>(W.bk_pad) * P.dim * 4);
  W.sn2 = place(cur, static_cast<size_t>(W.bk_pad) * P.dim * 2);
  W.hid = place(cur, static_cast<size_t>(W.bk_pad) * 4 * P.dim * 2);
  W.mlp = place(cur, static_cast<size_t>(W.bk_pad) * P.dim * 4);
  for (int i = 0; i < 2; ++i) W.cur[i] = place(cur, static_cast<size_t>(W.bk_pad) * P.dim * 4);
  W.bytes = rus(cur, 1024);
}

// activations the backward pass needs (rlsb_slot_attention_fwd with a tape): global + per iteration
struct SlotTape {
  size_t xn, kv;                                               // packed LN(X), packed k|v
  size_t sprev, sn, q, updp, sprevp, gi, gh, snew, sn2, hid;  // offsets inside one iteration block
  size_t iter0, iter_bytes;
  size_t bytes;
};
void make_slot_tape(const SlotPlan& P, const SlotWs& W, SlotTape& T) {
  size_t cur = 0;
  T.xn = place(cur, static_cast<size_t>(W.bt_pad) * P.dim * 2);
  T.kv = place(cur, static_cast<size_t>(W.bt_pad) * 2 * P.dim * 2);
  T.iter0 = rus(cur, 1024);
  size_t c = 0;
  const size_t bk = static_cast<size_t>(W.bk_pad);
  T.sprev = place(c, bk * P.dim * 4);
  T.sn = place(c, bk * P.dim * 2);
  T.q = place(c, bk * P.dim * 4);
  T.updp = place(c, bk * P.dim * 2);
  T.sprevp = place(c, bk * P.dim * 2);
  T.gi = place(c, bk * W.ld_g * 4);
  T.gh = place(c, bk * W.ld_g * 4);
  T.snew = place(c, bk * P.dim * 4);
  T.sn2 = place(c, bk * P.dim * 2);
  T.hid = place(c, bk * 4 * P.dim * 2);
  T.iter_bytes = rus(c, 1024);
  T.bytes = T.iter0 + T.iter_bytes * P.iters;
}

#define RLSB_TRY(expr)      \
  do {                      \
    int _e = (expr);        \
    if (_e != 0) return _e; \
  } while (0)
#define RLSB_CUDA(expr)                                   \
  do {                                                    \
    cudaError_t _e = (expr);                              \
    if (_e != cudaSuccess) return static_cast<int>(_e);   \
  } while (0)

int grid_for(long long total, int block, int cap) {
  long long g = (total + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace
}  // namespace rlsb

using namespace rlsb;

extern "C" size_t rlsb_slot_attention_packed_bytes(const rlsb_slot_cfg* cfg) {
  SlotPlan P;
  if (!cfg || make_slot_plan(*cfg, P) != 0) return 0;
  return P.bytes;
}

extern "C" size_t rlsb_slot_attention_workspace_bytes(const rlsb_slot_cfg* cfg, int64_t B) {
  SlotPlan P;
  if (!cfg || B <= 0 || make_slot_plan(*cfg, P) != 0) return 0;
  SlotWs W;
  make_slot_ws(P, B, W);
  return W.bytes;
}

extern "C" int rlsb_slot_attention_pack(const rlsb_slot_cfg* cfg, const rlsb_slot_params* p, void* packed,
                                        void* stream_) {
  if (!cfg || !p || !packed) return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  SlotPlan P;
  RLSB_TRY(make_slot_plan(*cfg, P));
  uint8_t* base = static_cast<uint8_t*>(packed);
  auto pack_w = [&](const WPlan& L, const float* w, const float* b) -> int {
    if (!w) return -32;
    PackSeg seg{0, 0, L.K};
    RLSB_TRY(launch_pack(w, L.K, L.N, reinterpret_cast<__nv_bfloat16*>(base + L.w_off), L.RB, L.NB * L.RB, L.K, 1,
                         &seg, s));
    const int n_pad = L.NB * L.RB;
    copy_pad_kernel2<<<(n_pad + 255) / 256, 256, 0, s>>>(b, L.N, reinterpret_cast<float*>(base + L.b_off), n_pad);
    count_launch();
    return static_cast<int>(cudaGetLastError());
  };
  RLSB_TRY(pack_w(P.kv, p->inputs_proj_w, nullptr));
  RLSB_TRY(pack_w(P.q, p->slots_proj_w, nullptr));
  RLSB_TRY(pack_w(P.ih, p->gru_w_ih, p->gru_b_ih));
  RLSB_TRY(pack_w(P.hh, p->gru_w_hh, p->gru_b_hh));
  RLSB_TRY(pack_w(P.m1, p->mlp_w1, p->mlp_b1));
  RLSB_TRY(pack_w(P.m2, p->mlp_w2, p->mlp_b2));
  auto pack_t = [&](const WPlan& L, const float* w) -> int {   // L.N = in-features (rows), L.K = out-features
    return launch_pack_transposed(w, L.N, L.K, L.N, reinterpret_cast<__nv_bfloat16*>(base + L.w_off), L.RB,
                                  L.NB * L.RB, L.K, s);
  };
  RLSB_TRY(pack_t(P.t_kv, p->inputs_proj_w));
  RLSB_TRY(pack_t(P.t_q, p->slots_proj_w));
  RLSB_TRY(pack_t(P.t_ih, p->gru_w_ih));
  RLSB_TRY(pack_t(P.t_hh, p->gru_w_hh));
  RLSB_TRY(pack_t(P.t_m1, p->mlp_w1));
  RLSB_TRY(pack_t(P.t_m2, p->mlp_w2));
  slot_ones_tile_kernel<<<8, 128, 0, s>>>(reinterpret_cast<__nv_bfloat16*>(base + P.ones_off));
  count_launch();
  const float* lsrc[6] = {p->inputs_norm_g, p->inputs_norm_b, p->slots_norm_g, p->slots_norm_b,
                          p->slots_norm2_g, p->slots_norm2_b};
  const size_t loff[6] = {P.ln_in_g, P.ln_in_b, P.ln_s_g, P.ln_s_b, P.ln_2_g, P.ln_2_b};
  for (int i = 0; i < 6; ++i) {
    if (!lsrc[i]) return -33;
    RLSB_CUDA(cudaMemcpyAsync(base + loff[i], lsrc[i], static_cast<size_t>(P.dim) * 4, cudaMemcpyDeviceToDevice, s));
  }
  return 0;
}

static int slot_fwd_impl(const rlsb_slot_cfg* cfg, const void* packed, int64_t B, const float* X,
                         const float* prev_slots, float* out_slots, float* out_attn, void* workspace, void* tape_,
                         void* stream_) {
  if (!cfg || !packed || !X || !prev_slots || !out_slots || !workspace || B <= 0) return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  SlotPlan P;
  RLSB_TRY(make_slot_plan(*cfg, P));
  SlotWs W;
  make_slot_ws(P, B, W);
  const long long BT = B * P.T, BK = B * P.K;
  if (BT > (1LL << 30)) return -3;
  const int dim = P.dim;
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  uint8_t* tape = static_cast<uint8_t*>(tape_);
  SlotTape TP{};
  if (tape) make_slot_tape(P, W, TP);
  int tape_it = 0;   // iteration whose tape block receives the activations
  // a buffer lives in the workspace, or — when a tape is recorded — in the tape (global / per-iteration block)
  auto loc = [&](size_t ws_off, size_t tape_off, bool per_iter) -> uint8_t* {
    if (!tape) return ws + ws_off;
    return tape + (per_iter ? TP.iter0 + TP.iter_bytes * tape_it + tape_off : tape_off);
  };
  auto bf = [&](size_t off) { return reinterpret_cast<__nv_bfloat16*>(ws + off); };
  auto f32 = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
  auto pf = [&](size_t off) { return reinterpret_cast<const float*>(pk + off); };
  auto gemm = [&](const WPlan& L, const __nv_bfloat16* A, int M) {
    GemmParams g{};
    g.A[0] = A; g.a_ktiles[0] = L.K / 64; g.n_seg = 1;
    g.W = reinterpret_cast<const __nv_bfloat16*>(pk + L.w_off);
    g.RB = L.RB; g.NB = L.NB; g.G = 1; g.M = M; g.m_tiles = (M + 127) / 128; g.N = L.N;
    g.bias = pf(L.b_off);
    return g;
  };
  auto ln_pack = [&](const float* a, const float* b, float* sum_out, long long rows, int rows_pad, const float* gam,
                     const float* bet, __nv_bfloat16* out) -> int {
    const long long threads = static_cast<long long>(rows_pad) * 32;
    layernorm_pack_kernel<<<grid_for(threads, 256, 148 * 16), 256, 0, s>>>(a, b, sum_out, rows, rows_pad, dim, gam, bet,
                                                                           1e-5f, out, dim);
    count_launch();
    return static_cast<int>(cudaGetLastError());
  };

  // ---- k, v = W_kv LN(X)            slot_attention.py:54 ------------------------------------------
  __nv_bfloat16* xn_img = reinterpret_cast<__nv_bfloat16*>(loc(W.xn, TP.xn, false));
  __nv_bfloat16* kv_img = reinterpret_cast<__nv_bfloat16*>(loc(W.kv, TP.kv, false));
  RLSB_TRY(ln_pack(X, nullptr, nullptr, BT, W.bt_pad, pf(P.ln_in_g), pf(P.ln_in_b), xn_img));
  {
    GemmParams g = gemm(P.kv, xn_img, static_cast<int>(BT));
    g.act = ACT_NONE; g.out_bf16 = kv_img; g.out_kpad = 2 * dim;
    RLSB_TRY(launch_gemm(g, EPI_LN_ACT, s));
  }
  RLSB_CUDA(cudaMemcpyAsync(f32(W.cur[0]), prev_slots, static_cast<size_t>(BK) * dim * 4, cudaMemcpyDeviceToDevice, s));
  static PerDeviceOnce attr_once[kMaxSlots + 1];
  unsigned long long dev_bit = 0;
  const size_t attn_smem = (static_cast<size_t>(P.K) * dim + ((static_cast<size_t>(P.K) * P.T + 3) & ~static_cast<size_t>(3)) +
                            static_cast<size_t>(kAttnWarps / 2) * P.K * dim + kMaxSlots) * sizeof(float);
  int cur = 0;
  for (int it = 0; it < P.iters; ++it) {
    float* s_prev = f32(W.cur[cur]);
    float* s_out = (it == P.iters - 1) ? out_slots : f32(W.cur[cur ^ 1]);
    tape_it = it;
    __nv_bfloat16* sn_img = reinterpret_cast<__nv_bfloat16*>(loc(W.sn, TP.sn, true));
    float* q_buf = reinterpret_cast<float*>(loc(W.q, TP.q, true));
    __nv_bfloat16* updp_img = reinterpret_cast<__nv_bfloat16*>(loc(W.updp, TP.updp, true));
    __nv_bfloat16* sprevp_img = reinterpret_cast<__nv_bfloat16*>(loc(W.sprevp, TP.sprevp, true));
    float* gi_buf = reinterpret_cast<float*>(loc(W.gi, TP.gi, true));
    float* gh_buf = reinterpret_cast<float*>(loc(W.gh, TP.gh, true));
    float* snew_buf = reinterpret_cast<float*>(loc(W.snew, TP.snew, true));
    __nv_bfloat16* sn2_img = reinterpret_cast<__nv_bfloat16*>(loc(W.sn2, TP.sn2, true));
    __nv_bfloat16* hid_img = reinterpret_cast<__nv_bfloat16*>(loc(W.hid, TP.hid, true));
    if (tape)
      RLSB_CUDA(cudaMemcpyAsync(loc(0, TP.sprev, true), s_prev, static_cast<size_t>(BK) * dim * 4,
                                cudaMemcpyDeviceToDevice, s));
    // q = W_q LN(slots)             slot_attention.py:66-67
    RLSB_TRY(ln_pack(s_prev, nullptr, nullptr, BK, W.bk_pad, pf(P.ln_s_g), pf(P.ln_s_b), sn_img));
    {
      GemmParams g = gemm(P.q, sn_img, static_cast<int>(BK));
      g.out_f32 = q_buf; g.ldo = dim;
      RLSB_TRY(launch_gemm(g, EPI_PLAIN, s));
    }
    // attention                     slot_attention.py:69-74
    {
      AttnArgs a{kv_img, q_buf, P.T, P.K, dim, 1.0f / sqrtf(static_cast<float>(dim)), 1e-8f,
                 (it == P.iters - 1) ? out_attn : nullptr, f32(W.upd), updp_img};
#define RLSB_ATTN(KK)                                                                                   \
  case KK:                                                                                              \
    if (attr_once[KK].need(dev_bit)) {                                                                  \
      RLSB_CUDA(cudaFuncSetAttribute(slot_attn_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     200 * 1024));                                                      \
      attr_once[KK].done(dev_bit);                                                                      \
    }                                                                                                   \
    slot_attn_kernel<KK><<<static_cast<unsigned>(B), kAttnThreads, attn_smem, s>>>(a);                  \
    break;
      switch (P.K) {
        RLSB_ATTN(1) RLSB_ATTN(2) RLSB_ATTN(3) RLSB_ATTN(4) RLSB_ATTN(5) RLSB_ATTN(6) RLSB_ATTN(7) RLSB_ATTN(8)
        default: return -31;
      }
#undef RLSB_ATTN
      count_launch();
      RLSB_CUDA(cudaGetLastError());
    }
    // slots = GRUCell(updates, slots_prev)        slot_attention.py:75
    {
      PackSeg seg{0, 0, dim};
      RLSB_TRY(launch_pack(s_prev, dim, static_cast<int>(BK), sprevp_img, 128, W.bk_pad, dim, 1, &seg, s));
      GemmParams g1 = gemm(P.ih, updp_img, static_cast<int>(BK));
      g1.out_f32 = gi_buf; g1.ldo = W.ld_g;
      RLSB_TRY(launch_gemm(g1, EPI_PLAIN, s));
      GemmParams g2 = gemm(P.hh, sprevp_img, static_cast<int>(BK));
      g2.out_f32 = gh_buf; g2.ldo = W.ld_g;
      RLSB_TRY(launch_gemm(g2, EPI_PLAIN, s));
      slot_gru_kernel<<<grid_for(BK * dim, 256, 148 * 8), 256, 0, s>>>(gi_buf, gh_buf, W.ld_g, s_prev, BK, dim,
                                                                       snew_buf);
      count_launch();
      RLSB_CUDA(cudaGetLastError());
    }
    // slots = slots + W2 ReLU(W1 LN(slots) + b1) + b2           slot_attention.py:76
    {
      RLSB_TRY(ln_pack(snew_buf, nullptr, nullptr, BK, W.bk_pad, pf(P.ln_2_g), pf(P.ln_2_b), sn2_img));
      GemmParams g1 = gemm(P.m1, sn2_img, static_cast<int>(BK));
      g1.act = ACT_RELU; g1.out_bf16 = hid_img; g1.out_kpad = 4 * dim;
      RLSB_TRY(launch_gemm(g1, EPI_LN_ACT, s));
      GemmParams g2 = gemm(P.m2, hid_img, static_cast<int>(BK));
      g2.out_f32 = f32(W.mlp); g2.ldo = dim;
      RLSB_TRY(launch_gemm(g2, EPI_PLAIN, s));
      add_rows_kernel<<<grid_for(BK * dim, 256, 148 * 8), 256, 0, s>>>(snew_buf, f32(W.mlp), BK * dim, s_out);
      count_launch();
      RLSB_CUDA(cudaGetLastError());
    }
    cur ^= 1;
  }
  return 0;
}

extern "C" int rlsb_slot_attention_fwd(const rlsb_slot_cfg* cfg, const void* packed, int64_t B, const float* X,
                                       const float* prev_slots, float* out_slots, float* out_attn, void* workspace,
                                       void* stream_) {
  return slot_fwd_impl(cfg, packed, B, X, prev_slots, out_slots, out_attn, workspace, nullptr, stream_);
}

extern "C" size_t rlsb_slot_attention_tape_bytes(const rlsb_slot_cfg* cfg, int64_t B) {
  SlotPlan P;
  if (!cfg || B <= 0 || make_slot_plan(*cfg, P) != 0) return 0;
  SlotWs W;
  make_slot_ws(P, B, W);
  SlotTape T;
  make_slot_tape(P, W, T);
  return T.bytes;
}

extern "C" int rlsb_slot_attention_fwd_tape(const rlsb_slot_cfg* cfg, const void* packed, int64_t B, const float* X,
                                            const float* prev_slots, float* out_slots, float* out_attn, void* tape,
                                            void* workspace, void* stream_) {
  if (!tape) return -1;
  return slot_fwd_impl(cfg, packed, B, X, prev_slots, out_slots, out_attn, workspace, tape, stream_);
}

// =============================================================================================
// K3 backward (autograd of SlotAttention.forward, vision/slot_attention.py:52-77)
// =============================================================================================
namespace rlsb {
namespace {

// ---- attention backward, one CTA per frame ---------------------------------------------------
struct AttnBwdArgs {
  const __nv_bfloat16* kv;   // packed [BT_pad x 2*dim]
  const float* q;            // [B*K][dim]
  const float* dupd;         // [B*K][dim]  d loss / d updates
  int T, K, dim;
  float scale, eps;
  __nv_bfloat16* dq_packed;  // packed [BK_pad x dim]  d loss / d q
  float* dkv;                // [BT][2*dim] fp32, accumulated over iterations (+=)
};

template <int K>
__global__ void __launch_bounds__(kAttnThreads) slot_attn_bwd_kernel(const AttnBwdArgs a) {
  extern __shared__ float sm[];
  const int dim = a.dim, T = a.T;
  const int chunks = dim >> 3;
  float* sq = sm;                         // [K][dim]
  float* sdu = sq + K * dim;              // [K][dim]
  float* ss = sdu + K * dim;              // [K][T] softmax over slots
  float* sat = ss + K * T;                // [K][T] normalised attention
  float* sda = sat + K * T;               // [K][T] d attn -> d logits
  float* sred = sda + K * T;              // [warps][K][dim]
  float* srs = sred + kAttnWarps * K * dim;  // [K] row sums
  float* sc = srs + kMaxSlots;            // [K]
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < K * dim; i += kAttnThreads) {
    sq[i] = a.q[static_cast<size_t>(b) * K * dim + i];
    sdu[i] = a.dupd[static_cast<size_t>(b) * K * dim + i];
  }
  __syncthreads();
  const int kpad = 2 * dim;
  // (1) recompute the forward attention: logits, softmax over slots, + eps
  for (int j = warp; j < T; j += kAttnWarps) {
    const size_t row = static_cast<size_t>(b) * T + j;
    float acc[K];
#pragma unroll
    for (int i = 0; i < K; ++i) acc[i] = 0.f;
    for (int c = lane; c < chunks; c += 32) {
      float kf[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(a.kv + packed_index(row, static_cast<size_t>(c * 8),
                                                                       static_cast<size_t>(kpad), kTileM))), kf);
#pragma unroll
      for (int i = 0; i < K; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[i] = fmaf(sq[i * dim + c * 8 + e], kf[e], acc[i]);
    }
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    if (lane == 0) {
      float mx = acc[0] * a.scale;
#pragma unroll
      for (int i = 1; i < K; ++i) mx = fmaxf(mx, acc[i] * a.scale);
      float e[K], den = 0.f;
#pragma unroll
      for (int i = 0; i < K; ++i) {
        e[i] = __expf(acc[i] * a.scale - mx);
        den += e[i];
      }
#pragma unroll
      for (int i = 0; i < K; ++i) {
        ss[i * T + j] = e[i] / den;
        sat[i * T + j] = e[i] / den + a.eps;
      }
    }
  }
  __syncthreads();
  if (warp < K) {
    float s = 0.f;
    for (int j = lane; j < T; j += 32) s += sat[warp * T + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) srs[warp] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * T; i += kAttnThreads) sat[i] = sat[i] / srs[i / T];
  __syncthreads();
  // (2) stream v: d attn_ij = dupd_i . v_j ;  d v_j = sum_i attn_ij dupd_i
  for (int j = warp; j < T; j += kAttnWarps) {
    const size_t row = static_cast<size_t>(b) * T + j;
    float w[K], dot[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
      w[i] = sat[i * T + j];
      dot[i] = 0.f;
    }
    for (int c = lane; c < chunks; c += 32) {
      float vf[8], dv[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(a.kv + packed_index(row, static_cast<size_t>(dim + c * 8),
                                                                       static_cast<size_t>(kpad), kTileM))), vf);
#pragma unroll
      for (int e = 0; e < 8; ++e) dv[e] = 0.f;
#pragma unroll
      for (int i = 0; i < K; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float du = sdu[i * dim + c * 8 + e];
          dot[i] = fmaf(du, vf[e], dot[i]);
          dv[e] = fmaf(w[i], du, dv[e]);
        }
      float4* po = reinterpret_cast<float4*>(a.dkv + row * kpad + dim + c * 8);
      float4 o0 = po[0], o1 = po[1];
      o0.x += dv[0]; o0.y += dv[1]; o0.z += dv[2]; o0.w += dv[3];
      o1.x += dv[4]; o1.y += dv[5]; o1.z += dv[6]; o1.w += dv[7];
      po[0] = o0; po[1] = o1;
    }
#pragma unroll
    for (int i = 0; i < K; ++i) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dot[i] += __shfl_xor_sync(0xffffffffu, dot[i], o);
      if (lane == 0) sda[i * T + j] = dot[i];
    }
  }
  __syncthreads();
  // (3) through the token renormalisation and the softmax over slots
  if (warp < K) {
    float s = 0.f;
    for (int j = lane; j < T; j += 32) s = fmaf(sat[warp * T + j], sda[warp * T + j], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) sc[warp] = s;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < T; j += kAttnThreads) {
    float da[K], t = 0.f;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      da[i] = (sda[i * T + j] - sc[i]) / srs[i];       // d loss / d (softmax + eps)
      t = fmaf(ss[i * T + j], da[i], t);
    }
#pragma unroll
    for (int i = 0; i < K; ++i) sda[i * T + j] = ss[i * T + j] * (da[i] - t) * a.scale;   // d loss / d (q_i . k_j)
  }
  __syncthreads();
  // (4) stream k: d q_i = sum_j dl_ij k_j ;  d k_j = sum_i dl_ij q_i
  {
    float acc[K][2][8];
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[i][u][e] = 0.f;
    for (int j = warp; j < T; j += kAttnWarps) {
      const size_t row = static_cast<size_t>(b) * T + j;
      float dl[K];
#pragma unroll
      for (int i = 0; i < K; ++i) dl[i] = sda[i * T + j];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c = lane + 32 * u;
        if (c < chunks) {
          float kf[8], dk[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(a.kv + packed_index(row, static_cast<size_t>(c * 8),
                                                                           static_cast<size_t>(kpad), kTileM))), kf);
#pragma unroll
          for (int e = 0; e < 8; ++e) dk[e] = 0.f;
#pragma unroll
          for (int i = 0; i < K; ++i)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              acc[i][u][e] = fmaf(dl[i], kf[e], acc[i][u][e]);
              dk[e] = fmaf(dl[i], sq[i * dim + c * 8 + e], dk[e]);
            }
          float4* po = reinterpret_cast<float4*>(a.dkv + row * kpad + c * 8);
          float4 o0 = po[0], o1 = po[1];
          o0.x += dk[0]; o0.y += dk[1]; o0.z += dk[2]; o0.w += dk[3];
          o1.x += dk[4]; o1.y += dk[5]; o1.z += dk[6]; o1.w += dk[7];
          po[0] = o0; po[1] = o1;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c = lane + 32 * u;
        if (c < chunks) {
#pragma unroll
          for (int e = 0; e < 8; ++e) sred[(warp * K + i) * dim + c * 8 + e] = acc[i][u][e];
        }
      }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < K * chunks; idx += kAttnThreads) {
    const int i = idx / chunks, c = idx - i * chunks;
    float s[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] = 0.f;
    for (int wv = 0; wv < kAttnWarps; ++wv) {
      const float* p = sred + (wv * K + i) * dim + c * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) s[e] += p[e];
    }
    const size_t r = static_cast<size_t>(b) * K + i;
    *reinterpret_cast<uint4*>(a.dq_packed + packed_index(r, static_cast<size_t>(c * 8), static_cast<size_t>(dim), kTileM)) =
        make_uint4(bf2(s[0], s[1]), bf2(s[2], s[3]), bf2(s[4], s[5]), bf2(s[6], s[7]));
  }
}

// ---- nn.GRUCell backward (gates recomputed from gi, gh) ---------------------------------------
__global__ void slot_gru_bwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh, long long ld,
                                    const float* __restrict__ s_prev, const float* __restrict__ d_snew, long long rows,
                                    int rows_pad, int dim, __nv_bfloat16* __restrict__ dgi_p,
                                    __nv_bfloat16* __restrict__ dgh_p, float* __restrict__ d_sprev) {
  const int chunks = dim >> 3;
  const long long total = static_cast<long long>(rows_pad) * chunks;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / chunks;
    const int c = static_cast<int>(i - r * chunks) * 8;
    float o[6][8];
#pragma unroll
    for (int g = 0; g < 6; ++g)
#pragma unroll
      for (int e = 0; e < 8; ++e) o[g][e] = 0.f;
    if (r < rows) {
      const float* a = gi + r * ld;
      const float* bb = gh + r * ld;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int d = c + e;
        const float rg = 1.0f / (1.0f + __expf(-(a[d] + bb[d])));
        const float zg = 1.0f / (1.0f + __expf(-(a[dim + d] + bb[dim + d])));
        const float bn = bb[2 * dim + d];
        const float n = tanhf(a[2 * dim + d] + rg * bn);
        const float g = d_snew[r * dim + d];
        const float sp = s_prev[r * dim + d];
        const float dn = g * (1.0f - zg) * (1.0f - n * n);
        const float dz = g * (sp - n) * zg * (1.0f - zg);
        const float dr = dn * bn * rg * (1.0f - rg);
        o[0][e] = dr; o[1][e] = dz; o[2][e] = dn;           // d gi
        o[3][e] = dr; o[4][e] = dz; o[5][e] = dn * rg;      // d gh
        d_sprev[r * dim + d] = g * zg;
      }
    }
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const size_t idx = packed_index(static_cast<size_t>(r), static_cast<size_t>(g * dim + c), static_cast<size_t>(3 * dim), kTileM);
      *reinterpret_cast<uint4*>(dgi_p + idx) = make_uint4(bf2(o[g][0], o[g][1]), bf2(o[g][2], o[g][3]),
                                                          bf2(o[g][4], o[g][5]), bf2(o[g][6], o[g][7]));
      *reinterpret_cast<uint4*>(dgh_p + idx) = make_uint4(bf2(o[3 + g][0], o[3 + g][1]), bf2(o[3 + g][2], o[3 + g][3]),
                                                          bf2(o[3 + g][4], o[3 + g][5]), bf2(o[3 + g][6], o[3 + g][7]));
    }
  }
}

// ---- LayerNorm backward: dx = add0 + add1 + rstd (g - mean(g) - x_hat mean(g x_hat)), g = dy * gamma ----------
// one warp per row (grid-stride); per-block partial d_gamma / d_beta -> part[block][2][C]; C % 8 == 0, C <= 512
constexpr int kLnBwdThreads = 256;
__global__ void __launch_bounds__(kLnBwdThreads) ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                               long long ld_dy, const float* __restrict__ gamma,
                                                               const float* __restrict__ add0,
                                                               const float* __restrict__ add1, long long rows, int C,
                                                               float eps, float* __restrict__ dx,
                                                               float* __restrict__ part) {
  __shared__ float sred[kLnBwdThreads / 32][2][512];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const int chunks = C >> 3;
  float ag[2][8], ab[2][8];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int e = 0; e < 8; ++e) ag[u][e] = ab[u][e] = 0.f;
  for (long long r = warp0; r < rows; r += nwarps) {
    float xv[2][8], gv[2][8];
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int ch = lane + 32 * u;
#pragma unroll
      for (int e = 0; e < 8; ++e) xv[u][e] = gv[u][e] = 0.f;
      if (ch < chunks) {
        const float4* px = reinterpret_cast<const float4*>(x + r * C + ch * 8);
        const float4 a0 = px[0], a1 = px[1];
        xv[u][0] = a0.x; xv[u][1] = a0.y; xv[u][2] = a0.z; xv[u][3] = a0.w;
        xv[u][4] = a1.x; xv[u][5] = a1.y; xv[u][6] = a1.z; xv[u][7] = a1.w;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          gv[u][e] = dy[r * ld_dy + ch * 8 + e];
          s += xv[u][e];
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / static_cast<float>(C);
    float m2 = 0.f;
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (lane + 32 * u < chunks) {
#pragma unroll
        for (int e = 0; e < 8; ++e) m2 = fmaf(xv[u][e] - mean, xv[u][e] - mean, m2);
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, o);
    const float rstd = 1.0f / sqrtf(m2 / static_cast<float>(C) + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int ch = lane + 32 * u;
      if (ch < chunks) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xh = (xv[u][e] - mean) * rstd;
          const float d = gv[u][e];
          ag[u][e] = fmaf(d, xh, ag[u][e]);
          ab[u][e] += d;
          const float g = d * __ldg(gamma + ch * 8 + e);
          xv[u][e] = xh;
          gv[u][e] = g;
          s1 += g;
          s2 = fmaf(g, xh, s2);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float m1 = s1 / static_cast<float>(C), mm2 = s2 / static_cast<float>(C);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int ch = lane + 32 * u;
      if (ch < chunks) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float v = rstd * (gv[u][e] - m1 - xv[u][e] * mm2);
          const size_t idx = static_cast<size_t>(r) * C + ch * 8 + e;
          if (add0) v += add0[idx];
          if (add1) v += add1[idx];
          dx[idx] = v;
        }
      }
    }
  }
  // per-block partial column sums (fixed order: warps 0..7)
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int ch = lane + 32 * u;
    if (ch < chunks) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        sred[warp][0][ch * 8 + e] = ag[u][e];
        sred[warp][1][ch * 8 + e] = ab[u][e];
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += kLnBwdThreads) {
    const int which = i / C, col = i - which * C;
    float acc = 0.f;
    for (int wv = 0; wv < kLnBwdThreads / 32; ++wv) acc += sred[wv][which][col];
    part[(static_cast<size_t>(blockIdx.x) * 2 + which) * C + col] = acc;
  }
}

__global__ void ln_param_reduce_kernel(const float* __restrict__ part, int blocks, int C, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * C) return;
  const int which = i / C, col = i - which * C;
  float acc = 0.f;
  for (int b = 0; b < blocks; ++b) acc += part[(static_cast<size_t>(b) * 2 + which) * C + col];
  float* dst = which == 0 ? dgamma : dbeta;
  dst[col] = accumulate ? dst[col] + acc : acc;
}

struct SlotBwdWs {
  size_t dmlp, dpre1, tmp1, dsnew, dgi, dgh, dsdir, dupd, tmp2, dq, tmp3, ds[2], dkv, dkvp, dxn, lnpart, partial;
  size_t bytes;
};
constexpr int kLnBwdBlocks = 148 * 2;

int make_slot_bwd_ws(const SlotPlan& P, const SlotWs& W, long long B, SlotBwdWs& S) {
  const size_t bk = static_cast<size_t>(W.bk_pad), bt = static_cast<size_t>(W.bt_pad);
  const size_t dim = P.dim;
  size_t cur = 0;
  S.dmlp = place(cur, bk * dim * 2);
  S.dpre1 = place(cur, bk * 4 * dim * 2);
  S.tmp1 = place(cur, bk * dim * 4);
  S.dsnew = place(cur, bk * dim * 4);
  S.dgi = place(cur, bk * 3 * dim * 2);
  S.dgh = place(cur, bk * 3 * dim * 2);
  S.dsdir = place(cur, bk * dim * 4);
  S.dupd = place(cur, bk * dim * 4);
  S.tmp2 = place(cur, bk * dim * 4);
  S.dq = place(cur, bk * dim * 2);
  S.tmp3 = place(cur, bk * dim * 4);
  for (int i = 0; i < 2; ++i) S.ds[i] = place(cur, bk * dim * 4);
  S.dkv = place(cur, bt * 2 * dim * 4);
  S.dkvp = place(cur, bt * 2 * dim * 2);
  S.dxn = place(cur, bt * dim * 4);
  S.lnpart = place(cur, static_cast<size_t>(kLnBwdBlocks) * 2 * dim * 4);
  // weight-gradient partial tiles: the largest contraction decides
  size_t pmax = 0;
  const int shapes[6][3] = {{static_cast<int>(dim), 4 * static_cast<int>(dim), W.bk_pad},       // dW2: dY dim, X 4dim
                            {4 * static_cast<int>(dim), static_cast<int>(dim), W.bk_pad},       // dW1
                            {3 * static_cast<int>(dim), static_cast<int>(dim), W.bk_pad},       // dW_ih / dW_hh
                            {static_cast<int>(dim), static_cast<int>(dim), W.bk_pad},           // dW_q
                            {2 * static_cast<int>(dim), static_cast<int>(dim), W.bt_pad},       // dW_kv
                            {0, 0, 0}};
  for (int i = 0; i < 5; ++i) {
    WgradParams wp{};
    wp.n_tiles = shapes[i][0] / 64; wp.G = 1; wp.m_tiles = shapes[i][2] / 128;
    wp.n_seg = 2; wp.x_ktiles[0] = shapes[i][1] / 64; wp.x_ktiles[1] = 1;
    const int e = plan_wgrad(wp);
    if (e != 0) return e;
    const size_t b = wgrad_partial_bytes(wp);
    if (b > pmax) pmax = b;
  }
  S.partial = place(cur, pmax);
  S.bytes = rus(cur, 1024);
  (void)B;
  return 0;
}

}  // namespace
}  // namespace rlsb

extern "C" size_t rlsb_slot_attention_bwd_workspace_bytes(const rlsb_slot_cfg* cfg, int64_t B) {
  SlotPlan P;
  if (!cfg || B <= 0 || make_slot_plan(*cfg, P) != 0) return 0;
  SlotWs W;
  make_slot_ws(P, B, W);
  SlotBwdWs S;
  if (make_slot_bwd_ws(P, W, B, S) != 0) return 0;
  return S.bytes;
}

extern "C" int rlsb_slot_attention_bwd(const rlsb_slot_cfg* cfg, const void* packed, int64_t B, const float* X,
                                       const void* tape_, const float* d_out_slots, const rlsb_slot_grads* grads,
                                       float* dX, float* d_prev_slots, void* workspace, void* stream_) {
  if (!cfg || !packed || !X || !tape_ || !d_out_slots || !grads || !dX || !d_prev_slots || !workspace || B <= 0) return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  SlotPlan P;
  RLSB_TRY(make_slot_plan(*cfg, P));
  SlotWs W;
  make_slot_ws(P, B, W);
  SlotTape TP;
  make_slot_tape(P, W, TP);
  SlotBwdWs S;
  RLSB_TRY(make_slot_bwd_ws(P, W, B, S));
  const long long BT = B * P.T, BK = B * P.K;
  const int dim = P.dim;
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  const uint8_t* tape = static_cast<const uint8_t*>(tape_);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  auto bf = [&](size_t off) { return reinterpret_cast<__nv_bfloat16*>(ws + off); };
  auto f32 = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
  auto pf = [&](size_t off) { return reinterpret_cast<const float*>(pk + off); };
  auto tbf = [&](int it, size_t off) { return reinterpret_cast<const __nv_bfloat16*>(tape + TP.iter0 + TP.iter_bytes * it + off); };
  auto tf = [&](int it, size_t off) { return reinterpret_cast<const float*>(tape + TP.iter0 + TP.iter_bytes * it + off); };
  const __nv_bfloat16* ones = reinterpret_cast<const __nv_bfloat16*>(pk + P.ones_off);
  const int bk_tiles = W.bk_pad / 128, bt_tiles = W.bt_pad / 128;

  // dX-type GEMM: out = A[M x K] * Wt (transposed image, rows = in-features)
  auto dx_gemm = [&](const WPlan& T, const __nv_bfloat16* A, int M, int m_tiles) {
    GemmParams g{};
    g.A[0] = A; g.a_ktiles[0] = T.K / 64; g.n_seg = 1;
    g.W = reinterpret_cast<const __nv_bfloat16*>(pk + T.w_off);
    g.RB = T.RB; g.NB = T.NB; g.G = 1; g.M = M; g.m_tiles = m_tiles; g.N = T.N;
    return g;
  };
  // weight (+ bias) gradient: dW[n][k] (+)= dY^T X, db (+)= dY^T 1
  auto wgrad = [&](const __nv_bfloat16* dY, int n, const __nv_bfloat16* Xp, int k, int m_tiles, float* dW, float* db,
                   int accumulate) -> int {
    WgradParams wp{};
    wp.dY = dY; wp.n_tiles = n / 64; wp.G = 1; wp.m_tiles = m_tiles;
    wp.n_seg = 2;
    wp.X[0] = Xp; wp.x_ktiles[0] = k / 64; wp.x_mtile_stride[0] = static_cast<long long>(k) * 128;
    wp.X[1] = ones; wp.x_ktiles[1] = 1; wp.x_mtile_stride[1] = 0;
    wp.partial = f32(S.partial);
    RLSB_TRY(plan_wgrad(wp));
    RLSB_TRY(launch_wgrad(wp, s));
    WgradReduceParams rp{};
    rp.partial = wp.partial; rp.splits = wp.splits; rp.G = 1; rp.rows_pad = wp.n_slices * 128; rp.ld = wp.kt_total * 64;
    rp.w_dst[0] = dW; rp.b_dst[0] = db; rp.n_out[0] = n; rp.ld_dst = k; rp.n_seg = 1;
    rp.seg[0] = PackSeg{0, 0, k};
    rp.ones_col = k; rp.accumulate = accumulate;
    return launch_wgrad_reduce(rp, s);
  };
  auto ln_bwd = [&](const float* x, const float* dy, const float* gamma, const float* add0, const float* add1,
                    long long rows, float* dx, float* dgamma, float* dbeta, int accumulate) -> int {
    long long want = (rows + 7) / 8;
    const int blocks = static_cast<int>(want < kLnBwdBlocks ? (want < 1 ? 1 : want) : kLnBwdBlocks);
    ln_bwd_kernel<<<blocks, kLnBwdThreads, 0, s>>>(x, dy, dim, gamma, add0, add1, rows, dim, 1e-5f, dx, f32(S.lnpart));
    count_launch();
    RLSB_CUDA(cudaGetLastError());
    ln_param_reduce_kernel<<<(2 * dim + 255) / 256, 256, 0, s>>>(f32(S.lnpart), blocks, dim, dgamma, dbeta, accumulate);
    count_launch();
    return static_cast<int>(cudaGetLastError());
  };

  RLSB_CUDA(cudaMemsetAsync(ws + S.dkv, 0, static_cast<size_t>(BT) * 2 * dim * 4, s));
  static PerDeviceOnce attr_once[kMaxSlots + 1];
  unsigned long long dev_bit = 0;
  const size_t attn_smem = (2 * static_cast<size_t>(P.K) * dim + 3 * static_cast<size_t>(P.K) * P.T +
                            static_cast<size_t>(kAttnWarps) * P.K * dim + 2 * kMaxSlots) * sizeof(float);
  const float* ds = d_out_slots;   // d loss / d (slots leaving iteration it)
  int cur = 0;
  for (int it = P.iters - 1; it >= 0; --it) {
    const int acc = (it != P.iters - 1);   // later iterations ADD to the parameter gradients
    // ---- slots = snew + W2 ReLU(W1 LN2(snew) + b1) + b2 ---------------------------------------------
    {
      PackSeg seg{0, 0, dim};
      RLSB_TRY(launch_pack(ds, dim, static_cast<int>(BK), bf(S.dmlp), 128, W.bk_pad, dim, 1, &seg, s));
      RLSB_TRY(wgrad(bf(S.dmlp), dim, tbf(it, TP.hid), 4 * dim, bk_tiles, grads->mlp_w2, grads->mlp_b2, acc));
      GemmParams g = dx_gemm(P.t_m2, bf(S.dmlp), static_cast<int>(BK), bk_tiles);   // d hidden, ReLU' fused
      g.act = ACT_RELU; g.bwd_pre = tbf(it, TP.hid); g.out_bf16 = bf(S.dpre1); g.out_kpad = 4 * dim; g.group_major = 1;
      RLSB_TRY(launch_gemm(g, EPI_BWD, s));
      RLSB_TRY(wgrad(bf(S.dpre1), 4 * dim, tbf(it, TP.sn2), dim, bk_tiles, grads->mlp_w1, grads->mlp_b1, acc));
      GemmParams g2 = dx_gemm(P.t_m1, bf(S.dpre1), static_cast<int>(BK), bk_tiles);
      g2.out_f32 = f32(S.tmp1); g2.ldo = dim;
      RLSB_TRY(launch_gemm(g2, EPI_PLAIN, s));
      // d snew = d slots + LN2 backward
      RLSB_TRY(ln_bwd(tf(it, TP.snew), f32(S.tmp1), pf(P.ln_2_g), ds, nullptr, BK, f32(S.dsnew), grads->slots_norm2_g,
                      grads->slots_norm2_b, acc));
    }
    // ---- snew = GRUCell(updates, slots_prev) -------------------------------------------------------------
    {
      slot_gru_bwd_kernel<<<grid_for(static_cast<long long>(W.bk_pad) * (dim / 8), 256, 148 * 8), 256, 0, s>>>(
          tf(it, TP.gi), tf(it, TP.gh), W.ld_g, tf(it, TP.sprev), f32(S.dsnew), BK, W.bk_pad, dim, bf(S.dgi), bf(S.dgh),
          f32(S.dsdir));
      count_launch();
      RLSB_CUDA(cudaGetLastError());
      RLSB_TRY(wgrad(bf(S.dgi), 3 * dim, tbf(it, TP.updp), dim, bk_tiles, grads->gru_w_ih, grads->gru_b_ih, acc));
      RLSB_TRY(wgrad(bf(S.dgh), 3 * dim, tbf(it, TP.sprevp), dim, bk_tiles, grads->gru_w_hh, grads->gru_b_hh, acc));
      GemmParams g1 = dx_gemm(P.t_ih, bf(S.dgi), static_cast<int>(BK), bk_tiles);
      g1.out_f32 = f32(S.dupd); g1.ldo = dim;
      RLSB_TRY(launch_gemm(g1, EPI_PLAIN, s));
      GemmParams g2 = dx_gemm(P.t_hh, bf(S.dgh), static_cast<int>(BK), bk_tiles);
      g2.out_f32 = f32(S.tmp2); g2.ldo = dim;
      RLSB_TRY(launch_gemm(g2, EPI_PLAIN, s));
    }
    // ---- updates = attention(q, k, v) ------------------------------------------------------------------------
    {
      AttnBwdArgs a{reinterpret_cast<const __nv_bfloat16*>(tape + TP.kv), tf(it, TP.q), f32(S.dupd), P.T, P.K, dim,
                    1.0f / sqrtf(static_cast<float>(dim)), 1e-8f, bf(S.dq), f32(S.dkv)};
      if (BK != W.bk_pad)   // padding rows of the d q operand image
        RLSB_CUDA(cudaMemsetAsync(ws + S.dq + static_cast<size_t>(bk_tiles - 1) * 128 * dim * 2, 0,
                                  static_cast<size_t>(128) * dim * 2, s));
#define RLSB_ATTNB(KK)                                                                                      \
  case KK:                                                                                                  \
    if (attr_once[KK].need(dev_bit)) {                                                                      \
      RLSB_CUDA(cudaFuncSetAttribute(slot_attn_bwd_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     200 * 1024));                                                          \
      attr_once[KK].done(dev_bit);                                                                          \
    }                                                                                                       \
    slot_attn_bwd_kernel<KK><<<static_cast<unsigned>(B), kAttnThreads, attn_smem, s>>>(a);                  \
    break;
      switch (P.K) {
        RLSB_ATTNB(1) RLSB_ATTNB(2) RLSB_ATTNB(3) RLSB_ATTNB(4) RLSB_ATTNB(5) RLSB_ATTNB(6) RLSB_ATTNB(7) RLSB_ATTNB(8)
        default: return -31;
      }
#undef RLSB_ATTNB
      count_launch();
      RLSB_CUDA(cudaGetLastError());
    }
    // ---- q = W_q LN_s(slots_prev) ------------------------------------------------------------------------------
    {
      RLSB_TRY(wgrad(bf(S.dq), dim, tbf(it, TP.sn), dim, bk_tiles, grads->slots_proj_w, nullptr, acc));
      GemmParams g = dx_gemm(P.t_q, bf(S.dq), static_cast<int>(BK), bk_tiles);
      g.out_f32 = f32(S.tmp3); g.ldo = dim;
      RLSB_TRY(launch_gemm(g, EPI_PLAIN, s));
      // d slots_prev = GRU direct path + W_hh path + LN_s backward
      float* ds_next = (it == 0) ? d_prev_slots : f32(S.ds[cur]);
      RLSB_TRY(ln_bwd(tf(it, TP.sprev), f32(S.tmp3), pf(P.ln_s_g), f32(S.dsdir), f32(S.tmp2), BK, ds_next,
                      grads->slots_norm_g, grads->slots_norm_b, acc));
      ds = ds_next;
      cur ^= 1;
    }
  }
  // ---- k, v = W_kv LN_in(X) ---------------------------------------------------------------------------------------
  {
    PackSeg seg{0, 0, 2 * dim};
    RLSB_TRY(launch_pack(f32(S.dkv), 2 * dim, static_cast<int>(BT), bf(S.dkvp), 128, W.bt_pad, 2 * dim, 1, &seg, s));
    RLSB_TRY(wgrad(bf(S.dkvp), 2 * dim, reinterpret_cast<const __nv_bfloat16*>(tape + TP.xn), dim, bt_tiles,
                   grads->inputs_proj_w, nullptr, 0));
    GemmParams g = dx_gemm(P.t_kv, bf(S.dkvp), static_cast<int>(BT), bt_tiles);
    g.out_f32 = f32(S.dxn); g.ldo = dim;
    RLSB_TRY(launch_gemm(g, EPI_PLAIN, s));
    RLSB_TRY(ln_bwd(X, f32(S.dxn), pf(P.ln_in_g), nullptr, nullptr, BT, dX, grads->inputs_norm_g, grads->inputs_norm_b, 0));
  }
  return 0;
}
