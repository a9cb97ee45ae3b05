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
__global__ void __launch_bounds__(kAttnThreads) slot_attn_kernel(const AttnArgs a) {
  extern __shared__ float sm[];
  const int dim = a.dim, T = a.T;
  const int chunks = dim >> 3;  // 16-byte chunks per k (or v) row
  float* sq = sm;                      // [K][dim]
  float* sattn = sq + K * dim;         // [K][T]
  float* sred = sattn + K * T;         // [warps][K][dim]
  float* srs = sred + kAttnWarps * K * dim;  // [K] row sums
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < K * dim; i += kAttnThreads) sq[i] = a.q[static_cast<size_t>(b) * K * dim + i];
  __syncthreads();
  const int kpad = 2 * dim;
  // ---- logits, softmax over slots -------------------------------------------------------------
  for (int j = warp; j < T; j += kAttnWarps) {
    const size_t row = static_cast<size_t>(b) * T + j;
    float acc[K];
#pragma unroll
    for (int i = 0; i < K; ++i) acc[i] = 0.f;
    for (int c = lane; c < chunks; c += 32) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(
          a.kv + packed_index(row, static_cast<size_t>(c * 8), static_cast<size_t>(kpad), kTileM)));
      float kf[8];
      unpack8(u, kf);
#pragma unroll
      for (int i = 0; i < K; ++i) {
        const float* qi = sq + i * dim + c * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[i] = fmaf(qi[e], kf[e], acc[i]);
      }
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
      for (int i = 0; i < K; ++i) sattn[i * T + j] = e[i] / den + a.eps;
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
  // ---- updates = attn . v  (each warp accumulates a token subset, then a block reduction) ----------
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
      float w[K];
#pragma unroll
      for (int i = 0; i < K; ++i) w[i] = sattn[i * T + j];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c = lane + 32 * u;
        if (c < chunks) {
          const uint4 x = __ldg(reinterpret_cast<const uint4*>(
              a.kv + packed_index(row, static_cast<size_t>(dim + c * 8), static_cast<size_t>(kpad), kTileM)));
          float vf[8];
          unpack8(x, vf);
#pragma unroll
          for (int i = 0; i < K; ++i)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[i][u][e] = fmaf(w[i], vf[e], acc[i][u][e]);
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
    float4* po = reinterpret_cast<float4*>(a.upd_out + r * dim + c * 8);
    po[0] = make_float4(s[0], s[1], s[2], s[3]);
    po[1] = make_float4(s[4], s[5], s[6], s[7]);
    *reinterpret_cast<uint4*>(a.upd_packed + packed_index(r, static_cast<size_t>(c * 8), static_cast<size_t>(dim), kTileM)) =
        make_uint4(bf2(s[0], s[1]), bf2(s[2], s[3]), bf2(s[4], s[5]), bf2(s[6], s[7]));
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
  const float* lsrc[6] = {p->inputs_norm_g, p->inputs_norm_b, p->slots_norm_g, p->slots_norm_b,
                          p->slots_norm2_g, p->slots_norm2_b};
  const size_t loff[6] = {P.ln_in_g, P.ln_in_b, P.ln_s_g, P.ln_s_b, P.ln_2_g, P.ln_2_b};
  for (int i = 0; i < 6; ++i) {
    if (!lsrc[i]) return -33;
    RLSB_CUDA(cudaMemcpyAsync(base + loff[i], lsrc[i], static_cast<size_t>(P.dim) * 4, cudaMemcpyDeviceToDevice, s));
  }
  return 0;
}

extern "C" int rlsb_slot_attention_fwd(const rlsb_slot_cfg* cfg, const void* packed, int64_t B, const float* X,
                                       const float* prev_slots, float* out_slots, float* out_attn, void* workspace,
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
  RLSB_TRY(ln_pack(X, nullptr, nullptr, BT, W.bt_pad, pf(P.ln_in_g), pf(P.ln_in_b), bf(W.xn)));
  {
    GemmParams g = gemm(P.kv, bf(W.xn), static_cast<int>(BT));
    g.act = ACT_NONE; g.out_bf16 = bf(W.kv); g.out_kpad = 2 * dim;
    RLSB_TRY(launch_gemm(g, EPI_LN_ACT, s));
  }
  RLSB_CUDA(cudaMemcpyAsync(f32(W.cur[0]), prev_slots, static_cast<size_t>(BK) * dim * 4, cudaMemcpyDeviceToDevice, s));
  static bool attr_done[kMaxSlots + 1] = {};
  const size_t attn_smem = (static_cast<size_t>(P.K) * dim + static_cast<size_t>(P.K) * P.T +
                            static_cast<size_t>(kAttnWarps) * P.K * dim + kMaxSlots) * sizeof(float);
  int cur = 0;
  for (int it = 0; it < P.iters; ++it) {
    float* s_prev = f32(W.cur[cur]);
    float* s_out = (it == P.iters - 1) ? out_slots : f32(W.cur[cur ^ 1]);
    // q = W_q LN(slots)             slot_attention.py:66-67
    RLSB_TRY(ln_pack(s_prev, nullptr, nullptr, BK, W.bk_pad, pf(P.ln_s_g), pf(P.ln_s_b), bf(W.sn)));
    {
      GemmParams g = gemm(P.q, bf(W.sn), static_cast<int>(BK));
      g.out_f32 = f32(W.q); g.ldo = dim;
      RLSB_TRY(launch_gemm(g, EPI_PLAIN, s));
    }
    // attention                     slot_attention.py:69-74
    {
      AttnArgs a{bf(W.kv), f32(W.q), P.T, P.K, dim, 1.0f / sqrtf(static_cast<float>(dim)), 1e-8f,
                 (it == P.iters - 1) ? out_attn : nullptr, f32(W.upd), bf(W.updp)};
#define RLSB_ATTN(KK)                                                                                   \
  case KK:                                                                                              \
    if (!attr_done[KK])                                                                                 \
      RLSB_CUDA(cudaFuncSetAttribute(slot_attn_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     200 * 1024));                                                      \
    attr_done[KK] = true;                                                                               \
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
      RLSB_TRY(launch_pack(s_prev, dim, static_cast<int>(BK), bf(W.sprevp), 128, W.bk_pad, dim, 1, &seg, s));
      GemmParams g1 = gemm(P.ih, bf(W.updp), static_cast<int>(BK));
      g1.out_f32 = f32(W.gi); g1.ldo = W.ld_g;
      RLSB_TRY(launch_gemm(g1, EPI_PLAIN, s));
      GemmParams g2 = gemm(P.hh, bf(W.sprevp), static_cast<int>(BK));
      g2.out_f32 = f32(W.gh); g2.ldo = W.ld_g;
      RLSB_TRY(launch_gemm(g2, EPI_PLAIN, s));
      slot_gru_kernel<<<grid_for(BK * dim, 256, 148 * 8), 256, 0, s>>>(f32(W.gi), f32(W.gh), W.ld_g, s_prev, BK, dim,
                                                                       f32(W.snew));
      count_launch();
      RLSB_CUDA(cudaGetLastError());
    }
    // slots = slots + W2 ReLU(W1 LN(slots) + b1) + b2           slot_attention.py:76
    {
      RLSB_TRY(ln_pack(f32(W.snew), nullptr, nullptr, BK, W.bk_pad, pf(P.ln_2_g), pf(P.ln_2_b), bf(W.sn2)));
      GemmParams g1 = gemm(P.m1, bf(W.sn2), static_cast<int>(BK));
      g1.act = ACT_RELU; g1.out_bf16 = bf(W.hid); g1.out_kpad = 4 * dim;
      RLSB_TRY(launch_gemm(g1, EPI_LN_ACT, s));
      GemmParams g2 = gemm(P.m2, bf(W.hid), static_cast<int>(BK));
      g2.out_f32 = f32(W.mlp); g2.ldo = dim;
      RLSB_TRY(launch_gemm(g2, EPI_PLAIN, s));
      add_rows_kernel<<<grid_for(BK * dim, 256, 148 * 8), 256, 0, s>>>(f32(W.snew), f32(W.mlp), BK * dim, s_out);
      count_launch();
      RLSB_CUDA(cudaGetLastError());
    }
    cur ^= 1;
  }
  return 0;
}
