// rlsb_kernels.cu — HBM-bound kernels of the imagination path: operand packing, LayerNorm /
// GRU-gate application, categorical sampling, head post-processing and the lambda-return scan.
#include <cstdlib>
#include "rlsb_kernels.cuh"

#include "rlsb_count.cuh"
#include "rlsb_detmath.h"
#include "rlsb_gemm.cuh"
#include "rlsb_ptx.cuh"
#include "rlsb_rowops.cuh"

namespace rlsb {

namespace {

using namespace rowops;

// x - float(bf16(x)): what the "lo" image of the split-operand contraction mode carries (exact in fp32)
__device__ __forceinline__ float bf16_residual(float x) { return x - __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ float act_apply(float x, int act) {
  if (act == ACT_ELU) return x > 0.f ? x : expm1f(x);
  if (act == ACT_RELU) return fmaxf(x, 0.f);
  return x;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ------------------------------------------------------------------------------------------
// pack: one thread per 16-byte destination chunk (8 bf16)
// ------------------------------------------------------------------------------------------
struct PackArgs {
  const float* src;
  long long ld_src;
  int rows_src;
  __nv_bfloat16* dst;
  int RB, rows_dst_pad, k_pad, n_seg;
  PackSeg seg[8];
  int perm_D = 0;   // > 0: GRU weight for the fused cell (EPI_GRU): destination row 192 nb + 64 gate + u <- source row
                    // gate * D + 64 nb + u (n-block nb holds [reset | candidate | update] of hidden units [64 nb, 64 nb + 64))
};
__host__ __device__ __forceinline__ long long gru_perm_row(long long r, int D) {
  const long long nb = r / 192;
  const int j = static_cast<int>(r - nb * 192);
  return static_cast<long long>(j >> 6) * D + nb * 64 + (j & 63);
}
__global__ void copy_gru_perm_kernel(const float* src, int D, float* dst, float fill) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 3 * D) dst[i] = src ? src[gru_perm_row(i, D)] : fill;
}

__global__ void pack_kernel(const PackArgs a) {
  const long long chunks_per_row = a.k_pad >> 3;
  const long long total = static_cast<long long>(a.rows_dst_pad) * chunks_per_row;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / chunks_per_row;
    const long long srow = a.perm_D ? gru_perm_row(row, a.perm_D) : row;
    const int k0 = static_cast<int>(i - row * chunks_per_row) << 3;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float x = 0.f;
      if (row < a.rows_src) {
        const int k = k0 + j;
        for (int s = 0; s < a.n_seg; ++s) {
          const int off = k - a.seg[s].dst_k0;
          if (off >= 0 && off < a.seg[s].len) {
            x = __ldg(a.src + srow * a.ld_src + a.seg[s].src_c0 + off);
            if (a.seg[s].part) x = bf16_residual(x);
          }
        }
      }
      v[j] = x;
    }
    uint4 pk = make_uint4(bf2(v[0], v[1]), bf2(v[2], v[3]), bf2(v[4], v[5]), bf2(v[6], v[7]));
    const size_t idx = packed_index(static_cast<size_t>(row), static_cast<size_t>(k0),
                                    static_cast<size_t>(a.k_pad), a.RB);
    *reinterpret_cast<uint4*>(a.dst + idx) = pk;
  }
}

// ---- multi-job variants (see batch_begin / batch_end) -----------------------------------------------------------
constexpr int kPackJobs = 20;
struct PackJobs {
  int n;
  long long first[kPackJobs + 1];   // prefix sums of 16-byte destination chunks
  PackArgs job[kPackJobs];
};
__global__ void pack_multi_kernel(const __grid_constant__ PackJobs J) {
  const long long total = J.first[J.n];
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    int j = 0;
    while (j + 1 < J.n && i >= J.first[j + 1]) ++j;
    const PackArgs& a = J.job[j];
    const long long li = i - J.first[j];
    const long long chunks_per_row = a.k_pad >> 3;
    const long long row = li / chunks_per_row;
    const long long srow = a.perm_D ? gru_perm_row(row, a.perm_D) : row;
    const int k0 = static_cast<int>(li - row * chunks_per_row) << 3;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float x = 0.f;
      if (row < a.rows_src) {
        const int k = k0 + e;
        for (int s = 0; s < a.n_seg; ++s) {
          const int off = k - a.seg[s].dst_k0;
          if (off >= 0 && off < a.seg[s].len) {
            x = __ldg(a.src + srow * a.ld_src + a.seg[s].src_c0 + off);
            if (a.seg[s].part) x = bf16_residual(x);
          }
        }
      }
      v[e] = x;
    }
    const size_t idx = packed_index(static_cast<size_t>(row), static_cast<size_t>(k0), static_cast<size_t>(a.k_pad), a.RB);
    *reinterpret_cast<uint4*>(a.dst + idx) = make_uint4(bf2(v[0], v[1]), bf2(v[2], v[3]), bf2(v[4], v[5]), bf2(v[6], v[7]));
  }
}

constexpr int kPadJobs = 96;
struct PadJob {
  const float* src;
  float* dst;
  int n, n_pad;
  float fill;
};
struct PadJobs {
  int n;
  int first[kPadJobs + 1];   // prefix sums of n_pad
  PadJob job[kPadJobs];
};
__global__ void copy_pad_multi_kernel(const __grid_constant__ PadJobs J) {
  const int total = J.first[J.n];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int lo = 0, hi = J.n;   // job j with first[j] <= i < first[j + 1]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (i >= J.first[mid]) lo = mid;
      else hi = mid;
    }
    const PadJob& a = J.job[lo];
    const int li = i - J.first[lo];
    a.dst[li] = (a.src && li < a.n) ? a.src[li] : a.fill;
  }
}
__global__ void copy_pad_single_kernel(const float* __restrict__ src, int n, float* __restrict__ dst, int n_pad, float fill) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) dst[i] = (src && i < n) ? src[i] : fill;
}

struct PackTArgs {
  const float* src;
  long long ld_src;
  int cols;                 // out-features of the Linear (columns of the transposed operand)
  __nv_bfloat16* dst;
  int RB, rows_dst_pad, k_pad;
  int dst_k0, k_len;        // only the K range [dst_k0, dst_k0 + k_len) is written (k_len % 8 == 0)
  int n_seg;
  PackSeg seg[3];           // row segments: dst_k0 = first padded ROW, src_c0 = first in-feature, len
};

__global__ void pack_transposed_kernel(const PackTArgs a) {
  const long long chunks = a.k_len >> 3;
  const long long total = static_cast<long long>(a.rows_dst_pad) * chunks;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    // consecutive threads take consecutive ROWS of the same chunk: coalesced reads of src[c][r..]
    const long long ch = i / a.rows_dst_pad;
    const int row = static_cast<int>(i - ch * a.rows_dst_pad);
    const int c0 = static_cast<int>(ch) << 3;   // column of the transposed operand relative to dst_k0
    int src_col = -1;
    for (int s = 0; s < a.n_seg; ++s) {
      const int off = row - a.seg[s].dst_k0;
      if (off >= 0 && off < a.seg[s].len) src_col = a.seg[s].src_c0 + off;
    }
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      v[j] = (src_col >= 0 && c0 + j < a.cols) ? __ldg(a.src + static_cast<long long>(c0 + j) * a.ld_src + src_col) : 0.f;
    const size_t idx = packed_index(static_cast<size_t>(row), static_cast<size_t>(a.dst_k0 + c0),
                                    static_cast<size_t>(a.k_pad), a.RB);
    *reinterpret_cast<uint4*>(a.dst + idx) = make_uint4(bf2(v[0], v[1]), bf2(v[2], v[3]), bf2(v[4], v[5]), bf2(v[6], v[7]));
  }
}

constexpr int kPackTJobs = 28;
struct PackTJobs {
  int n;
  long long first[kPackTJobs + 1];   // prefix sums of 16-byte destination chunks
  PackTArgs job[kPackTJobs];
};
__global__ void pack_transposed_multi_kernel(const __grid_constant__ PackTJobs J) {
  const long long total = J.first[J.n];
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    int j = 0;
    while (j + 1 < J.n && i >= J.first[j + 1]) ++j;
    const PackTArgs& a = J.job[j];
    const long long li = i - J.first[j];
    const long long ch = li / a.rows_dst_pad;
    const int row = static_cast<int>(li - ch * a.rows_dst_pad);
    const int c0 = static_cast<int>(ch) << 3;
    int src_col = -1;
    for (int s = 0; s < a.n_seg; ++s) {
      const int off = row - a.seg[s].dst_k0;
      if (off >= 0 && off < a.seg[s].len) src_col = a.seg[s].src_c0 + off;
    }
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e)
      v[e] = (src_col >= 0 && c0 + e < a.cols) ? __ldg(a.src + static_cast<long long>(c0 + e) * a.ld_src + src_col) : 0.f;
    const size_t idx = packed_index(static_cast<size_t>(row), static_cast<size_t>(a.dst_k0 + c0),
                                    static_cast<size_t>(a.k_pad), a.RB);
    *reinterpret_cast<uint4*>(a.dst + idx) = make_uint4(bf2(v[0], v[1]), bf2(v[2], v[3]), bf2(v[4], v[5]), bf2(v[6], v[7]));
  }
}

struct LaunchBatch {
  bool active = false;
  cudaStream_t stream = nullptr;
  PackJobs packs{};
  PackTJobs packts{};
  PadJobs pads{};
};
thread_local LaunchBatch t_batch;

// ------------------------------------------------------------------------------------------
// LayerNorm statistics from the per-(row, n-block) partials (sum, sum of squares) the GEMM
// epilogue wrote: one warp works on one row, lane b fetches block b, shuffle-reduce.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void row_stats_warp(const float* stats, int NB, int m_pad, int m, int N,
                                               float eps, float& mean, float& rstd) {
  const float2* st = reinterpret_cast<const float2*>(stats);
  const int lane = threadIdx.x & 31;
  float s = 0.f, q = 0.f;
  for (int b = lane; b < NB; b += 32) {
    const float2 v = __ldg(&st[static_cast<size_t>(b) * m_pad + m]);
    s += v.x;
    q += v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  const float inv_n = 1.0f / static_cast<float>(N);
  mean = s * inv_n;
  rstd = 1.0f / sqrtf(fmaxf(q * inv_n - mean * mean, 0.f) + eps);
}

struct LnActArgs {
  const float* scratch;
  long long ld;
  const float* stats;
  int NB, RB, M, m_pad, N;
  const float* gamma;
  const float* beta;
  float eps;
  int act;
  __nv_bfloat16* out;
  int out_kpad;
};

__device__ __forceinline__ void load8(const float* p, bool vec, int valid, float (&v)[8]) {
  if (vec && valid >= 8) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = j < valid ? p[j] : 0.f;
  }
}
__device__ __forceinline__ void load8_ro(const float* p, bool vec, int valid, float (&v)[8]) {
  if (vec && valid >= 8) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = j < valid ? __ldg(p + j) : 0.f;
  }
}

// LayerNorm + activation over the fp32 pre-activations a split-row GEMM left in `scratch`: one warp per row.
// A lane owns the 8-column chunks lane, lane + 32, ... (one 16-byte store each into the packed operand image); up
// to four chunks (eight 16-byte loads) are in flight per lane before the row statistics are reduced — HBM-bound:
// 4 B read + 2 B written per element.
__global__ void __launch_bounds__(256, 4) ln_act_kernel(const LnActArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const int cpr = a.out_kpad >> 3;
  const int lane = threadIdx.x & 31;
  const int warp0 = static_cast<int>((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int nwarps = static_cast<int>((gridDim.x * blockDim.x) >> 5);
  const bool vec = ((a.ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.scratch) & 15) == 0);
  const bool pvec = a.gamma && ((reinterpret_cast<uintptr_t>(a.gamma) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(a.beta) & 15) == 0);
  for (int m = warp0; m < a.m_pad; m += nwarps) {
    const bool row_ok = m < a.M;
    const float* src = a.scratch + static_cast<size_t>(m) * a.ld;
    float mean = 0.f, rstd = 1.f;
    for (int cb = 0; cb < cpr; cb += 128) {   // batches of 4 chunks per lane
      float v[4][8];
      int valid[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c0 = (cb + 32 * i + lane) << 3;
        valid[i] = (row_ok && cb + 32 * i + lane < cpr) ? max(0, min(8, a.N - c0)) : 0;
        if (valid[i] > 0) load8(src + c0, vec, valid[i], v[i]);
      }
      if (cb == 0 && row_ok && a.gamma) row_stats_warp(a.stats, a.NB, a.m_pad, m, a.N, a.eps, mean, rstd);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int chunk = cb + 32 * i + lane;
        if (chunk >= cpr) continue;
        const int c0 = chunk << 3;
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = 0.f;
        if (valid[i] > 0) {
          float g[8], b[8];
          if (a.gamma) {
            load8_ro(a.gamma + c0, pvec, valid[i], g);
            load8_ro(a.beta + c0, pvec, valid[i], b);
            const float nmr = -mean * rstd;
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = fmaf(fmaf(v[i][j], rstd, nmr), g[j], b[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = v[i][j];
          }
          if (a.act == ACT_ELU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = fmaxf(y[j], ex2_approx(1.4426950408889634f * fminf(y[j], 0.f)) - 1.0f);
          } else if (a.act == ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = fmaxf(y[j], 0.f);
          }
          if (valid[i] < 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j >= valid[i]) y[j] = 0.f;
          }
        }
        const size_t idx = packed_index(static_cast<size_t>(m), static_cast<size_t>(c0),
                                        static_cast<size_t>(a.out_kpad), kTileM);
        *reinterpret_cast<uint4*>(a.out + idx) = make_uint4(bf2(y[0], y[1]), bf2(y[2], y[3]), bf2(y[4], y[5]), bf2(y[6], y[7]));
      }
    }
  }
}

struct GruArgs {
  const float* scratch;
  long long ld;
  const float* stats;
  int NB, RB, M, m_pad, D;
  const float* gamma;
  const float* beta;
  float eps, update_bias;
  const float* h_prev;
  long long ld_h;
  float* h_next;
  long long ld_hn;
  __nv_bfloat16* h_packed;
  int kpad;
};

// generic path (D or a leading dimension not a multiple of 4): one warp per (row, 32 eight-column chunks)
__global__ void gru_gate_kernel_generic(const GruArgs a) {
  const int chunks_per_row = a.kpad >> 3;
  const int gpr = (chunks_per_row + 31) >> 5;
  const long long total = static_cast<long long>(a.m_pad) * gpr;
  const int D = a.D;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long wi = warp0; wi < total; wi += nwarps) {
    const int m = static_cast<int>(wi / gpr);
    const int chunk = static_cast<int>(wi - static_cast<long long>(m) * gpr) * 32 + lane;
    const int c0 = chunk << 3;
    const bool row_ok = m < a.M;
    const bool act = row_ok && chunk < chunks_per_row && c0 < D;
    const int valid = act ? min(8, D - c0) : 0;
    float pr[8], pc[8], pu[8], hp[8], gr[8], gc[8], gu[8], br[8], bc[8], bu[8];
    if (act) {
      const float* src = a.scratch + static_cast<size_t>(m) * a.ld + c0;
      load8(src, false, valid, pr);
      load8(src + D, false, valid, pc);
      load8(src + 2 * D, false, valid, pu);
      load8(a.h_prev + static_cast<size_t>(m) * a.ld_h + c0, false, valid, hp);
      load8(a.gamma + c0, false, valid, gr);
      load8(a.gamma + D + c0, false, valid, gc);
      load8(a.gamma + 2 * D + c0, false, valid, gu);
      load8(a.beta + c0, false, valid, br);
      load8(a.beta + D + c0, false, valid, bc);
      load8(a.beta + 2 * D + c0, false, valid, bu);
    }
    float mean = 0.f, rstd = 1.f;
    if (row_ok) row_stats_warp(a.stats, a.NB, a.m_pad, m, 3 * D, a.eps, mean, rstd);
    if (chunk >= chunks_per_row) continue;
    float y[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = 0.f;
    if (act) {
      float* hn = a.h_next + static_cast<size_t>(m) * a.ld_hn + c0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < valid) {
          const float r = fast_sigmoid((pr[j] - mean) * rstd * gr[j] + br[j]);
          const float cand = fast_tanh(r * ((pc[j] - mean) * rstd * gc[j] + bc[j]));
          const float u = fast_sigmoid((pu[j] - mean) * rstd * gu[j] + bu[j] + a.update_bias);
          y[j] = u * cand + (1.0f - u) * hp[j];
          hn[j] = y[j];
        }
      }
    }
    const size_t idx = packed_index(static_cast<size_t>(m), static_cast<size_t>(c0),
                                    static_cast<size_t>(a.kpad), kTileM);
    *reinterpret_cast<uint4*>(a.h_packed + idx) = make_uint4(bf2(y[0], y[1]), bf2(y[2], y[3]), bf2(y[4], y[5]), bf2(y[6], y[7]));
  }
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// GRU gates (common.py:69-81) on the fp32 pre-activations of the split-row GRU contraction: one warp per row, a lane
// owns the float4 column groups lane, lane + 32, ...; two groups (eight streaming 16-byte loads) are in flight per lane
// before the row statistics are reduced.  HBM-bound: 3·4 + 4 B read, 4 + 2 B written per hidden unit.
__global__ void __launch_bounds__(256) gru_gate_kernel(const GruArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const int D = a.D;
  const int lane = threadIdx.x & 31;
  const int warp0 = static_cast<int>((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int nwarps = static_cast<int>((gridDim.x * blockDim.x) >> 5);
  const int quads = a.kpad >> 2;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int m = warp0; m < a.m_pad; m += nwarps) {
    const bool row_ok = m < a.M;
    const float* src = a.scratch + static_cast<size_t>(m) * a.ld;
    const float* hp = a.h_prev + static_cast<size_t>(m) * a.ld_h;
    float* hn = a.h_next + static_cast<size_t>(m) * a.ld_hn;
    float mean = 0.f, rstd = 1.f;
    for (int qb = 0; qb < quads; qb += 64) {
      float4 pr[2], pc[2], pu[2], hv[2];
      bool ok[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c = (qb + 32 * i + lane) << 2;
        ok[i] = row_ok && c < D;
        pr[i] = pc[i] = pu[i] = hv[i] = zero4;
        if (ok[i]) {
          pr[i] = *reinterpret_cast<const float4*>(src + c);
          pc[i] = *reinterpret_cast<const float4*>(src + D + c);
          pu[i] = *reinterpret_cast<const float4*>(src + 2 * D + c);
          hv[i] = *reinterpret_cast<const float4*>(hp + c);
        }
      }
      if (qb == 0 && row_ok) row_stats_warp(a.stats, a.NB, a.m_pad, m, 3 * D, a.eps, mean, rstd);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int q = qb + 32 * i + lane;
        if (q >= quads) continue;
        const int c = q << 2;
        float y[4] = {0.f, 0.f, 0.f, 0.f};
        if (ok[i]) {
          const float4 gr = ldg4(a.gamma + c), gc = ldg4(a.gamma + D + c), gu = ldg4(a.gamma + 2 * D + c);
          const float4 br = ldg4(a.beta + c), bc = ldg4(a.beta + D + c), bu = ldg4(a.beta + 2 * D + c);
          const float vr[4] = {pr[i].x, pr[i].y, pr[i].z, pr[i].w}, vc[4] = {pc[i].x, pc[i].y, pc[i].z, pc[i].w};
          const float vu[4] = {pu[i].x, pu[i].y, pu[i].z, pu[i].w}, vh[4] = {hv[i].x, hv[i].y, hv[i].z, hv[i].w};
          const float wgr[4] = {gr.x, gr.y, gr.z, gr.w}, wgc[4] = {gc.x, gc.y, gc.z, gc.w}, wgu[4] = {gu.x, gu.y, gu.z, gu.w};
          const float wbr[4] = {br.x, br.y, br.z, br.w}, wbc[4] = {bc.x, bc.y, bc.z, bc.w}, wbu[4] = {bu.x, bu.y, bu.z, bu.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float r = fast_sigmoid((vr[j] - mean) * rstd * wgr[j] + wbr[j]);
            const float cand = fast_tanh(r * ((vc[j] - mean) * rstd * wgc[j] + wbc[j]));
            const float u = fast_sigmoid((vu[j] - mean) * rstd * wgu[j] + wbu[j] + a.update_bias);
            y[j] = u * cand + (1.0f - u) * vh[j];
          }
          *reinterpret_cast<float4*>(hn + c) = make_float4(y[0], y[1], y[2], y[3]);
        }
        const size_t idx = packed_index(static_cast<size_t>(m), static_cast<size_t>(c), static_cast<size_t>(a.kpad), kTileM);
        *reinterpret_cast<uint2*>(a.h_packed + idx) = make_uint2(bf2(y[0], y[1]), bf2(y[2], y[3]));
      }
    }
  }
}

// Latent sampling kernel: see rowops::sample_latent_items (rlsb_rowops.cuh)
__global__ void __launch_bounds__(256, 4) sample_latent_kernel(const SampleLatentArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  sample_latent_items<false>(a, 0, static_cast<long long>(a.M) * a.groups, warp0, nwarps);
}

// generic categorical (rows x classes) with explicit uniforms: indices only
__global__ void sample_categorical_kernel(const float* __restrict__ logits,
                                          const float* __restrict__ uniforms, long long rows,
                                          int classes, int32_t* __restrict__ idx_out) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= rows) return;
  const float* l = logits + i * classes;
  const float* u = uniforms + i * classes;
  float best = 0.f;
  int best_k = 0;
  for (int k = 0; k < classes; ++k) {
    const float s = __fadd_rn(__ldg(l + k), rlsb_gumbel(__ldg(u + k)));
    if (k == 0 || s > best) {
      best = s;
      best_k = k;
    }
  }
  idx_out[i] = best_k;
}

// ------------------------------------------------------------------------------------------
// head post-processing: reward / value / discount read-out + action draw (one thread per row)
// ------------------------------------------------------------------------------------------
__global__ void head_finish_kernel(const HeadFinishParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= p.m_pad) return;
  const bool valid = m < p.M;
  if (valid) {
    if (p.reward_out)   // g_reward < 0: the reward head was not evaluated for this step (rlsb_imagine_cfg::last_step_value_only)
      p.reward_out[m] = p.g_reward >= 0 ? p.head_out[p.g_reward * p.group_stride + static_cast<long long>(m) * p.ldo] : 0.f;
    if (p.g_critic >= 0 && p.value_out)
      p.value_out[m] = p.head_out[p.g_critic * p.group_stride + static_cast<long long>(m) * p.ldo];
    if (p.discount_out) {
      float d = 1.0f;
      if (p.g_discount >= 0 && !p.first_step) {
        // torch Bernoulli(logits).mode: (probs >= 0.5), NaN where probs == 0.5 (world_model.py:137)
        const float x = p.head_out[p.g_discount * p.group_stride + static_cast<long long>(m) * p.ldo];
        const float pr = sigmoidf_(x);
        d = pr > 0.5f ? 1.0f : (pr == 0.5f ? (p.nan_on_tie ? __int_as_float(0x7fc00000) : 1.0f) : 0.0f);
      }
      p.discount_out[m] = d;
    }
  }
  if (!p.want_action || p.g_actor < 0) return;
  const float* ao = p.head_out + p.g_actor * p.group_stride + static_cast<long long>(m) * p.ldo;
  // zero the padded packed row first (a_kpad is 64 for every shipped config)
  float act[64];
  for (int k = 0; k < p.a_kpad && k < 64; ++k) act[k] = 0.f;
  if (valid && p.precomp) {
    for (int k = 0; k < p.A; ++k) act[k] = p.precomp[static_cast<size_t>(m) * p.A + k];
    if (p.action_out)
      for (int k = 0; k < p.A; ++k) p.action_out[static_cast<size_t>(m) * p.A + k] = act[k];
  } else if (valid) {
    if (p.discrete) {
      float best = 0.f;
      int best_k = 0;
      for (int k = 0; k < p.A; ++k) {
        const float u = p.noise.explicit_noise
                            ? __ldg(p.noise.explicit_noise + static_cast<size_t>(m) * p.noise.ld + k)
                            : rlsb_noise_uniform(p.noise.seed_ptr ? __ldg(p.noise.seed_ptr) : p.noise.seed,
                                                 p.noise.row_offset + m, p.noise.step, 1u, k);
        const float s = __fadd_rn(ao[k], rlsb_gumbel(u));
        if (k == 0 || s > best) {
          best = s;
          best_k = k;
        }
        if (p.actor_raw_out) p.actor_raw_out[static_cast<size_t>(m) * p.A + k] = ao[k];
      }
      act[best_k] = 1.0f;
    } else {
      // TruncatedNormal(tanh(mu), 2*sigmoid(s/2)+0.1).rsample() == unclamped Normal.rsample
      // (dists.py:108-129 overrides only sample(); dreamer_v2.py:87 calls rsample)
      for (int k = 0; k < p.A; ++k) {
        const float mu = tanhf(ao[k]);
        const float sd = 2.0f * sigmoidf_(ao[p.A + k] * 0.5f) + 0.1f;
        const float e = noise_normal(p.noise, m, 1u, k);
        act[k] = mu + e * sd;
        if (p.actor_raw_out) {
          p.actor_raw_out[static_cast<size_t>(m) * 2 * p.A + k] = ao[k];
          p.actor_raw_out[static_cast<size_t>(m) * 2 * p.A + p.A + k] = ao[p.A + k];
        }
      }
    }
    if (p.action_out)
      for (int k = 0; k < p.A; ++k) p.action_out[static_cast<size_t>(m) * p.A + k] = act[k];
  }
  if (p.action_packed) {
    // slotted RSSM: every slot row (n * K + k) of the img_in operand sees the same action
    // (rssm_slots_attention.py:170: action.unsqueeze(2).repeat(...))
    const int rep = p.action_repeat > 1 ? p.action_repeat : 1;
    for (int r = 0; r < rep; ++r) {
      const size_t row = static_cast<size_t>(m) * rep + r;
      if (row >= static_cast<size_t>(p.action_rows_pad)) break;
      for (int c0 = 0; c0 < p.a_kpad; c0 += 8) {
        uint4 pk = make_uint4(bf2(act[c0], act[c0 + 1]), bf2(act[c0 + 2], act[c0 + 3]),
                              bf2(act[c0 + 4], act[c0 + 5]), bf2(act[c0 + 6], act[c0 + 7]));
        const size_t idx = packed_index(row, static_cast<size_t>(c0), static_cast<size_t>(p.a_kpad), kTileM);
        *reinterpret_cast<uint4*>(p.action_packed + idx) = pk;
        if (p.action_packed_lo) {
          float l[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) l[j] = bf16_residual(act[c0 + j]);
          *reinterpret_cast<uint4*>(p.action_packed_lo + idx) =
              make_uint4(bf2(l[0], l[1]), bf2(l[2], l[3]), bf2(l[4], l[5]), bf2(l[6], l[7]));
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// K2: lambda-return reverse scan + shifted cumprod weights + advantage
//   time-major layout (T, N): one thread per start state, operations in the reference's order
//   (ac.py:57-58: rs[i] + ds[i] * ((1-l)*vs[i+1] + l*V)), each rounded once => bit-identical
//   to the fp32 torch loop.
// ------------------------------------------------------------------------------------------
template <int VEC>
__global__ void lambda_return_tm_kernel(const float* __restrict__ r, const float* __restrict__ v,
                                        const float* __restrict__ d, int T, long long N, float c1,
                                        float c2, float* __restrict__ vs, float* __restrict__ w,
                                        float* __restrict__ adv) {
  const long long n0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * VEC;
  if (n0 >= N) return;
  const int H = T - 1;
  float V[VEC], vnext[VEC];
  if (VEC == 4) {
    const float4 t = *reinterpret_cast<const float4*>(v + static_cast<size_t>(H) * N + n0);
    V[0] = t.x; V[1 % VEC] = t.y; V[2 % VEC] = t.z; V[3 % VEC] = t.w;
  } else {
    V[0] = v[static_cast<size_t>(H) * N + n0];
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) vnext[j] = V[j];
  for (int t = H - 1; t >= 0; --t) {
    float rr[VEC], dd[VEC], vv[VEC];
    const size_t off = static_cast<size_t>(t) * N + n0;
    if (VEC == 4) {
      const float4 a = *reinterpret_cast<const float4*>(r + off);
      const float4 b = *reinterpret_cast<const float4*>(d + off);
      const float4 c = *reinterpret_cast<const float4*>(v + off);
      rr[0] = a.x; rr[1 % VEC] = a.y; rr[2 % VEC] = a.z; rr[3 % VEC] = a.w;
      dd[0] = b.x; dd[1 % VEC] = b.y; dd[2 % VEC] = b.z; dd[3 % VEC] = b.w;
      vv[0] = c.x; vv[1 % VEC] = c.y; vv[2 % VEC] = c.z; vv[3 % VEC] = c.w;
    } else {
      rr[0] = r[off]; dd[0] = d[off]; vv[0] = v[off];
    }
    float out[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float mix = __fadd_rn(__fmul_rn(c1, vnext[j]), __fmul_rn(c2, V[j]));
      out[j] = __fadd_rn(rr[j], __fmul_rn(dd[j], mix));
    }
    if (VEC == 4) {
      *reinterpret_cast<float4*>(vs + off) = make_float4(out[0], out[1 % VEC], out[2 % VEC], out[3 % VEC]);
    } else {
      vs[off] = out[0];
    }
    if (adv && t <= H - 2 && t + 1 <= H - 1) {
      // adv[t] = vs[t+1] - v[t]
      if (VEC == 4) {
        *reinterpret_cast<float4*>(adv + off) =
            make_float4(__fsub_rn(V[0], vv[0]), __fsub_rn(V[1 % VEC], vv[1 % VEC]),
                        __fsub_rn(V[2 % VEC], vv[2 % VEC]), __fsub_rn(V[3 % VEC], vv[3 % VEC]));
      } else {
        adv[off] = __fsub_rn(V[0], vv[0]);
      }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      V[j] = out[j];
      vnext[j] = vv[j];
    }
  }
  if (w) {
    // w[0] = 1, w[t] = w[t-1] * d[t-1]   (dreamer_v2.py:194-197)
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 1.0f;
    for (int t = 0; t < T; ++t) {
      const size_t off = static_cast<size_t>(t) * N + n0;
      if (VEC == 4) {
        *reinterpret_cast<float4*>(w + off) = make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]);
        if (t + 1 < T) {  // d[H] is never read: the caller may pass only H rows
          const float4 b = *reinterpret_cast<const float4*>(d + off);
          acc[0] = __fmul_rn(acc[0], b.x);
          acc[1 % VEC] = __fmul_rn(acc[1 % VEC], b.y);
          acc[2 % VEC] = __fmul_rn(acc[2 % VEC], b.z);
          acc[3 % VEC] = __fmul_rn(acc[3 % VEC], b.w);
        }
      } else {
        w[off] = acc[0];
        if (t + 1 < T) acc[0] = __fmul_rn(acc[0], d[off]);
      }
    }
  }
}

// batch-major layout (N, T), T <= 32: warp-shuffle reverse scan over affine maps
//   V_t = a_t + b_t * V_{t+1},  a_t = r_t + d_t*(1-l)*v_{t+1},  b_t = d_t*l
__global__ void lambda_return_bm_kernel(const float* __restrict__ r, const float* __restrict__ v,
                                        const float* __restrict__ d, int T, int Tp, long long N,
                                        float c1, float c2, float* __restrict__ vs,
                                        float* __restrict__ w, float* __restrict__ adv) {
  const int rows_per_warp = 32 / Tp;
  const long long warp_global = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int sub = lane / Tp;   // which row inside the warp
  const int t = lane - sub * Tp;
  const long long n = warp_global * rows_per_warp + sub;
  const int H = T - 1;
  const bool active = (n < N) && (t < T);
  float rt = 0.f, dt = 0.f, vt = 0.f;
  if (active) {
    const size_t off = static_cast<size_t>(n) * T + t;
    rt = r[off]; dt = d[off]; vt = v[off];
  }
  const float vnext = __shfl_down_sync(0xffffffffu, vt, 1);
  float a, b;
  if (t < H) {
    a = rt + dt * (c1 * vnext);
    b = dt * c2;
  } else {  // t == H holds the bootstrap, t > H padding
    a = (t == H) ? vt : 0.f;
    b = 0.f;
  }
  // shifted cumprod: inclusive scan of d, then shift by one
  float cp = dt;
  for (int off = 1; off < Tp; off <<= 1) {
    const float a2 = __shfl_down_sync(0xffffffffu, a, off);
    const float b2 = __shfl_down_sync(0xffffffffu, b, off);
    const float c_up = __shfl_up_sync(0xffffffffu, cp, off);
    if (t + off < Tp) {
      a = a + b * a2;
      b = b * b2;
    }
    if (t >= off) cp = cp * c_up;
  }
  float wt = __shfl_up_sync(0xffffffffu, cp, 1);
  if (t == 0) wt = 1.0f;
  const float Vn = __shfl_down_sync(0xffffffffu, a, 1);  // V_{t+1}
  if (active) {
    if (t < H) vs[static_cast<size_t>(n) * H + t] = a;
    if (w) w[static_cast<size_t>(n) * T + t] = wt;
    if (adv && t < H - 1) adv[static_cast<size_t>(n) * (H - 1) + t] = Vn - vt;
  }
}

// backward of the reverse scan = forward scan (time-major); gradients w.r.t. r, v, d
__global__ void lambda_return_bwd_kernel(const float* __restrict__ g_vs, const float* __restrict__ v,
                                         const float* __restrict__ d, const float* __restrict__ vs,
                                         int T, long long N, float c1, float c2,
                                         float* __restrict__ g_r, float* __restrict__ g_v,
                                         float* __restrict__ g_d) {
  const long long n = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (n >= N) return;
  const int H = T - 1;
  float G = 0.f;       // dL/dV_t (total)
  float carry = 0.f;   // contribution flowing from V_{t-1}
  if (g_v) g_v[n] = 0.f;
  for (int t = 0; t < H; ++t) {
    const size_t off = static_cast<size_t>(t) * N + n;
    G = g_vs[off] + carry;
    const float dt = d[off];
    const float vn = v[off + N];
    const float Vn = (t + 1 < H) ? vs[off + N] : vn;  // V_{t+1}; V_H = v_H
    if (g_r) g_r[off] = G;
    if (g_d) g_d[off] = G * (c1 * vn + c2 * Vn);
    if (g_v) {
      float gv = G * dt * c1;
      if (t + 1 == H) gv += G * dt * c2;  // bootstrap V_H = v_H
      g_v[off + N] = gv;
    }
    carry = G * dt * c2;
  }
  if (g_d) g_d[static_cast<size_t>(H) * N + n] = 0.f;
  if (g_r) g_r[static_cast<size_t>(H) * N + n] = 0.f;
}

inline int grid_for(long long total, int block, int cap = 148 * 16) {
  long long g = (total + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace

namespace {
int flush_packs() {
  PackJobs& J = t_batch.packs;
  if (J.n == 0) return 0;
  pack_multi_kernel<<<grid_for(J.first[J.n], 256), 256, 0, t_batch.stream>>>(J);
  count_launch();
  J.n = 0;
  return static_cast<int>(cudaGetLastError());
}
int flush_packts() {
  PackTJobs& J = t_batch.packts;
  if (J.n == 0) return 0;
  pack_transposed_multi_kernel<<<grid_for(J.first[J.n], 256, 148 * 8), 256, 0, t_batch.stream>>>(J);
  count_launch();
  J.n = 0;
  return static_cast<int>(cudaGetLastError());
}
int flush_pads() {
  PadJobs& J = t_batch.pads;
  if (J.n == 0) return 0;
  copy_pad_multi_kernel<<<grid_for(J.first[J.n], 256, 148 * 4), 256, 0, t_batch.stream>>>(J);
  count_launch();
  J.n = 0;
  return static_cast<int>(cudaGetLastError());
}
}  // namespace

int launch_pack(const float* src, long long ld_src, int rows_src, __nv_bfloat16* dst, int RB,
                int rows_dst_pad, int k_pad, int n_seg, const PackSeg* segs, cudaStream_t stream) {
  return launch_pack_perm(src, ld_src, rows_src, dst, RB, rows_dst_pad, k_pad, n_seg, segs, 0, stream);
}

int launch_copy_gru_perm(const float* src, int D, float* dst, float fill, cudaStream_t stream) {
  if (D <= 0 || (D % 64) != 0) return -1;
  copy_gru_perm_kernel<<<(3 * D + 255) / 256, 256, 0, stream>>>(src, D, dst, fill);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_pack_perm(const float* src, long long ld_src, int rows_src, __nv_bfloat16* dst, int RB, int rows_dst_pad, int k_pad,
                     int n_seg, const PackSeg* segs, int gru_perm_D, cudaStream_t stream) {
  if (n_seg < 0 || n_seg > 8 || (k_pad % 64) != 0 || RB <= 0 || (rows_dst_pad % RB) != 0) return -1;
  if (gru_perm_D && (RB != 192 || (gru_perm_D % 64) != 0 || rows_src != 3 * gru_perm_D || rows_dst_pad != rows_src)) return -1;
  PackArgs a{};
  a.src = src; a.ld_src = ld_src; a.rows_src = rows_src; a.dst = dst; a.RB = RB;
  a.rows_dst_pad = rows_dst_pad; a.k_pad = k_pad; a.n_seg = n_seg; a.perm_D = gru_perm_D;
  for (int s = 0; s < n_seg; ++s) a.seg[s] = segs[s];
  const long long total = static_cast<long long>(rows_dst_pad) * (k_pad / 8);
  if (t_batch.active && t_batch.stream == stream) {
    if (t_batch.packs.n == kPackJobs) {
      const int e = flush_packs();
      if (e != 0) return e;
    }
    PackJobs& J = t_batch.packs;
    if (J.n == 0) J.first[0] = 0;
    J.job[J.n] = a;
    J.first[J.n + 1] = J.first[J.n] + total;
    ++J.n;
    return 0;
  }
  pack_kernel<<<grid_for(total, 256), 256, 0, stream>>>(a);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_copy_pad(const float* src, int n, float* dst, int n_pad, float fill, cudaStream_t stream) {
  if (n_pad <= 0) return 0;
  if (t_batch.active && t_batch.stream == stream) {
    PadJobs& J = t_batch.pads;
    if (J.n == kPadJobs || (J.n > 0 && J.first[J.n] > (1 << 30) - n_pad)) {
      const int e = flush_pads();
      if (e != 0) return e;
    }
    if (J.n == 0) J.first[0] = 0;
    J.job[J.n] = PadJob{src, dst, n, n_pad, fill};
    J.first[J.n + 1] = J.first[J.n] + n_pad;
    ++J.n;
    return 0;
  }
  copy_pad_single_kernel<<<(n_pad + 255) / 256, 256, 0, stream>>>(src, n, dst, n_pad, fill);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

void batch_begin(cudaStream_t stream) {
  static const bool enabled = [] {
    const char* e = getenv("RLSB_PACK_BATCH");   // RLSB_PACK_BATCH=0: one launch per pack / pad job (A/B runs)
    return !(e && e[0] == '0');
  }();
  t_batch.active = enabled;
  t_batch.stream = stream;
  t_batch.packs.n = 0;
  t_batch.packts.n = 0;
  t_batch.pads.n = 0;
}

int batch_end() {
  const int e = flush_packs();
  const int e1 = flush_packts();
  const int e2 = flush_pads();
  t_batch.active = false;
  return e != 0 ? e : (e1 != 0 ? e1 : e2);
}

int launch_pack_transposed_seg(const float* src, long long ld_src, int cols, __nv_bfloat16* dst, int RB,
                               int rows_dst_pad, int k_pad, int dst_k0, int k_len, int n_seg, const PackSeg* segs,
                               cudaStream_t stream) {
  if (!src || !dst || (k_pad & 63) || (rows_dst_pad % RB) || (k_len & 7) || (dst_k0 & 7) || dst_k0 + k_len > k_pad ||
      cols > k_len || n_seg < 1 || n_seg > 3)
    return -1;
  PackTArgs a{src, ld_src, cols, dst, RB, rows_dst_pad, k_pad, dst_k0, k_len, n_seg, {}};
  for (int i = 0; i < n_seg; ++i) a.seg[i] = segs[i];
  const long long total = static_cast<long long>(rows_dst_pad) * (k_len >> 3);
  if (t_batch.active && t_batch.stream == stream) {
    if (t_batch.packts.n == kPackTJobs) {
      const int e = flush_packts();
      if (e != 0) return e;
    }
    PackTJobs& J = t_batch.packts;
    if (J.n == 0) J.first[0] = 0;
    J.job[J.n] = a;
    J.first[J.n + 1] = J.first[J.n] + total;
    ++J.n;
    return 0;
  }
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  pack_transposed_kernel<<<blocks, 256, 0, stream>>>(a);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_pack_transposed(const float* src, long long ld_src, int cols, int rows, __nv_bfloat16* dst, int RB,
                           int rows_dst_pad, int k_pad, cudaStream_t stream) {
  if (rows > rows_dst_pad) return -1;
  PackSeg seg{0, 0, rows};
  return launch_pack_transposed_seg(src, ld_src, cols, dst, RB, rows_dst_pad, k_pad, 0, k_pad, 1, &seg, stream);
}

int launch_ln_act(const float* scratch, long long ld, const float* stats, int NB, int RB, int M,
                  int m_pad, int N, const float* gamma, const float* beta, float eps, int act,
                  __nv_bfloat16* out, int out_kpad, cudaStream_t stream) {
  LnActArgs a{scratch, ld, stats, NB, RB, M, m_pad, N, gamma, beta, eps, act, out, out_kpad};
  // one warp per row; small row counts still spread over every SM (one warp per block)
  const int block = m_pad >= 148 * 8 ? 256 : 32;
  const cudaError_t le = launch_pdl(ln_act_kernel, grid_for(static_cast<long long>(m_pad) * 32, block, 148 * 8), block, 0,
                                    stream, a);
  if (le != cudaSuccess) return static_cast<int>(le);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_gru_gate(const float* scratch, long long ld, const float* stats, int NB, int RB, int M,
                    int m_pad, int D, const float* gamma, const float* beta, float eps,
                    float update_bias, const float* h_prev, long long ld_h, float* h_next,
                    long long ld_hn, __nv_bfloat16* h_next_packed, int kpad, cudaStream_t stream) {
  GruArgs a{scratch, ld, stats, NB, RB, M, m_pad, D, gamma, beta, eps, update_bias,
            h_prev, ld_h, h_next, ld_hn, h_next_packed, kpad};
  const bool vec = ((ld & 3) == 0) && ((D & 3) == 0) && ((ld_h & 3) == 0) && ((ld_hn & 3) == 0) &&
                   (((reinterpret_cast<uintptr_t>(scratch) | reinterpret_cast<uintptr_t>(h_prev) |
                      reinterpret_cast<uintptr_t>(h_next) | reinterpret_cast<uintptr_t>(gamma) |
                      reinterpret_cast<uintptr_t>(beta)) & 15) == 0);
  if (vec) {
    const int block = m_pad >= 148 * 8 ? 256 : 32;
    const cudaError_t le = launch_pdl(gru_gate_kernel, grid_for(static_cast<long long>(m_pad) * 32, block, 148 * 8), block,
                                      0, stream, a);
    if (le != cudaSuccess) return static_cast<int>(le);
  } else {
    const long long total = static_cast<long long>(m_pad) * ((kpad / 8 + 31) / 32) * 32;
    gru_gate_kernel_generic<<<grid_for(total, 256, 148 * 32), 256, 0, stream>>>(a);
  }
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_sample_latent(const float* logits, long long ld, int M, int groups, int classes,
                         NoiseSpec noise, uint8_t* idx_out, __nv_bfloat16* onehot_packed, int kpad,
                         float* onehot_f32, long long ld_f32, cudaStream_t stream) {
  if (classes != 32) return -1;
  SampleLatentArgs a{logits, ld, M, groups, noise, idx_out, onehot_packed, kpad, onehot_f32, ld_f32};
  const long long total = static_cast<long long>(M) * groups;
  if (total <= 0) return 0;
  // eight lanes per (row, group): four groups per warp iteration
  const long long warps = (total + 3) / 4;
  const int block = warps >= 148 * 8 ? 256 : 32;
  const cudaError_t le = launch_pdl(sample_latent_kernel, grid_for(warps * 32, block, 148 * 64), block, 0, stream, a);
  if (le != cudaSuccess) return static_cast<int>(le);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_sample_categorical(const float* logits, const float* uniforms, long long rows, int classes,
                              int32_t* idx_out, cudaStream_t stream) {
  if (rows <= 0) return 0;
  const int block = 128;
  sample_categorical_kernel<<<static_cast<unsigned>((rows + block - 1) / block), block, 0, stream>>>(
      logits, uniforms, rows, classes, idx_out);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_head_finish(const HeadFinishParams& p, cudaStream_t stream) {
  if (p.a_kpad > 64 || (p.a_kpad % 8) != 0) return -1;
  const int block = 128;
  const cudaError_t le = launch_pdl(head_finish_kernel, (p.m_pad + block - 1) / block, block, 0, stream, p);
  if (le != cudaSuccess) return static_cast<int>(le);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_lambda_return(const float* r, const float* v, const float* d, int T, long long N,
                         double lambda_, float* vs, float* w, float* adv, int layout_batch_major,
                         cudaStream_t stream) {
  if (T < 2 || N <= 0) return -1;
  // (1 - lambda) is formed in double like the Python expression in ac.py:58, then rounded to fp32
  const float c1 = static_cast<float>(1.0 - lambda_);
  const float c2 = static_cast<float>(lambda_);
  const int block = 256;
  if (layout_batch_major) {
    if (T > 32) return -2;
    int Tp = 1;
    while (Tp < T) Tp <<= 1;
    const long long warps = (N + (32 / Tp) - 1) / (32 / Tp);
    const long long threads = warps * 32;
    lambda_return_bm_kernel<<<static_cast<unsigned>((threads + block - 1) / block), block, 0, stream>>>(
        r, v, d, T, Tp, N, c1, c2, vs, w, adv);
  } else {
    const bool vec = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(r) | reinterpret_cast<uintptr_t>(v) |
                                       reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(vs) |
                                       reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(adv)) % 16 == 0);
    if (vec) {
      const long long threads = N / 4;
      lambda_return_tm_kernel<4><<<static_cast<unsigned>((threads + block - 1) / block), block, 0, stream>>>(
          r, v, d, T, N, c1, c2, vs, w, adv);
    } else {
      lambda_return_tm_kernel<1><<<static_cast<unsigned>((N + block - 1) / block), block, 0, stream>>>(
          r, v, d, T, N, c1, c2, vs, w, adv);
    }
  }
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_lambda_return_bwd(const float* g_vs, const float* v, const float* d, const float* vs,
                             int T, long long N, double lambda_, float* g_r, float* g_v, float* g_d,
                             cudaStream_t stream) {
  if (T < 2 || N <= 0) return -1;
  const float c1 = static_cast<float>(1.0 - lambda_);
  const float c2 = static_cast<float>(lambda_);
  const int block = 256;
  lambda_return_bwd_kernel<<<static_cast<unsigned>((N + block - 1) / block), block, 0, stream>>>(
      g_vs, v, d, vs, T, N, c1, c2, g_r, g_v, g_d);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

}  // namespace rlsb
