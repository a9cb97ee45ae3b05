// rlsb_wgrad.cuh — weight-gradient contraction  dW[n][k] = sum_m dY[m][n] * X[m][k]  on tcgen05.
//
// Both operands are read straight from the packed SWIZZLE_128B activation images the forward /
// backward GEMMs wrote (row = m, 128-byte rows of 64 columns): seen from this contraction they are
// MN-major UMMA operands (the contraction index m is the row index), so no transposed copies exist.
// Replaces (reference): the autograd of nn.Linear inside loss.backward() for the actor / critic
// MLPs (utils/optimizer.py:55-57 -> agents/dreamer/ac.py:68-81,113-146).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "rlsb_kernels.cuh"

namespace rlsb {

constexpr int kWgMaxSeg = 4;
constexpr int kWgMaxGroups = 4;

struct WgradParams {
  // dY: packed bf16 [G][M_pad x n_tiles*64], row block 128
  const __nv_bfloat16* dY;
  long long dy_group_stride;  // elements
  int n_tiles;
  // X: K segments, each a packed image [M_pad x x_ktiles*64]; x_mtile_stride = elements between
  // consecutive M tiles (x_ktiles*8192 normally, 0 for a tile shared by all M tiles, e.g. the
  // "ones" tile whose column 0 yields the bias gradient)
  const __nv_bfloat16* X[kWgMaxSeg];
  int x_ktiles[kWgMaxSeg];
  long long x_group_stride[kWgMaxSeg];
  long long x_mtile_stride[kWgMaxSeg];
  int n_seg;
  int G;
  int m_tiles;
  float* partial;  // [splits][G][n_slices*128][kt_total*64] fp32
  // ---- derived by plan_wgrad ----
  int kt_total, kc_tiles, n_chunks, n_slices, splits;
};

// fills the derived fields (work decomposition for the current device); returns 0 or < 0
int plan_wgrad(WgradParams& p);
size_t wgrad_partial_bytes(const WgradParams& p);
int launch_wgrad(const WgradParams& p, cudaStream_t stream);

struct WgradReduceParams {
  const float* partial;
  int splits, G, rows_pad, ld;   // geometry of `partial`
  float* w_dst[kWgMaxGroups];    // nn.Linear weight gradient [n_out][ld_dst] per group (may be nullptr)
  float* b_dst[kWgMaxGroups];    // bias gradient [n_out] per group (may be nullptr)
  int n_out[kWgMaxGroups];
  int ld_dst;
  int n_seg;
  PackSeg seg[3];                // dst_k0 = first padded column, src_c0 = first weight column, len
  int ones_col;                  // padded column that holds the bias gradient, or -1
  int accumulate;                // 1: dst += sum (contributions of several launches, e.g. slot-attention iterations)
};
int launch_wgrad_reduce(const WgradReduceParams& p, cudaStream_t stream);

// column sums written by the EPI_BWD epilogue: col_part[cta][G][2][RB] -> d_gamma, d_beta (per group pointers)
struct ColsumReduceParams {
  const float* col_part;
  int ctas, G, RB, N;
  float* dgamma[kWgMaxGroups];
  float* dbeta[kWgMaxGroups];
};
int launch_colsum_reduce(const ColsumReduceParams& p, cudaStream_t stream);

}  // namespace rlsb
