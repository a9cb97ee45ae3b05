// rlsb_gemm.cuh — parameters of the tcgen05 row-block GEMM used by every contraction on the
// imagination path (reference call sites: rssm.py:179-192, common.py:69-81, fc_nn.py:4-23).
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace rlsb {

constexpr int kTileM = 128;      // rows per M tile == TMEM lanes
constexpr int kTileK = 64;       // bf16 elements per 128-byte swizzled row
constexpr int kMaxSeg = 8;       // K segments (concatenated inputs, e.g. cat[x, h]; slotted heads: 2 per slot)
constexpr int kGemmThreads = 576;  // warp0 = bulk-copy producer, warp1 = MMA issuer, warps2-17 = epilogue

enum GemmEpilogue : int {
  EPI_PLAIN = 0,   // out_f32[m][col] = acc + bias                         (row-major fp32)
  EPI_STATS = 1,   // EPI_PLAIN + per-(row, n-block) (mean, M2) partials    (for a later LayerNorm)
  EPI_LN_ACT = 2,  // full row in TMEM: [LayerNorm] -> activation -> packed bf16 (+ optional saves for backward)
  EPI_LN_ACT_SAVE = 4,  // EPI_LN_ACT that also stores save_pre / save_rstd (selected by the launcher when save_pre != nullptr)
  EPI_BWD = 3,     // full row in TMEM: acc = dL/d(act output); ELU' and LayerNorm backward -> packed bf16 dL/d(pre-LN)
  EPI_GRU = 5,     // GRU cell (common.py:69-81) fused into its contraction: n-block nb holds the [reset | candidate | update]
                   // pre-activations of hidden units [64 nb, 64 nb + 64) (RB = 192, weight rows permuted by the packer); the
                   // joint LayerNorm over all 3D columns goes through the cross-block exchange (xstats); writes only h'
};

enum Activation : int { ACT_NONE = 0, ACT_ELU = 1, ACT_RELU = 2 };

struct GemmParams {
  // ---- A: packed bf16 [M_pad x K_s] (row block 128), one pointer per K segment -------------
  const __nv_bfloat16* A[kMaxSeg];
  int a_ktiles[kMaxSeg];            // K_s / 64
  long long a_group_stride[kMaxSeg];  // elements between groups (0 = shared by all groups)
  int n_seg;
  // ---- B: packed bf16 weights [(G*NB*RB) x K_total] with row block RB ------------------------
  const __nv_bfloat16* W;
  int RB;  // rows (output columns) per n-block: multiple of 32, <= 512
  int NB;  // n-blocks per group
  int G;   // groups (independent heads sharing the launch)
  // ---- problem ------------------------------------------------------------------------------
  int M;        // valid rows
  int m_tiles;  // ceil(M / 128)
  int N;        // valid output columns per group (<= NB*RB)
  const float* bias;  // [G][NB*RB] (zero padded) or nullptr
  // ---- EPI_PLAIN / EPI_STATS ----------------------------------------------------------------
  float* out_f32;              // [G][M_pad][ldo]
  long long ldo;
  long long out_group_stride;  // elements
  float* stats;                // [G][NB][M_pad][2] = (sum, sum of squares) over the block's valid columns
  // ---- EPI_LN_ACT ---------------------------------------------------------------------------
  const float* ln_gamma;  // [G][RB] or nullptr (=> no LayerNorm)
  const float* ln_beta;
  float ln_eps;
  int act;
  __nv_bfloat16* out_bf16;  // packed [G][M_pad x out_kpad]
  int out_kpad;             // multiple of 64, >= N
  long long out_bf16_group_stride;
  // ---- stacked row blocks (H per-step images of m_pad rows each): row m is valid iff
  //      (m % row_period) < row_valid; row_period == 0 => valid iff m < M.  Invalid rows are
  //      written as zeros by EPI_LN_ACT / EPI_BWD so they never contribute to a weight gradient.
  int row_period, row_valid;
  // ---- EPI_LN_ACT, training forward: keep what the backward pass needs ---------------------
  __nv_bfloat16* save_pre;  // packed like out_bf16: normalised x_hat (LayerNorm) or the pre-activation (no LN)
  float* save_rstd;         // [G][M_pad] 1/sqrt(var + eps) per row (LayerNorm only) or nullptr
  // ---- EPI_BWD: acc[m][n] = dL/dy, y = act(gamma * pre + beta) (LN) or act(pre) ---------------
  //   dL/da = acc * act'(a);  LN: dxh = dL/da * gamma, out = rstd * (dxh - mean(dxh) - pre * mean(dxh * pre))
  //   column sums over valid rows: d_gamma = sum dL/da * pre, d_beta = sum dL/da -> col_part[cta][g][2][RB]
  const __nv_bfloat16* bwd_pre;  // packed [G][M_pad x out_kpad] (same geometry as out_bf16)
  const float* bwd_rstd;         // [G][M_pad] or nullptr
  float* col_part;               // [gridDim.x][G][2][RB] (zeroed by the launcher) or nullptr
  int group_major;               // work order: all M tiles of group 0, then group 1, ... (EPI_BWD)
  // ---- one group of a grouped launch living in other buffers (K1's actor head, whose activations rlsb_ac_update reuses
  //      instead of recomputing them): group alt_group_p1 - 1 (0 = none) reads its segment-0 A operand from alt_A (if
  //      non-null), writes its packed output to alt_out_bf16, and is the only group that stores save_pre / save_rstd
  //      (both then point at that group's image: no group stride is applied)
  int alt_group_p1;
  const __nv_bfloat16* alt_A;
  __nv_bfloat16* alt_out_bf16;
  // ---- LayerNorm over a row that spans NB > 1 n-blocks (EPI_LN_ACT with NB > 1, EPI_GRU): every CTA publishes its block's
  //      (sum, sum of squares) per row in xstats[nb][m] = two 64-bit words {value, tag} and gathers the NB blocks' words of its
  //      rows, polling until they carry the launch's tag (= the slot's previous tag + 1) — the CTAs of a launch are co-resident
  //      (grid <= #SMs, one CTA per SM) and walk the work items n-block-fastest, so the partners of a tile are running or
  //      finished.  xstats: [NB][m_tiles * 128][2] uint64, owned by ONE layer (every launch must write every slot of a row once),
  //      zeroed once.
  unsigned long long* xstats;
  // ---- EPI_GRU: h' = u * cand + (1 - u) * h ------------------------------------------------
  const float* gru_h_prev;   // fp32 [M][gru_ld_h]
  long long gru_ld_h;
  float* gru_h_next;         // fp32 [M][gru_ld_hn]; the packed bf16 image goes to out_bf16 (out_kpad = NB * 64)
  long long gru_ld_hn;
  float gru_update_bias;
  // profiling (set_gemm_trace): [gridDim.x][64 tiles][8] %globaltimer stamps of every CTA's first 64 EPI_GRU tiles — 0 tile's
  // parameters staged, 1 accumulator ready, 2 copied to registers (TMEM buffer released), 3 statistics published,
  // 4 all partners' statistics gathered, 5 totals known, 6 gates done, 7 bulk store issued; nullptr = off
  unsigned long long* trace;
  // ---- set by launch_gemm (callers leave it 0): the full-row epilogue assembles each 128 x 64 output tile (16 KB, contiguous
  //      in the packed image) in shared memory and writes it with one bulk copy instead of 16-byte stores scattered over 32 rows
  int staged_out;
  // ---- set by launch_gemm: 1 / N as the host's correctly rounded fp32 quotient (== the device's IEEE division, without the
  //      FCHK + MUFU.RCP + Newton sequence every epilogue thread would run per tile); used when a block holds all N columns
  float inv_n;
  // ---- set by launch_gemm (EPI_BWD): bytes of the shared-memory buffer the epilogue threads prefetch their saved x_hat /
  //      pre-activation chunks into with cp.async (0 = the epilogue reads them from global memory, twice)
  int xbuf_bytes;
};

// Launch on `stream`; returns cudaError_t as int (0 = ok) or a negative argument error.
int launch_gemm(const GemmParams& p, int epilogue, cudaStream_t stream);
// number of CTAs launch_gemm will use for `p` (size of the col_part buffer of EPI_BWD)
int gemm_grid_size(const GemmParams& p);

// CTAs per thread-block cluster that share one weight block through TMA multicast (1, 2 or 4)
void set_gemm_cluster_size(int cs);
int gemm_cluster_size();
// full-row epilogues write their output tiles through shared memory + bulk copies (1, default) or with 16-byte stores (0);
// any other value only queries.  Returns the value in effect.
int set_gemm_staged_output(int on);
// profiling: device buffer for GemmParams::trace of the following EPI_GRU launches (nullptr = off)
void set_gemm_trace(unsigned long long* device_buffer);

}  // namespace rlsb
