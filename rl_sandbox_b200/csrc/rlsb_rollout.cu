// rlsb_rollout.cu — K1 as ONE persistent kernel: the whole H-step imagination rollout
// (DreamerV2.imagine_trajectory, agents/dreamer_v2.py:68-96) of a 128-row block of start states runs inside one
// thread-block cluster, from the first actor evaluation to the last latent draw, without leaving the GPU.
//
// Why a cluster per row block.  Start states are independent, so a row block's rollout depends on nothing
// outside itself: the only synchronisation the 15 x ~11 dependent layers need is between the CTAs that share the row
// block — a hardware cluster barrier (~0.2 us) instead of a kernel boundary (launch + prologue + pipeline refill +
// teardown: 10-20 us per layer at the configured 800 start states, where the chained rollout of rlsb_imagine.cu is a
// string of ~230 such launches).  Every layer's output columns are split over the C CTAs of the cluster:
//   * each CTA streams its own slab of the layer's weights ([k-tile][NC x 64] bf16, re-ordered per CTA by
//     rlsb_rollout_pack) with bulk copies into a shared-memory ring and accumulates a 128 x NC tile in TMEM with
//     tcgen05.mma (M = 128, kind::f16);
//   * LayerNorm rows span several CTAs: every CTA reduces its columns to a per-row (sum, sum of squares), publishes
//     them in its own shared memory, and after one cluster barrier reads its peers' partials through distributed
//     shared memory (ld.shared::cluster); the accumulator waits in TMEM meanwhile;
//   * the GRU (common.py:69-81) is fused into its contraction's epilogue: a CTA owns the (reset, candidate, update)
//     pre-activations of the same D / C hidden units, the joint LayerNorm over 3D goes through the DSMEM exchange, and
//     the epilogue applies gates and the convex update and writes only h' (fp32 + the packed bf16 image);
//   * the reward / value / discount read-out with the action draw, and the 32 x 32 straight-through latent draw, run as
//     row phases on the epilogue warps of all CTAs (rows split over the cluster).
// Activations travel between the CTAs of a cluster through the packed operand images in global memory (L2 resident:
// a row block's images are a few hundred KB), written by the epilogues and read back by the next layer's bulk copies.
//
// Same inputs, outputs, workspace, tape and noise contract as rlsb_imagine_fwd (flat RSSM); the packed weights are its
// own (rlsb_rollout_pack).  Meant for the launch-bound regime (a few thousand start states); the chained rollout with
// CTA-pair MMAs over L2-shared weights stays the path for the 16 k - 256 k sweep.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/rlsb.h"
#include "rlsb_count.cuh"
#include "rlsb_detmath.h"
#include "rlsb_gemm.cuh"
#include "rlsb_imagine_plan.cuh"
#include "rlsb_kernels.cuh"
#include "rlsb_ptx.cuh"
#include "rlsb_rowops.cuh"
#include "rlsb_bwd_rowops.cuh"

namespace rlsb {
namespace ro {

constexpr int kMaxC = 16;
constexpr int kGruChunksHost = 5;   // == kGruChunks of the kernel
constexpr int kThreads = kGemmThreads;   // warp 0 producer, warp 1 MMA issuer, warps 2-17 epilogue / row phases
constexpr int kEpiThreads = 512;

// ---- one layer as the kernel sees it -------------------------------------------------------------------------------
struct RLayer {
  const __nv_bfloat16* W;   // [ranks][kt][NC x 64] packed bf16 slabs (SWIZZLE_128B rows, one slab per CTA rank)
  const float* bias;        // [ranks][NC]
  const float* gamma;       // [ranks][NC] (LayerNorm layers)
  const float* beta;
  int NC;                   // slab rows = accumulator columns per CTA (multiple of 16, <= 512)
  int kt;                   // k tiles
  int cpg;                  // CTAs per LayerNorm group (heads: per head; RSSM layers: all active ranks)
  int ranks;                // active ranks = groups x cpg
  int n;                    // valid columns per LayerNorm group
  int wpad;                 // GRU: padded slice width (NC = 3 wpad)
  short col0[kMaxC];        // per rank inside its group: first logical column ...
  short width[kMaxC];       // ... and valid columns
};

struct RolloutParams {
  RLayer head[5], img_in, gru, prior1, prior2;
  int C, M, m_pad, H;
  int D, S, A, Aout, Dp, Sp, Ap, Hp, G, groups;
  int g_actor, g_reward, g_discount, g_critic;
  int discrete, layer_norm, nan_on_tie, last_step_value_only;
  float eps;
  // packed state images: img(t) = base + (pingpong ? (t & 1) : t) * step
  __nv_bfloat16* himg; long long himg_step; int himg_pingpong;
  __nv_bfloat16* zimg; long long zimg_step; int zimg_pingpong;
  __nv_bfloat16 *abf, *xbf, *ybf, *hid[2];
  float* head_out;          // [G][m_pad][32]
  // outputs (rlsb_imagine_out)
  float *determ, *logits, *stoch, *actions, *rewards, *discounts, *values, *actor_raw;
  uint8_t* stoch_idx;
  // noise (rlsb_noise)
  const float *latent_uniforms, *action_noise, *precomp;
  uint64_t seed; const uint64_t* seed_ptr; uint32_t row_offset;
  // activation tape (rlsb_imagine_bwd) or nullptr
  uint8_t* tape; size_t tape_step;
  size_t tp_head_pre[4], tp_head_rstd[4], tp_x_pre, tp_x_rstd, tp_gru_scratch, tp_gru_stats, tp_y_pre, tp_y_rstd;
  long long tape_ld_scratch; int tape_gru_nb;
  // profiling aid (rlsb_rollout_set_trace): [H+1][11 phases][8] globaltimer stamps of cluster 0, see `trp` in the kernel;
  // nullptr = off
  unsigned long long* trace;
  int kg_max;   // k tiles per pipeline stage, at most (RLSB_ROLLOUT_KG, default 4)
  int dbg;   // timing experiments only (RLSB_ROLLOUT_DEBUG; results are garbage): 1 = no MMAs, 2 = no weight copies, 4 = no A copies
};

// split n columns into `parts` slices whose boundaries are multiples of 8 (16-byte chunks of the packed images)
inline void split8(int n, int parts, short* col0, short* width) {
  const int units = (n + 7) / 8;
  const int base = units / parts, rem = units % parts;
  int c = 0;
  for (int i = 0; i < parts; ++i) {
    const int u = base + (i < rem ? 1 : 0);
    col0[i] = static_cast<short>(c * 8);
    int end = (c + u) * 8;
    if (end > n) end = n;
    width[i] = static_cast<short>(u > 0 ? end - c * 8 : 0);
    c += u;
  }
}

// ---- host plan: geometry of every layer and its place in the packed blob ------------------------------------------
struct HLayer {
  int NC = 0, kt = 0, cpg = 1, ranks = 0, n = 0, wpad = 0, groups = 1;
  short col0[kMaxC] = {}, width[kMaxC] = {};
  size_t w_off = 0, bias_off = 0, g_off = 0, b_off = 0;
};
struct RPlan {
  int C = 8;
  k1::Plan P;
  HLayer head[5], img_in, gru, prior1, prior2;
  // transposed layers of the backward rollout (cfg.with_backward): slab rows = in-features
  bool bwd = false;
  HLayer t_head[5], t_prior2, t_prior1, t_gru, t_img_in;
  size_t gru_ln_off = 0;   // fp32 [2][3D]: determ_recurrent._norm gamma | beta in the reference's order (gate backward row code)
  size_t bytes = 0;
};

inline int make_rplan(const rlsb_imagine_cfg& cfg, int C, RPlan& R) {
  if (C != 4 && C != 8 && C != 16) return -30;
  rlsb_imagine_cfg c = cfg;
  c.parity = 0;
  c.slots = 0;
  if (k1::make_plan(c, R.P) != 0) return -31;
  if (cfg.slots > 1 || cfg.parity) return -32;   // flat RSSM, bf16 contractions
  const k1::Plan& P = R.P;
  if (P.G > C) return -33;
  R.C = C;
  size_t cur = 0;
  auto finish = [&](HLayer& L) {
    int maxw = 0;
    for (int i = 0; i < L.cpg; ++i) maxw = L.width[i] > maxw ? L.width[i] : maxw;
    if (L.wpad > 0) L.NC = 3 * L.wpad;
    else L.NC = k1::ru(maxw, 16);
    if (L.NC < 16) L.NC = 16;
    if (L.NC > 512) return -34;
    L.ranks = L.groups * L.cpg;
    L.w_off = k1::place(cur, static_cast<size_t>(L.ranks) * L.kt * L.NC * 128);
    L.bias_off = k1::place(cur, static_cast<size_t>(L.ranks) * L.NC * 4);
    L.g_off = k1::place(cur, static_cast<size_t>(L.ranks) * L.NC * 4);
    L.b_off = k1::place(cur, static_cast<size_t>(L.ranks) * L.NC * 4);
    return 0;
  };
  // RSSM layers: every CTA of the cluster takes a slice (fewer when the layer has fewer 8-column chunks than CTAs)
  auto rssm = [&](HLayer& L, int n, int kt) {
    L.n = n; L.kt = kt; L.groups = 1;
    int parts = C;
    while (parts > 1 && (n + 7) / 8 < parts) parts >>= 1;
    L.cpg = parts;
    split8(n, parts, L.col0, L.width);
    return finish(L);
  };
  int e;
  if ((e = rssm(R.img_in, P.D, (P.Sp + P.Ap) / 64)) != 0) return e;
  {
    HLayer& L = R.gru;   // slices of hidden units; a CTA holds the three gates of its units: [r | c | u], each wpad wide
    L.n = 3 * P.D; L.kt = 2 * P.Dp / 64; L.groups = 1;
    int parts = C;
    while (parts > 1 && (P.D + 7) / 8 < parts) parts >>= 1;
    L.cpg = parts;
    split8(P.D, parts, L.col0, L.width);
    int maxw = 0;
    for (int i = 0; i < parts; ++i) maxw = L.width[i] > maxw ? L.width[i] : maxw;
    L.wpad = k1::ru(maxw, 16);
    if (L.wpad > 8 * 4 * kGruChunksHost) return -36;
    if ((e = finish(L)) != 0) return e;
  }
  if ((e = rssm(R.prior1, P.D, P.Dp / 64)) != 0) return e;
  if ((e = rssm(R.prior2, P.S, P.Dp / 64)) != 0) return e;
  for (int l = 0; l < 5; ++l) {
    HLayer& L = R.head[l];
    L.groups = P.G;
    L.kt = (l == 0) ? (P.Dp + P.Sp) / 64 : P.Hp / 64;
    if (l < 4) {
      L.n = P.Hd;
      L.cpg = C / P.G;
      while (L.cpg > 1 && (P.Hd + 7) / 8 < L.cpg) --L.cpg;
      split8(P.Hd, L.cpg, L.col0, L.width);
    } else {
      L.n = P.Aout;   // per group: Aout (actor) or 1 — the slab is Aout wide for every group, unused rows are zero
      L.cpg = 1;
      L.col0[0] = 0;
      L.width[0] = static_cast<short>(P.Aout);
    }
    if ((e = finish(L)) != 0) return e;
  }
  // ---- backward rollout (rlsb_rollout_bwd): dX contractions against the transposed weights ------------------
  R.bwd = P.bwd;
  if (R.bwd) {
    // (a cluster size whose slices do not fit the LayerNorm-backward epilogue's register plan — at most 4 chunks = 128
    // accumulator columns per CTA — leaves the forward kernel available: R.bwd = false, rlsb_rollout_bwd refuses)
    const size_t cur0 = cur;
    auto tlayer = [&](HLayer& L, int groups, int n_rows, int kt, int cpg) {
      L.groups = groups; L.n = n_rows; L.kt = kt; L.cpg = cpg;
      while (L.cpg > 1 && (n_rows + 7) / 8 < L.cpg) --L.cpg;
      split8(n_rows, L.cpg, L.col0, L.width);
      return finish(L);
    };
    const int cpg_h = C / P.Gb > 0 ? C / P.Gb : 1;
    int eb = 0;
    for (int l = 1; l < 5 && eb == 0; ++l) eb = tlayer(R.t_head[l], P.Gb, P.Hd, k1::ru(P.head[l].N, 64) / 64, cpg_h);
    if (eb == 0) eb = tlayer(R.t_head[0], 1, P.Dp + P.Sp, P.Gb * P.Hp / 64, C);
    if (eb == 0) eb = tlayer(R.t_prior2, 1, P.D, P.Sp / 64, C);
    if (eb == 0) eb = tlayer(R.t_prior1, 1, P.D, P.Dp / 64, C);
    if (eb == 0) eb = tlayer(R.t_gru, 2, P.D, P.G3p / 64, C / 2);
    if (eb == 0) eb = tlayer(R.t_img_in, 1, P.Sp + P.Ap, P.Dp / 64, C);
    for (int l = 1; l < 5 && eb == 0; ++l)
      if (R.t_head[l].NC > 128) eb = -37;
    if (eb == 0 && (R.t_prior2.NC > 128 || R.t_gru.NC > 128 || P.Gb > 4)) eb = -37;
    // gate-backward row phase: 16 / (128 / C) warps per row, at most two 4-unit slices per lane
    if (eb == 0 && (C < 8 || P.D / 4 > 2 * 32 * (C / 8))) eb = -39;
    if (eb == 0) {
      R.gru_ln_off = k1::place(cur, static_cast<size_t>(2) * 3 * P.D * 4);
    } else {
      R.bwd = false;
      cur = cur0;
    }
  }
  R.bytes = k1::rus(cur, 1024);
  return 0;
}

// ---- weight re-pack: nn.Linear fp32 (out, in) -> per-CTA slabs --------------------------------------------------------
struct RowSeg {
  int dst_r0, src_i0, len;   // rows [dst_r0, dst_r0 + len) of the (padded) slab row axis <- in-features src_i0 ...
};
struct RPackJob {
  const float* w; long long ld;
  const float* b; const float* g; const float* be;   // bias / LayerNorm gamma / beta by slab row (nullable)
  __nv_bfloat16* W; float* bias; float* gamma; float* beta;   // destinations of this job's first rank
  // mode 0: slab rows = out-features (a contiguous slice per rank), K = in-features through `seg`
  // mode 1: GRU triplets [r | c | u] of a slice of hidden units
  // mode 2: TRANSPOSED (the dX contractions of the backward rollout): slab rows = in-features through `rseg` (the padded row
  //         axis is sliced per rank), K = out-features through `seg`; only the K range [k_lo, k_hi) is written (several
  //         Linears can share one K axis: K-concatenated head groups)
  int ranks, NC, kt, n_out, mode, D, wpad, n_seg;
  PackSeg seg[4];
  int n_rseg, k_lo, k_hi;
  RowSeg rseg[2];
  short col0[kMaxC], width[kMaxC];
};
constexpr int kMaxJobs = 48;
struct RPackJobs {
  int n;
  int first_block[kMaxJobs + 1];
  RPackJob job[kMaxJobs];
};

namespace {

__device__ __forceinline__ int rpack_src_row(const RPackJob& j, int rank, int r) {
  if (j.mode == 0) return (r < j.width[rank] && j.col0[rank] + r < j.n_out) ? j.col0[rank] + r : -1;
  if (j.mode == 2) {
    if (r >= j.width[rank]) return -1;
    const int rg = j.col0[rank] + r;
    for (int i = 0; i < j.n_rseg; ++i)
      if (rg >= j.rseg[i].dst_r0 && rg < j.rseg[i].dst_r0 + j.rseg[i].len) return j.rseg[i].src_i0 + rg - j.rseg[i].dst_r0;
    return -1;
  }
  const int gate = r / j.wpad, u = r - gate * j.wpad;
  return (gate < 3 && u < j.width[rank]) ? gate * j.D + j.col0[rank] + u : -1;
}

// one thread per 16-byte chunk of a slab; the first threads of a job's first block also re-order bias / gamma / beta
__global__ void __launch_bounds__(256) rpack_kernel(const __grid_constant__ RPackJobs J) {
  int ji = 0;
  while (ji + 1 < J.n && static_cast<int>(blockIdx.x) >= J.first_block[ji + 1]) ++ji;
  const RPackJob& j = J.job[ji];
  const long long local = (static_cast<long long>(blockIdx.x) - J.first_block[ji]) * blockDim.x + threadIdx.x;
  const long long chunks = static_cast<long long>(j.ranks) * j.kt * j.NC * 8;
  if (local < static_cast<long long>(j.ranks) * j.NC) {
    const int rank = static_cast<int>(local / j.NC), r = static_cast<int>(local % j.NC);
    const int src = rpack_src_row(j, rank, r);
    j.bias[local] = (src >= 0 && j.b) ? j.b[src] : 0.f;
    if (j.gamma) j.gamma[local] = (src >= 0 && j.g) ? j.g[src] : 1.f;
    if (j.beta) j.beta[local] = (src >= 0 && j.be) ? j.be[src] : 0.f;
  }
  if (local >= chunks) return;
  const int ch = static_cast<int>(local & 7);
  const long long rr = local >> 3;
  const int r = static_cast<int>(rr % j.NC);
  const long long t2 = rr / j.NC;
  const int kt = static_cast<int>(t2 % j.kt);
  const int rank = static_cast<int>(t2 / j.kt);
  const int src = rpack_src_row(j, rank, r);
  const int k0 = kt * 64 + ch * 8;
  if (j.mode == 2 && (k0 < j.k_lo || k0 >= j.k_hi)) return;   // another job's share of the K axis
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = 0.f;
  if (src >= 0) {
    const float* row = j.w + static_cast<long long>(src) * j.ld;
    for (int s = 0; s < j.n_seg; ++s) {
      const PackSeg& sg = j.seg[s];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = k0 + e - sg.dst_k0;
        if (k >= 0 && k < sg.len)
          v[e] = j.mode == 2 ? j.w[static_cast<long long>(sg.src_c0 + k) * j.ld + src] : row[sg.src_c0 + k];
      }
    }
  }
  __nv_bfloat16* dst = j.W + ((static_cast<size_t>(rank) * j.kt + kt) * j.NC + r) * 64 + ((ch ^ (r & 7)) << 3);
  *reinterpret_cast<uint4*>(dst) = make_uint4(rowops::bf2(v[0], v[1]), rowops::bf2(v[2], v[3]), rowops::bf2(v[4], v[5]),
                                              rowops::bf2(v[6], v[7]));
}

// ---- device helpers ---------------------------------------------------------------------------------------------
struct RCtl {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full;
  uint32_t tmem_base;
  uint32_t pad;
  alignas(16) float bias[512];
  float gamma[512];
  float beta[512];
  float2 part[4][kTileM];    // per column-quarter partial (sum, sumsq) of each row
  // per-row (sum, sumsq) of the current LayerNorm layer from every CTA of this CTA's LayerNorm group: each CTA PUSHES its
  // partial into slot [its index in the group] of all its peers (st.shared::cluster) before the cluster barrier and reads
  // only its own shared memory after it (two sets: a fast peer may already push the next layer's)
  float2 xstat[2][kMaxC][kTileM];
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// profiling stamps: `p` = this phase's 8 slots (nullptr unless this thread is one of cluster 0's tracer threads)
__device__ __forceinline__ void tr(unsigned long long* p, int slot) {
  if (p) p[slot] = globaltimer_ns();
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void st_dsmem_f2(uint32_t cluster_addr, float2 v) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void epi_bar(int id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kEpiThreads) : "memory");
}
__device__ __forceinline__ float elu1(float x) {
  return fmaxf(x, rowops::ex2_approx(1.4426950408889634f * fminf(x, 0.f)) - 1.0f);
}
// position of (row r of the 128-row block, column col) inside a packed image's row block (elements)
__device__ __forceinline__ size_t pk_off(int r, int col) {
  return static_cast<size_t>(col >> 6) * (kTileM * kTileK) + static_cast<size_t>(r) * kTileK +
         static_cast<size_t>((((col & 63) >> 3) ^ (r & 7)) << 3) + (col & 7);
}

struct OpA {   // A operand of a layer for this row block: up to four K segments (concatenated inputs)
  const __nv_bfloat16* A[4];
  int kt[4];
  int nseg;
};

struct Roles {
  uint32_t pe;   // producer: parity to wait for on empty[s] (bit s)
  uint32_t pf;   // MMA issuer: parity to wait for on full[s]
  uint32_t tf;   // epilogue: parity to wait for on tmem_full
};

// A pipeline stage holds a GROUP of up to `kg` consecutive k tiles of one K segment (their A tiles are contiguous in the
// packed image, and so are the slab's): one "full" barrier, one tcgen05.commit per group instead of per k tile — the
// commit is what paces a main loop of small MMAs (it is not overlapped with the MMAs that follow it).
struct Ring {
  int kg;            // k tiles per stage
  int stages;
  uint32_t a_bytes;  // bytes reserved for the A tiles of a stage (kg x 16 KB)
  uint32_t stage_bytes;
};
__device__ __forceinline__ Ring make_ring(int ring_bytes, int NC, int kg_max) {
  const int per_kt = 16384 + NC * 128;
  int kg = ring_bytes / (2 * per_kt);   // at least two stages
  if (kg > kg_max) kg = kg_max;
  if (kg < 1) kg = 1;
  Ring r;
  r.kg = kg;
  r.a_bytes = static_cast<uint32_t>(kg) * 16384u;
  r.stage_bytes = static_cast<uint32_t>(kg * per_kt);
  r.stages = ring_bytes / static_cast<int>(r.stage_bytes);
  if (r.stages > 8) r.stages = 8;
  return r;
}

// k tiles of the n-th group of a K segment of `left` remaining tiles
__device__ __forceinline__ int group_tiles(int n, int kg, int left) {
  (void)n;   // (single-tile first groups were tried: the first MMA starts 0.2 us earlier, the main loops run 0.8 us longer)
  return kg < left ? kg : left;
}

// producer thread: stream the A tiles of this row block and this CTA's weight slab through the stage ring
__device__ __forceinline__ void produce(RCtl* ctl, uint8_t* ring, int ring_bytes, const RLayer& L, int rank, const OpA& a,
                                        Roles& st, int dbg, int kg_max, unsigned long long* trp) {
  const Ring R = make_ring(ring_bytes, L.NC, kg_max);
  const uint32_t b_kt = static_cast<uint32_t>(L.NC) * 128u;
  tr(trp, 0);
  const __nv_bfloat16* w = L.W + static_cast<size_t>(rank) * L.kt * L.NC * 64;
  fence_proxy_async_all();   // the images were written with ordinary stores by other CTAs before the cluster barrier
  int s = 0, ktg = 0, ng = 0;
  for (int sg = 0; sg < a.nseg; ++sg) {
    for (int kt = 0, g = 0; kt < a.kt[sg]; kt += g, ++ng) {
      g = group_tiles(ng, R.kg, a.kt[sg] - kt);
      mbar_wait(&ctl->empty[s], (st.pe >> s) & 1u);
      st.pe ^= 1u << s;
      uint8_t* sa = ring + static_cast<size_t>(s) * R.stage_bytes;
      const uint32_t ab = static_cast<uint32_t>(g) * 16384u, bb = static_cast<uint32_t>(g) * b_kt;
      if (elect_one()) {
        mbar_expect_tx(&ctl->full[s], ((dbg & 4) ? 0u : ab) + ((dbg & 2) ? 0u : bb));
        if (!(dbg & 4)) bulk_g2s(sa, a.A[sg] + static_cast<size_t>(kt) * (kTileM * kTileK), ab, &ctl->full[s]);
        if (!(dbg & 2)) bulk_g2s(sa + R.a_bytes, w + static_cast<size_t>(ktg) * L.NC * 64, bb, &ctl->full[s]);
      }
      __syncwarp();
      ktg += g;
      if (++s == R.stages) s = 0;
    }
  }
}

// MMA thread: acc[128 x NC] (TMEM columns 0..NC) = sum over k tiles of A_tile * W_tile^T
__device__ __forceinline__ void issue_mma(RCtl* ctl, uint8_t* ring, int ring_bytes, const RLayer& L, uint32_t tmem_base,
                                          const OpA& a, Roles& st, int dbg, int kg_max, unsigned long long* trp) {
  const Ring R = make_ring(ring_bytes, L.NC, kg_max);
  const uint32_t b_kt = static_cast<uint32_t>(L.NC) * 128u;
  const int n0 = L.NC > 256 ? 256 : L.NC, n1 = L.NC - n0;
  const uint32_t idesc0 = make_idesc_bf16(kTileM, static_cast<uint32_t>(n0));
  const uint32_t idesc1 = n1 > 0 ? make_idesc_bf16(kTileM, static_cast<uint32_t>(n1)) : 0u;
  tc_fence_after();   // the previous layer's epilogue has drained TMEM (cluster barrier in between)
  int s = 0, ng = 0;
  uint32_t acc = 0u;
  for (int sg = 0; sg < a.nseg; ++sg) {
    for (int kt = 0, g = 0; kt < a.kt[sg]; kt += g, ++ng) {
      g = group_tiles(ng, R.kg, a.kt[sg] - kt);
      mbar_wait(&ctl->full[s], (st.pf >> s) & 1u);
      st.pf ^= 1u << s;
      tc_fence_after();
      if (acc == 0u) tr(trp, 1);
      const uint32_t sa = smem_u32(ring + static_cast<size_t>(s) * R.stage_bytes);
      if (elect_one()) {
        for (int j = 0; j < g; ++j) {
          const uint64_t adesc = make_smem_desc_sw128(sa + static_cast<uint32_t>(j) * 16384u);
          const uint32_t sb = sa + R.a_bytes + static_cast<uint32_t>(j) * b_kt;
          const uint64_t bdesc0 = make_smem_desc_sw128(sb);
          const uint64_t bdesc1 = make_smem_desc_sw128(sb + static_cast<uint32_t>(n0) * 128u);
#pragma unroll
          for (int kk = 0; kk < kTileK / 16; ++kk) {
            if (dbg & 1) break;
            umma_bf16(tmem_base, adesc + static_cast<uint64_t>(kk * 2), bdesc0 + static_cast<uint64_t>(kk * 2), idesc0, acc);
            if (n1 > 0)
              umma_bf16(tmem_base + 256u, adesc + static_cast<uint64_t>(kk * 2), bdesc1 + static_cast<uint64_t>(kk * 2),
                        idesc1, acc);
            acc = 1u;
          }
        }
        umma_commit(&ctl->empty[s]);
      }
      __syncwarp();
      acc = 1u;
      if (++s == R.stages) s = 0;
    }
  }
  tr(trp, 2);
  if (elect_one()) umma_commit(&ctl->tmem_full);
  __syncwarp();
}

// epilogue warps: bring this rank's bias / gamma / beta into shared memory, then wait for the accumulator
__device__ __forceinline__ void epi_begin(RCtl* ctl, const RLayer& L, int rank, int tid_e, bool ln, Roles& st) {
  const size_t off = static_cast<size_t>(rank) * L.NC;
  for (int i = tid_e; i < L.NC; i += kEpiThreads) {
    ctl->bias[i] = __ldg(L.bias + off + i);
    if (ln) {
      ctl->gamma[i] = __ldg(L.gamma + off + i);
      ctl->beta[i] = __ldg(L.beta + off + i);
    }
  }
  epi_bar(1);
  mbar_wait(&ctl->tmem_full, st.tf & 1u);
  st.tf ^= 1u;
  tc_fence_after();
}

// every thread of every CTA of the cluster; `writer`: this thread stored to global memory that other CTAs' bulk
// copies (async proxy) will read after the barrier
__device__ __forceinline__ void layer_end(bool writer) {
  if (writer) {
    fence_proxy_async_all();
    tc_fence_before();
  }
  __syncwarp();
  cluster_sync_all();
}

struct RowStat {
  float mean, rstd;
};
// LayerNorm statistics of a row whose columns are spread over the 4 column-quarter warps of this CTA and over the
// `cpg` CTAs [g0, g0 + cpg) of the cluster.  Called by all 512 epilogue threads; contains the cluster barrier that the
// other warps of the CTA (and the CTAs without work in this layer) match with a bare cluster_sync_all().
__device__ __forceinline__ RowStat ln_exchange(RCtl* ctl, int cq, int row, float sum, float sq, int g0, int cpg, int gi,
                                               int n, float eps, int par, float2* total = nullptr) {
  ctl->part[cq][row] = make_float2(sum, sq);
  epi_bar(2);
  if (cq == 0) {
    const float2 a0 = ctl->part[0][row], a1 = ctl->part[1][row], a2 = ctl->part[2][row], a3 = ctl->part[3][row];
    const float2 mine = make_float2((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y));
    const uint32_t slot = smem_u32(&ctl->xstat[par][gi][row]);
    for (int r = 0; r < cpg; ++r) st_dsmem_f2(mapa_cluster(slot, static_cast<uint32_t>(g0 + r)), mine);
  }
  __syncwarp();
  cluster_sync_all();
  float s = 0.f, q = 0.f;
  for (int r = 0; r < cpg; ++r) {   // same order in every CTA: identical statistics everywhere
    const float2 v = ctl->xstat[par][r][row];
    s += v.x;
    q += v.y;
  }
  if (total) *total = make_float2(s, q);
  RowStat st;
  const float inv_n = 1.0f / static_cast<float>(n);
  st.mean = s * inv_n;
  st.rstd = 1.0f / sqrtf(fmaxf(q * inv_n - st.mean * st.mean, 0.f) + eps);
  return st;
}

__device__ __forceinline__ uint4 pack8(const float (&y)[8]) {
  return make_uint4(rowops::bf2(y[0], y[1]), rowops::bf2(y[2], y[3]), rowops::bf2(y[4], y[5]), rowops::bf2(y[6], y[7]));
}

// ---------------------------------------------------------------------------------------------------------------
// Linear -> [LayerNorm over the group's CTAs] -> ELU -> packed bf16 image (+ x_hat / pre-activation and 1/std for the
// backward pass).  fc_nn.py:14-20, rssm.py:179,192.
// ---------------------------------------------------------------------------------------------------------------
struct LnActOut {
  __nv_bfloat16* out;       // row block of the output image (row stride 64 elements per k tile, see pk_off)
  __nv_bfloat16* save_pre;  // same geometry, or nullptr
  float* save_rstd;         // [128] of this row block (group already applied), or nullptr
  int out_kpad;
};

__device__ __forceinline__ void ld8f(const float* p, float (&f)[8]) {   // 32-byte aligned shared-memory parameters
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

__device__ __forceinline__ void epi_ln_act(RCtl* ctl, const RLayer& L, int rank, bool ln, const LnActOut& o, uint32_t tmem_d,
                                           int cq, int row, bool row_ok, float eps, int par, unsigned long long* trp) {
  const int gi = rank % L.cpg, g0 = rank - gi;
  const int width = L.width[gi], col0 = L.col0[gi];
  const int n_chunks = L.NC >> 3;
  const int mine = n_chunks > cq ? (n_chunks - cq + 3) >> 2 : 0;
  RowStat st{0.f, 1.f};
  if (ln) {
    float sum = 0.f, sq = 0.f, sumb = 0.f, sqb = 0.f;
    tmem_sweep(tmem_d, cq, mine, [&](const uint32_t (&r)[8], int i) {
      const int c = (cq + 4 * i) * 8;
      if (c >= width) return;
      float b[8];
      ld8f(&ctl->bias[c], b);
      if (c + 8 <= width) {   // complete chunk: no per-element predicates, two accumulator pairs
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const float v0 = __uint_as_float(r[j]) + b[j], v1 = __uint_as_float(r[j + 1]) + b[j + 1];
          sum += v0;
          sumb += v1;
          sq = fmaf(v0, v0, sq);
          sqb = fmaf(v1, v1, sqb);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (c + j < width) {
            const float v = __uint_as_float(r[j]) + b[j];
            sum += v;
            sq = fmaf(v, v, sq);
          }
        }
      }
    });
    tr(trp, 4);
    st = ln_exchange(ctl, cq, row, sum + sumb, sq + sqb, g0, L.cpg, gi, L.n, eps, par);
    tr(trp, 5);
  }
  const float nmr = -st.mean * st.rstd;
  const bool save = o.save_pre != nullptr;
  tmem_sweep(tmem_d, cq, mine, [&](const uint32_t (&r)[8], int i) {
    const int c = (cq + 4 * i) * 8;
    if (c >= width) return;
    float y[8], x[8], b[8];
    ld8f(&ctl->bias[c], b);
    if (ln) {
      float g[8], be[8];
      ld8f(&ctl->gamma[c], g);
      ld8f(&ctl->beta[c], be);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        x[j] = fmaf(__uint_as_float(r[j]) + b[j], st.rstd, nmr);
        y[j] = elu1(fmaf(x[j], g[j], be[j]));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        x[j] = __uint_as_float(r[j]) + b[j];
        y[j] = elu1(x[j]);
      }
    }
    if (c + 8 > width || (save && !row_ok)) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (c + j >= width || (save && !row_ok)) {
          y[j] = 0.f;
          x[j] = 0.f;
        }
      }
    }
    const size_t off = pk_off(row, col0 + c);
    *reinterpret_cast<uint4*>(o.out + off) = pack8(y);
    if (save) *reinterpret_cast<uint4*>(o.save_pre + off) = pack8(x);
  });
  if (gi == L.cpg - 1) {   // padding columns of the image: zeros (they meet zero weights, but must be finite)
    for (int ch = ((col0 + width + 7) >> 3) + cq; ch < (o.out_kpad >> 3); ch += 4) {
      const size_t off = pk_off(row, ch * 8);
      *reinterpret_cast<uint4*>(o.out + off) = make_uint4(0u, 0u, 0u, 0u);
      if (save) *reinterpret_cast<uint4*>(o.save_pre + off) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  if (ln && o.save_rstd && cq == 0 && gi == 0) o.save_rstd[row] = st.rstd;
  tr(trp, 6);
}

// the compiler must not move a read of freshly loaded TMEM registers above the tcgen05.wait::ld that completes them
__device__ __forceinline__ void tie8(uint32_t (&r)[8]) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]));
}
// epi_ln_act with the thread's accumulator columns held in registers between the statistics and the output pass: every
// TMEM load of the thread is in flight at once and TMEM is read once (MAXCH = 8-column chunks per thread, at most)
template <int MAXCH>
__device__ __forceinline__ void epi_ln_act_reg(RCtl* ctl, const RLayer& L, int rank, bool ln, const LnActOut& o, uint32_t tmem_d,
                                               int cq, int row, bool row_ok, float eps, int par, unsigned long long* trp) {
  const int gi = rank % L.cpg, g0 = rank - gi;
  const int width = L.width[gi], col0 = L.col0[gi];
  const int n_chunks = L.NC >> 3;
  const int mine = n_chunks > cq ? (n_chunks - cq + 3) >> 2 : 0;
  uint32_t r[MAXCH][8];
#pragma unroll
  for (int i = 0; i < MAXCH; ++i)
    if (i < mine) tmem_ld8(tmem_d + static_cast<uint32_t>((cq + 4 * i) * 8), r[i]);
  tmem_ld_wait();
  float sum = 0.f, sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXCH; ++i) {
    const int c = (cq + 4 * i) * 8;
    if (i < mine && c < width) {
      tie8(r[i]);
      float b[8];
      ld8f(&ctl->bias[c], b);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = __uint_as_float(r[i][j]) + b[j];
        r[i][j] = __float_as_uint(v);   // the pre-activation stays in the register the accumulator arrived in
        if (c + j < width) {
          sum += v;
          sq = fmaf(v, v, sq);
        }
      }
    }
  }
  RowStat st{0.f, 1.f};
  if (ln) {
    tr(trp, 4);
    st = ln_exchange(ctl, cq, row, sum, sq, g0, L.cpg, gi, L.n, eps, par);
    tr(trp, 5);
  }
  const float nmr = -st.mean * st.rstd;
  const bool save = o.save_pre != nullptr;
#pragma unroll
  for (int i = 0; i < MAXCH; ++i) {
    const int c = (cq + 4 * i) * 8;
    if (i < mine && c < width) {
      float y[8], x[8], g[8], be[8];
      if (ln) {
        ld8f(&ctl->gamma[c], g);
        ld8f(&ctl->beta[c], be);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a = __uint_as_float(r[i][j]);
        if (ln) {
          a = fmaf(a, st.rstd, nmr);
          x[j] = a;
          a = fmaf(a, g[j], be[j]);
        } else {
          x[j] = a;
        }
        y[j] = elu1(a);
        if (c + j >= width || (save && !row_ok)) {
          y[j] = 0.f;
          x[j] = 0.f;
        }
      }
      const size_t off = pk_off(row, col0 + c);
      *reinterpret_cast<uint4*>(o.out + off) = pack8(y);
      if (save) *reinterpret_cast<uint4*>(o.save_pre + off) = pack8(x);
    }
  }
  if (gi == L.cpg - 1) {   // padding columns of the image: zeros (they meet zero weights, but must be finite)
    for (int ch = ((col0 + width + 7) >> 3) + cq; ch < (o.out_kpad >> 3); ch += 4) {
      const size_t off = pk_off(row, ch * 8);
      *reinterpret_cast<uint4*>(o.out + off) = make_uint4(0u, 0u, 0u, 0u);
      if (save) *reinterpret_cast<uint4*>(o.save_pre + off) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  if (ln && o.save_rstd && cq == 0 && gi == 0) o.save_rstd[row] = st.rstd;
  tr(trp, 6);
}
// size dispatch: chunks per thread = ceil(NC / 32)
__device__ __forceinline__ void epi_ln_act_any(RCtl* ctl, const RLayer& L, int rank, bool ln, const LnActOut& o, uint32_t tmem_d,
                                               int cq, int row, bool row_ok, float eps, int par, unsigned long long* trp) {
  // (the register-resident variant costs more in spills than it saves in TMEM reads while the whole rollout is one
  // function compiled for 96 registers per thread: kept for reference, switched off)
  constexpr bool kRegEpilogue = false;
  const int per_thread = (L.NC + 31) >> 5;
  if (kRegEpilogue && per_thread <= 4) epi_ln_act_reg<4>(ctl, L, rank, ln, o, tmem_d, cq, row, row_ok, eps, par, trp);
  else epi_ln_act(ctl, L, rank, ln, o, tmem_d, cq, row, row_ok, eps, par, trp);
}

// Linear -> fp32 row-major (fc_nn.py:21 head outputs; rssm.py:192 prior logits)
__device__ __forceinline__ void epi_plain(RCtl* ctl, const RLayer& L, int rank, float* orow, int n_valid, uint32_t tmem_d,
                                          int cq, bool row_ok) {
  const int n_chunks = L.NC >> 3;
  const int mine = n_chunks > cq ? (n_chunks - cq + 3) >> 2 : 0;
  tmem_sweep(tmem_d, cq, mine, [&](const uint32_t (&r)[8], int i) {
    const int c = (cq + 4 * i) * 8;
    if (c >= n_valid || !row_ok) return;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]) + ctl->bias[c + j];
    if (c + 8 <= n_valid) {
      *reinterpret_cast<float4*>(orow + c) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(orow + c + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c + j < n_valid) orow[c + j] = v[j];
    }
  });
}

// GRU contraction epilogue (common.py:69-81): joint LayerNorm over the 3D pre-activations (all CTAs), gates, update
struct GruOut {
  const float* h_prev;   // determ[t] row (D floats)
  float* h_next;         // determ[t+1] row
  __nv_bfloat16* himg;   // row block of the packed h image of step t+1
  float* tape_pre;       // tape: this row of the fp32 pre-activations [3D] or nullptr
  float2* tape_stats;    // tape: [nb][m_pad] partial statistics, this row block's first row, or nullptr
  int tape_nb, m_pad;
};
constexpr int kGruChunks = 5;   // 8-column chunks of hidden units per epilogue thread (wpad <= 160)
// h_prev of this thread's hidden units, fetched while the contraction is still running
template <int NCH>
__device__ __forceinline__ void gru_prefetch_h(const RLayer& L, int rank, const float* h_prev, int cq, bool row_ok,
                                               float (&hp)[NCH][8]) {
  const int width = L.width[rank], col0 = L.col0[rank];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = (cq + 4 * i) * 8;
    if (row_ok && c + 8 <= width) {
      const float4 h0 = __ldcg(reinterpret_cast<const float4*>(h_prev + col0 + c));
      const float4 h1 = __ldcg(reinterpret_cast<const float4*>(h_prev + col0 + c + 4));
      hp[i][0] = h0.x; hp[i][1] = h0.y; hp[i][2] = h0.z; hp[i][3] = h0.w;
      hp[i][4] = h1.x; hp[i][5] = h1.y; hp[i][6] = h1.z; hp[i][7] = h1.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) hp[i][j] = (row_ok && c + j < width) ? __ldcg(h_prev + col0 + c + j) : 0.f;
    }
  }
}
__device__ __forceinline__ void epi_gru(RCtl* ctl, const RLayer& L, int rank, int D, int Dp, const GruOut& o, uint32_t tmem_d,
                                        int cq, int row, bool row_ok, float eps, int par, const float (&hpre)[kGruChunks][8],
                                        unsigned long long* trp) {
  const int width = L.width[rank], col0 = L.col0[rank], wpad = L.wpad;
  const int n_chunks = wpad >> 3;
  float sum = 0.f, sq = 0.f;
  for (int jc = cq; jc < n_chunks; jc += 4) {
    const int c = jc * 8;
    if (c >= width) break;
    uint32_t r0[8], r1[8], r2[8];
    tmem_ld8(tmem_d + static_cast<uint32_t>(c), r0);
    tmem_ld8(tmem_d + static_cast<uint32_t>(wpad + c), r1);
    tmem_ld8(tmem_d + static_cast<uint32_t>(2 * wpad + c), r2);
    float b0[8], b1[8], b2[8];
    ld8f(&ctl->bias[c], b0);
    ld8f(&ctl->bias[wpad + c], b1);
    ld8f(&ctl->bias[2 * wpad + c], b2);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (c + j < width) {
        const float a = __uint_as_float(r0[j]) + b0[j];
        const float b = __uint_as_float(r1[j]) + b1[j];
        const float u = __uint_as_float(r2[j]) + b2[j];
        sum += (a + b) + u;
        sq = fmaf(a, a, fmaf(b, b, fmaf(u, u, sq)));
      }
    }
  }
  tr(trp, 4);
  float2 total;
  const RowStat st = ln_exchange(ctl, cq, row, sum, sq, 0, L.cpg, rank, L.n, eps, par, &total);
  tr(trp, 5);
  if (o.tape_stats && cq == 0 && rank == 0) {
    // the backward pass sums the per-block partials of the chained rollout: hand it the total in block 0
    o.tape_stats[row] = total;
    for (int b = 1; b < o.tape_nb; ++b) o.tape_stats[static_cast<size_t>(b) * o.m_pad + row] = make_float2(0.f, 0.f);
  }
  const float nmr = -st.mean * st.rstd;
#pragma unroll
  for (int i = 0; i < kGruChunks; ++i) {
    const int c = (cq + 4 * i) * 8;
    if (c >= width || c >= wpad) break;
    const float (&hp)[8] = hpre[i];
    uint32_t r0[8], r1[8], r2[8];
    tmem_ld8(tmem_d + static_cast<uint32_t>(c), r0);
    tmem_ld8(tmem_d + static_cast<uint32_t>(wpad + c), r1);
    tmem_ld8(tmem_d + static_cast<uint32_t>(2 * wpad + c), r2);
    tmem_ld_wait();
    const bool tape = o.tape_pre != nullptr && row_ok;
    float gate[8], y[8];
    {   // reset gate
      float b[8], g[8], be[8];
      ld8f(&ctl->bias[c], b);
      ld8f(&ctl->gamma[c], g);
      ld8f(&ctl->beta[c], be);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float p = __uint_as_float(r0[j]) + b[j];
        if (tape && c + j < width) o.tape_pre[col0 + c + j] = p;
        gate[j] = rowops::fast_sigmoid(fmaf(fmaf(p, st.rstd, nmr), g[j], be[j]));
      }
    }
    {   // candidate
      float b[8], g[8], be[8];
      ld8f(&ctl->bias[wpad + c], b);
      ld8f(&ctl->gamma[wpad + c], g);
      ld8f(&ctl->beta[wpad + c], be);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float p = __uint_as_float(r1[j]) + b[j];
        if (tape && c + j < width) o.tape_pre[D + col0 + c + j] = p;
        gate[j] = rowops::fast_tanh(gate[j] * fmaf(fmaf(p, st.rstd, nmr), g[j], be[j]));
      }
    }
    {   // update gate and the convex update
      float b[8], g[8], be[8];
      ld8f(&ctl->bias[2 * wpad + c], b);
      ld8f(&ctl->gamma[2 * wpad + c], g);
      ld8f(&ctl->beta[2 * wpad + c], be);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float p = __uint_as_float(r2[j]) + b[j];
        if (tape && c + j < width) o.tape_pre[2 * D + col0 + c + j] = p;
        const float u = rowops::fast_sigmoid(fmaf(fmaf(p, st.rstd, nmr), g[j], be[j]) - 1.0f);
        y[j] = (row_ok && c + j < width) ? u * gate[j] + (1.0f - u) * hp[j] : 0.f;
      }
    }
    if (row_ok) {
      if (c + 8 <= width) {
        *reinterpret_cast<float4*>(o.h_next + col0 + c) = make_float4(y[0], y[1], y[2], y[3]);
        *reinterpret_cast<float4*>(o.h_next + col0 + c + 4) = make_float4(y[4], y[5], y[6], y[7]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c + j < width) o.h_next[col0 + c + j] = y[j];
      }
    }
    *reinterpret_cast<uint4*>(o.himg + pk_off(row, col0 + c)) = pack8(y);
  }
  if (rank == L.cpg - 1) {   // padding columns [D rounded up to 8, Dp) of the h image
    for (int ch = ((D + 7) >> 3) + cq; ch < (Dp >> 3); ch += 4)
      *reinterpret_cast<uint4*>(o.himg + pk_off(row, ch * 8)) = make_uint4(0u, 0u, 0u, 0u);
  }
  tr(trp, 6);
}

// epi_gru for slices of at most 64 hidden units (two 8-column chunks per thread): the 3 x 2 accumulator chunks stay in
// registers between the statistics and the gate pass
__device__ __forceinline__ void epi_gru_reg(RCtl* ctl, const RLayer& L, int rank, int D, int Dp, const GruOut& o, uint32_t tmem_d,
                                            int cq, int row, bool row_ok, float eps, int par, const float (&hpre)[2][8],
                                            unsigned long long* trp) {
  constexpr int MJ = 2;
  const int width = L.width[rank], col0 = L.col0[rank], wpad = L.wpad;
  uint32_t r[3][MJ][8];
#pragma unroll
  for (int i = 0; i < MJ; ++i) {
    const int c = (cq + 4 * i) * 8;
    if (c < width) {
#pragma unroll
      for (int gt = 0; gt < 3; ++gt) tmem_ld8(tmem_d + static_cast<uint32_t>(gt * wpad + c), r[gt][i]);
    }
  }
  tmem_ld_wait();
  float sum = 0.f, sq = 0.f;
#pragma unroll
  for (int i = 0; i < MJ; ++i) {
    const int c = (cq + 4 * i) * 8;
    if (c < width) {
#pragma unroll
      for (int gt = 0; gt < 3; ++gt) {
        tie8(r[gt][i]);
        float b[8];
        ld8f(&ctl->bias[gt * wpad + c], b);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v = __uint_as_float(r[gt][i][j]) + b[j];
          r[gt][i][j] = __float_as_uint(v);
          if (c + j < width) {
            sum += v;
            sq = fmaf(v, v, sq);
          }
        }
      }
    }
  }
  tr(trp, 4);
  float2 total;
  const RowStat st = ln_exchange(ctl, cq, row, sum, sq, 0, L.cpg, rank, L.n, eps, par, &total);
  tr(trp, 5);
  if (o.tape_stats && cq == 0 && rank == 0) {
    o.tape_stats[row] = total;
    for (int b = 1; b < o.tape_nb; ++b) o.tape_stats[static_cast<size_t>(b) * o.m_pad + row] = make_float2(0.f, 0.f);
  }
  const float nmr = -st.mean * st.rstd;
#pragma unroll
  for (int i = 0; i < MJ; ++i) {
    const int c = (cq + 4 * i) * 8;
    if (c < width) {
      float gr[8], gc[8], gu[8], br[8], bc[8], bu[8], y[8];
      ld8f(&ctl->gamma[c], gr);
      ld8f(&ctl->gamma[wpad + c], gc);
      ld8f(&ctl->gamma[2 * wpad + c], gu);
      ld8f(&ctl->beta[c], br);
      ld8f(&ctl->beta[wpad + c], bc);
      ld8f(&ctl->beta[2 * wpad + c], bu);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float pr = __uint_as_float(r[0][i][j]), pc = __uint_as_float(r[1][i][j]), pu = __uint_as_float(r[2][i][j]);
        if (o.tape_pre && row_ok && c + j < width) {
          o.tape_pre[col0 + c + j] = pr;
          o.tape_pre[D + col0 + c + j] = pc;
          o.tape_pre[2 * D + col0 + c + j] = pu;
        }
        const float rg = rowops::fast_sigmoid(fmaf(fmaf(pr, st.rstd, nmr), gr[j], br[j]));
        const float cand = rowops::fast_tanh(rg * fmaf(fmaf(pc, st.rstd, nmr), gc[j], bc[j]));
        const float u = rowops::fast_sigmoid(fmaf(fmaf(pu, st.rstd, nmr), gu[j], bu[j]) - 1.0f);
        y[j] = (row_ok && c + j < width) ? u * cand + (1.0f - u) * hpre[i][j] : 0.f;
      }
      if (row_ok) {
        if (c + 8 <= width) {
          *reinterpret_cast<float4*>(o.h_next + col0 + c) = make_float4(y[0], y[1], y[2], y[3]);
          *reinterpret_cast<float4*>(o.h_next + col0 + c + 4) = make_float4(y[4], y[5], y[6], y[7]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (c + j < width) o.h_next[col0 + c + j] = y[j];
        }
      }
      *reinterpret_cast<uint4*>(o.himg + pk_off(row, col0 + c)) = pack8(y);
    }
  }
  if (rank == L.cpg - 1) {   // padding columns [D rounded up to 8, Dp) of the h image
    for (int ch = ((D + 7) >> 3) + cq; ch < (Dp >> 3); ch += 4)
      *reinterpret_cast<uint4*>(o.himg + pk_off(row, ch * 8)) = make_uint4(0u, 0u, 0u, 0u);
  }
  tr(trp, 6);
}

// ---------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) rollout_kernel(const __grid_constant__ RolloutParams P, const int ring_bytes) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  RCtl* ctl = reinterpret_cast<RCtl*>(ring + ring_bytes);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int C = P.C;
  const int rank = static_cast<int>(cluster_ctarank());
  const int rb = static_cast<int>(blockIdx.x) / C;   // row block of this cluster
  pdl_launch_dependents();

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < 8; ++s) {
        mbar_init(&ctl->full[s], 1);
        mbar_init(&ctl->empty[s], 1);
      }
      mbar_init(&ctl->tmem_full, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(&ctl->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();
  const uint32_t tmem_base = ctl->tmem_base;
  pdl_wait();

  Roles st{0xffu, 0u, 0u};
  const int q = warp & 3;                 // TMEM lane quarter this warp may read
  const int cq = (warp - 2) >> 2;         // column quarter (epilogue warps)
  const int row = q * 32 + lane;          // row of the block this epilogue thread owns
  const int tid_e = static_cast<int>(threadIdx.x) - 64;
  const int we = warp - 2;                // epilogue warp index 0..15 (row phases)
  const uint32_t tmem_d = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
  const int m = rb * kTileM + row;
  const bool row_ok = m < P.M;
  const size_t blk_D = static_cast<size_t>(rb) * (P.Dp >> 6) * (kTileM * kTileK);
  const size_t blk_S = static_cast<size_t>(rb) * (P.Sp >> 6) * (kTileM * kTileK);
  const size_t blk_A = static_cast<size_t>(rb) * (P.Ap >> 6) * (kTileM * kTileK);
  const size_t blk_H = static_cast<size_t>(rb) * (P.Hp >> 6) * (kTileM * kTileK);
  const size_t hid_gs = static_cast<size_t>(P.m_pad) * P.Hp;
  const size_t ND = static_cast<size_t>(P.M) * P.D, NS = static_cast<size_t>(P.M) * P.S;
  int lnpar = 0;   // parity of the xstat slot (alternates per LayerNorm layer: a peer may still be reading the other)

  auto himg = [&](int t) { return P.himg + static_cast<size_t>(P.himg_pingpong ? (t & 1) : t) * P.himg_step + blk_D; };
  auto zimg = [&](int t) { return P.zimg + static_cast<size_t>(P.zimg_pingpong ? (t & 1) : t) * P.zimg_step + blk_S; };
  auto tp = [&](int t, size_t off) { return P.tape + static_cast<size_t>(t) * P.tape_step + off; };

  int phase = 0;   // phase index inside the step (profiling stamps)
  const bool tracer = P.trace != nullptr && blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 32 || threadIdx.x == 64);
  unsigned long long* trp = nullptr;   // this phase's stamps: 0 producer starts, 1 first stage landed, 2 MMAs issued, 3 accumulator
                                       // ready, 4 statistics pass done, 5 statistics exchanged, 6 outputs stored, 7 phase finished
  auto stamp = [&](int t, int ph) {    // call at the top of phase `ph` of step t
    trp = tracer ? P.trace + (static_cast<size_t>(t) * 11 + ph) * 8 : nullptr;
  };
  // one contraction layer, every role: `active` = this CTA has a slab in it
  auto run_mainloop = [&](const RLayer& L, const OpA& a, bool active) {
    // whole warps in uniform control flow; one elected lane issues the bulk copies / tcgen05 instructions (their operands
    // then live in uniform registers: no ELECT / R2UR.BROADCAST loop around every UBLKCP / UTCHMMA)
    if (warp == 0) {
      if (active) produce(ctl, ring, ring_bytes, L, rank, a, st, P.dbg, P.kg_max, trp);
    } else if (warp == 1) {
      if (active) issue_mma(ctl, ring, ring_bytes, L, tmem_base, a, st, P.dbg, P.kg_max, trp);
    }
  };

  for (int t = 0; t <= P.H; ++t) {
    phase = 0;
    // ================= heads on s_t = cat[h_t, z_t]: actor, reward, discount, target critic =================
    const bool crit_only = P.last_step_value_only != 0 && t == P.H && !P.tape && P.g_critic >= 0;
    for (int l = 0; l < 5; ++l) {
      stamp(t, phase);
      const RLayer& L = P.head[l];
      const int grp = rank / L.cpg;
      const bool active = rank < L.ranks && (!crit_only || grp == P.g_critic);
      OpA a;
      if (l == 0) {
        a.nseg = 2;
        a.A[0] = himg(t); a.kt[0] = P.Dp >> 6;
        a.A[1] = zimg(t); a.kt[1] = P.Sp >> 6;
      } else {
        a.nseg = 1;
        a.A[0] = P.hid[(l - 1) & 1] + static_cast<size_t>(grp) * hid_gs + blk_H; a.kt[0] = P.Hp >> 6;
        a.A[1] = nullptr; a.kt[1] = 0;
      }
      run_mainloop(L, a, active);
      if (l < 4) {
        const bool ln = (l == 0) || P.layer_norm;
        if (warp >= 2 && active) {
          epi_begin(ctl, L, rank, tid_e, ln, st);
          tr(trp, 3);
          LnActOut o;
          o.out = P.hid[l & 1] + static_cast<size_t>(grp) * hid_gs + blk_H;
          o.out_kpad = P.Hp;
          o.save_pre = P.tape ? reinterpret_cast<__nv_bfloat16*>(tp(t, P.tp_head_pre[l])) + static_cast<size_t>(grp) * hid_gs + blk_H
                              : nullptr;
          o.save_rstd = (P.tape && ln) ? reinterpret_cast<float*>(tp(t, P.tp_head_rstd[l])) + static_cast<size_t>(grp) * P.m_pad +
                                             rb * kTileM
                                       : nullptr;
          epi_ln_act_any(ctl, L, rank, ln, o, tmem_d, cq, row, row_ok, P.eps, lnpar, trp);
        } else if (ln) {
          __syncwarp();
          cluster_sync_all();   // the statistics exchange of the CTAs that do have work
        }
        if (ln) lnpar ^= 1;
      } else if (warp >= 2 && active) {
        epi_begin(ctl, L, rank, tid_e, false, st);
        tr(trp, 3);
        float* orow = P.head_out + (static_cast<size_t>(grp) * P.m_pad + m) * 32;
        epi_plain(ctl, L, rank, orow, grp == P.g_actor ? P.Aout : 1, tmem_d, cq, row_ok);
      }
      layer_end(warp >= 2 && active);
      tr(trp, 7);
      ++phase;
    }
    // ---- reward / value / discount read-out and the action draw: one warp per row, rows split over the cluster ----
    stamp(t, phase);
    if (warp >= 2) {
      const int rpc = kTileM / C;
      const bool want_action = t < P.H && !crit_only;
      for (int r = we; r < rpc; r += 16) {
        const int rr = rank * rpc + r;
        const int mm = rb * kTileM + rr;
        const bool valid = mm < P.M;
        const size_t gs = static_cast<size_t>(P.m_pad) * 32;
        if (valid && lane == 0) {
          P.rewards[static_cast<size_t>(t) * P.M + mm] =
              (P.g_reward >= 0 && !crit_only) ? __ldcg(P.head_out + P.g_reward * gs + static_cast<size_t>(mm) * 32) : 0.f;
          if (P.g_critic >= 0 && P.values)
            P.values[static_cast<size_t>(t) * P.M + mm] = __ldcg(P.head_out + P.g_critic * gs + static_cast<size_t>(mm) * 32);
          float d = 1.0f;
          if (P.g_discount >= 0 && t > 0 && !crit_only) {
            // torch Bernoulli(logits).mode: (probs >= 0.5), NaN where probs == 0.5 (world_model.py:137)
            const float x = __ldcg(P.head_out + P.g_discount * gs + static_cast<size_t>(mm) * 32);
            const float pr = 1.0f / (1.0f + expf(-x));
            d = pr > 0.5f ? 1.0f : (pr == 0.5f ? (P.nan_on_tie ? __int_as_float(0x7fc00000) : 1.0f) : 0.0f);
          }
          P.discounts[static_cast<size_t>(t) * P.M + mm] = d;
        }
        if (!want_action) continue;
        const float* ao = P.head_out + P.g_actor * gs + static_cast<size_t>(mm) * 32;
        const int k = lane;
        float act = 0.f;
        if (valid && P.precomp) {
          if (k < P.A) act = __ldg(P.precomp + (static_cast<size_t>(t) * P.M + mm) * P.A + k);
        } else if (valid && P.discrete) {
          // argmax_k fl(logit_k + G(u_k)), first maximum wins (aten::multinomial's exponential race on logits)
          float s = -INFINITY;
          if (k < P.A) {
            const float u = P.action_noise
                                ? __ldg(P.action_noise + (static_cast<size_t>(t) * P.M + mm) * P.A + k)
                                : rlsb_noise_uniform(P.seed_ptr ? __ldg(P.seed_ptr) : P.seed, P.row_offset + mm,
                                                     static_cast<uint32_t>(t), 1u, k);
            const float lg = __ldcg(ao + k);
            s = __fadd_rn(lg, rlsb_gumbel(u));
            if (P.actor_raw) P.actor_raw[(static_cast<size_t>(t) * P.M + mm) * P.A + k] = lg;
            if (s != s) s = (k == 0) ? INFINITY : -INFINITY;   // the sequential scan never replaces with / from a NaN
          }
          float bs = s;
          int bk = k;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, bs, o);
            const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
            if (os > bs || (os == bs && ok < bk)) {
              bs = os;
              bk = ok;
            }
          }
          act = (k == bk) ? 1.0f : 0.f;
        } else if (valid) {
          // TruncatedNormal(tanh(mu), 2 sigmoid(s / 2) + 0.1).rsample() == the unclamped Normal (dists.py:108-129)
          if (k < P.A) {
            const float a_mu = __ldcg(ao + k), a_sd = __ldcg(ao + P.A + k);
            const float mu = tanhf(a_mu);
            const float sd = 2.0f * (1.0f / (1.0f + expf(-a_sd * 0.5f))) + 0.1f;
            NoiseSpec ns{};
            ns.explicit_noise = P.action_noise ? P.action_noise + static_cast<size_t>(t) * P.M * P.A : nullptr;
            ns.ld = P.A; ns.seed = P.seed; ns.seed_ptr = P.seed_ptr; ns.step = static_cast<uint32_t>(t);
            ns.row_offset = P.row_offset;
            act = mu + rowops::noise_normal(ns, mm, 1u, k) * sd;
            if (P.actor_raw) {
              P.actor_raw[(static_cast<size_t>(t) * P.M + mm) * 2 * P.A + k] = a_mu;
              P.actor_raw[(static_cast<size_t>(t) * P.M + mm) * 2 * P.A + P.A + k] = a_sd;
            }
          }
        }
        if (valid && k < P.A) P.actions[(static_cast<size_t>(t + 1) * P.M + mm) * P.A + k] = act;
        // the action operand of img_in: row rr of this block's image (Ap = 64 columns), zeros beyond A and in padding rows
        __nv_bfloat16* arow = P.abf + blk_A;
        arow[pk_off(rr, k)] = __float2bfloat16_rn((valid && k < P.A) ? act : 0.f);
        arow[pk_off(rr, k + 32)] = __float2bfloat16_rn(0.f);
      }
    }
    layer_end(warp >= 2);
    tr(trp, 7);
    ++phase;
    if (t == P.H) break;

    // ================= x = ELU(LN?(W_in [z, a] + b))                                    rssm.py:179 =================
    {
      stamp(t, phase);
      const RLayer& L = P.img_in;
      const bool active = rank < L.ranks;
      OpA a;
      a.nseg = 2;
      a.A[0] = zimg(t); a.kt[0] = P.Sp >> 6;
      a.A[1] = P.abf + blk_A; a.kt[1] = P.Ap >> 6;
      run_mainloop(L, a, active);
      const bool ln = P.layer_norm != 0;
      if (warp >= 2 && active) {
        epi_begin(ctl, L, rank, tid_e, ln, st);
          tr(trp, 3);
        LnActOut o;
        o.out = P.xbf + blk_D;
        o.out_kpad = P.Dp;
        o.save_pre = P.tape ? reinterpret_cast<__nv_bfloat16*>(tp(t + 1, P.tp_x_pre)) + blk_D : nullptr;
        o.save_rstd = (P.tape && ln) ? reinterpret_cast<float*>(tp(t + 1, P.tp_x_rstd)) + rb * kTileM : nullptr;
        epi_ln_act_any(ctl, L, rank, ln, o, tmem_d, cq, row, row_ok, P.eps, lnpar, trp);
      } else if (ln) {
        __syncwarp();
        cluster_sync_all();
      }
      if (ln) lnpar ^= 1;
      layer_end(warp >= 2 && active);
      tr(trp, 7);
      ++phase;
    }
    // ================= h' = GRU(x, h)                                      rssm.py:181, common.py:69-81 =================
    {
      stamp(t, phase);
      const RLayer& L = P.gru;
      const bool active = rank < L.ranks;
      OpA a;
      a.nseg = 2;
      a.A[0] = P.xbf + blk_D; a.kt[0] = P.Dp >> 6;
      a.A[1] = himg(t); a.kt[1] = P.Dp >> 6;
      run_mainloop(L, a, active);
      if (warp >= 2 && active) {
        const float* hprev = P.determ + static_cast<size_t>(t) * ND + static_cast<size_t>(m) * P.D;
        GruOut o;
        o.h_prev = hprev;
        o.h_next = P.determ + static_cast<size_t>(t + 1) * ND + static_cast<size_t>(m) * P.D;
        o.himg = himg(t + 1);
        o.tape_pre = P.tape ? reinterpret_cast<float*>(tp(t + 1, P.tp_gru_scratch)) + static_cast<size_t>(m) * P.tape_ld_scratch
                            : nullptr;
        o.tape_stats = P.tape ? reinterpret_cast<float2*>(tp(t + 1, P.tp_gru_stats)) + rb * kTileM : nullptr;
        o.tape_nb = P.tape_gru_nb;
        o.m_pad = P.m_pad;
        if (false && L.wpad <= 64) {
          float hp[2][8];
          gru_prefetch_h<2>(L, rank, hprev, cq, row_ok, hp);
          epi_begin(ctl, L, rank, tid_e, true, st);
          tr(trp, 3);
          epi_gru_reg(ctl, L, rank, P.D, P.Dp, o, tmem_d, cq, row, row_ok, P.eps, lnpar, hp, trp);
        } else {
          float hp[kGruChunks][8];
          gru_prefetch_h<kGruChunks>(L, rank, hprev, cq, row_ok, hp);
          epi_begin(ctl, L, rank, tid_e, true, st);
          tr(trp, 3);
          epi_gru(ctl, L, rank, P.D, P.Dp, o, tmem_d, cq, row, row_ok, P.eps, lnpar, hp, trp);
        }
      } else {
        __syncwarp();
        cluster_sync_all();
      }
      lnpar ^= 1;
      layer_end(warp >= 2 && active);
      tr(trp, 7);
      ++phase;
    }
    // ================= prior logits = W2 ELU(LN?(W1 h' + b1)) + b2                         rssm.py:192 =================
    {
      stamp(t, phase);
      const RLayer& L = P.prior1;
      const bool active = rank < L.ranks;
      OpA a;
      a.nseg = 1;
      a.A[0] = himg(t + 1); a.kt[0] = P.Dp >> 6;
      a.A[1] = nullptr; a.kt[1] = 0;
      run_mainloop(L, a, active);
      const bool ln = P.layer_norm != 0;
      if (warp >= 2 && active) {
        epi_begin(ctl, L, rank, tid_e, ln, st);
          tr(trp, 3);
        LnActOut o;
        o.out = P.ybf + blk_D;
        o.out_kpad = P.Dp;
        o.save_pre = P.tape ? reinterpret_cast<__nv_bfloat16*>(tp(t + 1, P.tp_y_pre)) + blk_D : nullptr;
        o.save_rstd = (P.tape && ln) ? reinterpret_cast<float*>(tp(t + 1, P.tp_y_rstd)) + rb * kTileM : nullptr;
        epi_ln_act_any(ctl, L, rank, ln, o, tmem_d, cq, row, row_ok, P.eps, lnpar, trp);
      } else if (ln) {
        __syncwarp();
        cluster_sync_all();
      }
      if (ln) lnpar ^= 1;
      layer_end(warp >= 2 && active);
      tr(trp, 7);
      ++phase;
    }
    {
      stamp(t, phase);
      const RLayer& L = P.prior2;
      const bool active = rank < L.ranks;
      OpA a;
      a.nseg = 1;
      a.A[0] = P.ybf + blk_D; a.kt[0] = P.Dp >> 6;
      a.A[1] = nullptr; a.kt[1] = 0;
      run_mainloop(L, a, active);
      if (warp >= 2 && active) {
        epi_begin(ctl, L, rank, tid_e, false, st);
        tr(trp, 3);
        float* orow = P.logits + static_cast<size_t>(t + 1) * NS + static_cast<size_t>(m) * P.S + L.col0[rank];
        epi_plain(ctl, L, rank, orow, L.width[rank], tmem_d, cq, row_ok);
      }
      layer_end(warp >= 2 && active);
      tr(trp, 7);
      ++phase;
    }
    // ================= z' ~ OneHotCategoricalST(logits)                                     rssm.py:34-37 =================
    stamp(t, phase);
    if (warp >= 2) {
      const int rpc = kTileM / C;
      const long long row0 = static_cast<long long>(rb) * kTileM + rank * rpc;
      long long row1 = row0 + rpc;
      if (row1 > P.M) row1 = P.M;
      if (row0 < row1) {
        rowops::SampleLatentArgs sa{};
        sa.logits = P.logits + static_cast<size_t>(t + 1) * NS;
        sa.ld = P.S;
        sa.M = P.M;
        sa.groups = P.groups;
        sa.noise.explicit_noise = P.latent_uniforms ? P.latent_uniforms + static_cast<size_t>(t) * NS : nullptr;
        sa.noise.ld = P.S;
        sa.noise.seed = P.seed;
        sa.noise.seed_ptr = P.seed_ptr;
        sa.noise.step = static_cast<uint32_t>(t);
        sa.noise.row_offset = P.row_offset;
        sa.idx_out = P.stoch_idx + static_cast<size_t>(t + 1) * P.M * P.groups;
        sa.onehot_packed = P.zimg + static_cast<size_t>(P.zimg_pingpong ? ((t + 1) & 1) : (t + 1)) * P.zimg_step;
        sa.kpad = P.Sp;
        sa.onehot_f32 = P.stoch ? P.stoch + static_cast<size_t>(t + 1) * NS : nullptr;
        sa.ld_f32 = P.S;
        rowops::sample_latent_items<true>(sa, row0 * P.groups, row1 * P.groups, we, 16);
      }
    }
    layer_end(warp >= 2);
    tr(trp, 7);
    ++phase;
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// the backward rollout as ONE persistent kernel (rlsb_rollout_bwd): d loss / d actions through the chain of
// rlsb_imagine_bwd.cu, a thread-block cluster per 128 start states, the same layer machinery as the forward kernel
// ---------------------------------------------------------------------------------------------------------------
struct RolloutBwdParams {
  RLayer t_head[5], t_prior2, t_prior1, t_gru, t_img_in;
  int C, M, m_pad, H;
  int D, S, A, Dp, Sp, Ap, Hp, G3p, Gb, gb0, gb_reward, gb_critic, groups, layer_norm, gru_nb;
  float eps;
  // forward results
  const float *determ, *logits;
  const uint8_t* tape; size_t tape_step;
  size_t tp_head_pre[4], tp_head_rstd[4], tp_x_pre, tp_x_rstd, tp_gru_scratch, tp_gru_stats, tp_y_pre, tp_y_rstd;
  long long tape_ld_scratch;
  const float *gru_gamma, *gru_beta;   // determ_recurrent._norm (3D), un-permuted (the gate kernel's row code reads them)
  // gradients in / out
  const float *g_rewards, *g_values;
  float* g_actions;
  // workspace (k1::BwdWorkspace)
  __nv_bfloat16 *dy4, *dh[2], *g_logits, *dp1, *g_pre, *dp_in;
  float *g_s, *g_hprior, *g_hdirect, *g_hgru, *g_za;
  long long ldS, ldZA;
  int kg_max;
  unsigned long long* trace;   // profiling: [H+1][12 phases][8] %globaltimer stamps of cluster 0 (slots 0-2 producer / MMA,
                               // 7 phase finished), nullptr = off
};

struct BwdEpi {
  const __nv_bfloat16* pre;   // row block of the saved x_hat / pre-activation image (pk_off geometry)
  const float* rstd;          // [128] of this row block (LayerNorm) or nullptr
  __nv_bfloat16* out;         // row block of the output image
  int out_kpad;
};
// (the register-hungry row phases are separate functions: the 64 live values of a softmax group or the gate gradients of a row
// must not inflate the register allocation of the contraction phases around them)
// straight-through softmax backward (rssm.py:34-37; bwdops::st_softmax_bwd_item restated for eight lanes per (row, group):
// a lane holds four classes, max / sum / dot are three-step butterflies) — 4x the parallelism and no 64-value local arrays
__device__ __noinline__ void bwd_softmax_phase(const float* logits, int S, const float* g_s_z, long long ldS, const float* g_za,
                                               long long ldZA, __nv_bfloat16* g_logits, int Sp, int row0, int rows, int M, int groups,
                                               int tid_e) {
  const int lane = tid_e & 31, q = lane & 7;
  const int items = rows * groups;
  for (int base = (tid_e >> 5) * 4; base < items; base += (kEpiThreads >> 5) * 4) {
    int it = base + (lane >> 3);
    const bool live = it < items && row0 + it / groups < M;
    if (!(it < items)) it = items - 1;
    const int mm = min(row0 + it / groups, M - 1), g = it % groups;
    const float4 l4 = __ldg(reinterpret_cast<const float4*>(logits + static_cast<size_t>(mm) * S + g * 32) + q);
    float4 gz = __ldcg(reinterpret_cast<const float4*>(g_s_z + static_cast<size_t>(mm) * ldS + g * 32) + q);
    if (g_za) {
      const float4 b = __ldcg(reinterpret_cast<const float4*>(g_za + static_cast<size_t>(mm) * ldZA + g * 32) + q);
      gz.x += b.x; gz.y += b.y; gz.z += b.z; gz.w += b.w;
    }
    float mx = fmaxf(fmaxf(l4.x, l4.y), fmaxf(l4.z, l4.w));
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float e0 = __expf(l4.x - mx), e1 = __expf(l4.y - mx), e2 = __expf(l4.z - mx), e3 = __expf(l4.w - mx);
    float se = (e0 + e1) + (e2 + e3);
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    const float inv = 1.0f / se;
    e0 *= inv; e1 *= inv; e2 *= inv; e3 *= inv;
    float dot = fmaf(gz.x, e0, fmaf(gz.y, e1, fmaf(gz.z, e2, gz.w * e3)));
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (live) {
      // classes [4q, 4q + 4) of the group: half of a 16-byte chunk of the packed image
      const int col = g * 32 + q * 4;
      __nv_bfloat16* dst = g_logits + packed_index(static_cast<size_t>(mm), static_cast<size_t>(col & ~7), static_cast<size_t>(Sp), kTileM) +
                           (col & 4);
      *reinterpret_cast<uint2*>(dst) = make_uint2(rowops::bf2(e0 * (gz.x - dot), e1 * (gz.y - dot)),
                                                   rowops::bf2(e2 * (gz.z - dot), e3 * (gz.w - dot)));
    }
  }
}

// GRU gates + joint LayerNorm backward (common.py:69-81; bwdops::gru_gate_bwd_row restated): the `wpr` = 16 / rows warps of a
// row take 4-unit slices of the D hidden units (a lane holds at most `kGateUnits` slices), the row sums of the LayerNorm
// backward meet in shared memory, and the gate gradients of pass 1 stay in registers for pass 2.
constexpr int kGateUnits = 2;
__device__ __noinline__ void bwd_gate_phase(const bwdops::GruBwdArgs& a, int row0, int rows, int we, int lane, float2* red) {
  const int D = a.D;
  const int wpr = rows >= 16 ? 1 : 16 / rows;          // warps per row
  const int r_local = we / wpr, half = we - r_local * wpr;
  const bool has_row = r_local < rows;
  const int m = row0 + (has_row ? r_local : 0);
  const bool valid = has_row && m < a.M;
  const int units = D >> 2;
  const float inv_n = 1.0f / static_cast<float>(3 * D);
  float mean = 0.f, rstd = 1.f;
  if (valid) {
    const float2* stp = reinterpret_cast<const float2*>(a.stats);
    float s = 0.f, q = 0.f;
    for (int b = 0; b < a.NB; ++b) {
      const float2 v = __ldg(&stp[static_cast<size_t>(b) * a.m_pad + m]);
      s += v.x;
      q += v.y;
    }
    mean = s * inv_n;
    rstd = 1.0f / sqrtf(fmaxf(q * inv_n - mean * mean, 0.f) + a.eps);
  }
  const float* src = a.scratch + static_cast<size_t>(m) * a.ld;
  float dr[kGateUnits][4], dc[kGateUnits][4], du[kGateUnits][4], xr[kGateUnits][4], xc[kGateUnits][4], xu[kGateUnits][4];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int k = 0; k < kGateUnits; ++k) {
    const int u = half * 32 + lane + k * 32 * wpr;
    if (valid && u < units) {
      const int j0 = u * 4;
      const float4 pr = *reinterpret_cast<const float4*>(src + j0), pc = *reinterpret_cast<const float4*>(src + D + j0),
                   pu = *reinterpret_cast<const float4*>(src + 2 * D + j0);
      const float4 hp = *reinterpret_cast<const float4*>(a.h_prev + static_cast<size_t>(m) * a.ld_h + j0);
      float gh[4] = {0.f, 0.f, 0.f, 0.f};
      for (int i = 0; i < a.n_gh; ++i) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(a.gh[i] + static_cast<size_t>(m) * a.ld_gh[i] + j0));
        gh[0] += v.x; gh[1] += v.y; gh[2] += v.z; gh[3] += v.w;
      }
      const float4 gr4 = __ldg(reinterpret_cast<const float4*>(a.gamma + j0)), gc4 = __ldg(reinterpret_cast<const float4*>(a.gamma + D + j0)),
                   gu4 = __ldg(reinterpret_cast<const float4*>(a.gamma + 2 * D + j0));
      const float4 br4 = __ldg(reinterpret_cast<const float4*>(a.beta + j0)), bc4 = __ldg(reinterpret_cast<const float4*>(a.beta + D + j0)),
                   bu4 = __ldg(reinterpret_cast<const float4*>(a.beta + 2 * D + j0));
      const float vr[4] = {pr.x, pr.y, pr.z, pr.w}, vc[4] = {pc.x, pc.y, pc.z, pc.w}, vu[4] = {pu.x, pu.y, pu.z, pu.w};
      const float vh[4] = {hp.x, hp.y, hp.z, hp.w};
      const float ggr[4] = {gr4.x, gr4.y, gr4.z, gr4.w}, ggc[4] = {gc4.x, gc4.y, gc4.z, gc4.w}, ggu[4] = {gu4.x, gu4.y, gu4.z, gu4.w};
      const float bbr[4] = {br4.x, br4.y, br4.z, br4.w}, bbc[4] = {bc4.x, bc4.y, bc4.z, bc4.w}, bbu[4] = {bu4.x, bu4.y, bu4.z, bu4.w};
      float ghd[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        xr[k][e] = (vr[e] - mean) * rstd;
        xc[k][e] = (vc[e] - mean) * rstd;
        xu[k][e] = (vu[e] - mean) * rstd;
        const float nr = fmaf(xr[k][e], ggr[e], bbr[e]);
        const float nc = fmaf(xc[k][e], ggc[e], bbc[e]);
        const float nu = fmaf(xu[k][e], ggu[e], bbu[e]) + a.update_bias;
        const float r = bwdops::fsig(nr);
        const float c = bwdops::ftanh(r * nc);
        const float uu = bwdops::fsig(nu);
        const float g_u = gh[e] * (c - vh[e]);
        const float g_t = gh[e] * uu * (1.0f - c * c);
        dr[k][e] = g_t * nc * r * (1.0f - r) * ggr[e];
        dc[k][e] = g_t * r * ggc[e];
        du[k][e] = g_u * uu * (1.0f - uu) * ggu[e];
        ghd[e] = gh[e] * (1.0f - uu);
        s1 += dr[k][e] + dc[k][e] + du[k][e];
        s2 += dr[k][e] * xr[k][e] + dc[k][e] * xc[k][e] + du[k][e] * xu[k][e];
      }
      *reinterpret_cast<float4*>(a.g_hdirect + static_cast<size_t>(m) * D + j0) = make_float4(ghd[0], ghd[1], ghd[2], ghd[3]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane == 0) red[we] = make_float2(s1, s2);
  epi_bar(3);
  float t1 = 0.f, t2 = 0.f;
  for (int h = 0; h < wpr; ++h) {
    const float2 v = red[r_local * wpr + h];
    t1 += v.x;
    t2 += v.y;
  }
  const float m1 = t1 * inv_n, m2 = t2 * inv_n;
  if (has_row) {
    const int row = m & 127;
    __nv_bfloat16* img = a.g_pre + static_cast<size_t>(m >> 7) * (a.kpad >> 6) * (kTileM * kTileK);
    auto put4 = [&](int col, float o0, float o1, float o2, float o3) {
      *reinterpret_cast<uint2*>(img + pk_off(row, col & ~7) + (col & 4)) = make_uint2(rowops::bf2(o0, o1), rowops::bf2(o2, o3));
    };
#pragma unroll
    for (int k = 0; k < kGateUnits; ++k) {
      const int u = half * 32 + lane + k * 32 * wpr;
      if (u < units) {
        const int j0 = u * 4;
        if (valid) {
          put4(j0, rstd * (dr[k][0] - m1 - xr[k][0] * m2), rstd * (dr[k][1] - m1 - xr[k][1] * m2),
               rstd * (dr[k][2] - m1 - xr[k][2] * m2), rstd * (dr[k][3] - m1 - xr[k][3] * m2));
          put4(D + j0, rstd * (dc[k][0] - m1 - xc[k][0] * m2), rstd * (dc[k][1] - m1 - xc[k][1] * m2),
               rstd * (dc[k][2] - m1 - xc[k][2] * m2), rstd * (dc[k][3] - m1 - xc[k][3] * m2));
          put4(2 * D + j0, rstd * (du[k][0] - m1 - xu[k][0] * m2), rstd * (du[k][1] - m1 - xu[k][1] * m2),
               rstd * (du[k][2] - m1 - xu[k][2] * m2), rstd * (du[k][3] - m1 - xu[k][3] * m2));
        } else {   // padding rows of the operand image: zeros
          put4(j0, 0.f, 0.f, 0.f, 0.f);
          put4(D + j0, 0.f, 0.f, 0.f, 0.f);
          put4(2 * D + j0, 0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    // padding columns [3D, kpad)
    for (int c = (3 * D >> 2) + half * 32 + lane; c < (a.kpad >> 2); c += 32 * wpr) put4(c * 4, 0.f, 0.f, 0.f, 0.f);
  }
}
template <int kBwdChunks>   // 8-column chunks per epilogue thread, at most: 2 (NC <= 64) or 4 (NC <= 128)
__device__ __forceinline__ void epi_bwd_n(RCtl* ctl, const RLayer& L, int rank, bool ln, const BwdEpi& o, uint32_t tmem_d, int cq,
                                          int row, bool row_ok, int tid_e, int par, Roles& st, unsigned long long* trp) {
  const int gi = rank % L.cpg, g0 = rank - gi;
  const int width = L.width[gi], col0 = L.col0[gi];
  const int n_chunks = L.NC >> 3;
  const int mine = n_chunks > cq ? (n_chunks - cq + 3) >> 2 : 0;
  // the saved image and 1/std come from the forward pass: fetch them while the contraction is still running
  uint4 pb[kBwdChunks];
#pragma unroll
  for (int i = 0; i < kBwdChunks; ++i) {
    const int c = (cq + 4 * i) * 8;
    pb[i] = (i < mine && c < width) ? __ldg(reinterpret_cast<const uint4*>(o.pre + pk_off(row, col0 + c))) : make_uint4(0u, 0u, 0u, 0u);
  }
  const float rstd = (ln && o.rstd) ? __ldg(o.rstd + row) : 1.0f;
  epi_begin(ctl, L, rank, tid_e, ln, st);
  tr(trp, 3);
  uint32_t r[kBwdChunks][8];
#pragma unroll
  for (int i = 0; i < kBwdChunks; ++i)
    if (i < mine) tmem_ld8(tmem_d + static_cast<uint32_t>((cq + 4 * i) * 8), r[i]);
  tmem_ld_wait();
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < kBwdChunks; ++i) {
    const int c = (cq + 4 * i) * 8;
    if (i < mine && c < width) {
      tie8(r[i]);
      float x[8], g[8], be[8];
      const uint32_t w[4] = {pb[i].x, pb[i].y, pb[i].z, pb[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        x[2 * j] = __uint_as_float(w[j] << 16);
        x[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
      }
      if (ln) {
        ld8f(&ctl->gamma[c], g);
        ld8f(&ctl->beta[c], be);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = ln ? fmaf(x[j], g[j], be[j]) : x[j];
        const float d = a > 0.f ? 1.0f : rowops::ex2_approx(a * 1.4426950408889634f);   // ELU'
        float dxh = (row_ok && c + j < width) ? __uint_as_float(r[i][j]) * d : 0.f;
        if (ln) dxh *= g[j];
        s1 += dxh;
        s2 = fmaf(dxh, x[j], s2);
        r[i][j] = __float_as_uint(dxh);
      }
    }
  }
  float m1r = 0.f, m2r = 0.f;
  tr(trp, 4);
  if (ln) {
    float2 total;
    ln_exchange(ctl, cq, row, s1, s2, g0, L.cpg, gi, L.n, 0.f, par, &total);
    tr(trp, 5);
    const float inv_n = 1.0f / static_cast<float>(L.n);
    m1r = -total.x * inv_n * rstd;
    m2r = -total.y * inv_n * rstd;
  }
#pragma unroll
  for (int i = 0; i < kBwdChunks; ++i) {
    const int c = (cq + 4 * i) * 8;
    if (i < mine && c < width) {
      float y[8];
      const uint32_t w[4] = {pb[i].x, pb[i].y, pb[i].z, pb[i].w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dxh = __uint_as_float(r[i][j]);
        const float xh = (j & 1) ? __uint_as_float(w[j >> 1] & 0xffff0000u) : __uint_as_float(w[j >> 1] << 16);
        // rstd * (dxh - mean(dxh) - x_hat * mean(dxh * x_hat))
        y[j] = ln ? ((row_ok && c + j < width) ? fmaf(xh, m2r, fmaf(dxh, rstd, m1r)) : 0.f) : dxh;
      }
      *reinterpret_cast<uint4*>(o.out + pk_off(row, col0 + c)) = pack8(y);
    }
  }
  if (gi == L.cpg - 1) {
    for (int ch = ((col0 + width + 7) >> 3) + cq; ch < (o.out_kpad >> 3); ch += 4)
      *reinterpret_cast<uint4*>(o.out + pk_off(row, ch * 8)) = make_uint4(0u, 0u, 0u, 0u);
  }
  tr(trp, 6);
}

__device__ __noinline__ void epi_bwd(RCtl* ctl, const RLayer& L, int rank, bool ln, const BwdEpi& o, uint32_t tmem_d, int cq,
                                     int row, bool row_ok, int tid_e, int par, Roles& st, unsigned long long* trp) {
  if (L.NC <= 64) epi_bwd_n<2>(ctl, L, rank, ln, o, tmem_d, cq, row, row_ok, tid_e, par, st, trp);
  else epi_bwd_n<4>(ctl, L, rank, ln, o, tmem_d, cq, row, row_ok, tid_e, par, st, trp);
}

__global__ void __launch_bounds__(kThreads, 1) rollout_bwd_kernel(const __grid_constant__ RolloutBwdParams P, const int ring_bytes) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  RCtl* ctl = reinterpret_cast<RCtl*>(ring + ring_bytes);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int C = P.C;
  const int rank = static_cast<int>(cluster_ctarank());
  const int rb = static_cast<int>(blockIdx.x) / C;
  pdl_launch_dependents();
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < 8; ++s) {
        mbar_init(&ctl->full[s], 1);
        mbar_init(&ctl->empty[s], 1);
      }
      mbar_init(&ctl->tmem_full, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(&ctl->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();
  const uint32_t tmem_base = ctl->tmem_base;
  pdl_wait();

  Roles st{0xffu, 0u, 0u};
  const int q = warp & 3;
  const int cq = (warp - 2) >> 2;
  const int row = q * 32 + lane;
  const int tid_e = static_cast<int>(threadIdx.x) - 64;
  const int we = warp - 2;
  const uint32_t tmem_d = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
  const int m = rb * kTileM + row;
  const bool row_ok = m < P.M;
  const size_t tile = static_cast<size_t>(kTileM) * kTileK;
  const size_t blk_64 = static_cast<size_t>(rb) * tile;
  const size_t blk_D = static_cast<size_t>(rb) * (P.Dp >> 6) * tile;
  const size_t blk_S = static_cast<size_t>(rb) * (P.Sp >> 6) * tile;
  const size_t blk_H = static_cast<size_t>(rb) * (P.Hp >> 6) * tile;
  const size_t blk_G = static_cast<size_t>(rb) * (P.G3p >> 6) * tile;
  const size_t hid_gs = static_cast<size_t>(P.m_pad) * P.Hp;
  const size_t ND = static_cast<size_t>(P.M) * P.D, NS = static_cast<size_t>(P.M) * P.S;
  const int rpc = kTileM / C;   // rows of the block per CTA in the row phases
  const bool lncfg = P.layer_norm != 0;
  int lnpar = 0;
  const bool tracer = P.trace != nullptr && blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 32 || threadIdx.x == 64);
  unsigned long long* trp = nullptr;
  int phase = 0;
  auto stamp = [&](int t) { trp = tracer ? P.trace + (static_cast<size_t>(t) * 12 + phase) * 8 : nullptr; };
  auto done = [&]() {
    tr(trp, 7);
    ++phase;
  };
  auto tp = [&](int t, size_t off) { return P.tape + static_cast<size_t>(t) * P.tape_step + off; };
  auto run_mainloop = [&](const RLayer& L, const OpA& a, bool active) {
    if (warp == 0) {
      if (active) produce(ctl, ring, ring_bytes, L, rank, a, st, 0, P.kg_max, trp);
    } else if (warp == 1) {
      if (active) issue_mma(ctl, ring, ring_bytes, L, tmem_base, a, st, 0, P.kg_max, trp);
    }
  };
  // d loss / d a_t = columns [Sp, Sp + A) of the img_in dX the previous step left in g_za
  auto extract_actions = [&](int t_act) {
    if (warp < 2) return;
    for (int i = tid_e; i < rpc * P.A; i += kEpiThreads) {
      const int mm = rb * kTileM + rank * rpc + i / P.A, k = i % P.A;
      if (mm < P.M)
        P.g_actions[(static_cast<size_t>(t_act) * P.M + mm) * P.A + k] = __ldcg(P.g_za + static_cast<size_t>(mm) * P.ldZA + P.Sp + k);
    }
  };

  for (int t = P.H; t >= 1; --t) {
    phase = 0;
    stamp(t);
    // ---- d loss / d (reward, value) of state t -> head gradients; the previous step's action gradient ----------------
    if (t < P.H) extract_actions(t);
    if (warp >= 2 && tid_e < rpc) {
      const int mm = rb * kTileM + rank * rpc + tid_e;
      bwdops::head_grad_row(mm, P.g_rewards + static_cast<size_t>(t) * P.M, P.g_values + static_cast<size_t>(t) * P.M, P.M,
                            P.m_pad, P.Gb, P.gb_reward, P.gb_critic, P.dy4);
    }
    layer_end(warp >= 2);
    done();
    stamp(t);
    // ---- reward head + target critic, layers 4 .. 1: dX with the ELU' / LayerNorm backward of the layer below ---------
    for (int l = 4; l >= 1; --l) {
      const RLayer& L = P.t_head[l];
      const int grp = rank / L.cpg;
      const bool active = rank < L.ranks;
      const bool ln = (l - 1 == 0) || lncfg;
      OpA a;
      a.nseg = 1;
      a.A[0] = (l == 4) ? P.dy4 + static_cast<size_t>(grp) * P.m_pad * 64 + blk_64
                        : P.dh[(l + 1) & 1] + static_cast<size_t>(grp) * hid_gs + blk_H;
      a.kt[0] = L.kt;
      run_mainloop(L, a, active);
      if (warp >= 2 && active) {
        BwdEpi o;
        o.pre = reinterpret_cast<const __nv_bfloat16*>(tp(t, P.tp_head_pre[l - 1])) + static_cast<size_t>(P.gb0 + grp) * hid_gs + blk_H;
        o.rstd = ln ? reinterpret_cast<const float*>(tp(t, P.tp_head_rstd[l - 1])) + static_cast<size_t>(P.gb0 + grp) * P.m_pad + rb * kTileM
                    : nullptr;
        o.out = P.dh[l & 1] + static_cast<size_t>(grp) * hid_gs + blk_H;
        o.out_kpad = P.Hp;
        epi_bwd(ctl, L, rank, ln, o, tmem_d, cq, row, row_ok, tid_e, lnpar, st, trp);
      } else if (ln) {
        __syncwarp();
        cluster_sync_all();
      }
      if (ln) lnpar ^= 1;
      layer_end(warp >= 2 && active);
      done();
      stamp(t);
    }
    // ---- layer 0 of the heads: the groups are K segments of one contraction -> d loss / d [h_t, z_t] ------------------
    {
      const RLayer& L = P.t_head[0];
      const bool active = rank < L.ranks;
      OpA a;
      a.nseg = P.Gb;
      for (int i = 0; i < P.Gb; ++i) {
        a.A[i] = P.dh[1] + static_cast<size_t>(i) * hid_gs + blk_H;
        a.kt[i] = P.Hp >> 6;
      }
      run_mainloop(L, a, active);
      if (warp >= 2 && active) {
        epi_begin(ctl, L, rank, tid_e, false, st);
        epi_plain(ctl, L, rank, P.g_s + static_cast<size_t>(m) * P.ldS + L.col0[rank], L.width[rank], tmem_d, cq, row_ok);
      }
      layer_end(warp >= 2 && active);
      done();
      stamp(t);
    }
    // ---- z_t = onehot + p - p.detach(): d loss / d prior logits ------------------------------------------------------------
    if (warp >= 2)
      bwd_softmax_phase(P.logits + static_cast<size_t>(t) * NS, P.S, P.g_s + P.Dp, P.ldS, t < P.H ? P.g_za : nullptr, P.ldZA,
                        P.g_logits, P.Sp, rb * kTileM + rank * rpc, rpc, P.M, P.groups, tid_e);
    layer_end(warp >= 2);
    done();
    stamp(t);
    // ---- prior MLP: logits -> y (ELU' / LayerNorm backward of prior 1) -> h_t ----------------------------------------------
    {
      const RLayer& L = P.t_prior2;
      const bool active = rank < L.ranks;
      OpA a;
      a.nseg = 1;
      a.A[0] = P.g_logits + blk_S; a.kt[0] = P.Sp >> 6;
      run_mainloop(L, a, active);
      if (warp >= 2 && active) {
        BwdEpi o;
        o.pre = reinterpret_cast<const __nv_bfloat16*>(tp(t, P.tp_y_pre)) + blk_D;
        o.rstd = lncfg ? reinterpret_cast<const float*>(tp(t, P.tp_y_rstd)) + rb * kTileM : nullptr;
        o.out = P.dp1 + blk_D;
        o.out_kpad = P.Dp;
        epi_bwd(ctl, L, rank, lncfg, o, tmem_d, cq, row, row_ok, tid_e, lnpar, st, trp);
      } else if (lncfg) {
        __syncwarp();
        cluster_sync_all();
      }
      if (lncfg) lnpar ^= 1;
      layer_end(warp >= 2 && active);
      done();
      stamp(t);
    }
    {
      const RLayer& L = P.t_prior1;
      const bool active = rank < L.ranks;
      OpA a;
      a.nseg = 1;
      a.A[0] = P.dp1 + blk_D; a.kt[0] = P.Dp >> 6;
      run_mainloop(L, a, active);
      if (warp >= 2 && active) {
        epi_begin(ctl, L, rank, tid_e, false, st);
        epi_plain(ctl, L, rank, P.g_hprior + static_cast<size_t>(m) * P.D + L.col0[rank], L.width[rank], tmem_d, cq, row_ok);
      }
      layer_end(warp >= 2 && active);
      done();
      stamp(t);
    }
    // ---- h_t = GRU(x_t, h_{t-1}): gates + joint LayerNorm backward, one warp per row -------------------------------------------
    if (warp >= 2) {
      bwdops::GruBwdArgs ga{};
      ga.scratch = reinterpret_cast<const float*>(tp(t, P.tp_gru_scratch)); ga.ld = P.tape_ld_scratch;
      ga.stats = reinterpret_cast<const float*>(tp(t, P.tp_gru_stats));
      ga.NB = P.gru_nb; ga.M = P.M; ga.m_pad = P.m_pad; ga.D = P.D;
      ga.gamma = P.gru_gamma; ga.beta = P.gru_beta;
      ga.eps = P.eps; ga.update_bias = -1.0f;
      ga.h_prev = P.determ + static_cast<size_t>(t - 1) * ND; ga.ld_h = P.D;
      ga.gh[0] = P.g_s; ga.ld_gh[0] = P.ldS;
      ga.gh[1] = P.g_hprior; ga.ld_gh[1] = P.D;
      ga.n_gh = 2;
      if (t < P.H) {
        ga.gh[2] = P.g_hdirect; ga.ld_gh[2] = P.D;
        ga.gh[3] = P.g_hgru; ga.ld_gh[3] = P.D;
        ga.n_gh = 4;
      }
      ga.g_pre = P.g_pre; ga.kpad = P.G3p;
      ga.g_hdirect = P.g_hdirect;
      bwd_gate_phase(ga, rb * kTileM + rank * rpc, rpc, we, lane, &ctl->part[0][0]);
    }
    layer_end(warp >= 2);
    done();
    stamp(t);
    // ---- d / d x_t (with img_in's ELU' / LayerNorm backward) and d / d h_{t-1} through the gates: one phase, two groups -------
    {
      const RLayer& L = P.t_gru;
      const int grp = rank / L.cpg;
      const bool active = rank < L.ranks;
      OpA a;
      a.nseg = 1;
      a.A[0] = P.g_pre + blk_G; a.kt[0] = P.G3p >> 6;
      run_mainloop(L, a, active);
      if (warp >= 2 && active && grp == 0) {
        BwdEpi o;
        o.pre = reinterpret_cast<const __nv_bfloat16*>(tp(t, P.tp_x_pre)) + blk_D;
        o.rstd = lncfg ? reinterpret_cast<const float*>(tp(t, P.tp_x_rstd)) + rb * kTileM : nullptr;
        o.out = P.dp_in + blk_D;
        o.out_kpad = P.Dp;
        epi_bwd(ctl, L, rank, lncfg, o, tmem_d, cq, row, row_ok, tid_e, lnpar, st, trp);
      } else {
        if (warp >= 2 && active) {
          epi_begin(ctl, L, rank, tid_e, false, st);
          const int gi = rank % L.cpg;
          epi_plain(ctl, L, rank, P.g_hgru + static_cast<size_t>(m) * P.D + L.col0[gi], L.width[gi], tmem_d, cq, row_ok);
        }
        if (lncfg) {
          __syncwarp();
          cluster_sync_all();
        }
      }
      if (lncfg) lnpar ^= 1;
      layer_end(warp >= 2 && active);
      done();
      stamp(t);
    }
    // ---- x_t = ELU(LN?(W_in [z_{t-1}, a_{t-1}])): d loss / d z_{t-1}, d loss / d a_{t-1} -------------------------------------------
    {
      const RLayer& L = P.t_img_in;
      const bool active = rank < L.ranks;
      OpA a;
      a.nseg = 1;
      a.A[0] = P.dp_in + blk_D; a.kt[0] = P.Dp >> 6;
      run_mainloop(L, a, active);
      if (warp >= 2 && active) {
        epi_begin(ctl, L, rank, tid_e, false, st);
        epi_plain(ctl, L, rank, P.g_za + static_cast<size_t>(m) * P.ldZA + L.col0[rank], L.width[rank], tmem_d, cq, row_ok);
      }
      layer_end(warp >= 2 && active);
      done();
      stamp(t);
    }
  }
  extract_actions(0);

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int g_rollout_cluster = 0;   // 0: not read yet
unsigned long long* g_rollout_trace = nullptr;

int rollout_cluster_size() {
  if (g_rollout_cluster == 0) {
    g_rollout_cluster = 8;
    if (const char* env = getenv("RLSB_ROLLOUT_CLUSTER")) {
      const int c = atoi(env);
      if (c == 4 || c == 8 || c == 16) g_rollout_cluster = c;
    }
  }
  return g_rollout_cluster;
}

#define RLSB_TRY(expr)      \
  do {                      \
    int _e = (expr);        \
    if (_e != 0) return _e; \
  } while (0)

void fill_layer(RLayer& d, const HLayer& h, const uint8_t* pk) {
  d.W = reinterpret_cast<const __nv_bfloat16*>(pk + h.w_off);
  d.bias = reinterpret_cast<const float*>(pk + h.bias_off);
  d.gamma = reinterpret_cast<const float*>(pk + h.g_off);
  d.beta = reinterpret_cast<const float*>(pk + h.b_off);
  d.NC = h.NC; d.kt = h.kt; d.cpg = h.cpg; d.ranks = h.ranks; d.n = h.n; d.wpad = h.wpad;
  for (int i = 0; i < kMaxC; ++i) {
    d.col0[i] = h.col0[i];
    d.width[i] = h.width[i];
  }
}

}  // namespace
}  // namespace ro
}  // namespace rlsb

using namespace rlsb;
using namespace rlsb::ro;

extern "C" int rlsb_rollout_cluster_size(void) { return rollout_cluster_size(); }

// clusters of `cluster` CTAs of the persistent rollout kernel the current device keeps resident at once (a cluster needs its
// CTAs inside one GPC: fewer than #SMs / cluster in general); 0 on error.  Cached per cluster size.
extern "C" int rlsb_rollout_max_clusters(int cluster) {
  if (cluster != 4 && cluster != 8 && cluster != 16) return 0;
  static int cached[3] = {-1, -1, -1};
  int& slot = cached[cluster == 4 ? 0 : cluster == 8 ? 1 : 2];
  if (slot >= 0) return slot;
  if (cudaFuncSetAttribute(rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
      cudaFuncSetAttribute(rollout_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  const int ring_bytes = (227 * 1024 - 1024 - static_cast<int>(sizeof(RCtl)) - 256) / 1024 * 1024;
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3(static_cast<unsigned>(cluster * 64));
  lc.blockDim = dim3(kThreads);
  lc.dynamicSmemBytes = static_cast<size_t>(ring_bytes) + sizeof(RCtl) + 1024;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(cluster);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  lc.attrs = attr;
  lc.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, rollout_kernel, &lc) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  slot = n;
  return n;
}

extern "C" void rlsb_rollout_set_trace(void* device_buffer) { g_rollout_trace = static_cast<unsigned long long*>(device_buffer); }

namespace {
int cluster_of(const rlsb_imagine_cfg& cfg) {
  const int c = cfg.rollout_cluster;
  return (c == 4 || c == 8 || c == 16) ? c : rollout_cluster_size();
}
}  // namespace

extern "C" size_t rlsb_rollout_packed_bytes(const rlsb_imagine_cfg* cfg) {
  RPlan R;
  if (!cfg || make_rplan(*cfg, cluster_of(*cfg), R) != 0) return 0;
  return R.bytes;
}

extern "C" int rlsb_rollout_bwd_supported(const rlsb_imagine_cfg* cfg) {
  RPlan R;
  if (!cfg || make_rplan(*cfg, cluster_of(*cfg), R) != 0) return 0;
  return R.bwd ? 1 : 0;
}

extern "C" int rlsb_rollout_pack(const rlsb_imagine_cfg* cfg, const rlsb_imagine_params* prm, void* packed, void* stream_) {
  if (!cfg || !prm || !packed) return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  RPlan R;
  RLSB_TRY(make_rplan(*cfg, cluster_of(*cfg), R));
  const k1::Plan& P = R.P;
  uint8_t* base = static_cast<uint8_t*>(packed);
  RPackJobs J{};
  int blocks = 0;
  auto add = [&](const HLayer& L, int group, const float* w, long long ld, const float* b, const float* g, const float* be,
                 int n_out, int mode, int n_seg, const PackSeg* segs) -> int {
    if (J.n >= kMaxJobs) return -35;
    if (!w) return -20;
    RPackJob& j = J.job[J.n];
    j.w = w; j.ld = ld; j.b = b; j.g = g; j.be = be;
    const size_t r0 = static_cast<size_t>(group) * L.cpg;
    j.W = reinterpret_cast<__nv_bfloat16*>(base + L.w_off) + r0 * L.kt * L.NC * 64;
    j.bias = reinterpret_cast<float*>(base + L.bias_off) + r0 * L.NC;
    j.gamma = reinterpret_cast<float*>(base + L.g_off) + r0 * L.NC;
    j.beta = reinterpret_cast<float*>(base + L.b_off) + r0 * L.NC;
    j.ranks = L.cpg; j.NC = L.NC; j.kt = L.kt; j.n_out = n_out; j.mode = mode; j.D = P.D; j.wpad = L.wpad; j.n_seg = n_seg;
    for (int i = 0; i < n_seg; ++i) j.seg[i] = segs[i];
    for (int i = 0; i < kMaxC; ++i) {
      j.col0[i] = L.col0[i];
      j.width[i] = L.width[i];
    }
    J.first_block[J.n] = blocks;
    const long long chunks = static_cast<long long>(L.cpg) * L.kt * L.NC * 8;
    blocks += static_cast<int>((chunks + 255) / 256);
    ++J.n;
    J.first_block[J.n] = blocks;
    return 0;
  };
  {
    PackSeg sg[2] = {{0, 0, P.S}, {P.Sp, P.S, P.A}};
    RLSB_TRY(add(R.img_in, 0, prm->img_in_w, P.S + P.A, prm->img_in_b, prm->img_in_ln_g, prm->img_in_ln_b, P.D, 0, 2, sg));
  }
  {
    PackSeg sg[2] = {{0, 0, P.D}, {P.Dp, P.D, P.D}};
    RLSB_TRY(add(R.gru, 0, prm->gru_w, 2 * P.D, prm->gru_b, prm->gru_ln_g, prm->gru_ln_b, 3 * P.D, 1, 2, sg));
  }
  {
    PackSeg sg[1] = {{0, 0, P.D}};
    RLSB_TRY(add(R.prior1, 0, prm->prior1_w, P.D, prm->prior1_b, prm->prior1_ln_g, prm->prior1_ln_b, P.D, 0, 1, sg));
    RLSB_TRY(add(R.prior2, 0, prm->prior2_w, P.D, prm->prior2_b, nullptr, nullptr, P.S, 0, 1, sg));
  }
  for (int l = 0; l < 5; ++l) {
    for (int g = 0; g < P.G; ++g) {
      const rlsb_mlp_params* hp = (g == P.g_actor) ? &prm->actor : (g == P.g_reward) ? &prm->reward
                                  : (g == P.g_discount) ? &prm->discount : &prm->critic;
      const int n_out = (l == 4) ? ((g == P.g_actor) ? P.Aout : 1) : P.Hd;
      if (l == 0) {
        PackSeg sg[2] = {{0, 0, P.D}, {P.Dp, P.D, P.S}};
        RLSB_TRY(add(R.head[l], g, hp->w[l], P.D + P.S, hp->b[l], hp->ln_g[l], hp->ln_b[l], n_out, 0, 2, sg));
      } else {
        PackSeg sg[1] = {{0, 0, P.Hd}};
        RLSB_TRY(add(R.head[l], g, hp->w[l], P.Hd, hp->b[l], l < 4 ? hp->ln_g[l] : nullptr, l < 4 ? hp->ln_b[l] : nullptr,
                     n_out, 0, 1, sg));
      }
    }
  }
  if (R.bwd) {
    // transposed slabs: rows = in-features (LayerNorm gamma / beta of the layer BELOW ride along, indexed by the same rows)
    auto addt = [&](const HLayer& L, int group, const float* w, long long ld, const float* g, const float* be, int n_out,
                    int dst_k0, int k_lo, int k_hi, int n_rseg, const RowSeg* rs) -> int {
      PackSeg sg[1] = {{dst_k0, 0, n_out}};
      RLSB_TRY(add(L, group, w, ld, nullptr, g, be, n_out, 2, 1, sg));
      RPackJob& j = J.job[J.n - 1];
      j.n_rseg = n_rseg;
      for (int i = 0; i < n_rseg; ++i) j.rseg[i] = rs[i];
      j.k_lo = k_lo; j.k_hi = k_hi;
      return 0;
    };
    for (int l = 1; l < 5; ++l) {
      for (int gb = 0; gb < P.Gb; ++gb) {
        const int g = P.gb0 + gb;
        const rlsb_mlp_params* hp = (g == P.g_reward) ? &prm->reward : (g == P.g_discount) ? &prm->discount : &prm->critic;
        const int n_out = (l == 4) ? 1 : P.Hd;
        RowSeg rs[1] = {{0, 0, P.Hd}};
        RLSB_TRY(addt(R.t_head[l], gb, hp->w[l], P.Hd, hp->ln_g[l - 1], hp->ln_b[l - 1], n_out, 0, 0, R.t_head[l].kt * 64, 1, rs));
      }
    }
    for (int gb = 0; gb < P.Gb; ++gb) {   // layer 0: the gradient-carrying groups share one K axis
      const int g = P.gb0 + gb;
      const rlsb_mlp_params* hp = (g == P.g_reward) ? &prm->reward : (g == P.g_discount) ? &prm->discount : &prm->critic;
      RowSeg rs[2] = {{0, 0, P.D}, {P.Dp, P.D, P.S}};
      RLSB_TRY(addt(R.t_head[0], 0, hp->w[0], P.D + P.S, nullptr, nullptr, P.Hd, gb * P.Hp, gb * P.Hp, (gb + 1) * P.Hp, 2, rs));
    }
    {
      RowSeg rs[1] = {{0, 0, P.D}};
      RLSB_TRY(addt(R.t_prior2, 0, prm->prior2_w, P.D, prm->prior1_ln_g, prm->prior1_ln_b, P.S, 0, 0, P.Sp, 1, rs));
      RLSB_TRY(addt(R.t_prior1, 0, prm->prior1_w, P.D, nullptr, nullptr, P.D, 0, 0, P.Dp, 1, rs));
      // GRU weight (3D, 2D): in-features [x | h]; group 0 = d / d x (with img_in's LayerNorm below), group 1 = d / d h
      RLSB_TRY(addt(R.t_gru, 0, prm->gru_w, 2 * P.D, prm->img_in_ln_g, prm->img_in_ln_b, 3 * P.D, 0, 0, P.G3p, 1, rs));
      RowSeg rh[1] = {{0, P.D, P.D}};
      RLSB_TRY(addt(R.t_gru, 1, prm->gru_w, 2 * P.D, nullptr, nullptr, 3 * P.D, 0, 0, P.G3p, 1, rh));
      RowSeg ri[2] = {{0, 0, P.S}, {P.Sp, P.S, P.A}};
      RLSB_TRY(addt(R.t_img_in, 0, prm->img_in_w, P.S + P.A, nullptr, nullptr, P.D, 0, 0, P.Dp, 2, ri));
    }
  }
  rpack_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(J);
  count_launch();
  if (R.bwd) {
    float* dst = reinterpret_cast<float*>(base + R.gru_ln_off);
    RLSB_TRY(launch_copy_pad(prm->gru_ln_g, 3 * P.D, dst, 3 * P.D, 1.f, s));
    RLSB_TRY(launch_copy_pad(prm->gru_ln_b, 3 * P.D, dst + 3 * P.D, 3 * P.D, 0.f, s));
  }
  return static_cast<int>(cudaGetLastError());
}

extern "C" int rlsb_rollout_fwd(const rlsb_imagine_cfg* cfg, const void* packed, int64_t N, const float* h0, const float* z0,
                                const float* logits0, const rlsb_noise* noise, const rlsb_imagine_out* out, void* workspace,
                                void* stream_) {
  if (!cfg || !packed || !h0 || !z0 || !noise || !out || !workspace || N <= 0) return -1;
  if (!out->determ || !out->logits || !out->stoch_idx || !out->actions || !out->rewards || !out->discounts) return -2;
  if (N > (1LL << 24)) return -3;
  if (out->actor_slots) return -7;   // the update recomputes the actor forward at these sizes
  if ((out->determ_packed != nullptr) != (out->stoch_packed != nullptr)) return -4;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  const int C = cluster_of(*cfg);
  RPlan R;
  RLSB_TRY(make_rplan(*cfg, C, R));
  const k1::Plan& P = R.P;
  k1::Workspace W;
  k1::make_workspace(P, N, W);
  const int M = static_cast<int>(N);
  const int m_pad = W.m_pad;
  const int H = cfg->H;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  auto bf = [&](size_t off) { return reinterpret_cast<__nv_bfloat16*>(ws + off); };
  const size_t ND = static_cast<size_t>(M) * P.D, NS = static_cast<size_t>(M) * P.S;
  k1::Tape TP{};
  if (out->tape) {
    if (!P.bwd) return -5;
    k1::make_tape(P, N, H, TP);
  }
  const bool keep = out->determ_packed != nullptr;

  RolloutParams rp{};
  for (int l = 0; l < 5; ++l) fill_layer(rp.head[l], R.head[l], pk);
  fill_layer(rp.img_in, R.img_in, pk);
  fill_layer(rp.gru, R.gru, pk);
  fill_layer(rp.prior1, R.prior1, pk);
  fill_layer(rp.prior2, R.prior2, pk);
  rp.C = C; rp.M = M; rp.m_pad = m_pad; rp.H = H;
  rp.D = P.D; rp.S = P.S; rp.A = P.A; rp.Aout = P.Aout; rp.Dp = P.Dp; rp.Sp = P.Sp; rp.Ap = P.Ap; rp.Hp = P.Hp; rp.G = P.G;
  rp.groups = cfg->groups;
  rp.g_actor = P.g_actor; rp.g_reward = P.g_reward; rp.g_discount = P.g_discount; rp.g_critic = P.g_critic;
  rp.discrete = cfg->discrete; rp.layer_norm = cfg->layer_norm; rp.nan_on_tie = cfg->discount_nan_on_tie;
  rp.last_step_value_only = cfg->last_step_value_only;
  rp.eps = 1e-5f;
  if (keep) {
    rp.himg = static_cast<__nv_bfloat16*>(out->determ_packed); rp.himg_step = static_cast<long long>(m_pad) * P.Dp; rp.himg_pingpong = 0;
    rp.zimg = static_cast<__nv_bfloat16*>(out->stoch_packed); rp.zimg_step = static_cast<long long>(m_pad) * P.Sp; rp.zimg_pingpong = 0;
  } else {
    rp.himg = bf(W.hbf[0]); rp.himg_step = static_cast<long long>((W.hbf[1] - W.hbf[0]) / 2); rp.himg_pingpong = 1;
    rp.zimg = bf(W.zbf[0]); rp.zimg_step = static_cast<long long>((W.zbf[1] - W.zbf[0]) / 2); rp.zimg_pingpong = 1;
  }
  rp.abf = bf(W.abf); rp.xbf = bf(W.xbf); rp.ybf = bf(W.ybf); rp.hid[0] = bf(W.hid[0]); rp.hid[1] = bf(W.hid[1]);
  rp.head_out = reinterpret_cast<float*>(ws + W.head_out);
  rp.determ = out->determ; rp.logits = out->logits; rp.stoch = out->stoch; rp.actions = out->actions;
  rp.rewards = out->rewards; rp.discounts = out->discounts; rp.values = out->values; rp.actor_raw = out->actor_raw;
  rp.stoch_idx = out->stoch_idx;
  rp.latent_uniforms = noise->latent_uniforms; rp.action_noise = noise->action_noise; rp.precomp = noise->precomp_actions;
  rp.seed = noise->seed; rp.seed_ptr = noise->seed_device; rp.row_offset = noise->row_offset;
  rp.tape = static_cast<uint8_t*>(out->tape);
  if (rp.tape) {
    rp.tape_step = TP.step_bytes;
    for (int l = 0; l < 4; ++l) {
      rp.tp_head_pre[l] = TP.head_pre[l];
      rp.tp_head_rstd[l] = TP.head_rstd[l];
    }
    rp.tp_x_pre = TP.x_pre; rp.tp_x_rstd = TP.x_rstd; rp.tp_gru_scratch = TP.gru_scratch; rp.tp_gru_stats = TP.gru_stats;
    rp.tp_y_pre = TP.y_pre; rp.tp_y_rstd = TP.y_rstd;
    rp.tape_ld_scratch = TP.ld_scratch;
    rp.tape_gru_nb = P.gru.NB;
  }
  rp.trace = g_rollout_trace;
  if (const char* env = getenv("RLSB_ROLLOUT_DEBUG")) rp.dbg = atoi(env);
  rp.kg_max = 4;
  if (const char* env = getenv("RLSB_ROLLOUT_KG")) rp.kg_max = atoi(env) >= 1 ? atoi(env) : 1;
  auto himg = [&](int t) { return rp.himg + static_cast<size_t>(rp.himg_pingpong ? (t & 1) : t) * rp.himg_step; };
  auto zimg = [&](int t) { return rp.zimg + static_cast<size_t>(rp.zimg_pingpong ? (t & 1) : t) * rp.zimg_step; };

  // ---- start state (as rlsb_imagine_fwd) ----------------------------------------------------------------------
  {
    PackSeg seg[1] = {{0, 0, P.D}};
    RLSB_TRY(launch_pack(h0, P.D, M, himg(0), 128, m_pad, P.Dp, 1, seg, s));
    PackSeg segz[1] = {{0, 0, P.S}};
    RLSB_TRY(launch_pack(z0, P.S, M, zimg(0), 128, m_pad, P.Sp, 1, segz, s));
    const size_t tile_row_bytes = static_cast<size_t>(P.Sp / 64) * 128 * 64 * 2;
    cudaError_t e = cudaSuccess;
    if (M != m_pad) {   // rows >= N of the one-hot images are never written by the draw: clear their last row block
      for (int t = 1; t <= (keep ? H : 1) && e == cudaSuccess; ++t)
        e = cudaMemsetAsync(reinterpret_cast<uint8_t*>(zimg(t)) + static_cast<size_t>(m_pad / 128 - 1) * tile_row_bytes, 0,
                            tile_row_bytes, s);
    }
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaMemcpyAsync(out->determ, h0, ND * 4, cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) return static_cast<int>(e);
    if (logits0) e = cudaMemcpyAsync(out->logits, logits0, NS * 4, cudaMemcpyDeviceToDevice, s);
    else e = cudaMemsetAsync(out->logits, 0, NS * 4, s);
    if (e != cudaSuccess) return static_cast<int>(e);
    if (out->stoch) {
      e = cudaMemcpyAsync(out->stoch, z0, NS * 4, cudaMemcpyDeviceToDevice, s);
      if (e != cudaSuccess) return static_cast<int>(e);
    }
    e = cudaMemsetAsync(out->actions, 0, static_cast<size_t>(N) * P.A * 4, s);
    if (e != cudaSuccess) return static_cast<int>(e);
    RLSB_TRY(launch_onehot_to_idx(z0, M, cfg->groups, cfg->classes, out->stoch_idx, s));
  }

  // ---- the rollout: one launch ----------------------------------------------------------------------------------
  const int ring_bytes = (227 * 1024 - 1024 - static_cast<int>(sizeof(RCtl)) - 256) / 1024 * 1024;
  const size_t smem = static_cast<size_t>(ring_bytes) + sizeof(RCtl) + 1024;
  static PerDeviceOnce attr_once;
  unsigned long long dev_bit = 0;
  cudaError_t e;
  if (attr_once.need(dev_bit)) {
    e = cudaFuncSetAttribute(rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaFuncSetAttribute(rollout_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_once.done(dev_bit);
  }
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3(static_cast<unsigned>(m_pad / 128 * C));
  lc.blockDim = dim3(kThreads);
  lc.dynamicSmemBytes = smem;
  lc.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(C);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = g_pdl;
  lc.attrs = attr;
  lc.numAttrs = 2;
  e = cudaLaunchKernelEx(&lc, rollout_kernel, rp, ring_bytes);
  if (e != cudaSuccess) return static_cast<int>(e);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}


extern "C" int rlsb_rollout_bwd(const rlsb_imagine_cfg* cfg, const void* packed, int64_t N, const rlsb_imagine_out* fwd,
                                const float* g_rewards, const float* g_values, float* g_actions, void* workspace, void* stream_) {
  if (!cfg || !packed || !fwd || !g_rewards || !g_values || !g_actions || !workspace || N <= 0) return -1;
  if (!fwd->tape || !fwd->determ || !fwd->logits) return -2;
  if (N > (1LL << 24)) return -3;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  const int C = cluster_of(*cfg);
  RPlan R;
  RLSB_TRY(make_rplan(*cfg, C, R));
  if (!R.bwd) return -15;
  const k1::Plan& P = R.P;
  if ((P.D & 7) != 0) return -15;
  const int H = cfg->H;
  k1::Tape TP;
  k1::make_tape(P, N, H, TP);
  k1::BwdWorkspace W;
  k1::make_bwd_workspace(P, N, W);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  auto bf = [&](size_t off) { return reinterpret_cast<__nv_bfloat16*>(ws + off); };
  auto f32 = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };

  RolloutBwdParams rp{};
  for (int l = 0; l < 5; ++l) fill_layer(rp.t_head[l], R.t_head[l], pk);
  fill_layer(rp.t_prior2, R.t_prior2, pk);
  fill_layer(rp.t_prior1, R.t_prior1, pk);
  fill_layer(rp.t_gru, R.t_gru, pk);
  fill_layer(rp.t_img_in, R.t_img_in, pk);
  rp.C = C; rp.M = static_cast<int>(N); rp.m_pad = W.m_pad; rp.H = H;
  rp.D = P.D; rp.S = P.S; rp.A = P.A; rp.Dp = P.Dp; rp.Sp = P.Sp; rp.Ap = P.Ap; rp.Hp = P.Hp; rp.G3p = P.G3p;
  rp.Gb = P.Gb; rp.gb0 = P.gb0; rp.gb_reward = P.g_reward - P.gb0; rp.gb_critic = P.g_critic - P.gb0;
  rp.groups = cfg->groups; rp.layer_norm = cfg->layer_norm; rp.gru_nb = P.gru.NB;
  rp.eps = 1e-5f;
  rp.determ = fwd->determ; rp.logits = fwd->logits;
  rp.tape = static_cast<const uint8_t*>(fwd->tape); rp.tape_step = TP.step_bytes;
  for (int l = 0; l < 4; ++l) {
    rp.tp_head_pre[l] = TP.head_pre[l];
    rp.tp_head_rstd[l] = TP.head_rstd[l];
  }
  rp.tp_x_pre = TP.x_pre; rp.tp_x_rstd = TP.x_rstd; rp.tp_gru_scratch = TP.gru_scratch; rp.tp_gru_stats = TP.gru_stats;
  rp.tp_y_pre = TP.y_pre; rp.tp_y_rstd = TP.y_rstd;
  rp.tape_ld_scratch = TP.ld_scratch;
  // the gate backward's row code reads the GRU LayerNorm parameters in the reference's (3D) order (the forward slab's are
  // permuted per CTA): a plain copy kept in the blob
  rp.gru_gamma = reinterpret_cast<const float*>(pk + R.gru_ln_off);
  rp.gru_beta = rp.gru_gamma + 3 * P.D;
  rp.g_rewards = g_rewards; rp.g_values = g_values; rp.g_actions = g_actions;
  rp.dy4 = bf(W.dy4); rp.dh[0] = bf(W.dh[0]); rp.dh[1] = bf(W.dh[1]); rp.g_logits = bf(W.g_logits); rp.dp1 = bf(W.dp1);
  rp.g_pre = bf(W.g_pre); rp.dp_in = bf(W.dp_in);
  rp.g_s = f32(W.g_s); rp.g_hprior = f32(W.g_hprior); rp.g_hdirect = f32(W.g_hdirect); rp.g_hgru = f32(W.g_hgru);
  rp.g_za = f32(W.g_za);
  rp.ldS = W.ldS; rp.ldZA = W.ldZA;
  rp.kg_max = 4;
  if (const char* env = getenv("RLSB_ROLLOUT_KG")) rp.kg_max = atoi(env) >= 1 ? atoi(env) : 1;
  rp.trace = g_rollout_trace;

  const int ring_bytes = (227 * 1024 - 1024 - static_cast<int>(sizeof(RCtl)) - 256) / 1024 * 1024;
  const size_t smem = static_cast<size_t>(ring_bytes) + sizeof(RCtl) + 1024;
  static PerDeviceOnce attr_once;
  unsigned long long dev_bit = 0;
  cudaError_t e;
  if (attr_once.need(dev_bit)) {
    e = cudaFuncSetAttribute(rollout_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaFuncSetAttribute(rollout_bwd_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_once.done(dev_bit);
  }
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3(static_cast<unsigned>(W.m_pad / 128 * C));
  lc.blockDim = dim3(kThreads);
  lc.dynamicSmemBytes = smem;
  lc.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(C);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = g_pdl;
  lc.attrs = attr;
  lc.numAttrs = 2;
  e = cudaLaunchKernelEx(&lc, rollout_bwd_kernel, rp, ring_bytes);
  if (e != cudaSuccess) return static_cast<int>(e);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}
