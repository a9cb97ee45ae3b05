// rlsb_gemm.cu — persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   acc[128 x RB] (TMEM, fp32) = sum over K tiles of  A_tile[128 x 64] * W_tile[RB x 64]^T
//
// * operands live in HBM in the packed SWIZZLE_128B tile image (rlsb_ptx.cuh::packed_index), so
//   each pipeline stage is filled by two contiguous cp.async.bulk copies (TMA engine, UBLKCP)
//   that complete on an mbarrier; no tensor maps, no per-element address math;
// * one elected thread issues tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) into TMEM; smem
//   stages are released with tcgen05.commit; accumulators are double buffered in TMEM when
//   RB <= 256 so the epilogue of tile i overlaps the main loop of tile i+1;
// * four epilogue warps own one TMEM lane (= one output row) per thread, which makes the
//   LayerNorm statistics of rssm.py:136-152 / fc_nn.py:14-21 a per-thread reduction.
//
// Replaces (reference): nn.Linear + nn.LayerNorm + nn.ELU chains in
// agents/dreamer/rssm.py:136-152, agents/dreamer/common.py:58-75, utils/fc_nn.py:14-22.
#include <cstdlib>

#include "rlsb_gemm.cuh"
#include "rlsb_count.cuh"
#include "rlsb_ptx.cuh"

namespace rlsb {

namespace {

constexpr int kEpiWarps = 16;                       // 4 TMEM lane quarters x 4 column quarters
constexpr int kEpiThreads = kEpiWarps * 32;         // 512
constexpr int kMaxRB = 512;

struct SmemCtl {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
  alignas(16) float bias[2][kMaxRB];    // double-buffered per-tile epilogue parameters
  float gamma[2][kMaxRB];
  float beta[2][kMaxRB];
  float2 part[4][kTileM];   // per column-quarter partial (sum, sumsq) of each row
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int ACT>
__device__ __forceinline__ float act_apply(float x) {
  // ELU as the reference evaluates it on the CPU: exp(x) - 1 for x <= 0 (ATen elu kernel)
  if (ACT == ACT_ELU) return x > 0.f ? x : ex2_approx(x * 1.4426950408889634f) - 1.0f;
  if (ACT == ACT_RELU) return fmaxf(x, 0.f);
  return x;
}

// pass 2 of the full-row epilogue for one 8-column chunk: bias, [LayerNorm affine], activation,
// bf16 pack, one 16-byte store into the packed operand image.
template <int ACT, bool LN, bool SAVE>
__device__ __forceinline__ void ln_act_chunk(const uint32_t (&r)[8], int c, int n_valid, uint32_t s_bias,
                                             uint32_t s_gam, uint32_t s_bet, float rstd, float nmr,
                                             __nv_bfloat16* dst, __nv_bfloat16* dst_pre, bool row_ok) {
  float y[8];
  float pre[8];
  if (c >= n_valid || (SAVE && !row_ok)) {  // padding columns of the block / invalid row (training forward)
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = 0.f;
    if (SAVE) {
#pragma unroll
      for (int j = 0; j < 8; ++j) pre[j] = 0.f;
    }
  } else {
    const float4 b0 = lds128(s_bias + 4u * c), b1 = lds128(s_bias + 4u * c + 16u);
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    if (LN) {
      const float4 g0 = lds128(s_gam + 4u * c), g1 = lds128(s_gam + 4u * c + 16u);
      const float4 e0 = lds128(s_bet + 4u * c), e1 = lds128(s_bet + 4u * c + 16u);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = fmaf(__uint_as_float(r[j]) + bb[j], rstd, nmr);  // (x - mean) * rstd
        if (SAVE) pre[j] = xh;
        y[j] = act_apply<ACT>(fmaf(xh, gg[j], ee[j]));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = __uint_as_float(r[j]) + bb[j];
        if (SAVE) pre[j] = a;
        y[j] = act_apply<ACT>(a);
      }
    }
    if (c + 8 > n_valid) {  // partial chunk (N not a multiple of 8)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c + j >= n_valid) {
          y[j] = 0.f;
          if (SAVE) pre[j] = 0.f;
        }
    }
  }
  *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]),
                                              pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
  if (SAVE)
    *reinterpret_cast<uint4*>(dst_pre) = make_uint4(pack_bf16x2(pre[0], pre[1]), pack_bf16x2(pre[2], pre[3]),
                                                    pack_bf16x2(pre[4], pre[5]), pack_bf16x2(pre[6], pre[7]));
}

template <int ACT, bool LN, bool SAVE>
__device__ __forceinline__ void ln_act_pass2(uint32_t tmem_d, int cq, int my_chunks, int n_valid, int col0, int row,
                                             uint32_t s_bias, uint32_t s_gam, uint32_t s_bet, float rstd, float nmr,
                                             __nv_bfloat16* obase, __nv_bfloat16* pbase, bool row_ok) {
  for (int i0 = 0; i0 < my_chunks; i0 += 2) {
    uint32_t r0[8], r1[8];
    const int c0 = (cq + 4 * i0) * 8, c1 = c0 + 32;
    const bool two = i0 + 1 < my_chunks;
    tmem_ld8(tmem_d + static_cast<uint32_t>(c0), r0);
    if (two) tmem_ld8(tmem_d + static_cast<uint32_t>(c1), r1);
    tmem_ld_wait();
    {
      const int oc = col0 + c0;
      const size_t off = static_cast<size_t>(oc >> 6) * (kTileM * kTileK) + ((((oc & 63) >> 3) ^ (row & 7)) << 3);
      ln_act_chunk<ACT, LN, SAVE>(r0, c0, n_valid, s_bias, s_gam, s_bet, rstd, nmr, obase + off,
                                  SAVE ? pbase + off : nullptr, row_ok);
    }
    if (two) {
      const int oc = col0 + c1;
      const size_t off = static_cast<size_t>(oc >> 6) * (kTileM * kTileK) + ((((oc & 63) >> 3) ^ (row & 7)) << 3);
      ln_act_chunk<ACT, LN, SAVE>(r1, c1, n_valid, s_bias, s_gam, s_bet, rstd, nmr, obase + off,
                                  SAVE ? pbase + off : nullptr, row_ok);
    }
  }
}


__device__ __forceinline__ void epi_bar(int id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kEpiThreads) : "memory");
}

// work item -> (m_tile, g, nb).  The n-block index runs fastest so that the CTAs resident at any
// moment share a handful of A tiles (L2 hits) while the whole weight matrix stays L2 resident.
// With a cluster of `cs` CTAs, the CTAs of one cluster take `cs` consecutive M tiles of the SAME
// (g, nb): they consume the same weight block, which is fetched once per cluster (multicast).
__device__ __forceinline__ void decode_work(const GemmParams& p, int w, int cs, int rank, int& m_tile, int& gnb) {
  const int per_m = p.G * p.NB;
  const int m_super = w / per_m;
  gnb = w - m_super * per_m;
  m_tile = m_super * cs + rank;
}

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const GemmParams p, const int stages, const int nbuf, const int cs) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atom
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const uint32_t a_bytes = kTileM * kTileK * 2;                 // 16 KB
  const uint32_t b_bytes = static_cast<uint32_t>(p.RB) * kTileK * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(smem + static_cast<size_t>(stages) * stage_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  int kt_total = 0;
  for (int s = 0; s < p.n_seg; ++s) kt_total += p.a_ktiles[s];
  const int total_work = p.G * p.NB * ((p.m_tiles + cs - 1) / cs);   // (cluster-)work items
  const int buf_cols = (nbuf == 2) ? 256 : 512;
  const int rank = cs > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int cluster_id = static_cast<int>(blockIdx.x) / cs;
  const int num_clusters = static_cast<int>(gridDim.x) / cs;
  const uint16_t cta_mask = static_cast<uint16_t>((1u << cs) - 1u);

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < stages; ++s) {
        mbar_init(&ctl->full[s], 1);
        mbar_init(&ctl->empty[s], static_cast<uint32_t>(cs));   // released by every CTA of the cluster
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&ctl->tmem_full[b], 1);
        mbar_init(&ctl->tmem_empty[b], kEpiWarps);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(&ctl->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (cs > 1) cluster_sync_all();   // peers' barriers are initialised before anyone signals them
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ===================== producer: bulk copies into the stage ring =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t b_slice = b_bytes / static_cast<uint32_t>(cs);
      for (int w = cluster_id; w < total_work; w += num_clusters) {
        int m_tile, gnb;
        decode_work(p, w, cs, rank, m_tile, gnb);
        if (m_tile >= p.m_tiles) m_tile = p.m_tiles - 1;   // padding CTA of the last cluster: valid loads, no stores
        const int g = gnb / p.NB;
        const __nv_bfloat16* wsrc =
            p.W + static_cast<size_t>(gnb) * kt_total * (static_cast<size_t>(p.RB) * kTileK);
        int kt_glob = 0;
        for (int s = 0; s < p.n_seg; ++s) {
          const __nv_bfloat16* asrc = p.A[s] + static_cast<size_t>(g) * p.a_group_stride[s] +
                                      static_cast<size_t>(m_tile) * p.a_ktiles[s] *
                                          (kTileM * kTileK);
          for (int kt = 0; kt < p.a_ktiles[s]; ++kt, ++kt_glob) {
            mbar_wait(&ctl->empty[stage], phase ^ 1u);
            uint8_t* sa = smem + static_cast<size_t>(stage) * stage_bytes;
            uint8_t* sb = sa + a_bytes;
            mbar_expect_tx(&ctl->full[stage], stage_bytes);
            bulk_g2s(sa, asrc + static_cast<size_t>(kt) * (kTileM * kTileK), a_bytes,
                     &ctl->full[stage]);
            const __nv_bfloat16* wtile = wsrc + static_cast<size_t>(kt_glob) * (static_cast<size_t>(p.RB) * kTileK);
            if (cs == 1) {
              bulk_g2s(sb, wtile, b_bytes, &ctl->full[stage]);
            } else {
              // this CTA fetches rows [rank*RB/cs, (rank+1)*RB/cs) of the block for the whole cluster
              bulk_g2s_multicast(sb + static_cast<size_t>(rank) * b_slice,
                                 reinterpret_cast<const uint8_t*>(wtile) + static_cast<size_t>(rank) * b_slice,
                                 b_slice, &ctl->full[stage], cta_mask);
            }
            if (++stage == stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================================
    if (lane == 0) {
      const int n_chunk0 = p.RB > 256 ? 256 : p.RB;
      const int n_chunk1 = p.RB - n_chunk0;
      const uint32_t idesc0 = make_idesc_bf16(kTileM, static_cast<uint32_t>(n_chunk0));
      const uint32_t idesc1 = n_chunk1 > 0 ? make_idesc_bf16(kTileM, static_cast<uint32_t>(n_chunk1)) : 0u;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = cluster_id; w < total_work; w += num_clusters, ++it) {
        const int buf = (nbuf == 2) ? (it & 1) : 0;
        const uint32_t use = (nbuf == 2) ? static_cast<uint32_t>(it >> 1) : static_cast<uint32_t>(it);
        mbar_wait(&ctl->tmem_empty[buf], (use & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * buf_cols);
        for (int kt = 0; kt < kt_total; ++kt) {
          mbar_wait(&ctl->full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + static_cast<size_t>(stage) * stage_bytes);
          const uint32_t sb = sa + a_bytes;
          const uint64_t adesc = make_smem_desc_sw128(sa);
          const uint64_t bdesc0 = make_smem_desc_sw128(sb);
          const uint64_t bdesc1 = make_smem_desc_sw128(sb + 256u * 128u);
#pragma unroll
          for (int kk = 0; kk < kTileK / 16; ++kk) {
            const uint32_t acc = (kt > 0 || kk > 0) ? 1u : 0u;
            // +32 bytes per 16-element K step == +2 in the (addr >> 4) start-address field
            umma_bf16(tmem_d, adesc + static_cast<uint64_t>(kk * 2), bdesc0 + static_cast<uint64_t>(kk * 2),
                      idesc0, acc);
            if (n_chunk1 > 0)
              umma_bf16(tmem_d + 256u, adesc + static_cast<uint64_t>(kk * 2),
                        bdesc1 + static_cast<uint64_t>(kk * 2), idesc1, acc);
          }
          // frees the smem stage (in every CTA of the cluster: peers multicast into it) when these MMAs retire
          if (cs == 1) umma_commit(&ctl->empty[stage]);
          else umma_commit_multicast(&ctl->empty[stage], cta_mask);
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&ctl->tmem_full[buf]);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue: 16 warps ==============================================
    // warp -> (q, cq): q = TMEM lane quarter it may address (hardware: warp id % 4) = 32 rows,
    // cq = which quarter of the 8-column chunks it owns (chunk c belongs to cq = c % 4).
    // Row statistics are combined across the 4 column quarters through shared memory.
    const int q = warp & 3;
    const int cq = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int tid_e = threadIdx.x - 64;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int m_pad = p.m_tiles * kTileM;
    const int my_chunks = p.RB >> 5;  // (RB / 8) / 4
    const uint32_t s_part = smem_u32(&ctl->part[0][0]);
    int it = 0;
    for (int w = cluster_id; w < total_work; w += num_clusters, ++it) {
      int m_tile, gnb;
      decode_work(p, w, cs, rank, m_tile, gnb);
      const bool tile_ok = m_tile < p.m_tiles;   // false only for the padding CTA of the last cluster
      const int g = gnb / p.NB;
      const int nb = gnb - g * p.NB;
      const int buf = (nbuf == 2) ? (it & 1) : 0;
      const uint32_t use = (nbuf == 2) ? static_cast<uint32_t>(it >> 1) : static_cast<uint32_t>(it);
      const int col0 = nb * p.RB;                   // first column of this block inside the group
      const int n_valid = min(p.RB, p.N - col0);    // valid columns in this block (may be <= 0)
      const int pb = it & 1;
      const uint32_t s_bias = smem_u32(&ctl->bias[pb][0]);
      const uint32_t s_gam = smem_u32(&ctl->gamma[pb][0]);
      const uint32_t s_bet = smem_u32(&ctl->beta[pb][0]);
      const bool has_ln = (EPI == EPI_LN_ACT) && (p.ln_gamma != nullptr);
      // ---- stage this tile's parameters (overlaps the tile's main loop) ------------------------
      {
        const size_t poff = static_cast<size_t>(g) * p.NB * p.RB + col0;
        for (int i = tid_e; i < p.RB; i += kEpiThreads) {
          sts32(s_bias + 4u * i, p.bias ? __ldg(p.bias + poff + i) : 0.f);
          if (has_ln) {
            sts32(s_gam + 4u * i, __ldg(p.ln_gamma + static_cast<size_t>(g) * p.RB + i));
            sts32(s_bet + 4u * i, __ldg(p.ln_beta + static_cast<size_t>(g) * p.RB + i));
          }
        }
      }
      epi_bar(1);
      mbar_wait(&ctl->tmem_full[buf], use & 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * buf_cols) + lane_addr;
      const int m = tile_ok ? m_tile * kTileM + row : p.M + kTileM;   // padding tile: every row invalid

      // ---- pass 1: statistics (EPI_STATS / LayerNorm) and/or plain fp32 output ------------------
      float mean = 0.f, rstd = 1.f;
      if (EPI == EPI_PLAIN || EPI == EPI_STATS || has_ln) {
        float sum = 0.f, sq = 0.f;
        float* orow = (EPI == EPI_LN_ACT) ? nullptr
                                          : p.out_f32 + static_cast<size_t>(g) * p.out_group_stride +
                                                static_cast<size_t>(m) * p.ldo + col0;
        const bool vec_ok = (EPI != EPI_LN_ACT) && ((p.ldo & 3) == 0) && ((col0 & 3) == 0) &&
                            ((reinterpret_cast<uintptr_t>(p.out_f32) & 15) == 0);
        const bool store = (EPI != EPI_LN_ACT) && (m < p.M);
        auto pass1_chunk = [&](const uint32_t (&r)[8], int c) {
          const float4 b0 = lds128(s_bias + 4u * c), b1 = lds128(s_bias + 4u * c + 16u);
          float v[8];
          v[0] = __uint_as_float(r[0]) + b0.x; v[1] = __uint_as_float(r[1]) + b0.y;
          v[2] = __uint_as_float(r[2]) + b0.z; v[3] = __uint_as_float(r[3]) + b0.w;
          v[4] = __uint_as_float(r[4]) + b1.x; v[5] = __uint_as_float(r[5]) + b1.y;
          v[6] = __uint_as_float(r[6]) + b1.z; v[7] = __uint_as_float(r[7]) + b1.w;
          if (c + 8 <= n_valid) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              sum += v[j];
              sq = fmaf(v[j], v[j], sq);
            }
            if (store) {
              if (vec_ok) {
                *reinterpret_cast<float4*>(orow + c) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(orow + c + 4) = make_float4(v[4], v[5], v[6], v[7]);
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) orow[c + j] = v[j];
              }
            }
          } else if (c < n_valid) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (c + j < n_valid) {
                sum += v[j];
                sq = fmaf(v[j], v[j], sq);
                if (store) orow[c + j] = v[j];
              }
            }
          }
        };
        for (int i0 = 0; i0 < my_chunks; i0 += 2) {
          uint32_t r0[8], r1[8];
          const int c0 = (cq + 4 * i0) * 8, c1 = c0 + 32;
          const bool two = i0 + 1 < my_chunks;
          tmem_ld8(tmem_d + static_cast<uint32_t>(c0), r0);
          if (two) tmem_ld8(tmem_d + static_cast<uint32_t>(c1), r1);
          tmem_ld_wait();
          pass1_chunk(r0, c0);
          if (two) pass1_chunk(r1, c1);
        }
        if (EPI == EPI_STATS || has_ln) {
          sts64(s_part + 8u * (cq * kTileM + row), sum, sq);
          epi_bar(2);
          const float2 a0 = lds64(s_part + 8u * row), a1 = lds64(s_part + 8u * (kTileM + row)),
                       a2 = lds64(s_part + 8u * (2 * kTileM + row)), a3 = lds64(s_part + 8u * (3 * kTileM + row));
          const float tsum = (a0.x + a1.x) + (a2.x + a3.x);
          const float tsq = (a0.y + a1.y) + (a2.y + a3.y);
          if (EPI == EPI_STATS) {
            // per-(row, n-block) partials (sum, sum of squares); the consumer combines the blocks
            if (cq == 0 && tile_ok)
              reinterpret_cast<float2*>(p.stats)[static_cast<size_t>(gnb) * m_pad + m] = make_float2(tsum, tsq);
          } else {
            const float inv_n = 1.0f / static_cast<float>(n_valid);
            mean = tsum * inv_n;
            const float var = fmaxf(tsq * inv_n - mean * mean, 0.f);
            rstd = 1.0f / sqrtf(var + p.ln_eps);
          }
        }
      }

      // ---- pass 2 (EPI_LN_ACT): normalise, activate, write the packed bf16 operand image --------
      if (EPI == EPI_LN_ACT && tile_ok) {
        const float nmr = -mean * rstd;
        __nv_bfloat16* obase = p.out_bf16 + static_cast<size_t>(g) * p.out_bf16_group_stride +
                               static_cast<size_t>(m_tile) * (p.out_kpad >> 6) * (kTileM * kTileK) +
                               static_cast<size_t>(row) * kTileK;
        const int tot_chunks = p.RB >> 3;
#define RLSB_P2(ACT, LN) \
  ln_act_pass2<ACT, LN>(tmem_d, cq, my_chunks, n_valid, col0, row, s_bias, s_gam, s_bet, rstd, nmr, obase)
        if (has_ln) {
          if (p.act == ACT_ELU) RLSB_P2(ACT_ELU, true);
          else if (p.act == ACT_RELU) RLSB_P2(ACT_RELU, true);
          else RLSB_P2(ACT_NONE, true);
        } else {
          if (p.act == ACT_ELU) RLSB_P2(ACT_ELU, false);
          else if (p.act == ACT_RELU) RLSB_P2(ACT_RELU, false);
          else RLSB_P2(ACT_NONE, false);
        }
#undef RLSB_P2
        // zero the padding columns [RB, out_kpad) of the packed image (last block only)
        if (nb == p.NB - 1) {
          for (int ch = tot_chunks + cq; ch < ((p.out_kpad - col0) >> 3); ch += 4) {
            const int oc = col0 + ch * 8;
            __nv_bfloat16* trow = obase + static_cast<size_t>(oc >> 6) * (kTileM * kTileK);
            const int chunk = ((oc & 63) >> 3) ^ (row & 7);
            *reinterpret_cast<uint4*>(trow + chunk * 8) = make_uint4(0u, 0u, 0u, 0u);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->tmem_empty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();   // nobody leaves while a peer may still write its smem / barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int g_num_sms = 0;
int g_cluster_size = 2;   // CTAs per cluster sharing a weight block (1, 2 or 4); RLSB_CLUSTER overrides

}  // namespace

void set_gemm_cluster_size(int cs) {
  if (cs == 1 || cs == 2 || cs == 4) g_cluster_size = cs;
}
int gemm_cluster_size() { return g_cluster_size; }

int launch_gemm(const GemmParams& p, int epilogue, cudaStream_t stream) {
  if (p.RB <= 0 || p.RB > 512 || (p.RB % 32) != 0) return -1;
  if (p.n_seg < 1 || p.n_seg > kMaxSeg) return -2;
  if (epilogue == EPI_LN_ACT && p.NB != 1 && p.ln_gamma != nullptr) return -3;  // LayerNorm needs the whole row
  if (p.M <= 0 || p.m_tiles != (p.M + kTileM - 1) / kTileM) return -4;
  if (epilogue == EPI_LN_ACT && ((p.out_kpad % 64) != 0 || p.out_kpad < p.N)) return -5;
  if (g_num_sms == 0) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return static_cast<int>(e);
    if (const char* env = getenv("RLSB_CLUSTER")) set_gemm_cluster_size(atoi(env));
  }
  const int stage_bytes = kTileM * kTileK * 2 + p.RB * kTileK * 2;
  const int budget = 227 * 1024 - 1024 /*align*/ - static_cast<int>(sizeof(SmemCtl)) - 256;
  static_assert(sizeof(SmemCtl) < 20 * 1024, "control block grew");
  int stages = budget / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return -6;
  const int nbuf = p.RB <= 256 ? 2 : 1;
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + sizeof(SmemCtl) + 1024;
  // cluster size along M: CTAs of a cluster share the weight block (TMA multicast)
  int cs = g_cluster_size;
  while (cs > 1 && ((p.RB / cs) % 8 != 0 || p.m_tiles < cs)) cs >>= 1;
  const int total_work = p.G * p.NB * ((p.m_tiles + cs - 1) / cs);
  int clusters = g_num_sms / cs;
  if (total_work < clusters) clusters = total_work;
  const int grid = clusters * cs;

  cudaError_t e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(cs);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#define RLSB_LAUNCH(EPI)                                                                         \
  do {                                                                                           \
    static bool attr_done = false;                                                               \
    if (!attr_done) {                                                                            \
      e = cudaFuncSetAttribute(gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                               227 * 1024);                                                      \
      if (e != cudaSuccess) return static_cast<int>(e);                                          \
      attr_done = true;                                                                          \
    }                                                                                            \
    e = cudaLaunchKernelEx(&cfg, gemm_kernel<EPI>, p, stages, nbuf, cs);                         \
    if (e != cudaSuccess) return static_cast<int>(e);                                            \
  } while (0)
  switch (epilogue) {
    case EPI_PLAIN: RLSB_LAUNCH(EPI_PLAIN); break;
    case EPI_STATS: RLSB_LAUNCH(EPI_STATS); break;
    case EPI_LN_ACT: RLSB_LAUNCH(EPI_LN_ACT); break;
    default: return -7;
  }
#undef RLSB_LAUNCH
  count_launch();
  e = cudaGetLastError();
  return static_cast<int>(e);
}

}  // namespace rlsb
