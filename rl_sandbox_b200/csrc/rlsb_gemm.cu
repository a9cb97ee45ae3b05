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
// * sixteen epilogue warps (4 TMEM lane quarters x 4 column quarters) own one output row per thread, which makes the
//   LayerNorm statistics of rssm.py:136-152 / fc_nn.py:14-21 a per-thread reduction; where a row spans several n-blocks
//   (EPI_LN_ACT with NB > 1, EPI_GRU: the whole GRUCell of common.py:69-81) the blocks' CTAs exchange their row statistics
//   through tagged 64-bit words in global memory (GemmParams::xstats) — no second kernel, no fp32 round trip.
//
// Replaces (reference): nn.Linear + nn.LayerNorm + nn.ELU chains in
// agents/dreamer/rssm.py:136-152, agents/dreamer/common.py:58-75, utils/fc_nn.py:14-22.
#include <cstdlib>
#include <mutex>

#include "rlsb_gemm.cuh"
#include "rlsb_count.cuh"
#include "rlsb_ptx.cuh"
#include "rlsb_rowops.cuh"

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
  uint64_t peer_full[8];   // CTA pair: the odd CTA's stage is filled (signalled across the pair by its relay thread)
  uint32_t tmem_base;
  uint32_t pad;
  alignas(16) float bias[2][kMaxRB];    // double-buffered per-tile epilogue parameters
  float gamma[2][kMaxRB];
  float beta[2][kMaxRB];
  float2 part[4][kTileM];   // per column-quarter partial (sum, sumsq) of each row
  float2 part2[4][kTileM];  // cross-block LayerNorm: per gathering thread partial totals of each row
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
// activation on a pair.  ELU is evaluated as max(a, exp(min(a, 0)) - 1): identical to the reference's
// "a > 0 ? a : exp(a) - 1" (e^a - 1 >= a everywhere, and the exp argument never overflows), without predicates.
template <int ACT>
__device__ __forceinline__ void act_pair(f32x2 a2, float& y0, float& y1) {
  float a0, a1;
  unpk2(a2, a0, a1);
  if (ACT == ACT_ELU) {
    const f32x2 t = fmul2(pk2(fminf(a0, 0.f), fminf(a1, 0.f)), pk2(1.4426950408889634f, 1.4426950408889634f));
    float t0, t1;
    unpk2(t, t0, t1);
    const f32x2 em = fadd2(pk2(ex2_approx(t0), ex2_approx(t1)), pk2(-1.0f, -1.0f));
    float e0, e1;
    unpk2(em, e0, e1);
    y0 = fmaxf(a0, e0);
    y1 = fmaxf(a1, e1);
  } else if (ACT == ACT_RELU) {
    y0 = fmaxf(a0, 0.f);
    y1 = fmaxf(a1, 0.f);
  } else {
    y0 = a0;
    y1 = a1;
  }
}

// pass 2 of the full-row epilogue for one 8-column chunk: bias, [LayerNorm affine], activation,
// bf16 pack, one 16-byte store into the packed operand image (packed fp32x2 arithmetic).
template <int ACT, bool LN, bool SAVE>
__device__ __forceinline__ void ln_act_chunk_words(const uint32_t (&r)[8], int c, int n_valid, uint32_t s_bias,
                                                   uint32_t s_gam, uint32_t s_bet, f32x2 rstd2, f32x2 nmr2, bool row_ok,
                                                   uint4& o, uint4& q) {
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
    f32x2 b[4], v[4];
    lds_2x2(s_bias + 4u * c, b[0], b[1]);
    lds_2x2(s_bias + 4u * c + 16u, b[2], b[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = fadd2(pk2u(r[2 * i], r[2 * i + 1]), b[i]);
    if (LN) {
      f32x2 g[4], e[4];
      lds_2x2(s_gam + 4u * c, g[0], g[1]);
      lds_2x2(s_gam + 4u * c + 16u, g[2], g[3]);
      lds_2x2(s_bet + 4u * c, e[0], e[1]);
      lds_2x2(s_bet + 4u * c + 16u, e[2], e[3]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[i] = ffma2(v[i], rstd2, nmr2);  // (x - mean) * rstd
        if (SAVE) unpk2(v[i], pre[2 * i], pre[2 * i + 1]);
        act_pair<ACT>(ffma2(v[i], g[i], e[i]), y[2 * i], y[2 * i + 1]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (SAVE) unpk2(v[i], pre[2 * i], pre[2 * i + 1]);
        act_pair<ACT>(v[i], y[2 * i], y[2 * i + 1]);
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
  o = make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
  if (SAVE)
    q = make_uint4(pack_bf16x2(pre[0], pre[1]), pack_bf16x2(pre[2], pre[3]), pack_bf16x2(pre[4], pre[5]),
                   pack_bf16x2(pre[6], pre[7]));
}
template <int ACT, bool LN, bool SAVE>
__device__ __forceinline__ void ln_act_chunk(const uint32_t (&r)[8], int c, int n_valid, uint32_t s_bias,
                                             uint32_t s_gam, uint32_t s_bet, f32x2 rstd2, f32x2 nmr2,
                                             __nv_bfloat16* dst, __nv_bfloat16* dst_pre, bool row_ok) {
  uint4 o, q = make_uint4(0u, 0u, 0u, 0u);
  ln_act_chunk_words<ACT, LN, SAVE>(r, c, n_valid, s_bias, s_gam, s_bet, rstd2, nmr2, row_ok, o, q);
  *reinterpret_cast<uint4*>(dst) = o;
  if (SAVE && dst_pre) *reinterpret_cast<uint4*>(dst_pre) = q;
}

// the same for a chunk known to be complete (all 8 columns valid): no bounds checks, parameter addresses = base + immediate
template <int ACT, bool LN, bool SAVE>
__device__ __forceinline__ void ln_act_chunk_full_words(const uint32_t (&r)[8], uint32_t s_bias, uint32_t s_gam,
                                                        uint32_t s_bet, f32x2 rstd2, f32x2 nmr2, bool row_ok, uint4& o,
                                                        uint4& q) {
  float y[8];
  float pre[8];
  f32x2 b[4], v[4];
  lds_2x2(s_bias, b[0], b[1]);
  lds_2x2(s_bias + 16u, b[2], b[3]);
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = fadd2(pk2u(r[2 * i], r[2 * i + 1]), b[i]);
  if (LN) {
    f32x2 g[4], e[4];
    lds_2x2(s_gam, g[0], g[1]);
    lds_2x2(s_gam + 16u, g[2], g[3]);
    lds_2x2(s_bet, e[0], e[1]);
    lds_2x2(s_bet + 16u, e[2], e[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i] = ffma2(v[i], rstd2, nmr2);  // (x - mean) * rstd
      if (SAVE) unpk2(v[i], pre[2 * i], pre[2 * i + 1]);
      act_pair<ACT>(ffma2(v[i], g[i], e[i]), y[2 * i], y[2 * i + 1]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (SAVE) unpk2(v[i], pre[2 * i], pre[2 * i + 1]);
      act_pair<ACT>(v[i], y[2 * i], y[2 * i + 1]);
    }
  }
  o = make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
  if (SAVE) {   // invalid rows of the training forward stay zero (they are contraction indices of the weight gradient)
    q = make_uint4(pack_bf16x2(pre[0], pre[1]), pack_bf16x2(pre[2], pre[3]), pack_bf16x2(pre[4], pre[5]),
                   pack_bf16x2(pre[6], pre[7]));
    if (!row_ok) {
      o = make_uint4(0u, 0u, 0u, 0u);
      q = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}
template <int ACT, bool LN, bool SAVE>
__device__ __forceinline__ void ln_act_chunk_full(const uint32_t (&r)[8], uint32_t s_bias, uint32_t s_gam, uint32_t s_bet,
                                                  f32x2 rstd2, f32x2 nmr2, __nv_bfloat16* dst, __nv_bfloat16* dst_pre,
                                                  bool row_ok) {
  uint4 o, q = make_uint4(0u, 0u, 0u, 0u);
  ln_act_chunk_full_words<ACT, LN, SAVE>(r, s_bias, s_gam, s_bet, rstd2, nmr2, row_ok, o, q);
  if (SAVE && dst_pre) *reinterpret_cast<uint4*>(dst_pre) = q;
  *reinterpret_cast<uint4*>(dst) = o;
}

template <int ACT, bool LN, bool SAVE>
__device__ __forceinline__ void ln_act_pass2(uint32_t tmem_d, int cq, int my_chunks, int n_valid, int col0, int row,
                                             uint32_t s_bias, uint32_t s_gam, uint32_t s_bet, float rstd, float nmr,
                                             __nv_bfloat16* obase, __nv_bfloat16* pbase, bool row_ok, int n_fast) {
  const f32x2 rstd2 = pk2(rstd, rstd), nmr2 = pk2(nmr, nmr);
  int done = 0;
  if (n_fast >= 4) {
    // chunk i of this warp = columns (cq + 4 i) * 8 ..: k-tile (col0 >> 6) + (i >> 1), 16-byte position cq or cq + 4 inside
    // the tile's 128-byte row (XOR-swizzled by the row) — two fixed offsets and a tile pointer that moves every other chunk
    const int sw = row & 7;
    const int pos0 = (cq ^ sw) << 3, pos1 = ((cq + 4) ^ sw) << 3;
    __nv_bfloat16* o = obase + static_cast<size_t>(col0 >> 6) * (kTileM * kTileK);
    __nv_bfloat16* q = (SAVE && pbase) ? pbase + static_cast<size_t>(col0 >> 6) * (kTileM * kTileK) : nullptr;
    uint32_t sb = s_bias + 32u * cq, sg = s_gam + 32u * cq, se = s_bet + 32u * cq;
    done = tmem_sweep_groups(
        tmem_d + static_cast<uint32_t>(cq * 8), n_fast,
        [&](const uint32_t (&r)[8], auto J) {
          constexpr int j = decltype(J)::value;
          constexpr int tile_off = (j >> 1) * (kTileM * kTileK);
          const int pos = (j & 1) ? pos1 : pos0;
          ln_act_chunk_full<ACT, LN, SAVE>(r, sb + 128u * j, sg + 128u * j, se + 128u * j, rstd2, nmr2,
                                           o + tile_off + pos, (SAVE && q) ? q + tile_off + pos : nullptr, row_ok);
        },
        [&]() {
          o += 2 * (kTileM * kTileK);
          if (SAVE && q) q += 2 * (kTileM * kTileK);
          sb += 512u;
          sg += 512u;
          se += 512u;
        });
  }
  auto checked = [&](const uint32_t (&r)[8], int i) {   // bounds-checked (partial / padding chunks, odd shapes)
    const int c = (cq + 4 * i) * 8;
    const int oc = col0 + c;
    const size_t off = static_cast<size_t>(oc >> 6) * (kTileM * kTileK) + ((((oc & 63) >> 3) ^ (row & 7)) << 3);
    ln_act_chunk<ACT, LN, SAVE>(r, c, n_valid, s_bias, s_gam, s_bet, rstd2, nmr2, obase + off,
                                (SAVE && pbase) ? pbase + off : nullptr, row_ok);
  };
  if (done == 0) {
    tmem_sweep(tmem_d, cq, my_chunks, checked);
    return;
  }
  for (int i = done; i < my_chunks; ++i) {   // at most three chunks left
    uint32_t r[8];
    tmem_ld8(tmem_d + static_cast<uint32_t>((cq + 4 * i) * 8), r);
    tmem_ld_wait();
    checked(r, i);
  }
}


__device__ __forceinline__ void trace_stamp(unsigned long long* p, int slot) {
  if (p) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p[slot] = t;
  }
}
__device__ __forceinline__ void epi_bar(int id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kEpiThreads) : "memory");
}

// Cross-block LayerNorm (GemmParams::xstats): a block publishes a row's (sum, sum of squares) as two 64-bit words
// {value, tag} — 64-bit relaxed gpu-scope stores are single-copy atomic, so a reader that finds the expected tag in a word
// has its value: no fence, no counter, no atomics.  The tag of a launch is the previous tag of the slot + 1: every launch
// writes every slot of a row exactly once, so all NB slots of a row carry the same tag between launches (zero them once).
__device__ __forceinline__ void xst_store(unsigned long long* slot, float sum, float sq, uint32_t tag) {
  const unsigned long long hi = static_cast<unsigned long long>(tag) << 32;
  asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(slot), "l"(hi | __float_as_uint(sum)) : "memory");
  asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(slot + 1), "l"(hi | __float_as_uint(sq)) : "memory");
}
__device__ __forceinline__ uint32_t xst_tag(const unsigned long long* slot) {
  unsigned long long w;
  asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(w) : "l"(slot) : "memory");
  return static_cast<uint32_t>(w >> 32);
}
// sum over the blocks b = b_first, b_first + 4, ... < nb_total of row m: every block's words are polled until they carry `tag`
// (normally the first read; a partner that never publishes means the CTAs of the launch are not co-resident: trap after
// 4 M polls — a launch error, not a hang).  Up to four blocks are in flight at once.
__device__ __forceinline__ void xst_gather(const unsigned long long* base, size_t block_stride, int b_first, int nb_total,
                                           uint32_t tag, float& s, float& q) {
  s = 0.f;
  q = 0.f;
  for (int b0 = b_first; b0 < nb_total; b0 += 16) {
    unsigned long long w0[4], w1[4];
    bool ok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ok[i] = !(b0 + 4 * i < nb_total);
      w0[i] = w1[i] = 0ull;
    }
    for (unsigned int polls = 0;; ++polls) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (!ok[i]) {
          const unsigned long long* src = base + static_cast<size_t>(b0 + 4 * i) * block_stride;
          asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(w0[i]), "=l"(w1[i]) : "l"(src) : "memory");
        }
      }
      bool all = true;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (!ok[i]) ok[i] = static_cast<uint32_t>(w0[i] >> 32) == tag && static_cast<uint32_t>(w1[i] >> 32) == tag;
        all = all && ok[i];
      }
      if (all) break;
      if (polls > (1u << 22)) __trap();   // ~ seconds of polling (counted in polls, not in wall time: a pre-empted context does not trip it)
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (b0 + 4 * i < nb_total) {
        s += __uint_as_float(static_cast<uint32_t>(w0[i]));
        q += __uint_as_float(static_cast<uint32_t>(w1[i]));
      }
    }
  }
}

// Pass 2 of the full-row epilogue with the output assembled in shared memory.  A thread owns one row, so its 16-byte
// chunks land in 32 different 128-byte lines per warp store — the LSU then bounds the whole kernel (a bias-only epilogue
// took 77 us where the main loop needs 24).  A 128 x 64 tile of the packed image is 16 KB of contiguous global memory in
// exactly the layout it has here, so the 16 epilogue warps write their two chunks of k-tile kt into a 16 KB slot with
// conflict-free STS.128 and one thread sends the slot with a single bulk copy (two slots: the copy of k-tile kt - 1 drains
// while kt is assembled).  ph / pending carry the slot parity and "a copy is in flight" across tiles.
template <int ACT, bool LN, bool SAVE>
__device__ __forceinline__ void ln_act_pass2_staged(uint32_t tmem_d, int cq, int n_valid, int kt_out, int row, int tid_e,
                                                    uint32_t s_bias, uint32_t s_gam, uint32_t s_bet, float rstd, float nmr,
                                                    bool row_ok, uint8_t* stage_out, uint8_t* stage_pre,
                                                    __nv_bfloat16* gout, __nv_bfloat16* gpre, uint32_t& ph, bool& pending) {
  const f32x2 rstd2 = pk2(rstd, rstd), nmr2 = pk2(nmr, nmr);
  const uint32_t sw = static_cast<uint32_t>(row & 7);
  const uint32_t so0 = smem_u32(stage_out) + static_cast<uint32_t>(row) * 128u;
  const uint32_t sp0 = smem_u32(stage_pre) + static_cast<uint32_t>(row) * 128u;
  const bool save_here = SAVE && gpre != nullptr;
  for (int kt = 0; kt < kt_out; ++kt) {
    const uint32_t slot = ph & 1u;
    uint32_t r0[8] = {}, r1[8] = {};
    const int ch0 = kt * 8 + cq, ch1 = ch0 + 4;
    const int c0 = ch0 * 8, c1 = ch1 * 8;
    if (c0 < n_valid) tmem_ld8(tmem_d + static_cast<uint32_t>(c0), r0);
    if (c1 < n_valid) tmem_ld8(tmem_d + static_cast<uint32_t>(c1), r1);
    tmem_ld_wait2(r0, r1);
    uint4 o, q = make_uint4(0u, 0u, 0u, 0u);
    if (c0 + 8 <= n_valid)
      ln_act_chunk_full_words<ACT, LN, SAVE>(r0, s_bias + 4u * c0, s_gam + 4u * c0, s_bet + 4u * c0, rstd2, nmr2, row_ok, o, q);
    else
      ln_act_chunk_words<ACT, LN, SAVE>(r0, c0, n_valid, s_bias, s_gam, s_bet, rstd2, nmr2, row_ok, o, q);
    sts128(so0 + slot * 16384u + ((static_cast<uint32_t>(cq) ^ sw) << 4), o);
    if (save_here) sts128(sp0 + slot * 16384u + ((static_cast<uint32_t>(cq) ^ sw) << 4), q);
    if (c1 + 8 <= n_valid)
      ln_act_chunk_full_words<ACT, LN, SAVE>(r1, s_bias + 4u * c1, s_gam + 4u * c1, s_bet + 4u * c1, rstd2, nmr2, row_ok, o, q);
    else
      ln_act_chunk_words<ACT, LN, SAVE>(r1, c1, n_valid, s_bias, s_gam, s_bet, rstd2, nmr2, row_ok, o, q);
    sts128(so0 + slot * 16384u + ((static_cast<uint32_t>(cq + 4) ^ sw) << 4), o);
    if (save_here) sts128(sp0 + slot * 16384u + ((static_cast<uint32_t>(cq + 4) ^ sw) << 4), q);
    // the copy of the previous k-tile has finished reading the other slot before anybody passes the barrier and
    // starts to overwrite it
    if (tid_e == 0 && pending) bulk_wait_read<0>();
    fence_proxy_async_smem();   // this thread's slot writes are visible to the bulk-copy engine
    epi_bar(4);
    if (tid_e == 0) {
      bulk_s2g(gout + static_cast<size_t>(kt) * (kTileM * kTileK), stage_out + slot * 16384u, 16384u);
      if (save_here) bulk_s2g(gpre + static_cast<size_t>(kt) * (kTileM * kTileK), stage_pre + slot * 16384u, 16384u);
      bulk_commit();
    }
    pending = true;
    ph ^= 1u;
  }
}

// work item -> (m_tile, g, nb).  The n-block index runs fastest so that the CTAs resident at any
// moment share a handful of A tiles (L2 hits) while the whole weight matrix stays L2 resident.
// With a cluster of `cs` CTAs, the CTAs of one cluster take `cs` consecutive M tiles of the SAME
// (g, nb): they consume the same weight block, which is fetched once per cluster (multicast).
__device__ __forceinline__ void decode_work(const GemmParams& p, int w, int cs, int rank, int& m_tile, int& gnb) {
  if (p.group_major) {   // all M tiles of (g, nb) 0, then 1, ...: a CTA meets each group once (EPI_BWD column sums)
    const int m_supers = (p.m_tiles + cs - 1) / cs;
    gnb = w / m_supers;
    m_tile = (w - gnb * m_supers) * cs + rank;
    return;
  }
  const int per_m = p.G * p.NB;
  const int m_super = w / per_m;
  gnb = w - m_super * per_m;
  m_tile = m_super * cs + rank;
}

__device__ __forceinline__ bool row_is_valid(const GemmParams& p, int m) {
  return p.row_period > 0 ? (m < p.M && (m % p.row_period) < p.row_valid) : (m < p.M);
}

// column sums over the 32 rows a warp holds, for the 8 columns of one chunk: 9 shuffles instead of 40.
// On return the lanes with (lane & 3) == 0 hold the total of column
//   ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1).
__device__ __forceinline__ float colsum8(const float (&v)[8], int lane) {
  float w4[4], w2[2];
  const bool h16 = (lane & 16) != 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float mine = h16 ? v[4 + j] : v[j];
    const float other = h16 ? v[j] : v[4 + j];
    w4[j] = mine + __shfl_xor_sync(0xffffffffu, other, 16);
  }
  const bool h8 = (lane & 8) != 0;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float mine = h8 ? w4[2 + j] : w4[j];
    const float other = h8 ? w4[j] : w4[2 + j];
    w2[j] = mine + __shfl_xor_sync(0xffffffffu, other, 8);
  }
  const bool h4 = (lane & 4) != 0;
  float s = (h4 ? w2[1] : w2[0]) + __shfl_xor_sync(0xffffffffu, h4 ? w2[0] : w2[1], 4);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  return s;
}

__device__ __forceinline__ void red_shared_add(uint32_t a, float x) {
  asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory");
}

__device__ __forceinline__ void unpack_bf16x8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

template <int ACT>
__device__ __forceinline__ float act_grad(float a) {
  if (ACT == ACT_ELU) return a > 0.f ? 1.0f : ex2_approx(a * 1.4426950408889634f);
  if (ACT == ACT_RELU) return a > 0.f ? 1.0f : 0.f;
  return 1.0f;
}

// EPI_BWD, one 8-column chunk: g = dL/dy (TMEM) and the saved pre (x_hat or pre-activation) ->
// da = g * act'(a), dxh = da * gamma.  Returns da / dxh in place; pre is unpacked into `pre`.
template <int ACT, bool LN>
__device__ __forceinline__ void bwd_chunk(const uint32_t (&r)[8], const uint4& pre_bits, int c, int n_valid,
                                          uint32_t s_gam, uint32_t s_bet, bool row_ok, float (&da)[8],
                                          float (&dxh)[8], float (&pre)[8]) {
  unpack_bf16x8(pre_bits, pre);
  if (LN) {
    const float4 g0 = lds128(s_gam + 4u * c), g1 = lds128(s_gam + 4u * c + 16u);
    const float4 e0 = lds128(s_bet + 4u * c), e1 = lds128(s_bet + 4u * c + 16u);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool ok = row_ok && (c + j < n_valid);
      const float a = fmaf(pre[j], gg[j], ee[j]);
      da[j] = ok ? __uint_as_float(r[j]) * act_grad<ACT>(a) : 0.f;
      dxh[j] = da[j] * gg[j];
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool ok = row_ok && (c + j < n_valid);
      da[j] = ok ? __uint_as_float(r[j]) * act_grad<ACT>(pre[j]) : 0.f;
      dxh[j] = da[j];
    }
  }
}

// bwd_chunk for a complete chunk of a valid row: no predicates, packed fp32x2 arithmetic, parameters at base + immediate
template <int ACT, bool LN>
__device__ __forceinline__ void bwd_chunk_full(const uint32_t (&r)[8], const uint4& pre_bits, uint32_t s_gam, uint32_t s_bet,
                                               f32x2 (&da)[4], f32x2 (&dxh)[4], f32x2 (&pre)[4]) {
  {
    float f[8];
    unpack_bf16x8(pre_bits, f);
#pragma unroll
    for (int i = 0; i < 4; ++i) pre[i] = pk2(f[2 * i], f[2 * i + 1]);
  }
  f32x2 g[4];
  if (LN) {
    f32x2 e[4];
    lds_2x2(s_gam, g[0], g[1]);
    lds_2x2(s_gam + 16u, g[2], g[3]);
    lds_2x2(s_bet, e[0], e[1]);
    lds_2x2(s_bet + 16u, e[2], e[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) e[i] = ffma2(pre[i], g[i], e[i]);   // a = x_hat * gamma + beta
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f32x2 d = pk2(1.0f, 1.0f);
      if (ACT == ACT_ELU) {   // act'(a) = a > 0 ? 1 : exp(a) = exp(min(a, 0))
        float a0, a1;
        unpk2(e[i], a0, a1);
        const f32x2 t = fmul2(pk2(fminf(a0, 0.f), fminf(a1, 0.f)), pk2(1.4426950408889634f, 1.4426950408889634f));
        float t0, t1;
        unpk2(t, t0, t1);
        d = pk2(ex2_approx(t0), ex2_approx(t1));
      } else if (ACT == ACT_RELU) {
        float a0, a1;
        unpk2(e[i], a0, a1);
        d = pk2(a0 > 0.f ? 1.0f : 0.f, a1 > 0.f ? 1.0f : 0.f);
      }
      da[i] = fmul2(pk2u(r[2 * i], r[2 * i + 1]), d);
      dxh[i] = fmul2(da[i], g[i]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f32x2 d = pk2(1.0f, 1.0f);
      float a0, a1;
      unpk2(pre[i], a0, a1);
      if (ACT == ACT_ELU) {
        const f32x2 t = fmul2(pk2(fminf(a0, 0.f), fminf(a1, 0.f)), pk2(1.4426950408889634f, 1.4426950408889634f));
        float t0, t1;
        unpk2(t, t0, t1);
        d = pk2(ex2_approx(t0), ex2_approx(t1));
      } else if (ACT == ACT_RELU) {
        d = pk2(a0 > 0.f ? 1.0f : 0.f, a1 > 0.f ? 1.0f : 0.f);
      }
      da[i] = fmul2(pk2u(r[2 * i], r[2 * i + 1]), d);
      dxh[i] = da[i];
    }
  }
}

// the whole EPI_BWD epilogue of one tile (see GemmParams); returns nothing, writes out_bf16 / smem column sums
template <int ACT, bool LN>
__device__ __forceinline__ void bwd_tile(const GemmParams& p, uint32_t tmem_d, int cq, int my_chunks, int n_valid,
                                         int col0, int row, int lane, bool row_ok, uint32_t s_gam, uint32_t s_bet,
                                         uint32_t s_part, uint32_t s_acc, float rstd,
                                         const __nv_bfloat16* pbase, __nv_bfloat16* obase, int n_fast, uint32_t xs) {
  // xs != 0: this thread's chunk i of the saved image was prefetched to shared memory at xs + i * 8192 (see the kernel)
  float s1 = 0.f, s2 = 0.f;
  const bool colsum = LN && p.col_part != nullptr;
  // ---- complete chunks of a warp whose 32 rows are all valid: software-pipelined sweep (TMEM and the saved x_hat of a
  // chunk are fetched two chunks ahead), no predicates, per-chunk addresses = group base + immediate ------------------
  int done = 0;
  const int sw = row & 7;
  const int pos0 = (cq ^ sw) << 3, pos1 = ((cq + 4) ^ sw) << 3;   // see ln_act_pass2
  const size_t tile0 = static_cast<size_t>(col0 >> 6) * (kTileM * kTileK);
  if (n_fast >= 4 && __all_sync(0xffffffffu, row_ok)) {
    const __nv_bfloat16* q = pbase + tile0;
    __nv_bfloat16* o = obase + tile0;
    uint32_t sg = s_gam + 32u * cq, se = s_bet + 32u * cq, sa = s_acc + 32u * cq;
    uint32_t xg = xs;
    f32x2 s1a = pk2(0.f, 0.f), s1b = s1a, s2a = s1a, s2b = s1a;
    done = tmem_sweep_groups_g(
        tmem_d + static_cast<uint32_t>(cq * 8), n_fast,
        [&](int j) {
          return xs ? lds128u(xg + 8192u * static_cast<uint32_t>(j))
                    : *reinterpret_cast<const uint4*>(q + (j >> 1) * (kTileM * kTileK) + ((j & 1) ? pos1 : pos0));
        },
        [&](const uint32_t (&r)[8], const uint4& pb, auto J, uint32_t t) {
          constexpr int j = decltype(J)::value;
          f32x2 da[4], dxh[4], pre[4];
          bwd_chunk_full<ACT, LN>(r, pb, sg + 128u * j, se + 128u * j, da, dxh, pre);
          if (LN) {
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float x0, x1;
              unpk2(dxh[i], x0, x1);
              w[2 * i] = __float_as_uint(x0);
              w[2 * i + 1] = __float_as_uint(x1);
            }
            tmem_st8(t + 32u * j, w);
            s1a = fadd2(s1a, dxh[0]);
            s1b = fadd2(s1b, dxh[1]);
            s2a = ffma2(dxh[0], pre[0], s2a);
            s2b = ffma2(dxh[1], pre[1], s2b);
            s1a = fadd2(s1a, dxh[2]);
            s1b = fadd2(s1b, dxh[3]);
            s2a = ffma2(dxh[2], pre[2], s2a);
            s2b = ffma2(dxh[3], pre[3], s2b);
            if (colsum) {
              float dg[8], db[8];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                unpk2(fmul2(da[i], pre[i]), dg[2 * i], dg[2 * i + 1]);
                unpk2(da[i], db[2 * i], db[2 * i + 1]);
              }
              const float cg = colsum8(dg, lane);
              const float cb = colsum8(db, lane);
              if ((lane & 3) == 0) {
                const uint32_t col4 = 4u * (((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1));
                red_shared_add(sa + 128u * j + col4, cg);
                red_shared_add(sa + 128u * j + col4 + 4u * p.RB, cb);
              }
            }
          } else {
            float y[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) unpk2(dxh[i], y[2 * i], y[2 * i + 1]);
            *reinterpret_cast<uint4*>(o + (j >> 1) * (kTileM * kTileK) + ((j & 1) ? pos1 : pos0)) =
                make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
          }
        },
        [&]() {
          q += 2 * (kTileM * kTileK);
          o += 2 * (kTileM * kTileK);
          sg += 512u;
          se += 512u;
          sa += 512u;
          xg += 4u * 8192u;
        });
    if (LN) {
      float a, b;
      unpk2(fadd2(s1a, s1b), a, b);
      s1 = a + b;
      unpk2(fadd2(s2a, s2b), a, b);
      s2 = a + b;
    }
  }
  for (int i = done; i < my_chunks; ++i) {
    const int c = (cq + 4 * i) * 8;
    if (c >= n_valid) break;   // warp-uniform
    uint32_t r[8];
    tmem_ld8(tmem_d + static_cast<uint32_t>(c), r);
    const int oc = col0 + c;   // column inside the whole row (col0 != 0 only without LayerNorm)
    const size_t off = static_cast<size_t>(oc >> 6) * (kTileM * kTileK) + ((((oc & 63) >> 3) ^ (row & 7)) << 3);
    const uint4 pb = xs ? lds128u(xs + 8192u * static_cast<uint32_t>(i)) : *reinterpret_cast<const uint4*>(pbase + off);
    tmem_ld_wait();
    float da[8], dxh[8], pre[8];
    bwd_chunk<ACT, LN>(r, pb, c, n_valid, s_gam, s_bet, row_ok, da, dxh, pre);
    if (LN) {
      {   // park dxh in the accumulator's columns: pass 2 then needs neither gamma / beta nor the activation derivative
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = __float_as_uint(dxh[j]);
        tmem_st8(tmem_d + static_cast<uint32_t>(c), w);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1 += dxh[j];
        s2 = fmaf(dxh[j], pre[j], s2);
      }
      if (colsum) {
        float dg[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) dg[j] = da[j] * pre[j];
        const float cg = colsum8(dg, lane);
        const float cb = colsum8(da, lane);
        if ((lane & 3) == 0) {
          const int col = c + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
          red_shared_add(s_acc + 4u * col, cg);
          red_shared_add(s_acc + 4u * (p.RB + col), cb);
        }
      }
    } else {
      *reinterpret_cast<uint4*>(obase + off) = make_uint4(pack_bf16x2(dxh[0], dxh[1]), pack_bf16x2(dxh[2], dxh[3]),
                                                          pack_bf16x2(dxh[4], dxh[5]), pack_bf16x2(dxh[6], dxh[7]));
    }
  }
  if (LN) {
    tmem_st_wait();
    sts64(s_part + 8u * (cq * kTileM + row), s1, s2);
    epi_bar(2);
    const float2 a0 = lds64(s_part + 8u * row), a1 = lds64(s_part + 8u * (kTileM + row)),
                 a2 = lds64(s_part + 8u * (2 * kTileM + row)), a3 = lds64(s_part + 8u * (3 * kTileM + row));
    const float inv_n = n_valid == p.N ? p.inv_n : 1.0f / static_cast<float>(n_valid);
    const float m1 = ((a0.x + a1.x) + (a2.x + a3.x)) * inv_n;
    const float m2 = ((a0.y + a1.y) + (a2.y + a3.y)) * inv_n;
    const float nm1r = -m1 * rstd, nm2r = -m2 * rstd;
    if (done > 0) {   // rstd * (dxh - m1 - x_hat * m2) over the complete chunks, pipelined as pass 1
      const __nv_bfloat16* q = pbase + tile0;
      __nv_bfloat16* o = obase + tile0;
      const f32x2 rstd2 = pk2(rstd, rstd), nm1r2 = pk2(nm1r, nm1r), nm2r2 = pk2(nm2r, nm2r);
      uint32_t xg = xs;
      tmem_sweep_groups_g(
          tmem_d + static_cast<uint32_t>(cq * 8), done,
          [&](int j) {
            return xs ? lds128u(xg + 8192u * static_cast<uint32_t>(j))
                      : *reinterpret_cast<const uint4*>(q + (j >> 1) * (kTileM * kTileK) + ((j & 1) ? pos1 : pos0));
          },
          [&](const uint32_t (&r)[8], const uint4& pb, auto J, uint32_t) {
            constexpr int j = decltype(J)::value;
            float f[8], y[8];
            unpack_bf16x8(pb, f);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              unpk2(ffma2(pk2(f[2 * i], f[2 * i + 1]), nm2r2, ffma2(pk2u(r[2 * i], r[2 * i + 1]), rstd2, nm1r2)), y[2 * i],
                    y[2 * i + 1]);
            *reinterpret_cast<uint4*>(o + (j >> 1) * (kTileM * kTileK) + ((j & 1) ? pos1 : pos0)) =
                make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
          },
          [&]() {
            q += 2 * (kTileM * kTileK);
            o += 2 * (kTileM * kTileK);
            xg += 4u * 8192u;
          });
    }
    for (int i = done; i < my_chunks; ++i) {
      const int c = (cq + 4 * i) * 8;
      if (c >= n_valid) break;
      uint32_t r[8];
      tmem_ld8(tmem_d + static_cast<uint32_t>(c), r);   // dxh (0 in invalid rows / columns)
      const int oc = col0 + c;
      const size_t off = static_cast<size_t>(oc >> 6) * (kTileM * kTileK) + ((((oc & 63) >> 3) ^ (row & 7)) << 3);
      const uint4 pb = xs ? lds128u(xs + 8192u * static_cast<uint32_t>(i)) : *reinterpret_cast<const uint4*>(pbase + off);
      tmem_ld_wait();
      float pre[8], o[8];
      unpack_bf16x8(pb, pre);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool ok = row_ok && (c + j < n_valid);
        // rstd * (dxh - m1 - pre * m2)
        o[j] = ok ? fmaf(pre[j], nm2r, fmaf(__uint_as_float(r[j]), rstd, nm1r)) : 0.f;
      }
      *reinterpret_cast<uint4*>(obase + off) = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                                          pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
    }
  }
}

template <int EPI, bool PAIR>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const GemmParams p, const int stages, const int nbuf, const int cs) {
  constexpr bool pair = PAIR;   // compile-time: a kernel that contains cta_group::2 instructions can only be launched
                                // with an even cluster size
  // pair != 0 (cs == 2, RB <= 256): the two CTAs of a cluster run ONE tcgen05.mma.cta_group::2 per k-step over their two
  // M tiles: each stages its own A tile and HALF of the weight block (rows [rank RB/2, (rank+1) RB/2)), so an SM receives
  // 16 + RB/4 KB per k-step instead of 16 + RB/8 ... 16 + RB/4 x 2 with multicast — the L2 -> SM delivery is what bounds
  // the wide contractions.  Rank 0 issues the MMAs and owns the barriers the issue depends on; rank 1's MMA warp relays
  // "my stage is full" across the pair.
  constexpr bool kLnAct = (EPI == EPI_LN_ACT || EPI == EPI_LN_ACT_SAVE);
  pdl_launch_dependents();   // the next kernel of the stream may be scheduled now (it waits before reading memory)
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atom
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const uint32_t a_bytes = kTileM * kTileK * 2;                 // 16 KB
  const uint32_t b_bytes = static_cast<uint32_t>(p.RB) * kTileK * 2 / (pair ? 2u : 1u);   // bytes staged per CTA
  const uint32_t stage_bytes = a_bytes + b_bytes;
  // staged output (ln_act_pass2_staged): two 16 KB slots for the output image (+ two for x_hat in the saving variant)
  uint8_t* stage_out = smem + static_cast<size_t>(stages) * stage_bytes;
  uint8_t* stage_pre = stage_out + 32768;
  const uint32_t staging_bytes = (p.staged_out || EPI == EPI_GRU) ? (EPI == EPI_LN_ACT_SAVE ? 65536u : 32768u) : 0u;
  uint8_t* xbuf = stage_out + staging_bytes;   // EPI_BWD: per-thread slots of the saved image (GemmParams::xbuf_bytes)
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(xbuf + (EPI == EPI_BWD ? static_cast<uint32_t>(p.xbuf_bytes) : 0u));

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
        // multicast: released by the MMA warp of every CTA of the cluster; pair: by the one issuing CTA
        mbar_init(&ctl->empty[s], pair ? 1u : static_cast<uint32_t>(cs));
        mbar_init(&ctl->peer_full[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&ctl->tmem_full[b], 1);
        mbar_init(&ctl->tmem_empty[b], pair ? 2 * kEpiWarps : kEpiWarps);   // pair: both CTAs' epilogues drain
      }
      fence_mbar_init();
    }
    __syncwarp();
    if constexpr (pair) {
      tmem_alloc2(&ctl->tmem_base, 512);
      tmem_relinquish2();
    } else {
      tmem_alloc(&ctl->tmem_base, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (cs > 1) cluster_sync_all();   // peers' barriers are initialised before anyone signals them
  const uint32_t tmem_base = ctl->tmem_base;
  pdl_wait();   // everything above (barriers, TMEM) overlapped the predecessor's tail; its results are visible from here

  if (warp == 0) {
    // ===================== producer: bulk copies into the stage ring =====================
    // (whole warp in uniform control flow, one elected lane issues: see the MMA warp)
    {
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
          const __nv_bfloat16* abase = (s == 0 && p.alt_A != nullptr && g + 1 == p.alt_group_p1)
                                           ? p.alt_A : p.A[s] + static_cast<size_t>(g) * p.a_group_stride[s];
          const __nv_bfloat16* asrc = abase + static_cast<size_t>(m_tile) * p.a_ktiles[s] * (kTileM * kTileK);
          for (int kt = 0; kt < p.a_ktiles[s]; ++kt, ++kt_glob) {
            mbar_wait(&ctl->empty[stage], phase ^ 1u);
            uint8_t* sa = smem + static_cast<size_t>(stage) * stage_bytes;
            uint8_t* sb = sa + a_bytes;
            const __nv_bfloat16* wtile = wsrc + static_cast<size_t>(kt_glob) * (static_cast<size_t>(p.RB) * kTileK);
            if (elect_one()) {
            mbar_expect_tx(&ctl->full[stage], stage_bytes);
            bulk_g2s(sa, asrc + static_cast<size_t>(kt) * (kTileM * kTileK), a_bytes,
                     &ctl->full[stage]);
            if (pair) {
              // this CTA's half of the rows of each MMA's column range (<= 256 columns per instruction), at the same
              // offsets in both CTAs: rows [rank n0/2, +n0/2) and, for blocks wider than 256, [n0 + rank n1/2, +n1/2)
              const uint32_t n0 = p.RB > 256 ? 256u : static_cast<uint32_t>(p.RB), n1 = static_cast<uint32_t>(p.RB) - n0;
              const uint32_t h0 = n0 / 2u * 128u, h1 = n1 / 2u * 128u;   // bytes
              const uint8_t* wb = reinterpret_cast<const uint8_t*>(wtile);
              bulk_g2s(sb, wb + static_cast<size_t>(rank) * h0, h0, &ctl->full[stage]);
              if (n1 > 0) bulk_g2s(sb + h0, wb + static_cast<size_t>(n0) * 128u + static_cast<size_t>(rank) * h1, h1,
                                   &ctl->full[stage]);
            } else if (cs == 1) {
              bulk_g2s(sb, wtile, b_bytes, &ctl->full[stage]);
            } else {
              // this CTA fetches rows [rank*RB/cs, (rank+1)*RB/cs) of the block for the whole cluster
              bulk_g2s_multicast(sb + static_cast<size_t>(rank) * b_slice,
                                 reinterpret_cast<const uint8_t*>(wtile) + static_cast<size_t>(rank) * b_slice,
                                 b_slice, &ctl->full[stage], cta_mask);
            }
            }
            __syncwarp();
            if (++stage == stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================================================
    // The WHOLE warp walks the issue loop in uniform control flow and one elected lane executes the tcgen05 instructions:
    // descriptors, TMEM addresses and predicates then live in uniform registers.  With the loop inside `if (lane == 0)`
    // the compiler cannot prove them warp-uniform and wraps every UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY
    // "waterfall" (5 R2UR per MMA) — ~2x the 128-cycle floor of a 128 x 256 x 16 MMA per issue.
    if (pair && rank == 1) {
      // odd CTA of a pair: no MMAs to issue; tell the even CTA when each of MY stages has landed
      int stage = 0;
      uint32_t phase = 0;
      for (int w = cluster_id; w < total_work; w += num_clusters) {
        for (int kt = 0; kt < kt_total; ++kt) {
          mbar_wait(&ctl->full[stage], phase);
          if (elect_one()) mbar_arrive_cluster(mapa_cluster(smem_u32(&ctl->peer_full[stage]), 0u));
          __syncwarp();
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    } else {
      const int n_chunk0 = p.RB > 256 ? 256 : p.RB;
      const int n_chunk1 = p.RB - n_chunk0;
      const uint32_t mma_m = pair ? 2 * kTileM : kTileM;
      const uint32_t idesc0 = make_idesc_bf16(mma_m, static_cast<uint32_t>(n_chunk0));
      const uint32_t idesc1 = n_chunk1 > 0 ? make_idesc_bf16(mma_m, static_cast<uint32_t>(n_chunk1)) : 0u;
      // rows of the first column range staged per CTA (pair: half of them)
      const uint32_t b1_off = static_cast<uint32_t>(n_chunk0) / (pair ? 2u : 1u) * 128u;
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
          if (pair) mbar_wait(&ctl->peer_full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + static_cast<size_t>(stage) * stage_bytes);
          const uint32_t sb = sa + a_bytes;
          const uint64_t adesc = make_smem_desc_sw128(sa);
          const uint64_t bdesc0 = make_smem_desc_sw128(sb);
          const uint64_t bdesc1 = make_smem_desc_sw128(sb + b1_off);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < kTileK / 16; ++kk) {
              const uint32_t acc = (kt > 0 || kk > 0) ? 1u : 0u;
              // +32 bytes per 16-element K step == +2 in the (addr >> 4) start-address field
              if constexpr (pair) {
                umma_bf16_2cta(tmem_d, adesc + static_cast<uint64_t>(kk * 2), bdesc0 + static_cast<uint64_t>(kk * 2),
                               idesc0, acc);
                if (n_chunk1 > 0)
                  umma_bf16_2cta(tmem_d + 256u, adesc + static_cast<uint64_t>(kk * 2),
                                 bdesc1 + static_cast<uint64_t>(kk * 2), idesc1, acc);
              } else {
                umma_bf16(tmem_d, adesc + static_cast<uint64_t>(kk * 2), bdesc0 + static_cast<uint64_t>(kk * 2),
                          idesc0, acc);
                if (n_chunk1 > 0)
                  umma_bf16(tmem_d + 256u, adesc + static_cast<uint64_t>(kk * 2),
                            bdesc1 + static_cast<uint64_t>(kk * 2), idesc1, acc);
              }
            }
            // frees the smem stage (in every CTA of the cluster: peers multicast into it) when these MMAs retire
            if constexpr (pair) umma_commit2_multicast(&ctl->empty[stage], cta_mask);
            else if (cs == 1) umma_commit(&ctl->empty[stage]);
            else umma_commit_multicast(&ctl->empty[stage], cta_mask);
          }
          __syncwarp();
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        // accumulator complete -> epilogue (pair: of both CTAs)
        if (elect_one()) {
          if constexpr (pair) umma_commit2_multicast(&ctl->tmem_full[buf], cta_mask);
          else umma_commit(&ctl->tmem_full[buf]);
        }
        __syncwarp();
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
    const uint32_t s_part2 = smem_u32(&ctl->part2[0][0]);
    // EPI_BWD: per-CTA column sums (d_gamma | d_beta) of the current group live in the (unused) bias buffers
    const uint32_t s_acc = smem_u32(&ctl->bias[0][0]);
    const bool bwd_colsum = (EPI == EPI_BWD) && p.col_part != nullptr && p.ln_gamma != nullptr;
    int g_acc = -1;
    auto flush_colsum = [&](int g_old) {
      epi_bar(3);
      if (g_old >= 0) {
        float* dst = p.col_part + (static_cast<size_t>(blockIdx.x) * p.G + g_old) * 2 * p.RB;
        for (int i = tid_e; i < 2 * p.RB; i += kEpiThreads) dst[i] = lds32(s_acc + 4u * i);
      }
      for (int i = tid_e; i < 2 * p.RB; i += kEpiThreads) sts32(s_acc + 4u * i, 0.f);
      epi_bar(3);
    };
    // EPI_BWD: the saved x_hat / pre-activation chunks this thread needs (both passes) arrive in shared memory by cp.async,
    // issued one tile ahead — their latency (52 % of the epilogue warps' stall samples were long-scoreboard waits on these
    // 16-byte loads, profiles/r02_d_bwd_*) hides behind the previous tile's pass 2 and this tile's main loop
    const uint32_t xs = (EPI == EPI_BWD && p.xbuf_bytes > 0) ? smem_u32(xbuf) + static_cast<uint32_t>(tid_e) * 16u : 0u;
    auto prefetch_saved = [&](int wn) {
      if (xs == 0u || wn >= total_work) return;
      int mt, gn;
      decode_work(p, wn, cs, rank, mt, gn);
      if (mt >= p.m_tiles) return;
      const int gg = gn / p.NB;
      const int nv = min(p.RB, p.N);
      const __nv_bfloat16* src = p.bwd_pre + static_cast<size_t>(gg) * p.out_bf16_group_stride +
                                 static_cast<size_t>(mt) * (p.out_kpad >> 6) * (kTileM * kTileK) + static_cast<size_t>(row) * kTileK;
      for (int i = 0; i < my_chunks; ++i) {
        const int c = (cq + 4 * i) * 8;
        if (c >= nv) break;
        cp_async16(xs + 8192u * static_cast<uint32_t>(i),
                   src + static_cast<size_t>(c >> 6) * (kTileM * kTileK) + ((((c & 63) >> 3) ^ (row & 7)) << 3));
      }
      cp_async_commit();
    };
    if (EPI == EPI_BWD) prefetch_saved(cluster_id);
    int it = 0;
    uint32_t out_ph = 0;        // staged output: slot parity, carried across tiles
    bool out_pending = false;   //                a bulk store of this CTA is in flight
    for (int w = cluster_id; w < total_work; w += num_clusters, ++it) {
      int m_tile, gnb;
      decode_work(p, w, cs, rank, m_tile, gnb);
      const bool tile_ok = m_tile < p.m_tiles;   // false only for the padding CTA of the last cluster
      if (bwd_colsum && gnb / p.NB != g_acc) {
        flush_colsum(g_acc);
        g_acc = gnb / p.NB;
      }
      const int g = gnb / p.NB;
      const int nb = gnb - g * p.NB;
      const int buf = (nbuf == 2) ? (it & 1) : 0;
      const uint32_t use = (nbuf == 2) ? static_cast<uint32_t>(it >> 1) : static_cast<uint32_t>(it);
      const int col0 = nb * p.RB;                   // first column of this block inside the group
      const int n_valid = min(p.RB, p.N - col0);    // valid columns in this block (may be <= 0)
      // chunks of this warp that are complete and need no bounds checks (full-row LayerNorm epilogues)
      const int n_fast = ((kLnAct || EPI == EPI_BWD) && (n_valid & 7) == 0 && (col0 & 63) == 0 && (n_valid >> 3) > cq)
                             ? min(my_chunks, ((n_valid >> 3) - cq + 3) >> 2) : 0;
      const int pb = it & 1;
      const uint32_t s_bias = smem_u32(&ctl->bias[pb][0]);
      const uint32_t s_gam = smem_u32(&ctl->gamma[pb][0]);
      const uint32_t s_bet = smem_u32(&ctl->beta[pb][0]);
      const bool has_ln = (kLnAct || EPI == EPI_BWD || EPI == EPI_GRU) && (p.ln_gamma != nullptr);
      // LayerNorm over a row that spans the NB n-blocks of the group: statistics meet in global memory (GemmParams::xstats)
      const bool xln = (kLnAct || EPI == EPI_GRU) && has_ln && p.NB > 1;
      // ---- stage this tile's parameters (overlaps the tile's main loop) ------------------------
      {
        const size_t poff = static_cast<size_t>(g) * p.NB * p.RB + col0;
        const size_t lnoff = xln ? poff : static_cast<size_t>(g) * p.RB;
        for (int i = tid_e; i < p.RB; i += kEpiThreads) {
          if (EPI != EPI_BWD) sts32(s_bias + 4u * i, p.bias ? __ldg(p.bias + poff + i) : 0.f);
          if (has_ln) {
            sts32(s_gam + 4u * i, __ldg(p.ln_gamma + lnoff + i));
            sts32(s_bet + 4u * i, __ldg(p.ln_beta + lnoff + i));
          }
        }
      }
      unsigned long long* trp = (EPI == EPI_GRU && p.trace && tid_e == 0 && it < 64)
                                    ? p.trace + (static_cast<size_t>(blockIdx.x) * 64 + it) * 8 : nullptr;
      epi_bar(1);
      trace_stamp(trp, 0);
      mbar_wait(&ctl->tmem_full[buf], use & 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * buf_cols) + lane_addr;
      const int m = tile_ok ? m_tile * kTileM + row : p.M + kTileM;   // padding tile: every row invalid
      const bool row_ok = tile_ok && row_is_valid(p, m);
      // this block's row totals (identical in the row's four threads) -> totals over all NB blocks of the row
      unsigned long long* xrow = xln && tile_ok ? p.xstats + 2 * static_cast<size_t>(m) : nullptr;   // + 2 m_pad per block
      const uint32_t xtag = xrow ? xst_tag(xrow + 2 * static_cast<size_t>(nb) * m_pad) + 1u : 0u;
      auto xln_exchange = [&](float tsum, float tsq, float& s_all, float& q_all) {
        if (xrow && cq == 0) xst_store(xrow + 2 * static_cast<size_t>(nb) * m_pad, tsum, tsq, xtag);
        trace_stamp(trp, 3);
        // the row's four threads (one per column quarter) gather a quarter of the blocks each and meet in shared memory:
        // summed in a fixed order, the same totals in every CTA of the row block
        float ps = 0.f, pq = 0.f;
        if (xrow) xst_gather(xrow, 2 * static_cast<size_t>(m_pad), cq, p.NB, xtag, ps, pq);
        trace_stamp(trp, 4);
        sts64(s_part2 + 8u * (cq * kTileM + row), ps, pq);
        epi_bar(2);
        const float2 t0 = lds64(s_part2 + 8u * row), t1 = lds64(s_part2 + 8u * (kTileM + row)),
                     t2 = lds64(s_part2 + 8u * (2 * kTileM + row)), t3 = lds64(s_part2 + 8u * (3 * kTileM + row));
        s_all = (t0.x + t1.x) + (t2.x + t3.x);
        q_all = (t0.y + t1.y) + (t2.y + t3.y);
      };

      if constexpr (EPI == EPI_GRU) {
        trace_stamp(trp, 1);
        // thread (row, cq) owns the chunks cq + 4 i, i < 6, of the block's 192 columns: i = 0, 1 reset | 2, 3 candidate |
        // 4, 5 update pre-activations of the SAME hidden units 64 nb + 8 (cq + 4 j), j = i & 1.  The accumulator is copied
        // to registers once (48 values) and its TMEM buffer handed back at once, so that waiting for the partner blocks'
        // statistics never stalls the MMA warp.
        float v[6][8];
        float sum = 0.f, sq = 0.f;
        {
          uint32_t r[6][8];
#pragma unroll
          for (int i = 0; i < 6; ++i) tmem_ld8(tmem_d + static_cast<uint32_t>((cq + 4 * i) * 8), r[i]);
          tmem_ld_wait2(r[0], r[1]);
          tmem_ld_wait2(r[2], r[3]);
          tmem_ld_wait2(r[4], r[5]);
          f32x2 sum2 = pk2(0.f, 0.f), sq2 = pk2(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const uint32_t sb = s_bias + 32u * static_cast<uint32_t>(cq + 4 * i);
            f32x2 b[4];
            lds_2x2(sb, b[0], b[1]);
            lds_2x2(sb + 16u, b[2], b[3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const f32x2 x = fadd2(pk2u(r[i][2 * e], r[i][2 * e + 1]), b[e]);
              sum2 = fadd2(sum2, x);
              sq2 = ffma2(x, x, sq2);
              unpk2(x, v[i][2 * e], v[i][2 * e + 1]);
            }
          }
          float s0, s1, q0, q1;
          unpk2(sum2, s0, s1);
          unpk2(sq2, q0, q1);
          sum = s0 + s1;
          sq = q0 + q1;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (pair && rank == 1) mbar_arrive_cluster(mapa_cluster(smem_u32(&ctl->tmem_empty[buf]), 0u));
          else mbar_arrive(&ctl->tmem_empty[buf]);
        }
        trace_stamp(trp, 2);
        // h of the previous step for this thread's 2 x 8 hidden units (in flight across the exchange)
        const int u0 = nb * 64;
        float4 hp[2][2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          hp[j][0] = hp[j][1] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_ok) {
            const float4* src = reinterpret_cast<const float4*>(p.gru_h_prev + static_cast<size_t>(m) * p.gru_ld_h + u0 + (cq + 4 * j) * 8);
            hp[j][0] = __ldg(src);
            hp[j][1] = __ldg(src + 1);
          }
        }
        sts64(s_part + 8u * (cq * kTileM + row), sum, sq);
        epi_bar(2);
        float tsum, tsq;
        {
          const float2 a0 = lds64(s_part + 8u * row), a1 = lds64(s_part + 8u * (kTileM + row)),
                       a2 = lds64(s_part + 8u * (2 * kTileM + row)), a3 = lds64(s_part + 8u * (3 * kTileM + row));
          tsum = (a0.x + a1.x) + (a2.x + a3.x);
          tsq = (a0.y + a1.y) + (a2.y + a3.y);
        }
        float s_all, q_all;
        xln_exchange(tsum, tsq, s_all, q_all);
        trace_stamp(trp, 5);
        const float mean = s_all * p.inv_n;
        const float rstd = 1.0f / sqrtf(fmaxf(q_all * p.inv_n - mean * mean, 0.f) + p.ln_eps);
        const float nmr = -mean * rstd;
        const uint32_t slot = out_ph & 1u;
        const uint32_t so = smem_u32(stage_out) + slot * 16384u + static_cast<uint32_t>(row) * 128u;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint32_t cr = 32u * static_cast<uint32_t>(cq + 4 * j);   // byte offset of the reset chunk's parameters
          float y[8];
          const float h[8] = {hp[j][0].x, hp[j][0].y, hp[j][0].z, hp[j][0].w, hp[j][1].x, hp[j][1].y, hp[j][1].z, hp[j][1].w};
#pragma unroll
          for (int e4 = 0; e4 < 2; ++e4) {
            const float4 gr = lds128(s_gam + cr + 16u * e4), gc = lds128(s_gam + cr + 256u + 16u * e4),
                         gu = lds128(s_gam + cr + 512u + 16u * e4);
            const float4 br = lds128(s_bet + cr + 16u * e4), bc = lds128(s_bet + cr + 256u + 16u * e4),
                         bu = lds128(s_bet + cr + 512u + 16u * e4);
            const float wgr[4] = {gr.x, gr.y, gr.z, gr.w}, wgc[4] = {gc.x, gc.y, gc.z, gc.w}, wgu[4] = {gu.x, gu.y, gu.z, gu.w};
            const float wbr[4] = {br.x, br.y, br.z, br.w}, wbc[4] = {bc.x, bc.y, bc.z, bc.w}, wbu[4] = {bu.x, bu.y, bu.z, bu.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int k = 4 * e4 + e;
              const float rg = rowops::fast_sigmoid(fmaf(fmaf(v[j][k], rstd, nmr), wgr[e], wbr[e]));
              const float cand = rowops::fast_tanh(rg * fmaf(fmaf(v[2 + j][k], rstd, nmr), wgc[e], wbc[e]));
              const float ug = rowops::fast_sigmoid(fmaf(fmaf(v[4 + j][k], rstd, nmr), wgu[e], wbu[e]) + p.gru_update_bias);
              y[k] = row_ok ? fmaf(ug, cand - h[k], h[k]) : 0.f;
            }
          }
          if (row_ok) {
            float4* dst = reinterpret_cast<float4*>(p.gru_h_next + static_cast<size_t>(m) * p.gru_ld_hn + u0 + (cq + 4 * j) * 8);
            dst[0] = make_float4(y[0], y[1], y[2], y[3]);
            dst[1] = make_float4(y[4], y[5], y[6], y[7]);
          }
          sts128(so + ((static_cast<uint32_t>(cq + 4 * j) ^ static_cast<uint32_t>(row & 7)) << 4),
                 make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7])));
        }
        trace_stamp(trp, 6);
        // the packed bf16 image of h': k-tile nb of the row block, one 16 KB bulk store (two slots alternate)
        if (tid_e == 0 && out_pending) bulk_wait_read<0>();
        fence_proxy_async_smem();
        epi_bar(4);
        if (tid_e == 0 && tile_ok) {
          bulk_s2g(p.out_bf16 + (static_cast<size_t>(m_tile) * (p.out_kpad >> 6) + nb) * (kTileM * kTileK),
                   stage_out + slot * 16384u, 16384u);
          bulk_commit();
          out_pending = true;
        }
        out_ph ^= 1u;
        trace_stamp(trp, 7);
        continue;
      }

      if (EPI == EPI_BWD) {
        if (tile_ok) {
          const size_t img = static_cast<size_t>(g) * p.out_bf16_group_stride +
                             static_cast<size_t>(m_tile) * (p.out_kpad >> 6) * (kTileM * kTileK) +
                             static_cast<size_t>(row) * kTileK;
          const float rs = (has_ln && row_ok) ? __ldg(p.bwd_rstd + static_cast<size_t>(g) * m_pad + m) : 0.f;
          if (xs) cp_async_wait_all();
#define RLSB_B(ACT, LN) \
  bwd_tile<ACT, LN>(p, tmem_d, cq, my_chunks, n_valid, col0, row, lane, row_ok, s_gam, s_bet, s_part, s_acc, rs, \
                    p.bwd_pre + img, p.out_bf16 + img, n_fast, xs)
          if (has_ln) {
            if (p.act == ACT_ELU) RLSB_B(ACT_ELU, true);
            else RLSB_B(ACT_NONE, true);
          } else {
            if (p.act == ACT_ELU) RLSB_B(ACT_ELU, false);
            else if (p.act == ACT_RELU) RLSB_B(ACT_RELU, false);
            else RLSB_B(ACT_NONE, false);
          }
#undef RLSB_B
          // zero the padding columns [N rounded down to a chunk .. out_kpad) of the packed image (last block)
          if (nb == p.NB - 1) {
            const int n_end = col0 + n_valid;
            for (int ch = (n_end >> 3) + cq; ch < (p.out_kpad >> 3); ch += 4) {
              if (ch * 8 < n_end) continue;   // partial chunk was written above
              __nv_bfloat16* trow = p.out_bf16 + img + static_cast<size_t>(ch >> 3) * (kTileM * kTileK);
              *reinterpret_cast<uint4*>(trow + (((ch & 7) ^ (row & 7)) << 3)) = make_uint4(0u, 0u, 0u, 0u);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (pair && rank == 1) mbar_arrive_cluster(mapa_cluster(smem_u32(&ctl->tmem_empty[buf]), 0u));
          else mbar_arrive(&ctl->tmem_empty[buf]);
        }
        prefetch_saved(w + num_clusters);
        continue;
      }

      // ---- pass 1: statistics (EPI_STATS / LayerNorm) and/or plain fp32 output ------------------
      float mean = 0.f, rstd = 1.f;
      if (EPI == EPI_PLAIN || EPI == EPI_STATS || has_ln) {
        float sum = 0.f, sq = 0.f;
        float* orow = (kLnAct) ? nullptr
                                          : p.out_f32 + static_cast<size_t>(g) * p.out_group_stride +
                                                static_cast<size_t>(m) * p.ldo + col0;
        const bool vec_ok = (!kLnAct) && ((p.ldo & 3) == 0) && ((col0 & 3) == 0) &&
                            ((reinterpret_cast<uintptr_t>(p.out_f32) & 15) == 0);
        const bool store = (!kLnAct) && row_ok;
        f32x2 sum2 = pk2(0.f, 0.f), sq2 = pk2(0.f, 0.f);
        auto pass1_chunk = [&](const uint32_t (&r)[8], int c) {
          f32x2 b[4], v2[4];
          lds_2x2(s_bias + 4u * c, b[0], b[1]);
          lds_2x2(s_bias + 4u * c + 16u, b[2], b[3]);
#pragma unroll
          for (int i = 0; i < 4; ++i) v2[i] = fadd2(pk2u(r[2 * i], r[2 * i + 1]), b[i]);
          if (c + 8 <= n_valid) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              sum2 = fadd2(sum2, v2[i]);
              sq2 = ffma2(v2[i], v2[i], sq2);
            }
            if (store) {
              float v[8];
#pragma unroll
              for (int i = 0; i < 4; ++i) unpk2(v2[i], v[2 * i], v[2 * i + 1]);
              if (vec_ok) {
                *reinterpret_cast<float4*>(orow + c) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(orow + c + 4) = make_float4(v[4], v[5], v[6], v[7]);
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) orow[c + j] = v[j];
              }
            }
          } else if (c < n_valid) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) unpk2(v2[i], v[2 * i], v[2 * i + 1]);
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
        int done1 = 0;
        if ((EPI == EPI_PLAIN || EPI == EPI_STATS) && p.staged_out) {
          // fp32 row-major output through shared memory: a thread owns a row, so its 32 bytes per chunk would go to a
          // different 128-byte line per lane (32 partial lines per warp store).  The sixteen warps instead write one chunk
          // each of a 128 x 32 slab (16 KB, rows of 128 bytes with the 16-byte slots XOR-swizzled by the row), and after a
          // barrier every warp stores eight rows of the slab as full 128-byte lines.  Two slabs alternate: the barrier of
          // slab s + 1 also says that everybody has finished reading slab s - 1's buffer.
          const uint32_t sbuf0 = smem_u32(stage_out);
          const int wq = warp - 2;                       // 0..15: rows [8 wq, 8 wq + 8) of a slab on the way out
          const uint32_t swz = static_cast<uint32_t>(row & 7);
          float* gbase = p.out_f32 + static_cast<size_t>(g) * p.out_group_stride + col0;
          uint32_t ra[8] = {}, rb[8] = {};
          const int nsl = my_chunks;                     // chunk i of this warp lies in slab i (columns [32 i, 32 i + 32))
          tmem_ld8(tmem_d + static_cast<uint32_t>(cq * 8), ra);
          for (int i = 0; i < nsl; ++i) {
            const int c = (cq + 4 * i) * 8;
            if (i & 1) tmem_ld_wait2(rb, ra); else tmem_ld_wait2(ra, rb);
            if (i + 1 < nsl) {
              if (i & 1) tmem_ld8(tmem_d + static_cast<uint32_t>(c + 32), ra);
              else tmem_ld8(tmem_d + static_cast<uint32_t>(c + 32), rb);
            }
            f32x2 b[4], v2[4];
            lds_2x2(s_bias + 4u * c, b[0], b[1]);
            lds_2x2(s_bias + 4u * c + 16u, b[2], b[3]);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              v2[e] = fadd2((i & 1) ? pk2u(rb[2 * e], rb[2 * e + 1]) : pk2u(ra[2 * e], ra[2 * e + 1]), b[e]);
            if (c + 8 <= n_valid) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                sum2 = fadd2(sum2, v2[e]);
                sq2 = ffma2(v2[e], v2[e], sq2);
              }
            }
            const uint32_t sbuf = sbuf0 + static_cast<uint32_t>(i & 1) * 16384u;
            {
              float f[8];
#pragma unroll
              for (int e = 0; e < 4; ++e) unpk2(v2[e], f[2 * e], f[2 * e + 1]);
              const uint32_t rowa = sbuf + static_cast<uint32_t>(row) * 128u;
              sts128(rowa + ((static_cast<uint32_t>(2 * cq) ^ swz) << 4),
                     make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])));
              sts128(rowa + ((static_cast<uint32_t>(2 * cq + 1) ^ swz) << 4),
                     make_uint4(__float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7])));
            }
            epi_bar(4);
            if (tile_ok && col0 + 32 * i < p.N) {   // (N is a multiple of 32 in this mode: a slab is valid or padding)
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int rr = wq * 8 + h * 4 + (lane >> 3);
                const uint32_t slot = static_cast<uint32_t>(lane & 7);
                const float4 v = lds128(sbuf + static_cast<uint32_t>(rr) * 128u + ((slot ^ static_cast<uint32_t>(rr & 7)) << 4));
                const int mm = m_tile * kTileM + rr;
                if (row_is_valid(p, mm))
                  *reinterpret_cast<float4*>(gbase + static_cast<size_t>(mm) * p.ldo + 32 * i + 4 * static_cast<int>(slot)) = v;
              }
            }
          }
          done1 = my_chunks;
        }
        if (kLnAct && n_fast >= 4) {   // complete chunks: statistics only, two independent accumulator pairs
          f32x2 sumb = pk2(0.f, 0.f), sqb = pk2(0.f, 0.f);
          uint32_t sb = s_bias + 32u * cq;
          done1 = tmem_sweep_groups(
              tmem_d + static_cast<uint32_t>(cq * 8), n_fast,
              [&](const uint32_t (&r)[8], auto J) {
                constexpr int j = decltype(J)::value;
                f32x2 b[4], v2[4];
                lds_2x2(sb + 128u * j, b[0], b[1]);
                lds_2x2(sb + 128u * j + 16u, b[2], b[3]);
#pragma unroll
                for (int i = 0; i < 4; ++i) v2[i] = fadd2(pk2u(r[2 * i], r[2 * i + 1]), b[i]);
                sum2 = fadd2(sum2, v2[0]);
                sumb = fadd2(sumb, v2[1]);
                sq2 = ffma2(v2[0], v2[0], sq2);
                sqb = ffma2(v2[1], v2[1], sqb);
                sum2 = fadd2(sum2, v2[2]);
                sumb = fadd2(sumb, v2[3]);
                sq2 = ffma2(v2[2], v2[2], sq2);
                sqb = ffma2(v2[3], v2[3], sqb);
              },
              [&]() { sb += 512u; });
          sum2 = fadd2(sum2, sumb);
          sq2 = fadd2(sq2, sqb);
        }
        if (done1 >= my_chunks) {
          // everything went through the staged sweep
        } else if (done1 == 0) {
          tmem_sweep(tmem_d, cq, my_chunks, [&](const uint32_t (&r)[8], int i) { pass1_chunk(r, (cq + 4 * i) * 8); });
        } else {
          for (int i = done1; i < my_chunks; ++i) {
            uint32_t r[8];
            tmem_ld8(tmem_d + static_cast<uint32_t>((cq + 4 * i) * 8), r);
            tmem_ld_wait();
            pass1_chunk(r, (cq + 4 * i) * 8);
          }
        }
        {
          float s0, s1, q0, q1;
          unpk2(sum2, s0, s1);
          unpk2(sq2, q0, q1);
          sum += s0 + s1;
          sq += q0 + q1;
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
            float ts = tsum, tq = tsq;
            if (xln) xln_exchange(tsum, tsq, ts, tq);
            const float inv_n = (xln || n_valid == p.N) ? p.inv_n : 1.0f / static_cast<float>(n_valid);
            mean = ts * inv_n;
            const float var = fmaxf(tq * inv_n - mean * mean, 0.f);
            rstd = 1.0f / sqrtf(var + p.ln_eps);
            if (p.save_rstd && cq == 0 && tile_ok) {
              if (p.alt_group_p1 == 0) p.save_rstd[static_cast<size_t>(g) * m_pad + m] = rstd;
              else if (g + 1 == p.alt_group_p1) p.save_rstd[m] = rstd;
            }
          }
        }
      }

      // ---- pass 2 (EPI_LN_ACT): normalise, activate, write the packed bf16 operand image --------
      if (kLnAct && tile_ok) {
        const float nmr = -mean * rstd;
        const bool alt = (g + 1 == p.alt_group_p1);
        const size_t tile_off = static_cast<size_t>(m_tile) * (p.out_kpad >> 6) * (kTileM * kTileK) +
                                static_cast<size_t>(row) * kTileK;
        const size_t goff = static_cast<size_t>(g) * p.out_bf16_group_stride;
        __nv_bfloat16* obase = (alt ? p.alt_out_bf16 : p.out_bf16 + goff) + tile_off;
        const int tot_chunks = p.RB >> 3;
        // saves: every group (training forward of rlsb_ac_update) or, with an alt group, that group alone
        __nv_bfloat16* pbase = !p.save_pre ? nullptr
                               : (p.alt_group_p1 == 0 ? p.save_pre + goff + tile_off
                                                      : (alt ? p.save_pre + tile_off : nullptr));
        // staged: tile pointers without the row term (the slot already is the 128-row tile)
        // (NB > 1: this block's RB / 64 tiles of the image)
        __nv_bfloat16* gout = obase - static_cast<size_t>(row) * kTileK + static_cast<size_t>(col0 >> 6) * (kTileM * kTileK);
        __nv_bfloat16* gpre = pbase ? pbase - static_cast<size_t>(row) * kTileK : nullptr;
        const int kt_out = p.NB == 1 ? (p.out_kpad >> 6) : (p.RB >> 6);
#define RLSB_P2(ACT, LN, SAVE)                                                                                            \
  do {                                                                                                                    \
    if (p.staged_out)                                                                                                     \
      ln_act_pass2_staged<ACT, LN, SAVE>(tmem_d, cq, n_valid, kt_out, row, tid_e, s_bias, s_gam, s_bet, rstd, nmr, \
                                         row_ok, stage_out, stage_pre, gout, gpre, out_ph, out_pending);                  \
    else                                                                                                                  \
      ln_act_pass2<ACT, LN, SAVE>(tmem_d, cq, my_chunks, n_valid, col0, row, s_bias, s_gam, s_bet, rstd, nmr, obase,      \
                                  pbase, row_ok, n_fast);                                                                 \
  } while (0)
        if (EPI == EPI_LN_ACT_SAVE) {   // training forward (ELU only): also keep x_hat / the pre-activation
          if (has_ln) RLSB_P2(ACT_ELU, true, EPI == EPI_LN_ACT_SAVE);
          else RLSB_P2(ACT_ELU, false, EPI == EPI_LN_ACT_SAVE);
        } else if (has_ln) {
          if (p.act == ACT_ELU) RLSB_P2(ACT_ELU, true, false);
          else if (p.act == ACT_RELU) RLSB_P2(ACT_RELU, true, false);
          else RLSB_P2(ACT_NONE, true, false);
        } else {
          if (p.act == ACT_ELU) RLSB_P2(ACT_ELU, false, false);
          else if (p.act == ACT_RELU) RLSB_P2(ACT_RELU, false, false);
          else RLSB_P2(ACT_NONE, false, false);
        }
#undef RLSB_P2
        // zero the padding columns [RB, out_kpad) of the packed image (last block only; the staged path wrote them)
        if (nb == p.NB - 1 && !p.staged_out) {
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
      if (lane == 0) {
        if (pair && rank == 1) mbar_arrive_cluster(mapa_cluster(smem_u32(&ctl->tmem_empty[buf]), 0u));
        else mbar_arrive(&ctl->tmem_empty[buf]);
      }
    }
    if (bwd_colsum) flush_colsum(g_acc);
    if ((kLnAct || EPI == EPI_GRU) && tid_e == 0 && out_pending) bulk_wait<0>();   // staged output: every bulk store has landed
  }

  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();   // nobody leaves while a peer may still write its smem / barriers
  if (warp == 1) {
    tc_fence_after();
    if constexpr (pair) tmem_dealloc2(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

int g_num_sms = 0;
int g_pair = 2;   // cta_group::2 for clusters of 2 (1: only n-blocks <= 256 columns, 0: multicast only); RLSB_PAIR overrides
int g_cluster_size = 2;   // CTAs per cluster sharing a weight block (1, 2 or 4); RLSB_CLUSTER overrides
int g_staged = 1;         // full-row epilogues write their output through shared memory + bulk copies (RLSB_STAGED=0: 16-byte stores)
int g_staged_f32 = 1;     // fp32 row-major outputs go through shared-memory slabs and leave as full lines (RLSB_STAGED_F32=0)
unsigned long long* g_trace = nullptr;
int g_gru_clusters = 0;   // experiment: cap on the clusters of an EPI_GRU launch (RLSB_GRU_CLUSTERS)
int g_bwd_xbuf = 1;       // EPI_BWD prefetches its saved x_hat chunks into shared memory with cp.async (RLSB_BWD_XBUF=0: global loads)

}  // namespace

void set_gemm_trace(unsigned long long* device_buffer) { g_trace = device_buffer; }

int set_gemm_staged_output(int on) {
  if (on == 0 || on == 1) g_staged = on;
  return g_staged;
}
void set_gemm_cluster_size(int cs) {
  if (cs == 1 || cs == 2 || cs == 4) g_cluster_size = cs;
}
int gemm_cluster_size() { return g_cluster_size; }

namespace {
int init_device_info() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return static_cast<int>(e);
    if (const char* env = getenv("RLSB_CLUSTER")) set_gemm_cluster_size(atoi(env));
    if (const char* env = getenv("RLSB_PAIR")) g_pair = atoi(env);   // 0: multicast only, 1: pairs for RB <= 256, 2: all
    if (const char* env = getenv("RLSB_STAGED")) g_staged = atoi(env);
    if (const char* env = getenv("RLSB_STAGED_F32")) g_staged_f32 = atoi(env);
    if (const char* env = getenv("RLSB_BWD_XBUF")) g_bwd_xbuf = atoi(env);
    if (const char* env = getenv("RLSB_GRU_CLUSTERS")) g_gru_clusters = atoi(env);
  }
  return 0;
}
int pick_cluster(const GemmParams& p) {
  int cs = g_cluster_size;
  while (cs > 1 && ((p.RB / cs) % 8 != 0 || p.m_tiles < cs)) cs >>= 1;
  return cs;
}
}  // namespace

namespace {
// clusters of a launch configuration the current device keeps resident at once (cudaOccupancyMaxActiveClusters), cached per
// (device, kernel variant, cluster size, shared memory)
template <class K>
int resident_clusters(K kernel, const cudaLaunchConfig_t& cfg, int variant) {
  struct Entry { int dev, variant, cs; size_t smem; int n; };
  static Entry cache[32];
  static int used = 0;
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  const int cs = static_cast<int>(cfg.attrs[0].val.clusterDim.x);
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < used; ++i)
    if (cache[i].dev == dev && cache[i].variant == variant && cache[i].cs == cs && cache[i].smem == cfg.dynamicSmemBytes)
      return cache[i].n;
  cudaLaunchConfig_t probe = cfg;
  probe.numAttrs = 1;   // the cluster dimension only
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kernel, &probe) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (used < 32) cache[used++] = Entry{dev, variant, cs, cfg.dynamicSmemBytes, n};
  return n;
}
}  // namespace

int gemm_grid_size(const GemmParams& p) {
  if (init_device_info() != 0) return 0;
  const int cs = pick_cluster(p);
  const int total_work = p.G * p.NB * ((p.m_tiles + cs - 1) / cs);
  int clusters = g_num_sms / cs;
  if (total_work < clusters) clusters = total_work;
  return clusters * cs;
}

int launch_gemm(const GemmParams& p, int epilogue, cudaStream_t stream) {
  if (p.RB <= 0 || p.RB > 512 || (p.RB % 32) != 0) return -1;
  if (p.n_seg < 1 || p.n_seg > kMaxSeg) return -2;
  if (epilogue == EPI_LN_ACT && p.save_pre) epilogue = EPI_LN_ACT_SAVE;
  // LayerNorm needs the whole row: one block, or the cross-block exchange
  const bool xln = (epilogue == EPI_LN_ACT || epilogue == EPI_GRU) && p.NB != 1 && p.ln_gamma != nullptr;
  if ((epilogue == EPI_LN_ACT || epilogue == EPI_LN_ACT_SAVE) && p.NB != 1 && p.ln_gamma != nullptr && !p.xstats) return -3;
  if (xln && (!p.xstats || p.G != 1 || p.N != p.NB * p.RB || p.row_period != 0)) return -3;
  if (epilogue == EPI_GRU && (p.RB != 192 || !p.ln_gamma || !p.gru_h_prev || !p.gru_h_next || !p.out_bf16 || p.G != 1 ||
                              p.out_kpad != p.NB * 64 || (p.gru_ld_h % 4) != 0 || (p.gru_ld_hn % 4) != 0 ||
                              (reinterpret_cast<uintptr_t>(p.gru_h_prev) % 16) != 0 || (reinterpret_cast<uintptr_t>(p.gru_h_next) % 16) != 0))
    return -10;
  if (p.M <= 0 || p.m_tiles != (p.M + kTileM - 1) / kTileM) return -4;
  if ((epilogue == EPI_LN_ACT || epilogue == EPI_LN_ACT_SAVE || epilogue == EPI_BWD) && ((p.out_kpad % 64) != 0 || p.out_kpad < p.N)) return -5;
  if (epilogue == EPI_BWD && ((p.NB != 1 && p.ln_gamma) || !p.bwd_pre || !p.out_bf16 || !p.group_major ||
                              (p.ln_gamma && !p.bwd_rstd) || (p.ln_gamma && p.act == ACT_RELU)))
    return -8;
  if (epilogue == EPI_LN_ACT_SAVE && (p.act != ACT_ELU || p.NB != 1)) return -9;
  {
    const int ie = init_device_info();
    if (ie != 0) return ie;
  }
  // cluster size along M: CTAs of a cluster share the weight block (TMA multicast, or one MMA over the CTA pair)
  const int cs = pick_cluster(p);
  const int pair = (g_pair && cs == 2 && (p.RB % 32) == 0 && (p.RB <= 256 || g_pair > 1)) ? 1 : 0;
  const int stage_bytes = kTileM * kTileK * 2 + p.RB * kTileK * 2 / (pair ? 2 : 1);
  // full-row epilogues that own whole output tiles assemble them in shared memory (ln_act_pass2_staged)
  const bool staged_f32 = g_staged_f32 && (epilogue == EPI_PLAIN || epilogue == EPI_STATS) && (p.N % 32) == 0 &&
                          (p.ldo % 4) == 0 && p.out_f32 != nullptr && (reinterpret_cast<uintptr_t>(p.out_f32) % 16) == 0 &&
                          (p.out_group_stride % 4) == 0;
  const bool staged = staged_f32 || (g_staged && (epilogue == EPI_LN_ACT || epilogue == EPI_LN_ACT_SAVE) &&
                                     (p.NB == 1 || ((p.RB % 64) == 0 && p.NB * p.RB == p.out_kpad)));
  const int staging_bytes = (staged || epilogue == EPI_GRU) ? (epilogue == EPI_LN_ACT_SAVE ? 65536 : 32768) : 0;
  int budget = 227 * 1024 - 1024 /*align*/ - static_cast<int>(sizeof(SmemCtl)) - 256 - staging_bytes;
  static_assert(sizeof(SmemCtl) < 24 * 1024, "control block grew");
  // EPI_BWD: a 16-byte slot per epilogue thread and chunk for the saved image (two stages must still fit)
  // (latency-bound launches only — a tile or two per CTA: dino step 5.52 -> 5.30 ms; with many tiles per CTA the loads of
  // the next tile already overlap and the deeper stage ring is worth more: sweep step 33.5 vs 34.0 ms; RLSB_BWD_XBUF=2 forces it)
  int xbuf_bytes = (epilogue == EPI_BWD && p.NB == 1 && g_bwd_xbuf &&
                    (g_bwd_xbuf > 1 || p.m_tiles * p.G <= 2 * g_num_sms)) ? (p.RB >> 5) * kEpiThreads * 16 : 0;
  if (xbuf_bytes > 0 && (budget - xbuf_bytes) / stage_bytes < 2) xbuf_bytes = 0;
  budget -= xbuf_bytes;
  int stages = budget / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return -6;
  const int nbuf = p.RB <= 256 ? 2 : 1;
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + staging_bytes + xbuf_bytes + sizeof(SmemCtl) + 1024;
  GemmParams q = p;
  q.staged_out = staged ? 1 : 0;
  q.xbuf_bytes = xbuf_bytes;
  q.inv_n = 1.0f / static_cast<float>(p.N);
  q.trace = epilogue == EPI_GRU ? g_trace : nullptr;
  const int total_work = p.G * p.NB * ((p.m_tiles + cs - 1) / cs);
  int clusters = g_num_sms / cs;
  if (epilogue == EPI_GRU && g_gru_clusters > 0 && g_gru_clusters < clusters) clusters = g_gru_clusters;
  if (total_work < clusters) clusters = total_work;
  const int grid = clusters * cs;

  cudaError_t e;
  if (epilogue == EPI_BWD && p.col_part && p.ln_gamma) {
    e = cudaMemsetAsync(p.col_part, 0, static_cast<size_t>(grid) * p.G * 2 * p.RB * sizeof(float), stream);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(cs);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = g_pdl;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
#define RLSB_LAUNCH(EPI)                                                                         \
  do {                                                                                           \
    static PerDeviceOnce attr_once;                                                              \
    unsigned long long dev_bit = 0;                                                              \
    if (attr_once.need(dev_bit)) {                                                               \
      e = cudaFuncSetAttribute(gemm_kernel<EPI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                               227 * 1024);                                                      \
      if (e != cudaSuccess) return static_cast<int>(e);                                          \
      e = cudaFuncSetAttribute(gemm_kernel<EPI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                               227 * 1024);                                                      \
      if (e != cudaSuccess) return static_cast<int>(e);                                          \
      attr_once.done(dev_bit);                                                                   \
    }                                                                                            \
    if (xln) { /* the cross-block exchange spins on partner CTAs: every cluster of the launch must be resident */ \
      const int cap = pair ? resident_clusters(gemm_kernel<EPI, true>, cfg, EPI * 2 + 1)         \
                           : resident_clusters(gemm_kernel<EPI, false>, cfg, EPI * 2);           \
      if (cap <= 0) return -11;                                                                  \
      if (clusters > cap) cfg.gridDim = dim3(static_cast<unsigned>(cap * cs));                   \
    }                                                                                            \
    if (pair) e = cudaLaunchKernelEx(&cfg, gemm_kernel<EPI, true>, q, stages, nbuf, cs);         \
    else e = cudaLaunchKernelEx(&cfg, gemm_kernel<EPI, false>, q, stages, nbuf, cs);             \
    if (e != cudaSuccess) return static_cast<int>(e);                                            \
  } while (0)
  switch (epilogue) {
    case EPI_PLAIN: RLSB_LAUNCH(EPI_PLAIN); break;
    case EPI_STATS: RLSB_LAUNCH(EPI_STATS); break;
    case EPI_LN_ACT: RLSB_LAUNCH(EPI_LN_ACT); break;
    case EPI_BWD: RLSB_LAUNCH(EPI_BWD); break;
    case EPI_LN_ACT_SAVE: RLSB_LAUNCH(EPI_LN_ACT_SAVE); break;
    case EPI_GRU: RLSB_LAUNCH(EPI_GRU); break;
    default: return -7;
  }
#undef RLSB_LAUNCH
  count_launch();
  e = cudaGetLastError();
  return static_cast<int>(e);
}

}  // namespace rlsb
