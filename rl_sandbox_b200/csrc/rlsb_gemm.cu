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
#include "rlsb_gemm.cuh"
#include "rlsb_count.cuh"
#include "rlsb_ptx.cuh"

namespace rlsb {

namespace {

struct SmemCtl {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ float act_apply(float x, int act) {
  if (act == ACT_ELU) return x > 0.f ? x : expm1f(x);
  if (act == ACT_RELU) return fmaxf(x, 0.f);
  return x;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const GemmParams p, const int stages, const int nbuf) {
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
  const int total_work = p.G * p.NB * p.m_tiles;
  const int buf_cols = (nbuf == 2) ? 256 : 512;

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < stages; ++s) {
        mbar_init(&ctl->full[s], 1);
        mbar_init(&ctl->empty[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&ctl->tmem_full[b], 1);
        mbar_init(&ctl->tmem_empty[b], 128);
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
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ===================== producer: bulk copies into the stage ring =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int m_tile = w % p.m_tiles;
        const int gnb = w / p.m_tiles;  // g * NB + nb
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
            bulk_g2s(sb, wsrc + static_cast<size_t>(kt_glob) * (static_cast<size_t>(p.RB) * kTileK),
                     b_bytes, &ctl->full[stage]);
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
      for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++it) {
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
          umma_commit(&ctl->empty[stage]);  // frees the smem stage when these MMAs retire
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&ctl->tmem_full[buf]);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue: 4 warps, one TMEM lane (row) per thread ===============
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;        // row inside the 128-row tile
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int m_pad = p.m_tiles * kTileM;
    int it = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++it) {
      const int m_tile = w % p.m_tiles;
      const int gnb = w / p.m_tiles;
      const int g = gnb / p.NB;
      const int nb = gnb - g * p.NB;
      const int buf = (nbuf == 2) ? (it & 1) : 0;
      const uint32_t use = (nbuf == 2) ? static_cast<uint32_t>(it >> 1) : static_cast<uint32_t>(it);
      mbar_wait(&ctl->tmem_full[buf], use & 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * buf_cols) + lane_addr;
      const int m = m_tile * kTileM + row;
      const int col0 = nb * p.RB;                     // first column of this block inside the group
      const int n_valid = min(p.RB, p.N - col0);      // valid columns in this block (may be <= 0)
      const float* bias = p.bias ? p.bias + static_cast<size_t>(g) * p.NB * p.RB + col0 : nullptr;

      if (EPI == EPI_PLAIN || EPI == EPI_STATS) {
        float* orow = p.out_f32 + static_cast<size_t>(g) * p.out_group_stride +
                      static_cast<size_t>(m) * p.ldo + col0;
        const bool vec_ok = ((p.ldo & 3) == 0) && ((col0 & 3) == 0);
        float sum = 0.f;
        for (int c = 0; c < p.RB; c += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_d + static_cast<uint32_t>(c), r);
          tmem_ld_wait();
          if (m < p.M && c < n_valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float v[4];
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                v[t] = __uint_as_float(r[j + t]) + (bias ? __ldg(bias + c + j + t) : 0.f);
                if (c + j + t < n_valid) sum += v[t];
              }
              if (vec_ok && c + j + 3 < n_valid) {
                *reinterpret_cast<float4*>(orow + c + j) = make_float4(v[0], v[1], v[2], v[3]);
              } else {
#pragma unroll
                for (int t = 0; t < 4; ++t)
                  if (c + j + t < n_valid) orow[c + j + t] = v[t];
              }
            }
          }
        }
        if (EPI == EPI_STATS) {
          const float mean = n_valid > 0 ? sum / static_cast<float>(n_valid) : 0.f;
          float m2 = 0.f;
          for (int c = 0; c < p.RB; c += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_d + static_cast<uint32_t>(c), r);
            tmem_ld_wait();
            if (c < n_valid) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (c + j < n_valid) {
                  const float d = __uint_as_float(r[j]) + (bias ? __ldg(bias + c + j) : 0.f) - mean;
                  m2 += d * d;
                }
              }
            }
          }
          float2* st = reinterpret_cast<float2*>(p.stats) +
                       (static_cast<size_t>(gnb) * m_pad + static_cast<size_t>(m));
          *st = make_float2(mean, m2);
        }
      } else {  // EPI_LN_ACT : the block holds the whole row (NB == 1)
        const float* gam = p.ln_gamma ? p.ln_gamma + static_cast<size_t>(g) * p.RB : nullptr;
        const float* bet = p.ln_beta ? p.ln_beta + static_cast<size_t>(g) * p.RB : nullptr;
        float mean = 0.f, rstd = 1.f;
        if (gam) {
          float sum = 0.f;
          for (int c = 0; c < n_valid; c += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_d + static_cast<uint32_t>(c), r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c + j < n_valid) sum += __uint_as_float(r[j]) + (bias ? __ldg(bias + c + j) : 0.f);
          }
          mean = sum / static_cast<float>(n_valid);
          float m2 = 0.f;
          for (int c = 0; c < n_valid; c += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_d + static_cast<uint32_t>(c), r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c + j < n_valid) {
                const float d = __uint_as_float(r[j]) + (bias ? __ldg(bias + c + j) : 0.f) - mean;
                m2 += d * d;
              }
          }
          rstd = 1.0f / sqrtf(m2 / static_cast<float>(n_valid) + p.ln_eps);
        }
        __nv_bfloat16* obase = p.out_bf16 + static_cast<size_t>(g) * p.out_bf16_group_stride;
        const int out_ktiles = p.out_kpad >> 6;
        for (int c = 0; c < p.out_kpad; c += 32) {
          uint32_t r[32];
          if (c < p.RB) {  // warp-uniform
            tmem_ld32(tmem_d + static_cast<uint32_t>(c), r);
            tmem_ld_wait();
          }
          float y[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float v = 0.f;
            if (c + j < n_valid) {
              v = __uint_as_float(r[j]) + (bias ? __ldg(bias + c + j) : 0.f);
              if (gam) v = (v - mean) * rstd * __ldg(gam + c + j) + __ldg(bet + c + j);
              v = act_apply(v, p.act);
            }
            y[j] = v;
          }
          // packed store: tile (m_tile, c/64), row `row`, chunks (c%64)/8 .. +3, swizzled
          const int kt = c >> 6;
          __nv_bfloat16* trow = obase + (static_cast<size_t>(m_tile) * out_ktiles + kt) *
                                            (kTileM * kTileK) +
                                static_cast<size_t>(row) * kTileK;
          const int chunk0 = (c & 63) >> 3;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint4 pk;
            pk.x = pack_bf16x2(y[ch * 8 + 0], y[ch * 8 + 1]);
            pk.y = pack_bf16x2(y[ch * 8 + 2], y[ch * 8 + 3]);
            pk.z = pack_bf16x2(y[ch * 8 + 4], y[ch * 8 + 5]);
            pk.w = pack_bf16x2(y[ch * 8 + 6], y[ch * 8 + 7]);
            const int chunk = (chunk0 + ch) ^ (row & 7);
            *reinterpret_cast<uint4*>(trow + chunk * 8) = pk;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&ctl->tmem_empty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int g_num_sms = 0;

}  // namespace

int launch_gemm(const GemmParams& p, int epilogue, cudaStream_t stream) {
  if (p.RB <= 0 || p.RB > 512 || (p.RB % 32) != 0) return -1;
  if (p.n_seg < 1 || p.n_seg > kMaxSeg) return -2;
  if (epilogue == EPI_LN_ACT && p.NB != 1) return -3;
  if (p.M <= 0 || p.m_tiles != (p.M + kTileM - 1) / kTileM) return -4;
  if (epilogue == EPI_LN_ACT && ((p.out_kpad % 64) != 0 || p.out_kpad < p.N)) return -5;
  if (g_num_sms == 0) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  const int stage_bytes = kTileM * kTileK * 2 + p.RB * kTileK * 2;
  const int budget = 227 * 1024 - 1024 /*align*/ - static_cast<int>(sizeof(SmemCtl)) - 256;
  int stages = budget / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return -6;
  const int nbuf = p.RB <= 256 ? 2 : 1;
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + sizeof(SmemCtl) + 1024;
  const int total_work = p.G * p.NB * p.m_tiles;
  const int grid = total_work < g_num_sms ? total_work : g_num_sms;

  cudaError_t e;
#define RLSB_LAUNCH(EPI)                                                                         \
  do {                                                                                           \
    static bool attr_done = false;                                                               \
    if (!attr_done) {                                                                            \
      e = cudaFuncSetAttribute(gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                               227 * 1024);                                                      \
      if (e != cudaSuccess) return static_cast<int>(e);                                          \
      attr_done = true;                                                                          \
    }                                                                                            \
    gemm_kernel<EPI><<<grid, kGemmThreads, smem, stream>>>(p, stages, nbuf);                     \
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
