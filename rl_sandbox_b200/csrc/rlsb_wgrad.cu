// rlsb_wgrad.cu — weight gradients of the actor / critic MLPs on tcgen05 (see rlsb_wgrad.cuh).
//
//   dW[g][n][k] = sum_m dY[g][m][n] * X[g][m][k]
//
// One CTA owns one output tile (128 n-rows x up to 512 k-columns, fp32 in TMEM) and a contiguous
// range of M tiles ("split"); partial tiles are written to HBM and summed by wgrad_reduce_kernel
// in a fixed order (deterministic; no atomics), which also scatters them into the nn.Linear
// weight / bias gradient layout.
//
// The contraction runs over the ROW index of the packed activation images, i.e. both operands are
// MN-major for the tensor core.  A pipeline stage holds 32 rows of every participating 64-column
// tile: 4 KB pieces, each one contiguous cp.async.bulk, laid out so that consecutive 64-column
// groups are 4 KB apart (LBO) and 8-row groups 1 KB apart (SBO).
#include "rlsb_wgrad.cuh"

#include <cstdlib>

#include "rlsb_count.cuh"
#include "rlsb_gemm.cuh"
#include "rlsb_ptx.cuh"

namespace rlsb {

namespace {

constexpr int kWgThreads = 192;       // warp 0 producer, warp 1 MMA issuer, warps 2-5 epilogue
constexpr int kWgRows = 32;           // contraction rows per stage
constexpr uint32_t kPiece = kWgRows * 128;  // bytes of one 64-column tile slice
constexpr int kWgMaxStages = 8;

struct WgCtl {
  uint64_t full[kWgMaxStages];
  uint64_t empty[kWgMaxStages];
  uint64_t tmem_full;
  uint32_t tmem_base;
  uint32_t pad;
};

// cs > 1: a cluster of cs CTAs works on cs consecutive n-slices of the same (group, k chunk, split).  They contract
// the same X pieces, so CTA r fetches pieces r, r + cs, ... and multicasts them into every CTA of the cluster (the
// dY pieces are private): the L2 -> SM traffic per 32-row stage drops from (2 + 8) to (2 + 8 / cs) pieces.
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_kernel(const WgradParams p, const int stages, const int cs) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- which output tile / which M range -----------------------------------------------------
  const int items = p.G * p.n_slices * p.n_chunks;
  const int item = static_cast<int>(blockIdx.x) % items;
  const int split = static_cast<int>(blockIdx.x) / items;
  const int ns = item % p.n_slices;                      // n-slice fastest: the CTAs of a cluster differ in ns only
  const int kc = (item / p.n_slices) % p.n_chunks;
  const int g = item / (p.n_chunks * p.n_slices);
  const int rank = cs > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const uint16_t cta_mask = static_cast<uint16_t>((1u << cs) - 1u);
  const int m0 = static_cast<int>(static_cast<long long>(split) * p.m_tiles / p.splits);
  const int m1 = static_cast<int>(static_cast<long long>(split + 1) * p.m_tiles / p.splits);
  const int kt0 = kc * p.kc_tiles;
  const int nkt = min(p.kc_tiles, p.kt_total - kt0);   // k tiles of this chunk (1..8)
  const int na = min(2, p.n_tiles - ns * 2);            // dY tiles of this slice (1..2)
  const uint32_t stage_bytes = static_cast<uint32_t>(2 + p.kc_tiles) * kPiece;
  WgCtl* ctl = reinterpret_cast<WgCtl*>(smem + static_cast<size_t>(stages) * stage_bytes);

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < stages; ++s) {
        mbar_init(&ctl->full[s], 1);
        mbar_init(&ctl->empty[s], static_cast<uint32_t>(cs));   // every CTA of the cluster releases the slot
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
  if (cs > 1) cluster_sync_all();   // peers' barriers are initialised before anyone signals them
  const uint32_t tmem_base = ctl->tmem_base;

  // Producer and MMA warp walk their loops as whole warps in uniform control flow; one elected lane issues the bulk
  // copies / tcgen05 instructions, whose operands then sit in uniform registers (inside `if (lane == 0)` the compiler wraps
  // every UBLKCP / UTCHMMA in an ELECT + R2UR.BROADCAST loop, several times the cost of the two MMAs a 32-row stage holds).
  if (warp == 0) {
    {
      // resolve the source of every k tile of this chunk once
      const __nv_bfloat16* xsrc[8];
      long long xstride[8];
      for (int j = 0; j < nkt; ++j) {
        int kt = kt0 + j, s = 0;
        while (kt >= p.x_ktiles[s]) {
          kt -= p.x_ktiles[s];
          ++s;
        }
        xsrc[j] = p.X[s] + static_cast<size_t>(g) * p.x_group_stride[s] + static_cast<size_t>(kt) * (kTileM * kTileK);
        xstride[j] = p.x_mtile_stride[s];
      }
      const __nv_bfloat16* ysrc = p.dY + static_cast<size_t>(g) * p.dy_group_stride +
                                  static_cast<size_t>(ns) * 2 * (kTileM * kTileK);
      const long long ystride = static_cast<long long>(p.n_tiles) * (kTileM * kTileK);
      const uint32_t tx = static_cast<uint32_t>(na + nkt) * kPiece;
      int stage = 0;
      uint32_t phase = 0;
      for (int mt = m0; mt < m1; ++mt) {
        for (int sub = 0; sub < kTileM / kWgRows; ++sub) {
          mbar_wait(&ctl->empty[stage], phase ^ 1u);
          uint8_t* sa = smem + static_cast<size_t>(stage) * stage_bytes;
          const size_t roff = static_cast<size_t>(sub) * kWgRows * kTileK;   // elements
          if (elect_one()) {
          mbar_expect_tx(&ctl->full[stage], tx);
          for (int a = 0; a < na; ++a)
            bulk_g2s(sa + a * kPiece, ysrc + mt * ystride + static_cast<size_t>(a) * (kTileM * kTileK) + roff, kPiece,
                     &ctl->full[stage]);
          if (cs == 1) {
            for (int j = 0; j < nkt; ++j)
              bulk_g2s(sa + (2 + j) * kPiece, xsrc[j] + mt * xstride[j] + roff, kPiece, &ctl->full[stage]);
          } else {
            for (int j = rank; j < nkt; j += cs)
              bulk_g2s_multicast(sa + (2 + j) * kPiece, xsrc[j] + mt * xstride[j] + roff, kPiece, &ctl->full[stage],
                                 cta_mask);
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
  } else if (warp == 1) {
    {
      const int n0 = nkt > 4 ? 256 : nkt * 64;
      const int n1 = nkt * 64 - n0;
      const uint32_t idesc0 = make_idesc_bf16_mn(128, static_cast<uint32_t>(n0));
      const uint32_t idesc1 = n1 > 0 ? make_idesc_bf16_mn(128, static_cast<uint32_t>(n1)) : 0u;
      int stage = 0;
      uint32_t phase = 0;
      bool first = true;
      for (int mt = m0; mt < m1; ++mt) {
        for (int sub = 0; sub < kTileM / kWgRows; ++sub) {
          mbar_wait(&ctl->full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + static_cast<size_t>(stage) * stage_bytes);
          const uint64_t adesc = make_smem_desc_mn_sw128(sa, kPiece);
          const uint64_t bdesc0 = make_smem_desc_mn_sw128(sa + 2 * kPiece, kPiece);
          const uint64_t bdesc1 = make_smem_desc_mn_sw128(sa + 6 * kPiece, kPiece);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < kWgRows / 16; ++kk) {
              const uint32_t acc = (first && kk == 0) ? 0u : 1u;
              const uint64_t adv = static_cast<uint64_t>(kk) * (2048u >> 4);   // 16 rows of 128 bytes
              umma_bf16(tmem_base, adesc + adv, bdesc0 + adv, idesc0, acc);
              if (n1 > 0) umma_bf16(tmem_base + 256u, adesc + adv, bdesc1 + adv, idesc1, acc);
            }
            if (cs == 1) umma_commit(&ctl->empty[stage]);
            else umma_commit_multicast(&ctl->empty[stage], cta_mask);
          }
          __syncwarp();
          first = false;
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      if (elect_one()) umma_commit(&ctl->tmem_full);
      __syncwarp();
    }
  } else {
    // ---- epilogue: TMEM -> fp32 partial tile ---------------------------------------------------
    const int q = warp & 3;
    const int n = q * 32 + lane;
    const int n_glob = ns * 128 + n;
    const int rows_pad = p.n_slices * 128;
    const int ld = p.kt_total * 64;
    mbar_wait(&ctl->tmem_full, 0u);
    tc_fence_after();
    float* dst = p.partial + ((static_cast<size_t>(split) * p.G + g) * rows_pad + n_glob) * ld + kt0 * 64;
    const bool row_ok = n_glob < p.n_tiles * 64;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    for (int c = 0; c < nkt * 64; c += 32) {
      uint32_t r[32];
      tmem_ld32(taddr + static_cast<uint32_t>(c), r);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(dst + c + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();   // nobody leaves while a peer may still signal its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

__global__ void wgrad_reduce_kernel(const WgradReduceParams p) {
  const long long total = static_cast<long long>(p.G) * p.rows_pad * p.ld;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int col = static_cast<int>(i % p.ld);
    const int n = static_cast<int>((i / p.ld) % p.rows_pad);
    const int g = static_cast<int>(i / (static_cast<long long>(p.ld) * p.rows_pad));
    if (n >= p.n_out[g]) continue;
    float* dst = nullptr;
    if (col == p.ones_col) {
      if (p.b_dst[g]) dst = p.b_dst[g] + n;
    } else if (p.w_dst[g]) {
      for (int s = 0; s < p.n_seg; ++s) {
        const int off = col - p.seg[s].dst_k0;
        if (off >= 0 && off < p.seg[s].len) dst = p.w_dst[g] + static_cast<size_t>(n) * p.ld_dst + p.seg[s].src_c0 + off;
      }
    }
    if (!dst) continue;
    float acc = 0.f;
    const size_t stride = static_cast<size_t>(p.G) * p.rows_pad * p.ld;
    const float* src = p.partial + i;
#pragma unroll 8
    for (int s = 0; s < p.splits; ++s) acc += __ldg(src + s * stride);   // (same order: the loads of eight splits are in flight)
    *dst = p.accumulate ? *dst + acc : acc;
  }
}

// eight lanes per output column: each sums every eighth CTA's partial (all loads in flight), then a three-step butterfly —
// a fixed order, so the sums are reproducible (a single thread walking the <= 148 partials took 22 us of L2 round trips)
__global__ void colsum_reduce_kernel(const ColsumReduceParams p) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = t >> 3, l = t & 7;
  const int total = p.G * 2 * p.RB;
  const int col = i % p.RB;
  const int which = (i / p.RB) & 1;
  const int g = i / (2 * p.RB);
  float* dst = nullptr;
  if (i < total && col < p.N) dst = which == 0 ? p.dgamma[g] : p.dbeta[g];
  float acc = 0.f;
  if (dst) {
    const size_t stride = static_cast<size_t>(p.G) * 2 * p.RB;
#pragma unroll 4
    for (int c = l; c < p.ctas; c += 8) acc += __ldg(p.col_part + c * stride + i);
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (dst && l == 0) dst[col] = acc;
}

PerDeviceOnce g_wg_attr_once;
int g_wg_sms = 0;
int g_wg_cluster = 4;          // max CTAs per cluster (1, 2 or 4); RLSB_WGRAD_CLUSTER overrides
int g_wg_max_ctas[5] = {0, 0, 0, 0, 0};   // co-resident CTAs for cluster size 1 / 2 / 4 (148 / 148 / 132 on B200)

int wg_stage_bytes(const WgradParams& p) { return (2 + p.kc_tiles) * static_cast<int>(kPiece); }
int wg_stages(const WgradParams& p) {
  int stages = (227 * 1024 - 1024 - static_cast<int>(sizeof(WgCtl)) - 256) / wg_stage_bytes(p);
  return stages > kWgMaxStages ? kWgMaxStages : stages;
}
int wg_cluster(const WgradParams& p) {
  int cs = g_wg_cluster;
  while (cs > 1 && (p.n_slices % cs) != 0) cs >>= 1;
  return cs;
}

}  // namespace

int plan_wgrad(WgradParams& p) {
  if (p.n_seg < 1 || p.n_seg > kWgMaxSeg || p.G < 1 || p.G > kWgMaxGroups || p.n_tiles < 1 || p.m_tiles < 1) return -1;
  if (g_wg_sms == 0) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaDeviceGetAttribute(&g_wg_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return static_cast<int>(e);
    if (const char* env = getenv("RLSB_WGRAD_CLUSTER")) {
      const int v = atoi(env);
      if (v == 1 || v == 2 || v == 4) g_wg_cluster = v;
    }
    g_wg_max_ctas[1] = g_wg_sms;
    {   // also the first device's attribute: the occupancy query below needs it
      unsigned long long bit = 0;
      g_wg_attr_once.need(bit);
      e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return static_cast<int>(e);
      g_wg_attr_once.done(bit);
    }
    for (int cs = 2; cs <= 4; cs *= 2) {   // how many clusters of this size the device holds at once
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(static_cast<unsigned>(g_wg_sms / cs * cs));
      cfg.blockDim = dim3(kWgThreads);
      cfg.dynamicSmemBytes = 200 * 1024;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = static_cast<unsigned>(cs);
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      int n = 0;
      e = cudaOccupancyMaxActiveClusters(&n, wgrad_kernel, &cfg);
      if (e != cudaSuccess || n < 1) {
        (void)cudaGetLastError();
        n = g_wg_sms / cs / 2;   // conservative
      }
      g_wg_max_ctas[cs] = n * cs;
    }
  }
  {   // the shared-memory attribute is per device: a process that moved to another GPU sets it there too
    unsigned long long bit = 0;
    if (g_wg_attr_once.need(bit)) {
      const cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return static_cast<int>(e);
      g_wg_attr_once.done(bit);
    }
  }
  p.kt_total = 0;
  for (int s = 0; s < p.n_seg; ++s) p.kt_total += p.x_ktiles[s];
  const int chunks = (p.kt_total + 7) / 8;
  p.kc_tiles = (p.kt_total + chunks - 1) / chunks;
  p.n_chunks = (p.kt_total + p.kc_tiles - 1) / p.kc_tiles;
  p.n_slices = (p.n_tiles + 1) / 2;
  const int items = p.G * p.n_slices * p.n_chunks;
  int splits = g_wg_max_ctas[wg_cluster(p)] / items;   // one wave of co-resident CTAs (clusters)
  if (splits < 1) splits = 1;
  if (splits > p.m_tiles) splits = p.m_tiles;
  p.splits = splits;
  return 0;
}

size_t wgrad_partial_bytes(const WgradParams& p) {
  return static_cast<size_t>(p.splits) * p.G * p.n_slices * 128 * p.kt_total * 64 * sizeof(float);
}

int launch_wgrad(const WgradParams& p, cudaStream_t stream) {
  if (!p.dY || !p.partial || p.splits < 1 || p.kc_tiles < 1 || p.kc_tiles > 8) return -1;
  const int stage_bytes = wg_stage_bytes(p);
  const int stages = wg_stages(p);
  if (stages < 2) return -2;
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + sizeof(WgCtl) + 1024;
  const int cs = wg_cluster(p);
  const int grid = p.G * p.n_slices * p.n_chunks * p.splits;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kWgThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(cs);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, wgrad_kernel, p, stages, cs);
  if (e != cudaSuccess) return static_cast<int>(e);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_wgrad_reduce(const WgradReduceParams& p, cudaStream_t stream) {
  const long long total = static_cast<long long>(p.G) * p.rows_pad * p.ld;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  wgrad_reduce_kernel<<<blocks, 256, 0, stream>>>(p);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_colsum_reduce(const ColsumReduceParams& p, cudaStream_t stream) {
  const int total = p.G * 2 * p.RB * 8;   // eight lanes per output
  colsum_reduce_kernel<<<(total + 127) / 128, 128, 0, stream>>>(p);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

}  // namespace rlsb
