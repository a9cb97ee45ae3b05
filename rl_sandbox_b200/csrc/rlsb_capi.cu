// rlsb_capi.cu — extern "C" surface declared in include/rlsb.h (everything except the K1
// entry points, which live next to their planner in rlsb_imagine.cu, and K3 in rlsb_slot.cu).
#include <cstdlib>

#include "../../include/rlsb.h"

#include "rlsb_count.cuh"
#include "rlsb_detmath.h"
#include "rlsb_gemm.cuh"
#include "rlsb_imagine_plan.cuh"
#include "rlsb_kernels.cuh"
#include "rlsb_wgrad.cuh"

using namespace rlsb;

namespace rlsb {
std::atomic<long long> g_launches{0};
int g_pdl = [] {
  const char* env = getenv("RLSB_PDL");
  return (env && atoi(env) == 0) ? 0 : 1;
}();
}

namespace {
__global__ void philox_uniform_kernel(uint64_t seed, uint32_t n0, uint32_t t, uint32_t stream_id,
                                      int per_row, long long count, float* __restrict__ out) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= count) return;
  const uint32_t n = n0 + static_cast<uint32_t>(i / per_row);
  const uint32_t e = static_cast<uint32_t>(i % per_row);
  out[i] = rlsb_noise_uniform(seed, n, t, stream_id, e);
}
}  // namespace

extern "C" int rlsb_abi_version(void) { return RLSB_ABI_VERSION; }

extern "C" long long rlsb_launch_count(int reset) {
  const long long v = g_launches.load();
  if (reset) g_launches.store(0);
  return v;
}

extern "C" long long rlsb_launch_count_add(long long n) {
  g_launches.fetch_add(n);
  return g_launches.load();
}

extern "C" int rlsb_set_cluster_size(int cs) {
  set_gemm_cluster_size(cs);
  return gemm_cluster_size();
}

extern "C" int rlsb_set_staged_output(int on) { return set_gemm_staged_output(on); }

extern "C" void rlsb_gemm_set_trace(void* device_buffer) { set_gemm_trace(static_cast<unsigned long long*>(device_buffer)); }

extern "C" int rlsb_set_fused_rssm(int on) {
  if (on == 0 || on == 1 || on == 2) k1::g_fused_rssm = on;   // 2: img_in / prior1 only (A/B runs)
  return k1::g_fused_rssm;
}

extern "C" int rlsb_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return -100;
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10 || minor != 0) return -101;  // built for sm_100a only; no fallback path exists
  return 0;
}

extern "C" const char* rlsb_error_string(int code) {
  if (code == 0) return "ok";
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  switch (code) {
    case -100: return "no CUDA device";
    case -101: return "device is not sm_100 (B200); librlsb has no fallback";
    case -11: return "cross-block LayerNorm: the device reports no resident cluster for this launch configuration";
    default: return "invalid argument";
  }
}

extern "C" int rlsb_lambda_return_fwd(const float* r, const float* v, const float* d, int T, int64_t N,
                                      double lambda_, float* vs, float* w, float* adv,
                                      int layout_batch_major, void* stream) {
  if (!r || !v || !d || !vs) return -1;
  return launch_lambda_return(r, v, d, T, N, lambda_, vs, w, adv, layout_batch_major,
                              static_cast<cudaStream_t>(stream));
}

extern "C" int rlsb_lambda_return_bwd(const float* g_vs, const float* v, const float* d, const float* vs,
                                      int T, int64_t N, double lambda_, float* g_r, float* g_v, float* g_d,
                                      void* stream) {
  if (!g_vs || !v || !d || !vs) return -1;
  return launch_lambda_return_bwd(g_vs, v, d, vs, T, N, lambda_, g_r, g_v, g_d,
                                  static_cast<cudaStream_t>(stream));
}

extern "C" int rlsb_sample_categorical(const float* logits, const float* uniforms, int64_t rows, int classes,
                                       int32_t* idx, void* stream) {
  if (!logits || !uniforms || !idx || classes <= 0) return -1;
  return launch_sample_categorical(logits, uniforms, rows, classes, idx, static_cast<cudaStream_t>(stream));
}

extern "C" int rlsb_sample_latent(const float* logits, int64_t rows, int groups, const float* uniforms, uint64_t seed,
                                  uint32_t row_offset, uint32_t step, uint8_t* idx, float* onehot_f32, void* stream) {
  if (!logits || !idx || rows <= 0 || rows > (1LL << 30) || groups <= 0 || groups > 64) return -1;
  NoiseSpec ns{};
  ns.explicit_noise = uniforms; ns.ld = static_cast<long long>(groups) * 32; ns.seed = seed; ns.seed_ptr = nullptr;
  ns.step = step; ns.row_offset = row_offset;
  return launch_sample_latent(logits, static_cast<long long>(groups) * 32, static_cast<int>(rows), groups, 32, ns, idx,
                              nullptr, 0, onehot_f32, static_cast<long long>(groups) * 32,
                              static_cast<cudaStream_t>(stream));
}

extern "C" int rlsb_philox_uniform(uint64_t seed, uint32_t n0, uint32_t t, uint32_t stream_id, int per_row,
                                   int64_t count, float* out, void* stream) {
  if (!out || per_row <= 0 || count <= 0) return -1;
  philox_uniform_kernel<<<static_cast<unsigned>((count + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      seed, n0, t, stream_id, per_row, count, out);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

extern "C" int rlsb_pack_rows(const float* src, int64_t ld_src, int rows_src, void* dst_bf16, int row_block,
                              int rows_dst_pad, int k_pad, int dst_k0, int src_c0, int len, void* stream) {
  if (!src || !dst_bf16) return -1;
  PackSeg seg{dst_k0, src_c0, len};
  return launch_pack(src, ld_src, rows_src, static_cast<__nv_bfloat16*>(dst_bf16), row_block, rows_dst_pad,
                     k_pad, 1, &seg, static_cast<cudaStream_t>(stream));
}

extern "C" int rlsb_gemm_bias(const void* a_packed, int k_pad, const void* w_packed, int rb, int n_blocks,
                              const float* bias_padded, int M, int N, float* out, int64_t ldo, float* stats,
                              void* stream) {
  if (!a_packed || !w_packed || !out || (k_pad % 64) != 0) return -1;
  GemmParams g{};
  g.A[0] = static_cast<const __nv_bfloat16*>(a_packed);
  g.a_ktiles[0] = k_pad / 64;
  g.n_seg = 1;
  g.W = static_cast<const __nv_bfloat16*>(w_packed);
  g.RB = rb; g.NB = n_blocks; g.G = 1;
  g.M = M; g.m_tiles = (M + 127) / 128; g.N = N;
  g.bias = bias_padded;
  g.out_f32 = out; g.ldo = ldo; g.stats = stats;
  return launch_gemm(g, stats ? EPI_STATS : EPI_PLAIN, static_cast<cudaStream_t>(stream));
}

// ---- GRUCell as one launch (EPI_GRU) ---------------------------------------------------------------------------------
namespace {
struct GruCellLayout {
  int kx, kp;               // padded K of x, of cat[x, h]
  size_t w, b, g, e, bytes; // weight image, bias, LayerNorm gain / offset (permuted, fp32 [3D])
};
bool gru_cell_layout(int Dx, int D, GruCellLayout& L) {
  if (Dx <= 0 || D <= 0 || (D % 64) != 0 || 3 * D <= 512) return false;
  L.kx = (Dx + 63) / 64 * 64;
  L.kp = L.kx + D;
  L.w = 0;
  L.b = (static_cast<size_t>(3) * D * L.kp * 2 + 1023) / 1024 * 1024;
  L.g = L.b + static_cast<size_t>(3) * D * 4;
  L.e = L.g + static_cast<size_t>(3) * D * 4;
  L.bytes = L.e + static_cast<size_t>(3) * D * 4;
  return true;
}
}  // namespace

extern "C" size_t rlsb_gru_cell_packed_bytes(int Dx, int D) {
  GruCellLayout L;
  return gru_cell_layout(Dx, D, L) ? L.bytes : 0;
}

extern "C" size_t rlsb_gru_cell_workspace_bytes(int D, int M) {
  if (D <= 0 || (D % 64) != 0 || M <= 0) return 0;
  const size_t m_tiles = (static_cast<size_t>(M) + 127) / 128;
  return static_cast<size_t>(D / 64) * m_tiles * 128 * 16;   // tagged per-block row statistics
}

extern "C" int rlsb_gru_cell_pack(const float* weight, const float* bias, const float* ln_gamma, const float* ln_beta, int Dx,
                                  int D, void* packed, void* stream) {
  GruCellLayout L;
  if (!weight || !packed || !gru_cell_layout(Dx, D, L)) return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* base = static_cast<uint8_t*>(packed);
  PackSeg segs[2] = {{0, 0, Dx}, {L.kx, Dx, D}};
  int e = launch_pack_perm(weight, Dx + D, 3 * D, reinterpret_cast<__nv_bfloat16*>(base + L.w), 192, 3 * D, L.kp, 2, segs, D, s);
  if (e == 0) e = launch_copy_gru_perm(bias, D, reinterpret_cast<float*>(base + L.b), 0.f, s);
  if (e == 0) e = launch_copy_gru_perm(ln_gamma, D, reinterpret_cast<float*>(base + L.g), 1.f, s);
  if (e == 0) e = launch_copy_gru_perm(ln_beta, D, reinterpret_cast<float*>(base + L.e), 0.f, s);
  return e;
}

extern "C" int rlsb_gru_cell_fwd(const void* packed, int Dx, int D, const void* x_packed, const void* h_packed,
                                 const float* h_prev, int M, float update_bias, float eps, float* h_next, void* h_next_packed,
                                 void* workspace, void* stream) {
  GruCellLayout L;
  if (!packed || !x_packed || !h_packed || !h_prev || !h_next || !h_next_packed || !workspace || M <= 0 ||
      !gru_cell_layout(Dx, D, L))
    return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const uint8_t* base = static_cast<const uint8_t*>(packed);
  const int m_tiles = (M + 127) / 128;
  GemmParams g{};
  g.n_seg = 2;
  g.A[0] = static_cast<const __nv_bfloat16*>(x_packed); g.a_ktiles[0] = L.kx / 64;
  g.A[1] = static_cast<const __nv_bfloat16*>(h_packed); g.a_ktiles[1] = D / 64;
  g.W = reinterpret_cast<const __nv_bfloat16*>(base + L.w);
  g.RB = 192; g.NB = D / 64; g.G = 1;
  g.M = M; g.m_tiles = m_tiles; g.N = 3 * D;
  g.bias = reinterpret_cast<const float*>(base + L.b);
  g.ln_gamma = reinterpret_cast<const float*>(base + L.g);
  g.ln_beta = reinterpret_cast<const float*>(base + L.e);
  g.ln_eps = eps;
  const size_t xbytes = static_cast<size_t>(D / 64) * m_tiles * 128 * 16;
  g.xstats = static_cast<unsigned long long*>(workspace);
  const cudaError_t ce = cudaMemsetAsync(g.xstats, 0, xbytes, s);   // all slots of a row start from the same tag
  if (ce != cudaSuccess) return static_cast<int>(ce);
  g.gru_h_prev = h_prev; g.gru_ld_h = D;
  g.gru_h_next = h_next; g.gru_ld_hn = D;
  g.gru_update_bias = update_bias;
  g.out_bf16 = static_cast<__nv_bfloat16*>(h_next_packed); g.out_kpad = D;
  return launch_gemm(g, EPI_GRU, s);
}

extern "C" int rlsb_gemm_ln_act(const void* a_packed, int k_pad, const void* w_packed, int rb,
                                const float* bias_padded, int M, int N, const float* gamma, const float* beta,
                                float eps, int act, void* out_packed, int out_kpad, void* stream) {
  if (!a_packed || !w_packed || !out_packed || (k_pad % 64) != 0) return -1;
  GemmParams g{};
  g.A[0] = static_cast<const __nv_bfloat16*>(a_packed);
  g.a_ktiles[0] = k_pad / 64;
  g.n_seg = 1;
  g.W = static_cast<const __nv_bfloat16*>(w_packed);
  g.RB = rb; g.NB = 1; g.G = 1;
  g.M = M; g.m_tiles = (M + 127) / 128; g.N = N;
  g.bias = bias_padded;
  g.ln_gamma = gamma; g.ln_beta = beta; g.ln_eps = eps; g.act = act;
  g.out_bf16 = static_cast<__nv_bfloat16*>(out_packed); g.out_kpad = out_kpad;
  return launch_gemm(g, EPI_LN_ACT, static_cast<cudaStream_t>(stream));
}

extern "C" size_t rlsb_gemm_wgrad_workspace_bytes(int n_pad, int k_pad, int M) {
  WgradParams wp{};
  wp.n_tiles = n_pad / 64; wp.G = 1; wp.m_tiles = (M + 127) / 128; wp.n_seg = 1; wp.x_ktiles[0] = k_pad / 64;
  if (n_pad <= 0 || k_pad <= 0 || (n_pad % 64) || (k_pad % 64) || M <= 0 || plan_wgrad(wp) != 0) return 0;
  return wgrad_partial_bytes(wp);
}

extern "C" int rlsb_gemm_wgrad(const void* dy_packed, int n_pad, const void* x_packed, int k_pad, int M, float* out,
                               void* workspace, void* stream) {
  if (!dy_packed || !x_packed || !out || !workspace || (n_pad % 64) || (k_pad % 64) || M <= 0) return -1;
  WgradParams wp{};
  wp.dY = static_cast<const __nv_bfloat16*>(dy_packed);
  wp.n_tiles = n_pad / 64; wp.G = 1; wp.m_tiles = (M + 127) / 128;
  wp.n_seg = 1;
  wp.X[0] = static_cast<const __nv_bfloat16*>(x_packed);
  wp.x_ktiles[0] = k_pad / 64;
  wp.x_mtile_stride[0] = static_cast<long long>(k_pad) * 128;
  wp.partial = static_cast<float*>(workspace);
  int e = plan_wgrad(wp);
  if (e != 0) return e;
  e = launch_wgrad(wp, static_cast<cudaStream_t>(stream));
  if (e != 0) return e;
  WgradReduceParams rp{};
  rp.partial = wp.partial; rp.splits = wp.splits; rp.G = 1; rp.rows_pad = wp.n_slices * 128; rp.ld = k_pad;
  rp.w_dst[0] = out; rp.n_out[0] = n_pad; rp.ld_dst = k_pad; rp.n_seg = 1;
  rp.seg[0] = PackSeg{0, 0, k_pad};
  rp.ones_col = -1;
  return launch_wgrad_reduce(rp, static_cast<cudaStream_t>(stream));
}
