// rlsb_imagine_bwd.cu — backward pass of the imagination rollout (K1) w.r.t. the sampled actions.
//
// Needed when the actor is continuous (rho != 1): loss_actor contains -mean((1 - rho) * vs * w)
// (agents/dreamer/ac.py:121-123) and the lambda-returns depend on the actions through
//   a_t -> x_{t+1} = ELU(LN?(W_in [z_t, a_t]))          rssm.py:179
//       -> h_{t+1} = GRU(x_{t+1}, h_t)                   common.py:69-81
//       -> prior logits -> z_{t+1} = onehot + p - p.detach()   rssm.py:34-37,192
//       -> reward head / target critic on [h, z]          world_model.py:135, ac.py:65
// Only ACTIVATION gradients are propagated: the actor loss never updates the world model or the
// target critic (their .grad is discarded by the next zero_grad), so no weight gradient is formed
// here; the actor's own parameter gradients follow from g_actions in rlsb_ac_update.
//
// Per step (t = H .. 1) the chain is 4 + 1 + 1 + 1 + 2 + 1 tcgen05 GEMMs against transposed weight
// images (EPI_BWD fuses ELU' and the LayerNorm backward into the dX epilogue) plus three small
// HBM-bound kernels (head gradient packing, straight-through softmax backward, GRU gate backward).
#include "../../include/rlsb.h"
#include "rlsb_count.cuh"
#include "rlsb_gemm.cuh"
#include "rlsb_imagine_plan.cuh"
#include "rlsb_kernels.cuh"
#include "rlsb_ptx.cuh"
#include "rlsb_bwd_rowops.cuh"

namespace rlsb {

using namespace k1;

namespace {

using bwdops::GruBwdArgs;

// d loss / d (head outputs) of the reward head and the target critic -> packed [Gb][m_pad x 64]
__global__ void head_grad_kernel(const float* __restrict__ g_r, const float* __restrict__ g_v, int M, int m_pad,
                                 int Gb, int gb_reward, int gb_critic, __nv_bfloat16* __restrict__ dy4) {
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= m_pad) return;
  bwdops::head_grad_row(m, g_r, g_v, M, m_pad, Gb, gb_reward, gb_critic, dy4);
}

// z = onehot + p - p.detach(), p = softmax(logits) over each group of 32 classes (rssm.py:34-37):
// g_logit_j = p_j (g_z_j - sum_k g_z_k p_k).  One thread per (row, group); g_z = ga (+ gb).
__global__ void st_softmax_bwd_kernel(const float* __restrict__ logits, long long ld_l, const float* __restrict__ ga,
                                      long long ld_a, const float* __restrict__ gb, long long ld_b, int M, int groups,
                                      __nv_bfloat16* __restrict__ out, int kpad) {
  pdl_launch_dependents();
  pdl_wait();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(M) * groups) return;
  const int m = static_cast<int>(i / groups);
  const int g = static_cast<int>(i - static_cast<long long>(m) * groups);
  bwdops::st_softmax_bwd_item<false>(m, g, logits, ld_l, ga, ld_a, gb, ld_b, out, kpad);
}

// one warp per row, lane -> chunks of 8 consecutive j (D % 8 == 0)
__global__ void gru_gate_bwd_kernel(const GruBwdArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  if (warp >= a.m_pad) return;
  bwdops::gru_gate_bwd_row<false>(a, static_cast<int>(warp), lane);
}

__global__ void extract_cols_kernel(const float* __restrict__ src, long long ld, int col0, int M, int n,
                                    float* __restrict__ dst) {
  pdl_launch_dependents();
  pdl_wait();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(M) * n) return;
  const int m = static_cast<int>(i / n), k = static_cast<int>(i % n);
  dst[i] = src[static_cast<size_t>(m) * ld + col0 + k];
}

#define RLSB_TRY(expr)            \
  do {                            \
    int _e = (expr);              \
    if (_e != 0) return _e;       \
  } while (0)
#define RLSB_CUDA_OK()                                  \
  do {                                                  \
    cudaError_t _ce = cudaGetLastError();               \
    if (_ce != cudaSuccess) return static_cast<int>(_ce); \
  } while (0)

}  // namespace

}  // namespace rlsb

using namespace rlsb;

extern "C" size_t rlsb_imagine_bwd_workspace_bytes(const rlsb_imagine_cfg* cfg, int64_t N) {
  Plan P;
  if (!cfg || N <= 0 || make_plan(*cfg, P) != 0 || !P.bwd) return 0;
  BwdWorkspace W;
  make_bwd_workspace(P, N, W);
  return W.bytes;
}

extern "C" int rlsb_imagine_bwd(const rlsb_imagine_cfg* cfg, const void* packed, int64_t N,
                                const rlsb_imagine_out* fwd, const float* g_rewards, const float* g_values,
                                float* g_actions, void* workspace, void* stream_) {
  if (!cfg || !packed || !fwd || !g_rewards || !g_values || !g_actions || !workspace || N <= 0) return -1;
  if (!fwd->tape || !fwd->determ || !fwd->logits) return -2;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  Plan P;
  RLSB_TRY(make_plan(*cfg, P));
  if (!P.bwd || (P.D & 7) != 0) return -15;
  const int H = cfg->H;
  Tape TP;
  make_tape(P, N, H, TP);
  BwdWorkspace W;
  make_bwd_workspace(P, N, W);
  const int M = static_cast<int>(N);
  const int m_pad = W.m_pad;
  const int m_tiles = m_pad / 128;
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const uint8_t* tape = static_cast<const uint8_t*>(fwd->tape);
  auto tp = [&](int t, size_t off) { return tape + static_cast<size_t>(t) * TP.step_bytes + off; };
  auto bf = [&](size_t off) { return reinterpret_cast<__nv_bfloat16*>(ws + off); };
  auto f32 = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
  auto pbf = [&](size_t off) { return reinterpret_cast<const __nv_bfloat16*>(pk + off); };
  auto pf = [&](size_t off) { return reinterpret_cast<const float*>(pk + off); };
  const bool ln = cfg->layer_norm != 0;
  const float eps = 1e-5f;
  const size_t ND = static_cast<size_t>(N) * P.D, NS = static_cast<size_t>(N) * P.S;
  const long long hid_gs = static_cast<long long>(m_pad) * P.Hp;
  const int lnp = ru(P.Hd, 32);

  auto base = [&]() {
    GemmParams g{};
    g.M = M; g.m_tiles = m_tiles;
    g.ln_eps = eps;
    g.NB = 1; g.G = 1;
    g.n_seg = 1;
    return g;
  };

  for (int t = H; t >= 1; --t) {
    // ---- reward head + target critic at state t: d loss / d [h_t, z_t] -> g_s -------------------------
    if (launch_pdl(head_grad_kernel, static_cast<unsigned>((m_pad + 127) / 128), 128, 0, s, g_rewards + static_cast<size_t>(t) * N,
                                                          g_values + static_cast<size_t>(t) * N, M, m_pad, P.Gb,
                                                          P.g_reward - P.gb0, P.g_critic - P.gb0, bf(W.dy4)) != cudaSuccess) return static_cast<int>(cudaGetLastError());
    count_launch();
    RLSB_CUDA_OK();
    const __nv_bfloat16* dy = bf(W.dy4);
    long long dy_gs = static_cast<long long>(m_pad) * 64;
    for (int l = 4; l >= 1; --l) {
      const TLayer& T = P.t_head[l];
      const bool has_ln = (l - 1 == 0) || ln;
      GemmParams g = base();
      g.G = P.Gb;
      g.A[0] = dy; g.a_ktiles[0] = T.kp / 64; g.a_group_stride[0] = dy_gs;
      g.W = pbf(T.off) + static_cast<size_t>(P.gb0) * T.NB * T.RB * T.kp;
      g.RB = T.RB; g.N = P.Hd;
      g.ln_gamma = has_ln ? pf(P.head[l - 1].g_off) + static_cast<size_t>(P.gb0) * lnp : nullptr;
      g.ln_beta = has_ln ? pf(P.head[l - 1].b_off) + static_cast<size_t>(P.gb0) * lnp : nullptr;
      g.act = ACT_ELU;
      g.bwd_pre = reinterpret_cast<const __nv_bfloat16*>(tp(t, TP.head_pre[l - 1])) + static_cast<size_t>(P.gb0) * hid_gs;
      g.bwd_rstd = has_ln ? reinterpret_cast<const float*>(tp(t, TP.head_rstd[l - 1])) + static_cast<size_t>(P.gb0) * m_pad
                          : nullptr;
      g.out_bf16 = bf(W.dh[l & 1]); g.out_kpad = P.Hp; g.out_bf16_group_stride = hid_gs;
      g.group_major = 1;
      RLSB_TRY(launch_gemm(g, EPI_BWD, s));
      dy = bf(W.dh[l & 1]);
      dy_gs = hid_gs;
    }
    {
      const TLayer& T = P.t_head[0];   // layer 0: the groups are K segments of one contraction
      GemmParams g = base();
      g.n_seg = P.Gb;
      for (int i = 0; i < P.Gb; ++i) {
        g.A[i] = dy + static_cast<size_t>(i) * hid_gs;
        g.a_ktiles[i] = P.Hp / 64;
        g.a_group_stride[i] = 0;
      }
      g.W = pbf(T.off); g.RB = T.RB; g.NB = T.NB; g.N = P.Dp + P.Sp;
      g.out_f32 = f32(W.g_s); g.ldo = W.ldS;
      RLSB_TRY(launch_gemm(g, EPI_PLAIN, s));
    }
    // ---- z_t -> prior logits -> prior MLP -> h_t ---------------------------------------------------------
    {
      const long long tot = static_cast<long long>(M) * cfg->groups;
      if (launch_pdl(st_softmax_bwd_kernel, static_cast<unsigned>(static_cast<unsigned>((tot + 127) / 128)), 128, 0, s, 
          fwd->logits + static_cast<size_t>(t) * NS, P.S, f32(W.g_s) + P.Dp, W.ldS,
          t < H ? f32(W.g_za) : nullptr, W.ldZA, M, cfg->groups, bf(W.g_logits), P.Sp) != cudaSuccess) return static_cast<int>(cudaGetLastError());
      count_launch();
      RLSB_CUDA_OK();
      GemmParams g = base();
      g.A[0] = bf(W.g_logits); g.a_ktiles[0] = P.Sp / 64;
      g.W = pbf(P.t_prior2.off); g.RB = P.t_prior2.RB; g.N = P.D;
      g.ln_gamma = ln ? pf(P.prior1.g_off) : nullptr;
      g.ln_beta = ln ? pf(P.prior1.b_off) : nullptr;
      g.act = ACT_ELU;
      g.bwd_pre = reinterpret_cast<const __nv_bfloat16*>(tp(t, TP.y_pre));
      g.bwd_rstd = ln ? reinterpret_cast<const float*>(tp(t, TP.y_rstd)) : nullptr;
      g.out_bf16 = bf(W.dp1); g.out_kpad = P.Dp;
      g.group_major = 1;
      RLSB_TRY(launch_gemm(g, EPI_BWD, s));
      GemmParams g1 = base();
      g1.A[0] = bf(W.dp1); g1.a_ktiles[0] = P.Dp / 64;
      g1.W = pbf(P.t_prior1.off); g1.RB = P.t_prior1.RB; g1.NB = P.t_prior1.NB; g1.N = P.D;
      g1.out_f32 = f32(W.g_hprior); g1.ldo = P.D;
      RLSB_TRY(launch_gemm(g1, EPI_PLAIN, s));
    }
    // ---- h_t = GRU(x_t, h_{t-1}): gates + joint LayerNorm backward ---------------------------------------
    {
      GruBwdArgs a{};
      a.scratch = reinterpret_cast<const float*>(tp(t, TP.gru_scratch)); a.ld = TP.ld_scratch;
      a.stats = reinterpret_cast<const float*>(tp(t, TP.gru_stats));
      a.NB = P.gru.NB; a.M = M; a.m_pad = m_pad; a.D = P.D;
      a.gamma = pf(P.gru.g_off); a.beta = pf(P.gru.b_off);
      a.eps = eps; a.update_bias = -1.0f;
      a.h_prev = fwd->determ + static_cast<size_t>(t - 1) * ND; a.ld_h = P.D;
      a.gh[0] = f32(W.g_s); a.ld_gh[0] = W.ldS;
      a.gh[1] = f32(W.g_hprior); a.ld_gh[1] = P.D;
      a.n_gh = 2;
      if (t < H) {
        a.gh[2] = f32(W.g_hdirect); a.ld_gh[2] = P.D;
        a.gh[3] = f32(W.g_hgru); a.ld_gh[3] = P.D;
        a.n_gh = 4;
      }
      a.g_pre = bf(W.g_pre); a.kpad = P.G3p;
      a.g_hdirect = f32(W.g_hdirect);
      if (launch_pdl(gru_gate_bwd_kernel, static_cast<unsigned>((m_pad * 32 + 255) / 256), 256, 0, s, a) != cudaSuccess) return static_cast<int>(cudaGetLastError());
      count_launch();
      RLSB_CUDA_OK();
      // d loss / d x_t with the ELU' / LayerNorm backward of the img_in layer fused
      GemmParams g = base();
      g.A[0] = bf(W.g_pre); g.a_ktiles[0] = P.G3p / 64;
      g.W = pbf(P.t_gru_x.off); g.RB = P.t_gru_x.RB; g.N = P.D;
      g.ln_gamma = ln ? pf(P.img_in.g_off) : nullptr;
      g.ln_beta = ln ? pf(P.img_in.b_off) : nullptr;
      g.act = ACT_ELU;
      g.bwd_pre = reinterpret_cast<const __nv_bfloat16*>(tp(t, TP.x_pre));
      g.bwd_rstd = ln ? reinterpret_cast<const float*>(tp(t, TP.x_rstd)) : nullptr;
      g.out_bf16 = bf(W.dp_in); g.out_kpad = P.Dp;
      g.group_major = 1;
      RLSB_TRY(launch_gemm(g, EPI_BWD, s));
      // d loss / d h_{t-1} through the gates
      GemmParams gh = base();
      gh.A[0] = bf(W.g_pre); gh.a_ktiles[0] = P.G3p / 64;
      gh.W = pbf(P.t_gru_h.off); gh.RB = P.t_gru_h.RB; gh.NB = P.t_gru_h.NB; gh.N = P.D;
      gh.out_f32 = f32(W.g_hgru); gh.ldo = P.D;
      RLSB_TRY(launch_gemm(gh, EPI_PLAIN, s));
    }
    // ---- x_t = ELU(LN?(W_in [z_{t-1}, a_{t-1}])): d loss / d z_{t-1}, d loss / d a_{t-1} -----------------
    {
      GemmParams g = base();
      g.A[0] = bf(W.dp_in); g.a_ktiles[0] = P.Dp / 64;
      g.W = pbf(P.t_img_in.off); g.RB = P.t_img_in.RB; g.NB = P.t_img_in.NB; g.N = P.Sp + P.Ap;
      g.out_f32 = f32(W.g_za); g.ldo = W.ldZA;
      RLSB_TRY(launch_gemm(g, EPI_PLAIN, s));
      const long long tot = static_cast<long long>(M) * P.A;
      if (launch_pdl(extract_cols_kernel, static_cast<unsigned>(static_cast<unsigned>((tot + 255) / 256)), 256, 0, s, 
          f32(W.g_za), W.ldZA, P.Sp, M, P.A, g_actions + static_cast<size_t>(t - 1) * N * P.A) != cudaSuccess) return static_cast<int>(cudaGetLastError());
      count_launch();
      RLSB_CUDA_OK();
    }
  }
  return 0;
}
