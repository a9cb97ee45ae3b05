// rlsb_imagine_plan.cuh — layout of the packed weight blob, the activation workspace and the
// activation tape of K1 (shared by the forward rollout, rlsb_imagine.cu, and its backward pass,
// rlsb_imagine_bwd.cu).
#pragma once
#include <cstddef>
#include <cstdint>

#include <cuda_runtime.h>

#include "../../include/rlsb.h"

namespace rlsb {
namespace k1 {

inline int ru(int x, int m) { return (x + m - 1) / m * m; }
inline size_t rus(size_t x, size_t m) { return (x + m - 1) / m * m; }

struct LayerPlan {
  int N = 0;       // valid outputs per group
  int RB = 0, NB = 0, G = 1;
  int kp = 0;      // padded K (sum of segments)
  size_t w_off = 0;     // bytes into packed blob (bf16 tiles)
  size_t bias_off = 0;  // fp32 [G][NB*RB]
  size_t g_off = 0, b_off = 0;  // fp32 LN params: [G][RB] (full-row) or [N] (stats path)
  bool fullrow = false;
};

inline void plan_nb(LayerPlan& L) {
  if (ru(L.N, 32) <= 512) {
    L.NB = 1;
    L.RB = ru(L.N, 32);
    L.fullrow = true;
  } else {
    L.NB = (L.N + 255) / 256;
    L.RB = ru((L.N + L.NB - 1) / L.NB, 32);
    L.fullrow = false;
  }
}

// Fused RSSM epilogues of the chained rollout (rlsb_set_fused_rssm, default on): LayerNorm + ELU of img_in / prior1 and the
// whole GRU cell run inside their contractions' epilogues even when a row spans several n-blocks (GemmParams::xstats);
// 0 = contraction -> fp32 pre-activations -> ln_act_kernel / gru_gate_kernel.  Read by make_plan: re-pack after a change.
extern int g_fused_rssm;

// transposed weight image for a backward (dX) GEMM: rows = in-features (padded), K = out-features
struct TLayer {
  int RB = 0, NB = 0, kp = 0;
  size_t off = 0;
};

struct Plan {
  int D, S, A, Hd, Dp, Sp, Ap, Hp, Aout, G;
  int g_actor, g_reward, g_discount, g_critic;
  LayerPlan img_in, gru, prior1, prior2, head[5];
  // ---- slotted RSSM (cfg.slots > 1): mixer blocks between the GRU and the prior logits ----
  int K = 1, nblk = 0;
  LayerPlan mix_qkv, mix_fc;
  size_t mix_pre_g = 0, mix_pre_b = 0, mix_fcn_g = 0, mix_fcn_b = 0;   // fp32 [D]
  // ---- split-operand contraction mode (cfg.parity): every weight image carries [Whi | Wlo | Whi] per input segment,
  //      i.e. LayerPlan::kp is three times the padded input width (rlsb_imagine_parity.cu)
  bool parity = false;
  // ---- GRU cell fused into its contraction (EPI_GRU): P.gru has NB = D / 64 blocks of RB = 192 permuted rows
  bool gru_fused = false;
  // ---- backward (cfg.with_backward): dX operands ----
  bool bwd = false;
  int Gb = 0, gb0 = 0;       // head groups that carry gradient: gb0 .. gb0+Gb-1 (reward .. target critic)
  int G3p = 0;               // 3D padded to 64 (K of the GRU dX GEMMs)
  TLayer t_head[5];          // l = 1..4: [G][RB=ru(Hd,32)][kp]; l = 0: K-concatenated groups, rows = Dp+Sp
  TLayer t_prior2, t_prior1, t_gru_x, t_gru_h, t_img_in;
  size_t packed_bytes;
};

// a LayerNorm / ELU layer of the RSSM runs with its epilogue fused (one block per row, or the cross-block exchange)
inline bool ln_layer_fused(const Plan& P, const LayerPlan& L) {
  return L.fullrow || (g_fused_rssm != 0 && (L.RB % 64) == 0 && L.NB * L.RB == P.Dp && L.N == P.Dp);
}

inline size_t place(size_t& cursor, size_t bytes) {
  cursor = rus(cursor, 1024);
  size_t off = cursor;
  cursor += bytes;
  return off;
}

inline int make_plan(const rlsb_imagine_cfg& c, Plan& P) {
  if (c.classes != 32 || c.groups <= 0 || c.groups > 64) return -10;
  if (c.D <= 0 || c.A <= 0 || c.hidden <= 0 || c.H <= 0) return -11;
  P.D = c.D; P.S = c.groups * c.classes; P.A = c.A; P.Hd = c.hidden;
  P.Dp = ru(P.D, 64); P.Sp = ru(P.S, 64); P.Ap = ru(P.A, 64); P.Hp = ru(P.Hd, 64);
  P.Aout = c.discrete ? c.A : 2 * c.A;
  if (P.Aout > 32 || P.Ap > 64) return -12;
  if (ru(P.Hd, 32) > 512) return -13;
  P.K = c.slots > 1 ? c.slots : 1;
  P.nblk = P.K > 1 ? c.attention_blocks : 0;
  if (P.K > 1 && (P.K > 4 || P.D > 512 || c.with_backward || P.nblk < 0)) return -16;
  P.parity = c.parity != 0;
  if (P.parity && (P.K > 1 || c.with_backward)) return -17;   // verification mode: flat RSSM, forward only
  const int kmul = P.parity ? 3 : 1;
  int g = 0;
  P.g_actor = g++;
  P.g_reward = g++;
  P.g_discount = c.predict_discount ? g++ : -1;
  P.g_critic = c.with_critic ? g++ : -1;
  P.G = g;

  size_t cur = 0;
  auto finish = [&](LayerPlan& L, int ln_len_per_group) {
    plan_nb(L);
    L.w_off = place(cur, static_cast<size_t>(L.G) * L.NB * L.RB * L.kp * 2);
    L.bias_off = place(cur, static_cast<size_t>(L.G) * L.NB * L.RB * 4);
    L.g_off = place(cur, static_cast<size_t>(L.G) * ln_len_per_group * 4);
    L.b_off = place(cur, static_cast<size_t>(L.G) * ln_len_per_group * 4);
  };
  P.img_in.N = P.D; P.img_in.kp = kmul * (P.Sp + P.Ap); finish(P.img_in, ru(P.D, 32));
  P.gru.N = 3 * P.D; P.gru.kp = kmul * 2 * P.Dp; finish(P.gru, 3 * P.D);
  // (finish() placed NB * RB >= 3 D rows; the fused layout has exactly 3 D of them)
  P.gru_fused = g_fused_rssm == 1 && !c.with_backward && !P.parity && P.K == 1 && (P.D % 64) == 0 && !P.gru.fullrow;
  if (P.gru_fused) {
    P.gru.NB = P.D / 64;
    P.gru.RB = 192;
  }
  P.prior1.N = P.D; P.prior1.kp = kmul * P.Dp;   finish(P.prior1, ru(P.D, 32));
  P.prior2.N = P.S; P.prior2.kp = kmul * P.Dp;   finish(P.prior2, 32);
  for (int l = 0; l < 5; ++l) {
    LayerPlan& L = P.head[l];
    L.G = P.G;
    L.N = (l == 4) ? P.Aout : P.Hd;
    L.kp = kmul * ((l == 0) ? P.K * (P.Dp + P.Sp) : P.Hp);
    finish(L, ru(L.N, 32));
  }
  if (P.K > 1) {
    P.mix_qkv.N = 3 * P.D; P.mix_qkv.kp = P.Dp; finish(P.mix_qkv, 32);
    P.mix_fc.N = P.D; P.mix_fc.kp = P.Dp;       finish(P.mix_fc, 32);
    P.mix_pre_g = place(cur, static_cast<size_t>(P.D) * 4);
    P.mix_pre_b = place(cur, static_cast<size_t>(P.D) * 4);
    P.mix_fcn_g = place(cur, static_cast<size_t>(P.D) * 4);
    P.mix_fcn_b = place(cur, static_cast<size_t>(P.D) * 4);
  }
  P.bwd = c.with_backward != 0;
  if (P.bwd) {
    if (!P.img_in.fullrow || !P.prior1.fullrow || P.g_critic < 0) return -15;   // backward: D <= 512 and a critic
    P.gb0 = P.g_reward;
    P.Gb = P.g_critic - P.g_reward + 1;
    P.G3p = ru(3 * P.D, 64);
    auto tplace = [&](TLayer& T, int rows, int kp, int groups) {
      LayerPlan tmp;
      tmp.N = rows;
      plan_nb(tmp);
      T.RB = tmp.RB; T.NB = tmp.NB; T.kp = kp;
      T.off = place(cur, static_cast<size_t>(groups) * T.NB * T.RB * kp * 2);
    };
    for (int l = 1; l < 5; ++l) tplace(P.t_head[l], P.Hd, ru(P.head[l].N, 64), P.G);
    tplace(P.t_head[0], P.Dp + P.Sp, P.Gb * P.Hp, 1);
    tplace(P.t_prior2, P.D, P.Sp, 1);
    tplace(P.t_prior1, P.D, P.Dp, 1);
    tplace(P.t_gru_x, P.D, P.G3p, 1);
    tplace(P.t_gru_h, P.D, P.G3p, 1);
    tplace(P.t_img_in, P.Sp + P.Ap, P.Dp, 1);
  }
  P.packed_bytes = rus(cur, 1024);
  return 0;
}

// activation tape written by the forward rollout when rlsb_imagine_out::tape != NULL
struct Tape {
  size_t head_pre[4], head_rstd[4], x_pre, x_rstd, gru_scratch, gru_stats, y_pre, y_rstd;  // offsets inside a step
  size_t step_bytes;
  long long ld_scratch;
  int m_pad;
  size_t bytes;
};

inline void make_tape(const Plan& P, long long N, int H, Tape& T) {
  const size_t m_pad = static_cast<size_t>(ru(static_cast<int>(N), 128));
  T.m_pad = static_cast<int>(m_pad);
  size_t cur = 0;
  for (int l = 0; l < 4; ++l) {
    T.head_pre[l] = place(cur, static_cast<size_t>(P.G) * m_pad * P.Hp * 2);
    T.head_rstd[l] = place(cur, static_cast<size_t>(P.G) * m_pad * 4);
  }
  T.x_pre = place(cur, m_pad * P.Dp * 2);
  T.x_rstd = place(cur, m_pad * 4);
  T.ld_scratch = ru(3 * P.D, 4);
  T.gru_scratch = place(cur, m_pad * T.ld_scratch * 4);
  T.gru_stats = place(cur, static_cast<size_t>(P.gru.NB) * m_pad * 2 * 4);
  T.y_pre = place(cur, m_pad * P.Dp * 2);
  T.y_rstd = place(cur, m_pad * 4);
  T.step_bytes = rus(cur, 1024);
  T.bytes = T.step_bytes * static_cast<size_t>(H + 1);
}

struct Workspace {
  size_t hbf[2], zbf[2], abf, xbf, ybf, hid[2], scratch, stats, head_out;
  size_t xstats[3];     // tagged row statistics of the cross-block LayerNorm (GemmParams::xstats), one array per layer:
  size_t xstats_bytes;  // img_in, GRU, prior1 ([NB][ms_pad][2] uint64 each; contiguous, cleared once per rollout)
  // slotted RSSM: per-slot operand planes for the heads and the mixer's buffers
  size_t hplanes, zplanes, hpost, mix_ln, mix_qkv, mix_upd, mix_fc;
  // split-operand mode (Plan::parity): fp32 pre-activation buffers, the residual ("lo") images and a zero image
  size_t p_pre = 0, p_head_pre = 0, p_hlo[2] = {0, 0}, p_zero = 0, p_zero_bytes = 0, p_hid_lo[2] = {0, 0}, p_alo = 0,
         p_xlo = 0, p_ylo = 0;
  long long p_ld = 0, p_ld_head = 0;
  long long ld_scratch, ld_qkv;
  int m_pad;    // rows of the head operands (start states, padded)
  int ms_pad;   // rows of the RSSM operands (start states x slots, padded)
  size_t bytes;
};

inline void make_workspace(const Plan& P, long long N, Workspace& W) {
  const int m_pad = ru(static_cast<int>(N), 128);
  const int ms_pad = ru(static_cast<int>(N) * P.K, 128);
  W.m_pad = m_pad;
  W.ms_pad = ms_pad;
  size_t cur = 0;
  for (int i = 0; i < 2; ++i) W.hbf[i] = place(cur, static_cast<size_t>(ms_pad) * P.Dp * 2);
  for (int i = 0; i < 2; ++i) W.zbf[i] = place(cur, static_cast<size_t>(ms_pad) * P.Sp * 2);
  W.abf = place(cur, static_cast<size_t>(ms_pad) * P.Ap * 2);
  W.xbf = place(cur, static_cast<size_t>(ms_pad) * P.Dp * 2);
  W.ybf = place(cur, static_cast<size_t>(ms_pad) * P.Dp * 2);
  for (int i = 0; i < 2; ++i) W.hid[i] = place(cur, static_cast<size_t>(P.G) * m_pad * P.Hp * 2);
  W.ld_scratch = ru(3 * P.D, 4);
  // fp32 pre-activations of the unfused stages (3 D floats per row): not needed when every RSSM layer is fused
  const bool need_scratch = P.K > 1 || !P.gru_fused || !ln_layer_fused(P, P.img_in) || !ln_layer_fused(P, P.prior1);
  W.scratch = place(cur, need_scratch ? static_cast<size_t>(ms_pad) * W.ld_scratch * 4 : 0);
  int nbmax = P.gru.NB;
  if (P.img_in.NB > nbmax) nbmax = P.img_in.NB;
  W.stats = place(cur, static_cast<size_t>(nbmax) * ms_pad * 2 * 4);
  W.head_out = place(cur, static_cast<size_t>(P.G) * m_pad * 32 * 4);
  {
    const LayerPlan* xl[3] = {&P.img_in, &P.gru, &P.prior1};
    size_t tot = 0;
    for (int i = 0; i < 3; ++i) tot += rus(static_cast<size_t>(xl[i]->NB) * ms_pad * 16, 1024);
    size_t off = place(cur, tot);
    W.xstats_bytes = tot;
    for (int i = 0; i < 3; ++i) {
      W.xstats[i] = off;
      off += rus(static_cast<size_t>(xl[i]->NB) * ms_pad * 16, 1024);
    }
  }
  W.ld_qkv = ru(3 * P.D, 4);
  W.hplanes = W.zplanes = W.hpost = W.mix_ln = W.mix_qkv = W.mix_upd = W.mix_fc = 0;
  if (P.K > 1) {
    W.hplanes = place(cur, static_cast<size_t>(P.K) * m_pad * P.Dp * 2);
    W.zplanes = place(cur, static_cast<size_t>(P.K) * m_pad * P.Sp * 2);
    W.hpost = place(cur, static_cast<size_t>(ms_pad) * P.D * 4);
    W.mix_ln = place(cur, static_cast<size_t>(ms_pad) * P.Dp * 2);
    W.mix_qkv = place(cur, static_cast<size_t>(ms_pad) * W.ld_qkv * 4);
    W.mix_upd = place(cur, static_cast<size_t>(ms_pad) * P.Dp * 2);
    W.mix_fc = place(cur, static_cast<size_t>(ms_pad) * P.D * 4);
  }
  if (P.parity) {
    const size_t mp = static_cast<size_t>(m_pad);
    W.p_ld = ru(3 * P.D, 4);
    W.p_ld_head = ru(P.Hd, 4);
    W.p_pre = place(cur, mp * W.p_ld * 4);
    W.p_head_pre = place(cur, static_cast<size_t>(P.G) * mp * W.p_ld_head * 4);
    for (int i = 0; i < 2; ++i) W.p_hlo[i] = place(cur, mp * P.Dp * 2);
    W.p_zero_bytes = mp * P.Sp * 2;
    W.p_zero = place(cur, W.p_zero_bytes);
    for (int i = 0; i < 2; ++i) W.p_hid_lo[i] = place(cur, static_cast<size_t>(P.G) * mp * P.Hp * 2);
    W.p_alo = place(cur, mp * P.Ap * 2);
    W.p_xlo = place(cur, mp * P.Dp * 2);
    W.p_ylo = place(cur, mp * P.Dp * 2);
  }
  W.bytes = rus(cur, 1024);
}



// workspace of the backward rollout (rlsb_imagine_bwd / rlsb_rollout_bwd)
struct BwdWorkspace {
  size_t dy4, dh[2], g_s, g_logits, dp1, g_hprior, g_pre, g_hdirect, dp_in, g_hgru, g_za;
  int m_pad;
  long long ldS, ldZA;
  size_t bytes;
};

inline void make_bwd_workspace(const Plan& P, long long N, BwdWorkspace& W) {
  const size_t m_pad = static_cast<size_t>(ru(static_cast<int>(N), 128));
  W.m_pad = static_cast<int>(m_pad);
  W.ldS = P.Dp + P.Sp;
  W.ldZA = P.Sp + P.Ap;
  size_t cur = 0;
  W.dy4 = place(cur, static_cast<size_t>(P.Gb) * m_pad * 64 * 2);
  for (int i = 0; i < 2; ++i) W.dh[i] = place(cur, static_cast<size_t>(P.Gb) * m_pad * P.Hp * 2);
  W.g_s = place(cur, m_pad * W.ldS * 4);
  W.g_logits = place(cur, m_pad * P.Sp * 2);
  W.dp1 = place(cur, m_pad * P.Dp * 2);
  W.g_hprior = place(cur, m_pad * P.D * 4);
  W.g_pre = place(cur, m_pad * P.G3p * 2);
  W.g_hdirect = place(cur, m_pad * P.D * 4);
  W.dp_in = place(cur, m_pad * P.Dp * 2);
  W.g_hgru = place(cur, m_pad * P.D * 4);
  W.g_za = place(cur, m_pad * W.ldZA * 4);
  W.bytes = rus(cur, 1024);
}

}  // namespace k1

// start-state one-hot rows -> uint8 class indices (rlsb_imagine.cu)
int launch_onehot_to_idx(const float* z, long long rows, int groups, int classes, uint8_t* idx, cudaStream_t s);
// the rollout in the split-operand contraction mode (rlsb_imagine_parity.cu)
int imagine_fwd_parity(const rlsb_imagine_cfg* cfg, const k1::Plan& P, const void* packed, int64_t N, const float* h0,
                       const float* z0, const float* logits0, const rlsb_noise* noise, const rlsb_imagine_out* out,
                       void* workspace, cudaStream_t s);
}  // namespace rlsb
