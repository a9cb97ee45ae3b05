// rlsb_imagine.cu — K1: the imagination rollout (DreamerV2.imagine_trajectory,
// agents/dreamer_v2.py:68-96) as a stream-ordered chain of tcgen05 GEMMs with fused epilogues
// plus small HBM-bound kernels; one C-ABI call enqueues all H steps (CUDA-graph capturable).
//
// Data layout in HBM
//   * recurrent state h: fp32 row-major inside the caller's `determ` output (H+1, N, D) — the
//     GRU update h' = u*c + (1-u)*h is carried in fp32 — plus a packed bf16 image (ping-pong)
//     that feeds the tensor cores;
//   * stochastic state z: uint8 class indices (H+1, N, 32) + packed bf16 one-hot image;
//   * every bf16 operand (activations and weights) is stored as SWIZZLE_128B tile images so a
//     pipeline stage is one contiguous bulk copy (rlsb_ptx.cuh::packed_index).
#include <cstdio>
#include <cmath>
#include <cstring>

#include "../../include/rlsb.h"
#include "rlsb_count.cuh"
#include "rlsb_gemm.cuh"
#include "rlsb_imagine_plan.cuh"
#include "rlsb_kernels.cuh"

namespace rlsb {

namespace k1 {
int g_fused_rssm = [] {
  const char* e = getenv("RLSB_FUSED_RSSM");
  return (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;   // 2: LayerNorm + ELU layers only, GRU cell unfused
}();
}  // namespace k1

using namespace k1;

namespace {

int copy_pad(const float* src, int n, float* dst, int n_pad, float fill, cudaStream_t s) {
  return launch_copy_pad(src, n, dst, n_pad, fill, s);
}

// one-hot fp32 rows -> uint8 class index per group (start state only)
__global__ void onehot_to_idx_kernel(const float* __restrict__ z, long long N, int groups, int classes,
                                     uint8_t* __restrict__ idx) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= N * groups) return;
  const float* p = z + i * classes;
  int best = 0;
  float bv = p[0];
  for (int k = 1; k < classes; ++k)
    if (p[k] > bv) { bv = p[k]; best = k; }
  idx[i] = static_cast<uint8_t>(best);
}

#define RLSB_TRY(expr)            \
  do {                            \
    int _e = (expr);              \
    if (_e != 0) return _e;       \
  } while (0)

// weight image of one Linear: plain, or — Plan::parity — the split layout [Whi | Wlo | Whi] per input segment (each
// block as wide as the segment's padded width; LayerPlan::kp is already three times the padded input width)
int pack_weight(const Plan& P, const float* src, long long ld, int rows, __nv_bfloat16* dst, const LayerPlan& L, int n,
                const PackSeg* segs, cudaStream_t s) {
  if (!P.parity) return launch_pack(src, ld, rows, dst, L.RB, L.NB * L.RB, L.kp, n, segs, s);
  if (3 * n > 8) return -22;
  PackSeg ex[8];
  const int kp0 = L.kp / 3;
  for (int i = 0; i < n; ++i) {
    const int w = (i + 1 < n ? segs[i + 1].dst_k0 : kp0) - segs[i].dst_k0;
    for (int j = 0; j < 3; ++j) ex[3 * i + j] = PackSeg{3 * segs[i].dst_k0 + j * w, segs[i].src_c0, segs[i].len, j == 1 ? 1 : 0};
  }
  return launch_pack(src, ld, rows, dst, L.RB, L.NB * L.RB, L.kp, 3 * n, ex, s);
}

const rlsb_mlp_params* head_params(const rlsb_imagine_params& p, const Plan& P, int g) {
  if (g == P.g_actor) return &p.actor;
  if (g == P.g_reward) return &p.reward;
  if (g == P.g_discount) return &p.discount;
  return &p.critic;
}

}  // namespace

int launch_onehot_to_idx(const float* z, long long rows, int groups, int classes, uint8_t* idx, cudaStream_t s) {
  const long long tot = rows * groups;
  onehot_to_idx_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, s>>>(z, rows, groups, classes, idx);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

}  // namespace rlsb

using namespace rlsb;

extern "C" size_t rlsb_imagine_packed_bytes(const rlsb_imagine_cfg* cfg) {
  Plan P;
  if (!cfg || make_plan(*cfg, P) != 0) return 0;
  return P.packed_bytes;
}

extern "C" size_t rlsb_imagine_workspace_bytes(const rlsb_imagine_cfg* cfg, int64_t N) {
  Plan P;
  if (!cfg || N <= 0 || make_plan(*cfg, P) != 0) return 0;
  Workspace W;
  make_workspace(P, N, W);
  return W.bytes;
}

extern "C" size_t rlsb_imagine_tape_bytes(const rlsb_imagine_cfg* cfg, int64_t N) {
  Plan P;
  if (!cfg || N <= 0 || make_plan(*cfg, P) != 0 || !P.bwd) return 0;
  Tape T;
  make_tape(P, N, cfg->H, T);
  return T.bytes;
}

extern "C" int rlsb_imagine_pack(const rlsb_imagine_cfg* cfg, const rlsb_imagine_params* prm, void* packed,
                                 void* stream_) {
  if (!cfg || !prm || !packed) return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  Plan P;
  RLSB_TRY(make_plan(*cfg, P));
  LaunchBatchScope batch(s);   // the small pack / pad launches below are queued and issued as multi-job kernels
  uint8_t* base = static_cast<uint8_t*>(packed);
  auto wptr = [&](const LayerPlan& L) { return reinterpret_cast<__nv_bfloat16*>(base + L.w_off); };
  auto fptr = [&](size_t off) { return reinterpret_cast<float*>(base + off); };

  // ---- RSSM layers (rssm.py:136-152, common.py:58-63) ----
  {
    const LayerPlan& L = P.img_in;  // input = cat[stoch, action]
    PackSeg segs[2] = {{0, 0, P.S}, {P.Sp, P.S, P.A}};
    RLSB_TRY(pack_weight(P, prm->img_in_w, P.S + P.A, L.N, wptr(L), L, 2, segs, s));
    RLSB_TRY(copy_pad(prm->img_in_b, L.N, fptr(L.bias_off), L.NB * L.RB, 0.f, s));
    if (prm->img_in_ln_g) {
      RLSB_TRY(copy_pad(prm->img_in_ln_g, L.N, fptr(L.g_off), ru(L.N, 32), 1.f, s));
      RLSB_TRY(copy_pad(prm->img_in_ln_b, L.N, fptr(L.b_off), ru(L.N, 32), 0.f, s));
    }
  }
  {
    const LayerPlan& L = P.gru;  // input = cat[x, h]
    PackSeg segs[2] = {{0, 0, P.D}, {P.Dp, P.D, P.D}};
    if (P.gru_fused) {
      // fused cell: n-block nb = [reset | candidate | update] rows of hidden units [64 nb, 64 nb + 64), vectors likewise
      RLSB_TRY(launch_pack_perm(prm->gru_w, 2 * P.D, L.N, wptr(L), L.RB, L.NB * L.RB, L.kp, 2, segs, P.D, s));
      RLSB_TRY(launch_copy_gru_perm(prm->gru_b, P.D, fptr(L.bias_off), 0.f, s));
      RLSB_TRY(launch_copy_gru_perm(prm->gru_ln_g, P.D, fptr(L.g_off), 1.f, s));
      RLSB_TRY(launch_copy_gru_perm(prm->gru_ln_b, P.D, fptr(L.b_off), 0.f, s));
    } else {
      RLSB_TRY(pack_weight(P, prm->gru_w, 2 * P.D, L.N, wptr(L), L, 2, segs, s));
      RLSB_TRY(copy_pad(prm->gru_b, L.N, fptr(L.bias_off), L.NB * L.RB, 0.f, s));
      RLSB_TRY(copy_pad(prm->gru_ln_g, L.N, fptr(L.g_off), L.N, 1.f, s));
      RLSB_TRY(copy_pad(prm->gru_ln_b, L.N, fptr(L.b_off), L.N, 0.f, s));
    }
  }
  {
    const LayerPlan& L = P.prior1;
    PackSeg segs[1] = {{0, 0, P.D}};
    RLSB_TRY(pack_weight(P, prm->prior1_w, P.D, L.N, wptr(L), L, 1, segs, s));
    RLSB_TRY(copy_pad(prm->prior1_b, L.N, fptr(L.bias_off), L.NB * L.RB, 0.f, s));
    if (prm->prior1_ln_g) {
      RLSB_TRY(copy_pad(prm->prior1_ln_g, L.N, fptr(L.g_off), ru(L.N, 32), 1.f, s));
      RLSB_TRY(copy_pad(prm->prior1_ln_b, L.N, fptr(L.b_off), ru(L.N, 32), 0.f, s));
    }
  }
  {
    const LayerPlan& L = P.prior2;
    PackSeg segs[1] = {{0, 0, P.D}};
    RLSB_TRY(pack_weight(P, prm->prior2_w, P.D, L.N, wptr(L), L, 1, segs, s));
    RLSB_TRY(copy_pad(prm->prior2_b, L.N, fptr(L.bias_off), L.NB * L.RB, 0.f, s));
  }
  // ---- heads (fc_nn.py:4-23): groups share one launch per layer ----
  for (int l = 0; l < 5; ++l) {
    const LayerPlan& L = P.head[l];
    for (int g = 0; g < P.G; ++g) {
      const rlsb_mlp_params* hp = head_params(*prm, P, g);
      if (!hp->w[l]) return -20;
      const int n_out = (l == 4) ? ((g == P.g_actor) ? P.Aout : 1) : P.Hd;
      __nv_bfloat16* dst = wptr(L) + static_cast<size_t>(g) * L.NB * L.RB * L.kp;
      if (l == 0) {  // input = cat[determ, stoch]  (rssm.py:29-31), per slot when slotted
        PackSeg segs[8];
        for (int k = 0; k < P.K; ++k) {
          segs[2 * k] = PackSeg{k * (P.Dp + P.Sp), k * (P.D + P.S), P.D};
          segs[2 * k + 1] = PackSeg{k * (P.Dp + P.Sp) + P.Dp, k * (P.D + P.S) + P.D, P.S};
        }
        RLSB_TRY(pack_weight(P, hp->w[l], static_cast<long long>(P.K) * (P.D + P.S), n_out, dst, L, 2 * P.K, segs, s));
      } else {
        PackSeg segs[1] = {{0, 0, P.Hd}};
        RLSB_TRY(pack_weight(P, hp->w[l], P.Hd, n_out, dst, L, 1, segs, s));
      }
      if (l == 0 && P.K > 1 && prm->pos_enc) {
        // State.combined_slots adds the constant pos_enc to [h_k, z_k] (rssm_slots_attention.py:37-41):
        // W (s + p) + b = W s + (b + W p)
        RLSB_TRY(launch_bias_fold(hp->w[0], static_cast<long long>(P.K) * (P.D + P.S), n_out, P.K * (P.D + P.S),
                                  prm->pos_enc, hp->b[0], fptr(L.bias_off) + static_cast<size_t>(g) * L.NB * L.RB,
                                  L.NB * L.RB, s));
      } else {
        RLSB_TRY(copy_pad(hp->b[l], n_out, fptr(L.bias_off) + static_cast<size_t>(g) * L.NB * L.RB,
                          L.NB * L.RB, 0.f, s));
      }
      if (l < 4) {
        const int lnp = ru(L.N, 32);
        // fc_nn.py:15 — the first LayerNorm always exists; later ones only with layer_norm
        RLSB_TRY(copy_pad(hp->ln_g[l], L.N, fptr(L.g_off) + static_cast<size_t>(g) * lnp, lnp, 1.f, s));
        RLSB_TRY(copy_pad(hp->ln_b[l], L.N, fptr(L.b_off) + static_cast<size_t>(g) * lnp, lnp, 0.f, s));
      }
      if (P.bwd) {   // transposed images for the dX GEMMs of rlsb_imagine_bwd
        const TLayer& T = P.t_head[l];
        __nv_bfloat16* tb = reinterpret_cast<__nv_bfloat16*>(base + T.off);
        if (l >= 1) {
          PackSeg rs[1] = {{0, 0, P.Hd}};
          RLSB_TRY(launch_pack_transposed_seg(hp->w[l], P.Hd, n_out, tb + static_cast<size_t>(g) * T.NB * T.RB * T.kp,
                                              T.RB, T.NB * T.RB, T.kp, 0, T.kp, 1, rs, s));
        } else if (g >= P.gb0 && g < P.gb0 + P.Gb) {
          // layer 0: the gradient-carrying groups share one K axis (their dX contributions add up)
          PackSeg rs[2] = {{0, 0, P.D}, {P.Dp, P.D, P.S}};
          RLSB_TRY(launch_pack_transposed_seg(hp->w[0], P.D + P.S, P.Hd, tb, T.RB, T.NB * T.RB, T.kp,
                                              (g - P.gb0) * P.Hp, P.Hp, 2, rs, s));
        }
      }
    }
  }
  if (P.K > 1) {   // slot mixer (rssm_slots_attention.py:141-145)
    if (!prm->mix_qkv_w || !prm->mix_fc_w || !prm->mix_fc_b || !prm->mix_pre_norm_g || !prm->mix_pre_norm_b ||
        !prm->mix_fc_norm_g || !prm->mix_fc_norm_b)
      return -21;
    PackSeg seg[1] = {{0, 0, P.D}};
    RLSB_TRY(launch_pack(prm->mix_qkv_w, P.D, 3 * P.D, wptr(P.mix_qkv), P.mix_qkv.RB, P.mix_qkv.NB * P.mix_qkv.RB,
                         P.mix_qkv.kp, 1, seg, s));
    RLSB_TRY(launch_pack(prm->mix_fc_w, P.D, P.D, wptr(P.mix_fc), P.mix_fc.RB, P.mix_fc.NB * P.mix_fc.RB,
                         P.mix_fc.kp, 1, seg, s));
    RLSB_TRY(copy_pad(prm->mix_fc_b, P.D, fptr(P.mix_fc.bias_off), P.mix_fc.NB * P.mix_fc.RB, 0.f, s));
    RLSB_TRY(copy_pad(prm->mix_pre_norm_g, P.D, fptr(P.mix_pre_g), P.D, 1.f, s));
    RLSB_TRY(copy_pad(prm->mix_pre_norm_b, P.D, fptr(P.mix_pre_b), P.D, 0.f, s));
    RLSB_TRY(copy_pad(prm->mix_fc_norm_g, P.D, fptr(P.mix_fcn_g), P.D, 1.f, s));
    RLSB_TRY(copy_pad(prm->mix_fc_norm_b, P.D, fptr(P.mix_fcn_b), P.D, 0.f, s));
  }
  if (P.bwd) {
    auto tptr = [&](const TLayer& T) { return reinterpret_cast<__nv_bfloat16*>(base + T.off); };
    {
      PackSeg rs[1] = {{0, 0, P.D}};
      RLSB_TRY(launch_pack_transposed_seg(prm->prior2_w, P.D, P.S, tptr(P.t_prior2), P.t_prior2.RB,
                                          P.t_prior2.NB * P.t_prior2.RB, P.t_prior2.kp, 0, P.t_prior2.kp, 1, rs, s));
      RLSB_TRY(launch_pack_transposed_seg(prm->prior1_w, P.D, P.D, tptr(P.t_prior1), P.t_prior1.RB,
                                          P.t_prior1.NB * P.t_prior1.RB, P.t_prior1.kp, 0, P.t_prior1.kp, 1, rs, s));
      // GRU weight (3D, 2D): in-features [x | h]
      RLSB_TRY(launch_pack_transposed_seg(prm->gru_w, 2 * P.D, 3 * P.D, tptr(P.t_gru_x), P.t_gru_x.RB,
                                          P.t_gru_x.NB * P.t_gru_x.RB, P.t_gru_x.kp, 0, P.t_gru_x.kp, 1, rs, s));
      PackSeg rh[1] = {{0, P.D, P.D}};
      RLSB_TRY(launch_pack_transposed_seg(prm->gru_w, 2 * P.D, 3 * P.D, tptr(P.t_gru_h), P.t_gru_h.RB,
                                          P.t_gru_h.NB * P.t_gru_h.RB, P.t_gru_h.kp, 0, P.t_gru_h.kp, 1, rh, s));
      PackSeg ri[2] = {{0, 0, P.S}, {P.Sp, P.S, P.A}};
      RLSB_TRY(launch_pack_transposed_seg(prm->img_in_w, P.S + P.A, P.D, tptr(P.t_img_in), P.t_img_in.RB,
                                          P.t_img_in.NB * P.t_img_in.RB, P.t_img_in.kp, 0, P.t_img_in.kp, 2, ri, s));
    }
  }
  return batch.end();
}

extern "C" int rlsb_imagine_fwd(const rlsb_imagine_cfg* cfg, const void* packed, int64_t N, const float* h0,
                                const float* z0, const float* logits0, const rlsb_noise* noise,
                                const rlsb_imagine_out* out, void* workspace, void* stream_) {
  if (!cfg || !packed || !h0 || !z0 || !noise || !out || !workspace || N <= 0) return -1;
  if (!out->determ || !out->logits || !out->stoch_idx || !out->actions || !out->rewards || !out->discounts)
    return -2;
  if (N > (1LL << 30)) return -3;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  Plan P;
  RLSB_TRY(make_plan(*cfg, P));
  if (P.parity) return imagine_fwd_parity(cfg, P, packed, N, h0, z0, logits0, noise, out, workspace, s);
  Workspace W;
  make_workspace(P, N, W);
  const int M = static_cast<int>(N);          // start states: rows of the head operands
  const int m_pad = W.m_pad;
  const int m_tiles = m_pad / 128;
  const int K = P.K;                           // slots (1 = flat RSSM)
  if (N * K > (1LL << 30)) return -3;
  const int Ms = M * K;                        // rows of the RSSM operands, ordered (n, slot)
  const int ms_pad = W.ms_pad;
  const int ms_tiles = ms_pad / 128;
  const int H = cfg->H;
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  auto wbf = [&](const LayerPlan& L) { return reinterpret_cast<const __nv_bfloat16*>(pk + L.w_off); };
  auto pf = [&](size_t off) { return reinterpret_cast<const float*>(pk + off); };
  auto bf = [&](size_t off) { return reinterpret_cast<__nv_bfloat16*>(ws + off); };
  float* scratch = reinterpret_cast<float*>(ws + W.scratch);
  float* stats = reinterpret_cast<float*>(ws + W.stats);
  float* head_out = reinterpret_cast<float*>(ws + W.head_out);
  auto xstats = [&](int i) { return reinterpret_cast<unsigned long long*>(ws + W.xstats[i]); };   // img_in, GRU, prior1
  const bool ln = cfg->layer_norm != 0;
  const float eps = 1e-5f;
  const size_t ND = static_cast<size_t>(Ms) * P.D, NS = static_cast<size_t>(Ms) * P.S;
  if (K > 1 && (out->determ_packed || out->stoch_packed || out->tape)) return -6;

  // packed bf16 state images: ping-pong in the workspace, or — when the caller keeps them for the
  // actor-critic update (rlsb_ac_update) — one slot per step in the caller's buffers
  Tape TP{};
  uint8_t* tape = static_cast<uint8_t*>(out->tape);
  if (tape) {
    if (!P.bwd) return -5;
    make_tape(P, N, H, TP);
  }
  auto tp = [&](int t, size_t off) { return tape + static_cast<size_t>(t) * TP.step_bytes + off; };
  const bool keep = out->determ_packed != nullptr && out->stoch_packed != nullptr;
  // the actor head's activations of steps 0..H-1 go straight into the update's workspace (rlsb_ac_actor_slots)
  const rlsb_actor_slots* slots = out->actor_slots;
  if (slots && (tape || !keep || K > 1 || slots->steps != H || slots->m_pad != m_pad || slots->Hp != P.Hp)) return -7;
  auto slot_img = [&](void* const* base, int l, int t) {
    return static_cast<__nv_bfloat16*>(base[l]) + static_cast<size_t>(t) * m_pad * P.Hp;
  };
  if ((out->determ_packed != nullptr) != (out->stoch_packed != nullptr)) return -4;
  auto himg = [&](int t) {
    return keep ? static_cast<__nv_bfloat16*>(out->determ_packed) + static_cast<size_t>(t) * ms_pad * P.Dp
                : bf(W.hbf[t & 1]);
  };
  auto zimg = [&](int t) {
    return keep ? static_cast<__nv_bfloat16*>(out->stoch_packed) + static_cast<size_t>(t) * ms_pad * P.Sp
                : bf(W.zbf[t & 1]);
  };

  // ---- start state ---------------------------------------------------------------------------
  {
    PackSeg seg[1] = {{0, 0, P.D}};
    RLSB_TRY(launch_pack(h0, P.D, Ms, himg(0), 128, ms_pad, P.Dp, 1, seg, s));
    PackSeg segz[1] = {{0, 0, P.S}};
    RLSB_TRY(launch_pack(z0, P.S, Ms, zimg(0), 128, ms_pad, P.Sp, 1, segz, s));
    // rows >= N of the other one-hot images are never written by the sampler: clear their last M tile
    const size_t tile_row_bytes = static_cast<size_t>(P.Sp / 64) * 128 * 64 * 2;
    cudaError_t e = cudaSuccess;
    if (Ms != ms_pad) {
      for (int t = 1; t <= (keep ? H : 1) && e == cudaSuccess; ++t)
        e = cudaMemsetAsync(reinterpret_cast<uint8_t*>(zimg(t)) + static_cast<size_t>(ms_tiles - 1) * tile_row_bytes, 0,
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
    // tagged statistics of the cross-block LayerNorm: all slots of a row must start from the same tag
    if (g_fused_rssm) e = cudaMemsetAsync(ws + W.xstats[0], 0, W.xstats_bytes, s);
    if (e != cudaSuccess) return static_cast<int>(e);
    RLSB_TRY(launch_onehot_to_idx(z0, Ms, cfg->groups, cfg->classes, out->stoch_idx, s));
  }

  // heads work on M start states; RSSM layers on Ms = M * slots rows
  auto base_gemm = [&](const LayerPlan& L, bool rssm_rows = true) {
    GemmParams g{};
    g.W = wbf(L); g.RB = L.RB; g.NB = L.NB; g.G = L.G;
    g.M = rssm_rows ? Ms : M; g.m_tiles = rssm_rows ? ms_tiles : m_tiles; g.N = L.N;
    g.bias = pf(L.bias_off);
    g.ln_eps = eps;
    return g;
  };

  // Linear -> [LN] -> ELU -> packed bf16, for a single-group RSSM layer
  auto rssm_layer = [&](const LayerPlan& L, GemmParams g, bool has_ln, __nv_bfloat16* outp, uint8_t* save_pre,
                        uint8_t* save_rstd) -> int {
    if (L.fullrow) {
      g.ln_gamma = has_ln ? pf(L.g_off) : nullptr;
      g.ln_beta = has_ln ? pf(L.b_off) : nullptr;
      g.act = ACT_ELU;
      g.out_bf16 = outp; g.out_kpad = P.Dp; g.out_bf16_group_stride = 0;
      g.save_pre = reinterpret_cast<__nv_bfloat16*>(save_pre);
      g.save_rstd = has_ln ? reinterpret_cast<float*>(save_rstd) : nullptr;
      return launch_gemm(g, EPI_LN_ACT, s);
    }
    if (!save_pre && ln_layer_fused(P, L)) {
      // the row spans NB n-blocks: LayerNorm statistics meet across the blocks' CTAs (GemmParams::xstats)
      g.ln_gamma = has_ln ? pf(L.g_off) : nullptr;
      g.ln_beta = has_ln ? pf(L.b_off) : nullptr;
      g.act = ACT_ELU;
      g.out_bf16 = outp; g.out_kpad = P.Dp; g.out_bf16_group_stride = 0;
      g.xstats = xstats(&L == &P.img_in ? 0 : 2);
      return launch_gemm(g, EPI_LN_ACT, s);
    }
    g.out_f32 = scratch; g.ldo = W.ld_scratch; g.out_group_stride = 0; g.stats = stats;
    RLSB_TRY(launch_gemm(g, has_ln ? EPI_STATS : EPI_PLAIN, s));
    return launch_ln_act(scratch, W.ld_scratch, stats, L.NB, L.RB, Ms, ms_pad, L.N,
                         has_ln ? pf(L.g_off) : nullptr, has_ln ? pf(L.b_off) : nullptr, eps, ACT_ELU,
                         outp, P.Dp, s);
  };

  for (int t = 0; t <= H; ++t) {
    const __nv_bfloat16* hb = himg(t);
    const __nv_bfloat16* zb = zimg(t);
    // ---- heads on s_t = cat[h_t, z_t]: actor, reward, discount, target critic -------------------
    // Training callers (cfg.last_step_value_only) read rows 0..H-1 of rewards / discounts (ac.py:57-58,
    // dreamer_v2.py:192-197) and only the bootstrap value of state H: at t == H the target-critic group runs alone
    // (a quarter of that step's head work); rewards[H] := 0, discounts[H] := 1.
    const bool crit_only = cfg->last_step_value_only != 0 && t == H && !tape && K == 1 && P.g_critic >= 0;
    const int g0 = crit_only ? P.g_critic : 0;
    for (int l = 0; l < 5; ++l) {
      const LayerPlan& L = P.head[l];
      GemmParams g = base_gemm(L, false);
      if (crit_only) {   // group g0 only: every per-group base pointer moves to that group's slot
        g.G = 1;
        g.W = wbf(L) + static_cast<size_t>(g0) * L.NB * L.RB * L.kp;
        g.bias = pf(L.bias_off) + static_cast<size_t>(g0) * L.NB * L.RB;
      }
      const size_t hid_g0 = static_cast<size_t>(g0) * m_pad * P.Hp;
      if (l == 0 && K > 1) {
        // slotted State.combined (rssm_slots_attention.py:33-43): cat over slots of [h_k, z_k]; the (n, slot)-ordered
        // images are first gathered into one operand plane per slot (pos_enc is folded into the bias)
        RLSB_TRY(launch_slot_gather(hb, M, K, P.Dp, bf(W.hplanes), m_pad, s));
        RLSB_TRY(launch_slot_gather(zb, M, K, P.Sp, bf(W.zplanes), m_pad, s));
        g.n_seg = 2 * K;
        for (int k = 0; k < K; ++k) {
          g.A[2 * k] = bf(W.hplanes) + static_cast<size_t>(k) * m_pad * P.Dp;
          g.a_ktiles[2 * k] = P.Dp / 64; g.a_group_stride[2 * k] = 0;
          g.A[2 * k + 1] = bf(W.zplanes) + static_cast<size_t>(k) * m_pad * P.Sp;
          g.a_ktiles[2 * k + 1] = P.Sp / 64; g.a_group_stride[2 * k + 1] = 0;
        }
      } else if (l == 0) {
        g.n_seg = 2;
        g.A[0] = hb; g.a_ktiles[0] = P.Dp / 64; g.a_group_stride[0] = 0;
        g.A[1] = zb; g.a_ktiles[1] = P.Sp / 64; g.a_group_stride[1] = 0;
      } else {
        g.n_seg = 1;
        g.A[0] = bf(W.hid[(l - 1) & 1]) + hid_g0; g.a_ktiles[0] = P.Hp / 64;
        g.a_group_stride[0] = static_cast<long long>(m_pad) * P.Hp;
      }
      if (l < 4) {
        const bool has_ln = (l == 0) || ln;
        const size_t ln_g0 = static_cast<size_t>(g0) * ru(L.N, 32);
        g.ln_gamma = has_ln ? pf(L.g_off) + ln_g0 : nullptr;
        g.ln_beta = has_ln ? pf(L.b_off) + ln_g0 : nullptr;
        g.act = ACT_ELU;
        g.out_bf16 = bf(W.hid[l & 1]) + hid_g0; g.out_kpad = P.Hp;
        g.out_bf16_group_stride = static_cast<long long>(m_pad) * P.Hp;
        if (tape) {
          g.save_pre = reinterpret_cast<__nv_bfloat16*>(tp(t, TP.head_pre[l]));
          g.save_rstd = has_ln ? reinterpret_cast<float*>(tp(t, TP.head_rstd[l])) : nullptr;
        }
        if (slots && t < H) {   // the actor group reads / writes its slot of the update's images and keeps x_hat, rstd
          g.alt_group_p1 = P.g_actor + 1;
          g.alt_A = l > 0 ? slot_img(slots->x, l - 1, t) : nullptr;
          g.alt_out_bf16 = slot_img(slots->x, l, t);
          g.save_pre = slot_img(slots->pre, l, t);
          g.save_rstd = has_ln ? slots->rstd[l] + static_cast<size_t>(t) * m_pad : nullptr;
        }
        RLSB_TRY(launch_gemm(g, EPI_LN_ACT, s));
      } else {
        g.out_f32 = head_out + static_cast<size_t>(g0) * m_pad * 32; g.ldo = 32;
        g.out_group_stride = static_cast<long long>(m_pad) * 32;
        if (slots && t < H) {
          g.alt_group_p1 = P.g_actor + 1;
          g.alt_A = slot_img(slots->x, 3, t);
        }
        RLSB_TRY(launch_gemm(g, EPI_PLAIN, s));
      }
    }
    HeadFinishParams hf{};
    hf.head_out = head_out; hf.ldo = 32; hf.group_stride = static_cast<long long>(m_pad) * 32;
    hf.g_actor = P.g_actor; hf.g_reward = P.g_reward; hf.g_discount = P.g_discount; hf.g_critic = P.g_critic;
    if (crit_only) hf.g_actor = hf.g_reward = hf.g_discount = -1;   // not evaluated: reward := 0, discount := 1
    hf.M = M; hf.m_pad = m_pad; hf.A = P.A; hf.discrete = cfg->discrete;
    hf.first_step = (t == 0); hf.want_action = (t < H); hf.nan_on_tie = cfg->discount_nan_on_tie;
    hf.noise.explicit_noise = noise->action_noise ? noise->action_noise + static_cast<size_t>(t) * N * P.A : nullptr;
    hf.noise.ld = P.A; hf.noise.seed = noise->seed; hf.noise.seed_ptr = noise->seed_device; hf.noise.step = static_cast<uint32_t>(t);
    hf.noise.row_offset = noise->row_offset;
    hf.reward_out = out->rewards + static_cast<size_t>(t) * N;
    hf.discount_out = out->discounts + static_cast<size_t>(t) * N;
    hf.value_out = out->values ? out->values + static_cast<size_t>(t) * N : nullptr;
    hf.action_out = (t < H) ? out->actions + static_cast<size_t>(t + 1) * N * P.A : nullptr;
    hf.actor_raw_out = (out->actor_raw && t < H) ? out->actor_raw + static_cast<size_t>(t) * N * P.Aout : nullptr;
    hf.precomp = (noise->precomp_actions && t < H) ? noise->precomp_actions + static_cast<size_t>(t) * N * P.A : nullptr;
    hf.action_packed = bf(W.abf); hf.a_kpad = P.Ap;
    hf.action_repeat = K; hf.action_rows_pad = ms_pad;
    RLSB_TRY(launch_head_finish(hf, s));
    if (t == H) break;

    // ---- x = ELU(LN?(W_in [z, a] + b))                                   rssm.py:179 ----------
    {
      GemmParams g = base_gemm(P.img_in);
      g.n_seg = 2;
      g.A[0] = zb; g.a_ktiles[0] = P.Sp / 64;
      g.A[1] = bf(W.abf); g.a_ktiles[1] = P.Ap / 64;
      RLSB_TRY(rssm_layer(P.img_in, g, ln, bf(W.xbf), tape ? tp(t + 1, TP.x_pre) : nullptr,
                          tape ? tp(t + 1, TP.x_rstd) : nullptr));
    }
    // ---- h' = GRU(x, h)                                    rssm.py:181, common.py:69-81 ----------
    {
      GemmParams g = base_gemm(P.gru);
      g.n_seg = 2;
      g.A[0] = bf(W.xbf); g.a_ktiles[0] = P.Dp / 64;
      g.A[1] = hb; g.a_ktiles[1] = P.Dp / 64;
      if (P.gru_fused) {
        // the whole cell in the contraction's epilogue: only h' leaves the kernel (fp32 state + packed bf16 operand image)
        g.ln_gamma = pf(P.gru.g_off); g.ln_beta = pf(P.gru.b_off);
        g.xstats = xstats(1);
        g.gru_h_prev = out->determ + static_cast<size_t>(t) * ND; g.gru_ld_h = P.D;
        g.gru_h_next = out->determ + static_cast<size_t>(t + 1) * ND; g.gru_ld_hn = P.D;
        g.gru_update_bias = -1.0f;
        g.out_bf16 = himg(t + 1); g.out_kpad = P.Dp;
        RLSB_TRY(launch_gemm(g, EPI_GRU, s));
      } else {
      // with a tape the pre-LayerNorm gate activations of this transition are kept (slot t+1)
      float* gsc = tape ? reinterpret_cast<float*>(tp(t + 1, TP.gru_scratch)) : scratch;
      float* gst = tape ? reinterpret_cast<float*>(tp(t + 1, TP.gru_stats)) : stats;
      g.out_f32 = gsc; g.ldo = W.ld_scratch; g.stats = gst;
      RLSB_TRY(launch_gemm(g, EPI_STATS, s));
      RLSB_TRY(launch_gru_gate(gsc, W.ld_scratch, gst, P.gru.NB, P.gru.RB, Ms, ms_pad, P.D,
                               pf(P.gru.g_off), pf(P.gru.b_off), eps, -1.0f,
                               out->determ + static_cast<size_t>(t) * ND, P.D,
                               out->determ + static_cast<size_t>(t + 1) * ND, P.D, himg(t + 1), P.Dp, s));
      }
    }
    // ---- prior logits = W2 ELU(LN?(W1 h' + b1)) + b2                      rssm.py:192 ----------
    const __nv_bfloat16* prior_in = himg(t + 1);
    if (K > 1) {
      // ---- slot mixer (rssm_slots_attention.py:186-203): determ_post = h'; per block
      //      q,k,v = W_qkv LN(determ_post); attn over slots; determ_post += W_fc LN(attn v) + b_fc.
      //      Only the prior logits see determ_post; the state keeps the un-mixed h' (:207).
      float* hpost = reinterpret_cast<float*>(ws + W.hpost);
      float* qkv = reinterpret_cast<float*>(ws + W.mix_qkv);
      float* fco = reinterpret_cast<float*>(ws + W.mix_fc);
      cudaError_t e = cudaMemcpyAsync(hpost, out->determ + static_cast<size_t>(t + 1) * ND, ND * 4,
                                      cudaMemcpyDeviceToDevice, s);
      if (e != cudaSuccess) return static_cast<int>(e);
      for (int b = 0; b <= P.nblk; ++b) {
        const bool last = b == P.nblk;
        // determ_post += fc output of the previous block; then pre_norm (or, after the last block, a plain pack)
        RLSB_TRY(launch_residual_ln_pack(hpost, P.D, b > 0 ? fco : nullptr, P.D, Ms, ms_pad, P.D,
                                         last ? nullptr : pf(P.mix_pre_g), last ? nullptr : pf(P.mix_pre_b), eps,
                                         bf(W.mix_ln), P.Dp, s));
        if (last) break;
        GemmParams gq = base_gemm(P.mix_qkv);
        gq.bias = nullptr;
        gq.n_seg = 1;
        gq.A[0] = bf(W.mix_ln); gq.a_ktiles[0] = P.Dp / 64;
        gq.out_f32 = qkv; gq.ldo = W.ld_qkv;
        RLSB_TRY(launch_gemm(gq, EPI_PLAIN, s));
        RLSB_TRY(launch_mixer_attn(qkv, W.ld_qkv, M, K, P.D, cfg->symmetric_qk, 1.0f / sqrtf(static_cast<float>(P.D)),
                                   1e-8f, cfg->mixer_coeff, pf(P.mix_fcn_g), pf(P.mix_fcn_b), eps, bf(W.mix_upd),
                                   P.Dp, ms_pad, s));
        GemmParams gf = base_gemm(P.mix_fc);
        gf.n_seg = 1;
        gf.A[0] = bf(W.mix_upd); gf.a_ktiles[0] = P.Dp / 64;
        gf.out_f32 = fco; gf.ldo = P.D;
        RLSB_TRY(launch_gemm(gf, EPI_PLAIN, s));
      }
      prior_in = bf(W.mix_ln);
    }
    {
      GemmParams g = base_gemm(P.prior1);
      g.n_seg = 1;
      g.A[0] = prior_in; g.a_ktiles[0] = P.Dp / 64;
      RLSB_TRY(rssm_layer(P.prior1, g, ln, bf(W.ybf), tape ? tp(t + 1, TP.y_pre) : nullptr,
                          tape ? tp(t + 1, TP.y_rstd) : nullptr));
      GemmParams g2 = base_gemm(P.prior2);
      g2.n_seg = 1;
      g2.A[0] = bf(W.ybf); g2.a_ktiles[0] = P.Dp / 64;
      g2.out_f32 = out->logits + static_cast<size_t>(t + 1) * NS; g2.ldo = P.S;
      RLSB_TRY(launch_gemm(g2, EPI_PLAIN, s));
    }
    // ---- z' ~ OneHotCategoricalST(logits)                                rssm.py:34-37 ----------
    {
      NoiseSpec ns{};
      ns.explicit_noise = noise->latent_uniforms ? noise->latent_uniforms + static_cast<size_t>(t) * NS : nullptr;
      ns.ld = P.S; ns.seed = noise->seed; ns.seed_ptr = noise->seed_device; ns.step = static_cast<uint32_t>(t);
      ns.row_offset = noise->row_offset * static_cast<uint32_t>(K);   // rows are (global start state, slot)
      RLSB_TRY(launch_sample_latent(out->logits + static_cast<size_t>(t + 1) * NS, P.S, Ms, cfg->groups,
                                    cfg->classes, ns, out->stoch_idx + static_cast<size_t>(t + 1) * Ms * cfg->groups,
                                    zimg(t + 1), P.Sp,
                                    out->stoch ? out->stoch + static_cast<size_t>(t + 1) * NS : nullptr, P.S, s));
    }
  }
  return 0;
}
