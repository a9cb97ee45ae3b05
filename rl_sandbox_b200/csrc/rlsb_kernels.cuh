// rlsb_kernels.cuh — launchers for the non-GEMM kernels (packing, LayerNorm/gates, sampling,
// lambda-return scan, slot attention).  All launchers return cudaError_t as int.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace rlsb {

struct PackSeg {
  int dst_k0;  // first destination column (inside the padded K axis)
  int src_c0;  // first source column
  int len;     // columns
  int part = 0;  // 0: bf16(x); 1: the residual bf16(x - float(bf16(x))) — the "lo" image of the split-operand
                 // (bf16 x 3) contraction mode, rlsb_imagine_cfg::parity (launch_pack only)
};

// fp32 row-major [rows_src x *] (leading dim ld_src) -> packed bf16 [rows_dst_pad x k_pad] with
// row block RB; everything not covered by a segment (and rows >= rows_src) is written as zero.
// Batching of the small launches of a weight (re-)pack: between batch_begin(stream) and batch_end() every launch_pack and
// launch_copy_pad on that stream is queued instead of launched; batch_end() issues them as a handful of multi-job kernels
// (the job table travels in the kernel parameters).  The queued jobs must be independent of each other and of anything
// launched in between — they are: each writes its own region of a packed blob from fp32 parameters.  A pack is ~140 tiny
// launches per optimizer step otherwise, a third of all launches of the step at 800 start states.
void batch_begin(cudaStream_t stream);
int batch_end();
// dst[i] = i < n ? src[i] : fill for i < n_pad (src may be nullptr: all fill)
int launch_copy_pad(const float* src, int n, float* dst, int n_pad, float fill, cudaStream_t stream);
// scope guard: an early error return inside a pack function must not leave the thread's batch open
struct LaunchBatchScope {
  explicit LaunchBatchScope(cudaStream_t stream) { batch_begin(stream); }
  ~LaunchBatchScope() { if (!done_) batch_end(); }
  int end() { done_ = true; return batch_end(); }
 private:
  bool done_ = false;
};

int launch_pack(const float* src, long long ld_src, int rows_src, __nv_bfloat16* dst, int RB,
                int rows_dst_pad, int k_pad, int n_seg, const PackSeg* segs, cudaStream_t stream);
// launch_pack with the GRU weight's row permutation for the fused cell (EPI_GRU, rlsb_gemm.cuh): gru_perm_D = D > 0 takes
// destination row 192 nb + 64 gate + u from source row gate * D + 64 nb + u (RB must be 192, rows 3 D); 0 = launch_pack
int launch_pack_perm(const float* src, long long ld_src, int rows_src, __nv_bfloat16* dst, int RB, int rows_dst_pad, int k_pad,
                     int n_seg, const PackSeg* segs, int gru_perm_D, cudaStream_t stream);
// the same permutation for the GRU's fp32 vectors of length 3 D (bias, LayerNorm gain / offset); src == nullptr: all `fill`
int launch_copy_gru_perm(const float* src, int D, float* dst, float fill, cudaStream_t stream);

// transposed variant: logical T[r][c] = src[c * ld_src + r] for r < rows, c < cols (nn.Linear weight
// [cols = out, rows = in] -> operand whose rows are the in-features), zero padded
int launch_pack_transposed(const float* src, long long ld_src, int cols, int rows, __nv_bfloat16* dst, int RB,
                           int rows_dst_pad, int k_pad, cudaStream_t stream);

// general form: the rows of the transposed operand are gathered from in-feature segments
// (seg.dst_k0 = first padded row, seg.src_c0 = first in-feature, seg.len) and only the K range
// [dst_k0, dst_k0 + k_len) is written (several Linears can share one K axis: K-concatenated groups)
int launch_pack_transposed_seg(const float* src, long long ld_src, int cols, __nv_bfloat16* dst, int RB,
                               int rows_dst_pad, int k_pad, int dst_k0, int k_len, int n_seg, const PackSeg* segs,
                               cudaStream_t stream);

// scratch fp32 [M_pad x ld] + per-block (mean, M2) partials -> [LayerNorm] -> act -> packed bf16
int launch_ln_act(const float* scratch, long long ld, const float* stats, int NB, int RB, int M,
                  int m_pad, int N, const float* gamma, const float* beta, float eps, int act,
                  __nv_bfloat16* out, int out_kpad, cudaStream_t stream);

// GRU gate update (reference: common.py:69-81).  scratch = W[x,h]+b over 3D columns
// (reset | cand | update), LayerNorm over all 3D jointly, then
//   r = sigmoid(p_r); c = tanh(r * p_c); u = sigmoid(p_u + update_bias); h' = u*c + (1-u)*h
int launch_gru_gate(const float* scratch, long long ld, const float* stats, int NB, int RB, int M,
                    int m_pad, int D, const float* gamma, const float* beta, float eps,
                    float update_bias, const float* h_prev, long long ld_h, float* h_next,
                    long long ld_hn, __nv_bfloat16* h_next_packed, int kpad, cudaStream_t stream);

struct NoiseSpec {
  const float* explicit_noise;  // uniforms (categorical) / normals (continuous) or nullptr
  long long ld;                 // row stride of the explicit tensor (elements)
  uint64_t seed;                // Philox key when explicit_noise == nullptr
  const uint64_t* seed_ptr;     // device-resident key overriding `seed` (CUDA-graph replays change it without re-capture)
  uint32_t step;                // imagination step (Philox counter word)
  uint32_t row_offset;          // global index of local row 0 (multi-GPU sharding)
};

// 32x32 (groups x classes) straight-through categorical draw by Gumbel-max with the
// deterministic transform of rlsb_detmath.h (reference: rssm.py:34-37, dists.py:177-179).
int launch_sample_latent(const float* logits, long long ld, int M, int groups, int classes,
                         NoiseSpec noise, uint8_t* idx_out, __nv_bfloat16* onehot_packed, int kpad,
                         float* onehot_f32, long long ld_f32, cudaStream_t stream);

// standalone categorical sampler on arbitrary (rows x groups x classes<=64) logits: indices only
int launch_sample_categorical(const float* logits, const float* uniforms, long long rows, int classes,
                              int32_t* idx_out, cudaStream_t stream);

struct HeadFinishParams {
  const float* head_out;  // [G][m_pad][ldo]
  long long ldo, group_stride;
  int g_actor, g_reward, g_discount, g_critic;  // group ids, -1 = absent
  int M, m_pad, A;
  int discrete;
  int first_step;  // t == 0: discount := 1
  int nan_on_tie;  // Bernoulli.mode NaN at p == 0.5 (reference-exact) vs tie -> 1
  int want_action; // t < H
  NoiseSpec noise;
  float* reward_out;    // [M]
  float* discount_out;  // [M]
  float* value_out;     // [M]
  float* action_out;    // [M][A]   (actions[t+1])
  float* actor_raw_out; // [M][A or 2A] raw actor head output (logits / mean,std pre-activations) or nullptr
  const float* precomp; // [M][A] action to replay instead of sampling, or nullptr
  __nv_bfloat16* action_packed;  // [action_rows_pad x a_kpad]
  __nv_bfloat16* action_packed_lo;  // residual image of the action (split-operand mode) or nullptr
  int a_kpad;
  int action_repeat;             // slots (each action row is written `action_repeat` times), 0/1 = flat
  int action_rows_pad;           // rows of the action image
};
int launch_head_finish(const HeadFinishParams& p, cudaStream_t stream);

// K2 (reference: ac.py:52-66 + dreamer_v2.py:192-197 + ac.py:118)
// ---- split-operand ("bf16 x 3") contraction mode: rlsb_imagine_cfg::parity ------------------------------------------
// Every activation x is carried as two packed bf16 images, hi = bf16(x) and lo = bf16(x - hi); with the weights split the
// same way, x.w = hi.Whi + hi.Wlo + lo.Whi + O(2^-17 |x||w|) — three K segments of the tcgen05 contraction, fp32
// accumulation, i.e. fp32-grade results from the bf16 tensor-core path.
// fp32 pre-activations [G][m_pad x ld] (bias already added) -> [LayerNorm over N columns] -> act -> hi / lo images
int launch_ln_act_split(const float* pre, long long ld, long long group_stride, int G, int M, int m_pad, int N,
                        const float* gamma, const float* beta, int ln_group_stride, float eps, int act,
                        __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int out_kpad, long long out_group_stride,
                        cudaStream_t stream);
// GRU gates (common.py:69-81) with its own two-pass LayerNorm statistics and libm-grade sigmoid / tanh
int launch_gru_gate_split(const float* pre, long long ld, int M, int m_pad, int D, const float* gamma, const float* beta,
                          float eps, float update_bias, const float* h_prev, long long ld_h, float* h_next, long long ld_hn,
                          __nv_bfloat16* h_hi, __nv_bfloat16* h_lo, int kpad, cudaStream_t stream);

int launch_lambda_return(const float* r, const float* v, const float* d, int T, long long N,
                         double lambda_, float* vs, float* w, float* adv, int layout_batch_major,
                         cudaStream_t stream);
int launch_lambda_return_bwd(const float* g_vs, const float* v, const float* d, const float* vs,
                             int T, long long N, double lambda_, float* g_r, float* g_v, float* g_d,
                             cudaStream_t stream);

}  // namespace rlsb

namespace rlsb {
// ---- slotted RSSM (rssm_slots_attention.py:166-209), D <= 512, slots <= 4 -------------------------------
// x[m][:] += add[m][:] (add may be nullptr), then packed = LayerNorm(x) * gamma + beta (gamma nullptr: packed = x)
int launch_residual_ln_pack(float* x, long long ld, const float* add, long long ld_add, int M, int m_pad, int D,
                            const float* gamma, const float* beta, float eps, __nv_bfloat16* out, int kpad,
                            cudaStream_t stream);
// one warp per start state n: q_i, k_j, v_j = rows n*K+{i,j} of qkv (q | k | v, each D wide);
// attn = softmax_j(scale q_i.k_j) + eps, renormalised, blended with the identity by `coeff`;
// upd_i = sum_j attn_ij v_j; packed row n*K+i = LayerNorm(upd_i) * gamma + beta
int launch_mixer_attn(const float* qkv, long long ld, int N, int K, int D, int symmetric_qk, float scale,
                      float attn_eps, float coeff, const float* gamma, const float* beta, float ln_eps,
                      __nv_bfloat16* out, int kpad, int rows_pad, cudaStream_t stream);
// packed slot-minor image (row n*K+k) -> K plane images [K][m_pad x kpad] (row n), padding rows zeroed
int launch_slot_gather(const __nv_bfloat16* src, int N, int K, int kpad, __nv_bfloat16* dst, int m_pad,
                       cudaStream_t stream);
// bias'[j] = bias[j] + sum_i W[j][i] * pos[i]  (the constant pos_enc added to the head input folds into the bias)
int launch_bias_fold(const float* W, long long ld, int n_out, int n_in, const float* pos, const float* bias,
                     float* out, int out_pad, cudaStream_t stream);
}  // namespace rlsb
