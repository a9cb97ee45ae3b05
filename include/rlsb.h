/* rlsb.h — C ABI of librlsb.so: the B200-native (sm_100a) kernels behind the DreamerV2
 * imagination + lambda-return + actor-critic hot path of Midren/rl_sandbox.
 *
 * The reference has NO native/FFI layer (it is pure PyTorch), so there is no pre-existing
 * binding to match; each entry point below names the reference Python code it replaces
 * (paths relative to the reference checkout, rl_sandbox/...).  INTEGRATION.md shows the
 * ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - no allocation, no synchronisation and no exceptions inside: the caller passes
 *     workspaces, everything is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - return value: 0 = ok, < 0 = argument error, > 0 = cudaError_t;
 *   - re-entrant per stream; fp32 tensors are row-major with the reference's shapes.
 */
#ifndef RLSB_H
#define RLSB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLSB_ABI_VERSION 6

/* ---- library / device ------------------------------------------------------------------- */
int rlsb_abi_version(void);
/* 0 when the current device is sm_100 (B200); negative otherwise (the library has no fallback) */
int rlsb_check_device(void);
const char* rlsb_error_string(int code);
/* number of CUDA kernels this library has launched since load (or since the last reset != 0) */
long long rlsb_launch_count(int reset);
/* a CUDA-graph replay launches the kernels recorded at capture time without passing through the library:
 * the caller that replays a graph adds the captured launch count here so the counter stays truthful */
long long rlsb_launch_count_add(long long n);
/* tuning: CTAs per thread-block cluster sharing one weight block via TMA multicast (1, 2 or 4;
 * default 2, env RLSB_CLUSTER).  Returns the value in effect. */
int rlsb_set_cluster_size(int cs);
/* tuning: the full-row GEMM epilogues (LayerNorm / activation -> packed bf16 image) assemble each 128 x 64 output tile in
 * shared memory and write it with one bulk copy (1, default; env RLSB_STAGED) or store 16 bytes per thread (0) — same
 * bits either way.  Any other value only queries.  Returns the value in effect. */
int rlsb_set_staged_output(int on);
/* tuning: the chained rollout (rlsb_imagine_fwd) runs LayerNorm + ELU of img_in / prior1 (rssm.py:179, :192) and the whole
 * GRUCell (common.py:69-81) inside the epilogues of their contractions even where a row spans several n-blocks — the row
 * statistics of the blocks' CTAs meet in global memory — instead of writing fp32 pre-activations for a second kernel
 * (1, default; env RLSB_FUSED_RSSM; applies to flat RSSMs with D % 64 == 0 and no tape).  It changes the packed layout of the
 * GRU weight: call rlsb_imagine_pack again after a change.  Any other value only queries.  Returns the value in effect. */
int rlsb_set_fused_rssm(int on);
/* profiling: %globaltimer stamps of the fused GRU cell's epilogue ([CTAs][64 tiles][8] uint64, see GemmParams::trace in
 * csrc/rlsb_gemm.cuh; scripts/gru_cell_trace.py); NULL = off */
void rlsb_gemm_set_trace(void* device_buffer);

/* ---- K2: lambda-return + shifted-cumprod weights + advantage --------------------------------
 * replaces ImaginativeCritic._lambda_return (agents/dreamer/ac.py:52-62), the discount
 * shift+cumprod of DreamerV2.train (agents/dreamer_v2.py:192-197) and the advantage of
 * ImaginativeActor.calculate_loss (agents/dreamer/ac.py:118).
 *   r, v, d : (T, N) fp32 time-major (layout_batch_major = 0) or (N, T) (= 1), T = H + 1
 *   vs      : (H, N)     V_t = r_t + d_t * ((1-lambda) v_{t+1} + lambda V_{t+1}),  V_H = v_H
 *   w       : (T, N)     w_0 = 1, w_t = w_{t-1} * d_{t-1}                 (may be NULL)
 *   adv     : (H-1, N)   adv_t = vs_{t+1} - v_t                           (may be NULL)
 * lambda_ is a double because the reference forms (1 - lambda) in Python double arithmetic before
 * it meets the fp32 tensors; time-major results are bit-identical to the reference's fp32 loop. */
int rlsb_lambda_return_fwd(const float* r, const float* v, const float* d, int T, int64_t N,
                           double lambda_, float* vs, float* w, float* adv,
                           int layout_batch_major, void* stream);
/* gradients of sum(g_vs * vs) w.r.t. r, v, d (time-major); any output may be NULL */
int rlsb_lambda_return_bwd(const float* g_vs, const float* v, const float* d, const float* vs,
                           int T, int64_t N, double lambda_, float* g_r, float* g_v, float* g_d,
                           void* stream);

/* ---- categorical sampler (bit-exact test surface) -------------------------------------------
 * replaces OneHotCategoricalStraightThrough(...).sample() as used by Dist / DistLayer('onehot')
 * (agents/dreamer/common.py:27-28, utils/dists.py:177-179): idx = argmax_k(logits_k + g(u_k)),
 * g(u) = -log(-log u) evaluated with the bit-reproducible log of rlsb_detmath.h; ties -> lowest k.
 *   logits, uniforms : (rows, classes) fp32;  idx : (rows) int32 */
int rlsb_sample_categorical(const float* logits, const float* uniforms, int64_t rows, int classes,
                            int32_t* idx, void* stream);
/* The latent draw of K1 / K5 on its own (rssm.py:34-37: State.stoch over groups x 32 classes), as the rollout
 * launches it: screening pass with fast logs, every group whose winner is not separated from the runner-up by more
 * than the screening error is redrawn in the reference order with the bit-reproducible transform, so idx equals
 * rlsb_sample_categorical's on the same logits and uniforms, always.
 *   logits : (rows, groups*32) fp32;  uniforms : same shape, or NULL -> Philox(seed; row_offset + row, step, 0, e/4);
 *   idx : (rows, groups) uint8;  onehot_f32 : (rows, groups*32) fp32 or NULL */
int rlsb_sample_latent(const float* logits, int64_t rows, int groups, const float* uniforms, uint64_t seed,
                       uint32_t row_offset, uint32_t step, uint8_t* idx, float* onehot_f32, void* stream);
/* Philox4x32-10 uniforms exactly as the imagination kernels draw them (for RNG parity tests):
 * out[i] = uniform(seed, n = n0 + i / per_row, t, stream_id, e = i % per_row) */
int rlsb_philox_uniform(uint64_t seed, uint32_t n0, uint32_t t, uint32_t stream_id, int per_row,
                        int64_t count, float* out, void* stream);

/* ---- operand packing + tcgen05 GEMM (building blocks, exported for tests) -------------------
 * Packed layout: see rl_sandbox_b200/csrc/rlsb_ptx.cuh::packed_index (SWIZZLE_128B tile image). */
int rlsb_pack_rows(const float* src, int64_t ld_src, int rows_src, void* dst_bf16, int row_block,
                   int rows_dst_pad, int k_pad, int dst_k0, int src_c0, int len, void* stream);
/* out[M, N] (fp32, ld = ldo) = A[M, K] * W[N, K]^T + bias ; A packed (row block 128),
 * W packed with row block `rb` (n_blocks * rb >= N).  Thin wrapper over the kernel every
 * layer of the imagination path uses (replaces nn.Linear, utils/fc_nn.py:14-21). */
/* GRUCell (agents/dreamer/common.py:58-81: Linear(cat[x, h]) -> LayerNorm over the 3D pre-activations -> reset / candidate /
 * update gates -> h' = u * cand + (1 - u) * h) as ONE launch: the contraction's epilogue applies LayerNorm and gates and only
 * h' leaves the kernel (fp32 + the packed bf16 operand image).  The rollout (rlsb_imagine_fwd) uses the same kernel; this entry
 * point exposes the cell alone (parity tests against the reference module, bench.py's roofline of the dominant kernel).
 *   D % 64 == 0, 3 D > 512; weight: nn.Linear(Dx + D, 3 D).weight ([3D][Dx + D] row-major), bias / LayerNorm gain / offset [3D]
 *   x_packed / h_packed: packed bf16 images (row block 128) of x [M][Dx] (K padded to 64) and h [M][D]; h_prev fp32 [M][D]
 *   h_next fp32 [M][D]; h_next_packed: packed bf16 [M_pad x D]; workspace: rlsb_gru_cell_workspace_bytes(D, M) */
size_t rlsb_gru_cell_packed_bytes(int Dx, int D);
size_t rlsb_gru_cell_workspace_bytes(int D, int M);
int rlsb_gru_cell_pack(const float* weight, const float* bias, const float* ln_gamma, const float* ln_beta, int Dx, int D,
                       void* packed, void* stream);
int rlsb_gru_cell_fwd(const void* packed, int Dx, int D, const void* x_packed, const void* h_packed, const float* h_prev,
                      int M, float update_bias, float eps, float* h_next, void* h_next_packed, void* workspace, void* stream);
int rlsb_gemm_bias(const void* a_packed, int k_pad, const void* w_packed, int rb, int n_blocks,
                   const float* bias_padded, int M, int N, float* out, int64_t ldo, float* stats,
                   void* stream);
/* fused full-row variant: out = act(LayerNorm(A W^T + b)) written as packed bf16 [M_pad, out_kpad];
 * gamma/beta NULL => no LayerNorm; act: 0 none, 1 ELU, 2 ReLU  (fc_nn.py:14-21, rssm.py:136-152) */
int rlsb_gemm_ln_act(const void* a_packed, int k_pad, const void* w_packed, int rb,
                     const float* bias_padded, int M, int N, const float* gamma, const float* beta,
                     float eps, int act, void* out_packed, int out_kpad, void* stream);

/* out[n_pad, k_pad] (fp32 row-major) = dY[M, n_pad]^T X[M, k_pad]: the weight-gradient contraction of
 * nn.Linear's autograd; both operands are the packed row-block-128 activation images (read as MN-major
 * tensor-core operands, no transposes).  Rows >= M of the images must be zero.  Test surface of the kernel
 * rlsb_ac_update uses for every layer. */
size_t rlsb_gemm_wgrad_workspace_bytes(int n_pad, int k_pad, int M);
int rlsb_gemm_wgrad(const void* dy_packed, int n_pad, const void* x_packed, int k_pad, int M, float* out,
                    void* workspace, void* stream);

/* ---- K1: imagination rollout ----------------------------------------------------------------
 * replaces DreamerV2.imagine_trajectory (agents/dreamer_v2.py:68-96) with everything it calls:
 * ImaginativeActor.forward (agents/dreamer/ac.py:103-104), WorldModel.predict_next
 * (agents/dreamer/world_model.py:131-140), RSSM.predict_next (agents/dreamer/rssm.py:176-193),
 * GRUCell.forward (agents/dreamer/common.py:69-81), State.stoch sampling (rssm.py:34-37) and the
 * target-critic read of ImaginativeCritic.lambda_return (agents/dreamer/ac.py:65). */
typedef struct {
  int32_t D;                /* rssm_dim (deterministic state width) */
  int32_t groups;           /* latent_dim  = 32 categorical variables */
  int32_t classes;          /* latent_classes = 32 */
  int32_t A;                /* actions_num */
  int32_t hidden;           /* head MLP width (400) */
  int32_t discrete;         /* 1: one-hot categorical actor, 0: truncated-normal actor (2A outputs) */
  int32_t layer_norm;       /* config `layer_norm` (first head LN and the GRU LN exist regardless) */
  int32_t predict_discount; /* discount head present */
  int32_t with_critic;      /* also evaluate the target critic on every state */
  int32_t H;                /* horizon */
  /* torch's Bernoulli(logits).mode is (p >= 0.5) with NaN where p == 0.5 exactly
   * (world_model.py:137).  1 = reproduce that NaN (reference-exact), 0 = ties resolve to 1 —
   * a single NaN discount poisons every parameter through the losses. */
  int32_t discount_nan_on_tie;
  /* 1: the packed blob also holds the transposed weight images rlsb_imagine_bwd needs
   * (continuous actors, rho != 1: dynamics back-propagation, ac.py:121-123); requires D <= 512 */
  int32_t with_backward;
  /* slotted RSSM (agents/dreamer/rssm_slots_attention.py:166-209): slots > 1 folds the slots into the
   * row axis (row = n * slots + k, the reference's (batch, slots) order), runs `attention_blocks` mixer
   * blocks on the GRU output before the prior logits, and feeds the heads cat_k[h_k, z_k] + pos_enc.
   * 0 or 1 = the flat RSSM of rssm.py.  mixer_coeff = attention_scheduler.val (1.0 once warmed up). */
  int32_t slots;
  int32_t attention_blocks;
  int32_t symmetric_qk;
  float mixer_coeff;
  /* 1: split-operand ("bf16 x 3") contractions — every operand x travels as bf16(x) and bf16(x - bf16(x)), a Linear is
   * one tcgen05 contraction over [hi.Whi | hi.Wlo | lo.Whi] with fp32 accumulation: results within ~1e-5 of the
   * reference's fp32 path (north star: rtol 1e-3) at 3x the tensor work.  The verification mode of the parity tests;
   * flat RSSM, forward only (no tape, no actor_slots).  The packed blob depends on it: pack with the same cfg. */
  int32_t parity;
  /* 1: at step H evaluate the target-critic head only — values[H] is the bootstrap of the lambda-return, while
   * rewards[H] / discounts[H] are never read by the update (ac.py:57-58, dreamer_v2.py:192-197) and are written as 0 / 1.
   * For the training path; imagine_trajectory's public result keeps the reference's rewards[H] (flag 0). */
  int32_t last_step_value_only;
  /* rlsb_rollout_* only: CTAs per thread-block cluster (one cluster carries 128 start states): 4, 8 or 16; 0 = the
   * library default (rlsb_rollout_cluster_size()).  The packed blob depends on it: pack with the same value. */
  int32_t rollout_cluster;
} rlsb_imagine_cfg;

/* fp32 parameters in nn.Linear layout (weight = [out, in] row-major); NULL = absent.
 * mlp arrays are indexed by Linear layer 0..4 and LayerNorm 0..3 (fc_nn.py Sequential indices
 * 0,3,6,9,12 and 1,4,7,10). */
typedef struct {
  const float* w[5];
  const float* b[5];
  const float* ln_g[4];
  const float* ln_b[4];
} rlsb_mlp_params;

typedef struct {
  const float* img_in_w;  const float* img_in_b;     /* pre_determ_recurrent.0  (D, S+A)      */
  const float* img_in_ln_g; const float* img_in_ln_b;/* pre_determ_recurrent.1  (D) or NULL   */
  const float* gru_w;     const float* gru_b;        /* determ_recurrent._layer (3D, 2D)      */
  const float* gru_ln_g;  const float* gru_ln_b;     /* determ_recurrent._norm  (3D)          */
  const float* prior1_w;  const float* prior1_b;     /* ensemble_prior_estimator.0 (D, D)     */
  const float* prior1_ln_g; const float* prior1_ln_b;/* ensemble_prior_estimator.1 or NULL    */
  const float* prior2_w;  const float* prior2_b;     /* ensemble_prior_estimator.3 (S, D)     */
  rlsb_mlp_params actor;                             /* actor.actor.*                         */
  rlsb_mlp_params reward;                            /* world_model.reward_predictor.*        */
  rlsb_mlp_params discount;                          /* world_model.discount_predictor.*      */
  rlsb_mlp_params critic;                            /* critic.target_critic.*                */
  /* slotted RSSM only (NULL otherwise) */
  const float* mix_qkv_w;                            /* hidden_attention_proj.weight (3D, D), no bias */
  const float* mix_pre_norm_g; const float* mix_pre_norm_b;   /* pre_norm (D)                        */
  const float* mix_fc_w; const float* mix_fc_b;      /* fc (D, D)                             */
  const float* mix_fc_norm_g; const float* mix_fc_norm_b;     /* fc_norm (D)                          */
  const float* pos_enc;                              /* world_model.pos_enc (slots, D + S)    */
} rlsb_imagine_params;

typedef struct {
  /* explicit noise (parity mode) or NULL (Philox mode keyed by seed) */
  const float* latent_uniforms;  /* (H, N, groups*classes) */
  const float* action_noise;     /* (H, N, A): uniforms if discrete else standard normals */
  uint64_t seed;
  uint32_t row_offset;           /* global index of start state 0 of this shard */
  /* (H, N, A) actions to replay instead of sampling the actor (metrics caller of
   * imagine_trajectory(state, precomp_actions, horizon), dreamer_v2.py:83-84), or NULL */
  const float* precomp_actions;
  /* device-resident Philox key; when non-NULL it overrides `seed`, so a CUDA graph that captured the call
   * draws fresh noise on every replay (the caller rewrites the 8 bytes between replays) */
  const uint64_t* seed_device;
} rlsb_noise;

/* Slices of the rlsb_ac_update workspace that hold the ACTOR's forward activations (group 0 of the update's
 * actor | critic images): the rollout evaluates ImaginativeActor.forward (agents/dreamer/ac.py:103-104) on every
 * state anyway (agents/dreamer_v2.py:86), with the same weights and on the same packed state images the update
 * reads, so it can leave layer outputs, x_hat and 1/std where ImaginativeActor.calculate_loss's backward pass
 * (ac.py:113-146 under optimizer.py:55-57) expects them, and the update skips that half of its forward. */
typedef struct rlsb_actor_slots {
  void* x[4];        /* packed bf16 output of hidden layer l, step t at row t * m_pad: [H][m_pad x Hp] */
  void* pre[4];      /* packed bf16 x_hat (LayerNorm) or pre-activation of layer l, same geometry */
  float* rstd[4];    /* [H][m_pad] 1/sqrt(var + eps) of layer l */
  int64_t m_pad;     /* rows per step image = rlsb_packed_rows(N) */
  int32_t Hp;        /* padded hidden width */
  int32_t steps;     /* H */
} rlsb_actor_slots;

typedef struct {
  float* determ;        /* (H+1, N, D)              row 0 = start state (written by the call)   */
  float* logits;        /* (H+1, N, groups*classes)  row 0 = start logits (copied)               */
  uint8_t* stoch_idx;   /* (H+1, N, groups)          row 0 = argmax of the start one-hot         */
  float* stoch;         /* (H+1, N, groups*classes) one-hot fp32, or NULL                        */
  float* actions;       /* (H+1, N, A)               row 0 = 0                                   */
  float* rewards;       /* (H+1, N)                                                              */
  float* discounts;     /* (H+1, N)                  row 0 = 1                                   */
  float* values;        /* (H+1, N) target critic, or NULL                                       */
  float* actor_raw;     /* (H, N, A or 2A) raw actor head outputs per step, or NULL              */
  /* packed bf16 tile images of the states, one slot of rlsb_packed_rows(N) rows per step, kept for
   * rlsb_ac_update: (H+1) x [rows x round_up(D,64)] and (H+1) x [rows x round_up(groups*classes,64)]
   * (both or neither; NULL = the rollout ping-pongs inside its workspace) */
  void* determ_packed;
  void* stoch_packed;
  /* activation tape for rlsb_imagine_bwd (rlsb_imagine_tape_bytes bytes), or NULL */
  void* tape;
  /* where the actor head's activations of steps 0..H-1 are kept for rlsb_ac_update (filled by rlsb_ac_actor_slots;
   * needs determ_packed / stoch_packed, no tape, a flat RSSM), or NULL: the update then recomputes the actor forward */
  const struct rlsb_actor_slots* actor_slots;
} rlsb_imagine_out;

/* bytes needed for packed weights / activation workspace for N start states */
size_t rlsb_imagine_packed_bytes(const rlsb_imagine_cfg* cfg);
size_t rlsb_imagine_workspace_bytes(const rlsb_imagine_cfg* cfg, int64_t N);
/* fp32 nn.Linear parameters -> packed bf16 tile images + padded fp32 vectors */
int rlsb_imagine_pack(const rlsb_imagine_cfg* cfg, const rlsb_imagine_params* params, void* packed,
                      void* stream);
/* h0: (N, D) fp32; z0: (N, groups*classes) fp32 one-hot; logits0: (N, groups*classes) or NULL.
 * Slotted RSSM: every per-state tensor has slots times the rows, ordered (n, slot): h0 (N*slots, D),
 * determ (H+1, N, slots, D), logits / stoch (H+1, N, slots, S), stoch_idx (H+1, N, slots, groups),
 * latent_uniforms (H, N, slots, S); actions / rewards / discounts / values stay per start state.
 * determ_packed / stoch_packed / tape are not available with slots > 1. */
int rlsb_imagine_fwd(const rlsb_imagine_cfg* cfg, const void* packed, int64_t N, const float* h0,
                     const float* z0, const float* logits0, const rlsb_noise* noise,
                     const rlsb_imagine_out* out, void* workspace, void* stream);

/* ---- K1 as ONE persistent kernel (csrc/rlsb_rollout.cu) ------------------------------------------------------------
 * The same rollout — DreamerV2.imagine_trajectory's loop (agents/dreamer_v2.py:68-96) with everything it calls, see
 * rlsb_imagine_fwd — executed by a single kernel launch: each 128-row block of start states is carried through all H
 * steps by one thread-block cluster (rlsb_rollout_cluster_size() CTAs, env RLSB_ROLLOUT_CLUSTER = 4 | 8 | 16); layers are
 * split by output columns over the CTAs, LayerNorm statistics travel through distributed shared memory, the GRU gates
 * (agents/dreamer/common.py:69-81) are fused into their contraction's epilogue, and the only synchronisation between
 * layers is the hardware cluster barrier.  For the launch-bound sizes (the configured 16 x 50 = 800 start states); the
 * chained rlsb_imagine_fwd stays the path of the 16 k - 256 k sweep.
 * Same cfg / noise / out / workspace (rlsb_imagine_workspace_bytes) / tape contract as rlsb_imagine_fwd for the flat
 * RSSM (slots <= 1, parity == 0, out->actor_slots == NULL); the packed weights are this kernel's own. */
int rlsb_rollout_cluster_size(void);
/* clusters of `cluster` (4 | 8 | 16) CTAs of the persistent kernels that the current device keeps resident at once
 * (cudaOccupancyMaxActiveClusters: a cluster lives inside one GPC, so this is less than #SMs / cluster in general).  More row
 * blocks than this run as a second wave at twice the time: callers pick the cluster size / the chained rollout by it. */
int rlsb_rollout_max_clusters(int cluster);
/* profiling aid: a device buffer of (H + 1) x 11 x 8 uint64 that the next rlsb_rollout_fwd launches fill with
 * %globaltimer stamps of cluster 0 — per step the 11 phases head layers 0-4, read-out / action draw, img_in, GRU, prior 1,
 * prior 2, latent draw; per phase 0 producer starts, 1 first stage landed, 2 MMAs issued, 3 accumulator ready,
 * 4 statistics pass done, 5 statistics exchanged, 6 outputs stored, 7 phase finished — or NULL to switch it off */
void rlsb_rollout_set_trace(void* device_buffer);
size_t rlsb_rollout_packed_bytes(const rlsb_imagine_cfg* cfg);
int rlsb_rollout_pack(const rlsb_imagine_cfg* cfg, const rlsb_imagine_params* params, void* packed, void* stream);
int rlsb_rollout_fwd(const rlsb_imagine_cfg* cfg, const void* packed, int64_t N, const float* h0, const float* z0,
                     const float* logits0, const rlsb_noise* noise, const rlsb_imagine_out* out, void* workspace,
                     void* stream);
/* rlsb_imagine_bwd (the backward of the rollout w.r.t. the sampled actions: agents/dreamer_v2.py:199-207 through
 * agents/dreamer/rssm.py:176-193, common.py:69-81, rssm.py:34-37) as ONE persistent kernel on the same cluster machinery:
 * 12 phases per step (head gradients, four dX contractions with the ELU' / LayerNorm backward in the epilogue — row sums over
 * DSMEM —, head layer 0, straight-through softmax backward, prior MLP, GRU gate backward, the two GRU dX contractions as one
 * phase, img_in).  Same arguments and workspace (rlsb_imagine_bwd_workspace_bytes) as rlsb_imagine_bwd; `packed` is the
 * rlsb_rollout_pack blob of a cfg with with_backward = 1; `fwd->tape` may come from either forward kernel. */
int rlsb_rollout_bwd_supported(const rlsb_imagine_cfg* cfg);   /* 1: with_backward and the cluster size fits the kernel's plan */
int rlsb_rollout_bwd(const rlsb_imagine_cfg* cfg, const void* packed, int64_t N, const rlsb_imagine_out* fwd,
                     const float* g_rewards, const float* g_values, float* g_actions, void* workspace, void* stream);

/* backward of the rollout w.r.t. the sampled actions (activation gradients only: the world model and the
 * target critic receive no parameter update from the actor loss, dreamer_v2.py:199-207):
 *   g_rewards, g_values : (H+1, N) d loss / d rewards[t], d loss / d values[t]  (rlsb_lambda_return_bwd)
 *   g_actions           : (H, N, A) d loss / d a_t, a_t = the action sampled in state t (out.actions[t+1])
 * Chain per step (reference autograd of predict_next): reward / critic heads -> [h_t, z_t];
 * z_t = onehot + probs - probs.detach() -> prior logits -> prior MLP -> h_t; GRU + LayerNorm backward ->
 * (x_t, h_{t-1}); x_t -> (z_{t-1}, a_{t-1}).  The forward call must have been given out.tape. */
size_t rlsb_imagine_tape_bytes(const rlsb_imagine_cfg* cfg, int64_t N);
size_t rlsb_imagine_bwd_workspace_bytes(const rlsb_imagine_cfg* cfg, int64_t N);
int rlsb_imagine_bwd(const rlsb_imagine_cfg* cfg, const void* packed, int64_t N, const rlsb_imagine_out* fwd,
                     const float* g_rewards, const float* g_values, float* g_actions, void* workspace,
                     void* stream);

/* ---- K4: actor-critic update -------------------------------------------------------------------
 * replaces the loss half of DreamerV2.train (for a discrete actor rho == 1 and nothing differentiates
 * through the rollout, agents/dreamer/ac.py:90-92,121-125; for a continuous actor the dynamics term enters
 * through g_actions)
 * (agents/dreamer_v2.py:199-207): ImaginativeCritic.calculate_loss (agents/dreamer/ac.py:68-81),
 * ImaginativeActor.calculate_loss (ac.py:113-146) and the two loss.backward() calls of
 * Optimizer.step (utils/optimizer.py:55-57).  Outputs are the fp32 parameter gradients in nn.Linear /
 * nn.LayerNorm layout (the caller all-reduces, clips and applies AdamW) and the scalar losses / metrics.
 * The states come as the packed bf16 images rlsb_imagine_fwd leaves in determ_packed / stoch_packed;
 * steps 0..H-1 are used (critic: all H, actor: the first H-1, ac.py / dreamer_v2.py:203-206). */
typedef struct {
  int32_t D, groups, classes, A, hidden;
  int32_t discrete;        /* 1: one-hot categorical actor; 0: TruncatedNormal actor (2A outputs) */
  int32_t layer_norm;
  int32_t H;               /* imagination horizon */
  float rho;               /* reinforce fraction (1 for discrete actors) */
  float eta;               /* entropy scale */
  int32_t metrics_samples; /* draws per element behind actor/avg_val, avg_sd, min_val, max_val (ac.py:137); 0 = skip */
  int32_t actor_fwd_in_rollout; /* 1: rlsb_imagine_fwd filled the actor's slots (rlsb_imagine_out::actor_slots) */
} rlsb_ac_cfg;

typedef struct {
  float* w[5];
  float* b[5];
  float* ln_g[4];
  float* ln_b[4];          /* NULL where the MLP has no LayerNorm */
} rlsb_mlp_grads;

enum {
  RLSB_AC_LOSS_CRITIC = 0,
  RLSB_AC_LOSS_ACTOR_REINFORCE = 1,
  RLSB_AC_LOSS_ACTOR_DYNAMICS = 2,
  RLSB_AC_LOSS_ACTOR_ENTROPY = 3,
  RLSB_AC_LOSS_ACTOR = 4,
  RLSB_AC_CRITIC_AVG_TARGET = 5,
  RLSB_AC_CRITIC_AVG_LAMBDA = 6,
  RLSB_AC_CRITIC_AVG_PRED = 7,
  RLSB_AC_ACTOR_AVG_VAL = 8,
  RLSB_AC_ACTOR_MEAN_VAL = 9,
  RLSB_AC_ACTOR_AVG_SD = 10,
  RLSB_AC_ACTOR_MIN_VAL = 11,
  RLSB_AC_ACTOR_MAX_VAL = 12,
  RLSB_AC_SCALARS = 16
};

/* rows of one packed per-step image for N start states (N rounded up to the 128-row tile) */
size_t rlsb_packed_rows(int64_t N);
size_t rlsb_ac_packed_bytes(const rlsb_ac_cfg* cfg);
size_t rlsb_ac_workspace_bytes(const rlsb_ac_cfg* cfg, int64_t N);
/* actor = ImaginativeActor.actor.*, critic = ImaginativeCritic.critic.* (the trained copy, not the target) */
int rlsb_ac_pack(const rlsb_ac_cfg* cfg, const rlsb_mlp_params* actor, const rlsb_mlp_params* critic,
                 void* packed, void* stream);
/* vs: (H, N) lambda-returns; w: (H+1, N) cumprod weights; values: (H+1, N) target critic (baseline and
 * critic/avg_target_value); actions: (H+1, N, A) one-hot (row t+1 = action taken in state t);
 * g_actions: (H, N, A) d loss_actor / d a_t from rlsb_imagine_bwd when rho != 1 (continuous actor), else NULL;
 * scalars: RLSB_AC_SCALARS floats (device).  seed (or the device-resident seed_device, if non-NULL) keys the
 * Philox stream of the metric draws. */
/* The loss half of rlsb_ac_update on its own: head_out is (2, H * rlsb_packed_rows(N), 32) fp32 — group 0 the actor's
 * raw outputs (logits, or mean | std pre-activations) of states 0..H-1 in columns [0, A or 2A), group 1 the critic's
 * value in column 0; step t occupies rows [t * rlsb_packed_rows(N), ... + N).  Writes the RLSB_AC_SCALARS losses /
 * metrics of ImaginativeCritic.calculate_loss / ImaginativeActor.calculate_loss (ac.py:68-81,113-146); the dynamics
 * term is reported, gradients are not formed.  With head outputs from a rollout in the split-operand mode
 * (rlsb_imagine_cfg::parity) this is the fp32-grade evaluation of the losses. */
int rlsb_ac_losses(const rlsb_ac_cfg* cfg, int64_t N, const float* head_out, const float* vs, const float* w,
                   const float* values, const float* actions, uint64_t seed, float* scalars, void* workspace,
                   void* stream);
/* the actor's slices of `workspace` (rlsb_ac_workspace_bytes(cfg, N) bytes) for rlsb_imagine_out::actor_slots */
int rlsb_ac_actor_slots(const rlsb_ac_cfg* cfg, int64_t N, void* workspace, rlsb_actor_slots* slots);
int rlsb_ac_update(const rlsb_ac_cfg* cfg, const void* packed, int64_t N, const void* determ_packed,
                   const void* stoch_packed, const float* vs, const float* w, const float* values,
                   const float* actions, const float* g_actions, uint64_t seed, const uint64_t* seed_device,
                   const rlsb_mlp_grads* actor_grads,
                   const rlsb_mlp_grads* critic_grads, float* scalars, void* workspace, void* stream);

/* ---- K3: slot attention ---------------------------------------------------------------------
 * replaces SlotAttention.forward (rl_sandbox/vision/slot_attention.py:52-77) for explicit
 * prev_slots (the caller draws the initial slots, slot_attention.py:46-50 /
 * world_model_slots_attention.py:278-279):
 *   k, v = W_kv LN(X)                      once              (tcgen05 GEMM, bf16 k/v kept in HBM)
 *   per iteration: q = W_q LN(slots); attn = softmax_over_slots(scale q k^T) + eps;
 *                  attn /= sum_over_tokens; upd = attn v      (one CTA per frame streams k, v once)
 *                  slots = GRUCell(upd, slots); slots += W2 ReLU(W1 LN(slots) + b1) + b2
 * X: (B, tokens, dim) fp32, prev_slots: (B, slots, dim) fp32  ->  out_slots (B, slots, dim) fp32,
 * out_attn (B, slots, tokens) fp32 = the last iteration's normalised attention (may be NULL). */
typedef struct {
  int32_t slots;   /* 4  */
  int32_t dim;     /* 384 (multiple of 64, <= 512) */
  int32_t tokens;  /* 196 */
  int32_t iters;   /* slots_iter_num */
} rlsb_slot_cfg;

typedef struct {
  const float* inputs_norm_g; const float* inputs_norm_b;   /* inputs_norm   (dim)            */
  const float* inputs_proj_w;                               /* inputs_proj   (2 dim, dim)     */
  const float* slots_norm_g;  const float* slots_norm_b;    /* slots_norm    (dim)            */
  const float* slots_proj_w;                                /* slots_proj    (dim, dim)       */
  const float* gru_w_ih; const float* gru_w_hh;             /* slots_reccur  (3 dim, dim)     */
  const float* gru_b_ih; const float* gru_b_hh;             /*               (3 dim)          */
  const float* slots_norm2_g; const float* slots_norm2_b;   /* slots_norm_2  (dim)            */
  const float* mlp_w1; const float* mlp_b1;                 /* slots_proj_2.0 (4 dim, dim)    */
  const float* mlp_w2; const float* mlp_b2;                 /* slots_proj_2.2 (dim, 4 dim)    */
} rlsb_slot_params;

size_t rlsb_slot_attention_packed_bytes(const rlsb_slot_cfg* cfg);
size_t rlsb_slot_attention_workspace_bytes(const rlsb_slot_cfg* cfg, int64_t B);
int rlsb_slot_attention_pack(const rlsb_slot_cfg* cfg, const rlsb_slot_params* params, void* packed,
                             void* stream);
int rlsb_slot_attention_fwd(const rlsb_slot_cfg* cfg, const void* packed, int64_t B, const float* X,
                            const float* prev_slots, float* out_slots, float* out_attn,
                            void* workspace, void* stream);

/* ---- K5: world-model observe scan ------------------------------------------------
 * replaces the T sequential RSSM.forward calls of WorldModel.calculate_loss
 * (agents/dreamer/world_model.py:187-202 -> agents/dreamer/rssm.py:176-209: predict_next + update_current,
 * posterior sampled straight-through) and, with rlsb_observe_bwd, their autograd (BPTT over the T steps with
 * parameter gradients).  Time-major tensors: embed (T, B, E) encoder output, actions (T, B, A) already
 * multiplied by (1 - is_first) (world_model.py:191); the initial state is zero (get_initial_state).
 * The tape (rlsb_observe_tape_bytes) must be ZERO-INITIALISED by the caller. */
typedef struct {
  int32_t D, groups, classes, A;
  int32_t E;            /* encoder embedding width (4 * 384 for the conv encoder, rssm.py:156) */
  int32_t layer_norm;
  int32_t T;            /* steps (batch_cluster_size) */
} rlsb_observe_cfg;

typedef struct {
  const float* img_in_w;  const float* img_in_b;  const float* img_in_ln_g; const float* img_in_ln_b;  /* pre_determ_recurrent */
  const float* gru_w;     const float* gru_b;     const float* gru_ln_g;    const float* gru_ln_b;     /* determ_recurrent    */
  const float* prior1_w;  const float* prior1_b;  const float* prior1_ln_g; const float* prior1_ln_b;  /* ensemble_prior_estimator.0/.1 */
  const float* prior2_w;  const float* prior2_b;                                                         /* ensemble_prior_estimator.3    */
  const float* post1_w;   const float* post1_b;   const float* post1_ln_g;  const float* post1_ln_b;   /* stoch_net.0/.1 (D + E inputs) */
  const float* post2_w;   const float* post2_b;                                                          /* stoch_net.3                   */
} rlsb_observe_params;

typedef struct {
  float* img_in_w;  float* img_in_b;  float* img_in_ln_g; float* img_in_ln_b;
  float* gru_w;     float* gru_b;     float* gru_ln_g;    float* gru_ln_b;
  float* prior1_w;  float* prior1_b;  float* prior1_ln_g; float* prior1_ln_b;
  float* prior2_w;  float* prior2_b;
  float* post1_w;   float* post1_b;   float* post1_ln_g;  float* post1_ln_b;
  float* post2_w;   float* post2_b;
} rlsb_observe_grads;

typedef struct {
  float* prior_logits;  /* (T, B, groups*classes) */
  float* post_logits;   /* (T, B, groups*classes) */
  float* determ;        /* (T, B, D) */
  uint8_t* stoch_idx;   /* (T, B, groups) posterior sample */
  float* stoch;         /* (T, B, groups*classes) one-hot, or NULL */
} rlsb_observe_out;

size_t rlsb_observe_packed_bytes(const rlsb_observe_cfg* cfg);
size_t rlsb_observe_tape_bytes(const rlsb_observe_cfg* cfg, int64_t B);
size_t rlsb_observe_bwd_workspace_bytes(const rlsb_observe_cfg* cfg, int64_t B);
int rlsb_observe_pack(const rlsb_observe_cfg* cfg, const rlsb_observe_params* params, void* packed, void* stream);
/* noise: latent_uniforms (T, B, groups*classes) or the Philox key (stream 0, step t, row = row_offset + b) */
int rlsb_observe_fwd(const rlsb_observe_cfg* cfg, const void* packed, int64_t B, const float* embed, const float* actions,
                     const rlsb_noise* noise, const rlsb_observe_out* out, void* tape, void* stream);
/* g_* : d loss / d (the four forward outputs), each may be NULL; the straight-through gradient of `stoch` is routed
 * into the posterior logits inside.  grads: every parameter gradient (LayerNorm entries may be NULL without
 * layer_norm); g_embed: (T, B, E). */
int rlsb_observe_bwd(const rlsb_observe_cfg* cfg, const void* packed, int64_t B, const void* tape,
                     const rlsb_observe_out* fwd, const float* g_prior_logits, const float* g_post_logits,
                     const float* g_determ, const float* g_stoch, const rlsb_observe_grads* grads, float* g_embed,
                     void* workspace, void* stream);

/* K3 with autograd: the forward records an activation tape (rlsb_slot_attention_tape_bytes bytes); the backward
 * returns d loss / d X (B, tokens, dim), d loss / d prev_slots (B, slots, dim) and the gradient of every parameter of
 * rlsb_slot_params -- same field names, nn.Linear / nn.LayerNorm / nn.GRUCell layouts; all required -- the autograd of
 * SlotAttention.forward (vision/slot_attention.py:52-77) inside the world-model loss.  `packed` is the blob of
 * rlsb_slot_attention_pack (it also carries the transposed weight images). */
typedef struct {
  float* inputs_norm_g; float* inputs_norm_b;
  float* inputs_proj_w;
  float* slots_norm_g;  float* slots_norm_b;
  float* slots_proj_w;
  float* gru_w_ih; float* gru_w_hh;
  float* gru_b_ih; float* gru_b_hh;
  float* slots_norm2_g; float* slots_norm2_b;
  float* mlp_w1; float* mlp_b1;
  float* mlp_w2; float* mlp_b2;
} rlsb_slot_grads;

size_t rlsb_slot_attention_tape_bytes(const rlsb_slot_cfg* cfg, int64_t B);
size_t rlsb_slot_attention_bwd_workspace_bytes(const rlsb_slot_cfg* cfg, int64_t B);
int rlsb_slot_attention_fwd_tape(const rlsb_slot_cfg* cfg, const void* packed, int64_t B, const float* X,
                                 const float* prev_slots, float* out_slots, float* out_attn, void* tape,
                                 void* workspace, void* stream);
int rlsb_slot_attention_bwd(const rlsb_slot_cfg* cfg, const void* packed, int64_t B, const float* X, const void* tape,
                            const float* d_out_slots, const rlsb_slot_grads* grads, float* dX, float* d_prev_slots,
                            void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RLSB_H */
