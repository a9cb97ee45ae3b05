// rlsb_mixer.cu — small kernels of the slotted RSSM step (agents/dreamer/rssm_slots_attention.py:166-209):
// the slot-mixing attention blocks that run between the GRU and the prior logits, the slot gather that
// turns the (n, slot)-ordered state images into per-slot operand planes for the heads, and the fold of
// the constant positional encoding into the first head layer's bias.
#include "rlsb_kernels.cuh"

#include "rlsb_count.cuh"
#include "rlsb_gemm.cuh"
#include "rlsb_ptx.cuh"

namespace rlsb {

namespace {

constexpr int kMaxPerLane = 16;   // D <= 512

__device__ __forceinline__ uint16_t bf16_bits(float x) {
  __nv_bfloat16 v = __float2bfloat16_rn(x);
  return *reinterpret_cast<uint16_t*>(&v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void residual_ln_pack_kernel(float* __restrict__ x, long long ld, const float* __restrict__ add,
                                        long long ld_add, int M, int m_pad, int D, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ out,
                                        int kpad) {
  const int lane = threadIdx.x & 31;
  const long long row = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  if (row >= m_pad) return;
  uint16_t* o16 = reinterpret_cast<uint16_t*>(out);
  if (row >= M) {
    for (int c = lane; c < kpad; c += 32)
      o16[packed_index(static_cast<size_t>(row), static_cast<size_t>(c), static_cast<size_t>(kpad), kTileM)] = 0;
    return;
  }
  float v[kMaxPerLane];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + 32 * i;
    v[i] = 0.f;
    if (c < D) {
      float t = x[row * ld + c];
      if (add) {
        t += add[row * ld_add + c];
        x[row * ld + c] = t;
      }
      v[i] = t;
      s += t;
    }
  }
  float mean = 0.f, rstd = 1.f;
  if (gamma) {
    mean = warp_sum(s) / static_cast<float>(D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i)
      if (lane + 32 * i < D) q += (v[i] - mean) * (v[i] - mean);
    rstd = rsqrtf(warp_sum(q) / static_cast<float>(D) + eps);
  }
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + 32 * i;
    if (c < kpad) {
      float y = 0.f;
      if (c < D) y = gamma ? (v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c) : v[i];
      o16[packed_index(static_cast<size_t>(row), static_cast<size_t>(c), static_cast<size_t>(kpad), kTileM)] = bf16_bits(y);
    }
  }
}

template <int K>
__global__ void mixer_attn_kernel(const float* __restrict__ qkv, long long ld, int N, int D, int symmetric_qk,
                                  float scale, float attn_eps, float coeff, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, float ln_eps, __nv_bfloat16* __restrict__ out,
                                  int kpad, int rows_pad) {
  const int lane = threadIdx.x & 31;
  const long long n = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  uint16_t* o16 = reinterpret_cast<uint16_t*>(out);
  if (n >= N) {
    // rows beyond N*K of the operand image: zeros (one warp per K padding rows)
    for (int i = 0; i < K; ++i) {
      const long long row = n * K + i;
      if (row < rows_pad)
        for (int c = lane; c < kpad; c += 32)
          o16[packed_index(static_cast<size_t>(row), static_cast<size_t>(c), static_cast<size_t>(kpad), kTileM)] = 0;
    }
    return;
  }
  const float* base = qkv + n * K * ld;
  // qk[i][j] = q_i . k_j  (fp32, rssm_slots_attention.py:192)
  float qk[K][K];
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int j = 0; j < K; ++j) qk[i][j] = 0.f;
  for (int c = lane; c < D; c += 32) {
    float q[K], k[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
      q[i] = base[i * ld + c];
      k[i] = symmetric_qk ? q[i] : base[i * ld + D + c];
    }
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
      for (int j = 0; j < K; ++j) qk[i][j] = fmaf(q[i], k[j], qk[i][j]);
  }
  float attn[K][K];
#pragma unroll
  for (int i = 0; i < K; ++i) {
    float mx = -3.0e38f;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      qk[i][j] = warp_sum(qk[i][j]) * scale;
      mx = fmaxf(mx, qk[i][j]);
    }
    float se = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      attn[i][j] = expf(qk[i][j] - mx);
      se += attn[i][j];
    }
    float tot = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      attn[i][j] = attn[i][j] / se + attn_eps;
      tot += attn[i][j];
    }
#pragma unroll
    for (int j = 0; j < K; ++j) attn[i][j] = coeff * (attn[i][j] / tot) + (1.0f - coeff) * (i == j ? 1.0f : 0.f);
  }
  // updates_i = sum_j attn_ij v_j, then fc_norm
  float u[K][kMaxPerLane];
  float s[K];
#pragma unroll
  for (int i = 0; i < K; ++i) s[i] = 0.f;
#pragma unroll
  for (int t = 0; t < kMaxPerLane; ++t) {
    const int c = lane + 32 * t;
    float v[K];
#pragma unroll
    for (int j = 0; j < K; ++j) v[j] = c < D ? base[j * ld + 2 * D + c] : 0.f;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < K; ++j) a = fmaf(attn[i][j], v[j], a);
      u[i][t] = a;
      s[i] += a;
    }
  }
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const float mean = warp_sum(s[i]) / static_cast<float>(D);
    float q = 0.f;
#pragma unroll
    for (int t = 0; t < kMaxPerLane; ++t)
      if (lane + 32 * t < D) q += (u[i][t] - mean) * (u[i][t] - mean);
    const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(D) + ln_eps);
    const size_t row = static_cast<size_t>(n) * K + i;
#pragma unroll
    for (int t = 0; t < kMaxPerLane; ++t) {
      const int c = lane + 32 * t;
      if (c < kpad) {
        const float y = c < D ? (u[i][t] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c) : 0.f;
        o16[packed_index(row, static_cast<size_t>(c), static_cast<size_t>(kpad), kTileM)] = bf16_bits(y);
      }
    }
  }
}

__global__ void slot_gather_kernel(const __nv_bfloat16* __restrict__ src, int N, int K, int kpad,
                                   __nv_bfloat16* __restrict__ dst, int m_pad) {
  const int chunks = kpad >> 3;
  const long long total = static_cast<long long>(K) * m_pad * chunks;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % chunks);
    const int n = static_cast<int>((i / chunks) % m_pad);
    const int k = static_cast<int>(i / (static_cast<long long>(chunks) * m_pad));
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (n < N)
      v = *reinterpret_cast<const uint4*>(src + packed_index(static_cast<size_t>(n) * K + k, static_cast<size_t>(ch) * 8,
                                                            static_cast<size_t>(kpad), kTileM));
    *reinterpret_cast<uint4*>(dst + static_cast<size_t>(k) * m_pad * kpad +
                              packed_index(static_cast<size_t>(n), static_cast<size_t>(ch) * 8,
                                           static_cast<size_t>(kpad), kTileM)) = v;
  }
}

__global__ void bias_fold_kernel(const float* __restrict__ W, long long ld, int n_out, int n_in,
                                 const float* __restrict__ pos, const float* __restrict__ bias,
                                 float* __restrict__ out, int out_pad) {
  const int lane = threadIdx.x & 31;
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (j >= out_pad) return;
  float acc = 0.f;
  if (j < n_out)
    for (int i = lane; i < n_in; i += 32) acc = fmaf(__ldg(W + static_cast<long long>(j) * ld + i), __ldg(pos + i), acc);
  acc = warp_sum(acc);
  if (lane == 0) out[j] = j < n_out ? acc + (bias ? bias[j] : 0.f) : 0.f;
}

}  // namespace

int launch_residual_ln_pack(float* x, long long ld, const float* add, long long ld_add, int M, int m_pad, int D,
                            const float* gamma, const float* beta, float eps, __nv_bfloat16* out, int kpad,
                            cudaStream_t stream) {
  if (!x || !out || D > 32 * kMaxPerLane || kpad > 32 * kMaxPerLane || (kpad & 63)) return -1;
  const long long threads = static_cast<long long>(m_pad) * 32;
  residual_ln_pack_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, stream>>>(
      x, ld, add, ld_add, M, m_pad, D, gamma, beta, eps, out, kpad);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_mixer_attn(const float* qkv, long long ld, int N, int K, int D, int symmetric_qk, float scale,
                      float attn_eps, float coeff, const float* gamma, const float* beta, float ln_eps,
                      __nv_bfloat16* out, int kpad, int rows_pad, cudaStream_t stream) {
  if (!qkv || !out || !gamma || !beta || D > 32 * kMaxPerLane || kpad > 32 * kMaxPerLane || K < 2 || K > 4) return -1;
  const long long warps = (rows_pad + K - 1) / K;
  const unsigned blocks = static_cast<unsigned>((warps * 32 + 127) / 128);
#define RLSB_MIX(KK)                                                                                              \
  mixer_attn_kernel<KK><<<blocks, 128, 0, stream>>>(qkv, ld, N, D, symmetric_qk, scale, attn_eps, coeff, gamma, \
                                                    beta, ln_eps, out, kpad, rows_pad)
  if (K == 2) RLSB_MIX(2);
  else if (K == 3) RLSB_MIX(3);
  else RLSB_MIX(4);
#undef RLSB_MIX
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_slot_gather(const __nv_bfloat16* src, int N, int K, int kpad, __nv_bfloat16* dst, int m_pad,
                       cudaStream_t stream) {
  const long long total = static_cast<long long>(K) * m_pad * (kpad >> 3);
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  slot_gather_kernel<<<blocks, 256, 0, stream>>>(src, N, K, kpad, dst, m_pad);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_bias_fold(const float* W, long long ld, int n_out, int n_in, const float* pos, const float* bias,
                     float* out, int out_pad, cudaStream_t stream) {
  bias_fold_kernel<<<(out_pad * 32 + 127) / 128, 128, 0, stream>>>(W, ld, n_out, n_in, pos, bias, out, out_pad);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

}  // namespace rlsb
