// ppo_ops.cu — the small (HBM / latency bound) pieces of the PPO loop as single kernels for sm_100a:
//   vss_policy_sample     Normal(mean, exp(logstd)).sample() + log_prob().sum(1)          (ppo…:155-164)
//   vss_ppo_loss          minibatch gather + advantage normalisation + clipped surrogate,
//                         value and entropy losses, AND their gradients w.r.t. the network
//                         outputs, plus the logging statistics                            (ppo…:314-352)
//   vss_convert_bf16_batch  bf16 (optionally transposed / K-padded) copies of the fp32 master weights
//   vss_clip_adam         clip_grad_norm_ + Adam on the flat parameter buffer              (ppo…:353-354)
// In the reference each of these is 15-100 torch launches; at minibatch 131 072 they cost more than
// a third of the tensor-core time of the MLPs they sit between.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <string>

#include "../../include/vss_b200.h"
#include "ppo_sample.cuh"
#include "vss_lane.cuh"

namespace ppo {

using vss::U4;


__global__ void k_bump32(uint32_t* ctr) { *ctr += 1u; }

// ---- action sampling ----------------------------------------------------------------------
// One thread per row (the arithmetic lives in ppo_sample.cuh, shared with the fused MLP forward).
template <int A>
__global__ void __launch_bounds__(256)
k_policy_sample(const float* __restrict__ mean, const float* __restrict__ logstd, long long M, uint32_t seed_lo,
                uint32_t seed_hi, const uint32_t* __restrict__ counter, float* __restrict__ action,
                float* __restrict__ logprob) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  float mu[A], act[A];
#pragma unroll
  for (int a = 0; a < A; ++a) mu[a] = mean[i * A + a];
  logprob[i] = sample_row<A>(mu, logstd, i, *counter, seed_lo, seed_hi, act);
#pragma unroll
  for (int a = 0; a < A; ++a) action[i * A + a] = act[a];
}

// ---- rows of the rollout that ended an episode ---------------------------------------------
// list[k] = index of the k-th non-zero flag (any order), *count = how many there are (may exceed cap: the caller
// checks), for flags[0 .. n). One warp-aggregated atomic per warp.
__global__ void __launch_bounds__(256)
k_compact_nonzero(const float* __restrict__ flags, long long n, long long* __restrict__ list, int cap, int* count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool on = i < n && flags[i] != 0.0f;
  const unsigned m = __ballot_sync(0xffffffffu, on);
  if (m == 0u) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == __ffs((int)m) - 1) base = atomicAdd(count, __popc(m));
  base = __shfl_sync(0xffffffffu, base, __ffs((int)m) - 1);
  if (on) {
    const int k = base + __popc(m & ((1u << lane) - 1u));
    if (k < cap) list[k] = i;
  }
}
// dst[list[k]] = src[k] for k < min(*count, cap)
__global__ void __launch_bounds__(256)
k_scatter_rows(float* __restrict__ dst, const long long* __restrict__ list, const float* __restrict__ src,
               const int* __restrict__ count, int cap) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < min(*count, cap)) dst[list[k]] = src[k];
}

// ---- minibatch loss + gradients -----------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// sum and sum of squares of adv[inds[i]] in fp64 -> acc[0], acc[1]
__global__ void __launch_bounds__(256)
k_adv_stats(const float* __restrict__ adv, const long long* __restrict__ inds, long long B, double* acc) {
  double s = 0.0, q = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (long long)gridDim.x * blockDim.x) {
    const double a = (double)adv[inds ? inds[i] : i];
    s += a; q += a * a;
  }
  __shared__ double red[2][8];
  s = warp_sum(s); q = warp_sum(q);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { red[0][w] = s; red[1][w] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { ts += red[0][k]; tq += red[1][k]; }
    atomicAdd(acc, ts); atomicAdd(acc + 1, tq);
  }
}

struct LossArgs {
  const float *mean, *value, *logstd, *b_action, *b_logprob, *b_adv, *b_ret, *b_val;
  const long long* inds;
  long long B;
  float clip, ent_coef, vf_coef;
  int norm_adv, clip_vloss;
  float *d_mean, *d_value, *d_logstd, *stats;
  const double* adv_acc;
};

// stats: 0 pg_loss, 1 v_loss, 2 entropy, 3 old_approx_kl, 4 approx_kl, 5 clipfrac, 6 loss
constexpr int N_STATS = 8;

template <int A>
__global__ void __launch_bounds__(256)
k_ppo_loss(const LossArgs a) {
  const float invB = 1.0f / (float)a.B;
  float amean = 0.0f, ainv = 1.0f;
  if (a.norm_adv) {  // (x - mean) / (std_unbiased + 1e-8), ppo…:327-328
    const double n = (double)a.B, s = a.adv_acc[0], q = a.adv_acc[1];
    const double var = fmax((q - s * s / n) / (n - 1.0), 0.0);
    amean = (float)(s / n);
    ainv = 1.0f / ((float)sqrt(var) + 1e-8f);
  }
  float ls[A], inv_var[A];
#pragma unroll
  for (int k = 0; k < A; ++k) { ls[k] = a.logstd[k]; const float sd = expf(ls[k]); inv_var[k] = 1.0f / (sd * sd); }
  float s_pg = 0.f, s_v = 0.f, s_okl = 0.f, s_kl = 0.f, s_cf = 0.f, s_dls[A];
#pragma unroll
  for (int k = 0; k < A; ++k) s_dls[k] = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.B; i += (long long)gridDim.x * blockDim.x) {
    const long long j = a.inds ? a.inds[i] : i;
    float d[A], lp = 0.0f;
#pragma unroll
    for (int k = 0; k < A; ++k) {
      d[k] = a.b_action[j * A + k] - a.mean[i * A + k];
      lp += -(d[k] * d[k]) * (0.5f * inv_var[k]) - ls[k] - HALF_LOG_2PI;
    }
    const float logratio = lp - a.b_logprob[j];
    const float ratio = expf(logratio);
    s_okl += -logratio;
    s_kl += (ratio - 1.0f) - logratio;
    s_cf += fabsf(ratio - 1.0f) > a.clip ? 1.0f : 0.0f;
    const float adv = (a.b_adv[j] - amean) * ainv;
    const float lo = 1.0f - a.clip, hi = 1.0f + a.clip;
    const float t1 = -adv * ratio, t2 = -adv * fminf(fmaxf(ratio, lo), hi);
    s_pg += fmaxf(t1, t2);
    // d max(t1, t2) / d ratio: inside the clip range both branches carry -adv (torch splits the tie
    // half and half, clamp passes the gradient on [lo, hi]); outside only the unclipped branch can
    const bool inside = ratio >= lo && ratio <= hi;
    const float g_ratio = (inside || t1 > t2) ? -adv : 0.0f;
    const float g_lp = g_ratio * ratio * invB;
#pragma unroll
    for (int k = 0; k < A; ++k) {
      a.d_mean[i * A + k] = g_lp * d[k] * inv_var[k];
      s_dls[k] += g_lp * (d[k] * d[k] * inv_var[k] - 1.0f);
    }
    const float v = a.value[i], ret = a.b_ret[j];
    float g_v;
    if (a.clip_vloss) {  // ppo…:337-345
      const float v0 = a.b_val[j], dv = v - v0;
      const float vc = v0 + fminf(fmaxf(dv, -a.clip), a.clip);
      const float u = (v - ret) * (v - ret), c = (vc - ret) * (vc - ret);
      s_v += 0.5f * fmaxf(u, c);
      const bool in = dv >= -a.clip && dv <= a.clip;
      if (u > c) g_v = v - ret;
      else if (u < c) g_v = in ? (vc - ret) : 0.0f;
      else g_v = 0.5f * (v - ret) + (in ? 0.5f * (vc - ret) : 0.0f);
    } else {
      s_v += 0.5f * (v - ret) * (v - ret);
      g_v = v - ret;
    }
    a.d_value[i] = a.vf_coef * g_v * invB;
  }
  // block reduction -> one atomic per statistic per block
  constexpr int NS = 5 + A;
  __shared__ float red[NS][8];
  float vals[NS] = {s_pg, s_v, s_okl, s_kl, s_cf};
#pragma unroll
  for (int k = 0; k < A; ++k) vals[5 + k] = s_dls[k];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NS; ++k) {
    const float r = warp_sum(vals[k]);
    if (lane == 0) red[k][w] = r;
  }
  __syncthreads();
  if (threadIdx.x < NS) {
    float t = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[threadIdx.x][k];
    if (threadIdx.x < 5) {
      const int slot[5] = {0, 1, 3, 4, 5};
      atomicAdd(a.stats + slot[threadIdx.x], t * invB);
      if (threadIdx.x == 0) atomicAdd(a.stats + 6, t * invB);                 // loss += pg
      if (threadIdx.x == 1) atomicAdd(a.stats + 6, a.vf_coef * t * invB);     // loss += vf_coef * v
    } else {
      atomicAdd(a.d_logstd + (threadIdx.x - 5), t);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // entropy of Normal does not depend on the sample
    float ent = 0.f;
#pragma unroll
    for (int k = 0; k < A; ++k) { ent += 0.5f + HALF_LOG_2PI + ls[k]; atomicAdd(a.d_logstd + k, -a.ent_coef); }
    atomicAdd(a.stats + 2, ent);
    atomicAdd(a.stats + 6, -a.ent_coef * ent);
  }
}

// ---- weight copies ------------------------------------------------------------------------
struct ConvJob { const float* src; __nv_bfloat16* dst; int rows, cols, ld_dst, transpose; };
struct ConvJobs { ConvJob j[VSS_MAX_CONVERT_JOBS]; };

// dst[r, c] = src[r, c] (or dst[c, r] = src[r, c] when transposing); columns of dst beyond the
// source are left as they are (the caller zeroes the padding once).
__global__ void __launch_bounds__(256)
k_convert_bf16(const ConvJobs jobs) {
  const ConvJob J = jobs.j[blockIdx.y];
  __shared__ float tile[32][33];
  const int tiles_c = (J.cols + 31) / 32, tiles_r = (J.rows + 31) / 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int t = blockIdx.x; t < tiles_c * tiles_r; t += gridDim.x) {
    const int r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
    if (!J.transpose) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = r0 + ty + 8 * k, c = c0 + tx;
        if (r < J.rows && c < J.cols) J.dst[(size_t)r * J.ld_dst + c] = __float2bfloat16(J.src[(size_t)r * J.cols + c]);
      }
    } else {
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = r0 + ty + 8 * k, c = c0 + tx;
        tile[ty + 8 * k][tx] = (r < J.rows && c < J.cols) ? J.src[(size_t)r * J.cols + c] : 0.0f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + 8 * k, r = r0 + tx;  // dst row = source column
        if (r < J.rows && c < J.cols) J.dst[(size_t)c * J.ld_dst + r] = __float2bfloat16(tile[tx][ty + 8 * k]);
      }
    }
  }
}

// ---- clip_grad_norm_ + Adam ---------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_sumsq(const float* __restrict__ g, long long n, float* acc) {
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = g[i];
    s += v * v;
  }
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[k];
    atomicAdd(acc, t);
  }
}

// state: [0] step count t (as float), [1] learning rate, [2] scratch: sum of squares of the gradient
__global__ void __launch_bounds__(256)
k_clip_adam(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
            const float* __restrict__ state, float grad_scale, float max_norm, float b1, float b2, float eps) {
  const float t = state[0] + 1.0f, lr = state[1];
  // clip_grad_norm_: coefficient min(1, max_norm / (norm + 1e-6)) on the (rank-averaged) gradient
  const float norm = sqrtf(state[2]) * grad_scale;
  const float coef = grad_scale * fminf(max_norm / (norm + 1e-6f), 1.0f);
  const float bc1 = 1.0f - powf(b1, t), bc2 = 1.0f - powf(b2, t);
  const float step = lr / bc1, isq2 = 1.0f / sqrtf(bc2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = m[i] * b1 + (1.0f - b1) * gi;
    const float vi = v[i] * b2 + (1.0f - b2) * gi * gi;
    m[i] = mi; v[i] = vi; g[i] = gi;
    p[i] -= step * (mi / (sqrtf(vi) * isq2 + eps));
  }
}
__global__ void k_adam_tick(float* state) { state[0] += 1.0f; state[2] = 0.0f; }

thread_local std::string g_err;
static int bad(const char* what) { g_err = what; return VSS_E_INVALID; }
static int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return VSS_OK;
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return VSS_E_CUDA;
}

}  // namespace ppo

extern "C" {

VSS_API const char* vss_ppo_last_error(void) { return ppo::g_err.c_str(); }

VSS_API int vss_policy_sample(const float* mean, const float* logstd, int64_t M, int A, uint64_t seed,
                              uint32_t* counter, float* action, float* logprob, void* stream) {
  if (!mean || !logstd || !counter || !action || !logprob || M <= 0) return ppo::bad("vss_policy_sample: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((M + 255) / 256);
  const uint32_t lo = (uint32_t)seed, hi = (uint32_t)(seed >> 32);
  if (A == 2) ppo::k_policy_sample<2><<<grid, 256, 0, st>>>(mean, logstd, M, lo, hi, counter, action, logprob);
  else if (A == 6) ppo::k_policy_sample<6><<<grid, 256, 0, st>>>(mean, logstd, M, lo, hi, counter, action, logprob);
  else return ppo::bad("vss_policy_sample: action width must be 2 or 6");
  ppo::k_bump32<<<1, 1, 0, st>>>(counter);
  return ppo::check_launch("vss_policy_sample");
}

VSS_API int vss_compact_nonzero(const float* flags, int64_t n, int64_t* list, int cap, int* count, void* stream) {
  if (!flags || !list || !count || n <= 0 || cap <= 0) return ppo::bad("vss_compact_nonzero: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(count, 0, sizeof(int), st) != cudaSuccess) return ppo::check_launch("vss_compact_nonzero: memset");
  ppo::k_compact_nonzero<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(flags, n, reinterpret_cast<long long*>(list), cap, count);
  return ppo::check_launch("vss_compact_nonzero");
}

VSS_API int vss_scatter_rows_f32(float* dst, const int64_t* list, const float* src, const int* count, int cap, void* stream) {
  if (!dst || !list || !src || !count || cap <= 0) return ppo::bad("vss_scatter_rows_f32: bad argument");
  ppo::k_scatter_rows<<<(unsigned)((cap + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      dst, reinterpret_cast<const long long*>(list), src, count, cap);
  return ppo::check_launch("vss_scatter_rows_f32");
}

VSS_API int vss_ppo_loss(const float* mean, const float* value, const float* logstd, const float* b_action,
                         const float* b_logprob, const float* b_adv, const float* b_ret, const float* b_val,
                         const int64_t* inds, int64_t B, int A, float clip_coef, float ent_coef, float vf_coef,
                         int norm_adv, int clip_vloss, float* d_mean, float* d_value, float* d_logstd, float* stats,
                         double* scratch, void* stream) {
  if (!mean || !value || !logstd || !b_action || !b_logprob || !b_adv || !b_ret || !d_mean || !d_value || !d_logstd ||
      !stats || !scratch || B <= 1 || (clip_vloss && !b_val))
    return ppo::bad("vss_ppo_loss: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(stats, 0, sizeof(float) * ppo::N_STATS, st) != cudaSuccess ||
      cudaMemsetAsync(scratch, 0, sizeof(double) * 2, st) != cudaSuccess)
    return ppo::check_launch("vss_ppo_loss: memset");
  const unsigned grid = (unsigned)std::min<int64_t>((B + 255) / 256, 148 * 4);
  const long long* idx = reinterpret_cast<const long long*>(inds);
  if (norm_adv) ppo::k_adv_stats<<<grid, 256, 0, st>>>(b_adv, idx, B, scratch);
  ppo::LossArgs a{mean, value, logstd, b_action, b_logprob, b_adv, b_ret, b_val, idx, B, clip_coef, ent_coef, vf_coef,
                  norm_adv, clip_vloss, d_mean, d_value, d_logstd, stats, scratch};
  if (A == 2) ppo::k_ppo_loss<2><<<grid, 256, 0, st>>>(a);
  else if (A == 6) ppo::k_ppo_loss<6><<<grid, 256, 0, st>>>(a);
  else return ppo::bad("vss_ppo_loss: action width must be 2 or 6");
  return ppo::check_launch("vss_ppo_loss");
}

VSS_API int vss_convert_bf16_batch(const vss_convert_job* jobs, int njobs, void* stream) {
  if (!jobs || njobs <= 0 || njobs > VSS_MAX_CONVERT_JOBS) return ppo::bad("vss_convert_bf16_batch: bad argument");
  ppo::ConvJobs J;
  int max_tiles = 1;
  for (int k = 0; k < njobs; ++k) {
    const vss_convert_job& s = jobs[k];
    if (!s.src || !s.dst || s.rows <= 0 || s.cols <= 0 || s.ld_dst < (s.transpose ? s.rows : s.cols))
      return ppo::bad("vss_convert_bf16_batch: bad job");
    J.j[k] = ppo::ConvJob{s.src, reinterpret_cast<__nv_bfloat16*>(s.dst), s.rows, s.cols, s.ld_dst, s.transpose};
    max_tiles = std::max(max_tiles, ((s.rows + 31) / 32) * ((s.cols + 31) / 32));
  }
  ppo::k_convert_bf16<<<dim3((unsigned)std::min(max_tiles, 148), (unsigned)njobs), 256, 0, (cudaStream_t)stream>>>(J);
  return ppo::check_launch("vss_convert_bf16_batch");
}

VSS_API int vss_clip_adam(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float* state,
                          float grad_scale, float max_grad_norm, float beta1, float beta2, float eps, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !state || n <= 0) return ppo::bad("vss_clip_adam: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8);
  ppo::k_sumsq<<<grid, 256, 0, st>>>(grads, n, state + 2);
  ppo::k_clip_adam<<<grid, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, state, grad_scale, max_grad_norm,
                                         beta1, beta2, eps);
  ppo::k_adam_tick<<<1, 1, 0, st>>>(state);
  return ppo::check_launch("vss_clip_adam");
}

}  // extern "C"
