// ppo_sample.cuh - Normal(mean, exp(logstd)).sample() + log_prob().sum(1) of one row (ppo_continuous_action_isaacgym.py:
// 155-164), shared by vss_policy_sample (ppo_ops.cu) and the fused MLP forward (mlp_fused.cu) so that both draw the same
// numbers. Philox4x32-10, counter = (row, call index, stream), key = seed: the stream of normals depends only on
// (seed, call, row), not on the launch shape.
#pragma once
#include "vss_lane.cuh"

namespace ppo {

constexpr float HALF_LOG_2PI = 0.9189385332046727f;

// action[0..A) = mean + exp(logstd) * z; returns the summed log-probability of the action.
template <int A>
__device__ __forceinline__ float sample_row(const float (&mu)[A], const float* __restrict__ logstd, long long i, uint32_t call,
                                            uint32_t seed_lo, uint32_t seed_hi, float (&act)[A]) {
  float z[(A + 3) / 4 * 4];
#pragma unroll
  for (int b = 0; b < (A + 3) / 4; ++b) {
    const vss::U4 g = vss::philox4x32_10(vss::U4{(uint32_t)i, (uint32_t)(i >> 32), call, 0x50504f00u + b}, seed_lo, seed_hi);
    const uint32_t u[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float rad = sqrtf(-2.0f * logf(vss::u01_open(u[2 * h])));
      float sn, cs;
      sincosf(6.283185307179586f * vss::u01(u[2 * h + 1]), &sn, &cs);
      z[4 * b + 2 * h] = rad * cs;
      z[4 * b + 2 * h + 1] = rad * sn;
    }
  }
  float lp = 0.0f;
#pragma unroll
  for (int a = 0; a < A; ++a) {
    const float ls = logstd[a], sd = expf(ls);
    act[a] = mu[a] + sd * z[a];
    const float d = act[a] - mu[a];
    lp += -(d * d) / (2.0f * sd * sd) - ls - HALF_LOG_2PI;
  }
  return lp;
}

}  // namespace ppo
