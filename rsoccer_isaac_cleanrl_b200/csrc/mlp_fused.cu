// mlp_fused.cu — the Agent's whole tanh MLP forward (obs -> 256 -> 512 -> 512 -> 256 -> head) in ONE launch (sm_100a).
//
// Replaces, for the per-step policy / value calls of the rollout (ppo_continuous_action_isaacgym.py:155-164, 259-261:
// `agent.get_action_and_value(next_obs)` = actor_mean(x), critic(x)), the chain of 2 x (4 GEMM launches + head launch):
// at 4 096 rows those ten launches are pure launch and pipeline-fill latency (6.8 us per GEMM for 0.5 GFLOP).
//
// One CTA owns 128 rows of the batch for one network (grid = row blocks x networks) and walks the four hidden layers:
//   * the activations of the row block never leave the SM: a 128 x 512 bf16 buffer in shared memory, in the
//     K-major 128-byte-swizzled layout tcgen05.mma reads as its A operand (what TMA would have produced);
//   * warp 0 streams the weights (1.06 MB per network, L2-resident) through a 3 x 32 KB ring with TMA, running
//     ahead across layer boundaries;
//   * warp 1: one elected thread issues tcgen05.mma (M = 128, N = 256, K = 16); a layer's accumulator
//     (128 x 256 / 512 fp32) is the whole TMEM (512 columns);
//   * warps 2..: epilogue — tcgen05.ld (software-pipelined), + bias, tanh.approx, bf16, written back INTO the
//     activation buffer as the next layer's A operand (fence.proxy.async, mbarrier to the MMA thread);
//     the last hidden layer feeds the 256 -> {1, 2, 6} head from registers, so only (rows, n_out) floats go to HBM.
// Same arithmetic per element as the one-GEMM-per-launch path (tc_gemm.cu: fp32 accumulation over K in the same
// order, (acc + bias) -> tanh.approx.f32 -> bf16 round-to-nearest), so the hidden activations are the same bits; the
// head sums its 256 products in a different order (fp32).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <string>

#include "../../include/vss_b200.h"
#include "tc_common.cuh"

namespace tc {

constexpr int FM_NCHUNK = 256;                      // weight rows per ring stage = UMMA N
constexpr int FM_STAGE_BYTES = FM_NCHUNK * BK * 2;  // 32 KB
constexpr int FM_STAGES = 3;
constexpr int FM_BLOCK_BYTES = BM * BK * 2;         // one 128 x 64 k-block of the activation buffer: 16 KB
constexpr int FM_ACT_BYTES = 8 * FM_BLOCK_BYTES;    // 128 rows x 512 columns bf16
constexpr int FM_SMEM = FM_ACT_BYTES + FM_STAGES * FM_STAGE_BYTES + 128 /*barriers*/ + 1024 /*alignment slack*/;
constexpr int FM_MAX_NETS = 2;

__host__ __device__ constexpr int fm_k(int l) { return l == 0 ? 64 : (l == 1 ? 256 : 512); }
__host__ __device__ constexpr int fm_n(int l) { return (l == 0 || l == 3) ? 256 : 512; }

struct FusedNet {
  CUtensorMap w[4];       // bf16 [fm_n(l), fm_k(l)] row-major, box 64 x 256
  const float* bias[4];
  const float* head_w;    // f32 [n_out, 256]
  const float* head_b;    // f32 [n_out]
  float* out;             // f32 [M, n_out]
  int n_out;              // 1, 2 or 6
  int pad_;
};
struct FusedArgs {
  CUtensorMap x;          // bf16 [M, 64] (observations, zero-padded to 64 columns), box 64 x 128
  FusedNet net[FM_MAX_NETS];
  int M;
  unsigned long long* stamps;  // profiling (vss_mlp_forward_fused_timed): CTA (0,0) records globaltimer at its phase boundaries
};

__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// stamps[0] kernel entry, [1] setup done, [2 + 2l] layer l accumulator complete (seen by epilogue warp 2),
// [3 + 2l] layer l epilogue done by warp 2, [10] exit
__device__ __forceinline__ void stamp(const FusedArgs& g, int i) {
  if (g.stamps && blockIdx.x == 0 && blockIdx.y == 0) g.stamps[i] = gtime_ns();
}

// tanh(acc + bias) of 32 consecutive columns of one row, as 16 packed bf16 pairs
__device__ __forceinline__ void bias_tanh_pack(const uint32_t (&v)[32], const float* __restrict__ bias, uint32_t (&packed)[16]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 bv = __ldg(reinterpret_cast<const float4*>(bias) + j);
    const float a = tanh_fast(__uint_as_float(v[4 * j]) + bv.x);
    const float b = tanh_fast(__uint_as_float(v[4 * j + 1]) + bv.y);
    const float c = tanh_fast(__uint_as_float(v[4 * j + 2]) + bv.z);
    const float d = tanh_fast(__uint_as_float(v[4 * j + 3]) + bv.w);
    __nv_bfloat162 p = __floats2bfloat162_rn(a, b), q = __floats2bfloat162_rn(c, d);
    packed[2 * j] = *reinterpret_cast<uint32_t*>(&p);
    packed[2 * j + 1] = *reinterpret_cast<uint32_t*>(&q);
  }
}

// Hidden layer: 32 columns [c, c + 32) of this thread's row into the activation buffer (K-major SWIZZLE_128B:
// k-block c / 64, row pitch 128 B, 16-byte piece j of a row stored at piece j ^ (row % 8)).
__device__ __forceinline__ void store_act_chunk(const uint32_t (&v)[32], const float* __restrict__ bias, uint8_t* act,
                                                int row, int c) {
  uint32_t packed[16];
  bias_tanh_pack(v, bias + c, packed);
  uint8_t* base = act + (c >> 6) * FM_BLOCK_BYTES + row * 128;
  const int j0 = (c & 63) >> 3, sw = row & 7;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(base + (((j0 + j) ^ sw) << 4)) =
        make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
}

// Last hidden layer: the same values (rounded to bf16 like the stored activations of the unfused path) times the
// head's weight rows, accumulated per output.
template <int NO>
__device__ __forceinline__ void head_chunk(const uint32_t (&v)[32], const float* __restrict__ bias,
                                           const float* __restrict__ head_w, int c, float (&hacc)[6]) {
  uint32_t packed[16];
  bias_tanh_pack(v, bias + c, packed);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const __nv_bfloat162 p = *reinterpret_cast<const __nv_bfloat162*>(&packed[2 * j]);
    const __nv_bfloat162 q = *reinterpret_cast<const __nv_bfloat162*>(&packed[2 * j + 1]);
    const float x0 = __bfloat162float(p.x), x1 = __bfloat162float(p.y), x2 = __bfloat162float(q.x), x3 = __bfloat162float(q.y);
#pragma unroll
    for (int a = 0; a < NO; ++a) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(head_w + a * 256 + c) + j);
      hacc[a] = fmaf(x3, w.w, fmaf(x2, w.z, fmaf(x1, w.y, fmaf(x0, w.x, hacc[a]))));
    }
  }
}

template <int EW>  // epilogue warps: 4 (one per TMEM lane quarter) or 8 (two per quarter, half of the columns each)
__global__ void __launch_bounds__(64 + 32 * EW, 1)
k_mlp_fwd(const __grid_constant__ FusedArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* act = smem;
  uint8_t* ring = smem + FM_ACT_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + FM_STAGES * FM_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + FM_STAGES;
  uint64_t* x_full = empty_bar + FM_STAGES;   // the observation tile has landed in k-block 0
  uint64_t* acc_full = x_full + 1;            // all MMAs of the current layer have retired (phase = layer & 1)
  uint64_t* act_ready = acc_full + 1;         // the epilogue has written the next layer's A operand (phase = layer & 1)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(act_ready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const FusedNet& net = g.net[blockIdx.y];
  const int m0 = blockIdx.x * BM;
  if (threadIdx.x == 64) stamp(g, 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < FM_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(x_full, 1); mbar_init(acc_full, 1); mbar_init(act_ready, EW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 64) stamp(g, 1);

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer: the observation tile, then every weight stage of the four layers in order
      mbar_expect_tx(x_full, FM_BLOCK_BYTES);
      tma_load_2d(act, &g.x, x_full, 0, m0);
      uint32_t s = 0, ph = 0;
#pragma unroll 1
      for (int l = 0; l < 4; ++l) {
        const int chunks = fm_n(l) / FM_NCHUNK, kbs = fm_k(l) / BK;
        for (int ch = 0; ch < chunks; ++ch)
          for (int kb = 0; kb < kbs; ++kb) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            mbar_expect_tx(&full_bar[s], FM_STAGE_BYTES);
            tma_load_2d(ring + s * FM_STAGE_BYTES, &net.w[l], &full_bar[s], kb * BK, ch * FM_NCHUNK);
            if (++s == FM_STAGES) { s = 0; ph ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---- MMA issuer
      constexpr uint32_t idesc = make_idesc(BM, FM_NCHUNK, false);
      const uint32_t act_addr = smem_u32(act), ring_addr = smem_u32(ring);
      uint32_t s = 0, ph = 0;
#pragma unroll 1
      for (int l = 0; l < 4; ++l) {
        if (l == 0) mbar_wait(x_full, 0); else mbar_wait(act_ready, (uint32_t)(l - 1) & 1u);
        tc_fence_after();
        const int chunks = fm_n(l) / FM_NCHUNK, kbs = fm_k(l) / BK;
        for (int ch = 0; ch < chunks; ++ch) {
          const uint32_t tmem_d = tmem_base + (uint32_t)(ch * FM_NCHUNK);
          for (int kb = 0; kb < kbs; ++kb) {
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint64_t adesc = make_desc_k128(act_addr + kb * FM_BLOCK_BYTES);
            const uint64_t bdesc = make_desc_k128(ring_addr + s * FM_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            umma_commit(&empty_bar[s]);  // frees the ring stage when these MMAs retire
            if (++s == FM_STAGES) { s = 0; ph ^= 1; }
          }
        }
        umma_commit(acc_full);  // the layer's accumulator is complete, and the activation buffer is no longer read
      }
    }
  } else {
    // ---- epilogue: this thread owns row q * 32 + lane of the block (TMEM lane), warp half h takes columns [h * N / 2 ...)
    const int q = warp & 3, half = (warp - 2) >> 2, row = q * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    float hacc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int l = 0; l < 4; ++l) {
      const int per_warp = fm_n(l) / (EW / 4), c_begin = half * per_warp, c_end = c_begin + per_warp;
      const float* __restrict__ bias = net.bias[l];
      mbar_wait(acc_full, (uint32_t)l & 1u);
      tc_fence_after();
      if (threadIdx.x == 64) stamp(g, 2 + 2 * l);
      uint32_t va[32], vb[32];
      tmem_ld32_issue(t_row + (uint32_t)c_begin, va);
#pragma unroll 1
      for (int c = c_begin; c < c_end; c += 64) {
        tmem_ld_wait();
        tmem_ld32_issue(t_row + (uint32_t)(c + 32), vb);
        if (l < 3) store_act_chunk(va, bias, act, row, c);
        else if (net.n_out == 1) head_chunk<1>(va, bias, net.head_w, c, hacc);
        else if (net.n_out == 2) head_chunk<2>(va, bias, net.head_w, c, hacc);
        else head_chunk<6>(va, bias, net.head_w, c, hacc);
        tmem_ld_wait();
        if (c + 64 < c_end) tmem_ld32_issue(t_row + (uint32_t)(c + 64), va);
        if (l < 3) store_act_chunk(vb, bias, act, row, c + 32);
        else if (net.n_out == 1) head_chunk<1>(vb, bias, net.head_w, c + 32, hacc);
        else if (net.n_out == 2) head_chunk<2>(vb, bias, net.head_w, c + 32, hacc);
        else head_chunk<6>(vb, bias, net.head_w, c + 32, hacc);
      }
      if (l < 3) {
        // generic-proxy writes -> visible to the tensor core's (async proxy) reads; TMEM reads done before the
        // next layer's MMAs overwrite the accumulator
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(act_ready);
      }
      if (threadIdx.x == 64) stamp(g, 3 + 2 * l);
    }
    // head: combine the column halves (the activation buffer is free: every MMA has retired), add the bias, store
    if (EW == 8) {
      float* part = reinterpret_cast<float*>(act) + row * 8;
      if (half == 1) {
#pragma unroll
        for (int a = 0; a < 6; ++a) part[a] = hacc[a];
      }
      asm volatile("bar.sync 1, %0;" ::"r"(32 * EW) : "memory");
      if (half == 0) {
#pragma unroll
        for (int a = 0; a < 6; ++a) hacc[a] += part[a];
      }
    }
    if (half == 0 && m0 + row < g.M) {
      float* o = net.out + (size_t)(m0 + row) * net.n_out;
      for (int a = 0; a < net.n_out; ++a) o[a] = hacc[a] + __ldg(net.head_b + a);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 64) stamp(g, 10);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace tc

extern thread_local std::string g_tc_error;

extern "C" {

// The whole MLP forward for up to two networks that share the input rows (actor mean and critic value of the
// same observations): out_i [M, n_out_i] f32 = head_i(tanh-MLP_i(x)). See include/vss_b200.h.
VSS_API int vss_mlp_forward_fused(const void* x16, int ldx, int M, const vss_mlp_net* nets, int n_nets, int epilogue_warps,
                                  void* stream) {
  return vss_mlp_forward_fused_timed(x16, ldx, M, nets, n_nets, epilogue_warps, nullptr, stream);
}

// The same launch; the first CTA also records 11 globaltimer stamps (ns) at its phase boundaries into `stamps`
// (device memory): [0] entry, [1] barriers + TMEM ready, [2 + 2l] accumulator of layer l complete, [3 + 2l] epilogue of
// layer l done (one epilogue warp's view), [10] exit. A profiling hook: profiles/mlp_fused_bench.py.
VSS_API int vss_mlp_forward_fused_timed(const void* x16, int ldx, int M, const vss_mlp_net* nets, int n_nets,
                                        int epilogue_warps, unsigned long long* stamps, void* stream) {
  using namespace tc;
  if (!x16 || !nets || M <= 0 || n_nets < 1 || n_nets > FM_MAX_NETS || ldx < 64 || (ldx & 7)) {
    g_tc_error = "vss_mlp_forward_fused: bad argument (x16 [M, 64] bf16 with ldx >= 64, 1 or 2 networks)";
    return VSS_E_INVALID;
  }
  if (epilogue_warps == 0) epilogue_warps = 8;
  if (epilogue_warps != 4 && epilogue_warps != 8) { g_tc_error = "vss_mlp_forward_fused: epilogue_warps must be 0 (default), 4 or 8"; return VSS_E_INVALID; }
  FusedArgs a;
  memset(&a, 0, sizeof(a));
  a.M = M;
  a.stamps = stamps;
  if (!make_map(&a.x, x16, M, 64, ldx, BM)) { g_tc_error = "vss_mlp_forward_fused: cuTensorMapEncodeTiled failed (x)"; return VSS_E_CUDA; }
  for (int i = 0; i < n_nets; ++i) {
    const vss_mlp_net& s = nets[i];
    if (!s.head_w || !s.head_b || !s.out || (s.n_out != 1 && s.n_out != 2 && s.n_out != 6)) {
      g_tc_error = "vss_mlp_forward_fused: bad network (head 256 -> 1, 2 or 6)";
      return VSS_E_INVALID;
    }
    for (int l = 0; l < 4; ++l) {
      if (!s.w[l] || !s.b[l] || (reinterpret_cast<uintptr_t>(s.w[l]) & 15u) || (reinterpret_cast<uintptr_t>(s.b[l]) & 15u)) {
        g_tc_error = "vss_mlp_forward_fused: weights / biases must be non-null and 16-byte aligned";
        return VSS_E_INVALID;
      }
      if (!make_map(&a.net[i].w[l], s.w[l], fm_n(l), fm_k(l), fm_k(l), FM_NCHUNK)) {
        g_tc_error = "vss_mlp_forward_fused: cuTensorMapEncodeTiled failed (weights)";
        return VSS_E_CUDA;
      }
      a.net[i].bias[l] = s.b[l];
    }
    if (reinterpret_cast<uintptr_t>(s.head_w) & 15u) { g_tc_error = "vss_mlp_forward_fused: head_w must be 16-byte aligned"; return VSS_E_INVALID; }
    a.net[i].head_w = s.head_w; a.net[i].head_b = s.head_b; a.net[i].out = s.out; a.net[i].n_out = s.n_out;
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_fwd<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, FM_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mlp_fwd<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, FM_SMEM);
    if (e != cudaSuccess) { g_tc_error = std::string("vss_mlp_forward_fused: ") + cudaGetErrorString(e); return VSS_E_CUDA; }
    configured = true;
  }
  const dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)n_nets);
  cudaStream_t st = (cudaStream_t)stream;
  if (epilogue_warps == 4) k_mlp_fwd<4><<<grid, 64 + 32 * 4, FM_SMEM, st>>>(a);
  else k_mlp_fwd<8><<<grid, 64 + 32 * 8, FM_SMEM, st>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_tc_error = std::string("vss_mlp_forward_fused: ") + cudaGetErrorString(e); return VSS_E_CUDA; }
  return VSS_OK;
}

}  // extern "C"
