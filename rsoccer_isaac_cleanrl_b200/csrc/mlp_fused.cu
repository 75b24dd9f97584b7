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
//   * warp 1: one elected thread issues tcgen05.mma (M = 128, N = 256, K = 16); accumulators of 256 columns alternate
//     between the two halves of the TMEM (512 columns), so a chunk's MMAs run while the previous chunk is drained;
//   * warps 2..: epilogue — tcgen05.ld (software-pipelined), + bias, tanh.approx, bf16, written back INTO the
//     activation buffer as the next layer's A operand (fence.proxy.async, one mbarrier per 64-column block, so the next
//     layer's MMAs start block by block); the last hidden layer feeds the 256 -> {1, 2, 6} head from registers and, for
//     the actor, the action sample (ppo_sample.cuh), so only (rows, n_out) floats, actions and log-probs go to HBM.
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
#include "ppo_sample.cuh"
#include "tc_common.cuh"

namespace tc {

constexpr int FM_NCHUNK = 256;                      // weight rows per ring stage = UMMA N
constexpr int FM_STAGE_BYTES = FM_NCHUNK * BK * 2;  // 32 KB
constexpr int FM_STAGES = 3;
constexpr int FM_BLOCK_BYTES = BM * BK * 2;         // one 128 x 64 k-block of the activation buffer: 16 KB
constexpr int FM_ACT_BYTES = 8 * FM_BLOCK_BYTES;    // 128 rows x 512 columns bf16
constexpr int FM_SMEM = FM_ACT_BYTES + FM_STAGES * FM_STAGE_BYTES + 256 /*barriers*/ + 1024 /*alignment slack*/;
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
  // optional, network 0 only: sample the action from Normal(out, exp(logstd)) in the same launch (null action = off)
  const float* logstd;
  const uint32_t* counter;   // device word: call index of the sampling stream (the caller advances it)
  uint32_t call_offset, seed_lo, seed_hi;
  float* action;             // f32 [M, n_out]
  float* logprob;            // f32 [M]
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

// ---- the schedule ---------------------------------------------------------------------------------------------
// Six accumulator tasks t = 0..5: (layer, chunk of 256 output columns) = (0,0) (1,0) (1,1) (2,0) (2,1) (3,0); task t
// accumulates in TMEM half t & 1 (columns [256 (t & 1), +256)), so the MMAs of task t + 1 run while the epilogue
// drains task t. The activation buffer has 8 k-blocks (128 rows x 64 columns); the input of layer l lives in
//   layer 0: block 0 (the observation tile);   layer 1: logical k-block j -> block j;
//   layers 2, 3: logical k-block j -> block (j + 4) % 8
// so that the epilogue of a chunk always writes blocks the MMAs still in flight do not read - except task (2,0),
// whose output replaces the blocks task (2,1) is reading: it waits per block (blk_free) for those reads to retire.
// The MMAs of the next layer start on a k-block as soon as the epilogue has written it (blk_ready per block).
__host__ __device__ constexpr int fm_task_layer(int t) { return t == 0 ? 0 : (t <= 2 ? 1 : (t <= 4 ? 2 : 3)); }
__host__ __device__ constexpr int fm_task_chunk(int t) { return (t == 2 || t == 4) ? 1 : 0; }
__device__ __forceinline__ int fm_phys(int layer, int kb) { return layer <= 1 ? kb : ((kb + 4) & 7); }
// does layer l read block p?
__device__ __forceinline__ int fm_reads(int l, int p) { return l == 0 ? (p == 0) : (l == 1 ? (p < 4) : 1); }

template <int EW>  // epilogue warps: 4 (one per TMEM lane quarter) or 8 (two per quarter, half of the columns each)
__global__ void __launch_bounds__(64 + 32 * EW, 1)
k_mlp_fwd(const __grid_constant__ FusedArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* act = smem;
  uint8_t* ring = smem + FM_ACT_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + FM_STAGES * FM_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + FM_STAGES;
  uint64_t* x_full = empty_bar + FM_STAGES;   // the observation tile has landed in block 0
  uint64_t* acc_full = x_full + 1;            // [2] every MMA of the task in this TMEM half has retired
  uint64_t* acc_free = acc_full + 2;          // [2] every epilogue warp has read this TMEM half out
  uint64_t* blk_ready = acc_free + 2;         // [8] the epilogue has written this block (next layer's A operand)
  uint64_t* blk_free = blk_ready + 8;         // [8] the MMAs of the layer that read this block have retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(blk_free + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const FusedNet& net = g.net[blockIdx.y];
  const int m0 = blockIdx.x * BM;
  if (threadIdx.x == 64) stamp(g, 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < FM_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(x_full, 1);
    for (int h = 0; h < 2; ++h) { mbar_init(&acc_full[h], 1); mbar_init(&acc_free[h], EW); }
    for (int b = 0; b < 8; ++b) { mbar_init(&blk_ready[b], EW); mbar_init(&blk_free[b], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 64) stamp(g, 1);

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer: the observation tile, then every weight stage in the order the MMAs consume them
      mbar_expect_tx(x_full, FM_BLOCK_BYTES);
      tma_load_2d(act, &g.x, x_full, 0, m0);
      uint32_t s = 0, ph = 0;
#pragma unroll 1
      for (int t = 0; t < 6; ++t) {
        const int l = fm_task_layer(t), ch = fm_task_chunk(t), kbs = fm_k(l) / BK;
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
      for (int t = 0; t < 6; ++t) {
        const int l = fm_task_layer(t), ch = fm_task_chunk(t), kbs = fm_k(l) / BK, h = t & 1;
        const bool last_chunk = (ch + 1) * FM_NCHUNK == fm_n(l);
        // the epilogue has drained the task that used this TMEM half before (t - 2); passes at once for t < 2
        mbar_wait(&acc_free[h], (((uint32_t)t >> 1) & 1u) ^ 1u);
        if (l == 0) mbar_wait(x_full, 0);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(h * FM_NCHUNK);
        for (int kb = 0; kb < kbs; ++kb) {
          const int p = fm_phys(l, kb);
          if (l > 0) {  // written by the epilogue of layer l - 1: completion number l (blocks 0-3) or l - 1 (blocks 4-7)
            mbar_wait(&blk_ready[p], (uint32_t)((p < 4 ? l : l - 1) - 1) & 1u);
          }
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint64_t adesc = make_desc_k128(act_addr + p * FM_BLOCK_BYTES);
          const uint64_t bdesc = make_desc_k128(ring_addr + s * FM_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty_bar[s]);                 // frees the ring stage when these MMAs retire
          if (last_chunk) umma_commit(&blk_free[p]);  // ... and the activation block: no later MMA of this layer reads it
          if (++s == FM_STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&acc_full[h]);
      }
    }
  } else {
    // ---- epilogue: this thread owns row q * 32 + lane of the block (TMEM lane). All epilogue warps work on the SAME
    // 64-column k-block at a time (EW = 8: the two warps of a lane quarter take 32 columns of it each; EW = 4: one warp
    // takes both halves), so the blocks become ready one after the other in the order the next layer's MMAs consume
    // them, and only the last block's MMAs remain when the epilogue of a chunk ends.
    constexpr int CPB = 8 / EW;          // 32-column TMEM loads per k-block per warp
    constexpr int NLD = 4 * CPB;         // ... per 256-column chunk
    const int q = warp & 3, half = (warp - 2) >> 2, row = q * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    const int col_of_warp = CPB == 1 ? 32 * half : 0;
    float hacc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int t = 0; t < 6; ++t) {
      const int l = fm_task_layer(t), ch = fm_task_chunk(t), h = t & 1;
      const int c0 = ch * FM_NCHUNK + col_of_warp;  // output column of this warp's first load in this task
      const float* __restrict__ bias = net.bias[l];
      mbar_wait(&acc_full[h], ((uint32_t)t >> 1) & 1u);
      tc_fence_after();
      if (threadIdx.x == 64 && ch == 0) stamp(g, 2 + 2 * l);
      const uint32_t tc0 = t_row + (uint32_t)(h * FM_NCHUNK + col_of_warp);
      constexpr int STEP = CPB == 1 ? 64 : 32;   // column distance between this warp's consecutive loads
      uint32_t v[2][32];
      tmem_ld32_issue(tc0, v[0]);
#pragma unroll
      for (int i = 0; i < NLD; ++i) {
        tmem_ld_wait();
        if (i + 1 < NLD) {
          tmem_ld32_issue(tc0 + (uint32_t)(STEP * (i + 1)), v[(i + 1) & 1]);
        } else {  // this warp has read its share of the TMEM half: hand it back before the arithmetic
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_free[h]);
        }
        const int c = c0 + STEP * i;
        if (l < 3) {
          const int p = fm_phys(l + 1, c >> 6);
          if (i % CPB == 0) {
            // the layers up to l that read block p have all retired their MMAs (completion number n; n = 0: never read)
            int n = 0;
            for (int l2 = 0; l2 <= l; ++l2) n += fm_reads(l2, p);
            mbar_wait(&blk_free[p], (uint32_t)(n - 1) & 1u);
          }
          uint32_t packed[16];
          bias_tanh_pack(v[i & 1], bias + c, packed);
          uint8_t* base = act + p * FM_BLOCK_BYTES + row * 128;
          const int j0 = (c & 63) >> 3, sw = row & 7;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(base + (((j0 + j) ^ sw) << 4)) =
                make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          if (i % CPB == CPB - 1) {  // this warp's share of the block is written: generic-proxy writes -> visible to the tensor core
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&blk_ready[p]);
          }
        } else if (net.n_out == 1) head_chunk<1>(v[i & 1], bias, net.head_w, c, hacc);
        else if (net.n_out == 2) head_chunk<2>(v[i & 1], bias, net.head_w, c, hacc);
        else head_chunk<6>(v[i & 1], bias, net.head_w, c, hacc);
      }
      if (threadIdx.x == 64 && (ch + 1) * FM_NCHUNK == fm_n(l)) stamp(g, 3 + 2 * l);
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
      const long long i = m0 + row;
#pragma unroll
      for (int a = 0; a < 6; ++a)
        if (a < net.n_out) hacc[a] += __ldg(net.head_b + a);
      if (net.out) {
        float* o = net.out + (size_t)i * net.n_out;
#pragma unroll
        for (int a = 0; a < 6; ++a)
          if (a < net.n_out) o[a] = hacc[a];
      }
      if (blockIdx.y == 0 && g.action) {  // Agent.get_action_and_value: probs.sample(), log_prob(action).sum(1)
        const uint32_t call = __ldg(g.counter) + g.call_offset;
        if (net.n_out == 2) {
          const float mu[2] = {hacc[0], hacc[1]};
          float act[2];
          g.logprob[i] = ppo::sample_row<2>(mu, g.logstd, i, call, g.seed_lo, g.seed_hi, act);
          *reinterpret_cast<float2*>(g.action + 2 * i) = make_float2(act[0], act[1]);
        } else {
          const float mu[6] = {hacc[0], hacc[1], hacc[2], hacc[3], hacc[4], hacc[5]};
          float act[6];
          g.logprob[i] = ppo::sample_row<6>(mu, g.logstd, i, call, g.seed_lo, g.seed_hi, act);
#pragma unroll
          for (int a = 0; a < 6; ++a) g.action[6 * i + a] = act[a];
        }
      }
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
VSS_API int vss_mlp_forward_fused(const void* x16, int ldx, int M, const vss_mlp_net* nets, int n_nets,
                                  const vss_mlp_sampling* sampling, int epilogue_warps, void* stream) {
  return vss_mlp_forward_fused_timed(x16, ldx, M, nets, n_nets, sampling, epilogue_warps, nullptr, stream);
}

// The same launch; the first CTA also records 11 globaltimer stamps (ns) at its phase boundaries into `stamps`
// (device memory): [0] entry, [1] barriers + TMEM ready, [2 + 2l] accumulator of layer l complete, [3 + 2l] epilogue of
// layer l done (one epilogue warp's view), [10] exit. A profiling hook: profiles/mlp_fused_bench.py.
VSS_API int vss_mlp_forward_fused_timed(const void* x16, int ldx, int M, const vss_mlp_net* nets, int n_nets,
                                        const vss_mlp_sampling* sampling, int epilogue_warps, unsigned long long* stamps,
                                        void* stream) {
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
    if (!s.head_w || !s.head_b || (!s.out && !(i == 0 && sampling)) || (s.n_out != 1 && s.n_out != 2 && s.n_out != 6)) {
      g_tc_error = "vss_mlp_forward_fused: bad network (head 256 -> 1, 2 or 6; out may be NULL only for a sampled network 0)";
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
  if (sampling) {
    if (!sampling->logstd || !sampling->counter || !sampling->action || !sampling->logprob ||
        (nets[0].n_out != 2 && nets[0].n_out != 6) || (reinterpret_cast<uintptr_t>(sampling->action) & 7u)) {
      g_tc_error = "vss_mlp_forward_fused: sampling needs logstd, counter, action (8-byte aligned), logprob and an action width of 2 or 6";
      return VSS_E_INVALID;
    }
    a.logstd = sampling->logstd; a.counter = sampling->counter; a.call_offset = sampling->call_offset;
    a.seed_lo = (uint32_t)sampling->seed; a.seed_hi = (uint32_t)(sampling->seed >> 32);
    a.action = sampling->action; a.logprob = sampling->logprob;
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
