// peer_reduce.cu — the PPO gradient all-reduce as ONE kernel over NVLink peer memory (sm_100a).
//
// The only collective of the path is the sum of the flat gradient (≈ 1.08 M floats,
// 4.3 MB) over the ranks, 32 times per update, between a minibatch's backward pass and its
// clip + Adam step. Through NCCL that exchange costs ~140 us per minibatch at 8 GPUs (PPO-sa update
// 43.0 -> 47.6 ms). Here every rank keeps its flat gradient in a buffer that the other ranks of the
// node map through CUDA IPC, and one kernel per rank
//   1. tells every peer "my gradient of step e is complete" (a release store of e into the peer's
//      flag word for this rank) and waits until all peers have said so,
//   2. reads the same element of all P buffers over NVLink (16-byte loads, fixed rank order, so all
//      ranks compute bit-identical sums) and writes the sum into a local buffer,
//   3. tells every peer "I have read your gradient of step e" and waits until all peers have read
//      its own — after the kernel the gradient buffer may be overwritten by the next backward pass.
// No host involvement, graph-capturable (the step counter e lives on the device). NVSwitch gives
// every GPU full bandwidth to every peer, so the one-shot form (P - 1 remote reads per element,
// 30 MB per rank at 8 GPUs) is bound by latency, not by links. Every wait is bounded and traps.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <new>
#include <string>

#include "../../include/vss_b200.h"

namespace peer {

constexpr int MAX_RANKS = 16;
constexpr int FLAG_WORDS = 2 * MAX_RANKS;     // [0, MAX) "gradient ready" flags, [MAX, 2 MAX) "gradient read" flags

struct Args {
  const float* grad[MAX_RANKS];               // every rank's gradient buffer (own = local pointer)
  uint32_t* flags[MAX_RANKS];                 // every rank's flag words
  float* out;                                 // local: the sum
  long long n4;                               // float4 elements
  int rank, world;
  uint32_t* epoch;                            // local device counter of calls
  uint32_t* grid_ctr;                         // local: CTAs that have finished reading
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float* p) {  // system-scope load: never served from a stale L1 line
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
// bounded wait until flag word `p` has reached epoch `e` (wrap-around safe)
__device__ __forceinline__ void wait_flag(const uint32_t* p, uint32_t e, int what) {
  for (uint32_t i = 0; i < (1u << 26); ++i) {
    if ((int32_t)(ld_acquire_sys(p) - e) >= 0) return;
    __nanosleep(64);
  }
  printf("vss_peer_allreduce: wait %d timed out (block %d)\n", what, blockIdx.x);
  __trap();
}

__global__ void __launch_bounds__(256)
k_peer_allreduce(const __grid_constant__ Args a) {
  __shared__ uint32_t s_epoch;
  const int tid = threadIdx.x;
  if (tid == 0) s_epoch = *reinterpret_cast<volatile uint32_t*>(a.epoch) + 1u;  // (advanced by the last CTA, below)
  __syncthreads();
  const uint32_t e = s_epoch;
  // 1. my gradient is complete (this kernel is stream-ordered after the backward pass): tell the peers
  if (blockIdx.x == 0 && tid < a.world && tid != a.rank) {
    __threadfence_system();
    st_release_sys(a.flags[tid] + a.rank, e);
  }
  if (tid < a.world && tid != a.rank) wait_flag(a.flags[a.rank] + tid, e, 1);
  __syncthreads();
  // 2. the sum, in rank order
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < a.n4; i += stride) {
    float4 s = ld_peer(a.grad[0] + 4 * i);
    for (int p = 1; p < a.world; ++p) {
      const float4 v = ld_peer(a.grad[p] + 4 * i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<float4*>(a.out)[i] = s;
  }
  // 3. the last CTA to finish: "rank r has read everybody", then wait until everybody has read rank r
  __syncthreads();
  __shared__ bool s_last;
  if (tid == 0) s_last = atomicAdd(a.grid_ctr, 1u) == gridDim.x - 1u;
  __syncthreads();
  if (!s_last) return;
  if (tid < a.world && tid != a.rank) {
    st_release_sys(a.flags[tid] + MAX_RANKS + a.rank, e);
    wait_flag(a.flags[a.rank] + MAX_RANKS + tid, e, 2);
  }
  __syncthreads();
  if (tid == 0) { *a.grid_ctr = 0u; *a.epoch = e; }
}

}  // namespace peer

struct vss_peer_group {
  int device, rank, world;
  int64_t n;                    // floats (padded to a multiple of 4)
  float* region;                // own allocation: n floats + flag words + epoch + grid counter
  void* peer_base[peer::MAX_RANKS];
  peer::Args args;
  bool connected;
};

static thread_local std::string g_peer_error;
static int pfail(int code, const char* what, cudaError_t e = cudaSuccess) {
  g_peer_error = what;
  if (e != cudaSuccess) { g_peer_error += ": "; g_peer_error += cudaGetErrorString(e); }
  return code;
}
static size_t region_bytes(int64_t n) { return sizeof(float) * (size_t)n + sizeof(uint32_t) * (peer::FLAG_WORDS + 4); }

extern "C" {

VSS_API const char* vss_peer_last_error(void) { return g_peer_error.c_str(); }

VSS_API int vss_peer_create(vss_peer* out, int device, int rank, int world, int64_t num_floats) {
  if (!out || rank < 0 || world < 2 || world > peer::MAX_RANKS || rank >= world || num_floats <= 0)
    return pfail(VSS_E_INVALID, "vss_peer_create: bad argument (2 <= world <= 16)");
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return pfail(VSS_E_CUDA, "vss_peer_create: cudaSetDevice", e);
  vss_peer_group* h = new (std::nothrow) vss_peer_group();
  if (!h) return pfail(VSS_E_NOMEM, "vss_peer_create: host allocation failed");
  memset(h, 0, sizeof(*h));
  h->device = device; h->rank = rank; h->world = world; h->n = (num_floats + 3) / 4 * 4;
  e = cudaMalloc(&h->region, region_bytes(h->n));   // (cudaMalloc memory can be exported through CUDA IPC)
  if (e == cudaSuccess) e = cudaMemset(h->region, 0, region_bytes(h->n));
  if (e != cudaSuccess) { cudaFree(h->region); delete h; return pfail(VSS_E_NOMEM, "vss_peer_create: cudaMalloc", e); }
  *out = h;
  return VSS_OK;
}

VSS_API int vss_peer_ipc_handle(vss_peer h, void* handle64) {
  if (!h || !handle64) return pfail(VSS_E_INVALID, "vss_peer_ipc_handle: null");
  static_assert(sizeof(cudaIpcMemHandle_t) == VSS_PEER_HANDLE_BYTES, "handle size");
  cudaIpcMemHandle_t hd;
  cudaError_t e = cudaIpcGetMemHandle(&hd, h->region);
  if (e != cudaSuccess) return pfail(VSS_E_CUDA, "vss_peer_ipc_handle: cudaIpcGetMemHandle", e);
  memcpy(handle64, &hd, sizeof(hd));
  return VSS_OK;
}

VSS_API int vss_peer_connect(vss_peer h, const void* all_handles) {
  if (!h || !all_handles) return pfail(VSS_E_INVALID, "vss_peer_connect: null");
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess) return pfail(VSS_E_CUDA, "vss_peer_connect: cudaSetDevice", e);
  for (int p = 0; p < h->world; ++p) {
    void* base = h->region;
    if (p != h->rank) {
      cudaIpcMemHandle_t hd;
      memcpy(&hd, static_cast<const char*>(all_handles) + (size_t)p * VSS_PEER_HANDLE_BYTES, sizeof(hd));
      e = cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) return pfail(VSS_E_CUDA, "vss_peer_connect: cudaIpcOpenMemHandle (no peer access between these GPUs?)", e);
      h->peer_base[p] = base;
    }
    h->args.grad[p] = static_cast<const float*>(base);
    h->args.flags[p] = reinterpret_cast<uint32_t*>(static_cast<float*>(base) + h->n);
  }
  h->args.n4 = h->n / 4; h->args.rank = h->rank; h->args.world = h->world;
  h->args.epoch = h->args.flags[h->rank] + peer::FLAG_WORDS;
  h->args.grid_ctr = h->args.epoch + 1;
  h->connected = true;
  return VSS_OK;
}

VSS_API float* vss_peer_buffer(vss_peer h) { return h ? h->region : nullptr; }
VSS_API int64_t vss_peer_num_floats(vss_peer h) { return h ? h->n : 0; }

VSS_API int vss_peer_allreduce(vss_peer h, float* out_sum, void* stream) {
  if (!h || !out_sum) return pfail(VSS_E_INVALID, "vss_peer_allreduce: null");
  if (!h->connected) return pfail(VSS_E_INVALID, "vss_peer_allreduce: vss_peer_connect has not been called");
  if (reinterpret_cast<uintptr_t>(out_sum) & 15u) return pfail(VSS_E_INVALID, "vss_peer_allreduce: out_sum must be 16-byte aligned");
  peer::Args a = h->args;
  a.out = out_sum;
  // one CTA per SM at most (all resident: the kernel's last phase waits on other GPUs), 1024 elements per CTA at least
  const long long want = (a.n4 + 1023) / 1024;
  const unsigned grid = (unsigned)(want < 1 ? 1 : (want > 148 ? 148 : want));
  peer::k_peer_allreduce<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return pfail(VSS_E_CUDA, "vss_peer_allreduce: launch", e);
  return VSS_OK;
}

VSS_API int vss_peer_destroy(vss_peer h) {
  if (!h) return VSS_OK;
  cudaSetDevice(h->device);
  for (int p = 0; p < h->world; ++p)
    if (p != h->rank && h->peer_base[p]) cudaIpcCloseMemHandle(h->peer_base[p]);
  cudaFree(h->region);
  delete h;
  return VSS_OK;
}

}  // extern "C"
