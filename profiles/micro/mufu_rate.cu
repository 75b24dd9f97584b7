// Throughput of MUFU.TANH vs MUFU.EX2 vs MUFU.RCP per SM (lanes per clock), one CTA of 1024 threads per SM,
// 8 independent chains per thread. Build: nvcc -gencode arch=compute_100a,code=sm_100a -o mufu_rate mufu_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__device__ __forceinline__ float f(float x) {
  float y;
  if (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int OP>
__global__ void k(float* out, int iters, long long* cycles) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = 0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = f<OP>(v[i]);
  __syncthreads();
  const long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  const char* names[3] = {"MUFU.TANH", "MUFU.EX2", "MUFU.RCP"};
  for (int op = 0; op < 3; ++op) for (int threads : {128, 256, 1024}) {
    for (int rep = 0; rep < 2; ++rep) {
      if (op == 0) k<0><<<148, threads>>>(out, iters, cyc); else if (op == 1) k<1><<<148, threads>>>(out, iters, cyc); else k<2><<<148, threads>>>(out, iters, cyc);
      cudaDeviceSynchronize();
    }
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    printf("%s, %4d threads per SM: %.2f lanes per clock per SM\n", names[op], threads, (double)threads * 8 * iters / c);
  }
  return 0;
}
