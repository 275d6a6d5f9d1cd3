// Microbenchmark: FFMA vs FFMA2 (fma.rn.f32x2) issue throughput on sm_100a, 8 warps per SM, 48 accumulators per thread.
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                     rc = *reinterpret_cast<unsigned long long *>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2 *>(&rd);
}
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(const float *in, float *out, int iters) {
  float2 acc[24];
  for (int i = 0; i < 24; ++i) acc[i] = make_float2(in[i], in[i + 1]);
  float2 a = make_float2(in[threadIdx.x & 7], in[(threadIdx.x & 7) + 1]), w = make_float2(in[3], in[4]);
  for (int j = 0; j < iters; ++j) {
#pragma unroll
    for (int i = 0; i < 24; ++i) {
      if (MODE == 0) { acc[i].x = fmaf(a.x, w.x, acc[i].x); acc[i].y = fmaf(a.y, w.y, acc[i].y); }
      else acc[i] = ffma2(a, w, acc[i]);
    }
    a.x += 1e-9f; w.y -= 1e-9f;
  }
  float s = 0.f;
  for (int i = 0; i < 24; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float *in, *out;
  cudaMalloc(&in, 4096); cudaMalloc(&out, 148 * 256 * 4);
  cudaMemset(in, 0, 4096);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 100000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148, 256>>>(in, out, iters); else k<1><<<148, 256>>>(in, out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = 148.0 * 256 * 48.0 * iters;
      printf("mode %s: %.3f ms, %.1f TFLOP/s, %.1f FMA/clk/SM @1.965GHz\n", mode ? "FFMA2" : "FFMA ", ms, 2 * fma / ms * 1e-9, fma / 148 / (ms * 1e-3 * 1.965e9));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
