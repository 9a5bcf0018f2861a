// micro-probe: scalar FFMA vs packed f32x2 FFMA2 on sm_100a -- throughput (ILP 8) and dependent-chain latency
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, bool PACKED>
__global__ void k(float* out, int iters, float a, float b) {
  float2 acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
  const float2 A = make_float2(a, a * 1.0001f), B = make_float2(b, b * 0.9999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (PACKED) acc[i] = __ffma2_rn(acc[i], A, B);
      else { acc[i].x = fmaf(acc[i].x, A.x, B.x); acc[i].y = fmaf(acc[i].y, A.y, B.y); }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP, bool PACKED>
void run(const char* name, int warps_per_sm) {
  float* d; cudaMalloc(&d, 148 * 1024 * 4 * 4);
  const int iters = 20000;
  dim3 grid(148), block(warps_per_sm * 32);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ILP, PACKED><<<grid, block>>>(d, 100, 0.999f, 0.001f);
  cudaEventRecord(e0); k<ILP, PACKED><<<grid, block>>>(d, iters, 0.999f, 0.001f); cudaEventRecord(e1);
  cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fma = 2.0 * ILP * (double)iters * 148 * warps_per_sm * 32;   // scalar-equivalent FMAs
  printf("%-28s warps/SM %2d  ILP %d : %.3f ms  %.2f TFMA/s (x2 = TFLOP/s)\n", name, warps_per_sm, ILP, ms, fma / ms / 1e9);
  cudaFree(d);
}
int main() {
  for (int w : {4, 8, 16, 32}) {
    run<1, false>("scalar dependent chain", w); run<1, true>("packed dependent chain", w);
    run<8, false>("scalar ILP8", w); run<8, true>("packed ILP8", w);
  }
  return 0;
}
