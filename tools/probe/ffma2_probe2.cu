// micro-probe 2: FFMA2 / FFMA throughput when every instruction reads THREE DISTINCT register operands (the situation
// of the EQ scan kernels: y = b0*u + s1 has three live operands), versus the accumulate form acc = acc*A + B with two
// loop-invariant operands that tools/probe/ffma2_probe.cu measured.  ILP 8, 148 CTAs, 4..32 warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
template <bool PACKED>
__global__ void k3(float* out, int iters, float s) {
  float2 a[8], b[8], c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
    b[i] = make_float2(0.5f + 1e-3f * i, 0.25f + s * i);
    c[i] = make_float2(1e-3f * i, s);
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (PACKED) a[i] = __ffma2_rn(b[i], c[(i + 1) & 7], a[i]);
      else { a[i].x = fmaf(b[i].x, c[(i + 1) & 7].x, a[i].x); a[i].y = fmaf(b[i].y, c[(i + 1) & 7].y, a[i].y); }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (PACKED) b[i] = __ffma2_rn(a[i], c[i], b[(i + 3) & 7]);
      else { b[i].x = fmaf(a[i].x, c[i].x, b[(i + 3) & 7].x); b[i].y = fmaf(a[i].y, c[i].y, b[(i + 3) & 7].y); }
    }
  }
  float r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y + b[i].x + b[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// FADD / FADD2 (two distinct source pairs) and the accumulate form with two loop-invariant operands
template <int MODE>   // 0 scalar add, 1 packed add, 2 scalar fma invariant operands, 3 packed fma invariant operands
__global__ void k2(float* out, int iters, float s) {
  float2 a[8], b[8];
  const float2 A = make_float2(0.999f, 0.9991f + s), B = make_float2(s, 2 * s);
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i); b[i] = make_float2(1e-3f * i, s * i); }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { a[i].x = a[i].x + b[(i + 1) & 7].x; a[i].y = a[i].y + b[(i + 1) & 7].y; }
      if (MODE == 1) a[i] = __fadd2_rn(a[i], b[(i + 1) & 7]);
      if (MODE == 2) { a[i].x = fmaf(a[i].x, A.x, B.x); a[i].y = fmaf(a[i].y, A.y, B.y); }
      if (MODE == 3) a[i] = __ffma2_rn(a[i], A, B);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { b[i].x = b[i].x + a[(i + 3) & 7].x; b[i].y = b[i].y + a[(i + 3) & 7].y; }
      if (MODE == 1) b[i] = __fadd2_rn(b[i], a[(i + 3) & 7]);
      if (MODE == 2) { b[i].x = fmaf(b[i].x, A.x, B.x); b[i].y = fmaf(b[i].y, A.y, B.y); }
      if (MODE == 3) b[i] = __ffma2_rn(b[i], A, B);
    }
  }
  float r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y + b[i].x + b[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run2(const char* name, int warps_per_sm) {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  const int iters = 10000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k2<MODE><<<148, warps_per_sm * 32>>>(d, 100, 1e-6f);
  cudaEventRecord(e0); k2<MODE><<<148, warps_per_sm * 32>>>(d, iters, 1e-6f); cudaEventRecord(e1);
  cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);
  const bool packed = MODE & 1;
  const double ops = 2.0 * 16 * (double)iters * 148 * warps_per_sm * 32;
  const double cyc_per_inst = (ms * 1e-3 * 1.965e9) / ((packed ? 16.0 : 32.0) * iters * warps_per_sm / 4.0);
  printf("%-34s warps/SM %2d : %.3f ms  %.2f Tops/s  ~%.2f cycles per warp-instruction per SMSP @1965 MHz\n", name,
         warps_per_sm, ms, ops / ms / 1e9, cyc_per_inst);
  cudaFree(d);
}
template <bool PACKED>
void run(const char* name, int warps_per_sm) {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  const int iters = 10000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k3<PACKED><<<148, warps_per_sm * 32>>>(d, 100, 1e-6f);
  cudaEventRecord(e0); k3<PACKED><<<148, warps_per_sm * 32>>>(d, iters, 1e-6f); cudaEventRecord(e1);
  cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fma = 2.0 * 16 * (double)iters * 148 * warps_per_sm * 32;    // scalar-equivalent FMAs
  const double cyc_per_inst = (ms * 1e-3 * 1.965e9) / ((PACKED ? 16.0 : 32.0) * iters * warps_per_sm / 4.0);
  printf("%-34s warps/SM %2d : %.3f ms  %.2f TFMA/s  ~%.2f cycles per warp-instruction per SMSP @1965 MHz\n", name,
         warps_per_sm, ms, fma / ms / 1e9, cyc_per_inst);
  cudaFree(d);
}
int main() {
  for (int w : {4, 8, 16, 32}) { run<false>("scalar FFMA, 3 distinct operands", w); run<true>("packed FFMA2, 3 distinct operands", w); }
  for (int w : {8, 32}) {
    run2<0>("scalar FADD, 2 distinct operands", w); run2<1>("packed FADD2, 2 distinct operands", w);
    run2<2>("scalar FFMA, invariant a*A+B", w); run2<3>("packed FFMA2, invariant a*A+B", w);
  }
  return 0;
}
