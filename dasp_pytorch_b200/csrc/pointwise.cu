// gain / distortion forward + backward (reference: dasp_pytorch/functional.py:10-29, 65-78).
//
// Row model: a "row" is the contiguous run of samples that shares one parameter value.
//   gain       : row = one batch item   (chs*N samples), parameter gain_db[b]
//   distortion : row = one (b, c) pair  (N samples),     parameter drive_db[b*chs + c]
// Roofline: pure HBM streaming.  fwd 8 B/sample (read x, write y); bwd 12 B/sample (read g,
// read x, write gx) -- tanh is recomputed from x in the backward instead of re-reading y so
// the backward stays at 12 B/sample and needs x anyway for the drive gradient.
// Loads/stores are 128-bit, streaming (no L1 allocate), 4 independent vectors in flight per
// thread.  The parameter gradient is reduced deterministically: one partial per CTA into a
// caller workspace, then a second tiny kernel sums the partials of each row (no atomics, so
// 1-GPU and sharded N-GPU runs are bit-identical per item).
#include "common.cuh"
#include "../../include/dasp_b200.h"

namespace dasp {
namespace {

constexpr int kThreads = 256;
constexpr int kVecPerThread = 4;                              // float4 per thread per tile
constexpr int kTile = kThreads * kVecPerThread * 4;           // 4096 samples per CTA

enum class Op { Gain, Tanh };

__device__ __forceinline__ float4 ld_stream(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st_stream(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

template <Op OP>
__device__ __forceinline__ float fwd1(float x, float k) {
  if (OP == Op::Gain) return x * k;
  return tanhf(x * k);   // accurate tanh: tanh.approx (5e-4) would break the 1e-4 parity gate
}

// grid = rows * tiles (flat: CTA id = row * tiles + tile)
template <Op OP, bool VEC>
__global__ void __launch_bounds__(kThreads) pointwise_fwd_kernel(const float* __restrict__ x,
                                                                 const float* __restrict__ p_db,
                                                                 float* __restrict__ y, int64_t n, int tiles) {
  const int64_t row = blockIdx.x / tiles;
  const int tile = blockIdx.x - (unsigned)(row * tiles);
  const float k = db_to_lin(__ldg(p_db + row));
  const float* xr = x + row * n;
  float* yr = y + row * n;
  const int64_t t0 = (int64_t)tile * kTile;
  if (VEC) {
    float4 v[kVecPerThread];
#pragma unroll
    for (int j = 0; j < kVecPerThread; ++j) {
      int64_t i = t0 + (int64_t)(j * kThreads + threadIdx.x) * 4;
      if (i < n) v[j] = ld_stream(xr + i);
    }
#pragma unroll
    for (int j = 0; j < kVecPerThread; ++j) {
      int64_t i = t0 + (int64_t)(j * kThreads + threadIdx.x) * 4;
      if (i < n) {
        float4 o;
        o.x = fwd1<OP>(v[j].x, k); o.y = fwd1<OP>(v[j].y, k);
        o.z = fwd1<OP>(v[j].z, k); o.w = fwd1<OP>(v[j].w, k);
        st_stream(yr + i, o);
      }
    }
  } else {
    for (int j = threadIdx.x; j < kTile; j += kThreads) {
      int64_t i = t0 + j;
      if (i < n) yr[i] = fwd1<OP>(xr[i], k);
    }
  }
}

// per element: returns gx, accumulates the parameter-gradient integrand into acc
template <Op OP>
__device__ __forceinline__ float bwd1(float g, float x, float k, float& acc) {
  if (OP == Op::Gain) {
    acc = fmaf(g, x, acc);            // d/dk (k x) = x ; chain to dB in the reduce kernel
    return g * k;
  }
  float y = tanhf(x * k);
  float d = g * fmaf(-y, y, 1.0f);    // g (1 - y^2)
  acc = fmaf(d, x, acc);
  return d * k;
}

template <Op OP, bool VEC>
__global__ void __launch_bounds__(kThreads) pointwise_bwd_kernel(const float* __restrict__ gy,
                                                                 const float* __restrict__ x,
                                                                 const float* __restrict__ p_db,
                                                                 float* __restrict__ gx,
                                                                 float* __restrict__ partials, int64_t n,
                                                                 int tiles) {
  const int64_t row = blockIdx.x / tiles;
  const int tile = blockIdx.x - (unsigned)(row * tiles);
  const float k = db_to_lin(__ldg(p_db + row));
  const float* xr = x + row * n;
  const float* gr = gy + row * n;
  float* gxr = gx + row * n;
  const int64_t t0 = (int64_t)tile * kTile;
  float acc = 0.f;
  if (VEC) {
    float4 vx[kVecPerThread], vg[kVecPerThread];
#pragma unroll
    for (int j = 0; j < kVecPerThread; ++j) {
      int64_t i = t0 + (int64_t)(j * kThreads + threadIdx.x) * 4;
      if (i < n) { vx[j] = ld_stream(xr + i); vg[j] = ld_stream(gr + i); }
    }
#pragma unroll
    for (int j = 0; j < kVecPerThread; ++j) {
      int64_t i = t0 + (int64_t)(j * kThreads + threadIdx.x) * 4;
      if (i < n) {
        float4 o;
        o.x = bwd1<OP>(vg[j].x, vx[j].x, k, acc); o.y = bwd1<OP>(vg[j].y, vx[j].y, k, acc);
        o.z = bwd1<OP>(vg[j].z, vx[j].z, k, acc); o.w = bwd1<OP>(vg[j].w, vx[j].w, k, acc);
        st_stream(gxr + i, o);
      }
    }
  } else {
    for (int j = threadIdx.x; j < kTile; j += kThreads) {
      int64_t i = t0 + j;
      if (i < n) gxr[i] = bwd1<OP>(gr[i], xr[i], k, acc);
    }
  }
  __shared__ float warp_part[kThreads / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) s += warp_part[w];
    partials[blockIdx.x] = s;   // == row * tiles + tile
  }
}

// one warp per row: g_param[row] = ln10/20 * k * sum(partials[row, :])
__global__ void reduce_param_grad_kernel(const float* __restrict__ partials, const float* __restrict__ p_db,
                                         float* __restrict__ g_param, int64_t rows, int tiles) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  for (int t = lane; t < tiles; t += 32) s += (double)partials[row * tiles + t];
  s = warp_sum(s);
  if (lane == 0) g_param[row] = (float)(s * (double)kLn10Over20 * (double)db_to_lin(p_db[row]));
}

template <Op OP>
int launch_fwd(const float* x, const float* p_db, float* y, int64_t rows, int64_t n, cudaStream_t st) {
  DASP_REQUIRE(rows >= 0 && n >= 0, "pointwise fwd: negative size");
  if (rows == 0 || n == 0) return DASP_OK;          // empty tensors have null data pointers
  DASP_REQUIRE(x && p_db && y, "pointwise fwd: null pointer");
  const int64_t tiles64 = (n + kTile - 1) / kTile;
  DASP_REQUIRE(rows * tiles64 < (1ll << 31), "pointwise fwd: rows*tiles = %lld exceeds the grid limit",
               (long long)(rows * tiles64));
  const int tiles = (int)tiles64;
  const unsigned grid = (unsigned)(rows * tiles64);
  const bool vec = (n % 4 == 0) && aligned16(x) && aligned16(y);
  if (vec) pointwise_fwd_kernel<OP, true><<<grid, kThreads, 0, st>>>(x, p_db, y, n, tiles);
  else     pointwise_fwd_kernel<OP, false><<<grid, kThreads, 0, st>>>(x, p_db, y, n, tiles);
  DASP_LAUNCH_OK("pointwise_fwd_kernel");
  return DASP_OK;
}

template <Op OP>
int launch_bwd(const float* gy, const float* x, const float* p_db, float* gx, float* g_param, float* ws,
               int64_t ws_floats, int64_t rows, int64_t n, cudaStream_t st) {
  DASP_REQUIRE(rows >= 0 && n >= 0, "pointwise bwd: negative size");
  if (rows == 0) return DASP_OK;
  DASP_REQUIRE(g_param != nullptr, "pointwise bwd: null g_param");
  if (n == 0) { DASP_CUDA_OK(cudaMemsetAsync(g_param, 0, sizeof(float) * rows, st)); return DASP_OK; }
  DASP_REQUIRE(gy && x && p_db && gx, "pointwise bwd: null pointer");
  const int64_t tiles64 = (n + kTile - 1) / kTile;
  DASP_REQUIRE(rows * tiles64 < (1ll << 31), "pointwise bwd: rows*tiles = %lld exceeds the grid limit",
               (long long)(rows * tiles64));
  const int tiles = (int)tiles64;
  if (ws == nullptr || ws_floats < rows * tiles) {
    set_error("pointwise bwd: workspace needs %lld floats, got %lld", (long long)(rows * tiles),
              (long long)ws_floats);
    return DASP_ERR_WORKSPACE;
  }
  const unsigned grid = (unsigned)(rows * tiles64);
  const bool vec = (n % 4 == 0) && aligned16(x) && aligned16(gy) && aligned16(gx);
  if (vec) pointwise_bwd_kernel<OP, true><<<grid, kThreads, 0, st>>>(gy, x, p_db, gx, ws, n, tiles);
  else     pointwise_bwd_kernel<OP, false><<<grid, kThreads, 0, st>>>(gy, x, p_db, gx, ws, n, tiles);
  DASP_LAUNCH_OK("pointwise_bwd_kernel");
  const int rows_per_block = 8;
  reduce_param_grad_kernel<<<(unsigned)((rows + rows_per_block - 1) / rows_per_block), rows_per_block * 32, 0, st>>>(
      ws, p_db, g_param, rows, tiles);
  DASP_LAUNCH_OK("reduce_param_grad_kernel");
  return DASP_OK;
}

}  // namespace
}  // namespace dasp

using namespace dasp;

extern "C" {

int64_t dasp_pointwise_bwd_workspace_floats(int64_t rows, int64_t n) {
  return rows * ((n + kTile - 1) / kTile);
}

int dasp_gain_fwd(const float* x, const float* gain_db, float* y, int64_t bs, int64_t chs, int64_t n,
                  void* stream) {
  return launch_fwd<Op::Gain>(x, gain_db, y, bs, chs * n, (cudaStream_t)stream);
}
int dasp_gain_bwd(const float* gy, const float* x, const float* gain_db, float* gx, float* g_gain_db,
                  float* ws, int64_t ws_floats, int64_t bs, int64_t chs, int64_t n, void* stream) {
  return launch_bwd<Op::Gain>(gy, x, gain_db, gx, g_gain_db, ws, ws_floats, bs, chs * n, (cudaStream_t)stream);
}
int dasp_distortion_fwd(const float* x, const float* drive_db, float* y, int64_t rows, int64_t n, void* stream) {
  return launch_fwd<Op::Tanh>(x, drive_db, y, rows, n, (cudaStream_t)stream);
}
int dasp_distortion_bwd(const float* gy, const float* x, const float* drive_db, float* gx, float* g_drive_db,
                        float* ws, int64_t ws_floats, int64_t rows, int64_t n, void* stream) {
  return launch_bwd<Op::Tanh>(gy, x, drive_db, gx, g_drive_db, ws, ws_floats, rows, n, (cudaStream_t)stream);
}

}  // extern "C"
