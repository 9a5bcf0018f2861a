// stereo_widener / stereo_panner / stereo_bus forward + backward
// (reference: dasp_pytorch/functional.py:580-604, 607-636, 32-62).  SURVEY.md 8f rank 3: the remaining
// exported processors.  All three are streaming mixes -- HBM bound, 128-bit accesses, the parameter gradient
// reduced deterministically (one partial per CTA, then a second tiny kernel; no atomics).
//
//   widener:  mid/side scaling collapses to  left = L + c R,  right = c L + R,  c = 1 - 2 width
//             (mid=(L+R)/sqrt2 * 2(1-w), side=(L-R)/sqrt2 * 2w, functional.py:592-604)
//   panner :  out[b, 0|1, t, :] = x[b, t, :] * sqrt((pi/2 - th) (2/pi) cos th) | sqrt(th (2/pi) sin th), th = pan pi/2
//   bus    :  out[b, c, :] = sum_t x[b, c, t, :] * 10^(send_db[b, t] / 20)
#include <math.h>

#include "common.cuh"

namespace dasp {
namespace {

constexpr int kThreads = 256;
constexpr int kTile = kThreads * 4 * 2;      // samples per CTA (2 float4 per thread)

__device__ __forceinline__ float4 ld4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ float dot4(float4 a, float4 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }
__device__ __forceinline__ float4 axpby(float a, float4 x, float b, float4 y) {
  return make_float4(fmaf(a, x.x, b * y.x), fmaf(a, x.y, b * y.y), fmaf(a, x.z, b * y.z), fmaf(a, x.w, b * y.w));
}

// block-wide sum of up to two values; thread 0 gets the result
__device__ __forceinline__ void block_sum2(float& a, float& b) {
  __shared__ float wp[2][kThreads / 32];
  a = warp_sum(a); b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) { wp[0][threadIdx.x >> 5] = a; wp[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) { s0 += wp[0][w]; s1 += wp[1][w]; }
    a = s0; b = s1;
  }
}

// generic two-row mixer:  o0 = a00 r0 + a01 r1,  o1 = a10 r0 + a11 r1  over one tile of two rows.
// The backward of widener and panner and both forwards are instances of it; VEC = 16-byte aligned rows.
template <bool VEC, bool REDUCE>
__device__ __forceinline__ void mix_tile(const float* r0, const float* r1, float* o0, float* o1, float a00, float a01,
                                         float a10, float a11, const float* q0, const float* q1, int64_t t0, int64_t n,
                                         float& acc0, float& acc1) {
  // REDUCE: additionally acc0 += sum r0*q0[+..], acc1 += sum r1*q1 (q may alias the other row)
  if (VEC) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int64_t i = t0 + (int64_t)(j * kThreads + threadIdx.x) * 4;
      if (i < n) {
        const float4 v0 = r0 ? ld4(r0 + i) : make_float4(0, 0, 0, 0);
        const float4 v1 = r1 ? ld4(r1 + i) : make_float4(0, 0, 0, 0);
        if (o0) st4(o0 + i, axpby(a00, v0, a01, v1));
        if (o1) st4(o1 + i, axpby(a10, v0, a11, v1));
        if (REDUCE) { acc0 += dot4(v0, ld4(q0 + i)); acc1 += dot4(v1, ld4(q1 + i)); }
      }
    }
  } else {
    for (int j = threadIdx.x; j < kTile; j += kThreads) {
      const int64_t i = t0 + j;
      if (i < n) {
        const float v0 = r0 ? r0[i] : 0.f, v1 = r1 ? r1[i] : 0.f;
        if (o0) o0[i] = fmaf(a00, v0, a01 * v1);
        if (o1) o1[i] = fmaf(a10, v0, a11 * v1);
        if (REDUCE) { acc0 = fmaf(v0, q0[i], acc0); acc1 = fmaf(v1, q1[i], acc1); }
      }
    }
  }
}

// ------------------------------------------------------------------ widener
template <bool VEC, bool BWD>
__global__ void __launch_bounds__(kThreads) widener_kernel(const float* __restrict__ in, const float* __restrict__ x,
                                                           const float* __restrict__ width, float* __restrict__ out,
                                                           float* __restrict__ part, int64_t n, int tiles) {
  // forward: in = x, out = y.  backward: in = gy, out = gx, x = forward input (for d width).
  const int64_t b = blockIdx.x / tiles;
  const int tile = blockIdx.x - (unsigned)(b * tiles);
  const float c = 1.0f - 2.0f * width[b];
  const float* i0 = in + (b * 2) * n;
  const float* i1 = i0 + n;
  float a0 = 0.f, a1 = 0.f;
  // d/dw: y_l = L + c R, y_r = c L + R  =>  dL/dw = -2 sum (g_l R + g_r L)
  mix_tile<VEC, BWD>(i0, i1, out + (b * 2) * n, out + (b * 2 + 1) * n, 1.f, c, c, 1.f,
                     BWD ? x + (b * 2 + 1) * n : nullptr, BWD ? x + (b * 2) * n : nullptr, (int64_t)tile * kTile, n, a0, a1);
  if (BWD) {
    block_sum2(a0, a1);
    if (threadIdx.x == 0) part[blockIdx.x] = -2.0f * (a0 + a1);
  }
}

// ------------------------------------------------------------------ panner
__device__ __forceinline__ void pan_gains(float pan, float& lg, float& rg, float& dlg, float& drg) {
  const float hp = 1.5707963267948966f, tw = 0.6366197723675814f;     // pi/2, 2/pi
  const float th = pan * hp;
  float s, c;
  sincosf(th, &s, &c);
  const float fl = (hp - th) * tw * c, fr = th * tw * s;
  lg = sqrtf(fl); rg = sqrtf(fr);
  // d/dpan = (pi/2) d/dth ; d sqrt(f) = f' / (2 sqrt f)
  dlg = hp * (tw * (-c - (hp - th) * s)) / (2.0f * lg);
  drg = hp * (tw * (s + th * c)) / (2.0f * rg);
}

template <bool VEC, bool BWD>
__global__ void __launch_bounds__(kThreads) panner_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                          const float* __restrict__ pan, float* __restrict__ out,
                                                          float* __restrict__ part, int64_t n, int tracks, int tiles) {
  // rows r = b*tracks + t.  forward: out (bs, 2, tracks, n).  backward: g (bs, 2, tracks, n) -> out = gx (bs, tracks, n)
  const int64_t r = blockIdx.x / tiles;
  const int tile = blockIdx.x - (unsigned)(r * tiles);
  const int64_t b = r / tracks, t = r - b * tracks;
  float lg, rg, dlg, drg;
  pan_gains(pan[r], lg, rg, dlg, drg);
  const int64_t o0 = ((b * 2 + 0) * tracks + t) * n, o1 = ((b * 2 + 1) * tracks + t) * n;
  float a0 = 0.f, a1 = 0.f;
  if (!BWD) {
    mix_tile<VEC, false>(x + r * n, nullptr, out + o0, out + o1, lg, 0.f, rg, 0.f, nullptr, nullptr, (int64_t)tile * kTile,
                         n, a0, a1);
  } else {
    // gx = lg g0 + rg g1 ; partials S0 = sum g0 x, S1 = sum g1 x
    mix_tile<VEC, true>(g + o0, g + o1, out + r * n, nullptr, lg, rg, 0.f, 0.f, x + r * n, x + r * n, (int64_t)tile * kTile,
                        n, a0, a1);
    block_sum2(a0, a1);
    if (threadIdx.x == 0) part[blockIdx.x] = dlg * a0 + drg * a1;
  }
}

// ------------------------------------------------------------------ bus
// forward: out[b, c, i] = sum_t x[b, c, t, i] s[b, t]   grid = (bs*2) * tiles
template <bool VEC>
__global__ void __launch_bounds__(kThreads) bus_fwd_kernel(const float* __restrict__ x, const float* __restrict__ send_db,
                                                           float* __restrict__ y, int64_t n, int tracks, int tiles) {
  const int64_t bc = blockIdx.x / tiles;
  const int tile = blockIdx.x - (unsigned)(bc * tiles);
  const int64_t b = bc >> 1;
  const float* xr = x + bc * tracks * n;
  const int64_t t0 = (int64_t)tile * kTile;
  if (VEC) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int64_t i = t0 + (int64_t)(j * kThreads + threadIdx.x) * 4;
      if (i >= n) continue;
      float4 acc = make_float4(0, 0, 0, 0);
      for (int t = 0; t < tracks; ++t) {
        const float s = db_to_lin(send_db[b * tracks + t]);
        const float4 v = ld4(xr + (int64_t)t * n + i);
        acc.x = fmaf(s, v.x, acc.x); acc.y = fmaf(s, v.y, acc.y); acc.z = fmaf(s, v.z, acc.z); acc.w = fmaf(s, v.w, acc.w);
      }
      st4(y + bc * n + i, acc);
    }
  } else {
    for (int j = threadIdx.x; j < kTile; j += kThreads) {
      const int64_t i = t0 + j;
      if (i >= n) continue;
      float acc = 0.f;
      for (int t = 0; t < tracks; ++t) acc = fmaf(db_to_lin(send_db[b * tracks + t]), xr[(int64_t)t * n + i], acc);
      y[bc * n + i] = acc;
    }
  }
}
// backward: gx[b, c, t, i] = g[b, c, i] s[b, t];  part[(b*2+c)*tracks + t][tile] = sum_i g x    grid = (bs*2*tracks) * tiles
template <bool VEC>
__global__ void __launch_bounds__(kThreads) bus_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                           const float* __restrict__ send_db, float* __restrict__ gx,
                                                           float* __restrict__ part, int64_t n, int tracks, int tiles) {
  const int64_t row = blockIdx.x / tiles;            // (b, c, t)
  const int tile = blockIdx.x - (unsigned)(row * tiles);
  const int64_t bc = row / tracks, t = row - bc * tracks, b = bc >> 1;
  const float s = db_to_lin(send_db[b * tracks + t]);
  float a0 = 0.f, a1 = 0.f;
  mix_tile<VEC, true>(g + bc * n, nullptr, gx + row * n, nullptr, s, 0.f, 0.f, 0.f, x + row * n, x + row * n,
                      (int64_t)tile * kTile, n, a0, a1);
  block_sum2(a0, a1);
  if (threadIdx.x == 0) part[blockIdx.x] = a0;
}

// out[i] = scale(i) * sum_{j<cnt} part[i*cnt + j]   (one warp per output)
// mode 0: plain sum (widener, panner)    mode 1: bus: out[b*tracks+t] = ln10/20 * s * (sum over c in {0,1} and tiles)
__global__ void stereo_reduce_kernel(const float* __restrict__ part, const float* __restrict__ send_db,
                                     float* __restrict__ out, int64_t nout, int cnt, int mode, int tracks) {
  const int64_t o = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (o >= nout) return;
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  if (mode == 0) {
    for (int j = lane; j < cnt; j += 32) s += (double)part[o * cnt + j];
  } else {
    const int64_t b = o / tracks, t = o - b * tracks;
    for (int c = 0; c < 2; ++c)
      for (int j = lane; j < cnt; j += 32) s += (double)part[(((b * 2 + c) * tracks) + t) * cnt + j];
  }
  s = warp_sum(s);
  if (lane == 0) out[o] = (float)(mode == 0 ? s : s * (double)kLn10Over20 * (double)db_to_lin(send_db[o]));
}

inline int tiles_of(int64_t n) { return (int)((n + kTile - 1) / kTile); }
inline bool vec_ok(int64_t n, std::initializer_list<const void*> ptrs) {
  if (n % 4) return false;
  for (const void* p : ptrs) if (!aligned16(p)) return false;
  return true;
}
int reduce(const float* part, const float* send_db, float* out, int64_t nout, int cnt, int mode, int tracks, cudaStream_t st) {
  stereo_reduce_kernel<<<(unsigned)((nout + 7) / 8), 256, 0, st>>>(part, send_db, out, nout, cnt, mode, tracks);
  DASP_LAUNCH_OK("stereo_reduce_kernel");
  return DASP_OK;
}

}  // namespace
}  // namespace dasp

using namespace dasp;

extern "C" {

// floats of scratch for the three *_bwd calls: rows * ceil(n / tile) with rows = bs | bs*tracks | bs*2*tracks
int64_t dasp_stereo_bwd_workspace_floats(int64_t rows, int64_t n) { return rows * tiles_of(n > 0 ? n : 1); }

int dasp_widener_fwd(const float* x, const float* width, float* y, int64_t bs, int64_t n, void* stream) {
  DASP_REQUIRE(bs >= 0 && n >= 0, "widener fwd: negative size");
  if (bs == 0 || n == 0) return DASP_OK;
  DASP_REQUIRE(x && width && y, "widener fwd: null pointer");
  const int tiles = tiles_of(n);
  DASP_REQUIRE(bs * tiles < (1ll << 31), "widener fwd: grid too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (vec_ok(n, {x, y})) widener_kernel<true, false><<<(unsigned)(bs * tiles), kThreads, 0, st>>>(x, nullptr, width, y, nullptr, n, tiles);
  else                   widener_kernel<false, false><<<(unsigned)(bs * tiles), kThreads, 0, st>>>(x, nullptr, width, y, nullptr, n, tiles);
  DASP_LAUNCH_OK("widener_kernel");
  return DASP_OK;
}
int dasp_widener_bwd(const float* gy, const float* x, const float* width, float* gx, float* g_width, float* ws,
                     int64_t ws_floats, int64_t bs, int64_t n, void* stream) {
  DASP_REQUIRE(bs >= 0 && n >= 0, "widener bwd: negative size");
  if (bs == 0) return DASP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  DASP_REQUIRE(g_width != nullptr, "widener bwd: null g_width");
  if (n == 0) { DASP_CUDA_OK(cudaMemsetAsync(g_width, 0, sizeof(float) * bs, st)); return DASP_OK; }
  DASP_REQUIRE(gy && x && width && gx, "widener bwd: null pointer");
  const int tiles = tiles_of(n);
  if (!ws || ws_floats < bs * tiles) { set_error("widener bwd: workspace needs %lld floats", (long long)(bs * tiles)); return DASP_ERR_WORKSPACE; }
  if (vec_ok(n, {gy, x, gx})) widener_kernel<true, true><<<(unsigned)(bs * tiles), kThreads, 0, st>>>(gy, x, width, gx, ws, n, tiles);
  else                        widener_kernel<false, true><<<(unsigned)(bs * tiles), kThreads, 0, st>>>(gy, x, width, gx, ws, n, tiles);
  DASP_LAUNCH_OK("widener_kernel<bwd>");
  return reduce(ws, nullptr, g_width, bs, tiles, 0, 1, st);
}

int dasp_panner_fwd(const float* x, const float* pan, float* y, int64_t bs, int64_t tracks, int64_t n, void* stream) {
  DASP_REQUIRE(bs >= 0 && tracks >= 0 && n >= 0, "panner fwd: negative size");
  if (bs * tracks == 0 || n == 0) return DASP_OK;
  DASP_REQUIRE(x && pan && y, "panner fwd: null pointer");
  const int tiles = tiles_of(n);
  const int64_t grid = bs * tracks * tiles;
  DASP_REQUIRE(grid < (1ll << 31), "panner fwd: grid too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (vec_ok(n, {x, y})) panner_kernel<true, false><<<(unsigned)grid, kThreads, 0, st>>>(x, nullptr, pan, y, nullptr, n, (int)tracks, tiles);
  else                   panner_kernel<false, false><<<(unsigned)grid, kThreads, 0, st>>>(x, nullptr, pan, y, nullptr, n, (int)tracks, tiles);
  DASP_LAUNCH_OK("panner_kernel");
  return DASP_OK;
}
int dasp_panner_bwd(const float* gy, const float* x, const float* pan, float* gx, float* g_pan, float* ws,
                    int64_t ws_floats, int64_t bs, int64_t tracks, int64_t n, void* stream) {
  DASP_REQUIRE(bs >= 0 && tracks >= 0 && n >= 0, "panner bwd: negative size");
  const int64_t rows = bs * tracks;
  if (rows == 0) return DASP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  DASP_REQUIRE(g_pan != nullptr, "panner bwd: null g_pan");
  if (n == 0) { DASP_CUDA_OK(cudaMemsetAsync(g_pan, 0, sizeof(float) * rows, st)); return DASP_OK; }
  DASP_REQUIRE(gy && x && pan && gx, "panner bwd: null pointer");
  const int tiles = tiles_of(n);
  if (!ws || ws_floats < rows * tiles) { set_error("panner bwd: workspace needs %lld floats", (long long)(rows * tiles)); return DASP_ERR_WORKSPACE; }
  if (vec_ok(n, {gy, x, gx})) panner_kernel<true, true><<<(unsigned)(rows * tiles), kThreads, 0, st>>>(x, gy, pan, gx, ws, n, (int)tracks, tiles);
  else                        panner_kernel<false, true><<<(unsigned)(rows * tiles), kThreads, 0, st>>>(x, gy, pan, gx, ws, n, (int)tracks, tiles);
  DASP_LAUNCH_OK("panner_kernel<bwd>");
  return reduce(ws, nullptr, g_pan, rows, tiles, 0, 1, st);
}

int dasp_bus_fwd(const float* x, const float* send_db, float* y, int64_t bs, int64_t tracks, int64_t n, void* stream) {
  DASP_REQUIRE(bs >= 0 && tracks >= 0 && n >= 0, "bus fwd: negative size");
  if (bs == 0 || n == 0) return DASP_OK;
  DASP_REQUIRE(y != nullptr && (tracks == 0 || (x && send_db)), "bus fwd: null pointer");
  const int tiles = tiles_of(n);
  DASP_REQUIRE(bs * 2 * tiles < (1ll << 31), "bus fwd: grid too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (vec_ok(n, {x, y})) bus_fwd_kernel<true><<<(unsigned)(bs * 2 * tiles), kThreads, 0, st>>>(x, send_db, y, n, (int)tracks, tiles);
  else                   bus_fwd_kernel<false><<<(unsigned)(bs * 2 * tiles), kThreads, 0, st>>>(x, send_db, y, n, (int)tracks, tiles);
  DASP_LAUNCH_OK("bus_fwd_kernel");
  return DASP_OK;
}
int dasp_bus_bwd(const float* gy, const float* x, const float* send_db, float* gx, float* g_send_db, float* ws,
                 int64_t ws_floats, int64_t bs, int64_t tracks, int64_t n, void* stream) {
  DASP_REQUIRE(bs >= 0 && tracks >= 0 && n >= 0, "bus bwd: negative size");
  const int64_t rows = bs * 2 * tracks;
  if (rows == 0) return DASP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  DASP_REQUIRE(g_send_db != nullptr, "bus bwd: null g_send_db");
  if (n == 0) { DASP_CUDA_OK(cudaMemsetAsync(g_send_db, 0, sizeof(float) * bs * tracks, st)); return DASP_OK; }
  DASP_REQUIRE(gy && x && send_db && gx, "bus bwd: null pointer");
  const int tiles = tiles_of(n);
  if (!ws || ws_floats < rows * tiles) { set_error("bus bwd: workspace needs %lld floats", (long long)(rows * tiles)); return DASP_ERR_WORKSPACE; }
  DASP_REQUIRE(rows * tiles < (1ll << 31), "bus bwd: grid too large");
  if (vec_ok(n, {gy, x, gx})) bus_bwd_kernel<true><<<(unsigned)(rows * tiles), kThreads, 0, st>>>(gy, x, send_db, gx, ws, n, (int)tracks, tiles);
  else                        bus_bwd_kernel<false><<<(unsigned)(rows * tiles), kThreads, 0, st>>>(gy, x, send_db, gx, ws, n, (int)tracks, tiles);
  DASP_LAUNCH_OK("bus_bwd_kernel");
  return reduce(ws, send_db, g_send_db, bs * tracks, tiles, 1, (int)tracks, st);
}

}  // extern "C"
