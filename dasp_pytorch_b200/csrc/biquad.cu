// parametric_eq forward + backward: six cascaded biquads as time-parallel linear recurrences
// (reference: dasp_pytorch/functional.py:118-272, signal.py:242-306 design, :136-166 filtering).
//
// The reference never runs a recursion: it samples H = prod B_k/A_k on a 2^17-point FFT grid
// and multiplies spectra (signal.py:7-39).  That equals the zero-state IIR whenever the cascade's
// impulse response has died out inside n_fft - N samples (DESIGN.md "FSM vs recursion"), and it
// is what this file computes directly, in one pass over HBM.
//
// Realisation ("sigma form").  Direct forms are unusable in fp32 for the 20 Hz..2 kHz sections
// (a1 ~ -2, a2 ~ 1: the information sits in 1+a1+a2 ~ 1e-5).  Each section is instead run as
//       s1' = sg*s1 +    s2 + be1*u          sg  = -a1/2          be1 = b1 - a1*b0
//       s2' = q *s1 + sg*s2 + B2 *u          q   = sg^2 - a2      B2  = (b2 - a2*b0) + sg*be1
//       y   = s1 + b0*u
// whose state matrix [[sg,1],[q,sg]] has equal diagonal entries, i.e. is diagonally similar to a
// NORMAL matrix (rotation-scaling for complex poles, symmetric for real poles): round-off is
// amplified by 1/(1-r), not 1/((1-r) w0^2), and the small quantity q = -(r sin w0)^2 is stored
// directly instead of as a difference of O(1) numbers.  6 FMA per sample per section, all fp32;
// measured against the fp64 reference this is 100-1000x more accurate than the reference's own
// fp32 path (numerics table).  Coefficients and all matrix powers are designed in fp64.
//
// Parallelisation.  One CTA per (item, channel) row walks its N samples through the TMA tile
// pipeline (tile_pipe.cuh).  In a tile each thread owns E consecutive samples and, per section:
//   local zero-state pass -> Kogge-Stone shuffle scan of the 2-vector end states with the
//   precomputed powers A^(E*2^k) -> cross-warp carry in shared memory -> fix-up
//   y[j] += (A^j c_in)_1 from a per-item table.  (A is constant in time, so the scan operator is
//   a matrix power, not a generic 2x2 pair product.)
// The tile-to-tile carries stay in registers; the forward stores them per tile as checkpoints.
//
// Packed fp32x2 math.  Measured on B200 (tools/probe/ffma2_probe.cu): scalar FFMA peaks at 21 TFMA/s, the
// Blackwell-only packed FFMA2 (fma.rn.f32x2, __ffma2_rn) at 33 TFMA/s, and the scalar version of this kernel
// already sat at 74 % of the scalar peak.  Every thread therefore processes a PAIR of rows -- the left and
// right channel of a stereo item, or two neighbouring mono items -- with each float2 lane holding one row:
// states, coefficients, matrix powers and scan operands are all (rowA, rowB) pairs, so the whole recurrence,
// the shuffle scan and the fix-up run on FFMA2/FADD2/FMUL2.  An odd last row is paired with itself.
//
// Backward.  State-space adjoint (SURVEY.md A.3 restated for the sigma form): with lam = adjoint
// state,  lam[n] = A^T lam[n+1] + (g[n],0);  gu[n] = be1*lam1[n+1] + B2*lam2[n+1] + b0*g[n];
//   d sg = sum lam[n+1].s[n],  d q = sum lam2[n+1] s1[n],  d be1 = sum lam1[n+1] u[n],
//   d B2 = sum lam2[n+1] u[n], d b0 = sum g[n] u[n].
// Tiles are swept in reverse time order; per tile the six section inputs are recomputed from the
// checkpoint into thread-private shared memory, then sections are unwound 6 -> 1.  The 30 sums per
// row are reduced deterministically; a second tiny kernel adds the channels of an item and applies
// the fp64 Jacobian d(sg,q,be1,B2,b0)/d(gain_dB, fc, Q) (forward-mode dual numbers).
#include <math.h>

#include "common.cuh"
#include "tile_pipe.cuh"

namespace dasp {
namespace {

#ifndef DASP_EQ_E
#define DASP_EQ_E 11
#endif
#ifndef DASP_EQ_WARPS_PER_SM
#define DASP_EQ_WARPS_PER_SM 24
#endif
constexpr int kE = DASP_EQ_E;    // samples per thread per tile (odd: conflict-free stride-E smem access)
constexpr int kStages = 3;
constexpr int kSections = 6;
constexpr int kNumPowTables = kE + 5 + 32 + 1;   // A^j (j<E) | A^(E 2^k) (k<5) | A^(E lane) | A^(32E)

// ------------------------------------------------------------------ coefficient design (fp64)
// forward-mode dual number with 3 directional derivatives (gain_dB, fc, Q)
struct Dual3 {
  double v, d[3];
};
__host__ __device__ inline Dual3 mk(double v) { return {v, {0, 0, 0}}; }
__host__ __device__ inline Dual3 operator+(Dual3 a, Dual3 b) { return {a.v + b.v, {a.d[0] + b.d[0], a.d[1] + b.d[1], a.d[2] + b.d[2]}}; }
__host__ __device__ inline Dual3 operator-(Dual3 a, Dual3 b) { return {a.v - b.v, {a.d[0] - b.d[0], a.d[1] - b.d[1], a.d[2] - b.d[2]}}; }
__host__ __device__ inline Dual3 operator-(Dual3 a) { return {-a.v, {-a.d[0], -a.d[1], -a.d[2]}}; }
__host__ __device__ inline Dual3 operator*(Dual3 a, Dual3 b) {
  return {a.v * b.v, {a.d[0] * b.v + a.v * b.d[0], a.d[1] * b.v + a.v * b.d[1], a.d[2] * b.v + a.v * b.d[2]}};
}
__host__ __device__ inline Dual3 operator/(Dual3 a, Dual3 b) {
  const double inv = 1.0 / b.v, r = a.v * inv;
  return {r, {(a.d[0] - r * b.d[0]) * inv, (a.d[1] - r * b.d[1]) * inv, (a.d[2] - r * b.d[2]) * inv}};
}
__host__ __device__ inline Dual3 operator*(double s, Dual3 a) { return {s * a.v, {s * a.d[0], s * a.d[1], s * a.d[2]}}; }
__host__ __device__ inline Dual3 operator+(double s, Dual3 a) { return {s + a.v, {a.d[0], a.d[1], a.d[2]}}; }
__host__ __device__ inline Dual3 operator-(double s, Dual3 a) { return {s - a.v, {-a.d[0], -a.d[1], -a.d[2]}}; }
__host__ __device__ inline Dual3 chain(Dual3 a, double f, double df) { return {f, {df * a.d[0], df * a.d[1], df * a.d[2]}}; }
__host__ __device__ inline Dual3 dsin(Dual3 a) { return chain(a, sin(a.v), cos(a.v)); }
__host__ __device__ inline Dual3 dcos(Dual3 a) { return chain(a, cos(a.v), -sin(a.v)); }
__host__ __device__ inline Dual3 dexp(Dual3 a) { double e = exp(a.v); return chain(a, e, e); }
__host__ __device__ inline Dual3 dsqrt(Dual3 a) { double s = sqrt(a.v); return chain(a, s, 0.5 / s); }

struct SigmaCoef {   // [sg, q, be1, B2, b0]
  Dual3 c[5];
};

// RBJ cookbook biquad (signal.py:242-306) -> sigma-form coefficients, with derivatives.
// kind: 0 = low shelf (:268-274), 1 = peaking (:275-281), 2 = high shelf (:261-267).
__host__ __device__ inline SigmaCoef design_section(double gain_db, double fc, double qf, double sr, int kind) {
  Dual3 g = {gain_db, {1, 0, 0}}, f = {fc, {0, 1, 0}}, Q = {qf, {0, 0, 1}};
  Dual3 A = dexp((0.05756462732485114 /* ln10/40 */) * g);
  Dual3 w0 = (6.283185307179586 / sr) * f;
  Dual3 alpha = dsin(w0) / (2.0 * Q);
  Dual3 cw = dcos(w0);
  Dual3 b0, b1, b2, a0, a1, a2;
  if (kind == 1) {
    b0 = 1.0 + alpha * A; b1 = -2.0 * cw; b2 = 1.0 - alpha * A;
    a0 = 1.0 + alpha / A; a1 = -2.0 * cw; a2 = 1.0 - alpha / A;
  } else {
    const double sgn = (kind == 0) ? 1.0 : -1.0;     // the shelves differ in the sign of the cos terms
    Dual3 s = 2.0 * dsqrt(A) * alpha;
    Dual3 ap1 = 1.0 + A, am1 = A - mk(1.0);
    Dual3 t = sgn * (am1 * cw);
    b0 = A * (ap1 - t + s);
    b1 = (sgn * 2.0) * (A * (am1 - sgn * (ap1 * cw)));
    b2 = A * (ap1 - t - s);
    a0 = ap1 + t + s;
    a1 = (-sgn * 2.0) * (am1 + sgn * (ap1 * cw));
    a2 = ap1 + t - s;
  }
  b0 = b0 / a0; b1 = b1 / a0; b2 = b2 / a0; a1 = a1 / a0; a2 = a2 / a0;
  SigmaCoef o;
  Dual3 sg = -0.5 * a1;
  Dual3 be1 = b1 - a1 * b0;
  o.c[0] = sg;
  o.c[1] = sg * sg - a2;
  o.c[2] = be1;
  o.c[3] = (b2 - a2 * b0) + sg * be1;
  o.c[4] = b0;
  return o;
}
__host__ __device__ inline int section_kind(int k) { return k == 0 ? 0 : (k == 5 ? 2 : 1); }

// 2x2 fp64 matrix power of [[sg,1],[q,sg]] by binary exponentiation
struct M2d { double a, b, c, d; };
__device__ inline M2d mm(const M2d& x, const M2d& y) {
  return {x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d};
}
__device__ inline M2d mpow(double sg, double q, unsigned n) {
  M2d r = {1, 0, 0, 1}, b = {sg, 1.0, q, sg};
  while (n) {
    if (n & 1u) r = mm(r, b);
    b = mm(b, b);
    n >>= 1;
  }
  return r;
}

// ------------------------------------------------------------------ packed helpers
typedef float2 f2;    // .x = row A of the pair, .y = row B
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 addp(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 zero2() { return make_float2(0.f, 0.f); }
struct St { f2 s1, s2; };                      // 2-vector state of one section, both rows
struct Mp { f2 a, b, c, d; };                  // 2x2 matrix [[a,b],[c,d]] per row

__device__ __forceinline__ St mv(const Mp& m, const St& v) {           // M v
  return {fma2(m.a, v.s1, mul2(m.b, v.s2)), fma2(m.c, v.s1, mul2(m.d, v.s2))};
}
__device__ __forceinline__ St mtv(const Mp& m, const St& v) {          // M^T v
  return {fma2(m.a, v.s1, mul2(m.c, v.s2)), fma2(m.b, v.s1, mul2(m.d, v.s2))};
}
__device__ __forceinline__ St adds(const St& x, const St& y) { return {addp(x.s1, y.s1), addp(x.s2, y.s2)}; }
__device__ __forceinline__ f2 shfl_up2(f2 v, int d) {
  return make_float2(__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d));
}
__device__ __forceinline__ f2 shfl_dn2(f2 v, int d) {
  return make_float2(__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d));
}
__device__ __forceinline__ f2 shfl_at2(f2 v, int l) {
  return make_float2(__shfl_sync(0xffffffffu, v.x, l), __shfl_sync(0xffffffffu, v.y, l));
}

// ------------------------------------------------------------------ shared-memory layout
// per-pair tables (fp32, both rows interleaved), built once per CTA.  Power-table entries per section:
//   [0, E)      A^j            (fix-up / zero-input response)
//   [E, E+5)    A^(E 2^s)      (Kogge-Stone steps)
//   E+5         A^(32 E)       (warp-to-warp carry)
//   E+6 ..      A^(E lane), lane = 0..31   (only when a CTA has more than one warp)
template <int W>
struct __align__(16) PairTables {
  static constexpr int kNT = kE + 5 + 1 + (W > 1 ? 32 : 0);
  f2 cf[kSections][6];                 // sg, q, be1, B2, b0, pad
  Mp pw[kSections][kNT];
};
constexpr size_t kHdrBars = 64;                                   // mbarriers
constexpr size_t kHdrAgg = 16 * 8 * sizeof(St);                   // up to 16 scan slots x 8 warps
template <int W>
__host__ __device__ constexpr size_t hdr_bytes() { return ((kHdrBars + kHdrAgg + sizeof(PairTables<W>) + 127) / 128) * 128; }

template <int W>
struct Smem {
  uint64_t* bars; St* agg; PairTables<W>* tb; float* stages;
  __device__ __forceinline__ Smem(unsigned char* base) {
    bars = reinterpret_cast<uint64_t*>(base);
    agg = reinterpret_cast<St*>(base + kHdrBars);
    tb = reinterpret_cast<PairTables<W>*>(base + kHdrBars + kHdrAgg);
    stages = reinterpret_cast<float*>(base + hdr_bytes<W>());
  }
};

struct EqParams {
  const float* x;        // (bs, C, N)
  const float* gy;       // backward
  float* y;              // forward out / backward gx
  const float* params;   // (bs, 18): gain_dB, fc, Q per section, signature order
  float* ckpt;           // (pairs, ntiles, 6, 4): section states (s1A, s1B, s2A, s2B) entering each tile
  float* partial;        // (rows, 30) backward: per-row coefficient-gradient sums
  int64_t n;
  int64_t rows;
  int chs;
  int ntiles;
  float sample_rate;
  int bulk;
};

struct PairIO {     // forward: buffer 0 = row A, 1 = row B (in place)
  const float* a; const float* b; float* ya; float* yb;
  __device__ __forceinline__ const float* src(int i) const { return i == 0 ? a : b; }
  __device__ __forceinline__ float* dst(int i) const { return i == 0 ? ya : yb; }
};
struct PairIOBwd {  // backward: buffers 0,1 = x rows (read only), 2,3 = gy rows -> gx rows
  const float* xa; const float* xb; const float* ga; const float* gb; float* oa; float* ob;
  __device__ __forceinline__ const float* src(int i) const { return i == 0 ? xa : (i == 1 ? xb : (i == 2 ? ga : gb)); }
  __device__ __forceinline__ float* dst(int i) const { return i == 2 ? oa : (i == 3 ? ob : nullptr); }
};

// build the per-pair tables: every thread of the CTA participates; ends with __syncthreads()
template <int W>
__device__ void build_tables(PairTables<W>& tb, const float* params, int64_t item_a, int64_t item_b, float sample_rate) {
  __shared__ double cfd[2][kSections][2];   // sg, q in fp64 for the matrix powers
  const int tid = threadIdx.x;
  for (int t = tid; t < 2 * kSections; t += blockDim.x) {
    const int r = t / kSections, k = t - r * kSections;
    const float* pp = params + (r == 0 ? item_a : item_b) * 18 + 3 * k;
    const SigmaCoef sc = design_section((double)pp[0], (double)pp[1], (double)pp[2], (double)sample_rate, section_kind(k));
    float* dst = reinterpret_cast<float*>(&tb.cf[k][0]) + r;
#pragma unroll
    for (int j = 0; j < 5; ++j) dst[2 * j] = (float)sc.c[j].v;
    dst[10] = 0.f;
    cfd[r][k][0] = sc.c[0].v;
    cfd[r][k][1] = sc.c[1].v;
  }
  __syncthreads();
  constexpr int NT = PairTables<W>::kNT;
  for (int idx = tid; idx < 2 * kSections * NT; idx += blockDim.x) {
    const int r = idx / (kSections * NT), rem = idx - r * (kSections * NT);
    const int k = rem / NT, e = rem - k * NT;
    unsigned n;
    if (e < kE) n = (unsigned)e;
    else if (e < kE + 5) n = (unsigned)kE << (e - kE);
    else if (e == kE + 5) n = (unsigned)(kE * 32);
    else n = (unsigned)(kE * (e - kE - 6));
    const M2d m = mpow(cfd[r][k][0], cfd[r][k][1], n);
    float* dst = reinterpret_cast<float*>(&tb.pw[k][e]) + r;
    dst[0] = (float)m.a; dst[2] = (float)m.b; dst[4] = (float)m.c; dst[6] = (float)m.d;
  }
  __syncthreads();
}

// ------------------------------------------------------------------ 2-state block scans (both rows packed)
// forward in time.  v = this thread's end state after a zero-state local pass; c_tile = state entering the tile
// (updated to the state leaving it).  Returns the state entering this thread's chunk.
template <int W>
__device__ __forceinline__ St scan_fwd2(St v, St& c_tile, const PairTables<W>& tb, int k, St* agg, int lane, int warp) {
  if (W == 1) {
    // single warp: fold the tile carry into lane 0 (v0 += A^E c), then the inclusive scan yields every
    // thread's TRUE end state and the incoming state is simply the left neighbour's
    if (lane == 0) v = adds(v, mv(tb.pw[k][kE + 0], c_tile));
#pragma unroll
    for (int s = 0; s < 5; ++s) {
      St u = {shfl_up2(v.s1, 1 << s), shfl_up2(v.s2, 1 << s)};
      if (lane >= (1 << s)) v = adds(v, mv(tb.pw[k][kE + s], u));
    }
    St in = {shfl_up2(v.s1, 1), shfl_up2(v.s2, 1)};
    if (lane == 0) in = c_tile;
    c_tile = {shfl_at2(v.s1, 31), shfl_at2(v.s2, 31)};
    return in;
  }
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    St u = {shfl_up2(v.s1, 1 << s), shfl_up2(v.s2, 1 << s)};
    if (lane >= (1 << s)) v = adds(v, mv(tb.pw[k][kE + s], u));
  }
  St excl = {shfl_up2(v.s1, 1), shfl_up2(v.s2, 1)};
  if (lane == 0) excl = {zero2(), zero2()};
  if (lane == 31) agg[warp] = v;
  __syncthreads();
  St c = c_tile, c_warp = c_tile;
  const Mp wm = tb.pw[k][kE + 5];
#pragma unroll
  for (int w = 0; w < W; ++w) {
    if (w == warp) c_warp = c;
    c = adds(mv(wm, c), agg[w]);
  }
  c_tile = c;
  return adds(excl, mv(tb.pw[k][kE + 6 + lane], c_warp));
}

// reverse in time (adjoint): v = adjoint state at this thread's first sample after a zero-terminal local
// reverse pass; c_tile = adjoint state at the first sample of the NEXT tile.  Uses transposed powers.
template <int W>
__device__ __forceinline__ St scan_rev2(St v, St& c_tile, const PairTables<W>& tb, int k, St* agg, int lane, int warp) {
  if (W == 1) {
    if (lane == 31) v = adds(v, mtv(tb.pw[k][kE + 0], c_tile));
#pragma unroll
    for (int s = 0; s < 5; ++s) {
      St u = {shfl_dn2(v.s1, 1 << s), shfl_dn2(v.s2, 1 << s)};
      if (lane + (1 << s) < 32) v = adds(v, mtv(tb.pw[k][kE + s], u));
    }
    St in = {shfl_dn2(v.s1, 1), shfl_dn2(v.s2, 1)};
    if (lane == 31) in = c_tile;
    c_tile = {shfl_at2(v.s1, 0), shfl_at2(v.s2, 0)};
    return in;
  }
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    St u = {shfl_dn2(v.s1, 1 << s), shfl_dn2(v.s2, 1 << s)};
    if (lane + (1 << s) < 32) v = adds(v, mtv(tb.pw[k][kE + s], u));
  }
  St excl = {shfl_dn2(v.s1, 1), shfl_dn2(v.s2, 1)};
  if (lane == 31) excl = {zero2(), zero2()};
  if (lane == 0) agg[warp] = v;
  __syncthreads();
  St c = c_tile, c_warp = c_tile;
  const Mp wm = tb.pw[k][kE + 5];
#pragma unroll
  for (int w = W - 1; w >= 0; --w) {
    if (w == warp) c_warp = c;
    c = adds(mtv(wm, c), agg[w]);
  }
  c_tile = c;
  // distance from the first sample of thread lane+1 to the first sample of the next warp: (31-lane) chunks
  return adds(excl, mtv(tb.pw[k][kE + 6 + (31 - lane)], c_warp));
}

struct Cf { f2 sg, q, be1, B2, b0; };
template <int W>
__device__ __forceinline__ Cf load_cf(const PairTables<W>& tb, int k) {
  return {tb.cf[k][0], tb.cf[k][1], tb.cf[k][2], tb.cf[k][3], tb.cf[k][4]};
}

// zero-state local pass of one section over the thread's E samples (in place); returns the end state
__device__ __forceinline__ St local_pass(f2 (&v)[kE], const Cf& c) {
  f2 s1 = zero2(), s2 = zero2();
#pragma unroll
  for (int j = 0; j < kE; ++j) {
    const f2 u = v[j];
    v[j] = fma2(c.b0, u, s1);
    const f2 t1 = fma2(c.be1, u, fma2(c.sg, s1, s2));
    s2 = fma2(c.B2, u, fma2(c.q, s1, mul2(c.sg, s2)));
    s1 = t1;
  }
  return {s1, s2};
}

__device__ __forceinline__ void load_pair(f2 (&v)[kE], const float* a, const float* b, int64_t n0, int64_t n) {
#pragma unroll
  for (int j = 0; j < kE; ++j) v[j] = (n0 + j < n) ? make_float2(a[j], b[j]) : zero2();
}

// =============================================================================== forward
template <int W>
__global__ void __launch_bounds__(W * 32) eq_fwd_kernel(EqParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem<W> sm(smem_raw);
  PairTables<W>& tb = *sm.tb;
  const int64_t row_a = 2 * (int64_t)blockIdx.x;
  const bool has_b = row_a + 1 < p.rows;
  const int64_t row_b = has_b ? row_a + 1 : row_a;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tile_len = W * 32 * kE;

  build_tables<W>(tb, p.params, row_a / p.chs, row_b / p.chs, p.sample_rate);

  TileGeom g{p.n, tile_len, p.ntiles, false};
  PairIO rows{p.x + row_a * p.n, p.x + row_b * p.n, p.y + row_a * p.n, has_b ? p.y + row_b * p.n : nullptr};
  TilePipe<kStages> pipe;
  pipe.init(sm.bars, sm.stages, 2, tile_len, p.bulk != 0);
  pipe.prologue(g, rows);

  St carry[kSections];
#pragma unroll
  for (int k = 0; k < kSections; ++k) carry[k] = {zero2(), zero2()};
  const int off = threadIdx.x * kE;

  for (int i = 0; i < p.ntiles; ++i) {
    pipe.acquire(i, g, rows);
    float* ba = pipe.buf(i % kStages, 0) + off;
    float* bb = pipe.buf(i % kStages, 1) + off;
    const int64_t n0 = (int64_t)i * tile_len + off;
    if (p.ckpt && threadIdx.x == 0) {
      St* ck = reinterpret_cast<St*>(p.ckpt + ((int64_t)blockIdx.x * p.ntiles + i) * 24);
#pragma unroll
      for (int k = 0; k < kSections; ++k) ck[k] = carry[k];
    }
    f2 v[kE];
    load_pair(v, ba, bb, n0, p.n);

#pragma unroll
    for (int k = 0; k < kSections; ++k) {
      const Cf c = load_cf<W>(tb, k);
      const St end = local_pass(v, c);
      const St cin = scan_fwd2<W>(end, carry[k], tb, k, sm.agg + ((i * kSections + k) & 1) * 8, lane, warp);
#pragma unroll
      for (int j = 0; j < kE; ++j) {
        const Mp& m = tb.pw[k][j];                       // y[j] += (A^j c_in)_1
        v[j] = fma2(m.a, cin.s1, fma2(m.b, cin.s2, v[j]));
      }
    }
#pragma unroll
    for (int j = 0; j < kE; ++j) { ba[j] = v[j].x; bb[j] = v[j].y; }
    pipe.release(i, g, rows);
  }
  pipe.drain();
}

// =============================================================================== backward
constexpr int kBwdStages = 2;

template <int W>
__global__ void __launch_bounds__(W * 32) eq_bwd_kernel(EqParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem<W> sm(smem_raw);
  PairTables<W>& tb = *sm.tb;
  __shared__ double red[W][kSections * 5 * 2];
  const int64_t row_a = 2 * (int64_t)blockIdx.x;
  const bool has_b = row_a + 1 < p.rows;
  const int64_t row_b = has_b ? row_a + 1 : row_a;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tile_len = W * 32 * kE;
  const int nthr = W * 32;

  build_tables<W>(tb, p.params, row_a / p.chs, row_b / p.chs, p.sample_rate);

  TileGeom g{p.n, tile_len, p.ntiles, true};
  PairIOBwd rows{p.x + row_a * p.n, p.x + row_b * p.n, p.gy + row_a * p.n, p.gy + row_b * p.n,
                 p.y + row_a * p.n, has_b ? p.y + row_b * p.n : nullptr};
  TilePipe<kBwdStages> pipe;
  pipe.init(sm.bars, sm.stages, 4, tile_len, p.bulk != 0);
  pipe.prologue(g, rows);
  // behind the pipeline stages: thread-private scratch for the inputs of sections 1..5 (u_1..u_5, packed pairs)
  // and the 30 running sums per thread, laid out [sum][thread] so that warps touch consecutive words
  f2* scratch = reinterpret_cast<f2*>(sm.stages + (size_t)kBwdStages * 4 * tile_len);
  f2* accs = scratch + (size_t)(kSections - 1) * tile_len;
  for (int q = 0; q < kSections * 5; ++q) accs[q * nthr + threadIdx.x] = zero2();

  St adj[kSections];          // adjoint state at the first sample of the next tile, per section
#pragma unroll
  for (int k = 0; k < kSections; ++k) adj[k] = {zero2(), zero2()};
  const int off = threadIdx.x * kE;

  for (int i = 0; i < p.ntiles; ++i) {
    pipe.acquire(i, g, rows);
    const int st = i % kBwdStages;
    const int tile = g.tile_of(i);
    const int64_t n0 = (int64_t)tile * tile_len + off;
    const float* xa = pipe.buf(st, 0) + off;
    const float* xb = pipe.buf(st, 1) + off;
    float* ga = pipe.buf(st, 2) + off;
    float* gb = pipe.buf(st, 3) + off;
    const St* ck = reinterpret_cast<const St*>(p.ckpt + ((int64_t)blockIdx.x * p.ntiles + tile) * 24);

    // ---- phase F: recompute the section inputs u_1..u_5 and every section's incoming state ----
    St cin[kSections];
    {
      f2 v[kE];
      load_pair(v, xa, xb, n0, p.n);
#pragma unroll
      for (int k = 0; k < kSections; ++k) {
        const Cf c = load_cf<W>(tb, k);
        const St end = local_pass(v, c);
        St ct = ck[k];
        cin[k] = scan_fwd2<W>(end, ct, tb, k, sm.agg + (2 * k + 0) * 8, lane, warp);
        if (k < kSections - 1) {
          f2* uk = scratch + (size_t)k * tile_len + off;
#pragma unroll
          for (int j = 0; j < kE; ++j) {
            const Mp& m = tb.pw[k][j];
            v[j] = fma2(m.a, cin[k].s1, fma2(m.b, cin[k].s2, v[j]));
            uk[j] = v[j];
          }
        }
      }
    }

    // ---- phase B: unwind the sections 6 -> 1 ----
    f2 gq[kE];
    load_pair(gq, ga, gb, n0, p.n);
#pragma unroll
    for (int k = kSections - 1; k >= 0; --k) {
      const Cf c = load_cf<W>(tb, k);
      f2 u[kE], s1[kE], s2[kE];
      if (k == 0) {
        load_pair(u, xa, xb, n0, p.n);
      } else {
        const f2* uk = scratch + (size_t)(k - 1) * tile_len + off;
#pragma unroll
        for (int j = 0; j < kE; ++j) u[j] = (n0 + j < p.n) ? uk[j] : zero2();
      }
      {  // true-state forward pass: s[j] = state BEFORE sample j
        f2 a1 = cin[k].s1, a2 = cin[k].s2;
#pragma unroll
        for (int j = 0; j < kE; ++j) {
          s1[j] = a1; s2[j] = a2;
          const f2 t1 = fma2(c.be1, u[j], fma2(c.sg, a1, a2));
          a2 = fma2(c.B2, u[j], fma2(c.q, a1, mul2(c.sg, a2)));
          a1 = t1;
        }
      }
      St agg_v;
      {  // zero-terminal reverse pass: only the value reaching the chunk's first sample is needed
        f2 l1 = zero2(), l2 = zero2();
#pragma unroll
        for (int j = kE - 1; j >= 0; --j) {
          const f2 t1 = fma2(c.sg, l1, fma2(c.q, l2, gq[j]));
          l2 = fma2(c.sg, l2, l1);
          l1 = t1;
        }
        agg_v = {l1, l2};
      }
      const St din = scan_rev2<W>(agg_v, adj[k], tb, k, sm.agg + (2 * k + 1) * 8, lane, warp);
      f2 a0 = zero2(), a1s = zero2(), a2s = zero2(), a3 = zero2(), a4 = zero2();
      {  // final reverse pass with the true terminal adjoint state
        f2 l1 = din.s1, l2 = din.s2;      // lambda[n+1] while processing sample n
#pragma unroll
        for (int j = kE - 1; j >= 0; --j) {
          const f2 gj = gq[j];
          a0 = fma2(l1, s1[j], fma2(l2, s2[j], a0));
          a1s = fma2(l2, s1[j], a1s);
          a2s = fma2(l1, u[j], a2s);
          a3 = fma2(l2, u[j], a3);
          a4 = fma2(gj, u[j], a4);
          gq[j] = fma2(c.be1, l1, fma2(c.B2, l2, mul2(c.b0, gj)));
          const f2 t1 = fma2(c.sg, l1, fma2(c.q, l2, gj));
          l2 = fma2(c.sg, l2, l1);
          l1 = t1;
        }
      }
      f2* ac = accs + (size_t)(k * 5) * nthr + threadIdx.x;
      ac[0 * nthr] = addp(ac[0 * nthr], a0);
      ac[1 * nthr] = addp(ac[1 * nthr], a1s);
      ac[2 * nthr] = addp(ac[2 * nthr], a2s);
      ac[3 * nthr] = addp(ac[3 * nthr], a3);
      ac[4 * nthr] = addp(ac[4 * nthr], a4);
    }
#pragma unroll
    for (int j = 0; j < kE; ++j) { ga[j] = gq[j].x; gb[j] = gq[j].y; }
    pipe.release(i, g, rows);
  }
  pipe.drain();

  // ---- deterministic block reduction of the 30 sums per row (fp64) ----
  for (int q = 0; q < kSections * 5; ++q) {
    const f2 v = accs[q * nthr + threadIdx.x];
    const double sa = warp_sum((double)v.x), sb = warp_sum((double)v.y);
    if (lane == 0) { red[warp][2 * q] = sa; red[warp][2 * q + 1] = sb; }
  }
  __syncthreads();
  if (threadIdx.x < kSections * 5 * 2) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < W; ++w) s += red[w][threadIdx.x];
    const int q = threadIdx.x >> 1, r = threadIdx.x & 1;
    if (r == 0) p.partial[row_a * 30 + q] = (float)s;
    else if (has_b) p.partial[row_b * 30 + q] = (float)s;
  }
}

// one thread per (item, section): add the channel partials, apply the fp64 Jacobian
__global__ void eq_param_grad_kernel(const float* __restrict__ partial, const float* __restrict__ params,
                                     float* __restrict__ gparams, int64_t bs, int chs, float sample_rate) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= bs * kSections) return;
  const int64_t item = idx / kSections;
  const int k = (int)(idx - item * kSections);
  double gc[5] = {0, 0, 0, 0, 0};
  for (int c = 0; c < chs; ++c) {
    const float* pr = partial + ((int64_t)item * chs + c) * 30 + k * 5;
#pragma unroll
    for (int q = 0; q < 5; ++q) gc[q] += (double)pr[q];
  }
  const float* pp = params + item * 18 + 3 * k;
  const SigmaCoef sc = design_section((double)pp[0], (double)pp[1], (double)pp[2], (double)sample_rate, section_kind(k));
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 5; ++q) s += gc[q] * sc.c[q].d[d];
    gparams[item * 18 + 3 * k + d] = (float)s;
  }
}

// ---- host side -----------------------------------------------------------------------------
int pick_warps(int64_t pairs) {
  const int64_t want = (int64_t)DASP_EQ_WARPS_PER_SM * sm_count();
  int w = 1;
  while (w < 4 && pairs * w < want) w *= 2;
  return w;
}
template <int W> size_t smem_fwd() { return hdr_bytes<W>() + (size_t)kStages * 2 * (W * 32 * kE) * 4; }
template <int W> size_t smem_bwd() {
  const size_t tile = (size_t)W * 32 * kE;
  return hdr_bytes<W>() + (size_t)kBwdStages * 4 * tile * 4 + (size_t)(kSections - 1) * tile * 8 +
         (size_t)kSections * 5 * (W * 32) * 8;
}

template <int W>
int launch_fwd_w(const EqParams& p, int64_t pairs, cudaStream_t st) {
  const size_t smem = smem_fwd<W>();
  DASP_CUDA_OK(cudaFuncSetAttribute(eq_fwd_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  eq_fwd_kernel<W><<<(unsigned)pairs, W * 32, smem, st>>>(p);
  DASP_LAUNCH_OK("eq_fwd_kernel");
  return DASP_OK;
}
template <int W>
int launch_bwd_w(const EqParams& p, int64_t pairs, cudaStream_t st) {
  const size_t smem = smem_bwd<W>();
  DASP_CUDA_OK(cudaFuncSetAttribute(eq_bwd_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  eq_bwd_kernel<W><<<(unsigned)pairs, W * 32, smem, st>>>(p);
  DASP_LAUNCH_OK("eq_bwd_kernel");
  return DASP_OK;
}

}  // namespace
}  // namespace dasp

using namespace dasp;

extern "C" {

int64_t dasp_eq_tile_len(int64_t rows) { return (int64_t)pick_warps((rows + 1) / 2) * 32 * kE; }
int64_t dasp_eq_ckpt_floats(int64_t rows, int64_t n) {
  const int64_t tile = dasp_eq_tile_len(rows);
  const int64_t ntiles = n > 0 ? (n + tile - 1) / tile : 1;
  return ((rows + 1) / 2) * ntiles * 24;
}
int64_t dasp_eq_bwd_workspace_floats(int64_t bs, int64_t chs) { return bs * chs * 30; }

int dasp_eq_fwd(const float* x, const float* params, float* y, float* ckpt, int64_t bs, int64_t chs, int64_t n,
                float sample_rate, void* stream) {
  DASP_REQUIRE(bs >= 0 && chs >= 1 && n >= 0, "eq fwd: bad shape bs=%lld chs=%lld n=%lld", (long long)bs,
               (long long)chs, (long long)n);
  if (bs == 0 || n == 0) return DASP_OK;
  DASP_REQUIRE(x && params && y, "eq fwd: null pointer");
  DASP_REQUIRE(sample_rate > 0.f, "eq fwd: sample_rate must be positive");
  const int64_t rows = bs * chs, pairs = (rows + 1) / 2;
  DASP_REQUIRE(pairs < (1ll << 31), "eq fwd: too many rows");
  const int w = pick_warps(pairs);
  const int tile_len = w * 32 * kE;
  EqParams p{};
  p.x = x; p.y = y; p.params = params; p.ckpt = ckpt; p.n = n; p.rows = rows; p.chs = (int)chs;
  p.ntiles = (int)((n + tile_len - 1) / tile_len); p.sample_rate = sample_rate;
  p.bulk = (n % 4 == 0) && aligned16(x) && aligned16(y);
  cudaStream_t st = (cudaStream_t)stream;
  switch (w) {
    case 1: return launch_fwd_w<1>(p, pairs, st);
    case 2: return launch_fwd_w<2>(p, pairs, st);
    default: return launch_fwd_w<4>(p, pairs, st);
  }
}

int dasp_eq_bwd(const float* gy, const float* x, const float* params, const float* ckpt, float* gx,
                float* gparams, float* ws, int64_t ws_floats, int64_t bs, int64_t chs, int64_t n,
                float sample_rate, void* stream) {
  DASP_REQUIRE(bs >= 0 && chs >= 1 && n >= 0, "eq bwd: bad shape bs=%lld chs=%lld n=%lld", (long long)bs,
               (long long)chs, (long long)n);
  if (bs == 0) return DASP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  DASP_REQUIRE(gparams != nullptr, "eq bwd: null gparams");
  if (n == 0) { DASP_CUDA_OK(cudaMemsetAsync(gparams, 0, sizeof(float) * 18 * bs, st)); return DASP_OK; }
  DASP_REQUIRE(gy && x && params && ckpt && gx, "eq bwd: null pointer");
  const int64_t rows = bs * chs, pairs = (rows + 1) / 2;
  DASP_REQUIRE(pairs < (1ll << 31), "eq bwd: too many rows");
  if (ws == nullptr || ws_floats < rows * 30) {
    set_error("eq bwd: workspace needs %lld floats, got %lld", (long long)(rows * 30), (long long)ws_floats);
    return DASP_ERR_WORKSPACE;
  }
  const int w = pick_warps(pairs);
  const int tile_len = w * 32 * kE;
  EqParams p{};
  p.x = x; p.gy = gy; p.y = gx; p.params = params; p.ckpt = const_cast<float*>(ckpt); p.partial = ws; p.n = n;
  p.rows = rows; p.chs = (int)chs; p.ntiles = (int)((n + tile_len - 1) / tile_len); p.sample_rate = sample_rate;
  p.bulk = (n % 4 == 0) && aligned16(x) && aligned16(gy) && aligned16(gx);
  int rc;
  switch (w) {
    case 1: rc = launch_bwd_w<1>(p, pairs, st); break;
    case 2: rc = launch_bwd_w<2>(p, pairs, st); break;
    default: rc = launch_bwd_w<4>(p, pairs, st); break;
  }
  if (rc != DASP_OK) return rc;
  const int64_t tot = bs * kSections;
  eq_param_grad_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(ws, params, gparams, bs, (int)chs, sample_rate);
  DASP_LAUNCH_OK("eq_param_grad_kernel");
  return DASP_OK;
}

}  // extern "C"
