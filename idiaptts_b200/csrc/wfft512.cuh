// One-warp transforms of 512 complex points (= 1024 real points through the packed half-size trick), the building block of the
// warp-per-unit kernels (synth_fast.cu: one pulse per warp; cheaptrick_fast.cu: one frame per warp).  Radix 16 x 16 x 2 Stockham
// passes with sixteen points per lane, the first pass fed from registers; only __syncwarp between passes.  Single precision on
// packed f32x2 arithmetic (fft32.cuh) and a double-precision twin for the one transform whose rounding noise matters
// (CheapTrick's waveform -> power spectrum: weak bins sit 80 dB under the strong ones).
#pragma once
#include "fft32.cuh"

namespace b2w {
namespace w512 {

using f32::ZQ;
using f32::cadd;
using f32::cmul;
using f32::cmul_mi;
using f32::cscale;
using f32::csub;

constexpr int kN = 1024;   // real length
constexpr int kM = kN / 2; // complex length
constexpr int kK = kM + 1; // bins

// Forward complex FFT of 512 points by one warp.  v[q] = x[lane + 32 q] on entry; the spectrum ends up in z (natural order,
// padded layout ZQ).  The caller guarantees that no lane still reads z.
// Passes two and three, out of line: the warp-per-unit kernels call the transform five to seven times per unit, and with a few
// warps per sub-partition at different points of a 150 KB instruction stream they stall on instruction fetch (ncu: "no
// instruction" 26 % of the stalls of render_fast_kernel).  One shared copy of the big block keeps the kernel inside the
// instruction cache; only the register-fed first pass stays inline.
static __device__ __noinline__ void wfft512_tail(float2* z, const float2* tw16, const float2* tw512, int lane) {
  float2 v[16];
  float2* zi = z + lane + (lane >> 4);  // ZQ(lane)
#pragma unroll
  for (int q = 0; q < 16; ++q) v[q] = zi[34 * q];  // ZQ(lane + 32 q)
  __syncwarp();
  {
    const int k = lane & 15;
    f32::apply_twiddles<16>(v, tw16[k]);
    f32::dft16(v);
    float2* zo = z + 17 * (lane - k) + k;  // ZQ(16 (lane - k) + k + 16 q) = ... + 17 q
#pragma unroll
    for (int q = 0; q < 16; ++q) zo[17 * q] = v[q];
  }
  __syncwarp();
#pragma unroll
  for (int b = 0; b < 8; ++b) {  // radix 2, sub-length 256: (j, j + 256), in place
    float2* pa = zi + 34 * b;
    const float2 a = pa[0];
    const float2 t = cmul(pa[272], tw512[lane + 32 * b]);
    pa[0] = cadd(a, t);
    pa[272] = csub(a, t);
  }
  __syncwarp();
}
__device__ __forceinline__ void wfft512(float2* z, float2* v, const float2* tw16, const float2* tw512, int lane) {
  f32::dft16(v);
  {
    float2* zo = z + 17 * lane;  // ZQ(16 lane + q)
#pragma unroll
    for (int q = 0; q < 16; ++q) zo[q] = v[q];
  }
  __syncwarp();
  wfft512_tail(z, tw16, tw512, lane);
}

// Walks the bins k = lane + 32 j <= 512 of the length-1024 REAL transform whose packed half-size transform sits in z:
// f(k, X[k]).  X[k] = E + w^k O with E = (Z[k] + conj Z[M - k]) / 2, O = (Z[k] - conj Z[M - k]) / (2 i).
template <typename F>
__device__ __forceinline__ void for_real_bins(const float2* z, const float2* twn, int lane, F f) {
  const float2* za = z + lane + (lane >> 4);
  const int mir = (kM - lane) & (kM - 1);
  const float2* zb = z + mir + (mir >> 4);
#pragma unroll
  for (int j = 0; j <= 16; ++j) {
    if (j == 16 && lane != 0) break;  // bin M belongs to lane 0
    const int k = lane + 32 * j;
    const float2 a = (j == 16) ? z[0] : za[34 * j];
    const float2 bq = (lane == 0) ? z[(j == 0 || j == 16) ? 0 : (kM - 32 * j) / 16 * 17] : zb[-34 * j];
    const float2 b = float2{bq.x, -bq.y};
    const float2 e = cscale(cadd(a, b), 0.5f);
    const float2 o = cmul_mi(cscale(csub(a, b), 0.5f));
    const float2 w = (j == 16) ? float2{-1.0f, 0.0f} : twn[k];
    f(k, cadd(e, cmul(w, o)));
  }
}


// ---- double precision twin ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 dadd(double2 a, double2 b) { return double2{a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ double2 dsub(double2 a, double2 b) { return double2{a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ double2 dmul(double2 a, double2 w) { return double2{a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x}; }
__device__ __forceinline__ double2 dmul_mi(double2 a) { return double2{a.y, -a.x}; }
__device__ __forceinline__ void ddft4(double2& v0, double2& v1, double2& v2, double2& v3) {
  const double2 a0 = dadd(v0, v2), a1 = dsub(v0, v2);
  const double2 a2 = dadd(v1, v3), a3 = dmul_mi(dsub(v1, v3));
  v0 = dadd(a0, a2);
  v1 = dadd(a1, a3);
  v2 = dsub(a0, a2);
  v3 = dsub(a1, a3);
}
__device__ __forceinline__ void ddft16(double2* v) {
  const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173, h = 0.70710678118654752440;
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) ddft4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
  v[5] = dmul(v[5], double2{c1, -s1});
  v[6] = double2{h * (v[6].x + v[6].y), h * (v[6].y - v[6].x)};
  v[7] = dmul(v[7], double2{s1, -c1});
  v[9] = double2{h * (v[9].x + v[9].y), h * (v[9].y - v[9].x)};
  v[10] = dmul_mi(v[10]);
  v[11] = double2{h * (v[11].y - v[11].x), -h * (v[11].x + v[11].y)};
  v[13] = dmul(v[13], double2{s1, -c1});
  v[14] = double2{h * (v[14].y - v[14].x), -h * (v[14].x + v[14].y)};
  v[15] = dmul(v[15], double2{-c1, s1});
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) ddft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
  double2 t;
  t = v[1]; v[1] = v[4]; v[4] = t;
  t = v[2]; v[2] = v[8]; v[8] = t;
  t = v[3]; v[3] = v[12]; v[12] = t;
  t = v[6]; v[6] = v[9]; v[9] = t;
  t = v[7]; v[7] = v[13]; v[13] = t;
  t = v[11]; v[11] = v[14]; v[14] = t;
}
// Same contract as wfft512 with double2 data.  tw256: exp(-2 pi i k / 256), k < 256 (the second pass reads the sixteen powers
// q k directly: no multiplication chain, full accuracy); tw512: exp(-2 pi i k / 512), k < 256.
__device__ __forceinline__ void wfft512_f64(double2* z, double2* v, const double2* tw256, const double2* tw512, int lane) {
  ddft16(v);
  {
    double2* zo = z + 17 * lane;
#pragma unroll
    for (int q = 0; q < 16; ++q) zo[q] = v[q];
  }
  __syncwarp();
  double2* zi = z + lane + (lane >> 4);
#pragma unroll
  for (int q = 0; q < 16; ++q) v[q] = zi[34 * q];
  __syncwarp();
  {
    const int k = lane & 15;
#pragma unroll
    for (int q = 1; q < 16; ++q) v[q] = dmul(v[q], tw256[q * k]);
    ddft16(v);
    double2* zo = z + 17 * (lane - k) + k;
#pragma unroll
    for (int q = 0; q < 16; ++q) zo[17 * q] = v[q];
  }
  __syncwarp();
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    double2* pa = zi + 34 * b;
    const double2 a = pa[0];
    const double2 t = dmul(pa[272], tw512[lane + 32 * b]);
    pa[0] = dadd(a, t);
    pa[272] = dsub(a, t);
  }
  __syncwarp();
}
// f(k, X[k]) for the bins k = lane + 32 j <= 512 of the real transform (double); twn: exp(-2 pi i k / 1024), k < 512
template <typename F>
__device__ __forceinline__ void for_real_bins_f64(const double2* z, const double2* twn, int lane, F f) {
  const double2* za = z + lane + (lane >> 4);
  const int mir = (kM - lane) & (kM - 1);
  const double2* zb = z + mir + (mir >> 4);
#pragma unroll
  for (int j = 0; j <= 16; ++j) {
    if (j == 16 && lane != 0) break;
    const int k = lane + 32 * j;
    const double2 a = (j == 16) ? z[0] : za[34 * j];
    const double2 bq = (lane == 0) ? z[(j == 0 || j == 16) ? 0 : (kM - 32 * j) / 16 * 17] : zb[-34 * j];
    const double2 b = double2{bq.x, -bq.y};
    const double2 e = double2{0.5 * (a.x + b.x), 0.5 * (a.y + b.y)};
    const double2 o = dmul_mi(double2{0.5 * (a.x - b.x), 0.5 * (a.y - b.y)});
    const double2 w = (j == 16) ? double2{-1.0, 0.0} : twn[k];
    f(k, dadd(e, dmul(w, o)));
  }
}

}  // namespace w512
}  // namespace b2w
