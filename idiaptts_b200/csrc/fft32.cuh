// Single-precision shared-memory complex FFT for one CTA, radix-16 Stockham passes on packed f32x2 arithmetic.
//
// Used by the fast D4C path (d4c_fast.cu).  Differences to fft.cuh (the fp64 / generic version):
//   * every complex add / subtract is ONE FADD2 and every complex multiply TWO instructions (FMUL2 + FFMA2 with operand
//     swizzles), through inline `add/mul/fma.rn.f32x2` PTX -- nvcc does not form the packed instructions by itself;
//   * N = 16 * 16 * R3 (R3 = 4, 8, 16 for N = 1024, 2048, 4096) with NT = N / 16 threads: three passes, the first one fed from
//     registers (fft32_first_pass), so a transform makes 2.5 shared-memory round trips instead of 3.5;
//   * buffer layout: logical index i lives at ZQ(i) = i + (i >> 4): with 8-byte elements the stride-16 stores of the first
//     pass then hit all 32 banks once per half-warp.
// All transforms are forward (exp(-i ...)).
#pragma once
#include "common.cuh"

namespace b2w {
namespace f32 {

typedef unsigned long long u64;

__host__ __device__ constexpr int ZQ(int i) { return i + (i >> 4); }
__host__ __device__ constexpr int zq_size(int M) { return M + (M >> 4) + 1; }

__device__ __forceinline__ u64 as64(float2 a) { return *reinterpret_cast<u64*>(&a); }
__device__ __forceinline__ float2 as2(u64 a) { return *reinterpret_cast<float2*>(&a); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
  u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(as64(a)), "l"(as64(b)));
  return as2(r);
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
  u64 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(as64(a)), "l"(as64(b)));
  return as2(r);
}
__device__ __forceinline__ float2 pmul(float2 a, float2 b) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(as64(a)), "l"(as64(b)));
  return as2(r);
}
__device__ __forceinline__ float2 pfma(float2 a, float2 b, float2 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(as64(a)), "l"(as64(b)), "l"(as64(c)));
  return as2(r);
}
// a * w: (ax wx - ay wy, ax wy + ay wx); ptxas folds the broadcasts, the swap and the half negation into operand modifiers
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
  const float2 t = pmul(float2{a.y, a.y}, float2{w.y, w.x});
  return pfma(float2{a.x, a.x}, w, float2{-t.x, t.y});
}
__device__ __forceinline__ float2 cmul_mi(float2 a) { return float2{a.y, -a.x}; }  // * (-i), folded into the consumer
__device__ __forceinline__ float2 cscale(float2 a, float s) { return pmul(a, float2{s, s}); }

__device__ __forceinline__ void dft4(float2& v0, float2& v1, float2& v2, float2& v3) {
  const float2 a0 = cadd(v0, v2), a1 = csub(v0, v2);
  const float2 a2 = cadd(v1, v3), a3 = cmul_mi(csub(v1, v3));
  v0 = cadd(a0, a2);
  v1 = cadd(a1, a3);
  v2 = csub(a0, a2);
  v3 = csub(a1, a3);
}
__device__ __forceinline__ void dft4(float2* v) { dft4(v[0], v[1], v[2], v[3]); }

// natural order in, natural order out
__device__ __forceinline__ void dft8(float2* v) {
  const float h = 0.70710678118654752440f;
  float2 a0 = cadd(v[0], v[4]), b0 = csub(v[0], v[4]);
  float2 a1 = cadd(v[1], v[5]), b1 = csub(v[1], v[5]);
  float2 a2 = cadd(v[2], v[6]), b2 = csub(v[2], v[6]);
  float2 a3 = cadd(v[3], v[7]), b3 = csub(v[3], v[7]);
  b1 = cscale(cadd(b1, cmul_mi(b1)), h);             // * (1 - i) / sqrt2
  b2 = cmul_mi(b2);
  b3 = cscale(csub(cmul_mi(b3), b3), h);             // * (-1 - i) / sqrt2
  dft4(a0, a1, a2, a3);
  dft4(b0, b1, b2, b3);
  v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
  v[1] = b0; v[3] = b1; v[5] = b2; v[7] = b3;
}

// 16 = 4 x 4: n = 4 n1 + n2, k = k1 + 4 k2; inner DFTs over n1, twiddles W16^(n2 k1), outer DFTs over n2
__device__ __forceinline__ void dft16(float2* v) {
  const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
  // inner: for every n2 the 4-point DFT of v[n2], v[4 + n2], v[8 + n2], v[12 + n2] -> y[n2][k1] left in v[4 k1 + n2]
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) dft4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
  // twiddles W16^(n2 k1) on v[4 k1 + n2]
  v[5] = cmul(v[5], float2{c1, -s1});                       // W^1
  v[6] = cscale(cadd(v[6], cmul_mi(v[6])), h);              // W^2 = (1 - i) / sqrt2
  v[7] = cmul(v[7], float2{s1, -c1});                       // W^3
  v[9] = cscale(cadd(v[9], cmul_mi(v[9])), h);              // W^2
  v[10] = cmul_mi(v[10]);                                   // W^4 = -i
  v[11] = cscale(csub(cmul_mi(v[11]), v[11]), h);           // W^6 = (-1 - i) / sqrt2
  v[13] = cmul(v[13], float2{s1, -c1});                     // W^3
  v[14] = cscale(csub(cmul_mi(v[14]), v[14]), h);           // W^6
  v[15] = cmul(v[15], float2{-c1, s1});                     // W^9
  // outer: for every k1 the 4-point DFT over n2 of v[4 k1 + n2] -> X[k1 + 4 k2] left in v[4 k1 + k2]
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
  // natural order: X[k1 + 4 k2] sits in v[4 k1 + k2] -> transpose the 4 x 4 register tile (renaming only)
  float2 t;
  t = v[1]; v[1] = v[4]; v[4] = t;
  t = v[2]; v[2] = v[8]; v[8] = t;
  t = v[3]; v[3] = v[12]; v[12] = t;
  t = v[6]; v[6] = v[9]; v[9] = t;
  t = v[7]; v[7] = v[13]; v[13] = t;
  t = v[11]; v[11] = v[14]; v[14] = t;
}

template <int R> __device__ __forceinline__ void dftR(float2* v) {
  if (R == 16) dft16(v);
  else if (R == 8) dft8(v);
  else dft4(v);
}

template <int N> struct Plan;  // N = 16 * 16 * R3
template <> struct Plan<1024> { static constexpr int R3 = 4; };
template <> struct Plan<2048> { static constexpr int R3 = 8; };
template <> struct Plan<4096> { static constexpr int R3 = 16; };

// Per-CTA twiddle table: [0, 16) exp(-2 pi i k / 256), k < 16 (second pass); [16, 272) exp(-2 pi i k / N), k < 256 (third pass).
constexpr int kTwEntries = 16 + 256;

template <int N, int NT>
__device__ __forceinline__ void tw_fill(float2* tws, const double2* __restrict__ tw, int tid) {
  for (int e = tid; e < kTwEntries; e += NT) {
    const int idx = e < 16 ? e * (kTwN / 256) : (e - 16) * (kTwN / N);
    const double2 w = __ldg(&tw[idx]);
    tws[e] = float2{(float)w.x, (float)w.y};
  }
}

// v[q] *= w^q, q = 1 .. R-1, powers formed by a multiplication tree of depth <= 4
template <int R>
__device__ __forceinline__ void apply_twiddles(float2* v, float2 w1) {
  const float2 w2 = cmul(w1, w1);
  const float2 w3 = cmul(w2, w1);
  v[1] = cmul(v[1], w1);
  v[2] = cmul(v[2], w2);
  v[3] = cmul(v[3], w3);
  if (R >= 8) {
    const float2 w4 = cmul(w2, w2);
    v[4] = cmul(v[4], w4);
    v[5] = cmul(v[5], cmul(w4, w1));
    v[6] = cmul(v[6], cmul(w4, w2));
    const float2 w7 = cmul(w4, w3);
    v[7] = cmul(v[7], w7);
    if (R == 16) {
      const float2 w8 = cmul(w4, w4);
      v[8] = cmul(v[8], w8);
      v[9] = cmul(v[9], cmul(w8, w1));
      v[10] = cmul(v[10], cmul(w8, w2));
      v[11] = cmul(v[11], cmul(w8, w3));
      v[12] = cmul(v[12], cmul(w8, w4));
      v[13] = cmul(v[13], cmul(w8, cmul(w4, w1)));
      v[14] = cmul(v[14], cmul(w8, cmul(w4, w2)));
      v[15] = cmul(v[15], cmul(w8, w7));
    }
  }
}

// Address arithmetic: every index handed to ZQ below has the form base + c with c a multiple of 16, so ZQ(base + c) =
// ZQ(base) + c + c / 16 -- written out by hand (one base per thread and pass, constant offsets), the compiler cannot prove it.

// First pass (sub-length 1) with inputs already in registers: v[q] = x[tid + q * NT], NT = N / 16.  The caller guarantees
// that no thread still reads z (a __syncthreads() since the last read).  Ends with a barrier.
template <int N, int NT>
__device__ __forceinline__ void fft32_first_pass(float2* z, float2* v, int tid) {
  static_assert(N == 16 * NT, "one radix-16 butterfly per thread");
  dft16(v);
  float2* zo = z + 17 * tid;  // ZQ(16 tid + q) = 17 tid + q
#pragma unroll
  for (int q = 0; q < 16; ++q) zo[q] = v[q];
  __syncthreads();
}

// Passes 2 and 3.  On return the spectrum is in z (natural order) and visible to all threads.
template <int N, int NT>
__device__ __forceinline__ void fft32_tail(float2* z, const float2* tws, int tid) {
  constexpr int R3 = Plan<N>::R3;
  const float2* zi = z + tid + (tid >> 4);  // ZQ(tid)
  {  // pass 2: radix 16, sub-length 16, one butterfly per thread
    float2 v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = zi[q * (NT + NT / 16)];
    __syncthreads();
    const int k = tid & 15;
    apply_twiddles<16>(v, tws[k]);
    dft16(v);
    float2* zo = z + 17 * (tid - k) + k;  // j0 = 16 (tid - k) + k, ZQ(j0 + 16 q) = j0 + (tid - k) + 17 q
#pragma unroll
    for (int q = 0; q < 16; ++q) zo[17 * q] = v[q];
    __syncthreads();
  }
  {  // pass 3: radix R3, sub-length 256, 16 / R3 butterflies per thread
    constexpr int NB = N / R3;
    constexpr int BPT = NB / NT;
    float2 v[BPT][R3];
#pragma unroll
    for (int b = 0; b < BPT; ++b) {
#pragma unroll
      for (int q = 0; q < R3; ++q) v[b][q] = zi[(b * NT + q * NB) / 16 * 17];
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < BPT; ++b) {
      const int j = tid + b * NT;
      const int k = j & 255;
      apply_twiddles<R3>(v[b], tws[16 + k]);
      dftR<R3>(v[b]);
      const int j0 = (j - k) * R3 + k;
      float2* zo = z + j0 + (j0 >> 4);
#pragma unroll
      for (int q = 0; q < R3; ++q) zo[272 * q] = v[b][q];
    }
    __syncthreads();
  }
}

// After the FFT of z = x1 + i x2 (two real sequences of length N): X1[k] and X2[k], k in [0, N/2], from a = Z[k] and
// bq = Z[(N - k) mod N].
__device__ __forceinline__ void split_pair(float2 a, float2 bq, float2& x1, float2& x2) {
  const float2 b = float2{bq.x, -bq.y};
  x1 = cscale(cadd(a, b), 0.5f);
  x2 = cmul_mi(cscale(csub(a, b), 0.5f));
}
// Walks the bins k = tid, tid + NT, ... <= N / 2 of a packed pair transform: f(k, X1[k], X2[k]).  Both addresses move by
// constant strides (ZQ(k + NT) = ZQ(k) + NT + NT / 16; the mirrored bin the other way; bin 0 mirrors onto itself).
template <int N, int NT, typename F>
__device__ __forceinline__ void for_pair_bins(const float2* z, int tid, F f) {
  constexpr int ST = NT + NT / 16;
  const float2* za = z + tid + (tid >> 4);
  const int mir = (N - tid) & (N - 1);
  const float2* zb = z + mir + (mir >> 4);
#pragma unroll
  for (int j = 0; j <= 8; ++j) {
    if (j == 8 && tid != 0) break;  // bin N / 2 belongs to thread 0
    float2 x1, x2;
    // tid == 0: mirror of bin 0 is bin 0 (j = 0), afterwards N - NT j = ZQ offset (N - NT j) * 17 / 16
    const float2 bq = (tid == 0) ? z[j == 0 ? 0 : (N - j * NT) / 16 * 17] : zb[-j * ST];
    split_pair(za[j * ST], bq, x1, x2);
    f(tid + j * NT, x1, x2);
  }
}

}  // namespace f32
}  // namespace b2w
