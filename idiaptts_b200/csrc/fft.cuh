// Shared-memory complex FFT for one CTA (Stockham autosort, radix-8/4/2 butterflies held in registers,
// in-place in a padded shared buffer), plus the two "real input" post-passes WORLD needs.
//
//   buffer layout: M complex points, logical index i stored at ZP(i) = i + (i >> 3) (one pad element every 8
//   complex values keeps both the stride-R stores of the first pass and the strided loads conflict-free for
//   16-byte elements).  Size in elements: zp_size(M).
//
// All transforms are forward (exp(-i...)).  Twiddles come from the global table tw[k] = exp(-2 pi i k / kTwN).
#pragma once
#include "common.cuh"

namespace b2w {

template <typename T> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

__host__ __device__ constexpr int ZP(int i) { return i + (i >> 3); }
__host__ __device__ constexpr int zp_size(int M) { return M + (M >> 3) + 1; }

template <typename V> __device__ __forceinline__ V cadd(V a, V b) { return V{a.x + b.x, a.y + b.y}; }
template <typename V> __device__ __forceinline__ V csub(V a, V b) { return V{a.x - b.x, a.y - b.y}; }
template <typename V> __device__ __forceinline__ V cmul(V a, V b) {
  return V{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
// multiply by -i (forward DFT quarter turn)
template <typename V> __device__ __forceinline__ V cmul_mi(V a) { return V{a.y, -a.x}; }

template <typename T, typename V>
__device__ __forceinline__ V tw_load(const double2* __restrict__ tw, int idx) {
  double2 w = __ldg(&tw[idx]);
  return V{(T)w.x, (T)w.y};
}

// natural-order in, natural-order out forward DFTs of 2, 4, 8 points
template <typename V> __device__ __forceinline__ void dft2(V* v) {
  V a = v[0], b = v[1];
  v[0] = cadd(a, b);
  v[1] = csub(a, b);
}
template <typename V> __device__ __forceinline__ void dft4(V* v) {
  V a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]);
  V a2 = cadd(v[1], v[3]), a3 = cmul_mi(csub(v[1], v[3]));
  v[0] = cadd(a0, a2);
  v[1] = cadd(a1, a3);
  v[2] = csub(a0, a2);
  v[3] = csub(a1, a3);
}
template <typename T, typename V> __device__ __forceinline__ void dft8(V* v) {
  const T h = (T)0.70710678118654752440;
  // decimation in frequency: first stage pairs (p, p+4)
  V a0 = cadd(v[0], v[4]), b0 = csub(v[0], v[4]);
  V a1 = cadd(v[1], v[5]), b1 = csub(v[1], v[5]);
  V a2 = cadd(v[2], v[6]), b2 = csub(v[2], v[6]);
  V a3 = cadd(v[3], v[7]), b3 = csub(v[3], v[7]);
  // twiddles W8^p on the odd branch: W8^1 = (1 - i)/sqrt2, W8^2 = -i, W8^3 = (-1 - i)/sqrt2
  b1 = V{h * (b1.x + b1.y), h * (b1.y - b1.x)};
  b2 = cmul_mi(b2);
  b3 = V{h * (b3.y - b3.x), -h * (b3.x + b3.y)};
  V e[4] = {a0, a1, a2, a3};
  V o[4] = {b0, b1, b2, b3};
  dft4(e);
  dft4(o);
  v[0] = e[0]; v[2] = e[1]; v[4] = e[2]; v[6] = e[3];
  v[1] = o[0]; v[3] = o[1]; v[5] = o[2]; v[7] = o[3];
}

template <typename T, int R, typename V> __device__ __forceinline__ void dftR(V* v) {
  if (R == 8) dft8<T, V>(v);
  else if (R == 4) dft4<V>(v);
  else dft2<V>(v);
}

// One Stockham pass of radix R on M points where the already-transformed sub-length is Ns.
// Every thread first pulls all of its butterflies' inputs into registers, the CTA synchronises, then results are
// scattered back into the same buffer (in-place autosort).
template <typename T, int M, int NT, int R, int Ns>
__device__ __forceinline__ void fft_pass(typename Vec2<T>::type* z, const double2* __restrict__ tw, int tid) {
  using V = typename Vec2<T>::type;
  constexpr int NB = M / R;                     // butterflies in this pass
  constexpr int BPT = (NB + NT - 1) / NT;       // per thread
  V v[BPT][R];
#pragma unroll
  for (int b = 0; b < BPT; ++b) {
    const int j = tid + b * NT;
    if (NB % NT == 0 || j < NB) {
#pragma unroll
      for (int q = 0; q < R; ++q) v[b][q] = z[ZP(j + q * NB)];
    }
  }
  __syncthreads();
#pragma unroll
  for (int b = 0; b < BPT; ++b) {
    const int j = tid + b * NT;
    if (NB % NT == 0 || j < NB) {
      const int k = j & (Ns - 1);
      if (Ns > 1) {
#pragma unroll
        for (int q = 1; q < R; ++q) v[b][q] = cmul(v[b][q], tw_load<T, V>(tw, (q * k) * (kTwN / (Ns * R))));
      }
      dftR<T, R, V>(v[b]);
      const int j0 = (j - k) * R + k;
#pragma unroll
      for (int q = 0; q < R; ++q) z[ZP(j0 + q * Ns)] = v[b][q];
    }
  }
  __syncthreads();
}

// Forward complex FFT of M points held in z (padded layout).  Caller must __syncthreads() after filling z; on return
// the result is visible to all threads.
template <typename T, int M, int NT>
__device__ __forceinline__ void cfft(typename Vec2<T>::type* z, const double2* __restrict__ tw, int tid) {
  static_assert(M == 256 || M == 512 || M == 1024 || M == 2048 || M == 4096, "unsupported FFT length");
  if (M == 256) {
    fft_pass<T, M, NT, 8, 1>(z, tw, tid);
    fft_pass<T, M, NT, 8, 8>(z, tw, tid);
    fft_pass<T, M, NT, 4, 64>(z, tw, tid);
  } else if (M == 512) {
    fft_pass<T, M, NT, 8, 1>(z, tw, tid);
    fft_pass<T, M, NT, 8, 8>(z, tw, tid);
    fft_pass<T, M, NT, 8, 64>(z, tw, tid);
  } else if (M == 1024) {
    fft_pass<T, M, NT, 8, 1>(z, tw, tid);
    fft_pass<T, M, NT, 8, 8>(z, tw, tid);
    fft_pass<T, M, NT, 4, 64>(z, tw, tid);
    fft_pass<T, M, NT, 4, 256>(z, tw, tid);
  } else if (M == 2048) {
    fft_pass<T, M, NT, 8, 1>(z, tw, tid);
    fft_pass<T, M, NT, 8, 8>(z, tw, tid);
    fft_pass<T, M, NT, 8, 64>(z, tw, tid);
    fft_pass<T, M, NT, 4, 512>(z, tw, tid);
  } else {
    fft_pass<T, M, NT, 8, 1>(z, tw, tid);
    fft_pass<T, M, NT, 8, 8>(z, tw, tid);
    fft_pass<T, M, NT, 8, 64>(z, tw, tid);
    fft_pass<T, M, NT, 8, 512>(z, tw, tid);
  }
}

// ---- shared-memory twiddles ---------------------------------------------------------------------------------------------------
// The FFT kernels are bound by the L1/LSU data pipe, and a gathered 16-byte global load costs one 32-byte sector per lane.
// cfft_s therefore reads ONE twiddle per butterfly, w = exp(-2 pi i k / (Ns R)), from a small per-CTA shared-memory table
// (contiguous in k: 4 wavefronts per warp) and forms w^2 .. w^(R-1) by multiplication (a few ulp, on the idle fp64 pipe).
//   layout (complex entries):  [0, 8)    pass with Ns = 8      w(k),  k < 8
//                              [8, 72)   pass with Ns = 64     w(k),  k < 64
//                              [72, 136) pass with Ns >= 256   w(8 a), a < Ns / 8      (two-level: w(k) = w(8 a) w(b))
//                              [136,144) pass with Ns >= 256   w(b),  b < 8
constexpr int kFftTwEntries = 144;

template <int M> struct FftPlan;  // radices of the passes, first to last
template <> struct FftPlan<256> { static constexpr int n = 3; static constexpr int R[4] = {8, 8, 4, 1}; };
template <> struct FftPlan<512> { static constexpr int n = 3; static constexpr int R[4] = {8, 8, 8, 1}; };
template <> struct FftPlan<1024> { static constexpr int n = 4; static constexpr int R[4] = {8, 8, 4, 4}; };
template <> struct FftPlan<2048> { static constexpr int n = 4; static constexpr int R[4] = {8, 8, 8, 4}; };
template <> struct FftPlan<4096> { static constexpr int n = 4; static constexpr int R[4] = {8, 8, 8, 8}; };

// Fills the table for cfft_s<T, M>; call once per CTA (all threads), followed by a __syncthreads() before the first FFT.
template <typename T, int M, int NT>
__device__ __forceinline__ void fft_tw_fill(typename Vec2<T>::type* tws, const double2* __restrict__ tw, int tid) {
  using V = typename Vec2<T>::type;
  using P = FftPlan<M>;
  constexpr int R1 = P::R[1], R2 = P::R[2], R3 = P::R[3], Ns3 = P::R[0] * P::R[1] * P::R[2];
  for (int e = tid; e < kFftTwEntries; e += NT) {
    int idx = 0;  // index into the global table of kTwN-th roots
    if (e < 8) idx = e * (kTwN / (8 * R1));
    else if (e < 72) idx = (e - 8) * (kTwN / (64 * R2));
    else if (P::n == 4) {
      constexpr int step = kTwN / (Ns3 * R3);
      idx = e < 136 ? ((e - 72) * 8 < Ns3 ? (e - 72) * 8 * step : 0) : (e - 136) * step;
    }
    tws[e] = tw_load<T, V>(tw, idx);
  }
}

template <typename T, int R, int Ns, typename V>
__device__ __forceinline__ V fft_tw_w1(const V* tws, int k) {
  if (Ns == 8) return tws[k];
  if (Ns == 64) return tws[8 + k];
  return cmul(tws[72 + (k >> 3)], tws[136 + (k & 7)]);
}

template <typename T, int M, int NT, int R, int Ns>
__device__ __forceinline__ void fft_pass_s(typename Vec2<T>::type* z, const typename Vec2<T>::type* tws, int tid) {
  using V = typename Vec2<T>::type;
  constexpr int NB = M / R;
  constexpr int BPT = (NB + NT - 1) / NT;
  V v[BPT][R];
#pragma unroll
  for (int b = 0; b < BPT; ++b) {
    const int j = tid + b * NT;
    if (NB % NT == 0 || j < NB) {
#pragma unroll
      for (int q = 0; q < R; ++q) v[b][q] = z[ZP(j + q * NB)];
    }
  }
  __syncthreads();
#pragma unroll
  for (int b = 0; b < BPT; ++b) {
    const int j = tid + b * NT;
    if (NB % NT == 0 || j < NB) {
      const int k = j & (Ns - 1);
      if (Ns > 1) {
        const V w1 = fft_tw_w1<T, R, Ns, V>(tws, k);
        const V w2 = cmul(w1, w1);
        v[b][1] = cmul(v[b][1], w1);
        v[b][2] = cmul(v[b][2], w2);
        const V w3 = cmul(w2, w1);
        v[b][3] = cmul(v[b][3], w3);
        if (R == 8) {
          const V w4 = cmul(w2, w2);
          v[b][4] = cmul(v[b][4], w4);
          v[b][5] = cmul(v[b][5], cmul(w4, w1));
          v[b][6] = cmul(v[b][6], cmul(w3, w3));
          v[b][7] = cmul(v[b][7], cmul(w4, w3));
        }
      }
      dftR<T, R, V>(v[b]);
      const int j0 = (j - k) * R + k;
#pragma unroll
      for (int q = 0; q < R; ++q) z[ZP(j0 + q * Ns)] = v[b][q];
    }
  }
  __syncthreads();
}

// First Stockham pass (Ns = 1) of an M-point FFT whose inputs are already in registers: v[q] = x[tid + q * NT] with
// NT == M / R (one butterfly per thread).  Producers (windowing) therefore never stage the FFT input in shared memory and the
// pass has no loads.  The caller guarantees that no thread still reads z (a __syncthreads() since the last read).
template <typename T, int M, int NT>
__device__ __forceinline__ void fft_first_pass_regs(typename Vec2<T>::type* z, typename Vec2<T>::type* v, int tid) {
  using V = typename Vec2<T>::type;
  constexpr int R = FftPlan<M>::R[0];
  static_assert(M / R == NT, "one first-pass butterfly per thread");
  dftR<T, R, V>(v);
#pragma unroll
  for (int q = 0; q < R; ++q) z[ZP(tid * R + q)] = v[q];
  __syncthreads();
}

// Passes 2.. of cfft_s after fft_first_pass_regs.
template <typename T, int M, int NT>
__device__ __forceinline__ void cfft_s_tail(typename Vec2<T>::type* z, const typename Vec2<T>::type* tws, int tid) {
  using P = FftPlan<M>;
  constexpr int R0 = P::R[0], R1 = P::R[1], R2 = P::R[2], R3 = P::R[3];
  fft_pass_s<T, M, NT, R1, R0>(z, tws, tid);
  fft_pass_s<T, M, NT, R2, R0 * R1>(z, tws, tid);
  if constexpr (P::n == 4) fft_pass_s<T, M, NT, R3, R0 * R1 * R2>(z, tws, tid);
}

// cfft with shared-memory twiddles (see above); same contract as cfft.
template <typename T, int M, int NT>
__device__ __forceinline__ void cfft_s(typename Vec2<T>::type* z, const typename Vec2<T>::type* tws, int tid) {
  using P = FftPlan<M>;
  constexpr int R0 = P::R[0], R1 = P::R[1], R2 = P::R[2], R3 = P::R[3];
  fft_pass_s<T, M, NT, R0, 1>(z, tws, tid);
  fft_pass_s<T, M, NT, R1, R0>(z, tws, tid);
  fft_pass_s<T, M, NT, R2, R0 * R1>(z, tws, tid);
  if constexpr (P::n == 4) fft_pass_s<T, M, NT, R3, R0 * R1 * R2>(z, tws, tid);
}

// Phasor pair for the real-FFT post-pass when bins are visited as k = tid, tid + NT, ...: w = exp(-2 pi i tid / N) and the
// stride exp(-2 pi i NT / N); rfft_bin_w takes the running w, the caller advances it with cmul(w, step).
template <typename T, int N, int NT>
__device__ __forceinline__ void rfft_rot_init(const double2* __restrict__ tw, int tid, typename Vec2<T>::type& w0,
                                              typename Vec2<T>::type& step) {
  using V = typename Vec2<T>::type;
  w0 = tw_load<T, V>(tw, tid * (kTwN / N));
  step = tw_load<T, V>(tw, NT * (kTwN / N));
}

template <typename T, int N>
__device__ __forceinline__ typename Vec2<T>::type rfft_bin_w(const typename Vec2<T>::type* z, int k,
                                                             typename Vec2<T>::type w) {
  using V = typename Vec2<T>::type;
  constexpr int M = N / 2;
  const V a = z[ZP(k & (M - 1))];
  const V bq = z[ZP((M - k) & (M - 1))];
  const V b = V{bq.x, -bq.y};
  const V e = V{(T)0.5 * (a.x + b.x), (T)0.5 * (a.y + b.y)};
  const V d = V{(T)0.5 * (a.x - b.x), (T)0.5 * (a.y - b.y)};
  return cadd(e, cmul(w, cmul_mi(d)));
}

// Real sequence x[0..N) stored as N reals: real index n lives in complex slot n>>1 (component n&1) of the padded buffer.
template <typename T>
__device__ __forceinline__ T& zreal(typename Vec2<T>::type* z, int n) {
  return reinterpret_cast<T*>(&z[ZP(n >> 1)])[n & 1];
}

// After cfft<T, N/2> of the packed real sequence: X[k], k in [0, N/2], of the length-N real DFT.
template <typename T, int N>
__device__ __forceinline__ typename Vec2<T>::type rfft_bin(const typename Vec2<T>::type* z, const double2* __restrict__ tw,
                                                           int k) {
  using V = typename Vec2<T>::type;
  constexpr int M = N / 2;
  const V a = z[ZP(k & (M - 1))];
  const V bq = z[ZP((M - k) & (M - 1))];
  const V b = V{bq.x, -bq.y};                         // conj(Z[M-k])
  const V e = V{(T)0.5 * (a.x + b.x), (T)0.5 * (a.y + b.y)};
  const V d = V{(T)0.5 * (a.x - b.x), (T)0.5 * (a.y - b.y)};
  const V o = cmul_mi(d);                             // (A - B) / (2i)
  const V w = tw_load<T, V>(tw, k * (kTwN / N));      // exp(-2 pi i k / N); k = N/2 -> index kTwN/2 (= -1)
  return cadd(e, cmul(w, o));
}

// After cfft<T, N> of z = x1 + i*x2 (two real sequences of length N): X1[k] and X2[k], k in [0, N/2].
template <typename T, int N>
__device__ __forceinline__ void two_real_bins(const typename Vec2<T>::type* z, int k, typename Vec2<T>::type& x1,
                                              typename Vec2<T>::type& x2) {
  using V = typename Vec2<T>::type;
  const V a = z[ZP(k)];
  const V bq = z[ZP((N - k) & (N - 1))];
  const V b = V{bq.x, -bq.y};
  x1 = V{(T)0.5 * (a.x + b.x), (T)0.5 * (a.y + b.y)};
  const V d = V{(T)0.5 * (a.x - b.x), (T)0.5 * (a.y - b.y)};
  x2 = cmul_mi(d);
}

}  // namespace b2w
