// WORLD common.cpp DCCorrection / LinearSmoothing for one CTA on a half spectrum held in shared memory.
#pragma once
#include "common.cuh"

namespace b2w {

// P[0 .. upper-1) += P evaluated at (f0 - axis[i]); S is scratch (>= upper doubles).  N = fft size of the spectrum.
template <int NT, int N>
__device__ __forceinline__ void dc_correction(double* P, double* S, double f0, double fs, int* status) {
  constexpr int K = N / 2 + 1;
  const int tid = threadIdx.x;
  int upper = 2 + (int)(f0 * N / fs);
  if (upper + 1 > K) {  // f0 above ~fs/2: outside WORLD's domain, keep memory safe and flag it
    if (tid == 0) atomicOr(status, B2W_STATUS_F0_TOO_HIGH);
    upper = K - 1;
  }
  const double inv_dx = -(double)N / fs;
  for (int i = tid; i < upper - 1; i += NT) S[i] = interp1q_at(f0, inv_dx, P, upper + 1, (double)i * fs / N);
  __syncthreads();
  for (int i = tid; i < upper - 1; i += NT) P[i] += S[i];
  __syncthreads();
}

// Rectangular smoothing of `width` Hz via a mirrored cumulative sum (block scan) and two linear interpolations per bin.
// in: K doubles (not modified); S: scratch of K + 2*BMAX + 2 doubles; emit(k, value) consumes bin k (may write `in`'s
// storage only if it does not alias S).  All threads must call; ends with the CTA synchronised on S reads only if the
// caller syncs afterwards.
template <int NT, int N, int BMAX, typename Emit>
__device__ __forceinline__ void linear_smoothing(const double* in, double* S, double width, double fs, double* red,
                                                 int* status, Emit emit) {
  constexpr int H = N / 2;
  constexpr int K = H + 1;
  const int tid = threadIdx.x;
  int bnd = (int)(width * N / fs) + 1;
  if (bnd > BMAX) {
    if (tid == 0) atomicOr(status, B2W_STATUS_F0_TOO_HIGH);
    bnd = BMAX;
  }
  const int L = H + 2 * bnd + 1;
  for (int i = tid; i < L; i += NT) {
    double v;
    if (i < bnd) v = in[bnd - i];
    else if (i < H + bnd) v = in[i - bnd];
    else v = in[H - (i - (H + bnd))];
    S[i] = v * fs / N;
  }
  __syncthreads();
  block_scan_inclusive<NT>(S, L, red);
  const double origin_f = -(bnd - 0.5) * fs / N;
  const double inv_dfi = (double)N / fs;
  const double inv_width = 1.0 / width;
  for (int k = tid; k < K; k += NT) {
    const double fax = (double)k / N * fs - width / 2.0;
    const double low = interp1q_at(origin_f, inv_dfi, S, L, fax);
    const double high = interp1q_at(origin_f, inv_dfi, S, L, fax + width);
    emit(k, (high - low) * inv_width);
  }
}

}  // namespace b2w
