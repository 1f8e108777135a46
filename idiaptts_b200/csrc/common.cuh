// Shared device/host helpers for libb200world (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/b200world.h"

namespace b2w {

constexpr double kPi = 3.1415926535897932384;
constexpr double kMySafeGuardMinimum = 1e-12;
constexpr double kEps = 2.2204460492503131e-16;
constexpr double kDefaultF0 = 500.0;
constexpr double kFrequencyInterval = 3000.0;
constexpr double kUpperLimit = 15000.0;
constexpr double kFloorF0D4C = 47.0;
constexpr double kLog2 = 0.69314718055994529;

// host-side error plumbing -------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define B2W_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      b2w::set_error(__VA_ARGS__);    \
      return -1;                      \
    }                                 \
  } while (0)

// Twiddle table exp(-2*pi*i*k/kTwN), k in [0, kTwN): fp64, lives in global memory, L1/L2 resident.
constexpr int kTwN = 4096;
const double2* twiddle_table(cudaStream_t stream);  // lazily built once per device (device code computes it with sincospi)

// One out-of-line copy of the double-precision sincos: inlined it is ~250 instructions (with its large-argument path) per call
// site, and the FFT kernels are instruction-fetch sensitive (ncu: 14-16 % "no instruction" stalls).
#ifdef B2W_INLINE_SINCOS
static __device__ __forceinline__ double2 sincos_ol(double x) {
#else
static __device__ __noinline__ double2 sincos_ol(double x) {
#endif
  double s, c;
  sincos(x, &s, &c);
  return make_double2(c, s);
}
#define B2W_SINCOS(ARG_, SPTR_, CPTR_) do { const double2 cs_ = b2w::sincos_ol(ARG_); *(CPTR_) = cs_.x; *(SPTR_) = cs_.y; } while (0)

// device helpers --------------------------------------------------------------------------------------------------
__device__ __forceinline__ int mround_pos(double x) {  // WORLD matlab_round
  return x > 0 ? (int)(x + 0.5) : (int)(x - 0.5);
}

template <int NT>
__device__ __forceinline__ double block_sum(double v, double* scratch /* >= NT/32 doubles */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (NT == 32) return v;
  __syncthreads();  // protect scratch reuse
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double r = 0.0;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) r += scratch[w];
  return r;
}

template <int NT>
__device__ __forceinline__ void block_sum2(double& a, double& b, double* scratch /* >= 2*NT/32 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (NT == 32) return;
  __syncthreads();
  if (lane == 0) {
    scratch[2 * warp] = a;
    scratch[2 * warp + 1] = b;
  }
  __syncthreads();
  double ra = 0.0, rb = 0.0;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) {
    ra += scratch[2 * w];
    rb += scratch[2 * w + 1];
  }
  a = ra;
  b = rb;
}

template <int NT>
__device__ __forceinline__ void block_sum4(double& a, double& b, double& c, double& d, double* scratch /* >= 4*NT/32 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
    d += __shfl_xor_sync(0xffffffffu, d, o);
  }
  if (NT == 32) return;
  __syncthreads();
  if (lane == 0) {
    scratch[4 * warp] = a;
    scratch[4 * warp + 1] = b;
    scratch[4 * warp + 2] = c;
    scratch[4 * warp + 3] = d;
  }
  __syncthreads();
  double ra = 0.0, rb = 0.0, rc = 0.0, rd = 0.0;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) {
    ra += scratch[4 * w];
    rb += scratch[4 * w + 1];
    rc += scratch[4 * w + 2];
    rd += scratch[4 * w + 3];
  }
  a = ra; b = rb; c = rc; d = rd;
}

// In-place inclusive prefix sum of a[0..L) in shared memory (fp64). Each thread scans a contiguous chunk.
template <int NT>
__device__ __forceinline__ void block_scan_inclusive(double* a, int L, double* scratch /* >= NT/32 + 1 */) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk = (L + NT - 1) / NT;
  const int beg = min(tid * chunk, L), end = min(beg + chunk, L);
  double s = 0.0;
  for (int i = beg; i < end; ++i) {
    s += a[i];
    a[i] = s;
  }
  // exclusive scan of the per-thread totals
  double incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  __syncthreads();
  if (lane == 31) scratch[warp] = incl;
  __syncthreads();
  double off = incl - s;
  for (int w = 0; w < warp; ++w) off += scratch[w];
  for (int i = beg; i < end; ++i) a[i] += off;
  __syncthreads();
}

// Waveform sample with pre-emphasis, identical arithmetic to AudioProcessing.get_raw (separate multiply and subtract).
template <int DT>
__device__ __forceinline__ double raw_sample(const void* x, int64_t i) {
  if (DT == B2W_F64) return reinterpret_cast<const double*>(x)[i];
  if (DT == B2W_F32) return (double)reinterpret_cast<const float*>(x)[i];
  return (double)reinterpret_cast<const int16_t*>(x)[i] * (1.0 / 32768.0);  // exact: power of two
}
template <int DT>
__device__ __forceinline__ double emph_sample(const void* x, int64_t base, int idx, double p) {
  double v = raw_sample<DT>(x, base + idx);
  if (p != 0.0 && idx > 0) v = __dsub_rn(v, __dmul_rn(p, raw_sample<DT>(x, base + idx - 1)));
  return v;
}

// WORLD interp1Q evaluated at one point on an array in shared/global memory: y[base] + (y[base+1]-y[base])*frac with
// delta_y[last] = 0.
// inv_dx = 1 / dx is passed instead of dx: one fp64 division per frame instead of one per bin.  (The product may differ from
// the quotient in the last bit; the interpolant is continuous across the integer boundary, so the result moves by ~1e-16.)
__device__ __forceinline__ double interp1q_at(double x0, double inv_dx, const double* y, int y_len, double xi) {
  double pos = (xi - x0) * inv_dx;
  int base = (int)pos;
  double frac = pos - base;
  double y0 = y[base];
  double dy = (base + 1 < y_len) ? (y[base + 1] - y0) : 0.0;
  return y0 + dy * frac;
}

__host__ __device__ __forceinline__ int num_aperiodicities(int fs) {
  double v = fs / 2.0 - kFrequencyInterval;
  if (v > kUpperLimit) v = kUpperLimit;
  return (int)(v / kFrequencyInterval);
}

// interp1 of the coarse aperiodicity (dB) at frequency f: knots 0, 3000, ..., 3000*nap, fs/2 with values
// -60, coarse[0..nap), -1e-12 (WORLD GetAperiodicity / interp1 with histc semantics).
__device__ __forceinline__ double coarse_db_at(const double* coarse, int nap, double fs_half, double f) {
  int k = (int)(f / kFrequencyInterval) + 1;  // first knot strictly greater than f, knots spaced 3000
  if (k > nap + 1) k = nap + 1;
  // knot k-1 <= f < knot k ; the last interval ends at fs/2
  double xl = (k - 1) * kFrequencyInterval;
  double xr = (k == nap + 1) ? fs_half : k * kFrequencyInterval;
  double yl = (k - 1 == 0) ? -60.0 : coarse[k - 2];
  double yr = (k == nap + 1) ? -kMySafeGuardMinimum : coarse[k - 1];
  double s = (f - xl) / (xr - xl);
  return yl + s * (yr - yl);
}

}  // namespace b2w
