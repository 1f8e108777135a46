// D4C band aperiodicity with the LoveTrain voicing stage -- the fast path of the fused extraction: single-precision FFTs on
// packed f32x2 arithmetic, double precision only where cancellation demands it (the cumulative sums of the three smoothings)
// and for the final ratio.
//
// Replaces pyworld.d4c + pyworld.code_aperiodicity as reached from WorldFeatLabelGen.world_extract_features
// (idiaptts/src/data_preparation/world/WorldFeatLabelGen.py:792, :805).  Same algorithm and the same CTA-per-frame structure as
// d4c.cu (WORLD d4c.cpp restated), re-balanced after the round-1 profile (the fp64 kernel spent 80 % of its instructions
// outside the FFTs and spilled):
//   * N / 16 threads per frame, three radix-16 FFT passes (fft32.cuh), first pass fed from the windowing registers;
//   * the centroid pair (w, w * (n + 1)) is packed as (w, w * (n + 1) / wlen): in single precision the two halves of a packed
//     transform must have comparable magnitudes, otherwise the small one inherits the rounding noise of the large one
//     (8e-4 dB without the scaling, 2e-5 dB with it -- numpy emulation);
//   * the waveform segment a frame needs (all four windows lie within 2.25 T0 of the frame centre) is staged once in shared
//     memory, already pre-emphasised;
//   * smoothing: only the half spectrum itself is scanned (in registers, eight bins per thread, one padded shared-memory write
//     per bin); WORLD's mirrored extension is a function of that prefix sum and only matters for the first / last few bins;
//     the bins sit on an integer grid, so the two interpolation fractions are per-frame constants and all addresses move by
//     constant strides;
//   * order statistics: 11-bit key = float exponent + 3 mantissa bits (a handful of candidates per bucket), both bands of a
//     packed transform in one pass (their counts share a histogram word), the band values stay in registers between the
//     histogram, candidate and summation steps;
//   * reductions use two alternating scratch buffers: one barrier per reduction.
// LoveTrain's threshold decision (ap0 <= 0.85) is protected by a guard band: a frame whose single-precision ap0 lies within
// `guard` of the threshold is marked 2 in `voiced` and re-evaluated by the fp64 kernel (d4c.cu, todo mode), so the voiced /
// unvoiced decisions are those of the fp64 path.  Measured ap0 error of this kernel: ~2e-7; guard = 1e-5.
#include "fft32.cuh"

namespace b2w {
namespace {

using f32::ZQ;
using f32::cadd;
using f32::cmul;
using f32::csub;

#ifndef B2W_D4CF_CTAS
#define B2W_D4CF_CTAS 5  // resident CTAs per SM at N = 2048 the register budget is set for
#endif
constexpr int kBMax = 256;  // static bound of the smoothing half-width in bins (f0 * N / fs + 1)

template <int N>
struct Smem {
  static constexpr int NT = N / 16;
  static constexpr int H = N / 2;
  static constexpr int K = H + 1;
  static constexpr int KP = (K + 3) & ~3;
  static constexpr int fft_bytes = f32::zq_size(N) * 8;
  static constexpr int s_bytes = (K + 2 * (K >> 3) + 4) * 8;                    // smoothing scan (padded doubles), aliases the FFT buffer
  static constexpr int hist_off = ((ZQ(H) + 1) * 8 + 15) & ~15;                 // selection: band powers stay in z[0 .. ZQ(H)]
  static constexpr int hist_ints = 2048 + 32;
  static constexpr int sel_bytes = hist_off + hist_ints * 4;
  static constexpr int z_raw = fft_bytes > s_bytes ? (fft_bytes > sel_bytes ? fft_bytes : sel_bytes)
                                                   : (s_bytes > sel_bytes ? s_bytes : sel_bytes);
  static constexpr int z_bytes = (z_raw + 15) & ~15;
  static constexpr int a_off = z_bytes;
  static constexpr int b_off = a_off + KP * 4;
  static constexpr int nut_max = 3 * N / 4 + 4;
  static constexpr int nut_off = b_off + KP * 4;
  static constexpr int seg_max = N + N / 8 + 8;                                 // staged waveform segment (floats)
  static constexpr int seg_off = nut_off + nut_max * 4;
  static constexpr int tw_off = seg_off + seg_max * 4;
  static constexpr int red_off = tw_off + f32::kTwEntries * 8;
  static constexpr int red_doubles = 2 * 4 * (NT / 32 > 0 ? NT / 32 : 1);       // two buffers of up to 4 values per warp
  static constexpr int total_bytes = red_off + red_doubles * 8;
};

enum { kHann = 0, kBlackman = 1 };

// ---- reductions: shuffle tree, then one exchange through shared memory; two alternating buffers = one barrier per call --------
template <int NT, int NV, typename T>
__device__ __forceinline__ void block_sum_n(T (&v)[NV], double* red, int& phase) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  }
  if (NT == 32) return;
  constexpr int NW = NT / 32;
  T* r = reinterpret_cast<T*>(red + phase * (4 * NW));
  phase ^= 1;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) r[warp * NV + i] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    T s = r[i];
#pragma unroll
    for (int w = 1; w < NW; ++w) s += r[w * NV + i];
    v[i] = s;
  }
}

// ---- windows -----------------------------------------------------------------------------------------------------------------
// cos(theta (i - half)) for i = tid, tid + NT, ... by rotation: this thread's start phasor and the stride phasor
struct WinRot {
  float c0, s0, cd, sd;
  int half;
};
template <int NT>
__device__ __forceinline__ WinRot make_winrot(double f0w, double ratio, double fs) {
  WinRot r;
  r.half = mround_pos(ratio * fs / f0w / 2.0);
  const double th = (2.0 / ratio / fs) * f0w;  // angle per sample in units of pi
  sincospif((float)(th * (double)((int)threadIdx.x - r.half)), &r.s0, &r.c0);
  sincospif((float)(th * (double)NT), &r.sd, &r.cd);
  return r;
}
__device__ __forceinline__ float winval(int type, float c) {
  return (type == kHann) ? fmaf(0.5f, c, 0.5f) : fmaf(0.08f, fmaf(2.0f * c, c, -1.0f), fmaf(0.5f, c, 0.42f));
}

template <int DT>
__device__ __forceinline__ float sample_f(const void* x, int64_t base, int idx, double p, float pf) {
  if (DT == B2W_I16) {  // value / 32768 is exact in single precision; one rounding in the pre-emphasis
    const int16_t* xs = reinterpret_cast<const int16_t*>(x) + base;
    float v = (float)xs[idx];
    if (pf != 0.0f && idx > 0) v = fmaf(-pf, (float)xs[idx - 1], v);
    return v * (1.0f / 32768.0f);
  }
  return (float)emph_sample<DT>(x, base, idx, p);
}

// WORLD d4c.cpp GetWindowedWaveform, in registers: w[q] = seg[off + i] * win(i) for i = tid + q * NT < wlen, 0 beyond (the
// zero padding of the FFT input); seg = the staged (clamped, pre-emphasised) waveform, off = origin - half - segment start.
// Partial sums for the DC removal in sw / swin.
template <int NT, int N>
__device__ __forceinline__ void window_raw(float (&w)[16], float& sw, float& swin, const float* seg, int off, const WinRot& rot,
                                           int type) {
  const int wlen = min(2 * rot.half + 1, N);
  const float* sp = seg + off + threadIdx.x;
  const int rem = wlen - (int)threadIdx.x;  // i < wlen  <=>  q * NT < rem
  float c = rot.c0, s = rot.s0;
#pragma unroll
  for (int q = 0; q < 16; ++q) w[q] = 0.0f;
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    if (q * NT >= wlen) break;  // CTA-uniform
    if (q * NT < rem) {
      const float win = winval(type, c);
      w[q] = sp[q * NT] * win;
      sw += w[q];
      swin += win;
    }
    const float cn = c * rot.cd - s * rot.sd;
    s = fmaf(s, rot.cd, c * rot.sd);
    c = cn;
  }
}
// second half of GetWindowedWaveform: w[i] -= win(i) * (sum w / sum win)
template <int NT, int N>
__device__ __forceinline__ void window_dc(float (&w)[16], const WinRot& rot, int type, float coef) {
  const int wlen = min(2 * rot.half + 1, N);
  const int rem = wlen - (int)threadIdx.x;
  float c = rot.c0, s = rot.s0;
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    if (q * NT >= wlen) break;
    if (q * NT < rem) w[q] = fmaf(-winval(type, c), coef, w[q]);
    const float cn = c * rot.cd - s * rot.sd;
    s = fmaf(s, rot.cd, c * rot.sd);
    c = cn;
  }
}

// ---- WORLD common.cpp DCCorrection on a float half spectrum -------------------------------------------------------------------
template <int NT, int N>
__device__ __noinline__ void dc_correction_f(float* P, double f0, double fs, int* status) {
  constexpr int K = N / 2 + 1;
  const int tid = threadIdx.x;
  int upper = 2 + (int)(f0 * N / fs);
  if (upper + 1 > K || upper - 1 > 2 * NT) {  // f0 far above WORLD's domain: keep memory safe and flag it
    if (tid == 0) atomicOr(status, B2W_STATUS_F0_TOO_HIGH);
    upper = min(K - 1, 2 * NT + 1);
  }
  const double inv_dx = -(double)N / fs;
  float add[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int i = tid + e * NT;
    add[e] = 0.0f;
    if (i < upper - 1) {
      const double pos = ((double)i * fs / N - f0) * inv_dx;
      const int base = (int)pos;
      const float frac = (float)(pos - base);
      const float y0 = P[base];
      const float dy = (base + 1 < upper + 1) ? (P[base + 1] - y0) : 0.0f;
      add[e] = fmaf(dy, frac, y0);
    }
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int i = tid + e * NT;
    if (i < upper - 1) P[i] += add[e];
  }
  __syncthreads();
}

// ---- WORLD common.cpp LinearSmoothing ------------------------------------------------------------------------------------------
// Rectangular smoothing of `width` Hz.  WORLD builds the cumulative sum S of the spectrum extended by mirroring (bnd bins on
// either side) and takes (interp(S, f + width / 2) - interp(S, f - width / 2)) / width.  With C[j] = sum_{m <= j} in[m] over the
// half spectrum alone, m = i - bnd:
//     S[i] = C[bnd] - C[-m - 1]                            m < 0          (left mirror)
//          = (C[bnd] - C[0]) + C[m]                        0 <= m < H
//          = (C[bnd] - C[0]) + C[H-1] + C[H] - C[2H-m-1]   m >= H         (right mirror)
// so only C is scanned (fp64: differences of a cumulative sum over a spectrum spanning many decades cancel catastrophically in
// single precision), in registers, eight bins per thread; the constant of the middle case cancels in the difference, and the
// mirror cases are evaluated only by the few bins near 0 and N / 2 that reach them.  Bin k needs S at k + d_lo and k + d_hi
// with per-frame constants, so the two interpolation fractions are constants too.  C is stored padded, CP(j) = j + 2 (j >> 3):
// a thread's eight values go out as four conflict-free 16-byte stores, and CP(j + NT) = CP(j) + NT + NT / 4.
// in: K floats, 16-byte aligned (only read before the first barrier, so out may alias it); C: K + K / 4 + 4 doubles of scratch.
// out[k] = smoothed(in)[k], or sub[k] - smoothed(in)[k] when sub != nullptr.  r: NT / 32 doubles of reduction scratch.
// Ends WITHOUT a barrier after the output loop.
// Out of line (like the FFT tail and the selection): with several CTAs per SM at different phases the kernel is
// instruction-fetch sensitive, one copy of each big block keeps the footprint small.
__device__ __forceinline__ int CP(int j) { return j + ((j >> 3) << 1); }

template <int NT, int N>
__device__ __noinline__ void smooth_f(const float* in, double* C, double width, double fs, double* r, int* status,
                                      float* out, const float* sub) {
  constexpr int H = N / 2;
  constexpr int K = H + 1;
  static_assert(H == 8 * NT, "eight bins per thread");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double wbins = width * N / fs;
  int bnd = (int)wbins + 1;
  if (bnd > kBMax) {
    if (tid == 0) atomicOr(status, B2W_STATUS_F0_TOO_HIGH);
    bnd = kBMax;
  }
  double loc[8];
  {
    const float4 a = reinterpret_cast<const float4*>(in)[2 * tid], b = reinterpret_cast<const float4*>(in)[2 * tid + 1];
    loc[0] = (double)a.x;
    loc[1] = loc[0] + (double)a.y;
    loc[2] = loc[1] + (double)a.z;
    loc[3] = loc[2] + (double)a.w;
    loc[4] = loc[3] + (double)b.x;
    loc[5] = loc[4] + (double)b.y;
    loc[6] = loc[5] + (double)b.z;
    loc[7] = loc[6] + (double)b.w;
  }
  const float nyq = in[H];
  double incl = loc[7];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) r[warp] = incl;
  __syncthreads();  // also: every thread has read its part of `in`
  double off = incl - loc[7];
  for (int w = 0; w < warp; ++w) off += r[w];
  {
    double2* dst = reinterpret_cast<double2*>(C + 10 * tid);  // CP(8 tid) = 10 tid
#pragma unroll
    for (int e = 0; e < 4; ++e) dst[e] = make_double2(off + loc[2 * e], off + loc[2 * e + 1]);
    if (tid == NT - 1) C[CP(H)] = off + loc[7] + (double)nyq;
  }
  __syncthreads();
  // WORLD: low / high = interp1Q(origin, fs / N, S, k fs / N -+ width / 2) with origin = -(bnd - 0.5) fs / N
  const double d_lo = (double)bnd - 0.5 - 0.5 * wbins;
  const double d_hi = d_lo + wbins;
  const int b_lo = (int)d_lo, b_hi = (int)d_hi;
  const double f_lo = d_lo - b_lo, f_hi = d_hi - b_hi;
  const double scale = fs / N / width;
  // the mirrored extension, relative to the constant of the middle case
  const double c_0 = C[0], c_top = C[CP(H - 1)] + C[CP(H)];
  auto Sx = [&](int m) -> double {
    if (m < 0) return c_0 - C[CP(-m - 1)];   // = (C[bnd] - C[-m-1]) - (C[bnd] - C[0])
    if (m < H) return C[CP(m)];
    return c_top - C[CP(2 * H - m - 1)];
  };
  const int m_lo = tid + b_lo - bnd, m_hi = tid + b_hi - bnd;       // bin k = tid + j NT reads C at m + j NT (and + 1)
  const double* p_lo0 = C + CP(m_lo);                               // CP() with an arithmetic shift: also valid strides for m < 0
  const double* p_lo1 = C + CP(m_lo + 1);
  const double* p_hi0 = C + CP(m_hi);
  const double* p_hi1 = C + CP(m_hi + 1);
  constexpr int ST = NT + NT / 4;
  // bins whose four samples lie inside the half spectrum: k in [k_lo, k_hi]
  const int k_lo = bnd - b_lo, k_hi = H - 2 - (b_hi - bnd);
#pragma unroll
  for (int j = 0; j <= 8; ++j) {
    if (j == 8 && tid != 0) break;
    const int k = tid + j * NT;
    if (k >= k_lo && k <= k_hi) {
      const double l0 = p_lo0[j * ST], l1 = p_lo1[j * ST], h0 = p_hi0[j * ST], h1 = p_hi1[j * ST];
      const float v = (float)((fma(h1 - h0, f_hi, h0) - fma(l1 - l0, f_lo, l0)) * scale);
      out[k] = sub ? sub[k] - v : v;
    }
  }
  // the few bins at either end that reach into the mirrored extension (one rolled copy of the general form)
  const int n_lo = min(k_lo, K), n_edge = n_lo + max(0, K - 1 - max(k_hi, n_lo - 1));
#pragma unroll 1
  for (int e = tid; e < n_edge; e += NT) {
    const int k = e < n_lo ? e : max(k_hi, n_lo - 1) + 1 + (e - n_lo);
    const int ml = k + b_lo - bnd, mh = k + b_hi - bnd;
    const double l0 = Sx(ml), l1 = Sx(ml + 1), h0 = Sx(mh), h1 = Sx(mh + 1);
    const float v = (float)((fma(h1 - h0, f_hi, h0) - fma(l1 - l0, f_lo, l0)) * scale);
    out[k] = sub ? sub[k] - v : v;
  }
}

// ---- order statistics -----------------------------------------------------------------------------------------------------------
// For the two band power spectra P_b(i) = zp[ZQ(i)].{x, y}, i < K (values >= 0): the sum of the (K - r_excl) smallest values and
// the sum of all of them, without sorting -- histogram of an 11-bit key (float exponent + 3 mantissa bits), locate the bucket of
// the r_excl-th largest value, rank the few candidates of that bucket against each other, sum everything below the threshold.
// Both bands go through ONE pass: their counts share a histogram word (16 bits each, K < 65536, so sums never carry), the
// values stay in registers between the histogram, candidate and summation steps.  Thread 0 writes
// out_db[b] = min(0, 10 log10(small_b / total_b) + bias).  r: 4 NT / 32 doubles of reduction scratch.  Ends with a barrier.
template <int NT, int N>
__device__ __noinline__ void band_ratios(const float2* zp, bool has2, int r_excl, int* hist, float* cand0, float* cand1, double* r,
                                         double bias, double* out_db) {
  constexpr int BPT = 2048 / NT;
  constexpr int NW = NT / 32;
  constexpr int ST = NT + NT / 16;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // meta: [0] bucket0 [1] need0 [2] bucket1 [3] need1 [4] #cand0 [5] #cand1 [6] ties0 [7] ties1 [8] tau0 [9] tau1 [10 ..] warp totals
  int* meta = hist + 2048;
  for (int i = tid; i < (2048 + 32) / 4; i += NT) reinterpret_cast<int4*>(hist)[i] = make_int4(0, 0, 0, 0);
  __syncthreads();
  float2 vals[9];
  {
    const float2* zk = zp + tid + (tid >> 4);
#pragma unroll
    for (int j = 0; j <= 8; ++j) {
      vals[j] = float2{-1.0f, -1.0f};
      if (j < 8 || tid == 0) {
        vals[j] = zk[j * ST];
        atomicAdd(&hist[__float_as_uint(vals[j].x) >> 20], 1);
        if (has2) atomicAdd(&hist[__float_as_uint(vals[j].y) >> 20], 0x10000);
      }
    }
  }
  __syncthreads();
  {
    // thread t owns buckets [BPT t, BPT t + BPT); packed suffix counts from the top over all threads
    int mine = 0;
#pragma unroll
    for (int e = 0; e < BPT / 4; ++e) {
      // 16-byte loads at a 64-byte thread stride: rotate the order per thread pair so a quarter warp covers all banks
      const int ee = (BPT == 16) ? (e ^ ((tid >> 1) & 3)) : e;
      const int4 q = reinterpret_cast<const int4*>(hist)[(BPT / 4) * tid + ee];
      mine += q.x + q.y + q.z + q.w;
    }
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_down_sync(0xffffffffu, incl, o);
      if (lane + o < 32) incl += v;
    }
    if (lane == 0) meta[10 + warp] = incl;
    __syncthreads();
    int above = incl - mine;
    for (int w = warp + 1; w < NW; ++w) above += meta[10 + w];
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int ab = (above >> (16 * b)) & 0xffff, mb = (mine >> (16 * b)) & 0xffff;
      if (ab < r_excl && r_excl <= ab + mb) {  // one thread per band: walk its buckets from the top
        int cum = ab;
#pragma unroll 1
        for (int e = BPT - 1; e >= 0; --e) {
          const int h = (hist[BPT * tid + e] >> (16 * b)) & 0xffff;
          if (cum + h >= r_excl) {
            meta[2 * b] = BPT * tid + e;
            meta[2 * b + 1] = r_excl - cum;  // the need-th largest inside this bucket is the threshold
            break;
          }
          cum += h;
        }
      }
    }
  }
  __syncthreads();
  const int bucket0 = meta[0], need0 = meta[1], bucket1 = meta[2], need1 = meta[3];
#pragma unroll
  for (int j = 0; j <= 8; ++j) {
    if (vals[j].x >= 0.0f) {
      if ((int)(__float_as_uint(vals[j].x) >> 20) == bucket0) cand0[atomicAdd(&meta[4], 1)] = vals[j].x;
      if (has2 && (int)(__float_as_uint(vals[j].y) >> 20) == bucket1) cand1[atomicAdd(&meta[5], 1)] = vals[j].y;
    }
  }
  __syncthreads();
  const int nc0 = meta[4], nc1 = meta[5];
  for (int i = tid; i < nc0 + nc1; i += NT) {
    const bool second = i >= nc0;
    const float* c = second ? cand1 : cand0;
    const int nc = second ? nc1 : nc0, need = second ? need1 : need0;
    const float v = c[second ? i - nc0 : i];
    int g = 0, eq = 0;
    for (int j = 0; j < nc; ++j) {
      const float u = c[j];
      g += (u > v);
      eq += (u == v);
    }
    if (g < need && need <= g + eq) {  // v is the threshold (all tied candidates write the same values)
      meta[8 + second] = __float_as_int(v);
      meta[6 + second] = eq - (need - g);  // copies of the threshold that stay on the "small" side
    }
  }
  __syncthreads();
  const float tau0 = __int_as_float(meta[8]), tau1 = __int_as_float(meta[9]);
  const int ties0 = meta[6], ties1 = meta[7];
  float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};  // small0, total0, small1, total1
#pragma unroll
  for (int j = 0; j <= 8; ++j) {
    if (vals[j].x >= 0.0f) {
      acc[1] += vals[j].x;
      if (vals[j].x < tau0) acc[0] += vals[j].x;
      acc[3] += vals[j].y;
      if (vals[j].y < tau1) acc[2] += vals[j].y;
    }
  }
  double st[4] = {(double)acc[0], (double)acc[1], (double)acc[2], (double)acc[3]};
  int ph = 0;
  block_sum_n<NT, 4>(st, r, ph);
  if (tid == 0) {
    out_db[0] = fmin(0.0, 10.0 * log10((st[0] + (double)ties0 * (double)tau0) / st[1]) + bias);
    if (has2) out_db[1] = fmin(0.0, 10.0 * log10((st[2] + (double)ties1 * (double)tau1) / st[3]) + bias);
  }
  __syncthreads();
}

template <int N>
__device__ __noinline__ void fft_tail(float2* z, const float2* tws) {
  f32::fft32_tail<N, N / 16>(z, tws, threadIdx.x);
}

template <int N, int XDT>
__global__ void __launch_bounds__(N / 16, B2W_D4CF_CTAS * 2048 / N)
d4c_fast_kernel(b2w_batch b, double threshold, double guard, double* __restrict__ coarse_db, uint8_t* __restrict__ voiced,
                const double2* __restrict__ tw, int* __restrict__ status) {
  using SM = Smem<N>;
  constexpr int NT = N / 16;
  constexpr int H = N / 2;
  constexpr int K = H + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* z = reinterpret_cast<float2*>(smem_raw);
  double* S = reinterpret_cast<double*>(smem_raw);
  int* hist = reinterpret_cast<int*>(smem_raw + SM::hist_off);
  float* A = reinterpret_cast<float*>(smem_raw + SM::a_off);
  float* B = reinterpret_cast<float*>(smem_raw + SM::b_off);
  float* nut = reinterpret_cast<float*>(smem_raw + SM::nut_off);
  float* seg = reinterpret_cast<float*>(smem_raw + SM::seg_off);
  float2* tws = reinterpret_cast<float2*>(smem_raw + SM::tw_off);
  double* red = reinterpret_cast<double*>(smem_raw + SM::red_off);
  const int tid = threadIdx.x;
  int phase = 0;
  auto red_next = [&]() {  // the reduction scratch alternates between two buffers (one barrier per reduction)
    double* r = red + phase * (4 * (NT / 32));
    phase ^= 1;
    return r;
  };
  const double fs = (double)b.fs;
  const float pf = (float)b.preemphasis;
  const int nap = num_aperiodicities(b.fs);
  const int wl = (int)(kFrequencyInterval * N / fs) * 2 + 1;  // Nuttall window length
  const int hw = wl / 2;
  const int boundary = mround_pos(N * 8.0 / wl);
  f32::tw_fill<N, NT>(tws, tw, tid);
  for (int i = tid; i < wl; i += NT) {
    const double x = (double)i / (wl - 1.0);
    nut[i] = (float)(0.355768 - 0.487396 * cospi(2.0 * x) + 0.144232 * cospi(4.0 * x) - 0.012604 * cospi(6.0 * x));
  }
  // LoveTrain band edges (cumulative powers at 100, 4000, 7900 Hz); clamped for fs < 15.8 kHz where WORLD reads past its spectrum
  int lt_b0 = (int)ceil(100.0 * N / fs), lt_b1 = (int)ceil(4000.0 * N / fs), lt_b2 = (int)ceil(7900.0 * N / fs);
  lt_b1 = min(lt_b1, H);
  lt_b2 = min(lt_b2, H);
  __syncthreads();

  for (int64_t frame = blockIdx.x; frame < b.num_frames; frame += gridDim.x) {
    const double f0 = b.f0[frame];
    if (f0 == 0.0) {
      if (tid == 0) voiced[frame] = 0;
      continue;
    }
    const int u = b.frame_utt[frame];
    const int64_t s0 = b.utt_sample_offset[u];
    const int xlen = (int)(b.utt_sample_offset[u + 1] - s0);
    const double tpos = b.t[frame];
    const double f0_lt = fmax(f0, 40.0);
    const double f0c = fmax(f0, kFloorF0D4C);
    const WinRot rot4 = make_winrot<NT>(f0c, 4.0, fs);

    // ---- 0: stage the waveform segment all four windows read: [origin(t - T0/4) - half4, origin(t + T0/4) + half4] -------
    // (the LoveTrain window, 3 T0 at max(f0, 40), is always shorter than the 4 T0 window at max(f0, 47))
    const int origin_c = mround_pos(__dadd_rn(__dmul_rn(tpos, fs), 0.001));
    const int origin_m = mround_pos(__dadd_rn(__dmul_rn(tpos - 0.25 / f0c, fs), 0.001));
    const int origin_p = mround_pos(__dadd_rn(__dmul_rn(tpos + 0.25 / f0c, fs), 0.001));
    const int seg_lo = origin_m - rot4.half;
    {
      const int seg_len = min(origin_p + rot4.half - seg_lo + 1, SM::seg_max);
      for (int j = tid; j < seg_len; j += NT) {
        const int idx = max(0, min(xlen - 1, seg_lo + j));
        seg[j] = sample_f<XDT>(b.x, s0, idx, b.preemphasis, pf);
      }
    }
    __syncthreads();

    // ---- 1: packed FFT: re = LoveTrain window (Blackman, 3 T0), im = smoothed-power window (Hann, 4 T0) ------------------
    {
      const WinRot rot_lt = make_winrot<NT>(f0_lt, 3.0, fs);
      float w1[16], w2[16];
      float sums[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      window_raw<NT, N>(w1, sums[0], sums[1], seg, origin_c - rot_lt.half - seg_lo, rot_lt, kBlackman);
      window_raw<NT, N>(w2, sums[2], sums[3], seg, origin_c - rot4.half - seg_lo, rot4, kHann);
      block_sum_n<NT, 4>(sums, red, phase);
      window_dc<NT, N>(w1, rot_lt, kBlackman, sums[0] / sums[1]);
      window_dc<NT, N>(w2, rot4, kHann, sums[2] / sums[3]);
      float2 v[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) v[q] = float2{w1[q], w2[q]};
      f32::fft32_first_pass<N, NT>(z, v, tid);
    }
    fft_tail<N>(z, tws);
    {
      float c[2] = {0.0f, 0.0f};
      f32::for_pair_bins<N, NT>(z, tid, [&](int k, float2 x1, float2 x2) {
        const float p1 = fmaf(x1.x, x1.x, x1.y * x1.y);
        if (k > lt_b0 && k <= lt_b2) {
          c[1] += p1;
          if (k <= lt_b1) c[0] += p1;
        }
        B[k] = fmaf(x2.x, x2.x, x2.y * x2.y);
      });
      block_sum_n<NT, 2>(c, red, phase);  // its barrier also orders the B writes / z reads before what follows
      const double ap0 = (double)c[0] / (double)c[1];
      if (fabs(ap0 - threshold) < guard) {  // too close to call in single precision: the fp64 kernel decides (CTA-uniform)
        if (tid == 0) voiced[frame] = 2;
        continue;
      }
      if (ap0 <= threshold) {  // LoveTrain says unvoiced
        if (tid == 0) voiced[frame] = 0;
        continue;
      }
    }

    // ---- 2: smoothed power spectrum: DC correction + smoothing of width f0 ---------------------------------------------
    dc_correction_f<NT, N>(B, f0c, fs, status);
    smooth_f<NT, N>(B, S, f0c, fs, red_next(), status, B, nullptr);
    __syncthreads();

    // ---- 3: static centroid from two time-shifted Blackman (4 T0) windows ------------------------------------------------
    const int wlen4 = min(2 * rot4.half + 1, N);
    const float cscale_in = 1.0f / (float)wlen4;  // keeps (w, w * (n + 1)) at comparable magnitudes inside one packed FFT
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
      {
        float w[16];
        float sums[2] = {0.0f, 0.0f};
        window_raw<NT, N>(w, sums[0], sums[1], seg, (side == 0 ? 0 : origin_p - origin_m), rot4, kBlackman);
        block_sum_n<NT, 2>(sums, red, phase);
        window_dc<NT, N>(w, rot4, kBlackman, sums[0] / sums[1]);
        float pw[1] = {0.0f};  // the window spans exactly the 2 round(2 fs / f0) + 1 samples WORLD normalises over
#pragma unroll
        for (int q = 0; q < 16; ++q) pw[0] = fmaf(w[q], w[q], pw[0]);
        block_sum_n<NT, 1>(pw, red, phase);
        const float inv = rsqrtf(pw[0]);
        const float ramp0 = (float)(tid + 1) * cscale_in, ramp_step = (float)NT * cscale_in;
        float2 v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float wn = w[q] * inv;
          v[q] = float2{wn, wn * fmaf((float)q, ramp_step, ramp0)};
        }
        f32::fft32_first_pass<N, NT>(z, v, tid);
      }
      fft_tail<N>(z, tws);
      const float back = (float)wlen4;
      f32::for_pair_bins<N, NT>(z, tid, [&](int k, float2 x1, float2 x2) {
        const float cen = fmaf(x2.x, x1.x, x1.y * x2.y) * back;
        A[k] = side == 0 ? cen : A[k] + cen;
      });
      __syncthreads();
    }
    dc_correction_f<NT, N>(A, f0c, fs, status);

    // ---- 4: static group delay, smoothed twice ---------------------------------------------------------------------------
    for (int k = tid; k < K; k += NT) B[k] = A[k] / B[k];
    __syncthreads();
    smooth_f<NT, N>(B, S, f0c / 2.0, fs, red_next(), status, A, nullptr);
    __syncthreads();
    smooth_f<NT, N>(A, S, f0c, fs, red_next(), status, B, A);
    __syncthreads();
    // B now holds the static group delay; A and the staged segment are free (candidate lists of the selection)

    // ---- 5: per-band Nuttall-windowed segment -> power spectrum -> share of the smallest bins ----------------------
#pragma unroll 1
    for (int band = 0; band < nap; band += 2) {
      const int center1 = (int)(kFrequencyInterval * (band + 1) * N / fs);
      const bool has2 = band + 1 < nap;
      const int center2 = has2 ? (int)(kFrequencyInterval * (band + 2) * N / fs) : center1;
      {
        const float* g1 = B + center1 - hw + tid;
        const float* g2 = B + center2 - hw + tid;
        const float* nw = nut + tid;
        const int rem = wl - tid;
        float2 v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = float2{0.0f, 0.0f};
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          if (q * NT >= wl) break;
          if (q * NT < rem) {
            const float w = nw[q * NT];
            v[q] = float2{g1[q * NT] * w, has2 ? g2[q * NT] * w : 0.0f};
          }
        }
        f32::fft32_first_pass<N, NT>(z, v, tid);
      }
      fft_tail<N>(z, tws);
      // bins k and N - k are read by this thread only, so the two power spectra replace z[k] in place
      {
        float2* zk = z + tid + (tid >> 4);
        f32::for_pair_bins<N, NT>(z, tid, [&](int k, float2 x1, float2 x2) {
          zk[(k - tid) / 16 * 17] = float2{fmaf(x1.x, x1.x, x1.y * x1.y), fmaf(x2.x, x2.x, x2.y * x2.y)};
        });
      }
      __syncthreads();
      band_ratios<NT, N>(z, has2, boundary + 1, hist, A, seg, red_next(), (f0c - 100.0) / 50.0, coarse_db + frame * nap + band);
    }
    if (tid == 0) voiced[frame] = 1;
  }
}

}  // namespace

template <int N>
static int launch_d4c_fast(const b2w_batch* b, double threshold, double guard, double* coarse_db, uint8_t* voiced, int* status,
                           cudaStream_t st) {
  const double2* tw = twiddle_table(st);
  if (!tw) return check_launch("twiddle table");
  const int smem = Smem<N>::total_bytes;
  const int64_t cap = (int64_t)148 * B2W_D4CF_CTAS * 8;  // a few waves: dynamic balance of the uneven frame costs, cheap CTA set-up
  const int grid = (int)(b->num_frames < cap ? b->num_frames : cap);
#define B2W_D4CF_LAUNCH(XDT)                                                                                  \
  do {                                                                                                        \
    cudaFuncSetAttribute(d4c_fast_kernel<N, XDT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);         \
    d4c_fast_kernel<N, XDT><<<grid, N / 16, smem, st>>>(*b, threshold, guard, coarse_db, voiced, tw, status); \
  } while (0)
  if (b->x_dtype == B2W_F64) B2W_D4CF_LAUNCH(B2W_F64);
  else if (b->x_dtype == B2W_F32) B2W_D4CF_LAUNCH(B2W_F32);
  else B2W_D4CF_LAUNCH(B2W_I16);
#undef B2W_D4CF_LAUNCH
  return check_launch("d4c_fast_kernel");
}

// Called by b2w_d4c_coarse (d4c.cu): the single-precision pass; frames it leaves marked 2 are re-evaluated there in fp64.
int d4c_fast_pass(const b2w_batch* b, int n4, double threshold, double guard, double* coarse_db, uint8_t* voiced, int* status,
                  cudaStream_t st) {
  switch (n4) {
    case 1024: return launch_d4c_fast<1024>(b, threshold, guard, coarse_db, voiced, status, st);
    case 2048: return launch_d4c_fast<2048>(b, threshold, guard, coarse_db, voiced, status, st);
    case 4096: return launch_d4c_fast<4096>(b, threshold, guard, coarse_db, voiced, status, st);
    default: set_error("b2w_d4c_coarse: unsupported D4C fft size %d at fs=%d", n4, b->fs); return -1;
  }
}

}  // namespace b2w
