// F0 estimation on the device: pyworld.dio (speed = 1) + pyworld.stonemask, i.e. the F0 half of
// pyworld.wav2world (reference call sites: world/WorldFeatLabelGen.py:792, world/LF0LabelGen.py:263-264).
//
// DIO.  WORLD filters the whole utterance in the frequency domain (one FFT of the padded signal, one inverse FFT per band).
// Both filters are short FIRs (low cut: delta minus a unit-sum Hann of 2*round(fs/50)+1 taps; band b: a Nuttall window of
// 4*round(fs/boundary_b/2) taps), and WORLD sizes its FFT so that the circular convolution never wraps: the result equals
// the LINEAR convolution, which is what the kernels below compute directly in fp64 -- tile-parallel, register-tiled (8
// outputs per thread from a sliding window, one tap + one sample loaded per 8 FMAs), no utterance-sized FFT, no
// utterance-sized intermediate besides the low-cut signal itself.  The band kernel detects the four kinds of zero crossings
// on the tile it has just filtered and appends the refined crossing positions to ordered per-(utterance, band, kind) lists
// (block scan), so the band-filtered signals never reach HBM.
//
// StoneMask needs the spectrum of a Blackman-windowed segment and of its derivative-windowed twin at 2 + 6 harmonic bins
// only: a direct DFT at those bins (rotation recurrences) replaces the two FFTs.
#include "common.cuh"

namespace b2w {

constexpr int kDioTile = 2048;      // outputs per FIR tile (256 threads x 8)
constexpr int kDioThreads = 256;
constexpr double kDioCutOff = 50.0;
constexpr double kDioMaximumValue = 100000.0;
constexpr int kDioMaxBands = 16;

__host__ __device__ __forceinline__ int mround_hd(double x) { return x > 0 ? (int)(x + 0.5) : (int)(x - 0.5); }
__host__ __device__ __forceinline__ int pad9(int i) { return i + (i >> 3); }

struct DioGeom {
  int fs, nb;
  int lowcut_h;      // (N - 1) / 2 of the low-cut Hann, N = 2 * round(fs / 50) + 1
  int pad;           // padding of the low-cut signal on both sides: 2 * half_average_length of band 0
  int half[kDioMaxBands];
  double boundary[kDioMaxBands];
  double f0_floor, f0_ceil, allowed_range, frame_period;
};

static int dio_geom(int fs, double f0_floor, double f0_ceil, double channels, double frame_period, double allowed_range,
                    DioGeom* g) {
  if (!(f0_floor > 0 && f0_ceil > f0_floor && channels > 0)) return -1;
  int nb = 1 + (int)(log(f0_ceil / f0_floor) / kLog2 * channels);
  if (nb < 1 || nb > kDioMaxBands) return -1;
  g->fs = fs;
  g->nb = nb;
  for (int i = 0; i < nb; ++i) {
    g->boundary[i] = f0_floor * pow(2.0, (i + 1) / channels);
    g->half[i] = mround_hd(fs / g->boundary[i] / 2.0);
    if (g->half[i] < 1) return -1;
  }
  g->lowcut_h = mround_hd(fs / kDioCutOff);
  g->pad = 2 * g->half[0];
  g->f0_floor = f0_floor;
  g->f0_ceil = f0_ceil;
  g->allowed_range = allowed_range;
  g->frame_period = frame_period;
  return 0;
}

// workspace layout (bytes offsets, all 16-byte aligned) ----------------------------------------------------------
struct DioWs {
  int64_t mean, ylc, ev, counts, cands, scores, s1, neg, pos, total;
  int64_t ev_plane;  // doubles per (band, kind) plane
};
static DioWs dio_ws(int64_t S, int64_t U, int64_t F, const DioGeom& g) {
  DioWs w;
  auto al = [](int64_t v) { return (v + 15) & ~(int64_t)15; };
  int64_t o = 0;
  w.mean = o; o = al(o + 8 * U);
  w.ylc = o; o = al(o + 8 * (S + U * (1 + 2 * (int64_t)g.pad)));
  w.ev_plane = S / 2 + 3 * U + 4;
  w.ev = o; o = al(o + 8 * w.ev_plane * 4 * g.nb);
  w.counts = o; o = al(o + 4 * U * g.nb * 4);
  w.cands = o; o = al(o + 8 * F * g.nb);
  w.scores = o; o = al(o + 8 * F * g.nb);
  w.s1 = o; o = al(o + 8 * F);
  w.neg = o; o = al(o + 4 * F);
  w.pos = o; o = al(o + 4 * F);
  w.total = o;
  return w;
}

// ---------------------------------------------------------------------------------------------------------------
// 1. mean of y[0 .. L] (y[L] = 0: WORLD's y_length = x_length + 1 at decimation ratio 1)
// ---------------------------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(256) dio_mean_kernel(const void* x, const int64_t* soff, double p, double* mean) {
  __shared__ double scratch[8];
  const int u = blockIdx.x;
  const int64_t base = soff[u];
  const int L = (int)(soff[u + 1] - base);
  double s = 0.0;
  for (int i = threadIdx.x; i < L; i += 256) s += emph_sample<DT>(x, base, i, p);
  s = block_sum<256>(s, scratch);
  if (threadIdx.x == 0) mean[u] = s / (double)(L + 1);
}

// ---------------------------------------------------------------------------------------------------------------
// FIR tile: out[o] = sum_j wr[j] * s[o + j], o = 8 t + c, s in the 9/8-padded staging buffer, wr zero-padded to a multiple of 8
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fir8(const double* __restrict__ s, const double* __restrict__ wr, int ntaps8, double acc[8]) {
  const int o = 8 * threadIdx.x;
  double r[16];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    acc[c] = 0.0;
    r[c] = s[pad9(o + c)];
  }
  for (int j0 = 0; j0 < ntaps8; j0 += 8) {
#pragma unroll
    for (int c = 0; c < 8; ++c) r[8 + c] = s[pad9(o + j0 + 8 + c)];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const double wv = wr[j0 + jj];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = fma(wv, r[jj + c], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) r[c] = r[8 + c];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 2. low cut: ylc[n] = y[n] - sum_j hann[j + h] y[n - j], n in [-pad, L + pad]; tile-parallel over the padded output
// ---------------------------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(kDioThreads) dio_lowcut_kernel(const void* x, const int64_t* soff, int U, double p,
                                                                 const double* mean, int h, int pad, double* ylc_all) {
  extern __shared__ double sm[];
  const int ntaps = 2 * h + 1, ntaps8 = (ntaps + 7) & ~7;
  double* wr = sm;                       // [ntaps8]
  double* s = sm + ntaps8;               // staging, pad9(kDioTile + ntaps8 + 8)
  const int tid = threadIdx.x;
  // Hann taps, unit sum (DesignLowCutFilter); symmetric, so reversed == forward
  {
    double part = 0.0;
    for (int i = tid; i < ntaps8; i += kDioThreads) {
      double v = 0.0;
      if (i < ntaps) v = 0.5 - 0.5 * cospi((double)(i + 1) * 2.0 / (double)(ntaps + 1));
      wr[i] = v;
      part += v;
    }
    __shared__ double scratch[8];
    const double tot = block_sum<kDioThreads>(part, scratch);
    for (int i = tid; i < ntaps8; i += kDioThreads) wr[i] = wr[i] / tot;
    __syncthreads();
  }
  // this CTA's slice [g0, g1) of the concatenated padded outputs
  const int64_t stride_u = 1 + 2 * (int64_t)pad;
  const int64_t g0 = (int64_t)blockIdx.x * kDioTile;
  const int64_t total = soff[U] + U * stride_u;
  const int64_t g1 = min(g0 + kDioTile, total);
  if (g0 >= total) return;
  // first utterance whose padded range contains g0: start_u = soff[u] + u * stride_u
  int lo = 0, hi = U - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (soff[mid] + mid * stride_u <= g0) lo = mid; else hi = mid - 1;
  }
  for (int u = lo; u < U; ++u) {
    const int64_t start = soff[u] + u * stride_u;
    if (start >= g1) break;
    const int64_t base = soff[u];
    const int L = (int)(soff[u + 1] - base);
    const int64_t end = start + L + stride_u;
    const int q0 = (int)(max(g0, start) - start), q1 = (int)(min(g1, end) - start);  // output slots [q0, q1) of utterance u
    const double mu = mean[u];
    // staging index i <-> y position n = (q0 - pad) - h + i   (out[o] = y[n0 + o] - sum_j wr[j] y[n0 + o - h + j])
    const int n_lo = q0 - pad - h;
    const int n_stage = kDioTile + ntaps8 + 8;
    __syncthreads();
    for (int i = tid; i < n_stage; i += kDioThreads) {
      const int n = n_lo + i;
      double v = 0.0;
      if (n >= 0 && n < L) v = emph_sample<DT>(x, base, n, p) - mu;
      else if (n == L) v = -mu;
      s[pad9(i)] = v;
    }
    __syncthreads();
    double acc[8];
    fir8(s, wr, ntaps8, acc);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int o = 8 * tid + c;
      if (q0 + o < q1) ylc_all[start + q0 + o] = s[pad9(o + h)] - acc[c];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 3. band filter + four zero-crossing event lists; one CTA per (utterance, band), tiles walked in order
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kDioThreads) dio_band_kernel(const int64_t* soff, int U, DioGeom g, const double* ylc_all,
                                                               double* ev, int64_t ev_plane, int32_t* counts) {
  extern __shared__ double sm[];
  // band-major CTA order: band 0 has the longest filter (8 x the work of the last band), so the heavy CTAs start first and
  // the tail of the grid is made of short ones
  const int band = blockIdx.x / U, u = blockIdx.x % U;
  const int H = g.half[band], ntaps = 4 * H, ntaps8 = (ntaps + 7) & ~7;
  double* wr = sm;                                   // [ntaps8] reversed Nuttall
  double* s = wr + ntaps8;                           // staging pad9(kDioTile + ntaps8 + 8)
  double* sig = s + pad9(kDioTile + ntaps8 + 8) + 1; // pad9(kDioTile + 8)
  __shared__ unsigned long long warp_tot[kDioThreads / 32];
  __shared__ unsigned long long tile_tot;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < ntaps8; i += kDioThreads) {
    double v = 0.0;
    const int k = ntaps - 1 - i;  // reversed
    if (k >= 0) {
      const double tmp = (double)k / (ntaps - 1.0);
      v = 0.355768 - 0.487396 * cospi(2.0 * tmp) + 0.144232 * cospi(4.0 * tmp) - 0.012604 * cospi(6.0 * tmp);
    }
    wr[i] = v;
  }
  const int64_t stride_u = 1 + 2 * (int64_t)g.pad;
  const int64_t base = soff[u];
  const int L = (int)(soff[u + 1] - base);
  const int ylen = L + 1;
  const double* ylc = ylc_all + base + u * stride_u + g.pad;   // ylc[n], n in [-pad, L + pad]
  const int64_t start_u = (base >> 1) + 3 * (int64_t)u;
  double* list[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) list[k] = ev + (int64_t)(band * 4 + k) * ev_plane + start_u;
  int run[4] = {0, 0, 0, 0};
  const int step = kDioTile - 2;
  for (int o0 = 0; o0 < ylen - 1; o0 += step) {
    // sig[o0 + o] = sum_k w[k] ylc[o0 + o + 2H - k] = sum_j wr[j] ylc[o0 + o + 2H - (ntaps - 1) + j]
    const int n_lo = o0 + 2 * H - (ntaps - 1);
    const int n_stage = kDioTile + ntaps8 + 8;
    __syncthreads();
    for (int i = tid; i < n_stage; i += kDioThreads) {
      const int n = n_lo + i;
      s[pad9(i)] = (n >= -g.pad && n <= L + g.pad) ? ylc[n] : 0.0;
    }
    __syncthreads();
    double acc[8];
    fir8(s, wr, ntaps8, acc);
#pragma unroll
    for (int c = 0; c < 8; ++c) sig[pad9(8 * tid + c)] = acc[c];
    __syncthreads();
    // events at positions i = o0 + o, o in [0, step): kinds 0 negative-going, 1 positive-going, 2 peak, 3 dip
    double v[10];
#pragma unroll
    for (int c = 0; c < 10; ++c) {
      const int o = 8 * tid + c;
      v[c] = (o < kDioTile) ? sig[pad9(o)] : 0.0;
    }
    unsigned flags[4] = {0, 0, 0, 0};
    unsigned long long cnt = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int o = 8 * tid + c;
      const int i = o0 + o;
      const bool in1 = (o < step) && (i + 1 < ylen);
      const bool in2 = (o < step) && (i + 2 < ylen);
      const double a = v[c], b = v[c + 1], d0 = v[c] - v[c + 1], d1 = v[c + 1] - v[c + 2];
      const bool e0 = in1 && (0.0 < a) && (b <= 0.0);
      const bool e1 = in1 && (0.0 < -a) && (-b <= 0.0);
      const bool e2 = in2 && (0.0 < d0) && (d1 <= 0.0);
      const bool e3 = in2 && (0.0 < -d0) && (-d1 <= 0.0);
      flags[0] |= (unsigned)e0 << c; flags[1] |= (unsigned)e1 << c; flags[2] |= (unsigned)e2 << c; flags[3] |= (unsigned)e3 << c;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) cnt |= (unsigned long long)__popc(flags[k]) << (16 * k);
    // block exclusive scan of the packed counts
    unsigned long long incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    unsigned long long excl = incl - cnt;
    for (int w = 0; w < warp; ++w) excl += warp_tot[w];
    if (tid == kDioThreads - 1) tile_tot = excl + cnt;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int pos = run[k] + (int)((excl >> (16 * k)) & 0xffff);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (flags[k] >> c & 1) {  // events are rare: the refinement (an fp64 division) is only done for them
          const double a = (k < 2) ? v[c] : v[c] - v[c + 1];
          const double b = (k < 2) ? v[c + 1] : v[c + 1] - v[c + 2];
          list[k][pos++] = (double)(o0 + 8 * tid + c + 1) - a / (b - a);
        }
      }
    }
    __syncthreads();
    const unsigned long long tt = tile_tot;
#pragma unroll
    for (int k = 0; k < 4; ++k) run[k] += (int)((tt >> (16 * k)) & 0xffff);
  }
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) counts[((int64_t)u * g.nb + band) * 4 + k] = run[k];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 4. candidates and scores: one thread per (band, frame)
// ---------------------------------------------------------------------------------------------------------------
__global__ void dio_candidates_kernel(const int64_t* soff, const int32_t* frame_utt, const double* t, int64_t F, DioGeom g,
                                      const double* ev, int64_t ev_plane, const int32_t* counts, double* cands, double* scores) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= F * g.nb) return;
  const int band = (int)(idx / F);
  const int64_t f = idx - (int64_t)band * F;
  const int u = frame_utt[f];
  const double tf = t[f], fs = (double)g.fs;
  const int64_t start_u = (soff[u] >> 1) + 3 * (int64_t)u;
  double val[4];
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ne = counts[((int64_t)u * g.nb + band) * 4 + k];
    const int ni = ne >= 2 ? ne - 1 : 0;  // intervals
    if (ni - 2 <= 0) ok = false;
  }
  double cand = 0.0, score = kDioMaximumValue;
  if (ok) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double* e = ev + (int64_t)(band * 4 + k) * ev_plane + start_u;
      const int ni = counts[((int64_t)u * g.nb + band) * 4 + k] - 1;
      // first interval location strictly greater than tf (histc), clipped to [1, ni - 1]
      int lo = 0, hi = ni;  // searchsorted right over loc[j] = (e[j] + e[j+1]) / 2 / fs
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const double loc = (e[mid] + e[mid + 1]) / 2.0 / fs;
        if (loc <= tf) lo = mid + 1; else hi = mid;
      }
      int kk = lo < 1 ? 1 : (lo > ni - 1 ? ni - 1 : lo);
      const double e0 = e[kk - 1], e1 = e[kk], e2 = e[kk + 1];
      const double x0 = (e0 + e1) / 2.0 / fs, x1 = (e1 + e2) / 2.0 / fs;
      const double y0 = fs / (e1 - e0), y1 = fs / (e2 - e1);
      const double sfrac = (tf - x0) / (x1 - x0);
      val[k] = y0 + sfrac * (y1 - y0);
    }
    cand = (val[0] + val[1] + val[2] + val[3]) / 4.0;
    const double d0 = val[0] - cand, d1 = val[1] - cand, d2 = val[2] - cand, d3 = val[3] - cand;
    score = sqrt((d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3) / 3.0);
    const double bf = g.boundary[band];
    if (cand > bf || cand < bf / 2.0 || cand > g.f0_ceil || cand < g.f0_floor) {
      cand = 0.0;
      score = kDioMaximumValue;
    }
  }
  cands[idx] = cand;
  scores[idx] = score / (cand + kMySafeGuardMinimum);
}

// ---------------------------------------------------------------------------------------------------------------
// 5. best contour + FixF0Contour (steps 1-4); one CTA per utterance, the extension steps are sequential (thread 0)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double dio_best(const double* cands, const double* scores, int64_t F, int nb, int64_t f) {
  double best = cands[f], tmp = scores[f];
  for (int j = 1; j < nb; ++j) {
    const double sc = scores[(int64_t)j * F + f];
    if (tmp > sc) {
      tmp = sc;
      best = cands[(int64_t)j * F + f];
    }
  }
  return best;
}

__device__ __forceinline__ double dio_select_best(double current_f0, double past_f0, const double* cands, int64_t F, int nb,
                                                  int64_t target, double allowed_range) {
  const double ref = (current_f0 * 3.0 - past_f0) / 2.0;
  double best = cands[target];
  double err = fabs(ref - best);
  for (int i = 1; i < nb; ++i) {
    const double c = cands[(int64_t)i * F + target];
    const double e = fabs(ref - c);
    if (e < err) {
      err = e;
      best = c;
    }
  }
  if (fabs(1.0 - best / ref) > allowed_range) return 0.0;
  return best;
}

__global__ void __launch_bounds__(128) dio_fix_kernel(const int64_t* foff, int64_t F, DioGeom g, int step2_sections,
                                                      const double* cands, const double* scores, double* s1, int32_t* neg,
                                                      int32_t* pos, double* f0_out) {
  const int u = blockIdx.x, tid = threadIdx.x;
  const int64_t f0 = foff[u];
  const int T = (int)(foff[u + 1] - f0);
  const int vrm = (int)(0.5 + 1000.0 / g.frame_period / g.f0_floor) * 2 + 1;
  if (T <= vrm) {
    for (int i = tid; i < T; i += 128) f0_out[f0 + i] = 0.0;
    return;
  }
  // step 1: reject jumps of the best contour (ends zeroed)
  for (int i = tid; i < T; i += 128) {
    double r = 0.0;
    if (i >= vrm) {
      const double bi = (i < T - vrm) ? dio_best(cands, scores, F, g.nb, f0 + i) : 0.0;
      const double bp = (i - 1 >= vrm && i - 1 < T - vrm) ? dio_best(cands, scores, F, g.nb, f0 + i - 1) : 0.0;
      r = fabs((bi - bp) / (kMySafeGuardMinimum + bi)) < g.allowed_range ? bi : 0.0;
    }
    s1[f0 + i] = r;
  }
  __syncthreads();
  double* s2 = f0_out + f0;
  if (!step2_sections) {
    // step 2 (the variant the reference's fixtures were produced with): erosion by voice_range_minimum frames
    const int center = (vrm - 1) / 2;
    for (int i = tid; i < T; i += 128) {
      double r = s1[f0 + i];
      if (i >= center && i < T - center) {
        for (int j = -center; j <= center; ++j)
          if (s1[f0 + i + j] == 0.0) r = 0.0;
      }
      s2[i] = r;
    }
    __syncthreads();
  } else {
    // later WORLD variant: drop voiced sections shorter than voice_range_minimum (vuv[0] = vuv[T-1] = 0 in GetBoundaryList)
    for (int i = tid; i < T; i += 128) s2[i] = s1[f0 + i];
    __syncthreads();
    if (tid == 0) {
      int nbnd = 0, b_start = 0;
      for (int i = 0; i < T - 1; ++i) {
        const int v0 = (i == 0) ? 0 : (s1[f0 + i] > 0.0);
        const int v1 = (i + 1 == T - 1) ? 0 : (s1[f0 + i + 1] > 0.0);
        if (v1 - v0 != 0) {
          const int bpos = i + (nbnd & 1);
          if (nbnd & 1) {
            if (bpos - b_start < vrm)
              for (int j = b_start; j <= bpos; ++j) s2[j] = 0.0;
          } else {
            b_start = bpos;
          }
          ++nbnd;
        }
      }
    }
    __syncthreads();
  }
  if (tid != 0) return;
  // voiced section boundaries of step 2
  int npos = 0, nneg = 0;
  int32_t* ng = neg + f0;
  int32_t* ps = pos + f0;
  {
    double prev = s2[0];
    for (int i = 1; i < T; ++i) {
      const double cur = s2[i];
      if (cur == 0.0 && prev != 0.0) ng[nneg++] = i - 1;
      else if (prev == 0.0 && cur != 0.0) ps[npos++] = i;
      prev = cur;
    }
  }
  // step 3: extend forward
  for (int i = 0; i < nneg; ++i) {
    const int limit = (i == nneg - 1) ? T - 1 : ng[i + 1];
    for (int j = ng[i]; j < limit; ++j) {
      const double r = dio_select_best(s2[j], s2[j - 1], cands, F, g.nb, f0 + j + 1, g.allowed_range);
      s2[j + 1] = r;
      if (r == 0.0) break;
    }
  }
  // step 4: extend backward
  for (int i = npos - 1; i >= 0; --i) {
    const int limit = (i == 0) ? 1 : ps[i - 1];
    for (int j = ps[i]; j > limit; --j) {
      const double r = dio_select_best(s2[j], s2[j + 1], cands, F, g.nb, f0 + j - 1, g.allowed_range);
      s2[j - 1] = r;
      if (r == 0.0) break;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// StoneMask: one CTA (128 threads) per frame
// ---------------------------------------------------------------------------------------------------------------
constexpr int kSmThreads = 128;
constexpr double kFloorF0StoneMask = 40.0;

// sums over n of v[n] * exp(-2 pi i idx n / fft) for main and diff waveforms; thread handles n = tid + 128 m
__device__ __forceinline__ void sm_dft_bin(const double* mw, const double* dw, int n_len, int idx, int fft_size, double out[4]) {
  const int tid = threadIdx.x;
  const int id = ((idx % fft_size) + fft_size) % fft_size;
  double s0, c0, ss, cs;
  sincospi(-2.0 * (double)(((long long)id * tid) % fft_size) / (double)fft_size, &s0, &c0);
  sincospi(-2.0 * (double)(((long long)id * kSmThreads) % fft_size) / (double)fft_size, &ss, &cs);
  double mr = 0.0, mi = 0.0, dr = 0.0, di = 0.0;
  for (int n = tid; n < n_len; n += kSmThreads) {
    const double a = mw[n], b = dw[n];
    mr = fma(a, c0, mr); mi = fma(a, s0, mi);
    dr = fma(b, c0, dr); di = fma(b, s0, di);
    const double cn = c0 * cs - s0 * ss;
    s0 = s0 * cs + c0 * ss;
    c0 = cn;
  }
  out[0] = mr; out[1] = mi; out[2] = dr; out[3] = di;
}

// FixF0 of stonemask.cpp over nh harmonics of f0_in; all threads return the same value
__device__ double sm_fix_f0(const double* mw, const double* dw, int n_len, int fft_size, double fs, double f0_in, int nh,
                            double* scratch) {
  double num = 0.0, den = 0.0;
  for (int i = 0; i < nh; ++i) {
    const int idx = mround_pos(f0_in * fft_size / fs * (i + 1));
    double o[4];
    sm_dft_bin(mw, dw, n_len, idx, fft_size, o);
    block_sum4<kSmThreads>(o[0], o[1], o[2], o[3], scratch);
    const double power = o[0] * o[0] + o[1] * o[1];
    const double numer = o[0] * o[3] - o[1] * o[2];
    const double inst = (power == 0.0) ? 0.0 : (double)idx * fs / fft_size + numer / power * fs / 2.0 / kPi;
    const double amp = sqrt(power);
    num += amp * inst;
    den += amp * (i + 1.0);
  }
  return num / (den + kMySafeGuardMinimum);
}

template <int DT>
__global__ void __launch_bounds__(kSmThreads) stonemask_kernel(b2w_batch b, int n_max, double* refined) {
  extern __shared__ double sm[];
  __shared__ double scratch[4 * kSmThreads / 32];
  double* mw = sm;             // main waveform  [n_max]
  double* dw = sm + n_max;     // diff waveform  [n_max]
  double* win = sm + 2 * n_max;  // main window  [n_max]
  const int64_t f = blockIdx.x;
  const double f0 = b.f0[f];
  const double fs = (double)b.fs;
  if (f0 <= kFloorF0StoneMask || f0 > fs / 12.0) {
    if (threadIdx.x == 0) refined[f] = 0.0;
    return;
  }
  const int u = b.frame_utt[f];
  const int64_t base = b.utt_sample_offset[u];
  const int L = (int)(b.utt_sample_offset[u + 1] - base);
  const double pos = b.t[f];
  const int half = (int)(1.5 * fs / f0 + 1.0);
  const int n_len = 2 * half + 1;
  if (n_len > n_max) {  // cannot happen for f0 > 40 Hz (n_max is sized from it)
    if (threadIdx.x == 0) refined[f] = f0;
    return;
  }
  const double wlen = (2.0 * half + 1.0) / fs;
  const int fft_size = 1 << (2 + (int)(log(half * 2.0 + 1.0) / kLog2));
  const int tid = threadIdx.x;
  for (int i = tid; i < n_len; i += kSmThreads) {
    const double bt = (double)(i - half) / fs;
    const int raw = mround_pos(__dmul_rn(__dadd_rn(pos, bt), fs));
    const double tmp = (raw - 1.0) / fs - pos;
    const double a = tmp / wlen;
    win[i] = 0.42 + 0.5 * cospi(2.0 * a) + 0.08 * cospi(4.0 * a);
    int si = raw - 1;
    si = si < 0 ? 0 : (si > L - 1 ? L - 1 : si);
    mw[i] = emph_sample<DT>(b.x, base, si, b.preemphasis);
  }
  __syncthreads();
  for (int i = tid; i < n_len; i += kSmThreads) {
    double d;
    if (i == 0) d = -win[1] / 2.0;
    else if (i == n_len - 1) d = win[n_len - 2] / 2.0;
    else d = -(win[i + 1] - win[i - 1]) / 2.0;
    const double x = mw[i];
    dw[i] = x * d;
  }
  __syncthreads();
  for (int i = tid; i < n_len; i += kSmThreads) mw[i] *= win[i];
  __syncthreads();
  double mean_f0 = 0.0;
  const double tentative = sm_fix_f0(mw, dw, n_len, fft_size, fs, f0, 2, scratch);
  if (!(tentative <= 0.0 || tentative > f0 * 2)) mean_f0 = sm_fix_f0(mw, dw, n_len, fft_size, fs, tentative, 6, scratch);
  if (fabs(mean_f0 - f0) > f0 * 0.2) mean_f0 = f0;
  if (tid == 0) refined[f] = mean_f0;
}

}  // namespace b2w

// =================================================================================================================
// C ABI
// =================================================================================================================
extern "C" int32_t b2w_dio_num_bands(double f0_floor, double f0_ceil, double channels_in_octave) {
  b2w::DioGeom g;
  if (b2w::dio_geom(16000, f0_floor, f0_ceil, channels_in_octave, 5.0, 0.1, &g)) return -1;
  return g.nb;
}

extern "C" int64_t b2w_dio_workspace_bytes(int64_t num_samples, int32_t num_utts, int64_t num_frames, int32_t fs, double f0_floor,
                                           double f0_ceil, double channels_in_octave) {
  b2w::DioGeom g;
  if (b2w::dio_geom(fs, f0_floor, f0_ceil, channels_in_octave, 5.0, 0.1, &g)) return -1;
  return b2w::dio_ws(num_samples, num_utts, num_frames, g).total;
}

extern "C" int b2w_dio(const b2w_batch* b, int64_t num_samples, const int64_t* utt_frame_offset, double f0_floor, double f0_ceil,
                       double channels_in_octave, double frame_period, double allowed_range, int32_t step2_sections,
                       void* workspace, double* f0_out, void* stream) {
  using namespace b2w;
  B2W_REQUIRE(b && b->x && b->utt_sample_offset && b->frame_utt && b->t && utt_frame_offset && workspace && f0_out,
              "b2w_dio: null argument");
  B2W_REQUIRE(b->x_dtype == B2W_F64 || b->x_dtype == B2W_F32 || b->x_dtype == B2W_I16, "b2w_dio: bad x_dtype %d", b->x_dtype);
  B2W_REQUIRE(b->fs >= 4000 && b->fs <= 96000, "b2w_dio: fs %d out of range", b->fs);
  B2W_REQUIRE(frame_period > 0 && allowed_range > 0, "b2w_dio: bad frame_period / allowed_range");
  DioGeom g;
  B2W_REQUIRE(dio_geom(b->fs, f0_floor, f0_ceil, channels_in_octave, frame_period, allowed_range, &g) == 0,
              "b2w_dio: bad f0_floor %g / f0_ceil %g / channels_in_octave %g", f0_floor, f0_ceil, channels_in_octave);
  const int U = b->num_utts;
  const int64_t F = b->num_frames, S = num_samples;
  if (U <= 0 || F <= 0) return 0;
  B2W_REQUIRE(S >= 0 && S + U * (1 + 2 * (int64_t)g.pad) < ((int64_t)1 << 40), "b2w_dio: bad num_samples");
  const DioWs w = dio_ws(S, U, F, g);
  char* ws = (char*)workspace;
  double* mean = (double*)(ws + w.mean);
  double* ylc = (double*)(ws + w.ylc);
  double* ev = (double*)(ws + w.ev);
  int32_t* counts = (int32_t*)(ws + w.counts);
  double* cands = (double*)(ws + w.cands);
  double* scores = (double*)(ws + w.scores);
  double* s1 = (double*)(ws + w.s1);
  int32_t* neg = (int32_t*)(ws + w.neg);
  int32_t* pos = (int32_t*)(ws + w.pos);
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  // 1. mean
  switch (b->x_dtype) {
    case B2W_F64: dio_mean_kernel<B2W_F64><<<U, 256, 0, st>>>(b->x, b->utt_sample_offset, b->preemphasis, mean); break;
    case B2W_F32: dio_mean_kernel<B2W_F32><<<U, 256, 0, st>>>(b->x, b->utt_sample_offset, b->preemphasis, mean); break;
    default: dio_mean_kernel<B2W_I16><<<U, 256, 0, st>>>(b->x, b->utt_sample_offset, b->preemphasis, mean); break;
  }
  if ((rc = check_launch("dio_mean_kernel"))) return rc;
  // 2. low cut
  {
    const int ntaps8 = (2 * g.lowcut_h + 1 + 7) & ~7;
    const size_t smem = sizeof(double) * (ntaps8 + pad9(kDioTile + ntaps8 + 8) + 1);
    B2W_REQUIRE(smem <= 200 * 1024, "b2w_dio: fs %d too high for the low-cut tile", b->fs);
    const int64_t total = S + U * (1 + 2 * (int64_t)g.pad);
    const unsigned grid = (unsigned)((total + kDioTile - 1) / kDioTile);
#define B2W_LAUNCH_LOWCUT(DT_)                                                                                      \
  do {                                                                                                              \
    cudaFuncSetAttribute(dio_lowcut_kernel<DT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
    dio_lowcut_kernel<DT_><<<grid, kDioThreads, smem, st>>>(b->x, b->utt_sample_offset, U, b->preemphasis, mean,    \
                                                            g.lowcut_h, g.pad, ylc);                                \
  } while (0)
    switch (b->x_dtype) {
      case B2W_F64: B2W_LAUNCH_LOWCUT(B2W_F64); break;
      case B2W_F32: B2W_LAUNCH_LOWCUT(B2W_F32); break;
      default: B2W_LAUNCH_LOWCUT(B2W_I16); break;
    }
#undef B2W_LAUNCH_LOWCUT
    if ((rc = check_launch("dio_lowcut_kernel"))) return rc;
  }
  // 3. bands
  {
    const int ntaps8 = (4 * g.half[0] + 7) & ~7;
    const size_t smem = sizeof(double) * (ntaps8 + pad9(kDioTile + ntaps8 + 8) + 1 + pad9(kDioTile + 8) + 1);
    B2W_REQUIRE(smem <= 200 * 1024, "b2w_dio: fs %d / f0_floor %g too large for the band tile", b->fs, f0_floor);
    cudaFuncSetAttribute(dio_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dio_band_kernel<<<(unsigned)(U * g.nb), kDioThreads, smem, st>>>(b->utt_sample_offset, U, g, ylc, ev, w.ev_plane, counts);
    if ((rc = check_launch("dio_band_kernel"))) return rc;
  }
  // 4. candidates
  {
    const int64_t n = F * g.nb;
    dio_candidates_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(b->utt_sample_offset, b->frame_utt, b->t, F, g, ev,
                                                                       w.ev_plane, counts, cands, scores);
    if ((rc = check_launch("dio_candidates_kernel"))) return rc;
  }
  // 5. contour
  dio_fix_kernel<<<U, 128, 0, st>>>(utt_frame_offset, F, g, step2_sections, cands, scores, s1, neg, pos, f0_out);
  return check_launch("dio_fix_kernel");
}

extern "C" int b2w_stonemask(const b2w_batch* b, double* refined_f0, void* stream) {
  using namespace b2w;
  B2W_REQUIRE(b && b->x && b->utt_sample_offset && b->frame_utt && b->f0 && b->t && refined_f0, "b2w_stonemask: null argument");
  B2W_REQUIRE(b->x_dtype == B2W_F64 || b->x_dtype == B2W_F32 || b->x_dtype == B2W_I16, "b2w_stonemask: bad x_dtype %d", b->x_dtype);
  B2W_REQUIRE(b->fs >= 4000 && b->fs <= 96000, "b2w_stonemask: fs %d out of range", b->fs);
  if (b->num_frames <= 0) return 0;
  const int n_max = 2 * (int)(1.5 * b->fs / kFloorF0StoneMask + 1.0) + 1;
  const size_t smem = sizeof(double) * 3 * (size_t)n_max;
  B2W_REQUIRE(smem <= 200 * 1024, "b2w_stonemask: fs %d too high", b->fs);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)b->num_frames;
#define B2W_LAUNCH_SM(DT_)                                                                                    \
  do {                                                                                                        \
    cudaFuncSetAttribute(stonemask_kernel<DT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
    stonemask_kernel<DT_><<<grid, kSmThreads, smem, st>>>(*b, n_max, refined_f0);                             \
  } while (0)
  switch (b->x_dtype) {
    case B2W_F64: B2W_LAUNCH_SM(B2W_F64); break;
    case B2W_F32: B2W_LAUNCH_SM(B2W_F32); break;
    default: B2W_LAUNCH_SM(B2W_I16); break;
  }
#undef B2W_LAUNCH_SM
  return check_launch("stonemask_kernel");
}
