// CheapTrick spectral envelope, the fast path of the fused extraction: one WARP per analysis frame, mixed precision.
//
// Replaces pyworld.cheaptrick as reached from WorldFeatLabelGen.world_extract_features
// (idiaptts/src/data_preparation/world/WorldFeatLabelGen.py:792, pyworld.wav2world) when the envelope goes to the mel-cepstrum
// kernel as a float32 plane (fft size 1024: fs <= 32 kHz).  Same algorithm as cheaptrick_kernel (cheaptrick.cu, fp64 throughout,
// what compat.pyworld.cheaptrick keeps using); what changes is the mapping and the precision of each stage:
//   * a frame is owned by one warp (w512: radix 16 x 16 x 2 transforms, sixteen points per lane, first pass fed from
//     registers): no CTA barrier, reductions are shuffles, eight frames per CTA share the constant tables;
//   * transform 1 (windowed waveform -> power spectrum) stays in DOUBLE precision: its weak bins sit up to 80 dB under the strong
//     ones and the rounding noise of a single-precision transform (relative to the strong bins) is what a numpy emulation showed
//     to cost 9e-5 relative in the envelope on the reference's utterances -- too close to the 1e-4 tolerance;
//   * DC correction in single precision, the smoothing's cumulative sum in double (as everywhere: differences of a cumulative sum
//     over many decades), scanned in registers, the mirrored extension evaluated from the prefix sum of the half spectrum alone;
//   * transforms 2 and 3 (log spectrum -> cepstrum -> liftered envelope) in single precision on the mean-removed log spectrum.
// Envelope error against the fp64 oracle: <= 5e-5 relative (emulation over all 11 579 fixture frames; tolerance 1e-4).
#include "wfft512.cuh"

namespace b2w {
namespace {

using namespace w512;

constexpr int kWarps = 8;      // frames per CTA
constexpr int kBMaxCt = 128;   // static bound of the smoothing half-width in bins (2/3 f0 N / fs + 1)

__device__ __forceinline__ int CP(int j) { return j + ((j >> 3) << 1); }  // padded index of the prefix sum (see d4c_fast.cu)

struct Smem {
  static constexpr int z_bytes = ((f32::zq_size(kM) * 16) + 15) & ~15;           // FFT buffer (double2 for transform 1), prefix sum
  static constexpr int p_bytes = ((kK + 3) & ~3) * 4;                           // power spectrum / log spectrum / cepstrum (floats)
  static constexpr int warp_bytes = z_bytes + p_bytes;
  static constexpr int tw256d_off = kWarps * warp_bytes;                        // double2 exp(-2 pi i k / 256), k < 256
  static constexpr int tw512d_off = tw256d_off + 256 * 16;                      // double2 exp(-2 pi i k / 512), k < 256
  static constexpr int twnd_off = tw512d_off + 256 * 16;                        // double2 exp(-2 pi i k / 1024), k < 512
  static constexpr int tw16_off = twnd_off + 512 * 16;                          // float2 tables of the single-precision transforms
  static constexpr int tw512_off = tw16_off + 16 * 8;
  static constexpr int twn_off = tw512_off + 256 * 8;
  static constexpr int total_bytes = twn_off + 512 * 8;
  static_assert((kK + 2 * (kK >> 3) + 4) * 8 <= z_bytes, "the padded prefix sum must fit in the FFT buffer");
};

template <int DT>
__device__ __forceinline__ float sample_f(const void* x, int64_t base, int idx, int xlen, double p, float pf) {
  idx = max(0, min(xlen - 1, idx));
  if (DT == B2W_I16) {  // value / 32768 is exact in single precision; one rounding in the pre-emphasis
    const int16_t* xs = reinterpret_cast<const int16_t*>(x) + base;
    float v = (float)xs[idx];
    if (pf != 0.0f && idx > 0) v = fmaf(-pf, (float)xs[idx - 1], v);
    return v * (1.0f / 32768.0f);
  }
  return (float)emph_sample<DT>(x, base, idx, p);
}

template <int XDT>
__global__ void __launch_bounds__(32 * kWarps, 2)
cheaptrick_fast_kernel(b2w_batch b, double q1, float* __restrict__ sp, int64_t sp_stride, const double2* __restrict__ tw,
                       int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double2* tw256d = reinterpret_cast<double2*>(smem_raw + Smem::tw256d_off);
  double2* tw512d = reinterpret_cast<double2*>(smem_raw + Smem::tw512d_off);
  double2* twnd = reinterpret_cast<double2*>(smem_raw + Smem::twnd_off);
  float2* tw16 = reinterpret_cast<float2*>(smem_raw + Smem::tw16_off);
  float2* tw512 = reinterpret_cast<float2*>(smem_raw + Smem::tw512_off);
  float2* twn = reinterpret_cast<float2*>(smem_raw + Smem::twn_off);
  for (int e = threadIdx.x; e < 512; e += 32 * kWarps) {
    const double2 wn = __ldg(&tw[e * (kTwN / 1024)]);
    twnd[e] = wn;
    twn[e] = float2{(float)wn.x, (float)wn.y};
    if (e < 256) {
      const double2 a = __ldg(&tw[e * (kTwN / 256)]), c = __ldg(&tw[e * (kTwN / 512)]);
      tw256d[e] = a;
      tw512d[e] = c;
      tw512[e] = float2{(float)c.x, (float)c.y};
      if (e < 16) tw16[e] = float2{(float)a.x, (float)a.y};
    }
  }
  __syncthreads();  // the only CTA barrier: tables

  unsigned char* mine = smem_raw + warp * Smem::warp_bytes;
  double2* zd = reinterpret_cast<double2*>(mine);
  float2* zf = reinterpret_cast<float2*>(mine);
  double* C = reinterpret_cast<double*>(mine);
  float* P = reinterpret_cast<float*>(mine + Smem::z_bytes);
  float* seg = reinterpret_cast<float*>(mine);  // the frame's waveform segment, staged in the (still idle) FFT buffer
  const double fs = (double)b.fs;
  const float pf = (float)b.preemphasis;
  const double f0_floor = 3.0 * fs / (kN - 3.0);

  for (int64_t frame = (int64_t)blockIdx.x * kWarps + warp; frame < b.num_frames; frame += (int64_t)gridDim.x * kWarps) {
    const int u = b.frame_utt[frame];
    const int64_t s0 = b.utt_sample_offset[u];
    const int xlen = (int)(b.utt_sample_offset[u + 1] - s0);
    double f0 = b.f0[frame];
    if (f0 <= f0_floor) f0 = kDefaultF0;
    const int half = mround_pos(1.5 * fs / f0);
    const int wlen = 2 * half + 1;  // <= 1021 for f0 > f0_floor
    const int origin = mround_pos(__dadd_rn(__dmul_rn(b.t[frame], fs), 0.001));

    // the samples under the window, read once, coalesced (the two window passes below would otherwise wait on strided global
    // loads: the first version of this kernel spent most of its time there)
    for (int i = lane; i < wlen; i += 32) seg[i] = sample_f<XDT>(b.x, s0, origin + i - half, xlen, b.preemphasis, pf);
    __syncwarp();

    // ---- A: Hann window by rotation recurrence (samples 2 n and 2 n + 1, n = lane + 32 q), the three sums -------------------
    const double theta = kPi * f0 / 1.5 / fs;  // angle per sample
    double cd, sd, c1, s1;                     // stride of 64 samples, and of one sample
    sincos(theta * 64.0, &sd, &cd);
    sincos(theta, &s1, &c1);
    double c_start, s_start;
    sincos(theta * (double)(2 * lane - half), &s_start, &c_start);
    double sum_w2 = 0.0, sum_xw = 0.0, sum_w = 0.0;
    {
      double c = c_start, s = s_start;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        if (64 * q >= wlen) break;  // warp-uniform
        const int i = 2 * (lane + 32 * q);
        if (i < wlen) {
          const double w0 = 0.5 * c + 0.5;
          const double x0 = (double)seg[i];
          sum_w2 += w0 * w0;
          sum_xw += x0 * w0;
          sum_w += w0;
          if (i + 1 < wlen) {
            const double w1 = 0.5 * (c * c1 - s * s1) + 0.5;
            const double x1 = (double)seg[i + 1];
            sum_w2 += w1 * w1;
            sum_xw += x1 * w1;
            sum_w += w1;
          }
        }
        const double cn = c * cd - s * sd;
        s = s * cd + c * sd;
        c = cn;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sum_w2 += __shfl_xor_sync(0xffffffffu, sum_w2, o);
      sum_xw += __shfl_xor_sync(0xffffffffu, sum_xw, o);
      sum_w += __shfl_xor_sync(0xffffffffu, sum_w, o);
    }
    // normalised window wn = w / sqrt(sum w^2); windowed waveform x wn - wn (sum x wn / sum wn)
    const double inv_avg = 1.0 / sqrt(sum_w2);
    const double coef = sum_xw / sum_w;  // = (sum x wn) / (sum wn): the normalisation cancels
    {
      double2 v[16];
      double c = c_start, s = s_start;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        v[q] = double2{0.0, 0.0};
        if (64 * q < wlen) {
          const int i = 2 * (lane + 32 * q);
          if (i < wlen) {
            const double w0 = (0.5 * c + 0.5) * inv_avg;
            v[q].x = ((double)seg[i] - coef) * w0;
            if (i + 1 < wlen) {
              const double w1 = (0.5 * (c * c1 - s * s1) + 0.5) * inv_avg;
              v[q].y = ((double)seg[i + 1] - coef) * w1;
            }
          }
          const double cn = c * cd - s * sd;
          s = s * cd + c * sd;
          c = cn;
        }
      }
      __syncwarp();  // every lane has read its samples: the transform may overwrite the staging area
      wfft512_f64(zd, v, tw256d, tw512d, lane);
    }
    for_real_bins_f64(zd, twnd, lane, [&](int k, double2 X) { P[k] = (float)(X.x * X.x + X.y * X.y); });
    __syncwarp();

    // ---- D: DC correction (fold the spectrum below f0 back around f0) ----------------------------------------------------------
    {
      int upper = 2 + (int)(f0 * kN / fs);
      if (upper + 1 > kK || upper - 1 > 96) {  // f0 far above WORLD's domain: keep memory safe and flag it
        if (lane == 0) atomicOr(status, B2W_STATUS_F0_TOO_HIGH);
        upper = min(kK - 1, 97);
      }
      const double inv_dx = -(double)kN / fs;
      float add[3];
#pragma unroll
      for (int e = 0; e < 3; ++e) {
        const int i = lane + 32 * e;
        add[e] = 0.0f;
        if (i < upper - 1) {
          const double pos = ((double)i * fs / kN - f0) * inv_dx;
          const int base = (int)pos;
          const float frac = (float)(pos - base);
          const float y0 = P[base];
          const float dy = (base + 1 < upper + 1) ? (P[base + 1] - y0) : 0.0f;
          add[e] = fmaf(dy, frac, y0);
        }
      }
      __syncwarp();
#pragma unroll
      for (int e = 0; e < 3; ++e) {
        const int i = lane + 32 * e;
        if (i < upper - 1) P[i] += add[e];
      }
      __syncwarp();
    }

    // ---- E: rectangular smoothing of width 2 f0 / 3 (WORLD LinearSmoothing), log, mean removal ----------------------------------
    float lp[17];
    float mean;
    {
      const double width = f0 * 2.0 / 3.0;
      const double wbins = width * kN / fs;
      int bnd = (int)wbins + 1;
      if (bnd > kBMaxCt) {
        if (lane == 0) atomicOr(status, B2W_STATUS_F0_TOO_HIGH);
        bnd = kBMaxCt;
      }
      // prefix sum C[j] = sum_{m <= j} P[m] over the half spectrum: sixteen bins per lane, scanned in registers
      double loc[16];
      {
        const float4* src = reinterpret_cast<const float4*>(P) + 4 * lane;
        double run = 0.0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 a = src[e];
          loc[4 * e] = run += (double)a.x;
          loc[4 * e + 1] = run += (double)a.y;
          loc[4 * e + 2] = run += (double)a.z;
          loc[4 * e + 3] = run += (double)a.w;
        }
      }
      const float nyq = P[kM];
      double incl = loc[15];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const double off = incl - loc[15];
      __syncwarp();  // every lane has read P (and the FFT buffer is free): C may overwrite the buffer
      {
        double2* dst = reinterpret_cast<double2*>(C + 20 * lane);  // CP(16 lane) = 20 lane
#pragma unroll
        for (int e = 0; e < 8; ++e) dst[e + (e >> 2)] = make_double2(off + loc[2 * e], off + loc[2 * e + 1]);  // 2-double pad after 8
        if (lane == 31) C[CP(kM)] = off + loc[15] + (double)nyq;
      }
      __syncwarp();
      const double d_lo = (double)bnd - 0.5 - 0.5 * wbins;
      const double d_hi = d_lo + wbins;
      const int b_lo = (int)d_lo, b_hi = (int)d_hi;
      const double f_lo = d_lo - b_lo, f_hi = d_hi - b_hi;
      const double scale = fs / kN / width;
      const double c_0 = C[0], c_top = C[CP(kM - 1)] + C[CP(kM)];
      auto Sx = [&](int m) -> double {  // cumulative sum of the mirrored spectrum, relative to the constant of the middle case
        if (m < 0) return c_0 - C[CP(-m - 1)];
        if (m < kM) return C[CP(m)];
        return c_top - C[CP(2 * kM - m - 1)];
      };
      const int k_lo = bnd - b_lo, k_hi = kM - 2 - (b_hi - bnd);
      const int m_lo = lane + b_lo - bnd, m_hi = lane + b_hi - bnd;
      const double* p_lo0 = C + CP(m_lo);
      const double* p_lo1 = C + CP(m_lo + 1);
      const double* p_hi0 = C + CP(m_hi);
      const double* p_hi1 = C + CP(m_hi + 1);
      float acc = 0.0f;
#pragma unroll
      for (int j = 0; j <= 16; ++j) {
        lp[j] = 0.0f;
        if (j == 16 && lane != 0) break;
        const int k = lane + 32 * j;
        double l0, l1, h0, h1;
        if (k >= k_lo && k <= k_hi) {
          l0 = p_lo0[40 * j]; l1 = p_lo1[40 * j]; h0 = p_hi0[40 * j]; h1 = p_hi1[40 * j];  // CP(m + 32) = CP(m) + 40
        } else {
          const int ml = k + b_lo - bnd, mh = k + b_hi - bnd;
          l0 = Sx(ml); l1 = Sx(ml + 1); h0 = Sx(mh); h1 = Sx(mh + 1);
        }
        const float v = (float)((fma(h1 - h0, f_hi, h0) - fma(l1 - l0, f_lo, l0)) * scale);
        lp[j] = logf(v + (float)kEps);
        acc += (k == 0 || k == kM) ? lp[j] : 2.0f * lp[j];  // the mirrored sequence of length N
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      mean = acc * (1.0f / kN);
      __syncwarp();  // C has been read: P (float) is rewritten next, the buffer after that
#pragma unroll
      for (int j = 0; j <= 16; ++j) {
        if (j == 16 && lane != 0) break;
        P[lane + 32 * j] = lp[j] - mean;
      }
      __syncwarp();
    }

    // ---- F: cepstrum of the (mean-removed) log spectrum, liftering, envelope ------------------------------------------------------
    {
      float2 v[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {  // mirrored sequence x[m] = L[m] (m <= 512), L[1024 - m] beyond, packed in pairs
        const int m = 2 * (lane + 32 * q);
        v[q] = (m < kM) ? float2{P[m], P[m + 1]} : (m == kM ? float2{P[kM], P[kM - 1]} : float2{P[kN - m], P[kN - m - 1]});
      }
      __syncwarp();
      wfft512(zf, v, tw16, tw512, lane);
      const float a1 = (float)(f0 / fs);  // pi f0 k / fs in units of pi
      const float q1f = (float)q1;
      for_real_bins(zf, twn, lane, [&](int k, float2 X) {
        float sn, cs;
        sincospif(a1 * (float)k, &sn, &cs);
        const float arg = (float)kPi * a1 * (float)k;
        const float smooth = (k == 0) ? 1.0f : sn / arg;
        const float comp = (1.0f - 2.0f * q1f) + 2.0f * q1f * (1.0f - 2.0f * sn * sn);  // cos(2x) = 1 - 2 sin^2 x
        P[k] = X.x * smooth * comp * (1.0f / kN);
      });
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int m = 2 * (lane + 32 * q);
        v[q] = (m < kM) ? float2{P[m], P[m + 1]} : (m == kM ? float2{P[kM], P[kM - 1]} : float2{P[kN - m], P[kN - m - 1]});
      }
      __syncwarp();
      wfft512(zf, v, tw16, tw512, lane);
      float* out = sp + frame * sp_stride;
      for_real_bins(zf, twn, lane, [&](int k, float2 X) { out[k] = expf(X.x + mean); });
      __syncwarp();
    }
  }
}

}  // namespace

int cheaptrick_fast_launch(const b2w_batch* b, double q1, float* sp, int64_t sp_stride, int* status, cudaStream_t st) {
  const double2* tw = twiddle_table(st);
  if (!tw) return check_launch("twiddle table");
  const int smem = Smem::total_bytes;
  const int64_t want = (b->num_frames + kWarps - 1) / kWarps;
  const int64_t cap = (int64_t)148 * 2 * 16;  // a few waves of resident CTAs; the rest is the grid-stride loop
  const int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
#define B2W_CTF_LAUNCH(XDT)                                                                                       \
  do {                                                                                                            \
    cudaFuncSetAttribute(cheaptrick_fast_kernel<XDT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);          \
    cheaptrick_fast_kernel<XDT><<<grid, 32 * kWarps, smem, st>>>(*b, q1, sp, sp_stride, tw, status);               \
  } while (0)
  if (b->x_dtype == B2W_F64) B2W_CTF_LAUNCH(B2W_F64);
  else if (b->x_dtype == B2W_F32) B2W_CTF_LAUNCH(B2W_F32);
  else B2W_CTF_LAUNCH(B2W_I16);
#undef B2W_CTF_LAUNCH
  return check_launch("cheaptrick_fast_kernel");
}

}  // namespace b2w
