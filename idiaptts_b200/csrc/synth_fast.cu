// WORLD synthesis, the per-pulse response kernel of the batched fast path: one WARP per pulse, single-precision FFTs on packed
// f32x2 arithmetic, float32 responses.
//
// Replaces the per-pulse body of pyworld.synthesize (WORLD synthesis.cpp GetOneFrameSegment) as reached from
// WorldFeatLabelGen.world_features_to_raw (idiaptts/src/data_preparation/world/WorldFeatLabelGen.py:943) through
// Synthesiser.run_world_synth (idiaptts/src/Synthesiser.py:39-80).  Same arithmetic as render_kernel (synth.cu, fp64, what
// compat.pyworld.synthesize keeps using), re-cast for throughput:
//   * a pulse is owned by one warp: the seven 512-point complex transforms of a pulse (two minimum-phase spectra = four, the
//     noise spectrum, two inverse real transforms) run as radix 16 x 16 x 2 Stockham passes with sixteen points per lane, every
//     first pass fed from registers; there is no CTA barrier anywhere, only __syncwarp, and reductions are shuffles;
//   * the periodic response never touches memory: both inverse transforms leave sample pairs (2n, 2n + 1), n = lane + 32 j, in
//     the same lanes, so the periodic pair waits in registers until the aperiodic one exists;
//   * WORLD's DC remover (Hann-shaped weights) is a 2 KB table per CTA; sqrt(1 - cos^2) of the fractional-delay phasor is
//     taken as |sin|;
//   * everything that DECIDES something stays in double precision and equal to the fp64 kernel: frame indices, interpolation
//     weight, the voiced test ar[0] > 0.999, the aperiodic ratio itself (1 - ar loses all accuracy in single precision).
// Responses are float32 [pulse][N]; overlap_add sums them in double, in pulse order (bit-reproducible).
// Accuracy against the fp64 path: resynthesis SNR ~ 110 dB (tolerance 60 dB).
#include "wfft512.cuh"

namespace b2w {
namespace {

using namespace w512;           // kN = 1024 (the fft size this kernel is built for: fs <= 32 kHz; other sizes take the fp64 kernel),
                                // kM = 512 complex points, kK = 513 bins, wfft512, for_real_bins
constexpr int kWarps = 4;       // pulses per CTA

struct Smem {
  static constexpr int z_bytes = ((f32::zq_size(kM) * 8) + 15) & ~15;   // FFT buffer
  static constexpr int x_bytes = ((kK + 1) * 8 + 15) & ~15;             // spectrum X / log spectrum / cepstrum (aliased)
  static constexpr int l_bytes = ((kK + 3) & ~3) * 4;                   // aperiodic log spectrum, parked during the periodic half
  static constexpr int warp_bytes = z_bytes + x_bytes + l_bytes;
  static constexpr int tw16_off = kWarps * warp_bytes;                  // exp(-2 pi i k / 256), k < 16
  static constexpr int tw512_off = tw16_off + 16 * 8;                   // exp(-2 pi i k / 512), k < 256
  static constexpr int twn_off = tw512_off + 256 * 8;                   // exp(-2 pi i k / 1024), k < 512
  static constexpr int dcr_off = twn_off + 512 * 8;                     // DC remover weights, i < 512
  static constexpr int total_bytes = dcr_off + 512 * 4;
};

// WORLD GetMinimumPhaseSpectrum: log-amplitude L[0 .. 512] (floats, may alias X) -> complex spectrum X[0 .. 512].
// (Out of line, like the transform tail: called twice per pulse.)
__device__ __noinline__ void minimum_phase(const float* L, float2* z, float2* X, const float2* tw16, const float2* tw512,
                                              const float2* twn, int lane) {
  float* c = reinterpret_cast<float*>(X);
  float2 v[16];
  // mirrored log spectrum as a packed real sequence: x[m] = L[m] (m <= 512), L[1024 - m] beyond
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const int m = 2 * (lane + 32 * q);
    v[q] = (m < kM) ? float2{L[m], L[m + 1]} : (m == kM ? float2{L[kM], L[kM - 1]} : float2{L[kN - m], L[kN - m - 1]});
  }
  __syncwarp();  // every lane has read L before the cepstrum overwrites it
  wfft512(z, v, tw16, tw512, lane);
  // cepstrum (real for a symmetric input), folded: c[0], 2 c[1 .. 511], c[512], zeros
  for_real_bins(z, twn, lane, [&](int k, float2 s) { c[k] = (k == 0 || k == kM) ? s.x : 2.0f * s.x; });
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const int m = 2 * (lane + 32 * q);
    v[q] = (m < kM) ? float2{c[m], c[m + 1]} : (m == kM ? float2{c[kM], 0.0f} : float2{0.0f, 0.0f});
  }
  __syncwarp();
  wfft512(z, v, tw16, tw512, lane);
  for_real_bins(z, twn, lane, [&](int k, float2 s) {
    const float mag = __expf(s.x * (1.0f / kN));
    float sn, cs;
    __sincosf(s.y * (1.0f / kN), &sn, &cs);
    X[k] = float2{mag * cs, mag * sn};
  });
  __syncwarp();
}

// Unnormalised inverse real FFT of the Hermitian half spectrum X[0 .. 512]: on return out[j] = (x[2 n], x[2 n + 1]) for
// n = lane + 32 j, j < 16.  (A c2r transform ignores the imaginary parts of the DC and Nyquist bins: WORLD / FFTW semantics.)
__device__ __noinline__ void inverse_real_core(const float2* X, float2* z, const float2* tw16, const float2* tw512, const float2* twn,
                                               int lane) {
  float2 v[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const int k = lane + 32 * q;
    float2 a = X[k];
    float2 bq = X[kM - k];
    if (k == 0) {
      a.y = 0.0f;
      bq.y = 0.0f;
    }
    const float2 b = float2{bq.x, -bq.y};  // conj(X[M - k]) = X[k + M]
    const float2 e = cadd(a, b);
    const float2 d = csub(a, b);
    const float2 w = twn[k];
    const float2 o = cmul(d, float2{w.x, -w.y});  // W^-k
    // Y = e + i o; a forward FFT of conj(Y) followed by a conjugate is the inverse transform
    v[q] = float2{e.x - o.y, -(e.y + o.x)};
  }
  wfft512(z, v, tw16, tw512, lane);
}
// (the transform itself is out of line -- two calls per pulse --, only the sixteen result loads are inline: an array reference
// through a real call would live in local memory)
__device__ __forceinline__ void inverse_real(const float2* X, float2* z, const float2* tw16, const float2* tw512, const float2* twn,
                                             int lane, float2 (&out)[16]) {
  inverse_real_core(X, z, tw16, tw512, twn, lane);
  const float2* zi = z + lane + (lane >> 4);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float2 r = zi[34 * j];
    out[j] = float2{r.x, -r.y};
  }
  __syncwarp();
}

template <typename PT>
__device__ __forceinline__ double plane_at(const void* p, int64_t idx) { return (double)reinterpret_cast<const PT*>(p)[idx]; }

template <typename PT>
__global__ void __launch_bounds__(32 * kWarps, 4)
render_fast_kernel(const void* __restrict__ sp, const void* __restrict__ ap, const int64_t* __restrict__ utt_frame_offset,
                   const int64_t* __restrict__ utt_pulse_offset, const int* __restrict__ num_pulses, int num_utts,
                   const int* __restrict__ pulse_index, const double* __restrict__ pulse_shift,
                   const uint8_t* __restrict__ pulse_vuv, const double* __restrict__ randn_table, int64_t randn_len, int fs_i,
                   double frame_period_ms, float* __restrict__ response, const double2* __restrict__ tw, double dc_rs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2* tw16 = reinterpret_cast<float2*>(smem_raw + Smem::tw16_off);
  float2* tw512 = reinterpret_cast<float2*>(smem_raw + Smem::tw512_off);
  float2* twn = reinterpret_cast<float2*>(smem_raw + Smem::twn_off);
  float* dcr = reinterpret_cast<float*>(smem_raw + Smem::dcr_off);
  for (int e = threadIdx.x; e < 16 + 256 + 512; e += 32 * kWarps) {
    const int idx = e < 16 ? e * (kTwN / 256) : (e < 272 ? (e - 16) * (kTwN / 512) : (e - 272) * (kTwN / 1024));
    const double2 w = __ldg(&tw[idx]);
    (e < 16 ? tw16[e] : (e < 272 ? tw512[e - 16] : twn[e - 272])) = float2{(float)w.x, (float)w.y};
  }
  for (int i = threadIdx.x; i < kM; i += 32 * kWarps) dcr[i] = (float)((0.5 - 0.5 * cospi(2.0 * (i + 1.0) / (1.0 + kN))) / dc_rs);
  __syncthreads();  // the only CTA barrier: tables

  unsigned char* mine = smem_raw + warp * Smem::warp_bytes;
  float2* z = reinterpret_cast<float2*>(mine);
  float2* X = reinterpret_cast<float2*>(mine + Smem::z_bytes);
  float* Lx = reinterpret_cast<float*>(X);  // periodic log spectrum: consumed before X is written
  float* La = reinterpret_cast<float*>(mine + Smem::z_bytes + Smem::x_bytes);
  const double fs = (double)fs_i;
  const double fp = frame_period_ms / 1000.0;
  const int64_t total_rows = utt_pulse_offset[num_utts];

  // one pulse slot (row of the slab layout) per warp, grid-strided
  for (int64_t row = (int64_t)blockIdx.x * kWarps + warp; row < total_rows; row += (int64_t)gridDim.x * kWarps) {
    int lo = 0, hi = num_utts;  // utt_pulse_offset[lo] <= row < utt_pulse_offset[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (utt_pulse_offset[mid] <= row) lo = mid; else hi = mid;
    }
    const int u = lo;
    const int64_t poff = utt_pulse_offset[u];
    const int p = (int)(row - poff);
    const int P = num_pulses[u];
    if (p >= P) continue;  // warp-uniform
    const int64_t f_off = utt_frame_offset[u];
    const int T = (int)(utt_frame_offset[u + 1] - f_off);
    const int n_p = pulse_index[poff + p];
    const int n_next = pulse_index[poff + min(P - 1, p + 1)];
    const int n_first = pulse_index[poff];
    const int noise_size = n_next - n_p;
    const bool cur_vuv = pulse_vuv[poff + p] != 0;
    const double cur_time = (double)n_p / fs;
    const int fl = min(T - 1, (int)floor(cur_time / fp));
    const int ce = min(T - 1, (int)ceil(cur_time / fp));
    const double w = cur_time / fp - fl;
    const int64_t r0 = (f_off + fl) * kK, r1 = (f_off + ce) * kK;
    // interpolated spectral envelope and aperiodic ratio (WORLD GetSpectralEnvelope / GetAperiodicRatio), in double like the
    // fp64 kernel; the two log spectra are what single precision takes over
    auto env_ar = [&](int k, double& env, double& ar) {
      const double s0 = fabs(plane_at<PT>(sp, r0 + k));
      double a0 = plane_at<PT>(ap, r0 + k);
      a0 = fmax(0.001, fmin(0.999999999999, a0));
      a0 *= a0;
      if (fl == ce) {
        env = s0;
        ar = a0;
      } else {
        const double s1 = fabs(plane_at<PT>(sp, r1 + k));
        double a1 = plane_at<PT>(ap, r1 + k);
        a1 = fmax(0.001, fmin(0.999999999999, a1));
        a1 *= a1;
        env = (1.0 - w) * s0 + w * s1;
        ar = (1.0 - w) * a0 + w * a1;
      }
    };
    double env0, ar0;
    env_ar(0, env0, ar0);
    const bool has_periodic = cur_vuv && !(ar0 > 0.999);
#pragma unroll 1   // (unrolling by four to batch the plane loads was measured SLOWER: 6.0 against 5.6 ms per 345 k pulses)
    for (int k = lane; k < kK; k += 32) {
      double env, ar;
      env_ar(k, env, ar);
      if (has_periodic) Lx[k] = 0.5f * __logf((float)(env * (1.0 - ar) + kMySafeGuardMinimum));
      La[k] = 0.5f * __logf((float)(cur_vuv ? env * ar : env));
    }
    __syncwarp();

    float2 per[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) per[j] = float2{0.0f, 0.0f};
    if (has_periodic) {
      minimum_phase(Lx, z, X, tw16, tw512, twn, lane);
      // fractional time shift: multiply bin k by cos(c k) - i |sin(c k)| (WORLD writes sqrt(1 - cos^2))
      const float coef = (float)(2.0 * kPi * pulse_shift[poff + p] * fs / kN);
      for (int k = lane; k < kK; k += 32) {
        const float2 v = X[k];
        float sn, re2;
        __sincosf(coef * (float)k, &sn, &re2);
        const float im2 = fabsf(sn);
        X[k] = float2{v.x * re2 + v.y * im2, v.y * re2 - v.x * im2};
      }
      __syncwarp();
      inverse_real(X, z, tw16, tw512, twn, lane, per);
      // WORLD RemoveDCComponent on the fft-shifted response s = (i + 512) mod 1024: dc = sum of the second half = the samples
      // x[0 .. 511] = pairs j < 8; the first half is REPLACED by -dc r[s], the second half gets -dc r[1023 - s]
      float dcs = 0.0f;
#pragma unroll
      for (int j = 0; j < 8; ++j) dcs += per[j].x + per[j].y;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dcs += __shfl_xor_sync(0xffffffffu, dcs, o);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n2 = 2 * (lane + 32 * j);
        if (j < 8) {  // s = n2 + 512 (and + 1): r index 1023 - s = 511 - n2 (and 510 - n2)
          per[j].x -= dcs * dcr[kM - 1 - n2];
          per[j].y -= dcs * dcr[kM - 2 - n2];
        } else {      // s = n2 - 512 (and + 1)
          per[j].x = -dcs * dcr[n2 - kM];
          per[j].y = -dcs * dcr[n2 - kM + 1];
        }
      }
    }
    // aperiodic response: minimum phase of sqrt(env ar) (voiced) or sqrt(env) (unvoiced) ...
    minimum_phase(La, z, X, tw16, tw512, twn, lane);
    // ... times the noise spectrum of this pulse's slice of the randn stream (WORLD GetNoiseSpectrum)
    {
      const int64_t start = (int64_t)n_p - n_first;
      double acc = 0.0;
      for (int i = lane; i < noise_size; i += 32) acc += (start + i < randn_len) ? randn_table[start + i] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      const double mean = noise_size > 0 ? acc / noise_size : 0.0;
      float2 v[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int i = 2 * (lane + 32 * q);
        v[q] = float2{0.0f, 0.0f};
        if (i < noise_size && start + i < randn_len) v[q].x = (float)(randn_table[start + i] - mean);
        if (i + 1 < noise_size && start + i + 1 < randn_len) v[q].y = (float)(randn_table[start + i + 1] - mean);
      }
      wfft512(z, v, tw16, tw512, lane);
      for_real_bins(z, twn, lane, [&](int k, float2 zn) { X[k] = cmul(X[k], zn); });
      __syncwarp();
    }
    float2 apr[16];
    inverse_real(X, z, tw16, tw512, twn, lane, apr);
    const float sqrt_noise = sqrtf((float)noise_size);
    float* out = response + row * (int64_t)kN;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int s = (2 * (lane + 32 * j) + kM) & (kN - 1);
      *reinterpret_cast<float2*>(out + s) = float2{(per[j].x * sqrt_noise + apr[j].x) * (1.0f / kN),
                                                  (per[j].y * sqrt_noise + apr[j].y) * (1.0f / kN)};
    }
  }
}

}  // namespace

int render_fast_launch(const void* sp, const void* ap, int plane_dtype, const int64_t* utt_frame_offset,
                       const int64_t* utt_pulse_offset, const int* num_pulses, int num_utts, const int* pulse_index,
                       const double* pulse_shift, const uint8_t* pulse_vuv, const double* randn_table, int64_t randn_len, int fs,
                       double frame_period_ms, int64_t total_rows, float* response, cudaStream_t st) {
  const double2* tw = twiddle_table(st);
  if (!tw) return check_launch("twiddle table");
  double dc_rs = 0.0;  // WORLD GetDCRemover: sum of the Hann-shaped weights (both halves), a constant of the fft size
  for (int i = 0; i < kN / 2; ++i) dc_rs += 2.0 * (0.5 - 0.5 * cos(2.0 * kPi * (i + 1.0) / (1.0 + kN)));
  const int smem = Smem::total_bytes;
  const int64_t want = (total_rows + kWarps - 1) / kWarps;
  const int64_t cap = (int64_t)148 * 4 * 16;  // a few waves of resident CTAs; the rest is the grid-stride loop
  const int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
  if (plane_dtype == B2W_F64) {
    cudaFuncSetAttribute(render_fast_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    render_fast_kernel<double><<<grid, 32 * kWarps, smem, st>>>(sp, ap, utt_frame_offset, utt_pulse_offset, num_pulses, num_utts,
                                                                pulse_index, pulse_shift, pulse_vuv, randn_table, randn_len, fs,
                                                                frame_period_ms, response, tw, dc_rs);
  } else {
    cudaFuncSetAttribute(render_fast_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    render_fast_kernel<float><<<grid, 32 * kWarps, smem, st>>>(sp, ap, utt_frame_offset, utt_pulse_offset, num_pulses, num_utts,
                                                               pulse_index, pulse_shift, pulse_vuv, randn_table, randn_len, fs,
                                                               frame_period_ms, response, tw, dc_rs);
  }
  return check_launch("render_fast_kernel");
}

}  // namespace b2w
