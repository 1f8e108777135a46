// WORLD synthesis on a ragged batch: pulse placement, per-pulse minimum-phase responses, atomic-free overlap-add.
//
// Replaces pyworld.synthesize(f0, sp, ap, fs) as called by WorldFeatLabelGen.world_features_to_raw
// (idiaptts/src/data_preparation/world/WorldFeatLabelGen.py:943) inside Synthesiser.run_world_synth
// (idiaptts/src/Synthesiser.py:51-65, a serial loop over utterances).  Algorithm: WORLD synthesis.cpp.
//
//   timebase     three kernels: per-sample phase increments (parallel), the running phase (strictly sequential per
//                utterance, as WORLD does it, because pulse positions are threshold decisions on its rounding), pulse
//                detection + ordered compaction (one CTA per utterance).
//   randn table  WORLD's randn() is xorshift128 (sum of 12 draws); Synthesis() reseeds it, and pulse p consumes exactly
//                12 * (n_{p+1} - n_p) steps, so the noise of pulse p is the slice [n_p - n_0, n_{p+1} - n_0) of ONE fixed
//                sequence.  The table kernel regenerates that sequence in parallel by GF(2) jump-ahead.
//   render       one CTA per pulse: interpolated envelope / aperiodicity, two minimum-phase spectra (each two real
//                FFTs), noise FFT, two inverse real FFTs, DC removal -> response[pulse][fft_size].
//   overlap-add  gather: each output sample sums, in pulse order (= the order WORLD accumulates in), the <= fft_size /
//                pulse-spacing responses that cover it.  No atomics, bit-reproducible.
#include <mutex>

#include "fft.cuh"

namespace b2w {

// ---- xorshift128 jump-ahead ------------------------------------------------------------------------------------------------
constexpr int kRandChunk = 256;   // randn values per thread
constexpr int kJumpLevels = 24;   // jump matrices T^(2^i), T = step^(12 * kRandChunk): up to 2^24 chunks
struct JumpTables {
  uint32_t col[kJumpLevels][128][4];  // column k of level i = image of state bit k
};
__device__ JumpTables g_jump;

struct XsState { uint32_t x, y, z, w; };
__host__ __device__ inline void xs_step(XsState& s) {
  const uint32_t t = s.x ^ (s.x << 11);
  s.x = s.y; s.y = s.z; s.z = s.w;
  s.w = (s.w ^ (s.w >> 19)) ^ (t ^ (t >> 8));
}

typedef uint32_t BitMat[128][4];  // column-major: m[k] = M e_k packed as (x, y, z, w)
static void bm_apply(const BitMat m, const uint32_t in[4], uint32_t out[4]) {
  out[0] = out[1] = out[2] = out[3] = 0;
  for (int k = 0; k < 128; ++k)
    if ((in[k >> 5] >> (k & 31)) & 1u)
      for (int q = 0; q < 4; ++q) out[q] ^= m[k][q];
}
static void bm_mul(const BitMat a, const BitMat b, BitMat c) {  // c = a * b
  for (int k = 0; k < 128; ++k) bm_apply(a, b[k], c[k]);
}

static int ensure_jump_tables() {
  static std::mutex mu;
  static bool done[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  std::lock_guard<std::mutex> lock(mu);
  if (done[dev]) return 0;
  static JumpTables host;
  static bool built = false;
  if (!built) {
    static BitMat step, acc, tmp, sq;
    for (int k = 0; k < 128; ++k) {
      XsState s{0, 0, 0, 0};
      (&s.x)[k >> 5] = 1u << (k & 31);
      xs_step(s);
      step[k][0] = s.x; step[k][1] = s.y; step[k][2] = s.z; step[k][3] = s.w;
    }
    // acc = step^(12 * kRandChunk) by square-and-multiply
    for (int k = 0; k < 128; ++k)
      for (int q = 0; q < 4; ++q) acc[k][q] = (q == (k >> 5)) ? (1u << (k & 31)) : 0u;  // identity
    memcpy(sq, step, sizeof(BitMat));
    for (unsigned e = 12u * kRandChunk; e; e >>= 1) {
      if (e & 1u) { bm_mul(sq, acc, tmp); memcpy(acc, tmp, sizeof(BitMat)); }
      bm_mul(sq, sq, tmp);
      memcpy(sq, tmp, sizeof(BitMat));
    }
    for (int i = 0; i < kJumpLevels; ++i) {
      memcpy(host.col[i], acc, sizeof(BitMat));
      bm_mul(acc, acc, tmp);
      memcpy(acc, tmp, sizeof(BitMat));
    }
    built = true;
  }
  if (cudaMemcpyToSymbol(g_jump, &host, sizeof(JumpTables)) != cudaSuccess) return -1;
  done[dev] = true;
  return 0;
}

__global__ void randn_table_kernel(double* __restrict__ table, int64_t n) {
  const int64_t chunk = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t first = chunk * kRandChunk;
  if (first >= n) return;
  uint32_t s[4] = {123456789u, 362436069u, 521288629u, 88675123u};  // randn_reseed()
  for (int lvl = 0; lvl < kJumpLevels; ++lvl) {
    if ((chunk >> lvl) & 1) {
      uint32_t o[4] = {0, 0, 0, 0};
      for (int k = 0; k < 128; ++k) {
        if ((s[k >> 5] >> (k & 31)) & 1u) {
#pragma unroll
          for (int q = 0; q < 4; ++q) o[q] ^= g_jump.col[lvl][k][q];
        }
      }
      s[0] = o[0]; s[1] = o[1]; s[2] = o[2]; s[3] = o[3];
    }
  }
  XsState st{s[0], s[1], s[2], s[3]};
  const int64_t last = min(n, first + kRandChunk);
  for (int64_t i = first; i < last; ++i) {
    uint32_t tmp = 0;
#pragma unroll
    for (int r = 0; r < 12; ++r) {
      xs_step(st);
      tmp += st.w >> 4;
    }
    table[i] = (double)tmp / 268435456.0 - 6.0;
  }
}

// ---- time base -----------------------------------------------------------------------------------------------------------------
constexpr int kTbThreads = 256;

// coarse contour value i in [0, T]: index T is WORLD's linear extrapolation 2 c[T-1] - c[T-2]
__device__ __forceinline__ void coarse_at(const double* __restrict__ f0, int T, double lowest_f0, int i, double& cf0, double& cv) {
  if (i < T) {
    const double v = f0[i];
    cf0 = v < lowest_f0 ? 0.0 : v;
    cv = cf0 == 0.0 ? 0.0 : 1.0;
  } else {
    double a0, v0, a1, v1;
    coarse_at(f0, T, lowest_f0, T - 1, a1, v1);
    if (T >= 2) coarse_at(f0, T, lowest_f0, T - 2, a0, v0); else { a0 = a1; v0 = v1; }
    cf0 = __dsub_rn(__dmul_rn(a1, 2.0), a0);
    cv = __dsub_rn(__dmul_rn(v1, 2.0), v0);
  }
}

// per-sample F0 used for the phase (500 Hz where the interpolated VUV is <= 0.5) and the thresholded VUV
__device__ __forceinline__ double sample_f0(const double* __restrict__ f0, int T, double lowest_f0, double fp, double fs, int n,
                                            int& k, double& vuv_out) {
  const double time = (double)n / fs;
  while (k < T && time >= __dmul_rn((double)k, fp)) ++k;     // coarse_t[k-1] <= time < coarse_t[k], k in [1, T]
  const double xl = __dmul_rn((double)(k - 1), fp), xr = __dmul_rn((double)k, fp);
  const double s = __ddiv_rn(__dsub_rn(time, xl), __dsub_rn(xr, xl));
  double fl, vl, fr, vr;
  coarse_at(f0, T, lowest_f0, k - 1, fl, vl);
  coarse_at(f0, T, lowest_f0, k, fr, vr);
  const double if0 = __dadd_rn(fl, __dmul_rn(s, __dsub_rn(fr, fl)));
  const double iv = __dadd_rn(vl, __dmul_rn(s, __dsub_rn(vr, vl)));
  vuv_out = iv > 0.5 ? 1.0 : 0.0;
  return vuv_out == 0.0 ? kDefaultF0 : if0;
}

// (1) phase increment of every output sample, parallel: inc[n] = 2 pi f0(n) / fs
__global__ void __launch_bounds__(kTbThreads)
phase_inc_kernel(const double* __restrict__ f0_all, const int64_t* __restrict__ utt_frame_offset,
                 const int64_t* __restrict__ utt_out_offset, int fs_i, double frame_period_ms, int fft_size,
                 double* __restrict__ phase) {
  const int u = blockIdx.y;
  const double* f0 = f0_all + utt_frame_offset[u];
  const int T = (int)(utt_frame_offset[u + 1] - utt_frame_offset[u]);
  const int64_t yoff = utt_out_offset[u];
  const int ylen = (int)(utt_out_offset[u + 1] - yoff);
  const int n = blockIdx.x * kTbThreads + threadIdx.x;
  if (n >= ylen || T < 1) return;
  const double fs = (double)fs_i;
  const double fp = frame_period_ms / 1000.0;
  const double lowest_f0 = (double)(fs_i / fft_size) + 1.0;
  // first guess of the frame interval by a multiplication (the two searches below and in sample_f0 correct it in either direction;
  // they, the time and the interpolation weight keep WORLD's exact divisions)
  int k = max(1, min(T, (int)((double)n * (1.0 / (fs * fp)))));
  while (k > 1 && (double)n / fs < __dmul_rn((double)(k - 1), fp)) --k;
  double v;
  const double f = sample_f0(f0, T, lowest_f0, fp, fs, n, k, v);
  phase[yoff + n] = __ddiv_rn(__dmul_rn(2.0 * kPi, f), fs);
}

// (2) WORLD accumulates the phase sequentially (total[n] = total[n-1] + inc[n]); floating-point addition is not
// associative and pulse positions are threshold decisions on fmod(total, 2 pi) - at 16 / 48 kHz the 500 Hz unvoiced
// default even produces exact ties - so the running sum is kept strictly sequential: one warp per utterance stages
// 1024 increments at a time in shared memory (coalesced), lane 0 runs the dependent DADD chain, the warp writes the
// totals back (coalesced).  ~10 cycles per sample, all utterances in parallel.
constexpr int kSeqChunk = 1024;
__global__ void __launch_bounds__(32) phase_scan_kernel(const int64_t* __restrict__ utt_out_offset, double* __restrict__ phase) {
  __shared__ double buf[kSeqChunk];
  const int u = blockIdx.x;
  const int lane = threadIdx.x;
  const int64_t yoff = utt_out_offset[u];
  const int ylen = (int)(utt_out_offset[u + 1] - yoff);
  double total = 0.0;
  for (int c0 = 0; c0 < ylen; c0 += kSeqChunk) {
    const int cn = min(kSeqChunk, ylen - c0);
    for (int i = lane; i < cn; i += 32) buf[i] = phase[yoff + c0 + i];
    __syncwarp();
    if (lane == 0) {
#pragma unroll 8
      for (int i = 0; i < cn; ++i) {
        total = __dadd_rn(total, buf[i]);
        buf[i] = total;
      }
    }
    __syncwarp();
    for (int i = lane; i < cn; i += 32) phase[yoff + c0 + i] = buf[i];
    __syncwarp();
  }
}

// (2') The same running sum, parallel AND bit-identical (scripts/exact_phase_scan_prototype.py, tests/test_oracle_synthesis.py):
// while the total stays inside one binade [2^e, 2^(e+1)) it is an integer multiple T of u = 2^(e-52) and
//     fl(T u + x) = (T + X + d) u,  X = floor(x / u),  rho = x - X u,  d = [rho > u/2], and for an exact tie d = (T + X) & 1,
// i.e. one addition is the integer map T -> T + X + d(parity of T).  A run of additions is described by two integers (the
// increment for an even / an odd incoming T) and these pairs compose associatively, so a block scan over them reproduces every
// sequentially rounded total.  A binade crossing (T reaching 2^53) is found by the scan; the crossing addition itself is one
// ordinary rounded add, after which the pass restarts with u doubled (~17 crossings per utterance).
// One CTA per utterance, tiles of 2048 increments (8 per thread).
constexpr int kScanThreads = 256;
constexpr int kScanPer = 8;
constexpr int kScanTile = kScanThreads * kScanPer;
constexpr long long kScanSat = 1ll << 60;   // increments saturate here: anything >= 2^53 is "crossed", only that matters

struct PMap { long long e, o; };            // increment for an even / odd incoming T
__device__ __forceinline__ long long psat(long long v) { return v > kScanSat ? kScanSat : v; }
__device__ __forceinline__ PMap pcompose(PMap a, PMap b) {  // a first, then b
  PMap r;
  r.e = psat(a.e + ((a.e & 1) ? b.o : b.e));
  r.o = psat(a.o + (((1 + a.o) & 1) ? b.o : b.e));
  return r;
}

__global__ void __launch_bounds__(kScanThreads) phase_scan_exact_kernel(const int64_t* __restrict__ utt_out_offset,
                                                                        double* __restrict__ phase) {
  __shared__ PMap warp_map[kScanThreads / 32];
  __shared__ double sh_t;
  __shared__ int sh_cross;
  const int u_idx = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t yoff = utt_out_offset[u_idx];
  const int n = (int)(utt_out_offset[u_idx + 1] - yoff);
  double* ph = phase + yoff;
  // prologue: the first samples sequentially (0 + x is exact, the total doubles every few samples at the start)
  const int pro = min(n, 64);
  if (tid == 0) {
    double t = 0.0;
    for (int i = 0; i < pro; ++i) {
      t = __dadd_rn(t, ph[i]);
      ph[i] = t;
    }
    sh_t = t;
  }
  __syncthreads();
  double t = sh_t;
  if (!(t >= 1e-200) || !(t < 1e200)) {  // degenerate totals (zero / tiny / huge / NaN): keep the plain sequential chain
    if (tid == 0) {
      for (int i = pro; i < n; ++i) {
        t = __dadd_rn(t, ph[i]);
        ph[i] = t;
      }
    }
    return;
  }
  for (int tile0 = pro; tile0 < n; tile0 += kScanTile) {
    const int tile1 = min(n, tile0 + kScanTile);
    const int base = tile0 + tid * kScanPer;      // this thread's elements base .. base + 7
    double x[kScanPer];
#pragma unroll
    for (int k = 0; k < kScanPer; ++k) x[k] = (base + k < tile1) ? ph[base + k] : 0.0;
    int p = tile0;                                 // elements before p are committed
    while (p < tile1) {
      // unit of the current binade
      const int e = (int)((__double_as_longlong(t) >> 52) & 0x7ff) - 1023;
      const double u = __longlong_as_double((long long)(e - 52 + 1023) << 52);
      const double inv_u = __longlong_as_double((long long)(52 - e + 1023) << 52);
      const long long T0 = (long long)(t * inv_u);
      const long long limit = 1ll << 53;
      PMap m[kScanPer];
      PMap mine = {0, 0};
#pragma unroll
      for (int k = 0; k < kScanPer; ++k) {
        m[k].e = m[k].o = 0;
        if (base + k >= p && base + k < tile1) {
          const double xs = x[k] * inv_u;          // exact power-of-two scaling
          const double Xf = floor(xs);
          const double rho = xs - Xf;              // exact, in [0, 1)
          long long X = (Xf >= 1.1529215046068469e18) ? kScanSat : (long long)Xf;
          const long long b = psat(X + (rho > 0.5 ? 1 : 0));
          const bool tie = rho == 0.5;
          m[k].e = b + ((tie && (X & 1)) ? 1 : 0);
          m[k].o = b + ((tie && !(X & 1)) ? 1 : 0);
        }
        mine = pcompose(mine, m[k]);
      }
      // block exclusive scan of the per-thread maps under pcompose
      PMap incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        PMap prev;
        prev.e = __shfl_up_sync(0xffffffffu, incl.e, o);
        prev.o = __shfl_up_sync(0xffffffffu, incl.o, o);
        if (lane >= o) incl = pcompose(prev, incl);
      }
      PMap excl;
      excl.e = __shfl_up_sync(0xffffffffu, incl.e, 1);
      excl.o = __shfl_up_sync(0xffffffffu, incl.o, 1);
      if (lane == 0) excl.e = excl.o = 0;
      __syncthreads();                             // previous pass has read warp_map / sh_cross / sh_t
      if (lane == 31) warp_map[warp] = incl;
      if (tid == 0) sh_cross = tile1;
      __syncthreads();
      PMap pre = {0, 0};
      for (int w = 0; w < warp; ++w) pre = pcompose(pre, warp_map[w]);
      pre = pcompose(pre, excl);
      // expand this thread's elements from its incoming T
      long long Tk = T0 + ((T0 & 1) ? pre.o : pre.e);
      bool alive = Tk < limit;                     // false: a crossing happened before this thread's first element
      double outv[kScanPer];
      int my_cross = tile1;
      long long T_before_cross = Tk;
#pragma unroll
      for (int k = 0; k < kScanPer; ++k) {
        outv[k] = 0.0;
        if (base + k >= p && base + k < tile1 && alive) {
          const long long Tn = Tk + ((Tk & 1) ? m[k].o : m[k].e);
          if (Tn >= limit) {
            alive = false;
            my_cross = base + k;
            T_before_cross = Tk;
          } else {
            Tk = Tn;
            outv[k] = (double)Tk * u;              // exact: Tk < 2^53, u a power of two
          }
        }
      }
      if (my_cross < tile1) atomicMin(&sh_cross, my_cross);
      __syncthreads();
      const int cross = sh_cross;
#pragma unroll
      for (int k = 0; k < kScanPer; ++k)
        if (base + k >= p && base + k < min(cross, tile1)) ph[base + k] = outv[k];
      if (cross < tile1) {
        // the crossing addition: an ordinary rounded add from the last committed total
        if (my_cross == cross) {
          const double tb = (double)T_before_cross * u;
          const double tn = __dadd_rn(tb, x[cross - base]);
          ph[cross] = tn;
          sh_t = tn;
        }
        p = cross + 1;
      } else {
        // the owner of the last element publishes the total
        if (tile1 - 1 >= base && tile1 - 1 < base + kScanPer) sh_t = (double)Tk * u;
        p = tile1;
      }
      __syncthreads();
      t = sh_t;
    }
  }
}

// (3) pulse detection and ordered compaction: a pulse sits at sample i when |wrap[i+1] - wrap[i]| > pi, wrap = fmod(total, 2 pi).
// Every utterance is cut into chunks of kPulseChunk jumps, one CTA each (a 256-utterance batch used to keep 256 CTAs busy for
// 1.4 ms with strided loads): pass 1 counts the pulses of every chunk, pass 2 adds up the counts of the chunks before it and
// writes its pulses in order.  Loads are coalesced, one fmod per jump (the neighbour's value comes by shuffle).
constexpr int kPulseChunk = 4096;

// fmod(x, 2 pi) for x >= 0, bit-identical to the C library function (whose result is the EXACT remainder) in a handful of
// instructions instead of the ~450 of the generic fmod: with q = floor(fl(x / y)) in {floor(x / y), floor(x / y) + 1} (rounding is
// monotonic, so the quotient is never under-estimated), r = x - q y is a multiple of ulp(y) = 2^-50 of magnitude < 8, hence
// exactly representable, and the fused multiply-add delivers it without rounding; q one too large shows as r < 0.
__device__ __forceinline__ double fmod_two_pi(double x) {
  const double y = 2.0 * kPi;
  if (x < y) return x;
  const double q = floor(__ddiv_rn(x, y));
  double r = __fma_rn(-q, y, x);
  if (r < 0.0) r = __fma_rn(-(q - 1.0), y, x);
  return r;
}

template <bool WRITE>
__global__ void __launch_bounds__(kTbThreads)
pulse_chunk_kernel(const double* __restrict__ f0_all, const int64_t* __restrict__ utt_frame_offset,
                   const int64_t* __restrict__ utt_out_offset, const int64_t* __restrict__ utt_pulse_offset, int fs_i,
                   double frame_period_ms, int fft_size, const double* __restrict__ phase, int* __restrict__ chunk_counts,
                   int max_chunks, int* __restrict__ pulse_index, double* __restrict__ pulse_shift, uint8_t* __restrict__ pulse_vuv,
                   int* __restrict__ num_pulses, int* __restrict__ status) {
  __shared__ int sh_w[kTbThreads / 32];
  // write pass: the pulses of the chunk are first collected (sample index + the two wrapped phases) and then finished by consecutive
  // threads -- a pulse occurs once in ~100 samples, so finishing it where it is found ran the expensive part (three fp64 divisions,
  // the frame search, the F0 / VUV interpolation) with one or two active lanes per warp, ~50 times per chunk
  constexpr int kList = WRITE ? 640 : 1;   // >= kPulseChunk * 1200 / 8000: the slab bound of b2w_synth_max_pulses down to fs = 8 kHz
  __shared__ int li[kList];
  __shared__ double lw0[kList], lw1[kList];
  const int u = blockIdx.y, c = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = (int)(utt_frame_offset[u + 1] - utt_frame_offset[u]);
  const int64_t yoff = utt_out_offset[u];
  const int ylen = (int)(utt_out_offset[u + 1] - yoff);
  const int njump = (T < 1 || ylen < 2) ? 0 : ylen - 1;
  const int j0 = c * kPulseChunk;
  int* my_count = chunk_counts + (int64_t)u * max_chunks + c;
  if (j0 >= njump) {
    if (!WRITE && tid == 0) *my_count = 0;
    if (WRITE && c == 0 && njump == 0 && tid == 0) num_pulses[u] = 0;
    return;
  }
  const int j1 = min(njump, j0 + kPulseChunk);
  const double two_pi = 2.0 * kPi;
  const double* tot = phase + yoff;
  int running = 0;
  if (WRITE) {
    for (int k = 0; k < c; ++k) running += chunk_counts[(int64_t)u * max_chunks + k];  // <= a few dozen values, L2 hits
  }
  const double* f0 = f0_all + utt_frame_offset[u];
  const int64_t poff = utt_pulse_offset[u];
  const int cap = (int)(utt_pulse_offset[u + 1] - poff);
  const double fs = (double)fs_i;
  const double fp = frame_period_ms / 1000.0;
  const double lowest_f0 = (double)(fs_i / fft_size) + 1.0;
  const int run0 = running;  // pulses of the utterance before this chunk
  int warp_total = 0;  // pass 1: pulses seen by this warp
  for (int s0 = j0; s0 < j1; s0 += kTbThreads) {
    const int i = s0 + tid;
    const bool valid = i < j1;
    // wrap[i] for this lane's jump; wrap[i+1] from the next lane (the last lane of a warp computes it itself)
    const double w0 = valid ? fmod_two_pi(tot[i]) : 0.0;
    double w1 = __shfl_down_sync(0xffffffffu, w0, 1);
    if (valid && (lane == 31 || i + 1 >= j1)) w1 = fmod_two_pi(tot[i + 1]);
    const bool is_pulse = valid && fabs(w1 - w0) > kPi;
    const unsigned ball = __ballot_sync(0xffffffffu, is_pulse);
    if (!WRITE) {
      warp_total += __popc(ball);
      continue;
    }
    if (lane == 0) sh_w[warp] = __popc(ball);
    __syncthreads();
    int before = 0, slab = 0;
#pragma unroll
    for (int w = 0; w < kTbThreads / 32; ++w) {
      const int v = sh_w[w];
      before += (w < warp) ? v : 0;
      slab += v;
    }
    auto finish = [&](int slot, int i_, double w0_, double w1_) {
      const double y1 = w0_ - two_pi;
      const double x = -y1 / (w1_ - y1);
      int k = max(1, min(T, (int)((double)i_ / fs / fp)));
      while (k > 1 && (double)i_ / fs < __dmul_rn((double)(k - 1), fp)) --k;
      double v;
      sample_f0(f0, T, lowest_f0, fp, fs, i_, k, v);
      pulse_index[poff + slot] = i_;
      pulse_shift[poff + slot] = x / fs;
      pulse_vuv[poff + slot] = v > 0.5 ? 1 : 0;
    };
    if (is_pulse) {
      const int slot = running + before + __popc(ball & ((1u << lane) - 1u));
      if (slot < cap) {
        const int loc = slot - run0;
        if (loc < kList) {
          li[loc] = i;
          lw0[loc] = w0;
          lw1[loc] = w1;
        } else {
          finish(slot, i, w0, w1);  // (more pulses in one chunk than the list holds: only with an F0 far above WORLD's range)
        }
      }
    }
    running += slab;
    __syncthreads();  // sh_w is rewritten by the next slab (and, after the last one, the list is complete)
    if (s0 + kTbThreads >= j1) {
      const int n_list = min(min(running, cap) - run0, kList);
      for (int t = tid; t < n_list; t += kTbThreads) finish(run0 + t, li[t], lw0[t], lw1[t]);
    }
  }
  if (!WRITE) {
    if (lane == 0) sh_w[warp] = warp_total;
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < kTbThreads / 32; ++w) t += sh_w[w];
      *my_count = t;
    }
  } else if (j1 == njump && tid == 0) {  // the last chunk of the utterance knows the total
    if (running > cap) atomicOr(status, B2W_STATUS_F0_TOO_HIGH);
    num_pulses[u] = min(running, cap);
  }
}

// ---- render -------------------------------------------------------------------------------------------------------------------
template <int N>
struct RenderSmem {
  static constexpr int M = N / 2;
  static constexpr int K = N / 2 + 1;
  static constexpr int z_doubles = 2 * zp_size(M);
  static constexpr int x_doubles = 2 * (K + 1);   // complex spectrum staging
  static constexpr int env_doubles = K + 1;
  static constexpr int ar_doubles = K + 1;
  static constexpr int per_doubles = N;            // periodic response
  static constexpr int red_doubles = 32;
  static constexpr int tw_doubles = 2 * kFftTwEntries;  // shared-memory twiddles of cfft_s
  static constexpr int total_bytes =
      (z_doubles + x_doubles + env_doubles + ar_doubles + per_doubles + red_doubles + tw_doubles) * 8;
};

// WORLD GetMinimumPhaseSpectrum for the log-amplitude L[0..N/2] held in `logsp` -> X[0..N/2] (complex) in `X`.
template <int N, int NT>
__device__ __forceinline__ void minimum_phase(const double* logsp, double2* z, double2* X, const double2* tws, double2 rw0,
                                              double2 rstep, int tid) {
  constexpr int M = N / 2, H = N / 2, K = H + 1;
  for (int k = tid; k < K; k += NT) {
    const double v = logsp[k];
    zreal<double>(z, k) = v;
    if (k > 0 && k < H) zreal<double>(z, N - k) = v;
  }
  __syncthreads();
  cfft_s<double, M, NT>(z, tws, tid);
  // cepstrum (real for a symmetric input), folded: c[0], 2 c[1..N/2-1], c[N/2], zeros
  {
    double2 rw = rw0;
    for (int k = tid; k < K; k += NT) {
      const double c = rfft_bin_w<double, N>(z, k, rw).x;
      rw = cmul(rw, rstep);
      X[k].x = (k == 0 || k == H) ? c : 2.0 * c;
    }
  }
  __syncthreads();
  for (int n = tid; n < N; n += NT) zreal<double>(z, n) = (n <= H) ? X[n].x : 0.0;
  __syncthreads();
  cfft_s<double, M, NT>(z, tws, tid);
  double2 rw = rw0;
  for (int k = tid; k < K; k += NT) {
    const double2 s = rfft_bin_w<double, N>(z, k, rw);
    rw = cmul(rw, rstep);
    const double mag = exp(s.x / N);
    double sn, cs;
    sincos(s.y / N, &sn, &cs);
    X[k] = make_double2(mag * cs, mag * sn);
  }
  __syncthreads();
}

// Unnormalised inverse real FFT of the Hermitian half spectrum X[0..N/2], fft-shifted: out(j) = x[(j + N/2) mod N].
template <int N, int NT, typename Emit>
__device__ __forceinline__ void inverse_real_shifted(const double2* X, double2* z, const double2* tws, double2 rw0, double2 rstep,
                                                     int tid, Emit emit) {
  constexpr int M = N / 2;
  double2 rw = rw0;
  for (int k = tid; k < M; k += NT) {
    double2 a = X[k];
    double2 bq = X[M - k];
    if (k == 0) {  // a c2r transform ignores the imaginary parts of the DC and Nyquist bins (WORLD / FFTW semantics);
      a.y = 0.0;   // the fractional time shift makes the Nyquist bin complex
      bq.y = 0.0;
    }
    const double2 b = make_double2(bq.x, -bq.y);                 // conj(X[M - k]) = X[k + M]
    const double2 e = make_double2(a.x + b.x, a.y + b.y);
    const double2 d = make_double2(a.x - b.x, a.y - b.y);
    const double2 wc = make_double2(rw.x, -rw.y);                // W^-k
    rw = cmul(rw, rstep);
    const double2 o = cmul(d, wc);
    // Y = e + i * o ; store conj(Y) so that a forward FFT followed by a conjugate gives the inverse transform
    z[ZP(k)] = make_double2(e.x - o.y, -(e.y + o.x));
  }
  __syncthreads();
  cfft_s<double, M, NT>(z, tws, tid);
  for (int n = tid; n < M; n += NT) {
    const double2 v = z[ZP(n)];
    // x[2n] = Re, x[2n+1] = -Im ; shifted position j = (i + N/2) mod N
    emit((2 * n + M) & (N - 1), v.x);
    emit((2 * n + 1 + M) & (N - 1), -v.y);
  }
  __syncthreads();
}

template <typename PT>
__device__ __forceinline__ double plane_at(const void* p, int64_t idx) { return (double)reinterpret_cast<const PT*>(p)[idx]; }

#ifndef B2W_RENDER_DIV
#define B2W_RENDER_DIV 8  // threads per pulse = N / B2W_RENDER_DIV (128 at N = 1024: measured 5 % faster than 64)
#endif
template <int N, typename PT>
__global__ void __launch_bounds__(N / B2W_RENDER_DIV)
render_kernel(const void* __restrict__ sp, const void* __restrict__ ap, const int64_t* __restrict__ utt_frame_offset,
              const int64_t* __restrict__ utt_pulse_offset, const int* __restrict__ num_pulses,
              const int* __restrict__ pulse_index, const double* __restrict__ pulse_shift,
              const uint8_t* __restrict__ pulse_vuv, const double* __restrict__ randn_table, int64_t randn_len, int fs_i,
              double frame_period_ms, double* __restrict__ response, const double2* __restrict__ tw, double dc_rs) {
  constexpr int NT = N / B2W_RENDER_DIV;
  constexpr int H = N / 2, K = H + 1;
  const int u = blockIdx.y;
  const int P = num_pulses[u];
  const int p = blockIdx.x;
  if (p >= P) return;
  extern __shared__ double smem[];
  double2* z = reinterpret_cast<double2*>(smem);
  double2* X = reinterpret_cast<double2*>(smem + RenderSmem<N>::z_doubles);
  double* env = smem + RenderSmem<N>::z_doubles + RenderSmem<N>::x_doubles;
  double* ar = env + RenderSmem<N>::env_doubles;
  double* periodic = ar + RenderSmem<N>::ar_doubles;
  double* red = periodic + RenderSmem<N>::per_doubles;
  double2* tws = reinterpret_cast<double2*>(red + RenderSmem<N>::red_doubles);
  const int tid = threadIdx.x;
  fft_tw_fill<double, N / 2, NT>(tws, tw, tid);
  double2 rw0, rstep;  // phasors of the real-FFT pre / post passes (bins tid, tid + NT, ...)
  rfft_rot_init<double, N, NT>(tw, tid, rw0, rstep);
  const double fs = (double)fs_i;
  const double fp = frame_period_ms / 1000.0;
  const int64_t poff = utt_pulse_offset[u];
  const int64_t f_off = utt_frame_offset[u];
  const int T = (int)(utt_frame_offset[u + 1] - f_off);
  const int n_p = pulse_index[poff + p];
  const int n_next = pulse_index[poff + min(P - 1, p + 1)];
  const int n_first = pulse_index[poff];
  const int noise_size = n_next - n_p;
  const double cur_vuv = pulse_vuv[poff + p] ? 1.0 : 0.0;
  const double cur_time = (double)n_p / fs;
  const int fl = min(T - 1, (int)floor(cur_time / fp));
  const int ce = min(T - 1, (int)ceil(cur_time / fp));
  const double w = cur_time / fp - fl;
  // interpolated spectral envelope and aperiodic ratio (WORLD GetSpectralEnvelope / GetAperiodicRatio)
  for (int k = tid; k < K; k += NT) {
    const double s0 = fabs(plane_at<PT>(sp, (f_off + fl) * K + k));
    double a0 = plane_at<PT>(ap, (f_off + fl) * K + k);
    a0 = fmax(0.001, fmin(0.999999999999, a0));
    a0 *= a0;
    if (fl == ce) {
      env[k] = s0;
      ar[k] = a0;
    } else {
      const double s1 = fabs(plane_at<PT>(sp, (f_off + ce) * K + k));
      double a1 = plane_at<PT>(ap, (f_off + ce) * K + k);
      a1 = fmax(0.001, fmin(0.999999999999, a1));
      a1 *= a1;
      env[k] = (1.0 - w) * s0 + w * s1;
      ar[k] = (1.0 - w) * a0 + w * a1;
    }
  }
  __syncthreads();
  const bool has_periodic = !(cur_vuv <= 0.5 || ar[0] > 0.999);
  if (has_periodic) {
    double* L = periodic;  // log-spectrum staging shares the periodic buffer until the response is emitted
    for (int k = tid; k < K; k += NT) L[k] = log(env[k] * (1.0 - ar[k]) + kMySafeGuardMinimum) / 2.0;
    __syncthreads();
    minimum_phase<N, NT>(L, z, X, tws, rw0, rstep, tid);
    // fractional time shift: multiply bin k by (cos(c k) - i sqrt(1 - cos^2(c k)))
    const double coef = 2.0 * kPi * pulse_shift[poff + p] * fs / N;
    for (int k = tid; k < K; k += NT) {
      const double2 v = X[k];
      const double re2 = cos(coef * k);
      const double im2 = sqrt(1.0 - re2 * re2);
      X[k] = make_double2(v.x * re2 + v.y * im2, v.y * re2 - v.x * im2);
    }
    __syncthreads();
    inverse_real_shifted<N, NT>(X, z, tws, rw0, rstep, tid, [&](int j, double v) { periodic[j] = v; });
    // remove the DC component (WORLD RemoveDCComponent with GetDCRemover's Hann-shaped weights)
    // dc_remover[i] = hann(i) / sum, hann(i) = 0.5 - 0.5 cos(2 pi (i + 1) / (1 + N)): the sum depends on N only (dc_rs, computed
    // once on the host), the cosines of i = tid, tid + NT, ... come from a rotation recurrence instead of one cos() each
    double dcs = 0.0;
    for (int i = tid; i < H; i += NT) dcs += periodic[H + i];
    dcs = block_sum<NT>(dcs, red);
    {
      double c, sn, cd, sd;
      sincos(2.0 * kPi * (tid + 1.0) / (1.0 + N), &sn, &c);
      sincos(2.0 * kPi * (double)NT / (1.0 + N), &sd, &cd);
      const double scale = dcs / dc_rs;
      for (int i = tid; i < H; i += NT) {
        const double r = (0.5 - 0.5 * c) * scale;  // dc_remover[i] == dc_remover[N-1-i]
        periodic[i] = -r;
        periodic[N - 1 - i] -= r;
        const double cn = c * cd - sn * sd;
        sn = sn * cd + c * sd;
        c = cn;
      }
    }
    __syncthreads();
  }
  // aperiodic response: minimum phase of sqrt(env * ar) (voiced) or sqrt(env) (unvoiced) ...
  for (int k = tid; k < K; k += NT) {
    const double e = env[k];
    env[k] = (cur_vuv != 0.0) ? log(e * ar[k]) / 2.0 : log(e) / 2.0;
  }
  __syncthreads();
  minimum_phase<N, NT>(env, z, X, tws, rw0, rstep, tid);
  // ... times the noise spectrum of this pulse's slice of the randn stream (WORLD GetNoiseSpectrum); the noise FFT reuses the
  // FFT buffer after the minimum-phase spectrum is in X, and its bins are multiplied into X on the fly (no noise buffer)
  {
    const int64_t start = (int64_t)n_p - n_first;
    double acc = 0.0;
    for (int i = tid; i < noise_size; i += NT) acc += (start + i < randn_len) ? randn_table[start + i] : 0.0;
    const double tot = block_sum<NT>(acc, red);
    const double mean = noise_size > 0 ? tot / noise_size : 0.0;
    for (int i = tid; i < N; i += NT) {
      double v = 0.0;
      if (i < noise_size && start + i < randn_len) v = randn_table[start + i] - mean;
      zreal<double>(z, i) = v;
    }
    __syncthreads();
    cfft_s<double, N / 2, NT>(z, tws, tid);
    double2 rw = rw0;
    for (int k = tid; k < K; k += NT) {
      X[k] = cmul(X[k], rfft_bin_w<double, N>(z, k, rw));
      rw = cmul(rw, rstep);
    }
  }
  __syncthreads();
  const double sqrt_noise = sqrt((double)noise_size);
  double* out = response + (poff + p) * (int64_t)N;
  inverse_real_shifted<N, NT>(X, z, tws, rw0, rstep, tid, [&](int j, double v) {
    const double per_v = has_periodic ? periodic[j] : 0.0;
    out[j] = (per_v * sqrt_noise + v) / N;
  });
}

// ---- overlap-add ---------------------------------------------------------------------------------------------------------------
constexpr int kOlaTile = 1024;  // output samples per CTA (4 per thread): the two binary searches are amortised over 8 KB of output
template <typename OT, typename RT>
__global__ void __launch_bounds__(256)
overlap_add_kernel(const RT* __restrict__ response, const int64_t* __restrict__ utt_out_offset,
                   const int64_t* __restrict__ utt_pulse_offset, const int* __restrict__ num_pulses,
                   const int* __restrict__ pulse_index, int fft_size, OT* __restrict__ y) {
  const int u = blockIdx.y;
  const int64_t yoff = utt_out_offset[u];
  const int ylen = (int)(utt_out_offset[u + 1] - yoff);
  const int n0 = blockIdx.x * kOlaTile;
  if (n0 >= ylen) return;
  const int64_t poff = utt_pulse_offset[u];
  const int P = num_pulses[u];
  const int* idx = pulse_index + poff;
  const int H = fft_size / 2;
  // first pulse whose response can reach sample n0: n_p + H >= n0  <=>  n_p >= n0 - H
  int lo = 0, hi = P;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (idx[mid] < n0 - H) lo = mid + 1; else hi = mid;
  }
  // one past the last pulse whose response starts inside this tile: n_p - H + 1 <= n0 + kOlaTile - 1
  int end = lo;
  hi = P;
  while (end < hi) {
    const int mid = (end + hi) >> 1;
    if (idx[mid] - H + 1 <= n0 + kOlaTile - 1) end = mid + 1; else hi = mid;
  }
  // Two pulses x four samples per step: eight independent streaming 8-byte loads in flight per thread (the kernel is HBM-bound;
  // one dependent load per thread and pulse left the memory system under-subscribed).  Every sample still sums its pulses in
  // pulse order: bit-identical to the sequential loop.
  const RT* rbase = response + poff * (int64_t)fft_size;
  const int n = n0 + threadIdx.x;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  int p = lo;
  for (; p + 2 <= end; p += 2) {
    const int s0 = idx[p] - H + 1, s1 = idx[p + 1] - H + 1;
    double v[2][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j0 = n + 256 * q - s0, j1 = n + 256 * q - s1;
      v[0][q] = (j0 >= 0 && j0 < fft_size) ? (double)__ldcs(rbase + (int64_t)p * fft_size + j0) : 0.0;
      v[1][q] = (j1 >= 0 && j1 < fft_size) ? (double)__ldcs(rbase + (int64_t)(p + 1) * fft_size + j1) : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      acc[q] += v[0][q];
      acc[q] += v[1][q];
    }
  }
  for (; p < end; ++p) {
    const int s0 = idx[p] - H + 1;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = n + 256 * q - s0;
      if (j >= 0 && j < fft_size) acc[q] += (double)__ldcs(rbase + (int64_t)p * fft_size + j);
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (n + 256 * q < ylen) y[yoff + n + 256 * q] = (OT)acc[q];
}

// y[n] = x[n] + p y[n-1] (scipy.signal.lfilter([1], [1, -p])) on the float32-rounded waveform, one thread per utterance
__global__ void deemphasis_kernel(const int64_t* __restrict__ utt_out_offset, int num_utts, double p, double* __restrict__ y) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= num_utts) return;
  const int64_t b = utt_out_offset[u], e = utt_out_offset[u + 1];
  double prev = 0.0;
  for (int64_t i = b; i < e; ++i) {
    prev = __dadd_rn((double)(float)y[i], __dmul_rn(p, prev));
    y[i] = prev;
  }
}

}  // namespace b2w

extern "C" int64_t b2w_synth_max_pulses(int64_t y_length, int32_t fs) {
  return (int64_t)((double)y_length * 1200.0 / (double)fs) + 64;
}

extern "C" int64_t b2w_synth_timebase_chunks(int64_t max_out_per_utt) {
  return (max_out_per_utt + b2w::kPulseChunk - 1) / b2w::kPulseChunk;
}

extern "C" int b2w_synth_randn_table(double* table, int64_t n, void* stream) {
  using namespace b2w;
  B2W_REQUIRE(table, "b2w_synth_randn_table: null argument");
  if (n == 0) return 0;
  B2W_REQUIRE(n / kRandChunk < ((int64_t)1 << kJumpLevels), "b2w_synth_randn_table: table too long");
  B2W_REQUIRE(ensure_jump_tables() == 0, "b2w_synth_randn_table: cannot upload jump tables");
  const int64_t chunks = (n + kRandChunk - 1) / kRandChunk;
  randn_table_kernel<<<(unsigned)((chunks + 63) / 64), 64, 0, (cudaStream_t)stream>>>(table, n);
  return check_launch("randn_table_kernel");
}

extern "C" int b2w_synth_timebase(const double* f0, const int64_t* utt_frame_offset, const int64_t* utt_out_offset,
                                  const int64_t* utt_pulse_offset, int32_t num_utts, int64_t max_out_per_utt, int32_t fs,
                                  double frame_period_ms, int32_t fft_size, double* phase_ws, int32_t* chunk_ws,
                                  int32_t* pulse_index, double* pulse_shift, uint8_t* pulse_vuv, int32_t* num_pulses,
                                  int32_t* status, void* stream) {
  using namespace b2w;
  B2W_REQUIRE(f0 && utt_frame_offset && utt_out_offset && utt_pulse_offset && phase_ws && chunk_ws && pulse_index && pulse_shift &&
                  pulse_vuv && num_pulses && status,
              "b2w_synth_timebase: null argument");
  B2W_REQUIRE(num_utts <= 65535, "b2w_synth_timebase: at most 65535 utterances per call");
  if (num_utts == 0 || max_out_per_utt == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((max_out_per_utt + kTbThreads - 1) / kTbThreads), (unsigned)num_utts);
  phase_inc_kernel<<<grid, kTbThreads, 0, st>>>(f0, utt_frame_offset, utt_out_offset, fs, frame_period_ms, fft_size, phase_ws);
  int rc = check_launch("phase_inc_kernel");
  if (rc) return rc;
#ifdef B2W_SEQUENTIAL_PHASE_SCAN
  phase_scan_kernel<<<num_utts, 32, 0, st>>>(utt_out_offset, phase_ws);
  rc = check_launch("phase_scan_kernel");
#else
  phase_scan_exact_kernel<<<num_utts, kScanThreads, 0, st>>>(utt_out_offset, phase_ws);
  rc = check_launch("phase_scan_exact_kernel");
#endif
  if (rc) return rc;
  const int max_chunks = (int)((max_out_per_utt + kPulseChunk - 1) / kPulseChunk);
  int* chunk_counts = chunk_ws;
  dim3 pgrid((unsigned)max_chunks, (unsigned)num_utts);
  pulse_chunk_kernel<false><<<pgrid, kTbThreads, 0, st>>>(f0, utt_frame_offset, utt_out_offset, utt_pulse_offset, fs, frame_period_ms,
                                                         fft_size, phase_ws, chunk_counts, max_chunks, pulse_index, pulse_shift,
                                                         pulse_vuv, num_pulses, status);
  rc = check_launch("pulse_chunk_kernel<count>");
  if (rc) return rc;
  pulse_chunk_kernel<true><<<pgrid, kTbThreads, 0, st>>>(f0, utt_frame_offset, utt_out_offset, utt_pulse_offset, fs, frame_period_ms,
                                                        fft_size, phase_ws, chunk_counts, max_chunks, pulse_index, pulse_shift,
                                                        pulse_vuv, num_pulses, status);
  return check_launch("pulse_chunk_kernel<write>");
}

extern "C" int b2w_synth_render(const void* sp, const void* ap, int32_t plane_dtype, const int64_t* utt_frame_offset,
                                const int64_t* utt_pulse_offset, const int32_t* num_pulses, int32_t num_utts,
                                const int32_t* pulse_index, const double* pulse_shift, const uint8_t* pulse_vuv,
                                const double* randn_table, int64_t randn_table_len, int32_t fs, double frame_period_ms,
                                int32_t fft_size, int64_t max_pulses_per_utt, double* response, void* stream) {
  using namespace b2w;
  B2W_REQUIRE(sp && ap && utt_frame_offset && utt_pulse_offset && num_pulses && pulse_index && pulse_shift && pulse_vuv &&
                  randn_table && response,
              "b2w_synth_render: null argument");
  B2W_REQUIRE(plane_dtype == B2W_F64 || plane_dtype == B2W_F32, "b2w_synth_render: bad plane dtype %d", plane_dtype);
  B2W_REQUIRE(num_utts <= 65535, "b2w_synth_render: at most 65535 utterances per call");
  if (num_utts == 0 || max_pulses_per_utt == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const double2* tw = twiddle_table(st);
  if (!tw) return check_launch("twiddle table");
  dim3 grid((unsigned)max_pulses_per_utt, (unsigned)num_utts);
  double dc_rs = 0.0;  // WORLD GetDCRemover: sum of the Hann-shaped weights (both halves), a constant of the fft size
  for (int i = 0; i < fft_size / 2; ++i) dc_rs += 2.0 * (0.5 - 0.5 * cos(2.0 * kPi * (i + 1.0) / (1.0 + fft_size)));
#define B2W_RENDER_LAUNCH(NN, PT)                                                                                          \
  do {                                                                                                                     \
    const int smem = RenderSmem<NN>::total_bytes;                                                                          \
    cudaFuncSetAttribute(render_kernel<NN, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                        \
    render_kernel<NN, PT><<<grid, NN / B2W_RENDER_DIV, smem, st>>>(sp, ap, utt_frame_offset, utt_pulse_offset, num_pulses, pulse_index, \
                                                       pulse_shift, pulse_vuv, randn_table, randn_table_len, fs,           \
                                                       frame_period_ms, response, tw, dc_rs);                              \
  } while (0)
  switch (fft_size) {
    case 512: if (plane_dtype == B2W_F64) B2W_RENDER_LAUNCH(512, double); else B2W_RENDER_LAUNCH(512, float); break;
    case 1024: if (plane_dtype == B2W_F64) B2W_RENDER_LAUNCH(1024, double); else B2W_RENDER_LAUNCH(1024, float); break;
    case 2048: if (plane_dtype == B2W_F64) B2W_RENDER_LAUNCH(2048, double); else B2W_RENDER_LAUNCH(2048, float); break;
    case 4096: if (plane_dtype == B2W_F64) B2W_RENDER_LAUNCH(4096, double); else B2W_RENDER_LAUNCH(4096, float); break;
    default: set_error("b2w_synth_render: unsupported fft_size %d", fft_size); return -1;
  }
#undef B2W_RENDER_LAUNCH
  return check_launch("render_kernel");
}

static int overlap_add_launch(const void* response, int response_is_f32, const int64_t* utt_out_offset,
                              const int64_t* utt_pulse_offset, const int32_t* num_pulses, int32_t num_utts, const int32_t* pulse_index,
                              int32_t fft_size, int64_t max_out_per_utt, double deemphasis, void* y, int32_t y_dtype, void* stream) {
  using namespace b2w;
  B2W_REQUIRE(response && utt_out_offset && utt_pulse_offset && num_pulses && pulse_index && y,
              "b2w_synth_overlap_add: null argument");
  B2W_REQUIRE(y_dtype == B2W_F64 || y_dtype == B2W_F32, "b2w_synth_overlap_add: bad y dtype %d", y_dtype);
  B2W_REQUIRE(deemphasis == 0.0 || y_dtype == B2W_F64, "b2w_synth_overlap_add: de-emphasis needs a float64 output");
  B2W_REQUIRE(num_utts <= 65535, "b2w_synth_overlap_add: at most 65535 utterances per call");
  if (num_utts == 0 || max_out_per_utt == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((max_out_per_utt + kOlaTile - 1) / kOlaTile), (unsigned)num_utts);
#define B2W_OLA(OT, RT) \
  overlap_add_kernel<OT, RT><<<grid, 256, 0, st>>>((const RT*)response, utt_out_offset, utt_pulse_offset, num_pulses, pulse_index, fft_size, (OT*)y)
  if (response_is_f32) {
    if (y_dtype == B2W_F64) B2W_OLA(double, float); else B2W_OLA(float, float);
  } else {
    if (y_dtype == B2W_F64) B2W_OLA(double, double); else B2W_OLA(float, double);
  }
#undef B2W_OLA
  int rc = check_launch("overlap_add_kernel");
  if (rc) return rc;
  if (deemphasis != 0.0) {
    deemphasis_kernel<<<(num_utts + 63) / 64, 64, 0, st>>>(utt_out_offset, num_utts, deemphasis, (double*)y);
    return check_launch("deemphasis_kernel");
  }
  return 0;
}

extern "C" int b2w_synth_overlap_add(const double* response, const int64_t* utt_out_offset, const int64_t* utt_pulse_offset,
                                     const int32_t* num_pulses, int32_t num_utts, const int32_t* pulse_index, int32_t fft_size,
                                     int64_t max_out_per_utt, double deemphasis, void* y, int32_t y_dtype, void* stream) {
  return overlap_add_launch(response, 0, utt_out_offset, utt_pulse_offset, num_pulses, num_utts, pulse_index, fft_size,
                            max_out_per_utt, deemphasis, y, y_dtype, stream);
}

extern "C" int b2w_synth_overlap_add_f32(const float* response, const int64_t* utt_out_offset, const int64_t* utt_pulse_offset,
                                         const int32_t* num_pulses, int32_t num_utts, const int32_t* pulse_index, int32_t fft_size,
                                         int64_t max_out_per_utt, double deemphasis, void* y, int32_t y_dtype, void* stream) {
  return overlap_add_launch(response, 1, utt_out_offset, utt_pulse_offset, num_pulses, num_utts, pulse_index, fft_size,
                            max_out_per_utt, deemphasis, y, y_dtype, stream);
}

namespace b2w {
int render_fast_launch(const void* sp, const void* ap, int plane_dtype, const int64_t* utt_frame_offset,
                       const int64_t* utt_pulse_offset, const int* num_pulses, int num_utts, const int* pulse_index,
                       const double* pulse_shift, const uint8_t* pulse_vuv, const double* randn_table, int64_t randn_len, int fs,
                       double frame_period_ms, int64_t total_rows, float* response, cudaStream_t st);  // synth_fast.cu
}

// The batched fast path's response kernel: one warp per pulse, single-precision transforms, float32 responses [total_rows, 1024]
// (total_rows = utt_pulse_offset[num_utts] on the host).  fft_size must be 1024 (fs <= 32 kHz); other sizes: b2w_synth_render.
extern "C" int b2w_synth_render_f32(const void* sp, const void* ap, int32_t plane_dtype, const int64_t* utt_frame_offset,
                                    const int64_t* utt_pulse_offset, const int32_t* num_pulses, int32_t num_utts,
                                    const int32_t* pulse_index, const double* pulse_shift, const uint8_t* pulse_vuv,
                                    const double* randn_table, int64_t randn_table_len, int32_t fs, double frame_period_ms,
                                    int32_t fft_size, int64_t total_rows, float* response, void* stream) {
  using namespace b2w;
  B2W_REQUIRE(sp && ap && utt_frame_offset && utt_pulse_offset && num_pulses && pulse_index && pulse_shift && pulse_vuv &&
                  randn_table && response,
              "b2w_synth_render_f32: null argument");
  B2W_REQUIRE(plane_dtype == B2W_F64 || plane_dtype == B2W_F32, "b2w_synth_render_f32: bad plane dtype %d", plane_dtype);
  B2W_REQUIRE(fft_size == 1024, "b2w_synth_render_f32: fft_size %d (only 1024; use b2w_synth_render)", fft_size);
  if (num_utts == 0 || total_rows == 0) return 0;
  return render_fast_launch(sp, ap, plane_dtype, utt_frame_offset, utt_pulse_offset, num_pulses, num_utts, pulse_index, pulse_shift,
                            pulse_vuv, randn_table, randn_table_len, fs, frame_period_ms, total_rows, response, (cudaStream_t)stream);
}
