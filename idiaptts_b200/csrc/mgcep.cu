// Mel-GENERALISED cepstral analysis and its synthesis-side inverse (sp_type = "mgc", gamma = -1/3).
//
// Replaces pysptk.mgcep(amp_sp, order, alpha, gamma, eps=1e-8, etype=1, itype=3) called through AudioProcessing.extract_mgc
// (idiaptts/src/data_preparation/audio/AudioProcessing.py:123-140) and exp(Re pysptk.mgc2sp(mgc, alpha, gamma, fftlen)) in
// AudioProcessing.mgc_to_amp_sp (:259-275).  PARITY UNPINNED: SPTK is not available and the reference holds no fixture for this
// branch; the kernels follow oracle/mgc_np.py, a restatement of the published definition (Tokuda et al. 1994):
//     H(w) = (1 + gamma C(w))^(1/gamma),  C(w) = sum_m c(m) exp(-j m w~(w)),  w~ = the all-pass warped frequency
//     c = argmin E,  E = mean_w [ I(w) / |H(w)|^2 + log |H(w)|^2 ]              (I = periodogram; convex for -1 <= gamma <= 0)
// solved by an exact Newton iteration whose Hessian is Toeplitz + Hankel, like SPTK's:
//     G = 1 + gamma C,  P = I |G|^(-2/gamma),  q = 1 / G
//     grad(m)  = mean 2 (1 - P) Re(e^{-j m w~} q)
//     H(m, n)  = t(|m - n|) + h(m + n),  t(k) = mean 2 P |q|^2 cos(k w~),  h(k) = mean (2 P - 2 gamma (1 - P)) Re(e^{-j k w~} q^2)
// Every mean over frequency is a contraction of a per-bin weight tile against a constant cos / sin table, so the kernel has the
// shape of mcep.cu: a CTA owns a tile of F frames, two forward contractions (C), five reductions (grad, t, h), one warp-level
// LDL^T solve per frame and iteration (mcep_solve.cuh), SPTK's stopping rule on epsilon = mean(P) exp(mean log |H|^2).
#include "common.cuh"
#include "mcep_solve.cuh"
#include "mcep_tile.cuh"

namespace b2w {

constexpr int kMgF = 8;  // frames per CTA (five weight tiles of [F][K] floats live in shared memory)

struct MgcepParams {
  const void* in;
  int in_is_power;
  int64_t num_frames;
  int K, KP, m, MP, NP0, NP2, NBk, chol_floats, u_floats;
  int miniter, maxiter;
  float threshold, eps, gamma;
  const float* m0t;      // [K][NP0]: log periodogram -> linear mel-cepstrum (+ start value), shared with mcep
  const float* fwd_cos;  // [MP][K]   cos(m w~_j)
  const float* fwd_sin;  // [MP][K]   sin(m w~_j)
  const float* red_cos;  // [K][NP2]  W_j cos(n w~_j), W = bin weights of the mean over the circle
  const float* red_sin;  // [K][NP2]  W_j sin(n w~_j)
  void* out;
  int out_dtype;
  int64_t out_stride;
  int* iters;
  int* status;
};

template <typename IT>
__device__ __forceinline__ float mg_load_per(const MgcepParams& p, int64_t frame, int j) {
  const double v = (double)reinterpret_cast<const IT*>(p.in)[frame * p.K + j];
  return (float)(p.in_is_power ? v + (double)p.eps : v * v + (double)p.eps);
}

// per-bin quantities of one (frame, bin): the five reduction weights and the two terms of epsilon
struct BinOut {
  float a_re, a_im, tau, b_re, b_im, P, logH2;
};
__device__ __forceinline__ BinOut mg_bin(float per, float Cre, float Cim, float gamma) {
  BinOut o;
  const float Gre = fmaf(gamma, Cre, 1.0f), Gim = gamma * Cim;
  const float g2 = fmaf(Gre, Gre, Gim * Gim);
  const float lg = logf(g2);
  o.logH2 = lg / gamma;
  o.P = per * expf(-o.logH2);
  const float inv = 1.0f / g2;
  const float qre = Gre * inv, qim = -Gim * inv;  // q = conj(G) / |G|^2
  const float a = 2.0f * (1.0f - o.P);
  o.a_re = a * qre;
  o.a_im = a * qim;
  o.tau = 2.0f * o.P * inv;
  const float beta = 2.0f * o.P - 2.0f * gamma * (1.0f - o.P);
  o.b_re = beta * (qre * qre - qim * qim);
  o.b_im = beta * (2.0f * qre * qim);
  return o;
}

template <typename IT>
__global__ void __launch_bounds__(kMcThreads) mgcep_kernel(MgcepParams p) {
  constexpr int F = kMgF;
  extern __shared__ float smf[];
  const int K = p.K, KP = p.KP, MP = p.MP, NP2 = p.NP2, m = p.m;
  float* U = smf;                          // five weight tiles [5][F][KP]; aliased by the per-warp solve workspaces
  float* c = U + p.u_floats;               // [F][MP] coefficients
  float* gr = c + F * MP;                  // [F][MP] gradient (also: scratch of the initial gc2gc)
  float* tt = gr + F * MP;                 // [F][MP] Toeplitz sequence
  float* hh = tt + F * MP;                 // [F][NP2] Hankel sequence (also: output of the initial-value contraction)
  float* sP = hh + F * NP2;                // [F] sum W P
  float* sL = sP + F;                      // [F] sum W log |H|^2
  float* prev = sL + F;                    // [F] previous epsilon (< 0: none yet)
  int* act = reinterpret_cast<int*>(prev + F);
  int* itc = act + F;
  uint16_t* tri = reinterpret_cast<uint16_t*>(itc + F);
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t frame0 = (int64_t)blockIdx.x * F;
  const int nvalid = (int)min((int64_t)F, p.num_frames - frame0);
  const float gamma = p.gamma;
  const int tile = F * KP;

  for (int pi = tid; pi < p.NBk * (p.NBk - 1) / 2; pi += kMcThreads) {
    int a_ = 0, q = pi;
    while (q > a_) { q -= a_ + 1; ++a_; }
    tri[pi] = (uint16_t)((a_ << 8) | q);
  }
  for (int i = tid; i < F * MP; i += kMcThreads) c[i] = 0.f;
  // ---- initial value (SPTK): linear mel-cepstrum of the log periodogram ... ------------------------------------------------
  bool zero_per = false;
  for (int f = 0; f < F; ++f) {
    for (int j = tid; j < K; j += kMcThreads) {
      float per = 1.f;
      if (f < nvalid) per = mg_load_per<IT>(p, frame0 + f, j);
      if (!(per > 0.f)) zero_per = true;
      U[f * KP + j] = logf(per);
    }
  }
  if (zero_per) atomicOr(p.status, B2W_STATUS_ZERO_PERIODOGRAM);
  __syncthreads();
  gemm_tile_dispatch<F>(U, KP, K, p.m0t, p.NP0, m + 1, hh, NP2);
  __syncthreads();
  // ... converted to gamma with the gnorm / gc2gc / ignorm recursions (mgc2mgc with equal alpha), one thread per frame
  if (tid < F) {
    const float* c1 = hh + tid * NP2;      // gamma 0 coefficients; gnorm(0): K = exp(c1[0]), c1[1..] unchanged
    float* c2 = gr + tid * MP;             // normalised coefficients at gamma
    for (int i = 1; i <= m; ++i) {
      float ss2 = 0.f;                     // g1 = 0: only the g2 term of gc2gc remains
      for (int k = 1; k <= i - 1; ++k) ss2 = fmaf((float)k * c1[k], c2[i - k], ss2);
      c2[i] = c1[i] + gamma * ss2 / (float)i;
    }
    const float kg = expf(gamma * c1[0]);  // ignorm: K^gamma
    float* cf = c + tid * MP;
    cf[0] = (kg - 1.0f) / gamma;
    for (int i = 1; i <= m; ++i) cf[i] = c2[i] * kg;
    act[tid] = tid < nvalid ? 1 : 0;
    itc[tid] = 0;
    prev[tid] = -1.f;
  }
  __syncthreads();

  const float wmid = 2.0f / (float)(2 * (K - 1)), wend = 1.0f / (float)(2 * (K - 1));
  for (int it = 1; it <= p.maxiter; ++it) {
    if (tid < F) { sP[tid] = 0.f; sL[tid] = 0.f; }
    __syncthreads();
    // ---- forward: C = c . (cos - j sin), then the per-bin weights ---------------------------------------------------------
    {
      float cr0[F], cr1[F], ci0[F], ci1[F];
      float accP[F], accL[F];
#pragma unroll
      for (int f = 0; f < F; ++f) { accP[f] = 0.f; accL[f] = 0.f; }
      for (int j0 = tid; j0 + 1 < K; j0 += 2 * kMcThreads) {  // K - 1 is a multiple of 256 for every supported fft size
        const int j1 = j0 + kMcThreads;
        const bool v1 = j1 + 1 < K;
        const int j1s = v1 ? j1 : j0;
        two_columns<F>(c, MP, p.fwd_cos, K, j0, j1s, 0, cr0, cr1);
        two_columns<F>(c, MP, p.fwd_sin, K, j0, j1s, 0, ci0, ci1);
        const float w0 = j0 == 0 ? wend : wmid;
#pragma unroll
        for (int f = 0; f < F; ++f) {
          float per0 = 1.f, per1 = 1.f;
          if (f < nvalid) {
            per0 = mg_load_per<IT>(p, frame0 + f, j0);
            per1 = mg_load_per<IT>(p, frame0 + f, j1s);
          }
          const BinOut o0 = mg_bin(per0, cr0[f], -ci0[f], gamma);
          float* u = U + f * KP + j0;
          u[0] = o0.a_re; u[tile] = o0.a_im; u[2 * tile] = o0.tau; u[3 * tile] = o0.b_re; u[4 * tile] = o0.b_im;
          accP[f] = fmaf(w0, o0.P, accP[f]);
          accL[f] = fmaf(w0, o0.logH2, accL[f]);
          if (v1) {
            const BinOut o1 = mg_bin(per1, cr1[f], -ci1[f], gamma);
            float* u1 = U + f * KP + j1;
            u1[0] = o1.a_re; u1[tile] = o1.a_im; u1[2 * tile] = o1.tau; u1[3 * tile] = o1.b_re; u1[4 * tile] = o1.b_im;
            accP[f] = fmaf(wmid, o1.P, accP[f]);
            accL[f] = fmaf(wmid, o1.logH2, accL[f]);
          }
        }
      }
      if (tid < F) {  // the Nyquist column, one frame per thread
        const int j = K - 1;
        float cr = 0.f, ci = 0.f;
        for (int k = 0; k <= m; ++k) {
          cr = fmaf(c[tid * MP + k], __ldg(p.fwd_cos + (int64_t)k * K + j), cr);
          ci = fmaf(c[tid * MP + k], __ldg(p.fwd_sin + (int64_t)k * K + j), ci);
        }
        float per = 1.f;
        if (tid < nvalid) per = mg_load_per<IT>(p, frame0 + tid, j);
        const BinOut o = mg_bin(per, cr, -ci, gamma);
        float* u = U + tid * KP + j;
        u[0] = o.a_re; u[tile] = o.a_im; u[2 * tile] = o.tau; u[3 * tile] = o.b_re; u[4 * tile] = o.b_im;
        atomicAdd(&sP[tid], wend * o.P);
        atomicAdd(&sL[tid], wend * o.logH2);
      }
#pragma unroll
      for (int f = 0; f < F; ++f) {
        float a = accP[f], b = accL[f];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (lane == 0) {
          atomicAdd(&sP[f], a);
          atomicAdd(&sL[f], b);
        }
      }
    }
    __syncthreads();
    // ---- reductions over frequency: gradient, Toeplitz and Hankel sequences ------------------------------------------------
    gemm_tile_dispatch<F>(U, KP, K, p.red_cos, NP2, m + 1, gr, MP);
    gemm_tile_dispatch<F>(U + 2 * tile, KP, K, p.red_cos, NP2, m + 1, tt, MP);
    gemm_tile_dispatch<F>(U + 3 * tile, KP, K, p.red_cos, NP2, 2 * m + 1, hh, NP2);
    __syncthreads();
    gemm_tile_dispatch<F, true>(U + tile, KP, K, p.red_sin, NP2, m + 1, gr, MP);
    gemm_tile_dispatch<F, true>(U + 4 * tile, KP, K, p.red_sin, NP2, 2 * m + 1, hh, NP2);
    __syncthreads();
    // ---- SPTK's stopping rule on epsilon ------------------------------------------------------------------------------------
    if (tid < F && act[tid]) {
      const float e = sP[tid] * expf(sL[tid]);
      if (it >= p.miniter && prev[tid] >= 0.f && fabsf((e - prev[tid]) / e) < p.threshold) {
        act[tid] = 0;
        itc[tid] = it;
      } else {
        prev[tid] = e;
      }
    }
    __syncthreads();
    int any = 0;
    if (tid < F) any = act[tid];
    if (!__syncthreads_or(any)) break;
    // ---- Newton step for the frames still active: one warp per frame (the weight tiles are free now) ------------------------
    {
      const int warp = tid >> 5;
      float* ws = U + warp * p.chol_floats;
      for (int f = warp; f < F; f += kMcWarps) {
        if (!act[f]) continue;
        float* xo = ws + p.chol_floats - pad4(m + 1);
        const bool ok = warp_ldl_solve<kBlk>(tt + f * MP, nullptr, m + 1, p.NBk, tri, ws, xo, hh + f * NP2, gr + f * MP);
        if (!ok) {
          if (lane == 0) {
            atomicOr(p.status, B2W_STATUS_SOLVE_FAILED);
            act[f] = 0;
            itc[f] = it;
          }
        } else {
          for (int k = lane; k <= m; k += 32) c[f * MP + k] -= xo[k];
        }
        __syncwarp();
      }
    }
    __syncthreads();
  }
  if (tid < F && act[tid]) {
    itc[tid] = p.maxiter;
    atomicOr(p.status, B2W_STATUS_NOT_CONVERGED);
  }
  __syncthreads();
  for (int i = tid; i < nvalid * (m + 1); i += kMcThreads) {
    const int f = i / (m + 1), k = i - f * (m + 1);
    const float v = c[f * MP + k];
    if (p.out_dtype == B2W_F64) reinterpret_cast<double*>(p.out)[(frame0 + f) * p.out_stride + k] = (double)v;
    else reinterpret_cast<float*>(p.out)[(frame0 + f) * p.out_stride + k] = v;
  }
  if (p.iters && tid < nvalid) p.iters[frame0 + tid] = itc[tid];
}

// |H(w_j)| = |1 + gamma sum_m c(m) e^{-j m w~_j}|^(1/gamma): AudioProcessing.mgc_to_amp_sp without the detour through a
// 512-term cepstrum (pysptk.mgc2sp = mgc2mgc to gamma 0 + FFT; the two agree to 3e-8 relative, oracle/mgc_np.py)
template <typename MT, typename OT>
__global__ void __launch_bounds__(kMcThreads) mgc2sp_kernel(const MT* __restrict__ mgc, int64_t stride, int64_t num_frames, int K,
                                                            int m, float gamma, const float* __restrict__ fwd_cos,
                                                            const float* __restrict__ fwd_sin, OT* __restrict__ out) {
  constexpr int F = 16;
  extern __shared__ float smf[];
  const int MP = pad4(m + 1);
  float* c = smf;
  const int tid = threadIdx.x;
  const int64_t frame0 = (int64_t)blockIdx.x * F;
  const int nvalid = (int)min((int64_t)F, num_frames - frame0);
  for (int i = tid; i < F * MP; i += kMcThreads) {
    const int f = i / MP, k = i - f * MP;
    c[i] = (f < nvalid && k <= m) ? (float)mgc[(frame0 + f) * stride + k] : 0.f;
  }
  __syncthreads();
  const float inv_2g = 0.5f / gamma;
  for (int j = tid; j < K; j += kMcThreads) {
    float cr[F], ci[F];
#pragma unroll
    for (int f = 0; f < F; ++f) { cr[f] = 0.f; ci[f] = 0.f; }
    for (int k = 0; k <= m; ++k) {
      const float cs = __ldg(fwd_cos + (int64_t)k * K + j), sn = __ldg(fwd_sin + (int64_t)k * K + j);
#pragma unroll
      for (int f = 0; f < F; ++f) {
        const float v = c[f * MP + k];
        cr[f] = fmaf(v, cs, cr[f]);
        ci[f] = fmaf(v, sn, ci[f]);
      }
    }
#pragma unroll
    for (int f = 0; f < F; ++f) {
      if (f < nvalid) {
        const float Gre = fmaf(gamma, cr[f], 1.0f), Gim = gamma * ci[f];
        out[(frame0 + f) * K + j] = (OT)expf(inv_2g * logf(fmaf(Gre, Gre, Gim * Gim)));
      }
    }
  }
}

}  // namespace b2w

extern "C" int b2w_mgcep(const void* in, int32_t in_dtype, int32_t in_is_power, int64_t num_frames, int32_t fft_size, int32_t order,
                         double gamma, int32_t miniter, int32_t maxiter, double threshold, double eps, const float* m0t,
                         const float* fwd_cos, const float* fwd_sin, const float* red_cos, const float* red_sin, void* mgc,
                         int32_t mgc_dtype, int64_t mgc_stride, int32_t* iters, int32_t* status, void* stream) {
  using namespace b2w;
  B2W_REQUIRE(in && m0t && fwd_cos && fwd_sin && red_cos && red_sin && mgc && status, "b2w_mgcep: null argument");
  B2W_REQUIRE(in_dtype == B2W_F64 || in_dtype == B2W_F32, "b2w_mgcep: bad in_dtype %d", in_dtype);
  B2W_REQUIRE(mgc_dtype == B2W_F64 || mgc_dtype == B2W_F32, "b2w_mgcep: bad mgc_dtype %d", mgc_dtype);
  B2W_REQUIRE(fft_size == 512 || fft_size == 1024 || fft_size == 2048 || fft_size == 4096, "b2w_mgcep: unsupported fft_size %d",
              fft_size);
  B2W_REQUIRE(order >= 1 && 2 * order + 1 <= 256, "b2w_mgcep: order %d out of range [1, 127]", order);
  B2W_REQUIRE(gamma < 0.0 && gamma >= -1.0, "b2w_mgcep: gamma %g outside [-1, 0) (gamma = 0 is b2w_mcep)", gamma);
  B2W_REQUIRE(mgc_stride >= order + 1, "b2w_mgcep: mgc_stride too small");
  B2W_REQUIRE(maxiter >= 1 && miniter >= 1, "b2w_mgcep: bad iteration limits");
  if (num_frames == 0) return 0;
  MgcepParams p;
  p.in = in; p.in_is_power = in_is_power; p.num_frames = num_frames;
  p.K = fft_size / 2 + 1; p.KP = pad4(p.K); p.m = order; p.MP = pad4(order + 1);
  p.NP0 = pad4(order + 2); p.NP2 = pad4(2 * order + 1);
  p.NBk = (order + 1 + 3) / 4;
  p.chol_floats = (p.NBk * (p.NBk + 1) / 2) * kBlk + p.NBk * kBlk + 8 * p.NBk + pad4(order + 1);
  p.miniter = miniter; p.maxiter = maxiter; p.threshold = (float)threshold; p.eps = (float)eps; p.gamma = (float)gamma;
  p.m0t = m0t; p.fwd_cos = fwd_cos; p.fwd_sin = fwd_sin; p.red_cos = red_cos; p.red_sin = red_sin;
  p.out = mgc; p.out_dtype = mgc_dtype; p.out_stride = mgc_stride; p.iters = iters; p.status = status;
  const int F = kMgF;
  const int tiles = 5 * F * p.KP, ws = kMcWarps * p.chol_floats;
  p.u_floats = tiles > ws ? tiles : ws;
  const size_t smem = sizeof(float) * ((size_t)p.u_floats + 3 * (size_t)F * p.MP + (size_t)F * p.NP2 + 3 * F) + sizeof(int) * 2 * F +
                      sizeof(uint16_t) * (size_t)(p.NBk * (p.NBk - 1) / 2 + 2);
  B2W_REQUIRE(smem <= 227 * 1024, "b2w_mgcep: order %d / fft_size %d need %zu bytes of shared memory", order, fft_size, smem);
  const int64_t grid = (num_frames + F - 1) / F;
  B2W_REQUIRE(grid < (int64_t)1 << 31, "b2w_mgcep: too many frames in one call");
  cudaStream_t st = (cudaStream_t)stream;
  if (in_dtype == B2W_F64) {
    cudaFuncSetAttribute(mgcep_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    mgcep_kernel<double><<<(unsigned)grid, kMcThreads, smem, st>>>(p);
  } else {
    cudaFuncSetAttribute(mgcep_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    mgcep_kernel<float><<<(unsigned)grid, kMcThreads, smem, st>>>(p);
  }
  return check_launch("mgcep_kernel");
}

extern "C" int b2w_mgc2sp(const void* mgc, int32_t mgc_dtype, int64_t mgc_stride, int64_t num_frames, int32_t fft_size, int32_t order,
                          double gamma, const float* fwd_cos, const float* fwd_sin, void* amp, int32_t amp_dtype, void* stream) {
  using namespace b2w;
  B2W_REQUIRE(mgc && fwd_cos && fwd_sin && amp, "b2w_mgc2sp: null argument");
  B2W_REQUIRE(mgc_dtype == B2W_F64 || mgc_dtype == B2W_F32, "b2w_mgc2sp: bad mgc_dtype %d", mgc_dtype);
  B2W_REQUIRE(amp_dtype == B2W_F64 || amp_dtype == B2W_F32, "b2w_mgc2sp: bad amp_dtype %d", amp_dtype);
  B2W_REQUIRE(order >= 0 && order < 512, "b2w_mgc2sp: bad order %d", order);
  B2W_REQUIRE(gamma != 0.0, "b2w_mgc2sp: gamma = 0 is b2w_mc2sp");
  if (num_frames == 0) return 0;
  const int K = fft_size / 2 + 1;
  const size_t smem = sizeof(float) * 16 * pad4(order + 1);
  const int64_t grid = (num_frames + 15) / 16;
  B2W_REQUIRE(grid < (int64_t)1 << 31, "b2w_mgc2sp: too many frames in one call");
  cudaStream_t st = (cudaStream_t)stream;
#define B2W_MGC2SP(MT, OT)                                                                                                      \
  mgc2sp_kernel<MT, OT><<<(unsigned)grid, kMcThreads, smem, st>>>((const MT*)mgc, mgc_stride, num_frames, K, order, (float)gamma, \
                                                                  fwd_cos, fwd_sin, (OT*)amp)
  if (mgc_dtype == B2W_F64) {
    if (amp_dtype == B2W_F64) B2W_MGC2SP(double, double); else B2W_MGC2SP(double, float);
  } else {
    if (amp_dtype == B2W_F64) B2W_MGC2SP(float, double); else B2W_MGC2SP(float, float);
  }
#undef B2W_MGC2SP
  return check_launch("mgc2sp_kernel");
}
