// Mel-cepstral analysis (SPTK mcep, Newton-Raphson UELS) and synthesis-side mc -> spectrum, as dense contractions
// against precomputed all-pass warping matrices.
//
// Replaces pysptk.mcep(amp_sp, order, alpha, eps=1e-8, etype=1, itype=3) called once per frame through
// AudioProcessing.extract_mcep (idiaptts/src/data_preparation/audio/AudioProcessing.py:143-153) and
// pysptk.mgc2sp(..., gamma=0) in mcep_to_amp_sp (:248-256).
//
// SPTK runs, per frame and per Newton iteration: freqt(mc -> N/2, -alpha), FFT, exp, IFFT, frqtr(-> 2m, alpha), and a
// (Toeplitz + Hankel) solve.  freqt/frqtr are linear maps with constant coefficients and so are the real (I)FFTs, so
// each pair collapses into one constant matrix (built in fp64 on the host by b2w_mcep_tables_host):
//     C   = mc  . Cmat            [F x (m+1)] x [(m+1) x K]          (freqt + FFT)
//     P   = per * exp(-2 C)                                           (elementwise)
//     r~  = P   . M2^T            [F x K] x [K x (2m+1)]              (IFFT + frqtr)
//     mc0 = log(per) . M0^T       [F x K] x [K x (m+2)]               (initial value, + start value s)
// followed by a per-frame (m+1) x (m+1) SPD solve (blocked LDL^T, 4x4 register tiles, one warp per frame).
// A CTA owns a tile of F frames and iterates until every frame of the tile has met SPTK's stopping rule.
//
// This file holds the fp32 CUDA-core version of the contractions (v1).
#include <vector>

#include "common.cuh"
#include "mcep_solve.cuh"
#include "mcep_tile.cuh"

namespace b2w {

struct McepParams {
  const void* in;
  int in_is_power;
  int64_t num_frames;
  int K;       // fft_size/2 + 1
  int KP;      // padded row length of the smem tile (multiple of 4)
  int m;       // order
  int MP;      // pad4(m + 1)
  int NP0;     // pad4(m + 2): row stride of m0t
  int NP2;     // pad4(2m + 1): row stride of m2t, and of the rt tile
  int NBk;     // number of 4-wide blocks of the solve: ceil((m+1)/4)
  int chol_floats;  // per-warp solve workspace in floats
  int u_floats;     // size of the shared tile/workspace union in floats
  int miniter, maxiter;
  float threshold, eps;
  float alpha;
  const float* m0t;
  const float* cmat;
  const float* m2t;
  void* mc_out;
  int mc_dtype;
  int64_t mc_stride;
  int* iters;
  int* status;
};

template <typename IT>
__device__ __forceinline__ float load_per(const McepParams& p, int64_t frame, int j) {
  const double v = (double)reinterpret_cast<const IT*>(p.in)[frame * p.K + j];
  return (float)(p.in_is_power ? v + (double)p.eps : v * v + (double)p.eps);
}

template <int F, typename IT>
__global__ void __launch_bounds__(kMcThreads) mcep_kernel(McepParams p) {
  extern __shared__ float smf[];
  float* U = smf;                         // tile [F][KP] (log per / P), later aliased by the per-warp solve workspaces
  float* mc = U + p.u_floats;             // [F][MP]
  float* rt = mc + F * p.MP;              // [F][NP2]
  float* al = rt + F * p.NP2;             // [MP]
  float* sv = al + p.MP;                  // [F] start / previous r~[0]
  int* act = reinterpret_cast<int*>(sv + F);  // [F] 1 = still iterating
  int* itc = act + F;                         // [F] iteration count at exit
  uint16_t* tri = reinterpret_cast<uint16_t*>(itc + F);  // [NBk (NBk-1) / 2] flat pair index -> (a << 8 | q), q <= a
  const int tid = threadIdx.x;
  const int K = p.K, KP = p.KP, MP = p.MP, m = p.m;
  const int64_t frame0 = (int64_t)blockIdx.x * F;
  const int nvalid = (int)min((int64_t)F, p.num_frames - frame0);

  for (int pi = tid; pi < p.NBk * (p.NBk - 1) / 2; pi += kMcThreads) {
    int a_ = 0, q = pi;
    while (q > a_) { q -= a_ + 1; ++a_; }
    tri[pi] = (uint16_t)((a_ << 8) | q);
  }
  // (-alpha)^k, zero padded
  if (tid < MP) al[tid] = (tid <= m) ? powf(-p.alpha, (float)tid) : 0.f;
  if (tid == 0) al[0] = 1.f;
  for (int i = tid; i < F * MP; i += kMcThreads) mc[i] = 0.f;
  // log periodogram tile
  bool zero_per = false;
  for (int f = 0; f < F; ++f) {
    for (int j = tid; j < K; j += kMcThreads) {
      float per = 1.f;
      if (f < nvalid) per = load_per<IT>(p, frame0 + f, j);
      if (!(per > 0.f)) zero_per = true;
      U[f * KP + j] = logf(per);
    }
  }
  if (zero_per) atomicOr(p.status, B2W_STATUS_ZERO_PERIODOGRAM);
  __syncthreads();
  // initial value: [mc | s] = log(per) . M0^T
  gemm_tile_dispatch<F>(U, KP, K, p.m0t, p.NP0, m + 2, rt, p.NP2);
  __syncthreads();
  for (int i = tid; i < F * (m + 1); i += kMcThreads) {
    const int f = i / (m + 1), k = i - f * (m + 1);
    mc[f * MP + k] = rt[f * p.NP2 + k];
  }
  if (tid < F) {
    sv[tid] = rt[tid * p.NP2 + m + 1];
    act[tid] = tid < nvalid ? 1 : 0;
    itc[tid] = 0;
  }
  __syncthreads();

  for (int it = 1; it <= p.maxiter; ++it) {
    // P = per * exp(-2 mc . Cmat): every thread owns the two columns tid and tid + 256 (and, for K - 1 = 1024 or 2048,
    // further pairs), half of the tile's frames at a time
    {
      constexpr int FH = F >= 16 ? 16 : F;
      float acc0[FH], acc1[FH];
      for (int j0 = tid; j0 + 1 < K; j0 += 2 * kMcThreads) {  // K - 1 is a multiple of 256 for every supported fft size
        const int j1 = j0 + kMcThreads;
        const bool v1 = j1 + 1 < K;
        const int j1s = v1 ? j1 : j0;
        for (int fb = 0; fb < F; fb += FH) {
          two_columns<FH>(mc, MP, p.cmat, K, j0, j1s, fb, acc0, acc1);
#pragma unroll
          for (int f = 0; f < FH; ++f) {
            float per0 = 1.f, per1 = 1.f;
            if (fb + f < nvalid) {
              per0 = load_per<IT>(p, frame0 + fb + f, j0);
              per1 = load_per<IT>(p, frame0 + fb + f, j1s);
            }
            U[(fb + f) * KP + j0] = per0 * expf(-2.f * acc0[f]);
            if (v1) U[(fb + f) * KP + j1] = per1 * expf(-2.f * acc1[f]);
          }
        }
      }
      if (tid < F) {  // the Nyquist column, one frame per thread
        float c = 0.f;
        for (int k = 0; k <= m; ++k) c = fmaf(mc[tid * MP + k], __ldg(p.cmat + (int64_t)k * K + (K - 1)), c);
        float per = 1.f;
        if (tid < nvalid) per = load_per<IT>(p, frame0 + tid, K - 1);
        U[tid * KP + K - 1] = per * expf(-2.f * c);
      }
    }
    __syncthreads();
    gemm_tile_dispatch<F>(U, KP, K, p.m2t, p.NP2, 2 * m + 1, rt, p.NP2);
    __syncthreads();
    // SPTK's stopping rule on r~[0]
    if (tid < F && act[tid]) {
      const float t = rt[tid * p.NP2];
      if (it >= p.miniter) {
        if (fabsf((t - sv[tid]) / t) < p.threshold) {
          act[tid] = 0;
          itc[tid] = it;
        } else {
          sv[tid] = t;
        }
      }
    }
    __syncthreads();
    int any = 0;
    if (tid < F) any = act[tid];
    if (!__syncthreads_or(any)) break;
    // Newton step for the frames still active: one warp per frame (the tile U is free now and holds the workspaces)
    {
      const int warp = tid >> 5, lane = tid & 31;
      float* ws = U + warp * p.chol_floats;
      for (int f = warp; f < F; f += kMcWarps) {
        if (!act[f]) continue;
        float* xo = ws + p.chol_floats - pad4(m + 1);  // tail of the workspace
        const bool ok = warp_ldl_solve<kBlk>(rt + f * p.NP2, al, m + 1, p.NBk, tri, ws, xo);
        if (!ok) {
          if (lane == 0) {
            atomicOr(p.status, B2W_STATUS_SOLVE_FAILED);
            act[f] = 0;
            itc[f] = it;
          }
        } else {
          for (int k = lane; k <= m; k += 32) mc[f * MP + k] += xo[k];
        }
        __syncwarp();
      }
    }
    __syncthreads();
  }
  // frames that ran out of iterations
  if (tid < F && act[tid]) {
    itc[tid] = p.maxiter;
    atomicOr(p.status, B2W_STATUS_NOT_CONVERGED);
  }
  __syncthreads();
  for (int i = tid; i < nvalid * (m + 1); i += kMcThreads) {
    const int f = i / (m + 1), k = i - f * (m + 1);
    const float v = mc[f * MP + k];
    if (p.mc_dtype == B2W_F64) reinterpret_cast<double*>(p.mc_out)[(frame0 + f) * p.mc_stride + k] = (double)v;
    else reinterpret_cast<float*>(p.mc_out)[(frame0 + f) * p.mc_stride + k] = v;
  }
  if (p.iters && tid < nvalid) p.iters[frame0 + tid] = itc[tid];
}

// out[f][j] = (exp?)(scale * sum_k mc[f][k] cmat[k][j])
template <int F, typename MT, typename OT>
__global__ void __launch_bounds__(kMcThreads) mc2sp_kernel(const MT* __restrict__ mcg, int64_t mc_stride, int64_t num_frames,
                                                           int K, int m, const float* __restrict__ cmat, float scale,
                                                           int do_exp, OT* __restrict__ out) {
  extern __shared__ float smf[];
  const int MP = pad4(m + 1);
  float* mc = smf;  // [F][MP]
  const int tid = threadIdx.x;
  const int64_t frame0 = (int64_t)blockIdx.x * F;
  const int nvalid = (int)min((int64_t)F, num_frames - frame0);
  for (int i = tid; i < F * MP; i += kMcThreads) {
    const int f = i / MP, k = i - f * MP;
    mc[i] = (f < nvalid && k <= m) ? (float)mcg[(frame0 + f) * mc_stride + k] : 0.f;
  }
  __syncthreads();
  auto emit = [&](int f, int j, float accv) {
    const float v = scale * accv;
    // do_exp 2: the power spectrum as world_features_to_raw builds it (W:924): float32 amplitude, squared in float64
    const float a = do_exp ? expf(v) : v;
    out[(frame0 + f) * K + j] = (do_exp == 2) ? (OT)((double)a * (double)a) : (OT)a;
  };
  // Two columns per thread (j and j + 256) share every 16-byte broadcast load of the coefficients; K = 2^p + 1, so after the
  // column pairs one Nyquist column is left, which is spread over the threads as (frame, column) dot products instead of
  // costing a third, almost empty pass (513 columns on 256 threads used to run at 67 % thread utilisation).
  int jdone = 0;
  for (int jb = 0; jb + 2 * kMcThreads <= K; jb += 2 * kMcThreads) {
    const int ja = jb + tid, jc = jb + kMcThreads + tid;
    float acc0[F], acc1[F];
#pragma unroll
    for (int f = 0; f < F; ++f) acc0[f] = acc1[f] = 0.f;
    for (int k = 0; k < MP; k += 4) {
      float wa[4], wc[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        wa[q] = (k + q <= m) ? __ldg(cmat + (int64_t)(k + q) * K + ja) : 0.f;
        wc[q] = (k + q <= m) ? __ldg(cmat + (int64_t)(k + q) * K + jc) : 0.f;
      }
#pragma unroll
      for (int f = 0; f < F; ++f) {
        const float4 c4 = *reinterpret_cast<const float4*>(mc + f * MP + k);
        acc0[f] = fmaf(c4.x, wa[0], acc0[f]);
        acc0[f] = fmaf(c4.y, wa[1], acc0[f]);
        acc0[f] = fmaf(c4.z, wa[2], acc0[f]);
        acc0[f] = fmaf(c4.w, wa[3], acc0[f]);
        acc1[f] = fmaf(c4.x, wc[0], acc1[f]);
        acc1[f] = fmaf(c4.y, wc[1], acc1[f]);
        acc1[f] = fmaf(c4.z, wc[2], acc1[f]);
        acc1[f] = fmaf(c4.w, wc[3], acc1[f]);
      }
    }
#pragma unroll
    for (int f = 0; f < F; ++f) {
      if (f < nvalid) {
        emit(f, ja, acc0[f]);
        emit(f, jc, acc1[f]);
      }
    }
    jdone = jb + 2 * kMcThreads;
  }
  // remaining columns: one (frame, column) dot product per thread, same summation order over k
  for (int e = tid; e < F * (K - jdone); e += kMcThreads) {
    const int f = e % F, j = jdone + e / F;
    if (f >= nvalid) continue;
    float accv = 0.f;
    for (int k = 0; k <= m; ++k) accv = fmaf(mc[f * MP + k], __ldg(cmat + (int64_t)k * K + j), accv);
    emit(f, j, accv);
  }
}

// ---- host: fp64 construction of the warping matrices -------------------------------------------------------------------
// freqt as a matrix: column p of the result is freqt(e_p); vectorised over the unit inputs.
static void freqt_matrix_host(int m1, int m2, double a, bool is_frqtr, std::vector<double>& out /* [(m2+1) x (m1+1)] */) {
  const int nin = m1 + 1, nout = m2 + 1;
  const double b = 1.0 - a * a;
  std::vector<double> g((size_t)nout * nin, 0.0), d((size_t)nout * nin, 0.0);
  for (int i = -m1; i <= 0; ++i) {
    d = g;
    const int src = -i;
    for (int p = 0; p < nin; ++p) {
      const double e = (p == src) ? 1.0 : 0.0;
      g[p] = is_frqtr ? e : e + a * d[p];
    }
    if (nout > 1) {
      if (is_frqtr) {
        for (int p = 0; p < nin; ++p) g[nin + p] = d[p] + a * (d[nin + p] - g[p]);
      } else {
        for (int p = 0; p < nin; ++p) g[nin + p] = b * d[p] + a * d[nin + p];
      }
    }
    for (int j = 2; j < nout; ++j) {
      double* gj = &g[(size_t)j * nin];
      const double* gjm = &g[(size_t)(j - 1) * nin];
      const double* dj = &d[(size_t)j * nin];
      const double* djm = &d[(size_t)(j - 1) * nin];
      for (int p = 0; p < nin; ++p) gj[p] = djm[p] + a * (dj[p] - gjm[p]);
    }
  }
  out.swap(g);
}

}  // namespace b2w

extern "C" int32_t b2w_mcep_pad(int32_t n) { return b2w::pad4(n); }

extern "C" int b2w_mcep_tables_host(int32_t order, double alpha, int32_t fft_size, double* h_m0t, double* h_cmat,
                                    double* h_m2t) {
  using namespace b2w;
  B2W_REQUIRE(order >= 1 && fft_size >= 8 && (fft_size & (fft_size - 1)) == 0, "b2w_mcep_tables_host: bad order/fft_size");
  B2W_REQUIRE(h_m0t && h_cmat && h_m2t, "b2w_mcep_tables_host: null argument");
  const int N = fft_size, f2 = N / 2, K = f2 + 1, m = order;
  const int np0 = pad4(m + 2), np2 = pad4(2 * m + 1);
  std::vector<double> A, Bm, R;
  freqt_matrix_host(f2, m, alpha, false, A);       // [m+1][K]
  freqt_matrix_host(m, f2, -alpha, false, Bm);     // [K][m+1]
  freqt_matrix_host(f2, 2 * m, alpha, true, R);    // [2m+1][K]
  std::vector<double> cosv(N);
  for (int i = 0; i < N; ++i) cosv[i] = cos(2.0 * kPi * i / N);
  // Fd[n][j] = D[n] * w_j cos(2 pi n j / N) / N   (real IFFT of the mirrored half spectrum, then c0/2, c[N/2]/2)
  std::vector<double> Fm((size_t)K * K);
  for (int n = 0; n < K; ++n)
    for (int j = 0; j < K; ++j) {
      const double w = (j == 0 || j == f2) ? 1.0 : 2.0;
      Fm[(size_t)n * K + j] = w * cosv[(int)(((int64_t)n * j) % N)] / N;
    }
  for (int j = 0; j < K; ++j) {
    double* row = h_m0t + (size_t)j * np0;
    for (int k = 0; k < np0; ++k) row[k] = 0.0;
    for (int k = 0; k <= m; ++k) {
      double s = 0.0;
      for (int n = 0; n < K; ++n) {
        const double dn = (n == 0 || n == f2) ? 0.5 : 1.0;
        s += A[(size_t)k * K + n] * dn * Fm[(size_t)n * K + j];
      }
      row[k] = s;
    }
    row[m + 1] = 0.5 * Fm[j];  // c[0] after halving: SPTK's start value s
  }
  for (int k = 0; k < pad4(m + 1); ++k)
    for (int j = 0; j < K; ++j) {
      double s = 0.0;
      if (k <= m)
        for (int n = 0; n < K; ++n) s += cosv[(int)(((int64_t)n * j) % N)] * Bm[(size_t)n * (m + 1) + k];
      h_cmat[(size_t)k * K + j] = s;
    }
  for (int j = 0; j < K; ++j) {
    double* row = h_m2t + (size_t)j * np2;
    for (int r = 0; r < np2; ++r) row[r] = 0.0;
    for (int r = 0; r <= 2 * m; ++r) {
      double s = 0.0;
      for (int n = 0; n < K; ++n) s += R[(size_t)r * K + n] * Fm[(size_t)n * K + j];
      row[r] = s;
    }
  }
  return 0;
}

extern "C" int b2w_mcep(const void* in, int32_t in_dtype, int32_t in_is_power, int64_t num_frames, int32_t fft_size,
                        int32_t order, double alpha, int32_t miniter, int32_t maxiter, double threshold, double eps,
                        const float* m0t, const float* cmat, const float* m2t, void* mc, int32_t mc_dtype, int64_t mc_stride,
                        int32_t* iters, int32_t* status, void* stream) {
  using namespace b2w;
  B2W_REQUIRE(in && m0t && cmat && m2t && mc && status, "b2w_mcep: null argument");
  B2W_REQUIRE(in_dtype == B2W_F64 || in_dtype == B2W_F32, "b2w_mcep: bad in_dtype %d", in_dtype);
  B2W_REQUIRE(mc_dtype == B2W_F64 || mc_dtype == B2W_F32, "b2w_mcep: bad mc_dtype %d", mc_dtype);
  B2W_REQUIRE(fft_size == 512 || fft_size == 1024 || fft_size == 2048 || fft_size == 4096,
              "b2w_mcep: unsupported fft_size %d", fft_size);
  B2W_REQUIRE(order >= 1 && 2 * order + 1 <= 256, "b2w_mcep: order %d out of range [1, 127]", order);
  B2W_REQUIRE(mc_stride >= order + 1, "b2w_mcep: mc_stride too small");
  B2W_REQUIRE(maxiter >= 1 && miniter >= 1, "b2w_mcep: bad iteration limits");
  if (num_frames == 0) return 0;
  McepParams p;
  p.in = in; p.in_is_power = in_is_power; p.num_frames = num_frames;
  p.K = fft_size / 2 + 1; p.KP = pad4(p.K); p.m = order; p.MP = pad4(order + 1);
  p.NP0 = pad4(order + 2); p.NP2 = pad4(2 * order + 1);
  p.NBk = (order + 1 + 3) / 4;
  p.chol_floats = (p.NBk * (p.NBk + 1) / 2) * kBlk + p.NBk * kBlk + 8 * p.NBk + pad4(order + 1);
  p.miniter = miniter; p.maxiter = maxiter; p.threshold = (float)threshold; p.eps = (float)eps; p.alpha = (float)alpha;
  p.m0t = m0t; p.cmat = cmat; p.m2t = m2t; p.mc_out = mc; p.mc_dtype = mc_dtype; p.mc_stride = mc_stride;
  p.iters = iters; p.status = status;
  cudaStream_t st = (cudaStream_t)stream;
  // tile height by fft size so the tile fits shared memory next to the small per-frame vectors
  const int F = fft_size <= 1024 ? 32 : (fft_size == 2048 ? 16 : 8);
  const int tile_floats = F * p.KP;
  const int ws_floats = kMcWarps * p.chol_floats;
  p.u_floats = tile_floats > ws_floats ? tile_floats : ws_floats;
  const size_t smem = sizeof(float) * ((size_t)p.u_floats + (size_t)F * p.MP + (size_t)F * p.NP2 + p.MP + F) + sizeof(int) * 2 * F +
                      sizeof(uint16_t) * (size_t)(p.NBk * (p.NBk - 1) / 2 + 2);
  B2W_REQUIRE(smem <= 227 * 1024, "b2w_mcep: order %d needs %zu bytes of shared memory", order, smem);
  const int64_t grid = (num_frames + F - 1) / F;
  B2W_REQUIRE(grid < (int64_t)1 << 31, "b2w_mcep: too many frames in one call");
#define B2W_MCEP_LAUNCH(FF, IT)                                                                          \
  do {                                                                                                   \
    cudaFuncSetAttribute(mcep_kernel<FF, IT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    mcep_kernel<FF, IT><<<(unsigned)grid, kMcThreads, smem, st>>>(p);                                    \
  } while (0)
  if (in_dtype == B2W_F64) {
    if (F == 32) B2W_MCEP_LAUNCH(32, double); else if (F == 16) B2W_MCEP_LAUNCH(16, double); else B2W_MCEP_LAUNCH(8, double);
  } else {
    if (F == 32) B2W_MCEP_LAUNCH(32, float); else if (F == 16) B2W_MCEP_LAUNCH(16, float); else B2W_MCEP_LAUNCH(8, float);
  }
#undef B2W_MCEP_LAUNCH
  return check_launch("mcep_kernel");
}

extern "C" int b2w_mc2sp(const void* mc, int32_t mc_dtype, int64_t mc_stride, int64_t num_frames, int32_t fft_size,
                         int32_t order, const float* cmat, double scale, int32_t do_exp, void* out, int32_t out_dtype,
                         void* stream) {
  using namespace b2w;
  B2W_REQUIRE(mc && cmat && out, "b2w_mc2sp: null argument");
  B2W_REQUIRE(mc_dtype == B2W_F64 || mc_dtype == B2W_F32, "b2w_mc2sp: bad mc_dtype %d", mc_dtype);
  B2W_REQUIRE(out_dtype == B2W_F64 || out_dtype == B2W_F32, "b2w_mc2sp: bad out_dtype %d", out_dtype);
  B2W_REQUIRE(order >= 0 && order < 512, "b2w_mc2sp: bad order %d", order);
  if (num_frames == 0) return 0;
  constexpr int F = 16;
  const int K = fft_size / 2 + 1;
  const size_t smem = sizeof(float) * F * pad4(order + 1);
  const int64_t grid = (num_frames + F - 1) / F;
  B2W_REQUIRE(grid < (int64_t)1 << 31, "b2w_mc2sp: too many frames in one call");
  cudaStream_t st = (cudaStream_t)stream;
#define B2W_MC2SP_LAUNCH(MT, OT)                                                                                      \
  mc2sp_kernel<F, MT, OT><<<(unsigned)grid, kMcThreads, smem, st>>>((const MT*)mc, mc_stride, num_frames, K, order, cmat, \
                                                                    (float)scale, do_exp, (OT*)out)
  if (mc_dtype == B2W_F64) {
    if (out_dtype == B2W_F64) B2W_MC2SP_LAUNCH(double, double); else B2W_MC2SP_LAUNCH(double, float);
  } else {
    if (out_dtype == B2W_F64) B2W_MC2SP_LAUNCH(float, double); else B2W_MC2SP_LAUNCH(float, float);
  }
#undef B2W_MC2SP_LAUNCH
  return check_launch("mc2sp_kernel");
}
