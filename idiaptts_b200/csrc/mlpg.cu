// MLPG: maximum-probability parameter generation for every (utterance, feature dimension) of a ragged batch.
//
// Replaces MLPG.generation (idiaptts/misc/mlpg.py:94-127, called from WorldFeatLabelGen._postprocess_world :357-415 on the
// network output [static | delta | delta-delta]): the reference builds, per dimension, the banded precision matrix
// P = sum_w W_w^T diag(tau_w) W_w and b = sum_w W_w^T (mu_w tau_w) with bandmat (:54-92) and solves P x = b by banded Cholesky
// (:125).  With the three fixed windows [1], [-0.5 0 0.5], [1 -2 1] (:95-99) P is pentadiagonal with closed-form entries
//     P[i][i]   = tau0 + 0.25 (tau1[i-1] + tau1[i+1]) + tau2[i-1] + 4 tau2[i] + tau2[i+1]
//     P[i][i+1] = -2 (tau2[i] + tau2[i+1])          P[i][i+2] = tau2[i+1] - 0.25 tau1[i+1]
//     b[i]      = tau0 mu0[i] + 0.5 (tau1 mu1)[i-1] - 0.5 (tau1 mu1)[i+1] + (tau2 mu2)[i-1] - 2 (tau2 mu2)[i] + (tau2 mu2)[i+1]
// (terms outside the utterance vanish; tau1, tau2 of the first and last frame are 1e-11, :113-116).  fp64 like the reference.
//
// P does not depend on the data, and its rows 0 .. T - 3 do not depend on T either (only the last two rows see the switched-off
// experts of the final frame), so the L D L^T factors of those rows are the same for EVERY utterance of a dimension.  Two kernels:
//   mlpg_factor_kernel  one thread per dimension runs the factor recurrence (a chain of dependent divisions, ~1 k cycles per row for
//                       a lone warp) ONCE into a table [row][4][D] = (l1, l2, d, 1 / d) -- and stops as soon as the state repeats bit for bit (the
//                       recurrence is a fixed function of (d_{i-1}, d_{i-2}, l1_{i-1}), so a repeated state stays repeated): the
//                       factors of a diagonally dominant band converge within a few dozen rows, later rows read the last entry;
//   mlpg_solve_kernel   one thread per (utterance, dimension): forward substitution with the tabulated factors (two dependent
//                       FMAs per row; the last two rows and utterances shorter than 6 frames are factored in place), z = y (1 / d) to
//                       the workspace [frame][D], backward substitution.  Rows are read kCh at a time into registers, one chunk
//                       ahead of the chunk being processed, so no load sits in a dependency chain.
// The results do not depend on the batch an utterance is solved in (same operations on the same values in every path).
#include "common.cuh"

namespace b2w {

constexpr int kCh = 8;

struct Taus {
  double t0, t1i, t2i, tedge;
  int T;
  __device__ __forceinline__ double tau1(int t) const { return (t < 0 || t >= T) ? 0.0 : ((t == 0 || t == T - 1) ? tedge : t1i); }
  __device__ __forceinline__ double tau2(int t) const { return (t < 0 || t >= T) ? 0.0 : ((t == 0 || t == T - 1) ? tedge : t2i); }
};

__device__ __forceinline__ Taus make_taus(const double* __restrict__ var3, int D, int d, int T) {
  Taus t;
  t.t0 = 1.0 / var3[d];
  t.t1i = 1.0 / var3[D + d];
  t.t2i = 1.0 / var3[2 * D + d];
  t.tedge = 1.0 / 100000000000.0;
  t.T = T;
  return t;
}

// row i of the factorisation from the state of the two rows before it
__device__ __forceinline__ void factor_row(const Taus& ta, int i, double d1, double d2, double l1p, double& l1, double& l2, double& di) {
  const double pii = ta.t0 + 0.25 * (ta.tau1(i - 1) + ta.tau1(i + 1)) + ta.tau2(i - 1) + 4.0 * ta.tau2(i) + ta.tau2(i + 1);
  const double pi1 = (i >= 1) ? -2.0 * (ta.tau2(i - 1) + ta.tau2(i)) : 0.0;          // P[i][i-1]
  const double pi2 = (i >= 2) ? ta.tau2(i - 1) - 0.25 * ta.tau1(i - 1) : 0.0;        // P[i][i-2]
  l2 = (i >= 2) ? pi2 / d2 : 0.0;
  l1 = (i >= 1) ? (pi1 - l2 * d2 * l1p) / d1 : 0.0;
  di = pii - l1 * l1 * d1 - l2 * l2 * d2;
}

constexpr int kNoEnd = 0x3fffffff;  // "utterance length" of the table: its rows never see a final frame

// table [row][4][D] = (l1, l2, d, 1 / d) of rows 0 .. last[d]; rows beyond last[d] equal row last[d]
__global__ void mlpg_factor_kernel(const double* __restrict__ var3, const int64_t* __restrict__ frame_off, int num_utts, int D,
                                   double* __restrict__ table, int* __restrict__ last) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  int64_t tmax = 0;
  for (int u = 0; u < num_utts; ++u) tmax = max(tmax, frame_off[u + 1] - frame_off[u]);
  const Taus ta = make_taus(var3, D, d, kNoEnd);
  double d1 = 0.0, d2 = 0.0, l1p = 0.0;
  int i = 0;
  for (; i < tmax; ++i) {
    double l1, l2, di;
    factor_row(ta, i, d1, d2, l1p, l1, l2, di);
    double* w = table + ((int64_t)i * 4) * D + d;
    w[0] = l1;
    w[D] = l2;
    w[2 * D] = di;
    w[3 * D] = 1.0 / di;
    const bool fixed = i >= 4 && di == d1 && d1 == d2 && l1 == l1p;  // rows >= 2 share P: the same state gives the same row again
    d2 = d1; d1 = di; l1p = l1;
    if (fixed) break;
  }
  last[d] = (int)min((int64_t)i, max(tmax - 1, (int64_t)0));
}

template <typename FT>
__global__ void __launch_bounds__(128) mlpg_solve_kernel(const FT* __restrict__ feats, int64_t feat_stride, const double* __restrict__ var3,
                                                         const int64_t* __restrict__ frame_off, int num_utts, int D,
                                                         const double* __restrict__ table, const int* __restrict__ last,
                                                         double* __restrict__ zws, double* __restrict__ out, int64_t out_stride) {
  const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (gid >= (int64_t)num_utts * D) return;
  const int u = (int)(gid / D), d = (int)(gid - (int64_t)u * D);
  const int64_t f0 = frame_off[u];
  const int T = (int)(frame_off[u + 1] - f0);
  if (T <= 0) return;
  const Taus ta = make_taus(var3, D, d, T);
  const int lastrow = last[d];
  const int first_own = T < 6 ? 0 : T - 2;  // rows factored here: the last two (they see the final frame), or all of a short utterance
  double el1[6], el2[6];                    // their l1, l2 (row first_own + k), for the backward sweep
  const FT* col = feats + f0 * feat_stride + d;
  const double* tab = table + d;
  const int64_t D4 = 4 * (int64_t)D;
  double* zcol = zws + f0 * D + d;
  double* ocol = out + f0 * out_stride + d;
  // rows i0 .. i0 + kCh - 1: static mean of row i, delta / delta-delta means of row i + 1 (0 beyond the utterance)
  auto load_fwd = [&](int i0, double (&a0)[kCh], double (&a1)[kCh], double (&a2)[kCh]) {
    const FT* p = col + (int64_t)i0 * feat_stride;
    if (i0 + kCh < T) {  // whole chunk and its look-ahead row inside the utterance
#pragma unroll
      for (int k = 0; k < kCh; ++k) {
        a0[k] = (double)p[0];
        a1[k] = (double)p[feat_stride + D];
        a2[k] = (double)p[feat_stride + 2 * D];
        p += feat_stride;
      }
    } else {
#pragma unroll
      for (int k = 0; k < kCh; ++k) {
        const int i = i0 + k;
        a0[k] = i < T ? (double)p[0] : 0.0;
        a1[k] = i + 1 < T ? (double)p[feat_stride + D] : 0.0;
        a2[k] = i + 1 < T ? (double)p[feat_stride + 2 * D] : 0.0;
        p += feat_stride;
      }
    }
  };
  // factors of the rows beyond the table's last row (the recurrence is stationary there): kept in registers
  const double sl1 = tab[(int64_t)lastrow * D4], sl2 = tab[(int64_t)lastrow * D4 + D], sd = tab[(int64_t)lastrow * D4 + 2 * D],
               srd = tab[(int64_t)lastrow * D4 + 3 * D];
  // forward: y = L^-1 b, z = D^-1 y
  double d1 = 0.0, d2 = 0.0, l1p = 0.0;  // d_{i-1}, d_{i-2}, l1_{i-1}
  double y1 = 0.0, y2 = 0.0;             // y_{i-1}, y_{i-2}
  double m1m = 0.0, m1c = (double)col[D] * ta.tau1(0), m2m = 0.0, m2c = (double)col[2 * D] * ta.tau2(0);  // (tau mu) at i-1 and i
  auto fwd_chunk = [&](int i0, const double (&c0)[kCh], const double (&c1)[kCh], const double (&c2)[kCh]) {
    double* zp = zcol + (int64_t)i0 * D;
    const bool interior = i0 >= 2 && i0 + kCh <= first_own;  // every expert of rows i - 1 .. i + 1 switched on, factors tabulated
    if (interior && i0 > lastrow) {
#pragma unroll
      for (int k = 0; k < kCh; ++k) {
        const double m1n = c1[k] * ta.t1i, m2n = c2[k] * ta.t2i;
        const double bi = ta.t0 * c0[k] + 0.5 * m1m - 0.5 * m1n + m2m - 2.0 * m2c + m2n;
        const double yi = bi - sl1 * y1 - sl2 * y2;
        zp[0] = yi * srd;
        zp += D;
        y2 = y1; y1 = yi;
        m1m = m1c; m1c = m1n; m2m = m2c; m2c = m2n;
      }
      d2 = sd;
      d1 = sd;
      l1p = sl1;
      return;
    }
    double tl1[kCh], tl2[kCh], td[kCh], trd[kCh];
    if (i0 + kCh - 1 <= lastrow) {  // tabulated factors of this chunk (shared by all utterances: cache hits)
      const double* w = tab + (int64_t)i0 * D4;
#pragma unroll
      for (int k = 0; k < kCh; ++k) {
        tl1[k] = w[0];
        tl2[k] = w[D];
        td[k] = w[2 * D];
        trd[k] = w[3 * D];
        w += D4;
      }
    } else {
#pragma unroll
      for (int k = 0; k < kCh; ++k) {
        const double* w = tab + (int64_t)min(i0 + k, lastrow) * D4;
        tl1[k] = w[0];
        tl2[k] = w[D];
        td[k] = w[2 * D];
        trd[k] = w[3 * D];
      }
    }
    if (interior) {
#pragma unroll
      for (int k = 0; k < kCh; ++k) {
        const double m1n = c1[k] * ta.t1i, m2n = c2[k] * ta.t2i;
        const double bi = ta.t0 * c0[k] + 0.5 * m1m - 0.5 * m1n + m2m - 2.0 * m2c + m2n;
        const double yi = bi - tl1[k] * y1 - tl2[k] * y2;
        zp[0] = yi * trd[k];
        zp += D;
        y2 = y1; y1 = yi;
        m1m = m1c; m1c = m1n; m2m = m2c; m2c = m2n;
      }
      d2 = td[kCh - 2];
      d1 = td[kCh - 1];
      l1p = tl1[kCh - 1];
      return;
    }
#pragma unroll
    for (int k = 0; k < kCh; ++k) {
      const int i = i0 + k;
      if (i < T) {
        const double m1n = c1[k] * ta.tau1(i + 1), m2n = c2[k] * ta.tau2(i + 1);
        const double bi = ta.t0 * c0[k] + 0.5 * m1m - 0.5 * m1n + m2m - 2.0 * m2c + m2n;
        double l1 = tl1[k], l2 = tl2[k], di = td[k], rdi = trd[k];
        if (i >= first_own) {
          factor_row(ta, i, d1, d2, l1p, l1, l2, di);
          rdi = 1.0 / di;
          el1[i - first_own] = l1;
          el2[i - first_own] = l2;
        }
        const double yi = bi - l1 * y1 - l2 * y2;
        zp[0] = yi * rdi;
        d2 = d1; d1 = di; l1p = l1; y2 = y1; y1 = yi;
        m1m = m1c; m1c = m1n; m2m = m2c; m2c = m2n;
      }
      zp += D;
    }
  };
  double pa0[kCh], pa1[kCh], pa2[kCh], pb0[kCh], pb1[kCh], pb2[kCh];  // two chunk buffers, used alternately (no copies)
  load_fwd(0, pa0, pa1, pa2);
  for (int i0 = 0; i0 < T; i0 += 2 * kCh) {
    load_fwd(i0 + kCh, pb0, pb1, pb2);
    fwd_chunk(i0, pa0, pa1, pa2);
    if (i0 + kCh < T) {
      load_fwd(i0 + 2 * kCh, pa0, pa1, pa2);
      fwd_chunk(i0 + kCh, pb0, pb1, pb2);
    }
  }
  // backward: x_i = z_i - l1_{i+1} x_{i+1} - l2_{i+2} x_{i+2}; chunks of rows i0 - k, k < kCh, loaded one chunk ahead
  // (a0, a1 = tabulated l1, l2 of the rows -- not loaded for chunks that lie wholly in the stationary part)
  auto bwd_stationary = [&](int i0) { return i0 < first_own && i0 - kCh + 1 > lastrow; };
  auto load_bwd = [&](int i0, double (&a0)[kCh], double (&a1)[kCh], double (&a2)[kCh]) {
    if (bwd_stationary(i0)) {
      const double* z = zcol + (int64_t)i0 * D;
#pragma unroll
      for (int k = 0; k < kCh; ++k) {
        a2[k] = z[0];
        z -= D;
      }
    } else if (i0 - kCh + 1 >= 0 && i0 <= lastrow) {
      const double* w = tab + (int64_t)i0 * D4;
      const double* z = zcol + (int64_t)i0 * D;
#pragma unroll
      for (int k = 0; k < kCh; ++k) {
        a0[k] = w[0];
        a1[k] = w[D];
        a2[k] = z[0];
        w -= D4;
        z -= D;
      }
    } else {
#pragma unroll
      for (int k = 0; k < kCh; ++k) {
        const int i = i0 - k;
        const int r = max(i, 0);
        const double* w = tab + (int64_t)min(r, lastrow) * D4;
        a0[k] = w[0];
        a1[k] = w[D];
        a2[k] = i >= 0 ? zcol[(int64_t)r * D] : 0.0;
      }
    }
  };
  double x1 = 0.0, x2 = 0.0, l1n = 0.0, l2n = 0.0, l2nn = 0.0;  // x_{i+1}, x_{i+2}, l1_{i+1}, l2_{i+1}, l2_{i+2}
  auto bwd_chunk = [&](int i0, const double (&c0)[kCh], const double (&c1)[kCh], const double (&c2)[kCh]) {
    double* op = ocol + (int64_t)i0 * out_stride;
    if (bwd_stationary(i0)) {
#pragma unroll
      for (int k = 0; k < kCh; ++k) {
        const double xi = c2[k] - l1n * x1 - l2nn * x2;
        op[0] = xi;
        op -= out_stride;
        x2 = x1; x1 = xi;
        l2nn = l2n;
        l1n = sl1;
        l2n = sl2;
      }
    } else if (i0 < first_own && i0 - kCh + 1 >= 0) {
#pragma unroll
      for (int k = 0; k < kCh; ++k) {
        const double xi = c2[k] - l1n * x1 - l2nn * x2;
        op[0] = xi;
        op -= out_stride;
        x2 = x1; x1 = xi;
        l2nn = l2n;
        l1n = c0[k];
        l2n = c1[k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < kCh; ++k) {
        const int i = i0 - k;
        if (i >= 0) {
          const double xi = c2[k] - l1n * x1 - l2nn * x2;
          op[0] = xi;
          x2 = x1; x1 = xi;
          l2nn = l2n;
          l1n = i >= first_own ? el1[i - first_own] : c0[k];
          l2n = i >= first_own ? el2[i - first_own] : c1[k];
        }
        op -= out_stride;
      }
    }
  };
  load_bwd(T - 1, pa0, pa1, pa2);
  for (int i0 = T - 1; i0 >= 0; i0 -= 2 * kCh) {
    load_bwd(i0 - kCh, pb0, pb1, pb2);
    bwd_chunk(i0, pa0, pa1, pa2);
    if (i0 - kCh >= 0) {
      load_bwd(i0 - 2 * kCh, pa0, pa1, pa2);
      bwd_chunk(i0 - kCh, pb0, pb1, pb2);
    }
  }
}

}  // namespace b2w

// workspace: z [F][D], the factor table [<= F rows][4][D], the table's last row per dimension (ints, padded to doubles)
extern "C" int64_t b2w_mlpg_workspace_doubles(int64_t num_frames, int32_t D) { return 5 * num_frames * (int64_t)D + (D + 1) / 2 + 1; }

extern "C" int b2w_mlpg(const void* feats, int32_t feats_dtype, int64_t feat_stride, const double* var3, const int64_t* frame_off,
                        int32_t num_utts, int32_t D, int64_t num_frames, double* workspace, double* out, int64_t out_stride,
                        void* stream) {
  using namespace b2w;
  B2W_REQUIRE(feats && var3 && frame_off && workspace && out, "b2w_mlpg: null argument");
  B2W_REQUIRE(feats_dtype == B2W_F64 || feats_dtype == B2W_F32, "b2w_mlpg: bad feats_dtype %d", feats_dtype);
  B2W_REQUIRE(D >= 1 && feat_stride >= 3 * (int64_t)D && out_stride >= D && num_frames >= 0, "b2w_mlpg: bad D %d / strides", D);
  if (num_utts <= 0 || num_frames == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  double* zws = workspace;
  double* table = workspace + num_frames * (int64_t)D;
  int* last = reinterpret_cast<int*>(table + 4 * num_frames * (int64_t)D);
  mlpg_factor_kernel<<<(D + 31) / 32, 32, 0, st>>>(var3, frame_off, num_utts, D, table, last);
  if (int rc = check_launch("mlpg_factor_kernel")) return rc;
  const int64_t threads = (int64_t)num_utts * D;
  const unsigned grid = (unsigned)((threads + 127) / 128);
  if (feats_dtype == B2W_F64)
    mlpg_solve_kernel<double><<<grid, 128, 0, st>>>((const double*)feats, feat_stride, var3, frame_off, num_utts, D, table, last, zws, out,
                                                    out_stride);
  else
    mlpg_solve_kernel<float><<<grid, 128, 0, st>>>((const float*)feats, feat_stride, var3, frame_off, num_utts, D, table, last, zws, out,
                                                   out_stride);
  return check_launch("mlpg_solve_kernel");
}
