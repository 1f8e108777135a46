// MLPG: maximum-probability parameter generation for every (utterance, feature dimension) of a ragged batch.
//
// Replaces MLPG.generation (idiaptts/misc/mlpg.py:94-127, called from WorldFeatLabelGen._postprocess_world :357-415 on the
// network output [static | delta | delta-delta]): the reference builds, per dimension, the banded precision matrix
// P = sum_w W_w^T diag(tau_w) W_w and b = sum_w W_w^T (mu_w tau_w) with bandmat (:54-92) and solves P x = b by banded Cholesky
// (:125).  With the three fixed windows [1], [-0.5 0 0.5], [1 -2 1] (:95-99) P is pentadiagonal with closed-form entries
//     P[i][i]   = tau0 + 0.25 (tau1[i-1] + tau1[i+1]) + tau2[i-1] + 4 tau2[i] + tau2[i+1]
//     P[i][i+1] = -2 (tau2[i] + tau2[i+1])          P[i][i+2] = tau2[i+1] - 0.25 tau1[i+1]
//     b[i]      = tau0 mu0[i] + 0.5 (tau1 mu1)[i-1] - 0.5 (tau1 mu1)[i+1] + (tau2 mu2)[i-1] - 2 (tau2 mu2)[i] + (tau2 mu2)[i+1]
// (terms outside the utterance vanish; tau1, tau2 of the first and last frame are 1e-11, :113-116).  One thread owns one
// (utterance, dimension): LDL^T forward sweep keeping two rows of state in registers, factors to a workspace laid out
// [frame][3][D] (coalesced across dimensions), backward sweep.  fp64 like the reference.
#include "common.cuh"

namespace b2w {

template <typename FT>
__global__ void __launch_bounds__(128) mlpg_kernel(const FT* __restrict__ feats, int64_t feat_stride, const double* __restrict__ var3,
                                                   const int64_t* __restrict__ frame_off, int num_utts, int D, double* __restrict__ ws,
                                                   double* __restrict__ out, int64_t out_stride) {
  const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (gid >= (int64_t)num_utts * D) return;
  const int u = (int)(gid / D), d = (int)(gid - (int64_t)u * D);
  const int64_t f0 = frame_off[u];
  const int T = (int)(frame_off[u + 1] - f0);
  if (T <= 0) return;
  const double t0 = 1.0 / var3[d], t1i = 1.0 / var3[D + d], t2i = 1.0 / var3[2 * D + d], tedge = 1.0 / 100000000000.0;
  auto tau1 = [&](int t) { return (t < 0 || t >= T) ? 0.0 : ((t == 0 || t == T - 1) ? tedge : t1i); };
  auto tau2 = [&](int t) { return (t < 0 || t >= T) ? 0.0 : ((t == 0 || t == T - 1) ? tedge : t2i); };
  auto mu = [&](int t, int w) { return (t < 0 || t >= T) ? 0.0 : (double)feats[(f0 + t) * feat_stride + w * D + d]; };
  // forward: L D L^T of the pentadiagonal matrix, y = L^-1 b
  double d1 = 0.0, d2 = 0.0;      // d_{i-1}, d_{i-2}
  double l1p = 0.0;               // l1_{i-1}
  double y1 = 0.0, y2 = 0.0;      // y_{i-1}, y_{i-2}
  double m1m = 0.0, m1c = mu(0, 1) * tau1(0), m2m = 0.0, m2c = mu(0, 2) * tau2(0);  // (tau mu) at i-1 and i
  for (int i = 0; i < T; ++i) {
    const double m1n = mu(i + 1, 1) * tau1(i + 1), m2n = mu(i + 1, 2) * tau2(i + 1);
    const double pii = t0 + 0.25 * (tau1(i - 1) + tau1(i + 1)) + tau2(i - 1) + 4.0 * tau2(i) + tau2(i + 1);
    const double pi1 = (i >= 1) ? -2.0 * (tau2(i - 1) + tau2(i)) : 0.0;          // P[i][i-1]
    const double pi2 = (i >= 2) ? tau2(i - 1) - 0.25 * tau1(i - 1) : 0.0;        // P[i][i-2]
    const double bi = t0 * mu(i, 0) + 0.5 * m1m - 0.5 * m1n + m2m - 2.0 * m2c + m2n;
    const double l2 = (i >= 2) ? pi2 / d2 : 0.0;
    const double l1 = (i >= 1) ? (pi1 - l2 * d2 * l1p) / d1 : 0.0;
    const double di = pii - l1 * l1 * d1 - l2 * l2 * d2;
    const double yi = bi - l1 * y1 - l2 * y2;
    double* w = ws + ((f0 + i) * 3) * D + d;
    w[0] = l1;
    w[D] = l2;
    w[2 * D] = yi / di;
    d2 = d1; d1 = di; l1p = l1; y2 = y1; y1 = yi;
    m1m = m1c; m1c = m1n; m2m = m2c; m2c = m2n;
  }
  // backward: x_i = y_i / d_i - l1_{i+1} x_{i+1} - l2_{i+2} x_{i+2}
  double x1 = 0.0, x2 = 0.0, l1n = 0.0, l2n = 0.0, l2nn = 0.0;  // x_{i+1}, x_{i+2}, l1_{i+1}, l2_{i+1}, l2_{i+2}
  for (int i = T - 1; i >= 0; --i) {
    const double* w = ws + ((f0 + i) * 3) * D + d;
    const double xi = w[2 * D] - l1n * x1 - l2nn * x2;
    out[(f0 + i) * out_stride + d] = xi;
    x2 = x1; x1 = xi;
    l2nn = l2n;
    l1n = w[0];
    l2n = w[D];
  }
}

}  // namespace b2w

extern "C" int64_t b2w_mlpg_workspace_doubles(int64_t num_frames, int32_t D) { return 3 * num_frames * (int64_t)D; }

extern "C" int b2w_mlpg(const void* feats, int32_t feats_dtype, int64_t feat_stride, const double* var3, const int64_t* frame_off,
                        int32_t num_utts, int32_t D, double* workspace, double* out, int64_t out_stride, void* stream) {
  using namespace b2w;
  B2W_REQUIRE(feats && var3 && frame_off && workspace && out, "b2w_mlpg: null argument");
  B2W_REQUIRE(feats_dtype == B2W_F64 || feats_dtype == B2W_F32, "b2w_mlpg: bad feats_dtype %d", feats_dtype);
  B2W_REQUIRE(D >= 1 && feat_stride >= 3 * (int64_t)D && out_stride >= D, "b2w_mlpg: bad D %d / strides", D);
  if (num_utts <= 0) return 0;
  const int64_t threads = (int64_t)num_utts * D;
  const unsigned grid = (unsigned)((threads + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
  if (feats_dtype == B2W_F64)
    mlpg_kernel<double><<<grid, 128, 0, st>>>((const double*)feats, feat_stride, var3, frame_off, num_utts, D, workspace, out, out_stride);
  else
    mlpg_kernel<float><<<grid, 128, 0, st>>>((const float*)feats, feat_stride, var3, frame_off, num_utts, D, workspace, out, out_stride);
  return check_launch("mlpg_kernel");
}
