// Objective metrics of WORLD feature rows on the device, one launch for a ragged batch of utterances.
//
// Replaces the per-utterance numpy code of Metrics (idiaptts/src/Metrics.py): mcd_k (:84-92, nnmnkwii melcd, c0 excluded),
// f0_rmse (:94-106), gross_pitch_error (:108-126), voicing_decision_error (:150-155), f0_frame_error (:128-148, derived on the
// host from the same sums) and aperiodicity_distortion (:157-164).  Rows are [coded_sp(D) | lf0 | vuv | bap(nap)] as produced by
// WorldFeatLabelGen.convert_from_world_features; one thread per frame, fp64 sums per utterance:
//   acc[u][0] sum_t ||c_org - c_out||_2 over bins 1..D-1        acc[u][1] sum_t vuv_org (exp lf0_org - exp lf0_out)^2
//   acc[u][2] sum_t vuv_org                                     acc[u][3] sum_t [|lf0_org - lf0_out| > 0.2 lf0_org] vuv_org vuv_out
//   acc[u][4] sum_t vuv_org vuv_out                             acc[u][5] sum_t [vuv_org != vuv_out]
//   acc[u][6] nap > 1: sum_t ||bap_org - bap_out||_2 over bins 1..nap-1;  nap == 1: sum_t (bap_org - bap_out)^2
//   acc[u][7] frames
#include "common.cuh"

namespace b2w {

__global__ void __launch_bounds__(256) world_metrics_kernel(const float* __restrict__ org, const float* __restrict__ out, int64_t stride,
                                                            const int32_t* __restrict__ frame_utt, int64_t num_frames, int D, int nap,
                                                            double* __restrict__ acc) {
  const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= num_frames) return;
  const float* a = org + f * stride;
  const float* b = out + f * stride;
  double s = 0.0;
  for (int d = 1; d < D; ++d) {
    const double e = (double)a[d] - (double)b[d];
    s += e * e;
  }
  const double lo = a[D], lx = b[D], vo = a[D + 1], vx = b[D + 1];
  const double fo = exp(lo), fx = exp(lx);
  double sb = 0.0;
  if (nap > 1) {
    for (int d = 1; d < nap; ++d) {
      const double e = (double)a[D + 2 + d] - (double)b[D + 2 + d];
      sb += e * e;
    }
    sb = sqrt(sb);
  } else {
    const double e = (double)a[D + 2] - (double)b[D + 2];
    sb = e * e;
  }
  double* r = acc + (int64_t)frame_utt[f] * 8;
  atomicAdd(r + 0, sqrt(s));
  atomicAdd(r + 1, vo * (fo - fx) * (fo - fx));
  atomicAdd(r + 2, vo);
  atomicAdd(r + 3, (fabs(lo - lx) > 0.2 * lo ? 1.0 : 0.0) * vo * vx);
  atomicAdd(r + 4, vo * vx);
  atomicAdd(r + 5, vo != vx ? 1.0 : 0.0);
  atomicAdd(r + 6, sb);
  atomicAdd(r + 7, 1.0);
}

}  // namespace b2w

extern "C" int b2w_world_metrics(const float* org, const float* out, int64_t stride, const int32_t* frame_utt, int64_t num_frames,
                                 int32_t num_coded_sps, int32_t num_bap, double* acc, void* stream) {
  using namespace b2w;
  B2W_REQUIRE(org && out && frame_utt && acc, "b2w_world_metrics: null argument");
  B2W_REQUIRE(num_coded_sps >= 1 && num_bap >= 1 && stride >= num_coded_sps + 2 + num_bap, "b2w_world_metrics: bad dimensions");
  if (num_frames == 0) return 0;
  world_metrics_kernel<<<(unsigned)((num_frames + 255) / 256), 256, 0, (cudaStream_t)stream>>>(org, out, stride, frame_utt, num_frames,
                                                                                              num_coded_sps, num_bap, acc);
  return check_launch("world_metrics_kernel");
}
