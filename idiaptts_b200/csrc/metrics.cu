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

// Row-cooperative form for rows that can be read as float4 (stride a multiple of 4, 16-byte aligned planes): eight lanes share a
// frame (a warp reads four whole rows per request instead of 32 scattered ones), the squared differences are reduced over the eight
// lanes by shuffles, and a group walks kMetricRows consecutive frames keeping the eight sums of the current utterance in registers,
// so there are 8 fp64 atomics per group and utterance instead of 8 per frame.  Same terms, summed in a different order (fp64).
constexpr int kMetricRows = 16;

__global__ void __launch_bounds__(256) world_metrics_rows_kernel(const float* __restrict__ org, const float* __restrict__ out,
                                                                 int64_t stride, const int32_t* __restrict__ frame_utt,
                                                                 int64_t num_frames, int D, int nap, double* __restrict__ acc) {
  const int sub = threadIdx.x & 7;
  const int64_t group = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3;
  const int64_t r0 = group * kMetricRows;
  if (r0 >= num_frames) return;  // whole groups leave together: the shuffles below stay inside a group of eight lanes
  const int64_t r1 = min(num_frames, r0 + kMetricRows);
  const unsigned gmask = 0xffu << ((threadIdx.x & 31) & ~7);
  const int W = D + 2 + nap;
  const int b_lo = (nap > 1) ? D + 3 : D + 2, b_hi = D + 2 + nap;  // bap columns entering the distortion
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0;
  int cur = frame_utt[r0];
  auto flush = [&](int u) {
    if (sub == 0) {
      double* r = acc + (int64_t)u * 8;
      atomicAdd(r + 0, a0); atomicAdd(r + 1, a1); atomicAdd(r + 2, a2); atomicAdd(r + 3, a3);
      atomicAdd(r + 4, a4); atomicAdd(r + 5, a5); atomicAdd(r + 6, a6); atomicAdd(r + 7, a7);
    }
    a0 = a1 = a2 = a3 = a4 = a5 = a6 = a7 = 0.0;
  };
  for (int64_t f = r0; f < r1; ++f) {
    const int u = frame_utt[f];
    if (u != cur) {
      flush(cur);
      cur = u;
    }
    const float* a = org + f * stride;
    const float* b = out + f * stride;
    double s = 0.0, sb = 0.0;
    for (int c = 4 * sub; c < W; c += 32) {
      const float4 va = *reinterpret_cast<const float4*>(a + c), vb = *reinterpret_cast<const float4*>(b + c);
      const float ea[4] = {va.x, va.y, va.z, va.w}, eb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int col = c + k;  // columns past W (row padding) match neither range
        const double e = (double)ea[k] - (double)eb[k];
        if (col >= 1 && col < D) s += e * e;
        if (col >= b_lo && col < b_hi) sb += e * e;
      }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      s += __shfl_xor_sync(gmask, s, o);
      sb += __shfl_xor_sync(gmask, sb, o);
    }
    if (sub == 0) {
      const double lo = a[D], lx = b[D], vo = a[D + 1], vx = b[D + 1];
      const double fo = exp(lo), fx = exp(lx);
      a0 += sqrt(s);
      a1 += vo * (fo - fx) * (fo - fx);
      a2 += vo;
      a3 += (fabs(lo - lx) > 0.2 * lo ? 1.0 : 0.0) * vo * vx;
      a4 += vo * vx;
      a5 += vo != vx ? 1.0 : 0.0;
      a6 += (nap > 1) ? sqrt(sb) : sb;
      a7 += 1.0;
    }
  }
  flush(cur);
}

}  // namespace b2w

extern "C" int b2w_world_metrics(const float* org, const float* out, int64_t stride, const int32_t* frame_utt, int64_t num_frames,
                                 int32_t num_coded_sps, int32_t num_bap, double* acc, void* stream) {
  using namespace b2w;
  B2W_REQUIRE(org && out && frame_utt && acc, "b2w_world_metrics: null argument");
  B2W_REQUIRE(num_coded_sps >= 1 && num_bap >= 1 && stride >= num_coded_sps + 2 + num_bap, "b2w_world_metrics: bad dimensions");
  if (num_frames == 0) return 0;
  if (stride % 4 == 0 && (reinterpret_cast<uintptr_t>(org) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const int64_t groups = (num_frames + kMetricRows - 1) / kMetricRows;
    world_metrics_rows_kernel<<<(unsigned)((groups * 8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(org, out, stride, frame_utt,
                                                                                                    num_frames, num_coded_sps, num_bap, acc);
    return check_launch("world_metrics_rows_kernel");
  }
  world_metrics_kernel<<<(unsigned)((num_frames + 255) / 256), 256, 0, (cudaStream_t)stream>>>(org, out, stride, frame_utt, num_frames,
                                                                                              num_coded_sps, num_bap, acc);
  return check_launch("world_metrics_kernel");
}
