// Tile-level building blocks shared by the CUDA-core mel-cepstrum kernels (mcep.cu: SPTK mcep; mgcep.cu: SPTK mgcep):
// a CTA of kMcThreads threads owns a tile of F frames in shared memory and contracts it against constant tables in L2.
#pragma once
#include "common.cuh"
#include "mcep_solve.cuh"

namespace b2w {

constexpr int kBlk = 20;  // padded block stride of the solve workspace (bank-conflict free float4 accesses)
constexpr int kMcThreads = 256;
constexpr int kMcWarps = kMcThreads / 32;


// out[f][n] = sum_j tile[f][j] * mt[j][n], n < nout.  Thread (c = tid % (NPAD/2), group = tid / (NPAD/2)) owns the two
// output columns c and c + NPAD/2 for FG = 2 F NPAD / 512... frames of its group: every float4 broadcast of a tile row
// feeds 8 FMAs, the matrix columns are read coalesced from L2.
template <int F, int NPAD, bool ACC = false>
__device__ __forceinline__ void gemm_tile_by_matrix(const float* __restrict__ tile, int KP, int K, const float* __restrict__ mt,
                                                    int ldm, int nout, float* __restrict__ out, int ldo) {
  constexpr int HC = NPAD / 2;              // threads per frame group
  constexpr int G = kMcThreads / HC;        // frame groups
  constexpr int FG = F / G;                 // frames per thread
  static_assert(FG >= 1 && FG * G == F, "tile height does not fit this mapping");
  const int c0 = threadIdx.x % HC;
  const int c1 = c0 + HC;
  const int f0 = (threadIdx.x / HC) * FG;
  const bool v0 = c0 < nout, v1 = c1 < nout;
  if (!v0) return;
  const int c1s = v1 ? c1 : c0;  // keep loads in bounds
  float acc0[FG], acc1[FG];
#pragma unroll
  for (int f = 0; f < FG; ++f) { acc0[f] = 0.f; acc1[f] = 0.f; }
  const int K4 = K & ~3;
  for (int j = 0; j < K4; j += 4) {
    const float* r0 = mt + (int64_t)j * ldm;
    const float a0 = __ldg(r0 + c0), a1 = __ldg(r0 + ldm + c0), a2 = __ldg(r0 + 2 * ldm + c0), a3 = __ldg(r0 + 3 * ldm + c0);
    const float b0 = __ldg(r0 + c1s), b1 = __ldg(r0 + ldm + c1s), b2 = __ldg(r0 + 2 * ldm + c1s), b3 = __ldg(r0 + 3 * ldm + c1s);
#pragma unroll
    for (int f = 0; f < FG; ++f) {
      const float4 p4 = *reinterpret_cast<const float4*>(tile + (f0 + f) * KP + j);
      acc0[f] = fmaf(p4.x, a0, acc0[f]); acc1[f] = fmaf(p4.x, b0, acc1[f]);
      acc0[f] = fmaf(p4.y, a1, acc0[f]); acc1[f] = fmaf(p4.y, b1, acc1[f]);
      acc0[f] = fmaf(p4.z, a2, acc0[f]); acc1[f] = fmaf(p4.z, b2, acc1[f]);
      acc0[f] = fmaf(p4.w, a3, acc0[f]); acc1[f] = fmaf(p4.w, b3, acc1[f]);
    }
  }
  for (int j = K4; j < K; ++j) {
    const float a = __ldg(mt + (int64_t)j * ldm + c0), b = __ldg(mt + (int64_t)j * ldm + c1s);
#pragma unroll
    for (int f = 0; f < FG; ++f) {
      const float t = tile[(f0 + f) * KP + j];
      acc0[f] = fmaf(t, a, acc0[f]);
      acc1[f] = fmaf(t, b, acc1[f]);
    }
  }
#pragma unroll
  for (int f = 0; f < FG; ++f) {
    if (ACC) {  // out += (the second half of a cos / sin pair)
      out[(f0 + f) * ldo + c0] += acc0[f];
      if (v1) out[(f0 + f) * ldo + c1] += acc1[f];
    } else {
      out[(f0 + f) * ldo + c0] = acc0[f];
      if (v1) out[(f0 + f) * ldo + c1] = acc1[f];
    }
  }
}

template <int F, bool ACC = false>
__device__ __forceinline__ void gemm_tile_dispatch(const float* tile, int KP, int K, const float* mt, int ldm, int nout,
                                                   float* out, int ldo) {
  if (nout <= 64) gemm_tile_by_matrix<F, 64, ACC>(tile, KP, K, mt, ldm, nout, out, ldo);
  else if (nout <= 128) gemm_tile_by_matrix<F, 128, ACC>(tile, KP, K, mt, ldm, nout, out, ldo);
  else gemm_tile_by_matrix<F, 256, ACC>(tile, KP, K, mt, ldm, nout, out, ldo);
}

// C[f][j] = sum_k mc[f][k] * cmat[k][j] for the two columns j0, j1 and FH frames starting at fbase (registers)
template <int FH>
__device__ __forceinline__ void two_columns(const float* __restrict__ mc, int MP, const float* __restrict__ cmat, int K, int j0,
                                            int j1, int fbase, float* acc0, float* acc1) {
#pragma unroll
  for (int f = 0; f < FH; ++f) { acc0[f] = 0.f; acc1[f] = 0.f; }
  for (int k = 0; k < MP; k += 4) {
    // cmat is stored with pad4(m+1) rows (zero rows past m), mc rows are zero padded the same way
    const float* r0 = cmat + (int64_t)k * K;
    const float a0 = __ldg(r0 + j0), a1 = __ldg(r0 + K + j0), a2 = __ldg(r0 + 2 * K + j0), a3 = __ldg(r0 + 3 * K + j0);
    const float b0 = __ldg(r0 + j1), b1 = __ldg(r0 + K + j1), b2 = __ldg(r0 + 2 * K + j1), b3 = __ldg(r0 + 3 * K + j1);
#pragma unroll
    for (int f = 0; f < FH; ++f) {
      const float4 m4 = *reinterpret_cast<const float4*>(mc + (fbase + f) * MP + k);
      acc0[f] = fmaf(m4.x, a0, acc0[f]); acc1[f] = fmaf(m4.x, b0, acc1[f]);
      acc0[f] = fmaf(m4.y, a1, acc0[f]); acc1[f] = fmaf(m4.y, b1, acc1[f]);
      acc0[f] = fmaf(m4.z, a2, acc0[f]); acc1[f] = fmaf(m4.z, b2, acc1[f]);
      acc0[f] = fmaf(m4.w, a3, acc0[f]); acc1[f] = fmaf(m4.w, b3, acc1[f]);
    }
  }
}


}  // namespace b2w
