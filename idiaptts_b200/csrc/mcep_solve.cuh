// One-warp blocked LDL^T solve used by the mel-cepstrum Newton step (CUDA-core and tensor-core kernels share it).
#pragma once
#include "common.cuh"

namespace b2w {

__host__ __device__ inline int pad4(int n) { return (n + 3) & ~3; }

// floats of workspace one warp needs for an n x n system with block stride KB (n padded to NBk = ceil(n / 4) blocks)
__host__ __device__ inline int ldl_workspace_floats(int NBk, int KB) { return (NBk * (NBk + 1) / 2) * KB + NBk * KB + 8 * NBk; }

// ---- blocked LDL^T solve of the (m+1)x(m+1) system M d = b with M[i][k] = rt[|i-k|] + rt[i+k], one warp per frame ------
// 4x4 blocks of the lower triangle, block (I, Kb <= I) at (I (I+1) / 2 + Kb) * KB floats; a 20-float stride keeps the
// float4 accesses of lanes working on consecutive blocks on distinct banks.  The right-hand side rides along as an
// extra column (forward substitution fused into the panel step); `tri` maps a flat pair index to (a, q), q <= a.
template <int KB>
__device__ __forceinline__ int blk_index(int I, int Kb) { return (I * (I + 1) / 2 + Kb) * KB; }

// returns false (warp-uniform) when a pivot is not positive
template <int KB>
__device__ __forceinline__ bool warp_ldl_solve(const float* __restrict__ rt, const float* __restrict__ al, int n, int NBk,
                                              const uint16_t* __restrict__ tri, float* __restrict__ ws,
                                              float* __restrict__ x_out) {
  const int lane = threadIdx.x & 31;
  const int np = 4 * NBk;
  float* A = ws;                                     // NBk (NBk+1) / 2 blocks
  float* Wp = A + (NBk * (NBk + 1) / 2) * KB;      // NBk blocks: panel W = L D
  float* dv = Wp + NBk * KB;                       // np reciprocal pivots
  float* bv = dv + np;                               // np right-hand side -> y -> solution
  // build: one float4 (row r of block (I, Kb)) per lane and step
  for (int I = 0; I < NBk; ++I) {
    for (int e = lane; e < 4 * (I + 1); e += 32) {
      const int Kb = e >> 2, r = e & 3;
      const int i = 4 * I + r, k = 4 * Kb;
      float4 v;
      if (i < n && k + 3 < n) {
        v.x = rt[abs(i - k)] + rt[i + k];
        v.y = rt[abs(i - k - 1)] + rt[i + k + 1];
        v.z = rt[abs(i - k - 2)] + rt[i + k + 2];
        v.w = rt[abs(i - k - 3)] + rt[i + k + 3];
      } else {
        float t[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) t[c] = (i < n && k + c < n) ? rt[abs(i - k - c)] + rt[i + k + c] : (i == k + c ? 1.f : 0.f);
        v = make_float4(t[0], t[1], t[2], t[3]);
      }
      *reinterpret_cast<float4*>(A + blk_index<KB>(I, Kb) + 4 * r) = v;
    }
  }
  for (int i = lane; i < np; i += 32) bv[i] = (i < n) ? rt[i] - al[i] : 0.f;
  __syncwarp();
  bool ok = true;
  for (int J = 0; J < NBk; ++J) {
    // (a) diagonal block, redundantly in every lane (broadcast loads)
    const float* a = A + blk_index<KB>(J, J);
    const float4 q0 = *reinterpret_cast<const float4*>(a), q1 = *reinterpret_cast<const float4*>(a + 4),
                 q2 = *reinterpret_cast<const float4*>(a + 8), q3 = *reinterpret_cast<const float4*>(a + 12);
    const float d0 = q0.x, r0 = 1.f / d0;
    const float l10 = q1.x * r0, l20 = q2.x * r0, l30 = q3.x * r0;
    const float d1 = q1.y - l10 * q1.x, r1 = 1.f / d1;
    const float l21 = (q2.y - l20 * q1.x) * r1, l31 = (q3.y - l30 * q1.x) * r1;
    const float d2 = q2.z - l20 * q2.x - l21 * (q2.y - l20 * q1.x), r2 = 1.f / d2;
    const float l32 = (q3.z - l30 * q2.x - l31 * (q2.y - l20 * q1.x)) * r2;
    const float d3 = q3.w - l30 * q3.x - l31 * (q3.y - l30 * q1.x) - l32 * (q3.z - l30 * q2.x - l31 * (q2.y - l20 * q1.x));
    const float r3 = 1.f / d3;
    if (!(d0 > 0.f && d1 > 0.f && d2 > 0.f && d3 > 0.f)) ok = false;
    // forward substitution inside the diagonal block: y_J = L_JJ^-1 b_J (b_J already holds b - sum_{K<J} L_JK y_K)
    const float4 bj = *reinterpret_cast<const float4*>(bv + 4 * J);
    const float y0 = bj.x;
    const float y1 = bj.y - l10 * y0;
    const float y2 = bj.z - l20 * y0 - l21 * y1;
    const float y3 = bj.w - l30 * y0 - l31 * y1 - l32 * y2;
    __syncwarp();
    if (lane == 0) {
      float* w = A + blk_index<KB>(J, J);
      w[4] = l10; w[8] = l20; w[9] = l21; w[12] = l30; w[13] = l31; w[14] = l32;
      *reinterpret_cast<float4*>(dv + 4 * J) = make_float4(r0, r1, r2, r3);
      *reinterpret_cast<float4*>(bv + 4 * J) = make_float4(y0, y1, y2, y3);
    }
    // (b) panel, one block row per lane: W = A_IJ L_JJ^-T, L_IJ = W D^-1, b_I -= L_IJ y_J
    {
      const int r = lane & 3;
      for (int I = J + 1 + (lane >> 2); I < NBk; I += 8) {
        float* X = A + blk_index<KB>(I, J) + 4 * r;
        const float4 x = *reinterpret_cast<const float4*>(X);
        const float w0 = x.x;
        const float w1 = x.y - l10 * w0;
        const float w2 = x.z - l20 * w0 - l21 * w1;
        const float w3 = x.w - l30 * w0 - l31 * w1 - l32 * w2;
        *reinterpret_cast<float4*>(Wp + I * KB + 4 * r) = make_float4(w0, w1, w2, w3);
        const float4 l = make_float4(w0 * r0, w1 * r1, w2 * r2, w3 * r3);
        *reinterpret_cast<float4*>(X) = l;
        bv[4 * I + r] -= l.x * y0 + l.y * y1 + l.z * y2 + l.w * y3;
      }
    }
    __syncwarp();
    // (c) trailing update A_IK -= W_IJ L_KJ^T for J < Kb <= I
    const int rr = NBk - 1 - J;
    const int cnt = rr * (rr + 1) / 2;
    for (int p = lane; p < cnt; p += 32) {
      const int code = tri[p];
      const int I = J + 1 + (code >> 8), Kb = J + 1 + (code & 255);
      const float* W = Wp + I * KB;
      const float* Lk = A + blk_index<KB>(Kb, J);
      float* T = A + blk_index<KB>(I, Kb);
      float4 lk[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) lk[c] = *reinterpret_cast<const float4*>(Lk + 4 * c);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 w = *reinterpret_cast<const float4*>(W + 4 * r);
        float4 t = *reinterpret_cast<const float4*>(T + 4 * r);
        t.x -= w.x * lk[0].x + w.y * lk[0].y + w.z * lk[0].z + w.w * lk[0].w;
        t.y -= w.x * lk[1].x + w.y * lk[1].y + w.z * lk[1].z + w.w * lk[1].w;
        t.z -= w.x * lk[2].x + w.y * lk[2].y + w.z * lk[2].z + w.w * lk[2].w;
        t.w -= w.x * lk[3].x + w.y * lk[3].y + w.z * lk[3].z + w.w * lk[3].w;
        *reinterpret_cast<float4*>(T + 4 * r) = t;
      }
    }
    __syncwarp();
  }
  ok = __all_sync(0xffffffffu, ok);
  // z = D^-1 y
  for (int i = lane; i < np; i += 32) bv[i] *= dv[i];
  __syncwarp();
  // backward substitution L^T x = z
  for (int J = NBk - 1; J >= 0; --J) {
    const float* L = A + blk_index<KB>(J, J);
    const float4 zj = *reinterpret_cast<const float4*>(bv + 4 * J);
    const float x3 = zj.w;
    const float x2 = zj.z - L[14] * x3;
    const float x1 = zj.y - L[9] * x2 - L[13] * x3;
    const float x0 = zj.x - L[4] * x1 - L[8] * x2 - L[12] * x3;
    __syncwarp();
    if (lane == 0) *reinterpret_cast<float4*>(bv + 4 * J) = make_float4(x0, x1, x2, x3);
    {
      const int c = lane & 3;
      for (int Kb = lane >> 2; Kb < J; Kb += 8) {
        const float* Lj = A + blk_index<KB>(J, Kb);
        bv[4 * Kb + c] -= Lj[c] * x0 + Lj[4 + c] * x1 + Lj[8 + c] * x2 + Lj[12 + c] * x3;
      }
    }
    __syncwarp();
  }
  for (int i = lane; i < n; i += 32) x_out[i] = bv[i];
  __syncwarp();
  return ok;
}


}  // namespace b2w
