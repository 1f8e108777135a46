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

// returns false (warp-uniform) when a pivot is not positive.
// General form (mgcep.cu): M[i][k] = rt[|i-k|] + hk[i+k] with a separate Hankel sequence hk (nullptr: hk = rt, the mcep case) and
// right-hand side rhs (nullptr: rt[i] - al[i], the mcep case).
template <int KB>
__device__ __forceinline__ bool warp_ldl_solve(const float* __restrict__ rt, const float* __restrict__ al, int n, int NBk,
                                              const uint16_t* __restrict__ tri, float* __restrict__ ws,
                                              float* __restrict__ x_out, const float* __restrict__ hk = nullptr,
                                              const float* __restrict__ rhs = nullptr) {
  if (hk == nullptr) hk = rt;
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
        v.x = rt[abs(i - k)] + hk[i + k];
        v.y = rt[abs(i - k - 1)] + hk[i + k + 1];
        v.z = rt[abs(i - k - 2)] + hk[i + k + 2];
        v.w = rt[abs(i - k - 3)] + hk[i + k + 3];
      } else {
        float t[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) t[c] = (i < n && k + c < n) ? rt[abs(i - k - c)] + hk[i + k + c] : (i == k + c ? 1.f : 0.f);
        v = make_float4(t[0], t[1], t[2], t[3]);
      }
      *reinterpret_cast<float4*>(A + blk_index<KB>(I, Kb) + 4 * r) = v;
    }
  }
  for (int i = lane; i < np; i += 32) bv[i] = (i < n) ? (rhs ? rhs[i] : rt[i] - al[i]) : 0.f;
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


// ---- register-resident LDL^T solve of the same system (compile-time size NS, NS % 4 == 0, NS <= 64) --------------------------------
// The blocked solver above keeps the matrix in shared memory and is bound by dependent shared-memory round trips and by issue
// slots (~7.6 k warp instructions for NS = 60).  Here lane l owns rows l and NS-1-l of the lower triangle IN REGISTERS
// (l < NS/2; l + 1 and NS - l entries: balanced), the code is fully unrolled so every register index is static, and per pivot
// column j the only exchange is: the UNSCALED column W[k][j] = A[k][j] is written once to shared memory (packed column-major,
// kept for the back substitution) and read back as broadcast float4 vectors, while d_j and y_j travel by shuffle:
//       A[i][k] -= (W[i][j] / d_j) W[k][j],   y_i -= (W[i][j] / d_j) y_j              (two f32x2 FMAs per float4 of column)
// Entries of a lane's register rows beyond the row length hold garbage that is never read by valid data.
// Back substitution runs in axpy form on the packed columns: lane c accumulates t_c = sum_i W[i][c] x_i for c = lane, lane + 32.
__host__ __device__ constexpr int rr_col_start(int j) { return ((j + 1) / 4) * 4; }  // first stored row of column j
// Packed column storage: row k >= rr_col_start(j) of column j lives at Wc[rr_col_base(NS, j) + k].  Column starts are float4
// aligned and skewed so that base(j) = 4 j (mod 32): in the back substitution lane c reads column c, and the skew spreads the
// 32 lanes' float4 loads over all banks (4 wavefronts per request, the minimum for 512 bytes).
__host__ __device__ constexpr int rr_col_base(int NS, int j, bool want_end = false) {
  int end = 0, base = 0;
  for (int c = 0; c <= j; ++c) {
    const int st = rr_col_start(c);
    int off = end;
    off += (((4 * c + st) % 32 - off % 32) + 32) % 32;
    base = off - st;
    end = off + (NS - st);
  }
  return want_end ? end : base;
}
__host__ __device__ constexpr int rr_workspace_floats(int NS) { return rr_col_base(NS, NS - 2, true) + 64 + 128; }

__device__ __forceinline__ void ffma2(float2& c, const float2 a, const float2 b) {
#ifdef B2W_SOLVE_SCALAR_FMA   // experiment: two scalar FMAs instead of one packed one (bit-identical results)
  c.x = fmaf(a.x, b.x, c.x);
  c.y = fmaf(a.y, b.y, c.y);
  return;
#endif
  unsigned long long cc = *reinterpret_cast<unsigned long long*>(&c);
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(cc)
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  c = *reinterpret_cast<float2*>(&cc);
}

// rt: r~[0 .. 2 NS - 2] with at least 64 readable floats BEFORE rt[0] (garbage allowed); al: (-alpha)^k; Wc: per-warp scratch of
// rr_col_base(NS, NS - 2, true) floats.  On return lane l holds x[l] in x0 and x[l + 32] in x1.  Returns false (warp-uniform) when a
// pivot is not positive.
// colbase: shared-memory table of rr_col_base(NS, j), j < NS - 1 (filled once per CTA).
template <int NS>
__device__ __forceinline__ bool warp_rr_solve(const float* __restrict__ rt, const float* __restrict__ al, float* __restrict__ Wc,
                                             const int* __restrict__ colbase, float& x0, float& x1) {
  static_assert(NS % 4 == 0 && NS >= 8 && NS <= 64, "unsupported system size");
  constexpr int H = NS / 2;            // lanes that own rows
  constexpr int HP = (H + 3) & ~3;     // register row A, padded to whole float4 groups
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const bool act = lane < H;
  const int iA = act ? lane : 0, iB = act ? NS - 1 - lane : NS - 1;  // idle lanes mirror lane 0 (they never store)
  float2 a[HP / 2], b[NS / 2];
  {
    const float* rA = rt + iA;
    const float* rB = rt + iB;
#pragma unroll
    for (int k = 0; k < HP; k += 2) a[k / 2] = make_float2(rA[-k] + rA[k], rA[-k - 1] + rA[k + 1]);
#pragma unroll
    for (int k = 0; k < NS; k += 2) b[k / 2] = make_float2(rB[-k] + rB[k], rB[-k - 1] + rB[k + 1]);
  }
  float yA = rt[iA] - al[iA], yB = rt[iB] - al[iB];
  float rdk0 = 0.f, rdk1 = 0.f, zk0 = 0.f, zk1 = 0.f;  // 1 / d_c and z_c = y_c / d_c for c = lane, lane + 32
  bool ok = true;
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    // pivot and right-hand side of row j from their owner
    const float dsrc = (j < H) ? ((j & 1) ? a[j / 2].y : a[j / 2].x) : ((j & 1) ? b[j / 2].y : b[j / 2].x);
    const float ysrc = (j < H) ? yA : yB;
    const int owner = (j < H) ? j : NS - 1 - j;
    const float d = __shfl_sync(FULL, dsrc, owner);
    const float yj = __shfl_sync(FULL, ysrc, owner);
    ok = ok && (d > 0.f);
    float rd = __fdividef(1.f, d);  // MUFU.RCP + one Newton step: correctly rounded up to the last bit or two
    rd = fmaf(fmaf(-d, rd, 1.f), rd, rd);
    if (lane == (j & 31)) {
      if (j < 32) { rdk0 = rd; zk0 = yj * rd; }
      else { rdk1 = rd; zk1 = yj * rd; }
    }
    if (j == NS - 1) break;
    // this lane's entries of column j (0 where the row is not below the diagonal)
    float wA = 0.f;
    if (j < H) wA = (lane > j && act) ? ((j & 1) ? a[j / 2].y : a[j / 2].x) : 0.f;
    const float wB = (iB > j && act) ? ((j & 1) ? b[j / 2].y : b[j / 2].x) : 0.f;
    const int st = rr_col_start(j);
    float* col = Wc + colbase[j];  // indexed by the absolute row
    if (st < H) { if (act && iA >= st) col[iA] = wA; }
    if (act && iB >= st) col[iB] = wB;
    const float lA = wA * rd, lB = wB * rd;
    yA -= lA * yj;
    yB -= lB * yj;
    const float2 nA = make_float2(-lA, -lA), nB = make_float2(-lB, -lB);
    __syncwarp();
#pragma unroll
    for (int k4 = st; k4 < NS; k4 += 4) {
      const float4 c = *reinterpret_cast<const float4*>(col + k4);
      const float2 c01 = make_float2(c.x, c.y), c23 = make_float2(c.z, c.w);
      if (k4 < H) {
        ffma2(a[k4 / 2], nA, c01);
        ffma2(a[k4 / 2 + 1], nA, c23);
      }
      ffma2(b[k4 / 2], nB, c01);
      ffma2(b[k4 / 2 + 1], nB, c23);
    }
  }
  ok = __all_sync(FULL, ok);
  // backward substitution  x_c = z_c - (1 / d_c) sum_{i > c} W[i][c] x_i
  const int c0i = lane < NS - 1 ? lane : 0, c1i = lane + 32 < NS - 1 ? lane + 32 : 0;
  const float* c0 = Wc + colbase[c0i];
  const float* c1 = Wc + colbase[c1i];
  float t0 = 0.f, t1 = 0.f;
  x0 = 0.f;
  x1 = 0.f;
  __syncwarp();
  float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0;
#pragma unroll
  for (int i = NS - 1; i >= 0; --i) {
    if ((i & 3) == 3) {  // rows i-3 .. i of this lane's two columns (rows above a column's start are other columns' data:
      w0 = *reinterpret_cast<const float4*>(c0 + i - 3);  // they only reach accumulators that are already closed)
      w1 = *reinterpret_cast<const float4*>(c1 + i - 3);
    }
    const float xv = (i < 32) ? zk0 - rdk0 * t0 : zk1 - rdk1 * t1;
    const float xi = __shfl_sync(FULL, xv, i & 31);
    if (lane == (i & 31)) {
      if (i < 32) x0 = xi;
      else x1 = xi;
    }
    if (i > 0) {  // contributions to the columns still open (c < i); closed accumulators collect garbage that is never read
      const float e0 = (i & 3) == 3 ? w0.w : (i & 3) == 2 ? w0.z : (i & 3) == 1 ? w0.y : w0.x;
      const float e1 = (i & 3) == 3 ? w1.w : (i & 3) == 2 ? w1.z : (i & 3) == 1 ? w1.y : w1.x;
      t0 += e0 * xi;
      t1 += e1 * xi;
    }
  }
  return ok;
}

}  // namespace b2w
