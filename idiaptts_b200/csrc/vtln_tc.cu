// Neural-VTLN all-pass warp on the tensor cores, for the common case of a warping factor that is constant over long runs
// of rows (one alpha per speaker / utterance, BASELINE.json configs[4]).
//
// Same operator and reference call site as vtln.cu (AllPassWarp.forward, layers/AllPassWarp.py:148-173).  There the warp is
// an O(n^2) recursion per row in one thread's registers: 3600 dependent FMAs per 480 bytes of traffic at n = 60, i.e.
// CUDA-core bound at ~12 % of the HBM roofline.  Here a tile of 128 (row, block) units that share one alpha is a GEMM
//     Y[128 x n] = X'[128 x n] . B^T,   B[j][r] = S2_j A(alpha)[j][r] S1_r      (A = SPTK freqt matrix = W(alpha)^T)
// on tcgen05.mma kind::tf32 with the 3xTF32 split (fp32-accurate), accumulator in tensor memory.  The matrix is built in
// shared memory by one warp with a wavefront recursion (column r of A is M^r e0: A[j][r] = A[j-1][r-1] + alpha (A[j][r-1] -
// A[j-1][r])) and cached across the consecutive tiles a persistent CTA owns, so it is rebuilt only when alpha changes.
// Tiles whose units do not share one alpha are flagged and left to the recursion kernel of vtln.cu (second launch).
// The kernel is HBM-bound by design: 8 n + 4 bytes per unit, one coalesced read and one coalesced write of every tile.
#include "common.cuh"
#include "umma.cuh"

namespace b2w {

constexpr int kVtcF = 128;       // units per tile = UMMA M
constexpr int kVtcNP = 64;       // padded n (K and N of the GEMM)
constexpr int kVtcThreads = 256;
constexpr int kVtcStageStride = kVtcNP + 1;  // output staging row stride (floats): conflict-free scalar stores
constexpr uint32_t kVtcABytes = kVtcF * kVtcNP * 4;   // 32 KB, one of hi / lo
constexpr uint32_t kVtcBBytes = kVtcNP * kVtcNP * 4;  // 16 KB, one of hi / lo

struct VtcSmem {
  static constexpr uint32_t a_hi = 0, a_lo = kVtcABytes, b_hi = 2 * kVtcABytes, b_lo = 2 * kVtcABytes + kVtcBBytes;
  static constexpr uint32_t vec = 2 * kVtcABytes + 2 * kVtcBBytes;      // mean[64], std[64], 1/std[64]
  static constexpr uint32_t misc = vec + 3 * kVtcNP * 4;                // mbarrier, tmem slot
  static constexpr uint32_t total = misc + 64;
  static_assert(kVtcF * kVtcStageStride * 4 <= 2 * kVtcABytes, "output staging aliases the A tiles");
};

// One warp: writes B = S2 A(alpha) S1 (hi / lo TF32 tiles, K-major) for the n x n freqt matrix of `a`.
__device__ __forceinline__ void vtc_build_matrix(float* b_hi, float* b_lo, float a, int n) {
  const int lane = threadIdx.x & 31;
  const float bcoef = 1.f - a * a;
  float cur[2] = {0.f, 0.f}, nb2[2] = {0.f, 0.f};
  for (int d = 0; d <= 2 * n - 2; ++d) {
    // neighbours' values of the previous step (row j - 1, same column r)
    const float up0 = __shfl_up_sync(0xffffffffu, cur[0], 1);
    float up1 = __shfl_up_sync(0xffffffffu, cur[1], 1);
    const float wrap = __shfl_sync(0xffffffffu, cur[0], 31);
    if (lane == 0) up1 = wrap;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = lane + 32 * h, r = d - j;
      const float up = h ? up1 : up0;
      if (j < n && r >= 0 && r < n) {
        float v;
        if (r == 0) v = (j == 0) ? 1.f : 0.f;
        else if (j == 0) v = a * cur[h];
        else if (j == 1) v = fmaf(bcoef, nb2[h], a * cur[h]);
        else v = fmaf(a, cur[h] - up, nb2[h]);
        cur[h] = v;
        float s = v;
        if (j == 0) s *= 2.f;   // S2
        if (r == 0) s *= 0.5f;  // S1
        float hi, lo;
        umma::split_tf32(s, hi, lo);
        const uint32_t off = umma::tile_off(kVtcNP, j, r) / 4;
        b_hi[off] = hi;
        b_lo[off] = lo;
      }
      nb2[h] = up;  // A[j-1][r] becomes A[j-1][(r+1)-1] of the next step
    }
  }
}

__global__ void __launch_bounds__(kVtcThreads, 2)
allpass_tc_forward_kernel(const float* __restrict__ x, const float* __restrict__ alpha, int64_t units, int n, int blocks,
                          const float* __restrict__ mean, const float* __restrict__ std_dev, float* __restrict__ y,
                          uint8_t* __restrict__ tile_mixed, int64_t num_tiles) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* a_hi = reinterpret_cast<float*>(smem + VtcSmem::a_hi);
  float* a_lo = reinterpret_cast<float*>(smem + VtcSmem::a_lo);
  float* b_hi = reinterpret_cast<float*>(smem + VtcSmem::b_hi);
  float* b_lo = reinterpret_cast<float*>(smem + VtcSmem::b_lo);
  float* stage = reinterpret_cast<float*>(smem);  // aliases the A tiles once the MMAs have completed
  float* vmean = reinterpret_cast<float*>(smem + VtcSmem::vec);
  float* vstd = vmean + kVtcNP;
  float* vrstd = vstd + kVtcNP;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + VtcSmem::misc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // contiguous tile range of this CTA (matrix reuse across the tiles of one speaker)
  const int64_t per = (num_tiles + gridDim.x - 1) / gridDim.x;
  const int64_t t_begin = (int64_t)blockIdx.x * per;
  const int64_t t_end = min(num_tiles, t_begin + per);
  if (t_begin >= t_end) return;

  if (tid == 0) {
    umma::mbar_init(bar, 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, 64);
  for (int i = tid; i < (int)((2 * kVtcABytes + 2 * kVtcBBytes) / 4); i += kVtcThreads) reinterpret_cast<float*>(smem)[i] = 0.f;
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = umma::idesc_tf32(kVtcF, kVtcNP);
  const int nq = n / 4;                    // float4 chunks per unit
  constexpr int kPre = 8;                  // float4 chunks per thread (n <= 64: 128 * 16 / 256)
  // chunk ownership: item i of warp w covers rows 8 rg .. 8 rg + 7 and chunks 4 kg .. 4 kg + 3 with (rg, kg) = ((w + 8 i) / 4,
  // (w + 8 i) % 4); lane -> (row rg * 8 + (lane & 7), chunk 4 kg + (lane >> 3)).  A quarter warp then holds 8 distinct rows of
  // one K-chunk (conflict-free 16-byte stores into the K-major tile) while every row still contributes 64 contiguous bytes
  // to the global access (full 32-byte sectors).
  const int my_r = lane & 7, my_kq = lane >> 3;
  uint32_t phase = 0;
  float cached_alpha = 0.f;
  bool have_matrix = false;
  int cached_blk = -1;

  // raw tile rows travel global -> registers one tile ahead, so their latency hides behind the previous tile's GEMM / epilogue
  float4 pre[kPre];
  float pre_alpha = 0.f;
  auto prefetch = [&](int64_t t) {
    const int64_t u0 = t * kVtcF;
    const int nun = (int)min((int64_t)kVtcF, units - u0);
    const float4* src = reinterpret_cast<const float4*>(x + u0 * n);
#pragma unroll
    for (int i = 0; i < kPre; ++i) {
      const int item = warp + 8 * i;
      const int r = 8 * (item >> 2) + my_r, kq = 4 * (item & 3) + my_kq;
      pre[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nun && kq < nq) pre[i] = __ldg(src + r * nq + kq);
    }
    pre_alpha = (tid < nun) ? alpha[(u0 + tid) / blocks] : 0.f;
  };
  auto gemm = [&]() {  // one thread: D = A . B^T as 8 K-steps x 3 split products
    umma::tc_fence_after_sync();
    const uint32_t a_lbo = kVtcF * 16, b_lbo = kVtcNP * 16;
    umma::mma_3xtf32<kVtcNP / 8>(tmem, umma::smem_desc(umma::smem_u32(a_hi), a_lbo, 128), umma::smem_desc(umma::smem_u32(a_lo), a_lbo, 128),
                                 umma::smem_desc(umma::smem_u32(b_hi), b_lbo, 128), umma::smem_desc(umma::smem_u32(b_lo), b_lbo, 128),
                                 2 * a_lbo, 2 * b_lbo, idesc, false);
    umma::mma_commit(bar);
  };
  // Accumulator rows -> normalised output (warps 0-3 own TMEM lanes 32 w .. 32 w + 31 = tile rows).  Regular tiles go through
  // the staging buffer (which aliases the A tiles: every MMA reading A must be complete) and leave with coalesced stores;
  // the rare two-run tiles write rows [lo, hi) straight to global memory because their A tile is needed by a second GEMM.
  auto emit_rows = [&](bool direct, int lo, int hi, float* ydst) {
    if (!direct) {
      // all eight warps: warp w reads TMEM lane quarter w & 3 (its rows) and column half w >> 2
      const int row = 32 * (warp & 3) + lane;
      const int c0 = 32 * (warp >> 2);
      const uint32_t taddr = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + c0;
      float v[16];
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        umma::tmem_ld16(taddr + 16 * cb, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int c = c0 + 16 * cb + i;
          stage[row * kVtcStageStride + c] = (v[i] - vmean[c]) * vrstd[c];
        }
      }
    } else if (warp < 4) {
      const int row = 32 * warp + lane;
      const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16);
      float v[16];
#pragma unroll
      for (int cb = 0; cb < kVtcNP / 16; ++cb) {
        umma::tmem_ld16(taddr + 16 * cb, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = (v[i] - vmean[16 * cb + i]) * vrstd[16 * cb + i];
        if (!direct) {
#pragma unroll
          for (int i = 0; i < 16; ++i) stage[row * kVtcStageStride + 16 * cb + i] = v[i];
        } else if (row >= lo && row < hi) {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            if (16 * cb + i < n) *reinterpret_cast<float4*>(ydst + (int64_t)row * n + 16 * cb + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      }
    }
  };

  prefetch(t_begin);
  for (int64_t t = t_begin; t < t_end; ++t) {
    const int64_t u0 = t * kVtcF;
    const int nun = (int)min((int64_t)kVtcF, units - u0);
    // ---- alpha runs of this tile: one value, or two contiguous runs (a speaker boundary); anything else -> recursion kernel --
    const int64_t row0 = u0 / blocks;
    const float a0 = alpha[row0];
    const float a1 = alpha[(u0 + nun - 1) / blocks];
    const int blk0 = (int)(u0 - row0 * blocks);
    const bool norm_ok = blocks == 1 || (mean == nullptr && std_dev == nullptr);
    int s0 = nun;   // length of the first run if the tile is well formed
    bool ok = norm_ok;
    if (!__syncthreads_and(tid >= nun || pre_alpha == a0)) {  // not one alpha for the whole tile (rare): look for two runs
      s0 = __syncthreads_count(tid < nun && pre_alpha == a0);
      bool fine = true;
      if (tid < nun) fine = (tid < s0) ? (pre_alpha == a0) : (pre_alpha == a1);
      ok = __syncthreads_and(fine) && norm_ok;
    }
    if (tid == 0) tile_mixed[t] = ok ? 0 : 1;
    if (!ok) {
      if (t + 1 < t_end) prefetch(t + 1);
      continue;
    }
    // ---- normalisation vectors of this tile's block -----------------------------------------------------------------------
    bool rewritten = false;
    if (cached_blk != blk0) {
      if (tid < kVtcNP) {
        const bool in = tid < n;
        vmean[tid] = (in && mean) ? mean[blk0 * n + tid] : 0.f;
        const float sd = (in && std_dev) ? std_dev[blk0 * n + tid] : 1.f;
        vstd[tid] = sd;
        vrstd[tid] = 1.f / sd;
      }
      rewritten = true;
    }
    cached_blk = blk0;
    // ---- matrix of the first run --------------------------------------------------------------------------------------------
    if (!have_matrix || a0 != cached_alpha) {
      if (warp == 0) vtc_build_matrix(b_hi, b_lo, a0, n);
      cached_alpha = a0;
      have_matrix = true;
      rewritten = true;
    }
    if (rewritten) __syncthreads();  // vectors + matrix visible (block-uniform condition)
    // ---- A = hi / lo split of the de-normalised tile ----------------------------------------------------------------------------
    {
#pragma unroll
      for (int i = 0; i < kPre; ++i) {
        const int item = warp + 8 * i;
        const int r = 8 * (item >> 2) + my_r, kq = 4 * (item & 3) + my_kq;
        if (kq < nq) {
          const int k = 4 * kq;
          float4 v = pre[i];
          if (r < nun) {
            v.x = fmaf(v.x, vstd[k], vmean[k]);
            v.y = fmaf(v.y, vstd[k + 1], vmean[k + 1]);
            v.z = fmaf(v.z, vstd[k + 2], vmean[k + 2]);
            v.w = fmaf(v.w, vstd[k + 3], vmean[k + 3]);
          }
          float4 h, l;
          umma::split_tf32(v.x, h.x, l.x);
          umma::split_tf32(v.y, h.y, l.y);
          umma::split_tf32(v.z, h.z, l.z);
          umma::split_tf32(v.w, h.w, l.w);
          const uint32_t off = umma::tile_off(kVtcF, r, k) / 4;
          *reinterpret_cast<float4*>(a_hi + off) = h;
          *reinterpret_cast<float4*>(a_lo + off) = l;
        }
      }
    }
    // K padding (columns n .. 63): the output staging of the previous tile aliases the A tiles, keep the pad exactly zero
    for (int e = tid; e < kVtcF * (kVtcNP / 4 - nq); e += kVtcThreads) {
      const int r = e % kVtcF, k = n + 4 * (e / kVtcF);
      const uint32_t off = umma::tile_off(kVtcF, r, k) / 4;
      *reinterpret_cast<float4*>(a_hi + off) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(a_lo + off) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    umma::fence_proxy_async();
    umma::tc_fence_before_sync();
    __syncthreads();
    if (tid == 0) gemm();
    if (t + 1 < t_end) prefetch(t + 1);  // in flight during the GEMM, the epilogue and the stores
    umma::mbar_wait(bar, phase);
    phase ^= 1;
    umma::tc_fence_after_sync();
    if (s0 < nun) {
      // two runs (a speaker boundary inside the tile): rows of the first run leave directly, then the matrix of the second
      // run -- the one the following tiles need anyway -- is built and the GEMM repeated for the remaining rows
      emit_rows(true, 0, s0, y + u0 * n);
      umma::tc_fence_before_sync();
      __syncthreads();  // every warp has read the first accumulator
      if (warp == 0) vtc_build_matrix(b_hi, b_lo, a1, n);
      cached_alpha = a1;
      umma::fence_proxy_async();
      umma::tc_fence_before_sync();
      __syncthreads();
      if (tid == 0) gemm();
      umma::mbar_wait(bar, phase);
      phase ^= 1;
      umma::tc_fence_after_sync();
      emit_rows(true, s0, nun, y + u0 * n);
      umma::tc_fence_before_sync();
    } else {
      emit_rows(false, 0, nun, nullptr);  // the only MMAs that read the A tiles are complete (barrier waited by every thread)
      umma::tc_fence_before_sync();
      __syncthreads();
      float4* dst = reinterpret_cast<float4*>(y + u0 * n);
#pragma unroll
      for (int i = 0; i < kPre; ++i) {
        const int item = warp + 8 * i;
        const int r = 8 * (item >> 2) + my_r, kq = 4 * (item & 3) + my_kq;
        if (r < nun && kq < nq) {
          const float* sp = stage + r * kVtcStageStride + 4 * kq;
          dst[r * nq + kq] = make_float4(sp[0], sp[1], sp[2], sp[3]);
        }
      }
    }
    __syncthreads();  // staging (= A tiles) is rewritten by the next tile
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 64);
}

// ---- backward -------------------------------------------------------------------------------------------------------------------
// gx = (gy / std) . Bb^T . std,  Bb[r][j] = S1_r A[j][r] S2_j   (the transposed forward matrix)
// galpha(unit) = < gy / std , X' . Bt^T >,  Bt[j][r] = S2_j dA/dalpha[j][r] S1_r   (tangent of the same wavefront recursion)
// Two GEMMs per tile into two accumulators; tiles whose units do not share ONE alpha go to the recursion kernel.
struct VtcBwdSmem {
  static constexpr uint32_t g_hi = 0, g_lo = kVtcABytes, x_hi = 2 * kVtcABytes, x_lo = 3 * kVtcABytes;
  static constexpr uint32_t bb_hi = 4 * kVtcABytes, bb_lo = bb_hi + kVtcBBytes, bt_hi = bb_lo + kVtcBBytes, bt_lo = bt_hi + kVtcBBytes;
  static constexpr uint32_t vec = bt_lo + kVtcBBytes;
  static constexpr uint32_t misc = vec + 3 * kVtcNP * 4;
  static constexpr uint32_t total = misc + 64;
};

__device__ __forceinline__ void vtc_build_matrices_bwd(float* bb_hi, float* bb_lo, float* bt_hi, float* bt_lo, float a, int n) {
  const int lane = threadIdx.x & 31;
  const float bcoef = 1.f - a * a;
  float ce[2] = {0.f, 0.f}, ct[2] = {0.f, 0.f}, ne[2] = {0.f, 0.f}, nt[2] = {0.f, 0.f};
  for (int d = 0; d <= 2 * n - 2; ++d) {
    const float ue0 = __shfl_up_sync(0xffffffffu, ce[0], 1), ut0 = __shfl_up_sync(0xffffffffu, ct[0], 1);
    float ue1 = __shfl_up_sync(0xffffffffu, ce[1], 1), ut1 = __shfl_up_sync(0xffffffffu, ct[1], 1);
    const float we = __shfl_sync(0xffffffffu, ce[0], 31), wt = __shfl_sync(0xffffffffu, ct[0], 31);
    if (lane == 0) { ue1 = we; ut1 = wt; }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = lane + 32 * h, r = d - j;
      const float ue = h ? ue1 : ue0, ut = h ? ut1 : ut0;
      if (j < n && r >= 0 && r < n) {
        float e, t;
        if (r == 0) { e = (j == 0) ? 1.f : 0.f; t = 0.f; }
        else if (j == 0) { e = a * ce[h]; t = fmaf(a, ct[h], ce[h]); }
        else if (j == 1) { e = fmaf(bcoef, ne[h], a * ce[h]); t = fmaf(-2.f * a, ne[h], fmaf(bcoef, nt[h], fmaf(a, ct[h], ce[h]))); }
        else { const float diff = ce[h] - ue; e = fmaf(a, diff, ne[h]); t = nt[h] + diff + a * (ct[h] - ut); }
        ce[h] = e;
        ct[h] = t;
        float sc = 1.f;
        if (j == 0) sc *= 2.f;
        if (r == 0) sc *= 0.5f;
        float hi, lo;
        umma::split_tf32(e * sc, hi, lo);
        uint32_t off = umma::tile_off(kVtcNP, r, j) / 4;   // Bb[n = r][k = j]
        bb_hi[off] = hi;
        bb_lo[off] = lo;
        umma::split_tf32(t * sc, hi, lo);
        off = umma::tile_off(kVtcNP, j, r) / 4;            // Bt[n = j][k = r]
        bt_hi[off] = hi;
        bt_lo[off] = lo;
      }
      ne[h] = ue;
      nt[h] = ut;
    }
  }
}

__global__ void __launch_bounds__(kVtcThreads, 1)
allpass_tc_backward_kernel(const float* __restrict__ gy, const float* __restrict__ x, const float* __restrict__ alpha, int64_t units, int n,
                           int blocks, const float* __restrict__ mean, const float* __restrict__ std_dev, float* __restrict__ gx,
                           float* __restrict__ galpha_unit, uint8_t* __restrict__ tile_mixed, int64_t num_tiles) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* g_hi = reinterpret_cast<float*>(smem + VtcBwdSmem::g_hi);
  float* g_lo = reinterpret_cast<float*>(smem + VtcBwdSmem::g_lo);
  float* x_hi = reinterpret_cast<float*>(smem + VtcBwdSmem::x_hi);
  float* x_lo = reinterpret_cast<float*>(smem + VtcBwdSmem::x_lo);
  float* bb_hi = reinterpret_cast<float*>(smem + VtcBwdSmem::bb_hi);
  float* bb_lo = reinterpret_cast<float*>(smem + VtcBwdSmem::bb_lo);
  float* bt_hi = reinterpret_cast<float*>(smem + VtcBwdSmem::bt_hi);
  float* bt_lo = reinterpret_cast<float*>(smem + VtcBwdSmem::bt_lo);
  float* stage = x_hi;  // gx staging aliases the X tiles once both GEMMs have completed (the G tiles stay intact for the dot)
  float* vmean = reinterpret_cast<float*>(smem + VtcBwdSmem::vec);
  float* vstd = vmean + kVtcNP;
  float* vrstd = vstd + kVtcNP;
  __shared__ float ga_half[kVtcF];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + VtcBwdSmem::misc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t per = (num_tiles + gridDim.x - 1) / gridDim.x;
  const int64_t t_begin = (int64_t)blockIdx.x * per;
  const int64_t t_end = min(num_tiles, t_begin + per);
  if (t_begin >= t_end) return;
  if (tid == 0) {
    umma::mbar_init(bar, 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, 128);
  for (int i = tid; i < (int)(VtcBwdSmem::vec / 4); i += kVtcThreads) reinterpret_cast<float*>(smem)[i] = 0.f;
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = umma::idesc_tf32(kVtcF, kVtcNP);
  const int nq = n / 4;
  constexpr int kPre = 8;
  const int my_r = lane & 7, my_kq = lane >> 3;
  uint32_t phase = 0;
  float cached_alpha = 0.f;
  bool have_matrix = false;
  int cached_blk = -1;
  float4 pg[kPre], px[kPre];
  float pre_alpha = 0.f;
  auto prefetch = [&](int64_t t) {
    const int64_t u0 = t * kVtcF;
    const int nun = (int)min((int64_t)kVtcF, units - u0);
    const float4* sg = reinterpret_cast<const float4*>(gy + u0 * n);
    const float4* sx = reinterpret_cast<const float4*>(x + u0 * n);
#pragma unroll
    for (int i = 0; i < kPre; ++i) {
      const int item = warp + 8 * i;
      const int r = 8 * (item >> 2) + my_r, kq = 4 * (item & 3) + my_kq;
      pg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      px[i] = pg[i];
      if (r < nun && kq < nq) {
        pg[i] = __ldg(sg + r * nq + kq);
        px[i] = __ldg(sx + r * nq + kq);
      }
    }
    pre_alpha = (tid < nun) ? alpha[(u0 + tid) / blocks] : 0.f;
  };
  auto put = [&](float* hi_t, float* lo_t, int r, int k, float4 v) {
    float4 h, l;
    umma::split_tf32(v.x, h.x, l.x);
    umma::split_tf32(v.y, h.y, l.y);
    umma::split_tf32(v.z, h.z, l.z);
    umma::split_tf32(v.w, h.w, l.w);
    const uint32_t off = umma::tile_off(kVtcF, r, k) / 4;
    *reinterpret_cast<float4*>(hi_t + off) = h;
    *reinterpret_cast<float4*>(lo_t + off) = l;
  };

  prefetch(t_begin);
  for (int64_t t = t_begin; t < t_end; ++t) {
    const int64_t u0 = t * kVtcF;
    const int nun = (int)min((int64_t)kVtcF, units - u0);
    const int64_t row0 = u0 / blocks;
    const float a0 = alpha[row0];
    const int blk0 = (int)(u0 - row0 * blocks);
    const bool norm_ok = blocks == 1 || (mean == nullptr && std_dev == nullptr);
    const bool ok = __syncthreads_and(tid >= nun || pre_alpha == a0) && norm_ok;
    if (tid == 0) tile_mixed[t] = ok ? 0 : 1;
    if (!ok) {
      if (t + 1 < t_end) prefetch(t + 1);
      continue;
    }
    bool rewritten = false;
    if (cached_blk != blk0) {
      if (tid < kVtcNP) {
        const bool in = tid < n;
        vmean[tid] = (in && mean) ? mean[blk0 * n + tid] : 0.f;
        const float sd = (in && std_dev) ? std_dev[blk0 * n + tid] : 1.f;
        vstd[tid] = sd;
        vrstd[tid] = 1.f / sd;
      }
      rewritten = true;
    }
    cached_blk = blk0;
    if (!have_matrix || a0 != cached_alpha) {
      if (warp == 0) vtc_build_matrices_bwd(bb_hi, bb_lo, bt_hi, bt_lo, a0, n);
      cached_alpha = a0;
      have_matrix = true;
      rewritten = true;
    }
    if (rewritten) __syncthreads();  // block-uniform condition
#pragma unroll
    for (int i = 0; i < kPre; ++i) {
      const int item = warp + 8 * i;
      const int r = 8 * (item >> 2) + my_r, kq = 4 * (item & 3) + my_kq;
      if (kq < nq) {
        const int k = 4 * kq;
        float4 g = pg[i], v = px[i];
        if (r < nun) {
          g.x *= vrstd[k]; g.y *= vrstd[k + 1]; g.z *= vrstd[k + 2]; g.w *= vrstd[k + 3];
          v.x = fmaf(v.x, vstd[k], vmean[k]);
          v.y = fmaf(v.y, vstd[k + 1], vmean[k + 1]);
          v.z = fmaf(v.z, vstd[k + 2], vmean[k + 2]);
          v.w = fmaf(v.w, vstd[k + 3], vmean[k + 3]);
        }
        put(g_hi, g_lo, r, k, g);
        put(x_hi, x_lo, r, k, v);
      }
    }
    // K padding of the X tiles (the gx staging of the previous tile aliases them); the G tiles' padding is never written
    for (int e = tid; e < kVtcF * (kVtcNP / 4 - nq); e += kVtcThreads) {
      const int r = e % kVtcF, k = n + 4 * (e / kVtcF);
      const uint32_t off = umma::tile_off(kVtcF, r, k) / 4;
      *reinterpret_cast<float4*>(x_hi + off) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(x_lo + off) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    umma::fence_proxy_async();
    umma::tc_fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      umma::tc_fence_after_sync();
      const uint32_t a_lbo = kVtcF * 16, b_lbo = kVtcNP * 16;
      umma::mma_3xtf32<kVtcNP / 8>(tmem, umma::smem_desc(umma::smem_u32(g_hi), a_lbo, 128), umma::smem_desc(umma::smem_u32(g_lo), a_lbo, 128),
                                   umma::smem_desc(umma::smem_u32(bb_hi), b_lbo, 128), umma::smem_desc(umma::smem_u32(bb_lo), b_lbo, 128),
                                   2 * a_lbo, 2 * b_lbo, idesc, false);
      umma::mma_3xtf32<kVtcNP / 8>(tmem + 64, umma::smem_desc(umma::smem_u32(x_hi), a_lbo, 128), umma::smem_desc(umma::smem_u32(x_lo), a_lbo, 128),
                                   umma::smem_desc(umma::smem_u32(bt_hi), b_lbo, 128), umma::smem_desc(umma::smem_u32(bt_lo), b_lbo, 128),
                                   2 * a_lbo, 2 * b_lbo, idesc, false);
      umma::mma_commit(bar);
    }
    if (t + 1 < t_end) prefetch(t + 1);
    umma::mbar_wait(bar, phase);
    phase ^= 1;
    umma::tc_fence_after_sync();
    {
      // all eight warps: warp w owns TMEM lane quarter w & 3 (its rows) and column half w >> 2; the two halves of a row's
      // d alpha dot product meet in shared memory
      const int row = 32 * (warp & 3) + lane;
      const int c0 = 32 * (warp >> 2);
      const uint32_t taddr = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + c0;
      float v[16];
      float ga = 0.f;
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        umma::tmem_ld16(taddr + 16 * cb, v);  // D1: gradient w.r.t. the de-normalised input
#pragma unroll
        for (int i = 0; i < 16; ++i) stage[row * kVtcStageStride + c0 + 16 * cb + i] = v[i] * vstd[c0 + 16 * cb + i];
        umma::tmem_ld16(taddr + 64 + 16 * cb, v);  // D2: d y / d alpha
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const uint32_t off = umma::tile_off(kVtcF, row, c0 + 16 * cb + i) / 4;
          const float4 gh = *reinterpret_cast<const float4*>(g_hi + off), gl = *reinterpret_cast<const float4*>(g_lo + off);
          ga = fmaf(gh.x + gl.x, v[i], ga);
          ga = fmaf(gh.y + gl.y, v[i + 1], ga);
          ga = fmaf(gh.z + gl.z, v[i + 2], ga);
          ga = fmaf(gh.w + gl.w, v[i + 3], ga);
        }
      }
      if (warp >= 4) ga_half[row] = ga;
      umma::tc_fence_before_sync();
      __syncthreads();
      if (warp < 4 && row < nun) galpha_unit[u0 + row] = ga + ga_half[row];
    }
    {
      float4* dst = reinterpret_cast<float4*>(gx + u0 * n);
#pragma unroll
      for (int i = 0; i < kPre; ++i) {
        const int item = warp + 8 * i;
        const int r = 8 * (item >> 2) + my_r, kq = 4 * (item & 3) + my_kq;
        if (r < nun && kq < nq) {
          const float* sp = stage + r * kVtcStageStride + 4 * kq;
          dst[r * nq + kq] = make_float4(sp[0], sp[1], sp[2], sp[3]);
        }
      }
    }
    __syncthreads();
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 128);
}

}  // namespace b2w

// second launch: the recursion kernel of vtln.cu restricted to the flagged tiles
extern "C" int b2w_allpass_forward_masked(const float* x, const float* alpha, int64_t rows, int32_t n, int32_t blocks, const float* mean,
                                          const float* std_dev, float* y, const uint8_t* tile_mask, void* stream);

extern "C" int b2w_allpass_forward_tc(const float* x, const float* alpha, int64_t rows, int32_t n, int32_t blocks, const float* mean,
                                      const float* std_dev, float* y, uint8_t* tile_flags, void* stream) {
  using namespace b2w;
  B2W_REQUIRE(x && alpha && y && tile_flags, "b2w_allpass_forward_tc: null argument");
  B2W_REQUIRE(n >= 4 && n <= kVtcNP && n % 4 == 0 && blocks >= 1,
              "b2w_allpass_forward_tc: n %d must be a multiple of 4 in [4, 64] (use b2w_allpass_forward)", n);
  B2W_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
              "b2w_allpass_forward_tc: x / y must be 16-byte aligned");
  if (rows == 0) return 0;
  const int64_t units = rows * blocks;
  const int64_t num_tiles = (units + kVtcF - 1) / kVtcF;
  const int grid = (int)(num_tiles < 2 * 148 ? num_tiles : 2 * 148);
  cudaStream_t st = (cudaStream_t)stream;
  cudaFuncSetAttribute(allpass_tc_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VtcSmem::total);
  allpass_tc_forward_kernel<<<grid, kVtcThreads, VtcSmem::total, st>>>(x, alpha, units, n, blocks, mean, std_dev, y, tile_flags, num_tiles);
  int rc = check_launch("allpass_tc_forward_kernel");
  if (rc) return rc;
  return b2w_allpass_forward_masked(x, alpha, rows, n, blocks, mean, std_dev, y, tile_flags, stream);
}

extern "C" int b2w_allpass_backward_masked(const float* grad_y, const float* x, const float* alpha, int64_t rows, int32_t n, int32_t blocks,
                                           const float* mean, const float* std_dev, float* grad_x, float* grad_alpha, float* unit_workspace,
                                           const uint8_t* tile_mask, void* stream);

extern "C" int b2w_allpass_backward_tc(const float* grad_y, const float* x, const float* alpha, int64_t rows, int32_t n, int32_t blocks,
                                       const float* mean, const float* std_dev, float* grad_x, float* grad_alpha, float* unit_workspace,
                                       uint8_t* tile_flags, void* stream) {
  using namespace b2w;
  B2W_REQUIRE(grad_y && x && alpha && grad_x && grad_alpha && unit_workspace && tile_flags, "b2w_allpass_backward_tc: null argument");
  B2W_REQUIRE(n >= 4 && n <= kVtcNP && n % 4 == 0 && blocks >= 1,
              "b2w_allpass_backward_tc: n %d must be a multiple of 4 in [4, 64] (use b2w_allpass_backward)", n);
  B2W_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(grad_y) | reinterpret_cast<uintptr_t>(grad_x)) & 15) == 0,
              "b2w_allpass_backward_tc: grad_y / x / grad_x must be 16-byte aligned");
  if (rows == 0) return 0;
  const int64_t units = rows * blocks;
  const int64_t num_tiles = (units + kVtcF - 1) / kVtcF;
  const int grid = (int)(num_tiles < 148 ? num_tiles : 148);
  cudaStream_t st = (cudaStream_t)stream;
  cudaFuncSetAttribute(allpass_tc_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VtcBwdSmem::total);
  allpass_tc_backward_kernel<<<grid, kVtcThreads, VtcBwdSmem::total, st>>>(grad_y, x, alpha, units, n, blocks, mean, std_dev, grad_x,
                                                                           unit_workspace, tile_flags, num_tiles);
  int rc = check_launch("allpass_tc_backward_kernel");
  if (rc) return rc;
  return b2w_allpass_backward_masked(grad_y, x, alpha, rows, n, blocks, mean, std_dev, grad_x, grad_alpha, unit_workspace, tile_flags, stream);
}
