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
constexpr uint32_t kVtcBBytes = kVtcNP * kVtcNP * 4;  // 16 KB, one of hi / lo


// One warp: writes B = S2 A(alpha) S1 (hi / lo TF32 tiles, K-major) for the n x n freqt matrix of `a`.
// Row by row: with (cn, cu) = (1 - a^2, 0) for row 1 and (1, 1) below it, the freqt recursion
//     A[j][r] = cn A[j-1][r-1] + a (A[j][r-1] - cu A[j-1][r]),     A[j][0] = (j == 0),   A[0][r] = a^r
// is, for a fixed row, a first-order linear recurrence over r with the CONSTANT coefficient a and a forcing term that only needs the
// previous row: A[j][r] = a A[j][r-1] + g_r.  Lane l owns columns 2 l and 2 l + 1; a five-step warp scan over affine maps (the
// multipliers are powers of a^2) resolves the recurrence, so a row costs ~40 instructions instead of a 119-step wavefront.
// Every value is split into hi / lo TF32 as it is stored (side effects, off the dependent chain of the recurrence).
__device__ __forceinline__ void vtc_build_matrix(float* b_hi, float* b_lo, float a, int n) {
  const int lane = threadIdx.x & 31;
  const int r0 = 2 * lane;
  float pw[5];  // a^2, a^4, a^8, a^16, a^32
  pw[0] = a * a;
#pragma unroll
  for (int i = 1; i < 5; ++i) pw[i] = pw[i - 1] * pw[i - 1];
  // row 0: a^r
  float p0 = powf(fabsf(a), (float)r0);
  if (a == 0.f) p0 = r0 == 0 ? 1.f : 0.f;   // (a < 0: even powers only, the odd one follows by multiplication)
  float p1 = p0 * a;
  const uint32_t cb = (uint32_t)(r0 >> 2) * (kVtcNP * 4) + (r0 & 3);  // float offset of column r0 inside a row of the K-major tile
  const bool in = r0 < n;  // n is even: both columns of a lane are inside or outside
  float left = __shfl_up_sync(0xffffffffu, p1, 1);            // A[j-1][r0 - 1] for the next row (0 for column -1)
  if (lane == 0) left = 0.f;
  for (int j = 0; j < n; ++j) {
    if (j > 0) {
      const float cn = j == 1 ? 1.f - a * a : 1.f, cu = j == 1 ? 0.f : a;
      const float g0 = lane == 0 ? 0.f : fmaf(cn, left, -cu * p0);   // column 0 of rows >= 1 is zero, and so is its forcing term
      const float g1 = fmaf(cn, p0, -cu * p1);
      // this lane's pair as an affine map of the incoming carry c = A[j][r0 - 1]:  A[j][r0] = a c + g0,  A[j][r1] = a^2 c + x
      float x = fmaf(a, g0, g1);
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const float t = __shfl_up_sync(0xffffffffu, x, 1 << i);
        if (lane >= (1 << i)) x = fmaf(pw[i], t, x);
      }
      float carry = __shfl_up_sync(0xffffffffu, x, 1);        // A[j][r0 - 1]: the carry of this row and `left` of the next one
      if (lane == 0) carry = 0.f;
      p0 = lane == 0 ? 0.f : fmaf(a, carry, g0);
      p1 = x;
      left = carry;
    }
    if (in) {  // split and store (off the dependent chain of the recurrence)
      const float s2 = j == 0 ? 2.f : 1.f;                    // S2; S1 halves column 0
      const uint32_t off = umma::tile_off(kVtcNP, j, 0) / 4 + cb;
      float2 hi, lo;
      umma::split_tf32(p0 * s2 * (lane == 0 ? 0.5f : 1.f), hi.x, lo.x);
      umma::split_tf32(p1 * s2, hi.y, lo.y);
      *reinterpret_cast<float2*>(b_hi + off) = hi;
      *reinterpret_cast<float2*>(b_lo + off) = lo;
    }
  }
  __syncwarp();  // (entries outside n x n stay zero: the buffers are cleared once)
}

// ---- forward: a persistent, warp-specialised pipeline ------------------------------------------------------------------------
// One CTA per SM walks a contiguous range of 128-unit tiles (the matrix is reused while alpha stays the same).  A tile of raw rows
// is ONE contiguous block of 128 n floats, so it travels as a single bulk async copy each way.  Nobody executes a CTA barrier
// inside the tile loop; every hand-over is an mbarrier:
//   producer warp          global -> raw stage ring (3 x 32 KB); it also classifies the tile (one alpha / two runs / mixed) and
//                          publishes that with the stage, so the other roles never wait for a global load inside the loop
//   builder warp           walks the CTA's tiles AHEAD of everybody else and builds B = S2 A(alpha) S1 for every new alpha into
//                          one of two matrix buffers (119 dependent wavefront steps: far too long to sit in the tile loop)
//   converter / epilogue   sixteen warps, a thread owns a quarter row (its TMEM lane, 16 columns), software pipelined:
//                            tile i:     raw row -> de-normalise -> hi / lo TF32 -> A operand in TENSOR MEMORY (tcgen05.st)
//                            tile i - 1: accumulator row (tcgen05.ld) -> normalise -> raw output stage
//                          so the MMA of tile i runs under the epilogue of tile i - 1 (A and D are double buffered in TMEM)
//   issuer warp            issues the 24 TS-form MMAs of a tile under elect.sync, switches matrix buffers when alpha changes
//   store thread           one bulk store per tile from the output stage
// Barrier parities are stateless (tile i uses completion i / slots of the barrier of slot i % slots): every tile, mixed or not,
// completes every barrier of its slots exactly once.  Matrix buffers: build k goes to buffer k & 1, full[k & 1] / free[k & 1].
constexpr int kVtfThreads = 640;          // warp 0 issuer, 1 producer, 2 builder, 3 store, 4-19 converter / epilogue
constexpr int kVtfNST = 3;                // raw input stages
constexpr int kVtfGW = 16;                // converter / epilogue warps
constexpr uint32_t kVtfStageBytes = kVtcF * kVtcNP * 4;  // 32 KB (n <= 64)
struct VtfSmem {
  static constexpr uint32_t in = 0;                                   // [NST][32 KB]
  static constexpr uint32_t out = kVtfNST * kVtfStageBytes;           // [2][32 KB]
  static constexpr uint32_t bmat = out + 2 * kVtfStageBytes;          // [2][hi 16 KB | lo 16 KB]
  static constexpr uint32_t vec = bmat + 4 * kVtcBBytes;              // mean[64], std[64], 1/std[64]
  static constexpr uint32_t bars = vec + 3 * kVtcNP * 4;              // see the kernel
  static constexpr uint32_t nbars = 2 * kVtfNST + 12;
  static constexpr uint32_t misc = bars + nbars * 8;                  // tmem slot, tile info ring
  static constexpr uint32_t total = misc + 16 + 8 * 32;
  static_assert(misc % 16 == 0, "tile info ring is read as int4");
};
constexpr int kVtfTmA = 0;     // A[b] hi at 128 b, lo at 128 b + 64
constexpr int kVtfTmD = 256;   // D[b] at 256 + 64 b

__device__ long long g_vtf_prof[16];
#ifdef B2W_VTF_PROF
#define VPROF_DECL long long vp_t = clock64(), vp_acc[16] = {0}; const long long vp_t0 = vp_t; const bool vp_on = blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 128)
#define VPROF_LAP(i) do { if (vp_on) { const long long n_ = clock64(); vp_acc[i] += n_ - vp_t; vp_t = n_; } } while (0)
#define VPROF_FLUSH(lo, hi) do { if (vp_on) { for (int i_ = lo; i_ < hi; ++i_) g_vtf_prof[i_] = vp_acc[i_]; } } while (0)
#else
#define VPROF_DECL
#define VPROF_LAP(i)
#define VPROF_FLUSH(lo, hi)
#endif

// Alpha runs of a tile, evaluated by one warp: ok = one alpha, or two contiguous runs (a speaker boundary); n0 = units of the
// first run.  `ar` holds the alphas of units lane + 32 h, loaded by the caller two tiles ahead.
struct VtfRuns {
  float a0, a1;
  int n0;
  bool ok;
};
__device__ __forceinline__ VtfRuns vtf_runs(const float (&ar)[4], float a0, float a1, int nun, int lane) {
  VtfRuns r;
  r.a0 = a0;
  r.a1 = a1;
  r.n0 = 0;
#pragma unroll
  for (int h = 0; h < 4; ++h) r.n0 += __popc(__ballot_sync(0xffffffffu, lane + 32 * h < nun && ar[h] == a0));
  bool fine = true;
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const int u = lane + 32 * h;
    if (u < nun) fine = fine && (u < r.n0 ? ar[h] == a0 : ar[h] == a1);
  }
  r.ok = __all_sync(0xffffffffu, fine);
  return r;
}

__global__ void __launch_bounds__(kVtfThreads, 1)
allpass_tc_forward_kernel(const float* __restrict__ x, const float* __restrict__ alpha, int64_t units, int n, int blocks,
                          const float* __restrict__ mean, const float* __restrict__ std_dev, float* __restrict__ y,
                          uint8_t* __restrict__ tile_mixed, int64_t num_tiles) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* vmean = reinterpret_cast<float*>(smem + VtfSmem::vec);
  float* vstd = vmean + kVtcNP;
  float* vrstd = vstd + kVtcNP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + VtfSmem::bars);
  uint64_t* bar_full = bars;                    // [NST] raw tile + its classification have landed (bytes + one arrival)
  uint64_t* bar_empty = bar_full + kVtfNST;     // [NST] every converter warp has read the raw tile
  uint64_t* bar_afull = bar_empty + kVtfNST;    // [2] A[b] written (16 warps)
  uint64_t* bar_dfull = bar_afull + 2;          // [2] the MMAs of the tile in D[b] have completed
  uint64_t* bar_ofull = bar_dfull + 2;          // [2] output stage [b] written (16 warps)
  uint64_t* bar_ofree = bar_ofull + 2;          // [2] the bulk store has read output stage [b]
  uint64_t* bar_bfull = bar_ofree + 2;          // [2] matrix buffer [k & 1] holds build k
  uint64_t* bar_bfree = bar_bfull + 2;          // [2] every MMA that read matrix buffer [k & 1] has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + VtfSmem::misc);
  int4* tinfo = reinterpret_cast<int4*>(smem + VtfSmem::misc + 16);  // [8] (ok, n0, nun, -) of tile i at i & 7, then [8] (alpha of the first / last unit)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t per = (num_tiles + gridDim.x - 1) / gridDim.x;
  const int64_t t_begin = (int64_t)blockIdx.x * per;
  const int64_t t_end = min(num_tiles, t_begin + per);
  if (t_begin >= t_end) return;
  const int cnt = (int)(t_end - t_begin);

  if (tid == 0) {
    for (int i = 0; i < kVtfNST; ++i) {
      umma::mbar_init(&bar_full[i], 2);
      umma::mbar_init(&bar_empty[i], kVtfGW);
    }
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(&bar_afull[i], kVtfGW);
      umma::mbar_init(&bar_dfull[i], 1);
      umma::mbar_init(&bar_ofull[i], kVtfGW);
      umma::mbar_init(&bar_ofree[i], 1);
      umma::mbar_init(&bar_bfull[i], 1);
      umma::mbar_init(&bar_bfree[i], 1);
    }
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
  // one set of normalisation vectors serves the whole call (tiles of other shapes are left to the recursion kernel)
  const bool norm_ok = blocks == 1 || (mean == nullptr && std_dev == nullptr);
  const bool has_norm = mean != nullptr || std_dev != nullptr;
  if (tid < kVtcNP) {
    const bool in = tid < n && blocks == 1;
    vmean[tid] = (in && mean) ? mean[tid] : 0.f;
    const float sd = (in && std_dev) ? std_dev[tid] : 1.f;
    vstd[tid] = sd;
    vrstd[tid] = 1.f / sd;
  }
  for (int i = tid; i < (int)(4 * kVtcBBytes / 4); i += kVtfThreads) reinterpret_cast<float*>(smem + VtfSmem::bmat)[i] = 0.f;
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const int nq = n / 4;
  const uint32_t row_bytes = (uint32_t)n * 4u;
  auto wait_slot = [&](uint64_t* arr, int slots, int i) { umma::mbar_wait(&arr[i % slots], (uint32_t)(i / slots) & 1u); };
  auto alpha_of = [&](int64_t u) { return blocks == 1 ? alpha[u] : alpha[u / blocks]; };  // (no 64-bit division in the common case)
  auto tile_units = [&](int i) { return (int)min((int64_t)kVtcF, units - (t_begin + i) * kVtcF); };
  // A scanning warp keeps the alphas of the next two tiles in registers (units lane + 32 h, first and last unit): a dependent
  // global load inside the loop waits behind megabytes of queued bulk traffic (~2 k cycles).
  float run_ar0[4], run_ar1[4], run_a00 = 0.f, run_a01 = 0.f, run_a10 = 0.f, run_a11 = 0.f;  // (named slots: no dynamic register indexing)
  auto load_into = [&](int i, float (&ar)[4], float& a0, float& a1) {
    const int64_t u0 = (t_begin + i) * kVtcF;
    const int nun = tile_units(i);
    a0 = alpha_of(u0);
    a1 = alpha_of(u0 + nun - 1);
#pragma unroll
    for (int h = 0; h < 4; ++h) ar[h] = lane + 32 * h < nun ? alpha_of(u0 + lane + 32 * h) : 0.f;
  };
  auto load_runs = [&](int i) {
    if (i >= cnt) return;
    if (i & 1) load_into(i, run_ar1, run_a01, run_a11);
    else load_into(i, run_ar0, run_a00, run_a10);
  };
  auto runs_of = [&](int i) {  // classification of tile i; refills the slot with tile i + 2
    VtfRuns r;
    if (i & 1) r = vtf_runs(run_ar1, run_a01, run_a11, tile_units(i), lane);
    else r = vtf_runs(run_ar0, run_a00, run_a10, tile_units(i), lane);
    load_runs(i + 2);
    r.ok = r.ok && norm_ok;
    return r;
  };
  VPROF_DECL;

  if (warp == 1) {
    // ---- producer -----------------------------------------------------------------------------------------------------------
    // A tile with two alpha runs (a speaker boundary) is emitted as two partial segments, so everybody downstream only ever sees
    // segments with ONE alpha (rows beyond the segment's length are zero / ignored, like in the last tile).  (ok, last, units,
    // first unit relative to the CTA's range) and the alpha travel with the stage.
    load_runs(0);
    load_runs(1);
    int sidx = 0;
    for (int i = 0; i < cnt; ++i) {
      const int nun = tile_units(i);
      const VtfRuns r = runs_of(i);
      const int nseg = (r.ok && r.n0 < nun) ? 2 : 1;
      for (int sg = 0; sg < nseg; ++sg, ++sidx) {
        if (lane == 0) {
          const int s = sidx % kVtfNST;
          const int off = sg ? r.n0 : 0, cu = nseg == 2 ? (sg ? nun - r.n0 : r.n0) : nun;
          const int ustart = i * kVtcF + off;
          if (sidx >= kVtfNST) wait_slot(bar_empty, kVtfNST, sidx - kVtfNST);
          umma::mbar_expect_tx(&bar_full[s], (uint32_t)cu * row_bytes);
          umma::bulk_g2s(smem + VtfSmem::in + s * kVtfStageBytes, x + (t_begin * kVtcF + ustart) * n, (uint32_t)cu * row_bytes, &bar_full[s]);
          tinfo[sidx & 7] = make_int4(r.ok ? 1 : 0, (i == cnt - 1 && sg == nseg - 1) ? 1 : 0, cu, ustart);  // (slot last read 8 segments ago)
          tinfo[8 + (sidx & 7)] = make_int4(__float_as_int(sg ? r.a1 : r.a0), 0, 0, 0);
          umma::mbar_arrive(&bar_full[s]);
        }
      }
      if (lane == 0) tile_mixed[t_begin + i] = r.ok ? 0 : 1;
      __syncwarp();
    }
  } else if (warp == 2) {
    // ---- builder: the sequence of matrices the issuer will ask for, built ahead into alternating buffers ------------------------
    float cached = 0.f;
    bool have = false;
    int k = 0;
    auto build = [&](float a) {
      if (k >= 2) umma::mbar_wait(&bar_bfree[k & 1], (uint32_t)((k >> 1) - 1) & 1u);
      float* bh = reinterpret_cast<float*>(smem + VtfSmem::bmat + (k & 1) * 2 * kVtcBBytes);
#ifdef B2W_VTF_PROF
      const long long tb0 = clock64();
#endif
      vtc_build_matrix(bh, bh + kVtcBBytes / 4, a, n);
#ifdef B2W_VTF_PROF
      const long long tb1 = clock64();
#endif
      umma::fence_proxy_async();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&bar_bfull[k & 1]);
#ifdef B2W_VTF_PROF
      if (blockIdx.x == 0 && lane == 0) { g_vtf_prof[4] += tb1 - tb0; g_vtf_prof[7] += clock64() - tb1; g_vtf_prof[5] += 1; if (k == 0) g_vtf_prof[6] = tb0 - vp_t0; }
#endif
      ++k;
      cached = a;
      have = true;
    };
    build(alpha_of(t_begin * kVtcF));  // the matrix of the very first unit, before anything is classified (the issuer mirrors this)
    load_runs(0);
    load_runs(1);
    for (int i = 0; i < cnt; ++i) {
      const int nun = tile_units(i);
      const VtfRuns r = runs_of(i);
      if (!r.ok) continue;
      if (r.a0 != cached) build(r.a0);
      if (r.n0 < nun) build(r.a1);
    }
  } else if (warp == 0) {
    // ---- issuer: the MMAs (whole warp converged, operands warp-uniform, elect.sync) -----------------------------------------------
    const uint32_t tm_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t sm_u = __shfl_sync(0xffffffffu, umma::smem_u32(smem), 0);
    const uint32_t idesc = umma::idesc_tf32(kVtcF, kVtcNP);
    float cached = 0.f;
    bool have = false;
    int k = 0;  // matrices taken over so far: the current one is build k - 1 in buffer (k - 1) & 1
    auto commit_to = [&](uint64_t* bar) {
      if (umma::elect_one())
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         sm_u + (uint32_t)VtfSmem::bars + 8u * (uint32_t)(bar - bars))
                     : "memory");
      __syncwarp();
    };
    auto next_matrix = [&](float a) {
      umma::mbar_wait(&bar_bfull[k & 1], (uint32_t)(k >> 1) & 1u);
      if (k >= 1) commit_to(&bar_bfree[(k - 1) & 1]);  // the buffer we leave is free when the MMAs issued so far have completed
      ++k;
      cached = a;
      have = true;
    };
    auto mma = [&](int b, uint32_t d_col) {
      if (umma::elect_one()) {
        const uint32_t b_lbo = kVtcNP * 16;
        const uint32_t ah = tm_u + kVtfTmA + 128 * b;
        const uint32_t bh = sm_u + VtfSmem::bmat + ((k - 1) & 1) * 2 * kVtcBBytes;
        umma::mma_3xtf32_ts<kVtcNP / 8>(tm_u + d_col, ah, ah + 64, umma::smem_desc(bh, b_lbo, 128), umma::smem_desc(bh + kVtcBBytes, b_lbo, 128),
                                        2 * b_lbo, idesc, false);
      }
      __syncwarp();
    };
    next_matrix(alpha_of(t_begin * kVtcF));
    for (int i = 0;; ++i) {
      const int b = i & 1;
      wait_slot(bar_afull, 2, i);  // (the converters arrive after the producer's description of segment i has become visible to them)
      VPROF_LAP(2);  // issuer: wait A
      const int4 ti = tinfo[i & 7];
      const float a0 = __int_as_float(tinfo[8 + (i & 7)].x);
      if (ti.x && a0 != cached) next_matrix(a0);
      VPROF_LAP(1);  // issuer: wait matrix
      umma::tc_fence_after_sync();
      if (ti.x) mma(b, kVtfTmD + 64 * b);
      commit_to(&bar_dfull[b]);
      VPROF_LAP(3);  // issuer: MMA issue
      if (ti.y) break;
    }
    VPROF_FLUSH(0, 4);
  } else if (warp == 3) {
    // ---- store thread: one bulk store per tile ------------------------------------------------------------------------------------
    if (lane == 0) {
      for (int j = 0;; ++j) {
        wait_slot(bar_ofull, 2, j);
        const int4 ti = tinfo[j & 7];
        if (ti.x) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(y + (t_begin * kVtcF + ti.w) * n),
                       "r"(umma::smem_u32(smem + VtfSmem::out + (j & 1) * kVtfStageBytes)), "r"((uint32_t)ti.z * row_bytes)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        umma::mbar_arrive(&bar_ofree[j & 1]);
        if (ti.y) break;
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else {
    // ---- converter / epilogue -----------------------------------------------------------------------------------------------------
    const int gw = warp - 4;
    const int q = warp & 3;                    // TMEM lane quarter of this warp
    const int qc = gw >> 2;                    // columns 16 qc .. 16 qc + 15 of the row
    const int row = 32 * q + lane;
    const uint32_t t_row = tmem + ((uint32_t)(32 * q) << 16);
    int total = 0x7fffffff;  // number of segments, known when the one flagged `last` arrives
    for (int i = 0; i <= total; ++i) {
      if (i < total) {
        const int s = i % kVtfNST, b = i & 1;
        wait_slot(bar_full, kVtfNST, i);
        VPROF_LAP(8);  // group: wait raw tile
        const int4 ti = tinfo[i & 7];
        if (ti.y) total = i + 1;
        float hi[16], lo[16];
        if (ti.x) {
          const float4* src = reinterpret_cast<const float4*>(smem + VtfSmem::in + s * kVtfStageBytes + (size_t)row * row_bytes) + 4 * qc;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (4 * qc + c < nq && row < ti.z) {
              w = src[c];
              if (has_norm) {
                const float4 sd = reinterpret_cast<const float4*>(vstd)[4 * qc + c], mu = reinterpret_cast<const float4*>(vmean)[4 * qc + c];
                w.x = fmaf(w.x, sd.x, mu.x);
                w.y = fmaf(w.y, sd.y, mu.y);
                w.z = fmaf(w.z, sd.z, mu.z);
                w.w = fmaf(w.w, sd.w, mu.w);
              }
            }
            umma::split_tf32(w.x, hi[4 * c], lo[4 * c]);
            umma::split_tf32(w.y, hi[4 * c + 1], lo[4 * c + 1]);
            umma::split_tf32(w.z, hi[4 * c + 2], lo[4 * c + 2]);
            umma::split_tf32(w.w, hi[4 * c + 3], lo[4 * c + 3]);
          }
        }
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(&bar_empty[s]);  // the raw stage may be refilled
        if (ti.x) {
          umma::tmem_st16(t_row + kVtfTmA + 128 * b + 16 * qc, hi);
          umma::tmem_st16(t_row + kVtfTmA + 128 * b + 64 + 16 * qc, lo);
          umma::tmem_st_wait();
        }
        umma::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(&bar_afull[b]);
        VPROF_LAP(9);  // group: read + convert + tensor-memory store
      }
      if (i >= 1) {
        const int j = i - 1, b = j & 1;
        const int4 tj = tinfo[j & 7];
        wait_slot(bar_dfull, 2, j);
        VPROF_LAP(10);  // group: wait MMA
        umma::tc_fence_after_sync();
        if (j >= 2) wait_slot(bar_ofree, 2, j - 2);
        VPROF_LAP(11);  // group: wait output stage
        if (tj.x) {
          float* orow = reinterpret_cast<float*>(smem + VtfSmem::out + b * kVtfStageBytes) + (size_t)row * n;
          float d[16];
          umma::tmem_ld16(t_row + kVtfTmD + 64 * b + 16 * qc, d);
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            const int c = 16 * qc + e;
            if (c < n && row < tj.z) {
              float4 o = make_float4(d[e], d[e + 1], d[e + 2], d[e + 3]);
              if (has_norm) {
                const float4 mu = *reinterpret_cast<const float4*>(vmean + c), rs = *reinterpret_cast<const float4*>(vrstd + c);
                o = make_float4((o.x - mu.x) * rs.x, (o.y - mu.y) * rs.y, (o.z - mu.z) * rs.z, (o.w - mu.w) * rs.w);
              }
              *reinterpret_cast<float4*>(orow + c) = o;
            }
          }
          umma::fence_proxy_async();
        }
        umma::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(&bar_ofull[b]);
        VPROF_LAP(12);  // group: epilogue
      }
    }
    VPROF_FLUSH(8, 13);
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

// ---- backward -------------------------------------------------------------------------------------------------------------------
// gx = (gy / std) . Bb^T . std,  Bb[r][j] = S1_r A[j][r] S2_j   (the transposed forward matrix)
// galpha(unit) = < gy / std , X' . Bt^T >,  Bt[j][r] = S2_j dA/dalpha[j][r] S1_r   (tangent of the same recursion)
// Two GEMMs per tile into two accumulators; tiles whose units do not share ONE alpha go to the recursion kernel.
//
// Same pipeline as the forward kernel, at OPERAND granularity: the stage ring, the converters and the tensor pipe work on the items
// G(0), X(0), G(1), X(1), ...  Tensor memory: A_G | A_X (hi / lo, 128 columns each, single buffered: A_G is free again when the
// first GEMM of the tile has completed, A_X when the second has), D1[2], D2[2] (64 columns each, by tile parity).  The sixteen
// converter warps run  G(i) -> X(i) -> epilogue(i - 1);  the epilogue writes its quarter rows of gx straight to global memory
// (64 contiguous bytes per thread) and leaves the partial dot products of d alpha in shared memory for the reduce warp.

// Row-wise construction of both matrices (see vtc_build_matrix): with E = A(alpha) and T = dA / dalpha,
//     E[j][r] = a E[j][r-1] + gE_r,   T[j][r] = a T[j][r-1] + gT_r          (first-order recurrences over r, resolved by warp scans)
//     row 0:   gE = [r == 0],                              gT = E[0][r-1]
//     row 1:   gE = (1 - a^2) E[0][r-1],                   gT = -2 a E[0][r-1] + (1 - a^2) T[0][r-1] + E[1][r-1]
//     row j:   gE = E[j-1][r-1] - a E[j-1][r],             gT = T[j-1][r-1] - a T[j-1][r] + E[j][r-1] - E[j-1][r]
// and column 0 of every row below the first is zero.
__device__ __forceinline__ void vtc_build_matrices_bwd(float* bb_hi, float* bb_lo, float* bt_hi, float* bt_lo, float a, int n) {
  const int lane = threadIdx.x & 31;
  const int r0 = 2 * lane;
  float pw[5];  // a^2, a^4, a^8, a^16, a^32
  pw[0] = a * a;
#pragma unroll
  for (int i = 1; i < 5; ++i) pw[i] = pw[i - 1] * pw[i - 1];
  const float bcoef = 1.f - a * a;
  // v[r] = a v[r-1] + g[r] for this lane's two columns (v[-1] = 0): affine maps of the incoming carry, Kogge-Stone over the lanes
  auto scan_row = [&](float g0, float g1, float& v0, float& v1, float& left) {  // left = v[r0 - 1] (0 for column -1)
    float x = fmaf(a, g0, g1);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const float t = __shfl_up_sync(0xffffffffu, x, 1 << i);
      if (lane >= (1 << i)) x = fmaf(pw[i], t, x);
    }
    float carry = __shfl_up_sync(0xffffffffu, x, 1);
    if (lane == 0) carry = 0.f;
    v0 = fmaf(a, carry, g0);
    v1 = x;
    left = carry;
  };
  float e0 = 0.f, e1 = 0.f, t0 = 0.f, t1 = 0.f, el = 0.f, tl = 0.f;  // previous row: this lane's columns and the one to their left
  const bool in = r0 < n;
  const uint32_t cbt = (uint32_t)(r0 >> 2) * (kVtcNP * 4) + (r0 & 3);  // Bt[n = j][k = r]: float offset of column r0 inside row j
  for (int j = 0; j < n; ++j) {
    float ne0, ne1, nt0, nt1, nel, ntl;
    if (j == 0) {
      scan_row(lane == 0 ? 1.f : 0.f, 0.f, ne0, ne1, nel);
      scan_row(nel, ne0, nt0, nt1, ntl);
    } else {
      const float cn = j == 1 ? bcoef : 1.f, cu = j == 1 ? 0.f : a;
      scan_row(lane == 0 ? 0.f : fmaf(cn, el, -cu * e0), fmaf(cn, e0, -cu * e1), ne0, ne1, nel);
      float gt0, gt1;
      if (j == 1) {
        gt0 = fmaf(-2.f * a, el, fmaf(bcoef, tl, nel));
        gt1 = fmaf(-2.f * a, e0, fmaf(bcoef, t0, ne0));
      } else {
        gt0 = (tl - a * t0) + (nel - e0);
        gt1 = (t0 - a * t1) + (ne0 - e1);
      }
      scan_row(lane == 0 ? 0.f : gt0, gt1, nt0, nt1, ntl);
    }
    e0 = ne0; e1 = ne1; t0 = nt0; t1 = nt1; el = nel; tl = ntl;
    if (in) {  // split and store (off the dependent chain of the recurrences)
      const float s2 = j == 0 ? 2.f : 1.f, s1 = lane == 0 ? 0.5f : 1.f;
      // Bb[n = r][k = j]
      const uint32_t kb = (uint32_t)(j >> 2) * (kVtcNP * 4) + (j & 3);
      const uint32_t ob0 = umma::tile_off(kVtcNP, r0, 0) / 4 + kb, ob1 = umma::tile_off(kVtcNP, r0 + 1, 0) / 4 + kb;
      float hi, lo;
      umma::split_tf32(e0 * s2 * s1, hi, lo);
      bb_hi[ob0] = hi;
      bb_lo[ob0] = lo;
      umma::split_tf32(e1 * s2, hi, lo);
      bb_hi[ob1] = hi;
      bb_lo[ob1] = lo;
      float2 th, tlo;
      umma::split_tf32(t0 * s2 * s1, th.x, tlo.x);
      umma::split_tf32(t1 * s2, th.y, tlo.y);
      const uint32_t ot = umma::tile_off(kVtcNP, j, 0) / 4 + cbt;
      *reinterpret_cast<float2*>(bt_hi + ot) = th;
      *reinterpret_cast<float2*>(bt_lo + ot) = tlo;
    }
  }
  __syncwarp();  // (entries outside n x n stay zero: the buffers are cleared once)
}

constexpr int kVtbThreads = 640;          // warp 0 issuer, 1 producer, 2 builder, 3 reduce, 4-19 converter / epilogue
struct VtbSmem {                          // offsets that do not depend on the stage size
  static constexpr uint32_t bmat = 0;                                 // [2][bb_hi | bb_lo | bt_hi | bt_lo] 2 x 64 KB
  static constexpr uint32_t vec = 8 * kVtcBBytes;                     // mean[64], std[64], 1/std[64]
  static constexpr uint32_t gpart = vec + 3 * kVtcNP * 4;             // [2][4][128] partial dot products
  static constexpr uint32_t bars = gpart + 2 * 4 * kVtcF * 4;
  static constexpr uint32_t nbars = 2 * 4 + 14;                       // full[4], empty[4], then see the kernel
  static constexpr uint32_t misc = bars + nbars * 8;                  // tmem slot, tile info ring
  static constexpr uint32_t stages = (misc + 16 + 8 * 32 + 127) & ~127u;  // [nst][stage_bytes]
  static_assert(misc % 16 == 0 && stages % 128 == 0, "alignment of the tile info ring / the bulk copy destinations");
};
constexpr int kVtbTmAG = 0, kVtbTmAX = 128, kVtbTmD1 = 256, kVtbTmD2 = 384;

__global__ void __launch_bounds__(kVtbThreads, 1)
allpass_tc_backward_kernel(const float* __restrict__ gy, const float* __restrict__ x, const float* __restrict__ alpha, int64_t units, int n,
                           int blocks, const float* __restrict__ mean, const float* __restrict__ std_dev, float* __restrict__ gx,
                           float* __restrict__ galpha_unit, uint8_t* __restrict__ tile_mixed, int64_t num_tiles, int nst,
                           uint32_t stage_bytes) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* vmean = reinterpret_cast<float*>(smem + VtbSmem::vec);
  float* vstd = vmean + kVtcNP;
  float* vrstd = vstd + kVtcNP;
  float* gpart = reinterpret_cast<float*>(smem + VtbSmem::gpart);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + VtbSmem::bars);
  uint64_t* bar_full = bars;              // [nst <= 4] operand tile (+ the tile's classification with the G item) has landed
  uint64_t* bar_empty = bars + 4;         // [nst <= 4] every converter warp has read the operand tile
  uint64_t* bar_ag = bars + 8;            // A_G written (16 warps), once per tile
  uint64_t* bar_ax = bars + 9;            // A_X written
  uint64_t* bar_d1 = bars + 10;           // [2] first GEMM of the tile complete: D1[b] ready, A_G free
  uint64_t* bar_d2 = bars + 12;           // [2] second GEMM complete: D2[b] ready, A_X free
  uint64_t* bar_ofull = bars + 14;        // [2] partial dot products of the tile written (16 warps)
  uint64_t* bar_ofree = bars + 16;        // [2] the reduce warp has consumed them
  uint64_t* bar_bfull = bars + 18;        // [2] matrix set [k & 1] holds build k
  uint64_t* bar_bfree = bars + 20;        // [2] every MMA that read matrix set [k & 1] has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + VtbSmem::misc);
  int4* tinfo = reinterpret_cast<int4*>(smem + VtbSmem::misc + 16);  // [8] (ok, -, nun, -) of tile i at i & 7, then [8] (alpha)
  uint8_t* stages = smem + VtbSmem::stages;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t per = (num_tiles + gridDim.x - 1) / gridDim.x;
  const int64_t t_begin = (int64_t)blockIdx.x * per;
  const int64_t t_end = min(num_tiles, t_begin + per);
  if (t_begin >= t_end) return;
  const int cnt = (int)(t_end - t_begin);

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      umma::mbar_init(&bar_full[i], 2);
      umma::mbar_init(&bar_empty[i], kVtfGW);
    }
    umma::mbar_init(bar_ag, kVtfGW);
    umma::mbar_init(bar_ax, kVtfGW);
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(&bar_d1[i], 1);
      umma::mbar_init(&bar_d2[i], 1);
      umma::mbar_init(&bar_ofull[i], kVtfGW);
      umma::mbar_init(&bar_ofree[i], 1);
      umma::mbar_init(&bar_bfull[i], 1);
      umma::mbar_init(&bar_bfree[i], 1);
    }
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
  const bool norm_ok = blocks == 1 || (mean == nullptr && std_dev == nullptr);
  const bool has_norm = mean != nullptr || std_dev != nullptr;
  if (tid < kVtcNP) {
    const bool in = tid < n && blocks == 1;
    vmean[tid] = (in && mean) ? mean[tid] : 0.f;
    const float sd = (in && std_dev) ? std_dev[tid] : 1.f;
    vstd[tid] = sd;
    vrstd[tid] = 1.f / sd;
  }
  for (int i = tid; i < (int)(8 * kVtcBBytes / 4); i += kVtbThreads) reinterpret_cast<float*>(smem + VtbSmem::bmat)[i] = 0.f;
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const int nq = n / 4;
  const uint32_t row_bytes = (uint32_t)n * 4u;
  auto wait_idx = [&](uint64_t* bar, int index) { umma::mbar_wait(bar, (uint32_t)index & 1u); };
  VPROF_DECL;
  auto wait_item = [&](uint64_t* arr, int k) { umma::mbar_wait(&arr[k % nst], (uint32_t)(k / nst) & 1u); };  // stage barriers, by item
  auto alpha_of = [&](int64_t u) { return blocks == 1 ? alpha[u] : alpha[u / blocks]; };
  auto tile_units = [&](int i) { return (int)min((int64_t)kVtcF, units - (t_begin + i) * kVtcF); };
  float run_ar0[4], run_ar1[4], run_a00 = 0.f, run_a01 = 0.f, run_a10 = 0.f, run_a11 = 0.f;
  auto load_into = [&](int i, float (&ar)[4], float& a0, float& a1) {
    const int64_t u0 = (t_begin + i) * kVtcF;
    const int nun = tile_units(i);
    a0 = alpha_of(u0);
    a1 = alpha_of(u0 + nun - 1);
#pragma unroll
    for (int h = 0; h < 4; ++h) ar[h] = lane + 32 * h < nun ? alpha_of(u0 + lane + 32 * h) : 0.f;
  };
  auto load_runs = [&](int i) {
    if (i >= cnt) return;
    if (i & 1) load_into(i, run_ar1, run_a01, run_a11);
    else load_into(i, run_ar0, run_a00, run_a10);
  };
  auto runs_of = [&](int i) {  // classification of tile i (refills the slot with tile i + 2)
    VtfRuns r;
    if (i & 1) r = vtf_runs(run_ar1, run_a01, run_a11, tile_units(i), lane);
    else r = vtf_runs(run_ar0, run_a00, run_a10, tile_units(i), lane);
    load_runs(i + 2);
    r.ok = r.ok && norm_ok;
    return r;
  };

  if (warp == 1) {
    // ---- producer: items 2 i (gy tile) and 2 i + 1 (x tile) ---------------------------------------------------------------------
    // Segments with ONE alpha, as in the forward kernel; a segment is two items: its gy rows (with the description) and its x rows.
    load_runs(0);
    load_runs(1);
    int sidx = 0;
    for (int i = 0; i < cnt; ++i) {
      const int nun = tile_units(i);
      const VtfRuns r = runs_of(i);
      const int nseg = (r.ok && r.n0 < nun) ? 2 : 1;
      for (int sg = 0; sg < nseg; ++sg, ++sidx) {
        if (lane == 0) {
          const int off = sg ? r.n0 : 0, cu = nseg == 2 ? (sg ? nun - r.n0 : r.n0) : nun;
          const int ustart = i * kVtcF + off;
          const uint32_t bytes = (uint32_t)cu * row_bytes;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int k = 2 * sidx + h, s = k % nst;
            if (k >= nst) wait_item(bar_empty, k - nst);
            umma::mbar_expect_tx(&bar_full[s], bytes);
            umma::bulk_g2s(stages + (size_t)s * stage_bytes, (h ? x : gy) + (t_begin * kVtcF + ustart) * n, bytes, &bar_full[s]);
            if (h == 0) {
              tinfo[sidx & 7] = make_int4(r.ok ? 1 : 0, (i == cnt - 1 && sg == nseg - 1) ? 1 : 0, cu, ustart);
              tinfo[8 + (sidx & 7)] = make_int4(__float_as_int(sg ? r.a1 : r.a0), 0, 0, 0);
            }
            umma::mbar_arrive(&bar_full[s]);
          }
        }
      }
      if (lane == 0) tile_mixed[t_begin + i] = r.ok ? 0 : 1;
      __syncwarp();
    }
  } else if (warp == 2) {
    // ---- builder ----------------------------------------------------------------------------------------------------------------
    float cached = 0.f;
    int k = 0;
    auto build = [&](float a) {
      if (k >= 2) umma::mbar_wait(&bar_bfree[k & 1], (uint32_t)((k >> 1) - 1) & 1u);
      float* bh = reinterpret_cast<float*>(smem + VtbSmem::bmat + (k & 1) * 4 * kVtcBBytes);
      vtc_build_matrices_bwd(bh, bh + kVtcBBytes / 4, bh + 2 * (kVtcBBytes / 4), bh + 3 * (kVtcBBytes / 4), a, n);
      umma::fence_proxy_async();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&bar_bfull[k & 1]);
      ++k;
      cached = a;
    };
    build(alpha_of(t_begin * kVtcF));
    load_runs(0);
    load_runs(1);
    for (int i = 0; i < cnt; ++i) {
      const VtfRuns r = runs_of(i);
      if (!r.ok) continue;
      if (r.a0 != cached) build(r.a0);
      if (r.n0 < tile_units(i)) build(r.a1);
    }
  } else if (warp == 0) {
    // ---- issuer ---------------------------------------------------------------------------------------------------------------------
    const uint32_t tm_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t sm_u = __shfl_sync(0xffffffffu, umma::smem_u32(smem), 0);
    const uint32_t idesc = umma::idesc_tf32(kVtcF, kVtcNP);
    float cached = 0.f;
    int k = 0;
    auto commit_to = [&](uint64_t* bar) {
      if (umma::elect_one())
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         sm_u + (uint32_t)VtbSmem::bars + 8u * (uint32_t)(bar - bars))
                     : "memory");
      __syncwarp();
    };
    auto next_matrix = [&](float a) {
      umma::mbar_wait(&bar_bfull[k & 1], (uint32_t)(k >> 1) & 1u);
      if (k >= 1) commit_to(&bar_bfree[(k - 1) & 1]);
      ++k;
      cached = a;
    };
    auto mma = [&](uint32_t a_col, uint32_t mat, uint32_t d_col) {  // mat: 0 = Bb, 1 = Bt of the current set
      if (umma::elect_one()) {
        const uint32_t b_lbo = kVtcNP * 16;
        const uint32_t bh = sm_u + VtbSmem::bmat + ((k - 1) & 1) * 4 * kVtcBBytes + mat * 2 * kVtcBBytes;
        umma::mma_3xtf32_ts<kVtcNP / 8>(tm_u + d_col, tm_u + a_col, tm_u + a_col + 64, umma::smem_desc(bh, b_lbo, 128),
                                        umma::smem_desc(bh + kVtcBBytes, b_lbo, 128), 2 * b_lbo, idesc, false);
      }
      __syncwarp();
    };
    next_matrix(alpha_of(t_begin * kVtcF));
    for (int i = 0;; ++i) {
      const int b = i & 1;
      wait_idx(bar_ag, i);
      VPROF_LAP(0);  // issuer: wait A_G
      const int4 ti = tinfo[i & 7];
      const float a0 = __int_as_float(tinfo[8 + (i & 7)].x);
      if (ti.x && a0 != cached) next_matrix(a0);
      VPROF_LAP(1);  // issuer: wait matrix
      umma::tc_fence_after_sync();
      if (ti.x) mma(kVtbTmAG, 0, kVtbTmD1 + 64 * b);
      commit_to(&bar_d1[b]);
      VPROF_LAP(2);  // issuer: GEMM 1 issue
      wait_idx(bar_ax, i);
      VPROF_LAP(3);  // issuer: wait A_X
      umma::tc_fence_after_sync();
      if (ti.x) mma(kVtbTmAX, 1, kVtbTmD2 + 64 * b);
      commit_to(&bar_d2[b]);
      VPROF_LAP(4);  // issuer: GEMM 2 issue
      if (ti.y) break;
    }
    VPROF_FLUSH(0, 5);
  } else if (warp == 3) {
    // ---- store / reduce warp: gx of a segment leaves with one bulk store from the stage its rows were staged in (the stage that held
    // the x rows of the NEXT segment), which is then handed back to the producer; d alpha of every unit = sum of the four quarter-row
    // partial dot products
    for (int j = 0;; ++j) {
      const int b = j & 1;
      wait_idx(&bar_ofull[b], j >> 1);
      const int4 tj = tinfo[j & 7];
      const int s = (2 * (j + 1) + 1) % nst;  // staging = stage of item X(j + 1) (for the last segment: a stage nobody uses any more)
      if (tj.x) {
        if (lane == 0) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gx + (t_begin * kVtcF + tj.w) * n),
                       "r"(umma::smem_u32(stages + (size_t)s * stage_bytes)), "r"((uint32_t)tj.z * row_bytes)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        const float* gp = gpart + b * 4 * kVtcF;
        const int64_t u0 = t_begin * kVtcF + tj.w;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int row = lane + 32 * h;
          if (row < tj.z) galpha_unit[u0 + row] = (gp[row] + gp[kVtcF + row]) + (gp[2 * kVtcF + row] + gp[3 * kVtcF + row]);
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      __syncwarp();
      if (lane == 0) {
        umma::mbar_arrive(&bar_ofree[b]);
        if (!tj.y) umma::mbar_arrive_n(&bar_empty[s], kVtfGW);  // on behalf of the sixteen converter warps that read the x rows
      }
      if (tj.y) break;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else {
    // ---- converter / epilogue -----------------------------------------------------------------------------------------------------
    const int gw = warp - 4;
    const int q = warp & 3;                    // TMEM lane quarter of this warp
    const int qc = gw >> 2;                    // columns 16 qc .. 16 qc + 15 of the row
    const int row = 32 * q + lane;
    const uint32_t t_row = tmem + ((uint32_t)(32 * q) << 16);
    float gprev[16];                           // this thread's part of gy / std of the previous tile (for the dot product)
#pragma unroll
    for (int e = 0; e < 16; ++e) gprev[e] = 0.f;
    int total = 0x7fffffff;  // number of segments, known when the one flagged `last` arrives
    for (int i = 0; i <= total; ++i) {
      float gcur[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) gcur[e] = 0.f;
      if (i < total) {
        const int kg = 2 * i, kx = 2 * i + 1;
        // ---- G(i) -> A_G ----
        wait_item(bar_full, kg);
        VPROF_LAP(8);  // group: wait gy rows
        const int4 ti = tinfo[i & 7];
        if (ti.y) total = i + 1;
        if (i >= 1) wait_idx(&bar_d1[(i - 1) & 1], (i - 1) >> 1);  // the first GEMM of the previous tile has read A_G
        VPROF_LAP(9);  // group: wait A_G free
        float hi[16], lo[16];
        if (ti.x) {
          const float4* src = reinterpret_cast<const float4*>(stages + (size_t)(kg % nst) * stage_bytes + (size_t)row * row_bytes) + 4 * qc;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (4 * qc + c < nq && row < ti.z) {
              w = src[c];
              if (has_norm) {
                const float4 rs = reinterpret_cast<const float4*>(vrstd)[4 * qc + c];
                w.x *= rs.x; w.y *= rs.y; w.z *= rs.z; w.w *= rs.w;
              }
            }
            gcur[4 * c] = w.x; gcur[4 * c + 1] = w.y; gcur[4 * c + 2] = w.z; gcur[4 * c + 3] = w.w;
            umma::split_tf32(w.x, hi[4 * c], lo[4 * c]);
            umma::split_tf32(w.y, hi[4 * c + 1], lo[4 * c + 1]);
            umma::split_tf32(w.z, hi[4 * c + 2], lo[4 * c + 2]);
            umma::split_tf32(w.w, hi[4 * c + 3], lo[4 * c + 3]);
          }
        }
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(&bar_empty[kg % nst]);
        if (ti.x) {
          umma::tmem_st16(t_row + kVtbTmAG + 16 * qc, hi);
          umma::tmem_st16(t_row + kVtbTmAG + 64 + 16 * qc, lo);
          umma::tmem_st_wait();
        }
        umma::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(bar_ag);
        VPROF_LAP(10);  // group: convert G
        // ---- X(i) -> A_X ----
        wait_item(bar_full, kx);
        if (i >= 1) wait_idx(&bar_d2[(i - 1) & 1], (i - 1) >> 1);  // the second GEMM of the previous tile has read A_X
        VPROF_LAP(11);  // group: wait x rows + A_X free
        if (ti.x) {
          const float4* src = reinterpret_cast<const float4*>(stages + (size_t)(kx % nst) * stage_bytes + (size_t)row * row_bytes) + 4 * qc;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (4 * qc + c < nq && row < ti.z) {
              w = src[c];
              if (has_norm) {
                const float4 sd = reinterpret_cast<const float4*>(vstd)[4 * qc + c], mu = reinterpret_cast<const float4*>(vmean)[4 * qc + c];
                w.x = fmaf(w.x, sd.x, mu.x);
                w.y = fmaf(w.y, sd.y, mu.y);
                w.z = fmaf(w.z, sd.z, mu.z);
                w.w = fmaf(w.w, sd.w, mu.w);
              }
            }
            umma::split_tf32(w.x, hi[4 * c], lo[4 * c]);
            umma::split_tf32(w.y, hi[4 * c + 1], lo[4 * c + 1]);
            umma::split_tf32(w.z, hi[4 * c + 2], lo[4 * c + 2]);
            umma::split_tf32(w.w, hi[4 * c + 3], lo[4 * c + 3]);
          }
        }
        // (the stage of the x rows is NOT released here: it becomes the staging buffer of the previous segment's gx rows and is
        // handed back to the producer by the store warp once the bulk store has read it; segment 0 has no predecessor)
        if (i == 0) {
          __syncwarp();
          if (lane == 0) umma::mbar_arrive(&bar_empty[kx % nst]);
        }
        if (ti.x) {
          umma::tmem_st16(t_row + kVtbTmAX + 16 * qc, hi);
          umma::tmem_st16(t_row + kVtbTmAX + 64 + 16 * qc, lo);
          umma::tmem_st_wait();
        }
        umma::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(bar_ax);
        VPROF_LAP(12);  // group: convert X
      }
      if (i >= 1) {
        // ---- epilogue of tile i - 1 ----
        const int j = i - 1, b = j & 1;
        const int4 tj = tinfo[j & 7];
        if (i == total) {  // (before that both GEMMs of segment i - 1 have been waited for above)
          wait_idx(&bar_d1[b], j >> 1);
          wait_idx(&bar_d2[b], j >> 1);
        }
        umma::tc_fence_after_sync();
        if (j >= 2) wait_idx(&bar_ofree[b], (j >> 1) - 1);
        if (tj.x) {
          float d[16];
          umma::tmem_ld16(t_row + kVtbTmD1 + 64 * b + 16 * qc, d);  // gradient w.r.t. the de-normalised input
          float* orow = reinterpret_cast<float*>(stages + (size_t)((2 * i + 1) % nst) * stage_bytes) + (size_t)row * n;
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            const int c = 16 * qc + e;
            if (c < n && row < tj.z) {
              float4 o = make_float4(d[e], d[e + 1], d[e + 2], d[e + 3]);
              if (has_norm) {
                const float4 sd = *reinterpret_cast<const float4*>(vstd + c);
                o.x *= sd.x; o.y *= sd.y; o.z *= sd.z; o.w *= sd.w;
              }
              *reinterpret_cast<float4*>(orow + c) = o;
            }
          }
          umma::tmem_ld16(t_row + kVtbTmD2 + 64 * b + 16 * qc, d);  // d y / d alpha
          float ga = 0.f;
#pragma unroll
          for (int e = 0; e < 16; ++e) ga = fmaf(gprev[e], d[e], ga);
          gpart[(b * 4 + qc) * kVtcF + row] = ga;
          umma::fence_proxy_async();  // the staged gx rows are read by a bulk store
        }
        umma::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(&bar_ofull[b]);
        VPROF_LAP(13);  // group: epilogue
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) gprev[e] = gcur[e];
    }
    VPROF_FLUSH(8, 14);
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

}  // namespace b2w

extern "C" int b2w_vtf_prof_read(long long* out16) { return (int)cudaMemcpyFromSymbol(out16, b2w::g_vtf_prof, sizeof(long long) * 16); }

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
  const int grid = (int)(num_tiles < 148 ? num_tiles : 148);
  cudaStream_t st = (cudaStream_t)stream;
  cudaFuncSetAttribute(allpass_tc_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VtfSmem::total);
  allpass_tc_forward_kernel<<<grid, kVtfThreads, VtfSmem::total, st>>>(x, alpha, units, n, blocks, mean, std_dev, y, tile_flags, num_tiles);
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
  const uint32_t stage_bytes = ((uint32_t)kVtcF * (uint32_t)n * 4u + 127u) & ~127u;
  const int nst = (VtbSmem::stages + 3 * stage_bytes <= 227 * 1024) ? 3 : 2;   // three operand stages when they fit (n <= 60)
  const uint32_t smem = VtbSmem::stages + (uint32_t)nst * stage_bytes;
  cudaFuncSetAttribute(allpass_tc_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  allpass_tc_backward_kernel<<<grid, kVtbThreads, smem, st>>>(grad_y, x, alpha, units, n, blocks, mean, std_dev, grad_x,
                                                              blocks == 1 ? grad_alpha : unit_workspace,
                                                              tile_flags, num_tiles, nst, stage_bytes);
  int rc = check_launch("allpass_tc_backward_kernel");
  if (rc) return rc;
  return b2w_allpass_backward_masked(grad_y, x, alpha, rows, n, blocks, mean, std_dev, grad_x, grad_alpha, unit_workspace, tile_flags, stream);
}
