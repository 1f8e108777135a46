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


// One warp: writes B = S2 A(alpha) S1 (hi / lo TF32 tiles, K-major) for the n x n freqt matrix of `a`.
// Row by row: with (cn, cu) = (1 - a^2, 0) for row 1 and (1, 1) below it, the freqt recursion
//     A[j][r] = cn A[j-1][r-1] + a (A[j][r-1] - cu A[j-1][r]),     A[j][0] = (j == 0),   A[0][r] = a^r
// is, for a fixed row, a first-order linear recurrence over r with the CONSTANT coefficient a and a forcing term that only needs the
// previous row: A[j][r] = a A[j][r-1] + g_r.  Lane l owns columns 2 l and 2 l + 1; a five-step warp scan over affine maps (the
// multipliers are powers of a^2) resolves the recurrence, so a row costs ~40 instructions instead of a 119-step wavefront.
// The unsplit fp32 values go to their final position in the hi tile; a second, fully parallel sweep splits them into hi / lo.
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
  for (int j = 0; j < n; ++j) {
    if (j > 0) {
      const float cn = j == 1 ? 1.f - a * a : 1.f, cu = j == 1 ? 0.f : a;
      float left = __shfl_up_sync(0xffffffffu, p1, 1);       // A[j-1][r0 - 1]
      if (lane == 0) left = 0.f;
      const float g0 = lane == 0 ? 0.f : fmaf(cn, left, -cu * p0);   // column 0 of rows >= 1 is zero, and so is its forcing term
      const float g1 = fmaf(cn, p0, -cu * p1);
      // this lane's pair as an affine map of the incoming carry c = A[j][r0 - 1]:  A[j][r0] = a c + g0,  A[j][r1] = a^2 c + x
      float x = fmaf(a, g0, g1);
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const float t = __shfl_up_sync(0xffffffffu, x, 1 << i);
        if (lane >= (1 << i)) x = fmaf(pw[i], t, x);
      }
      float carry = __shfl_up_sync(0xffffffffu, x, 1);        // A[j][r0 - 1]
      if (lane == 0) carry = 0.f;
      p0 = lane == 0 ? 0.f : fmaf(a, carry, g0);
      p1 = x;
    }
    if (in) {
      const float s2 = j == 0 ? 2.f : 1.f;                    // S2; S1 halves column 0
      const uint32_t off = umma::tile_off(kVtcNP, j, 0) / 4 + cb;
      *reinterpret_cast<float2*>(b_hi + off) = make_float2(p0 * s2 * (lane == 0 ? 0.5f : 1.f), p1 * s2);
    }
  }
  __syncwarp();
  for (int e = lane; e < kVtcNP * kVtcNP / 4; e += 32) {  // (entries outside n x n stay zero: the buffers are cleared once)
    const float4 v = reinterpret_cast<const float4*>(b_hi)[e];
    float4 hi, lo;
    umma::split_tf32(v.x, hi.x, lo.x);
    umma::split_tf32(v.y, hi.y, lo.y);
    umma::split_tf32(v.z, hi.z, lo.z);
    umma::split_tf32(v.w, hi.w, lo.w);
    reinterpret_cast<float4*>(b_hi)[e] = hi;
    reinterpret_cast<float4*>(b_lo)[e] = lo;
  }
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
constexpr int kVtfTmD2 = 384;  // second accumulator of a two-run tile at 384 + 64 b

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
    for (int i = 0; i < kVtfNST && i < cnt; ++i) {  // the first tiles are on their way before anything else happens
      if (lane == 0) {
        umma::mbar_expect_tx(&bar_full[i], (uint32_t)tile_units(i) * row_bytes);
        umma::bulk_g2s(smem + VtfSmem::in + i * kVtfStageBytes, x + (t_begin + i) * kVtcF * n, (uint32_t)tile_units(i) * row_bytes, &bar_full[i]);
      }
    }
    load_runs(0);
    load_runs(1);
    for (int i = 0; i < cnt; ++i) {
      const int s = i % kVtfNST;
      const int nun = tile_units(i);
      if (lane == 0 && i >= kVtfNST) {
        wait_slot(bar_empty, kVtfNST, i - kVtfNST);
        umma::mbar_expect_tx(&bar_full[s], (uint32_t)nun * row_bytes);
        umma::bulk_g2s(smem + VtfSmem::in + s * kVtfStageBytes, x + (t_begin + i) * kVtcF * n, (uint32_t)nun * row_bytes, &bar_full[s]);
      }
      __syncwarp();
      const VtfRuns r = runs_of(i);
      if (lane == 0) {
        tinfo[i & 7] = make_int4(r.ok ? 1 : 0, r.n0, nun, 0);  // (slot i & 7 was last read for tile i - 8: long retired)
        tinfo[8 + (i & 7)] = make_int4(__float_as_int(r.a0), __float_as_int(r.a1), 0, 0);
        tile_mixed[t_begin + i] = r.ok ? 0 : 1;
        umma::mbar_arrive(&bar_full[s]);
      }
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
    for (int i = 0; i < cnt; ++i) {
      const int b = i & 1;
      wait_slot(bar_afull, 2, i);  // (the converters arrive after the producer's classification of tile i has become visible to them)
      VPROF_LAP(2);  // issuer: wait A
      const int4 ti = tinfo[i & 7], ta = tinfo[8 + (i & 7)];
      const bool ok = ti.x != 0;
      const float a0 = __int_as_float(ta.x), a1 = __int_as_float(ta.y);
      if (ok && a0 != cached) next_matrix(a0);
      VPROF_LAP(1);  // issuer: wait matrix
      umma::tc_fence_after_sync();
      if (ok) mma(b, kVtfTmD + 64 * b);
      if (ok && ti.y < ti.z) {  // rows of the second run: same A operand, the next speaker's matrix, second accumulator
        next_matrix(a1);
        mma(b, kVtfTmD2 + 64 * b);
      }
      commit_to(&bar_dfull[b]);
      VPROF_LAP(3);  // issuer: MMA issue
    }
    VPROF_FLUSH(0, 4);
  } else if (warp == 3) {
    // ---- store thread: one bulk store per tile ------------------------------------------------------------------------------------
    if (lane == 0) {
      for (int j = 0; j < cnt; ++j) {
        wait_slot(bar_ofull, 2, j);
        const int4 ti = tinfo[j & 7];
        if (ti.x) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(y + (t_begin + j) * kVtcF * n),
                       "r"(umma::smem_u32(smem + VtfSmem::out + (j & 1) * kVtfStageBytes)), "r"((uint32_t)ti.z * row_bytes)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        umma::mbar_arrive(&bar_ofree[j & 1]);
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
    for (int i = 0; i <= cnt; ++i) {
      if (i < cnt) {
        const int s = i % kVtfNST, b = i & 1;
        wait_slot(bar_full, kVtfNST, i);
        VPROF_LAP(8);  // group: wait raw tile
        const int4 ti = tinfo[i & 7];
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
          if (tj.y < tj.z) {  // two runs: rows of the second run take the second accumulator (warp-wide load, per-row choice)
            float d2[16];
            umma::tmem_ld16(t_row + kVtfTmD2 + 64 * b + 16 * qc, d2);
#pragma unroll
            for (int e = 0; e < 16; ++e) d[e] = row < tj.y ? d[e] : d2[e];
          }
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
    if (warp == 0) {  // converged warp, warp-uniform operands, MMAs under elect.sync (see the forward kernel)
      umma::tc_fence_after_sync();
      const uint32_t a_lbo = kVtcF * 16, b_lbo = kVtcNP * 16;
      const uint32_t tm_u = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t sm_u = __shfl_sync(0xffffffffu, umma::smem_u32(smem), 0);
      if (umma::elect_one()) {
        umma::mma_3xtf32<kVtcNP / 8>(tm_u, umma::smem_desc(sm_u + VtcBwdSmem::g_hi, a_lbo, 128), umma::smem_desc(sm_u + VtcBwdSmem::g_lo, a_lbo, 128),
                                     umma::smem_desc(sm_u + VtcBwdSmem::bb_hi, b_lbo, 128), umma::smem_desc(sm_u + VtcBwdSmem::bb_lo, b_lbo, 128),
                                     2 * a_lbo, 2 * b_lbo, idesc, false);
        umma::mma_3xtf32<kVtcNP / 8>(tm_u + 64, umma::smem_desc(sm_u + VtcBwdSmem::x_hi, a_lbo, 128), umma::smem_desc(sm_u + VtcBwdSmem::x_lo, a_lbo, 128),
                                     umma::smem_desc(sm_u + VtcBwdSmem::bt_hi, b_lbo, 128), umma::smem_desc(sm_u + VtcBwdSmem::bt_lo, b_lbo, 128),
                                     2 * a_lbo, 2 * b_lbo, idesc, false);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sm_u + (uint32_t)VtcBwdSmem::misc) : "memory");
      }
      __syncwarp();
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
  cudaFuncSetAttribute(allpass_tc_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VtcBwdSmem::total);
  allpass_tc_backward_kernel<<<grid, kVtcThreads, VtcBwdSmem::total, st>>>(grad_y, x, alpha, units, n, blocks, mean, std_dev, grad_x,
                                                                           unit_workspace, tile_flags, num_tiles);
  int rc = check_launch("allpass_tc_backward_kernel");
  if (rc) return rc;
  return b2w_allpass_backward_masked(grad_y, x, alpha, rows, n, blocks, mean, std_dev, grad_x, grad_alpha, unit_workspace, tile_flags, stream);
}
