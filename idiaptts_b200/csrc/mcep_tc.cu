// Mel-cepstral analysis (SPTK mcep Newton loop) with the dense all-pass-warp contractions on the 5th-generation tensor
// cores: tcgen05.mma kind::tf32 with the 3xTF32 split (fp32-accurate), accumulators in tensor memory, constant matrices
// streamed by 1-D bulk async copies (mbarrier completion) from a pre-tiled global stream.
//
// Same mathematics and the same reference call site as mcep.cu (pysptk.mcep, AudioProcessing.py:146); this is the
// production path for order <= 63.  One CTA = 128 frames (the UMMA M dimension), 512 threads.
//
//   per Newton iteration, for each chunk of 32 frequency bins j0 .. j0+31 (17 chunks at K = 513):
//     GEMM1   D1[128 x 32]   = mc[128 x 64] . Cmat[64 x 32 chunk]          24 MMAs (8 K-steps x 3 split products)
//     epilogue (8 warps)      P = per * exp(-2 D1)  -> hi/lo TF32 in TENSOR MEMORY (tcgen05.st; GEMM2 takes its A operand from
//                             there, double buffered: no shared-memory round trip, no proxy fence)
//     GEMM2   D2[128 x 128] += P[128 x 32] . M2^T[32 chunk x 128]          12 MMAs
//   The GEMM phase is warp specialised: thread 32 only streams the matrices (the Cmat and M2^T halves of a stage are refilled
//   independently, each as soon as its own GEMM has completed), thread 0 only issues MMAs and never waits for one to complete
//   (GEMM1 of chunk c + 1 is in the tensor pipe while chunk c's epilogue runs: D1 is double buffered in tensor memory), warps
//   4-11 are the epilogue (one TMEM lane = one frame per thread, 16 bins each), everything is handed over through mbarriers
//   -- no CTA barrier inside the chunk loop.
//   then D2 row f = r~ of frame f: stopping rule on r~[0], and for the frames still iterating one warp each builds and
//   solves the (m+1) x (m+1) Toeplitz-plus-Hankel system (mcep_solve.cuh) and updates mc (fp32, shared memory).
//   Pass 0 (initial value) uses the same machinery: P = log(per), the stream holds M0^T instead of M2^T, GEMM1 is skipped.
#include "common.cuh"
#include "mcep_solve.cuh"
#include "umma.cuh"

namespace b2w {

constexpr int kTcF = 128;        // frames per CTA = UMMA M
constexpr int kTcThreads = 512;  // 16 warps: TMEM lane quarter q = warp & 3, column group g = warp >> 2
constexpr int kTcIssuer = 0;     // the thread that issues the MMAs
constexpr int kTcProducer = 32;  // the thread that streams the constant matrices (bulk async copies)
constexpr int kTcEpiWarp0 = 4;   // epilogue warps 4 .. 15: three sets of four (one warp per TMEM lane quarter)
constexpr int kTcEpiSets = 3;
constexpr int kTcSolveWarps = 16;
constexpr int kTcBK = 32;        // bins per chunk
constexpr int kTcMP = 64;        // padded cepstral dimension (K of GEMM1)
constexpr int kTcMS = 68;        // row stride of the fp32 mel-cepstra in shared memory (16-byte aligned rows, 8 rows span all banks)
constexpr int kTcNST = 3;        // stages of the constant-matrix ring
constexpr int kTcN2 = 128;       // padded r~ length (N of GEMM2)
constexpr int kTcKB = 20;        // padded block stride of the solve workspace (bank-conflict free float4 accesses)
constexpr uint32_t kB1Bytes = kTcBK * kTcMP * 4;   // one of hi / lo of the Cmat chunk   [N = 16 rows (bins)] x [K = 64]
constexpr uint32_t kB2Bytes = kTcN2 * kTcBK * 4;   // one of hi / lo of the M2^T chunk   [N = 128 rows]       x [K = 16]
constexpr uint32_t kStageBytes = 2 * kB1Bytes + 2 * kB2Bytes;  // 48 KB per chunk: [B1 hi | B1 lo | B2 hi | B2 lo]
constexpr int kTcBars = 4 * kTcNST + 4;
#ifndef B2W_TC_COPIES
#define B2W_TC_COPIES 4
#endif
constexpr int kTcCopies = B2W_TC_COPIES;  // replicas of the pre-tiled streams: CTA b reads replica b % kTcCopies (all CTAs stream the same 816 KB
                                          // per pass at about the same time; replicas spread that over more L2 lines / slices)
// tensor memory map (columns): D1[0] 0..31 | D2 32..159 | D1[1] 160..191 | A2[b] hi / lo at 192 + 64 b / 224 + 64 b | A1 hi 320..383 | A1 lo 384..447
constexpr int kTmemCols = 512;
constexpr int kTmA2 = 192;
constexpr int kTmA1 = 320;

// Phase timing of CTA 0 (build with -DB2W_MCEP_PROF via scripts/build_variant.py; read with b2w_mcep_prof_read): slots 0-7 are
// the issuer thread, 8-13 one epilogue thread, 14-15 the pass as seen by thread 32.
__device__ long long g_mcep_prof[16];
#ifdef B2W_MCEP_PROF
#define PROF_DECL long long prof_t = 0, prof_acc[16] = {0}; const bool prof_on = blockIdx.x == 0 && (tid == kTcIssuer || tid == 128)
#define PROF_START() do { if (prof_on) prof_t = clock64(); } while (0)
#define PROF_LAP(i) do { if (prof_on) { const long long n_ = clock64(); prof_acc[i] += n_ - prof_t; prof_t = n_; } } while (0)
#define PROF_FLUSH(lo, hi) do { if (prof_on) { for (int i_ = lo; i_ < hi; ++i_) g_mcep_prof[i_] = prof_acc[i_]; } } while (0)
#else
#define PROF_DECL
#define PROF_START()
#define PROF_LAP(i)
#define PROF_FLUSH(lo, hi)
#endif

struct McepTcParams {
  const void* in;
  int in_is_power;
  int in_vec4;       // rows are 16-byte aligned fp32: read with float4 loads
  int64_t in_stride; // elements
  int64_t num_frames;
  int K, m, NBk, nchunks;
  int ws_floats;   // per-warp solve workspace
  int miniter, maxiter;
  float threshold, eps, alpha;
  const float* stream0;  // pre-tiled [nchunks][kStageBytes]: pass 0 (M0^T in the B2 slots)
  const float* stream1;  // pre-tiled [nchunks][kStageBytes]: Newton passes (Cmat chunk, M2^T chunk)
  void* mc_out;
  int mc_dtype;
  int64_t mc_stride;
  int* iters;
  int* status;
};

// Builds both streams from the fp32 matrices of b2w_mcep_tables_host (m0t [K, np0], cmat [MP, K], m2t [K, np2]).
__global__ void mcep_tc_pretile_kernel(const float* __restrict__ m0t, int np0, const float* __restrict__ cmat, const float* __restrict__ m2t,
                                       int np2, int K, int m, int nchunks, float* __restrict__ stream0, float* __restrict__ stream1) {
  const int stage_floats = kStageBytes / 4;
  const int total = nchunks * stage_floats;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int c = e / stage_floats;
    int o = e - c * stage_floats;
    const int j0 = c * kTcBK;
    float v0 = 0.f, v1 = 0.f;  // values for stream0 / stream1 at this slot BEFORE the hi/lo selection
    bool lo;
    if (o < 2 * (int)(kB1Bytes / 4)) {
      // B1 tile: rows n = bin within chunk (16), K = cepstral index (64)
      lo = o >= (int)(kB1Bytes / 4);
      if (lo) o -= kB1Bytes / 4;
      const int kc = o / (kTcBK * 4), rem = o - kc * (kTcBK * 4);   // chunk stride = R * 16 B = R * 4 floats
      const int n = (rem >> 5) * 8 + ((rem >> 2) & 7), k = kc * 4 + (rem & 3);
      const int j = j0 + n;
      if (j < K && k <= m) v1 = cmat[(int64_t)k * K + j];
    } else {
      o -= 2 * (kB1Bytes / 4);
      lo = o >= (int)(kB2Bytes / 4);
      if (lo) o -= kB2Bytes / 4;
      const int kc = o / (kTcN2 * 4), rem = o - kc * (kTcN2 * 4);
      const int n = (rem >> 5) * 8 + ((rem >> 2) & 7), k = kc * 4 + (rem & 3);
      const int j = j0 + k;
      if (j < K) {
        if (n <= 2 * m) v1 = m2t[(int64_t)j * np2 + n];
        if (n <= m + 1) v0 = m0t[(int64_t)j * np0 + n];
      }
    }
    float h0, l0, h1, l1;
    umma::split_tf32(v0, h0, l0);
    umma::split_tf32(v1, h1, l1);
    for (int r = 0; r < kTcCopies; ++r) {
      stream0[(size_t)r * total + e] = lo ? l0 : h0;
      stream1[(size_t)r * total + e] = lo ? l1 : h1;
    }
  }
}

// Periodogram values of row `frame`, bins jb .. jb + 7 (raw input; 1 where the bin is outside the spectrum so that log / the
// zero check stay quiet).  Lanes hold different rows, so every load instruction costs one sector per lane: with padded rows
// two float4 loads replace eight scalar ones.
template <typename IT>
__device__ __forceinline__ void tc_load_raw8(const McepTcParams& p, int64_t frame, int jb, bool valid, float (&v)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 1.f;
  if (!valid) return;
  const IT* rp = reinterpret_cast<const IT*>(p.in) + frame * p.in_stride + jb;
  if (sizeof(IT) == 4 && p.in_vec4) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (jb + 4 * h < p.K) {  // the padded row covers the whole float4 (in_stride is a multiple of 4 >= K)
        const float4 q4 = __ldg(reinterpret_cast<const float4*>(rp) + h);
        v[4 * h] = q4.x; v[4 * h + 1] = q4.y; v[4 * h + 2] = q4.z; v[4 * h + 3] = q4.w;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (jb + i < p.K) v[i] = (float)rp[i];
  }
}

// NS = m + 1 when the register-resident solver is compiled for this order (mcep_solve.cuh), 0 = generic blocked solver
template <typename IT, int NS>
__global__ void __launch_bounds__(kTcThreads, 1) mcep_tc_kernel(McepTcParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  // ---- shared memory map ----------------------------------------------------------------------------------------------
  // GEMM phase:  [stage 0 48K | stage 1 48K | stage 2 48K]  (both A operands live in tensor memory)
  // solve phase: the same region holds the 16 per-warp workspaces
  uint8_t* region = smem_raw;
  uint8_t* stage_base = region;
  const uint32_t gemm_bytes = kTcNST * kStageBytes;
  const uint32_t ws_bytes = (uint32_t)kTcSolveWarps * (uint32_t)p.ws_floats * 4u;
  const uint32_t region_bytes = gemm_bytes > ws_bytes ? gemm_bytes : ws_bytes;
  float* mc = reinterpret_cast<float*>(region + region_bytes);  // [128][kTcMS] fp32
  float* al = mc + kTcF * kTcMS;                                  // [64]
  float* sv = al + kTcMP;                                         // [128]
  int* act = reinterpret_cast<int*>(sv + kTcF);                   // [128]
  int* itc = act + kTcF;                                          // [128]
  int* qcnt = itc + kTcF;                                         // [4] work counters of the lane quarters
  uint64_t* bars = reinterpret_cast<uint64_t*>(qcnt + 4);         // see below
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kTcBars);
  int* colbase = reinterpret_cast<int*>(tmem_slot + 2);            // [64] packed-column offsets of the register-resident solver
  uint16_t* tri = reinterpret_cast<uint16_t*>(colbase + 64);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, g = warp >> 2;
  const int row = 32 * q + lane;  // this thread's TMEM lane = frame within the tile
  const int K = p.K, m = p.m;
  const int64_t frame0 = (int64_t)blockIdx.x * kTcF;
  const int nvalid = (int)min((int64_t)kTcF, p.num_frames - frame0);
  uint64_t* bar_d1free = bars;                  // [2] every epilogue warp has read D1[c & 1]
  uint64_t* bar_a2 = bars + 2;                  // [2] every epilogue warp has written its part of A2[c & 1] (tensor memory)
  uint64_t* bar_full1 = bars + 4;               // [NST] Cmat half of stage [c % NST] loaded (bytes)
  uint64_t* bar_full2 = bars + 4 + kTcNST;      // [NST] M2^T half of the stage loaded (bytes)
  uint64_t* bar_g1 = bars + 4 + 2 * kTcNST;     // [NST] GEMM1 of the chunk in this stage complete: D1 ready, the Cmat half is free
  uint64_t* bar_g2 = bars + 4 + 3 * kTcNST;     // [NST] GEMM2 of the chunk in this stage complete: A2[c & 1] and the M2^T half are free

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) umma::mbar_init(&bars[i], 8);  // two half-chunk items x four warps, one elected arrival each
    for (int i = 4; i < kTcBars; ++i) umma::mbar_init(&bars[i], 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, kTmemCols);
  for (int pi = tid; pi < p.NBk * (p.NBk - 1) / 2; pi += kTcThreads) {
    int a_ = 0, qq = pi;
    while (qq > a_) { qq -= a_ + 1; ++a_; }
    tri[pi] = (uint16_t)((a_ << 8) | qq);
  }
  if (NS > 0 && tid < 64) colbase[tid] = rr_col_base(NS > 0 ? NS : 8, tid < NS - 1 ? tid : 0);
  if (tid < kTcMP) al[tid] = (tid <= m) ? powf(-p.alpha, (float)tid) : 0.f;
  if (tid == 0) al[0] = 1.f;
  if (tid < kTcF) {
    act[tid] = tid < nvalid ? 1 : 0;
    itc[tid] = 0;
    sv[tid] = 0.f;
  }
  for (int i = tid; i < kTcF * kTcMS; i += kTcThreads) mc[i] = 0.f;
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_d2 = tmem + ((uint32_t)(32 * q) << 16) + 32;   // D2: columns 32..159
  const uint32_t idesc1 = umma::idesc_tf32(kTcF, kTcBK);
  const uint32_t idesc2 = umma::idesc_tf32(kTcF, kTcN2);
  // Barrier parities are STATELESS: the barrier of slot c % n completes once per chunk with that residue, so the completion a role
  // waits for has the index (earlier passes) * (chunks per pass on that slot) + c / n, and its parity is the low bit.  No role has
  // to see every completion, and no phase word is carried through the solve phase.  (A wait is only ever issued when the previous
  // completion of the same barrier is known to have happened -- the tensor pipe completes in order and every role touches each
  // slot at most n chunks apart -- so the parity test cannot be fooled by a barrier two phases behind.)
  const int nch = p.nchunks;
  auto wait_slot = [&](uint64_t* arr, int n, int c, int passes_before) {
    const int slot = c % n;
    const int per_pass = (nch - slot + n - 1) / n;
    umma::mbar_wait(&arr[slot], (uint32_t)(passes_before * per_pass + c / n) & 1u);
  };
  bool zero_per = false;
  PROF_DECL;
  PROF_START();

  for (int pass = 0; pass <= p.maxiter; ++pass) {
    const float* stream = (pass == 0 ? p.stream0 : p.stream1) + (size_t)(blockIdx.x % kTcCopies) * ((size_t)p.nchunks * (kStageBytes / 4));
    // ---- A1 = hi/lo split of mc (Newton passes) ----------------------------------------------------------------------
    if (pass > 0) {  // thread (q, g, lane): row 32 q + lane, coefficients 16 g .. 16 g + 15 -> tensor memory columns of A1 hi / lo
      float hi[16], lo[16];
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        const float4 v = *reinterpret_cast<const float4*>(mc + row * kTcMS + 16 * g + 4 * i4);
        umma::split_tf32(v.x, hi[4 * i4], lo[4 * i4]);
        umma::split_tf32(v.y, hi[4 * i4 + 1], lo[4 * i4 + 1]);
        umma::split_tf32(v.z, hi[4 * i4 + 2], lo[4 * i4 + 2]);
        umma::split_tf32(v.w, hi[4 * i4 + 3], lo[4 * i4 + 3]);
      }
      const uint32_t t_a1 = tmem + ((uint32_t)(32 * q) << 16) + kTmA1 + 16 * g;
      umma::tmem_st16(t_a1, hi);
      umma::tmem_st16(t_a1 + kTcMP, lo);
      umma::tmem_st_wait();
    }
    umma::fence_proxy_async();  // the solve phase wrote the stage region through the generic proxy; the bulk copies come next
    umma::tc_fence_before_sync();
    __syncthreads();
    umma::tc_fence_after_sync();
    if (tid == kTcProducer) {
      // ---- producer: streams the constant matrices.  The two halves of a stage have their own barriers: the Cmat half is free as
      // soon as GEMM1 of the chunk NST back has completed (a whole chunk earlier than its GEMM2), the M2^T half when that GEMM2 has.
      for (int c = 0, s = 0; c < nch; ++c, s = (s + 1 == kTcNST ? 0 : s + 1)) {
        uint8_t* dst = stage_base + s * kStageBytes;
        const uint8_t* src = reinterpret_cast<const uint8_t*>(stream) + (size_t)c * kStageBytes;
        if (pass > 0) {
          if (c >= kTcNST) wait_slot(bar_g1, kTcNST, c - kTcNST, pass - 1);
          umma::mbar_expect_tx(&bar_full1[s], 2 * kB1Bytes);
          umma::bulk_g2s(dst, src, 2 * kB1Bytes, &bar_full1[s]);
        }
        if (c >= kTcNST) wait_slot(bar_g2, kTcNST, c - kTcNST, pass);
        umma::mbar_expect_tx(&bar_full2[s], 2 * kB2Bytes);
        umma::bulk_g2s(dst + 2 * kB1Bytes, src + 2 * kB1Bytes, 2 * kB2Bytes, &bar_full2[s]);
      }
    } else if (warp == kTcIssuer / 32) {
      // ---- issuer warp: keeps the tensor pipe fed; it never waits for an MMA to complete.  The WHOLE warp runs this branch and the
      // MMAs are issued under elect.sync: ptxas then keeps the descriptors in uniform registers and emits the tcgen05.mma back to
      // back; under a plain `if (tid == 0)` it wraps EVERY tcgen05.mma into an elect / R2UR / branch loop (~60 cycles per MMA).
      const uint32_t tm_u = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t st_u = __shfl_sync(0xffffffffu, umma::smem_u32(stage_base), 0);
      const uint32_t bar_u = __shfl_sync(0xffffffffu, umma::smem_u32(bars), 0);
      auto commit = [&](uint64_t* bar) {
        const uint32_t addr = bar_u + 8u * (uint32_t)(bar - bars);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(addr) : "memory");
      };
      auto gemm1 = [&](int c, int s) {  // D1[c & 1] = mc . Cmat chunk (stage s)
        const uint32_t b1h = st_u + s * kStageBytes, b1l = b1h + kB1Bytes, b_lbo = kTcBK * 16;
        if (umma::elect_one()) {
          umma::mma_3xtf32_ts<kTcMP / 8>(tm_u + ((c & 1) ? 160 : 0), tm_u + kTmA1, tm_u + kTmA1 + kTcMP, umma::smem_desc(b1h, b_lbo, 128),
                                         umma::smem_desc(b1l, b_lbo, 128), 2 * b_lbo, idesc1, false);
          commit(&bar_g1[s]);
        }
        __syncwarp();
      };
      PROF_LAP(0);  // outside the GEMM phase (A1 split, solves, barriers)
      if (pass > 0) {
        wait_slot(bar_full1, kTcNST, 0, pass - 1);
        umma::tc_fence_after_sync();
        gemm1(0, 0);
      }
      for (int c = 0, s = 0; c < nch; ++c) {
        const int sn = (s + 1 == kTcNST ? 0 : s + 1);
        if (pass > 0 && c + 1 < nch) {  // GEMM1 of the next chunk goes ahead of this chunk's GEMM2
          PROF_LAP(1);
          wait_slot(bar_full1, kTcNST, c + 1, pass - 1);
          PROF_LAP(2);  // wait Cmat half
          if (c + 1 >= 2) wait_slot(bar_d1free, 2, c - 1, pass - 1);  // D1[(c + 1) & 1] was last read by the epilogue of chunk c - 1
          PROF_LAP(3);  // wait D1 free
          umma::tc_fence_after_sync();
          gemm1(c + 1, sn);
          PROF_LAP(4);  // GEMM1 issue
        }
        wait_slot(bar_full2, kTcNST, c, pass);
        PROF_LAP(7);  // wait M2^T half
        wait_slot(bar_a2, 2, c, pass);  // the epilogue has written P of chunk c
        PROF_LAP(5);  // wait A2
        umma::tc_fence_after_sync();
        {
          const uint32_t b2h = st_u + s * kStageBytes + 2 * kB1Bytes, b2l = b2h + kB2Bytes, b_lbo = kTcN2 * 16;
          const uint32_t a2h = tm_u + kTmA2 + 64 * (c & 1), a2l = a2h + 32;
          if (umma::elect_one()) {
            umma::mma_3xtf32_ts<kTcBK / 8>(tm_u + 32, a2h, a2l, umma::smem_desc(b2h, b_lbo, 128), umma::smem_desc(b2l, b_lbo, 128), 2 * b_lbo,
                                           idesc2, c > 0);
            commit(&bar_g2[s]);
          }
          __syncwarp();
        }
        PROF_LAP(6);  // GEMM2 issue
        s = sn;
      }
      if (pass > 0) {  // the D1 releases of the last two chunks (nothing depends on them: every completion gets a wait)
        for (int c = (nch >= 2 ? nch - 2 : 0); c < nch; ++c) wait_slot(bar_d1free, 2, c, pass - 1);
      }
    } else if (warp >= kTcEpiWarp0) {
      // ---- epilogue: P = per * exp(-2 D1) (pass 0: log per) -> hi / lo TF32 columns of A2 in tensor memory ---------------------
      // Twelve warps = three sets of four (one warp per TMEM lane quarter).  A work item is half a chunk (128 rows x 16 bins); the
      // sets take the items round robin, so consecutive items of a set are 1.5 chunks apart.
      constexpr int CPT = kTcBK / 2;  // bins per thread and item
      static_assert(CPT == 16, "the tensor-memory load / store below move 16 columns per thread");
      const int nitems = 2 * nch;
      float pern[CPT];                // raw periodogram values, prefetched one item ahead
      int item = (warp >> 2) - 1;
#pragma unroll
      for (int h = 0; h < CPT / 8; ++h)
        tc_load_raw8<IT>(p, frame0 + row, (item >> 1) * kTcBK + CPT * (item & 1) + 8 * h, row < nvalid && item < nitems,
                         *reinterpret_cast<float(*)[8]>(pern + 8 * h));
      for (; item < nitems; item += kTcEpiSets) {
        const int c = item >> 1, eh = item & 1;
        const int s = c & 1;  // D1 / A2 slot
        const int j0 = c * kTcBK + CPT * eh;
        float perv[CPT];
#pragma unroll
        for (int i = 0; i < CPT; ++i) perv[i] = (j0 + i < K) ? (p.in_is_power ? pern[i] + p.eps : fmaf(pern[i], pern[i], p.eps)) : 1.f;
        if (item + kTcEpiSets < nitems) {
          const int nx = item + kTcEpiSets;
#pragma unroll
          for (int h = 0; h < CPT / 8; ++h)
            tc_load_raw8<IT>(p, frame0 + row, (nx >> 1) * kTcBK + CPT * (nx & 1) + 8 * h, row < nvalid, *reinterpret_cast<float(*)[8]>(pern + 8 * h));
        }
        float cv[CPT];
#pragma unroll
        for (int i = 0; i < CPT; ++i) cv[i] = 0.f;
        PROF_LAP(8);  // epilogue: loads / prefetch issue (+ everything outside the GEMM phase)
        if (pass > 0) {
          wait_slot(bar_g1, kTcNST, c, pass - 1);
          PROF_LAP(9);  // epilogue: wait GEMM1
          umma::tc_fence_after_sync();
          umma::tmem_ld16(tmem + ((uint32_t)(32 * q) << 16) + (s ? 160 : 0) + CPT * eh, cv);
          umma::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) umma::mbar_arrive(&bar_d1free[s]);
        }
        float ph[CPT], pl[CPT];
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
          if (!(perv[i] > 0.f)) zero_per = true;
          const float val = pass == 0 ? logf(perv[i]) : perv[i] * expf(-2.f * cv[i]);
          umma::split_tf32((j0 + i < K) ? val : 0.f, ph[i], pl[i]);
        }
        PROF_LAP(10);  // epilogue: tmem ld + exp + split
        if (c >= 2) wait_slot(bar_g2, kTcNST, c - 2, pass);  // A2[s] (double buffered) was last read by GEMM2 of chunk c - 2
        PROF_LAP(11);  // epilogue: wait A2 free
        {
          const uint32_t t_a2 = tmem + ((uint32_t)(32 * q) << 16) + kTmA2 + 64 * s + CPT * eh;
          umma::tmem_st16(t_a2, ph);
          umma::tmem_st16(t_a2 + 32, pl);
          umma::tmem_st_wait();
        }
        umma::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(&bar_a2[s]);
        PROF_LAP(12);  // epilogue: A2 stores + fence + arrive
      }
      if (nch >= 2) wait_slot(bar_g2, kTcNST, nch - 2, pass);  // (observed by nobody else: keeps every completion waited for)
      wait_slot(bar_g2, kTcNST, nch - 1, pass);  // the last chunk's GEMM2 completes D2
    }
    umma::tc_fence_before_sync();
    __syncthreads();  // D2 is complete (the epilogue warps and the issuer have waited for the last GEMM2)
    umma::tc_fence_after_sync();

    if (pass == 0) {
      // D2[:, 0..m] = initial mel-cepstrum, D2[:, m+1] = SPTK's start value s
      float v[16];
      umma::tmem_ld16(t_d2 + 16 * g, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int k = 16 * g + i;
        if (k <= m) mc[row * kTcMS + k] = v[i];
        if (k == m + 1) sv[row] = v[i];
      }
      umma::tc_fence_before_sync();
      __syncthreads();
      continue;
    }
    // ---- stopping rule on r~[0] (one thread per frame) -----------------------------------------------------------
    {
      float v[4];
      umma::tmem_ld4(t_d2, v);
      if (g == 0 && act[row]) {
        const float t = v[0];
        if (pass >= p.miniter) {
          if (fabsf((t - sv[row]) / t) < p.threshold) {
            act[row] = 0;
            itc[row] = pass;
          } else {
            sv[row] = t;
          }
        }
      }
      if (tid < 4) qcnt[tid] = 0;
    }
    __syncthreads();
    int any = 0;
    if (tid < kTcF) any = act[tid];
    if (!__syncthreads_or(any)) break;
    // ---- Newton step: the warps of a lane quarter share its 32 frames through a work counter -----------------------
    if (warp < kTcSolveWarps) {
      float* ws = reinterpret_cast<float*>(region) + warp * p.ws_floats;
      // generic: [blocked LDL^T workspace | r~ row 128 | x 64];  register-resident: [packed columns | 64 pad | r~ row 128]
      float* rtrow = ws + (NS > 0 ? rr_workspace_floats(NS > 0 ? NS : 8) - kTcN2 : ldl_workspace_floats(p.NBk, kTcKB));
      float* xo = rtrow + kTcN2;                                // [64] (generic solver only)
      for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(&qcnt[q], 1);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= 32) break;
        const int f = 32 * q + idx;
        if (!act[f]) continue;
        // pull row f of D2 out of tensor memory: every lane receives its own row, lane idx keeps it
#pragma unroll
        for (int cb = 0; cb < kTcN2 / 16; ++cb) {
          float v[16];
          umma::tmem_ld16(t_d2 + 16 * cb, v);
          if (lane == idx) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(rtrow + 16 * cb + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
        }
        __syncwarp();
        bool ok;
        if constexpr (NS > 0) {
          float x0, x1;
          ok = warp_rr_solve<NS>(rtrow, al, ws, colbase, x0, x1);
          if (ok) {
            if (lane < NS) mc[f * kTcMS + lane] += x0;
            if (lane + 32 < NS) mc[f * kTcMS + lane + 32] += x1;
          }
        } else {
          ok = warp_ldl_solve<kTcKB>(rtrow, al, m + 1, p.NBk, tri, ws, xo);
          if (ok) {
            for (int k = lane; k <= m; k += 32) mc[f * kTcMS + k] += xo[k];
          }
        }
        if (!ok && lane == 0) {
          atomicOr(p.status, B2W_STATUS_SOLVE_FAILED);
          act[f] = 0;
          itc[f] = pass;
        }
        __syncwarp();
      }
    }
    umma::tc_fence_before_sync();
    __syncthreads();
  }
#ifdef B2W_MCEP_PROF
  if (tid == kTcIssuer) PROF_FLUSH(0, 8);
  if (tid == 128) PROF_FLUSH(8, 14);
#endif
  if (zero_per) atomicOr(p.status, B2W_STATUS_ZERO_PERIODOGRAM);
  if (tid < kTcF && act[tid]) {
    itc[tid] = p.maxiter;
    atomicOr(p.status, B2W_STATUS_NOT_CONVERGED);
  }
  __syncthreads();
  for (int i = tid; i < nvalid * (m + 1); i += kTcThreads) {
    const int f = i / (m + 1), k = i - f * (m + 1);
    const float v = mc[f * kTcMS + k];
    if (p.mc_dtype == B2W_F64) reinterpret_cast<double*>(p.mc_out)[(frame0 + f) * p.mc_stride + k] = (double)v;
    else reinterpret_cast<float*>(p.mc_out)[(frame0 + f) * p.mc_stride + k] = v;
  }
  if (p.iters && tid < nvalid) p.iters[frame0 + tid] = itc[tid];
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, kTmemCols);
}

}  // namespace b2w

extern "C" int b2w_mcep_prof_read(long long* out16) {
  return (int)cudaMemcpyFromSymbol(out16, b2w::g_mcep_prof, sizeof(long long) * 16);
}

extern "C" int64_t b2w_mcep_tc_stream_floats(int32_t fft_size) {
  const int K = fft_size / 2 + 1;
  const int nchunks = (K + b2w::kTcBK - 1) / b2w::kTcBK;
  return (int64_t)nchunks * (b2w::kStageBytes / 4) * b2w::kTcCopies;
}

extern "C" int b2w_mcep_tc_pretile(int32_t order, int32_t fft_size, const float* m0t, const float* cmat, const float* m2t,
                                   float* stream0, float* stream1, void* stream) {
  using namespace b2w;
  B2W_REQUIRE(m0t && cmat && m2t && stream0 && stream1, "b2w_mcep_tc_pretile: null argument");
  B2W_REQUIRE(order >= 1 && order <= 62, "b2w_mcep_tc_pretile: order %d out of range [1, 62]", order);
  const int K = fft_size / 2 + 1;
  const int nchunks = (K + kTcBK - 1) / kTcBK;
  mcep_tc_pretile_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(m0t, pad4(order + 2), cmat, m2t, pad4(2 * order + 1), K, order, nchunks,
                                                                 stream0, stream1);
  return check_launch("mcep_tc_pretile_kernel");
}

extern "C" int b2w_mcep_tc(const void* in, int32_t in_dtype, int32_t in_is_power, int64_t in_stride, int64_t num_frames,
                           int32_t fft_size, int32_t order,
                           double alpha, int32_t miniter, int32_t maxiter, double threshold, double eps, const float* stream0,
                           const float* stream1, void* mc, int32_t mc_dtype, int64_t mc_stride, int32_t* iters, int32_t* status,
                           void* stream) {
  using namespace b2w;
  B2W_REQUIRE(in && stream0 && stream1 && mc && status, "b2w_mcep_tc: null argument");
  B2W_REQUIRE(in_dtype == B2W_F64 || in_dtype == B2W_F32, "b2w_mcep_tc: bad in_dtype %d", in_dtype);
  B2W_REQUIRE(mc_dtype == B2W_F64 || mc_dtype == B2W_F32, "b2w_mcep_tc: bad mc_dtype %d", mc_dtype);
  B2W_REQUIRE(order >= 1 && order <= 59, "b2w_mcep_tc: order %d out of range [1, 59] (use b2w_mcep)", order);
  B2W_REQUIRE(fft_size >= 64 && (fft_size & (fft_size - 1)) == 0, "b2w_mcep_tc: bad fft_size %d", fft_size);
  B2W_REQUIRE(mc_stride >= order + 1 && maxiter >= 1 && miniter >= 1, "b2w_mcep_tc: bad stride / iteration limits");
  B2W_REQUIRE(in_stride >= fft_size / 2 + 1, "b2w_mcep_tc: in_stride %lld < fft_size/2+1", (long long)in_stride);
  if (num_frames == 0) return 0;
  McepTcParams p;
  p.in = in; p.in_is_power = in_is_power; p.num_frames = num_frames;
  p.in_stride = in_stride;
  p.in_vec4 = (in_dtype == B2W_F32 && in_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0) ? 1 : 0;
  p.K = fft_size / 2 + 1; p.m = order; p.NBk = (order + 1 + 3) / 4;
  p.nchunks = (p.K + kTcBK - 1) / kTcBK;
  p.ws_floats = ldl_workspace_floats(p.NBk, kTcKB) + kTcN2 + kTcMP;
  if (order + 1 == 20 || order + 1 == 40 || order + 1 == 60) p.ws_floats = rr_workspace_floats(order + 1);  // register-resident solver
  p.miniter = miniter; p.maxiter = maxiter; p.threshold = (float)threshold; p.eps = (float)eps; p.alpha = (float)alpha;
  p.stream0 = stream0; p.stream1 = stream1; p.mc_out = mc; p.mc_dtype = mc_dtype; p.mc_stride = mc_stride;
  p.iters = iters; p.status = status;
  const uint32_t gemm_bytes = kTcNST * kStageBytes;
  const uint32_t ws_bytes = (uint32_t)kTcSolveWarps * (uint32_t)p.ws_floats * 4u;
  const uint32_t region_bytes = gemm_bytes > ws_bytes ? gemm_bytes : ws_bytes;
  const size_t smem = region_bytes + sizeof(float) * (kTcF * kTcMS + kTcMP + kTcF) + sizeof(int) * (2 * kTcF + 4) + 8 * kTcBars + 8 + 64 * sizeof(int) +
                      sizeof(uint16_t) * (size_t)(p.NBk * (p.NBk - 1) / 2 + 2) + 16;
  B2W_REQUIRE(smem <= 227 * 1024, "b2w_mcep_tc: %zu bytes of shared memory needed", smem);
  const int64_t grid = (num_frames + kTcF - 1) / kTcF;
  B2W_REQUIRE(grid < ((int64_t)1 << 31), "b2w_mcep_tc: too many frames in one call");
  cudaStream_t st = (cudaStream_t)stream;
#define B2W_TC_LAUNCH(IT, NS)                                                                                  \
  do {                                                                                                         \
    cudaFuncSetAttribute(mcep_tc_kernel<IT, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
    mcep_tc_kernel<IT, NS><<<(unsigned)grid, kTcThreads, smem, st>>>(p);                                       \
  } while (0)
#define B2W_TC_DISPATCH(IT)                        \
  do {                                             \
    if (order + 1 == 60) B2W_TC_LAUNCH(IT, 60);    \
    else if (order + 1 == 40) B2W_TC_LAUNCH(IT, 40); \
    else if (order + 1 == 20) B2W_TC_LAUNCH(IT, 20); \
    else B2W_TC_LAUNCH(IT, 0);                     \
  } while (0)
  if (in_dtype == B2W_F64) B2W_TC_DISPATCH(double);
  else B2W_TC_DISPATCH(float);
#undef B2W_TC_DISPATCH
#undef B2W_TC_LAUNCH
  return check_launch("mcep_tc_kernel");
}
