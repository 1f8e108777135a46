// Mel-cepstral analysis (SPTK mcep Newton loop) with the dense all-pass-warp contractions on the 5th-generation tensor
// cores: tcgen05.mma kind::tf32 with the 3xTF32 split (fp32-accurate), accumulators in tensor memory, constant matrices
// streamed by 1-D bulk async copies (mbarrier completion) from a pre-tiled global stream.
//
// Same mathematics and the same reference call site as mcep.cu (pysptk.mcep, AudioProcessing.py:146); this is the
// production path for order <= 63.  One CTA = 128 frames (the UMMA M dimension), 512 threads.
//
//   per Newton iteration, for each chunk of 32 frequency bins j0 .. j0+31 (17 chunks at K = 513):
//     GEMM1   D1[128 x 32]   = mc[128 x 64] . Cmat[64 x 32 chunk]          24 MMAs (8 K-steps x 3 split products)
//     epilogue (all threads)  P = per * exp(-2 D1)  -> hi/lo TF32 tiles in shared memory (A operand of GEMM2)
//     GEMM2   D2[128 x 128] += P[128 x 32] . M2^T[32 chunk x 128]          12 MMAs
//   The GEMM phase is warp specialised: thread 0 only streams the matrices and issues MMAs (GEMM1 of chunk c + 1 is in the tensor
//   pipe while chunk c's epilogue runs: D1 is double buffered in tensor memory), warps 4-11 are the epilogue (one TMEM lane =
//   one frame per thread, 16 bins each), everything is handed over through mbarriers -- no CTA barrier inside the chunk loop.
//   then D2 row f = r~ of frame f: stopping rule on r~[0], and for the frames still iterating one warp each builds and
//   solves the (m+1) x (m+1) Toeplitz-plus-Hankel system (mcep_solve.cuh) and updates mc (fp32, shared memory).
//   Pass 0 (initial value) uses the same machinery: P = log(per), the stream holds M0^T instead of M2^T, GEMM1 is skipped.
#include "common.cuh"
#include "mcep_solve.cuh"
#include "umma.cuh"

namespace b2w {

constexpr int kTcF = 128;        // frames per CTA = UMMA M
#ifdef B2W_TC_ISSUER_WARP            // experiment: a 17th warp issues, all 16 others are epilogue warps (120 registers / thread)
constexpr int kTcThreads = 544;
constexpr int kTcIssuer = 512;
constexpr int kTcProducer = 513;
constexpr int kTcEpiWarp0 = 0;
constexpr int kTcEpiThreads = 512;
#else
constexpr int kTcThreads = 512;  // 16 warps: TMEM lane quarter q = warp & 3, column group g = warp >> 2
constexpr int kTcIssuer = 0;     // the thread that issues the MMAs
constexpr int kTcProducer = 32;  // the thread that streams the constant matrices (bulk async copies)
constexpr int kTcEpiWarp0 = 4;   // epilogue warps 4 .. 11 (two per TMEM lane quarter)
constexpr int kTcEpiThreads = 256;
#endif
constexpr int kTcSolveWarps = 16;
constexpr int kTcBK = 32;        // bins per chunk
constexpr int kTcMP = 64;        // padded cepstral dimension (K of GEMM1)
constexpr int kTcN2 = 128;       // padded r~ length (N of GEMM2)
constexpr int kTcKB = 20;        // padded block stride of the solve workspace (bank-conflict free float4 accesses)
constexpr uint32_t kB1Bytes = kTcBK * kTcMP * 4;   // one of hi / lo of the Cmat chunk   [N = 16 rows (bins)] x [K = 64]
constexpr uint32_t kB2Bytes = kTcN2 * kTcBK * 4;   // one of hi / lo of the M2^T chunk   [N = 128 rows]       x [K = 16]
constexpr uint32_t kStageBytes = 2 * kB1Bytes + 2 * kB2Bytes;  // 48 KB per chunk: [B1 hi | B1 lo | B2 hi | B2 lo]
constexpr uint32_t kA1Bytes = kTcF * kTcMP * 4;    // 32 KB, one of hi / lo
constexpr uint32_t kA2Bytes = kTcF * kTcBK * 4;    // 16 KB, one of hi / lo
constexpr int kTcBars = 12;
constexpr int kTmemCols = 256;                     // D1[0] at columns 0..31, D2 at columns 32..159, D1[1] at columns 160..191

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
    stream0[e] = lo ? l0 : h0;
    stream1[e] = lo ? l1 : h1;
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
  // GEMM phase:  [A1 hi 32K | A1 lo 32K | stage 0 48K | stage 1 48K | A2 hi 16K | A2 lo 16K]  = 192 KB
  // solve phase: the same region holds the 16 per-warp workspaces
  uint8_t* region = smem_raw;
  float* a1_hi = reinterpret_cast<float*>(region);
  float* a1_lo = reinterpret_cast<float*>(region + kA1Bytes);
  uint8_t* stage_base = region + 2 * kA1Bytes;
  float* a2_hi = reinterpret_cast<float*>(stage_base + 2 * kStageBytes);
  float* a2_lo = reinterpret_cast<float*>(stage_base + 2 * kStageBytes + kA2Bytes);
  const uint32_t gemm_bytes = 2 * kA1Bytes + 2 * kStageBytes + 2 * kA2Bytes;
  const uint32_t ws_bytes = (uint32_t)kTcSolveWarps * (uint32_t)p.ws_floats * 4u;
  const uint32_t region_bytes = gemm_bytes > ws_bytes ? gemm_bytes : ws_bytes;
  float* mc = reinterpret_cast<float*>(region + region_bytes);  // [128][64] fp32
  float* al = mc + kTcF * kTcMP;                                  // [64]
  float* sv = al + kTcMP;                                         // [128]
  int* act = reinterpret_cast<int*>(sv + kTcF);                   // [128]
  int* itc = act + kTcF;                                          // [128]
  int* qcnt = itc + kTcF;                                         // [4] work counters of the lane quarters
  uint64_t* bars = reinterpret_cast<uint64_t*>(qcnt + 4);         // full_b1[2], g1[2], d1_free[2], a2_full, -, g2[2], full_b2[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kTcBars);
  int* colbase = reinterpret_cast<int*>(tmem_slot + 2);            // [64] packed-column offsets of the register-resident solver
  uint16_t* tri = reinterpret_cast<uint16_t*>(colbase + 64);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, g = warp >> 2;
  const int row = 32 * q + lane;  // this thread's TMEM lane = frame within the tile
  const int K = p.K, m = p.m;
  const int64_t frame0 = (int64_t)blockIdx.x * kTcF;
  const int nvalid = (int)min((int64_t)kTcF, p.num_frames - frame0);
  uint64_t* bar_full1 = bars;       // [2] Cmat half of stage [slot] loaded (bytes)
  uint64_t* bar_g1 = bars + 2;      // [2] GEMM1 into D1[slot] complete: the Cmat half of the stage is free
  uint64_t* bar_d1free = bars + 4;  // [2] every epilogue thread has read D1[slot]
  uint64_t* bar_a2 = bars + 6;      //     every epilogue thread has written its part of A2
  uint64_t* bar_g2 = bars + 8;      // [2] GEMM2 of a chunk in stage [slot] complete: A2 and the M2^T half are free, D2 accumulated
  uint64_t* bar_full2 = bars + 10;  // [2] M2^T half of stage [slot] loaded (bytes)

  if (tid == 0) {
    umma::mbar_init(&bar_full1[0], 1);
    umma::mbar_init(&bar_full1[1], 1);
    umma::mbar_init(&bar_full2[0], 1);
    umma::mbar_init(&bar_full2[1], 1);
    umma::mbar_init(&bar_g1[0], 1);
    umma::mbar_init(&bar_g1[1], 1);
    umma::mbar_init(&bar_d1free[0], kTcEpiThreads);
    umma::mbar_init(&bar_d1free[1], kTcEpiThreads);
    umma::mbar_init(bar_a2, kTcEpiThreads);
    umma::mbar_init(&bar_g2[0], 1);
    umma::mbar_init(&bar_g2[1], 1);
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
  for (int i = tid; i < kTcF * kTcMP; i += kTcThreads) mc[i] = 0.f;
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_d2 = tmem + ((uint32_t)(32 * q) << 16) + 32;   // D2: columns 32..159
  const uint32_t idesc1 = umma::idesc_tf32(kTcF, kTcBK);
  const uint32_t idesc2 = umma::idesc_tf32(kTcF, kTcN2);
  // barrier parities: every completion of a barrier is consumed by exactly one wait of each role that uses it
  uint32_t ph_full1[2] = {0, 0}, ph_full2[2] = {0, 0}, ph_g1[2] = {0, 0}, ph_d1free[2] = {0, 0}, ph_a2 = 0, ph_g2[2] = {0, 0};
  const bool is_epi = warp >= kTcEpiWarp0 && warp < kTcEpiWarp0 + kTcEpiThreads / 32;
  const int eh = (warp - kTcEpiWarp0) >> 2;  // epilogue column group: bins CPT eh .. CPT eh + CPT - 1 of the chunk
  bool zero_per = false;
  PROF_DECL;
  PROF_START();

  for (int pass = 0; pass <= p.maxiter; ++pass) {
    const float* stream = pass == 0 ? p.stream0 : p.stream1;
    // ---- A1 = hi/lo split of mc (Newton passes) ----------------------------------------------------------------------
    if (pass > 0) {
      for (int e = tid; e < kTcF * (kTcMP / 4); e += kTcThreads) {
        const int r = e & (kTcF - 1), kc = e >> 7;  // consecutive threads -> consecutive rows: conflict-free 16 B stores
        const float4 v = *reinterpret_cast<const float4*>(mc + r * kTcMP + 4 * kc);
        float4 h, l;
        umma::split_tf32(v.x, h.x, l.x);
        umma::split_tf32(v.y, h.y, l.y);
        umma::split_tf32(v.z, h.z, l.z);
        umma::split_tf32(v.w, h.w, l.w);
        const uint32_t off = umma::tile_off(kTcF, r, 4 * kc) / 4;
        *reinterpret_cast<float4*>(a1_hi + off) = h;
        *reinterpret_cast<float4*>(a1_lo + off) = l;
      }
      umma::fence_proxy_async();
    }
    __syncthreads();
    const int nch = p.nchunks;
    if (tid == kTcIssuer) {
      // ---- issuer: stream the matrices, keep the tensor pipe fed ------------------------------------------------------------
      auto load = [&](int c) {
        const int s = c & 1;
        umma::mbar_expect_tx(&bar_full[s], kStageBytes);
        umma::bulk_g2s(stage_base + s * kStageBytes, reinterpret_cast<const uint8_t*>(stream) + (size_t)c * kStageBytes, kStageBytes,
                       &bar_full[s]);
      };
      const uint32_t a_lbo = kTcF * 16;
      auto gemm1 = [&](int c) {  // D1[c & 1] = mc . Cmat chunk
        const int s = c & 1;
        const uint32_t b1h = umma::smem_u32(stage_base + s * kStageBytes), b1l = b1h + kB1Bytes, b_lbo = kTcBK * 16;
        umma::mma_3xtf32<kTcMP / 8>(tmem + (s ? 160 : 0), umma::smem_desc(umma::smem_u32(a1_hi), a_lbo, 128),
                                    umma::smem_desc(umma::smem_u32(a1_lo), a_lbo, 128), umma::smem_desc(b1h, b_lbo, 128),
                                    umma::smem_desc(b1l, b_lbo, 128), 2 * a_lbo, 2 * b_lbo, idesc1, false);
        umma::mma_commit(&bar_g1[s]);
      };
      PROF_LAP(0);  // outside the GEMM phase (A1 split, solves, barriers)
      for (int c = 0; c < 2 && c < nch; ++c) load(c);
      if (pass > 0) {
        umma::mbar_wait(&bar_full[0], ph_full[0]);
        ph_full[0] ^= 1;
        umma::tc_fence_after_sync();
        gemm1(0);
      }
      for (int c = 0; c < nch; ++c) {
        const int s = c & 1;
        if (pass > 0) {
          if (c + 1 < nch) {  // GEMM1 of the next chunk goes ahead of this chunk's GEMM2
            const int sn = (c + 1) & 1;
            PROF_LAP(1);
            umma::mbar_wait(&bar_full[sn], ph_full[sn]);
            ph_full[sn] ^= 1;
            PROF_LAP(2);  // wait stage
            if (c + 1 >= 2) {  // D1[sn] was last read by the epilogue of chunk c - 1
              umma::mbar_wait(&bar_d1free[sn], ph_d1free[sn]);
              ph_d1free[sn] ^= 1;
            }
            PROF_LAP(3);  // wait D1 free
            umma::tc_fence_after_sync();
            gemm1(c + 1);
            PROF_LAP(4);  // GEMM1 issue
          }
        } else {
          umma::mbar_wait(&bar_full[s], ph_full[s]);
          ph_full[s] ^= 1;
        }
        umma::mbar_wait(bar_a2, ph_a2);  // the epilogue has written P of chunk c
        ph_a2 ^= 1;
        PROF_LAP(5);  // wait A2
        umma::tc_fence_after_sync();
        {
          const uint32_t b2h = umma::smem_u32(stage_base + s * kStageBytes) + 2 * kB1Bytes, b2l = b2h + kB2Bytes, b_lbo = kTcN2 * 16;
          umma::mma_3xtf32<kTcBK / 8>(tmem + 32, umma::smem_desc(umma::smem_u32(a2_hi), a_lbo, 128),
                                      umma::smem_desc(umma::smem_u32(a2_lo), a_lbo, 128), umma::smem_desc(b2h, b_lbo, 128),
                                      umma::smem_desc(b2l, b_lbo, 128), 2 * a_lbo, 2 * b_lbo, idesc2, c > 0);
          umma::mma_commit(bar_g2);
        }
        PROF_LAP(6);  // GEMM2 issue
        umma::mbar_wait(bar_g2, ph_g2);  // GEMM2(c) complete: stage s is free
        ph_g2 ^= 1;
        PROF_LAP(7);  // wait GEMM2
        if (c + 2 < nch) load(c + 2);
      }
      if (pass > 0) {  // consume the D1 releases of the last two chunks (keeps the parities in step)
        for (int c = (nch >= 2 ? nch - 2 : 0); c < nch; ++c) {
          umma::mbar_wait(&bar_d1free[c & 1], ph_d1free[c & 1]);
          ph_d1free[c & 1] ^= 1;
        }
      }
    } else if (is_epi) {
      // ---- epilogue warps: P = per * exp(-2 D1) (pass 0: log per) -> hi / lo TF32 tiles of A2 ------------------------------------
      constexpr int CPT = kTcBK / (kTcEpiThreads / 128);  // bins per thread (16 with 8 epilogue warps)
      float pern[CPT];                // raw periodogram values, prefetched one chunk ahead
#pragma unroll
      for (int h = 0; h < CPT / 8; ++h)
        tc_load_raw8<IT>(p, frame0 + row, CPT * eh + 8 * h, row < nvalid, *reinterpret_cast<float(*)[8]>(pern + 8 * h));
      for (int c = 0; c < nch; ++c) {
        const int s = c & 1;
        const int j0 = c * kTcBK + CPT * eh;
        float perv[CPT];
#pragma unroll
        for (int i = 0; i < CPT; ++i) perv[i] = (j0 + i < K) ? (p.in_is_power ? pern[i] + p.eps : fmaf(pern[i], pern[i], p.eps)) : 1.f;
        if (c + 1 < nch) {
#pragma unroll
          for (int h = 0; h < CPT / 8; ++h)
            tc_load_raw8<IT>(p, frame0 + row, j0 + kTcBK + 8 * h, row < nvalid, *reinterpret_cast<float(*)[8]>(pern + 8 * h));
        }
        float cv[CPT];
#pragma unroll
        for (int i = 0; i < CPT; ++i) cv[i] = 0.f;
        PROF_LAP(8);  // epilogue: loads / prefetch issue (+ everything outside the GEMM phase)
        if (pass > 0) {
          umma::mbar_wait(&bar_g1[s], ph_g1[s]);
          ph_g1[s] ^= 1;
          PROF_LAP(9);  // epilogue: wait GEMM1
          umma::tc_fence_after_sync();
          if constexpr (CPT == 16) umma::tmem_ld16(tmem + ((uint32_t)(32 * q) << 16) + (s ? 160 : 0) + CPT * eh, cv);
          else umma::tmem_ld8(tmem + ((uint32_t)(32 * q) << 16) + (s ? 160 : 0) + CPT * eh, cv);
          umma::tc_fence_before_sync();
          umma::mbar_arrive(&bar_d1free[s]);
        }
        float ph[CPT], pl[CPT];
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
          if (!(perv[i] > 0.f)) zero_per = true;
          const float val = pass == 0 ? logf(perv[i]) : perv[i] * expf(-2.f * cv[i]);
          umma::split_tf32((j0 + i < K) ? val : 0.f, ph[i], pl[i]);
        }
        PROF_LAP(10);  // epilogue: tmem ld + exp + split
        if (c > 0) {  // A2 is single-buffered: GEMM2 of the previous chunk must have consumed it
          umma::mbar_wait(bar_g2, ph_g2);
          ph_g2 ^= 1;
        }
        PROF_LAP(11);  // epilogue: wait A2 free (GEMM2 of the previous chunk)
#pragma unroll
        for (int h4 = 0; h4 < CPT / 4; ++h4) {
          const uint32_t off = umma::tile_off(kTcF, row, CPT * eh + 4 * h4) / 4;
          *reinterpret_cast<float4*>(a2_hi + off) = make_float4(ph[4 * h4], ph[4 * h4 + 1], ph[4 * h4 + 2], ph[4 * h4 + 3]);
          *reinterpret_cast<float4*>(a2_lo + off) = make_float4(pl[4 * h4], pl[4 * h4 + 1], pl[4 * h4 + 2], pl[4 * h4 + 3]);
        }
        umma::fence_proxy_async();
        umma::mbar_arrive(bar_a2);
        PROF_LAP(12);  // epilogue: A2 stores + fence + arrive
      }
      umma::mbar_wait(bar_g2, ph_g2);  // the last chunk's GEMM2 completes D2
      ph_g2 ^= 1;
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
        if (k <= m) mc[row * kTcMP + k] = v[i];
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
            if (lane < NS) mc[f * kTcMP + lane] += x0;
            if (lane + 32 < NS) mc[f * kTcMP + lane + 32] += x1;
          }
        } else {
          ok = warp_ldl_solve<kTcKB>(rtrow, al, m + 1, p.NBk, tri, ws, xo);
          if (ok) {
            for (int k = lane; k <= m; k += 32) mc[f * kTcMP + k] += xo[k];
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
    const float v = mc[f * kTcMP + k];
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
  return (int64_t)nchunks * (b2w::kStageBytes / 4);
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
  B2W_REQUIRE(order >= 1 && order <= 62, "b2w_mcep_tc: order %d out of range [1, 62] (use b2w_mcep)", order);
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
  if (order + 1 == 20 || order + 1 == 40 || order + 1 == 60) p.ws_floats = max(p.ws_floats, rr_workspace_floats(order + 1));
  p.miniter = miniter; p.maxiter = maxiter; p.threshold = (float)threshold; p.eps = (float)eps; p.alpha = (float)alpha;
  p.stream0 = stream0; p.stream1 = stream1; p.mc_out = mc; p.mc_dtype = mc_dtype; p.mc_stride = mc_stride;
  p.iters = iters; p.status = status;
  const uint32_t gemm_bytes = 2 * kA1Bytes + 2 * kStageBytes + 2 * kA2Bytes;
  const uint32_t ws_bytes = (uint32_t)kTcSolveWarps * (uint32_t)p.ws_floats * 4u;
  const uint32_t region_bytes = gemm_bytes > ws_bytes ? gemm_bytes : ws_bytes;
  const size_t smem = region_bytes + sizeof(float) * (kTcF * kTcMP + kTcMP + kTcF) + sizeof(int) * (2 * kTcF + 4) + 8 * 8 + 8 + 64 * sizeof(int) +
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
