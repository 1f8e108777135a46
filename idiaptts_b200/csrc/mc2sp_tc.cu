// Mel-cepstrum -> spectrum on the 5th-generation tensor cores: out[f][j] = (exp?)(scale * sum_k mc[f][k] Cmat[k][j]).
//
// Replaces Re pysptk.mgc2sp(mc, alpha, gamma = 0, fftlen) + np.exp as reached from AudioProcessing.mcep_to_amp_sp
// (idiaptts/src/data_preparation/audio/AudioProcessing.py:248-257) through decode_sp (:304-327) on the batched synthesis path
// (Synthesiser.run_world_synth, idiaptts/src/Synthesiser.py:39-80).  Same mathematics as mc2sp_kernel (mcep.cu, CUDA cores: 0.86 ms
// per 328 k frames, fp32-FMA bound): the all-pass warp + real FFT is the constant matrix Cmat [60 x 513], so a tile of 128 frames is
// the GEMM  D[128 x 32 bins] = mc[128 x 64] . Cmat chunk  per 32-bin chunk -- exactly GEMM 1 of mcep_tc_kernel, and it reads the same
// pre-tiled hi / lo TF32 stream (b2w_mcep_tc_pretile; only the Cmat half of every stage is fetched).  3xTF32 (fp32-accurate:
// exp() amplifies the error of the exponent by |C| ~ 10).
//   warp 0        issues the 24 TS-form MMAs of a chunk under elect.sync (A = mc hi / lo in tensor memory, written once per tile)
//   warp 1        producer: 16 KB bulk copies of the Cmat chunks into a three-stage ring
//   warps 2 - 9   load + split the mc rows of the tile (thread = row x half of the coefficients), then per chunk: tcgen05.ld of
//                 the accumulator (double buffered), exp, a padded staging tile, and the rows leave as 128-byte coalesced stores
// Two CTAs per SM (192 of 512 TMEM columns, 83 KB of shared memory each) cover each other's per-tile bubbles.  All barriers are
// mbarriers indexed by the CTA's running chunk counter (stateless parities).
#include "common.cuh"
#include "umma.cuh"

namespace b2w {
namespace {

constexpr int kF = 128;                 // frames per tile = UMMA M
constexpr int kBK = 32;                 // bins per chunk (the chunking of the mcep_tc stream)
constexpr int kMP = 64;                 // padded cepstral dimension
constexpr int kNST = 3;                 // stages of the Cmat ring
constexpr int kThreads = 320;
constexpr int kEpiWarps = 8;
constexpr uint32_t kB1Bytes = kBK * kMP * 4;                      // one of hi / lo of a Cmat chunk
constexpr uint32_t kStageStride = 2 * kB1Bytes + 2 * 128 * kBK * 4;  // stride of a stage in the mcep_tc stream (48 KB)
constexpr int kStgStride = kBK + 1;     // staging row stride (floats): conflict-free column writes and row reads
struct Smem {
  static constexpr uint32_t b = 0;                                   // [NST][hi 8 KB | lo 8 KB]
  static constexpr uint32_t stg = kNST * 2 * kB1Bytes;               // [2][128][33] floats
  static constexpr uint32_t bars = stg + 2 * kF * kStgStride * 4;    // full[NST], empty[NST], a_full, d_full[2], d_free[2]
  static constexpr uint32_t misc = bars + (2 * kNST + 5) * 8;
  static constexpr uint32_t total = misc + 16;
  static_assert(bars % 8 == 0, "mbarrier alignment");
};
constexpr int kTmA = 0;      // mc hi at 0..63, lo at 64..127
constexpr int kTmD = 128;    // D[b] at 128 + 32 b

struct Params {
  const void* mc;
  int mc_is_f64;
  int64_t mc_stride, num_frames;
  int K, m, nchunks;
  float scale;
  int do_exp;
  float* out;
  const float* stream1;
};

__global__ void __launch_bounds__(kThreads, 2) mc2sp_tc_kernel(Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* stg = reinterpret_cast<float*>(smem + Smem::stg);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
  uint64_t* bar_full = bars;                 // [NST] Cmat chunk landed (bytes)
  uint64_t* bar_empty = bars + kNST;         // [NST] the MMAs that read the stage have completed (tcgen05.commit)
  uint64_t* bar_afull = bars + 2 * kNST;     // mc hi / lo of the tile written (8 warps)
  uint64_t* bar_dfull = bar_afull + 1;       // [2] the MMAs of the chunk in D[b] have completed
  uint64_t* bar_dfree = bar_dfull + 2;       // [2] every epilogue warp has read D[b]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Smem::misc);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t tiles = (p.num_frames + kF - 1) / kF;
  const int nch = p.nchunks;
  if ((int64_t)blockIdx.x >= tiles) return;
  const int my_tiles = (int)((tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);

  if (tid == 0) {
    for (int i = 0; i < 2 * kNST; ++i) umma::mbar_init(&bars[i], 1);
    umma::mbar_init(bar_afull, kEpiWarps);
    umma::mbar_init(&bar_dfull[0], 1);
    umma::mbar_init(&bar_dfull[1], 1);
    umma::mbar_init(&bar_dfree[0], kEpiWarps);
    umma::mbar_init(&bar_dfree[1], kEpiWarps);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, 256);
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  auto wait_idx = [&](uint64_t* bar, int index) { umma::mbar_wait(bar, (uint32_t)index & 1u); };

  if (warp == 1) {
    // ---- producer -----------------------------------------------------------------------------------------------------------------
    if (lane == 0) {
      for (int t = 0, g = 0; t < my_tiles; ++t) {
        for (int c = 0; c < nch; ++c, ++g) {
          const int s = g % kNST;
          if (g >= kNST) wait_idx(&bar_empty[s], g / kNST - 1);
          umma::mbar_expect_tx(&bar_full[s], 2 * kB1Bytes);
          umma::bulk_g2s(smem + Smem::b + s * 2 * kB1Bytes, reinterpret_cast<const uint8_t*>(p.stream1) + (size_t)c * kStageStride, 2 * kB1Bytes,
                         &bar_full[s]);
        }
      }
    }
  } else if (warp == 0) {
    // ---- issuer (whole warp converged, operands warp-uniform, MMAs under elect.sync: see mcep_tc.cu) ----------------------------------
    const uint32_t tm_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t sm_u = __shfl_sync(0xffffffffu, umma::smem_u32(smem), 0);
    const uint32_t idesc = umma::idesc_tf32(kF, kBK);
    auto commit_to = [&](uint64_t* bar) {
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       sm_u + (uint32_t)Smem::bars + 8u * (uint32_t)(bar - bars))
                   : "memory");
    };
    for (int t = 0, g = 0; t < my_tiles; ++t) {
      wait_idx(bar_afull, t);
      for (int c = 0; c < nch; ++c, ++g) {
        const int s = g % kNST, b = g & 1;
        wait_idx(&bar_full[s], g / kNST);
        if (g >= 2) wait_idx(&bar_dfree[b], (g >> 1) - 1);
        umma::tc_fence_after_sync();
        if (umma::elect_one()) {
          const uint32_t bh = sm_u + Smem::b + s * 2 * kB1Bytes, b_lbo = kBK * 16;
          umma::mma_3xtf32_ts<kMP / 8>(tm_u + kTmD + kBK * b, tm_u + kTmA, tm_u + kTmA + kMP, umma::smem_desc(bh, b_lbo, 128),
                                       umma::smem_desc(bh + kB1Bytes, b_lbo, 128), 2 * b_lbo, idesc, false);
          commit_to(&bar_dfull[b]);
          commit_to(&bar_empty[s]);
        }
        __syncwarp();
      }
    }
  } else {
    // ---- loaders / epilogue ----------------------------------------------------------------------------------------------------
    const int ew = warp - 2;                   // 0 .. 7
    const int q = warp & 3;                    // TMEM lane quarter of this warp
    const int half = ew >> 2;                  // coefficients 32 half .. 32 half + 31 of the row; accumulator columns 16 half .. + 15
    const int row = 32 * q + lane;
    const uint32_t t_row = tmem + ((uint32_t)(32 * q) << 16);
    for (int t = 0, g = 0; t < my_tiles; ++t) {
      const int64_t frame0 = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * kF;
      const int nvalid = (int)min((int64_t)kF, p.num_frames - frame0);
      {  // this thread's 32 coefficients -> hi / lo TF32 -> tensor memory (the MMAs of the previous tile have completed: its last
         // accumulator has been waited for below)
        float v[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = 0.f;
        if (row < nvalid) {
          const int64_t base = (frame0 + row) * p.mc_stride + 32 * half;
          if (p.mc_is_f64) {
            const double* src = reinterpret_cast<const double*>(p.mc) + base;
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (32 * half + e <= p.m) v[e] = (float)src[e];
          } else {
            const float* src = reinterpret_cast<const float*>(p.mc) + base;
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (32 * half + e <= p.m) v[e] = src[e];
          }
        }
        float hi[16], lo[16];
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
#pragma unroll
          for (int e = 0; e < 16; ++e) umma::split_tf32(v[16 * cb + e], hi[e], lo[e]);
          umma::tmem_st16(t_row + kTmA + 32 * half + 16 * cb, hi);
          umma::tmem_st16(t_row + kTmA + kMP + 32 * half + 16 * cb, lo);
        }
        umma::tmem_st_wait();
        umma::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(bar_afull);
      }
      for (int c = 0; c < nch; ++c, ++g) {
        const int b = g & 1;
        wait_idx(&bar_dfull[b], g >> 1);
        umma::tc_fence_after_sync();
        float d[16];
        umma::tmem_ld16(t_row + kTmD + kBK * b + 16 * half, d);
        umma::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(&bar_dfree[b]);
        float* srow = stg + (size_t)b * kF * kStgStride + row * kStgStride + 16 * half;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float x = p.scale * d[e];
          // do_exp 2: the power spectrum as world_features_to_raw builds it (W:924): float32 amplitude, squared in float64
          const float a = p.do_exp ? expf(x) : x;
          srow[e] = p.do_exp == 2 ? (float)((double)a * (double)a) : a;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");  // the staging tile of this chunk is complete (and the other one is no longer read)
        // rows leave as 128-byte coalesced stores: warp w takes rows w, w + 8, ...; lane = bin within the chunk
        const int j = kBK * c + lane;
        if (j < p.K) {
          const float* sb = stg + (size_t)b * kF * kStgStride + lane;
          for (int r = ew; r < nvalid; r += kEpiWarps) p.out[(frame0 + r) * p.K + j] = sb[r * kStgStride];
        }
      }
    }
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

}  // namespace
}  // namespace b2w

// Tensor-core version of b2w_mc2sp for order <= 59 and a float32 output plane (the batched synthesis path).  stream1 = the pre-tiled
// Newton stream of b2w_mcep_tc_pretile for the same (order, alpha, fft_size).
extern "C" int b2w_mc2sp_tc(const void* mc, int32_t mc_dtype, int64_t mc_stride, int64_t num_frames, int32_t fft_size, int32_t order,
                            const float* stream1, double scale, int32_t do_exp, float* out, void* stream) {
  using namespace b2w;
  B2W_REQUIRE(mc && stream1 && out, "b2w_mc2sp_tc: null argument");
  B2W_REQUIRE(mc_dtype == B2W_F64 || mc_dtype == B2W_F32, "b2w_mc2sp_tc: bad mc_dtype %d", mc_dtype);
  B2W_REQUIRE(order >= 1 && order <= 59, "b2w_mc2sp_tc: order %d out of range [1, 59] (use b2w_mc2sp)", order);
  B2W_REQUIRE(fft_size >= 64 && (fft_size & (fft_size - 1)) == 0, "b2w_mc2sp_tc: bad fft_size %d", fft_size);
  B2W_REQUIRE(mc_stride >= order + 1 && do_exp >= 0 && do_exp <= 2, "b2w_mc2sp_tc: bad stride / do_exp");
  if (num_frames == 0) return 0;
  Params p;
  p.mc = mc; p.mc_is_f64 = mc_dtype == B2W_F64; p.mc_stride = mc_stride; p.num_frames = num_frames;
  p.K = fft_size / 2 + 1; p.m = order; p.nchunks = (p.K + kBK - 1) / kBK;
  p.scale = (float)scale; p.do_exp = do_exp; p.out = out; p.stream1 = stream1;
  const int64_t tiles = (num_frames + kF - 1) / kF;
  const int grid = (int)(tiles < 2 * 148 ? tiles : 2 * 148);
  cudaFuncSetAttribute(mc2sp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem::total);
  mc2sp_tc_kernel<<<grid, kThreads, Smem::total, (cudaStream_t)stream>>>(p);
  return check_launch("mc2sp_tc_kernel");
}
