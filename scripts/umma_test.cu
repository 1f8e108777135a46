// Self-test of the tcgen05 building blocks used by the tensor-core mel-cepstrum kernel: D[128 x N] = A[128 x K] . Bt[N x K]^T
// with the 3xTF32 split (A_hi B_hi + A_hi B_lo + A_lo B_hi, fp32 accumulation in tensor memory).  One CTA.
// The B operand arrives through a 1-D bulk async copy from a pre-tiled global buffer, exactly as in the production kernel.
#include "common.cuh"
#include "umma.cuh"

namespace b2w {

// Bt [N x K] row-major fp32 -> [hi tile | lo tile] in the K-major core-matrix order (what the kernels bulk-copy)
__global__ void umma_pretile_kernel(const float* __restrict__ bt, int N, int K, float* __restrict__ tiled) {
  const int total = N * K;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int n = e / K, k = e - n * K;
    float hi, lo;
    umma::split_tf32(bt[e], hi, lo);
    const uint32_t off = umma::tile_off(N, n, k) / 4;
    tiled[off] = hi;
    tiled[N * K + off] = lo;
  }
}

__global__ void __launch_bounds__(128) umma_test_kernel(const float* __restrict__ a, const float* __restrict__ b_tiled, int N, int K,
                                                        float* __restrict__ d) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* a_hi = reinterpret_cast<float*>(smem);
  float* a_lo = a_hi + 128 * K;
  float* b_hi = a_lo + 128 * K;
  float* b_lo = b_hi + N * K;
  uint64_t* bar_b = reinterpret_cast<uint64_t*>(b_lo + N * K);
  uint64_t* bar_mma = bar_b + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    umma::mbar_init(bar_b, 1);
    umma::mbar_init(bar_mma, 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, 256);
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    const uint32_t bytes = 2u * N * K * 4u;
    umma::mbar_expect_tx(bar_b, bytes);
    umma::bulk_g2s(b_hi, b_tiled, bytes, bar_b);
  }
  // stage A: split and scatter into the core-matrix layout
  for (int e = tid; e < 128 * K; e += 128) {
    const int r = e / K, k = e - r * K;
    float hi, lo;
    umma::split_tf32(a[e], hi, lo);
    const uint32_t off = umma::tile_off(128, r, k) / 4;
    a_hi[off] = hi;
    a_lo[off] = lo;
  }
  umma::fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    umma::mbar_wait(bar_b, 0);
    umma::tc_fence_after_sync();
    const uint32_t idesc = umma::idesc_tf32(128, N);
    const uint32_t a_lbo = 128 * 16, b_lbo = (uint32_t)N * 16, sbo = 128;
    bool acc = false;
    for (int s = 0; s < K / 8; ++s) {
      const uint64_t ah = umma::smem_desc(umma::smem_u32(a_hi) + 2 * s * a_lbo, a_lbo, sbo);
      const uint64_t al = umma::smem_desc(umma::smem_u32(a_lo) + 2 * s * a_lbo, a_lbo, sbo);
      const uint64_t bh = umma::smem_desc(umma::smem_u32(b_hi) + 2 * s * b_lbo, b_lbo, sbo);
      const uint64_t bl = umma::smem_desc(umma::smem_u32(b_lo) + 2 * s * b_lbo, b_lbo, sbo);
      umma::mma_tf32(tmem, al, bh, idesc, acc);   // small terms first
      umma::mma_tf32(tmem, ah, bl, idesc, true);
      umma::mma_tf32(tmem, ah, bh, idesc, true);
      acc = true;
    }
    umma::mma_commit(bar_mma);
  }
  umma::mbar_wait(bar_mma, 0);
  umma::tc_fence_after_sync();
  // epilogue: warp w owns TMEM lanes 32 w .. 32 w + 31 = rows of D
  const int row = 32 * warp + lane;
  for (int c = 0; c < N; c += 16) {
    float v[16];
    umma::tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) d[row * N + c + i] = v[i];
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

// Timing probe (scripts/gpu_umma_probe.py): cycles from the first tcgen05.mma to the commit's mbarrier arrival for `count`
// MMAs of shape 128 x N x 8 (kind::tf32) spread round-robin over `nacc` independent accumulators (zeroed operands).
__global__ void __launch_bounds__(128) umma_probe_kernel(int N, int count, int nacc, int M, int f16, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* a = reinterpret_cast<float*>(smem);             // 128 x 64
  float* b = a + 128 * 64;                                // 256 x 64 at most... only N x 64 used
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (128 * 64 + 128 * 64) * 4);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 2 * 128 * 64; i += 128) a[i] = 0.f;
  if (tid == 0) { umma::mbar_init(bar, 1); umma::mbar_fence_init(); }
  if (warp == 0) umma::tmem_alloc(slot, 512);
  umma::fence_proxy_async();
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = *slot;
  uint32_t phase = 0;
  for (int rep = 0; rep < 4; ++rep) {
    long long t0 = 0;
    if (tid == 0) {
      // kind::f16 probe: bf16 operands (format 1), fp32 accumulate, K = 16 per instruction (same 32 bytes per row)
      const uint32_t idesc = f16 ? ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24))
                                 : umma::idesc_tf32(M, N);
      const uint32_t a_lbo = (uint32_t)M * 16, b_lbo = (uint32_t)N * 16;
      const uint64_t ad0 = umma::smem_desc(umma::smem_u32(a), a_lbo, 128), bd0 = umma::smem_desc(umma::smem_u32(b), b_lbo, 128);
      t0 = clock64();
      if (f16) {
        for (int i = 0; i < count; ++i) {
          const int ks = i & 7, acc = i % nacc;
          const uint64_t ad = ad0 + ks * ((2 * a_lbo) >> 4), bd = bd0 + ks * ((2 * b_lbo) >> 4);
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem + acc * N),
                       "l"(ad), "l"(bd), "r"(idesc), "r"((uint32_t)(i >= nacc))
                       : "memory");
        }
      } else {
        // groups of 24 = one 3xTF32 product over K = 64 with unrolled, precomputed descriptors
        for (int i = 0; i < count; i += 24)
          umma::mma_3xtf32<8>(tmem + ((i / 24) % nacc) * N, ad0, ad0, bd0, bd0, 2 * a_lbo, 2 * b_lbo, idesc, i >= 24 * nacc);
      }
      umma::mma_commit(bar);
    }
    umma::mbar_wait(bar, phase);
    phase ^= 1;
    if (tid == 0) out[rep] = clock64() - t0;
    umma::tc_fence_after_sync();
    __syncthreads();
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

}  // namespace b2w

extern "C" int b2w_probe_umma(int32_t n, int32_t count, int32_t nacc, int32_t m, int32_t f16, long long* out4, void* stream) {
  using namespace b2w;
  B2W_REQUIRE(out4 && n >= 16 && n <= 128 && n % 16 == 0 && nacc >= 1 && nacc * n <= 512 && (m == 64 || m == 128), "b2w_probe_umma: bad arguments");
  const size_t smem = (size_t)(2 * 128 * 64) * 4 + 64;
  cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(n, count, nacc, m, f16, out4);
  return check_launch("umma_probe_kernel");
}

extern "C" int b2w_test_umma_gemm(const float* a, const float* bt, int32_t n, int32_t k, float* b_tiled_ws, float* d, void* stream) {
  using namespace b2w;
  B2W_REQUIRE(a && bt && b_tiled_ws && d, "b2w_test_umma_gemm: null argument");
  B2W_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0 && k >= 8 && k % 8 == 0, "b2w_test_umma_gemm: need N % 16 == 0 (<= 256), K % 8 == 0");
  const size_t smem = (size_t)(2 * 128 * k + 2 * n * k) * 4 + 64;
  B2W_REQUIRE(smem <= 227 * 1024, "b2w_test_umma_gemm: tile too large");
  cudaStream_t st = (cudaStream_t)stream;
  umma_pretile_kernel<<<64, 256, 0, st>>>(bt, n, k, b_tiled_ws);
  cudaFuncSetAttribute(umma_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  umma_test_kernel<<<1, 128, smem, st>>>(a, b_tiled_ws, n, k, d);
  return check_launch("umma_test_kernel");
}
