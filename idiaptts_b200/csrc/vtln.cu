// Neural-VTLN all-pass warp of mel-cepstra, forward and backward, without ever materialising a warp matrix.
//
// Replaces AllPassWarp.forward (idiaptts/src/neural_networks/pytorch/layers/AllPassWarp.py:148-173): the reference builds
// a [T*B, n, n] matrix per call from a polynomial tensor (einsum, :186-205, overflows in float32 for n >= ~35) and applies
// it with bmm.  That matrix is exactly the transpose of SPTK's freqt matrix, W(alpha) = A(alpha)^T, and
// A(alpha)^T == frqtr-matrix(-alpha).  So per row and n-block
//     forward :  y  = S2 . freqt(S1 x, alpha)                    S1 = diag(1/2, 1, ..), S2 = diag(2, 1, ..)  (:162-171)
//     backward:  gx = S1 . frqtr(S2 gy, -alpha),   galpha = < S2 gy , d freqt(S1 x, alpha) / d alpha >
// each an O(n^2) recursion held entirely in one thread's registers.  The optional mean / std_dev fold
// AllPassWarpLayer._denormalise / _normalise (layers/AllPassWarpLayer.py:186-200) into the same pass.
#include "common.cuh"

namespace b2w {

constexpr int kVtThreads = 128;

// stage `units` consecutive (row, block) vectors of n floats between global memory (contiguous) and shared [unit][n+1]
template <bool LOAD>
__device__ __forceinline__ void stage(float* __restrict__ sh, float* gl, int64_t first_elem, int valid_elems, int n) {
  for (int e = threadIdx.x; e < valid_elems; e += kVtThreads) {
    const int un = e / n, r = e - un * n;
    if (LOAD) sh[un * (n + 1) + r] = gl[first_elem + e];
    else gl[first_elem + e] = sh[un * (n + 1) + r];
  }
}

template <int NMAX>
__global__ void __launch_bounds__(kVtThreads)
allpass_forward_kernel(const float* __restrict__ x, const float* __restrict__ alpha, int64_t rows, int n, int blocks,
                       const float* __restrict__ mean, const float* __restrict__ std_dev, float* __restrict__ y,
                       const uint8_t* __restrict__ tile_mask) {
  extern __shared__ float sh[];  // [kVtThreads][n+1]
  const int64_t units = rows * blocks;
  // With a tile mask (second launch behind the tensor-core kernel of vtln_tc.cu, which has done every tile whose byte is zero) a
  // small grid walks the mask; without one every CTA owns the tile of its index.
  const int64_t num_tiles = (units + kVtThreads - 1) / kVtThreads;
  if (tile_mask) {  // one parallel look at this CTA's share of the mask: usually nothing is flagged and the CTA is done
    bool any = false;
    for (int64_t tile = blockIdx.x + (int64_t)threadIdx.x * gridDim.x; tile < num_tiles; tile += (int64_t)kVtThreads * gridDim.x)
      any = any || tile_mask[tile] != 0;
    if (!__syncthreads_or(any)) return;
  }
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
  if (tile_mask && !tile_mask[tile]) continue;
  const int64_t u0 = tile * kVtThreads;
  const int nun = (int)min((int64_t)kVtThreads, units - u0);
  stage<true>(sh, const_cast<float*>(x), u0 * n, nun * n, n);
  __syncthreads();
  const int tid = threadIdx.x;
  if (tid < nun) {
    const int64_t unit = u0 + tid;
    const int64_t row = unit / blocks;
    const int blk = (int)(unit - row * blocks);
    const float a = alpha[row];
    const float b = 1.f - a * a;
    float* v = sh + tid * (n + 1);
    float g[NMAX];
#pragma unroll
    for (int j = 0; j < NMAX; ++j) g[j] = 0.f;
    for (int r = n - 1; r >= 0; --r) {
      float xin = v[r];
      if (std_dev) xin *= std_dev[blk * n + r];
      if (mean) xin += mean[blk * n + r];
      if (r == 0) xin *= 0.5f;
      const float old0 = g[0];
      g[0] = fmaf(a, old0, xin);
      float prev_old = g[1];
      g[1] = fmaf(b, old0, a * prev_old);
      float prev_new = g[1];
#pragma unroll
      for (int j = 2; j < NMAX; ++j) {
        const float old = g[j];
        g[j] = fmaf(a, old - prev_new, prev_old);
        prev_old = old;
        prev_new = g[j];
      }
    }
    g[0] *= 2.f;
#pragma unroll
    for (int j = 0; j < NMAX; ++j) {
      if (j < n) {
        float o = g[j];
        if (mean) o -= mean[blk * n + j];
        if (std_dev) o /= std_dev[blk * n + j];
        v[j] = o;
      }
    }
  }
  __syncthreads();
  stage<false>(sh, y, u0 * n, nun * n, n);
  __syncthreads();
  }
}

template <int NMAX>
__global__ void __launch_bounds__(kVtThreads)
allpass_backward_kernel(const float* __restrict__ gy, const float* __restrict__ x, const float* __restrict__ alpha,
                        int64_t rows, int n, int blocks, const float* __restrict__ mean, const float* __restrict__ std_dev,
                        float* __restrict__ gx, float* __restrict__ galpha_unit, const uint8_t* __restrict__ tile_mask) {
  extern __shared__ float sh[];  // [2][kVtThreads][n+1]
  float* shx = sh;
  float* shg = sh + kVtThreads * (n + 1);
  const int64_t units = rows * blocks;
  const int64_t num_tiles = (units + kVtThreads - 1) / kVtThreads;
  if (tile_mask) {  // (see allpass_forward_kernel)
    bool any = false;
    for (int64_t tile = blockIdx.x + (int64_t)threadIdx.x * gridDim.x; tile < num_tiles; tile += (int64_t)kVtThreads * gridDim.x)
      any = any || tile_mask[tile] != 0;
    if (!__syncthreads_or(any)) return;
  }
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
  if (tile_mask && !tile_mask[tile]) continue;
  const int64_t u0 = tile * kVtThreads;
  const int nun = (int)min((int64_t)kVtThreads, units - u0);
  stage<true>(shx, const_cast<float*>(x), u0 * n, nun * n, n);
  stage<true>(shg, const_cast<float*>(gy), u0 * n, nun * n, n);
  __syncthreads();
  const int tid = threadIdx.x;
  if (tid < nun) {
    const int64_t unit = u0 + tid;
    const int64_t row = unit / blocks;
    const int blk = (int)(unit - row * blocks);
    const float a = alpha[row];
    const float b = 1.f - a * a;
    float* vx = shx + tid * (n + 1);
    float* vg = shg + tid * (n + 1);
    // upstream gradient w.r.t. the un-normalised output, with the c0 doubling folded in: gyp = S2 (gy / std)
    // ---- galpha: tangent of the forward recursion --------------------------------------------------------------------
    float ga = 0.f;
    {
      float g[NMAX], t[NMAX];
#pragma unroll
      for (int j = 0; j < NMAX; ++j) { g[j] = 0.f; t[j] = 0.f; }
      for (int r = n - 1; r >= 0; --r) {
        float xin = vx[r];
        if (std_dev) xin *= std_dev[blk * n + r];
        if (mean) xin += mean[blk * n + r];
        if (r == 0) xin *= 0.5f;
        const float old0 = g[0], told0 = t[0];
        g[0] = fmaf(a, old0, xin);
        t[0] = fmaf(a, told0, old0);
        float prev_old = g[1], tprev_old = t[1];
        g[1] = fmaf(b, old0, a * prev_old);
        t[1] = fmaf(-2.f * a, old0, fmaf(b, told0, fmaf(a, tprev_old, prev_old)));
        float prev_new = g[1], tprev_new = t[1];
#pragma unroll
        for (int j = 2; j < NMAX; ++j) {
          const float old = g[j], told = t[j];
          const float diff = old - prev_new;
          g[j] = fmaf(a, diff, prev_old);
          t[j] = tprev_old + diff + a * (told - tprev_new);
          prev_old = old; tprev_old = told;
          prev_new = g[j]; tprev_new = t[j];
        }
      }
#pragma unroll
      for (int j = 0; j < NMAX; ++j) {
        if (j < n) {
          float gyj = vg[j];
          if (std_dev) gyj /= std_dev[blk * n + j];
          if (j == 0) gyj *= 2.f;
          ga = fmaf(gyj, t[j], ga);
        }
      }
    }
    // ---- gx = S1 frqtr(S2 gy, -alpha) ---------------------------------------------------------------------------------------
    {
      const float na = -a;
      float g[NMAX];
#pragma unroll
      for (int j = 0; j < NMAX; ++j) g[j] = 0.f;
      for (int r = n - 1; r >= 0; --r) {
        float gin = vg[r];
        if (std_dev) gin /= std_dev[blk * n + r];
        if (r == 0) gin *= 2.f;
        float prev_old = g[0];
        g[0] = gin;
        float prev_new = gin;
#pragma unroll
        for (int j = 1; j < NMAX; ++j) {
          const float old = g[j];
          g[j] = fmaf(na, old - prev_new, prev_old);
          prev_old = old;
          prev_new = g[j];
        }
      }
      g[0] *= 0.5f;
#pragma unroll
      for (int j = 0; j < NMAX; ++j) {
        if (j < n) {
          float o = g[j];
          if (std_dev) o *= std_dev[blk * n + j];
          vx[j] = o;
        }
      }
    }
    galpha_unit[unit] = ga;
  }
  __syncthreads();
  stage<false>(shx, gx, u0 * n, nun * n, n);
  __syncthreads();
  }
}

// galpha[row] = sum over the row's blocks of the per-unit contributions
__global__ void reduce_blocks_kernel(const float* __restrict__ unit_vals, int64_t rows, int blocks, float* __restrict__ out) {
  const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (row >= rows) return;
  float s = 0.f;
  for (int b = 0; b < blocks; ++b) s += unit_vals[row * blocks + b];
  out[row] = s;
}

}  // namespace b2w

extern "C" int b2w_allpass_forward_masked(const float* x, const float* alpha, int64_t rows, int32_t n, int32_t blocks,
                                          const float* mean, const float* std_dev, float* y, const uint8_t* tile_mask, void* stream);
extern "C" int b2w_allpass_forward(const float* x, const float* alpha, int64_t rows, int32_t n, int32_t blocks,
                                   const float* mean, const float* std_dev, float* y, void* stream) {
  return b2w_allpass_forward_masked(x, alpha, rows, n, blocks, mean, std_dev, y, nullptr, stream);
}
// tile_mask (may be NULL): one byte per tile of 128 (row, block) units; only tiles with a non-zero byte are computed
extern "C" int b2w_allpass_forward_masked(const float* x, const float* alpha, int64_t rows, int32_t n, int32_t blocks,
                                          const float* mean, const float* std_dev, float* y, const uint8_t* tile_mask, void* stream) {
  using namespace b2w;
  B2W_REQUIRE(x && alpha && y, "b2w_allpass_forward: null argument");
  B2W_REQUIRE(n >= 2 && n <= 128 && blocks >= 1, "b2w_allpass_forward: n %d (2..128) / blocks %d out of range", n, blocks);
  if (rows == 0) return 0;
  const int64_t units = rows * blocks;
  int64_t grid = (units + kVtThreads - 1) / kVtThreads;
  B2W_REQUIRE(grid < ((int64_t)1 << 31), "b2w_allpass_forward: too many rows");
  if (tile_mask && grid > 2 * 148) grid = 2 * 148;  // flagged tiles are rare: a resident grid that scans the mask
  const size_t smem = sizeof(float) * kVtThreads * (n + 1);
  cudaStream_t st = (cudaStream_t)stream;
#define B2W_VT_FWD(NMAX)                                                                                   \
  do {                                                                                                     \
    cudaFuncSetAttribute(allpass_forward_kernel<NMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    allpass_forward_kernel<NMAX><<<(unsigned)grid, kVtThreads, smem, st>>>(x, alpha, rows, n, blocks, mean, std_dev, y, tile_mask); \
  } while (0)
  if (n <= 32) B2W_VT_FWD(32); else if (n <= 64) B2W_VT_FWD(64); else B2W_VT_FWD(128);
#undef B2W_VT_FWD
  return check_launch("allpass_forward_kernel");
}

extern "C" int b2w_allpass_backward_masked(const float* grad_y, const float* x, const float* alpha, int64_t rows, int32_t n, int32_t blocks,
                                           const float* mean, const float* std_dev, float* grad_x, float* grad_alpha, float* unit_workspace,
                                           const uint8_t* tile_mask, void* stream);
extern "C" int b2w_allpass_backward(const float* grad_y, const float* x, const float* alpha, int64_t rows, int32_t n,
                                    int32_t blocks, const float* mean, const float* std_dev, float* grad_x, float* grad_alpha,
                                    float* unit_workspace, void* stream) {
  return b2w_allpass_backward_masked(grad_y, x, alpha, rows, n, blocks, mean, std_dev, grad_x, grad_alpha, unit_workspace, nullptr, stream);
}
extern "C" int b2w_allpass_backward_masked(const float* grad_y, const float* x, const float* alpha, int64_t rows, int32_t n, int32_t blocks,
                                           const float* mean, const float* std_dev, float* grad_x, float* grad_alpha, float* unit_workspace,
                                           const uint8_t* tile_mask, void* stream) {
  using namespace b2w;
  B2W_REQUIRE(grad_y && x && alpha && grad_x && grad_alpha && unit_workspace, "b2w_allpass_backward: null argument");
  B2W_REQUIRE(n >= 2 && n <= 128 && blocks >= 1, "b2w_allpass_backward: n %d (2..128) / blocks %d out of range", n, blocks);
  if (rows == 0) return 0;
  const int64_t units = rows * blocks;
  int64_t grid = (units + kVtThreads - 1) / kVtThreads;
  B2W_REQUIRE(grid < ((int64_t)1 << 31), "b2w_allpass_backward: too many rows");
  if (tile_mask && grid > 2 * 148) grid = 2 * 148;  // flagged tiles are rare: a resident grid that scans the mask
  const size_t smem = sizeof(float) * 2 * kVtThreads * (n + 1);
  cudaStream_t st = (cudaStream_t)stream;
#define B2W_VT_BWD(NMAX)                                                                                    \
  do {                                                                                                      \
    cudaFuncSetAttribute(allpass_backward_kernel<NMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    allpass_backward_kernel<NMAX><<<(unsigned)grid, kVtThreads, smem, st>>>(grad_y, x, alpha, rows, n, blocks, mean, std_dev, \
                                                                            grad_x, blocks == 1 ? grad_alpha : unit_workspace, tile_mask); \
  } while (0)
  if (n <= 32) B2W_VT_BWD(32); else if (n <= 64) B2W_VT_BWD(64); else B2W_VT_BWD(128);
#undef B2W_VT_BWD
  int rc = check_launch("allpass_backward_kernel");
  if (rc) return rc;
  if (blocks == 1) return 0;  // a unit is a row: the per-unit values went straight to grad_alpha
  reduce_blocks_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(unit_workspace, rows, blocks, grad_alpha);
  return check_launch("reduce_blocks_kernel");
}
