// Label preparation around the vocoder kernels: lf0 / vuv with Merlin-style interpolation, np.gradient deltas, and the
// corpus normalisation statistics.
//
// Replaces (reference paths under idiaptts/):
//   src/data_preparation/world/WorldFeatLabelGen.py:798-802  lf0 = log(f0.clip(1e-10)) as float32, threshold, interpolate_lin
//   misc/utils.py:40-86                                      interpolate_lin (pure-python double loop)
//   misc/utils.py:103-105                                    compute_deltas = np.gradient(...).astype(float32)
//   misc/normalisation/MeanStdDevExtractor.py:43-47          add_sample: N += T, sum x, sum x^2
//   misc/normalisation/MeanCovarianceExtractor.py:44-48      add_sample: sum x, sum x^T x
#include "common.cuh"

namespace b2w {

// One warp per utterance.  The reference's interpolation is a sequential scan; here the control flow of that scan stays
// sequential per utterance (gap by gap, warp-uniform), while everything inside it is spread over the lanes: the log / threshold
// pass, the searches for the next unvoiced / voiced frame (ballots over 32 frames at a time) and the fills.  Same float32
// operation order as the reference, no fused multiply-add.  (Round 1 ran one THREAD per utterance: 7 ms for any batch size,
// strided 4-byte accesses in a dependent chain; this form is ~100 x shorter.)
__device__ __forceinline__ int warp_find(const float* d, int64_t stride, int from, int n, bool want_voiced, int lane) {
  for (int base = from; base < n; base += 32) {  // first k >= from with (d[k] > 0) == want_voiced, or n
    const int k = base + lane;
    const bool hit = k < n && ((d[(int64_t)k * stride] > 0.f) == want_voiced);
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (m) return base + __ffs(m) - 1;
  }
  return n;
}

__global__ void __launch_bounds__(128) lf0_vuv_kernel(const double* __restrict__ f0, const int64_t* __restrict__ utt_frame_offset,
                                                      int num_utts, float log_thr, float lf0_zero, float* lf0, float* vuv,
                                                      int64_t stride) {
  const int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (u >= num_utts) return;  // warp-uniform
  const int64_t beg = utt_frame_offset[u];
  const int n = (int)(utt_frame_offset[u + 1] - beg);
  float* d = lf0 + beg * stride;
  float* v = vuv + beg * stride;
  const double* f = f0 + beg;
  // pass 1: thresholded log-F0 and the voicing flag (taken BEFORE interpolation, utils.py:50-52)
  for (int i = lane; i < n; i += 32) {
    const float x = (float)fmax(f[i], 1e-10);  // np.log(..., dtype=float32) casts its input to float32 first
    float l = (float)log((double)x);           // correctly rounded float32 logarithm
    if (l <= log_thr) l = lf0_zero;
    d[(int64_t)i * stride] = l;
    v[(int64_t)i * stride] = l > 0.f ? 1.f : 0.f;
  }
  __syncwarp();
  // pass 2: fill the unvoiced gaps [g, j)
  int i = 0;
  while (i < n) {
    const int g = warp_find(d, stride, i, n, false, lane);
    if (g >= n) break;
    // the value the reference carries as last_value: frame g - 1 is voiced (gaps are maximal), 0 before the first voiced frame
    const float last_value = g > 0 ? d[(int64_t)(g - 1) * stride] : 0.f;
    const int j = warp_find(d, stride, g + 1, n, true, lane);  // n when no voiced frame follows
    if (j < n - 1) {
      const float dj = d[(int64_t)j * stride];
      if (last_value > 0.f) {
        const float step = __fdiv_rn(__fsub_rn(dj, last_value), (float)(j - g));
        for (int k = g + lane; k < j; k += 32) d[(int64_t)k * stride] = __fadd_rn(last_value, __fmul_rn(step, (float)(k - g + 1)));
      } else {
        for (int k = g + lane; k < j; k += 32) d[(int64_t)k * stride] = dj;
      }
      i = j;
    } else {
      // "end of data": also taken when the next voiced frame is the LAST frame, which is then overwritten too
      for (int k = g + lane; k < n; k += 32) d[(int64_t)k * stride] = last_value;
      break;
    }
    __syncwarp();
  }
}

__device__ __forceinline__ int find_utt(const int64_t* __restrict__ off, int num_utts, int64_t frame) {
  int lo = 0, hi = num_utts;  // off[lo] <= frame < off[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= frame) lo = mid; else hi = mid;
  }
  return lo;
}

// np.gradient along time in float32: interior (x[i+1] - x[i-1]) / 2, one-sided differences at both ends
__device__ __forceinline__ float grad_at(const float* __restrict__ col, int64_t stride, int i, int n) {
  if (n < 2) return 0.f;
  if (i == 0) return __fsub_rn(col[stride], col[0]);
  if (i == n - 1) return __fsub_rn(col[(int64_t)(n - 1) * stride], col[(int64_t)(n - 2) * stride]);
  return __fmul_rn(__fsub_rn(col[(int64_t)(i + 1) * stride], col[(int64_t)(i - 1) * stride]), 0.5f);
}

__global__ void deltas_kernel(const float* __restrict__ feats, int64_t fstride, int dim,
                              const int64_t* __restrict__ utt_frame_offset, int num_utts, int64_t total_frames,
                              float* __restrict__ deltas, float* __restrict__ ddeltas, int64_t ostride) {
  const int64_t total = total_frames * dim;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t frame = e / dim;
    const int c = (int)(e - frame * dim);
    const int u = find_utt(utt_frame_offset, num_utts, frame);
    const int64_t beg = utt_frame_offset[u];
    const int n = (int)(utt_frame_offset[u + 1] - beg);
    const int i = (int)(frame - beg);
    const float* col = feats + beg * fstride + c;
    const float d = grad_at(col, fstride, i, n);
    if (deltas) deltas[frame * ostride + c] = d;
    if (ddeltas) {
      // gradient of the float32 deltas: recompute the neighbouring deltas (identical float32 values)
      float dd = 0.f;
      if (n >= 2) {
        if (i == 0) dd = __fsub_rn(grad_at(col, fstride, 1, n), d);
        else if (i == n - 1) dd = __fsub_rn(d, grad_at(col, fstride, n - 2, n));
        else dd = __fmul_rn(__fsub_rn(grad_at(col, fstride, i + 1, n), grad_at(col, fstride, i - 1, n)), 0.5f);
      }
      ddeltas[frame * ostride + c] = dd;
    }
  }
}

// Vectorised form of the same arithmetic for rows that can be read as float4 (dim, both strides multiples of 4, 16-byte aligned
// pointers): a thread owns four columns and kDeltaRows consecutive frames and slides a five-row window x[i-2 .. i+2] down the
// utterance, so a feature row is read once (plus a two-row halo per thread) instead of up to nine times, the utterance is looked
// up once per thread instead of once per element, and a warp moves whole 256-byte row segments.  Same float32 operations in the
// same order as grad_at: bit-identical results.
constexpr int kDeltaRows = 16;

__device__ __forceinline__ float grad5(const float (&x)[5], int k, int i, int n) {  // gradient at row i + k, x = rows i-2 .. i+2
  const int j = i + k;
  if (n < 2) return 0.f;
  if (j == 0) return __fsub_rn(x[k + 3], x[k + 2]);
  if (j == n - 1) return __fsub_rn(x[k + 2], x[k + 1]);
  return __fmul_rn(__fsub_rn(x[k + 3], x[k + 1]), 0.5f);
}

__device__ __forceinline__ void delta_pair(const float (&x)[5], int i, int n, float& d, float& dd) {
  d = grad5(x, 0, i, n);
  dd = 0.f;
  if (n >= 2) {
    if (i == 0) dd = __fsub_rn(grad5(x, 1, i, n), d);
    else if (i == n - 1) dd = __fsub_rn(d, grad5(x, -1, i, n));
    else dd = __fmul_rn(__fsub_rn(grad5(x, 1, i, n), grad5(x, -1, i, n)), 0.5f);
  }
}

__global__ void __launch_bounds__(256) deltas_rows_kernel(const float* __restrict__ feats, int64_t fstride, int dim4,
                                                          const int64_t* __restrict__ utt_frame_offset, int num_utts,
                                                          int64_t total_frames, float* __restrict__ deltas,
                                                          float* __restrict__ ddeltas, int64_t ostride) {
  const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int cv = (int)(gid % dim4);
  const int64_t r0 = (gid / dim4) * kDeltaRows;
  if (r0 >= total_frames) return;
  const int64_t r1 = min(total_frames, r0 + kDeltaRows);
  int u = find_utt(utt_frame_offset, num_utts, r0);
  int64_t beg = utt_frame_offset[u], end = utt_frame_offset[u + 1];
  int n = (int)(end - beg), i = (int)(r0 - beg);
  const float* base = feats + 4 * cv;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 w[5];
  auto row = [&](int idx) { return (idx >= 0 && idx < n) ? *reinterpret_cast<const float4*>(base + (beg + idx) * fstride) : zero; };
#pragma unroll
  for (int k = 0; k < 5; ++k) w[k] = row(i - 2 + k);
  for (int64_t frame = r0; frame < r1; ++frame) {
    if (frame == end) {  // next (non-empty) utterance: restart the window
      do {
        ++u;
        beg = utt_frame_offset[u];
        end = utt_frame_offset[u + 1];
      } while (end == beg);
      n = (int)(end - beg);
      i = 0;
#pragma unroll
      for (int k = 0; k < 5; ++k) w[k] = row(i - 2 + k);
    }
    const float4 nxt = row(i + 3);  // in flight while this row is finished
    float4 d, dd;
    {
      const float x0[5] = {w[0].x, w[1].x, w[2].x, w[3].x, w[4].x}, x1[5] = {w[0].y, w[1].y, w[2].y, w[3].y, w[4].y};
      const float x2[5] = {w[0].z, w[1].z, w[2].z, w[3].z, w[4].z}, x3[5] = {w[0].w, w[1].w, w[2].w, w[3].w, w[4].w};
      delta_pair(x0, i, n, d.x, dd.x);
      delta_pair(x1, i, n, d.y, dd.y);
      delta_pair(x2, i, n, d.z, dd.z);
      delta_pair(x3, i, n, d.w, dd.w);
    }
    if (deltas) *reinterpret_cast<float4*>(deltas + frame * ostride + 4 * cv) = d;
    if (ddeltas) *reinterpret_cast<float4*>(ddeltas + frame * ostride + 4 * cv) = dd;
    w[0] = w[1];
    w[1] = w[2];
    w[2] = w[3];
    w[3] = w[4];
    w[4] = nxt;
    ++i;
  }
}

// sums[c] += sum_t x[t][c]; sums[dim + c] += sum_t x[t][c]^2 (fp64).  blockDim.x = columns handled per pass (<= 256),
// blockDim.y row lanes; one fp64 atomic per column and block.
__global__ void stats_kernel(const float* __restrict__ feats, int64_t fstride, int dim, int64_t num_frames,
                             int64_t rows_per_block, double* __restrict__ sums) {
  extern __shared__ double sh[];  // [2][blockDim.y][blockDim.x]
  const int64_t r0 = blockIdx.x * rows_per_block;
  const int64_t r1 = min(num_frames, r0 + rows_per_block);
  const int nx = blockDim.x, ny = blockDim.y;
  for (int c0 = 0; c0 < dim; c0 += nx) {
    const int c = c0 + threadIdx.x;
    double s = 0.0, q = 0.0;
    if (c < dim) {
      for (int64_t r = r0 + threadIdx.y; r < r1; r += ny) {
        const double x = (double)feats[r * fstride + c];
        s += x;
        q += x * x;
      }
    }
    sh[(0 * ny + threadIdx.y) * nx + threadIdx.x] = s;
    sh[(1 * ny + threadIdx.y) * nx + threadIdx.x] = q;
    __syncthreads();
    if (threadIdx.y == 0 && c < dim) {
      double ts = 0.0, tq = 0.0;
      for (int y = 0; y < ny; ++y) {
        ts += sh[(0 * ny + y) * nx + threadIdx.x];
        tq += sh[(1 * ny + y) * nx + threadIdx.x];
      }
      atomicAdd(&sums[c], ts);
      atomicAdd(&sums[dim + c], tq);
    }
    __syncthreads();
  }
}

// gram[a][b] += sum_t x[t][a] x[t][b] (fp64).  blockIdx.y picks a 16 x 16 tile of (a, b) pairs (one pair per thread),
// blockIdx.x a slab of rows; rows are staged 32 at a time in shared memory.  Only the add_deltas /
// MeanCovarianceExtractor recipe needs it.
constexpr int kGramRows = 32;
__global__ void gram_kernel(const float* __restrict__ feats, int64_t fstride, int dim, int64_t num_frames,
                            int64_t rows_per_block, double* __restrict__ gram) {
  __shared__ float sa[kGramRows][17], sb[kGramRows][17];
  const int tiles = (dim + 15) / 16;
  const int a0 = (blockIdx.y / tiles) * 16, b0 = (blockIdx.y % tiles) * 16;
  const int ta = threadIdx.x >> 4, tb = threadIdx.x & 15;
  const int64_t r0 = blockIdx.x * rows_per_block;
  const int64_t r1 = min(num_frames, r0 + rows_per_block);
  double acc = 0.0;
  for (int64_t rs = r0; rs < r1; rs += kGramRows) {
    const int nr = (int)min((int64_t)kGramRows, r1 - rs);
    __syncthreads();
    for (int e = threadIdx.x; e < kGramRows * 16; e += blockDim.x) {
      const int r = e >> 4, c = e & 15;
      float va = 0.f, vb = 0.f;
      if (r < nr) {
        if (a0 + c < dim) va = feats[(rs + r) * fstride + a0 + c];
        if (b0 + c < dim) vb = feats[(rs + r) * fstride + b0 + c];
      }
      sa[r][c] = va;
      sb[r][c] = vb;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < kGramRows; ++r) acc += (double)sa[r][ta] * (double)sb[r][tb];
  }
  if (a0 + ta < dim && b0 + tb < dim) atomicAdd(&gram[(int64_t)(a0 + ta) * dim + b0 + tb], acc);
}

}  // namespace b2w

// ---- trainer-facing batch ---------------------------------------------------------------------------------------
// Ragged feature rows -> the padded, normalised tensor a trainer consumes, in one pass: WorldFeatLabelGen.preprocess_sample
// ((x - mean) / std_dev in float32, world/WorldFeatLabelGen.py:279-336) followed by ModularModelHandlerPyTorch.prepare_batch
// (pad_sequence with zeros to the longest utterance + sequence_mask, :389-499).  One warp per (t, b) row, lanes over the
// feature columns: coalesced reads and writes, the kernel is HBM-bound.  Separate subtract / divide (and multiply / add in the
// inverse): bit-identical to numpy's float32 arithmetic.
namespace b2w {
__global__ void __launch_bounds__(256) pad_normalise_kernel(const float* __restrict__ feats, int64_t feat_stride, int width,
                                                            const int64_t* __restrict__ foff, int U, int t_max,
                                                            const float* __restrict__ mean, const float* __restrict__ std_dev,
                                                            int batch_first, float* __restrict__ out, float* __restrict__ mask) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t rows = (int64_t)t_max * U;
  for (int64_t r = warp0; r < rows; r += nwarps) {
    // r enumerates the OUTPUT rows in memory order
    const int b = batch_first ? (int)(r / t_max) : (int)(r % U);
    const int t = batch_first ? (int)(r % t_max) : (int)(r / U);
    const int64_t f0 = foff[b];
    const bool valid = t < (int)(foff[b + 1] - f0);
    float* o = out + r * width;
    if (valid) {
      const float* x = feats + (f0 + t) * feat_stride;
      for (int w = lane; w < width; w += 32) {
        const float m = mean ? mean[w] : 0.f, sd = std_dev ? std_dev[w] : 1.f;
        o[w] = __fdiv_rn(__fsub_rn(x[w], m), sd);
      }
    } else {
      for (int w = lane; w < width; w += 32) o[w] = 0.f;
    }
    if (mask && lane == 0) mask[r] = valid ? 1.f : 0.f;
  }
}

// float4 variants (width, strides and pointers multiples of 4 floats): one thread per 16-byte chunk, two chunks in flight
__device__ __forceinline__ float4 norm4(float4 x, float4 m, float4 sd) {
  return make_float4(__fdiv_rn(__fsub_rn(x.x, m.x), sd.x), __fdiv_rn(__fsub_rn(x.y, m.y), sd.y), __fdiv_rn(__fsub_rn(x.z, m.z), sd.z),
                     __fdiv_rn(__fsub_rn(x.w, m.w), sd.w));
}
__device__ __forceinline__ float4 denorm4(float4 x, float4 m, float4 sd) {
  return make_float4(__fadd_rn(__fmul_rn(x.x, sd.x), m.x), __fadd_rn(__fmul_rn(x.y, sd.y), m.y), __fadd_rn(__fmul_rn(x.z, sd.z), m.z),
                     __fadd_rn(__fmul_rn(x.w, sd.w), m.w));
}

__global__ void __launch_bounds__(256) pad_normalise4_kernel(const float4* __restrict__ feats, int64_t stride4, int w4,
                                                             const int64_t* __restrict__ foff, int U, int t_max,
                                                             const float4* __restrict__ mean, const float4* __restrict__ std_dev,
                                                             int batch_first, float4* __restrict__ out, float* __restrict__ mask) {
  const int64_t total = (int64_t)t_max * U * w4;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f), one = make_float4(1.f, 1.f, 1.f, 1.f);
  for (int64_t q0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q0 < total; q0 += 2 * step) {
    float4 x[2];
    int c[2];
    bool valid[2], live[2];
    int64_t r[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int64_t q = q0 + k * step;
      live[k] = q < total;
      valid[k] = false;
      x[k] = zero;
      c[k] = 0;
      r[k] = 0;
      if (live[k]) {
        r[k] = q / w4;
        c[k] = (int)(q - r[k] * w4);
        const int b = batch_first ? (int)(r[k] / t_max) : (int)(r[k] % U);
        const int t = batch_first ? (int)(r[k] % t_max) : (int)(r[k] / U);
        const int64_t f0 = foff[b];
        valid[k] = t < (int)(foff[b + 1] - f0);
        if (valid[k]) x[k] = __ldcs(feats + (f0 + t) * stride4 + c[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (!live[k]) continue;
      float4 v = zero;
      if (valid[k]) v = norm4(x[k], mean ? mean[c[k]] : zero, std_dev ? std_dev[c[k]] : one);
      __stcs(out + r[k] * w4 + c[k], v);
      if (mask && c[k] == 0) mask[r[k]] = valid[k] ? 1.f : 0.f;
    }
  }
}

__global__ void __launch_bounds__(256) unpad_denormalise4_kernel(const float4* __restrict__ padded, int w4,
                                                                 const int64_t* __restrict__ foff, const int32_t* __restrict__ frame_utt,
                                                                 int64_t F, int U, int t_max, const float4* __restrict__ mean,
                                                                 const float4* __restrict__ std_dev, int batch_first,
                                                                 float4* __restrict__ feats, int64_t stride4) {
  const int64_t total = F * w4;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f), one = make_float4(1.f, 1.f, 1.f, 1.f);
  for (int64_t q0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q0 < total; q0 += 2 * step) {
    float4 x[2];
    int c[2];
    int64_t f[2];
    bool live[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int64_t q = q0 + k * step;
      live[k] = q < total;
      x[k] = zero;
      c[k] = 0;
      f[k] = 0;
      if (live[k]) {
        f[k] = q / w4;
        c[k] = (int)(q - f[k] * w4);
        const int b = frame_utt[f[k]];
        const int t = (int)(f[k] - foff[b]);
        const int64_t r = batch_first ? (int64_t)b * t_max + t : (int64_t)t * U + b;
        x[k] = __ldcs(padded + r * w4 + c[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k)
      if (live[k]) __stcs(feats + f[k] * stride4 + c[k], denorm4(x[k], mean ? mean[c[k]] : zero, std_dev ? std_dev[c[k]] : one));
  }
}

__global__ void __launch_bounds__(256) unpad_denormalise_kernel(const float* __restrict__ padded, int width,
                                                                const int64_t* __restrict__ foff, const int32_t* __restrict__ frame_utt,
                                                                int64_t F, int U, int t_max, const float* __restrict__ mean,
                                                                const float* __restrict__ std_dev, int batch_first,
                                                                float* __restrict__ feats, int64_t feat_stride) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t f = warp0; f < F; f += nwarps) {
    const int b = frame_utt[f];
    const int t = (int)(f - foff[b]);
    const int64_t r = batch_first ? (int64_t)b * t_max + t : (int64_t)t * U + b;
    const float* x = padded + r * width;
    float* o = feats + f * feat_stride;
    for (int w = lane; w < width; w += 32) {
      const float m = mean ? mean[w] : 0.f, sd = std_dev ? std_dev[w] : 1.f;
      o[w] = __fadd_rn(__fmul_rn(x[w], sd), m);
    }
  }
}
}  // namespace b2w

extern "C" int b2w_pad_normalise(const float* feats, int64_t feat_stride, int32_t width, const int64_t* utt_frame_offset,
                                 int32_t num_utts, int32_t t_max, const float* mean, const float* std_dev, int32_t batch_first,
                                 float* out, float* mask, void* stream) {
  using namespace b2w;
  B2W_REQUIRE(feats && utt_frame_offset && out, "b2w_pad_normalise: null argument");
  B2W_REQUIRE(width >= 1 && feat_stride >= width && t_max >= 0, "b2w_pad_normalise: bad width %d / stride / t_max", width);
  const int64_t rows = (int64_t)t_max * num_utts;
  if (num_utts <= 0 || rows == 0) return 0;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (width % 4 == 0 && feat_stride % 4 == 0 && al16(feats) && al16(out) && al16(mean) && al16(std_dev)) {
    const int64_t quads = rows * (width / 4);
    int64_t g4 = (quads + 511) / 512;
    if (g4 > 148 * 16) g4 = 148 * 16;
    pad_normalise4_kernel<<<(unsigned)g4, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(feats), feat_stride / 4, width / 4, utt_frame_offset, num_utts, t_max,
        reinterpret_cast<const float4*>(mean), reinterpret_cast<const float4*>(std_dev), batch_first, reinterpret_cast<float4*>(out), mask);
    return check_launch("pad_normalise4_kernel");
  }
  int64_t g = (rows + 7) / 8;
  if (g > 148 * 32) g = 148 * 32;
  pad_normalise_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(feats, feat_stride, width, utt_frame_offset, num_utts, t_max,
                                                                     mean, std_dev, batch_first, out, mask);
  return check_launch("pad_normalise_kernel");
}

extern "C" int b2w_unpad_denormalise(const float* padded, int32_t width, const int64_t* utt_frame_offset, const int32_t* frame_utt,
                                     int64_t num_frames, int32_t num_utts, int32_t t_max, const float* mean, const float* std_dev,
                                     int32_t batch_first, float* feats, int64_t feat_stride, void* stream) {
  using namespace b2w;
  B2W_REQUIRE(padded && utt_frame_offset && frame_utt && feats, "b2w_unpad_denormalise: null argument");
  B2W_REQUIRE(width >= 1 && feat_stride >= width && t_max >= 0, "b2w_unpad_denormalise: bad width %d / stride / t_max", width);
  if (num_utts <= 0 || num_frames <= 0) return 0;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (width % 4 == 0 && feat_stride % 4 == 0 && al16(feats) && al16(padded) && al16(mean) && al16(std_dev)) {
    const int64_t quads = num_frames * (width / 4);
    int64_t g4 = (quads + 511) / 512;
    if (g4 > 148 * 16) g4 = 148 * 16;
    unpad_denormalise4_kernel<<<(unsigned)g4, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(padded), width / 4, utt_frame_offset, frame_utt, num_frames, num_utts, t_max,
        reinterpret_cast<const float4*>(mean), reinterpret_cast<const float4*>(std_dev), batch_first, reinterpret_cast<float4*>(feats),
        feat_stride / 4);
    return check_launch("unpad_denormalise4_kernel");
  }
  int64_t g = (num_frames + 7) / 8;
  if (g > 148 * 32) g = 148 * 32;
  unpad_denormalise_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(padded, width, utt_frame_offset, frame_utt, num_frames,
                                                                         num_utts, t_max, mean, std_dev, batch_first, feats,
                                                                         feat_stride);
  return check_launch("unpad_denormalise_kernel");
}

extern "C" int b2w_lf0_vuv(const double* f0, const int64_t* utt_frame_offset, int32_t num_utts, double f0_silence_threshold,
                           double lf0_zero, float* lf0, float* vuv, int64_t out_stride, void* stream) {
  using namespace b2w;
  B2W_REQUIRE(f0 && utt_frame_offset && lf0 && vuv, "b2w_lf0_vuv: null argument");
  B2W_REQUIRE(out_stride >= 1, "b2w_lf0_vuv: bad stride");
  if (num_utts == 0) return 0;
  const float log_thr = (float)log(f0_silence_threshold);
  lf0_vuv_kernel<<<(num_utts + 3) / 4, 128, 0, (cudaStream_t)stream>>>(f0, utt_frame_offset, num_utts, log_thr,
                                                                        (float)lf0_zero, lf0, vuv, out_stride);
  return check_launch("lf0_vuv_kernel");
}

extern "C" int b2w_deltas(const float* feats, int64_t feat_stride, int32_t dim, const int64_t* utt_frame_offset,
                          int32_t num_utts, int64_t num_frames, float* deltas, float* ddeltas, int64_t out_stride,
                          void* stream) {
  using namespace b2w;
  B2W_REQUIRE(feats && utt_frame_offset && (deltas || ddeltas), "b2w_deltas: null argument");
  B2W_REQUIRE(dim >= 1 && feat_stride >= dim && out_stride >= dim, "b2w_deltas: bad dim/stride");
  if (num_utts == 0 || num_frames == 0) return 0;
  const auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (dim % 4 == 0 && feat_stride % 4 == 0 && out_stride % 4 == 0 && al16(feats) && al16(deltas) && al16(ddeltas)) {
    const int dim4 = dim / 4;
    const int64_t threads = ((num_frames + kDeltaRows - 1) / kDeltaRows) * dim4;
    deltas_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        feats, feat_stride, dim4, utt_frame_offset, num_utts, num_frames, deltas, ddeltas, out_stride);
    return check_launch("deltas_rows_kernel");
  }
  const int64_t total = num_frames * dim;
  int64_t g = (total + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  deltas_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(feats, feat_stride, dim, utt_frame_offset, num_utts, num_frames,
                                                               deltas, ddeltas, out_stride);
  return check_launch("deltas_kernel");
}

extern "C" int b2w_stats_accumulate(const float* feats, int64_t feat_stride, int32_t dim, int64_t num_frames, double* sums,
                                    double* gram, void* stream) {
  using namespace b2w;
  B2W_REQUIRE(feats && sums, "b2w_stats_accumulate: null argument");
  B2W_REQUIRE(dim >= 1 && feat_stride >= dim, "b2w_stats_accumulate: bad dim/stride");
  if (num_frames == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  {
    int nx = ((dim + 31) / 32) * 32;
    if (nx > 256) nx = 256;
    const int ny = 256 / nx > 0 ? 256 / nx : 1;
    int64_t nblk = 148 * 4;
    int64_t rows = (num_frames + nblk - 1) / nblk;
    if (rows < 64) rows = 64;
    nblk = (num_frames + rows - 1) / rows;
    stats_kernel<<<(unsigned)nblk, dim3(nx, ny), sizeof(double) * 2 * nx * ny, st>>>(feats, feat_stride, dim, num_frames, rows,
                                                                                      sums);
    int rc = check_launch("stats_kernel");
    if (rc) return rc;
  }
  if (gram) {
    const int tiles = (dim + 15) / 16;
    int64_t nblk = 148;
    int64_t rows = (num_frames + nblk - 1) / nblk;
    if (rows < kGramRows) rows = kGramRows;
    nblk = (num_frames + rows - 1) / rows;
    gram_kernel<<<dim3((unsigned)nblk, tiles * tiles), 256, 0, st>>>(feats, feat_stride, dim, num_frames, rows, gram);
    return check_launch("gram_kernel");
  }
  return 0;
}
