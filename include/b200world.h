/*
 * b200world.h - C ABI of libb200world.so: the WORLD vocoder feature hot path of IdiapTTS as hand-written
 * CUDA for sm_100a (B200).
 *
 * This is the drop-in boundary. Today the reference crosses into native code at the pyworld / pysptk Cython
 * wrappers (single-threaded CPU C/C++, one utterance or one frame per call); every entry point below names the
 * reference call site it replaces (paths relative to the reference repo root):
 *
 *   W  = idiaptts/src/data_preparation/world/WorldFeatLabelGen.py
 *   A  = idiaptts/src/data_preparation/audio/AudioProcessing.py
 *   U  = idiaptts/misc/utils.py
 *   N  = idiaptts/misc/normalisation/MeanStdDevExtractor.py   (NC = MeanCovarianceExtractor.py)
 *   L  = idiaptts/src/neural_networks/pytorch/layers/AllPassWarp.py   (LL = AllPassWarpLayer.py)
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer unless its name starts with h_ (host). The caller (PyTorch) owns all
 *     memory; the library never allocates, frees or synchronises (exceptions are documented per call).
 *   - Ragged batches: utterance u owns samples [utt_sample_offset[u], utt_sample_offset[u+1]) of the packed
 *     waveform and frames [utt_frame_offset[u], utt_frame_offset[u+1]) of every per-frame array.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream). Calls are asynchronous.
 *   - Return value: 0 ok; < 0 argument error (nothing launched); > 0 cudaError_t of the launch.
 *     b2w_last_error() returns a thread-local message for the last non-zero return.
 *   - `status` arrays (int32, device) receive per-call data-dependent error flags (see B2W_STATUS_*); the
 *     caller reads them when it synchronises anyway.
 *   - Re-entrant per (device, stream). No global mutable state.
 */
#ifndef B200WORLD_H_
#define B200WORLD_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2W_VERSION 1

#if defined(__GNUC__)
#define B2W_API __attribute__((visibility("default")))
#else
#define B2W_API
#endif

/* element types of waveform / plane arguments */
#define B2W_F64 0
#define B2W_F32 1
#define B2W_I16 2 /* waveform only: value / 32768 (what soundfile.read returns for PCM16, A:114) */

/* bits of the device-side status word */
#define B2W_STATUS_F0_TOO_HIGH 1   /* smoothing half-width exceeds the kernel's static bound (f0 > ~fs/8) */
#define B2W_STATUS_ZERO_PERIODOGRAM 2 /* pysptk: "zero(s) are found in periodogram" (RuntimeError) */
#define B2W_STATUS_SOLVE_FAILED 4  /* pysptk: "failed to compute mcep; error occured in theq" */
#define B2W_STATUS_NOT_CONVERGED 8 /* informational: Newton loop hit maxiter (pysptk returns normally) */

B2W_API int b2w_version(void);
B2W_API const char* b2w_last_error(void);

/* Ragged batch of utterances with a cached F0 track: the common input of the analysis kernels. */
typedef struct {
  const void* x;                    /* packed waveform, x_dtype */
  int32_t x_dtype;                  /* B2W_F64 | B2W_F32 | B2W_I16 */
  int32_t num_utts;
  double preemphasis;               /* y[n] = x[n] - p*x[n-1], y[0] = x[0], per utterance, applied on load (A:117-118) */
  const int64_t* utt_sample_offset; /* [num_utts + 1] */
  const int32_t* frame_utt;         /* [num_frames] utterance index of every frame */
  const double* f0;                 /* [num_frames] Hz, 0 = unvoiced */
  const double* t;                  /* [num_frames] temporal positions in seconds (pyworld: i * frame_period / 1000) */
  int64_t num_frames;
  int32_t fs;
  int32_t reserved;
} b2w_batch;

/* ---- CheapTrick: replaces pyworld.cheaptrick inside pyworld.wav2world (W:792). ------------------------------
 * sp [num_frames, fft_size/2+1] power spectral envelope, sp_dtype B2W_F64 (pyworld-compatible) or B2W_F32
 * (fused extract path), row stride sp_stride elements (>= fft_size/2+1; the fused path pads rows to a multiple of 8 floats so
 * that b2w_mcep_tc can read them with aligned 16-byte loads). fft_size in {512, 1024, 2048, 4096}. status: 1 int32.
 * B2W_F64 planes are computed in double precision throughout (1e-9 relative against WORLD); a B2W_F32 plane at fft_size 1024
 * takes the mixed-precision kernel of the fused path (double-precision waveform transform and cumulative sums, single-precision
 * cepstral transforms: <= 5e-5 relative, tolerance 1e-4). */
B2W_API int b2w_cheaptrick(const b2w_batch* b, int32_t fft_size, double q1, void* sp, int32_t sp_dtype, int64_t sp_stride,
                   int32_t* status, void* stream);

/* ---- D4C: replaces pyworld.d4c inside pyworld.wav2world (W:792). ------------------------------------------
 * Stage 1 (the expensive one): LoveTrain voicing + per-band coarse aperiodicity.
 *   coarse_db [num_frames, nap] dB (undefined where voiced == 0), voiced [num_frames] uint8.
 *   nap = b2w_num_aperiodicities(fs); the D4C fft size is derived from fs as WORLD does.
 * b2w_d4c_coarse is the fused-extraction path: single-precision FFTs, fp64 cumulative sums; a frame whose LoveTrain ratio
 * lies within 1e-5 of `threshold` is re-evaluated in double precision, so the voiced / unvoiced decisions are those of
 * b2w_d4c_coarse_f64; coarse_db agrees with it to ~1e-4 dB.  b2w_d4c_coarse_f64 computes everything in double precision
 * (what compat.pyworld.d4c uses). */
B2W_API int b2w_d4c_coarse(const b2w_batch* b, double threshold, double* coarse_db, uint8_t* voiced, int32_t* status,
                   void* stream);
B2W_API int b2w_d4c_coarse_f64(const b2w_batch* b, double threshold, double* coarse_db, uint8_t* voiced, int32_t* status,
                   void* stream);
/* Stage 2a: expand to the pyworld.d4c result ap [num_frames, fft_size/2+1] f64 (1 - 1e-12 on unvoiced rows). */
B2W_API int b2w_d4c_expand(const double* coarse_db, const uint8_t* voiced, int64_t num_frames, int32_t fs, int32_t fft_size,
                   double* ap, void* stream);
/* Stage 2b (fused extract): pyworld.code_aperiodicity(d4c(...)) (W:805) without materialising ap.
 *   bap [num_frames, nap] f32 written with row stride bap_stride (elements). */
B2W_API int b2w_bap_from_coarse(const double* coarse_db, const uint8_t* voiced, int64_t num_frames, int32_t fs,
                        int32_t fft_size, float* bap, int64_t bap_stride, void* stream);

/* ---- codec: pyworld.code_aperiodicity (W:805) / pyworld.decode_aperiodicity (W:940). ----------------------- */
B2W_API int b2w_code_aperiodicity(const double* ap, int64_t num_frames, int32_t fs, int32_t fft_size, double* bap,
                          void* stream);
B2W_API int b2w_decode_aperiodicity(const double* bap, int64_t num_frames, int32_t fs, int32_t fft_size, double* ap,
                            void* stream);
/* float32 plane for the batched fast synthesis path (b2w_synth_render_f32 reads either type) */
B2W_API int b2w_decode_aperiodicity_f32(const double* bap, int64_t num_frames, int32_t fs, int32_t fft_size, float* ap, void* stream);

/* ---- mel-cepstral analysis: replaces pysptk.mcep(itype=3|4, etype=1) (A:146). --------------------------------
 * Host helper (runs on the CPU inside the call, fp64): builds the three precomputed all-pass warping matrices
 * for (order, alpha, fft_size); K = fft_size/2+1, m = order:
 *   h_m0t  [K, np0]  np0 = b2w_mcep_pad(m+2)  column n<=m: initial mel-cepstrum from log periodogram
 *                                              (IFFT, c0/2, c[N/2]/2, freqt(alpha)); column m+1: c[0] (start value s)
 *   h_cmat [b2w_mcep_pad(m+1), K]  C(w_j) = sum_k mc[k] * cmat[k][j]  (freqt(-alpha) to order N/2, then real FFT;
 *                                              rows past m are zero)
 *   h_m2t  [K, np2]  np2 = b2w_mcep_pad(2m+1)  r~[n] = sum_j m2t[j][n] * P[j]  (real IFFT, then frqtr(alpha) to 2m)
 * The caller converts to float32 and uploads. */
B2W_API int32_t b2w_mcep_pad(int32_t n);
B2W_API int b2w_mcep_tables_host(int32_t order, double alpha, int32_t fft_size, double* h_m0t, double* h_cmat, double* h_m2t);
/* in: spectrum plane [num_frames, K] (in_dtype F64|F32), in_is_power 0 = amplitude (x*x + eps), 1 = power (x + eps)
 * out: mc [num_frames, order+1] (mc_dtype F64|F32) with row stride mc_stride elements; iters [num_frames] int32
 * (may be NULL). Newton-Raphson UELS, miniter/maxiter/threshold as pysptk (2, 30, 1e-3). */
B2W_API int b2w_mcep(const void* in, int32_t in_dtype, int32_t in_is_power, int64_t num_frames, int32_t fft_size,
             int32_t order, double alpha, int32_t miniter, int32_t maxiter, double threshold, double eps,
             const float* m0t, const float* cmat, const float* m2t, void* mc, int32_t mc_dtype, int64_t mc_stride,
             int32_t* iters, int32_t* status, void* stream);
/* Tensor-core version of b2w_mcep for order <= 59 (the production path): the contractions run as tcgen05.mma kind::tf32
 * with the 3xTF32 split, 128 frames per CTA, accumulators in tensor memory, the matrices streamed by bulk async copies
 * from two pre-tiled device buffers of b2w_mcep_tc_stream_floats(fft_size) floats each, which b2w_mcep_tc_pretile builds
 * from the fp32 device copies of the b2w_mcep_tables_host matrices. */
B2W_API int64_t b2w_mcep_tc_stream_floats(int32_t fft_size);
B2W_API int b2w_mcep_tc_pretile(int32_t order, int32_t fft_size, const float* m0t, const float* cmat, const float* m2t,
                                float* stream0, float* stream1, void* stream);
/* in_stride: row stride of `in` in elements (>= K) */
B2W_API int b2w_mcep_tc(const void* in, int32_t in_dtype, int32_t in_is_power, int64_t in_stride, int64_t num_frames,
                        int32_t fft_size, int32_t order, double alpha, int32_t miniter, int32_t maxiter, double threshold, double eps,
                        const float* stream0, const float* stream1, void* mc, int32_t mc_dtype, int64_t mc_stride,
                        int32_t* iters, int32_t* status, void* stream);
/* log-amplitude spectrum from mel-cepstra: Re pysptk.mgc2sp(mc, alpha, gamma=0, fftlen) (A:252), optionally
 * exponentiated (A:256 np.exp(...)) : out[f][j] = (do_exp ? exp : id)(scale * sum_k mc[f][k] cmat[k][j]).
 * scale = 1 for mgc2sp / mcep_to_amp_sp, 2 (with do_exp) gives the power spectrum; do_exp = 2: the float32 amplitude squared in
 * float64 (what world_features_to_raw feeds to synthesis, W:924). */
B2W_API int b2w_mc2sp(const void* mc, int32_t mc_dtype, int64_t mc_stride, int64_t num_frames, int32_t fft_size,
              int32_t order, const float* cmat, double scale, int32_t do_exp, void* out, int32_t out_dtype,
              void* stream);
/* Tensor-core version for order <= 59 and a float32 output plane (the batched synthesis path, Synthesiser.py:39-80 -> A:304-327):
 * tiles of 128 frames as tcgen05.mma kind::tf32 GEMMs with the 3xTF32 split against the Cmat chunks of the pre-tiled Newton
 * stream of b2w_mcep_tc_pretile (stream1, same order / alpha / fft_size). */
B2W_API int b2w_mc2sp_tc(const void* mc, int32_t mc_dtype, int64_t mc_stride, int64_t num_frames, int32_t fft_size, int32_t order,
                 const float* stream1, double scale, int32_t do_exp, float* out, void* stream);

/* ---- label preparation: lf0 / vuv (W:798-802, U:40-86 interpolate_lin). -----------------------------------
 * One ragged batch; float32 arithmetic identical to the reference. lf0/vuv written with element stride
 * out_stride (so they can land directly in the packed [T, D] feature rows). */
B2W_API int b2w_lf0_vuv(const double* f0, const int64_t* utt_frame_offset, int32_t num_utts, double f0_silence_threshold,
                double lf0_zero, float* lf0, float* vuv, int64_t out_stride, void* stream);

/* ---- deltas + normalisation statistics (U:103-105 np.gradient, N:43-47 add_sample, NC:44-48). -------------
 * feats [num_frames, dim] f32 (row stride feat_stride); deltas/ddeltas may be NULL. */
B2W_API int b2w_deltas(const float* feats, int64_t feat_stride, int32_t dim, const int64_t* utt_frame_offset,
               int32_t num_utts, int64_t num_frames, float* deltas, float* ddeltas, int64_t out_stride, void* stream);
/* sums [2*dim] f64 += (sum x, sum x^2) over all rows; gram [dim*dim] f64 += X^T X if not NULL.
 * Accumulates (atomically, fp64) into the caller's buffers, which the caller zeroes once per corpus. */
B2W_API int b2w_stats_accumulate(const float* feats, int64_t feat_stride, int32_t dim, int64_t num_frames, double* sums,
                         double* gram, void* stream);

/* ---- synthesis: replaces pyworld.synthesize (W:943) on a ragged batch. ---------------------------------------
 * Inputs are the pyworld arguments per frame: f0 [F], sp [F, K] power, ap [F, K] (sp/ap dtype F64|F32).
 * Stage 1 - time base: per-sample phase increments (parallel), running phase (sequential per utterance, exactly
 *   WORLD's rounding: pulse positions are threshold decisions), pulse detection (chunk-parallel: count pass + ordered write
 *   pass).  phase_ws: one double per output sample; chunk_ws: num_utts * b2w_synth_timebase_chunks(max_out_per_utt) int32.  pulse_index/pulse_shift/pulse_vuv: [capacity] per-utterance slabs at utt_pulse_offset[u] (caller sizes
 *   them with b2w_synth_max_pulses); num_pulses [num_utts] int32 out.
 * Stage 2 - render: one CTA per pulse builds the periodic + aperiodic minimum-phase responses and writes
 *   response [total pulses, fft_size] f64.
 * Stage 3 - overlap-add: atomic-free gather; every output sample sums, in pulse order, the responses covering it. */
B2W_API int64_t b2w_synth_max_pulses(int64_t y_length, int32_t fs);
B2W_API int b2w_synth_randn_table(double* table, int64_t n, void* stream); /* WORLD randn() stream after randn_reseed(): xorshift128 with GF(2) jump-ahead */
B2W_API int64_t b2w_synth_timebase_chunks(int64_t max_out_per_utt); /* chunk_ws needs num_utts * this many int32 */
B2W_API int b2w_synth_timebase(const double* f0, const int64_t* utt_frame_offset, const int64_t* utt_out_offset,
                               const int64_t* utt_pulse_offset, int32_t num_utts, int64_t max_out_per_utt, int32_t fs,
                               double frame_period_ms, int32_t fft_size, double* phase_ws, int32_t* chunk_ws, int32_t* pulse_index,
                               double* pulse_shift, uint8_t* pulse_vuv, int32_t* num_pulses, int32_t* status, void* stream);
B2W_API int b2w_synth_render(const void* sp, const void* ap, int32_t plane_dtype, const int64_t* utt_frame_offset,
                     const int64_t* utt_pulse_offset, const int32_t* num_pulses, int32_t num_utts,
                     const int32_t* pulse_index, const double* pulse_shift, const uint8_t* pulse_vuv,
                     const double* randn_table, int64_t randn_table_len, int32_t fs, double frame_period_ms,
                     int32_t fft_size, int64_t max_pulses_per_utt, double* response, void* stream);
/* Fast path of the batched synthesis (Synthesiser.run_world_synth): one warp per pulse, single-precision transforms, float32
 * responses [total_rows, fft_size] with total_rows = utt_pulse_offset[num_utts]; fft_size must be 1024 (fs <= 32 kHz).  Pulse
 * decisions (frame indices, voiced test) are those of b2w_synth_render; waveform SNR against it ~ 110 dB. */
B2W_API int b2w_synth_render_f32(const void* sp, const void* ap, int32_t plane_dtype, const int64_t* utt_frame_offset,
                     const int64_t* utt_pulse_offset, const int32_t* num_pulses, int32_t num_utts,
                     const int32_t* pulse_index, const double* pulse_shift, const uint8_t* pulse_vuv,
                     const double* randn_table, int64_t randn_table_len, int32_t fs, double frame_period_ms,
                     int32_t fft_size, int64_t total_rows, float* response, void* stream);
B2W_API int b2w_synth_overlap_add_f32(const float* response, const int64_t* utt_out_offset, const int64_t* utt_pulse_offset,
                          const int32_t* num_pulses, int32_t num_utts, const int32_t* pulse_index, int32_t fft_size,
                          int64_t max_out_per_utt, double deemphasis, void* y, int32_t y_dtype, void* stream);
B2W_API int b2w_synth_overlap_add(const double* response, const int64_t* utt_out_offset, const int64_t* utt_pulse_offset,
                          const int32_t* num_pulses, int32_t num_utts, const int32_t* pulse_index, int32_t fft_size,
                          int64_t max_out_per_utt, double deemphasis, void* y, int32_t y_dtype, void* stream);

/* ---- Neural-VTLN all-pass warp: replaces AllPassWarp.forward's einsum+bmm (L:148-173, L:186-205). ------------
 * x, y: [rows, blocks*n] f32; alpha [rows] f32 (already combined, L:176-184). Per n-block:
 * x'[0] = x[0]/2, y = x' W(alpha), y[0] *= 2 with W(alpha) = freqt-matrix(n-1 -> n-1, alpha)^T evaluated by the
 * fp64-free all-pass recursion in registers (never materialises the [rows, n, n] tensor).
 * mean/std_dev (may be NULL, [blocks*n]) fold AllPassWarpLayer._denormalise/_normalise (LL:186-200) into the kernel. */
B2W_API int b2w_allpass_forward(const float* x, const float* alpha, int64_t rows, int32_t n, int32_t blocks,
                        const float* mean, const float* std_dev, float* y, void* stream);
/* backward: grad_x [rows, blocks*n], grad_alpha [rows] from grad_y (and x, alpha): gx = S1 frqtr(S2 gy, -alpha)
 * (A(alpha)^T == frqtr-matrix(-alpha)), galpha by the tangent of the forward recursion. */
/* Tensor-core version of b2w_allpass_forward for n % 4 == 0, n <= 64: tiles of 128 consecutive (row, block) units that share
 * one alpha run as a 3xTF32 tcgen05 GEMM against the warp matrix built on chip (HBM-bound); the other tiles are flagged in
 * tile_flags [ceil(rows * blocks / 128)] bytes (workspace) and computed by the recursion kernel in a second launch. */
B2W_API int b2w_allpass_forward_tc(const float* x, const float* alpha, int64_t rows, int32_t n, int32_t blocks, const float* mean,
                                   const float* std_dev, float* y, uint8_t* tile_flags, void* stream);
B2W_API int b2w_allpass_forward_masked(const float* x, const float* alpha, int64_t rows, int32_t n, int32_t blocks, const float* mean,
                                       const float* std_dev, float* y, const uint8_t* tile_mask, void* stream);
/* Tensor-core backward for n % 4 == 0, n <= 64: per tile of 128 units that share alpha two 3xTF32 GEMMs (grad_x through the
 * transposed warp matrix, d y / d alpha through the tangent matrix of the same wavefront recursion); other tiles are flagged in
 * tile_flags and computed by the recursion kernel.  unit_workspace [rows * blocks] floats, tile_flags [ceil(rows*blocks/128)]. */
B2W_API int b2w_allpass_backward_tc(const float* grad_y, const float* x, const float* alpha, int64_t rows, int32_t n, int32_t blocks,
                                    const float* mean, const float* std_dev, float* grad_x, float* grad_alpha, float* unit_workspace,
                                    uint8_t* tile_flags, void* stream);
B2W_API int b2w_allpass_backward_masked(const float* grad_y, const float* x, const float* alpha, int64_t rows, int32_t n, int32_t blocks,
                                        const float* mean, const float* std_dev, float* grad_x, float* grad_alpha, float* unit_workspace,
                                        const uint8_t* tile_mask, void* stream);
B2W_API int b2w_allpass_backward(const float* grad_y, const float* x, const float* alpha, int64_t rows, int32_t n,
                         int32_t blocks, const float* mean, const float* std_dev, float* grad_x, float* grad_alpha,
                         float* unit_workspace /* [rows * blocks] */, void* stream);

/* ---- MLPG: replaces MLPG.generation (idiaptts/misc/mlpg.py:94-127) as called from
 * WorldFeatLabelGen._postprocess_world (W:357-415).  feats [F, >= 3 D] rows [static(D) | delta(D) | delta-delta(D)] (F64|F32, row stride
 * feat_stride), var3 [3 D] = the diagonal of the covariance, frame_off [num_utts + 1], num_frames = F = frame_off[num_utts] (the
 * host knows it); out [F, D] fp64 (row stride out_stride); workspace: b2w_mlpg_workspace_doubles(F, D) doubles (the intermediate
 * vector [F][D] and the factor table shared by all utterances of a dimension, csrc/mlpg.cu). */
B2W_API int64_t b2w_mlpg_workspace_doubles(int64_t num_frames, int32_t D);
B2W_API int b2w_mlpg(const void* feats, int32_t feats_dtype, int64_t feat_stride, const double* var3, const int64_t* frame_off,
                     int32_t num_utts, int32_t D, int64_t num_frames, double* workspace, double* out, int64_t out_stride, void* stream);
/* ---- trainer-facing batch: for feature rows that are already on the device, replaces
 * WorldFeatLabelGen.preprocess_sample ((x - mean) / std_dev, W:279-336) + ModularModelHandlerPyTorch.prepare_batch
 * (idiaptts/src/neural_networks/pytorch/ModularModelHandlerPyTorch.py:389-499: pad_sequence to the longest utterance with
 * zeros + sequence_mask) and, after inference, the inverse (postprocess_sample de-normalisation, W:338-355, back to ragged rows).
 * feats [num_frames, width] f32 (row stride feat_stride); out [t_max, num_utts, width] (batch_first 0) or
 * [num_utts, t_max, width] (1); mask (may be NULL) same leading shape, 1.0 on real frames; mean / std_dev [width] f32 or NULL.
 * float32 arithmetic with separate operations: bit-identical to numpy. */
B2W_API int b2w_pad_normalise(const float* feats, int64_t feat_stride, int32_t width, const int64_t* utt_frame_offset,
                              int32_t num_utts, int32_t t_max, const float* mean, const float* std_dev, int32_t batch_first,
                              float* out, float* mask, void* stream);
B2W_API int b2w_unpad_denormalise(const float* padded, int32_t width, const int64_t* utt_frame_offset, const int32_t* frame_utt,
                                  int64_t num_frames, int32_t num_utts, int32_t t_max, const float* mean, const float* std_dev,
                                  int32_t batch_first, float* feats, int64_t feat_stride, void* stream);

/* ---- F0 estimation: pyworld.dio (speed = 1) and pyworld.stonemask, the F0 half of pyworld.wav2world
 * (W:792; world/LF0LabelGen.py:263-264).  b->f0 is ignored by b2w_dio and is the initial track for b2w_stonemask; b->t holds
 * i * frame_period / 1000 per utterance (as pyworld.dio returns).  num_samples = utt_sample_offset[num_utts] (the host knows
 * it; the library never reads device memory on the host).  utt_frame_offset [num_utts + 1].
 * DIO's two filters run as direct linear convolutions (WORLD sizes its FFT so that the circular convolution never wraps).
 * step2_sections: 0 = FixF0Contour step 2 as the reference's fixtures were produced (erosion by voice_range_minimum frames),
 * 1 = the later WORLD variant (drop voiced sections shorter than voice_range_minimum).
 * workspace: b2w_dio_workspace_bytes(...) bytes (~ 8 + 112 * num_bands / 7 bytes per sample: chunk long corpora). */
B2W_API int32_t b2w_dio_num_bands(double f0_floor, double f0_ceil, double channels_in_octave);
B2W_API int64_t b2w_dio_workspace_bytes(int64_t num_samples, int32_t num_utts, int64_t num_frames, int32_t fs, double f0_floor,
                                        double f0_ceil, double channels_in_octave);
B2W_API int b2w_dio(const b2w_batch* b, int64_t num_samples, const int64_t* utt_frame_offset, double f0_floor, double f0_ceil,
                    double channels_in_octave, double frame_period, double allowed_range, int32_t step2_sections,
                    void* workspace, double* f0_out, void* stream);
B2W_API int b2w_stonemask(const b2w_batch* b, double* refined_f0, void* stream);

/* ---- objective metrics: the sums behind Metrics.mcd_k / f0_rmse / gross_pitch_error / voicing_decision_error /
 * f0_frame_error / aperiodicity_distortion (idiaptts/src/Metrics.py:84-164) for a ragged batch.  org / out: float32 rows
 * [coded_sp(D) | lf0 | vuv | bap(nap)] with row stride `stride`; frame_utt [num_frames]; acc [num_utts][8] fp64, ZEROED by the caller
 * (layout in csrc/metrics.cu). */
B2W_API int b2w_world_metrics(const float* org, const float* out, int64_t stride, const int32_t* frame_utt, int64_t num_frames,
                              int32_t num_coded_sps, int32_t num_bap, double* acc, void* stream);
/* measurement aid (bench.py): launches a pure fp64 FMA kernel, returns the number of FMAs it executes (or -1) */
B2W_API int64_t b2w_probe_fp64_fma(int32_t iters, double* scratch, void* stream);
/* development aid: phase cycle counters of mcep_tc CTA 0 (non-zero only in a -DB2W_MCEP_PROF build) */
B2W_API int b2w_mcep_prof_read(long long* out16);
B2W_API int b2w_vtf_prof_read(long long* out16);   /* same for allpass_tc_forward_kernel (-DB2W_VTF_PROF builds) */

/* ---- generalised mel-cepstrum (sp_type = "mgc"): replaces pysptk.mgcep(amp_sp, order, alpha, gamma, eps = 1e-8,
 * etype = 1, itype = 3) in AudioProcessing.extract_mgc (A:123-140) and exp(Re pysptk.mgc2sp(mgc, alpha, gamma, fftlen)) in
 * AudioProcessing.mgc_to_amp_sp (A:259-275).  PARITY UNPINNED (csrc/mgcep.cu, oracle/mgc_np.py).  gamma in [-1, 0).
 * Tables (float32, device; layouts in csrc/mgcep.cu, built by ops.MgcTables): m0t [K, pad4(order + 2)] of b2w_mcep_tables_host,
 * fwd_cos / fwd_sin [pad4(order + 1), K], red_cos / red_sin [K, pad4(2 order + 1)]. */
B2W_API int b2w_mgcep(const void* in, int32_t in_dtype, int32_t in_is_power, int64_t num_frames, int32_t fft_size, int32_t order,
                      double gamma, int32_t miniter, int32_t maxiter, double threshold, double eps, const float* m0t,
                      const float* fwd_cos, const float* fwd_sin, const float* red_cos, const float* red_sin, void* mgc,
                      int32_t mgc_dtype, int64_t mgc_stride, int32_t* iters, int32_t* status, void* stream);
B2W_API int b2w_mgc2sp(const void* mgc, int32_t mgc_dtype, int64_t mgc_stride, int64_t num_frames, int32_t fft_size, int32_t order,
                       double gamma, const float* fwd_cos, const float* fwd_sin, void* amp, int32_t amp_dtype, void* stream);

/* ---- corpus file IO (host only, no device work, no stream): the formats on either side of the GPU batch, handled
 * for a whole shard by a pool of `threads` host threads (0 = one per hardware thread, at most 64).  Errors: -1 and
 * b2w_last_error() names the first file that failed; nothing is partially reported as success.
 *
 * wav in: replaces the per-utterance soundfile.read of AudioProcessing.get_raw (A:108-120) inside the gen_data loop (W:996-1013).
 * b2w_wav_probe parses the RIFF headers (PCM, or extensible with PCM sub-format): per file the number of sample frames, sampling
 * rate, bits per sample, channels and the byte offset of the data chunk.  b2w_wav_read_i16 reads 16-bit mono files into ONE packed
 * buffer (samples of file i at utt_sample_offset[i] .. [i + 1]) -- the int16 waveform layout of b2w_batch, so the (pinned) buffer
 * goes to the device as it is; the kernels scale by 1 / 32768 on load, which is the float64 soundfile returns.  Other widths /
 * channel counts are the caller's general path.
 *
 * npz out / in: replaces numpy.savez per feature and utterance in LabelGen.save_output / WorldFeatLabelGen.save_output
 * (LabelGen.py:63-101, W:1121-1172) and numpy.load in WorldFeatLabelGen.load_sample (W:459-567).  One call handles one feature
 * directory: file i is a ZIP archive (stored) with one member "<keys[k]>.npy" per key, a C-ordered little-endian float32 matrix
 * [utt_frame_offset[i + 1] - utt_frame_offset[i], cols[k]] = rows utt_frame_offset[i] .. of the packed host matrix `feats` (row
 * stride feat_stride floats), columns col_offset[k] .. + cols[k].  Archives are byte-compatible with numpy (numpy.load reads
 * them; the reader takes numpy.savez archives, ZIP64 headers included).  Compressed archives (savez_compressed) and other dtypes
 * are refused with a message so that the caller can fall back to numpy.  b2w_npz_probe returns the shape of one key per file (1-D
 * arrays report cols = 1); b2w_npz_read_f32 checks every shape against the offsets it was given and, with verify_crc, the CRC-32
 * the way zipfile does. */
B2W_API int b2w_wav_probe(const char* const* paths, int32_t num_files, int64_t* num_samples, int32_t* fs, int32_t* bits,
                          int32_t* channels, int64_t* data_offset, int32_t threads);
B2W_API int b2w_wav_read_i16(const char* const* paths, int32_t num_files, const int64_t* data_offset, const int64_t* utt_sample_offset,
                             int16_t* samples, int32_t threads);
/* wav out: the soundfile.write of Synthesiser.run_world_synth (idiaptts/src/Synthesiser.py:68-73) for a whole batch: file i gets the
 * float32 samples [utt_sample_offset[i], utt_sample_offset[i + 1]) of the packed host buffer as 16-bit mono PCM,
 * clip(round_half_even(32767 x), -32768, 32767). */
B2W_API int b2w_wav_write_pcm16(const char* const* paths, int32_t num_files, const int64_t* utt_sample_offset, const float* samples,
                                int32_t fs, int32_t threads);
B2W_API int b2w_npz_write_f32(const char* const* paths, int32_t num_files, const char* const* keys, int32_t num_keys,
                              const int32_t* col_offset, const int32_t* cols, const int64_t* utt_frame_offset, const float* feats,
                              int64_t feat_stride, int32_t threads);
B2W_API int b2w_npz_probe(const char* const* paths, int32_t num_files, const char* key, int64_t* rows, int32_t* cols, int32_t threads);
B2W_API int b2w_npz_read_f32(const char* const* paths, int32_t num_files, const char* const* keys, int32_t num_keys,
                             const int32_t* col_offset, const int32_t* cols, const int64_t* utt_frame_offset, float* feats,
                             int64_t feat_stride, int32_t verify_crc, int32_t threads);

/* development aid (tests): the archives' CRC-32 of a host buffer; variant 0 = table code, 1 = carry-less-multiplication folding
 * where the CPU has it (what the writers / readers use) */
B2W_API uint32_t b2w_crc32(const void* data, int64_t n, int32_t variant);

/* ---- scalar helpers (host, pure functions): pyworld.get_cheaptrick_fft_size (A:60), get_num_aperiodicities
 * (A:71), and the D4C transform size. */
B2W_API int32_t b2w_cheaptrick_fft_size(int32_t fs, double f0_floor);
B2W_API int32_t b2w_num_aperiodicities(int32_t fs);
B2W_API int32_t b2w_d4c_fft_size(int32_t fs);

#ifdef __cplusplus
}
#endif
#endif /* B200WORLD_H_ */
