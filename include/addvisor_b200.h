/* addvisor_b200.h - C ABI of libaddvisor_sm100.so
 *
 * B200 (sm_100a) kernels for the ADDvisor explanation-evaluation hot path.  The reference
 * (davidcombei/xAI-Audio-Deepfakes) is pure Python and has no FFI of its own; the boundary
 * a maintainer binds is its Python call surface.  Each entry point below names the
 * reference call it replaces (paths relative to the reference checkout).  The ctypes
 * binding that ships with this repo is xai-audio-deepfakes_b200/_lib.py; INTEGRATION.md
 * shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer marked "dev" is a device pointer on the current CUDA device; "host" is host memory
 *   - `stream` is a cudaStream_t passed as void*; all compute entry points are asynchronous on it
 *   - return value: 0 (ADV_OK) or a negative adv_status; nothing throws; adv_strerror() names codes
 *   - nothing is allocated behind the caller's back except plan-owned tables
 *   - spectra are "frame-major": element (b, t, f) of a [B, T, F] complex64 array, i.e. the memory
 *     layout torch.stft itself produces (its [B,F,T] result has strides (F*T, 1, F))
 *   - there is NO CPU fallback: without a CUDA device every compute call returns ADV_ERR_CUDA
 */
#ifndef ADDVISOR_B200_H
#define ADDVISOR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum adv_status {
    ADV_OK = 0,
    ADV_ERR_INVALID = -1,     /* bad argument (null pointer, negative size, ...) */
    ADV_ERR_UNSUPPORTED = -2, /* geometry outside what the kernels implement (n_fft not 512/1024, ...) */
    ADV_ERR_NOLA = -3,        /* window overlap-add envelope vanishes: torch.istft "window overlap add min: 1" */
    ADV_ERR_SHORT_INPUT = -4, /* reflect padding needs n_in > n_fft/2 (torch.stft raises too) */
    ADV_ERR_CUDA = -5,        /* CUDA runtime error; see adv_last_cuda_error() */
    ADV_ERR_SHAPE = -6        /* mask / spectrum shape inconsistent with the plan */
} adv_status;

typedef enum adv_mask_mode {
    ADV_MASK_LOG1P = 0, /* LMAC_metrics.py:136-143,151-153: expm1(m * log1p(|X|)) * exp(i*phase) */
    ADV_MASK_LINEAR = 1 /* loss_function.py:36-45:          (m * |X|) * exp(i*phase)            */
} adv_mask_mode;
/* OR-ed into `mode`: bins outside a mask smaller than the spectrum ([Fm,Tm] < [F,T]) are removed from BOTH outputs,
 * which is what the reference's crop of magnitude and phase to the mask's extent amounts to (LMAC_metrics.py:136-139,
 * loss_function.py:36-41).  Without the flag the mask is zero-extended: out-of-mask bins go to the masked-out branch. */
enum { ADV_MASK_DROP_OUTSIDE = 0x100 };

typedef struct adv_plan adv_plan;
typedef struct adv_c64 { float re, im; } adv_c64;

int adv_version(void);
/* Programmatic dependent launch for the transform / normalise / metric kernels (a kernel's set-up overlaps the tail of
 * the previous kernel of the same stream; every kernel waits for its predecessor before touching data).  Process-wide,
 * on by default; returns the previous setting.  Multi-stream pipelines that co-schedule kernels on one SM may prefer it
 * off (pipeline.PipelinedPool does). */
int adv_set_pdl(int on);
const char* adv_strerror(int status);
const char* adv_last_cuda_error(void);

/* ---- plan = AudioProcessor.__init__ geometry (audioprocessor.py:22-44) bound to one clip length --------
 * window_host: win_length taps or NULL for torch's default rectangular window (audioprocessor.py:102-108
 * passes no window).  The window is centred in n_fft like torch.stft does.  n_frames = 1 + n_in / hop for
 * the stft direction; n_out is torch.istft's `length` (audioprocessor.py:121-129) or hop*(n_frames-1)
 * for the length=None call sites (hifigan.py:223-225).  Fails with ADV_ERR_NOLA exactly when torch.istft
 * would raise on the envelope check. */
int adv_plan_create(adv_plan** out, int n_fft, int hop, int win_length, const float* window_host,
                    int n_frames, int n_in, int n_out);
void adv_plan_destroy(adv_plan* plan);
int adv_plan_bins(const adv_plan* plan);
int adv_plan_frames(const adv_plan* plan);
/* number of output tiles per clip the istft / explain kernels use for batch size B
 * (= second dimension of the `stats` array below) */
int adv_plan_tiles(const adv_plan* plan, int batch);        /* adv_explain / adv_explain_spec */
int adv_plan_tiles_istft(const adv_plan* plan, int batch);  /* adv_istft */

/* ---- AudioProcessor.compute_stft (audioprocessor.py:82-112): torch.stft + .abs() + .angle() -----------
 * wav: dev float [B][wav_stride], first n_in samples of each row are used (caller pads/crops, :83-98).
 * X: dev complex64 [B][T][F] frame-major; mag / phase: dev float [B][T][F] or NULL to skip. */
int adv_stft(const adv_plan* plan, const float* wav, int64_t wav_stride, int batch,
             adv_c64* X, float* mag, float* phase, void* stream);

/* Same transform with options.  ADV_STFT_ZERO_PAD: frames that reach past the clip edges read zeros instead of
 * torch.stft's reflect padding - with the synthesis window and the input pre-multiplied by the reciprocal
 * overlap-add envelope this is the adjoint of adv_istft, i.e. the backward pass of compute_invert_stft in the
 * training loss (loss_function.py:46-47 under autograd, train_addvisor.py:364-366). */
enum { ADV_STFT_ZERO_PAD = 1 };
int adv_stft_ex(const adv_plan* plan, const float* wav, int64_t wav_stride, int batch,
                adv_c64* X, float* mag, float* phase, int flags, void* stream);
/* copies the plan's reciprocal overlap-add envelope 1 / (n_fft * sum_t w^2) over [0, n_out) to dev float out[n_out] */
int adv_plan_inv_env(const adv_plan* plan, float* out, void* stream);
/* Gradient of the linear mask path (loss_function.py:36-47) with respect to the mask:
 *   gm[b][f][t] = c_f * Re(conj(X[b][t][f]) * A[b][t][f]),  f < Fm, t < Tm,   c_f = 1 for DC / Nyquist, 2 otherwise,
 * A = adv_stft_ex(ZERO_PAD) of (grad_rel - grad_irr) * inv_env.  X: element strides (sb, st, sf); A frame-major
 * [B][T][F]; gm dev float [B][Fm][Tm] (the mask network's layout). */
int adv_mask_grad_linear(const adv_c64* X, int64_t sb, int64_t st, int64_t sf, const adv_c64* A, int batch, int F, int T,
                         int Fm, int Tm, float* gm, void* stream);

/* ---- AudioProcessor.compute_invert_stft (audioprocessor.py:117-131): torch.istft ----------------------
 * X: dev complex64, element (b,t,f) at X[b*sb + t*st + f*sf] (element strides); out: dev float [B][n_out].
 * stats (nullable): dev double [B][tiles][2] per-tile (sum, sum of squares) of the output, for the
 * normaliser that follows in extract_features (classifier_embedder.py:59-63). */
int adv_istft(const adv_plan* plan, const adv_c64* X, int64_t sb, int64_t st, int64_t sf, int batch,
              float* out, double* stats, void* stream);

/* ---- fused explanation resynthesis: the LMAC_metrics.py:125-158 loop body without the classifier ------
 * wave + mask -> STFT -> mask / (1-mask) applied to the (log-)magnitude with the original phase ->
 * two iSTFTs.  Spectra never leave the SM.  mask: dev float [B][Fm][Tm] contiguous, Fm <= F, Tm <= T,
 * cells outside the mask count as mask = 0 (or are dropped from both outputs: mode | ADV_MASK_DROP_OUTSIDE).
 * rel / irr: dev float [B][n_out].
 * stats (nullable): dev double [B][tiles][4] = per-tile (sum rel, sumsq rel, sum irr, sumsq irr). */
int adv_explain(const adv_plan* plan, const float* wav, int64_t wav_stride, const float* mask, int Fm, int Tm,
                int mode, int batch, float* rel, float* irr, double* stats, void* stream);
/* same, starting from an existing spectrum (collate_fn already ran compute_stft, LMAC_metrics.py:109-114) */
int adv_explain_spec(const adv_plan* plan, const adv_c64* X, int64_t sb, int64_t st, int64_t sf,
                     const float* mask, int Fm, int Tm, int mode, int batch,
                     float* rel, float* irr, double* stats, void* stream);

/* ---- standalone mask-apply on (magnitude, phase) as the reference writes it ---------------------------
 * mag / phase: dev float [B][T][F] frame-major; mask as above; rel / irr: dev complex64 [B][T][F]. */
int adv_mask_apply(const float* mag, const float* phase, const float* mask, int batch, int T, int F,
                   int Fm, int Tm, int mode, adv_c64* rel, adv_c64* irr, void* stream);

/* ---- zero_mean_unit_var_norm (classifier_embedder.py:59-63): (x - mean) / (std_unbiased + 1e-7) -------
 * adv_row_stats: per-row partial (sum, sumsq) -> stats dev double [B][parts][2], parts = adv_row_stats_parts(n).
 * adv_normalize: in/out dev float [B][n] (may alias); stats dev double [B][parts][width], the (sum, sumsq)
 * pair at column `col`.  Two launches, no host sync. */
int adv_row_stats_parts(int n);
int adv_row_stats(const float* in, int batch, int n, double* stats, void* stream);
int adv_normalize(const float* in, float* out, int batch, int n, const double* stats, int parts, int width,
                  int col, void* stream);
/* both outputs of adv_explain normalised in place by one launch; stats = its [B][parts][4] array */
int adv_normalize_pair(float* rel, float* irr, int batch, int n, const double* stats, int parts, void* stream);

/* adv_normalize_pair and the metric reduction of the same batch in ONE launch (an extra CTA of the normaliser's grid):
 * the two are independent, and the fused form keeps the 4 us reduction from holding an SM after the normaliser is done.
 * n_logits <= 1024 (else ADV_ERR_UNSUPPORTED: use adv_lmac_reduce); flags / scores / sums as for adv_lmac_reduce. */
int adv_normalize_pair_lmac(float* rel, float* irr, int batch, int n, const double* stats, int parts, const float* p,
                            const float* theta, const float* q, int n_logits, int flags, float* scores, double* sums,
                            void* stream);

/* ---- LMAC metrics (LMAC_metrics.py:31-73,160-172; sigmoid of classifier_embedder.py:36) ---------------
 * p / theta / q: dev float [n] classifier outputs for the clean, masked-in and masked-out clips;
 * flags: ADV_LMAC_LOGITS applies the logistic sigmoid first; ADV_LMAC_ACCUMULATE adds into `sums`
 * instead of overwriting it (running totals over the batches of one evaluation).  scores (nullable): dev float [n][7] =
 * (faithfulness, fidelity, AD, AI, AG, pc, oc) per sample, pc / oc = get_score_for_predicted_class of p / theta.  sums: dev double [6] = the five sums and the count
 * (the vector the multi-GPU driver all-reduces).  block_partials: dev double scratch
 * [adv_lmac_blocks(n)][5]. */
enum { ADV_LMAC_LOGITS = 1, ADV_LMAC_ACCUMULATE = 2 };
int adv_lmac_blocks(int n);
int adv_lmac_reduce(const float* p, const float* theta, const float* q, int n, int flags,
                    float* scores, double* sums, double* block_partials, unsigned int* counter, void* stream);

/* ---- gradient-saliency time-domain mask (captum_saliency.py:136-143) ----------------------------------
 * m = |attr| / (max_row |attr| + 1e-8); rel = wave*m; irr = wave*(1-m).  rowmax: dev float [B] scratch. */
int adv_td_mask(const float* wave, const float* attr, int batch, int n, float* mask_out, float* rel, float* irr,
                float* rowmax, void* stream);

/* ---- ADDvisor mask head (addvisor.py:57-60,82): sigmoid(conv1x1(32 -> 1)) -----------------------------
 * y1: dev float [B][C][HW] contiguous, w: dev float [C], bias: dev float [1]; mask: dev float [B][HW]. */
int adv_mask_head(const float* y1, const float* w, const float* bias, int batch, int channels, int64_t hw,
                  float* mask, void* stream);

/* ---- band swap (train_logReg_swapping.py:64-75, hifigan.py:206-214): rows [f_lo, f_hi) of `voc` replace
 * those of `real`; all three frame-major [B][T][F] complex64. */
int adv_band_swap(const adv_c64* real, const adv_c64* voc, int batch, int T, int F, int f_lo, int f_hi,
                  adv_c64* out, void* stream);

/* all bands of the fabrication loop in one launch: out[k][b][t][f] = (edges[k] <= f < edges[k+1]) ? voc : real,
 * k < n_bands; edges: dev int [n_bands + 1]; out: dev complex64 [n_bands][B][T][F] (one batched adv_istft follows) */
int adv_band_swap_multi(const adv_c64* real, const adv_c64* voc, int batch, int T, int F, const int* edges, int n_bands,
                        adv_c64* out, void* stream);

/* ==== tensor-core (tcgen05 / TMEM) entry points ========================================================== */

/* ---- mel filterbank projection: MelSpectrogram / mel_spectogram (audioprocessor.py:38-44, hifigan.py:163-178)
 * X: dev complex64 [rows = B*T][F] frame-major spectrum (adv_stft output); fb_hi / fb_lo: dev float [NM][Kpad],
 * the transposed filterbank split into its tf32-representable part and the remainder (3xTF32 GEMM, fp32 accuracy),
 * NM = n_mels rounded up to 64 / 80 / 128, Kpad a multiple of 32 >= F.  out: dev float [B][n_mels][T] =
 * sum_f fb[f][mel] * |X|^power, then log(max(., clip)) when log_compress != 0. */
int adv_mel_project(const adv_c64* X, int64_t rows, int T, int F, const float* fb_hi, const float* fb_lo, int Kpad,
                    int n_mels, float power, int log_compress, float clip, float* out, void* stream);

/* The same front-end in ONE launch (n_fft 1024): framed STFT of the plan -> |X|^power -> filterbank contraction on
 * tcgen05 (bf16 hi/lo split, three passes: 16 mantissa bits) -> log(max(., clip)); the spectrum never leaves the SM.
 * wav: dev float [B][wav_stride] (plan n_in samples used, reflect-padded like torch.stft centre=True); out: dev float
 * [B][n_mels][T].  The filterbank is BAND-COMPRESSED by the host (mel.MelSpectrogram): for each of the eight 64-bin K
 * chunks of bins 0..511 only the mel columns [n0, n0 + n) that are non-zero there are stored, as a [n][64] bf16
 * K-major SWIZZLE_128B tile (element (r, k) at byte r*128 + (((k>>3) ^ (r&7)) << 4) + (k&7)*2); fb_tiles = all hi
 * tiles, then (from byte lo_base) all lo tiles at the same offsets; chunk_table: HOST int [8][3] = {n0, n, byte offset}
 * (n a multiple of 8, offsets multiples of 1024); fb_nyq: dev float [ceil16(n_mels)] = the bank's row of bin 512 (added
 * by the epilogue).  Returns ADV_ERR_UNSUPPORTED when the tiles do not fit in shared memory next to the 128 KB operand
 * tile (a dense bank): the caller then uses adv_stft + adv_mel_project. */
int adv_mel_fused(const adv_plan* plan, const float* wav, int64_t wav_stride, int batch, const void* fb_tiles, int fb_bytes,
                  int lo_base, const int* chunk_table, const float* fb_nyq, int n_mels, float power, int log_compress,
                  float clip, float* out, void* stream);

/* ---- HiFi-GAN generator layers (SpeechBrain HifiganGenerator behind hifi_gan.decode_batch, hifigan.py:180) -----
 * adv_conv1d_bf16: channels-last bf16 conv1d as an implicit GEMM, "same" length, odd tap count:
 *   out[b,l,n] = out_scale * ( bias[n] + sum_{tap,ci} w[n][tap*Cin+ci] * lrelu(in[b, l+(tap-center)*dil, ci], pre_slope)
 *                              + resid[b,l,n] )
 * in [B][L][Cin], w [N][Kpad] (Kpad multiple of 64, zero padded), resid / out [B][L][N]; Cin % 8 == 0, N % 16 == 0;
 * out (raw) and / or out_act = LeakyReLU(act_slope) of the result are written;
 * pad_reflect selects reflect instead of zero padding; pre_slope = 1 disables the input activation.
 * A transposed conv (stride s) is this conv with 3 taps and N = s*Cout phase-stacked weights. */
int adv_conv1d_bf16(const void* in, const void* w, const float* bias, const void* resid, void* out, void* out_act,
                    int batch, int L, int Cin, int taps, int dil, int N, int Kpad, int pad_reflect, float pre_slope,
                    float act_slope, float out_scale, void* stream);
/* Epilogue of the wide conv kernel (C_in, N multiples of 128): 1 (default) = output blocks staged in shared memory and
 * written by the TMA engine (cp.async.bulk.tensor), 0 = per-thread 16-byte stores.  Process-wide; returns the previous
 * setting (A/B measurements: scripts/bench_vocoder.py --epilogue). */
int adv_set_conv_epilogue(int tma);

/* Same contraction on the production pipeline: operands fetched by TMA (3-D tensor map over [B][L][Cin], the tap
 * is a row coordinate, zero padding = TMA out-of-bounds fill), warp-specialised producer / MMA / epilogue roles,
 * persistent CTAs, double-buffered TMEM accumulators.  Zero padding only; Cin = 32 or a multiple of 64;
 * N % 32 == 0; w [N][taps*Cin] unpadded.  out_raw and / or out_act (LeakyReLU(act_slope) of the result, for the
 * consuming layer) are written; the input is used as stored (no activation on load). */
int adv_conv1d_bf16_tma(const void* in, const void* w, const float* bias, const void* resid, void* out_raw,
                        void* out_act, int batch, int L, int Cin, int taps, int dil, int N, float act_slope,
                        float out_scale, void* stream);
/* Fused residual unit of a HiFi-GAN ResBlock1 (hifigan.py:180 via SpeechBrain's generator), C = 32 or 64 channels:
 *   out = conv2(lrelu(conv1(lrelu(x)) + b1)) + b2 + x,  conv1 dilation `dil`, conv2 dilation 1, both `taps` wide,
 * zero "same" padding.  x / out dev bf16 [B][L][C]; w1 / w2 dev bf16 [C][taps*C] (K index = tap*C + ci); biases fp32.
 * One persistent tcgen05 kernel: the intermediate never leaves the SM.  ADV_ERR_UNSUPPORTED when both weight sets do
 * not fit in shared memory (C = 64 with 11 taps): run the two convs with adv_conv1d_bf16_tma instead. */
int adv_resunit_bf16(const void* x, const void* w1, const float* b1, const void* w2, const float* b2, void* out, int batch,
                     int L, int C, int taps, int dil, float slope, void* stream);
/* mel [B][C][T] fp32 -> channels-last bf16 [B][T+2*pad][Cpad], replicate-padded in time (inference_padding) */
int adv_mel_to_channels_last(const float* mel, int batch, int C, int T, int pad, int Cpad, void* out, void* stream);
/* MRF average of the three resblock outputs (bf16, n elements), then LeakyReLU(act_slope) (1 = identity) */
int adv_avg3_bf16(const void* a, const void* b, const void* c, int64_t n, float act_slope, void* out, void* stream);
/* Halo layout for reflect "same" padding on the TMA conv kernels (SpeechBrain's nnet.CNN.Conv1d default padding_mode,
 * i.e. what hifi_gan.decode_batch computes at hifigan.py:180): an activation is [B][L + 2 H][C] with its L rows at
 * offset H; the conv kernels run over all L + 2 H rows with zero padding and adv_halo_fix_bf16 rewrites the halo rows of
 * a produced tensor in place - mode 1: reflection of the interior (row -j := row j, row L-1+j := row L-1-j), mode 0:
 * zeros (input of a transposed conv).  C % 8 == 0; mode 1 needs H < L. */
int adv_halo_fix_bf16(void* buf, int batch, int L, int C, int H, int mode, void* stream);
/* out interior = LeakyReLU_slope(mean of a [, b [, c]] interiors) (the MRF average of HifiganGenerator.forward, or a
 * plain re-layout with one input); inputs [B][L + 2 h_in][C], out [B][L + 2 h_out][C] with zero (mode 0) / reflected
 * (mode 1) halo rows. */
int adv_avg_relayout_bf16(const void* a, const void* b, const void* c, int batch, int L, int C, int h_in, int h_out, int mode,
                          float act_slope, void* out, void* stream);
/* LeakyReLU(slope) -> conv_post (C=32 -> 1, 7 taps, w [taps][C] fp32) -> tanh; out dev float [B][L] */
int adv_post_conv_tanh(const void* in, const float* w, const float* bias, int batch, int L, int C, int taps, float slope,
                       int pad_reflect, float* out, void* stream);

/* ---- AudioProcessor.load_audio for a BATCH of clips (audioprocessor.py:49-63): decode + resample + pad / crop in one
 * launch.  in: dev int16 PCM (in_is_pcm16 != 0; scaled by 1/32768 like torchaudio.load) or dev float, clip b = the
 * lens[b] samples from element offs[b] (dev int64 / dev int arrays); orig / newf: source / target rate divided by their
 * gcd; h: dev float [newf][2*width + orig] polyphase sinc filter of torchaudio.transforms.Resample (defaults), range:
 * dev int [newf][2] first / past-last non-zero tap of every phase; out: dev float [B][n_out],
 * out[b][j] = sum_k h[j % newf][k] * x_b[(j / newf) * orig + k - width] for j < ceil(newf * len_b / orig), zero beyond
 * (F.pad to audio_length * target_sr) and cropped at n_out.  orig == newf: conversion and pad / crop only (h may be NULL). */
int adv_resample_rows(const void* in, int in_is_pcm16, const long long* offs, const int* lens, int batch, int orig, int newf,
                      int width, const float* h, const int* range, int n_out, float* out, void* stream);

/* Cross-correlation alignment shift of align_waveforms (hifigan.py:113-136):
 *   cc[j] = sum_i ref[j + i - n_deg] * deg[i], j = 0 .. n_ref + n_deg;  *shift = argmax_j cc[j] - n_deg (first maximum).
 * Direct fp32 accumulation like the reference's conv1d.  ws_val / ws_idx: dev scratch of adv_xcorr_blocks(n_ref, n_deg)
 * floats / ints; shift: dev int.  Asynchronous on `stream`. */
int adv_xcorr_blocks(int n_ref, int n_deg);
int adv_xcorr_shift(const float* ref, int n_ref, const float* deg, int n_deg, float* ws_val, int* ws_idx, int* shift,
                    void* stream);

/* The same correlation through the frequency domain (overlap-save with 1024-point frames, hop 512): O(N log N)
 * instead of the O(N^2) of the reference's conv1d.  The caller runs adv_stft_ex(ADV_STFT_ZERO_PAD) on the padded deg
 * (512-tap rectangular window) and ref (1024-tap rectangular window), then
 *   adv_xcorr_fd_mac: Z[0] = 0, Z[q + 1][k] = (-1)^k sum_b conj(D[b + 1][k]) R[q + b + 1][k], q < nq, b < nb
 *                     (D, R, Z frame-major [frames][bins]; R has t_r frames, missing frames count as zero),
 * then adv_istft (n_fft 1024, hop 512, win 512): cc[j] sits at output sample j + 256, and
 *   adv_argmax_first: *out = (first arg-max of x[0 .. n)) - sub.
 * Host composition: hifigan.align_shift. */
int adv_xcorr_fd_mac(const adv_c64* D, int nb, const adv_c64* R, int t_r, adv_c64* Z, int nq, int bins, void* stream);
int adv_argmax_first(const float* x, int n, int sub, int* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADDVISOR_B200_H */
