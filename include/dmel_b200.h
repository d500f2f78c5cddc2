/* dmel_b200.h — C ABI of the B200 dMel tokenization path.
 *
 * The reference (ishine/dmel_codec) has no FFI for this path: it is a Python
 * nn.Module (dmel_codec/utils/spectrogram.py) calling torch ops.  These entry
 * points are what a binding for that module would call instead; each one names
 * the reference code it replaces.  No torch types cross this boundary: plain
 * pointers, sizes and a CUDA stream handle.
 *
 * Conventions
 *   - every function returns 0 on success, a negative DMEL_ERR_* otherwise, and
 *     never throws; dmel_last_error() gives the message for the calling thread;
 *   - "dev" pointers are device memory on the plan's GPU, "host" pointers are
 *     host memory; the library never allocates or frees caller tensors;
 *   - work is queued on `stream` (a cudaStream_t passed as void*, NULL = the
 *     legacy default stream) and the call returns without synchronising,
 *     except the *_host_* functions, which return after the result is in the
 *     caller's host buffer;
 *   - tensors are dense row-major: waveforms (B, row_stride >= L) float32,
 *     log-mel (B, n_mels, T) float32, codes (B, n_mels, T) uint8, with
 *     T = dmel_plan_num_frames(plan, L).
 */
#ifndef DMEL_B200_H
#define DMEL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMEL_ABI_VERSION 9

#define DMEL_OK 0
#define DMEL_ERR_INVALID (-1)     /* bad argument (shape, null pointer, L <= reflect pad ...) */
#define DMEL_ERR_UNSUPPORTED (-2) /* geometry this build has no kernel for */
#define DMEL_ERR_CUDA (-3)        /* CUDA runtime error, text in dmel_last_error() */
#define DMEL_ERR_NO_DEVICE (-4)   /* no CUDA device: there is no CPU fallback */

#define DMEL_DTYPE_F32 0
#define DMEL_DTYPE_BF16 1

typedef struct dmel_plan dmel_plan;
typedef struct dmel_stream dmel_stream;

int dmel_abi_version(void);
const char* dmel_last_error(void);

/* Geometry + constants of one transform.  Replaces the lazily cached
 * mel_basis / hann_window of LinearSpectrogram.forward
 * (reference dmel_codec/utils/spectrogram.py:43-56).
 *   mel_basis_host : (n_mels, n_fft/2+1) float32, row-major  (librosa.filters.mel, :45-52)
 *   window_host    : (n_fft) float32, already centre-padded if win_length < n_fft (:53)
 *   center         : 0/1, the `center` flag handed to torch.stft (:70)
 * The reflect pad (n_fft - hop)/2 of :58-62 is implied.  Supported n_fft: powers of two from 64 to 2048 (1024 and
 * 2048 have kernels of their own; a shorter frame runs zero-extended on the 1024-point kernel, which is exact);
 * anything else returns DMEL_ERR_UNSUPPORTED.
 * The plan binds to the CUDA device that is current at the time of the call. */
int dmel_plan_create(int n_fft, int hop_length, int n_mels, int center,
                     const float* mel_basis_host, const float* window_host,
                     dmel_plan** out);
void dmel_plan_destroy(dmel_plan* plan);

/* Writes a one-line JSON description of the launch configuration chosen for this geometry
 * (frame tile, CTAs per SM, shared memory, banded filterbank size) into buf. Diagnostics only. */
int dmel_plan_describe(const dmel_plan* plan, char* buf, size_t buf_len);

/* T for a row of n_samples (reference: shape of the torch.stft output, :64-75);
 * <= 0 when the row is too short.  n_samples must exceed the reflect pad. */
long long dmel_plan_num_frames(const dmel_plan* plan, long long n_samples);

/* waveform -> log-mel.  Replaces LinearSpectrogram.forward
 * (reference dmel_codec/utils/spectrogram.py:41-81). */
int dmel_logmel_f32(dmel_plan* plan, const float* wav_dev, long long n_rows, long long n_samples,
                    long long row_stride, float* logmel_dev, void* stream);

/* waveform -> masked log-mel in the encoder's dtype, one launch.  Frames at or past
 * lengths[b] / hop are written as 0 and never computed; out_dev is (B, n_mels, T) float32
 * (DMEL_DTYPE_F32) or bfloat16 (DMEL_DTYPE_BF16, round-to-nearest-even of the float32 value).
 * lengths_dev NULL = no masking.  Replaces the transform + cast + sequence_mask multiply of
 * VQGAN.encode_unquantized (reference models/codec_lit_modules.py:486-507, mask rule
 * utils/utils.py:48-55); the caller's (B*G, n_mels/G, T) group view of the result is free.
 * row_sum_dev (optional, (B, n_mels) float32, zeroed by the caller): += the sum over time of the
 * values written, per row and channel - the mels.mean(-1) behind the training step's `quality`
 * statistic (models/codec_lit_modules.py:173) without another pass over the mel (float atomics:
 * the summation order, hence the last bit, varies from launch to launch). */
int dmel_logmel_masked(dmel_plan* plan, const float* wav_dev, long long n_rows, long long n_samples,
                       long long row_stride, const int32_t* lengths_dev, int out_dtype,
                       void* out_dev, float* row_sum_dev, void* stream);

/* Calibration pass: running per-channel min / max of the log-mel of this batch
 * over valid frames (t < lengths[b] / hop; lengths_dev may be NULL = all T).
 * min_dev / max_dev (n_mels) are UPDATED in place (initialise to +inf / -inf),
 * so batches and ranks compose.  Not in the reference;
 * valid-frame rule from reference models/codec_lit_modules.py:176-177. */
int dmel_minmax_f32(dmel_plan* plan, const float* wav_dev, long long n_rows, long long n_samples,
                    long long row_stride, const int32_t* lengths_dev,
                    float* min_dev, float* max_dev, void* stream);

/* dmel_logmel_f32 and dmel_minmax_f32 in one launch: writes the log-mel of every valid frame (frames at or past
 * lengths[b] / hop are written as 0 and never computed, as in dmel_logmel_masked; all frames when lengths_dev is
 * NULL) and folds them into the running per-channel min / max.  With the log-mel of a shard kept in HBM the
 * calibrate-then-encode job of a dataset needs the STFT only once: pass 2 is dmel_quantize_u8 over the
 * stored tensor (bit-identical to the fused encode, which quantises the same float32 values). */
int dmel_logmel_minmax_f32(dmel_plan* plan, const float* wav_dev, long long n_rows, long long n_samples,
                           long long row_stride, const int32_t* lengths_dev, float* logmel_dev,
                           float* min_dev, float* max_dev, void* stream);

/* waveform -> uint8 dMel codes, fused.  code = clamp(floor((x - lo_c) * scale_c), 0, K-1)
 * with scale_c = K / (hi_c - lo_c) supplied by the caller (float32, n_mels each).
 * Frames at or past lengths[b] / hop get code 0.  Optional outputs (NULL to skip):
 *   logmel_dev     : (B, n_mels, T) the pre-quantisation values
 *   near_edge_dev  : one uint64, incremented by the number of valid values closer
 *                    than edge_eps (log-mel units) to an interior bin edge.
 * Stands where the reference calls encode_mel_transform then quantises
 * (reference models/codec_lit_modules.py:486-513). */
int dmel_encode_u8(dmel_plan* plan, const float* wav_dev, long long n_rows, long long n_samples,
                   long long row_stride, const int32_t* lengths_dev,
                   const float* lo_dev, const float* scale_dev, int n_bins,
                   uint8_t* codes_dev, float* logmel_dev,
                   unsigned long long* near_edge_dev, float edge_eps, void* stream);

/* The quantiser's forward in one launch: codes as dmel_encode_u8, and for every code its bin centre
 * mel_hat = lo_c + (code + 0.5) * step_c (step_c = (hi_c - lo_c) / K supplied by the caller, float32;
 * multiply and add rounded separately, i.e. bit-identical to dmel_dequantize_f32 on those codes), 0 at frames
 * at or past lengths[b] / hop.  Mirrors forward() of the reference's quantiser module, which returns the
 * quantised features together with the indices (models/modules/dowmsample_fsq.py:86-122). */
int dmel_encode_decode_u8(dmel_plan* plan, const float* wav_dev, long long n_rows, long long n_samples,
                          long long row_stride, const int32_t* lengths_dev,
                          const float* lo_dev, const float* scale_dev, const float* step_dev, int n_bins,
                          uint8_t* codes_dev, float* mel_hat_dev, void* stream);

/* dmel_encode_u8 for int16 PCM waveforms (the format audio is stored and shipped in): value = sample / 32768,
 * the scale is folded into the window taps, so the codes are bit-identical to dmel_encode_u8 on the
 * float32 tensor x / 32768.  Half the bytes per sample over PCIe and from HBM.  Where the reference
 * converts decoded PCM to float32 on the host (dataset/lhotse_tts_dataset.py:29-33, then
 * models/codec_lit_modules.py:487 `audios.float()`), this takes the PCM directly.
 * DMEL_ERR_UNSUPPORTED if the geometry does not fit the register-lean kernel variant. */
int dmel_encode_pcm16_u8(dmel_plan* plan, const int16_t* wav_dev, long long n_rows, long long n_samples,
                         long long row_stride, const int32_t* lengths_dev,
                         const float* lo_dev, const float* scale_dev, int n_bins,
                         uint8_t* codes_dev, void* stream);

/* Windowed form for streaming: writes frames [t_begin, t_begin + t_count) only (t_count < 0 = to the
 * end), for rows of which only samples [src_base, n_samples) are resident, at wav_dev[row][0 ...].
 * n_samples is the row length the reflect padding refers to: while a stream is open pass the number
 * of samples received so far and ask only for frames that end inside them; at end of stream pass the
 * final length and the remaining frames get the reference's right-edge reflection.  The buffer must
 * reach back to the first tap of frame t_begin (sample t_begin*hop - (n_fft-hop)/2, or 0).
 * Output tensors are (n_rows, n_mels, t_count).  codes_dev or logmel_dev may be NULL (not both).
 * The reference has no streaming mode; frame for frame the result equals dmel_encode_u8 on the
 * whole row (reference utils/spectrogram.py:41-81 applied to the complete waveform). */
int dmel_encode_frames_u8(dmel_plan* plan, const float* wav_dev, long long n_rows, long long row_stride,
                          long long src_base, long long n_samples, long long t_begin, long long t_count,
                          const float* lo_dev, const float* scale_dev, int n_bins,
                          uint8_t* codes_dev, float* logmel_dev, void* stream);

/* Streaming encoder (BASELINE configs[3], 80 ms chunks): a lock-step batch of n_streams audio
 * streams with a device-side history buffer of capacity_samples per stream.  The reference has no
 * streaming mode; the contract is that the codes of all pushes and the final flush, concatenated
 * along T, equal dmel_encode_u8 on the whole waveform bit for bit (reference
 * utils/spectrogram.py:41-81 applied to the complete row).
 *   dmel_stream_pending : frames the next push of n_incoming samples (at_end = 0) or the flush
 *                         (at_end = 1, n_incoming = 0) will emit; size the codes buffer with it.
 *   dmel_stream_push    : appends chunk ((n_streams, chunk_stride >= n) float32, device OR host
 *                         memory) and encodes every frame whose last tap has arrived into
 *                         codes_dev (n_streams, n_mels, codes_frames); codes_frames must equal
 *                         dmel_stream_pending.  One copy + one kernel launch on `stream`.
 *   dmel_stream_flush   : end of stream: the remaining frames (right-edge reflection), then reset. */
int dmel_stream_create(dmel_plan* plan, int n_streams, long long capacity_samples, dmel_stream** out);
void dmel_stream_destroy(dmel_stream* s);
int dmel_stream_reset(dmel_stream* s);
long long dmel_stream_pending(const dmel_stream* s, long long n_incoming, int at_end);
int dmel_stream_push(dmel_stream* s, const float* chunk, long long n, long long chunk_stride,
                     const float* lo_dev, const float* scale_dev, int n_bins,
                     uint8_t* codes_dev, long long codes_frames, long long* n_frames_out, void* stream);
int dmel_stream_flush(dmel_stream* s, const float* lo_dev, const float* scale_dev, int n_bins,
                      uint8_t* codes_dev, long long codes_frames, long long* n_frames_out, void* stream);

/* Zero-copy form of a push for latency-bound callers (80 ms chunks): bind the quantiser and the CUDA stream once,
 * then per chunk ask where the next n samples of every stream go (a window of the library's history buffer:
 * where_dev[stream * row_stride + i]), have the producer write them there, and commit - ONE kernel launch, no
 * chunk copy, four arguments.  Same codes as dmel_stream_push. */
int dmel_stream_bind(dmel_stream* s, const float* lo_dev, const float* scale_dev, int n_bins, void* stream);
int dmel_stream_input(dmel_stream* s, long long n, float** where_dev, long long* row_stride);
int dmel_stream_commit(dmel_stream* s, long long n, uint8_t* codes_dev, long long codes_frames, long long* n_frames_out);

/* Same as dmel_encode_u8 with HOST buffers: chunks rows through pinned staging,
 * overlapping H2D, kernel and D2H on internal streams; returns when codes_host
 * is complete.  lengths_host may be NULL; lo/scale are host arrays too. */
int dmel_encode_host_u8(dmel_plan* plan, const float* wav_host, long long n_rows, long long n_samples,
                        long long row_stride, const int32_t* lengths_host,
                        const float* lo_host, const float* scale_host, int n_bins,
                        uint8_t* codes_host);

/* dmel_encode_host_u8 for int16 PCM host buffers. */
int dmel_encode_host_pcm16_u8(dmel_plan* plan, const int16_t* wav_host, long long n_rows, long long n_samples,
                              long long row_stride, const int32_t* lengths_host,
                              const float* lo_host, const float* scale_host, int n_bins,
                              uint8_t* codes_host);

/* ---------------------------------------------------------------------------------------------
 * The general entry point: every input layout and every output of the fused kernel in one call.
 * The functions above are this call with a few fields filled in.
 *
 * Input layouts
 *   padded   wav_dev is (n_rows, row_stride >= n_samples): what the reference's collate yields
 *            (reference dataset/lhotse_tts_dataset.py:46-65, right zero-padded rows).
 *   ragged   offsets_dev != NULL: row b is wav_dev[offsets[b] .. offsets[b+1]) (n_rows + 1 int64 entries,
 *            device memory), n_samples = the longest row; no padding is stored or read.  Offsets that are
 *            multiples of 4 samples (8 for int16) keep the bulk-copy fast path; to align them, leave slack
 *            between rows and pass lengths_dev as well: row b is then the first lengths[b] samples at offsets[b].
 *   own_length = 1 (implied by ragged): each row is an utterance of its own length (lengths_dev[b], or the
 *            offsets' difference) and is transformed as if the reference ran on it ALONE: the reflect padding
 *            happens at the utterance's own end, it has lengths[b] / hop frames, everything past them is written
 *            as 0.  With own_length = 0 and lengths_dev the reference's behaviour on the padded batch is kept
 *            (reflection at the end of the padded row, utils/spectrogram.py:58-62 sees the collated tensor).
 *            min_row_samples: the caller's lower bound on the row lengths; must exceed the reflect pad
 *            (the reference's F.pad raises otherwise) - the library cannot read device-side lengths.
 *   row_gain_dev  per-row gain applied to the samples (float32 input only): per-utterance peak normalisation,
 *            reference dataset/lhotse_tts_dataset.py:29-33; fill it with dmel_row_peak_gain_f32.
 * Outputs: any subset of codes (+ bin centres), log-mel (float32 / bfloat16, optionally masked), running
 * min / max; all (n_rows, n_mels, T) with T = dmel_plan_num_frames(plan, n_samples).  Statistics and codes
 * cannot be asked for in one call (calibration precedes quantisation). */
typedef struct dmel_io {
  size_t struct_size;            /* sizeof(dmel_io): lets the library reject a caller built against another layout */
  const void* wav_dev;           /* float32, or int16 PCM with wav_is_pcm16 = 1 */
  int wav_is_pcm16;
  long long n_rows, n_samples, row_stride;
  const long long* offsets_dev;  /* ragged layout, or NULL */
  const int32_t* lengths_dev;    /* valid samples per row, or NULL */
  int own_length;
  long long min_row_samples;     /* only read with own_length / ragged */
  const float* row_gain_dev;     /* or NULL */
  /* quantiser (needed for codes_dev) */
  const float* lo_dev;
  const float* scale_dev;
  const float* step_dev;         /* needed for mel_hat_dev */
  int n_bins;
  /* outputs, NULL to skip */
  uint8_t* codes_dev;
  float* mel_hat_dev;            /* bin centre of every code (needs codes_dev) */
  void* logmel_dev;              /* float32, or bfloat16 with logmel_is_bf16 = 1 */
  int logmel_is_bf16;
  int mask_invalid;              /* log-mel of frames at or past lengths[b] / hop written as 0 */
  float* min_dev;                /* running per-channel min / max, updated in place */
  float* max_dev;
} dmel_io;

int dmel_run(dmel_plan* plan, const dmel_io* io, void* stream);

/* gain_dev[b] = target_peak / max|x| over the valid samples of row b (a peak below FLT_MIN counts as 1):
 * the gain of `librosa.util.normalize(audio) * 0.95` (reference dataset/lhotse_tts_dataset.py:32) with
 * target_peak = 0.95.  One HBM-bound pass over the waveform (4 bytes per sample) and a tiny second launch;
 * padded (offsets_dev NULL) or ragged layout as in dmel_io; lengths_dev may be NULL. */
int dmel_row_peak_gain_f32(const float* wav_dev, long long n_rows, long long n_samples, long long row_stride,
                           const long long* offsets_dev, const int32_t* lengths_dev, float target_peak,
                           float* gain_dev, void* stream);

/* Stand-alone quantiser stages on an existing (B, n_mels, T) log-mel tensor. */
int dmel_quantize_u8(const float* logmel_dev, long long n_rows, int n_mels, long long n_frames,
                     const float* lo_dev, const float* scale_dev, int n_bins,
                     uint8_t* codes_dev, void* stream);
/* The same with a valid-frame count per batch row (n_valid_dev: n_rows int32, or NULL = dmel_quantize_u8): frames at
 * or past it get code 0, which is what the fused encode writes there, and their log-mel is not read.  Pass 2 of the
 * calibrate-then-encode job over right-padded batches (reference collate: dataset/lhotse_tts_dataset.py:46-65; the
 * valid-frame rule: models/codec_lit_modules.py:176-177). */
int dmel_quantize_masked_u8(const float* logmel_dev, long long n_rows, int n_mels, long long n_frames,
                            const int32_t* n_valid_dev, const float* lo_dev, const float* scale_dev, int n_bins,
                            uint8_t* codes_dev, void* stream);

/* codes -> bin-centre log-mel.  table_dev is (n_mels, n_bins) float32:
 * table[c][k] = lo_c + (k + 0.5) * (hi_c - lo_c) / K.  Codes >= n_bins read as n_bins-1. */
int dmel_dequantize_f32(const uint8_t* codes_dev, long long n_rows, int n_mels, long long n_frames,
                        const float* table_dev, int n_bins, float* logmel_dev, void* stream);

/* What the quantiser derives from its statistics, one launch; each output may be NULL:
 *   scale_dev[c] = hi - lo > 0 ? n_bins / (hi - lo) : 0     step_dev[c] = (hi - lo) / n_bins     (float32, IEEE division)
 *   *ready_dev   = 1 iff lo[c] <= hi[c] for every channel: the statistics have seen at least one frame. */
int dmel_quantizer_derive_f32(const float* lo_dev, const float* hi_dev, int n_mels, int n_bins, float* scale_dev,
                              float* step_dev, int32_t* ready_dev, void* stream);

/* Running min/max over valid frames of an existing log-mel tensor
 * (n_valid_dev: frames per row, or NULL). Updates min_dev / max_dev in place. */
int dmel_tensor_minmax_f32(const float* logmel_dev, long long n_rows, int n_mels, long long n_frames,
                           const int32_t* n_valid_dev, float* min_dev, float* max_dev, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Next to the path: the elementwise core of the reference's LEARNED quantiser.
 * Finite scalar quantisation of the per-group latents to codes and indices, as GroupedResidualFSQ(levels,
 * num_quantizers = 1, groups = G) of vector_quantize_pytorch computes them between its learned project_in and
 * project_out (reference models/modules/dowmsample_fsq.py:39-44, :95, :130-137), and the language model's id_shift
 * (models/modules/lm_process_input.py:301-313) fused into the same pass.
 *   zp_dev      (B, T, G, D) float32, D = n_levels <= 8
 *   codes_dev   (B, T, G, D) float32 in [-1, 1], or NULL
 *   indices_dev (B, G, T) int64 in [0, prod(levels)): the layout DownsampleFiniteScalarQuantize.encode returns; or NULL
 *   lm_ids_dev  (B, T, G) int64 = index + g * codebook_size, or NULL
 * tanh is evaluated as 1 - 2 / (1 + e^{2x}) on the hardware's ex2 / rcp approximations (absolute error below 2e-7, the size
 * of tanhf's own last bit): a level can differ from a float32 tanhf implementation only where tanh(z + shift) * half_l - offset
 * lies within ~1e-6 of a rounding boundary (x.5), the same caveat any two float32 tanh implementations carry. */
int dmel_fsq_encode(const float* zp_dev, long long n_rows, long long n_steps, int n_groups, const int* levels,
                    int n_levels, float* codes_dev, long long* indices_dev, long long* lm_ids_dev, int codebook_size,
                    void* stream);
/* indices (B, G, T) int64 -> codes (B, T, G, D): FSQ.indices_to_codes (the input of the learned project_out). */
int dmel_fsq_decode(const long long* indices_dev, long long n_rows, long long n_steps, int n_groups, const int* levels,
                    int n_levels, float* codes_dev, void* stream);

/* BigVGAN's anti-aliased Snake activation, one pass: 2x upsample (12-tap FIR, replicate padding) -> x + sin^2(a x) / (b + 1e-9)
 * -> 2x downsample (12-tap FIR, replicate padding).  Replaces Activation1d.forward of the reference
 * (models/modules/bigvgan/alias_free_activation/torch/act.py:24-29, resample.py:10-58, activations.py:101-111) and its fused
 * kernel for sm_70 / sm_80 (.../cuda/anti_alias_activation_cuda.cu:44-179), whose interface this mirrors: filters as the
 * modules hold them (12 taps each, host arrays), alpha and beta per channel in LOG scale (device arrays; pass the same array
 * twice for the one-parameter Snake).
 *   x_dev, y_dev : (B, C, T) float32 */
int dmel_antialias_snake_f32(const float* x_dev, long long n_rows, int n_channels, long long n_steps, const float* up_taps_host,
                             const float* down_taps_host, const float* log_alpha_dev, const float* log_beta_dev, float* y_dev,
                             void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DMEL_B200_H */
