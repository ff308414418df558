/*
 * rnnspeech_b200 -- C ABI of the B200-native acoustic-model hot path
 * (features -> LSTM stack fwd/bwd -> CTC loss/grad/decode -> clip + Adam).
 *
 * The reference (domerin0/rnn-speech) is pure Python on TensorFlow-1 + librosa
 * and has NO FFI of its own for this path: the "interface each entry point
 * replaces" is therefore the Python method / TF op the reference calls, cited
 * as file:line relative to the reference root.  INTEGRATION.md shows the
 * ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every function returns 0 on success, a negative rs_status on error;
 *     rs_last_error() returns a thread-local message for the last failure;
 *   - all `*_d` / data pointers are DEVICE pointers owned by the caller
 *     (e.g. torch tensors' data_ptr()); the library never allocates device
 *     memory: scratch is passed in, sized by the matching *_workspace_bytes;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it
 *     and nothing synchronises the host unless stated;
 *   - thread-compatible: no internal threads, no global mutable state beyond
 *     the thread-local error string;  one process per GPU rank;
 *   - no CPU fallback: every entry point needs an sm_100 device.
 */
#ifndef RNNSPEECH_B200_H_
#define RNNSPEECH_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  RS_OK = 0,
  RS_ERR_INVALID = -1,      /* bad argument (reference: ValueError / InvalidArgumentError) */
  RS_ERR_CUDA = -2,         /* CUDA runtime error, message holds cudaGetErrorString */
  RS_ERR_UNSUPPORTED = -3,  /* shape outside what the kernels support */
  RS_ERR_WORKSPACE = -4     /* workspace / reserve buffer too small */
} rs_status;

int rs_version(void);
const char* rs_last_error(void);
/* Number of SMs of the current device (grid sizing is a multiple of this). */
int rs_sm_count(void);
/* Number of kernels this library has launched in this process so far. */
uint64_t rs_launch_count(void);
/* CRC-32C of a host buffer, continuing from `crc` (0 to start): TensorFlow's tensor-bundle checksum, for the
 * checkpoint writer (rnn-speech_b200/tf_checkpoint.py). */
uint32_t rs_crc32c(const void* data, size_t n, uint32_t crc);

/* ------------------------------------------------------------------------
 * (a) Feature extraction.  Replaces AudioProcessor.process_signal(sig, sr)
 *     -> _extract_fbank / _extract_mfcc, util/audioprocessor.py:52-161, for a
 *     whole batch of utterances in one call, and the padded-batch contract of
 *     AcousticModel.build_dataset (models/AcousticModel.py:809-827) plus the
 *     time-major transpose of create_training_rnn (:147-152).
 *
 *   pcm_d      float32 PCM, utterances concatenated
 *   offsets_d  int64[B+1] sample offsets into pcm_d (utterance b = [off[b], off[b+1]))
 *   max_samples  host copy of the longest utterance length (grid sizing)
 *   sr         sample rate; frame = round(.025 sr), hop = round(.01 sr)
 *   Tmax       max_input_seq_length: frames >= Tmax are dropped, rows
 *              [nframes, Tmax) are zero-filled
 *   delta_mode RS_DELTA_INTERP (librosa >= 0.6.1, savgol 'interp') or
 *              RS_DELTA_EDGE (librosa <= 0.6.0, edge replicate, /20)
 *   time_major 0: out is [B, Tmax, F]   1: out is [Tmax, B, F]
 *   out_d      float32 features, F = 120 (fbank) / n_mfcc (mfcc)
 *   nframes_d  int32[B]: PRE-truncation frame count (the reference's
 *              `length` return value, util/audioprocessor.py:156-161)
 * ------------------------------------------------------------------------ */
#define RS_DELTA_INTERP 0
#define RS_DELTA_EDGE 1
#define RS_FBANK_DIM 120

size_t rs_fbank_workspace_bytes(int B, int64_t max_samples, int sr);
/* ceil(|n - round(.025 sr)| / round(.01 sr)): util/audioprocessor.py:92 */
int64_t rs_fbank_num_frames(int64_t n_samples, int sr);
/* host-only: the mel weights [40*257] / Hamming window [512] the kernels use (CPU tests) */
int rs_fbank_tables_host(int sr, float* melw_out, float* window_out, int* frame_length, int* frame_step);
int rs_fbank_forward(const float* pcm_d, const int64_t* offsets_d, int B, int64_t max_samples,
                     int sr, int Tmax, int delta_mode, int time_major,
                     float* out_d, int32_t* nframes_d, void* ws_d, size_t ws_bytes, void* stream);

size_t rs_mfcc_workspace_bytes(int B, int64_t max_samples, int sr);
/* 1 + n / round(.01 sr)  (librosa stft, center=True) */
int64_t rs_mfcc_num_frames(int64_t n_samples, int sr);
int rs_mfcc_forward(const float* pcm_d, const int64_t* offsets_d, int B, int64_t max_samples,
                    int sr, int Tmax, int n_mfcc, int time_major,
                    float* out_d, int32_t* nframes_d, void* ws_d, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------
 * (a0) Audio front door.  Replaces what AudioProcessor.process_audio_file reaches
 *     through librosa.load(file_name, mono=True) (util/audioprocessor.py:49) AFTER
 *     the container has been decoded to PCM: int16 -> float32 (x * 2^-15), mean over
 *     channels, and resampling to 22 050 Hz with resampy's 'kaiser_best' band-limited
 *     sinc interpolation (64 zero crossings, 512 table samples per crossing, Kaiser
 *     beta 14.7696..., roll-off 0.94759...), padded to ceil(n * ratio) samples like
 *     librosa.util.fix_length.  The output feeds rs_fbank_forward / rs_mfcc_forward
 *     on the device without a host round trip.
 *
 *   pcm_d       RS_PCM_F32: float32 samples, RS_PCM_S16: int16 samples; `channels` (1..8) interleaved
 *   offsets_d   int64[B+1] FRAME offsets (samples per channel) of the utterances in pcm_d
 *   out_offsets_d int64[B+1]: out_offsets[b+1] - out_offsets[b] = rs_resample_num_samples(n_b, ...)
 *   max_out_samples  host copy of the longest output length (grid sizing)
 * ------------------------------------------------------------------------ */
#define RS_PCM_F32 0
#define RS_PCM_S16 1

size_t rs_resample_workspace_bytes(int B, int64_t max_out_samples);
/* ceil(n * sr_out / sr_in): the length librosa.load returns */
int64_t rs_resample_num_samples(int64_t n_samples, int sr_in, int sr_out);
/* host-only: the one-sided interpolation filter [64*512+1] the kernels build (CPU tests) */
int rs_resample_filter_host(double* win_out, int* num_table);
int rs_resample_forward(const void* pcm_d, int pcm_format, int channels, const int64_t* offsets_d, int B,
                        int64_t max_out_samples, int sr_in, int sr_out, float* out_d,
                        const int64_t* out_offsets_d, void* ws_d, size_t ws_bytes, void* stream);
/* Host-only FLAC decoder (the container librosa.load opens through audioread / soundfile for the LibriSpeech
 * files the reference trains on, util/audioprocessor.py:49, util/dataprocessor.py:207-243).  data / nbytes: a whole
 * .flac file in HOST memory;  out: interleaved int32 samples [frames * channels] in HOST memory, or NULL to read the
 * stream parameters only (out_capacity < 0 with out == NULL: walk every frame and count instead of trusting STREAMINFO);  md5: the 16-byte signature of the unencoded audio from STREAMINFO (zeros if unset).
 * Every frame's CRC-8 / CRC-16 is checked; a mismatch is RS_ERR_INVALID. */
int rs_flac_decode_host(const uint8_t* data, size_t nbytes, int32_t* out, int64_t out_capacity,
                        int* sample_rate, int* channels, int* bits_per_sample, int64_t* total_frames,
                        uint8_t* md5);
/* int16 interleaved -> float32 mono without a rate change (files already at the target rate) */
int rs_pcm16_to_f32(const int16_t* pcm_d, int64_t frames, int channels, float* out_d, void* stream);
/* float32 interleaved -> float32 mono (mean over channels), likewise */
int rs_pcm_f32_to_mono(const float* pcm_d, int64_t frames, int channels, float* out_d, void* stream);

/* ------------------------------------------------------------------------
 * (b) Acoustic model: input dense -> L x LSTM (TF BasicLSTMCell semantics,
 *     gate order i,j,f,o, forget_bias 1.0, per-step in/out dropout,
 *     dynamic_rnn sequence-length masking, persistent state in/out) ->
 *     output dense.  Replaces AcousticModel._build_base_rnn
 *     (models/AcousticModel.py:189-317) and the BPTT half of
 *     _add_training_on_rnn (:386-401).
 *
 *   params_d   float32 flat parameter buffer, TF variable layouts in the
 *              reference's checkpoint order (models/AcousticModel.py:515-527):
 *              input_w[F,H] input_b[H] {kernel_l[2H,4H] bias_l[4H]}xL
 *              output_w[H,C] output_b[C]
 *   x_d        float32 [T, B, F] time-major features
 *   len_d      int32 [B] valid frames per item (0 allowed)
 *   state_in_d / state_out_d   float32 [L, 2, B, H] (c then h per layer);
 *              state_in_d may be NULL (zeros); state_out_d may be NULL or
 *              alias state_in_d
 *   keep_in / keep_out  dropout keep probabilities (1.0 = identity)
 *   seed       dropout seed for this call (masks = f(seed, layer, t, b, h))
 *   logits_d   float32 [T, B, C]
 *   reserve_d  activations kept for backward (NULL => inference, nothing saved)
 *   grads_d    float32 flat buffer, same layout as params_d; backward
 *              ACCUMULATES (+=) into it (the reference's accumulate_gradients_op,
 *              models/AcousticModel.py:392-401)
 * ------------------------------------------------------------------------ */
typedef struct rs_am rs_am;

int rs_am_create(rs_am** out, int num_layers, int hidden_size, int input_dim, int num_labels,
                 int batch_size, int max_T);
void rs_am_destroy(rs_am* am);
int64_t rs_am_param_count(const rs_am* am);
/* offset (in floats) of a named variable inside the flat buffer: which = 0 input_w,
 * 1 input_b, 2 kernel (layer), 3 bias (layer), 4 output_w, 5 output_b */
int64_t rs_am_param_offset(const rs_am* am, int which, int layer);
/* 1 if this shape runs on the tcgen05 kernels (hidden_size % 64 == 0, batch <= 64, weights
 * fit in shared memory), 0 if it runs on the fp32 FFMA kernels. */
int rs_am_uses_tensor_cores(const rs_am* am);
/* Batch normalisation of the stack's input over the batch axis, no scale / offset, eps 1e-3 (the reference's
 * `normalization` constructor argument, models/AcousticModel.py:253-259; config.ini batch_normalization).  Call
 * before sizing the reserve / workspace. */
int rs_am_set_normalization(rs_am* am, int enable);
size_t rs_am_reserve_bytes(const rs_am* am);
size_t rs_am_workspace_bytes(const rs_am* am);
int rs_am_forward(rs_am* am, const float* params_d, const float* x_d, const int32_t* len_d, int T,
                  const float* state_in_d, float* state_out_d,
                  float keep_in, float keep_out, uint64_t seed,
                  float* logits_d, void* reserve_d, void* ws_d, size_t ws_bytes, void* stream);
/* Weight planes (the bf16 hi/lo images of the parameters the kernels read) are kept in the caller's workspace.  By
 * default they are re-packed on every rs_am_forward / rs_am_backward.  A caller that sets a non-zero version vouches
 * that params_d only changes when the version does and that nothing else writes the workspace between calls; the
 * planes are then re-packed once per version (the reference's variables likewise change only in train_step_op,
 * models/AcousticModel.py:404-406). */
int rs_am_set_params_version(rs_am* am, uint64_t version);
/* Measurement hooks: when enabled, CUDA events are recorded on the launching stream around
 * each recurrent kernel launch (one per layer, or one per time chunk of a layer in the
 * pipelined schedule).  rs_am_recurrent_ms returns the summed duration of a layer's launches
 * in the last call; rs_am_recurrent_trace writes (start, stop) of each launch in ms after the
 * top of that call and returns how many launches it wrote.  backward = 0 | 1. */
int rs_am_enable_timing(rs_am* am, int enable);
int rs_am_recurrent_ms(rs_am* am, int backward, int layer, float* ms);
int rs_am_recurrent_trace(rs_am* am, int backward, int layer, float* start_stop_ms, int max_launches);
/* Debug: %globaltimer stamps [T][8] (uint64) of CTA 0 of the layer-0 recurrent kernels. */
int rs_am_set_debug_timeline(rs_am* am, void* fwd_d, void* bwd_d);
/* keep_in / keep_out / seed must repeat the values given to the matching forward
 * (the dropout masks are recomputed, not stored).  The reserve is consumed:
 * one backward per forward. */
int rs_am_backward(rs_am* am, const float* params_d, const float* x_d, const int32_t* len_d, int T,
                   float keep_in, float keep_out, uint64_t seed,
                   const float* dlogits_d, void* reserve_d, float* grads_d,
                   void* ws_d, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------
 * (c) CTC.  Replaces tf.nn.ctc_loss(sparse_labels, logits, seq_len,
 *     ignore_longer_outputs_than_inputs=True) (models/AcousticModel.py:357)
 *     including its gradient, and provides the greedy decode the north star
 *     names as the prediction parity surface (:312-314).
 *
 *   logits_d   float32 [T, B, C] time-major, softmax applied inside
 *   labels_d   int32, label sequences concatenated; label_offsets_d int32[B+1]
 *   len_d      int32[B]
 *   blank      blank index (reference: C-1 = 79)
 *   beta_skip  RS_CTC_BETA_SOURCE (TF rule, default) / RS_CTC_BETA_DEST
 *   loss_d     float32[B];  grad_d float32 [T,B,C] or NULL (loss only)
 * ------------------------------------------------------------------------ */
#define RS_CTC_BETA_SOURCE 0
#define RS_CTC_BETA_DEST 1

size_t rs_ctc_workspace_bytes(int T, int B, int C, int max_label_len);
int rs_ctc_loss_grad(const float* logits_d, const int32_t* labels_d, const int32_t* label_offsets_d,
                     const int32_t* len_d, int T, int B, int C, int max_label_len, int blank,
                     int beta_skip, float* loss_d, float* grad_d,
                     void* ws_d, size_t ws_bytes, void* stream);
/* out_d int32 [B, T] (decoded ids, row-padded with -1), out_len_d int32 [B] */
int rs_ctc_greedy_decode(const float* logits_d, const int32_t* len_d, int T, int B, int C, int blank,
                         int32_t* out_d, int32_t* out_len_d, void* stream);
/* Beam search decoder.  Replaces tf.nn.ctc_beam_search_decoder(logits, seq_len) -- the reference's `prediction`
 * (models/AcousticModel.py:312-314: beam_width 100, top_paths 1, merge_repeated True, blank = C-1).
 *   normalize   1: log-softmax the scores per frame (TF >= 1.12), 0: subtract the maximum only; the decoded
 *               path is the same, only out_score_d differs
 *   out_d       int32 [B, T] top path, row-padded with -1;  out_len_d int32 [B];  out_score_d float32 [B] or NULL
 *   limits      beam_width <= 128, C <= 128, beam_width * C <= 8184 */
size_t rs_ctc_beam_workspace_bytes(int T, int B);
int rs_ctc_beam_search(const float* logits_d, const int32_t* len_d, int T, int B, int C, int beam_width,
                       int merge_repeated, int normalize, int32_t* out_d, int32_t* out_len_d,
                       float* out_score_d, void* ws_d, size_t ws_bytes, void* stream);

/* Edit distance between decoded and reference label sequences.  Replaces tf.edit_distance(prediction,
 * sparse_labels, normalize=True), the reference's error rate (models/AcousticModel.py:370).
 *   hyp_d [B, hyp_ld] int32 (as written by the decoders), hyp_len_d [B]; truth_d flat int32 + truth_offsets_d [B+1]
 *   dist_d int32 [B];  rate_d float32 [B] = dist / len(truth) (inf: empty truth, non-empty hypothesis) or NULL */
int rs_edit_distance(const int32_t* hyp_d, const int32_t* hyp_len_d, int hyp_ld, const int32_t* truth_d,
                     const int32_t* truth_offsets_d, int B, int max_truth_len, int32_t* dist_d, float* rate_d,
                     void* stream);

/* ------------------------------------------------------------------------
 * Update rule.  Replaces tf.clip_by_global_norm + AdamOptimizer.apply_gradients
 * (models/AcousticModel.py:388,404-406).
 *   rs_sumsq: sumsq_d[0] = sum(g^2)  (float64 accumulator, device)
 *   rs_clip_adam_step: reads sumsq_d[0] on device (no host sync):
 *       g' = g * clip / max(sqrt(sumsq), clip);  TF ApplyAdam with step `step` (1-based)
 * ------------------------------------------------------------------------ */
int rs_sumsq(const float* g_d, int64_t n, double* sumsq_d, void* stream);
int rs_clip_adam_step(float* params_d, const float* grads_d, float* m_d, float* v_d, int64_t n,
                      const double* sumsq_d, float clip, float lr, float beta1, float beta2,
                      float eps, int64_t step, void* stream);

/* ------------------------------------------------------------------------
 * Step-protocol helpers, so that a training step launches no framework kernels.
 *   rs_accumulate_mean: dst_d[0] += (1/n) sum_i v_d[i] / (div_d ? div_d[i] : 1); count_d[0] += 1 when given.
 *       Replaces acc_mean_loss_op / acc_error_rate_op / increase_mini_batch_op
 *       (models/AcousticModel.py:361-383: mean_loss = mean(ctc_loss / seq_len), accumulated over mini-batches).
 *   rs_memset_zero: replaces the accumulator / gradient-accumulator zeroing ops of start_batch (:662-670).
 *   rs_memcpy_h2d_async: ONE asynchronous copy of a staged (pinned) mini-batch to the device
 *       (the feed of models/AcousticModel.py:147-152 / the tf.data prefetch of :819-827).
 * ------------------------------------------------------------------------ */
int rs_accumulate_mean(const float* v_d, const int32_t* div_d, int n, float* dst_d, float* count_d, void* stream);
int rs_memset_zero(void* dst_d, size_t bytes, void* stream);
int rs_memcpy_h2d_async(void* dst_d, const void* src_host, size_t bytes, void* stream);

/* ------------------------------------------------------------------------
 * Data parallelism (not in the reference: SURVEY 8e).  One process per GPU; ONE all-reduce(SUM, fp32) of the flat
 * gradient buffer per optimizer step -- summing over ranks is the reference's gradient accumulation over
 * mini_batch_size mini-batches (models/AcousticModel.py:386-401).  NCCL is bound at run time (dlopen of
 * libnccl.so.2, RS_NCCL_LIB overrides); the library links only cudart.
 *   rs_comm_unique_id: rank 0 creates the 128-byte id and hands it to the other ranks out of band
 *   rs_comm_init:      collective over all ranks, on the current device
 *   rs_allreduce_sum:  in place, enqueued on `stream`
 * ------------------------------------------------------------------------ */
#define RS_COMM_ID_BYTES 128
int rs_comm_unique_id(void* id_out, size_t id_bytes);
int rs_comm_init(void** comm_out, const void* unique_id, size_t id_bytes, int rank, int world);
int rs_allreduce_sum(void* comm, float* buf_d, int64_t n, void* stream);
void rs_comm_destroy(void* comm);

#ifdef RS_DIAG
/* ------------------------------------------------------------------------
 * Diagnostics: NOT part of the product library.  These hooks exist only in librnnspeech_b200_diag.so (the same
 * sources built with -DRS_DIAG; rnn-speech_b200/build.py builds both), which the GPU tests and tests/gpu_diag.py load.
 * Self-test of the tcgen05 / TMEM / descriptor plumbing:
 * D[128,N] = A[128,K] * B[N,K]^T on one CTA; split != 0 uses the bf16x3 split.
 * ------------------------------------------------------------------------ */
int rs_tc_selftest(const float* A_d, const float* B_d, float* D_d, int N, int K, int split, void* stream);
/* tcgen05.mma issue-rate microbenchmark (one CTA, SS operands): out_cycles_d int64[2] =
 * {cycles to issue, cycles until the last MMA completed}. */
int rs_tc_mma_bench(int M, int N, int count, int variant, int nacc, void* out_cycles_d, void* stream);
/* Self-test of the TMEM-resident A operand ("TS" MMA) with the hi/lo stacking of the recurrent
 * kernels: A_d [64,K], B_d [32,K] fp32 -> raw accumulator D_d [128 TMEM lanes][64 columns]
 * (lane 32q+i = A_hi row 16q+i, lane 32q+16+i = A_lo row 16q+i; columns 0..31 = B_hi, 32..63 = B_lo).
 * K multiple of 64, <= 768.  out_cycles_d int64[2] (or NULL) as in rs_tc_mma_bench, for `reps` passes. */
int rs_tc_ts_selftest(const float* A_d, const float* B_d, float* D_d, int K, int variant, int reps,
                      void* out_cycles_d, void* stream);
/* Test hook for the production tcgen05 GEMM: C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) from fp32
 * inputs (split to bf16 planes inside); products = 1 (plain bf16) or 3 (bf16x3).
 * scratch_d: 2*(M*K + N*K)*2 + 64 bytes.  K multiple of 8. */
int rs_gemm_tc_test(const float* A_d, const float* B_d, const float* bias_d, float* C_d, int M, int N, int K,
                    int products, void* scratch_d, size_t scratch_bytes, void* stream);
/* Measurement hook: average ms of `reps` launches of the same GEMM (operands split once); bn = 0 | 128 | 256
 * forces the tile width, max_ctas / tiles_per_cta shape the grid. */
int rs_gemm_tc_bench(const float* A_d, const float* B_d, float* C_d, int M, int N, int K, int products,
                     int bn, int max_ctas, int tiles_per_cta, int accumulate, int reps, void* scratch_d,
                     size_t scratch_bytes, float* ms_out, void* stream);

#endif /* RS_DIAG */

#ifdef __cplusplus
}
#endif
#endif /* RNNSPEECH_B200_H_ */
