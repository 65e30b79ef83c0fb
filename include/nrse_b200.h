/*
 * nrse_b200.h -- C ABI of the B200-native BYOL noisy-view hot path.
 *
 * The reference (sunYtokki/Noise-Robust-Speech-Embedding) is pure Python and has no FFI;
 * its boundary for this path is the Python module surface.  Each entry point
 * below replaces the body of one reference function; the host-side Python mirror in
 * noise-robust-speech-embedding_b200/ binds these with ctypes (INTEGRATION.md shows the stub
 * a maintainer of the reference would add).
 *
 * Conventions
 *   - every function returns 0 (NRSE_OK) or a negative nrse_status; nothing throws;
 *   - pointers are DEVICE pointers owned by the caller unless the name ends in _host;
 *   - no entry point allocates device memory or synchronises the stream;
 *   - `stream` is a cudaStream_t (CUstream), e.g. torch.cuda.current_stream().cuda_stream;
 *   - re-entrant per stream; the only global state is the lazily resolved driver entry point for TMA
 *     descriptor encoding, one-time kernel attributes and the tuning knobs (nrse_*_set_*).
 *   - sm_100a only.  There is no CPU or other-arch fallback.
 */
#ifndef NRSE_B200_H_
#define NRSE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* nrse_stream_t;

typedef enum {
  NRSE_OK = 0,
  NRSE_ERR_INVALID_ARG = -1,
  NRSE_ERR_UNSUPPORTED = -2,  /* shape outside what the kernels implement */
  NRSE_ERR_CUDA = -3,         /* a CUDA runtime/driver call failed: see nrse_last_cuda_error() */
  NRSE_ERR_WORKSPACE = -4,    /* caller-provided workspace too small */
  NRSE_ERR_NO_DEVICE = -5     /* not an sm_100 device */
} nrse_status;

#define NRSE_DTYPE_F32 0
#define NRSE_DTYPE_BF16 1

#define NRSE_NORM_LAYER 0 /* wavlm-large: LayerNorm over channels on every conv layer */
#define NRSE_NORM_GROUP 1 /* wavlm-base(-plus): GroupNorm(512,512) on layer 0 only     */

#define NRSE_FRONTEND_LAYERS 7
#define NRSE_FRONTEND_CHANNELS 512

int nrse_version(void);
const char* nrse_strerror(int status);
/* cudaError_t (as int) of the last failing CUDA call made by this library on the calling thread. */
int nrse_last_cuda_error(void);
/* 0 if the current device is compute capability 10.x, else NRSE_ERR_NO_DEVICE / NRSE_ERR_CUDA. */
int nrse_check_device(void);
/* 1 if this library was built with -DNRSE_EXPERIMENTS (scripts-only timing build whose NRSE_EXPERIMENT environment
 * hooks produce wrong results), 0 for the product build.  smoke() and bench.py refuse a library that returns 1. */
int nrse_experiments_build(void);

/* ---------------------------------------------------------------------------------------------
 * Fused tensor health check: one launch over n_tensors <= NRSE_CHECK_MAX_TENSORS fp32 tensors, one 64-byte record
 * per tensor.  Replaces check_audio_tensor, ref:src/utils/debugging_utils.py:4-30 (called four times per step,
 * ref:train_byol.py:52-59; >= 4 passes and >= 4 host synchronisations per tensor in the reference).
 *   tensors_host / numel_host  HOST arrays of device pointers (4-byte aligned) and element counts
 *   out     DEVICE [n_tensors] records; flags bit 0 = NaN present, bit 1 = Inf present, bit 2 = sum|x| < min_threshold,
 *           bit 3 = max|x| > max_threshold -- the reference's four tests, which it reports in this order;
 *           abs_max / max / min / abs_sum / sum / sumsq / numel are what its DEBUG statistics are derived from
 *           (mean = sum / numel, unbiased std from sumsq).  One cudaMemsetAsync + one kernel on `stream`.
 * ------------------------------------------------------------------------------------------- */
#define NRSE_CHECK_MAX_TENSORS 8
typedef struct {
  int32_t flags;
  float abs_max;
  float max;
  float min;
  double abs_sum;
  double sum;
  double sumsq;
  int64_t numel;
  int64_t reserved;
  int64_t reserved2;
} nrse_tensor_check;
int nrse_check_tensors_f32(const float* const* tensors_host, const int64_t* numel_host, int n_tensors,
                           float max_threshold, float min_threshold, nrse_tensor_check* out, nrse_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Fused SNR mix + peak normalisation + z-normalisation.
 * Replaces, per utterance row b:
 *   add_noise_to_speech          ref:src/data/augment.py:4-66
 *   peak normalisation           ref:src/data/noisy_speech_dataset.py:88-116      (peak_norm=1)
 *   HF zero_mean_unit_var_norm   hf:models/wav2vec2/feature_extraction_wav2vec2.py:95,
 *                                called at ref:src/data/noisy_speech_dataset.py:120-129 and
 *                                ref:src/data/emotion_dataset.py:198-203
 * peak_norm=1 (BYOL pre-training): clean_out and noisy_out are both written.
 * peak_norm=0 (emotion fine-tune, ref:src/data/emotion_dataset.py:177-203): only noisy_out is
 *   written (clean_out may be NULL); a failed mix keeps the clean waveform (:193-194).
 * peak_norm=2: add_noise_to_speech alone -- noisy_out receives the un-normalised mix speech + scale*noise
 *   (rows with status != 0 carry the clean waveform; the Python wrapper returns None for them).
 * clean [B,L], noise [B,L_noise] (truncated if longer, tiled if shorter, augment.py:16-21),
 * snr_idx [B] indexes snr_db_table_host[n_snr] (dB; HOST pointer, n_snr <= 32).
 * status[b] = 0 or the number of the reference's `return None`/`continue` exit that row b would
 *   have taken (1..14, see nrse_mix_status_name); in BYOL mode such rows are zero-filled.
 * ------------------------------------------------------------------------------------------- */
int nrse_mix_normalize_f32(const float* clean, const float* noise, const int32_t* snr_idx,
                           const double* snr_db_table_host, int n_snr,
                           float* clean_out, float* noisy_out, int32_t* status,
                           int B, int L, int L_noise, int peak_norm, nrse_stream_t stream);
/* Device-side retry ("try another noise file, draw another SNR", ref:src/data/noisy_speech_dataset.py:58-84): same
 * operation, but only rows whose status[b] != 0 on entry are processed -- with the noise AND the SNR index of row
 * (b + noise_row_shift) % B, i.e. an independent re-draw of both, as the reference's next attempt makes -- and get their
 * outputs and status rewritten; all other rows are left untouched.  snr_idx itself is never modified;
 * snr_idx_used (nullable, [B]) receives, for every row processed by THIS call, the table index it was mixed at (seed
 * it with a copy of snr_idx), so that the batch's "snr" labels follow the re-draw.  The CTAs of good rows exit at
 * once, so a retry launch on a healthy batch costs a kernel launch and no memory traffic, and the host never has to
 * read the status to decide whether to retry. */
int nrse_mix_normalize_retry_f32(const float* clean, const float* noise, const int32_t* snr_idx,
                                 const double* snr_db_table_host, int n_snr,
                                 float* clean_out, float* noisy_out, int32_t* status, int32_t* snr_idx_used,
                                 int B, int L, int L_noise, int peak_norm, int noise_row_shift,
                                 nrse_stream_t stream);
/* After the retries: every row whose status is still != 0 takes over clean_out / noisy_out (and snr_idx_used) of the
 * nearest following good row of the batch -- the reference moves on to the next item when one is unusable
 * (ref:src/data/noisy_speech_dataset.py:60-66) and never emits a zero waveform.  status is left as it is (it keeps
 * reporting why the row failed).  Good rows' CTAs exit at once: no data moves on a healthy batch, no host sync ever.
 * clean_out and snr_idx_used are nullable. */
int nrse_mix_substitute_rows_f32(float* clean_out, float* noisy_out, const int32_t* status, int32_t* snr_idx_used, int B,
                                 int L, nrse_stream_t stream);
/* The whole attempt loop of NoiseRobustSpeechDataset.__getitem__ (ref:src/data/noisy_speech_dataset.py:55-149) for a batch,
 * in one host call and THREE launches whatever max_attempts is:
 *   1. nrse_mix_normalize_f32 on every row, recording the SNR index used in snr_idx_used [B] (required);
 *   2. (max_attempts > 1, B > 1, peak_norm = 1 -- the emotion path never retries, ref:src/data/emotion_dataset.py:190-194)
 *      one retry launch: a row that was rejected is redone INSIDE the launch with the noise
 *      and SNR draw of rows b + 1, b + 2, ... (mod B) until it passes or max_attempts - 1 further attempts are used up --
 *      the same donors, in the same order, as max_attempts - 1 calls of nrse_mix_normalize_retry_f32 with
 *      noise_row_shift = 1, 2, ...; the CTAs of good rows exit at once;
 *   3. one finishing launch: substitute_bad_rows != 0 copies the nearest following good row over every row that is still
 *      rejected (nrse_mix_substitute_rows_f32); snr_labels_out (nullable, [B] int64) receives
 *      snr_label_table[snr_idx_used[b]] (snr_label_table: DEVICE [n_snr] int64, the "snr" entry of the reference's item
 *      dict, :140-144); n_rejected (nullable, DEVICE int32 scalar) receives the number of rows with status != 0 -- what
 *      the host logs, copied back asynchronously by the caller.
 * The host never reads the status to decide anything; replaces 6 launches + 6 small torch kernels per batch of the
 * earlier retry-per-launch form. */
int nrse_mix_batch_f32(const float* clean, const float* noise, const int32_t* snr_idx, const double* snr_db_table_host,
                       int n_snr, float* clean_out, float* noisy_out, int32_t* status, int32_t* snr_idx_used,
                       const int64_t* snr_label_table, int64_t* snr_labels_out, int32_t* n_rejected, int B, int L,
                       int L_noise, int peak_norm, int max_attempts, int substitute_bad_rows, nrse_stream_t stream);
const char* nrse_mix_status_name(int status_code);
/* 4 (default): on-chip resident -- the CTAs of a cluster (1..8 per row) keep their segment of the row in registers and
 * shared memory between the three passes, exchanges by st.async + mbarrier; needs 16-byte aligned rows, L % 4 == 0,
 * L_noise >= L and a row of at most 8 x 40960 samples (20 s).  5: the same, forcing the 1024-thread one-CTA-per-SM shape.
 * 0: always the generic kernel (4 CTAs per row, pass 1 from HBM, passes 2-3 re-read from L2), which is also what rows
 * outside the resident kernel's limits (unaligned, tiled noise, > 20 s) fall back to by themselves. */
int nrse_mix_set_variant(int variant);
/* tuning: CTAs per row of the resident kernel (1..8); 0 = automatic (default) */
int nrse_mix_set_cluster(int ctas_per_row);
/* tuning: shared-memory carveout (percent of 228 KB) of the resident kernels; -1 = just what the CTAs need (default) */
int nrse_mix_set_carveout(int percent);

/* ---------------------------------------------------------------------------------------------
 * Multi-tensor EMA:  target = decay*target + one_minus_decay*online   (fp32, in place,
 * rounding identical to the reference expression: two products, one sum, no FMA).
 * Replaces BYOLSpeechModel._update_target_network, ref:src/models/byol.py:62-73.
 *
 * The launch works on a chunk table built once per model:
 *   nrse_ema_plan_chunks_host splits n_tensors tensors (host arrays of device pointers and element
 *   counts) into chunks of at most chunk_elems elements; it returns the number of chunks and, when
 *   the output arrays are non-NULL (capacity max_chunks), fills them.  Copy the three arrays to the
 *   device and pass them to nrse_ema_chunks_f32.
 * ------------------------------------------------------------------------------------------- */
int64_t nrse_ema_plan_chunks_host(const uint64_t* target_ptrs_host, const uint64_t* online_ptrs_host,
                                  const int64_t* numel_host, int n_tensors, int64_t chunk_elems,
                                  uint64_t* chunk_target_host, uint64_t* chunk_online_host,
                                  int32_t* chunk_numel_host, int64_t max_chunks);
int nrse_ema_chunks_f32(const uint64_t* chunk_target, const uint64_t* chunk_online,
                        const int32_t* chunk_numel, int64_t n_chunks,
                        float decay, float one_minus_decay, nrse_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Fused optimizer tail of the BYOL step: gradient-norm clip + AdamW + EMA of the target network, fp32,
 * two launches (one read of every gradient for the norm, then ONE pass that reads p, g, m, v, t and
 * writes p, m, v, t).  Replaces, in this order (ref:train_byol.py:67-71):
 *   torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)       ref:train_byol.py:67
 *   optimizer.step()   (torch.optim.AdamW, torch/optim/adam.py::_single_tensor_adam arithmetic)   :70
 *   model._update_target_network()                                         ref:src/models/byol.py:62-73
 *
 * nrse_optim_plan_chunks_host splits n_tensors parameter tensors into chunks of <= chunk_elems elements.
 *   Per tensor (HOST arrays of DEVICE addresses): p = parameter, g = gradient (0: no gradient this step,
 *   AdamW skips the tensor as torch does), m / v = exp_avg / exp_avg_sq (required when g != 0), t = EMA
 *   target twin (0: none).  Tensors with neither g nor t produce no chunk.  Returns the number of chunks;
 *   when chunk_ptrs_host / chunk_numel_host are non-NULL it fills them: chunk_ptrs_host is FIVE
 *   consecutive arrays of max_chunks entries (p | g | m | v | t), i.e. the pitch is max_chunks.
 * nrse_grad_sqnorm_chunks_f32 reads every gradient of a table once and writes nrse_optim_partials_count()
 *   fp64 partial sums of squares (one per CTA of a fixed grid; deterministic) to `partials`.
 * nrse_clip_adamw_ema_chunks_f32 runs one optimizer step over a DEVICE copy of the table.
 *   step >= 1 is the value of torch's state['step'] AFTER its increment (one value per call: parameters
 *   whose step counts differ go into separate tables / calls, their norm partials side by side);
 *   max_grad_norm <= 0 disables clipping, else every CTA sums partials[0..n_partials) in a fixed order and
 *   scales the gradients by min(1, max_grad_norm / (norm + 1e-6)) on the fly;
 *   grad_norm_out (nullable, device [1]) receives the total gradient norm clip_grad_norm_ returns.
 *   The gradients themselves are not modified (the reference scales them in place; nothing reads them
 *   afterwards: the next step starts with zero_grad).
 * ------------------------------------------------------------------------------------------- */
int64_t nrse_optim_plan_chunks_host(const uint64_t* p_host, const uint64_t* g_host, const uint64_t* m_host,
                                    const uint64_t* v_host, const uint64_t* t_host, const int64_t* numel_host,
                                    int n_tensors, int64_t chunk_elems, uint64_t* chunk_ptrs_host,
                                    int32_t* chunk_numel_host, int64_t max_chunks);
int nrse_optim_partials_count(void);
int nrse_grad_sqnorm_chunks_f32(const uint64_t* chunk_g, const int32_t* chunk_numel, int64_t n_chunks,
                                double* partials, nrse_stream_t stream);
int nrse_clip_adamw_ema_chunks_f32(const uint64_t* chunk_ptrs, int64_t chunk_pitch, const int32_t* chunk_numel,
                                   int64_t n_chunks, double lr, double beta1, double beta2, double eps,
                                   double weight_decay, int64_t step, double max_grad_norm, double ema_decay,
                                   const double* partials, int n_partials, float* grad_norm_out,
                                   nrse_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Batched Attentive Statistics Pooling, forward and backward (fp32).  Replaces the per-utterance loop of
 * AttentiveStatisticsPooling.forward, ref:src/models/pool.py:37-58 (consumer of the encoder on the emotion
 * fine-tune step, ref:src/models/emotion.py:60-79).  The linear layer `sap_linear` stays a library GEMM
 * run by the caller over all B*T frames; these entry points do everything after it.
 *   x, hl      [B,T,D] row-major fp32: encoder output and sap_linear(x) (pre-tanh); D % 4 == 0, T <= 4096
 *   attention  [D]     the attention vector (ref:src/models/pool.py:33)
 *   lens       [B]     int32 valid frames per utterance, min(compute_length_from_mask(mask), T)  (:44,47)
 *   out        [B,2D]  (mu | rh):  w = softmax_t(tanh(hl) . attention) over t < len,  mu = sum_t w x,
 *                      rh = sqrt(clamp(sum_t w x^2 - mu^2, min=1e-5))                  (:48-55)
 *   weights    [B,T]   the softmax weights (0 for t >= len), kept for the backward
 *   logits_ws  [B,T]   fp32 scratch;  workspace: nrse_asp_pool_bwd_workspace_bytes(B, T, D) bytes of scratch
 * Backward: grad_x is the DIRECT gradient w.r.t. x (the path through sap_linear comes back from the
 * caller's GEMM backward on grad_hl); grad_attention [D] is written (not accumulated) from per-CTA partial
 * rows summed in a fixed order: deterministic, no atomics.
 * ------------------------------------------------------------------------------------------- */
int nrse_asp_pool_fwd(const float* x, const float* hl, const float* attention, const int32_t* lens, float* out,
                      float* weights, float* logits_ws, int B, int T, int D, nrse_stream_t stream);
size_t nrse_asp_pool_bwd_workspace_bytes(int B, int T, int D);
int nrse_asp_pool_bwd(const float* x, const float* hl, const float* attention, const int32_t* lens,
                      const float* out, const float* weights, const float* grad_out, float* grad_x,
                      float* grad_hl, float* grad_attention, void* workspace, size_t workspace_bytes,
                      int B, int T, int D, nrse_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * BYOL loss with fused +1e-10, L2 normalisation (eps 1e-10), row dot product, clamp and mean.
 * Replaces byol_loss, ref:src/models/byol.py:104-129.  One launch, no host sync.
 *   p, z   [B,D] row-major, dtype NRSE_DTYPE_F32 or NRSE_DTYPE_BF16 (D % 4 == 0 for f32, % 8 for bf16)
 *   loss   [1] fp32 = 2 - 2*mean_b clamp(<p^,z^>, -1, 1)
 *   saved  [B,4] fp32 = (||p+1e-10||, ||z+1e-10||, unclamped similarity, 0) for the backward
 *   row_sim (nullable) [B] fp32 clamped similarities (evaluate_byol.py:55 uses these per row)
 *   flags  (nullable) [1] int32: the reference's two diagnostics (ref:src/models/byol.py:109-122), evaluated inside
 *          the same launch: bit 0 / 1 = NaN in p / z before normalisation, bit 2 / 3 = NaN in p / z after it
 *          (an input NaN, or an Inf that normalises to Inf / Inf)
 * Backward w.r.t. p only (the target branch is under no_grad, ref:src/models/byol.py:94-96).
 * ------------------------------------------------------------------------------------------- */
int nrse_byol_loss_fwd(const void* p, const void* z, float* loss, float* saved, float* row_sim, int32_t* flags,
                       int B, int D, int dtype, nrse_stream_t stream);
int nrse_byol_loss_bwd(const void* p, const void* z, const float* saved, const float* grad_loss,
                       void* grad_p, int B, int D, int dtype, nrse_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * WavLM convolutional feature encoder, forward.
 * Replaces WavLMFeatureEncoder.forward, hf:models/wavlm/modeling_wavlm.py:779-789 (reached from
 * ref:src/models/encoder.py:25): 7 x { Conv1d(bias=False) ; LayerNorm over C | GroupNorm | none ;
 * exact GELU }, k=(10,3,3,3,3,2,2), s=(5,2,2,2,2,2,2), C=512.
 *
 * Activations are channels-last [B, P_i, 512] bf16 with a per-utterance frame pitch P_i >= T_i chosen
 * so that P_{i-1} = s_i * P_i (nrse_conv_frontend_geometry); then output frame m = b*P_i + t of layer i
 * reads the contiguous window x[(s_i*m) .. (s_i*m + k_i)) * 512 of layer i-1, and the conv is one
 * implicit GEMM [B*P_i, k_i*512] x [k_i*512, 512] with no im2col.
 *
 *   x          [B,L] fp32 waveform (already z-normalised)
 *   w0         [512,10] fp32   (= conv_layers.0.conv.weight [512,1,10])
 *   w_packed   6 device pointers, layer i=1..6: bf16 [512, k_i*512] with K index = tap*512 + c_in
 *              (nrse_conv_frontend_pack_weights converts from the checkpoint layout [512,512,k] fp32)
 *   gamma/beta 7 device pointers [512] fp32; NULL where the layer has no norm (group mode, i >= 1)
 *   y          [B, P_6, 512] y_dtype (bf16 or f32); frames t >= T_6 of each utterance are padding
 *   workspace  nrse_conv_frontend_workspace_bytes(B, L) bytes, 1024-byte aligned
 *   acts_out   nullable array of 6 HOST-visible slots receiving the device addresses (inside the
 *              workspace) of the layer 0..5 outputs, for the backward / debugging
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* w0;
  const void* w_packed[NRSE_FRONTEND_LAYERS - 1];
  const float* gamma[NRSE_FRONTEND_LAYERS];
  const float* beta[NRSE_FRONTEND_LAYERS];
} nrse_frontend_params;

/* T[i], P[i] for i = 0..6.  Returns NRSE_ERR_INVALID_ARG if L < 400 (empty output). */
int nrse_conv_frontend_geometry(int L, int32_t* T_out_host, int32_t* P_out_host);
size_t nrse_conv_frontend_workspace_bytes(int B, int L);
/* w [512,512,k] fp32 (checkpoint layout) -> packed bf16 [512, k*512] */
int nrse_conv_frontend_pack_weights(const float* w, void* w_packed, int k, nrse_stream_t stream);
int nrse_conv_frontend_fwd(const float* x, const nrse_frontend_params* params_host, int norm_mode,
                           void* y, int y_dtype, void* workspace, size_t workspace_bytes,
                           uint64_t* acts_out_host, int B, int L, nrse_stream_t stream);

/* The two building blocks of nrse_conv_frontend_fwd, exported for per-layer parity tests and for callers
 * that keep their own activation buffers.
 *   layer 0 (WavLMLayerNormConvLayer / WavLMGroupNormConvLayer with C_in = 1, hf:...modeling_wavlm.py:703-751):
 *     x [B,L] fp32 -> out [B*P0, 512] bf16 (frames t >= T0 zero-filled).  norm_mode LAYER: gamma/beta [512];
 *     GROUP: gamma/beta [512] and gn_scratch = the tail of the frontend workspace (statistics over time).
 *   layers 1..6 (stride 2, k = 3 or 2): act_prev [rows_prev = 2*rows_out, 512] bf16 -> out [rows_out, 512]
 *     (bf16 or f32); gamma/beta NULL = no normalisation (WavLMNoLayerNormConvLayer, :682-700). */
int nrse_conv_layer0_fwd(const float* x, const float* w0, const float* gamma, const float* beta, int norm_mode,
                         void* out, void* gn_scratch, int B, int L, int T0, int P0, nrse_stream_t stream);
int nrse_conv_layer_fwd(const void* act_prev, int64_t rows_prev, const void* w_packed, int k, int stride,
                        const float* gamma, const float* beta, void* out, int out_dtype, int64_t rows_out,
                        nrse_stream_t stream);
/* Tile decomposition of the tcgen05 kernel: 1 = one CTA owns all 512 channels of a 128-frame tile,
 * 2 = a 2-CTA cluster splits the channels and exchanges LayerNorm partials through DSMEM,
 * 3 = as 2, but the inference forward of the GEMM layers runs the 2-SM UMMA kernel (tcgen05.mma.cta_group::2: the CTA
 *     pair splits the FRAMES, each CTA stages half of the weight rows; same results to bf16 rounding),
 * 4 = as 2, but nrse_conv_frontend_fwd runs layers 1-3 on the 2-SM kernel (default; the choice never depends on the
 *     batch, so an utterance's features do not depend on what else is in the batch). */
int nrse_conv_frontend_set_variant(int variant);
/* Layer-0 kernel in LayerNorm mode: 0 = SIMT (warp per frame); 1 = tensor cores (hi/lo-split K=32 UMMA, LayerNorm +
 * GELU epilogue; always used by the training forward); 2 = tensor cores with LayerNorm folded into the GEMM operands
 * (K=48, GELU-only epilogue; inference forward); 3 = 2 with 16 epilogue warps (default). */
int nrse_conv_frontend_set_layer0_variant(int variant);
/* Tile order of the GEMM layers inside nrse_conv_frontend_fwd / _fwd_train: 1 (default) = consecutive layers walk their
 * tiles in opposite directions (each layer starts on the rows its producer wrote last, which are still in L2);
 * 0 = every layer first-to-last.  Results do not depend on it. */
int nrse_conv_frontend_set_tile_order(int alternate);
/* 1 = nrse_conv_frontend_bwd (LayerNorm mode) runs the LayerNorm + GELU backward of layers 0-5 inside the epilogue of the
 * data-gradient GEMM of the layer above (nrse_conv_layer_dgrad_lnbwd) instead of as kernels of their own: dOut_i never
 * makes the round trip through HBM.  0 = separate kernels.  Same results up to bf16 rounding of dOut_i (the fused form
 * keeps it in fp32). */
int nrse_conv_frontend_set_bwd_fusion(int on);
/* Number of SMs (8..148, default 148) the persistent kernels of the conv frontend / feature projection spread over.  The
 * data-parallel training step lowers it while the gradient all-reduce is in flight: NCCL's kernels need SMs of their own,
 * and a machine filled with one resident persistent CTA per SM would make them wait for the end of every kernel. */
int nrse_conv_frontend_set_sm_budget(int sms);
/* 1: the TMA producer of the GEMM layers bulk-prefetches the next tile's input frames into L2 (default 0: measured
 * 2-3 % slower at 64 x 4 s -- the operand feed is not HBM-latency bound). */
int nrse_conv_frontend_set_l2_prefetch(int on);

/* ---------------------------------------------------------------------------------------------
 * Training forward and backward of the feature encoder, both norm modes.
 * Replaces what autograd records through hf:models/wavlm/modeling_wavlm.py:703-727 x 7 (LayerNorm mode, wavlm-large)
 * or :730-751 + :682-700 x 6 (GroupNorm mode, wavlm-base) for the online branch (ref:train_byol.py:66
 * `loss.backward()`; partial unfreeze: ref:src/models/emotion.py:114-129).  The training forward additionally keeps,
 * on a caller-provided tape, the bf16 activations of layers 0..5, `xhat` of every layer (the normalised pre-affine
 * activation; for a layer without a normalisation the pre-GELU activation itself) and the statistics the backward
 * needs (1/std per frame in LayerNorm mode, 1/std per (utterance, channel) of layer 0 in GroupNorm mode).
 * Backward, per layer from 6 down:  dOut -> dZ (norm + GELU backward, also dgamma / dbeta), dW = dZ^T A (tcgen05,
 * MN-major operands, split-K, fp32 atomics), dX = dZ W (two forward-like GEMMs over even / odd frames).  Layer 0:
 * SIMT weight gradient (LayerNorm mode) / one fused pass that never materialises dZ (GroupNorm mode).
 * The input waveform gradient is not computed (nothing upstream of the waveform is trainable in the reference).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* wt_even[NRSE_FRONTEND_LAYERS - 1]; /* bf16 [512 c_in, n_even*512]: taps (0,2) for k=3, (0) for k=2 */
  const void* wt_odd[NRSE_FRONTEND_LAYERS - 1];  /* bf16 [512 c_in, 512]: tap 1 */
} nrse_frontend_bwd_weights;
/* Gradients, fp32, in the CHECKPOINT layouts.  Every pointer is nullable: NULL = "this tensor needs no gradient"
 * (frozen parameter).  Nothing below the lowest layer that wants a gradient is computed, and a layer's weight-gradient
 * GEMM is skipped when its dw is NULL.  Non-NULL tensors are ACCUMULATED into (+=): the caller zeroes them (one memset
 * over its gradient arena).  dgamma[i] / dbeta[i] must be both NULL or both set; GroupNorm mode has them for layer 0 only. */
typedef struct {
  float* dw0;                             /* [512, 1, 10] */
  float* dw[NRSE_FRONTEND_LAYERS - 1];    /* [512, 512, k_i] */
  float* dgamma[NRSE_FRONTEND_LAYERS];    /* [512] */
  float* dbeta[NRSE_FRONTEND_LAYERS];
} nrse_frontend_grads;

size_t nrse_conv_frontend_tape_bytes(int B, int L);
int nrse_conv_frontend_fwd_train(const float* x, const nrse_frontend_params* params_host, int norm_mode, void* y,
                                 int y_dtype, void* tape, size_t tape_bytes, int B, int L, nrse_stream_t stream);
/* w [512,512,k] fp32 -> the two data-gradient operands described above */
int nrse_conv_frontend_pack_weights_dgrad(const float* w, void* wt_even, void* wt_odd, int k, nrse_stream_t stream);
size_t nrse_conv_frontend_bwd_workspace_bytes(int B, int L);
/* dy [B, dy_pitch, 512] fp32: the gradient of the features; frame (b, t < T_6) at row b * dy_pitch + t (dy_pitch = T_6 for
 * a compact gradient, P_6 for a pitch-padded one; rows t >= T_6 are never read). */
int nrse_conv_frontend_bwd(const float* x, const nrse_frontend_params* params_host,
                           const nrse_frontend_bwd_weights* bwd_weights_host, int norm_mode, const void* tape,
                           const float* dy, int dy_pitch, const nrse_frontend_grads* grads_host, void* workspace,
                           size_t workspace_bytes, int B, int L, nrse_stream_t stream);
/* building blocks (per-layer parity tests).
 *   nrse_ln_gelu_bwd: dOut -> dZ.  dout bf16 [rows, 512] (may alias dz) or fp32 [B, dout_pitch, 512]; gamma == NULL
 *     selects the no-norm form dZ = dOut gelu'(xhat) (xhat = the pre-GELU activation); dgamma / dbeta nullable, accumulated.
 *   nrse_conv_layer_wgrad: dW += dZ^T A, fp32, packed K order [512, tap*512 + c] or (ckpt_layout) [512, 512, k].
 *   nrse_conv_layer0_gn_bwd: GroupNorm-mode layer 0, dOut0 bf16 [B*P0, 512] + xhat0 + per-(b,c) 1/std [B,512] ->
 *     dw0 / dgamma / dbeta (nullable, accumulated); scratch = nrse_conv_layer0_gn_bwd_scratch_bytes(B) bytes. */
int nrse_ln_gelu_bwd(const void* dout, int dout_dtype, int dout_pitch, const void* xhat, const float* rstd,
                     const float* gamma, const float* beta, void* dz, float* dgamma, float* dbeta, int64_t rows, int P,
                     int T, nrse_stream_t stream);
int nrse_conv_layer0_wgrad(const float* x, const void* dz0, float* dw0, int B, int L, int T0, int P0,
                           nrse_stream_t stream);
size_t nrse_conv_layer0_gn_bwd_scratch_bytes(int B);
int nrse_conv_layer0_gn_bwd(const float* x, const void* dout0, const void* xhat0, const float* gn_rstd,
                            const float* gamma, const float* beta, float* dw0, float* dgamma, float* dbeta,
                            void* scratch, int B, int L, int T0, int P0, nrse_stream_t stream);
int nrse_conv_layer_wgrad(const void* dz, const void* act_prev, int64_t rows_out, int k, float* dw, int ckpt_layout,
                          nrse_stream_t stream);
int nrse_conv_layer_dgrad(const void* dz, int64_t rows_out, const void* wt_even, const void* wt_odd, int k, void* dx,
                          nrse_stream_t stream);
/* nrse_conv_layer_dgrad followed by nrse_ln_gelu_bwd of the layer below, in one kernel: dz_prev [2*rows_out, 512] bf16
 * receives dZ_{i-1}; dgamma_prev / dbeta_prev (both or neither, fp32 [512]) are accumulated into.  xhat_prev / rstd_prev:
 * what nrse_conv_frontend_fwd_train saved for layer i-1 (frame pitch P_prev, T_prev valid frames per utterance; padding
 * frames get zeros).  Replaces the autograd backward of Conv1d -> LayerNorm -> GELU across two layers,
 * hf:models/wavlm/modeling_wavlm.py:250-275.  As for nrse_conv_layer_dgrad, the rows of dz that belong to pitch padding
 * must be zero (every kernel of this library that writes a gradient buffer leaves them so). */
int nrse_conv_layer_dgrad_lnbwd(const void* dz, int64_t rows_out, const void* wt_even, const void* wt_odd, int k,
                                const void* xhat_prev, const float* rstd_prev, const float* gamma_prev,
                                const float* beta_prev, void* dz_prev, float* dgamma_prev, float* dbeta_prev, int P_prev,
                                int T_prev, nrse_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Feature projection: LayerNorm(512) + Linear(512 -> 1024) on the conv features, forward and backward.
 * Replaces WavLMFeatureProjection.forward, hf:models/wavlm/modeling_wavlm.py:93-105 (reached from
 * ref:src/models/encoder.py:25 through WavLMModel.forward, hf:...:1061-1064); the dropout behind it stays with the caller.
 *   pack     projection.weight [1024, 512] fp32 -> bf16 copy (forward operand) and its transpose [512, 1024] (backward)
 *   fwd      feats [B, feats_pitch, 512] fp32 / bf16 -- the conv frontend's channels-last output, read in place
 *            (feats_pitch = P_6 for the pitched kernel output, T for a compact tensor) -> hidden [B*T, 1024] fp32
 *            (= projection(layer_norm(feats)) + bias), norm_hidden (nullable) [B*T, 512] fp32 = layer_norm(feats), HF's
 *            second output.  One LayerNorm launch (writes the bf16 GEMM operand into `tape`) + one tcgen05 GEMM launch.
 *            tape: nrse_feature_projection_tape_bytes(B*T) bytes, 1024-aligned; always needed (it holds the GEMM operand);
 *            training != 0 additionally keeps xhat / rstd there for the backward.
 *   bwd      d_hidden [rows, 1024] fp32 -> d_feats [rows, 512] fp32 (nullable; the `dy` of nrse_conv_frontend_bwd with
 *            dy_pitch = T), d_ln_gamma / d_ln_beta [512], d_w [1024, 512], d_bias [1024]: all nullable, ACCUMULATED.
 *            workspace: nrse_feature_projection_bwd_workspace_bytes(rows) bytes, 1024-aligned.
 * ------------------------------------------------------------------------------------------- */
int nrse_feature_projection_pack(const float* w, void* w_bf16, void* wt_bf16, nrse_stream_t stream);
size_t nrse_feature_projection_tape_bytes(int64_t rows);
int nrse_feature_projection_fwd(const void* feats, int feats_dtype, int B, int T, int feats_pitch, const float* ln_gamma,
                                const float* ln_beta, float eps, const void* w_bf16, const float* bias, float* hidden,
                                float* norm_hidden, void* tape, int training, nrse_stream_t stream);
size_t nrse_feature_projection_bwd_workspace_bytes(int64_t rows);
int nrse_feature_projection_bwd(const float* d_hidden, const void* tape, const float* ln_gamma, const float* ln_beta,
                                const void* wt_bf16, float* d_feats, float* d_ln_gamma, float* d_ln_beta, float* d_w,
                                float* d_bias, void* workspace, int64_t rows, nrse_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Positional convolution embedding: Conv1d(1024, 1024, k = 128, padding = 64, groups = 16) with
 * weight-norm over dim 2, last frame dropped, exact GELU; forward and backward.
 * Replaces WavLMPositionalConvEmbedding.forward, hf:models/wavlm/modeling_wavlm.py:48-90 (the consumer of the feature
 * projection inside the encoder; reached from ref:src/models/encoder.py:25 and, on the emotion fine-tune step, from
 * ref:src/models/emotion.py:60-79).  wavlm-large geometry only (hidden 1024, 16 groups, 128 taps).
 *   pack   v [1024, 64, 128] = conv.parametrizations.weight.original1, g [128] = ...original0  ->  w = g v / ||v||_tap as
 *          two bf16 GEMM operand packs of nrse_pos_conv_pack_bytes() bytes each (forward; tap-reversed + transposed for the
 *          data gradient) and normsq [128] fp32 = ||v[:, :, tap]||^2 (kept for the backward).
 *   fwd    x [B, T, 1024] fp32 contiguous -> y [B, T, 1024] fp32; x_bf16 [B*T, 1024] receives the bf16 operand copy of x
 *          (the backward's weight gradient reads it again); z_save (nullable) [B*T, 1024] bf16 = the pre-GELU activation.
 *          One cast launch + one tcgen05 implicit-GEMM launch (no padded copy: the TMA engine zero-fills frames < 0, >= T).
 *   bwd    d_y [B, T, 1024] fp32 -> d_x (nullable) [B, T, 1024] fp32 written; d_v [1024, 64, 128], d_g [128] (both or
 *          neither), d_bias [1024] (nullable) ACCUMULATED.  workspace: nrse_pos_conv_bwd_workspace_bytes(B, T), 1024-aligned.
 * ------------------------------------------------------------------------------------------- */
size_t nrse_pos_conv_pack_bytes(void);
int nrse_pos_conv_pack(const float* v, const float* g, void* w_fwd, void* w_bwd, float* normsq, nrse_stream_t stream);
int nrse_pos_conv_fwd(const float* x, const void* w_fwd, const float* bias, float* y, void* x_bf16, void* z_save, int B, int T,
                      nrse_stream_t stream);
size_t nrse_pos_conv_bwd_workspace_bytes(int B, int T);
int nrse_pos_conv_bwd(const float* d_y, const void* x_bf16, const void* z_save, const void* w_bwd, const float* v,
                      const float* g, const float* normsq, float* d_x, float* d_v, float* d_g, float* d_bias,
                      void* workspace, int B, int T, nrse_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Gradient all-reduce (mean over ranks) through NVSwitch multicast memory: the one collective of the data-parallel step
 * (what DistributedDataParallel's all-reduce computes around ref:train_byol.py:66-70).
 *   multicast_ptr  the MULTICAST address of a symmetric fp32 allocation present on all `world` ranks (host side:
 *                  torch.distributed._symmetric_memory.rendezvous(...).multicast_ptr)
 *   elem_offset / numel  the range to reduce (multiples of 4 elements); rank r sums and rewrites the r-th slice of it:
 *                  multimem.ld_reduce.add (the switch returns the sum over the ranks) -> x 1/world -> multimem.st (the
 *                  switch stores to every rank).  One launch of <= max_ctas (0 = 148) register-light CTAs that co-reside
 *                  with the persistent conv kernels running concurrently on another stream.
 * The caller orders the ranks: a symmetric-memory barrier on `stream` before the launch (every rank's gradients are
 * written) and after it (every slice is stored). */
int nrse_multimem_allreduce_mean_f32(void* multicast_ptr, int64_t elem_offset, int64_t numel, int rank, int world,
                                     int max_ctas, nrse_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NRSE_B200_H_ */
