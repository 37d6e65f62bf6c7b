/*
 * speech_inpainting_b200.h - C-ABI of the B200-native Speech-Inpainting hot path.
 *
 * The reference (Fireflies-17/Speech-Inpainting) is pure Python with no FFI of its own;
 * its "operator boundary" for the north-star path is the nn.Module call surface
 *   CustomModel.forward / HubertModel.forward        I_ea/model.py:80-89
 *   HubertFeatureReader.get_feats / extract_features I_da/src/hubert_feature_reader.py:44-67
 *   Generator.forward                                I_ea/hifi_gan/models.py:107-123, I_da/src/models.py:209-225
 *   CodeGenerator.forward                            I_da/src/model.py:121-189
 * plus the glue in I_ea/predict.py:85-207 and I_da/scripts/inpainting.py:181-259.
 * Each entry point below names the reference lines it replaces.  INTEGRATION.md shows the
 * ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers + sizes only; every pointer is a DEVICE pointer unless named host_*.
 *   - no allocation inside the library: outputs and workspaces are caller-owned.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*).
 *   - return 0 (SIB_OK) or a negative/positive error code; sib_last_error() gives the text.
 *     No error is swallowed and there is NO CPU fallback.
 *   - activations are "frame-major": [batch][time][channels], channels contiguous.
 *   - re-entrant across devices: no global mutable state except a thread-local error string.
 */
#ifndef SPEECH_INPAINTING_B200_H_
#define SPEECH_INPAINTING_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIB_ABI_VERSION 2
#define SIB_MAX_TAPS 128

enum sib_status { SIB_OK = 0, SIB_ERR_INVALID = 1, SIB_ERR_CUDA = 2, SIB_ERR_UNSUPPORTED = 3 };
enum sib_act { SIB_ACT_NONE = 0, SIB_ACT_GELU = 1, SIB_ACT_LRELU = 2, SIB_ACT_TANH = 3 };
enum sib_dtype { SIB_F32 = 0, SIB_BF16 = 1, SIB_I16 = 2 /* int16 PCM, sib_resample input only */ };

typedef void* sib_stream_t; /* cudaStream_t */

int sib_abi_version(void);
const char* sib_last_error(void);
/* number of kernels this library has launched from the calling thread (bench `gpu_launches`) */
long long sib_launch_count(void);
/* Programmatic dependent launch of the tensor-core / LayerNorm kernels (on by default; SIB_NO_PDL=1 in the environment
 * turns it off).  Returns the previous setting.  Off = plain stream-ordered launches, e.g. for per-kernel timing. */
int sib_set_pdl(int enabled);

/* ------------------------------------------------------------------------------------------
 * Generic frame-major 1-D convolution / linear layer with fused epilogue.
 *   y[b,t,co] = post( out_scale * ( bias[co] + sum_j sum_ci pre(x[b, t*stride + tap_offset[j], ci]) * w[g][j][ci][co]
 *                                    + residual[b,t,co] + (accumulate ? y[b,t,co] : 0) ) )
 *   (res_after_act=1 moves `+ residual` outside post())
 * rows outside [0,t_in) read as zero (== zero padding).  Covers
 *   torch.nn.Conv1d (+dilation/padding/stride/groups)   HF:106-175 conv layers, HF:83-92 pos-conv,
 *                                                       models.py:36-43 ResBlock convs, :108 conv_pre
 *   torch.nn.ConvTranspose1d (poly-phase: c_out = stride*Cout)  models.py:110-111 `ups[i]`
 *   torch.nn.Linear (batch=1, n_taps=1)                 HF:228-229, 320-324, 343, 363-367; model.py:88
 * w layout: [groups][n_taps][c_in/groups][c_out/groups], c_out contiguous.
 * ------------------------------------------------------------------------------------------ */
typedef struct sib_conv_desc {
  int32_t batch, t_in, t_out, c_in, c_out, groups;
  int32_t n_taps, stride;
  int32_t tap_offset[SIB_MAX_TAPS];
  int64_t x_batch_stride, y_batch_stride, r_batch_stride; /* elements */
  int32_t x_row_stride, y_row_stride, r_row_stride;       /* elements */
  int32_t pre_act;   /* SIB_ACT_NONE | SIB_ACT_LRELU applied to x on load */
  float pre_slope;
  int32_t post_act;  /* sib_act */
  float post_slope;
  float out_scale;
  int32_t accumulate;
  int32_t res_after_act; /* 0: residual joins the sum before post(); 1: y = post(...) + residual (HF:440-441) */
  /* bf16 tensor-core path only */
  float act2_slope;      /* second output y_act = leaky_relu(y, act2_slope) */
} sib_conv_desc;

int sib_conv1d_f32(const sib_conv_desc* d, const float* x, const float* w, const float* bias,
                   const float* residual, float* y, sib_stream_t stream);

/* conv with a single output channel + activation: models.py:119-121 lrelu(0.01) -> conv_post -> tanh.
 * x [B,T,C] frame-major, w [k][C], y [B,T].  */
int sib_conv1d_cout1_f32(const float* x, const float* w, const float* bias, float* y, int batch, int t,
                         int c, int k, int pad, float pre_slope, int post_act, sib_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * HuBERT conv0 (Conv1d 1->C, k=10, s=5) fused with its norm (HF:154-175 / HF:127-151).
 *   mode 0: GroupNorm statistics only -> partial[b][tile][c][2] (sum, sumsq); tile = 64 frames
 *   mode 1: recompute conv0, GroupNorm(C groups) with mean/rstd[b][c], affine, exact GELU -> y[B,T0,C]
 *   mode 2: conv0 + bias only (feat_extract_norm == "layer"; LayerNorm+GELU follow as sib_layernorm)
 * sib_gn_finalize reduces the partials to mean / rstd (double accumulation, eps 1e-5).
 * ------------------------------------------------------------------------------------------ */
int sib_conv0_f32(int mode, const float* wave, int batch, int n_samples, int64_t wave_batch_stride,
                  const float* w /*[C][k]*/, const float* bias /*nullable*/, int c, int k, int stride, int t0,
                  float* partial, const float* mean, const float* rstd, const float* gamma,
                  const float* beta, float* y, sib_stream_t stream);
int sib_conv0_num_tiles(int t0);
/* GroupNorm(C, C) statistics of conv0's output in closed form from the waveform's lag sums (conv0 is linear in the
 * samples): mean[b,c], rstd[b,c] = 1/sqrt(var + eps), same outputs as mode 0 + sib_gn_finalize_f32 without
 * evaluating conv0 (HF:154-175, torch GroupNorm biased variance).  k = 10, stride = 5 only. */
int sib_conv0_gn_stats_f32(const float* wave, int batch, int n_samples, int64_t wave_batch_stride, const float* w,
                           const float* bias, int c, int k, int stride, int t0, float eps, float* mean, float* rstd,
                           sib_stream_t stream);
int sib_gn_finalize_f32(const float* partial, int batch, int n_tiles, int c, int t0, float eps, float* mean,
                        float* rstd, sib_stream_t stream);

/* LayerNorm over the channel axis of `rows` rows (HF:396,398,442,613,147,228; model.py:87):
 *   y = act( LN(x + residual) * gamma + beta ), residual nullable, act in {NONE, GELU}. In-place allowed. */
int sib_layernorm_f32(const float* x, const float* residual, const float* gamma, const float* beta, float* y,
                      int64_t rows, int c, float eps, int post_act, sib_stream_t stream);

/* Multi-head self-attention (HF:234-259 eager path): qkv packed [B,T,3H] (q|k|v), softmax in fp32,
 * scale d^-1/2, keys >= key_len[b] masked (key_len nullable => no padding). out [B,T,H]. head_dim = 64. */
int sib_attention_f32(const float* qkv, const int32_t* key_len, float* out, int batch, int t, int heads,
                      int head_dim, sib_stream_t stream);

/* zero frames t >= key_len[b] of h[B,T,C] (HF:429-432). */
int sib_zero_padded_frames_f32(float* h, const int32_t* key_len, int batch, int t, int c, sib_stream_t stream);

/* ------------------------------------------------------------------------------------------ glue */
/* a1: wave[b, lo[b]:hi[b]] = 0 (numpy slice semantics; I_ea/predict.py:133, dataset.py:82).
 * add_eps != 0 first adds it to every sample: y_inp = (y + 1e-6) * mask (I_da inpainting.py:187-191). */
int sib_zero_ranges_f32(float* wave, int batch, int n, const int32_t* lo, const int32_t* hi, float add_eps,
                        sib_stream_t stream);
/* a2: per-utterance (x-mean)/sqrt(var+eps); lengths nullable (HF feature_extraction_wav2vec2.py:78-97,
 * eps 1e-7, tail -> 0) ; eps 1e-5 reproduces F.layer_norm(x, x.shape) (hubert_feature_reader.py:53-54). */
int sib_znorm_f32(const float* x, float* y, int batch, int n, const int32_t* lengths, float eps,
                  sib_stream_t stream);
/* 8f row 1 (I_ea/predict.py:99-103): y = normalize(x with [lo[b], hi[b]) zeroed) * scale, where normalize is
 * librosa.util.normalize (x / max|x|; rows whose peak is below FLT_MIN are left unscaled).  lo / hi nullable (no mask).
 * Feeds sib_mel_spectrogram_f32(hop 441, pad 312) = get_mel (I_ea/dataset/mel_dump.py:96-98). */
int sib_mask_peak_normalize_f32(const float* x, float* y, int batch, int n, const int32_t* lo, const int32_t* hi,
                                float scale, sib_stream_t stream);
/* a11: ragged gather of masked frames: out[off[b]+i, :] = src[b, pos[b]+i, :], i < len[b] (predict.py:164-168) */
int sib_gather_frames_f32(const float* src, int batch, int t, int d, const int32_t* pos, const int32_t* len,
                          const int32_t* off, float* out, sib_stream_t stream);
/* a12: labels[m] = argmax_k cos(v[m], cc[k]) (loss_fn.py:44-46), cc = centred codebook [K,D]; ties -> lowest k */
int sib_cos_argmax_f32(const float* v, const float* cc, int m, int k, int d, int64_t* labels, sib_stream_t stream);
/* a17: labels[m] = argmin_k ||f[m]-mu[k]||^2 (sklearn KMeans.predict, inpainting.py:204-205) */
int sib_l2_argmin_f32(const float* f, const float* mu, int m, int k, int d, int64_t* labels, sib_stream_t stream);
/* a17 at scale: argmin_k ||f - mu_k||^2 = argmax_k (f . mu_k - 0.5 ||mu_k||^2) - the form sklearn evaluates for float32
 * features.  sib_row_sqnorm_f32 gives the bias (out[r] = scale * sum_c x[r,c]^2, scale = -0.5), sib_conv1d_f32 (a linear
 * layer, w = mu^T) the scores [M, K], sib_row_argmax_f32 the labels (ties -> lowest k). */
int sib_row_sqnorm_f32(const float* x, int rows, int d, float scale, float* out, sib_stream_t stream);
int sib_row_argmax_f32(const float* s, int rows, int k, int64_t* labels, sib_stream_t stream);
/* a10 head on the gathered frames (I_ea/model.py:75-78,88 Linear(H, 80) after LayerNorm): y[m,:] = x[m,:] @ w + bias for a
 * NARROW output (n <= 128), w [k][n] (= sib_conv1d_f32 layout of a linear layer); one CTA per row. */
int sib_linear_skinny_f32(const float* x, const float* w, const float* bias, float* y, int m, int k, int n,
                          sib_stream_t stream);
/* a13: mel[b,:,pos[b]+i] = cc[labels[off[b]+i]] + center, i < len[b]; mel is channels-first [B,D,T] (predict.py:184-187).
 * cc has k rows; a label outside [0, k) is a device-side assert (printf + trap), as torch indexing does. */
int sib_paste_centroids_f32(float* mel, int batch, int d, int t, const float* cc, const float* center,
                            const int64_t* labels, const int32_t* pos, const int32_t* len, const int32_t* off, int k,
                            sib_stream_t stream);
/* a14: extend_mel (inference_modified.py:16-19): linear resize along time by 441/256, align_corners=False.
 * in channels-first [B,D,T]; out channels-first [B,D,Tm] (frame_major=0) or frame-major [B,Tm,D] (1). */
int sib_extend_mel_f32(const float* in, float* out, int batch, int d, int t, int tm, int frame_major,
                       sib_stream_t stream);
/* [B,C,T] <-> [B,T,C] */
int sib_transpose_f32(const float* in, float* out, int batch, int rows, int cols, sib_stream_t stream);
/* a18 front (I_da/src/model.py:141-172): out[b,t,:] = emb_c[code[b,t]] | emb_p[zp[b,t/rep]] | spk[b]  (frame-major).
 * emb_c has n_codes rows, emb_p n_bins rows; an index outside its table is a device-side assert, as nn.Embedding does. */
int sib_embed_concat_f32(const int64_t* code, const int64_t* zp, const float* spk, const float* emb_c,
                         const float* emb_p, float* out, int batch, int t, int t_p, int e, int e_spk, int n_codes,
                         int n_bins, sib_stream_t stream);
/* weight-norm folding (`remove_weight_norm()` I_ea/hifi_gan/models.py:125-132, I_da/src/models.py:227-233: dim 0;
 * HF pos-conv weight_norm(dim=2) HF:59-78): w = v * (g / ||v||).  v, w [outer][inner] fp32.
 * norm_dim_last = 0: one norm per row (g [outer]); 1: one norm per LAST-axis index over all rows (g [inner]). */
int sib_weight_norm_fold_f32(const float* v, const float* g, float* w, int outer, int inner, int norm_dim_last,
                             sib_stream_t stream);
/* a19: int16 = (int16)(int32)trunc(y*32768) (dataset.py:241-243, predict.py:125-127) */
int sib_pack_int16_f32(const float* y, int16_t* out, int64_t n, sib_stream_t stream);

/* a20: log-mel spectrogram (meldataset.py:49-79 / mel_dump.py:40-98): reflect pad, hann-1024 STFT,
 * sqrt(re^2+im^2+1e-9), mel basis [n_mels][513] (sparse rows), log(clamp(.,1e-5)). out [B,n_mels,frames]. */
int sib_mel_spectrogram_f32(const float* wave, int batch, int n, int hop, int pad, const float* mel_basis,
                            int n_mels, float* out, int frames, void* workspace /* sib_mel_workspace_bytes(n_mels) */,
                            sib_stream_t stream);
size_t sib_mel_workspace_bytes(int n_mels);

/* ------------------------------------------------------------------------------------------ 8f rows 3 and 4
 * 8f row 3: rational poly-phase resampler, the device side of `librosa.load(path, sr=22050)` / `sr=16000`
 * (I_ea/predict.py:79-80) and `resampy.resample(data, sr, 16000)` (I_da/scripts/preprocess.py:43-45):
 *   y[b, n] = sum_{j < taps} filt[j][n % up] * x[b, (n*down)/up + first + j],   x == 0 outside [0, len_in[b]) (nullable)
 * x is fp32 or int16 PCM (SIB_I16: scaled by 1/32768 on load, the soundfile / librosa convention); `filt` is the
 * windowed-sinc prototype sampled per output residue, [taps][up] with the residue contiguous (built on the host by
 * speech_inpainting_b200.audio.resample_filter).  up = down = taps = 1, first = 0, filt = {1} is a plain PCM -> float
 * conversion.  n_out is the caller's ceil(n_in*up/down). */
int sib_resample(const void* x, int x_dtype, int batch, int n_in, int64_t x_batch_stride, const int32_t* len_in,
                 const float* filt, int up, int down, int taps, int first, float* y, int n_out, int64_t y_batch_stride,
                 const int32_t* len_out /* nullable: y[b, n >= len_out[b]] = 0 */, sib_stream_t stream);
size_t sib_resample_smem_bytes(int down, int taps);
/* 8f row 4: SI-SDR per utterance (I_ea/metrics.py:127-141): a = (eps + <r,e>)/(<r,r> + eps),
 * out[b] = 10 log10((eps + |a r|^2)/(eps + |e - a r|^2)); est/ref [B,n] fp32, lengths nullable, double accumulation in
 * a fixed order (bit-reproducible).  eps = FLT_EPSILON reproduces the reference on float32 arrays. */
int sib_si_sdr_f32(const float* est, const float* ref, int batch, int n, const int32_t* lengths, float eps, float* out,
                   void* workspace /* sib_si_sdr_workspace_bytes(batch) */, sib_stream_t stream);
size_t sib_si_sdr_workspace_bytes(int batch);
/* out[b] = mean_i |a[b,i] - b[b,i]| over n elements per utterance: the reduction of mel-L1
 * (F.l1_loss(y_mel, y_g_hat_mel), I_ea/hifi_gan/train.py:224-227) on two sib_mel_spectrogram_f32 outputs. */
int sib_abs_diff_mean_f32(const float* a, const float* b, int batch, int64_t n, float* out,
                          void* workspace /* sib_abs_diff_workspace_bytes(batch) */, sib_stream_t stream);
size_t sib_abs_diff_workspace_bytes(int batch);

/* ------------------------------------------------------------------------------------------
 * bf16 tensor-core path (tcgen05 + TMA), see DESIGN.md.  Same contract as sib_conv1d_f32 with
 * bf16 x / w / y / residual, fp32 bias and accumulation.  Pre-activation: leaky-relu only (slope in (0, 1]) and only when
 * the kernel runs in its halo mode (stride 1, > 1 evenly spaced taps): the landed A tile is activated in place in shared
 * memory; sib_conv1d_bf16_pre_act_supported() tells - otherwise producers write the activated tensor (y_act) instead.
 * w layout: [groups][c_in/g / cc][n_taps][c_out/g][cc] - one K-major slab per (channel chunk, tap); cc from
 * sib_conv1d_bf16_kblock.
 * Requires c_in/groups % 16 == 0, c_out/groups % 8 == 0.  stride > 1: valid convolution, groups = 1, dense rows, and
 * x readable up to ceil(t_in/stride)*stride rows in the last batch item.
 * ------------------------------------------------------------------------------------------ */
int sib_conv1d_bf16(const sib_conv_desc* d, const void* x, const void* w, const float* bias,
                    const void* residual, void* y, void* y_act /* nullable 2nd output = lrelu(y, act2_slope) */,
                    sib_stream_t stream);
/* Fused ResBlock1 unit (I_ea/hifi_gan/models.py:36-43; I_da/src/models.py ResBlock1.forward) for the narrow stages:
 *   y = (conv2_{k,1}(lrelu(conv1_{k,dilation}(lrelu(x, slope_in)) + b1, slope_mid)) + b2 + x [+ y_old]) * out_scale
 *   y_act = lrelu(y, act2_slope)            (optional second output for an unfused consumer)
 * x, y [B, T, C] frame-major bf16, C in {16, 32, 64}; w1 / w2 in the sib_conv1d_bf16 layout ([k][c_out][c_in]);
 * both convolutions zero-pad ("same"), exactly like the two Conv1d modules.  The intermediate never leaves the SM.
 * y must not alias x.  sib_resunit_bf16_supported() says whether (c, k, dilation) fits (weights stay resident). */
typedef struct sib_resunit_desc {
  int32_t batch, t, c, k, dilation;
  int32_t accumulate;                 /* y += result before scaling (MRF sum, models.py:113-118) */
  float slope_in, slope_mid, out_scale, act2_slope;
  int64_t x_batch_stride, y_batch_stride; /* elements */
  int32_t x_row_stride, y_row_stride;
} sib_resunit_desc;
int sib_resunit_bf16(const sib_resunit_desc* d, const void* x, const void* w1, const float* b1, const void* w2,
                     const float* b2, void* y, void* y_act /* nullable */, sib_stream_t stream);
int sib_resunit_bf16_supported(int c, int k, int dilation, int accumulate, int has_y_act);
int sib_conv1d_bf16_pre_act_supported(const sib_conv_desc* d, int has_residual, int has_y_act);

/* LayerNorm folded into the neighbouring linear layers (HF:388-405 post-LN layers: x1 = LN(t1) feeds FFN-in AND is the
 * residual of FFN-out): the normalised tensor is only stored RAW (t, bf16) together with partial row statistics
 * stats[row][SIB_LN_SLOTS][2] = (sum, sum of squares) over slices of the row.
 *   SIB_LN_APPLY:    y = act(LN(t) W + b) computed as r (t W' - mu s) + c with W' = diag(gamma) W (passed as w),
 *                    s[n] = sum_k w'[k][n] (colsum, of the bf16-rounded w'), c = beta W + b (passed as bias); x = raw t.
 *   SIB_LN_RESIDUAL: y = x W + b + LN(t) with the residual tile = raw t (gamma / beta / stats_in; stats_in null = the
 *                    residual is already normalised); if stats_out is given, the partial statistics of the rows of y.
 * A row's statistics are the sums over ALL slots (unused slots must stay zero: a buffer is always written by the same
 * layer shape).  Linear layers only (one tap, one group, c_out % 64 == 0). */
#define SIB_LN_SLOTS 32
enum { SIB_LN_APPLY = 1, SIB_LN_RESIDUAL = 2 };
typedef struct sib_ln_fold {
  const float* stats_in;
  const float* colsum;
  const float* gamma;
  const float* beta;
  float* stats_out;
  int32_t mode;
  int32_t n_norm; /* length of the normalised rows (hidden size) */
  float eps;
} sib_ln_fold;
int sib_linear_ln_bf16(const sib_conv_desc* d, const sib_ln_fold* ln, const void* x, const void* w, const float* bias,
                       const void* residual, void* y, sib_stream_t stream);

/* Tile-level dataflow between consecutive launches of the transformer loop (HF:388-405 / 525-548: out-proj -> LayerNorm ->
 * FFN-in -> FFN-out -> LayerNorm -> next QKV are all ROW-WISE dependent).  A kernel boundary is a grid-wide barrier: a
 * 75-tile GEMM on 74 SM pairs leaves 73 pairs idle for a whole tile time, and every boundary pays the drain of one
 * kernel plus the pipeline fill of the next.  With a sib_flow the launch does not wait for the previous grid
 * (programmatic dependent launch without griddepcontrol.wait): its CTAs start as SMs become free and read the rows of
 * the 128-row block rb (row / 128) only once wait[rb] has reached the target; whoever writes a 32-row x n-column part
 * of block rb adds n to signal[rb] after the stores have completed (release / acquire at gpu scope).  A full block is
 * therefore worth 4 x N (N = row length of the producer's output); a LayerNorm producer adds N / 32 per VALID row, so
 * the last block's target differs (wait_target_last).  Counters are zeroed by the caller before the chain starts
 * (sib_fill_zero) and hold ceil(rows / 256) * 2 entries.  Arithmetic is untouched: results are bit-identical to the
 * plain launches.  Every input of a launch that carries `wait` must be covered by the flags directly or transitively
 * (weights aside); rows must be flat (batch == 1), linear layers only. */
typedef struct sib_flow {
  const int32_t* wait;   /* nullable: the launch then orders itself after the previous grid as usual */
  int32_t* signal;       /* nullable */
  int32_t wait_target, wait_target_last;
} sib_flow;
int sib_linear_flow_bf16(const sib_conv_desc* d, const sib_flow* flow, const void* x, const void* w, const float* bias,
                         const void* residual, void* y, sib_stream_t stream);
int sib_layernorm_flow_bf16(const void* x, const void* residual, const float* gamma, const float* beta, void* y,
                            int64_t rows, int c, float eps, const sib_flow* flow, sib_stream_t stream);
/* attention inside the chain: the counters run over the FLAT rows b * t + frame.  wait = the QKV projection's counters (an
 * utterance is read once every row block it touches is complete); signal: 2 (= head_dim / 32) per stored row and head, i.e.
 * a block is complete at valid_rows x (heads x head_dim) / 32 like a LayerNorm producer's.  head_dim 64. */
int sib_attention_flow_bf16(const void* qkv, const int32_t* key_len, void* out, int batch, int t, int heads, int head_dim,
                            const sib_flow* flow, sib_stream_t stream);
int sib_fill_zero(void* p, int64_t bytes, sib_stream_t stream);

/* K-block geometry the kernel uses for c_in/groups: cc channels x tb taps per pipeline stage (cc*tb = 64). */
int sib_conv1d_bf16_kblock(int c_in_per_group, int* cc, int* tb);

/* dtype-generic variants of the bandwidth-bound kernels for the bf16 plans (dtypes are sib_dtype; statistics,
 * softmax and accumulation stay fp32).  Same reference lines as their _f32 counterparts above. */
int sib_layernorm(const void* x, int x_dtype, const void* residual, int r_dtype, const float* gamma, const float* beta,
                  void* y, int y_dtype, int64_t rows, int c, float eps, int post_act, sib_stream_t stream);
int sib_attention(const void* qkv, int dtype, const int32_t* key_len, void* out, int batch, int t, int heads,
                  int head_dim, sib_stream_t stream);
int sib_conv0(int mode, const float* wave, int batch, int n_samples, int64_t wave_batch_stride, const float* w,
              const float* bias, int c, int k, int stride, int t0, float* partial, const float* mean, const float* rstd,
              const float* gamma, const float* beta, void* y, int y_dtype, sib_stream_t stream);
int sib_zero_padded_frames(void* h, int dtype, const int32_t* key_len, int batch, int t, int c, sib_stream_t stream);
int sib_conv1d_cout1(const void* x, int x_dtype, const float* w, const float* bias, float* y, int batch, int t, int c,
                     int k, int pad, float pre_slope, int post_act, sib_stream_t stream);
int sib_extend_mel(const float* in, void* out, int out_dtype, int batch, int d, int t, int tm, int frame_major,
                   sib_stream_t stream);
int sib_cast_f32_to_bf16(const float* in, void* out, int64_t n, sib_stream_t stream);
int sib_cast_bf16_to_f32(const void* in, float* out, int64_t n, sib_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SPEECH_INPAINTING_B200_H_ */
