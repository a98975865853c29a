/* svt_b200.h -- C ABI of the B200-native AMT inference hot path (libsvt_b200.so).
 *
 * Plain C, no torch types: device pointers are raw CUDA device addresses (e.g. tensor.data_ptr()),
 * `stream` is a cudaStream_t passed as void*.  Every function returns 0 on success or a non-zero
 * status; svt_last_error() returns the message of the calling thread's last failure.  There is no CPU
 * fallback anywhere: without a CUDA device the compute entry points fail with SVT_ERR_NO_DEVICE.
 *
 * Each entry point replaces one interface of guxm2021/SVT_SpeechBrain (paths relative to the
 * reference root; "HF:" = transformers/models/wav2vec2/modeling_wav2vec2.py):
 *
 *   svt_encoder_*          HuggingFaceWav2Vec2.__init__/forward      MIR_ST500/huggingface_interface.py:89-179,263-298
 *                          (Wav2Vec2Model.forward HF:1327-1383) + the AMT head
 *                          speechbrain.nnet.linear.Linear             speechbrain/nnet/linear.py:41-76
 *                          as applied in AMT.compute_forward          MIR_ST500/train_audio_ssl.py:28-48
 *   svt_fusion_*           FusionRCA.__init__/forward                 N20EMv2/audio_visual/fusion.py:186-210
 *   svt_frame_postproc     per-frame sigmoid inputs / argmax          MIR_ST500/train_audio_ssl.py:93-100
 *   svt_frame2note         frame2note                                 MIR_ST500/utils.py:82-149
 *   svt_op_*               kernel-level hooks (GEMM = nn.Linear / nn.Conv1d, attention = HF:438-463,
 *                          layer norm = nn.LayerNorm) exported so the parity tests bisect through the same ABI.
 */
#ifndef SVT_B200_H_
#define SVT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVT_OK 0
#define SVT_ERR_INVALID_ARGUMENT 1
#define SVT_ERR_CUDA 2
#define SVT_ERR_NOT_FINALIZED 3
#define SVT_ERR_UNKNOWN_TENSOR 4
#define SVT_ERR_WORKSPACE_TOO_SMALL 5
#define SVT_ERR_UNSUPPORTED 6
#define SVT_ERR_NO_DEVICE 7

#define SVT_MAX_CONV_LAYERS 8

int svt_version(void);
const char* svt_last_error(void);
/* number of CUDA devices visible (0 on a CPU-only box); never fails */
int svt_device_count(void);
/* kernels launched by this library since it was loaded (all streams); used by bench.py's gpu_launches */
long long svt_debug_launch_count(void);

/* process-wide tuning / test switches.  "attention_impl": 0 auto (default), 1 force the mma.sync kernel,
 * 2 force the tcgen05/TMEM kernel (head_dim 64 only).  "gemm_impl": 0 auto (CTA-pair cta_group::2 kernel when
 * N % 256 == 0 and the problem has enough tiles, else the one-CTA kernel), 1 force the one-CTA kernel, 2 one-CTA kernel with
 * 128-column tiles, 3 force the CTA-pair kernel.  "ln_fold": 1 (default) folds the two per-layer LayerNorms of pre-LN
 * (stable_layer_norm) transformer layers into the neighbouring GEMMs, 0 runs them as separate kernels.  "rowln_fuse": 1
 * (default) runs conv -> LayerNorm(512) -> GELU of the layer-norm feature extractors as one kernel per layer
 * (svt_op_gemm_rowln), 0 as a GEMM followed by a LayerNorm kernel.  "conv0_impl": 0 (default) tensor-core first conv layer for
 * layer-norm models, 1 the SIMT kernel.  "resid_epilogue": which per-configuration epilogue instantiations of the GEMM kernels
 * may run: 2 (default) all, 1 only the residual GEMMs', 0 the generic kernel (identical results).  "resid_bf16": 1 keeps the
 * residual stream of the folded pre-LN layers in bf16 only (faster, logits error x 1.6; default 0). */
int svt_set_option(const char* name, int value);

/* development aid: device buffer of 4 x 256 int64 that CTA 0 of the tcgen05 attention kernel fills with clock64()
 * stamps of its pipeline phases (NULL switches it off, the default); see tools/attention_trace.py */
void svt_debug_attention_trace(void* dev_buffer_8k);

/* ------------------------------------------------------------------ wav2vec2-style SSL encoder + head */
typedef struct svt_encoder svt_encoder;

/* Mirrors the fields of HF Wav2Vec2Config that the forward depends on (configuration_wav2vec2.py). */
typedef struct svt_encoder_config {
  int hidden_size;         /* D: 1024 large / 768 base */
  int num_layers;          /* 24 / 12 */
  int num_heads;           /* 16 / 12 (head dim must be 64 or 128) */
  int ffn_size;            /* 4096 / 3072 */
  int num_conv_layers;     /* 7 */
  int conv_dim;            /* 512 (all layers) */
  int conv_kernel[SVT_MAX_CONV_LAYERS]; /* 10,3,3,3,3,2,2 */
  int conv_stride[SVT_MAX_CONV_LAYERS]; /* 5,2,2,2,2,2,2 */
  int conv_bias;           /* config.conv_bias */
  int feat_norm_layer;     /* 1: feat_extract_norm == "layer" (large); 0: "group" (base) */
  int stable_layer_norm;   /* config.do_stable_layer_norm */
  int pos_conv_kernel;     /* 128 */
  int pos_conv_groups;     /* 16 */
  float layer_norm_eps;    /* 1e-5 */
  int normalize_wav;       /* lobe.normalize_wav: F.layer_norm(wav, wav.shape), huggingface_interface.py:288-289 */
  int output_norm;         /* lobe.output_norm:  F.layer_norm(out, out.shape),  huggingface_interface.py:295-296 */
  int feat_proj_norm;      /* 1: LayerNorm before the feature projection (wav2vec2, HuBERT large); 0: none (HuBERT base,
                              HubertConfig.feat_proj_layer_norm = False) */
  int pos_conv_layers;     /* 0: one weight-normed grouped conv + GELU (wav2vec2, HuBERT).  n > 0: data2vec-audio's stack
                              of n [grouped conv (kernel pos_conv_kernel, plain weights) -> LayerNorm without affine ->
                              GELU] (Data2VecAudioPositionalConvEmbedding; n = 5, kernel 19) */
  int rel_pos_buckets;     /* 0: none.  > 0: WavLM's gated relative position bias (WavLMConfig.num_buckets = 320) */
  int rel_pos_max_distance; /* WavLMConfig.max_bucket_distance = 800 */
  int pos_conv_batch_norm; /* HubertConfig.conv_pos_batch_norm: eval-mode BatchNorm1d before a plain (not weight-normed)
                              positional conv (HF modeling_hubert.py, HubertPositionalConvEmbedding) */
} svt_encoder_config;

int svt_encoder_create(const svt_encoder_config* cfg, svt_encoder** out);
void svt_encoder_destroy(svt_encoder* enc);
/* Register one fp32 HOST tensor under its reference state_dict name ("model." prefix optional), e.g.
 * "model.encoder.layers.3.attention.q_proj.weight".  Both weight-norm spellings of the positional conv are
 * accepted (parametrizations.weight.original{0,1} / weight_g, weight_v).  Unknown names are ignored
 * (e.g. masked_spec_embed) and reported by return value SVT_ERR_UNKNOWN_TENSOR only if `strict` != 0. */
int svt_encoder_set_tensor(svt_encoder* enc, const char* name, const float* host, const int64_t* shape, int ndim,
                           int strict);
/* AMT head = speechbrain Linear: w (n_out, D) fp32 host, b (n_out) or NULL; n_out <= 32 */
int svt_encoder_set_head(svt_encoder* enc, const float* w, const float* b, int n_out);
/* Scope of the two whole-tensor layer norms (huggingface_interface.py:288-289,295-296).  0 (default): one mean / variance
 * over the whole call, exactly what one reference forward of a (B, L) batch computes.  1: one per clip, i.e. what B
 * separate reference calls of batch size 1 compute -- the reference's evaluation loop (train_audio_ssl.py:85-90 asserts
 * batch size 1), so a batch of utterances reproduces that loop in one call. */
int svt_encoder_set_norm_per_clip(svt_encoder* enc, int per_clip);
/* Pack into kernel layouts on the device (bf16 cast, conv weights tap-major, QKV concatenation with the
 * d_h^-0.5 scale folded into q, weight-norm recomposition).  May be called again after set_tensor. */
int svt_encoder_finalize(svt_encoder* enc);
/* frames produced for L samples (no padding), e.g. 160000 -> 499 */
int svt_encoder_num_frames(const svt_encoder* enc, int n_samples);
size_t svt_encoder_workspace_bytes(const svt_encoder* enc, int batch, int n_samples);
/* wav_dev: (B, L) fp32 device.  feats_dev: (B, T, D) fp32 device or NULL.  logits_dev: (B, T, n_out) fp32
 * device or NULL (requires a head).  The whole-tensor norms span exactly this call's B clips, as in the
 * reference.  Asynchronous on `stream`. */
int svt_encoder_forward(svt_encoder* enc, const float* wav_dev, int batch, int n_samples, void* workspace_dev,
                        size_t workspace_bytes, float* feats_dev, float* logits_dev, void* stream);
/* End-to-end convenience used by the host-buffer benchmark leg: pinned/pageable HOST wav (B, L) -> H2D ->
 * forward -> D2H logits (B, T, n_out) into host memory; synchronises `stream` before returning. */
int svt_encoder_forward_host(svt_encoder* enc, const float* wav_host, int batch, int n_samples, void* workspace_dev,
                             size_t workspace_bytes, float* wav_stage_dev, float* logits_stage_dev,
                             float* logits_host, void* stream);

/* ------------------------------------------------------------------ host-buffer serving pipeline
 * The reference's evaluation loop hands one HOST batch at a time to the model (speechbrain/core.py:1221-1226).  A
 * pipeline keeps `depth` batches in flight: the H2D copy of batch k + 1 runs on a copy stream under the forward of
 * batch k, the D2H of its logits follows on the compute stream.  Memory is the caller's: workspace as for
 * svt_encoder_forward, wav_stage_dev depth x (batch x n_samples) floats, logits_stage_dev depth x (batch x T x n_out)
 * floats; host buffers must be pinned and stay valid until svt_pipeline_wait(ticket) returns.  Needs a head. */
typedef struct svt_pipeline svt_pipeline;
int svt_pipeline_create(svt_encoder* enc, int batch, int n_samples, int depth, void* workspace_dev, size_t workspace_bytes,
                        float* wav_stage_dev, float* logits_stage_dev, svt_pipeline** out);
void svt_pipeline_destroy(svt_pipeline* p);
int svt_pipeline_submit(svt_pipeline* p, const float* wav_host_pinned, float* logits_host_pinned, long long* ticket);
int svt_pipeline_wait(svt_pipeline* p, long long ticket);

/* ------------------------------------------------------------------ residual cross-attention fusion */
typedef struct svt_fusion svt_fusion;
typedef struct svt_fusion_config {
  int d_model; /* 1024 */
  int nhead;   /* 8  (head dim 128) */
  int d_ffn;   /* 3072 */
  float alpha; /* 0.5 */
} svt_fusion_config;

int svt_fusion_create(const svt_fusion_config* cfg, svt_fusion** out);
void svt_fusion_destroy(svt_fusion* f);
/* names as in FusionRCA.state_dict(): "fusion.layer1.self_att.att.in_proj_weight", ...; the `pe` buffer is
 * recomputed on the device and ignored if passed. */
int svt_fusion_set_tensor(svt_fusion* f, const char* name, const float* host, const int64_t* shape, int ndim,
                          int strict);
int svt_fusion_finalize(svt_fusion* f);
size_t svt_fusion_workspace_bytes(const svt_fusion* f, int batch, int t_audio);
/* audio_dev (B, Ta, D), video_dev (B, Tv, D) fp32 device -> out_dev (B, Ta, D) fp32 device. */
int svt_fusion_forward(svt_fusion* f, const float* audio_dev, const float* video_dev, int batch, int t_audio,
                       int t_video, void* workspace_dev, size_t workspace_bytes, float* out_dev, void* stream);

/* ------------------------------------------------------------------ AV-HuBERT video stream (lip ROIs -> features) */
/* Replaces FairseqAVHubertPretrain.forward({"video": x, "audio": None})  N20EMv2/video_only/fairseq_interface.py:454-485
 * = AVHubertModel.extract_finetune (hubert.py:688-739): ResEncoder (resnet.py:133-171) -> Linear 512->D -> concat with a
 * zero audio stream -> LayerNorm(2D) -> Linear 2D->D -> fairseq TransformerEncoder (layer_norm_first) -> output LN. */
typedef struct svt_video svt_video;
typedef struct svt_video_config {
  int embed_dim;        /* cfg.encoder_embed_dim: 1024 (large) */
  int num_layers;       /* cfg.encoder_layers: 24 */
  int num_heads;        /* cfg.encoder_attention_heads: 16 */
  int ffn_size;         /* cfg.encoder_ffn_embed_dim: 4096 */
  int conv_pos;         /* cfg.conv_pos: 128 */
  int conv_pos_groups;  /* cfg.conv_pos_groups: 16 */
  float layer_norm_eps; /* 1e-5 */
  int input_norm;       /* wrapper's input_norm: F.layer_norm(video, video.shape), fairseq_interface.py:473-474 */
  int output_norm;      /* wrapper's output_norm: F.layer_norm(out, out.shape),     fairseq_interface.py:482-483 */
} svt_video_config;

int svt_video_create(const svt_video_config* cfg, svt_video** out);
void svt_video_destroy(svt_video* v);
/* fp32 HOST tensors under the reference module's state_dict names ("model." prefix optional):
 * feature_extractor_video.resnet.{frontend3D,trunk}.*, feature_extractor_video.proj.*, layer_norm.*, post_extract_proj.*,
 * encoder.{pos_conv.0.*, layers.N.{self_attn.*, self_attn_layer_norm.*, fc1.*, fc2.*, final_layer_norm.*}, layer_norm.*}.
 * Tensors the video-only forward never reads (feature_extractor_audio.*, mask_emb, final_proj, ...) are ignored. */
int svt_video_set_tensor(svt_video* v, const char* name, const float* host, const int64_t* shape, int ndim, int strict);
/* folds every BatchNorm into its conv, packs bf16 kernel layouts */
int svt_video_finalize(svt_video* v);
size_t svt_video_workspace_bytes(const svt_video* v, int batch, int n_frames);
/* video_dev: (B, 1, T, 88, 88) fp32 device (already cropped / normalised as in video_only/train_video_ssl.py:445-457);
 * feats_dev: (B, T, D) fp32 device.  Asynchronous on `stream`. */
int svt_video_forward(svt_video* v, const float* video_dev, int batch, int n_frames, void* workspace_dev, size_t workspace_bytes,
                      float* feats_dev, void* stream);

/* Evaluation transform of the video recipe (video_only/train_video_ssl.py:454-457, utils.py:45-84) on the device:
 * uint8 grey frames (n_frames, height, width) -> x / 255 -> CenterCrop(crop) -> (x - mean) / stdev, fp32
 * (n_frames, crop, crop) = the (B, 1, T, 88, 88) input of svt_video_forward (crop 88, mean 0.421, stdev 0.165). */
int svt_video_transform_u8(const uint8_t* frames_dev, long long n_frames, int height, int width, int crop, float mean,
                           float stdev, float* out_dev, void* stream);

/* HOST helper: bucket of a relative position (key index - query index) exactly as WavLMAttention._relative_positions_bucket
 * computes it (fp32), exported so the CPU tests can pin it against torch. */
int svt_wavlm_relative_bucket(int relative_position, int num_buckets, int max_distance);

/* ------------------------------------------------------------------ frame post-processing + note decoding */
/* logits_dev (n_frames, n_out) fp32 device -> octave / pitch-class argmax (first maximum wins, torch
 * semantics) into int32 device arrays.  Columns [oct_off, oct_off+n_oct) and [pc_off, pc_off+n_pc). */
int svt_frame_postproc(const float* logits_dev, int n_frames, int n_out, int oct_off, int n_oct, int pc_off, int n_pc,
                       int32_t* oct_dev, int32_t* pc_dev, void* stream);
/* HOST function (the reference decodes on the host too).  p_on / p_off: fp32 sigmoid outputs, oct / pc: class
 * ids.  Writes up to max_notes rows of [onset_s, offset_s, midi] (float64) and *n_notes.  Comparisons are
 * done in fp32 against (float)threshold, pitch mode ties follow CPython's set iteration order, time stamps
 * are frame_size * i in float64 -- bit-exact with the reference.  n_frames == 1 with an onset above threshold
 * reproduces the reference's ValueError as SVT_ERR_INVALID_ARGUMENT. */
int svt_frame2note(const float* p_on, const float* p_off, const int32_t* oct, const int32_t* pc, int n_frames,
                   double onset_thres, double offset_thres, double frame_size, double* notes_out, int max_notes,
                   int* n_notes);

/* ------------------------------------------------------------------ kernel-level hooks (device pointers) */
/* C[M,N] = act(A[M,K] W[N,K]^T + bias) (+ resid); A rows start every a_row_stride elements (a_row_stride < K
 * gives the overlapping-row implicit conv1d view; k_inner = contiguous run, K % k_inner == 0).
 * act: 0 none, 1 GELU(erf), 2 ReLU.  Exactly as used by the encoder. */
int svt_op_gemm(const void* a_bf16, long long a_row_stride, int k_inner, const void* w_bf16, const float* bias,
                const float* resid, float* out_f32, void* out_bf16, int M, int N, int K, int ld_out, int act,
                void* stream);
/* conv feature-extractor layer of the layer-norm models (HF modeling_wav2vec2.py:275-299: conv -> LayerNorm(512) -> GELU) in
 * ONE kernel: out = act(LN(A W^T + bias) * gamma + beta), N = 512, bf16 rows normalised in place while still in L2
 * (option "rowln_fuse" chooses between this and svt_op_gemm + svt_op_layer_norm inside the encoder). */
int svt_op_gemm_rowln(const void* a_bf16, long long a_row_stride, int k_inner, const void* w_bf16, const float* bias,
                      const float* gamma, const float* beta, float eps, int gelu, void* out_bf16, int M, int N, int K,
                      void* stream);
/* The two halves of a LayerNorm folded around GEMMs (pre-LN transformer layers, option "ln_fold").  Producer
 * (row_stats_out != NULL, out_f32 != NULL, N % 256 == 0): as svt_op_gemm, and row_stats_out[row][j] = (sum, sum of
 * squares) of columns 128 j .. 128 j + 127 of the fp32 output row.  Consumer (ln_stats != NULL): a holds un-normalised rows x, w = W o gamma,
 * colsum[n] = sum_k w[n][k], bias = beta.W + b; out_bf16 = act(rstd * (x.w - mean * colsum) + bias) with mean / rstd
 * from ln_stats[row][0 .. K / 128) (K in {256, 512, 768, 1024}).  svt_op_row_stats_cast: y = bf16(x) and
 * stats[row][D / 128][2] written. */
int svt_op_gemm_ln(const void* a_bf16, const void* w_bf16, const float* bias, const float* colsum, const float* ln_stats,
                   float ln_eps, float* row_stats_out, const float* resid, float* out_f32, void* out_bf16, int M, int N,
                   int K, int act, void* stream);
int svt_op_row_stats_cast(const float* x, int rows, int D, void* y_bf16, float* stats, void* stream);
/* grouped "same"-padded conv1d over time as used for the positional embedding: x (clips, clip_rows, D) bf16,
 * w packed [G][taps][64][64] bf16, out_f32[row, :] = resid + gelu(conv + bias) for valid rows. */
int svt_op_posconv(const void* x_bf16, const void* w_packed, const float* bias, const float* resid, float* out_f32,
                   int clips, int clip_rows, int t_valid, int D, int groups, int taps, void* stream);
/* pack a (D, D/G, taps) fp32 device conv weight (already weight-norm recomposed) into [G][taps][64][64] bf16 */
int svt_op_pack_posconv(const float* w_f32_dev, int D, int groups, int taps, void* out_bf16, void* stream);
int svt_op_attention(const void* q, const void* k, const void* v, void* o, int ldq, int ldk, int ldv, int ldo, int Tq,
                     int Tk, int q_clip_rows, int k_clip_rows, int clips, int heads, int head_dim, void* stream);
int svt_op_layer_norm(const float* x_f32, const void* x_bf16, const float* gamma, const float* beta, void* y_bf16,
                      float* y_f32, int rows, int D, float eps, int gelu, void* stream);
/* y[rows, n_out] = x[rows, D] W[n_out, D]^T + b, all fp32 device, n_out <= 32: the AMT head
 * (speechbrain.nnet.linear.Linear with n_neurons = 20) as a standalone op. */
int svt_op_linear_small(const float* x, int rows, int D, const float* w, const float* b, int n_out, float* y,
                        void* stream);
int svt_op_conv0(const float* wav, int B, int L, const float* w_kc, const float* bias, const float* gamma,
                 const float* beta, int normalize, void* out_bf16, int t_alloc, double* stats_scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SVT_B200_H_ */
