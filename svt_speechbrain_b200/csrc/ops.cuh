// Non-GEMM device ops of the hot path + their host launchers.
#pragma once
#include "common.cuh"

namespace svt {

// ---- attention (attention.cu)
struct AttentionArgs {
  const __nv_bfloat16* q = nullptr;  // row (clip*q_clip_rows + t), col head*head_dim + d
  const __nv_bfloat16* k = nullptr;
  const __nv_bfloat16* v = nullptr;
  __nv_bfloat16* o = nullptr;
  int ldq = 0, ldk = 0, ldv = 0, ldo = 0;
  int Tq = 0, Tk = 0;                   // valid rows per clip
  int q_clip_rows = 0, k_clip_rows = 0; // allocated rows per clip
  int clips = 0, heads = 0, head_dim = 0;
  // WavLM gated relative position bias (HF modeling_wavlm.py, WavLMAttention.forward): the score of (query i, key j) gets
  // gate[(clip * q_clip_rows + i) * heads + head] * rel_tab[head * rel_tab_stride + (j - i) + Tk - 1] added before the
  // softmax.  Self-attention only (Tq == Tk); rel_tab_stride >= 2 * Tk - 1 + 128 with zeros after the 2 * Tk - 1 entries
  // (the tcgen05 kernel reads whole 128-key blocks).
  const float* rel_tab = nullptr;
  int rel_tab_stride = 0;
  const float* gate = nullptr;
};
int attention_bf16(const AttentionArgs& a, cudaStream_t stream);       // dispatcher
int attention_bf16_tc(const AttentionArgs& a, cudaStream_t stream);    // tcgen05 / TMEM kernel (attention_tc.cu), head_dim 64
// process-wide options (svt_set_option): "attention_impl" 0 = auto, 1 = mma.sync kernel, 2 = tcgen05 kernel
int get_option_attention_impl();
int get_option_ln_fold();
int get_option_resid_bf16();  // "resid_bf16": 1 = bf16-only residual stream in the folded pre-LN layers (measurement switch)
int get_option_rowln_fuse();  // "rowln_fuse": 1 = LayerNorm(512) + GELU of the conv feature extractor fused into the conv GEMMs
// development aid: device buffer of 4 x 256 int64 clock stamps written by CTA 0 of the tcgen05 attention kernel
void set_attention_trace_buffer(long long* dev_ptr);

// ---- row ops (rowops.cu)
// y = LayerNorm(x) over the last dim (biased variance), optional GELU afterwards.
// x is fp32 or bf16 (exactly one non-null); writes bf16 and/or fp32.  D % 128 == 0, D <= 2048.
struct LayerNormArgs {
  const float* x_f32 = nullptr;
  const __nv_bfloat16* x_bf16 = nullptr;
  const float* gamma = nullptr;
  const float* beta = nullptr;
  __nv_bfloat16* y_bf16 = nullptr;
  float* y_f32 = nullptr;
  int rows = 0, D = 0;
  float eps = 1e-5f;
  int gelu = 0;
  // optional: accumulate sum / sum-of-squares of the fp32 OUTPUT over rows with (row % clip_rows) < clip_valid
  double* stats = nullptr;  // [2], or [clips][2] when stats_stride == 2 (per-clip statistics)
  int clip_rows = 0, clip_valid = 0;
  int stats_stride = 0;
};
int layer_norm(const LayerNormArgs& a, cudaStream_t stream);
// Folded-LayerNorm helpers (GemmArgs::ln_stats): y = bf16(x) plus per-row (sum, sum of squares) -> stats [rows][2];
// and the per-feature vectors of a folded weight: colsum[n] = sum_k w_packed[n][k], bias[n] += scale * w_f32[n].beta
// WavLM: gate[row][head] = ga * (gb * const[head] - 1) + 2 with (ga, gb) = sigmoid(w2 . x_head + b2), x = rows [M, heads*64]
// bf16, w2 [2][64] / b2 [2] = gru_rel_pos_linear with its 2 x 4 output groups summed (the sums commute with the Linear)
int wavlm_gate(const __nv_bfloat16* x, int rows, int heads, const float* w2, const float* b2, const float* head_const,
               float* gate, cudaStream_t stream);
// the same gate from UN-normalised rows + their folded-LayerNorm statistics (w2g = w2 o gamma per head [heads][2][64], cg its
// row sums [heads][2], dg = w2 . beta + b2 [heads][2])
int wavlm_gate_ln(const __nv_bfloat16* x, const float* stats, int rows, int heads, const float* w2g, const float* cg,
                  const float* dg, const float* head_const, float eps, float* gate, cudaStream_t stream);
// bucket of a relative position (key - query) as WavLMAttention._relative_positions_bucket computes it in fp32 (host)
int wavlm_relative_bucket(int relative_position, int num_buckets, int max_distance);
int row_stats_cast(const float* x, int rows, int D, __nv_bfloat16* y, float* stats, cudaStream_t stream);
int ln_fold_vectors(const float* w_f32, const __nv_bfloat16* w_packed, const float* beta, float scale, int N, int K,
                    float* colsum, float* bias, cudaStream_t stream);

// sum / sum of squares of an fp32 tensor (whole-tensor layer norm, huggingface_interface.py:289)
int tensor_stats(const float* x, size_t n, double* stats /*[2], zeroed here*/, cudaStream_t stream);
// the same per clip: x (clips, n_per_clip) -> stats [clips][2]
int tensor_stats_per_clip(const float* x, int clips, size_t n_per_clip, double* stats, cudaStream_t stream);

// conv layer 0 (C_in = 1): wav (B, L) fp32 -> (B, t_alloc, C) bf16 channel-last.
//   layer mode : (x - mean) * rstd -> conv(k, s) + bias -> LN over C -> GELU          (large)
//   group mode : (x - mean) * rstd -> conv(k, s) [+ bias], pre-norm bf16 + per-(clip, channel) sums (base)
struct Conv0Args {
  const float* wav = nullptr;
  int B = 0, L = 0, T = 0, t_alloc = 0;
  int C = 512, k = 10, stride = 5;
  const float* w = nullptr;     // [k][C] fp32 (tap-major)
  const float* bias = nullptr;  // [C] or null
  const float* gamma = nullptr; // LN affine (layer mode)
  const float* beta = nullptr;
  const double* in_stats = nullptr;  // [2] sum, sumsq over the B*L input (null: no input normalisation)
  int stats_stride = 0;              // 2: in_stats is [B][2], every clip normalised by its own statistics
  __nv_bfloat16* out = nullptr;
  int layer_mode = 1;
  double* chan_stats = nullptr;  // group mode: [B][C][2] sums (zeroed by the launcher)
  // layer mode: tables of the tensor-core kernel (conv0_build_tables; conv0_tables_bytes() bytes, 16-byte aligned).
  // null (or option "conv0_impl" = 1) runs the SIMT kernel.
  const void* tc_tables = nullptr;
};
int conv0_forward(const Conv0Args& a, cudaStream_t stream);
size_t conv0_tables_bytes();
int conv0_build_tables(const float* w_kc, const float* bias, void* tables, cudaStream_t stream);
int get_option_conv0_impl();  // "conv0_impl": 0 auto (tensor-core kernel for layer-norm models), 1 SIMT kernel

// group-norm apply (+GELU) for the base model's layer 0: in-place on (B, t_alloc, C) bf16
int groupnorm_gelu_apply(__nv_bfloat16* x, const double* chan_stats, const float* gamma, const float* beta, int B, int T,
                         int t_alloc, int C, cudaStream_t stream);

// Output whole-tensor norm + head + frame post-processing:
//   feats[b,t,:] = (x[b*clip_rows+t,:] - mean) * rstd    (mean/rstd from stats over valid rows; identity if !output_norm)
//   logits[b,t,:] = feats . Wh^T + bh                     (n_out <= 32)
struct HeadArgs {
  const float* x = nullptr;  // [clips*clip_rows, D] fp32
  int clips = 0, clip_rows = 0, T = 0, D = 0;
  const double* stats = nullptr;  // [2] or null (no output norm); [clips][2] when stats_stride == 2
  int stats_stride = 0;
  float eps = 1e-5f;
  const float* w = nullptr;  // [n_out, D] fp32 (null: no head)
  const float* b = nullptr;
  int n_out = 0;
  float* feats = nullptr;   // [clips, T, D] compact, or null
  float* logits = nullptr;  // [clips, T, n_out] compact, or null
};
int head_forward(const HeadArgs& a, cudaStream_t stream);

// per-frame argmax over the octave / pitch-class logits (train_audio_ssl.py:95-99); first max wins
int frame_argmax(const float* logits, int n_frames, int n_out, int oct_off, int n_oct, int pc_off, int n_pc, int32_t* oct,
                 int32_t* pc, cudaStream_t stream);

// ---- weight packing helpers (pack.cu): generic strided fp32 -> bf16/fp32 gather
//   dst[i0][i1][i2][i3] (contiguous) = scale * src[i0*s0 + i1*s1 + i2*s2 + i3*s3] * (vec ? vec[i_vecdim] : 1)
struct PackArgs {
  const float* src = nullptr;
  int dims[4] = {1, 1, 1, 1};
  long long strides[4] = {0, 0, 0, 0};
  float scale = 1.f;
  const float* vec = nullptr;
  int vec_dim = 0;
};
int pack_bf16(const PackArgs& a, __nv_bfloat16* dst, cudaStream_t stream);
int pack_f32(const PackArgs& a, float* dst, cudaStream_t stream);
// weight-norm(dim=2) factor g[j] / ||v[:,:,j]||  for the positional conv (HF modeling_wav2vec2.py:343-355)
int weight_norm_scale(const float* v, const float* g, int d0, int d1, int taps, float* out, cudaStream_t stream);
// x[b,t,:] += pe[t,:]  (sinusoidal table computed on the fly) -> bf16 + fp32 copies; zero rows t >= T_valid of src
int add_positional_encoding(const float* x, int clips, int T_src, int T, int D, float* out_f32, __nv_bfloat16* out_bf16,
                            cudaStream_t stream);
// out = a + b (fp32, n elements)
int add_f32(const float* a, const float* b, float* out, size_t n, cudaStream_t stream);
int cast_f32_to_bf16(const float* x, __nv_bfloat16* y, size_t n, cudaStream_t stream);
// y[row][c] = bf16(x[row][c] * scale[c] + shift[c]): eval-mode BatchNorm1d over channel-last rows
int channel_affine_bf16(const float* x, const float* scale, const float* shift, __nv_bfloat16* y, size_t rows, int D,
                        cudaStream_t stream);

}  // namespace svt
