// tcgen05 / TMEM / TMA GEMM used by every dense contraction on the hot path
// (conv-as-GEMM, feature projection, QKV / out-proj / FFN, positional conv, fusion projections).
#pragma once
#include "common.cuh"

namespace svt {

enum GemmAct : int { kActNone = 0, kActGelu = 1, kActRelu = 2, kActPRelu = 3 };
constexpr int kMaxGemmTaps = 9;
// below this many 256 x 256 work items the one-CTA kernel with 128-column tiles is used (see gemm_bf16_tc)
constexpr long long kSmallProblemPairUnits = 33;  // measured: profiles/r1_small_m_gemm_sweep.txt (tools/kernel_bench.py smallm)

// C[row, n] = act( sum_k A[row, k] * W[n, k] + bias[n] ) (+ resid[row, n])
//
// A (bf16) is described as a 3-D TMA view so that the same kernel covers
//   * plain row-major activations            dims (K, 1, M)
//   * strided conv1d as implicit GEMM         dims (C_in, taps, M) with row stride = conv_stride * C_in
//     (channel-last activations: the `taps` input rows of one output frame are contiguous in HBM, so
//      im2col is nothing but an overlapping row stride -- no data is materialised)
//   * the grouped positional conv             dims (D, T, B), one (clip, 128-frame) tile per CTA tile and
//     one tap per k-block, the tap shift and the zero padding both done by TMA coordinates / OOB fill.
// W (bf16, K-major, i.e. the nn.Linear layout [N, K]) is a 2-D TMA view.
struct GemmArgs {
  // ---- operands
  const __nv_bfloat16* a = nullptr;
  uint64_t a_dims[3] = {0, 1, 0};          // elements
  uint64_t a_strides[2] = {0, 0};          // elements, for dims 1 and 2
  const __nv_bfloat16* w = nullptr;        // [w_rows, K_w] row-major
  int w_rows = 0;                          // total rows of the weight view
  int w_cols = 0;                          // inner (K) extent of the weight view
  // ---- problem
  int mode = 0;                            // 0 linear, 1 positional conv
  int M = 0;                               // linear: flat rows. posconv: unused
  int N = 0;                               // output columns
  int K = 0;                               // contraction length (posconv: taps * 64)
  int k_inner = 0;                         // linear: a_dims[0]
  // posconv geometry
  int n_clips = 1, clip_rows = 0 /*T_alloc*/, clip_valid = 0 /*T*/, pad_left = 0, taps = 0;
  int group_size = 64;                     // channels per group (weights zero-padded to 64 x 64 per tap)
  // ---- epilogue
  const float* bias = nullptr;             // [N] fp32 or null
  const float* resid = nullptr;            // fp32 [rows, ld_resid] or null (may alias out_f32)
  float* out_f32 = nullptr;
  __nv_bfloat16* out_bf16 = nullptr;
  int ld_out = 0;                          // leading dimension (elements) of out_f32 / out_bf16 / resid
  int act = kActNone;
  const float* alpha = nullptr;            // [N] per-column slope for kActPRelu
  const __nv_bfloat16* resid_bf16 = nullptr;  // bf16 residual [rows, ld_out] added before the activation (ResNet blocks)
  const uint8_t* row_mask = nullptr;       // [rows]: rows with mask 0 are written as zeros (padding ring of a feature map)
  // LayerNorm folded around the GEMM (pre-LN transformer layers).  The GEMM that PRODUCES the residual stream adds
  // (sum, sum of squares) of every 128-column slice of its output rows to row_stats_out [M][N / 128][2] (one writer
  // per slot, so deterministic and nothing to zero); the GEMM that CONSUMES LN(x) (ln_stats, same layout) multiplies the
  // un-normalised bf16 rows by W' = gamma o W and finishes y = rstd * (x.W' - mean * colsum) + bias in its epilogue,
  // with colsum[n] = sum_k W'[n][k] and bias[n] = beta.W[n] + b[n] (bf16-only outputs, LN width = K).
  float* row_stats_out = nullptr;
  const float* ln_stats = nullptr;
  const float* ln_colsum = nullptr;
  float ln_eps = 1e-5f;
  // Row LayerNorm over ALL N output columns (+ optional GELU) fused behind the GEMM (conv feature-extractor layers of the
  // layer-norm models, HF:275-299: conv -> LN(512) -> GELU): y = act(LN(x.W^T + bias) * gamma + beta), bf16 output only.
  // Built for N = 512 in the CTA-pair kernel (gemm_rowln_supported): a pair computes both 256-column tiles of a row
  // block back to back, writes them un-normalised (bf16, they stay in L2) with per-row statistics kept in shared
  // memory, then normalises the block in place -- the rows never make an extra round trip through HBM.
  const float* rowln_gamma = nullptr;      // [N]
  const float* rowln_beta = nullptr;       // [N]
  float rowln_eps = 1e-5f;
  int rowln_gelu = 0;
  // ---- shifted-row taps (linear mode only): k-block kb reads A rows (row + tap_off[kb / (k_inner / 64)]), columns
  // (kb % (k_inner / 64)) * 64 .. + 64; K = n_taps * k_inner.  This is a 2-D convolution over feature maps stored as
  // flat rows [frame][y][x] with a zero padding ring: tap (dy, dx) is the constant row offset dy * W_padded + dx.
  int n_taps = 0;
  int tap_off[kMaxGemmTaps] = {0};
  // optional second bf16 output written transposed per clip is not needed (attention reads row-major V)
};

int gemm_bf16_tc(const GemmArgs& g, cudaStream_t stream);
// true when g (with rowln_gamma set) can run with the LayerNorm fused; otherwise run the GEMM plain + layer_norm()
bool gemm_rowln_supported(const GemmArgs& g);

// bf16 tiled tensor map with 128-byte swizzle; strides in elements for dims 1..rank-1
int encode_bf16_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                    const uint32_t* box);

}  // namespace svt
