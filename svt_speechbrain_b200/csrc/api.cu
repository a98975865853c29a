// C-ABI glue: error state, device probing, kernel-level hooks.
#include <atomic>
#include <mutex>

#include "gemm_epilogue.cuh"
#include "model.cuh"

namespace svt {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
static std::atomic<long long> g_launches{0};
void note_kernel_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(); }
static std::atomic<int> g_attention_impl{0};
static std::atomic<int> g_gemm_impl{0};
static std::atomic<int> g_ln_fold{1};
static std::atomic<int> g_rowln_fuse{1};
static std::atomic<int> g_conv0_impl{0};
static std::atomic<int> g_resid_bf16{0};
static std::atomic<int> g_resid_epilogue{2};
int get_option_resid_epilogue() { return g_resid_epilogue.load(std::memory_order_relaxed); }
int get_option_resid_bf16() { return g_resid_bf16.load(std::memory_order_relaxed); }
int get_option_conv0_impl() { return g_conv0_impl.load(std::memory_order_relaxed); }
int get_option_ln_fold() { return g_ln_fold.load(std::memory_order_relaxed); }
int get_option_rowln_fuse() { return g_rowln_fuse.load(std::memory_order_relaxed); }
int get_option_gemm_impl() { return g_gemm_impl.load(std::memory_order_relaxed); }
int get_option_attention_impl() { return g_attention_impl.load(std::memory_order_relaxed); }
int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

}  // namespace svt

using namespace svt;

extern "C" {

int svt_version(void) { return 100; }
long long svt_debug_launch_count(void) { return launch_count(); }
const char* svt_last_error(void) { return g_last_error.c_str(); }
int svt_set_option(const char* name, int value) {
  if (name == nullptr) return fail(kInvalidArgument, "null argument");
  const std::string n(name);
  if (n == "attention_impl") {
    if (value < 0 || value > 2) return fail(kInvalidArgument, "attention_impl must be 0 (auto), 1 (mma.sync) or 2 (tcgen05)");
    g_attention_impl.store(value);
    return kOk;
  }
  if (n == "gemm_impl") {
    if (value < 0 || value > 3)
      return fail(kInvalidArgument, "gemm_impl must be 0 (auto), 1 (one-CTA kernel), 2 (one-CTA kernel, 128-column tiles) or 3 (CTA-pair kernel)");
    g_gemm_impl.store(value);
    return kOk;
  }
  if (n == "ln_fold") {
    if (value < 0 || value > 1) return fail(kInvalidArgument, "ln_fold must be 0 (separate LayerNorm kernels) or 1 (folded)");
    g_ln_fold.store(value);
    return kOk;
  }
  if (n == "resid_epilogue") {
    if (value < 0 || value > 2)
      return fail(kInvalidArgument, "resid_epilogue must be 0 (generic epilogue kernel), 1 (+ specialised residual epilogue) or 2 (+ specialised QKV / FFN-1 epilogues)");
    g_resid_epilogue.store(value);
    return kOk;
  }
  if (n == "resid_bf16") {
    if (value < 0 || value > 1) return fail(kInvalidArgument, "resid_bf16 must be 0 (fp32 residual stream) or 1 (bf16 only; measurement switch)");
    g_resid_bf16.store(value);
    return kOk;
  }
  if (n == "conv0_impl") {
    if (value < 0 || value > 1) return fail(kInvalidArgument, "conv0_impl must be 0 (auto: tensor-core kernel for layer-norm models) or 1 (SIMT kernel)");
    g_conv0_impl.store(value);
    return kOk;
  }
  if (n == "rowln_fuse") {
    if (value < 0 || value > 1) return fail(kInvalidArgument, "rowln_fuse must be 0 (conv GEMM + separate LayerNorm kernel) or 1 (fused)");
    g_rowln_fuse.store(value);
    return kOk;
  }
  return fail(kInvalidArgument, "unknown option " + n);
}
void svt_debug_attention_trace(void* dev_buffer_8k) { set_attention_trace_buffer(static_cast<long long*>(dev_buffer_8k)); }
int svt_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int svt_frame_postproc(const float* logits_dev, int n_frames, int n_out, int oct_off, int n_oct, int pc_off, int n_pc,
                       int32_t* oct_dev, int32_t* pc_dev, void* stream) {
  if (logits_dev == nullptr || oct_dev == nullptr || pc_dev == nullptr) return fail(kInvalidArgument, "null argument");
  if (oct_off < 0 || pc_off < 0 || n_oct <= 0 || n_pc <= 0 || oct_off + n_oct > n_out || pc_off + n_pc > n_out)
    return fail(kInvalidArgument, "frame_postproc: column ranges outside the logits");
  return frame_argmax(logits_dev, n_frames, n_out, oct_off, n_oct, pc_off, n_pc, oct_dev, pc_dev,
                      static_cast<cudaStream_t>(stream));
}

int svt_op_gemm(const void* a_bf16, long long a_row_stride, int k_inner, const void* w_bf16, const float* bias,
                const float* resid, float* out_f32, void* out_bf16, int M, int N, int K, int ld_out, int act,
                void* stream) {
  if (a_bf16 == nullptr || w_bf16 == nullptr) return fail(kInvalidArgument, "null argument");
  if (k_inner <= 0 || K % k_inner != 0) return fail(kInvalidArgument, "K must be a multiple of k_inner");
  GemmArgs g;
  g.a = static_cast<const __nv_bfloat16*>(a_bf16);
  g.a_dims[0] = k_inner; g.a_dims[1] = K / k_inner; g.a_dims[2] = M;
  g.a_strides[0] = k_inner; g.a_strides[1] = static_cast<uint64_t>(a_row_stride);
  g.w = static_cast<const __nv_bfloat16*>(w_bf16); g.w_rows = N; g.w_cols = K;
  g.M = M; g.N = N; g.K = K; g.k_inner = k_inner;
  g.bias = bias; g.resid = resid; g.out_f32 = out_f32; g.out_bf16 = static_cast<__nv_bfloat16*>(out_bf16);
  g.ld_out = ld_out; g.act = act;
  return gemm_bf16_tc(g, static_cast<cudaStream_t>(stream));
}

int svt_op_gemm_rowln(const void* a_bf16, long long a_row_stride, int k_inner, const void* w_bf16, const float* bias,
                      const float* gamma, const float* beta, float eps, int gelu, void* out_bf16, int M, int N, int K,
                      void* stream) {
  if (a_bf16 == nullptr || w_bf16 == nullptr || gamma == nullptr || beta == nullptr || out_bf16 == nullptr)
    return fail(kInvalidArgument, "null argument");
  if (k_inner <= 0 || K % k_inner != 0) return fail(kInvalidArgument, "K must be a multiple of k_inner");
  GemmArgs g;
  g.a = static_cast<const __nv_bfloat16*>(a_bf16);
  g.a_dims[0] = k_inner; g.a_dims[1] = K / k_inner; g.a_dims[2] = M;
  g.a_strides[0] = k_inner; g.a_strides[1] = static_cast<uint64_t>(a_row_stride);
  g.w = static_cast<const __nv_bfloat16*>(w_bf16); g.w_rows = N; g.w_cols = K;
  g.M = M; g.N = N; g.K = K; g.k_inner = k_inner;
  g.bias = bias; g.out_bf16 = static_cast<__nv_bfloat16*>(out_bf16); g.ld_out = N; g.act = kActNone;
  g.rowln_gamma = gamma; g.rowln_beta = beta; g.rowln_eps = eps; g.rowln_gelu = gelu;
  if (!gemm_rowln_supported(g)) return fail(kUnsupported, "fused row LayerNorm: needs N = 512 and the CTA-pair GEMM (K % 64 == 0)");
  return gemm_bf16_tc(g, static_cast<cudaStream_t>(stream));
}

int svt_op_gemm_ln(const void* a_bf16, const void* w_bf16, const float* bias, const float* colsum, const float* ln_stats,
                   float ln_eps, float* row_stats_out, const float* resid, float* out_f32, void* out_bf16, int M, int N,
                   int K, int act, void* stream) {
  if (a_bf16 == nullptr || w_bf16 == nullptr) return fail(kInvalidArgument, "null argument");
  GemmArgs g;
  g.a = static_cast<const __nv_bfloat16*>(a_bf16);
  g.a_dims[0] = K; g.a_dims[1] = 1; g.a_dims[2] = M;
  g.a_strides[0] = K; g.a_strides[1] = static_cast<uint64_t>(K);
  g.w = static_cast<const __nv_bfloat16*>(w_bf16); g.w_rows = N; g.w_cols = K;
  g.M = M; g.N = N; g.K = K; g.k_inner = K;
  g.bias = bias; g.resid = resid; g.out_f32 = out_f32; g.out_bf16 = static_cast<__nv_bfloat16*>(out_bf16);
  g.ld_out = N; g.act = act;
  g.row_stats_out = row_stats_out; g.ln_stats = ln_stats; g.ln_colsum = colsum; g.ln_eps = ln_eps;
  return gemm_bf16_tc(g, static_cast<cudaStream_t>(stream));
}

int svt_wavlm_relative_bucket(int relative_position, int num_buckets, int max_distance) {
  return wavlm_relative_bucket(relative_position, num_buckets, max_distance);
}

int svt_op_row_stats_cast(const float* x, int rows, int D, void* y_bf16, float* stats, void* stream) {
  if (x == nullptr || y_bf16 == nullptr || stats == nullptr) return fail(kInvalidArgument, "null argument");
  return row_stats_cast(x, rows, D, static_cast<__nv_bfloat16*>(y_bf16), stats, static_cast<cudaStream_t>(stream));
}

int svt_op_pack_posconv(const float* w_f32_dev, int D, int groups, int taps, void* out_bf16, void* stream) {
  if (w_f32_dev == nullptr || out_bf16 == nullptr) return fail(kInvalidArgument, "null argument");
  return pack_posconv_weight(w_f32_dev, nullptr, D, groups, taps, static_cast<__nv_bfloat16*>(out_bf16),
                             static_cast<cudaStream_t>(stream));
}

int svt_op_posconv(const void* x_bf16, const void* w_packed, const float* bias, const float* resid, float* out_f32,
                   int clips, int clip_rows, int t_valid, int D, int groups, int taps, void* stream) {
  if (x_bf16 == nullptr || w_packed == nullptr || out_f32 == nullptr) return fail(kInvalidArgument, "null argument");
  GemmArgs g;
  g.mode = 1;
  g.a = static_cast<const __nv_bfloat16*>(x_bf16);
  g.a_dims[0] = D; g.a_dims[1] = t_valid; g.a_dims[2] = clips;
  g.a_strides[0] = D; g.a_strides[1] = static_cast<uint64_t>(clip_rows) * D;
  g.w = static_cast<const __nv_bfloat16*>(w_packed); g.w_rows = groups * taps * 64; g.w_cols = 64;
  g.N = D; g.K = taps * 64;
  g.n_clips = clips; g.clip_rows = clip_rows; g.clip_valid = t_valid; g.pad_left = taps / 2; g.taps = taps;
  g.group_size = D / groups;
  g.bias = bias; g.resid = resid; g.out_f32 = out_f32; g.ld_out = D; g.act = kActGelu;
  return gemm_bf16_tc(g, static_cast<cudaStream_t>(stream));
}

int svt_op_attention(const void* q, const void* k, const void* v, void* o, int ldq, int ldk, int ldv, int ldo, int Tq,
                     int Tk, int q_clip_rows, int k_clip_rows, int clips, int heads, int head_dim, void* stream) {
  if (q == nullptr || k == nullptr || v == nullptr || o == nullptr) return fail(kInvalidArgument, "null argument");
  AttentionArgs a;
  a.q = static_cast<const __nv_bfloat16*>(q); a.k = static_cast<const __nv_bfloat16*>(k);
  a.v = static_cast<const __nv_bfloat16*>(v); a.o = static_cast<__nv_bfloat16*>(o);
  a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo; a.Tq = Tq; a.Tk = Tk;
  a.q_clip_rows = q_clip_rows; a.k_clip_rows = k_clip_rows; a.clips = clips; a.heads = heads; a.head_dim = head_dim;
  return attention_bf16(a, static_cast<cudaStream_t>(stream));
}

int svt_op_layer_norm(const float* x_f32, const void* x_bf16, const float* gamma, const float* beta, void* y_bf16,
                      float* y_f32, int rows, int D, float eps, int gelu, void* stream) {
  LayerNormArgs a;
  a.x_f32 = x_f32; a.x_bf16 = static_cast<const __nv_bfloat16*>(x_bf16); a.gamma = gamma; a.beta = beta;
  a.y_bf16 = static_cast<__nv_bfloat16*>(y_bf16); a.y_f32 = y_f32; a.rows = rows; a.D = D; a.eps = eps; a.gelu = gelu;
  return layer_norm(a, static_cast<cudaStream_t>(stream));
}

int svt_op_linear_small(const float* x, int rows, int D, const float* w, const float* b, int n_out, float* y,
                        void* stream) {
  if (x == nullptr || w == nullptr || y == nullptr) return fail(kInvalidArgument, "null argument");
  HeadArgs h;
  h.x = x; h.clips = 1; h.clip_rows = rows; h.T = rows; h.D = D; h.w = w; h.b = b; h.n_out = n_out; h.logits = y;
  return head_forward(h, static_cast<cudaStream_t>(stream));
}

int svt_op_conv0(const float* wav, int B, int L, const float* w_kc, const float* bias, const float* gamma,
                 const float* beta, int normalize, void* out_bf16, int t_alloc, double* stats_scratch, void* stream) {
  if (wav == nullptr || w_kc == nullptr || out_bf16 == nullptr) return fail(kInvalidArgument, "null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (normalize) {
    if (stats_scratch == nullptr) return fail(kInvalidArgument, "stats scratch required");
    SVT_TRY(tensor_stats(wav, static_cast<size_t>(B) * L, stats_scratch, s));
  }
  Conv0Args a;
  a.wav = wav; a.B = B; a.L = L; a.T = (L - 10) / 5 + 1; a.t_alloc = t_alloc;
  a.w = w_kc; a.bias = bias; a.gamma = gamma; a.beta = beta; a.in_stats = normalize ? stats_scratch : nullptr;
  a.out = static_cast<__nv_bfloat16*>(out_bf16); a.layer_mode = 1;
  void* tables = nullptr;
  if (get_option_conv0_impl() != 1) {  // test hook: tables built per call (the encoder builds them once at finalize)
    SVT_CUDA(cudaMalloc(&tables, conv0_tables_bytes()));
    const int rc = conv0_build_tables(w_kc, bias, tables, s);
    if (rc != kOk) { cudaFree(tables); return rc; }
    a.tc_tables = tables;
  }
  const int rc = conv0_forward(a, s);
  if (tables != nullptr) {
    cudaStreamSynchronize(s);
    cudaFree(tables);
  }
  return rc;
}

}  // extern "C"
