// Residual cross-attention fusion (FusionRCA, N20EMv2/audio_visual/fusion.py:186-210) on the device.
//   align video to audio frames (:196-203)  ->  + sinusoidal PE (:59-60)
//   layer1(kv = audio, q = video), layer2(kv = video, q = audio) (:63-78), each (:137-183, post-LN):
//     S = MHA(kv, kv, kv), X = MHA(q, kv, kv) with SHARED weights; y = LN1(kv + a*S + (1-a)*X);
//     out = LN2(y + W2 relu(W1 y))
//   result = out1 + out2 (:209)
// K/V are projected once per layer (the reference projects them twice), and because out_proj is linear the
// two attention contexts go through ONE GEMM with K = 2D against [a*Wo | (1-a)*Wo].
#include <cmath>

#include "model.cuh"

using namespace svt;

namespace {

struct FusPlan {
  int M;
  size_t off_af, off_vf, off_ab, off_vb, off_qkv, off_qc, off_ctx2, off_y, off_yb, off_mid, off_o1, off_o2, total;
};
FusPlan make_plan(const svt_fusion* f, int B, int T) {
  FusPlan p{};
  p.M = B * T;
  const size_t M = p.M, D = f->cfg.d_model, F = f->cfg.d_ffn;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
  p.off_af = take(M * D * 4); p.off_vf = take(M * D * 4);
  p.off_ab = take(M * D * 2); p.off_vb = take(M * D * 2);
  p.off_qkv = take(M * 3 * D * 2); p.off_qc = take(M * D * 2);
  p.off_ctx2 = take(M * 2 * D * 2);
  p.off_y = take(M * D * 4); p.off_yb = take(M * D * 2);
  p.off_mid = take(M * F * 2);
  p.off_o1 = take(M * D * 4); p.off_o2 = take(M * D * 4);
  p.total = off;
  return p;
}

int lin(const __nv_bfloat16* a, int M, const __nv_bfloat16* w, const float* b, int N, int K, const float* resid,
        float* of, __nv_bfloat16* ob, int act, cudaStream_t s) {
  GemmArgs g;
  g.a = a;
  g.a_dims[0] = K; g.a_dims[1] = 1; g.a_dims[2] = M;
  g.a_strides[0] = K; g.a_strides[1] = K;
  g.w = w; g.w_rows = N; g.w_cols = K;
  g.M = M; g.N = N; g.K = K; g.k_inner = K;
  g.bias = b; g.resid = resid; g.out_f32 = of; g.out_bf16 = ob; g.ld_out = N; g.act = act;
  return gemm_bf16_tc(g, s);
}

int pack_mat(DevicePool& pool, const RawTensor* t, int N, int K, float scale, __nv_bfloat16** out) {
  SVT_TRY(pool.alloc_t<__nv_bfloat16>(static_cast<size_t>(N) * K, out));
  PackArgs a;
  a.src = t->dev; a.dims[2] = N; a.dims[3] = K; a.strides[2] = K; a.strides[3] = 1; a.scale = scale;
  return pack_bf16(a, *out, 0);
}
int pack_vecf(DevicePool& pool, const RawTensor* t, int n, float** out) {
  SVT_TRY(pool.alloc_t<float>(n, out));
  SVT_CUDA(cudaMemcpy(*out, t->dev, sizeof(float) * n, cudaMemcpyDeviceToDevice));
  return kOk;
}

int finalize_fusion(svt_fusion* f) {
  const int D = f->cfg.d_model, F = f->cfg.d_ffn, H = f->cfg.nhead, dh = D / H;
  f->pool.release();
  const float qs = 1.0f / std::sqrt(static_cast<float>(dh));
  for (int l = 0; l < 2; ++l) {
    const std::string p = "fusion.layer" + std::to_string(l + 1) + ".";
    svt_fusion::Layer& L = f->layer[l];
    const RawTensor *w, *b, *t;
    SVT_TRY(f->reg.require(p + "self_att.att.in_proj_weight", {3 * D, D}, &w));
    SVT_TRY(f->reg.require(p + "self_att.att.in_proj_bias", {3 * D}, &b));
    SVT_TRY(f->pool.alloc_t<__nv_bfloat16>(static_cast<size_t>(3) * D * D, &L.qkv.w));
    SVT_TRY(f->pool.alloc_t<float>(3 * D, &L.qkv.b));
    for (int j = 0; j < 3; ++j) {  // rows [0:D] = Wq (scaled), [D:2D] = Wk, [2D:3D] = Wv
      PackArgs a;
      a.src = w->dev + static_cast<size_t>(j) * D * D; a.dims[2] = D; a.dims[3] = D; a.strides[2] = D; a.strides[3] = 1;
      a.scale = j == 0 ? qs : 1.f;
      SVT_TRY(pack_bf16(a, L.qkv.w + static_cast<size_t>(j) * D * D, 0));
      PackArgs bb;
      bb.src = b->dev + static_cast<size_t>(j) * D; bb.dims[3] = D; bb.strides[3] = 1; bb.scale = a.scale;
      SVT_TRY(pack_f32(bb, L.qkv.b + static_cast<size_t>(j) * D, 0));
    }
    L.qkv.N = 3 * D; L.qkv.K = D;
    // out2[n][0:D] = alpha * Wo[n][:], out2[n][D:2D] = (1 - alpha) * Wo[n][:]
    SVT_TRY(f->reg.require(p + "self_att.att.out_proj.weight", {D, D}, &w));
    SVT_TRY(f->pool.alloc_t<__nv_bfloat16>(static_cast<size_t>(2) * D * D, &L.out2.w));
    {
      __nv_bfloat16* tmp;
      SVT_TRY(f->pool.alloc_t<__nv_bfloat16>(static_cast<size_t>(D) * D, &tmp));
      for (int half = 0; half < 2; ++half) {
        PackArgs a;
        a.src = w->dev; a.dims[2] = D; a.dims[3] = D; a.strides[2] = D; a.strides[3] = 1;
        a.scale = half == 0 ? f->cfg.alpha : 1.f - f->cfg.alpha;
        SVT_TRY(pack_bf16(a, tmp, 0));
        SVT_CUDA(cudaMemcpy2D(L.out2.w + static_cast<size_t>(half) * D, sizeof(__nv_bfloat16) * 2 * D, tmp,
                              sizeof(__nv_bfloat16) * D, sizeof(__nv_bfloat16) * D, D, cudaMemcpyDeviceToDevice));
      }
    }
    SVT_TRY(f->reg.require(p + "self_att.att.out_proj.bias", {D}, &t));
    SVT_TRY(pack_vecf(f->pool, t, D, &L.out2.b));
    L.out2.N = D; L.out2.K = 2 * D;
    SVT_TRY(f->reg.require(p + "pos_ffn.ffn.0.weight", {F, D}, &w));
    SVT_TRY(pack_mat(f->pool, w, F, D, 1.f, &L.ff1.w));
    SVT_TRY(f->reg.require(p + "pos_ffn.ffn.0.bias", {F}, &t));
    SVT_TRY(pack_vecf(f->pool, t, F, &L.ff1.b));
    L.ff1.N = F; L.ff1.K = D;
    SVT_TRY(f->reg.require(p + "pos_ffn.ffn.3.weight", {D, F}, &w));
    SVT_TRY(pack_mat(f->pool, w, D, F, 1.f, &L.ff2.w));
    SVT_TRY(f->reg.require(p + "pos_ffn.ffn.3.bias", {D}, &t));
    SVT_TRY(pack_vecf(f->pool, t, D, &L.ff2.b));
    L.ff2.N = D; L.ff2.K = F;
    SVT_TRY(f->reg.require(p + "norm1.norm.weight", {D}, &t)); SVT_TRY(pack_vecf(f->pool, t, D, &L.n1.g));
    SVT_TRY(f->reg.require(p + "norm1.norm.bias", {D}, &t));   SVT_TRY(pack_vecf(f->pool, t, D, &L.n1.b));
    SVT_TRY(f->reg.require(p + "norm2.norm.weight", {D}, &t)); SVT_TRY(pack_vecf(f->pool, t, D, &L.n2.g));
    SVT_TRY(f->reg.require(p + "norm2.norm.bias", {D}, &t));   SVT_TRY(pack_vecf(f->pool, t, D, &L.n2.b));
  }
  SVT_CUDA(cudaDeviceSynchronize());
  f->reg.clear();
  f->finalized = true;
  return kOk;
}

}  // namespace

extern "C" {

int svt_fusion_create(const svt_fusion_config* cfg, svt_fusion** out) {
  if (cfg == nullptr || out == nullptr) return fail(kInvalidArgument, "null argument");
  if (cfg->d_model % 128 != 0 || cfg->d_model > 2048) return fail(kUnsupported, "d_model must be a multiple of 128, <= 2048");
  if (cfg->nhead <= 0 || cfg->d_model % cfg->nhead != 0) return fail(kInvalidArgument, "bad nhead");
  const int dh = cfg->d_model / cfg->nhead;
  if (dh != 64 && dh != 128) return fail(kUnsupported, "head dim must be 64 or 128");
  if (cfg->d_ffn % 64 != 0) return fail(kUnsupported, "d_ffn must be a multiple of 64");
  svt_fusion* f = new svt_fusion();
  f->cfg = *cfg;
  *out = f;
  return kOk;
}
void svt_fusion_destroy(svt_fusion* f) { delete f; }

int svt_fusion_set_tensor(svt_fusion* f, const char* name, const float* host, const int64_t* shape, int ndim, int strict) {
  if (f == nullptr || name == nullptr || host == nullptr) return fail(kInvalidArgument, "null argument");
  if (svt_device_count() <= 0) return fail(kNoDevice, "no CUDA device");
  std::string n(name);
  if (n.find("positional_encoding.pe") != std::string::npos) return kOk;  // recomputed on the device
  if (n.rfind("fusion.layer", 0) != 0) return strict ? fail(kUnknownTensor, "unknown tensor " + n) : static_cast<int>(kOk);
  f->finalized = false;
  return f->reg.set(n, host, shape, ndim);
}
int svt_fusion_finalize(svt_fusion* f) {
  if (f == nullptr) return fail(kInvalidArgument, "null argument");
  if (svt_device_count() <= 0) return fail(kNoDevice, "no CUDA device");
  return finalize_fusion(f);
}
size_t svt_fusion_workspace_bytes(const svt_fusion* f, int batch, int t_audio) {
  if (f == nullptr || batch <= 0 || t_audio <= 0) return 0;
  return make_plan(f, batch, t_audio).total;
}

int svt_fusion_forward(svt_fusion* f, const float* audio, const float* video, int B, int Ta, int Tv, void* ws,
                       size_t ws_bytes, float* out, void* stream) {
  if (f == nullptr || audio == nullptr || video == nullptr || ws == nullptr || out == nullptr)
    return fail(kInvalidArgument, "null argument");
  if (!f->finalized) return fail(kNotFinalized, "svt_fusion_finalize has not been called");
  if (B <= 0 || Ta <= 0 || Tv <= 0) return fail(kInvalidArgument, "empty input");
  if (Ta > 2500) return fail(kInvalidArgument, "more frames than PositionalEncoding max_len (2500)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const FusPlan p = make_plan(f, B, Ta);
  if (ws_bytes < p.total) return fail(kWorkspaceTooSmall, "workspace too small: need " + std::to_string(p.total));
  uint8_t* base = static_cast<uint8_t*>(ws);
  float* af = reinterpret_cast<float*>(base + p.off_af);
  float* vf = reinterpret_cast<float*>(base + p.off_vf);
  __nv_bfloat16* ab = reinterpret_cast<__nv_bfloat16*>(base + p.off_ab);
  __nv_bfloat16* vb = reinterpret_cast<__nv_bfloat16*>(base + p.off_vb);
  __nv_bfloat16* qkv = reinterpret_cast<__nv_bfloat16*>(base + p.off_qkv);
  __nv_bfloat16* qc = reinterpret_cast<__nv_bfloat16*>(base + p.off_qc);
  __nv_bfloat16* ctx2 = reinterpret_cast<__nv_bfloat16*>(base + p.off_ctx2);
  float* y = reinterpret_cast<float*>(base + p.off_y);
  __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(base + p.off_yb);
  __nv_bfloat16* mid = reinterpret_cast<__nv_bfloat16*>(base + p.off_mid);
  float* o1 = reinterpret_cast<float*>(base + p.off_o1);
  float* o2 = reinterpret_cast<float*>(base + p.off_o2);
  const int D = f->cfg.d_model, H = f->cfg.nhead, dh = D / H, M = p.M;

  // frame alignment + positional encoding (video truncated / zero padded to Ta frames)
  SVT_TRY(add_positional_encoding(audio, B, Ta, Ta, D, af, ab, s));
  SVT_TRY(add_positional_encoding(video, B, Tv, Ta, D, vf, vb, s));

  for (int l = 0; l < 2; ++l) {
    const svt_fusion::Layer& L = f->layer[l];
    const float* kvf = l == 0 ? af : vf;
    const __nv_bfloat16* kvb = l == 0 ? ab : vb;
    const __nv_bfloat16* qb = l == 0 ? vb : ab;
    float* o = l == 0 ? o1 : o2;
    // [Qs | K | V] from the kv stream, Qc from the query stream (same Wq)
    SVT_TRY(lin(kvb, M, L.qkv.w, L.qkv.b, 3 * D, D, nullptr, nullptr, qkv, kActNone, s));
    SVT_TRY(lin(qb, M, L.qkv.w, L.qkv.b, D, D, nullptr, nullptr, qc, kActNone, s));
    AttentionArgs a;
    a.k = qkv + D; a.v = qkv + 2 * D; a.ldk = a.ldv = 3 * D;
    a.Tq = Ta; a.Tk = Ta; a.q_clip_rows = Ta; a.k_clip_rows = Ta; a.clips = B; a.heads = H; a.head_dim = dh;
    a.ldo = 2 * D;
    a.q = qkv; a.ldq = 3 * D; a.o = ctx2;          // self attention -> columns [0, D)
    SVT_TRY(attention_bf16(a, s));
    a.q = qc; a.ldq = D; a.o = ctx2 + D;           // cross attention -> columns [D, 2D)
    SVT_TRY(attention_bf16(a, s));
    // y = LN1(kv + alpha*O(S') + (1-alpha)*O(X'))
    SVT_TRY(lin(ctx2, M, L.out2.w, L.out2.b, D, 2 * D, kvf, y, nullptr, kActNone, s));
    LayerNormArgs ln;
    ln.x_f32 = y; ln.gamma = L.n1.g; ln.beta = L.n1.b; ln.y_f32 = y; ln.y_bf16 = yb; ln.rows = M; ln.D = D; ln.eps = 1e-6f;
    SVT_TRY(layer_norm(ln, s));
    // out = LN2(y + W2 relu(W1 y))
    SVT_TRY(lin(yb, M, L.ff1.w, L.ff1.b, f->cfg.d_ffn, D, nullptr, nullptr, mid, kActRelu, s));
    SVT_TRY(lin(mid, M, L.ff2.w, L.ff2.b, D, f->cfg.d_ffn, y, o, nullptr, kActNone, s));
    LayerNormArgs l2;
    l2.x_f32 = o; l2.gamma = L.n2.g; l2.beta = L.n2.b; l2.y_f32 = o; l2.rows = M; l2.D = D; l2.eps = 1e-6f;
    SVT_TRY(layer_norm(l2, s));
  }
  return add_f32(o1, o2, out, static_cast<size_t>(M) * D, s);
}

}  // extern "C"
