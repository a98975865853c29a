// wav2vec2-style SSL encoder + AMT head: weight packing and forward orchestration.
// Reference semantics: MIR_ST500/huggingface_interface.py:263-298 wrapping HF Wav2Vec2Model
// (modeling_wav2vec2.py:382-434 feature encoder/projection, :326-379 positional conv,
//  :576-655 encoder layers, :658-803 encoders) and speechbrain/nnet/linear.py:61-76 (head).
#include <cmath>
#include <cstring>

#include "model.cuh"

namespace svt {

// ------------------------------------------------------------------------------------ pool / registry
int DevicePool::alloc(size_t bytes, void** out) {
  void* p = nullptr;
  SVT_CUDA(cudaMalloc(&p, bytes > 0 ? bytes : 16));
  ptrs_.push_back(p);
  *out = p;
  return kOk;
}
void DevicePool::release() {
  for (void* p : ptrs_) cudaFree(p);
  ptrs_.clear();
}

int WeightRegistry::set(const std::string& name, const float* host, const int64_t* shape, int ndim) {
  RawTensor t;
  t.shape.assign(shape, shape + ndim);
  const size_t n = t.numel();
  auto it = t_.find(name);
  if (it != t_.end()) {
    cudaFree(it->second.dev);
    t_.erase(it);
  }
  SVT_CUDA(cudaMalloc(&t.dev, sizeof(float) * (n > 0 ? n : 1)));
  SVT_CUDA(cudaMemcpy(t.dev, host, sizeof(float) * n, cudaMemcpyHostToDevice));
  t_[name] = t;
  return kOk;
}
const RawTensor* WeightRegistry::find(const std::string& name) const {
  auto it = t_.find(name);
  return it == t_.end() ? nullptr : &it->second;
}
int WeightRegistry::require(const std::string& name, std::initializer_list<int64_t> shape, const RawTensor** out) const {
  const RawTensor* t = find(name);
  if (t == nullptr) return fail(kUnknownTensor, "missing tensor: " + name);
  if (shape.size() > 0) {
    std::vector<int64_t> want(shape);
    if (want != t->shape) {
      std::string s = "shape mismatch for " + name + ": got (";
      for (auto d : t->shape) s += std::to_string(d) + ",";
      s += ") expected (";
      for (auto d : want) s += std::to_string(d) + ",";
      return fail(kInvalidArgument, s + ")");
    }
  }
  *out = t;
  return kOk;
}
void WeightRegistry::clear() {
  for (auto& kv : t_) cudaFree(kv.second.dev);
  t_.clear();
}

namespace {

int pack_vec(DevicePool& pool, const WeightRegistry& reg, const std::string& name, int64_t n, float scale, float** out) {
  const RawTensor* t;
  SVT_TRY(reg.require(name, {n}, &t));
  SVT_TRY(pool.alloc_t<float>(static_cast<size_t>(n), out));
  PackArgs a;
  a.src = t->dev;
  a.dims[3] = static_cast<int>(n);
  a.strides[3] = 1;
  a.scale = scale;
  return pack_f32(a, *out, 0);
}
int zero_vec(DevicePool& pool, int64_t n, float value, float** out) {
  SVT_TRY(pool.alloc_t<float>(static_cast<size_t>(n), out));
  std::vector<float> h(static_cast<size_t>(n), value);
  SVT_CUDA(cudaMemcpy(*out, h.data(), sizeof(float) * n, cudaMemcpyHostToDevice));
  return kOk;
}
int pack_norm(DevicePool& pool, const WeightRegistry& reg, const std::string& prefix, int64_t n, NormW* out) {
  SVT_TRY(pack_vec(pool, reg, prefix + "weight", n, 1.f, &out->g));
  return pack_vec(pool, reg, prefix + "bias", n, 1.f, &out->b);
}
// dst rows [row0, row0+N) of a [*, K] bf16 matrix <- scale * src (N, K)
int pack_rows(const RawTensor* t, int N, int K, float scale, __nv_bfloat16* dst, int row0, const float* col_scale = nullptr) {
  PackArgs a;
  a.src = t->dev;
  a.dims[2] = N; a.dims[3] = K;
  a.strides[2] = K; a.strides[3] = 1;
  a.scale = scale;
  a.vec = col_scale; a.vec_dim = 3;
  return pack_bf16(a, dst + static_cast<size_t>(row0) * K, 0);
}
// LayerNorm(gamma, beta) folded into the Linear that consumes it: rows [row0, row0 + N) of dst.w <- scale * W o gamma,
// dst.colsum <- row sums of the packed bf16 weight, dst.b (already scale * b) += scale * W.beta
int fold_norm_rows(const RawTensor* w, int N, int K, float scale, const NormW& norm, LinearW* dst, int row0) {
  SVT_TRY(pack_rows(w, N, K, scale, dst->w, row0, norm.g));
  return ln_fold_vectors(w->dev, dst->w + static_cast<size_t>(row0) * K, norm.b, scale, N, K, dst->colsum + row0,
                         dst->b + row0, 0);
}
int pack_linear(DevicePool& pool, const WeightRegistry& reg, const std::string& prefix, int N, int K, LinearW* out) {
  const RawTensor* w;
  SVT_TRY(reg.require(prefix + "weight", {N, K}, &w));
  SVT_TRY(pool.alloc_t<__nv_bfloat16>(static_cast<size_t>(N) * K, &out->w));
  SVT_TRY(pack_rows(w, N, K, 1.f, out->w, 0));
  SVT_TRY(pack_vec(pool, reg, prefix + "bias", N, 1.f, &out->b));
  out->N = N; out->K = K;
  return kOk;
}

}  // namespace

int pack_posconv_weight(const float* w_dev, const float* tap_scale, int D, int groups, int taps, __nv_bfloat16* dst,
                        cudaStream_t stream) {
  // dst[g][j][co][ci] (64 x 64 per tap, zero padded) = w[(g*Dg+co)][ci][j] * tap_scale[j]
  const int Dg = D / groups;
  SVT_CUDA(cudaMemsetAsync(dst, 0, sizeof(__nv_bfloat16) * static_cast<size_t>(groups) * taps * 64 * 64, stream));
  if (Dg == 64) {
    PackArgs a;
    a.src = w_dev;
    a.dims[0] = groups; a.dims[1] = taps; a.dims[2] = 64; a.dims[3] = 64;
    a.strides[0] = static_cast<long long>(Dg) * Dg * taps; a.strides[1] = 1;
    a.strides[2] = static_cast<long long>(Dg) * taps; a.strides[3] = taps;
    a.vec = tap_scale; a.vec_dim = 1;
    return pack_bf16(a, dst, stream);
  }
  return fail(kUnsupported, "positional conv with channels-per-group != 64 is packed on the host side");
}

}  // namespace svt

using namespace svt;

// ------------------------------------------------------------------------------------ geometry
int svt_encoder::conv_out_len(int L, int upto) const {
  int t = L;
  for (int i = 0; i <= upto; ++i) {
    if (t < cfg.conv_kernel[i]) return 0;
    t = (t - cfg.conv_kernel[i]) / cfg.conv_stride[i] + 1;
  }
  return t;
}
int svt_encoder::t_alloc0(int L) const {
  int prod = 1;
  for (int i = 1; i < cfg.num_conv_layers; ++i) prod *= cfg.conv_stride[i];
  if (prod % 4 != 0) prod *= 4;
  const int t0 = conv_out_len(L, 0);
  return (t0 + prod - 1) / prod * prod;
}

// ------------------------------------------------------------------------------------ packing
namespace svt {
int encoder_finalize(svt_encoder* e);
}
int svt::encoder_finalize(svt_encoder* e) {
  const svt_encoder_config& c = e->cfg;
  const int D = c.hidden_size, F = c.ffn_size, C = c.conv_dim;
  e->pool.release();
  e->conv.clear(); e->conv_norm.clear(); e->layers.clear();
  DevicePool& pool = e->pool;
  const WeightRegistry& reg = e->reg;
  const RawTensor* t;

  if (!e->transformer_only) {
  // conv layer 0: (C, 1, k) -> [k][C] fp32
  const std::string fe = "feature_extractor.conv_layers.";
  SVT_TRY(reg.require(fe + "0.conv.weight", {C, 1, c.conv_kernel[0]}, &t));
  SVT_TRY(pool.alloc_t<float>(static_cast<size_t>(C) * c.conv_kernel[0], &e->conv0_w));
  {
    PackArgs a;
    a.src = t->dev;
    a.dims[2] = c.conv_kernel[0]; a.dims[3] = C;
    a.strides[2] = 1; a.strides[3] = c.conv_kernel[0];
    SVT_TRY(pack_f32(a, e->conv0_w, 0));
  }
  if (c.conv_bias) SVT_TRY(pack_vec(pool, reg, fe + "0.conv.bias", C, 1.f, &e->conv0_b));
  else SVT_TRY(zero_vec(pool, C, 0.f, &e->conv0_b));
  SVT_TRY(pack_norm(pool, reg, fe + "0.layer_norm.", C, &e->conv0_norm));  // LN (large) or GroupNorm (base)
  e->conv0_tab = nullptr;
  if (c.feat_norm_layer && C == 512 && c.conv_kernel[0] == 10 && c.conv_stride[0] == 5) {
    SVT_TRY(pool.alloc(conv0_tables_bytes(), &e->conv0_tab));
    SVT_TRY(conv0_build_tables(e->conv0_w, e->conv0_b, e->conv0_tab, 0));
  }

  // conv layers 1..: (C, C, k) -> [C][k][C_in] bf16 so that K index = tap * C_in + ci
  for (int i = 1; i < c.num_conv_layers; ++i) {
    const int k = c.conv_kernel[i];
    const std::string p = fe + std::to_string(i) + ".";
    SVT_TRY(reg.require(p + "conv.weight", {C, C, k}, &t));
    LinearW lw;
    lw.N = C; lw.K = k * C;
    SVT_TRY(pool.alloc_t<__nv_bfloat16>(static_cast<size_t>(C) * k * C, &lw.w));
    PackArgs a;
    a.src = t->dev;
    a.dims[1] = C; a.dims[2] = k; a.dims[3] = C;
    a.strides[1] = static_cast<long long>(C) * k; a.strides[2] = 1; a.strides[3] = k;
    SVT_TRY(pack_bf16(a, lw.w, 0));
    if (c.conv_bias) SVT_TRY(pack_vec(pool, reg, p + "conv.bias", C, 1.f, &lw.b));
    else SVT_TRY(zero_vec(pool, C, 0.f, &lw.b));
    e->conv.push_back(lw);
    NormW nw;
    if (c.feat_norm_layer) SVT_TRY(pack_norm(pool, reg, p + "layer_norm.", C, &nw));
    e->conv_norm.push_back(nw);
  }
  if (c.feat_proj_norm) SVT_TRY(pack_norm(pool, reg, "feature_projection.layer_norm.", C, &e->proj_norm));
  SVT_TRY(pack_linear(pool, reg, "feature_projection.projection.", D, C, &e->proj));
  }  // !transformer_only

  // positional conv weights packed per (group, tap): dst[g][j][co][ci] = v[g*Dg+co][ci][j] * scale[j]
  auto pack_pos = [&](const RawTensor* v, const float* scale, __nv_bfloat16** dst) -> int {
    const int G = c.pos_conv_groups, taps = c.pos_conv_kernel, Dg = D / G;
    SVT_TRY(pool.alloc_t<__nv_bfloat16>(static_cast<size_t>(G) * taps * 64 * 64, dst));
    SVT_CUDA(cudaMemset(*dst, 0, sizeof(__nv_bfloat16) * static_cast<size_t>(G) * taps * 64 * 64));
    if (Dg == 64) return pack_posconv_weight(v->dev, scale, D, G, taps, *dst, 0);
    // generic (base: Dg = 48): build on the host, small one-time cost
    std::vector<float> hv(v->numel()), hs(taps, 1.f);
    SVT_CUDA(cudaMemcpy(hv.data(), v->dev, sizeof(float) * hv.size(), cudaMemcpyDeviceToHost));
    if (scale != nullptr) SVT_CUDA(cudaMemcpy(hs.data(), scale, sizeof(float) * taps, cudaMemcpyDeviceToHost));
    std::vector<__nv_bfloat16> hp(static_cast<size_t>(G) * taps * 64 * 64, __float2bfloat16(0.f));
    for (int gi = 0; gi < G; ++gi)
      for (int j = 0; j < taps; ++j)
        for (int co = 0; co < Dg; ++co)
          for (int ci = 0; ci < Dg; ++ci)
            hp[((static_cast<size_t>(gi) * taps + j) * 64 + co) * 64 + ci] =
                __float2bfloat16(hv[(static_cast<size_t>(gi * Dg + co) * Dg + ci) * taps + j] * hs[j]);
    SVT_CUDA(cudaMemcpy(*dst, hp.data(), sizeof(__nv_bfloat16) * hp.size(), cudaMemcpyHostToDevice));
    return kOk;
  };
  e->pos_stack.clear();
  if (c.pos_conv_layers > 0) {
    // data2vec-audio: plain conv weights, one set per stacked layer (HF Data2VecAudioPositionalConvLayer)
    const int taps = c.pos_conv_kernel, Dg = D / c.pos_conv_groups;
    for (int i = 0; i < c.pos_conv_layers; ++i) {
      const std::string pc = "encoder.pos_conv_embed.layers." + std::to_string(i) + ".conv.";
      const RawTensor* w;
      SVT_TRY(reg.require(pc + "weight", {D, Dg, taps}, &w));
      svt_encoder::PosLayer pl;
      SVT_TRY(pack_pos(w, nullptr, &pl.w));
      SVT_TRY(pack_vec(pool, reg, pc + "bias", D, 1.f, &pl.b));
      e->pos_stack.push_back(pl);
    }
    SVT_TRY(zero_vec(pool, D, 1.f, &e->ones));
    SVT_TRY(zero_vec(pool, D, 0.f, &e->zeros));
  } else if (c.pos_conv_batch_norm) {
    // HuBERT conv_pos_batch_norm: BatchNorm1d (running statistics) in front of a plain grouped conv
    const int taps = c.pos_conv_kernel, Dg = D / c.pos_conv_groups;
    const std::string bn = "encoder.pos_conv_embed.batch_norm.";
    const RawTensor *w, *g, *b, *rm, *rv;
    SVT_TRY(reg.require("encoder.pos_conv_embed.conv.weight", {D, Dg, taps}, &w));
    SVT_TRY(reg.require(bn + "weight", {D}, &g));
    SVT_TRY(reg.require(bn + "bias", {D}, &b));
    SVT_TRY(reg.require(bn + "running_mean", {D}, &rm));
    SVT_TRY(reg.require(bn + "running_var", {D}, &rv));
    std::vector<float> hg(D), hb(D), hm(D), hv(D), sc(D), sh(D);
    SVT_CUDA(cudaMemcpy(hg.data(), g->dev, sizeof(float) * D, cudaMemcpyDeviceToHost));
    SVT_CUDA(cudaMemcpy(hb.data(), b->dev, sizeof(float) * D, cudaMemcpyDeviceToHost));
    SVT_CUDA(cudaMemcpy(hm.data(), rm->dev, sizeof(float) * D, cudaMemcpyDeviceToHost));
    SVT_CUDA(cudaMemcpy(hv.data(), rv->dev, sizeof(float) * D, cudaMemcpyDeviceToHost));
    for (int i = 0; i < D; ++i) {
      sc[i] = hg[i] / std::sqrt(hv[i] + 1e-5f);  // nn.BatchNorm1d default eps
      sh[i] = hb[i] - hm[i] * sc[i];
    }
    SVT_TRY(pool.alloc_t<float>(D, &e->pos_bn_scale));
    SVT_TRY(pool.alloc_t<float>(D, &e->pos_bn_shift));
    SVT_CUDA(cudaMemcpy(e->pos_bn_scale, sc.data(), sizeof(float) * D, cudaMemcpyHostToDevice));
    SVT_CUDA(cudaMemcpy(e->pos_bn_shift, sh.data(), sizeof(float) * D, cudaMemcpyHostToDevice));
    SVT_TRY(pack_pos(w, nullptr, &e->pos_w));
    SVT_TRY(pack_vec(pool, reg, "encoder.pos_conv_embed.conv.bias", D, 1.f, &e->pos_b));
  } else {
    // weight-norm recomposition folded here (HF:343-355)
    const int taps = c.pos_conv_kernel, Dg = D / c.pos_conv_groups;
    const std::string pc = "encoder.pos_conv_embed.conv.";
    const RawTensor* g = reg.find(pc + "parametrizations.weight.original0");
    const RawTensor* v = reg.find(pc + "parametrizations.weight.original1");
    if (g == nullptr) g = reg.find(pc + "weight_g");
    if (v == nullptr) v = reg.find(pc + "weight_v");
    if (g == nullptr || v == nullptr) return fail(kUnknownTensor, "missing positional conv weight-norm tensors");
    if (v->shape != std::vector<int64_t>({D, Dg, taps}) || g->numel() != static_cast<size_t>(taps))
      return fail(kInvalidArgument, "positional conv weight shapes do not match the config");
    float* scale;
    SVT_TRY(pool.alloc_t<float>(taps, &scale));
    SVT_TRY(weight_norm_scale(v->dev, g->dev, D, Dg, taps, scale, 0));
    SVT_TRY(pack_pos(v, scale, &e->pos_w));
    SVT_TRY(pack_vec(pool, reg, pc + "bias", D, 1.f, &e->pos_b));
  }
  SVT_TRY(pack_norm(pool, reg, "encoder.layer_norm.", D, &e->enc_norm));

  const int H = c.num_heads, dh = D / H;
  const float qscale = 1.0f / std::sqrt(static_cast<float>(dh));
  for (int l = 0; l < c.num_layers; ++l) {
    const std::string p = "encoder.layers." + std::to_string(l) + ".";
    svt_encoder::Layer L;
    SVT_TRY(pack_norm(pool, reg, p + "layer_norm.", D, &L.ln1));
    SVT_TRY(pack_norm(pool, reg, p + "final_layer_norm.", D, &L.ln2));
    // QKV concatenation, q rows and bias pre-scaled by d_h^-0.5 (HF:513 scales q before QK^T)
    L.qkv.N = 3 * D; L.qkv.K = D;
    SVT_TRY(pool.alloc_t<__nv_bfloat16>(static_cast<size_t>(3) * D * D, &L.qkv.w));
    SVT_TRY(pool.alloc_t<float>(static_cast<size_t>(3) * D, &L.qkv.b));
    const char* names[3] = {"q_proj.", "k_proj.", "v_proj."};
    for (int j = 0; j < 3; ++j) {
      const float sc = (j == 0) ? qscale : 1.f;
      const RawTensor *w, *b;
      SVT_TRY(reg.require(p + "attention." + names[j] + "weight", {D, D}, &w));
      SVT_TRY(reg.require(p + "attention." + names[j] + "bias", {D}, &b));
      SVT_TRY(pack_rows(w, D, D, sc, L.qkv.w, j * D));
      PackArgs a;
      a.src = b->dev; a.dims[3] = D; a.strides[3] = 1; a.scale = sc;
      SVT_TRY(pack_f32(a, L.qkv.b + static_cast<size_t>(j) * D, 0));
    }
    SVT_TRY(pack_linear(pool, reg, p + "attention.out_proj.", D, D, &L.out));
    SVT_TRY(pack_linear(pool, reg, p + "feed_forward.intermediate_dense.", F, D, &L.ff1));
    SVT_TRY(pack_linear(pool, reg, p + "feed_forward.output_dense.", D, F, &L.ff2));
    if (c.rel_pos_buckets > 0) {
      // WavLM gate: (Linear(64 -> 8)).view(2, 4).sum(-1) == Linear(64 -> 2) with the 4-row groups of W / b summed
      const RawTensor *gw, *gb, *gc;
      SVT_TRY(reg.require(p + "attention.gru_rel_pos_linear.weight", {8, dh}, &gw));
      SVT_TRY(reg.require(p + "attention.gru_rel_pos_linear.bias", {8}, &gb));
      gc = reg.find(p + "attention.gru_rel_pos_const");
      if (gc == nullptr || gc->numel() != static_cast<size_t>(H)) return fail(kUnknownTensor, "missing " + p + "attention.gru_rel_pos_const");
      std::vector<float> hw(8 * dh), hb8(8), w2(2 * dh, 0.f), b2(2, 0.f);
      SVT_CUDA(cudaMemcpy(hw.data(), gw->dev, sizeof(float) * hw.size(), cudaMemcpyDeviceToHost));
      SVT_CUDA(cudaMemcpy(hb8.data(), gb->dev, sizeof(float) * 8, cudaMemcpyDeviceToHost));
      for (int r = 0; r < 8; ++r) {
        for (int k = 0; k < dh; ++k) w2[(r / 4) * dh + k] += hw[r * dh + k];
        b2[r / 4] += hb8[r];
      }
      SVT_TRY(pool.alloc_t<float>(w2.size(), &L.gate_w2));
      SVT_TRY(pool.alloc_t<float>(2, &L.gate_b2));
      SVT_TRY(pool.alloc_t<float>(H, &L.gate_const));
      SVT_CUDA(cudaMemcpy(L.gate_w2, w2.data(), sizeof(float) * w2.size(), cudaMemcpyHostToDevice));
      SVT_CUDA(cudaMemcpy(L.gate_b2, b2.data(), sizeof(float) * 2, cudaMemcpyHostToDevice));
      SVT_CUDA(cudaMemcpy(L.gate_const, gc->dev, sizeof(float) * H, cudaMemcpyDeviceToDevice));
      if (transformer_rowstats_bytes(e, 1) > 0) {
        // the same gate on un-normalised rows (LayerNorm fold): per head w2 o gamma_h, its row sums, w2 . beta_h + b2
        std::vector<float> gam(D), bet(D), w2g(static_cast<size_t>(H) * 2 * dh), cg(2 * H, 0.f), dg(2 * H);
        SVT_CUDA(cudaMemcpy(gam.data(), L.ln1.g, sizeof(float) * D, cudaMemcpyDeviceToHost));
        SVT_CUDA(cudaMemcpy(bet.data(), L.ln1.b, sizeof(float) * D, cudaMemcpyDeviceToHost));
        for (int hh = 0; hh < H; ++hh)
          for (int r = 0; r < 2; ++r) {
            double csum = 0.0, dsum = b2[r];
            for (int k = 0; k < dh; ++k) {
              const float wg = w2[r * dh + k] * gam[hh * dh + k];
              w2g[(static_cast<size_t>(hh) * 2 + r) * dh + k] = wg;
              csum += wg;
              dsum += static_cast<double>(w2[r * dh + k]) * bet[hh * dh + k];
            }
            cg[2 * hh + r] = static_cast<float>(csum);
            dg[2 * hh + r] = static_cast<float>(dsum);
          }
        SVT_TRY(pool.alloc_t<float>(w2g.size(), &L.gate_w2g));
        SVT_TRY(pool.alloc_t<float>(cg.size(), &L.gate_cg));
        SVT_TRY(pool.alloc_t<float>(dg.size(), &L.gate_dg));
        SVT_CUDA(cudaMemcpy(L.gate_w2g, w2g.data(), sizeof(float) * w2g.size(), cudaMemcpyHostToDevice));
        SVT_CUDA(cudaMemcpy(L.gate_cg, cg.data(), sizeof(float) * cg.size(), cudaMemcpyHostToDevice));
        SVT_CUDA(cudaMemcpy(L.gate_dg, dg.data(), sizeof(float) * dg.size(), cudaMemcpyHostToDevice));
      }
    }
    if (transformer_rowstats_bytes(e, 1) > 0) {
      // second copies of the two Linears that read a LayerNorm, with the norm folded in
      L.qkv_ln.N = 3 * D; L.qkv_ln.K = D;
      SVT_TRY(pool.alloc_t<__nv_bfloat16>(static_cast<size_t>(3) * D * D, &L.qkv_ln.w));
      SVT_TRY(pool.alloc_t<float>(static_cast<size_t>(3) * D, &L.qkv_ln.b));
      SVT_TRY(pool.alloc_t<float>(static_cast<size_t>(3) * D, &L.qkv_ln.colsum));
      SVT_CUDA(cudaMemcpy(L.qkv_ln.b, L.qkv.b, sizeof(float) * 3 * D, cudaMemcpyDeviceToDevice));
      for (int j = 0; j < 3; ++j) {
        const RawTensor* w;
        SVT_TRY(reg.require(p + "attention." + names[j] + "weight", {D, D}, &w));
        SVT_TRY(fold_norm_rows(w, D, D, (j == 0) ? qscale : 1.f, L.ln1, &L.qkv_ln, j * D));
      }
      L.ff1_ln.N = F; L.ff1_ln.K = D;
      SVT_TRY(pool.alloc_t<__nv_bfloat16>(static_cast<size_t>(F) * D, &L.ff1_ln.w));
      SVT_TRY(pool.alloc_t<float>(static_cast<size_t>(F), &L.ff1_ln.b));
      SVT_TRY(pool.alloc_t<float>(static_cast<size_t>(F), &L.ff1_ln.colsum));
      SVT_CUDA(cudaMemcpy(L.ff1_ln.b, L.ff1.b, sizeof(float) * F, cudaMemcpyDeviceToDevice));
      const RawTensor* w;
      SVT_TRY(reg.require(p + "feed_forward.intermediate_dense.weight", {F, D}, &w));
      SVT_TRY(fold_norm_rows(w, F, D, 1.f, L.ln2, &L.ff1_ln, 0));
    }
    e->layers.push_back(L);
  }
  e->drop_rel_tabs();  // built from the previous weights
  e->rel_embed.clear();
  if (c.rel_pos_buckets > 0) {
    const RawTensor* emb;
    SVT_TRY(reg.require("encoder.layers.0.attention.rel_attn_embed.weight", {c.rel_pos_buckets, H}, &emb));
    e->rel_embed.resize(emb->numel());
    SVT_CUDA(cudaMemcpy(e->rel_embed.data(), emb->dev, sizeof(float) * emb->numel(), cudaMemcpyDeviceToHost));
  }
  SVT_CUDA(cudaDeviceSynchronize());
  e->reg.clear();  // fp32 staging copies no longer needed
  e->finalized = true;
  return kOk;
}

// ------------------------------------------------------------------------------------ workspace plan
namespace {
struct EncPlan {
  int B, L, T0a, Tn, Tna, M;  // Tn: valid output frames, Tna: allocated frames per clip, M = B * Tna
  size_t off_stats, off_chan, off_bufA, off_bufB, off_h, off_hb, off_qkv, off_ctx, off_mid, off_pre, off_rowstats, off_gate, total;
};
EncPlan make_plan(const svt_encoder* e, int B, int L) {
  const svt_encoder_config& c = e->cfg;
  EncPlan p{};
  p.B = B; p.L = L;
  p.T0a = e->t_alloc0(L);
  int ta = p.T0a;
  for (int i = 1; i < c.num_conv_layers; ++i) ta /= c.conv_stride[i];
  p.Tna = ta;
  p.Tn = e->conv_out_len(L, c.num_conv_layers - 1);
  p.M = B * p.Tna;
  const size_t C = c.conv_dim, D = c.hidden_size, F = c.ffn_size;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
  p.off_stats = take(64 + sizeof(double) * 4 * static_cast<size_t>(B));  // [B][2] input + [B][2] output statistics
  p.off_chan = take(sizeof(double) * 2 * C * B);
  const size_t slack = 2 * C * 8 * 2;  // overlapping conv rows read up to (k - stride) * C elements past the end
  p.off_bufA = take(static_cast<size_t>(B) * p.T0a * C * 2 + slack);
  const int T1a = p.T0a / (c.num_conv_layers > 1 ? c.conv_stride[1] : 1);
  p.off_bufB = take(static_cast<size_t>(B) * T1a * C * 2 + slack);
  p.off_h = take(static_cast<size_t>(p.M) * D * 4);
  p.off_hb = take(static_cast<size_t>(p.M) * D * 2);
  p.off_qkv = take(static_cast<size_t>(p.M) * 3 * D * 2);
  p.off_ctx = take(static_cast<size_t>(p.M) * D * 2);
  p.off_mid = take(static_cast<size_t>(p.M) * F * 2);
  p.off_pre = take(static_cast<size_t>(p.M) * D * 4);
  p.off_rowstats = take(transformer_rowstats_bytes(e, static_cast<size_t>(p.M)));
  p.off_gate = take(c.rel_pos_buckets > 0 ? sizeof(float) * static_cast<size_t>(p.M) * c.num_heads : 0);
  p.total = off;
  return p;
}

int linear(const __nv_bfloat16* a, int M, const LinearW& w, const float* resid, float* out_f32, __nv_bfloat16* out_bf16,
           int act, cudaStream_t s, float* row_stats_out = nullptr, const float* ln_stats = nullptr, float ln_eps = 0.f,
           const __nv_bfloat16* resid_bf16 = nullptr) {
  GemmArgs g;
  g.row_stats_out = row_stats_out;
  g.resid_bf16 = resid_bf16;
  if (ln_stats != nullptr) { g.ln_stats = ln_stats; g.ln_colsum = w.colsum; g.ln_eps = ln_eps; }
  g.a = a;
  g.a_dims[0] = w.K; g.a_dims[1] = 1; g.a_dims[2] = M;
  g.a_strides[0] = w.K; g.a_strides[1] = w.K;
  g.w = w.w; g.w_rows = w.N; g.w_cols = w.K;
  g.M = M; g.N = w.N; g.K = w.K; g.k_inner = w.K;
  g.bias = w.b; g.resid = resid; g.out_f32 = out_f32; g.out_bf16 = out_bf16; g.ld_out = w.N; g.act = act;
  return gemm_bf16_tc(g, s);
}
}  // namespace

// ------------------------------------------------------------------------------------ transformer body
// positional conv embedding + N encoder layers + final / per-layer LayerNorms (HF:658-803), shared by the audio
// encoder and the AV-HuBERT video stream (fairseq TransformerEncoder, same graph).  In: h (fp32 residual stream) and
// hb (its bf16 copy), rows = clip * Ta + t.  Out: *final_x points at the fp32 rows to normalise / hand to the head;
// if stats_out != nullptr it receives sum / sum-of-squares of those rows over t < T (whole-tensor output norm).
namespace svt {
// row stride of the WavLM position-bias table of a T-frame clip: 2T - 1 entries + one key block of zeros, 16-byte multiple
static int rel_tab_stride(int T) { return (2 * T - 1 + 128 + 3) / 4 * 4; }
size_t transformer_rowstats_bytes(const svt_encoder* e, size_t M) {
  // two [M][D / 128][2] fp32 buffers; 0 when the LayerNorm fold does not apply (post-LN model or unsupported width)
  const int D = e->cfg.hidden_size;
  const bool ok = e->cfg.stable_layer_norm && D % 256 == 0 && D <= 1024;
  return ok ? sizeof(float) * 2 * M * static_cast<size_t>(D / 128) * 2 : 0;
}
int encoder_transformer_forward(const svt_encoder* e, int B, int T, int Ta, const TransformerBuffers& tb, double* stats_out,
                                int stats_stride, const float** final_x_out, cudaStream_t s) {
  const svt_encoder_config& c = e->cfg;
  const int D = c.hidden_size, H = c.num_heads, dh = D / H;
  const int M = B * Ta;
  const float eps = c.layer_norm_eps;
  float* h = tb.h;
  __nv_bfloat16* hb = tb.hb;
  __nv_bfloat16* qkv = tb.qkv;
  __nv_bfloat16* ctx = tb.ctx;
  __nv_bfloat16* mid = tb.mid;
  float* pre = tb.pre;
  const bool want_stats = stats_out != nullptr;
  auto pos_conv = [&](const __nv_bfloat16* a, const __nv_bfloat16* w, const float* bias) {
    GemmArgs g;
    g.mode = 1;
    g.a = a;
    g.a_dims[0] = D; g.a_dims[1] = T; g.a_dims[2] = B;
    g.a_strides[0] = D; g.a_strides[1] = static_cast<uint64_t>(Ta) * D;
    g.w = w; g.w_rows = c.pos_conv_groups * c.pos_conv_kernel * 64; g.w_cols = 64;
    g.N = D; g.K = c.pos_conv_kernel * 64;
    g.n_clips = B; g.clip_rows = Ta; g.clip_valid = T; g.pad_left = c.pos_conv_kernel / 2; g.taps = c.pos_conv_kernel;
    g.group_size = D / c.pos_conv_groups;
    g.bias = bias; g.ld_out = D;
    return g;
  };
  if (c.pos_conv_layers == 0) {
    // ---- positional conv embedding + residual: h += GELU(conv(h) + b)   (HuBERT conv_pos_batch_norm: conv(BN(h)))
    if (c.pos_conv_batch_norm) SVT_TRY(channel_affine_bf16(h, e->pos_bn_scale, e->pos_bn_shift, hb, M, D, s));
    GemmArgs g = pos_conv(hb, e->pos_w, e->pos_b);
    g.resid = h; g.out_f32 = h; g.act = kActGelu;
    SVT_TRY(gemm_bf16_tc(g, s));
  } else {
    // ---- data2vec-audio: p <- GELU(LN(conv(p) + b)) pos_conv_layers times starting from h, then h += p.
    // conv output in fp32 (pre), the normalised rows ping-pong through two free bf16 buffers (ctx, qkv)
    SVT_CUDA(cudaMemsetAsync(pre, 0, static_cast<size_t>(M) * D * 4, s));  // rows t >= T stay zero
    const __nv_bfloat16* a = hb;
    for (int i = 0; i < c.pos_conv_layers; ++i) {
      const bool last = i == c.pos_conv_layers - 1;
      GemmArgs g = pos_conv(a, e->pos_stack[i].w, e->pos_stack[i].b);
      g.out_f32 = pre; g.act = kActNone;
      SVT_TRY(gemm_bf16_tc(g, s));
      __nv_bfloat16* nxt = (i & 1) ? qkv : ctx;
      LayerNormArgs ln;
      ln.x_f32 = pre; ln.gamma = e->ones; ln.beta = e->zeros; ln.rows = M; ln.D = D; ln.eps = 1e-5f; ln.gelu = 1;
      ln.y_bf16 = last ? nullptr : nxt; ln.y_f32 = last ? pre : nullptr;
      SVT_TRY(layer_norm(ln, s));
      a = nxt;
    }
    SVT_TRY(add_f32(h, pre, h, static_cast<size_t>(M) * D, s));
  }
  SVT_CUDA(cudaMemsetAsync(ctx, 0, static_cast<size_t>(M) * D * 2, s));  // rows t >= T are never written by attention

  auto ln_rows = [&](const float* x, const NormW& w, __nv_bfloat16* yb, float* yf, double* stats) {
    LayerNormArgs ln;
    ln.x_f32 = x; ln.gamma = w.g; ln.beta = w.b; ln.y_bf16 = yb; ln.y_f32 = yf;
    ln.rows = M; ln.D = D; ln.eps = eps; ln.stats = stats; ln.clip_rows = Ta; ln.clip_valid = T;
    ln.stats_stride = stats_stride;
    return layer_norm(ln, s);
  };
  // WavLM: Toeplitz position-bias table of this T (built once per T on the host, cached on the device) and, per
  // layer, the per-row gates from the attention input rows
  const float* rel_tab = nullptr;
  if (c.rel_pos_buckets > 0) {
    auto it = e->rel_tabs.find(T);
    if (it == e->rel_tabs.end()) {
      const int stride = rel_tab_stride(T);
      std::vector<float> tab(static_cast<size_t>(H) * stride, 0.f);  // zero tail: the kernels read whole key blocks
      for (int d = -(T - 1); d <= T - 1; ++d) {
        const int bucket = wavlm_relative_bucket(d, c.rel_pos_buckets, c.rel_pos_max_distance);
        for (int hh = 0; hh < H; ++hh)
          tab[static_cast<size_t>(hh) * stride + d + T - 1] = e->rel_embed[static_cast<size_t>(bucket) * H + hh];
      }
      if (e->rel_tabs.size() >= svt_encoder::kMaxRelTabs) {  // many distinct clip lengths: start the cache over
        SVT_CUDA(cudaStreamSynchronize(s));
        e->drop_rel_tabs();
      }
      float* dev = nullptr;
      SVT_CUDA(cudaMalloc(&dev, sizeof(float) * tab.size()));
      SVT_CUDA(cudaMemcpyAsync(dev, tab.data(), sizeof(float) * tab.size(), cudaMemcpyHostToDevice, s));
      SVT_CUDA(cudaStreamSynchronize(s));  // `tab` is pageable host memory going out of scope
      it = e->rel_tabs.emplace(T, dev).first;
    }
    rel_tab = it->second;
  }
  auto attend = [&](const svt_encoder::Layer& Lw, const __nv_bfloat16* x_rows, const float* x_stats = nullptr) {
    AttentionArgs a;
    a.q = qkv; a.k = qkv + D; a.v = qkv + 2 * D; a.o = ctx;
    a.ldq = a.ldk = a.ldv = 3 * D; a.ldo = D;
    a.Tq = T; a.Tk = T; a.q_clip_rows = Ta; a.k_clip_rows = Ta; a.clips = B; a.heads = H; a.head_dim = dh;
    if (rel_tab != nullptr) {
      if (x_stats != nullptr)  // x_rows are un-normalised (LayerNorm fold): normalise inside the gate
        SVT_TRY(wavlm_gate_ln(x_rows, x_stats, M, H, Lw.gate_w2g, Lw.gate_cg, Lw.gate_dg, Lw.gate_const, eps, tb.gate, s));
      else
        SVT_TRY(wavlm_gate(x_rows, M, H, Lw.gate_w2, Lw.gate_b2, Lw.gate_const, tb.gate, s));
      a.rel_tab = rel_tab; a.rel_tab_stride = rel_tab_stride(T); a.gate = tb.gate;
    }
    return attention_bf16(a, s);
  };
  if (want_stats) SVT_CUDA(cudaMemsetAsync(stats_out, 0, 2 * sizeof(double) * (stats_stride > 0 ? B : 1), s));
  const float* final_x = nullptr;

  if (c.stable_layer_norm && get_option_ln_fold() != 0 && c.num_layers > 0 && transformer_rowstats_bytes(e, M) > 0) {
    // pre-LN layers (HF:612-655) with both per-layer LayerNorms folded around the GEMMs: the GEMM that writes the
    // residual stream also writes its bf16 copy and per-row (sum, sum of squares); the GEMM that reads LN(h) runs on
    // the un-normalised copy with gamma folded into W and finishes the normalisation in its epilogue.
    float* st = tb.rowstats;                                    // statistics of the rows entering a layer
    float* st_mid = st + 2 * static_cast<size_t>(M) * (D / 128);  // ... and of the rows after the attention block
    SVT_TRY(row_stats_cast(h, M, D, hb, st, s));
    // Option "resid_bf16" (measurement only, default off): the residual stream lives in its bf16 copy alone -- the two
    // residual GEMMs read and write 2 instead of 4 + 2 bytes per element -- at the price of one bf16 rounding per residual
    // add (profiles/r2_resid_bf16_experiment.txt: the logits error doubles, so it is not the default).
    const bool rb = get_option_resid_bf16() != 0;
    for (int l = 0; l < c.num_layers; ++l) {
      const svt_encoder::Layer& Lw = e->layers[l];
      SVT_TRY(linear(hb, M, Lw.qkv_ln, nullptr, nullptr, qkv, kActNone, s, nullptr, st, eps));
      SVT_TRY(attend(Lw, hb, st));
      if (rb) SVT_TRY(linear(ctx, M, Lw.out, nullptr, nullptr, hb, kActNone, s, st_mid, nullptr, 0.f, hb));
      else SVT_TRY(linear(ctx, M, Lw.out, h, h, hb, kActNone, s, st_mid));
      SVT_TRY(linear(hb, M, Lw.ff1_ln, nullptr, nullptr, mid, kActGelu, s, nullptr, st_mid, eps));
      if (rb) SVT_TRY(linear(mid, M, Lw.ff2, nullptr, nullptr, hb, kActNone, s, st, nullptr, 0.f, hb));
      else SVT_TRY(linear(mid, M, Lw.ff2, h, h, hb, kActNone, s, st));
    }
    if (rb) {
      LayerNormArgs ln;
      ln.x_bf16 = hb; ln.gamma = e->enc_norm.g; ln.beta = e->enc_norm.b; ln.y_f32 = pre;
      ln.rows = M; ln.D = D; ln.eps = eps; ln.stats = want_stats ? stats_out : nullptr; ln.clip_rows = Ta; ln.clip_valid = T;
      ln.stats_stride = stats_stride;
      SVT_TRY(layer_norm(ln, s));
    } else {
      SVT_TRY(ln_rows(h, e->enc_norm, nullptr, pre, want_stats ? stats_out : nullptr));
    }
    final_x = pre;
  } else if (c.stable_layer_norm) {
    // pre-LN layers (HF:612-655) + final encoder LN (HF:792)
    for (int l = 0; l < c.num_layers; ++l) {
      const svt_encoder::Layer& Lw = e->layers[l];
      SVT_TRY(ln_rows(h, Lw.ln1, hb, nullptr, nullptr));
      SVT_TRY(linear(hb, M, Lw.qkv, nullptr, nullptr, qkv, kActNone, s));
      SVT_TRY(attend(Lw, hb));
      SVT_TRY(linear(ctx, M, Lw.out, h, h, nullptr, kActNone, s));
      SVT_TRY(ln_rows(h, Lw.ln2, hb, nullptr, nullptr));
      SVT_TRY(linear(hb, M, Lw.ff1, nullptr, nullptr, mid, kActGelu, s));
      SVT_TRY(linear(mid, M, Lw.ff2, h, h, nullptr, kActNone, s));
    }
    SVT_TRY(ln_rows(h, e->enc_norm, nullptr, pre, want_stats ? stats_out : nullptr));
    final_x = pre;
  } else {
    // post-LN layers (HF:576-609), encoder LN before the stack (HF:692)
    const bool no_layers = c.num_layers == 0;
    SVT_TRY(ln_rows(h, e->enc_norm, hb, h, (want_stats && no_layers) ? stats_out : nullptr));
    for (int l = 0; l < c.num_layers; ++l) {
      const svt_encoder::Layer& Lw = e->layers[l];
      const bool last = l == c.num_layers - 1;
      SVT_TRY(linear(hb, M, Lw.qkv, nullptr, nullptr, qkv, kActNone, s));
      SVT_TRY(attend(Lw, hb));
      SVT_TRY(linear(ctx, M, Lw.out, h, h, nullptr, kActNone, s));
      SVT_TRY(ln_rows(h, Lw.ln1, hb, h, nullptr));
      SVT_TRY(linear(hb, M, Lw.ff1, nullptr, nullptr, mid, kActGelu, s));
      SVT_TRY(linear(mid, M, Lw.ff2, h, h, nullptr, kActNone, s));
      SVT_TRY(ln_rows(h, Lw.ln2, hb, h, (want_stats && last) ? stats_out : nullptr));
    }
    final_x = h;
  }
  *final_x_out = final_x;
  return kOk;
}
}  // namespace svt

// ------------------------------------------------------------------------------------ forward
static int forward_impl(svt_encoder* e, const float* wav, int B, int L, void* ws, size_t ws_bytes, float* feats,
                        float* logits, cudaStream_t s) {
  const svt_encoder_config& c = e->cfg;
  if (!e->finalized) return fail(kNotFinalized, "svt_encoder_finalize has not been called");
  if (B <= 0 || L <= 0) return fail(kInvalidArgument, "batch and n_samples must be positive");
  const EncPlan p = make_plan(e, B, L);
  if (p.Tn <= 0) return fail(kInvalidArgument, "input too short for the conv stack");
  if (ws_bytes < p.total) return fail(kWorkspaceTooSmall, "workspace too small: need " + std::to_string(p.total));
  if (logits != nullptr && e->head_w == nullptr) return fail(kInvalidArgument, "logits requested but no head set");
  uint8_t* base = static_cast<uint8_t*>(ws);
  double* stats_in = reinterpret_cast<double*>(base + p.off_stats);
  double* stats_out = stats_in + 2 * static_cast<size_t>(B);
  // a single clip is its own normalisation scope either way: route it through the per-clip statistics so that its
  // result is bit-identical to the same clip inside a larger per-clip batch
  const int stats_stride = (e->norm_per_clip || B == 1) ? 2 : 0;
  double* chan = reinterpret_cast<double*>(base + p.off_chan);
  __nv_bfloat16* bufA = reinterpret_cast<__nv_bfloat16*>(base + p.off_bufA);
  __nv_bfloat16* bufB = reinterpret_cast<__nv_bfloat16*>(base + p.off_bufB);
  float* h = reinterpret_cast<float*>(base + p.off_h);
  __nv_bfloat16* hb = reinterpret_cast<__nv_bfloat16*>(base + p.off_hb);
  __nv_bfloat16* qkv = reinterpret_cast<__nv_bfloat16*>(base + p.off_qkv);
  __nv_bfloat16* ctx = reinterpret_cast<__nv_bfloat16*>(base + p.off_ctx);
  __nv_bfloat16* mid = reinterpret_cast<__nv_bfloat16*>(base + p.off_mid);
  float* pre = reinterpret_cast<float*>(base + p.off_pre);
  const int C = c.conv_dim, D = c.hidden_size;
  const float eps = c.layer_norm_eps;

  // ---- A1 + conv layer 0 (fused input normalisation)
  if (c.normalize_wav) {
    if (stats_stride > 0) SVT_TRY(tensor_stats_per_clip(wav, B, static_cast<size_t>(L), stats_in, s));
    else SVT_TRY(tensor_stats(wav, static_cast<size_t>(B) * L, stats_in, s));
  }
  Conv0Args c0;
  c0.wav = wav; c0.B = B; c0.L = L; c0.T = e->conv_out_len(L, 0); c0.t_alloc = p.T0a;
  c0.C = C; c0.k = c.conv_kernel[0]; c0.stride = c.conv_stride[0];
  c0.w = e->conv0_w; c0.bias = e->conv0_b; c0.gamma = e->conv0_norm.g; c0.beta = e->conv0_norm.b;
  c0.in_stats = c.normalize_wav ? stats_in : nullptr;
  c0.stats_stride = stats_stride;
  c0.out = bufA; c0.layer_mode = c.feat_norm_layer; c0.chan_stats = chan; c0.tc_tables = e->conv0_tab;
  SVT_TRY(conv0_forward(c0, s));
  if (!c.feat_norm_layer)
    SVT_TRY(groupnorm_gelu_apply(bufA, chan, e->conv0_norm.g, e->conv0_norm.b, B, c0.T, p.T0a, C, s));

  // ---- conv layers 1..n-1 as implicit GEMMs over overlapping channel-last rows
  __nv_bfloat16* cur = bufA;
  __nv_bfloat16* nxt = bufB;
  int ta = p.T0a;
  for (int i = 1; i < c.num_conv_layers; ++i) {
    const int k = c.conv_kernel[i], st = c.conv_stride[i];
    const int ta_out = ta / st;
    const int M = B * ta_out;
    GemmArgs g;
    g.a = cur;
    g.a_dims[0] = C; g.a_dims[1] = k; g.a_dims[2] = M;
    g.a_strides[0] = C; g.a_strides[1] = static_cast<uint64_t>(st) * C;
    g.w = e->conv[i - 1].w; g.w_rows = C; g.w_cols = k * C;
    g.M = M; g.N = C; g.K = k * C; g.k_inner = C;
    g.bias = e->conv[i - 1].b; g.out_bf16 = nxt; g.ld_out = C;
    g.act = c.feat_norm_layer ? kActNone : kActGelu;
    bool fused_ln = false;
    if (c.feat_norm_layer && get_option_rowln_fuse() != 0) {
      // conv -> LN(512) -> GELU in one kernel (the rows are normalised in place while still in L2)
      g.rowln_gamma = e->conv_norm[i - 1].g; g.rowln_beta = e->conv_norm[i - 1].b; g.rowln_eps = 1e-5f; g.rowln_gelu = 1;
      fused_ln = gemm_rowln_supported(g);
      if (!fused_ln) g.rowln_gamma = g.rowln_beta = nullptr;
    }
    SVT_TRY(gemm_bf16_tc(g, s));
    if (c.feat_norm_layer && !fused_ln) {
      LayerNormArgs ln;
      ln.x_bf16 = nxt; ln.y_bf16 = nxt; ln.gamma = e->conv_norm[i - 1].g; ln.beta = e->conv_norm[i - 1].b;
      ln.rows = M; ln.D = C; ln.eps = 1e-5f; ln.gelu = 1;
      SVT_TRY(layer_norm(ln, s));
    }
    std::swap(cur, nxt);
    ta = ta_out;
  }
  const int M = p.M, T = p.Tn, Ta = p.Tna;

  // ---- feature projection: LN(512) -> Linear(512 -> D); fp32 residual stream + bf16 copy for the pos-conv
  {
    if (c.feat_proj_norm) {
      LayerNormArgs ln;
      ln.x_bf16 = cur; ln.y_bf16 = cur; ln.gamma = e->proj_norm.g; ln.beta = e->proj_norm.b;
      ln.rows = M; ln.D = C; ln.eps = eps;
      SVT_TRY(layer_norm(ln, s));
    }
    SVT_TRY(linear(cur, M, e->proj, nullptr, h, hb, kActNone, s));
  }
  // ---- positional conv + transformer layers (shared with the video stream)
  const bool want_stats = c.output_norm != 0;
  const float* final_x = nullptr;
  {
    TransformerBuffers tb;
    tb.h = h; tb.hb = hb; tb.qkv = qkv; tb.ctx = ctx; tb.mid = mid; tb.pre = pre;
    tb.rowstats = reinterpret_cast<float*>(base + p.off_rowstats);
    tb.gate = reinterpret_cast<float*>(base + p.off_gate);
    SVT_TRY(encoder_transformer_forward(e, B, T, Ta, tb, want_stats ? stats_out : nullptr, stats_stride, &final_x, s));
  }
  // ---- A7 whole-tensor output norm + head
  HeadArgs ha;
  ha.x = final_x; ha.clips = B; ha.clip_rows = Ta; ha.T = T; ha.D = D;
  ha.stats = want_stats ? stats_out : nullptr; ha.eps = 1e-5f; ha.stats_stride = stats_stride;
  ha.w = logits != nullptr ? e->head_w : nullptr; ha.b = e->head_b; ha.n_out = e->head_n;
  ha.feats = feats; ha.logits = logits;
  if (feats != nullptr || logits != nullptr) SVT_TRY(head_forward(ha, s));
  return kOk;
}

// ------------------------------------------------------------------------------------ C ABI
extern "C" {

int svt_encoder_create(const svt_encoder_config* cfg, svt_encoder** out) {
  if (cfg == nullptr || out == nullptr) return fail(kInvalidArgument, "null argument");
  const svt_encoder_config& c = *cfg;
  if (c.num_conv_layers < 2 || c.num_conv_layers > SVT_MAX_CONV_LAYERS) return fail(kInvalidArgument, "num_conv_layers out of range");
  if (c.conv_dim != 512) return fail(kUnsupported, "conv_dim must be 512");
  if (c.hidden_size % 128 != 0 || c.hidden_size > 2048) return fail(kUnsupported, "hidden_size must be a multiple of 128, <= 2048");
  if (c.num_heads <= 0 || c.hidden_size % c.num_heads != 0) return fail(kInvalidArgument, "bad num_heads");
  const int dh = c.hidden_size / c.num_heads;
  if (dh != 64 && dh != 128) return fail(kUnsupported, "head dim must be 64 or 128");
  if (c.ffn_size % 64 != 0) return fail(kUnsupported, "ffn_size must be a multiple of 64");
  if (c.pos_conv_groups <= 0 || c.hidden_size % c.pos_conv_groups != 0) return fail(kInvalidArgument, "bad pos_conv_groups");
  if (c.pos_conv_kernel <= 0 || c.pos_conv_layers < 0 || c.pos_conv_layers > 16) return fail(kInvalidArgument, "bad positional conv kernel / depth");
  if (c.rel_pos_buckets < 0 || c.rel_pos_buckets % 4 != 0 || (c.rel_pos_buckets > 0 && (c.rel_pos_max_distance <= c.rel_pos_buckets / 4 || dh != 64)))
    return fail(kInvalidArgument, "bad relative position bias configuration (needs head dim 64)");
  const int dg = c.hidden_size / c.pos_conv_groups;
  if (dg > 64 || dg % 16 != 0) return fail(kUnsupported, "positional conv channels per group must be a multiple of 16, <= 64");
  for (int i = 0; i < c.num_conv_layers; ++i)
    if (c.conv_kernel[i] <= 0 || c.conv_stride[i] <= 0 || c.conv_kernel[i] < c.conv_stride[i])
      return fail(kInvalidArgument, "bad conv kernel/stride");
  svt_encoder* e = new svt_encoder();
  e->cfg = c;
  *out = e;
  return kOk;
}

void svt_encoder_destroy(svt_encoder* enc) { delete enc; }

int svt_encoder_set_tensor(svt_encoder* enc, const char* name, const float* host, const int64_t* shape, int ndim,
                           int strict) {
  if (enc == nullptr || name == nullptr || host == nullptr) return fail(kInvalidArgument, "null argument");
  if (svt_device_count() <= 0) return fail(kNoDevice, "no CUDA device");
  std::string n(name);
  if (n.rfind("model.", 0) == 0) n = n.substr(6);
  const bool known = n.rfind("feature_extractor.", 0) == 0 || n.rfind("feature_projection.", 0) == 0 ||
                     n.rfind("encoder.", 0) == 0;
  if (!known) return strict ? fail(kUnknownTensor, "unknown tensor " + n) : static_cast<int>(kOk);
  enc->finalized = false;
  return enc->reg.set(n, host, shape, ndim);
}

int svt_encoder_set_head(svt_encoder* enc, const float* w, const float* b, int n_out) {
  if (enc == nullptr || w == nullptr) return fail(kInvalidArgument, "null argument");
  if (n_out <= 0 || n_out > 32) return fail(kUnsupported, "head n_out must be in 1..32");
  if (svt_device_count() <= 0) return fail(kNoDevice, "no CUDA device");
  const size_t D = enc->cfg.hidden_size;
  if (enc->head_w != nullptr) { cudaFree(enc->head_w); cudaFree(enc->head_b); enc->head_w = enc->head_b = nullptr; }
  SVT_CUDA(cudaMalloc(&enc->head_w, sizeof(float) * n_out * D));
  SVT_CUDA(cudaMalloc(&enc->head_b, sizeof(float) * n_out));
  SVT_CUDA(cudaMemcpy(enc->head_w, w, sizeof(float) * n_out * D, cudaMemcpyHostToDevice));
  if (b != nullptr) SVT_CUDA(cudaMemcpy(enc->head_b, b, sizeof(float) * n_out, cudaMemcpyHostToDevice));
  else SVT_CUDA(cudaMemset(enc->head_b, 0, sizeof(float) * n_out));
  enc->head_n = n_out;
  return kOk;
}

int svt_encoder_set_norm_per_clip(svt_encoder* enc, int per_clip) {
  if (enc == nullptr) return fail(kInvalidArgument, "null argument");
  enc->norm_per_clip = per_clip != 0;
  return kOk;
}

int svt_encoder_finalize(svt_encoder* enc) {
  if (enc == nullptr) return fail(kInvalidArgument, "null argument");
  if (svt_device_count() <= 0) return fail(kNoDevice, "no CUDA device");
  return encoder_finalize(enc);
}

int svt_encoder_num_frames(const svt_encoder* enc, int n_samples) {
  return enc == nullptr ? 0 : enc->conv_out_len(n_samples, enc->cfg.num_conv_layers - 1);
}

size_t svt_encoder_workspace_bytes(const svt_encoder* enc, int batch, int n_samples) {
  if (enc == nullptr || batch <= 0 || n_samples <= 0) return 0;
  return make_plan(enc, batch, n_samples).total;
}

int svt_encoder_forward(svt_encoder* enc, const float* wav_dev, int batch, int n_samples, void* workspace_dev,
                        size_t workspace_bytes, float* feats_dev, float* logits_dev, void* stream) {
  if (enc == nullptr || wav_dev == nullptr || workspace_dev == nullptr) return fail(kInvalidArgument, "null argument");
  return forward_impl(enc, wav_dev, batch, n_samples, workspace_dev, workspace_bytes, feats_dev, logits_dev,
                      static_cast<cudaStream_t>(stream));
}

int svt_encoder_forward_host(svt_encoder* enc, const float* wav_host, int batch, int n_samples, void* workspace_dev,
                             size_t workspace_bytes, float* wav_stage_dev, float* logits_stage_dev,
                             float* logits_host, void* stream) {
  if (enc == nullptr || wav_host == nullptr || wav_stage_dev == nullptr || logits_stage_dev == nullptr ||
      logits_host == nullptr)
    return fail(kInvalidArgument, "null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t in_bytes = sizeof(float) * static_cast<size_t>(batch) * n_samples;
  SVT_CUDA(cudaMemcpyAsync(wav_stage_dev, wav_host, in_bytes, cudaMemcpyHostToDevice, s));
  SVT_TRY(forward_impl(enc, wav_stage_dev, batch, n_samples, workspace_dev, workspace_bytes, nullptr, logits_stage_dev, s));
  const int T = svt_encoder_num_frames(enc, n_samples);
  const size_t out_bytes = sizeof(float) * static_cast<size_t>(batch) * T * enc->head_n;
  SVT_CUDA(cudaMemcpyAsync(logits_host, logits_stage_dev, out_bytes, cudaMemcpyDeviceToHost, s));
  SVT_CUDA(cudaStreamSynchronize(s));
  return kOk;
}

// ------------------------------------------------------------------------------------ host-buffer pipeline
// Serving loop for HOST batches: the H2D copy of batch k + 1 runs on its own stream under the forward of batch k, the
// D2H of the logits follows the forward on the compute stream.  Caller-owned memory, library-owned streams / events.
struct svt_pipeline {
  svt_encoder* enc = nullptr;
  int B = 0, L = 0, T = 0, depth = 0;
  void* ws = nullptr;
  size_t ws_bytes = 0;
  float* wav_stage = nullptr;
  float* logits_stage = nullptr;
  cudaStream_t copy = nullptr, compute = nullptr;
  std::vector<cudaEvent_t> h2d, free_, done;
  long long submitted = 0;
};

int svt_pipeline_create(svt_encoder* enc, int batch, int n_samples, int depth, void* workspace_dev, size_t workspace_bytes,
                        float* wav_stage_dev, float* logits_stage_dev, svt_pipeline** out) {
  if (enc == nullptr || out == nullptr || workspace_dev == nullptr || wav_stage_dev == nullptr || logits_stage_dev == nullptr)
    return fail(kInvalidArgument, "null argument");
  if (!enc->finalized || enc->head_w == nullptr) return fail(kNotFinalized, "pipeline: encoder must be finalized and have a head");
  if (batch <= 0 || n_samples <= 0 || depth < 1 || depth > 8) return fail(kInvalidArgument, "pipeline: bad batch / samples / depth");
  const EncPlan p = make_plan(enc, batch, n_samples);
  if (p.Tn <= 0) return fail(kInvalidArgument, "input too short for the conv stack");
  if (workspace_bytes < p.total) return fail(kWorkspaceTooSmall, "workspace too small: need " + std::to_string(p.total));
  svt_pipeline* pl = new svt_pipeline();
  pl->enc = enc; pl->B = batch; pl->L = n_samples; pl->T = p.Tn; pl->depth = depth;
  pl->ws = workspace_dev; pl->ws_bytes = workspace_bytes; pl->wav_stage = wav_stage_dev; pl->logits_stage = logits_stage_dev;
  if (cudaStreamCreateWithFlags(&pl->copy, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&pl->compute, cudaStreamNonBlocking) != cudaSuccess) {
    delete pl;
    return fail(kCudaError, "pipeline: cannot create streams");
  }
  pl->h2d.resize(depth); pl->free_.resize(depth); pl->done.resize(depth);
  for (int i = 0; i < depth; ++i) {
    cudaEventCreateWithFlags(&pl->h2d[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&pl->free_[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&pl->done[i], cudaEventDisableTiming);
  }
  *out = pl;
  return kOk;
}

void svt_pipeline_destroy(svt_pipeline* pl) {
  if (pl == nullptr) return;
  if (pl->compute != nullptr) cudaStreamSynchronize(pl->compute);
  if (pl->copy != nullptr) cudaStreamSynchronize(pl->copy);
  for (auto e : pl->h2d) cudaEventDestroy(e);
  for (auto e : pl->free_) cudaEventDestroy(e);
  for (auto e : pl->done) cudaEventDestroy(e);
  if (pl->copy != nullptr) cudaStreamDestroy(pl->copy);
  if (pl->compute != nullptr) cudaStreamDestroy(pl->compute);
  delete pl;
}

int svt_pipeline_submit(svt_pipeline* pl, const float* wav_host_pinned, float* logits_host_pinned, long long* ticket) {
  if (pl == nullptr || wav_host_pinned == nullptr || logits_host_pinned == nullptr || ticket == nullptr)
    return fail(kInvalidArgument, "null argument");
  const long long k = pl->submitted;
  const int slot = static_cast<int>(k % pl->depth);
  const size_t in_elems = static_cast<size_t>(pl->B) * pl->L;
  const size_t out_elems = static_cast<size_t>(pl->B) * pl->T * pl->enc->head_n;
  float* wav_dev = pl->wav_stage + static_cast<size_t>(slot) * in_elems;
  float* lg_dev = pl->logits_stage + static_cast<size_t>(slot) * out_elems;
  if (k >= pl->depth) {
    // the slot's previous occupant must have left the device: its forward no longer reads wav_dev, its D2H is complete
    SVT_CUDA(cudaStreamWaitEvent(pl->copy, pl->free_[slot], 0));
    SVT_CUDA(cudaEventSynchronize(pl->done[slot]));
  }
  SVT_CUDA(cudaMemcpyAsync(wav_dev, wav_host_pinned, sizeof(float) * in_elems, cudaMemcpyHostToDevice, pl->copy));
  SVT_CUDA(cudaEventRecord(pl->h2d[slot], pl->copy));
  SVT_CUDA(cudaStreamWaitEvent(pl->compute, pl->h2d[slot], 0));
  SVT_TRY(forward_impl(pl->enc, wav_dev, pl->B, pl->L, pl->ws, pl->ws_bytes, nullptr, lg_dev, pl->compute));
  SVT_CUDA(cudaEventRecord(pl->free_[slot], pl->compute));
  SVT_CUDA(cudaMemcpyAsync(logits_host_pinned, lg_dev, sizeof(float) * out_elems, cudaMemcpyDeviceToHost, pl->compute));
  SVT_CUDA(cudaEventRecord(pl->done[slot], pl->compute));
  *ticket = k;
  pl->submitted = k + 1;
  return kOk;
}

int svt_pipeline_wait(svt_pipeline* pl, long long ticket) {
  if (pl == nullptr) return fail(kInvalidArgument, "null argument");
  if (ticket < 0 || ticket >= pl->submitted) return fail(kInvalidArgument, "pipeline: unknown ticket");
  if (ticket + pl->depth < pl->submitted) return kOk;  // its slot has been reused: that only happens after completion
  SVT_CUDA(cudaEventSynchronize(pl->done[static_cast<int>(ticket % pl->depth)]));
  return kOk;
}

}  // extern "C"
