// AV-HuBERT video stream: lip-ROI clips (B, 1, T, 88, 88) -> features (B, T, D).
// Reference semantics: N20EMv2/video_only/fairseq_interface.py:454-485 (FairseqAVHubertPretrain.forward),
// hubert.py:688-739 (AVHubertModel.extract_finetune with audio = None), hubert.py:311-326 (SubModel),
// resnet.py:37-171 (ResEncoder: Conv3d front end + ResNet-18 trunk with PReLU), and fairseq's wav2vec2
// TransformerEncoder (layer_norm_first), whose graph is the one csrc/encoder.cu already runs.
//
// Everything dense runs on the tcgen05 GEMM (gemm_tc.cu):
//   * Conv3d(1->64, 5x7x7, stride 1x2x2): one gather kernel writes the [frames*44*44, 256] bf16 patch matrix
//     (K = 245 padded to 256), the GEMM applies the BN-folded weights with a bias + PReLU epilogue;
//   * every 3x3 / 1x1 conv of the trunk is an implicit GEMM over feature maps kept as FLAT rows
//     [frame][y][x] x channels with a zero padding ring: tap (dy, dx) is a constant row shift of the TMA
//     coordinate, so no patch matrix exists; BatchNorm is folded into weights + bias, the PReLU, the residual add and
//     the re-zeroing of the padding ring are epilogue options.  Stride-2 convs are evaluated at full resolution and
//     subsampled (three of the seventeen convs; to be replaced by a strided TMA view);
//   * SubModel.proj writes straight into the upper half of the [frames, 2D] concat buffer whose lower (audio) half
//     stays zero (hubert.py:700-708), LayerNorm(2D) and post_extract_proj follow, then the shared transformer body.
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "model.cuh"

using namespace svt;

namespace {

constexpr int kImg = 88;       // lip ROI (reference crops to 88 x 88, video_only/train_video_ssl.py:445-457)
constexpr int kF0 = 44;        // after the stride-2 Conv3d
constexpr int kFrontK = 64;    // 7 * 7 = 49 spatial taps of one frame padded to 64; the 5 temporal taps are row shifts
constexpr int kFrontT = 5;     // temporal taps of the Conv3d
constexpr int kRes[4] = {22, 11, 6, 3};
constexpr int kPlanes[4] = {64, 128, 256, 512};
constexpr float kBnEps = 1e-5f;

struct HostTensor {
  std::vector<float> v;
  std::vector<int64_t> shape;
};

struct ConvW {
  __nv_bfloat16* w = nullptr;  // [cout][taps][cin] (BN scale folded)
  float* bias = nullptr;       // BN shift
  float* alpha = nullptr;      // PReLU slopes applied after this conv (+ residual), or null
  int cin = 0, cout = 0, taps = 0;
};

// ------------------------------------------------------------------------------------------ kernels
// Spatial patch matrix of the Conv3d front end: row (n, y, x) = frame slot n = b * Ta + t, output pixel (y, x) of 44 x 44;
// column k = dy * 7 + dx < 49 holds video[b, t, 2y + dy - 3, 2x + dx - 3] (zero outside the image), optionally
// whole-tensor normalised first; columns 49..63 and frame slots t >= T are zero.  The temporal extent of the kernel
// (5 frames) is NOT unrolled into columns: the GEMM walks it as 5 taps that shift the A rows by whole frames
// (44 * 44 rows), and Ta >= T + 2 keeps two all-zero frame slots between clips, which is the Conv3d's zero padding in
// time.  One CTA = one output row y of one frame slot: the 7 input rows it touches are staged in shared memory.
__global__ void __launch_bounds__(256) frontend_patch_kernel(const float* __restrict__ video, int B, int T, int Ta,
                                                             const double* __restrict__ in_stats, double inv_count,
                                                             __nv_bfloat16* __restrict__ col) {
  __shared__ float tile[7][kImg];
  const int y = blockIdx.x % kF0;
  const int n = blockIdx.x / kF0;
  const int t = n % Ta, b = n / Ta;
  __nv_bfloat16* out = col + (static_cast<size_t>(n) * kF0 + y) * kF0 * kFrontK;
  if (t >= T) {  // padding frame slot: all-zero patches
    for (int i = threadIdx.x; i < kF0 * kFrontK / 8; i += 256) reinterpret_cast<uint4*>(out)[i] = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  float mean = 0.f, rstd = 1.f;
  if (in_stats != nullptr) {
    const double m = in_stats[0] * inv_count;
    const double var = in_stats[1] * inv_count - m * m;
    mean = static_cast<float>(m);
    rstd = static_cast<float>(1.0 / sqrt((var > 0 ? var : 0) + 1e-5));
  }
  for (int i = threadIdx.x; i < 7 * kImg; i += 256) {
    const int xi = i % kImg, dy = i / kImg;
    const int yi = 2 * y + dy - 3;
    float v = 0.f;  // zero padding of the (normalised) input
    if (yi >= 0 && yi < kImg) v = (__ldg(video + ((static_cast<size_t>(b) * T + t) * kImg + yi) * kImg + xi) - mean) * rstd;
    tile[dy][xi] = v;
  }
  __syncthreads();
  // work item = (x, 8-column chunk c8): 44 * 8 items, 16 bytes each, consecutive threads write consecutive chunks
  for (int item = threadIdx.x; item < kF0 * (kFrontK / 8); item += 256) {
    const int x = item >> 3, c8 = item & 7;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = c8 * 8 + i;
      const int dy = k / 7, dx = k - dy * 7;
      const int xi = 2 * x + dx - 3;
      v[i] = (k < 49 && xi >= 0 && xi < kImg) ? tile[dy][xi] : 0.f;
    }
    *reinterpret_cast<uint4*>(out + static_cast<size_t>(x) * kFrontK + c8 * 8) =
        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  }
}

// Evaluation transform of the video recipe (video_only/train_video_ssl.py:454-457, utils.py:45-84): uint8 frames
// (n, H, W) -> /255 -> centre crop -> (x - mean) / std, fp32 (n, crop, crop)
__global__ void __launch_bounds__(256) video_transform_kernel(const uint8_t* __restrict__ frames, long long n, int H, int W,
                                                              int crop, int y0, int x0, float mean, float stdev,
                                                              float* __restrict__ out) {
  const long long total = n * crop * crop;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % crop), y = static_cast<int>((i / crop) % crop);
    const long long f = i / (static_cast<long long>(crop) * crop);
    const float v = static_cast<float>(frames[(f * H + y0 + y) * W + x0 + x]);
    out[i] = (v / 255.0f - mean) / stdev;
  }
}

// MaxPool3d((1,3,3), stride (1,2,2), pad (0,1,1)): [N][44][44][64] -> padded [N][24][24][64] (ring = 0)
__global__ void __launch_bounds__(256) maxpool_kernel(const __nv_bfloat16* __restrict__ f0, int N, __nv_bfloat16* __restrict__ out) {
  constexpr int kR0 = 22, Hp = kR0 + 2, C8 = 64 / 8;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(N) * Hp * Hp * C8) return;
  const int c8 = static_cast<int>(i % C8);
  const int xp = static_cast<int>((i / C8) % Hp), yp = static_cast<int>((i / (C8 * Hp)) % Hp);
  const int n = static_cast<int>(i / (C8 * Hp * Hp));
  uint4 r = make_uint4(0u, 0u, 0u, 0u);
  if (yp >= 1 && yp <= kR0 && xp >= 1 && xp <= kR0) {
    float m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
    for (int dy = 0; dy < 3; ++dy) {
      const int yi = 2 * (yp - 1) - 1 + dy;
      if (yi < 0 || yi >= kF0) continue;
      for (int dx = 0; dx < 3; ++dx) {
        const int xi = 2 * (xp - 1) - 1 + dx;
        if (xi < 0 || xi >= kF0) continue;
        const uint4 q = *reinterpret_cast<const uint4*>(f0 + ((static_cast<size_t>(n) * kF0 + yi) * kF0 + xi) * 64 + c8 * 8);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          m[2 * k] = fmaxf(m[2 * k], __low2float(h[k]));
          m[2 * k + 1] = fmaxf(m[2 * k + 1], __high2float(h[k]));
        }
      }
    }
    r = make_uint4(pack_bf16x2(m[0], m[1]), pack_bf16x2(m[2], m[3]), pack_bf16x2(m[4], m[5]), pack_bf16x2(m[6], m[7]));
  }
  *reinterpret_cast<uint4*>(out + i * 8) = r;
}

// stride-2 sampling of a padded feature map: dst interior (y', x') <- src interior (2y', 2x'); dst ring = 0
__global__ void __launch_bounds__(256) subsample_kernel(const __nv_bfloat16* __restrict__ src, int N, int Hs, int C,
                                                        __nv_bfloat16* __restrict__ dst) {
  const int Hd = (Hs + 1) / 2, Hsp = Hs + 2, Hdp = Hd + 2, C8 = C / 8;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(N) * Hdp * Hdp * C8) return;
  const int c8 = static_cast<int>(i % C8);
  const int xp = static_cast<int>((i / C8) % Hdp), yp = static_cast<int>((i / (static_cast<long long>(C8) * Hdp)) % Hdp);
  const int n = static_cast<int>(i / (static_cast<long long>(C8) * Hdp * Hdp));
  uint4 r = make_uint4(0u, 0u, 0u, 0u);
  if (yp >= 1 && yp <= Hd && xp >= 1 && xp <= Hd)
    r = *reinterpret_cast<const uint4*>(src + ((static_cast<size_t>(n) * Hsp + 2 * (yp - 1) + 1) * Hsp + 2 * (xp - 1) + 1) * C + c8 * 8);
  *reinterpret_cast<uint4*>(dst + i * 8) = r;
}

// 1 = interior pixel, 0 = padding ring, for N padded Hp x Hp maps stored as flat rows
__global__ void __launch_bounds__(256) ring_mask_kernel(uint8_t* __restrict__ mask, long long rows, int Hp) {
  const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int x = static_cast<int>(r % Hp), y = static_cast<int>((r / Hp) % Hp);
  mask[r] = (x >= 1 && x <= Hp - 2 && y >= 1 && y <= Hp - 2) ? 1 : 0;
}

// AdaptiveAvgPool2d(1) over the 3 x 3 interior of padded 5 x 5 maps: [N][5][5][C] -> out[n * ld + c] (bf16)
__global__ void __launch_bounds__(256) avgpool_kernel(const __nv_bfloat16* __restrict__ x, int N, int C, __nv_bfloat16* __restrict__ out,
                                                      int ld) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(N) * C) return;
  const int c = static_cast<int>(i % C), n = static_cast<int>(i / C);
  float s = 0.f;
  for (int y = 1; y <= 3; ++y)
    for (int xx = 1; xx <= 3; ++xx) s += __bfloat162float(x[((static_cast<size_t>(n) * 5 + y) * 5 + xx) * C + c]);
  out[static_cast<size_t>(n) * ld + c] = __float2bfloat16(s * (1.0f / 9.0f));
}

template <typename T>
int upload(DevicePool& pool, const std::vector<T>& h, T** out) {
  SVT_TRY(pool.alloc_t<T>(h.size(), out));
  SVT_CUDA(cudaMemcpy(*out, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
  return kOk;
}

}  // namespace

struct svt_video {
  svt_video_config cfg{};
  std::map<std::string, HostTensor> host;  // front end / glue tensors (fp32, host) until finalize
  svt_encoder* enc = nullptr;              // transformer body (positional conv + layers), transformer_only
  DevicePool pool;
  bool finalized = false;
  ConvW front;                             // Conv3d as [64][5][64]
  struct Block { ConvW c1, c2, ds; bool has_ds = false; };
  Block blocks[4][2];
  LinearW proj, post;
  NormW cat_norm;
  ~svt_video() { delete enc; }
};

namespace {

const HostTensor* find(const svt_video* v, const std::string& name) {
  auto it = v->host.find(name);
  return it == v->host.end() ? nullptr : &it->second;
}
int need(const svt_video* v, const std::string& name, std::vector<int64_t> shape, const HostTensor** out) {
  const HostTensor* t = find(v, name);
  if (t == nullptr) return fail(kUnknownTensor, "missing tensor: " + name);
  if (t->shape != shape) return fail(kInvalidArgument, "shape mismatch for " + name);
  *out = t;
  return kOk;
}

// BatchNorm (eval) folded into scale / shift per output channel
int bn_fold(const svt_video* v, const std::string& p, int c, std::vector<float>* scale, std::vector<float>* shift) {
  const HostTensor *g, *b, *m, *var;
  SVT_TRY(need(v, p + "weight", {c}, &g));
  SVT_TRY(need(v, p + "bias", {c}, &b));
  SVT_TRY(need(v, p + "running_mean", {c}, &m));
  SVT_TRY(need(v, p + "running_var", {c}, &var));
  scale->resize(c);
  shift->resize(c);
  for (int i = 0; i < c; ++i) {
    const float s = g->v[i] / std::sqrt(var->v[i] + kBnEps);
    (*scale)[i] = s;
    (*shift)[i] = b->v[i] - m->v[i] * s;
  }
  return kOk;
}

// conv (cout, cin, k, k) + BN -> bf16 [cout][k*k][cin], bias; optional PReLU slopes
int pack_conv(svt_video* v, const std::string& conv, const std::string& bn, const std::string& prelu, int cin, int cout, int k,
              ConvW* out) {
  const HostTensor* w;
  SVT_TRY(need(v, conv + "weight", {cout, cin, k, k}, &w));
  std::vector<float> scale, shift;
  SVT_TRY(bn_fold(v, bn, cout, &scale, &shift));
  const int taps = k * k;
  std::vector<__nv_bfloat16> pw(static_cast<size_t>(cout) * taps * cin);
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < cin; ++ci)
      for (int t = 0; t < taps; ++t)
        pw[(static_cast<size_t>(co) * taps + t) * cin + ci] = __float2bfloat16(w->v[(static_cast<size_t>(co) * cin + ci) * taps + t] * scale[co]);
  SVT_TRY(upload(v->pool, pw, &out->w));
  SVT_TRY(upload(v->pool, shift, &out->bias));
  if (!prelu.empty()) {
    const HostTensor* a;
    SVT_TRY(need(v, prelu + "weight", {cout}, &a));
    SVT_TRY(upload(v->pool, a->v, &out->alpha));
  }
  out->cin = cin; out->cout = cout; out->taps = taps;
  return kOk;
}

int pack_linear_host(svt_video* v, const std::string& p, int N, int K, LinearW* out) {
  const HostTensor *w, *b;
  SVT_TRY(need(v, p + "weight", {N, K}, &w));
  SVT_TRY(need(v, p + "bias", {N}, &b));
  std::vector<__nv_bfloat16> pw(w->v.size());
  for (size_t i = 0; i < pw.size(); ++i) pw[i] = __float2bfloat16(w->v[i]);
  SVT_TRY(upload(v->pool, pw, &out->w));
  SVT_TRY(upload(v->pool, b->v, &out->b));
  out->N = N; out->K = K;
  return kOk;
}

int finalize_video(svt_video* v) {
  const int D = v->cfg.embed_dim;
  v->pool.release();
  const std::string r = "feature_extractor_video.resnet.";
  // ---- Conv3d front end (64, 1, 5, 7, 7) + BN3d + PReLU -> [64][5 temporal taps][64] (k = dy*7 + dx, zero padded)
  {
    const HostTensor* w;
    SVT_TRY(need(v, r + "frontend3D.0.weight", {64, 1, 5, 7, 7}, &w));
    std::vector<float> scale, shift;
    SVT_TRY(bn_fold(v, r + "frontend3D.1.", 64, &scale, &shift));
    std::vector<__nv_bfloat16> pw(static_cast<size_t>(64) * kFrontT * kFrontK, __float2bfloat16(0.f));
    for (int co = 0; co < 64; ++co)
      for (int dt = 0; dt < kFrontT; ++dt)
        for (int k = 0; k < 49; ++k)
          pw[(static_cast<size_t>(co) * kFrontT + dt) * kFrontK + k] =
              __float2bfloat16(w->v[(static_cast<size_t>(co) * kFrontT + dt) * 49 + k] * scale[co]);
    SVT_TRY(upload(v->pool, pw, &v->front.w));
    SVT_TRY(upload(v->pool, shift, &v->front.bias));
    const HostTensor* a;
    SVT_TRY(need(v, r + "frontend3D.2.weight", {64}, &a));
    SVT_TRY(upload(v->pool, a->v, &v->front.alpha));
    v->front.cin = kFrontK; v->front.cout = 64; v->front.taps = kFrontT;
  }
  // ---- ResNet-18 trunk (resnet.py:79-131): BasicBlock = conv1-bn1-prelu1-conv2-bn2 (+ downsample(x)) - prelu2
  int inpl = 64;
  for (int li = 0; li < 4; ++li) {
    for (int bi = 0; bi < 2; ++bi) {
      const std::string p = r + "trunk.layer" + std::to_string(li + 1) + "." + std::to_string(bi) + ".";
      const int cin = bi == 0 ? inpl : kPlanes[li];
      svt_video::Block& blk = v->blocks[li][bi];
      SVT_TRY(pack_conv(v, p + "conv1.", p + "bn1.", p + "relu1.", cin, kPlanes[li], 3, &blk.c1));
      SVT_TRY(pack_conv(v, p + "conv2.", p + "bn2.", p + "relu2.", kPlanes[li], kPlanes[li], 3, &blk.c2));
      blk.has_ds = find(v, p + "downsample.0.weight") != nullptr;
      if (blk.has_ds) SVT_TRY(pack_conv(v, p + "downsample.0.", p + "downsample.1.", "", cin, kPlanes[li], 1, &blk.ds));
      if (bi == 0 && li > 0 && !blk.has_ds) return fail(kUnknownTensor, "missing downsample branch of " + p);
    }
    inpl = kPlanes[li];
  }
  SVT_TRY(pack_linear_host(v, "feature_extractor_video.proj.", D, 512, &v->proj));
  SVT_TRY(pack_linear_host(v, "post_extract_proj.", D, 2 * D, &v->post));
  {
    const HostTensor *g, *b;
    SVT_TRY(need(v, "layer_norm.weight", {2 * D}, &g));
    SVT_TRY(need(v, "layer_norm.bias", {2 * D}, &b));
    SVT_TRY(upload(v->pool, g->v, &v->cat_norm.g));
    SVT_TRY(upload(v->pool, b->v, &v->cat_norm.b));
  }
  SVT_TRY(encoder_finalize(v->enc));
  v->host.clear();
  v->finalized = true;
  return kOk;
}

struct VideoPlan {
  int B, T, Ta, N;
  size_t off_stats, off_col, off_f0, off_buf[4], off_mask[4], off_cat, off_h, off_hb, off_qkv, off_ctx, off_mid, off_pre, off_rowstats, total;
  size_t buf_elems;
};
VideoPlan make_plan(const svt_video* v, int B, int T) {
  VideoPlan p{};
  // two all-zero frame slots after every clip = the temporal zero padding the front end's frame-shift taps read
  p.B = B; p.T = T; p.Ta = (T + 2 + 3) / 4 * 4; p.N = B * p.Ta;
  const size_t N = p.N, D = v->cfg.embed_dim, F = v->cfg.ffn_size;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
  p.off_stats = take(64);
  p.off_col = take(N * kF0 * kF0 * kFrontK * 2 + 4096);
  p.off_f0 = take(N * kF0 * kF0 * 64 * 2);
  p.buf_elems = N * 24 * 24 * 128;  // largest map: the full-resolution evaluation of layer2's stride-2 conv
  for (int i = 0; i < 4; ++i) p.off_buf[i] = take(p.buf_elems * 2 + 4096);
  for (int i = 0; i < 4; ++i) p.off_mask[i] = take(N * (kRes[i] + 2) * (kRes[i] + 2));
  p.off_cat = take(N * 2 * D * 2);
  p.off_h = take(N * D * 4);
  p.off_hb = take(N * D * 2);
  p.off_qkv = take(N * 3 * D * 2);
  p.off_ctx = take(N * D * 2);
  p.off_mid = take(N * F * 2);
  p.off_pre = take(N * D * 4);
  p.off_rowstats = take(transformer_rowstats_bytes(v->enc, N));
  p.total = off;
  return p;
}

// out[rows, cout] = act( conv_{taps}(x)[rows] + bias (+ resid) ), ring rows zeroed; x / out / resid: flat padded maps
int conv_gemm(const ConvW& w, const __nv_bfloat16* x, long long rows, int Hp, const __nv_bfloat16* resid, const uint8_t* mask,
              bool prelu, __nv_bfloat16* out, cudaStream_t s) {
  GemmArgs g;
  g.a = x;
  g.a_dims[0] = w.cin; g.a_dims[1] = 1; g.a_dims[2] = static_cast<uint64_t>(rows);
  g.a_strides[0] = w.cin; g.a_strides[1] = w.cin;
  g.w = w.w; g.w_rows = w.cout; g.w_cols = w.taps * w.cin;
  g.M = static_cast<int>(rows); g.N = w.cout; g.K = w.taps * w.cin; g.k_inner = w.cin;
  g.bias = w.bias; g.out_bf16 = out; g.ld_out = w.cout;
  g.resid_bf16 = resid; g.row_mask = mask;
  if (prelu) { g.act = kActPRelu; g.alpha = w.alpha; }
  if (w.taps == 9) {
    g.n_taps = 9;
    for (int dy = 0; dy < 3; ++dy)
      for (int dx = 0; dx < 3; ++dx) g.tap_off[dy * 3 + dx] = (dy - 1) * Hp + (dx - 1);
  } else if (w.taps == kFrontT) {  // front end: tap = frame shift, Hp = rows per frame
    g.n_taps = kFrontT;
    for (int dt = 0; dt < kFrontT; ++dt) g.tap_off[dt] = (dt - kFrontT / 2) * Hp;
  }
  return gemm_bf16_tc(g, s);
}

int grid_for(long long n, int block) { return static_cast<int>((n + block - 1) / block); }

int forward_video(svt_video* v, const float* video, int B, int T, void* ws, size_t ws_bytes, float* feats, cudaStream_t s) {
  if (!v->finalized) return fail(kNotFinalized, "svt_video_finalize has not been called");
  if (B <= 0 || T <= 0) return fail(kInvalidArgument, "batch and frames must be positive");
  const VideoPlan p = make_plan(v, B, T);
  if (ws_bytes < p.total) return fail(kWorkspaceTooSmall, "workspace too small: need " + std::to_string(p.total));
  if (static_cast<long long>(p.N) * kF0 * kF0 > 2000000000LL) return fail(kUnsupported, "too many frames for one call");
  uint8_t* base = static_cast<uint8_t*>(ws);
  double* stats_in = reinterpret_cast<double*>(base + p.off_stats);
  double* stats_out = stats_in + 2;
  __nv_bfloat16* col = reinterpret_cast<__nv_bfloat16*>(base + p.off_col);
  __nv_bfloat16* f0 = reinterpret_cast<__nv_bfloat16*>(base + p.off_f0);
  __nv_bfloat16* buf[4];
  uint8_t* mask[4];
  for (int i = 0; i < 4; ++i) {
    buf[i] = reinterpret_cast<__nv_bfloat16*>(base + p.off_buf[i]);
    mask[i] = base + p.off_mask[i];
  }
  __nv_bfloat16* cat = reinterpret_cast<__nv_bfloat16*>(base + p.off_cat);
  const int N = p.N, D = v->cfg.embed_dim, Ta = p.Ta;

  // ---- optional whole-tensor input LN (fairseq_interface.py:473-474), folded into the gather
  if (v->cfg.input_norm) SVT_TRY(tensor_stats(video, static_cast<size_t>(B) * T * kImg * kImg, stats_in, s));
  // ---- Conv3d front end as patch matrix + GEMM (BN folded, PReLU epilogue)
  {
    const long long rows = static_cast<long long>(N) * kF0 * kF0;
    frontend_patch_kernel<<<N * kF0, 256, 0, s>>>(video, B, T, Ta, v->cfg.input_norm ? stats_in : nullptr,
                                                  1.0 / (static_cast<double>(B) * T * kImg * kImg), col);
    SVT_POST_LAUNCH();
    SVT_TRY(conv_gemm(v->front, col, rows, kF0 * kF0, nullptr, nullptr, true, f0, s));
    maxpool_kernel<<<grid_for(static_cast<long long>(N) * 24 * 24 * 8, 256), 256, 0, s>>>(f0, N, buf[0]);
    SVT_POST_LAUNCH();
  }
  for (int i = 0; i < 4; ++i) {
    const long long rows = static_cast<long long>(N) * (kRes[i] + 2) * (kRes[i] + 2);
    ring_mask_kernel<<<grid_for(rows, 256), 256, 0, s>>>(mask[i], rows, kRes[i] + 2);
    SVT_POST_LAUNCH();
  }
  // ---- ResNet-18 trunk; `cur` indexes the buffer holding the current block input
  int cur = 0;
  for (int li = 0; li < 4; ++li) {
    const int Hp = kRes[li] + 2;
    const long long rows = static_cast<long long>(N) * Hp * Hp;
    for (int bi = 0; bi < 2; ++bi) {
      const svt_video::Block& blk = v->blocks[li][bi];
      const int a = (cur + 1) & 3, b2 = (cur + 2) & 3, c3 = (cur + 3) & 3;
      if (bi == 0 && li > 0) {
        // stride-2 block: conv1 at the input resolution, then keep every other pixel; downsample branch on the
        // subsampled input; conv2 at the new resolution adds it
        const int Hs = kRes[li - 1], Hsp = Hs + 2;
        const long long rows_in = static_cast<long long>(N) * Hsp * Hsp;
        SVT_TRY(conv_gemm(blk.c1, buf[cur], rows_in, Hsp, nullptr, nullptr, true, buf[a], s));
        subsample_kernel<<<grid_for(rows * (blk.c1.cout / 8), 256), 256, 0, s>>>(buf[a], N, Hs, blk.c1.cout, buf[b2]);
        SVT_POST_LAUNCH();
        subsample_kernel<<<grid_for(rows * (blk.ds.cin / 8), 256), 256, 0, s>>>(buf[cur], N, Hs, blk.ds.cin, buf[a]);
        SVT_POST_LAUNCH();
        SVT_TRY(conv_gemm(blk.ds, buf[a], rows, Hp, nullptr, mask[li], false, buf[c3], s));
        SVT_TRY(conv_gemm(blk.c2, buf[b2], rows, Hp, buf[c3], mask[li], true, buf[cur], s));
        // output landed in buf[cur]
      } else {
        SVT_TRY(conv_gemm(blk.c1, buf[cur], rows, Hp, nullptr, mask[li], true, buf[a], s));
        SVT_TRY(conv_gemm(blk.c2, buf[a], rows, Hp, buf[cur], mask[li], true, buf[b2], s));
        cur = b2;
      }
    }
  }
  // ---- avgpool -> SubModel.proj into the video half of the concat buffer (audio half = 0) -> LN(2D) -> post_extract_proj
  __nv_bfloat16* pooled = buf[(cur + 1) & 3];
  avgpool_kernel<<<grid_for(static_cast<long long>(N) * 512, 256), 256, 0, s>>>(buf[cur], N, 512, pooled, 512);
  SVT_POST_LAUNCH();
  SVT_CUDA(cudaMemsetAsync(cat, 0, static_cast<size_t>(N) * 2 * D * 2, s));
  {
    GemmArgs g;
    g.a = pooled;
    g.a_dims[0] = 512; g.a_dims[1] = 1; g.a_dims[2] = N;
    g.a_strides[0] = 512; g.a_strides[1] = 512;
    g.w = v->proj.w; g.w_rows = D; g.w_cols = 512;
    g.M = N; g.N = D; g.K = 512; g.k_inner = 512;
    g.bias = v->proj.b; g.out_bf16 = cat + D; g.ld_out = 2 * D;
    SVT_TRY(gemm_bf16_tc(g, s));
  }
  {
    LayerNormArgs ln;
    ln.x_bf16 = cat; ln.y_bf16 = cat; ln.gamma = v->cat_norm.g; ln.beta = v->cat_norm.b;
    ln.rows = N; ln.D = 2 * D; ln.eps = v->cfg.layer_norm_eps;
    SVT_TRY(layer_norm(ln, s));
  }
  TransformerBuffers tb;
  tb.h = reinterpret_cast<float*>(base + p.off_h);
  tb.hb = reinterpret_cast<__nv_bfloat16*>(base + p.off_hb);
  tb.qkv = reinterpret_cast<__nv_bfloat16*>(base + p.off_qkv);
  tb.ctx = reinterpret_cast<__nv_bfloat16*>(base + p.off_ctx);
  tb.mid = reinterpret_cast<__nv_bfloat16*>(base + p.off_mid);
  tb.pre = reinterpret_cast<float*>(base + p.off_pre);
  tb.rowstats = reinterpret_cast<float*>(base + p.off_rowstats);
  {
    GemmArgs g;
    g.a = cat;
    g.a_dims[0] = 2 * D; g.a_dims[1] = 1; g.a_dims[2] = N;
    g.a_strides[0] = 2 * D; g.a_strides[1] = 2 * D;
    g.w = v->post.w; g.w_rows = D; g.w_cols = 2 * D;
    g.M = N; g.N = D; g.K = 2 * D; g.k_inner = 2 * D;
    g.bias = v->post.b; g.out_f32 = tb.h; g.out_bf16 = tb.hb; g.ld_out = D;
    SVT_TRY(gemm_bf16_tc(g, s));
  }
  // ---- transformer body (fairseq TransformerEncoder == csrc/encoder.cu graph) + whole-tensor output LN
  const bool want_stats = v->cfg.output_norm != 0;
  const float* final_x = nullptr;
  SVT_TRY(encoder_transformer_forward(v->enc, B, T, Ta, tb, want_stats ? stats_out : nullptr, 0, &final_x, s));
  HeadArgs ha;
  ha.x = final_x; ha.clips = B; ha.clip_rows = Ta; ha.T = T; ha.D = D;
  ha.stats = want_stats ? stats_out : nullptr; ha.eps = 1e-5f;
  ha.feats = feats;
  return head_forward(ha, s);
}

std::string fairseq_to_hf(std::string n) {
  auto rep = [&](const std::string& from, const std::string& to) {
    size_t pos = n.find(from);
    if (pos != std::string::npos) n.replace(pos, from.size(), to);
  };
  rep("encoder.pos_conv.0.", "encoder.pos_conv_embed.conv.");
  rep(".self_attn_layer_norm.", ".layer_norm.");
  rep(".self_attn.", ".attention.");
  rep(".fc1.", ".feed_forward.intermediate_dense.");
  rep(".fc2.", ".feed_forward.output_dense.");
  return n;
}

}  // namespace

extern "C" {

int svt_video_create(const svt_video_config* cfg, svt_video** out) {
  if (cfg == nullptr || out == nullptr) return fail(kInvalidArgument, "null argument");
  const svt_video_config& c = *cfg;
  if (c.embed_dim != 256 && c.embed_dim != 512 && c.embed_dim != 1024)
    return fail(kUnsupported, "embed_dim must be 256, 512 or 1024 (row kernels are built for D and 2D in {256, 512, 1024, 2048})");
  if (c.num_heads <= 0 || c.embed_dim % c.num_heads != 0) return fail(kInvalidArgument, "bad num_heads");
  const int dh = c.embed_dim / c.num_heads;
  if (dh != 64 && dh != 128) return fail(kUnsupported, "head dim must be 64 or 128");
  if (c.ffn_size % 64 != 0) return fail(kUnsupported, "ffn_size must be a multiple of 64");
  if (c.conv_pos_groups <= 0 || c.embed_dim % c.conv_pos_groups != 0) return fail(kInvalidArgument, "bad conv_pos_groups");
  const int dg = c.embed_dim / c.conv_pos_groups;
  if (dg > 64 || dg % 16 != 0) return fail(kUnsupported, "positional conv channels per group must be a multiple of 16, <= 64");
  svt_video* v = new svt_video();
  v->cfg = c;
  v->enc = new svt_encoder();
  v->enc->transformer_only = true;
  svt_encoder_config& e = v->enc->cfg;
  e.hidden_size = c.embed_dim; e.num_layers = c.num_layers; e.num_heads = c.num_heads; e.ffn_size = c.ffn_size;
  e.num_conv_layers = 0; e.conv_dim = 512;
  e.stable_layer_norm = 1;  // fairseq layer_norm_first = True (AV-HuBERT large / base checkpoints)
  e.pos_conv_kernel = c.conv_pos; e.pos_conv_groups = c.conv_pos_groups; e.layer_norm_eps = c.layer_norm_eps;
  e.normalize_wav = 0; e.output_norm = c.output_norm;
  *out = v;
  return kOk;
}

void svt_video_destroy(svt_video* v) { delete v; }

int svt_video_set_tensor(svt_video* v, const char* name, const float* host, const int64_t* shape, int ndim, int strict) {
  if (v == nullptr || name == nullptr || host == nullptr) return fail(kInvalidArgument, "null argument");
  if (svt_device_count() <= 0) return fail(kNoDevice, "no CUDA device");
  std::string n(name);
  if (n.rfind("model.", 0) == 0) n = n.substr(6);
  v->finalized = false;
  if (n.rfind("encoder.", 0) == 0) {
    v->enc->finalized = false;
    return v->enc->reg.set(fairseq_to_hf(n), host, shape, ndim);
  }
  const bool known = n.rfind("feature_extractor_video.", 0) == 0 || n.rfind("layer_norm.", 0) == 0 ||
                     n.rfind("post_extract_proj.", 0) == 0;
  if (!known) return strict ? fail(kUnknownTensor, "unknown tensor " + n) : static_cast<int>(kOk);
  HostTensor t;
  t.shape.assign(shape, shape + ndim);
  size_t cnt = 1;
  for (int i = 0; i < ndim; ++i) cnt *= static_cast<size_t>(shape[i]);
  t.v.assign(host, host + cnt);
  v->host[n] = std::move(t);
  return kOk;
}

int svt_video_finalize(svt_video* v) {
  if (v == nullptr) return fail(kInvalidArgument, "null argument");
  if (svt_device_count() <= 0) return fail(kNoDevice, "no CUDA device");
  return finalize_video(v);
}

size_t svt_video_workspace_bytes(const svt_video* v, int batch, int n_frames) {
  if (v == nullptr || batch <= 0 || n_frames <= 0) return 0;
  return make_plan(v, batch, n_frames).total;
}

int svt_video_transform_u8(const uint8_t* frames_dev, long long n_frames, int height, int width, int crop, float mean,
                           float stdev, float* out_dev, void* stream) {
  if (frames_dev == nullptr || out_dev == nullptr) return fail(kInvalidArgument, "null argument");
  if (n_frames <= 0 || crop <= 0 || crop > height || crop > width || stdev == 0.f)
    return fail(kInvalidArgument, "video transform: bad geometry");
  // CenterCrop of the reference: delta = int(round(w - tw) / 2.)  (utils.py:79-80)
  const int x0 = (width - crop) / 2, y0 = (height - crop) / 2;
  const long long total = n_frames * crop * crop;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(num_sms()) * 16));
  video_transform_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(frames_dev, n_frames, height, width, crop, y0, x0,
                                                                              mean, stdev, out_dev);
  SVT_POST_LAUNCH();
  return kOk;
}

int svt_video_forward(svt_video* v, const float* video_dev, int batch, int n_frames, void* workspace_dev, size_t workspace_bytes,
                      float* feats_dev, void* stream) {
  if (v == nullptr || video_dev == nullptr || workspace_dev == nullptr || feats_dev == nullptr)
    return fail(kInvalidArgument, "null argument");
  return forward_video(v, video_dev, batch, n_frames, workspace_dev, workspace_bytes, feats_dev, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
