// HBM-bound row kernels of the hot path: layer norms, the C_in=1 first conv layer (fused with the
// whole-tensor input normalisation, LayerNorm and GELU), whole-tensor output norm + linear head,
// frame post-processing.  All use 16-byte vectorised, warp-coalesced accesses and warp-shuffle reductions.
#include <cmath>
#include "ops.cuh"

namespace svt {

namespace {

__device__ __forceinline__ float4 ld_bf16x4(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st_bf16x4(__nv_bfloat16* p, float a, float b, float c, float d) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
}

__device__ __forceinline__ double block_sum_double(double v, double* sh) {
  // warp reduce then 8-warp smem reduce (blockDim.x == 256)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += sh[i];
  __syncthreads();
  return t;  // valid in thread 0
}

// ------------------------------------------------------------------------------------ LayerNorm
// one warp per row, NV float4 groups per lane (D = NV * 128)
template <int NV>
__global__ void __launch_bounds__(256)
layer_norm_kernel(const float* __restrict__ xf, const __nv_bfloat16* __restrict__ xb, const float* __restrict__ gamma,
                  const float* __restrict__ beta, __nv_bfloat16* __restrict__ yb, float* __restrict__ yf, int rows,
                  float eps, int gelu, double* __restrict__ stats, int clip_rows, int clip_valid, int stats_stride) {
  constexpr int D = NV * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  float s_out = 0.f, ss_out = 0.f;
  if (row < rows) {
    float4 v[NV];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const size_t off = static_cast<size_t>(row) * D + (i * 32 + lane) * 4;
      v[i] = (xf != nullptr) ? *reinterpret_cast<const float4*>(xf + off) : ld_bf16x4(xb + off);
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(sum) * (1.0f / D);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
      float4 y;
      y.x = v[i].x * rstd * g.x + b.x;
      y.y = v[i].y * rstd * g.y + b.y;
      y.z = v[i].z * rstd * g.z + b.z;
      y.w = v[i].w * rstd * g.w + b.w;
      if (gelu) { y.x = gelu_erf(y.x); y.y = gelu_erf(y.y); y.z = gelu_erf(y.z); y.w = gelu_erf(y.w); }
      const size_t off = static_cast<size_t>(row) * D + c;
      if (yb != nullptr) st_bf16x4(yb + off, y.x, y.y, y.z, y.w);
      if (yf != nullptr) *reinterpret_cast<float4*>(yf + off) = y;
      s_out += (y.x + y.y) + (y.z + y.w);
      ss_out += (y.x * y.x + y.y * y.y) + (y.z * y.z + y.w * y.w);
    }
    if (stats != nullptr && (row % clip_rows) >= clip_valid) { s_out = 0.f; ss_out = 0.f; }
  }
  if (stats != nullptr && stats_stride > 0) {
    // per-clip statistics: a block's 8 rows may straddle two clips, so every warp (= row) adds to its own clip's pair
    const float a = warp_sum(s_out), b = warp_sum(ss_out);
    if (lane == 0 && row < rows) {
      double* st = stats + static_cast<size_t>(row / clip_rows) * stats_stride;
      atomicAdd(st, static_cast<double>(a));
      atomicAdd(st + 1, static_cast<double>(b));
    }
  } else if (stats != nullptr) {
    __shared__ double sh[8];
    const double a = block_sum_double(static_cast<double>(s_out), sh);
    const double b = block_sum_double(static_cast<double>(ss_out), sh);
    if (threadIdx.x == 0) { atomicAdd(stats, a); atomicAdd(stats + 1, b); }
  }
}

// ------------------------------------------------------------------------------------ tensor stats
// blockIdx.y = segment (clip): x + y * seg_stride, stats + y * stats_stride (both 0 for the whole-tensor form).
// A segment may start at any 4-byte address (a row of a batch whose length is not a multiple of 4, a sliced view of a
// song): up to 3 leading scalars bring the body to a 16-byte boundary, the remainder is a scalar tail.
__global__ void __launch_bounds__(256) tensor_stats_kernel(const float* __restrict__ x, size_t n, double* stats,
                                                           size_t seg_stride, int stats_stride) {
  x += static_cast<size_t>(blockIdx.y) * seg_stride;
  stats += static_cast<size_t>(blockIdx.y) * stats_stride;
  size_t head = ((16 - (reinterpret_cast<uintptr_t>(x) & 15)) & 15) >> 2;
  if (head > n) head = n;
  const float* xa = x + head;
  const size_t nb = n - head;
  float s = 0.f, ss = 0.f;
  const size_t n4 = nb / 4;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(xa) + i);
    s += (v.x + v.y) + (v.z + v.w);
    ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  if (blockIdx.x == 0) {
    if (threadIdx.x < head) {
      const float v = x[threadIdx.x];
      s += v; ss += v * v;
    } else if (threadIdx.x >= 4 && threadIdx.x - 4 < (nb & 3)) {
      const float v = xa[n4 * 4 + threadIdx.x - 4];
      s += v; ss += v * v;
    }
  }
  __shared__ double sh[8];
  const double a = block_sum_double(static_cast<double>(s), sh);
  const double b = block_sum_double(static_cast<double>(ss), sh);
  if (threadIdx.x == 0) { atomicAdd(stats, a); atomicAdd(stats + 1, b); }
}

__device__ __forceinline__ void mean_rstd_from_stats(const double* stats, double n, float eps, float& mean, float& rstd) {
  const double m = stats[0] / n;
  double var = stats[1] / n - m * m;
  if (var < 0.0) var = 0.0;
  mean = static_cast<float>(m);
  rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

// ------------------------------------------------------------------------------------ conv layer 0
// C = 512, k = 10, stride = 5.  One warp computes 4 consecutive frames; lane owns channels
// {i*128 + lane*4 + e : i<4, e<4}.  Weights [k][C] fp32 live in shared memory.
constexpr int kC0 = 512, kK0 = 10, kS0 = 5, kR0 = 4;
constexpr int kConv0GroupsPerWarp = 5;

template <bool kLayerMode>
__global__ void __launch_bounds__(256)
conv0_kernel(const float* __restrict__ wav, int L, int T, int t_alloc, const float* __restrict__ w,
             const float* __restrict__ bias, const float* __restrict__ gamma, const float* __restrict__ beta,
             const double* __restrict__ in_stats, double n_in, __nv_bfloat16* __restrict__ out,
             double* __restrict__ chan_stats, int stats_stride) {
  __shared__ __align__(16) float sw[kK0 * kC0];
  __shared__ __align__(16) float sbias[kC0];
  __shared__ __align__(16) float sgamma[kC0];
  __shared__ __align__(16) float sbeta[kC0];
  for (int i = threadIdx.x; i < kK0 * kC0; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < kC0; i += blockDim.x) {
    sbias[i] = bias != nullptr ? bias[i] : 0.f;
    sgamma[i] = kLayerMode ? gamma[i] : 1.f;
    sbeta[i] = kLayerMode ? beta[i] : 0.f;
  }
  __syncthreads();
  float mean = 0.f, rstd = 1.f;
  if (in_stats != nullptr) mean_rstd_from_stats(in_stats + static_cast<size_t>(blockIdx.y) * stats_stride, n_in, 1e-5f, mean, rstd);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int clip = blockIdx.y;
  const float* x = wav + static_cast<size_t>(clip) * L;
  __nv_bfloat16* o = out + static_cast<size_t>(clip) * t_alloc * kC0;
  float cs[16], css[16];  // group mode: per-channel sums over this warp's frames
#pragma unroll
  for (int i = 0; i < 16; ++i) cs[i] = css[i] = 0.f;

  for (int gi = 0; gi < kConv0GroupsPerWarp; ++gi) {
    const int t0 = ((blockIdx.x * 8 + warp) * kConv0GroupsPerWarp + gi) * kR0;
    if (t0 >= t_alloc) break;
    // 25 input samples cover 4 frames; lane l holds sample 5*t0 + l
    const int si = kS0 * t0 + lane;
    float xv = (lane < kS0 * (kR0 - 1) + kK0 && si < L) ? (__ldg(x + si) - mean) * rstd : 0.f;
    float acc[kR0][16];
#pragma unroll
    for (int r = 0; r < kR0; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[r][i] = 0.f;
#pragma unroll
    for (int j = 0; j < kK0; ++j) {
      float4 wv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) wv[i] = *reinterpret_cast<const float4*>(&sw[j * kC0 + i * 128 + lane * 4]);
#pragma unroll
      for (int r = 0; r < kR0; ++r) {
        const float xs = __shfl_sync(0xffffffffu, xv, kS0 * r + j);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[r][4 * i + 0] = fmaf(xs, wv[i].x, acc[r][4 * i + 0]);
          acc[r][4 * i + 1] = fmaf(xs, wv[i].y, acc[r][4 * i + 1]);
          acc[r][4 * i + 2] = fmaf(xs, wv[i].z, acc[r][4 * i + 2]);
          acc[r][4 * i + 3] = fmaf(xs, wv[i].w, acc[r][4 * i + 3]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kR0; ++r) {
      const int t = t0 + r;
      __nv_bfloat16* orow = o + static_cast<size_t>(t) * kC0;
      if (t >= T) {  // allocation padding rows: keep finite
#pragma unroll
        for (int i = 0; i < 4; ++i) st_bf16x4(orow + i * 128 + lane * 4, 0.f, 0.f, 0.f, 0.f);
        continue;
      }
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 b = *reinterpret_cast<const float4*>(&sbias[i * 128 + lane * 4]);
        acc[r][4 * i + 0] += b.x; acc[r][4 * i + 1] += b.y; acc[r][4 * i + 2] += b.z; acc[r][4 * i + 3] += b.w;
        sum += (acc[r][4 * i] + acc[r][4 * i + 1]) + (acc[r][4 * i + 2] + acc[r][4 * i + 3]);
      }
      if (kLayerMode) {
        const float mu = warp_sum(sum) * (1.0f / kC0);
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) { acc[r][i] -= mu; sq = fmaf(acc[r][i], acc[r][i], sq); }
        const float rs = rsqrtf(warp_sum(sq) * (1.0f / kC0) + 1e-5f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 g = *reinterpret_cast<const float4*>(&sgamma[i * 128 + lane * 4]);
          const float4 b = *reinterpret_cast<const float4*>(&sbeta[i * 128 + lane * 4]);
          st_bf16x4(orow + i * 128 + lane * 4, gelu_erf(acc[r][4 * i + 0] * rs * g.x + b.x),
                    gelu_erf(acc[r][4 * i + 1] * rs * g.y + b.y), gelu_erf(acc[r][4 * i + 2] * rs * g.z + b.z),
                    gelu_erf(acc[r][4 * i + 3] * rs * g.w + b.w));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) { cs[i] += acc[r][i]; css[i] = fmaf(acc[r][i], acc[r][i], css[i]); }
#pragma unroll
        for (int i = 0; i < 4; ++i)
          st_bf16x4(orow + i * 128 + lane * 4, acc[r][4 * i], acc[r][4 * i + 1], acc[r][4 * i + 2], acc[r][4 * i + 3]);
      }
    }
  }
  if (!kLayerMode) {
    double* st = chan_stats + static_cast<size_t>(clip) * kC0 * 2;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = i * 128 + lane * 4 + e;
        atomicAdd(st + 2 * c, static_cast<double>(cs[4 * i + e]));
        atomicAdd(st + 2 * c + 1, static_cast<double>(css[4 * i + e]));
      }
  }
}

// ------------------------------------------------------------------------------------ conv layer 0 on the tensor cores
// Layer-norm models (wav2vec2-large ...): (x - mean) * rstd -> conv(k = 10, s = 5) + bias -> LN(512) -> GELU -> bf16.
// The SIMT kernel above spends 10 FFMA + shuffles per output on the dot products and two warp reductions per frame on the
// LayerNorm statistics; it is bound by instruction issue (76 % of the issue slots, 0.23 of the HBM roof).  Here
//   * the statistics of a frame come from closed forms of its 10 input samples:  sum_c y_c = W1 . x + B1  and
//     sum_c y_c^2 = x^T G x + 2 H . x + B2  with G = W^T W (10 x 10), H = W^T b -- ~80 FFMA per FRAME, one lane per frame;
//   * the dot products run as mma.sync.m16n8k16 (bf16 in, fp32 accumulate) with fp32 precision kept by splitting both
//     operands into bf16 hi + lo parts (x_hi w_hi + x_lo w_hi + x_hi w_lo: relative error ~2^-17);
//   * the K = 16 operand has six spare slots: slot 10 carries rstd_f (A) x bias_c (B) and slot 11 carries -mean_f rstd_f
//     (A) x 1 (B), and the taps of A are pre-scaled by rstd_f, so the accumulator IS the normalised value
//     z = (conv + bias - mean_f) * rstd_f and the epilogue is one FFMA (gamma, beta) + GELU per element;
//   * the columns of the B operand are permuted so that four consecutive n-tiles leave 8 CONTIGUOUS channels in a thread:
//     one 16-byte store per row and thread, a quad covers 64 bytes.
// A warp handles 32 frames (two m16 tiles) at a time so every B fragment / (gamma, beta) vector read from shared memory
// feeds two MMAs rows.  Tables (conv0_build_tables): B fragments [64 n-tiles][32 lanes] uint4 {hi k0-1, hi k8-9, lo k0-1,
// lo k8-9} + the 77 closed-form coefficients.
constexpr int kC0Tiles = kC0 / 8;                      // 64 n-tiles
constexpr int kC0FragBytes = kC0Tiles * 32 * 16;       // 32 KB
constexpr int kC0QuadFloats = 80;                      // G' 55 (upper triangle, off-diagonals doubled), 2H 10, W1 10, B1, B2
constexpr int kC0FramesPerWarp = 32;

__device__ __forceinline__ int conv0_tile_channel(int tile, int col) {  // column `col` (0..7) of n-tile `tile` -> channel
  return 32 * (tile >> 2) + 8 * (col >> 1) + 2 * (tile & 3) + (col & 1);
}
__device__ __forceinline__ void split_bf16(float v, float& hi, float& lo) {
  hi = __bfloat162float(__float2bfloat16_rn(v));
  lo = v - hi;
}

__global__ void __launch_bounds__(256) conv0_tables_kernel(const float* __restrict__ w /*[k][C]*/, const float* __restrict__ bias,
                                                           uint4* __restrict__ frag, float* __restrict__ quad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kC0Tiles * 32) {
    const int tile = i >> 5, lane = i & 31, g = lane >> 2, tig = lane & 3;
    const int c = conv0_tile_channel(tile, g);
    auto val = [&](int k) -> float {
      if (k < kK0) return w[k * kC0 + c];
      if (k == 10) return bias != nullptr ? bias[c] : 0.f;
      return k == 11 ? 1.f : 0.f;
    };
    float hi[4], lo[4];
    const int ks[4] = {2 * tig, 2 * tig + 1, 2 * tig + 8, 2 * tig + 9};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      split_bf16(val(ks[j]), hi[j], lo[j]);
      if (ks[j] == 11) lo[j] = 0.f;
    }
    frag[i] = make_uint4(pack_bf16x2(hi[0], hi[1]), pack_bf16x2(hi[2], hi[3]), pack_bf16x2(lo[0], lo[1]), pack_bf16x2(lo[2], lo[3]));
  }
  if (blockIdx.x == 0 && threadIdx.x < 77) {
    // closed-form coefficients in double: entry e < 55 is G'(j, k), j <= k
    const int e = threadIdx.x;
    double acc = 0.0;
    if (e < 55) {
      int j = 0, rem = e;
      while (rem >= kK0 - j) { rem -= kK0 - j; ++j; }
      const int k = j + rem;
      for (int c = 0; c < kC0; ++c) acc += static_cast<double>(w[j * kC0 + c]) * w[k * kC0 + c];
      if (k != j) acc *= 2.0;
    } else if (e < 65) {
      for (int c = 0; c < kC0; ++c) acc += 2.0 * w[(e - 55) * kC0 + c] * (bias != nullptr ? bias[c] : 0.f);
    } else if (e < 75) {
      for (int c = 0; c < kC0; ++c) acc += w[(e - 65) * kC0 + c];
    } else if (e == 75) {
      for (int c = 0; c < kC0; ++c) acc += bias != nullptr ? bias[c] : 0.f;
    } else {
      for (int c = 0; c < kC0; ++c) acc += bias != nullptr ? static_cast<double>(bias[c]) * bias[c] : 0.0;
    }
    quad[e] = static_cast<float>(acc);
  }
}

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// first MMA of an accumulator: C = 0 comes from the zero register instead of 4 MOVs per fragment
__device__ __forceinline__ void mma_bf16_16816_zero(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%10, %10, %10, %10};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}

__global__ void __launch_bounds__(256, 2)
conv0_tc_kernel(const float* __restrict__ wav, int L, int T, int t_alloc, const uint4* __restrict__ frag,
                const float* __restrict__ quad, const float* __restrict__ gamma, const float* __restrict__ beta,
                const double* __restrict__ in_stats, double n_in, __nv_bfloat16* __restrict__ out, int stats_stride, int iters) {
  __shared__ __align__(16) uint4 sfrag[kC0Tiles * 32];
  __shared__ __align__(16) float sgamma[kC0];
  __shared__ __align__(16) float sbeta[kC0];
  __shared__ float squad[kC0QuadFloats];
  for (int i = threadIdx.x; i < kC0Tiles * 32; i += blockDim.x) sfrag[i] = frag[i];
  for (int i = threadIdx.x; i < kC0; i += blockDim.x) { sgamma[i] = gamma[i]; sbeta[i] = beta[i]; }
  if (threadIdx.x < 77) squad[threadIdx.x] = quad[threadIdx.x];
  __syncthreads();
  float mean = 0.f, rstd = 1.f;
  if (in_stats != nullptr) mean_rstd_from_stats(in_stats + static_cast<size_t>(blockIdx.y) * stats_stride, n_in, 1e-5f, mean, rstd);
  const float in_shift = -mean * rstd;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  const float* x = wav + static_cast<size_t>(blockIdx.y) * L;
  __nv_bfloat16* o = out + static_cast<size_t>(blockIdx.y) * t_alloc * kC0;

  for (int it = 0; it < iters; ++it) {
    const int f0 = ((blockIdx.x * iters + it) * 8 + warp) * kC0FramesPerWarp;
    if (f0 >= t_alloc) break;
    // ---- LayerNorm statistics of frame f0 + lane from its 10 normalised samples (closed forms)
    float mu = 0.f, rs = 0.f;
    {
      const int f = f0 + lane;
      float xv[kK0];
#pragma unroll
      for (int j = 0; j < kK0; ++j) xv[j] = (f < T) ? fmaf(__ldg(x + kS0 * f + j), rstd, in_shift) : 0.f;
      float s1 = squad[75], s2 = squad[76];
      int e = 0;
#pragma unroll
      for (int j = 0; j < kK0; ++j) {
        s1 = fmaf(squad[65 + j], xv[j], s1);
        float row = squad[55 + j];
#pragma unroll
        for (int k = j; k < kK0; ++k) row = fmaf(squad[e++], xv[k], row);
        s2 = fmaf(row, xv[j], s2);
      }
      mu = s1 * (1.0f / kC0);
      rs = rsqrtf(fmaxf(s2 * (1.0f / kC0) - mu * mu, 0.f) + 1e-5f);
    }
    // ---- A fragments (hi / lo) of the two 16-frame tiles: taps scaled by the frame's rstd, slot 10 = rstd, slot 11 = -mu rstd
    uint32_t ahi[2][4], alo[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int fl = 16 * mt + 8 * h + g;            // frame of fragment rows g (h = 0) / g + 8 (h = 1), local index
        const int f = f0 + fl;
        const float rs_f = __shfl_sync(0xffffffffu, rs, fl);
        const float mu_f = __shfl_sync(0xffffffffu, mu, fl);
        const bool live = f < T;
        const float* xf = x + kS0 * f;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;     // k = 2 tig, 2 tig + 1, 2 tig + 8, 2 tig + 9
        if (live) {
          v0 = fmaf(__ldg(xf + 2 * tig), rstd, in_shift) * rs_f;
          v1 = fmaf(__ldg(xf + 2 * tig + 1), rstd, in_shift) * rs_f;
          if (tig == 0) {
            v2 = fmaf(__ldg(xf + 8), rstd, in_shift) * rs_f;
            v3 = fmaf(__ldg(xf + 9), rstd, in_shift) * rs_f;
          } else if (tig == 1) {
            v2 = rs_f;
            v3 = -mu_f * rs_f;
          }
        }
        float h0, l0, h1, l1, h2, l2, h3, l3;
        split_bf16(v0, h0, l0); split_bf16(v1, h1, l1); split_bf16(v2, h2, l2); split_bf16(v3, h3, l3);
        ahi[mt][h] = pack_bf16x2(h0, h1); ahi[mt][2 + h] = pack_bf16x2(h2, h3);
        alo[mt][h] = pack_bf16x2(l0, l1); alo[mt][2 + h] = pack_bf16x2(l2, l3);
      }
    // ---- 16 groups of four n-tiles = 32 channels; a thread ends up with channels 32 u + 8 tig .. + 7 of its four rows
#pragma unroll 1
    for (int u = 0; u < kC0Tiles / 4; ++u) {
      float acc[2][4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 bf = sfrag[(4 * u + i) * 32 + lane];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma_bf16_16816_zero(acc[mt][i], ahi[mt], bf.x, bf.y);
          mma_bf16_16816(acc[mt][i], alo[mt], bf.x, bf.y);
          mma_bf16_16816(acc[mt][i], ahi[mt], bf.z, bf.w);
        }
      }
      const int c0 = 32 * u + 8 * tig;
      const float4 g0 = *reinterpret_cast<const float4*>(&sgamma[c0]), g1 = *reinterpret_cast<const float4*>(&sgamma[c0 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&sbeta[c0]), b1 = *reinterpret_cast<const float4*>(&sbeta[c0 + 4]);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int f = f0 + 16 * mt + 8 * h + g;
          if (f >= t_alloc) continue;
          uint4 pk = make_uint4(0u, 0u, 0u, 0u);   // allocation padding rows (f >= T) stay zero
          if (f < T) {
            // accumulator registers 2 h, 2 h + 1 of n-tile i are channels c0 + 2 i, c0 + 2 i + 1
            const float y0 = gelu_erf(fmaf(acc[mt][0][2 * h], g0.x, b0.x)), y1 = gelu_erf(fmaf(acc[mt][0][2 * h + 1], g0.y, b0.y));
            const float y2 = gelu_erf(fmaf(acc[mt][1][2 * h], g0.z, b0.z)), y3 = gelu_erf(fmaf(acc[mt][1][2 * h + 1], g0.w, b0.w));
            const float y4 = gelu_erf(fmaf(acc[mt][2][2 * h], g1.x, b1.x)), y5 = gelu_erf(fmaf(acc[mt][2][2 * h + 1], g1.y, b1.y));
            const float y6 = gelu_erf(fmaf(acc[mt][3][2 * h], g1.z, b1.z)), y7 = gelu_erf(fmaf(acc[mt][3][2 * h + 1], g1.w, b1.w));
            pk = make_uint4(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3), pack_bf16x2(y4, y5), pack_bf16x2(y6, y7));
          }
          *reinterpret_cast<uint4*>(o + static_cast<size_t>(f) * kC0 + c0) = pk;
        }
    }
  }
}

// group-norm apply + GELU, in place.  block = 256 threads = 4 frames x 64 channel-octets
__global__ void __launch_bounds__(256)
groupnorm_gelu_kernel(__nv_bfloat16* __restrict__ x, const double* __restrict__ chan_stats,
                      const float* __restrict__ gamma, const float* __restrict__ beta, int T, int t_alloc, int C,
                      int frames_per_block) {
  const int clip = blockIdx.y;
  const int c0 = (threadIdx.x & 63) * 8;
  const int sub = threadIdx.x >> 6;
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const double* st = chan_stats + (static_cast<size_t>(clip) * C + c0 + e) * 2;
    const double m = st[0] / T;
    const double var = fmax(st[1] / T - m * m, 0.0);
    const float rs = static_cast<float>(1.0 / sqrt(var + 1e-5));
    sc[e] = rs * gamma[c0 + e];
    sh[e] = beta[c0 + e] - static_cast<float>(m) * sc[e];
  }
  const int tbeg = blockIdx.x * frames_per_block;
  for (int t = tbeg + sub; t < tbeg + frames_per_block && t < T; t += 4) {
    __nv_bfloat16* p = x + (static_cast<size_t>(clip) * t_alloc + t) * C + c0;
    const float4 a = ld_bf16x4(p), b = ld_bf16x4(p + 4);
    st_bf16x4(p, gelu_erf(a.x * sc[0] + sh[0]), gelu_erf(a.y * sc[1] + sh[1]), gelu_erf(a.z * sc[2] + sh[2]),
              gelu_erf(a.w * sc[3] + sh[3]));
    st_bf16x4(p + 4, gelu_erf(b.x * sc[4] + sh[4]), gelu_erf(b.y * sc[5] + sh[5]), gelu_erf(b.z * sc[6] + sh[6]),
              gelu_erf(b.w * sc[7] + sh[7]));
  }
}

// ------------------------------------------------------------------------------------ output norm + head
constexpr int head_rows(int nv) { return nv > 8 ? 2 : 4; }  // frames per warp and pass (register budget)
template <int NV>
__global__ void __launch_bounds__(256)
head_kernel(const float* __restrict__ x, int clips, int clip_rows, int T, const double* __restrict__ stats, float eps,
            const float* __restrict__ w, const float* __restrict__ b, int n_out, float* __restrict__ feats,
            float* __restrict__ logits, int w_in_smem, int stats_stride) {
  constexpr int D = NV * 128;
  constexpr int kHeadRows = head_rows(NV);
  extern __shared__ __align__(16) float sw[];
  if (w != nullptr && w_in_smem)
    for (int i = threadIdx.x; i < n_out * D; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  float mean = 0.f, rstd = 1.f;
  if (stats != nullptr && stats_stride == 0) mean_rstd_from_stats(stats, static_cast<double>(clips) * T * D, eps, mean, rstd);
  const float* wp = w_in_smem ? sw : w;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = clips * T;
  // kHeadRows frames per warp and pass: every weight vector read from shared memory feeds kHeadRows dot products (the
  // kernel is bound by those reads, not by HBM)
  int cached_clip = -1;
  for (int fr0 = (blockIdx.x * 8 + warp) * kHeadRows; fr0 < total; fr0 += gridDim.x * 8 * kHeadRows) {
    float4 v[kHeadRows][NV];
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r) {
      const int fr = fr0 + r;
      if (fr < total) {
        const int clip = fr / T, t = fr % T;
        if (stats != nullptr && stats_stride > 0 && clip != cached_clip) {
          mean_rstd_from_stats(stats + static_cast<size_t>(clip) * stats_stride, static_cast<double>(T) * D, eps, mean, rstd);
          cached_clip = clip;
        }
        const float* xr = x + (static_cast<size_t>(clip) * clip_rows + t) * D;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          float4 a = *reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4);
          a.x = (a.x - mean) * rstd; a.y = (a.y - mean) * rstd; a.z = (a.z - mean) * rstd; a.w = (a.w - mean) * rstd;
          if (feats != nullptr) *reinterpret_cast<float4*>(feats + static_cast<size_t>(fr) * D + (i * 32 + lane) * 4) = a;
          v[r][i] = a;
        }
      } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[r][i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    if (logits != nullptr) {
      float mine[kHeadRows];
#pragma unroll
      for (int r = 0; r < kHeadRows; ++r) mine[r] = 0.f;
      for (int n = 0; n < n_out; ++n) {
        float acc[kHeadRows];
#pragma unroll
        for (int r = 0; r < kHeadRows; ++r) acc[r] = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const float4 ww = *reinterpret_cast<const float4*>(wp + static_cast<size_t>(n) * D + (i * 32 + lane) * 4);
#pragma unroll
          for (int r = 0; r < kHeadRows; ++r) {
            acc[r] = fmaf(v[r][i].x, ww.x, acc[r]); acc[r] = fmaf(v[r][i].y, ww.y, acc[r]);
            acc[r] = fmaf(v[r][i].z, ww.z, acc[r]); acc[r] = fmaf(v[r][i].w, ww.w, acc[r]);
          }
        }
        const float bn = b != nullptr ? b[n] : 0.f;
#pragma unroll
        for (int r = 0; r < kHeadRows; ++r) {
          const float tot = warp_sum(acc[r]);
          if (lane == n) mine[r] = tot + bn;
        }
      }
#pragma unroll
      for (int r = 0; r < kHeadRows; ++r)
        if (lane < n_out && fr0 + r < total) logits[static_cast<size_t>(fr0 + r) * n_out + lane] = mine[r];
    }
  }
}

__global__ void frame_argmax_kernel(const float* __restrict__ logits, int n_frames, int n_out, int oct_off, int n_oct,
                                    int pc_off, int n_pc, int32_t* __restrict__ oct, int32_t* __restrict__ pc) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_frames) return;
  const float* r = logits + static_cast<size_t>(f) * n_out;
  int bo = 0, bp = 0;
  float vo = r[oct_off], vp = r[pc_off];
  for (int i = 1; i < n_oct; ++i) if (r[oct_off + i] > vo) { vo = r[oct_off + i]; bo = i; }  // first max wins
  for (int i = 1; i < n_pc; ++i) if (r[pc_off + i] > vp) { vp = r[pc_off + i]; bp = i; }
  oct[f] = bo;
  pc[f] = bp;
}

// ------------------------------------------------------------------------------------ misc
__global__ void add_pe_kernel(const float* __restrict__ x, int T_src, int T, int D, float* __restrict__ of,
                              __nv_bfloat16* __restrict__ ob) {
  // x: [clips, T_src, D]; out: [clips, T, D]; rows t >= T_src are zero before adding the encoding (fusion.py:198-203)
  const int clip = blockIdx.y, t = blockIdx.x;
  const float kNegLog = -(logf(10000.0f) / static_cast<float>(D));
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float denom = expf(static_cast<float>(c & ~1) * kNegLog);
    const float ang = static_cast<float>(t) * denom;
    const float pe = (c & 1) ? cosf(ang) : sinf(ang);
    const float v = (t < T_src ? x[(static_cast<size_t>(clip) * T_src + t) * D + c] : 0.f) + pe;
    const size_t o = (static_cast<size_t>(clip) * T + t) * D + c;
    if (of != nullptr) of[o] = v;
    if (ob != nullptr) ob[o] = __float2bfloat16(v);
  }
}

__global__ void add_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, size_t n) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) o[i] = a[i] + b[i];
}
__global__ void cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, size_t n) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = __float2bfloat16(x[i]);
}

// y[row][c] = bf16(x[row][c] * scale[c] + shift[c])  (eval-mode BatchNorm1d over channel-last rows)
__global__ void channel_affine_bf16_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                           const float* __restrict__ shift, __nv_bfloat16* __restrict__ y, size_t n4, int D4) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const int c = static_cast<int>(i % D4) * 4;
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + c));
    const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + c));
    st_bf16x4(y + i * 4, fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
  }
}

template <typename OutT>
__global__ void pack_kernel(const float* __restrict__ src, int d0, int d1, int d2, int d3, long long s0, long long s1,
                            long long s2, long long s3, float scale, const float* __restrict__ vec, int vec_dim,
                            OutT* __restrict__ dst) {
  const size_t n = static_cast<size_t>(d0) * d1 * d2 * d3;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    size_t r = i;
    const int i3 = r % d3; r /= d3;
    const int i2 = r % d2; r /= d2;
    const int i1 = r % d1; r /= d1;
    const int i0 = static_cast<int>(r);
    float v = src[i0 * s0 + i1 * s1 + i2 * s2 + i3 * s3] * scale;
    if (vec != nullptr) v *= vec[vec_dim == 0 ? i0 : (vec_dim == 1 ? i1 : (vec_dim == 2 ? i2 : i3))];
    if constexpr (sizeof(OutT) == 2) dst[i] = __float2bfloat16(v); else dst[i] = v;
  }
}

__global__ void weight_norm_scale_kernel(const float* __restrict__ v, const float* __restrict__ g, int n01, int taps,
                                         float* __restrict__ out) {
  const int j = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < n01; i += blockDim.x) {
    const double x = v[static_cast<size_t>(i) * taps + j];
    s += x * x;
  }
  __shared__ double sh[8];
  const double tot = block_sum_double(s, sh);
  if (threadIdx.x == 0) out[j] = static_cast<float>(static_cast<double>(g[j]) / sqrt(tot));
}

// one warp per row: y = bf16(x) and (sum, sum of squares) of every 128-column slice of the fp32 row -- the first link
// of the folded-LayerNorm chain in the layout of GemmArgs::row_stats_out (later links come out of GEMM epilogues)
__global__ void row_stats_cast_kernel(const float* __restrict__ x, int rows, int D, __nv_bfloat16* __restrict__ y,
                                      float* __restrict__ stats) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D);
  uint2* yr = reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * D);
  const int slots = D / 128;
  for (int j = 0; j < slots; ++j) {  // one pass of the warp = one 128-column slot
    const float4 v = xr[j * 32 + lane];
    float s = (v.x + v.y) + (v.z + v.w);
    float ss = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
    yr[j * 32 + lane] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    s = warp_sum(s);
    ss = warp_sum(ss);
    if (lane == 0) *reinterpret_cast<float2*>(stats + 2 * (static_cast<size_t>(row) * slots + j)) = make_float2(s, ss);
  }
}

// one warp per output feature n: colsum[n] = sum_k float(w_packed[n][k]) (what the tensor cores multiply the row mean
// by), bias[n] += scale * sum_k w_f32[n][k] * beta[k]
__global__ void ln_fold_vectors_kernel(const float* __restrict__ w_f32, const __nv_bfloat16* __restrict__ w_packed,
                                       const float* __restrict__ beta, float scale, int N, int K,
                                       float* __restrict__ colsum, float* __restrict__ bias) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  float cs = 0.f, d = 0.f;
  for (int k = lane; k < K; k += 32) {
    cs += __bfloat162float(w_packed[static_cast<size_t>(n) * K + k]);
    d = fmaf(w_f32[static_cast<size_t>(n) * K + k], beta[k], d);
  }
  cs = warp_sum(cs);
  d = warp_sum(d);
  if (lane == 0) {
    colsum[n] = cs;
    bias[n] += scale * d;
  }
}

// WavLM gate: one warp per row, lane pair (2h, 2h + 1) owns head h (64 channels)
__global__ void __launch_bounds__(256) wavlm_gate_kernel(const __nv_bfloat16* __restrict__ x, int rows, int heads,
                                                         const float* __restrict__ w2, const float* __restrict__ b2,
                                                         const float* __restrict__ head_const, float* __restrict__ gate) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int D = heads * 64;
  for (int base = 0; base < D; base += 1024) {  // 32 lanes x 32 channels per pass
    const int c0 = base + lane * 32;
    float a = 0.f, b = 0.f;
    if (c0 < D) {
      const __nv_bfloat16* xr = x + static_cast<size_t>(row) * D + c0;
      const int w0 = (lane & 1) * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 v = ld_bf16x4(xr + i);
        const float4 wa = __ldg(reinterpret_cast<const float4*>(w2 + w0 + i));
        const float4 wb = __ldg(reinterpret_cast<const float4*>(w2 + 64 + w0 + i));
        a += v.x * wa.x + v.y * wa.y + v.z * wa.z + v.w * wa.w;
        b += v.x * wb.x + v.y * wb.y + v.z * wb.z + v.w * wb.w;
      }
    }
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    b += __shfl_xor_sync(0xffffffffu, b, 1);
    if (c0 < D && (lane & 1) == 0) {
      const int h = c0 >> 6;
      const float ga = 1.f / (1.f + __expf(-(a + b2[0])));
      const float gb = 1.f / (1.f + __expf(-(b + b2[1])));
      gate[static_cast<size_t>(row) * heads + h] = ga * (gb * head_const[h] - 1.f) + 2.f;
    }
  }
}

// WavLM gate with the layer's first LayerNorm folded in (option "ln_fold"): x holds the UN-normalised bf16 rows, stats
// their [rows][D / 128][2] partial (sum, sum of squares); w2g = w2 o gamma per head [heads][2][64], cg = its row sums
// [heads][2], dg = w2 . beta + b2 [heads][2]:  w2 . LN(x)_head + b2 = rstd * (w2g . x_head - mean * cg) + dg
__global__ void __launch_bounds__(256) wavlm_gate_ln_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ stats,
                                                            int rows, int heads, const float* __restrict__ w2g,
                                                            const float* __restrict__ cg, const float* __restrict__ dg,
                                                            const float* __restrict__ head_const, float eps,
                                                            float* __restrict__ gate) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int D = heads * 64, slots = D / 128;
  float sum = 0.f, sumsq = 0.f;  // fixed order over the slots, as the GEMM epilogue does
  for (int j = 0; j < slots; ++j) {
    const float2 st = *reinterpret_cast<const float2*>(stats + 2 * (static_cast<size_t>(row) * slots + j));
    sum += st.x; sumsq += st.y;
  }
  const float mean = sum / D;
  const float rstd = rsqrtf(fmaxf(sumsq / D - mean * mean, 0.f) + eps);
  for (int base = 0; base < D; base += 1024) {
    const int c0 = base + lane * 32;
    float a = 0.f, b = 0.f;
    const int h = c0 >> 6;
    if (c0 < D) {
      const __nv_bfloat16* xr = x + static_cast<size_t>(row) * D + c0;
      const float* wa = w2g + static_cast<size_t>(h) * 128 + (lane & 1) * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 v = ld_bf16x4(xr + i);
        const float4 pa = __ldg(reinterpret_cast<const float4*>(wa + i));
        const float4 pb = __ldg(reinterpret_cast<const float4*>(wa + 64 + i));
        a += v.x * pa.x + v.y * pa.y + v.z * pa.z + v.w * pa.w;
        b += v.x * pb.x + v.y * pb.y + v.z * pb.z + v.w * pb.w;
      }
    }
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    b += __shfl_xor_sync(0xffffffffu, b, 1);
    if (c0 < D && (lane & 1) == 0) {
      const float za = rstd * (a - mean * cg[2 * h]) + dg[2 * h];
      const float zb = rstd * (b - mean * cg[2 * h + 1]) + dg[2 * h + 1];
      const float ga = 1.f / (1.f + __expf(-za));
      const float gb = 1.f / (1.f + __expf(-zb));
      gate[static_cast<size_t>(row) * heads + h] = ga * (gb * head_const[h] - 1.f) + 2.f;
    }
  }
}

template <typename F>
int dispatch_nv(int D, F&& f) {
  switch (D / 128) {
    case 2: return f(std::integral_constant<int, 2>{});
    case 4: return f(std::integral_constant<int, 4>{});
    case 6: return f(std::integral_constant<int, 6>{});
    case 8: return f(std::integral_constant<int, 8>{});
    case 16: return f(std::integral_constant<int, 16>{});
    default: return fail(kUnsupported, "row op: D must be one of 256/512/768/1024/2048, got " + std::to_string(D));
  }
}

}  // namespace

int layer_norm(const LayerNormArgs& a, cudaStream_t stream) {
  if (a.rows <= 0) return kOk;
  if (a.D % 128 != 0) return fail(kUnsupported, "layer_norm: D % 128 != 0");
  if ((a.x_f32 == nullptr) == (a.x_bf16 == nullptr)) return fail(kInvalidArgument, "layer_norm: exactly one input");
  const int clip_rows = a.clip_rows > 0 ? a.clip_rows : 1;
  const int clip_valid = a.clip_rows > 0 ? a.clip_valid : 1;
  return dispatch_nv(a.D, [&](auto nv) {
    layer_norm_kernel<decltype(nv)::value><<<ceil_div(a.rows, 8), 256, 0, stream>>>(
        a.x_f32, a.x_bf16, a.gamma, a.beta, a.y_bf16, a.y_f32, a.rows, a.eps, a.gelu, a.stats, clip_rows, clip_valid,
        a.stats_stride);
    SVT_POST_LAUNCH();
    return static_cast<int>(kOk);
  });
}

int wavlm_gate(const __nv_bfloat16* x, int rows, int heads, const float* w2, const float* b2, const float* head_const,
               float* gate, cudaStream_t stream) {
  if (rows <= 0) return kOk;
  wavlm_gate_kernel<<<ceil_div(rows, 8), 256, 0, stream>>>(x, rows, heads, w2, b2, head_const, gate);
  SVT_POST_LAUNCH();
  return kOk;
}

int wavlm_gate_ln(const __nv_bfloat16* x, const float* stats, int rows, int heads, const float* w2g, const float* cg,
                  const float* dg, const float* head_const, float eps, float* gate, cudaStream_t stream) {
  if (rows <= 0) return kOk;
  if (heads % 2 != 0) return fail(kInvalidArgument, "wavlm_gate_ln: the statistics slots need heads * 64 % 128 == 0");
  wavlm_gate_ln_kernel<<<ceil_div(rows, 8), 256, 0, stream>>>(x, stats, rows, heads, w2g, cg, dg, head_const, eps, gate);
  SVT_POST_LAUNCH();
  return kOk;
}

int wavlm_relative_bucket(int relative_position, int num_buckets, int max_distance) {
  // fp32 arithmetic in the order torch evaluates WavLMAttention._relative_positions_bucket
  const int nb = num_buckets / 2;
  int bucket = relative_position > 0 ? nb : 0;
  const int rel = relative_position < 0 ? -relative_position : relative_position;
  const int max_exact = nb / 2;
  if (rel < max_exact) return bucket + rel;
  float v = std::log(static_cast<float>(rel) / static_cast<float>(max_exact));
  v = v / static_cast<float>(std::log(static_cast<double>(max_distance) / max_exact));
  v = v * static_cast<float>(nb - max_exact);
  long large = static_cast<long>(static_cast<float>(max_exact) + v);
  if (large > nb - 1) large = nb - 1;
  return bucket + static_cast<int>(large);
}

int row_stats_cast(const float* x, int rows, int D, __nv_bfloat16* y, float* stats, cudaStream_t stream) {
  if (rows <= 0) return kOk;
  if (D % 128 != 0) return fail(kInvalidArgument, "row_stats_cast: D % 128 != 0");
  row_stats_cast_kernel<<<ceil_div(rows, 8), 256, 0, stream>>>(x, rows, D, y, stats);
  SVT_POST_LAUNCH();
  return kOk;
}

int ln_fold_vectors(const float* w_f32, const __nv_bfloat16* w_packed, const float* beta, float scale, int N, int K,
                    float* colsum, float* bias, cudaStream_t stream) {
  ln_fold_vectors_kernel<<<ceil_div(N, 8), 256, 0, stream>>>(w_f32, w_packed, beta, scale, N, K, colsum, bias);
  SVT_POST_LAUNCH();
  return kOk;
}

int tensor_stats(const float* x, size_t n, double* stats, cudaStream_t stream) {
  SVT_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(double), stream));
  if ((reinterpret_cast<uintptr_t>(x) & 3) != 0) return fail(kInvalidArgument, "tensor_stats: input must be 4-byte aligned");
  const int grid = num_sms() * 4;
  tensor_stats_kernel<<<grid, 256, 0, stream>>>(x, n, stats, 0, 0);
  SVT_POST_LAUNCH();
  return kOk;
}

int tensor_stats_per_clip(const float* x, int clips, size_t n_per_clip, double* stats, cudaStream_t stream) {
  SVT_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(double) * clips, stream));
  if ((reinterpret_cast<uintptr_t>(x) & 3) != 0) return fail(kInvalidArgument, "tensor_stats: input must be 4-byte aligned");
  // CTAs per clip depend on the clip length only, never on the batch: a clip's partial sums (and so its statistics and
  // everything downstream) are the same whichever batch or rank it is processed in
  int gx = static_cast<int>(std::min<size_t>((n_per_clip / 4 + 4095) / 4096, 1024));
  if (gx < 1) gx = 1;
  tensor_stats_kernel<<<dim3(gx, clips), 256, 0, stream>>>(x, n_per_clip, stats, n_per_clip, 2);
  SVT_POST_LAUNCH();
  return kOk;
}

size_t conv0_tables_bytes() { return kC0FragBytes + sizeof(float) * kC0QuadFloats; }
int conv0_build_tables(const float* w_kc, const float* bias, void* tables, cudaStream_t stream) {
  conv0_tables_kernel<<<ceil_div(kC0Tiles * 32, 256), 256, 0, stream>>>(
      w_kc, bias, static_cast<uint4*>(tables), reinterpret_cast<float*>(static_cast<uint8_t*>(tables) + kC0FragBytes));
  SVT_POST_LAUNCH();
  return kOk;
}

int conv0_forward(const Conv0Args& a, cudaStream_t stream) {
  if (a.C != kC0 || a.k != kK0 || a.stride != kS0)
    return fail(kUnsupported, "conv0: only C=512, kernel=10, stride=5 (every wav2vec2/HuBERT checkpoint) is built");
  if (a.t_alloc % kR0 != 0) return fail(kInvalidArgument, "conv0: t_alloc % 4 != 0");
  const double n_in = a.stats_stride > 0 ? static_cast<double>(a.L) : static_cast<double>(a.B) * a.L;
  if (a.layer_mode && a.tc_tables != nullptr && get_option_conv0_impl() != 1) {
    // tensor-core kernel: a block covers iters x 8 warps x 32 frames; ~5 waves of two blocks per SM
    const int block_tiles = ceil_div(a.t_alloc, 8 * kC0FramesPerWarp);
    int iters = ceil_div(block_tiles * a.B, 10 * num_sms());
    if (iters < 1) iters = 1;
    if (iters > 8) iters = 8;
    dim3 grid(ceil_div(block_tiles, iters), a.B);
    const uint4* frag = static_cast<const uint4*>(a.tc_tables);
    const float* quad = reinterpret_cast<const float*>(static_cast<const uint8_t*>(a.tc_tables) + kC0FragBytes);
    conv0_tc_kernel<<<grid, 256, 0, stream>>>(a.wav, a.L, a.T, a.t_alloc, frag, quad, a.gamma, a.beta, a.in_stats, n_in, a.out,
                                              a.stats_stride, iters);
    SVT_POST_LAUNCH();
    return kOk;
  }
  const int groups = a.t_alloc / kR0;
  dim3 grid(ceil_div(groups, 8 * kConv0GroupsPerWarp), a.B);
  if (a.layer_mode) {
    conv0_kernel<true><<<grid, 256, 0, stream>>>(a.wav, a.L, a.T, a.t_alloc, a.w, a.bias, a.gamma, a.beta, a.in_stats,
                                                 n_in, a.out, nullptr, a.stats_stride);
  } else {
    SVT_CUDA(cudaMemsetAsync(a.chan_stats, 0, sizeof(double) * 2 * a.C * a.B, stream));
    conv0_kernel<false><<<grid, 256, 0, stream>>>(a.wav, a.L, a.T, a.t_alloc, a.w, a.bias, nullptr, nullptr, a.in_stats,
                                                  n_in, a.out, a.chan_stats, a.stats_stride);
  }
  SVT_POST_LAUNCH();
  return kOk;
}

int groupnorm_gelu_apply(__nv_bfloat16* x, const double* chan_stats, const float* gamma, const float* beta, int B, int T,
                         int t_alloc, int C, cudaStream_t stream) {
  if (C != 512) return fail(kUnsupported, "groupnorm: C must be 512");
  const int fpb = 64;
  dim3 grid(ceil_div(T, fpb), B);
  groupnorm_gelu_kernel<<<grid, 256, 0, stream>>>(x, chan_stats, gamma, beta, T, t_alloc, C, fpb);
  SVT_POST_LAUNCH();
  return kOk;
}

int head_forward(const HeadArgs& a, cudaStream_t stream) {
  if (a.n_out > 32) return fail(kUnsupported, "head: n_out must be <= 32");
  if (a.clips * a.T <= 0) return kOk;
  return dispatch_nv(a.D, [&](auto nv) {
    constexpr int NV = decltype(nv)::value;
    size_t smem = (a.w != nullptr) ? sizeof(float) * a.n_out * a.D : 0;
    int w_in_smem = 1;
    if (smem > 160 * 1024) { smem = 0; w_in_smem = 0; }
    static std::atomic<unsigned long long> attr_seen{0};
    if (first_use_on_device(attr_seen)) SVT_CUDA(cudaFuncSetAttribute(head_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    int grid = ceil_div(a.clips * a.T, 8 * head_rows(NV));
    if (grid > num_sms()) grid = num_sms();
    head_kernel<NV><<<grid, 256, smem, stream>>>(a.x, a.clips, a.clip_rows, a.T, a.stats, a.eps, a.w, a.b, a.n_out,
                                                 a.feats, a.logits, w_in_smem, a.stats_stride);
    SVT_POST_LAUNCH();
    return static_cast<int>(kOk);
  });
}

int frame_argmax(const float* logits, int n_frames, int n_out, int oct_off, int n_oct, int pc_off, int n_pc, int32_t* oct,
                 int32_t* pc, cudaStream_t stream) {
  if (n_frames <= 0) return kOk;
  frame_argmax_kernel<<<ceil_div(n_frames, 256), 256, 0, stream>>>(logits, n_frames, n_out, oct_off, n_oct, pc_off, n_pc,
                                                                  oct, pc);
  SVT_POST_LAUNCH();
  return kOk;
}

int pack_bf16(const PackArgs& a, __nv_bfloat16* dst, cudaStream_t stream) {
  const size_t n = static_cast<size_t>(a.dims[0]) * a.dims[1] * a.dims[2] * a.dims[3];
  const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, 4096));
  pack_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(a.src, a.dims[0], a.dims[1], a.dims[2], a.dims[3], a.strides[0],
                                                      a.strides[1], a.strides[2], a.strides[3], a.scale, a.vec,
                                                      a.vec_dim, dst);
  SVT_POST_LAUNCH();
  return kOk;
}
int pack_f32(const PackArgs& a, float* dst, cudaStream_t stream) {
  const size_t n = static_cast<size_t>(a.dims[0]) * a.dims[1] * a.dims[2] * a.dims[3];
  const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, 4096));
  pack_kernel<float><<<grid, 256, 0, stream>>>(a.src, a.dims[0], a.dims[1], a.dims[2], a.dims[3], a.strides[0],
                                              a.strides[1], a.strides[2], a.strides[3], a.scale, a.vec, a.vec_dim, dst);
  SVT_POST_LAUNCH();
  return kOk;
}

int weight_norm_scale(const float* v, const float* g, int d0, int d1, int taps, float* out, cudaStream_t stream) {
  weight_norm_scale_kernel<<<taps, 256, 0, stream>>>(v, g, d0 * d1, taps, out);
  SVT_POST_LAUNCH();
  return kOk;
}

int add_positional_encoding(const float* x, int clips, int T_src, int T, int D, float* out_f32, __nv_bfloat16* out_bf16,
                            cudaStream_t stream) {
  dim3 grid(T, clips);
  add_pe_kernel<<<grid, 256, 0, stream>>>(x, T_src, T, D, out_f32, out_bf16);
  SVT_POST_LAUNCH();
  return kOk;
}

int add_f32(const float* a, const float* b, float* out, size_t n, cudaStream_t stream) {
  add_f32_kernel<<<num_sms() * 4, 256, 0, stream>>>(a, b, out, n);
  SVT_POST_LAUNCH();
  return kOk;
}
int channel_affine_bf16(const float* x, const float* scale, const float* shift, __nv_bfloat16* y, size_t rows, int D,
                        cudaStream_t stream) {
  if (D % 4 != 0) return fail(kInvalidArgument, "channel_affine: D % 4 != 0");
  channel_affine_bf16_kernel<<<num_sms() * 4, 256, 0, stream>>>(x, scale, shift, y, rows * (D / 4), D / 4);
  SVT_POST_LAUNCH();
  return kOk;
}
int cast_f32_to_bf16(const float* x, __nv_bfloat16* y, size_t n, cudaStream_t stream) {
  cast_bf16_kernel<<<num_sms() * 4, 256, 0, stream>>>(x, y, n);
  SVT_POST_LAUNCH();
  return kOk;
}

}  // namespace svt
