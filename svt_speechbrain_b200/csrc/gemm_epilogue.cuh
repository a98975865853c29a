// Epilogue shared by the one-CTA (gemm_tc.cu) and CTA-pair (gemm_tc2.cu) tcgen05 GEMM kernels:
// TMEM accumulator tile (128 rows x BN fp32 columns of one CTA) -> bias / GELU / ReLU -> (+ fp32 residual) ->
// fp32 and / or bf16 rows in HBM.  8 epilogue warps per CTA: warp w owns TMEM lanes 32 * (w % 4) .. + 31 and
// column half (w - 4) / 4 of the tile.
#pragma once
#include "gemm_tc.cuh"

namespace svt {

constexpr int kEpiWarp0 = 4;
constexpr int kNumEpiWarps = 8;
constexpr int kEpiStageBytes = 4096;  // per-warp 32 x 32 fp32 transpose tile

struct GemmEpiParams {
  int M;  // valid rows of the whole problem (linear mode)
  const float* bias;
  const float* resid;
  float* out_f32;
  __nv_bfloat16* out_bf16;
  int ld_out, act;
};

#ifdef __CUDACC__
// Pull a tile's residual rows towards L2 ahead of time (the fp32 residual stream is the only operand of these
// kernels that does not arrive through TMA).  Called by all 256 epilogue threads.
template <int BN>
__device__ __forceinline__ void gemm_prefetch_resid(const GemmEpiParams& p, int row0, int valid, int col0, int n_valid) {
  constexpr int kLinesPerRow = BN * 4 / 128;
  const int et = threadIdx.x - kEpiWarp0 * 32;
  for (int i = et; i < 128 * kLinesPerRow; i += kNumEpiWarps * 32) {
    const int r = i / kLinesPerRow, l = i % kLinesPerRow;
    if (r < valid && l * 32 < n_valid)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p.resid + static_cast<size_t>(row0 + r) * p.ld_out + col0 + l * 32));
  }
}

// row0 / valid: first output row of this CTA's 128-row tile and how many of its rows exist; col0 / n_valid: first
// output column and valid columns of the tile; tmem_acc: TMEM address (lane 0) of the accumulator stage.
template <int BN>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmEpiParams& p, int row0, int valid, int col0, int n_valid,
                                                   uint32_t tmem_acc, int quad, int half, int lane, uint8_t* stage_mine) {
  constexpr int kColsPerWarp = BN / 2;
  const bool f32_path = (p.out_f32 != nullptr || p.resid != nullptr);
  const int c4 = lane & 7;
#pragma unroll 1
  for (int c = 0; c < kColsPerWarp; c += 32) {
    const int col_in_tile = half * kColsPerWarp + c;
    const int col = col0 + col_in_tile;
    const int nv = n_valid - col_in_tile;  // valid columns of this 32-wide chunk
    // residual chunk in the coalesced store mapping (lane = 4 columns of row 4i + lane/8), all eight loads
    // issued before the TMEM read so their latency overlaps it (resid may alias out_f32: loads come first)
    float4 rs[8];
    if (p.resid != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + (lane >> 3);
        rs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (quad * 32 + rr < valid && 4 * c4 < nv)
          rs[i] = *reinterpret_cast<const float4*>(p.resid + static_cast<size_t>(row0 + quad * 32 + rr) * static_cast<size_t>(p.ld_out) +
                                                   static_cast<size_t>(col + 4 * c4));
      }
    }
    uint32_t r[32];
    tmem_ld32(tmem_acc + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(col_in_tile), r);
    tmem_ld_wait();
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (p.bias != nullptr) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (4 * j < nv) {
          const float4 b = __ldg(b4 + j);
          v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
        }
      }
    }
    if (p.act == kActGelu) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
    } else if (p.act == kActRelu) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
    }
    // Thread = row in TMEM, but HBM wants lanes along columns: transpose the 32 x 32 chunk through this
    // warp's private XOR-swizzled staging tile (conflict-free both ways), then do coalesced row segments.
    if (f32_path) {
      float* st = reinterpret_cast<float*>(stage_mine);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(st + lane * 32 + ((q ^ (lane & 7)) << 2)) =
            make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + (lane >> 3);
        float4 a = *reinterpret_cast<const float4*>(st + rr * 32 + ((c4 ^ (rr & 7)) << 2));
        if (quad * 32 + rr < valid && 4 * c4 < nv) {
          const size_t off = static_cast<size_t>(row0 + quad * 32 + rr) * static_cast<size_t>(p.ld_out) +
                             static_cast<size_t>(col + 4 * c4);
          if (p.resid != nullptr) { a.x += rs[i].x; a.y += rs[i].y; a.z += rs[i].z; a.w += rs[i].w; }
          if (p.out_f32 != nullptr) *reinterpret_cast<float4*>(p.out_f32 + off) = a;
          if (p.out_bf16 != nullptr)
            *reinterpret_cast<uint2*>(p.out_bf16 + off) = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
        }
      }
      __syncwarp();
    } else {
      uint8_t* st = stage_mine;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(st + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) =
            make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                       pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
      __syncwarp();
      const int sl = lane & 3;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = 8 * i + (lane >> 2);
        const uint4 a = *reinterpret_cast<const uint4*>(st + rr * 64 + ((sl ^ ((rr >> 1) & 3)) << 4));
        if (quad * 32 + rr < valid && 8 * sl < nv)
          *reinterpret_cast<uint4*>(p.out_bf16 + static_cast<size_t>(row0 + quad * 32 + rr) * static_cast<size_t>(p.ld_out) +
                                    static_cast<size_t>(col + 8 * sl)) = a;
      }
      __syncwarp();
    }
  }
}
#endif  // __CUDACC__

// CTA-pair (cta_group::2) kernel, gemm_tc2.cu
bool gemm_pair_supported(const GemmArgs& g);
int gemm_bf16_tc_pair(const GemmArgs& g, cudaStream_t stream);
// process-wide option "gemm_impl": 0 auto (pair kernel when supported), 1 force the one-CTA kernel
int get_option_gemm_impl();

}  // namespace svt
