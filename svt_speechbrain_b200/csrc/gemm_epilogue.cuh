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
  const float* alpha;               // kActPRelu slopes
  const __nv_bfloat16* resid_bf16;  // bf16 residual added BEFORE the activation
  const uint8_t* row_mask;          // rows with mask 0 are written as zeros
  // LayerNorm folded around the GEMM (pre-LN transformer layers, see GemmArgs):
  float* row_stats_out;     // producer: [M][N / 128][2] = (sum, sum of squares) of each 128-column slice of the final
                            // fp32 output rows (plain stores, one writer per slot: deterministic, nothing to zero)
  const float* ln_stats;    // consumer: [M][K / 128][2] partial statistics of the un-normalised A rows
  int ln_slots;             // K / 128 (<= kMaxLnSlots)
  const float* ln_colsum;   // consumer: [N] sums of the (gamma-folded) bf16 weight rows; bias holds beta.W + b
  float ln_inv_d, ln_eps;   // 1 / (LN width = K), LN epsilon
  // row LayerNorm over all N columns fused behind the GEMM (GemmArgs::rowln_*, CTA-pair kernel only)
  const float* rowln_gamma;
  const float* rowln_beta;
  float rowln_eps;
  int rowln_gelu;
};

// What the epilogue fetches one tile ahead of its use.
constexpr int kMaxLnSlots = 8;
struct GemmEpiPrefetch {
  float4 bias;
  float4 colsum;
  float4 st[kMaxLnSlots / 2];  // this thread's row: (sum, sumsq) x 2 slots per element
};

#ifdef __CUDACC__
// Pull a tile's residual rows towards L2 ahead of time (the fp32 residual stream is the only operand of these
// kernels that does not arrive through TMA).  Called by all 256 epilogue threads.
template <int BN>
__device__ __forceinline__ void gemm_prefetch_resid(const GemmEpiParams& p, int row0, int valid, int col0, int n_valid) {
  constexpr int kLinesPerRow = BN * 4 / 128;
  const int et = static_cast<int>(threadIdx.x) % (kNumEpiWarps * 32);  // epilogue warps are physical warps 0-7
  for (int i = et; i < 128 * kLinesPerRow; i += kNumEpiWarps * 32) {
    const int r = i / kLinesPerRow, l = i % kLinesPerRow;
    if (r < valid && l * 32 < n_valid)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p.resid + static_cast<size_t>(row0 + r) * p.ld_out + col0 + l * 32));
  }
}

__device__ __forceinline__ void epi_sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 epi_lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// This warp's slice of the bias (and, with a folded LayerNorm, of the weight column sums plus this thread's row
// statistics) for one tile: BN / 2 columns, 4 per lane, fetched one tile ahead of its use.
template <int BN>
__device__ __forceinline__ GemmEpiPrefetch gemm_epi_prefetch(const GemmEpiParams& p, int row0, int valid, int col0, int n_valid,
                                                             int quad, int half, int lane) {
  GemmEpiPrefetch f;
  f.bias = f.colsum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < kMaxLnSlots / 2; ++i) f.st[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int c = half * (BN / 2) + 4 * lane;
  if (4 * lane < BN / 2 && c < n_valid) {
    if (p.bias != nullptr) f.bias = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + c));
    if (p.ln_colsum != nullptr) f.colsum = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + col0 + c));
  }
  if (p.ln_stats != nullptr && quad * 32 + lane < valid) {
    const float4* st = reinterpret_cast<const float4*>(p.ln_stats + 2 * static_cast<size_t>(row0 + quad * 32 + lane) * p.ln_slots);
#pragma unroll
    for (int i = 0; i < kMaxLnSlots / 2; ++i)
      if (2 * i < p.ln_slots) f.st[i] = st[i];
  }
  return f;
}

// row0 / valid: first output row of this CTA's 128-row tile and how many of its rows exist; col0 / n_valid: first
// output column and valid columns of the tile; tmem_acc: TMEM address (lane 0) of the accumulator stage;
// stage_addr: shared-space address of this warp's 4 KB staging tile; bias4: gemm_load_bias_slice of this tile.
//
// Thread = row in TMEM, but HBM wants lanes along columns, so every 32 x 32 chunk is transposed through the warp's
// private XOR-swizzled staging tile (conflict-free both ways) and leaves as whole row segments:
//   bf16-only outputs : bias (broadcast reads of the slice parked in the upper half of the staging tile) and
//                       activation on the row-per-thread registers, packed bf16 through 2 KB of staging;
//   fp32 / residual   : raw accumulators through 4 KB of staging, then bias, activation and the fp32 residual
//                       (prefetched ahead of the TMEM read) in the coalesced mapping, where a lane keeps the
//                       same 4 columns for all 8 of its rows.
template <int BN, int kMode = 0>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmEpiParams& p, int row0, int valid, int col0, int n_valid,
                                                   uint32_t tmem_acc, int quad, int half, int lane, uint32_t stage_addr,
                                                   const GemmEpiPrefetch& pf, float2* row_sums = nullptr) {
  constexpr int kColsPerWarp = BN / 2;
  // kMode 2 / 3 (own kernel instantiations, gemm_epi_mode() on the host): bf16-only output with the folded LayerNorm and
  // GELU (FFN-1) / no activation (QKV) known at compile time -- no runtime switches in the 32-element chunk body.
  constexpr bool kFixedBf16 = kMode == 2 || kMode == 3 || kMode == 4;  // 4: bias only (conv GEMMs with the fused row LayerNorm)
  // (kMode 5 / 6 take their own branch below and never reach the bias staging of the bf16 path)
  const bool f32_path = !kFixedBf16 && (p.out_f32 != nullptr || p.resid != nullptr || p.resid_bf16 != nullptr ||
                                        p.row_mask != nullptr || p.act == kActPRelu);
  const int act = kMode == 2 ? static_cast<int>(kActGelu) : (kMode == 3 || kMode == 4 ? static_cast<int>(kActNone) : p.act);
  const int c4 = lane & 7;
  const uint32_t bias_slot = stage_addr + 2048, colsum_slot = stage_addr + 2048 + 512;
  const bool ln = kMode == 2 || kMode == 3 || (kMode == 0 && p.ln_stats != nullptr);  // host side guarantees !f32_path and bias != nullptr with it
  float ln_rstd = 1.f, ln_shift = 0.f;
  if (!f32_path && p.bias != nullptr) {
    if (4 * lane < kColsPerWarp) {
      epi_sts128(bias_slot + 16 * lane, __float_as_uint(pf.bias.x), __float_as_uint(pf.bias.y), __float_as_uint(pf.bias.z),
                 __float_as_uint(pf.bias.w));
      if (ln)
        epi_sts128(colsum_slot + 16 * lane, __float_as_uint(pf.colsum.x), __float_as_uint(pf.colsum.y),
                   __float_as_uint(pf.colsum.z), __float_as_uint(pf.colsum.w));
    }
    __syncwarp();
  }
  if (ln) {
    // y = rstd * (x.W' - mean * colsum) + (beta.W + b): the row statistics were accumulated by the producing GEMM
    float sum = 0.f, sumsq = 0.f;  // fixed summation order over the slots
#pragma unroll
    for (int i = 0; i < kMaxLnSlots / 2; ++i) {
      sum += pf.st[i].x; sumsq += pf.st[i].y;
      sum += pf.st[i].z; sumsq += pf.st[i].w;
    }
    const float mean = sum * p.ln_inv_d;
    const float var = fmaxf(sumsq * p.ln_inv_d - mean * mean, 0.f);
    ln_rstd = rsqrtf(var + p.ln_eps);
    ln_shift = -mean * ln_rstd;
  }
  float ps[8], pss[8];  // producer side: per-lane partial row statistics over this warp's columns
#pragma unroll
  for (int i = 0; i < 8; ++i) ps[i] = pss[i] = 0.f;
  // The TMEM read of chunk c + 1 is issued before the math of chunk c and stays in flight under it.  The chunk loop is
  // unrolled by two over two register arrays that swap roles, so no accumulator is ever copied between registers.
  const uint32_t tmem_row = tmem_acc + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(half * kColsPerWarp);

  // one 32-column chunk whose accumulators are arriving in `acc`; `nxt` receives the following chunk's TMEM read
  auto chunk_cols = [&](int c, int* col, int* nv) {
    const int col_in_tile = half * kColsPerWarp + c;
    *col = col0 + col_in_tile;
    *nv = n_valid - col_in_tile;  // valid columns of this 32-wide chunk
  };
  // kMode = 1 (own kernel instantiation, chosen on the host by gemm_epi_resid_fast): the residual GEMMs of the transformer
  // layers -- bias + fp32 residual in place + bf16 copy + row statistics over whole 256-column tiles.  Same order of memory
  // operations as the generic path below (the chunk's eight residual loads in one burst ahead of the TMEM wait), but the
  // pointers and row predicates are formed once per tile, the accumulators go to the staging tile straight from the TMEM
  // registers and none of the other variants' branches are compiled in: the generic path spends ~29 instructions per
  // element at a per-warp IPC of 0.14, and that, not HBM, paces the out-projection (profiles/r2_rejected_experiments.txt).
  if constexpr (kMode == 1) {
    const int rbase = quad * 32 + (lane >> 3);
    uint32_t vmask = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) vmask |= (rbase + 4 * i < valid) ? (1u << i) : 0u;
    const size_t ld = static_cast<size_t>(p.ld_out);
    const size_t base = static_cast<size_t>(row0 + rbase) * ld + static_cast<size_t>(col0 + half * kColsPerWarp + 4 * c4);
    const float* rp = p.resid + base;
    float* of = p.out_f32 + base;
    __nv_bfloat16* ob = p.out_bf16 + base;
    const float* bp = p.bias + col0 + half * kColsPerWarp + 4 * c4;
    const size_t step = 4 * ld;
    const bool want_stats = p.row_stats_out != nullptr;
    uint32_t rn[32];
    tmem_ld32(tmem_row, rn);
#pragma unroll 1
    for (int c = 0; c < kColsPerWarp; c += 32) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp + c));
      float4 rs[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        rs[i] = ((vmask >> i) & 1u) ? *reinterpret_cast<const float4*>(rp + i * step + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      tmem_ld_wait_regs(rn);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        epi_sts128(stage_addr + lane * 128 + ((q ^ (lane & 7)) << 4), rn[4 * q], rn[4 * q + 1], rn[4 * q + 2], rn[4 * q + 3]);
      if (c + 32 < kColsPerWarp) tmem_ld32(tmem_row + c + 32, rn);
      __syncwarp();
      uint4 raw[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + (lane >> 3);
        raw[i] = epi_lds128(stage_addr + rr * 128 + ((c4 ^ (rr & 7)) << 4));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 a = make_float4(__uint_as_float(raw[i].x) + b4.x + rs[i].x, __uint_as_float(raw[i].y) + b4.y + rs[i].y,
                                     __uint_as_float(raw[i].z) + b4.z + rs[i].z, __uint_as_float(raw[i].w) + b4.w + rs[i].w);
        if ((vmask >> i) & 1u) {
          if (want_stats) {
            ps[i] += (a.x + a.y) + (a.z + a.w);
            pss[i] = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, pss[i]))));
          }
          *reinterpret_cast<float4*>(of + i * step + c) = a;
          *reinterpret_cast<uint2*>(ob + i * step + c) = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
        }
      }
      __syncwarp();
    }
  } else if constexpr (kMode == 5 || kMode == 6) {
    // kMode 5 / 6 (one-CTA kernel, the ResNet convolutions of the video stream): out = ring_mask * PReLU(acc + bias
    // (+ bf16 residual)) as bf16, whole BN-column tiles.  Same order of memory operations as the generic path; row
    // predicates, the padding-ring mask and the pointers are formed once per tile.
    const int rbase = quad * 32 + (lane >> 3);
    uint32_t vmask = 0, keep = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (rbase + 4 * i < valid) {
        vmask |= 1u << i;
        if (p.row_mask == nullptr || p.row_mask[static_cast<size_t>(row0 + rbase + 4 * i)] != 0) keep |= 1u << i;
      }
    }
    const size_t ld = static_cast<size_t>(p.ld_out);
    const size_t base = static_cast<size_t>(row0 + rbase) * ld + static_cast<size_t>(col0 + half * kColsPerWarp + 4 * c4);
    __nv_bfloat16* ob = p.out_bf16 + base;
    const __nv_bfloat16* rbp = (kMode == 6) ? p.resid_bf16 + base : nullptr;
    const float* bp = p.bias + col0 + half * kColsPerWarp + 4 * c4;
    const float* ap = p.alpha + col0 + half * kColsPerWarp + 4 * c4;
    const size_t step = 4 * ld;
    uint32_t rn[32];
    tmem_ld32(tmem_row, rn);
#pragma unroll 1
    for (int c = 0; c < kColsPerWarp; c += 32) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp + c));
      const float4 al = __ldg(reinterpret_cast<const float4*>(ap + c));
      uint2 rbv[8];
      if constexpr (kMode == 6) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          rbv[i] = ((vmask >> i) & 1u) ? *reinterpret_cast<const uint2*>(rbp + i * step + c) : make_uint2(0u, 0u);
      }
      tmem_ld_wait_regs(rn);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        epi_sts128(stage_addr + lane * 128 + ((q ^ (lane & 7)) << 4), rn[4 * q], rn[4 * q + 1], rn[4 * q + 2], rn[4 * q + 3]);
      if (c + 32 < kColsPerWarp) tmem_ld32(tmem_row + c + 32, rn);
      __syncwarp();
      uint4 raw[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + (lane >> 3);
        raw[i] = epi_lds128(stage_addr + rr * 128 + ((c4 ^ (rr & 7)) << 4));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 a = make_float4(__uint_as_float(raw[i].x) + b4.x, __uint_as_float(raw[i].y) + b4.y,
                               __uint_as_float(raw[i].z) + b4.z, __uint_as_float(raw[i].w) + b4.w);
        if constexpr (kMode == 6) {
          const __nv_bfloat162 r01 = *reinterpret_cast<const __nv_bfloat162*>(&rbv[i].x);
          const __nv_bfloat162 r23 = *reinterpret_cast<const __nv_bfloat162*>(&rbv[i].y);
          a.x += __low2float(r01); a.y += __high2float(r01); a.z += __low2float(r23); a.w += __high2float(r23);
        }
        a.x = a.x > 0.f ? a.x : a.x * al.x; a.y = a.y > 0.f ? a.y : a.y * al.y;
        a.z = a.z > 0.f ? a.z : a.z * al.z; a.w = a.w > 0.f ? a.w : a.w * al.w;
        if (!((keep >> i) & 1u)) a = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((vmask >> i) & 1u)
          *reinterpret_cast<uint2*>(ob + i * step + c) = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
      }
      __syncwarp();
    }
  } else if (f32_path) {
    // fp32 / residual outputs: one register array, copied once per chunk
    uint32_t rn[32];
    tmem_ld32(tmem_row, rn);
#pragma unroll 1
    for (int c = 0; c < kColsPerWarp; c += 32) {
      int col, nv;
      chunk_cols(c, &col, &nv);
      uint32_t acc[32];
      // residual + bias of this chunk in the coalesced mapping (lane = 4 columns of rows 4i + lane / 8), all loads
      // issued before the TMEM read so their latency overlaps it (resid may alias out_f32: loads come first)
      float4 rs[8];
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), al = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias != nullptr && 4 * c4 < nv) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4 * c4));
      if (p.act == kActPRelu && 4 * c4 < nv) al = __ldg(reinterpret_cast<const float4*>(p.alpha + col + 4 * c4));
      uint2 rb[8];
      uint32_t keep = 0xffu;  // bit i: row 4i + lane / 8 is kept (not a padding-ring row)
      if (p.resid_bf16 != nullptr || p.row_mask != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = 4 * i + (lane >> 3);
          rb[i] = make_uint2(0u, 0u);
          if (quad * 32 + rr < valid && 4 * c4 < nv) {
            const size_t row = static_cast<size_t>(row0 + quad * 32 + rr);
            if (p.resid_bf16 != nullptr)
              rb[i] = *reinterpret_cast<const uint2*>(p.resid_bf16 + row * static_cast<size_t>(p.ld_out) + static_cast<size_t>(col + 4 * c4));
            if (p.row_mask != nullptr && p.row_mask[row] == 0) keep &= ~(1u << i);
          }
        }
      }
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
      tmem_ld_wait_regs(rn);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = rn[j];
      if (c + 32 < kColsPerWarp) tmem_ld32(tmem_row + c + 32, rn);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        epi_sts128(stage_addr + lane * 128 + ((q ^ (lane & 7)) << 4), acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
      __syncwarp();
      uint4 raw[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + (lane >> 3);
        raw[i] = epi_lds128(stage_addr + rr * 128 + ((c4 ^ (rr & 7)) << 4));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + (lane >> 3);
        float4 a = make_float4(__uint_as_float(raw[i].x) + b4.x, __uint_as_float(raw[i].y) + b4.y,
                               __uint_as_float(raw[i].z) + b4.z, __uint_as_float(raw[i].w) + b4.w);
        if (p.resid_bf16 != nullptr) {
          const __nv_bfloat162 r01 = *reinterpret_cast<const __nv_bfloat162*>(&rb[i].x);
          const __nv_bfloat162 r23 = *reinterpret_cast<const __nv_bfloat162*>(&rb[i].y);
          a.x += __low2float(r01); a.y += __high2float(r01); a.z += __low2float(r23); a.w += __high2float(r23);
        }
        if (p.act == kActGelu) {
          a.x = gelu_erf(a.x); a.y = gelu_erf(a.y); a.z = gelu_erf(a.z); a.w = gelu_erf(a.w);
        } else if (p.act == kActRelu) {
          a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
        } else if (p.act == kActPRelu) {
          a.x = a.x > 0.f ? a.x : a.x * al.x; a.y = a.y > 0.f ? a.y : a.y * al.y;
          a.z = a.z > 0.f ? a.z : a.z * al.z; a.w = a.w > 0.f ? a.w : a.w * al.w;
        }
        if (!((keep >> i) & 1u)) a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (quad * 32 + rr < valid && 4 * c4 < nv) {
          const size_t off = static_cast<size_t>(row0 + quad * 32 + rr) * static_cast<size_t>(p.ld_out) +
                             static_cast<size_t>(col + 4 * c4);
          if (p.resid != nullptr) { a.x += rs[i].x; a.y += rs[i].y; a.z += rs[i].z; a.w += rs[i].w; }
          if (p.row_stats_out != nullptr) {
            ps[i] += (a.x + a.y) + (a.z + a.w);
            pss[i] = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, pss[i]))));
          }
          if (p.out_f32 != nullptr) *reinterpret_cast<float4*>(p.out_f32 + off) = a;
          if (p.out_bf16 != nullptr)
            *reinterpret_cast<uint2*>(p.out_bf16 + off) = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
        }
      }
      __syncwarp();
    }
  } else {
    // bf16-only outputs (QKV, FFN-1 with GELU: issue-bound): the loop is unrolled by two over two register arrays
    // that swap roles, so no accumulator is ever copied between registers
    float rsum = 0.f, rsq = 0.f;  // row_sums: this thread's row over this warp's columns (after bias, before activation)
    auto chunk = [&](uint32_t (&acc)[32], uint32_t (&nxt)[32], int c) {
      int col, nv;
      chunk_cols(c, &col, &nv);
      tmem_ld_wait_regs(acc);
      if (c + 32 < kColsPerWarp) tmem_ld32(tmem_row + c + 32, nxt);
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
      if (ln) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 b = epi_lds128(bias_slot + (c + 4 * j) * 4);
          const uint4 cs = epi_lds128(colsum_slot + (c + 4 * j) * 4);
          v[4 * j + 0] = fmaf(v[4 * j + 0], ln_rstd, fmaf(ln_shift, __uint_as_float(cs.x), __uint_as_float(b.x)));
          v[4 * j + 1] = fmaf(v[4 * j + 1], ln_rstd, fmaf(ln_shift, __uint_as_float(cs.y), __uint_as_float(b.y)));
          v[4 * j + 2] = fmaf(v[4 * j + 2], ln_rstd, fmaf(ln_shift, __uint_as_float(cs.z), __uint_as_float(b.z)));
          v[4 * j + 3] = fmaf(v[4 * j + 3], ln_rstd, fmaf(ln_shift, __uint_as_float(cs.w), __uint_as_float(b.w)));
        }
      } else if (p.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 b = epi_lds128(bias_slot + (c + 4 * j) * 4);  // same address in every lane: broadcast
          v[4 * j + 0] += __uint_as_float(b.x); v[4 * j + 1] += __uint_as_float(b.y);
          v[4 * j + 2] += __uint_as_float(b.z); v[4 * j + 3] += __uint_as_float(b.w);
        }
      }
      if (row_sums != nullptr) {
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          s0 += v[j]; s1 += v[j + 1];
          q0 = fmaf(v[j], v[j], q0); q1 = fmaf(v[j + 1], v[j + 1], q1);
        }
        rsum += s0 + s1; rsq += q0 + q1;
      }
      if (act == kActGelu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
      } else if (act == kActRelu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        epi_sts128(stage_addr + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4), pack_bf16x2(v[8 * q], v[8 * q + 1]),
                   pack_bf16x2(v[8 * q + 2], v[8 * q + 3]), pack_bf16x2(v[8 * q + 4], v[8 * q + 5]),
                   pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
      __syncwarp();
      const int sl = lane & 3;
      uint4 o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = 8 * i + (lane >> 2);
        o[i] = epi_lds128(stage_addr + rr * 64 + ((sl ^ ((rr >> 1) & 3)) << 4));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = 8 * i + (lane >> 2);
        if (quad * 32 + rr < valid && 8 * sl < nv)
          *reinterpret_cast<uint4*>(p.out_bf16 + static_cast<size_t>(row0 + quad * 32 + rr) * static_cast<size_t>(p.ld_out) +
                                    static_cast<size_t>(col + 8 * sl)) = o[i];
      }
      __syncwarp();
    };
    uint32_t ra[32], rb2[32];
    tmem_ld32(tmem_row, ra);
#pragma unroll 1
    for (int c = 0; c < kColsPerWarp; c += 64) {
      chunk(ra, rb2, c);
      if (c + 32 < kColsPerWarp) chunk(rb2, ra, c + 32);
    }
    if (row_sums != nullptr) *row_sums = make_float2(rsum, rsq);
  }
  if constexpr (BN == 256) {
    if (p.row_stats_out != nullptr) {
      // lanes 8r .. 8r + 7 hold partials of row 4i + r over this warp's 128 columns = one slot of that row
      const int slots = p.ld_out >> 7;
      const int slot = (col0 >> 7) + half;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          ps[i] += __shfl_xor_sync(0xffffffffu, ps[i], o);
          pss[i] += __shfl_xor_sync(0xffffffffu, pss[i], o);
        }
        const int rr = 4 * i + (lane >> 3);
        if (c4 == 0 && quad * 32 + rr < valid)
          *reinterpret_cast<float2*>(p.row_stats_out + 2 * (static_cast<size_t>(row0 + quad * 32 + rr) * slots + slot)) =
              make_float2(ps[i], pss[i]);
      }
    }
  }
}
#endif  // __CUDACC__

// the residual GEMMs of the transformer layers qualify for the specialised epilogue (kMode = 1)
inline bool gemm_epi_resid_fast(const GemmEpiParams& p, int N) {
  return N % 256 == 0 && p.resid != nullptr && p.out_f32 != nullptr && p.out_bf16 != nullptr && p.resid_bf16 == nullptr &&
         p.row_mask == nullptr && p.act == kActNone && p.bias != nullptr && p.ln_stats == nullptr && p.rowln_gamma == nullptr;
}
// epilogue variant of a CTA-pair launch: 1 residual GEMM, 2 folded LayerNorm + GELU (FFN-1), 3 folded LayerNorm (QKV), 0 generic
inline int gemm_epi_mode(const GemmEpiParams& p, int N) {
  if (gemm_epi_resid_fast(p, N)) return 1;
  const bool bf16_ln = p.ln_stats != nullptr && p.bias != nullptr && p.ln_colsum != nullptr && p.out_bf16 != nullptr &&
                       p.out_f32 == nullptr && p.resid == nullptr && p.resid_bf16 == nullptr && p.row_mask == nullptr &&
                       p.rowln_gamma == nullptr && p.row_stats_out == nullptr;
  if (bf16_ln && p.act == kActGelu) return 2;
  if (bf16_ln && p.act == kActNone) return 3;
  if (p.rowln_gamma != nullptr && p.bias != nullptr) return 4;  // gemm_rowln_supported() has checked the rest
  return 0;
}
// process-wide option "resid_epilogue": which specialised epilogue instantiations of the CTA-pair kernel may be used:
// 0 none (generic kernel), 1 the residual GEMMs, 2 (default) also QKV / FFN-1 with the folded LayerNorm
int get_option_resid_epilogue();

// CTA-pair (cta_group::2) kernel, gemm_tc2.cu
bool gemm_pair_supported(const GemmArgs& g);
int gemm_bf16_tc_pair(const GemmArgs& g, cudaStream_t stream);
// process-wide option "gemm_impl": 0 auto (pair kernel when supported), 1 force the one-CTA kernel
int get_option_gemm_impl();

}  // namespace svt
