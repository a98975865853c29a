// CTA-pair variant of the bf16 GEMM (cta_group::2): two CTAs of one cluster (the two SMs of a TPC) share every
// MMA.  One 256 x 256 output tile per pair and step: CTA r stages A rows [128 r, 128 r + 128) and W rows
// [128 r, 128 r + 128) of the tile; the leader's single thread issues tcgen05.mma.cta_group::2 (M = 256,
// N = 256, K = 16), which reads both CTAs' shared memory and writes accumulator rows [128 r, +128) into CTA r's
// TMEM.  Compared with the one-CTA kernel (gemm_tc.cu) each SM stages and reads half of the B operand:
// shared-memory traffic per SM drops from 96 + 96 B/clk (TMA fill + MMA reads) to 64 + 64 B/clk, which is what
// capped the one-CTA kernel near 80 % of the tensor peak; the smaller stage also buys a 6-deep ring.
//
// Barrier protocol (all mbarriers live at the same shared-memory offsets in both CTAs):
//   full[s]   (leader's copy only)  : armed by the leader's producer with the bytes of BOTH CTAs; both producers'
//                                     TMA loads complete_tx on it (cp.async.bulk.tensor ... cta_group::2)
//   empty[s]  (each CTA's own copy) : tcgen05.commit ... multicast::cluster from the leader's MMA thread
//   tfull[a]  (each CTA's own copy) : multicast commit after the last k-block of a tile
//   tempty[a] (leader's copy only)  : 2 x 8 epilogue warps arrive (the peer's through shared::cluster)
// Roles per CTA as in gemm_tc.cu: warp 0 TMA producer, warp 1 TMEM allocator (+ MMA issuer in the leader),
// warps 4-11 epilogue over the CTA's own 128 accumulator rows.
#include "gemm_tc.cuh"
#include "gemm_epilogue.cuh"

namespace svt {

namespace {

constexpr int BM = 128;   // rows per CTA (256 per pair)
constexpr int BN2 = 256;  // columns per pair tile
constexpr int BK = 64;
constexpr int kThreads = 384;
constexpr int kStages2 = 6;
constexpr int kABytes = BM * BK * 2;         // 16 KB
constexpr int kBBytes = (BN2 / 2) * BK * 2;  // 16 KB: this CTA's half of the W tile
constexpr int kStageBytes2 = kABytes + kBBytes;
constexpr int kTmemCols2 = 512;              // two accumulator stages of 256 columns
constexpr int kSmemBytes2 = kStages2 * kStageBytes2 + kNumEpiWarps * kEpiStageBytes + 1024 + 256;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all prior MMAs of this thread retire) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
      : "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

// Fused row LayerNorm (GemmArgs::rowln_*), N = 2 * BN2 = 512 columns.  Every epilogue warp parks the (sum, sum of squares) of
// its 32 rows x 128 columns in the unused last kilobyte of its staging tile: [unit parity][n tile][lane] float2.
constexpr int kRowLnSlot = 3072;
constexpr int kRowLnN = 2 * BN2;

// Second half of the fused LayerNorm: this CTA's 128 rows x 512 columns were just written un-normalised (bf16) by its own
// epilogue warps and are still in L2; epilogue warp ew normalises rows [16 ew, 16 ew + 16) in place, a lane owning columns
// [8 lane, +8) and [256 + 8 lane, +8) (whole 512-byte halves of a row per warp instruction).
__device__ __forceinline__ void rowln_finish(const GemmEpiParams& p, int row0, int valid, const uint8_t* stage_area, int upar,
                                             int ew, int lane) {
  float g[16], b[16];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const float4 gv = __ldg(reinterpret_cast<const float4*>(p.rowln_gamma + h * BN2 + 8 * lane + 4 * v));
      const float4 bv = __ldg(reinterpret_cast<const float4*>(p.rowln_beta + h * BN2 + 8 * lane + 4 * v));
      g[8 * h + 4 * v + 0] = gv.x; g[8 * h + 4 * v + 1] = gv.y; g[8 * h + 4 * v + 2] = gv.z; g[8 * h + 4 * v + 3] = gv.w;
      b[8 * h + 4 * v + 0] = bv.x; b[8 * h + 4 * v + 1] = bv.y; b[8 * h + 4 * v + 2] = bv.z; b[8 * h + 4 * v + 3] = bv.w;
    }
  // The rows come back from L2 (~700 cycles) and two epilogue warps per sub-partition cannot hide that with math, so they
  // are fetched asynchronously (cp.async.cg, L2 only: coherent with the other warps' stores ordered by the barrier) three
  // rows ahead into the first 3 KB of this warp's staging tile; a lane copies and later reads only its own 2 x 16 bytes.
  constexpr int kAhead = 3;
  const int r_begin = 16 * ew;
  const int r_end = min(r_begin + 16, valid);
  auto row_ptr = [&](int r) { return p.out_bf16 + static_cast<size_t>(row0 + r) * static_cast<size_t>(p.ld_out); };
  const uint32_t slot0 = smem_u32(stage_area + ew * kEpiStageBytes) + 16 * lane;
  auto fetch = [&](int r) {
    if (r < r_end) {
      const uint32_t dst = slot0 + static_cast<uint32_t>(((r - r_begin) % kAhead) * 1024);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(row_ptr(r) + 8 * lane) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 512), "l"(row_ptr(r) + BN2 + 8 * lane) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");  // one group per row, empty past the end: the wait below counts groups
  };
#pragma unroll
  for (int i = 0; i < kAhead; ++i) fetch(r_begin + i);
#pragma unroll 1
  for (int r = r_begin; r < r_end; ++r) {
    asm volatile("cp.async.wait_group %0;" ::"n"(kAhead - 1) : "memory");
    const uint32_t src = slot0 + static_cast<uint32_t>(((r - r_begin) % kAhead) * 1024);
    const uint4 x0 = epi_lds128(src), x1 = epi_lds128(src + 512);
    // four partial statistics of the row: column halves (warps q, q + 4) of the two tiles, fixed order
    const uint8_t* st = stage_area + (r >> 5) * kEpiStageBytes + kRowLnSlot + upar * 512 + (r & 31) * 8;
    float sum = 0.f, sq = 0.f;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh)
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const float2 v = *reinterpret_cast<const float2*>(st + hh * 4 * kEpiStageBytes + t * 256);
        sum += v.x; sq += v.y;
      }
    const float mean = sum * (1.0f / kRowLnN);
    const float rstd = rsqrtf(fmaxf(sq * (1.0f / kRowLnN) - mean * mean, 0.f) + p.rowln_eps);
    const float shift = -mean * rstd;
    const uint32_t w[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
    float y[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float lo = __uint_as_float(w[j] << 16), hi = __uint_as_float(w[j] & 0xffff0000u);
      y[2 * j] = fmaf(fmaf(lo, rstd, shift), g[2 * j], b[2 * j]);
      y[2 * j + 1] = fmaf(fmaf(hi, rstd, shift), g[2 * j + 1], b[2 * j + 1]);
    }
    fetch(r + kAhead);  // this row's slot is free again: its 32 bytes are in registers
    if (p.rowln_gelu) {
#pragma unroll
      for (int j = 0; j < 16; ++j) y[j] = gelu_erf(y[j]);
    }
    *reinterpret_cast<uint4*>(row_ptr(r) + 8 * lane) =
        make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
    *reinterpret_cast<uint4*>(row_ptr(r) + BN2 + 8 * lane) =
        make_uint4(pack_bf16x2(y[8], y[9]), pack_bf16x2(y[10], y[11]), pack_bf16x2(y[12], y[13]), pack_bf16x2(y[14], y[15]));
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// kEpi: 0 = every epilogue variant; 1 = only the residual path of gemm_epilogue_tile (out-projection / FFN-2); 2 / 3 = folded
// LayerNorm with / without GELU (FFN-1 / QKV); 4 = bias only, with the fused row LayerNorm behind it (conv feature extractor)
template <int kEpi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmEpiParams p,
                int k_inner, int num_kb, int m_pairs, int n_tiles, int n_inner) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_area = smem + kStages2 * kStageBytes2;  // 8 epilogue warps x 4 KB transpose tiles
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_area + kNumEpiWarps * kEpiStageBytes);
  uint64_t* empty_bar = full_bar + kStages2;
  uint64_t* tfull_bar = empty_bar + kStages2;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  // logical warp = role index; physical warps 0-7 run the epilogue and 8-11 the producer / MMA issuer, because the
  // sub-partition scheduler favours the higher warp id: the issuer must not starve behind epilogue math
  const int warp = ((threadIdx.x >> 5) + 4) % 12;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  // Work unit = n_inner consecutive tiles (tile = m_pair * n_tiles + n_tile, n fastest) handled back to back by one pair:
  // n_inner = 1 normally; n_inner = n_tiles with the fused row LayerNorm, so that a pair owns whole output rows.
  const int num_units = m_pairs * n_tiles / n_inner;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages2; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * kNumEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair<kTmemCols2>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised and its TMEM is allocated before anything touches them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // register budget: the producer / issuer warpgroup needs few registers, the epilogue warpgroups get the rest
  if (warp < kEpiWarp0) {
  setmaxnreg_dec<64>();
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    // This loop runs once per k-block on a sub-partition it shares with two epilogue warps, so every instruction in it
    // competes with the epilogue math for issue slots: a slow producer, not the tensor pipe, was what capped the GELU
    // GEMM at 66 % pipe activity (profiles/r2_ffn1_producer_stall.txt: the MMA warp waited on `full` 45 % of its time
    // while this warp was busy 87 % of its time, mostly in the integer division of k0 by k_inner).  Hence: no division
    // (the (column, tap) coordinates of the A view advance incrementally), addresses formed by adds from per-role constants.
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t full0_leader = mapa_shared(smem_u32(&full_bar[0]), 0);
    const uint32_t smem_base = smem_u32(smem);
    const int b_row_off = static_cast<int>(rank) * (BN2 / 2);
    for (int unit = pair; unit < num_units; unit += num_pairs)
    for (int ui = 0; ui < n_inner; ++ui) {
      const int tile = unit * n_inner + ui;
      const int n_tile = tile % n_tiles;
      const int m_row = ((tile / n_tiles) * 2 + static_cast<int>(rank)) * BM;
      const int b_row = n_tile * BN2 + b_row_off;
      int kc = 0, tap = 0;  // k-block kb covers columns [kc, kc + 64) of tap `tap`: kb * 64 = tap * k_inner + kc
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          const uint32_t sa = smem_base + static_cast<uint32_t>(stage * kStageBytes2);
          const uint32_t bar = full0_leader + static_cast<uint32_t>(stage * 8);
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * kStageBytes2);
          asm volatile(
              "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
              ::"r"(sa), "l"(&tmA), "r"(bar), "r"(kc), "r"(tap), "r"(m_row)
              : "memory");
          asm volatile(
              "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
              ::"r"(sa + kABytes), "l"(&tmB), "r"(bar), "r"(kb * BK), "r"(b_row)
              : "memory");
        }
        __syncwarp();
        kc += BK;
        if (kc == k_inner) { kc = 0; ++tap; }
        if (++stage == kStages2) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN2);
      // operand descriptors of stage 0; a stage advances the 16-byte-unit address field by kStageBytes2 / 16
      const uint64_t da0 = make_sw128_kmajor_desc(smem_u32(smem));
      const uint64_t db0 = make_sw128_kmajor_desc(smem_u32(smem) + kABytes);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int unit = pair; unit < num_units; unit += num_pairs)
      for (int ui = 0; ui < n_inner; ++ui) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);  // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = da0 + static_cast<uint64_t>(stage * (kStageBytes2 >> 4));
          const uint64_t db = db0 + static_cast<uint64_t>(stage * (kStageBytes2 >> 4));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_pair(d_tmem, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                             (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit_pair(&empty_bar[stage]);
            if (kb == num_kb - 1) umma_commit_pair(&tfull_bar[as]);
          }
          __syncwarp();
          if (++stage == kStages2) { stage = 0; phase ^= 1; }
        }
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  }
  } else {
    setmaxnreg_inc<216>();
    // ------------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
    const int quad = warp & 3;
    const int half = (warp - kEpiWarp0) >> 2;
    const uint32_t stage_mine = smem_u32(stage_area + (warp - kEpiWarp0) * kEpiStageBytes);
    const bool rowln = p.rowln_gamma != nullptr;
    auto tile_of = [&](int unit, int ui, int* row0, int* col0) {
      const int tile = unit * n_inner + ui;
      *row0 = ((tile / n_tiles) * 2 + static_cast<int>(rank)) * BM;
      *col0 = (tile % n_tiles) * BN2;
    };
    GemmEpiPrefetch pf_next{};
    if (pair < num_units) {
      int r0, c0;
      tile_of(pair, 0, &r0, &c0);
      pf_next = gemm_epi_prefetch<BN2>(p, r0, p.M - r0, c0, BN2, quad, half, lane);
    }
    const uint32_t leader_tempty0 = mapa_shared(smem_u32(&tempty_bar[0]), 0);
    int as = 0;
    uint32_t aphase = 0;
    int upar = 0;  // parity of the unit: double-buffers the row statistics of the fused LayerNorm
    for (int unit = pair; unit < num_units; unit += num_pairs) {
      for (int ui = 0; ui < n_inner; ++ui) {
        int row0, col0;
        tile_of(unit, ui, &row0, &col0);
        const int valid = p.M - row0;
        const GemmEpiPrefetch pf_cur = pf_next;
        {
          int nu = unit, ni = ui + 1;
          if (ni == n_inner) { ni = 0; nu += num_pairs; }
          if (nu < num_units) {
            int nrow0, ncol0;
            tile_of(nu, ni, &nrow0, &ncol0);
            pf_next = gemm_epi_prefetch<BN2>(p, nrow0, p.M - nrow0, ncol0, BN2, quad, half, lane);
            if (p.resid != nullptr) gemm_prefetch_resid<BN2>(p, nrow0, p.M - nrow0, ncol0, BN2);
          }
        }
        mbar_wait(&tfull_bar[as], aphase);
        tc_fence_after();
        float2 rs = make_float2(0.f, 0.f);
        gemm_epilogue_tile<BN2, kEpi>(p, row0, valid, col0, BN2, tmem_base + static_cast<uint32_t>(as * BN2), quad, half, lane,
                                      stage_mine, pf_cur, rowln ? &rs : nullptr);
        // all of this warp's TMEM reads are complete -> release the accumulator stage to the leader's MMA thread
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(leader_tempty0 + static_cast<uint32_t>(as * 8));
        if (++as == 2) { as = 0; aphase ^= 1; }
        if (rowln)  // (sum, sum of squares) of row quad * 32 + lane over this warp's 128 columns of tile ui
          *reinterpret_cast<float2*>(stage_area + (warp - kEpiWarp0) * kEpiStageBytes + kRowLnSlot + (upar * 2 + ui) * 256 + lane * 8) = rs;
      }
      if (rowln) {
        // every epilogue warp has stored its un-normalised part of these 128 rows and published its statistics
        asm volatile("bar.sync 1, 256;" ::: "memory");
        int row0, col0;
        tile_of(unit, 0, &row0, &col0);
        rowln_finish(p, row0, p.M - row0, stage_area, upar, warp - kEpiWarp0, lane);
        upar ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody exits (or frees TMEM) while the peer can still read its smem or signal its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair<kTmemCols2>(tmem_base);
  }
}

}  // namespace

bool gemm_rowln_supported(const GemmArgs& g) {
  return gemm_pair_supported(g) && g.N == kRowLnN && g.ld_out == kRowLnN && g.out_bf16 != nullptr && g.out_f32 == nullptr &&
         g.resid == nullptr && g.resid_bf16 == nullptr && g.row_mask == nullptr && g.act == kActNone && g.n_taps == 0 &&
         g.ln_stats == nullptr && g.row_stats_out == nullptr && g.rowln_beta != nullptr && get_option_gemm_impl() != 1;
}

bool gemm_pair_supported(const GemmArgs& g) {
  return g.mode == 0 && g.N % BN2 == 0 && g.K % BK == 0 && (g.k_inner > 0 ? g.k_inner : g.K) % BK == 0 && num_sms() % 2 == 0;
}

int gemm_bf16_tc_pair(const GemmArgs& g, cudaStream_t stream) {
  if (!gemm_pair_supported(g)) return fail(kUnsupported, "gemm (cta pair): needs N % 256 == 0 and K % 64 == 0");
  if (g.ld_out % 8 != 0) return fail(kInvalidArgument, "gemm: ld_out must be a multiple of 8");
  GemmEpiParams p{};
  p.M = g.M; p.bias = g.bias; p.resid = g.resid; p.out_f32 = g.out_f32; p.out_bf16 = g.out_bf16;
  p.ld_out = g.ld_out; p.act = g.act; p.alpha = g.alpha; p.resid_bf16 = g.resid_bf16; p.row_mask = g.row_mask;
  p.row_stats_out = g.row_stats_out; p.ln_stats = g.ln_stats; p.ln_colsum = g.ln_colsum;
  p.ln_slots = g.K / 128; p.ln_inv_d = 1.0f / static_cast<float>(g.K); p.ln_eps = g.ln_eps;
  int n_inner = 1;
  if (g.rowln_gamma != nullptr) {
    if (!gemm_rowln_supported(g)) return fail(kUnsupported, "gemm: fused row LayerNorm needs N = 512 bf16 rows in the CTA-pair kernel");
    p.rowln_gamma = g.rowln_gamma; p.rowln_beta = g.rowln_beta; p.rowln_eps = g.rowln_eps; p.rowln_gelu = g.rowln_gelu;
    n_inner = g.N / BN2;
  }
  const int k_inner = g.k_inner > 0 ? g.k_inner : g.K;
  CUtensorMap tmA, tmB;
  const uint32_t abox[3] = {BK, 1, BM};
  SVT_TRY(encode_bf16_map(&tmA, g.a, 3, g.a_dims, g.a_strides, abox));
  const uint64_t wd[2] = {static_cast<uint64_t>(g.w_cols), static_cast<uint64_t>(g.w_rows)};
  const uint64_t ws[1] = {static_cast<uint64_t>(g.w_cols)};
  const uint32_t wb[2] = {BK, BN2 / 2};
  SVT_TRY(encode_bf16_map(&tmB, g.w, 2, wd, ws, wb));
  static std::atomic<unsigned long long> attr_seen{0};
  if (first_use_on_device(attr_seen)) {
    SVT_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes2));
    SVT_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes2));
    SVT_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes2));
    SVT_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes2));
    SVT_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes2));
  }
  const int m_pairs = ceil_div(g.M, 2 * BM);
  const int n_tiles = g.N / BN2;
  const int units = m_pairs * n_tiles / n_inner;
  const int max_pairs = num_sms() / 2;
  const int grid = 2 * (units < max_pairs ? units : max_pairs);
  int epi = gemm_epi_mode(p, g.N);
  const int opt = get_option_resid_epilogue();  // 0: generic kernel only, 1: residual variant only, 2 (default): all variants
  if (opt == 0 || (opt == 1 && epi != 1)) epi = 0;
  switch (epi) {
    case 1: gemm_tc2_kernel<1><<<grid, kThreads, kSmemBytes2, stream>>>(tmA, tmB, p, k_inner, g.K / BK, m_pairs, n_tiles, n_inner); break;
    case 2: gemm_tc2_kernel<2><<<grid, kThreads, kSmemBytes2, stream>>>(tmA, tmB, p, k_inner, g.K / BK, m_pairs, n_tiles, n_inner); break;
    case 3: gemm_tc2_kernel<3><<<grid, kThreads, kSmemBytes2, stream>>>(tmA, tmB, p, k_inner, g.K / BK, m_pairs, n_tiles, n_inner); break;
    case 4: gemm_tc2_kernel<4><<<grid, kThreads, kSmemBytes2, stream>>>(tmA, tmB, p, k_inner, g.K / BK, m_pairs, n_tiles, n_inner); break;
    default: gemm_tc2_kernel<0><<<grid, kThreads, kSmemBytes2, stream>>>(tmA, tmB, p, k_inner, g.K / BK, m_pairs, n_tiles, n_inner);
  }
  SVT_POST_LAUNCH();
  return kOk;
}

}  // namespace svt
