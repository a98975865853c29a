// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma
// (128 x BN x 16, fp32 accumulators in TMEM, double-buffered) -> tcgen05.ld epilogue with fused
// bias / GELU / ReLU / fp32 residual and bf16 or fp32 stores.
//
// Roles (384 threads, 1 CTA per SM):
//   warp 0      : TMA producer (one elected lane)
//   warp 1      : TMEM allocator + MMA issuer (one elected lane)
//   warps 2-3   : idle (keep the epilogue warps aligned to TMEM lane quadrants: warp_id % 4)
//   warps 4-11  : epilogue; warp w owns TMEM lanes 32*(w%4).. and column half (w-4)/4 of the tile
#include "gemm_tc.cuh"
#include "gemm_epilogue.cuh"

#include <mutex>

namespace svt {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kThreads = 384;

template <int BN>
struct Cfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;  // two accumulator stages (power of two)
  static constexpr int kSmemBytes =
      kStages * kStageBytes + kNumEpiWarps * kEpiStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct KParams {
  int mode, M, N, K, k_inner;
  int n_tiles, m_tiles, num_kb;
  int tiles_per_clip, clip_rows, clip_valid, pad_left;
  int n_stride;  // output-column (and posconv input-channel) offset per n-tile
  int n_valid;   // valid output columns per n-tile (<= BN)
  int n_taps, kb_per_tap;  // shifted-row taps (0: off)
  int tap_off[kMaxGemmTaps];
  GemmEpiParams e;
};

// Epilogue of the tap-paired positional conv.  The accumulator holds D'[r][s * 64 + co] = sum over k-blocks of
// x[t0 + r + 4kb - pad, :] . W_{4kb + s}[co, :] for r < 128, s < 4, so the conv output of frame t0 + r' is
// sum_s D'[r' + s][s * 64 + co]: every 128-row tile yields kPosRows = 125 frames.  The four column blocks are added
// into a [128][64] fp32 tile in shared memory with a row shift of s (phase s; named barriers between phases because
// shifted rows cross the warps' lane quadrants), then all 8 warps read it back as whole 256-byte rows and apply
// bias + GELU + the fp32 residual.  acc tile = the 32 KB staging area; 16-byte slots XOR-swizzled by (row & 7).
constexpr int kPosRows = 125;
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ void posconv_epilogue_tile(const GemmEpiParams& p, int row0, int valid, int col0, int n_valid,
                                                      uint32_t tmem_acc, int quad, int half, int lane, uint32_t acc_addr,
                                                      uint64_t* tempty) {
  const int r = quad * 32 + lane;  // accumulator row of this thread
  const uint32_t tmem_row = tmem_acc + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
  for (int s = 0; s < 4; ++s) {
    if ((s >> 1) == half) {
      const int dst = r - s;  // output row this block contributes to
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_row + static_cast<uint32_t>(s * 64 + c), v);
        tmem_ld_wait();
        if (dst >= 0) {
          const uint32_t row_addr = acc_addr + dst * 256;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const uint32_t a = row_addr + ((((c >> 2) + q) ^ (dst & 7)) << 4);
            if (s == 0) {
              epi_sts128(a, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            } else {
              const uint4 o = epi_lds128(a);
              epi_sts128(a, __float_as_uint(__uint_as_float(o.x) + __uint_as_float(v[4 * q])),
                         __float_as_uint(__uint_as_float(o.y) + __uint_as_float(v[4 * q + 1])),
                         __float_as_uint(__uint_as_float(o.z) + __uint_as_float(v[4 * q + 2])),
                         __float_as_uint(__uint_as_float(o.w) + __uint_as_float(v[4 * q + 3])));
            }
          }
        }
      }
      if ((s & 1) == 1) {  // this warp's last TMEM read of the tile: release the accumulator stage
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty);
      }
    }
    epi_bar_sync();
  }
  // read back: warp w = 4 * half + quad owns rows 16w .. 16w + 15, two rows per instruction (16 lanes x float4 each)
  const int w = half * 4 + quad;
  const int slot = lane & 15;
  const int ncol = 4 * slot;
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.bias != nullptr && ncol < n_valid) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + ncol));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = 16 * w + 2 * i + (lane >> 4);
    const uint4 o = epi_lds128(acc_addr + rr * 256 + ((slot ^ (rr & 7)) << 4));
    if (rr < kPosRows && rr < valid && ncol < n_valid) {
      float4 a = make_float4(__uint_as_float(o.x) + b4.x, __uint_as_float(o.y) + b4.y, __uint_as_float(o.z) + b4.z,
                             __uint_as_float(o.w) + b4.w);
      if (p.act == kActGelu) { a.x = gelu_erf(a.x); a.y = gelu_erf(a.y); a.z = gelu_erf(a.z); a.w = gelu_erf(a.w); }
      const size_t off = static_cast<size_t>(row0 + rr) * static_cast<size_t>(p.ld_out) + static_cast<size_t>(col0 + ncol);
      if (p.resid != nullptr) {
        const float4 q = *reinterpret_cast<const float4*>(p.resid + off);
        a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
      }
      if (p.out_f32 != nullptr) *reinterpret_cast<float4*>(p.out_f32 + off) = a;
      if (p.out_bf16 != nullptr) *reinterpret_cast<uint2*>(p.out_bf16 + off) = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
    }
  }
  epi_bar_sync();  // the next tile's phase 0 overwrites the acc tile
}

// kEpi: 0 = every epilogue variant; 5 / 6 = only the PReLU + padding-ring-mask (+ bf16 residual) epilogue of the video
// stream's ResNet convolutions (gemm_epilogue_tile<BN, kEpi>)
template <int BN, int kEpi>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const KParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle atoms need 1024-byte aligned tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_area = smem + C::kStages * C::kStageBytes;  // 8 epilogue warps x 4 KB transpose tiles
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_area + kNumEpiWarps * kEpiStageBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tfull_bar = empty_bar + C::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  // logical warp = role index; physical warps 0-7 run the epilogue and 8-11 the producer / MMA issuer, because the
  // sub-partition scheduler favours the higher warp id: the issuer must not starve behind epilogue math
  const int warp = ((threadIdx.x >> 5) + 4) % 12;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kNumEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // register budget: the producer / issuer warpgroup needs few registers, the epilogue warpgroups get the rest
  if (warp < kEpiWarp0) {
  setmaxnreg_dec<64>();
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // The whole warp walks the (warp-uniform) loop so that addresses and coordinates live in uniform registers;
    // one elected lane arms the barrier and issues the copies.
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles;
      const int m_tile = tile / p.n_tiles;
      const int clip = m_tile / p.tiles_per_clip;
      const int tt = m_tile % p.tiles_per_clip;
      // (column, tap) coordinates of the A view advance incrementally: no integer division in the per-k-block loop, which
      // shares its sub-partition's issue slots with two epilogue warps (see the producer of gemm_tc2.cu)
      int kc = 0, tap = 0;
      const int k_wrap = (p.mode == 0 && p.n_taps > 0) ? p.kb_per_tap * BK : p.k_inner;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * C::kStageBytes;
        uint8_t* sb = sa + C::kABytes;
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], C::kStageBytes);
          if (p.mode == 0 && p.n_taps > 0) {
            // implicit 2-D conv over flat padded rows: tap = constant row shift, OOB rows read as zero
            tma_load_3d(sa, &tmA, &full_bar[stage], kc, 0, m_tile * BM + p.tap_off[tap]);
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n_tile * BN);
          } else if (p.mode == 0) {
            tma_load_3d(sa, &tmA, &full_bar[stage], kc, tap, m_tile * BM);
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n_tile * BN);
          } else if (p.mode == 1) {
            // positional conv: group = n_tile, tap = kb; frames shifted by (tap - pad_left), OOB -> 0
            tma_load_3d(sa, &tmA, &full_bar[stage], n_tile * p.n_stride, tt * BM + kb - p.pad_left, clip);
            tma_load_2d(sb, &tmB, &full_bar[stage], 0, (n_tile * p.num_kb + kb) * BN);
          } else {
            // tap-paired positional conv: k-block kb = taps 4kb .. 4kb+3 as 4 x 64 accumulator columns against ONE
            // A tile shifted by (4kb - pad_left); the tile yields kPosRows output frames (see posconv_epilogue_tile)
            tma_load_3d(sa, &tmA, &full_bar[stage], n_tile * p.n_stride, tt * kPosRows + 4 * kb - p.pad_left, clip);
            tma_load_2d(sb, &tmB, &full_bar[stage], 0, (n_tile * p.num_kb + kb) * BN);
          }
        }
        __syncwarp();
        kc += BK;
        if (kc == k_wrap) { kc = 0; ++tap; }
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform loop, one elected lane issues)
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
    // operand descriptors of stage 0; a stage advances the 16-byte-unit address field by kStageBytes / 16
    const uint64_t da0 = make_sw128_kmajor_desc(smem_u32(smem));
    const uint64_t db0 = make_sw128_kmajor_desc(smem_u32(smem) + C::kABytes);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[as], aphase ^ 1);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t da = da0 + static_cast<uint64_t>(stage * (C::kStageBytes >> 4));
        const uint64_t db = db0 + static_cast<uint64_t>(stage * (C::kStageBytes >> 4));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the swizzle row: +2 in 16-byte units
            umma_bf16(d_tmem, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                      (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (kb == p.num_kb - 1) umma_commit(&tfull_bar[as]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }
  } else {
    setmaxnreg_inc<216>();
    // ------------------------------------------------------------------ epilogue
    const int quad = warp & 3;
    const int half = (warp - kEpiWarp0) >> 2;
    const uint32_t stage_mine = smem_u32(stage_area + (warp - kEpiWarp0) * kEpiStageBytes);
    int as = 0;
    uint32_t aphase = 0;
    // row0 / valid rows of a tile's output block
    auto tile_rows = [&](int m_tile, int step, int* row0, int* valid) {
      if (p.mode == 0) {
        *row0 = m_tile * BM;
        *valid = p.M - *row0;
      } else {
        const int clip = m_tile / p.tiles_per_clip, tt = m_tile % p.tiles_per_clip;
        *row0 = clip * p.clip_rows + tt * step;
        *valid = p.clip_valid - tt * step;
      }
    };
    GemmEpiPrefetch pf_next{};
    if (blockIdx.x < num_tiles) {
      int r0, vl;
      tile_rows(blockIdx.x / p.n_tiles, BM, &r0, &vl);
      pf_next = gemm_epi_prefetch<BN>(p.e, r0, vl, (blockIdx.x % p.n_tiles) * p.n_stride, p.n_valid, quad, half, lane);
    }
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles;
      const int m_tile = tile / p.n_tiles;
      int row0, valid;
      tile_rows(m_tile, BM, &row0, &valid);
      // next tile: its bias slice goes to registers, its residual rows are pulled towards L2
      const GemmEpiPrefetch pf_cur = pf_next;
      {
        const int nt = tile + gridDim.x;
        if (nt < num_tiles) {
          const int nn = nt % p.n_tiles, nm = nt / p.n_tiles;
          int nrow0, nvalid;
          tile_rows(nm, p.mode == 2 ? kPosRows : BM, &nrow0, &nvalid);
          pf_next = gemm_epi_prefetch<BN>(p.e, nrow0, nvalid, nn * p.n_stride, p.n_valid, quad, half, lane);
          if (p.e.resid != nullptr) gemm_prefetch_resid<BN>(p.e, nrow0, nvalid, nn * p.n_stride, p.n_valid);
        }
      }
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      if (BN == 256 && p.mode == 2) {
        const int clip = m_tile / p.tiles_per_clip;
        const int tt = m_tile % p.tiles_per_clip;
        posconv_epilogue_tile(p.e, clip * p.clip_rows + tt * kPosRows, p.clip_valid - tt * kPosRows, n_tile * p.n_stride, p.n_valid,
                              tmem_base + static_cast<uint32_t>(as * BN), quad, half, lane, smem_u32(stage_area), &tempty_bar[as]);
      } else {
        gemm_epilogue_tile<BN, kEpi>(p.e, row0, valid, n_tile * p.n_stride, p.n_valid, tmem_base + static_cast<uint32_t>(as * BN), quad,
                               half, lane, stage_mine, pf_cur);
        // all of this warp's TMEM reads are complete (wait::ld above) -> release the accumulator stage
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[as]);
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

}  // namespace

int encode_bf16_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                    const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return fail(kCudaError, "cuTensorMapEncodeTiled entry point not available (no driver?)");
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_elems[i - 1] * 2;
  }
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim,
                  gstr, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    std::string s = "cuTensorMapEncodeTiled failed, code " + std::to_string(static_cast<int>(r)) + " rank " +
                    std::to_string(rank) + " dims";
    for (int i = 0; i < rank; ++i) s += " " + std::to_string(dims[i]);
    s += " strides";
    for (int i = 0; i + 1 < rank; ++i) s += " " + std::to_string(strides_elems[i]);
    return fail(kCudaError, s);
  }
  return kOk;
}

namespace {

template <int BN>
int launch(const GemmArgs& g, const KParams& kp, cudaStream_t stream) {
  using C = Cfg<BN>;
  CUtensorMap tmA, tmB;
  uint32_t abox[3];
  if (g.mode == 0) { abox[0] = BK; abox[1] = 1; abox[2] = BM; } else { abox[0] = BK; abox[1] = BM; abox[2] = 1; }
  SVT_TRY(encode_bf16_map(&tmA, g.a, 3, g.a_dims, g.a_strides, abox));
  const uint64_t wd[2] = {static_cast<uint64_t>(g.w_cols), static_cast<uint64_t>(g.w_rows)};
  const uint64_t ws[1] = {static_cast<uint64_t>(g.w_cols)};
  const uint32_t wb[2] = {BK, BN};
  SVT_TRY(encode_bf16_map(&tmB, g.w, 2, wd, ws, wb));
  static std::atomic<unsigned long long> attr_seen{0};
  if (first_use_on_device(attr_seen)) {
    SVT_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    SVT_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    SVT_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
  }
  const int tiles = kp.m_tiles * kp.n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  // the ResNet convolutions of the video stream (bias + PReLU + padding-ring mask (+ bf16 residual) -> bf16, whole tiles)
  const bool conv_epi = g.mode == 0 && g.act == kActPRelu && g.alpha != nullptr && g.bias != nullptr && g.out_bf16 != nullptr &&
                        g.out_f32 == nullptr && g.resid == nullptr && g.row_stats_out == nullptr && g.ln_stats == nullptr &&
                        g.N % BN == 0 && get_option_resid_epilogue() == 2;
  if (conv_epi && g.resid_bf16 != nullptr) gemm_tc_kernel<BN, 6><<<grid, kThreads, C::kSmemBytes, stream>>>(tmA, tmB, kp);
  else if (conv_epi) gemm_tc_kernel<BN, 5><<<grid, kThreads, C::kSmemBytes, stream>>>(tmA, tmB, kp);
  else gemm_tc_kernel<BN, 0><<<grid, kThreads, C::kSmemBytes, stream>>>(tmA, tmB, kp);
  SVT_POST_LAUNCH();
  return kOk;
}

}  // namespace

int gemm_bf16_tc(const GemmArgs& g, cudaStream_t stream) {
  if (g.ln_stats != nullptr) {
    const bool f32_path = g.out_f32 != nullptr || g.resid != nullptr || g.resid_bf16 != nullptr || g.row_mask != nullptr ||
                          g.act == kActPRelu;
    if (g.mode != 0 || f32_path || g.bias == nullptr || g.ln_colsum == nullptr || g.out_bf16 == nullptr)
      return fail(kInvalidArgument, "gemm: a folded LayerNorm needs mode 0, bf16-only output, bias and column sums");
    if (g.K % 128 != 0 || g.K / 128 > kMaxLnSlots || g.K / 128 % 2 != 0)
      return fail(kUnsupported, "gemm: a folded LayerNorm needs K in {256, 512, 768, 1024}");
  }
  if (g.row_stats_out != nullptr && (g.mode != 0 || (g.out_f32 == nullptr && g.resid_bf16 == nullptr) || g.N % 256 != 0 || g.ld_out != g.N))
    return fail(kInvalidArgument, "gemm: row statistics need mode 0, an fp32 output (or a bf16 residual) and N % 256 == 0 (= ld_out)");
  // Kernel / tile choice.  The CTA-pair kernel (256 x 256 per SM pair) wins whenever there are enough tiles to keep the
  // 74 pairs busy for a few waves; small row counts (a handful of clips) run on the one-CTA kernel with 128-column
  // tiles, which cuts the problem into 4x as many work items.  A producer of LayerNorm statistics needs 256-column
  // tiles (one 128-column slot per epilogue warp).
  const int impl = get_option_gemm_impl();
  const bool pair_ok = gemm_pair_supported(g) && g.n_taps == 0;
  bool want_small = false;
  if (g.mode == 0 && g.N % 128 == 0 && g.row_stats_out == nullptr) {
    const long long units_pair = static_cast<long long>(ceil_div(g.M, 256)) * (g.N / 256 > 0 ? g.N / 256 : 1);
    want_small = impl == 2 || (impl == 0 && units_pair < kSmallProblemPairUnits);
  }
  if (g.rowln_gamma != nullptr) return gemm_bf16_tc_pair(g, stream);  // callers check gemm_rowln_supported first
  if (impl == 3 && pair_ok) return gemm_bf16_tc_pair(g, stream);
  if (impl == 0 && pair_ok && !want_small && g.M >= 1024) return gemm_bf16_tc_pair(g, stream);
  KParams kp{};
  kp.mode = g.mode;
  kp.M = g.M;
  kp.N = g.N;
  kp.K = g.K;
  kp.k_inner = g.k_inner > 0 ? g.k_inner : g.K;
  kp.e.M = g.M;
  kp.e.bias = g.bias;
  kp.e.resid = g.resid;
  kp.e.out_f32 = g.out_f32;
  kp.e.out_bf16 = g.out_bf16;
  kp.e.ld_out = g.ld_out;
  kp.e.act = g.act;
  kp.e.alpha = g.alpha;
  kp.e.resid_bf16 = g.resid_bf16;
  kp.e.row_mask = g.row_mask;
  kp.e.row_stats_out = g.row_stats_out;
  kp.e.ln_stats = g.ln_stats; kp.e.ln_colsum = g.ln_colsum;
  kp.e.ln_slots = g.K / 128; kp.e.ln_inv_d = 1.0f / static_cast<float>(g.K); kp.e.ln_eps = g.ln_eps;
  if (g.act == kActPRelu && g.alpha == nullptr) return fail(kInvalidArgument, "gemm: PReLU needs per-column slopes");
  kp.n_taps = 0;
  if (g.n_taps > 0) {
    if (g.mode != 0 || g.n_taps > kMaxGemmTaps || kp.k_inner % BK != 0 || g.K != g.n_taps * kp.k_inner)
      return fail(kInvalidArgument, "gemm: bad tap configuration");
    kp.n_taps = g.n_taps;
    kp.kb_per_tap = kp.k_inner / BK;
    for (int i = 0; i < g.n_taps; ++i) kp.tap_off[i] = g.tap_off[i];
  }
  if (g.mode == 0 && (g.K % BK != 0 || kp.k_inner % BK != 0))
    return fail(kInvalidArgument, "gemm: K must be a multiple of 64");
  if (g.ld_out % 8 != 0) return fail(kInvalidArgument, "gemm: ld_out must be a multiple of 8");
  int bn;
  if (g.mode == 1 && g.taps % 4 == 0 && get_option_gemm_impl() != 1) {
    // tap-paired positional conv (mode 2 in the kernel): N = 256 = 4 taps x 64 output channels per k-block
    if (g.group_size <= 0 || g.group_size > 64 || g.group_size % 8 != 0 || g.N % g.group_size != 0)
      return fail(kInvalidArgument, "posconv: channels per group must be a multiple of 8 and <= 64");
    bn = 256;
    kp.mode = 2;
    kp.tiles_per_clip = ceil_div(g.clip_valid, kPosRows);
    kp.m_tiles = g.n_clips * kp.tiles_per_clip;
    kp.n_tiles = g.N / g.group_size;
    kp.n_stride = g.group_size;
    kp.n_valid = g.group_size;
    kp.num_kb = g.taps / 4;
    kp.clip_rows = g.clip_rows;
    kp.clip_valid = g.clip_valid;
    kp.pad_left = g.pad_left;
  } else if (g.mode == 1) {
    bn = 64;
    if (g.group_size <= 0 || g.group_size > 64 || g.group_size % 8 != 0 || g.N % g.group_size != 0)
      return fail(kInvalidArgument, "posconv: channels per group must be a multiple of 8 and <= 64");
    kp.tiles_per_clip = ceil_div(g.clip_valid, BM);
    kp.m_tiles = g.n_clips * kp.tiles_per_clip;
    kp.n_tiles = g.N / g.group_size;
    kp.n_stride = g.group_size;
    kp.n_valid = g.group_size;
    kp.num_kb = g.taps;
    kp.clip_rows = g.clip_rows;
    kp.clip_valid = g.clip_valid;
    kp.pad_left = g.pad_left;
  } else {
    bn = (g.N % 256 == 0 && !want_small) ? 256 : (g.N % 128 == 0 ? 128 : 64);
    if (g.N % bn != 0) return fail(kInvalidArgument, "gemm: N must be a multiple of 64");
    kp.m_tiles = ceil_div(g.M, BM);
    kp.n_tiles = g.N / bn;
    kp.num_kb = g.K / BK;
    kp.tiles_per_clip = 1;
    kp.n_stride = bn;
    kp.n_valid = bn;
  }
  if (kp.m_tiles <= 0 || kp.n_tiles <= 0 || kp.num_kb <= 0) return fail(kInvalidArgument, "gemm: empty problem");
  switch (bn) {
    case 256: return launch<256>(g, kp, stream);
    case 128: return launch<128>(g, kp, stream);
    default: return launch<64>(g, kp, stream);
  }
}

}  // namespace svt
