// Fused softmax attention on the 5th-gen tensor cores (head dim 64): S = Q K^T and O = P V are tcgen05.mma with
// accumulators in TMEM, Q/K/V tiles arrive by TMA (128B swizzle), probabilities go registers -> bf16 -> swizzled
// shared memory (A operand of the second MMA); scores never touch HBM.
//
// One CTA = 256 query rows (two 128-row tiles) of one (clip, head); keys/values stream in blocks of 128.
//   warp 0      : TMA producer (Q tiles once, then K_j / V_j through 3-deep rings)
//   warp 1      : TMEM allocator + MMA issuer
//   warps 4-7   : softmax / output for query tile 0 (thread = query row, TMEM lane quadrant = warp % 4)
//   warps 8-11  : same for query tile 1.  The two tiles ping-pong on the tensor pipe: while one tile's
//                 softmax runs on the MUFU/FMA pipes, the other tile's MMAs run.
// Per key block j and tile t:   S_j = Q K_j^T (4 MMAs, N = 128)   ->   softmax: two passes over S in TMEM
// (row max, then exp2 / row sum / bf16 P to smem)   ->   O_j = P_j V_j (8 MMAs, N = 64, fresh accumulator)
// -> running output kept in REGISTERS: o = (o + O_{j-1}) * exp2(m_{j-1} - m_j), so TMEM never needs rescaling.
// V is consumed in its natural [key][d] layout as an MN-major B operand.
//
// Replaces HF Wav2Vec2Attention's core (modeling_wav2vec2.py:438-463,530-544); the 1/sqrt(d_h) scale is folded
// into the packed q-projection, there is no mask except the key-length tail.
#include "gemm_tc.cuh"
#include "ops.cuh"

namespace svt {

namespace {

constexpr int kDh = 64;
constexpr int kTileQ = 128;        // query rows per tile (UMMA M)
constexpr int kBlockK = 128;       // keys per block (UMMA N of S, K of PV)
constexpr int kRing = 3;           // K / V ring depth
constexpr int kTileBytes = 128 * kDh * 2;  // 16 KB: Q tile, K block, V block
constexpr int kPBytes = 128 * kBlockK * 2;  // 32 KB per query tile (two 64-key halves of 16 KB)
constexpr int kThreadsTc = 384;
constexpr int kSmemTc = 2 * kTileBytes + 2 * kRing * kTileBytes + 2 * kPBytes + 1024 + 256;
constexpr int kTmemColsTc = 512;   // S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384)

__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  // MN-major operand, 128-byte swizzle: 64 MN-elements contiguous per 128-B row, 8 K-rows per 1024-B atom
  // (stride byte offset between 8-row groups), leading byte offset = stride between 64-wide MN atoms.
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(kThreadsTc, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ out, int ldo, int Tq, int Tk,
                    int q_clip_rows, int q_col0, int k_col0, int v_col0) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                // 2 tiles
  uint8_t* sK = sQ + 2 * kTileBytes;                 // ring
  uint8_t* sV = sK + kRing * kTileBytes;             // ring
  uint8_t* sP = sV + kRing * kTileBytes;             // 2 tiles x 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kPBytes);
  uint64_t* q_full = bars;             // 1
  uint64_t* k_full = bars + 1;         // kRing
  uint64_t* k_empty = k_full + kRing;  // kRing
  uint64_t* v_full = k_empty + kRing;
  uint64_t* v_empty = v_full + kRing;
  uint64_t* s_full = v_empty + kRing;  // 2
  uint64_t* p_full = s_full + 2;       // 2
  uint64_t* o_full = p_full + 2;       // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * kTileQ;
  const int head = blockIdx.y;
  const int clip = blockIdx.z;
  const int n_tiles = (q0 + kTileQ < Tq) ? 2 : 1;
  const int nb = (Tk + kBlockK - 1) / kBlockK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < kRing; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 4);  // one arrive per softmax warp of the tile
      mbar_init(&o_full[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemColsTc>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, n_tiles * kTileBytes);
      for (int t = 0; t < n_tiles; ++t)
        tma_load_3d(sQ + t * kTileBytes, &tmQ, q_full, q_col0 + head * kDh, q0 + t * kTileQ, clip);
      for (int j = 0; j < nb; ++j) {
        const int s = j % kRing;
        const uint32_t ph = (j / kRing) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_expect_tx(&k_full[s], kTileBytes);
        tma_load_3d(sK + s * kTileBytes, &tmK, &k_full[s], k_col0 + head * kDh, j * kBlockK, clip);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_expect_tx(&v_full[s], kTileBytes);
        tma_load_3d(sV + s * kTileBytes, &tmV, &v_full[s], v_col0 + head * kDh, j * kBlockK, clip);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(kTileQ, kBlockK);            // 128 x 128, both K-major
      constexpr uint32_t idesc_pv = make_idesc_bf16(kTileQ, kDh) | (1u << 16);  // 128 x 64, B (= V) MN-major
      auto issue_pv = [&](int t, int jb) {  // O_t = P_t(jb) V_jb
        const uint32_t pa = smem_u32(sP + t * kPBytes);
        const uint32_t vb = smem_u32(sV + (jb % kRing) * kTileBytes);
#pragma unroll
        for (int i = 0; i < kBlockK / 16; ++i) {
          const uint64_t da = make_sw128_kmajor_desc(pa + (i >> 2) * (kPBytes / 2) + (i & 3) * 32);
          const uint64_t db = make_sw128_mnmajor_desc(vb + i * 2048, 1024);
          umma_bf16(tmem_base + 256 + t * kDh, da, db, idesc_pv, i > 0 ? 1u : 0u);
        }
        umma_commit(&o_full[t]);
      };
      mbar_wait(q_full, 0);
      for (int j = 0; j < nb; ++j) {
        const int s = j % kRing;
        mbar_wait(&k_full[s], (j / kRing) & 1);
        if (j > 0) mbar_wait(&v_full[(j - 1) % kRing], ((j - 1) / kRing) & 1);
        for (int t = 0; t < n_tiles; ++t) {
          if (j > 0) mbar_wait(&p_full[t], (j - 1) & 1);  // S_{j-1} consumed, P_{j-1} in smem, O_{j-2} consumed
          tc_fence_after();
          const uint32_t qa = smem_u32(sQ + t * kTileBytes);
          const uint32_t kb = smem_u32(sK + s * kTileBytes);
#pragma unroll
          for (int k = 0; k < kDh / 16; ++k)
            umma_bf16(tmem_base + t * kBlockK, make_sw128_kmajor_desc(qa + k * 32), make_sw128_kmajor_desc(kb + k * 32),
                      idesc_s, k > 0 ? 1u : 0u);
          umma_commit(&s_full[t]);
          if (j > 0) issue_pv(t, j - 1);
        }
        umma_commit(&k_empty[s]);
        if (j > 0) umma_commit(&v_empty[(j - 1) % kRing]);
      }
      mbar_wait(&v_full[(nb - 1) % kRing], ((nb - 1) / kRing) & 1);
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait(&p_full[t], (nb - 1) & 1);
        tc_fence_after();
        issue_pv(t, nb - 1);
      }
      umma_commit(&v_empty[(nb - 1) % kRing]);
    }
  } else if (warp >= 4 && (warp - 4) / 4 < n_tiles) {
    // ------------------------------------------------------------------ softmax + output, one thread per query row
    const int t = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;  // row inside the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_addr + static_cast<uint32_t>(t * kBlockK);
    const uint32_t o_addr = tmem_base + lane_addr + static_cast<uint32_t>(256 + t * kDh);
    uint8_t* p_row = sP + t * kPBytes + row * 128;
    constexpr float kLog2e = 1.4426950408889634f;
    float o[kDh];
#pragma unroll
    for (int i = 0; i < kDh; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;

    for (int j = 0; j < nb; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      const int key_base = j * kBlockK;
      const bool tail = (key_base + kBlockK > Tk);
      // pass 1: block row max
      float m_blk = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < kBlockK; c += 32) {
        uint32_t r[32];
        tmem_ld32(s_addr + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float v = __uint_as_float(r[i]);
          if (tail && key_base + c + i >= Tk) v = -INFINITY;
          m_blk = fmaxf(m_blk, v);
        }
      }
      const float m_new = fmaxf(m_run, m_blk);  // finite: every block holds at least one valid key
      const float corr = exp2f((m_run - m_new) * kLog2e);
      if (j > 0) {
        // O_{j-1} is complete (this also means P_{j-1} has been read, so sP may be rewritten below)
        mbar_wait(&o_full[t], (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < kDh; c += 32) {
          uint32_t r[32];
          tmem_ld32(o_addr + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c + i] = (o[c + i] + __uint_as_float(r[i])) * corr;
        }
      }
      // pass 2: probabilities -> bf16 -> swizzled smem (K-major A operand of the PV MMA), row sum
      const float mb = m_new * kLog2e;
      float l_blk = 0.f;
#pragma unroll 1
      for (int c = 0; c < kBlockK; c += 32) {
        uint32_t r[32];
        tmem_ld32(s_addr + c, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = exp2f(fmaf(__uint_as_float(r[i]), kLog2e, -mb));
          float p1 = exp2f(fmaf(__uint_as_float(r[i + 1]), kLog2e, -mb));
          if (tail) {
            if (key_base + c + i >= Tk) p0 = 0.f;
            if (key_base + c + i + 1 >= Tk) p1 = 0.f;
          }
          l_blk += p0 + p1;
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
        uint8_t* dst = p_row + (c >> 6) * (kPBytes / 2);  // 64-key half
        const int chunk0 = (c & 32) >> 3;                // 16-byte chunk index inside the 128-byte row: 0 or 4
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(dst + (((chunk0 + q) ^ (row & 7)) << 4)) =
              make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      }
      l_run = l_run * corr + l_blk;
      m_run = m_new;
      // publish: my smem writes -> async proxy, my TMEM reads are done
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
    }
    // last block's O
    mbar_wait(&o_full[t], (nb - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / l_run;
    const int qrow = q0 + t * kTileQ + row;
    __nv_bfloat16* dst = out + (static_cast<size_t>(clip) * q_clip_rows + qrow) * ldo + head * kDh;
#pragma unroll
    for (int c = 0; c < kDh; c += 32) {
      uint32_t r[32];
      tmem_ld32(o_addr + c, r);
      tmem_ld_wait();
      if (qrow < Tq) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = (o[c + 8 * q + i] + __uint_as_float(r[8 * q + i])) * inv;
          *reinterpret_cast<uint4*>(dst + c + 8 * q) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                                                pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
        }
      }
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    __syncwarp();
    tmem_dealloc<kTmemColsTc>(tmem_base);
  }
}

}  // namespace

// q/k/v: bf16 matrices with row = clip * clip_rows + t and leading dimension ld*; only rows t < T of each clip are
// visible to TMA (out-of-range rows read as zero), so clips never see each other's frames.
int attention_bf16_tc(const AttentionArgs& a, cudaStream_t stream) {
  if (a.head_dim != kDh) return fail(kUnsupported, "attention_tc: head_dim must be 64");
  if ((a.ldq | a.ldk | a.ldv | a.ldo) % 8 != 0) return fail(kInvalidArgument, "attention_tc: leading dims must be multiples of 8");
  CUtensorMap tmQ, tmK, tmV;
  const uint32_t box[3] = {kDh, 128, 1};
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(a.heads) * kDh, static_cast<uint64_t>(a.Tq), static_cast<uint64_t>(a.clips)};
    const uint64_t str[2] = {static_cast<uint64_t>(a.ldq), static_cast<uint64_t>(a.q_clip_rows) * a.ldq};
    SVT_TRY(encode_bf16_map(&tmQ, a.q, 3, dims, str, box));
  }
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(a.heads) * kDh, static_cast<uint64_t>(a.Tk), static_cast<uint64_t>(a.clips)};
    const uint64_t strk[2] = {static_cast<uint64_t>(a.ldk), static_cast<uint64_t>(a.k_clip_rows) * a.ldk};
    const uint64_t strv[2] = {static_cast<uint64_t>(a.ldv), static_cast<uint64_t>(a.k_clip_rows) * a.ldv};
    SVT_TRY(encode_bf16_map(&tmK, a.k, 3, dims, strk, box));
    SVT_TRY(encode_bf16_map(&tmV, a.v, 3, dims, strv, box));
  }
  static bool attr_set = false;
  if (!attr_set) {
    SVT_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTc));
    attr_set = true;
  }
  dim3 grid(ceil_div(a.Tq, 2 * kTileQ), a.heads, a.clips);
  attention_tc_kernel<<<grid, kThreadsTc, kSmemTc, stream>>>(tmQ, tmK, tmV, a.o, a.ldo, a.Tq, a.Tk, a.q_clip_rows, 0, 0, 0);
  SVT_POST_LAUNCH();
  return kOk;
}

}  // namespace svt
