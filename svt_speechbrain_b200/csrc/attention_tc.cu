// Fused softmax attention on the 5th-gen tensor cores (head dim 64): S = Q K^T and O = P V are tcgen05.mma with
// accumulators in TMEM; Q/K/V tiles arrive by TMA (128B swizzle); probabilities go registers -> bf16 -> swizzled
// shared memory (A operand of the second MMA); scores never touch HBM.
//
// Persistent kernel, one CTA per SM, work item = 256 query rows (two 128-row tiles) of one (clip, head);
// keys / values stream in blocks of 128 through TMA rings that run ahead across work items (Q double-buffered).
//   warp 0      : TMA producer
//   warp 1      : TMEM allocator + issuer of the S = Q K^T MMAs
//   warp 2      : issuer of the O += P V MMAs
//   warps 4-7   : softmax for query tile 0 (thread = query row, TMEM lane quadrant = warp % 4)
//   warps 8-11  : softmax for query tile 1; the two tiles take strict turns on the exponential section
// Per key block j and tile t:
//   S_j = Q K_j^T (4 MMAs, N = 128) into the tile's S columns
//   softmax warps copy their S row to registers and release the S columns at once, so S_{j+1} is computed
//   while the exponentials of block j are still being evaluated; row max (FMNMX3), then a hand software-
//   pipelined loop p = 2^(s log2e - m) (FFMA four elements ahead of the MUFU.EX2, row-sum FADD / bf16 pack eight
//   behind, 16-byte stores of the swizzled P row sixteen behind), so one warp alone runs at 9.3 cycles per
//   exponential against the 8-cycle MUFU issue interval measured on B200 (tools/micro/)
//   O += P_j V_j (8 MMAs, N = 64, V consumed in its natural [key][d] layout as an MN-major B operand)
// O stays in TMEM across key blocks.  The running max m is only raised (and O rescaled in TMEM, the row sum in
// its register, by 2^(m_old - m_new)) when a block's max exceeds it by more than 8 in the log2 domain, so
// probabilities are bounded by 2^8 and the rescale is rare; the decision is warp-uniform.
// The output of an item (O / l -> bf16 -> HBM, staged through the idle P tile for coalesced rows) is written at the
// start of the NEXT item, so the last P V of an item completes in the shadow of the barrier round trip.
// Registers: the producer / issuer warpgroup shrinks to 64 registers per thread (setmaxnreg), the two softmax
// warpgroups grow to 216 to hold a 128-wide fp32 score row.
//
// Replaces HF Wav2Vec2Attention's core (modeling_wav2vec2.py:438-463,530-544); the 1/sqrt(d_h) scale is folded
// into the packed q-projection, there is no mask except the key-length tail.
#include "gemm_tc.cuh"
#include "ops.cuh"

namespace svt {

long long* attention_trace_buffer();

namespace {

constexpr int kDh = 64;
constexpr int kTileQ = 128;        // query rows per tile (UMMA M)
constexpr int kBlockK = 128;       // keys per block (UMMA N of S, K of PV)
constexpr int kRingK = 3;          // K ring depth
constexpr int kRingV = 2;          // V ring depth
constexpr int kTileBytes = 128 * kDh * 2;   // 16 KB: Q tile, K block, V block
constexpr int kPBytes = 128 * kBlockK * 2;  // 32 KB per query tile (two 64-key halves of 16 KB)
constexpr int kThreadsTc = 384;
constexpr int kSmemTc = 4 * kTileBytes + (kRingK + kRingV) * kTileBytes + 2 * kPBytes + 1024 + 256;
constexpr int kTmemColsTc = 512;   // S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384)
constexpr int kColO = 256;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units

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

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// named barriers 1 / 2: the two softmax warpgroups take strict turns on the MUFU-bound exponential section (one
// warp per SM sub-partition saturates the MUFU unit with the pipelined loop below), so that one tile's
// exponentials always overlap the other tile's TMEM loads, row max, output and barrier round trips
__device__ __forceinline__ void turn_wait(int t) { asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory"); }
__device__ __forceinline__ void turn_pass(int t) { asm volatile("bar.arrive %0, 256;" ::"r"(2 - t) : "memory"); }


// keys >= n_valid of the last block do not exist (TMA zero-filled them): set their scores to -inf once, so that the
// row max ignores them and 2^(-inf) = 0 drops them from P and from the row sum (n_valid is warp-uniform)
__device__ __forceinline__ void mask_tail(float (&sc)[kBlockK], int n_valid) {
#pragma unroll
  for (int c = 0; c < kBlockK; c += 32) {
    const int nv = n_valid - c;  // valid entries of this 32-wide chunk (<= 0: none)
    if (nv < 32) {
#pragma unroll
      for (int i = 0; i < 32; ++i) sc[c + i] = (i < nv) ? sc[c + i] : -INFINITY;
    }
  }
}

// max of a 128-wide score row held in registers
__device__ __forceinline__ float row_max(const float (&sc)[kBlockK]) {
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
  float m4 = -INFINITY, m5 = -INFINITY, m6 = -INFINITY, m7 = -INFINITY;
#pragma unroll
  for (int i = 0; i < kBlockK; i += 16) {
    m0 = max3(m0, sc[i], sc[i + 1]);
    m1 = max3(m1, sc[i + 2], sc[i + 3]);
    m2 = max3(m2, sc[i + 4], sc[i + 5]);
    m3 = max3(m3, sc[i + 6], sc[i + 7]);
    m4 = max3(m4, sc[i + 8], sc[i + 9]);
    m5 = max3(m5, sc[i + 10], sc[i + 11]);
    m6 = max3(m6, sc[i + 12], sc[i + 13]);
    m7 = max3(m7, sc[i + 14], sc[i + 15]);
  }
  return fmaxf(max3(m0, m1, m2), max3(max3(m3, m4, m5), m6, m7));
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// p = 2^(s log2e - m) -> bf16 -> this thread's row of the swizzled P tile (K-major A operand: two 64-key halves of
// [128 rows][128 B]); returns the fp32 row sum of the block; scores of -inf (masked tail keys) give 0.
// Hand software-pipelined (volatile asm pins the order): the FFMA of element i + 4, the MUFU.EX2 of element i, the
// row-sum FADD / bf16 pack of element i - 8 and the 16-byte store of the chunk ending at i - 16 are issued
// together, so no instruction waits on a result that was produced just before it.
__device__ __forceinline__ float row_probs(const float (&sc)[kBlockK], uint32_t p_row_addr, int row, float m) {
  const float neg_m = -m;
  float x[kBlockK];
  uint32_t pk[kBlockK / 2];
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  constexpr int kAhead = 4, kBehind = 8;
#pragma unroll
  for (int i = 0; i < kAhead; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(x[i]) : "f"(sc[i]), "f"(kLog2e), "f"(neg_m));
#pragma unroll
  for (int i = 0; i < kBlockK + kBehind + 8; ++i) {
    if (i + kAhead < kBlockK)
      asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(x[i + kAhead]) : "f"(sc[i + kAhead]), "f"(kLog2e), "f"(neg_m));
    if (i < kBlockK) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    const int j = i - kBehind;
    if (j >= 0 && j < kBlockK) {
      if ((j & 3) == 0) asm volatile("add.f32 %0, %0, %1;" : "+f"(s0) : "f"(x[j]));
      if ((j & 3) == 1) asm volatile("add.f32 %0, %0, %1;" : "+f"(s1) : "f"(x[j]));
      if ((j & 3) == 2) asm volatile("add.f32 %0, %0, %1;" : "+f"(s2) : "f"(x[j]));
      if ((j & 3) == 3) asm volatile("add.f32 %0, %0, %1;" : "+f"(s3) : "f"(x[j]));
      if (j & 1) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[j >> 1]) : "f"(x[j]), "f"(x[j - 1]));
    }
    const int q = i - kBehind - 8;
    if (q >= 7 && q < kBlockK && (q & 7) == 7) {
      const int ch = q >> 3;  // 16-byte chunk = 8 keys; chunk ch lives in 64-key half ch / 8
      st_shared_v4(p_row_addr + (ch >> 3) * (kPBytes / 2) + (((ch & 7) ^ (row & 7)) << 4), pk[4 * ch], pk[4 * ch + 1],
                   pk[4 * ch + 2], pk[4 * ch + 3]);
    }
  }
  return (s0 + s1) + (s2 + s3);
}

__global__ void __launch_bounds__(kThreadsTc, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ out, int ldo, int Tq, int Tk,
                    int q_clip_rows, int heads, int n_qpairs, int n_items, long long* __restrict__ trace,
                    const float* __restrict__ rel_tab, int rel_stride, const float* __restrict__ gate) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                // 2 buffers x 2 tiles
  uint8_t* sK = sQ + 4 * kTileBytes;                 // ring
  uint8_t* sV = sK + kRingK * kTileBytes;            // ring
  uint8_t* sP = sV + kRingV * kTileBytes;            // 2 tiles x 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kPBytes);
  uint64_t* q_full = bars;               // 2
  uint64_t* q_empty = q_full + 2;        // 2
  uint64_t* k_full = q_empty + 2;        // kRingK
  uint64_t* k_empty = k_full + kRingK;   // kRingK
  uint64_t* v_full = k_empty + kRingK;   // kRingV
  uint64_t* v_empty = v_full + kRingV;   // kRingV
  uint64_t* s_full = v_empty + kRingV;   // 2: S_g of tile t is in TMEM
  uint64_t* s_free = s_full + 2;         // 2: the tile's S columns have been copied to registers
  uint64_t* p_full = s_free + 2;         // 2: P_g of tile t is in smem (and any O rescale is done)
  uint64_t* o_full = p_full + 2;         // 2: O += P_g V_g has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nb = (Tk + kBlockK - 1) / kBlockK;
  const int tail_valid = Tk - (nb - 1) * kBlockK;  // valid keys of the last block (1..128)
  // development aid (svt_debug_attention_trace): clock64 stamps of CTA 0's phases, 4 recorders x 256 events
  int trace_n = 0;
  auto stamp = [&](int recorder) {
    if (trace != nullptr && blockIdx.x == 0 && lane == 0 && trace_n < 256) trace[recorder * 256 + trace_n++] = clock64();
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int s = 0; s < 2; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); }
    for (int s = 0; s < kRingK; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    for (int s = 0; s < kRingV; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 4);  // one arrive per softmax warp of the tile
      mbar_init(&p_full[t], 4);
      mbar_init(&o_full[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemColsTc>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    setmaxnreg_dec<64>();
    // Producer / issuer warps walk their (warp-uniform) loops with all lanes so that descriptors, addresses and
    // coordinates stay in uniform registers; one elected lane issues the TMA / tcgen05 instructions
    // (elect.sync picks the same leader every time, so MMAs and their commits come from one thread).
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer
      int li = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++li) {
        const int qp = item % n_qpairs;
        const int head = (item / n_qpairs) % heads;
        const int clip = item / (n_qpairs * heads);
        const int qb = li & 1;
        mbar_wait(&q_empty[qb], ((li >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&q_full[qb], 2 * kTileBytes);
          tma_load_3d(sQ + (2 * qb) * kTileBytes, &tmQ, &q_full[qb], head * kDh, qp * 2 * kTileQ, clip);
          tma_load_3d(sQ + (2 * qb + 1) * kTileBytes, &tmQ, &q_full[qb], head * kDh, qp * 2 * kTileQ + kTileQ, clip);
        }
        __syncwarp();
        for (int j = 0; j < nb; ++j) {
          const int g = li * nb + j;
          const int sk = g % kRingK, sv = g % kRingV;
          mbar_wait(&k_empty[sk], ((g / kRingK) & 1) ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&k_full[sk], kTileBytes);
            tma_load_3d(sK + sk * kTileBytes, &tmK, &k_full[sk], head * kDh, j * kBlockK, clip);
          }
          __syncwarp();
          mbar_wait(&v_empty[sv], ((g / kRingV) & 1) ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&v_full[sv], kTileBytes);
            tma_load_3d(sV + sv * kTileBytes, &tmV, &v_full[sv], head * kDh, j * kBlockK, clip);
          }
          __syncwarp();
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- S = Q K^T issuer
      constexpr uint32_t idesc_s = make_idesc_bf16(kTileQ, kBlockK);  // 128 x 128, both operands K-major
      const uint32_t q_base = smem_u32(sQ), k_base = smem_u32(sK);
      int li = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++li) {
        const int qb = li & 1;
        mbar_wait(&q_full[qb], (li >> 1) & 1);
        for (int j = 0; j < nb; ++j) {
          const int g = li * nb + j;
          const int sk = g % kRingK;
          mbar_wait(&k_full[sk], (g / kRingK) & 1);
          const uint64_t kd = make_sw128_kmajor_desc(k_base + sk * kTileBytes);
          for (int t = 0; t < 2; ++t) {
            if (g > 0) mbar_wait(&s_free[t], (g - 1) & 1);  // previous scores of this tile are in registers
            tc_fence_after();
            const uint64_t qd = make_sw128_kmajor_desc(q_base + (2 * qb + t) * kTileBytes);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < kDh / 16; ++k)  // +32 B along the swizzled K row = +2 in descriptor units
                umma_bf16(tmem_base + t * kBlockK, qd + 2 * k, kd + 2 * k, idesc_s, k > 0 ? 1u : 0u);
              umma_commit(&s_full[t]);
              if (t == 1) {
                umma_commit(&k_empty[sk]);
                if (j == nb - 1) umma_commit(&q_empty[qb]);  // last use of this item's Q tiles
              }
            }
            __syncwarp();
            stamp(2);
          }
        }
      }
    } else if (warp == 2) {
      // ---------------------------------------------------------------- O += P V issuer
      constexpr uint32_t idesc_pv = make_idesc_bf16(kTileQ, kDh) | (1u << 16);  // 128 x 64, B (= V) MN-major
      const uint32_t p_base = smem_u32(sP), v_base = smem_u32(sV);
      int li = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++li) {
        for (int j = 0; j < nb; ++j) {
          const int g = li * nb + j;
          const int sv = g % kRingV;
          mbar_wait(&v_full[sv], (g / kRingV) & 1);
          const uint64_t vd = make_sw128_mnmajor_desc(v_base + sv * kTileBytes, 1024);
          for (int t = 0; t < 2; ++t) {
            mbar_wait(&p_full[t], g & 1);
            tc_fence_after();
            stamp(3);
            const uint64_t pd = make_sw128_kmajor_desc(p_base + t * kPBytes);
            if (elect_one()) {
#pragma unroll
              for (int i = 0; i < kBlockK / 16; ++i) {
                // P: k-slice i = 64-key half (i / 4) + 32 B * (i % 4); V: 16 key rows = 2048 B per k-slice
                const uint64_t da = pd + static_cast<uint64_t>((i >> 2) * (kPBytes / 2 / 16) + (i & 3) * 2);
                const uint64_t db = vd + static_cast<uint64_t>(i * (2048 / 16));
                umma_bf16(tmem_base + kColO + t * kDh, da, db, idesc_pv, (j > 0 || i > 0) ? 1u : 0u);
              }
              umma_commit(&o_full[t]);
              if (t == 1) umma_commit(&v_empty[sv]);
            }
            __syncwarp();
            stamp(3);
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + output, one thread per query row
    setmaxnreg_inc<216>();
    const int t = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;  // row inside the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_addr + static_cast<uint32_t>(t * kBlockK);
    const uint32_t o_addr = tmem_base + lane_addr + static_cast<uint32_t>(kColO + t * kDh);
    const uint32_t p_row_addr = smem_u32(sP + t * kPBytes + row * 128);
    const bool rec = (quad == 0);
    if (t == 1) turn_pass(1);  // tile 0 takes the first turn

    // O / l of a finished item -> bf16 rows in HBM (waits for the item's last P V).  Thread = row in TMEM but HBM
    // wants lanes along a row: the warp stages its 32 x 128 B through its own rows of the P tile (idle between
    // the item's last P V and the next exponentials), then stores 4 full rows per instruction.
    auto write_output = [&](int item, int g_last, float l) {
      const int qp = item % n_qpairs;
      const int head = (item / n_qpairs) % heads;
      const int clip = item / (n_qpairs * heads);
      mbar_wait(&o_full[t], g_last & 1);
      tc_fence_after();
      const float inv = 1.0f / l;
#pragma unroll
      for (int c = 0; c < kDh; c += 16) {  // 16 columns at a time: the next item's score row is live in registers
        uint32_t r[16];
        tmem_ld16(o_addr + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[8 * q + i]) * inv;
          st_shared_v4(p_row_addr + ((((c >> 3) + q) ^ (row & 7)) << 4), pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                       pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
        }
      }
      tc_fence_before();  // ordered before this thread's next p_full arrive, i.e. before O is overwritten
      __syncwarp();
      const int q_row0 = qp * 2 * kTileQ + t * kTileQ + quad * 32;
      const uint32_t warp_rows = p_row_addr - lane * 128;
      const int ch = lane & 7;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + (lane >> 3);
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "r"(warp_rows + rr * 128 + ((ch ^ (rr & 7)) << 4)));
        if (q_row0 + rr < Tq)
          *reinterpret_cast<uint4*>(out + (static_cast<size_t>(clip) * q_clip_rows + q_row0 + rr) * ldo + head * kDh + ch * 8) = v;
      }
      __syncwarp();  // the rows are rewritten by other lanes' probabilities next
    };

    int li = 0;
    int prev_item = -1;
    float prev_l = 1.f;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++li) {
      float m_run = 0.f;  // log2 domain
      float l_run = 0.f;
      // WavLM gated relative position bias: score(i, j) += gate[i] * table[j - i]; this thread's row i is fixed per item
      float bias_gate = 0.f;
      const float* bias_row = nullptr;
      if (rel_tab != nullptr) {
        const int head = (item / n_qpairs) % heads;
        const int clip = item / (n_qpairs * heads);
        int qi = (item % n_qpairs) * 2 * kTileQ + t * kTileQ + row;
        if (qi > Tq - 1) qi = Tq - 1;  // padding rows of the last tile: any in-range address, results are discarded
        bias_gate = gate[(static_cast<size_t>(clip) * q_clip_rows + qi) * heads + head];
        bias_row = rel_tab + static_cast<size_t>(head) * rel_stride + (Tk - 1 - qi);
      }
      for (int j = 0; j < nb; ++j) {
        const int g = li * nb + j;
        if (rec) stamp(t);
        mbar_wait(&s_full[t], g & 1);
        tc_fence_after();
        if (rec) stamp(t);
        // previous item's output (its o_full wait also protects the P tile that is overwritten below); done here,
        // before the score row occupies the registers, while S_0 of the new item is already waiting in TMEM
        if (j == 0 && prev_item >= 0) write_output(prev_item, g - 1, prev_l);
        float sc[kBlockK];
        {
          uint32_t(&r)[kBlockK] = reinterpret_cast<uint32_t(&)[kBlockK]>(sc);
          tmem_ld32(s_addr, reinterpret_cast<uint32_t(&)[32]>(r[0]));
          tmem_ld32(s_addr + 32, reinterpret_cast<uint32_t(&)[32]>(r[32]));
          tmem_ld32(s_addr + 64, reinterpret_cast<uint32_t(&)[32]>(r[64]));
          tmem_ld32(s_addr + 96, reinterpret_cast<uint32_t(&)[32]>(r[96]));
          tmem_ld_wait();
        }
        // the S columns are free again: S_{g+1} of this tile runs on the tensor pipe under the exponentials below
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);
        if (rec) stamp(t);
        if (bias_row != nullptr) {
          // the table is padded with kBlockK zeros per head, so the key tail of the last block reads in range
          const float* tb = bias_row + j * kBlockK;
#pragma unroll
          for (int c = 0; c < kBlockK; ++c) sc[c] = fmaf(bias_gate, __ldg(tb + c), sc[c]);
        }
        if (j == nb - 1 && tail_valid < kBlockK) mask_tail(sc, tail_valid);
        const float m_blk = row_max(sc) * kLog2e;
        if (j == 0) {
          m_run = m_blk;
        } else {
          // PV_{g-1} reads the P tile and writes O: it must be complete before P is overwritten below
          mbar_wait(&o_full[t], (g - 1) & 1);
          if (__any_sync(0xffffffffu, m_blk > m_run + kRescaleThreshold)) {
            // rare: raise the running max and rescale O (TMEM) and the row sum
            tc_fence_after();
            const float m_new = fmaxf(m_run, m_blk);
            const float f = ex2_approx(m_run - m_new);
            m_run = m_new;
            l_run *= f;
#pragma unroll
            for (int c = 0; c < kDh; c += 32) {
              uint32_t r[32];
              tmem_ld32(o_addr + c, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
              tmem_st32(o_addr + c, r);
            }
            tmem_st_wait();
          }
        }
        if (rec) stamp(t);
        turn_wait(t);
        l_run += row_probs(sc, p_row_addr, row, m_run);
        turn_pass(t);
        if (rec) stamp(t);
        // publish: my smem writes -> async proxy; my TMEM writes (rescale) are done
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
        if (rec) stamp(t);
      }
      prev_item = item;
      prev_l = l_run;
    }
    if (prev_item >= 0) write_output(prev_item, li * nb - 1, prev_l);
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

static long long* g_attention_trace = nullptr;
long long* attention_trace_buffer() { return g_attention_trace; }
void set_attention_trace_buffer(long long* p) { g_attention_trace = p; }

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
  static std::atomic<unsigned long long> attr_seen{0};
  if (first_use_on_device(attr_seen)) SVT_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTc));
  const int n_qpairs = ceil_div(a.Tq, 2 * kTileQ);
  const int n_items = n_qpairs * a.heads * a.clips;
  const int grid = n_items < num_sms() ? n_items : num_sms();
  attention_tc_kernel<<<grid, kThreadsTc, kSmemTc, stream>>>(tmQ, tmK, tmV, a.o, a.ldo, a.Tq, a.Tk, a.q_clip_rows, a.heads,
                                                            n_qpairs, n_items, attention_trace_buffer(), a.rel_tab, a.rel_tab_stride,
                                                            a.gate);
  SVT_POST_LAUNCH();
  return kOk;
}

}  // namespace svt
