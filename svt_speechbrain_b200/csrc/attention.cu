// Fused flash-style softmax attention (no mask other than the key-length tail), bf16 in / bf16 out,
// fp32 online softmax.  One CTA = 128 query rows of one (clip, head); 8 warps x 16 rows; keys/values
// streamed in 64-row blocks through a double-buffered cp.async ring; scores never touch HBM.
//
// Replaces: HF Wav2Vec2Attention core (modeling_wav2vec2.py:438-463,530-544) and the nn.MultiheadAttention
// math path used by FusionRCA (speechbrain/nnet/attention.py:762-769).  The 1/sqrt(d_h) scale is folded
// into the packed q-projection weights, so this kernel computes softmax(q k^T) v.
//
// Round-1 note: the two contractions here use the legacy warp-level mma.sync.m16n8k16 path.  With d_h = 64
// the kernel is bounded by the exp throughput (one MUFU per 4*d_h flops), not by the tensor pipe; a
// tcgen05 (TMEM-resident S/P) version is the planned upgrade.
#include "ops.cuh"

namespace svt {

namespace {

constexpr int kBQ = 128;  // query rows per CTA
constexpr int kBK = 64;   // keys per block
constexpr int kAttnThreads = 256;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = smem_u32(smem_dst);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// smem tile [rows][DH] bf16, 16-byte chunks XOR-swizzled by (row & 7) so ldmatrix is conflict-free
template <int DH>
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
  return static_cast<uint32_t>(row * (DH * 2) + ((chunk ^ (row & 7)) << 4));
}

template <int DH, int ROWS>
__device__ __forceinline__ void load_tile(uint8_t* smem, const __nv_bfloat16* g, size_t ld, int row0, int rows_valid) {
  constexpr int kChunks = DH / 8;  // 16-byte chunks per row
  for (int i = threadIdx.x; i < ROWS * kChunks; i += kAttnThreads) {
    const int r = i / kChunks;
    const int c = i % kChunks;
    const bool ok = (row0 + r) < rows_valid;
    const __nv_bfloat16* src = g + static_cast<size_t>(ok ? (row0 + r) : 0) * ld + c * 8;
    cp_async16(smem + tile_off<DH>(r, c), src, ok);
  }
}

template <int DH>
__global__ void __launch_bounds__(kAttnThreads, (DH == 64) ? 2 : 1)
attention_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                 const __nv_bfloat16* __restrict__ v, __nv_bfloat16* __restrict__ o, int ldq, int ldk, int ldv, int ldo,
                 int Tq, int Tk, int q_clip_rows, int k_clip_rows, const float* __restrict__ rel_tab,
                 const float* __restrict__ gate, int heads, int rel_stride) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kBQ * DH * 2;
  uint8_t* sV = sK + 2 * kBK * DH * 2;

  const int q0 = blockIdx.x * kBQ;
  const int head = blockIdx.y;
  const int clip = blockIdx.z;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const __nv_bfloat16* qg = q + static_cast<size_t>(clip) * q_clip_rows * ldq + head * DH;
  const __nv_bfloat16* kg = k + static_cast<size_t>(clip) * k_clip_rows * ldk + head * DH;
  const __nv_bfloat16* vg = v + static_cast<size_t>(clip) * k_clip_rows * ldv + head * DH;

  const int n_blocks = (Tk + kBK - 1) / kBK;
  load_tile<DH, kBQ>(sQ, qg, ldq, q0, Tq);
  load_tile<DH, kBK>(sK, kg, ldk, 0, Tk);
  load_tile<DH, kBK>(sV, vg, ldv, 0, Tk);
  cp_async_commit();

  constexpr int KS = DH / 16;  // k-steps for QK^T
  constexpr int NT = DH / 8;   // output n-tiles
  uint32_t qf[KS][4];
  float oacc[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  constexpr float kLog2e = 1.4426950408889634f;

  for (int blk = 0; blk < n_blocks; ++blk) {
    const int buf = blk & 1;
    if (blk + 1 < n_blocks) {
      load_tile<DH, kBK>(sK + (buf ^ 1) * kBK * DH * 2, kg, ldk, (blk + 1) * kBK, Tk);
      load_tile<DH, kBK>(sV + (buf ^ 1) * kBK * DH * 2, vg, ldv, (blk + 1) * kBK, Tk);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (blk == 0) {
      const int mi = lane >> 3, r = lane & 7;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int row = warp * 16 + (mi & 1) * 8 + r;
        const int chunk = ks * 2 + (mi >> 1);
        ldsm_x4(smem_u32(sQ + tile_off<DH>(row, chunk)), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }
    const uint8_t* bK = sK + buf * kBK * DH * 2;
    const uint8_t* bV = sV + buf * kBK * DH * 2;

    // ---- S = Q K^T for this warp's 16 rows x 64 keys
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    {
      const int mi = lane >> 3, r = lane & 7;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {  // pairs of 8-key n-tiles
          const int krow = np * 16 + (mi >> 1) * 8 + r;
          const int chunk = ks * 2 + (mi & 1);
          uint32_t b0, b1, b2, b3;
          ldsm_x4(smem_u32(bK + tile_off<DH>(krow, chunk)), b0, b1, b2, b3);
          mma16816(s[2 * np], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b0, b1);
          mma16816(s[2 * np + 1], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b2, b3);
        }
      }
    }
    // ---- mask the key tail, online softmax (rows g and g+8 of the warp's 16)
    const int key0 = blk * kBK + (lane & 3) * 2;
    if (rel_tab != nullptr) {
      // WavLM: + gate[row] * table[key - row] (Toeplitz bias, one 2 Tk - 1 entry table per head)
      const int ra = q0 + warp * 16 + (lane >> 2), rb = ra + 8;
      const float* tb = rel_tab + static_cast<size_t>(head) * rel_stride + (Tk - 1);
      const float* gp = gate + static_cast<size_t>(clip) * q_clip_rows * heads + head;
      const float ga = ra < Tq ? gp[static_cast<size_t>(ra) * heads] : 0.f;
      const float gb = rb < Tq ? gp[static_cast<size_t>(rb) * heads] : 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int kk = key0 + nt * 8;
        if (kk < Tk) {
          if (ra < Tq) s[nt][0] = fmaf(ga, __ldg(tb + kk - ra), s[nt][0]);
          if (rb < Tq) s[nt][2] = fmaf(gb, __ldg(tb + kk - rb), s[nt][2]);
        }
        if (kk + 1 < Tk) {
          if (ra < Tq) s[nt][1] = fmaf(ga, __ldg(tb + kk + 1 - ra), s[nt][1]);
          if (rb < Tq) s[nt][3] = fmaf(gb, __ldg(tb + kk + 1 - rb), s[nt][3]);
        }
      }
    }
    if (blk == n_blocks - 1) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int kk = key0 + nt * 8;
        if (kk >= Tk) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
        if (kk + 1 >= Tk) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
      }
    }
    float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
    }
    float corr[2], rs[2] = {0.f, 0.f};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      corr[h] = exp2f((m_run[h] - mx[h]) * kLog2e);  // first block: exp2(-inf) = 0
      m_run[h] = mx[h];
    }
    uint32_t pf[4][4];  // P as bf16 A-fragments for the 4 key k-steps
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f((s[nt][0] - mx[0]) * kLog2e);
      const float p1 = exp2f((s[nt][1] - mx[0]) * kLog2e);
      const float p2 = exp2f((s[nt][2] - mx[1]) * kLog2e);
      const float p3 = exp2f((s[nt][3] - mx[1]) * kLog2e);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) l_run[h] = l_run[h] * corr[h] + rs[h];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      oacc[nt][0] *= corr[0]; oacc[nt][1] *= corr[0];
      oacc[nt][2] *= corr[1]; oacc[nt][3] *= corr[1];
    }
    // ---- O += P V
    {
      const int mi = lane >> 3, r = lane & 7;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {  // 16 keys per step
#pragma unroll
        for (int np = 0; np < NT / 2; ++np) {  // pairs of 8-wide d tiles
          const int vrow = ks * 16 + (mi & 1) * 8 + r;
          const int chunk = np * 2 + (mi >> 1);
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(smem_u32(bV + tile_off<DH>(vrow, chunk)), b0, b1, b2, b3);
          mma16816(oacc[2 * np], pf[ks][0], pf[ks][1], pf[ks][2], pf[ks][3], b0, b1);
          mma16816(oacc[2 * np + 1], pf[ks][0], pf[ks][1], pf[ks][2], pf[ks][3], b2, b3);
        }
      }
    }
    __syncthreads();  // everyone done with this K/V buffer before it is refilled
  }

  // ---- finalize: divide by the row sums, store bf16
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  const int r0 = q0 + warp * 16 + (lane >> 2);
  __nv_bfloat16* og = o + static_cast<size_t>(clip) * q_clip_rows * ldo + head * DH + (lane & 3) * 2;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    if (r0 < Tq)
      *reinterpret_cast<uint32_t*>(og + static_cast<size_t>(r0) * ldo + nt * 8) =
          pack_bf16x2(oacc[nt][0] * inv0, oacc[nt][1] * inv0);
    if (r0 + 8 < Tq)
      *reinterpret_cast<uint32_t*>(og + static_cast<size_t>(r0 + 8) * ldo + nt * 8) =
          pack_bf16x2(oacc[nt][2] * inv1, oacc[nt][3] * inv1);
  }
}

template <int DH>
int launch_attention(const AttentionArgs& a, cudaStream_t stream) {
  const int smem = (kBQ + 4 * kBK) * DH * 2;
  static std::atomic<unsigned long long> attr_seen{0};
  if (first_use_on_device(attr_seen)) SVT_CUDA(cudaFuncSetAttribute(attention_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid(ceil_div(a.Tq, kBQ), a.heads, a.clips);
  attention_kernel<DH><<<grid, kAttnThreads, smem, stream>>>(a.q, a.k, a.v, a.o, a.ldq, a.ldk, a.ldv, a.ldo, a.Tq, a.Tk,
                                                             a.q_clip_rows, a.k_clip_rows, a.rel_tab, a.gate, a.heads, a.rel_tab_stride);
  SVT_POST_LAUNCH();
  return kOk;
}

}  // namespace

int attention_bf16(const AttentionArgs& a, cudaStream_t stream) {
  if (a.Tq <= 0 || a.Tk <= 0 || a.clips <= 0 || a.heads <= 0) return fail(kInvalidArgument, "attention: empty problem");
  if ((a.ldq | a.ldk | a.ldv) % 8 != 0 || a.ldo % 2 != 0) return fail(kInvalidArgument, "attention: misaligned leading dims");
  const int impl = get_option_attention_impl();
  if (a.rel_tab != nullptr) {
    if (a.gate == nullptr || a.Tq != a.Tk || a.rel_tab_stride < 2 * a.Tk - 1 + 128)
      return fail(kInvalidArgument, "attention: relative position bias needs gates, Tq == Tk and a padded table");
    if (a.head_dim == 64 && impl != 1 && a.ldo % 8 == 0) return attention_bf16_tc(a, stream);
    if (a.head_dim == 64) return launch_attention<64>(a, stream);
    if (a.head_dim == 128) return launch_attention<128>(a, stream);
    return fail(kUnsupported, "attention: head_dim must be 64 or 128");
  }
  if (impl == 2 && a.head_dim != 64) return fail(kUnsupported, "attention: the tcgen05 kernel is built for head_dim 64");
  if (a.head_dim == 64 && impl != 1 && a.ldo % 8 == 0) return attention_bf16_tc(a, stream);
  if (a.head_dim == 64) return launch_attention<64>(a, stream);
  if (a.head_dim == 128) return launch_attention<128>(a, stream);
  return fail(kUnsupported, "attention: head_dim must be 64 or 128");
}

}  // namespace svt
