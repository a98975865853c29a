// Micro-benchmark of the attention kernel's exponential section (row_probs) in isolation:
// 128-wide fp32 score row in registers -> 2^(s log2e - m) -> bf16 -> swizzled smem row.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
constexpr int kBlockK = 128;
constexpr int kPBytes = 128 * kBlockK * 2;
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
template <int VARIANT>
__device__ __forceinline__ void row_probs(const float (&sc)[kBlockK], uint32_t p_row_addr, int row, float m, int n_valid) {
  const float neg_m = -m;
#pragma unroll
  for (int c = 0; c < kBlockK; c += 32) {
    uint32_t pk[16];
    if (c < n_valid) {
      float x[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = fmaf(sc[c + i], kLog2e, neg_m);
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = ex2_approx(x[i]);
      if (c + 32 > n_valid) {
#pragma unroll
        for (int i = 0; i < 32; ++i) if (c + i >= n_valid) x[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 32; i += 2) pk[i >> 1] = pack_bf16x2(x[i], x[i + 1]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[i] = 0u;
    }
    if (VARIANT == 1) { asm volatile("" ::"r"(pk[0]), "r"(pk[5]), "r"(pk[10]), "r"(pk[15])); continue; }
    const uint32_t dst = p_row_addr + (c >> 6) * (kPBytes / 2);
    const int chunk0 = (c & 32) >> 3;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      st_shared_v4(dst + (((chunk0 + q) ^ (row & 7)) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
  }
}
// hand software-pipelined: FFMA runs 4 elements ahead of MUFU, packs run 8 elements behind, stores 16 behind
__device__ __forceinline__ void row_probs_sp(const float (&sc)[kBlockK], uint32_t p_row_addr, int row, float m) {
  const float neg_m = -m;
  float x[kBlockK];
  uint32_t pk[kBlockK / 2];
  constexpr int kA = 4, kP = 8;
#pragma unroll
  for (int i = 0; i < kA; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(x[i]) : "f"(sc[i]), "f"(kLog2e), "f"(neg_m));
#pragma unroll
  for (int i = 0; i < kBlockK + kP + 8; ++i) {
    if (i + kA < kBlockK) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(x[i + kA]) : "f"(sc[i + kA]), "f"(kLog2e), "f"(neg_m));
    if (i < kBlockK) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    const int j = i - kP;  // pack pair (j-1, j) when j is odd
    if (j >= 1 && j < kBlockK && (j & 1)) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[j >> 1]) : "f"(x[j]), "f"(x[j - 1]));
    const int q = i - kP - 8;  // store 16-byte chunk q/8 when q % 8 == 7
    if (q >= 7 && q < kBlockK && (q & 7) == 7) {
      const int ch = q >> 3;  // chunk of 8 keys
      const uint32_t dst = p_row_addr + (ch >> 3) * (kPBytes / 2) + ((((ch & 7)) ^ (row & 7)) << 4);
      st_shared_v4(dst, pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
    }
  }
}

// as row_probs_sp, but every 4th exponential is evaluated on the FMA / ALU pipes (Cody-Waite + degree-3 polynomial,
// exponent spliced in with an integer add) as an 8-stage software-pipelined chain, taking 25 % off the MUFU pipe
__device__ __forceinline__ float row_probs_poly(const float (&sc)[kBlockK], uint32_t p_row_addr, int row, float m) {
  const float neg_m = -m;
  float x[kBlockK];
  float tt[kBlockK / 4], pq[kBlockK / 4];  // magic-number sums / scratch of the polynomial elements
  uint32_t pk[kBlockK / 2];
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  constexpr int kAhead = 4, kBehind = 12;
  constexpr float kMagic = 12582912.0f;
#pragma unroll
  for (int i = 0; i < kAhead; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(x[i]) : "f"(sc[i]), "f"(kLog2e), "f"(neg_m));
#pragma unroll
  for (int i = 0; i < kBlockK + kBehind + 8; ++i) {
    if (i + kAhead < kBlockK)
      asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(x[i + kAhead]) : "f"(sc[i + kAhead]), "f"(kLog2e), "f"(neg_m));
    if (i < kBlockK && (i & 3) != 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
#pragma unroll
    for (int st = 0; st < 10; ++st) {  // stage st (ONE instruction) of the polynomial chain of element e = i - st
      const int e = i - st;
      if (e < 0 || e >= kBlockK || (e & 3) != 3) continue;
      float& t = tt[e >> 2];
      float& pp = pq[e >> 2];
      if (st == 0) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[e]) : "f"(-126.0f));
      if (st == 1) asm volatile("add.f32 %0, %1, %2;" : "=f"(t) : "f"(x[e]), "f"(kMagic));
      if (st == 2) asm volatile("sub.f32 %0, %1, %2;" : "=f"(pp) : "f"(t), "f"(kMagic));
      if (st == 3) asm volatile("sub.f32 %0, %0, %1;" : "+f"(x[e]) : "f"(pp));
      if (st == 4) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(pp) : "f"(0.05517166769240671f), "f"(x[e]), "f"(0.24261112208902874f));
      if (st == 5) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(pp) : "f"(x[e]), "f"(0.6932609857127234f));
      if (st == 6) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(x[e]) : "f"(pp), "f"(x[e]), "f"(0.999928073552223f));
      if (st == 7) asm volatile("shl.b32 %0, %1, 23;" : "=r"(reinterpret_cast<uint32_t&>(t)) : "r"(__float_as_uint(t)));
      if (st == 8) asm volatile("add.s32 %0, %1, %2;" : "=r"(reinterpret_cast<uint32_t&>(x[e])) : "r"(__float_as_uint(t)), "r"(__float_as_uint(x[e])));
    }
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
      const int ch = q >> 3;
      st_shared_v4(p_row_addr + (ch >> 3) * (kPBytes / 2) + (((ch & 7) ^ (row & 7)) << 4), pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
    }
  }
  return (s0 + s1) + (s2 + s3);
}

template <int VARIANT>
__global__ void __launch_bounds__(256, 1) k(const float* in, float* out, long long* cyc, int iters, int n_valid) {
  extern __shared__ uint8_t smem[];
  float sc[kBlockK];
  const int row = threadIdx.x & 127, t = threadIdx.x >> 7;
#pragma unroll
  for (int i = 0; i < kBlockK; ++i) sc[i] = in[(threadIdx.x * 131 + i * 7) & 4095];
  const uint32_t p_row = static_cast<uint32_t>(__cvta_generic_to_shared(smem)) + t * kPBytes + row * 128;
  __syncthreads();
  long long t0 = clock64();
  float m = 0.5f;
  for (int it = 0; it < iters; ++it) {
    if (VARIANT == 3) m += 1e-9f * row_probs_poly(sc, p_row, row, m); else if (VARIANT == 2) row_probs_sp(sc, p_row, row, m); else row_probs<VARIANT>(sc, p_row, row, m, n_valid);
    m += 0.001f;
#pragma unroll
    for (int i = 0; i < kBlockK; i += 16) sc[i] += 0.01f;  // keep the compiler from hoisting
  }
  long long t1 = clock64();
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = reinterpret_cast<float*>(smem)[threadIdx.x] + sc[5];
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int VARIANT>
void run(const char* name, int threads, int n_valid) {
  float *in, *out; long long* cyc;
  cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0, 4096 * 4); cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(k<VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kPBytes);
  const int iters = 500;
  k<VARIANT><<<148, threads, 2 * kPBytes>>>(in, out, cyc, iters, n_valid);
  k<VARIANT><<<148, threads, 2 * kPBytes>>>(in, out, cyc, iters, n_valid);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-34s threads=%d n_valid=%d: %7.1f cycles per 128-key row block (%s)\n", name, threads, n_valid, double(c) / iters, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  run<0>("row_probs (with STS)", 128, 128);
  run<0>("row_probs (with STS)", 256, 128);
  run<1>("row_probs (no STS)", 128, 128);
  run<1>("row_probs (no STS)", 256, 128);
  run<0>("row_probs tail (with STS)", 128, 115);
  run<2>("row_probs hand-pipelined", 128, 128);
  run<2>("row_probs hand-pipelined", 256, 128);
  run<3>("row_probs pipelined + 1/4 poly", 128, 128);
  run<3>("row_probs pipelined + 1/4 poly", 256, 128);
  return 0;
}
