// Micro-benchmark: issue cost of the exact-erf GELU (common.cuh gelu_erf, scalar FFMA) vs a packed fma.rn.f32x2
// formulation, and raw fma.rn.f32 vs fma.rn.f32x2 issue rates, on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gelu_bench gelu_bench.cu && ./gelu_bench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float gelu_scalar(float x) {
  const float ax = fabsf(x);
  float p = fmaf(5.38297490493278e-06f, ax, 4.889063711743802e-05f);
  p = fmaf(p, ax, 3.8003574445610866e-05f);
  p = fmaf(p, ax, 0.0032776263542473316f);
  p = fmaf(p, ax, 0.02114100567996502f);
  p = fmaf(p, ax, 0.04986734688282013f);
  p = fmaf(p, ax, 1.0f);
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p));
  r *= r; r *= r; r *= r; r *= r;
  return fmaf(-0.5f * ax, r, fmaxf(x, 0.f));
}

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// two GELUs at once
__device__ __forceinline__ void gelu_packed(float& x0, float& x1) {
  const u64 ax = pk(fabsf(x0), fabsf(x1));
  u64 p = fma2(pk(5.38297490493278e-06f, 5.38297490493278e-06f), ax, pk(4.889063711743802e-05f, 4.889063711743802e-05f));
  p = fma2(p, ax, pk(3.8003574445610866e-05f, 3.8003574445610866e-05f));
  p = fma2(p, ax, pk(0.0032776263542473316f, 0.0032776263542473316f));
  p = fma2(p, ax, pk(0.02114100567996502f, 0.02114100567996502f));
  p = fma2(p, ax, pk(0.04986734688282013f, 0.04986734688282013f));
  p = fma2(p, ax, pk(1.0f, 1.0f));
  float p0, p1;
  upk(p, p0, p1);
  asm("rcp.approx.ftz.f32 %0, %0;" : "+f"(p0));
  asm("rcp.approx.ftz.f32 %0, %0;" : "+f"(p1));
  u64 r = pk(p0, p1);
  r = mul2(r, r); r = mul2(r, r); r = mul2(r, r); r = mul2(r, r);
  const u64 h = mul2(ax, pk(-0.5f, -0.5f));
  const u64 y = fma2(h, r, pk(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
  upk(y, x0, x1);
}

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = -0.001f * (threadIdx.x + i) + 0.3f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = gelu_scalar(x[i]) + 0.25f;
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) { gelu_packed(x[i], x[i + 1]); x[i] += 0.25f; x[i + 1] += 0.25f; }
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i]));
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        u64 v = pk(x[i], x[i + 1]);
        asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(v));
        asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(v));
        upk(v, x[i], x[i + 1]);
      }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, double elems_per_iter) {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  for (int warps : {4, 8, 16}) {
    k<MODE><<<148, warps * 32>>>(out, cyc, iters);
    k<MODE><<<148, warps * 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-34s warps/SMSP=%d  SMSP cycles per element-warp = %6.2f\n", name, warps / 4,
           double(c) / (iters * elems_per_iter) / (warps / 4));
  }
}

int main() {
  run<0>("gelu scalar (16 per iter)", 16);
  run<1>("gelu packed f32x2 (16 per iter)", 16);
  run<2>("fma.rn.f32 (16 per iter)", 16);
  run<3>("fma.rn.f32x2 (16 elem-fma per iter)", 16);
  return 0;
}
