// Micro-benchmark: MUFU.EX2 issue rate per SM sub-partition on sm_100a (f32, f16x2, bf16x2) for 1/2/4 warps per SMSP.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu && ./mufu_bench
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float x[16];
  unsigned h[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { x[i] = -0.001f * (threadIdx.x + i); h[i] = 0xB800B800u + i; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i]));
      if (MODE == 5) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(x[i]), "f"(x[(i + 1) & 15]));
      if (MODE == 6) asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(h[i]) : "r"(__float_as_uint(x[i])), "r"(h[(i + 1) & 15]));
      if (MODE == 7) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(x[(i + 1) & 15]), "f"(x[(i + 2) & 15]));
      if (MODE == 8 && (i & 1) == 0) {  // realistic pair: 2 FFMA + 2 MUFU + 1 F2FP
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[i]) : "f"(1.0001f));
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[i + 1]) : "f"(1.0001f));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i + 1]));
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(x[i]), "f"(x[i + 1]));
      }
      if (MODE == 9 && (i & 1) == 0) {  // pair with PRMT truncation pack
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[i]) : "f"(1.0001f));
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[i + 1]) : "f"(1.0001f));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i + 1]));
        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(h[i]) : "r"(__float_as_uint(x[i])), "r"(__float_as_uint(x[i + 1])));
      }
      if (MODE == 4) { asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[i]) : "f"(1.0001f)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  for (int warps : {4, 8, 16, 32}) {
    k<MODE><<<148, warps * 32>>>(out, cyc, iters);
    k<MODE><<<148, warps * 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double per_warp_inst = double(c) / (iters * 16.0);
    printf("%-28s warps/SMSP=%d  cycles per warp-instr (per warp) = %6.2f  => SMSP issue interval = %5.2f cycles\n", name, warps / 4,
           per_warp_inst, per_warp_inst / (warps / 4));
  }
}

int main() {
  run<0>("ex2.approx.ftz.f32");
  run<1>("ex2.approx.f16x2");
  run<2>("ex2.approx.ftz.bf16x2");
  run<3>("fma.rn.f32");
  run<4>("fma + ex2 f32 (2 instr)");
  run<5>("cvt.rn.bf16x2.f32 (F2FP)");
  run<6>("prmt");
  run<7>("max.f32 3-input");
  run<8>("pair: 2fma+2ex2+F2FP (x8)");
  run<9>("pair: 2fma+2ex2+PRMT (x8)");
  return 0;
}
