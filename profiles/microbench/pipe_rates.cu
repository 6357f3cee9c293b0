// Micro-benchmark (B200, sm_100a): issue/pipe rates of scalar FFMA, packed FFMA2 (fma.rn.f32x2),
// MUFU.EX2 and mixes of them -- the numbers DESIGN.md's compute-bound model is based on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float& dx, float& dy, float ax, float ay, float bx, float by) {
  asm volatile("{.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rd, {%0, %1};\n\t"
               "fma.rn.f32x2 rd, ra, rb, rd;\n\tmov.b64 {%0, %1}, rd;}"
               : "+f"(dx), "+f"(dy) : "f"(ax), "f"(ay), "f"(bx), "f"(by));
}
__device__ __forceinline__ float ex2(float v) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {  // 16 scalar FFMA
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
    } else if (MODE == 1) {  // 8 FFMA2 (= 16 fma)
#pragma unroll
      for (int i = 0; i < 16; i += 2) ffma2(acc[i], acc[i + 1], a, a, b, b);
    } else if (MODE == 2) {  // 16 MUFU.EX2
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = ex2(acc[i]);
    } else if (MODE == 3) {  // 16 FFMA + 2 MUFU (fit-like mix: 8 fp per exp)
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
      acc[0] = ex2(acc[0]); acc[8] = ex2(acc[8]);
    } else if (MODE == 4) {  // 8 FFMA2 + 2 MUFU
#pragma unroll
      for (int i = 0; i < 16; i += 2) ffma2(acc[i], acc[i + 1], a, a, b, b);
      acc[0] = ex2(acc[0]); acc[8] = ex2(acc[8]);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double ops_per_iter_per_thread) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 8, threads = 256, iters = 4096;
  float* out; cudaMalloc(&out, sizeof(float) * blocks * threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, iters, 0.999f, 1e-3f);
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) k<MODE><<<blocks, threads>>>(out, iters, 0.999f, 1e-3f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  double total = ops_per_iter_per_thread * iters * (double)blocks * threads;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-28s %8.3f ms  %8.2f Gop/s  %6.2f op/clk/SM (at %d MHz)\n", name, ms, total / ms / 1e6,
         total / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
  cudaFree(out);
}

int main() {
  run<0>("16x FFMA (scalar)", 16);
  run<1>("8x FFMA2 (=16 fma)", 16);
  run<2>("16x MUFU.EX2", 16);
  run<3>("16 FFMA + 2 EX2 (count fma)", 16);
  run<4>("8 FFMA2 + 2 EX2 (count fma)", 16);
  return 0;
}
