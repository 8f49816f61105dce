// Microbenchmark: MUFU exp2 throughput, fp32 vs packed f16x2 (decides the attention softmax formulation).
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__global__ void k_f32(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f;
  for (int i = 0; i < iters; ++i) {
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0));
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2));
    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
    a0 -= 1.0f; a1 -= 1.0f; a2 -= 1.0f; a3 -= 1.0f;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}
__global__ void k_f16x2(float* out, int iters) {
  unsigned a0 = 0x30003000u + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
  const unsigned one = 0xbc00bc00u;  // (-1, -1)
  for (int i = 0; i < iters; ++i) {
    asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a0));
    asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a1));
    asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a2));
    asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a3));
    asm volatile("add.f16x2 %0, %0, %1;" : "+r"(a0) : "r"(one));
    asm volatile("add.f16x2 %0, %0, %1;" : "+r"(a1) : "r"(one));
    asm volatile("add.f16x2 %0, %0, %1;" : "+r"(a2) : "r"(one));
    asm volatile("add.f16x2 %0, %0, %1;" : "+r"(a3) : "r"(one));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(a0 ^ a1 ^ a2 ^ a3);
}
int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  cudaEvent_t s, e;
  cudaEventCreate(&s); cudaEventCreate(&e);
  const int iters = 4096, blocks = 148 * 4, threads = 512;
  for (int which = 0; which < 2; ++which) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(s);
      if (which == 0) k_f32<<<blocks, threads>>>(out, iters); else k_f16x2<<<blocks, threads>>>(out, iters);
      cudaEventRecord(e); cudaEventSynchronize(e);
      float ms; cudaEventElapsedTime(&ms, s, e);
      double ops = double(blocks) * threads * iters * 4;            // MUFU instructions (per thread)
      double vals = ops * (which == 0 ? 1 : 2);
      if (rep == 1) printf("%s: %.3f ms  %.1f G MUFU-instr-lanes/s  %.1f G exp-values/s  (%.2f values/clk/SM @1.9GHz)\n",
                           which == 0 ? "ex2.f32  " : "ex2.f16x2", ms, ops / ms / 1e6, vals / ms / 1e6, vals / (ms * 1e-3) / 148 / 1.9e9);
    }
  }
  return 0;
}
