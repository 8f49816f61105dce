// Microbenchmark: does cvt.rn.f16x2.f32 (F2FP.PACK_AB) share the MUFU (XU) pipe?  (decides the attention P-conversion path)
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f;
  unsigned acc = 0;
  for (int i = 0; i < iters; ++i) {
    if (MODE & 1) {  // 4 MUFU
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
    }
    if (MODE & 2) {  // 4 F2FP (pack 2 floats -> half2)
      unsigned r0, r1, r2, r3;
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r0) : "f"(a0), "f"(a1));
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r1) : "f"(a1), "f"(a2));
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r2) : "f"(a2), "f"(a3));
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r3) : "f"(a3), "f"(a0));
      acc ^= r0 ^ r1 ^ r2 ^ r3;
    }
    if (MODE & 4) {  // 8 FFMA for reference
      a0 = fmaf(a0, 1.0001f, -1.0f); a1 = fmaf(a1, 1.0001f, -1.0f); a2 = fmaf(a2, 1.0001f, -1.0f); a3 = fmaf(a3, 1.0001f, -1.0f);
      a0 = fmaf(a0, 0.9999f, 1.0f); a1 = fmaf(a1, 0.9999f, 1.0f); a2 = fmaf(a2, 0.9999f, 1.0f); a3 = fmaf(a3, 0.9999f, 1.0f);
    } else { a0 -= 1.0f; a1 -= 1.0f; a2 -= 1.0f; a3 -= 1.0f; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __uint_as_float(acc);
}
template <int MODE> void run(const char* name, float* out) {
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  const int iters = 4096, blocks = 148 * 4, threads = 512;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(s); k<MODE><<<blocks, threads>>>(out, iters); cudaEventRecord(e); cudaEventSynchronize(e);
    float ms; cudaEventElapsedTime(&ms, s, e);
    if (rep) printf("%-28s %.3f ms  -> %.2f clk per loop iteration per SM-subpartition-warp-slot (@1.9 GHz, 16 warps/SMSP)\n", name, ms,
                    ms * 1e-3 * 1.9e9 / iters / 16.0);
  }
}
int main() {
  float* out; cudaMalloc(&out, 148 * 4 * 512 * sizeof(float));
  run<1>("4 MUFU", out); run<2>("4 F2FP", out); run<3>("4 MUFU + 4 F2FP", out); run<4>("8 FFMA", out); run<5>("4 MUFU + 8 FFMA", out);
  return 0;
}
