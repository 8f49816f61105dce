// Microbenchmark: MUFU.EX2 cadence seen by 1, 2, 3, 4 warps per SM sub-partition, alone and with the softmax loop's
// companion instructions (1 FFMA per exponential before it, 1 F2FP per two exponentials after it).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float c, float mc) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i * 0.01f;
  unsigned acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float x = a[i];
      if (MODE & 1) x = fmaf(x, c, -mc);
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(a[i]) : "f"(x));
    }
    if (MODE & 2) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        unsigned r;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[i + 1]));
        acc ^= r;
      }
    }
  }
  long long t1 = clock64();
  float s = __uint_as_float(acc);
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0) / (iters * 16.0f);
}
template <int MODE> void run(const char* name, float* out) {
  for (int warps_per_smsp = 1; warps_per_smsp <= 4; ++warps_per_smsp) {
    k<MODE><<<148, 128 * warps_per_smsp>>>(out, 2048, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    k<MODE><<<148, 128 * warps_per_smsp>>>(out, 2048, 1.0001f, 0.5f);
    float v; cudaMemcpy(&v, out, 4, cudaMemcpyDeviceToHost);
    printf("%-22s %d warp(s)/SMSP: %.2f clk per MUFU per warp -> %.2f clk per MUFU per SMSP\n", name, warps_per_smsp, v, v / warps_per_smsp);
  }
}
int main() {
  float* out; cudaMalloc(&out, 148 * 512 * sizeof(float));
  run<0>("MUFU only", out); run<1>("FFMA+MUFU", out); run<3>("FFMA+MUFU+F2FP/2", out);
  return 0;
}
