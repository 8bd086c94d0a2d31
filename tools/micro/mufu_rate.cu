// Microbenchmark: MUFU.EX2 / FFMA / F2FP issue rates per SM on this GPU (clock64-based).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate tools/micro/mufu_rate.cu && ./mufu_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      if (MODE == 2) { unsigned d; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(d) : "f"(a[i])); a[i] = __uint_as_float(d); }
      if (MODE == 3) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
    }
  }
  const long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int threads) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  k<MODE><<<148, threads>>>(out, cyc, iters); cudaDeviceSynchronize();
  k<MODE><<<148, threads>>>(out, cyc, iters); cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  printf("%-10s threads %4d: %.2f lane-ops/clk/SM\n", name, threads, (double)threads * 8 * iters / c);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int t : {128, 256, 512, 1024}) { run<0>("ex2", t); run<3>("rcp", t); run<2>("f2fp.bf16x2", t); run<1>("ffma", t); }
  return 0;
}
