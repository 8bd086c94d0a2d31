// Microbenchmark: cycles per tcgen05.mma (kind::f16, bf16 -> f32, K = 16) on one SM / one CTA pair, by operand
// source (A in TMEM = "TS", A in shared memory = "SS"), N and cta_group.  The similarity kernel issues
// TS MMAs with N = 64: if their rate falls short of the max(M,128)*N/(256*cta_group) floor, the A-operand
// read out of TMEM (4 KB per MMA) is the limiter and a wider N per MMA amortises it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I aurora_b200/csrc -o tools/micro/mma_rate tools/micro/mma_rate.cu
//   ./tools/micro/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace aur::ptx;

constexpr uint32_t kDescHi = 0x40004040u;  // SBO 1024 B, version 1, SWIZZLE_128B

template <int CG, int N, bool TS, bool ALT>
__global__ void __launch_bounds__(128, 1) k(long long* cyc, int iters, int rand_data) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  if (CG == 2) cluster_sync_all();
  // operands: constant (low switching activity) or pseudo-random bf16 in +-[0.5, 2) like real embeddings -- the
  // tensor pipe's power draw, and with it the clock the power cap allows, depends on the data
  uint32_t lcg = 0x9E3779B9u * (threadIdx.x + 1) + blockIdx.x;
  auto rnd = [&]() { lcg = lcg * 1664525u + 1013904223u; const uint32_t h = lcg >> 16;
                     return rand_data ? (((h & 0x80FFu) | 0x3F00u) | ((((lcg >> 8) & 0x80FFu) | 0x3F00u) << 16)) : 0x3c003c00u; };
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = rnd();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc<CG>(&tptr, 512); tmem_relinquish<CG>(); }
  fence_proxy_async_smem();
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tb = tptr;
  {   // A operand in TMEM: 128 lanes x 128 columns of bf16 pairs
    uint32_t v[16];
    for (int c = 0; c < 128; c += 16) {
      for (int j = 0; j < 16; ++j) v[j] = rnd();
      tmem_st_x16(tb + (static_cast<uint32_t>((warp & 3) * 32) << 16) + c, v);
    }
    tmem_wait_st();
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
  }
  if (warp == 0 && rank == 0) {
    const uint32_t idesc = idesc_bf16_f32(128 * CG, N);
    const uint32_t b_lo = (smem_u32(smem) & 0x3FFFFu) >> 4;
    const uint32_t a_lo = ((smem_u32(smem) + 32768u) & 0x3FFFFu) >> 4;
    const uint32_t d0 = tb + (N > 128 ? 256 : 384);   // accumulator columns 384.. like the similarity kernel
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const uint32_t d = ALT ? d0 + (it & 1) * (N <= 64 ? 64 : 0) : d0;
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            if (TS) mma_ts_bf16<CG>(d, tb + kb * 32 + k4 * 8, pack_u64(b_lo + kb * ((N / CG) * 128 >> 4) + k4 * 2, kDescHi), idesc, 1u);
            else mma_ss_bf16<CG>(d, pack_u64(a_lo + k4 * 2, kDescHi), pack_u64(b_lo + kb * ((N / CG) * 128 >> 4) + k4 * 2, kDescHi), idesc, 1u);
          }
      }
      mma_commit<CG>(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (elect_one()) cyc[blockIdx.x] = t1 - t0;
  } else if (warp == 0) {
    mbar_wait(&bar, 0);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<CG>(tb, 512);
}

template <int CG, int N, bool TS, bool ALT>
void run(int grid, int rand_data = 0, int sustain_ms = 0) {
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  cudaMemset(cyc, 0, 148 * 8);
  const int iters = sustain_ms ? 40000 : 2000;
  auto kern = k<CG, N, TS, ALT>;
  const size_t smem = 64 * 1024 + 2048;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0.f;
  // sustained mode: launch back to back for sustain_ms so the power cap settles, time the last launch
  const int reps = sustain_ms ? 1000 : 2;
  float total = 0.f;
  for (int rep = 0; rep < reps; ++rep) {
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, cyc, iters, rand_data);
    cudaEventRecord(e1);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cg%d N%d %s: %s\n", CG, N, TS ? "TS" : "SS", cudaGetErrorString(e)); return; }
    cudaEventElapsedTime(&ms, e0, e1);
    total += ms;
    if (sustain_ms && total > sustain_ms) break;
  }
  long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  double s = 0; int n = 0;
  for (int i = 0; i < grid; i += CG) { s += (double)h[i]; ++n; }
  const double per = s / n / (iters * 16.0);
  const double floor_c = 128.0 * CG * N / (256.0 * CG);
  const double flops = 2.0 * 128 * CG * N * 16 * 16.0 * iters * (grid / CG);
  printf("cta_group %d  N %3d  %s  %s  grid %3d %s: %6.1f cycles / MMA  (floor %5.1f, x%.2f)", CG, N, TS ? "TS (A in TMEM)" : "SS (A in smem)",
         ALT ? "2 accumulators" : "1 accumulator ", grid, rand_data ? "random data  " : "constant data", per, floor_c, per / floor_c);
  if (sustain_ms) printf("  sustained: %.3f ms/launch, SM clock %.0f MHz, %.0f TFLOP/s", ms, s / n / (ms * 1e3), flops / (ms * 1e-3) / 1e12);
  printf("\n");
  cudaFree(cyc);
}

int main() {
  for (int grid : {2, 148}) {
    run<2, 64, true, true>(grid);
    run<2, 64, true, false>(grid);
    run<2, 128, true, false>(grid);
    run<2, 32, true, true>(grid);
    run<2, 64, false, true>(grid);
    run<2, 128, false, false>(grid);
    run<1, 64, true, true>(grid);
    run<1, 128, true, false>(grid);
    run<1, 64, false, true>(grid);
    run<2, 256, false, false>(grid);
  }
  // under the power cap, all SMs, ~1.5 s each: what the tensor pipe sustains by tile shape and data
  for (int rd : {0, 1}) {
    run<2, 64, true, true>(148, rd, 1500);
    run<2, 128, true, false>(148, rd, 1500);
    run<2, 128, false, false>(148, rd, 1500);
    run<2, 256, false, false>(148, rd, 1500);
  }
  return 0;
}
