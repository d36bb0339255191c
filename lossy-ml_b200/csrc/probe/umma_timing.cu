// Microbenchmark: cycles per tcgen05.mma (M=128, K=16, kind::f16, SWIZZLE_NONE K-major operands at
// 16-byte row pitch) as a function of N, for a dependent accumulate chain in one TMEM accumulator
// and for two alternating accumulators.  One CTA per SM on `ctas` SMs.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../umma.cuh"

__global__ void __launch_bounds__(128) timing_kernel(int N, int reps, int alternate, int lbo_a, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) umma::mbar_init(&bar, 1);
  umma::fence_proxy_async_smem();
  umma::fence_barrier_init();
  if (threadIdx.x < 32) umma::tmem_alloc(&tmem_base_s, 512);
  umma::tcgen05_fence_before();
  __syncthreads();
  umma::tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma::make_idesc(128, N, 1);
    const uint32_t sa = umma::smem_u32(smem), sb = sa + 16 * 1024;
    const uint64_t da = umma::make_desc(sa, lbo_a, 128);
    const uint64_t db = umma::make_desc(sb, N * 16, 128);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint32_t d = tmem + ((alternate && (r & 1)) ? 256 : 0);
      umma::mma_f16(d, da + (uint64_t)((r & 7) * 3), db, idesc, r > 1);
    }
    umma::commit(&bar);
    umma::mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  umma::tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) umma::tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int reps = 2000;
  int Ns[] = {16, 32, 48, 80, 96, 128, 256};
  for (int lbo : {2048, 3680}) {
    for (int alt = 0; alt < 2; ++alt)
      for (int ctas : {1, 148})
        for (int N : Ns) {
          timing_kernel<<<ctas, 128, 64 * 1024>>>(N, reps, alt, lbo, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long h[148];
          cudaMemcpy(h, d, ctas * 8, cudaMemcpyDeviceToHost);
          long long mx = 0;
          for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
          printf("lbo=%d alt=%d ctas=%3d N=%3d : %.1f cycles/MMA  (floor N/2 = %d)\n", lbo, alt, ctas, N, (double)mx / reps, N / 2);
        }
  }
  return 0;
}
