// Microbenchmark of the handshake latencies the plane-sweep kernels depend on:
//   (a) tcgen05.commit -> mbarrier arrival seen by the issuing thread (0 and 14 MMAs in flight)
//   (b) mbarrier ping-pong between two warps (arrive -> try_wait wake-up), with try_wait and test_wait
//   (c) commit -> wake-up of ANOTHER warp
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../umma.cuh"

__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(umma::smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

__global__ void __launch_bounds__(128) sync_kernel(int reps, int nmma, int mode, long long* out, int a_shift = 0, int k_stride = 32) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t barA, barB;
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < 32 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { umma::mbar_init(&barA, 1); umma::mbar_init(&barB, 1); }
  umma::fence_proxy_async_smem();
  umma::fence_barrier_init();
  if (threadIdx.x < 32) umma::tmem_alloc(&tmem_base_s, 512);
  umma::tcgen05_fence_before();
  __syncthreads();
  umma::tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = umma::make_idesc(128, 32, 1);
  const uint64_t da = umma::make_desc(umma::smem_u32(smem), 2048, 128);
  const uint64_t db = umma::make_desc(umma::smem_u32(smem) + 16384, 32 * 16, 128);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (mode == 0) {          // (a) commit latency seen by the issuer
    if (threadIdx.x == 0) {
      long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        for (int k = 0; k < nmma; ++k) umma::mma_f16(tmem, da, db, idesc, k > 0);
        umma::commit(&barA);
        umma::mbar_wait(&barA, r & 1);
      }
      out[0] = clock64() - t0;
    }
  } else if (mode == 1 || mode == 2) {   // (b) ping-pong warp0 <-> warp1 (1: try_wait, 2: test_wait)
    if (lane == 0 && warp < 2) {
      long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        if (warp == 0) {
          umma::mbar_arrive(&barA);
          if (mode == 1) umma::mbar_wait(&barB, r & 1); else while (!mbar_test_wait(&barB, r & 1)) {}
        } else {
          if (mode == 1) umma::mbar_wait(&barA, r & 1); else while (!mbar_test_wait(&barA, r & 1)) {}
          umma::mbar_arrive(&barB);
        }
      }
      if (warp == 0) out[0] = clock64() - t0;
    }
  } else if (mode == 3) {   // (c) warp0 commits (nmma MMAs), warp1 waits then arrives barB, warp0 waits barB
    if (lane == 0 && warp < 2) {
      long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        if (warp == 0) {
          for (int k = 0; k < nmma; ++k) umma::mma_f16(tmem, da, db, idesc, k > 0);
          umma::commit(&barA);
          umma::mbar_wait(&barB, r & 1);
        } else {
          umma::mbar_wait(&barA, r & 1);
          umma::mbar_arrive(&barB);
        }
      }
      if (warp == 0) out[0] = clock64() - t0;
    }
  }
  else if (mode == 6 || mode == 7 || mode == 8) {
    // issue-side cost only: 6 = {14 MMAs, commit} no waits; 7 = {try_wait on a completed barrier, 14 MMAs};
    // 8 = {14 MMAs} then a commit every 4th step
    if (threadIdx.x == 0) {
      umma::mbar_arrive(&barB);            // complete phase 0 of barB once
      long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        if (mode == 7) umma::mbar_wait(&barB, 0);
        for (int k = 0; k < nmma; ++k) umma::mma_f16(tmem + (r & 1) * 40, da, db, idesc, k > 0);
        if (mode == 6) umma::commit(&barA);
        if (mode == 8 && (r & 3) == 3) umma::commit(&barA);
      }
      umma::commit(&barA);
      long long t1 = clock64();
      out[0] = t1 - t0;
      // drain
      for (int i = 0; i < 100000; ++i) { if (umma::mbar_try_wait(&barA, 0) || umma::mbar_try_wait(&barA, 1)) {} }
    }
  }
  else if (mode >= 10) {    // mode 10+W: W warps issue concurrently into their own accumulators (N=80)
    const int W = mode - 10;
    const uint32_t idesc80 = umma::make_idesc(128, 80, 1);
    const uint64_t db80 = umma::make_desc(umma::smem_u32(smem) + 16384, 80 * 16, 128);
    __syncthreads();
    if (lane == 0 && warp < W) {
      long long t0 = clock64();
      for (int r = 0; r < reps; ++r)
        for (int k = 0; k < nmma; ++k) umma::mma_f16(tmem + warp * 80, da + (uint64_t)((a_shift + k * k_stride) >> 4), db80, idesc80, k > 0);
      umma::commit(warp == 0 ? &barA : &barB);
      if (warp == 0) { umma::mbar_wait(&barA, 0); out[0] = clock64() - t0; }
    }
  }
  else if (mode >= 4) {   // double/triple-buffered MMA -> consumer -> MMA pipeline (consumer does nothing)
    __shared__ uint64_t tf[4], te[4];
    const int nb = mode - 2;          // mode 4: 2 buffers, mode 5: 3 buffers
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) { umma::mbar_init(&tf[i], 1); umma::mbar_init(&te[i], 1); } umma::fence_barrier_init(); }
    __syncthreads();
    if (warp == 0 && lane == 0) {
      long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        const int b = r % nb, ph = (r / nb) & 1;
        umma::mbar_wait(&te[b], ph ^ 1);
        umma::tcgen05_fence_after();
        for (int k = 0; k < nmma; ++k) umma::mma_f16(tmem + b * 40, da, db, idesc, k > 0);
        umma::commit(&tf[b]);
      }
      out[0] = clock64() - t0;
    } else if (warp == 1 && lane == 0) {
      for (int r = 0; r < reps; ++r) {
        const int b = r % nb, ph = (r / nb) & 1;
        umma::mbar_wait(&tf[b], ph);
        umma::tcgen05_fence_after();
        umma::tcgen05_fence_before();
        umma::mbar_arrive(&te[b]);
      }
    }
  }
  umma::tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) umma::tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(sync_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  const int reps = 2000;
  struct { int mode, nmma; const char* what; } cases[] = {
    {0, 0, "commit->own wait, 0 MMAs"}, {0, 1, "commit->own wait, 1 MMA (N=80)"}, {0, 14, "commit->own wait, 14 MMAs"},
    {1, 0, "ping-pong round trip, try_wait"}, {2, 0, "ping-pong round trip, test_wait spin"},
    {3, 0, "commit(0 MMA)->other warp->arrive->wait"}, {3, 14, "commit(14 MMA)->other warp->arrive->wait"},
    {4, 14, "pipeline 2 buffers, 14 MMAs/step"}, {5, 14, "pipeline 3 buffers, 14 MMAs/step"},
    {6, 14, "issue only: 14 MMAs + commit, no waits"}, {7, 14, "issue only: completed-wait + 14 MMAs"},
    {8, 14, "issue only: 14 MMAs, commit every 4th"}, {6, 0, "issue only: commit alone"},
    {11, 14, "1 warp issuing 14 MMAs/step (N=80)"}, {12, 14, "2 warps issuing concurrently, 14 MMAs/step each"},
    {13, 14, "3 warps issuing concurrently"}, {14, 14, "4 warps issuing concurrently"},
    {4, 0, "pipeline 2 buffers, 0 MMAs/step"}, {4, 28, "pipeline 2 buffers, 28 MMAs/step"}, {4, 6, "pipeline 2 buffers, 6 MMAs/step"}};
  for (int shift : {0, 16, 64, 112}) for (int ks : {0, 32, 128, 816}) {
    sync_kernel<<<1, 128, 48 * 1024>>>(reps, 14, 12, d, shift, ks);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("2 issuers, A start shift %3d B, per-MMA A stride %4d B: %6.1f cycles per MMA (aggregate)\n", shift, ks, (double)h / reps / 28);
  }
  for (auto& c : cases) {
    sync_kernel<<<1, 128, 48 * 1024>>>(reps, c.nmma, c.mode, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-45s %8.1f cycles/iter\n", c.what, (double)h / reps);
  }
  return 0;
}
