// Standalone probe of the tcgen05 operand scheme the convolution kernels rely on:
//   * SWIZZLE_NONE K-major operands whose rows sit at a 16-byte pitch (SBO = 128 B), so a
//     "shifted view" of an activation plane is just a different descriptor start address;
//   * LBO as a free byte distance between the two 8-element K halves of one MMA (lets a K step
//     pair channel group 2 of one tap with channel group 0 of the next);
//   * N in {16, 48, 80, 96}, accumulation over several K steps, tcgen05.ld 32x32b read-back.
// Prints max |D - D_ref| per case; exit code 0 iff all cases match.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../umma.cuh"

struct Step { uint32_t a_off, a_lbo, b_off, b_lbo; };

__global__ void __launch_bounds__(128) probe_kernel(const uint4* __restrict__ a_img, int a_bytes,
                                                    const uint4* __restrict__ b_img, int b_bytes,
                                                    const Step* __restrict__ steps, int nsteps, int N,
                                                    int fmt, float* __restrict__ d_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((a_bytes + 1023) / 1024) * 1024;
  for (int i = threadIdx.x; i < a_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(sa)[i] = a_img[i];
  for (int i = threadIdx.x; i < b_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(sb)[i] = b_img[i];
  if (threadIdx.x == 0) umma::mbar_init(&bar, 1);
  umma::fence_proxy_async_smem();
  umma::fence_barrier_init();
  if (threadIdx.x < 32) umma::tmem_alloc(&tmem_base_s, 128);
  umma::tcgen05_fence_before();
  __syncthreads();
  umma::tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma::make_idesc(128, N, fmt);
    for (int s = 0; s < nsteps; ++s) {
      const Step st = steps[s];
      const uint64_t da = umma::make_desc(umma::smem_u32(sa) + st.a_off, st.a_lbo, 128);
      const uint64_t db = umma::make_desc(umma::smem_u32(sb) + st.b_off, st.b_lbo, 128);
      umma::mma_f16(tmem, da, db, idesc, s > 0);
    }
    umma::commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::tcgen05_fence_after();
  const int warp = threadIdx.x >> 5;
  const int row = threadIdx.x;
  for (int c = 0; c < N; c += 8) {
    uint32_t v[8];
    umma::tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    umma::tmem_ld_wait();
    for (int k = 0; k < 8; ++k) d_out[row * N + c + k] = __uint_as_float(v[k]);
  }
  umma::tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) umma::tmem_dealloc(tmem, 128);
}

static float bf(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
static uint16_t tobf(float f) { __nv_bfloat16 b = __float2bfloat16(f); uint16_t h; memcpy(&h, &b, 2); return h; }

// case: activation planes [G][S][8] with row r of tap `off` at slot r+off; weights [khalf][N][8] per K step
static int run_case(int N, int G, int S, const std::vector<int>& tap_offs) {
  const int M = 128;
  std::vector<uint16_t> act((size_t)G * S * 8);
  srand(1234 + N);
  for (auto& v : act) v = tobf((rand() % 2001 - 1000) / 500.0f);
  // K groups: for every tap, G channel groups; flattened list, paired into K steps (last padded)
  struct Grp { int tap, g; };
  std::vector<Grp> groups;
  for (int g = 0; g < G; ++g)   // group-major: operand addresses increase, so every LBO is >= 0
    for (size_t t = 0; t < tap_offs.size(); ++t) groups.push_back({(int)t, g});
  const int nsteps = (int)(groups.size() + 1) / 2;
  std::vector<uint16_t> wt((size_t)nsteps * 2 * N * 8, 0);
  std::vector<float> wref((size_t)groups.size() * N * 8);
  for (size_t gi = 0; gi < groups.size(); ++gi)
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < 8; ++k) {
        uint16_t h = tobf((rand() % 2001 - 1000) / 4000.0f);
        wt[(((size_t)(gi / 2) * 2 + (gi & 1)) * N + n) * 8 + k] = h;
        wref[(gi * N + n) * 8 + k] = bf(h);
      }
  std::vector<Step> steps(nsteps);
  auto a_addr = [&](const Grp& g) { return (uint32_t)(((size_t)g.g * S + tap_offs[g.tap]) * 16); };
  for (int s = 0; s < nsteps; ++s) {
    const Grp g0 = groups[2 * s];
    const bool has1 = (size_t)(2 * s + 1) < groups.size();
    const Grp g1 = has1 ? groups[2 * s + 1] : g0;   // padded half: any valid plane, zero weights
    uint32_t a0 = a_addr(g0), a1 = a_addr(g1);
    if (a1 < a0) { fprintf(stderr, "negative LBO in test construction\n"); return 1; }
    steps[s] = {a0, a1 - a0, (uint32_t)((size_t)s * 2 * N * 16), (uint32_t)(N * 16)};
  }
  std::vector<float> ref((size_t)M * N, 0.f);
  for (size_t gi = 0; gi < groups.size(); ++gi)
    for (int r = 0; r < M; ++r)
      for (int n = 0; n < N; ++n) {
        float acc = 0;
        for (int k = 0; k < 8; ++k)
          acc += bf(act[((size_t)groups[gi].g * S + r + tap_offs[groups[gi].tap]) * 8 + k]) * wref[(gi * N + n) * 8 + k];
        ref[(size_t)r * N + n] += acc;
      }
  void *da, *db, *ds; float* dd;
  cudaMalloc(&da, act.size() * 2); cudaMalloc(&db, wt.size() * 2); cudaMalloc(&ds, steps.size() * sizeof(Step));
  cudaMalloc(&dd, (size_t)M * N * 4);
  cudaMemcpy(da, act.data(), act.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, wt.data(), wt.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(ds, steps.data(), steps.size() * sizeof(Step), cudaMemcpyHostToDevice);
  const int a_bytes = (int)act.size() * 2, b_bytes = (int)wt.size() * 2;
  const int smem = ((a_bytes + 1023) / 1024) * 1024 + b_bytes + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 128, smem>>>((const uint4*)da, a_bytes, (const uint4*)db, b_bytes, (const Step*)ds, nsteps, N, 1, dd);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("case N=%d: CUDA error %s\n", N, cudaGetErrorString(e)); return 1; }
  std::vector<float> out((size_t)M * N);
  cudaMemcpy(out.data(), dd, out.size() * 4, cudaMemcpyDeviceToHost);
  double mx = 0;
  for (size_t i = 0; i < out.size(); ++i) mx = fmax(mx, fabs((double)out[i] - ref[i]));
  printf("case N=%3d G=%d taps=%zu ksteps=%d smem=%d: max|D-ref| = %.3e %s\n", N, G, tap_offs.size(), nsteps, smem, mx,
         mx < 1e-3 ? "OK" : "MISMATCH");
  cudaFree(da); cudaFree(db); cudaFree(ds); cudaFree(dd);
  return mx < 1e-3 ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run_case(32, 2, 128, {0});                       // one plain K step, aligned
  bad += run_case(32, 2, 160, {0, 1, 7, 19});             // shifted (unaligned) start addresses
  bad += run_case(96, 3, 260, {0, 1, 2, 51, 52, 53, 102, 103, 104});   // 3^3-conv-like, 27 groups -> 14 K steps
  bad += run_case(80, 3, 260, {0, 1, 2, 51, 52, 53, 102, 103, 104});
  bad += run_case(48, 3, 200, {0, 1, 25, 26});
  bad += run_case(16, 3, 300, {0, 1, 2, 3, 51, 52, 53, 54});
  printf(bad ? "PROBE FAILED (%d)\n" : "PROBE PASSED\n", bad);
  return bad;
}
