// HBM write-bandwidth probe: what a store-only kernel reaches on this GPU, for the access patterns of the
// convolution epilogues (one 16-byte vector per lane = 512 contiguous bytes per warp, scattered over group planes).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hbm_store_probe hbm_store_probe.cu && ./hbm_store_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

// mode 0: contiguous grid-stride 16-byte stores; 1: same with st.global.cs; 2: warp w writes 512-byte pieces that are
// `stride` bytes apart (each warp owns `pieces` consecutive pieces of a plane, planes interleaved like the epilogue)
template <int MODE>
__global__ void store_kernel(uint4* __restrict__ out, size_t n_vec, size_t piece_stride_vec, int pieces) {
  const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3u, 4u);
  if (MODE == 0 || MODE == 1) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) {
      if (MODE == 1) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(out + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      else out[i] = v;
    }
  } else {
    // the buffer is `pieces` planes of n_vec / pieces vectors; warp w, iteration k writes vector 32*(w + k*warps) + lane
    // of EVERY plane (like one voxel tile row writing its three channel-group planes of several time planes)
    const size_t plane = n_vec / pieces;
    const size_t warps = (size_t)gridDim.x * blockDim.x / 32;
    const size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
    const int lane = threadIdx.x & 31;
    for (size_t r = w; r * 32 + 31 < plane; r += warps)
      for (int pl = 0; pl < pieces; ++pl) out[(size_t)pl * plane + r * 32 + lane] = v;
    (void)piece_stride_vec;
  }
}

template <int MODE>
float run(uint4* buf, size_t n_vec, int grid, int block, int pieces) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  store_kernel<MODE><<<grid, block>>>(buf, n_vec, 0, pieces);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) store_kernel<MODE><<<grid, block>>>(buf, n_vec, 0, pieces);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main() {
  const size_t bytes = (size_t)1600 << 20;
  uint4* buf;
  cudaMalloc(&buf, bytes);
  const size_t n_vec = bytes / 16;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaMemset(buf, 0, bytes);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) cudaMemsetAsync(buf, 0, bytes);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("cudaMemset            %.3f ms  %.2f TB/s\n", ms / 5, bytes / (ms / 5) / 1e9);
  for (int grid : {148, 296, 592, 1184, 2368}) {
    for (int block : {256, 512, 1024}) {
      float a = run<0>(buf, n_vec, grid, block, 1), b = run<1>(buf, n_vec, grid, block, 1);
      float c = run<2>(buf, n_vec, grid, block, 3), d = run<2>(buf, n_vec, grid, block, 48);
      printf("grid %4d block %4d: contiguous %.2f TB/s  .cs %.2f TB/s  3 planes %.2f TB/s  48 planes %.2f TB/s\n", grid, block,
             bytes / a / 1e9, bytes / b / 1e9, bytes / c / 1e9, bytes / d / 1e9);
    }
  }
  return 0;
}
