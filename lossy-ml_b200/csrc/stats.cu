// K1: statistics of channel 0 (compress.py:64-65) and the exhaustive error-bound check.
// Two-level fixed-order reductions in fp64: per-thread strided partial -> warp shuffle tree ->
// per-block partial in the workspace -> one final block.  The grid size is a constant, so the
// summation order (and therefore the result) does not depend on the GPU or the run.
#include "common.cuh"

namespace {

constexpr int kBlocks = 148 * 8;
constexpr int kThreads = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block reduce of (a: sum, b: sum, c: max); result valid in thread 0
__device__ __forceinline__ void block_reduce3(double& a, double& b, double& c) {
  __shared__ double sa[kThreads / 32], sb[kThreads / 32], sc[kThreads / 32];
  a = warp_sum(a); b = warp_sum(b); c = warp_max(c);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { sa[wid] = a; sb[wid] = b; sc[wid] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = 0; b = 0; c = 0;
    for (int w = 0; w < kThreads / 32; ++w) { a += sa[w]; b += sb[w]; c = fmax(c, sc[w]); }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads)
stats_partial_kernel(const float* __restrict__ x, long long n, int C, bool vec, double* part) {
  double s = 0, s2 = 0, mx = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec && C == 2) {  // one float4 = two voxels (x0, c1, x0', c1')
    const float4* p = reinterpret_cast<const float4*>(x);
    const long long n2 = n / 2;
    for (long long i = tid; i < n2; i += stride) {
      const float4 v = __ldg(p + i);
      const double a = v.x, b = v.z;
      s += a; s2 += a * a; mx = fmax(mx, fabs(a));
      s += b; s2 += b * b; mx = fmax(mx, fabs(b));
    }
    if ((n & 1) && tid == 0) {
      const double a = x[(n - 1) * 2];
      s += a; s2 += a * a; mx = fmax(mx, fabs(a));
    }
  } else {
    for (long long i = tid; i < n; i += stride) {
      const double a = __ldg(x + i * C);
      s += a; s2 += a * a; mx = fmax(mx, fabs(a));
    }
  }
  block_reduce3(s, s2, mx);
  if (threadIdx.x == 0) {
    part[blockIdx.x * 3 + 0] = s;
    part[blockIdx.x * 3 + 1] = s2;
    part[blockIdx.x * 3 + 2] = mx;
  }
}

__global__ void __launch_bounds__(kThreads)
final3_kernel(const double* __restrict__ part, int nparts, double n, double* out) {
  double s = 0, s2 = 0, mx = 0;
  for (int i = threadIdx.x; i < nparts; i += kThreads) {
    s += part[i * 3 + 0]; s2 += part[i * 3 + 1]; mx = fmax(mx, part[i * 3 + 2]);
  }
  block_reduce3(s, s2, mx);
  if (threadIdx.x == 0) { out[0] = s; out[1] = s2; out[2] = mx; out[3] = n; }
}

__global__ void __launch_bounds__(kThreads)
verify_partial_kernel(const float* __restrict__ x, int C, const float* __restrict__ xhat, long long n,
                      double abs_err, double* part) {
  double viol = 0, dummy = 0, mx = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double d = fabs((double)__ldg(x + i * C) - (double)__ldg(xhat + i));
    if (!(d <= abs_err)) viol += 1.0;   // NaN counts as a violation
    mx = fmax(mx, d);
  }
  block_reduce3(viol, dummy, mx);
  if (threadIdx.x == 0) {
    part[blockIdx.x * 3 + 0] = viol;
    part[blockIdx.x * 3 + 1] = 0;
    part[blockIdx.x * 3 + 2] = mx;
  }
}

__global__ void verify_final_kernel(const double* __restrict__ part, int nparts, double* out) {
  double v = 0, d = 0, mx = 0;
  for (int i = threadIdx.x; i < nparts; i += kThreads) { v += part[i * 3]; mx = fmax(mx, part[i * 3 + 2]); }
  block_reduce3(v, d, mx);
  if (threadIdx.x == 0) { out[0] = v; out[1] = mx; }
}

}  // namespace

extern "C" size_t lc_stats_workspace_bytes(void) { return (size_t)kBlocks * 3 * sizeof(double); }

extern "C" int lc_stats(const float* d_x, int64_t n_vox, int C, double* d_out, void* d_ws, void* stream) {
  LC_REQUIRE(d_x && d_out && d_ws, "null argument");
  LC_REQUIRE(n_vox > 0 && C >= 1, "empty input");
  cudaStream_t s = as_stream(stream);
  const bool vec = (reinterpret_cast<uintptr_t>(d_x) & 15) == 0;
  stats_partial_kernel<<<kBlocks, kThreads, 0, s>>>(d_x, n_vox, C, vec, reinterpret_cast<double*>(d_ws));
  LC_LAUNCH_CHECK();
  final3_kernel<<<1, kThreads, 0, s>>>(reinterpret_cast<double*>(d_ws), kBlocks, (double)n_vox, d_out);
  LC_LAUNCH_CHECK();
  return LC_OK;
}

// Shard-invariant form: the field is cut into blocks of `block_vox` voxels (the codec uses one 16-step time chunk,
// 16*H*W) and every block is reduced on its own in the fixed order above; d_out receives {sum, sum^2, max|x|, n}
// per block.  The caller adds the blocks in index order, so a time-sharded field (every rank reduces its own
// blocks, the partials are exchanged) gets bit for bit the statistics of the single-GPU call.
extern "C" int lc_stats_blocks(const float* d_x, int64_t n_vox, int64_t block_vox, int C, double* d_out, void* d_ws,
                               void* stream) {
  LC_REQUIRE(d_x && d_out && d_ws, "null argument");
  LC_REQUIRE(n_vox > 0 && block_vox > 0 && C >= 1, "empty input");
  cudaStream_t s = as_stream(stream);
  int b = 0;
  for (int64_t o = 0; o < n_vox; o += block_vox, ++b) {
    const int64_t n = (n_vox - o < block_vox) ? (n_vox - o) : block_vox;
    const float* x = d_x + o * C;
    const bool vec = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    stats_partial_kernel<<<kBlocks, kThreads, 0, s>>>(x, n, C, vec, reinterpret_cast<double*>(d_ws));
    LC_LAUNCH_CHECK();
    final3_kernel<<<1, kThreads, 0, s>>>(reinterpret_cast<double*>(d_ws), kBlocks, (double)n, d_out + 4 * b);
    LC_LAUNCH_CHECK();
  }
  return LC_OK;
}

extern "C" int lc_verify_bound(const float* d_x, int C, const float* d_xhat, int64_t n_vox, double abs_err,
                               double* d_out, void* d_ws, void* stream) {
  LC_REQUIRE(d_x && d_xhat && d_out && d_ws, "null argument");
  LC_REQUIRE(n_vox > 0 && C >= 1, "empty input");
  cudaStream_t s = as_stream(stream);
  verify_partial_kernel<<<kBlocks, kThreads, 0, s>>>(d_x, C, d_xhat, n_vox, abs_err,
                                                     reinterpret_cast<double*>(d_ws));
  LC_LAUNCH_CHECK();
  verify_final_kernel<<<1, kThreads, 0, s>>>(reinterpret_cast<double*>(d_ws), kBlocks, d_out);
  LC_LAUNCH_CHECK();
  return LC_OK;
}
