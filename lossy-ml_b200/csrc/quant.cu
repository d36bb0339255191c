// Residual / mask / quantise / compaction (compress.py:114-135) and its inverse
// (decompress.py:75-104), plus first-difference and cumsum helpers.
//
// Bit-exactness with NumPy: every float op is an explicit round-to-nearest intrinsic so nvcc
// cannot contract a*b+c into an FMA; the divide is IEEE (__fdiv_rn), the
// rounding is half-to-even (__float2int_rn == np.round().astype(int32) for in-range values).
#include <cuda_fp16.h>
#include "scan.cuh"

using namespace lcscan;

namespace {

constexpr int kItems = 8;
constexpr int kTile = kThreads * kItems;

__device__ __forceinline__ void load8_strided(const float* __restrict__ x, int C, long long i0,
                                              long long n, bool vec, float v[kItems]) {
  if (vec && C == 2 && i0 + kItems <= n) {
    const float4* p = reinterpret_cast<const float4*>(x + i0 * 2);
#pragma unroll
    for (int k = 0; k < kItems / 2; ++k) {
      float4 a = __ldg(p + k);
      v[2 * k] = a.x;
      v[2 * k + 1] = a.z;
    }
  } else if (vec && C == 1 && i0 + kItems <= n) {
    const float4* p = reinterpret_cast<const float4*>(x + i0);
#pragma unroll
    for (int k = 0; k < kItems / 4; ++k) {
      float4 a = __ldg(p + k);
      v[4 * k] = a.x; v[4 * k + 1] = a.y; v[4 * k + 2] = a.z; v[4 * k + 3] = a.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < kItems; ++k) v[k] = (i0 + k < n) ? __ldg(x + (i0 + k) * C) : 0.f;
  }
}

__device__ __forceinline__ int classify(float x0, float r, float mean, float stdv, float thr,
                                        float thr2, int& q) {
  const float rr = __fadd_rn(__fmul_rn(r, stdv), mean);   // compress.py:114
  const float e = __fsub_rn(x0, rr);                      // compress.py:117
  const float a = fabsf(e);
  q = 0;
  if (!(a > thr)) return 0;                               // compress.py:121 (|e| <= thr -> 0)
  q = __float2int_rn(__fdiv_rn(a, thr2));                 // compress.py:135
  return e > 0.f ? 1 : 2;                                 // compress.py:122-123
}

// decompress.py:86-104 for one element: the value the decoder adds to the de-standardised reconstruction
__device__ __forceinline__ float decoded_value(float r, float mean, float stdv, int m, int q, double thr2) {
  // val = float32(mask); val[val>0] = float64(q)*thr2 -> float32; negated where mask == 2
  float val = (float)m;
  if (m > 0) {
    val = (float)((double)q * thr2);
    if (m == 2) val = -val;
  }
  const float rr = __fadd_rn(__fmul_rn(r, stdv), mean);   // decompress.py:98
  return __fadd_rn(rr, val);                              // decompress.py:104
}

// CHECK: also decode every element exactly as reconstruct_kernel will and count |x - xhat| > abs_err in
// fp64 (out3[1]) with the largest error (out3[2], bits of a non-negative double) -- the strict mode's bound
// check without a second pass over the field.
template <bool CHECK>
__global__ void __launch_bounds__(kThreads)
quantize_kernel(const float* __restrict__ x, int C, const float* __restrict__ recon, long long n,
                float mean, float stdv, float thr, float thr2, int8_t* __restrict__ mask,
                int32_t* __restrict__ qout, unsigned long long* __restrict__ count,
                unsigned long long* ws, bool vec_x, bool vec_r, bool vec_m, double thr2_dec,
                double abs_err) {
  const int tile = take_ticket(ws);
  const long long i0 = (long long)tile * kTile + (long long)threadIdx.x * kItems;
  float xv[kItems], rv[kItems];
  load8_strided(x, C, i0, n, vec_x, xv);
  load8_strided(recon, 1, i0, n, vec_r, rv);
  int q[kItems];
  int m[kItems];
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    m[k] = (i0 + k < n) ? classify(xv[k], rv[k], mean, stdv, thr, thr2, q[k]) : 0;
    cnt += (m[k] != 0);
  }
  if (i0 + kItems <= n && vec_m) {
    unsigned long long pk = 0;
#pragma unroll
    for (int k = 0; k < kItems; ++k) pk |= (unsigned long long)(m[k] & 0xff) << (8 * k);
    *reinterpret_cast<unsigned long long*>(mask + i0) = pk;
  } else {
#pragma unroll
    for (int k = 0; k < kItems; ++k)
      if (i0 + k < n) mask[i0 + k] = (int8_t)m[k];
  }
  if (CHECK) {
    int bad = 0;
    double worst = 0.0;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
      if (i0 + k < n) {
        const float y = decoded_value(rv[k], mean, stdv, m[k], q[k], thr2_dec);
        const double d = fabs((double)xv[k] - (double)y);
        if (!(d <= abs_err)) ++bad;            // NaN counts as a violation
        worst = fmax(worst, d);
      }
    }
    __shared__ double s_worst[kThreads / 32];
    __shared__ int s_bad[kThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      bad += __shfl_xor_sync(0xffffffffu, bad, o);
      worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
    }
    if ((threadIdx.x & 31) == 0) { s_worst[threadIdx.x >> 5] = worst; s_bad[threadIdx.x >> 5] = bad; }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int w = 1; w < kThreads / 32; ++w) { worst = fmax(worst, s_worst[w]); bad += s_bad[w]; }
      if (bad) atomicAdd(count + 1, (unsigned long long)bad);
      if (worst > 0.0) atomicMax(count + 2, (unsigned long long)__double_as_longlong(worst));
    }
  }
  // compaction through shared memory: the tile's outliers are gathered in C order and written out as one
  // coalesced run (per-thread 4-byte stores would touch every 32-byte sector several times, each a partial write)
  __shared__ int32_t s_q[kTile];
  int total;
  const int texcl = block_exclusive<int>(cnt, total);
  {
    int o = texcl;
#pragma unroll
    for (int k = 0; k < kItems; ++k)
      if (m[k] != 0) s_q[o++] = q[k];
  }
  const unsigned long long base = tile_prefix(ws, tile, (unsigned long long)total);   // contains a __syncthreads
  for (int i = threadIdx.x; i < total; i += kThreads) qout[base + i] = s_q[i];
  if ((long long)(tile + 1) * kTile >= n && threadIdx.x == 0) *count = base + (unsigned long long)total;
}

__global__ void __launch_bounds__(kThreads)
reconstruct_kernel(const float* __restrict__ recon, const int8_t* __restrict__ mask,
                   const int32_t* __restrict__ q, long long n_q, long long n, float mean, float stdv, double thr2,
                   float* __restrict__ out, unsigned long long* ws, bool vec_r, bool vec_m,
                   bool vec_o) {
  const int tile = take_ticket(ws);
  const long long i0 = (long long)tile * kTile + (long long)threadIdx.x * kItems;
  float rv[kItems];
  load8_strided(recon, 1, i0, n, vec_r, rv);
  int m[kItems];
  if (i0 + kItems <= n && vec_m) {
    const unsigned long long pk = *reinterpret_cast<const unsigned long long*>(mask + i0);
#pragma unroll
    for (int k = 0; k < kItems; ++k) m[k] = (int)(int8_t)((pk >> (8 * k)) & 0xff);
  } else {
#pragma unroll
    for (int k = 0; k < kItems; ++k) m[k] = (i0 + k < n) ? (int)mask[i0 + k] : 0;
  }
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) cnt += (m[k] > 0);
  int total;
  const int texcl = block_exclusive<int>(cnt, total);
  const unsigned long long base = tile_prefix(ws, tile, (unsigned long long)total);
  long long o = (long long)base + texcl;
  float y[kItems];
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    int qk = 0;
    if (m[k] > 0) {               // never read past the code list, whatever the mask of a damaged blob claims
      if (o < n_q) qk = q[o];
      ++o;
    }
    y[k] = decoded_value(rv[k], mean, stdv, m[k], qk, thr2);
  }
  // the number of outliers the mask holds, for the caller's consistency check against the code count
  if ((long long)(tile + 1) * kTile >= n && threadIdx.x == 0) ws[0] = base + (unsigned long long)total;
  if (i0 + kItems <= n && vec_o) {
    float4* po = reinterpret_cast<float4*>(out + i0);
    po[0] = make_float4(y[0], y[1], y[2], y[3]);
    po[1] = make_float4(y[4], y[5], y[6], y[7]);
  } else {
#pragma unroll
    for (int k = 0; k < kItems; ++k)
      if (i0 + k < n) out[i0 + k] = y[k];
  }
}

// Latent quantisation: the bottleneck tensor is stored as IEEE half precision (round to nearest even, saturating),
// split into its low and high byte planes (only the high one -- sign, exponent, two mantissa bits -- is
// compressible).  The fp32 latent is overwritten with the dequantised values, which is what the decoder consumes in
// compress() and decompress() alike, so the reconstruction both sides see is identical.
__global__ void latent_quantize_kernel(float* __restrict__ lat, long long n, uint8_t* __restrict__ lo,
                                       uint8_t* __restrict__ hi) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    // The half-precision code is produced in a 32-bit register (cvt.rn.f16x2 with a zero upper half) and its bytes
    // are taken with 32-bit integer ops: ptxas 12.9 turns `st.u8` of a 16-bit register written by cvt.rn.f16.f32
    // into a numeric half->u8 conversion (F2I.U8.F16), which zeroed the low byte plane.
    const float v = fminf(fmaxf(lat[i], -65504.f), 65504.f);                // NaN propagates
    unsigned w;
    float back;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(0.f), "f"(v));
    asm("{ .reg .b16 l, h; mov.b32 {l, h}, %1; cvt.f32.f16 %0, l; }" : "=f"(back) : "r"(w));
    lo[i] = (uint8_t)(w & 0xffu);
    hi[i] = (uint8_t)((w >> 8) & 0xffu);
    lat[i] = back;
  }
}

__global__ void latent_dequantize_kernel(const uint8_t* __restrict__ lo, const uint8_t* __restrict__ hi, long long n,
                                         float* __restrict__ lat) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
  {
    const unsigned short u = (unsigned short)((unsigned)lo[i] | ((unsigned)hi[i] << 8));
    float back;
    asm("cvt.f32.f16 %0, %1;" : "=f"(back) : "h"(u));
    lat[i] = back;
  }
}

template <typename T>
__global__ void diff_kernel(const T* __restrict__ in, long long n, T* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = (i == 0) ? in[0] : (T)(in[i] - in[i - 1]);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
cumsum_kernel(const T* __restrict__ in, long long n, T* __restrict__ out, unsigned long long* ws) {
  const int tile = take_ticket(ws);
  const long long i0 = (long long)tile * kTile + (long long)threadIdx.x * kItems;
  long long v[kItems];
  long long s = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    v[k] = (i0 + k < n) ? (long long)in[i0 + k] : 0;
    s += v[k];
  }
  long long total;
  const long long texcl = block_exclusive<long long>(s, total);
  const unsigned long long base = tile_prefix(ws, tile, (unsigned long long)total);
  // 62-bit modular arithmetic: sign-extend
  long long run = (long long)(base << 2) >> 2;
  run += texcl;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    run += v[k];
    if (i0 + k < n) out[i0 + k] = (T)run;
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; }
inline long long tiles_for(long long n) { return (n + kTile - 1) / kTile; }

}  // namespace

extern "C" size_t lc_scan_workspace_bytes(int64_t n) { return ws_bytes(tiles_for(n < 1 ? 1 : n)); }

extern "C" int lc_quantize(const float* d_x, int C, const float* d_recon, int64_t n_vox, float mean,
                           float stdv, float thr, float thr2, int8_t* d_mask, int32_t* d_q,
                           unsigned long long* d_count, void* d_ws, void* stream) {
  LC_REQUIRE(d_x && d_recon && d_mask && d_q && d_count && d_ws, "null argument");
  LC_REQUIRE(n_vox > 0 && C >= 1, "empty input");
  cudaStream_t s = as_stream(stream);
  const long long nt = tiles_for(n_vox);
  LC_CUDA(cudaMemsetAsync(d_ws, 0, ws_bytes(nt), s));
  quantize_kernel<false><<<(unsigned)nt, kThreads, 0, s>>>(
      d_x, C, d_recon, n_vox, mean, stdv, thr, thr2, d_mask, d_q, d_count,
      reinterpret_cast<unsigned long long*>(d_ws), aligned16(d_x), aligned16(d_recon), aligned8(d_mask), 0.0, 0.0);
  LC_LAUNCH_CHECK();
  return LC_OK;
}

extern "C" int lc_quantize_checked(const float* d_x, int C, const float* d_recon, int64_t n_vox, float mean,
                                   float stdv, float thr, float thr2, double thr2_dec, double abs_err,
                                   int8_t* d_mask, int32_t* d_q, unsigned long long* d_out3, void* d_ws,
                                   void* stream) {
  LC_REQUIRE(d_x && d_recon && d_mask && d_q && d_out3 && d_ws, "null argument");
  LC_REQUIRE(n_vox > 0 && C >= 1, "empty input");
  cudaStream_t s = as_stream(stream);
  const long long nt = tiles_for(n_vox);
  LC_CUDA(cudaMemsetAsync(d_ws, 0, ws_bytes(nt), s));
  LC_CUDA(cudaMemsetAsync(d_out3, 0, 3 * sizeof(unsigned long long), s));
  quantize_kernel<true><<<(unsigned)nt, kThreads, 0, s>>>(
      d_x, C, d_recon, n_vox, mean, stdv, thr, thr2, d_mask, d_q, d_out3,
      reinterpret_cast<unsigned long long*>(d_ws), aligned16(d_x), aligned16(d_recon), aligned8(d_mask),
      thr2_dec, abs_err);
  LC_LAUNCH_CHECK();
  return LC_OK;
}

extern "C" int lc_reconstruct(const float* d_recon, const int8_t* d_mask, const int32_t* d_q, int64_t n_q,
                              int64_t n_vox, float mean, float stdv, double thr2, float* d_out,
                              void* d_ws, void* stream) {
  LC_REQUIRE(d_recon && d_mask && d_out && d_ws, "null argument");
  LC_REQUIRE(n_vox > 0 && n_q >= 0 && (n_q == 0 || d_q), "empty input");
  cudaStream_t s = as_stream(stream);
  const long long nt = tiles_for(n_vox);
  LC_CUDA(cudaMemsetAsync(d_ws, 0, ws_bytes(nt), s));
  reconstruct_kernel<<<(unsigned)nt, kThreads, 0, s>>>(d_recon, d_mask, d_q, n_q, n_vox, mean, stdv, thr2,
                                                       d_out, reinterpret_cast<unsigned long long*>(d_ws),
                                                       aligned16(d_recon), aligned8(d_mask),
                                                       aligned16(d_out));
  LC_LAUNCH_CHECK();
  return LC_OK;
}

extern "C" int lc_latent_quantize(float* d_latent, int64_t n, uint8_t* d_lo, uint8_t* d_hi, void* stream) {
  if (n <= 0) return LC_OK;
  LC_REQUIRE(d_latent && d_lo && d_hi, "null argument");
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  latent_quantize_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(d_latent, n, d_lo, d_hi);
  LC_LAUNCH_CHECK();
  return LC_OK;
}

extern "C" int lc_latent_dequantize(const uint8_t* d_lo, const uint8_t* d_hi, int64_t n, float* d_latent,
                                    void* stream) {
  if (n <= 0) return LC_OK;
  LC_REQUIRE(d_latent && d_lo && d_hi, "null argument");
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  latent_dequantize_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(d_lo, d_hi, n, d_latent);
  LC_LAUNCH_CHECK();
  return LC_OK;
}

extern "C" int lc_diff_i32(const int32_t* d_in, int64_t n, int32_t* d_out, void* stream) {
  if (n <= 0) return LC_OK;
  LC_REQUIRE(d_in && d_out && d_in != d_out, "bad argument");
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  diff_kernel<int32_t><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(d_in, n, d_out);
  LC_LAUNCH_CHECK();
  return LC_OK;
}

extern "C" int lc_diff_i8(const int8_t* d_in, int64_t n, int8_t* d_out, void* stream) {
  if (n <= 0) return LC_OK;
  LC_REQUIRE(d_in && d_out && d_in != d_out, "bad argument");
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  diff_kernel<int8_t><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(d_in, n, d_out);
  LC_LAUNCH_CHECK();
  return LC_OK;
}

extern "C" int lc_cumsum_i32(const int32_t* d_in, int64_t n, int32_t* d_out, void* d_ws, void* stream) {
  if (n <= 0) return LC_OK;
  LC_REQUIRE(d_in && d_out && d_ws, "null argument");
  cudaStream_t s = as_stream(stream);
  const long long nt = tiles_for(n);
  LC_CUDA(cudaMemsetAsync(d_ws, 0, ws_bytes(nt), s));
  cumsum_kernel<int32_t><<<(unsigned)nt, kThreads, 0, s>>>(d_in, n, d_out,
                                                           reinterpret_cast<unsigned long long*>(d_ws));
  LC_LAUNCH_CHECK();
  return LC_OK;
}

extern "C" int lc_cumsum_i8(const int8_t* d_in, int64_t n, int8_t* d_out, void* d_ws, void* stream) {
  if (n <= 0) return LC_OK;
  LC_REQUIRE(d_in && d_out && d_ws, "null argument");
  cudaStream_t s = as_stream(stream);
  const long long nt = tiles_for(n);
  LC_CUDA(cudaMemsetAsync(d_ws, 0, ws_bytes(nt), s));
  cumsum_kernel<int8_t><<<(unsigned)nt, kThreads, 0, s>>>(d_in, n, d_out,
                                                          reinterpret_cast<unsigned long long*>(d_ws));
  LC_LAUNCH_CHECK();
  return LC_OK;
}
