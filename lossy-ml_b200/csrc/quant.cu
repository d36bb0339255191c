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

// CHECK: also decode every element exactly as reconstruct_kernel will and count |x - xhat| > abs_err (out3[1])
// with the largest error (out3[2], bits of a non-negative double) -- the strict mode's bound check without a second
// pass over the field.  The comparison is exact without fp64 on the common path: x and xhat agree to within a
// factor of two (Sterbenz), so d = fl32(x - xhat) is the exact difference, and |d| <= abs_err holds iff
// |d| <= chk_f, the largest float not above abs_err; only elements whose difference is not small against |x|
// (tiny values) take the fp64 comparison.  Per-tile verdicts go to 64-way partial words in the workspace BEFORE the
// tile publishes its scan aggregate, so the last tile -- whose look-back transitively waited for every publication --
// can fold them into out3 (two atomics per tile on one address each were a serial chain of 12 000 L2 round trips).
constexpr int kPart = 64;
// elements per thread of the quantiser.  16 (4096-element tiles: half the look-backs and block barriers per
// element, 128 registers) was measured against 8 on the benchmark field: 0.255 / 0.291 ms (plain / checked) against
// 0.242 / 0.276 ms -- the kernel is not paced by its barriers alone
#ifndef LC_QITEMS
#define LC_QITEMS 8
#endif
constexpr int kQItems = LC_QITEMS;
constexpr int kQTile = kThreads * kQItems;

template <bool CHECK, bool FULL>
__device__ __forceinline__ void quantize_tile(const float* __restrict__ x, int C, const float* __restrict__ recon,
                                              long long n, float mean, float stdv, float thr, float thr2,
                                              int8_t* __restrict__ mask, int32_t* __restrict__ qout,
                                              unsigned long long* __restrict__ count, unsigned long long* ws,
                                              double thr2_dec, double abs_err, float chk_f, int tile, int n_tiles,
                                              int32_t* s_q, double* s_worst, int* s_bad) {
  const long long i0 = (long long)tile * kQTile + (long long)threadIdx.x * kQItems;
  float xv[kQItems], rv[kQItems];
  if (FULL) {
    const float4* px = reinterpret_cast<const float4*>(x + i0 * 2);
#pragma unroll
    for (int k = 0; k < kQItems / 2; ++k) {
      const float4 a = __ldg(px + k);
      xv[2 * k] = a.x;
      xv[2 * k + 1] = a.z;
    }
    const float4* pr = reinterpret_cast<const float4*>(recon + i0);
#pragma unroll
    for (int k = 0; k < kQItems / 4; ++k) {
      const float4 a = __ldg(pr + k);
      rv[4 * k] = a.x; rv[4 * k + 1] = a.y; rv[4 * k + 2] = a.z; rv[4 * k + 3] = a.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < kQItems; ++k) {
      xv[k] = (i0 + k < n) ? __ldg(x + (i0 + k) * C) : 0.f;
      rv[k] = (i0 + k < n) ? __ldg(recon + i0 + k) : 0.f;
    }
  }
  int q[kQItems];
  unsigned long long pk[kQItems / 8] = {};
  int cnt = 0, bad = 0;
  float worst = 0.f;
  double worst_d = 0.0;
#pragma unroll
  for (int k = 0; k < kQItems; ++k) {
    const bool in = FULL || (i0 + k < n);
    const float rr = __fadd_rn(__fmul_rn(rv[k], stdv), mean);   // compress.py:114
    const float e = __fsub_rn(xv[k], rr);                       // compress.py:117
    const float a = fabsf(e);
    const bool out = in && (a > thr);                           // compress.py:121 (|e| <= thr -> 0; NaN -> 0)
    int m = 0;
    q[k] = 0;
    if (out) {
      q[k] = __float2int_rn(__fdiv_rn(a, thr2));                // compress.py:135
      m = e > 0.f ? 1 : 2;                                      // compress.py:122-123
    }
    cnt += out;
    pk[k >> 3] |= (unsigned long long)m << (8 * (k & 7));
    if (CHECK && in) {
      // decompress.py:86-104: val = fl32(fl64(q) * thr2), negated where mask == 2; xhat = fl32(rr + val)
      float val = 0.f;
      if (out) {
        val = (float)((double)q[k] * thr2_dec);
        if (m == 2) val = -val;
      }
      const float y = __fadd_rn(rr, val);
      const float d = fabsf(__fsub_rn(xv[k], y));
      if (d <= 0.49f * fabsf(xv[k])) {          // x/2 <= xhat <= 2x: the float difference is exact
        bad += !(d <= chk_f);
        worst = fmaxf(worst, d);
      } else {                                  // tiny values (or NaN): compare in fp64
        const double dd = fabs((double)xv[k] - (double)y);
        bad += !(dd <= abs_err);
        worst_d = fmax(worst_d, dd);
      }
    }
  }
  if (FULL) {
#pragma unroll
    for (int j = 0; j < kQItems / 8; ++j) reinterpret_cast<unsigned long long*>(mask + i0)[j] = pk[j];
  } else {
#pragma unroll
    for (int k = 0; k < kQItems; ++k)
      if (i0 + k < n) mask[i0 + k] = (int8_t)((pk[k >> 3] >> (8 * (k & 7))) & 0xff);
  }
  // compaction through shared memory: the tile's outliers are gathered in C order and written out as one
  // coalesced run (per-thread 4-byte stores would touch every 32-byte sector several times, each a partial write)
  if (CHECK) {
    double w = fmax((double)worst, worst_d);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      bad += __shfl_xor_sync(0xffffffffu, bad, o);
      w = fmax(w, __shfl_xor_sync(0xffffffffu, w, o));
    }
    if ((threadIdx.x & 31) == 0) { s_worst[threadIdx.x >> 5] = w; s_bad[threadIdx.x >> 5] = bad; }
  }
  int total;
  const int texcl = block_exclusive<int>(cnt, total);      // contains __syncthreads: s_worst / s_bad are visible after it
  {
    int o = texcl;
#pragma unroll
    for (int k = 0; k < kQItems; ++k)
      if ((pk[k >> 3] >> (8 * (k & 7))) & 0xff) s_q[o++] = q[k];
  }
  unsigned long long* part = ws + 1 + n_tiles + 1;         // [kPart] violation counts, [kPart] worst-error bits
  if (CHECK && threadIdx.x == 0) {
    double w = s_worst[0];
    int bsum = s_bad[0];
#pragma unroll
    for (int i = 1; i < kThreads / 32; ++i) { w = fmax(w, s_worst[i]); bsum += s_bad[i]; }
    if (bsum) atomicAdd(part + (tile & (kPart - 1)), (unsigned long long)bsum);
    const unsigned long long wb = (unsigned long long)__double_as_longlong(w);
    unsigned long long* pw = part + kPart + (tile & (kPart - 1));
    if (wb > ld_status(pw)) atomicMax(pw, wb);
    __threadfence();                                       // before this tile's aggregate is published
  }
  const unsigned long long base = tile_prefix(ws, tile, (unsigned long long)total);   // contains a __syncthreads
  for (int i = threadIdx.x; i < total; i += kThreads) qout[base + i] = s_q[i];
  if (tile == n_tiles - 1 && threadIdx.x == 0) {
    count[0] = base + (unsigned long long)total;
    if (CHECK) {
      __threadfence();
      unsigned long long b2 = 0, w2 = 0;
      for (int i = 0; i < kPart; ++i) {
        b2 += ld_status(part + i);
        const unsigned long long v = ld_status(part + kPart + i);
        w2 = v > w2 ? v : w2;
      }
      count[1] = b2;
      count[2] = w2;
    }
  }
}

template <bool CHECK>
__global__ void __launch_bounds__(kThreads, (kQItems > 8) ? 2 : 4)
quantize_kernel(const float* __restrict__ x, int C, const float* __restrict__ recon, long long n,
                float mean, float stdv, float thr, float thr2, int8_t* __restrict__ mask,
                int32_t* __restrict__ qout, unsigned long long* __restrict__ count,
                unsigned long long* ws, bool vec, double thr2_dec, double abs_err, float chk_f, int n_tiles) {
  __shared__ int32_t s_q[kQTile];
  __shared__ double s_worst[kThreads / 32];
  __shared__ int s_bad[kThreads / 32];
  const int tile = take_ticket(ws);
  // block-uniform: whole tile in range, two interleaved channels, every pointer aligned for vector access
  if (vec && (long long)(tile + 1) * kQTile <= n)
    quantize_tile<CHECK, true>(x, C, recon, n, mean, stdv, thr, thr2, mask, qout, count, ws, thr2_dec, abs_err, chk_f,
                               tile, n_tiles, s_q, s_worst, s_bad);
  else
    quantize_tile<CHECK, false>(x, C, recon, n, mean, stdv, thr, thr2, mask, qout, count, ws, thr2_dec, abs_err,
                                chk_f, tile, n_tiles, s_q, s_worst, s_bad);
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

// status words of the look-back scan + the 2 x 64 partial verdict words of quantize_kernel<true>
extern "C" size_t lc_scan_workspace_bytes(int64_t n) { return ws_bytes(tiles_for(n < 1 ? 1 : n)) + 2 * kPart * 8; }

// largest float that does not exceed the (positive) double v
static float float_not_above(double v) {
  float f = (float)v;
  if ((double)f > v) f = nextafterf(f, 0.f);
  return f;
}

extern "C" int lc_quantize(const float* d_x, int C, const float* d_recon, int64_t n_vox, float mean,
                           float stdv, float thr, float thr2, int8_t* d_mask, int32_t* d_q,
                           unsigned long long* d_count, void* d_ws, void* stream) {
  LC_REQUIRE(d_x && d_recon && d_mask && d_q && d_count && d_ws, "null argument");
  LC_REQUIRE(n_vox > 0 && C >= 1, "empty input");
  cudaStream_t s = as_stream(stream);
  const long long nt = (n_vox + kQTile - 1) / kQTile;
  LC_CUDA(cudaMemsetAsync(d_ws, 0, ws_bytes(nt), s));
  const bool vec = C == 2 && aligned16(d_x) && aligned16(d_recon) && aligned8(d_mask);
  quantize_kernel<false><<<(unsigned)nt, kThreads, 0, s>>>(
      d_x, C, d_recon, n_vox, mean, stdv, thr, thr2, d_mask, d_q, d_count,
      reinterpret_cast<unsigned long long*>(d_ws), vec, 0.0, 0.0, 0.f, (int)nt);
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
  LC_REQUIRE(abs_err > 0.0, "abs_err must be positive");
  const long long nt = (n_vox + kQTile - 1) / kQTile;
  LC_CUDA(cudaMemsetAsync(d_ws, 0, ws_bytes(nt) + 2 * kPart * 8, s));
  const bool vec = C == 2 && aligned16(d_x) && aligned16(d_recon) && aligned8(d_mask);
  quantize_kernel<true><<<(unsigned)nt, kThreads, 0, s>>>(
      d_x, C, d_recon, n_vox, mean, stdv, thr, thr2, d_mask, d_q, d_out3,
      reinterpret_cast<unsigned long long*>(d_ws), vec, thr2_dec, abs_err, float_not_above(abs_err), (int)nt);
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
