// Single-pass chained scan with decoupled look-back, shared by every kernel that needs a
// global exclusive prefix (outlier compaction, residual scatter, cumsum, escape lists,
// Huffman bit offsets).  One 64-bit word per tile carries {status:2, value:62} so a single
// store publishes both; tile ids follow the block index (in-order block dispatch gives forward progress).
#pragma once
#include "common.cuh"

namespace lcscan {

constexpr int kThreads = 256;
constexpr unsigned long long kMask = (1ull << 62) - 1;
constexpr unsigned long long kAgg = 1ull << 62;
constexpr unsigned long long kPre = 2ull << 62;

// workspace: word 0 = ticket, words 1.. = tile status
__host__ __device__ inline size_t ws_bytes(long long n_tiles) { return (size_t)(n_tiles + 2) * 8; }

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v));
}

// Returns this block's tile id (all threads).  Tiles are taken in block-index order: the hardware dispatches the
// blocks of a 1-D grid in increasing blockIdx, so every tile a block waits on in the look-back has already
// started (the same assumption CUB's single-pass scan makes).  An atomic ticket instead cost every block a
// global round trip plus a barrier before its first load -- half of all stall samples of the quantiser.
__device__ __forceinline__ int take_ticket(unsigned long long* ws) {
  (void)ws;
  return (int)blockIdx.x;
}

// Block-wide: publish `aggregate` (valid in thread 0) for `tile`, look back, return the
// exclusive prefix of the tile to all threads.  One status per lane and round trip: wider windows (2, 4, 8
// statuses per lane) were measured and are slower (quantiser 0.26 -> 0.27 / 0.28 / 0.31 ms, reconstruct 0.14 -> 0.18 /
// 0.19 / 0.23 ms on the benchmark field) -- the walk is short, the extra loads only add latency.  Value domain: unsigned 62-bit (wrapping sums
// of signed values are fine as long as the caller truncates consistently).
__device__ __forceinline__ unsigned long long tile_prefix(unsigned long long* ws, int tile,
                                                          unsigned long long aggregate) {
  __shared__ unsigned long long s_excl;
  unsigned long long* status = ws + 1;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    if (tile == 0) {
      if (lane == 0) {
        st_status(status + 0, kPre | (aggregate & kMask));
        s_excl = 0;
      }
    } else {
      if (lane == 0) st_status(status + tile, kAgg | (aggregate & kMask));
      unsigned long long excl = 0;
      int base = tile - 1;
      while (true) {
        const int idx = base - lane;
        unsigned long long st = kPre;  // virtual tiles before 0: prefix 0
        if (idx >= 0) {
          do { st = ld_status(status + idx); } while ((st >> 62) == 0);
        }
        const unsigned pm = __ballot_sync(0xffffffffu, (st >> 62) == 2);
        const int first = pm ? (__ffs(pm) - 1) : 32;
        unsigned long long v = (lane <= first) ? (st & kMask) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        excl += v;
        if (pm) break;
        base -= 32;
      }
      excl &= kMask;
      if (lane == 0) {
        st_status(status + tile, kPre | ((excl + aggregate) & kMask));
        s_excl = excl;
      }
    }
  }
  __syncthreads();
  return s_excl;
}

// Block exclusive scan of one value per thread (kThreads threads); returns exclusive prefix,
// total in `total` (all threads).
template <typename T>
__device__ __forceinline__ T block_exclusive(T v, T& total) {
  __shared__ T s_warp[kThreads / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  T inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) s_warp[wid] = inc;
  __syncthreads();
  T wbase = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) {
    T x = s_warp[w];
    if (w < wid) wbase += x;
    tot += x;
  }
  total = tot;
  __syncthreads();
  return wbase + inc - v;
}

}  // namespace lcscan
