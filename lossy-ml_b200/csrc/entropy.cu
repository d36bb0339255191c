// Entropy stage: byte-symbol histogram, symbol mapping of the code / mask streams and Huffman
// bit packing / unpacking from a host-built table (the host builder mirrors huffman.py:20-45).
#include "scan.cuh"

using namespace lcscan;

namespace {

// ---------------------------------------------------------------- histogram (warp-aggregated)
__global__ void __launch_bounds__(256) hist_kernel(const uint8_t* __restrict__ sym, long long n,
                                                   uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[8][256];   // one private histogram per warp: no inter-warp contention
  for (int i = threadIdx.x; i < 8 * 256; i += 256) (&sh[0][0])[i] = 0;
  __syncthreads();
  uint32_t* mine = sh[threadIdx.x >> 5];
  const long long nvec = n / 16;
  const uint4* p = reinterpret_cast<const uint4*>(sym);
  const bool vec = (reinterpret_cast<uintptr_t>(sym) & 15) == 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    for (long long i = tid; i < nvec; i += stride) {
      const uint4 v = __ldg(p + i);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const uint32_t s = (w[k] >> (8 * b)) & 0xff;
          // warp-aggregate: lanes holding the same symbol elect one leader that adds the count
          const unsigned peers = __match_any_sync(__activemask(), s);
          if ((__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&mine[s], __popc(peers));
        }
      }
    }
    for (long long i = nvec * 16 + tid; i < n; i += stride) atomicAdd(&mine[sym[i]], 1u);
  } else {
    for (long long i = tid; i < n; i += stride) atomicAdd(&mine[sym[i]], 1u);
  }
  __syncthreads();
  for (int s = threadIdx.x; s < 256; s += 256) {
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[w][s];
    if (t) atomicAdd(&hist[s], t);
  }
}

// ---------------------------------------------------------------- histograms of the code stream
// Both candidate symbolisations of the int32 codes in one pass: hist[0..255] of (q + bias0) and
// hist[256..511] of (q[i] - q[i-1] + bias1), out-of-range values counted under the escape symbol.
// Per-lane shared-memory atomics on one private pair of histograms per warp.  A warp-aggregated form was measured
// on the benchmark field (16.6 M codes, 67 % of them the same symbol): electing a leader per distinct symbol with
// ballots / __match_any_sync and adding the group's size took 88 us against 23 us (2.9 TB/s, 74 % issue slots) for
// the plain form -- the shared-memory atomic unit of sm_100 already merges lanes that hit the same address, so the
// election only adds instructions.  (hist_kernel keeps its aggregation: its 16 symbols per load amortise the match.)
__global__ void __launch_bounds__(256) hist_codes_kernel(const int32_t* __restrict__ q, long long n, int esc,
                                                         int bias0, int bias1, uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[8][512];   // one private pair of histograms per warp
  for (int i = threadIdx.x; i < 8 * 512; i += 256) (&sh[0][0])[i] = 0;
  __syncthreads();
  uint32_t* mine = sh[threadIdx.x >> 5];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int cur = q[i];
    const int prev = (i > 0) ? q[i - 1] : 0;
    int s0 = cur + bias0, s1 = cur - prev + bias1;
    if (s0 < 0 || s0 >= esc) s0 = esc;
    if (s1 < 0 || s1 >= esc) s1 = esc;
    atomicAdd(&mine[s0], 1u);
    atomicAdd(&mine[256 + s1], 1u);
  }
  __syncthreads();
  for (int s = threadIdx.x; s < 512; s += 256) {
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[w][s];
    if (t) atomicAdd(&hist[s], t);
  }
}

// ---------------------------------------------------------------- trits <-> bytes
__global__ void pack_trits_kernel(const int8_t* __restrict__ m, long long n, uint8_t* __restrict__ out) {
  const long long nb = (n + 4) / 5;
  for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < nb;
       b += (long long)gridDim.x * blockDim.x) {
    const long long i = b * 5;
    int v = 0, mul = 1;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int t = (i + k < n) ? (int)m[i + k] : 0;
      v += t * mul;
      mul *= 3;
    }
    out[b] = (uint8_t)v;
  }
}

// 16 output bytes (80 mask values) per thread: five 16-byte loads, one 16-byte store, and the histogram of the
// packed bytes in the same pass (zero bytes -- most of a sparse mask -- are counted per warp through a ballot, the
// others per byte), so that the entropy stage needs no separate pass to decide how to code the mask.
__global__ void __launch_bounds__(256) pack_trits_hist_kernel(const int8_t* __restrict__ m, long long n,
                                                              uint8_t* __restrict__ out, uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[256];
  for (int i = threadIdx.x; i < 256; i += 256) sh[i] = 0;
  __syncthreads();
  const long long nb = (n + 4) / 5;
  const long long ng = (nb + 15) / 16;                 // groups of 16 output bytes
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g - (threadIdx.x & 31) < ng; g += stride) {
    int zeros = 0;
    if (g < ng) {
      const long long b0 = g * 16, i0 = b0 * 5;
      uint8_t v[16];
      if (vec && i0 + 80 <= n) {
        uint32_t w[20];
        const uint4* p = reinterpret_cast<const uint4*>(m + i0);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const uint4 t = __ldg(p + k);
          w[4 * k] = t.x; w[4 * k + 1] = t.y; w[4 * k + 2] = t.z; w[4 * k + 3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          int acc = 0, mul = 1;
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            const int idx = j * 5 + k;
            acc += (int)((w[idx >> 2] >> (8 * (idx & 3))) & 0xff) * mul;
            mul *= 3;
          }
          v[j] = (uint8_t)acc;
        }
        *reinterpret_cast<uint4*>(out + b0) =
            make_uint4(v[0] | (v[1] << 8) | (v[2] << 16) | ((uint32_t)v[3] << 24),
                       v[4] | (v[5] << 8) | (v[6] << 16) | ((uint32_t)v[7] << 24),
                       v[8] | (v[9] << 8) | (v[10] << 16) | ((uint32_t)v[11] << 24),
                       v[12] | (v[13] << 8) | (v[14] << 16) | ((uint32_t)v[15] << 24));
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (v[j] == 0) ++zeros;
          else atomicAdd(&sh[v[j]], 1u);
        }
      } else {
        for (int j = 0; j < 16 && b0 + j < nb; ++j) {
          int acc = 0, mul = 1;
          for (int k = 0; k < 5; ++k) {
            const long long i = i0 + j * 5 + k;
            acc += ((i < n) ? (int)m[i] : 0) * mul;
            mul *= 3;
          }
          out[b0 + j] = (uint8_t)acc;
          if (acc == 0) ++zeros;
          else atomicAdd(&sh[acc], 1u);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) zeros += __shfl_xor_sync(0xffffffffu, zeros, o);
    if ((threadIdx.x & 31) == 0 && zeros) atomicAdd(&sh[0], (uint32_t)zeros);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += 256)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

__global__ void unpack_trits_kernel(const uint8_t* __restrict__ in, long long n, int8_t* __restrict__ m) {
  const long long nb = (n + 4) / 5;
  for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < nb;
       b += (long long)gridDim.x * blockDim.x) {
    int v = in[b];
    const long long i = b * 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      if (i + k < n) m[i + k] = (int8_t)(v % 3);
      v /= 3;
    }
  }
}

// ---------------------------------------------------------------- zero-run tokens
// A Huffman code spends at least one bit per symbol, so a mask whose 5-trit bytes are almost all zero (large
// abs_error) cannot go below n/40 bytes however small its entropy is.  Runs of zero bytes are therefore folded
// into the alphabet itself: bytes 0..242 are literals (0 = a single zero byte) and 243+k stands for 2^(k+1) zero
// bytes (k = 0..4).  Runs are confined to the 32 bytes one thread owns and are decomposed greedily, largest power
// of two first, so tokenising is thread-local bit arithmetic; the token stream stays a byte stream for the same
// Huffman packer.  Whether a mask is coded plain or tokenised is decided from the two histograms.  The same
// kernels serve the code stream at large abs_error (nearly every code is 1): run symbol and first run token are
// parameters (the five run tokens must be unused by the stream, which the caller checks on its histogram).
constexpr int kRleSeg = 32;                        // bytes per thread = longest run
constexpr int kRleTile = kThreads * kRleSeg;       // 8192 bytes per block

__global__ void __launch_bounds__(kThreads)
rle_tokens_kernel(const uint8_t* __restrict__ sym, long long n, int run_sym, int run_first,
                  uint8_t* __restrict__ tok, unsigned long long* __restrict__ n_tok, uint32_t* __restrict__ hist,
                  uint32_t* __restrict__ hist_in, unsigned long long* ws) {
  __shared__ uint8_t s_tok[kRleTile];
  __shared__ uint32_t s_hist[256], s_hin[256];
  if (threadIdx.x < 256) { s_hist[threadIdx.x] = 0; s_hin[threadIdx.x] = 0; }
  __syncthreads();   // the shared tables above are read by every thread below
  const int tile = take_ticket(ws);
  const long long i0 = (long long)tile * kRleTile + (long long)threadIdx.x * kRleSeg;
  uint8_t b[kRleSeg];
  int len = 0;                         // valid bytes of this thread's segment
  if (i0 + kRleSeg <= n && (reinterpret_cast<uintptr_t>(sym) & 15) == 0) {
    const uint4 v0 = *reinterpret_cast<const uint4*>(sym + i0), v1 = *reinterpret_cast<const uint4*>(sym + i0 + 16);
    const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int k = 0; k < kRleSeg; ++k) b[k] = (w[k >> 2] >> (8 * (k & 3))) & 0xff;
    len = kRleSeg;
  } else {
#pragma unroll
    for (int k = 0; k < kRleSeg; ++k) b[k] = (i0 + k < n) ? sym[i0 + k] : 0;
    len = (i0 < n) ? (int)((n - i0 < kRleSeg) ? (n - i0) : kRleSeg) : 0;
  }
  unsigned nz = 0;                     // bit k: byte k is not part of a run (padding beyond len never is)
#pragma unroll
  for (int k = 0; k < kRleSeg; ++k) nz |= (unsigned)(b[k] != run_sym || k >= len) << k;
  // token starts: every literal, and inside a zero run [s, s+R) the offsets j with j = R with its low bits cleared
  unsigned start = 0;
  uint8_t val[kRleSeg];
#pragma unroll
  for (int k = 0; k < kRleSeg; ++k) {
    val[k] = b[k];
    if (k < len) {
      if ((nz >> k) & 1u) {
        start |= 1u << k;
      } else {
        const unsigned below = nz & ((1u << k) - 1u);
        const int s0 = below ? (32 - __clz(below)) : 0;
        const unsigned above = (k == 31) ? 0u : (nz >> (k + 1)) << (k + 1);
        const int e = above ? (__ffs(above) - 1) : kRleSeg;
        const int j = k - s0, R = e - s0;
        const int hb = 31 - __clz((unsigned)(R ^ j));       // R > j, so R ^ j != 0
        if ((j & ((2 << hb) - 1)) == 0) {
          start |= 1u << k;
          val[k] = hb ? (uint8_t)(run_first - 1 + hb) : (uint8_t)run_sym;
        }
      }
    }
  }
  if (hist_in) {
    // histogram of the input bytes in the same pass (the plain-coding candidate's table needs it): the run symbol
    // is counted per warp, the rest per byte
    int runs = __popc(~nz);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) runs += __shfl_xor_sync(0xffffffffu, runs, o);
    if ((threadIdx.x & 31) == 0 && runs) atomicAdd(&s_hin[run_sym], (unsigned)runs);
#pragma unroll
    for (int k = 0; k < kRleSeg; ++k)
      if (k < len && b[k] != run_sym) atomicAdd(&s_hin[b[k]], 1u);
  }
  const int cnt = __popc(start);
  int total;
  const int texcl = block_exclusive<int>(cnt, total);
  {
    int o = texcl;
#pragma unroll
    for (int k = 0; k < kRleSeg; ++k)
      if ((start >> k) & 1u) {
        s_tok[o++] = val[k];
        atomicAdd(&s_hist[val[k]], 1u);
      }
  }
  const unsigned long long base = tile_prefix(ws, tile, (unsigned long long)total);   // contains a __syncthreads
  for (int i = threadIdx.x; i < total; i += kThreads) tok[base + i] = s_tok[i];
  if (threadIdx.x < 256 && s_hist[threadIdx.x]) atomicAdd(&hist[threadIdx.x], s_hist[threadIdx.x]);
  if (hist_in && threadIdx.x < 256 && s_hin[threadIdx.x]) atomicAdd(&hist_in[threadIdx.x], s_hin[threadIdx.x]);
  if ((long long)(tile + 1) * kRleTile >= n && threadIdx.x == 0) *n_tok = base + (unsigned long long)total;
}

// tokens -> bytes: exclusive scan of the token lengths gives every literal its position; the output was filled
// with the run symbol
constexpr int kExpItems = 16;
__global__ void __launch_bounds__(kThreads)
rle_expand_kernel(const uint8_t* __restrict__ tok, long long n_tok, int run_sym, int run_first,
                  uint8_t* __restrict__ out, long long n_out, unsigned long long* ws) {
  const int tile = take_ticket(ws);
  const long long i0 = (long long)tile * (kThreads * kExpItems) + (long long)threadIdx.x * kExpItems;
  uint8_t t[kExpItems];
  int sum = 0;
#pragma unroll
  for (int k = 0; k < kExpItems; ++k) {
    t[k] = (i0 + k < n_tok) ? tok[i0 + k] : 0;
    const int r = (int)t[k] - run_first;                   // 0..4: a run of 2 << r
    if (i0 + k < n_tok) sum += (r < 0 || r > 4) ? 1 : (2 << r);
  }
  int total;
  const int texcl = block_exclusive<int>(sum, total);
  const unsigned long long base = tile_prefix(ws, tile, (unsigned long long)total);
  long long pos = (long long)base + texcl;
#pragma unroll
  for (int k = 0; k < kExpItems; ++k) {
    if (i0 + k < n_tok) {
      const int r = (int)t[k] - run_first;
      if (r < 0 || r > 4) {
        if (t[k] != run_sym && pos < n_out) out[pos] = t[k];
        pos += 1;
      } else {
        pos += 2 << r;
      }
    }
  }
}

// ---------------------------------------------------------------- int32 codes <-> byte symbols
constexpr int kItems = 8;
constexpr int kTile = kThreads * kItems;

__global__ void __launch_bounds__(kThreads)
symbolize_kernel(const int32_t* __restrict__ q, long long n, int diff, int bias, int esc,
                 uint8_t* __restrict__ sym, int32_t* __restrict__ escv,
                 unsigned long long* __restrict__ nesc, unsigned long long* ws, bool vec) {
  const int tile = take_ticket(ws);
  const long long i0 = (long long)tile * kTile + (long long)threadIdx.x * kItems;
  int cur[kItems];
  const bool full = vec && i0 + kItems <= n;
  if (full) {             // two 16-byte loads, one 8-byte store per thread
    const int4 a = __ldg(reinterpret_cast<const int4*>(q + i0)), b = __ldg(reinterpret_cast<const int4*>(q + i0) + 1);
    cur[0] = a.x; cur[1] = a.y; cur[2] = a.z; cur[3] = a.w; cur[4] = b.x; cur[5] = b.y; cur[6] = b.z; cur[7] = b.w;
  } else {
#pragma unroll
    for (int k = 0; k < kItems; ++k) cur[k] = (i0 + k < n) ? q[i0 + k] : 0;
  }
  int v[kItems];
  unsigned emask = 0;
  unsigned long long pk = 0;
  int prev = (diff && i0 > 0 && i0 - 1 < n) ? q[i0 - 1] : 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    v[k] = diff ? (cur[k] - prev) : cur[k];
    prev = cur[k];
    const int s = v[k] + bias;
    const bool e = (i0 + k < n) && (s < 0 || s >= esc);
    emask |= (unsigned)e << k;
    pk |= (unsigned long long)(uint8_t)(e ? esc : s) << (8 * k);
  }
  if (full) {
    *reinterpret_cast<unsigned long long*>(sym + i0) = pk;
  } else {
#pragma unroll
    for (int k = 0; k < kItems; ++k)
      if (i0 + k < n) sym[i0 + k] = (uint8_t)(pk >> (8 * k));
  }
  int total;
  const int texcl = block_exclusive<int>(__popc(emask), total);
  const unsigned long long base = tile_prefix(ws, tile, (unsigned long long)total);
  if (emask) {
    long long o = (long long)base + texcl;
#pragma unroll
    for (int k = 0; k < kItems; ++k)
      if ((emask >> k) & 1u) escv[o++] = v[k];
  }
  if ((long long)(tile + 1) * kTile >= n && threadIdx.x == 0) *nesc = base + (unsigned long long)total;
}

__global__ void __launch_bounds__(kThreads)
unsymbolize_kernel(const uint8_t* __restrict__ sym, long long n, int bias, int esc,
                   const int32_t* __restrict__ escv, int32_t* __restrict__ q, unsigned long long* ws, bool vec) {
  const int tile = take_ticket(ws);
  const long long i0 = (long long)tile * kTile + (long long)threadIdx.x * kItems;
  int s[kItems];
  const bool full = vec && i0 + kItems <= n;
  if (full) {
    const unsigned long long pk = __ldg(reinterpret_cast<const unsigned long long*>(sym + i0));
#pragma unroll
    for (int k = 0; k < kItems; ++k) s[k] = (int)((pk >> (8 * k)) & 0xff);
  } else {
#pragma unroll
    for (int k = 0; k < kItems; ++k) s[k] = (i0 + k < n) ? (int)sym[i0 + k] : 0;
  }
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) cnt += (i0 + k < n) && (s[k] == esc);
  int total;
  const int texcl = block_exclusive<int>(cnt, total);
  const unsigned long long base = tile_prefix(ws, tile, (unsigned long long)total);
  long long o = (long long)base + texcl;
  int out[kItems];
#pragma unroll
  for (int k = 0; k < kItems; ++k) out[k] = (i0 + k < n && s[k] == esc) ? escv[o++] : (s[k] - bias);
  if (full) {
    int4* po = reinterpret_cast<int4*>(q + i0);
    po[0] = make_int4(out[0], out[1], out[2], out[3]);
    po[1] = make_int4(out[4], out[5], out[6], out[7]);
  } else {
#pragma unroll
    for (int k = 0; k < kItems; ++k)
      if (i0 + k < n) q[i0 + k] = out[k];
  }
}

// Escape-free forms (the caller knows from its histogram / the container header that no value leaves the symbol
// range): plain element-wise maps, 16 codes per thread, no scan.
__global__ void __launch_bounds__(256) symbolize_plain_kernel(const int32_t* __restrict__ q, long long n, int diff,
                                                              int bias, uint8_t* __restrict__ sym, bool vec) {
  const long long ng = (n + 15) / 16;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += (long long)gridDim.x * blockDim.x) {
    const long long i0 = g * 16;
    int prev = (diff && i0 > 0) ? q[i0 - 1] : 0;
    if (vec && i0 + 16 <= n) {
      uint32_t w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int4 a = __ldg(reinterpret_cast<const int4*>(q + i0) + k);
        const int c[4] = {a.x, a.y, a.z, a.w};
        uint32_t pk = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int v = diff ? (c[j] - prev) : c[j];
          prev = c[j];
          pk |= (uint32_t)(uint8_t)(v + bias) << (8 * j);
        }
        w[k] = pk;
      }
      *reinterpret_cast<uint4*>(sym + i0) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
      for (int k = 0; k < 16 && i0 + k < n; ++k) {
        const int c = q[i0 + k];
        sym[i0 + k] = (uint8_t)((diff ? (c - prev) : c) + bias);
        prev = c;
      }
    }
  }
}

__global__ void __launch_bounds__(256) unsymbolize_plain_kernel(const uint8_t* __restrict__ sym, long long n, int bias,
                                                                int32_t* __restrict__ q, bool vec) {
  const long long ng = (n + 15) / 16;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += (long long)gridDim.x * blockDim.x) {
    const long long i0 = g * 16;
    if (vec && i0 + 16 <= n) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(sym + i0));
      const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        reinterpret_cast<int4*>(q + i0)[k] =
            make_int4((int)(w[k] & 0xff) - bias, (int)((w[k] >> 8) & 0xff) - bias, (int)((w[k] >> 16) & 0xff) - bias,
                      (int)(w[k] >> 24) - bias);
    } else {
      for (int k = 0; k < 16 && i0 + k < n; ++k) q[i0 + k] = (int)sym[i0 + k] - bias;
    }
  }
}

// ---------------------------------------------------------------- Huffman pack
// one tile = kThreads threads x 16 consecutive symbols, cut into independently decodable segments of `seg`
// symbols (a power of two between 64 and the tile: 1024 for the large streams, 256 for the small latent plane
// whose decoding sits on decompress()'s critical path -- one thread decodes one segment)
constexpr int kSegItems = 16;
constexpr int kTileSyms = kThreads * kSegItems;

__global__ void __launch_bounds__(kThreads)
huff_encode_kernel(const uint8_t* __restrict__ sym, long long n, const uint32_t* __restrict__ code,
                   const uint8_t* __restrict__ len, uint32_t* __restrict__ out,
                   unsigned long long* __restrict__ seg_bits, unsigned long long* __restrict__ total_bits,
                   unsigned long long* ws, int seg_threads) {
  __shared__ uint32_t s_code[256];
  __shared__ uint8_t s_len[256];
  if (threadIdx.x < 256) {
    s_code[threadIdx.x] = code[threadIdx.x];
    s_len[threadIdx.x] = len[threadIdx.x];
  }
  __syncthreads();   // the shared tables above are read by every thread below
  const int tile = take_ticket(ws);
  const long long i0 = (long long)tile * kTileSyms + (long long)threadIdx.x * kSegItems;
  uint8_t s[kSegItems];
  if (i0 + kSegItems <= n && (reinterpret_cast<uintptr_t>(sym) & 15) == 0) {
    const uint4 v = *reinterpret_cast<const uint4*>(sym + i0);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < kSegItems; ++k) s[k] = (w[k >> 2] >> (8 * (k & 3))) & 0xff;
  } else {
#pragma unroll
    for (int k = 0; k < kSegItems; ++k) s[k] = (i0 + k < n) ? sym[i0 + k] : 0;
  }
  int bits = 0;
#pragma unroll
  for (int k = 0; k < kSegItems; ++k) bits += (i0 + k < n) ? (int)s_len[s[k]] : 0;
  int total;
  const int texcl = block_exclusive<int>(bits, total);
  const unsigned long long base = tile_prefix(ws, tile, (unsigned long long)total);
  if ((threadIdx.x & (seg_threads - 1)) == 0 && i0 < n)
    seg_bits[(long long)tile * (kThreads / seg_threads) + threadIdx.x / seg_threads] = base + (unsigned long long)texcl;
  if (threadIdx.x == 0 && (long long)(tile + 1) * kTileSyms >= n) *total_bits = base + (unsigned long long)total;
  // emit: big-endian bit order within the byte stream == big-endian within 32-bit words that are
  // byte-swapped on store.  Accumulate into a 64-bit window and flush 32 bits at a time.
  unsigned long long pos = base + (unsigned long long)texcl;
  unsigned long long word_idx = pos >> 5;
  int fill = (int)(pos & 31);            // bits already used in the current word (by a neighbour)
  unsigned long long acc = 0;            // top `fill + pending` bits valid, left-aligned at bit 63
  int have = fill;
#pragma unroll
  for (int k = 0; k < kSegItems; ++k) {
    if (i0 + k < n) {
      const int l = s_len[s[k]];
      if (l) {
        acc |= (unsigned long long)s_code[s[k]] << (64 - have - l);
        have += l;
        if (have >= 32) {
          const uint32_t w = (uint32_t)(acc >> 32);
          atomicOr(&out[word_idx], __byte_perm(w, 0, 0x0123));
          ++word_idx;
          acc <<= 32;
          have -= 32;
        }
      }
    }
  }
  if (have > 0) {
    const uint32_t w = (uint32_t)(acc >> 32);
    if (w) atomicOr(&out[word_idx], __byte_perm(w, 0, 0x0123));
  }
}

// ---------------------------------------------------------------- Huffman unpack
// One thread per segment: a 64-bit bit window refilled with aligned 32-bit words (the stream is
// big-endian within bytes), 12-bit direct table in shared memory, tree walk for longer codes,
// symbols stored four at a time.
__global__ void __launch_bounds__(128)
huff_decode_kernel(const uint32_t* __restrict__ words, long long n_words, long long n, int seg,
                   const unsigned long long* __restrict__ seg_bits, const uint16_t* __restrict__ lut,
                   int lut_bits, const int32_t* __restrict__ tree, uint8_t* __restrict__ sym) {
  extern __shared__ uint16_t s_lut[];
  for (int i = threadIdx.x; i < (1 << lut_bits); i += blockDim.x) s_lut[i] = lut[i];
  __syncthreads();
  const long long nseg = (n + seg - 1) / seg;
  const long long sg = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (sg >= nseg) return;
  // Reads never leave the buffer whatever a corrupted container holds: the start offset is clamped and the word
  // pointer saturates at the last word (a damaged stream then decodes to garbage symbols, which the caller's
  // length / count checks reject, instead of faulting).
  const uint32_t* const wend = words + (n_words - 1);
  unsigned long long pos = seg_bits[sg];
  if ((long long)(pos >> 5) > n_words - 3) pos = 0;
  const uint32_t* wp = words + (pos >> 5);
  unsigned long long buf = ((unsigned long long)__byte_perm(wp[0], 0, 0x0123) << 32) | __byte_perm(wp[1], 0, 0x0123);
  wp += 2;
  uint32_t ahead = *wp;                                  // next word, loaded one refill early (the stream has 16 spare bytes)
  int used = (int)(pos & 31);
  const long long i0 = sg * seg;
  const long long i1 = (i0 + seg < n) ? i0 + seg : n;
  const bool one_symbol = (tree[0] < 0 && tree[1] < 0 && tree[0] == tree[1]);
  uint32_t pack = 0;
  for (long long i = i0; i < i1; ++i) {
    const unsigned long long win = buf << used;          // >= 33 valid bits at the top
    const uint16_t e = s_lut[win >> (64 - lut_bits)];
    int l = e >> 8;
    int s = e & 0xff;
    if (l == 0) {   // code longer than lut_bits (or the empty code of a one-symbol alphabet)
      if (one_symbol) {
        s = ~tree[0];
      } else {
        int node = 0;
        while (l < 32) {                                   // codes are at most 32 bits (lc_huff_build_decode checks)
          const int bit = (int)((win >> (63 - l)) & 1);
          ++l;
          const int nx = tree[2 * node + bit];
          if (nx < 0) { s = ~nx; break; }
          node = nx;
        }
      }
    }
    used += l;
    if (used >= 32) {
      buf = (buf << 32) | __byte_perm(ahead, 0, 0x0123);
      used -= 32;
      wp = (wp < wend) ? wp + 1 : wp;
      ahead = *wp;
    }
    const int k = (int)(i - i0) & 3;
    pack |= (uint32_t)s << (8 * k);
    if (k == 3) {
      *reinterpret_cast<uint32_t*>(sym + i - 3) = pack;
      pack = 0;
    }
  }
  const int rem = (int)(i1 - i0) & 3;
  for (int k = 0; k < rem; ++k) sym[i1 - rem + k] = (uint8_t)(pack >> (8 * k));
}

// Staged form: the bit stream of the block's kDecSegs consecutive segments is copied into shared memory first
// (coalesced), and every thread decodes its segment from there, storing 16 symbols at a time.  The plain kernel above
// reads the stream word by word from global memory with one word of look-ahead: a refill every ~13 symbols that
// costs an L2 / DRAM round trip the two or three resident warps per SM cannot hide (measured: 233 ns per symbol on
// the 5 M-symbol mask stream, 239 us for a kernel that moves 8 MB).  Blocks whose window does not fit (barely
// compressible data) take the plain path.
constexpr int kDecSegs = 64;
constexpr int kDecWinWords = 24 * 1024;        // 96 KB of bit stream per block

__global__ void __launch_bounds__(kDecSegs)
huff_decode_staged_kernel(const uint32_t* __restrict__ words, long long n_words, long long n, int seg,
                          const unsigned long long* __restrict__ seg_bits, const uint16_t* __restrict__ lut,
                          int lut_bits, const int32_t* __restrict__ tree, uint8_t* __restrict__ sym) {
  extern __shared__ uint32_t s_dec[];
  uint32_t* s_words = s_dec;
  uint16_t* s_lut = reinterpret_cast<uint16_t*>(s_dec + kDecWinWords);
  for (int i = threadIdx.x; i < (1 << lut_bits); i += blockDim.x) s_lut[i] = lut[i];
  const long long nseg = (n + seg - 1) / seg;
  const long long sg0 = (long long)blockIdx.x * kDecSegs;
  const long long sg = sg0 + threadIdx.x;
  // the block's window of the stream: from the first word of its first segment to the last word any of its segments
  // can touch (the start of the next block's first segment, or the end of the buffer) plus the look-ahead words
  long long w0 = (long long)(seg_bits[sg0] >> 5);
  long long w1 = (sg0 + kDecSegs < nseg) ? (long long)(seg_bits[sg0 + kDecSegs] >> 5) + 3 : n_words;
  if (w0 > n_words - 3) w0 = 0;
  if (w1 > n_words) w1 = n_words;
  const long long nw = w1 - w0;
  const bool staged = nw > 0 && nw <= kDecWinWords;
  if (staged)
    for (long long i = threadIdx.x; i < nw; i += blockDim.x) s_words[i] = __ldg(words + w0 + i);
  __syncthreads();
  if (sg >= nseg) return;
  unsigned long long pos = seg_bits[sg];
  if ((long long)(pos >> 5) > n_words - 3 || (long long)(pos >> 5) < w0) pos = (unsigned long long)w0 << 5;
  const uint32_t* base = staged ? s_words : words + w0;
  const long long last = nw - 1;
  long long wi = (long long)(pos >> 5) - w0;
  if (wi > last - 2) wi = last - 2 < 0 ? 0 : last - 2;
  unsigned long long buf = ((unsigned long long)__byte_perm(base[wi], 0, 0x0123) << 32) | __byte_perm(base[wi + 1], 0, 0x0123);
  wi += 2;
  uint32_t ahead = base[wi];
  int used = (int)(pos & 31);
  const long long i0 = sg * seg;
  const long long i1 = (i0 + seg < n) ? i0 + seg : n;
  const bool one_symbol = (tree[0] < 0 && tree[1] < 0 && tree[0] == tree[1]);
  unsigned long long plo = 0, phi = 0;                  // 16 decoded symbols, stored as one 16-byte vector
  for (long long i = i0; i < i1; ++i) {
    const unsigned long long win = buf << used;          // >= 33 valid bits at the top
    const uint16_t e = s_lut[win >> (64 - lut_bits)];
    int l = e >> 8;
    int sy = e & 0xff;
    if (l == 0) {   // code longer than lut_bits (or the empty code of a one-symbol alphabet)
      if (one_symbol) {
        sy = ~tree[0];
      } else {
        int node = 0;
        while (l < 32) {
          const int bit = (int)((win >> (63 - l)) & 1);
          ++l;
          const int nx = tree[2 * node + bit];
          if (nx < 0) { sy = ~nx; break; }
          node = nx;
        }
      }
    }
    used += l;
    if (used >= 32) {
      buf = (buf << 32) | __byte_perm(ahead, 0, 0x0123);
      used -= 32;
      wi = (wi < last) ? wi + 1 : wi;                    // saturates: a damaged stream decodes to garbage, never faults
      ahead = base[wi];
    }
    const int k = (int)(i - i0) & 15;
    const unsigned long long v = (unsigned long long)sy << (8 * (k & 7));
    if (k < 8) plo |= v; else phi |= v;
    if (k == 15) {
      *reinterpret_cast<uint4*>(sym + i - 15) =
          make_uint4((uint32_t)plo, (uint32_t)(plo >> 32), (uint32_t)phi, (uint32_t)(phi >> 32));
      plo = phi = 0;
    }
  }
  const int rem = (int)(i1 - i0) & 15;
  for (int k = 0; k < rem; ++k) sym[i1 - rem + k] = (uint8_t)((k < 8 ? plo : phi) >> (8 * (k & 7)));
}

inline unsigned grid_for(long long n, int per_block) {
  long long b = (n + per_block - 1) / per_block;
  if (b > 148 * 32) b = 148 * 32;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace

extern "C" int lc_hist_u8(const uint8_t* d_sym, int64_t n, uint32_t* d_hist, void* stream) {
  LC_REQUIRE(d_hist, "null argument");
  cudaStream_t s = as_stream(stream);
  LC_CUDA(cudaMemsetAsync(d_hist, 0, 256 * sizeof(uint32_t), s));
  if (n <= 0) return LC_OK;
  LC_REQUIRE(d_sym, "null argument");
  hist_kernel<<<grid_for(n, 256 * 64), 256, 0, s>>>(d_sym, n, d_hist);
  LC_LAUNCH_CHECK();
  return LC_OK;
}

extern "C" int lc_hist_codes(const int32_t* d_q, int64_t n, int esc, int bias0, int bias1, uint32_t* d_hist,
                             void* stream) {
  LC_REQUIRE(d_hist, "null argument");
  LC_REQUIRE(esc >= 1 && esc <= 255, "escape symbol out of range");
  cudaStream_t s = as_stream(stream);
  LC_CUDA(cudaMemsetAsync(d_hist, 0, 512 * sizeof(uint32_t), s));
  if (n <= 0) return LC_OK;
  LC_REQUIRE(d_q, "null argument");
  hist_codes_kernel<<<grid_for(n, 256 * 16), 256, 0, s>>>(d_q, n, esc, bias0, bias1, d_hist);
  LC_LAUNCH_CHECK();
  return LC_OK;
}

extern "C" int lc_pack_trits(const int8_t* d_mask, int64_t n, uint8_t* d_sym, void* stream) {
  if (n <= 0) return LC_OK;
  LC_REQUIRE(d_mask && d_sym, "null argument");
  pack_trits_kernel<<<grid_for((n + 4) / 5, 256), 256, 0, as_stream(stream)>>>(d_mask, n, d_sym);
  LC_LAUNCH_CHECK();
  return LC_OK;
}

extern "C" int lc_pack_trits_hist(const int8_t* d_mask, int64_t n, uint8_t* d_sym, uint32_t* d_hist, void* stream) {
  LC_REQUIRE(d_hist, "null argument");
  cudaStream_t s = as_stream(stream);
  LC_CUDA(cudaMemsetAsync(d_hist, 0, 256 * sizeof(uint32_t), s));
  if (n <= 0) return LC_OK;
  LC_REQUIRE(d_mask && d_sym, "null argument");
  pack_trits_hist_kernel<<<grid_for((n + 79) / 80, 256), 256, 0, s>>>(d_mask, n, d_sym, d_hist);
  LC_LAUNCH_CHECK();
  return LC_OK;
}

extern "C" int lc_unpack_trits(const uint8_t* d_sym, int64_t n, int8_t* d_mask, void* stream) {
  if (n <= 0) return LC_OK;
  LC_REQUIRE(d_mask && d_sym, "null argument");
  unpack_trits_kernel<<<grid_for((n + 4) / 5, 256), 256, 0, as_stream(stream)>>>(d_sym, n, d_mask);
  LC_LAUNCH_CHECK();
  return LC_OK;
}

extern "C" size_t lc_rle_workspace_bytes(int64_t n) {
  const long long nt = ((n < 1 ? 1 : n) + kRleTile - 1) / kRleTile;
  const long long ne = ((n < 1 ? 1 : n) + kThreads * kExpItems - 1) / (kThreads * kExpItems);
  return ws_bytes(nt > ne ? nt : ne);
}

extern "C" int lc_rle_tokens(const uint8_t* d_sym, int64_t n, int run_sym, int run_first, uint8_t* d_tok,
                             unsigned long long* d_n_tok, uint32_t* d_hist, uint32_t* d_hist_in, void* d_ws,
                             void* stream) {
  LC_REQUIRE(d_sym && d_tok && d_n_tok && d_hist && d_ws, "null argument");
  LC_REQUIRE(n > 0, "empty input");
  LC_REQUIRE(run_sym >= 0 && run_sym < 256 && run_first >= 0 && run_first + 4 < 256 &&
                 (run_sym < run_first || run_sym > run_first + 4), "bad run symbols");
  cudaStream_t s = as_stream(stream);
  const long long nt = (n + kRleTile - 1) / kRleTile;
  LC_CUDA(cudaMemsetAsync(d_ws, 0, ws_bytes(nt), s));
  LC_CUDA(cudaMemsetAsync(d_hist, 0, 256 * sizeof(uint32_t), s));
  LC_CUDA(cudaMemsetAsync(d_n_tok, 0, sizeof(unsigned long long), s));
  if (d_hist_in) LC_CUDA(cudaMemsetAsync(d_hist_in, 0, 256 * sizeof(uint32_t), s));
  rle_tokens_kernel<<<(unsigned)nt, kThreads, 0, s>>>(d_sym, n, run_sym, run_first, d_tok, d_n_tok, d_hist, d_hist_in,
                                                      reinterpret_cast<unsigned long long*>(d_ws));
  LC_LAUNCH_CHECK();
  return LC_OK;
}

extern "C" int lc_rle_expand(const uint8_t* d_tok, int64_t n_tok, int run_sym, int run_first, uint8_t* d_sym,
                             int64_t n, void* d_ws, void* stream) {
  LC_REQUIRE(d_sym && d_ws && n > 0, "bad argument");
  LC_REQUIRE(run_sym >= 0 && run_sym < 256 && run_first >= 0 && run_first + 4 < 256, "bad run symbols");
  cudaStream_t s = as_stream(stream);
  LC_CUDA(cudaMemsetAsync(d_sym, run_sym, (size_t)n, s));
  if (n_tok <= 0) return LC_OK;
  LC_REQUIRE(d_tok, "null argument");
  const long long nt = (n_tok + kThreads * kExpItems - 1) / (kThreads * kExpItems);
  LC_CUDA(cudaMemsetAsync(d_ws, 0, ws_bytes(nt), s));
  rle_expand_kernel<<<(unsigned)nt, kThreads, 0, s>>>(d_tok, n_tok, run_sym, run_first, d_sym, n,
                                                      reinterpret_cast<unsigned long long*>(d_ws));
  LC_LAUNCH_CHECK();
  return LC_OK;
}

extern "C" int lc_symbolize_i32(const int32_t* d_q, int64_t n, int diff, int bias, int esc, uint8_t* d_sym,
                                int32_t* d_esc, unsigned long long* d_nesc, void* d_ws, void* stream) {
  LC_REQUIRE(d_nesc && d_ws, "null argument");
  cudaStream_t s = as_stream(stream);
  LC_CUDA(cudaMemsetAsync(d_nesc, 0, sizeof(unsigned long long), s));
  if (n <= 0) return LC_OK;
  LC_REQUIRE(d_q && d_sym && d_esc, "null argument");
  LC_REQUIRE(esc >= 1 && esc <= 255, "escape symbol out of range");
  const long long nt = (n + kTile - 1) / kTile;
  LC_CUDA(cudaMemsetAsync(d_ws, 0, ws_bytes(nt), s));
  const bool vec = (reinterpret_cast<uintptr_t>(d_q) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_sym) & 7) == 0;
  symbolize_kernel<<<(unsigned)nt, kThreads, 0, s>>>(d_q, n, diff, bias, esc, d_sym, d_esc, d_nesc,
                                                     reinterpret_cast<unsigned long long*>(d_ws), vec);
  LC_LAUNCH_CHECK();
  return LC_OK;
}

extern "C" int lc_symbolize_plain_i32(const int32_t* d_q, int64_t n, int diff, int bias, uint8_t* d_sym, void* stream) {
  if (n <= 0) return LC_OK;
  LC_REQUIRE(d_q && d_sym, "null argument");
  const bool vec = ((reinterpret_cast<uintptr_t>(d_q) | reinterpret_cast<uintptr_t>(d_sym)) & 15) == 0;
  symbolize_plain_kernel<<<grid_for((n + 15) / 16, 256), 256, 0, as_stream(stream)>>>(d_q, n, diff, bias, d_sym, vec);
  LC_LAUNCH_CHECK();
  return LC_OK;
}

extern "C" int lc_unsymbolize_i32(const uint8_t* d_sym, int64_t n, int diff, int bias, int esc,
                                  const int32_t* d_esc, int32_t* d_q, void* d_ws, void* stream) {
  if (n <= 0) return LC_OK;
  LC_REQUIRE(d_q && d_sym && d_ws, "null argument");
  cudaStream_t s = as_stream(stream);
  if (d_esc == nullptr) {      // the container holds no escapes: element-wise map, no scan
    const bool vec = ((reinterpret_cast<uintptr_t>(d_q) | reinterpret_cast<uintptr_t>(d_sym)) & 15) == 0;
    unsymbolize_plain_kernel<<<grid_for((n + 15) / 16, 256), 256, 0, s>>>(d_sym, n, bias, d_q, vec);
    LC_LAUNCH_CHECK();
    if (diff) return lc_cumsum_i32(d_q, n, d_q, d_ws, stream);
    return LC_OK;
  }
  const long long nt = (n + kTile - 1) / kTile;
  LC_CUDA(cudaMemsetAsync(d_ws, 0, ws_bytes(nt), s));
  const bool vec = (reinterpret_cast<uintptr_t>(d_q) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_sym) & 7) == 0;
  unsymbolize_kernel<<<(unsigned)nt, kThreads, 0, s>>>(d_sym, n, bias, esc, d_esc, d_q,
                                                       reinterpret_cast<unsigned long long*>(d_ws), vec);
  LC_LAUNCH_CHECK();
  if (diff) return lc_cumsum_i32(d_q, n, d_q, d_ws, stream);
  return LC_OK;
}

extern "C" size_t lc_huff_max_bytes(int64_t n) { return (size_t)(n < 0 ? 0 : n) * 4 + 64; }

extern "C" size_t lc_huff_workspace_bytes(int64_t n, int seg) {
  (void)seg;
  const long long nt = ((n < 1 ? 1 : n) + kTileSyms - 1) / kTileSyms;
  return ws_bytes(nt);
}

extern "C" int lc_huff_encode(const uint8_t* d_sym, int64_t n, const uint32_t* d_code, const uint8_t* d_len,
                              int seg, uint8_t* d_out, unsigned long long* d_seg_bits,
                              unsigned long long* d_total_bits, void* d_ws, void* stream) {
  LC_REQUIRE(d_total_bits && d_ws, "null argument");
  LC_REQUIRE(seg >= 64 && seg <= kTileSyms && (seg & (seg - 1)) == 0, "segment length must be a power of two in 64..4096");
  cudaStream_t s = as_stream(stream);
  LC_CUDA(cudaMemsetAsync(d_total_bits, 0, sizeof(unsigned long long), s));
  if (n <= 0) return LC_OK;
  LC_REQUIRE(d_sym && d_code && d_len && d_out && d_seg_bits, "null argument");
  LC_REQUIRE((reinterpret_cast<uintptr_t>(d_out) & 3) == 0, "output must be 4-byte aligned");
  const long long nt = (n + kTileSyms - 1) / kTileSyms;
  LC_CUDA(cudaMemsetAsync(d_ws, 0, ws_bytes(nt), s));
  huff_encode_kernel<<<(unsigned)nt, kThreads, 0, s>>>(d_sym, n, d_code, d_len,
                                                       reinterpret_cast<uint32_t*>(d_out), d_seg_bits,
                                                       d_total_bits,
                                                       reinterpret_cast<unsigned long long*>(d_ws), seg / kSegItems);
  LC_LAUNCH_CHECK();
  return LC_OK;
}

extern "C" int lc_huff_decode(const uint8_t* d_bits, int64_t n_bytes, int64_t n, int seg,
                              const unsigned long long* d_seg_bits, const uint16_t* d_lut, int lut_bits,
                              const int32_t* d_tree, uint8_t* d_sym, void* stream) {
  if (n <= 0) return LC_OK;
  LC_REQUIRE(d_bits && d_seg_bits && d_lut && d_tree && d_sym, "null argument");
  LC_REQUIRE(lut_bits >= 1 && lut_bits <= 14 && seg > 0, "bad decode table");
  LC_REQUIRE(n_bytes >= 16 && n_bytes % 4 == 0, "the bit buffer must be a multiple of 4 bytes with 16 spare bytes");
  const long long nseg = (n + seg - 1) / seg;
  // the caller guarantees the bit buffer is readable up to the last byte holding a code bit and
  // zero padded to 8 more bytes; pass the conservative size bound
  LC_REQUIRE((reinterpret_cast<uintptr_t>(d_bits) & 3) == 0 && (reinterpret_cast<uintptr_t>(d_sym) & 3) == 0 &&
                 seg % 4 == 0, "bit stream and symbol buffers must be 4-byte aligned");
  if (seg % 16 == 0 && (reinterpret_cast<uintptr_t>(d_sym) & 15) == 0 && !getenv("LOSSYCOMP_PLAIN_DECODE")) {
    const size_t smem = (size_t)kDecWinWords * 4 + (size_t)(1 << lut_bits) * 2;
    static bool attr_done[64] = {};
    int dev = 0;
    LC_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
      LC_CUDA(cudaFuncSetAttribute(huff_decode_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    huff_decode_staged_kernel<<<(unsigned)((nseg + kDecSegs - 1) / kDecSegs), kDecSegs, smem, as_stream(stream)>>>(
        reinterpret_cast<const uint32_t*>(d_bits), n_bytes / 4, n, seg, d_seg_bits, d_lut, lut_bits, d_tree, d_sym);
    LC_LAUNCH_CHECK();
    return LC_OK;
  }
  huff_decode_kernel<<<(unsigned)((nseg + 127) / 128), 128, (size_t)(1 << lut_bits) * 2,
                       as_stream(stream)>>>(reinterpret_cast<const uint32_t*>(d_bits), n_bytes / 4, n, seg, d_seg_bits,
                                            d_lut, lut_bits, d_tree, d_sym);
  LC_LAUNCH_CHECK();
  return LC_OK;
}
