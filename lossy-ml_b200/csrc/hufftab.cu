// Host-side Huffman table construction for the GPU bit packer (no device code).
//
// Same tree as the reference's lossycomp/huffman.py: leaves ordered by (count, symbol) (huffman.py:71);
// before every merge the reference re-sorts its queue stably by count and appends the new parent at the
// end (huffman.py:20-33).  A stable sort keeps equal counts in queue order and a parent is younger than
// everything queued, so the queue order is (count, creation index): selecting the two smallest of that
// key reproduces the merges.  First taken = left = bit 0 (huffman.py:24-27, 40-43); a code is the
// root-to-leaf path (huffman.py:35-45).  
#include "common.cuh"

namespace {

constexpr int kMaxSyms = 256;

// lengths (and codes when they fit 32 bits) for `n` leaves already ordered by (count, symbol)
int build_lengths(const unsigned long long* cnt, int n, int* par, int* bit, int* length) {
  if (n == 1) { length[0] = 0; return 0; }
  // two-queue merge: the leaves are sorted and the parents come out in non-decreasing count order, so the
  // smallest (count, index) is at the head of one of the two queues; on equal counts the leaf is older
  unsigned long long key[2 * kMaxSyms];
  for (int i = 0; i < n; ++i) { key[i] = cnt[i]; bit[i] = 0; }
  int nxt = n, leaf = 0, inner = n;
  for (int merges = 0; merges < n - 1; ++merges) {
    int pick[2];
    for (int t = 0; t < 2; ++t)
      pick[t] = (leaf < n && (inner >= nxt || key[leaf] <= key[inner])) ? leaf++ : inner++;
    par[pick[0]] = nxt; bit[pick[0]] = 0;
    par[pick[1]] = nxt; bit[pick[1]] = 1;
    key[nxt] = key[pick[0]] + key[pick[1]];
    bit[nxt] = 0;
    ++nxt;
  }
  const int root = nxt - 1;
  length[root] = 0;
  int deepest = 0;
  for (int node = root - 1; node >= 0; --node) {     // parents are younger (larger index)
    length[node] = length[par[node]] + 1;
    if (node < n && length[node] > deepest) deepest = length[node];
  }
  return deepest;
}

}  // namespace

extern "C" int lc_huff_build_table(const int64_t* hist, int max_bits, uint32_t* code, uint8_t* len,
                                   uint8_t* used, int* reference_tree) {
  LC_REQUIRE(hist && code && len && used && reference_tree, "null argument");
  LC_REQUIRE(max_bits >= 8 && max_bits <= 32, "max_bits must be in 8..32");
  for (int s = 0; s < kMaxSyms; ++s) { code[s] = 0; len[s] = 0; used[s] = 0; }
  for (int shift = 0; shift < 64; ++shift) {
    // counts, flattened just enough (>> shift, floor 1) to cap the depth when the true tree is too deep
    int sym[kMaxSyms];
    unsigned long long cnt[kMaxSyms];
    int n = 0;
    for (int s = 0; s < kMaxSyms; ++s) {
      LC_REQUIRE(hist[s] >= 0, "negative count");
      if (hist[s] > 0) {
        unsigned long long c = (unsigned long long)hist[s] >> shift;
        sym[n] = s;
        cnt[n++] = c ? c : 1ull;
      }
    }
    if (n == 0) { sym[0] = 0; cnt[0] = 1; n = 1; }
    // order by (count, symbol): symbols are already ascending, so a stable insertion sort by count does it
    for (int i = 1; i < n; ++i) {
      const int s = sym[i];
      const unsigned long long c = cnt[i];
      int j = i - 1;
      while (j >= 0 && cnt[j] > c) { sym[j + 1] = sym[j]; cnt[j + 1] = cnt[j]; --j; }
      sym[j + 1] = s; cnt[j + 1] = c;
    }
    int par[2 * kMaxSyms], bit[2 * kMaxSyms], length[2 * kMaxSyms];
    if (build_lengths(cnt, n, par, bit, length) > max_bits) continue;
    unsigned long long path[2 * kMaxSyms];
    const int root = 2 * n - 2;
    path[root] = 0;
    for (int node = root - 1; node >= 0; --node) path[node] = (path[par[node]] << 1) | (unsigned)bit[node];
    for (int i = 0; i < n; ++i) {
      code[sym[i]] = (n == 1) ? 0u : (uint32_t)path[i];
      len[sym[i]] = (uint8_t)((n == 1) ? 0 : length[i]);
      used[sym[i]] = 1;
    }
    *reference_tree = (shift == 0);
    return LC_OK;
  }
  LC_REQUIRE(false, "could not cap the code length");
  return LC_OK;
}
