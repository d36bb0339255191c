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

// Decode-side arrays of a prefix code (what lc_huff_decode consumes): a 2^lut_bits direct table of (len << 8 | sym)
// for codes no longer than lut_bits, and a binary tree (node = {child0, child1}, leaves are ~sym, 0 = absent) that
// holds every code.  A one-symbol alphabet (empty code) is the tree {~sym, ~sym}.  Returns the number of int32
// entries written to `tree` (<= 2 * 511) through n_tree.
extern "C" int lc_huff_build_decode(const uint32_t* code, const uint8_t* len, const uint8_t* used, int lut_bits,
                                    uint16_t* lut, int32_t* tree, int* n_tree) {
  LC_REQUIRE(code && len && used && lut && tree && n_tree, "null argument");
  LC_REQUIRE(lut_bits >= 1 && lut_bits <= 16, "lut_bits must be in 1..16");
  for (int i = 0; i < (1 << lut_bits); ++i) lut[i] = 0;
  int n_used = 0, only = 0;
  for (int s = 0; s < kMaxSyms; ++s)
    if (used[s]) { ++n_used; only = s; }
  if (n_used <= 1) {
    tree[0] = ~only;
    tree[1] = ~only;
    *n_tree = 2;
    return LC_OK;
  }
  int nodes = 1;
  tree[0] = 0;
  tree[1] = 0;
  for (int s = 0; s < kMaxSyms; ++s) {
    if (!used[s]) continue;
    const int L = len[s];
    LC_REQUIRE(L >= 1 && L <= 32, "code length out of range");
    const uint32_t c = code[s];
    LC_REQUIRE(L == 32 || (c >> L) == 0, "code has bits above its length");
    if (L <= lut_bits) {
      const uint32_t base = c << (lut_bits - L);
      for (uint32_t k = 0; k < (1u << (lut_bits - L)); ++k) lut[base + k] = (uint16_t)((L << 8) | s);
    }
    int node = 0;
    for (int i = 0; i < L; ++i) {
      const int bit = (int)((c >> (L - 1 - i)) & 1u);
      if (i == L - 1) {
        LC_REQUIRE(tree[2 * node + bit] == 0, "not a prefix code (a code repeats or extends another)");
        tree[2 * node + bit] = ~s;
      } else {
        LC_REQUIRE(tree[2 * node + bit] >= 0, "not a prefix code (a code extends another)");
        if (tree[2 * node + bit] == 0) {
          LC_REQUIRE(nodes < kMaxSyms - 1, "not a prefix code (more inner nodes than a 256-leaf tree has)");
          tree[2 * nodes] = 0;
          tree[2 * nodes + 1] = 0;
          tree[2 * node + bit] = nodes++;
        }
        node = tree[2 * node + bit];
      }
    }
  }
  // a Huffman code is complete: every inner node has two children.  The decoder relies on it (every bit
  // pattern reaches a leaf within 32 steps, whatever a corrupted bit stream contains).
  for (int i = 0; i < 2 * nodes; ++i) LC_REQUIRE(tree[i] != 0, "incomplete prefix code");
  *n_tree = 2 * nodes;
  return LC_OK;
}
