"""Host-side Huffman table builder for the GPU bit packer.

Mirrors the reference's lossycomp/huffman.py: the same (non-canonical) tree construction and
tie-breaking -- leaves ordered by (count, symbol) (huffman.py:71), the queue stably re-sorted by
count before each merge, the two smallest popped (first = left = '0'), the parent appended at
the end (huffman.py:20-33), code = root-to-leaf path (huffman.py:35-45) -- so that the bit
stream the GPU emits for a symbol sequence equals the '0101...' string `getHuffmanCode` returns
for the same sequence.  `getHuffmanCode` / `decode_huffman` keep the reference's names and
return conventions for string input; `build_table` is the integer-symbol entry the codec uses.
"""
from __future__ import annotations

from typing import Dict, Hashable, List, Sequence, Tuple

import numpy as np

MAX_CODE_BITS = 32      # the GPU packer takes codes of at most 32 bits
LUT_BITS = 12           # direct decode table; longer codes walk the tree


def _tree_code_ints(counts: Sequence[int]) -> Tuple[List[int], List[int]]:
    """counts: leaf counts already ordered by (count, symbol).  Returns (code, length) per leaf.

    The reference re-sorts its queue stably by count before every merge and appends the new parent
    at the end (huffman.py:20-33).  A stable sort keeps equal counts in queue order and a parent is
    younger than everything already queued, so the queue order is exactly (count, creation index):
    a heap on that key pops the same two nodes without the O(n^2) re-sorting.  First popped = left
    = bit 0 (huffman.py:24-27, 40-43)."""
    import heapq
    n = len(counts)
    heap = [(int(counts[i]), i) for i in range(n)]      # already sorted => already a heap
    par = [0] * (2 * n)
    bit = [0] * (2 * n)
    nxt = n
    while len(heap) > 1:
        fa, a = heapq.heappop(heap)
        fb, b = heapq.heappop(heap)
        par[a] = nxt
        par[b] = nxt
        bit[b] = 1
        heapq.heappush(heap, (fa + fb, nxt))
        nxt += 1
    root = nxt - 1
    code = [0] * nxt
    length = [0] * nxt
    for node in range(root - 1, -1, -1):                # parents are always younger (larger id)
        p = par[node]
        code[node] = (code[p] << 1) | bit[node]
        length[node] = length[p] + 1
    if n == 1:
        return [0], [0]
    return code[:n], length[:n]


def _tree_codes(leaves: Sequence[Tuple[Hashable, int]]) -> Dict[Hashable, str]:
    """leaves: [(symbol, count)] ordered by (count, symbol) -> {symbol: '0101' code string}."""
    code, length = _tree_code_ints([c for _, c in leaves])
    return {sym: (format(code[i], "b").zfill(length[i]) if length[i] else "")
            for i, (sym, _) in enumerate(leaves)}


def getHuffmanCode(string):
    """Reference signature (huffman.py:63): returns [bit string, {char: code}]."""
    freqs: Dict[str, int] = {}
    for ch in string:
        freqs[ch] = freqs.get(ch, 0) + 1
    leaves = sorted(freqs.items(), key=lambda kv: (kv[1], kv[0]))
    table = _tree_codes(leaves)
    return ["".join(table[ch] for ch in string), table]


def decode_huffman(input_string, char_store, freq_store):
    """Reference signature (huffman.py:48): greedy prefix match of `input_string` against the codes."""
    inv = {code: ch for ch, code in zip(char_store, freq_store)}
    out, cur = [], ""
    for b in input_string:
        cur += b
        if cur in inv:
            out.append(inv[cur])
            cur = ""
    return "".join(out)


class Table:
    """A prefix code over byte symbols in the arrays the C-ABI wants."""

    def __init__(self, codes: Dict[int, str] = None, reference_tree: bool = False, arrays=None):
        self.reference_tree = reference_tree
        self.code = np.zeros(256, dtype=np.uint32)
        self.len = np.zeros(256, dtype=np.uint8)
        self.used = np.zeros(256, dtype=bool)
        if arrays is not None:
            syms, code, length = arrays
            assert max(length) <= MAX_CODE_BITS
            self.code[syms] = code
            self.len[syms] = length
            self.used[syms] = True
        else:
            for s, bits in codes.items():
                assert len(bits) <= MAX_CODE_BITS
                self.code[s] = int(bits, 2) if bits else 0
                self.len[s] = len(bits)
                self.used[s] = True

    @property
    def codes(self) -> Dict[int, str]:
        return {int(s): (format(int(self.code[s]), "b").zfill(int(self.len[s])) if self.len[s] else "")
                for s in np.nonzero(self.used)[0]}

    # -- decode side ---------------------------------------------------------
    def decode_arrays(self, lut_bits: int = LUT_BITS):
        lut = np.zeros(1 << lut_bits, dtype=np.uint16)
        nodes: List[List[int]] = [[0, 0]]
        codes = self.codes
        if len(codes) == 1:
            (s,) = codes.keys()
            return lut, np.array([~s, ~s], dtype=np.int32)
        for s, bits in codes.items():
            L = len(bits)
            if L <= lut_bits:
                base = int(bits, 2) << (lut_bits - L)
                lut[base:base + (1 << (lut_bits - L))] = (L << 8) | s
            node = 0
            for i, b in enumerate(bits):
                bi = int(b)
                if i == L - 1:
                    nodes[node][bi] = ~s
                else:
                    if nodes[node][bi] <= 0:
                        nodes.append([0, 0])
                        nodes[node][bi] = len(nodes) - 1
                    node = nodes[node][bi]
        return lut, np.array(nodes, dtype=np.int32).reshape(-1)

    def decode_arrays_native(self, lut_bits: int = LUT_BITS):
        """decode_arrays() through the library's host entry point lc_huff_build_decode (same arrays)."""
        import ctypes as C

        from . import _native as N
        lut = np.empty(1 << lut_bits, dtype=np.uint16)
        tree = np.empty(1022, dtype=np.int32)
        used = self.used.astype(np.uint8)
        n = C.c_int(0)
        N.check(N.lib().lc_huff_build_decode(self.code.ctypes.data, self.len.ctypes.data, used.ctypes.data, lut_bits,
                                             lut.ctypes.data, tree.ctypes.data, C.byref(n)), "lc_huff_build_decode")
        return lut, tree[:n.value]

    # -- serialisation (explicit codes: the tree is not canonical) -----------
    _REC = np.dtype([("s", "u1"), ("l", "u1"), ("c", "<u4")])

    def to_bytes(self) -> bytes:
        syms = np.nonzero(self.used)[0]
        rec = np.empty(syms.size, dtype=self._REC)
        rec["s"], rec["l"], rec["c"] = syms, self.len[syms], self.code[syms]
        return bytes([syms.size & 0xFF, syms.size >> 8]) + rec.tobytes()

    @staticmethod
    def from_bytes(b) -> Tuple["Table", int]:
        n = b[0] | (b[1] << 8)
        rec = np.frombuffer(b, dtype=Table._REC, count=n, offset=2)
        return Table(arrays=(rec["s"].astype(np.int64), rec["c"], rec["l"])), 2 + 6 * n

    def encoded_bits(self, hist) -> int:
        return int((np.asarray(hist, dtype=np.int64) * self.len.astype(np.int64)).sum())


def build_table_native(hist) -> Table:
    """build_table() through the library's host entry point lc_huff_build_table (what the codec calls:
    three tables per compress() cost 0.4 ms in Python, microseconds in C); same result bit for bit."""
    import ctypes as C

    from . import _native as N
    h = np.ascontiguousarray(hist, dtype=np.int64)
    assert h.size == 256
    t = Table({}, False)
    used = np.zeros(256, dtype=np.uint8)
    ref = C.c_int(0)
    N.check(N.lib().lc_huff_build_table(h.ctypes.data, MAX_CODE_BITS, t.code.ctypes.data, t.len.ctypes.data,
                                        used.ctypes.data, C.byref(ref)), "lc_huff_build_table")
    t.used = used.astype(bool)
    t.reference_tree = bool(ref.value)
    return t


def build_table(hist) -> Table:
    """hist: 256 counts.  Reference tree when its depth fits the packer, otherwise the same
    construction on counts flattened just enough to cap the depth at MAX_CODE_BITS."""
    hist = np.asarray(hist, dtype=np.int64)
    shift = 0
    while True:
        h = hist if shift == 0 else np.where(hist > 0, np.maximum(hist >> shift, 1), 0)
        syms = np.nonzero(h)[0]
        if syms.size == 0:
            syms = np.array([0])
            cnt = np.array([1])
        else:
            cnt = h[syms]
        order = np.lexsort((syms, cnt))                 # by (count, symbol), huffman.py:71
        syms, cnt = syms[order], cnt[order]
        code, length = _tree_code_ints(cnt.tolist())
        if max(length) <= MAX_CODE_BITS:
            return Table(reference_tree=(shift == 0), arrays=(syms, code, length))
        shift += 1
