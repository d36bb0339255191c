"""Minimal pure-Python reader for the HDF5 subset Keras 2.4 writes (superblock v0,
old-style groups, v1 object headers, contiguous little-endian float32 datasets).

h5py is not available in the build image, so the weight converter
(tools/convert_weights.py) uses this instead.  Only what is needed to walk
`/model_weights/<model>/<layer>/{kernel:0,bias:0}` is implemented.
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5File:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        b = self.buf
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise ValueError("not an HDF5 file")
        if b[8] != 0:
            raise ValueError("only superblock v0 supported")
        size_off, size_len = b[13], b[14]
        if size_off != 8 or size_len != 8:
            raise ValueError("only 8-byte offsets/lengths supported")
        # 8 sig + 8 (versions/sizes) + 2+2 (K values) + 4 flags + 4*8 addresses
        root_entry = 8 + 8 + 4 + 4 + 32
        self.root = self._symtab_entry(root_entry)[1]

    # -- low level ---------------------------------------------------------
    def _u(self, off, n):
        return int.from_bytes(self.buf[off:off + n], "little")

    def _symtab_entry(self, off):
        name_off = self._u(off, 8)
        ohdr = self._u(off + 8, 8)
        return name_off, ohdr

    def _messages(self, ohdr):
        """Yield (type, body_offset, size) for a v1 object header, following continuations."""
        b = self.buf
        if b[ohdr] != 1:
            raise ValueError("only v1 object headers supported")
        nmsgs = self._u(ohdr + 2, 2)
        hdr_size = self._u(ohdr + 8, 4)
        blocks = [(ohdr + 16, hdr_size)]
        seen = 0
        while blocks and seen < nmsgs:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and seen < nmsgs:
                mtype = self._u(pos, 2)
                msize = self._u(pos + 2, 2)
                body = pos + 8
                seen += 1
                if mtype == 0x10:  # continuation
                    blocks.append((self._u(body, 8), self._u(body + 8, 8)))
                else:
                    yield mtype, body, msize
                pos = body + msize

    def _group_children(self, ohdr):
        btree = heap = None
        for mtype, body, _ in self._messages(ohdr):
            if mtype == 0x11:
                btree, heap = self._u(body, 8), self._u(body + 8, 8)
        if btree is None:
            return None
        if self.buf[heap:heap + 4] != b"HEAP":
            raise ValueError("bad local heap")
        heap_data = self._u(heap + 24, 8)
        out = {}
        self._walk_btree(btree, heap_data, out)
        return out

    def _walk_btree(self, node, heap_data, out):
        b = self.buf
        if b[node:node + 4] != b"TREE":
            raise ValueError("bad B-tree node")
        level = b[node + 5]
        nent = self._u(node + 6, 2)
        # keys(8) and children(8) interleaved from +24: key0 child0 key1 child1 ...
        for i in range(nent):
            child = self._u(node + 24 + 8 + i * 16, 8)
            if level > 0:
                self._walk_btree(child, heap_data, out)
            else:
                if b[child:child + 4] != b"SNOD":
                    raise ValueError("bad symbol node")
                nsyms = self._u(child + 6, 2)
                for s in range(nsyms):
                    name_off, ohdr = self._symtab_entry(child + 8 + s * 40)
                    p = heap_data + name_off
                    q = b.index(b"\x00", p)
                    out[b[p:q].decode()] = ohdr

    # -- public ------------------------------------------------------------
    def children(self, ohdr=None):
        return self._group_children(self.root if ohdr is None else ohdr)

    def resolve(self, path):
        node = self.root
        for part in [p for p in path.split("/") if p]:
            kids = self._group_children(node)
            if kids is None or part not in kids:
                raise KeyError(path)
            node = kids[part]
        return node

    def dataset(self, ohdr):
        shape = addr = None
        for mtype, body, _ in self._messages(ohdr):
            if mtype == 0x01:  # dataspace v1
                rank = self.buf[body + 1]
                shape = tuple(self._u(body + 8 + 8 * i, 8) for i in range(rank))
            elif mtype == 0x03:  # datatype
                cls = self.buf[body] & 0x0F
                size = self._u(body + 4, 4)
                if cls != 1 or size != 4:
                    raise ValueError("only float32 datasets supported")
            elif mtype == 0x08:  # layout
                if self.buf[body] != 3 or self.buf[body + 1] != 1:
                    raise ValueError("only contiguous v3 layout supported")
                addr = self._u(body + 2, 8)
        if shape is None or addr is None or addr == UNDEF:
            raise ValueError("not a dataset")
        n = int(np.prod(shape)) if shape else 1
        return np.frombuffer(self.buf, "<f4", count=n, offset=addr).reshape(shape).copy()

    def walk_datasets(self, path):
        """Return {relative/path: ndarray} for every dataset below `path`."""
        out = {}

        def rec(node, prefix):
            kids = self._group_children(node)
            if kids is None:
                out[prefix] = self.dataset(node)
                return
            for name, child in kids.items():
                rec(child, f"{prefix}/{name}" if prefix else name)

        rec(self.resolve(path), "")
        return out
