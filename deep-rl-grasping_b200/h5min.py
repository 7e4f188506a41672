"""Minimal HDF5 reader for Keras 2.2.4 weight files (h5py is not a dependency).

Handles exactly what ``keras.Model.save_weights`` / ``model.save`` produced for
/root/reference/encoder_files/*/model.h5: superblock version 0, version-1 object headers, groups as
symbol tables (B-tree ``TREE`` nodes + ``SNOD`` symbol nodes + local heaps), little-endian IEEE float
datasets with CONTIGUOUS layout.  Anything else raises ``NotImplementedError``.
"""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np


class H5File:
    def __init__(self, path: str):
        self.b = open(path, "rb").read()
        if self.b[:8] != b"\x89HDF\r\n\x1a\n":
            raise ValueError("not an HDF5 file")
        if self.b[8] != 0:
            raise NotImplementedError(f"superblock version {self.b[8]}")
        self.so, self.sl = self.b[13], self.b[14]              # size of offsets / lengths
        if (self.so, self.sl) != (8, 8):
            raise NotImplementedError("only 8-byte offsets/lengths")
        # superblock v0: ... base addr, free-space addr, EOF addr, driver addr at 24.., root symbol-table entry at 56
        root = 24 + 4 * 8
        self.root_header = self._u64(root + 8)
        self.root_btree, self.root_heap = self._u64(root + 24), self._u64(root + 32)

    def _u64(self, o): return struct.unpack_from("<Q", self.b, o)[0]
    def _u32(self, o): return struct.unpack_from("<I", self.b, o)[0]
    def _u16(self, o): return struct.unpack_from("<H", self.b, o)[0]

    # ---- groups
    def _heap_data(self, heap_addr):
        assert self.b[heap_addr:heap_addr + 4] == b"HEAP"
        return self._u64(heap_addr + 24)

    def _name(self, heap_addr, off):
        d = self._heap_data(heap_addr) + off
        e = self.b.index(b"\x00", d)
        return self.b[d:e].decode()

    def _btree_entries(self, addr, heap):
        assert self.b[addr:addr + 4] == b"TREE", "bad B-tree node"
        level, n = self.b[addr + 5], self._u16(addr + 6)
        p = addr + 24
        out = {}
        for i in range(n):
            child = self._u64(p + 8)           # key_i (8) then child_i (8)
            p += 16
            if level > 0:
                out.update(self._btree_entries(child, heap))
            else:
                assert self.b[child:child + 4] == b"SNOD"
                ns = self._u16(child + 6)
                q = child + 8
                for _ in range(ns):
                    name_off, hdr = self._u64(q), self._u64(q + 8)
                    cache = self._u32(q + 16)
                    out[self._name(heap, name_off)] = (hdr, (self._u64(q + 24), self._u64(q + 32)) if cache == 1 else None)
                    q += 40
        return out

    def _messages(self, hdr):
        if self.b[hdr] != 1:
            raise NotImplementedError("object header version != 1")
        nmsg, size = self._u16(hdr + 2), self._u32(hdr + 8)
        blocks = [(hdr + 16, size)]
        msgs = []
        while blocks and len(msgs) < nmsg + 64:
            p, sz = blocks.pop(0)
            end = p + sz
            while p + 8 <= end:
                t, s = self._u16(p), self._u16(p + 2)
                body = p + 8
                if t == 0x10:                   # continuation
                    blocks.append((self._u64(body), self._u64(body + 8)))
                else:
                    msgs.append((t, body, s))
                p = body + s
        return msgs

    def _children(self, hdr, cached=None):
        if cached is not None:
            return self._btree_entries(cached[0], cached[1])
        for t, body, _ in self._messages(hdr):
            if t == 0x11:                       # symbol table message
                return self._btree_entries(self._u64(body), self._u64(body + 8))
        return None

    def _dataset(self, hdr):
        shape, dtype, addr = None, None, None
        for t, body, s in self._messages(hdr):
            if t == 0x01:                       # dataspace
                ver, rank = self.b[body], self.b[body + 1]
                o = body + (8 if ver == 1 else 4)
                shape = tuple(self._u64(o + 8 * i) for i in range(rank))
            elif t == 0x03:                     # datatype
                cls, size = self.b[body] & 0x0F, self._u32(body + 4)
                if cls != 1 or (self.b[body + 1] & 1):
                    raise NotImplementedError("only little-endian floating point datasets")
                dtype = {4: np.float32, 8: np.float64}[size]
            elif t == 0x08:                     # layout
                ver = self.b[body]
                if ver == 3:
                    if self.b[body + 1] != 1:
                        raise NotImplementedError("only contiguous layout")
                    addr = self._u64(body + 2)
                elif ver in (1, 2):
                    rank, lcls = self.b[body + 1], self.b[body + 2]
                    if lcls != 1:
                        raise NotImplementedError("only contiguous layout")
                    addr = self._u64(body + 8)
                else:
                    raise NotImplementedError(f"layout version {ver}")
        if shape is None or dtype is None or addr is None:
            return None
        n = int(np.prod(shape)) if shape else 1
        return np.frombuffer(self.b, dtype=dtype, count=n, offset=addr).reshape(shape).copy()

    def datasets(self) -> Dict[str, np.ndarray]:
        """All datasets, keyed by their full path (e.g. 'model_weights/conv2d_1/conv2d_1/kernel:0')."""
        out = {}

        def walk(hdr, cached, prefix):
            ch = self._children(hdr, cached)
            if ch is None:
                d = self._dataset(hdr)
                if d is not None:
                    out[prefix] = d
                return
            for name, (h2, c2) in ch.items():
                walk(h2, c2, f"{prefix}/{name}" if prefix else name)
        walk(self.root_header, (self.root_btree, self.root_heap), "")
        return out


def load_keras_weights(path: str) -> Dict[str, np.ndarray]:
    """{'conv2d_1/kernel': array, ...} from a Keras model.h5 / weights.h5."""
    ds = H5File(path).datasets()
    out = {}
    for k, v in ds.items():
        parts = k.split("/")
        if parts[-1].endswith(":0") and len(parts) >= 2:
            out[parts[-2] + "/" + parts[-1][:-2]] = v
    return out
