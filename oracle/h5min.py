"""Minimal HDF5 reader/writer for the files on the Sayram-2D hot path boundary.

TEST INFRASTRUCTURE (oracle side).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg may import this module.

The reference loads its diffusion table through xtensor-io/HighFive/libhdf5
(reference: source/Cases/Albert_Young_IO.cc:17-36); none of those exist in this
image, so this module parses the subset of the HDF5 file format that
D/AlbertYoung_chorus.h5 uses: superblock v0, v1 group B-tree + local heap +
symbol-table nodes, v1 object headers, contiguous layout (v3 layout message),
little-endian IEEE f64 fixed-point-free datasets, no filters.

It walks the symbol table instead of hard-coding byte offsets.
"""
from __future__ import annotations

import struct
import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(RuntimeError):
    pass


def _u(fmt, buf, off):
    return struct.unpack_from("<" + fmt, buf, off)


class H5File:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.b = fh.read()
        b = self.b
        if b[:8] != _SIG:
            raise H5FormatError("not an HDF5 file: " + str(path))
        if b[8] != 0:
            raise H5FormatError("only superblock version 0 is supported")
        if b[13] != 8 or b[14] != 8:
            raise H5FormatError("only 8-byte offsets/lengths are supported")
        self.base = _u("Q", b, 24)[0]
        # root group symbol table entry starts at byte 56
        _, _, cache_type = _u("QQI", b, 56)
        if cache_type != 1:
            raise H5FormatError("root group has no cached B-tree/heap addresses")
        btree, heap = _u("QQ", b, 56 + 24)
        self.datasets = {}
        self._walk_group(btree, heap, "")

    # -- group structures -------------------------------------------------
    def _heap_data(self, heap_addr):
        b = self.b
        if b[heap_addr:heap_addr + 4] != b"HEAP":
            raise H5FormatError("bad local heap signature")
        _size, _free, data_addr = _u("QQQ", b, heap_addr + 8)
        return data_addr

    def _name(self, heap_data, off):
        b = self.b
        s = heap_data + off
        e = b.index(b"\x00", s)
        return b[s:e].decode("ascii")

    def _walk_group(self, btree_addr, heap_addr, prefix):
        heap_data = self._heap_data(heap_addr)
        self._walk_btree(btree_addr, heap_data, prefix)

    def _walk_btree(self, addr, heap_data, prefix):
        b = self.b
        if b[addr:addr + 4] != b"TREE":
            raise H5FormatError("bad B-tree signature")
        node_type, level, nent = _u("BBH", b, addr + 4)
        if node_type != 0:
            raise H5FormatError("expected a group B-tree")
        # keys and children interleaved after sig(4)+type/level/n(4)+2 siblings(16)
        p = addr + 24
        for k in range(nent):
            child = _u("Q", b, p + 8 + k * 16)[0]
            if level > 0:
                self._walk_btree(child, heap_data, prefix)
            else:
                self._walk_snod(child, heap_data, prefix)

    def _walk_snod(self, addr, heap_data, prefix):
        b = self.b
        if b[addr:addr + 4] != b"SNOD":
            raise H5FormatError("bad symbol-table node signature")
        nsym = _u("H", b, addr + 6)[0]
        for k in range(nsym):
            e = addr + 8 + k * 40
            name_off, ohdr, cache_type = _u("QQI", b, e)
            name = prefix + "/" + self._name(heap_data, name_off)
            if cache_type == 1:  # sub-group with cached addresses
                bt, hp = _u("QQ", b, e + 24)
                self._walk_group(bt, hp, name)
            else:
                info = self._object_header(ohdr)
                if info is not None:
                    self.datasets[name] = info
                elif "stab" in self._last_msgs:
                    bt, hp = self._last_msgs["stab"]
                    self._walk_group(bt, hp, name)

    # -- object header ----------------------------------------------------
    def _object_header(self, addr):
        b = self.b
        version = b[addr]
        if version != 1:
            raise H5FormatError("only version-1 object headers are supported")
        nmsg = _u("H", b, addr + 2)[0]
        hsize = _u("I", b, addr + 8)[0]
        blocks = [(addr + 16, hsize)]
        msgs = {}
        seen = 0
        while blocks and seen < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and seen < nmsg:
                mtype, msize, _flags = _u("HHB", b, p)
                body = p + 8
                seen += 1
                if mtype == 0x0001:  # dataspace
                    ver, rank = b[body], b[body + 1]
                    if ver == 1:
                        dims = _u("%dQ" % rank, b, body + 8)
                    elif ver == 2:
                        dims = _u("%dQ" % rank, b, body + 4)
                    else:
                        raise H5FormatError("dataspace version")
                    msgs["shape"] = tuple(int(d) for d in dims)
                elif mtype == 0x0003:  # datatype
                    cls = b[body] & 0x0F
                    bits0 = b[body + 1]
                    size_b = _u("I", b, body + 4)[0]
                    msgs["dtype"] = (cls, bits0 & 1, size_b)
                elif mtype == 0x0008:  # data layout
                    ver = b[body]
                    if ver != 3:
                        raise H5FormatError("only layout message v3 supported")
                    lclass = b[body + 1]
                    if lclass != 1:
                        raise H5FormatError("only contiguous layout supported")
                    daddr, dsize = _u("QQ", b, body + 2)
                    msgs["data"] = (daddr, dsize)
                elif mtype == 0x000B:
                    raise H5FormatError("filtered datasets are not supported")
                elif mtype == 0x0010:  # continuation
                    caddr, clen = _u("QQ", b, body)
                    blocks.append((caddr, clen))
                elif mtype == 0x0011:  # symbol table message (a group)
                    msgs["stab"] = _u("QQ", b, body)
                p = body + msize
        self._last_msgs = msgs
        if "data" not in msgs:
            return None
        return msgs

    def read(self, name):
        if not name.startswith("/"):
            name = "/" + name
        if name not in self.datasets:
            raise KeyError(name)
        m = self.datasets[name]
        cls, big_endian, size_b = m["dtype"]
        if cls != 1 or size_b != 8 or big_endian:
            raise H5FormatError("only little-endian f64 datasets are supported")
        daddr, dsize = m["data"]
        if daddr == _UNDEF:
            raise H5FormatError("dataset has no storage")
        n = int(np.prod(m["shape"])) if m["shape"] else 1
        arr = np.frombuffer(self.b, dtype="<f8", count=n, offset=self.base + daddr)
        return arr.reshape(m["shape"]).copy()


def load_d_table(path):
    """alpha0[deg], E[MeV], Daa, Dap, Dpp exactly as Albert_Young_IO.cc:21-35 reads them."""
    f = H5File(path)
    return {k: f.read("/" + k) for k in ("alpha0", "E", "Daa", "Dap", "Dpp")}
