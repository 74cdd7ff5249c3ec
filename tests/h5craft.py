"""Hand-packed HDF5 files for the reader tests (tests only; no libhdf5 / h5py in this image).

An INDEPENDENT writer of the on-disk structures h5lite::File claims to read, following the HDF5 file format
specification (version 3.0): superblock v0 with old-style groups or superblock v2 / v3 with version-2 object headers and
compact link messages; contiguous, compact and chunked (version-1 B-tree, node type 1) layouts; the filter pipeline
message (v1 / v2) with shuffle, deflate (Python's zlib) and fletcher32; IEEE f64 / f32 and fixed-point datatypes.
Checksums of the v2 structures are written as zero (the reader does not verify them)."""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
DTYPES = {"f8": (1, 8, 0x20, 0x3F), "f4": (1, 4, 0x20, 0x1F), "i4": (0, 4, 0x08, 0), "u2": (0, 2, 0x00, 0), "i8": (0, 8, 0x08, 0), "u1": (0, 1, 0, 0)}


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def msg_dataspace(shape):
    return 0x0001, struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", s) for s in shape)


def msg_datatype(code):
    cls, size, bits0, bits1 = DTYPES[code]
    head = struct.pack("<BBBBI", 0x10 | cls, bits0, bits1, 0, size)
    if cls == 1:   # bit offset, precision, exponent location / size, mantissa location / size, bias
        prop = struct.pack("<HHBBBBI", 0, 8 * size, 52 if size == 8 else 23, 11 if size == 8 else 8, 0, 52 if size == 8 else 23, 1023 if size == 8 else 127)
    else:
        prop = struct.pack("<HH", 0, 8 * size)
    return 0x0003, head + prop


def msg_filters(filters, version):
    out = struct.pack("<BB", version, len(filters)) + (b"\0" * 6 if version == 1 else b"")
    for fid, cd in filters:
        if version == 1:
            out += struct.pack("<HHHH", fid, 0, 1, len(cd)) + b"".join(struct.pack("<I", c) for c in cd) + (b"\0" * 4 if len(cd) % 2 else b"")
        else:
            out += struct.pack("<HHH", fid, 1, len(cd)) + b"".join(struct.pack("<I", c) for c in cd)
    return 0x000B, out


def header_v1(messages):
    body = b""
    for t, m in messages:
        m = _pad8(m)
        body += struct.pack("<HHB3x", t, len(m), 0) + m
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


def header_v2(messages, track_order=False):
    body = b""
    for t, m in messages:
        body += struct.pack("<BHB", t, len(m), 0) + (struct.pack("<H", 0) if track_order else b"") + m
    flags = 0x02 | (0x04 if track_order else 0)      # 4-byte chunk-0 size
    return b"OHDR" + struct.pack("<BBI", 2, flags, len(body)) + body + b"\0\0\0\0"


class Crafter:
    """Accumulates the file image; `place` appends a blob at an 8-byte boundary and returns its address."""

    def __init__(self, reserve):
        self.img = bytearray(reserve)

    def place(self, blob):
        self.img += b"\0" * (-len(self.img) % 8)
        a = len(self.img)
        self.img += blob
        return a

    def dataset_messages(self, arr, code, layout, chunk=None, filters=(), pipeline_version=1):
        arr = np.ascontiguousarray(arr.astype("<" + code))
        msgs = [msg_dataspace(arr.shape), msg_datatype(code)]
        es = arr.dtype.itemsize
        if layout == "contiguous":
            a = self.place(arr.tobytes())
            msgs.append((0x0008, struct.pack("<BBQQ", 3, 1, a, arr.nbytes)))
        elif layout == "compact":
            msgs.append((0x0008, struct.pack("<BBH", 3, 0, arr.nbytes) + arr.tobytes()))
        else:
            rank = arr.ndim
            grid = [-(-s // c) for s, c in zip(arr.shape, chunk)]
            entries = []
            for idx in np.ndindex(*grid):
                off = [i * c for i, c in zip(idx, chunk)]
                block = np.zeros(chunk, dtype=arr.dtype)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(off, chunk, arr.shape))
                block[tuple(slice(0, s.stop - s.start) for s in sl)] = arr[sl]
                raw = block.tobytes()
                for fid, cd in filters:
                    if fid == 2:
                        n = len(raw) // es
                        raw = np.frombuffer(raw, dtype=np.uint8).reshape(n, es).T.tobytes()
                    elif fid == 1:
                        raw = zlib.compress(raw, cd[0] if cd else 6)
                    elif fid == 3:
                        raw = raw + b"\xde\xad\xbe\xef"
                entries.append((len(raw), off, self.place(raw)))
            # one leaf node per 3 chunks, one level-1 node above them when there is more than one leaf
            def node(level, items):    # items: (size, offsets, child address)
                key = lambda size, off: struct.pack("<II", size, 0) + b"".join(struct.pack("<Q", o) for o in off) + struct.pack("<Q", 0)
                b = b"TREE" + struct.pack("<BBHQQ", 1, level, len(items), UNDEF, UNDEF)
                for size, off, child in items:
                    b += key(size, off) + struct.pack("<Q", child)
                b += key(0, [s for s in arr.shape])
                return self.place(b)
            leaves = [entries[k:k + 3] for k in range(0, len(entries), 3)]
            if len(leaves) == 1:
                root = node(0, leaves[0])
            else:
                root = node(1, [(lf[0][0], lf[0][1], node(0, lf)) for lf in leaves])
            if filters:
                msgs.append(msg_filters(filters, pipeline_version))
            msgs.append((0x0008, struct.pack("<BBBQ", 3, 2, rank + 1, root) + b"".join(struct.pack("<I", c) for c in chunk) + struct.pack("<I", es)))
        return msgs


def write_old_style(path, datasets):
    """Superblock v0, one root group with a v1 B-tree / local heap / one symbol-table node, v1 object headers.
    datasets: name -> dict(data=, code=, layout=, chunk=, filters=, pipeline_version=)."""
    c = Crafter(96)
    names = sorted(datasets)
    heap_data = b"\0" * 8
    name_off = {}
    for n in names:
        name_off[n] = len(heap_data)
        heap_data += _pad8(n.encode() + b"\0")
    hdr = {n: c.place(header_v1(c.dataset_messages(d["data"], d["code"], d["layout"], d.get("chunk"), d.get("filters", ()), d.get("pipeline_version", 1))))
           for n, d in datasets.items()}
    heap_data_at = c.place(heap_data)
    heap = c.place(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, heap_data_at))
    snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
    for n in names:
        snod += struct.pack("<QQII16x", name_off[n], hdr[n], 0, 0)
    snod_at = c.place(snod)
    tree = c.place(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod_at, name_off[names[-1]]))
    root_hdr = c.place(header_v1([(0x0011, struct.pack("<QQ", tree, heap))]))
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, len(c.img), UNDEF)
    sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", tree, heap)
    c.img[:len(sb)] = sb
    open(path, "wb").write(bytes(c.img))


def write_new_style(path, tree, superblock_version=3, track_order=False):
    """Superblock v2 / v3, version-2 object headers, groups as compact link messages (the second group's links sit in a
    continuation block).  tree: nested dict, leaves = dict(data=, code=, layout=, ...)."""
    c = Crafter(48)

    def link(name, addr):
        return 0x0006, struct.pack("<BBB", 1, 0x00, len(name)) + name.encode() + struct.pack("<Q", addr)

    def put(node, use_continuation):
        if "data" in node:
            return c.place(header_v2(c.dataset_messages(node["data"], node["code"], node["layout"], node.get("chunk"), node.get("filters", ()), node.get("pipeline_version", 2)),
                                     track_order))
        links = [link(n, put(child, True)) for n, child in node.items()]
        info = (0x0002, struct.pack("<BBQQ", 0, 0, UNDEF, UNDEF))
        if use_continuation and len(links) > 1:
            blob = b"OCHK"
            for t, m in links[1:]:
                blob += struct.pack("<BHB", t, len(m), 0) + (struct.pack("<H", 0) if track_order else b"") + m
            blob += b"\0\0\0\0"
            at = c.place(blob)
            return c.place(header_v2([info, links[0], (0x0010, struct.pack("<QQ", at, len(blob)))], track_order))
        return c.place(header_v2([info] + links, track_order))

    root = put(tree, False)
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBQQQQI", superblock_version, 8, 8, 0, 0, UNDEF, len(c.img), root, 0)
    c.img[:len(sb)] = sb
    open(path, "wb").write(bytes(c.img))
