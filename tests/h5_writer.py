"""TEST INFRASTRUCTURE ONLY: writes small HDF5 files laid out like a DSEC sequence's ``events.h5`` /
``rectify_map.h5`` so that ``cmda_b200.h5lite`` (and ``store_io.convert_dsec_h5`` on top of it) can be exercised
without h5py, which is not in this image.  The writer follows the HDF5 File Format Specification independently of
the reader (version 0 superblock, version 1 object headers, symbol-table groups with a version 1 B-tree + local
heap, chunked datasets indexed by a version 1 B-tree, filter pipeline message version 1) -- the structures h5py
emits with its default ``libver='earliest'`` -- and frames Blosc chunks the way c-blosc 1.x does (16-byte header,
block start table, byte shuffle, one compressed stream per byte of the type).
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


# ------------------------------------------------------------------------------------------------ Blosc-1 frames
def blosc_compress(data: bytes, typesize: int, cname: str = "zstd", shuffle: bool = True, blocksize: int = 4096) -> bytes:
    import pyarrow as pa
    code = {"blosclz": 0, "lz4": 1, "snappy": 2, "zlib": 3, "zstd": 4}[cname]
    nbytes = len(data)
    flags = (1 if shuffle and typesize > 1 else 0) | (code << 5)
    nblocks = max(1, (nbytes + blocksize - 1) // blocksize)
    header = 16 + 4 * nblocks
    body, bstarts = bytearray(), []
    for bi in range(nblocks):
        block = data[bi * blocksize:(bi + 1) * blocksize]
        bsize = len(block)
        if flags & 1:
            nel = bsize // typesize
            block = np.frombuffer(block[:nel * typesize], np.uint8).reshape(nel, typesize).T.tobytes() + block[nel * typesize:]
        split = typesize <= 16 and blocksize // typesize >= 128 and bsize == blocksize
        nstreams = typesize if split else 1
        neblock = bsize // nstreams
        bstarts.append(header + len(body))
        for k in range(nstreams):
            stream = block[k * neblock:(k + 1) * neblock]
            if code == 3:
                comp = zlib.compress(stream, 5)
            elif code == 0:
                comp = stream                                   # stored: blosclz streams are optional in a frame
            else:
                comp = pa.compress(stream, codec={1: "lz4_raw", 2: "snappy", 4: "zstd"}[code], asbytes=True)
            if len(comp) >= len(stream):
                comp = stream                                   # incompressible stream: stored as is (csize == neblock)
            body += struct.pack("<i", len(comp)) + comp
    cbytes = header + len(body)
    head = struct.pack("<BBBBIII", 2, 1, flags, typesize, nbytes, blocksize, cbytes)
    return head + struct.pack(f"<{nblocks}i", *bstarts) + bytes(body)


def blosc_memcpy_frame(data: bytes, typesize: int) -> bytes:
    return struct.pack("<BBBBIII", 2, 1, 0x02, typesize, len(data), len(data), len(data) + 16) + data


# ------------------------------------------------------------------------------------------------ HDF5 structures
class _Out:
    def __init__(self):
        self.buf = bytearray()

    def tell(self):
        return len(self.buf)

    def align(self, n=8):
        self.buf += b"\0" * ((-len(self.buf)) % n)

    def write(self, b):
        a = len(self.buf)
        self.buf += b
        return a

    def patch(self, addr, b):
        self.buf[addr:addr + len(b)] = b


def _msg(mtype, body):
    body = body + b"\0" * ((-len(body)) % 8)
    return struct.pack("<HHB3x", mtype, len(body), 0) + body


def _object_header(msgs):
    data = b"".join(msgs)
    return struct.pack("<BxHII4x", 1, len(msgs), 1, len(data)) + data


def _datatype(dt: np.dtype):
    dt = np.dtype(dt)
    if dt.kind in "ui":
        bits = 0x08 if dt.kind == "i" else 0
        return struct.pack("<BBBBI", 0x10 | 0, bits, 0, 0, dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "f" and dt.itemsize == 4:
        return struct.pack("<BBBBI", 0x10 | 1, 0x20, 31, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    if dt.kind == "f" and dt.itemsize == 8:
        return struct.pack("<BBBBI", 0x10 | 1, 0x20, 63, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    raise NotImplementedError(dt)


def _dataspace(shape):
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(s)) for s in shape)


def _filter_pipeline(filters):
    body = struct.pack("<BB6x", 1, len(filters))
    for fid, name, cd in filters:
        nm = name.encode() + b"\0"
        nm += b"\0" * ((-len(nm)) % 8)
        body += struct.pack("<HHHH", fid, len(nm), 1, len(cd)) + nm + struct.pack(f"<{len(cd)}I", *cd)
        if len(cd) % 2:
            body += b"\0" * 4
    return body


def _write_dataset(out: _Out, arr: np.ndarray, chunks=None, filters=(), cname="zstd"):
    """Returns the address of the dataset's object header."""
    arr = np.asarray(arr, order="C")                           # (ascontiguousarray would turn a scalar into shape (1,))
    dt = arr.dtype.newbyteorder("<")
    msgs = [_msg(0x0001, _dataspace(arr.shape)), _msg(0x0003, _datatype(dt))]
    if chunks is None:
        out.align()
        addr = out.write(arr.astype(dt).tobytes()) if arr.size else UNDEF
        msgs.append(_msg(0x0008, struct.pack("<BBQQ", 3, 1, addr, arr.nbytes)))
    else:
        rank = arr.ndim
        entries = []                                            # (chunk offsets, address, stored size)
        grid = [range(0, arr.shape[d], chunks[d]) for d in range(rank)]
        for offs in np.stack(np.meshgrid(*grid, indexing="ij"), -1).reshape(-1, rank):
            block = np.zeros(chunks, dtype=dt)
            sl = tuple(slice(int(o), min(int(o) + c, s)) for o, c, s in zip(offs, chunks, arr.shape))
            block[tuple(slice(0, s.stop - s.start) for s in sl)] = arr[sl]
            raw = block.tobytes()
            for fid, _, _ in filters:
                if fid == 2:
                    raw = np.frombuffer(raw, np.uint8).reshape(-1, dt.itemsize).T.tobytes()
                elif fid == 1:
                    raw = zlib.compress(raw, 4)
                elif fid == 32001:
                    raw = blosc_compress(raw, dt.itemsize, cname=cname) if cname != "memcpy" else blosc_memcpy_frame(raw, dt.itemsize)
            out.align()
            entries.append((tuple(int(o) for o in offs), out.write(raw), len(raw)))
        # one leaf node (level 0) holds every chunk; the final key carries the dataset's extent
        out.align()
        node = bytearray(b"TREE" + struct.pack("<BBHQQ", 1, 0, len(entries), UNDEF, UNDEF))
        for offs, addr, size in entries:
            node += struct.pack("<II", size, 0) + b"".join(struct.pack("<Q", o) for o in offs) + struct.pack("<Q", 0)
            node += struct.pack("<Q", addr)
        node += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", int(s)) for s in arr.shape) + struct.pack("<Q", 0)
        btree = out.write(bytes(node))
        layout = struct.pack("<BBB", 3, 2, rank + 1) + struct.pack("<Q", btree) + \
            b"".join(struct.pack("<I", int(c)) for c in chunks) + struct.pack("<I", dt.itemsize)
        msgs.append(_msg(0x0008, layout))
        if filters:
            msgs.append(_msg(0x000B, _filter_pipeline(filters)))
    out.align()
    return out.write(_object_header(msgs))


def _write_group(out: _Out, children: dict):
    """children: name -> object header address.  Returns (object header address, B-tree address, heap address)."""
    names = sorted(children)
    heap = bytearray(b"\0" * 8)                                 # offset 0: the empty name
    offsets = {}
    for n in names:
        offsets[n] = len(heap)
        heap += n.encode() + b"\0"
        heap += b"\0" * ((-len(heap)) % 8)
    out.align()
    data_addr = out.write(bytes(heap))
    out.align()
    heap_addr = out.write(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), UNDEF, data_addr))
    out.align()
    snod = bytearray(b"SNOD" + struct.pack("<BxH", 1, len(names)))
    for n in names:
        snod += struct.pack("<QQII16x", offsets[n], children[n], 0, 0)
    snod_addr = out.write(bytes(snod))
    out.align()
    tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod_addr, offsets[names[-1]] if names else 0)
    btree_addr = out.write(tree)
    out.align()
    hdr = out.write(_object_header([_msg(0x0011, struct.pack("<QQ", btree_addr, heap_addr))]))
    return hdr, btree_addr, heap_addr


def write_h5(path: str, tree: dict):
    """``tree``: name -> numpy array | dict (sub-group) | (array, dict(chunks=..., filters=[...], cname=...))."""
    out = _Out()
    out.write(b"\0" * 96)                                       # superblock (56 bytes) + root symbol table entry (40)

    def emit(node):
        children = {}
        for name, val in node.items():
            if isinstance(val, dict):
                children[name] = emit(val)[0]
            elif isinstance(val, tuple):
                arr, opt = val
                filters = [{"deflate": (1, "deflate", (4,)), "shuffle": (2, "shuffle", (np.dtype(arr.dtype).itemsize,)),
                            "blosc": (32001, "blosc", (2, 2, np.dtype(arr.dtype).itemsize, 0, 5, 1, 4))}[f] for f in opt.get("filters", [])]
                children[name] = _write_dataset(out, np.asarray(arr), opt.get("chunks"), filters, opt.get("cname", "zstd"))
            else:
                children[name] = _write_dataset(out, np.asarray(val))
        return _write_group(out, children)

    root_hdr, root_btree, root_heap = emit(tree)
    eof = out.tell()
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0) + \
        struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF) + struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", root_btree, root_heap)
    assert len(sb) == 96
    out.patch(0, sb)
    with open(path, "wb") as fh:
        fh.write(out.buf)
