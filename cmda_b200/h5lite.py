"""A minimal HDF5 reader for the two files of a DSEC sequence the event path opens -- ``events.h5``
(``events/{t,x,y,p}``, ``ms_to_idx``, ``t_offset``) and ``rectify_map.h5`` (``rectify_map``) -- the way the
reference opens them with h5py + hdf5plugin (mmseg/datasets/dsec.py:3-4, 287-291, 342-345;
create_dsec_dataset_txt.py:14-18).  Neither package is in this image and there is no network, so the subset of
the HDF5 file format those files use is read directly (HDF5 File Format Specification, version 3.0):

  * superblock version 0 - 3; version 1 object headers (+ continuation blocks) and version 2 ("OHDR") headers;
  * old-style groups (symbol table message -> version 1 B-tree + local heap + "SNOD" nodes, what h5py writes by
    default) and compact new-style groups (link messages in the object header);
  * datasets: simple / scalar dataspaces, little-endian fixed-point and IEEE float datatypes, compact, contiguous
    and chunked (version 1 B-tree) layouts, layout message version 3;
  * the filter pipeline: deflate (1), shuffle (2) and Blosc (32001, the filter of hdf5plugin that DSEC's files are
    written with).  A Blosc chunk is a self-describing Blosc-1 frame: byte shuffle + split streams compressed with
    zstd / lz4 / zlib / snappy (decoded with pyarrow's codecs and zlib) or blosclz (decoded here).

Read-only, whole-chunk granular: ``ds[a:b]`` on a 1-D dataset touches only the chunks the slice overlaps, which is
how ``store_io.convert_dsec_h5`` streams a 400 M-event sequence into the decoded cache.  This is file I/O, not
part of the timed path; it exists so that the cache builder has actually run on the reference's on-disk format.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

__all__ = ["File", "Dataset", "Group", "blosc_decompress"]

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


# --------------------------------------------------------------------------------------------- Blosc-1 frames
def _blosclz_decompress(src: bytes, n_out: int) -> bytes:
    """blosclz (FastLZ level-1 family) block decoder: literal runs and (length, distance) matches."""
    out = bytearray()
    ip, n = 0, len(src)
    ctrl = src[ip] & 31
    ip += 1
    while True:
        if ctrl >= 32:
            length = (ctrl >> 5) - 1
            ofs = (ctrl & 31) << 8
            if length == 7 - 1:
                while True:
                    code = src[ip]
                    ip += 1
                    length += code
                    if code != 255:
                        break
            code = src[ip]
            ip += 1
            length += 3
            ref = len(out) - ofs - code - 1
            if code == 255 and ofs == (31 << 8):       # 16-bit distance extension
                ofs = (src[ip] << 8) | src[ip + 1]
                ip += 2
                ref = len(out) - ofs - 8191 - 1
            if ref < 0:
                raise ValueError("corrupt blosclz stream")
            for _ in range(length):                    # overlapping copies are the point of the format
                out.append(out[ref])
                ref += 1
        else:
            run = ctrl + 1
            out += src[ip:ip + run]
            ip += run
        if ip >= n or len(out) >= n_out:
            break
        ctrl = src[ip]
        ip += 1
    return bytes(out[:n_out])


def _codec_decompress(code: int, data: bytes, n_out: int) -> bytes:
    if code == 3:
        return zlib.decompress(data)
    if code == 0:
        return _blosclz_decompress(data, n_out)
    import pyarrow as pa
    name = {1: "lz4_raw", 2: "snappy", 4: "zstd"}.get(code)
    if name is None:
        raise NotImplementedError(f"Blosc compressor code {code}")
    return pa.decompress(data, decompressed_size=n_out, codec=name).to_pybytes()


def blosc_decompress(frame: bytes) -> bytes:
    """One Blosc-1 frame (what the HDF5 Blosc filter stores per chunk) -> the chunk's bytes."""
    flags, typesize = frame[2], frame[3]
    nbytes, blocksize, cbytes = struct.unpack_from("<III", frame, 4)
    if flags & 0x02:                                        # memcpyed: stored, not compressed
        return bytes(frame[16:16 + nbytes])
    if flags & 0x04:
        raise NotImplementedError("Blosc bit-shuffle")
    code = flags >> 5
    nblocks = (nbytes + blocksize - 1) // blocksize
    bstarts = struct.unpack_from(f"<{nblocks}i", frame, 16)
    out = bytearray(nbytes)
    for bi in range(nblocks):
        bsize = min(blocksize, nbytes - bi * blocksize)
        leftover = bsize != blocksize
        split = not (flags & 0x10) and typesize <= 16 and blocksize // typesize >= 128 and not leftover
        nstreams = typesize if split else 1
        neblock = bsize // nstreams
        ip = bstarts[bi]
        block = bytearray()
        for _ in range(nstreams):
            (csz,) = struct.unpack_from("<i", frame, ip)
            ip += 4
            block += frame[ip:ip + csz] if csz == neblock else _codec_decompress(code, bytes(frame[ip:ip + csz]), neblock)
            ip += csz
        if flags & 0x01 and typesize > 1:                   # byte shuffle: typesize planes of bsize // typesize bytes
            nel = bsize // typesize
            body = np.frombuffer(bytes(block[:nel * typesize]), dtype=np.uint8).reshape(typesize, nel).T.tobytes()
            block = bytearray(body) + block[nel * typesize:]
        out[bi * blocksize:bi * blocksize + bsize] = block
    return bytes(out)


def _unshuffle(data: bytes, itemsize: int) -> bytes:
    """HDF5's own shuffle filter (id 2)."""
    n = len(data) // itemsize
    if itemsize <= 1 or n == 0:
        return data
    body = np.frombuffer(data[:n * itemsize], dtype=np.uint8).reshape(itemsize, n).T.tobytes()
    return body + data[n * itemsize:]


# --------------------------------------------------------------------------------------------- file structures
class _Reader:
    def __init__(self, fh):
        self.fh = fh
        self.so = self.sl = 8

    def at(self, addr: int, n: int) -> bytes:
        self.fh.seek(addr)
        b = self.fh.read(n)
        if len(b) != n:
            raise EOFError(f"short read at {addr}")
        return b

    def off(self, buf, pos):
        return int.from_bytes(buf[pos:pos + self.so], "little")

    def length(self, buf, pos):
        return int.from_bytes(buf[pos:pos + self.sl], "little")


def _parse_messages_v1(r: _Reader, addr: int):
    hdr = r.at(addr, 16)
    if hdr[0] != 1:
        raise NotImplementedError(f"object header version {hdr[0]}")
    n_msgs, size = struct.unpack_from("<H", hdr, 2)[0], struct.unpack_from("<I", hdr, 8)[0]
    blocks = [(addr + 16, size)]
    msgs = []
    while blocks and len(msgs) < n_msgs:
        a, n = blocks.pop(0)
        buf = r.at(a, n)
        pos = 0
        while pos + 8 <= n and len(msgs) < n_msgs:
            mtype, msize, _flags = struct.unpack_from("<HHB", buf, pos)
            body = buf[pos + 8:pos + 8 + msize]
            pos += 8 + msize
            if mtype == 0x0010:                             # continuation
                blocks.append((r.off(body, 0), r.length(body, r.so)))
            msgs.append((mtype, body))
    return msgs


def _parse_messages_v2(r: _Reader, addr: int):
    head = r.at(addr, 6)
    flags = head[5]
    pos = addr + 6
    if flags & 0x20:
        pos += 16                                           # four timestamps
    if flags & 0x10:
        pos += 4                                            # max compact / min dense attributes
    nb = 1 << (flags & 3)
    size = int.from_bytes(r.at(pos, nb), "little")
    pos += nb
    msgs = []
    blocks = [(pos, size)]
    track = bool(flags & 0x04)
    while blocks:
        a, n = blocks.pop(0)
        buf = r.at(a, n)
        p = 0
        while p + 4 <= n:
            mtype, msize, _mflags = buf[p], struct.unpack_from("<H", buf, p + 1)[0], buf[p + 3]
            p += 4 + (2 if track else 0)
            body = buf[p:p + msize]
            p += msize
            if mtype == 0x10:
                ca, cl = r.off(body, 0), r.length(body, r.so)
                blocks.append((ca + 4, cl - 8))             # skip "OCHK", drop the checksum
            elif mtype != 0:
                msgs.append((mtype, body))
    return msgs


def _messages(r: _Reader, addr: int):
    return _parse_messages_v2(r, addr) if r.at(addr, 4) == b"OHDR" else _parse_messages_v1(r, addr)


def _dtype_of(body: bytes) -> np.dtype:
    cls, bits0 = body[0] & 0x0F, body[1]
    size = struct.unpack_from("<I", body, 4)[0]
    if bits0 & 1:
        raise NotImplementedError("big-endian datatype")
    if cls == 0:
        return np.dtype(("<i" if bits0 & 0x08 else "<u") + str(size))
    if cls == 1:
        return np.dtype("<f" + str(size))
    raise NotImplementedError(f"datatype class {cls}")


def _shape_of(body: bytes):
    ver, rank = body[0], body[1]
    pos = 8 if ver == 1 else 4
    return tuple(int.from_bytes(body[pos + 8 * i:pos + 8 * i + 8], "little") for i in range(rank))


def _filters_of(body: bytes):
    ver, n = body[0], body[1]
    pos = 8 if ver == 1 else 2
    out = []
    for _ in range(n):
        fid = struct.unpack_from("<H", body, pos)[0]
        pos += 2
        name_len = 0
        if ver == 1 or fid >= 256:
            name_len = struct.unpack_from("<H", body, pos)[0]
            pos += 2
        _flags, ncd = struct.unpack_from("<HH", body, pos)
        pos += 4
        if name_len:
            pos += (name_len + 7) // 8 * 8 if ver == 1 else name_len
        cd = struct.unpack_from(f"<{ncd}I", body, pos)
        pos += 4 * ncd
        if ver == 1 and ncd % 2:
            pos += 4
        out.append((fid, cd))
    return out


class Dataset:
    def __init__(self, r: _Reader, addr: int, name: str):
        self._r, self.name = r, name
        self._filters, self._layout = [], None
        self.shape, self.dtype = (), None
        for mtype, body in _messages(r, addr):
            if mtype == 0x0001:
                self.shape = _shape_of(body)
            elif mtype == 0x0003:
                self.dtype = _dtype_of(body)
            elif mtype == 0x000B:
                self._filters = _filters_of(body)
            elif mtype == 0x0008:
                self._layout = self._parse_layout(body)
        if self.dtype is None or self._layout is None:
            raise ValueError(f"{name}: not a dataset")
        self.size = int(np.prod(self.shape)) if self.shape else 1
        self._chunks = None

    def _parse_layout(self, body):
        ver, cls = body[0], body[1]
        if ver != 3:
            raise NotImplementedError(f"data layout message version {ver}")
        r = self._r
        if cls == 0:
            n = struct.unpack_from("<H", body, 2)[0]
            return ("compact", bytes(body[4:4 + n]))
        if cls == 1:
            return ("contiguous", r.off(body, 2), r.length(body, 2 + r.so))
        nd = body[2]
        btree = r.off(body, 3)
        dims = struct.unpack_from(f"<{nd}I", body, 3 + r.so)
        return ("chunked", btree, tuple(dims[:-1]))

    def __len__(self):
        return self.shape[0]

    # ---- chunk index: version 1 B-tree, node type 1 --------------------------------------------------
    def _walk(self, addr, rank, out):
        r = self._r
        head = r.at(addr, 8 + 2 * r.so)
        if head[:4] != b"TREE" or head[4] != 1:
            raise ValueError("corrupt chunk B-tree")
        level, used = head[5], struct.unpack_from("<H", head, 6)[0]
        key = 8 + 8 * (rank + 1)
        body = r.at(addr + 8 + 2 * r.so, used * (key + r.so) + key)
        for i in range(used):
            p = i * (key + r.so)
            nbytes, mask = struct.unpack_from("<II", body, p)
            offs = struct.unpack_from(f"<{rank}Q", body, p + 8)
            child = r.off(body, p + key)
            if level:
                self._walk(child, rank, out)
            else:
                out.append((offs, child, nbytes, mask))

    def _chunk_list(self):
        if self._chunks is None:
            out = []
            if self._layout[1] != _UNDEF:
                self._walk(self._layout[1], len(self.shape), out)
            self._chunks = sorted(out)
        return self._chunks

    def _decode(self, raw: bytes, mask: int) -> bytes:
        for k in range(len(self._filters) - 1, -1, -1):    # the pipeline, in reverse
            if mask & (1 << k):
                continue
            fid = self._filters[k][0]
            if fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:
                raw = _unshuffle(raw, self.dtype.itemsize)
            elif fid == 32001:
                raw = blosc_decompress(raw)
            else:
                raise NotImplementedError(f"HDF5 filter {fid}")
        return raw

    def _read_all(self) -> np.ndarray:
        kind = self._layout[0]
        if kind == "compact":
            return np.frombuffer(self._layout[1], dtype=self.dtype, count=self.size).reshape(self.shape).copy()
        if kind == "contiguous":
            if self._layout[1] == _UNDEF:
                return np.zeros(self.shape, self.dtype)
            raw = self._r.at(self._layout[1], self.size * self.dtype.itemsize)
            return np.frombuffer(raw, dtype=self.dtype).reshape(self.shape).copy()
        return self._read_rows(0, self.shape[0] if self.shape else 1)

    def _read_rows(self, a: int, b: int) -> np.ndarray:
        """Rows [a, b) of the first axis of a chunked dataset: only the chunks that overlap them are read."""
        cshape = self._layout[2]
        out = np.zeros((max(b - a, 0),) + tuple(self.shape[1:]), dtype=self.dtype)
        for offs, addr, nbytes, mask in self._chunk_list():
            if offs[0] >= b or offs[0] + cshape[0] <= a:
                continue
            chunk = np.frombuffer(self._decode(self._r.at(addr, nbytes), mask), dtype=self.dtype,
                                  count=int(np.prod(cshape))).reshape(cshape)
            src = tuple(slice(max(a - offs[0], 0) if d == 0 else 0,
                              min(cshape[d], (b if d == 0 else self.shape[d]) - offs[d])) for d in range(len(cshape)))
            dst = tuple(slice(offs[d] + src[d].start - (a if d == 0 else 0), offs[d] + src[d].stop - (a if d == 0 else 0))
                        for d in range(len(cshape)))
            out[dst] = chunk[src]
        return out

    def __getitem__(self, key):
        if key == () or key is Ellipsis:
            arr = self._read_all()
            return arr[()] if not self.shape else arr
        if isinstance(key, slice) and self.shape and self._layout[0] == "chunked":
            a, b, step = key.indices(self.shape[0])
            if step == 1:
                return self._read_rows(a, b)
        return self._read_all()[key]

    def __array__(self, dtype=None, copy=None):
        arr = self._read_all()
        return arr if dtype is None else arr.astype(dtype)


class Group:
    def __init__(self, r: _Reader, addr: int, name: str):
        self._r, self.name = r, name
        self._links = {}
        for mtype, body in _messages(r, addr):
            if mtype == 0x0011:                             # symbol table: B-tree + local heap
                self._read_symbol_table(r.off(body, 0), r.off(body, r.so))
            elif mtype == 0x0006:                           # link message (compact new-style group)
                self._read_link(body)
            elif mtype == 0x0002:
                raise NotImplementedError("dense link storage (fractal heap)")

    def _read_link(self, body):
        flags = body[1]
        pos = 2
        ltype = 0
        if flags & 0x08:
            ltype = body[pos]
            pos += 1
        if flags & 0x04:
            pos += 8
        if flags & 0x10:
            pos += 1
        nb = 1 << (flags & 3)
        n = int.from_bytes(body[pos:pos + nb], "little")
        pos += nb
        name = body[pos:pos + n].decode()
        pos += n
        if ltype == 0:
            self._links[name] = self._r.off(body, pos)

    def _read_symbol_table(self, btree, heap):
        r = self._r
        h = r.at(heap, 8 + 2 * r.sl + r.so)
        if h[:4] != b"HEAP":
            raise ValueError("corrupt local heap")
        data_addr = r.off(h, 8 + 2 * r.sl)
        data = r.at(data_addr, r.length(h, 8))

        def walk(addr):
            head = r.at(addr, 8 + 2 * r.so)
            if head[:4] != b"TREE" or head[4] != 0:
                raise ValueError("corrupt group B-tree")
            level, used = head[5], struct.unpack_from("<H", head, 6)[0]
            body = r.at(addr + 8 + 2 * r.so, used * (r.sl + r.so) + r.sl)
            for i in range(used):
                child = r.off(body, i * (r.sl + r.so) + r.sl)
                if level:
                    walk(child)
                    continue
                node = r.at(child, 8)
                if node[:4] != b"SNOD":
                    raise ValueError("corrupt symbol table node")
                n = struct.unpack_from("<H", node, 6)[0]
                ents = r.at(child + 8, n * (2 * r.so + 24))
                for k in range(n):
                    e = k * (2 * r.so + 24)
                    name_off, obj = r.off(ents, e), r.off(ents, e + r.so)
                    self._links[data[name_off:data.index(b"\0", name_off)].decode()] = obj

        walk(btree)

    def keys(self):
        return sorted(self._links)

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __getitem__(self, path: str):
        node = self
        parts = [p for p in path.split("/") if p]
        for i, part in enumerate(parts):
            if not isinstance(node, Group) or part not in node._links:
                raise KeyError(path)
            addr = node._links[part]
            full = "/".join(parts[:i + 1])
            kinds = {m for m, _ in _messages(self._r, addr)}
            node = Dataset(self._r, addr, full) if 0x0008 in kinds else Group(self._r, addr, full)
        return node


class File(Group):
    """``h5lite.File(path)['events/t'][a:b]`` -- the h5py calls of the reference, read-only."""

    def __init__(self, path: str, mode: str = "r"):
        if mode != "r":
            raise ValueError("h5lite is read-only")
        self._fh = open(path, "rb")
        r = _Reader(self._fh)
        base = 0
        while True:                                         # the superblock sits at 0, 512, 1024, ...
            self._fh.seek(base)
            if self._fh.read(8) == _SIG:
                break
            base = 512 if base == 0 else base * 2
            if base > 1 << 24:
                raise ValueError(f"{path}: not an HDF5 file")
        ver = r.at(base + 8, 1)[0]
        if ver in (0, 1):
            head = r.at(base + 8, 16)
            r.so, r.sl = head[5], head[6]
            pos = base + 24 + (4 if ver == 1 else 0)        # past the K values and the consistency flags
            pos += 4 * r.so                                 # base, free-space, end-of-file, driver-info addresses
            entry = r.at(pos, 2 * r.so + 24)
            root = r.off(entry, r.so)
        elif ver in (2, 3):
            head = r.at(base + 8, 4)
            r.so, r.sl = head[1], head[2]
            root = r.off(r.at(base + 12 + 3 * r.so, r.so), 0)
        else:
            raise NotImplementedError(f"superblock version {ver}")
        if r.so != 8 or r.sl != 8:
            raise NotImplementedError("offsets / lengths that are not 8 bytes wide")
        super().__init__(r, root, "/")

    def close(self):
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
