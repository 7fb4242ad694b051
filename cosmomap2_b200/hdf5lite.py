"""
Minimal HDF5 reader / writer in NumPy, for the files the reference exchanges through h5py
(utilities/IOfiles.py): the classic on-disk layout libhdf5 1.8 / h5py write by default --
superblock version 0 or 1, version-1 object headers, groups as symbol tables (B-tree v1 + local heap
+ SNOD nodes), integer and IEEE float datasets of either byte order, stored compact, contiguous or
chunked (B-tree v1, optional shuffle / deflate / fletcher32 filters).

``h5py`` is not part of this image; when it is importable ``cosmomap2_b200.IOfiles`` uses it and this
module is only the fallback.  What is not covered (superblock >= 2, version-2 object headers,
compound / variable-length types, attributes) raises ``NotImplementedError`` naming the feature.

Pinned by the reference's own fixtures data/testcase_block_diag_{3,4}.hdf5 (written by h5py through
``write_to_hdf5``, IOfiles.py:277-300; copies under tests/golden/), which the reader decodes and the
writer reproduces structure for structure.
"""
import mmap
import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5Error(IOError):
    pass


# =============================================================================================
# reader
# =============================================================================================
class _Dataset(object):
    def __init__(self, f, name, shape, dtype, layout):
        self._f, self.name, self.shape, self.dtype, self._layout = f, name, tuple(shape), dtype, layout

    def __getitem__(self, key):
        arr = self._read()
        if key is Ellipsis or key == ():
            return arr
        return arr[key]

    def __len__(self):
        return self.shape[0]

    def _read(self):
        f, lay = self._f, self._layout
        n = int(np.prod(self.shape, dtype=np.int64)) if len(self.shape) else 1
        nbytes = n * self.dtype.itemsize
        kind = lay["class"]
        if kind == "compact":
            raw = lay["data"][:nbytes]
        elif kind == "contiguous":
            if lay["address"] == UNDEF:                   # never written: the fill value (zeros)
                raw = b"\0" * nbytes
            else:
                raw = f._buf[lay["address"]:lay["address"] + nbytes]
        else:
            return self._read_chunked()
        out = np.frombuffer(raw, dtype=self.dtype, count=n).reshape(self.shape)
        return out.astype(self.dtype.newbyteorder("="))   # native byte order, as h5py returns it

    def _read_chunked(self):
        f, lay = self._f, self._layout
        rank = len(self.shape)
        cdims = lay["chunk"][:rank]
        out = np.zeros(self.shape, dtype=self.dtype.newbyteorder("="))
        csize = int(np.prod(cdims)) * self.dtype.itemsize
        for offs, addr, size, mask in f._chunks(lay["address"], rank):
            raw = f._buf[addr:addr + size]
            for i, (fid, _name, cvals) in reversed(list(enumerate(lay["filters"]))):
                if mask & (1 << i):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:                             # shuffle: bytes of equal significance together
                    es = cvals[0] if cvals else self.dtype.itemsize
                    a = np.frombuffer(raw, dtype=np.uint8)
                    ne = len(a) // es
                    raw = a[:ne * es].reshape(es, ne).T.tobytes() + a[ne * es:].tobytes()
                elif fid == 3:                             # fletcher32 checksum trailer
                    raw = raw[:-4]
                else:
                    raise NotImplementedError("HDF5 filter id %d (%s) is not supported" % (fid, _name))
            chunk = np.frombuffer(raw[:csize], dtype=self.dtype).reshape(cdims)
            sl_out, sl_in = [], []
            for o, c, s in zip(offs, cdims, self.shape):
                hi = min(o + c, s)
                sl_out.append(slice(o, hi))
                sl_in.append(slice(0, hi - o))
            out[tuple(sl_out)] = chunk[tuple(sl_in)]
        return out


class _Group(object):
    def __init__(self, f, name, links):
        self._f, self.name, self._links = f, name, links

    def keys(self):
        return list(self._links)

    def __contains__(self, key):
        try:
            self[key]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, _Group) or part not in node._links:
                raise KeyError("%s: no object %r" % (node.name, part))
            node = node._f._object(node._links[part], (node.name.rstrip("/") + "/" + part))
        return node


class File(_Group):
    """``File(path)`` -- read-only view with the h5py idioms the reference uses: ``f['a/b'][...]``,
    ``f['grp']['dset'][...]``, ``f.close()``."""

    def __init__(self, filename, mode="r"):
        if mode != "r":
            raise ValueError("hdf5lite.File reads; use hdf5lite.write(...) to create files")
        self._fh = open(filename, "rb")
        try:                                               # mapped, not read: CES files are GBs
            self._buf = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
        except ValueError:                                 # empty file
            self._buf = b""
        self.filename = filename
        self._cache = {}
        root = self._superblock()
        _Group.__init__(self, self, "/", self._object_links(root))

    def close(self):
        if isinstance(self._buf, mmap.mmap):
            self._buf.close()
        self._buf = b""
        if self._fh is not None:
            self._fh.close()
            self._fh = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- low level ----------------------------------------------------------------------------
    def _u(self, off, n):
        return int.from_bytes(self._buf[off:off + n], "little")

    def _superblock(self):
        b = self._buf
        if b[:8] != SIGNATURE:
            raise Hdf5Error("%s is not an HDF5 file" % self.filename)
        ver = b[8]
        if ver > 1:
            raise NotImplementedError("HDF5 superblock version %d (libver='latest' files) needs h5py" % ver)
        if b[13] != 8 or b[14] != 8:
            raise NotImplementedError("HDF5 offsets/lengths of %d/%d bytes" % (b[13], b[14]))
        p = 24 + (4 if ver == 1 else 0)
        self._base = self._u(p, 8)
        root_entry = p + 32
        return self._u(root_entry + 8, 8)                  # object header address of the root group

    def _messages(self, addr):
        """(type, flags, payload bytes) of a version-1 object header, continuations followed."""
        b = self._buf
        if b[addr:addr + 4] == b"OHDR":
            raise NotImplementedError("version-2 object headers (libver='latest' files) need h5py")
        if b[addr] != 1:
            raise Hdf5Error("unsupported object header version %d at 0x%x" % (b[addr], addr))
        nmsg = self._u(addr + 2, 2)
        size = self._u(addr + 8, 4)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = self._u(p, 2), self._u(p + 2, 2), b[p + 4]
                body = b[p + 8:p + 8 + msize]
                if mtype == 0x0010:                        # continuation
                    blocks.append((self._u(p + 8, 8), self._u(p + 16, 8)))
                out.append((mtype, flags, body))
                p += 8 + msize
        return out

    def _heap_string(self, heap_addr, off):
        b = self._buf
        if b[heap_addr:heap_addr + 4] != b"HEAP":
            raise Hdf5Error("bad local heap at 0x%x" % heap_addr)
        data = self._u(heap_addr + 24, 8)
        end = b.find(b"\0", data + off)
        return b[data + off:end].decode("utf-8")

    def _btree_group(self, addr, heap, links):
        b = self._buf
        if b[addr:addr + 4] != b"TREE" or b[addr + 4] != 0:
            raise Hdf5Error("bad group B-tree node at 0x%x" % addr)
        level, used = b[addr + 5], self._u(addr + 6, 2)
        p = addr + 24
        for i in range(used):
            child = self._u(p + 8 + 16 * i, 8)
            if level > 0:
                self._btree_group(child, heap, links)
                continue
            if b[child:child + 4] != b"SNOD":
                raise Hdf5Error("bad symbol table node at 0x%x" % child)
            nsym = self._u(child + 6, 2)
            for k in range(nsym):
                e = child + 8 + 40 * k
                links[self._heap_string(heap, self._u(e, 8))] = self._u(e + 8, 8)

    def _chunks(self, addr, rank):
        """(offsets, address, size, filter mask) of every chunk under a chunk B-tree (node type 1)."""
        b = self._buf
        if addr == UNDEF:
            return
        if b[addr:addr + 4] != b"TREE" or b[addr + 4] != 1:
            raise Hdf5Error("bad chunk B-tree node at 0x%x" % addr)
        level, used = b[addr + 5], self._u(addr + 6, 2)
        ksize = 8 + 8 * (rank + 1)
        p = addr + 24
        for i in range(used):
            key = p + i * (ksize + 8)
            size, mask = self._u(key, 4), self._u(key + 4, 4)
            offs = [self._u(key + 8 + 8 * d, 8) for d in range(rank)]
            child = self._u(key + ksize, 8)
            if level > 0:
                for c in self._chunks(child, rank):
                    yield c
            else:
                yield offs, child, size, mask

    def _object_links(self, addr):
        links = {}
        for mtype, _flags, body in self._messages(addr):
            if mtype == 0x0011:                            # symbol table: B-tree + heap
                bt, heap = int.from_bytes(body[:8], "little"), int.from_bytes(body[8:16], "little")
                self._btree_group(bt, heap, links)
            elif mtype == 0x0006:                          # link message (compact new-style group)
                ver, lf = body[0], body[1]
                p = 2
                ltype = 0
                if lf & 0x08:
                    ltype = body[p]
                    p += 1
                if lf & 0x04:
                    p += 8
                if lf & 0x10:
                    p += 1
                lsz = 1 << (lf & 3)
                ln = int.from_bytes(body[p:p + lsz], "little")
                p += lsz
                name = body[p:p + ln].decode("utf-8")
                p += ln
                if ver == 1 and ltype == 0:
                    links[name] = int.from_bytes(body[p:p + 8], "little")
        return links

    def _object(self, addr, name):
        if addr in self._cache:
            return self._cache[addr]
        msgs = self._messages(addr)
        types = [m[0] for m in msgs]
        if 0x0008 not in types:                            # no data layout: a group
            obj = _Group(self, name, self._object_links(addr))
        else:
            shape, dtype, layout, filters = (), None, None, []
            for mtype, _flags, body in msgs:
                if mtype == 0x0001:
                    shape = _parse_dataspace(body)
                elif mtype == 0x0003:
                    dtype = _parse_datatype(body)
                elif mtype == 0x000B:
                    filters = _parse_filters(body)
                elif mtype == 0x0008:
                    layout = _parse_layout(body)
            if dtype is None or layout is None:
                raise Hdf5Error("%s: dataset without datatype / layout message" % name)
            layout["filters"] = filters
            obj = _Dataset(self, name, shape, dtype, layout)
        self._cache[addr] = obj
        return obj


def _parse_dataspace(body):
    ver, rank, flags = body[0], body[1], body[2]
    if ver == 1:
        p = 8
    elif ver == 2:
        if body[3] == 2:                                   # null dataspace
            return (0,)
        p = 4
    else:
        raise NotImplementedError("dataspace message version %d" % ver)
    return tuple(int.from_bytes(body[p + 8 * i:p + 8 * i + 8], "little") for i in range(rank))


def _parse_datatype(body):
    cls, ver = body[0] & 0x0F, body[0] >> 4
    bits0 = body[1]
    size = int.from_bytes(body[4:8], "little")
    order = ">" if bits0 & 1 else "<"
    if cls == 0:
        signed = bool(bits0 & 0x08)
        return np.dtype("%s%s%d" % (order, "i" if signed else "u", size))
    if cls == 1:
        if size not in (2, 4, 8):
            raise NotImplementedError("floating-point type of %d bytes" % size)
        return np.dtype("%sf%d" % (order, size))
    names = {2: "time", 3: "string", 4: "bitfield", 5: "opaque", 6: "compound", 7: "reference", 8: "enum",
             9: "variable-length", 10: "array"}
    raise NotImplementedError("HDF5 datatype class %d (%s, message version %d) is not supported"
                              % (cls, names.get(cls, "?"), ver))


def _parse_layout(body):
    ver = body[0]
    if ver == 3:
        cls = body[1]
        if cls == 0:
            n = int.from_bytes(body[2:4], "little")
            return {"class": "compact", "data": bytes(body[4:4 + n])}
        if cls == 1:
            return {"class": "contiguous", "address": int.from_bytes(body[2:10], "little"),
                    "size": int.from_bytes(body[10:18], "little")}
        if cls == 2:
            rank1 = body[2]
            addr = int.from_bytes(body[3:11], "little")
            dims = [int.from_bytes(body[11 + 4 * i:15 + 4 * i], "little") for i in range(rank1)]
            return {"class": "chunked", "address": addr, "chunk": dims}
        raise NotImplementedError("data layout class %d" % cls)
    if ver in (1, 2):
        rank, cls = body[1], body[2]
        p = 8
        addr = UNDEF
        if cls != 0:
            addr = int.from_bytes(body[p:p + 8], "little")
            p += 8
        dims = [int.from_bytes(body[p + 4 * i:p + 4 * i + 4], "little") for i in range(rank)]
        p += 4 * rank
        if cls == 1:
            return {"class": "contiguous", "address": addr, "size": None}
        if cls == 2:
            return {"class": "chunked", "address": addr, "chunk": dims}
        n = int.from_bytes(body[p:p + 4], "little")
        return {"class": "compact", "data": bytes(body[p + 4:p + 4 + n])}
    raise NotImplementedError("data layout message version %d (libver='latest' files) needs h5py" % ver)


def _parse_filters(body):
    ver, nf = body[0], body[1]
    p = 8 if ver == 1 else 2
    out = []
    for _ in range(nf):
        fid = int.from_bytes(body[p:p + 2], "little")
        if ver == 1 or fid >= 256:
            nlen = int.from_bytes(body[p + 2:p + 4], "little")
            p += 4
        else:
            nlen = 0
            p += 2
        p += 2                                             # flags
        ncv = int.from_bytes(body[p:p + 2], "little")
        p += 2
        name = body[p:p + nlen].split(b"\0")[0].decode("ascii", "replace")
        p += nlen if ver != 1 else (nlen + 7) // 8 * 8
        cvals = [int.from_bytes(body[p + 4 * i:p + 4 * i + 4], "little") for i in range(ncv)]
        p += 4 * ncv
        if ver == 1 and ncv % 2:
            p += 4
        out.append((fid, name, cvals))
    return out


# =============================================================================================
# writer: nested dict {name: ndarray | dict} -> classic-layout file with contiguous datasets
# =============================================================================================
_LEAF_K, _INTERNAL_K, _CHUNK_K = 4, 16, 32


class Chunked(object):
    """``Chunked(array, chunks, deflate=None, shuffle=False)`` as a value of the tree handed to
    ``write``: the dataset is stored in chunks (what h5py does for ``chunks=True`` -- the reference's
    Ritz-vector files, IOfiles.py:230), optionally shuffled and deflated at level ``deflate``."""

    def __init__(self, array, chunks, deflate=None, shuffle=False):
        self.array = np.asarray(array)
        self.chunks = tuple(int(c) for c in chunks)
        if len(self.chunks) != self.array.ndim or self.array.ndim == 0:
            raise ValueError("one chunk extent per dimension of a non-scalar array")
        self.deflate, self.shuffle = deflate, shuffle


def _pad8(n):
    return (n + 7) // 8 * 8


def _msg(mtype, body, flags=0):
    body = body + b"\0" * (_pad8(len(body)) - len(body))
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _datatype_msg(dt):
    dt = np.dtype(dt)
    be = 1 if dt.byteorder == ">" or (dt.byteorder == "=" and not np.little_endian) else 0
    if dt.kind in "iu":
        bits = be | (0x08 if dt.kind == "i" else 0)
        return struct.pack("<B3BI", 0x10, bits, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        if dt.itemsize == 8:
            sign, eloc, esz, msz, bias = 63, 52, 11, 52, 1023
        else:
            sign, eloc, esz, msz, bias = 31, 23, 8, 23, 127
        return (struct.pack("<B3BI", 0x11, 0x20 | be, sign, 0, dt.itemsize)
                + struct.pack("<HHBBBBI", 0, 8 * dt.itemsize, eloc, esz, 0, msz, bias))
    raise NotImplementedError("hdf5lite.write: dtype %s" % dt)


class _Writer(object):
    def __init__(self):
        self.buf = bytearray(96)                           # superblock (56) + root symbol table entry (40)

    def alloc(self, n):
        off = len(self.buf)
        self.buf += b"\0" * _pad8(n)
        return off

    def put(self, off, data):
        self.buf[off:off + len(data)] = data

    def chunked(self, ch):
        arr = ch.array
        rank, es = arr.ndim, arr.dtype.itemsize
        filters = b""
        nfilt = 0
        if ch.shuffle:
            filters += struct.pack("<HHHH", 2, 8, 1, 1) + b"shuffle\0" + struct.pack("<II", es, 0)
            nfilt += 1
        if ch.deflate is not None:
            filters += struct.pack("<HHHH", 1, 8, 1, 1) + b"deflate\0" + struct.pack("<II", int(ch.deflate), 0)
            nfilt += 1
        grid = [range(0, s, c) for s, c in zip(arr.shape, ch.chunks)]
        keys = []
        for offs in np.ndindex(*[len(g) for g in grid]):
            o = [g[i] for g, i in zip(grid, offs)]
            block = np.zeros(ch.chunks, dtype=arr.dtype)
            sl = tuple(slice(a, min(a + c, s)) for a, c, s in zip(o, ch.chunks, arr.shape))
            part = arr[sl]
            block[tuple(slice(0, n) for n in part.shape)] = part
            raw = block.tobytes()
            if ch.shuffle:
                raw = np.frombuffer(raw, dtype=np.uint8).reshape(-1, es).T.tobytes()
            if ch.deflate is not None:
                raw = zlib.compress(raw, int(ch.deflate))
            addr = self.alloc(len(raw))
            self.put(addr, raw)
            keys.append((len(raw), o, addr))
        if len(keys) > 2 * _CHUNK_K:
            raise NotImplementedError("hdf5lite.write: more than %d chunks per dataset" % (2 * _CHUNK_K))
        ksize = 8 + 8 * (rank + 1)
        bt = self.alloc(24 + (2 * _CHUNK_K + 1) * ksize + 2 * _CHUNK_K * 8)
        body = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(keys), UNDEF, UNDEF)
        for size, o, addr in keys:
            body += struct.pack("<II", size, 0) + b"".join(struct.pack("<Q", v) for v in o) + struct.pack("<QQ", 0, addr)
        body += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", v) for v in arr.shape) + struct.pack("<Q", 0)
        self.put(bt, body)
        space = struct.pack("<BBB5x", 1, rank, 1)
        for _ in range(2):
            space += b"".join(struct.pack("<Q", v) for v in arr.shape)
        layout = struct.pack("<BBB", 3, 2, rank + 1) + struct.pack("<Q", bt)
        layout += b"".join(struct.pack("<I", c) for c in ch.chunks) + struct.pack("<I", es)
        msgs = (_msg(0x0001, space) + _msg(0x0003, _datatype_msg(arr.dtype), flags=1)
                + _msg(0x0005, struct.pack("<BBBB4x", 2, 1, 0, 1), flags=1))
        nmsg = 4
        if nfilt:
            msgs += _msg(0x000B, struct.pack("<BB6x", 1, nfilt) + filters, flags=1)
            nmsg += 1
        msgs += _msg(0x0008, layout)
        return self.header(msgs, nmsg)

    def dataset(self, arr):
        if isinstance(arr, Chunked):
            return self.chunked(arr)
        arr = np.asarray(arr)
        if arr.dtype == np.bool_:
            arr = arr.astype(np.uint8)
        rank = arr.ndim
        space = struct.pack("<BBB5x", 1, rank, 1 if rank else 0)
        for _ in range(2 if rank else 0):
            space += b"".join(struct.pack("<Q", s) for s in arr.shape)
        raw = np.ascontiguousarray(arr).tobytes()
        data_addr = self.alloc(len(raw)) if raw else UNDEF
        if raw:
            self.put(data_addr, raw)
        msgs = (_msg(0x0001, space) + _msg(0x0003, _datatype_msg(arr.dtype), flags=1)
                + _msg(0x0005, struct.pack("<BBBB4x", 2, 2, 2, 1), flags=1)
                + _msg(0x0008, struct.pack("<BBQQ", 3, 1, data_addr, len(raw))))
        return self.header(msgs, 4)

    def header(self, msgs, nmsg):
        off = self.alloc(16 + len(msgs))
        self.put(off, struct.pack("<BBHII4x", 1, 0, nmsg, 1, len(msgs)) + msgs)
        return off

    def group(self, tree):
        """Write the members, then heap + symbol nodes + B-tree; returns (header, btree, heap)."""
        names = sorted(tree)                               # symbol tables are ordered by name
        entries = []
        for name in names:
            val = tree[name]
            if isinstance(val, dict):
                hdr, bt, heap = self.group(val)
                entries.append((name, hdr, 1, struct.pack("<QQ", bt, heap)))
            else:
                entries.append((name, self.dataset(val), 0, b"\0" * 16))
        # local heap: offset 0 is the empty string, names follow, each padded to 8 bytes
        heap_data = bytearray(8)
        offs = []
        for name, *_ in entries:
            offs.append(len(heap_data))
            nb = name.encode("utf-8") + b"\0"
            heap_data += nb + b"\0" * (_pad8(len(nb)) - len(nb))
        free_off = len(heap_data)
        heap_data += struct.pack("<QQ", 1, 16)             # one free block closing the segment
        heap = self.alloc(32)
        data = self.alloc(len(heap_data))
        self.put(data, heap_data)
        self.put(heap, b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, data))
        # symbol table nodes of up to 2*LEAF_K entries each, under one level-0 B-tree node
        per = 2 * _LEAF_K
        snods = []
        for i in range(0, max(len(entries), 1), per):
            part = list(zip(offs[i:i + per], entries[i:i + per]))
            node = self.alloc(8 + 40 * per)
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
            for off, (_name, hdr, ctype, scratch) in part:
                body += struct.pack("<QQI4x", off, hdr, ctype) + scratch
            self.put(node, body)
            snods.append((node, part[-1][0] if part else 0))
        if len(snods) > 2 * _INTERNAL_K:
            raise NotImplementedError("hdf5lite.write: more than %d objects in one group" % (per * 2 * _INTERNAL_K))
        bt = self.alloc(24 + (2 * _INTERNAL_K + 1) * 8 + 2 * _INTERNAL_K * 8)
        body = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF) + struct.pack("<Q", 0)
        for node, last_off in snods:
            body += struct.pack("<QQ", node, last_off)
        self.put(bt, body)
        hdr = self.header(_msg(0x0011, struct.pack("<QQ", bt, heap)), 1)
        return hdr, bt, heap


def write(filename, tree):
    """Create ``filename`` from a nested dict: ``{"obspix": array, "bolo_pair_0": {"pixel": array, ...}}``.
    Arrays keep their dtype and byte order (``np.dtype('>i4')`` gives the reference's STD_I32BE);
    every dataset is stored contiguously."""
    w = _Writer()
    hdr, bt, heap = w.group(tree)
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, _LEAF_K, _INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, len(w.buf), UNDEF)
    sb += struct.pack("<QQI4x", 0, hdr, 1) + struct.pack("<QQ", bt, heap)
    w.put(0, sb)
    with open(filename, "wb") as fh:
        fh.write(bytes(w.buf))
