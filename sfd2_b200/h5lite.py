"""Minimal pure-Python HDF5 reader / writer for the reference's feature and match files.

The reference stores everything through h5py (extract_localization.py:235-272, hloc/match_features.py:84-121; read back
by hloc/triangulation.py:57-111 and it_loc): one group per image (or per pair) holding a few plain numeric arrays.  That
needs only the oldest, most widely readable corner of the format, which is what h5py itself writes by default
(libver='earliest'):

  superblock version 0, version-1 object headers, "old style" groups (symbol-table message -> v1 B-tree of symbol-table
  nodes + local heap), contiguous little-endian datasets of fixed-point / IEEE floating-point type.

This module writes exactly those structures (HDF5 File Format Specification, sections II.A, III.A, III.D, III.E, IV.A) and
reads them back - including files written by libhdf5 itself (tests read a MATLAB v7.3 sample that ships with scipy).
No compression, chunking, attributes, strings or links: not needed for these files and rejected when met.

    with File(path, "w") as f:  f.create_group("db/1.jpg").create_dataset("keypoints", data=arr)
    with File(path, "r") as f:  f["db/1.jpg"]["keypoints"].__array__()          # the h5py subset the reference uses
"""
import os
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b"\x89HDF\r\n\x1a\n"
LEAF_K, INTERNAL_K = 32, 64           # symbols per node = 2 * LEAF_K, children per B-tree node = 2 * INTERNAL_K
SUPERBLOCK_BYTES = 96
HEAP_FREE_NULL = 1


class H5Error(RuntimeError):
    pass


def _pad8(n):
    return (n + 7) & ~7


# ------------------------------------------------------------------------------------------ datatypes
def _encode_dtype(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.byteorder == ">":
        raise H5Error("big-endian arrays are not supported")
    if dt.kind == "f" and dt.itemsize in (2, 4, 8):
        size = dt.itemsize
        exp_size, mant = {2: (5, 10), 4: (8, 23), 8: (11, 52)}[size]
        bias = (1 << (exp_size - 1)) - 1
        # class 1 (floating point), version 1; bits 0-7: LE, no padding, mantissa normalisation 2 (msb implied);
        # bits 8-15: sign bit position
        head = struct.pack("<BBBBI", 0x11, 0x20, size * 8 - 1, 0, size)
        return head + struct.pack("<HHBBBBI", 0, size * 8, mant, exp_size, 0, mant, bias)
    if dt.kind in "iu" and dt.itemsize in (1, 2, 4, 8):
        head = struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize)
        return head + struct.pack("<HH", 0, dt.itemsize * 8)
    raise H5Error(f"dtype {dt} is not supported (numeric little-endian arrays only)")


def _decode_dtype(b: bytes) -> np.dtype:
    cls, ver = b[0] & 0x0F, b[0] >> 4
    size = struct.unpack_from("<I", b, 4)[0]
    if ver not in (1, 2, 3):
        raise H5Error(f"datatype version {ver}")
    big = b[1] & 1
    if cls == 0:
        signed = (b[1] >> 3) & 1
        return np.dtype(("i" if signed else "u") + str(size)).newbyteorder(">" if big else "<")
    if cls == 1:
        if size not in (2, 4, 8):
            raise H5Error(f"float of {size} bytes")
        return np.dtype("f" + str(size)).newbyteorder(">" if big else "<")
    raise H5Error(f"datatype class {cls} is not supported")


# ------------------------------------------------------------------------------------------ writer
class _WDataset:
    def __init__(self, shape, dtype, addr, nbytes):
        self.shape, self.dtype, self.addr, self.nbytes = tuple(shape), np.dtype(dtype), addr, nbytes


class _WGroup:
    def __init__(self, file, name):
        self._file, self.name = file, name
        self.children = {}           # name -> _WGroup | _WDataset | ("raw", header address) for untouched foreign objects

    # h5py surface used by the reference
    def create_group(self, name):
        return self._file._create_group(self, name)

    def create_dataset(self, name, data=None, **kw):
        if data is None:
            raise H5Error("create_dataset needs data=")
        return self._file._create_dataset(self, name, data)

    def __contains__(self, name):
        return self._file._lookup(self, name) is not None

    def __getitem__(self, name):
        node = self._file._lookup(self, name)
        if node is None:
            raise KeyError(name)
        return self._file._wrap(node)

    def keys(self):
        return list(self.children.keys())

    def items(self):
        return [(k, self._file._wrap(v)) for k, v in self.children.items()]


class _DatasetView:
    """What feature_file[name][k] returns: supports .__array__(), np.asarray, .shape, .dtype, [()] / [...]."""

    def __init__(self, file, ds):
        self._file, self._ds = file, ds
        self.shape, self.dtype = ds.shape, ds.dtype

    def __array__(self, dtype=None, copy=None):
        a = self._file._read_data(self._ds)
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, idx):
        return self.__array__()[idx]

    def __len__(self):
        return self.shape[0]


class File(_WGroup):
    """h5py.File look-alike for modes 'r', 'w', 'a' (subset).  Writing streams the raw array data to disk as datasets are
    created and emits all metadata (object headers, B-trees, heaps, superblock) on close()."""

    def __init__(self, path, mode="r"):
        if mode not in ("r", "w", "a"):
            raise H5Error(f"mode {mode!r}")
        self.path, self.mode = str(path), mode
        _WGroup.__init__(self, self, "/")
        self._fh = None
        self._dirty = False
        exists = os.path.exists(self.path)
        if mode == "r" or (mode == "a" and exists):
            self._fh = open(self.path, "rb" if mode == "r" else "r+b")
            self._load()
            if mode == "a":
                self._fh.seek(0, os.SEEK_END)
                self._end = _pad8(self._fh.tell())
        else:
            self._fh = open(self.path, "w+b")
            self._fh.write(b"\0" * SUPERBLOCK_BYTES)        # superblock goes here on close
            self._end = SUPERBLOCK_BYTES
            self._base = 0
            self._dirty = True

    # ---- tree helpers
    def _walk(self, parent, name, create):
        parts = [p for p in name.split("/") if p]
        g = parent if not name.startswith("/") else self
        for p in parts[:-1]:
            nxt = g.children.get(p)
            if nxt is None:
                if not create:
                    return None, None
                nxt = _WGroup(self, p)
                g.children[p] = nxt
                self._dirty = True
            if not isinstance(nxt, _WGroup):
                raise H5Error(f"'{p}' is not a group")
            g = nxt
        return g, (parts[-1] if parts else None)

    def _lookup(self, parent, name):
        g, leaf = self._walk(parent, name, False)
        if g is None:
            return None
        return g if leaf is None else g.children.get(leaf)

    def _wrap(self, node):
        return _DatasetView(self, node) if isinstance(node, _WDataset) else node

    def _create_group(self, parent, name):
        self._check_writable()
        g, leaf = self._walk(parent, name, True)
        if leaf in g.children:
            raise ValueError(f"Unable to create group (name already exists): {name}")
        g.children[leaf] = _WGroup(self, leaf)
        self._dirty = True
        return g.children[leaf]

    def _create_dataset(self, parent, name, data):
        self._check_writable()
        g, leaf = self._walk(parent, name, True)
        if leaf in g.children:
            raise ValueError(f"Unable to create dataset (name already exists): {name}")
        a = np.ascontiguousarray(data)
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("<"))
        _encode_dtype(a.dtype)                       # raises on unsupported types before anything is written
        addr = UNDEF
        if a.nbytes:
            addr = self._end
            self._fh.seek(addr)
            self._fh.write(a.tobytes())
            self._end = _pad8(addr + a.nbytes)
        ds = _WDataset(a.shape, a.dtype, addr, a.nbytes)
        g.children[leaf] = ds
        self._dirty = True
        return _DatasetView(self, ds)

    def _check_writable(self):
        if self.mode == "r":
            raise H5Error("file is open read-only")

    def _read_data(self, ds):
        if ds.nbytes == 0:
            return np.zeros(ds.shape, ds.dtype)
        self._fh.seek(ds.addr)
        buf = self._fh.read(ds.nbytes)
        if len(buf) != ds.nbytes:
            raise H5Error("truncated dataset")
        return np.frombuffer(buf, dtype=ds.dtype).reshape(ds.shape).copy()

    # ---- writing the metadata
    def _alloc(self, blob: bytes):
        addr = self._end
        self._fh.seek(addr)
        self._fh.write(blob)
        self._end = _pad8(addr + len(blob))
        if self._end > addr + len(blob):
            self._fh.write(b"\0" * (self._end - addr - len(blob)))
        return addr

    @staticmethod
    def _message(mtype, data, flags=0):
        data = data + b"\0" * (_pad8(len(data)) - len(data))
        return struct.pack("<HHBBBB", mtype, len(data), flags, 0, 0, 0) + data

    def _object_header(self, messages):
        body = b"".join(messages)
        return struct.pack("<BBHII", 1, 0, len(messages), 1, len(body)) + b"\0" * 4 + body

    def _write_dataset(self, ds):
        rank = len(ds.shape)
        space = struct.pack("<BBB5x", 1, rank, 0) + b"".join(struct.pack("<Q", d) for d in ds.shape)
        layout = struct.pack("<BBQQ", 3, 1, (ds.addr - self._base) if ds.nbytes else UNDEF, ds.nbytes)
        fill = struct.pack("<BBBB", 2, 2, 2, 0)           # version 2, late allocation, write if set, undefined
        msgs = [self._message(0x0001, space), self._message(0x0003, _encode_dtype(ds.dtype), 1),
                self._message(0x0005, fill, 1), self._message(0x0008, layout)]
        return self._alloc(self._object_header(msgs))

    def _write_group(self, g):
        """-> (object header address, B-tree address, heap address), all absolute."""
        entries = []
        for name in sorted(g.children, key=lambda s: s.encode()):      # symbol-table order = strcmp order
            node = g.children[name]
            if isinstance(node, _WGroup):
                oh, bt, hp = self._write_group(node)
                entries.append((name, oh, 1, bt, hp))
            elif isinstance(node, _WDataset):
                entries.append((name, self._write_dataset(node), 0, 0, 0))
            else:
                raise H5Error("unsupported object")
        # local heap: offset 0 = "" (the B-tree's first key), then the names, then one free block
        data = bytearray(8)
        offs = []
        for name, *_ in entries:
            offs.append(len(data))
            nb = name.encode() + b"\0"
            data += nb + b"\0" * (_pad8(len(nb)) - len(nb))
        free_off = len(data)
        data += struct.pack("<QQ", HEAP_FREE_NULL, 16)             # free block: no successor, 16 bytes long
        data_addr = self._alloc(bytes(data))
        heap_addr = self._alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(data), free_off, data_addr - self._base))
        # symbol-table nodes of <= 2 * LEAF_K entries each
        per = 2 * LEAF_K
        leaves = []                                              # (address, heap offset of the last name)
        for i in range(0, len(entries), per):
            chunk = list(zip(entries[i:i + per], offs[i:i + per]))
            blob = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk)))
            for (name, oh, cache, bt, hp), off in chunk:
                scratch = struct.pack("<QQ", bt - self._base, hp - self._base) if cache == 1 else b"\0" * 16
                blob += struct.pack("<QQII", off, oh - self._base, cache, 0) + scratch
            blob += b"\0" * (8 + per * 40 - len(blob))
            leaves.append((self._alloc(bytes(blob)), chunk[-1][1]))
        # v1 B-tree (node type 0) over the leaves; more than 2 * INTERNAL_K children -> another level
        level = 0
        nodes = leaves
        first_key = 0
        while True:
            per_node = 2 * INTERNAL_K
            parents = []
            groups_ = [nodes[i:i + per_node] for i in range(0, len(nodes), per_node)] or [[]]     # empty group: a node with 0 entries
            addrs = []
            # reserve addresses first: siblings point at each other
            node_size = 24 + (2 * per_node + 1) * 8
            base_addr = self._end
            for k in range(len(groups_)):
                addrs.append(base_addr + k * _pad8(node_size))
            for k, grp in enumerate(groups_):
                left = addrs[k - 1] - self._base if k > 0 else UNDEF
                right = addrs[k + 1] - self._base if k + 1 < len(groups_) else UNDEF
                blob = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, level, len(grp), left, right))
                key = first_key if k == 0 else groups_[k - 1][-1][1]
                blob += struct.pack("<Q", key)
                for child_addr, last_key in grp:
                    blob += struct.pack("<QQ", child_addr - self._base, last_key)
                blob += b"\0" * (node_size - len(blob))
                got = self._alloc(bytes(blob))
                assert got == addrs[k]
                parents.append((got, grp[-1][1] if grp else 0))
            if len(parents) == 1:
                btree_addr = parents[0][0]
                break
            nodes, level = parents, level + 1
        stab = struct.pack("<QQ", btree_addr - self._base, heap_addr - self._base)
        oh = self._alloc(self._object_header([self._message(0x0011, stab, 0)]))
        return oh, btree_addr, heap_addr

    def flush(self):
        if self.mode == "r" or not self._dirty:
            return
        oh, bt, hp = self._write_group(self)
        eof = self._end
        sb = SIG + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", self._base, UNDEF, eof - self._base, UNDEF)
        sb += struct.pack("<QQII", 0, oh - self._base, 1, 0) + struct.pack("<QQ", bt - self._base, hp - self._base)
        assert len(sb) == SUPERBLOCK_BYTES
        self._fh.seek(self._sb_offset if hasattr(self, "_sb_offset") else 0)
        self._fh.write(sb)
        self._fh.flush()
        self._dirty = False

    def close(self):
        if self._fh is not None:
            self.flush()
            self._fh.close()
            self._fh = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- reading
    def _load(self):
        fh = self._fh
        off = 0
        while True:                                  # the superblock may sit behind a user block: 0, 512, 1024, ...
            fh.seek(off)
            if fh.read(8) == SIG:
                break
            off = 512 if off == 0 else off * 2
            if off > (1 << 24):
                raise H5Error("not an HDF5 file")
        self._sb_offset = off
        fh.seek(off + 8)
        ver, _, _, _, _, so, sl, _ = struct.unpack("<8B", fh.read(8))
        if ver not in (0, 1) or so != 8 or sl != 8:
            raise H5Error(f"superblock version {ver} / offset size {so}: only what h5py's default (libver='earliest') writes is read")
        self._file_leaf_k, self._file_int_k, _ = struct.unpack("<HHI", fh.read(8))
        if ver == 1:
            fh.read(4)
        base, _, eof, _ = struct.unpack("<QQQQ", fh.read(32))
        self._base = base                            # every address in the file is relative to it
        _, root_oh, cache, _ = struct.unpack("<QQII", fh.read(24))
        fh.read(16)
        if (self._file_leaf_k, self._file_int_k) != (LEAF_K, INTERNAL_K) and self.mode == "a":
            raise H5Error("appending to a file with other B-tree ranks than this writer's is not supported: rewrite it")
        self._read_group_into(self, self._base + root_oh)

    def _read_at(self, addr, n):
        self._fh.seek(addr)
        b = self._fh.read(n)
        if len(b) != n:
            raise H5Error("truncated file")
        return b

    def _messages(self, addr):
        """All messages of a version-1 object header (following continuation blocks)."""
        ver, _, nmsgs, _, size = struct.unpack("<BBHII", self._read_at(addr, 12))
        if ver != 1:
            raise H5Error("only version-1 object headers (libver='earliest') are read")
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsgs:
            p, n = blocks.pop(0)
            buf = self._read_at(p, n)
            q = 0
            while q + 8 <= n and len(out) < nmsgs:
                t, s, fl = struct.unpack_from("<HHB", buf, q)
                data = buf[q + 8:q + 8 + s]
                q += 8 + s
                out.append((t, data))
                if t == 0x0010:
                    caddr, clen = struct.unpack("<QQ", data[:16])
                    blocks.append((self._base + caddr, clen))
        return out

    def _read_group_into(self, g, oh_addr):
        stab = None
        for t, data in self._messages(oh_addr):
            if t == 0x0011:
                stab = struct.unpack("<QQ", data[:16])
            elif t in (0x0002, 0x0006):
                raise H5Error("new-style (link message) groups are not supported; write the file with libver='earliest'")
        if stab is None:
            raise H5Error("group without a symbol table")
        bt, hp = self._base + stab[0], self._base + stab[1]
        hb = self._read_at(hp, 32)
        if hb[:4] != b"HEAP":
            raise H5Error("bad local heap")
        dsize, _, daddr = struct.unpack("<QQQ", hb[8:32])
        heap = self._read_at(self._base + daddr, dsize)
        for name_off, obj, cache, scratch in self._btree_symbols(bt):
            end = heap.index(b"\0", name_off)
            name = heap[name_off:end].decode()
            self._read_object_into(g, name, self._base + obj)

    def _btree_symbols(self, addr):
        b = self._read_at(addr, 24)
        if b[:4] != b"TREE":
            raise H5Error("bad B-tree node")
        ntype, level, used = struct.unpack("<BBH", b[4:8])
        if ntype != 0:
            raise H5Error("not a group B-tree")
        body = self._read_at(addr + 24, (2 * used + 1) * 8)
        for i in range(used):
            child = struct.unpack_from("<Q", body, 8 + 16 * i)[0] + self._base
            if level > 0:
                yield from self._btree_symbols(child)
            else:
                h = self._read_at(child, 8)
                if h[:4] != b"SNOD":
                    raise H5Error("bad symbol table node")
                n = struct.unpack("<H", h[6:8])[0]
                ents = self._read_at(child + 8, 40 * n)
                for k in range(n):
                    name_off, obj, cache, _ = struct.unpack_from("<QQII", ents, 40 * k)
                    yield name_off, obj, cache, ents[40 * k + 24:40 * k + 40]

    def _read_object_into(self, g, name, oh_addr):
        msgs = self._messages(oh_addr)
        types = {t for t, _ in msgs}
        if 0x0011 in types:
            child = _WGroup(self, name)
            g.children[name] = child
            self._read_group_into(child, oh_addr)
            return
        shape = dtype = layout = None
        for t, data in msgs:
            if t == 0x0001:
                ver, rank, flags = data[0], data[1], data[2]
                q = 8 if ver == 1 else 4
                shape = struct.unpack_from("<" + "Q" * rank, data, q) if rank else ()
            elif t == 0x0003:
                dtype = _decode_dtype(data)
            elif t == 0x0008:
                ver = data[0]
                if ver == 3:
                    cls = data[1]
                    if cls == 1:
                        layout = struct.unpack_from("<QQ", data, 2)
                    elif cls == 0:
                        n = struct.unpack_from("<H", data, 2)[0]
                        layout = ("compact", data[4:4 + n])
                    else:
                        raise H5Error(f"dataset '{name}': chunked storage is not supported")
                elif ver in (1, 2):
                    rank, cls = data[1], data[2]
                    if cls != 1:
                        raise H5Error(f"dataset '{name}': only contiguous storage is supported")
                    addr = struct.unpack_from("<Q", data, 8)[0]
                    dims = struct.unpack_from("<" + "I" * rank, data, 16)
                    layout = (addr, int(np.prod(dims)))
                else:
                    raise H5Error(f"layout message version {ver}")
            elif t in (0x000B,):
                raise H5Error(f"dataset '{name}': filters (compression) are not supported")
        if shape is None or dtype is None or layout is None:
            raise H5Error(f"object '{name}' is neither a group nor a plain dataset")
        nbytes = int(np.prod(shape)) * dtype.itemsize if len(shape) else dtype.itemsize
        if layout[0] == "compact":
            addr = self._alloc_compact(layout[1])
        else:
            addr = self._base + layout[0] if layout[0] != UNDEF else UNDEF
        g.children[name] = _WDataset(shape, dtype, addr, nbytes if addr != UNDEF else 0)

    def _alloc_compact(self, raw):
        raise H5Error("compact datasets are not supported")
