"""Minimal pure-Python HDF5 reader/writer for Keras weight files.

The reference stores its networks with Keras 2.3.1 / h5py (`model.save_weights`, `model.save`:
lattice_vae.py:149-151,339-341, unet.py:261-264,378-379).  h5py / libhdf5 are not available in this image, so
this module implements the subset of the HDF5 file format those files use — the "classic" layout libhdf5
writes by default (`libver='earliest'`):

    superblock v0/v1 · object header v1 (+ continuation blocks) · groups as symbol tables
    (B-tree v1 + SNOD nodes + local heap) · contiguous / compact dataset layout (layout message v1-v3) ·
    fixed-point / IEEE float / fixed-length string datatypes · attribute messages v1-v3 · dataspace v1/v2

Chunked or filtered datasets, variable-length strings, new-style (v2 object header / fractal heap) groups and
superblock v2+ are outside that subset: reading them raises H5FormatError with a message that says which feature
was met (attributes of unsupported type are skipped).  `tools/keras_h5_to_npz.py` is the offline converter for such
files on a machine that has h5py.

API:  read(path) -> Group   (Group.attrs: dict, Group.keys(), Group[name] -> Group | numpy array, '/'-paths work)
      write(path, tree)      tree = {"attrs": {...}, "items": {name: subtree | numpy array}} (see Node)
"""
from __future__ import annotations

import struct

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(ValueError):
    pass


def is_hdf5(path) -> bool:
    """True when an HDF5 signature sits at offset 0, 512, 1024, ... (a user block may precede the superblock)."""
    try:
        with open(path, "rb") as f:
            off = 0
            for _ in range(12):
                f.seek(off)
                if f.read(8) == SIGNATURE:
                    return True
                off = 512 if off == 0 else off * 2
    except OSError:
        return False
    return False


# ======================================================================================================
# reader
# ======================================================================================================
class Group:
    def __init__(self, f, attrs, links):
        self._f, self.attrs, self._links = f, attrs, links  # links: name -> object header address

    def keys(self):
        return list(self._links)

    def __contains__(self, name):
        return name in self._links

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group) or part not in node._links:
                raise KeyError(path)
            node = node._f._object(node._links[part])
        return node


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        off = 0
        while True:
            if buf[off:off + 8] == SIGNATURE:
                break
            off = 512 if off == 0 else off * 2
            if off + 8 > len(buf):
                raise H5FormatError("not an HDF5 file (no signature)")
        ver = buf[off + 8]
        if ver > 1:
            raise H5FormatError(f"HDF5 superblock version {ver} (new-style file, libver='latest'): only the classic format "
                                "Keras/h5py write by default is supported — convert with tools/keras_h5_to_npz.py")
        self.O, self.L = buf[off + 13], buf[off + 14]
        if self.O != 8 or self.L != 8:
            raise H5FormatError(f"HDF5 offsets/lengths of {self.O}/{self.L} bytes are not supported (expected 8/8)")
        p = off + 24 + (4 if ver == 1 else 0)
        self.base = self._u(p, 8)
        p += 4 * 8  # base, free-space, end-of-file, driver-info addresses
        # root symbol table entry
        ohdr = self._u(p + 8, 8)
        self.root_addr = ohdr
        self._cache = {}

    # ---- primitives ----
    def _u(self, p, n):
        return int.from_bytes(self.b[p:p + n], "little")

    def _abs(self, addr):
        return self.base + addr

    # ---- object header v1 ----
    def _messages(self, addr):
        p = self._abs(addr)
        if self.b[p:p + 4] == b"OHDR":
            raise H5FormatError("version-2 object header (new-style HDF5 file) is not supported — convert with "
                                "tools/keras_h5_to_npz.py")
        if self.b[p] != 1:
            raise H5FormatError(f"object header version {self.b[p]} at {p:#x}")
        nmsg, size = self._u(p + 2, 2), self._u(p + 8, 4)
        blocks = [(p + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            q, left = blocks.pop(0)
            end = q + left
            while q + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = self._u(q, 2), self._u(q + 2, 2), self.b[q + 4]
                body = q + 8
                if mtype == 0x10:  # continuation
                    blocks.append((self._abs(self._u(body, 8)), self._u(body + 8, 8)))
                out.append((mtype, body, msize, flags))
                q = body + msize
        return out

    def _object(self, addr):
        if addr in self._cache:
            return self._cache[addr]
        msgs = self._messages(addr)
        attrs, symtab, space, dtype, layout = {}, None, None, None, None
        for mtype, p, size, flags in msgs:
            if mtype == 0x11:
                symtab = (self._u(p, 8), self._u(p + 8, 8))
            elif mtype == 0x01:
                space = self._dataspace(p)
            elif mtype == 0x03:
                dtype = self._datatype(p)
            elif mtype == 0x08:
                layout = p
            elif mtype == 0x0C:
                name, val = self._attribute(p)
                if val is not None:
                    attrs[name] = val
            elif mtype in (0x02, 0x06):
                raise H5FormatError("new-style group (link info / link messages) is not supported — convert with "
                                    "tools/keras_h5_to_npz.py")
            elif mtype == 0x0B:
                raise H5FormatError("filtered (compressed) dataset is not supported")
        if symtab is not None:
            obj = Group(self, attrs, self._links(*symtab))
        elif layout is not None:
            obj = self._dataset(space, dtype, layout)
        else:
            obj = Group(self, attrs, {})
        self._cache[addr] = obj
        return obj

    # ---- groups ----
    def _heap_name(self, heap_addr, off):
        h = self._abs(heap_addr)
        if self.b[h:h + 4] != b"HEAP":
            raise H5FormatError("bad local heap signature")
        data = self._abs(self._u(h + 24, 8))
        e = self.b.index(b"\0", data + off)
        return self.b[data + off:e].decode("utf8")

    def _links(self, btree, heap):
        links = {}

        def walk(addr):
            p = self._abs(addr)
            if self.b[p:p + 4] == b"SNOD":
                n = self._u(p + 6, 2)
                q = p + 8
                for _ in range(n):
                    links[self._heap_name(heap, self._u(q, 8))] = self._u(q + 8, 8)
                    q += 40
                return
            if self.b[p:p + 4] != b"TREE":
                raise H5FormatError(f"expected TREE/SNOD at {p:#x}")
            n = self._u(p + 6, 2)
            q = p + 24 + 8  # skip key 0
            for _ in range(n):
                walk(self._u(q, 8))
                q += 16

        walk(btree)
        return links

    # ---- datasets / attributes ----
    def _dataspace(self, p):
        ver, rank, flags = self.b[p], self.b[p + 1], self.b[p + 2]
        if ver == 1:
            q = p + 8
        elif ver == 2:
            if self.b[p + 3] == 2:
                return None  # null dataspace
            q = p + 4
        else:
            raise H5FormatError(f"dataspace version {ver}")
        return tuple(self._u(q + 8 * i, 8) for i in range(rank))

    def _datatype(self, p):
        """-> (numpy dtype | None, size in bytes of the message)."""
        cls, ver = self.b[p] & 15, self.b[p] >> 4
        bits0 = self.b[p + 1]
        size = self._u(p + 4, 4)
        if cls == 0:  # fixed point
            signed = bool(bits0 & 8)
            dt = np.dtype(("<" if not bits0 & 1 else ">") + ("i" if signed else "u") + str(size))
            return dt, 8 + 4
        if cls == 1:  # IEEE float
            if size not in (2, 4, 8):
                return None, 8 + 12
            return np.dtype(("<" if not bits0 & 1 else ">") + "f" + str(size)), 8 + 12
        if cls == 3:  # fixed-length string
            return np.dtype("S" + str(size)), 8
        return None, 8  # variable length, compound, ...: not needed for weight files

    def _attribute(self, p):
        ver = self.b[p]
        nsz, tsz, ssz = self._u(p + 2, 2), self._u(p + 4, 2), self._u(p + 6, 2)
        q = p + 8 + (1 if ver == 3 else 0)
        pad = (lambda n: (n + 7) // 8 * 8) if ver == 1 else (lambda n: n)
        name = self.b[q:q + nsz].split(b"\0")[0].decode("utf8")
        q += pad(nsz)
        dt, _ = self._datatype(q)
        q += pad(tsz)
        shape = self._dataspace(q)
        q += pad(ssz)
        if dt is None or shape is None:
            return name, None
        n = int(np.prod(shape)) if shape else 1
        arr = np.frombuffer(self.b, dtype=dt, count=n, offset=q).reshape(shape)
        return name, (arr.copy() if shape else arr.reshape(()).item())

    def _dataset(self, shape, dtype, p):
        dt = dtype[0] if dtype else None
        if dt is None or shape is None:
            raise H5FormatError("dataset of unsupported datatype / dataspace")
        n = int(np.prod(shape)) if shape else 1
        ver = self.b[p]
        if ver == 3:
            cls = self.b[p + 1]
            if cls == 1:
                addr = self._u(p + 2, 8)
            elif cls == 0:
                return np.frombuffer(self.b, dtype=dt, count=n, offset=p + 4).reshape(shape).copy()
            else:
                raise H5FormatError("chunked dataset layout is not supported (Keras writes contiguous datasets) — convert "
                                    "with tools/keras_h5_to_npz.py")
        elif ver in (1, 2):
            cls = self.b[p + 2]
            if cls == 1:
                addr = self._u(p + 8, 8)
            elif cls == 0:
                rank = self.b[p + 1]
                return np.frombuffer(self.b, dtype=dt, count=n, offset=p + 8 + 4 * rank + 4).reshape(shape).copy()
            else:
                raise H5FormatError("chunked dataset layout is not supported")
        else:
            raise H5FormatError(f"data layout message version {ver}")
        if addr == UNDEF:
            return np.zeros(shape, dtype=dt.newbyteorder("="))
        a = np.frombuffer(self.b, dtype=dt, count=n, offset=self._abs(addr)).reshape(shape)
        return a.astype(dt.newbyteorder("=")) if dt.kind != "S" else a.copy()


def read(path) -> Group:
    with open(path, "rb") as f:
        r = _Reader(f.read())
    root = r._object(r.root_addr)
    if not isinstance(root, Group):
        raise H5FormatError("root object is not a group")
    return root


# ======================================================================================================
# writer
# ======================================================================================================
class Node:
    """A group to write: attrs {name: str | bytes | numpy array (float / int / 'S')}, items {name: Node | numpy array}."""

    def __init__(self, attrs=None, items=None):
        self.attrs = dict(attrs or {})
        self.items = dict(items or {})

    def group(self, name) -> "Node":
        return self.items.setdefault(name, Node())


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _dtype_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, dt.itemsize)  # class 3 v1, null-padded ASCII (what h5py writes for 'S')
    if dt.kind == "f" and dt.itemsize in (4, 8):
        ebits, mbits = (8, 23) if dt.itemsize == 4 else (11, 52)
        head = struct.pack("<BBBBI", 0x11, 0x20, 8 * dt.itemsize - 1, 0, dt.itemsize)
        return head + struct.pack("<HHBBBBI", 0, 8 * dt.itemsize, mbits, ebits, 0, mbits, (1 << (ebits - 1)) - 1)
    if dt.kind in "iu":
        head = struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize)
        return head + struct.pack("<HH", 0, 8 * dt.itemsize)
    raise H5FormatError(f"cannot write dtype {dt}")


def _space_msg(shape) -> bytes:
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(s)) for s in shape)


def _msg(mtype, body, flags=0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _attr_msg(name, value) -> bytes:
    if isinstance(value, str):
        value = value.encode("utf8")
    if isinstance(value, bytes):
        arr = np.array(value, dtype="S%d" % max(1, len(value)))
    else:
        arr = np.asarray(value)
        if arr.dtype.kind == "U":
            arr = np.char.encode(arr, "utf8")
        if arr.dtype.kind == "S" and arr.dtype.itemsize == 0:
            arr = arr.astype("S1")
    if arr.dtype.byteorder == ">":
        arr = arr.astype(arr.dtype.newbyteorder("<"))
    nm = name.encode("utf8") + b"\0"
    dtm, spm = _dtype_msg(arr.dtype), _space_msg(arr.shape)
    body = struct.pack("<BxHHH", 1, len(nm), len(dtm), len(spm)) + _pad8(nm) + _pad8(dtm) + _pad8(spm) + arr.tobytes()
    if len(body) > 64000:
        raise H5FormatError(f"attribute {name} too large for one object-header message")
    return _msg(0x0C, body)


def _ohdr(msgs) -> bytes:
    body = b"".join(msgs)
    return struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body


class _Writer:
    LEAF_K, NODE_K = 4, 16

    def __init__(self):
        self.buf = bytearray()

    def alloc(self, data: bytes, align=8) -> int:
        self.buf += b"\0" * (-len(self.buf) % align)
        addr = len(self.buf)
        self.buf += data
        return addr

    def dataset(self, arr) -> int:
        arr = np.ascontiguousarray(arr)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        raw = arr.tobytes()
        daddr = self.alloc(raw) if raw else UNDEF
        msgs = [
            _msg(0x01, _space_msg(arr.shape)),
            _msg(0x03, _dtype_msg(arr.dtype), flags=1),
            _msg(0x05, struct.pack("<BBBBI", 1, 2, 2, 1, 0)),  # fill value: late allocation, default fill
            _msg(0x08, struct.pack("<BBQQ", 3, 1, daddr, len(raw))),
        ]
        return self.alloc(_ohdr(msgs))

    def group(self, node: Node):
        """-> (object header address, b-tree address, local heap address)."""
        children = {}
        for name, item in node.items.items():
            if isinstance(item, Node):
                children[name] = self.group(item)
            else:
                children[name] = (self.dataset(item), None, None)
        names = sorted(children, key=lambda s: s.encode("utf8"))
        if len(names) > 2 * self.LEAF_K * 2 * self.NODE_K:
            raise H5FormatError("too many links in one group for a single-level B-tree")
        # local heap: offset 0 = empty string, then the names (8-byte aligned, as libhdf5 lays them out)
        heap_data = bytearray(b"\0" * 8)
        offs = {}
        for n in names:
            offs[n] = len(heap_data)
            heap_data += _pad8(n.encode("utf8") + b"\0")
        free_off = len(heap_data)
        heap_data += struct.pack("<QQ", 1, 16)  # one free block at the end (next = 1: none)
        heap_data_addr = self.alloc(bytes(heap_data))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, heap_data_addr))
        # symbol-table nodes of up to 2*LEAF_K entries each, in name order
        per = 2 * self.LEAF_K
        snods, keys = [], [0]
        for i in range(0, len(names), per):
            chunk = names[i:i + per]
            body = b"SNOD" + struct.pack("<BxH", 1, len(chunk))
            for n in chunk:
                ohdr, bt, hp = children[n]
                if bt is None:
                    body += struct.pack("<QQII16x", offs[n], ohdr, 0, 0)
                else:
                    body += struct.pack("<QQIIQQ", offs[n], ohdr, 1, 0, bt, hp)
            body += b"\0" * (40 * (per - len(chunk)))
            snods.append(self.alloc(body))
            keys.append(offs[chunk[-1]])
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF)
        for i, a in enumerate(snods):
            tree += struct.pack("<QQ", keys[i], a)
        tree += struct.pack("<Q", keys[-1])
        tree += b"\0" * (16 * (2 * self.NODE_K - len(snods)))
        bt_addr = self.alloc(tree)
        msgs = [_msg(0x11, struct.pack("<QQ", bt_addr, heap_addr))] + [_attr_msg(k, v) for k, v in node.attrs.items()]
        return self.alloc(_ohdr(msgs)), bt_addr, heap_addr


def write(path, root: Node):
    w = _Writer()
    w.buf += b"\0" * 96  # superblock v0 (56 bytes) + root symbol-table entry (40 bytes)
    ohdr, bt, hp = w.group(root)
    eof = len(w.buf)
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, _Writer.LEAF_K, _Writer.NODE_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQIIQQ", 0, ohdr, 1, 0, bt, hp)
    assert len(sb) == 96
    w.buf[0:96] = sb
    with open(path, "wb") as f:
        f.write(bytes(w.buf))
