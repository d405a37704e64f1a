"""Minimal HDF5 writer/reader for the simulation's output files (no h5py in this image).

Files follow what h5py/libhdf5 1.10 produce for ``h5py.File(name, 'w').create_dataset(key, data=arr)``
(the only call the reference makes: simulation_utilities/sim_archive.py:13-23): superblock
version 0, root group as a symbol table (one B-tree node + local heap + one symbol-table node),
version-1 object headers of 256 bytes with dataspace / datatype / fill-value / contiguous-layout /
modification-time messages, raw data contiguous.  The allocation order mimics the library
(2 KiB metadata and small-data blocks), so that files with the shapes of the shipped tutorial
outputs are byte-identical to them apart from the time stamps -- which is how the writer is tested
(tests/test_h5lite.py).  Readers that matter: h5py (used by the reference's SimInfo,
analysis/extract_sim_info.py:29-54) and read_h5() below.

Supported dtypes: float64, float32, int64, int32.  Up to 8 datasets per file, all in the root group.
"""
import struct
import time

import numpy as np

__all__ = ["write_h5", "read_h5", "File"]

UNDEF = 0xFFFFFFFFFFFFFFFF
BLOCK = 2048
GROUP_LEAF_K, GROUP_INTERNAL_K = 4, 16
HEAP_DATA_SIZE = 88
OHDR_DATA = 256                      # bytes of message data in every dataset object header


def _pad8(n):
    return (n + 7) // 8 * 8


def _dtype_message(dt):
    dt = np.dtype(dt)
    if dt == np.float64:
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 0x3F, 0x00, 8, 0, 64, 52, 11, 0, 52, 1023)
    if dt == np.float32:
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 0x1F, 0x00, 4, 0, 32, 23, 8, 0, 23, 127)
    if dt == np.int64:
        return struct.pack("<BBBBIHH", 0x10, 0x08, 0x00, 0x00, 8, 0, 64) + b"\x00" * 4
    if dt == np.int32:
        return struct.pack("<BBBBIHH", 0x10, 0x08, 0x00, 0x00, 4, 0, 32) + b"\x00" * 4
    raise TypeError(f"h5lite cannot store dtype {dt}")


def _message(mtype, body, flags=0):
    body = body + b"\x00" * (_pad8(len(body)) - len(body))
    return struct.pack("<HHBBBB", mtype, len(body), flags, 0, 0, 0) + body


def _object_header(arr, data_addr, mtime):
    rank = arr.ndim
    dims = struct.pack(f"<{rank}Q", *arr.shape) if rank else b""
    space = struct.pack("<BBBBI", 1, rank, 1, 0, 0) + dims + dims            # version 1, max dims present
    dtype_body = _dtype_message(arr.dtype)
    msgs = _message(0x0001, space)
    msgs += _message(0x0003, dtype_body, flags=1)
    msgs += _message(0x0005, struct.pack("<BBBBI", 2, 2, 2, 1, 0), flags=1)   # fill value (v2, late alloc, defined)
    layout = struct.pack("<BBQQ", 3, 1, data_addr if arr.nbytes else UNDEF, arr.nbytes)
    msgs += _message(0x0008, layout)
    msgs += _message(0x0012, struct.pack("<BBBBI", 1, 0, 0, 0, int(mtime) & 0xFFFFFFFF))
    nil = OHDR_DATA - len(msgs) - 8
    assert nil >= 0, "object header overflow"
    msgs += struct.pack("<HHBBBB", 0, nil, 0, 0, 0, 0) + b"\x00" * nil
    return struct.pack("<BBHII", 1, 0, 6, 1, OHDR_DATA) + b"\x00" * 4 + msgs


class _Alloc:
    """libhdf5-like file space allocation: metadata and small raw data come from 2 KiB aggregator
    blocks, large raw data from the end of file."""

    def __init__(self):
        self.eof = BLOCK
        self.meta = [0, BLOCK]          # [next free, end] of the current metadata block
        self.small = None

    def take_meta(self, size):
        if self.meta[0] + size > self.meta[1]:
            self.meta = [self.eof, self.eof + max(BLOCK, size)]
            self.eof = self.meta[1]
        addr = self.meta[0]
        self.meta[0] += size
        return addr

    def take_raw(self, size):
        if size >= BLOCK:
            addr = self.eof
            self.eof += size
            return addr
        if self.small is None or self.small[0] + size > self.small[1]:
            self.small = [self.eof, self.eof + BLOCK]
            self.eof = self.small[1]
        addr = self.small[0]
        self.small[0] += size
        return addr

    def final_eof(self):
        """At close the library returns the unused tail of whichever aggregator block ends the file."""
        eof = self.eof
        for blk in (self.meta, self.small):
            if blk is not None and blk[1] == eof and blk[0] < eof and blk[1] > BLOCK:
                eof = blk[0]
        return eof


def write_h5(fname, keys, values, mtime=None):
    """Equivalent of SimArchivist.save_h5 (reference sim_archive.py:13-23)."""
    mtime = time.time() if mtime is None else mtime
    arrays = []
    for v in values:
        a = np.asarray(v)
        if a.dtype == np.bool_ or (a.dtype.kind in "iu" and a.dtype != np.int32):
            a = a.astype(np.int64)                 # Python int lists become int64, like h5py does
        elif a.dtype.kind == "f" and a.dtype not in (np.float64, np.float32):
            a = a.astype(np.float64)
        arrays.append(np.ascontiguousarray(a))
    keys = [str(k) for k in keys]
    if len(keys) != len(arrays) or not 1 <= len(keys) <= 2 * GROUP_LEAF_K:
        raise ValueError("h5lite writes between 1 and 8 datasets per file")
    if len(set(keys)) != len(keys):
        raise ValueError("dataset names must be unique")

    al = _Alloc()
    root_ohdr = al.take_meta(96 + 40) + 96                                    # superblock, then root object header
    btree = al.take_meta(24 + (2 * GROUP_INTERNAL_K + 1) * 8 + 2 * GROUP_INTERNAL_K * 8)
    heap = al.take_meta(32)
    heap_data = al.take_meta(HEAP_DATA_SIZE)
    # local heap: offset 0 holds the empty string, names follow 8-byte aligned
    name_off, used = {}, 8
    for k in keys:
        name_off[k] = used
        used += _pad8(len(k.encode()) + 1)
    if used + 16 > HEAP_DATA_SIZE:
        raise ValueError("dataset names too long for the fixed-size local heap")
    ohdr_addr, data_addr, snod = {}, {}, None
    for k, a in zip(keys, arrays):
        ohdr_addr[k] = al.take_meta(16 + OHDR_DATA)
        if snod is None:                                                      # first link creates the symbol-table node
            snod = al.take_meta(8 + 2 * GROUP_LEAF_K * 40)
        data_addr[k] = al.take_raw(a.nbytes) if a.nbytes else UNDEF
    eof = al.final_eof()

    buf = bytearray(eof)
    # superblock (version 0) + root symbol-table entry
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, GROUP_LEAF_K, GROUP_INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root_ohdr, 1, 0) + struct.pack("<QQ", btree, heap)
    buf[0:len(sb)] = sb
    root = struct.pack("<BBHII", 1, 0, 1, 1, 24) + b"\x00" * 4 + _message(0x0011, struct.pack("<QQ", btree, heap))
    buf[root_ohdr:root_ohdr + len(root)] = root
    last = max(keys)
    node = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod, name_off[last])
    buf[btree:btree + len(node)] = node
    hp = b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, HEAP_DATA_SIZE, used, heap_data)
    buf[heap:heap + len(hp)] = hp
    hd = bytearray(HEAP_DATA_SIZE)
    for k in keys:
        kb = k.encode()
        hd[name_off[k]:name_off[k] + len(kb)] = kb
    hd[used:used + 16] = struct.pack("<QQ", 1, HEAP_DATA_SIZE - used)          # free block: no next, size
    buf[heap_data:heap_data + HEAP_DATA_SIZE] = hd
    sn = b"SNOD" + struct.pack("<BBH", 1, 0, len(keys))
    for k in sorted(keys):
        sn += struct.pack("<QQII", name_off[k], ohdr_addr[k], 0, 0) + b"\x00" * 16
    buf[snod:snod + len(sn)] = sn
    for k, a in zip(keys, arrays):
        oh = _object_header(a, data_addr[k], mtime)
        buf[ohdr_addr[k]:ohdr_addr[k] + len(oh)] = oh
        if a.nbytes:
            buf[data_addr[k]:data_addr[k] + a.nbytes] = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
    with open(fname, "wb") as fh:
        fh.write(bytes(buf))


# ----------------------------------------------------------------------------- reader
class _Reader:
    def __init__(self, raw):
        self.b = raw
        if raw[:8] != b"\x89HDF\r\n\x1a\n" or raw[8] != 0:
            raise ValueError("not an HDF5 file with a version-0 superblock")
        self.root_ohdr = struct.unpack_from("<Q", raw, 64)[0]

    def messages(self, addr):
        """Yield (type, body) of a version-1 object header, following continuation blocks."""
        ver, _, nmsg, _, size = struct.unpack_from("<BBHII", self.b, addr)
        if ver != 1:
            raise ValueError("only version-1 object headers are supported")
        blocks, seen = [(addr + 16, size)], 0
        while blocks and seen < nmsg:
            pos, left = blocks.pop(0)
            while left >= 8 and seen < nmsg:
                mtype, msize, _ = struct.unpack_from("<HHB", self.b, pos)
                body = self.b[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                left -= 8 + msize
                seen += 1
                if mtype == 0x0010:
                    blocks.append(struct.unpack_from("<QQ", body, 0))
                else:
                    yield mtype, body

    def group_entries(self, btree, heap):
        heap_data = struct.unpack_from("<Q", self.b, heap + 24)[0]

        def walk(node):
            if self.b[node:node + 4] == b"SNOD":
                n = struct.unpack_from("<H", self.b, node + 6)[0]
                for i in range(n):
                    off, oh = struct.unpack_from("<QQ", self.b, node + 8 + 40 * i)
                    end = self.b.index(b"\x00", heap_data + off)
                    yield self.b[heap_data + off:end].decode(), oh
                return
            assert self.b[node:node + 4] == b"TREE"
            used = struct.unpack_from("<H", self.b, node + 6)[0]
            for i in range(used):
                child = struct.unpack_from("<Q", self.b, node + 24 + 8 + 16 * i)[0]
                yield from walk(child)
        yield from walk(btree)

    def load(self, addr, prefix, out):
        space = dtype = layout = None
        for mtype, body in self.messages(addr):
            if mtype == 0x0011:
                bt, hp = struct.unpack_from("<QQ", body, 0)
                for name, oh in self.group_entries(bt, hp):
                    self.load(oh, f"{prefix}{name}/", out)
                return
            if mtype == 0x0001:
                rank = body[1]
                space = struct.unpack_from(f"<{rank}Q", body, 8) if rank else ()
            elif mtype == 0x0003:
                cls, size = body[0] & 0x0F, struct.unpack_from("<I", body, 4)[0]
                dtype = {(1, 8): "<f8", (1, 4): "<f4", (0, 8): "<i8", (0, 4): "<i4"}.get((cls, size))
            elif mtype == 0x0008 and body[0] == 3 and body[1] == 1:
                layout = struct.unpack_from("<QQ", body, 2)
        if space is not None and dtype and layout:
            a, nbytes = layout
            n = int(np.prod(space)) if space else 1
            arr = np.frombuffer(self.b, dtype, n, a).reshape(space) if nbytes else np.zeros(space, dtype)
            out[prefix.rstrip("/")] = arr.copy()


def read_h5(fname):
    """{dataset path: array} for every contiguous float/int dataset in the file (groups are walked)."""
    with open(fname, "rb") as fh:
        raw = fh.read()
    r, out = _Reader(raw), {}
    r.load(r.root_ohdr, "", out)
    return out


class File:
    """Tiny h5py.File look-alike for the two access patterns the reference uses
    (``with File(f,'w') as hf: hf.create_dataset(k, data=v)`` and ``with File(f,'r') as hf: hf[k][:]``)."""

    def __init__(self, fname, mode="r"):
        self.fname, self.mode, self._k, self._v = fname, mode, [], []
        self._d = read_h5(fname) if mode == "r" else None

    def create_dataset(self, key, data=None):
        self._k.append(key)
        self._v.append(data)

    def __getitem__(self, key):
        return self._d[key]

    def keys(self):
        return list(self._d)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if self.mode != "r" and self._k:
            write_h5(self.fname, self._k, self._v)
        return False
