"""Minimal pure-Python HDF5 reader/writer for Keras `save_weights` files (no h5py in this image).

Reads what Keras 2.1.6 wrote for the reference's weight files (model.py:38-41 `save_weights`,
SURVEY.md appendix B): superblock v0, old-style groups (symbol-table message -> v1 B-tree + local heap
+ SNOD), v1 object headers with continuation blocks, contiguous little-endian IEEE float datasets.
`read_weights(path)` returns {"conv2d_1/kernel": ndarray, ...}.
"""
import struct

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(ValueError):
    pass


class _File:
    def __init__(self, data):
        if data[:8] != _SIG:
            raise H5Error("not an HDF5 file")
        if data[8] != 0 or data[13] != 8 or data[14] != 8:
            raise H5Error("only superblock v0 with 8-byte offsets/lengths is supported")
        self.d = data
        self.base = struct.unpack_from("<Q", data, 24)[0]
        # root symbol-table entry starts at 56: name offset, object header address
        self.root_header = struct.unpack_from("<Q", data, 64)[0]

    # -- object header v1 ---------------------------------------------------------------------------
    def messages(self, addr):
        d = self.d
        version, _, nmsg, _, hsize = struct.unpack_from("<BBHII", d, addr)
        if version != 1:
            raise H5Error("only v1 object headers are supported")
        out = []
        blocks = [(addr + 16, hsize)]
        while blocks and len(out) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", d, pos)
                body = pos + 8
                if mtype == 0x10:                      # continuation
                    off, length = struct.unpack_from("<QQ", d, body)
                    blocks.append((off + self.base, length))
                out.append((mtype, body, msize))
                pos = body + msize
        return out

    # -- old-style group traversal ------------------------------------------------------------------
    def group_entries(self, addr):
        """name -> object header address for the group whose header is at addr."""
        for mtype, body, _ in self.messages(addr):
            if mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", self.d, body)
                return self._walk_btree(btree + self.base, self._heap_data(heap + self.base))
        return None

    def _heap_data(self, addr):
        if self.d[addr:addr + 4] != b"HEAP":
            raise H5Error("bad local heap")
        return struct.unpack_from("<Q", self.d, addr + 24)[0] + self.base

    def _walk_btree(self, addr, heap):
        d = self.d
        if d[addr:addr + 4] == b"SNOD":
            n = struct.unpack_from("<H", d, addr + 6)[0]
            out = {}
            for i in range(n):
                e = addr + 8 + 40 * i
                name_off, hdr = struct.unpack_from("<QQ", d, e)
                s = heap + name_off
                out[d[s:d.index(b"\x00", s)].decode()] = hdr + self.base
            return out
        if d[addr:addr + 4] != b"TREE":
            raise H5Error("bad B-tree node")
        _ntype, level, used = struct.unpack_from("<BBH", d, addr + 4)
        out = {}
        pos = addr + 24            # after signature(4) type(1) level(1) entries(2) left(8) right(8)
        pos += 8                   # key 0
        for _ in range(used):
            child = struct.unpack_from("<Q", d, pos)[0]
            out.update(self._walk_btree(child + self.base, heap))
            pos += 16              # child pointer + next key
        return out

    # -- dataset --------------------------------------------------------------------------------------
    def dataset(self, addr):
        d = self.d
        shape = dtype = data = None
        for mtype, body, _ in self.messages(addr):
            if mtype == 0x01:                                  # dataspace
                ver, rank, flags = struct.unpack_from("<BBB", d, body)
                off = body + (8 if ver == 1 else 4)
                shape = struct.unpack_from("<%dQ" % rank, d, off) if rank else ()
            elif mtype == 0x03:                                # datatype
                cls = d[body] & 0x0F
                size = struct.unpack_from("<I", d, body + 4)[0]
                big = d[body + 1] & 1
                if cls == 1:
                    dtype = np.dtype(("<", ">")[big] + "f%d" % size)
                elif cls == 0:
                    signed = (d[body + 1] >> 3) & 1
                    dtype = np.dtype(("<", ">")[big] + ("u", "i")[signed] + "%d" % size)
                else:
                    dtype = None
            elif mtype == 0x08:                                # layout
                ver, cls = d[body], d[body + 1]
                if ver != 3 or cls != 1:
                    raise H5Error("only contiguous v3 layouts are supported")
                a, size = struct.unpack_from("<QQ", d, body + 2)
                data = (a + self.base, size)
        if shape is None or dtype is None or data is None:
            return None
        n = int(np.prod(shape)) if shape else 1
        if data[0] == _UNDEF + self.base or n == 0:
            return np.zeros(shape, dtype=dtype)
        return np.frombuffer(self.d, dtype=dtype, count=n, offset=data[0]).reshape(shape).copy()


def read_tree(path):
    """Returns {"/a/b/c": ndarray} for every dataset of the file."""
    with open(path, "rb") as f:
        h5 = _File(f.read())
    out = {}

    def walk(addr, prefix):
        entries = h5.group_entries(addr)
        if entries is None:
            arr = h5.dataset(addr)
            if arr is not None:
                out[prefix] = arr
            return
        for name, child in entries.items():
            walk(child, prefix + "/" + name)
    walk(h5.root_header + h5.base, "")
    return out


def read_weights(path):
    """Keras save_weights layout '/<layer>/<layer>/<param>:0'  ->  {'<layer>/<param>': float32 array}."""
    out = {}
    for k, v in read_tree(path).items():
        parts = k.strip("/").split("/")
        if len(parts) == 3 and parts[0] == parts[1] and parts[2].endswith(":0"):
            out["%s/%s" % (parts[0], parts[2][:-2])] = np.asarray(v, dtype=np.float32)
    if not out:
        raise H5Error("no Keras weights found in %s" % path)
    return out
