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

    # -- attributes (message 0x0C, version 1): fixed-length string arrays and numeric arrays ---------------
    def attributes(self, addr):
        d = self.d
        out = {}
        for mtype, body, _ in self.messages(addr):
            if mtype != 0x0C or d[body] != 1:
                continue
            nsz, tsz, ssz = struct.unpack_from("<HHH", d, body + 2)
            pad = lambda x: (x + 7) & ~7
            pos = body + 8
            name = d[pos:pos + nsz].split(b"\x00")[0].decode()
            pos += pad(nsz)
            cls, size = d[pos] & 0x0F, struct.unpack_from("<I", d, pos + 4)[0]
            big = d[pos + 1] & 1
            tpos = pos
            pos += pad(tsz)
            ver, rank = d[pos], d[pos + 1]
            dims = struct.unpack_from("<%dQ" % rank, d, pos + (8 if ver == 1 else 4)) if rank else ()
            pos += pad(ssz)
            n = int(np.prod(dims)) if dims else 1
            if cls == 3:
                vals = [d[pos + i * size:pos + (i + 1) * size].split(b"\x00")[0] for i in range(n)]
                out[name] = vals if dims else vals[0]
            elif cls in (0, 1):
                kind = "f" if cls == 1 else ("i" if (d[tpos + 1] >> 3) & 1 else "u")
                out[name] = np.frombuffer(d, dtype=np.dtype(("<", ">")[big] + kind + "%d" % size), count=n, offset=pos).reshape(dims).copy()
            else:
                out[name] = None                        # variable-length strings etc.: present, not decoded
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


def _canonical_layer_names(names):
    """Keras numbers auto-named layers per process, not per model: the second ResidualCNN built in a process owns conv2d_31..60,
    batch_normalization_31..60 and dense_2.  `load_weights` matches layers by position, so such files load in the reference; here
    each auto-named family is renumbered from 1 in ascending suffix order (= creation = topological order)."""
    import re
    fam = {}
    for n in names:
        m = re.fullmatch(r"(conv2d|batch_normalization|dense)_(\d+)", n)
        if m:
            fam.setdefault(m.group(1), []).append(int(m.group(2)))
    ren = {}
    for f, idx in fam.items():
        for new, old in enumerate(sorted(set(idx)), start=1):
            ren["%s_%d" % (f, old)] = "%s_%d" % (f, new)
    return ren


def read_weights(path):
    """Keras save_weights layout '/<layer>/<layer>/<param>:0' (optionally under '/model_weights', which is where a full
    `model.save` file keeps it)  ->  {'<layer>/<param>': float32 array} with canonical layer numbering."""
    raw = {}
    for k, v in read_tree(path).items():
        parts = k.strip("/").split("/")
        if parts and parts[0] == "model_weights":
            parts = parts[1:]
        if len(parts) == 3 and parts[0] == parts[1] and parts[2].endswith(":0"):
            raw[(parts[0], parts[2][:-2])] = np.asarray(v, dtype=np.float32)
    if not raw:
        raise H5Error("no Keras weights found in %s" % path)
    ren = _canonical_layer_names({layer for layer, _ in raw})
    return {"%s/%s" % (ren.get(layer, layer), param): v for (layer, param), v in raw.items()}


# =====================================================================================================
# Writer: the same subset, enough for Keras `save_weights` files (model.py:38-41) and the reference's training-data
# files (utils.save_train_data, utils.py:48-56): superblock v0, old-style groups (one symbol-table node per group:
# the superblock's group-leaf K is sized for the largest group), v1 object headers, contiguous little-endian
# datasets, v1 attribute messages holding numeric arrays or fixed-length string arrays (what h5py writes for a list
# of bytes — Keras' layer_names / weight_names).

def _pad8(b):
    return b + b"\x00" * (-len(b) % 8)


def _msg(mtype, body, flags=0):
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, dt.itemsize)                       # string, null-padded, ASCII
    if dt.kind == "f":
        spec = {4: (31, 23, 8, 0, 23, 127), 8: (63, 52, 11, 0, 52, 1023)}[dt.itemsize]
        sign, eloc, esz, mloc, msz, bias = spec
        return (struct.pack("<BBBBI", 0x11, 0x20, sign, 0, dt.itemsize) +
                struct.pack("<HHBBBBI", 0, dt.itemsize * 8, eloc, esz, mloc, msz, bias))
    if dt.kind in "iu":
        return struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
    raise H5Error("unsupported dtype %s" % dt)


def _space_msg(shape):
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(x)) for x in shape)


def _attr_msg(name, value):
    if isinstance(value, (list, tuple)) and (len(value) == 0 or isinstance(value[0], (bytes, str))):
        vals = [v.encode() if isinstance(v, str) else v for v in value]
        arr = np.array(vals, dtype="S%d" % max([len(v) for v in vals] + [1])) if vals else np.zeros((0,), dtype=np.float64)
    elif isinstance(value, (bytes, str)):
        arr = np.array(value.encode() if isinstance(value, str) else value)
    else:
        arr = np.asarray(value)
    shape = arr.shape                              # (np.ascontiguousarray would promote a scalar to shape (1,))
    arr = np.ascontiguousarray(arr)
    if arr.dtype.byteorder == ">":
        arr = arr.astype(arr.dtype.newbyteorder("<"))
    nm = name.encode() + b"\x00"
    t, sp = _dtype_msg(arr.dtype), _space_msg(shape)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(t), len(sp)) + _pad8(nm) + _pad8(t) + _pad8(sp) + arr.tobytes()
    return _msg(0x0C, body)


class _Writer:
    def __init__(self, leaf_k):
        self.buf = bytearray(96)                  # superblock (56) + root symbol-table entry (40), filled in at the end
        self.leaf_k = leaf_k

    def alloc(self, data, align=8):
        self.buf += b"\x00" * (-len(self.buf) % align)
        addr = len(self.buf)
        self.buf += data
        return addr

    def header(self, messages):
        body = b"".join(messages)
        return self.alloc(struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body)

    def dataset(self, arr, attrs):
        shape = np.shape(arr)
        arr = np.ascontiguousarray(arr)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        data = self.alloc(arr.tobytes()) if arr.size else _UNDEF
        msgs = [_msg(0x01, _space_msg(shape)), _msg(0x03, _dtype_msg(arr.dtype), flags=1),
                _msg(0x05, struct.pack("<BBBB", 2, 2, 2, 0)),                              # fill value v2: not defined
                _msg(0x08, struct.pack("<BBQQ", 3, 1, data, arr.nbytes))]                  # contiguous layout v3
        return self.header(msgs + [_attr_msg(k, v) for k, v in attrs.items()])

    def group(self, children, attrs):
        """children: {name: object header address}; returns the group's object header address"""
        names = sorted(children, key=lambda s: s.encode())
        if len(names) > 2 * self.leaf_k:
            raise H5Error("group too large for the symbol-table node size")
        heap = bytearray(8)                                   # offset 0: the empty string
        offs = {}
        for nm in names:
            offs[nm] = len(heap)
            heap += _pad8(nm.encode() + b"\x00")
        free = len(heap)
        heap += struct.pack("<QQ", 1, 16)                     # one free block: next = 1 (end of list), size 16
        heap_data = self.alloc(bytes(heap))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free, heap_data))
        snod = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(names)))
        for nm in names:
            snod += struct.pack("<QQII16x", offs[nm], children[nm], 0, 0)
        snod += b"\x00" * (8 + 2 * self.leaf_k * 40 - len(snod))
        snod_addr = self.alloc(bytes(snod))
        last = offs[names[-1]] if names else 0
        tree = (b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, _UNDEF, _UNDEF) +
                struct.pack("<QQQ", 0, snod_addr, last))
        tree += b"\x00" * (24 + (2 * 16 + 1) * 8 + 2 * 16 * 8 - len(tree))        # full node for internal K = 16
        tree_addr = self.alloc(tree)
        hdr = self.header([_msg(0x11, struct.pack("<QQ", tree_addr, heap_addr))] + [_attr_msg(k, v) for k, v in attrs.items()])
        return hdr, tree_addr, heap_addr

    def finish(self, root_hdr, root_tree, root_heap):
        eof = len(self.buf)
        sb = _SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.leaf_k, 16, 0)
        sb += struct.pack("<QQQQ", 0, _UNDEF, eof, _UNDEF)
        sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", root_tree, root_heap)     # root entry, cached btree/heap
        self.buf[:len(sb)] = sb
        return bytes(self.buf)


def write_tree(path, tree, attrs=None):
    """tree: nested dict {name: ndarray | dict}; attrs: {"/path/to/object": {attr name: value}} ("/" = root)."""
    attrs = attrs or {}

    def max_children(node):
        return max([len(node)] + [max_children(v) for v in node.values() if isinstance(v, dict)])
    w = _Writer(max(4, (max_children(tree) + 1) // 2))

    def emit(node, prefix):
        kids = {}
        for name, v in node.items():
            p = prefix + "/" + name
            kids[name] = emit(v, p)[0] if isinstance(v, dict) else w.dataset(v, attrs.get(p, {}))
        return w.group(kids, attrs.get(prefix or "/", {}))
    data = w.finish(*emit(tree, ""))
    with open(path, "wb") as f:
        f.write(data)
    return path


# model.layers order of the reference's ResidualCNN (model.py:58-145) as Keras 2.1.6 names them: load_weights matches
# weights to layers by this order (keras/engine/topology.py load_weights_from_hdf5_group)
def keras_layer_order():
    names = ["input_1", "conv2d_1", "batch_normalization_1", "activation_1"]
    act = 1
    for b in range(9):
        for j in range(3):
            i = 2 + 3 * b + j
            names += ["conv2d_%d" % i, "batch_normalization_%d" % i]
            if j == 2:
                names.append("add_%d" % (b + 1))
            act += 1
            names.append("activation_%d" % act)
    names += ["conv2d_30", "conv2d_29", "batch_normalization_30", "batch_normalization_29", "activation_30", "activation_29",
              "flatten_2", "flatten_1", "dense_1", "policy_head", "value_head"]
    return names


_PARAM_ORDER = {"conv2d": ["kernel", "bias"], "batch_normalization": ["gamma", "beta", "moving_mean", "moving_variance"],
                "dense": ["kernel", "bias"], "policy_head": ["kernel", "bias"], "value_head": ["kernel", "bias"]}


def write_weights(path, weights, keras_version="2.1.6", backend="tensorflow"):
    """{'<layer>/<param>': array} -> a Keras `save_weights` file: /<layer>/<layer>/<param>:0 float32 datasets, root attribute
    layer_names (model.layers order), per-layer attribute weight_names."""
    tree, attrs = {}, {"/": {"layer_names": [n.encode() for n in keras_layer_order()], "backend": backend.encode(),
                             "keras_version": keras_version.encode()}}
    have = {k.split("/")[0] for k in weights}
    for layer in keras_layer_order():
        if layer in have:
            kind = next(k for k in _PARAM_ORDER if layer.startswith(k))
            params = [p for p in _PARAM_ORDER[kind] if "%s/%s" % (layer, p) in weights]
            tree[layer] = {layer: {p + ":0": np.asarray(weights["%s/%s" % (layer, p)], dtype=np.float32) for p in params}}
            attrs["/" + layer] = {"weight_names": [("%s/%s:0" % (layer, p)).encode() for p in params]}
        else:
            tree[layer] = {}
            attrs["/" + layer] = {"weight_names": []}
    return write_tree(path, tree, attrs)
