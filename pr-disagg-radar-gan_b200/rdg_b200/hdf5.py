"""Minimal pure-Python HDF5 reader / writer for Keras `model.save(...h5)` weight files.

The reference loads `trained_models/gen_*.h5` with tf.keras.models.load_model
(raindisagg_gan_pretrained.py:43-45) and writes checkpoints with generator.save / critic.save
(gan_train_cwgangp_pixelnorm.py:520-521).  Neither h5py nor libhdf5 exist in this image, so
this module implements the subset of the HDF5 file format those files use (h5py 2.10 /
HDF5 1.10.4, libver='earliest'; pr-disagg-env.yml:40-41):

  superblock v0/v1, symbol-table groups (v1 B-tree + SNOD + local heap), v1 object headers
  (+ continuation blocks), dataspace v1/v2, datatypes: IEEE float, fixed-point, fixed-length
  and variable-length strings (global heap), contiguous / compact layouts, attributes v1-v3.

Anything else (chunked / filtered datasets, v2 object headers, dense links) raises
NotImplementedError naming the feature.  UNVERIFIED against a real Keras file: the shipped
.h5 blobs are absent from /root/reference and no HDF5 library is available to produce one;
the reader is exercised on files produced by the writer below, which follows the same spec.
"""
from __future__ import annotations

import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(n):
    return (n + 7) & ~7


# ============================================================================ reader
class _Reader:
    def __init__(self, buf):
        # The superblock sits at offset 0 or, after a user block, at 512, 1024, 2048, ... (HDF5 spec III.A); every file
        # address is relative to the superblock's base address, which is the user-block size for files libhdf5 writes.
        sb = 0
        while buf[sb:sb + 8] != SIG:
            sb = 512 if sb == 0 else sb * 2
            if sb + 8 > len(buf):
                raise ValueError("not an HDF5 file (bad signature)")
        ver = buf[sb + 8]
        if ver not in (0, 1):
            raise NotImplementedError(f"HDF5 superblock version {ver} (only 0/1: libver='earliest')")
        so, sl = buf[sb + 13], buf[sb + 14]
        if so != 8 or sl != 8:
            raise NotImplementedError("HDF5 files with offset/length size != 8")
        pos = sb + (24 if ver == 0 else 28)
        self.base = struct.unpack_from("<Q", buf, pos)[0]
        if self.base > sb:
            raise ValueError("HDF5 base address beyond the superblock")
        self.b = buf[self.base:] if self.base else buf      # addresses below index this view directly
        root_entry = pos - self.base + 32
        self.root = self._symtab_entry(root_entry)

    # -- low level
    def u(self, fmt, off):
        return struct.unpack_from("<" + fmt, self.b, off)

    def _symtab_entry(self, off):
        name_off, ohdr, cache, _ = self.u("QQII", off)
        scratch = self.b[off + 24:off + 40]
        return dict(name_off=name_off, ohdr=ohdr, cache=cache, scratch=scratch)

    def _heap_name(self, heap_addr, off):
        assert self.b[heap_addr:heap_addr + 4] == b"HEAP", "bad local heap"
        data_addr = self.u("Q", heap_addr + 24)[0]
        s = data_addr + off
        e = self.b.index(b"\0", s)
        return self.b[s:e].decode("utf8")

    # -- object headers
    def messages(self, addr):
        """Yield (type, flags, data_offset, size) of a v1 object header, following continuations."""
        if self.b[addr:addr + 4] == b"OHDR":
            raise NotImplementedError("HDF5 v2 object headers (file written with libver='latest')")
        ver, _, nmsg, _, hsize = self.u("BBHII", addr)
        if ver != 1:
            raise NotImplementedError(f"object header version {ver}")
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, mflags = self.u("HHB", p)
                d = p + 8
                if mtype == 0x0010:
                    caddr, clen = self.u("QQ", d)
                    blocks.append((caddr, clen))
                out.append((mtype, mflags, d, msize))
                p = d + msize
        return out

    # -- datatype / dataspace
    def _datatype(self, off):
        cv = self.b[off]
        cls, ver = cv & 0x0F, cv >> 4
        bits = self.b[off + 1:off + 4]
        size = self.u("I", off + 4)[0]
        if cls == 1:
            if bits[0] & 1:
                raise NotImplementedError("big-endian float datasets")
            return dict(kind="float", size=size, np=np.dtype("<f%d" % size))
        if cls == 0:
            signed = bool(bits[0] & 0x08)
            return dict(kind="int", size=size, np=np.dtype("<%s%d" % ("i" if signed else "u", size)))
        if cls == 3:
            return dict(kind="string", size=size, np=np.dtype("S%d" % size))
        if cls == 9:
            vtype = bits[0] & 0x0F          # 0 sequence, 1 string
            return dict(kind="vlen_str" if vtype == 1 else "vlen", size=size, np=None)
        raise NotImplementedError(f"HDF5 datatype class {cls}")

    def _dataspace(self, off):
        ver = self.b[off]
        rank = self.b[off + 1]
        if ver == 1:
            p = off + 8
        elif ver == 2:
            p = off + 4
        else:
            raise NotImplementedError(f"dataspace version {ver}")
        return tuple(self.u("Q" * rank, p)) if rank else ()

    def _vlen_string(self, off):
        length, gaddr, idx = self.u("IQI", off)
        if self.b[gaddr:gaddr + 4] != b"GCOL":
            raise ValueError("bad global heap collection")
        csize = self.u("Q", gaddr + 8)[0]
        p, end = gaddr + 16, gaddr + csize
        while p + 16 <= end:
            oidx, _, _, osize = self.u("HHIQ", p)
            if oidx == idx:
                return bytes(self.b[p + 16:p + 16 + length])
            if oidx == 0:
                break
            p += 16 + _pad8(osize)
        raise ValueError("global heap object not found")

    def _decode(self, dt, shape, data_off):
        n = int(np.prod(shape)) if shape else 1
        if dt["kind"] == "vlen_str":
            vals = [self._vlen_string(data_off + 16 * i) for i in range(n)]
            arr = np.array(vals, dtype=object)
            return arr.reshape(shape) if shape else arr.reshape(())[()]
        if dt["np"] is None:
            raise NotImplementedError("variable-length sequence data")
        arr = np.frombuffer(self.b, dtype=dt["np"], count=n, offset=data_off)
        return arr.reshape(shape).copy() if shape else arr[0]

    def attributes(self, ohdr):
        out = {}
        for mtype, _, d, msize in self.messages(ohdr):
            if mtype != 0x000C:
                continue
            ver = self.b[d]
            nsz, dtsz, dssz = self.u("HHH", d + 2)
            p = d + 8
            if ver == 3:
                p += 1  # name character set encoding
            name = bytes(self.b[p:p + nsz]).split(b"\0")[0].decode("utf8")
            if ver == 1:
                p += _pad8(nsz); dt = self._datatype(p); p += _pad8(dtsz); shape = self._dataspace(p); p += _pad8(dssz)
            elif ver in (2, 3):
                p += nsz; dt = self._datatype(p); p += dtsz; shape = self._dataspace(p); p += dssz
            else:
                raise NotImplementedError(f"attribute message version {ver}")
            out[name] = self._decode(dt, shape, p)
        return out

    # -- groups
    def _group_tables(self, ohdr):
        for mtype, _, d, _ in self.messages(ohdr):
            if mtype == 0x0011:
                return self.u("QQ", d)
            if mtype in (0x0002, 0x0006):
                raise NotImplementedError("new-style (link message) groups: file written with libver='latest'")
        return None

    def is_group(self, ohdr):
        return any(m[0] == 0x0011 for m in self.messages(ohdr))

    def children(self, ohdr):
        """Ordered dict name -> object header address (B-tree order = sorted by name)."""
        tabs = self._group_tables(ohdr)
        if tabs is None:
            raise ValueError("object is not a group")
        btree, heap = tabs
        out = {}
        self._walk_btree(btree, heap, out)
        return out

    def _walk_btree(self, addr, heap, out):
        if self.b[addr:addr + 4] != b"TREE":
            raise ValueError("bad group B-tree node")
        ntype, level, used = self.u("BBH", addr + 4)
        if ntype != 0:
            raise ValueError("unexpected B-tree node type")
        p = addr + 24
        for i in range(used):
            child = self.u("Q", p + 8 + 16 * i)[0]
            if level > 0:
                self._walk_btree(child, heap, out)
            else:
                if self.b[child:child + 4] != b"SNOD":
                    raise ValueError("bad symbol table node")
                nsym = self.u("H", child + 6)[0]
                for k in range(nsym):
                    e = self._symtab_entry(child + 8 + 40 * k)
                    out[self._heap_name(heap, e["name_off"])] = e["ohdr"]

    # -- datasets
    def dataset(self, ohdr):
        dt = shape = None
        layout = None
        for mtype, _, d, msize in self.messages(ohdr):
            if mtype == 0x0001:
                shape = self._dataspace(d)
            elif mtype == 0x0003:
                dt = self._datatype(d)
            elif mtype == 0x0008:
                ver = self.b[d]
                if ver == 3:
                    cls = self.b[d + 1]
                    if cls == 1:
                        layout = ("contig", self.u("Q", d + 2)[0])
                    elif cls == 0:
                        layout = ("compact", d + 4)
                    else:
                        raise NotImplementedError("chunked dataset layout (Keras weight files are contiguous)")
                elif ver in (1, 2):
                    rank, cls = self.b[d + 1], self.b[d + 2]
                    if cls == 1:
                        layout = ("contig", self.u("Q", d + 8)[0])
                    else:
                        raise NotImplementedError("layout message v1/v2 non-contiguous")
                else:
                    raise NotImplementedError(f"data layout version {ver}")
            elif mtype == 0x000B:
                raise NotImplementedError("filtered (compressed) datasets")
        if dt is None or shape is None or layout is None:
            raise ValueError("object is not a dataset")
        kind, off = layout
        if kind == "contig":
            if off == UNDEF:
                return np.zeros(shape, dt["np"])
        return self._decode(dt, shape, off)


class H5File:
    """Read-only view: f['a/b/c'] -> numpy array or H5Group; .attrs on groups and the file."""

    def __init__(self, path_or_bytes):
        if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
            buf = bytes(path_or_bytes)
        else:
            with open(path_or_bytes, "rb") as f:
                buf = f.read()
        self._r = _Reader(buf)
        self.root = H5Group(self._r, self._r.root["ohdr"], "/")

    def __getitem__(self, k):
        return self.root[k]

    @property
    def attrs(self):
        return self.root.attrs

    def keys(self):
        return self.root.keys()


class H5Group:
    def __init__(self, r, ohdr, name):
        self._r, self._ohdr, self.name = r, ohdr, name
        self._children = None

    def _kids(self):
        if self._children is None:
            self._children = self._r.children(self._ohdr)
        return self._children

    def keys(self):
        return list(self._kids().keys())

    def __contains__(self, k):
        return k in self._kids()

    @property
    def attrs(self):
        return self._r.attributes(self._ohdr)

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            kids = node._kids()
            if part not in kids:
                raise KeyError(f"{part!r} not in {node.name}")
            o = kids[part]
            node = H5Group(self._r, o, node.name.rstrip("/") + "/" + part) if self._r.is_group(o) else _Leaf(self._r, o)
        return node.value() if isinstance(node, _Leaf) else node


class _Leaf:
    def __init__(self, r, ohdr):
        self._r, self._ohdr = r, ohdr

    def value(self):
        return self._r.dataset(self._ohdr)

    def _kids(self):
        raise KeyError("dataset has no children")


def _as_str(x):
    return x.decode("utf8") if isinstance(x, (bytes, np.bytes_)) else str(x)


def load_keras_weights(path):
    """Return the weight arrays of a Keras .h5 model file in `model.get_weights()` order.

    Walks `model_weights` using the `layer_names` / `weight_names` attributes: layer
    order and weight order come from those attributes, not from (counter-dependent) layer names.
    Also accepts weight-only files (`model.save_weights`), where the root is the weights group.
    """
    f = H5File(path)
    g = f["model_weights"] if "model_weights" in f.root else f.root
    attrs = g.attrs
    if "layer_names" not in attrs:
        raise ValueError("no layer_names attribute: not a Keras weight file")
    out = []
    for ln in np.atleast_1d(attrs["layer_names"]):
        lg = g[_as_str(ln)]
        wn = lg.attrs.get("weight_names", [])
        for w in np.atleast_1d(wn):
            out.append(np.asarray(lg[_as_str(w)], dtype=np.float32))
    return out


# ============================================================================ writer
class _Writer:
    LEAF_K, INT_K = 32, 16   # symbol-table node holds up to 2*LEAF_K entries

    def __init__(self):
        self.buf = bytearray(96)      # superblock placeholder

    def alloc(self, data, align=8):
        while len(self.buf) % align:
            self.buf.append(0)
        off = len(self.buf)
        self.buf += data
        return off

    # -- messages
    @staticmethod
    def _msg(mtype, data, flags=0):
        data = bytes(data) + b"\0" * (_pad8(len(data)) - len(data))
        return struct.pack("<HHB3x", mtype, len(data), flags) + data

    @staticmethod
    def _dtype_msg(dt):
        dt = np.dtype(dt)
        if dt.kind == "f" and dt.itemsize == 4:     # IEEE binary32 little-endian
            return struct.pack("<BBBBI", 0x11, 0x20, 0x1F, 0x00, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
        if dt.kind == "f" and dt.itemsize == 8:     # IEEE binary64 little-endian
            return struct.pack("<BBBBI", 0x11, 0x20, 0x3F, 0x00, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
        if dt.kind in "iu":
            b0 = 0x08 if dt.kind == "i" else 0x00
            return struct.pack("<BBBBI", 0x10, b0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
        if dt.kind == "S":
            return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, dt.itemsize)   # null-padded ASCII
        raise NotImplementedError(f"dtype {dt}")

    @staticmethod
    def _dspace_msg(shape):
        return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", s) for s in shape)

    def _attr_msg(self, name, value):
        arr = np.asarray(value)
        if arr.dtype.kind == "U":
            arr = np.char.encode(arr, "utf8")
        if arr.dtype.kind == "O":
            arr = np.array([v if isinstance(v, bytes) else str(v).encode("utf8") for v in arr.ravel()]).reshape(arr.shape)
        nm = name.encode("utf8") + b"\0"
        dt, ds = self._dtype_msg(arr.dtype), self._dspace_msg(arr.shape)
        body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds))
        body += nm + b"\0" * (_pad8(len(nm)) - len(nm))
        body += dt + b"\0" * (_pad8(len(dt)) - len(dt))
        body += ds + b"\0" * (_pad8(len(ds)) - len(ds))
        body += np.ascontiguousarray(arr).tobytes()
        if len(body) > 64000:
            raise ValueError(f"attribute {name!r} too large for a v1 object header")
        return self._msg(0x000C, body)

    def _ohdr(self, msgs):
        body = b"".join(msgs)
        return self.alloc(struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body)

    # -- objects
    def dataset(self, arr):
        arr = np.ascontiguousarray(arr)
        data_off = self.alloc(arr.tobytes()) if arr.size else UNDEF
        msgs = [self._msg(0x0001, self._dspace_msg(arr.shape)),
                self._msg(0x0003, self._dtype_msg(arr.dtype), flags=1),
                self._msg(0x0005, struct.pack("<BBBB", 2, 2, 0, 0)),
                self._msg(0x0008, struct.pack("<BBQQ", 3, 1, data_off, arr.nbytes))]
        return self._ohdr(msgs), None

    def group(self, children, attrs):
        """children: dict name -> (ohdr, scratch or None). Returns (ohdr, (btree, heap))."""
        names = sorted(children.keys(), key=lambda s: s.encode("utf8"))
        if len(names) > 2 * self.LEAF_K:
            raise NotImplementedError("groups with more than 64 members")
        # local heap data segment: offset 0 = empty string
        seg = bytearray(8)
        offs = {}
        for n in names:
            offs[n] = len(seg)
            nb = n.encode("utf8") + b"\0"
            seg += nb + b"\0" * (_pad8(len(nb)) - len(nb))
        free_off = len(seg)
        seg += struct.pack("<QQ", 1, 32) + b"\0" * 16          # one free block (next = none, size 32)
        seg_addr = self.alloc(bytes(seg))
        heap = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(seg), free_off, seg_addr))
        # symbol table node
        snod = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(names)))
        for n in names:
            ohdr, scratch = children[n]
            if scratch is not None:
                snod += struct.pack("<QQII", offs[n], ohdr, 1, 0) + struct.pack("<QQ", *scratch)
            else:
                snod += struct.pack("<QQII", offs[n], ohdr, 0, 0) + b"\0" * 16
        snod += b"\0" * (8 + 40 * 2 * self.LEAF_K - len(snod))
        snod_addr = self.alloc(bytes(snod))
        node = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, UNDEF, UNDEF))
        node += struct.pack("<Q", 0)
        if names:
            node += struct.pack("<QQ", snod_addr, offs[names[-1]])
        node += b"\0" * (24 + (2 * self.INT_K + 1) * 8 + 2 * self.INT_K * 8 - len(node))
        btree = self.alloc(bytes(node))
        msgs = [self._msg(0x0011, struct.pack("<QQ", btree, heap))]
        msgs += [self._attr_msg(k, v) for k, v in (attrs or {}).items()]
        return self._ohdr(msgs), (btree, heap)

    def finish(self, root):
        ohdr, (btree, heap) = root
        sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.LEAF_K, self.INT_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, ohdr, 1, 0) + struct.pack("<QQ", btree, heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def _build(w, tree):
    """tree: {"attrs": {...}, "children": {name: ndarray | tree}}"""
    kids = {}
    for name, v in tree.get("children", {}).items():
        kids[name] = _build(w, v) if isinstance(v, dict) else w.dataset(v)
    return w.group(kids, tree.get("attrs"))


def write_h5(path, tree):
    w = _Writer()
    data = w.finish(_build(w, tree))
    if path is not None:
        with open(path, "wb") as f:
            f.write(data)
    return data


def save_keras_weights(path, weights, kind="generator", keras_version="2.2.4-tf", model_config=None):
    """Write weights in the tree Keras 2.2.4-tf `model.save` produces for the reference's models
    (SURVEY A2): model_weights/<sequential>/<layer>/{kernel:0,bias:0} with layer_names /
    weight_names attributes.  `kind`: "generator" (Dense + 4 Conv3D) or "critic" (4 Conv3D + Dense)."""
    if kind == "generator":
        seq, names = "sequential", ["dense", "conv3d", "conv3d_1", "conv3d_2", "conv3d_3"]
        top = ["input_1", "input_2", "flatten", "concatenate", seq]
    elif kind == "critic":
        seq, names = "sequential_1", ["conv3d_4", "conv3d_5", "conv3d_6", "conv3d_7", "dense_1"]
        top = ["input_3", "reshape", "input_4", "lambda", "concatenate_1", seq]
    else:
        raise ValueError(kind)
    if len(weights) != 10:
        raise ValueError("expected 10 weight tensors")
    layers, wnames = {}, []
    for i, ln in enumerate(names):
        layers[ln] = {"children": {"kernel:0": np.asarray(weights[2 * i], np.float32),
                                   "bias:0": np.asarray(weights[2 * i + 1], np.float32)}}
        wnames += [f"{ln}/kernel:0", f"{ln}/bias:0"]
    mw = {}
    for t in top:
        if t == seq:
            mw[t] = {"attrs": {"weight_names": np.array([n.encode() for n in wnames])}, "children": layers}
        else:
            mw[t] = {"attrs": {"weight_names": np.zeros((0,), "S1")}}
    root_attrs = {"keras_version": np.bytes_(keras_version.encode()), "backend": np.bytes_(b"tensorflow")}
    if model_config is not None:
        root_attrs["model_config"] = np.bytes_(model_config.encode())
    tree = {"attrs": root_attrs,
            "children": {"model_weights": {"attrs": {"layer_names": np.array([t.encode() for t in top]),
                                                     "backend": np.bytes_(b"tensorflow"),
                                                     "keras_version": np.bytes_(keras_version.encode())},
                                           "children": mw}}}
    return write_h5(path, tree)
