"""A dependency-free reader for the subset of HDF5 that Keras weight files use (model.save() / model.save_weights() through h5py with
default settings; the reference loads them with model.load_weights, deeplabv3p/model.py:102-103).

Written from the HDF5 File Format Specification (version 3.0 of the document, the layouts libhdf5 1.8 / 1.10 write by default):
  * superblock versions 0-3
  * object headers version 1 (with continuation blocks) and version 2 ('OHDR')
  * old-style groups: symbol-table message -> version-1 B-tree ('TREE', node type 0, any depth) -> symbol nodes ('SNOD') with names in a
    local heap ('HEAP'); new-style compact groups: link messages (hard links).  Dense groups (fractal heaps) are not supported.
  * datasets with contiguous or compact layout (data-layout message version 3; versions 1-2 for contiguous), little- or big-endian IEEE
    floats and fixed-point integers of 1-8 bytes, fixed-length strings.  Chunked / filtered datasets raise H5Error (Keras writes neither).
  * attributes (message versions 1-3) of those types plus variable-length strings through the global heap ('GCOL') — layer_names,
    weight_names, keras_version, backend.

The interface mirrors the part of h5py the exporters use: File / Group have keys(), __contains__, __getitem__ (paths with '/'), attrs;
Dataset has shape, dtype and ds[()] -> numpy array.  h5py is not part of this image, so the parser is checked against files produced by
the spec-following writer in tests/h5_writer.py (same default layout libhdf5 uses: superblock 0, symbol-table groups, contiguous data);
it has not been run against a file written by libhdf5 itself.
"""
from __future__ import annotations

import struct
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

SIGNATURE = b'\x89HDF\r\n\x1a\n'
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(ValueError):
    pass


class _Reader:
    def __init__(self, buf: bytes):
        self.buf = buf
        self.O = 8      # size of offsets
        self.L = 8      # size of lengths
        self.base = 0

    def u(self, off: int, n: int) -> int:
        if off < 0 or off + n > len(self.buf):
            raise H5Error('read of %d bytes at %d is outside the file (%d bytes)' % (n, off, len(self.buf)))
        return int.from_bytes(self.buf[off:off + n], 'little')

    def off(self, pos: int) -> int:
        return self.u(pos, self.O)

    def len_(self, pos: int) -> int:
        return self.u(pos, self.L)

    def bytes_(self, off: int, n: int) -> bytes:
        if off < 0 or off + n > len(self.buf):
            raise H5Error('read of %d bytes at %d is outside the file (%d bytes)' % (n, off, len(self.buf)))
        return self.buf[off:off + n]


def _pad8(n: int) -> int:
    return (n + 7) & ~7


# ------------------------------------------------------------------------------------------------ datatypes / dataspaces
class _Datatype:
    def __init__(self, cls: int, size: int, np_dtype: Optional[np.dtype], vlen_string: bool = False):
        self.cls, self.size, self.np_dtype, self.vlen_string = cls, size, np_dtype, vlen_string


def _parse_datatype(r: _Reader, pos: int) -> Tuple[_Datatype, int]:
    """-> (datatype, bytes consumed)."""
    b0 = r.u(pos, 1)
    cls, version = b0 & 0x0F, b0 >> 4
    bits = r.u(pos + 1, 3)
    size = r.u(pos + 4, 4)
    if version < 1 or version > 3:
        raise H5Error('datatype message version %d' % version)
    if cls == 0:      # fixed point: properties bit offset (2), precision (2)
        order = '>' if bits & 1 else '<'
        signed = bool(bits & 0x08)
        if size not in (1, 2, 4, 8):
            raise H5Error('fixed-point size %d' % size)
        return _Datatype(cls, size, np.dtype('%s%s%d' % (order, 'i' if signed else 'u', size))), 8 + 4
    if cls == 1:      # floating point: properties 12 bytes
        order = '>' if bits & 1 else '<'
        if size not in (2, 4, 8):
            raise H5Error('floating-point size %d' % size)
        return _Datatype(cls, size, np.dtype('%sf%d' % (order, size))), 8 + 12
    if cls == 3:      # fixed-length string, no properties
        return _Datatype(cls, size, np.dtype('S%d' % size)), 8
    if cls == 9:      # variable length: base type follows
        vtype = bits & 0x0F
        base, used = _parse_datatype(r, pos + 8)
        return _Datatype(cls, size, None, vlen_string=(vtype == 1)), 8 + used
    raise H5Error('datatype class %d is not supported (floats, integers and strings only)' % cls)


def _parse_dataspace(r: _Reader, pos: int) -> Tuple[Tuple[int, ...], int]:
    version = r.u(pos, 1)
    rank = r.u(pos + 1, 1)
    flags = r.u(pos + 2, 1)
    if version == 1:
        p = pos + 8
    elif version == 2:
        if r.u(pos + 3, 1) == 2:      # null dataspace
            return (0,), 4
        p = pos + 4
    else:
        raise H5Error('dataspace message version %d' % version)
    dims = tuple(r.len_(p + i * r.L) for i in range(rank))
    p += rank * r.L
    if flags & 1:
        p += rank * r.L
    return dims, p - pos


# ------------------------------------------------------------------------------------------------ object headers
class _Message:
    def __init__(self, mtype: int, pos: int, size: int):
        self.type, self.pos, self.size = mtype, pos, size


def _messages(r: _Reader, addr: int) -> List[_Message]:
    """Every header message of the object at `addr` (absolute), following continuation blocks."""
    a = addr + r.base
    out: List[_Message] = []
    if r.bytes_(a, 4) == b'OHDR':                       # version 2
        if r.u(a + 4, 1) != 2:
            raise H5Error('object header version %d' % r.u(a + 4, 1))
        flags = r.u(a + 5, 1)
        p = a + 6
        if flags & 0x20:
            p += 16                                      # four time stamps
        if flags & 0x10:
            p += 4                                       # max compact / min dense attributes
        csize = 1 << (flags & 3)
        chunk0 = r.u(p, csize)
        p += csize
        tracked = bool(flags & 0x04)
        blocks = [(p, chunk0)]
        while blocks:
            start, size = blocks.pop(0)
            q, end = start, start + size
            while q + 4 <= end:
                mtype = r.u(q, 1)
                msize = r.u(q + 1, 2)
                q += 4 + (2 if tracked else 0)
                if q + msize > end + 4:                  # the gap before the checksum
                    break
                if mtype == 0x10:
                    co, cl = r.off(q), r.len_(q + r.O)
                    if r.bytes_(co + r.base, 4) != b'OCHK':
                        raise H5Error('continuation block without OCHK signature')
                    blocks.append((co + r.base + 4, cl - 8))
                elif mtype != 0:
                    out.append(_Message(mtype, q, msize))
                q += msize
        return out
    version = r.u(a, 1)
    if version != 1:
        raise H5Error('object header version %d at %d' % (version, addr))
    nmsgs = r.u(a + 2, 2)
    size = r.u(a + 8, 4)
    blocks = [(a + 16, size)]
    while blocks and len(out) < 65536:
        start, bsize = blocks.pop(0)
        q, end = start, start + bsize
        while q + 8 <= end and nmsgs > 0:
            mtype = r.u(q, 2)
            msize = r.u(q + 2, 2)
            q += 8
            nmsgs -= 1
            if mtype == 0x10:
                blocks.append((r.off(q) + r.base, r.len_(q + r.O)))
            elif mtype != 0:
                out.append(_Message(mtype, q, msize))
            q += msize
    return out


def _global_heap_object(r: _Reader, heap_addr: int, index: int) -> bytes:
    a = heap_addr + r.base
    if r.bytes_(a, 4) != b'GCOL':
        raise H5Error('global heap collection without GCOL signature')
    size = r.len_(a + 8)
    p, end = a + 8 + r.L, a + size
    while p + 8 + r.L <= end:
        idx = r.u(p, 2)
        osize = r.len_(p + 8)
        if idx == 0:
            break
        if idx == index:
            return r.bytes_(p + 8 + r.L, osize)
        p += 8 + r.L + _pad8(osize)
    raise H5Error('global heap object %d not found' % index)


def _decode(r: _Reader, dt: _Datatype, dims: Tuple[int, ...], raw_pos: int):
    n = 1
    for d in dims:
        n *= d
    if dt.cls == 9:
        if not dt.vlen_string:
            raise H5Error('variable-length sequences are not supported')
        esize = 4 + r.O + 4
        vals = []
        for i in range(n):
            p = raw_pos + i * esize
            ln = r.u(p, 4)
            vals.append(_global_heap_object(r, r.off(p + 4), r.u(p + 4 + r.O, 4))[:ln].decode('utf-8'))
        return vals[0] if dims == () else np.array(vals, dtype=object).reshape(dims)
    a = np.frombuffer(r.bytes_(raw_pos, n * dt.size), dtype=dt.np_dtype, count=n).reshape(dims)
    if dt.cls == 3:
        a = np.char.rstrip(a, b'\x00') if a.size else a
    return a[()] if dims == () else a.copy()


class Dataset:
    def __init__(self, r: _Reader, name: str, msgs: List[_Message]):
        self._r, self.name = r, name
        self._dt: Optional[_Datatype] = None
        self.shape: Tuple[int, ...] = ()
        self._layout: Optional[Tuple[str, int, int]] = None
        self.attrs = _attributes(r, msgs)
        for m in msgs:
            if m.type == 0x03:
                self._dt, _ = _parse_datatype(r, m.pos)
            elif m.type == 0x01:
                self.shape, _ = _parse_dataspace(r, m.pos)
            elif m.type == 0x08:
                self._layout = self._parse_layout(m)
            elif m.type == 0x0B:
                raise H5Error('%s: filtered (compressed) datasets are not supported' % name)
        if self._dt is None or self._layout is None:
            raise H5Error('%s: not a dataset (no datatype / layout message)' % name)

    def _parse_layout(self, m: _Message) -> Tuple[str, int, int]:
        r = self._r
        version = r.u(m.pos, 1)
        if version == 3:
            cls = r.u(m.pos + 1, 1)
            if cls == 0:
                return 'compact', m.pos + 4, r.u(m.pos + 2, 2)
            if cls == 1:
                return 'contiguous', r.off(m.pos + 2), r.len_(m.pos + 2 + r.O)
            raise H5Error('%s: chunked layout is not supported (Keras writes contiguous datasets)' % self.name)
        if version in (1, 2):
            rank, cls = r.u(m.pos + 1, 1), r.u(m.pos + 2, 1)
            if cls != 1:
                raise H5Error('%s: layout version %d class %d is not supported' % (self.name, version, cls))
            return 'contiguous', r.off(m.pos + 8), 0
        raise H5Error('%s: data layout message version %d' % (self.name, version))

    @property
    def dtype(self) -> np.dtype:
        return self._dt.np_dtype if self._dt.np_dtype is not None else np.dtype(object)

    def __getitem__(self, key):
        if key != () and key is not Ellipsis:
            return self[()][key]
        kind, a, _ = self._layout
        if kind == 'contiguous':
            if a == UNDEF:                 # never written: fill value (zeros)
                return np.zeros(self.shape, self.dtype)
            a += self._r.base
        return _decode(self._r, self._dt, self.shape, a)


def _attributes(r: _Reader, msgs: List[_Message]) -> Dict[str, object]:
    out: Dict[str, object] = {}
    for m in msgs:
        if m.type != 0x0C:
            continue
        version = r.u(m.pos, 1)
        nsize, tsize, ssize = r.u(m.pos + 2, 2), r.u(m.pos + 4, 2), r.u(m.pos + 6, 2)
        p = m.pos + 8
        if version == 1:
            name = r.bytes_(p, nsize)
            p += _pad8(nsize)
            dt, _ = _parse_datatype(r, p)
            p += _pad8(tsize)
            dims, _ = _parse_dataspace(r, p)
            p += _pad8(ssize)
        elif version in (2, 3):
            if version == 3:
                p += 1
            name = r.bytes_(p, nsize)
            p += nsize
            dt, _ = _parse_datatype(r, p)
            p += tsize
            dims, _ = _parse_dataspace(r, p)
            p += ssize
        else:
            raise H5Error('attribute message version %d' % version)
        try:
            out[name.split(b'\x00')[0].decode('utf-8')] = _decode(r, dt, dims, p)
        except H5Error:
            continue                    # an attribute type outside the subset does not make the file unreadable
    return out


class Group:
    def __init__(self, r: _Reader, name: str, addr: int, msgs: Optional[List[_Message]] = None):
        self._r, self.name, self._addr = r, name, addr
        self._msgs = msgs if msgs is not None else _messages(r, addr)
        self._links: Optional[Dict[str, int]] = None
        self.attrs = _attributes(r, self._msgs)

    # -- links
    def _load(self) -> Dict[str, int]:
        if self._links is not None:
            return self._links
        r = self._r
        links: Dict[str, int] = {}
        for m in self._msgs:
            if m.type == 0x11:                                    # symbol table: B-tree + local heap
                btree, heap = r.off(m.pos), r.off(m.pos + r.O)
                ha = heap + r.base
                if r.bytes_(ha, 4) != b'HEAP':
                    raise H5Error('local heap without HEAP signature')
                data = r.off(ha + 8 + 2 * r.L) + r.base
                for name_off, obj in self._btree_entries(btree):
                    end = r.buf.index(b'\x00', data + name_off)
                    links[r.buf[data + name_off:end].decode('utf-8')] = obj
            elif m.type == 0x06:                                  # link message (new-style compact group)
                p = m.pos
                if r.u(p, 1) != 1:
                    raise H5Error('link message version %d' % r.u(p, 1))
                flags = r.u(p + 1, 1)
                p += 2
                ltype = 0
                if flags & 0x08:
                    ltype = r.u(p, 1)
                    p += 1
                if flags & 0x04:
                    p += 8
                if flags & 0x10:
                    p += 1
                lsize = 1 << (flags & 3)
                nlen = r.u(p, lsize)
                p += lsize
                name = r.bytes_(p, nlen).decode('utf-8')
                p += nlen
                if ltype == 0:
                    links[name] = r.off(p)
            elif m.type == 0x02:                                  # link info: dense storage?
                p = m.pos
                flags = r.u(p + 1, 1)
                p += 2 + (8 if flags & 1 else 0)
                if r.off(p) != UNDEF:
                    raise H5Error('%s: dense link storage (fractal heap) is not supported; re-save the file with h5py defaults' % self.name)
        self._links = links
        return links

    def _btree_entries(self, addr: int) -> Iterator[Tuple[int, int]]:
        r = self._r
        a = addr + r.base
        if r.bytes_(a, 4) != b'TREE':
            raise H5Error('B-tree node without TREE signature at %d' % addr)
        if r.u(a + 4, 1) != 0:
            raise H5Error('B-tree node type %d in a group' % r.u(a + 4, 1))
        level, used = r.u(a + 5, 1), r.u(a + 6, 2)
        p = a + 8 + 2 * r.O
        for i in range(used):
            child = r.off(p + r.L + i * (r.L + r.O))
            if level > 0:
                yield from self._btree_entries(child)
                continue
            s = child + r.base
            if r.bytes_(s, 4) != b'SNOD':
                raise H5Error('symbol table node without SNOD signature')
            n = r.u(s + 6, 2)
            esize = 2 * r.O + 8 + 16
            for k in range(n):
                e = s + 8 + k * esize
                yield r.off(e), r.off(e + r.O)

    # -- h5py-like interface
    def keys(self) -> List[str]:
        return sorted(self._load().keys())

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self._load())

    def __contains__(self, path: str) -> bool:
        try:
            self[path]
            return True
        except KeyError:
            return False

    def __getitem__(self, path: str):
        node = self
        parts = [p for p in path.split('/') if p]
        for i, part in enumerate(parts):
            if not isinstance(node, Group):
                raise KeyError(path)
            links = node._load()
            if part not in links:
                raise KeyError(path)
            addr = links[part]
            msgs = _messages(self._r, addr)
            full = (node.name.rstrip('/') + '/' + part)
            is_group = any(m.type in (0x11, 0x02, 0x06, 0x0A) for m in msgs) and not any(m.type == 0x08 for m in msgs)
            node = Group(self._r, full, addr, msgs) if is_group else Dataset(self._r, full, msgs)
        return node


class File(Group):
    """Read-only HDF5 file: `with h5lite.File(path) as f: f['model_weights/aspp0/aspp0/kernel:0'][()]`."""

    def __init__(self, path_or_bytes, mode: str = 'r'):
        if mode != 'r':
            raise H5Error('h5lite is read-only')
        if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
            buf = bytes(path_or_bytes)
        else:
            with open(path_or_bytes, 'rb') as fh:
                buf = fh.read()
        r = _Reader(buf)
        pos = 0
        while True:                      # the superblock may sit at 0, 512, 1024, ... (user block)
            if buf[pos:pos + 8] == SIGNATURE:
                break
            pos = 512 if pos == 0 else pos * 2
            if pos + 8 > len(buf):
                raise H5Error('not an HDF5 file (no superblock signature)')
        version = buf[pos + 8]
        if version in (0, 1):
            r.O, r.L = buf[pos + 13], buf[pos + 14]
            p = pos + 24 + (4 if version == 1 else 0)
            r.base = r.off(p)
            root_entry = p + 4 * r.O
            root = r.off(root_entry + r.O)
        elif version in (2, 3):
            r.O, r.L = buf[pos + 9], buf[pos + 10]
            r.base = r.off(pos + 12)
            root = r.off(pos + 12 + 3 * r.O)
        else:
            raise H5Error('superblock version %d' % version)
        if r.O not in (4, 8) or r.L not in (4, 8):
            raise H5Error('offset / length sizes %d / %d' % (r.O, r.L))
        super().__init__(r, '/', root)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass


def walk_datasets(group: Group, prefix: str = '') -> Iterator[Tuple[str, Dataset]]:
    for key in group.keys():
        item = group[key]
        path = prefix + '/' + key if prefix else key
        if isinstance(item, Group):
            yield from walk_datasets(item, path)
        else:
            yield path, item


KERAS_VARIABLES = ('kernel', 'bias', 'depthwise_kernel', 'gamma', 'beta', 'moving_mean', 'moving_variance')


def keras_weights(path_or_file) -> Dict[str, np.ndarray]:
    """{'<layer>/<variable>': fp32 array} for every Keras variable dataset of a weight file, matched BY NAME: a dataset path ending in
    `<layer>/<variable>:0` under any scope prefix (`model_weights/aspp0/aspp0/kernel:0`, `aspp0/aspp0_1/kernel:0`, ...).  Works for
    model.save() files (weights under /model_weights) and model.save_weights() files (weights at the root)."""
    f = path_or_file if isinstance(path_or_file, Group) else File(path_or_file)
    root = f['model_weights'] if 'model_weights' in f else f
    out: Dict[str, np.ndarray] = {}
    for path, ds in walk_datasets(root):
        parts = path.split('/')
        if len(parts) < 2:
            continue
        var = parts[-1].split(':')[0]
        if var not in KERAS_VARIABLES:
            continue
        layer = parts[0]                 # the top-level group is the Keras layer name; inner scopes may carry '_1' suffixes
        key = '%s/%s' % (layer, var)
        a = np.asarray(ds[()], np.float32)
        if key in out and not np.array_equal(out[key], a):
            raise H5Error('two different datasets map to %s' % key)
        out[key] = a
    return out
