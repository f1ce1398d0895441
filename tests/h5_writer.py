"""Test infrastructure: a minimal HDF5 WRITER that lays a tree of float / integer arrays out the way libhdf5 does with h5py's default
settings (what Keras' model.save_weights produces): version-0 superblock, old-style groups (symbol-table message -> version-1 B-tree of
symbol nodes, leaf K = 4, internal K = 16, names in a local heap), version-1 object headers, version-1 dataspace / attribute messages,
data-layout message version 3 with CONTIGUOUS raw data, fixed-length string array attributes (layer_names / weight_names) and
variable-length string attributes through a global heap collection (keras_version / backend).  Written from the HDF5 File Format
Specification; h5py / libhdf5 are not in this image, so dlv3p_b200.h5lite is exercised on this layout.

    write_h5(path, {'model_weights': {'aspp0': {'aspp0': {'kernel:0': array, ...}}}}, attrs={'/': {'backend': 'tensorflow'}, ...})
"""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 4, 16


def _pad8(b: bytes) -> bytes:
    return b + b'\x00' * (-len(b) % 8)


class _Writer:
    def __init__(self):
        self.buf = bytearray(96)            # superblock placeholder

    def alloc(self, data: bytes) -> int:
        self.buf += b'\x00' * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    def reserve(self, n: int) -> int:
        return self.alloc(b'\x00' * n)

    def put(self, addr: int, data: bytes) -> None:
        self.buf[addr:addr + len(data)] = data


def _datatype(a: np.ndarray) -> bytes:
    dt = a.dtype
    if dt.kind == 'f':
        size = dt.itemsize
        exp_bits, man_bits, bias = {2: (5, 10, 15), 4: (8, 23, 127), 8: (11, 52, 1023)}[size]
        big = dt.byteorder == '>'
        head = bytes([0x11, 0x20 | (1 if big else 0), size * 8 - 1, 0]) + struct.pack('<I', size)
        return head + struct.pack('<HHBBBBI', 0, size * 8, man_bits, exp_bits, 0, man_bits, bias)
    if dt.kind in 'iu':
        big = dt.byteorder == '>'
        bits = (1 if big else 0) | (0x08 if dt.kind == 'i' else 0)
        return bytes([0x10, bits, 0, 0]) + struct.pack('<I', dt.itemsize) + struct.pack('<HH', 0, dt.itemsize * 8)
    if dt.kind == 'S':
        return bytes([0x13, 0, 0, 0]) + struct.pack('<I', dt.itemsize)
    raise TypeError(dt)


def _dataspace(shape: Tuple[int, ...]) -> bytes:
    return bytes([1, len(shape), 0, 0, 0, 0, 0, 0]) + b''.join(struct.pack('<Q', d) for d in shape)


def _message(mtype: int, data: bytes, flags: int = 0) -> bytes:
    data = _pad8(data)
    return struct.pack('<HHB3x', mtype, len(data), flags) + data


def _object_header(msgs: List[bytes]) -> bytes:
    body = b''.join(msgs)
    return struct.pack('<BBHII4x', 1, 0, len(msgs), 1, len(body)) + body


def _attribute(w: _Writer, name: str, value) -> bytes:
    nm = name.encode() + b'\x00'
    if isinstance(value, str):              # variable-length UTF-8 string through a global heap collection
        s = value.encode('utf-8')
        obj = struct.pack('<HH4xQ', 1, 1, len(s)) + _pad8(s)
        size = 4096
        coll = b'GCOL' + bytes([1, 0, 0, 0]) + struct.pack('<Q', size) + obj
        coll += struct.pack('<HH4xQ', 0, 0, size - len(coll) - 16)
        heap = w.alloc(coll + b'\x00' * (size - len(coll)))
        dtype = bytes([0x19, 0x01, 0x01, 0x00]) + struct.pack('<I', 16) + bytes([0x13, 0, 0, 0]) + struct.pack('<I', 1)
        space = _dataspace(())
        data = struct.pack('<IQI', len(s), heap, 1)
    else:
        a = np.asarray(value)
        dtype, space, data = _datatype(a), _dataspace(a.shape), a.tobytes()
    return _message(0x0C, bytes([1, 0]) + struct.pack('<HHH', len(nm), len(dtype), len(space)) + _pad8(nm) + _pad8(dtype) + _pad8(space) + data)


def _write_dataset(w: _Writer, a: np.ndarray, chunked: bool = False) -> int:
    a = np.asarray(a)
    a = a.copy(order='C') if not a.flags.c_contiguous else a
    raw = w.alloc(a.tobytes()) if a.size else UNDEF
    if chunked:                              # only to check that the reader refuses it: layout class 2 with a dangling B-tree address
        layout = bytes([3, 2, len(a.shape) + 1]) + struct.pack('<Q', UNDEF) + b''.join(struct.pack('<I', d) for d in a.shape) + struct.pack('<I', a.dtype.itemsize)
    else:
        layout = bytes([3, 1]) + struct.pack('<QQ', raw, a.nbytes)
    fill = bytes([2, 2, 0, 0])              # fill value message version 2: late allocation, never write, undefined
    msgs = [_message(0x01, _dataspace(a.shape)), _message(0x03, _datatype(a), 1), _message(0x05, fill), _message(0x08, layout)]
    return w.alloc(_object_header(msgs))


def _write_group(w: _Writer, tree: Dict, attrs: Dict[str, Dict], path: str, chunked: Tuple[str, ...]) -> int:
    names = sorted(tree.keys())
    # children first (addresses are needed for the symbol table entries)
    entries = []
    for n in names:
        child_path = path.rstrip('/') + '/' + n
        v = tree[n]
        if isinstance(v, dict):
            entries.append((n, _write_group(w, v, attrs, child_path, chunked), v))
        else:
            entries.append((n, _write_dataset(w, np.asarray(v), child_path in chunked), None))
    # local heap: offset 0 = the empty string, then the names
    heap_data = bytearray(8)
    name_off = {}
    for n in names:
        name_off[n] = len(heap_data)
        heap_data += _pad8(n.encode() + b'\x00')
    free_off = len(heap_data)
    heap_data += struct.pack('<QQ', 1, 16 + 0)     # one free block: next = 1 (none), size
    heap_data += b'\x00' * 16
    struct.pack_into('<QQ', heap_data, free_off, 1, len(heap_data) - free_off)
    data_addr = w.alloc(bytes(heap_data))
    heap_addr = w.alloc(b'HEAP' + bytes([0, 0, 0, 0]) + struct.pack('<QQQ', len(heap_data), free_off, data_addr))
    # symbol nodes of up to 2 * LEAF_K entries
    leaves: List[Tuple[int, int]] = []             # (address, heap offset of the node's largest name)
    for i in range(0, max(len(entries), 1), 2 * LEAF_K):
        part = entries[i:i + 2 * LEAF_K]
        body = b'SNOD' + bytes([1, 0]) + struct.pack('<H', len(part))
        for n, addr, sub in part:
            if sub is not None:                    # cached symbol-table info of a group: B-tree and heap addresses (filled by the child)
                bt, hp = _GROUP_CACHE[addr]
                body += struct.pack('<QQII', name_off[n], addr, 1, 0) + struct.pack('<QQ', bt, hp)
            else:
                body += struct.pack('<QQII', name_off[n], addr, 0, 0) + b'\x00' * 16
        body += b'\x00' * ((2 * LEAF_K - len(part)) * 40)
        leaves.append((w.alloc(body), name_off[part[-1][0]] if part else 0))
    # B-tree levels of up to 2 * INTERNAL_K children
    level, nodes = 0, leaves
    while True:
        parents: List[Tuple[int, int]] = []
        for i in range(0, len(nodes), 2 * INTERNAL_K):
            part = nodes[i:i + 2 * INTERNAL_K]
            body = b'TREE' + bytes([0, level]) + struct.pack('<H', len(part)) + struct.pack('<QQ', UNDEF, UNDEF)
            body += struct.pack('<Q', 0)                                     # key 0: the empty string
            for addr, key in part:
                body += struct.pack('<QQ', addr, key)
            body += b'\x00' * ((2 * INTERNAL_K - len(part)) * 16)
            parents.append((w.alloc(body), part[-1][1]))
        # sibling pointers
        for k, (addr, _) in enumerate(parents):
            left = parents[k - 1][0] if k > 0 else UNDEF
            right = parents[k + 1][0] if k + 1 < len(parents) else UNDEF
            w.put(addr + 8, struct.pack('<QQ', left, right))
        if len(parents) == 1:
            btree = parents[0][0]
            break
        nodes, level = parents, level + 1
    msgs = [_message(0x11, struct.pack('<QQ', btree, heap_addr))]
    for k, v in attrs.get(path, {}).items():
        msgs.append(_attribute(w, k, v))
    addr = w.alloc(_object_header(msgs))
    _GROUP_CACHE[addr] = (btree, heap_addr)
    return addr


_GROUP_CACHE: Dict[int, Tuple[int, int]] = {}


def write_h5(path, tree: Dict, attrs: Dict[str, Dict] = None, chunked: Tuple[str, ...] = ()) -> bytes:
    """Write `tree` ({name: subtree | array}) to `path` (or only return the bytes when path is None)."""
    _GROUP_CACHE.clear()
    w = _Writer()
    root = _write_group(w, tree, attrs or {}, '/', tuple(chunked))
    bt, hp = _GROUP_CACHE[root]
    sb = b'\x89HDF\r\n\x1a\n' + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack('<HHI', LEAF_K, INTERNAL_K, 0)
    sb += struct.pack('<QQQQ', 0, UNDEF, len(w.buf), UNDEF)
    sb += struct.pack('<QQII', 0, root, 1, 0) + struct.pack('<QQ', bt, hp)
    assert len(sb) == 96
    w.put(0, sb)
    data = bytes(w.buf)
    if path is not None:
        with open(path, 'wb') as f:
            f.write(data)
    return data
