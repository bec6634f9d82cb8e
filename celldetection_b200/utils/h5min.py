"""A minimal HDF5 writer / reader for the result files of ``cpn_inference`` (the reference writes them with
``cd.to_h5`` -> h5py, /root/reference/celldetection/util/util.py:1357-1399; h5py / libhdf5 are not part of this image).

Written from the HDF5 File Format Specification (version 0 superblock -- what libhdf5 itself writes by default -- with
old-style groups: symbol-table message -> v1 B-tree -> symbol-table node + local heap; version 1 object headers;
contiguous data layout, version 3; little-endian fixed-point and IEEE floating-point datatypes; fixed-length string
attributes).  One root group, one flat level of datasets: exactly the layout ``h5py.File(f, 'w').create_dataset(k, data=v)``
produces for the result dict, minus chunking / compression (``to_h5`` passes none by default).

``read`` parses the same subset -- and is validated in the tests on a file written by libhdf5 itself (the MATLAB v7.3 file
shipped with scipy's test data), which pins this module's reading of the specification; ``write`` is validated through
``read``.  When h5py is importable ``utils.outputs.to_h5`` uses it instead.
"""
import struct

import numpy as np

SIGNATURE = b'\x89HDF\r\n\x1a\n'
UNDEF = 0xffffffffffffffff
LEAF_K = 16            # symbol-table nodes hold up to 2 * LEAF_K entries: one node for every result dict
INTERNAL_K = 16

MSG_DATASPACE, MSG_DATATYPE, MSG_FILL, MSG_LAYOUT, MSG_ATTRIBUTE, MSG_CONTINUATION, MSG_SYMBOL_TABLE = (
    0x1, 0x3, 0x5, 0x8, 0xC, 0x10, 0x11)


def _pad8(b):
    return b + b'\0' * (-len(b) % 8)


# ----------------------------------------------------------------------------------------------------------------------
# datatype / dataspace messages
# ----------------------------------------------------------------------------------------------------------------------
def _datatype_message(dt):
    dt = np.dtype(dt)
    if dt.kind in 'iub':
        size = dt.itemsize
        bits = 0x08 if dt.kind == 'i' else 0x00                       # bit 3: two's complement signed
        return struct.pack('<BBBBI', 0x10 | 0, bits, 0, 0, size) + struct.pack('<HH', 0, size * 8)
    if dt.kind == 'f' and dt.itemsize in (4, 8):
        size = dt.itemsize
        exp_bits, man_bits, bias = (8, 23, 127) if size == 4 else (11, 52, 1023)
        # bit field: little endian, mantissa normalisation 2 (implied msb), sign bit position in byte 1
        return (struct.pack('<BBBBI', 0x10 | 1, 0x20, size * 8 - 1, 0, size) +
                struct.pack('<HHBBBBI', 0, size * 8, man_bits, exp_bits, 0, man_bits, bias))
    if dt.kind == 'S':
        return struct.pack('<BBBBI', 0x10 | 3, 0x00, 0, 0, dt.itemsize)   # null-terminated, ASCII
    raise TypeError(f'h5min: unsupported dtype {dt}')


def _parse_datatype(b):
    cls, ver = b[0] & 0x0f, b[0] >> 4
    bits0, bits1 = b[1], b[2]
    size = struct.unpack_from('<I', b, 4)[0]
    if bits0 & 1 and cls in (0, 1):
        raise NotImplementedError('h5min: big-endian data')
    if cls == 0:
        return np.dtype(('<i' if bits0 & 0x08 else '<u') + str(size))
    if cls == 1:
        return np.dtype('<f' + str(size))
    if cls == 3:
        return np.dtype('S' + str(size))
    raise NotImplementedError(f'h5min: datatype class {cls} (version {ver})')


def _dataspace_message(shape):
    if shape == ():
        return struct.pack('<BBB5x', 1, 0, 0)
    return struct.pack('<BBB5x', 1, len(shape), 0) + b''.join(struct.pack('<Q', int(d)) for d in shape)


def _parse_dataspace(b):
    ver, rank, flags = b[0], b[1], b[2]
    off = 8 if ver == 1 else 4
    return tuple(struct.unpack_from('<Q', b, off + 8 * i)[0] for i in range(rank))


def _message(mtype, body, flags=0):
    body = _pad8(body)
    return struct.pack('<HHB3x', mtype, len(body), flags) + body


def _object_header(messages):
    body = b''.join(messages)
    # version 1 prefix: version, reserved, #messages, reference count, header size, 4 bytes of alignment padding
    return struct.pack('<BBHII4x', 1, 0, len(messages), 1, len(body)) + body


def _attribute_message(name, value):
    """Version 1 attribute message; ``value``: str / bytes (fixed-length string, scalar space) or an ndarray."""
    if isinstance(value, str):
        value = value.encode('utf-8')
    if isinstance(value, bytes):
        arr = np.array(value + b'\0', dtype=f'S{len(value) + 1}')
    else:
        arr = np.asarray(value)
        arr = arr if arr.ndim == 0 else np.ascontiguousarray(arr)
    nm = name.encode('utf-8') + b'\0'
    dtm, dsm = _datatype_message(arr.dtype), _dataspace_message(arr.shape)
    body = struct.pack('<BBHHH', 1, 0, len(nm), len(dtm), len(dsm)) + _pad8(nm) + _pad8(dtm) + _pad8(dsm) + arr.tobytes()
    return _message(MSG_ATTRIBUTE, body)


# ----------------------------------------------------------------------------------------------------------------------
# writer
# ----------------------------------------------------------------------------------------------------------------------
def write(filename, datasets, attributes=None):
    """``datasets``: {name: array-like}; ``attributes``: {dataset name: {attribute name: str | bytes | ndarray}}."""
    attributes = attributes or {}
    names = sorted(datasets, key=lambda s: s.encode('utf-8'))          # symbol-table entries are ordered by name
    if len(names) > 2 * LEAF_K:
        raise ValueError(f'h5min: at most {2 * LEAF_K} datasets per file')
    arrays = {}
    for k in names:
        a = np.asarray(datasets[k])
        if a.dtype == np.bool_:
            a = a.astype(np.uint8)                                     # (h5py stores numpy bool as an enum; kept plain here)
        if a.dtype.byteorder == '>':
            a = a.astype(a.dtype.newbyteorder('<'))
        arrays[k] = a if a.ndim == 0 else np.ascontiguousarray(a)      # (ascontiguousarray turns 0-d into 1-d)

    # local heap data segment: the empty string at offset 0, then the link names, 8-byte aligned, then one free block
    heap, name_off = bytearray(8), {}
    for k in names:
        name_off[k] = len(heap)
        heap += _pad8(k.encode('utf-8') + b'\0')
    free_off = len(heap)
    heap += struct.pack('<QQ', 1, 16)                                  # free block: next = 1 (end of list), size 16

    # file layout: superblock, root object header, B-tree node, heap header, heap data, symbol-table node, then per
    # dataset its object header followed by its (8-byte aligned) raw data
    sb_size = 8 + 8 + 4 + 4 + 4 * 8 + 40
    root_hdr_addr = sb_size
    root_hdr_size = 16 + 8 + 16
    btree_addr = root_hdr_addr + root_hdr_size
    btree_size = 24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8
    heap_hdr_addr = btree_addr + btree_size
    heap_data_addr = heap_hdr_addr + 32
    snod_addr = heap_data_addr + len(heap)
    snod_size = 8 + 2 * LEAF_K * 40
    cursor = snod_addr + snod_size

    headers, placed = {}, {}
    for k in names:
        a = arrays[k]
        hdr_addr = cursor
        msgs = [_message(MSG_DATASPACE, _dataspace_message(a.shape)),
                _message(MSG_DATATYPE, _datatype_message(a.dtype), flags=1),
                # late allocation, written if set, default value: the bytes libhdf5 itself stores for a plain dataset
                _message(MSG_FILL, struct.pack('<BBBBI', 1, 2, 2, 1, 0), flags=1),
                None]
        for an, av in (attributes.get(k) or {}).items():
            msgs.append(_attribute_message(an, av))
        size = 16 + sum(len(m) for m in msgs if m is not None) + 8 + 24   # + layout message (8 header + 24 body)
        data_addr = (hdr_addr + size + 7) // 8 * 8
        layout = struct.pack('<BBQQ', 3, 1, data_addr if a.nbytes else UNDEF, a.nbytes)
        msgs[3] = _message(MSG_LAYOUT, layout)
        hdr = _object_header(msgs)
        assert len(hdr) == size, (len(hdr), size)
        headers[k] = hdr
        placed[k] = (hdr_addr, data_addr)
        cursor = (data_addr + a.nbytes + 7) // 8 * 8
    eof = cursor

    out = bytearray(eof)

    def put(addr, b):
        out[addr:addr + len(b)] = b

    # superblock, version 0
    sb = SIGNATURE + struct.pack('<BBBBBBBB', 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack('<HHI', LEAF_K, INTERNAL_K, 0)
    sb += struct.pack('<QQQQ', 0, UNDEF, eof, UNDEF)
    sb += struct.pack('<QQII', 0, root_hdr_addr, 1, 0) + struct.pack('<QQ', btree_addr, heap_hdr_addr)   # root entry
    assert len(sb) == sb_size
    put(0, sb)
    put(root_hdr_addr, _object_header([_message(MSG_SYMBOL_TABLE, struct.pack('<QQ', btree_addr, heap_hdr_addr), flags=1)]))
    # B-tree: one leaf-level node with one child (the symbol-table node); key 0 = "", key 1 = the largest name
    bt = b'TREE' + struct.pack('<BBH', 0, 0, 1 if names else 0) + struct.pack('<QQ', UNDEF, UNDEF)
    if names:
        bt += struct.pack('<QQQ', 0, snod_addr, name_off[names[-1]])
    put(btree_addr, bt)
    put(heap_hdr_addr, b'HEAP' + struct.pack('<B3x', 0) + struct.pack('<QQQ', len(heap), free_off, heap_data_addr))
    put(heap_data_addr, bytes(heap))
    sn = b'SNOD' + struct.pack('<BBH', 1, 0, len(names))
    for k in names:
        sn += struct.pack('<QQII16x', name_off[k], placed[k][0], 0, 0)
    put(snod_addr, sn)
    for k in names:
        put(placed[k][0], headers[k])
        put(placed[k][1], arrays[k].tobytes())
    with open(filename, 'wb') as f:
        f.write(bytes(out))
    return filename


# ----------------------------------------------------------------------------------------------------------------------
# reader (same subset; also walks files written by libhdf5 as long as they stay inside it)
# ----------------------------------------------------------------------------------------------------------------------
class _File:
    def __init__(self, buf):
        self.buf = buf
        base = 0
        while buf[base:base + 8] != SIGNATURE:                         # a user block shifts the superblock to 512, 1024, ...
            base = 512 if base == 0 else base * 2
            if base >= len(buf):
                raise ValueError('h5min: not an HDF5 file')
        self.base = base
        ver = buf[base + 8]
        if ver not in (0, 1):
            raise NotImplementedError(f'h5min: superblock version {ver}')
        assert buf[base + 13] == 8 and buf[base + 14] == 8, 'h5min: 8-byte offsets / lengths only'
        o = base + 16
        self.leaf_k, self.internal_k = struct.unpack_from('<HH', buf, o)
        o += 8 + (4 if ver == 1 else 0)
        self.base_addr, _, self.eof, _ = struct.unpack_from('<QQQQ', buf, o)
        o += 32
        _, self.root_header, cache, _ = struct.unpack_from('<QQII', buf, o)
        self.base = self.base_addr                                      # every address in the file is relative to it

    def at(self, addr):
        return self.base + addr

    def messages(self, addr):
        """(type, flags, body) of every message of a version 1 object header, continuation blocks included."""
        buf, o = self.buf, self.at(addr)
        ver, _, nmsg, _, size = struct.unpack_from('<BBHII', buf, o)
        if ver != 1:
            raise NotImplementedError(f'h5min: object header version {ver}')
        blocks, out = [(o + 16, size)], []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            while left >= 8 and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from('<HHB', buf, p)
                body = bytes(buf[p + 8:p + 8 + msize])
                if mtype == MSG_CONTINUATION:
                    caddr, clen = struct.unpack_from('<QQ', body, 0)
                    blocks.append((self.at(caddr), clen))
                out.append((mtype, flags, body))
                p += 8 + msize
                left -= 8 + msize
        return out

    def heap_name(self, heap_addr, off):
        o = self.at(heap_addr)
        assert self.buf[o:o + 4] == b'HEAP'
        _, _, data_addr = struct.unpack_from('<QQQ', self.buf, o + 8)
        s = self.at(data_addr) + off
        e = self.buf.index(b'\0', s)
        return bytes(self.buf[s:e]).decode('utf-8')

    def group_links(self, btree_addr, heap_addr):
        """{name: object header address} of an old-style group."""
        links, stack = {}, [btree_addr]
        while stack:
            o = self.at(stack.pop())
            sig = bytes(self.buf[o:o + 4])
            if sig == b'TREE':
                ntype, level, used = struct.unpack_from('<BBH', self.buf, o + 4)
                assert ntype == 0
                for i in range(used):
                    stack.append(struct.unpack_from('<Q', self.buf, o + 24 + 8 + 16 * i)[0])
            elif sig == b'SNOD':
                n = struct.unpack_from('<H', self.buf, o + 6)[0]
                for i in range(n):
                    name_off, hdr = struct.unpack_from('<QQ', self.buf, o + 8 + 40 * i)
                    links[self.heap_name(heap_addr, name_off)] = hdr
            else:
                raise ValueError(f'h5min: unexpected group node {sig!r}')
        return links

    def attribute(self, body):
        ver, _, nlen, dlen, slen = struct.unpack_from('<BBHHH', body, 0)
        pad = (lambda n: (n + 7) // 8 * 8) if ver == 1 else (lambda n: n)
        o = 8
        name = body[o:o + nlen].split(b'\0')[0].decode('utf-8')
        o += pad(nlen)
        dt = _parse_datatype(body[o:o + dlen])
        o += pad(dlen)
        shape = _parse_dataspace(body[o:o + slen])
        o += pad(slen)
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        val = np.frombuffer(body, dtype=dt, count=n, offset=o).reshape(shape)
        if dt.kind == 'S' and shape == ():
            return name, bytes(val[()]).split(b'\0')[0].decode('utf-8')
        return name, val.copy()

    def node(self, addr):
        """('group', {name: addr}) or ('dataset', array | None, {attributes})."""
        msgs = self.messages(addr)
        for mtype, _, body in msgs:
            if mtype == MSG_SYMBOL_TABLE:
                return ('group', self.group_links(*struct.unpack_from('<QQ', body, 0)))
        shape = dt = data = None
        attrs = {}
        layout = None
        for mtype, _, body in msgs:
            if mtype == MSG_DATASPACE:
                shape = _parse_dataspace(body)
            elif mtype == MSG_DATATYPE:
                try:
                    dt = _parse_datatype(body)
                except NotImplementedError:
                    dt = None
            elif mtype == MSG_LAYOUT:
                layout = body
            elif mtype == MSG_ATTRIBUTE:
                try:
                    k, v = self.attribute(body)
                    attrs[k] = v
                except NotImplementedError:
                    pass
        if layout is not None and dt is not None and shape is not None:
            n = int(np.prod(shape, dtype=np.int64)) if shape else 1
            if layout[0] == 3 and layout[1] == 1:                      # version 3, contiguous
                daddr, dsize = struct.unpack_from('<QQ', layout, 2)
                data = (np.zeros(shape, dt) if daddr == UNDEF or n == 0 else
                        np.frombuffer(self.buf, dtype=dt, count=n, offset=self.at(daddr)).reshape(shape).copy())
            elif layout[0] == 3 and layout[1] == 0:                    # version 3, compact
                data = np.frombuffer(layout, dtype=dt, count=n, offset=4).reshape(shape).copy()
            elif layout[0] in (1, 2) and layout[2] == 1:               # versions 1 / 2 (older libhdf5), contiguous
                daddr = struct.unpack_from('<Q', layout, 8)[0]
                data = (np.zeros(shape, dt) if daddr == UNDEF or n == 0 else
                        np.frombuffer(self.buf, dtype=dt, count=n, offset=self.at(daddr)).reshape(shape).copy())
        return ('dataset', data, attrs)


def read(filename, with_attributes=False):
    """{dataset name: ndarray} of the root group (nested groups as nested dicts; datasets outside the subset -> None)."""
    with open(filename, 'rb') as f:
        h = _File(f.read())

    def walk(addr):
        kind, *rest = h.node(addr)
        if kind == 'group':
            return {k: walk(a) for k, a in rest[0].items()}
        return (rest[0], rest[1]) if with_attributes else rest[0]

    return walk(h.root_header)
