"""A minimal TIFF 6.0 / BigTIFF writer for the overlay images of ``cpn_inference`` (the reference calls
``tifffile.imwrite(dst, label_vis, compression='ZLIB', bigtiff=label_vis.size > 2 ** 28)``,
/root/reference/celldetection_scripts/cpn_inference.py:846; tifffile is not part of this image).

Strips of ~1 MiB, Adobe-deflate (compression tag 8, what tifffile's 'ZLIB' means) or uncompressed; uint8 / uint16 / float32
samples; gray, RGB or RGBA (unassociated alpha); little endian.  The tests read the files back with the installed OpenCV
(libtiff), classic and BigTIFF.
"""
import struct
import zlib

import numpy as np

_TYPES = {'H': 3, 'I': 4, 'Q': 16}        # SHORT, LONG, LONG8


def imwrite(filename, image, compression='ZLIB', bigtiff=None, level=6):
    a = np.ascontiguousarray(image)
    if a.ndim == 2:
        a = a[..., None]
    if a.ndim != 3 or a.shape[-1] not in (1, 3, 4):
        raise ValueError(f'tiffmin: Array[h, w(, 1|3|4)] expected, got {a.shape}')
    if a.dtype == np.bool_:
        a = a.astype(np.uint8) * 255
    fmt = {'u': 1, 'i': 2, 'f': 3}.get(a.dtype.kind)
    if fmt is None or a.dtype.itemsize not in (1, 2, 4) or (fmt == 3 and a.dtype.itemsize != 4):
        raise TypeError(f'tiffmin: unsupported dtype {a.dtype}')
    a = a.astype(a.dtype.newbyteorder('<'), copy=False)
    h, w, c = a.shape
    if bigtiff is None:
        bigtiff = a.size > 2 ** 28
    deflate = compression is not None and str(compression).upper() in ('ZLIB', 'DEFLATE', 'ADOBE_DEFLATE')
    if compression is not None and not deflate and str(compression).upper() != 'NONE':
        raise ValueError(f'tiffmin: unknown compression {compression!r}')
    row_bytes = w * c * a.dtype.itemsize
    rows = max(1, min(h, (1 << 20) // max(row_bytes, 1))) if h else 1
    strips = []
    for y in range(0, h, rows):
        raw = a[y:y + rows].tobytes()
        strips.append(zlib.compress(raw, level) if deflate else raw)

    off_t = 'Q' if bigtiff else 'I'
    head = (struct.pack('<2sHHHQ', b'II', 43, 8, 0, 0) if bigtiff else struct.pack('<2sHI', b'II', 42, 0))
    offsets, pos = [], len(head)
    for s in strips:
        offsets.append(pos)
        pos += len(s) + (len(s) & 1)                           # word alignment
    if not bigtiff and pos >= 2 ** 32 - 65536:
        raise ValueError('tiffmin: image too large for classic TIFF, pass bigtiff=True')

    tags = [(256, 'I', [w]), (257, 'I', [h]), (258, 'H', [a.dtype.itemsize * 8] * c), (259, 'H', [8 if deflate else 1]),
            (262, 'H', [2 if c >= 3 else 1]), (273, off_t, offsets), (277, 'H', [c]), (278, 'I', [rows]),
            (279, off_t, [len(s) for s in strips]), (284, 'H', [1])]
    if c == 4:
        tags.append((338, 'H', [2]))                           # ExtraSamples: unassociated alpha
    tags.append((339, 'H', [fmt] * c))
    tags.sort(key=lambda t: t[0])

    # IFD after the strips; values that do not fit the entry's value field go behind the IFD
    ifd_pos = pos
    n = len(tags)
    entry_size, value_room = (20, 8) if bigtiff else (12, 4)
    ifd_size = (8 if bigtiff else 2) + n * entry_size + (8 if bigtiff else 4)
    extra_pos, extra, entries = ifd_pos + ifd_size, b'', b''
    for tag, code, vals in tags:
        data = struct.pack('<' + code * len(vals), *vals)
        cnt = struct.pack('<Q' if bigtiff else '<I', len(vals))
        if len(data) <= value_room:
            field = data + b'\0' * (value_room - len(data))
        else:
            field = struct.pack('<Q' if bigtiff else '<I', extra_pos + len(extra))
            extra += data + b'\0' * (len(data) & 1)
        entries += struct.pack('<HH', tag, _TYPES[code]) + cnt + field
    ifd = (struct.pack('<Q', n) if bigtiff else struct.pack('<H', n)) + entries + (b'\0' * 8 if bigtiff else b'\0' * 4)
    head = (struct.pack('<2sHHHQ', b'II', 43, 8, 0, ifd_pos) if bigtiff else struct.pack('<2sHI', b'II', 42, ifd_pos))
    with open(filename, 'wb') as f:
        f.write(head)
        for s in strips:
            f.write(s)
            if len(s) & 1:
                f.write(b'\0')
        f.write(ifd)
        f.write(extra)
    return filename
