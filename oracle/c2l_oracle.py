"""TEST INFRASTRUCTURE ONLY -- CPU restatement of ``cd.data.contours2labels``, the label rasterisation
that follows the CPN hot path (/root/reference/celldetection/data/cpn.py:246-256 ``render_contour``, :292-358
``contours2labels``; called by celldetection_scripts/cpn_inference.py:809-813).

The polygon fill itself lives in a third-party dependency that is absent from /root/reference: OpenCV
(``opencv-python``, unpinned in the reference's requirements; 4.13.0 in the build container), ``cv2.drawContours(a,
[pts], 0, val, thickness=-1, offset)``.  Its published algorithm (modules/imgproc/src/drawing.cpp: ``CollectPolyEdges``
-> ``Line`` (8-connected Bresenham, left-to-right ``LineIterator``) for every polygon edge, then ``FillEdgeCollection``,
an integer scan-line fill on 16.16 fixed-point edge x-coordinates) is restated here in plain Python/numpy integers and
pinned against the installed cv2 (tests/test_oracle_golden.py, tests/golden/contours2labels.npz minted from the
reference's own ``contours2labels``).
"""
import numpy as np

XY_SHIFT = 16
XY_ONE = 1 << XY_SHIFT


def bresenham8(p0, p1):
    """Pixels of cv::Line / LineIterator(pt1, pt2, connectivity=8, leftToRight=true), endpoints inclusive."""
    (x0, y0), (x1, y1) = (int(p0[0]), int(p0[1])), (int(p1[0]), int(p1[1]))
    dx, dy = x1 - x0, y1 - y0
    if dx < 0:                       # leftToRight: start from the left end point
        x0, y0, x1, y1 = x1, y1, x0, y0
        dx, dy = -dx, -dy
    sx, sy = 1, (1 if dy >= 0 else -1)
    dy = abs(dy)
    steep = dy > dx
    if steep:
        dx, dy = dy, dx
    err = dx - (dy + dy)
    plus_delta, minus_delta = dx + dx, -(dy + dy)
    out = []
    x, y = x0, y0
    for _ in range(dx + 1):
        out.append((x, y))
        neg = err < 0
        err += minus_delta + (plus_delta if neg else 0)
        if steep:                    # major axis y
            y += sy
            if neg:
                x += sx
        else:
            x += sx
            if neg:
                y += sy
    return out


def fill_polygon(pts, width, height):
    """Boolean mask [height, width] of cv2.drawContours(zeros, [pts], 0, 1, thickness=-1) for integer vertices ``pts``
    [n, 2] (x, y) -- boundary lines plus scan-line interior (drawing.cpp CollectPolyEdges + FillEdgeCollection, shift 0,
    line_type 8).  Vertices must lie inside the image (render_contour guarantees it: the canvas is the bounding box)."""
    pts = np.asarray(pts, dtype=np.int64).reshape(-1, 2)
    mask = np.zeros((height, width), dtype=bool)
    n = len(pts)
    if n == 0:
        return mask
    edges = []                       # (y0, y1, x (16.16 at y0), dx per scan line)
    p0 = pts[-1]
    for i in range(n):
        p1 = pts[i]
        for (x, y) in bresenham8(p0, p1):
            if 0 <= x < width and 0 <= y < height:
                mask[y, x] = True
        if p0[1] != p1[1]:
            x0f, x1f = int(p0[0]) << XY_SHIFT, int(p1[0]) << XY_SHIFT
            num, den = x1f - x0f, int(p1[1]) - int(p0[1])
            dxl = abs(num) // abs(den) * (1 if (num >= 0) == (den >= 0) else -1)   # C++ integer division (truncation)
            if p0[1] < p1[1]:
                edges.append([int(p0[1]), int(p1[1]), x0f, dxl])
            else:
                edges.append([int(p1[1]), int(p0[1]), x1f, dxl])
        p0 = p1
    if len(edges) < 2:
        return mask
    y_min, y_max = min(e[0] for e in edges), max(e[1] for e in edges)
    for y in range(y_min, y_max):
        xs = sorted(e[2] + e[3] * (y - e[0]) for e in edges if e[0] <= y < e[1])
        for a, b in zip(xs[0::2], xs[1::2]):
            xa, xb = (a + XY_ONE - 1) >> XY_SHIFT, b >> XY_SHIFT
            if 0 <= y < height and xa < width and xb >= 0:
                mask[y, max(xa, 0):min(xb, width - 1) + 1] = True
    return mask


def render_contour(contour, val=1, dtype='int32'):
    """data/cpn.py:246-256 (reference=None, round=False, thickness=-1): canvas = bounding box of the contour."""
    contour = np.asarray(contour)
    xmin, ymin = np.floor(np.min(contour, axis=0)).astype('int')
    xmax, ymax = np.ceil(np.max(contour, axis=0)).astype('int')
    pts = np.array(contour, dtype=np.int32).reshape(-1, 2).astype(np.int64) - np.array([xmin, ymin])
    a = fill_polygon(pts, xmax - xmin + 1, ymax - ymin + 1).astype(dtype) * val
    return a, (xmin, xmax), (ymin, ymax)


def contours2labels(contours, size, rounded=True, clip=True, initial_depth=1, gap=3, dtype='int32'):
    """data/cpn.py:292-358 with ioa_thresh=None, sort_by=None: contour k gets label k + 1 in the first channel whose
    gap-dilated bounding-box region is still empty; channels are appended on demand."""
    labels = np.zeros(tuple(size) + (initial_depth,), dtype=dtype)
    lbl = 1
    for contour in contours:
        contour = np.array(contour, dtype=np.float32 if np.asarray(contour).dtype.kind == 'f' else None, copy=True)
        if rounded:
            contour = np.round(contour)
        if clip:
            np.clip(contour[..., 0], 0, size[1] - 1, out=contour[..., 0])
            np.clip(contour[..., 1], 0, size[0] - 1, out=contour[..., 1])
        a, (xmin, xmax), (ymin, ymax) = render_contour(contour, val=lbl, dtype=dtype)
        lbl += 1
        s = (labels[max(0, ymin - gap): gap + ymin + a.shape[0], max(0, xmin - gap): gap + xmin + a.shape[1]] > 0).sum((0, 1))
        i = next(i for i in range(labels.shape[2] + 1) if not (i < labels.shape[2] and np.any(s[i])))
        if i >= labels.shape[2]:
            labels = np.concatenate((labels, np.zeros(size, dtype=dtype)[..., None]), axis=-1)
        labels[ymin:ymin + a.shape[0], xmin:xmin + a.shape[1], i] += a
    return labels


def resolve_label_channels(labels, max_iter=999):
    """data/cpn.py:361-398 (method='dilation', kernel = cv2.getStructuringElement(MORPH_CROSS, (3, 3))): cv2.dilate is
    restated as the maximum over the pixel and its 4 neighbours (constant border = lowest value)."""
    labels = np.asarray(labels)
    mask_sm = np.sum(labels > 0, axis=-1)
    mask = mask_sm > 1
    if not mask.any():
        return labels.max(-1).astype(labels.dtype)
    lbl = np.zeros(labels.shape[:2], dtype='float64')
    core = mask_sm == 1
    lbl[core] = labels.max(-1)[core]
    for _ in range(max_iter):
        m = mask & (lbl <= 0)
        if not np.any(m):
            break
        p = np.pad(lbl, 1, constant_values=-np.inf)
        dil = np.maximum.reduce([p[1:-1, 1:-1], p[:-2, 1:-1], p[2:, 1:-1], p[1:-1, :-2], p[1:-1, 2:]])
        new = lbl.copy()
        new[m] = dil[m]
        if np.allclose(new, lbl):
            lbl = new
            break
        lbl = new
    return lbl.astype(labels.dtype)
