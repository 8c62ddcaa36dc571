"""heimdall_core.detection -- rust/heimdall-core/src/detection.rs:36-124 (find_contours)."""
from __future__ import annotations

from typing import List, Optional

import numpy as np


def _image3(image) -> np.ndarray:
    if not isinstance(image, np.ndarray) or image.dtype != np.uint8 or image.ndim != 3:
        raise TypeError("argument 'image': expected a 3-dimensional numpy array of uint8")
    return np.ascontiguousarray(image)


def _dfs_pop_order(pixels: set, start):
    """Order in which the reference's explicit-stack flood fill pops the pixels of one small blob
    (detection.rs:66-88: push order up, down, left, right; LIFO pop).  Pure result marshalling for blobs of at most
    100 pixels -- the labelling itself comes from the GPU."""
    order = []
    visited = {start}
    stack = [start]
    while stack:
        y, x = stack.pop()
        order.append((y, x))
        for q in ((y - 1, x), (y + 1, x), (y, x - 1), (y, x + 1)):
            if q in pixels and q not in visited:
                visited.add(q)
                stack.append(q)
    return order


def find_contours(image, min_area: Optional[float] = None, max_area: Optional[float] = None) -> List[dict]:
    """Blobs (4-connected, foreground `> 127`) with min_area <= area <= max_area (defaults 10, 10000) in discovery
    order: {"position": (cy, cx), "area": float, "pixel_count": int[, "points": [(y, x), ...] if <= 100 pixels]}."""
    from .batch import with_capacity_retry
    img = _image3(image)
    lo = 10.0 if min_area is None else float(min_area)
    hi = 10000.0 if max_area is None else float(max_area)
    recs, labels = with_capacity_retry(lambda d: d.find_contours(img, lo, hi, want_labels=True), *img.shape[:2])
    small = {int(r.label) for r in recs if r.pixel_count <= 100}
    by_label = {}
    if small:
        ys, xs = np.nonzero(labels)
        labs = labels[ys, xs]
        sel = np.isin(labs, np.fromiter(small, dtype=np.int32))
        for y, x, l in zip(ys[sel].tolist(), xs[sel].tolist(), labs[sel].tolist()):
            by_label.setdefault(l, []).append((y, x))
    out = []
    for r in recs:
        d = {"position": (int(r.y), int(r.x)), "area": float(r.area), "pixel_count": int(r.pixel_count)}
        if r.pixel_count <= 100:
            pts = by_label[int(r.label)]  # raster order; pts[0] is the seed
            d["points"] = _dfs_pop_order(set(pts), pts[0])
        out.append(d)
    return out
