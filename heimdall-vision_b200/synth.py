"""Synthetic bottle frames for tests and benchmarks.

Pattern of the reference's simulators (heimdall/core/acquisition.py:313-361, rust/heimdall-core/src/acquisition.rs:57-107)
scaled to (H, W) and emitted as 1-channel u8 (the camera default is 1280x1024 Mono8,
rust/heimdall-camera/src/lib.rs:80-93): background 220, 2-px bottle outline of 100, filled disc of 80, contaminants =
dark discs (value in [0,60), r in [15,30), heimdall/test_contamination.py:36-49) and specks (value 40, r in [3,10),
acquisition.py:345-351), plus uniform integer noise in [-noise, noise].  RNG: numpy default_rng(1234 + frame_index).
Pure numpy so the same bytes come out everywhere.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

BASE_SEED = 1234


def _disc(img: np.ndarray, cx: int, cy: int, r: int, value: int) -> None:
    h, w = img.shape
    y0, y1, x0, x1 = max(cy - r, 0), min(cy + r + 1, h), max(cx - r, 0), min(cx + r + 1, w)
    if y0 >= y1 or x0 >= x1:
        return
    yy, xx = np.mgrid[y0:y1, x0:x1]
    sel = (xx - cx) ** 2 + (yy - cy) ** 2 <= r * r
    img[y0:y1, x0:x1][sel] = value


def bottle_frame(h: int = 1024, w: int = 1280, index: int = 0, *, contaminants: Optional[int] = None,
                 p_defect: float = 0.3, noise: int = 3) -> np.ndarray:
    """One (h, w) u8 frame.  contaminants=None draws 0 (prob 1-p_defect) or 1..3 contaminants."""
    rng = np.random.default_rng(BASE_SEED + index)
    img = np.full((h, w), 220, np.int16)
    cx, cy = w // 2, h // 2
    bw, bh = min(w, h) // 3, min(w, h) // 2
    x0, x1, y0, y1 = cx - bw // 2, cx + bw // 2, cy - bh // 2, cy + bh // 2
    if bw >= 8 and bh >= 8:
        for t in range(2):
            img[y0 + t, x0:x1 + 1] = 100
            img[y1 - t, x0:x1 + 1] = 100
            img[y0:y1 + 1, x0 + t] = 100
            img[y0:y1 + 1, x1 - t] = 100
        _disc(img, cx, cy + bh // 2 - 20, max(bw // 2 - 5, 1), 80)
    if contaminants is None:
        contaminants = int(rng.integers(1, 4)) if rng.random() < p_defect else 0
    for _ in range(contaminants):
        big = rng.random() < 0.5
        r = int(rng.integers(15, 30)) if big else int(rng.integers(3, 10))
        val = int(rng.integers(0, 60)) if big else 40
        px = int(rng.integers(max(cx - bw // 3, 0), max(cx + bw // 3, 1)))
        py = int(rng.integers(max(cy - bh // 3, 0), max(cy + bh // 3, 1)))
        _disc(img, px, py, r, val)
    if noise > 0:
        img += rng.integers(-noise, noise + 1, size=(h, w), dtype=np.int16)
    return np.clip(img, 0, 255).astype(np.uint8)


def bottle_batch(n: int, h: int = 1024, w: int = 1280, start_index: int = 0, **kw) -> np.ndarray:
    """(n, h, w) u8, frame i uses seed BASE_SEED + start_index + i."""
    out = np.empty((n, h, w), np.uint8)
    for i in range(n):
        out[i] = bottle_frame(h, w, start_index + i, **kw)
    return out


def high_contamination_frame(h: int = 3000, w: int = 4096, index: int = 0, pitch: int = 30, salt: float = 0.01,
                             noise: int = 3) -> np.ndarray:
    """Config C4: specks r in [1,4] on a jittered grid of `pitch` px plus `salt` dark single pixels
    (>= 10k blobs per 12 MP frame): stresses the CCL merge and the statistics atomics."""
    rng = np.random.default_rng(BASE_SEED + 100000 + index)
    img = np.full((h, w), 220, np.int16)
    gy, gx = np.mgrid[pitch // 2:h:pitch, pitch // 2:w:pitch]
    gy = (gy + rng.integers(-pitch // 3, pitch // 3 + 1, size=gy.shape)).ravel()
    gx = (gx + rng.integers(-pitch // 3, pitch // 3 + 1, size=gx.shape)).ravel()
    rad = rng.integers(1, 5, size=gy.shape[0])
    for r in range(1, 5):
        sel = rad == r
        for dy in range(-r, r + 1):
            for dx in range(-r, r + 1):
                if dx * dx + dy * dy <= r * r:
                    yy = np.clip(gy[sel] + dy, 0, h - 1)
                    xx = np.clip(gx[sel] + dx, 0, w - 1)
                    img[yy, xx] = 40
    if salt > 0:
        k = int(salt * h * w)
        img[rng.integers(0, h, k), rng.integers(0, w, k)] = 20
    if noise > 0:
        img += rng.integers(-noise, noise + 1, size=(h, w), dtype=np.int16)
    return np.clip(img, 0, 255).astype(np.uint8)


def near_threshold_frame(h: int = 1024, w: int = 1280, index: int = 0, *, spots: int = 6, noise: int = 3,
                         background: int = 200) -> np.ndarray:
    """An empty line position (no bottle) with `spots` faint blemishes whose contrast straddles the detector's decision
    boundaries: blurred contrast around the adaptive threshold c = 25 (detection.rs:186,211), areas around min_size = 10 and shapes
    whose confidence falls on either side of 0.3 (detection.rs:250,298).  spots=0 gives a noisy frame with nothing to
    find.  Most such frames are NOT rejected -- the bottle frames always are (the reference flags the bottle outline
    itself) -- so these are the inputs that exercise reject == False on realistic data."""
    rng = np.random.default_rng(BASE_SEED + 200000 + index)
    img = np.full((h, w), background, np.int16)
    # slow illumination gradient, as a camera sees it (amplitude well below the threshold)
    yy, xx = np.mgrid[0:h, 0:w]
    img += ((xx * 6) // max(w, 1) + (yy * 4) // max(h, 1)).astype(np.int16)
    for _ in range(spots):
        depth = int(rng.integers(70, 150))             # blur - mean peaks near 0.27 * depth three pixels inside an edge:
                                                       # on either side of c = 25 for depths around 92
        kind = int(rng.integers(0, 3))
        cy, cx = int(rng.integers(20, max(h - 20, 21))), int(rng.integers(20, max(w - 20, 21)))
        if kind == 0:                                  # small square: area around min_size
            s = int(rng.integers(3, 12))
            img[cy:cy + s, cx:cx + s] -= depth
        elif kind == 1:                                # disc
            r = int(rng.integers(2, 12))
            y0, y1, x0, x1 = max(cy - r, 0), min(cy + r + 1, h), max(cx - r, 0), min(cx + r + 1, w)
            sy, sx = np.mgrid[y0:y1, x0:x1]
            sel = (sx - cx) ** 2 + (sy - cy) ** 2 <= r * r
            img[y0:y1, x0:x1][sel] -= depth
        else:                                          # thin streak: low fill of its bounding box -> high shape score
            ln = int(rng.integers(6, 40))
            for t in range(ln):
                y, x = cy + t // 3, cx + t
                if 0 <= y < h and 0 <= x < w:
                    img[y, x] -= depth
    if noise > 0:
        img += rng.integers(-noise, noise + 1, size=(h, w), dtype=np.int16)
    return np.clip(img, 0, 255).astype(np.uint8)
