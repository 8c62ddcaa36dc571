"""heimdall_core.acquisition -- rust/heimdall-core/src/acquisition.rs:10-107.

Host-side synthetic frame source (the reference's `acquire_image` ignores its source and always simulates); it is not
part of the accelerated path, it only completes the module surface."""
from __future__ import annotations

from typing import Optional

import numpy as np


def simulate_bottle_image(height: int = 480, width: int = 640) -> np.ndarray:
    """acquisition.rs:57-107: 220 background, 1-px bottle outline of 100, filled disc of 80."""
    img = np.full((height, width, 3), 220, np.uint8)
    cx, cy = width // 2, height // 2
    bw, bh = min(width, height) // 3, min(width, height) // 2
    y0, y1, x0, x1 = cy - bh // 2, cy + bh // 2, cx - bw // 2, cx + bw // 2
    img[y0, x0:x1] = 100
    img[y1 - 1, x0:x1] = 100
    img[y0:y1, x0] = 100
    img[y0:y1, x1 - 1] = 100
    ccy, r = cy + bh // 2 - 20, bw // 2 - 5
    yy, xx = np.mgrid[0:height, 0:width]
    dist = np.sqrt(((xx - cx) ** 2 + (yy - ccy) ** 2).astype(np.float64))
    img[dist < float(r)] = 80
    return img


def acquire_image(source_type: str, params: Optional[dict] = None) -> np.ndarray:
    if source_type in ("simulation", "file", "camera"):
        return simulate_bottle_image()
    raise ValueError(f"Unsupported source type: {source_type}")
