"""heimdall_core.processing -- rust/heimdall-core/src/processing.rs pyfunctions (preprocess_image, apply_threshold)."""
from __future__ import annotations

from typing import Optional

import numpy as np


def _image3(image) -> np.ndarray:
    if not isinstance(image, np.ndarray) or image.dtype != np.uint8 or image.ndim != 3:
        raise TypeError("argument 'image': expected a 3-dimensional numpy array of uint8")
    return np.ascontiguousarray(image)


def preprocess_image(image, grayscale: Optional[bool] = None, blur_size: Optional[int] = None) -> np.ndarray:
    """processing.rs:30-101: f64 grayscale (default True) then (2*(blur_size/2)+1)^2 box mean on the interior."""
    from .batch import default_detector
    img = _image3(image)
    g = True if grayscale is None else bool(grayscale)
    b = 0 if blur_size is None else int(blur_size)
    return default_detector().preprocess_image(img, g, b)


def apply_threshold(image, threshold_value: Optional[int] = None, adaptive: Optional[bool] = None,
                    inverse: Optional[bool] = None) -> np.ndarray:
    """processing.rs:104-185: global (`> thr`, default 127) or adaptive 11x11 mean with c = 2; `inverse` flips."""
    from .batch import default_detector
    img = _image3(image)
    thr = 127 if threshold_value is None else int(threshold_value)
    if not 0 <= thr <= 255:
        raise OverflowError("threshold_value out of range for u8")
    return default_detector().apply_threshold(img, thr, bool(adaptive), bool(inverse))
