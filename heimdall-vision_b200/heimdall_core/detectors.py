"""heimdall_core.detectors -- the reference's Python detector contract on the GPU.

`ContaminationDetector` mirrors heimdall/detectors/contamination_detector.py: same constructor (name, config with the keys
min_contaminant_size / max_contaminant_size / contrast_threshold / min_confidence / use_color and their defaults, :26-38),
same `detect(image, context) -> List[Defect]` and `__call__` (heimdall/detectors/base.py:41-84).  The work is done by
hv_python_detect (OpenCV-exact stages, external contours, polygon areas and moments, scores, filters); the returned `Defect`
objects carry the reference's metadata keys except "contour" (the CHAIN_APPROX_SIMPLE point list is not produced).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import numpy as np

from .batch import Detector, default_detector
from .results import Defect


class DefectDetector:
    """heimdall/detectors/base.py:41-84."""

    def __init__(self, name: str, config: Optional[Dict[str, Any]] = None):
        self.name = name
        self.config = config or {}

    def detect(self, image: np.ndarray, context: Optional[Dict[str, Any]] = None) -> List[Defect]:
        raise NotImplementedError("Subclasses must implement this method")

    def __call__(self, image: np.ndarray, context: Optional[Dict[str, Any]] = None) -> List[Defect]:
        return self.detect(image, {} if context is None else context)


class ContaminationDetector(DefectDetector):
    def __init__(self, name: str = "contamination_detector", config: Optional[Dict[str, Any]] = None,
                 detector: Optional[Detector] = None):
        super().__init__(name, config)
        self.min_contaminant_size = self.config.get("min_contaminant_size", 10)
        self.max_contaminant_size = self.config.get("max_contaminant_size", 3000)
        self.contrast_threshold = self.config.get("contrast_threshold", 15)
        self.min_confidence = self.config.get("min_confidence", 0.25)
        self.use_color = self.config.get("use_color", True)
        self._det = detector

    def detect(self, image: np.ndarray, context: Optional[Dict[str, Any]] = None) -> List[Defect]:
        img = np.asarray(image)
        if img.dtype != np.uint8 or img.ndim not in (2, 3) or (img.ndim == 3 and img.shape[2] != 3):
            raise ValueError("ContaminationDetector.detect needs an (h, w) gray or (h, w, 3) BGR uint8 image")
        det = self._det or default_detector()
        out = det.python_detect(img, self.min_contaminant_size, self.max_contaminant_size, self.contrast_threshold,
                                self.min_confidence, bool(self.use_color))
        return [Defect("contamination", d["position"], d["size"], d["confidence"],
                       {"intensity_diff": d["intensity_diff"], "shape_score": d["shape_score"], "color_score": d["color_score"],
                        "bounding_box": d["bounding_box"]}) for d in out]
