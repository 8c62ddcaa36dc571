"""heimdall_core.results -- the result side: what the reference does with a frame's defect list.

Mirrors, over the C ABI (hv_export_results, hv_draw_overlays):
  * `InspectionResult` / `Defect.to_dict()` (heimdall/inspection/base_inspector.py:11-64, heimdall/detectors/base.py:7-38):
    per-frame records with the same attributes and the same `to_dict()` keys, built from a `BatchResult`;
  * the dashboard's running statistics (dashboard.py:38-46, updated per image at :483-500);
  * overlays: the 7-px crosses of `process_image("contamination")` (rust/heimdall-core/src/processing.rs:371-401), bounding
    boxes (cv2.rectangle, thickness 1) and the dashboard's defect marker (cv2.circle radius 10 thickness 2, dashboard.py:462),
    drawn on the GPU.  Text labels (cv2.putText) are not reproduced.
"""
from __future__ import annotations

import ctypes as C
import time
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _abi as A
from .batch import BatchResult, Detector, _raise, default_detector


class Defect:
    """heimdall/detectors/base.py:7-38 (position is whatever the producing path reports: (row, col) for the Rust path)."""

    def __init__(self, defect_type: str, position: Tuple[int, int], size: float, confidence: float,
                 metadata: Optional[Dict[str, Any]] = None):
        self.defect_type = defect_type
        self.position = position
        self.size = size
        self.confidence = confidence
        self.metadata = metadata or {}

    def __str__(self) -> str:
        return f"Defect({self.defect_type}, pos={self.position}, size={self.size:.1f}, conf={self.confidence:.2f})"

    def to_dict(self) -> Dict[str, Any]:
        return {"type": self.defect_type, "position": self.position, "size": self.size, "confidence": self.confidence,
                **self.metadata}


class InspectionResult:
    """heimdall/inspection/base_inspector.py:11-64: same attributes, properties and to_dict() keys."""

    def __init__(self, inspection_id: str, timestamp: float, success: bool, defects: Optional[List[Defect]] = None,
                 images: Optional[Dict[str, np.ndarray]] = None, metadata: Optional[Dict[str, Any]] = None):
        self.inspection_id = inspection_id
        self.timestamp = timestamp
        self.success = success
        self.defects = defects or []
        self.images = images or {}
        self.metadata = metadata or {}
        self.processing_time = self.metadata.get("processing_time", 0)

    @property
    def has_defects(self) -> bool:
        return len(self.defects) > 0

    @property
    def defect_count(self) -> int:
        return len(self.defects)

    def to_dict(self) -> Dict[str, Any]:
        return {"inspection_id": self.inspection_id, "timestamp": self.timestamp, "success": self.success,
                "has_defects": self.has_defects, "defect_count": self.defect_count,
                "defects": [d.to_dict() for d in self.defects], "processing_time": self.processing_time,
                "metadata": self.metadata}

    def __str__(self) -> str:
        return f"InspectionResult(id={self.inspection_id}, success={self.success}, defects={self.defect_count})"


class DashboardStats:
    """dashboard.py:38-46 `processing_stats`, updated by hv_export_results exactly as dashboard.py:483-500 does per image."""

    def __init__(self, start_time: Optional[float] = None):
        self._s = A.hv_dashboard_stats()
        self._s.start_time = time.time() if start_time is None else float(start_time)

    def as_dict(self) -> Dict[str, Any]:
        return {"total_images": int(self._s.total_images), "total_defects": int(self._s.total_defects),
                "avg_processing_time": float(self._s.avg_processing_time_ms), "defect_rate": float(self._s.defect_rate),
                "start_time": float(self._s.start_time)}


def export_batch(res: BatchResult, *, inspector_id: str = "contamination", timestamp: Optional[float] = None,
                 processing_time: float = 0.0, first_sequence: int = 0, stats: Optional[DashboardStats] = None,
                 defect_type: str = "contamination") -> List[InspectionResult]:
    """One `InspectionResult` per frame of a batch (hv_export_results); `stats` is updated once per frame."""
    n = len(res.frames)
    recs = (A.hv_inspection_record * max(n, 1))()
    fr = np.ascontiguousarray(res.frames)
    ts = time.time() if timestamp is None else float(timestamp)
    st = A.lib.hv_export_results(fr.ctypes.data_as(C.POINTER(A.hv_frame_result)), n, ts, float(processing_time),
                                 int(first_sequence), recs, C.byref(stats._s) if stats is not None else None)
    if st != A.HV_OK:
        _raise(st, None)
    out = []
    for f in range(n):
        r = recs[f]
        defects = [Defect(defect_type, (int(d["y"]), int(d["x"])), float(d["size"]), float(d["confidence"]),
                          {"bounding_box": (int(d["xmin"]), int(d["ymin"]), int(d["xmax"] - d["xmin"] + 1),
                                            int(d["ymax"] - d["ymin"] + 1))})
                   for d in res.defects[r.defects_offset:r.defects_offset + r.defect_count]]
        ir = InspectionResult(f"{inspector_id}_{int(r.sequence)}", float(r.timestamp), bool(r.success), defects,
                              metadata={"inspector_id": inspector_id, "processing_time": float(r.processing_time)})
        assert ir.has_defects == bool(r.has_defects) and ir.defect_count == int(r.defect_count)
        out.append(ir)
    return out


CROSS, BOX, MARKER = A.HV_OVERLAY_CROSS, A.HV_OVERLAY_BOX, A.HV_OVERLAY_MARKER


def draw_overlays(image: np.ndarray, items: Sequence[tuple], detector: Optional[Detector] = None) -> np.ndarray:
    """items: (CROSS, y, x[, color]) | (MARKER, y, x[, color]) | (BOX, y, x, y1, x1[, color]); color defaults to (0, 0, 255).
    Returns a new (h, w, 3) u8 image with the overlays drawn in list order (hv_draw_overlays)."""
    img = np.ascontiguousarray(image, np.uint8)
    if img.ndim == 2:
        img = np.repeat(img[:, :, None], 3, axis=2)   # cv2.COLOR_GRAY2BGR
    if img.ndim != 3 or img.shape[2] != 3:
        raise ValueError("draw_overlays needs an (h, w) or (h, w, 3) uint8 image")
    out = img.copy()
    arr = (A.hv_overlay * max(len(items), 1))()
    for i, it in enumerate(items):
        kind = int(it[0])
        if kind == BOX:
            y, x, y1, x1 = (int(v) for v in it[1:5])
            color = it[5] if len(it) > 5 else (0, 0, 255)
        else:
            y, x = int(it[1]), int(it[2])
            y1 = x1 = 0
            color = it[3] if len(it) > 3 else (0, 0, 255)
        arr[i] = A.hv_overlay(kind, y, x, y1, x1, (C.c_uint8 * 3)(*[int(c) for c in color]), 0)
    det = detector or default_detector()
    h, w = out.shape[:2]
    with det._lock:
        st = A.lib.hv_draw_overlays(det._ctx, out.ctypes.data, h, w, arr, len(items))
        if st != A.HV_OK:
            _raise(st, det._ctx)
    return out


def visualize_defects(image: np.ndarray, res: BatchResult, f: int = 0, *, boxes: bool = True, markers: bool = True,
                      detector: Optional[Detector] = None) -> np.ndarray:
    """The dashboard's picture of one frame (dashboard.py:456-462): a marker at every defect position, plus its box."""
    items = []
    for d in res.defects_of(f):
        if boxes:
            items.append((BOX, int(d["ymin"]), int(d["xmin"]), int(d["ymax"]), int(d["xmax"])))
        if markers:
            items.append((MARKER, int(d["y"]), int(d["x"])))
    return draw_overlays(image, items, detector)
