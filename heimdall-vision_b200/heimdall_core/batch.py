"""Batch / stream interface to the CUDA backend -- the superset of the reference's one-frame-per-call API.

`Detector` owns one hv_ctx (one CUDA device).  The reference processes one frame per `detect_contamination` call on one
CPU thread (rust/heimdall-core/src/lib.rs:95-143); here whole batches (one second of line at 90k BPH = 25 frames) go
through one call, host-fed (`detect_batch`, `submit`/`wait`) or device-resident (`detect_device`, `enqueue_device`).
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _abi as A

_lib = A.lib


class HeimdallCudaError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


def _raise(status: int, ctx) -> None:
    msg = _lib.hv_last_error(ctx) if ctx else b""
    text = (msg or b"").decode() or _lib.hv_status_string(status).decode()
    # same mapping as the reference: its Result errors become ValueError (processing.rs:23-27, detection.rs:29-33)
    if status in (A.HV_ERR_INVALID_DIMENSIONS, A.HV_ERR_CHANNELS, A.HV_ERR_UNSUPPORTED, A.HV_ERR_INVALID_ARGUMENT):
        raise ValueError(text)
    raise HeimdallCudaError(status, text)


def make_params(min_size: float = 10.0, max_size: float = 3000.0, threshold: float = 25.0, *,
                min_confidence: float = 0.3, blur_mode: int = A.HV_BLUR_BOX, blur_ksize: int = 5,
                gauss_sigma: float = 0.0, morph_open_k: int = 0, morph_close_k: int = 0) -> A.hv_params:
    p = A.hv_params()
    _lib.hv_params_default(C.byref(p))
    p.min_size, p.max_size, p.threshold = float(min_size), float(max_size), float(threshold)
    p.min_confidence = float(min_confidence)
    p.blur_mode, p.blur_ksize, p.gauss_sigma = int(blur_mode), int(blur_ksize), float(gauss_sigma)
    p.morph_open_k, p.morph_close_k = int(morph_open_k), int(morph_close_k)
    return p


DEFECT_DTYPE = np.dtype([("y", "<i4"), ("x", "<i4"), ("size", "<f8"), ("confidence", "<f8"), ("ymin", "<i4"),
                         ("xmin", "<i4"), ("ymax", "<i4"), ("xmax", "<i4"), ("label", "<u4"), ("frame", "<u4")])
RESULT_DTYPE = np.dtype([("n_components", "<u4"), ("n_defects", "<u4"), ("defects_offset", "<u4"),
                         ("rejected", "<u4"), ("fg_pixels", "<u4"), ("status", "<i4")])
BLOB_DTYPE = np.dtype([("area", "<u4"), ("ymin", "<u4"), ("ymax", "<u4"), ("xmin", "<u4"), ("xmax", "<u4"),
                       ("reserved", "<u4"), ("sum_y", "<u8"), ("sum_x", "<u8")])
assert DEFECT_DTYPE.itemsize == C.sizeof(A.hv_defect)
assert RESULT_DTYPE.itemsize == C.sizeof(A.hv_frame_result)
assert BLOB_DTYPE.itemsize == C.sizeof(A.hv_blob)


class BatchResult:
    """Per-frame results of one batch: `frames` (structured array, RESULT_DTYPE) and `defects` (DEFECT_DTYPE)."""

    def __init__(self, frames: np.ndarray, defects: np.ndarray, status: int, debug: Optional[dict] = None):
        self.frames = frames
        self.defects = defects
        self.status = status
        self.debug = debug or {}

    def defects_of(self, f: int) -> np.ndarray:
        o, n = int(self.frames["defects_offset"][f]), int(self.frames["n_defects"][f])
        return self.defects[o:o + n]

    def as_dicts(self, f: int) -> List[dict]:
        """The reference's defect dicts (lib.rs:116-137): position=(row, col), size, confidence, metadata={}."""
        return [{"position": (int(d["y"]), int(d["x"])), "size": float(d["size"]),
                 "confidence": float(d["confidence"]), "metadata": {}} for d in self.defects_of(f)]

    @property
    def rejected(self) -> np.ndarray:
        return self.frames["rejected"].astype(bool)


class Detector:
    """One CUDA device, one hv_ctx.  The context is not thread-safe; every call into it is serialised with a lock."""

    def __init__(self, device: int = 0, *, max_blobs_per_frame: int = 0, max_defects_per_frame: int = 0,
                 num_slots: int = 0, profile: bool = False, keep_blur: bool = False, force_generic: bool = False,
                 global_ccl: bool = False, phase_timing: bool = False, defer_tail: bool = False):
        cfg = A.hv_config()
        _lib.hv_config_default(C.byref(cfg))
        cfg.max_blobs_per_frame = max_blobs_per_frame
        cfg.max_defects_per_frame = max_defects_per_frame
        cfg.num_slots = num_slots
        cfg.flags = ((A.HV_FLAG_PROFILE if profile else 0) | (A.HV_FLAG_KEEP_BLUR if keep_blur else 0) |
                     (A.HV_FLAG_FORCE_GENERIC if force_generic else 0) |
                     (A.HV_FLAG_GLOBAL_CCL if global_ccl else 0) |
                     (A.HV_FLAG_PHASE_TIMING if phase_timing else 0) |
                     (A.HV_FLAG_DEFER_TAIL if defer_tail else 0))
        self._ctx = C.c_void_p()
        self._lock = threading.RLock()
        self._inflight_frames: Dict[int, tuple] = {}  # ticket -> raw camera frames that must outlive their H2D copy
        self.device = device
        self.defect_cap = max_defects_per_frame or 256
        st = _lib.hv_create(device, C.byref(cfg), C.byref(self._ctx))
        if st != A.HV_OK:
            self._ctx = C.c_void_p()
            msg = (_lib.hv_last_error(None) or b"").decode()
            raise HeimdallCudaError(st, f"heimdall_core: cannot create CUDA context on device {device}: {msg}")

    def close(self) -> None:
        if getattr(self, "_ctx", None) is not None and self._ctx:
            _lib.hv_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers -----------------------------------------------------------------------------------------------
    @staticmethod
    def _as_batch(frames: np.ndarray) -> Tuple[np.ndarray, int, int, int, int]:
        a = np.asarray(frames)
        if a.dtype != np.uint8:
            raise TypeError("frames must be uint8")
        if a.ndim == 2:
            a = a[None, :, :, None]
        elif a.ndim == 3:
            a = a[None]
        elif a.ndim != 4:
            raise TypeError("frames must be (H,W), (H,W,C) or (N,H,W,C)")
        a = np.ascontiguousarray(a)
        n, h, w, c = a.shape
        return a, n, h, w, c

    def _out_arrays(self, n: int, defects_cap: Optional[int]):
        cap = defects_cap if defects_cap is not None else n * self.defect_cap
        return np.zeros(n, RESULT_DTYPE), np.zeros(max(cap, 1), DEFECT_DTYPE), cap

    # ---- host-fed ----------------------------------------------------------------------------------------------
    def detect_batch(self, frames: np.ndarray, params: Optional[A.hv_params] = None, *, debug: Sequence[str] = (),
                     defects_cap: Optional[int] = None, raise_on_capacity: bool = True) -> BatchResult:
        """frames: (N,H,W,C) / (H,W,C) / (H,W) uint8 in host memory.  debug: any of gray, blur, mask, labels, blobs."""
        a, n, h, w, c = self._as_batch(frames)
        res, dfx, cap = self._out_arrays(n, defects_cap)
        total = C.c_size_t(0)
        dbg_arrays: Dict[str, np.ndarray] = {}
        dbg = A.hv_debug_outputs()
        for name in debug:
            if name in ("gray", "blur", "mask"):
                dbg_arrays[name] = np.empty((n, h, w), np.uint8)
            elif name == "labels":
                dbg_arrays[name] = np.empty((n, h, w), np.int32)
            elif name == "blobs":
                stride = h * w // 2 + 1
                dbg_arrays[name] = np.zeros((n, stride), BLOB_DTYPE)
                dbg.blobs_stride = stride
            else:
                raise ValueError(f"unknown debug output {name}")
            setattr(dbg, name, dbg_arrays[name].ctypes.data)
        p = params if params is not None else make_params()
        with self._lock:
            st = _lib.hv_detect_batch(self._ctx, a.ctypes.data, n, h, w, c, 0, 0, C.byref(p),
                                      res.ctypes.data_as(C.POINTER(A.hv_frame_result)),
                                      dfx.ctypes.data_as(C.POINTER(A.hv_defect)), cap, C.byref(total),
                                      C.byref(dbg) if debug else None)
            if st != A.HV_OK and not (st == A.HV_ERR_CAPACITY and not raise_on_capacity):
                _raise(st, self._ctx)
        if "blobs" in dbg_arrays:
            dbg_arrays["blobs"] = [dbg_arrays["blobs"][f, :min(int(res["n_components"][f]), dbg_arrays["blobs"].shape[1])]
                                   for f in range(n)]
        return BatchResult(res, dfx[:total.value], st, dbg_arrays)

    def submit(self, frames_ptr: int, n: int, h: int, w: int, c: int = 1, params: Optional[A.hv_params] = None) -> int:
        """Asynchronous host-fed batch from a raw host pointer (ideally hv_host_alloc'ed / pinned)."""
        t = C.c_int64(0)
        p = params if params is not None else make_params()
        with self._lock:
            st = _lib.hv_submit(self._ctx, frames_ptr, n, h, w, c, 0, 0, C.byref(p), C.byref(t))
            if st != A.HV_OK:
                _raise(st, self._ctx)
        return t.value

    def wait(self, ticket: int, n: int, defects_cap: Optional[int] = None) -> BatchResult:
        res, dfx, cap = self._out_arrays(n, defects_cap)
        total = C.c_size_t(0)
        with self._lock:
            try:
                st = _lib.hv_wait(self._ctx, ticket, res.ctypes.data_as(C.POINTER(A.hv_frame_result)),
                                  dfx.ctypes.data_as(C.POINTER(A.hv_defect)), cap, C.byref(total))
            finally:
                self._inflight_frames.pop(ticket, None)  # hv_wait has returned: the H2D copy of the raw frames is over
            if st not in (A.HV_OK, A.HV_ERR_CAPACITY):
                _raise(st, self._ctx)
        return BatchResult(res, dfx[:total.value], st)

    # ---- frame feed (N1): raw camera frames ----------------------------------------------------------------------
    def convert_frame(self, frame) -> np.ndarray:
        """One `camera.CameraFrame` -> the (h, w, c) u8 image the reference hands to the detector."""
        fa = frame.as_abi()
        ch = int(_lib.hv_frame_channels(fa.pixel_format))
        out = np.empty((int(fa.height), int(fa.width), max(ch, 1)), np.uint8)
        oc = C.c_int32(0)
        with self._lock:
            st = _lib.hv_convert_frame(self._ctx, C.byref(fa), out.ctypes.data, C.byref(oc))
            if st != A.HV_OK:
                _raise(st, self._ctx)
        return out

    def submit_frames(self, frames, params: Optional[A.hv_params] = None) -> int:
        """Asynchronous batch of raw camera frames (identical geometry and pixel format); collect with `wait`."""
        arr = (A.hv_camera_frame * len(frames))(*[f.as_abi() for f in frames])
        t = C.c_int64(0)
        p = params if params is not None else make_params()
        with self._lock:
            st = _lib.hv_submit_frames(self._ctx, arr, len(frames), C.byref(p), C.byref(t))
            if st != A.HV_OK:
                _raise(st, self._ctx)
            self._inflight_frames[t.value] = (arr, list(frames))  # the raw data must outlive the copy; wait() drops it
        return t.value

    def detect_frames(self, frames, params: Optional[A.hv_params] = None) -> BatchResult:
        return self.wait(self.submit_frames(frames, params), len(frames))

    def host_alloc(self, nbytes: int) -> int:
        p = _lib.hv_host_alloc(self._ctx, nbytes)
        if not p:
            raise MemoryError("hv_host_alloc failed")
        return p

    def host_free(self, ptr: int) -> None:
        _lib.hv_host_free(self._ctx, ptr)

    # ---- device-resident -----------------------------------------------------------------------------------------
    @staticmethod
    def pipeline_depth() -> int:
        """How many sets of output planes to rotate through enqueue_device for full overlap (hv_pipeline_depth)."""
        return int(_lib.hv_pipeline_depth())

    def device_alloc(self, shape, dtype=np.uint8, compressible: bool = True) -> "DeviceArray":
        """Device buffer for detect_device / enqueue_device outputs (hv_device_alloc).  compressible=True asks for L2
        compute-data compression: the mostly-zero mask and label planes then cost less DRAM write time."""
        return DeviceArray(self, shape, dtype, compressible)

    def set_stream(self, cuda_stream: Optional[int]) -> None:
        """Run on the caller's stream (0 = legacy default stream, e.g. torch's current stream); None = own stream."""
        with self._lock:
            st = _lib.hv_set_stream(self._ctx, cuda_stream or None, 0 if cuda_stream is None else 1)
            if st != A.HV_OK:
                _raise(st, self._ctx)

    def enqueue_device(self, d_frames: int, n: int, h: int, w: int, c: int = 1,
                       params: Optional[A.hv_params] = None, d_mask: int = 0, d_labels: int = 0) -> int:
        """Enqueue one device-resident batch (no host synchronisation); returns its ticket.  The results travel to
        page-locked host memory behind the batch's last kernel; `fetch(ticket, n)` returns them.  The context keeps
        the last `pipeline_depth()` batches."""
        p = params if params is not None else make_params()
        t = C.c_int64(0)
        with self._lock:
            st = _lib.hv_enqueue_device(self._ctx, d_frames, n, h, w, c, 0, 0, C.byref(p), d_mask or None,
                                        d_labels or None, C.byref(t))
            if st != A.HV_OK:
                _raise(st, self._ctx)
        return t.value

    def fetch(self, ticket: int, n: int, defects_cap: Optional[int] = None, raise_on_capacity: bool = True) -> BatchResult:
        """Results of the batch `ticket` (hv_fetch_ticket); blocks until its read-back has arrived."""
        res, dfx, cap = self._out_arrays(n, defects_cap)
        return self.fetch_into(ticket, res, dfx, raise_on_capacity)

    def fetch_into(self, ticket: int, res: np.ndarray, dfx: np.ndarray, raise_on_capacity: bool = True) -> BatchResult:
        """Like fetch, into caller-provided arrays (RESULT_DTYPE[n], DEFECT_DTYPE[cap]): no allocation per batch."""
        total = C.c_size_t(0)
        with self._lock:
            st = _lib.hv_fetch_ticket(self._ctx, ticket, res.ctypes.data_as(C.POINTER(A.hv_frame_result)),
                                      dfx.ctypes.data_as(C.POINTER(A.hv_defect)), len(dfx), C.byref(total))
            if st != A.HV_OK and not (st == A.HV_ERR_CAPACITY and not raise_on_capacity):
                _raise(st, self._ctx)
        return BatchResult(res, dfx[:total.value], st)

    def flush(self) -> None:
        """With defer_tail: put the kernel held back for the latest batch onto the stream now (hv_flush)."""
        with self._lock:
            st = _lib.hv_flush(self._ctx)
            if st != A.HV_OK:
                _raise(st, self._ctx)

    def fetch_results(self, n: int, defects_cap: Optional[int] = None, raise_on_capacity: bool = True) -> BatchResult:
        """Results of the most recently enqueued batch."""
        res, dfx, cap = self._out_arrays(n, defects_cap)
        total = C.c_size_t(0)
        with self._lock:
            st = _lib.hv_fetch_results(self._ctx, res.ctypes.data_as(C.POINTER(A.hv_frame_result)),
                                       dfx.ctypes.data_as(C.POINTER(A.hv_defect)), cap, C.byref(total))
            if st != A.HV_OK and not (st == A.HV_ERR_CAPACITY and not raise_on_capacity):
                _raise(st, self._ctx)
        return BatchResult(res, dfx[:total.value], st)

    def fetch_debug(self, n: int, h: int, w: int, names: Sequence[str]) -> Dict[str, np.ndarray]:
        out: Dict[str, np.ndarray] = {}
        dbg = A.hv_debug_outputs()
        for name in names:
            if name in ("gray", "blur", "mask"):
                out[name] = np.empty((n, h, w), np.uint8)
            elif name == "labels":
                out[name] = np.empty((n, h, w), np.int32)
            elif name == "blobs":
                stride = h * w // 2 + 1
                out[name] = np.zeros((n, stride), BLOB_DTYPE)
                dbg.blobs_stride = stride
            else:
                raise ValueError(f"unknown debug output {name}")
            setattr(dbg, name, out[name].ctypes.data)
        with self._lock:
            st = _lib.hv_fetch_debug(self._ctx, C.byref(dbg))
            if st != A.HV_OK:
                _raise(st, self._ctx)
        return out

    def detect_device(self, d_frames: int, n: int, h: int, w: int, c: int = 1, params: Optional[A.hv_params] = None,
                      d_mask: int = 0, d_labels: int = 0) -> BatchResult:
        with self._lock:
            t = self.enqueue_device(d_frames, n, h, w, c, params, d_mask, d_labels)
            return self.fetch(t, n)

    # ---- single-frame utilities (processing.rs / detection.rs pyfunctions) ---------------------------------------
    def preprocess_image(self, img: np.ndarray, grayscale: bool, blur_size: int) -> np.ndarray:
        h, w, c = img.shape
        out = np.empty((h, w, 1 if grayscale else c), np.uint8)
        with self._lock:
            st = _lib.hv_preprocess_image(self._ctx, img.ctypes.data, h, w, c, int(grayscale), int(blur_size),
                                          out.ctypes.data)
            if st != A.HV_OK:
                _raise(st, self._ctx)
        return out

    def apply_threshold(self, img: np.ndarray, thr: int, adaptive: bool, inverse: bool) -> np.ndarray:
        h, w, c = img.shape
        out = np.empty((h, w, 1), np.uint8)
        with self._lock:
            st = _lib.hv_apply_threshold(self._ctx, img.ctypes.data, h, w, c, thr, int(adaptive), int(inverse),
                                         out.ctypes.data)
            if st != A.HV_OK:
                _raise(st, self._ctx)
        return out

    def morphology(self, mask: np.ndarray, open_k: int = 0, close_k: int = 0) -> np.ndarray:
        """cv2 MORPH_OPEN (rect open_k) then MORPH_CLOSE (rect close_k) of a binary (H, W) u8 mask on the GPU."""
        m = np.ascontiguousarray(mask, np.uint8)
        h, w = m.shape[:2]
        out = np.empty((h, w), np.uint8)
        with self._lock:
            st = _lib.hv_morphology(self._ctx, m.ctypes.data, h, w, int(open_k), int(close_k), out.ctypes.data)
            if st != A.HV_OK:
                _raise(st, self._ctx)
        return out

    def python_detector_stages(self, img: np.ndarray, contrast_threshold: float = 25.0, *, blur_ksize: int = 5,
                               block_size: int = 11, morph_open_k: int = 3, morph_close_k: int = 3) -> dict:
        """The stages of the reference's Python detector (contamination_detector.py:58-90) with OpenCV's arithmetic
        (hv_python_detector_stages): {"gray", "blurred", "binary", "labels8", "components"}; components = BLOB_DTYPE rows of
        the 8-connected components of `binary` in raster order of their first pixel."""
        a = np.ascontiguousarray(img, np.uint8)
        if a.ndim == 2:
            a = a[:, :, None]
        h, w, c = a.shape
        p = A.hv_pydet_params()
        _lib.hv_pydet_params_default(C.byref(p))
        p.contrast_threshold, p.blur_ksize, p.block_size = float(contrast_threshold), int(blur_ksize), int(block_size)
        p.morph_open_k, p.morph_close_k = int(morph_open_k), int(morph_close_k)
        out = {"gray": np.empty((h, w), np.uint8), "blurred": np.empty((h, w), np.uint8), "binary": np.empty((h, w), np.uint8),
               "labels8": np.empty((h, w), np.int32)}
        cap = h * w // 2 + 1
        comps = np.zeros(min(cap, 131072), BLOB_DTYPE)
        n = C.c_size_t(0)
        with self._lock:
            st = _lib.hv_python_detector_stages(self._ctx, a.ctypes.data, h, w, c, C.byref(p), out["gray"].ctypes.data,
                                                out["blurred"].ctypes.data, out["binary"].ctypes.data, out["labels8"].ctypes.data,
                                                comps.ctypes.data_as(C.POINTER(A.hv_blob)), len(comps), C.byref(n))
            if st != A.HV_OK:
                _raise(st, self._ctx)
        out["components"] = comps[:n.value]
        return out

    def python_detect(self, img: np.ndarray, min_size: float = 10.0, max_size: float = 3000.0,
                      contrast_threshold: float = 15.0, min_confidence: float = 0.25, use_color: bool = True) -> List[dict]:
        """`ContaminationDetector.detect` of the reference's Python path (contamination_detector.py:44-216; defaults :26-38)
        on the GPU (hv_python_detect): a list of the reference's `Defect.to_dict()` dictionaries -- type, position (x, y),
        size (cv2.contourArea), confidence, intensity_diff, shape_score, color_score, bounding_box -- in cv2.findContours'
        order.  `img`: (h, w) gray or (h, w, 3) BGR uint8.  The metadata entry "contour" is not produced."""
        a = np.ascontiguousarray(img, np.uint8)
        if a.ndim == 2:
            a = a[:, :, None]
        h, w, c = a.shape
        sp = A.hv_pydet_params()
        _lib.hv_pydet_params_default(C.byref(sp))
        sp.contrast_threshold = float(contrast_threshold)
        qp = A.hv_pydet_score_params()
        _lib.hv_pydet_score_params_default(C.byref(qp))
        qp.min_size, qp.max_size, qp.min_confidence, qp.use_color = float(min_size), float(max_size), float(min_confidence), int(use_color)
        cap = 4096
        out = (A.hv_pydefect * cap)()
        n, nc = C.c_size_t(0), C.c_size_t(0)
        with self._lock:
            st = _lib.hv_python_detect(self._ctx, a.ctypes.data, h, w, c, C.byref(sp), C.byref(qp), out, cap, C.byref(n), C.byref(nc))
            if st != A.HV_OK:
                _raise(st, self._ctx)
        return [{"type": "contamination", "position": (int(d.x), int(d.y)), "size": float(d.size), "confidence": float(d.confidence),
                 "intensity_diff": float(d.intensity_diff), "shape_score": float(d.shape_score), "color_score": float(d.color_score),
                 "bounding_box": (int(d.bx), int(d.by), int(d.bw), int(d.bh))} for d in out[:n.value]]

    def find_contours(self, img: np.ndarray, min_area: float, max_area: float, want_labels: bool = True):
        h, w, c = img.shape
        cap = h * w // 2 + 1
        recs = (A.hv_contour * cap)()
        n = C.c_size_t(0)
        labels = np.empty((h, w), np.int32) if want_labels else None
        with self._lock:
            st = _lib.hv_find_contours(self._ctx, img.ctypes.data, h, w, c, float(min_area), float(max_area), recs, cap,
                                       C.byref(n), labels.ctypes.data if labels is not None else None)
            if st != A.HV_OK:
                _raise(st, self._ctx)
        return [recs[i] for i in range(n.value)], labels

    def process_image(self, img: np.ndarray, pipeline: int):
        h, w, c = img.shape
        out = np.empty((h, w, 3), np.uint8)
        cap = h * w // 2 + 1
        centers = (A.hv_center * cap)()
        n = C.c_size_t(0)
        with self._lock:
            st = _lib.hv_process_image(self._ctx, img.ctypes.data, h, w, c, pipeline, out.ctypes.data, centers, cap,
                                       C.byref(n))
            if st != A.HV_OK:
                _raise(st, self._ctx)
        return out, [(int(centers[i].y), int(centers[i].x), float(centers[i].confidence)) for i in range(n.value)]

    # ---- statistics / measurement ----------------------------------------------------------------------------------
    def stats(self) -> dict:
        s = A.hv_line_stats()
        with self._lock:
            st = _lib.hv_stats_get(self._ctx, C.byref(s))
            if st != A.HV_OK:
                _raise(st, self._ctx)
        return {"frames_inspected": s.frames_inspected, "frames_rejected": s.frames_rejected,
                "total_defects": s.total_defects, "total_components": s.total_components,
                "total_defect_area": s.total_defect_area, "total_fg_pixels": s.total_fg_pixels,
                "area_hist": list(s.area_hist), "capacity_errors": s.capacity_errors}

    def stats_reset(self) -> None:
        with self._lock:
            _lib.hv_stats_reset(self._ctx)

    def stats_device_ptr(self) -> int:
        return int(_lib.hv_stats_device_ptr(self._ctx) or 0)

    def phase_times(self) -> List[int]:
        out = (C.c_uint64 * 256)()
        _lib.hv_debug_phase_times(self._ctx, out)
        return [int(x) for x in out]

    def launch_count(self) -> int:
        return int(_lib.hv_launch_count(self._ctx))

    def profile_enable(self, kernels: Optional[Sequence[int]] = None) -> None:
        """Time the given HV_K_* kernels with CUDA events on the launching stream (None = all, [] = off)."""
        mask = 0xFFFFFFFF if kernels is None else sum(1 << k for k in kernels)
        _lib.hv_profile_enable(self._ctx, mask)

    def profile(self) -> Dict[str, dict]:
        """Summed milliseconds and launch counts per kernel since the last call (waits for the events)."""
        ms = (C.c_float * A.HV_K_COUNT)()
        cnt = (C.c_uint32 * A.HV_K_COUNT)()
        _lib.hv_profile_get(self._ctx, ms, cnt)
        return {_lib.hv_kernel_name(k).decode(): {"ms": float(ms[k]), "launches": int(cnt[k])}
                for k in range(A.HV_K_COUNT)}


class DeviceArray:
    """A device buffer owned by a Detector's context: .ptr for the device-resident entry points, .get() / .set() copy."""

    def __init__(self, det: Detector, shape, dtype, compressible: bool):
        self._det = det
        self.shape = tuple(int(x) for x in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p, got = C.c_void_p(), C.c_int32(0)
        st = _lib.hv_device_alloc(det._ctx, self.nbytes, A.HV_ALLOC_COMPRESSIBLE if compressible else 0, C.byref(p), C.byref(got))
        if st != A.HV_OK:
            _raise(st, det._ctx)
        self.ptr = int(p.value)
        self.compressed = bool(got.value)

    def data_ptr(self) -> int:
        return self.ptr

    def get(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        """Copy entries [first, first+count) along axis 0 (default: everything) to a new host array."""
        count = self.shape[0] - first if count is None else count
        out = np.empty((count,) + self.shape[1:], self.dtype)
        row = int(np.prod(self.shape[1:])) * self.dtype.itemsize
        st = _lib.hv_device_read(self._det._ctx, out.ctypes.data, self.ptr + first * row, out.nbytes)
        if st != A.HV_OK:
            _raise(st, self._det._ctx)
        return out

    def set(self, arr: np.ndarray) -> None:
        arr = np.ascontiguousarray(arr, self.dtype)
        if arr.nbytes != self.nbytes:
            raise ValueError("size mismatch")
        st = _lib.hv_device_write(self._det._ctx, self.ptr, arr.ctypes.data, arr.nbytes)
        if st != A.HV_OK:
            _raise(st, self._det._ctx)

    def free(self) -> None:
        if self.ptr and self._det._ctx:
            _lib.hv_device_free(self._det._ctx, self.ptr)
        self.ptr = 0

    def __del__(self):  # pragma: no cover
        try:
            self.free()
        except Exception:
            pass


_unbounded: Dict[int, Tuple[int, Detector]] = {}


def unbounded_detector(h: int, w: int, device: Optional[int] = None) -> Detector:
    """A context whose per-frame tables hold the 4-connectivity maximum of an h x w frame (h*w/2 + 1 components and
    defects).  The reference has no limits (its lists are Vecs); the drop-in wrappers retry here when the default
    context reports HV_ERR_CAPACITY, so callers of the reference API never see a capacity error."""
    import os
    if device is None:
        device = int(os.environ.get("HEIMDALL_CUDA_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    need = h * w // 2 + 1
    with _default_lock:
        have = _unbounded.get(device)
        if have is None or have[0] < need:
            if have is not None:
                have[1].close()
            have = (need, Detector(device, max_blobs_per_frame=need, max_defects_per_frame=need))
            _unbounded[device] = have
        return have[1]


def with_capacity_retry(call, h: int, w: int):
    """call(detector) on the default context; on HV_ERR_CAPACITY once more on the unbounded one."""
    try:
        return call(default_detector())
    except HeimdallCudaError as e:
        if e.status != A.HV_ERR_CAPACITY:
            raise
    return call(unbounded_detector(h, w))


_default: Dict[int, Detector] = {}
_default_lock = threading.Lock()


def default_detector(device: Optional[int] = None) -> Detector:
    import os
    if device is None:
        device = int(os.environ.get("HEIMDALL_CUDA_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    with _default_lock:
        d = _default.get(device)
        if d is None:
            d = Detector(device)
            _default[device] = d
        return d
