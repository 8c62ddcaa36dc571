"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/hv_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
Nothing under heimdall-vision_b200/ imports it; the product path has no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libhv_oracle.so")


class HvoBlob(C.Structure):
    _fields_ = [("area", C.c_uint32), ("ymin", C.c_uint32), ("ymax", C.c_uint32), ("xmin", C.c_uint32),
                ("xmax", C.c_uint32), ("sum_y", C.c_uint64), ("sum_x", C.c_uint64)]


class HvoDefect(C.Structure):
    _fields_ = [("y", C.c_int32), ("x", C.c_int32), ("size", C.c_double), ("confidence", C.c_double),
                ("ymin", C.c_int32), ("xmin", C.c_int32), ("ymax", C.c_int32), ("xmax", C.c_int32),
                ("label", C.c_uint32)]


class HvoParams(C.Structure):
    _fields_ = [("min_size", C.c_double), ("max_size", C.c_double), ("threshold", C.c_double),
                ("gauss_ksize", C.c_int32), ("gauss_sigma", C.c_double), ("morph_open_k", C.c_int32),
                ("morph_close_k", C.c_int32)]


class HvoContour(C.Structure):
    _fields_ = [("y", C.c_int32), ("x", C.c_int32), ("confidence", C.c_double)]


class HvoContourRec(C.Structure):
    _fields_ = [("y", C.c_int32), ("x", C.c_int32), ("area", C.c_double), ("pixel_count", C.c_uint64),
                ("points_offset", C.c_uint64)]


BLOB_DTYPE = np.dtype([("area", "<u4"), ("ymin", "<u4"), ("ymax", "<u4"), ("xmin", "<u4"), ("xmax", "<u4"),
                       ("_pad", "<u4"), ("sum_y", "<u8"), ("sum_x", "<u8")])
assert BLOB_DTYPE.itemsize == C.sizeof(HvoBlob)

_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/libhv_oracle.so with the committed Makefile (gcc only)."""
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(os.path.join(_HERE, f))
                                              for f in ("hv_oracle.c", "hv_oracle.h", "Makefile"))):
        subprocess.run(["make", "-C", _HERE, "-B", "libhv_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        u8p, i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_int32)
        L.hvo_gray.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p]
        L.hvo_gray.restype = C.c_int
        L.hvo_box_blur.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        L.hvo_box_blur.restype = None
        L.hvo_adaptive_threshold.argtypes = [u8p, C.c_int, C.c_int, C.c_int32, C.c_int, u8p]
        L.hvo_adaptive_threshold.restype = None
        L.hvo_global_threshold.argtypes = [u8p, C.c_int, C.c_int, C.c_uint8, C.c_int, u8p]
        L.hvo_global_threshold.restype = None
        L.hvo_f64_as_i32.argtypes = [C.c_double]
        L.hvo_f64_as_i32.restype = C.c_int32
        L.hvo_label4.argtypes = [u8p, C.c_int, C.c_int, C.c_int, i32p, C.POINTER(HvoBlob), C.c_size_t, i32p]
        L.hvo_label4.restype = C.c_int64
        L.hvo_score_blobs.argtypes = [u8p, u8p, C.c_int, C.c_int, C.POINTER(HvoBlob), C.c_size_t, C.c_double,
                                      C.c_double, C.POINTER(HvoDefect), C.c_size_t]
        L.hvo_score_blobs.restype = C.c_int64
        L.hvo_detect_contamination.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.POINTER(HvoParams), u8p, u8p, u8p,
                                               i32p, C.POINTER(C.c_int64), C.POINTER(HvoDefect), C.c_size_t]
        L.hvo_detect_contamination.restype = C.c_int64
        L.hvo_preprocess_image.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        L.hvo_preprocess_image.restype = C.c_int
        L.hvo_apply_threshold.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_uint8, C.c_int, C.c_int, u8p]
        L.hvo_apply_threshold.restype = C.c_int
        L.hvo_basic_pipeline.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p]
        L.hvo_basic_pipeline.restype = C.c_int
        L.hvo_contamination_pipeline.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p, C.POINTER(HvoContour),
                                                 C.c_size_t]
        L.hvo_contamination_pipeline.restype = C.c_int64
        L.hvo_find_contours.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                        C.POINTER(HvoContourRec), C.c_size_t, i32p]
        L.hvo_find_contours.restype = C.c_int64
        L.hvo_gaussian_blur.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_double, u8p]
        L.hvo_gaussian_blur.restype = C.c_int
        L.hvo_gaussian_kernel_q8.argtypes = [C.c_int, C.c_double, C.POINTER(C.c_uint16)]
        L.hvo_gaussian_kernel_q8.restype = C.c_int
        L.hvo_morph.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        L.hvo_morph.restype = C.c_int
        L.hvo_bayer_to_rgb.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p]
        L.hvo_bayer_to_rgb.restype = C.c_int
        L.hvo_yuyv_to_rgb.argtypes = [u8p, C.c_int, C.c_int, u8p]
        L.hvo_yuyv_to_rgb.restype = C.c_int
        L.hvo_version.restype = C.c_char_p
        _lib = L
    return _lib


class OracleError(ValueError):
    pass


_ERR = {-1: "Invalid image dimensions: expected 3D array", -2: "stage requires a grayscale image",
        -3: "capacity exceeded", -4: "bad argument"}


def _check(rc: int) -> int:
    if rc < 0:
        raise OracleError(_ERR.get(int(rc), f"oracle error {rc}"))
    return int(rc)


def _u8(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


def _p(a: np.ndarray, t=C.c_uint8):
    return a.ctypes.data_as(C.POINTER(t))


def _hwc(img: np.ndarray) -> Tuple[np.ndarray, int, int, int]:
    img = _u8(img)
    if img.ndim == 2:
        img = img[:, :, None]
    if img.ndim != 3:
        raise OracleError("Invalid image dimensions: expected 3D array")
    h, w, c = img.shape
    return np.ascontiguousarray(img), h, w, c


def gray(img) -> np.ndarray:
    img, h, w, c = _hwc(img)
    out = np.empty((h, w), np.uint8)
    _check(lib().hvo_gray(_p(img), h, w, c, _p(out)))
    return out


def box_blur(src, radius: int = 2) -> np.ndarray:
    src, h, w, c = _hwc(src)
    out = np.empty_like(src)
    lib().hvo_box_blur(_p(src), h, w, c, radius, _p(out))
    return out[:, :, 0] if c == 1 else out


def adaptive_threshold(src, c: int, inverse: bool = True) -> np.ndarray:
    src = _u8(src)
    h, w = src.shape
    out = np.empty((h, w), np.uint8)
    lib().hvo_adaptive_threshold(_p(src), h, w, int(c), int(inverse), _p(out))
    return out


def global_threshold(src, thr: int = 127, inverse: bool = False) -> np.ndarray:
    src = _u8(src)
    h, w = src.shape
    out = np.empty((h, w), np.uint8)
    lib().hvo_global_threshold(_p(src), h, w, int(thr), int(inverse), _p(out))
    return out


def f64_as_i32(v: float) -> int:
    return int(lib().hvo_f64_as_i32(float(v)))


def label4(mask, fg_gt127: bool = False, want_pop_order: bool = False):
    """Returns (labels int32 HxW, blobs structured array[, pop_order])."""
    mask = _u8(mask)
    h, w = mask.shape
    labels = np.empty((h, w), np.int32)
    cap = h * w // 2 + 1
    blobs = np.zeros(cap, BLOB_DTYPE)
    pop = np.empty(h * w, np.int32) if want_pop_order else None
    n = _check(lib().hvo_label4(_p(mask), h, w, int(fg_gt127), _p(labels, C.c_int32),
                                blobs.ctypes.data_as(C.POINTER(HvoBlob)), cap,
                                _p(pop, C.c_int32) if pop is not None else None))
    blobs = blobs[:n].copy()
    if want_pop_order:
        return labels, blobs, pop
    return labels, blobs


@dataclass
class DetectResult:
    gray: np.ndarray
    blur: np.ndarray
    mask: np.ndarray
    labels: np.ndarray
    ncomp: int
    defects: List[dict]

    @property
    def reject(self) -> bool:  # heimdall/inspection/base_inspector.py:40-42
        return len(self.defects) > 0


def _defects_to_list(arr, n) -> List[dict]:
    return [dict(position=(arr[i].y, arr[i].x), size=arr[i].size, confidence=arr[i].confidence,
                 bbox=(arr[i].ymin, arr[i].xmin, arr[i].ymax, arr[i].xmax), label=arr[i].label)
            for i in range(n)]


def detect_contamination(img, min_size: float = 10.0, max_size: float = 3000.0, threshold: float = 25.0,
                         gauss_ksize: int = 0, gauss_sigma: float = 0.0, morph_open_k: int = 0,
                         morph_close_k: int = 0, cap: Optional[int] = None,
                         want_intermediates: bool = True) -> DetectResult:
    img, h, w, c = _hwc(img)
    p = HvoParams(min_size, max_size, threshold, gauss_ksize, gauss_sigma, morph_open_k, morph_close_k)
    cap = cap or (h * w // 2 + 1)
    defects = (HvoDefect * cap)()
    ncomp = C.c_int64(0)
    if want_intermediates:
        g = np.empty((h, w), np.uint8)
        b = np.empty((h, w), np.uint8)
        m = np.empty((h, w), np.uint8)
        lab = np.empty((h, w), np.int32)
        n = _check(lib().hvo_detect_contamination(_p(img), h, w, c, C.byref(p), _p(g), _p(b), _p(m),
                                                  _p(lab, C.c_int32), C.byref(ncomp), defects, cap))
    else:
        g = b = m = lab = None
        n = _check(lib().hvo_detect_contamination(_p(img), h, w, c, C.byref(p), None, None, None, None,
                                                  C.byref(ncomp), defects, cap))
    return DetectResult(g, b, m, lab, int(ncomp.value), _defects_to_list(defects, n))


def preprocess_image(img, grayscale: bool = True, blur_size: Optional[int] = None) -> np.ndarray:
    img, h, w, c = _hwc(img)
    out = np.empty((h, w, 1 if grayscale else c), np.uint8)
    _check(lib().hvo_preprocess_image(_p(img), h, w, c, int(grayscale), int(blur_size or 0), _p(out)))
    return out


def apply_threshold(img, threshold_value: int = 127, adaptive: bool = False, inverse: bool = False) -> np.ndarray:
    img, h, w, c = _hwc(img)
    out = np.empty((h, w, 1), np.uint8)
    rc = lib().hvo_apply_threshold(_p(img), h, w, c, int(threshold_value), int(adaptive), int(inverse), _p(out))
    if rc == -2:
        raise OracleError("Image processing error: Thresholding requires a grayscale image")
    _check(rc)
    return out


def basic_pipeline(img) -> np.ndarray:
    img, h, w, c = _hwc(img)
    out = np.empty((h, w, 3), np.uint8)
    _check(lib().hvo_basic_pipeline(_p(img), h, w, c, _p(out)))
    return out


def contamination_pipeline(img):
    img, h, w, c = _hwc(img)
    out = np.empty((h, w, 3), np.uint8)
    cap = h * w // 2 + 1
    cont = (HvoContour * cap)()
    n = _check(lib().hvo_contamination_pipeline(_p(img), h, w, c, _p(out), cont, cap))
    return out, [(cont[i].y, cont[i].x, cont[i].confidence) for i in range(n)]


def find_contours(img, min_area: float = 10.0, max_area: float = 10000.0) -> List[dict]:
    img, h, w, c = _hwc(img)
    cap = h * w // 2 + 1
    recs = (HvoContourRec * cap)()
    pop = np.empty(h * w, np.int32)
    rc = lib().hvo_find_contours(_p(img), h, w, c, min_area, max_area, recs, cap, _p(pop, C.c_int32))
    if rc == -2:
        raise OracleError("Detection error: Contour detection requires a grayscale or binary image")
    n = _check(rc)
    out = []
    for i in range(n):
        d = dict(position=(recs[i].y, recs[i].x), area=recs[i].area, pixel_count=int(recs[i].pixel_count))
        if recs[i].points_offset != 2 ** 64 - 1:
            o = int(recs[i].points_offset)
            pts = pop[o:o + int(recs[i].pixel_count)]
            d["points"] = [(int(q) // w, int(q) % w) for q in pts]
        out.append(d)
    return out


def gaussian_kernel_q8(ksize: int, sigma: float) -> np.ndarray:
    k = np.zeros(ksize, np.uint16)
    _check(lib().hvo_gaussian_kernel_q8(ksize, float(sigma), _p(k, C.c_uint16)))
    return k


def gaussian_blur(src, ksize: int, sigma: float = 0.0) -> np.ndarray:
    src = _u8(src)
    h, w = src.shape
    out = np.empty((h, w), np.uint8)
    _check(lib().hvo_gaussian_blur(_p(src), h, w, ksize, float(sigma), _p(out)))
    return out


MORPH_ERODE, MORPH_DILATE, MORPH_OPEN, MORPH_CLOSE = 0, 1, 2, 3


def morph(src, op: int, k: int) -> np.ndarray:
    src = _u8(src)
    h, w = src.shape
    out = np.empty((h, w), np.uint8)
    _check(lib().hvo_morph(_p(src), h, w, op, k, _p(out)))
    return out


BAYER_PATTERNS = {"RG": 0, "GB": 1, "GR": 2, "BG": 3}


def bayer_to_rgb(bayer, pattern: str) -> np.ndarray:
    """cv2.cvtColor(bayer, COLOR_Bayer<pattern>2RGB): the conversion rust/heimdall-camera/src/lib.rs:226-245 names."""
    b = _u8(np.asarray(bayer))
    if b.ndim != 2:
        raise OracleError("bayer frame must be (h, w)")
    out = np.empty(b.shape + (3,), np.uint8)
    _check(lib().hvo_bayer_to_rgb(_p(b), b.shape[0], b.shape[1], BAYER_PATTERNS[pattern], _p(out)))
    return out


def yuyv_to_rgb(yuyv) -> np.ndarray:
    """cv2.cvtColor(yuyv, COLOR_YUV2RGB_YUYV) for an (h, w, 2) u8 frame (lib.rs:246-250)."""
    s = _u8(np.asarray(yuyv))
    if s.ndim != 3 or s.shape[2] != 2:
        raise OracleError("YUYV frame must be (h, w, 2)")
    out = np.empty(s.shape[:2] + (3,), np.uint8)
    _check(lib().hvo_yuyv_to_rgb(_p(s), s.shape[0], s.shape[1], _p(out)))
    return out
