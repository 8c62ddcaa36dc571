"""ctypes binding of include/heimdall_cuda.h (the C ABI a Rust `heimdall-cuda` FFI crate would bind).

This module only declares the boundary; all computation happens in libheimdall_cuda.so (hand-written sm_100a
kernels).  There is no CPU fallback: a missing library raises ImportError (so the reference's
`heimdall/rust_bridge.py:20-26` probe reports the backend as unavailable) and a missing GPU makes every call raise
RuntimeError.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HEIMDALL_CUDA_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libheimdall_cuda.so")

HV_OK = 0
HV_ERR_INVALID_DIMENSIONS = -1
HV_ERR_INVALID_ARGUMENT = -2
HV_ERR_CUDA = -3
HV_ERR_CAPACITY = -4
HV_ERR_NO_DEVICE = -5
HV_ERR_UNSUPPORTED = -6
HV_ERR_CHANNELS = -7
HV_ERR_BAD_TICKET = -8

HV_FLAG_PROFILE = 2
HV_FLAG_KEEP_BLUR = 4
HV_FLAG_DEFER_TAIL = 64
HV_FLAG_FORCE_GENERIC = 8
HV_FLAG_GLOBAL_CCL = 16
HV_FLAG_PHASE_TIMING = 32

HV_BLUR_BOX, HV_BLUR_GAUSSIAN, HV_BLUR_NONE = 0, 1, 2
HV_ALLOC_COMPRESSIBLE = 1
HV_FRAMESET_ZERO_COPY = 1
HV_PIPELINE_BASIC, HV_PIPELINE_CONTAMINATION = 0, 1
HV_STATS_AREA_BINS = 16
HV_K_COUNT = 9
(HV_K_GRAY, HV_K_PREPROCESS, HV_K_MORPH, HV_K_CCL_MERGE, HV_K_CCL_FLATTEN, HV_K_CCL_SCAN, HV_K_CCL_LABEL,
 HV_K_SCORE, HV_K_CCL_FRAME) = range(9)


class hv_config(C.Structure):
    _fields_ = [("max_batch", C.c_int32), ("max_height", C.c_int32), ("max_width", C.c_int32),
                ("max_blobs_per_frame", C.c_int32), ("max_defects_per_frame", C.c_int32), ("num_slots", C.c_int32),
                ("flags", C.c_int32), ("reserved", C.c_int32)]


class hv_params(C.Structure):
    _fields_ = [("min_size", C.c_double), ("max_size", C.c_double), ("threshold", C.c_double),
                ("min_confidence", C.c_double), ("gauss_sigma", C.c_double), ("blur_mode", C.c_int32),
                ("blur_ksize", C.c_int32), ("morph_open_k", C.c_int32), ("morph_close_k", C.c_int32),
                ("reserved", C.c_int32 * 4)]


class hv_defect(C.Structure):
    _fields_ = [("y", C.c_int32), ("x", C.c_int32), ("size", C.c_double), ("confidence", C.c_double),
                ("ymin", C.c_int32), ("xmin", C.c_int32), ("ymax", C.c_int32), ("xmax", C.c_int32),
                ("label", C.c_uint32), ("frame", C.c_uint32)]


class hv_frame_result(C.Structure):
    _fields_ = [("n_components", C.c_uint32), ("n_defects", C.c_uint32), ("defects_offset", C.c_uint32),
                ("rejected", C.c_uint32), ("fg_pixels", C.c_uint32), ("status", C.c_int32)]


class hv_blob(C.Structure):
    _fields_ = [("area", C.c_uint32), ("ymin", C.c_uint32), ("ymax", C.c_uint32), ("xmin", C.c_uint32),
                ("xmax", C.c_uint32), ("reserved", C.c_uint32), ("sum_y", C.c_uint64), ("sum_x", C.c_uint64)]


class hv_debug_outputs(C.Structure):
    _fields_ = [("gray", C.c_void_p), ("blur", C.c_void_p), ("mask", C.c_void_p), ("labels", C.c_void_p),
                ("blobs", C.c_void_p), ("blobs_stride", C.c_size_t)]


class hv_line_stats(C.Structure):
    _fields_ = [("frames_inspected", C.c_uint64), ("frames_rejected", C.c_uint64), ("total_defects", C.c_uint64),
                ("total_components", C.c_uint64), ("total_defect_area", C.c_uint64), ("total_fg_pixels", C.c_uint64),
                ("area_hist", C.c_uint64 * HV_STATS_AREA_BINS), ("capacity_errors", C.c_uint64),
                ("reserved", C.c_uint64 * 9)]


class hv_contour(C.Structure):
    _fields_ = [("y", C.c_int32), ("x", C.c_int32), ("area", C.c_double), ("pixel_count", C.c_uint64),
                ("label", C.c_uint32), ("reserved", C.c_uint32)]


class hv_center(C.Structure):
    _fields_ = [("y", C.c_int32), ("x", C.c_int32), ("confidence", C.c_double)]


(HV_PIX_MONO8, HV_PIX_MONO16, HV_PIX_RGB8, HV_PIX_BGR8, HV_PIX_RGBA8, HV_PIX_BGRA8, HV_PIX_YUV422,
 HV_PIX_YUV422_PACKED, HV_PIX_BAYER_RG8, HV_PIX_BAYER_GB8, HV_PIX_BAYER_GR8, HV_PIX_BAYER_BG8) = range(12)


class hv_camera_frame(C.Structure):
    _fields_ = [("data", C.c_void_p), ("size", C.c_size_t), ("width", C.c_uint32), ("height", C.c_uint32),
                ("pixel_format", C.c_int32), ("camera", C.c_uint32), ("frame_id", C.c_uint64),
                ("timestamp_ns", C.c_uint64)]


HV_SYNC_FREERUN, HV_SYNC_SOFTWARE, HV_SYNC_HARDWARE = 0, 1, 2


class hv_frameset_config(C.Structure):
    _fields_ = [("n_cameras", C.c_int32), ("sets_per_batch", C.c_int32), ("sync_mode", C.c_int32),
                ("max_pending_sets", C.c_int32), ("flags", C.c_int32), ("reserved", C.c_int32)]


class hv_frameset_stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("frames_pushed", "sets_completed", "sets_dropped", "frames_dropped",
                                          "duplicates", "batches_submitted", "max_skew_ns", "sets_pending")]


class hv_pydet_params(C.Structure):
    _fields_ = [("contrast_threshold", C.c_double), ("blur_ksize", C.c_int32), ("block_size", C.c_int32),
                ("morph_open_k", C.c_int32), ("morph_close_k", C.c_int32), ("reserved", C.c_int32 * 4)]


class hv_pydet_score_params(C.Structure):
    _fields_ = [("min_size", C.c_double), ("max_size", C.c_double), ("min_confidence", C.c_double),
                ("use_color", C.c_int32), ("reserved", C.c_int32 * 3)]


class hv_pydefect(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("size", C.c_double), ("confidence", C.c_double),
                ("intensity_diff", C.c_double), ("shape_score", C.c_double), ("color_score", C.c_double),
                ("bx", C.c_int32), ("by", C.c_int32), ("bw", C.c_int32), ("bh", C.c_int32),
                ("label8", C.c_uint32), ("chain_len", C.c_uint32)]


class hv_inspection_record(C.Structure):
    _fields_ = [("sequence", C.c_uint64), ("timestamp", C.c_double), ("processing_time", C.c_double),
                ("success", C.c_uint32), ("has_defects", C.c_uint32), ("defect_count", C.c_uint32),
                ("defects_offset", C.c_uint32)]


class hv_dashboard_stats(C.Structure):
    _fields_ = [("total_images", C.c_uint64), ("total_defects", C.c_uint64), ("avg_processing_time_ms", C.c_double),
                ("defect_rate", C.c_double), ("start_time", C.c_double)]


class hv_overlay(C.Structure):
    _fields_ = [("kind", C.c_int32), ("y", C.c_int32), ("x", C.c_int32), ("y1", C.c_int32), ("x1", C.c_int32),
                ("color", C.c_uint8 * 3), ("reserved", C.c_uint8)]


HV_OVERLAY_CROSS, HV_OVERLAY_BOX, HV_OVERLAY_MARKER = 0, 1, 2

assert C.sizeof(hv_camera_frame) == 48
assert C.sizeof(hv_defect) == 48 and C.sizeof(hv_frame_result) == 24 and C.sizeof(hv_blob) == 40
assert C.sizeof(hv_line_stats) == 256

_vp, _sz, _i32, _i64, _f64 = C.c_void_p, C.c_size_t, C.c_int32, C.c_int64, C.c_double
_P = C.POINTER

# name -> (restype, argtypes); must list every HV_API symbol of include/heimdall_cuda.h (checked by the tests)
PROTOTYPES = {
    "hv_abi_version": (_i32, []),
    "hv_version": (C.c_char_p, []),
    "hv_status_string": (C.c_char_p, [_i32]),
    "hv_device_count": (_i32, []),
    "hv_params_default": (None, [_P(hv_params)]),
    "hv_config_default": (None, [_P(hv_config)]),
    "hv_create": (_i32, [_i32, _P(hv_config), _P(_vp)]),
    "hv_destroy": (None, [_vp]),
    "hv_last_error": (C.c_char_p, [_vp]),
    "hv_set_stream": (_i32, [_vp, _vp, _i32]),
    "hv_host_alloc": (_vp, [_vp, _sz]),
    "hv_host_free": (None, [_vp, _vp]),
    "hv_pipeline_depth": (_i32, []),
    "hv_device_alloc": (_i32, [_vp, _sz, C.c_uint32, _P(_vp), _P(_i32)]),
    "hv_device_free": (_i32, [_vp, _vp]),
    "hv_device_read": (_i32, [_vp, _vp, _vp, _sz]),
    "hv_device_write": (_i32, [_vp, _vp, _vp, _sz]),
    "hv_detect_batch": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _sz, _sz, _P(hv_params), _P(hv_frame_result),
                               _P(hv_defect), _sz, _P(_sz), _P(hv_debug_outputs)]),
    "hv_detect_batch_device": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _sz, _sz, _P(hv_params), _vp, _vp,
                                      _P(hv_frame_result), _P(hv_defect), _sz, _P(_sz)]),
    "hv_enqueue_device": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _sz, _sz, _P(hv_params), _vp, _vp, _P(_i64)]),
    "hv_fetch_ticket": (_i32, [_vp, _i64, _P(hv_frame_result), _P(hv_defect), _sz, _P(_sz)]),
    "hv_flush": (_i32, [_vp]),
    "hv_fetch_results": (_i32, [_vp, _P(hv_frame_result), _P(hv_defect), _sz, _P(_sz)]),
    "hv_fetch_debug": (_i32, [_vp, _P(hv_debug_outputs)]),
    "hv_submit": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _sz, _sz, _P(hv_params), _P(_i64)]),
    "hv_wait": (_i32, [_vp, _i64, _P(hv_frame_result), _P(hv_defect), _sz, _P(_sz)]),
    "hv_frame_channels": (_i32, [_i32]),
    "hv_convert_frame": (_i32, [_vp, _P(hv_camera_frame), _vp, _P(_i32)]),
    "hv_submit_frames": (_i32, [_vp, _P(hv_camera_frame), _i32, _P(hv_params), _P(_i64)]),
    "hv_frameset_create": (_i32, [_vp, _P(hv_frameset_config), _P(_vp)]),
    "hv_frameset_destroy": (None, [_vp]),
    "hv_frameset_last_error": (C.c_char_p, [_vp]),
    "hv_frameset_push": (_i32, [_vp, _P(hv_camera_frame), _P(hv_params), _P(_i64)]),
    "hv_frameset_flush": (_i32, [_vp, _P(hv_params), _P(_i64)]),
    "hv_frameset_batch_ids": (_i32, [_vp, _i64, _P(C.c_uint64), _i32, _P(_i32)]),
    "hv_frameset_wait": (_i32, [_vp, _i64, _P(hv_frame_result), _P(hv_defect), _sz, _P(_sz)]),
    "hv_frameset_get_stats": (_i32, [_vp, _P(hv_frameset_stats)]),
    "hv_preprocess_image": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "hv_apply_threshold": (_i32, [_vp, _vp, _i32, _i32, _i32, C.c_uint8, _i32, _i32, _vp]),
    "hv_morphology": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "hv_find_contours": (_i32, [_vp, _vp, _i32, _i32, _i32, _f64, _f64, _P(hv_contour), _sz, _P(_sz), _vp]),
    "hv_process_image": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _P(hv_center), _sz, _P(_sz)]),
    "hv_pydet_params_default": (None, [_P(hv_pydet_params)]),
    "hv_python_detector_stages": (_i32, [_vp, _vp, _i32, _i32, _i32, _P(hv_pydet_params), _vp, _vp, _vp, _vp, _P(hv_blob), _sz,
                                         _P(_sz)]),
    "hv_pydet_score_params_default": (None, [_P(hv_pydet_score_params)]),
    "hv_python_detect": (_i32, [_vp, _vp, _i32, _i32, _i32, _P(hv_pydet_params), _P(hv_pydet_score_params), _P(hv_pydefect), _sz,
                                _P(_sz), _P(_sz)]),
    "hv_export_results": (_i32, [_P(hv_frame_result), _i32, _f64, _f64, C.c_uint64, _P(hv_inspection_record),
                                 _P(hv_dashboard_stats)]),
    "hv_draw_overlays": (_i32, [_vp, _vp, _i32, _i32, _P(hv_overlay), _i32]),
    "hv_stats_get": (_i32, [_vp, _P(hv_line_stats)]),
    "hv_stats_reset": (_i32, [_vp]),
    "hv_stats_device_ptr": (_vp, [_vp]),
    "hv_launch_count": (C.c_uint64, [_vp]),
    "hv_profile_enable": (_i32, [_vp, C.c_uint32]),
    "hv_profile_get": (_i32, [_vp, _P(C.c_float), _P(C.c_uint32)]),
    "hv_kernel_name": (C.c_char_p, [_i32]),
    "hv_debug_phase_times": (_i32, [_vp, _P(C.c_uint64)]),
}


def load(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise ImportError(
            f"heimdall_core: CUDA backend library not found at {path}; build it with `python __graft_entry__.py` "
            "(or `make -C heimdall-vision_b200`). There is no CPU fallback.")
    try:
        lib = C.CDLL(path)
    except OSError as e:  # pragma: no cover
        raise ImportError(f"heimdall_core: cannot load {path}: {e}") from e
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here = ABI mismatch, fail loudly
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load()
