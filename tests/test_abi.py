"""CPU-only checks of the drop-in boundary: the shared library loads, exports exactly the symbols include/heimdall_cuda.h
declares, struct layouts agree between the header, the ctypes binding and numpy, and the product path fails loudly
without a GPU (no CPU fallback, no oracle import)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "heimdall_cuda.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"HV_API\s+[\w\s\*]+?\b(hv_\w+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    syms = header_symbols()
    for must in ("hv_create", "hv_destroy", "hv_detect_batch", "hv_detect_batch_device", "hv_enqueue_device",
                 "hv_fetch_results", "hv_submit", "hv_wait", "hv_preprocess_image", "hv_apply_threshold",
                 "hv_find_contours", "hv_process_image", "hv_stats_get", "hv_stats_device_ptr", "hv_last_error",
                 "hv_version", "hv_launch_count", "hv_profile_get", "hv_profile_enable"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    import heimdall_core._abi as A
    syms = header_symbols()
    assert sorted(A.PROTOTYPES) == syms, "ctypes PROTOTYPES and the header disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", A.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert set(syms) <= exported
    # nothing but the hv_ ABI leaks out of the library (kernels and helpers have hidden visibility)
    leaked = {s for s in exported if not s.startswith("hv_") and not s.startswith("_")}
    assert not leaked, leaked
    for s in syms:
        assert hasattr(A.lib, s)


def test_version_status_strings_and_defaults():
    import heimdall_core._abi as A
    assert A.lib.hv_abi_version() == 1
    assert b"sm_100a" in A.lib.hv_version()
    assert A.lib.hv_status_string(A.HV_ERR_INVALID_DIMENSIONS) == b"Invalid image dimensions: expected 3D array"
    p = A.hv_params()
    A.lib.hv_params_default(C.byref(p))
    # rust/heimdall-core/src/lib.rs:106-108, detection.rs:163,298
    assert (p.min_size, p.max_size, p.threshold, p.min_confidence) == (10.0, 3000.0, 25.0, 0.3)
    assert (p.blur_mode, p.blur_ksize, p.morph_open_k, p.morph_close_k) == (A.HV_BLUR_BOX, 5, 0, 0)
    assert A.lib.hv_kernel_name(1) == b"preprocess_mask"


def test_struct_layouts_match_the_header():
    """Compile a tiny C program against the public header and compare sizeof/offsetof with ctypes."""
    import heimdall_core._abi as A
    import tempfile
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "heimdall_cuda.h"
int main(void){
 printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(hv_config), sizeof(hv_params), sizeof(hv_defect),
   sizeof(hv_frame_result), sizeof(hv_blob), sizeof(hv_debug_outputs), sizeof(hv_line_stats), sizeof(hv_contour), sizeof(hv_center));
 printf("%zu %zu %zu %zu %zu\n", offsetof(hv_params, blur_mode), offsetof(hv_defect, confidence), offsetof(hv_defect, label),
   offsetof(hv_blob, sum_y), offsetof(hv_line_stats, capacity_errors));
 return 0; }
'''
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(td, "t")
        subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        l1, l2 = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines()
    sizes = [C.sizeof(t) for t in (A.hv_config, A.hv_params, A.hv_defect, A.hv_frame_result, A.hv_blob,
                                   A.hv_debug_outputs, A.hv_line_stats, A.hv_contour, A.hv_center)]
    assert list(map(int, l1.split())) == sizes
    offs = [A.hv_params.blur_mode.offset, A.hv_defect.confidence.offset, A.hv_defect.label.offset,
            A.hv_blob.sum_y.offset, A.hv_line_stats.capacity_errors.offset]
    assert list(map(int, l2.split())) == offs
    from heimdall_core import batch
    assert batch.DEFECT_DTYPE.fields["confidence"][1] == A.hv_defect.confidence.offset
    assert batch.BLOB_DTYPE.fields["sum_x"][1] == A.hv_blob.sum_x.offset


def test_module_surface_matches_lib_rs():
    """rust/heimdall-core/src/lib.rs:20-36: three functions + three submodules with these function names."""
    import inspect

    import heimdall_core as hc
    assert list(inspect.signature(hc.process_image).parameters) == ["image", "pipeline_type", "params"]
    assert list(inspect.signature(hc.detect_contamination).parameters) == ["image", "min_size", "max_size", "threshold"]
    assert list(inspect.signature(hc.benchmark_processing).parameters) == ["image", "iterations"]
    assert list(inspect.signature(hc.acquisition.acquire_image).parameters) == ["source_type", "params"]
    assert list(inspect.signature(hc.processing.preprocess_image).parameters) == ["image", "grayscale", "blur_size"]
    assert list(inspect.signature(hc.processing.apply_threshold).parameters) == ["image", "threshold_value", "adaptive",
                                                                                 "inverse"]
    assert list(inspect.signature(hc.detection.find_contours).parameters) == ["image", "min_area", "max_area"]


def test_argument_validation_happens_before_any_gpu_work():
    import heimdall_core as hc
    with pytest.raises(TypeError):
        hc.detect_contamination(np.zeros((4, 4), np.uint8))
    with pytest.raises(TypeError):
        hc.detect_contamination([[1, 2], [3, 4]])
    with pytest.raises(TypeError):
        hc.process_image(np.zeros((4, 4, 3), np.float64), "basic")
    with pytest.raises(ValueError, match="Unsupported pipeline type: fancy"):
        hc.process_image(np.zeros((4, 4, 3), np.uint8), "fancy")
    with pytest.raises(ValueError, match="Unsupported source type: usb"):
        hc.acquisition.acquire_image("usb")
    img = hc.acquisition.acquire_image("camera")
    assert img.shape == (480, 640, 3) and img.dtype == np.uint8
    # acquisition.rs:57-107: outline value 100 on the rectangle border, disc of 80, 220 elsewhere
    assert img[120, 300, 0] == 100 and img[340, 320, 1] == 80 and img[10, 10, 2] == 220


def test_no_cpu_fallback_and_no_oracle_on_the_product_path():
    """Without a CUDA device every compute call must raise; and nothing under heimdall-vision_b200/ may reference the
    oracle."""
    import heimdall_core as hc
    if hc._abi.lib.hv_device_count() == 0:
        with pytest.raises(hc.HeimdallCudaError) as ei:
            hc.Detector(0)
        assert ei.value.status == hc._abi.HV_ERR_NO_DEVICE and "no CPU fallback" in str(ei.value)
        with pytest.raises(hc.HeimdallCudaError):
            hc.detect_contamination(np.zeros((16, 16, 1), np.uint8))
    pkg = os.path.join(ROOT, "heimdall-vision_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "hv_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_reference_bridge_picks_the_module_up(monkeypatch):
    """`import heimdall_core` is all the reference bridge does (heimdall/rust_bridge.py:20-26); mimic its probe."""
    import importlib
    m = importlib.import_module("heimdall_core")
    assert all(hasattr(m, n) for n in ("process_image", "detect_contamination", "benchmark_processing"))
