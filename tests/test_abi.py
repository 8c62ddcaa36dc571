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
    assert A.lib.hv_abi_version() == 2
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


def test_rust_crate_is_generated_from_the_header():
    """rust/heimdall-cuda/src/ffi.rs (constants, #[repr(C)] structs, the extern "C" block) and build.rs (the .cu source
    list) are generated by tools/gen_rust_ffi.py from include/heimdall_cuda.h and the Makefile; the committed files must be
    what the generator produces now, every HV_API symbol must have its `pub fn`, every header struct its `pub struct` with
    the same field names in the same order, and build.rs must list exactly the Makefile's sources."""
    gen = os.path.join(ROOT, "tools", "gen_rust_ffi.py")
    r = subprocess.run(["python", gen, "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    ffi = open(os.path.join(ROOT, "rust", "heimdall-cuda", "src", "ffi.rs")).read()
    assert sorted(re.findall(r"pub fn (hv_\w+)\(", ffi)) == header_symbols()
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for body, name in re.findall(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", hdr, flags=re.S):
        c_fields = []
        for decl in body.split(";"):
            for part in decl.split(","):
                m = re.search(r"([A-Za-z_]\w*)\s*(\[[^\]]*\])?\s*$", part.strip())
                if m and part.strip():
                    c_fields.append(m.group(1))
        m = re.search(r"pub struct %s \{(.*?)\n\}" % name, ffi, flags=re.S)
        assert m, name
        assert re.findall(r"pub (\w+):", m.group(1)) == c_fields, name
    mk = open(os.path.join(ROOT, "heimdall-vision_b200", "Makefile")).read()
    srcs = [os.path.basename(s) for s in re.search(r"^SRCS\s*:=\s*(.*)$", mk, flags=re.M).group(1).split()]
    build_rs = open(os.path.join(ROOT, "rust", "heimdall-cuda", "build.rs")).read()
    listed = re.findall(r'"(\w+\.cu)"', build_rs)
    assert listed == srcs and set(srcs) == {f for f in os.listdir(os.path.join(ROOT, "heimdall-vision_b200", "csrc")) if f.endswith(".cu")}
    lib_rs = open(os.path.join(ROOT, "rust", "heimdall-cuda", "src", "lib.rs")).read()
    for used in set(re.findall(r"\b(hv_[a-z_]+)\(", lib_rs)):
        assert used in header_symbols(), used


REFERENCE = os.environ.get("HEIMDALL_REFERENCE_ROOT", "/root/reference")

_BRIDGE_PROBE = r'''
import json, sys
sys.path.insert(0, %(pkg)r)
sys.path.insert(0, %(ref)r)
import logging
logging.disable(logging.CRITICAL)
import numpy as np
import heimdall_core
calls = []
real = heimdall_core.detect_contamination
def spy(image, min_size=None, max_size=None, threshold=None):
    calls.append((image.shape, min_size, max_size, threshold))
    if %(stub)r:
        return {"defects": [{"position": (1, 2), "size": 3.0, "confidence": 0.5, "metadata": {}}], "processing_time": 0.0}
    return real(image, min_size, max_size, threshold)
heimdall_core.detect_contamination = spy
import heimdall.rust_bridge as rb                      # the UNMODIFIED reference bridge
img = np.full((64, 96, 1), 220, np.uint8)
img[20:30, 40:52] = 30
out = rb.RustBridge.detect_contamination(img, 10.0, 3000.0, 25.0)
print(json.dumps({"available": rb.RUST_AVAILABLE, "is_available": rb.RustBridge.is_available(),
                  "module": heimdall_core.__file__, "calls": [[list(c[0])] + list(c[1:]) for c in calls],
                  "defects": [[list(d["position"]), d["size"], d["confidence"]] for d in out["defects"]],
                  "keys": sorted(out)}))
'''


def _run_bridge(stub: bool):
    import json
    code = _BRIDGE_PROBE % {"pkg": os.path.join(ROOT, "heimdall-vision_b200"), "ref": REFERENCE, "stub": stub}
    r = subprocess.run(["python", "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "heimdall")), reason="reference checkout not present")
def test_unmodified_reference_bridge_routes_to_this_module():
    """heimdall/rust_bridge.py:20-26,118-137 of the reference, imported UNMODIFIED with this package on sys.path:
    `import heimdall_core` succeeds, RUST_AVAILABLE flips to True, and RustBridge.detect_contamination hands the frame and
    the three scalars to heimdall_core.detect_contamination and returns its dict untouched.  (CPU box: the call itself
    is answered by a stub here; with a GPU the same probe runs the real kernels, see the gpu test below.)"""
    out = _run_bridge(stub=True)
    assert out["available"] is True and out["is_available"] is True
    assert out["module"].startswith(os.path.join(ROOT, "heimdall-vision_b200"))
    assert out["calls"] == [[[64, 96, 1], 10.0, 3000.0, 25.0]]
    assert out["defects"] == [[[1, 2], 3.0, 0.5]] and out["keys"] == ["defects", "processing_time"]


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "heimdall")),
                    reason="reference checkout not present (it does not travel to the GPU box)")
def test_unmodified_reference_bridge_on_the_gpu(oracle):
    """The same probe with the real kernels behind it: what the reference's dashboard / benchmark would get."""
    out = _run_bridge(stub=False)
    img = np.full((64, 96, 1), 220, np.uint8)
    img[20:30, 40:52] = 30
    ref = oracle.detect_contamination(img)
    assert out["available"] is True and out["defects"] == [[list(d["position"]), d["size"], d["confidence"]] for d in ref.defects]
