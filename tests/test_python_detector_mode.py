"""Next-row N3: the stages of the reference's Python detector (heimdall/detectors/contamination_detector.py:58-90) on the GPU
against outputs of opencv-python itself (tests/golden/cv2_python_detector.npz, generator: tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

CASES = ["bottle_bgr", "texture_bgr", "blemish_gray", "colour_bgr"]


def _unpack(a, h, w):
    return (np.unpackbits(a)[:h * w].reshape(h, w) * 255).astype(np.uint8)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_stages_match_opencv(detector, golden_dir, case):
    z = np.load(os.path.join(golden_dir, "cv2_python_detector.npz"))
    img, C = z[f"{case}_img"], float(z[f"{case}_C"])
    h, w = img.shape[:2]
    out = detector.python_detector_stages(img, C)
    assert np.array_equal(out["gray"], z[f"{case}_gray"]), "cv2.cvtColor(BGR2GRAY)"
    assert np.array_equal(out["blurred"], z[f"{case}_blurred"]), "cv2.GaussianBlur(5,5,0)"
    assert np.array_equal(out["binary"], _unpack(z[f"{case}_morphed"], h, w)), "adaptiveThreshold(GAUSSIAN_C) + open + close"
    # the threshold stage alone (no morphology)
    pre = detector.python_detector_stages(img, C, morph_open_k=0, morph_close_k=0)["binary"]
    assert np.array_equal(pre, _unpack(z[f"{case}_binary"], h, w)), "cv2.adaptiveThreshold(GAUSSIAN_C, BINARY_INV, 11, C)"
    # 8-connected components in raster order of their first pixel: label plane, pixel areas, bounding boxes
    assert np.array_equal(out["labels8"], z[f"{case}_labels8"])
    st = z[f"{case}_stats8"]                                   # x, y, w, h, area
    comps = out["components"]
    assert len(comps) == len(st)
    assert np.array_equal(comps["area"], st[:, 4])
    assert np.array_equal(comps["xmin"], st[:, 0]) and np.array_equal(comps["ymin"], st[:, 1])
    assert np.array_equal(comps["xmax"] - comps["xmin"] + 1, st[:, 2]) and np.array_equal(comps["ymax"] - comps["ymin"] + 1, st[:, 3])
    # the outer borders cv2.findContours(RETR_EXTERNAL) traces: one per component here (no component inside another's hole)
    rects = sorted((int(c["xmin"]), int(c["ymin"]), int(c["xmax"] - c["xmin"] + 1), int(c["ymax"] - c["ymin"] + 1)) for c in comps)
    assert rects == [tuple(int(v) for v in r) for r in z[f"{case}_ext_rects"]]


@pytest.mark.gpu
def test_reference_python_results_are_consistent(detector, golden_dir):
    """tests/golden/reference_python.json holds what the UNMODIFIED reference fallback returns for three bottle frames (no
    defects: after GAUSSIAN_C thresholding with C = 25 and open + close nothing is left).  The GPU stages agree: the final
    mask of those frames has no component inside the detector's size limits."""
    import synth
    ref = json.load(open(os.path.join(golden_dir, "reference_python.json")))["results"]
    for idx, defects in ref.items():
        fr = synth.bottle_frame(240, 320, int(idx), contaminants=2)
        out = detector.python_detector_stages(np.dstack([fr, fr, fr]), 25.0)
        big = [c for c in out["components"] if 10 <= int(c["area"]) <= 3000]
        assert len(defects) == 0 and len(big) == 0


def test_eight_connectivity_differs_from_four_on_diagonals():
    """Host-side sanity of the test vectors themselves: the golden label planes are 8-connected (a diagonal pair is one
    component), unlike the Rust path's 4-connected labels."""
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cv2_python_detector.npz")
    z = np.load(golden)
    lab = z["texture_bgr_labels8"]
    fg = lab > 0
    diag = fg[1:, 1:] & fg[:-1, :-1] & ~fg[1:, :-1] & ~fg[:-1, 1:]
    if diag.any():
        ys, xs = np.nonzero(diag)
        assert (lab[ys + 1, xs + 1] == lab[ys, xs]).all()


REF_CASES = ["blemish_gray", "blemish_bgr", "colour_bgr", "colour_bgr_nocolor", "texture_bgr", "nested_gray"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", REF_CASES)
def test_python_detect_equals_the_unmodified_reference_detector(detector, golden_dir, case):
    """hv_python_detect against tests/golden/reference_python_detect.json: what the reference's own
    `ContaminationDetector.detect` (imported unmodified from /root/reference by tests/golden/make_golden.py) returned for
    the same frames and configurations -- every defect, in its order, with exact equality of the integer fields and of
    the doubles (contour area, confidence, the three scores).  Covers gray and BGR input, the colour score on and off,
    non-default size limits, nested shapes (a ring with an island in its hole: RETR_EXTERNAL keeps the ring only)."""
    ref = json.load(open(os.path.join(golden_dir, "reference_python_detect.json")))[case]
    img = np.load(os.path.join(golden_dir, "reference_python_detect_frames.npz"))[case]
    cfg = ref["config"]
    got = detector.python_detect(img, cfg.get("min_contaminant_size", 10), cfg.get("max_contaminant_size", 3000),
                                 cfg.get("contrast_threshold", 15), cfg.get("min_confidence", 0.25), cfg.get("use_color", True))
    exp = ref["defects"]
    assert len(exp) > 0
    assert [list(d["position"]) for d in got] == [d["position"] for d in exp]
    assert [list(d["bounding_box"]) for d in got] == [d["bounding_box"] for d in exp]
    for key in ("size", "confidence", "intensity_diff", "shape_score", "color_score"):
        assert [d[key] for d in got] == [d[key] for d in exp], key


def test_detector_mirror_has_the_reference_contract():
    """heimdall_core.detectors.ContaminationDetector: constructor defaults of contamination_detector.py:26-38 and the
    detect() / __call__ contract of heimdall/detectors/base.py:41-84 (no GPU needed for this part)."""
    import inspect

    from heimdall_core import detectors as D
    d = D.ContaminationDetector()
    assert (d.name, d.min_contaminant_size, d.max_contaminant_size, d.contrast_threshold, d.min_confidence, d.use_color) == \
        ("contamination_detector", 10, 3000, 15, 0.25, True)
    d2 = D.ContaminationDetector(config={"contrast_threshold": 25, "min_confidence": 0.3})   # what rust_bridge.py:143-149 passes
    assert (d2.contrast_threshold, d2.min_confidence, d2.min_contaminant_size) == (25, 0.3, 10)
    assert list(inspect.signature(d.detect).parameters) == ["image", "context"]
    assert issubclass(D.ContaminationDetector, D.DefectDetector) and callable(d)
    with pytest.raises(ValueError):
        d.detect(np.zeros((4, 4, 2), np.uint8))


@pytest.mark.gpu
def test_detector_mirror_returns_the_reference_defects(detector, golden_dir):
    from heimdall_core import detectors as D
    ref = json.load(open(os.path.join(golden_dir, "reference_python_detect.json")))["blemish_bgr"]
    img = np.load(os.path.join(golden_dir, "reference_python_detect_frames.npz"))["blemish_bgr"]
    got = D.ContaminationDetector(config=ref["config"], detector=detector)(img)
    assert [(list(d.position), d.size, d.confidence, d.defect_type) for d in got] == \
        [(e["position"], e["size"], e["confidence"], "contamination") for e in ref["defects"]]
    dd = got[0].to_dict()
    assert set(dd) == {"type", "position", "size", "confidence", "intensity_diff", "shape_score", "color_score", "bounding_box"}
