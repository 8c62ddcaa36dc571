#!/usr/bin/env python3
"""Regenerates the committed golden fixtures under tests/golden/.  Run in the dev container only (it needs
opencv-python and, for the reference-derived vectors, /root/reference); the tests read the committed files and never
touch /root/reference.

Sources of truth:
  cv2_gaussian.npz / cv2_morph.npz / cv2_ccl.npz / cv2_pixfmt.npz : outputs of opencv-python (the reference's real third-party
      dependency for its Python path; version printed into meta.json) on seeded inputs.
  reference_python.json : outputs of the UNMODIFIED reference Python code imported from /root/reference
      (heimdall/rust_bridge.py RustBridge.detect_contamination -> heimdall/detectors/contamination_detector.py) on
      seeded synthetic frames; pins the fallback path.
  fixture_frames.npz + rust_path.json : the reference's own fixture images contaminated_{1,2,3}.jpg decoded with this
      cv2 build, and the Rust-path results of the oracle on them (regression vectors; the reference ships no expected
      outputs for them -- parity of the Rust path itself is pinned by the hand-derived KATs in tests/test_oracle.py).
  bottle_expectations.json : oracle results on the synthetic bottle frames used by the GPU parity tests and bench.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import importlib.util  # noqa: E402

import cv2  # noqa: E402

# load synth.py by path: the package directory must NOT be on sys.path here, or the reference bridge below would pick
# up our drop-in `heimdall_core` instead of its own Python fallback
_spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "heimdall-vision_b200", "synth.py"))
synth = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(synth)
from oracle import oracle as O  # noqa: E402

GAUSS_CASES = [(3, 0), (5, 0), (7, 0), (9, 0), (11, 0), (13, 0), (15, 0), (7, 1.0), (13, 2.0), (15, 3.0), (5, 1.1),
               (9, 1.5), (15, 2.5), (3, 0.8)]
MORPH_KS = [2, 3, 4, 5, 7, 9, 11, 13, 15]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def pixfmt():
    """cv2_pixfmt.npz: camera pixel-format conversions named by rust/heimdall-camera/src/lib.rs:226-252 (next-row N1)."""
    rng = np.random.default_rng(20261019)
    d = {}
    bayer_a = rng.integers(0, 256, (37, 50), dtype=np.uint8)
    bayer_b = synth.bottle_frame(96, 128, 11, contaminants=2)  # a structured mosaic
    d["bayer_a"], d["bayer_b"] = bayer_a, bayer_b
    for pat in ("RG", "GB", "GR", "BG"):
        code = getattr(cv2, f"COLOR_Bayer{pat}2RGB")
        d[f"a_{pat}"] = cv2.cvtColor(bayer_a, code)
        d[f"b_{pat}"] = cv2.cvtColor(bayer_b, code)
    yuyv = rng.integers(0, 256, (24, 34, 2), dtype=np.uint8)
    d["yuyv"] = yuyv
    d["yuyv_rgb"] = cv2.cvtColor(yuyv, cv2.COLOR_YUV2RGB_YUYV)
    np.savez_compressed(os.path.join(HERE, "cv2_pixfmt.npz"), **d)


MORPH_PIPE_CASES = [("240x333_11", (240, 333, 11, {"contaminants": 3})), ("1024x1280_1", (1024, 1280, 1, {"contaminants": 2})),
                    ("96x128_7", (96, 128, 7, {"contaminants": 2}))]
MORPH_PIPE_KS = [(3, 3), (3, 0), (0, 5), (5, 7), (9, 3), (15, 15)]


def morph_pipeline():
    """cv2_morph_pipeline.npz: cv2.morphologyEx(MORPH_OPEN k_open) then (MORPH_CLOSE k_close), MORPH_RECT, applied by
    opencv-python to the reference-path mask of synthetic bottle frames (heimdall/detectors/contamination_detector.py:81-87
    is the call sequence).  The GPU test runs the same frames through the detector WITH morphology and must reproduce
    these masks bit for bit: the fused morphology kernels are pinned to OpenCV itself, not to the oracle's restatement
    (the pre-morphology mask both sides start from is checked separately).  Masks are stored bit-packed."""
    d = {}
    for name, (h, w, idx, kw) in MORPH_PIPE_CASES:
        fr = synth.bottle_frame(h, w, idx, **kw)
        pre = O.detect_contamination(fr).mask
        d[f"{name}_pre"] = np.packbits(pre > 0)
        for ko, kc in MORPH_PIPE_KS:
            m = pre
            if ko:
                m = cv2.morphologyEx(m, cv2.MORPH_OPEN, cv2.getStructuringElement(cv2.MORPH_RECT, (ko, ko)))
            if kc:
                m = cv2.morphologyEx(m, cv2.MORPH_CLOSE, cv2.getStructuringElement(cv2.MORPH_RECT, (kc, kc)))
            d[f"{name}_o{ko}_c{kc}"] = np.packbits(m > 0)
    np.savez_compressed(os.path.join(HERE, "cv2_morph_pipeline.npz"), **d)


def overlays():
    """cv2_overlays.npz: what opencv-python draws for the result-side overlays (next-row N4): cv2.rectangle thickness 1 for
    boxes (any corner order, partly outside the image), cv2.circle(c, 10, color, 2) for the dashboard's defect marker
    (dashboard.py:462) at centres whose marker lies inside the image.  Items are (kind, y, x, y1, x1, b, g, r)."""
    rng = np.random.default_rng(20261020)
    h, w = 96, 140
    base = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    items = []
    img = base.copy()
    for _ in range(14):
        x, y = int(rng.integers(-6, w + 6)), int(rng.integers(-6, h + 6))
        x1, y1 = x + int(rng.integers(-30, 40)), y + int(rng.integers(-30, 40))
        col = [int(c) for c in rng.integers(0, 256, 3)]
        cv2.rectangle(img, (x, y), (x1, y1), col, 1)
        items.append([1, y, x, y1, x1] + col)
    for _ in range(10):
        x, y = int(rng.integers(12, w - 12)), int(rng.integers(12, h - 12))
        col = [int(c) for c in rng.integers(0, 256, 3)]
        cv2.circle(img, (x, y), 10, col, 2)
        items.append([2, y, x, 0, 0] + col)
    np.savez_compressed(os.path.join(HERE, "cv2_overlays.npz"), base=base, items=np.array(items, np.int32), out=img)


def python_detector():
    """cv2_python_detector.npz (next-row N3): the stages of the reference's Python detector
    (heimdall/detectors/contamination_detector.py:58-90) run through opencv-python itself on seeded frames:
    cvtColor(BGR2GRAY) -> GaussianBlur(5,5,0) -> adaptiveThreshold(GAUSSIAN_C, BINARY_INV, 11, C) -> MORPH_OPEN 3x3 ->
    MORPH_CLOSE 3x3 -> findContours(RETR_EXTERNAL); plus the 8-connected components of the final mask
    (connectedComponentsWithStats, relabelled in raster order of each component's first pixel)."""
    rng = np.random.default_rng(20261021)
    frames = {}
    fr = synth.bottle_frame(240, 320, 3, contaminants=2)
    frames["bottle_bgr"] = (np.dstack([fr, fr, fr]), 25)
    tex = rng.integers(0, 256, (72, 100, 3), dtype=np.uint8)
    frames["texture_bgr"] = (tex, 10)
    blem = synth.near_threshold_frame(200, 260, 7, spots=8)
    frames["blemish_gray"] = (blem, 5)
    col = np.dstack([synth.bottle_frame(160, 208, 9, contaminants=3), synth.bottle_frame(160, 208, 10, contaminants=1),
                     synth.bottle_frame(160, 208, 11, contaminants=2)])
    frames["colour_bgr"] = (col, 12.7)
    d = {}
    for name, (img, C) in frames.items():
        gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) if img.ndim == 3 else img
        blurred = cv2.GaussianBlur(gray, (5, 5), 0)
        binary = cv2.adaptiveThreshold(blurred, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY_INV, 11, C)
        ker = cv2.getStructuringElement(cv2.MORPH_RECT, (3, 3))
        m = cv2.morphologyEx(cv2.morphologyEx(binary, cv2.MORPH_OPEN, ker), cv2.MORPH_CLOSE, ker)
        n, lab, st, _ = cv2.connectedComponentsWithStats(m, connectivity=8, ltype=cv2.CV_32S)
        flat = lab.ravel()
        _, first = np.unique(flat, return_index=True)          # first raster index of every label (0 = background)
        order = np.argsort(first[1:]) + 1                        # labels by first pixel
        remap = np.zeros(n, np.int32)
        remap[order] = np.arange(1, n)
        canon = remap[lab]
        stats = st[order]                                        # x, y, w, h, area in canonical order
        contours, _ = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        rects = sorted(cv2.boundingRect(c) for c in contours)
        d[f"{name}_img"] = img
        d[f"{name}_C"] = np.float64(C)
        d[f"{name}_gray"], d[f"{name}_blurred"] = gray, blurred
        d[f"{name}_binary"], d[f"{name}_morphed"] = np.packbits(binary > 0), np.packbits(m > 0)
        d[f"{name}_labels8"] = canon.astype(np.int32)
        d[f"{name}_stats8"] = stats.astype(np.int32)
        d[f"{name}_ext_rects"] = np.array(rects, np.int32).reshape(-1, 4)
    np.savez_compressed(os.path.join(HERE, "cv2_python_detector.npz"), **d)
    for name in frames:
        print(name, "components", len(d[f"{name}_stats8"]), "external contours", len(d[f"{name}_ext_rects"]))


def python_detect():
    """reference_python_detect.json (next-row N3): results of the UNMODIFIED reference detector
    (heimdall/detectors/contamination_detector.py `ContaminationDetector.detect`, imported from /root/reference) on seeded
    frames that do contain contamination it finds: every defect's to_dict() without the contour point list."""
    sys.path.insert(0, "/root/reference")
    import logging
    logging.disable(logging.CRITICAL)
    from heimdall.detectors.contamination_detector import ContaminationDetector
    rng = np.random.default_rng(20261022)
    frames = {}
    blem = synth.near_threshold_frame(200, 260, 7, spots=8)
    frames["blemish_gray"] = (blem, dict(contrast_threshold=5))
    frames["blemish_bgr"] = (np.dstack([blem, np.roll(blem, 3, 1), np.roll(blem, 5, 0)]), dict(contrast_threshold=5, min_confidence=0.3))
    col = np.dstack([synth.bottle_frame(160, 208, 9, contaminants=3), synth.bottle_frame(160, 208, 10, contaminants=1),
                     synth.bottle_frame(160, 208, 11, contaminants=2)])
    frames["colour_bgr"] = (col, dict(contrast_threshold=12.7))
    frames["colour_bgr_nocolor"] = (col, dict(contrast_threshold=12.7, use_color=False, min_contaminant_size=4, max_contaminant_size=500))
    tex = rng.integers(0, 256, (72, 100, 3), dtype=np.uint8)
    frames["texture_bgr"] = (tex, dict(contrast_threshold=10, min_contaminant_size=2))
    rings = np.full((120, 160), 200, np.uint8)   # nested shapes: a ring with an island in its hole, a blob with two holes
    cv2.circle(rings, (40, 60), 30, 60, 6)
    cv2.circle(rings, (40, 60), 8, 50, -1)
    cv2.rectangle(rings, (90, 30), (150, 90), 40, -1)
    cv2.rectangle(rings, (100, 40), (115, 55), 200, -1)
    cv2.circle(rings, (135, 72), 7, 210, -1)
    frames["nested_gray"] = (rings, dict(contrast_threshold=15, max_contaminant_size=100000))
    out = {}
    for name, (img, cfg) in frames.items():
        det = ContaminationDetector(config=cfg)
        defects = det.detect(img)
        recs = []
        for d in defects:
            dd = d.to_dict()
            recs.append({"position": [int(v) for v in dd["position"]], "size": float(dd["size"]), "confidence": float(dd["confidence"]),
                         "intensity_diff": float(dd["intensity_diff"]), "shape_score": float(dd["shape_score"]),
                         "color_score": float(dd["color_score"]), "bounding_box": [int(v) for v in dd["bounding_box"]]})
        out[name] = {"config": cfg, "defects": recs}
        print(name, len(recs), "defects")
    np.savez_compressed(os.path.join(HERE, "reference_python_detect_frames.npz"), **{k: v[0] for k, v in frames.items()})
    json.dump(out, open(os.path.join(HERE, "reference_python_detect.json"), "w"), indent=1)
    logging.disable(logging.NOTSET)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "python_detect":
        python_detect()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "python_detector":
        python_detector()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "overlays":
        overlays()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "pixfmt":
        pixfmt()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "morph_pipeline":
        morph_pipeline()
        return
    pixfmt()
    morph_pipeline()
    overlays()
    python_detector()
    if os.path.isdir("/root/reference/heimdall"):
        python_detect()
    rng = np.random.default_rng(20261018)
    meta = {"opencv": cv2.__version__, "numpy": np.__version__}

    # ---- Gaussian ------------------------------------------------------------------------------------------
    src = rng.integers(0, 256, (61, 83), dtype=np.uint8)
    src2 = synth.bottle_frame(96, 128, 7, contaminants=2)
    g = {"src": src, "src2": src2}
    for k, s in GAUSS_CASES:
        g[f"a_k{k}_s{s}"] = cv2.GaussianBlur(src, (k, k), s)
        g[f"b_k{k}_s{s}"] = cv2.GaussianBlur(src2, (k, k), s)
    np.savez_compressed(os.path.join(HERE, "cv2_gaussian.npz"), **g)

    # ---- morphology ------------------------------------------------------------------------------------------
    m1 = (rng.random((70, 101)) < 0.55).astype(np.uint8) * 255
    m2 = np.zeros((64, 96), np.uint8)
    cv2.circle(m2, (40, 30), 17, 255, -1)
    cv2.rectangle(m2, (60, 5), (90, 50), 255, 2)
    m2[rng.random(m2.shape) < 0.03] ^= 255
    mm = {"m1": m1, "m2": m2}
    ops = {"erode": cv2.MORPH_ERODE, "dilate": cv2.MORPH_DILATE, "open": cv2.MORPH_OPEN, "close": cv2.MORPH_CLOSE}
    for k in MORPH_KS:
        ker = cv2.getStructuringElement(cv2.MORPH_RECT, (k, k))
        for name, op in ops.items():
            mm[f"m1_{name}_{k}"] = cv2.morphologyEx(m1, op, ker)
            mm[f"m2_{name}_{k}"] = cv2.morphologyEx(m2, op, ker)
    np.savez_compressed(os.path.join(HERE, "cv2_morph.npz"), **mm)

    # ---- CCL -------------------------------------------------------------------------------------------------
    cc = {}
    for i, (shape, p) in enumerate([((40, 70), 0.3), ((64, 64), 0.5), ((33, 97), 0.62), ((50, 50), 0.05)]):
        m = (rng.random(shape) < p).astype(np.uint8) * 255
        n, lab, st, _ = cv2.connectedComponentsWithStats(m, connectivity=4, ltype=cv2.CV_32S)
        cc[f"mask{i}"] = m
        cc[f"labels{i}"] = lab.astype(np.int32)
        cc[f"stats{i}"] = st[1:].astype(np.int32)  # x, y, w, h, area per component
    np.savez_compressed(os.path.join(HERE, "cv2_ccl.npz"), **cc)

    # ---- reference fixtures + Rust-path regression vectors ---------------------------------------------------------
    frames = {}
    rust_path = {}
    for i in (1, 2, 3):
        p = f"/root/reference/contaminated_{i}.jpg"
        if not os.path.exists(p):
            continue
        img = cv2.imread(p, cv2.IMREAD_COLOR)  # BGR; the bridge hands this array to heimdall_core unchanged
        frames[f"contaminated_{i}"] = img
        r = O.detect_contamination(img)
        rust_path[f"contaminated_{i}"] = {
            "sha256": sha(img), "shape": list(img.shape), "fg_pixels": int((r.mask == 255).sum()),
            "ncomp": r.ncomp, "mask_sha256": sha(r.mask), "labels_sha256": sha(r.labels),
            "defects": [[d["position"][0], d["position"][1], d["size"], d["confidence"]] for d in r.defects]}
    if frames:
        np.savez_compressed(os.path.join(HERE, "fixture_frames.npz"), **frames)
        json.dump(rust_path, open(os.path.join(HERE, "rust_path.json"), "w"), indent=1)

    # ---- unmodified reference Python fallback ---------------------------------------------------------------------
    if os.path.isdir("/root/reference/heimdall"):
        sys.path.insert(0, "/root/reference")
        import logging
        logging.disable(logging.CRITICAL)
        from heimdall.rust_bridge import RustBridge, RUST_AVAILABLE
        assert not RUST_AVAILABLE  # nothing named heimdall_core may be importable while generating these
        ref = {}
        for idx in (0, 3, 5):
            fr = synth.bottle_frame(240, 320, idx, contaminants=2)
            out = RustBridge.detect_contamination(np.dstack([fr, fr, fr]), 10.0, 3000.0, 25.0)
            ref[str(idx)] = [{"position": list(map(int, d["position"])), "size": float(d["size"]),
                              "confidence": float(d["confidence"])} for d in out["defects"]]
        json.dump({"frames": "synth.bottle_frame(240,320,idx,contaminants=2) replicated to 3 channels",
                   "results": ref}, open(os.path.join(HERE, "reference_python.json"), "w"), indent=1)
        logging.disable(logging.NOTSET)

    # ---- synthetic bottle expectations (oracle) --------------------------------------------------------------------
    exp = {}
    for (h, w, idx, kw) in [(1024, 1280, 0, {}), (1024, 1280, 1, {"contaminants": 2}), (1024, 1280, 2, {}),
                            (1024, 1280, 3, {"contaminants": 3}), (480, 640, 4, {"contaminants": 1}),
                            (2048, 2448, 5, {"contaminants": 3})]:
        fr = synth.bottle_frame(h, w, idx, **kw)
        r = O.detect_contamination(fr)
        exp[f"{h}x{w}_{idx}"] = {"kw": kw, "sha256": sha(fr), "fg_pixels": int((r.mask == 255).sum()),
                                 "ncomp": r.ncomp, "mask_sha256": sha(r.mask), "labels_sha256": sha(r.labels),
                                 "defects": [[d["position"][0], d["position"][1], d["size"], d["confidence"]]
                                             for d in r.defects]}
    json.dump(exp, open(os.path.join(HERE, "bottle_expectations.json"), "w"), indent=1)
    json.dump(meta, open(os.path.join(HERE, "meta.json"), "w"), indent=1)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
