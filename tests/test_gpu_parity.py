"""GPU parity tests: the CUDA path, called through the C ABI (heimdall_core -> libheimdall_cuda.so), against the CPU
oracle on the same inputs.  Bit-exact for masks, blurred intermediates, labels, blob statistics, defect lists
(including the f64 confidences) and reject decisions."""
import hashlib
import json
import os

import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def check_frame(oracle, det, img, *, min_size=10.0, max_size=3000.0, threshold=25.0, morph_open_k=0, morph_close_k=0,
                gauss=None, check_blur=True):
    """Run one frame through both implementations and compare everything."""
    import heimdall_core as hc
    kw = {}
    okw = {}
    if gauss is not None:
        kw.update(blur_mode=hc._abi.HV_BLUR_GAUSSIAN, blur_ksize=gauss[0], gauss_sigma=gauss[1])
        okw.update(gauss_ksize=gauss[0], gauss_sigma=gauss[1])
    p = hc.make_params(min_size, max_size, threshold, morph_open_k=morph_open_k, morph_close_k=morph_close_k, **kw)
    ref = oracle.detect_contamination(img, min_size, max_size, threshold, morph_open_k=morph_open_k,
                                      morph_close_k=morph_close_k, **okw)
    _, oblobs = oracle.label4(ref.mask)
    exp = [(d["position"], d["size"], d["confidence"], d["label"], d["bbox"]) for d in ref.defects]
    # run 1 materialises the blurred intermediate (which disables the flat-tile skip), run 2 is the production
    # configuration with the sparsity fast path enabled; both must match the oracle bit for bit
    for with_blur in ((True, False) if check_blur else (False,)):
        dbg = ["gray", "mask", "labels", "blobs"] + (["blur"] if with_blur else [])
        res = det.detect_batch(img, p, debug=dbg)
        assert np.array_equal(res.debug["gray"][0], ref.gray), "gray"
        if with_blur:
            assert np.array_equal(res.debug["blur"][0], ref.blur), "blur"
        assert np.array_equal(res.debug["mask"][0], ref.mask), "mask"
        assert int(res.frames["n_components"][0]) == ref.ncomp
        assert int(res.frames["fg_pixels"][0]) == int((ref.mask == 255).sum())
        assert np.array_equal(res.debug["labels"][0], ref.labels), "labels"
        gb = res.debug["blobs"][0]
        for fld in ("area", "ymin", "ymax", "xmin", "xmax", "sum_y", "sum_x"):
            assert np.array_equal(gb[fld], oblobs[fld]), fld
        got = [((int(d["y"]), int(d["x"])), float(d["size"]), float(d["confidence"]), int(d["label"]),
                (int(d["ymin"]), int(d["xmin"]), int(d["ymax"]), int(d["xmax"]))) for d in res.defects_of(0)]
        assert got == exp  # exact f64 equality on confidences
        assert bool(res.rejected[0]) == ref.reject
    return res, ref


# ---- known-answer tests through the GPU -------------------------------------------------------------
def test_kat_squares(oracle, detector):
    img = np.full((1024, 1280, 1), 220, np.uint8)
    img[500:520, 600:620] = 40
    res, _ = check_frame(oracle, detector, img)
    d = res.defects_of(0)
    assert len(d) == 1 and (d[0]["y"], d[0]["x"]) == (509, 609) and d[0]["size"] == 208.0
    assert d[0]["confidence"] == 0.844
    img = np.full((1024, 1280, 1), 220, np.uint8)
    img[500:504, 600:604] = 40
    res, _ = check_frame(oracle, detector, img)
    assert repr(float(res.defects_of(0)[0]["confidence"])) == "0.7999999999999999"


def test_kat_uniform_and_ramp(oracle, detector):
    res, _ = check_frame(oracle, detector, np.full((1024, 1280, 1), 220, np.uint8))
    assert not res.rejected[0] and int(res.frames["fg_pixels"][0]) == 0
    ramp = ((np.arange(1280) * 255) // 1280).astype(np.uint8)
    res, _ = check_frame(oracle, detector, np.tile(ramp, (1024, 1))[:, :, None])
    assert not res.rejected[0]


def test_kat5_gray_on_gpu(oracle, detector):
    v = np.arange(256, dtype=np.uint8)
    img = np.ascontiguousarray(np.broadcast_to(np.stack([v, v, v], -1)[None], (16, 256, 3)))
    res = detector.detect_batch(img, debug=["gray"])
    assert np.array_equal(res.debug["gray"][0], oracle.gray(img))
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    res = detector.detect_batch(img, debug=["gray"])
    assert np.array_equal(res.debug["gray"][0], oracle.gray(img))


# ---- shapes: tiny, ragged, non-multiple-of-4/32/128 widths, below the blur / window sizes ---------------------------
SHAPES = [(1, 1), (1, 7), (7, 1), (3, 3), (4, 4), (5, 5), (4, 40), (40, 4), (10, 10), (11, 11), (12, 13), (31, 33),
          (32, 128), (33, 129), (46, 144), (64, 127), (65, 257), (100, 130), (37, 1283)]


@pytest.mark.parametrize("shape", SHAPES)
def test_random_frames_all_shapes(oracle, detector, shape):
    h, w = shape
    rng = np.random.default_rng(h * 10007 + w)
    img = rng.integers(140, 256, (h, w, 1), dtype=np.uint8)
    dark = rng.random((h, w)) < 0.07
    img[dark] = rng.integers(0, 90, (int(dark.sum()), 1), dtype=np.uint8)
    check_frame(oracle, detector, img, min_size=1.0, threshold=10.0)
    check_frame(oracle, detector, img)


@pytest.mark.parametrize("c", [1, 3])
def test_random_textured(oracle, detector, c):
    rng = np.random.default_rng(99 + c)
    img = rng.integers(0, 256, (150, 200, c), dtype=np.uint8)  # no flat tile anywhere: exercises the full path
    check_frame(oracle, detector, img, min_size=1.0, max_size=1e9, threshold=3.0)
    check_frame(oracle, detector, img, min_size=3.0, max_size=40.0, threshold=0.0)


@pytest.mark.parametrize("thr", [0.0, 3.0, 25.0, 120.0, 255.0, 256.0, -1.0, -2.0])
def test_fast_path_interior_tiles(oracle, detector, thr):
    """Images large enough to contain interior 128x32 tiles (packed-u16x2 fast path) next to border tiles (generic
    path), textured so that no tile is flat; both paths must agree with the oracle and with each other."""
    import heimdall_core as hc
    rng = np.random.default_rng(int(thr) + 1000)
    img = rng.integers(0, 256, (200, 650, 1), dtype=np.uint8)
    img[60:140, 200:500] = (img[60:140, 200:500] // 8 + 100)          # a calmer region: thresholds near 25 matter
    img[90:110, 300:340] = 10
    res, ref = check_frame(oracle, detector, img, min_size=1.0, max_size=1e9, threshold=thr)
    gen = hc.Detector(0, force_generic=True, max_defects_per_frame=8192)
    try:
        r2 = gen.detect_batch(img, hc.make_params(1.0, 1e9, thr), debug=["mask", "labels"])
    finally:
        gen.close()
    assert np.array_equal(r2.debug["mask"][0], ref.mask) and np.array_equal(r2.debug["labels"][0], ref.labels)


def test_fast_path_extreme_values(oracle, detector):
    """Saturated inputs: all-255 blocks next to all-0 blocks maximise every packed 16-bit lane (30855 + 1, 61831)."""
    img = np.zeros((160, 520, 1), np.uint8)
    img[:, ::37] = 255
    img[40:120, 100:400] = 255
    img[70:90, 200:300] = 0
    for thr in (0.0, 25.0, 255.0):
        check_frame(oracle, detector, img, min_size=1.0, max_size=1e9, threshold=thr)
    check_frame(oracle, detector, np.full((160, 520, 1), 255, np.uint8), threshold=0.0)


@pytest.mark.parametrize("thr", [-1.0, -40.0, -300.0, 0.0, 0.9, 1e12, -1e12, float("nan"), 254.0, 255.0, 256.0, 300.0,
                                 2147483647.0, -2147483648.0, -2147483548.0, -2147483393.0, -2147483392.0, float("-inf")])
def test_threshold_edge_values(oracle, detector, thr):
    """`threshold as i32` saturates and `mean - c` wraps in i32 like the reference's release build (detection.rs:186,211):
    at and below -2147483393.0 pixels whose window mean reaches c + 2^31 are never foreground (-1e12: empty mask), one
    above that every pixel is."""
    rng = np.random.default_rng(17)
    img = rng.integers(0, 256, (40, 70, 1), dtype=np.uint8)
    res, ref = check_frame(oracle, detector, img, min_size=1.0, max_size=1e9, threshold=thr)
    if thr in (-1e12, -2147483648.0, float("-inf")):
        assert int(res.frames["fg_pixels"][0]) == 0
    if thr == -2147483392.0:
        assert int(res.frames["fg_pixels"][0]) == 40 * 70
    if thr == -2147483548.0:   # T = 100: foreground exactly where the window mean is below 100
        assert 0 < int(res.frames["fg_pixels"][0]) < 40 * 70
    # the same on a frame with interior 128x32 tiles (TMA kernel, 16-px aligned)
    big = rng.integers(60, 140, (96, 400, 1), dtype=np.uint8)
    check_frame(oracle, detector, big, min_size=1.0, max_size=1e9, threshold=thr, check_blur=False)


def test_size_filters_inclusive(oracle, detector):
    img = np.full((80, 120, 1), 220, np.uint8)
    img[20:24, 30:34] = 40   # ring of 24 px (KAT2 geometry)
    for mn, mx, n in [(24.0, 24.0, 1), (24.5, 100.0, 0), (1.0, 23.9, 0), (float("nan"), 100.0, 0)]:
        res, ref = check_frame(oracle, detector, img, min_size=mn, max_size=mx)
        assert len(res.defects_of(0)) == n == len(ref.defects)


# ---- synthetic bottle frames at the headline resolution, against the oracle and the committed goldens --------------
def test_bottle_frames_match_oracle_and_golden(oracle, detector, golden_dir):
    exp = json.load(open(os.path.join(golden_dir, "bottle_expectations.json")))
    for key in ("1024x1280_0", "1024x1280_1", "480x640_4"):
        e = exp[key]
        hw, idx = key.split("_")
        h, w = map(int, hw.split("x"))
        fr = synth.bottle_frame(h, w, int(idx), **e["kw"])
        assert sha(fr) == e["sha256"]
        res, _ = check_frame(oracle, detector, fr[:, :, None])
        assert sha(res.debug["mask"][0]) == e["mask_sha256"] and sha(res.debug["labels"][0]) == e["labels_sha256"]
        got = [[int(d["y"]), int(d["x"]), float(d["size"]), float(d["confidence"])] for d in res.defects_of(0)]
        assert got == e["defects"]


def test_golden_only_large_frame(detector, golden_dir):
    """2448x2048 frame checked against committed oracle results only (no oracle run: keeps the GPU suite fast)."""
    exp = json.load(open(os.path.join(golden_dir, "bottle_expectations.json")))
    e = exp["2048x2448_5"]
    fr = synth.bottle_frame(2048, 2448, 5, **e["kw"])
    assert sha(fr) == e["sha256"]
    res = detector.detect_batch(fr, debug=["mask", "labels"])
    assert sha(res.debug["mask"][0]) == e["mask_sha256"] and sha(res.debug["labels"][0]) == e["labels_sha256"]
    got = [[int(d["y"]), int(d["x"]), float(d["size"]), float(d["confidence"])] for d in res.defects_of(0)]
    assert got == e["defects"] and int(res.frames["n_components"][0]) == e["ncomp"]


def test_reference_fixture_images(oracle, detector, golden_dir):
    frames = np.load(os.path.join(golden_dir, "fixture_frames.npz"))
    exp = json.load(open(os.path.join(golden_dir, "rust_path.json")))
    for name, e in exp.items():
        res, _ = check_frame(oracle, detector, frames[name])
        got = [[int(d["y"]), int(d["x"]), float(d["size"]), float(d["confidence"])] for d in res.defects_of(0)]
        assert got == e["defects"] and sha(res.debug["mask"][0]) == e["mask_sha256"]


def test_batch_of_25_equals_single_frames(oracle, detector):
    """BASELINE config 2: one second of line (25 frames) in one call == 25 single-frame calls == oracle."""
    batch = synth.bottle_batch(25, 1024, 1280, start_index=100)
    res = detector.detect_batch(batch[..., None], debug=["mask", "labels"])
    assert res.frames.shape == (25,)
    for f in (0, 7, 24):
        ref = oracle.detect_contamination(batch[f][:, :, None])
        assert np.array_equal(res.debug["mask"][f], ref.mask) and np.array_equal(res.debug["labels"][f], ref.labels)
        got = [((int(d["y"]), int(d["x"])), float(d["size"]), float(d["confidence"])) for d in res.defects_of(f)]
        assert got == [(d["position"], d["size"], d["confidence"]) for d in ref.defects]
        assert all(int(d["frame"]) == f for d in res.defects_of(f))
    single = detector.detect_batch(batch[3][:, :, None])
    assert np.array_equal(single.defects_of(0)[["y", "x", "size", "confidence"]],
                          res.defects_of(3)[["y", "x", "size", "confidence"]])
    # offsets are a running sum
    off = np.concatenate([[0], np.cumsum(res.frames["n_defects"])[:-1]])
    assert np.array_equal(res.frames["defects_offset"], off)


# ---- CCL stress through find_contours (mask given directly) ----------------------------------------------------------
def _masks():
    out = {}
    h, w = 96, 160
    yy, xx = np.mgrid[0:h, 0:w]
    out["checkerboard"] = (((yy + xx) & 1) * 255).astype(np.uint8)          # max component count
    out["all_fg"] = np.full((h, w), 255, np.uint8)
    out["all_bg"] = np.zeros((h, w), np.uint8)
    comb = np.zeros((h, w), np.uint8)
    comb[::2, :] = 255
    comb[:, 0] = 255                                                        # comb: long vertical spine
    out["comb"] = comb
    comb2 = np.zeros((h, w), np.uint8)
    comb2[:, ::2] = 255
    comb2[h - 1, :] = 255                                                   # teeth joined only at the last row
    out["comb_bottom"] = comb2
    sp = np.zeros((h, w), np.uint8)                                         # rectangular spiral, 1-px wide
    t, b, l, r = 0, h - 1, 0, w - 1
    while t <= b and l <= r:
        sp[t, l:r + 1] = 255
        sp[t:b + 1, r] = 255
        if b - t >= 2:
            sp[b, l + 2:r + 1] = 255
        if r - l >= 2 and b - t >= 4:
            sp[t + 2:b + 1, l + 2] = 255
        t, b, l, r = t + 4, b - 4, l + 4, r - 4
    out["spiral"] = sp
    rng = np.random.default_rng(5)
    out["dense_random"] = ((rng.random((h, w)) < 0.59) * 255).astype(np.uint8)   # near the percolation threshold
    out["gt127"] = rng.integers(0, 256, (h, w), dtype=np.uint8)
    return out


@pytest.mark.parametrize("name", list(_masks().keys()))
def test_ccl_adversarial_masks(oracle, detector, name):
    m = _masks()[name]
    recs, labels = detector.find_contours(m[:, :, None], 0.0, 1e18)
    ref_labels, ref_blobs = oracle.label4(m, fg_gt127=True)
    assert np.array_equal(labels, ref_labels)
    assert [int(r.pixel_count) for r in recs] == ref_blobs["area"].tolist()
    assert [(r.y, r.x) for r in recs] == [(int(b["sum_y"] // b["area"]), int(b["sum_x"] // b["area"]))
                                          for b in ref_blobs]


def test_capacity_error_is_loud(detector):
    """More components than the stats table holds -> HV_ERR_CAPACITY, never a silent truncation."""
    import heimdall_core as hc
    det = hc.Detector(0, max_blobs_per_frame=100, max_defects_per_frame=8)
    try:
        yy, xx = np.mgrid[0:64, 0:64]
        m = (((yy + xx) & 1) * 255).astype(np.uint8)
        with pytest.raises(hc.HeimdallCudaError) as ei:
            det.find_contours(np.ascontiguousarray(m[:, :, None]), 0.0, 1e18)
        assert ei.value.status == hc._abi.HV_ERR_CAPACITY
        rng = np.random.default_rng(8)
        img = rng.integers(0, 256, (64, 64, 1), dtype=np.uint8)
        p = hc.make_params(1.0, 1e9, 0.0)
        with pytest.raises(hc.HeimdallCudaError):
            det.detect_batch(img, p)
        res = det.detect_batch(img, p, raise_on_capacity=False)
        assert res.status == hc._abi.HV_ERR_CAPACITY and res.frames["status"][0] == hc._abi.HV_ERR_CAPACITY
        assert res.frames["n_components"][0] > 8        # the true component count is still reported
        assert res.frames["n_defects"][0] == 8          # ... and the defect slots that exist are filled
        # a frame that fits is unaffected
        ok = det.detect_batch(np.full((64, 64, 1), 200, np.uint8))
        assert ok.status == 0 and ok.frames["n_components"][0] == 0
    finally:
        det.close()


# ---- extension stages ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k_open,k_close", [(3, 3), (3, 0), (0, 5), (5, 7), (15, 15), (2, 4), (9, 3)])
def test_morphology_open_close(oracle, detector, k_open, k_close):
    fr = synth.bottle_frame(240, 333, 11, contaminants=3)
    check_frame(oracle, detector, fr[:, :, None], min_size=1.0, morph_open_k=k_open, morph_close_k=k_close)
    rng = np.random.default_rng(k_open * 31 + k_close)
    img = rng.integers(0, 256, (90, 101, 1), dtype=np.uint8)
    check_frame(oracle, detector, img, min_size=1.0, max_size=1e9, threshold=2.0, morph_open_k=k_open,
                morph_close_k=k_close)


@pytest.mark.parametrize("k_open,k_close,shape", [(3, 3, (1024, 1280)), (7, 5, (1024, 1280)), (15, 13, (200, 256)),
                                                   (17, 0, (150, 192)), (0, 21, (150, 200)), (4, 6, (97, 128)),
                                                   (15, 15, (33, 31))])
def test_morphology_fused_kernel_and_fallback(oracle, detector, k_open, k_close, shape):
    """One-kernel open+close+expansion (kernel sizes up to 15, reach <= 28 px) against the oracle on tile-aligned and
    ragged shapes, frames smaller than the reach, and the multi-pass fallback for larger kernels; a batch of two frames
    so that neighbouring frames must not leak into each other's halo."""
    import heimdall_core as hc
    h, w = shape
    rng = np.random.default_rng(k_open * 100 + k_close)
    a = synth.bottle_frame(h, w, 77, contaminants=3) if min(h, w) >= 150 else rng.integers(0, 256, (h, w), dtype=np.uint8)
    bimg = rng.integers(0, 256, (h, w), dtype=np.uint8)
    bimg[: h // 2] = 128  # half flat, half texture: empty and busy tiles side by side
    batch = np.stack([a, bimg])[..., None]
    p = hc.make_params(1.0, 1e9, 2.0 if min(h, w) < 150 else 25.0, morph_open_k=k_open, morph_close_k=k_close)
    res = detector.detect_batch(batch, p, debug=["mask", "labels"])
    for f in range(2):
        ref = oracle.detect_contamination(batch[f], 1.0, 1e9, 2.0 if min(h, w) < 150 else 25.0, morph_open_k=k_open,
                                          morph_close_k=k_close)
        assert np.array_equal(res.debug["mask"][f], ref.mask)
        assert np.array_equal(res.debug["labels"][f], ref.labels)
        assert [(int(d["y"]), int(d["x"]), float(d["size"]), float(d["confidence"])) for d in res.defects_of(f)] == \
            [(d["position"][0], d["position"][1], d["size"], d["confidence"]) for d in ref.defects]


def test_morphology_entry_matches_opencv_golden(detector, golden_dir):
    """hv_morphology (separable bit-packed erode / dilate kernels) against committed outputs of cv2.morphologyEx on two
    masks: every operation and kernel size of tests/golden/make_golden.py, bit for bit."""
    z = np.load(os.path.join(golden_dir, "cv2_morph.npz"))
    for name in ("m1", "m2"):
        m = z[name]
        for k in (2, 3, 4, 5, 7, 9, 11, 13, 15):
            assert np.array_equal(detector.morphology(m, open_k=k), z[f"{name}_open_{k}"]), (name, "open", k)
            assert np.array_equal(detector.morphology(m, close_k=k), z[f"{name}_close_{k}"]), (name, "close", k)
        assert np.array_equal(detector.morphology(m), m)


@pytest.mark.parametrize("case", ["240x333_11", "1024x1280_1", "96x128_7"])
def test_morphology_pipeline_matches_opencv_golden(detector, golden_dir, case):
    """The morphology the detector runs between threshold and CCL (folded into K1 for 3x3 / 5x5 kernels, the tiles kernel
    up to 15x15) against masks produced by opencv-python itself: cv2 MORPH_OPEN then MORPH_CLOSE of the frame's
    pre-morphology mask, committed bit-packed in cv2_morph_pipeline.npz.  No oracle involved."""
    import heimdall_core as hc
    z = np.load(os.path.join(golden_dir, "cv2_morph_pipeline.npz"))
    hw, idx = case.split("_")
    h, w = map(int, hw.split("x"))
    kw = {"240x333_11": {"contaminants": 3}, "1024x1280_1": {"contaminants": 2}, "96x128_7": {"contaminants": 2}}[case]
    fr = synth.bottle_frame(h, w, int(idx), **kw)

    def unpack(a):
        return (np.unpackbits(a)[:h * w].reshape(h, w) * 255).astype(np.uint8)

    pre = detector.detect_batch(fr, debug=["mask"]).debug["mask"][0]
    assert np.array_equal(pre, unpack(z[f"{case}_pre"]))
    for ko, kc in [(3, 3), (3, 0), (0, 5), (5, 7), (9, 3), (15, 15)]:
        got = detector.detect_batch(fr, hc.make_params(morph_open_k=ko, morph_close_k=kc), debug=["mask"]).debug["mask"][0]
        assert np.array_equal(got, unpack(z[f"{case}_o{ko}_c{kc}"])), (case, ko, kc)
        assert np.array_equal(detector.morphology(pre, ko, kc), got)


@pytest.mark.parametrize("k,s", [(5, 0.0), (3, 0.0), (7, 1.0), (13, 2.0), (15, 3.0), (9, 1.5)])
def test_gaussian_blur_mode(oracle, detector, golden_dir, k, s):
    fr = synth.bottle_frame(200, 260, 21, contaminants=2)
    res, ref = check_frame(oracle, detector, fr[:, :, None], min_size=1.0, gauss=(k, s))
    z = np.load(os.path.join(golden_dir, "cv2_gaussian.npz"))
    key = f"b_k{k}_s{s if s else 0}"
    if key in z.files:
        got = detector.detect_batch(z["src2"], _gauss_params(k, s), debug=["blur"]).debug["blur"][0]
        assert np.array_equal(got, z[key])  # bit-exact vs cv2 (north star allows +-1 LSB)


@pytest.mark.parametrize("k,s", [(3, 0.0), (5, 0.0), (7, 1.0), (11, 0.0), (13, 2.0), (15, 3.0)])
@pytest.mark.parametrize("shape", [(224, 320), (96, 128), (33, 16), (160, 1296)])
def test_gaussian_fused_into_k1(oracle, detector, k, s, shape):
    """16-px aligned frames take the Gaussian variant of the TMA kernel (A7 fused into K1): interior and border tiles,
    flat and non-flat tiles, reflect-101 borders, a partial last tile column (1296 = 10 x 128 + 16)."""
    h, w = shape
    fr = synth.bottle_frame(h, w, 31, contaminants=2)
    check_frame(oracle, detector, fr[:, :, None], min_size=1.0, gauss=(k, s))
    rng = np.random.default_rng(k * 100 + h)
    tex = rng.integers(0, 256, (h, w, 1), dtype=np.uint8)      # every tile non-flat, foreground everywhere
    check_frame(oracle, detector, tex, min_size=1.0, threshold=5.0, gauss=(k, s), check_blur=(w <= 320))


@pytest.mark.parametrize("k", [3, 5, 7, 9, 11, 13, 15])
def test_gaussian_fast_taps_every_kernel_size(oracle, k):
    """The packed-arithmetic Gaussian of K1's interior tiles (vertical u16x2 pass, horizontal IDP.2A pass; halo 3 for k <= 7,
    halo 7 above) for every odd kernel size and several sigmas, on a texture where every tile is non-flat and on bottle frames;
    sigma = 0.1 gives a centre tap of 256, which does not fit a byte and must take the generic taps."""
    import heimdall_core as hc
    detector = hc.Detector(0, max_blobs_per_frame=400000, max_defects_per_frame=200000)
    rng = np.random.default_rng(500 + k)
    tex = rng.integers(0, 256, (352, 640, 1), dtype=np.uint8)
    smooth = (rng.integers(0, 256, (44, 80), dtype=np.uint8).repeat(8, 0).repeat(8, 1)[..., None] // 2 + tex // 2).astype(np.uint8)
    for s in (0.0, 0.7, 2.5):
        check_frame(oracle, detector, tex, min_size=1.0, threshold=5.0, gauss=(k, s), check_blur=False)
        check_frame(oracle, detector, smooth, min_size=1.0, threshold=12.0, gauss=(k, s), check_blur=(s == 0.7))
    fr = synth.bottle_frame(384, 640, 77 + k, contaminants=3)
    for s in (1.3, 0.1):  # (0.1: the blur is the identity; on the textures that is more components than the detector holds)
        check_frame(oracle, detector, fr[:, :, None], min_size=1.0, gauss=(k, s), check_blur=False)
    detector.close()


def _gauss_params(k, s):
    import heimdall_core as hc
    return hc.make_params(blur_mode=hc._abi.HV_BLUR_GAUSSIAN, blur_ksize=k, gauss_sigma=s)


# ---- the drop-in module surface (heimdall_core, lib.rs:14-39) ---------------------------------------------------------
def test_module_detect_contamination_contract(oracle):
    import heimdall_core as hc
    fr = synth.bottle_frame(480, 640, 4, contaminants=1)
    out = hc.detect_contamination(fr[:, :, None])
    assert set(out) == {"defects", "processing_time"} and isinstance(out["processing_time"], float)
    ref = oracle.detect_contamination(fr[:, :, None])
    assert [(d["position"], d["size"], d["confidence"]) for d in out["defects"]] == \
           [(d["position"], d["size"], d["confidence"]) for d in ref.defects]
    d0 = out["defects"][0]
    assert set(d0) == {"position", "size", "confidence", "metadata"} and d0["metadata"] == {}
    assert isinstance(d0["position"][0], int) and isinstance(d0["size"], float)
    out2 = hc.detect_contamination(fr[:, :, None], 50.0, 100.0, 10.0)
    ref2 = oracle.detect_contamination(fr[:, :, None], 50.0, 100.0, 10.0)
    assert [d["position"] for d in out2["defects"]] == [d["position"] for d in ref2.defects]
    with pytest.raises(TypeError):
        hc.detect_contamination(fr)                       # 2-D: PyReadonlyArray3 extraction fails
    with pytest.raises(TypeError):
        hc.detect_contamination(fr[:, :, None].astype(np.float32))
    with pytest.raises(ValueError, match="Invalid image dimensions"):
        hc.detect_contamination(np.zeros((8, 8, 2), np.uint8))
    # non-contiguous view
    big = np.dstack([fr, fr])
    out3 = hc.detect_contamination(big[:, :, 1:2])
    assert [d["position"] for d in out3["defects"]] == [d["position"] for d in ref.defects]


def test_module_process_image(oracle):
    import heimdall_core as hc
    fr = synth.bottle_frame(200, 300, 9, contaminants=2)
    img = np.dstack([fr, fr, fr])
    out = hc.process_image(img, "contamination")
    vis, contours = oracle.contamination_pipeline(img)
    assert np.array_equal(out["processed_image"], vis) and out["contours"] == contours
    out = hc.process_image(img, "basic")
    assert np.array_equal(out["processed_image"], oracle.basic_pipeline(img))
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (64, 80, 3), dtype=np.uint8)
    out = hc.process_image(img, "contamination", {})
    vis, contours = oracle.contamination_pipeline(img)
    assert np.array_equal(out["processed_image"], vis) and out["contours"] == contours
    with pytest.raises(ValueError, match="Unsupported pipeline type: nope"):
        hc.process_image(img, "nope")
    with pytest.raises(ValueError):
        hc.process_image(img[:, :, :1], "contamination")
    b = hc.benchmark_processing(img, 2)
    assert set(b) == {"basic_pipeline_time", "contamination_pipeline_time", "iterations"} and b["iterations"] == 2


def test_module_processing_and_detection_submodules(oracle):
    import heimdall_core as hc
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, (50, 70, 3), dtype=np.uint8)
    for gs, bs in [(None, None), (True, 5), (False, 3), (True, 9), (False, 0), (True, -3)]:
        got = hc.processing.preprocess_image(img, gs, bs)
        exp = oracle.preprocess_image(img, True if gs is None else gs, bs)
        assert np.array_equal(got, exp), (gs, bs)
    g = hc.processing.preprocess_image(img, True, 5)
    for thr, ad, inv in [(None, None, None), (100, False, True), (None, True, False), (None, True, True), (0, False, False)]:
        got = hc.processing.apply_threshold(g, thr, ad, inv)
        exp = oracle.apply_threshold(g, 127 if thr is None else thr, bool(ad), bool(inv))
        assert np.array_equal(got, exp), (thr, ad, inv)
    with pytest.raises(ValueError, match="Thresholding requires a grayscale image"):
        hc.processing.apply_threshold(img)
    m = np.zeros((60, 90, 1), np.uint8)
    m[2:6, 3:9] = 200
    m[10:25, 10:25] = 255
    m[40:43, 50:80] = 130
    m[30, 30] = 255
    got = hc.detection.find_contours(m)
    assert got == oracle.find_contours(m)                  # including the DFS-ordered "points"
    assert hc.detection.find_contours(m, 1.0, 30.0) == oracle.find_contours(m, 1.0, 30.0)
    with pytest.raises(ValueError, match="Contour detection requires"):
        hc.detection.find_contours(img)
    a = hc.acquisition.acquire_image("simulation")
    assert a.shape == (480, 640, 3) and a[0, 0, 0] == 220
    with pytest.raises(ValueError, match="Unsupported source type"):
        hc.acquisition.acquire_image("nope")


# ---- asynchronous and device-resident entry points ------------------------------------------------------------------
def test_submit_wait_pipeline(oracle, detector):
    import ctypes
    n, h, w = 4, 256, 320
    nbytes = n * h * w
    bufs = [detector.host_alloc(nbytes) for _ in range(3)]
    try:
        batches = [synth.bottle_batch(n, h, w, start_index=200 + 10 * i) for i in range(3)]
        for b, p in zip(batches, bufs):
            ctypes.memmove(p, b.ctypes.data, nbytes)
        tickets = [detector.submit(p, n, h, w) for p in bufs]
        for t, b in zip(tickets, batches):
            res = detector.wait(t, n)
            for f in range(n):
                ref = oracle.detect_contamination(b[f][:, :, None], want_intermediates=False)
                got = [((int(d["y"]), int(d["x"])), float(d["size"]), float(d["confidence"])) for d in res.defects_of(f)]
                assert got == [(d["position"], d["size"], d["confidence"]) for d in ref.defects]
        with pytest.raises(Exception):
            detector.wait(tickets[0], n)   # ticket already consumed
    finally:
        for p in bufs:
            detector.host_free(p)


def test_device_resident_path_with_torch(oracle, detector):
    import torch
    n, h, w = 3, 300, 420
    batch = synth.bottle_batch(n, h, w, start_index=300, contaminants=2)
    d_in = torch.from_numpy(batch).cuda()
    d_mask = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
    d_lab = torch.empty((n, h, w), dtype=torch.int32, device="cuda")
    detector.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        res = detector.detect_device(d_in.data_ptr(), n, h, w, 1, None, d_mask.data_ptr(), d_lab.data_ptr())
        torch.cuda.synchronize()
    finally:
        detector.set_stream(None)
    for f in range(n):
        ref = oracle.detect_contamination(batch[f][:, :, None])
        assert np.array_equal(d_mask[f].cpu().numpy(), ref.mask)
        assert np.array_equal(d_lab[f].cpu().numpy(), ref.labels)
        assert [(int(d["y"]), int(d["x"])) for d in res.defects_of(f)] == [d["position"] for d in ref.defects]


def test_line_stats_and_launch_count(oracle):
    import heimdall_core as hc
    det = hc.Detector(0, profile=True)
    try:
        batch = synth.bottle_batch(5, 240, 320, start_index=400)
        l0 = det.launch_count()
        res = det.detect_batch(batch[..., None])
        assert det.launch_count() - l0 >= 2      # K1 + fused per-frame CCL (global path: K1 + 5)
        s = det.stats()
        assert s["frames_inspected"] == 5 and s["frames_rejected"] == int(res.rejected.sum())
        assert s["total_defects"] == int(res.frames["n_defects"].sum())
        assert s["total_components"] == int(res.frames["n_components"].sum())
        assert s["total_fg_pixels"] == int(res.frames["fg_pixels"].sum())
        assert s["total_defect_area"] == int(res.defects["size"].sum())
        assert sum(s["area_hist"]) == s["total_defects"]
        prof = det.profile()
        assert prof["preprocess_mask"]["ms"] > 0 and prof["ccl_frame_fused"]["launches"] == 1
        det.stats_reset()
        assert det.stats()["frames_inspected"] == 0
    finally:
        det.close()


# ---- full-size, size-independent properties (BASELINE configs 3 and 4) -------------------------------------------------
def test_full_size_properties_12mp_high_contamination(detector):
    """4096x3000, >= 10k blobs: label map properties that do not need the oracle at this size:
    labels > 0 exactly on the mask, labels are 1..K with K = n_components, first occurrences are in increasing raster
    order (canonical numbering), per-label pixel counts equal the blob table, 4-neighbours of equal mask share labels."""
    import heimdall_core as hc
    fr = synth.high_contamination_frame(3000, 4096, 0)
    det = hc.Detector(0, max_blobs_per_frame=400000, max_defects_per_frame=200000)
    try:
        res = det.detect_batch(fr, hc.make_params(10.0, 3000.0, 25.0), debug=["mask", "labels", "blobs"])
    finally:
        det.close()
    mask, lab, blobs = res.debug["mask"][0], res.debug["labels"][0], res.debug["blobs"][0]
    k = int(res.frames["n_components"][0])
    assert k >= 10000 and len(blobs) == k
    assert np.array_equal(lab > 0, mask == 255)
    flat = lab.ravel()
    fg = flat[flat > 0]
    uniq, first = np.unique(fg, return_index=True)
    assert np.array_equal(uniq, np.arange(1, k + 1)) and np.all(np.diff(first) > 0)
    assert np.array_equal(np.bincount(fg, minlength=k + 1)[1:], blobs["area"])
    both_h = (mask[:, 1:] == 255) & (mask[:, :-1] == 255)
    assert np.array_equal(lab[:, 1:][both_h], lab[:, :-1][both_h])
    both_v = (mask[1:, :] == 255) & (mask[:-1, :] == 255)
    assert np.array_equal(lab[1:, :][both_v], lab[:-1, :][both_v])
    ys, xs = np.nonzero(lab)
    assert np.array_equal(np.bincount(lab[ys, xs], weights=ys, minlength=k + 1)[1:].astype(np.uint64), blobs["sum_y"])
    assert np.array_equal(np.bincount(lab[ys, xs], weights=xs, minlength=k + 1)[1:].astype(np.uint64), blobs["sum_x"])
    assert int(res.frames["n_defects"][0]) > 100 and res.rejected[0]


def test_idempotent_and_batch_invariant(detector):
    """Same frame anywhere in a batch gives the same answer (frames are independent), and re-running is identical."""
    fr = synth.bottle_frame(1024, 1280, 500, contaminants=3)
    other = synth.bottle_frame(1024, 1280, 501, contaminants=0)
    batch = np.stack([other, fr, other, fr])[..., None]
    a = detector.detect_batch(batch, debug=["labels"])
    b = detector.detect_batch(batch, debug=["labels"])
    assert np.array_equal(a.debug["labels"], b.debug["labels"])
    assert np.array_equal(a.debug["labels"][1], a.debug["labels"][3])
    assert np.array_equal(a.defects_of(1)[["y", "x", "size", "confidence"]],
                          a.defects_of(3)[["y", "x", "size", "confidence"]])


def test_dense_frames_defect_order_across_score_chunks(oracle):
    """Thousands of blobs per frame: the global-memory path scores them in chunks of 256 on different CTAs and stitches the
    defect list together with a look-back over the chunk counts -- the list must still be the oracle's, in its order,
    for every frame of the batch, twice in a row (the chunk counters clean up after themselves)."""
    import heimdall_core as hc
    frames = np.stack([synth.high_contamination_frame(1500, 2048, s) for s in (3, 4, 5)])
    det = hc.Detector(0, max_blobs_per_frame=100000, max_defects_per_frame=50000)
    try:
        for rep in range(2):
            res = det.detect_batch(frames[..., None], debug=["labels"])
            for f in range(len(frames)):
                ref = oracle.detect_contamination(frames[f][:, :, None])
                assert int(res.frames["n_components"][f]) == ref.ncomp > 2000
                assert np.array_equal(res.debug["labels"][f], ref.labels)
                got = [((int(d["y"]), int(d["x"])), float(d["size"]), float(d["confidence"]), int(d["label"]))
                       for d in res.defects_of(f)]
                exp = [(d["position"], d["size"], d["confidence"]) for d in ref.defects]
                assert [g[:3] for g in got] == exp and len(got) > 256
                assert [g[3] for g in got] == sorted(g[3] for g in got)
    finally:
        det.close()


def test_compressible_device_buffers(oracle, detector):
    """hv_device_alloc: output planes in L2-compressible memory give the same bytes as ordinary memory; write/read round
    trip; the flag degrades to plain memory when compression is unavailable."""
    n, h, w = 4, 256, 384
    batch = synth.bottle_batch(n, h, w, start_index=900, contaminants=2)
    d_in = detector.device_alloc((n, h, w), np.uint8, compressible=False)
    d_in.set(batch)
    assert np.array_equal(d_in.get(), batch) and not d_in.compressed
    outs = {}
    for comp in (False, True):
        m = detector.device_alloc((n, h, w), np.uint8, compressible=comp)
        l = detector.device_alloc((n, h, w), np.int32, compressible=comp)
        l.set(np.full((n, h, w), -7, np.int32))  # stale contents must not survive
        for rep in range(2):
            res = detector.detect_device(d_in.ptr, n, h, w, 1, None, m.ptr, l.ptr)
        outs[comp] = (m.get(), l.get(), res)
        assert np.array_equal(l.get(1, 2), outs[comp][1][1:3])
        m.free(), l.free()
    assert np.array_equal(outs[False][0], outs[True][0]) and np.array_equal(outs[False][1], outs[True][1])
    for f in range(n):
        ref = oracle.detect_contamination(batch[f][:, :, None])
        assert np.array_equal(outs[True][0][f], ref.mask) and np.array_equal(outs[True][1][f], ref.labels)
        assert [(int(d["y"]), int(d["x"])) for d in outs[True][2].defects_of(f)] == [d["position"] for d in ref.defects]
    d_in.free()


@pytest.mark.parametrize("n_sets,morph_k", [(1, 0), (2, 0), (3, 0), (5, 0), (2, 3), (5, 3), (5, 7)])
def test_batches_in_flight_with_rotating_output_sets(oracle, n_sets, morph_k):
    """Streaming use of enqueue_device: many batches enqueued back to back, the kernels of several of them in flight at
    once (K1 of a batch starts while the per-frame CCL kernels of the previous ones still run), output planes rotated
    over n_sets sets -- fewer than the library's pipeline depth means K1 has to wait on the device for the batch that
    last wrote the set.  Every set must end up holding exactly the oracle's planes of the last batch written into it,
    and the last batch's defect list must be the oracle's."""
    import heimdall_core as hc
    n, h, w = 4, 256, 384
    n_batches = 13
    batches = [synth.bottle_batch(n, h, w, start_index=7000 + 10 * i, contaminants=(i % 4)) for i in range(n_batches)]
    det = hc.Detector(0)
    try:
        assert det.pipeline_depth() >= 3
        d_in = [det.device_alloc((n, h, w), np.uint8, compressible=False) for _ in range(n_batches)]
        for a, bt in zip(d_in, batches):
            a.set(bt)
        masks = [det.device_alloc((n, h, w), np.uint8) for _ in range(n_sets)]
        labels = [det.device_alloc((n, h, w), np.int32) for _ in range(n_sets)]
        params = hc.make_params(morph_open_k=morph_k, morph_close_k=morph_k) if morph_k else None
        okw = dict(morph_open_k=morph_k, morph_close_k=morph_k) if morph_k else {}
        for rep in range(2):
            for i in range(n_batches):
                det.enqueue_device(d_in[i].ptr, n, h, w, 1, params, masks[i % n_sets].ptr, labels[i % n_sets].ptr)
            last = det.fetch_results(n)
            for k in range(n_sets):
                i = max(j for j in range(n_batches) if j % n_sets == k)
                got_m, got_l = masks[k].get(), labels[k].get()
                for f in range(n):
                    ref = oracle.detect_contamination(batches[i][f][:, :, None], **okw)
                    assert np.array_equal(got_m[f], ref.mask), (n_sets, rep, k, f)
                    assert np.array_equal(got_l[f], ref.labels), (n_sets, rep, k, f)
            for f in range(n):
                ref = oracle.detect_contamination(batches[n_batches - 1][f][:, :, None], **okw)
                assert [((int(d["y"]), int(d["x"])), float(d["size"]), float(d["confidence"])) for d in last.defects_of(f)] == \
                    [(d["position"], d["size"], d["confidence"]) for d in ref.defects]
    finally:
        det.close()


def test_small_ccl_build_overflow_escalates(oracle):
    """A frame with more foreground than the small per-frame CCL build holds (2048 non-zero words) is flagged, finished by
    the global path, and the next batches go through the big build -- results identical throughout."""
    import heimdall_core as hc
    busy = synth.high_contamination_frame(512, 640, 3)  # 2333 non-zero words, 402 components: fits the big build only
    calm = synth.bottle_frame(512, 640, 4, contaminants=1)
    batch = np.stack([calm, busy])[..., None]
    det = hc.Detector(0, max_blobs_per_frame=100000, max_defects_per_frame=20000)
    try:
        for rep in range(3):
            res = det.detect_batch(batch, debug=["mask", "labels"])
            for f in range(2):
                ref = oracle.detect_contamination(batch[f])
                assert np.array_equal(res.debug["mask"][f], ref.mask) and np.array_equal(res.debug["labels"][f], ref.labels)
                assert [((int(d["y"]), int(d["x"])), float(d["size"])) for d in res.defects_of(f)] == \
                    [(d["position"], d["size"]) for d in ref.defects]
    finally:
        det.close()


# ---- BASELINE configs[2] and [3] at their stated sizes, against the oracle ---------------------------------------------
def _compare_full(res, ref, f=0):
    assert np.array_equal(res.debug["mask"][f], ref.mask), "mask"
    assert np.array_equal(res.debug["labels"][f], ref.labels), "labels"
    assert int(res.frames["n_components"][f]) == ref.ncomp
    got = [((int(d["y"]), int(d["x"])), float(d["size"]), float(d["confidence"])) for d in res.defects_of(f)]
    assert got == [(d["position"], d["size"], d["confidence"]) for d in ref.defects]
    assert bool(res.rejected[f]) == ref.reject


def test_12mp_high_contamination_matches_oracle(oracle):
    """configs[3]: one 4096x3000 frame with >= 10k blobs -- mask, label plane and defect list equal to the oracle's
    (the per-frame kernel flags the frame, the global-memory CCL kernels finish it)."""
    import heimdall_core as hc
    fr = synth.high_contamination_frame(3000, 4096, 1)
    ref = oracle.detect_contamination(fr[:, :, None])
    assert ref.ncomp >= 10000 and len(ref.defects) > 100
    for global_ccl in (False, True):
        det = hc.Detector(0, max_blobs_per_frame=400000, max_defects_per_frame=200000, global_ccl=global_ccl)
        try:
            _compare_full(det.detect_batch(fr, debug=["mask", "labels"]), ref)
        finally:
            det.close()


@pytest.mark.parametrize("gauss,morph", [((5, 0.0), 0), ((15, 3.0), 0), (None, 3), (None, 15), ((5, 0.0), 3)])
def test_5mp_gaussian_and_morphology_match_oracle(oracle, detector, gauss, morph):
    """configs[2]: a 2448x2048 frame through the Gaussian variants of K1 (k = 5 sigma 0, k = 15 sigma 3) and through
    open + close (3x3: folded into K1; 15x15: the tiles kernel), each compared with the oracle: mask, labels, defects."""
    import heimdall_core as hc
    fr = synth.bottle_frame(2048, 2448, 5, contaminants=3)
    kw, okw = {}, {}
    if gauss is not None:
        kw.update(blur_mode=hc._abi.HV_BLUR_GAUSSIAN, blur_ksize=gauss[0], gauss_sigma=gauss[1])
        okw.update(gauss_ksize=gauss[0], gauss_sigma=gauss[1])
    if morph:
        kw.update(morph_open_k=morph, morph_close_k=morph)
        okw.update(morph_open_k=morph, morph_close_k=morph)
    ref = oracle.detect_contamination(fr[:, :, None], **okw)
    res = detector.detect_batch(fr, hc.make_params(**kw), debug=["mask", "labels"])
    _compare_full(res, ref)


# ---- frames that are NOT rejected --------------------------------------------------------------------------------------
def test_near_threshold_and_empty_frames(oracle, detector):
    """The bottle frames are always rejected (the reference flags the bottle outline itself), so reject == False would only
    ever be seen on the two flat KATs.  Here: empty line positions with camera noise and an illumination gradient, with
    and without faint blemishes whose blurred contrast, area and confidence straddle c = 25, min_size = 10 and 0.3.
    One batch of 24 frames at 256x320, one at the headline resolution; both decisions must occur and agree."""
    frames = np.stack([synth.near_threshold_frame(256, 320, i, spots=(0 if i % 4 == 0 else 3)) for i in range(24)])
    res = detector.detect_batch(frames[..., None], debug=["mask", "labels"])
    refs = [oracle.detect_contamination(frames[f][:, :, None]) for f in range(len(frames))]
    for f, ref in enumerate(refs):
        _compare_full(res, ref, f)
    decisions = [r.reject for r in refs]
    assert 5 <= sum(decisions) <= 19, decisions       # a real mix of accepted and rejected frames
    assert any(r.ncomp > 0 and not r.reject for r in refs)   # components seen, none of them a defect
    big = np.stack([synth.near_threshold_frame(1024, 1280, 100 + i, spots=(0 if i == 0 else 4)) for i in range(4)])
    res = detector.detect_batch(big[..., None], debug=["mask", "labels"])
    for f in range(len(big)):
        _compare_full(res, oracle.detect_contamination(big[f][:, :, None]), f)
    assert not res.rejected[0]
    # min_confidence boundary: a defect whose confidence is exactly the threshold is kept (>=, detection.rs:298)
    import heimdall_core as hc
    f0 = next(f for f, r in enumerate(refs) if r.defects)
    c0 = refs[f0].defects[0]["confidence"]
    for mc, keep in ((c0, True), (np.nextafter(c0, 2.0), False)):
        r2 = detector.detect_batch(frames[f0], hc.make_params(min_confidence=mc))
        assert (c0 in [float(d["confidence"]) for d in r2.defects_of(0)]) == keep


# ---- streaming: every batch's results reach the host ---------------------------------------------------------------------
def test_streaming_tickets_deliver_every_batch(oracle):
    """hv_enqueue_device returns a ticket per batch; the batch's results travel to pinned host memory behind its last
    kernel (copy stream ordered by the slot's device-side counter) and hv_fetch_ticket returns them while later batches
    are already running.  Every one of 17 batches is fetched -- lagging the enqueue by depth - 1 -- and equals the oracle;
    tickets older than the pipeline depth are refused; the line statistics count every frame."""
    import heimdall_core as hc
    n, h, w = 3, 256, 384
    n_batches = 17
    batches = [synth.bottle_batch(n, h, w, start_index=9000 + 10 * i, contaminants=(i % 4)) for i in range(n_batches)]
    refs = [[oracle.detect_contamination(b[f][:, :, None], want_intermediates=False) for f in range(n)] for b in batches]
    det = hc.Detector(0)
    try:
        depth = det.pipeline_depth()
        d_in = [det.device_alloc((n, h, w), np.uint8, compressible=False) for _ in range(n_batches)]
        for a, bt in zip(d_in, batches):
            a.set(bt)
        masks = [det.device_alloc((n, h, w), np.uint8) for _ in range(depth)]
        labels = [det.device_alloc((n, h, w), np.int32) for _ in range(depth)]
        tickets, fetched = [], {}

        def fetch(i):
            fetched[i] = det.fetch(tickets[i], n)

        for i in range(n_batches):
            tickets.append(det.enqueue_device(d_in[i].ptr, n, h, w, 1, None, masks[i % depth].ptr, labels[i % depth].ptr))
            if i >= depth - 1:
                fetch(i - (depth - 1))
        assert tickets == sorted(set(tickets)) and all(t > 0 for t in tickets)
        for i in range(n_batches - (depth - 1), n_batches):
            fetch(i)
        for i in range(n_batches):
            for f in range(n):
                got = [((int(d["y"]), int(d["x"])), float(d["size"]), float(d["confidence"])) for d in fetched[i].defects_of(f)]
                assert got == [(d["position"], d["size"], d["confidence"]) for d in refs[i][f].defects], (i, f)
                assert bool(fetched[i].rejected[f]) == refs[i][f].reject
        with pytest.raises(hc.HeimdallCudaError) as ei:
            det.fetch(tickets[0], n)                   # its scratch set has long been reused
        assert ei.value.status == hc._abi.HV_ERR_BAD_TICKET
        again = det.fetch(tickets[-1], n)              # a batch can be fetched more than once while it is held
        assert np.array_equal(again.defects, fetched[n_batches - 1].defects)
        s = det.stats()
        assert s["frames_inspected"] == n * n_batches
        assert s["total_defects"] == sum(len(r.defects) for b in refs for r in b)
        assert s["frames_rejected"] == sum(r.reject for b in refs for r in b)
    finally:
        det.close()


def test_streaming_finishes_flagged_frames_without_fetch(oracle):
    """A frame too busy for the per-frame CCL kernel in the middle of a stream of batches nobody fetches: the slot is
    retired before it is reused -- the global-memory kernels finish the flagged frame in the caller's planes, the line
    statistics count it -- and the batches behind it switch to the bigger build (the selection state is updated at every
    retirement, not only at a fetch)."""
    import heimdall_core as hc
    n, h, w = 2, 512, 640
    busy = synth.high_contamination_frame(h, w, 3)   # 2333 non-zero words: beyond the small build
    calm = [synth.bottle_frame(h, w, 40 + i, contaminants=1) for i in range(12)]
    batches = [np.stack([calm[i], busy if i == 3 else calm[(i + 5) % 12]]) for i in range(12)]
    det = hc.Detector(0, max_blobs_per_frame=100000, max_defects_per_frame=20000)
    try:
        depth = det.pipeline_depth()
        d_in = [det.device_alloc((n, h, w), np.uint8, compressible=False) for _ in batches]
        for a, bt in zip(d_in, batches):
            a.set(bt)
        masks = [det.device_alloc((n, h, w), np.uint8) for _ in range(len(batches))]
        labels = [det.device_alloc((n, h, w), np.int32) for _ in range(len(batches))]
        tickets = [det.enqueue_device(d_in[i].ptr, n, h, w, 1, None, masks[i].ptr, labels[i].ptr) for i in range(len(batches))]
        last = det.fetch(tickets[-1], n)
        exp_defects = 0
        for i, bt in enumerate(batches):
            got_m, got_l = masks[i].get(), labels[i].get()
            for f in range(n):
                ref = oracle.detect_contamination(bt[f][:, :, None])
                exp_defects += len(ref.defects)
                assert np.array_equal(got_m[f], ref.mask), (i, f)
                assert np.array_equal(got_l[f], ref.labels), (i, f)
        s = det.stats()
        assert s["frames_inspected"] == n * len(batches) and s["total_defects"] == exp_defects
        assert len(last.defects_of(0)) == len(oracle.detect_contamination(batches[-1][0][:, :, None]).defects)
        assert depth >= 2
    finally:
        det.close()


@pytest.mark.parametrize("global_ccl", [False, True])
def test_deferred_tail_streaming_with_work_between_batches(oracle, global_ccl):
    """HV_FLAG_DEFER_TAIL: the per-frame kernel of a batch goes onto the stream with the next call.  A streaming loop that
    records an event between every two batches, synchronizes the whole device while a tail is held back, rotates one and
    several plane sets, contains a frame the per-frame kernel cannot hold, and finally closes a context with a tail still
    held back: every batch's records, the planes after hv_flush, and the line statistics equal the oracle's."""
    import torch

    import heimdall_core as hc
    n, h, w = 2, 512, 640
    busy = synth.high_contamination_frame(h, w, 3)
    calm = [synth.bottle_frame(h, w, 140 + i, contaminants=i % 3) for i in range(14)]
    batches = [np.stack([calm[i], busy if i == 6 else calm[(i + 5) % 14]]) for i in range(14)]
    refs = [[oracle.detect_contamination(b[f][:, :, None]) for f in range(n)] for b in batches]
    st = torch.cuda.current_stream().cuda_stream
    for n_sets in (1, 6):
        # (global_ccl: every batch through the global-memory kernels, which the flag puts onto the slots' own streams)
        det = hc.Detector(0, max_blobs_per_frame=100000, max_defects_per_frame=20000, defer_tail=True, global_ccl=global_ccl)
        try:
            det.set_stream(st)
            depth = det.pipeline_depth()
            d_in = [torch.from_numpy(b).cuda() for b in batches]
            masks = [det.device_alloc((n, h, w), np.uint8) for _ in range(n_sets)]
            labels = [det.device_alloc((n, h, w), np.int32) for _ in range(n_sets)]
            tickets, fetched, events = [], {}, []
            for i in range(len(batches)):
                tickets.append(det.enqueue_device(d_in[i].data_ptr(), n, h, w, 1, None, masks[i % n_sets].ptr, labels[i % n_sets].ptr))
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                events.append(ev)
                if i == 4:
                    torch.cuda.synchronize()     # (must not wait for a kernel that is not on the stream yet)
                if i >= depth - 1:
                    fetched[i - depth + 1] = det.fetch(tickets[i - depth + 1], n)
            det.flush()
            torch.cuda.synchronize()
            k = (len(batches) - 1) % n_sets
            got_m, got_l = masks[k].get(), labels[k].get()
            for f in range(n):
                assert np.array_equal(got_m[f], refs[-1][f].mask) and np.array_equal(got_l[f], refs[-1][f].labels)
            for i in range(len(batches) - depth + 1, len(batches)):
                fetched[i] = det.fetch(tickets[i], n)
            for i in range(len(batches)):
                for f in range(n):
                    got = [((int(d["y"]), int(d["x"])), float(d["size"]), float(d["confidence"])) for d in fetched[i].defects_of(f)]
                    assert got == [(d["position"], d["size"], d["confidence"]) for d in refs[i][f].defects], (n_sets, i, f)
                    assert int(fetched[i].frames["n_components"][f]) == refs[i][f].ncomp
            t = det.enqueue_device(d_in[0].data_ptr(), n, h, w, 1, None, masks[0].ptr, labels[0].ptr)
            s = det.stats()                      # counts the batch whose tail was still held back
            assert s["frames_inspected"] == n * (len(batches) + 1)
            assert s["total_defects"] == sum(len(r.defects) for b in refs for r in b) + sum(len(r.defects) for r in refs[0])
            det.enqueue_device(d_in[1].data_ptr(), n, h, w, 1, None, masks[0].ptr, labels[0].ptr)
            assert t > 0
        finally:
            det.close()                          # with a tail held back
    torch.cuda.synchronize()


def test_deferred_tail_long_kernels_on_side_streams(oracle):
    """HV_FLAG_DEFER_TAIL with batches whose kernels behind K1 are long -- morphology through the tiles kernels on a batch too
    big for the counter chain (> 16384 tiles), and dense frames through the global-memory CCL kernels: they run on the slot's
    own stream beside the next batch's K1.  Records of every batch and the planes after hv_flush equal the oracle's."""
    import torch

    import heimdall_core as hc
    st = torch.cuda.current_stream().cuda_stream
    cases = [("morph7", [synth.bottle_frame(2048, 2448, 610 + i, contaminants=2 + i) for i in range(2)], 13,
              dict(morph_open_k=7, morph_close_k=7)),
             ("morph7small", [synth.bottle_frame(512, 640, 650 + i, contaminants=1 + i) for i in range(2)], 3,
              dict(morph_open_k=7, morph_close_k=7)),   # (small batch: the flag prefers the slot's stream to the counter chain)
             ("dense", [synth.high_contamination_frame(768, 1024, 5 + i) for i in range(2)], 4, {})]
    for name, distinct, n, mk in cases:
        h, w = distinct[0].shape
        refs = [oracle.detect_contamination(d[:, :, None], **mk) for d in distinct]
        batch = np.stack([distinct[i % 2] for i in range(n)])
        det = hc.Detector(0, max_blobs_per_frame=200000, max_defects_per_frame=60000, defer_tail=True,
                          global_ccl=(name == "dense"))
        try:
            det.set_stream(st)
            depth = det.pipeline_depth()
            d_in = torch.from_numpy(batch).cuda()
            masks = [det.device_alloc((n, h, w), np.uint8) for _ in range(depth)]
            labels = [det.device_alloc((n, h, w), np.int32) for _ in range(depth)]
            params = hc.make_params(**mk)
            tickets, fetched = [], []
            steps = 9
            for i in range(steps):
                tickets.append(det.enqueue_device(d_in.data_ptr(), n, h, w, 1, params, masks[i % depth].ptr, labels[i % depth].ptr))
                torch.cuda.Event().record()
                if i >= depth - 1:
                    fetched.append(det.fetch(tickets[i - depth + 1], n))
            det.flush()
            torch.cuda.synchronize()
            k = (steps - 1) % depth
            for f in (0, 1, n - 1):
                assert np.array_equal(masks[k].get(f, 1)[0], refs[f % 2].mask), (name, f)
                assert np.array_equal(labels[k].get(f, 1)[0], refs[f % 2].labels), (name, f)
            for i in range(steps - depth + 1, steps):
                fetched.append(det.fetch(tickets[i], n))
            assert len(fetched) == steps
            for r in fetched:
                for f in range(n):
                    got = [((int(d["y"]), int(d["x"])), float(d["size"]), float(d["confidence"])) for d in r.defects_of(f)]
                    assert got == [(d["position"], d["size"], d["confidence"]) for d in refs[f % 2].defects], (name, f)
            assert det.stats()["frames_inspected"] == n * steps
        finally:
            det.close()


def test_enqueue_refuses_graph_capture():
    """The kernels of consecutive batches are ordered by device-side counters whose expected values are kernel arguments:
    replaying a captured graph would replay stale values.  hv_enqueue_device refuses a capturing stream."""
    import torch

    import heimdall_core as hc
    n, h, w = 2, 128, 256
    det = hc.Detector(0)
    try:
        d_in = torch.from_numpy(synth.bottle_batch(n, h, w, start_index=1)).cuda()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            det.set_stream(side.cuda_stream)
            det.enqueue_device(d_in.data_ptr(), n, h, w)          # fine outside a capture
            g = torch.cuda.CUDAGraph()
            g.capture_begin(capture_error_mode="relaxed")
            try:
                with pytest.raises(ValueError, match="cannot be captured"):
                    det.enqueue_device(d_in.data_ptr(), n, h, w)
            finally:
                g.capture_end()
            det.enqueue_device(d_in.data_ptr(), n, h, w)          # and fine again afterwards
            assert det.fetch_results(n).frames.shape == (n,)
        torch.cuda.synchronize()
    finally:
        det.set_stream(None)
        det.close()


def test_ccl_path_selection_follows_the_frames(oracle):
    """Dense frames move the context to the global-memory CCL kernels (after one batch that the per-frame kernel flags);
    it stays there -- no blind re-tries, which would cost such batches twice -- until the results of a batch show sparse
    frames again, and then returns to the per-frame kernel.  Results equal the oracle's throughout."""
    import heimdall_core as hc
    dense = np.stack([synth.high_contamination_frame(768, 1024, s) for s in (1, 2)])
    sparse = synth.bottle_batch(2, 768, 1024, start_index=60, contaminants=1)
    det = hc.Detector(0, max_blobs_per_frame=100000, max_defects_per_frame=30000, profile=True)
    try:
        def run(batch):
            det.profile()
            res = det.detect_batch(batch[..., None], debug=["labels"])
            for f in range(len(batch)):
                ref = oracle.detect_contamination(batch[f][:, :, None])
                assert np.array_equal(res.debug["labels"][f], ref.labels)
                assert [((int(d["y"]), int(d["x"])), float(d["size"])) for d in res.defects_of(f)] == \
                    [(d["position"], d["size"]) for d in ref.defects]
            p = det.profile()
            return p["ccl_frame_fused"]["launches"], p["ccl_merge"]["launches"]
        assert run(sparse) == (1, 0)
        f1 = run(dense)
        assert f1[1] == 1                       # flagged (small and/or big build) and finished by the global kernels
        for _ in range(9):
            assert run(dense) in ((1, 1), (0, 1))
        assert run(dense) == (0, 1)             # settled on the global path: no probing of the per-frame kernel
        assert run(sparse) == (0, 1)            # still there for one batch: its results say "sparse" ...
        assert run(sparse) == (1, 0)            # ... and the per-frame kernel is back
    finally:
        det.close()
