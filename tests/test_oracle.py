"""CPU tests of the oracle (tests may use oracle/; the product never does).

The reference has no tests or golden vectors for this path, so the oracle is pinned by
  * the hand-derived known-answer tests of SURVEY.md 8c (KAT1-KAT5),
  * an independent literal Python transcription of detection.rs (tests/ref_literal.py),
  * OpenCV as an independent labelling implementation and as the spec of the extension stages (committed cv2 outputs),
  * regression vectors on the reference's own fixture images.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import ref_literal

H, W = 1024, 1280


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_kat1_square_20(oracle):
    img = np.full((H, W, 1), 220, np.uint8)
    img[500:520, 600:620] = 40
    r = oracle.detect_contamination(img)
    assert r.blur[510, 596:604].tolist() == [220, 220, 184, 148, 112, 76, 40, 40]
    assert int((r.mask == 255).sum()) == 208 and r.ncomp == 1
    assert len(r.defects) == 1
    d = r.defects[0]
    assert d["position"] == (509, 609) and d["size"] == 208.0 and d["bbox"] == (500, 600, 519, 619)
    assert d["confidence"] == 0.7 + 0.3 * (1 - 208 / 400) == 0.844
    assert r.reject


def test_kat2_square_4(oracle):
    img = np.full((H, W, 1), 220, np.uint8)
    img[500:504, 600:604] = 40
    r = oracle.detect_contamination(img)
    assert int((r.mask == 255).sum()) == 24 and r.ncomp == 1
    d = r.defects[0]
    assert d["position"] == (501, 601) and d["size"] == 24.0 and d["bbox"] == (499, 599, 504, 604)
    assert repr(d["confidence"]) == "0.7999999999999999"


def test_kat3_uniform_and_kat4_ramp(oracle):
    r = oracle.detect_contamination(np.full((H, W, 1), 220, np.uint8))
    assert int(r.mask.sum()) == 0 and r.defects == [] and not r.reject
    ramp = ((np.arange(W) * 255) // W).astype(np.uint8)
    r = oracle.detect_contamination(np.tile(ramp, (H, 1))[:, :, None])
    assert int(r.mask.sum()) == 0 and not r.reject


KAT5 = {1, 2, 4, 8, 11, 13, 16, 22, 26, 27, 32, 39, 44, 51, 52, 54, 57, 59, 61, 64, 78, 81, 88, 91, 95, 101, 102, 104, 108,
        114, 115, 118, 122, 127, 128, 143, 149, 156, 157, 162, 169, 175, 176, 182, 183, 190, 195, 202, 203, 204, 208, 209,
        216, 223, 225, 228, 230, 235, 236, 239, 241, 244, 249, 251, 254}


def test_kat5_gray_f64_truncation(oracle):
    v = np.arange(256, dtype=np.uint8)
    g = oracle.gray(np.stack([v, v, v], -1)[None])[0]
    assert {int(i) for i in v if int(g[i]) == int(i) - 1} == KAT5
    assert all(int(g[i]) == (i - 1 if i in KAT5 else i) for i in range(256))
    # pure python doubles agree (no FMA anywhere)
    assert all(ref_literal.f64_as_u8(0.299 * float(i) + 0.587 * float(i) + 0.114 * float(i)) == int(g[i])
               for i in range(256))


def test_f64_as_i32(oracle):
    for v, e in [(25.0, 25), (25.9, 25), (-25.9, -25), (float("nan"), 0), (1e20, 2 ** 31 - 1), (-1e20, -2 ** 31),
                 (float("inf"), 2 ** 31 - 1)]:
        assert oracle.f64_as_i32(v) == e == ref_literal.f64_as_i32(v)


@pytest.mark.parametrize("shape,c,seed", [((23, 31), 1, 0), ((17, 40), 3, 1), ((12, 12), 1, 2), ((4, 9), 1, 3),
                                          ((9, 4), 3, 4), ((1, 1), 1, 5), ((30, 11), 1, 6), ((40, 45), 1, 7)])
def test_oracle_equals_literal_python(oracle, shape, c, seed):
    rng = np.random.default_rng(seed)
    h, w = shape
    base = rng.integers(150, 256, (h, w, c), dtype=np.uint8)
    dark = rng.random((h, w)) < 0.08
    base[dark] = rng.integers(0, 80, (int(dark.sum()), c), dtype=np.uint8)
    if h > 14 and w > 14:
        base[5:12, 4:13] = 20
    for thr, mn, mx in [(25.0, 10.0, 3000.0), (5.0, 1.0, 50.0), (-3.0, 2.0, 1e9)]:
        a = oracle.detect_contamination(base, mn, mx, thr)
        b = ref_literal.detect_contamination(base, mn, mx, thr)
        assert np.array_equal(a.gray, np.array(b["gray"], np.uint8))
        assert np.array_equal(a.blur, np.array(b["blurred"], np.uint8))
        assert np.array_equal(a.mask, np.array(b["binary"], np.uint8))
        assert np.array_equal(a.labels, np.array(b["labels"], np.int32))
        assert a.ncomp == b["ncomp"]
        assert [(d["position"], d["size"], d["confidence"], d["label"]) for d in a.defects] == \
               [(d["position"], d["size"], d["confidence"], d["label"]) for d in b["defects"]]


def test_labels_match_opencv_golden(oracle, golden_dir):
    z = np.load(os.path.join(golden_dir, "cv2_ccl.npz"))
    for i in range(4):
        lab, blobs = oracle.label4(z[f"mask{i}"])
        assert np.array_equal(lab, z[f"labels{i}"])
        st = z[f"stats{i}"]
        assert np.array_equal(blobs["area"], st[:, 4])
        assert np.array_equal(blobs["xmin"], st[:, 0]) and np.array_equal(blobs["ymin"], st[:, 1])
        assert np.array_equal(blobs["xmax"] - blobs["xmin"] + 1, st[:, 2])
        assert np.array_equal(blobs["ymax"] - blobs["ymin"] + 1, st[:, 3])


def test_labels_match_opencv_live(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    for p in (0.05, 0.4, 0.6, 0.9):
        m = (rng.random((120, 150)) < p).astype(np.uint8) * 255
        lab, blobs = oracle.label4(m)
        n, cl, st, _ = cv2.connectedComponentsWithStats(m, connectivity=4, ltype=cv2.CV_32S)
        assert np.array_equal(lab, cl) and len(blobs) == n - 1


def test_gaussian_matches_opencv_golden(oracle, golden_dir):
    z = np.load(os.path.join(golden_dir, "cv2_gaussian.npz"))
    for key in z.files:
        if key in ("src", "src2"):
            continue
        which, k, s = key.split("_")
        src = z["src"] if which == "a" else z["src2"]
        out = oracle.gaussian_blur(src, int(k[1:]), float(s[1:]))
        assert np.array_equal(out, z[key]), key  # bit-exact, tighter than the +-1 LSB the north star allows


def test_morphology_matches_opencv_golden(oracle, golden_dir):
    z = np.load(os.path.join(golden_dir, "cv2_morph.npz"))
    ops = {"erode": oracle.MORPH_ERODE, "dilate": oracle.MORPH_DILATE, "open": oracle.MORPH_OPEN,
           "close": oracle.MORPH_CLOSE}
    for key in z.files:
        if key in ("m1", "m2"):
            continue
        which, name, k = key.split("_")
        assert np.array_equal(oracle.morph(z[which], ops[name], int(k)), z[key]), key


def test_reference_fixture_regression(oracle, golden_dir):
    frames = np.load(os.path.join(golden_dir, "fixture_frames.npz"))
    exp = json.load(open(os.path.join(golden_dir, "rust_path.json")))
    for name, e in exp.items():
        img = frames[name]
        assert sha(img) == e["sha256"]
        r = oracle.detect_contamination(img)
        assert int((r.mask == 255).sum()) == e["fg_pixels"] and r.ncomp == e["ncomp"]
        assert sha(r.mask) == e["mask_sha256"] and sha(r.labels) == e["labels_sha256"]
        assert [[d["position"][0], d["position"][1], d["size"], d["confidence"]] for d in r.defects] == e["defects"]
    # the numbers recorded at survey time
    assert [exp[f"contaminated_{i}"]["fg_pixels"] for i in (1, 2, 3)] == [3541, 3720, 3734]
    assert [exp[f"contaminated_{i}"]["ncomp"] for i in (1, 2, 3)] == [35, 35, 37]
    assert all(len(exp[f"contaminated_{i}"]["defects"]) == 11 for i in (1, 2, 3))


def test_pipelines_and_utilities(oracle):
    rng = np.random.default_rng(5)
    img = rng.integers(180, 256, (40, 52, 3), dtype=np.uint8)
    img[10:20, 12:25] = 15
    vis, contours = oracle.contamination_pipeline(img)
    assert vis.shape == (40, 52, 3) and all(c[2] == 0.75 for c in contours) and len(contours) >= 1
    cy, cx, _ = contours[0]
    assert tuple(vis[cy, cx]) == (0, 0, 255)
    b = oracle.basic_pipeline(img)
    assert set(np.unique(b)) <= {0, 255} and np.array_equal(b[..., 0], b[..., 2])
    with pytest.raises(ValueError):
        oracle.contamination_pipeline(img[:, :, :1])
    pre = oracle.preprocess_image(img, True, 5)
    assert np.array_equal(pre[:, :, 0], oracle.box_blur(oracle.gray(img), 2))
    assert np.array_equal(oracle.preprocess_image(img, False, None), img)
    g1 = pre
    t = oracle.apply_threshold(g1, adaptive=True, inverse=True)
    assert np.array_equal(t[:, :, 0], oracle.adaptive_threshold(g1[:, :, 0], 2, True))
    t2 = oracle.apply_threshold(g1, 100, False, False)
    assert np.array_equal(t2[:, :, 0], (g1[:, :, 0] > 100).astype(np.uint8) * 255)
    with pytest.raises(ValueError):
        oracle.apply_threshold(img)
    m = np.zeros((30, 30, 1), np.uint8)
    m[2:6, 3:9] = 200
    m[10:25, 10:25] = 255
    fc = oracle.find_contours(m)
    assert [c["pixel_count"] for c in fc] == [24, 225] and "points" in fc[0] and "points" not in fc[1]
    assert fc[0]["points"][0] == (2, 3) and len(fc[0]["points"]) == 24
