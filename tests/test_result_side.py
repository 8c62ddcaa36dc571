"""Result side: InspectionResult-shaped export, dashboard statistics, overlays."""
import os
import sys

import numpy as np
import pytest

import synth

REFERENCE = os.environ.get("HEIMDALL_REFERENCE_ROOT", "/root/reference")


def _batch_result():
    from heimdall_core import batch as B
    fr = np.zeros(4, B.RESULT_DTYPE)
    fr["n_defects"] = [0, 2, 1, 0]
    fr["defects_offset"] = [0, 0, 2, 3]
    fr["status"] = [0, 0, 0, -4]
    df = np.zeros(3, B.DEFECT_DTYPE)
    df["y"], df["x"] = [10, 20, 30], [11, 21, 31]
    df["size"], df["confidence"] = [12.0, 250.0, 31.0], [0.5, 0.8125, 0.7999999999999999]
    df["ymin"], df["xmin"], df["ymax"], df["xmax"] = [8, 15, 28], [9, 16, 29], [12, 25, 32], [13, 26, 33]
    return B.BatchResult(fr, df, 0)


def test_export_matches_the_dashboard_update_rule():
    """hv_export_results against a literal restatement of dashboard.py:483-500 (per image: counters, EMA with the first
    sample taken as is, defect rate in per cent) -- exact f64 equality."""
    from heimdall_core import results as R
    res = _batch_result()
    st = R.DashboardStats(start_time=123.0)
    stats = {"total_images": 0, "total_defects": 0, "avg_processing_time": 0, "defect_rate": 0, "start_time": 123.0}
    seq = 0
    for rep, pt in enumerate((0.0031, 0.0007, 0.01234)):
        out = R.export_batch(res, timestamp=1000.5 + rep, processing_time=pt, first_sequence=seq, stats=st)
        seq += len(out)
        for f in range(len(res.frames)):
            n_def = int(res.frames["n_defects"][f])
            stats["total_images"] += 1
            stats["total_defects"] += n_def
            if stats["avg_processing_time"] == 0:
                stats["avg_processing_time"] = pt * 1000
            else:
                stats["avg_processing_time"] = (0.9 * stats["avg_processing_time"] + 0.1 * pt * 1000)
            if stats["total_images"] > 0:
                stats["defect_rate"] = (stats["total_defects"] / stats["total_images"] * 100)
        assert st.as_dict() == stats
        assert [o.inspection_id for o in out] == [f"contamination_{seq - 4 + f}" for f in range(4)]
        assert [o.has_defects for o in out] == [False, True, True, False]
        assert [o.success for o in out] == [True, True, True, False]          # a frame whose status is an error
        assert out[1].defects[1].position == (20, 21) and out[2].defects[0].confidence == 0.7999999999999999
        assert out[1].processing_time == pt and out[1].timestamp == 1000.5 + rep


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "heimdall")), reason="reference checkout not present")
def test_export_equals_the_reference_classes():
    """The mirrors give the same to_dict() as the reference's own InspectionResult / Defect fed with the same values
    (heimdall/inspection/base_inspector.py:11-64, heimdall/detectors/base.py:7-38), imported unmodified."""
    import subprocess
    import json
    code = r'''
import json, sys
sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(0, %r)
import logging; logging.disable(logging.CRITICAL)
from test_result_side import _batch_result
from heimdall_core import results as R
from heimdall.inspection.base_inspector import InspectionResult
from heimdall.detectors.base import Defect
out = R.export_batch(_batch_result(), timestamp=7.25, processing_time=0.004, first_sequence=40)
same = True
for o in out:
    ref = InspectionResult(o.inspection_id, o.timestamp, o.success,
                           [Defect(d.defect_type, d.position, d.size, d.confidence, dict(d.metadata)) for d in o.defects],
                           metadata=dict(o.metadata))
    same &= json.dumps(ref.to_dict(), sort_keys=True) == json.dumps(o.to_dict(), sort_keys=True)
    same &= ref.has_defects == o.has_defects and ref.defect_count == o.defect_count and str(ref) == str(o)
    same &= [str(a) for a in ref.defects] == [str(b) for b in o.defects]
print(json.dumps({"same": bool(same), "n": len(out)}))
''' % (os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "heimdall-vision_b200"), REFERENCE,
       os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert json.loads(r.stdout.strip().splitlines()[-1]) == {"same": True, "n": 4}


def _numpy_overlays(img, items):
    """Sequential restatement: crosses as processing.rs:371-401, boxes as a 1-px outline, in list order."""
    out = img.copy()
    h, w = out.shape[:2]
    for it in items:
        kind, col = it[0], it[-1]
        if kind == 0:
            y, x = it[1], it[2]
            for dy in range(-3, 4):
                if 0 <= y + dy < h and 0 <= x < w:
                    out[y + dy, x] = col
            for dx in range(-3, 4):
                if 0 <= y < h and 0 <= x + dx < w:
                    out[y, x + dx] = col
        elif kind == 1:
            ya, yb, xa, xb = min(it[1], it[3]), max(it[1], it[3]), min(it[2], it[4]), max(it[2], it[4])
            for yy in (ya, yb):
                if 0 <= yy < h:
                    out[yy, max(xa, 0):max(min(xb + 1, w), 0)] = col
            for xx in (xa, xb):
                if 0 <= xx < w:
                    out[max(ya, 0):max(min(yb + 1, h), 0), xx] = col
    return out


@pytest.mark.gpu
def test_overlays_match_opencv_golden_and_the_rust_crosses(oracle, detector, golden_dir):
    from heimdall_core import results as R
    z = np.load(os.path.join(golden_dir, "cv2_overlays.npz"))
    items = [(int(k), int(y), int(x), int(y1), int(x1), (int(b), int(g), int(r))) if k == 1 else (int(k), int(y), int(x), (int(b), int(g), int(r)))
             for k, y, x, y1, x1, b, g, r in z["items"]]
    got = R.draw_overlays(z["base"], items, detector)
    assert np.array_equal(got, z["out"])                 # boxes and markers, overlapping, in cv2's drawing order
    # crosses: the visualisation of process_image("contamination") is the mask replicated + crosses at the centres
    fr = synth.bottle_frame(200, 300, 9, contaminants=2)
    img3 = np.dstack([fr, fr, fr])
    vis, contours = oracle.contamination_pipeline(img3)
    res = detector.detect_batch(img3, __import__("heimdall_core").make_params(3.0, 1e18, 15.0, min_confidence=-1.0), debug=["mask"])
    mask3 = np.repeat(res.debug["mask"][0][:, :, None], 3, axis=2)
    centres = [(R.CROSS, cy, cx) for cy, cx, _ in contours]
    assert np.array_equal(R.draw_overlays(mask3, centres, detector), vis)
    # clipping, ordering and colours against the sequential numpy restatement
    rng = np.random.default_rng(3)
    base = rng.integers(0, 256, (40, 57, 3), dtype=np.uint8)
    its = []
    for i in range(60):
        col = tuple(int(c) for c in rng.integers(0, 256, 3))
        if i % 2:
            its.append((R.CROSS, int(rng.integers(-4, 44)), int(rng.integers(-4, 61)), col))
        else:
            y, x = int(rng.integers(-5, 45)), int(rng.integers(-5, 62))
            its.append((R.BOX, y, x, y + int(rng.integers(-20, 20)), x + int(rng.integers(-20, 20)), col))
    assert np.array_equal(R.draw_overlays(base, its, detector), _numpy_overlays(base, its))
    assert np.array_equal(R.draw_overlays(base, [], detector), base)
    assert np.array_equal(R.draw_overlays(base, its, detector), _numpy_overlays(base, its))   # the claim plane was released


@pytest.mark.gpu
def test_export_of_a_real_batch(oracle, detector):
    from heimdall_core import results as R
    batch = np.stack([synth.bottle_frame(256, 320, 5, contaminants=2), synth.near_threshold_frame(256, 320, 0, spots=0)])
    res = detector.detect_batch(batch[..., None])
    st = R.DashboardStats(0.0)
    out = R.export_batch(res, timestamp=1.0, processing_time=1e-4, stats=st)
    for f, o in enumerate(out):
        ref = oracle.detect_contamination(batch[f][:, :, None])
        assert o.has_defects == ref.reject == bool(res.rejected[f])
        assert [(d.position, d.size, d.confidence) for d in o.defects] == [(d["position"], d["size"], d["confidence"]) for d in ref.defects]
        assert [d.metadata["bounding_box"] for d in o.defects] == \
            [(d["bbox"][1], d["bbox"][0], d["bbox"][3] - d["bbox"][1] + 1, d["bbox"][2] - d["bbox"][0] + 1) for d in ref.defects]
    assert st.as_dict()["total_images"] == 2 and st.as_dict()["total_defects"] == len(out[0].defects)
    assert not out[1].has_defects
    vis = R.visualize_defects(batch[0], res, 0, detector=detector)
    assert vis.shape == (256, 320, 3) and (vis != np.repeat(batch[0][:, :, None], 3, 2)).any()
