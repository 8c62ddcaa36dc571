"""A small run of every hot kernel for compute-sanitizer (memcheck / racecheck / synccheck): the TMA preprocess kernel
(plain, Gaussian and morphology variants), the per-frame CCL kernel (small and big builds, overflow -> global path), the
global-memory CCL kernels on adversarial masks, the morphology tiles kernels, a few batches in flight with rotating output
sets.  Results are compared with the oracle so that a sanitizer-clean run is also a correct one.
usage: compute-sanitizer --tool <tool> python tools/sanitize_case.py"""
import sys
import numpy as np
sys.path.insert(0, 'heimdall-vision_b200'); sys.path.insert(0, '.')
import heimdall_core as hc, synth
from oracle import oracle as O

def same(res, f, ref):
    got = [((int(d["y"]), int(d["x"])), float(d["size"]), float(d["confidence"])) for d in res.defects_of(f)]
    return got == [(d["position"], d["size"], d["confidence"]) for d in ref.defects]

ok = True
det = hc.Detector(0, max_blobs_per_frame=100000, max_defects_per_frame=4096)
n, h, w = 3, 192, 384
batches = [synth.bottle_batch(n, h, w, start_index=50 + 10 * i, contaminants=2) for i in range(7)]
# streaming, 2 output sets, plain and morphology (3x3 folded into K1; 7x7 tiles kernels)
for params, okw in ((None, {}), (hc.make_params(morph_open_k=3, morph_close_k=3), dict(morph_open_k=3, morph_close_k=3)),
                    (hc.make_params(morph_open_k=7, morph_close_k=7), dict(morph_open_k=7, morph_close_k=7))):
    d_in = [det.device_alloc((n, h, w), np.uint8, False) for _ in batches]
    for a, b in zip(d_in, batches): a.set(b)
    masks = [det.device_alloc((n, h, w), np.uint8) for _ in range(2)]
    labels = [det.device_alloc((n, h, w), np.int32) for _ in range(2)]
    tickets = [det.enqueue_device(d_in[i].ptr, n, h, w, 1, params, masks[i % 2].ptr, labels[i % 2].ptr) for i in range(len(batches))]
    last = det.fetch(tickets[-1], n)
    i = len(batches) - 1
    gm, gl = masks[i % 2].get(), labels[i % 2].get()
    for f in range(n):
        ref = O.detect_contamination(batches[i][f][:, :, None], **okw)
        ok &= np.array_equal(gm[f], ref.mask) and np.array_equal(gl[f], ref.labels) and same(last, f, ref)
    for a in d_in + masks + labels: a.free()
print("streaming ok", ok)
# overflow of the small build -> big build -> global path
busy = synth.high_contamination_frame(256, 640, 3)
calm = synth.bottle_frame(256, 640, 4, contaminants=1)
for rep in range(2):
    res = det.detect_batch(np.stack([calm, busy])[..., None], debug=["labels"])
    for f, fr in enumerate((calm, busy)):
        ref = O.detect_contamination(fr[:, :, None])
        ok &= np.array_equal(res.debug["labels"][f], ref.labels) and same(res, f, ref)
print("escalation ok", ok)
# Gaussian variant of K1
fr = synth.bottle_frame(160, 256, 9, contaminants=2)
res = det.detect_batch(fr, hc.make_params(blur_mode=hc._abi.HV_BLUR_GAUSSIAN, blur_ksize=5, gauss_sigma=0.0), debug=["mask"])
ok &= np.array_equal(res.debug["mask"][0], O.detect_contamination(fr[:, :, None], gauss_ksize=5, gauss_sigma=0.0).mask)
# global-memory CCL kernels on adversarial masks
yy, xx = np.mgrid[0:64, 0:96]
for m in ((((yy + xx) & 1) * 255).astype(np.uint8), np.full((64, 96), 255, np.uint8)):
    recs, lab = det.find_contours(np.ascontiguousarray(m[:, :, None]), 0.0, 1e18)
    ok &= np.array_equal(lab, O.label4(m, fg_gt127=True)[0])
print("all ok", ok)
det.close()
sys.exit(0 if ok else 1)
