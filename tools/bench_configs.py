#!/usr/bin/env python3
"""Throughput of the BASELINE.json configs other than the headline one (configs[2..4]); the headline config is
bench.py's.  Device-resident inputs, CUDA events on the launching stream, steady state (hv_pipeline_depth() sets of output
planes in rotation).
Every case is checked against the oracle on one frame before it is timed (measurement infrastructure like bench.py: the
oracle is the checker here, never the thing measured).  Writes a markdown table.

usage: bench_configs.py [out.md] [--quick] [--only=<substring of a row name>]...
environment: BENCH_SETS (6) sets of output planes in rotation; BENCH_DEFER=1 creates every context with HV_FLAG_DEFER_TAIL
(the rows that name the flag always do)
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heimdall-vision_b200"))
sys.path.insert(0, ROOT)
import heimdall_core as hc  # noqa: E402
import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

PEAK = 6546.6
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
quick = "--quick" in sys.argv
only = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--only=")]  # run only the configs whose name contains one
out_md = next((a for a in sys.argv[1:] if not a.startswith("--")), None)
st = torch.cuda.current_stream().cuda_stream
rows = []


def tile_batch(distinct, n):
    """n frames from a few distinct ones (generation cost), each copy with its own +-1 noise so that bytes differ."""
    rng = np.random.default_rng(7)
    out = np.empty((n,) + distinct[0].shape, np.uint8)
    for i in range(n):
        out[i] = np.clip(distinct[i % len(distinct)].astype(np.int16) + rng.integers(-1, 2, distinct[0].shape), 0, 255)
    return out


NSETS = int(os.environ.get("BENCH_SETS", "6"))  # output-plane sets in rotation (= hv_pipeline_depth(): no batch waits for another's planes)


def run(name, batch, params, okw, steps, check_frames=(0,), defer=False):
    if only and not any(o in name for o in only):
        return True
    n, h, w = batch.shape
    det = hc.Detector(0, max_defects_per_frame=512 if h * w < 8_000_000 else 32768, defer_tail=defer or os.environ.get("BENCH_DEFER", "0") == "1")
    det.set_stream(st)
    d_in = torch.from_numpy(batch).cuda()
    outs = [(torch.empty((n, h, w), dtype=torch.uint8, device="cuda"), torch.empty((n, h, w), dtype=torch.int32, device="cuda"))
            for _ in range(NSETS)]
    res = det.detect_device(d_in.data_ptr(), n, h, w, 1, params, outs[0][0].data_ptr(), outs[0][1].data_ptr())
    ok = True
    for f in check_frames:
        ref = O.detect_contamination(batch[f][:, :, None], **okw)
        ok &= np.array_equal(outs[0][0][f].cpu().numpy(), ref.mask) and np.array_equal(outs[0][1][f].cpu().numpy(), ref.labels)
        ok &= [((int(d["y"]), int(d["x"])), float(d["size"]), float(d["confidence"])) for d in res.defects_of(f)] == \
            [(d["position"], d["size"], d["confidence"]) for d in ref.defects]
    l0 = det.launch_count()

    def step(i):
        det.enqueue_device(d_in.data_ptr(), n, h, w, 1, params, outs[i % NSETS][0].data_ptr(), outs[i % NSETS][1].data_ptr())
    for i in range(2 * det.pipeline_depth() + 1):  # every scratch slot of the library has seen this shape (first use allocates)
        step(i)
    torch.cuda.synchronize()
    l0 = det.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = (det.launch_count() - l0) / steps
    last = det.fetch_results(n)
    # per-kernel times of the same steps (a separate pass: the events serialise the overlap)
    det.profile_enable(None)
    for i in range(steps):
        step(i)
    torch.cuda.synchronize()
    prof = {k: round(v["ms"] / steps, 4) for k, v in det.profile().items() if v["launches"]}
    det.profile_enable([])
    print("   per-kernel ms/step:", prof, flush=True)
    fps = n / ms * 1e3
    gbs = 6.0 * n * h * w / ms / 1e6
    rows.append((name, f"{n} x {w}x{h}", "yes" if ok else "NO", f"{ms:.3f}", f"{fps:,.0f}", f"{gbs:,.0f}", f"{gbs / PEAK:.2f}",
                 f"{launches:.0f}", f"{int(last.frames['n_components'].sum()) / n:,.0f}",
                 f"{int(last.frames['n_defects'].sum()) / n:.1f}"))
    print(rows[-1], flush=True)
    det.close()
    del d_in, outs
    torch.cuda.empty_cache()
    return ok


t0 = time.time()
# ---- configs[2]: 5 MP 2448x2048 u8, batch of 256, blur sigma sweep and morphology kernel 3..15 ----------------------
n5 = 32 if quick else 256
five = tile_batch([synth.bottle_frame(2048, 2448, 500 + i, contaminants=i % 4) for i in range(8)], n5)
run("C3 5 MP, reference-exact (box 5x5, no morphology)", five, hc.make_params(), {}, 4 if quick else 10)
for sig in (0.0, 1.0, 2.0, 3.0):
    k = 5 if sig == 0 else min(2 * int(np.ceil(3 * sig)) + 1, 15)
    run(f"C3 5 MP, Gaussian k={k} sigma={sig:g}", five,
        hc.make_params(blur_mode=hc._abi.HV_BLUR_GAUSSIAN, blur_ksize=k, gauss_sigma=sig), dict(gauss_ksize=k, gauss_sigma=sig),
        3 if quick else 6)
for k in (3, 7, 11, 15):
    run(f"C3 5 MP, morphology open+close k={k}", five, hc.make_params(morph_open_k=k, morph_close_k=k),
        dict(morph_open_k=k, morph_close_k=k), 3 if quick else 6)
    if k > 3:  # (tiles kernels behind K1: with the flag they run on the slots' streams beside the next batch's K1)
        run(f"C3 5 MP, morphology open+close k={k}, HV_FLAG_DEFER_TAIL", five, hc.make_params(morph_open_k=k, morph_close_k=k),
            dict(morph_open_k=k, morph_close_k=k), 3 if quick else 6, defer=True)
del five
# ---- configs[3]: 12 MP 4096x3000 high-contamination frames (10k+ blobs per frame) -------------------------------------
n12 = 4 if quick else 16
twelve = tile_batch([synth.high_contamination_frame(3000, 4096, i) for i in range(2)], n12)
run("C4 12 MP, >10k blobs/frame (global-memory CCL path)", twelve, hc.make_params(), {}, 3 if quick else 6)
run("C4 12 MP, >10k blobs/frame, HV_FLAG_DEFER_TAIL", twelve, hc.make_params(),
    {}, 3 if quick else 6, defer=True)
del twelve
# ---- configs[4] on one GPU: 8 camera streams x 5 MP, one batch per stream round (bench.py --gpus N shards streams) ----
streams = tile_batch([synth.bottle_frame(2048, 2448, 900 + i, contaminants=i % 3) for i in range(8)], 8)
run("C5 8 streams x 5 MP, 1 frame per stream per step (1 GPU)", streams, hc.make_params(), {}, 10 if quick else 40)
streams4 = tile_batch([synth.bottle_frame(2048, 2448, 900 + i, contaminants=i % 3) for i in range(8)], 32)
run("C5 8 streams x 5 MP, 4 frame sets per step (FrameSet batcher, sets_per_batch=4; 1 GPU)", streams4, hc.make_params(), {},
    10 if quick else 40)

# ---- configs[0]: one 1280x1024 frame through the drop-in module call (host array in, Python dict out) -------------------
lat_note = ""
if not only or any("C1" in o for o in only):
    one = synth.bottle_frame(1024, 1280, 1234, contaminants=2)[:, :, None]
    ref1 = O.detect_contamination(one)
    out = hc.detect_contamination(one)
    ok1 = [(d["position"], d["size"], d["confidence"]) for d in out["defects"]] == \
        [(d["position"], d["size"], d["confidence"]) for d in ref1.defects]
    for _ in range(20):
        hc.detect_contamination(one)
    reps = 200
    t1 = time.perf_counter()
    for _ in range(reps):
        hc.detect_contamination(one)
    lat_us = (time.perf_counter() - t1) / reps * 1e6
    t1 = time.perf_counter()
    for _ in range(3):
        O.detect_contamination(one, want_intermediates=False)
    cpu_us = (time.perf_counter() - t1) / 3 * 1e6
    lat_note = (f"\nC1 (configs[0]): one 1280x1024 frame through `heimdall_core.detect_contamination` (pageable numpy array in, "
                f"list of dicts out, synchronous): {lat_us:.0f} us per call, parity {'yes' if ok1 else 'NO'}; the oracle port of "
                f"the reference's CPU path on one host core: {cpu_us / 1e3:.1f} ms per frame.\n")
    print(lat_note, flush=True)

hdr = ("config", "batch", "oracle parity", "ms/step", "frames/s", "alg GB/s", "frac of measured HBM", "launches/step",
       "components/frame", "defects/frame")
text = "| " + " | ".join(hdr) + " |\n|" + "---|" * len(hdr) + "\n" + "\n".join("| " + " | ".join(r) + " |" for r in rows)
text = (f"BASELINE.json configs[2..4] on one B200 (device-resident inputs, CUDA events, outputs rotate over {NSETS} "
        f"buffer sets; algorithmic bytes = 6 B/px; measured HBM peak {PEAK:.1f} GB/s).  Generated by tools/bench_configs.py"
        f"{' --quick' if quick else ''} in {time.time() - t0:.0f} s.\n\n" + text + "\n" + lat_note)
print(text)
if out_md:
    open(out_md, "w").write(text)
