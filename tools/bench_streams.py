#!/usr/bin/env python3
"""BASELINE.json configs[4]: a sustained stream of 8 cameras x 5 MP (2448x2048 Mono8) sharded across 1 / 2 / 4 / 8 B200,
through the reference-facing feed: hv_frameset_push (FrameSet batcher, rust/heimdall-gige/src/frame.rs:127-185) ->
hv_submit_frames -> hv_wait, host frames in, per-frame results out, with the NCCL all-reduce of the line statistics every
25 frames per stream (the reference's per-camera fan-out is rust/heimdall-gige/src/lib.rs:584-616).

  python tools/bench_streams.py [--triggers T] [--sets-per-batch S] [--out digests.json] [--expect digests.json]
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_streams.py ...

Stream s is owned by rank s mod N (hv_dist.streams_of); each rank runs one detector context, one batcher over its own
cameras, and pins its host thread to its own share of the cores.  Every frame's result is reduced to a digest keyed by
(stream, trigger): --out writes them, --expect compares with the digests of another run (1 GPU vs N GPUs must be identical);
one frame set per rank is also compared with the oracle.  Prints one JSON line (rank 0).
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "heimdall-vision_b200"))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=8)
    ap.add_argument("--triggers", type=int, default=200, help="frames per stream")
    ap.add_argument("--sets-per-batch", type=int, default=4)
    ap.add_argument("--height", type=int, default=2048)
    ap.add_argument("--width", type=int, default=2448)
    ap.add_argument("--slots", type=int, default=3)
    ap.add_argument("--stats-every", type=int, default=25, help="all-reduce the line statistics every this many frames per stream")
    ap.add_argument("--copy", action="store_true", help="pageable frames, copied into the batcher's pinned slabs (default: pinned frames, referenced)")
    ap.add_argument("--out", default=None)
    ap.add_argument("--expect", default=None)
    args = ap.parse_args()
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    import torch
    import torch.distributed as dist

    import heimdall_core as hc
    import heimdall_core.camera as cam
    import hv_dist
    import synth

    # host thread placement: each rank gets its own contiguous share of the cores this process may use
    cores = sorted(os.sched_getaffinity(0))
    share = max(1, len(cores) // world)
    mine = cores[rank * share:(rank + 1) * share] or cores
    os.sched_setaffinity(0, mine)
    numa = "n/a"
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local_rank)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        numa = open(f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node").read().strip()
    except Exception:
        pass

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    h, w, S = args.height, args.width, args.sets_per_batch
    streams = hv_dist.streams_of(rank, world, args.streams)
    ncam = len(streams)
    det = hc.Detector(local_rank, num_slots=args.slots, max_defects_per_frame=512)
    # a small pool of distinct frames per stream (trigger t of stream s shows pool[s][t % P])
    P = 4
    # ... in page-locked memory, as a camera driver's ring of DMA buffers would be: the batcher references such frames
    # instead of copying them (HV_FRAMESET_ZERO_COPY; --copy: pageable frames, copied into the batcher's slabs on arrival)
    import ctypes
    pool = {}
    for s in streams:
        pool[s] = []
        for k in range(P):
            fr = synth.bottle_frame(h, w, 900 + 16 * s + k, contaminants=(s + k) % 4)
            if not args.copy:
                ptr = det.host_alloc(h * w)
                ctypes.memmove(ptr, fr.ctypes.data, h * w)
                fr = np.ctypeslib.as_array((ctypes.c_uint8 * (h * w)).from_address(ptr)).reshape(h, w)
            pool[s].append(fr)
    batcher = cam.FrameSetBatcher(det, n_cameras=ncam, sets_per_batch=S, sync_mode=cam.SyncMode.Hardware,
                                  zero_copy=not args.copy) if ncam else None

    stats_view = torch.as_tensor(hv_dist.CudaArrayView(det.stats_device_ptr()), device=dev)
    stats_buf = torch.zeros(hv_dist.STATS_WORDS, dtype=torch.int64, device=dev)
    side = torch.cuda.Stream(device=dev)

    def reduce_stats():
        if world > 1:
            with torch.cuda.stream(side):
                stats_buf.copy_(stats_view)
                dist.all_reduce(stats_buf)

    digests = {}
    inflight = []

    def pack(y, x, size, conf):
        a = np.empty(len(y), dtype=[("y", "<i4"), ("x", "<i4"), ("size", "<f8"), ("confidence", "<f8")])
        a["y"], a["x"], a["size"], a["confidence"] = y, x, size, conf
        return a.tobytes()

    def collect(ticket):
        ids = batcher.batch_ids(ticket)
        res = batcher.wait(ticket)
        for k, fid in enumerate(ids):
            for c, s in enumerate(streams):
                f = k * ncam + c
                d = res.defects_of(f)
                hsh = hashlib.sha256(pack(d["y"], d["x"], d["size"], d["confidence"])).hexdigest()[:16]
                digests[f"{s}:{fid}"] = [int(res.frames["n_components"][f]), int(res.frames["n_defects"][f]),
                                         int(res.frames["rejected"][f]), hsh]

    def run(triggers, first_trigger=0):
        for t in range(first_trigger, first_trigger + triggers):
            for c, s in enumerate(streams):
                fr = cam.CameraFrame(pool[s][t % P], w, h, cam.PixelFormat.Mono8, frame_id=t, camera=c)
                while True:
                    try:
                        tk = batcher.push(fr)
                        break
                    except hc.HeimdallCudaError as e:   # every slot in flight: collect the oldest batch, then retry the hand-over
                        if e.status != hc._abi.HV_ERR_CAPACITY:
                            raise
                        collect(inflight.pop(0))
                        tk = batcher.flush()
                        break
                if tk:
                    inflight.append(tk)
                    if len(inflight) >= args.slots:
                        collect(inflight.pop(0))
            if args.stats_every > 0 and (t + 1) % args.stats_every == 0:
                reduce_stats()
        while inflight:
            collect(inflight.pop(0))

    warm = 2 * S * args.slots
    if ncam:
        run(warm)
    torch.cuda.synchronize()
    det.stats_reset()
    digests.clear()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    if ncam:
        run(args.triggers, first_trigger=warm)
    torch.cuda.synchronize()
    secs = time.perf_counter() - t0
    frames_rank = ncam * (args.triggers // S) * S if ncam else 0  # (a trailing incomplete batch stays queued)
    # one frame set of this rank against the oracle
    oracle_ok = True
    if ncam:
        from oracle import oracle as O
        O.build()
        s0 = streams[0]
        t_chk = warm
        ref = O.detect_contamination(pool[s0][t_chk % P][:, :, None], want_intermediates=False)
        hsh = hashlib.sha256(pack([d["position"][0] for d in ref.defects], [d["position"][1] for d in ref.defects],
                                  [d["size"] for d in ref.defects], [d["confidence"] for d in ref.defects])).hexdigest()[:16]
        got = digests.get(f"{s0}:{t_chk}")
        oracle_ok = got is not None and got[1] == len(ref.defects) and got[0] == ref.ncomp and got[3] == hsh
    local = {"rank": rank, "streams": streams, "frames": frames_rank, "seconds": secs, "cores": mine[:2] + ["..."] + mine[-1:],
             "h2d_gbs": frames_rank * h * w / secs / 1e9 if secs > 0 else 0.0, "oracle_ok": bool(oracle_ok),
             "batcher": batcher.stats() if batcher else {}, "frames_inspected": det.stats()["frames_inspected"]}
    allr, alld = [local], [digests]
    if world > 1:
        torch.cuda.synchronize()
        with torch.cuda.stream(side):
            side.wait_stream(torch.cuda.current_stream())
            stats_buf.copy_(stats_view)
            dist.all_reduce(stats_buf)
        torch.cuda.synchronize()
        allr = [None] * world
        alld = [None] * world
        dist.all_gather_object(allr, local)
        dist.all_gather_object(alld, digests)
        total_stats = hv_dist.stats_dict(stats_buf.cpu().numpy())
    else:
        total_stats = {k: v for k, v in det.stats().items() if k != "area_hist"}
    if rank == 0:
        merged = {}
        for d in alld:
            merged.update(d)
        identical = None
        if args.expect and os.path.exists(args.expect):
            exp = json.load(open(args.expect))
            identical = exp == merged
        if args.out:
            json.dump(merged, open(args.out, "w"))
        tmax = max(r["seconds"] for r in allr)
        total_frames = sum(r["frames"] for r in allr)
        line = {"config": f"configs[4]: {args.streams} camera streams x {w}x{h} Mono8, {args.triggers} triggers per stream, frame sets "
                          f"of each rank's cameras, {S} sets per batch, {args.slots} batches in flight, host frames -> hv_frameset_push "
                          f"-> hv_submit_frames -> hv_wait",
                "n_gpus": world, "frames": total_frames, "seconds_max_over_ranks": tmax, "value": total_frames / tmax,
                "unit": "frames/s (end to end, host frames in, results out)", "h2d_gbs_total": total_frames * h * w / tmax / 1e9,
                "per_rank": allr, "stats_allreduced": total_stats,
                "stats_consistent": total_stats["frames_inspected"] == sum(r["frames_inspected"] for r in allr) == total_frames,
                "results_identical_to_expected": identical, "oracle_ok_all_ranks": all(r["oracle_ok"] for r in allr),
                "frame_digests": len(merged), "numa_note": f"GPU NUMA node {numa}; {len(cores)} cores visible, {share} per rank"}
        print(json.dumps(line))
    det.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
