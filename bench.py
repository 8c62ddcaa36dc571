#!/usr/bin/env python3
"""bench.py -- frames/s and achieved HBM roofline of the contamination-inspection hot path on B200.

Workload (BASELINE.json configs[1]): one step = one pass of the hot path (blur -> adaptive threshold -> CCL -> blob
statistics -> scoring -> reject decision) over a batch of 25 synthetic 1280x1024 u8 bottle frames -- one second of
line at 90 000 bottles/h -- with the batch's results delivered to the host.  Inputs rotate over a pool of distinct
batches whose total size exceeds the 126 MB L2.  Timing: --repeats windows of --steps steps, median window.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
  python bench.py --impl reference ...                            the reference's CPU path (oracle port, all host threads)

N > 1: launched by torchrun, one rank per GPU; frames are independent so every rank processes its own batches
(weak scaling, no collective on the data path); the 256-byte line-statistics vector is all-reduced over NCCL once per
step on a side stream.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import ctypes
import gc
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "heimdall-vision_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

_REAL_STDOUT = None


def emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


METRIC = "frames/s (contamination pipeline: blur->threshold->CCL->blob stats->reject), device-resident inputs"
UNIT = "frames/s"
ALG_BYTES_PER_PX = 6  # input 1 B + final mask 1 B + i32 label map 4 B per pixel


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ---------------------------------------------------------------------------------------------------------------------
# clocks: sample SM clock and throttle reasons while the GPU is under load
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
    NOTE = {"sw_power_cap": 0x4}

    def __init__(self, index: int, period_s: float = 0.01):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], 0, None
        self._stop = threading.Event()
        self._thr = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.reasons |= r
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.ok:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self) -> dict:
        if self._thr:
            self._stop.set()
            self._thr.join(timeout=2)
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        names = [n for n, bit in {**self.BAD, **self.NOTE}.items() if self.reasons & bit]
        return {"sm_mhz": int(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": names,
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference's Rust path, timed on this box's host cores
# ---------------------------------------------------------------------------------------------------------------------
def cpu_frames_per_s(frames: np.ndarray, threads: int, budget_s: float):
    """frames: (n,h,w) u8.  Runs the oracle on as many frames as fit in about budget_s; returns (fps, n_done, secs)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle as O
    O.lib()
    n = frames.shape[0]

    def one(i):
        O.detect_contamination(frames[i % n][:, :, None], want_intermediates=False)
        return 1

    t0 = time.perf_counter()
    one(0)
    per = max(time.perf_counter() - t0, 1e-3)
    total = int(max(threads, min(budget_s / per * threads, 20 * n)))
    total = max(threads, (total // threads) * threads)
    t0 = time.perf_counter()
    if threads == 1:
        for i in range(total):
            one(i)
    else:
        with ThreadPoolExecutor(threads) as ex:  # ctypes releases the GIL inside the C call
            list(ex.map(one, range(total)))
    dt = time.perf_counter() - t0
    return total / dt, total, dt


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ---------------------------------------------------------------------------------------------------------------------
def run_reference(args, rank: int, world: int) -> None:
    """The reference arm: the reference's own CPU algorithm (its Rust sources cannot be built here: no cargo/rustc, see
    DESIGN.md) as restated by the oracle port, with all host threads over independent frames.  One step = the same
    batch of `--frames` frames the CUDA arm processes per step."""
    if rank != 0:
        return
    import synth
    h, w, nf = args.height, args.width, args.frames
    threads = host_threads()
    frames = synth.bottle_batch(nf, h, w, start_index=0)
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle as O
    O.build()
    O.lib()

    def one(i):
        O.detect_contamination(frames[i][:, :, None], want_intermediates=False)

    ex = ThreadPoolExecutor(threads)  # one pool for the whole run; ctypes releases the GIL inside the C call

    def step():
        list(ex.map(one, range(nf)))

    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    ex.shutdown()
    fps = nf * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"all {nf} frames of the batch per step x {args.steps} steps over {threads} threads; "
                                   f"oracle port of detection.rs (the Rust reference cannot be built here); CPU: {cpu_model()}"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, world: int) -> dict:
    """The part of `config` both arms share (the driver compares it)."""
    return {"workload": f"batch of {args.frames} synthetic {args.width}x{args.height} u8 bottle frames per GPU per step "
                        f"(BASELINE configs[1]: one second of line at 90k BPH)",
            "frames_per_step_per_gpu": args.frames, "height": args.height, "width": args.width, "channels": 1,
            "params": "reference defaults min_size=10 max_size=3000 threshold=25"}


# the sources of the two kernels whose DRAM traffic profiles/traffic.json records (K1 and the per-frame CCL kernel)
TRAFFIC_SOURCES = ("expand_tile.cuh", "hv_common.cuh", "k_ccl_frame.cu", "k_preprocess.cu", "score_device.cuh")


def peak_gbs() -> float:
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def other_configs(hc, synth, torch, stream, peak, skip_parity, defer_tail=True):
    """BASELINE.json configs[2] and [3] at their frame sizes, device-resident, with the streaming loop of the headline
    (enqueue + fetch of the batch depth-1 back).  Batches are smaller than the configs' 256 / full-line counts to keep the
    default run short; frac = 6 B/px x pixels / time / measured HBM peak."""
    from oracle import oracle as O
    out = {}
    G = hc._abi.HV_BLUR_GAUSSIAN
    cases = [("c2_5mp_box", (64, 2048, 2448), hc.make_params(), {}, "bottle"),
             ("c2_5mp_gauss_k5", (64, 2048, 2448), hc.make_params(blur_mode=G, blur_ksize=5, gauss_sigma=0.0), dict(gauss_ksize=5, gauss_sigma=0.0), "bottle"),
             ("c2_5mp_gauss_k15_s3", (64, 2048, 2448), hc.make_params(blur_mode=G, blur_ksize=15, gauss_sigma=3.0), dict(gauss_ksize=15, gauss_sigma=3.0), "bottle"),
             ("c2_5mp_open_close_3", (64, 2048, 2448), hc.make_params(morph_open_k=3, morph_close_k=3), dict(morph_open_k=3, morph_close_k=3), "bottle"),
             ("c2_5mp_open_close_15", (64, 2048, 2448), hc.make_params(morph_open_k=15, morph_close_k=15), dict(morph_open_k=15, morph_close_k=15), "bottle"),
             ("c3_12mp_10k_blobs", (16, 3000, 4096), hc.make_params(), {}, "dense")]
    cache = {}
    for name, (n, h, w), prm, okw, kind in cases:
        key = (kind, n, h, w)
        if key not in cache:
            cache.clear()
            gen = (lambda i: synth.bottle_frame(h, w, 500 + i, contaminants=i % 4)) if kind == "bottle" else \
                (lambda i: synth.high_contamination_frame(h, w, i))
            distinct = [gen(i) for i in range(4 if kind == "bottle" else 2)]
            host = np.stack([distinct[i % len(distinct)] for i in range(n)])
            cache[key] = (host, torch.from_numpy(host).cuda())
        host, d_in = cache[key]
        det = hc.Detector(torch.cuda.current_device(), max_defects_per_frame=512 if kind == "bottle" else 32768,
                          defer_tail=defer_tail)   # (as the headline loop: HV_FLAG_DEFER_TAIL unless --no-defer-tail)
        det.set_stream(stream.cuda_stream)
        depth = det.pipeline_depth()
        n_out = depth                 # as many sets of output planes as the library keeps batches in flight
        outs = [(det.device_alloc((n, h, w), np.uint8), det.device_alloc((n, h, w), np.int32)) for _ in range(n_out)]
        res_buf, dfx_buf, _ = det._out_arrays(n, None)   # fetched into the same arrays every step (the dense case returns
                                                         # 180 k defects per batch: allocating 25 MB per fetch is host-bound)
        res = det.detect_device(d_in.data_ptr(), n, h, w, 1, prm, outs[0][0].ptr, outs[0][1].ptr)
        ok = None
        if not skip_parity:
            ref = O.detect_contamination(host[0][:, :, None], **okw)
            ok = bool(np.array_equal(outs[0][0].get(0, 1)[0], ref.mask) and np.array_equal(outs[0][1].get(0, 1)[0], ref.labels) and
                      [((int(d["y"]), int(d["x"])), float(d["size"]), float(d["confidence"])) for d in res.defects_of(0)] ==
                      [(d["position"], d["size"], d["confidence"]) for d in ref.defects])
            if not ok:
                raise SystemExit(f"{name}: results differ from the oracle; refusing to report a number")
        tickets = []

        def step(i):
            tickets.append(det.enqueue_device(d_in.data_ptr(), n, h, w, 1, prm, outs[i % n_out][0].ptr, outs[i % n_out][1].ptr))
            if len(tickets) >= n_out:
                det.fetch_into(tickets.pop(0), res_buf, dfx_buf)
        for i in range(2 * depth):
            step(i)
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        evs[0].record(stream)
        for r in range(3):
            for i in range(5):
                step(i)
            evs[r + 1].record(stream)
        while tickets:
            det.fetch_into(tickets.pop(0), res_buf, dfx_buf)
        torch.cuda.synchronize()
        ms = float(np.median([evs[r].elapsed_time(evs[r + 1]) for r in range(3)])) / 5
        gbs = ALG_BYTES_PER_PX * n * h * w / (ms * 1e-3) / 1e9
        out[name] = {"batch": f"{n} x {w}x{h}", "ms_per_step": ms, "frames_per_s": n / (ms * 1e-3), "achieved_gbs": gbs,
                     "frac": gbs / peak, "parity_checked": ok}
        if kind == "dense":
            out[name]["note"] = ("host-bound: every batch's 180 k defect records (8.6 MB) are delivered to the host and unpacked by "
                                 "one thread inside the timed region; device-side the step is 1.15 ms (0.16; 0.79 ms = 0.23 with "
                                 "HV_FLAG_DEFER_TAIL), profiles/r05_configs.md")
        for a, b in outs:
            a.free(), b.free()
        det.close()
    cache.clear()
    torch.cuda.empty_cache()
    return out


def gpu_numa_affinity(torch, dev_index: int):
    """Restrict this thread to the CPUs local to the GPU (sysfs local_cpulist of its PCI function); returns what
    numa_restore needs, or None when the topology is not visible."""
    try:
        pr = torch.cuda.get_device_properties(dev_index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        cpus = set()
        for part in open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        before = os.sched_getaffinity(0)
        local = cpus & before
        if not local or local == before:
            return None
        os.sched_setaffinity(0, local)
        return {"before": before, "cpus": len(local), "bdf": bdf}
    except (OSError, ValueError, AttributeError):
        return None


def numa_restore(state) -> None:
    if state:
        os.sched_setaffinity(0, state["before"])


def csrc_sha16() -> str:
    """Fingerprint of the kernel sources: the ncu-measured DRAM traffic in profiles/ is only quoted for the code it was
    captured from."""
    import hashlib
    hsh = hashlib.sha256()
    d = os.path.join(PKG, "csrc")
    for f in TRAFFIC_SOURCES:
        hsh.update(open(os.path.join(d, f), "rb").read())
    return hsh.hexdigest()[:16]


# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args, rank: int, local_rank: int, world: int) -> None:
    import torch
    import torch.distributed as dist

    import heimdall_core as hc
    import hv_dist
    import synth
    from heimdall_core.batch import DEFECT_DTYPE, RESULT_DTYPE

    h, w, nf = args.height, args.width, args.frames
    K, W, R = args.steps, args.warmup, args.repeats
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- inputs: a pool of distinct batches larger than L2 (126 MB), different on every rank -------------------------
    batch_bytes = nf * h * w
    pool_n = max(2, int(np.ceil(args.pool_mb * 1e6 / batch_bytes)))
    base = synth.bottle_batch(nf, h, w, start_index=rank * 100000)
    pool_host = []
    rng = np.random.default_rng(4242 + rank)
    for p in range(pool_n):
        # new noise realisation + frame permutation per pool entry: distinct bytes, same statistics, cheap to generate
        perm = rng.permutation(nf)
        b = base[perm].astype(np.int16) + rng.integers(-1, 2, size=base.shape, dtype=np.int16)
        pool_host.append(np.clip(b, 0, 255).astype(np.uint8))
    pool_dev = [torch.from_numpy(b).to(dev) for b in pool_host]
    det = hc.Detector(local_rank, num_slots=args.slots, defer_tail=not args.no_defer_tail)
    # The output planes come from the library's allocator (hv_device_alloc): memory with L2 compute-data compression, so
    # the almost entirely zero mask / label planes cost less DRAM write time.  --no-compress: plain cudaMalloc memory.
    depth = det.pipeline_depth()  # output sets in rotation = batches the library keeps in flight on the device
    d_mask = [det.device_alloc((nf, h, w), np.uint8, not args.no_compress) for _ in range(depth)]
    d_labels = [det.device_alloc((nf, h, w), np.int32, not args.no_compress) for _ in range(depth)]
    out_mem = ("L2-compressible (cuMemCreate, CU_MEM_ALLOCATION_COMP_GENERIC)" if d_labels[0].compressed
               else "plain device memory")
    stream = torch.cuda.current_stream()
    det.set_stream(stream.cuda_stream)
    params = hc.make_params()
    params_morph = hc.make_params(morph_open_k=3, morph_close_k=3)   # contamination_detector.py:81-87: 3x3 open, 3x3 close

    def check_against_oracle(res, batch_host, mask_t, labels_t, frames, **okw):
        from oracle import oracle as O
        O.build()
        ok = True
        for f in frames:
            ref = O.detect_contamination(batch_host[f][:, :, None], **okw)
            ok = ok and (np.array_equal(mask_t.get(f, 1)[0], ref.mask) and
                         np.array_equal(labels_t.get(f, 1)[0], ref.labels) and
                         [((int(d["y"]), int(d["x"])), float(d["size"]), float(d["confidence"]))
                          for d in res.defects_of(f)] ==
                         [(d["position"], d["size"], d["confidence"]) for d in ref.defects] and
                         bool(res.rejected[f]) == ref.reject)
        return ok

    # ---- parity gate, on EVERY rank: the timed configuration must agree with the oracle before any number counts ------
    res0 = det.detect_device(pool_dev[0].data_ptr(), nf, h, w, 1, params, d_mask[0].data_ptr(), d_labels[0].data_ptr())
    parity = None
    if not args.skip_parity:
        parity = check_against_oracle(res0, pool_host[0], d_mask[0], d_labels[0], (0, nf - 1))
        if not parity:
            raise SystemExit(f"rank {rank}: parity check against the oracle FAILED; refusing to report a number")

    # ---- line statistics all-reduce (the only collective): 32 x u64, side stream, every --stats-every steps ----------------
    stats_view = torch.as_tensor(hv_dist.CudaArrayView(det.stats_device_ptr()), device=dev)
    stats_buf = torch.zeros(hv_dist.STATS_WORDS, dtype=torch.int64, device=dev)
    side = torch.cuda.Stream(device=dev)

    # The snapshot + all-reduce pair costs about 50 us of host time (a copy and an NCCL call), which a 40 us step cannot hide
    # when the thread that issues it is the one that feeds the pipeline: a helper thread issues it instead (torch releases
    # the GIL inside both calls), on a side stream.  The launching thread only posts a request.
    import queue
    stats_q = queue.Queue()

    def stats_worker():
        torch.cuda.set_device(local_rank)
        while True:
            item = stats_q.get()
            if item is None:
                return
            with torch.cuda.stream(side):
                stats_buf.copy_(stats_view)
                dist.all_reduce(stats_buf)
            stats_q.task_done()

    stats_thread = None
    if world > 1:
        with torch.cuda.stream(side):
            stats_buf.copy_(stats_view)
            dist.all_reduce(stats_buf)          # communicator set-up outside any timed region
        torch.cuda.synchronize()
        stats_thread = threading.Thread(target=stats_worker, daemon=True)
        stats_thread.start()

    def reduce_stats(final=False):
        """Snapshot of the running line counters -> all-reduce, entirely on the side stream.  The counters are 64-bit
        atomics that only grow, so a report may be a few microseconds stale but every counter in it is consistent; the
        launching stream is never touched (an event between two kernels there would serialise K1 behind the per-frame
        CCL kernel and undo the programmatic-dependent-launch overlap).  final=True waits for the requests posted so far
        and orders the snapshot after everything enqueued on the launching stream: the totals reported at the end."""
        if world == 1:
            return
        if final:
            stats_q.join()
            side.wait_stream(stream)
            with torch.cuda.stream(side):
                stats_buf.copy_(stats_view)
                dist.all_reduce(stats_buf)
            return
        stats_q.put(1)

    # Streaming loop: one step = enqueue one batch + collect the results of the batch depth-1 steps back, which the copy
    # engine has meanwhile delivered to pinned host memory (hv_fetch_ticket).  EVERY batch's per-frame records and defect
    # list reach the host inside the timed region; `host` accumulates what arrived so that it can be checked against the
    # device-side line statistics afterwards.
    class Stream_:
        def __init__(self, prm):
            self.prm = prm
            self.tickets = [0] * depth
            self.res = [np.zeros(nf, RESULT_DTYPE) for _ in range(depth)]
            self.dfx = [np.zeros(nf * det.defect_cap, DEFECT_DTYPE) for _ in range(depth)]
            self.i = 0
            self.cstep = 0
            self.batches = 0
            # per-frame records of everything collected so far, summed field by field (one vectorised add per batch: the
            # launching thread has 40 us per step for two library calls and this)
            self.views = [r.view(np.uint32).reshape(nf, 6) for r in self.res]
            self.acc = np.zeros((nf, 6), np.uint64)
            self.last = None

        @property
        def frames(self):
            return self.batches * nf

        @property
        def rejected(self):
            return int(self.acc[:, 3].sum())

        @property
        def defects(self):
            return int(self.acc[:, 1].sum())

        def collect(self, slot):
            self.last = det.fetch_into(self.tickets[slot], self.res[slot], self.dfx[slot])
            np.add(self.acc, self.views[slot], out=self.acc)
            self.batches += 1
            self.tickets[slot] = 0

        def step(self, collective=True):
            i = self.i
            slot = i % depth
            self.tickets[slot] = det.enqueue_device(pool_dev[i % pool_n].data_ptr(), nf, h, w, 1, self.prm,
                                                    d_mask[slot].data_ptr(), d_labels[slot].data_ptr())
            nxt = (i + 1) % depth
            if self.tickets[nxt]:
                self.collect(nxt)
            if collective and args.stats_every > 0:
                self.cstep += 1     # counted from the start of the timed region: every rank issues the same number
                if self.cstep % args.stats_every == 0:
                    reduce_stats()
            self.i = i + 1

        def drain(self):
            for k in range(1, depth + 1):
                slot = (self.i + k - 1) % depth
                if self.tickets[slot]:
                    self.collect(slot)

    def timed_windows(S, repeats, collective=True):
        """`repeats` windows of K steps back to back, a CUDA event on the launching stream at every window boundary.
        Returns per-window milliseconds, launches and wall seconds."""
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(repeats + 1)]
        S.cstep = 0
        l0 = det.launch_count()
        gc.collect()
        gc.disable()  # a collection pause of the launching thread would show up as a pipeline bubble
        t0 = time.perf_counter()
        evs[0].record(stream)
        for r in range(repeats):
            for _ in range(K):
                S.step(collective)
            evs[r + 1].record(stream)
        S.drain()
        if world > 1:
            stats_q.join()   # every rank has issued the same number of all-reduces before anybody moves on
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        gc.enable()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        return [evs[r].elapsed_time(evs[r + 1]) for r in range(repeats)], det.launch_count() - l0, wall

    # ---- settle clocks under load, then W warm-up steps ------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    S = Stream_(params)
    t_end = time.perf_counter() + args.settle_s
    while time.perf_counter() < t_end:
        for _ in range(20):
            S.step(collective=False)  # time-based loop: ranks do different numbers of iterations, so no collective here
    S.drain()
    torch.cuda.synchronize()
    for _ in range(W):
        S.step(collective=False)
    S.drain()

    # ---- timed region -----------------------------------------------------------------------------------------------------
    stats_before = det.stats()
    host_before = (S.frames, S.rejected, S.defects)
    win_ms, launches, wall = timed_windows(S, R)
    stats_after = det.stats()
    host_delta = (S.frames - host_before[0], S.rejected - host_before[1], S.defects - host_before[2])
    dev_delta = (stats_after["frames_inspected"] - stats_before["frames_inspected"],
                 stats_after["frames_rejected"] - stats_before["frames_rejected"],
                 stats_after["total_defects"] - stats_before["total_defects"])
    if host_delta != dev_delta or host_delta[0] != R * K * nf:
        raise SystemExit(f"rank {rank}: results delivered to the host {host_delta} != device line statistics {dev_delta} "
                         f"(expected {R * K * nf} frames)")
    parity_after = None
    if not args.skip_parity:  # the overlapped steady state must still be bit-exact (every rank checks its own last step)
        j = S.i - 1
        parity_after = check_against_oracle(S.last, pool_host[j % pool_n], d_mask[j % depth], d_labels[j % depth], (1, nf - 2))
        if not parity_after:
            raise SystemExit(f"rank {rank}: parity check of the last timed step FAILED; refusing to report a number")
    clocks = sampler.stop()

    med, lo, hi = float(np.median(win_ms)), float(min(win_ms)), float(max(win_ms))
    per_rank = [[med, lo, hi, wall / (R * K) * 1e3]]
    if world > 1:
        t = torch.tensor(per_rank[0], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
        per_rank = [[float(x) for x in a.tolist()] for a in allr]
    ms_window = max(p[0] for p in per_rank)          # max over ranks of the per-rank median window
    value = world * nf * K / (ms_window * 1e-3)

    # ---- the same without the collective (N > 1): is the all-reduce visible at all? ----------------------------------------
    nocoll = None
    if world > 1 and args.stats_every > 0:
        w2, _, _ = timed_windows(S, max(5, R // 3), collective=False)
        t = torch.tensor([float(np.median(w2))], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nocoll = float(t.item()) / K

    # ---- sustained: one event pair around R*K steps, no event in between (an event record between two batches makes the
    #      next K1 wait for everything before it: one pipeline drain + fill per window) ----------------------------------------
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    S.cstep = 0
    e0.record(stream)
    for _ in range(R * K):
        S.step()
    e1.record(stream)
    S.drain()
    if world > 1:
        stats_q.join()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / (R * K)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sustained_ms = float(t.item())

    # ---- per-kernel durations: K steps with a CUDA event pair around every kernel, on the launching stream (the events
    #      serialise the kernels, so these are the isolated per-kernel times) ------------------------------------------------
    det.profile_enable(None)
    for _ in range(K):
        S.step(collective=False)
    S.drain()
    torch.cuda.synchronize()
    prof = det.profile()
    shares = {k: (v["ms"] / v["launches"] if v["launches"] else 0.0) for k, v in prof.items()}
    det.profile_enable([])

    # ---- the pipeline WITH morphology (blur -> threshold -> open 3x3 -> close 3x3 -> CCL): own parity gate, own timing ------
    morph = None
    if not args.no_morph:
        resm = det.detect_device(pool_dev[1 % pool_n].data_ptr(), nf, h, w, 1, params_morph, d_mask[0].data_ptr(),
                                 d_labels[0].data_ptr())
        okm = True
        if not args.skip_parity:
            okm = check_against_oracle(resm, pool_host[1 % pool_n], d_mask[0], d_labels[0], (0, nf - 1),
                                       morph_open_k=3, morph_close_k=3)
            if not okm:
                raise SystemExit(f"rank {rank}: morphology pipeline differs from the oracle; refusing to report a number")
        SM = Stream_(params_morph)
        for _ in range(max(W, 2 * depth)):
            SM.step(collective=False)
        SM.drain()
        l0 = det.launch_count()
        wm, lm, _ = timed_windows(SM, max(5, R // 3), collective=False)
        t = torch.tensor([float(np.median(wm))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        morph = {"ms_per_step": float(t.item()) / K, "windows": len(wm), "launches_per_step": lm / (len(wm) * K),
                 "parity_checked": bool(okm) and not args.skip_parity}

    # ---- end to end: pinned host frames -> H2D -> pipeline -> D2H of the results, through hv_submit / hv_wait -----------------
    n_pin = min(pool_n, max(args.slots, 3))
    pins = []
    # the staging buffers are allocated and first touched by a thread that runs on the GPU's own NUMA node (first-touch
    # placement): a ring on the other socket costs a third of the H2D rate on two-socket hosts
    numa = gpu_numa_affinity(torch, local_rank)
    for p in range(n_pin):
        ptr = det.host_alloc(batch_bytes)
        ctypes.memmove(ptr, pool_host[p].ctypes.data, batch_bytes)
        pins.append(ptr)
    numa_restore(numa)
    det.set_stream(None)
    inflight = []

    def e2e_steps(count):
        done = 0
        for j in range(count):
            if len(inflight) == args.slots:
                det.wait(inflight.pop(0), nf)
                done += 1
            inflight.append(det.submit(pins[j % n_pin], nf, h, w, 1, params))
        while inflight:
            det.wait(inflight.pop(0), nf)
            done += 1
        return done

    e2e_steps(max(W, 3))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_steps(K * max(1, R // 3))
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / max(1, R // 3)
    e2e_rank = [e2e_s]
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
        e2e_rank = [float(a.item()) for a in allr]
    e2e_value = world * nf * K / max(e2e_rank)
    for ptr in pins:
        det.host_free(ptr)
    d2h_bytes = nf * (24 + 4 + det.defect_cap * 48)

    # ---- the other BASELINE configs at their stated frame sizes (rank 0 at N = 1): parity of one frame against the oracle,
    #      then a short timed run (3 windows of 5 steps, median) -- the full tables are tools/bench_configs.py's ----------------
    other = None
    if world == 1 and not args.no_other_configs:
        other = other_configs(hc, synth, torch, stream, peak_gbs(), args.skip_parity, not args.no_defer_tail)

    torch.cuda.synchronize()
    if world > 1:
        reduce_stats(final=True)
        torch.cuda.synchronize()
        total_stats = hv_dist.stats_dict(stats_buf.cpu().numpy())
        fr = torch.tensor([det.stats()["frames_inspected"]], dtype=torch.int64, device=dev)
        dist.all_reduce(fr)
        if total_stats["frames_inspected"] != int(fr.item()):
            raise SystemExit("all-reduced line statistics do not add up to the ranks' own counters")
    else:
        total_stats = {k: v for k, v in det.stats().items() if k != "area_hist"}

    if world > 1:
        stats_q.put(None)
        stats_thread.join(timeout=10)
        torch.cuda.synchronize()
        dist.barrier()
    if rank != 0:
        det.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (K1 preprocess_mask) ---------------------------------------------------------------
    k1 = prof["preprocess_mask"]
    k1_ms = k1["ms"] / max(k1["launches"], 1)
    alg_bytes = ALG_BYTES_PER_PX * h * w * nf
    peak = peak_gbs()
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
        else "fallback (B200_PROFILING.md)"
    # DRAM bytes per launch from the ncu --set full capture of THIS code (profiles/traffic.json carries the fingerprint of
    # the kernel sources it was taken from; a capture of other code is not quoted)
    traffic, traffic_note = None, "no ncu capture of the current kernel sources under profiles/"
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if tj.get("csrc_sha16") == csrc_sha16():
            traffic = tj["k1_dram_bytes_per_launch"]
            traffic_note = (f"{tj['source']}: K1 {tj['k1_dram_bytes_per_launch'] / 1e6:.1f} MB + per-frame CCL kernel "
                            f"{tj['ccl_dram_bytes_per_launch'] / 1e6:.1f} MB of DRAM traffic per step")
        else:
            traffic_note = f"profiles/traffic.json was captured from other kernel sources ({tj.get('csrc_sha16')}): not quoted"
    except Exception:
        pass

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        fps_mt, n_mt, s_mt = cpu_frames_per_s(pool_host[0], threads, args.cpu_seconds / 2)
        fps_1t, n_1t, s_1t = cpu_frames_per_s(pool_host[0], 1, args.cpu_seconds / 2)
        cpu = {"value": fps_mt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{n_mt} frames of the same workload in {s_mt:.1f} s on {threads} threads (one frame per "
                         f"thread); single thread (faithful to the reference, which never uses rayon): "
                         f"{fps_1t:.2f} frames/s over {n_1t} frames; CPU: {cpu_model()}",
               "single_thread_value": fps_1t}

    step_ms = ms_window / K
    achieved = alg_bytes / (step_ms * 1e-3) / 1e9
    cfg = workload_config(args, world)
    cfg.update({
        "l2_policy": f"inputs rotate over a pool of {pool_n} distinct batches ({pool_n * batch_bytes / 1e6:.0f} MB "
                     f"> 126 MB L2); each step also writes {5 * batch_bytes / 1e6:.0f} MB of mask+labels",
        "output_memory": out_mem,
        "parallelism": f"dp{world} (frames sharded, no data-path collective; 256 B all-reduce of the running line statistics "
                       f"every {args.stats_every} steps = {args.stats_every * nf} frames per GPU, issued by a helper thread on a side stream)",
        "timed": f"{R} windows of K = {K} steps back to back, a CUDA event on the launching stream at every window boundary; "
                 f"ms_per_step = median window / K (max over ranks of the per-rank medians).  One step = hv_enqueue_device "
                 f"of one batch + hv_fetch_ticket of the batch {depth - 1} steps back: the per-frame records and defect "
                 f"lists of EVERY batch reach pinned host memory inside the timed region (copy stream ordered by the "
                 f"slot's device-side completion counter) and their sums are checked against the device's line statistics; "
                 f"K1 of step i+1 overlaps the per-frame CCL kernels of the previous steps (programmatic dependent launch, "
                 f"{depth} output sets in rotation)",
        "defer_tail": (None if args.no_defer_tail else
                       "HV_FLAG_DEFER_TAIL: call i enqueues the per-frame kernel of batch i - 2, then K1 of batch i, so the "
                       "window-boundary event sits behind a K1 and in front of a per-frame kernel whose input is long complete "
                       "instead of draining the K1 / per-frame overlap; every window still holds K K1 launches and K per-frame "
                       "launches (shifted by two batches), results are fetched by ticket as before.  --no-defer-tail measures "
                       "the plain order (DESIGN.md section 5)")})
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "repeats": R,
        "ms_per_step": step_ms, "ms_per_step_min": min(p[1] for p in per_rank) / K,
        "ms_per_step_max": max(p[2] for p in per_rank) / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": cfg,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": batch_bytes, "d2h_bytes_per_step": d2h_bytes,
                "how": f"hv_submit/hv_wait, {args.slots} batches in flight, pinned host frames, wall clock, "
                       f"{K * max(1, R // 3)} steps",
                "h2d_gbs_per_rank": [batch_bytes * K / s / 1e9 for s in e2e_rank],
                "staging_numa_local": bool(numa)},
        "gpu_launches": int(round(launches / R)),
        # The kernels of consecutive steps overlap (K1 launches follow each other without a gap, the per-frame CCL kernels
        # of the last few steps run beside them), so no kernel's own duration appears in the step: `kernel_ms` is the STEP
        # time, i.e. the throughput of the K1 -> CCL chain per batch, and `frac` is the pipeline's fraction of the HBM
        # roofline.  K1's duration when run alone is kernel_ms_isolated.
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_note": traffic_note, "kernel": "preprocess_mask (K1) -> ccl_frame_fused chain",
                     "kernel_ms": step_ms,
                     "kernel_ms_how": "step time = chain throughput: median timed window / K launches (CUDA events on the "
                                      "launching stream).  kernel_ms_isolated / isolated_frac: mean duration of K1 alone, "
                                      "with an event pair around every kernel in a separate pass of K steps (that "
                                      "serialises the kernels and removes the overlap)",
                     "kernel_ms_isolated": k1_ms, "isolated_achieved": alg_bytes / (k1_ms * 1e-3) / 1e9 if k1_ms else 0.0,
                     "isolated_frac": alg_bytes / (k1_ms * 1e-3) / 1e9 / peak if k1_ms else 0.0,
                     "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                     "sustained_ms_per_step": sustained_ms,
                     "sustained_frac": alg_bytes / (sustained_ms * 1e-3) / 1e9 / peak,
                     "sustained_how": f"one event pair around {R * K} steps, no event in between"},
        "kernel_ms_per_step": shares,
        "per_rank_ms_per_step": [{"median": p[0] / K, "min": p[1] / K, "max": p[2] / K, "wall": p[3]} for p in per_rank],
        "clocks": clocks,
        "parity_checked": bool(parity) and bool(parity_after),
        "parity_ranks": world,
        "results_delivered": {"frames": host_delta[0], "rejected": host_delta[1], "defects": host_delta[2],
                              "equal_to_device_line_stats": True},
        "wall_ms_per_step": wall / (R * K) * 1e3,
        "line_stats": total_stats,
        "last_step": {"rejected": int(S.last.rejected.sum()), "defects": int(S.last.frames["n_defects"].sum())},
    }
    if nocoll is not None:
        line["ms_per_step_no_collective"] = nocoll
    if other is not None:
        line["other_configs"] = other
    if morph is not None:
        m_ach = alg_bytes / (morph["ms_per_step"] * 1e-3) / 1e9
        line["roofline_morph"] = {"pipeline": "blur -> adaptive threshold -> open 3x3 -> close 3x3 -> CCL -> stats -> reject "
                                              "(heimdall/detectors/contamination_detector.py:81-87)",
                                  "bound": "hbm", "achieved": m_ach, "peak": peak, "unit": "GB/s", "frac": m_ach / peak,
                                  "value": world * nf / (morph["ms_per_step"] * 1e-3), **morph}
    if cpu is not None:
        line["cpu_baseline"] = cpu
    bad = [r for r in clocks.get("reasons", []) if r in ClockSampler.BAD]
    if bad:
        line["clock_warning"] = f"throttle reasons seen: {bad}"
    emit(line)
    det.close()
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    # Only the JSON line may reach stdout (NCCL and others print banners there): park fd 1 on stderr until the end.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--repeats", type=int, default=15, help="timed windows of --steps steps, back to back; the median counts")
    ap.add_argument("--no-defer-tail", action="store_true",
                    help="enqueue every batch's per-frame kernel with the batch itself (without HV_FLAG_DEFER_TAIL)")
    ap.add_argument("--no-morph", action="store_true", help="skip the roofline_morph sub-record")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the other_configs sub-record (BASELINE configs[2], [3])")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--frames", type=int, default=25)
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--pool-mb", type=float, default=260.0)
    ap.add_argument("--slots", type=int, default=3)
    ap.add_argument("--settle-s", type=float, default=1.5)
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-parity", action="store_true")
    ap.add_argument("--no-compress", action="store_true", help="output planes in plain cudaMalloc memory")
    ap.add_argument("--stats-every", type=int, default=20,
                    help="all-reduce the running line statistics every this many steps (N > 1)")
    args = ap.parse_args()
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
