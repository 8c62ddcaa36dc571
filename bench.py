#!/usr/bin/env python3
"""bench.py -- frames/s and achieved HBM roofline of the contamination-inspection hot path on B200.

Workload (BASELINE.json configs[1]): one step = one pass of the hot path (blur -> adaptive threshold -> CCL -> blob
statistics -> scoring -> reject decision) over a batch of 25 synthetic 1280x1024 u8 bottle frames -- one second of
line at 90 000 bottles/h.  Inputs rotate over a pool of distinct batches whose total size exceeds the 126 MB L2.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
  python bench.py --impl reference ...                            the reference's CPU path (oracle port, all host threads)

N > 1: launched by torchrun, one rank per GPU; frames are independent so every rank processes its own batches
(weak scaling, no collective on the data path); the 256-byte line-statistics vector is all-reduced over NCCL once per
step on a side stream.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import ctypes
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
    DESIGN.md) as restated by the oracle port, with all host threads over independent frames."""
    if rank != 0:
        return
    import synth
    h, w, nf = args.height, args.width, args.frames
    threads = host_threads()
    sample = max(1, min(nf, threads))          # frames per step: a bounded sample of the 25-frame batch
    frames = synth.bottle_batch(sample, h, w, start_index=0)
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle as O
    O.build()
    O.lib()

    def one(i):
        O.detect_contamination(frames[i][:, :, None], want_intermediates=False)

    ex = ThreadPoolExecutor(threads)  # one pool for the whole run: a step is `sample` frames, one per thread

    def step():
        list(ex.map(one, range(sample)))

    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    ex.shutdown()
    fps = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"batch of {nf} synthetic {w}x{h} u8 bottle frames (BASELINE configs[1])",
                   "frames_per_step_timed": sample, "height": h, "width": w},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} frames per step x {args.steps} steps, one frame per thread; oracle port "
                                   f"of detection.rs (the Rust reference cannot be built here); CPU: {cpu_model()}"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args, rank: int, local_rank: int, world: int) -> None:
    import torch
    import torch.distributed as dist

    import heimdall_core as hc
    import hv_dist
    import synth

    h, w, nf = args.height, args.width, args.frames
    K, W = args.steps, args.warmup
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
    # hv_pipeline_depth() sets of outputs in rotation: the kernels of several batches are in flight at once (K1 of step
    # i+1 starts while K1 of step i retires, the per-frame CCL kernels of the last few steps run next to them)
    det = hc.Detector(local_rank, num_slots=args.slots)
    # The output planes come from the library's allocator (hv_device_alloc): memory with L2 compute-data compression, so
    # the almost entirely zero mask / label planes cost less DRAM write time.  --no-compress: plain cudaMalloc memory.
    n_out = det.pipeline_depth()  # output sets in rotation = batches the library keeps in flight on the device
    d_mask = [det.device_alloc((nf, h, w), np.uint8, not args.no_compress) for _ in range(n_out)]
    d_labels = [det.device_alloc((nf, h, w), np.int32, not args.no_compress) for _ in range(n_out)]
    out_mem = ("L2-compressible (cuMemCreate, CU_MEM_ALLOCATION_COMP_GENERIC)" if d_labels[0].compressed
               else "plain device memory")
    stream = torch.cuda.current_stream()
    det.set_stream(stream.cuda_stream)
    params = hc.make_params()

    # ---- parity gate: the timed configuration must agree with the oracle before any number counts ----------------------
    parity = None
    def check_against_oracle(res, batch_host, mask_t, labels_t, frames):
        from oracle import oracle as O
        O.build()
        ok = True
        for f in frames:
            ref = O.detect_contamination(batch_host[f][:, :, None])
            ok = ok and (np.array_equal(mask_t.get(f, 1)[0], ref.mask) and
                         np.array_equal(labels_t.get(f, 1)[0], ref.labels) and
                         [((int(d["y"]), int(d["x"])), float(d["size"]), float(d["confidence"]))
                          for d in res.defects_of(f)] ==
                         [(d["position"], d["size"], d["confidence"]) for d in ref.defects])
        return ok

    res0 = det.detect_device(pool_dev[0].data_ptr(), nf, h, w, 1, params, d_mask[0].data_ptr(), d_labels[0].data_ptr())
    if rank == 0 and not args.skip_parity:
        parity = check_against_oracle(res0, pool_host[0], d_mask[0], d_labels[0], (0, nf - 1))
        if not parity:
            raise SystemExit("parity check against the oracle FAILED; refusing to report a number")

    # ---- line statistics all-reduce (the only collective): 32 x u64, side stream, every --stats-every steps ----------------
    stats_view = torch.as_tensor(hv_dist.CudaArrayView(det.stats_device_ptr()), device=dev)
    stats_buf = torch.zeros(hv_dist.STATS_WORDS, dtype=torch.int64, device=dev)
    side = torch.cuda.Stream(device=dev)

    def reduce_stats(final=False):
        """Snapshot of the running line counters -> all-reduce, entirely on the side stream.  The counters are 64-bit
        atomics that only grow, so a report may be a few microseconds stale but every counter in it is consistent; the
        launching stream is never touched (an event between two kernels there would serialise K1 behind the per-frame
        CCL kernel and undo the programmatic-dependent-launch overlap).  final=True orders the snapshot after everything
        enqueued so far: the totals reported at the end."""
        if world == 1:
            return
        if final:
            side.wait_stream(stream)
        with torch.cuda.stream(side):
            stats_buf.copy_(stats_view)
            dist.all_reduce(stats_buf)

    def step(i, collective=True):
        det.enqueue_device(pool_dev[i % pool_n].data_ptr(), nf, h, w, 1, params, d_mask[i % n_out].data_ptr(),
                           d_labels[i % n_out].data_ptr())
        if collective and (i + 1) % args.stats_every == 0:
            reduce_stats()

    # ---- settle clocks under load, then W warm-up steps ------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_end = time.perf_counter() + args.settle_s
    i = 0
    while time.perf_counter() < t_end:
        for _ in range(20):
            step(i, collective=False)  # time-based loop: ranks do different numbers of iterations, so no collective here
            i += 1
        torch.cuda.synchronize()
    for j in range(W):
        step(j)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    # ---- timed region: exactly K steps, CUDA events on the launching stream ------------------------------------------------
    # no per-kernel events inside the timed region: an event between two kernels would serialise them and hide the
    # overlap of K1(step i+1) with the per-frame CCL of step i that production runs get
    l0 = det.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for j in range(K):
        step(j)
    e1.record(stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches = det.launch_count() - l0
    ev_ms = e0.elapsed_time(e1)
    last = det.fetch_results(nf)           # results of the final step: proves the pipeline ran to completion
    clocks = sampler.stop()
    parity_after = None
    if rank == 0 and not args.skip_parity:  # the overlapped steady state must still be bit-exact
        j = K - 1
        parity_after = check_against_oracle(last, pool_host[j % pool_n], d_mask[j % n_out], d_labels[j % n_out], (1, nf - 2))
        if not parity_after:
            raise SystemExit("parity check of the last timed step FAILED; refusing to report a number")

    t = torch.tensor([ev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * nf * K / (ms_total * 1e-3)

    # ---- per-kernel durations: the same K steps again with a CUDA event pair around every kernel, on the launching
    #      stream (the events serialise the kernels, so these are the isolated per-kernel times) ------------------------------
    det.profile_enable(None)
    for j in range(K):
        step(j)
    torch.cuda.synchronize()
    prof = det.profile()
    shares = {k: (v["ms"] / v["launches"] if v["launches"] else 0.0) for k, v in prof.items()}
    det.profile_enable([])

    # ---- end to end: pinned host frames -> H2D -> pipeline -> D2H of the results, through hv_submit / hv_wait -----------------
    n_pin = min(pool_n, max(args.slots, 3))
    pins = []
    for p in range(n_pin):
        ptr = det.host_alloc(batch_bytes)
        ctypes.memmove(ptr, pool_host[p].ctypes.data, batch_bytes)
        pins.append(ptr)
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
    e2e_steps(K)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * nf * K / float(t.item())
    for ptr in pins:
        det.host_free(ptr)
    d2h_bytes = nf * (24 + det.defect_cap * 48)

    if world > 1:
        torch.cuda.synchronize()
        reduce_stats(final=True)
        torch.cuda.synchronize()
        total_stats = hv_dist.stats_dict(stats_buf.cpu().numpy())
    else:
        total_stats = {k: v for k, v in det.stats().items() if k != "area_hist"}

    if rank != 0:
        det.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (K1 preprocess_mask) ---------------------------------------------------------------
    k1 = prof["preprocess_mask"]
    k1_ms = k1["ms"] / max(k1["launches"], 1)
    alg_bytes = ALG_BYTES_PER_PX * h * w * nf
    achieved = alg_bytes / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else 0.0
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(mp["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))["dram_bytes_per_launch"]
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

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"batch of {nf} synthetic {w}x{h} u8 bottle frames per GPU per step "
                               f"(BASELINE configs[1]: one second of line at 90k BPH)",
                   "frames_per_step_per_gpu": nf, "height": h, "width": w, "channels": 1,
                   "params": "reference defaults min_size=10 max_size=3000 threshold=25",
                   "l2_policy": f"inputs rotate over a pool of {pool_n} distinct batches ({pool_n * batch_bytes / 1e6:.0f} MB "
                                f"> 126 MB L2); each step also writes {5 * batch_bytes / 1e6:.0f} MB of mask+labels",
                   "output_memory": out_mem,
                   "parallelism": f"dp{world} (frames sharded, no data-path collective; 256 B all-reduce of the running line statistics every {args.stats_every} steps = {args.stats_every * nf} frames per GPU, on a side stream)",
                   "timed": "K x hv_enqueue_device on one stream, CUDA events on that stream; K1 of step i+1 overlaps the "
                            "per-frame CCL kernels of the previous steps (programmatic dependent launch, device-side completion counters, "
                            f"{n_out} output sets in rotation); "
                            "results of every step stay on the device, the last step's are fetched and checked against "
                            "the oracle"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": batch_bytes, "d2h_bytes_per_step": d2h_bytes,
                "how": f"hv_submit/hv_wait, {args.slots} batches in flight, pinned host frames, wall clock"},
        "gpu_launches": int(launches),
        # K1 is the dominant kernel and the critical path: its launches follow each other without a gap (each one starts as
        # the CTAs of the previous one retire), the per-frame CCL kernels of the last few steps run next to them.  Its
        # average time per launch over the timed region is therefore the step time; the isolated time (events around
        # every launch, kernels serialised) is reported beside it.
        "roofline": {"bound": "hbm", "achieved": alg_bytes / (ms_total / K * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg_bytes / (ms_total / K * 1e-3) / 1e9 / peak,
                     "traffic": traffic, "kernel": "preprocess_mask (K1)", "kernel_ms": ms_total / K,
                     "kernel_ms_how": "timed region / K launches, CUDA events on the launching stream: K1 launches are "
                                      "back to back (programmatic dependent launch) and the other kernels run concurrently, "
                                      "so this is K1's sustained time per launch; kernel_ms_isolated / isolated_frac: "
                                      "mean over K launches with an event pair around every kernel in a separate pass of the "
                                      "same K steps (that serialises the kernels and removes the overlap)",
                     "kernel_ms_isolated": k1_ms, "isolated_achieved": achieved, "isolated_frac": achieved / peak,
                     "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                     "pipeline_achieved": alg_bytes / (ms_total / K * 1e-3) / 1e9,
                     "pipeline_frac": alg_bytes / (ms_total / K * 1e-3) / 1e9 / peak},
        "kernel_ms_per_step": shares,
        "clocks": clocks,
        "parity_checked": bool(parity) and bool(parity_after),
        "wall_ms_per_step": wall / K * 1e3,
        "line_stats": total_stats,
        "last_step": {"rejected": int(last.rejected.sum()), "defects": int(last.frames["n_defects"].sum())},
    }
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
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
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
    ap.add_argument("--stats-every", type=int, default=8,
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
