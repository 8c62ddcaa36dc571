"""CPU-only checks of the integer identities and bit tricks the kernels rely on (restated in numpy / Python)."""
import numpy as np
import pytest


def test_div25_mul_shift_is_exact_for_every_possible_sum():
    s = np.arange(0, 25 * 255 + 1, dtype=np.uint64)          # 5x5 sums of u8
    assert np.array_equal((s * 5243) >> 17, s // 25)
    assert not np.array_equal((np.arange(0, 70000, dtype=np.uint64) * 5243) >> 17, np.arange(0, 70000) // 25)


def test_division_free_threshold_test_is_equivalent():
    """px < floor(S/cnt) - c  <=>  (px + c + 1) * cnt <= S, and px > floor(S/cnt) - c  <=>  not((px + c) * cnt <= S)."""
    rng = np.random.default_rng(0)
    for cnt in [36, 42, 66, 77, 121, 1, 11]:
        px = rng.integers(0, 256, 200000)
        S = rng.integers(0, 255 * cnt + 1, 200000)
        for c in [-256, -30, -1, 0, 1, 2, 15, 25, 200, 256]:
            mean = S // cnt
            assert np.array_equal(px < mean - c, (px + c + 1) * cnt <= S)
            assert np.array_equal(px > mean - c, ~((px + c) * cnt <= S))


def _wrap_i32(v):
    return (v + 2 ** 31) % 2 ** 32 - 2 ** 31


def test_threshold_clamp_does_not_change_results():
    """`mean - c` is i32 arithmetic that wraps in the reference's release build (detection.rs:211; rust/Cargo.toml
    [profile.release] has no overflow-checks).  Clamping c to [-256, 256] is exact for u8 data as long as the
    subtraction does not wrap, i.e. for c > 255 - 2^31; at and below that the library switches to the wrapped form
    `mask = mean < c + 2^31` (hv_api.cu: threshold_plan)."""
    px = np.arange(256)[:, None]
    mean = np.arange(256)[None, :]
    for c, cc in [(10 ** 9, 256), (2 ** 31 - 1, 256), (257, 256), (-257, -256), (-2 ** 31 + 256, -256), (-10 ** 9, -256)]:
        assert np.array_equal(px < _wrap_i32(mean - c), px < mean - cc)
    for c in [-2 ** 31, -2 ** 31 + 1, -2 ** 31 + 100, -2 ** 31 + 255]:
        t = c + 2 ** 31                                     # 0..255: pixels whose mean reaches t wrap to "never"
        assert np.array_equal(px < _wrap_i32(mean - c), np.broadcast_to(mean < t, (256, 256)))


def test_oracle_wraps_like_a_release_build():
    from oracle import oracle as O
    import ref_literal
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (23, 31, 1), dtype=np.uint8)
    for thr in (-1e12, -2147483648.0, -2147483548.0, -2147483393.0, -2147483392.0, 2147483647.0, 1e12):
        got = O.detect_contamination(img, 1.0, 1e9, thr)
        lit = ref_literal.detect_contamination(img, 1.0, 1e9, thr)
        assert np.array_equal(got.mask, np.array(lit["binary"], np.uint8)), thr
    assert int(O.detect_contamination(img, threshold=-1e12).mask.sum()) == 0          # wraps everywhere: never foreground
    assert int((O.detect_contamination(img, threshold=-2147483392.0).mask == 255).all())  # one above: no wrap, always


def test_flat_tile_bound():
    """mean - blur <= max - min over the neighbourhood, so range <= c (c >= 0) implies an empty mask."""
    from oracle import oracle as O
    rng = np.random.default_rng(1)
    for c in (0, 3, 25):
        img = rng.integers(100, 100 + c + 1, (60, 80), dtype=np.uint8)   # range <= c everywhere
        assert int(O.detect_contamination(img, threshold=float(c)).mask.sum()) == 0


def test_nibble_gather_multiplier():
    for v in range(16):
        word = sum(((v >> k) & 1) * 0xFF << (8 * k) for k in range(4))
        assert (((word & 0x01010101) * 0x10204080) & 0xFFFFFFFF) >> 28 == v


def _run_start(m, bit):
    zeros_below = ~m & ((1 << bit) - 1) & 0xFFFFFFFF
    return zeros_below.bit_length() if zeros_below else 0


def test_run_start_and_word_run_decomposition():
    rng = np.random.default_rng(2)
    for m in [0xFFFFFFFF, 1, 0x80000000, 0xF0F0F0F0, 0x55555555] + [int(x) for x in rng.integers(0, 2 ** 32, 200)]:
        bits = [(m >> i) & 1 for i in range(32)]
        for b in range(32):
            if bits[b]:
                s = b
                while s > 0 and bits[s - 1]:
                    s -= 1
                assert _run_start(m, b) == s
        starts = m & ~(m << 1) & 0xFFFFFFFF
        assert bin(starts).count("1") == sum(1 for i in range(32) if bits[i] and (i == 0 or not bits[i - 1]))


def test_union_find_by_min_gives_raster_first_labels():
    """Host model of K2-K5: word-runs as nodes, union by minimum index, rank of roots in raster order."""
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    for shape, p in [((20, 70), 0.5), ((9, 33), 0.62), ((16, 64), 0.3)]:
        m = (rng.random(shape) < p)
        h, w = shape
        parent = {}

        def find(a):
            while parent[a] != a:
                a = parent[a]
            return a

        def union(a, b):
            a, b = find(a), find(b)
            if a != b:
                parent[max(a, b)] = min(a, b)

        node = -np.ones(shape, np.int64)
        for y in range(h):
            for x in range(w):
                if m[y, x]:
                    if x % 32 != 0 and m[y, x - 1]:
                        node[y, x] = node[y, x - 1]
                    else:
                        node[y, x] = y * w + x
                        parent[y * w + x] = y * w + x
        for y in range(h):
            for x in range(w):
                if m[y, x]:
                    if x > 0 and m[y, x - 1] and node[y, x - 1] != node[y, x]:
                        union(int(node[y, x]), int(node[y, x - 1]))
                    if y > 0 and m[y - 1, x]:
                        union(int(node[y, x]), int(node[y - 1, x]))
        roots = sorted({find(a) for a in parent})
        rank = {r: i + 1 for i, r in enumerate(roots)}
        lab = np.zeros(shape, np.int32)
        for y in range(h):
            for x in range(w):
                if m[y, x]:
                    lab[y, x] = rank[find(int(node[y, x]))]
        ref, _ = O.label4(m.astype(np.uint8) * 255)
        assert np.array_equal(lab, ref)


def test_gaussian_q8_kernels_sum_to_256():
    from oracle import oracle as O
    for k in range(1, 32, 2):
        for s in (0.0, 0.5, 1.0, 2.0, 3.0, 7.5):
            kk = O.gaussian_kernel_q8(k, s)
            assert int(kk.sum()) == 256 and np.array_equal(kk, kk[::-1])


def test_gaussian_packed_pass_model_equals_the_oracle():
    """The arithmetic of K1's packed Gaussian taps (k_preprocess.cu, gauss_fast_blur), modelled in numpy on the interior of
    an image: vertical pass first on the bytes (every lane stays below 2^16: a u16x2 lane never carries), horizontal pass as
    sums of two-tap products with a 32-bit accumulator that starts at 32768 (IDP.2A), blur = acc >> 16 -- equal to the
    oracle's rows-then-columns cv2.GaussianBlur restatement for every odd kernel size 3..15 whose taps fit a byte."""
    from oracle import oracle as O
    rng = np.random.default_rng(77)
    img = rng.integers(0, 256, (96, 128), dtype=np.uint8)
    img[:8] = 255                                            # saturated region: the lanes' worst case
    for k in range(3, 16, 2):
        for sg in (0.0, 0.7, 2.5, 0.1):
            taps = O.gaussian_kernel_q8(k, sg).astype(np.uint32)
            if taps.max() > 255:
                assert sg == 0.1                             # (the kernel sends these to the generic taps)
                continue
            r = k // 2
            v = np.zeros(img.shape, np.uint32)
            for t in range(k):                               # vertical: v[y] = sum_t taps[t] * img[y - r + t]
                v[r:-r] += taps[t] * img[t:img.shape[0] - 2 * r + t].astype(np.uint32)
            assert int(v.max()) <= 255 * 256 < 65536
            acc = np.full(img.shape, 32768, np.uint32)
            for m in range((k + 1) // 2):                    # horizontal, two taps per step (the odd tail pairs with a 0)
                t0, t1 = 2 * m, 2 * m + 1
                acc[:, r:-r] += taps[t0] * v[:, t0:img.shape[1] - 2 * r + t0]
                if t1 < k:
                    acc[:, r:-r] += taps[t1] * v[:, t1:img.shape[1] - 2 * r + t1]
            got = (acc >> 16).astype(np.uint8)
            ref = O.gaussian_blur(img, k, sg)
            assert np.array_equal(got[r:-r, r:-r], ref[r:-r, r:-r]), (k, sg)


def test_bench_reads_the_gpu_local_cpu_list(tmp_path, monkeypatch):
    """bench.py binds the thread that allocates the pinned staging buffers to the GPU's NUMA node (sysfs local_cpulist);
    without the sysfs entry, or when the list is the whole affinity mask, it leaves the affinity alone."""
    import os
    import sys
    import types
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    class Props:
        pci_domain_id, pci_bus_id, pci_device_id = 0, 0x1b, 0

    fake = types.SimpleNamespace(cuda=types.SimpleNamespace(get_device_properties=lambda i: Props()))
    before = os.sched_getaffinity(0)
    assert bench.gpu_numa_affinity(fake, 0) is None or os.sched_getaffinity(0) <= before   # (no such PCI function here)
    bench.numa_restore(None)
    os.sched_setaffinity(0, before)
    real_open = open
    first = sorted(before)[0]

    def fake_open(path, *a, **k):
        if str(path).endswith("0000:1b:00.0/local_cpulist"):
            import io
            return io.StringIO(f"{first}-{first}\n")
        return real_open(path, *a, **k)
    monkeypatch.setattr("builtins.open", fake_open)
    state = bench.gpu_numa_affinity(fake, 0)
    if len(before) > 1:
        assert state is not None and os.sched_getaffinity(0) == {first} and state["before"] == before
    bench.numa_restore(state)
    assert os.sched_getaffinity(0) == before


def test_synthetic_frames_are_deterministic():
    import synth
    a = synth.bottle_frame(128, 160, 3)
    b = synth.bottle_frame(128, 160, 3)
    assert np.array_equal(a, b) and a.dtype == np.uint8 and a.shape == (128, 160)
    assert not np.array_equal(a, synth.bottle_frame(128, 160, 4))
    assert synth.bottle_batch(3, 64, 96, start_index=5).shape == (3, 64, 96)
    hc = synth.high_contamination_frame(300, 400, 0)
    assert hc.shape == (300, 400) and (hc < 100).sum() > 1000
