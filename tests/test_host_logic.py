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


def test_synthetic_frames_are_deterministic():
    import synth
    a = synth.bottle_frame(128, 160, 3)
    b = synth.bottle_frame(128, 160, 3)
    assert np.array_equal(a, b) and a.dtype == np.uint8 and a.shape == (128, 160)
    assert not np.array_equal(a, synth.bottle_frame(128, 160, 4))
    assert synth.bottle_batch(3, 64, 96, start_index=5).shape == (3, 64, 96)
    hc = synth.high_contamination_frame(300, 400, 0)
    assert hc.shape == (300, 400) and (hc < 100).sum() > 1000
