// k_preprocess.cu -- K1: fused preprocess for the contamination path.
//
//   gray (u8) --5x5 box mean, floor, interior only--> blur --11x11 edge-truncated mean, floor--> mean
//   mask = 255 if (i32)blur < mean - c else 0
//
// Restates rust/heimdall-core/src/detection.rs:162-182 (blur) and :184-213 (adaptive threshold); the same loops
// appear at processing.rs:269-289 / 291-320 (c = 15) and processing.rs:131-164 (c = 2, selectable `inverse`).
//
// Division-free test (exact):  px < floor(S/cnt) - c  <=>  (px + c + 1) * cnt <= S      (cnt > 0, all integers)
//                              px > floor(S/cnt) - c  <=>  !((px + c) * cnt <= S)
// Box mean: floor(s/25) == (s*5243) >> 17 for every s <= 6375 (checked exhaustively in tests/test_host_logic.py).
//
// Outputs per tile, written exactly once: the u8 mask, the bit-packed mask (1 bit/px) and the i32 label plane
// initialised for the union-find CCL (0 = background or non-node, p+1 at the first pixel of every word-run, where
// p = y*w + x and a word-run is a maximal horizontal run inside one 32-pixel word).
//
// Sparsity fast path (exact, not an approximation): blur and mean are floor-averages of gray values inside the
// 15x15 neighbourhood N(p), so mean - blur <= max_N(gray) - min_N(gray).  If that range is <= c over a whole tile
// plus halo, no pixel of the tile can satisfy blur < mean - c and the tile is written as zeros without computing
// the sums.  On bottle frames most tiles are flat, which makes this kernel HBM-bound: 1 B/px read, 5.125 B/px written.
#include <cuda.h>

#include <algorithm>
#include <cudaTypedefs.h>

#include "hv_common.cuh"
#include "expand_tile.cuh"

namespace hv {

namespace {

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// exact floor(s / 25) for s <= 6375
__device__ __forceinline__ uint32_t div25(uint32_t s) { return (s * 5243u) >> 17; }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the thread is parked until the phase completes (or the hint expires) instead of
// re-issuing the test -- in the r03c capture 28 % of all executed warp instructions were this spin loop (the producer
// warp of every CTA waits nearly all the time), competing for issue slots with the warps that have work.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, uint32_t hint_ns) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(hint_ns)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tmap, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap *tmap, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Barriers of the tile pipeline.  NAMED = false: the whole CTA (__syncthreads).  NAMED = true: the 256 consumer threads of
// the warp-specialised TMA kernel only (named barrier 1), so that the producer warp never has to take part.
template <bool NAMED>
__device__ __forceinline__ void tile_sync() {
    if (NAMED)
        asm volatile("bar.sync 1, 256;" ::: "memory");
    else
        __syncthreads();
}
template <bool NAMED>
__device__ __forceinline__ bool tile_sync_or(uint32_t v) {
    if (NAMED) {
        uint32_t r;
        asm volatile(
            "{\n"
            ".reg .pred p, q;\n"
            "setp.ne.u32 p, %1, 0;\n"
            "bar.red.or.pred q, 1, 256, p;\n"
            "selp.u32 %0, 1, 0, q;\n"
            "}\n"
            : "=r"(r)
            : "r"(v)
            : "memory");
        return r != 0;
    }
    return __syncthreads_or((int)v) != 0;
}

// Shared-memory plan of one tile (TW x TH outputs, 256 threads).  Column index i of s_g / s_bl is image column
// x0 - 8 + i; row r of s_g is image row y0 - HALO + r; row r of s_v / s_bl / s_h11 is image row y0 - 5 + r.
//   s_g   [GH][GW]  u8   gray tile + halo                                  (dead after the blur phase)
//   s_f   [TH][TW]  u8   mask bytes of the tile            -- aliases s_g
//   s_v   [BH][VP]  u16  vertical 5-sums, column i stored at i + 4         (fast path only, dead after the blur phase)
//   s_h11 [BH][TW]  u16  horizontal 11-sums of the blur    -- aliases s_v
//   s_bl  [BH][BW]  u8   blurred tile + 5-px ring (logical columns)
template <int TW, int TH, int RB, int HXP = 8>
struct Tile {
    static constexpr int HALO = RB + kAdaptHalf;  // rows of gray needed above/below
    static constexpr int HX = HXP;                // column halo of the staged gray tile: 8 (aligned 8-byte loads) or 16
                                                  // (TMA: the box must start on a 16-byte boundary)
    static constexpr int GW = TW + 2 * HX;        // pitch of s_g
    static constexpr int GOFF = HX - 8;           // s_g column of logical column 0 (logical column i = image x0 - 8 + i)
    static constexpr int BW = TW + 16;            // logical width = pitch of s_bl
    static constexpr int GH = TH + 2 * HALO;
    static constexpr int BH = TH + 2 * kAdaptHalf;
    static constexpr int VP = BW + 8;  // pitch of s_v in u16 (column i is stored at i + 4; 4 spare on each side)
    static constexpr int G_BYTES = GH * GW;
    static constexpr int GR_BYTES = RB > 2 ? GH * BW * 2 : 0;  // Gaussian variant: row-filtered u16 tile, all GH rows
    static constexpr int U1_BASE = (BH * VP * 2 > BH * TW * 2) ? BH * VP * 2 : BH * TW * 2;
    static constexpr int U1_BYTES = GR_BYTES > U1_BASE ? GR_BYTES : U1_BASE;
    static constexpr int BL_BYTES = BH * BW;
    static_assert(HX >= 8 && (HX % 8) == 0, "column halo");
    static_assert(TH * TW <= G_BYTES, "s_f must fit inside s_g");
    static_assert((G_BYTES % 16) == 0 && (U1_BYTES % 16) == 0, "alignment");
};

// ---- fast path, interior tiles only (every blur pixel is an interior pixel, every window has 121 pixels) ---------
// Packed u16x2 arithmetic: sums of u8 never overflow a 16-bit lane (5x5: 6375, 11x11 of blur: 30855), so plain 32-bit
// adds/subs act on both lanes at once; VIMNMX.U16x2 gives the lane-wise compare for the threshold test.
template <int TW, int TH, int HXP, bool NAMED>
__device__ __forceinline__ void fast_blur_rb2(uint8_t *g_raw, uint8_t *u1_raw, uint8_t *bl_raw, int tid,
                                              uint64_t *empty_bar) {
    using T = Tile<TW, TH, 2, HXP>;
    uint8_t(*s_g)[T::GW] = reinterpret_cast<uint8_t(*)[T::GW]>(g_raw);
    uint16_t(*s_v)[T::VP] = reinterpret_cast<uint16_t(*)[T::VP]>(u1_raw);
    uint8_t(*s_bl)[T::BW] = reinterpret_cast<uint8_t(*)[T::BW]>(bl_raw);
    static_assert(TW == 128 && TH == 32, "thread mappings below are written for 128x32 tiles");

    // A. vertical 5-sums: thread = (column quad q, row segment of 7) -> 36 x 6 = 216 threads
    if (tid < 36 * 6) {
        const int q = tid % 36, seg = tid / 36;
        const int r0 = seg * 7;
        uint32_t lo[5], hi[5];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t v = *reinterpret_cast<const uint32_t *>(&s_g[r0 + k][4 * q + T::GOFF]);
            lo[k] = prmt(v, 0, 0x4140);
            hi[k] = prmt(v, 0, 0x4342);
        }
        uint32_t alo = lo[0] + lo[1] + lo[2] + lo[3], ahi = hi[0] + hi[1] + hi[2] + hi[3];
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const uint32_t v = *reinterpret_cast<const uint32_t *>(&s_g[r0 + k + 4][4 * q + T::GOFF]);
            const uint32_t nlo = prmt(v, 0, 0x4140), nhi = prmt(v, 0, 0x4342);
            alo += nlo;
            ahi += nhi;
            *reinterpret_cast<uint2 *>(&s_v[r0 + k][4 * q + 4]) = make_uint2(alo, ahi);  // column i at index i + 4
            alo -= lo[k % 5];
            ahi -= hi[k % 5];
            lo[(k + 4) % 5] = nlo;
            hi[(k + 4) % 5] = nhi;
        }
    }
    tile_sync<NAMED>();
    if (empty_bar && tid == 0) mbar_arrive(empty_bar);  // the gray stage is dead: the producer may refill it
    // B. horizontal 5-sums + /25 -> s_bl.  thread = (row, group of 36 columns) -> 42 x 4 = 168 threads.
    //    s_v stores column i at index i + 4, so the window i-2..i+2 of output i is s_v[i+2 .. i+6]; a group needs
    //    s_v[36g + 2 .. 36g + 42), loaded as 11 aligned 8-byte words starting at 36g.
    if (tid < 42 * 4) {
        const int g = tid & 3, r = tid >> 2;
        const uint16_t *vrow = &s_v[r][36 * g];
        uint32_t pk[22];
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const uint2 t = *reinterpret_cast<const uint2 *>(vrow + 4 * k);
            pk[2 * k] = t.x;
            pk[2 * k + 1] = t.y;
        }
        // element e (0..43) = s_v[36g + e]; outputs i = 36g + j (j = 0..35) use elements j+2 .. j+6
        auto el = [&](int e) -> uint32_t { return (e & 1) ? (pk[e >> 1] >> 16) : (pk[e >> 1] & 0xffffu); };
        uint32_t s = el(2) + el(3) + el(4) + el(5);
        uint32_t *brow = reinterpret_cast<uint32_t *>(&s_bl[r][36 * g]);
#pragma unroll
        for (int j4 = 0; j4 < 9; j4++) {
            uint32_t q[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int j = 4 * j4 + u;
                s += el(j + 6);
                q[u] = __umulhi(s, 5243u << 15);  // floor(s / 25)
                s -= el(j + 2);
            }
            brow[j4] = q[0] | (q[1] << 8) | (q[2] << 16) | (q[3] << 24);
        }
    }
    tile_sync<NAMED>();
}

// Phases C and D of the fast path: 11x11 window sums of the blurred tile (+ 5-px ring) in s_bl and the threshold test.
// BWP = pitch of s_bl.
template <int TW, int TH, int BWP, bool NAMED>
__device__ __forceinline__ void fast_threshold(uint8_t *f_raw, uint8_t *u1_raw, uint8_t *bl_raw, int tid, int cth) {
    uint8_t(*s_f)[TW] = reinterpret_cast<uint8_t(*)[TW]>(f_raw);
    uint16_t(*s_h11)[TW] = reinterpret_cast<uint16_t(*)[TW]>(u1_raw);
    uint8_t(*s_bl)[BWP] = reinterpret_cast<uint8_t(*)[BWP]>(bl_raw);
    static_assert(TW == 128 && TH == 32, "thread mappings below are written for 128x32 tiles");
    // C. horizontal 11-sums of the blur.  thread = (row, group of 32 output columns) -> 42 x 4 = 168 threads.
    //    output column c uses s_bl[c + 3 .. c + 13]; a group needs bytes 32g + 3 .. 32g + 44 = words 8g .. 8g + 11.
    if (tid < 42 * 4) {
        const int g = tid & 3, r = tid >> 2;
        const uint32_t *brow = reinterpret_cast<const uint32_t *>(&s_bl[r][32 * g]);
        uint32_t wd[12];
#pragma unroll
        for (int k = 0; k < 12; k++) wd[k] = brow[k];
        auto by = [&](int e) -> uint32_t { return (wd[e >> 2] >> (8 * (e & 3))) & 0xffu; };
        uint32_t s = 0;
#pragma unroll
        for (int e = 3; e < 13; e++) s += by(e);
        uint32_t *hrow = reinterpret_cast<uint32_t *>(&s_h11[r][32 * g]);
#pragma unroll
        for (int c2 = 0; c2 < 16; c2++) {
            s += by(2 * c2 + 13);
            const uint32_t a = s;
            s -= by(2 * c2 + 3);
            s += by(2 * c2 + 14);
            hrow[c2] = a | (s << 16);
            s -= by(2 * c2 + 4);
        }
    }
    tile_sync<NAMED>();
    // D. vertical 11-sums + threshold test.  thread = (column quad, segment of 8 rows) -> 32 x 4 = 128 threads.  (Four rows
    //    per thread keep all 256 threads busy, 14 row steps each instead of 18 for half of them: 39.65 vs 39.71 us per
    //    headline batch, and 55 % more instructions in this phase for the dense frames that are issue-bound -- not taken.)
    //    fg  <=>  (px + c + 1) * 121 <= S  <=>  S + 1 > px * 121 + (c + 1) * 121      (all lanes < 65536 for 0 <= c <= 255)
    constexpr int DSEG = 8;
    if (tid < 32 * (32 / DSEG)) {
        const int q = tid & 31, seg = tid >> 5;
        const int r0 = seg * DSEG;
        const uint32_t K2 = (uint32_t)((cth + 1) * 121) * 0x00010001u;
        uint32_t rl[11], rh[11];
        uint32_t alo = 0x00010001u, ahi = 0x00010001u;  // the "+ 1"
#pragma unroll
        for (int k = 0; k < 10; k++) {
            const uint2 t = *reinterpret_cast<const uint2 *>(&s_h11[r0 + k][4 * q]);
            rl[k] = t.x;
            rh[k] = t.y;
            alo += t.x;
            ahi += t.y;
        }
#pragma unroll
        for (int k = 0; k < DSEG; k++) {
            const uint2 t = *reinterpret_cast<const uint2 *>(&s_h11[r0 + k + 10][4 * q]);
            alo += t.x;
            ahi += t.y;
            const uint32_t px4 = *reinterpret_cast<const uint32_t *>(&s_bl[r0 + k + 5][4 * q + 8]);
            const uint32_t xlo = prmt(px4, 0, 0x4140) * 121u + K2, xhi = prmt(px4, 0, 0x4342) * 121u + K2;
            const uint32_t dlo = __vminu2(__vmaxu2(alo, xlo) - xlo, 0x00010001u);
            const uint32_t dhi = __vminu2(__vmaxu2(ahi, xhi) - xhi, 0x00010001u);
            *reinterpret_cast<uint32_t *>(&s_f[r0 + k][4 * q]) = prmt(dlo, dhi, 0x6420) * 255u;
            alo -= rl[k % 11];
            ahi -= rh[k % 11];
            rl[(k + 10) % 11] = t.x;
            rh[(k + 10) % 11] = t.y;
        }
    }
    tile_sync<NAMED>();
}

// ---- Gaussian variant (A7): cv2.GaussianBlur(k, k, sigma) of the staged tile, k <= 15 ------------------------------
// OpenCV's CV_8U path: rows in 8.8 fixed point (u16), columns in 16.16, result (s + 32768) >> 16, BORDER_REFLECT_101.
// The staged box has RB = 7 more rows than the box-blur variant needs, its out-of-image cells have been filled by
// reflection (see the consumer loop), so both passes are uniform.  Row pass over all GH rows and the BW logical columns
// into u1 (u16), column pass over the BH rows of the blur ring into s_bl; blur pixels outside the image are 0 (the
// truncated-window threshold test sums them).  4 outputs per thread-iteration share their taps' loads.
constexpr int kGaussRB = 7;       // stage halo of the variant for kernel sizes 9..15
constexpr int kGaussRBSmall = 3;  // ... for kernel sizes <= 7
template <int TW, int TH, int RB, int HXP, bool NAMED>
__device__ __forceinline__ void gauss_blur_tile(const BatchView &b, const PreprocessParams &p, uint8_t *g_raw, uint8_t *u1_raw,
                                                uint8_t *bl_raw, int f, int x0, int y0, int tid, uint64_t *empty_bar) {
    using T = Tile<TW, TH, RB, HXP>;
    uint8_t(*s_g)[T::GW] = reinterpret_cast<uint8_t(*)[T::GW]>(g_raw);
    uint16_t(*s_r)[T::BW] = reinterpret_cast<uint16_t(*)[T::BW]>(u1_raw);
    uint8_t(*s_bl)[T::BW] = reinterpret_cast<uint8_t(*)[T::BW]>(bl_raw);
    const int ks = p.gauss_ksize, R = ks >> 1;
    const int H = b.h, W = b.w;
    const int r_skip = RB - R;  // the column pass reads row-filtered rows r_skip .. GH - 1 - r_skip only
    for (int idx = tid; idx < (T::GH - 2 * r_skip) * (T::BW / 4); idx += 256) {
        const int r = r_skip + idx / (T::BW / 4), i = 4 * (idx % (T::BW / 4));
        const uint8_t *src = &s_g[r][T::GOFF + i - R];  // tap t of output j reads src[j + t]
        uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        uint32_t w0 = src[0], w1 = src[1], w2 = src[2];
        for (int t = 0; t < ks; t++) {
            const uint32_t w3 = src[t + 3], kk = p.gk[t];
            a0 += kk * w0, a1 += kk * w1, a2 += kk * w2, a3 += kk * w3;
            w0 = w1, w1 = w2, w2 = w3;
        }
        *reinterpret_cast<uint2 *>(&s_r[r][i]) = make_uint2(min(a0, 65535u) | (min(a1, 65535u) << 16),
                                                            min(a2, 65535u) | (min(a3, 65535u) << 16));
    }
    tile_sync<NAMED>();
    if (empty_bar && tid == 0) mbar_arrive(empty_bar);  // the gray stage is dead: the producer may refill it
    for (int idx = tid; idx < T::BH * (T::BW / 4); idx += 256) {
        const int r = idx / (T::BW / 4), i = 4 * (idx - r * (T::BW / 4));
        uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        for (int t = 0; t < ks; t++) {
            const uint2 v = *reinterpret_cast<const uint2 *>(&s_r[r + RB - R + t][i]);
            const uint32_t kk = p.gk[t];
            a0 += kk * (v.x & 0xffffu), a1 += kk * (v.x >> 16), a2 += kk * (v.y & 0xffffu), a3 += kk * (v.y >> 16);
        }
        const int gy = y0 - kAdaptHalf + r, gx = x0 - 8 + i;
        uint32_t o[4] = {min((a0 + 32768u) >> 16, 255u), min((a1 + 32768u) >> 16, 255u), min((a2 + 32768u) >> 16, 255u),
                         min((a3 + 32768u) >> 16, 255u)};
        const bool row_in = gy >= 0 && gy < H;
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (!row_in || gx + j < 0 || gx + j >= W) o[j] = 0;
        *reinterpret_cast<uint32_t *>(&s_bl[r][i]) = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
        if (p.write_blur && row_in && r >= kAdaptHalf && r < kAdaptHalf + TH) {
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (i + j >= 8 && i + j < 8 + TW && gx + j < W) b.blur[((size_t)f * H + gy) * W + gx + j] = (uint8_t)o[j];
        }
    }
    tile_sync<NAMED>();
}

// cv2.GaussianBlur(5, 5, 0) -- the reference's own setting (heimdall/detectors/contamination_detector.py:66) -- on interior
// tiles, in packed u16x2 arithmetic like the box blur.  OpenCV's 8.8 taps are [16, 64, 96, 64, 16] = 16 * [1, 4, 6, 4, 1]
// per axis and its result is (sum + 32768) >> 16; both passes are exact integer sums, so the order of the passes is free and
// the common factor 256 can be taken out: v = (g0 + g4) + 4 (g1 + g3) + 6 g2 over rows (<= 4080), h likewise over columns
// (<= 65280: fits a 16-bit lane), blur = (h + 128) >> 8.  Blur row r (image row y0 - 5 + r) is centred on stage row r + 7.
template <int TW, int TH, int RB, int HXP, bool NAMED>
__device__ __forceinline__ void gauss5_fast_blur(uint8_t *g_raw, uint8_t *u1_raw, uint8_t *bl_raw, int tid, uint64_t *empty_bar) {
    using T = Tile<TW, TH, RB, HXP>;
    uint8_t(*s_g)[T::GW] = reinterpret_cast<uint8_t(*)[T::GW]>(g_raw);
    uint16_t(*s_v)[T::VP] = reinterpret_cast<uint16_t(*)[T::VP]>(u1_raw);
    uint8_t(*s_bl)[T::BW] = reinterpret_cast<uint8_t(*)[T::BW]>(bl_raw);
    static_assert(TW == 128 && TH == 32 && T::GOFF == 8 && T::BH == 42 && T::VP >= 152, "thread mappings below");
    // A. vertical taps: thread = (quad q of 38 over logical columns -4 .. 147, segment of 7 blur rows) -> 228 threads.
    //    Logical column i sits at stage column 8 + i and is stored at s_v index i + 4.
    if (tid < 38 * 6) {
        const int q = tid % 38, seg = tid / 38;
        const int r0 = seg * 7;
        uint32_t lo[11], hi[11];
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const uint32_t v = *reinterpret_cast<const uint32_t *>(&s_g[r0 + RB - 2 + k][4 + 4 * q]);
            lo[k] = prmt(v, 0, 0x4140);
            hi[k] = prmt(v, 0, 0x4342);
        }
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const uint32_t vlo = (lo[k] + lo[k + 4]) + ((lo[k + 1] + lo[k + 3]) << 2) + lo[k + 2] * 6u;
            const uint32_t vhi = (hi[k] + hi[k + 4]) + ((hi[k + 1] + hi[k + 3]) << 2) + hi[k + 2] * 6u;
            *reinterpret_cast<uint2 *>(&s_v[r0 + k][4 * q]) = make_uint2(vlo, vhi);
        }
    }
    tile_sync<NAMED>();
    if (empty_bar && tid == 0) mbar_arrive(empty_bar);  // the gray stage is dead: the producer may refill it
    // B. horizontal taps + rounding: thread = (row, group of 36 columns) -> 42 x 4 = 168 threads; output i = 36g + j reads
    //    s_v indices i + 2 .. i + 6
    if (tid < 42 * 4) {
        const int g = tid & 3, r = tid >> 2;
        const uint16_t *vrow = &s_v[r][36 * g];
        uint32_t pk[22];
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const uint2 t = *reinterpret_cast<const uint2 *>(vrow + 4 * k);
            pk[2 * k] = t.x;
            pk[2 * k + 1] = t.y;
        }
        auto el = [&](int e) -> uint32_t { return (e & 1) ? (pk[e >> 1] >> 16) : (pk[e >> 1] & 0xffffu); };
        uint32_t *brow = reinterpret_cast<uint32_t *>(&s_bl[r][36 * g]);
#pragma unroll
        for (int j4 = 0; j4 < 9; j4++) {
            uint32_t o[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int j = 4 * j4 + u;
                const uint32_t hsum = (el(j + 2) + el(j + 6)) + ((el(j + 3) + el(j + 5)) << 2) + el(j + 4) * 6u;
                o[u] = (hsum + 128u) >> 8;
            }
            brow[j4] = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
        }
    }
    tile_sync<NAMED>();
}


// cv2.GaussianBlur(KS, KS, sigma) for any odd KS in 3..15 whose taps fit a byte, interior tiles.  Both passes of OpenCV's
// CV_8U path are exact integer sums (8.8 taps, rows into u16, columns into 16.16, one rounding at the end), so their order
// is free.  Vertical pass first, on the bytes, in packed u16x2 lanes (a lane never exceeds 255 * 256): a thread owns a column
// quad and seven output rows and streams the 7 + KS - 1 input rows through its fourteen accumulators, fully unrolled.
// Horizontal pass on the u16 sums with IDP.2A (two taps per instruction, two adjacent columns per 32-bit word): an item is
// (row, 24 output columns); the outputs whose window starts on an even element use the words as loaded, the others the same
// words shifted by one element, converted in place between the two halves.  acc starts at 32768: blur = acc >> 16.
// s_v element index = stage column (logical column + 8), pitch 160.  Blur row r is centred on stage row r + RB.
template <int TW, int TH, int RB, int HXP, bool NAMED, int KS>
__device__ __forceinline__ void gauss_fast_blur(const PreprocessParams &p, uint8_t *g_raw, uint8_t *u1_raw, uint8_t *bl_raw, int tid,
                                                uint64_t *empty_bar) {
    using T = Tile<TW, TH, RB, HXP>;
    constexpr int R = KS / 2, VPG = 160, S = 7;
    static_assert(TW == 128 && TH == 32 && HXP == 16 && T::GW == 160 && T::BH == 42 && T::BW == 144, "thread mappings below");
    static_assert(R <= RB && R <= 7 && T::BH * VPG * 2 <= T::U1_BYTES, "halo / scratch size");
    uint8_t(*s_g)[T::GW] = reinterpret_cast<uint8_t(*)[T::GW]>(g_raw);
    uint16_t(*s_v)[VPG] = reinterpret_cast<uint16_t(*)[VPG]>(u1_raw);
    uint8_t(*s_bl)[T::BW] = reinterpret_cast<uint8_t(*)[T::BW]>(bl_raw);
    uint32_t tap[R + 1];  // tap[d] = weight at distance d from the centre
#pragma unroll
    for (int d = 0; d <= R; d++) tap[d] = p.gk[R - d];
    // A. vertical: 40 column quads x 6 segments of 7 blur rows = 240 threads
    if (tid < 40 * 6) {
        const int q = tid % 40, seg = tid / 40;
        const int r0 = seg * S;
        uint32_t alo[S], ahi[S];
#pragma unroll
        for (int k = 0; k < S; k++) alo[k] = 0, ahi[k] = 0;
#pragma unroll
        for (int j = 0; j < S + KS - 1; j++) {
            const uint32_t v = *reinterpret_cast<const uint32_t *>(&s_g[r0 + RB - R + j][4 * q]);
            const uint32_t lo = prmt(v, 0, 0x4140), hi = prmt(v, 0, 0x4342);
#pragma unroll
            for (int k = 0; k < S; k++) {
                const int t = j - k;  // tap index of input row j for output row k
                if (t >= 0 && t < KS) {
                    const uint32_t wgt = tap[t < R ? R - t : t - R];
                    alo[k] += wgt * lo;
                    ahi[k] += wgt * hi;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < S; k++) *reinterpret_cast<uint2 *>(&s_v[r0 + k][4 * q]) = make_uint2(alo[k], ahi[k]);
    }
    tile_sync<NAMED>();
    if (empty_bar && tid == 0) mbar_arrive(empty_bar);  // the gray stage is dead: the producer may refill it
    // B. horizontal: 42 rows x 6 groups of 24 outputs = 252 threads.  Output j of group g is logical column 24g + j; its
    //    window starts at element 24g + j + 8 - R.  The group loads elements 24g .. 24g + 47 (24 words).
    if (tid < 42 * 6) {
        const int g = tid % 6, r = tid / 6;
        uint32_t pk[24];
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(&s_v[r][24 * g]);
#pragma unroll
            for (int k = 0; k < 6; k++) {
                const uint4 t = src[k];
                pk[4 * k] = t.x, pk[4 * k + 1] = t.y, pk[4 * k + 2] = t.z, pk[4 * k + 3] = t.w;
            }
        }
        // taps of the window, four to a register: byte m of tw[i] = weight of window element 4i + m (0 beyond KS - 1)
        constexpr int NT = (KS + 3) / 4;
        uint32_t tw[NT];
#pragma unroll
        for (int i = 0; i < NT; i++) {
            uint32_t w4 = 0;
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const int t = 4 * i + m;
                if (t < KS) w4 |= tap[t < R ? R - t : t - R] << (8 * m);
            }
            tw[i] = w4;
        }
        uint32_t out[6];
#pragma unroll
        for (int k = 0; k < 6; k++) out[k] = 0;
#pragma unroll
        for (int half = 0; half < 2; half++) {
            if (half == 1) {  // words shifted by one element: pk[i] = elements (2i + 1, 2i + 2)
#pragma unroll
                for (int i = 0; i < 23; i++) pk[i] = prmt(pk[i], pk[i + 1], 0x5432);
            }
#pragma unroll
            for (int j = 0; j < 24; j++) {
                const int s0 = j + 8 - R;  // first window element
                if ((s0 & 1) != half) continue;
                const int w0 = s0 >> 1;    // (half == 1: shifted word w0 holds elements s0, s0 + 1)
                uint32_t acc = 32768u;
#pragma unroll
                for (int m = 0; m < (KS + 1) / 2; m++) {
                    if (m & 1)
                        acc = __dp2a_hi(pk[w0 + m], tw[m >> 1], acc);
                    else
                        acc = __dp2a_lo(pk[w0 + m], tw[m >> 1], acc);
                }
                out[j >> 2] |= (acc >> 16) << (8 * (j & 3));
            }
        }
        uint32_t *brow = reinterpret_cast<uint32_t *>(&s_bl[r][24 * g]);
#pragma unroll
        for (int k = 0; k < 6; k++) brow[k] = out[k];
    }
    tile_sync<NAMED>();
}

// Everything after the gray tile (+ halo, zero outside the image) sits in shared memory and the flat decision is known:
// blur + threshold (fast or generic path) and the three outputs.  Block-uniform control flow; contains barriers.
template <int TW, int TH, int RB, int HXP, bool NAMED>
__device__ __forceinline__ void tile_compute_and_store(const BatchView &b, const PreprocessParams &p, uint32_t *bits_out,
                                                       uint8_t *g_raw, uint8_t *f_raw, uint8_t *u1_raw, uint8_t *bl_raw,
                                                       int f, int tile_x, int x0, int y0, bool flat, int tid,
                                                       uint64_t *empty_bar) {
    using T = Tile<TW, TH, RB, HXP>;
    uint8_t(*s_g)[T::GW] = reinterpret_cast<uint8_t(*)[T::GW]>(g_raw);
    uint8_t(*s_f)[TW] = reinterpret_cast<uint8_t(*)[TW]>(f_raw);
    uint16_t(*s_h11)[TW] = reinterpret_cast<uint16_t(*)[TW]>(u1_raw);
    uint8_t(*s_bl)[T::BW] = reinterpret_cast<uint8_t(*)[T::BW]>(bl_raw);
    const int H = b.h, W = b.w;
    const int cth = p.c_thresh;
    const bool vec = (W & 3) == 0;
    uint8_t *mask = b.mask + (size_t)f * H * W;
    int32_t *labels = b.labels + (size_t)f * H * W;
    // interior tile: the blur ring (tile +- 5) lies >= RB pixels inside the image, so no border rule applies anywhere
    const bool interior = x0 >= T::HALO && y0 >= T::HALO && x0 + TW + T::HALO <= W && y0 + TH + T::HALO <= H;

    if (flat && empty_bar && tid == 0) mbar_arrive(empty_bar);  // every thread is past its flat-test reads of the stage
    if (!flat) {
        const bool fast_thr = TW == 128 && TH == 32 && interior && p.inverse && cth >= 0 && cth <= 255 && !p.force_generic;
        if (RB > 2 && p.gauss_ksize > 0) {
            bool done = false;
            if constexpr (RB > 2 && HXP == 16) {
                if (fast_thr && !p.write_blur) {
                    const int ks = p.gauss_ksize;
                    if (ks == 5 && p.gk[0] == 16 && p.gk[1] == 64 && p.gk[2] == 96) {
                        gauss5_fast_blur<TW, TH, RB, HXP, NAMED>(g_raw, u1_raw, bl_raw, tid, empty_bar);
                        done = true;
                    } else if (p.gk[ks >> 1] <= 255) {  // (the centre tap is the largest: all of them fit a byte)
                        done = true;
                        if constexpr (RB == kGaussRBSmall) {
                            if (ks == 3) gauss_fast_blur<TW, TH, RB, HXP, NAMED, 3>(p, g_raw, u1_raw, bl_raw, tid, empty_bar);
                            else if (ks == 5) gauss_fast_blur<TW, TH, RB, HXP, NAMED, 5>(p, g_raw, u1_raw, bl_raw, tid, empty_bar);
                            else if (ks == 7) gauss_fast_blur<TW, TH, RB, HXP, NAMED, 7>(p, g_raw, u1_raw, bl_raw, tid, empty_bar);
                            else done = false;
                        } else {
                            if (ks == 9) gauss_fast_blur<TW, TH, RB, HXP, NAMED, 9>(p, g_raw, u1_raw, bl_raw, tid, empty_bar);
                            else if (ks == 11) gauss_fast_blur<TW, TH, RB, HXP, NAMED, 11>(p, g_raw, u1_raw, bl_raw, tid, empty_bar);
                            else if (ks == 13) gauss_fast_blur<TW, TH, RB, HXP, NAMED, 13>(p, g_raw, u1_raw, bl_raw, tid, empty_bar);
                            else if (ks == 15) gauss_fast_blur<TW, TH, RB, HXP, NAMED, 15>(p, g_raw, u1_raw, bl_raw, tid, empty_bar);
                            else done = false;
                        }
                    }
                }
            }
            if (!done) gauss_blur_tile<TW, TH, RB, HXP, NAMED>(b, p, g_raw, u1_raw, bl_raw, f, x0, y0, tid, empty_bar);
            if (fast_thr) fast_threshold<128, 32, T::BW, NAMED>(f_raw, u1_raw, bl_raw, tid, cth);
        } else if (RB == 2 && fast_thr && !p.write_blur) {
            fast_blur_rb2<128, 32, HXP, NAMED>(g_raw, u1_raw, bl_raw, tid, empty_bar);
            fast_threshold<128, 32, T::BW, NAMED>(f_raw, u1_raw, bl_raw, tid, cth);
        } else {
            // ---- generic path: any border, any c, either comparison direction --------------------------------------
            // 2. blur over the tile + 5-px ring (zero outside the image, pass-through outside the interior)
            for (int idx = tid; idx < T::BH * (TW + 2 * kAdaptHalf); idx += 256) {
                const int r = idx / (TW + 2 * kAdaptHalf), i = idx - r * (TW + 2 * kAdaptHalf) + (8 - kAdaptHalf);
                const int gy = y0 - kAdaptHalf + r, gx = x0 - 8 + i;
                const int ry = r + RB;
                uint32_t v = 0;
                if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                    if (RB > 0 && gy >= RB && gy < H - RB && gx >= RB && gx < W - RB) {
                        uint32_t s = 0;
#pragma unroll
                        for (int dy = -RB; dy <= RB; dy++)
#pragma unroll
                            for (int dx = -RB; dx <= RB; dx++) s += s_g[ry + dy][i + dx + T::GOFF];
                        v = (RB == 2) ? div25(s) : s / ((2 * RB + 1) * (2 * RB + 1));
                    } else {
                        v = s_g[ry][i + T::GOFF];
                    }
                    if (p.write_blur && r >= kAdaptHalf && r < kAdaptHalf + TH && i >= 8 && i < 8 + TW)
                        b.blur[((size_t)f * H + gy) * W + gx] = (uint8_t)v;
                }
                s_bl[r][i] = (uint8_t)v;
            }
            tile_sync<NAMED>();
            if (empty_bar && tid == 0) mbar_arrive(empty_bar);  // last read of the gray stage is behind us
        }
        if (!((RB > 2 && p.gauss_ksize > 0) ? fast_thr : (RB == 2 && fast_thr && !p.write_blur))) {
            // 3. horizontal 11-sums: output column c uses s_bl columns c + 3 .. c + 13
            for (int idx = tid; idx < T::BH * TW; idx += 256) {
                const int r = idx / TW, c = idx - r * TW;
                uint32_t s = 0;
#pragma unroll
                for (int k = 0; k < 2 * kAdaptHalf + 1; k++) s += s_bl[r][c + (8 - kAdaptHalf) + k];
                s_h11[r][c] = (uint16_t)s;
            }
            tile_sync<NAMED>();
            // 4. vertical 11-sums + threshold test (window truncated at the image border: cnt = rows * cols)
            for (int idx = tid; idx < TH * TW; idx += 256) {
                const int r = idx / TW, c = idx - r * TW;
                const int gy = y0 + r, gx = x0 + c;
                uint8_t fg = 0;
                if (gy < H && gx < W) {
                    int s = 0;
#pragma unroll
                    for (int k = 0; k < 2 * kAdaptHalf + 1; k++) s += s_h11[r + k][c];
                    const int rows = min(gy + kAdaptHalf, H - 1) - max(gy - kAdaptHalf, 0) + 1;
                    const int cols = min(gx + kAdaptHalf, W - 1) - max(gx - kAdaptHalf, 0) + 1;
                    const int cnt = rows * cols;
                    const int px = s_bl[r + kAdaptHalf][c + 8];
                    bool t = p.inverse ? ((px + cth + 1) * cnt <= s) : !((px + cth) * cnt <= s);
                    // i32 wrap of `mean - c` (see PreprocessParams::wrap_t1): where mean >= T the right-hand side wrapped to
                    // a large negative number, elsewhere it is a large positive one
                    if (p.wrap_t1) t = (s < (p.wrap_t1 - 1) * cnt) == (p.inverse != 0);
                    fg = t ? 255 : 0;
                }
                s_f[r][c] = fg;
            }
            tile_sync<NAMED>();
        }
    }

    // ---- 5. outputs: u8 mask, bit-packed mask, label plane ------------------------------------------------------
    if (flat && x0 + TW <= W && y0 + TH <= H && (W & 15) == 0) {
        // flat tile fully inside a 16-px aligned image: nothing but wide zero stores (one warp writes one 512-byte
        // label row per instruction)
        const int4 z = make_int4(0, 0, 0, 0);
        const size_t o0 = (size_t)f * H * W + (size_t)y0 * W + x0;
        if (p.init_labels) {
            int4 *dst = reinterpret_cast<int4 *>(b.labels + o0 + (size_t)(tid >> 5) * W) + (tid & 31);
            const size_t step = (size_t)2 * W;  // 8 rows of W int32 = 2*W int4
#pragma unroll
            for (int k = 0; k < TH / 8; k++) dst[k * step] = z;
        }
        if (p.write_mask) {
            static_assert(TH * (TW / 16) == 256, "one 16-pixel group per thread");
            *reinterpret_cast<int4 *>(b.mask + o0 + (size_t)(tid >> 3) * W + 16 * (tid & 7)) = z;
        }
        if (tid < TH * (TW / 32)) {
            const int r = tid / (TW / 32), wq = tid - r * (TW / 32);
            bits_out[((size_t)f * H + y0 + r) * b.ww + (x0 >> 5) + wq] = 0u;
        }
        if (b.rowflags && tid < TH && y0 + tid < H)
            b.rowflags[(size_t)f * b.rf_stride + rowflag_index(y0 + tid, tile_x, b.tiles_x)] = 0;
        if (b.tile_occ && tid == 0)
            *reinterpret_cast<uint32_t *>(b.tile_occ + 4 * tile_occ_index(b, f, tile_x, y0 / TH)) = 0u;
        return;
    }
    if (x0 + TW <= W && (W & 15) == 0) {
        // full-width tile of a 16-px aligned image: one thread per 16 pixels, 128-bit stores throughout
        static_assert(TH * (TW / 16) == 256, "one 16-pixel group per thread");
        const int r = tid >> 3, c16 = (tid & 7) * 16;
        const int gy = y0 + r;
        if (gy < H) {
            const uint4 m16 = flat ? make_uint4(0u, 0u, 0u, 0u) : *reinterpret_cast<const uint4 *>(&s_f[r][c16]);
            const size_t o = (size_t)gy * W + x0 + c16;
            if (p.write_mask) *reinterpret_cast<uint4 *>(mask + o) = m16;
            if (p.init_labels) {
                int4 *dst = reinterpret_cast<int4 *>(labels + o);
                if (!(m16.x | m16.y | m16.z | m16.w)) {
                    const int4 z = make_int4(0, 0, 0, 0);
                    dst[0] = z, dst[1] = z, dst[2] = z, dst[3] = z;
                } else {
                    // 16 foreground bits, run starts inside the 32-px word, then p+1 at the starts
                    auto nib = [](uint32_t w) { return ((w & 0x01010101u) * 0x10204080u) >> 28; };
                    const uint32_t fg = nib(m16.x) | (nib(m16.y) << 4) | (nib(m16.z) << 8) | (nib(m16.w) << 12);
                    const uint32_t prev = (c16 & 31) ? (s_f[r][c16 - 1] & 1u) : 0u;
                    const uint32_t starts = fg & ~((fg << 1) | prev);
                    const int base = gy * W + x0 + c16 + 1;
                    int lab[16];
#pragma unroll
                    for (int k = 0; k < 16; k++) lab[k] = ((starts >> k) & 1u) ? base + k : 0;
#pragma unroll
                    for (int k = 0; k < 4; k++) dst[k] = make_int4(lab[4 * k], lab[4 * k + 1], lab[4 * k + 2], lab[4 * k + 3]);
                }
            }
        }
    } else
        for (int idx = tid; idx < TH * (TW / 4); idx += 256) {
            const int r = idx / (TW / 4), c4 = (idx - r * (TW / 4)) * 4;
        const int gy = y0 + r, gx = x0 + c4;
        if (gy >= H || gx >= W) continue;
        const uint32_t m4 = flat ? 0u : *reinterpret_cast<const uint32_t *>(&s_f[r][c4]);
        int4 lab = make_int4(0, 0, 0, 0);
        if (m4 && p.init_labels) {
            const uint32_t prev = (c4 & 31) ? s_f[r][c4 - 1] : 0u;
            const int base = gy * W + gx + 1;
            const uint32_t f0 = m4 & 0xffu, f1 = (m4 >> 8) & 0xffu, f2 = (m4 >> 16) & 0xffu, f3 = m4 >> 24;
            lab.x = (f0 && !prev) ? base : 0;
            lab.y = (f1 && !f0) ? base + 1 : 0;
            lab.z = (f2 && !f1) ? base + 2 : 0;
            lab.w = (f3 && !f2) ? base + 3 : 0;
        }
        if (vec) {
            if (p.write_mask) *reinterpret_cast<uint32_t *>(mask + (size_t)gy * W + gx) = m4;
            if (p.init_labels) *reinterpret_cast<int4 *>(labels + (size_t)gy * W + gx) = lab;
        } else {
            const int la[4] = {lab.x, lab.y, lab.z, lab.w};
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (gx + k < W) {
                    if (p.write_mask) mask[(size_t)gy * W + gx + k] = (uint8_t)(m4 >> (8 * k));
                    if (p.init_labels) labels[(size_t)gy * W + gx + k] = la[k];
                }
        }
    }
    static_assert(TW == 128 && TH * (TW / 32) <= 256 && (TH * (TW / 32)) % 32 == 0, "bit-packing stage layout");
    if (tid < TH * (TW / 32)) {  // whole warps: 8 rows x 4 words per warp
        const int r = tid >> 2, wq = tid & 3;
        const int gy = y0 + r, gwx = (x0 >> 5) + wq;
        const bool inside = gy < H && gwx < b.ww;
        uint32_t word = 0;
        if (!flat && inside) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(&s_f[r][wq * 32]);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                // gather bit 0 of each of the 4 mask bytes into a nibble
                const uint32_t nib = ((src[k] & 0x01010101u) * 0x10204080u) >> 28;
                word |= nib << (4 * k);
            }
        }
        if (inside) bits_out[((size_t)f * H + gy) * b.ww + gwx] = word;
        // occupancy nibble of this (row, tile): which of its 4 words are non-zero
        const uint32_t bal = __ballot_sync(0xffffffffu, word != 0);
        if (b.rowflags && wq == 0 && gy < H)
            b.rowflags[(size_t)f * b.rf_stride + rowflag_index(gy, tile_x, b.tiles_x)] =
                (uint8_t)((bal >> (tid & 31)) & 0xfu);
        // tile occupancy for the morphology scan: one byte per group of 8 rows (= this warp), bit k = word k non-zero
        if (b.tile_occ && (tid & 31) == 0) {
            uint32_t o = bal | (bal >> 16);
            o |= o >> 8;
            o |= o >> 4;
            b.tile_occ[4 * tile_occ_index(b, f, tile_x, y0 / TH) + (tid >> 5)] = (uint8_t)(o & 0xfu);
        }
    }
}

template <int TW, int TH, int RB>
__global__ void __launch_bounds__(256, 8) k_preprocess(BatchView b, PreprocessParams p, uint32_t *bits_out) {
    using T = Tile<TW, TH, RB>;
    static_assert(TW % 32 == 0, "tile width must cover whole bitmask words");
    __shared__ __align__(16) uint8_t smem[T::G_BYTES + T::U1_BYTES + T::BL_BYTES];
    uint8_t(*s_g)[T::GW] = reinterpret_cast<uint8_t(*)[T::GW]>(smem);

    const int tid = threadIdx.x;
    const int f = blockIdx.z;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int H = b.h, W = b.w;
    const uint8_t *gray = b.gray + (size_t)f * b.gray_frame_stride;
    const size_t gpitch = b.gray_row_stride;
    const bool aligned_in = ((reinterpret_cast<uintptr_t>(gray) | gpitch) & 3) == 0;

    // ---- 1. stage the gray tile + halo in shared memory and test it for flatness on the way ------------------------
    // Flatness (conservative, exact in effect): if every in-image pixel of the tile + halo satisfies
    // |g - ref| <= floor(c/2) for one reference pixel, the range is <= c and the mask of the tile is empty (see the
    // header).  VABSDIFF4 gives the four byte distances; (d & 0x7f) + (127 - T) sets bit 7 of a byte iff d > T, and
    // d itself has bit 7 set iff d >= 128, so OR-accumulating both and testing 0x80808080 decides the whole tile.
    const int cth = p.c_thresh;
    const bool try_flat = p.inverse && !p.write_blur && !p.force_generic && cth >= 0;
    uint32_t ref4 = 0;
    if (try_flat) ref4 = 0x01010101u * __ldg(gray + (size_t)min(y0 + TH / 2, H - 1) * gpitch + min(x0 + TW / 2, W - 1));
    const uint32_t kq = 0x01010101u * (uint32_t)(127 - min(cth >> 1, 127));
    uint32_t acc = 0;
    auto flat_test = [&](uint32_t v) {
        const uint32_t d = __vabsdiffu4(v, ref4);
        acc |= d | ((d & 0x7f7f7f7fu) + kq);
    };
    const bool full_in = x0 >= T::HX && x0 + TW + T::HX <= W && y0 >= T::HALO && y0 + TH + T::HALO <= H &&
                         ((reinterpret_cast<uintptr_t>(gray) | gpitch) & 7) == 0;
    if (full_in) {
        // all loads of the thread are issued before the first use: one DRAM latency per tile instead of one per word
        const uint8_t *base = gray + (size_t)(y0 - T::HALO) * gpitch + (x0 - T::HX);
        constexpr int NW = T::GH * (T::GW / 8);
        constexpr int PER = (NW + 255) / 256;
        uint2 v[PER];
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const int idx = tid + 256 * k;
            const int r = idx / (T::GW / 8), cw = idx - r * (T::GW / 8);
            v[k] = make_uint2(ref4, ref4);
            if (idx < NW) v[k] = __ldg(reinterpret_cast<const uint2 *>(base + (size_t)r * gpitch + 8 * cw));
        }
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const int idx = tid + 256 * k;
            const int r = idx / (T::GW / 8), cw = idx - r * (T::GW / 8);
            flat_test(v[k].x);
            flat_test(v[k].y);
            if (idx < NW) *reinterpret_cast<uint2 *>(&s_g[r][8 * cw]) = v[k];
        }
    } else {
        for (int idx = tid; idx < T::GH * (T::GW / 4); idx += 256) {
            const int r = idx / (T::GW / 4), cw = idx - r * (T::GW / 4);
            const int gy = y0 - T::HALO + r, gx = x0 - T::HX + 4 * cw;
            uint32_t v = 0;
            if (gy >= 0 && gy < H) {
                const uint8_t *row = gray + (size_t)gy * gpitch;
                if (aligned_in && gx >= 0 && gx + 3 < W) {
                    v = __ldg(reinterpret_cast<const uint32_t *>(row + gx));
                    flat_test(v);
                } else {
                    uint32_t vt = ref4;  // out-of-image bytes must not fail the test
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int x = gx + k;
                        if (x >= 0 && x < W) {
                            const uint32_t q = __ldg(row + x);
                            v |= q << (8 * k);
                            vt = (vt & ~(0xffu << (8 * k))) | (q << (8 * k));
                        }
                    }
                    flat_test(vt);
                }
            }
            *reinterpret_cast<uint32_t *>(&s_g[r][4 * cw]) = v;
        }
    }
    // barrier (s_g complete) + block-wide OR in one instruction
    const bool flat = !__syncthreads_or((int)(acc & 0x80808080u)) && try_flat;
    // s_f aliases s_g (dead by the time the mask bytes are written)
    tile_compute_and_store<TW, TH, RB, 8, false>(b, p, bits_out, smem, smem, smem + T::G_BYTES,
                                                 smem + T::G_BYTES + T::U1_BYTES, f, blockIdx.x, x0, y0, flat, tid, nullptr);
}

constexpr int kTmaStages = 2;  // stages per CTA.  4 stages and 4 CTAs per SM ran the headline batch in 49.2 us, 2 stages and 5 CTAs per
                               // SM (37.9 KB and 40 registers each) in 47.1 us: the kernel is latency-bound, residency wins
constexpr int kK1Consumers = 256;              // threads that test, compute and store tiles
constexpr int kK1Threads = kK1Consumers + 32;  // + one producer warp (scheduler + TMA issue)

// ---------------------------------------------------------------------------------------------------------------------
// K1 with the morphology folded in (A8 for small kernels): cv2 MORPH_OPEN then MORPH_CLOSE with 3x3 / 5x5 rectangles
// (heimdall/detectors/contamination_detector.py:81-87 runs 3x3 open + 3x3 close) cannot reach further than MR = 4
// pixels, so the tile's final mask is a function of the thresholded mask of the tile + 4 px, which is a function of the
// gray tile + 11 px: the staged box grows from 46 to 54 rows (its 16-column halo already covers the 11 columns), the
// mask is computed for the tile + ring in shared memory, bit-packed, eroded / dilated / eroded there, and only the final
// planes are written.  No pre-morphology plane, no tile scan, no second pass over a quarter of the frame: the kernels
// behind K1 are exactly those of the plain pipeline.
//
// Geometry of a tile (x0, y0), MR = morphology reach; "box column" c is image column x0 - 16 + c:
//   gray stage s_g  [GH = 32 + 2 (7 + MR)][160]          row g  = image row y0 - 7 - MR + g
//   s_v   u16 [BH = 32 + 2 (5 + MR)][VP]  vertical 5-sums  row r  = image row y0 - 5 - MR + r, box column c at index c + 4
//   s_bl  u8  [BH][BP]                    blur             same rows, box column c at index c
//   s_h11 u16 [BH][HP]  (aliases s_v)     horizontal 11-sums of the blur, column j = box column 8 + j (image x0 - 8 + j)
//   s_f2  u8  [MH = 32 + 2 MR][FP]        pre-morphology mask bytes, row m = image row y0 - MR + m, column j as s_h11
//   s_ma / s_mb u32 [MH][8]               bit-packed mask, word k (at column k + 1) = image columns x0 + 32 (k - 1) ..+31
// ---------------------------------------------------------------------------------------------------------------------
template <int MRv>
struct MTile {
    static constexpr int MR = MRv;
    static constexpr int HX = 16;
    static constexpr int HALO = 2 + kAdaptHalf + MR;
    static constexpr int GW = 160;
    static constexpr int GH = 32 + 2 * HALO;
    static constexpr int BH = 32 + 2 * (kAdaptHalf + MR);
    static constexpr int MH = 32 + 2 * MR;
    static constexpr int VP = 168, BP = 168, HP = 144, FP = 144;
    static constexpr int G_BYTES = GH * GW;
    static constexpr int U1_BYTES = BH * VP * 2;
    static constexpr int BL_BYTES = BH * BP;
    static constexpr int F_BYTES = MH * FP;
    static constexpr int M_BYTES = MH * 8 * 4;  // rows padded to 8 words: word k at column k + 1
    static constexpr int SCRATCH = U1_BYTES + BL_BYTES + F_BYTES + 2 * M_BYTES + 32 * 4 * 4;
    static_assert(MR == 4, "thread mappings below: 6 x 9 gray rows, 7 x 6 mask rows");
    static_assert(BH * HP * 2 <= U1_BYTES && (U1_BYTES % 16) == 0 && (BL_BYTES % 16) == 0 && (F_BYTES % 16) == 0, "layout");
};

// pre-morphology mask bytes of the tile + ring, interior tiles (the whole staged box lies inside the image): the packed
// u16x2 pipeline of fast_blur_rb2 / fast_threshold over the larger region
template <int MR>
__device__ __forceinline__ void morph_fast_mask(uint8_t *g_raw, uint8_t *u1_raw, uint8_t *bl_raw, uint8_t *f_raw, int tid, int cth,
                                                uint64_t *empty_bar) {
    using T = MTile<MR>;
    uint8_t(*s_g)[T::GW] = reinterpret_cast<uint8_t(*)[T::GW]>(g_raw);
    uint16_t(*s_v)[T::VP] = reinterpret_cast<uint16_t(*)[T::VP]>(u1_raw);
    uint16_t(*s_h11)[T::HP] = reinterpret_cast<uint16_t(*)[T::HP]>(u1_raw);
    uint8_t(*s_bl)[T::BP] = reinterpret_cast<uint8_t(*)[T::BP]>(bl_raw);
    uint8_t(*s_f2)[T::FP] = reinterpret_cast<uint8_t(*)[T::FP]>(f_raw);
    // A. vertical 5-sums: thread = (column quad q of 40, segment of 9 rows) -> 240 threads; blur row r sums gray rows r..r+4
    if (tid < 40 * 6) {
        const int q = tid % 40, seg = tid / 40;
        const int r0 = seg * 9;
        const int nrows = min(9, T::BH - r0);
        uint32_t lo[5], hi[5];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t v = *reinterpret_cast<const uint32_t *>(&s_g[r0 + k][4 * q]);
            lo[k] = prmt(v, 0, 0x4140);
            hi[k] = prmt(v, 0, 0x4342);
        }
        uint32_t alo = lo[0] + lo[1] + lo[2] + lo[3], ahi = hi[0] + hi[1] + hi[2] + hi[3];
#pragma unroll
        for (int k = 0; k < 9; k++) {
            if (k < nrows) {
                const uint32_t v = *reinterpret_cast<const uint32_t *>(&s_g[r0 + k + 4][4 * q]);
                const uint32_t nlo = prmt(v, 0, 0x4140), nhi = prmt(v, 0, 0x4342);
                alo += nlo;
                ahi += nhi;
                *reinterpret_cast<uint2 *>(&s_v[r0 + k][4 * q + 4]) = make_uint2(alo, ahi);
                alo -= lo[k % 5];
                ahi -= hi[k % 5];
                lo[(k + 4) % 5] = nlo;
                hi[(k + 4) % 5] = nhi;
            }
        }
    }
    tile_sync<true>();
    if (tid == 0) mbar_arrive(empty_bar);  // the gray stage is dead: the producer may refill it
    // B. horizontal 5-sums + /25: thread = (row, group of 32 box columns) -> 50 x 5 = 250 threads.  Box column c sums
    //    s_v indices c + 2 .. c + 6; a group loads indices 32g .. 32g + 39 (columns 0, 1, 158, 159 come out as garbage
    //    and are never used).
    if (tid < T::BH * 5) {
        const int g = tid % 5, r = tid / 5;
        const uint16_t *vrow = &s_v[r][32 * g];
        uint32_t pk[20];
#pragma unroll
        for (int k = 0; k < 10; k++) {
            const uint2 t = *reinterpret_cast<const uint2 *>(vrow + 4 * k);
            pk[2 * k] = t.x;
            pk[2 * k + 1] = t.y;
        }
        auto el = [&](int e) -> uint32_t { return (e & 1) ? (pk[e >> 1] >> 16) : (pk[e >> 1] & 0xffffu); };
        uint32_t s = el(2) + el(3) + el(4) + el(5);
        uint32_t *brow = reinterpret_cast<uint32_t *>(&s_bl[r][32 * g]);
#pragma unroll
        for (int j4 = 0; j4 < 8; j4++) {
            uint32_t q[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int j = 4 * j4 + u;
                s += el(j + 6);
                q[u] = __umulhi(s, 5243u << 15);  // floor(s / 25)
                s -= el(j + 2);
            }
            brow[j4] = q[0] | (q[1] << 8) | (q[2] << 16) | (q[3] << 24);
        }
    }
    tile_sync<true>();
    // C. horizontal 11-sums of the blur: thread = (row, group of 36 output columns) -> 50 x 4 = 200 threads.  Output column
    //    j (box column 8 + j) sums box columns 3 + j .. 13 + j; a group loads the 13 words from box column 36g.
    if (tid < T::BH * 4) {
        const int g = tid & 3, r = tid >> 2;
        const uint32_t *brow = reinterpret_cast<const uint32_t *>(&s_bl[r][36 * g]);
        uint32_t wd[13];
#pragma unroll
        for (int k = 0; k < 13; k++) wd[k] = brow[k];
        auto by = [&](int e) -> uint32_t { return (wd[e >> 2] >> (8 * (e & 3))) & 0xffu; };
        uint32_t s = 0;
#pragma unroll
        for (int e = 3; e < 13; e++) s += by(e);
        uint32_t *hrow = reinterpret_cast<uint32_t *>(&s_h11[r][36 * g]);
#pragma unroll
        for (int c2 = 0; c2 < 18; c2++) {
            s += by(2 * c2 + 13);
            const uint32_t a = s;
            s -= by(2 * c2 + 3);
            s += by(2 * c2 + 14);
            hrow[c2] = a | (s << 16);
            s -= by(2 * c2 + 4);
        }
    }
    tile_sync<true>();
    // D. vertical 11-sums + threshold test: thread = (column quad of 36, segment of 6 mask rows) -> 252 threads.  Mask row m
    //    sums s_h11 rows m .. m + 10, its own pixel is blur row m + 5.
    if (tid < 36 * 7) {
        const int q = tid % 36, seg = tid / 36;
        const int r0 = seg * 6;
        const int nrows = min(6, T::MH - r0);
        const uint32_t K2 = (uint32_t)((cth + 1) * 121) * 0x00010001u;
        uint32_t rl[11], rh[11];
        uint32_t alo = 0x00010001u, ahi = 0x00010001u;  // the "+ 1"
#pragma unroll
        for (int k = 0; k < 10; k++) {
            const uint2 t = *reinterpret_cast<const uint2 *>(&s_h11[r0 + k][4 * q]);
            rl[k] = t.x;
            rh[k] = t.y;
            alo += t.x;
            ahi += t.y;
        }
#pragma unroll
        for (int k = 0; k < 6; k++) {
            if (k < nrows) {
                const uint2 t = *reinterpret_cast<const uint2 *>(&s_h11[r0 + k + 10][4 * q]);
                alo += t.x;
                ahi += t.y;
                const uint32_t px4 = *reinterpret_cast<const uint32_t *>(&s_bl[r0 + k + 5][4 * q + 8]);
                const uint32_t xlo = prmt(px4, 0, 0x4140) * 121u + K2, xhi = prmt(px4, 0, 0x4342) * 121u + K2;
                const uint32_t dlo = __vminu2(__vmaxu2(alo, xlo) - xlo, 0x00010001u);
                const uint32_t dhi = __vminu2(__vmaxu2(ahi, xhi) - xhi, 0x00010001u);
                *reinterpret_cast<uint32_t *>(&s_f2[r0 + k][4 * q]) = prmt(dlo, dhi, 0x6420) * 255u;
                alo -= rl[k % 11];
                ahi -= rh[k % 11];
                rl[(k + 10) % 11] = t.x;
                rh[(k + 10) % 11] = t.y;
            }
        }
    }
    tile_sync<true>();
}

// the same for tiles whose box is clipped by the image: per-pixel loops with the border rules (blur passes through outside
// the interior, the 11x11 window is truncated and divides by its own count), zero outside the image
template <int MR>
__device__ __forceinline__ void morph_generic_mask(const BatchView &b, uint8_t *g_raw, uint8_t *u1_raw, uint8_t *bl_raw,
                                                   uint8_t *f_raw, int x0, int y0, int tid, int cth, uint64_t *empty_bar) {
    using T = MTile<MR>;
    uint8_t(*s_g)[T::GW] = reinterpret_cast<uint8_t(*)[T::GW]>(g_raw);
    uint16_t(*s_h11)[T::HP] = reinterpret_cast<uint16_t(*)[T::HP]>(u1_raw);
    uint8_t(*s_bl)[T::BP] = reinterpret_cast<uint8_t(*)[T::BP]>(bl_raw);
    uint8_t(*s_f2)[T::FP] = reinterpret_cast<uint8_t(*)[T::FP]>(f_raw);
    const int H = b.h, W = b.w;
    // blur over box columns 2..157 of the BH ring rows (zero outside the image: TMA zero fill + explicit test)
    for (int idx = tid; idx < T::BH * 156; idx += kK1Consumers) {
        const int r = idx / 156, c = idx - r * 156 + 2;
        const int gy = y0 - kAdaptHalf - MR + r, gx = x0 - T::HX + c;
        uint32_t v = 0;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
            if (gy >= 2 && gy < H - 2 && gx >= 2 && gx < W - 2) {
                uint32_t s = 0;
#pragma unroll
                for (int dy = 0; dy < 5; dy++)
#pragma unroll
                    for (int dx = -2; dx <= 2; dx++) s += s_g[r + dy][c + dx];
                v = div25(s);
            } else {
                v = s_g[r + 2][c];
            }
        }
        s_bl[r][c] = (uint8_t)v;
    }
    tile_sync<true>();
    if (tid == 0) mbar_arrive(empty_bar);
    for (int idx = tid; idx < T::BH * T::HP; idx += kK1Consumers) {
        const int r = idx / T::HP, j = idx - r * T::HP;
        uint32_t s = 0;
#pragma unroll
        for (int k = 0; k < 2 * kAdaptHalf + 1; k++) s += s_bl[r][j + 3 + k];
        s_h11[r][j] = (uint16_t)s;
    }
    tile_sync<true>();
    for (int idx = tid; idx < T::MH * T::FP; idx += kK1Consumers) {
        const int m = idx / T::FP, j = idx - m * T::FP;
        const int gy = y0 - MR + m, gx = x0 - 8 + j;
        uint8_t fg = 0;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
            int s = 0;
#pragma unroll
            for (int k = 0; k < 2 * kAdaptHalf + 1; k++) s += s_h11[m + k][j];
            const int rows = min(gy + kAdaptHalf, H - 1) - max(gy - kAdaptHalf, 0) + 1;
            const int cols = min(gx + kAdaptHalf, W - 1) - max(gx - kAdaptHalf, 0) + 1;
            const int px = s_bl[m + kAdaptHalf][j + 8];
            fg = ((px + cth + 1) * rows * cols <= s) ? 255 : 0;
        }
        s_f2[m][j] = fg;
    }
    tile_sync<true>();
}

// One erode / dilate pass with a (2r+1) x (2r+1) rectangle over the bit-packed region, one thread per word.  OpenCV's
// default border ("outside the image never wins") is applied when a word is READ: bits outside the image count as 1 for
// erode and 0 for dilate.  Words and rows beyond the staged region read as the identity as well (what they would really
// hold is unknown, but nothing that far away can reach the tile).
template <bool DILATE, int R>
__device__ __forceinline__ uint32_t morph_rect_word(const uint32_t (*src)[8], int m, int k, int m_lo, int m_hi, const uint32_t (&inm)[3]) {
    // src rows are padded: word k of the region sits at column k + 1, columns 0 and 7 are never written and masked out by
    // inm[0] / inm[2] = 0 for k = 0 / k = 5
    uint32_t acc = DILATE ? 0u : 0xffffffffu;
    const int a_lo = max(m - R, m_lo), a_hi = min(m + R, m_hi);
    for (int mm = a_lo; mm <= a_hi; mm++) {
        uint32_t L = src[mm][k], M = src[mm][k + 1], Rw = src[mm][k + 2];
        L = DILATE ? (L & inm[0]) : (L | ~inm[0]);
        M = DILATE ? (M & inm[1]) : (M | ~inm[1]);
        Rw = DILATE ? (Rw & inm[2]) : (Rw | ~inm[2]);
        uint32_t h = M;
#pragma unroll
        for (int dx = 1; dx <= R; dx++) {
            const uint32_t tr = __funnelshift_r(M, Rw, dx), tl = __funnelshift_l(L, M, dx);
            h = DILATE ? (h | tr | tl) : (h & tr & tl);
        }
        acc = DILATE ? (acc | h) : (acc & h);
    }
    return acc;
}
template <bool DILATE>
__device__ __forceinline__ uint32_t morph_rect_word_r(const uint32_t (*src)[8], int m, int k, int r, int m_lo, int m_hi,
                                                      const uint32_t (&inm)[3]) {
    switch (r) {  // block-uniform
        case 1: return morph_rect_word<DILATE, 1>(src, m, k, m_lo, m_hi, inm);
        case 2: return morph_rect_word<DILATE, 2>(src, m, k, m_lo, m_hi, inm);
        case 3: return morph_rect_word<DILATE, 3>(src, m, k, m_lo, m_hi, inm);
        default: return morph_rect_word<DILATE, 4>(src, m, k, m_lo, m_hi, inm);
    }
}

// Everything after the flat decision for the morphology variant: pre-morphology mask of the tile + ring, bit-packing,
// erode(ro) -> dilate(ro + rc) -> erode(rc) (open = erode, dilate; close = dilate, erode; the two dilations in the middle are
// one), final planes.  Block-uniform control flow; contains barriers.
template <int MR>
__device__ __forceinline__ void morph_tile_compute_and_store(const BatchView &b, const PreprocessParams &p, uint32_t *bits_out,
                                                             uint8_t *g_raw, uint8_t *scratch, int f, int tx, int ty, bool flat,
                                                             bool box_inside, int tid, uint64_t *empty_bar) {
    using T = MTile<MR>;
    uint8_t *u1_raw = scratch, *bl_raw = u1_raw + T::U1_BYTES, *f_raw = bl_raw + T::BL_BYTES;
    uint32_t(*s_ma)[8] = reinterpret_cast<uint32_t(*)[8]>(f_raw + T::F_BYTES);
    uint32_t(*s_mb)[8] = reinterpret_cast<uint32_t(*)[8]>(f_raw + T::F_BYTES + T::M_BYTES);
    uint32_t(*s_w)[4] = reinterpret_cast<uint32_t(*)[4]>(f_raw + T::F_BYTES + 2 * T::M_BYTES);
    uint8_t(*s_f2)[T::FP] = reinterpret_cast<uint8_t(*)[T::FP]>(f_raw);
    const int H = b.h, W = b.w, WW = b.ww;
    const int x0 = tx * 128, y0 = ty * 32;
    bool any = false;
    if (flat) {
        if (tid == 0) mbar_arrive(empty_bar);  // every thread is past its flat-test reads of the stage
    } else {
        if (box_inside)
            morph_fast_mask<MR>(g_raw, u1_raw, bl_raw, f_raw, tid, p.c_thresh, empty_bar);
        else
            morph_generic_mask<MR>(b, g_raw, u1_raw, bl_raw, f_raw, x0, y0, tid, p.c_thresh, empty_bar);
        // bit-packing: thread = (mask row m, word k); word k holds image columns x0 + 32 (k - 1) .. + 31 = s_f2 columns
        // 32k - 24 .. 32k + 7, of which 0 .. 143 exist.  Bits outside the image are cleared.
        const int m = tid / 6, k = tid - 6 * m;
        const bool mine = tid < T::MH * 6;
        // in-image masks of the word and of its two neighbours, rows of the region inside the image: all ones / the whole
        // region when the box lies inside the image
        uint32_t inm[3] = {k > 0 ? 0xffffffffu : 0u, 0xffffffffu, k < 5 ? 0xffffffffu : 0u};
        int m_lo = 0, m_hi = T::MH - 1;
        if (!box_inside) {
            auto col_mask = [&](int kk) -> uint32_t {
                const int gx0 = x0 + 32 * (kk - 1);  // image column of bit 0 (tiles start at multiples of 128: a word lies
                if (kk < 0 || kk > 5 || gx0 < 0 || gx0 >= W) return 0u;  // inside or outside as a whole on the left)
                return gx0 + 32 > W ? (1u << (W - gx0)) - 1u : 0xffffffffu;
            };
            inm[0] = col_mask(k - 1), inm[1] = col_mask(k), inm[2] = col_mask(k + 1);
            m_lo = max(0, MR - y0), m_hi = min(T::MH - 1, MR + (H - 1 - y0));
        }
        const bool row_in = m >= m_lo && m <= m_hi;
        uint32_t word = 0;
        if (mine) {
            if (k >= 1 && k <= 4) {
                const uint32_t *src = reinterpret_cast<const uint32_t *>(&s_f2[m][32 * k - 24]);
#pragma unroll
                for (int i = 0; i < 8; i++) word |= (((src[i] & 0x01010101u) * 0x10204080u) >> 28) << (4 * i);
            } else if (k == 0) {
                const uint32_t *src = reinterpret_cast<const uint32_t *>(&s_f2[m][0]);
                word = ((((src[0] & 0x01010101u) * 0x10204080u) >> 28) << 24) | ((((src[1] & 0x01010101u) * 0x10204080u) >> 28) << 28);
            } else {
                const uint32_t *src = reinterpret_cast<const uint32_t *>(&s_f2[m][136]);
                word = (((src[0] & 0x01010101u) * 0x10204080u) >> 28) | ((((src[1] & 0x01010101u) * 0x10204080u) >> 28) << 4);
            }
            if (!row_in) word = 0;
            word &= inm[1];
            s_ma[m][k + 1] = word;
        }
        any = tile_sync_or<true>(word);  // also: s_ma complete
        if (any) {
            const int ro = p.morph_open_k > 0 ? (p.morph_open_k - 1) / 2 : 0, rc = p.morph_close_k > 0 ? (p.morph_close_k - 1) / 2 : 0;
            const uint32_t(*src)[8] = s_ma;
            uint32_t(*dst)[8] = s_mb;
            auto flip = [&]() {
                tile_sync<true>();
                const uint32_t(*t)[8] = dst;
                dst = const_cast<uint32_t(*)[8]>(src);
                src = t;
            };
            if (ro > 0) {
                if (mine) dst[m][k + 1] = row_in ? (morph_rect_word_r<false>(src, m, k, ro, m_lo, m_hi, inm) & inm[1]) : 0u;
                flip();
            }
            if (ro + rc > 0) {
                if (mine) dst[m][k + 1] = row_in ? (morph_rect_word_r<true>(src, m, k, ro + rc, m_lo, m_hi, inm) & inm[1]) : 0u;
                flip();
            }
            if (rc > 0) {
                if (mine) dst[m][k + 1] = row_in ? (morph_rect_word_r<false>(src, m, k, rc, m_lo, m_hi, inm) & inm[1]) : 0u;
                flip();
            }
            if (tid < 128) s_w[tid >> 2][tid & 3] = src[MR + (tid >> 2)][2 + (tid & 3)];
            tile_sync<true>();
        }
    }
    // final bit-mask words, occupancy records, then the mask bytes and the label plane
    uint32_t word = 0;
    if (tid < 128) {
        const int r = tid >> 2, wq = tid & 3;
        if (any) word = s_w[r][wq];
        const int gy = y0 + r, gwx = 4 * tx + wq;
        if (gy >= H || gwx >= WW) word = 0;
        if (gy < H && gwx < WW && !(p.sparse_aux && !any)) bits_out[((size_t)f * H + gy) * WW + gwx] = word;
        const uint32_t bal = __ballot_sync(0xffffffffu, word != 0);
        if (b.rowflags && wq == 0 && gy < H)
            b.rowflags[(size_t)f * b.rf_stride + rowflag_index(gy, tx, b.tiles_x)] = (uint8_t)((bal >> (tid & 31)) & 0xfu);
    }
    const bool any_out = tile_sync_or<true>(word);
    expand_tile(b, (size_t)f, tx, ty, s_w, any_out, tid);
    // s_w / s_ma / s_mb are rewritten by the next tile only behind its flat-test barrier, which every consumer reaches after
    // its reads here
}

// ---------------------------------------------------------------------------------------------------------------------
// K1 v3: persistent CTAs, TMA-staged tiles (cp.async.bulk.tensor.3d + mbarrier), double buffering, dynamic tile scheduler.
// The v2 kernel above is one CTA per tile: load -> barrier -> compute -> store, so every tile exposes a full DRAM latency
// and needs >= 6 resident CTAs per SM to hide it (ncu: long-scoreboard stalls dominate).  Here a CTA keeps fetching tiles
// from an atomic counter and the TMA unit loads tile k+1 (zero-filling outside the image) while the threads test, compute
// and store tile k; no thread ever issues a global load for pixels.
// ---------------------------------------------------------------------------------------------------------------------

template <int TW, int TH, int RB, int MR = 0, int NS = kTmaStages>
struct TmaSmem {
    using T = Tile<TW, TH, RB, 16>;  // 16-column halo: TMA boxes must start on a 16-byte boundary
    static constexpr int STAGE = (T::G_BYTES + 127) & ~127;
    static constexpr int BYTES = NS * STAGE + TH * TW + T::U1_BYTES + T::BL_BYTES;
};
template <int TW, int TH, int NS>
struct TmaSmem<TW, TH, 2, 4, NS> {  // box blur + 3x3 / 5x5 open + close folded in
    using T = MTile<4>;
    static constexpr int STAGE = (T::G_BYTES + 127) & ~127;
    static constexpr int BYTES = NS * STAGE + T::SCRATCH;
};


// Warp-specialised: warp 8 is the producer (one elected lane takes tile numbers from the scheduler and issues the TMA
// loads, running up to kTmaStages tiles ahead), warps 0..7 are the consumers.  In the first TMA version thread 0 did the
// refill between two tiles: the TMA issue sits behind that thread's own outstanding global stores, and every barrier of
// the next tile waited for it (0.4-0.7 us per tile of 1.3-5 us).  Stage hand-over: full[st] (TMA transaction barrier,
// producer -> consumers) and empty[st] (one consumer arrival after the last read of the stage, consumers -> producer).
template <int TW, int TH, int RB, int MR = 0, int NS = kTmaStages>
__global__ void __launch_bounds__(kK1Threads, (RB == 2 && MR == 0) ? 5 : 4)
    k_preprocess_tma(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ BatchView b,
                     const __grid_constant__ PreprocessParams p, uint32_t *bits_out, unsigned int *sched) {
    using T = typename TmaSmem<TW, TH, RB, MR, NS>::T;
    constexpr int STAGE = TmaSmem<TW, TH, RB, MR, NS>::STAGE;
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t full[NS], empty[NS];
    __shared__ int4 s_tile[NS];  // {tile number, frame, tile x, tile y}, decoded once by the producer
    uint8_t *f_raw = sm + NS * STAGE, *u1_raw = f_raw + TH * TW, *bl_raw = u1_raw + T::U1_BYTES;
    (void)u1_raw, (void)bl_raw;

    const int tid = threadIdx.x;
    const int H = b.h, W = b.w;
    const int tiles_x = b.tiles_x, tiles_y = (H + TH - 1) / TH;
    const int per_frame = tiles_x * tiles_y, total = per_frame * b.n;
    unsigned long long t_cta0 = 0;
    if (b.phase_ns && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_cta0));

    // the per-frame CCL kernel of this batch may be launched now (programmatic dependent launch): its CTAs become
    // resident as ours retire and wait there for this grid to complete, which takes its launch latency off the step
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) {
        // Launched while the kernels of the previous batch are still running (programmatic dependent launch): the slot this
        // batch writes was last used by the per-frame CCL kernel two batches back, which is normally long done -- make sure
        // (its completion counter; wrap-safe comparison) before anything of the slot, the tile counter included, is touched.
        if (b.ccl_done) {
            const volatile unsigned int *flag = b.ccl_done;
            while ((int)(*flag - b.ccl_wait_value) < 0) __nanosleep(100);
            for (int k = 0; k < b.ccl_wait_n; k++) {
                const volatile unsigned int *flag2 = b.ccl_wait_flag[k];
                while ((int)(*flag2 - b.ccl_wait_val[k]) < 0) __nanosleep(100);
            }
            __threadfence();
        }
#pragma unroll
        for (int k = 0; k < NS; k++) {
            mbar_init(&full[k], 1);
            mbar_init(&empty[k], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= kK1Consumers) {
        // ---- producer warp ---------------------------------------------------------------------------------------------
        if (tid != kK1Consumers) return;
        // Tiles claimed ahead are tiles no other CTA can take.  A non-flat tile costs four times a flat one, so with every
        // CTA holding NS claims (a third of the batch over the whole grid) the CTAs used to finish up to 17 us
        // apart.  The producer therefore runs `look` tiles ahead of the consumers -- p.lookahead in the steady state,
        // p.tail_lookahead once the tile numbers handed out are within p.tail_tiles of the end -- and claims a tile only
        // when it is about to issue its load.  Tile it may be issued once tile it - look has been released; the stage's
        // own barrier (tile it - NS) is implied because the consumers release in order.
        // p.claim_ahead (off): request the next tile number one tile early, so that the round trip of the atomic overlaps
        // the wait below.  Measured: the TMA wait per tile drops from 0.71 to 0.52 us, but one more tile per CTA is spoken
        // for and the step gets 0.9 us longer (49.7 -> 50.6 us): balance is worth more than the hidden latency.
        const int tail_from = total - p.tail_tiles;
        int look = p.lookahead, prev_t = 0;
        int pending = 0;
        if (p.claim_ahead) pending = p.static_sched ? (int)blockIdx.x : (int)atomicAdd(sched, 1u);
        for (int it = 0;; it++) {
            const int st = it % NS;
            if (prev_t >= tail_from) look = p.tail_lookahead;
            if (it >= look) mbar_wait(&empty[(it - look) % NS], (uint32_t)((it - look) / NS) & 1u, (uint32_t)p.wait_hint_ns);
            int t;
            if (p.claim_ahead) {
                t = pending;
                if (t < total) pending = p.static_sched ? t + (int)gridDim.x : (int)atomicAdd(sched, 1u);
            } else {
                t = p.static_sched ? (int)blockIdx.x + it * (int)gridDim.x : (int)atomicAdd(sched, 1u);
            }
            prev_t = t;
            const int f = t / per_frame, r = t - f * per_frame;
            const int ty = r / tiles_x, tx = r - ty * tiles_x;
            // bit 16 of .w: the whole staged box lies inside the image; bit 17: the whole tile lies inside the image
            const int x0p = tx * TW, y0p = ty * TH;
            const int fl = ((x0p >= T::HX && x0p + TW + T::HX <= W && y0p >= T::HALO && y0p + TH + T::HALO <= H) ? 0x10000 : 0) |
                           ((x0p + TW <= W && y0p + TH <= H) ? 0x20000 : 0);
            s_tile[st] = make_int4(t, f, tx, ty | fl);  // published by the arrive below (release) / the consumers' wait (acquire)
            if (t >= total) {
                mbar_arrive(&full[st]);  // end marker: tile numbers only grow, nothing is left for this CTA
                break;
            }
            mbar_expect_tx(&full[st], (uint32_t)T::G_BYTES);
            tma_load_3d(sm + st * STAGE, &tmap, &full[st], tx * TW - T::HX, ty * TH - T::HALO, f);
            // Tiles are handed out in order, so tile t + prefetch_tiles will be claimed by some CTA a few microseconds
            // from now: pull its box into the L2 (whoever loads it then finds it there; every tile is prefetched once).
            if (p.prefetch_tiles > 0 && t + p.prefetch_tiles < total) {
                const int tp = t + p.prefetch_tiles;
                const int fp = tp / per_frame, rp = tp - fp * per_frame;
                const int typ = rp / tiles_x, txp = rp - typ * tiles_x;
                tma_prefetch_3d(&tmap, txp * TW - T::HX, typ * TH - T::HALO, fp);
            }
        }
        // the last CTA to finish fetching rearms the scheduler for the next launch
        if (!p.static_sched) {
            __threadfence();
            const unsigned int d = atomicAdd(sched + 1, 1u);
            if (d == gridDim.x - 1) {
                sched[0] = 0;
                sched[1] = 0;
            }
        }
        return;
    }

    // ---- consumers -----------------------------------------------------------------------------------------------------
    // Most tiles are flat and do nothing but test 7 KB of shared memory and store 20 KB of zeros, so the instruction count
    // of that path is what the SMs' issue slots see (ncu, uniform frames: 2400 warp instructions per tile, IPC 2.0, issue
    // slots 50 % busy before this was hoisted).  Everything that depends only on the thread is computed here, once.
    const int cth = p.c_thresh;
    const bool try_flat = p.inverse && !p.write_blur && !p.force_generic && cth >= 0;
    const uint32_t kq = 0x01010101u * (uint32_t)(127 - min(cth >> 1, 127));
    // flat test of a box that lies entirely inside the image: columns [8, 152) of all GH rows as
    //   - 16-byte items over columns [16, 144): 8 per row -> rows 0..31 one per thread, rows 32..45 threads 0..111
    //   - 8-byte edge items (columns 8..15 and 144..151): 2 per row -> threads 128..219
    static_assert(TW == 128 && TH == 32 && T::GW == 160 && (MR > 0 || RB != 2 || T::GH == 46), "flat-test thread mapping");
    const uint32_t ft0 = (uint32_t)((tid >> 3) * T::GW + 16 + (tid & 7) * 16);
    const uint32_t ft1 = (uint32_t)((32 + (tid >> 3)) * T::GW + 16 + (tid & 7) * 16);  // tid < 112
    const uint32_t fte = (uint32_t)((((tid - 128) >> 1)) * T::GW + (((tid - 128) & 1) ? 144 : 8));  // 128 <= tid < 220
    const uint32_t ftref = (uint32_t)((T::HALO + TH / 2) * T::GW + T::HX + TW / 2);
    const int mj0 = tid % 10, mj1 = (tid + 256) % 10, mj2 = (tid + 512) % 10;  // morphology variant: item column of the thread's items
    (void)mj0, (void)mj1, (void)mj2;
    // zero stores of a flat tile that lies entirely inside the image (element offsets from the tile's first pixel / word)
    const uint32_t so_lab = (uint32_t)((tid >> 5) * W + 4 * (tid & 31)), so_lab_step = (uint32_t)(8 * W);
    const uint32_t so_mask = (uint32_t)((tid >> 3) * W + 16 * (tid & 7));
    const uint32_t so_bits = (uint32_t)((tid >> 2) * b.ww + (tid & 3));
    const bool fast_store_ok = (W & 15) == 0;
    auto absd = [&](uint32_t v, uint32_t ref4, uint32_t &acc) {
        const uint32_t d = __vabsdiffu4(v, ref4);
        acc |= d | ((d & 0x7f7f7f7fu) + kq);
    };
    for (int it = 0;; it++) {
        const int st = it % NS;
        unsigned long long t_a = 0, t_b = 0;
        if (b.phase_ns && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_a));
        mbar_wait(&full[st], (uint32_t)(it / NS) & 1u, (uint32_t)p.wait_hint_ns);
        if (b.phase_ns && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_b));
        const int4 cur = s_tile[st];
        if (cur.x >= total) break;
        const int f = cur.y, tx = cur.z, ty = cur.w & 0xffff;
        const bool box_inside = (cur.w & 0x10000) != 0, tile_inside = (cur.w & 0x20000) != 0;
        const int x0 = tx * TW, y0 = ty * TH;

        // flatness test straight from shared memory, over the in-image part of the tile + halo
        uint8_t *cur_stage = sm + st * STAGE;
        uint32_t acc = 0;
        if constexpr (MR > 0) {
            // morphology variant: every 16-byte item of the box that lies inside the image (the image is 16-px aligned, so an
            // item is inside or outside as a whole).  The mask of the tile + MR ring depends on columns [5, 155) of the box:
            // the outer four columns on either side are ignored.
            if (try_flat && box_inside) {
                // the box is one contiguous run of GH * 10 sixteen-byte items: items tid, tid + 256, tid + 512; which of them
                // is a row's first / last item (outer four columns ignored) is known per thread (mj0..2, hoisted)
                constexpr int NI = T::GH * (T::GW / 16);
                static_assert(NI > 512 && NI <= 768, "three items per thread");
                const uint32_t ref4 = 0x01010101u * cur_stage[ftref];
                uint4 v0 = *reinterpret_cast<const uint4 *>(cur_stage + 16 * tid);
                uint4 v1 = *reinterpret_cast<const uint4 *>(cur_stage + 16 * (tid + 256));
                if (mj0 == 0) v0.x = ref4;
                if (mj0 == 9) v0.w = ref4;
                if (mj1 == 0) v1.x = ref4;
                if (mj1 == 9) v1.w = ref4;
                absd(v0.x, ref4, acc), absd(v0.y, ref4, acc), absd(v0.z, ref4, acc), absd(v0.w, ref4, acc);
                absd(v1.x, ref4, acc), absd(v1.y, ref4, acc), absd(v1.z, ref4, acc), absd(v1.w, ref4, acc);
                if (tid + 512 < NI) {  // (the last NI - 512 items: one warp)
                    uint4 v2 = *reinterpret_cast<const uint4 *>(cur_stage + 16 * (tid + 512));
                    if (mj2 == 0) v2.x = ref4;
                    if (mj2 == 9) v2.w = ref4;
                    absd(v2.x, ref4, acc), absd(v2.y, ref4, acc), absd(v2.z, ref4, acc), absd(v2.w, ref4, acc);
                }
            } else if (try_flat) {
                constexpr int IPR = T::GW / 16;
                const int r_lo = max(0, T::HALO - y0), r_hi = min(T::GH, H - y0 + T::HALO);
                const int j_lo = x0 == 0 ? 1 : 0, j_hi = min(IPR, (W - x0 + T::HX) >> 4);
                const uint32_t ref4 = 0x01010101u * cur_stage[min(T::HALO + TH / 2, r_hi - 1) * T::GW + min(T::HX + TW / 2, 16 * j_hi - 1)];
                for (int idx = tid; idx < T::GH * IPR; idx += kK1Consumers) {
                    const int r = idx / IPR, j = idx - r * IPR;
                    if (r >= r_lo && r < r_hi && j >= j_lo && j < j_hi) {
                        uint4 v = *reinterpret_cast<const uint4 *>(cur_stage + r * T::GW + 16 * j);
                        if (j == 0) v.x = ref4;
                        if (j == IPR - 1) v.w = ref4;
                        absd(v.x, ref4, acc), absd(v.y, ref4, acc), absd(v.z, ref4, acc), absd(v.w, ref4, acc);
                    }
                }
            }
        } else if (RB != 2) {
            // Gaussian variant.  BORDER_REFLECT_101: cells of the box that lie outside the image (TMA zero fill) but within
            // reach of the filters (RB + 5 <= 12 px) take the value of their mirror image, which is inside the box.
            uint8_t(*s_g)[T::GW] = reinterpret_cast<uint8_t(*)[T::GW]>(cur_stage);
            if (!box_inside) {
                // rows above / below the image first (whole rows, 16-byte items, from their mirror rows), then the columns
                // left / right of it in every row (from their mirror columns, which lie inside the image).  The whole halo
                // is filled, 16 columns and 12 rows, so that a border tile can pass the flat test like any other.
                // (below the image only the HALO rows the filters can reach: the mirror of anything further lies outside
                // the box; such rows exist only in a frame's last, partial tile row)
                const int r_top = max(0, T::HALO - y0), r_bot = min(T::GH, H - y0 + T::HALO);
                const int n_rows = r_top + min(T::GH - r_bot, T::HALO);
                constexpr int IPR = T::GW / 16;
                for (int idx = tid; idx < n_rows * IPR; idx += kK1Consumers) {
                    const int k = idx / IPR, j = idx - k * IPR;
                    const int r = k < r_top ? k : r_bot + (k - r_top);
                    const int gy = y0 - T::HALO + r;
                    const int sy = gy < 0 ? -gy : 2 * (H - 1) - gy;
                    *reinterpret_cast<uint4 *>(&s_g[r][16 * j]) = *reinterpret_cast<const uint4 *>(&s_g[sy - (y0 - T::HALO)][16 * j]);
                }
                if (n_rows) tile_sync<true>();  // block-uniform
                const int c_r = W - x0 + T::HX;  // first box column right of the image
                const int n_left = x0 == 0 ? T::HX : 0, n_right = c_r < T::GW ? min(T::HX, T::GW - c_r) : 0;
                const int n_cols = n_left + n_right;
                for (int idx = tid; idx < T::GH * n_cols; idx += kK1Consumers) {
                    const int r = idx / n_cols, k = idx - r * n_cols;
                    const int c = k < n_left ? k : c_r + (k - n_left);
                    const int gx = x0 - T::HX + c;
                    const int sx = gx < 0 ? -gx : 2 * (W - 1) - gx;
                    s_g[r][c] = s_g[r][sx - (x0 - T::HX)];
                }
                if (n_cols) tile_sync<true>();
            }
            if (try_flat) {  // the rows within reach of the filters (blur radius + 5 of the halo rows)
                const uint32_t ref4 = 0x01010101u * s_g[T::HALO + TH / 2][T::HX + TW / 2];
                const int r_skip = RB - (p.gauss_ksize >> 1);
                const int nrows = T::GH - 2 * r_skip;
                const uint8_t *base = cur_stage + r_skip * T::GW;
                if constexpr (RB == kGaussRBSmall) {
                    // reach <= 8 columns: columns [8, 152), the mapping of the box-blur variant (8 sixteen-byte items per row,
                    // rows 0..31 one per thread, the rest threads 0..; the two 8-byte edge items of a row: threads 128..)
                    const uint4 v0 = *reinterpret_cast<const uint4 *>(base + ft0);
                    absd(v0.x, ref4, acc), absd(v0.y, ref4, acc), absd(v0.z, ref4, acc), absd(v0.w, ref4, acc);
                    if (tid < (nrows - 32) * 8) {
                        const uint4 v1 = *reinterpret_cast<const uint4 *>(base + ft1);
                        absd(v1.x, ref4, acc), absd(v1.y, ref4, acc), absd(v1.z, ref4, acc), absd(v1.w, ref4, acc);
                    } else if (tid >= 128 && tid < 128 + 2 * nrows) {
                        const uint2 v1 = *reinterpret_cast<const uint2 *>(base + fte);
                        absd(v1.x, ref4, acc), absd(v1.y, ref4, acc);
                    }
                } else {  // all 16-byte items of those rows (4 px more than needed on either side): one contiguous run
                    const int nv = nrows * (T::GW / 16);  // <= 560
                    const uint4 v0 = *reinterpret_cast<const uint4 *>(base + 16 * tid);  // (nv >= 440)
                    absd(v0.x, ref4, acc), absd(v0.y, ref4, acc), absd(v0.z, ref4, acc), absd(v0.w, ref4, acc);
                    if (tid + 256 < nv) {
                        const uint4 v1 = *reinterpret_cast<const uint4 *>(base + 16 * (tid + 256));
                        absd(v1.x, ref4, acc), absd(v1.y, ref4, acc), absd(v1.z, ref4, acc), absd(v1.w, ref4, acc);
                    }
                    if (tid + 512 < nv) {
                        const uint4 v2 = *reinterpret_cast<const uint4 *>(base + 16 * (tid + 512));
                        absd(v2.x, ref4, acc), absd(v2.y, ref4, acc), absd(v2.z, ref4, acc), absd(v2.w, ref4, acc);
                    }
                }
            }
        } else if (try_flat) {
            if (box_inside) {
                const uint32_t ref4 = 0x01010101u * cur_stage[ftref];
                const uint4 v0 = *reinterpret_cast<const uint4 *>(cur_stage + ft0);
                absd(v0.x, ref4, acc), absd(v0.y, ref4, acc), absd(v0.z, ref4, acc), absd(v0.w, ref4, acc);
                if (tid < 112) {
                    const uint4 v1 = *reinterpret_cast<const uint4 *>(cur_stage + ft1);
                    absd(v1.x, ref4, acc), absd(v1.y, ref4, acc), absd(v1.z, ref4, acc), absd(v1.w, ref4, acc);
                } else if (tid >= 128 && tid < 220) {
                    const uint2 v1 = *reinterpret_cast<const uint2 *>(cur_stage + fte);
                    absd(v1.x, ref4, acc), absd(v1.y, ref4, acc);
                }
            } else {
                // clipped box: 16-byte items (row, 16 columns); the image is 16-px aligned, so an item lies entirely inside
                // or outside of it (outside = zero fill, must not be tested).  Only the logical columns [8, 152) matter:
                // the outer halves of the first and last item of a row are ignored.
                uint8_t(*s_g)[T::GW] = reinterpret_cast<uint8_t(*)[T::GW]>(cur_stage);
                constexpr int IPR = T::GW / 16;  // items per row
                const int r_lo = max(0, T::HALO - y0), r_hi = min(T::GH, H - y0 + T::HALO);
                const int j_lo = x0 == 0 ? 1 : 0, j_hi = min(IPR, (W - x0 + T::HX) >> 4);
                const uint32_t ref4 =
                    0x01010101u * s_g[min(T::HALO + TH / 2, r_hi - 1)][min(T::HX + TW / 2, 16 * j_hi - 1)];
                constexpr int NV = T::GH * IPR;
                for (int idx = tid; idx < NV; idx += kK1Consumers) {
                    const int r = idx / IPR, j = idx - r * IPR;
                    if (r >= r_lo && r < r_hi && j >= j_lo && j < j_hi) {
                        uint4 v = *reinterpret_cast<const uint4 *>(&s_g[r][16 * j]);
                        if (j == 0) v.x = ref4, v.y = ref4;
                        if (j == IPR - 1) v.z = ref4, v.w = ref4;
                        absd(v.x, ref4, acc), absd(v.y, ref4, acc), absd(v.z, ref4, acc), absd(v.w, ref4, acc);
                    }
                }
            }
        }
        // barrier + block-wide OR in one instruction.  No barrier closes the tile: the scratch tiles of tile k are last
        // read before (u1, bl) or during (f) its output stage, and tile k+1 first writes them behind at least one
        // barrier that every consumer reaches only after it has finished tile k.
        const bool flat = !tile_sync_or<true>(acc & 0x80808080u) && try_flat;
        if (flat && tile_inside && fast_store_ok) {
            // flat tile fully inside a 16-px aligned image: nothing but wide zero stores (one warp writes one 512-byte
            // label row per instruction)
            if (tid == 0) mbar_arrive(&empty[st]);  // every thread is past its flat-test reads of the stage
            const size_t row0 = (size_t)f * H + y0;
            const size_t pix0 = row0 * W + x0;
            const int4 z = make_int4(0, 0, 0, 0);
            if (p.init_labels) {
                int32_t *dst = b.labels + pix0 + so_lab;
#pragma unroll
                for (int k = 0; k < TH / 8; k++) *reinterpret_cast<int4 *>(dst + k * so_lab_step) = z;
            }
            if (p.write_mask) *reinterpret_cast<int4 *>(b.mask + pix0 + so_mask) = z;
            // occupancy record of the tile: 32 bytes, one per row, all zero.  The 128 all-zero bit-mask words follow unless
            // only the fused per-frame CCL kernel reads this batch (it never looks at unflagged words).
            if (b.rowflags && tid < 2)
                *reinterpret_cast<int4 *>(b.rowflags + (size_t)f * b.rf_stride + rowflag_index(y0, tx, tiles_x) + 16 * tid) = z;
            if (b.tile_occ && tid == 2) reinterpret_cast<uint32_t *>(b.tile_occ)[cur.x] = 0u;  // (the tile number is its index)
            if (!p.sparse_aux && tid < TH * (TW / 32)) bits_out[row0 * b.ww + (x0 >> 5) + so_bits] = 0u;
        } else if constexpr (MR > 0) {
            morph_tile_compute_and_store<MR>(b, p, bits_out, cur_stage, sm + NS * STAGE, f, tx, ty, flat, box_inside, tid,
                                             &empty[st]);
        } else {
            tile_compute_and_store<TW, TH, RB, 16, true>(b, p, bits_out, cur_stage, f_raw, u1_raw, bl_raw, f, tx, x0, y0, flat,
                                                         tid, &empty[st]);
        }
        if (b.phase_ns && tid == 0) {  // debug: time per tile split into TMA wait and processing, flat vs non-flat
            unsigned long long t_c;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_c));
            atomicAdd(b.phase_ns + 200 + (flat ? 0 : 3), t_b - t_a);
            atomicAdd(b.phase_ns + 201 + (flat ? 0 : 3), t_c - t_b);
            atomicAdd(b.phase_ns + 202 + (flat ? 0 : 3), 1ull);
        }
    }
    // Launched ahead of the previous batch's per-frame CCL kernel's completion (programmatic dependent launch): this grid
    // must not complete before that one has, because the kernel after us relies on "K1 complete => everything before it
    // complete" when it lets the next K1 overwrite that batch's buffers (see k_ccl_frame).  A no-op otherwise.
    // With the slots' completion counters (b.ccl_done) that inference is not needed -- every K1 checks the counter of the
    // slot it is about to overwrite -- and the wait would only serialise the per-frame kernels of consecutive batches
    // (measured: step = duration of that kernel + 7 us, whatever K1 did).
    if (tid == 0 && !b.ccl_done) asm volatile("griddepcontrol.wait;" ::: "memory");
    // launch counter for the per-frame kernel of this batch: the last CTA whose consumers have stored their last tile
    // publishes it (and rearms the CTA count)
    if (b.k1_done) {
        tile_sync<true>();  // all 256 consumers of this CTA are past their stores
        if (tid == 0) {
            __threadfence();
            if (atomicAdd(b.k1_done - 1, 1u) == gridDim.x - 1) {
                b.k1_done[-1] = 0;
                __threadfence();
                atomicAdd(b.k1_done, 1u);
            }
        }
    }
    if (b.phase_ns && tid == 0) {  // debug: CTA lifetimes
        unsigned long long t_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
        atomicMin(b.phase_ns + 248, t_cta0);
        atomicMax(b.phase_ns + 249, t_end);
        atomicAdd(b.phase_ns + 250, t_end - t_cta0);
        atomicMax(b.phase_ns + 251, t_end - t_cta0);
        atomicMin(b.phase_ns + 252, t_end);
        atomicMax(b.phase_ns + 253, t_cta0);
    }
}

__global__ void __launch_bounds__(256) k_gray3(const uint8_t *img, int n, int h, int w, int c, size_t row_stride,
                                               size_t frame_stride, uint8_t *gray) {
    const size_t total = (size_t)n * h * w;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % w);
        const size_t t = i / w;
        const int y = (int)(t % h);
        const size_t f = t / h;
        const uint8_t *px = img + f * frame_stride + (size_t)y * row_stride + (size_t)x * c;
        gray[i] = gray_f64(px[0], px[1], px[2]);
    }
}

}  // namespace

cudaError_t launch_gray3(const uint8_t *d_img, int n, int h, int w, size_t row_stride, size_t frame_stride,
                         uint8_t *d_gray, cudaStream_t s) {
    const size_t total = (size_t)n * h * w;
    const int grid = (int)(((total + 255) / 256) < (size_t)(148 * 16) ? ((total + 255) / 256) : (size_t)(148 * 16));
    k_gray3<<<grid, 256, 0, s>>>(d_img, n, h, w, 3, row_stride, frame_stride, d_gray);
    return cudaGetLastError();
}

cudaError_t launch_gray_first3(const uint8_t *d_img, int h, int w, int c, uint8_t *d_gray, cudaStream_t s) {
    const size_t total = (size_t)h * w;
    const int grid = (int)(((total + 255) / 256) < (size_t)(148 * 16) ? ((total + 255) / 256) : (size_t)(148 * 16));
    k_gray3<<<grid, 256, 0, s>>>(d_img, 1, h, w, c, (size_t)w * c, (size_t)h * w * c, d_gray);
    return cudaGetLastError();
}

namespace {

PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;  // resolved once; immutable afterwards
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

int k1_ctas_per_sm() { return tunables().k1_ctas_per_sm; }

}  // namespace

// per device, once (hv_create): the TMA kernel's stages need more than the 48 KB static shared-memory limit
cudaError_t configure_preprocess_tma() {
    cudaError_t e = cudaFuncSetAttribute(k_preprocess_tma<128, 32, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         TmaSmem<128, 32, 2>::BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_preprocess_tma<128, 32, kGaussRB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             TmaSmem<128, 32, kGaussRB>::BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_preprocess_tma<128, 32, kGaussRBSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             TmaSmem<128, 32, kGaussRBSmall>::BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_preprocess_tma<128, 32, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             TmaSmem<128, 32, 2, 4>::BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_preprocess_tma<128, 32, 2, 4>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_preprocess_tma<128, 32, 2, 4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             TmaSmem<128, 32, 2, 4, 3>::BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_preprocess_tma<128, 32, 2, 4, 3>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_preprocess_tma<128, 32, 2, 0, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             TmaSmem<128, 32, 2, 0, 3>::BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_preprocess_tma<128, 32, 2, 0, 3>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    // ask for the largest shared-memory carve-out: residency of these kernels is limited by shared memory, not by L1
    e = cudaFuncSetAttribute(k_preprocess_tma<128, 32, 2>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_preprocess_tma<128, 32, kGaussRB>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_preprocess_tma<128, 32, kGaussRBSmall>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_preprocess<128, 32, 2>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_preprocess<128, 32, 0>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                cudaSharedmemCarveoutMaxShared);
}

// 3x3 / 5x5 open and close whose total reach is at most 4 px (MTile<4>), on a batch the TMA kernel takes at all, with
// the standard comparison (inverse, 0 <= c <= 255, no blur debug plane, no forced generic path)
bool preprocess_tma_morph_supported(const BatchView &b, const PreprocessParams &p, int open_k, int close_k) {
    auto ok_k = [](int k) { return k == 0 || k == 3 || k == 5; };
    if (!ok_k(open_k) || !ok_k(close_k) || (open_k == 0 && close_k == 0)) return false;
    if ((open_k > 0 ? open_k - 1 : 0) + (close_k > 0 ? close_k - 1 : 0) > 4) return false;
    const uintptr_t base = reinterpret_cast<uintptr_t>(b.gray);
    if (p.blur_radius != 2 || p.gauss_ksize > 0 || (base & 15) || (b.gray_row_stride & 15) || (b.gray_frame_stride & 15) || (b.w & 15) ||
        tunables().k1_no_tma || tunables().no_k1_morph)
        return false;
    return p.inverse && p.c_thresh >= 0 && p.c_thresh <= 255 && !p.write_blur && !p.force_generic && !p.wrap_t1 && p.write_mask &&
           p.init_labels && tensor_map_encoder() != nullptr;
}

// TMA path: 3-D tensor map {w, h, n} over the gray frames, box = tile + halo, zero fill outside the image.
// p.gauss_ksize > 0 selects the Gaussian variant (A7 fused into K1): same pipeline, RB = 7 rows more halo.
cudaError_t launch_preprocess_tma(const BatchView &b, const PreprocessParams &p, uint32_t *bits_out, unsigned int *sched,
                                  int num_sms, bool pdl, cudaStream_t s, bool *used) {
    *used = false;
    const bool gauss = p.gauss_ksize > 0;
    const uintptr_t base = reinterpret_cast<uintptr_t>(b.gray);
    if ((!gauss && p.blur_radius != 2) || !sched || (base & 15) || (b.gray_row_stride & 15) || (b.gray_frame_stride & 15) ||
        (b.w & 15) || tunables().k1_no_tma)
        return cudaSuccess;
    if (gauss && (p.gauss_ksize > 2 * kGaussRB + 1 || !(p.gauss_ksize & 1) || b.h < 16 || b.w < 16)) return cudaSuccess;
    auto enc = tensor_map_encoder();
    if (!enc) return cudaSuccess;
    const bool morph = !gauss && (p.morph_open_k > 0 || p.morph_close_k > 0);
    if (morph && !preprocess_tma_morph_supported(b, p, p.morph_open_k, p.morph_close_k)) return cudaSuccess;
    const bool gauss_small = gauss && p.gauss_ksize <= 2 * kGaussRBSmall + 1;  // the variant with the smaller halo
    const int gh = gauss ? (gauss_small ? Tile<128, 32, kGaussRBSmall, 16>::GH : Tile<128, 32, kGaussRB, 16>::GH) : (morph ? MTile<4>::GH : Tile<128, 32, 2, 16>::GH);
    CUtensorMap tmap;
    const cuuint64_t dims[3] = {(cuuint64_t)b.w, (cuuint64_t)b.h, (cuuint64_t)b.n};
    const cuuint64_t strides[2] = {(cuuint64_t)b.gray_row_stride, (cuuint64_t)b.gray_frame_stride};
    const cuuint32_t box[3] = {(cuuint32_t)Tile<128, 32, 2, 16>::GW, (cuuint32_t)gh, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t *>(b.gray), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return cudaSuccess;  // fall back to the non-TMA kernel
    const int tiles = b.tiles_x * ((b.h + 31) / 32) * b.n;
    const int gauss_ctas = tunables().k1_gauss_ctas;
    int grid = num_sms * (gauss ? (gauss_small ? tunables().k1_gauss_small_ctas : gauss_ctas) : (p.ctas_per_sm > 0 ? std::min(p.ctas_per_sm, k1_ctas_per_sm()) : k1_ctas_per_sm()));
    if (morph) grid = std::min(grid, num_sms * 4);  // 50 KB of shared memory per CTA
    if (grid > tiles) grid = tiles;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kK1Threads);
    // three stages when the kernel runs at three CTAs per SM next to the per-frame CCL kernel (p.stages, set by hv_api.cu)
    const bool s3 = !gauss && p.stages == 3;
    cfg.dynamicSmemBytes = gauss ? (gauss_small ? TmaSmem<128, 32, kGaussRBSmall>::BYTES : TmaSmem<128, 32, kGaussRB>::BYTES)
                                 : (morph ? (s3 ? TmaSmem<128, 32, 2, 4, 3>::BYTES : TmaSmem<128, 32, 2, 4>::BYTES)
                                          : (s3 ? TmaSmem<128, 32, 2, 0, 3>::BYTES : TmaSmem<128, 32, 2>::BYTES));
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    *used = true;
    PreprocessParams q = p;
    const Tunables &tun = tunables();
    q.lookahead = std::min(s3 ? std::max(tun.k1_lookahead, 3) : tun.k1_lookahead, s3 ? 3 : kTmaStages);
    q.tail_lookahead = std::min(tun.k1_tail_lookahead, q.lookahead);
    q.tail_tiles = tun.k1_tail_rounds * grid;
    q.prefetch_tiles = tun.k1_prefetch;
    // Requesting the next tile number one tile early hides the atomic's round trip when K1 runs at reduced residency next
    // to the per-frame kernel (three or four CTAs per SM: 39.70 -> 39.48 us per headline batch, 47.9 -> 47.3 with the
    // morphology variant, 155.6 -> 154.6 for 100 frames) and costs when it has all five (64 x 5 MP: 364.6 -> 374.0 us).
    q.claim_ahead = tun.k1_claim_ahead >= 0 ? tun.k1_claim_ahead : ((p.ctas_per_sm > 0 && p.ctas_per_sm < k1_ctas_per_sm()) ? 1 : 0);
    q.wait_hint_ns = tun.k1_wait_hint_ns;
    if (gauss_small) return cudaLaunchKernelEx(&cfg, k_preprocess_tma<128, 32, kGaussRBSmall>, tmap, b, q, bits_out, sched);
    if (gauss) return cudaLaunchKernelEx(&cfg, k_preprocess_tma<128, 32, kGaussRB>, tmap, b, q, bits_out, sched);
    if (morph && s3) return cudaLaunchKernelEx(&cfg, k_preprocess_tma<128, 32, 2, 4, 3>, tmap, b, q, bits_out, sched);
    if (morph) return cudaLaunchKernelEx(&cfg, k_preprocess_tma<128, 32, 2, 4>, tmap, b, q, bits_out, sched);
    if (s3) return cudaLaunchKernelEx(&cfg, k_preprocess_tma<128, 32, 2, 0, 3>, tmap, b, q, bits_out, sched);
    return cudaLaunchKernelEx(&cfg, k_preprocess_tma<128, 32, 2>, tmap, b, q, bits_out, sched);
}

cudaError_t launch_preprocess(const BatchView &b, const PreprocessParams &p, uint32_t *bits_out, cudaStream_t s) {
    constexpr int TW = 128, TH = 32;
    dim3 grid((b.w + TW - 1) / TW, (b.h + TH - 1) / TH, b.n);
    if (p.blur_radius == 2)
        k_preprocess<TW, TH, 2><<<grid, 256, 0, s>>>(b, p, bits_out);
    else if (p.blur_radius == 0)
        k_preprocess<TW, TH, 0><<<grid, 256, 0, s>>>(b, p, bits_out);
    else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}

}  // namespace hv
