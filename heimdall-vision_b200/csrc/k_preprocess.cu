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
#include "hv_common.cuh"

namespace hv {

namespace {

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// exact floor(s / 25) for s <= 6375
__device__ __forceinline__ uint32_t div25(uint32_t s) { return (s * 5243u) >> 17; }

// Shared-memory plan of one tile (TW x TH outputs, 256 threads).  Column index i of s_g / s_bl is image column
// x0 - 8 + i; row r of s_g is image row y0 - HALO + r; row r of s_v / s_bl / s_h11 is image row y0 - 5 + r.
//   s_g   [GH][GW]  u8   gray tile + halo                                  (dead after the blur phase)
//   s_f   [TH][TW]  u8   mask bytes of the tile            -- aliases s_g
//   s_v   [BH][VP]  u16  vertical 5-sums, column i stored at i + 4         (fast path only, dead after the blur phase)
//   s_h11 [BH][TW]  u16  horizontal 11-sums of the blur    -- aliases s_v
//   s_bl  [BH][GW]  u8   blurred tile + 5-px ring
template <int TW, int TH, int RB>
struct Tile {
    static constexpr int HALO = RB + kAdaptHalf;  // rows of gray needed above/below
    static constexpr int HX = 8;                  // column halo, padded to a multiple of 4 for aligned loads
    static constexpr int GW = TW + 2 * HX;
    static constexpr int GH = TH + 2 * HALO;
    static constexpr int BH = TH + 2 * kAdaptHalf;
    static constexpr int VP = GW + 8;  // pitch of s_v in u16 (column i is stored at i + 4; 4 spare on each side)
    static constexpr int G_BYTES = GH * GW;
    static constexpr int U1_BYTES = (BH * VP * 2 > BH * TW * 2) ? BH * VP * 2 : BH * TW * 2;
    static constexpr int BL_BYTES = BH * GW;
    static_assert(TH * TW <= G_BYTES, "s_f must fit inside s_g");
    static_assert((G_BYTES % 16) == 0 && (U1_BYTES % 16) == 0, "alignment");
};

// ---- fast path, interior tiles only (every blur pixel is an interior pixel, every window has 121 pixels) ---------
// Packed u16x2 arithmetic: sums of u8 never overflow a 16-bit lane (5x5: 6375, 11x11 of blur: 30855), so plain 32-bit
// adds/subs act on both lanes at once; VIMNMX.U16x2 gives the lane-wise compare for the threshold test.
template <int TW, int TH>
__device__ __forceinline__ void fast_tile_rb2(uint8_t *smem, int tid, int cth) {
    using T = Tile<TW, TH, 2>;
    uint8_t(*s_g)[T::GW] = reinterpret_cast<uint8_t(*)[T::GW]>(smem);
    uint8_t(*s_f)[TW] = reinterpret_cast<uint8_t(*)[TW]>(smem);
    uint16_t(*s_v)[T::VP] = reinterpret_cast<uint16_t(*)[T::VP]>(smem + T::G_BYTES);
    uint16_t(*s_h11)[TW] = reinterpret_cast<uint16_t(*)[TW]>(smem + T::G_BYTES);
    uint8_t(*s_bl)[T::GW] = reinterpret_cast<uint8_t(*)[T::GW]>(smem + T::G_BYTES + T::U1_BYTES);
    static_assert(TW == 128 && TH == 32, "thread mappings below are written for 128x32 tiles");

    // A. vertical 5-sums: thread = (column quad q, row segment of 7) -> 36 x 6 = 216 threads
    if (tid < 36 * 6) {
        const int q = tid % 36, seg = tid / 36;
        const int r0 = seg * 7;
        uint32_t lo[5], hi[5];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t v = *reinterpret_cast<const uint32_t *>(&s_g[r0 + k][4 * q]);
            lo[k] = prmt(v, 0, 0x4140);
            hi[k] = prmt(v, 0, 0x4342);
        }
        uint32_t alo = lo[0] + lo[1] + lo[2] + lo[3], ahi = hi[0] + hi[1] + hi[2] + hi[3];
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const uint32_t v = *reinterpret_cast<const uint32_t *>(&s_g[r0 + k + 4][4 * q]);
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
    __syncthreads();
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
    __syncthreads();
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
    __syncthreads();
    // D. vertical 11-sums + threshold test.  thread = (column quad, segment of 8 rows) -> 32 x 4 = 128 threads.
    //    fg  <=>  (px + c + 1) * 121 <= S  <=>  S + 1 > px * 121 + (c + 1) * 121      (all lanes < 65536 for 0 <= c <= 255)
    if (tid < 32 * 4) {
        const int q = tid & 31, seg = tid >> 5;
        const int r0 = seg * 8;
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
        for (int k = 0; k < 8; k++) {
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
    __syncthreads();
}

template <int TW, int TH, int RB>
__global__ void __launch_bounds__(256, 8) k_preprocess(BatchView b, PreprocessParams p, uint32_t *bits_out) {
    using T = Tile<TW, TH, RB>;
    static_assert(TW % 32 == 0, "tile width must cover whole bitmask words");
    __shared__ __align__(16) uint8_t smem[T::G_BYTES + T::U1_BYTES + T::BL_BYTES];
    uint8_t(*s_g)[T::GW] = reinterpret_cast<uint8_t(*)[T::GW]>(smem);
    uint8_t(*s_f)[TW] = reinterpret_cast<uint8_t(*)[TW]>(smem);  // aliases s_g (dead by then)
    uint16_t(*s_h11)[TW] = reinterpret_cast<uint16_t(*)[TW]>(smem + T::G_BYTES);
    uint8_t(*s_bl)[T::GW] = reinterpret_cast<uint8_t(*)[T::GW]>(smem + T::G_BYTES + T::U1_BYTES);

    const int tid = threadIdx.x;
    const int f = blockIdx.z;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int H = b.h, W = b.w;
    const uint8_t *gray = b.gray + (size_t)f * b.gray_frame_stride;
    const size_t gpitch = b.gray_row_stride;
    const bool aligned_in = ((reinterpret_cast<uintptr_t>(gray) | gpitch) & 3) == 0;
    const bool vec = (W & 3) == 0;
    uint8_t *mask = b.mask + (size_t)f * H * W;
    int32_t *labels = b.labels + (size_t)f * H * W;

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
    // interior tile: the blur ring (tile +- 5) lies >= RB pixels inside the image, so no border rule applies anywhere
    const bool interior = x0 >= T::HALO && y0 >= T::HALO && x0 + TW + T::HALO <= W && y0 + TH + T::HALO <= H;

    if (!flat) {
        if (RB == 2 && TW == 128 && TH == 32 && interior && p.inverse && !p.write_blur && cth >= 0 && cth <= 255 &&
            !p.force_generic) {
            fast_tile_rb2<128, 32>(smem, tid, cth);
        } else {
            // ---- generic path: any border, any c, either comparison direction --------------------------------------
            // 2. blur over the tile + 5-px ring (zero outside the image, pass-through outside the interior)
            for (int idx = tid; idx < T::BH * (TW + 2 * kAdaptHalf); idx += 256) {
                const int r = idx / (TW + 2 * kAdaptHalf), i = idx - r * (TW + 2 * kAdaptHalf) + (T::HX - kAdaptHalf);
                const int gy = y0 - kAdaptHalf + r, gx = x0 - T::HX + i;
                const int ry = r + RB;
                uint32_t v = 0;
                if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                    if (RB > 0 && gy >= RB && gy < H - RB && gx >= RB && gx < W - RB) {
                        uint32_t s = 0;
#pragma unroll
                        for (int dy = -RB; dy <= RB; dy++)
#pragma unroll
                            for (int dx = -RB; dx <= RB; dx++) s += s_g[ry + dy][i + dx];
                        v = (RB == 2) ? div25(s) : s / ((2 * RB + 1) * (2 * RB + 1));
                    } else {
                        v = s_g[ry][i];
                    }
                    if (p.write_blur && r >= kAdaptHalf && r < kAdaptHalf + TH && i >= T::HX && i < T::HX + TW)
                        b.blur[((size_t)f * H + gy) * W + gx] = (uint8_t)v;
                }
                s_bl[r][i] = (uint8_t)v;
            }
            __syncthreads();
            // 3. horizontal 11-sums: output column c uses s_bl columns c + 3 .. c + 13
            for (int idx = tid; idx < T::BH * TW; idx += 256) {
                const int r = idx / TW, c = idx - r * TW;
                uint32_t s = 0;
#pragma unroll
                for (int k = 0; k < 2 * kAdaptHalf + 1; k++) s += s_bl[r][c + (T::HX - kAdaptHalf) + k];
                s_h11[r][c] = (uint16_t)s;
            }
            __syncthreads();
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
                    const int px = s_bl[r + kAdaptHalf][c + T::HX];
                    const bool t = p.inverse ? ((px + cth + 1) * cnt <= s) : !((px + cth) * cnt <= s);
                    fg = t ? 255 : 0;
                }
                s_f[r][c] = fg;
            }
            __syncthreads();
        }
    }

    // ---- 5. outputs: u8 mask, bit-packed mask, label plane ------------------------------------------------------
    if (flat && x0 + TW <= W && y0 + TH <= H && (W & 15) == 0) {
        // flat tile fully inside a 16-px aligned image: nothing but wide zero stores (one warp writes one 512-byte
        // label row per instruction)
        const int4 z = make_int4(0, 0, 0, 0);
        if (p.init_labels) {
#pragma unroll
            for (int r = tid >> 5; r < TH; r += 8)
                *reinterpret_cast<int4 *>(labels + (size_t)(y0 + r) * W + x0 + 4 * (tid & 31)) = z;
        }
        if (p.write_mask) {
            for (int idx = tid; idx < TH * (TW / 16); idx += 256) {
                const int r = idx / (TW / 16), c16 = idx - r * (TW / 16);
                *reinterpret_cast<int4 *>(mask + (size_t)(y0 + r) * W + x0 + 16 * c16) = z;
            }
        }
        if (tid < TH * (TW / 32)) {
            const int r = tid / (TW / 32), wq = tid - r * (TW / 32);
            bits_out[((size_t)f * H + y0 + r) * b.ww + (x0 >> 5) + wq] = 0u;
        }
        if (b.rowflags && tid < TH) b.rowflags[(size_t)f * b.rf_stride + (size_t)(y0 + tid) * b.tiles_x + blockIdx.x] = 0;
        return;
    }
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
            b.rowflags[(size_t)f * b.rf_stride + (size_t)gy * b.tiles_x + blockIdx.x] =
                (uint8_t)((bal >> (tid & 31)) & 0xfu);
    }
}

// f64 gray with the reference's expression order: (0.299*c0 + 0.587*c1) + 0.114*c2, truncated
// (detection.rs:138-150).  __dmul_rn/__dadd_rn keep the compiler from contracting into FMAs.
__device__ __forceinline__ uint8_t gray_f64(uint32_t c0, uint32_t c1, uint32_t c2) {
    const double t0 = __dmul_rn(0.299, (double)c0);
    const double t1 = __dmul_rn(0.587, (double)c1);
    const double t2 = __dmul_rn(0.114, (double)c2);
    const double s = __dadd_rn(__dadd_rn(t0, t1), t2);
    int v = __double2int_rz(s);
    v = v < 0 ? 0 : (v > 255 ? 255 : v);
    return (uint8_t)v;
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
