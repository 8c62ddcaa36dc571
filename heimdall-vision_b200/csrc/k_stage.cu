// k_stage.cu -- the stages around the fused hot kernel: morphology on the bit-packed mask (A8), the OpenCV-exact
// Gaussian blur (A7), and the single-frame utilities behind heimdall_core.processing.* / process_image
// (rust/heimdall-core/src/processing.rs).  None of these is on the default (Rust-exact) detect path.
#include <algorithm>
#include <cstdlib>

#include "hv_common.cuh"
#include "expand_tile.cuh"

namespace hv {

namespace {

// ---------------------------------------------------------------------------------------------------------
// A8: rect k x k erode / dilate, anchor k/2, OpenCV default border (out-of-image never wins).
// Spec: heimdall/detectors/contamination_detector.py:81-87 (cv2.morphologyEx MORPH_OPEN / MORPH_CLOSE, 3x3 rect),
// heimdall/core/pipeline.py:290-332.  Works on 32 pixels per word: the horizontal window is a fold of funnel
// shifts over (left, mid, right) words, the vertical window a fold over rows.
// ---------------------------------------------------------------------------------------------------------
template <bool DILATE>
__device__ __forceinline__ uint32_t load_word(const uint32_t *rowp, int wx, int ww, uint32_t tail_mask) {
    // value outside the image: 0 for dilate, 1 for erode
    if (wx < 0 || wx >= ww) return DILATE ? 0u : 0xffffffffu;
    uint32_t v = rowp[wx];
    if (wx == ww - 1) v = DILATE ? (v & tail_mask) : (v | ~tail_mask);
    return v;
}

// Separable: a k x k rectangle is a 1 x k row window followed by a k x 1 column window (min and max commute with the
// product structure), and "out of the image never wins" holds per pass.  The first version did the full k x k fold per
// word (k rows x 2k funnel shifts: 156 us per pass for 16 x 5 MP at k = 15); the two passes cost k shifts + k loads.
template <bool DILATE>
__global__ void __launch_bounds__(256) k_morph_h(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, int n,
                                                 int h, int ww, int w, int k) {
    const int a = k / 2;                // anchor
    const int lo = -a, hi = k - 1 - a;  // window offsets [lo, hi]
    const uint32_t tail_mask = (w & 31) ? ((1u << (w & 31)) - 1u) : 0xffffffffu;
    const size_t total = (size_t)n * h * ww;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int wx = (int)(i % ww);
        const uint32_t *rowp = src + (i - wx);
        const uint32_t L = load_word<DILATE>(rowp, wx - 1, ww, tail_mask);
        const uint32_t M = load_word<DILATE>(rowp, wx, ww, tail_mask);
        const uint32_t R = load_word<DILATE>(rowp, wx + 1, ww, tail_mask);
        uint32_t acc = M;
        for (int dx = 1; dx <= hi; dx++) {  // pixel x + dx
            const uint32_t t = __funnelshift_r(M, R, dx);
            acc = DILATE ? (acc | t) : (acc & t);
        }
        for (int dx = 1; dx <= -lo; dx++) {  // pixel x - dx
            const uint32_t t = __funnelshift_l(L, M, dx);
            acc = DILATE ? (acc | t) : (acc & t);
        }
        if (wx == ww - 1) acc &= tail_mask;
        dst[i] = acc;
    }
}

template <bool DILATE>
__global__ void __launch_bounds__(256) k_morph_v(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, int n,
                                                 int h, int ww, int k) {
    const int a = k / 2;
    const int lo = -a, hi = k - 1 - a;
    const size_t words_per_frame = (size_t)h * ww;
    const size_t total = words_per_frame * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t f = i / words_per_frame;
        const int wi = (int)(i - f * words_per_frame);
        const int y = wi / ww;
        const int y0 = max(y + lo, 0), y1 = min(y + hi, h - 1);
        const uint32_t *colp = src + i - (size_t)(y - y0) * ww;
        uint32_t acc = DILATE ? 0u : 0xffffffffu;
        for (int yy = y0; yy <= y1; yy++, colp += ww) acc = DILATE ? (acc | __ldg(colp)) : (acc & __ldg(colp));
        dst[i] = acc;
    }
}

// Mask, labels and occupancy records from a bit-packed mask, one 128 x 32 tile per CTA-iteration.  Used after the
// multi-pass morphology (kernel sizes above 15).
__global__ void __launch_bounds__(256) k_expand_bits(BatchView b) {
    __shared__ uint32_t s_w[32][4];
    const int tid = threadIdx.x;
    const int H = b.h;
    const int tiles_y = (H + 31) / 32;
    const size_t per_frame = (size_t)tiles_y * b.tiles_x, total = per_frame * b.n;
    for (size_t t = blockIdx.x; t < total; t += gridDim.x) {
        const size_t f = t / per_frame;
        const int j = (int)(t - f * per_frame);
        const int ty = j / b.tiles_x, tx = j - ty * b.tiles_x;
        const int y0 = ty * 32;
        uint32_t word = 0;
        if (tid < 128) {
            const int r = tid >> 2, wq = tid & 3;
            if (y0 + r < H && 4 * tx + wq < b.ww) word = __ldg(b.bits + (f * H + y0 + r) * (size_t)b.ww + 4 * tx + wq);
            s_w[r][wq] = word;
            const uint32_t bal = __ballot_sync(0xffffffffu, word != 0);
            if (wq == 0 && y0 + r < H && b.rowflags)
                b.rowflags[f * b.rf_stride + rowflag_index(y0 + r, tx, b.tiles_x)] = (uint8_t)((bal >> (tid & 31)) & 0xfu);
        }
        const bool any = __syncthreads_or(word != 0);
        expand_tile(b, f, tx, ty, s_w, any, tid);
        __syncthreads();  // s_w is rewritten by the next tile
    }
}

// A8 in one pass (kernel sizes up to 15): K1 left the pre-morphology mask bit-packed; a CTA stages the words of a
// 128 x 32 tile plus the reach of open followed by close (at most 28 pixels = 28 rows and one word on either side:
// 88 x 6 words, 2 KB), runs the separable passes Eh Ev Dh Dv (open) Dh Dv Eh Ev (close) between two shared-memory
// buffers and expands the centre into the mask bytes, the label plane, the final bit-mask (bits_out, a buffer other
// than the input: neighbouring CTAs still read the original) and the occupancy records -- 5.1 B/px written once, where
// the multi-pass version re-read and re-wrote the bit plane eight times and expanded it in a ninth kernel.  A staged
// region without a set bit yields an empty tile (no pass can create foreground out of nothing), which is the common
// case and costs one 2 KB read and the zero stores.  "Out of the image never wins" (OpenCV's default border) is applied
// when a pass READS a word: bits outside the image count as 1 for erode and 0 for dilate, whatever the previous pass
// left there.  Garbage entering from outside the staged region travels at most the reach and never gets to the centre.
constexpr int kMorphReach = 28;
constexpr int kMorphRows = 32 + 2 * kMorphReach;
constexpr int kMorphCols = 6;

// One separable pass over the staged rows [r_lo, r_hi): window [-lo, +hi] along the row (H) or the column (V).  What
// depends only on the thread's words -- their position and the in-image masks of the word and its neighbours -- is
// computed once per tile (MorphItem), not once per pass: the passes themselves are a few loads, shifts and stores.
// packed: staged row | column << 8 | flags << 16, flags: 1 row inside the image, 2 / 4 / 8 word c-1 / c / c+1 inside the
// image (and the staged region), 16 / 32 / 64 that word is the row's last one (tail mask); 0xffffffff = no word
__device__ __forceinline__ uint32_t morph_word_mask(uint32_t item, uint32_t in_bit, uint32_t tail_bit, uint32_t tail_mask,
                                                    bool need_row) {
    const uint32_t fl = item >> 16;
    if (!(fl & in_bit) || (need_row && !(fl & 1u))) return 0u;
    return (fl & tail_bit) ? tail_mask : 0xffffffffu;
}

template <bool DILATE>
__device__ __forceinline__ void morph_pass_h(const uint32_t (*src)[kMorphCols], uint32_t (*dst)[kMorphCols], int lo, int hi,
                                             const uint32_t (&it)[3], uint32_t tail_mask) {
#pragma unroll
    for (int q = 0; q < 3; q++) {
        if (it[q] == 0xffffffffu) continue;
        const int r = it[q] & 0xff, c = (it[q] >> 8) & 0xff;
        // outside the image (or the staged region) a word reads as the identity: 0 for dilate, all ones for erode
        const uint32_t ml = morph_word_mask(it[q], 2u, 16u, tail_mask, true), mc = morph_word_mask(it[q], 4u, 32u, tail_mask, true),
                       mr = morph_word_mask(it[q], 8u, 64u, tail_mask, true);
        uint32_t L = c > 0 ? src[r][c - 1] : 0u, M = src[r][c], R = c < kMorphCols - 1 ? src[r][c + 1] : 0u;
        L = DILATE ? (L & ml) : (L | ~ml);
        M = DILATE ? (M & mc) : (M | ~mc);
        R = DILATE ? (R & mr) : (R | ~mr);
        uint32_t acc = M;
        for (int dx = 1; dx <= hi; dx++) {
            const uint32_t t = __funnelshift_r(M, R, dx);
            acc = DILATE ? (acc | t) : (acc & t);
        }
        for (int dx = 1; dx <= lo; dx++) {
            const uint32_t t = __funnelshift_l(L, M, dx);
            acc = DILATE ? (acc | t) : (acc & t);
        }
        dst[r][c] = acc;
    }
}

// rows [v_lo, v_hi] are the staged rows that lie inside the image (and inside the staged region)
template <bool DILATE>
__device__ __forceinline__ void morph_pass_v(const uint32_t (*src)[kMorphCols], uint32_t (*dst)[kMorphCols], int lo, int hi,
                                             const uint32_t (&it)[3], uint32_t tail_mask, int v_lo, int v_hi) {
#pragma unroll
    for (int q = 0; q < 3; q++) {
        if (it[q] == 0xffffffffu) continue;
        const int r = it[q] & 0xff, c = (it[q] >> 8) & 0xff;
        uint32_t acc = DILATE ? 0u : 0xffffffffu;
        const int a_lo = max(r - lo, v_lo), a_hi = min(r + hi, v_hi);
        for (int rr = a_lo; rr <= a_hi; rr++) {
            const uint32_t v = src[rr][c];
            acc = DILATE ? (acc | v) : (acc & v);
        }
        // the column mask of the word is the same in every row: apply it once
        const uint32_t colm = morph_word_mask(it[q], 4u, 32u, tail_mask, false);
        dst[r][c] = DILATE ? (acc & colm) : (acc | ~colm);
    }
}

// Which tiles can open + close change at all?  Only those with foreground within reach.  K1 left 4 bytes per tile (byte g:
// which of the tile's 4 words are non-zero in rows 8g..8g+7); one thread per tile looks at its 3 x 3 neighbourhood --
// of the tiles to the left only the last word matters (reach < 32 px), of those to the right the first, of the tiles above
// and below the row groups within reach -- and appends the tile to a work list (warp-aggregated).  A tile that is not
// listed keeps what K1 wrote for it (zeros: its own mask is empty); its occupancy record for the CCL is zeroed here.
// chain (optional): {scan CTAs done, scan launches done, tiles launches done} -- with it, and with K1's launch counter in
// b.k1_done, the kernels of the morphology variant hand over through counters like the plain path does (see
// k_ccl_frame.cu: griddepcontrol.wait would make every kernel wait for the whole previous batch).
__global__ void __launch_bounds__(256) k_morph_scan(BatchView b, int reach, uint8_t *rowflags_out, uint32_t *tile_list,
                                                    unsigned int *ctrl, unsigned int *chain) {
    if (chain && b.k1_done) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        if (threadIdx.x == 0) {
            const volatile unsigned int *flag = b.k1_done;
            while ((int)(*flag - b.k1_wait_value) < 0) __nanosleep(200);
            __threadfence();
        }
        __syncthreads();
    } else {
        asm volatile("griddepcontrol.wait;" ::: "memory");  // K1 may still be running (programmatic dependent launch)
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }
    const int tiles_y = (b.h + 31) / 32;
    const size_t per_frame = (size_t)tiles_y * b.tiles_x, total = per_frame * b.n;
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    bool busy = false;
    if (t < total) {
        const size_t f = t / per_frame;
        const int j = (int)(t - f * per_frame);
        const int ty = j / b.tiles_x, tx = j - ty * b.tiles_x;
        const uint32_t *occ = reinterpret_cast<const uint32_t *>(b.tile_occ) + f * per_frame;
        // row groups of the neighbours above / below that lie within reach: group g of the tile above is 25 - 8g rows
        // away (its last row), group g of the tile below 8g + 1
        uint32_t up_rows = 0, down_rows = 0;
        for (int g = 0; g < 4; g++) {
            if (25 - 8 * g <= reach) up_rows |= 0xffu << (8 * g);
            if (8 * g + 1 <= reach) down_rows |= 0xffu << (8 * g);
        }
        uint32_t acc = 0;
        for (int dy = -1; dy <= 1; dy++) {
            const int ny = ty + dy;
            if (ny < 0 || ny >= tiles_y) continue;
            const uint32_t rows = dy < 0 ? up_rows : (dy > 0 ? down_rows : 0xffffffffu);
            for (int dx = -1; dx <= 1; dx++) {
                const int nx = tx + dx;
                if (nx < 0 || nx >= b.tiles_x) continue;
                const uint32_t cols = dx < 0 ? 0x08080808u : (dx > 0 ? 0x01010101u : 0x0f0f0f0fu);
                acc |= __ldcg(occ + (size_t)ny * b.tiles_x + nx) & rows & cols;  // (L2: written by a K1 this CTA may have waited for)
            }
        }
        busy = acc != 0;
        if (!busy) {
            int4 *rec = reinterpret_cast<int4 *>(rowflags_out + f * b.rf_stride + rowflag_index(ty * 32, tx, b.tiles_x));
            rec[0] = make_int4(0, 0, 0, 0);
            rec[1] = make_int4(0, 0, 0, 0);
        }
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, busy);
    if (bal) {
        const int lane = threadIdx.x & 31;
        unsigned int base = 0;
        if (lane == 0) base = atomicAdd(ctrl, (unsigned int)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (busy) tile_list[base + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)t;
    }
    if (chain) {  // the last CTA of the launch publishes "scan done"
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(chain, 1u) == gridDim.x - 1) {
                chain[0] = 0;
                __threadfence();
                atomicAdd(chain + 1, 1u);
            }
        }
    }
}

// A8 for the listed tiles: persistent CTAs take entries of the work list through an atomic cursor (requested one
// iteration ahead, so its round trip is never waited for).  ctrl = {list length, cursor, CTAs done}; the last CTA to
// finish rearms all three for the next launch.
__global__ void __launch_bounds__(256) k_morph_tiles(BatchView b, int open_k, int close_k, uint32_t *bits_out,
                                                     uint8_t *rowflags_out, const uint32_t *tile_list, unsigned int *ctrl,
                                                     unsigned int *chain, unsigned int scan_expected) {
    __shared__ uint32_t s_a[kMorphRows][kMorphCols], s_b[kMorphRows][kMorphCols];
    __shared__ uint32_t s_w[32][4];
    __shared__ unsigned int s_next[2];
    const int tid = threadIdx.x;
    const int H = b.h, W = b.w, WW = b.ww;
    const int tiles_y = (H + 31) / 32;
    const size_t per_frame = (size_t)tiles_y * b.tiles_x;
    const uint32_t tail_mask = (W & 31) ? ((1u << (W & 31)) - 1u) : 0xffffffffu;
    // window [-lo, +hi] of the two kernel sizes (anchor k/2, as cv2), and the staged rows: the tile plus what open + close
    // can reach (4 rows for two 3 x 3 kernels, 28 for two 15 x 15 ones)
    const int lo_o = open_k > 0 ? open_k / 2 : 0, hi_o = open_k > 0 ? open_k - 1 - open_k / 2 : 0;
    const int lo_c = close_k > 0 ? close_k / 2 : 0, hi_c = close_k > 0 ? close_k - 1 - close_k / 2 : 0;
    const int reach = min(kMorphReach, 2 * max(lo_o, hi_o) + 2 * max(lo_c, hi_c));
    const int r_lo = kMorphReach - reach, r_hi = kMorphReach + 32 + reach;
    const int n_staged = (r_hi - r_lo) * kMorphCols;
    if (chain) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        if (tid == 0) {
            const volatile unsigned int *flag = chain + 1;
            while ((int)(*flag - scan_expected) < 0) __nanosleep(200);
            __threadfence();
        }
        __syncthreads();
    } else {
        asm volatile("griddepcontrol.wait;" ::: "memory");  // the scan (and K1 before it) have completed
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }
    const unsigned int count = *reinterpret_cast<volatile unsigned int *>(ctrl);
    // thread 0 requests the next list entry AND loads its tile number one iteration ahead (two dependent round trips, 16 %
    // of the warp samples of this kernel when they sat at the top of the iteration)
    __shared__ uint32_t s_tile[2];
    unsigned int pending = 0;
    uint32_t pending_tile = 0;
    if (tid == 0) {
        const unsigned int first = atomicAdd(ctrl + 1, 1u);
        s_next[0] = first;
        s_tile[0] = first < count ? __ldcg(tile_list + first) : 0u;
        pending = atomicAdd(ctrl + 1, 1u);
        pending_tile = pending < count ? __ldcg(tile_list + pending) : 0u;
    }
    __syncthreads();
    unsigned int cur = s_next[0];
    uint32_t t = s_tile[0];
    for (int it = 1; cur < count; it++) {
        if (tid == 0) {
            s_next[it & 1] = pending;
            s_tile[it & 1] = pending_tile;
            if (pending < count) {
                pending = atomicAdd(ctrl + 1, 1u);
                pending_tile = pending < count ? __ldcg(tile_list + pending) : 0u;
            }
        }
        // 32-bit arithmetic: the 64-bit divisions were a quarter of the kernel's instructions
        const uint32_t f32 = t / (uint32_t)per_frame;
        const size_t f = f32;
        const int j = (int)(t - f32 * (uint32_t)per_frame);
        const int ty = j / b.tiles_x, tx = j - ty * b.tiles_x;
        const int y0 = ty * 32, wx0 = tx * 4 - 1;
        const uint32_t *fb = b.bits + f * (size_t)H * WW;
        // all loads of the thread first, then the shared-memory stores: one global-memory latency per tile
        uint32_t v[3];
        uint32_t item[3];
        auto col_flags = [&](int cc, uint32_t in_bit, uint32_t tail_bit) -> uint32_t {
            const int gwx = wx0 + cc;
            if (cc < 0 || cc >= kMorphCols || gwx < 0 || gwx >= WW) return 0u;
            return in_bit | (gwx == WW - 1 ? tail_bit : 0u);
        };
#pragma unroll
        for (int q = 0; q < 3; q++) {
            const int idx = tid + 256 * q;
            v[q] = 0;
            item[q] = 0xffffffffu;
            if (idx >= n_staged) continue;  // (two 3 x 3 kernels stage 240 words: one per thread)
            const int r = r_lo + idx / kMorphCols, c = idx % kMorphCols;
            const int gy = y0 - kMorphReach + r, gwx = wx0 + c;
            const bool row_in = gy >= 0 && gy < H;
            item[q] = idx < n_staged ? ((uint32_t)r | ((uint32_t)c << 8) |
                                        (((row_in ? 1u : 0u) | col_flags(c - 1, 2u, 16u) | col_flags(c, 4u, 32u) | col_flags(c + 1, 8u, 64u)) << 16))
                                     : 0xffffffffu;
            if (idx < n_staged && row_in && gwx >= 0 && gwx < WW) {
                v[q] = __ldcg(fb + (size_t)gy * WW + gwx);
                if (gwx == WW - 1) v[q] &= tail_mask;
            }
        }
        // staged rows inside the image and the staged region (for the vertical passes)
        const int v_lo = max(r_lo, kMorphReach - y0), v_hi = min(r_hi - 1, kMorphReach + (H - 1 - y0));
        __syncthreads();  // the previous tile's readers of s_a / s_b / s_w are done; s_next[it & 1] is published
        const unsigned int nxt = s_next[it & 1];
        const uint32_t t_nxt = s_tile[it & 1];
#pragma unroll
        for (int q = 0; q < 3; q++) {
            const int idx = tid + 256 * q;
            if (idx < n_staged) s_a[r_lo + idx / kMorphCols][idx % kMorphCols] = v[q];
        }
        const bool any = __syncthreads_or((v[0] | v[1] | v[2]) != 0);  // also: s_a complete
        if (any) {
            const uint32_t(*src)[kMorphCols] = s_a;
            uint32_t(*dst)[kMorphCols] = s_b;
            auto flip = [&]() {
                __syncthreads();
                const uint32_t(*tsrc)[kMorphCols] = dst;
                dst = const_cast<uint32_t(*)[kMorphCols]>(src);
                src = tsrc;
            };
            // open = erode, dilate; close = dilate, erode.  The two dilations in the middle are one dilation with the
            // summed window (a maximum over in-image pixels of a maximum over in-image pixels, and the image is convex).
            if (open_k > 0) {
                morph_pass_h<false>(src, dst, lo_o, hi_o, item, tail_mask), flip();
                morph_pass_v<false>(src, dst, lo_o, hi_o, item, tail_mask, v_lo, v_hi), flip();
            }
            morph_pass_h<true>(src, dst, lo_o + lo_c, hi_o + hi_c, item, tail_mask), flip();
            morph_pass_v<true>(src, dst, lo_o + lo_c, hi_o + hi_c, item, tail_mask, v_lo, v_hi), flip();
            if (close_k > 0) {
                morph_pass_h<false>(src, dst, lo_c, hi_c, item, tail_mask), flip();
                morph_pass_v<false>(src, dst, lo_c, hi_c, item, tail_mask, v_lo, v_hi), flip();
            }
            if (tid < 128) {
                const int r = tid >> 2, wq = tid & 3;
                const int gwx = 4 * tx + wq;
                uint32_t w = src[kMorphReach + r][1 + wq];
                if (y0 + r >= H || gwx >= WW) w = 0;
                else if (gwx == WW - 1) w &= tail_mask;
                s_w[r][wq] = w;
            }
            __syncthreads();
        }
        // final bit-mask words + occupancy records (a second set of records and a second bit plane: other CTAs still
        // read K1's), then the mask bytes and the label plane of the tile
        uint32_t word = 0;
        if (tid < 128) {
            const int r = tid >> 2, wq = tid & 3;
            if (any) word = s_w[r][wq];
            if (y0 + r < H && 4 * tx + wq < WW) bits_out[(f * H + y0 + r) * (size_t)WW + 4 * tx + wq] = word;
            const uint32_t bal = __ballot_sync(0xffffffffu, word != 0);
            if (wq == 0 && y0 + r < H)
                rowflags_out[f * b.rf_stride + rowflag_index(y0 + r, tx, b.tiles_x)] = (uint8_t)((bal >> (tid & 31)) & 0xfu);
        }
        const bool any_out = __syncthreads_or(word != 0);
        expand_tile(b, f, tx, ty, s_w, any_out, tid);
        cur = nxt;
        t = t_nxt;
    }
    __syncthreads();  // every thread of the CTA is past its stores
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(ctrl + 2, 1u) == gridDim.x - 1) {
            ctrl[0] = 0, ctrl[1] = 0, ctrl[2] = 0;
            if (chain) {  // "tiles done": the per-frame CCL kernel of the batch may read the planes
                __threadfence();
                atomicAdd(chain + 2, 1u);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// A7: cv2.GaussianBlur for CV_8U: rows in 8.8 fixed point, columns in 16.16, BORDER_REFLECT_101.
// ---------------------------------------------------------------------------------------------------------
struct GaussKernel {
    uint16_t k[32];
    int ksize;
};

__device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

__global__ void __launch_bounds__(256) k_gauss_rows(const uint8_t *src, int n, int h, int w, GaussKernel g,
                                                    uint16_t *rows) {
    const size_t total = (size_t)n * h * w;
    const int r = g.ksize / 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % w);
        const uint8_t *rowp = src + (i - x);
        uint32_t s = 0;
        if (x >= r && x + r < w) {
            for (int t = -r; t <= r; t++) s += (uint32_t)g.k[t + r] * rowp[x + t];
        } else {
            for (int t = -r; t <= r; t++) s += (uint32_t)g.k[t + r] * rowp[reflect101(x + t, w)];
        }
        rows[i] = (uint16_t)min(s, 65535u);
    }
}

__global__ void __launch_bounds__(256) k_gauss_cols(const uint16_t *rows, int n, int h, int w, GaussKernel g,
                                                    uint8_t *dst) {
    const size_t total = (size_t)n * h * w;
    const int r = g.ksize / 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % w);
        const size_t t0 = i / w;
        const int y = (int)(t0 % h);
        const uint16_t *colp = rows + (t0 - y) * w + x;
        uint32_t s = 0;
        for (int t = -r; t <= r; t++) s += (uint32_t)g.k[t + r] * colp[(size_t)reflect101(y + t, h) * w];
        const uint32_t v = (s + 32768u) >> 16;
        dst[i] = (uint8_t)min(v, 255u);
    }
}

// ---------------------------------------------------------------------------------------------------------
// processing.rs utilities
// ---------------------------------------------------------------------------------------------------------
// processing.rs:66-94: (2r+1)^2 box mean per channel, interior only, floor division.
__global__ void __launch_bounds__(256) k_box_blur_generic(const uint8_t *src, int h, int w, int nch, int radius,
                                                          uint8_t *dst) {
    const size_t total = (size_t)h * w * nch;
    const uint32_t cnt = (2 * radius + 1) * (2 * radius + 1);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int ch = (int)(i % nch);
        const size_t t = i / nch;
        const int x = (int)(t % w), y = (int)(t / w);
        uint32_t v = src[i];
        if (y >= radius && y < h - radius && x >= radius && x < w - radius) {
            uint32_t s = 0;
            for (int dy = -radius; dy <= radius; dy++)
                for (int dx = -radius; dx <= radius; dx++) s += src[((size_t)(y + dy) * w + (x + dx)) * nch + ch];
            v = s / cnt;
        }
        dst[i] = (uint8_t)v;
    }
}

// processing.rs:165-178: global threshold.
__global__ void __launch_bounds__(256) k_threshold_global(const uint8_t *src, size_t total, int thr, int inverse,
                                                          uint8_t *dst) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int v = src[i];
        dst[i] = (inverse ? (v < thr) : (v > thr)) ? 255 : 0;
    }
}

// find_contours' foreground predicate `> 127` (detection.rs:64) -> bit-packed mask. One warp per word.
__global__ void __launch_bounds__(256) k_bits_from_gt127(const uint8_t *src, int n, int h, int w, int ww,
                                                         uint32_t *bits) {
    const size_t words_per_frame = (size_t)h * ww;
    const size_t total = words_per_frame * n;
    const int lane = threadIdx.x & 31;
    const size_t warp0 = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t i = warp0; i < total; i += nwarps) {
        const size_t f = i / words_per_frame;
        const int wi = (int)(i - f * words_per_frame);
        const int y = wi / ww, wx = wi - y * ww;
        const int x = wx * 32 + lane;
        const bool fg = x < w && src[(f * h + y) * (size_t)w + x] > 127;
        const uint32_t word = __ballot_sync(0xffffffffu, fg);
        if (lane == 0) bits[i] = word;
    }
}

// processing.rs:371-401 / 237-246: replicate the mask into 3 channels, then 7-pixel crosses (0,0,255).
__global__ void __launch_bounds__(256) k_replicate3(const uint8_t *mask, size_t total, uint8_t *out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint8_t v = mask[i];
        out[3 * i] = v;
        out[3 * i + 1] = v;
        out[3 * i + 2] = v;
    }
}

__global__ void __launch_bounds__(256) k_crosses(const hv_center *centers, int n_centers, int h, int w,
                                                 uint8_t *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_centers * 14) return;
    const hv_center c = centers[i / 14];
    const int j = i % 14;
    int y = c.y, x = c.x;
    if (j < 7)
        y += j - 3;
    else
        x += j - 10;
    if (y < 0 || y >= h || x < 0 || x >= w) return;
    uint8_t *q = out + ((size_t)y * w + x) * 3;
    q[0] = 0;
    q[1] = 0;
    q[2] = 255;
}

// ---------------------------------------------------------------------------------------------------------
// N4: overlays (crosses, boxes, markers) drawn in list order.  Three passes over the items' pixels: claim (the highest
// item index that touches a pixel owns it -- what sequential drawing leaves behind), paint, release the claims.
// ---------------------------------------------------------------------------------------------------------
// footprint of cv2.circle(img, c, 10, color, 2) around its centre, offsets -11..11, bit dx + 11 of row dy + 11
// (opencv-python 4.13; tests/test_result_side.py compares with committed cv2 output)
__constant__ uint32_t kMarkerRows[23] = {0x7f00, 0x1ff80, 0x7fff0, 0xf81f8, 0x1f007c, 0x1e003c, 0x3c001e, 0x38000e,
                                         0x78000f, 0x700007, 0x700007, 0x700007, 0x700007, 0x700007, 0x700007, 0x38000e,
                                         0x3c001e, 0x1e003c, 0x1f007c, 0xf81f8, 0x7fff0, 0x1ff80, 0x7f00};

// pixel k of an item, or false when k is beyond the item / not part of it
__device__ __forceinline__ bool overlay_pixel(const hv_overlay &it, int k, int *py, int *px) {
    if (it.kind == HV_OVERLAY_CROSS) {  // processing.rs:371-401: rows y-3..y+3 at x, then columns x-3..x+3 at y
        if (k >= 14) return false;
        *py = k < 7 ? it.y + k - 3 : it.y;
        *px = k < 7 ? it.x : it.x + k - 10;
        return true;
    }
    if (it.kind == HV_OVERLAY_MARKER) {
        if (k >= 23 * 23) return false;
        const int r = k / 23, c = k - 23 * r;
        if (!((kMarkerRows[r] >> c) & 1u)) return false;
        *py = it.y + r - 11;
        *px = it.x + c - 11;
        return true;
    }
    const int ya = min(it.y, it.y1), yb = max(it.y, it.y1), xa = min(it.x, it.x1), xb = max(it.x, it.x1);
    const int bw = xb - xa + 1, bh = yb - ya + 1;
    if (k < bw) return *py = ya, *px = xa + k, true;
    k -= bw;
    if (k < bw) return *py = yb, *px = xa + k, true;
    k -= bw;
    if (k < bh) return *py = ya + k, *px = xa, true;
    k -= bh;
    if (k < bh) return *py = ya + k, *px = xb, true;
    return false;
}

__device__ __forceinline__ int overlay_extent(const hv_overlay &it) {
    if (it.kind == HV_OVERLAY_CROSS) return 14;
    if (it.kind == HV_OVERLAY_MARKER) return 23 * 23;
    const long long e = 2ll * (llabs((long long)it.y1 - it.y) + 1) + 2ll * (llabs((long long)it.x1 - it.x) + 1);
    return (int)min(e, 2000000000ll);
}

template <int PASS>
__global__ void __launch_bounds__(128) k_overlays(const hv_overlay *items, int n, int h, int w, uint8_t *img, unsigned int *owner) {
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const hv_overlay it = items[i];
        const int ext = overlay_extent(it);
        for (int k = threadIdx.x; k < ext; k += blockDim.x) {
            int y, x;
            if (!overlay_pixel(it, k, &y, &x) || y < 0 || y >= h || x < 0 || x >= w) continue;
            const size_t p = (size_t)y * w + x;
            if (PASS == 0) {
                atomicMax(owner + p, (unsigned int)i + 1u);
            } else if (PASS == 1) {
                if (owner[p] == (unsigned int)i + 1u) {
                    img[3 * p] = it.color[0];
                    img[3 * p + 1] = it.color[1];
                    img[3 * p + 2] = it.color[2];
                }
            } else {
                owner[p] = 0u;
            }
        }
    }
}

int grid_for(size_t work_items, int per_block) {
    size_t g = (work_items + per_block - 1) / per_block;
    const size_t cap = 148 * 32;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace

cudaError_t launch_morph(const BatchView &b, int open_k, int close_k, int *n_launches, cudaStream_t s) {
    // K1 wrote the pre-morphology mask into b.bits; every erode/dilate is a row pass into bits_tmp and a column pass back
    // into bits, so the result lands in b.bits again.
    const size_t total = (size_t)b.n * b.h * b.ww;
    const int grid = grid_for(total, 256);
    int launches = 0;
    auto step = [&](bool dilate, int k) {
        if (dilate) {
            k_morph_h<true><<<grid, 256, 0, s>>>(b.bits, b.bits_tmp, b.n, b.h, b.ww, b.w, k);
            k_morph_v<true><<<grid, 256, 0, s>>>(b.bits_tmp, b.bits, b.n, b.h, b.ww, k);
        } else {
            k_morph_h<false><<<grid, 256, 0, s>>>(b.bits, b.bits_tmp, b.n, b.h, b.ww, b.w, k);
            k_morph_v<false><<<grid, 256, 0, s>>>(b.bits_tmp, b.bits, b.n, b.h, b.ww, k);
        }
        launches += 2;
    };
    if (open_k > 0) {
        step(false, open_k);
        step(true, open_k);
    }
    if (close_k > 0) {
        step(true, close_k);
        step(false, close_k);
    }
    if (n_launches) *n_launches = launches;
    return cudaGetLastError();
}

// open followed by close in one kernel; the final bit-mask goes to bits_out (must not be b.bits)
bool morph_expand_supported(int open_k, int close_k) {
    auto reach = [](int k) { return k > 0 ? 2 * std::max(k / 2, k - 1 - k / 2) : 0; };
    return reach(open_k) + reach(close_k) <= kMorphReach;
}

// persistent grids: exactly one wave (a second, partial wave would run at a fraction of the occupancy for as long again)
template <typename K>
int one_wave_grid(K kernel, int block) {
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    return sms * per_sm;
}

cudaError_t launch_morph_expand(const BatchView &b, int open_k, int close_k, uint32_t *bits_out, uint8_t *rowflags_out,
                                uint32_t *tile_list, unsigned int *ctrl, unsigned int *chain, unsigned int scan_expected,
                                int num_sms, bool pdl, cudaStream_t s) {
    const size_t tiles = (size_t)b.n * ((b.h + 31) / 32) * b.tiles_x;
    auto reach1 = [](int k) { return k > 0 ? 2 * std::max(k / 2, k - 1 - k / 2) : 0; };
    const int reach = std::min(kMorphReach, reach1(open_k) + reach1(close_k));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)((tiles + 255) / 256));
    cfg.blockDim = dim3(256);
    cfg.stream = s;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_morph_scan, b, reach, rowflags_out, tile_list, ctrl, chain);
    if (e != cudaSuccess) return e;
    // counter chain: one CTA per SM, so that the whole grid is resident next to K1 and the CCL kernels at once (it has to
    // be, to release the kernel behind it); otherwise one full wave
    static const int wave = one_wave_grid(k_morph_tiles, 256);
    const int chain_ctas = tunables().morph_tiles_per_sm;
    cfg.gridDim = dim3((unsigned)std::min<size_t>(tiles, (size_t)(chain ? std::min(num_sms * chain_ctas, wave) : wave)));
    cfg.numAttrs = 1;  // behind the scan, whatever preceded that
    return cudaLaunchKernelEx(&cfg, k_morph_tiles, b, open_k, close_k, bits_out, rowflags_out, (const uint32_t *)tile_list, ctrl,
                              chain, scan_expected);
}

cudaError_t launch_expand_bits(const BatchView &b, cudaStream_t s) {
    const size_t tiles = (size_t)b.n * ((b.h + 31) / 32) * b.tiles_x;
    static const int wave = one_wave_grid(k_expand_bits, 256);
    k_expand_bits<<<(int)std::min<size_t>(tiles, (size_t)wave), 256, 0, s>>>(b);
    return cudaGetLastError();
}

cudaError_t launch_gaussian_blur(const uint8_t *src, int n, int h, int w, const uint16_t *k_q8_host, int ksize,
                                 uint8_t *dst, uint16_t *tmp_rows, cudaStream_t s) {
    if (ksize < 1 || ksize > 31 || !(ksize & 1)) return cudaErrorInvalidValue;
    GaussKernel g;
    for (int i = 0; i < 32; i++) g.k[i] = i < ksize ? k_q8_host[i] : 0;
    g.ksize = ksize;
    const size_t total = (size_t)n * h * w;
    k_gauss_rows<<<grid_for(total, 256), 256, 0, s>>>(src, n, h, w, g, tmp_rows);
    k_gauss_cols<<<grid_for(total, 256), 256, 0, s>>>(tmp_rows, n, h, w, g, dst);
    return cudaGetLastError();
}

cudaError_t launch_box_blur_generic(const uint8_t *src, int h, int w, int nch, int radius, uint8_t *dst,
                                    cudaStream_t s) {
    k_box_blur_generic<<<grid_for((size_t)h * w * nch, 256), 256, 0, s>>>(src, h, w, nch, radius, dst);
    return cudaGetLastError();
}

cudaError_t launch_threshold_generic(const uint8_t *src, int h, int w, int adaptive, int c_or_thr, int inverse,
                                     uint8_t *dst, cudaStream_t s) {
    if (adaptive) return cudaErrorInvalidValue;  // adaptive goes through launch_preprocess<RB=0>
    k_threshold_global<<<grid_for((size_t)h * w, 256), 256, 0, s>>>(src, (size_t)h * w, c_or_thr, inverse, dst);
    return cudaGetLastError();
}

cudaError_t launch_bits_from_gt127(const uint8_t *src, int n, int h, int w, int ww, uint32_t *bits,
                                   cudaStream_t s) {
    k_bits_from_gt127<<<grid_for((size_t)n * h * ww * 32, 256), 256, 0, s>>>(src, n, h, w, ww, bits);
    return cudaGetLastError();
}

cudaError_t launch_visualise(const uint8_t *mask, int h, int w, const hv_center *d_centers, int n_centers,
                             uint8_t *out_hw3, cudaStream_t s) {
    k_replicate3<<<grid_for((size_t)h * w, 256), 256, 0, s>>>(mask, (size_t)h * w, out_hw3);
    if (n_centers > 0) k_crosses<<<(n_centers * 14 + 255) / 256, 256, 0, s>>>(d_centers, n_centers, h, w, out_hw3);
    return cudaGetLastError();
}

}  // namespace hv

namespace hv {
cudaError_t launch_overlays(const hv_overlay *d_items, int n, int h, int w, uint8_t *d_img, unsigned int *d_owner, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    const int grid = n < 148 * 8 ? n : 148 * 8;
    k_overlays<0><<<grid, 128, 0, s>>>(d_items, n, h, w, d_img, d_owner);
    k_overlays<1><<<grid, 128, 0, s>>>(d_items, n, h, w, d_img, d_owner);
    k_overlays<2><<<grid, 128, 0, s>>>(d_items, n, h, w, d_img, d_owner);
    return cudaGetLastError();
}
}  // namespace hv
