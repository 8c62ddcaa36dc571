// k_stage.cu -- the stages around the fused hot kernel: morphology on the bit-packed mask (A8), the OpenCV-exact
// Gaussian blur (A7), and the single-frame utilities behind heimdall_core.processing.* / process_image
// (rust/heimdall-core/src/processing.rs).  None of these is on the default (Rust-exact) detect path.
#include <algorithm>

#include "hv_common.cuh"

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

// Mask bytes, label-plane initialisation (p + 1 at the first pixel of every word-run, see k_preprocess.cu) and the
// occupancy records from a bit-packed mask, one 128 x 32 tile per CTA-iteration with the store pattern of K1: all-zero
// tiles (most of an inspection frame) are nothing but 128-bit zero stores.  Used after morphology.
__global__ void __launch_bounds__(256) k_expand_bits(BatchView b) {
    __shared__ uint32_t s_w[32][4];
    const int tid = threadIdx.x;
    const int H = b.h, W = b.w;
    const int tiles_y = (H + 31) / 32;
    const size_t per_frame = (size_t)tiles_y * b.tiles_x, total = per_frame * b.n;
    const bool vec = (W & 15) == 0;
    for (size_t t = blockIdx.x; t < total; t += gridDim.x) {
        const size_t f = t / per_frame;
        const int j = (int)(t - f * per_frame);
        const int ty = j / b.tiles_x, tx = j - ty * b.tiles_x;
        const int x0 = tx * 128, y0 = ty * 32;
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
        const size_t pix0 = (f * H + y0) * (size_t)W + x0;
        if (vec && x0 + 128 <= W) {
            // one thread per 16 pixels, 128-bit stores throughout
            const int r = tid >> 3, c16 = (tid & 7) * 16;
            if (y0 + r < H) {
                const uint32_t m16 = any ? (s_w[r][c16 >> 5] >> (c16 & 31)) & 0xffffu : 0u;
                const size_t o = pix0 + (size_t)r * W + c16;
                uint4 mv;
                uint32_t *mp = &mv.x;
#pragma unroll
                for (int q = 0; q < 4; q++) {  // 4 bits -> 4 bytes of 0x00 / 0xff
                    const uint32_t nib = (m16 >> (4 * q)) & 0xfu;
                    mp[q] = (((nib * 0x00204081u) & 0x01010101u) * 0xffu);
                }
                *reinterpret_cast<uint4 *>(b.mask + o) = mv;
                int4 *dst = reinterpret_cast<int4 *>(b.labels + o);
                if (!m16) {
                    const int4 z = make_int4(0, 0, 0, 0);
                    dst[0] = z, dst[1] = z, dst[2] = z, dst[3] = z;
                } else {
                    const uint32_t prev = (c16 & 31) ? (s_w[r][c16 >> 5] >> ((c16 & 31) - 1)) & 1u : 0u;
                    const uint32_t starts = m16 & ~((m16 << 1) | prev);
                    const int base = (y0 + r) * W + x0 + c16 + 1;
                    int lab[16];
#pragma unroll
                    for (int q = 0; q < 16; q++) lab[q] = ((starts >> q) & 1u) ? base + q : 0;
#pragma unroll
                    for (int q = 0; q < 4; q++) dst[q] = make_int4(lab[4 * q], lab[4 * q + 1], lab[4 * q + 2], lab[4 * q + 3]);
                }
            }
        } else {
            for (int idx = tid; idx < 32 * 128; idx += 256) {
                const int r = idx >> 7, c = idx & 127;
                const int gy = y0 + r, gx = x0 + c;
                if (gy >= H || gx >= W) continue;
                const uint32_t m = s_w[r][c >> 5];
                const int bit = c & 31;
                const bool fg = (m >> bit) & 1u;
                const bool start = fg && (bit == 0 || !((m >> (bit - 1)) & 1u));
                b.mask[pix0 + (size_t)r * W + c] = fg ? 255 : 0;
                b.labels[pix0 + (size_t)r * W + c] = start ? (gy * W + gx + 1) : 0;
            }
        }
        __syncthreads();  // s_w is rewritten by the next tile
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

cudaError_t launch_expand_bits(const BatchView &b, cudaStream_t s) {
    const size_t tiles = (size_t)b.n * ((b.h + 31) / 32) * b.tiles_x;
    k_expand_bits<<<(int)std::min<size_t>(tiles, 148 * 8), 256, 0, s>>>(b);
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
