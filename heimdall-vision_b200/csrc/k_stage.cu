// k_stage.cu -- the stages around the fused hot kernel: morphology on the bit-packed mask (A8), the OpenCV-exact
// Gaussian blur (A7), and the single-frame utilities behind heimdall_core.processing.* / process_image
// (rust/heimdall-core/src/processing.rs).  None of these is on the default (Rust-exact) detect path.
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

template <bool DILATE>
__global__ void __launch_bounds__(256) k_morph(const uint32_t *src, uint32_t *dst, int n, int h, int ww, int w,
                                               int k) {
    const int a = k / 2;             // anchor
    const int lo = -a, hi = k - 1 - a;  // window offsets [lo, hi]
    const uint32_t tail_mask = (w & 31) ? ((1u << (w & 31)) - 1u) : 0xffffffffu;
    const size_t words_per_frame = (size_t)h * ww;
    const size_t total = words_per_frame * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t f = i / words_per_frame;
        const int wi = (int)(i - f * words_per_frame);
        const int y = wi / ww, wx = wi - y * ww;
        const uint32_t *frame = src + f * words_per_frame;
        uint32_t acc = DILATE ? 0u : 0xffffffffu;
        for (int dy = lo; dy <= hi; dy++) {
            const int yy = y + dy;
            if (yy < 0 || yy >= h) continue;
            const uint32_t *rowp = frame + (size_t)yy * ww;
            const uint32_t L = load_word<DILATE>(rowp, wx - 1, ww, tail_mask);
            const uint32_t M = load_word<DILATE>(rowp, wx, ww, tail_mask);
            const uint32_t R = load_word<DILATE>(rowp, wx + 1, ww, tail_mask);
            uint32_t hacc = M;
            for (int dx = 1; dx <= hi; dx++) {  // pixel x+dx -> bit position shifts right
                const uint32_t s = __funnelshift_r(M, R, dx);
                hacc = DILATE ? (hacc | s) : (hacc & s);
            }
            for (int dx = 1; dx <= -lo; dx++) {  // pixel x-dx
                const uint32_t s = __funnelshift_l(L, M, dx);
                hacc = DILATE ? (hacc | s) : (hacc & s);
            }
            acc = DILATE ? (acc | hacc) : (acc & hacc);
        }
        if (wx == ww - 1) acc &= tail_mask;
        dst[i] = acc;
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
    // K1 wrote the pre-morphology mask into b.bits; every erode/dilate ping-pongs between bits and bits_tmp and the
    // chain always has an even number of steps, so the result lands in b.bits again.
    const size_t total = (size_t)b.n * b.h * b.ww;
    const int grid = grid_for(total, 256);
    uint32_t *cur = b.bits, *nxt = b.bits_tmp;
    int launches = 0;
    auto step = [&](bool dilate, int k) {
        if (dilate)
            k_morph<true><<<grid, 256, 0, s>>>(cur, nxt, b.n, b.h, b.ww, b.w, k);
        else
            k_morph<false><<<grid, 256, 0, s>>>(cur, nxt, b.n, b.h, b.ww, b.w, k);
        uint32_t *t = cur;
        cur = nxt;
        nxt = t;
        launches++;
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
