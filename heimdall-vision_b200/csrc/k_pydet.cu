// k_pydet.cu -- next-row N3: the stages of the reference's PYTHON detector (heimdall/detectors/contamination_detector.py:
// 58-90, the fallback behind heimdall/rust_bridge.py:139-161) with OpenCV's arithmetic, on the GPU:
//   cv2.cvtColor(BGR2GRAY)            fixed point, 15 fractional bits: (B*3735 + G*19235 + R*9798 + 16384) >> 15
//   cv2.GaussianBlur(5,5,0)           k_gauss_rows / k_gauss_cols of k_stage.cu (8.8 fixed point, reflect-101)
//   cv2.adaptiveThreshold(GAUSSIAN_C, THRESH_BINARY_INV, blockSize, C)
//                                      mean = u8(round-half-even(float32 Gaussian of the u8 image, BORDER_REPLICATE)),
//                                      mask = 255 if src - mean <= -floor(C) else 0.  The float32 filter follows OpenCV's
//                                      separable float path as dispatched on AVX2/FMA hosts (opencv-python 4.13): rows
//                                      acc = fma(src[t], k[t], acc) for t = 0..n-1 from 0; columns s = k[c] * row[c], then
//                                      s = fma(row[c+j] + row[c-j], k[c+j], s) for j = 1..n/2.  Another summation order can
//                                      differ where the float mean lands within an ulp of x.5 (about 3 pixels in 10^7).
//   cv2.morphologyEx(OPEN / CLOSE)    k_morph_h / k_morph_v of k_stage.cu on the bit-packed mask
//   8-connected components            the global-memory CCL kernels of k_ccl.cu with BatchView::conn8 set
// Not on the hot path: plain grid-stride kernels.
#include "hv_common.cuh"

namespace hv {

namespace {

__global__ void __launch_bounds__(256) k_gray_bgr_cv(const uint8_t *img, size_t px, uint8_t *gray) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < px; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t b = img[3 * i], g = img[3 * i + 1], r = img[3 * i + 2];
        gray[i] = (uint8_t)((b * 3735u + g * 19235u + r * 9798u + 16384u) >> 15);
    }
}

struct FloatKernel {
    float k[31];
    int n;
};

__global__ void __launch_bounds__(256) k_adapt_rows(const uint8_t *src, int h, int w, FloatKernel fk, float *rows) {
    const size_t total = (size_t)h * w;
    const int r = fk.n / 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % w);
        const uint8_t *rowp = src + (i - x);
        float acc = 0.f;
        for (int t = 0; t < fk.n; t++) {
            const int xx = min(max(x + t - r, 0), w - 1);  // BORDER_REPLICATE
            acc = __fmaf_rn((float)rowp[xx], fk.k[t], acc);
        }
        rows[i] = acc;
    }
}

__global__ void __launch_bounds__(256) k_adapt_cols_threshold(const uint8_t *src, const float *rows, int h, int w, FloatKernel fk,
                                                              int idelta, uint8_t *mask) {
    const size_t total = (size_t)h * w;
    const int r = fk.n / 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % w), y = (int)(i / w);
        float s = __fmul_rn(fk.k[r], rows[i]);
        for (int j = 1; j <= r; j++) {
            const float a = rows[(size_t)min(y + j, h - 1) * w + x], b = rows[(size_t)max(y - j, 0) * w + x];
            s = __fmaf_rn(__fadd_rn(a, b), fk.k[r + j], s);
        }
        int mean = __float2int_rn(s);  // cvRound: round half to even, then saturate_cast<uchar>
        mean = min(max(mean, 0), 255);
        mask[i] = ((int)src[i] - mean <= -idelta) ? 255 : 0;
    }
}

// ---- contour quantities of the external components ------------------------------------------------------------------------
// cv2.findContours(RETR_EXTERNAL) traces the outer border of every 8-connected component that does not lie inside a hole of
// another one, and cv2.drawContours(..., -1) of such a contour paints the component plus everything it encloses.  Both are
// functions of two label planes -- the 8-connected components of the mask and the 4-connected components of its
// complement -- and of the enclosure relation between them, resolved on the host into `root` tables: root8[l] / root4[l] =
// the external component that component l (foreground / background) belongs to or is enclosed by, 0 = none.

__global__ void __launch_bounds__(256) k_invert_bits(const uint32_t *in, uint32_t *out, int h, int ww, int w) {
    const size_t total = (size_t)h * ww;
    const uint32_t tail = (w & 31) ? ((1u << (w & 31)) - 1u) : 0xffffffffu;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t v = ~in[i];
        if ((int)(i % ww) == ww - 1) v &= tail;
        out[i] = v;
    }
}

// raster-first pixel (linear index) of every label of a label plane
__global__ void __launch_bounds__(256) k_first_pixel(const int32_t *labels, size_t px, uint32_t *first) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < px; i += (size_t)gridDim.x * blockDim.x) {
        const int32_t l = labels[i];
        if (l > 0) atomicMin(first + (l - 1), (uint32_t)i);
    }
}

// label of the OTHER plane at the pixel above a component's first pixel (0 in the first row): the background region above a
// foreground component decides whether that component is nested, the foreground component above a hole is its owner
__global__ void __launch_bounds__(256) k_label_above(const uint32_t *first, int n, int w, const int32_t *other, int32_t *above) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = first[i];
    above[i] = p >= (uint32_t)w ? other[p - w] : 0;
}

struct RegionSums {  // per external component, over its bounding box: inside / outside the filled contour
    unsigned long long cnt_in, cnt_out, gray_in, gray_out, ch_in[3], ch_out[3];
};

__device__ __forceinline__ bool in_region(int c1, const int32_t *l8, const int32_t *l4, const int32_t *root8, const int32_t *root4,
                                          int h, int w, int y, int x) {
    if (y < 0 || y >= h || x < 0 || x >= w) return false;
    const size_t p = (size_t)y * w + x;
    const int32_t a = l8[p];
    if (a > 0) return root8[a - 1] == c1;
    const int32_t b = l4[p];
    return b > 0 && root4[b - 1] == c1;
}

// one CTA per external component
__global__ void __launch_bounds__(256) k_region_sums(const int32_t *ext, int n_ext, const hv_blob *comps8, const int32_t *l8,
                                                     const int32_t *l4, const int32_t *root8, const int32_t *root4, int h, int w,
                                                     const uint8_t *gray, const uint8_t *bgr, RegionSums *out) {
    __shared__ unsigned long long s[14];
    const int e = blockIdx.x;
    if (e >= n_ext) return;
    const int c1 = ext[e];  // 1-based label of the component
    const hv_blob q = comps8[c1 - 1];
    const int bw = (int)(q.xmax - q.xmin + 1), bh = (int)(q.ymax - q.ymin + 1);
    if (threadIdx.x < 14) s[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long acc[14] = {0};
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) {
        const int y = (int)q.ymin + i / bw, x = (int)q.xmin + i % bw;
        const bool in = in_region(c1, l8, l4, root8, root4, h, w, y, x);
        const size_t p = (size_t)y * w + x;
        const int o = in ? 0 : 1;
        acc[o] += 1;
        acc[2 + o] += gray[p];
        if (bgr) {
            acc[4 + o * 3] += bgr[3 * p];
            acc[5 + o * 3] += bgr[3 * p + 1];
            acc[6 + o * 3] += bgr[3 * p + 2];
        }
    }
#pragma unroll
    for (int k = 0; k < 10; k++) {
        unsigned long long v = acc[k];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s[k], v);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        RegionSums r;
        r.cnt_in = s[0], r.cnt_out = s[1], r.gray_in = s[2], r.gray_out = s[3];
        for (int k = 0; k < 3; k++) r.ch_in[k] = s[4 + k], r.ch_out[k] = s[7 + k];
        out[e] = r;
    }
}

struct TraceOut {
    double a00, a10, a01;   // cv2 contourMoments accumulators over the closed chain
    uint32_t chain_len;
    uint32_t reserved;
};

// Moore-neighbour tracing of the outer border of the filled region, one thread per external component, starting at the
// component's raster-first pixel with the (background) pixel to its west as backtrack cell, scanning clockwise.  This is the
// chain cv2.findContours(CHAIN_APPROX_NONE) returns (same length, same polygon); the accumulators are cv2's contourMoments
// sums, exact in double (integer terms far below 2^53), so area and first moments are bit-identical with cv2.contourArea /
// cv2.moments of the CHAIN_APPROX_SIMPLE polygon (collinear points drop out of the sums).
__global__ void __launch_bounds__(64) k_trace_contours(const int32_t *ext, int n_ext, const uint32_t *first8, const int32_t *l8,
                                                       const int32_t *l4, const int32_t *root8, const int32_t *root4, int h, int w,
                                                       TraceOut *out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_ext) return;
    const int c1 = ext[e];
    const int dxs[8] = {-1, -1, 0, 1, 1, 1, 0, -1}, dys[8] = {0, -1, -1, -1, 0, 1, 1, 1};  // W NW N NE E SE S SW (clockwise)
    const uint32_t p0 = first8[c1 - 1];
    const int sx = (int)(p0 % (uint32_t)w), sy = (int)(p0 / (uint32_t)w);
    TraceOut r;
    r.a00 = r.a10 = r.a01 = 0.0;
    r.chain_len = 1;
    r.reserved = 0;
    bool any = false;
    for (int d = 0; d < 8; d++) any |= in_region(c1, l8, l4, root8, root4, h, w, sy + dys[d], sx + dxs[d]);
    if (any) {
        int cx = sx, cy = sy, bdir = 0;
        int fx = 0, fy = 0, fk = -1;  // the first move: the trace is closed when it is about to be repeated
        uint32_t len = 0;
        const uint32_t limit = 8u * (uint32_t)h * (uint32_t)w + 16u;
        while (len < limit) {
            int k = -1;
            for (int i = 1; i <= 8; i++) {
                const int d = (bdir + i) & 7;
                if (in_region(c1, l8, l4, root8, root4, h, w, cy + dys[d], cx + dxs[d])) {
                    k = d;
                    break;
                }
            }
            if (fk < 0)
                fx = cx, fy = cy, fk = k;
            else if (cx == fx && cy == fy && k == fk)
                break;
            const int nx = cx + dxs[k], ny = cy + dys[k];
            // contourMoments term of the edge (cx, cy) -> (nx, ny)
            const double dxy = (double)cx * (double)ny - (double)nx * (double)cy;
            r.a00 += dxy;
            r.a10 += dxy * (double)(cx + nx);
            r.a01 += dxy * (double)(cy + ny);
            // new backtrack cell: the cell examined just before k, seen from the new pixel
            const int pd = (k + 7) & 7;
            const int bx = cx + dxs[pd] - nx, by = cy + dys[pd] - ny;
            for (int d = 0; d < 8; d++)
                if (dxs[d] == bx && dys[d] == by) bdir = d;
            cx = nx, cy = ny;
            len++;
        }
        r.chain_len = len;
    }
    out[e] = r;
}

int grid_for(size_t items) {
    size_t g = (items + 255) / 256;
    return (int)(g < 1 ? 1 : (g > 148 * 32 ? 148 * 32 : g));
}

}  // namespace

cudaError_t launch_gray_bgr_cv(const uint8_t *d_img, int h, int w, uint8_t *d_gray, cudaStream_t s) {
    k_gray_bgr_cv<<<grid_for((size_t)h * w), 256, 0, s>>>(d_img, (size_t)h * w, d_gray);
    return cudaGetLastError();
}

cudaError_t launch_adaptive_gaussian(const uint8_t *d_src, int h, int w, const float *k_host, int ksize, int idelta,
                                     float *d_rows, uint8_t *d_mask, cudaStream_t s) {
    if (ksize < 3 || ksize > 31 || !(ksize & 1)) return cudaErrorInvalidValue;
    FloatKernel fk;
    for (int i = 0; i < 31; i++) fk.k[i] = i < ksize ? k_host[i] : 0.f;
    fk.n = ksize;
    k_adapt_rows<<<grid_for((size_t)h * w), 256, 0, s>>>(d_src, h, w, fk, d_rows);
    k_adapt_cols_threshold<<<grid_for((size_t)h * w), 256, 0, s>>>(d_src, d_rows, h, w, fk, idelta, d_mask);
    return cudaGetLastError();
}

}  // namespace hv

namespace hv {
cudaError_t launch_invert_bits(const uint32_t *in, uint32_t *out, int h, int ww, int w, cudaStream_t s) {
    k_invert_bits<<<grid_for((size_t)h * ww), 256, 0, s>>>(in, out, h, ww, w);
    return cudaGetLastError();
}
cudaError_t launch_first_pixel(const int32_t *labels, int h, int w, uint32_t *first, cudaStream_t s) {
    k_first_pixel<<<grid_for((size_t)h * w), 256, 0, s>>>(labels, (size_t)h * w, first);
    return cudaGetLastError();
}
cudaError_t launch_label_above(const uint32_t *first, int n, int w, const int32_t *other, int32_t *above, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    k_label_above<<<(n + 255) / 256, 256, 0, s>>>(first, n, w, other, above);
    return cudaGetLastError();
}
cudaError_t launch_region_sums(const int32_t *ext, int n_ext, const hv_blob *comps8, const int32_t *l8, const int32_t *l4,
                               const int32_t *root8, const int32_t *root4, int h, int w, const uint8_t *gray, const uint8_t *bgr,
                               void *out, cudaStream_t s) {
    if (n_ext <= 0) return cudaSuccess;
    k_region_sums<<<n_ext, 256, 0, s>>>(ext, n_ext, comps8, l8, l4, root8, root4, h, w, gray, bgr, static_cast<RegionSums *>(out));
    return cudaGetLastError();
}
cudaError_t launch_trace_contours(const int32_t *ext, int n_ext, const uint32_t *first8, const int32_t *l8, const int32_t *l4,
                                  const int32_t *root8, const int32_t *root4, int h, int w, void *out, cudaStream_t s) {
    if (n_ext <= 0) return cudaSuccess;
    k_trace_contours<<<(n_ext + 63) / 64, 64, 0, s>>>(ext, n_ext, first8, l8, l4, root8, root4, h, w, static_cast<TraceOut *>(out));
    return cudaGetLastError();
}
}  // namespace hv
