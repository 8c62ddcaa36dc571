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
