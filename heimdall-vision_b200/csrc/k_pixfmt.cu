// k_pixfmt.cu -- N1 frame feed: camera pixel formats -> detector input, on the device.
//
// The reference converts camera frames on the host: to_ndarray (rust/heimdall-camera/src/lib.rs:260-278) passes
// Mono8 / RGB8 / BGR8 / RGBA8 / BGRA8 bytes through unchanged, to_opencv_mat (lib.rs:203-257) names cv2.cvtColor with
// COLOR_Bayer{RG,GB,GR,BG}2RGB and COLOR_YUV2RGB_YUYV for the mosaic and YUV formats (in the reference those branches
// are unreachable: the `mat_type` match at lib.rs:207-216 returns ConversionError first; we implement what they name).
// Here the raw frame (1 B/px Bayer, 2 B/px YUYV) crosses PCIe and is converted by a kernel, either to interleaved RGB
// (hv_convert_frame) or straight to the f64 gray the detector's A1 stage would compute from that RGB (feed path:
// the RGB image is never materialised: 1-2 B/px read, 1 B/px written instead of 3 + 3 + 1).
//
// Bayer semantics (bit-exact with opencv-python 4.13.0, tests/golden/cv2_pixfmt.npz): bilinear interpolation on the
// interior (pair averages (a+b+1)>>1, cross / diagonal averages (a+b+c+d+2)>>2), the one-pixel border copies the
// nearest interior result, frames smaller than 3x3 give zeros.  YUYV: BT.601 limited range in 20-bit fixed point.
#include "hv_common.cuh"

namespace hv {

namespace {

// which neighbourhood average feeds output channel ch at a site of type t (0 centre, 1 horizontal pair, 2 vertical
// pair, 3 cross, 4 diagonal) and the site type at (y & 1, x & 1) per pattern (RG, GB, GR, BG), packed 2 bits each
__constant__ uint8_t c_site[4][3] = {{4, 3, 0}, {0, 3, 4}, {2, 0, 1}, {1, 0, 2}};
__constant__ uint8_t c_pat[4][2][2] = {{{0, 2}, {3, 1}}, {{3, 1}, {0, 2}}, {{2, 0}, {1, 3}}, {{1, 3}, {2, 0}}};

template <bool GRAY>
__global__ void __launch_bounds__(256) k_bayer(const uint8_t *__restrict__ src, int n, int h, int w, int pattern,
                                               uint8_t *__restrict__ dst) {
    const size_t total = (size_t)n * h * w;
    const bool tiny = h < 3 || w < 3;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % w);
        const size_t t = i / w;
        const int y = (int)(t % h);
        const size_t f = t / h;
        uint32_t r = 0, g = 0, bl = 0;
        if (!tiny) {
            const int yy = min(max(y, 1), h - 2), xx = min(max(x, 1), w - 2);
            const uint8_t *p = src + (f * h + yy) * (size_t)w + xx;
            const uint32_t a = __ldg(p - w - 1), b = __ldg(p - w), c = __ldg(p - w + 1);
            const uint32_t d = __ldg(p - 1), e = __ldg(p), g2 = __ldg(p + 1);
            const uint32_t h2 = __ldg(p + w - 1), i2 = __ldg(p + w), j = __ldg(p + w + 1);
            uint32_t v[5];
            v[0] = e;
            v[1] = (d + g2 + 1) >> 1;
            v[2] = (b + i2 + 1) >> 1;
            v[3] = (d + g2 + b + i2 + 2) >> 2;
            v[4] = (a + c + h2 + j + 2) >> 2;
            const uint8_t *s = c_site[c_pat[pattern][yy & 1][xx & 1]];
            r = v[s[0]], g = v[s[1]], bl = v[s[2]];
        }
        if (GRAY) {
            dst[i] = gray_f64(r, g, bl);
        } else {
            dst[3 * i] = (uint8_t)r, dst[3 * i + 1] = (uint8_t)g, dst[3 * i + 2] = (uint8_t)bl;
        }
    }
}

__device__ __forceinline__ uint32_t sat8(long long v) { return (uint32_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

// one thread per pixel pair (Y0 U Y1 V)
template <bool GRAY>
__global__ void __launch_bounds__(256) k_yuyv(const uint8_t *__restrict__ src, size_t pairs, uint8_t *__restrict__ dst) {
    const long long CY = 1220542, CUB = 2116026, CUG = -409993, CVG = -852492, CVR = 1673527, HALF = 1 << 19;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < pairs; i += (size_t)gridDim.x * blockDim.x) {
        const uchar4 q = __ldg(reinterpret_cast<const uchar4 *>(src) + i);
        const long long u = (long long)q.y - 128, v = (long long)q.w - 128;
        const long long ruv = HALF + CVR * v, guv = HALF + CVG * v + CUG * u, buv = HALF + CUB * u;
        const long long y0 = (long long)max((int)q.x - 16, 0) * CY, y1 = (long long)max((int)q.z - 16, 0) * CY;
        const uint32_t r0 = sat8((y0 + ruv) >> 20), g0 = sat8((y0 + guv) >> 20), b0 = sat8((y0 + buv) >> 20);
        const uint32_t r1 = sat8((y1 + ruv) >> 20), g1 = sat8((y1 + guv) >> 20), b1 = sat8((y1 + buv) >> 20);
        if (GRAY) {
            dst[2 * i] = gray_f64(r0, g0, b0);
            dst[2 * i + 1] = gray_f64(r1, g1, b1);
        } else {
            uint8_t *o = dst + 6 * i;
            o[0] = (uint8_t)r0, o[1] = (uint8_t)g0, o[2] = (uint8_t)b0, o[3] = (uint8_t)r1, o[4] = (uint8_t)g1, o[5] = (uint8_t)b1;
        }
    }
}

int grid_px(size_t items) {
    size_t g = (items + 255) / 256;
    const size_t cap = 148 * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

cudaError_t launch_bayer(const uint8_t *d_src, int n, int h, int w, int pattern, bool to_gray, uint8_t *d_dst,
                         cudaStream_t s) {
    if (pattern < 0 || pattern > 3) return cudaErrorInvalidValue;
    const int g = grid_px((size_t)n * h * w);
    if (to_gray)
        k_bayer<true><<<g, 256, 0, s>>>(d_src, n, h, w, pattern, d_dst);
    else
        k_bayer<false><<<g, 256, 0, s>>>(d_src, n, h, w, pattern, d_dst);
    return cudaGetLastError();
}

cudaError_t launch_yuyv(const uint8_t *d_src, int n, int h, int w, bool to_gray, uint8_t *d_dst, cudaStream_t s) {
    if (w & 1) return cudaErrorInvalidValue;
    const size_t pairs = (size_t)n * h * (w / 2);
    if (to_gray)
        k_yuyv<true><<<grid_px(pairs), 256, 0, s>>>(d_src, pairs, d_dst);
    else
        k_yuyv<false><<<grid_px(pairs), 256, 0, s>>>(d_src, pairs, d_dst);
    return cudaGetLastError();
}

}  // namespace hv
