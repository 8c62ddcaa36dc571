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

template <int TW, int TH, int RB>
struct Tile {
    static constexpr int HALO = RB + kAdaptHalf;  // rows of gray needed above/below
    static constexpr int HX = 8;                  // column halo, padded to a multiple of 4 for aligned loads
    static constexpr int GW = TW + 2 * HX;
    static constexpr int GH = TH + 2 * HALO;
    static constexpr int BW = TW + 2 * kAdaptHalf;  // blur columns needed
    static constexpr int BWP = BW + 2;              // padded pitch
    static constexpr int BH = TH + 2 * kAdaptHalf;
};

template <int TW, int TH, int RB>
__global__ void __launch_bounds__(256) k_preprocess(BatchView b, PreprocessParams p, uint32_t *bits_out) {
    using T = Tile<TW, TH, RB>;
    static_assert(TW % 32 == 0, "tile width must cover whole bitmask words");
    __shared__ __align__(16) uint8_t s_g[T::GH][T::GW];
    __shared__ __align__(16) uint8_t s_bl[T::BH][T::BWP];
    __shared__ __align__(16) uint16_t s_h11[T::BH][TW];
    __shared__ __align__(16) uint8_t s_f[TH][TW];
    __shared__ uint32_t s_red[2][8];

    const int tid = threadIdx.x;
    const int f = blockIdx.z;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int H = b.h, W = b.w;
    const uint8_t *gray = b.gray + (size_t)f * b.gray_frame_stride;
    const size_t gpitch = b.gray_row_stride;
    const bool aligned_in = ((reinterpret_cast<uintptr_t>(gray) | gpitch) & 3) == 0;

    // ---- 1. stage the gray tile + halo in shared memory, tracking min/max of the in-image pixels -------------
    uint32_t mn = 0x00ff00ffu, mx = 0u;  // packed u16x2 running min / max
    for (int idx = tid; idx < T::GH * (T::GW / 4); idx += 256) {
        const int r = idx / (T::GW / 4), cw = idx - r * (T::GW / 4);
        const int gy = y0 - T::HALO + r, gx = x0 - T::HX + 4 * cw;
        uint32_t v = 0;
        if (gy >= 0 && gy < H) {
            const uint8_t *row = gray + (size_t)gy * gpitch;
            if (aligned_in && gx >= 0 && gx + 3 < W) {
                v = __ldg(reinterpret_cast<const uint32_t *>(row + gx));
                const uint32_t lo = prmt(v, 0, 0x4140), hi = prmt(v, 0, 0x4342);
                mn = __vminu2(__vminu2(mn, lo), hi);
                mx = __vmaxu2(__vmaxu2(mx, lo), hi);
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int x = gx + k;
                    if (x >= 0 && x < W) {
                        const uint32_t q = __ldg(row + x);
                        v |= q << (8 * k);
                        mn = __vminu2(mn, q | (q << 16));
                        mx = __vmaxu2(mx, q | (q << 16));
                    }
                }
            }
        }
        *reinterpret_cast<uint32_t *>(&s_g[r][4 * cw]) = v;
    }
    // block-wide min / max
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        mn = __vminu2(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = __vmaxu2(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((tid & 31) == 0) {
        s_red[0][tid >> 5] = mn;
        s_red[1][tid >> 5] = mx;
    }
    __syncthreads();
    mn = s_red[0][0];
    mx = s_red[1][0];
#pragma unroll
    for (int k = 1; k < 8; k++) {
        mn = __vminu2(mn, s_red[0][k]);
        mx = __vmaxu2(mx, s_red[1][k]);
    }
    const int vmin = min(mn & 0xffffu, mn >> 16), vmax = max(mx & 0xffffu, mx >> 16);
    const int cth = p.c_thresh;
    const bool flat = p.inverse && !p.write_blur && cth >= 0 && (vmax - vmin) <= cth;

    if (!flat) {
        // ---- 2. blur over the tile + 5-px ring (zero outside the image, pass-through outside the interior) ---
        for (int idx = tid; idx < T::BH * T::BW; idx += 256) {
            const int r = idx / T::BW, c = idx - r * T::BW;
            const int gy = y0 - kAdaptHalf + r, gx = x0 - kAdaptHalf + c;
            const int ry = r + RB, rx = c + (T::HX - kAdaptHalf);
            uint32_t v = 0;
            if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                if (RB > 0 && gy >= RB && gy < H - RB && gx >= RB && gx < W - RB) {
                    uint32_t s = 0;
#pragma unroll
                    for (int dy = -RB; dy <= RB; dy++)
#pragma unroll
                        for (int dx = -RB; dx <= RB; dx++) s += s_g[ry + dy][rx + dx];
                    v = (RB == 2) ? div25(s) : s / ((2 * RB + 1) * (2 * RB + 1));
                } else {
                    v = s_g[ry][rx];
                }
                if (p.write_blur && r >= kAdaptHalf && r < kAdaptHalf + TH && c >= kAdaptHalf &&
                    c < kAdaptHalf + TW)
                    b.blur[((size_t)f * H + gy) * W + gx] = (uint8_t)v;
            }
            s_bl[r][c] = (uint8_t)v;
        }
        __syncthreads();
        // ---- 3. horizontal 11-sums ------------------------------------------------------------------------------
        for (int idx = tid; idx < T::BH * TW; idx += 256) {
            const int r = idx / TW, c = idx - r * TW;
            uint32_t s = 0;
#pragma unroll
            for (int k = 0; k < 2 * kAdaptHalf + 1; k++) s += s_bl[r][c + k];
            s_h11[r][c] = (uint16_t)s;
        }
        __syncthreads();
        // ---- 4. vertical 11-sums + threshold test -------------------------------------------------------------
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
                const int px = s_bl[r + kAdaptHalf][c + kAdaptHalf];
                const bool t = p.inverse ? ((px + cth + 1) * cnt <= s) : !((px + cth) * cnt <= s);
                fg = t ? 255 : 0;
            }
            s_f[r][c] = fg;
        }
        __syncthreads();
    }

    // ---- 5. outputs: u8 mask, bit-packed mask, label plane ------------------------------------------------------
    const bool vec = (W & 3) == 0;
    uint8_t *mask = b.mask + (size_t)f * H * W;
    int32_t *labels = b.labels + (size_t)f * H * W;
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
    for (int idx = tid; idx < TH * (TW / 32); idx += 256) {
        const int r = idx / (TW / 32), wq = idx - r * (TW / 32);
        const int gy = y0 + r, gwx = (x0 >> 5) + wq;
        if (gy >= H || gwx >= b.ww) continue;
        uint32_t word = 0;
        if (!flat) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(&s_f[r][wq * 32]);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                // gather bit 0 of each of the 4 mask bytes into a nibble
                const uint32_t nib = ((src[k] & 0x01010101u) * 0x10204080u) >> 28;
                word |= nib << (4 * k);
            }
        }
        bits_out[((size_t)f * H + gy) * b.ww + gwx] = word;
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
