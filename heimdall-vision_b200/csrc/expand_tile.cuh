// expand_tile.cuh -- mask bytes + label-plane initialisation of one 128 x 32 tile from its bit-packed rows; shared by the
// morphology tiles kernel / k_expand_bits (k_stage.cu) and the K1 variant with the morphology folded in (k_preprocess.cu).
#pragma once

#include "hv_common.cuh"

namespace hv {

// Mask bytes and label-plane initialisation (p + 1 at the first pixel of every word-run, see k_preprocess.cu) of one
// 128 x 32 tile from its bit-packed rows s_w[32][4], with the store pattern of K1: all-zero tiles (most of an
// inspection frame) are nothing but 128-bit zero stores.
__device__ __forceinline__ void expand_tile(const BatchView &b, size_t f, int tx, int ty, const uint32_t (*s_w)[4], bool any,
                                            int tid) {
    const int H = b.h, W = b.w;
    const int x0 = tx * 128, y0 = ty * 32;
    const size_t pix0 = (f * H + y0) * (size_t)W + x0;
    if ((W & 15) == 0 && x0 + 128 <= W) {
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
            const uint32_t m = any ? s_w[r][c >> 5] : 0u;
            const int bit = c & 31;
            const bool fg = (m >> bit) & 1u;
            const bool start = fg && (bit == 0 || !((m >> (bit - 1)) & 1u));
            b.mask[pix0 + (size_t)r * W + c] = fg ? 255 : 0;
            b.labels[pix0 + (size_t)r * W + c] = start ? (gy * W + gx + 1) : 0;
        }
    }
}

}  // namespace hv
