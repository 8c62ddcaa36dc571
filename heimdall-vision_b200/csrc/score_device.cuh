// score_device.cuh -- per-blob scoring shared by the global-path K6 (k_score.cu) and the fused per-frame kernel
// (k_ccl_frame.cu).  Restates rust/heimdall-core/src/detection.rs:247-311; see k_score.cu for the commentary.
#pragma once
#include "hv_common.cuh"

namespace hv {
namespace {

struct Scored {
    bool keep;
    hv_defect d;
};

// LOW_REG = false: all 50 probe loads in flight together (50 registers).  LOW_REG = true (the 64-register build of the
// per-frame kernel): the 25 mask bytes first, folded into a bit mask, then the 25 gray bytes -- two round trips, half the
// registers; with everything in flight that build spilled and its scoring phase took 20 us instead of 6.
template <bool LOW_REG = false>
__device__ __forceinline__ Scored score_blob(const BatchView &b, const ScoreParams &p, int f, uint32_t k,
                                             const hv_blob &q) {
    Scored out;
    out.keep = false;
    const double area = (double)q.area;
    if (!(area >= p.min_size && area <= p.max_size) || q.area == 0) return out;
    const int H = b.h, W = b.w;
    const uint64_t cy = q.sum_y / q.area, cx = q.sum_x / q.area;
    const int icy = (int)cy, icx = (int)cx;
    const int y_lo = max(icy - 2, 0), y_hi = min(icy + 2, H - 1);
    const int x_lo = max(icx - 2, 0), x_hi = min(icx + 2, W - 1);
    const uint8_t *gray = b.gray + (size_t)f * b.gray_frame_stride;
    const uint8_t *mask = b.mask + (size_t)f * H * W;
    uint32_t fg_sum = 0, bg_sum = 0, fg_cnt = 0, bg_cnt = 0;
    const bool words_ok = W >= 8 && (W & 3) == 0 && (b.gray_row_stride & 3) == 0 && (b.gray_frame_stride & 3) == 0 &&
                          ((reinterpret_cast<uintptr_t>(b.gray) | reinterpret_cast<uintptr_t>(b.mask)) & 3) == 0;
    if (LOW_REG && words_ok) {
        // The five bytes of a probe row lie inside the eight bytes that start at the 4-byte boundary at or below their first
        // one: two 32-bit loads per row and plane, 20 loads for the whole probe, all in flight together -- one round trip
        // with 20 registers instead of five round trips (one row each) or 50 registers.  Rows / columns outside the clamped
        // window are loaded from clamped addresses and ignored.
        const int xa = min(max(icx - 2, 0) & ~3, W - 8);  // first byte of the 8-byte window, inside the row
        uint32_t gw[5][2], mw[5][2];
#pragma unroll
        for (int dy = 0; dy < 5; dy++) {
            const int yc = min(max(icy - 2 + dy, 0), H - 1);
            const uint32_t *grow = reinterpret_cast<const uint32_t *>(gray + (size_t)yc * b.gray_row_stride + xa);
            const uint32_t *mrow = reinterpret_cast<const uint32_t *>(mask + (size_t)yc * W + xa);
            gw[dy][0] = grow[0], gw[dy][1] = grow[1];
            mw[dy][0] = __ldcg(mrow), mw[dy][1] = __ldcg(mrow + 1);
        }
#pragma unroll
        for (int dy = 0; dy < 5; dy++) {
            const int y = icy - 2 + dy;
            if (y < y_lo || y > y_hi) continue;
            const uint64_t g8 = (uint64_t)gw[dy][0] | ((uint64_t)gw[dy][1] << 32), m8 = (uint64_t)mw[dy][0] | ((uint64_t)mw[dy][1] << 32);
#pragma unroll
            for (int dx = 0; dx < 5; dx++) {
                const int x = icx - 2 + dx;
                if (x >= x_lo && x <= x_hi) {
                    const int sh = 8 * (x - xa);
                    const uint32_t g = (uint32_t)(g8 >> sh) & 0xffu, m = (uint32_t)(m8 >> sh) & 0xffu;
                    if (m == 255) {
                        fg_sum += g;
                        fg_cnt++;
                    } else {
                        bg_sum += g;
                        bg_cnt++;
                    }
                }
            }
        }
    } else if (LOW_REG) {
        // one probe row per round trip (10 loads in flight), accumulated at once
#pragma unroll 1
        for (int dy = -2; dy <= 2; dy++) {
            const int y = icy + dy;
            if (y < y_lo || y > y_hi) continue;
            const uint8_t *grow = gray + (size_t)y * b.gray_row_stride;
            const uint8_t *mrow = mask + (size_t)y * W;
            uint32_t g[5], m[5];
#pragma unroll
            for (int dx = 0; dx < 5; dx++) {
                const int xc = min(max(icx - 2 + dx, 0), W - 1);
                g[dx] = grow[xc];
                m[dx] = __ldcg(mrow + xc);
            }
#pragma unroll
            for (int dx = 0; dx < 5; dx++) {
                const int x = icx - 2 + dx;
                if (x >= x_lo && x <= x_hi) {
                    if (m[dx] == 255) {
                        fg_sum += g[dx];
                        fg_cnt++;
                    } else {
                        bg_sum += g[dx];
                        bg_cnt++;
                    }
                }
            }
        }
    } else {
        // 5x5 probe, fully unrolled with clamped coordinates so that all 50 loads are in flight together
        uint32_t gv[25], mv[25];
#pragma unroll
        for (int k = 0; k < 25; k++) {
            const int y = min(max(icy - 2 + k / 5, 0), H - 1), x = min(max(icx - 2 + k % 5, 0), W - 1);
            gv[k] = gray[(size_t)y * b.gray_row_stride + x];
            mv[k] = __ldcg(mask + (size_t)y * W + x);  // written by K1 while this CTA may already have been resident: L2
        }
#pragma unroll
        for (int k = 0; k < 25; k++) {
            const int y = icy - 2 + k / 5, x = icx - 2 + k % 5;
            if (y >= y_lo && y <= y_hi && x >= x_lo && x <= x_hi) {
                if (mv[k] == 255) {
                    fg_sum += gv[k];
                    fg_cnt++;
                } else {
                    bg_sum += gv[k];
                    bg_cnt++;
                }
            }
        }
    }
    const double fg_mean = fg_cnt ? __ddiv_rn((double)fg_sum, (double)fg_cnt) : 127.0;
    const double bg_mean = bg_cnt ? __ddiv_rn((double)bg_sum, (double)bg_cnt) : 127.0;
    const double idiff = fabs(__dsub_rn(bg_mean, fg_mean));
    const uint64_t rect = (uint64_t)(q.ymax - q.ymin + 1) * (uint64_t)(q.xmax - q.xmin + 1);
    const double shape = rect > 0 ? __dsub_rn(1.0, __ddiv_rn(area, (double)rect)) : 0.5;
    double iscore = __ddiv_rn(idiff, 30.0);
    if (!(iscore <= 1.0)) iscore = 1.0;
    const double conf = __dadd_rn(__dmul_rn(iscore, 0.7), __dmul_rn(shape, 0.3));
    if (conf >= p.min_confidence) {
        out.keep = true;
        out.d.y = icy;
        out.d.x = icx;
        out.d.size = area;
        out.d.confidence = conf;
        out.d.ymin = (int32_t)q.ymin;
        out.d.xmin = (int32_t)q.xmin;
        out.d.ymax = (int32_t)q.ymax;
        out.d.xmax = (int32_t)q.xmax;
        out.d.label = k + 1;
        out.d.frame = (uint32_t)f;
    }
    return out;
}

__device__ __forceinline__ int area_bin(uint32_t area) {
    const int bin = 31 - __clz(area | 1u);
    return bin < HV_STATS_AREA_BINS ? bin : HV_STATS_AREA_BINS - 1;
}


}  // namespace
}  // namespace hv
