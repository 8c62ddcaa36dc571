/*
 * hv_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see hv_oracle.h for the parity status).
 *
 * Single-threaded, scalar, deliberately literal: loop shapes, integer widths (u32 sums and counts, usize centroid
 * sums, i32 comparisons), floor divisions and f64 expression order follow the Rust sources statement by
 * statement.  Build with -ffp-contract=off (rustc never contracts a*b+c into an FMA).
 * Citations are relative to /root/reference/.
 */
#include "hv_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

const char *hvo_version(void) { return "hv_oracle 1.0 (restates heimdall-core detection.rs/processing.rs)"; }

/* Rust `f64 as u8`: truncate toward zero, saturate, NaN -> 0. */
static uint8_t f64_as_u8(double v) {
    if (!(v == v)) return 0;
    if (v <= 0.0) return 0;
    if (v >= 255.0) return 255;
    return (uint8_t)v;
}

/* Rust `f64 as i32` (detection.rs:186 `threshold as i32`). */
int32_t hvo_f64_as_i32(double v) {
    if (!(v == v)) return 0;
    if (v >= 2147483647.0) return INT32_MAX;
    if (v <= -2147483648.0) return INT32_MIN;
    return (int32_t)v;
}

/* detection.rs:135-160 (same expression at processing.rs:56,200,264). */
int hvo_gray(const uint8_t *img, int h, int w, int c, uint8_t *gray) {
    if (c == 3) {
        for (int i = 0; i < h; i++)
            for (int j = 0; j < w; j++) {
                const uint8_t *p = img + ((size_t)i * w + j) * 3;
                uint32_t r = p[0], g = p[1], b = p[2];
                /* (0.299 * r + 0.587 * g) + 0.114 * b, each op rounded separately */
                double t0 = 0.299 * (double)r;
                double t1 = 0.587 * (double)g;
                double t2 = 0.114 * (double)b;
                double s = t0 + t1;
                s = s + t2;
                gray[(size_t)i * w + j] = f64_as_u8(s);
            }
        return HVO_OK;
    }
    if (c == 1) {
        memcpy(gray, img, (size_t)h * w);
        return HVO_OK;
    }
    return HVO_ERR_DIMS;
}

/* gray from the first three channels unconditionally (processing.rs:47-59,191-203,255-267 index channels 0,1,2
 * without checking; fewer than 3 channels panics in the reference -> we return HVO_ERR_DIMS). */
static int gray_first3(const uint8_t *img, int h, int w, int c, uint8_t *gray) {
    if (c < 3) return HVO_ERR_DIMS;
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            const uint8_t *p = img + ((size_t)i * w + j) * c;
            double t0 = 0.299 * (double)(uint32_t)p[0];
            double t1 = 0.587 * (double)(uint32_t)p[1];
            double t2 = 0.114 * (double)(uint32_t)p[2];
            double s = t0 + t1;
            s = s + t2;
            gray[(size_t)i * w + j] = f64_as_u8(s);
        }
    return HVO_OK;
}

/* detection.rs:162-182 / processing.rs:66-94. dst = src outside the interior. */
void hvo_box_blur(const uint8_t *src, int h, int w, int nch, int radius, uint8_t *dst) {
    memcpy(dst, src, (size_t)h * w * nch);
    for (int i = radius; i < h - radius; i++)
        for (int j = radius; j < w - radius; j++)
            for (int ch = 0; ch < nch; ch++) {
                uint32_t sum = 0, count = 0;
                for (int bi = -radius; bi <= radius; bi++)
                    for (int bj = -radius; bj <= radius; bj++) {
                        sum += src[((size_t)(i + bi) * w + (j + bj)) * nch + ch];
                        count += 1;
                    }
                dst[((size_t)i * w + j) * nch + ch] = (uint8_t)(sum / count);
            }
}

/* detection.rs:184-213 / processing.rs:131-164,291-320. */
void hvo_adaptive_threshold(const uint8_t *src, int h, int w, int32_t c, int inverse, uint8_t *mask) {
    const int half = 11 / 2;
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            int start_i = i - half < 0 ? 0 : i - half; /* saturating_sub */
            int end_i = i + half < h - 1 ? i + half : h - 1;
            int start_j = j - half < 0 ? 0 : j - half;
            int end_j = j + half < w - 1 ? j + half : w - 1;
            uint32_t sum = 0, count = 0;
            for (int y = start_i; y <= end_i; y++)
                for (int x = start_j; x <= end_j; x++) {
                    sum += src[(size_t)y * w + x];
                    count += 1;
                }
            int32_t mean = (int32_t)(sum / count);
            int32_t px = src[(size_t)i * w + j];
            /* `mean - c` in i32 (detection.rs:211).  The reference's release profile (rust/Cargo.toml [profile.release])
             * does not enable overflow-checks, so the subtraction WRAPS: for c <= mean - 2^31 (only reachable with
             * threshold <= -2147483393.0, since `threshold as i32` saturates at i32::MIN) the right-hand side becomes a
             * large negative number and the inverse test is false.  Unsigned arithmetic is the well-defined C spelling of
             * that wrap. */
            int32_t rhs = (int32_t)((uint32_t)mean - (uint32_t)c);
            uint8_t v;
            if (inverse)
                v = (px < rhs) ? 255 : 0;
            else
                v = (px > rhs) ? 255 : 0;
            mask[(size_t)i * w + j] = v;
        }
}

/* processing.rs:165-178,227-235. */
void hvo_global_threshold(const uint8_t *src, int h, int w, uint8_t thr, int inverse, uint8_t *mask) {
    for (size_t k = 0, n = (size_t)h * w; k < n; k++)
        mask[k] = inverse ? (src[k] < thr ? 255 : 0) : (src[k] > thr ? 255 : 0);
}

/* detection.rs:215-245 (and :58-88, processing.rs:322-353). */
int64_t hvo_label4(const uint8_t *mask, int h, int w, int fg_gt127, int32_t *labels, hvo_blob *blobs,
                   size_t blob_cap, int32_t *pop_order) {
    size_t n = (size_t)h * w;
    uint8_t *visited = (uint8_t *)calloc(n ? n : 1, 1);
    int32_t *stack = (int32_t *)malloc((n ? n : 1) * sizeof(int32_t));
    if (!visited || !stack) {
        free(visited);
        free(stack);
        return HVO_ERR_ARG;
    }
    if (labels) memset(labels, 0, n * sizeof(int32_t));
    int64_t ncomp = 0;
    size_t pop_pos = 0;
#define HVO_FG(idx) (fg_gt127 ? (mask[idx] > 127) : (mask[idx] == 255))
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            size_t idx = (size_t)i * w + j;
            if (!HVO_FG(idx) || visited[idx]) continue;
            ncomp++;
            hvo_blob b;
            b.area = 0;
            b.ymin = b.ymax = (uint32_t)i;
            b.xmin = b.xmax = (uint32_t)j;
            b.sum_y = b.sum_x = 0;
            size_t sp = 0;
            stack[sp++] = (int32_t)idx;
            visited[idx] = 1;
            while (sp) {
                int32_t p = stack[--sp];
                int y = p / w, x = p % w;
                if (labels) labels[p] = (int32_t)ncomp;
                if (pop_order) pop_order[pop_pos] = p;
                pop_pos++;
                b.area++;
                b.sum_y += (uint64_t)y;
                b.sum_x += (uint64_t)x;
                if ((uint32_t)y < b.ymin) b.ymin = (uint32_t)y;
                if ((uint32_t)y > b.ymax) b.ymax = (uint32_t)y;
                if ((uint32_t)x < b.xmin) b.xmin = (uint32_t)x;
                if ((uint32_t)x > b.xmax) b.xmax = (uint32_t)x;
                /* neighbours: (y-1 sat, x), (y+1, x), (y, x-1 sat), (y, x+1); saturating_sub makes the pixel its
                 * own neighbour at the border, which is already visited */
                int ny[4] = {y > 0 ? y - 1 : 0, y + 1, y, y};
                int nx[4] = {x, x, x > 0 ? x - 1 : 0, x + 1};
                for (int k = 0; k < 4; k++) {
                    if (ny[k] < h && nx[k] < w) {
                        size_t q = (size_t)ny[k] * w + nx[k];
                        if (HVO_FG(q) && !visited[q]) {
                            stack[sp++] = (int32_t)q;
                            visited[q] = 1;
                        }
                    }
                }
            }
            if (blobs) {
                if ((size_t)(ncomp - 1) >= blob_cap) {
                    free(visited);
                    free(stack);
                    return HVO_ERR_CAPACITY;
                }
                blobs[ncomp - 1] = b;
            }
        }
#undef HVO_FG
    free(visited);
    free(stack);
    return ncomp;
}

/* detection.rs:247-311. */
int64_t hvo_score_blobs(const uint8_t *gray, const uint8_t *mask, int h, int w, const hvo_blob *blobs,
                        size_t nblobs, double min_size, double max_size, hvo_defect *defects, size_t cap) {
    int64_t nd = 0;
    for (size_t k = 0; k < nblobs; k++) {
        const hvo_blob *b = &blobs[k];
        double area = (double)b->area;
        if (!(area >= min_size && area <= max_size)) continue;
        uint64_t center_y = b->sum_y / b->area;
        uint64_t center_x = b->sum_x / b->area;
        uint32_t fg_sum = 0, bg_sum = 0, fg_count = 0, bg_count = 0;
        const uint64_t margin = 2;
        uint64_t start_i = center_y >= margin ? center_y - margin : 0;
        uint64_t end_i = center_y + margin < (uint64_t)(h - 1) ? center_y + margin : (uint64_t)(h - 1);
        uint64_t start_j = center_x >= margin ? center_x - margin : 0;
        uint64_t end_j = center_x + margin < (uint64_t)(w - 1) ? center_x + margin : (uint64_t)(w - 1);
        for (uint64_t y = start_i; y <= end_i; y++)
            for (uint64_t x = start_j; x <= end_j; x++) {
                size_t q = (size_t)y * w + x;
                if (mask[q] == 255) {
                    fg_sum += gray[q];
                    fg_count += 1;
                } else {
                    bg_sum += gray[q];
                    bg_count += 1;
                }
            }
        double fg_mean = fg_count > 0 ? (double)fg_sum / (double)fg_count : 127.0;
        double bg_mean = bg_count > 0 ? (double)bg_sum / (double)bg_count : 127.0;
        double intensity_diff = fabs(bg_mean - fg_mean);
        uint64_t rect_area = (uint64_t)(b->ymax - b->ymin + 1) * (uint64_t)(b->xmax - b->xmin + 1);
        double shape_score = rect_area > 0 ? 1.0 - (area / (double)rect_area) : 0.5;
        double intensity_score = intensity_diff / 30.0;
        if (!(intensity_score <= 1.0)) intensity_score = 1.0; /* f64::min(1.0); NaN impossible here */
        double t0 = intensity_score * 0.7;
        double t1 = shape_score * 0.3;
        double confidence = t0 + t1;
        if (confidence >= 0.3) {
            if ((size_t)nd >= cap) return HVO_ERR_CAPACITY;
            hvo_defect *d = &defects[nd++];
            d->y = (int32_t)center_y;
            d->x = (int32_t)center_x;
            d->size = area;
            d->confidence = confidence;
            d->ymin = (int32_t)b->ymin;
            d->xmin = (int32_t)b->xmin;
            d->ymax = (int32_t)b->ymax;
            d->xmax = (int32_t)b->xmax;
            d->label = (uint32_t)(k + 1);
        }
    }
    return nd;
}

int64_t hvo_detect_contamination(const uint8_t *img, int h, int w, int c, const hvo_params *p, uint8_t *gray_out,
                                 uint8_t *blur_out, uint8_t *mask_out, int32_t *labels_out, int64_t *ncomp_out,
                                 hvo_defect *defects, size_t cap) {
    if (h <= 0 || w <= 0) return HVO_ERR_ARG;
    if (c != 1 && c != 3) return HVO_ERR_DIMS;
    size_t n = (size_t)h * w;
    uint8_t *gray = (uint8_t *)malloc(n), *blur = (uint8_t *)malloc(n), *mask = (uint8_t *)malloc(n);
    uint8_t *tmp = (uint8_t *)malloc(n);
    hvo_blob *blobs = (hvo_blob *)malloc((n / 2 + 1) * sizeof(hvo_blob));
    int32_t *labels = labels_out ? labels_out : NULL;
    int64_t rc = HVO_ERR_ARG;
    if (!gray || !blur || !mask || !tmp || !blobs) goto done;
    rc = hvo_gray(img, h, w, c, gray);
    if (rc != HVO_OK) goto done;
    if (p->gauss_ksize > 0) {
        rc = hvo_gaussian_blur(gray, h, w, p->gauss_ksize, p->gauss_sigma, blur);
        if (rc != HVO_OK) goto done;
    } else {
        hvo_box_blur(gray, h, w, 1, 2, blur);
    }
    hvo_adaptive_threshold(blur, h, w, hvo_f64_as_i32(p->threshold), 1, mask);
    if (p->morph_open_k > 0) {
        rc = hvo_morph(mask, h, w, 2, p->morph_open_k, tmp);
        if (rc != HVO_OK) goto done;
        memcpy(mask, tmp, n);
    }
    if (p->morph_close_k > 0) {
        rc = hvo_morph(mask, h, w, 3, p->morph_close_k, tmp);
        if (rc != HVO_OK) goto done;
        memcpy(mask, tmp, n);
    }
    {
        int64_t nc = hvo_label4(mask, h, w, 0, labels, blobs, n / 2 + 1, NULL);
        if (nc < 0) {
            rc = nc;
            goto done;
        }
        if (ncomp_out) *ncomp_out = nc;
        rc = hvo_score_blobs(gray, mask, h, w, blobs, (size_t)nc, p->min_size, p->max_size, defects, cap);
    }
    if (gray_out) memcpy(gray_out, gray, n);
    if (blur_out) memcpy(blur_out, blur, n);
    if (mask_out) memcpy(mask_out, mask, n);
done:
    free(gray);
    free(blur);
    free(mask);
    free(tmp);
    free(blobs);
    return rc;
}

/* processing.rs:30-101. */
int hvo_preprocess_image(const uint8_t *img, int h, int w, int c, int grayscale, int blur_size, uint8_t *out) {
    int och = grayscale ? 1 : c;
    size_t n = (size_t)h * w * och;
    uint8_t *tmp = (uint8_t *)malloc(n ? n : 1);
    if (!tmp) return HVO_ERR_ARG;
    int rc = HVO_OK;
    if (grayscale) {
        rc = gray_first3(img, h, w, c, tmp);
        if (rc != HVO_OK) {
            free(tmp);
            return rc;
        }
    } else {
        memcpy(tmp, img, n);
    }
    if (blur_size > 0) {
        hvo_box_blur(tmp, h, w, och, blur_size / 2, out);
    } else {
        memcpy(out, tmp, n);
    }
    free(tmp);
    return HVO_OK;
}

/* processing.rs:104-185. */
int hvo_apply_threshold(const uint8_t *img, int h, int w, int c, uint8_t thr, int adaptive, int inverse,
                        uint8_t *out) {
    if (c != 1) return HVO_ERR_CHANNELS;
    if (adaptive)
        hvo_adaptive_threshold(img, h, w, 2, inverse, out);
    else
        hvo_global_threshold(img, h, w, thr, inverse, out);
    return HVO_OK;
}

/* processing.rs:188-249. */
int hvo_basic_pipeline(const uint8_t *img, int h, int w, int c, uint8_t *out_hw3) {
    size_t n = (size_t)h * w;
    uint8_t *gray = (uint8_t *)malloc(n ? n : 1), *blur = (uint8_t *)malloc(n ? n : 1);
    if (!gray || !blur) {
        free(gray);
        free(blur);
        return HVO_ERR_ARG;
    }
    int rc = gray_first3(img, h, w, c, gray);
    if (rc == HVO_OK) {
        hvo_box_blur(gray, h, w, 1, 2, blur);
        for (size_t k = 0; k < n; k++) {
            uint8_t v = blur[k] > 127 ? 255 : 0;
            out_hw3[3 * k] = out_hw3[3 * k + 1] = out_hw3[3 * k + 2] = v;
        }
    }
    free(gray);
    free(blur);
    return rc;
}

/* processing.rs:252-404. */
int64_t hvo_contamination_pipeline(const uint8_t *img, int h, int w, int c, uint8_t *out_hw3,
                                   hvo_contour *contours, size_t cap) {
    size_t n = (size_t)h * w;
    uint8_t *gray = (uint8_t *)malloc(n ? n : 1), *blur = (uint8_t *)malloc(n ? n : 1);
    uint8_t *mask = (uint8_t *)malloc(n ? n : 1);
    hvo_blob *blobs = (hvo_blob *)malloc((n / 2 + 1) * sizeof(hvo_blob));
    int64_t rc = HVO_ERR_ARG, nc = 0, nout = 0;
    if (!gray || !blur || !mask || !blobs) goto done;
    rc = gray_first3(img, h, w, c, gray);
    if (rc != HVO_OK) goto done;
    hvo_box_blur(gray, h, w, 1, 2, blur);
    hvo_adaptive_threshold(blur, h, w, 15, 1, mask);
    nc = hvo_label4(mask, h, w, 0, NULL, blobs, n / 2 + 1, NULL);
    if (nc < 0) {
        rc = nc;
        goto done;
    }
    for (size_t k = 0; k < n; k++) out_hw3[3 * k] = out_hw3[3 * k + 1] = out_hw3[3 * k + 2] = mask[k];
    for (int64_t k = 0; k < nc; k++) {
        if (blobs[k].area >= 3) {
            if ((size_t)nout >= cap) {
                rc = HVO_ERR_CAPACITY;
                goto done;
            }
            contours[nout].y = (int32_t)(blobs[k].sum_y / blobs[k].area);
            contours[nout].x = (int32_t)(blobs[k].sum_x / blobs[k].area);
            contours[nout].confidence = 0.75;
            nout++;
        }
    }
    /* crosses, in list order; vertical bar then horizontal bar (processing.rs:383-401) */
    for (int64_t k = 0; k < nout; k++) {
        int y = contours[k].y, x = contours[k].x;
        const int radius = 3;
        int i0 = y - radius < 0 ? 0 : y - radius, i1 = y + radius < h - 1 ? y + radius : h - 1;
        for (int i = i0; i <= i1; i++) {
            uint8_t *q = out_hw3 + ((size_t)i * w + x) * 3;
            q[0] = 0;
            q[1] = 0;
            q[2] = 255;
        }
        int j0 = x - radius < 0 ? 0 : x - radius, j1 = x + radius < w - 1 ? x + radius : w - 1;
        for (int j = j0; j <= j1; j++) {
            uint8_t *q = out_hw3 + ((size_t)y * w + j) * 3;
            q[0] = 0;
            q[1] = 0;
            q[2] = 255;
        }
    }
    rc = nout;
done:
    free(gray);
    free(blur);
    free(mask);
    free(blobs);
    return rc;
}

/* detection.rs:36-124. */
int64_t hvo_find_contours(const uint8_t *mask, int h, int w, int c, double min_area, double max_area,
                          hvo_contour_rec *recs, size_t cap, int32_t *pop_order) {
    if (c != 1) return HVO_ERR_CHANNELS;
    size_t n = (size_t)h * w;
    hvo_blob *blobs = (hvo_blob *)malloc((n / 2 + 1) * sizeof(hvo_blob));
    if (!blobs) return HVO_ERR_ARG;
    int64_t nc = hvo_label4(mask, h, w, 1, NULL, blobs, n / 2 + 1, pop_order);
    int64_t nout = 0;
    uint64_t off = 0;
    if (nc < 0) {
        free(blobs);
        return nc;
    }
    for (int64_t k = 0; k < nc; k++) {
        double area = (double)blobs[k].area;
        if (area >= min_area && area <= max_area) {
            if ((size_t)nout >= cap) {
                free(blobs);
                return HVO_ERR_CAPACITY;
            }
            recs[nout].y = (int32_t)(blobs[k].sum_y / blobs[k].area);
            recs[nout].x = (int32_t)(blobs[k].sum_x / blobs[k].area);
            recs[nout].area = area;
            recs[nout].pixel_count = blobs[k].area;
            recs[nout].points_offset = (blobs[k].area <= 100 && pop_order) ? off : UINT64_MAX;
            nout++;
        }
        off += blobs[k].area;
    }
    free(blobs);
    return nout;
}

/* ------------------------------------------------------------------------------------------------------------
 * Extension stages with OpenCV semantics (third-party dependency of the reference's Python path:
 * opencv-python, version unpinned in the reference (README.md:45); validated here against 4.13.0).
 * ---------------------------------------------------------------------------------------------------------- */

/* cv::getGaussianKernelBitExact + getGaussianKernelFixedPoint_ED (OpenCV 4.x modules/imgproc/src/smooth.dispatch.cpp):
 * double kernel normalised to sum 1, then rounded to 8 fractional bits with error diffusion from the ends toward
 * the centre, centre = 256 - 2*sum(others).  OpenCV evaluates exp() in softfloat; libm exp differs by at most an
 * ulp, which cannot change a coefficient unless a scaled value sits within 1e-13 of a rounding boundary. */
int hvo_gaussian_kernel_q8(int n, double sigma, uint16_t *k16) {
    if (n <= 0 || n > 31 || (n & 1) == 0) return HVO_ERR_ARG;
    double kd[31];
    int n2 = (n - 1) / 2;
    if (sigma <= 0 && n == 1) {
        kd[0] = 1.0;
    } else if (sigma <= 0 && n == 3) {
        kd[0] = 0.25, kd[1] = 0.5, kd[2] = 0.25;
    } else if (sigma <= 0 && n == 5) {
        kd[0] = 0.0625, kd[1] = 0.25, kd[2] = 0.375, kd[3] = 0.25, kd[4] = 0.0625;
    } else if (sigma <= 0 && n == 7) {
        kd[0] = 0.03125, kd[1] = 0.109375, kd[2] = 0.21875, kd[3] = 0.28125, kd[4] = 0.21875, kd[5] = 0.109375,
        kd[6] = 0.03125;
    } else if (sigma <= 0 && n == 9) {
        static const double v9[9] = {4.0 / 256, 13.0 / 256, 30.0 / 256, 51.0 / 256, 60.0 / 256,
                                     51.0 / 256, 30.0 / 256, 13.0 / 256, 4.0 / 256};
        memcpy(kd, v9, sizeof v9);
    } else {
        double sigmaX = sigma > 0 ? sigma : fma((double)n, 0.15, 0.35);
        double scale2X = -0.125 / (sigmaX * sigmaX);
        double values[16];
        double sum = 0.0;
        for (int i = 0, x = 1 - n; i < n2; i++, x += 2) {
            double t = exp((double)(x * x) * scale2X);
            values[i] = t;
            sum += t;
        }
        sum *= 2.0;
        sum += 1.0;
        double mul1 = 1.0 / sum;
        for (int i = 0; i < n2; i++) {
            double t = values[i] * mul1;
            kd[i] = t;
            kd[n - 1 - i] = t;
        }
        kd[n2] = 1.0 * mul1;
    }
    double err = 0.0;
    int64_t sum = 0;
    for (int i = 0; i < n2; i++) {
        double adj = kd[i] * 256.0 + err;
        int64_t v0 = (int64_t)nearbyint(adj); /* cvRound: round half to even */
        err = adj - (double)v0;
        k16[i] = (uint16_t)v0;
        k16[n - 1 - i] = (uint16_t)v0;
        sum += v0;
    }
    sum *= 2;
    k16[n2] = (uint16_t)(256 - sum);
    return HVO_OK;
}

static int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0)
            p = -p;
        else
            p = 2 * (len - 1) - p;
    }
    return p;
}

/* cv::GaussianBlur CV_8U fixed-point path: rows in 8.8 (u16), columns in 16.16, round-to-nearest at the end. */
int hvo_gaussian_blur(const uint8_t *src, int h, int w, int ksize, double sigma, uint8_t *dst) {
    uint16_t k[31];
    int rc = hvo_gaussian_kernel_q8(ksize, sigma, k);
    if (rc != HVO_OK) return rc;
    int r = ksize / 2;
    size_t n = (size_t)h * w;
    uint16_t *rows = (uint16_t *)malloc((n ? n : 1) * sizeof(uint16_t));
    if (!rows) return HVO_ERR_ARG;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            uint32_t s = 0;
            for (int t = -r; t <= r; t++) s += (uint32_t)k[t + r] * src[(size_t)y * w + reflect101(x + t, w)];
            rows[(size_t)y * w + x] = (uint16_t)(s > 65535u ? 65535u : s);
        }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            uint32_t s = 0;
            for (int t = -r; t <= r; t++) s += (uint32_t)k[t + r] * rows[(size_t)reflect101(y + t, h) * w + x];
            uint32_t v = (s + 32768u) >> 16;
            dst[(size_t)y * w + x] = (uint8_t)(v > 255u ? 255u : v);
        }
    free(rows);
    return HVO_OK;
}

/* rect kxk erode/dilate; anchor k/2; out-of-image pixels never win (OpenCV's default morphology border). */
static void morph1(const uint8_t *src, int h, int w, int dilate, int k, uint8_t *dst) {
    int a = k / 2;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int v = dilate ? 0 : 255;
            for (int dy = 0; dy < k; dy++) {
                int yy = y + dy - a;
                if (yy < 0 || yy >= h) continue;
                for (int dx = 0; dx < k; dx++) {
                    int xx = x + dx - a;
                    if (xx < 0 || xx >= w) continue;
                    int s = src[(size_t)yy * w + xx];
                    if (dilate ? s > v : s < v) v = s;
                }
            }
            dst[(size_t)y * w + x] = (uint8_t)v;
        }
}

int hvo_morph(const uint8_t *src, int h, int w, int op, int k, uint8_t *dst) {
    if (k <= 0 || op < 0 || op > 3) return HVO_ERR_ARG;
    if (op == 0 || op == 1) {
        morph1(src, h, w, op, k, dst);
        return HVO_OK;
    }
    size_t n = (size_t)h * w;
    uint8_t *tmp = (uint8_t *)malloc(n ? n : 1);
    if (!tmp) return HVO_ERR_ARG;
    if (op == 2) { /* open = erode then dilate */
        morph1(src, h, w, 0, k, tmp);
        morph1(tmp, h, w, 1, k, dst);
    } else { /* close = dilate then erode */
        morph1(src, h, w, 1, k, tmp);
        morph1(tmp, h, w, 0, k, dst);
    }
    free(tmp);
    return HVO_OK;
}

/* ---- N1: camera pixel formats (see hv_oracle.h) ------------------------------------------------------------------- */
/* Which neighbourhood average feeds output channel ch at a site of type t:
 *   0 centre, 1 horizontal pair, 2 vertical pair, 3 cross (4-neighbours), 4 diagonal (4 corners). */
static const uint8_t k_site[4][3] = {{4, 3, 0}, {0, 3, 4}, {2, 0, 1}, {1, 0, 2}};
/* site type at (y & 1, x & 1) for RG, GB, GR, BG (derived from cv2 by impulse responses, then checked exhaustively) */
static const uint8_t k_pat[4][2][2] = {{{0, 2}, {3, 1}}, {{3, 1}, {0, 2}}, {{2, 0}, {1, 3}}, {{1, 3}, {2, 0}}};

int hvo_bayer_to_rgb(const uint8_t *b, int h, int w, int pattern, uint8_t *rgb) {
    if (h <= 0 || w <= 0 || pattern < 0 || pattern > 3) return HVO_ERR_ARG;
    if (h < 3 || w < 3) {
        memset(rgb, 0, (size_t)h * w * 3);
        return HVO_OK;
    }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            /* border pixels copy the nearest interior result */
            int yy = y < 1 ? 1 : (y > h - 2 ? h - 2 : y), xx = x < 1 ? 1 : (x > w - 2 ? w - 2 : x);
            const uint8_t *p = b + (size_t)yy * w + xx;
            int v[5];
            v[0] = p[0];
            v[1] = (p[-1] + p[1] + 1) >> 1;
            v[2] = (p[-w] + p[w] + 1) >> 1;
            v[3] = (p[-1] + p[1] + p[-w] + p[w] + 2) >> 2;
            v[4] = (p[-w - 1] + p[-w + 1] + p[w - 1] + p[w + 1] + 2) >> 2;
            const uint8_t *t = k_site[k_pat[pattern][yy & 1][xx & 1]];
            for (int ch = 0; ch < 3; ch++) rgb[((size_t)y * w + x) * 3 + ch] = (uint8_t)v[t[ch]];
        }
    return HVO_OK;
}

static uint8_t sat_u8(int64_t v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

int hvo_yuyv_to_rgb(const uint8_t *s, int h, int w, uint8_t *rgb) {
    if (h <= 0 || w <= 0 || (w & 1)) return HVO_ERR_ARG;
    const int64_t CY = 1220542, CUB = 2116026, CUG = -409993, CVG = -852492, CVR = 1673527, HALF = 1 << 19;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x += 2) {
            const uint8_t *q = s + ((size_t)y * w + x) * 2; /* Y0 U Y1 V */
            int64_t u = (int64_t)q[1] - 128, v = (int64_t)q[3] - 128;
            int64_t ruv = HALF + CVR * v, guv = HALF + CVG * v + CUG * u, buv = HALF + CUB * u;
            for (int k = 0; k < 2; k++) {
                int64_t yy = (int64_t)q[2 * k] - 16;
                if (yy < 0) yy = 0;
                yy *= CY;
                uint8_t *o = rgb + ((size_t)y * w + x + k) * 3;
                o[0] = sat_u8((yy + ruv) >> 20);
                o[1] = sat_u8((yy + guv) >> 20);
                o[2] = sat_u8((yy + buv) >> 20);
            }
        }
    return HVO_OK;
}
