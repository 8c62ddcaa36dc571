/*
 * hv_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, never shipped, never on the product path).
 *
 * A plain-C, single-threaded restatement of the heimdall-vision Rust hot path
 * (rust/heimdall-core/src/detection.rs, processing.rs; defaults from lib.rs) plus the OpenCV-semantics
 * extension stages (Gaussian blur, rectangular morphology) that the reference only has on its Python side
 * (heimdall/detectors/contamination_detector.py:66,81-87).
 *
 * PARITY STATUS: "parity unpinned by the reference" for the Rust path -- the reference has no tests, golden
 * vectors or fixtures with expected outputs for it and its Rust sources cannot be compiled in this environment
 * (no cargo/rustc; SURVEY.md F6,F7,F9).  The oracle is pinned instead by (a) the hand-derived known-answer
 * tests of SURVEY.md section 8c, (b) an independent literal numpy/python transcription (tests/ref_literal.py),
 * (c) cv2.connectedComponentsWithStats(connectivity=4) for labels/stats.  The extension stages ARE pinned
 * against the reference's real third-party dependency: outputs of cv2 (opencv-python 4.13.0) committed under
 * tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 */
#ifndef HV_ORACLE_H
#define HV_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HVO_OK 0
#define HVO_ERR_DIMS (-1)     /* "Invalid image dimensions: expected 3D array" */
#define HVO_ERR_CHANNELS (-2) /* stage requires a 1-channel image */
#define HVO_ERR_CAPACITY (-3)
#define HVO_ERR_ARG (-4)

/* One blob (4-connected component), in discovery (= raster-first-pixel) order. */
typedef struct {
    uint32_t area;
    uint32_t ymin, ymax, xmin, xmax;
    uint64_t sum_y, sum_x;
} hvo_blob;

/* One emitted defect; mirrors detection.rs:12-18 (position = (row, col)). */
typedef struct {
    int32_t y, x;
    double size;
    double confidence;
    int32_t ymin, xmin, ymax, xmax;
    uint32_t label; /* 1-based canonical component label */
} hvo_defect;

/* A1: gray stage. detection.rs:135-160. c==3 -> f64 weighted sum truncated; c==1 -> copy; else HVO_ERR_DIMS. */
int hvo_gray(const uint8_t *img, int h, int w, int c, uint8_t *gray);

/* A2: box "Gaussian" blur, interior only. detection.rs:162-182 (radius 2); processing.rs:66-94 (radius = blur_size/2,
 * nch channels interleaved). */
void hvo_box_blur(const uint8_t *src, int h, int w, int nch, int radius, uint8_t *dst);

/* A3: adaptive mean threshold with edge-truncated 11x11 window. detection.rs:184-213; processing.rs:131-164.
 * inverse!=0: 255 if px < mean-c ; inverse==0: 255 if px > mean-c. */
void hvo_adaptive_threshold(const uint8_t *src, int h, int w, int32_t c, int inverse, uint8_t *mask);

/* A3': global threshold. processing.rs:165-178. */
void hvo_global_threshold(const uint8_t *src, int h, int w, uint8_t thr, int inverse, uint8_t *mask);

/* `threshold as i32` (detection.rs:186): saturating, truncating, NaN -> 0. */
int32_t hvo_f64_as_i32(double v);

/* A4: 4-connected flood fill in raster order (explicit stack, same neighbour order as detection.rs:219-245).
 * fg_gt127!=0 uses the `> 127` predicate of find_contours (detection.rs:64), else `== 255`.
 * labels: h*w int32, 0 = background, k = k-th component discovered. blobs (may be NULL) receives up to blob_cap
 * records. pop_order (may be NULL): h*w int32 receiving pixel linear indices in DFS pop order, component after
 * component (used by find_contours' "points"). Returns component count or a negative error. */
int64_t hvo_label4(const uint8_t *mask, int h, int w, int fg_gt127, int32_t *labels, hvo_blob *blobs,
                   size_t blob_cap, int32_t *pop_order);

/* A5/A5b/A6: score blobs -> defects. detection.rs:247-311. Returns number of defects (<= cap) or error. */
int64_t hvo_score_blobs(const uint8_t *gray, const uint8_t *mask, int h, int w, const hvo_blob *blobs,
                        size_t nblobs, double min_size, double max_size, hvo_defect *defects, size_t cap);

/* Whole path: detection.rs:127-317. Any of gray/blur/mask/labels may be NULL. morph_open_k / morph_close_k = 0
 * gives the reference-exact Rust path; >0 inserts OpenCV-semantics open then close on the mask before CCL
 * (contamination_detector.py:81-87). gauss_ksize>0 replaces the box blur by cv2.GaussianBlur(k,k,sigma). */
typedef struct {
    double min_size, max_size, threshold;
    int32_t gauss_ksize;
    double gauss_sigma;
    int32_t morph_open_k, morph_close_k;
} hvo_params;

int64_t hvo_detect_contamination(const uint8_t *img, int h, int w, int c, const hvo_params *p, uint8_t *gray,
                                 uint8_t *blur, uint8_t *mask, int32_t *labels, int64_t *ncomp,
                                 hvo_defect *defects, size_t cap);

/* processing.rs:30-101. out has (grayscale ? 1 : c) channels. blur_size <= 0 -> no blur. */
int hvo_preprocess_image(const uint8_t *img, int h, int w, int c, int grayscale, int blur_size, uint8_t *out);

/* processing.rs:104-185 (c must be 1). */
int hvo_apply_threshold(const uint8_t *img, int h, int w, int c, uint8_t thr, int adaptive, int inverse,
                        uint8_t *out);

/* processing.rs:188-249. img must have >=3 channels (the reference indexes channels 0..2 unconditionally). */
int hvo_basic_pipeline(const uint8_t *img, int h, int w, int c, uint8_t *out_hw3);

/* processing.rs:252-404. contours: triples (cy, cx) + fixed conf 0.75; returns count. */
typedef struct {
    int32_t y, x;
    double confidence;
} hvo_contour;
int64_t hvo_contamination_pipeline(const uint8_t *img, int h, int w, int c, uint8_t *out_hw3,
                                   hvo_contour *contours, size_t cap);

/* detection.rs:36-124. */
typedef struct {
    int32_t y, x;
    double area;
    uint64_t pixel_count;
    uint64_t points_offset; /* into pop_order when pixel_count <= 100, else UINT64_MAX */
} hvo_contour_rec;
int64_t hvo_find_contours(const uint8_t *mask, int h, int w, int c, double min_area, double max_area,
                          hvo_contour_rec *recs, size_t cap, int32_t *pop_order);

/* A7 extension: cv2.GaussianBlur(src,(k,k),sigma) for CV_8U, BORDER_REFLECT_101, OpenCV's 8.8 fixed-point path. */
int hvo_gaussian_blur(const uint8_t *src, int h, int w, int ksize, double sigma, uint8_t *dst);
/* Fixed-point (1/256) kernel used by hvo_gaussian_blur; k16 receives ksize coefficients summing to 256. */
int hvo_gaussian_kernel_q8(int ksize, double sigma, uint16_t *k16);

/* A8 extension: rect kxk erode/dilate/open/close with OpenCV default border (erode: +inf, dilate: -inf),
 * anchor at k/2. op: 0 erode, 1 dilate, 2 open, 3 close. */
int hvo_morph(const uint8_t *src, int h, int w, int op, int k, uint8_t *dst);

/* N1 (frame feed): camera pixel formats -> interleaved RGB, the conversions named by the reference's
 * to_opencv_mat (rust/heimdall-camera/src/lib.rs:226-252; note that in the reference itself these branches are
 * unreachable, the `mat_type` match at :207-216 returns ConversionError for Bayer and YUV frames first).
 * Semantics = cv2.cvtColor with COLOR_Bayer{RG,GB,GR,BG}2RGB (bilinear, 1-px border replicated from the interior;
 * frames smaller than 3x3 give zeros) and COLOR_YUV2RGB_YUYV (BT.601 limited range, 20-bit fixed point); pinned
 * against opencv-python 4.13.0 in tests/golden/cv2_pixfmt.npz.  pattern: 0 RG, 1 GB, 2 GR, 3 BG. */
int hvo_bayer_to_rgb(const uint8_t *bayer, int h, int w, int pattern, uint8_t *rgb);
int hvo_yuyv_to_rgb(const uint8_t *yuyv, int h, int w, uint8_t *rgb);

const char *hvo_version(void);

#ifdef __cplusplus
}
#endif
#endif
