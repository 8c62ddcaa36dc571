// hv_common.cuh -- shared device/host declarations for the sm_100a backend (internal; the public ABI is
// include/heimdall_cuda.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/heimdall_cuda.h"

namespace hv {

constexpr int kAdaptHalf = 5;  // 11x11 adaptive window (detection.rs:185)

// Geometry of one batch resident on the device. All arrays are tightly packed, frame-major.
struct BatchView {
    int n, h, w;
    int ww;                    // bitmask words per row = ceil(w / 32)
    const uint8_t *gray;       // n*h*w  (the input itself when c == 1 and tightly packed)
    size_t gray_row_stride;    // bytes
    size_t gray_frame_stride;  // bytes
    uint8_t *blur;             // n*h*w, optional (debug / non-fused blur modes)
    uint8_t *mask;             // n*h*w  final mask {0,255}
    uint32_t *bits;            // n*h*ww bit-packed final mask, bit x&31 of word x>>5
    uint32_t *bits_tmp;        // scratch for morphology
    uint8_t *rowflags;         // n * rf_stride: one byte per (row, 128-px tile): bit k = word 4*tile+k of that row is non-zero;
                               // tile-major: byte (y & 31) of the 32-byte record of tile (y >> 5, tx), see rowflag_index()
    int tiles_x;               // ceil(w / 128)
    size_t rf_stride;          // bytes per frame in rowflags: ceil(h / 32) * tiles_x * 32
    uint8_t *tile_occ;         // NULL or n * ceil(h/32) * tiles_x * 4: per 128 x 32 tile, byte g = which of the tile's 4 words are
                               // non-zero in rows 8g..8g+7 (written by K1 when the fused morphology follows)
    int32_t *labels;           // n*h*w  union-find parents (+1) during CCL, canonical labels afterwards
    uint32_t *rootbits;        // n*h*ww
    uint32_t *rankbase;        // n*h*ww: number of roots in the words before this one inside its 256-word segment
    uint32_t *segbase;         // 2*n*nseg: roots per 256-word segment (K3), then their exclusive prefixes per frame (K4)
    int nseg;                  // ceil(h*ww / 256)
    uint32_t *score_state;     // n*(score_chunks+1): per-chunk defect counts of the scoring kernel + a completion counter
                               // per frame; all zero between launches (the kernel cleans up after itself, see k_score.cu)
    int score_chunks;          // ceil(blob_cap / 256)
    uint32_t *ncomp;           // n
    uint32_t *fgcount;         // n
    hv_blob *blobs;            // n*blob_cap
    int blob_cap;
    hv_defect *defects;        // n*defect_cap
    int defect_cap;
    hv_frame_result *results;  // n
    hv_line_stats *stats;      // 1
    uint32_t *frame_flags;     // n: written by the fused per-frame kernel, 1 = frame needs the global-memory CCL path
    unsigned long long *phase_ns;  // 256 or NULL: per-phase timestamps of frame `phase_frame` in the fused kernel (debug)
    int phase_frame;               // which frame's CTA records the stamps (HV_PHASE_FRAME, default 0)
    unsigned int *ccl_done;        // NULL or the slot's completion counter: +1 per frame when the per-frame CCL kernel is
                                   // through with it (K1 of the batch that reuses the slot waits for ccl_wait_value)
    unsigned int ccl_wait_value;   // K1: counter value that means "every earlier per-frame kernel on this slot is done"
    unsigned int *k1_done;         // NULL or the slot's K1 launch counter (+1 when the last CTA of a K1 launch has stored its last
                                   // tile), k1_done[-1] counts the CTAs of the running launch; the per-frame kernel waits for
    unsigned int k1_wait_value;    // this value instead of griddepcontrol.wait
    int ccl_wait_n;                // K1: further counters to wait for (batches on other slots that wrote the same output planes)
    unsigned int *ccl_wait_flag[4];
    unsigned int ccl_wait_val[4];
    const uint32_t *frame_select;  // n or NULL: when set, the global-path kernels only touch frames with a non-zero entry
    int conn8;                     // global-memory CCL kernels: 8-connectivity (the Python detector's contours) instead of 4
};

struct PreprocessParams {
    int c_thresh;      // clamp(threshold as i32, -256, 256)
    int wrap_t1;       // 0 = off.  T + 1 when `mean - c` wraps in i32 for pixels with mean >= T = c + 2^31 (c <= 255 - 2^31,
                       // detection.rs:211 in a release build): the test becomes mean < T (c_thresh is then -256: generic path)
    int blur_radius;   // 0 = input is already blurred / no blur, 2 = fused 5x5 box
    int write_blur;    // also materialise the blurred image (debug)
    int write_mask;    // write the u8 mask (0 when morphology follows and rewrites it)
    int init_labels;   // write label zeros + word-run-start parents (0 when morphology follows)
    int inverse;       // 1: 255 if px < mean - c (detection path); 0: 255 if px > mean - c
    int force_generic; // testing: never take the packed fast path
    int static_sched;  // TMA kernel: static round-robin tile schedule instead of the atomic tile counter
    int gauss_ksize;   // > 0: the TMA kernel's Gaussian variant blurs with these taps (blur_radius is ignored), k <= 15
    uint16_t gk[16];   // OpenCV's 8.8 fixed-point kernel, sums to 256
    int lookahead;      // TMA kernel: tiles the producer runs ahead of the consumers (1..stages), and the same once the
    int tail_lookahead; // tile numbers handed out are within tail_tiles of the end (set by launch_preprocess_tma)
    int tail_tiles;
    int ctas_per_sm;    // TMA kernel: resident CTAs per SM to launch (0 = default); 4 leaves room for the small CCL build
    int stages;         // TMA kernel: 0 / 2 = two stages per CTA, 3 = three (box-blur variants; used at three CTAs per SM)
    int wait_hint_ns;   // TMA kernel: suspend-time hint of mbarrier.try_wait
    int claim_ahead;    // TMA kernel: request the next tile number one tile early (hides the atomic's round trip)
    int prefetch_tiles; // TMA kernel: L2-prefetch the box of the tile this many tile numbers ahead of every claimed tile (0 = off)
    int morph_open_k;  // TMA kernel, morphology variant: 3x3 / 5x5 open and close folded into K1 (0 = none); both 0 = the
    int morph_close_k; // plain kernel
    int sparse_aux;    // flat tiles do not write their (all-zero) bit-mask words: only the fused per-frame CCL kernel, which
                       // reads nothing but the words flagged in rowflags, may follow (densify_bits() repairs it otherwise)
};

// index of tile (tx, ty) of frame f in tile_occ (4 bytes per tile)
__host__ __device__ inline size_t tile_occ_index(const BatchView &b, size_t f, int tx, int ty) {
    return (f * ((b.h + 31) / 32) + ty) * b.tiles_x + tx;
}

// byte offset of the occupancy record of (row y, tile tx) inside a frame's rowflags
__host__ __device__ inline size_t rowflag_index(int y, int tx, int tiles_x) {
    return ((size_t)(y >> 5) * tiles_x + tx) * 32 + (y & 31);
}

struct ScoreParams {
    double min_size, max_size, min_confidence;
};

#ifdef __CUDACC__
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

#endif

// Tuning knobs.  Read from the environment ONCE per process, when the first context is created (hv_create); nothing on
// the enqueue path calls getenv.  Defaults are the measured optimum ("Experiment switches").  The HV_EXP_*
// switches, which make the library skip work, exist only in builds with -DHV_EXPERIMENTS (make EXPERIMENTS=1).
struct Tunables {
    int pipeline_depth = 6;       // HV_PIPELINE_DEPTH: scratch sets in rotation = batches in flight on the device (2..8)
    int k1_ctas_per_sm = 5;       // HV_K1_CTAS_PER_SM (1..5)
    int k1_ctas_coresident = 3;   // HV_K1_CTAS_CORESIDENT: K1 CTAs per SM next to the small per-frame CCL build (1..5)
    int k1_stages_coresident = 2; // HV_K1_STAGES: TMA stages of K1 when it runs at three CTAs per SM next to the CCL kernel (2 | 3)
    int defer_depth = 2;          // HV_DEFER_DEPTH: with HV_FLAG_DEFER_TAIL, how many batches' per-frame kernels are held back (1..3)
    int k1_gauss_ctas = 4;        // HV_K1_GAUSS_CTAS (1..4): Gaussian variant, kernel sizes 9..15
    int k1_gauss_small_ctas = 4;  // HV_K1_GAUSS_SMALL_CTAS (1..4; five with a 40-register build: 0.60 instead of 0.62): Gaussian variant, kernel sizes <= 7
    int k1_lookahead = 2;         // HV_K1_LOOKAHEAD: tiles the TMA producer runs ahead
    int k1_tail_lookahead = 1;    // HV_K1_TAIL_LOOKAHEAD
    int k1_tail_rounds = 0;       // HV_K1_TAIL_ROUNDS
    int k1_prefetch = 0;          // HV_K1_PREFETCH: L2 tensor prefetch distance in tiles (0 = off)
    int k1_claim_ahead = -1;      // HV_K1_CLAIM_AHEAD: 0 / 1, -1 = when K1 runs at reduced residency (see launch_preprocess_tma)
    int k1_wait_hint_ns = 10000000;  // HV_K1_WAIT_HINT_NS
    int ccl_small_max_tiles = 32768;  // HV_CCL_SMALL_MAX_TILES: batches with more 128x32 tiles use the big per-frame CCL build
    int morph_tiles_per_sm = 2;   // HV_MORPH_TILES_PER_SM
    int phase_frame = 0;          // HV_PHASE_FRAME (with HV_FLAG_PHASE_TIMING)
    bool k1_static = false;       // HV_K1_STATIC
    bool k1_no_tma = false;       // HV_K1_NO_TMA
    bool ccl_big = false;         // HV_CCL_BIG
    bool ccl_no_tiny = false;     // HV_CCL_NO_TINY: never use the tiny per-frame CCL build
    bool no_k1_flag = false;      // HV_NO_K1_FLAG
    bool no_early_k1 = false;     // HV_NO_EARLY_K1
    bool no_pdl = false;          // HV_NO_PDL
    bool no_pdl_tail = false;     // HV_NO_PDL_TAIL
    bool no_compression = false;  // HV_NO_COMPRESSION
    bool no_fused_gauss = false;  // HV_NO_FUSED_GAUSS
    bool no_fused_morph = false;  // HV_NO_FUSED_MORPH
    bool no_morph_chain = false;  // HV_NO_MORPH_CHAIN
    bool no_side_ccl = false;     // HV_NO_SIDE_CCL: with HV_FLAG_DEFER_TAIL, keep the global-memory CCL kernels on the launching stream
    bool no_k1_morph = false;     // HV_NO_K1_MORPH: never fold 3x3 / 5x5 open+close into K1 (use the tiles kernel)
    bool exp_k1_only = false;     // HV_EXP_K1_ONLY  (-DHV_EXPERIMENTS only): K1 chain alone, NO results
    int exp_ccl_stop = 0;         // HV_EXP_CCL_STOP (with HV_EXP_CCL_NOOP): the per-frame kernel returns after phase k
    bool exp_ccl_noop = false;    // HV_EXP_CCL_NOOP (-DHV_EXPERIMENTS only): per-frame kernels launched but idle, NO results
};
const Tunables &tunables();

#define HV_CUDA_TRY(expr)                         \
    do {                                          \
        cudaError_t _e = (expr);                  \
        if (_e != cudaSuccess) return _e;         \
    } while (0)

// ---- launch wrappers (one per kernel; each returns the launch status) -----------------------------------
cudaError_t launch_gray3(const uint8_t *d_img, int n, int h, int w, size_t row_stride, size_t frame_stride,
                         uint8_t *d_gray, cudaStream_t s);
cudaError_t launch_preprocess(const BatchView &b, const PreprocessParams &p, uint32_t *bits_out, cudaStream_t s);
cudaError_t launch_preprocess_tma(const BatchView &b, const PreprocessParams &p, uint32_t *bits_out, unsigned int *sched,
                                  int num_sms, bool pdl, cudaStream_t s, bool *used);
// can open_k / close_k be folded into the TMA kernel for this batch?  (3x3 / 5x5 rectangles, total reach <= 4 px)
bool preprocess_tma_morph_supported(const BatchView &b, const PreprocessParams &p, int open_k, int close_k);
cudaError_t launch_bits_to_mask_labels(const BatchView &b, cudaStream_t s);
cudaError_t launch_rowflags_from_bits(const BatchView &b, cudaStream_t s);
cudaError_t launch_densify_bits(const BatchView &b, cudaStream_t s);
cudaError_t launch_expand_bits(const BatchView &b, cudaStream_t s);
bool morph_expand_supported(int open_k, int close_k);
cudaError_t launch_morph_expand(const BatchView &b, int open_k, int close_k, uint32_t *bits_out, uint8_t *rowflags_out,
                                uint32_t *tile_list, unsigned int *ctrl, unsigned int *chain, unsigned int scan_expected,
                                int num_sms, bool pdl, cudaStream_t s);
cudaError_t launch_morph(const BatchView &b, int open_k, int close_k, int *n_launches, cudaStream_t s);
cudaError_t launch_ccl_merge(const BatchView &b, cudaStream_t s);
cudaError_t launch_ccl_flatten(const BatchView &b, cudaStream_t s);
cudaError_t launch_ccl_scan(const BatchView &b, cudaStream_t s);
cudaError_t launch_ccl_label(const BatchView &b, cudaStream_t s);
cudaError_t launch_score(const BatchView &b, const ScoreParams &p, cudaStream_t s);
cudaError_t launch_ccl_frame(const BatchView &b, const ScoreParams &p, bool pdl, int level, cudaStream_t s);
bool ccl_frame_supported(const BatchView &b);
cudaError_t configure_ccl_frame();
cudaError_t configure_preprocess_tma();

// generic single-frame stage kernels (python-facing utilities, not the hot path)
cudaError_t launch_box_blur_generic(const uint8_t *src, int h, int w, int nch, int radius, uint8_t *dst,
                                    cudaStream_t s);
cudaError_t launch_threshold_generic(const uint8_t *src, int h, int w, int adaptive, int c_or_thr, int inverse,
                                     uint8_t *dst, cudaStream_t s);
cudaError_t launch_gaussian_blur(const uint8_t *src, int n, int h, int w, const uint16_t *k_q8_host, int ksize,
                                 uint8_t *dst, uint16_t *tmp_rows, cudaStream_t s);
cudaError_t launch_bits_from_gt127(const uint8_t *src, int n, int h, int w, int ww, uint32_t *bits, cudaStream_t s);
cudaError_t launch_visualise(const uint8_t *mask, int h, int w, const hv_center *d_centers, int n_centers,
                             uint8_t *out_hw3, cudaStream_t s);
cudaError_t launch_gray_first3(const uint8_t *d_img, int h, int w, int c, uint8_t *d_gray, cudaStream_t s);
cudaError_t launch_collect_centers(const BatchView &b, uint32_t min_area, hv_center *d_centers, uint32_t *d_count,
                                   int cap, cudaStream_t s);
cudaError_t launch_bayer(const uint8_t *d_src, int n, int h, int w, int pattern, bool to_gray, uint8_t *d_dst,
                         cudaStream_t s);
cudaError_t launch_yuyv(const uint8_t *d_src, int n, int h, int w, bool to_gray, uint8_t *d_dst, cudaStream_t s);
cudaError_t launch_gray_bgr_cv(const uint8_t *d_img, int h, int w, uint8_t *d_gray, cudaStream_t s);
cudaError_t launch_adaptive_gaussian(const uint8_t *d_src, int h, int w, const float *k_host, int ksize, int idelta,
                                     float *d_rows, uint8_t *d_mask, cudaStream_t s);
cudaError_t launch_invert_bits(const uint32_t *in, uint32_t *out, int h, int ww, int w, cudaStream_t s);
cudaError_t launch_first_pixel(const int32_t *labels, int h, int w, uint32_t *first, cudaStream_t s);
cudaError_t launch_label_above(const uint32_t *first, int n, int w, const int32_t *other, int32_t *above, cudaStream_t s);
// out: n_ext records of 14 x u64 {cnt_in, cnt_out, gray_in, gray_out, ch_in[3], ch_out[3]} (k_pydet.cu RegionSums)
cudaError_t launch_region_sums(const int32_t *ext, int n_ext, const hv_blob *comps8, const int32_t *l8, const int32_t *l4,
                               const int32_t *root8, const int32_t *root4, int h, int w, const uint8_t *gray, const uint8_t *bgr,
                               void *out, cudaStream_t s);
// out: n_ext records {double a00, a10, a01; u32 chain_len, reserved} (k_pydet.cu TraceOut)
cudaError_t launch_trace_contours(const int32_t *ext, int n_ext, const uint32_t *first8, const int32_t *l8, const int32_t *l4,
                                  const int32_t *root8, const int32_t *root4, int h, int w, void *out, cudaStream_t s);
cudaError_t launch_overlays(const hv_overlay *d_items, int n, int h, int w, uint8_t *d_img, unsigned int *d_owner, cudaStream_t s);
cudaError_t launch_collect_contours(const BatchView &b, double min_area, double max_area, hv_contour *d_out,
                                    uint32_t *d_count, int cap, cudaStream_t s);

}  // namespace hv
