/*
 * heimdall_cuda.h -- C ABI of the B200 (sm_100a) backend for heimdall-vision's contamination-inspection hot path.
 *
 * This is the drop-in boundary: exactly what a Rust `heimdall-cuda` crate (build.rs + extern "C" block, see
 * rust/heimdall-cuda/ and INTEGRATION.md) or any other FFI (ctypes, cgo, JNI) binds.  Plain pointers and sizes,
 * fixed-width integers, POD structs; no C++ or torch types.  Every entry point cites the reference interface it
 * replaces (paths relative to the reference repository root).
 *
 * Threading: an hv_ctx is bound to one CUDA device and is NOT thread-safe; use one context per (thread, device).
 * The library keeps no global mutable state.  There is no CPU fallback: without a usable CUDA device hv_create fails.
 *
 * Results are bit-exact with oracle/hv_oracle.c, the line-for-line C restatement of the reference's Rust CPU path
 * (masks, labels in raster-first-pixel order, blob statistics, defect list order, confidences as IEEE f64, reject
 * decision).  The Rust path itself cannot be built in this environment and the reference holds no golden vectors for
 * it, so parity with the Rust binary is pinned only through that restatement ("Oracle and parity status").
 */
#ifndef HEIMDALL_CUDA_H
#define HEIMDALL_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define HV_API __declspec(dllexport)
#else
#define HV_API __attribute__((visibility("default")))
#endif

#define HV_ABI_VERSION 2

typedef struct hv_ctx hv_ctx;
typedef int32_t hv_status;

/* Status codes. The message texts returned by hv_last_error() reproduce the reference's error enums:
 * DetectionError (rust/heimdall-core/src/detection.rs:20-27), ProcessingError (processing.rs:11-21). */
enum {
    HV_OK = 0,
    HV_ERR_INVALID_DIMENSIONS = -1, /* "Invalid image dimensions: expected 3D array" (detection.rs:159) */
    HV_ERR_INVALID_ARGUMENT = -2,
    HV_ERR_CUDA = -3,               /* CUDA runtime failure; text in hv_last_error() */
    HV_ERR_CAPACITY = -4,           /* more blobs/defects than the configured capacity; never truncates silently */
    HV_ERR_NO_DEVICE = -5,          /* no CUDA device / wrong architecture: there is no CPU fallback */
    HV_ERR_UNSUPPORTED = -6,        /* "Unsupported pipeline type: ..." (lib.rs:80-84) and similar */
    HV_ERR_CHANNELS = -7,           /* stage needs a 1-channel image (processing.rs:119-121, detection.rs:50-52) */
    HV_ERR_BAD_TICKET = -8
};

/* Context sizing. Zero means "default". */
typedef struct {
    int32_t max_batch;              /* accepted for forward compatibility and ignored: scratch always grows on demand */
    int32_t max_height, max_width;  /* (likewise) */
    int32_t max_blobs_per_frame;    /* stats-table rows per frame; 0 -> min(h*w/2+1, 131072); h*w/2+1 is the 4-connectivity
                                       maximum: ask for it explicitly when frames can be that dense (40 B per row) */
    int32_t max_defects_per_frame;  /* device-side defect slots per frame; 0 -> 256 */
    int32_t num_slots;              /* in-flight batches for hv_submit/hv_wait; 0 -> 3 */
    int32_t flags;                  /* HV_FLAG_* */
    int32_t reserved;
} hv_config;

#define HV_FLAG_PROFILE 2u    /* record a CUDA event pair around every kernel (hv_profile_get) */
#define HV_FLAG_FORCE_GENERIC 8u /* testing: K1 never takes the packed fast path nor the flat-tile skip */
#define HV_FLAG_GLOBAL_CCL 16u   /* always use the global-memory CCL kernels (K2..K6), never the fused per-frame kernel */
#define HV_FLAG_PHASE_TIMING 32u /* fused CCL kernel records per-phase timestamps of frame 0 (hv_debug_phase_times) */
#define HV_FLAG_KEEP_BLUR 4u  /* materialise the blurred intermediate for hv_fetch_debug (disables the flat-tile skip) */
/* hv_enqueue_device holds the last kernel of a batch (the per-frame labelling + scoring kernel) and the read-back of its
 * results back for two calls: call i enqueues that kernel of batch i - 2, then the preprocess kernel of batch i.  Whatever
 * the caller enqueues on the stream between two batches (an event, the upload of the next frames) then sits behind a
 * preprocess kernel and in front of a per-frame kernel whose input has long been complete, instead of between a per-frame
 * kernel and the next batch's preprocess kernel, whose overlap is what keeps the device busy.  For callers that consume
 * the records (tickets), not the label plane in stream order: with this flag the label plane and the results of the last
 * two batches are complete only after hv_flush / hv_fetch_ticket / hv_stats_get (or two more hv_enqueue_device calls).
 * The mask plane is written by the preprocess kernel and is complete in stream order either way.  The held kernels read
 * the batch's input frames (the contrast probe of the scoring): the frames must stay unchanged until the batch has been
 * fetched or flushed.  Batches that take the global-memory CCL kernels (dense frames) run those on a stream of the
 * context's own, beside the next batch's preprocess kernel, under the same contract.  Rotate at least
 * hv_pipeline_depth() sets of output planes with it: a batch that writes planes one of the held batches wrote retires that
 * batch first, which puts its kernel on the stream and waits for it. */
#define HV_FLAG_DEFER_TAIL 64u

/* Blur selection for the preprocess stage. */
enum {
    HV_BLUR_BOX = 0,      /* detection.rs:162-182: (2r+1)^2 box mean, floor division, interior only; r = blur_ksize/2 */
    HV_BLUR_GAUSSIAN = 1, /* cv2.GaussianBlur(k,k,sigma) semantics (heimdall/detectors/contamination_detector.py:66) */
    HV_BLUR_NONE = 2
};

/* Per-call parameters. hv_params_default() fills the reference defaults
 * (rust/heimdall-core/src/lib.rs:106-108: 10.0 / 3000.0 / 25.0; detection.rs:163 radius 2, :185 window 11). */
typedef struct {
    double min_size;       /* inclusive lower area bound (detection.rs:250) */
    double max_size;       /* inclusive upper area bound */
    double threshold;      /* `c = threshold as i32` (detection.rs:186) */
    double min_confidence; /* 0.3 (detection.rs:298) */
    double gauss_sigma;    /* HV_BLUR_GAUSSIAN only; <=0 -> OpenCV's automatic sigma */
    int32_t blur_mode;     /* HV_BLUR_* */
    int32_t blur_ksize;    /* 5 */
    int32_t morph_open_k;  /* 0 = off (the Rust path has no morphology); k>0: rect kxk MORPH_OPEN on the mask
                              (contamination_detector.py:81-84) */
    int32_t morph_close_k; /* 0 = off; k>0: rect kxk MORPH_CLOSE after the open (contamination_detector.py:87) */
    int32_t reserved[4];
} hv_params;

/* One defect. Mirrors `Defect` (detection.rs:12-18): position = (row, col), size = area as f64.  The extra integer
 * fields are what the reference computes on the way (bbox for the shape score, 1-based component label). */
typedef struct {
    int32_t y, x;
    double size;
    double confidence;
    int32_t ymin, xmin, ymax, xmax;
    uint32_t label;
    uint32_t frame; /* index of the frame inside the batch */
} hv_defect;

/* Per-frame summary. `rejected` is the reject decision: n_defects > 0 (heimdall/inspection/base_inspector.py:40-42). */
typedef struct {
    uint32_t n_components; /* all 4-connected components of the mask, before any filter */
    uint32_t n_defects;
    uint32_t defects_offset; /* first defect of this frame in the caller's defects array */
    uint32_t rejected;
    uint32_t fg_pixels; /* mask pixels == 255 */
    int32_t status;     /* HV_OK or HV_ERR_CAPACITY for this frame */
} hv_frame_result;

/* Per-blob integer statistics, canonical (raster-first-pixel) order; row k describes label k+1. */
typedef struct {
    uint32_t area;
    uint32_t ymin, ymax, xmin, xmax;
    uint32_t reserved;
    uint64_t sum_y, sum_x;
} hv_blob;

/* Optional host-side copies of the intermediates for parity checks (any pointer may be NULL).
 * gray/blur/mask: n*h*w u8; labels: n*h*w i32 (0 = background, k = k-th component in raster-first-pixel order);
 * blobs: n * blobs_stride rows. */
typedef struct {
    uint8_t *gray;
    uint8_t *blur;
    uint8_t *mask;
    int32_t *labels;
    hv_blob *blobs;
    size_t blobs_stride;
} hv_debug_outputs;

/* Line-level statistics accumulated on the device across calls (dashboard.py:38-46,483-500;
 * heimdall/core/system.py:168-175).  All fields u64 so that the whole struct can be all-reduced with one
 * ncclAllReduce(ncclUint64, ncclSum) over hv_stats_device_ptr(). */
#define HV_STATS_AREA_BINS 16
typedef struct {
    uint64_t frames_inspected;
    uint64_t frames_rejected;
    uint64_t total_defects;
    uint64_t total_components;
    uint64_t total_defect_area;
    uint64_t total_fg_pixels;
    uint64_t area_hist[HV_STATS_AREA_BINS]; /* bin = floor(log2(area)), clamped */
    uint64_t capacity_errors;
    uint64_t reserved[9];
} hv_line_stats; /* 32 x u64 = 256 bytes */

/* Kernel indices for hv_profile_get(). */
enum {
    HV_K_GRAY = 0,       /* A1 gray (C==3 only) */
    HV_K_PREPROCESS = 1, /* A2+A3 fused blur + adaptive threshold + mask/bitmask/label-zero store */
    HV_K_MORPH = 2,      /* A8 open/close on the bit-packed mask (when enabled) */
    HV_K_CCL_MERGE = 3,  /* A4 union-find merge of word-runs */
    HV_K_CCL_FLATTEN = 4,
    HV_K_CCL_SCAN = 5,
    HV_K_CCL_LABEL = 6,  /* A4 canonical relabel + A5 blob statistics */
    HV_K_SCORE = 7,      /* A5b + A6 */
    HV_K_CCL_FRAME = 8,  /* A4 + A5 + A5b + A6 fused, one CTA per frame, union-find in shared memory (sparse masks) */
    HV_K_COUNT = 9
};

/* ---- library ------------------------------------------------------------------------------------------- */
HV_API int32_t hv_abi_version(void);
HV_API const char *hv_version(void);
HV_API const char *hv_status_string(hv_status s);
HV_API int32_t hv_device_count(void);
HV_API void hv_params_default(hv_params *p);
HV_API void hv_config_default(hv_config *c);

/* ---- context ------------------------------------------------------------------------------------------- */
HV_API hv_status hv_create(int32_t device, const hv_config *cfg, hv_ctx **out);
HV_API void hv_destroy(hv_ctx *ctx);
/* Text of the last failure on this context ("" if none). With ctx == NULL: last hv_create failure of this thread. */
HV_API const char *hv_last_error(const hv_ctx *ctx);
/* enable != 0: run the synchronous / device-resident entry points on the caller's CUDA stream (a cudaStream_t; NULL
 * is the legacy default stream, which is what torch.cuda.current_stream().cuda_stream returns by default).
 * enable == 0: back to the context's own non-blocking stream. */
HV_API hv_status hv_set_stream(hv_ctx *ctx, void *cuda_stream, int32_t enable);

/* Pinned host staging buffers (cudaHostAlloc). Frames handed to hv_detect_batch / hv_submit from such a buffer are
 * copied to the device without an intermediate host copy. */
HV_API void *hv_host_alloc(hv_ctx *ctx, size_t bytes);
HV_API void hv_host_free(hv_ctx *ctx, void *p);

/* Device buffers for the device-resident entry points (hv_enqueue_device: d_mask / d_labels).  With
 * HV_ALLOC_COMPRESSIBLE the buffer is created through the virtual-memory API with L2 compute-data compression
 * (CU_MEM_ALLOCATION_COMP_GENERIC): the mask and label planes of inspection frames are almost entirely zero, the L2
 * compresses such lines on their way to HBM and the write stream of the preprocess kernel -- 5 of the 6 algorithmic
 * bytes per pixel -- takes less DRAM time.  Reads (kernels, cudaMemcpy) see ordinary memory.  If the device does not
 * support compression the flag is ignored; *compressed_out (optional) reports what was obtained.  No counterpart in
 * the reference (host-only code). */
/* Number of scratch sets the device-resident entry point rotates through = how many batches may be in flight on the
 * device at once.  A caller that passes its own d_mask / d_labels planes to hv_enqueue_device should rotate this many
 * sets of them: a set that is reused earlier is still correct (the preprocess kernel of the new batch waits on the
 * device for the batch that last wrote it), it only costs overlap.  The input frames of a batch must stay untouched
 * until the batch has completed (stream-ordered work on the same stream is ordered behind it as usual). */
HV_API int32_t hv_pipeline_depth(void);

enum { HV_ALLOC_COMPRESSIBLE = 1 };
HV_API hv_status hv_device_alloc(hv_ctx *ctx, size_t bytes, uint32_t flags, void **d_ptr, int32_t *compressed_out);
HV_API hv_status hv_device_free(hv_ctx *ctx, void *d_ptr);
/* Blocking copies from / to such a buffer (ordered after the work enqueued so far on the context's stream). */
HV_API hv_status hv_device_read(hv_ctx *ctx, void *host_dst, const void *d_src, size_t bytes);
HV_API hv_status hv_device_write(hv_ctx *ctx, void *d_dst, const void *host_src, size_t bytes);

/* ---- the hot path --------------------------------------------------------------------------------------
 * Replaces heimdall_core.detect_contamination (rust/heimdall-core/src/lib.rs:95-143) ->
 * detection::detect_contamination (detection.rs:127-317), for a batch of n frames.
 *
 * frames: HOST memory, n frames of h x w x c u8 (c = 1 or 3), element (f,y,x,ch) at
 *         frames[f*frame_stride + y*row_stride + x*c + ch]; strides in bytes, 0 = tightly packed.
 * results: n entries. defects: up to defects_cap entries, frame-major, discovery order inside a frame.
 * Returns HV_OK, or HV_ERR_CAPACITY if any frame overflowed (per-frame status says which; the other frames are valid).
 */
HV_API hv_status hv_detect_batch(hv_ctx *ctx, const uint8_t *frames, int32_t n, int32_t h, int32_t w, int32_t c,
                                 size_t row_stride, size_t frame_stride, const hv_params *params,
                                 hv_frame_result *results, hv_defect *defects, size_t defects_cap,
                                 size_t *n_defects_total, const hv_debug_outputs *debug);

/* Same, frames already resident in DEVICE memory (c = 1 or 3, tightly packed rows unless strides given).
 * d_mask (n*h*w u8) and d_labels (n*h*w i32) are optional caller-owned DEVICE outputs; when NULL the context's
 * scratch is used.  results/defects are HOST memory. */
HV_API hv_status hv_detect_batch_device(hv_ctx *ctx, const uint8_t *d_frames, int32_t n, int32_t h, int32_t w,
                                        int32_t c, size_t row_stride, size_t frame_stride, const hv_params *params,
                                        uint8_t *d_mask, int32_t *d_labels, hv_frame_result *results,
                                        hv_defect *defects, size_t defects_cap, size_t *n_defects_total);

/* Device-resident, asynchronous, streaming: enqueue only (no host synchronisation on the launching stream).  The
 * batch's results (per-frame records, frame flags, defect table) are copied to page-locked host memory by the copy
 * engine on the context's own copy stream as soon as the batch's last kernel has finished (the stream waits for the
 * slot's device-side completion counter, so nothing is inserted between the kernels of consecutive batches), and
 * *ticket (optional) names the batch for hv_fetch_ticket.  The context keeps the last hv_pipeline_depth() batches:
 * enqueueing one more first retires the oldest -- waits for its read-back and, if the per-frame CCL kernel flagged
 * frames it could not hold, finishes them with the global-memory kernels -- so every batch is completed and counted
 * in the line statistics whether or not it is ever fetched, and the host runs at most hv_pipeline_depth() batches
 * ahead of the device.  The launch order of the kernels is tied to device-side counters whose expected values are
 * kernel arguments: a stream that is being captured into a CUDA graph is refused (HV_ERR_UNSUPPORTED). */
HV_API hv_status hv_enqueue_device(hv_ctx *ctx, const uint8_t *d_frames, int32_t n, int32_t h, int32_t w, int32_t c,
                                   size_t row_stride, size_t frame_stride, const hv_params *params, uint8_t *d_mask,
                                   int32_t *d_labels, int64_t *ticket);
/* HV_FLAG_DEFER_TAIL: puts the kernels held back (and their read-backs) onto the stream now and makes the stream wait for
 * the kernels that went onto streams of the context's own: after it, the planes of every enqueued batch are complete in
 * stream order (frames the per-frame kernel flagged are finished when their batch is fetched or retired, as without the
 * flag).  A no-op otherwise. */
HV_API hv_status hv_flush(hv_ctx *ctx);
/* Results of a batch enqueued with hv_enqueue_device, by ticket: blocks until its read-back has arrived.
 * HV_ERR_BAD_TICKET once the batch's scratch set has been reused (more than hv_pipeline_depth() - 1 batches later). */
HV_API hv_status hv_fetch_ticket(hv_ctx *ctx, int64_t ticket, hv_frame_result *results, hv_defect *defects,
                                 size_t defects_cap, size_t *n_defects_total);
/* Results of the most recently enqueued batch. */
HV_API hv_status hv_fetch_results(hv_ctx *ctx, hv_frame_result *results, hv_defect *defects, size_t defects_cap,
                                  size_t *n_defects_total);
/* Copy intermediates of the last enqueued batch to the host (parity checks). */
HV_API hv_status hv_fetch_debug(hv_ctx *ctx, const hv_debug_outputs *debug);

/* Pipelined host-fed form (camera streams): hv_submit copies/points at the host frames, enqueues H2D + kernels +
 * D2H on one of the context's slots and returns a ticket; hv_wait blocks for that ticket and fills the outputs.
 * Up to num_slots tickets may be in flight; the frames buffer must stay valid until hv_wait returns. */
HV_API hv_status hv_submit(hv_ctx *ctx, const uint8_t *frames, int32_t n, int32_t h, int32_t w, int32_t c,
                           size_t row_stride, size_t frame_stride, const hv_params *params, int64_t *ticket);
HV_API hv_status hv_wait(hv_ctx *ctx, int64_t ticket, hv_frame_result *results, hv_defect *defects,
                         size_t defects_cap, size_t *n_defects_total);

/* ---- frame feed (next-row N1) ----------------------------------------------------------------------------
 * Camera frames as heimdall-camera delivers them: `CameraFrame` (rust/heimdall-camera/src/lib.rs:111-132) with
 * `PixelFormat` (lib.rs:34-47).  The raw bytes cross PCIe once (1 B/px for Mono8 / Bayer, 2 B/px for YUV422) and are
 * converted on the device. */
enum {
    HV_PIX_MONO8 = 0,
    HV_PIX_MONO16 = 1,
    HV_PIX_RGB8 = 2,
    HV_PIX_BGR8 = 3,
    HV_PIX_RGBA8 = 4,
    HV_PIX_BGRA8 = 5,
    HV_PIX_YUV422 = 6,
    HV_PIX_YUV422_PACKED = 7,
    HV_PIX_BAYER_RG8 = 8,
    HV_PIX_BAYER_GB8 = 9,
    HV_PIX_BAYER_GR8 = 10,
    HV_PIX_BAYER_BG8 = 11
};
typedef struct {
    const uint8_t *data; /* raw frame bytes, host memory (pinned memory from hv_host_alloc avoids a staging copy) */
    size_t size;         /* bytes in data */
    uint32_t width, height;
    int32_t pixel_format; /* HV_PIX_* */
    uint32_t camera;      /* index of the camera / stream the frame came from (free for the caller) */
    uint64_t frame_id;
    uint64_t timestamp_ns;
} hv_camera_frame;

/* Channels of the image a frame converts to, or 0 when the format has no conversion: to_ndarray (lib.rs:260-278) gives
 * 1 / 3 / 4 channels for Mono8 / RGB8,BGR8 / RGBA8,BGRA8 (bytes unchanged); the cv2.cvtColor conversions named by
 * to_opencv_mat (lib.rs:226-252) give 3 (RGB) for the Bayer and YUV422 formats; Mono16 has none. */
HV_API int32_t hv_frame_channels(int32_t pixel_format);
/* One frame -> (height, width, channels) u8 image in host memory (the array the reference hands to the detector). */
HV_API hv_status hv_convert_frame(hv_ctx *ctx, const hv_camera_frame *frame, uint8_t *out_hwc, int32_t *out_channels);
/* hv_submit for n camera frames of identical geometry and pixel format: staging -> H2D -> conversion -> detect.
 * Formats that convert to 1 or 3 channels are accepted (detection.rs:138-160 rejects everything else:
 * HV_ERR_INVALID_DIMENSIONS); results are those of detect_contamination on the converted images.  Collect with
 * hv_wait.  The frames' data must stay valid until hv_wait returns. */
HV_API hv_status hv_submit_frames(hv_ctx *ctx, const hv_camera_frame *frames, int32_t n, const hv_params *params,
                                  int64_t *ticket);

/* ---- multi-camera frame sets (next-row N2) ------------------------------------------------------------------
 * `FrameSet` (rust/heimdall-gige/src/frame.rs:127-185): the frames all cameras of the line acquired for one trigger,
 * built by GigESystem::acquire_frames (rust/heimdall-gige/src/lib.rs:529-648; at most 4 Mono8 cameras, lib.rs:206-234)
 * under one of the three sync modes of rust/heimdall-gige/src/sync.rs:18-27.  The batcher takes frames as they arrive,
 * groups them into sets (by trigger number, or by arrival order in Freerun), and submits `sets_per_batch` complete
 * sets at a time to the detector: frame order inside a batch is set-major, camera-minor. */
typedef struct hv_frameset hv_frameset;
enum { HV_SYNC_FREERUN = 0, HV_SYNC_SOFTWARE = 1, HV_SYNC_HARDWARE = 2 };
typedef struct {
    int32_t n_cameras;        /* cameras per set, 1..16 (the reference configures up to 4) */
    int32_t sets_per_batch;   /* complete sets per detector batch */
    int32_t sync_mode;        /* HV_SYNC_*: Freerun matches the k-th frame of every camera, the others match frame_id */
    int32_t max_pending_sets; /* incomplete sets kept while cameras are late; the oldest is dropped beyond (0 -> 8) */
    int32_t flags;            /* HV_FRAMESET_* */
    int32_t reserved;
} hv_frameset_config;
/* Frames whose data already lies in page-locked host memory (hv_host_alloc / cudaHostAlloc / cudaHostRegister) are
 * referenced instead of copied into a slab: the caller keeps such a buffer valid and unchanged until hv_frameset_wait has
 * returned for the batch the frame went into (or the batcher reports the frame's set as dropped).  Pageable frames are
 * still copied.  Saves one pass over every frame on the host: 8 GB/s per thread for the copy against 55 GB/s of PCIe. */
#define HV_FRAMESET_ZERO_COPY 1
typedef struct {
    uint64_t frames_pushed;
    uint64_t sets_completed;
    uint64_t sets_dropped;   /* incomplete sets given up (a camera never delivered: lib.rs:590-606 fails the set) */
    uint64_t frames_dropped; /* frames of dropped sets + frames that arrived after their set was gone */
    uint64_t duplicates;     /* second delivery of the same (set, camera) */
    uint64_t batches_submitted;
    uint64_t max_skew_ns;    /* largest timestamp spread inside a completed set */
    uint64_t sets_pending;   /* open + complete-but-not-yet-batched sets right now */
} hv_frameset_stats;
HV_API hv_status hv_frameset_create(hv_ctx *ctx, const hv_frameset_config *cfg, hv_frameset **out);
HV_API void hv_frameset_destroy(hv_frameset *fs);
HV_API const char *hv_frameset_last_error(const hv_frameset *fs);
/* One frame (frame->camera = camera index, frame->frame_id = trigger number).  Its bytes are copied into a page-locked
 * slab before the call returns.  *ticket = the batch's ticket when this frame completed a batch, else 0. */
HV_API hv_status hv_frameset_push(hv_frameset *fs, const hv_camera_frame *frame, const hv_params *params,
                                  int64_t *ticket);
/* Retry the submission of a batch that hv_frameset_push could not hand over (it returned the detector's status, e.g.
 * HV_ERR_CAPACITY while all num_slots tickets were in flight; the complete sets stayed queued).  *ticket = 0 when there
 * is nothing to submit.  Complete sets queue up to max_pending_sets beyond one batch; past that the oldest is dropped
 * and counted in sets_dropped / frames_dropped. */
HV_API hv_status hv_frameset_flush(hv_frameset *fs, const hv_params *params, int64_t *ticket);
/* Set ids (trigger numbers / Freerun indices) of a submitted batch, ascending. */
HV_API hv_status hv_frameset_batch_ids(hv_frameset *fs, int64_t ticket, uint64_t *set_ids, int32_t cap, int32_t *n_sets);
/* hv_wait for a batch submitted by the batcher (also recycles its slabs): n_cameras * sets_per_batch results. */
HV_API hv_status hv_frameset_wait(hv_frameset *fs, int64_t ticket, hv_frame_result *results, hv_defect *defects,
                                  size_t defects_cap, size_t *n_defects_total);
HV_API hv_status hv_frameset_get_stats(const hv_frameset *fs, hv_frameset_stats *out);

/* ---- stage entry points (single frame, host memory) ----------------------------------------------------
 * heimdall_core.processing.preprocess_image (processing.rs:30-101): out has (grayscale ? 1 : c) channels;
 * blur_size <= 0 -> no blur.  grayscale != 0 requires c >= 3 (the reference indexes channels 1 and 2). */
HV_API hv_status hv_preprocess_image(hv_ctx *ctx, const uint8_t *img, int32_t h, int32_t w, int32_t c,
                                     int32_t grayscale, int32_t blur_size, uint8_t *out);
/* heimdall_core.processing.apply_threshold (processing.rs:104-185): c must be 1. */
HV_API hv_status hv_apply_threshold(hv_ctx *ctx, const uint8_t *img, int32_t h, int32_t w, int32_t c,
                                    uint8_t threshold_value, int32_t adaptive, int32_t inverse, uint8_t *out);

/* Rect morphology on a binary {0, 255} mask (`> 127` counts as set): cv2.morphologyEx(MORPH_OPEN, k_open x k_open) followed
 * by cv2.morphologyEx(MORPH_CLOSE, k_close x k_close), MORPH_RECT, OpenCV's default border; 0 skips an operation.  The
 * stage the reference's Python detector and pipeline run on their masks (heimdall/detectors/contamination_detector.py:
 * 81-87; heimdall/core/pipeline.py:290-332 `MorphologyStage`).  Kernel sizes up to 31. */
HV_API hv_status hv_morphology(hv_ctx *ctx, const uint8_t *mask, int32_t h, int32_t w, int32_t open_k, int32_t close_k,
                               uint8_t *out);

/* heimdall_core.detection.find_contours (detection.rs:36-124): foreground is `> 127`; blobs with
 * min_area <= area <= max_area in discovery order. labels (optional, h*w i32) receives the full label map. */
typedef struct {
    int32_t y, x;
    double area;
    uint64_t pixel_count;
    uint32_t label;
    uint32_t reserved;
} hv_contour;
HV_API hv_status hv_find_contours(hv_ctx *ctx, const uint8_t *img, int32_t h, int32_t w, int32_t c, double min_area,
                                  double max_area, hv_contour *contours, size_t cap, size_t *n_contours,
                                  int32_t *labels);

/* heimdall_core.process_image (lib.rs:42-92). pipeline: 0 = "basic" (processing.rs:188-249),
 * 1 = "contamination" (processing.rs:252-404). out_hw3: h*w*3 u8 visualisation. contours (cy, cx, 0.75) only for
 * pipeline 1. */
enum { HV_PIPELINE_BASIC = 0, HV_PIPELINE_CONTAMINATION = 1 };
typedef struct {
    int32_t y, x;
    double confidence;
} hv_center;
HV_API hv_status hv_process_image(hv_ctx *ctx, const uint8_t *img, int32_t h, int32_t w, int32_t c,
                                  int32_t pipeline, uint8_t *out_hw3, hv_center *contours, size_t cap,
                                  size_t *n_contours);

/* ---- Python-detector parity mode (next-row N3) ------------------------------------------------------------------
 * The stages of the reference's Python detector, `ContaminationDetector.detect` (heimdall/detectors/
 * contamination_detector.py:58-90; the fallback behind heimdall/rust_bridge.py:139-161), with OpenCV's arithmetic:
 *   gray     cv2.cvtColor(BGR2GRAY) for c == 3 (15-bit fixed point), the image itself for c == 1          (:58-62)
 *   blurred  cv2.GaussianBlur(gray, (blur_ksize, blur_ksize), 0)                                             (:66)
 *   binary   cv2.adaptiveThreshold(blurred, 255, ADAPTIVE_THRESH_GAUSSIAN_C, THRESH_BINARY_INV, block_size, C) (:70-77)
 *            then MORPH_OPEN and MORPH_CLOSE with k x k rectangles                                             (:81-87)
 *   comps    the 8-connected components of `binary` in raster order of their first pixel (area in pixels, bounding box,
 *            coordinate sums) and their label plane: the regions whose outer borders cv2.findContours(RETR_EXTERNAL)
 *            (:90) traces -- RETR_EXTERNAL additionally drops components that lie inside a hole of another one.
 * gray / blurred / binary are bit-exact with opencv-python 4.13 on the committed golden vectors (the float32 mean of
 * adaptiveThreshold follows OpenCV's AVX2/FMA summation order, see k_pydet.cu).  The contour quantities and the scores:
 * hv_python_detect below.  Any output pointer may be NULL. */
typedef struct {
    double contrast_threshold; /* C (contamination_detector.py:36 default 25) */
    int32_t blur_ksize;        /* 5 */
    int32_t block_size;        /* 11: odd, 3..31 */
    int32_t morph_open_k;      /* 3 */
    int32_t morph_close_k;     /* 3 */
    int32_t reserved[4];
} hv_pydet_params;
HV_API void hv_pydet_params_default(hv_pydet_params *p);
HV_API hv_status hv_python_detector_stages(hv_ctx *ctx, const uint8_t *img, int32_t h, int32_t w, int32_t c,
                                           const hv_pydet_params *params, uint8_t *gray, uint8_t *blurred,
                                           uint8_t *binary, int32_t *labels8, hv_blob *comps, size_t cap,
                                           size_t *n_comps);

/* The whole `ContaminationDetector.detect` (contamination_detector.py:44-216) on top of those stages: the external contours'
 * polygon area (cv2.contourArea), bounding rectangle, centre from the polygon moments (int(m10/m00), int(m01/m00)), the
 * filled-contour masks behind the intensity and colour scores, confidence = 0.5 intensity + 0.2 shape + 0.3 colour, the size
 * and confidence filters, in cv2.findContours' order.  How: the outer border of an 8-connected component with everything it
 * encloses is traced on the device (Moore tracing = cv2's CHAIN_APPROX_NONE chain; the moment accumulators are integer-valued
 * doubles, so area and moments are bit-identical with cv2's); cv2.drawContours(..., -1) of that contour is the component plus
 * the regions of its complement that do not reach the image border (checked against cv2 on thousands of random contours,
 * tests/golden/make_golden.py); RETR_EXTERNAL's nesting rule comes from the 4-connected components of the background.
 * position = (x, y) as the Python `Defect` has it.  Not returned: metadata["contour"] (the CHAIN_APPROX_SIMPLE point list). */
typedef struct {
    double min_size, max_size;   /* contour-area limits (contamination_detector.py:26-30: 10, 3000) */
    double min_confidence;       /* :36 default 0.25; heimdall/rust_bridge.py:148 passes 0.3 */
    int32_t use_color;           /* :38 default 1; only with a 3-channel image */
    int32_t reserved[3];
} hv_pydet_score_params;
typedef struct {
    int32_t x, y;                /* Defect.position = (cx, cy) */
    double size;                 /* cv2.contourArea */
    double confidence;
    double intensity_diff, shape_score, color_score; /* metadata */
    int32_t bx, by, bw, bh;      /* metadata["bounding_box"] = cv2.boundingRect */
    uint32_t label8;             /* 1-based label of the component in the labels8 plane of hv_python_detector_stages */
    uint32_t chain_len;          /* points of the CHAIN_APPROX_NONE contour */
} hv_pydefect;
HV_API void hv_pydet_score_params_default(hv_pydet_score_params *p);
HV_API hv_status hv_python_detect(hv_ctx *ctx, const uint8_t *img, int32_t h, int32_t w, int32_t c,
                                  const hv_pydet_params *stage_params, const hv_pydet_score_params *score_params,
                                  hv_pydefect *defects, size_t cap, size_t *n_defects, size_t *n_contours);

/* ---- result side (next-row N4) ---------------------------------------------------------------------------------
 * What the reference does with a frame's defect list once the detector has returned.
 *
 * hv_export_results: per-frame records in the shape of `InspectionResult` (heimdall/inspection/base_inspector.py:11-64:
 * inspection_id, timestamp, success, has_defects = len(defects) > 0, defect_count, processing_time) from the results of a
 * batch, and the dashboard's running statistics (dashboard.py:38-46 `processing_stats`, updated per image at :483-500:
 * total_images, total_defects, an exponential moving average of the processing time in milliseconds -- first sample
 * taken as is, then 0.9 * avg + 0.1 * t * 1000 -- and defect_rate = total_defects / total_images * 100).  Host-side
 * marshalling, no device work. */
typedef struct {
    uint64_t sequence;        /* running inspection number: inspection_id = "<inspector_id>_<sequence>" (base_inspector.py:109) */
    double timestamp;         /* seconds since the epoch, as given */
    double processing_time;   /* seconds, as given (per frame) */
    uint32_t success;         /* 1 unless the frame's status reports an error */
    uint32_t has_defects;     /* the reject predicate (base_inspector.py:40-42) */
    uint32_t defect_count;
    uint32_t defects_offset;  /* first defect of the frame in the batch's defect array */
} hv_inspection_record;
typedef struct {
    uint64_t total_images;
    uint64_t total_defects;
    double avg_processing_time_ms;
    double defect_rate;       /* per cent */
    double start_time;        /* untouched by the library (dashboard.py:45) */
} hv_dashboard_stats;
HV_API hv_status hv_export_results(const hv_frame_result *results, int32_t n, double timestamp, double processing_time,
                                   uint64_t first_sequence, hv_inspection_record *records, hv_dashboard_stats *stats);

/* Overlays on an (h, w, 3) u8 image in host memory, drawn on the device in list order (a later item paints over an earlier
 * one).  Kinds:
 *   HV_OVERLAY_CROSS  the 7-pixel cross of process_image("contamination") at (y, x): vertical bar then horizontal bar,
 *                     clipped at the image border (rust/heimdall-core/src/processing.rs:371-401);
 *   HV_OVERLAY_BOX    cv2.rectangle(img, (x, y), (x1, y1), color, 1): the one-pixel outline of the bounding box
 *                     (the box of heimdall/detectors/contamination_detector.py:249 with thickness 1);
 *   HV_OVERLAY_MARKER the dashboard's defect marker cv2.circle(img, (x, y), 10, color, 2) (dashboard.py:462; the same call
 *                     at base_inspector.py:186): OpenCV's footprint of that circle, identical to cv2 wherever the marker
 *                     lies inside the image and the clipped footprint otherwise (OpenCV re-rasterises clipped arcs). */
enum { HV_OVERLAY_CROSS = 0, HV_OVERLAY_BOX = 1, HV_OVERLAY_MARKER = 2 };
typedef struct {
    int32_t kind;
    int32_t y, x;      /* cross / marker centre; box: first corner */
    int32_t y1, x1;    /* box: opposite corner (inclusive) */
    uint8_t color[3];  /* in the image's channel order; the reference draws (0, 0, 255) */
    uint8_t reserved;
} hv_overlay;
HV_API hv_status hv_draw_overlays(hv_ctx *ctx, uint8_t *img_hw3, int32_t h, int32_t w, const hv_overlay *items, int32_t n);

/* ---- line statistics ------------------------------------------------------------------------------------ */
HV_API hv_status hv_stats_get(hv_ctx *ctx, hv_line_stats *out);
HV_API hv_status hv_stats_reset(hv_ctx *ctx);
/* Device pointer to the 32 x u64 stats vector, for the host's NCCL communicator (the only collective of this path). */
HV_API uint64_t *hv_stats_device_ptr(hv_ctx *ctx);

/* ---- measurement ---------------------------------------------------------------------------------------- */
/* Kernels launched by this context so far. */
HV_API uint64_t hv_launch_count(const hv_ctx *ctx);
/* Per-kernel device timing with CUDA events recorded on the launching stream.  kernel_mask: bit k enables kernel
 * HV_K_* = k (0 disables; HV_FLAG_PROFILE at creation enables all).  hv_profile_get waits for the recorded events,
 * returns the summed milliseconds and the number of timed launches per kernel since the previous call, and clears. */
HV_API hv_status hv_profile_enable(hv_ctx *ctx, uint32_t kernel_mask);
HV_API hv_status hv_profile_get(hv_ctx *ctx, float total_ms[HV_K_COUNT], uint32_t counts[HV_K_COUNT]);
HV_API const char *hv_kernel_name(int32_t k);
/* With HV_FLAG_PHASE_TIMING: globaltimer (ns) of every warp of frame 0's CTA at the phase boundaries of the fused
 * per-frame kernel, out_ns[stamp * 16 + warp]; entries not recorded are 0. */
HV_API hv_status hv_debug_phase_times(hv_ctx *ctx, uint64_t out_ns[256]);

#ifdef __cplusplus
}
#endif
#endif /* HEIMDALL_CUDA_H */
