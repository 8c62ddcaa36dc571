// hv_api.cu -- host side of the C ABI (include/heimdall_cuda.h): contexts, device scratch, pinned staging, streams,
// batch orchestration, result marshalling.  Mirrors the reference's PyO3 layer (rust/heimdall-core/src/lib.rs:42-178)
// at the granularity a Rust FFI crate would bind.  No CPU fallback anywhere: every result comes from the kernels.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "hv_common.cuh"

using namespace hv;

namespace {

int env_int(const char *name, int dflt, int lo, int hi) {
    const char *e = getenv(name);
    if (!e || !*e) return dflt;
    const int v = atoi(e);
    return v < lo ? lo : (v > hi ? hi : v);
}
bool env_flag(const char *name) { return getenv(name) != nullptr; }

hv::Tunables read_tunables() {
    hv::Tunables t;
    t.pipeline_depth = env_int("HV_PIPELINE_DEPTH", t.pipeline_depth, 2, 8);
    t.k1_ctas_per_sm = env_int("HV_K1_CTAS_PER_SM", t.k1_ctas_per_sm, 1, 5);
    t.k1_gauss_ctas = env_int("HV_K1_GAUSS_CTAS", t.k1_gauss_ctas, 1, 4);
    t.k1_gauss_small_ctas = env_int("HV_K1_GAUSS_SMALL_CTAS", t.k1_gauss_small_ctas, 1, 4);
    t.ccl_small_max_tiles = env_int("HV_CCL_SMALL_MAX_TILES", t.ccl_small_max_tiles, 0, 1 << 30);
    t.k1_ctas_coresident = env_int("HV_K1_CTAS_CORESIDENT", t.k1_ctas_coresident, 1, 5);
    t.defer_depth = env_int("HV_DEFER_DEPTH", t.defer_depth, 1, 3);
    t.no_side_ccl = env_flag("HV_NO_SIDE_CCL");
    t.k1_stages_coresident = env_int("HV_K1_STAGES", t.k1_stages_coresident, 2, 3);
    t.k1_lookahead = env_int("HV_K1_LOOKAHEAD", t.k1_lookahead, 1, 8);
    t.k1_tail_lookahead = env_int("HV_K1_TAIL_LOOKAHEAD", t.k1_tail_lookahead, 1, 8);
    t.k1_tail_rounds = env_int("HV_K1_TAIL_ROUNDS", t.k1_tail_rounds, 0, 1 << 20);
    t.k1_prefetch = env_int("HV_K1_PREFETCH", t.k1_prefetch, 0, 1 << 20);
    t.k1_claim_ahead = env_int("HV_K1_CLAIM_AHEAD", t.k1_claim_ahead, -1, 1);
    t.k1_wait_hint_ns = env_int("HV_K1_WAIT_HINT_NS", t.k1_wait_hint_ns, 0, 2000000000);
    t.morph_tiles_per_sm = env_int("HV_MORPH_TILES_PER_SM", t.morph_tiles_per_sm, 1, 8);
    t.phase_frame = env_int("HV_PHASE_FRAME", 0, 0, 1 << 20);
    t.k1_static = env_flag("HV_K1_STATIC");
    t.k1_no_tma = env_flag("HV_K1_NO_TMA");
    t.ccl_big = env_flag("HV_CCL_BIG");
    t.ccl_no_tiny = env_flag("HV_CCL_NO_TINY");
    t.no_k1_flag = env_flag("HV_NO_K1_FLAG");
    t.no_early_k1 = env_flag("HV_NO_EARLY_K1");
    t.no_pdl = env_flag("HV_NO_PDL");
    t.no_pdl_tail = env_flag("HV_NO_PDL_TAIL");
    t.no_compression = env_flag("HV_NO_COMPRESSION");
    t.no_fused_gauss = env_flag("HV_NO_FUSED_GAUSS");
    t.no_fused_morph = env_flag("HV_NO_FUSED_MORPH");
    t.no_morph_chain = env_flag("HV_NO_MORPH_CHAIN");
    t.no_k1_morph = env_flag("HV_NO_K1_MORPH");
#ifdef HV_EXPERIMENTS
    t.exp_k1_only = env_flag("HV_EXP_K1_ONLY");
    t.exp_ccl_noop = env_flag("HV_EXP_CCL_NOOP");
    t.exp_ccl_stop = env_int("HV_EXP_CCL_STOP", 0, 0, 9);
#endif
    return t;
}

// scratch sets rotated by the synchronous / device-resident entry points = batches in flight on the device (see
// enqueue_pipeline)
#define kSyncSlots (hv::tunables().pipeline_depth)

thread_local std::string g_create_error;

// driver entry points of the virtual-memory API, resolved through the runtime (no link-time dependency on libcuda)
template <typename F>
F drv_fn(const char *name) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<F>(p);
}

// Device memory with L2 compute-data compression (cuMemCreate + CU_MEM_ALLOCATION_COMP_GENERIC).  The mask and label
// planes are almost entirely zero; compressed lines cost less DRAM time on their way out of the L2 (measured on the
// headline batch: step 53.6 -> 49.2 us).  Returns false if the device or driver cannot do it (caller falls back to
// cudaMalloc).
struct VmmRec {
    unsigned long long handle = 0;
    size_t size = 0;
    bool compressed = false;
};
bool vmm_alloc_compressible(int device, size_t bytes, void **out, VmmRec *rec) {
    auto f_attr = drv_fn<decltype(&cuDeviceGetAttribute)>("cuDeviceGetAttribute");
    auto f_gran = drv_fn<decltype(&cuMemGetAllocationGranularity)>("cuMemGetAllocationGranularity");
    auto f_create = drv_fn<decltype(&cuMemCreate)>("cuMemCreate");
    auto f_props = drv_fn<decltype(&cuMemGetAllocationPropertiesFromHandle)>("cuMemGetAllocationPropertiesFromHandle");
    auto f_reserve = drv_fn<decltype(&cuMemAddressReserve)>("cuMemAddressReserve");
    auto f_map = drv_fn<decltype(&cuMemMap)>("cuMemMap");
    auto f_access = drv_fn<decltype(&cuMemSetAccess)>("cuMemSetAccess");
    auto f_release = drv_fn<decltype(&cuMemRelease)>("cuMemRelease");
    auto f_afree = drv_fn<decltype(&cuMemAddressFree)>("cuMemAddressFree");
    auto f_unmap = drv_fn<decltype(&cuMemUnmap)>("cuMemUnmap");
    int sup = 0;
    if (!(f_attr && f_gran && f_create && f_props && f_reserve && f_map && f_access && f_release && f_afree && f_unmap)) return false;
    if (tunables().no_compression) return false;
    if (f_attr(&sup, CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED, device) != CUDA_SUCCESS || !sup) return false;
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    prop.allocFlags.compressionType = CU_MEM_ALLOCATION_COMP_GENERIC;
    size_t gran = 0;
    if (f_gran(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || !gran) return false;
    const size_t size = (bytes + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle h;
    if (f_create(&h, size, &prop, 0) != CUDA_SUCCESS) return false;
    CUdeviceptr d = 0;
    const bool reserved = f_reserve(&d, size, 0, 0, 0) == CUDA_SUCCESS;
    const bool mapped = reserved && f_map(d, size, 0, h, 0) == CUDA_SUCCESS;
    CUmemAccessDesc a = {};
    a.location = prop.location;
    a.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (mapped && f_access(d, size, &a, 1) == CUDA_SUCCESS) {
        CUmemAllocationProp got = {};
        f_props(&got, h);
        rec->handle = (unsigned long long)h;
        rec->size = size;
        rec->compressed = got.allocFlags.compressionType == CU_MEM_ALLOCATION_COMP_GENERIC;
        *out = reinterpret_cast<void *>(d);
        return true;
    }
    if (mapped) f_unmap(d, size);
    if (reserved) f_afree(d, size);
    f_release(h);
    return false;
}
void vmm_free(void *p, const VmmRec &rec) {
    auto f_unmap = drv_fn<decltype(&cuMemUnmap)>("cuMemUnmap");
    auto f_release = drv_fn<decltype(&cuMemRelease)>("cuMemRelease");
    auto f_afree = drv_fn<decltype(&cuMemAddressFree)>("cuMemAddressFree");
    if (!(f_unmap && f_release && f_afree)) return;
    f_unmap((CUdeviceptr)p, rec.size);
    f_afree((CUdeviceptr)p, rec.size);
    f_release((CUmemGenericAllocationHandle)rec.handle);
}

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;  // elements
    bool want_compressible = false;  // output planes (mask, labels): try compressible memory first
    bool vmm = false;
    VmmRec rec;
    int device = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        release();
        if (want_compressible) {
            void *q = nullptr;
            if (vmm_alloc_compressible(device, n * sizeof(T), &q, &rec)) {
                p = static_cast<T *>(q);
                cap = n;
                vmm = true;
                return cudaSuccess;
            }
        }
        cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&p), n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() {
        if (p) {
            if (vmm) {
                cudaDeviceSynchronize();
                vmm_free(p, rec);
            } else {
                cudaFree(p);
            }
        }
        p = nullptr;
        cap = 0;
        vmm = false;
    }
};

template <typename T>
struct PinBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaHostAlloc(reinterpret_cast<void **>(&p), n * sizeof(T), cudaHostAllocPortable);
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    DevBuf<uint8_t> in, gray, blur, mask, rowflags, rowflags_tmp, tile_occ;
    DevBuf<uint32_t> tile_list;
    DevBuf<uint16_t> gauss_tmp;
    DevBuf<uint32_t> bits, bits_tmp, rootbits, rankbase, segbase, score_state, ncomp, fgcount, sched;
    uint32_t ccl_expected = 0;  // frames handed to the per-frame CCL kernel on this slot so far (sched[7] counts them done)
    uint32_t k1_expected = 0;   // K1 launches with a launch counter on this slot so far (sched[9] counts them done)
    uint32_t scan_expected = 0, tiles_expected = 0;  // likewise for the morphology kernels (sched[11], sched[12])
    cudaEvent_t copied = nullptr;   // device-resident batches: results / flags / defects have arrived in the pinned buffers
    bool pending = false;           // device-resident batch enqueued whose results have not been examined yet (retire_slot)
    bool defects_on_host = false;   // the read-back included the defect table (else it is fetched on demand)
    cudaStream_t batch_stream = nullptr;  // stream the batch's kernels were enqueued on
    PinBuf<uint8_t> h_stage;  // pinned staging for camera frames that arrive in pageable memory
    bool used_fused = false;  // the batch went through the fused per-frame CCL kernel
    bool used_small = false;  // ... its small build
    bool used_tiny = false;   // ... its tiny build
    bool tail_deferred = false;  // its per-frame kernel was held back when the batch was enqueued (HV_FLAG_DEFER_TAIL)
    bool tail_on_side = false;   // its global-memory CCL kernels went onto the slot's own stream (HV_FLAG_DEFER_TAIL)
    cudaStream_t tail_stream = nullptr;  // the stream its last kernel is on
    bool sparse_bits = false; // K1 left the bit-mask words of flat tiles unwritten (densify before any other reader)
    ScoreParams score{};
    DevBuf<int32_t> labels;
    DevBuf<hv_blob> blobs;
    DevBuf<hv_defect> defects;
    DevBuf<hv_frame_result> results;    // n results followed by n u32 frame flags (one read-back for both)
    PinBuf<hv_frame_result> h_results;  // same layout
    PinBuf<hv_defect> h_defects;
    // state of the batch currently held by the slot
    BatchView view{};
    bool has_batch = false;
    bool have_blur = false;
    int c = 1;
    const uint8_t *d_input = nullptr;  // device frames the batch was run on (for the gray debug copy, c == 1)
    int64_t ticket = -1;
    void release() {
        in.release(), gray.release(), blur.release(), mask.release(), gauss_tmp.release(), rowflags.release();
        rowflags_tmp.release(), tile_occ.release(), tile_list.release();
        bits.release(), bits_tmp.release(), rootbits.release(), rankbase.release(), ncomp.release();
        segbase.release(), score_state.release();
        fgcount.release(), labels.release(), blobs.release(), defects.release(), results.release();
        sched.release();
        h_results.release(), h_defects.release(), h_stage.release();
        if (done) cudaEventDestroy(done);
        if (copied) cudaEventDestroy(copied);
        copied = nullptr;
        if (stream) cudaStreamDestroy(stream);
        done = nullptr;
        stream = nullptr;
    }
};

}  // namespace

namespace hv {
const Tunables &tunables() {
    static const Tunables t = read_tunables();  // once per process (thread-safe static initialisation)
    return t;
}
}  // namespace hv

struct hv_ctx {
    int device = 0;
    int num_sms = 148;
    hv_config cfg{};
    std::string err;
    std::vector<Slot> slots;
    cudaStream_t user_stream = nullptr;
    bool use_user_stream = false;
    // results of device-resident batches travel on this stream: it waits (cuStreamWaitValue32) for the slot's completion
    // counter, so nothing is ever inserted between two kernels of the launching stream
    cudaStream_t copy_stream = nullptr;
    CUresult (*wait32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
    hv_line_stats *d_stats = nullptr;
    unsigned long long *d_phase_ns = nullptr;
    uint64_t launches = 0;
    int64_t next_ticket = 1;
    int next_slot = 0;
    // CCL path selection: sparse masks go through the fused per-frame kernel; if a batch needed the global-memory
    // fallback the next batches use the global path directly and the fused kernel is re-tried every 8th batch
    bool dense_hint = false;
    int sync_cur = 0;  // which of the synchronous slots holds the most recent batch
    // last batch enqueued through the device-resident path (for the programmatic-dependent-launch overlap decision)
    cudaStream_t last_stream = nullptr;
    bool last_valid = false, last_fused_tail = false;
    const void *last_mask = nullptr, *last_labels = nullptr;
    // The batches before that one whose per-frame CCL kernels may still be running when K1 of the batch being enqueued
    // starts (most recent first, at most kSyncSlots - 2 of them; the one that used the same slot is covered by the slot's
    // own counter): output planes, completion counter of their slot and the value that means "through".
    struct InFlight {
        const void *mask, *labels;
        unsigned int *done;
        uint32_t expected;
    };
    std::vector<InFlight> in_flight;
    unsigned int *last_done = nullptr;
    uint32_t last_expected = 0;
    // which build of the per-frame CCL kernel: the small one (co-resident with K1) until a frame did not fit in it
    bool ccl_small_ok = true;
    uint32_t ccl_small_retry = 0;
    // ... and the tiny one once a batch has reported (frame flags) that all its frames would fit it
    bool ccl_tiny_ok = false;
    uint32_t ccl_tiny_cooldown = 0;
    bool tail_used = false;  // kernels went onto slot streams since the last hv_flush
    // HV_FLAG_DEFER_TAIL: the per-frame kernel (and the read-back behind it) of the latest hv_enqueue_device batch, not yet
    // on the stream (flush_deferred)
    struct DeferredTail {
        int slot = -1;
        BatchView b;
        ScoreParams sp;
        int level = 0;
        bool pdl = false;
        cudaStream_t st = nullptr;
    };
    std::deque<DeferredTail> deferred;  // oldest first
    // profiling: event pairs recorded around kernels whose bit is set in prof_mask
    struct ProfRec {
        int k;
        cudaEvent_t a, b;
    };
    uint32_t prof_mask = 0;
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    // scratch for the single-frame utilities
    DevBuf<uint8_t> u_a, u_b, u_c;
    DevBuf<hv_center> u_centers;
    DevBuf<hv_contour> u_contours;
    DevBuf<uint32_t> u_count, u_owner;
    // buffers handed out by hv_device_alloc through the virtual-memory API (compressible memory)
    std::map<void *, VmmRec> vmm;
};

namespace {

hv_status fail(hv_ctx *ctx, hv_status s, const std::string &msg) {
    if (ctx) ctx->err = msg;
    return s;
}

hv_status fail_cuda(hv_ctx *ctx, cudaError_t e, const char *what) {
    std::string m = std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e);
    cudaGetLastError();  // clear the sticky-free error state
    return fail(ctx, HV_ERR_CUDA, m);
}

#define HV_TRY_CUDA(ctx, expr)                                         \
    do {                                                               \
        cudaError_t _e = (expr);                                       \
        if (_e != cudaSuccess) return fail_cuda((ctx), _e, #expr);     \
    } while (0)

// `c = threshold as i32` (detection.rs:186): truncating, saturating, NaN -> 0.  The comparison `px < mean - c` is i32
// arithmetic that WRAPS in the reference's release build (rust/Cargo.toml [profile.release]: no overflow-checks):
//   c > 255 - 2^31: no wrap for any u8 mean; beyond [-256, 256] the comparison is constant, so c is clamped there;
//   c <= 255 - 2^31 (threshold <= -2147483393.0): pixels whose window mean is >= T = c + 2^31 wrap to "never
//   foreground", the others are "always foreground": mask = mean < T.  *wrap_t1 = T + 1 (0 = no wrap).
int threshold_plan(double thr, int *wrap_t1) {
    long long c;
    if (!(thr == thr))
        c = 0;
    else if (thr >= 2147483647.0)
        c = 2147483647LL;
    else if (thr <= -2147483648.0)
        c = -2147483648LL;
    else
        c = (long long)thr;
    *wrap_t1 = 0;
    if (c <= 255LL - 2147483648LL) {
        *wrap_t1 = (int)(c + 2147483648LL) + 1;
        return -256;  // negative: no flat-tile skip, no packed fast path
    }
    if (c > 256) c = 256;
    if (c < -256) c = -256;
    return (int)c;
}

int blob_cap_for(const hv_ctx *ctx, int h, int w) {
    const long long max_possible = (long long)h * w / 2 + 1;
    long long cap = ctx->cfg.max_blobs_per_frame > 0 ? ctx->cfg.max_blobs_per_frame : 131072;
    return (int)std::min(cap, max_possible);
}

// results buffer: n hv_frame_result entries followed by n u32 frame flags, in units of hv_frame_result
size_t results_entries(int n) { return (size_t)n + ((size_t)n * sizeof(uint32_t) + sizeof(hv_frame_result) - 1) / sizeof(hv_frame_result); }
uint32_t *flags_after(hv_frame_result *r, int n) { return reinterpret_cast<uint32_t *>(r + n); }

int defect_cap_for(const hv_ctx *ctx) { return ctx->cfg.max_defects_per_frame > 0 ? ctx->cfg.max_defects_per_frame : 256; }

cudaEvent_t prof_event(hv_ctx *ctx) {
    if (!ctx->prof_pool.empty()) {
        cudaEvent_t e = ctx->prof_pool.back();
        ctx->prof_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

struct ProfScope {
    hv_ctx *ctx;
    int k;
    cudaStream_t s;
    cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(hv_ctx *c, int kk, cudaStream_t ss) : ctx(c), k(kk), s(ss) {
        if (ctx->prof_mask & (1u << k)) {
            a = prof_event(ctx);
            b = prof_event(ctx);
            if (a) cudaEventRecord(a, s);
        }
    }
    ~ProfScope() {
        if (a && b) {
            cudaEventRecord(b, s);
            ctx->prof_recs.push_back({k, a, b});
        }
    }
};

// OpenCV's 8-bit fixed-point Gaussian kernel (getGaussianKernelBitExact + getGaussianKernelFixedPoint_ED, 8 fractional
// bits): normalised double kernel, error-diffused rounding from the ends toward the centre, centre takes the rest.
bool gaussian_kernel_q8(int n, double sigma, uint16_t *k16) {
    if (n <= 0 || n > 31 || (n & 1) == 0) return false;
    double kd[31];
    const int n2 = (n - 1) / 2;
    static const double f3[] = {0.25, 0.5, 0.25};
    static const double f5[] = {0.0625, 0.25, 0.375, 0.25, 0.0625};
    static const double f7[] = {0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125};
    static const double f9[] = {4.0 / 256, 13.0 / 256, 30.0 / 256, 51.0 / 256, 60.0 / 256,
                                51.0 / 256, 30.0 / 256, 13.0 / 256, 4.0 / 256};
    const double *fixed = nullptr;
    if (sigma <= 0) fixed = n == 3 ? f3 : n == 5 ? f5 : n == 7 ? f7 : n == 9 ? f9 : nullptr;
    if (sigma <= 0 && n == 1) {
        kd[0] = 1.0;
    } else if (fixed) {
        for (int i = 0; i < n; i++) kd[i] = fixed[i];
    } else {
        const double sigmaX = sigma > 0 ? sigma : std::fma((double)n, 0.15, 0.35);
        const double scale2X = -0.125 / (sigmaX * sigmaX);
        double values[16];
        double sum = 0.0;
        for (int i = 0, x = 1 - n; i < n2; i++, x += 2) {
            values[i] = std::exp((double)(x * x) * scale2X);
            sum += values[i];
        }
        sum = sum * 2.0 + 1.0;
        const double mul1 = 1.0 / sum;
        for (int i = 0; i < n2; i++) kd[i] = kd[n - 1 - i] = values[i] * mul1;
        kd[n2] = mul1;
    }
    double err = 0.0;
    long long sum = 0;
    for (int i = 0; i < n2; i++) {
        const double adj = kd[i] * 256.0 + err;
        const long long v0 = (long long)std::nearbyint(adj);
        err = adj - (double)v0;
        k16[i] = k16[n - 1 - i] = (uint16_t)v0;
        sum += v0;
    }
    k16[n2] = (uint16_t)(256 - 2 * sum);
    return true;
}

cudaStream_t sync_stream(hv_ctx *ctx) { return ctx->use_user_stream ? ctx->user_stream : ctx->slots[0].stream; }
Slot &cur_sync_slot(hv_ctx *ctx) { return ctx->slots[ctx->sync_cur]; }

hv_status validate_shape(hv_ctx *ctx, int n, int h, int w, int c) {
    if (n <= 0 || h <= 0 || w <= 0) return fail(ctx, HV_ERR_INVALID_ARGUMENT, "batch, height and width must be positive");
    if (c != 1 && c != 3) return fail(ctx, HV_ERR_INVALID_DIMENSIONS, "Invalid image dimensions: expected 3D array");
    if ((long long)h * w >= 2147483647LL) return fail(ctx, HV_ERR_INVALID_ARGUMENT, "frame too large (h*w must fit in i32)");
    return HV_OK;
}

hv_status reserve_slot(hv_ctx *ctx, Slot &s, int n, int h, int w, bool need_in, size_t in_bytes, bool need_gray,
                       bool need_blur, bool need_gauss, bool need_mask, bool need_labels) {
    const size_t px = (size_t)n * h * w;
    const int ww = (w + 31) / 32;
    const size_t words = (size_t)n * h * ww;
    if (need_in) HV_TRY_CUDA(ctx, s.in.reserve(in_bytes));
    if (need_gray) HV_TRY_CUDA(ctx, s.gray.reserve(px));
    if (need_blur) HV_TRY_CUDA(ctx, s.blur.reserve(px));
    if (need_gauss) HV_TRY_CUDA(ctx, s.gauss_tmp.reserve(px));
    if (need_mask) HV_TRY_CUDA(ctx, s.mask.reserve(px));
    if (need_labels) HV_TRY_CUDA(ctx, s.labels.reserve(px));
    HV_TRY_CUDA(ctx, s.rowflags.reserve((size_t)n * ((h + 31) / 32) * ((w + 127) / 128) * 32));
    HV_TRY_CUDA(ctx, s.bits.reserve(words));
    HV_TRY_CUDA(ctx, s.bits_tmp.reserve(words));
    HV_TRY_CUDA(ctx, s.rootbits.reserve(words));
    HV_TRY_CUDA(ctx, s.rankbase.reserve(words));
    HV_TRY_CUDA(ctx, s.segbase.reserve(2 * (size_t)n * (((size_t)h * ww + 255) / 256)));
    {
        const size_t need = (size_t)n * (((size_t)blob_cap_for(ctx, h, w) + 255) / 256 + 1);
        if (need > s.score_state.cap) {  // all zero between launches: the scoring kernel cleans up after itself
            HV_TRY_CUDA(ctx, s.score_state.reserve(need));
            HV_TRY_CUDA(ctx, cudaMemset(s.score_state.p, 0, need * sizeof(uint32_t)));
        }
    }
    HV_TRY_CUDA(ctx, s.ncomp.reserve(n));
    HV_TRY_CUDA(ctx, s.fgcount.reserve(n));
    if (!s.sched.p) {  // K1's tile scheduler: {next tile, CTAs done}; the kernel rearms it itself
        // [0..1] K1's tile counter, [4..6] the fused morphology kernels, [7] frames the per-frame CCL kernel is through
        // with, [8] K1 CTAs done in the running launch, [9] K1 launches done, [10..12] morphology: scan CTAs done, scan
        // launches done, tiles launches done
        HV_TRY_CUDA(ctx, s.sched.reserve(16));
        HV_TRY_CUDA(ctx, cudaMemset(s.sched.p, 0, 16 * sizeof(uint32_t)));
    }
    HV_TRY_CUDA(ctx, s.blobs.reserve((size_t)n * blob_cap_for(ctx, h, w)));
    HV_TRY_CUDA(ctx, s.defects.reserve((size_t)n * defect_cap_for(ctx)));
    HV_TRY_CUDA(ctx, s.results.reserve(results_entries(n)));
    HV_TRY_CUDA(ctx, s.h_results.reserve(results_entries(n)));
    HV_TRY_CUDA(ctx, s.h_defects.reserve((size_t)n * defect_cap_for(ctx)));
    return HV_OK;
}

// The global-memory CCL path (K2..K6): robust for any density; b.frame_select restricts it to flagged frames.
hv_status enqueue_global_ccl(hv_ctx *ctx, const BatchView &b, const ScoreParams &sp, cudaStream_t st) {
    {
        ProfScope ps(ctx, HV_K_CCL_MERGE, st);
        HV_TRY_CUDA(ctx, launch_ccl_merge(b, st));
    }
    {
        ProfScope ps(ctx, HV_K_CCL_FLATTEN, st);
        HV_TRY_CUDA(ctx, launch_ccl_flatten(b, st));
    }
    {
        ProfScope ps(ctx, HV_K_CCL_SCAN, st);
        HV_TRY_CUDA(ctx, launch_ccl_scan(b, st));
    }
    {
        ProfScope ps(ctx, HV_K_CCL_LABEL, st);
        HV_TRY_CUDA(ctx, launch_ccl_label(b, st));
    }
    {
        ProfScope ps(ctx, HV_K_SCORE, st);
        HV_TRY_CUDA(ctx, launch_score(b, sp, st));
    }
    ctx->launches += 5;
    return HV_OK;
}

hv_status enqueue_async_readback(hv_ctx *ctx, Slot &s, cudaStream_t st);

// HV_FLAG_DEFER_TAIL.  The stream sequence of a streaming loop is K1(i), CCL(i), K1(i+1), CCL(i+1), ...: K1(i+1) is released
// by the per-frame kernel of batch i the moment that kernel is resident (programmatic launch), and that is what keeps the
// device busy while the per-frame kernel works through its latency.  Whatever the caller puts onto the stream between two
// calls -- an event, a copy -- lands between CCL(i) and K1(i+1) and turns that hand-over into plain stream order: K1(i+1)
// then waits for the whole per-frame kernel (about 40 us of mostly idle device for the headline batch).  With the tails of
// the last defer_depth batches held back, call i enqueues CCL(i - 2), K1(i): the caller's work lands behind a K1, in front
// of a per-frame kernel whose own K1 completed a batch earlier, and the next K1 is released as soon as that kernel is
// resident.  A per-frame kernel enqueued a batch late also finds its input complete: it is resident for its own 40-50 us,
// not for K1's duration on top (measured, no events between the calls: 40.7 -> 40.2 us per batch; with an event every 20
// batches: 44.2 -> 43.4 us with one tail held back -> 41.9 with two).  The price: a batch's label plane and results are
// complete in stream order only defer_depth calls later, or after hv_flush / hv_fetch_ticket -- which is why it is opt-in.
// The read-back is enqueued together with the kernel, never ahead of it, so a device-wide synchronize by the caller
// cannot wait for a kernel that has not been launched.
hv_status flush_deferred(hv_ctx *ctx, size_t keep = 0) {
    while (ctx->deferred.size() > keep) {
        hv_ctx::DeferredTail d = ctx->deferred.front();
        ctx->deferred.pop_front();
        Slot &s = ctx->slots[d.slot];
        HV_TRY_CUDA(ctx, launch_ccl_frame(d.b, d.sp, d.pdl, d.level, d.st));
        ctx->launches += 1;
        hv_status rs = enqueue_async_readback(ctx, s, d.st);
        if (rs != HV_OK) return rs;
    }
    return HV_OK;
}

// Enqueue the whole detect pipeline for frames already on the device.  No host synchronisation.
hv_status enqueue_pipeline(hv_ctx *ctx, Slot &s, cudaStream_t st, const uint8_t *d_frames, int n, int h, int w, int c,
                           size_t row_stride, size_t frame_stride, const hv_params &pr, uint8_t *d_mask,
                           int32_t *d_labels, bool want_blur, bool may_defer = false) {
    const Tunables &tun = tunables();
    {
        // the tails held back go onto the stream ahead of this batch's K1; a streaming call keeps the latest
        // (defer_depth - 1) of them back
        const bool streaming = may_defer && (ctx->cfg.flags & HV_FLAG_DEFER_TAIL) && !ctx->deferred.empty() &&
                               ctx->deferred.back().st == st;
        hv_status rf = flush_deferred(ctx, streaming ? (size_t)(tun.defer_depth - 1) : 0);
        if (rf != HV_OK) return rf;
    }
    if (row_stride == 0) row_stride = (size_t)w * c;
    if (frame_stride == 0) frame_stride = row_stride * h;
    if (row_stride < (size_t)w * c || frame_stride < row_stride * (size_t)h)
        return fail(ctx, HV_ERR_INVALID_ARGUMENT, "strides smaller than the frame");
    if (pr.blur_mode != HV_BLUR_BOX && pr.blur_mode != HV_BLUR_GAUSSIAN && pr.blur_mode != HV_BLUR_NONE)
        return fail(ctx, HV_ERR_INVALID_ARGUMENT, "unknown blur_mode");
    if (pr.morph_open_k < 0 || pr.morph_open_k > 31 || pr.morph_close_k < 0 || pr.morph_close_k > 31)
        return fail(ctx, HV_ERR_INVALID_ARGUMENT, "morphology kernel size must be in [0, 31]");
    const bool fused_box = pr.blur_mode == HV_BLUR_BOX && pr.blur_ksize / 2 == 2;
    const bool box_other = pr.blur_mode == HV_BLUR_BOX && !fused_box && pr.blur_ksize / 2 > 0;
    const bool gauss = pr.blur_mode == HV_BLUR_GAUSSIAN;
    uint16_t gk[32];
    if (gauss && !gaussian_kernel_q8(pr.blur_ksize, pr.gauss_sigma, gk))
        return fail(ctx, HV_ERR_INVALID_ARGUMENT, "Gaussian kernel size must be odd and in [1, 31]");
    const bool separate_blur = box_other || gauss;
    const bool morph_req = pr.morph_open_k > 0 || pr.morph_close_k > 0;

    hv_status rs = reserve_slot(ctx, s, n, h, w, false, 0, c == 3, separate_blur || want_blur, gauss, d_mask == nullptr,
                                d_labels == nullptr);
    if (rs != HV_OK) return rs;

    BatchView b{};
    b.n = n, b.h = h, b.w = w, b.ww = (w + 31) / 32;
    if (c == 3) {
        ProfScope ps(ctx, HV_K_GRAY, st);
        HV_TRY_CUDA(ctx, launch_gray3(d_frames, n, h, w, row_stride, frame_stride, s.gray.p, st));
        ctx->launches++;
        b.gray = s.gray.p;
        b.gray_row_stride = w;
        b.gray_frame_stride = (size_t)h * w;
    } else {
        b.gray = d_frames;
        b.gray_row_stride = row_stride;
        b.gray_frame_stride = frame_stride;
    }
    b.blur = s.blur.p;
    b.mask = d_mask ? d_mask : s.mask.p;
    b.bits = s.bits.p;
    b.bits_tmp = s.bits_tmp.p;
    b.rowflags = s.rowflags.p;
    b.tiles_x = (w + 127) / 128;
    b.rf_stride = (size_t)((h + 31) / 32) * b.tiles_x * 32;
    b.labels = d_labels ? d_labels : s.labels.p;
    b.rootbits = s.rootbits.p;
    b.rankbase = s.rankbase.p;
    b.segbase = s.segbase.p;
    b.nseg = (int)(((size_t)h * b.ww + 255) / 256);
    b.score_state = s.score_state.p;
    b.ncomp = s.ncomp.p;
    b.fgcount = s.fgcount.p;
    b.blobs = s.blobs.p;
    b.blob_cap = blob_cap_for(ctx, h, w);
    b.score_chunks = (b.blob_cap + 255) / 256;
    b.defects = s.defects.p;
    b.defect_cap = defect_cap_for(ctx);
    b.results = s.results.p;
    b.stats = ctx->d_stats;
    b.frame_flags = flags_after(s.results.p, n);
    b.frame_select = nullptr;
    b.ccl_done = (tun.no_early_k1 || tun.exp_k1_only) ? nullptr : s.sched.p + 7;
    b.ccl_wait_value = s.ccl_expected;
    // The caller's output planes are not tied to our slots: if this batch writes planes that one of the batches still in
    // flight wrote (a caller rotating fewer sets than we have slots), K1 also waits for that batch's per-frame kernel.
    b.ccl_wait_n = 0;
    bool conflict_overflow = false;
    if (b.ccl_done && ctx->last_valid && ctx->last_stream == st)
        for (const auto &q : ctx->in_flight)
            if (q.done && (q.mask == (const void *)b.mask || q.labels == (const void *)b.labels)) {
                if (b.ccl_wait_n == 4) {  // more conflicts than K1 can wait for: this batch is launched in plain stream order
                    conflict_overflow = true;
                    break;
                }
                b.ccl_wait_flag[b.ccl_wait_n] = q.done;
                b.ccl_wait_val[b.ccl_wait_n] = q.expected;
                b.ccl_wait_n++;
            }
    b.phase_ns = (ctx->cfg.flags & HV_FLAG_PHASE_TIMING) ? ctx->d_phase_ns : nullptr;
    if (tun.exp_ccl_noop) b.phase_frame = -12345 - tun.exp_ccl_stop;
    if (b.phase_ns) {
        cudaMemsetAsync(ctx->d_phase_ns + 192, 0, 64 * sizeof(unsigned long long), st);
        cudaMemsetAsync(ctx->d_phase_ns + 248, 0xff, sizeof(unsigned long long), st);
        cudaMemsetAsync(ctx->d_phase_ns + 252, 0xff, sizeof(unsigned long long), st);
        b.phase_frame = tun.phase_frame;
    }

    // Morphology with 3x3 / 5x5 rectangles (total reach <= 4 px) is folded into K1 itself (k_preprocess.cu, MTile): the
    // kernels behind K1 are then exactly those of the plain pipeline.  Anything else goes through the tiles kernels.
    PreprocessParams pp{};
    pp.c_thresh = threshold_plan(pr.threshold, &pp.wrap_t1);
    pp.inverse = 1;
    pp.force_generic = (ctx->cfg.flags & HV_FLAG_FORCE_GENERIC) ? 1 : 0;
    bool k1_morph = false;
    if (morph_req && fused_box && !want_blur) {
        PreprocessParams probe = pp;
        probe.blur_radius = 2, probe.write_mask = 1, probe.init_labels = 1;
        k1_morph = preprocess_tma_morph_supported(b, probe, pr.morph_open_k, pr.morph_close_k);
    }
    const bool morph = morph_req && !k1_morph;
    // CCL path of this batch (decided before K1 runs: the fused kernel lets K1 skip the all-zero bit-mask words)
    bool fused = !(ctx->cfg.flags & HV_FLAG_GLOBAL_CCL) && ccl_frame_supported(b);
    // After a batch that needed the global-memory kernels the context stays on them until the results of a later batch
    // say that its frames would fit the per-frame kernel again (retire_slot: few components, little foreground).  A probe
    // that fails costs the whole batch twice, so there is no blind re-try.
    if (fused && ctx->dense_hint) fused = false;
    // Small build of the per-frame kernel (co-resident with K1 CTAs): box-blur path, with or without the fused morphology.  After a frame did
    // not fit, the big build takes over until it reports (frame_flags bit 1) that a whole batch would have fitted again.
    const bool morph_fused_plan = morph && morph_expand_supported(pr.morph_open_k, pr.morph_close_k) && !tun.no_fused_morph;
    // counter chain through the morphology kernels too (K1 -> scan -> tiles -> CCL), HV_NO_MORPH_CHAIN: griddepcontrol.wait
    // (only for small batches: the chain needs the tiles kernel resident as a whole, two CTAs per SM, which pays when launch
    // gaps and serialisation dominate -- 8000 tiles: 78 -> 66 us -- and costs when the kernels are long -- 256 x 5 MP: 2.0 ->
    // 2.6 ms)
    // (HV_FLAG_DEFER_TAIL: long kernels behind K1 go onto the slot's own stream, see below; for the morphology tiles kernels
    //  that also beats the counter chain on small batches -- headline batch, open+close k = 7 / 15: 69.3 / 83.1 -> 65.2 / 78.3 us)
    const bool side_ok = may_defer && (ctx->cfg.flags & HV_FLAG_DEFER_TAIL) && ctx->prof_mask == 0 && s.stream != st &&
                         !tun.no_side_ccl;
    const bool morph_chain = morph_fused_plan && !tun.no_morph_chain && !side_ok &&
                             (size_t)n * ((h + 31) / 32) * ((w + 127) / 128) <= 16384;
    bool ccl_small = fused && (!morph || morph_chain) && !gauss && !box_other && c == 1 && b.ccl_done && !tun.ccl_big;
    if (!ctx->ccl_small_ok) ccl_small = false;  // until the big build reports frames that fit the small one again
    // Long kernels (many tiles per launch) do not need the gapless chain -- launch gaps and the drain of the per-frame kernel
    // are a small part of the step -- and K1 is faster with all its CTAs: the big build behind a full-occupancy K1.
    // Measured with open + close 3x3 folded into K1: 64 x 5 MP 466 -> 428 us per step; 25 x 1.3 MP 50.6 (small) vs 53.9 us;
    // plain pipeline on 1.3 MP frames: 250 frames 519 -> 415 us, 100 frames 172 vs 174, 64 frames 107 (small) vs 115.
    if ((size_t)n * ((h + 31) / 32) * ((w + 127) / 128) > (size_t)tun.ccl_small_max_tiles) ccl_small = false;
    // resident K1 CTAs per SM (0 = the kernel's default, 5).  Next to the small CCL build: 3.  Four would fit beside one CTA
    // of it, but those CTAs often land two to an SM, and two of them leave room for two K1 CTAs whatever K1 asked for --
    // with three the loss is one CTA instead of two (measured in one call: 41.9 us per step with 4, 41.1 with 3, although
    // K1 alone is slower with 3: 53.5 vs 50.4 us).  The morphology tiles kernel needs the room as well.
    pp.ctas_per_sm = ccl_small ? (k1_morph ? std::max(tun.k1_ctas_coresident, 4) : tun.k1_ctas_coresident) : 0;
    pp.stages = ccl_small ? tun.k1_stages_coresident : 2;
    pp.sparse_aux = (fused && !morph) ? 1 : 0;
    if (k1_morph) pp.morph_open_k = pr.morph_open_k, pp.morph_close_k = pr.morph_close_k;
    // Morphology in the fused kernel (k <= 15): K1 writes mask and labels as usual and the morphology kernel rewrites only
    // the tiles in reach of foreground.  Multi-pass fallback: K1 writes the bit plane only, everything is expanded later.
    if (morph_fused_plan) {
        const size_t ntiles = (size_t)n * ((h + 31) / 32) * b.tiles_x;
        HV_TRY_CUDA(ctx, s.rowflags_tmp.reserve((size_t)n * b.rf_stride));
        HV_TRY_CUDA(ctx, s.tile_occ.reserve(ntiles * 4));
        HV_TRY_CUDA(ctx, s.tile_list.reserve(ntiles));
        b.tile_occ = s.tile_occ.p;
    }
    pp.write_mask = (morph && !morph_fused_plan) ? 0 : 1;
    pp.init_labels = (morph && !morph_fused_plan) ? 0 : 1;
    BatchView kb = b;  // view handed to K1 (its "gray" may be a separately blurred image)
    // K1 (TMA kernel) publishes a launch counter for the per-frame CCL kernel that follows it directly
    unsigned int *k1_flag = (b.ccl_done && fused && (!morph || (morph_chain && ccl_small)) && !tun.no_k1_flag) ? s.sched.p + 9 : nullptr;
    kb.k1_done = k1_flag;
    // Gaussian blur fused into the TMA kernel (k <= 15, 16-px aligned frames); otherwise two separable passes first
    bool gauss_fused = false, k1_tma = false;
    if (gauss && pr.blur_ksize <= 15 && !(ctx->cfg.flags & HV_FLAG_FORCE_GENERIC) && !tun.no_fused_gauss) {
        PreprocessParams gp = pp;
        gp.gauss_ksize = pr.blur_ksize;
        for (int t = 0; t < 16; t++) gp.gk[t] = t < pr.blur_ksize ? gk[t] : 0;
        gp.blur_radius = 0;
        gp.write_blur = want_blur ? 1 : 0;
        ProfScope ps(ctx, HV_K_PREPROCESS, st);
        const bool pdl = ctx->last_valid && ctx->last_fused_tail && ctx->last_stream == st && ctx->prof_mask == 0 && c == 1 && !conflict_overflow &&
                         ctx->last_mask != (const void *)b.mask && ctx->last_labels != (const void *)b.labels &&
                         !tun.no_pdl;
        gp.static_sched = tun.k1_static ? 1 : 0;  // tiles handed out by an atomic counter (see k_preprocess.cu)
        HV_TRY_CUDA(ctx, launch_preprocess_tma(kb, gp, b.bits, s.sched.p, ctx->num_sms, pdl, st, &gauss_fused));
        if (gauss_fused) ctx->launches++;
        k1_tma = gauss_fused;
    }
    if (gauss_fused) {
        // nothing else to do before morphology / CCL
    } else if (separate_blur) {
        if (gauss) {
            // tightly packed gray required by the simple Gaussian kernels
            if (c == 1 && (b.gray_row_stride != (size_t)w || b.gray_frame_stride != (size_t)h * w))
                return fail(ctx, HV_ERR_UNSUPPORTED, "Gaussian blur needs tightly packed frames");
            ProfScope ps(ctx, HV_K_PREPROCESS, st);
            HV_TRY_CUDA(ctx, launch_gaussian_blur(b.gray, n, h, w, gk, pr.blur_ksize, s.blur.p, s.gauss_tmp.p, st));
            ctx->launches += 2;
        } else {
            if (c == 1 && (b.gray_row_stride != (size_t)w || b.gray_frame_stride != (size_t)h * w))
                return fail(ctx, HV_ERR_UNSUPPORTED, "non-default box radius needs tightly packed frames");
            for (int f = 0; f < n; f++) {
                HV_TRY_CUDA(ctx, launch_box_blur_generic(b.gray + (size_t)f * h * w, h, w, 1, pr.blur_ksize / 2,
                                                         s.blur.p + (size_t)f * h * w, st));
                ctx->launches++;
            }
        }
        kb.gray = s.blur.p;
        kb.gray_row_stride = w;
        kb.gray_frame_stride = (size_t)h * w;
        pp.blur_radius = 0;
        pp.write_blur = 0;
    } else {
        pp.blur_radius = fused_box ? 2 : 0;
        pp.write_blur = want_blur ? 1 : 0;
    }
    if (!gauss_fused) {
        ProfScope ps(ctx, HV_K_PREPROCESS, st);
        bool used_tma = false;
        // Programmatic dependent launch: if the kernel right before this one on the stream is the previous batch's fused
        // per-frame CCL kernel (25 CTAs, latency-bound) and the two batches share no buffers, K1 may start while it is
        // still running.  Anything else in between (copies, events, other kernels) means plain stream order.
        const bool pdl = ctx->last_valid && ctx->last_fused_tail && ctx->last_stream == st && ctx->prof_mask == 0 && c == 1 && !conflict_overflow &&
                         !separate_blur && ctx->last_mask != (const void *)b.mask && ctx->last_labels != (const void *)b.labels &&
                         !tun.no_pdl;
        pp.static_sched = tun.k1_static ? 1 : 0;  // tiles handed out by an atomic counter (see k_preprocess.cu)
        if (!(ctx->cfg.flags & HV_FLAG_FORCE_GENERIC))
            HV_TRY_CUDA(ctx, launch_preprocess_tma(kb, pp, b.bits, s.sched.p, ctx->num_sms, pdl, st, &used_tma));
        if (!used_tma) HV_TRY_CUDA(ctx, launch_preprocess(kb, pp, b.bits, st));
        k1_tma = used_tma;
        ctx->launches++;
    }
    if (k1_flag && k1_tma) {
        s.k1_expected++;
        b.k1_done = k1_flag;
        b.k1_wait_value = s.k1_expected;
    }
    bool morph_fused = false;
    // HV_FLAG_DEFER_TAIL, batches whose kernels behind K1 are long (morphology through the tiles kernels without the counter
    // chain, the global-memory CCL kernels): everything behind K1 goes onto the slot's own stream behind an event on K1, so
    // that K1 of the next batch runs beside those latency-bound kernels instead of behind them.  Same contract as the
    // deferred per-frame kernel: labels and results are complete after hv_flush / hv_fetch_ticket.
    cudaStream_t ts = st;
    auto to_side_stream = [&]() -> cudaError_t {
        if (ts != st) return cudaSuccess;
        cudaError_t e = cudaEventRecord(s.done, st);
        if (e != cudaSuccess) return e;
        e = cudaStreamWaitEvent(s.stream, s.done, 0);
        if (e != cudaSuccess) return e;
        ts = s.stream;
        s.tail_on_side = true;
        ctx->tail_used = true;
        return cudaSuccess;
    };
    if (morph) {
        ProfScope ps(ctx, HV_K_MORPH, st);
        if (side_ok && !morph_chain) HV_TRY_CUDA(ctx, to_side_stream());
        if (morph_fused_plan) {
            // open + close + expansion in one kernel, launched ahead of K1's completion; its result goes to the other
            // bit plane, which is the one the CCL reads from here on
            const bool pdl_mid = k1_tma && ctx->prof_mask == 0 && !tun.no_pdl && ts == st;
            // counter chain: the scan waits for K1's launch counter (b.k1_done), the tiles kernel for the scan's, and the
            // per-frame CCL kernel for the tiles kernel's (handed to it in the k1_done fields)
            unsigned int *chain = (b.k1_done && morph_chain) ? s.sched.p + 10 : nullptr;
            if (chain) s.scan_expected++;
            HV_TRY_CUDA(ctx, launch_morph_expand(b, pr.morph_open_k, pr.morph_close_k, b.bits_tmp, s.rowflags_tmp.p,
                                                 s.tile_list.p, s.sched.p + 4, chain, s.scan_expected, ctx->num_sms, pdl_mid, ts));
            std::swap(b.bits, b.bits_tmp);
            b.rowflags = s.rowflags_tmp.p;
            if (chain) {
                s.tiles_expected++;
                b.k1_done = chain + 2;
                b.k1_wait_value = s.tiles_expected;
            } else {
                b.k1_done = nullptr;  // the per-frame kernel waits with griddepcontrol.wait
            }
            ctx->launches += 2;
            morph_fused = true;
        } else {
            int nl = 0;
            HV_TRY_CUDA(ctx, launch_morph(b, pr.morph_open_k, pr.morph_close_k, &nl, ts));
            HV_TRY_CUDA(ctx, launch_expand_bits(b, ts));
            ctx->launches += nl + 1;
        }
    }
    ScoreParams sp{pr.min_size, pr.max_size, pr.min_confidence};
    bool deferred_now = false;
    cudaStream_t tail_stream = st;  // the stream the batch's last kernel is on
    if (fused && tun.exp_k1_only) {
        // experiment: K1 chain alone (results are NOT computed)
    } else if (fused) {
        ProfScope ps(ctx, HV_K_CCL_FRAME, st);
        // launched ahead of K1's completion when K1 (TMA kernel, which releases its dependents at once) is the kernel
        // right before it on the stream; the kernel waits for K1 itself (griddepcontrol.wait)
        const bool pdl_tail = k1_tma && (!morph || morph_fused) && ctx->prof_mask == 0 && !tun.no_pdl &&
                              !tun.no_pdl_tail;
        const bool ccl_tiny = ccl_small && ctx->ccl_tiny_ok && !tun.ccl_no_tiny;
        const int level = ccl_tiny ? 2 : (ccl_small ? 1 : 0);
        // (deferred only in the plain chain K1 -> per-frame kernel, where the kernel waits for K1's launch counter)
        if (may_defer && (ctx->cfg.flags & HV_FLAG_DEFER_TAIL) && pdl_tail && b.k1_done && !morph && ctx->wait32 && !b.phase_ns) {
            hv_ctx::DeferredTail d;
            d.slot = (int)(&s - ctx->slots.data());
            d.b = b, d.sp = sp, d.level = level, d.pdl = pdl_tail, d.st = st;
            ctx->deferred.push_back(d);
            deferred_now = true;
        } else {
            HV_TRY_CUDA(ctx, launch_ccl_frame(b, sp, pdl_tail, level, ts));
            ctx->launches += 1;
            tail_stream = ts;
        }
        s.ccl_expected += (uint32_t)n;
        s.used_small = ccl_small && !ccl_tiny;
        s.used_tiny = ccl_tiny;
    } else {
        if (side_ok) HV_TRY_CUDA(ctx, to_side_stream());  // (dense frames: the five global-memory CCL kernels)
        cudaStream_t cs = ts;
        if (morph_fused) {  // the global path scans every word: give the tiles the morphology skipped their zero words
            HV_TRY_CUDA(ctx, launch_densify_bits(b, cs));
            ctx->launches++;
        }
        hv_status rg = enqueue_global_ccl(ctx, b, sp, cs);
        if (rg != HV_OK) return rg;
        tail_stream = cs;
    }
    s.tail_deferred = deferred_now;
    s.tail_stream = tail_stream;
    s.used_fused = fused;
    s.sparse_bits = pp.sparse_aux != 0 || (morph_fused && fused);  // (fused morphology: tiles out of reach of foreground have no bit words)
    s.score = sp;
    if (!(ctx->last_valid && ctx->last_stream == st)) ctx->in_flight.clear();  // everything before this batch has completed
    else if (ctx->last_fused_tail && ctx->last_done) {
        ctx->in_flight.insert(ctx->in_flight.begin(), hv_ctx::InFlight{ctx->last_mask, ctx->last_labels, ctx->last_done, ctx->last_expected});
        if (ctx->in_flight.size() > (size_t)(kSyncSlots - 2)) ctx->in_flight.resize(kSyncSlots - 2);
    }
    ctx->last_valid = true;
    ctx->last_stream = st;
    ctx->last_fused_tail = fused;
    ctx->last_mask = b.mask;
    ctx->last_labels = b.labels;
    ctx->last_done = fused ? b.ccl_done : nullptr;
    ctx->last_expected = s.ccl_expected;
    s.view = b;
    s.has_batch = true;
    s.have_blur = (separate_blur && !gauss_fused) || want_blur;
    s.c = c;
    s.d_input = d_frames;
    return HV_OK;
}

// defect tables up to this size travel with every batch's results; larger ones (huge max_defects_per_frame x batch) are
// fetched on demand, only the rows in use
constexpr size_t kEagerDefectBytes = 2u << 20;

// Read-back of results + frame flags (one copy: they share a buffer) and, when `with_defects`, the defect table.
hv_status enqueue_readback(hv_ctx *ctx, Slot &s, cudaStream_t st, bool with_defects = true) {
    const BatchView &b = s.view;
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(s.h_results.p, b.results, sizeof(hv_frame_result) * b.n + sizeof(uint32_t) * b.n,
                                     cudaMemcpyDeviceToHost, st));
    if (with_defects)
        HV_TRY_CUDA(ctx, cudaMemcpyAsync(s.h_defects.p, b.defects, sizeof(hv_defect) * (size_t)b.n * b.defect_cap,
                                         cudaMemcpyDeviceToHost, st));
    s.defects_on_host = with_defects;
    return HV_OK;
}

// Device-resident batches: their results travel on the context's copy stream, which waits for the slot's completion
// counter (the per-frame CCL kernel bumps it once per frame) with a stream memory operation -- nothing is inserted
// between two kernels of the launching stream, so the kernels of consecutive batches keep overlapping.  Batches that
// did not go through the per-frame kernel (global-memory CCL path) are ordered by an event instead.
hv_status enqueue_async_readback(hv_ctx *ctx, Slot &s, cudaStream_t st) {
    const BatchView &b = s.view;
    bool ordered = false;
    if (s.used_fused && b.ccl_done && ctx->wait32)
        ordered = ctx->wait32(reinterpret_cast<CUstream>(ctx->copy_stream), reinterpret_cast<CUdeviceptr>(b.ccl_done),
                              s.ccl_expected, CU_STREAM_WAIT_VALUE_GEQ) == CUDA_SUCCESS;
    if (!ordered) {
        HV_TRY_CUDA(ctx, cudaEventRecord(s.done, st));
        HV_TRY_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, s.done, 0));
        ctx->last_valid = false;  // the event sits between this batch and the next: plain stream order across it
    }
    const bool eager = sizeof(hv_defect) * (size_t)b.n * b.defect_cap <= kEagerDefectBytes;
    hv_status rs = enqueue_readback(ctx, s, ctx->copy_stream, eager);
    if (rs != HV_OK) return rs;
    HV_TRY_CUDA(ctx, cudaEventRecord(s.copied, ctx->copy_stream));
    return HV_OK;
}

// After the read-back has completed: if the fused kernel flagged frames it could not hold in shared memory, run the
// global-memory path on exactly those frames (same stream), read back again and wait.
hv_status resolve_fallback(hv_ctx *ctx, Slot &s, cudaStream_t st) {
    if (!s.used_fused) return HV_OK;
    const uint32_t *h_flags = flags_after(s.h_results.p, s.view.n);
    bool any = false;
    bool all_fit_small = true, all_fit_tiny = true;
    for (int f = 0; f < s.view.n; f++) {
        any |= (h_flags[f] & 1u) != 0;
        all_fit_small &= (h_flags[f] & 3u) == 0;
        all_fit_tiny &= h_flags[f] == 0;
    }
    if (!any) {
        ctx->dense_hint = false;
        if (!s.used_small && !s.used_tiny && all_fit_small) ctx->ccl_small_ok = true;  // the big build reports that the small one would do
        // every build reports whether the tiny one would do; after a batch that overflowed it, it stays off for a while so
        // that a line whose frames hover around its capacity does not pay the fallback again and again
        if (ctx->ccl_tiny_cooldown > 0) ctx->ccl_tiny_cooldown--;
        ctx->ccl_tiny_ok = all_fit_tiny && ctx->ccl_tiny_cooldown == 0;
        return HV_OK;
    }
    if (s.used_tiny) {  // too much foreground for the tiny build: back to the small one (the global path below finishes the
        ctx->ccl_tiny_ok = false;  // flagged frames of this batch)
        ctx->ccl_tiny_cooldown = 64;
    } else if (s.used_small) {  // ... for the small build: the big one takes over
        ctx->ccl_small_ok = false;
        ctx->ccl_small_retry = 0;
    }
    ctx->last_valid = false;  // kernels that are not part of the K1 / per-frame chain go onto the stream
    BatchView b = s.view;
    b.frame_select = b.frame_flags;
    if (s.sparse_bits) {
        HV_TRY_CUDA(ctx, launch_densify_bits(b, st));
        ctx->launches++;
        s.sparse_bits = false;  // (for the selected frames, which are the only ones anybody reads from now on)
    }
    hv_status rs = enqueue_global_ccl(ctx, b, s.score, st);
    if (rs != HV_OK) return rs;
    s.used_fused = false;
    rs = enqueue_readback(ctx, s, st);
    if (rs != HV_OK) return rs;
    HV_TRY_CUDA(ctx, cudaStreamSynchronize(st));
    if (!s.used_small && !s.used_tiny) {  // (after the smaller builds the big one is tried before the global path becomes the default)
        ctx->dense_hint = true;
    }
    s.used_small = false;
    s.used_tiny = false;
    return HV_OK;
}

// A device-resident batch whose results nobody has looked at yet: wait for its read-back, finish the frames the
// per-frame kernel flagged (global path), update the CCL path selection.  Runs before the slot is reused, before a
// batch that writes the same caller-owned output planes is enqueued, and when the batch is fetched.
// CCL path selection from the results of a batch that went through the global-memory kernels: back to the per-frame
// kernel once every frame of a batch is sparse enough for its big build with a wide margin (components and foreground
// pixels; the kernel's real limits are non-zero words and word-runs, which the per-frame records do not carry).
void update_dense_hint(hv_ctx *ctx, const Slot &s) {
    if (!ctx->dense_hint || s.used_fused || !s.has_batch) return;
    bool sparse = true;
    for (int f = 0; f < s.view.n; f++) {
        const hv_frame_result &r = s.h_results.p[f];
        sparse &= r.status == HV_OK && r.n_components <= 1024u && r.fg_pixels <= 16384u;
    }
    if (sparse) ctx->dense_hint = false;
}

hv_status retire_slot(hv_ctx *ctx, Slot &s) {
    if (!s.pending) return HV_OK;
    for (size_t k = 0; k < ctx->deferred.size(); k++)
        if (&ctx->slots[ctx->deferred[k].slot] == &s) {  // its tail (and the older ones) go onto the stream now
            hv_status rf = flush_deferred(ctx, ctx->deferred.size() - 1 - k);
            if (rf != HV_OK) return rf;
            break;
        }
    HV_TRY_CUDA(ctx, cudaEventSynchronize(s.copied));
    s.pending = false;
    update_dense_hint(ctx, s);
    return resolve_fallback(ctx, s, s.batch_stream);
}

// next scratch set of the synchronous / device-resident entry points (its previous batch is retired first)
hv_status advance_sync_slot(hv_ctx *ctx) {
    const int next = (ctx->sync_cur + 1) % kSyncSlots;
    hv_status rs = retire_slot(ctx, ctx->slots[next]);
    if (rs != HV_OK) return rs;
    ctx->sync_cur = next;
    return HV_OK;
}

// After the stream has been synchronised: compact the per-frame defect blocks into the caller's arrays.
hv_status unpack_results(hv_ctx *ctx, Slot &s, hv_frame_result *results, hv_defect *defects, size_t defects_cap,
                         size_t *n_total) {
    const BatchView &b = s.view;
    size_t off = 0;
    hv_status rc = HV_OK;
    for (int f = 0; f < b.n; f++) {
        hv_frame_result r = s.h_results.p[f];
        const hv_defect *src = s.h_defects.p + (size_t)f * b.defect_cap;
        uint32_t nd = r.n_defects;
        if (r.status != HV_OK) rc = HV_ERR_CAPACITY;
        if (off + nd > defects_cap) {
            nd = (uint32_t)(defects_cap > off ? defects_cap - off : 0);
            r.status = HV_ERR_CAPACITY;
            rc = HV_ERR_CAPACITY;
        }
        r.defects_offset = (uint32_t)off;
        if (defects && nd) std::memcpy(defects + off, src, sizeof(hv_defect) * nd);
        r.n_defects = nd;
        off += nd;
        if (results) results[f] = r;
    }
    if (n_total) *n_total = off;
    if (rc != HV_OK)
        ctx->err = "capacity exceeded: more components or defects than max_blobs_per_frame / max_defects_per_frame / "
                   "defects_cap (see per-frame status)";
    return rc;
}

hv_status copy_debug(hv_ctx *ctx, Slot &s, cudaStream_t st, const hv_debug_outputs *dbg) {
    if (!dbg) return HV_OK;
    const BatchView &b = s.view;
    const size_t px = (size_t)b.n * b.h * b.w;
    if (dbg->gray)
        HV_TRY_CUDA(ctx, cudaMemcpy2DAsync(dbg->gray, b.w, b.gray, b.gray_row_stride, b.w, (size_t)b.n * b.h,
                                           cudaMemcpyDeviceToHost, st));
    if (dbg->blur) {
        if (!s.have_blur) return fail(ctx, HV_ERR_INVALID_ARGUMENT, "blur intermediate was not kept for this batch");
        HV_TRY_CUDA(ctx, cudaMemcpyAsync(dbg->blur, s.blur.p, px, cudaMemcpyDeviceToHost, st));
    }
    if (dbg->mask) HV_TRY_CUDA(ctx, cudaMemcpyAsync(dbg->mask, b.mask, px, cudaMemcpyDeviceToHost, st));
    if (dbg->labels)
        HV_TRY_CUDA(ctx, cudaMemcpyAsync(dbg->labels, b.labels, px * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (dbg->blobs) {
        const size_t stride = dbg->blobs_stride ? dbg->blobs_stride : (size_t)b.blob_cap;
        const size_t rows = std::min(stride, (size_t)b.blob_cap);
        HV_TRY_CUDA(ctx, cudaMemcpy2DAsync(dbg->blobs, stride * sizeof(hv_blob), b.blobs, (size_t)b.blob_cap * sizeof(hv_blob),
                                           rows * sizeof(hv_blob), b.n, cudaMemcpyDeviceToHost, st));
    }
    HV_TRY_CUDA(ctx, cudaStreamSynchronize(st));
    return HV_OK;
}

hv_status upload_frames(hv_ctx *ctx, Slot &s, cudaStream_t st, const uint8_t *frames, int n, int h, int w, int c,
                        size_t row_stride, size_t frame_stride) {
    const size_t row_bytes = (size_t)w * c;
    if (row_stride == 0) row_stride = row_bytes;
    if (frame_stride == 0) frame_stride = row_stride * h;
    if (row_stride < row_bytes || frame_stride < row_stride * (size_t)h)
        return fail(ctx, HV_ERR_INVALID_ARGUMENT, "strides smaller than the frame");
    HV_TRY_CUDA(ctx, s.in.reserve((size_t)n * h * row_bytes));
    if (row_stride == row_bytes && frame_stride == row_bytes * h) {
        HV_TRY_CUDA(ctx, cudaMemcpyAsync(s.in.p, frames, (size_t)n * h * row_bytes, cudaMemcpyHostToDevice, st));
    } else if (frame_stride == row_stride * h) {
        HV_TRY_CUDA(ctx, cudaMemcpy2DAsync(s.in.p, row_bytes, frames, row_stride, row_bytes, (size_t)n * h,
                                           cudaMemcpyHostToDevice, st));
    } else {
        for (int f = 0; f < n; f++)
            HV_TRY_CUDA(ctx, cudaMemcpy2DAsync(s.in.p + (size_t)f * h * row_bytes, row_bytes, frames + f * frame_stride,
                                               row_stride, row_bytes, h, cudaMemcpyHostToDevice, st));
    }
    return HV_OK;
}


}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
extern "C" {

int32_t hv_abi_version(void) { return HV_ABI_VERSION; }
const char *hv_version(void) { return "heimdall-cuda 0.1.0 (sm_100a)"; }

const char *hv_status_string(hv_status s) {
    switch (s) {
        case HV_OK: return "ok";
        case HV_ERR_INVALID_DIMENSIONS: return "Invalid image dimensions: expected 3D array";
        case HV_ERR_INVALID_ARGUMENT: return "invalid argument";
        case HV_ERR_CUDA: return "CUDA error";
        case HV_ERR_CAPACITY: return "capacity exceeded";
        case HV_ERR_NO_DEVICE: return "no usable CUDA device (there is no CPU fallback)";
        case HV_ERR_UNSUPPORTED: return "unsupported";
        case HV_ERR_CHANNELS: return "stage requires a 1-channel image";
        case HV_ERR_BAD_TICKET: return "unknown or already consumed ticket";
        default: return "unknown status";
    }
}

int32_t hv_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void hv_params_default(hv_params *p) {
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->min_size = 10.0;       // lib.rs:106
    p->max_size = 3000.0;     // lib.rs:107
    p->threshold = 25.0;      // lib.rs:108
    p->min_confidence = 0.3;  // detection.rs:298
    p->gauss_sigma = 0.0;
    p->blur_mode = HV_BLUR_BOX;
    p->blur_ksize = 5;  // detection.rs:163 (radius 2)
    p->morph_open_k = 0;
    p->morph_close_k = 0;
}

void hv_config_default(hv_config *c) {
    if (!c) return;
    std::memset(c, 0, sizeof(*c));
}

hv_status hv_create(int32_t device, const hv_config *cfg, hv_ctx **out) {
    if (!out) return HV_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        g_create_error = std::string("no usable CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "0 devices") +
                         "); this backend has no CPU fallback";
        return HV_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= ndev) {
        g_create_error = "device index out of range";
        return HV_ERR_INVALID_ARGUMENT;
    }
    cudaDeviceProp prop{};
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e);
        return HV_ERR_CUDA;
    }
    if (prop.major != 10) {
        g_create_error = "device is not sm_100 (Blackwell B200); kernels are built for sm_100a only";
        return HV_ERR_NO_DEVICE;
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return HV_ERR_CUDA;
    }
    (void)tunables();  // the environment is read here, once per process
    hv_ctx *ctx = new (std::nothrow) hv_ctx();
    if (!ctx) return HV_ERR_INVALID_ARGUMENT;
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    if (cfg) ctx->cfg = *cfg;
    // slots 0 and 1: the synchronous / device-resident entry points, used alternately so that consecutive batches do not
    // share scratch and K1 of batch i+1 may overlap the per-frame CCL of batch i; slots 2..: hv_submit / hv_wait
    const int nslots = kSyncSlots + (ctx->cfg.num_slots > 0 ? ctx->cfg.num_slots : 3);
    ctx->slots.resize(nslots);
    for (auto &s : ctx->slots) {
        // the slot's own output planes (used when the caller passes no device buffers) live in compressible memory
        s.mask.want_compressible = s.labels.want_compressible = true;
        s.mask.device = s.labels.device = device;
        if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming) != cudaSuccess) {
            g_create_error = "stream/event creation failed";
            hv_destroy(ctx);
            return HV_ERR_CUDA;
        }
    }
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
        g_create_error = "stream creation failed";
        hv_destroy(ctx);
        return HV_ERR_CUDA;
    }
    ctx->wait32 = drv_fn<decltype(&cuStreamWaitValue32)>("cuStreamWaitValue32");
    if (ctx->cfg.flags & HV_FLAG_PROFILE) ctx->prof_mask = 0xffffffffu;
    if (configure_ccl_frame() != cudaSuccess || configure_preprocess_tma() != cudaSuccess) {
        g_create_error = "cannot configure shared memory for the per-frame CCL kernel";
        hv_destroy(ctx);
        return HV_ERR_CUDA;
    }
    if (cudaMalloc(reinterpret_cast<void **>(&ctx->d_stats), sizeof(hv_line_stats)) != cudaSuccess ||
        cudaMemset(ctx->d_stats, 0, sizeof(hv_line_stats)) != cudaSuccess) {
        g_create_error = "stats allocation failed";
        hv_destroy(ctx);
        return HV_ERR_CUDA;
    }
    if (cudaMalloc(reinterpret_cast<void **>(&ctx->d_phase_ns), 256 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(ctx->d_phase_ns, 0, 256 * sizeof(unsigned long long)) != cudaSuccess) {
        g_create_error = "debug buffer allocation failed";
        hv_destroy(ctx);
        return HV_ERR_CUDA;
    }
    *out = ctx;
    return HV_OK;
}

void hv_destroy(hv_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    flush_deferred(ctx);
    cudaDeviceSynchronize();
    for (auto &s : ctx->slots) s.release();
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (auto &r : ctx->prof_recs) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    for (auto e : ctx->prof_pool) cudaEventDestroy(e);
    if (ctx->d_stats) cudaFree(ctx->d_stats);
    if (ctx->d_phase_ns) cudaFree(ctx->d_phase_ns);
    ctx->u_a.release(), ctx->u_b.release(), ctx->u_c.release();
    ctx->u_centers.release(), ctx->u_contours.release(), ctx->u_count.release(), ctx->u_owner.release();
    while (!ctx->vmm.empty()) hv_device_free(ctx, ctx->vmm.begin()->first);  // buffers the caller did not return
    delete ctx;
}

const char *hv_last_error(const hv_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

hv_status hv_set_stream(hv_ctx *ctx, void *cuda_stream, int32_t enable) {
    if (!ctx) return HV_ERR_INVALID_ARGUMENT;
    if (!ctx->deferred.empty()) {  // (they belong to the stream that is being replaced)
        cudaSetDevice(ctx->device);
        hv_status rf = flush_deferred(ctx);
        if (rf != HV_OK) return rf;
    }
    ctx->user_stream = enable ? reinterpret_cast<cudaStream_t>(cuda_stream) : nullptr;
    ctx->use_user_stream = enable != 0;
    return HV_OK;
}

void *hv_host_alloc(hv_ctx *ctx, size_t bytes) {
    if (!ctx) return nullptr;
    cudaSetDevice(ctx->device);
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        ctx->err = "cudaHostAlloc failed";
        return nullptr;
    }
    return p;
}

void hv_host_free(hv_ctx *ctx, void *p) {
    (void)ctx;
    if (p) cudaFreeHost(p);
}

int32_t hv_pipeline_depth(void) { return kSyncSlots; }

hv_status hv_device_alloc(hv_ctx *ctx, size_t bytes, uint32_t flags, void **d_ptr, int32_t *compressed_out) {
    if (!ctx || !d_ptr || bytes == 0) return fail(ctx, HV_ERR_INVALID_ARGUMENT, "bad argument");
    *d_ptr = nullptr;
    if (compressed_out) *compressed_out = 0;
    HV_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    HV_TRY_CUDA(ctx, cudaFree(nullptr));  // make sure the primary context exists before talking to the driver
    if (flags & HV_ALLOC_COMPRESSIBLE) {
        VmmRec rec;
        if (vmm_alloc_compressible(ctx->device, bytes, d_ptr, &rec)) {
            if (compressed_out) *compressed_out = rec.compressed ? 1 : 0;
            ctx->vmm[*d_ptr] = rec;
            return HV_OK;
        }
        // not supported / failed: ordinary memory below
    }
    HV_TRY_CUDA(ctx, cudaMalloc(d_ptr, bytes));
    return HV_OK;
}

hv_status hv_device_free(hv_ctx *ctx, void *d_ptr) {
    if (!ctx) return HV_ERR_INVALID_ARGUMENT;
    if (!d_ptr) return HV_OK;
    HV_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    HV_TRY_CUDA(ctx, cudaDeviceSynchronize());
    auto it = ctx->vmm.find(d_ptr);
    if (it == ctx->vmm.end()) {
        HV_TRY_CUDA(ctx, cudaFree(d_ptr));
        return HV_OK;
    }
    vmm_free(d_ptr, it->second);
    ctx->vmm.erase(it);
    return HV_OK;
}

hv_status hv_device_read(hv_ctx *ctx, void *host_dst, const void *d_src, size_t bytes) {
    if (!ctx || !host_dst || !d_src) return fail(ctx, HV_ERR_INVALID_ARGUMENT, "bad argument");
    HV_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = sync_stream(ctx);
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(host_dst, d_src, bytes, cudaMemcpyDeviceToHost, st));
    HV_TRY_CUDA(ctx, cudaStreamSynchronize(st));
    return HV_OK;
}

hv_status hv_device_write(hv_ctx *ctx, void *d_dst, const void *host_src, size_t bytes) {
    if (!ctx || !d_dst || !host_src) return fail(ctx, HV_ERR_INVALID_ARGUMENT, "bad argument");
    HV_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = sync_stream(ctx);
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(d_dst, host_src, bytes, cudaMemcpyHostToDevice, st));
    HV_TRY_CUDA(ctx, cudaStreamSynchronize(st));
    return HV_OK;
}

hv_status hv_enqueue_device(hv_ctx *ctx, const uint8_t *d_frames, int32_t n, int32_t h, int32_t w, int32_t c,
                            size_t row_stride, size_t frame_stride, const hv_params *params, uint8_t *d_mask,
                            int32_t *d_labels, int64_t *ticket) {
    if (!ctx || !d_frames) return HV_ERR_INVALID_ARGUMENT;
    if (ticket) *ticket = 0;
    hv_status rs = validate_shape(ctx, n, h, w, c);
    if (rs != HV_OK) return rs;
    HV_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = sync_stream(ctx);
    // The kernels of consecutive batches are ordered by device-side counters whose expected values are kernel
    // arguments: a captured graph would replay stale values and every wait would pass at once.  Refuse.
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) cudaGetLastError();
    if (cap != cudaStreamCaptureStatusNone)
        return fail(ctx, HV_ERR_UNSUPPORTED, "hv_enqueue_device cannot be captured into a CUDA graph (device-side batch counters)");
    hv_params pr;
    if (params)
        pr = *params;
    else
        hv_params_default(&pr);
    rs = advance_sync_slot(ctx);
    if (rs != HV_OK) return rs;
    // A batch still pending on another slot that wrote the same caller-owned planes must be finished first: if the
    // per-frame kernel flagged one of its frames, the global path completes it in those planes.
    if (d_mask || d_labels)
        for (int k = 0; k < kSyncSlots; k++) {
            Slot &q = ctx->slots[k];
            if (k != ctx->sync_cur && q.pending && q.has_batch &&
                ((d_mask && q.view.mask == d_mask) || (d_labels && q.view.labels == d_labels))) {
                rs = retire_slot(ctx, q);
                if (rs != HV_OK) return rs;
            }
        }
    Slot &s = cur_sync_slot(ctx);
    s.has_batch = false;
    rs = enqueue_pipeline(ctx, s, st, d_frames, n, h, w, c, row_stride, frame_stride, pr, d_mask, d_labels,
                          (ctx->cfg.flags & 4u) != 0, true);
    if (rs != HV_OK) return rs;
    s.batch_stream = st;
    s.ticket = ctx->next_ticket++;
    if (ticket) *ticket = s.ticket;
    if (tunables().exp_k1_only) return HV_OK;  // (experiment builds: nothing computes results)
    if (!s.tail_deferred) {  // (a deferred tail brings its read-back along: flush_deferred)
        rs = enqueue_async_readback(ctx, s, s.tail_stream);
        if (rs != HV_OK) return rs;
    }
    s.pending = true;
    return HV_OK;
}

hv_status hv_flush(hv_ctx *ctx) {
    if (!ctx) return HV_ERR_INVALID_ARGUMENT;
    HV_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    hv_status rs = flush_deferred(ctx);
    if (rs != HV_OK) return rs;
    if (ctx->tail_used) {  // the launching stream waits for the kernels that went onto slot streams
        cudaStream_t st = sync_stream(ctx);
        for (int k = 0; k < kSyncSlots; k++) {
            Slot &q = ctx->slots[k];
            if (!q.tail_on_side || q.stream == st) continue;
            HV_TRY_CUDA(ctx, cudaEventRecord(q.done, q.stream));
            HV_TRY_CUDA(ctx, cudaStreamWaitEvent(st, q.done, 0));
            q.tail_on_side = false;
        }
        ctx->tail_used = false;
        ctx->last_valid = false;
    }
    return HV_OK;
}

hv_status hv_fetch_ticket(hv_ctx *ctx, int64_t ticket, hv_frame_result *results, hv_defect *defects, size_t defects_cap,
                          size_t *n_defects_total) {
    if (!ctx) return HV_ERR_INVALID_ARGUMENT;
    Slot *sp = nullptr;
    for (int k = 0; k < kSyncSlots; k++)
        if (ctx->slots[k].has_batch && ctx->slots[k].ticket == ticket && ticket > 0) sp = &ctx->slots[k];
    if (!sp)
        return fail(ctx, HV_ERR_BAD_TICKET, "unknown ticket, or the batch's scratch set has been reused (only the last "
                                            "hv_pipeline_depth() batches can be fetched)");
    Slot &s = *sp;
    HV_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    hv_status rs = retire_slot(ctx, s);
    if (rs != HV_OK) return rs;
    if (!s.defects_on_host) {  // large defect tables: only the rows in use, now
        const BatchView &b = s.view;
        for (int f = 0; f < b.n; f++) {
            const uint32_t nd = std::min<uint32_t>(s.h_results.p[f].n_defects, (uint32_t)b.defect_cap);
            if (nd)
                HV_TRY_CUDA(ctx, cudaMemcpyAsync(s.h_defects.p + (size_t)f * b.defect_cap, b.defects + (size_t)f * b.defect_cap,
                                                 sizeof(hv_defect) * nd, cudaMemcpyDeviceToHost, ctx->copy_stream));
        }
        HV_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
        s.defects_on_host = true;
    }
    return unpack_results(ctx, s, results, defects, defects_cap, n_defects_total);
}

hv_status hv_fetch_results(hv_ctx *ctx, hv_frame_result *results, hv_defect *defects, size_t defects_cap,
                           size_t *n_defects_total) {
    if (!ctx) return HV_ERR_INVALID_ARGUMENT;
    Slot &s = cur_sync_slot(ctx);
    if (!s.has_batch) return fail(ctx, HV_ERR_INVALID_ARGUMENT, "no batch has been enqueued");
    return hv_fetch_ticket(ctx, s.ticket, results, defects, defects_cap, n_defects_total);
}

hv_status hv_fetch_debug(hv_ctx *ctx, const hv_debug_outputs *debug) {
    if (!ctx) return HV_ERR_INVALID_ARGUMENT;
    Slot &s = cur_sync_slot(ctx);
    if (!s.has_batch) return fail(ctx, HV_ERR_INVALID_ARGUMENT, "no batch has been enqueued");
    HV_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->last_valid = false;  // the copies below go onto the launching stream
    // flagged frames must have been completed by the global path before anything is copied
    hv_status rs = retire_slot(ctx, s);
    if (rs != HV_OK) return rs;
    return copy_debug(ctx, s, s.batch_stream ? s.batch_stream : sync_stream(ctx), debug);
}

hv_status hv_detect_batch_device(hv_ctx *ctx, const uint8_t *d_frames, int32_t n, int32_t h, int32_t w, int32_t c,
                                 size_t row_stride, size_t frame_stride, const hv_params *params, uint8_t *d_mask,
                                 int32_t *d_labels, hv_frame_result *results, hv_defect *defects, size_t defects_cap,
                                 size_t *n_defects_total) {
    int64_t ticket = 0;
    hv_status rs = hv_enqueue_device(ctx, d_frames, n, h, w, c, row_stride, frame_stride, params, d_mask, d_labels, &ticket);
    if (rs != HV_OK) return rs;
    return hv_fetch_ticket(ctx, ticket, results, defects, defects_cap, n_defects_total);
}

hv_status hv_detect_batch(hv_ctx *ctx, const uint8_t *frames, int32_t n, int32_t h, int32_t w, int32_t c,
                          size_t row_stride, size_t frame_stride, const hv_params *params, hv_frame_result *results,
                          hv_defect *defects, size_t defects_cap, size_t *n_defects_total,
                          const hv_debug_outputs *debug) {
    if (!ctx || !frames) return HV_ERR_INVALID_ARGUMENT;
    hv_status rs = validate_shape(ctx, n, h, w, c);
    if (rs != HV_OK) return rs;
    HV_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    hv_params pr;
    if (params)
        pr = *params;
    else
        hv_params_default(&pr);
    rs = advance_sync_slot(ctx);
    if (rs != HV_OK) return rs;
    ctx->last_valid = false;
    Slot &s = cur_sync_slot(ctx);
    s.has_batch = false;
    cudaStream_t st = sync_stream(ctx);
    s.batch_stream = st;
    s.ticket = ctx->next_ticket++;
    rs = upload_frames(ctx, s, st, frames, n, h, w, c, row_stride, frame_stride);
    if (rs != HV_OK) return rs;
    const bool want_blur = (debug && debug->blur) || (ctx->cfg.flags & 4u);
    rs = enqueue_pipeline(ctx, s, st, s.in.p, n, h, w, c, 0, 0, pr, nullptr, nullptr, want_blur);
    if (rs != HV_OK) return rs;
    rs = enqueue_readback(ctx, s, st);
    if (rs != HV_OK) return rs;
    HV_TRY_CUDA(ctx, cudaStreamSynchronize(st));
    update_dense_hint(ctx, s);
    rs = resolve_fallback(ctx, s, st);
    if (rs != HV_OK) return rs;
    hv_status rc = unpack_results(ctx, s, results, defects, defects_cap, n_defects_total);
    hv_status rd = copy_debug(ctx, s, st, debug);
    return rd != HV_OK ? rd : rc;
}

hv_status hv_submit(hv_ctx *ctx, const uint8_t *frames, int32_t n, int32_t h, int32_t w, int32_t c, size_t row_stride,
                    size_t frame_stride, const hv_params *params, int64_t *ticket) {
    if (!ctx || !frames || !ticket) return HV_ERR_INVALID_ARGUMENT;
    hv_status rs = validate_shape(ctx, n, h, w, c);
    if (rs != HV_OK) return rs;
    HV_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    hv_params pr;
    if (params)
        pr = *params;
    else
        hv_params_default(&pr);
    // find a free asynchronous slot (slots 1..)
    int pick = -1;
    const int nslots = (int)ctx->slots.size();
    const int nasync = nslots - kSyncSlots;
    for (int k = 0; k < nasync; k++) {
        const int idx = kSyncSlots + (ctx->next_slot + k) % nasync;
        if (ctx->slots[idx].ticket < 0) {
            pick = idx;
            break;
        }
    }
    if (pick < 0) return fail(ctx, HV_ERR_CAPACITY, "all slots are in flight: call hv_wait first");
    ctx->next_slot = (pick - kSyncSlots + 1) % nasync;
    Slot &s = ctx->slots[pick];
    rs = upload_frames(ctx, s, s.stream, frames, n, h, w, c, row_stride, frame_stride);
    if (rs != HV_OK) return rs;
    rs = enqueue_pipeline(ctx, s, s.stream, s.in.p, n, h, w, c, 0, 0, pr, nullptr, nullptr, false);
    if (rs != HV_OK) return rs;
    rs = enqueue_readback(ctx, s, s.stream);
    if (rs != HV_OK) return rs;
    HV_TRY_CUDA(ctx, cudaEventRecord(s.done, s.stream));
    s.ticket = ctx->next_ticket++;
    *ticket = s.ticket;
    return HV_OK;
}

hv_status hv_wait(hv_ctx *ctx, int64_t ticket, hv_frame_result *results, hv_defect *defects, size_t defects_cap,
                  size_t *n_defects_total) {
    if (!ctx) return HV_ERR_INVALID_ARGUMENT;
    for (size_t i = kSyncSlots; i < ctx->slots.size(); i++) {
        Slot &s = ctx->slots[i];
        if (s.ticket == ticket && ticket > 0) {
            HV_TRY_CUDA(ctx, cudaEventSynchronize(s.done));
            s.ticket = -1;
            update_dense_hint(ctx, s);
            hv_status rf = resolve_fallback(ctx, s, s.stream);
            if (rf != HV_OK) return rf;
            return unpack_results(ctx, s, results, defects, defects_cap, n_defects_total);
        }
    }
    return fail(ctx, HV_ERR_BAD_TICKET, "unknown or already consumed ticket");
}

// ---- frame feed (N1) --------------------------------------------------------------------------------------------------
int32_t hv_frame_channels(int32_t fmt) {
    switch (fmt) {
        case HV_PIX_MONO8: return 1;
        case HV_PIX_RGB8:
        case HV_PIX_BGR8: return 3;
        case HV_PIX_RGBA8:
        case HV_PIX_BGRA8: return 4;
        case HV_PIX_YUV422:
        case HV_PIX_YUV422_PACKED:
        case HV_PIX_BAYER_RG8:
        case HV_PIX_BAYER_GB8:
        case HV_PIX_BAYER_GR8:
        case HV_PIX_BAYER_BG8: return 3;
        default: return 0;
    }
}

}  // extern "C"

namespace {

bool is_bayer(int32_t fmt) { return fmt >= HV_PIX_BAYER_RG8 && fmt <= HV_PIX_BAYER_BG8; }
bool is_yuyv(int32_t fmt) { return fmt == HV_PIX_YUV422 || fmt == HV_PIX_YUV422_PACKED; }

// bytes per pixel of the raw frame, 0 = unknown format
size_t raw_bytes_per_px(int32_t fmt) {
    switch (fmt) {
        case HV_PIX_MONO8: return 1;
        case HV_PIX_MONO16: return 2;
        case HV_PIX_RGB8:
        case HV_PIX_BGR8: return 3;
        case HV_PIX_RGBA8:
        case HV_PIX_BGRA8: return 4;
        case HV_PIX_YUV422:
        case HV_PIX_YUV422_PACKED: return 2;
        default: return is_bayer(fmt) ? 1 : 0;
    }
}

hv_status check_frame(hv_ctx *ctx, const hv_camera_frame &fr) {
    if (!fr.data || fr.width == 0 || fr.height == 0) return fail(ctx, HV_ERR_INVALID_ARGUMENT, "empty camera frame");
    const size_t bpp = raw_bytes_per_px(fr.pixel_format);
    if (!bpp) return fail(ctx, HV_ERR_INVALID_ARGUMENT, "unknown pixel format");
    if ((unsigned long long)fr.width * fr.height >= 2147483647ULL) return fail(ctx, HV_ERR_INVALID_ARGUMENT, "frame too large");
    if (fr.size < (size_t)fr.width * fr.height * bpp)
        return fail(ctx, HV_ERR_INVALID_ARGUMENT, "camera frame data shorter than width * height * bytes per pixel");
    if (is_yuyv(fr.pixel_format) && (fr.width & 1))
        return fail(ctx, HV_ERR_INVALID_ARGUMENT, "YUV422 frames need an even width");
    return HV_OK;
}

}  // namespace

extern "C" {

hv_status hv_convert_frame(hv_ctx *ctx, const hv_camera_frame *frame, uint8_t *out, int32_t *out_channels) {
    if (!ctx || !frame || !out) return HV_ERR_INVALID_ARGUMENT;
    hv_status rs = check_frame(ctx, *frame);
    if (rs != HV_OK) return rs;
    const int32_t fmt = frame->pixel_format;
    const int32_t oc = hv_frame_channels(fmt);
    if (!oc)  // lib.rs:266-269 / 217-219
        return fail(ctx, HV_ERR_UNSUPPORTED, "Erreur de conversion d'image: Format de pixel non supporte pour la conversion");
    if (out_channels) *out_channels = oc;
    const int h = (int)frame->height, w = (int)frame->width;
    const size_t px = (size_t)h * w;
    if (!is_bayer(fmt) && !is_yuyv(fmt)) {  // to_ndarray: the bytes are the image
        std::memcpy(out, frame->data, px * oc);
        return HV_OK;
    }
    HV_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[0].stream;
    const size_t raw = px * raw_bytes_per_px(fmt);
    HV_TRY_CUDA(ctx, ctx->u_a.reserve(raw));
    HV_TRY_CUDA(ctx, ctx->u_b.reserve(px * 3));
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(ctx->u_a.p, frame->data, raw, cudaMemcpyHostToDevice, st));
    if (is_bayer(fmt))
        HV_TRY_CUDA(ctx, launch_bayer(ctx->u_a.p, 1, h, w, fmt - HV_PIX_BAYER_RG8, false, ctx->u_b.p, st));
    else
        HV_TRY_CUDA(ctx, launch_yuyv(ctx->u_a.p, 1, h, w, false, ctx->u_b.p, st));
    ctx->launches++;
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(out, ctx->u_b.p, px * 3, cudaMemcpyDeviceToHost, st));
    HV_TRY_CUDA(ctx, cudaStreamSynchronize(st));
    return HV_OK;
}

hv_status hv_submit_frames(hv_ctx *ctx, const hv_camera_frame *frames, int32_t n, const hv_params *params,
                           int64_t *ticket) {
    if (!ctx || !frames || !ticket || n <= 0) return HV_ERR_INVALID_ARGUMENT;
    const int32_t fmt = frames[0].pixel_format;
    const int h = (int)frames[0].height, w = (int)frames[0].width;
    for (int f = 0; f < n; f++) {
        hv_status rs = check_frame(ctx, frames[f]);
        if (rs != HV_OK) return rs;
        if (frames[f].pixel_format != fmt || (int)frames[f].height != h || (int)frames[f].width != w)
            return fail(ctx, HV_ERR_INVALID_ARGUMENT, "all frames of a batch must share geometry and pixel format");
    }
    const int32_t oc = hv_frame_channels(fmt);
    if (!oc) return fail(ctx, HV_ERR_UNSUPPORTED, "Erreur de conversion d'image: Format de pixel non supporte pour la conversion");
    if (oc != 1 && oc != 3) return fail(ctx, HV_ERR_INVALID_DIMENSIONS, "Invalid image dimensions: expected 3D array");
    HV_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    hv_params pr;
    if (params)
        pr = *params;
    else
        hv_params_default(&pr);
    int pick = -1;
    const int nslots = (int)ctx->slots.size();
    const int nasync = nslots - kSyncSlots;
    for (int k = 0; k < nasync; k++) {
        const int idx = kSyncSlots + (ctx->next_slot + k) % nasync;
        if (ctx->slots[idx].ticket < 0) {
            pick = idx;
            break;
        }
    }
    if (pick < 0) return fail(ctx, HV_ERR_CAPACITY, "all slots are in flight: call hv_wait first");
    ctx->next_slot = (pick - kSyncSlots + 1) % nasync;
    Slot &s = ctx->slots[pick];
    cudaStream_t st = s.stream;
    const size_t px = (size_t)h * w;
    const size_t raw = px * raw_bytes_per_px(fmt);
    HV_TRY_CUDA(ctx, s.in.reserve(raw * n));
    // frames already in page-locked memory go straight to the device; pageable ones pass through the slot's pinned
    // staging buffer so that the copy engine, not a hidden driver bounce buffer, moves them
    bool all_pinned = true;
    for (int f = 0; f < n && all_pinned; f++) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, frames[f].data) != cudaSuccess || at.type != cudaMemoryTypeHost) all_pinned = false;
    }
    cudaGetLastError();
    if (!all_pinned) {
        HV_TRY_CUDA(ctx, s.h_stage.reserve(raw * n));
        for (int f = 0; f < n; f++) std::memcpy(s.h_stage.p + raw * f, frames[f].data, raw);
        HV_TRY_CUDA(ctx, cudaMemcpyAsync(s.in.p, s.h_stage.p, raw * n, cudaMemcpyHostToDevice, st));
    } else {
        for (int f = 0; f < n; f++)
            HV_TRY_CUDA(ctx, cudaMemcpyAsync(s.in.p + raw * f, frames[f].data, raw, cudaMemcpyHostToDevice, st));
    }
    hv_status rs;
    if (is_bayer(fmt) || is_yuyv(fmt)) {
        // mosaic / YUV -> the gray the detector's A1 stage computes from the converted RGB image, in one kernel
        HV_TRY_CUDA(ctx, s.gray.reserve(px * n));
        ProfScope ps(ctx, HV_K_GRAY, st);
        if (is_bayer(fmt))
            HV_TRY_CUDA(ctx, launch_bayer(s.in.p, n, h, w, fmt - HV_PIX_BAYER_RG8, true, s.gray.p, st));
        else
            HV_TRY_CUDA(ctx, launch_yuyv(s.in.p, n, h, w, true, s.gray.p, st));
        ctx->launches++;
        rs = enqueue_pipeline(ctx, s, st, s.gray.p, n, h, w, 1, 0, 0, pr, nullptr, nullptr, false);
    } else {
        rs = enqueue_pipeline(ctx, s, st, s.in.p, n, h, w, oc, 0, 0, pr, nullptr, nullptr, false);
    }
    if (rs != HV_OK) return rs;
    rs = enqueue_readback(ctx, s, st);
    if (rs != HV_OK) return rs;
    HV_TRY_CUDA(ctx, cudaEventRecord(s.done, st));
    s.ticket = ctx->next_ticket++;
    *ticket = s.ticket;
    return HV_OK;
}

// ---- single-frame utilities ---------------------------------------------------------------------------------------
hv_status hv_preprocess_image(hv_ctx *ctx, const uint8_t *img, int32_t h, int32_t w, int32_t c, int32_t grayscale,
                              int32_t blur_size, uint8_t *out) {
    if (!ctx || !img || !out || h <= 0 || w <= 0 || c <= 0) return fail(ctx, HV_ERR_INVALID_ARGUMENT, "bad argument");
    if (grayscale && c < 3)
        return fail(ctx, HV_ERR_INVALID_DIMENSIONS,
                    "Invalid image dimensions: grayscale conversion indexes channels 0..2 (processing.rs:51-53)");
    HV_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[0].stream;
    const size_t in_bytes = (size_t)h * w * c;
    const int och = grayscale ? 1 : c;
    const size_t out_bytes = (size_t)h * w * och;
    HV_TRY_CUDA(ctx, ctx->u_a.reserve(in_bytes));
    HV_TRY_CUDA(ctx, ctx->u_b.reserve(out_bytes));
    HV_TRY_CUDA(ctx, ctx->u_c.reserve(out_bytes));
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(ctx->u_a.p, img, in_bytes, cudaMemcpyHostToDevice, st));
    const uint8_t *cur = ctx->u_a.p;
    if (grayscale) {
        HV_TRY_CUDA(ctx, launch_gray_first3(ctx->u_a.p, h, w, c, ctx->u_b.p, st));
        ctx->launches++;
        cur = ctx->u_b.p;
    }
    if (blur_size > 0) {
        HV_TRY_CUDA(ctx, launch_box_blur_generic(cur, h, w, och, blur_size / 2, ctx->u_c.p, st));
        ctx->launches++;
        cur = ctx->u_c.p;
    }
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(out, cur, out_bytes, cudaMemcpyDeviceToHost, st));
    HV_TRY_CUDA(ctx, cudaStreamSynchronize(st));
    return HV_OK;
}

hv_status hv_apply_threshold(hv_ctx *ctx, const uint8_t *img, int32_t h, int32_t w, int32_t c, uint8_t threshold_value,
                             int32_t adaptive, int32_t inverse, uint8_t *out) {
    if (!ctx || !img || !out || h <= 0 || w <= 0) return fail(ctx, HV_ERR_INVALID_ARGUMENT, "bad argument");
    if (c != 1) return fail(ctx, HV_ERR_CHANNELS, "Image processing error: Thresholding requires a grayscale image");
    HV_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    Slot &s = ctx->slots[0];
    if (retire_slot(ctx, s) != HV_OK) return HV_ERR_CUDA;  // the scratch set may hold a device-resident batch
    cudaStream_t st = s.stream;
    const size_t px = (size_t)h * w;
    hv_status rs = reserve_slot(ctx, s, 1, h, w, true, px, false, false, false, true, true);
    if (rs != HV_OK) return rs;
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(s.in.p, img, px, cudaMemcpyHostToDevice, st));
    if (adaptive) {
        BatchView b{};
        b.n = 1, b.h = h, b.w = w, b.ww = (w + 31) / 32;
        b.gray = s.in.p, b.gray_row_stride = w, b.gray_frame_stride = px;
        b.mask = s.mask.p, b.labels = s.labels.p, b.bits = s.bits.p;
        PreprocessParams pp{};
        pp.c_thresh = 2;  // processing.rs:134
        pp.blur_radius = 0;
        pp.write_mask = 1;
        pp.init_labels = 0;
        pp.inverse = inverse ? 1 : 0;
        HV_TRY_CUDA(ctx, launch_preprocess(b, pp, s.bits.p, st));
    } else {
        HV_TRY_CUDA(ctx, launch_threshold_generic(s.in.p, h, w, 0, threshold_value, inverse, s.mask.p, st));
    }
    ctx->launches++;
    s.has_batch = false;
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(out, s.mask.p, px, cudaMemcpyDeviceToHost, st));
    HV_TRY_CUDA(ctx, cudaStreamSynchronize(st));
    return HV_OK;
}

// cv2.morphologyEx(MORPH_OPEN, rect k_open) then (MORPH_CLOSE, rect k_close) on a binary mask
// (heimdall/detectors/contamination_detector.py:81-87; heimdall/core/pipeline.py:290-332 MorphologyStage).
hv_status hv_morphology(hv_ctx *ctx, const uint8_t *mask, int32_t h, int32_t w, int32_t open_k, int32_t close_k,
                        uint8_t *out) {
    if (!ctx || !mask || !out || h <= 0 || w <= 0) return fail(ctx, HV_ERR_INVALID_ARGUMENT, "bad argument");
    if (open_k < 0 || open_k > 31 || close_k < 0 || close_k > 31)
        return fail(ctx, HV_ERR_INVALID_ARGUMENT, "morphology kernel size must be in [0, 31]");
    if ((long long)h * w >= 2147483647LL) return fail(ctx, HV_ERR_INVALID_ARGUMENT, "frame too large");
    HV_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    Slot &s = ctx->slots[0];
    if (retire_slot(ctx, s) != HV_OK) return HV_ERR_CUDA;  // the scratch set may hold a device-resident batch
    cudaStream_t st = s.stream;
    const size_t px = (size_t)h * w;
    hv_status rs = reserve_slot(ctx, s, 1, h, w, true, px, false, false, false, true, true);
    if (rs != HV_OK) return rs;
    BatchView b{};
    b.n = 1, b.h = h, b.w = w, b.ww = (w + 31) / 32;
    b.mask = s.mask.p, b.bits = s.bits.p, b.bits_tmp = s.bits_tmp.p, b.labels = s.labels.p;
    b.rowflags = s.rowflags.p, b.tiles_x = (w + 127) / 128, b.rf_stride = (size_t)((h + 31) / 32) * b.tiles_x * 32;
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(s.in.p, mask, px, cudaMemcpyHostToDevice, st));
    HV_TRY_CUDA(ctx, launch_bits_from_gt127(s.in.p, 1, h, w, b.ww, b.bits, st));
    int nl = 0;
    HV_TRY_CUDA(ctx, launch_morph(b, open_k, close_k, &nl, st));
    HV_TRY_CUDA(ctx, launch_expand_bits(b, st));
    ctx->launches += 2 + nl;
    s.has_batch = false;
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(out, b.mask, px, cudaMemcpyDeviceToHost, st));
    HV_TRY_CUDA(ctx, cudaStreamSynchronize(st));
    return HV_OK;
}

static hv_status run_ccl_only(hv_ctx *ctx, Slot &s, cudaStream_t st, BatchView &b) {
    HV_TRY_CUDA(ctx, launch_ccl_merge(b, st));
    HV_TRY_CUDA(ctx, launch_ccl_flatten(b, st));
    HV_TRY_CUDA(ctx, launch_ccl_scan(b, st));
    HV_TRY_CUDA(ctx, launch_ccl_label(b, st));
    ctx->launches += 4;
    (void)s;
    return HV_OK;
}

static void fill_view(hv_ctx *ctx, Slot &s, BatchView &b, int h, int w) {
    b.n = 1, b.h = h, b.w = w, b.ww = (w + 31) / 32;
    b.mask = s.mask.p, b.bits = s.bits.p, b.bits_tmp = s.bits_tmp.p, b.labels = s.labels.p;
    b.rootbits = s.rootbits.p, b.rankbase = s.rankbase.p, b.ncomp = s.ncomp.p, b.fgcount = s.fgcount.p;
    b.segbase = s.segbase.p, b.nseg = (int)(((size_t)h * b.ww + 255) / 256);
    b.blobs = s.blobs.p, b.blob_cap = blob_cap_for(ctx, h, w);
    b.score_state = s.score_state.p, b.score_chunks = (b.blob_cap + 255) / 256;
    b.defects = s.defects.p, b.defect_cap = defect_cap_for(ctx);
    b.results = s.results.p, b.stats = ctx->d_stats;
    b.frame_flags = flags_after(s.results.p, 1), b.frame_select = nullptr;
}

hv_status hv_find_contours(hv_ctx *ctx, const uint8_t *img, int32_t h, int32_t w, int32_t c, double min_area,
                           double max_area, hv_contour *contours, size_t cap, size_t *n_contours, int32_t *labels) {
    if (!ctx || !img || h <= 0 || w <= 0) return fail(ctx, HV_ERR_INVALID_ARGUMENT, "bad argument");
    if (c != 1)
        return fail(ctx, HV_ERR_CHANNELS, "Detection error: Contour detection requires a grayscale or binary image");
    if ((long long)h * w >= 2147483647LL) return fail(ctx, HV_ERR_INVALID_ARGUMENT, "frame too large");
    HV_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    Slot &s = ctx->slots[0];
    if (retire_slot(ctx, s) != HV_OK) return HV_ERR_CUDA;  // the scratch set may hold a device-resident batch
    cudaStream_t st = s.stream;
    const size_t px = (size_t)h * w;
    hv_status rs = reserve_slot(ctx, s, 1, h, w, true, px, false, false, false, true, true);
    if (rs != HV_OK) return rs;
    BatchView b{};
    fill_view(ctx, s, b, h, w);
    const int ccap = (int)std::min<size_t>(cap, (size_t)b.blob_cap);
    HV_TRY_CUDA(ctx, ctx->u_contours.reserve(std::max(ccap, 1)));
    HV_TRY_CUDA(ctx, ctx->u_count.reserve(1));
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(s.in.p, img, px, cudaMemcpyHostToDevice, st));
    HV_TRY_CUDA(ctx, launch_bits_from_gt127(s.in.p, 1, h, w, b.ww, b.bits, st));
    HV_TRY_CUDA(ctx, launch_bits_to_mask_labels(b, st));
    ctx->launches += 2;
    rs = run_ccl_only(ctx, s, st, b);
    if (rs != HV_OK) return rs;
    HV_TRY_CUDA(ctx, launch_collect_contours(b, min_area, max_area, ctx->u_contours.p, ctx->u_count.p, ccap, st));
    ctx->launches++;
    uint32_t cnt = 0, ncomp = 0;
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(&cnt, ctx->u_count.p, sizeof(cnt), cudaMemcpyDeviceToHost, st));
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(&ncomp, b.ncomp, sizeof(ncomp), cudaMemcpyDeviceToHost, st));
    if (labels) HV_TRY_CUDA(ctx, cudaMemcpyAsync(labels, b.labels, px * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    HV_TRY_CUDA(ctx, cudaStreamSynchronize(st));
    s.has_batch = false;
    const uint32_t ncopy = std::min<uint32_t>(cnt, (uint32_t)ccap);
    if (contours && ncopy)
        HV_TRY_CUDA(ctx, cudaMemcpy(contours, ctx->u_contours.p, sizeof(hv_contour) * ncopy, cudaMemcpyDeviceToHost));
    if (n_contours) *n_contours = ncopy;
    if (ncomp > (uint32_t)b.blob_cap || cnt > (uint32_t)ccap)
        return fail(ctx, HV_ERR_CAPACITY, "capacity exceeded: more contours than max_blobs_per_frame / cap");
    return HV_OK;
}

hv_status hv_process_image(hv_ctx *ctx, const uint8_t *img, int32_t h, int32_t w, int32_t c, int32_t pipeline,
                           uint8_t *out_hw3, hv_center *contours, size_t cap, size_t *n_contours) {
    if (!ctx || !img || !out_hw3 || h <= 0 || w <= 0 || c <= 0) return fail(ctx, HV_ERR_INVALID_ARGUMENT, "bad argument");
    if (pipeline != HV_PIPELINE_BASIC && pipeline != HV_PIPELINE_CONTAMINATION)
        return fail(ctx, HV_ERR_UNSUPPORTED, "Unsupported pipeline type");
    if (c < 3)
        return fail(ctx, HV_ERR_INVALID_DIMENSIONS,
                    "Invalid image dimensions: the pipelines index channels 0..2 (processing.rs:195-197,259-261)");
    if ((long long)h * w >= 2147483647LL) return fail(ctx, HV_ERR_INVALID_ARGUMENT, "frame too large");
    HV_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    Slot &s = ctx->slots[0];
    if (retire_slot(ctx, s) != HV_OK) return HV_ERR_CUDA;  // the scratch set may hold a device-resident batch
    cudaStream_t st = s.stream;
    const size_t px = (size_t)h * w;
    hv_status rs = reserve_slot(ctx, s, 1, h, w, true, px * c, true, true, false, true, true);
    if (rs != HV_OK) return rs;
    HV_TRY_CUDA(ctx, ctx->u_a.reserve(px * 3));
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(s.in.p, img, px * c, cudaMemcpyHostToDevice, st));
    HV_TRY_CUDA(ctx, launch_gray_first3(s.in.p, h, w, c, s.gray.p, st));
    ctx->launches++;
    BatchView b{};
    fill_view(ctx, s, b, h, w);
    b.gray = s.gray.p, b.gray_row_stride = w, b.gray_frame_stride = px;
    size_t n_out = 0;
    hv_status rc = HV_OK;
    if (pipeline == HV_PIPELINE_BASIC) {
        HV_TRY_CUDA(ctx, launch_box_blur_generic(s.gray.p, h, w, 1, 2, s.blur.p, st));
        HV_TRY_CUDA(ctx, launch_threshold_generic(s.blur.p, h, w, 0, 127, 0, s.mask.p, st));
        HV_TRY_CUDA(ctx, launch_visualise(s.mask.p, h, w, nullptr, 0, ctx->u_a.p, st));
        ctx->launches += 3;
    } else {
        PreprocessParams pp{};
        pp.c_thresh = 15;  // processing.rs:293
        pp.blur_radius = 2;
        pp.write_mask = 1;
        pp.init_labels = 1;
        pp.inverse = 1;
        HV_TRY_CUDA(ctx, launch_preprocess(b, pp, b.bits, st));
        ctx->launches++;
        rs = run_ccl_only(ctx, s, st, b);
        if (rs != HV_OK) return rs;
        const int ccap = b.blob_cap;
        HV_TRY_CUDA(ctx, ctx->u_centers.reserve(ccap));
        HV_TRY_CUDA(ctx, ctx->u_count.reserve(1));
        HV_TRY_CUDA(ctx, launch_collect_centers(b, 3u, ctx->u_centers.p, ctx->u_count.p, ccap, st));  // processing.rs:355
        ctx->launches++;
        uint32_t cnt = 0, ncomp = 0;
        HV_TRY_CUDA(ctx, cudaMemcpyAsync(&cnt, ctx->u_count.p, sizeof(cnt), cudaMemcpyDeviceToHost, st));
        HV_TRY_CUDA(ctx, cudaMemcpyAsync(&ncomp, b.ncomp, sizeof(ncomp), cudaMemcpyDeviceToHost, st));
        HV_TRY_CUDA(ctx, cudaStreamSynchronize(st));
        if (ncomp > (uint32_t)b.blob_cap) rc = fail(ctx, HV_ERR_CAPACITY, "capacity exceeded: raise max_blobs_per_frame");
        const uint32_t ndraw = std::min<uint32_t>(cnt, (uint32_t)ccap);
        HV_TRY_CUDA(ctx, launch_visualise(s.mask.p, h, w, ctx->u_centers.p, (int)ndraw, ctx->u_a.p, st));
        ctx->launches += ndraw ? 2 : 1;
        n_out = std::min<size_t>(ndraw, cap);
        if (ndraw > cap) rc = fail(ctx, HV_ERR_CAPACITY, "capacity exceeded: contours array too small");
        if (contours && n_out)
            HV_TRY_CUDA(ctx, cudaMemcpyAsync(contours, ctx->u_centers.p, sizeof(hv_center) * n_out, cudaMemcpyDeviceToHost, st));
    }
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(out_hw3, ctx->u_a.p, px * 3, cudaMemcpyDeviceToHost, st));
    HV_TRY_CUDA(ctx, cudaStreamSynchronize(st));
    s.has_batch = false;
    if (n_contours) *n_contours = n_out;
    return rc;
}

// ---- Python-detector parity mode (N3) --------------------------------------------------------------------------------------
void hv_pydet_params_default(hv_pydet_params *p) {
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->contrast_threshold = 25.0;  // contamination_detector.py:36
    p->blur_ksize = 5;             // :66
    p->block_size = 11;            // :75
    p->morph_open_k = 3;           // :81-84
    p->morph_close_k = 3;          // :87
}

}  // extern "C"

namespace {

struct PyStages {
    BatchView b{};           // slot 0: bits / mask / labels (8-connected) / blobs of the final mask
    const uint8_t *d_gray = nullptr;
    const uint8_t *d_img = nullptr;
    uint32_t ncomp = 0;
};

// gray -> blur -> adaptive Gaussian threshold -> open / close -> 8-connected components, everything left on the device in
// slot 0 (the stream is synchronised on return)
hv_status pydet_run_stages(hv_ctx *ctx, const uint8_t *img, int32_t h, int32_t w, int32_t c, const hv_pydet_params *params,
                           PyStages *out) {
    if (!ctx || !img || h <= 0 || w <= 0) return fail(ctx, HV_ERR_INVALID_ARGUMENT, "bad argument");
    if (c != 1 && c != 3) return fail(ctx, HV_ERR_INVALID_DIMENSIONS, "Invalid image dimensions: expected 3D array");
    if ((long long)h * w >= 2147483647LL) return fail(ctx, HV_ERR_INVALID_ARGUMENT, "frame too large");
    hv_pydet_params pr;
    if (params)
        pr = *params;
    else
        hv_pydet_params_default(&pr);
    if (pr.block_size < 3 || pr.block_size > 31 || !(pr.block_size & 1))
        return fail(ctx, HV_ERR_INVALID_ARGUMENT, "block_size must be odd and in [3, 31]");
    if (pr.morph_open_k < 0 || pr.morph_open_k > 31 || pr.morph_close_k < 0 || pr.morph_close_k > 31)
        return fail(ctx, HV_ERR_INVALID_ARGUMENT, "morphology kernel size must be in [0, 31]");
    uint16_t gk[32];
    if (!gaussian_kernel_q8(pr.blur_ksize, 0.0, gk)) return fail(ctx, HV_ERR_INVALID_ARGUMENT, "Gaussian kernel size must be odd and in [1, 31]");
    // cv2.getGaussianKernel(block_size, 0, CV_32F): sigma = 0.3 * ((n - 1) * 0.5 - 1) + 0.8, computed in double, stored as float
    float fk[31];
    {
        const int n = pr.block_size;
        static const double t3[] = {0.25, 0.5, 0.25}, t5[] = {0.0625, 0.25, 0.375, 0.25, 0.0625},
                            t7[] = {0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125};
        const double *fixed = n == 3 ? t3 : n == 5 ? t5 : n == 7 ? t7 : nullptr;  // OpenCV's small_gaussian_tab
        if (fixed) {
            for (int i = 0; i < n; i++) fk[i] = (float)fixed[i];
        } else {
            const double sigma = ((n - 1) * 0.5 - 1) * 0.3 + 0.8;
            const double scale2x = -0.5 / (sigma * sigma);
            double kd[31], sum = 0;
            for (int i = 0; i < n; i++) {
                const double x = i - (n - 1) * 0.5;
                kd[i] = std::exp(scale2x * x * x);
                sum += kd[i];
            }
            for (int i = 0; i < n; i++) fk[i] = (float)(kd[i] / sum);
        }
    }
    HV_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    Slot &s = ctx->slots[0];
    if (retire_slot(ctx, s) != HV_OK) return HV_ERR_CUDA;
    cudaStream_t st = s.stream;
    const size_t px = (size_t)h * w;
    hv_status rs = reserve_slot(ctx, s, 1, h, w, true, px * c, true, true, true, true, true);
    if (rs != HV_OK) return rs;
    HV_TRY_CUDA(ctx, ctx->u_a.reserve(px * sizeof(float)));
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(s.in.p, img, px * c, cudaMemcpyHostToDevice, st));
    const uint8_t *d_gray = s.in.p;
    if (c == 3) {
        HV_TRY_CUDA(ctx, launch_gray_bgr_cv(s.in.p, h, w, s.gray.p, st));
        ctx->launches++;
        d_gray = s.gray.p;
    }
    HV_TRY_CUDA(ctx, launch_gaussian_blur(d_gray, 1, h, w, gk, pr.blur_ksize, s.blur.p, s.gauss_tmp.p, st));
    const double cf = std::floor(pr.contrast_threshold);  // THRESH_BINARY_INV: idelta = cvFloor(C)
    const int idelta = cf > 1e9 ? 1000000000 : (cf < -1e9 ? -1000000000 : (int)cf);
    HV_TRY_CUDA(ctx, launch_adaptive_gaussian(s.blur.p, h, w, fk, pr.block_size, idelta, reinterpret_cast<float *>(ctx->u_a.p),
                                              s.mask.p, st));
    BatchView b{};
    fill_view(ctx, s, b, h, w);
    b.rowflags = s.rowflags.p, b.tiles_x = (w + 127) / 128, b.rf_stride = (size_t)((h + 31) / 32) * b.tiles_x * 32;
    b.conn8 = 1;
    HV_TRY_CUDA(ctx, launch_bits_from_gt127(s.mask.p, 1, h, w, b.ww, b.bits, st));
    int nl = 0;
    HV_TRY_CUDA(ctx, launch_morph(b, pr.morph_open_k, pr.morph_close_k, &nl, st));
    HV_TRY_CUDA(ctx, launch_bits_to_mask_labels(b, st));
    ctx->launches += 6 + nl;
    rs = run_ccl_only(ctx, s, st, b);
    if (rs != HV_OK) return rs;
    uint32_t ncomp = 0;
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(&ncomp, b.ncomp, sizeof(ncomp), cudaMemcpyDeviceToHost, st));
    HV_TRY_CUDA(ctx, cudaStreamSynchronize(st));
    s.has_batch = false;
    if (ncomp > (uint32_t)b.blob_cap) return fail(ctx, HV_ERR_CAPACITY, "capacity exceeded: more components than max_blobs_per_frame");
    out->b = b;
    out->d_gray = d_gray;
    out->d_img = s.in.p;
    out->ncomp = ncomp;
    return HV_OK;
}

}  // namespace

extern "C" {

hv_status hv_python_detector_stages(hv_ctx *ctx, const uint8_t *img, int32_t h, int32_t w, int32_t c,
                                    const hv_pydet_params *params, uint8_t *gray, uint8_t *blurred, uint8_t *binary,
                                    int32_t *labels8, hv_blob *comps, size_t cap, size_t *n_comps) {
    PyStages ps;
    hv_status rs = pydet_run_stages(ctx, img, h, w, c, params, &ps);
    if (rs != HV_OK) return rs;
    Slot &s = ctx->slots[0];
    cudaStream_t st = s.stream;
    const size_t px = (size_t)h * w;
    if (gray) HV_TRY_CUDA(ctx, cudaMemcpyAsync(gray, ps.d_gray, px, cudaMemcpyDeviceToHost, st));
    if (blurred) HV_TRY_CUDA(ctx, cudaMemcpyAsync(blurred, s.blur.p, px, cudaMemcpyDeviceToHost, st));
    if (binary) HV_TRY_CUDA(ctx, cudaMemcpyAsync(binary, ps.b.mask, px, cudaMemcpyDeviceToHost, st));
    if (labels8) HV_TRY_CUDA(ctx, cudaMemcpyAsync(labels8, ps.b.labels, px * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    const size_t ncopy = std::min<size_t>(ps.ncomp, cap);
    if (comps && ncopy) HV_TRY_CUDA(ctx, cudaMemcpyAsync(comps, ps.b.blobs, sizeof(hv_blob) * ncopy, cudaMemcpyDeviceToHost, st));
    HV_TRY_CUDA(ctx, cudaStreamSynchronize(st));
    if (n_comps) *n_comps = ncopy;
    if (comps && ps.ncomp > cap) return fail(ctx, HV_ERR_CAPACITY, "capacity exceeded: more components than cap");
    return HV_OK;
}

void hv_pydet_score_params_default(hv_pydet_score_params *p) {
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->min_size = 10.0;         // contamination_detector.py:26
    p->max_size = 3000.0;       // :29
    p->min_confidence = 0.25;   // :35
    p->use_color = 1;           // :38
}

hv_status hv_python_detect(hv_ctx *ctx, const uint8_t *img, int32_t h, int32_t w, int32_t c, const hv_pydet_params *stage_params,
                           const hv_pydet_score_params *score_params, hv_pydefect *defects, size_t cap, size_t *n_defects,
                           size_t *n_contours) {
    if (n_defects) *n_defects = 0;
    if (n_contours) *n_contours = 0;
    hv_pydet_score_params sp;
    if (score_params)
        sp = *score_params;
    else
        hv_pydet_score_params_default(&sp);
    PyStages ps;
    hv_status rs = pydet_run_stages(ctx, img, h, w, c, stage_params, &ps);
    if (rs != HV_OK) return rs;
    const int n8 = (int)ps.ncomp;
    if (n8 == 0) return HV_OK;
    Slot &s0 = ctx->slots[0];
    Slot &s1 = ctx->slots[1];
    if (retire_slot(ctx, s1) != HV_OK) return HV_ERR_CUDA;
    cudaStream_t st = s0.stream;
    const size_t px = (size_t)h * w;
    // 4-connected components of the background (complement of the final mask), in the second scratch set
    rs = reserve_slot(ctx, s1, 1, h, w, false, 0, false, false, false, true, true);
    if (rs != HV_OK) return rs;
    s1.has_batch = false;
    BatchView b4{};
    fill_view(ctx, s1, b4, h, w);
    b4.rowflags = s1.rowflags.p, b4.tiles_x = (w + 127) / 128, b4.rf_stride = (size_t)((h + 31) / 32) * b4.tiles_x * 32;
    HV_TRY_CUDA(ctx, launch_invert_bits(ps.b.bits, b4.bits, h, b4.ww, w, st));
    HV_TRY_CUDA(ctx, launch_bits_to_mask_labels(b4, st));
    ctx->launches += 2;
    rs = run_ccl_only(ctx, s1, st, b4);
    if (rs != HV_OK) return rs;
    uint32_t n4u = 0;
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(&n4u, b4.ncomp, sizeof(n4u), cudaMemcpyDeviceToHost, st));
    HV_TRY_CUDA(ctx, cudaStreamSynchronize(st));
    if (n4u > (uint32_t)b4.blob_cap) return fail(ctx, HV_ERR_CAPACITY, "capacity exceeded: more background regions than max_blobs_per_frame");
    const int n4 = (int)n4u;
    // device scratch: first8 | first4 | above8 | above4 | root8 | root4 | ext | sums | trace
    const size_t o_first8 = 0, o_first4 = o_first8 + 4u * n8, o_above8 = o_first4 + 4u * n4, o_above4 = o_above8 + 4u * n8,
                 o_root8 = o_above4 + 4u * n4, o_root4 = o_root8 + 4u * n8, o_ext = o_root4 + 4u * n4,
                 o_sums = (o_ext + 4u * n8 + 15) & ~(size_t)15, o_trace = o_sums + 80u * n8, o_end = o_trace + 32u * n8;
    HV_TRY_CUDA(ctx, ctx->u_c.reserve(o_end + 64));
    uint8_t *base = ctx->u_c.p;
    auto u32p = [&](size_t o) { return reinterpret_cast<uint32_t *>(base + o); };
    auto i32p = [&](size_t o) { return reinterpret_cast<int32_t *>(base + o); };
    HV_TRY_CUDA(ctx, cudaMemsetAsync(base + o_first8, 0xff, 4u * (n8 + n4), st));
    HV_TRY_CUDA(ctx, launch_first_pixel(ps.b.labels, h, w, u32p(o_first8), st));
    HV_TRY_CUDA(ctx, launch_first_pixel(b4.labels, h, w, u32p(o_first4), st));
    HV_TRY_CUDA(ctx, launch_label_above(u32p(o_first8), n8, w, b4.labels, i32p(o_above8), st));
    HV_TRY_CUDA(ctx, launch_label_above(u32p(o_first4), n4, w, ps.b.labels, i32p(o_above4), st));
    ctx->launches += 4;
    std::vector<hv_blob> comps8(n8), comps4(n4);
    std::vector<uint32_t> first8(n8), first4(n4);
    std::vector<int32_t> above8(n8), above4(n4);
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(comps8.data(), ps.b.blobs, sizeof(hv_blob) * n8, cudaMemcpyDeviceToHost, st));
    if (n4) HV_TRY_CUDA(ctx, cudaMemcpyAsync(comps4.data(), b4.blobs, sizeof(hv_blob) * n4, cudaMemcpyDeviceToHost, st));
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(first8.data(), base + o_first8, 4u * n8, cudaMemcpyDeviceToHost, st));
    if (n4) HV_TRY_CUDA(ctx, cudaMemcpyAsync(first4.data(), base + o_first4, 4u * n4, cudaMemcpyDeviceToHost, st));
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(above8.data(), base + o_above8, 4u * n8, cudaMemcpyDeviceToHost, st));
    if (n4) HV_TRY_CUDA(ctx, cudaMemcpyAsync(above4.data(), base + o_above4, 4u * n4, cudaMemcpyDeviceToHost, st));
    HV_TRY_CUDA(ctx, cudaStreamSynchronize(st));
    // Enclosure: a background region that does not reach the image border is a hole of the foreground component above its
    // first pixel; a foreground component whose first pixel lies below a hole is nested in that hole's owner.  Everything
    // that encloses a component has an earlier first pixel, so one pass in raster order of the first pixels resolves it.
    std::vector<char> border4(n4);
    for (int j = 0; j < n4; j++)
        border4[j] = comps4[j].xmin == 0 || comps4[j].ymin == 0 || comps4[j].xmax == (uint32_t)w - 1 || comps4[j].ymax == (uint32_t)h - 1;
    std::vector<int32_t> root8(n8, 0), root4(n4, 0);
    {
        int i8 = 0, i4 = 0;  // labels of both planes are numbered in raster order of their first pixels
        while (i8 < n8 || i4 < n4) {
            const bool take8 = i4 >= n4 || (i8 < n8 && first8[i8] < first4[i4]);
            if (take8) {
                const int a = above8[i8];  // background label above the first pixel, 0 in the first row
                root8[i8] = (a <= 0 || border4[a - 1]) ? i8 + 1 : root4[a - 1];
                if (root8[i8] == 0) root8[i8] = i8 + 1;
                i8++;
            } else {
                const int a = above4[i4];  // foreground label above the hole's first pixel
                root4[i4] = (border4[i4] || a <= 0) ? 0 : root8[a - 1];
                i4++;
            }
        }
    }
    std::vector<int32_t> ext;
    for (int k = 0; k < n8; k++)
        if (root8[k] == k + 1) ext.push_back(k + 1);
    const int n_ext = (int)ext.size();
    if (n_contours) *n_contours = (size_t)n_ext;
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(base + o_root8, root8.data(), 4u * n8, cudaMemcpyHostToDevice, st));
    if (n4) HV_TRY_CUDA(ctx, cudaMemcpyAsync(base + o_root4, root4.data(), 4u * n4, cudaMemcpyHostToDevice, st));
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(base + o_ext, ext.data(), 4u * n_ext, cudaMemcpyHostToDevice, st));
    HV_TRY_CUDA(ctx, launch_region_sums(i32p(o_ext), n_ext, ps.b.blobs, ps.b.labels, b4.labels, i32p(o_root8), i32p(o_root4), h, w,
                                        ps.d_gray, c == 3 ? ps.d_img : nullptr, base + o_sums, st));
    HV_TRY_CUDA(ctx, launch_trace_contours(i32p(o_ext), n_ext, u32p(o_first8), ps.b.labels, b4.labels, i32p(o_root8), i32p(o_root4), h,
                                           w, base + o_trace, st));
    ctx->launches += 2;
    struct Sums {
        unsigned long long cnt_in, cnt_out, gray_in, gray_out, ch_in[3], ch_out[3];
    };
    struct Trace {
        double a00, a10, a01;
        uint32_t chain_len, reserved;
    };
    static_assert(sizeof(Sums) == 80 && sizeof(Trace) == 32, "device records");
    std::vector<Sums> sums(n_ext);
    std::vector<Trace> tr(n_ext);
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(sums.data(), base + o_sums, sizeof(Sums) * n_ext, cudaMemcpyDeviceToHost, st));
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(tr.data(), base + o_trace, sizeof(Trace) * n_ext, cudaMemcpyDeviceToHost, st));
    HV_TRY_CUDA(ctx, cudaStreamSynchronize(st));
    // contamination_detector.py:96-176, in cv2.findContours' order (the component found last comes first), in the
    // reference's expression order (doubles, no contraction)
    size_t nd = 0;
    hv_status rc = HV_OK;
    for (int e = n_ext - 1; e >= 0; e--) {
        const hv_blob &q = comps8[ext[e] - 1];
        const Trace &t = tr[e];
        const double area = std::fabs(t.a00 * 0.5);  // cv2.contourArea
        if (area < sp.min_size || area > sp.max_size) continue;
        const int bx = (int)q.xmin, by = (int)q.ymin, bw = (int)(q.xmax - q.xmin + 1), bh = (int)(q.ymax - q.ymin + 1);
        // cv2.moments of the contour: m00 = a00 * (+-0.5), m10 = a10 * (+-1/6), m01 = a01 * (+-1/6)
        double m00 = 0, m10 = 0, m01 = 0;
        if (std::fabs(t.a00) > 1.1920928955078125e-07) {
            const double h2 = t.a00 > 0 ? 0.5 : -0.5, h6 = t.a00 > 0 ? 0.16666666666666666666666666666667 : -0.16666666666666666666666666666667;
            m00 = t.a00 * h2, m10 = t.a10 * h6, m01 = t.a01 * h6;
        }
        if (!(m00 > 0)) continue;
        const int cx = (int)(m10 / m00), cy = (int)(m01 / m00);
        const Sums &u = sums[e];
        const double background = u.cnt_out ? (double)u.gray_out / (double)u.cnt_out : 127.0;
        const double foreground = u.cnt_in ? (double)u.gray_in / (double)u.cnt_in : 127.0;
        const double intensity_diff = std::fabs(background - foreground);
        const double intensity_score = std::min(1.0, intensity_diff / 30.0);
        const long long rect_area = (long long)bw * bh;
        const double area_ratio = rect_area > 0 ? area / (double)rect_area : 0.0;
        const double shape_score = 1.0 - area_ratio;
        double color_score = 0.5;
        if (sp.use_color && c == 3) {
            double color_diff = 0.0;
            for (int k = 0; k < 3; k++) {
                const double fg = u.cnt_in ? (double)u.ch_in[k] / (double)u.cnt_in : 127.0;
                const double bg = u.cnt_out ? (double)u.ch_out[k] / (double)u.cnt_out : 127.0;
                const double d = std::fabs(fg - bg);
                if (k == 0 || d > color_diff) color_diff = d;
            }
            color_score = std::min(1.0, color_diff / 30.0);
        }
        const double t1 = intensity_score * 0.5, t2 = shape_score * 0.2, t3 = color_score * 0.3;
        const double confidence = (t1 + t2) + t3;
        if (!(confidence >= sp.min_confidence)) continue;
        if (nd < cap && defects) {
            hv_pydefect &d = defects[nd];
            d.x = cx, d.y = cy;
            d.size = area, d.confidence = confidence;
            d.intensity_diff = intensity_diff, d.shape_score = shape_score, d.color_score = color_score;
            d.bx = bx, d.by = by, d.bw = bw, d.bh = bh;
            d.label8 = (uint32_t)ext[e];
            d.chain_len = t.chain_len;
        } else {
            rc = HV_ERR_CAPACITY;
        }
        nd++;
    }
    if (n_defects) *n_defects = std::min(nd, cap);
    if (rc != HV_OK) return fail(ctx, rc, "capacity exceeded: defects array too small");
    (void)px;
    return HV_OK;
}

// ---- result side (N4) ---------------------------------------------------------------------------------------------------
hv_status hv_export_results(const hv_frame_result *results, int32_t n, double timestamp, double processing_time,
                            uint64_t first_sequence, hv_inspection_record *records, hv_dashboard_stats *stats) {
    if (!results || n < 0 || (!records && !stats)) return HV_ERR_INVALID_ARGUMENT;
    for (int f = 0; f < n; f++) {
        const hv_frame_result &r = results[f];
        if (records) {
            hv_inspection_record &q = records[f];
            q.sequence = first_sequence + (uint64_t)f;
            q.timestamp = timestamp;
            q.processing_time = processing_time;
            q.success = r.status == HV_OK ? 1u : 0u;
            q.has_defects = r.n_defects > 0 ? 1u : 0u;  // base_inspector.py:40-42
            q.defect_count = r.n_defects;
            q.defects_offset = r.defects_offset;
        }
        if (stats) {  // dashboard.py:483-500, one image at a time, the reference's expression order
            stats->total_images += 1;
            stats->total_defects += r.n_defects;
            if (stats->avg_processing_time_ms == 0)
                stats->avg_processing_time_ms = processing_time * 1000;
            else
                stats->avg_processing_time_ms = 0.9 * stats->avg_processing_time_ms + 0.1 * processing_time * 1000;
            if (stats->total_images > 0)
                stats->defect_rate = (double)stats->total_defects / (double)stats->total_images * 100;
        }
    }
    return HV_OK;
}

hv_status hv_draw_overlays(hv_ctx *ctx, uint8_t *img, int32_t h, int32_t w, const hv_overlay *items, int32_t n) {
    if (!ctx || !img || h <= 0 || w <= 0 || n < 0 || (n > 0 && !items)) return fail(ctx, HV_ERR_INVALID_ARGUMENT, "bad argument");
    if ((long long)h * w >= 2147483647LL) return fail(ctx, HV_ERR_INVALID_ARGUMENT, "frame too large");
    for (int i = 0; i < n; i++)
        if (items[i].kind < HV_OVERLAY_CROSS || items[i].kind > HV_OVERLAY_MARKER) return fail(ctx, HV_ERR_INVALID_ARGUMENT, "unknown overlay kind");
    if (n == 0) return HV_OK;
    HV_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[0].stream;
    const size_t px = (size_t)h * w;
    HV_TRY_CUDA(ctx, ctx->u_a.reserve(px * 3));
    HV_TRY_CUDA(ctx, ctx->u_b.reserve(sizeof(hv_overlay) * (size_t)n));
    if (px > ctx->u_owner.cap) {  // all zero between calls: the kernels release what they claim
        HV_TRY_CUDA(ctx, ctx->u_owner.reserve(px));
        HV_TRY_CUDA(ctx, cudaMemsetAsync(ctx->u_owner.p, 0, px * sizeof(uint32_t), st));
    }
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(ctx->u_a.p, img, px * 3, cudaMemcpyHostToDevice, st));
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(ctx->u_b.p, items, sizeof(hv_overlay) * (size_t)n, cudaMemcpyHostToDevice, st));
    HV_TRY_CUDA(ctx, launch_overlays(reinterpret_cast<const hv_overlay *>(ctx->u_b.p), n, h, w, ctx->u_a.p, ctx->u_owner.p, st));
    ctx->launches += 3;
    HV_TRY_CUDA(ctx, cudaMemcpyAsync(img, ctx->u_a.p, px * 3, cudaMemcpyDeviceToHost, st));
    HV_TRY_CUDA(ctx, cudaStreamSynchronize(st));
    return HV_OK;
}

// ---- line statistics ----------------------------------------------------------------------------------------------
hv_status hv_stats_get(hv_ctx *ctx, hv_line_stats *out) {
    if (!ctx || !out) return HV_ERR_INVALID_ARGUMENT;
    HV_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    {
        hv_status rf = flush_deferred(ctx);  // every enqueued batch counts
        if (rf != HV_OK) return rf;
    }
    HV_TRY_CUDA(ctx, cudaDeviceSynchronize());
    HV_TRY_CUDA(ctx, cudaMemcpy(out, ctx->d_stats, sizeof(*out), cudaMemcpyDeviceToHost));
    return HV_OK;
}

hv_status hv_stats_reset(hv_ctx *ctx) {
    if (!ctx) return HV_ERR_INVALID_ARGUMENT;
    HV_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    {
        hv_status rf = flush_deferred(ctx);  // every enqueued batch counts
        if (rf != HV_OK) return rf;
    }
    HV_TRY_CUDA(ctx, cudaDeviceSynchronize());
    HV_TRY_CUDA(ctx, cudaMemset(ctx->d_stats, 0, sizeof(hv_line_stats)));
    return HV_OK;
}

uint64_t *hv_stats_device_ptr(hv_ctx *ctx) { return ctx ? reinterpret_cast<uint64_t *>(ctx->d_stats) : nullptr; }

uint64_t hv_launch_count(const hv_ctx *ctx) { return ctx ? ctx->launches : 0; }

hv_status hv_profile_enable(hv_ctx *ctx, uint32_t kernel_mask) {
    if (!ctx) return HV_ERR_INVALID_ARGUMENT;
    ctx->prof_mask = kernel_mask;
    return HV_OK;
}

hv_status hv_profile_get(hv_ctx *ctx, float total_ms[HV_K_COUNT], uint32_t counts[HV_K_COUNT]) {
    if (!ctx || !total_ms) return HV_ERR_INVALID_ARGUMENT;
    for (int k = 0; k < HV_K_COUNT; k++) {
        total_ms[k] = 0.f;
        if (counts) counts[k] = 0;
    }
    for (auto &r : ctx->prof_recs) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            total_ms[r.k] += ms;
            if (counts) counts[r.k]++;
        }
        ctx->prof_pool.push_back(r.a);
        ctx->prof_pool.push_back(r.b);
    }
    cudaGetLastError();
    ctx->prof_recs.clear();
    return HV_OK;
}

hv_status hv_debug_phase_times(hv_ctx *ctx, uint64_t out_ns[256]) {
    if (!ctx || !out_ns) return HV_ERR_INVALID_ARGUMENT;
    HV_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    HV_TRY_CUDA(ctx, cudaDeviceSynchronize());
    HV_TRY_CUDA(ctx, cudaMemcpy(out_ns, ctx->d_phase_ns, 256 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return HV_OK;
}

const char *hv_kernel_name(int32_t k) {
    static const char *names[HV_K_COUNT] = {"gray3",       "preprocess_mask", "morph",     "ccl_merge", "ccl_flatten",
                                            "ccl_scan",    "ccl_label",       "score",     "ccl_frame_fused"};
    return (k >= 0 && k < HV_K_COUNT) ? names[k] : "?";
}

}  // extern "C"
