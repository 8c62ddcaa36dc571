// hv_frameset.cu -- N2: multi-camera FrameSet batching in front of the frame feed (host side, C++).
//
// Reference: a `FrameSet` is the group of frames several cameras acquired for one trigger
// (rust/heimdall-gige/src/frame.rs:127-185: frames by camera id, has_all_cameras, one global frame_id), produced by
// GigESystem::acquire_frames (rust/heimdall-gige/src/lib.rs:529-648) for at most 4 Mono8 1920x1080 cameras
// (lib.rs:206-234) in one of three sync modes (rust/heimdall-gige/src/sync.rs:18-27): Freerun, Software, Hardware.
// The reference builds a set by awaiting every camera; a camera that keeps failing for 100 ms fails the whole set
// (lib.rs:590-606).  This batcher is the push-style equivalent for frames that arrive camera by camera:
//   * Software / Hardware: frames are matched by their trigger number (frame_id);  Freerun: the k-th frame of every
//     camera forms set k (there is no trigger to match on);
//   * a set is complete when all cameras delivered (has_all_cameras); `sets_per_batch` complete sets, in ascending
//     set order, form one detector batch (set-major, camera-minor) that goes through hv_submit_frames;
//   * frames are copied into page-locked slabs on arrival, so the camera buffer can be re-queued at once and the
//     H2D copy needs no further staging;
//   * at most `max_pending_sets` incomplete sets are kept; when a newer set arrives beyond that, the oldest incomplete
//     one is dropped (the counterpart of the reference's failed acquire) and accounted in the statistics, as are
//     duplicate deliveries and the largest timestamp skew inside a set.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "../../include/heimdall_cuda.h"

struct hv_frameset {
    hv_ctx *ctx = nullptr;
    hv_frameset_config cfg{};
    size_t frame_bytes = 0;  // fixed by the first frame pushed
    uint32_t width = 0, height = 0;
    int32_t pixel_format = -1;
    struct Set {
        uint64_t id = 0;
        uint8_t *slab = nullptr;  // n_cameras * frame_bytes, page-locked
        std::vector<uint8_t> have;
        std::vector<hv_camera_frame> meta;
        int count = 0;
        uint64_t t_min = ~0ull, t_max = 0;
    };
    std::map<uint64_t, Set> open;              // incomplete sets by id
    std::deque<Set> ready;                     // complete sets awaiting a full batch, ascending id
    std::vector<uint64_t> next_index;          // Freerun: next set index per camera
    std::vector<uint8_t *> free_slabs;
    struct InFlight {
        int64_t ticket;
        std::vector<Set> sets;
    };
    std::vector<InFlight> inflight;
    hv_frameset_stats st{};
    uint64_t dropped_below = 0;  // sets with id < this were dropped or already batched: late frames for them are discarded
    std::string err;
};

namespace {

uint8_t *take_slab(hv_frameset *fs) {
    if (!fs->free_slabs.empty()) {
        uint8_t *p = fs->free_slabs.back();
        fs->free_slabs.pop_back();
        return p;
    }
    void *p = nullptr;
    if (!fs->ctx) return static_cast<uint8_t *>(std::malloc(fs->frame_bytes * fs->cfg.n_cameras));  // dry mode
    if (cudaHostAlloc(&p, fs->frame_bytes * fs->cfg.n_cameras, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return static_cast<uint8_t *>(p);
}

void free_slab(hv_frameset *fs, uint8_t *p) {
    if (!p) return;
    if (fs->ctx)
        cudaFreeHost(p);
    else
        std::free(p);
}

hv_status fs_fail(hv_frameset *fs, hv_status s, const char *msg) {
    fs->err = msg;
    return s;
}

// Cut one batch off the complete sets (ascending id order) and hand it to the detector.  The sets leave `ready` only
// once hv_submit_frames has accepted them: when every slot of the context is still in flight (HV_ERR_CAPACITY) the
// batch stays queued, the status is returned to the caller -- who collects an earlier ticket with hv_frameset_wait and
// calls hv_frameset_flush, or simply pushes on: the next completed set retries -- and nothing is lost silently.  The
// queue of complete sets is bounded (max_pending_sets beyond a full batch): past that the oldest complete set is
// dropped and accounted like an incomplete one.
hv_status submit_ready(hv_frameset *fs, const hv_params *params, int64_t *ticket) {
    const int spb = fs->cfg.sets_per_batch;
    if ((int)fs->ready.size() < spb) return HV_OK;
    std::vector<hv_camera_frame> batch;
    for (int k = 0; k < spb; k++)
        for (auto &m : fs->ready[k].meta) batch.push_back(m);
    int64_t t = -(int64_t)(fs->st.batches_submitted + 1);
    hv_status rs = fs->ctx ? hv_submit_frames(fs->ctx, batch.data(), (int32_t)batch.size(), params, &t) : HV_OK;
    if (rs != HV_OK) {
        fs->err = hv_last_error(fs->ctx);
        while ((int)fs->ready.size() > spb + fs->cfg.max_pending_sets) {  // back-pressure has a bound
            fs->st.sets_dropped++;
            fs->st.frames_dropped += (uint64_t)fs->cfg.n_cameras;
            fs->dropped_below = std::max(fs->dropped_below, fs->ready.front().id + 1);
            fs->free_slabs.push_back(fs->ready.front().slab);
            fs->ready.pop_front();
        }
        return rs;
    }
    hv_frameset::InFlight fl;
    for (int k = 0; k < spb; k++) {
        fl.sets.push_back(std::move(fs->ready.front()));
        fs->ready.pop_front();
    }
    fs->dropped_below = std::max(fs->dropped_below, fl.sets.back().id + 1);
    // sets older than the batch that are still incomplete can never be delivered in order any more
    while (!fs->open.empty() && fs->open.begin()->first < fs->dropped_below) {
        auto old = fs->open.begin();
        fs->st.sets_dropped++;
        fs->st.frames_dropped += (uint64_t)old->second.count;
        fs->free_slabs.push_back(old->second.slab);
        fs->open.erase(old);
    }
    fl.ticket = t;
    fs->inflight.push_back(std::move(fl));
    fs->st.batches_submitted++;
    *ticket = t;
    return HV_OK;
}

}  // namespace

extern "C" {

hv_status hv_frameset_create(hv_ctx *ctx, const hv_frameset_config *cfg, hv_frameset **out) {
    // ctx == NULL: dry mode (host logic only: sets and batches are formed and accounted, nothing is submitted;
    // tickets are -1, -2, ...); used by the CPU tests of the batching rules
    if (!cfg || !out) return HV_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (cfg->n_cameras < 1 || cfg->n_cameras > 16 || cfg->sets_per_batch < 1 || cfg->sets_per_batch > 4096 ||
        cfg->sync_mode < HV_SYNC_FREERUN || cfg->sync_mode > HV_SYNC_HARDWARE)
        return HV_ERR_INVALID_ARGUMENT;
    hv_frameset *fs = new (std::nothrow) hv_frameset();
    if (!fs) return HV_ERR_INVALID_ARGUMENT;
    fs->ctx = ctx;
    fs->cfg = *cfg;
    if (fs->cfg.max_pending_sets <= 0) fs->cfg.max_pending_sets = 8;
    fs->next_index.assign(cfg->n_cameras, 0);
    *out = fs;
    return HV_OK;
}

void hv_frameset_destroy(hv_frameset *fs) {
    if (!fs) return;
    auto drop = [&](hv_frameset::Set &s) { free_slab(fs, s.slab); };
    for (auto &kv : fs->open) drop(kv.second);
    for (auto &s : fs->ready) drop(s);
    for (auto &f : fs->inflight)
        for (auto &s : f.sets) drop(s);
    for (auto *p : fs->free_slabs) free_slab(fs, p);
    delete fs;
}

const char *hv_frameset_last_error(const hv_frameset *fs) { return fs ? fs->err.c_str() : ""; }

hv_status hv_frameset_push(hv_frameset *fs, const hv_camera_frame *fr, const hv_params *params, int64_t *ticket) {
    if (!fs || !fr || !ticket) return HV_ERR_INVALID_ARGUMENT;
    *ticket = 0;
    if (!fr->data || fr->camera >= (uint32_t)fs->cfg.n_cameras) return fs_fail(fs, HV_ERR_INVALID_ARGUMENT, "bad camera index or empty frame");
    if (fs->pixel_format < 0) {  // the first frame fixes the geometry of the line
        const int32_t ch = hv_frame_channels(fr->pixel_format);
        if (ch != 1 && ch != 3) return fs_fail(fs, HV_ERR_INVALID_DIMENSIONS, "Invalid image dimensions: expected 3D array");
        const size_t bpp = (fr->pixel_format == HV_PIX_RGB8 || fr->pixel_format == HV_PIX_BGR8) ? 3
                           : (fr->pixel_format == HV_PIX_YUV422 || fr->pixel_format == HV_PIX_YUV422_PACKED) ? 2 : 1;
        fs->frame_bytes = (size_t)fr->width * fr->height * bpp;
        fs->width = fr->width, fs->height = fr->height, fs->pixel_format = fr->pixel_format;
        if (!fs->frame_bytes) return fs_fail(fs, HV_ERR_INVALID_ARGUMENT, "empty frame");
    }
    if (fr->width != fs->width || fr->height != fs->height || fr->pixel_format != fs->pixel_format || fr->size < fs->frame_bytes)
        return fs_fail(fs, HV_ERR_INVALID_ARGUMENT, "all cameras of a frame set must deliver the same geometry and pixel format");
    fs->st.frames_pushed++;
    const uint64_t id = fs->cfg.sync_mode == HV_SYNC_FREERUN ? fs->next_index[fr->camera]++ : fr->frame_id;
    if (id < fs->dropped_below) {  // its set is gone already
        fs->st.frames_dropped++;
        return HV_OK;
    }
    auto it = fs->open.find(id);
    if (it == fs->open.end()) {
        // a new set: make room first (oldest incomplete set goes)
        while ((int)fs->open.size() >= fs->cfg.max_pending_sets) {
            auto old = fs->open.begin();
            if (old->first > id) {  // the newcomer itself is the oldest: it is the one that is too late
                fs->st.frames_dropped++;
                return HV_OK;
            }
            fs->st.sets_dropped++;
            fs->st.frames_dropped += (uint64_t)old->second.count;
            fs->dropped_below = std::max(fs->dropped_below, old->first + 1);
            fs->free_slabs.push_back(old->second.slab);
            fs->open.erase(old);
        }
        hv_frameset::Set s;
        s.id = id;
        s.slab = take_slab(fs);
        if (!s.slab) return fs_fail(fs, HV_ERR_CUDA, "cudaHostAlloc failed for a frame-set slab");
        s.have.assign(fs->cfg.n_cameras, 0);
        s.meta.resize(fs->cfg.n_cameras);
        it = fs->open.emplace(id, std::move(s)).first;
    }
    hv_frameset::Set &s = it->second;
    if (s.have[fr->camera]) {
        fs->st.duplicates++;
        return HV_OK;
    }
    // zero-copy: a frame that already lies in page-locked memory is referenced where it is
    bool referenced = false;
    if (fs->ctx && (fs->cfg.flags & HV_FRAMESET_ZERO_COPY)) {
        cudaPointerAttributes at{};
        referenced = cudaPointerGetAttributes(&at, fr->data) == cudaSuccess && at.type == cudaMemoryTypeHost;
        cudaGetLastError();
    }
    if (!referenced) std::memcpy(s.slab + fs->frame_bytes * fr->camera, fr->data, fs->frame_bytes);
    s.have[fr->camera] = 1;
    s.meta[fr->camera] = *fr;
    if (!referenced) s.meta[fr->camera].data = s.slab + fs->frame_bytes * fr->camera;
    s.meta[fr->camera].size = fs->frame_bytes;
    s.meta[fr->camera].frame_id = id;
    s.t_min = std::min(s.t_min, fr->timestamp_ns);
    s.t_max = std::max(s.t_max, fr->timestamp_ns);
    if (++s.count < fs->cfg.n_cameras) return HV_OK;
    // complete (has_all_cameras): sets complete in any order, batches are cut in ascending id order
    fs->st.sets_completed++;
    fs->st.max_skew_ns = std::max(fs->st.max_skew_ns, s.t_max - s.t_min);
    hv_frameset::Set done = std::move(s);
    fs->open.erase(it);
    auto pos = std::lower_bound(fs->ready.begin(), fs->ready.end(), done.id,
                                [](const hv_frameset::Set &a, uint64_t v) { return a.id < v; });
    fs->ready.insert(pos, std::move(done));
    return submit_ready(fs, params, ticket);
}

hv_status hv_frameset_flush(hv_frameset *fs, const hv_params *params, int64_t *ticket) {
    if (!fs || !ticket) return HV_ERR_INVALID_ARGUMENT;
    *ticket = 0;
    return submit_ready(fs, params, ticket);
}

hv_status hv_frameset_batch_ids(hv_frameset *fs, int64_t ticket, uint64_t *set_ids, int32_t cap, int32_t *n_sets) {
    if (!fs) return HV_ERR_INVALID_ARGUMENT;
    for (auto &f : fs->inflight)
        if (f.ticket == ticket) {
            if (n_sets) *n_sets = (int32_t)f.sets.size();
            for (int k = 0; k < (int)f.sets.size() && k < cap; k++)
                if (set_ids) set_ids[k] = f.sets[k].id;
            return HV_OK;
        }
    return fs_fail(fs, HV_ERR_BAD_TICKET, "unknown or already consumed ticket");
}

hv_status hv_frameset_wait(hv_frameset *fs, int64_t ticket, hv_frame_result *results, hv_defect *defects,
                           size_t defects_cap, size_t *n_defects_total) {
    if (!fs) return HV_ERR_INVALID_ARGUMENT;
    for (size_t i = 0; i < fs->inflight.size(); i++)
        if (fs->inflight[i].ticket == ticket) {
            hv_status rs = fs->ctx ? hv_wait(fs->ctx, ticket, results, defects, defects_cap, n_defects_total)
                                   : HV_ERR_NO_DEVICE;
            // a failed wait says nothing about the H2D copy out of the slabs: make sure the device is through with
            // them before they are handed out again
            if (fs->ctx && rs != HV_OK && rs != HV_ERR_CAPACITY && cudaDeviceSynchronize() != cudaSuccess) cudaGetLastError();
            for (auto &q : fs->inflight[i].sets) fs->free_slabs.push_back(q.slab);
            fs->inflight.erase(fs->inflight.begin() + i);
            if (rs != HV_OK) fs->err = fs->ctx ? hv_last_error(fs->ctx) : "dry mode: nothing was submitted";
            return rs;
        }
    return fs_fail(fs, HV_ERR_BAD_TICKET, "unknown or already consumed ticket");
}

hv_status hv_frameset_get_stats(const hv_frameset *fs, hv_frameset_stats *out) {
    if (!fs || !out) return HV_ERR_INVALID_ARGUMENT;
    *out = fs->st;
    out->sets_pending = (uint64_t)(fs->open.size() + fs->ready.size());
    return HV_OK;
}

}  // extern "C"
