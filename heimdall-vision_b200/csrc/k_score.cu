// k_score.cu -- K6: per-blob filters, contrast probe, confidence and the reject decision.
//
// Restates rust/heimdall-core/src/detection.rs:247-311:
//   keep iff min_size <= area <= max_size (f64 compare, both ends inclusive)
//   centre = (floor(sum_y/n), floor(sum_x/n))
//   5x5 window clamped to the image around the centre: pixels with mask == 255 (ANY component) are foreground,
//   the rest background; means of GRAY (not the blurred image); an empty side counts as 127.0
//   shape = 1 - area/rect ; intensity = min(|bg-fg|/30, 1) ; confidence = intensity*0.7 + shape*0.3 ; keep iff >= 0.3
// All f64 operations use the round-to-nearest intrinsics so nvcc cannot contract a*b+c into an FMA (rustc never does;
// SURVEY.md KAT2 = 0.7999999999999999 depends on it).
// Defects are emitted in label order (= the reference's discovery order) with an order-preserving block scan.
// reject := n_defects > 0 (heimdall/inspection/base_inspector.py:40-42).
#include "hv_common.cuh"
#include "score_device.cuh"

namespace hv {

namespace {

// One CTA per chunk of 256 blobs (grid.x) and frame (grid.y).  The defect list keeps the label order, so a chunk needs the
// number of defects of all earlier chunks of its frame: every CTA publishes its count in score_state[chunk] (bit 31 =
// "there") and sums the counts before it (the CTAs of a frame are dispatched in chunk order, so the ones it spins on
// are running or done).  A completion counter behind the counts tells the last CTA of a frame to finish its look-back
// that nobody reads the counts any more: it zeroes them for the next launch.  With one CTA per frame a 12 MP frame
// with 14 k blobs spent 0.5 ms here (56 serial rounds of 50 dependent loads and four f64 divisions).
__global__ void __launch_bounds__(256) k_score(BatchView b, ScoreParams p) {
    __shared__ uint32_t s_warp[8], s_look[8];
    __shared__ uint32_t s_last;
    __shared__ unsigned long long s_area;
    __shared__ uint32_t s_hist[HV_STATS_AREA_BINS];
    const int f = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t chunk = blockIdx.x;
    if (b.frame_select && !(b.frame_select[f] & 1u)) return;
    const uint32_t ncomp = b.ncomp[f];
    const uint32_t nb = min(ncomp, (uint32_t)b.blob_cap);
    const uint32_t nchunks = max((nb + 255u) / 256u, 1u);
    if (chunk >= nchunks) return;
    uint32_t *state = b.score_state + (size_t)f * (b.score_chunks + 1);
    const hv_blob *blobs = b.blobs + (size_t)f * b.blob_cap;
    hv_defect *out = b.defects + (size_t)f * b.defect_cap;
    if (tid == 0) s_area = 0;
    if (tid < HV_STATS_AREA_BINS) s_hist[tid] = 0;

    const uint32_t k = chunk * 256u + tid;
    Scored sc;
    sc.keep = false;
    if (k < nb) sc = score_blob(b, p, f, k, blobs[k]);
    const uint32_t ballot = __ballot_sync(0xffffffffu, sc.keep);
    if (lane == 0) s_warp[wid] = __popc(ballot);
    __syncthreads();
    uint32_t pos = __popc(ballot & ((1u << lane) - 1u));
    uint32_t cnt = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        if (w < wid) pos += s_warp[w];
        cnt += s_warp[w];
    }
    if (tid == 0) atomicExch(state + chunk, cnt | 0x80000000u);
    // look-back over the earlier chunks of the frame
    uint32_t before = 0;
    for (uint32_t j = tid; j < chunk; j += 256) {
        uint32_t v;
        do {
            v = *reinterpret_cast<volatile uint32_t *>(state + j);
        } while (!(v >> 31));
        before += v & 0x7fffffffu;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    if (lane == 0) s_look[wid] = before;
    __syncthreads();
    uint32_t base = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) base += s_look[w];
    if (sc.keep) {
        if (base + pos < (uint32_t)b.defect_cap) out[base + pos] = sc.d;
        atomicAdd(&s_area, (unsigned long long)sc.d.size);
        atomicAdd(&s_hist[area_bin((uint32_t)sc.d.size)], 1u);
    }
    if (tid == 0) s_last = atomicAdd(state + b.score_chunks, 1u) == nchunks - 1u ? 1u : 0u;  // every look-back read is behind us
    __syncthreads();
    if (s_last) {  // every chunk of the frame has finished reading the counts
        for (uint32_t j = tid; j < nchunks; j += 256) state[j] = 0u;
        if (tid == 0) state[b.score_chunks] = 0u;
    }
    unsigned long long *st = reinterpret_cast<unsigned long long *>(b.stats);
    if (tid == 0) {
        if (s_area) atomicAdd(st + 4, s_area);  // total_defect_area
        if (chunk == nchunks - 1u) {
            const uint32_t nd = base + cnt;
            const bool overflow = ncomp > (uint32_t)b.blob_cap || nd > (uint32_t)b.defect_cap;
            hv_frame_result r;
            r.n_components = ncomp;
            r.n_defects = min(nd, (uint32_t)b.defect_cap);
            r.defects_offset = 0;
            r.rejected = nd > 0 ? 1u : 0u;
            r.fg_pixels = b.fgcount[f];
            r.status = overflow ? HV_ERR_CAPACITY : HV_OK;
            b.results[f] = r;
            atomicAdd(st + 0, 1ull);                             // frames_inspected
            atomicAdd(st + 1, (unsigned long long)r.rejected);  // frames_rejected
            atomicAdd(st + 2, (unsigned long long)nd);          // total_defects
            atomicAdd(st + 3, (unsigned long long)ncomp);       // total_components
            atomicAdd(st + 5, (unsigned long long)r.fg_pixels); // total_fg_pixels
            if (overflow) atomicAdd(st + 6 + HV_STATS_AREA_BINS, 1ull);
        }
    }
    if (tid < HV_STATS_AREA_BINS && s_hist[tid]) atomicAdd(st + 6 + tid, (unsigned long long)s_hist[tid]);
}

// (cy, cx) of every blob with area >= min_area, label order (processing.rs:355-366).
__global__ void __launch_bounds__(256) k_collect_centers(BatchView b, uint32_t min_area, hv_center *out,
                                                         uint32_t *count, int cap) {
    // single CTA, frame 0; ordered compaction as in k_score
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t nb = min(b.ncomp[0], (uint32_t)b.blob_cap);
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (uint32_t k0 = 0; k0 < nb; k0 += 256) {
        const uint32_t k = k0 + tid;
        bool keep = false;
        hv_blob q;
        if (k < nb) {
            q = b.blobs[k];
            keep = q.area >= min_area;
        }
        const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[wid] = __popc(ballot);
        __syncthreads();
        uint32_t pos = s_base + __popc(ballot & ((1u << lane) - 1u));
        uint32_t chunk_total = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            if (w < wid) pos += s_warp[w];
            chunk_total += s_warp[w];
        }
        if (keep && pos < (uint32_t)cap) {
            hv_center c;
            c.y = (int32_t)(q.sum_y / q.area);
            c.x = (int32_t)(q.sum_x / q.area);
            c.confidence = 0.75;
            out[pos] = c;
        }
        __syncthreads();
        if (tid == 0) s_base += chunk_total;
        __syncthreads();
    }
    if (tid == 0) *count = s_base;
}

// find_contours records (detection.rs:90-113): min_area <= area <= max_area, label order.
__global__ void __launch_bounds__(256) k_collect_contours(BatchView b, double min_area, double max_area,
                                                          hv_contour *out, uint32_t *count, int cap) {
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t nb = min(b.ncomp[0], (uint32_t)b.blob_cap);
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (uint32_t k0 = 0; k0 < nb; k0 += 256) {
        const uint32_t k = k0 + tid;
        bool keep = false;
        hv_blob q;
        if (k < nb) {
            q = b.blobs[k];
            const double area = (double)q.area;
            keep = q.area > 0 && area >= min_area && area <= max_area;
        }
        const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[wid] = __popc(ballot);
        __syncthreads();
        uint32_t pos = s_base + __popc(ballot & ((1u << lane) - 1u));
        uint32_t chunk_total = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            if (w < wid) pos += s_warp[w];
            chunk_total += s_warp[w];
        }
        if (keep && pos < (uint32_t)cap) {
            hv_contour c;
            c.y = (int32_t)(q.sum_y / q.area);
            c.x = (int32_t)(q.sum_x / q.area);
            c.area = (double)q.area;
            c.pixel_count = q.area;
            c.label = k + 1;
            c.reserved = 0;
            out[pos] = c;
        }
        __syncthreads();
        if (tid == 0) s_base += chunk_total;
        __syncthreads();
    }
    if (tid == 0) *count = s_base;
}

}  // namespace

cudaError_t launch_score(const BatchView &b, const ScoreParams &p, cudaStream_t s) {
    k_score<<<dim3(b.score_chunks > 0 ? b.score_chunks : 1, b.n), 256, 0, s>>>(b, p);
    return cudaGetLastError();
}

cudaError_t launch_collect_centers(const BatchView &b, uint32_t min_area, hv_center *d_centers, uint32_t *d_count,
                                   int cap, cudaStream_t s) {
    k_collect_centers<<<1, 256, 0, s>>>(b, min_area, d_centers, d_count, cap);
    return cudaGetLastError();
}

cudaError_t launch_collect_contours(const BatchView &b, double min_area, double max_area, hv_contour *d_out,
                                    uint32_t *d_count, int cap, cudaStream_t s) {
    k_collect_contours<<<1, 256, 0, s>>>(b, min_area, max_area, d_out, d_count, cap);
    return cudaGetLastError();
}

}  // namespace hv
