// k_ccl_frame.cu -- fused per-frame CCL + blob statistics + scoring for sparse masks: ONE kernel, one CTA per frame.
//
// Same semantics as the global-memory path (k_ccl.cu K2..K5 + k_score.cu K6; reference:
// rust/heimdall-core/src/detection.rs:215-311), different data placement.  The global path keeps its union-find
// parents inside the 4 B/px label plane, so every find/union step is a dependent DRAM round trip and the five
// kernels are latency-bound.  Inspection masks are sparse (a few thousand foreground pixels per 1.3 MP frame), so
// here the whole working set of a frame lives in shared memory:
//   phase 0  ordered compaction of the frame's non-zero bitmask words              -> widx / wbits   (raster order)
//   phase 1  word-runs become nodes, numbered in raster order (exclusive scan)      -> woff, node_px, node_len
//   phase 2  unions by minimum node id in shared memory (node id order == raster order of the runs' first pixels,
//            so every root is its component's raster-first run)
//   phase 3  flatten, rank the roots with a block scan                              -> canonical label = rank + 1
//   phase 4  per-run statistics into a shared-memory blob table, final labels to the label plane
//   phase 5  blob table -> global, scoring (shared with K6), ordered defect compaction, per-frame result, line stats
// A frame that does not fit (more than kCapW non-zero words, kCapN runs or kCapB components, or a side > 8192) is left
// untouched and flagged in frame_flags; the host then runs the global path on exactly those frames.
//
// The kernel is latency-bound by construction (25 CTAs for the headline batch), and every instruction is executed
// once per CTA, i.e. from a cold instruction cache: the first version (96 KB of SASS from 8-fold unrolled phase
// bodies, work split unevenly over the warps) spent 55 us per batch, 70 % of the warp samples waiting at barriers for
// one straggling warp.  This version keeps the code small (runtime loops, one shared scan routine), gives every
// thread an equal share of every phase, and has no dependent global-memory round trip that can be avoided.
#include <cstddef>

#include "hv_common.cuh"
#include "score_device.cuh"

namespace hv {

namespace {

// Two builds of the kernel.  Big: 512 threads, 4096 non-zero words / 8192 runs / 2048 components per frame, 170 KB of
// shared memory -- a CTA needs an SM to itself.  Small: 256 threads (72 registers), 2048 / 4096 / 512, 68 KB -- fits next to
// four resident CTAs of K1 (155 KB), so its CTAs become resident the moment the kernel is launched, release the next K1
// at once (see the top of the kernel) and simply wait there for their own K1 to finish: the chain of K1 launches has no
// gap and no SM is set aside.  hv_api.cu picks the build per batch (small first; a frame that does not fit is flagged).
template <int FT, int CW, int CN, int CB, int CE>
struct CclCfg {
    static constexpr int kFT = FT;         // threads per CTA
    static constexpr int kNW = FT / 32;    // warps per CTA
    static constexpr int kCapW = CW;       // non-zero words per frame
    static constexpr int kCapN = CN;       // word-runs (nodes) per frame
    static constexpr int kCapB = CB;       // components per frame
    static constexpr int kCapE = CE;       // residual union edges per frame
    static constexpr int kMaxEPT = CW / FT;  // compacted words per thread (blocked partition)
};
using CclBig = CclCfg<512, 4096, 8192, 2048, 4096>;
using CclSmall = CclCfg<256, 2048, 4096, 512, 1024>;
// Tiny: 128 threads, 768 / 1024 / 256, 22 KB -- for frames that are sparse after morphology (300-460 non-zero words on the
// headline frames with open + close 3x3).  It needs less than one K1 CTA does of everything (shared memory, registers,
// threads), so it can take the place of any single K1 CTA that retires.
using CclTiny = CclCfg<128, 768, 1024, 256, 512>;

template <typename C>
struct FrameSmem {
    union {  // the compacted words are dead after phase 2, the blob table is born in phase 3
        struct {
            uint32_t widx[C::kCapW];   // word index (y * ww + wx) of the compacted non-zero words, raster order
            uint32_t wbits[C::kCapW];
        };
        struct {
            uint32_t b_area[C::kCapB], b_sy[C::kCapB], b_sx[C::kCapB], b_ymin[C::kCapB], b_ymax[C::kCapB], b_xmin[C::kCapB],
                b_xmax[C::kCapB];
        };
    };
    uint32_t parent[C::kCapN];    // union-find parents (node ids)
    uint32_t node_px[C::kCapN];   // pixel index (y * w + x) of the first pixel of the run
    uint16_t woff[C::kCapW + 8];  // first node of every compacted word
    uint32_t edges[C::kCapE];     // residual union edges (u << 16 | v), see phase 2
    uint32_t n_edges;
    uint32_t cursor;              // phase 2: next word to hand out
    uint16_t rnk[C::kCapN];       // rank of the component among the roots (valid at root nodes)
    uint8_t node_len[C::kCapN];
    uint32_t warp_tmp[32];
    uint32_t warp_fg[32];
    uint32_t area_sum;
    uint32_t hist[HV_STATS_AREA_BINS];
};
static_assert(sizeof(FrameSmem<CclBig>) <= 227 * 1024, "FrameSmem must fit the 227 KB per-CTA shared memory of sm_100");
static_assert(sizeof(FrameSmem<CclSmall>) <= 71 * 1024, "the small build must fit next to four K1 CTAs");
static_assert(sizeof(FrameSmem<CclTiny>) <= 24 * 1024, "the tiny build must fit into the room one K1 CTA leaves");
static_assert(CclBig::kCapN <= 65536 && CclBig::kCapE <= 65536, "16-bit node ids in rnk / edges");

// exclusive scan of one value per thread across the CTA; returns the exclusive prefix, *total_out gets the block sum.
// Contains two __syncthreads().  One copy in the binary (see the header: code size is latency here).
template <int kNW>
__device__ __noinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *tmp, uint32_t *total_out) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();  // protect tmp from the previous use
    if (lane == 31) tmp[wid] = incl;
    __syncthreads();
    uint32_t w = lane < kNW ? tmp[lane] : 0u;
    uint32_t wi = w;
#pragma unroll
    for (int o = 1; o < kNW; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
    }
    *total_out = __shfl_sync(0xffffffffu, wi, kNW - 1);
    return __shfl_sync(0xffffffffu, wi - w, wid) + incl - v;
}

__device__ __forceinline__ uint32_t s_find(volatile uint32_t *parent, uint32_t a) {
    uint32_t p = parent[a];
    while (p != a) {
        a = p;
        p = parent[a];
    }
    return a;
}

// Lock-free union by minimum id (atomicMin keeps parents monotonically decreasing under concurrent unions).
__device__ __forceinline__ void s_union(uint32_t *parent, uint32_t a, uint32_t b) {
    while (true) {
        a = s_find(parent, a);
        b = s_find(parent, b);
        if (a == b) return;
        if (a < b) {
            const uint32_t t = a;
            a = b;
            b = t;
        }
        const uint32_t old = atomicMin(parent + a, b);
        if (old == a) return;
        a = old;
    }
}

// pointer jumping: every node shortens its own path (parent <- grandparent until the parent is a root); all nodes
// concurrently, so a chain of length n collapses in O(log n) rounds
template <int kFT>
__device__ __forceinline__ void pointer_jump(uint32_t *parent, uint32_t nn, int tid) {
    volatile uint32_t *vp = parent;
    for (uint32_t v = tid; v < nn; v += kFT) {
        uint32_t p = vp[v];
        while (true) {
            const uint32_t gp = vp[p];
            if (gp == p) break;
            vp[v] = gp;
            p = gp;
        }
    }
}

// index (within its word) of the run that contains bit `bit`, given the run-start mask of the word
__device__ __forceinline__ uint32_t run_index(uint32_t starts, int bit) {
    return __popc(starts & ((2u << bit) - 1u)) - 1u;  // (2u << 31) - 1 wraps to all ones, as intended
}

// Registers: the small build shares an SM with four K1 CTAs (4 x 288 threads x 40 registers = 46080 of 65536), which
// leaves 256 threads x 72 registers.
template <typename C>
__global__ void __launch_bounds__(C::kFT) __maxnreg__(C::kFT == 256 ? 72 : (C::kFT == 128 ? 80 : 128)) k_ccl_frame(BatchView b, ScoreParams sp) {
    constexpr int kFT = C::kFT, kNW = C::kNW, kCapW = C::kCapW, kCapN = C::kCapN, kCapB = C::kCapB, kCapE = C::kCapE;
    constexpr int kMaxEPT = C::kMaxEPT;
    using FrameSmem = hv::FrameSmem<C>;
    // Programmatic dependent launch, both ways.  (1) Let K1 of the next batch start as soon as every CTA of this kernel is
    // resident, which is when K1 of this batch is retiring its last CTAs: the tail of one K1 overlaps the head of the next.
    // That K1 works on the other slot, whose last user is the per-frame kernel of the batch before this one; it checks
    // that kernel's completion counter itself before it touches anything (see k_preprocess_tma), because nothing in the
    // launch order guarantees it any more.  (2) This kernel may have been launched while K1 of its own batch was still
    // running: wait until that grid has completed and its stores are visible.  Without the launch attribute both
    // instructions are no-ops.
    if (b.ccl_done) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        if (b.k1_done) {
            // K1 of this batch publishes a launch counter when its last CTA has stored its last tile.  Waiting for that
            // instead of griddepcontrol.wait matters: K1 was itself launched ahead of the previous batch's per-frame
            // kernel, and a grid does not count as complete before the grids ahead of it in the stream have completed,
            // so the hardware wait made the per-frame kernels of consecutive batches run strictly one after the other
            // (measured: step = duration of this kernel + 5..8 us in every configuration).  What K1 wrote is read
            // through the L2 below (ld.cg): a CTA that has been resident since before K1 finished has no guarantee
            // about its L1.
            if (threadIdx.x == 0) {
                const volatile unsigned int *flag = b.k1_done;
                while ((int)(*flag - b.k1_wait_value) < 0) __nanosleep(200);
                __threadfence();
            }
            __syncthreads();
        } else {
            asm volatile("griddepcontrol.wait;" ::: "memory");
        }
    } else {  // no completion counter: the next K1 may only start once K1 of this batch has completed
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }
#ifdef HV_EXPERIMENTS
    if (b.phase_frame == -12345) {  // experiment (HV_EXP_CCL_NOOP): the pipeline's structure without this kernel's work
        if (threadIdx.x == 0 && b.ccl_done) atomicAdd(b.ccl_done, 1u);
        return;
    }
#define HV_EXP_STOP(k)                                                       \
    if (b.phase_frame == -12345 - (k)) {                                     \
        if (threadIdx.x == 0) {                                              \
            hv_frame_result r0{};                                            \
            b.results[blockIdx.x] = r0;                                      \
            b.frame_flags[blockIdx.x] = 0u;                                  \
            __threadfence();                                                 \
            if (b.ccl_done) atomicAdd(b.ccl_done, 1u);                       \
        }                                                                    \
        return;                                                              \
    }
#else
#define HV_EXP_STOP(k)
#endif
    extern __shared__ __align__(16) uint8_t smem_raw[];
    FrameSmem &S = *reinterpret_cast<FrameSmem *>(smem_raw);
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int H = b.h, W = b.w, WW = b.ww;
    const uint32_t *bits = b.bits + (size_t)f * H * WW;
    // optional per-phase timestamps of one frame (HV_FLAG_PHASE_TIMING): phase_ns[stamp * 16 + warp] = globaltimer at
    // the moment the warp reaches the stamp (before the barrier that follows)
    int stamp_i = 0;
    unsigned long long t_start = 0;
    if (b.phase_ns && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    auto stamp = [&]() {
        if (b.phase_ns && f == b.phase_frame && lane == 0 && stamp_i < 12 && wid < 16) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            b.phase_ns[stamp_i * 16 + wid] = t;
        }
        stamp_i++;
    };
    stamp();  // 0

    // ---- phase 0: ordered compaction of the non-zero words ------------------------------------------------------------
    // K1 left one occupancy byte per (row, 128-px tile), bit k set <=> word 4*tile+k of that row is non-zero, stored as
    // 32-byte records per tile (tile-major, see rowflag_index).  The records of a band of 32 rows are contiguous, so
    // whole bands are staged in shared memory with 16-byte loads (one round trip; the staging area is the still unused
    // parent + node_px arrays), then thread t takes the contiguous rows [t*rpt, (t+1)*rpt) of the chunk and a block scan
    // of the popcounts gives its first output slot.  Frames with more bands than fit are staged in several chunks.
    static_assert(offsetof(FrameSmem, node_px) == offsetof(FrameSmem, parent) + sizeof(uint32_t) * kCapN, "staging area");
    static_assert(kNW <= 16, "phase stamps");
    const int TX = b.tiles_x;
    const int nbands = (H + 31) / 32;
    const int band_bytes = TX * 32;
    const uint8_t *rf = b.rowflags + (size_t)f * b.rf_stride;
    uint8_t *stage = reinterpret_cast<uint8_t *>(S.parent);
    const int bands_per_chunk = min(nbands, (int)(2 * sizeof(uint32_t) * kCapN) / band_bytes);
    if (tid < HV_STATS_AREA_BINS) S.hist[tid] = 0;
    if (tid == 0) S.area_sum = 0, S.n_edges = 0, S.cursor = 0;
    uint32_t nw = 0;
    bool too_big = H > 8192 || W > 8192 || bands_per_chunk < 1;
    for (int b0 = 0; b0 < nbands && !too_big; b0 += bands_per_chunk) {
        const int nb = min(bands_per_chunk, nbands - b0);
        const int bytes = nb * band_bytes;  // multiple of 32
        if (b0) __syncthreads();            // the previous chunk has been consumed
        for (int i = tid * 16; i < bytes; i += kFT * 16)
            *reinterpret_cast<uint4 *>(stage + i) =
                __ldcg(reinterpret_cast<const uint4 *>(rf + (size_t)b0 * band_bytes + i));  // (L2: see the wait above)
        __syncthreads();
        const int rows = min(nb * 32, H - b0 * 32);
        const int rpt = (rows + kFT - 1) / kFT;
        const int r0 = min(tid * rpt, rows), r1 = min(r0 + rpt, rows);
        uint32_t cnt = 0;
        for (int r = r0; r < r1; r++) {
            const uint8_t *rec = stage + (r >> 5) * band_bytes + (r & 31);
            for (int tx = 0; tx < TX; tx++) cnt += __popc(rec[tx * 32] & 0xfu);
        }
        uint32_t tot = 0;
        uint32_t pos = nw + block_exclusive_scan<kNW>(cnt, S.warp_tmp, &tot);
        nw += tot;
        if (nw > (uint32_t)kCapW) {  // block-uniform
            too_big = true;
            break;
        }
        if (cnt) {
            for (int r = r0; r < r1; r++) {
                const uint8_t *rec = stage + (r >> 5) * band_bytes + (r & 31);
                const uint32_t wrow = (uint32_t)(b0 * 32 + r) * WW;
                for (int tx = 0; tx < TX; tx++) {
                    uint32_t q = rec[tx * 32] & 0xfu;
                    while (q) {
                        S.widx[pos++] = wrow + 4 * tx + (__ffs(q) - 1);
                        q &= q - 1;
                    }
                }
            }
        }
    }
    if (too_big) {  // block-uniform
        if (tid == 0) {
            b.frame_flags[f] = 1u;
            __threadfence();  // the flag is read (copy engine, host) by whoever sees the counter below
            if (b.ccl_done) atomicAdd(b.ccl_done, 1u);  // nothing of this frame's slot is touched from here on
        }
        return;
    }
    stamp();  // 1
    __syncthreads();
    HV_EXP_STOP(1)

    // ---- phase 1: fetch the words; word-runs become nodes, numbered in raster order -----------------------------------------
    // blocked partition: thread t owns entries [t*ept, (t+1)*ept), so one scan orders the runs of the whole frame
    const int ept = ((int)nw + kFT - 1) / kFT;  // <= kMaxEPT
    const int e0 = min(tid * ept, (int)nw), e1 = min(e0 + ept, (int)nw);
    uint32_t local = 0, fgpx = 0;
    {
        // the gather itself is striped over the CTA (neighbouring lanes fetch neighbouring compacted words, which are mostly
        // neighbouring words of one row: a handful of L2 requests per warp instead of 32 per load), all loads of a thread in
        // flight together; the blocked partition reads the words back from shared memory
        uint32_t wv[kMaxEPT];
#pragma unroll
        for (int k = 0; k < kMaxEPT; k++) {
            const int e = tid + k * kFT;
            wv[k] = e < (int)nw ? __ldcg(bits + S.widx[e]) : 0u;  // one round trip
        }
#pragma unroll
        for (int k = 0; k < kMaxEPT; k++) {
            const int e = tid + k * kFT;
            if (e < (int)nw) S.wbits[e] = wv[k];
        }
        __syncthreads();
        for (int e = e0; e < e1; e++) {
            const uint32_t wvv = S.wbits[e];
            local += __popc(wvv & ~(wvv << 1));
            fgpx += __popc(wvv);
        }
    }
    uint32_t nn = 0;
    uint32_t off = block_exclusive_scan<kNW>(local, S.warp_tmp, &nn);
    if (nn > (uint32_t)kCapN) {  // block-uniform
        if (tid == 0) {
            b.frame_flags[f] = 1u;
            __threadfence();  // the flag is read (copy engine, host) by whoever sees the counter below
            if (b.ccl_done) atomicAdd(b.ccl_done, 1u);  // nothing of this frame's slot is touched from here on
        }
        return;
    }
    for (int e = e0; e < e1; e++) {
        S.woff[e] = (uint16_t)off;
        const uint32_t i = S.widx[e];
        const uint32_t y = i / WW, wx = i - y * WW;
        uint32_t rest = S.wbits[e];
        while (rest) {
            const int bit = __ffs(rest) - 1;
            const uint32_t shifted = ~(rest >> bit);
            const int len = shifted ? __ffs(shifted) - 1 : 32 - bit;
            rest = (bit + len >= 32) ? 0u : (rest & ~(((1u << len) - 1u) << bit));
            S.node_px[off] = y * W + wx * 32 + bit;
            S.node_len[off] = (uint8_t)len;
            off++;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) fgpx += __shfl_xor_sync(0xffffffffu, fgpx, o);
    if (lane == 0) S.warp_fg[wid] = fgpx;
    stamp();  // 2
    __syncthreads();
    HV_EXP_STOP(2)

    // ---- phase 2: unions --------------------------------------------------------------------------------------------------
    // Hooking all the runs of a long vertical line at once would leave a linked list as long as the line.  So the forest is
    // built in three steps:
    //   (a) every node points at its smallest "previous" neighbour (leftmost overlapping run of the row above, else the
    //       run ending the word to the left); previous neighbours always have smaller ids, so these pointers form trees
    //       whose roots are the trees' minimum ids.  Every other adjacency of the node (further overlapping runs above,
    //       the left neighbour when an upper one was chosen) goes to an edge list;
    //   (b) pointer jumping collapses every chain of length n in O(log n) rounds;
    //   (c) the listed edges are ordinary unions by minimum between the now depth-1 trees, ONE edge per thread: inside
    //       the per-word loops a union (a few hundred cycles of dependent shared-memory accesses and an atomic per
    //       retry) was serialised by the SIMT execution of the nested loops of all 32 lanes -- 8 us for 76 unions.
    // The words are handed out to the warps 32 at a time through a shared-memory cursor: words with many runs cluster (the
    // ragged fringe of a thresholded edge: 116 runs in the first 32 words of a bottle frame, 35 in any other 32), and with
    // a fixed assignment the warp that owned such a cluster took 12 us for this loop where the others took 5.
    while (true) {
        uint32_t e = 0;
        if (lane == 0) e = atomicAdd(&S.cursor, 32u);
        e = __shfl_sync(0xffffffffu, e, 0);
        if (e >= nw) break;
        e += lane;
        if (e >= nw) continue;
        const uint32_t i = S.widx[e], m = S.wbits[e];
        const uint32_t base = S.woff[e];
        uint32_t up = 0, ustarts = 0, ubase = 0;
        if (i >= (uint32_t)WW) {
            // the word above, if present, is at most WW - 1 entries back (entries are sorted by word index)
            const uint32_t target = i - WW;
            uint32_t lo = e > (uint32_t)WW ? e - WW : 0u, hi = e;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (S.widx[mid] < target)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            if (lo < e && S.widx[lo] == target) {
                up = S.wbits[lo];
                ustarts = up & ~(up << 1);
                ubase = S.woff[lo];
            }
        }
        const bool left_touch = (m & 1u) && e > 0 && S.widx[e - 1] == i - 1 && (i % WW) != 0 && (S.wbits[e - 1] >> 31);
        uint32_t rest = m, v = base;
        while (rest) {
            const int bit = __ffs(rest) - 1;
            const uint32_t shifted = ~(rest >> bit);
            const int len = shifted ? __ffs(shifted) - 1 : 32 - bit;
            const uint32_t runmask = (len >= 32 ? 0xffffffffu : ((1u << len) - 1u)) << bit;
            rest &= ~runmask;
            const uint32_t o = runmask & up;
            uint32_t p = v;
            if (o) {
                p = ubase + run_index(ustarts, __ffs(o) - 1);  // leftmost overlapping run above: smallest id
                // one segment per further run above that touches this run, plus the left neighbour
                uint32_t os = o & ~(o << 1);
                os &= os - 1;
                uint32_t extra = __popc(os) + ((bit == 0 && left_touch) ? 1u : 0u);
                if (extra) {
                    uint32_t slot = atomicAdd(&S.n_edges, extra);
                    if (slot + extra <= (uint32_t)kCapE) {
                        while (os) {
                            S.edges[slot++] = (v << 16) | (ubase + run_index(ustarts, __ffs(os) - 1));
                            os &= os - 1;
                        }
                        if (bit == 0 && left_touch) S.edges[slot] = (v << 16) | (base - 1);
                    }
                }
            } else if (bit == 0 && left_touch) {
                p = base - 1;
            }
            S.parent[v] = p;
            v++;
        }
    }
    stamp();  // 3
    __syncthreads();
    HV_EXP_STOP(3)
    const uint32_t ne = S.n_edges;
    if (ne > (uint32_t)kCapE) {  // block-uniform
        if (tid == 0) {
            b.frame_flags[f] = 1u;
            __threadfence();  // the flag is read (copy engine, host) by whoever sees the counter below
            if (b.ccl_done) atomicAdd(b.ccl_done, 1u);  // nothing of this frame's slot is touched from here on
        }
        return;
    }
    pointer_jump<kFT>(S.parent, nn, tid);
    stamp();  // 4
    __syncthreads();
    for (uint32_t k = tid; k < ne; k += kFT) {
        const uint32_t ed = S.edges[k];
        s_union(S.parent, ed >> 16, ed & 0xffffu);
    }
    stamp();  // 5
    __syncthreads();
    HV_EXP_STOP(5)

    // ---- phase 3: flatten, rank the roots ---------------------------------------------------------------------------------
    {
        volatile uint32_t *vp = S.parent;
        for (uint32_t v = tid; v < nn; v += kFT) {
            const uint32_t r = s_find(vp, v);
            vp[v] = r;  // concurrent readers see either the old ancestor or the root: both are valid ancestors
        }
    }
    __syncthreads();
    // blocked partition: thread t owns nodes [t*npt, (t+1)*npt)
    const uint32_t npt = (nn + kFT - 1) / kFT;
    const uint32_t v0 = min((uint32_t)tid * npt, nn), v1 = min(v0 + npt, nn);
    uint32_t nroots = 0;
    for (uint32_t v = v0; v < v1; v++) nroots += S.parent[v] == v ? 1u : 0u;
    uint32_t ncomp = 0;
    uint32_t rk = block_exclusive_scan<kNW>(nroots, S.warp_tmp, &ncomp);
    if (ncomp > (uint32_t)kCapB || ncomp > (uint32_t)b.blob_cap) {  // block-uniform
        if (tid == 0) {
            b.frame_flags[f] = 1u;
            __threadfence();  // the flag is read (copy engine, host) by whoever sees the counter below
            if (b.ccl_done) atomicAdd(b.ccl_done, 1u);  // nothing of this frame's slot is touched from here on
        }
        return;
    }
    for (uint32_t v = v0; v < v1; v++)
        if (S.parent[v] == v) S.rnk[v] = (uint16_t)rk++;
    for (uint32_t k = tid; k < ncomp; k += kFT) {
        S.b_area[k] = 0;
        S.b_sy[k] = 0;
        S.b_sx[k] = 0;
        S.b_ymin[k] = 0xffffffffu;
        S.b_ymax[k] = 0;
        S.b_xmin[k] = 0xffffffffu;
        S.b_xmax[k] = 0;
    }
    stamp();  // 6
    __syncthreads();
    HV_EXP_STOP(6)

    // ---- phase 4: labels + statistics, one node (run) per thread-iteration ----------------------------------------------
    // (Label stores: one thread per run, walking its own pixels.  A pixel-parallel version -- lane t takes pixel t of the
    // concatenated runs of the warp's 32 nodes, so that a run is one L2 request instead of one per pixel -- made this
    // kernel 2.4 us longer and the step no shorter: 40.35 vs 40.22 us.)
    int32_t *L = b.labels + (size_t)f * H * W;
    for (uint32_t v = tid; v < nn; v += kFT) {
        const uint32_t px = S.node_px[v], len = S.node_len[v];
        const uint32_t root = S.parent[v];
        const uint32_t rank = S.rnk[root];
        const uint32_t y = px / W, xs = px - y * W, xe = xs + len - 1;
        atomicAdd(&S.b_area[rank], len);
        atomicAdd(&S.b_sy[rank], y * len);
        atomicAdd(&S.b_sx[rank], ((xs + xe) * len) >> 1);
        atomicMax(&S.b_ymax[rank], y);
        atomicMin(&S.b_xmin[rank], xs);
        atomicMax(&S.b_xmax[rank], xe);
        if (root == v) S.b_ymin[rank] = y;  // the root run holds the raster-first pixel: its row is ymin
        int32_t *dst = L + px;
#ifdef HV_EXPERIMENTS
        if (b.phase_frame != -12345 - 9)
#endif
        for (uint32_t k = 0; k < len; k++) dst[k] = (int32_t)rank + 1;
    }
    stamp();  // 7
    __syncthreads();
    HV_EXP_STOP(7)

    // ---- phase 5: blob table -> global, scoring, ordered compaction, result ----------------------------------------------------
    hv_blob *blobs = b.blobs + (size_t)f * b.blob_cap;
    hv_defect *out = b.defects + (size_t)f * b.defect_cap;
    uint32_t nd_base = 0;
    for (uint32_t k0 = 0; k0 < ncomp; k0 += kFT) {
        const uint32_t k = k0 + tid;
        Scored sc;
        sc.keep = false;
        if (k < ncomp) {
            hv_blob q;
            q.area = S.b_area[k];
            q.ymin = S.b_ymin[k];
            q.ymax = S.b_ymax[k];
            q.xmin = S.b_xmin[k];
            q.xmax = S.b_xmax[k];
            q.reserved = 0;
            q.sum_y = S.b_sy[k];
            q.sum_x = S.b_sx[k];
            blobs[k] = q;
            sc = score_blob<(kFT < 512)>(b, sp, f, k, q);
        }
        uint32_t total = 0;
        const uint32_t dpos = nd_base + block_exclusive_scan<kNW>(sc.keep ? 1u : 0u, S.warp_tmp, &total);
        uint32_t a = sc.keep ? (uint32_t)sc.d.size : 0u;
        if (sc.keep) {
            if (dpos < (uint32_t)b.defect_cap) out[dpos] = sc.d;
            atomicAdd(&S.hist[area_bin(a)], 1u);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0 && a) atomicAdd(&S.area_sum, a);
        nd_base += total;
    }
    stamp();  // 8
    __syncthreads();
    if (tid == 0) {
        uint32_t fg = 0;
        for (int w = 0; w < kNW; w++) fg += S.warp_fg[w];
        const uint32_t nd = nd_base;
        const bool overflow = nd > (uint32_t)b.defect_cap;
        hv_frame_result r;
        r.n_components = ncomp;
        r.n_defects = min(nd, (uint32_t)b.defect_cap);
        r.defects_offset = 0;
        r.rejected = nd > 0 ? 1u : 0u;
        r.fg_pixels = fg;
        r.status = overflow ? HV_ERR_CAPACITY : HV_OK;
        b.results[f] = r;
        b.ncomp[f] = ncomp;
        b.fgcount[f] = fg;
        // bit 0 = "not done, needs the global path" (set on the early exits); bit 1 = done, but would not have fitted the
        // small build: the host uses it to decide when to go back to that build
        // bit 2 likewise for the tiny build
        const bool fits_small = nw <= (uint32_t)CclSmall::kCapW && nn <= (uint32_t)CclSmall::kCapN && ne <= (uint32_t)CclSmall::kCapE &&
                                ncomp <= (uint32_t)CclSmall::kCapB;
        const bool fits_tiny = nw <= (uint32_t)CclTiny::kCapW && nn <= (uint32_t)CclTiny::kCapN && ne <= (uint32_t)CclTiny::kCapE &&
                               ncomp <= (uint32_t)CclTiny::kCapB;
        b.frame_flags[f] = (fits_small ? 0u : 2u) | (fits_tiny ? 0u : 4u);
        unsigned long long *st = reinterpret_cast<unsigned long long *>(b.stats);
        atomicAdd(st + 0, 1ull);
        atomicAdd(st + 1, (unsigned long long)r.rejected);
        atomicAdd(st + 2, (unsigned long long)nd);
        atomicAdd(st + 3, (unsigned long long)ncomp);
        atomicAdd(st + 4, (unsigned long long)S.area_sum);
        atomicAdd(st + 5, (unsigned long long)fg);
        if (overflow) atomicAdd(st + 6 + HV_STATS_AREA_BINS, 1ull);
    }
    if (tid < HV_STATS_AREA_BINS && S.hist[tid])
        atomicAdd(reinterpret_cast<unsigned long long *>(b.stats) + 6 + tid, (unsigned long long)S.hist[tid]);
    // completion counter of the slot: the K1 that reuses the slot (two batches later) waits for it
    if (b.ccl_done) {
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicAdd(b.ccl_done, 1u);
        }
    }
    if (b.phase_ns && tid == 0 && f < 40) {  // debug: duration of this frame's CTA and its problem size
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        b.phase_ns[208 + f] = ((t - t_start) << 32) | ((unsigned long long)nw << 16) | (unsigned long long)ncomp;
    }
}

}  // namespace

bool ccl_frame_supported(const BatchView &b) { return b.h <= 8192 && b.w <= 8192; }

// per device, once (hv_create): opt in to the large dynamic shared-memory carve-out
cudaError_t configure_ccl_frame() {
    cudaError_t e = cudaFuncSetAttribute(k_ccl_frame<CclBig>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FrameSmem<CclBig>));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_ccl_frame<CclSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FrameSmem<CclSmall>));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_ccl_frame<CclTiny>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_ccl_frame<CclSmall>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

// level: 0 = big build, 1 = small, 2 = tiny
cudaError_t launch_ccl_frame(const BatchView &b, const ScoreParams &p, bool pdl, int level, cudaStream_t s) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(b.n);
    cfg.blockDim = dim3(level == 2 ? CclTiny::kFT : (level == 1 ? CclSmall::kFT : CclBig::kFT));
    cfg.dynamicSmemBytes = level == 2 ? sizeof(FrameSmem<CclTiny>) : (level == 1 ? sizeof(FrameSmem<CclSmall>) : sizeof(FrameSmem<CclBig>));
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    if (level == 2) return cudaLaunchKernelEx(&cfg, k_ccl_frame<CclTiny>, b, p);
    if (level == 1) return cudaLaunchKernelEx(&cfg, k_ccl_frame<CclSmall>, b, p);
    return cudaLaunchKernelEx(&cfg, k_ccl_frame<CclBig>, b, p);
}

}  // namespace hv
