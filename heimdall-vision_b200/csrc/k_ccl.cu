// k_ccl.cu -- K2..K5: 4-connected component labelling + per-blob statistics on the bit-packed mask.
//
// Reference semantics (rust/heimdall-core/src/detection.rs:215-245, same loops at :58-88 and processing.rs:322-353):
// a raster scan starts a DFS flood fill (4 neighbours) at every unvisited foreground pixel, so component k is the
// k-th component by the raster index of its first pixel.  The statistics the reference derives from each component's
// pixel list (detection.rs:247-255, 286-291) are area, sum of rows, sum of columns and the bounding box.
//
// Device formulation (no pixel lists, no DFS):
//   node      = "word-run": a maximal horizontal run of set bits inside one 32-pixel bitmask word.  Its id is the
//               linear pixel index p = y*w + x of its first pixel; its parent lives in labels[p] as parent+1
//               (K1 / bits_to_mask_labels initialise labels[p] = p+1 for every node and 0 elsewhere).
//   K2 merge  : union nodes that touch horizontally across a word boundary or overlap vertically; the union always
//               hooks the larger root under the smaller one (atomicMin), so every tree's root is the component's
//               minimum linear index = its raster-first pixel.
//   K3 flatten: every node points directly at its root; root nodes are flagged in rootbits, counted per word.
//   K4 scan   : exclusive prefix sum of the per-word root counts in raster order -> rank of every root; the
//               canonical label of a component is rank(root)+1, which is exactly the reference's discovery order.
//   K5 label  : every foreground pixel gets its canonical label (in place: a word's label slots are only ever read
//               and written by the thread that owns the word), blob statistics are accumulated with one set of
//               atomics per word-run.
// Only non-zero bitmask words do any work, so the cost follows the foreground, not the frame size.
#include "hv_common.cuh"

namespace hv {

namespace {

__device__ __forceinline__ int ld_parent(const int32_t *L, int a) {
    // parents are modified concurrently by atomics in L2: bypass L1
    return __ldcg(L + a) - 1;
}

__device__ __forceinline__ int uf_find(const int32_t *L, int a) {
    int p = ld_parent(L, a);
    while (p != a) {
        a = p;
        p = ld_parent(L, a);
    }
    return a;
}

// find with intermediate pointer jumping: every node visited on the way is re-pointed at its grandparent
// (atomicMin keeps the parent of a node monotonically decreasing, so concurrent unions are never undone).
__device__ __forceinline__ int uf_find_compress(int32_t *L, int a) {
    int p = ld_parent(L, a);
    if (p == a) return a;
    int prev = a;
    while (true) {
        const int next = ld_parent(L, p);
        if (next == p) return p;
        atomicMin(L + prev, next + 1);
        prev = p;
        p = next;
    }
}

// Lock-free union by minimum index.  Invariant: parent(a) <= a, so chains strictly decrease and terminate.
__device__ __forceinline__ void uf_union(int32_t *L, int a, int b) {
    while (true) {
        a = uf_find_compress(L, a);
        b = uf_find_compress(L, b);
        if (a == b) return;
        if (a < b) {
            const int t = a;
            a = b;
            b = t;
        }
        // a > b: try to hook root a under b
        const int old = atomicMin(L + a, b + 1) - 1;
        if (old == a) return;  // a was still a root: done
        a = old;               // somebody else hooked a under `old` first; keep merging old's tree with b
    }
}

// first bit of the run of set bits of m that contains bit `bit`
__device__ __forceinline__ int run_start(uint32_t m, int bit) {
    const uint32_t zeros_below = ~m & ((1u << bit) - 1u);
    return zeros_below ? 32 - __clz(zeros_below) : 0;
}

__global__ void __launch_bounds__(256) k_ccl_merge(BatchView b) {
    const size_t words_per_frame = (size_t)b.h * b.ww;
    const size_t total = words_per_frame * b.n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t m = b.bits[i];
        if (!m) continue;
        const size_t f = i / words_per_frame;
        if (b.frame_select && !(b.frame_select[f] & 1u)) continue;
        const int wi = (int)(i - f * words_per_frame);
        const int y = wi / b.ww, wx = wi - y * b.ww;
        int32_t *L = b.labels + f * (size_t)b.h * b.w;
        const int row = y * b.w + wx * 32;
        if ((m & 1u) && wx > 0) {
            const uint32_t left = b.bits[i - 1];
            if (left >> 31) uf_union(L, row, row - 32 + run_start(left, 31));
        }
        if (y > 0 && !b.conn8) {
            const uint32_t up = b.bits[i - b.ww];
            const uint32_t o = m & up;
            uint32_t starts = o & ~(o << 1);
            while (starts) {
                const int bit = __ffs(starts) - 1;
                starts &= starts - 1;
                uf_union(L, row + run_start(m, bit), row - b.w + run_start(up, bit));
            }
        } else if (y > 0) {
            // 8-connectivity: a run touches every run of the row above that meets it widened by one pixel on either side,
            // which may reach into the neighbouring words of that row
            const uint32_t up = b.bits[i - b.ww];
            const uint32_t upl = wx > 0 ? b.bits[i - b.ww - 1] : 0u, upr = wx + 1 < b.ww ? b.bits[i - b.ww + 1] : 0u;
            uint32_t rest = m;
            while (rest) {
                const int bit = __ffs(rest) - 1;
                const uint32_t shifted = ~(rest >> bit);
                const int len = shifted ? __ffs(shifted) - 1 : 32 - bit;
                const uint32_t runmask = (len >= 32 ? 0xffffffffu : ((1u << len) - 1u)) << bit;
                rest &= ~runmask;
                const int node = row + bit;
                const uint32_t touch = (runmask | (runmask << 1) | (runmask >> 1)) & up;
                uint32_t starts = touch & ~(touch << 1);
                while (starts) {
                    const int tb = __ffs(starts) - 1;
                    starts &= starts - 1;
                    uf_union(L, node, row - b.w + run_start(up, tb));
                }
                if (bit == 0 && (upl >> 31)) uf_union(L, node, row - b.w - 32 + run_start(upl, 31));
                if (bit + len == 32 && (upr & 1u)) uf_union(L, node, row - b.w + 32);
            }
        }
    }
}

// K3.  One block-iteration per (frame, segment of 256 consecutive words): every node of the segment is pointed at its
// root, the roots are flagged, and a block scan leaves in rankbase the number of roots in the words before each word
// inside the segment; the segment's total goes to segbase for K4.
__global__ void __launch_bounds__(256) k_ccl_flatten(BatchView b) {
    __shared__ uint32_t s_warp[8];
    const int words_per_frame = b.h * b.ww;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const size_t total = (size_t)b.nseg * b.n;
    for (size_t si = blockIdx.x; si < total; si += gridDim.x) {
        const size_t f = si / b.nseg;
        const int seg = (int)(si - f * b.nseg);
        if (b.frame_select && !(b.frame_select[f] & 1u)) continue;  // block-uniform
        const int wi = seg * 256 + tid;
        const size_t i = f * (size_t)words_per_frame + wi;
        const uint32_t m = wi < words_per_frame ? b.bits[i] : 0u;
        uint32_t roots = 0;
        if (m) {
            const int y = wi / b.ww, wx = wi - y * b.ww;
            int32_t *L = b.labels + f * (size_t)b.h * b.w;
            const int row = y * b.w + wx * 32;
            uint32_t starts = m & ~(m << 1);
            while (starts) {
                const int bit = __ffs(starts) - 1;
                starts &= starts - 1;
                const int s = row + bit;
                const int r = uf_find(L, s);
                if (r == s)
                    roots |= 1u << bit;
                else
                    L[s] = r + 1;
            }
        }
        const uint32_t c = __popc(roots);
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        __syncthreads();  // s_warp of the previous iteration has been read
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        uint32_t before = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            if (w < wid) before += s_warp[w];
            tot += s_warp[w];
        }
        if (wi < words_per_frame) {
            b.rootbits[i] = roots;
            b.rankbase[i] = before + incl - c;
        }
        if (tid == 0) b.segbase[si] = tot;
    }
}

// K4.  Exclusive scan of the per-segment root counts of a frame -> rank of every root (segbase[segment] +
// rankbase[word] + roots before it in its word); the canonical label of a component is that rank + 1, the reference's
// discovery order.  grid = (frames, R): every CTA sums the frame's counts (a few thousand values) to get the component
// count and resets its share of the blob table; CTA (f, 0) also writes the prefixes, ncomp and fgcount.
__global__ void __launch_bounds__(1024) k_ccl_scan(BatchView b) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_total;
    const int f = blockIdx.x;
    if (b.frame_select && !(b.frame_select[f] & 1u)) return;
    // counts in the first half of segbase, prefixes in the second: the other CTAs of the frame read the counts too
    const uint32_t *cnt = b.segbase + (size_t)f * b.nseg;
    uint32_t *pre = b.segbase + ((size_t)b.n + f) * b.nseg;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nseg = b.nseg;
    const int per = (nseg + 1023) / 1024;  // consecutive segments per thread
    const int i0 = min(tid * per, nseg), i1 = min(i0 + per, nseg);
    uint32_t sum = 0;
    for (int i = i0; i < i1; i++) sum += cnt[i];
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const uint32_t w = s_warp[lane];
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        s_warp[lane] = wi - w;
        if (lane == 31) s_total = wi;
    }
    __syncthreads();
    const uint32_t ncomp = s_total;
    if (blockIdx.y == 0) {
        uint32_t run = s_warp[wid] + incl - sum;
        for (int i = i0; i < i1; i++) {
            pre[i] = run;
            run += cnt[i];
        }
        if (tid == 0) {
            b.ncomp[f] = ncomp;
            b.fgcount[f] = 0;
        }
    }
    hv_blob *blobs = b.blobs + (size_t)f * b.blob_cap;
    const uint32_t nb = min(ncomp, (uint32_t)b.blob_cap);
    for (uint32_t k = blockIdx.y * 1024u + tid; k < nb; k += gridDim.y * 1024u) {
        hv_blob z;
        z.area = 0;
        z.ymin = 0xffffffffu;
        z.ymax = 0;
        z.xmin = 0xffffffffu;
        z.xmax = 0;
        z.reserved = 0;
        z.sum_y = 0;
        z.sum_x = 0;
        blobs[k] = z;
    }
}

// K5.  A warp owns a patch of 4 words x 8 rows (128 x 8 pixels) so that runs of the same blob in neighbouring rows
// and words meet in one warp: lanes whose current run carries the same label are grouped with __match_any_sync, their
// partial statistics are combined with redux.sync, and one lane per group issues the atomics.
__global__ void __launch_bounds__(256) k_ccl_label(BatchView b) {
    const int lane = threadIdx.x & 31;
    const int pw = (b.ww + 3) >> 2, ph = (b.h + 7) >> 3;  // patches per row / per column
    const size_t patches_per_frame = (size_t)pw * ph;
    const size_t total = patches_per_frame * b.n;
    const size_t words_per_frame = (size_t)b.h * b.ww;
    const size_t warp0 = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t pi = warp0; pi < total; pi += nwarps) {
        const size_t f = pi / patches_per_frame;
        if (b.frame_select && !(b.frame_select[f] & 1u)) continue;
        const int pj = (int)(pi - f * patches_per_frame);
        const int py = pj / pw, px = pj - py * pw;
        const int y = py * 8 + (lane >> 2), wx = px * 4 + (lane & 3);
        const bool inside = y < b.h && wx < b.ww;
        const uint32_t m = inside ? b.bits[f * words_per_frame + (size_t)y * b.ww + wx] : 0u;
        const uint32_t any = __ballot_sync(0xffffffffu, m != 0);
        if (!any) continue;
        int32_t *L = b.labels + f * (size_t)b.h * b.w;
        const uint32_t *rootbits = b.rootbits + f * words_per_frame;
        const uint32_t *rankbase = b.rankbase + f * words_per_frame;
        const uint32_t *segbase = b.segbase + ((size_t)b.n + f) * b.nseg;  // the prefixes (second half)
        hv_blob *blobs = b.blobs + f * (size_t)b.blob_cap;
        const int row = y * b.w + wx * 32;
        uint32_t fgsum = __popc(m);
#pragma unroll
        for (int o = 16; o; o >>= 1) fgsum += __shfl_xor_sync(0xffffffffu, fgsum, o);
        if (lane == 0) atomicAdd(b.fgcount + f, fgsum);
        uint32_t rest = m;
        while (true) {
            const uint32_t act = __ballot_sync(0xffffffffu, rest != 0);
            if (!act) break;
            if (rest) {
                const int bit = __ffs(rest) - 1;
                const uint32_t shifted = ~(rest >> bit);
                const int len = shifted ? __ffs(shifted) - 1 : 32 - bit;  // run [bit, bit+len)
                rest = (bit + len >= 32) ? 0u : (rest & ~(((1u << len) - 1u) << bit));
                const int s = row + bit;
                const int r = L[s] - 1;  // flattened by K3: the root itself
                const int ry = r / b.w, rx = r - ry * b.w;
                const int rw = ry * b.ww + (rx >> 5);
                const uint32_t rank = segbase[rw >> 8] + rankbase[rw] + __popc(rootbits[rw] & ((1u << (rx & 31)) - 1u));
                const int label = (int)rank + 1;
                for (int k = 0; k < len; k++) L[s + k] = label;
                const uint32_t xs = wx * 32 + bit, xe = xs + len - 1;
                const uint32_t grp = __match_any_sync(act, rank);
                const uint32_t area = __reduce_add_sync(grp, (uint32_t)len);
                const uint32_t sy = __reduce_add_sync(grp, (uint32_t)y * (uint32_t)len);       // <= 8 rows x 128 px x y
                const uint32_t sx = __reduce_add_sync(grp, ((xs + xe) * (uint32_t)len) >> 1);  // fits 32 bits
                const uint32_t ymax = __reduce_max_sync(grp, (uint32_t)y);
                const uint32_t xmin = __reduce_min_sync(grp, xs);
                const uint32_t xmax = __reduce_max_sync(grp, xe);
                if (rank < (uint32_t)b.blob_cap) {
                    hv_blob *q = blobs + rank;
                    if (r == s) q->ymin = (uint32_t)y;  // the root run is the raster-first pixel: its row is ymin
                    if (lane == __ffs(grp) - 1) {
                        atomicAdd(&q->area, area);
                        atomicAdd(reinterpret_cast<unsigned long long *>(&q->sum_y), (unsigned long long)sy);
                        atomicAdd(reinterpret_cast<unsigned long long *>(&q->sum_x), (unsigned long long)sx);
                        atomicMax(&q->ymax, ymax);
                        atomicMin(&q->xmin, xmin);
                        atomicMax(&q->xmax, xmax);
                    }
                }
            }
        }
    }
}

// mask bytes + label-plane initialisation from a bit-packed mask (used after morphology, and by find_contours).
__global__ void __launch_bounds__(256) k_bits_to_mask_labels(BatchView b) {
    const size_t words_per_frame = (size_t)b.h * b.ww;
    const size_t total = words_per_frame * b.n;
    const int lane = threadIdx.x & 31;
    // one warp per word: lane = pixel, so the 32 mask bytes / 32 labels of a word go out as one coalesced store
    const size_t warp0 = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t i = warp0; i < total; i += nwarps) {
        const uint32_t m = b.bits[i];
        const size_t f = i / words_per_frame;
        const int wi = (int)(i - f * words_per_frame);
        const int y = wi / b.ww, wx = wi - y * b.ww;
        const int x = wx * 32 + lane;
        if (x >= b.w) continue;
        const size_t px = (f * b.h + y) * (size_t)b.w + x;
        const bool fg = (m >> lane) & 1u;
        const bool start = fg && (lane == 0 || !((m >> (lane - 1)) & 1u));
        b.mask[px] = fg ? 255 : 0;
        b.labels[px] = start ? (y * b.w + x + 1) : 0;
    }
}

// occupancy nibbles from a bit-packed mask (needed when morphology rewrote the bits K1 had summarised)
__global__ void __launch_bounds__(256) k_rowflags_from_bits(BatchView b) {
    const size_t per_frame = (size_t)b.h * b.tiles_x;
    const size_t total = per_frame * b.n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t f = i / per_frame;
        const int j = (int)(i - f * per_frame);
        const int y = j / b.tiles_x, tx = j - y * b.tiles_x;
        const uint32_t *row = b.bits + (f * b.h + y) * (size_t)b.ww;
        uint32_t nib = 0;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (4 * tx + k < b.ww && row[4 * tx + k]) nib |= 1u << k;
        b.rowflags[f * b.rf_stride + rowflag_index(y, tx, b.tiles_x)] = (uint8_t)nib;
    }
}

// After a K1 launch with sparse_aux the bit-mask words of flat tiles were left unwritten (their occupancy records are all
// zero).  Zero them, so that the kernels that scan every word of the frame (global-memory CCL path) see a dense mask.
// One warp per tile: lane = row of the tile.
__global__ void __launch_bounds__(256) k_densify_bits(BatchView b) {
    const int tiles_y = (b.h + 31) / 32;
    const size_t per_frame = (size_t)tiles_y * b.tiles_x;
    const size_t total = per_frame * b.n;
    const int lane = threadIdx.x & 31;
    for (size_t t = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5; t < total; t += ((size_t)gridDim.x * blockDim.x) >> 5) {
        const size_t f = t / per_frame;
        if (b.frame_select && !(b.frame_select[f] & 1u)) continue;
        const int j = (int)(t - f * per_frame);
        const int ty = j / b.tiles_x, tx = j - ty * b.tiles_x;
        const int y = ty * 32 + lane;
        const uint32_t fl = y < b.h ? b.rowflags[f * b.rf_stride + rowflag_index(y, tx, b.tiles_x)] : 0u;
        if (__any_sync(0xffffffffu, fl != 0) || y >= b.h) continue;
        uint32_t *row = b.bits + (f * b.h + y) * (size_t)b.ww;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (4 * tx + k < b.ww) row[4 * tx + k] = 0u;
    }
}

int grid_for(size_t work_items, int per_block) {
    size_t g = (work_items + per_block - 1) / per_block;
    const size_t cap = 148 * 32;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace

cudaError_t launch_ccl_merge(const BatchView &b, cudaStream_t s) {
    k_ccl_merge<<<grid_for((size_t)b.n * b.h * b.ww, 256), 256, 0, s>>>(b);
    return cudaGetLastError();
}
cudaError_t launch_ccl_flatten(const BatchView &b, cudaStream_t s) {
    k_ccl_flatten<<<grid_for((size_t)b.n * b.nseg * 256, 256), 256, 0, s>>>(b);
    return cudaGetLastError();
}
cudaError_t launch_ccl_scan(const BatchView &b, cudaStream_t s) {
    // the blob-table reset is shared by a few CTAs per frame when the table can be large
    k_ccl_scan<<<dim3(b.n, b.blob_cap > 8192 ? 8 : 1), 1024, 0, s>>>(b);
    return cudaGetLastError();
}
cudaError_t launch_ccl_label(const BatchView &b, cudaStream_t s) {
    const size_t patches = (size_t)b.n * ((b.ww + 3) / 4) * ((b.h + 7) / 8);
    k_ccl_label<<<grid_for(patches * 32, 256), 256, 0, s>>>(b);
    return cudaGetLastError();
}
cudaError_t launch_rowflags_from_bits(const BatchView &b, cudaStream_t s) {
    k_rowflags_from_bits<<<grid_for((size_t)b.n * b.h * b.tiles_x, 256), 256, 0, s>>>(b);
    return cudaGetLastError();
}
cudaError_t launch_densify_bits(const BatchView &b, cudaStream_t s) {
    k_densify_bits<<<grid_for((size_t)b.n * ((b.h + 31) / 32) * b.tiles_x * 32, 256), 256, 0, s>>>(b);
    return cudaGetLastError();
}
cudaError_t launch_bits_to_mask_labels(const BatchView &b, cudaStream_t s) {
    k_bits_to_mask_labels<<<grid_for((size_t)b.n * b.h * b.ww * 32, 256), 256, 0, s>>>(b);
    return cudaGetLastError();
}

}  // namespace hv
