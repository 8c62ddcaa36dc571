#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tmap, uint8_t* out, int x, int y, int z) {
    __shared__ __align__(128) uint8_t sm[6656];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(6624) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
            ::"r"(smem_u32(sm)), "l"(&tmap), "r"(smem_u32(&bar)), "r"(x), "r"(y), "r"(z) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < 6624; i += blockDim.x) out[i] = sm[i];
}
int main() {
    int W = 1280, H = 1024, N = 3;
    std::vector<uint8_t> h((size_t)W * H * N);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + (i >> 10));
    uint8_t *d, *o; cudaMalloc(&d, h.size()); cudaMalloc(&o, 6656);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry %d %d %p\n", (int)e, (int)q, p);
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)p;
    CUtensorMap tmap;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[2] = {(cuuint64_t)W, (cuuint64_t)W * H};
    cuuint32_t box[3] = {144, 46, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r);
    for (int t = 0; t < 3; t++) {
        int x = t == 0 ? 120 : (t == 1 ? -8 : 1272), y = t == 0 ? 25 : (t == 1 ? -7 : 1000), z = t;
        k<<<1, 256>>>(tmap, o, x, y, z);
        e = cudaDeviceSynchronize();
        printf("run %d: %s\n", t, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<uint8_t> got(6624); cudaMemcpy(got.data(), o, 6624, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r2 = 0; r2 < 46; r2++) for (int c = 0; c < 144; c++) {
            int gx = x + c, gy = y + r2; uint8_t exp = (gx < 0 || gx >= W || gy < 0 || gy >= H) ? 0 : h[((size_t)z * H + gy) * W + gx];
            if (got[r2 * 144 + c] != exp) bad++;
        }
        printf("  mismatches %d\n", bad);
    }
    return 0;
}
