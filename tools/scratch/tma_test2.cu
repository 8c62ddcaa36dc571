#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
namespace cde = cuda::device::experimental;
using barrier = cuda::barrier<cuda::thread_scope_block>;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// variant 0: libcu++ wrappers, 3D
__global__ void k_lib(const __grid_constant__ CUtensorMap tmap, uint8_t* out, int x, int y, int z, int bytes) {
    __shared__ __align__(128) uint8_t sm[8192];
    #pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_3d_global_to_shared(sm, &tmap, x, y, z, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, bytes);
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[i];
}
// variant 1: raw PTX (as in the product kernel)
__global__ void k_raw(const __grid_constant__ CUtensorMap tmap, uint8_t* out, int x, int y, int z, int bytes) {
    __shared__ __align__(128) uint8_t sm[8192];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
            ::"r"(smem_u32(sm)), "l"(&tmap), "r"(smem_u32(&bar)), "r"(x), "r"(y), "r"(z) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[i];
}
int main(int argc, char** argv) {
    int variant = argc > 1 ? atoi(argv[1]) : 0;
    int bw = argc > 2 ? atoi(argv[2]) : 144, bh = argc > 3 ? atoi(argv[3]) : 46;
    int W = 1280, H = 1024, N = 3;
    std::vector<uint8_t> h((size_t)W * H * N);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + (i >> 10));
    uint8_t *d, *o; cudaMalloc(&d, h.size()); cudaMalloc(&o, 8192);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)p;
    CUtensorMap tmap;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[2] = {(cuuint64_t)W, (cuuint64_t)W * H};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d box %dx%d encode %d\n", variant, bw, bh, (int)r);
    int bytes = bw * bh;
    for (int t = 0; t < 3; t++) {
        int x = t == 0 ? 112 : (t == 1 ? -16 : 1264), y = t == 0 ? 25 : (t == 1 ? -7 : 1000), z = t;
        if (variant == 0) k_lib<<<1, 256>>>(tmap, o, x, y, z, bytes); else k_raw<<<1, 256>>>(tmap, o, x, y, z, bytes);
        cudaError_t e = cudaDeviceSynchronize();
        printf("run %d: %s\n", t, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<uint8_t> got(bytes); cudaMemcpy(got.data(), o, bytes, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r2 = 0; r2 < bh; r2++) for (int c = 0; c < bw; c++) {
            int gx = x + c, gy = y + r2; uint8_t exp = (gx < 0 || gx >= W || gy < 0 || gy >= H) ? 0 : h[((size_t)z * H + gy) * W + gx];
            if (got[r2 * bw + c] != exp) bad++;
        }
        printf("  mismatches %d\n", bad);
    }
    return 0;
}
