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
// a: barrier only
__global__ void k_bar(int* out) {
    #pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    auto t = bar.arrive(); bar.wait(std::move(t));
    if (threadIdx.x == 0) out[0] = 42;
}
// b: 1D bulk copy
__global__ void k_bulk(const int* src, int* out) {
    __shared__ alignas(128) int sm[1024];
    #pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) { cuda::memcpy_async(sm, src, cuda::aligned_size_t<16>(4096), bar); token = bar.arrive(); } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = sm[i];
}
// c: 2D int32 TMA (programming guide example)
__global__ void k_2d(const __grid_constant__ CUtensorMap tmap, int* out, int x, int y) {
    __shared__ alignas(128) int sm[16][16];
    #pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) { cde::cp_async_bulk_tensor_2d_global_to_shared(&sm, &tmap, x, y, bar); token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(sm)); }
    else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < 256; i += blockDim.x) out[i] = sm[i / 16][i % 16];
}
int main(int argc, char** argv) {
    int variant = argc > 1 ? atoi(argv[1]) : 0;
    int *d, *o; cudaMalloc(&d, 256 * 256 * 4); cudaMalloc(&o, 4096);
    std::vector<int> h(256 * 256); for (int i = 0; i < 65536; i++) h[i] = i;
    cudaMemcpy(d, h.data(), 65536 * 4, cudaMemcpyHostToDevice);
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0); 
    int drv=0, rt=0; cudaDriverGetVersion(&drv); cudaRuntimeGetVersion(&rt);
    printf("dev %s cc %d.%d drv %d rt %d | ", pr.name, pr.major, pr.minor, drv, rt);
    if (variant == 0) k_bar<<<1, 128>>>(o);
    else if (variant == 1) k_bulk<<<1, 128>>>(d, o);
    else {
        void *p = nullptr; cudaDriverEntryPointQueryResult q;
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        auto enc = (PFN_cuTensorMapEncodeTiled_v12000)p;
        CUtensorMap tmap;
        cuuint64_t dims[2] = {256, 256}; cuuint64_t strides[1] = {256 * 4}; cuuint32_t box[2] = {16, 16}; cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode %d | ", (int)r);
        unsigned char* tb = (unsigned char*)&tmap; for (int i = 0; i < 32; i++) printf("%02x", tb[i]); printf(" | ");
        k_2d<<<1, 128>>>(tmap, o, 32, 48);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d: %s", variant, cudaGetErrorString(e));
    if (e == cudaSuccess) { int v[4]; cudaMemcpy(v, o, 16, cudaMemcpyDeviceToHost); printf(" out %d %d %d %d", v[0], v[1], v[2], v[3]); }
    printf("\n");
    return 0;
}
