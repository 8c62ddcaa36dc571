// Does L2 compute-data compression (cuMemCreate + CU_MEM_ALLOCATION_COMP_GENERIC) speed up K1's write mix?
// 1 B/px read, 1 B/px mask write (zeros), 4 B/px label write (zeros, or `nzfrac` of the 128-B lines non-zero).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <functional>
#define DRV(name) ((decltype(&name))drv(#name))
static void* drv(const char* n) { void* p = nullptr; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint(n, &p, cudaEnableDefault, &q); if (!p) { printf("no %s\n", n); exit(1); } return p; }
static void* vmm_alloc(size_t bytes, bool comp, int* got) {
  CUmemAllocationProp prop = {}; prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 0;
  if (comp) prop.allocFlags.compressionType = CU_MEM_ALLOCATION_COMP_GENERIC;
  size_t gran = 0; DRV(cuMemGetAllocationGranularity)(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM);
  bytes = (bytes + gran - 1) / gran * gran;
  CUmemGenericAllocationHandle h; CUresult r = DRV(cuMemCreate)(&h, bytes, &prop, 0); if (r) { printf("cuMemCreate %d\n", r); exit(1); }
  CUmemAllocationProp p2 = {}; DRV(cuMemGetAllocationPropertiesFromHandle)(&p2, h); *got = p2.allocFlags.compressionType;
  CUdeviceptr d; r = DRV(cuMemAddressReserve)(&d, bytes, 0, 0, 0); if (r) { printf("reserve %d\n", r); exit(1); }
  r = DRV(cuMemMap)(d, bytes, 0, h, 0); if (r) { printf("map %d\n", r); exit(1); }
  CUmemAccessDesc a = {}; a.location = prop.location; a.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE; r = DRV(cuMemSetAccess)(d, bytes, &a, 1); if (r) { printf("access %d\n", r); exit(1); }
  return (void*)d; }
__global__ void k_mix(const uint4* in, int4* lab, uint4* mask, size_t npx16, unsigned nz_every) {
  const int lane = threadIdx.x & 31;
  for (size_t w0 = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) / 32 * 32; w0 < npx16; w0 += (size_t)gridDim.x * blockDim.x) {
    uint4 v = in[w0 + lane]; const int nz = (nz_every && ((w0 / 32) % nz_every) == 0) ? (int)(w0 + lane) * 2654435761u : 0;
    const int4 z = make_int4(nz, 0, nz, 0);
    mask[w0 + lane] = make_uint4(v.x & 0, 0, 0, 0); int4* l = lab + 4 * w0;
#pragma unroll
    for (int k = 0; k < 4; k++) l[32 * k + lane] = z; } }
__global__ void k_read(const int4* lab, size_t n, int* sink) { int acc = 0; for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { int4 v = lab[i]; acc |= v.x | v.y | v.z | v.w; } if (acc == 0x12345) *sink = acc; }
float timeit(std::function<void(int)> f, int reps) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); for (int i = 0; i < 8; i++) f(i); cudaEventRecord(a); for (int i = 0; i < reps; i++) f(i + 8); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps; }
int main() {
  cudaFree(0); int sup = -1; CUdevice dev; DRV(cuDeviceGet)(&dev, 0); DRV(cuDeviceGetAttribute)(&sup, CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED, dev); printf("generic compression supported: %d\n", sup);
  const int n = 25, H = 1024, W = 1280; const size_t npx = (size_t)n * H * W; const int NB = 8;
  uint8_t* in; cudaMalloc(&in, npx * NB); cudaMemset(in, 1, npx * NB); int* sink; cudaMalloc(&sink, 4);
  for (int comp = -1; comp < (sup == 1 ? 2 : 1); comp++) {
    int g1 = -1, g2 = -1; uint8_t* mask; int32_t* lab;
    if (comp < 0) { cudaMalloc(&mask, npx * NB); cudaMalloc(&lab, npx * 4 * NB); }  // plain cudaMalloc
    else { mask = (uint8_t*)vmm_alloc(npx * NB, comp, &g1); lab = (int32_t*)vmm_alloc(npx * 4 * NB, comp, &g2); }
    printf("comp requested %d, got mask %d labels %d\n", comp, g1, g2);
    for (unsigned nz : {0u, 1u}) for (int g : {592, 2368}) {
      float ms = timeit([&](int i) { int s = i % NB; k_mix<<<g, 256>>>((const uint4*)(in + npx * s), (int4*)(lab + npx * s), (uint4*)(mask + npx * s), npx / 16, nz); }, 24);
      printf("  comp %d nonzero-every %2u grid %4d: %6.1f us  %6.0f GB/s alg\n", comp, nz, g, ms * 1e3, npx * 6.0 / ms / 1e6); }
    float ms = timeit([&](int i) { int s = i % NB; k_read<<<2368, 256>>>((const int4*)(lab + npx * s), npx / 4, sink); }, 24);
    printf("  comp %d read back labels (last written pattern): %6.1f us %6.0f GB/s\n", comp, ms * 1e3, npx * 4.0 / ms / 1e6);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize())); return 0; }
