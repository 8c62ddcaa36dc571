// Does a burst L2 prefetch of the input ahead of the write stream beat interleaved reads (DRAM read/write turnaround)?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <functional>
// mode 0: plain; 1: prefetch.global.L2 whole slice first; 2: L2::evict_last prefetch; 3: cp.async.bulk.prefetch.L2
__global__ void k_mix(const uint4* in, int4* lab, uint4* mask, size_t npx16, int mode) {
  const int4 z = make_int4(0,0,0,0); const int lane = threadIdx.x & 31;
  if (mode) { // every CTA prefetches a contiguous slice of the input, 128 B per thread-iteration
    const size_t bytes = npx16 * 16, per = (bytes / gridDim.x + 127) & ~(size_t)127; const char* base = (const char*)in + per * blockIdx.x; const size_t lim = min(per, bytes - min(bytes, per * blockIdx.x));
    if (mode == 3) { if (threadIdx.x == 0) for (size_t o = 0; o < lim; o += 16384) { const unsigned sz = (unsigned)min((size_t)16384, lim - o); asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(base + o), "r"(sz) : "memory"); } }
    else for (size_t o = threadIdx.x * 128; o < lim; o += blockDim.x * 128) { if (mode == 1) asm volatile("prefetch.global.L2 [%0];" :: "l"(base + o)); else asm volatile("prefetch.global.L2::evict_last [%0];" :: "l"(base + o)); }
  }
  for (size_t w0 = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) / 32 * 32; w0 < npx16; w0 += (size_t)gridDim.x * blockDim.x) {
    uint4 v = in[w0 + lane]; mask[w0 + lane] = v; int4* l = lab + 4 * w0;
#pragma unroll
    for (int k = 0; k < 4; k++) l[32 * k + lane] = z; } }
float timeit(std::function<void(int)> f, int reps) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); for (int i = 0; i < 3; i++) f(i); cudaEventRecord(a); for (int i = 0; i < reps; i++) f(i + 3); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps; }
int main() {
  const int n = 25, H = 1024, W = 1280; const size_t npx = (size_t)n * H * W; const int NB = 8;
  uint8_t *in, *mask; int32_t* lab; cudaMalloc(&in, npx * NB); cudaMalloc(&mask, npx * NB); cudaMalloc(&lab, npx * 4 * NB); cudaMemset(in, 1, npx * NB);
  for (int mode = 0; mode < 4; mode++) for (int g : {592, 2368}) {
    float ms = timeit([&](int i) { int s = i % NB; k_mix<<<g, 256>>>((const uint4*)(in + npx * s), (int4*)(lab + npx * s), (uint4*)(mask + npx * s), npx / 16, mode); }, 24);
    printf("mode %d grid %4d: %6.1f us  %6.0f GB/s\n", mode, g, ms * 1e3, npx * 6.0 / ms / 1e6); }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize())); return 0; }
