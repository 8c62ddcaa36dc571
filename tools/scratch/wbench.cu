#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
// write-bandwidth micro-benchmarks: which store pattern gets closest to the HBM ceiling for a 4 B/px zero label plane
// + 1 B/px mask + 1 B/px read, at the K1 traffic mix (164 MB written, 33 MB read per launch)
__global__ void k_st128(int4* p, size_t n) { size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; const int4 z = make_int4(0,0,0,0);
  for (; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = z; }
__global__ void k_st128_cs(int4* p, size_t n) { size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; const int4 z = make_int4(0,0,0,0);
  for (; i < n; i += (size_t)gridDim.x * blockDim.x) __stcs(p + i, z); }
__global__ void k_st128_wt(int4* p, size_t n) { size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; const int4 z = make_int4(0,0,0,0);
  for (; i < n; i += (size_t)gridDim.x * blockDim.x) __stwt(p + i, z); }
// copy-like mix: read 1 B/px, write 5 B/px
__global__ void k_mix(const uint4* in, int4* lab, uint4* mask, size_t npx16) { size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; const int4 z = make_int4(0,0,0,0);
  for (; i < npx16; i += (size_t)gridDim.x * blockDim.x) { uint4 v = in[i]; uint4 m; m.x = v.x & 0; m.y = v.y & 0; m.z = v.z & 0; m.w = v.w & 0; mask[i] = m; lab[4*i] = z; lab[4*i+1] = z; lab[4*i+2] = z; lab[4*i+3] = z; } }
template <typename F> float timeit(F f, int reps) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); for (int i = 0; i < 3; i++) f(); cudaEventRecord(a); for (int i = 0; i < reps; i++) f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps; }
int main() {
  const size_t npx = 25ull * 1280 * 1024; int4* lab; uint4 *mask, *in; cudaMalloc(&lab, npx * 4 * 2); cudaMalloc(&mask, npx * 2); cudaMalloc(&in, npx * 8);
  cudaMemset(in, 1, npx * 8);
  const size_t n16 = npx * 4 / 16;
  for (int g : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
    float t1 = timeit([&] { k_st128<<<g, 256>>>(lab, n16); }, 20);
    float t2 = timeit([&] { k_st128_cs<<<g, 256>>>(lab, n16); }, 20);
    float t3 = timeit([&] { k_st128_wt<<<g, 256>>>(lab, n16); }, 20);
    printf("grid %5d: st128 %.1f us (%.0f GB/s)  st.cs %.1f us (%.0f GB/s)  st.wt %.1f us (%.0f GB/s)\n", g, t1 * 1e3, npx * 4 / t1 / 1e6, t2 * 1e3, npx * 4 / t2 / 1e6, t3 * 1e3, npx * 4 / t3 / 1e6);
  }
  float tm = timeit([&] { cudaMemsetAsync(lab, 0, npx * 4); }, 20);
  printf("cudaMemset 131 MB: %.1f us (%.0f GB/s)\n", tm * 1e3, npx * 4 / tm / 1e6);
  // alternate between two label buffers so writes cannot coalesce in L2 across launches
  int rot = 0;
  float t4 = timeit([&] { k_mix<<<148 * 16, 256>>>(in + (rot % 8) * (npx / 16), lab + (rot % 2) * (npx / 4), mask + (rot % 2) * (npx / 16), npx / 16); rot++; }, 20);
  printf("mix (1 B/px read + 5 B/px write, 197 MB): %.1f us (%.0f GB/s total)\n", t4 * 1e3, npx * 6 / t4 / 1e6);
  float tc = timeit([&] { cudaMemcpyAsync(lab, lab + npx / 4, npx * 4, cudaMemcpyDeviceToDevice); }, 20);
  printf("cudaMemcpy D2D 131 MB: %.1f us (%.0f GB/s r+w)\n", tc * 1e3, npx * 8 / tc / 1e6);
  return 0; }
