// Write-path micro-benchmarks on large rotating buffers (no L2 write absorption across launches).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <functional>
__global__ void k_st(int4* p, size_t n) { size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; const int4 z = make_int4(0,0,0,0);
  for (; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = z; }
__global__ void k_st_cs(int4* p, size_t n) { size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; const int4 z = make_int4(0,0,0,0);
  for (; i < n; i += (size_t)gridDim.x * blockDim.x) __stcs(p + i, z); }
__global__ void k_st_ef(int4* p, size_t n) { // L2 evict_first policy
  uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; i < n; i += (size_t)gridDim.x * blockDim.x) asm volatile("st.global.L2::cache_hint.v4.s32 [%0], {%1,%1,%1,%1}, %2;" :: "l"(p + i), "r"(0), "l"(pol) : "memory"); }
// contiguous chunk per CTA (each CTA owns n/grid consecutive int4)
__global__ void k_st_chunk(int4* p, size_t n) { const size_t per = n / gridDim.x; int4* q = p + per * blockIdx.x; const int4 z = make_int4(0,0,0,0);
  for (size_t i = threadIdx.x; i < per; i += blockDim.x) q[i] = z; }
// TMA bulk store from a zeroed smem buffer, 8 KB per bulk op, one thread issues
__global__ void k_st_bulk(char* p, size_t bytes) { __shared__ __align__(128) char z[8192];
  for (int i = threadIdx.x; i < 8192 / 16; i += blockDim.x) reinterpret_cast<int4*>(z)[i] = make_int4(0,0,0,0);
  __syncthreads(); asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) { const size_t nchunks = bytes / 8192; uint32_t s = (uint32_t)__cvta_generic_to_shared(z);
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x)
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 8192;" :: "l"(p + c * 8192), "r"(s) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory"); asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); } }
// mix: read R bytes/px-ish: in (1 B/px), write mask (1 B/px) + labels (4 B/px)
__global__ void k_mix(const uint4* in, int4* lab, uint4* mask, size_t npx16) { size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; const int4 z = make_int4(0,0,0,0);
  for (; i < npx16; i += (size_t)gridDim.x * blockDim.x) { uint4 v = in[i]; uint4 m; m.x = v.x & 0; m.y = v.y & 0; m.z = v.z & 0; m.w = v.w & 0; mask[i] = m; lab[4*i] = z; lab[4*i+1] = z; lab[4*i+2] = z; lab[4*i+3] = z; } }
__global__ void k_rd(const uint4* in, size_t n, unsigned* out) { size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; unsigned a = 0;
  for (; i < n; i += (size_t)gridDim.x * blockDim.x) { uint4 v = in[i]; a ^= v.x ^ v.y ^ v.z ^ v.w; } if (a == 0x12345) *out = a; }
__global__ void k_copy(const uint4* in, uint4* o, size_t n) { size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; i < n; i += (size_t)gridDim.x * blockDim.x) o[i] = in[i]; }
float timeit(std::function<void(int)> f, int reps) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); for (int i = 0; i < 3; i++) f(i); cudaEventRecord(a); for (int i = 0; i < reps; i++) f(i + 3); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps; }
int main() {
  const size_t CH = 164ull << 20;   // 164 MiB per launch (mask + labels of the headline batch)
  const int NB = 8;                 // rotate over 8 distinct regions = 1.3 GiB
  char* buf; cudaMalloc(&buf, CH * NB); cudaMemset(buf, 1, CH * NB); unsigned* out; cudaMalloc(&out, 4);
  auto P = [&](const char* name, float ms, double bytes) { printf("%-44s %7.1f us  %6.0f GB/s\n", name, ms * 1e3, bytes / ms / 1e6); };
  const size_t n16 = CH / 16;
  for (int g : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) { char nm[64];
    snprintf(nm, 64, "st.128 grid-stride grid=%d", g); P(nm, timeit([&](int i) { k_st<<<g, 256>>>((int4*)(buf + CH * (i % NB)), n16); }, 24), CH);
  }
  P("st.128.cs grid=2368", timeit([&](int i) { k_st_cs<<<2368, 256>>>((int4*)(buf + CH * (i % NB)), n16); }, 24), CH);
  P("st.128 L2 evict_first grid=2368", timeit([&](int i) { k_st_ef<<<2368, 256>>>((int4*)(buf + CH * (i % NB)), n16); }, 24), CH);
  P("st.128 chunk-per-CTA grid=592", timeit([&](int i) { k_st_chunk<<<592, 256>>>((int4*)(buf + CH * (i % NB)), n16); }, 24), CH);
  P("st.128 chunk-per-CTA grid=2368", timeit([&](int i) { k_st_chunk<<<2368, 256>>>((int4*)(buf + CH * (i % NB)), n16); }, 24), CH);
  P("TMA bulk store 8KB grid=592", timeit([&](int i) { k_st_bulk<<<592, 128>>>(buf + CH * (i % NB), CH); }, 24), CH);
  P("TMA bulk store 8KB grid=148", timeit([&](int i) { k_st_bulk<<<148, 128>>>(buf + CH * (i % NB), CH); }, 24), CH);
  P("cudaMemsetAsync", timeit([&](int i) { cudaMemsetAsync(buf + CH * (i % NB), 0, CH); }, 24), CH);
  P("read only 164 MiB grid=2368", timeit([&](int i) { k_rd<<<2368, 256>>>((const uint4*)(buf + CH * (i % NB)), n16, out); }, 24), CH);
  P("copy 82->82 MiB grid=2368", timeit([&](int i) { k_copy<<<2368, 256>>>((const uint4*)(buf + CH * (i % NB)), (uint4*)(buf + CH * ((i + 4) % NB)), n16 / 2); }, 24), CH);
  // K1 mix: 25 frames 1280x1024: 32.8 MB in, 32.8 MB mask, 131 MB labels
  const size_t npx = 25ull * 1280 * 1024;
  P("mix 1R:5W grid=2368", timeit([&](int i) { char* b = buf + CH * (i % NB); k_mix<<<2368, 256>>>((const uint4*)(buf + CH * ((i + 3) % NB)), (int4*)b, (uint4*)(b + npx * 4), npx / 16); }, 24), npx * 6.0);
  P("mix 1R:5W grid=592", timeit([&](int i) { char* b = buf + CH * (i % NB); k_mix<<<592, 256>>>((const uint4*)(buf + CH * ((i + 3) % NB)), (int4*)b, (uint4*)(b + npx * 4), npx / 16); }, 24), npx * 6.0);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0; }
