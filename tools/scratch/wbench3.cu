// Tile-shaped store patterns: how much of the write bandwidth survives the 128x32 tiling of K1?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <functional>
// coalesced mix: per 16 px: read 16 B, write 16 B mask, 64 B labels with warp-contiguous 512 B label stores
__global__ void k_mix_lin(const uint4* in, int4* lab, uint4* mask, size_t npx16) {
  const int4 z = make_int4(0,0,0,0); const int lane = threadIdx.x & 31;
  for (size_t w0 = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) / 32 * 32; w0 < npx16; w0 += (size_t)gridDim.x * blockDim.x) {
    uint4 v = in[w0 + lane]; mask[w0 + lane] = v;
    int4* l = lab + 4 * w0;  // 128 int4 for this warp's 32 groups
#pragma unroll
    for (int k = 0; k < 4; k++) l[32 * k + lane] = z; } }
// tile pattern: CTA handles tiles of TWxTH px of an n x H x W batch, tile index grid-strided (like K1 static schedule)
template <int TW, int TH>
__global__ void k_tile(const uint8_t* in, int32_t* lab, uint8_t* mask, int n, int H, int W, int do_read) {
  const int tx_n = W / TW, ty_n = H / TH, per = tx_n * ty_n, total = per * n; const int tid = threadIdx.x; const int4 z = make_int4(0,0,0,0);
  for (int t = blockIdx.x; t < total; t += gridDim.x) {
    const int f = t / per, r = t - f * per, ty = r / tx_n, tx = r - ty * tx_n; const size_t o0 = (size_t)f * H * W + (size_t)ty * TH * W + tx * TW;
    uint4 keepv[2] = {make_uint4(0,0,0,0), make_uint4(0,0,0,0)};
    if (do_read) { // read the tile as 16-byte pieces
      for (int i = tid; i < TW * TH / 16; i += 256) { const int rr = i / (TW / 16), c = i - rr * (TW / 16); keepv[(i / 256) & 1] = *reinterpret_cast<const uint4*>(in + o0 + (size_t)rr * W + 16 * c); } }
    for (int i = tid; i < TW * TH / 4; i += 256) { const int rr = i / (TW / 4), c = i - rr * (TW / 4); *reinterpret_cast<int4*>(lab + o0 + (size_t)rr * W + 4 * c) = z; }
    for (int i = tid; i < TW * TH / 16; i += 256) { const int rr = i / (TW / 16), c = i - rr * (TW / 16); *reinterpret_cast<uint4*>(mask + o0 + (size_t)rr * W + 16 * c) = keepv[(i / 256) & 1]; }
  } }
float timeit(std::function<void(int)> f, int reps) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); for (int i = 0; i < 3; i++) f(i); cudaEventRecord(a); for (int i = 0; i < reps; i++) f(i + 3); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps; }
int main() {
  const int n = 25, H = 1024, W = 1280; const size_t npx = (size_t)n * H * W; const int NB = 8;
  uint8_t *in, *mask; int32_t* lab; cudaMalloc(&in, npx * NB); cudaMalloc(&mask, npx * NB); cudaMalloc(&lab, npx * 4 * NB); cudaMemset(in, 1, npx * NB);
  auto P = [&](const char* name, float ms, double bytes) { printf("%-50s %7.1f us  %6.0f GB/s\n", name, ms * 1e3, bytes / ms / 1e6); };
  for (int g : {592, 1184, 2368}) { char nm[80]; snprintf(nm, 80, "linear coalesced mix 1R:5W grid=%d", g);
    P(nm, timeit([&](int i) { int s = i % NB; k_mix_lin<<<g, 256>>>((const uint4*)(in + npx * s), (int4*)(lab + npx * s), (uint4*)(mask + npx * s), npx / 16); }, 24), npx * 6.0); }
  for (int rd = 0; rd < 2; rd++) for (int g : {592, 1184, 2368}) { char nm[80];
    snprintf(nm, 80, "tile 128x32 %s grid=%d", rd ? "R+W" : "W only", g); P(nm, timeit([&](int i) { int s = i % NB; k_tile<128, 32><<<g, 256>>>(in + npx * s, lab + npx * s, mask + npx * s, n, H, W, rd); }, 24), npx * (5.0 + rd));
    snprintf(nm, 80, "tile 256x16 %s grid=%d", rd ? "R+W" : "W only", g); P(nm, timeit([&](int i) { int s = i % NB; k_tile<256, 16><<<g, 256>>>(in + npx * s, lab + npx * s, mask + npx * s, n, H, W, rd); }, 24), npx * (5.0 + rd));
    snprintf(nm, 80, "tile 1280x4 %s grid=%d", rd ? "R+W" : "W only", g); P(nm, timeit([&](int i) { int s = i % NB; k_tile<1280, 4><<<g, 256>>>(in + npx * s, lab + npx * s, mask + npx * s, n, H, W, rd); }, 24), npx * (5.0 + rd));
    snprintf(nm, 80, "tile 64x64 %s grid=%d", rd ? "R+W" : "W only", g); P(nm, timeit([&](int i) { int s = i % NB; k_tile<64, 64><<<g, 256>>>(in + npx * s, lab + npx * s, mask + npx * s, n, H, W, rd); }, 24), npx * (5.0 + rd)); }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0; }
