// store_probe.cu -- write-only bandwidth of the store patterns the GEMM epilogues use
// (why do the K = 32 shapes stop at ~2 TB/s of C writes?).  A measurement tool only.
//   nvcc -O2 -arch=sm_100a store_probe.cu -o store_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

// (a) linear: every thread writes float4s, grid-stride
__global__ void fill_linear(float4* p, size_t n4) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride)
    p[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}
// (b) column-major tiles, the tcgen05 epilogue's pattern: CTA = ROWS x COLS tile, warp w owns rows
// [32w, 32w+32), lane = row, one 4-byte store per column (a warp store = 128 contiguous bytes)
template <int ROWS, int COLS>
__global__ void fill_tiles(float* c, int m, int n, int tiles_m, int tiles_n, int persistent) {
  const int ntiles = tiles_m * tiles_n;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int group = 8 * tiles_n, first = (tile / group) * 8, gm = min(tiles_m - first, 8);
    const int m0 = (first + (tile % group) % gm) * ROWS, n0 = ((tile % group) / gm) * COLS;
    for (int r = threadIdx.x; r < ROWS; r += blockDim.x) {
      float* cp = c + (size_t)n0 * m + m0 + r;
#pragma unroll 16
      for (int j = 0; j < COLS; j++) cp[(size_t)j * m] = (float)j;
    }
    if (!persistent) break;
  }
}
// (c) same tiles, 16-byte stores: lane = 4 consecutive rows
template <int ROWS, int COLS>
__global__ void fill_tiles_v4(float* c, int m, int n, int tiles_m, int tiles_n) {
  const int ntiles = tiles_m * tiles_n;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int group = 8 * tiles_n, first = (tile / group) * 8, gm = min(tiles_m - first, 8);
    const int m0 = (first + (tile % group) % gm) * ROWS, n0 = ((tile % group) / gm) * COLS;
    constexpr int RG = ROWS / 4;
    for (int e = threadIdx.x; e < RG * COLS; e += blockDim.x) {
      const int rg = e % RG, j = e / RG;
      *reinterpret_cast<float4*>(c + (size_t)(n0 + j) * m + m0 + rg * 4) = make_float4(1.f, 2.f, 3.f, 4.f);
    }
  }
}
template <typename F> float time_ms(F f, int iters) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; i++) f();
  CK(cudaEventRecord(a)); for (int i = 0; i < iters; i++) f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b)); CK(cudaGetLastError()); return ms / iters;
}
int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 8192;  // n x n floats
  const size_t bytes = (size_t)n * n * 4;
  float* c; CK(cudaMalloc(&c, bytes));
  float* c2; CK(cudaMalloc(&c2, bytes));
  auto rep = [&](const char* what, float ms) { printf("%-58s %8.1f us  %7.1f GB/s\n", what, ms * 1e3, bytes / ms * 1e-6); };
  rep("memset", time_ms([&] { cudaMemsetAsync(c, 0, bytes); }, 20));
  rep("linear float4, 148x8 CTAs x 256", time_ms([&] { fill_linear<<<148 * 8, 256>>>((float4*)c, bytes / 16); }, 20));
  rep("linear float4, 148x32 CTAs x 256", time_ms([&] { fill_linear<<<148 * 32, 256>>>((float4*)c, bytes / 16); }, 20));
  rep("memcpy d2d (read + write counted once)", time_ms([&] { cudaMemcpyAsync(c2, c, bytes, cudaMemcpyDeviceToDevice); }, 20));
  const int t128 = n / 128, t256 = n / 256;
  rep("tiles 128x128, 128 thr, 4B/lane, persistent 148", time_ms([&] { fill_tiles<128, 128><<<148, 128>>>(c, n, n, t128, t128, 1); }, 20));
  rep("tiles 128x128, 128 thr, 4B/lane, one CTA per tile", time_ms([&] { fill_tiles<128, 128><<<t128 * t128, 128>>>(c, n, n, t128, t128, 0); }, 20));
  rep("tiles 128x256, 256 thr (2 col halves), persistent 148", time_ms([&] { fill_tiles<128, 128><<<148, 128>>>(c, n, n, t128, t128, 1); }, 20));
  rep("tiles 256x128, 256 thr, 4B/lane, persistent 148", time_ms([&] { fill_tiles<256, 128><<<148, 256>>>(c, n, n, t256, t128, 1); }, 20));
  rep("tiles 256x128, 256 thr, 4B/lane, persistent 296", time_ms([&] { fill_tiles<256, 128><<<296, 256>>>(c, n, n, t256, t128, 1); }, 20));
  rep("tiles 128x128, 256 thr, 16B/lane, persistent 148", time_ms([&] { fill_tiles_v4<128, 128><<<148, 256>>>(c, n, n, t128, t128); }, 20));
  rep("tiles 128x128, 256 thr, 16B/lane, persistent 592", time_ms([&] { fill_tiles_v4<128, 128><<<592, 256>>>(c, n, n, t128, t128); }, 20));
  rep("tiles 256x64, 256 thr, 16B/lane, persistent 592", time_ms([&] { fill_tiles_v4<256, 64><<<592, 256>>>(c, n, n, t256, n / 64); }, 20));
  rep("tiles 1024x32, 256 thr, 16B/lane, persistent 592", time_ms([&] { fill_tiles_v4<1024, 32><<<592, 256>>>(c, n, n, n / 1024, n / 32); }, 20));
  return 0;
}
