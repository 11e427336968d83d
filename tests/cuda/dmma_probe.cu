// dmma_probe.cu -- how fast can DMMA.8x8x4 issue as a function of warps per SM,
// independent accumulators per warp, and operand-register variety?  Answers what
// the DGEMM kernel's occupancy can reach before any memory traffic.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
constexpr int ITERS = 2048;

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// MI x NI accumulator tiles, MI distinct A regs, NI distinct B regs (like a warp tile)
template <int MI, int NI>
__global__ void k_tile(double* out, double seed) {
  double c[MI][NI][2], a[MI], b[NI];
#pragma unroll
  for (int i = 0; i < MI; i++) { a[i] = seed + i * 1e-9 + threadIdx.x * 1e-12;
#pragma unroll
    for (int j = 0; j < NI; j++) c[i][j][0] = c[i][j][1] = seed; }
#pragma unroll
  for (int j = 0; j < NI; j++) b[j] = 1e-9 * (j + 1);
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
      for (int j = 0; j < NI; j++) dmma(c[i][j][0], c[i][j][1], a[i], b[j]);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < MI; i++)
#pragma unroll
    for (int j = 0; j < NI; j++) s += c[i][j][0] + c[i][j][1];
  if (s == 123.456) out[0] = s;
}

template <int MI, int NI>
void run(const char* name, int sms, int warps_per_sm, double* buf) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int threads = warps_per_sm * 32;  // one CTA per SM
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(e0));
    k_tile<MI, NI><<<sms, threads>>>(buf, 1.0);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep && ms < best) best = ms;
  }
  const double flops = 2.0 * 256 * MI * NI * (double)ITERS * sms * warps_per_sm;
  printf("%-10s warps/SM=%2d  %8.2f TFLOP/s\n", name, warps_per_sm, flops / (best * 1e-3) * 1e-12);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  double* buf; CK(cudaMalloc(&buf, 1024));
  const int sms = p.multiProcessorCount;
  for (int w : {4, 8, 12, 16, 32}) {
    run<8, 4>("tile 8x4", sms, w, buf);
    run<4, 4>("tile 4x4", sms, w, buf);
    run<4, 2>("tile 4x2", sms, w, buf);
    run<2, 2>("tile 2x2", sms, w, buf);
  }
  return 0;
}
