// cublas_probe.cu -- kernel-only baselines of the calls the REFERENCE cuBLAS backend makes
// (cuBLAS/gemm.hh:53,134-146; cuBLAS/gemv.hh:130-137), with CUDA events, plus the
// element-wise error of its SGEMM under CUBLAS_TENSOR_OP_MATH.
// A measurement tool only: nothing in the product links cuBLAS.
//   nvcc -O2 -arch=sm_100a cublas_probe.cu -lcublas -o cublas_probe
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
#define CB(x) do { cublasStatus_t e = (x); if (e != CUBLAS_STATUS_SUCCESS) { printf("cuBLAS %s: %d\n", #x, (int)e); exit(1);} } while (0)

__global__ void fill(double* a, float* af, size_t n, double div, unsigned seed) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned x = (unsigned)i * 2654435761u + seed; x ^= x >> 16; x *= 2246822519u; x ^= x >> 13;
  const double v = (double)(x % 100) / div;   // the reference's value distribution
  a[i] = (double)(float)v; af[i] = (float)v;
}
template <typename F> float time_ms(F f, int iters) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; i++) f();
  CK(cudaEventRecord(a)); for (int i = 0; i < iters; i++) f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms / iters;
}
int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 8192;
  const size_t nn = (size_t)n * n;
  double *A, *B, *C; float *Af, *Bf, *Cf; double* err;
  CK(cudaMalloc(&A, nn * 8)); CK(cudaMalloc(&B, nn * 8)); CK(cudaMalloc(&C, nn * 8));
  CK(cudaMalloc(&Af, nn * 4)); CK(cudaMalloc(&Bf, nn * 4)); CK(cudaMalloc(&Cf, nn * 4)); CK(cudaMalloc(&err, 8));
  fill<<<(nn + 255) / 256, 256>>>(A, Af, nn, 7.0, 1); fill<<<(nn + 255) / 256, 256>>>(B, Bf, nn, 3.0, 2);
  cublasHandle_t h; CB(cublasCreate(&h));
  const double one = 1.0, zero = 0.0; const float onef = 1.f, zerof = 0.f;
  const double gflop = (2.0 * n * n * (double)n + (double)n * n) * 1e-9;
  float ms = time_ms([&] { CB(cublasGemmEx(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, CUDA_R_64F, n, B, CUDA_R_64F, n, &zero, C, CUDA_R_64F, n, CUBLAS_COMPUTE_64F, CUBLAS_GEMM_DEFAULT)); }, 5);
  printf("cuBLAS DGEMM %d^3 (COMPUTE_64F): %.3f ms  %.1f GFLOP/s\n", n, ms, gflop / ms * 1e3);
  // C (FP64 of the float-rounded inputs) is the reference for the SGEMM error
  for (int mode = 0; mode < 2; mode++) {
    CB(cublasSetMathMode(h, mode == 0 ? CUBLAS_TENSOR_OP_MATH : CUBLAS_DEFAULT_MATH));
    ms = time_ms([&] { CB(cublasGemmEx(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &onef, Af, CUDA_R_32F, n, Bf, CUDA_R_32F, n, &zerof, Cf, CUDA_R_32F, n, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT)); }, 5);
    // element-wise error on a 4M-element sample (host side)
    const size_t ns = nn < (4u << 20) ? nn : (4u << 20);
    std::vector<float> hc(ns); std::vector<double> hr(ns);
    CK(cudaMemcpy(hc.data(), Cf + (nn - ns) / 2, ns * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hr.data(), C + (nn - ns) / 2, ns * 8, cudaMemcpyDeviceToHost));
    double e = 0;
    for (size_t i = 0; i < ns; i++) e = fmax(e, fabs((double)hc[i] - hr[i]) / fabs(hr[i]));
    printf("cuBLAS SGEMM %d^3 (%s, as the reference backend %s): %.3f ms  %.1f GFLOP/s  max rel err vs FP64 = %.3e\n", n,
           mode == 0 ? "CUBLAS_TENSOR_OP_MATH" : "CUBLAS_DEFAULT_MATH", mode == 0 ? "sets it" : "would without it", ms, gflop / ms * 1e3, e);
  }
  CB(cublasSetMathMode(h, CUBLAS_DEFAULT_MATH));
  // GEMV on the same n x n operands
  ms = time_ms([&] { CB(cublasDgemv(h, CUBLAS_OP_N, n, n, &one, A, n, B, 1, &zero, C, 1)); }, 20);
  printf("cuBLAS DGEMV %d^2: %.4f ms  %.1f GB/s\n", n, ms, (nn + 2.0 * n) * 8 / ms * 1e-6);
  ms = time_ms([&] { CB(cublasSgemv(h, CUBLAS_OP_N, n, n, &onef, Af, n, Bf, 1, &zerof, Cf, 1)); }, 20);
  printf("cuBLAS SGEMV %d^2: %.4f ms  %.1f GB/s\n", n, ms, (nn + 2.0 * n) * 4 / ms * 1e-6);
  return 0;
}
