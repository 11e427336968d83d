// ref_cublas_driver.cc -- command-line front end of the REFERENCE's own cuBLAS backend classes.
//
// MEASUREMENT / TEST INFRASTRUCTURE ONLY (see oracle/blob_oracle.c header): this is the second
// reported baseline the north star asks for ("the reference's cuBLAS backend"), never a product
// path.  Built by oracle/Makefile into oracle/_ref/blobref-cublas from the sources where they
// lie under /root/reference (nothing is copied into this repo):
//   gpu::gemm_gpu<T>  /root/reference/cuBLAS/gemm.hh:14-274
//   gpu::gemv_gpu<T>  /root/reference/cuBLAS/gemv.hh:14-254
//   ::gemm<T>::compute (the harness timer)  /root/reference/include/kernels/gemm.hh:19-42
// driven exactly as include/doGemm.hh:326-340 drives them: initialise(mode, ...) then compute().
// It runs as its own process so that libcublas is never loaded into a process that also
// holds the product library.
//
//   blobref-cublas <sgemm|dgemm> <once|always|unified> M N K iters [repeats]
//   blobref-cublas <sgemv|dgemv> <once|always|unified> M N iters [repeats]
// prints one line per repeat: checksum seconds gflops   (seconds = the harness' own timer over
// `iters` iterations; gflops with the reference's FLOP formula, beta = 0)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#define GPU_CUBLAS
#include "cuBLAS/gemm.hh"
#include "cuBLAS/gemv.hh"

namespace {
gpuOffloadType parse_mode(const std::string& s) {
  if (s == "once") return gpuOffloadType::once;
  if (s == "always") return gpuOffloadType::always;
  if (s == "unified") return gpuOffloadType::unified;
  std::cerr << "unknown offload mode " << s << std::endl;
  exit(2);
}

template <typename T>
int run_gemm(gpuOffloadType mode, int m, int n, int k, int iters, int repeats) {
  gpu::gemm_gpu<T> g(iters);
  for (int r = 0; r < repeats; r++) {
    g.initialise(mode, m, n, k);
    time_checksum_gflop res = g.compute();
    const double flops = 2.0 * m * n * (double)k + (double)m * n;  // include/doGemm.hh:493-505, BETA == 0
    printf("%.17g %.9f %.3f\n", res.checksum, res.runtime, flops * iters / res.runtime * 1e-9);
    fflush(stdout);
  }
  return 0;
}

template <typename T>
int run_gemv(gpuOffloadType mode, int m, int n, int iters, int repeats) {
  gpu::gemv_gpu<T> g(iters);
  for (int r = 0; r < repeats; r++) {
    g.initialise(mode, m, n);
    time_checksum_gflop res = g.compute();
    const double flops = 2.0 * m * (double)n + (double)m;  // include/doGemv.hh:377-388, BETA == 0
    printf("%.17g %.9f %.3f\n", res.checksum, res.runtime, flops * iters / res.runtime * 1e-9);
    fflush(stdout);
  }
  return 0;
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 6) {
    std::cerr << "usage: blobref-cublas <sgemm|dgemm|sgemv|dgemv> <once|always|unified> M N [K] iters [repeats]\n";
    return 2;
  }
  const std::string kern = argv[1];
  const gpuOffloadType mode = parse_mode(argv[2]);
  if (kern == "sgemm" || kern == "dgemm") {
    if (argc < 7) return 2;
    const int m = atoi(argv[3]), n = atoi(argv[4]), k = atoi(argv[5]), iters = atoi(argv[6]);
    const int repeats = argc > 7 ? atoi(argv[7]) : 1;
    return kern == "dgemm" ? run_gemm<double>(mode, m, n, k, iters, repeats) : run_gemm<float>(mode, m, n, k, iters, repeats);
  }
  const int m = atoi(argv[3]), n = atoi(argv[4]), iters = atoi(argv[5]);
  const int repeats = argc > 6 ? atoi(argv[6]) : 1;
  return kern == "dgemv" ? run_gemv<double>(mode, m, n, iters, repeats) : run_gemv<float>(mode, m, n, iters, repeats);
}
