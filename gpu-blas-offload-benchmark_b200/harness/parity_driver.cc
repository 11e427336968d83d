// parity_driver.cc -- drives the B200 backend CLASSES (B200/gemm.hh,
// B200/gemv.hh) exactly the way include/doGemm.hh:326-340 and
// include/doGemv.hh:215-229 do -- initialise(mode, ...) then compute() -- and
// writes the full result matrix / vector to a file so tests can compare it
// element-wise with the oracle (the harness itself only ever looks at a 4- or
// 2-element checksum, include/kernels/gemm.hh:61-65).
//
// The reference frees the operands inside compute() (kernels/gemm.hh:39), so
// the result is captured by a subclass whose postCallKernelCleanup() copies
// C / y out.  Nothing else differs from what the harness executes.
//
//   b200_parity <sgemm|dgemm> <once|always|unified> M N K iters out.bin
//   b200_parity <sgemv|dgemv> <once|always|unified> M N iters out.bin
// prints: checksum seconds
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <type_traits>
#include <vector>

#include "B200/gemm.hh"
#include "B200/gemv.hh"

extern "C" int consume(void*, void*, void*) { return 0; }

namespace {

gpuOffloadType parse_mode(const std::string& s) {
  if (s == "once") return gpuOffloadType::once;
  if (s == "always") return gpuOffloadType::always;
  if (s == "unified") return gpuOffloadType::unified;
  std::cerr << "unknown offload mode " << s << std::endl;
  exit(2);
}

}  // namespace

// The capture has to happen between the end of the timed region and the
// cleanup, both of which are inside compute().  A subclass can override the
// private virtual postCallKernelCleanup() but cannot call the backend's own
// (private) implementation, so the operands of this one-shot process are
// simply left to process exit.
template <typename T>
class leaky_gemm : public gpu::gemm_gpu<T> {
 public:
  leaky_gemm(int iters, std::vector<T>* out) : gpu::gemm_gpu<T>(iters), out_(out) {}

 private:
  void postCallKernelCleanup() override {
    out_->assign(this->C_, this->C_ + (size_t)this->m_ * this->n_);
  }
  std::vector<T>* out_;
};

template <typename T>
class leaky_gemv : public gpu::gemv_gpu<T> {
 public:
  leaky_gemv(int iters, std::vector<T>* out) : gpu::gemv_gpu<T>(iters), out_(out) {}

 private:
  void postCallKernelCleanup() override { out_->assign(this->y_, this->y_ + (size_t)this->m_); }
  std::vector<T>* out_;
};

template <typename T>
int run_gemm(gpuOffloadType mode, int m, int n, int k, int iters, const char* path) {
  std::vector<T> out;
  leaky_gemm<T> g(iters, &out);
  g.initialise(mode, m, n, k);
  time_checksum_gflop r = g.compute();
  FILE* f = fopen(path, "wb");
  if (!f) return 3;
  fwrite(out.data(), sizeof(T), out.size(), f);
  fclose(f);
  printf("%.17g %.9f\n", r.checksum, r.runtime);
  return 0;
}

template <typename T>
int run_gemv(gpuOffloadType mode, int m, int n, int iters, const char* path) {
  std::vector<T> out;
  leaky_gemv<T> g(iters, &out);
  g.initialise(mode, m, n);
  time_checksum_gflop r = g.compute();
  FILE* f = fopen(path, "wb");
  if (!f) return 3;
  fwrite(out.data(), sizeof(T), out.size(), f);
  fclose(f);
  printf("%.17g %.9f\n", r.checksum, r.runtime);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 7) {
    std::cerr << "usage: b200_parity <sgemm|dgemm> <mode> M N K iters out.bin\n"
                 "       b200_parity <sgemv|dgemv> <mode> M N iters out.bin\n";
    return 2;
  }
  const std::string kern = argv[1];
  const gpuOffloadType mode = parse_mode(argv[2]);
  if (kern == "sgemm" || kern == "dgemm") {
    if (argc < 8) return 2;
    const int m = atoi(argv[3]), n = atoi(argv[4]), k = atoi(argv[5]), it = atoi(argv[6]);
    return kern == "sgemm" ? run_gemm<float>(mode, m, n, k, it, argv[7])
                           : run_gemm<double>(mode, m, n, k, it, argv[7]);
  }
  const int m = atoi(argv[3]), n = atoi(argv[4]), it = atoi(argv[5]);
  return kern == "sgemv" ? run_gemv<float>(mode, m, n, it, argv[6])
                         : run_gemv<double>(mode, m, n, it, argv[6]);
}
