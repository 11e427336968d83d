// ref_driver.cc -- thin C entry points around the REFERENCE's own CPU classes.
//
// TEST INFRASTRUCTURE ONLY (see oracle/blob_oracle.c header).  Built by
// oracle/Makefile into oracle/_ref/libblobref.so from the sources where they
// lie under /root/reference (nothing is copied into this repo):
//   cpu::gemm_cpu<T>  /root/reference/OpenBLAS/gemm.hh:14-60
//   cpu::gemv_cpu<T>  /root/reference/OpenBLAS/gemv.hh:14-56
//   ::gemm<T>::compute / initInputMatrices / calcChecksum
//                     /root/reference/include/kernels/gemm.hh:19-42,69-87,61-65
//   ::gemv<T> twins   /root/reference/include/kernels/gemv.hh:19-42,69-83,61-65
// The reference frees C / y inside compute() (kernels/gemm.hh:39), so the only
// addition here is a subclass whose postCallKernelCleanup() copies the result
// out before freeing -- the init, the cblas call, the timer and the checksum
// are the reference's, unmodified.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <type_traits>

#include <cblas.h>

#include "OpenBLAS/gemm.hh"
#include "OpenBLAS/gemv.hh"

namespace {

template <typename T>
class capturing_gemm : public cpu::gemm_cpu<T> {
 public:
  capturing_gemm(int iters, T* out) : cpu::gemm_cpu<T>(iters), out_(out) {}

 private:
  void postCallKernelCleanup() override {
    if (out_)
      std::memcpy(out_, this->C_, sizeof(T) * (size_t)this->m_ * this->n_);
    free(this->A_);
    free(this->B_);
    free(this->C_);
  }
  T* out_;
};

template <typename T>
class capturing_gemv : public cpu::gemv_cpu<T> {
 public:
  capturing_gemv(int iters, T* out) : cpu::gemv_cpu<T>(iters), out_(out) {}

 private:
  void postCallKernelCleanup() override {
    if (out_) std::memcpy(out_, this->y_, sizeof(T) * (size_t)this->m_);
    free(this->A_);
    free(this->x_);
    free(this->y_);
  }
  T* out_;
};

template <typename T>
int run_gemm(int m, int n, int k, int iters, T* c_out, double* checksum,
             double* seconds) {
  capturing_gemm<T> g(iters, c_out);
  g.initialise(m, n, k);
  time_checksum_gflop r = g.compute();
  if (checksum) *checksum = r.checksum;
  if (seconds) *seconds = r.runtime;
  return 0;
}

template <typename T>
int run_gemv(int m, int n, int iters, T* y_out, double* checksum,
             double* seconds) {
  capturing_gemv<T> g(iters, y_out);
  g.initialise(m, n);
  time_checksum_gflop r = g.compute();
  if (checksum) *checksum = r.checksum;
  if (seconds) *seconds = r.runtime;
  return 0;
}

}  // namespace

extern "C" {

int blobref_dgemm(int m, int n, int k, int iters, double* c_out,
                  double* checksum, double* seconds) {
  return run_gemm<double>(m, n, k, iters, c_out, checksum, seconds);
}
int blobref_sgemm(int m, int n, int k, int iters, float* c_out,
                  double* checksum, double* seconds) {
  return run_gemm<float>(m, n, k, iters, c_out, checksum, seconds);
}
int blobref_dgemv(int m, int n, int iters, double* y_out, double* checksum,
                  double* seconds) {
  return run_gemv<double>(m, n, iters, y_out, checksum, seconds);
}
int blobref_sgemv(int m, int n, int iters, float* y_out, double* checksum,
                  double* seconds) {
  return run_gemv<float>(m, n, iters, y_out, checksum, seconds);
}

/* Direct cblas calls on caller-supplied buffers (for transposed / beta != 0
 * cases the harness never exercises). trans: 0=N, 1=T. */
void blobref_cblas_dgemm(int ta, int tb, int m, int n, int k, double alpha,
                         const double* A, int lda, const double* B, int ldb,
                         double beta, double* C, int ldc) {
  cblas_dgemm(CblasColMajor, ta ? CblasTrans : CblasNoTrans,
              tb ? CblasTrans : CblasNoTrans, m, n, k, alpha, A, lda, B, ldb,
              beta, C, ldc);
}
void blobref_cblas_sgemm(int ta, int tb, int m, int n, int k, float alpha,
                         const float* A, int lda, const float* B, int ldb,
                         float beta, float* C, int ldc) {
  cblas_sgemm(CblasColMajor, ta ? CblasTrans : CblasNoTrans,
              tb ? CblasTrans : CblasNoTrans, m, n, k, alpha, A, lda, B, ldb,
              beta, C, ldc);
}
void blobref_cblas_dgemv(int trans, int m, int n, double alpha, const double* A,
                         int lda, const double* x, int incx, double beta,
                         double* y, int incy) {
  cblas_dgemv(CblasColMajor, trans ? CblasTrans : CblasNoTrans, m, n, alpha, A,
              lda, x, incx, beta, y, incy);
}
void blobref_cblas_sgemv(int trans, int m, int n, float alpha, const float* A,
                         int lda, const float* x, int incx, float beta, float* y,
                         int incy) {
  cblas_sgemv(CblasColMajor, trans ? CblasTrans : CblasNoTrans, m, n, alpha, A,
              lda, x, incx, beta, y, incy);
}

const char* blobref_blas_config(void) { return openblas_get_config(); }
int blobref_blas_threads(void) { return openblas_get_num_threads(); }
void blobref_blas_set_threads(int n) { openblas_set_num_threads(n); }

}  // extern "C"
