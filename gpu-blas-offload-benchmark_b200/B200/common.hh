#pragma once
// B200/common.hh -- shared plumbing of the GPU_LIB=B200 backend for GPU-BLOB.
//
// The backend's host side is plain C++17 (no CUDA headers): all device work
// goes through the extern "C" ABI of libb200blas.so (b200blas.h).  The error
// convention follows the reference's cudaCheckError / cublasCheckError
// (cuBLAS/common.hh:6-23): print "<what>: file:line: message" to stdout and
// exit(1).
#if defined GPU_B200

#include <cstdlib>
#include <iostream>
#include <string>

#include <b200blas.h>

#define b200CheckError(call)                                                  \
  do {                                                                        \
    if (int b200_rc_ = (call); b200_rc_ != 0) {                               \
      std::cout << "B200 error: " << __FILE__ << ":" << __LINE__ << ": "      \
                << b200_last_error() << std::endl;                            \
      exit(1);                                                                \
    }                                                                         \
  } while (false)

namespace gpu {
namespace b200detail {

/** One operand of a kernel: the host-visible pointer the harness reads and
 * writes (pinned, or managed in unified mode) and its device twin (the same
 * address in unified mode). */
template <typename T>
struct operand {
  T* host = nullptr;
  T* dev = nullptr;
  size_t bytes = 0;

  void acquire(b200_ctx* ctx, int mode, size_t count) {
    bytes = sizeof(T) * count;
    void *h = nullptr, *d = nullptr;
    b200CheckError(b200_alloc(ctx, mode, bytes, &h, &d));
    host = static_cast<T*>(h);
    dev = static_cast<T*>(d);
  }
  void release(b200_ctx* ctx, int mode) {
    b200CheckError(b200_free(ctx, mode, host, dev));
    host = dev = nullptr;
    bytes = 0;
  }
};

/** Lazily created, per-backend-object context (the harness keeps four backend
 * objects alive at once, src/main.cc:35-59). */
class context_holder {
 public:
  ~context_holder() {
    if (ctx_ != nullptr) b200CheckError(b200_ctx_destroy(ctx_));
  }
  b200_ctx* get() {
    if (ctx_ == nullptr) b200CheckError(b200_ctx_create(&ctx_, -1));
    return ctx_;
  }

 private:
  b200_ctx* ctx_ = nullptr;
};

/** Number of GPUs one problem is partitioned over: environment variable B200_NGPUS
 * (default 1).  The reference has no multi-GPU notion; this is read where it reads its
 * single device (cuBLAS/gemm.hh:45-63, cuBLAS/gemv.hh:45-60). */
inline int ngpus() {
  static const int n = [] {
    const char* e = std::getenv("B200_NGPUS");
    const int v = e ? std::atoi(e) : 1;
    return v > 1 ? v : 1;
  }();
  return n;
}

/** B200_UPLOAD_C=1: upload C / y in Once and Always mode even though beta == 0 makes it dead
 * traffic, exactly as the reference does (cuBLAS/gemm.hh:104-105,130-131,
 * cuBLAS/gemv.hh:100-101,126-127) -- for like-for-like transfer volumes against the cuBLAS
 * backend.  Default: skipped.  (The library reads the same variable for Always mode.) */
inline bool upload_c() {
  static const bool v = [] {
    const char* e = std::getenv("B200_UPLOAD_C");
    return e != nullptr && std::string(e) != "0" && std::string(e) != "off";
  }();
  return v;
}

/** Lazily created team of `ngpus()` devices (one per backend object, like the context). */
class team_holder {
 public:
  ~team_holder() {
    if (mg_ != nullptr) b200CheckError(b200_mg_destroy(mg_));
  }
  b200_mg* get() {
    if (mg_ == nullptr) b200CheckError(b200_mg_create(&mg_, ngpus()));
    return mg_;
  }

 private:
  b200_mg* mg_ = nullptr;
};

}  // namespace b200detail
}  // namespace gpu

#endif
