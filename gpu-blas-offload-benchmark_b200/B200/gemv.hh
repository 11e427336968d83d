#pragma once
// B200/gemv.hh -- gpu::gemv_gpu<T> for GPU_LIB=B200.
//
// Same plugin slot as the reference's cuBLAS/gemv.hh:14-254 (virtuals declared
// at include/kernels/gemv.hh:44-58 and include/kernels/GPU/gemv.hh:20), driven
// unchanged by include/doGemv.hh.  See B200/gemm.hh for the mode table; the
// operands here are A (m x n), x (n) and y (m).
#ifdef GPU_B200

#include <algorithm>
#include <type_traits>

#include "../include/kernels/GPU/gemv.hh"
#include "../include/utilities.hh"
#include "common.hh"

namespace gpu {

template <typename T>
class gemv_gpu : public gemv<T> {
 public:
  using gemv<T>::gemv;
  using gemv<T>::initInputMatrixVector;
  using gemv<T>::m_;
  using gemv<T>::n_;
  using gemv<T>::A_;
  using gemv<T>::x_;
  using gemv<T>::y_;
  using gemv<T>::offload_;
  using gemv<T>::vecIncrement_;

  void initialise(gpuOffloadType offload, int m, int n) override {
    offload_ = offload;
    m_ = m;
    n_ = n;
    const int mode = static_cast<int>(offload_);
    if (b200detail::ngpus() > 1) {  // row blocks of A and y over a team of GPUs, x replicated (csrc/mg.cu)
      void *h0 = nullptr, *h1 = nullptr, *h2 = nullptr;
      b200CheckError(b200_mg_problem_create(team_.get(), B200_MG_GEMV, (int)sizeof(T), mode, m_, n_, 0,
                                            &h0, &h1, &h2, &problem_));
      A_ = static_cast<T*>(h0);
      x_ = static_cast<T*>(h1);
      y_ = static_cast<T*>(h2);
      initInputMatrixVector();
      return;
    }
    b200_ctx* ctx = ctx_.get();
    a_.acquire(ctx, mode, (size_t)m_ * n_);
    x_op_.acquire(ctx, mode, (size_t)n_);
    y_op_.acquire(ctx, mode, (size_t)m_);
    A_ = a_.host;
    x_ = x_op_.host;
    y_ = y_op_.host;
    if (offload_ == gpuOffloadType::unified) {
      b200CheckError(b200_advise(ctx, a_.dev, a_.bytes, 1));
      b200CheckError(b200_advise(ctx, x_op_.dev, x_op_.bytes, 1));
      b200CheckError(b200_advise(ctx, y_op_.dev, y_op_.bytes, 0));
    }
    initInputMatrixVector();
  }

 private:
  void preLoopRequirements() override {
    if (problem_ != nullptr) {
      b200CheckError(b200_mg_problem_preloop(problem_, (double)beta));
      return;
    }
    b200_ctx* ctx = ctx_.get();
    if (offload_ == gpuOffloadType::once) {
      b200CheckError(b200_h2d_async(ctx, a_.dev, a_.host, a_.bytes, 0));
      b200CheckError(b200_h2d_async(ctx, x_op_.dev, x_op_.host, x_op_.bytes, 1));
      // BLAS never reads the output operand when beta == 0, so the reference's
      // third upload (cuBLAS/gemm.hh:104-105, cuBLAS/gemv.hh:100-101) is skipped
      // (B200_UPLOAD_C=1 restores it)
      if (beta != T(0) || b200detail::upload_c()) b200CheckError(b200_h2d_async(ctx, y_op_.dev, y_op_.host, y_op_.bytes, 2));
    } else if (offload_ == gpuOffloadType::unified) {
      b200CheckError(b200_prefetch_async(ctx, a_.dev, a_.bytes, 1, 0));
      b200CheckError(b200_prefetch_async(ctx, x_op_.dev, x_op_.bytes, 1, 1));
      b200CheckError(b200_prefetch_async(ctx, y_op_.dev, y_op_.bytes, 1, 2));
    }
  }

  void callGemv() override {
    if (problem_ != nullptr) {
      b200CheckError(b200_mg_problem_call(problem_, (double)alpha, (double)beta));
      return;
    }
    b200_ctx* ctx = ctx_.get();
    const int lda = std::max(1, m_);
    if (offload_ == gpuOffloadType::always) {
      if constexpr (std::is_same_v<T, float>)
        b200CheckError(b200_sgemv_offload(ctx, m_, n_, alpha, a_.host, a_.dev, lda, x_op_.host,
                                          x_op_.dev, beta, y_op_.host, y_op_.dev));
      else
        b200CheckError(b200_dgemv_offload(ctx, m_, n_, alpha, a_.host, a_.dev, lda, x_op_.host,
                                          x_op_.dev, beta, y_op_.host, y_op_.dev));
    } else {
      if constexpr (std::is_same_v<T, float>)
        b200CheckError(b200_sgemv(ctx, B200_OP_N, m_, n_, alpha, a_.dev, lda, x_op_.dev,
                                  vecIncrement_, beta, y_op_.dev, vecIncrement_));
      else
        b200CheckError(b200_dgemv(ctx, B200_OP_N, m_, n_, alpha, a_.dev, lda, x_op_.dev,
                                  vecIncrement_, beta, y_op_.dev, vecIncrement_));
    }
  }

  void postLoopRequirements() override {
    if (problem_ != nullptr) {
      b200CheckError(b200_mg_problem_postloop(problem_));
      return;
    }
    b200_ctx* ctx = ctx_.get();
    if (offload_ == gpuOffloadType::once) {
      b200CheckError(b200_d2h_async(ctx, y_op_.host, y_op_.dev, y_op_.bytes, 2));
      b200CheckError(b200_sync(ctx));
    } else if (offload_ == gpuOffloadType::unified) {
      b200CheckError(b200_prefetch_async(ctx, y_op_.dev, y_op_.bytes, 0, 2));
      b200CheckError(b200_sync(ctx));
    }
  }

  void postCallKernelCleanup() override {
    if (problem_ != nullptr) {
      b200CheckError(b200_mg_problem_destroy(problem_));
      problem_ = nullptr;
      return;
    }
    b200_ctx* ctx = ctx_.get();
    const int mode = static_cast<int>(offload_);
    a_.release(ctx, mode);
    x_op_.release(ctx, mode);
    y_op_.release(ctx, mode);
  }

  b200detail::context_holder ctx_;
  b200detail::team_holder team_;        // B200_NGPUS > 1
  b200_mg_problem* problem_ = nullptr;  // the current partitioned problem of the team
  b200detail::operand<T> a_, x_op_, y_op_;
  const T alpha = ALPHA;
  const T beta = BETA;
};

}  // namespace gpu
#endif
