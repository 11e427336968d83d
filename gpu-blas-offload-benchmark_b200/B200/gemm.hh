#pragma once
// B200/gemm.hh -- gpu::gemm_gpu<T> for GPU_LIB=B200.
//
// Implements the same plugin slot as the reference's cuBLAS/gemm.hh:14-274:
// the five virtuals of ::gemm<T> / gpu::gemm<T> (include/kernels/gemm.hh:44-58,
// include/kernels/GPU/gemm.hh:20) in the Once, Always and Unified offload
// modes, so include/doGemm.hh drives it unchanged.  Everything that touches
// the device goes through libb200blas.so (hand-written sm_100a kernels).
//
//   mode      preLoop                 callGemm                      postLoop
//   once      H2D A,B,C on 3 lanes    b200_?gemm (async)            D2H C, sync
//   always    -                       b200_?gemm_offload: slabbed   -
//                                     H2D || compute || D2H, sync
//   unified   advise + prefetch ->GPU b200_?gemm on managed ptrs    prefetch C ->CPU, sync
//
// B200_NGPUS=G (> 1) partitions every problem over a team of G GPUs through the same five
// virtuals (b200_mg_problem_*): column slabs of B and C, A uploaded 1/G per GPU over its own
// PCIe link and replicated over NVLink while the GEMM runs (csrc/mg.cu).
#ifdef GPU_B200

#include <algorithm>
#include <type_traits>

#include "../include/kernels/GPU/gemm.hh"
#include "../include/utilities.hh"
#include "common.hh"

namespace gpu {

template <typename T>
class gemm_gpu : public gemm<T> {
 public:
  using gemm<T>::gemm;
  using gemm<T>::initInputMatrices;
  using gemm<T>::m_;
  using gemm<T>::n_;
  using gemm<T>::k_;
  using gemm<T>::A_;
  using gemm<T>::B_;
  using gemm<T>::C_;
  using gemm<T>::offload_;

  /** Allocate the operands for this problem size and offload mode and fill the
   * inputs with the harness' generator (untimed). */
  void initialise(gpuOffloadType offload, int m, int n, int k) override {
    offload_ = offload;
    m_ = m;
    n_ = n;
    k_ = k;
    const int mode = static_cast<int>(offload_);
    if (b200detail::ngpus() > 1) {
      void *h0 = nullptr, *h1 = nullptr, *h2 = nullptr;
      b200CheckError(b200_mg_problem_create(team_.get(), B200_MG_GEMM, (int)sizeof(T), mode, m_, n_, k_,
                                            &h0, &h1, &h2, &problem_));
      A_ = static_cast<T*>(h0);
      B_ = static_cast<T*>(h1);
      C_ = static_cast<T*>(h2);
      initInputMatrices();
      return;
    }
    b200_ctx* ctx = ctx_.get();
    a_.acquire(ctx, mode, (size_t)m_ * k_);
    b_.acquire(ctx, mode, (size_t)k_ * n_);
    c_.acquire(ctx, mode, (size_t)m_ * n_);
    A_ = a_.host;
    B_ = b_.host;
    C_ = c_.host;
    if (offload_ == gpuOffloadType::unified) {
      b200CheckError(b200_advise(ctx, a_.dev, a_.bytes, 1));
      b200CheckError(b200_advise(ctx, b_.dev, b_.bytes, 1));
      b200CheckError(b200_advise(ctx, c_.dev, c_.bytes, 0));
    }
    if constexpr (std::is_same_v<T, float>)
      b200CheckError(b200_sgemm_prepare(ctx, m_, n_, k_, std::max(1, m_), std::max(1, k_)));
    initInputMatrices();
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
      b200CheckError(b200_h2d_async(ctx, b_.dev, b_.host, b_.bytes, 1));
      // BLAS never reads the output operand when beta == 0, so the reference's
      // third upload (cuBLAS/gemm.hh:104-105, cuBLAS/gemv.hh:100-101) is skipped
      // (B200_UPLOAD_C=1 restores it)
      if (beta != T(0) || b200detail::upload_c()) b200CheckError(b200_h2d_async(ctx, c_.dev, c_.host, c_.bytes, 2));
    } else if (offload_ == gpuOffloadType::unified) {
      b200CheckError(b200_prefetch_async(ctx, a_.dev, a_.bytes, 1, 0));
      b200CheckError(b200_prefetch_async(ctx, b_.dev, b_.bytes, 1, 1));
      b200CheckError(b200_prefetch_async(ctx, c_.dev, c_.bytes, 1, 2));
    }
  }

  void callGemm() override {
    if (problem_ != nullptr) {
      b200CheckError(b200_mg_problem_call(problem_, (double)alpha, (double)beta));
      return;
    }
    b200_ctx* ctx = ctx_.get();
    const int lda = std::max(1, m_), ldb = std::max(1, k_), ldc = std::max(1, m_);
    if (offload_ == gpuOffloadType::always) {
      if constexpr (std::is_same_v<T, float>)
        b200CheckError(b200_sgemm_offload(ctx, m_, n_, k_, alpha, a_.host, a_.dev, lda, b_.host,
                                          b_.dev, ldb, beta, c_.host, c_.dev, ldc));
      else
        b200CheckError(b200_dgemm_offload(ctx, m_, n_, k_, alpha, a_.host, a_.dev, lda, b_.host,
                                          b_.dev, ldb, beta, c_.host, c_.dev, ldc));
    } else {
      if constexpr (std::is_same_v<T, float>)
        b200CheckError(b200_sgemm(ctx, B200_OP_N, B200_OP_N, m_, n_, k_, alpha, a_.dev, lda,
                                  b_.dev, ldb, beta, c_.dev, ldc));
      else
        b200CheckError(b200_dgemm(ctx, B200_OP_N, B200_OP_N, m_, n_, k_, alpha, a_.dev, lda,
                                  b_.dev, ldb, beta, c_.dev, ldc));
    }
  }

  void postLoopRequirements() override {
    if (problem_ != nullptr) {
      b200CheckError(b200_mg_problem_postloop(problem_));
      return;
    }
    b200_ctx* ctx = ctx_.get();
    if (offload_ == gpuOffloadType::once) {
      b200CheckError(b200_d2h_async(ctx, c_.host, c_.dev, c_.bytes, 2));
      b200CheckError(b200_sync(ctx));
    } else if (offload_ == gpuOffloadType::unified) {
      b200CheckError(b200_prefetch_async(ctx, c_.dev, c_.bytes, 0, 2));
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
    b_.release(ctx, mode);
    c_.release(ctx, mode);
  }

  b200detail::context_holder ctx_;
  b200detail::team_holder team_;        // B200_NGPUS > 1
  b200_mg_problem* problem_ = nullptr;  // the current partitioned problem of the team
  b200detail::operand<T> a_, b_, c_;
  const T alpha = ALPHA;
  const T beta = BETA;
};

}  // namespace gpu
#endif
