/* b200blas.h -- the C ABI of libb200blas.so, the B200-native (sm_100a) BLAS
 * backend for GPU-BLOB (UoB-HPC/GPU-BLAS-Offload-Benchmark).
 *
 * This is the drop-in boundary: the `GPU_LIB=B200` backend classes in
 * B200/gemm.hh and B200/gemv.hh (gpu::gemm_gpu<T>, gpu::gemv_gpu<T>) are thin
 * host C++ over these entry points, exactly where the reference's cuBLAS
 * backend calls cuBLAS / the CUDA runtime.  Each entry point cites the
 * reference call it replaces (paths relative to the reference repository).
 *
 * Conventions
 *   - plain C, plain pointers and sizes; no CUDA or torch types in signatures
 *     (streams cross the boundary as void*).
 *   - every function returns 0 on success and non-zero on failure; it never
 *     throws and never exits.  b200_last_error() describes the last failure on
 *     the calling thread.  (The reference prints and exit(1)s at the call site,
 *     cuBLAS/common.hh:6-23; B200/common.hh does the same around this ABI.)
 *   - matrices are column-major with BLAS argument order and leading
 *     dimensions; alpha/beta are passed by value.
 *   - compute entry points take DEVICE or MANAGED addresses and are
 *     asynchronous on the context's compute stream, like cublasGemmEx; ordering
 *     against the copy lanes is handled inside the library with events (the
 *     reference gets it implicitly from legacy-default-stream semantics,
 *     cuBLAS/gemm.hh:60-62).
 *   - there is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef B200BLAS_H
#define B200BLAS_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_ctx b200_ctx;

/* Offload modes -- numeric values equal the reference's
 * enum class gpuOffloadType : uint8_t {always = 0, once, unified}
 * (include/utilities.hh:50-54). */
enum { B200_OFFLOAD_ALWAYS = 0, B200_OFFLOAD_ONCE = 1, B200_OFFLOAD_UNIFIED = 2 };

/* Transpose flags for the compute entry points (the harness only uses N). */
enum { B200_OP_N = 0, B200_OP_T = 1 };

/* Number of asynchronous copy lanes (the reference's s1_, s2_, s3_,
 * cuBLAS/gemm.hh:60-62). */
#define B200_NUM_LANES 3

/* ---- context ------------------------------------------------------------
 * Replaces cublasCreate + cublasSetMathMode + cudaGetDevice + 3x
 * cudaStreamCreate (cuBLAS/gemm.hh:46-63, cuBLAS/gemv.hh:46-60) and
 * cublasDestroy + 3x cudaStreamDestroy (cuBLAS/gemm.hh:26-36).
 * `device` < 0 selects the current CUDA device.  Contexts are independent:
 * the harness keeps up to four backend objects alive (src/main.cc:35-59). */
int b200_ctx_create(b200_ctx** out, int device);
int b200_ctx_destroy(b200_ctx* ctx);
/* Use a caller-owned stream (e.g. the torch current stream) as the compute
 * stream; NULL restores the context's own stream.  Not used by the harness. */
int b200_ctx_set_stream(b200_ctx* ctx, void* cuda_stream);
int b200_ctx_device(const b200_ctx* ctx, int* device, int* sm_count);
/* Kernel-selection knobs (the harness never needs them; tests and tuning do).
 *   "sgemm"      = "auto" | "tc" (force tcgen05 3xTF32) | "ffma" (force the FP32-pipe path)
 *   "dgemm_tile" = "auto" | "32" | "64" | "128"
 *   "dgemm_tma"  = "on" | "off"   (large aligned DGEMM: TMA-fed kernel vs cp.async kernel)
 *   "dgemm_streamk" = "on" | "off" (TMA kernel: stream-K over the partial last wave of tiles)
 *   "sgemm_streamk" = "on" | "off" (CTA-pair SGEMM kernel: stream-K over the partial last wave of tiles)
 *   "sgemm_pair" = "auto" | "on" | "off" (tcgen05 SGEMM: CTA-pair 256x256 tiles vs single-CTA 128x128)
 * Defaults come from the environment at context creation: B200_SGEMM, B200_SGEMM_PAIR,
 * B200_DGEMM_TILE, B200_DGEMM_TMA, B200_DGEMM_STREAMK. */
int b200_ctx_set_option(b200_ctx* ctx, const char* key, const char* value);

/* ---- memory -------------------------------------------------------------
 * Replaces cudaMallocHost + cudaMalloc (mode once/always, cuBLAS/gemm.hh:76-82)
 * or cudaMallocManaged (mode unified, cuBLAS/gemm.hh:70-73; then *host ==
 * *dev), and cudaFreeHost + cudaFree / cudaFree (cuBLAS/gemm.hh:223-237).
 * Allocation sits outside the reference's timed region, so freed pinned and
 * device blocks are cached per context and reused. */
int b200_alloc(b200_ctx* ctx, int mode, size_t bytes, void** host, void** dev);
int b200_free(b200_ctx* ctx, int mode, void* host, void* dev);

/* ---- data movement (asynchronous, lane in [0, B200_NUM_LANES)) ----------
 * Replace cudaMemcpyAsync H2D / D2H on s1_..s3_ (cuBLAS/gemm.hh:100-105,
 * 126-131,149-150,204-205) and cudaMemPrefetchAsync to the device / to
 * cudaCpuDeviceId (cuBLAS/gemm.hh:110-115,212-213). */
int b200_h2d_async(b200_ctx* ctx, void* dev, const void* host, size_t bytes, int lane);
int b200_d2h_async(b200_ctx* ctx, void* host, const void* dev, size_t bytes, int lane);
int b200_prefetch_async(b200_ctx* ctx, const void* managed, size_t bytes,
                        int to_device, int lane);
/* Unified-mode placement policy the reference lacks: preferred
 * location = device (+ accessed-by = CPU for inputs; cudaMemAdviseSetReadMostly is NOT
 * used -- it duplicates pages).  Applied to operands >= 64 MiB only: on small operands
 * the advice measurably slows the harness' unified rows (scripts/unified_ab.sh). */
int b200_advise(b200_ctx* ctx, const void* managed, size_t bytes, int is_input);

/* Replaces cudaDeviceSynchronize (cuBLAS/gemm.hh:152,207,215): returns once
 * all work issued through this context is complete. */
int b200_sync(b200_ctx* ctx);

/* ---- compute ------------------------------------------------------------
 * C(m x n) <- alpha * op(A)(m x k) * op(B)(k x n) + beta * C.
 * Replaces cublasGemmEx(..., CUDA_R_32F, CUBLAS_COMPUTE_32F, ...) and
 * cublasGemmEx(..., CUDA_R_64F, CUBLAS_COMPUTE_64F, ...)
 * (cuBLAS/gemm.hh:134-146,158-170,177-187). */
int b200_sgemm(b200_ctx* ctx, int transa, int transb, int m, int n, int k,
               float alpha, const float* A, int lda, const float* B, int ldb,
               float beta, float* C, int ldc);
int b200_dgemm(b200_ctx* ctx, int transa, int transb, int m, int n, int k,
               double alpha, const double* A, int lda, const double* B, int ldb,
               double beta, double* C, int ldc);

/* y <- alpha * op(A) * x + beta * y, A is m x n.
 * Replaces cublasSgemv / cublasDgemv (cuBLAS/gemv.hh:130-137,148-155,161-168). */
int b200_sgemv(b200_ctx* ctx, int trans, int m, int n, float alpha,
               const float* A, int lda, const float* x, int incx, float beta,
               float* y, int incy);
int b200_dgemv(b200_ctx* ctx, int trans, int m, int n, double alpha,
               const double* A, int lda, const double* x, int incx, double beta,
               double* y, int incy);

/* ---- Always-mode transfer engine ---------------------
 * One whole "always" iteration of the reference (cuBLAS/gemm.hh:124-153,
 * cuBLAS/gemv.hh:120-145): upload A, B/x and C/y from PINNED host memory,
 * compute, download C/y, and return with the result valid on the host.
 * The reference issues these back to back with no overlap.  Here large GEMMs
 * run as a K-chunk pipeline (C = sum_c A(:,Kc) B(Kc,:): chunk c+1 uploads while
 * chunk c computes; the last chunk is issued per column slab of C so downloads
 * overlap too), GEMVs stream A in column slabs, small problems go in one shot.
 * h* are host pointers, d* the device buffers from b200_alloc. */
int b200_sgemm_offload(b200_ctx* ctx, int m, int n, int k, float alpha,
                       const float* hA, float* dA, int lda, const float* hB,
                       float* dB, int ldb, float beta, float* hC, float* dC,
                       int ldc);
int b200_dgemm_offload(b200_ctx* ctx, int m, int n, int k, double alpha,
                       const double* hA, double* dA, int lda, const double* hB,
                       double* dB, int ldb, double beta, double* hC, double* dC,
                       int ldc);
int b200_sgemv_offload(b200_ctx* ctx, int m, int n, float alpha, const float* hA,
                       float* dA, int lda, const float* hx, float* dx, float beta,
                       float* hy, float* dy);
int b200_dgemv_offload(b200_ctx* ctx, int m, int n, double alpha,
                       const double* hA, double* dA, int lda, const double* hx,
                       double* dx, double beta, double* hy, double* dy);

/* ---- introspection / measurement ----------------------------------------
 * Name of the kernel path the last compute call on this context dispatched to
 * ("dgemm_dmma_128x128", "sgemm_3xtf32_tcgen05", "gemv_n_vec4", ...). */
const char* b200_last_kernel(const b200_ctx* ctx);
/* Number of kernels this library launched through `ctx` since creation. */
unsigned long long b200_launch_count(const b200_ctx* ctx);
/* Device-side timing of a region on the compute stream (CUDA events). */
int b200_timer_start(b200_ctx* ctx);
int b200_timer_stop(b200_ctx* ctx, float* milliseconds); /* synchronises */
/* Peak probes used as roofline denominators: which = "dfma", "dmma", "ffma",
 * "copy" ; result in GFLOP/s or GB/s. */
int b200_probe_peak(b200_ctx* ctx, const char* which, double* result);

const char* b200_last_error(void);
const char* b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200BLAS_H */
