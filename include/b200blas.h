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
 * the harness keeps up to four backend objects alive (src/main.cc:35-59).
 * Every call that takes a context runs on that context's device whatever device is
 * current in the calling thread, and leaves the thread's current device as it found it. */
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
 *   "sgemm_pair" = "auto" | "on" | "off" | "128" (tcgen05 SGEMM: CTA-pair 256x256 tiles, single-CTA 128x128,
 *                  CTA-pair 256x128 tiles)
 *   "sgemm_fused" = "auto" | "on" | "off" (3xTF32 operand split: fused into the GEMM -- TMA lands the raw
 *                  FP32 tiles, converter warps write the lo parts -- vs a separate pre-pass that writes
 *                  hi/lo copies of A and B into the workspace)
 *   "sgemm_tma_store" = "on" | "off" (tcgen05 SGEMM, beta == 0, 16-byte aligned C: tiles leave through
 *                  shared memory and cp.async.bulk.tensor stores vs register stores)
 *   "dgemm_smem_pad" = "off" | "on" (test knob: launch the TMA DGEMM kernel with the maximum shared-memory
 *                  carve-out, i.e. the minimum L1 -- the setting that exposes stage-release races)
 *   "tallk"      = "on" | "off"   (m, n <= 32 and k >= 1024: the cluster / distributed-shared-memory
 *                  split-K kernel vs the generic split-K paths)
 * Defaults come from the environment at context creation: B200_SGEMM, B200_SGEMM_PAIR, B200_SGEMM_FUSED,
 * B200_SGEMM_TMA_STORE, B200_TALLK, B200_DGEMM_TILE, B200_DGEMM_TMA,
 * B200_DGEMM_STREAMK, B200_UPLOAD_C.
 *   "upload_c"   = "off" | "on"  (Always mode: upload C / y even when beta == 0, as the reference
 *                  does at cuBLAS/gemm.hh:130-131 and cuBLAS/gemv.hh:126-127; default off -- BLAS
 *                  never reads the output operand when beta == 0)
 * b200_ctx_create also sets CUDA_MODULE_LOADING=EAGER in the process environment unless the
 * caller has set it (so no kernel is loaded inside a timed region). */
int b200_ctx_set_option(b200_ctx* ctx, const char* key, const char* value);

/* ---- memory -------------------------------------------------------------
 * Replaces cudaMallocHost + cudaMalloc (mode once/always, cuBLAS/gemm.hh:76-82)
 * or cudaMallocManaged (mode unified, cuBLAS/gemm.hh:70-73; then *host ==
 * *dev), and cudaFreeHost + cudaFree / cudaFree (cuBLAS/gemm.hh:223-237).
 * Allocation sits outside the reference's timed region, so freed pinned and
 * device blocks are cached per context and reused. */
int b200_alloc(b200_ctx* ctx, int mode, size_t bytes, void** host, void** dev);
int b200_free(b200_ctx* ctx, int mode, void* host, void* dev);
/* Device-only block from the same per-context cache (cudaMalloc alone): replicas and slabs
 * that have no host twin (multi-GPU teams, callers that keep data resident). */
int b200_alloc_device(b200_ctx* ctx, size_t bytes, void** dev);
int b200_free_device(b200_ctx* ctx, void* dev);

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
 * run as a dependency-driven block pipeline (A in K-chunks, B and C in column slabs, all
 * contiguous copies down one lane; block C(:,j) += A(:,Kc) B(Kc,j) is issued as soon as its
 * two pieces have arrived and a slab downloads as soon as its last block is done), GEMVs
 * stream A in column slabs, small problems go in one shot, tiny ones run zero-copy.
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

/* Pre-sizes the scratch an SGEMM of this shape needs (the 3xTF32 operand splits), outside the
 * caller's timed region -- what cublasCreate's workspace reservation does for cuBLAS
 * (cuBLAS/gemm.hh:46-52).  Optional: without it the first b200_sgemm of a larger shape grows
 * the scratch itself (one synchronising reallocation). */
int b200_sgemm_prepare(b200_ctx* ctx, int m, int n, int k, int lda, int ldb);

/* ---- multi-GPU teams -----------------------------------------
 * ONE GEMM / GEMV partitioned over the GPUs of an NVSwitch box.  The reference backend owns a
 * single device (cudaGetDevice in gemm_gpu::initialise, cuBLAS/gemm.hh:45-63,
 * cuBLAS/gemv.hh:45-60); with B200_NGPUS > 1 the B200 backend classes create a team there.
 *   GEMM: member g owns a column slab of B and C and a K-share of A; shares are pushed to
 *         every peer's replica of A over NVLink by the copy engines, sub-chunk by sub-chunk,
 *         and consumed chunk by chunk (C_slab += A(:,Kc) B_slab(Kc,:)) while later chunks
 *         are still in flight.
 *   GEMV: member g owns a compact row block of A and y; x is replicated the same way.
 * A team is either LOCAL (one process drives every member: b200_mg_create) or one RANK per
 * process (b200_mg_create_rank + the IPC calls below). */
typedef struct b200_mg b200_mg;
typedef struct b200_mg_problem b200_mg_problem;

#define B200_MG_HANDLE_BYTES 64
enum { B200_MG_GEMM = 0, B200_MG_GEMV = 1 };
/* phases of one member's step (OR them) */
enum {
  B200_MG_UPLOAD = 1,    /* host -> device: my K-share of A (my share of x), my slab of B (block of A) [, C / y] */
  B200_MG_REPLICATE = 2, /* push my share to every peer's replica and wait for theirs, chunk by chunk */
  B200_MG_COMPUTE = 4,   /* the kernels */
  B200_MG_DOWNLOAD = 8,  /* device -> host: my slab of C / block of y */
  B200_MG_GATHER = 16    /* push my result into member 0's full C / y over NVLink */
};

/* Pure host arithmetic (no CUDA): contiguous split of `total` into `parts`, boundaries at
 * multiples of `align` when every part still gets >= 2 * align; and how many members a problem
 * is worth spreading over (small problems stay on one GPU). */
int b200_mg_partition(long long total, int parts, int index, int align, long long* start, long long* count);
int b200_mg_active_gemm(int world, int elem, int m, int n, int k);
int b200_mg_active_gemv(int world, int elem, int m, int n);

int b200_mg_create(b200_mg** out, int ngpus); /* local team on devices 0..ngpus-1, peer access enabled */
int b200_mg_create_rank(b200_mg** out, int device, int rank, int world);
int b200_mg_destroy(b200_mg* mg);
int b200_mg_size(const b200_mg* mg);
int b200_mg_ctx(b200_mg* mg, int member, b200_ctx** ctx); /* a local member's context (allocation, options, timers) */
int b200_mg_sync(b200_mg* mg);                             /* everything issued through local members is complete */
/* Rank teams: 64-byte CUDA IPC handles, carried between processes by the caller.
 * export(dev = NULL) is this rank's flag page, import(dev = NULL) attaches a peer's; with a
 * pointer they export / map a device buffer obtained from b200_alloc (its base address). */
int b200_mg_export(b200_mg* mg, const void* dev, void* handle64);
int b200_mg_import(b200_mg* mg, int member, const void* handle64, void** dev);
int b200_mg_unmap(b200_mg* mg, void* dev);

typedef struct b200_mg_gemm_args {
  int elem;             /* 4 = float, 8 = double */
  int active;           /* members 0..active-1 take part */
  int m, n_slab, k;     /* my part: C_slab (m x n_slab) = alpha A (m x k) B_slab (k x n_slab) + beta C_slab */
  double alpha, beta;
  void* const* A_rep;   /* [active] every member's replica of A, addresses valid in this process */
  int lda;
  const void* hA_share; /* host: first element of MY K-share of A (NULL: already in A_rep[me]) */
  const void* hB;       /* host: my slab of B (NULL: already on the device) */
  void* dB;
  int ldb;
  void* hC;             /* host: my slab of C (destination; also the source when it is uploaded) */
  void* dC;
  int ldc;
  void* C_gather;       /* my slab inside member 0's full C (NULL: none) */
  int phases;
} b200_mg_gemm_args;

typedef struct b200_mg_gemv_args {
  int elem, active;
  int m_blk, n;         /* my part: y_blk (m_blk) = alpha A_blk (m_blk x n) x + beta y_blk */
  double alpha, beta;
  void* dA;             /* my compact row block on the device */
  int lda;
  const void* hA;       /* host: element (first row of my block, column 0) of a matrix with leading dimension h_lda */
  long long h_lda;
  void* const* x_rep;   /* [active] every member's replica of x */
  const void* hx_share; /* host: first element of MY share of x (NULL: already in x_rep[me]) */
  void* hy;
  void* dy;
  void* y_gather;       /* my block inside member 0's full y (NULL: none) */
  int phases;
} b200_mg_gemv_args;

/* One member's step; asynchronous.  Every member of the team must issue the same sequence of
 * steps.  RANK teams call these for their own member.  LOCAL teams use the *_team forms
 * (args[world], one entry per member): each member is issued from its own host thread, because
 * an issue may block on its device (module load, scratch growth) while that device waits for a
 * flag only a peer's step raises -- issuing the members one after the other can deadlock. */
int b200_mg_gemm(b200_mg* mg, int member, const b200_mg_gemm_args* args);
int b200_mg_gemv(b200_mg* mg, int member, const b200_mg_gemv_args* args);
int b200_mg_gemm_team(b200_mg* mg, const b200_mg_gemm_args* args);
int b200_mg_gemv_team(b200_mg* mg, const b200_mg_gemv_args* args);

/* Whole problems on a LOCAL team = the five virtuals of the backend class
 * (cuBLAS/gemm.hh:45-237, cuBLAS/gemv.hh:45-217).  create allocates the host-visible operands
 * the harness fills (h0 = A, h1 = B or x, h2 = C or y; pinned, or managed in unified mode) and
 * the per-member device buffers; k is ignored for GEMV. */
int b200_mg_problem_create(b200_mg* mg, int kind, int elem, int mode, int m, int n, int k,
                           void** h0, void** h1, void** h2, b200_mg_problem** out);
int b200_mg_problem_active(const b200_mg_problem* p);
int b200_mg_problem_preloop(b200_mg_problem* p, double beta);
int b200_mg_problem_call(b200_mg_problem* p, double alpha, double beta);
int b200_mg_problem_postloop(b200_mg_problem* p);
int b200_mg_problem_destroy(b200_mg_problem* p);

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
