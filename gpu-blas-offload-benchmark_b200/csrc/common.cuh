// common.cuh -- context, error plumbing and small device helpers shared by the
// translation units of libb200blas.so.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <vector>

#include "../../include/b200blas.h"

namespace b200 {

void set_error(const char* fmt, ...);

#define B200_CUDA(call)                                                        \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      b200::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,            \
                      cudaGetErrorString(e__));                                \
      return 1;                                                                \
    }                                                                          \
  } while (0)

#define B200_REQUIRE(cond, ...)                                                \
  do {                                                                         \
    if (!(cond)) {                                                             \
      b200::set_error(__VA_ARGS__);                                            \
      return 2;                                                                \
    }                                                                          \
  } while (0)

struct CachedBlock {
  void* ptr;
  size_t capacity;
  void* alias = nullptr;  // pinned host blocks: the address the DEVICE uses for the same memory
};

}  // namespace b200

// The opaque context of the C ABI.
struct b200_ctx {
  int device = 0;
  int sm_count = 148;
  size_t l2_bytes = 0;
  cudaStream_t own_compute = nullptr;  // non-blocking
  cudaStream_t compute = nullptr;      // own_compute or a caller stream
  cudaStream_t lane[B200_NUM_LANES] = {};
  cudaEvent_t lane_done[B200_NUM_LANES] = {};
  bool lane_pending[B200_NUM_LANES] = {};  // compute must wait for this lane
  bool lane_dirty[B200_NUM_LANES] = {};    // lane has un-synchronised work
  cudaEvent_t compute_done = nullptr;
  bool compute_dirty = false;      // kernels issued since the last b200_sync
  bool compute_unordered = false;  // kernels issued since compute_done was recorded
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  // event ring for the transfer engine (slab / chunk dependencies)
  static constexpr int kEvPool = 256;
  cudaEvent_t ev_pool[kEvPool] = {};
  int ev_pool_size = 0, ev_next = 0;

  // scratch: split-K / split-N partials, 3xTF32 operand splits
  void* workspace = nullptr;
  size_t workspace_bytes = 0;
  // self-resetting arrival counters for last-block reductions
  unsigned int* counters = nullptr;
  int num_counters = 0;

  std::vector<b200::CachedBlock> pinned_cache, device_cache, managed_cache;  // freed, reusable
  std::vector<b200::CachedBlock> live_host, live_dev, live_managed;          // handed out

  // kernel-selection knobs (b200_ctx_set_option)
  int opt_sgemm = 0;       // 0 auto, 1 tc, 2 ffma
  int opt_sgemm_streamk = 1;  // CTA-pair SGEMM kernel: 1 stream-K over the partial last wave, 0 whole tiles only
  int opt_sgemm_pair = 0;  // tcgen05 SGEMM tile config: 0 auto, 1 CTA-pair 256x256, 2 single-CTA 128x128, 3 CTA-pair 256x128
  int opt_sgemm_tma_store = 1;  // tcgen05 SGEMM epilogue: 1 staged through shared memory + TMA store when C allows, 0 register stores
  int opt_sgemm_fused = 0; // 3xTF32 operand split: 0 auto, 1 fused into the GEMM (converter warps), 2 separate pre-pass
  int opt_dgemm_tile = 0;  // 0 auto, 64, 128
  // Always mode: operand sets up to these sizes are computed in place from pinned host
  // memory (measured break-evens, scripts/zerocopy_ab.sh: GEMM re-reads tiles over PCIe,
  // GEMV touches every element once)
  size_t opt_zerocopy_gemm_bytes = 200u << 10;
  size_t opt_zerocopy_gemv_bytes = 16u << 20;
  int opt_tallk = 1;       // M, N <= 32 with K >= 1024: the cluster / DSMEM split-K kernel (gemm_tallk.cu); 0 = the generic split-K paths
  int opt_dgemm_min = 32;  // smallest m, n the DMMA kernels take (one 32x32 tile; below: DFMA SIMT path)
  int opt_dgemm_streamk = 1;  // 1: stream-K over the partial last wave of the TMA kernel, 0: whole tiles only
  int opt_dgemm_smem_pad = 0;  // test knob: launch the TMA DGEMM kernel with the maximum shared-memory carve-out (minimum L1)
  int opt_dgemm_tma = 1;   // 1: large aligned DGEMMs use the TMA-fed kernel, 0: cp.async kernel
  // Always mode: 1 = upload C / y even when beta == 0, as the reference does
  // (cuBLAS/gemm.hh:130-131, cuBLAS/gemv.hh:126-127); 0 = skip the dead upload.  B200_UPLOAD_C.
  int opt_upload_c = 0;

  const char* last_kernel = "none";
  unsigned long long launches = 0;
};

namespace b200 {

// cudaFuncSetAttribute is per device: one-time flags are kept per device id
constexpr int kMaxDevices = 64;

// Every ABI call runs with its context's device current and hands the calling thread back on
// the device it came with.  (Allocations and launches bind to the CURRENT device, not to the
// stream's: after a team call had left the thread on its last member's GPU, a single-GPU
// context grew its workspace on that GPU and ran its kernels through peer mappings.)  The
// common case -- already on the right device -- costs one cudaGetDevice.
struct DeviceScope {
  int prev = -1;
  bool changed = false;
  explicit DeviceScope(int device) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != device) changed = cudaSetDevice(device) == cudaSuccess;
  }
  ~DeviceScope() {
    if (changed) cudaSetDevice(prev);
  }
  DeviceScope(const DeviceScope&) = delete;
  DeviceScope& operator=(const DeviceScope&) = delete;
};
// Team calls walk over several devices: put the thread back where it was.
struct DeviceRestore {
  int prev = -1;
  DeviceRestore() {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
  }
  ~DeviceRestore() {
    int now = -1;
    if (prev >= 0 && cudaGetDevice(&now) == cudaSuccess && now != prev) cudaSetDevice(prev);
  }
};

int ensure_workspace(b200_ctx* ctx, size_t bytes);
// Transfer-engine plumbing shared by the single-GPU (b200blas.cu) and multi-GPU (mg.cu) engines.
int engine_event(b200_ctx* ctx, cudaEvent_t* ev);               // next event of the context's ring
int engine_begin(b200_ctx* ctx);                                // order a call after everything before it
int record_on(b200_ctx* ctx, cudaStream_t s, cudaEvent_t* ev);  // ring event recorded on s
// Managed-memory calls whose signature changed in CUDA 13 (cudaMemLocation).  device < 0 = the CPU.
cudaError_t mem_prefetch(const void* p, size_t bytes, int device, cudaStream_t s);
cudaError_t mem_advise(const void* p, size_t bytes, cudaMemoryAdvise advice, int device);
// Device-only block from the context's cache (no pinned twin).
int alloc_device(b200_ctx* ctx, size_t bytes, void** dev);
int free_device(b200_ctx* ctx, void* dev);
// Make the compute stream wait for every copy lane that has pending uploads.
int join_lanes_into_compute(b200_ctx* ctx);
inline void note_launch(b200_ctx* ctx, const char* name, int n = 1) {
  ctx->last_kernel = name;
  ctx->launches += (unsigned long long)n;
  ctx->compute_dirty = true;
  ctx->compute_unordered = true;
}

// ---- kernel-family entry points (one per .cu) -----------------------------
template <typename T>
int gemv_dispatch(b200_ctx* ctx, int trans, int m, int n, T alpha, const T* A,
                  int lda, const T* x, int incx, T beta, T* y, int incy);

template <typename T>
int gemm_simt_dispatch(b200_ctx* ctx, int ta, int tb, int m, int n, int k,
                       T alpha, const T* A, int lda, const T* B, int ldb, T beta,
                       T* C, int ldc);

// FP64 tensor-core (DMMA) path; returns -1 if the shape is not eligible.
int dgemm_dmma_dispatch(b200_ctx* ctx, int m, int n, int k, double alpha,
                        const double* A, int lda, const double* B, int ldb,
                        double beta, double* C, int ldc);

// TMA-fed variant of the large-tile DMMA kernel; returns -1 if not eligible.
int dgemm_tma_dispatch(b200_ctx* ctx, int m, int n, int k, double alpha,
                       const double* A, int lda, const double* B, int ldb,
                       double beta, double* C, int ldc);

// 3xTF32 tcgen05/TMEM path; returns -1 if the shape is not eligible.
int sgemm_tc_dispatch(b200_ctx* ctx, int m, int n, int k, float alpha,
                      const float* A, int lda, const float* B, int ldb,
                      float beta, float* C, int ldc);

int sgemm_tc_prepare(b200_ctx* ctx, int m, int n, int k, int lda, int ldb);

// One 32 x 32 tile of C, long K (the reference's M = N = 32 problem type): clusters of 8 CTAs,
// reduction through distributed shared memory; returns -1 if the shape is not eligible.
template <typename T>
int gemm_tallk_dispatch(b200_ctx* ctx, int m, int n, int k, T alpha, const T* A, int lda, const T* B, int ldb, T beta,
                        T* C, int ldc);

int probe_peak(b200_ctx* ctx, const char* which, double* result);

// ---- profiling hooks -----------------------------------
// B200_PROFILE=1 wraps every ABI call that issues device work in an NVTX range (header-only
// NVTX3: the tool's library is loaded only when a profiler is attached), so an nsys / ncu
// timeline shows "b200_dgemm 8192x8192x8192", "b200_dgemm_offload ...", "mg gemm step ..."
// next to the kernels.  Off by default: one getenv at start-up, one branch per call.
bool profile_enabled();
void profile_push(const char* fmt, ...);
void profile_pop();
struct ProfileRange {
  bool on;
  template <typename... Args>
  explicit ProfileRange(const char* fmt, Args... args) : on(profile_enabled()) {
    if (on) profile_push(fmt, args...);
  }
  ~ProfileRange() {
    if (on) profile_pop();
  }
};

// ---- device helpers --------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace b200
