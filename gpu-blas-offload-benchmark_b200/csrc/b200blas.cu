// b200blas.cu -- the extern "C" ABI of libb200blas.so (see include/b200blas.h):
// context, cached allocation, asynchronous copy lanes with explicit event
// ordering, kernel dispatch and the Always-mode transfer engine.
//
// Ordering model.  The reference relies on legacy-default-stream semantics:
// s1_..s3_ are blocking streams and cuBLAS runs on the NULL stream, so the GEMM
// implicitly waits for the three uploads and the D2H on s3_ implicitly waits
// for the GEMM (cuBLAS/gemm.hh:60-62,126-150).  Here every stream is
// non-blocking and the same ordering is explicit:
//   upload on lane L   -> record lane_done[L], mark lane_pending[L]
//   compute call       -> compute stream waits on every pending lane event
//   download on lane L -> lane waits on an event recorded after the last kernel
// so back-to-back kernels in Once mode cost one launch each and nothing else.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"

namespace b200 {

static thread_local char g_error[512] = "no error";

bool profile_enabled() {
  static const bool on = [] {
    const char* e = getenv("B200_PROFILE");
    return e != nullptr && strcmp(e, "0") != 0;
  }();
  return on;
}
void profile_push(const char* fmt, ...) {
  char buf[160];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  nvtxRangePushA(buf);
}
void profile_pop() { nvtxRangePop(); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int ensure_workspace(b200_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->workspace_bytes) return 0;
  // the old workspace may still be in use by queued kernels
  B200_CUDA(cudaStreamSynchronize(ctx->compute));
  if (ctx->workspace) B200_CUDA(cudaFree(ctx->workspace));
  ctx->workspace = nullptr;
  ctx->workspace_bytes = 0;
  size_t cap = std::max(bytes + bytes / 4, (size_t)(8u << 20));
  B200_CUDA(cudaMalloc(&ctx->workspace, cap));
  ctx->workspace_bytes = cap;
  return 0;
}

int join_lanes_into_compute(b200_ctx* ctx) {
  for (int l = 0; l < B200_NUM_LANES; l++) {
    if (ctx->lane_pending[l]) {
      B200_CUDA(cudaStreamWaitEvent(ctx->compute, ctx->lane_done[l], 0));
      ctx->lane_pending[l] = false;
    }
  }
  return 0;
}

// A copy lane must not overtake kernels that still read/write the buffer.
static int lanes_wait_for_compute(b200_ctx* ctx) {
  if (!ctx->compute_unordered) return 0;
  B200_CUDA(cudaEventRecord(ctx->compute_done, ctx->compute));
  for (int l = 0; l < B200_NUM_LANES; l++)
    B200_CUDA(cudaStreamWaitEvent(ctx->lane[l], ctx->compute_done, 0));
  ctx->compute_unordered = false;
  return 0;
}

// A sweep to d = 320 walks through ~15 capacity classes per operand; with 12 cached blocks
// the pool kept evicting (and re-creating) them, and a FRESH managed allocation costs
// 4-7 ms in the first timed region that prefetches it (profiles/r1_unified_notes.md).
constexpr size_t kMaxCachedBlocks = 48;

// Managed blocks: power-of-two classes.  Untouched managed pages cost nothing physical, a
// coarse class means a handful of allocations per process instead of one per size.
static size_t round_capacity_managed(size_t bytes) {
  size_t p = 64u << 10;
  while (p < bytes) p *= 2;
  return p;
}

static size_t round_capacity(size_t bytes) {
  const size_t min_gran = 64u << 10;
  if (bytes <= min_gran) return min_gran;
  size_t p = 1;
  while (p * 2 <= bytes) p *= 2;
  const size_t gran = std::max(min_gran, p / 8);
  return (bytes + gran - 1) / gran * gran;
}

static void* cache_take(std::vector<CachedBlock>& cache, size_t cap, void** alias = nullptr) {
  for (size_t i = 0; i < cache.size(); i++) {
    if (cache[i].capacity == cap) {
      void* p = cache[i].ptr;
      if (alias) *alias = cache[i].alias;
      cache.erase(cache.begin() + i);
      return p;
    }
  }
  return nullptr;
}

template <typename FreeFn>
static void cache_put(std::vector<CachedBlock>& cache, void* p, size_t cap,
                      size_t max_total, FreeFn free_fn, void* alias = nullptr) {
  if (cap > max_total) {
    free_fn(p);
    return;
  }
  cache.push_back({p, cap, alias});
  size_t total = 0;
  for (auto& b : cache) total += b.capacity;
  while (cache.size() > kMaxCachedBlocks || total > max_total) {
    total -= cache.front().capacity;
    free_fn(cache.front().ptr);
    cache.erase(cache.begin());
  }
}

// capacity bookkeeping for live blocks (ptr -> capacity)
static size_t live_pop(std::vector<CachedBlock>& v, void* p, void** alias = nullptr) {
  for (size_t i = 0; i < v.size(); i++)
    if (v[i].ptr == p) {
      const size_t c = v[i].capacity;
      if (alias) *alias = v[i].alias;
      v.erase(v.begin() + i);
      return c;
    }
  return 0;
}

// cudaMemPrefetchAsync / cudaMemAdvise take a cudaMemLocation from CUDA 13 on (the 4-argument
// forms the reference uses, cuBLAS/gemm.hh:110-115,212-213, are gone there); 12.2+ has the
// same thing as *_v2.  device < 0 selects the CPU.
cudaError_t mem_prefetch(const void* p, size_t bytes, int device, cudaStream_t s) {
#if CUDART_VERSION >= 12020
  cudaMemLocation loc;
  loc.type = device < 0 ? cudaMemLocationTypeHost : cudaMemLocationTypeDevice;
  loc.id = device < 0 ? 0 : device;
#if CUDART_VERSION >= 13000
  return cudaMemPrefetchAsync(p, bytes, loc, 0, s);
#else
  return cudaMemPrefetchAsync_v2(p, bytes, loc, 0, s);
#endif
#else
  return cudaMemPrefetchAsync(p, bytes, device < 0 ? cudaCpuDeviceId : device, s);
#endif
}

cudaError_t mem_advise(const void* p, size_t bytes, cudaMemoryAdvise advice, int device) {
#if CUDART_VERSION >= 12020
  cudaMemLocation loc;
  loc.type = device < 0 ? cudaMemLocationTypeHost : cudaMemLocationTypeDevice;
  loc.id = device < 0 ? 0 : device;
#if CUDART_VERSION >= 13000
  return cudaMemAdvise(p, bytes, advice, loc);
#else
  return cudaMemAdvise_v2(p, bytes, advice, loc);
#endif
#else
  return cudaMemAdvise(p, bytes, advice, device < 0 ? cudaCpuDeviceId : device);
#endif
}

// Free every cached (idle) block of this context: called when an allocation fails so that
// a long sweep does not die with most of HBM sitting in idle caches.
static void trim_caches(b200_ctx* ctx) {
  for (auto& b : ctx->pinned_cache) cudaFreeHost(b.ptr);
  for (auto& b : ctx->device_cache) cudaFree(b.ptr);
  for (auto& b : ctx->managed_cache) cudaFree(b.ptr);
  ctx->pinned_cache.clear();
  ctx->device_cache.clear();
  ctx->managed_cache.clear();
  cudaGetLastError();
}

static int take_device_block(b200_ctx* ctx, size_t cap, void** out) {
  void* d = cache_take(ctx->device_cache, cap);
  if (!d) {
    cudaError_t e = cudaMalloc(&d, cap);
    if (e != cudaSuccess) {
      cudaGetLastError();
      trim_caches(ctx);
      e = cudaMalloc(&d, cap);
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      set_error("b200_alloc: cudaMalloc(%zu) -> %s", cap, cudaGetErrorString(e));
      return 1;
    }
  }
  *out = d;
  return 0;
}

int alloc_device(b200_ctx* ctx, size_t bytes, void** dev) {
  const size_t cap = round_capacity(bytes ? bytes : 1);
  if (take_device_block(ctx, cap, dev)) return 1;
  ctx->live_dev.push_back({*dev, cap});
  return 0;
}

int free_device(b200_ctx* ctx, void* dev) {
  if (!dev) return 0;
  const size_t cap = live_pop(ctx->live_dev, dev);
  if (cap)
    cache_put(ctx->device_cache, dev, cap, (size_t)48 << 30, [](void* p) { cudaFree(p); });
  else
    B200_CUDA(cudaFree(dev));
  return 0;
}

// Device-side address of a pinned host pointer handed out by b200_alloc (nullptr if the
// pointer is not ours -- then the zero-copy path is not taken).
static void* device_alias(const b200_ctx* ctx, const void* host) {
  for (const auto& b : ctx->live_host) {
    const char* base = (const char*)b.ptr;
    if ((const char*)host >= base && (const char*)host < base + b.capacity && b.alias)
      return (char*)b.alias + ((const char*)host - base);
  }
  return nullptr;
}

}  // namespace b200

using namespace b200;

extern "C" {

const char* b200_last_error(void) { return g_error; }
const char* b200_version(void) { return "b200blas 0.1 (sm_100a)"; }

int b200_ctx_create(b200_ctx** out, int device) {
  B200_REQUIRE(out != nullptr, "b200_ctx_create: out is null");
  *out = nullptr;
  // Load every kernel of this library when the CUDA context is created instead
  // of on first launch: a lazily loaded kernel would otherwise be charged to
  // whichever timed region happens to launch it first.  Only effective when
  // this call is the one that initialises CUDA (it is, in the harness); an
  // explicit user setting wins.
  setenv("CUDA_MODULE_LOADING", "EAGER", 0);
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error("b200_ctx_create: no CUDA device (%s); this library has no CPU fallback",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return 1;
  }
  if (device < 0) B200_CUDA(cudaGetDevice(&device));
  B200_REQUIRE(device < count, "b200_ctx_create: device %d out of range (%d)", device, count);
  DeviceScope ds(device);
  cudaDeviceProp prop;
  B200_CUDA(cudaGetDeviceProperties(&prop, device));
  B200_REQUIRE(prop.major == 10, "b200_ctx_create: device %d is sm_%d%d; this build is sm_100a only",
               device, prop.major, prop.minor);
  b200_ctx* ctx = new b200_ctx();
  // any failure below destroys what has been created so far
  struct Guard {
    b200_ctx* c;
    ~Guard() { if (c) b200_ctx_destroy(c); }
  } guard{ctx};
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->l2_bytes = (size_t)prop.l2CacheSize;
  B200_CUDA(cudaStreamCreateWithFlags(&ctx->own_compute, cudaStreamNonBlocking));
  ctx->compute = ctx->own_compute;
  for (int l = 0; l < B200_NUM_LANES; l++) {
    B200_CUDA(cudaStreamCreateWithFlags(&ctx->lane[l], cudaStreamNonBlocking));
    B200_CUDA(cudaEventCreateWithFlags(&ctx->lane_done[l], cudaEventDisableTiming));
  }
  B200_CUDA(cudaEventCreateWithFlags(&ctx->compute_done, cudaEventDisableTiming));
  B200_CUDA(cudaEventCreate(&ctx->t0));
  B200_CUDA(cudaEventCreate(&ctx->t1));
  if (const char* e = getenv("B200_SGEMM")) ctx->opt_sgemm = !strcmp(e, "tc") ? 1 : !strcmp(e, "ffma") ? 2 : 0;
  if (const char* e = getenv("B200_SGEMM_STREAMK")) ctx->opt_sgemm_streamk = strcmp(e, "off") != 0 && strcmp(e, "0") != 0;
  if (const char* e = getenv("B200_SGEMM_PAIR")) ctx->opt_sgemm_pair = !strcmp(e, "on") ? 1 : !strcmp(e, "off") ? 2 : !strcmp(e, "128") ? 3 : 0;
  if (const char* e = getenv("B200_SGEMM_TMA_STORE")) ctx->opt_sgemm_tma_store = strcmp(e, "off") != 0 && strcmp(e, "0") != 0;
  if (const char* e = getenv("B200_TALLK")) ctx->opt_tallk = strcmp(e, "off") != 0 && strcmp(e, "0") != 0;
  if (const char* e = getenv("B200_SGEMM_FUSED")) ctx->opt_sgemm_fused = !strcmp(e, "on") ? 1 : !strcmp(e, "off") ? 2 : 0;
  if (const char* e = getenv("B200_ZEROCOPY_GEMM_BYTES")) ctx->opt_zerocopy_gemm_bytes = (size_t)atoll(e);
  if (const char* e = getenv("B200_ZEROCOPY_GEMV_BYTES")) ctx->opt_zerocopy_gemv_bytes = (size_t)atoll(e);
  if (const char* e = getenv("B200_DGEMM_MIN")) ctx->opt_dgemm_min = atoi(e);
  if (const char* e = getenv("B200_DGEMM_TMA")) ctx->opt_dgemm_tma = strcmp(e, "off") != 0 && strcmp(e, "0") != 0;
  if (const char* e = getenv("B200_DGEMM_STREAMK")) ctx->opt_dgemm_streamk = strcmp(e, "off") != 0 && strcmp(e, "0") != 0;
  if (const char* e = getenv("B200_DGEMM_TILE")) ctx->opt_dgemm_tile = (atoi(e) == 32 || atoi(e) == 64 || atoi(e) == 128) ? atoi(e) : 0;
  if (const char* e = getenv("B200_UPLOAD_C")) ctx->opt_upload_c = strcmp(e, "0") != 0 && strcmp(e, "off") != 0;
  ctx->num_counters = 1 << 16;
  B200_CUDA(cudaMalloc((void**)&ctx->counters, ctx->num_counters * sizeof(unsigned int)));
  B200_CUDA(cudaMemset(ctx->counters, 0, ctx->num_counters * sizeof(unsigned int)));
  // Scratch for split-K / split-N / stream-K partial tiles (<= ~56 MB whatever the problem
  // size) is reserved here, once, like cuBLAS reserves its workspace in cublasCreate;
  // b200_sgemm_prepare adds the 3xTF32 operand splits for one SGEMM shape.
  if (ensure_workspace(ctx, (size_t)96 << 20)) return 1;
  guard.c = nullptr;
  *out = ctx;
  return 0;
}

int b200_ctx_destroy(b200_ctx* ctx) {
  if (!ctx) return 0;
  DeviceScope ds(ctx->device);
  if (ctx->own_compute) cudaStreamSynchronize(ctx->own_compute);
  for (int l = 0; l < B200_NUM_LANES; l++)
    if (ctx->lane[l]) cudaStreamSynchronize(ctx->lane[l]);
  for (auto& b : ctx->pinned_cache) cudaFreeHost(b.ptr);
  for (auto& b : ctx->device_cache) cudaFree(b.ptr);
  for (auto& b : ctx->managed_cache) cudaFree(b.ptr);
  if (ctx->workspace) cudaFree(ctx->workspace);
  if (ctx->counters) cudaFree(ctx->counters);
  for (int l = 0; l < B200_NUM_LANES; l++) {
    if (ctx->lane[l]) cudaStreamDestroy(ctx->lane[l]);
    if (ctx->lane_done[l]) cudaEventDestroy(ctx->lane_done[l]);
  }
  if (ctx->compute_done) cudaEventDestroy(ctx->compute_done);
  for (int i = 0; i < ctx->ev_pool_size; i++) cudaEventDestroy(ctx->ev_pool[i]);
  if (ctx->t0) cudaEventDestroy(ctx->t0);
  if (ctx->t1) cudaEventDestroy(ctx->t1);
  if (ctx->own_compute) cudaStreamDestroy(ctx->own_compute);
  delete ctx;
  return 0;
}

int b200_ctx_set_stream(b200_ctx* ctx, void* cuda_stream) {
  B200_REQUIRE(ctx, "b200_ctx_set_stream: null context");
  ctx->compute = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_compute;
  return 0;
}

int b200_ctx_set_option(b200_ctx* ctx, const char* key, const char* value) {
  B200_REQUIRE(ctx && key && value, "b200_ctx_set_option: null argument");
  if (!strcmp(key, "sgemm")) {
    if (!strcmp(value, "auto")) ctx->opt_sgemm = 0;
    else if (!strcmp(value, "tc")) ctx->opt_sgemm = 1;
    else if (!strcmp(value, "ffma")) ctx->opt_sgemm = 2;
    else B200_REQUIRE(false, "b200_ctx_set_option: sgemm must be auto|tc|ffma, got '%s'", value);
    return 0;
  }
  if (!strcmp(key, "sgemm_streamk")) {
    if (!strcmp(value, "on") || !strcmp(value, "auto")) ctx->opt_sgemm_streamk = 1;
    else if (!strcmp(value, "off")) ctx->opt_sgemm_streamk = 0;
    else B200_REQUIRE(false, "b200_ctx_set_option: sgemm_streamk must be on|off, got '%s'", value);
    return 0;
  }
  if (!strcmp(key, "sgemm_pair")) {
    if (!strcmp(value, "auto")) ctx->opt_sgemm_pair = 0;
    else if (!strcmp(value, "on")) ctx->opt_sgemm_pair = 1;
    else if (!strcmp(value, "off")) ctx->opt_sgemm_pair = 2;
    else if (!strcmp(value, "128")) ctx->opt_sgemm_pair = 3;
    else B200_REQUIRE(false, "b200_ctx_set_option: sgemm_pair must be auto|on|off|128, got '%s'", value);
    return 0;
  }
  if (!strcmp(key, "sgemm_tma_store")) {
    if (!strcmp(value, "on") || !strcmp(value, "auto")) ctx->opt_sgemm_tma_store = 1;
    else if (!strcmp(value, "off")) ctx->opt_sgemm_tma_store = 0;
    else B200_REQUIRE(false, "b200_ctx_set_option: sgemm_tma_store must be on|off, got '%s'", value);
    return 0;
  }
  if (!strcmp(key, "dgemm_smem_pad")) {
    if (!strcmp(value, "on")) ctx->opt_dgemm_smem_pad = 1;
    else if (!strcmp(value, "off") || !strcmp(value, "auto")) ctx->opt_dgemm_smem_pad = 0;
    else B200_REQUIRE(false, "b200_ctx_set_option: dgemm_smem_pad must be on|off, got '%s'", value);
    return 0;
  }
  if (!strcmp(key, "tallk")) {
    if (!strcmp(value, "on") || !strcmp(value, "auto")) ctx->opt_tallk = 1;
    else if (!strcmp(value, "off")) ctx->opt_tallk = 0;
    else B200_REQUIRE(false, "b200_ctx_set_option: tallk must be on|off, got '%s'", value);
    return 0;
  }
  if (!strcmp(key, "sgemm_fused")) {
    if (!strcmp(value, "auto")) ctx->opt_sgemm_fused = 0;
    else if (!strcmp(value, "on")) ctx->opt_sgemm_fused = 1;
    else if (!strcmp(value, "off")) ctx->opt_sgemm_fused = 2;
    else B200_REQUIRE(false, "b200_ctx_set_option: sgemm_fused must be auto|on|off, got '%s'", value);
    return 0;
  }
  if (!strcmp(key, "dgemm_tile")) {
    if (!strcmp(value, "auto")) ctx->opt_dgemm_tile = 0;
    else if (!strcmp(value, "32")) ctx->opt_dgemm_tile = 32;
    else if (!strcmp(value, "64")) ctx->opt_dgemm_tile = 64;
    else if (!strcmp(value, "128")) ctx->opt_dgemm_tile = 128;
    else B200_REQUIRE(false, "b200_ctx_set_option: dgemm_tile must be auto|32|64|128, got '%s'", value);
    return 0;
  }
  if (!strcmp(key, "dgemm_streamk")) {
    if (!strcmp(value, "on") || !strcmp(value, "auto")) ctx->opt_dgemm_streamk = 1;
    else if (!strcmp(value, "off")) ctx->opt_dgemm_streamk = 0;
    else B200_REQUIRE(false, "b200_ctx_set_option: dgemm_streamk must be on|off, got '%s'", value);
    return 0;
  }
  if (!strcmp(key, "dgemm_tma")) {
    if (!strcmp(value, "on") || !strcmp(value, "auto")) ctx->opt_dgemm_tma = 1;
    else if (!strcmp(value, "off")) ctx->opt_dgemm_tma = 0;
    else B200_REQUIRE(false, "b200_ctx_set_option: dgemm_tma must be on|off, got '%s'", value);
    return 0;
  }
  if (!strcmp(key, "upload_c")) {
    if (!strcmp(value, "on") || !strcmp(value, "1")) ctx->opt_upload_c = 1;
    else if (!strcmp(value, "off") || !strcmp(value, "0")) ctx->opt_upload_c = 0;
    else B200_REQUIRE(false, "b200_ctx_set_option: upload_c must be on|off, got '%s'", value);
    return 0;
  }
  B200_REQUIRE(false, "b200_ctx_set_option: unknown key '%s'", key);
  return 0;
}

int b200_ctx_device(const b200_ctx* ctx, int* device, int* sm_count) {
  B200_REQUIRE(ctx, "b200_ctx_device: null context");
  if (device) *device = ctx->device;
  if (sm_count) *sm_count = ctx->sm_count;
  return 0;
}

int b200_alloc(b200_ctx* ctx, int mode, size_t bytes, void** host, void** dev) {
  B200_REQUIRE(ctx && host && dev, "b200_alloc: null argument");
  DeviceScope ds(ctx->device);
  *host = *dev = nullptr;
  if (bytes == 0) bytes = 1;
  if (mode == B200_OFFLOAD_UNIFIED) {
    // managed blocks are pooled too: cudaMallocManaged / cudaFree churn leaves
    // asynchronous UVM driver work (unmap, page-table teardown) that lands in the
    // next timed region
    const size_t cap = round_capacity_managed(bytes);
    void* p = cache_take(ctx->managed_cache, cap);
    if (!p) {
      cudaError_t e = cudaMallocManaged(&p, cap, cudaMemAttachGlobal);
      if (e != cudaSuccess) {
        cudaGetLastError();
        trim_caches(ctx);
        B200_CUDA(cudaMallocManaged(&p, cap, cudaMemAttachGlobal));
      }
    }
    ctx->live_managed.push_back({p, cap});
    *host = *dev = p;
    return 0;
  }
  B200_REQUIRE(mode == B200_OFFLOAD_ALWAYS || mode == B200_OFFLOAD_ONCE,
               "b200_alloc: unknown offload mode %d", mode);
  const size_t cap = round_capacity(bytes);
  void* alias = nullptr;
  void* h = cache_take(ctx->pinned_cache, cap, &alias);
  if (!h) {
    cudaError_t e = cudaMallocHost(&h, cap);
    if (e != cudaSuccess) {
      cudaGetLastError();
      trim_caches(ctx);
      B200_CUDA(cudaMallocHost(&h, cap));
    }
    // under UVA pinned memory is mapped into the device address space (zero-copy path)
    if (cudaHostGetDevicePointer(&alias, h, 0) != cudaSuccess) {
      alias = nullptr;
      cudaGetLastError();
    }
  }
  void* d = nullptr;
  if (take_device_block(ctx, cap, &d)) {
    // keep the pinned block (it may have come from the cache) for the next attempt
    cache_put(ctx->pinned_cache, h, cap, (size_t)6 << 30, [](void* p) { cudaFreeHost(p); }, alias);
    return 1;
  }
  ctx->live_host.push_back({h, cap, alias});
  ctx->live_dev.push_back({d, cap});
  *host = h;
  *dev = d;
  return 0;
}

int b200_free(b200_ctx* ctx, int mode, void* host, void* dev) {
  B200_REQUIRE(ctx, "b200_free: null context");
  DeviceScope ds(ctx->device);
  if (mode == B200_OFFLOAD_UNIFIED) {
    if (!host) return 0;
    if (ctx->compute_dirty || ctx->lane_dirty[0] || ctx->lane_dirty[1] || ctx->lane_dirty[2])
      if (b200_sync(ctx)) return 1;
    const size_t cap = live_pop(ctx->live_managed, host);
    if (cap)
      cache_put(ctx->managed_cache, host, cap, (size_t)24 << 30, [](void* p) { cudaFree(p); });
    else
      B200_CUDA(cudaFree(host));
    return 0;
  }
  // cached blocks may be handed out again immediately: queued work must be done
  if (ctx->compute_dirty || ctx->lane_dirty[0] || ctx->lane_dirty[1] || ctx->lane_dirty[2])
    if (b200_sync(ctx)) return 1;
  if (host) {
    void* alias = nullptr;
    const size_t cap = live_pop(ctx->live_host, host, &alias);
    if (cap)
      cache_put(ctx->pinned_cache, host, cap, (size_t)6 << 30, [](void* p) { cudaFreeHost(p); }, alias);
    else
      B200_CUDA(cudaFreeHost(host));
  }
  if (dev) {
    const size_t cap = live_pop(ctx->live_dev, dev);
    if (cap)
      cache_put(ctx->device_cache, dev, cap, (size_t)48 << 30, [](void* p) { cudaFree(p); });
    else
      B200_CUDA(cudaFree(dev));
  }
  return 0;
}

int b200_alloc_device(b200_ctx* ctx, size_t bytes, void** dev) {
  B200_REQUIRE(ctx && dev, "b200_alloc_device: null argument");
  DeviceScope ds(ctx->device);
  *dev = nullptr;
  return alloc_device(ctx, bytes, dev);
}

int b200_free_device(b200_ctx* ctx, void* dev) {
  B200_REQUIRE(ctx, "b200_free_device: null context");
  DeviceScope ds(ctx->device);
  if (ctx->compute_dirty || ctx->lane_dirty[0] || ctx->lane_dirty[1] || ctx->lane_dirty[2])
    if (b200_sync(ctx)) return 1;
  return free_device(ctx, dev);
}

int b200_h2d_async(b200_ctx* ctx, void* dev, const void* host, size_t bytes, int lane) {
  B200_REQUIRE(ctx && lane >= 0 && lane < B200_NUM_LANES, "b200_h2d_async: bad lane %d", lane);
  DeviceScope ds(ctx->device);
  if (bytes == 0) return 0;
  if (lanes_wait_for_compute(ctx)) return 1;
  B200_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->lane[lane]));
  B200_CUDA(cudaEventRecord(ctx->lane_done[lane], ctx->lane[lane]));
  ctx->lane_pending[lane] = true;
  ctx->lane_dirty[lane] = true;
  return 0;
}

int b200_d2h_async(b200_ctx* ctx, void* host, const void* dev, size_t bytes, int lane) {
  B200_REQUIRE(ctx && lane >= 0 && lane < B200_NUM_LANES, "b200_d2h_async: bad lane %d", lane);
  DeviceScope ds(ctx->device);
  if (bytes == 0) return 0;
  if (lanes_wait_for_compute(ctx)) return 1;
  B200_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->lane[lane]));
  // a later kernel may overwrite `dev`: it must wait for this copy to have read it (the
  // reference gets this from the legacy default stream)
  B200_CUDA(cudaEventRecord(ctx->lane_done[lane], ctx->lane[lane]));
  ctx->lane_pending[lane] = true;
  ctx->lane_dirty[lane] = true;
  return 0;
}

int b200_prefetch_async(b200_ctx* ctx, const void* managed, size_t bytes, int to_device, int lane) {
  B200_REQUIRE(ctx && lane >= 0 && lane < B200_NUM_LANES, "b200_prefetch_async: bad lane %d", lane);
  DeviceScope ds(ctx->device);
  if (bytes == 0) return 0;
  if (lanes_wait_for_compute(ctx)) return 1;
  B200_CUDA(mem_prefetch(managed, bytes, to_device ? ctx->device : -1, ctx->lane[lane]));
  B200_CUDA(cudaEventRecord(ctx->lane_done[lane], ctx->lane[lane]));
  ctx->lane_pending[lane] = true;
  ctx->lane_dirty[lane] = true;
  return 0;
}

int b200_advise(b200_ctx* ctx, const void* managed, size_t bytes, int is_input) {
  B200_REQUIRE(ctx && managed, "b200_advise: null argument");
  DeviceScope ds(ctx->device);
  if (bytes == 0) return 0;
  // Measured with the unmodified harness (scripts/unified_ab.sh, profiles/r1_unified_ab.txt):
  // on operands of a few MB the advice costs more than it saves (median 96 vs 63 us per
  // iteration, worse tails), from ~128 MB it is neutral.  It is therefore applied only to
  // large operands, where keeping the pages' home on the device protects them from
  // eviction; B200_ADVISE=all / none overrides.
  static const char* mode = getenv("B200_ADVISE");
  if (mode && !strcmp(mode, "none")) return 0;
  if (!(mode && !strcmp(mode, "all")) && bytes < ((size_t)64 << 20)) return 0;
  B200_CUDA(mem_advise(managed, bytes, cudaMemAdviseSetPreferredLocation, ctx->device));
  if (is_input) B200_CUDA(mem_advise(managed, bytes, cudaMemAdviseSetAccessedBy, -1));
  return 0;
}

int b200_sync(b200_ctx* ctx) {
  B200_REQUIRE(ctx, "b200_sync: null context");
  DeviceScope ds(ctx->device);
  B200_CUDA(cudaStreamSynchronize(ctx->compute));
  for (int l = 0; l < B200_NUM_LANES; l++) {
    if (ctx->lane_dirty[l]) B200_CUDA(cudaStreamSynchronize(ctx->lane[l]));
    ctx->lane_dirty[l] = false;
    ctx->lane_pending[l] = false;
  }
  ctx->compute_dirty = false;
  ctx->compute_unordered = false;
  return 0;
}

// ------------------------------------------------------------------ compute

static int check_gemm_args(const char* who, int ta, int tb, int m, int n, int k, const void* A,
                           int lda, const void* B, int ldb, const void* C, int ldc) {
  B200_REQUIRE((ta == 0 || ta == 1) && (tb == 0 || tb == 1), "%s: bad transpose flag", who);
  B200_REQUIRE(m >= 0 && n >= 0 && k >= 0, "%s: negative dimension (m=%d n=%d k=%d)", who, m, n, k);
  const int rowsA = ta ? k : m, rowsB = tb ? n : k;
  B200_REQUIRE(lda >= std::max(1, rowsA), "%s: lda=%d < max(1,%d)", who, lda, rowsA);
  B200_REQUIRE(ldb >= std::max(1, rowsB), "%s: ldb=%d < max(1,%d)", who, ldb, rowsB);
  B200_REQUIRE(ldc >= std::max(1, m), "%s: ldc=%d < max(1,%d)", who, ldc, m);
  if (m > 0 && n > 0)
    B200_REQUIRE(C != nullptr && (k == 0 || (A != nullptr && B != nullptr)), "%s: null operand", who);
  return 0;
}

int b200_dgemm(b200_ctx* ctx, int ta, int tb, int m, int n, int k, double alpha, const double* A,
               int lda, const double* B, int ldb, double beta, double* C, int ldc) {
  B200_REQUIRE(ctx, "b200_dgemm: null context");
  DeviceScope ds(ctx->device);
  ProfileRange pr("b200_dgemm %dx%dx%d", m, n, k);
  if (int r = check_gemm_args("b200_dgemm", ta, tb, m, n, k, A, lda, B, ldb, C, ldc)) return r;
  if (m == 0 || n == 0) return 0;
  if (join_lanes_into_compute(ctx)) return 1;
  if (!ta && !tb && k > 0 && alpha != 0.0) {
    int r = gemm_tallk_dispatch<double>(ctx, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
    if (r >= 0) return r;
    r = dgemm_dmma_dispatch(ctx, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
    if (r >= 0) return r;
  }
  return gemm_simt_dispatch<double>(ctx, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}

int b200_sgemm(b200_ctx* ctx, int ta, int tb, int m, int n, int k, float alpha, const float* A,
               int lda, const float* B, int ldb, float beta, float* C, int ldc) {
  B200_REQUIRE(ctx, "b200_sgemm: null context");
  DeviceScope ds(ctx->device);
  ProfileRange pr("b200_sgemm %dx%dx%d", m, n, k);
  if (int r = check_gemm_args("b200_sgemm", ta, tb, m, n, k, A, lda, B, ldb, C, ldc)) return r;
  if (m == 0 || n == 0) return 0;
  if (join_lanes_into_compute(ctx)) return 1;
  if (!ta && !tb && k > 0 && alpha != 0.0f) {
    int r = gemm_tallk_dispatch<float>(ctx, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
    if (r >= 0) return r;
    r = sgemm_tc_dispatch(ctx, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
    if (r >= 0) return r;
  }
  return gemm_simt_dispatch<float>(ctx, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}

int b200_sgemm_prepare(b200_ctx* ctx, int m, int n, int k, int lda, int ldb) {
  B200_REQUIRE(ctx, "b200_sgemm_prepare: null context");
  DeviceScope ds(ctx->device);
  if (m <= 0 || n <= 0 || k <= 0) return 0;
  return sgemm_tc_prepare(ctx, m, n, k, lda, ldb);
}

static int check_gemv_args(const char* who, int trans, int m, int n, const void* A, int lda,
                           const void* x, int incx, const void* y, int incy) {
  B200_REQUIRE(trans == 0 || trans == 1, "%s: bad transpose flag", who);
  B200_REQUIRE(m >= 0 && n >= 0, "%s: negative dimension (m=%d n=%d)", who, m, n);
  B200_REQUIRE(lda >= std::max(1, m), "%s: lda=%d < max(1,%d)", who, lda, m);
  B200_REQUIRE(incx != 0 && incy != 0, "%s: zero increment", who);
  if (m > 0 && n > 0) B200_REQUIRE(A && x && y, "%s: null operand", who);
  return 0;
}

int b200_dgemv(b200_ctx* ctx, int trans, int m, int n, double alpha, const double* A, int lda,
               const double* x, int incx, double beta, double* y, int incy) {
  B200_REQUIRE(ctx, "b200_dgemv: null context");
  DeviceScope ds(ctx->device);
  ProfileRange pr("b200_dgemv %dx%d", m, n);
  if (int r = check_gemv_args("b200_dgemv", trans, m, n, A, lda, x, incx, y, incy)) return r;
  if (join_lanes_into_compute(ctx)) return 1;
  return gemv_dispatch<double>(ctx, trans, m, n, alpha, A, lda, x, incx, beta, y, incy);
}

int b200_sgemv(b200_ctx* ctx, int trans, int m, int n, float alpha, const float* A, int lda,
               const float* x, int incx, float beta, float* y, int incy) {
  B200_REQUIRE(ctx, "b200_sgemv: null context");
  DeviceScope ds(ctx->device);
  ProfileRange pr("b200_sgemv %dx%d", m, n);
  if (int r = check_gemv_args("b200_sgemv", trans, m, n, A, lda, x, incx, y, incy)) return r;
  if (join_lanes_into_compute(ctx)) return 1;
  return gemv_dispatch<float>(ctx, trans, m, n, alpha, A, lda, x, incx, beta, y, incy);
}

}  // extern "C"

// ------------------------------------------------ Always-mode transfer engine

namespace b200 {

// Engine-private ordering: the slab / chunk pipelines know exactly which copy each
// kernel depends on, so they use their own events instead of the conservative
// "every lane waits for all earlier kernels" rule of b200_h2d_async (which would
// serialise upload j+1 behind kernel j).  Shared with the multi-GPU engine (mg.cu).
int engine_event(b200_ctx* ctx, cudaEvent_t* ev) {
  if (ctx->ev_pool_size == 0) {
    for (int i = 0; i < b200_ctx::kEvPool; i++)
      B200_CUDA(cudaEventCreateWithFlags(&ctx->ev_pool[i], cudaEventDisableTiming));
    ctx->ev_pool_size = b200_ctx::kEvPool;
  }
  *ev = ctx->ev_pool[ctx->ev_next];
  ctx->ev_next = (ctx->ev_next + 1) % ctx->ev_pool_size;
  return 0;
}

int engine_begin(b200_ctx* ctx) {
  // order this call after everything issued before it
  if (lanes_wait_for_compute(ctx)) return 1;
  if (join_lanes_into_compute(ctx)) return 1;
  return 0;
}

int record_on(b200_ctx* ctx, cudaStream_t s, cudaEvent_t* ev) {
  if (engine_event(ctx, ev)) return 1;
  B200_CUDA(cudaEventRecord(*ev, s));
  return 0;
}

}  // namespace b200

namespace {

// Bytes spanned by a column-major rows x cols block with leading dimension ld: the last
// column ends after `rows` elements, so a minimally sized caller buffer is never over-read.
inline size_t extent_bytes(size_t elem, int rows, int cols, int ld) {
  return cols > 0 && rows > 0 ? elem * ((size_t)ld * (cols - 1) + rows) : 0;
}

// Download a rows x cols block.  With ld > rows only the rows are written on the host: the
// padding between columns belongs to the caller (and, with beta == 0, holds nothing valid on
// the device).
inline cudaError_t download_block(void* host, const void* dev, size_t elem, int rows, int cols, int ld,
                                  cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  if (ld == rows || cols == 1)
    return cudaMemcpyAsync(host, dev, extent_bytes(elem, rows, cols, ld), cudaMemcpyDeviceToHost, st);
  return cudaMemcpy2DAsync(host, elem * ld, dev, elem * ld, elem * rows, cols, cudaMemcpyDeviceToHost, st);
}

template <typename T, typename GemmFn>
int gemm_offload(b200_ctx* ctx, int m, int n, int k, T alpha, const T* hA, T* dA, int lda,
                 const T* hB, T* dB, int ldb, T beta, T* hC, T* dC, int ldc, GemmFn gemm) {
  B200_REQUIRE(ctx && hA && dA && hB && dB && hC && dC, "gemm_offload: null argument");
  DeviceScope ds(ctx->device);
  ProfileRange pr("b200_%cgemm_offload %dx%dx%d", sizeof(T) == 8 ? 'd' : 's', m, n, k);
  B200_REQUIRE(m >= 0 && n >= 0 && k >= 0, "gemm_offload: negative dimension");
  if (m == 0 || n == 0) return 0;
  const size_t s = sizeof(T);
  const size_t bytesA = extent_bytes(s, m, k, lda), bytesB = extent_bytes(s, k, n, ldb),
               bytesC = extent_bytes(s, m, n, ldc);
  // B200_UPLOAD_C=1 reproduces the reference's third upload even though beta == 0 makes it dead
  const bool upload_c = beta != T(0) || ctx->opt_upload_c;

  // ---- small problems: one shot, fewest API calls (latency matters, not overlap).
  // C is only uploaded when beta != 0 -- with beta == 0 BLAS never reads it, so the
  // reference's third upload (cuBLAS/gemm.hh:130-131) is dead traffic.
  // (C is written over PCIe in 64-byte pieces by the small-tile kernels: beyond ~112 KiB of
  // output the one-shot path below is faster -- M=N=128..144, K=8..9: 38-44 us vs 36-40)
  if (bytesA + bytesB + bytesC <= ctx->opt_zerocopy_gemm_bytes && bytesC <= ((size_t)112 << 10) &&
      !ctx->opt_upload_c) {
    // ---- tiny problems: zero copy.  The kernel reads A and B straight from pinned host
    // memory over PCIe and writes C back the same way: one launch + one stream sync
    // instead of two uploads, a launch, a download and a sync.
    const T* zA = (const T*)device_alias(ctx, hA);
    const T* zB = (const T*)device_alias(ctx, hB);
    T* zC = (T*)device_alias(ctx, hC);
    if (zA && zB && zC) {
      if (engine_begin(ctx)) return 1;
      if (gemm(ctx, 0, 0, m, n, k, alpha, zA, lda, zB, ldb, beta, zC, ldc)) return 1;
      return b200_sync(ctx);
    }
  }
  if (bytesA + bytesB + bytesC < ((size_t)16 << 20) || k == 0) {
    if (b200_h2d_async(ctx, dA, hA, bytesA, 0)) return 1;
    if (b200_h2d_async(ctx, dB, hB, bytesB, 1)) return 1;
    if (upload_c)
      if (b200_h2d_async(ctx, dC, hC, bytesC, 2)) return 1;
    if (gemm(ctx, 0, 0, m, n, k, alpha, dA, lda, dB, ldb, beta, dC, ldc)) return 1;
    if (ldc == m) {
      if (b200_d2h_async(ctx, hC, dC, bytesC, 2)) return 1;
    } else {
      if (engine_begin(ctx)) return 1;
      B200_CUDA(download_block(hC, dC, s, m, n, ldc, ctx->lane[2]));
      ctx->lane_dirty[2] = true;
    }
    return b200_sync(ctx);
  }

  // ---- large problems: dependency-driven block pipeline.
  //   A is cut into nk K-chunks (contiguous column blocks of column-major A), B and C
  //   into ns column slabs (contiguous too); block (c, j) is
  //       C(:, j) (+)= A(:, Kc) * B(Kc, j)
  //   and needs only A_c and B_j on the device.  All uploads go down ONE lane in a
  //   fixed order, blocks are issued in the order their operands arrive, and a
  //   slab's download starts as soon as its last block is done.
  //   * compute-bound (DGEMM): uploads interleave A_0, B_0, A_1, B_1, ... so the GPU
  //     starts after 2 of 16 pieces and the rest of the upload hides behind compute;
  //   * transfer-bound (3xTF32 SGEMM): nk = 1, A first, then B slabs -- every slab's
  //     download overlaps the remaining uploads (PCIe is full duplex).
  //   Only contiguous copies are used: strided 2-D copies of B row-chunks measured
  //   ~10 GB/s against ~52 GB/s contiguous.
  if (engine_begin(ctx)) return 1;
  // B200_ENGINE_TRACE=1 prints a per-piece timeline (ms since the call) to stderr
  static const bool trace = getenv("B200_ENGINE_TRACE") != nullptr;
  struct Mark { cudaEvent_t ev; char what[24]; };
  std::vector<Mark> marks;
  auto mark = [&](cudaStream_t st, const char* fmt, int a, int b2) {
    if (!trace) return;
    Mark mk;
    cudaEventCreate(&mk.ev);
    cudaEventRecord(mk.ev, st);
    snprintf(mk.what, sizeof(mk.what), fmt, a, b2);
    marks.push_back(mk);
  };
  mark(ctx->lane[0], "start", 0, 0);
  const double flops = 2.0 * m * n * (double)k;
  const double rate = sizeof(T) == 8 ? 34e12 : ((lda % 4 == 0 && ldb % 4 == 0) ? 270e12 : 40e12);
  const bool compute_bound = flops / rate > (double)(bytesA + bytesB) / 50e9;
  // compute-bound: up to 8 x 8 pieces -- the first block starts after 2 of 16 pieces are up
  // (2.4 ms of a 19.4 ms upload at 8192^3) and only the last slab's download (1/8 of C) is
  // exposed at the end; 1024-wide blocks still run at ~0.95 of the big-GEMM rate (stream-K)
  int nk = compute_bound ? std::max(1, std::min(8, k / 1024)) : 1;
  int kc = ((k + nk - 1) / nk + 31) / 32 * 32;
  nk = (k + kc - 1) / kc;
  // column slabs of B and C: the last slab's GEMM and download are what is left exposed at the end
  // (and B_0 is what the first block waits for), so more, narrower slabs -- as long as a slab still
  // fills the machine (>= 512 columns: 256 DGEMM tiles with stream-K / 64 CTA-pair tiles)
  constexpr int kMaxK = 8, kMaxS = 16, kMax = kMaxS;
  static const char* slabs_env = getenv("B200_ENGINE_SLABS");  // tuning: upper bound on the slab count
  // measured, 8192^3 end to end: 8 slabs 36.59 ms (DGEMM) / 11.50 (SGEMM), 12 -> 35.97 / 11.32, 16 -> 36.27 / 11.45
  const int max_slabs = slabs_env ? std::max(1, std::min(kMaxS, atoi(slabs_env))) : 12;
  int ns = std::max(1, std::min(max_slabs, compute_bound ? n / 512 : n / 256));
  int cols = ((n + ns - 1) / ns + 127) / 128 * 128;
  ns = (n + cols - 1) / cols;
  static_assert(kMaxK <= kMax, "array sizes");
  cudaEvent_t evA[kMax], evB[kMax], ev;
  int posA[kMax], posB[kMax];
  bool waitedA[kMax] = {}, waitedB[kMax] = {};
  cudaStream_t up = ctx->lane[0], down = ctx->lane[2];
  if (upload_c) {
    B200_CUDA(cudaMemcpyAsync(dC, hC, bytesC, cudaMemcpyHostToDevice, up));
    if (record_on(ctx, up, &ev)) return 1;
    B200_CUDA(cudaStreamWaitEvent(ctx->compute, ev, 0));
  }
  auto upload_A = [&](int c, int pos) -> int {
    const int k0 = c * kc, kk = std::min(kc, k - k0);
    B200_CUDA(cudaMemcpyAsync(dA + (size_t)k0 * lda, hA + (size_t)k0 * lda, s * ((size_t)lda * (kk - 1) + m),
                              cudaMemcpyHostToDevice, up));
    posA[c] = pos;
    mark(up, "up A%d", c, 0);
    return record_on(ctx, up, &evA[c]);
  };
  auto upload_B = [&](int j, int pos) -> int {
    const int j0 = j * cols, nc = std::min(cols, n - j0);
    B200_CUDA(cudaMemcpyAsync(dB + (size_t)j0 * ldb, hB + (size_t)j0 * ldb, s * ((size_t)ldb * (nc - 1) + k),
                              cudaMemcpyHostToDevice, up));
    posB[j] = pos;
    mark(up, "up B%d", j, 0);
    return record_on(ctx, up, &evB[j]);
  };
  int pos = 0;
  if (compute_bound) {
    for (int i = 0; i < std::max(nk, ns); i++) {
      if (i < nk && upload_A(i, pos++)) return 1;
      if (i < ns && upload_B(i, pos++)) return 1;
    }
  } else {
    for (int c = 0; c < nk; c++)
      if (upload_A(c, pos++)) return 1;
    for (int j = 0; j < ns; j++)
      if (upload_B(j, pos++)) return 1;
  }
  // Block order by a small simulation of the pipeline: among the blocks whose operands
  // have (by estimate) already arrived, finish the lowest slab first so it downloads
  // while later slabs still compute; if nothing is ready, take the block that becomes
  // ready first.  (Issue order is execution order: the compute stream is FIFO.)
  struct Block { int c, j; bool issued; };
  Block blocks[kMax * kMax];
  int nb = 0, order[kMax * kMax];
  double start_time[kMax * kMax];
  bool mergeable[kMax * kMax] = {};
  for (int j = 0; j < ns; j++)
    for (int c = 0; c < nk; c++) blocks[nb++] = {c, j, false};
  {
    double arriveA[kMax], arriveB[kMax], t_up = 0.0;
    // replay the upload order to get arrival-time estimates
    struct Piece { bool isA; int idx; };
    Piece seq[2 * kMax];
    for (int c = 0; c < nk; c++) seq[posA[c]] = {true, c};
    for (int j = 0; j < ns; j++) seq[posB[j]] = {false, j};
    for (int i = 0; i < pos; i++) {
      const Piece& pc = seq[i];
      const int k0 = pc.idx * kc, kk = std::min(kc, k - k0);
      const int j0 = pc.idx * cols, nc = std::min(cols, n - j0);
      t_up += (pc.isA ? (double)s * lda * kk : (double)s * ldb * nc) / 55e9;
      (pc.isA ? arriveA : arriveB)[pc.idx] = t_up;
    }
    double now = 0.0;
    for (int i = 0; i < nb; i++) {
      start_time[i] = 0.0;
      int best = -1;
      double best_ready = 0.0;
      for (int b = 0; b < nb; b++) {
        if (blocks[b].issued) continue;
        const double ready = std::max(arriveA[blocks[b].c], arriveB[blocks[b].j]);
        if (ready <= now) { best = b; break; }  // blocks[] is slab-major: first ready = lowest slab
        if (best < 0 || ready < best_ready) { best = b; best_ready = ready; }
      }
      const Block& bl = blocks[best];
      const int kk = std::min(kc, k - bl.c * kc), nc = std::min(cols, n - bl.j * cols);
      start_time[i] = std::max(now, std::max(arriveA[bl.c], arriveB[bl.j]));
      now = start_time[i] + 2.0 * m * nc * (double)kk / rate;
      blocks[best].issued = true;
      order[i] = best;
    }
    // A run of consecutive K-chunks of one slab whose operands are all up before the run
    // starts is ONE GEMM over the joined K range (the late slabs of a compute-bound problem:
    // one 8192 x 1024 x 8192 call instead of eight x1024 accumulate blocks -- longer k loops,
    // one pass over C).  mergeable[i] = block order[i] continues the run of order[i-1].
    for (int i = 1; i < nb; i++) {
      const Block &p = blocks[order[i - 1]], &q = blocks[order[i]];
      int first = i - 1;
      while (first > 0 && mergeable[first]) first--;
      mergeable[i] = q.j == p.j && q.c == p.c + 1 && arriveA[q.c] <= start_time[first] &&
                     arriveB[q.j] <= start_time[first];
    }
  }
  int done_in_slab[kMax] = {};
  for (int i = 0; i < nb;) {
    int last = i;  // blocks order[i..last] form one GEMM call
    while (last + 1 < nb && mergeable[last + 1]) last++;
    const int c0 = blocks[order[i]].c, c1 = blocks[order[last]].c, j = blocks[order[i]].j;
    const int k0 = c0 * kc, kk = std::min(k, (c1 + 1) * kc) - k0;
    const int j0 = j * cols, nc = std::min(cols, n - j0);
    for (int c = c0; c <= c1; c++)
      if (!waitedA[c]) { B200_CUDA(cudaStreamWaitEvent(ctx->compute, evA[c], 0)); waitedA[c] = true; }
    if (!waitedB[j]) { B200_CUDA(cudaStreamWaitEvent(ctx->compute, evB[j], 0)); waitedB[j] = true; }
    if (gemm(ctx, 0, 0, m, nc, kk, alpha, dA + (size_t)k0 * lda, lda, dB + k0 + (size_t)j0 * ldb, ldb,
             done_in_slab[j] == 0 ? beta : T(1), dC + (size_t)j0 * ldc, ldc))
      return 1;
    mark(ctx->compute, c0 == c1 ? "gemm A%d B%d" : "gemm A%d.. B%d", c0, j);
    done_in_slab[j] += c1 - c0 + 1;
    if (done_in_slab[j] == nk) {
      if (record_on(ctx, ctx->compute, &ev)) return 1;
      B200_CUDA(cudaStreamWaitEvent(down, ev, 0));
      B200_CUDA(download_block(hC + (size_t)j0 * ldc, dC + (size_t)j0 * ldc, s, m, nc, ldc, down));
      mark(down, "down C%d", j, 0);
    }
    i = last + 1;
  }
  ctx->compute_unordered = false;  // every kernel above is already ordered before its download
  for (int l = 0; l < B200_NUM_LANES; l++) ctx->lane_dirty[l] = true;
  const int rc = b200_sync(ctx);
  if (trace && !marks.empty()) {
    fprintf(stderr, "[b200 engine] m=%d n=%d k=%d nk=%d ns=%d %s\n", m, n, k, nk, ns,
            compute_bound ? "compute-bound" : "transfer-bound");
    for (size_t i = 1; i < marks.size(); i++) {
      float ms = 0;
      cudaEventElapsedTime(&ms, marks[0].ev, marks[i].ev);
      fprintf(stderr, "[b200 engine] %8.3f ms  %s\n", ms, marks[i].what);
    }
    for (auto& mk : marks) cudaEventDestroy(mk.ev);
  }
  return rc;
}

template <typename T, typename GemvFn>
int gemv_offload(b200_ctx* ctx, int m, int n, T alpha, const T* hA, T* dA, int lda, const T* hx,
                 T* dx, T beta, T* hy, T* dy, GemvFn gemv) {
  B200_REQUIRE(ctx && hA && dA && hx && dx && hy && dy, "gemv_offload: null argument");
  DeviceScope ds(ctx->device);
  ProfileRange pr("b200_%cgemv_offload %dx%d", sizeof(T) == 8 ? 'd' : 's', m, n);
  B200_REQUIRE(m >= 0 && n >= 0, "gemv_offload: negative dimension");
  if (m == 0) return 0;
  const size_t s = sizeof(T);
  const size_t bytesA = extent_bytes(s, m, n, lda);
  const bool upload_y = beta != T(0) || ctx->opt_upload_c;
  if (bytesA <= ctx->opt_zerocopy_gemv_bytes && !ctx->opt_upload_c) {  // zero copy, see gemm_offload
    const T* zA = (const T*)device_alias(ctx, hA);
    const T* zx = (const T*)device_alias(ctx, hx);
    T* zy = (T*)device_alias(ctx, hy);
    if (zA && zx && zy) {
      if (engine_begin(ctx)) return 1;
      if (gemv(ctx, 0, m, n, alpha, zA, lda, zx, 1, beta, zy, 1)) return 1;
      return b200_sync(ctx);
    }
  }
  if (bytesA < ((size_t)8 << 20) || n < 2) {  // one shot
    if (b200_h2d_async(ctx, dA, hA, bytesA, 0)) return 1;
    if (b200_h2d_async(ctx, dx, hx, s * (size_t)n, 1)) return 1;
    if (upload_y)
      if (b200_h2d_async(ctx, dy, hy, s * (size_t)m, 2)) return 1;
    if (gemv(ctx, 0, m, n, alpha, dA, lda, dx, 1, beta, dy, 1)) return 1;
    if (b200_d2h_async(ctx, hy, dy, s * (size_t)m, 2)) return 1;
    return b200_sync(ctx);
  }
  // PCIe-bound: stream A up in column slabs, y accumulates (beta = 1 after the first
  // slab) so slab j's kernel hides behind slab j+1's upload.
  if (engine_begin(ctx)) return 1;
  cudaEvent_t ev;
  B200_CUDA(cudaMemcpyAsync(dx, hx, s * (size_t)n, cudaMemcpyHostToDevice, ctx->lane[2]));
  if (upload_y) B200_CUDA(cudaMemcpyAsync(dy, hy, s * (size_t)m, cudaMemcpyHostToDevice, ctx->lane[2]));
  if (record_on(ctx, ctx->lane[2], &ev)) return 1;
  B200_CUDA(cudaStreamWaitEvent(ctx->compute, ev, 0));
  const int slabs = (int)std::max<size_t>(2, std::min<size_t>(16, bytesA / ((size_t)16 << 20)));
  const int cols = (n + slabs - 1) / slabs;
  for (int j0 = 0, i = 0; j0 < n; j0 += cols, i++) {
    const int nc = std::min(cols, n - j0);
    cudaStream_t lane = ctx->lane[i & 1];
    B200_CUDA(cudaMemcpyAsync(dA + (size_t)j0 * lda, hA + (size_t)j0 * lda, s * ((size_t)lda * (nc - 1) + m),
                              cudaMemcpyHostToDevice, lane));
    if (record_on(ctx, lane, &ev)) return 1;
    B200_CUDA(cudaStreamWaitEvent(ctx->compute, ev, 0));
    if (gemv(ctx, 0, m, nc, alpha, dA + (size_t)j0 * lda, lda, dx + j0, 1, i == 0 ? beta : T(1), dy, 1))
      return 1;
  }
  if (record_on(ctx, ctx->compute, &ev)) return 1;
  B200_CUDA(cudaStreamWaitEvent(ctx->lane[2], ev, 0));
  B200_CUDA(cudaMemcpyAsync(hy, dy, s * (size_t)m, cudaMemcpyDeviceToHost, ctx->lane[2]));
  ctx->compute_unordered = false;
  for (int l = 0; l < B200_NUM_LANES; l++) ctx->lane_dirty[l] = true;
  return b200_sync(ctx);
}

}  // namespace

extern "C" {

int b200_dgemm_offload(b200_ctx* ctx, int m, int n, int k, double alpha, const double* hA, double* dA,
                       int lda, const double* hB, double* dB, int ldb, double beta, double* hC,
                       double* dC, int ldc) {
  return gemm_offload<double>(ctx, m, n, k, alpha, hA, dA, lda, hB, dB, ldb, beta, hC, dC, ldc, b200_dgemm);
}
int b200_sgemm_offload(b200_ctx* ctx, int m, int n, int k, float alpha, const float* hA, float* dA,
                       int lda, const float* hB, float* dB, int ldb, float beta, float* hC, float* dC,
                       int ldc) {
  return gemm_offload<float>(ctx, m, n, k, alpha, hA, dA, lda, hB, dB, ldb, beta, hC, dC, ldc, b200_sgemm);
}
int b200_dgemv_offload(b200_ctx* ctx, int m, int n, double alpha, const double* hA, double* dA, int lda,
                       const double* hx, double* dx, double beta, double* hy, double* dy) {
  return gemv_offload<double>(ctx, m, n, alpha, hA, dA, lda, hx, dx, beta, hy, dy, b200_dgemv);
}
int b200_sgemv_offload(b200_ctx* ctx, int m, int n, float alpha, const float* hA, float* dA, int lda,
                       const float* hx, float* dx, float beta, float* hy, float* dy) {
  return gemv_offload<float>(ctx, m, n, alpha, hA, dA, lda, hx, dx, beta, hy, dy, b200_sgemv);
}

// ------------------------------------------------------------ introspection

const char* b200_last_kernel(const b200_ctx* ctx) { return ctx ? ctx->last_kernel : "none"; }
unsigned long long b200_launch_count(const b200_ctx* ctx) { return ctx ? ctx->launches : 0; }

int b200_timer_start(b200_ctx* ctx) {
  B200_REQUIRE(ctx, "b200_timer_start: null context");
  DeviceScope ds(ctx->device);
  B200_CUDA(cudaEventRecord(ctx->t0, ctx->compute));
  return 0;
}
int b200_timer_stop(b200_ctx* ctx, float* ms) {
  B200_REQUIRE(ctx && ms, "b200_timer_stop: null argument");
  DeviceScope ds(ctx->device);
  B200_CUDA(cudaEventRecord(ctx->t1, ctx->compute));
  B200_CUDA(cudaEventSynchronize(ctx->t1));
  B200_CUDA(cudaEventElapsedTime(ms, ctx->t0, ctx->t1));
  return 0;
}

int b200_probe_peak(b200_ctx* ctx, const char* which, double* result) {
  B200_REQUIRE(ctx, "b200_probe_peak: null context");
  DeviceScope ds(ctx->device);
  return probe_peak(ctx, which, result);
}

}  // extern "C"
