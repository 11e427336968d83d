// probe.cu -- peak probes: the roofline denominators for DGEMM / SGEMM (FP64
// and FP32 pipe peaks are not in MEASURED_PEAKS.json, which only holds the HBM
// copy bandwidth and the bf16 tensor peak; SURVEY.md 9.6 Q3 asks for them to be
// measured on the box) and a copy-bandwidth cross-check.
#include <string.h>

#include "common.cuh"

namespace b200 {
namespace {

constexpr int kIters = 4096;

__global__ void __launch_bounds__(256) probe_dfma(double* out, double seed) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = seed + i + threadIdx.x;
  const double x = 1.0000001, y = 1e-9;
  for (int it = 0; it < kIters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = fma(a[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += a[i];
  if (s == 123.456) out[0] = s;  // never true; keeps the chain alive
}

__global__ void __launch_bounds__(256) probe_ffma(float* out, float seed) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = seed + i + threadIdx.x;
  const float x = 1.0000001f, y = 1e-9f;
  for (int it = 0; it < kIters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], x, y);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += a[i];
  if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(256) probe_dmma(double* out, double seed) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = seed + i;
  const double a = 1.0000001 + threadIdx.x * 1e-12, b = 1e-9;
  for (int it = 0; it < kIters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile(
          "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
          : "+d"(c[i][0]), "+d"(c[i][1])
          : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) probe_copy(const float4* __restrict__ src,
                                                  float4* __restrict__ dst, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    float4 v0 = __ldcs(src + i), v1 = __ldcs(src + i + stride),
           v2 = __ldcs(src + i + 2 * stride), v3 = __ldcs(src + i + 3 * stride);
    __stcs(dst + i, v0);
    __stcs(dst + i + stride, v1);
    __stcs(dst + i + 2 * stride, v2);
    __stcs(dst + i + 3 * stride, v3);
  }
  for (; i < n; i += stride) __stcs(dst + i, __ldcs(src + i));
}

}  // namespace

int probe_peak(b200_ctx* ctx, const char* which, double* result) {
  B200_REQUIRE(which && result, "probe_peak: null argument");
  cudaStream_t s = ctx->compute;
  cudaEvent_t e0, e1;
  B200_CUDA(cudaEventCreate(&e0));
  B200_CUDA(cudaEventCreate(&e1));
  const int blocks = ctx->sm_count * 8, threads = 256;
  float best_ms = 1e30f;
  double work = 0;  // flops or bytes per launch
  void* buf = nullptr;
  const size_t copy_bytes = (size_t)1 << 30;
  if (!strcmp(which, "copy")) {
    B200_CUDA(cudaMalloc(&buf, 2 * copy_bytes));
    B200_CUDA(cudaMemsetAsync(buf, 1, 2 * copy_bytes, s));
  } else {
    if (ensure_workspace(ctx, 4096)) return 1;
  }
  for (int rep = 0; rep < 6; rep++) {
    B200_CUDA(cudaEventRecord(e0, s));
    if (!strcmp(which, "dfma")) {
      probe_dfma<<<blocks, threads, 0, s>>>((double*)ctx->workspace, 1.0);
      work = 2.0 * 8 * kIters * (double)blocks * threads;
    } else if (!strcmp(which, "ffma")) {
      probe_ffma<<<blocks, threads, 0, s>>>((float*)ctx->workspace, 1.0f);
      work = 2.0 * 16 * kIters * (double)blocks * threads;
    } else if (!strcmp(which, "dmma")) {
      probe_dmma<<<blocks, threads, 0, s>>>((double*)ctx->workspace, 1.0);
      work = 2.0 * 8 * 4 * 8 * 8 * kIters * (double)blocks * (threads / 32);
    } else if (!strcmp(which, "copy")) {
      probe_copy<<<ctx->sm_count * 16, threads, 0, s>>>(
          (const float4*)buf, (float4*)((char*)buf + copy_bytes), copy_bytes / 16);
      work = 2.0 * (double)copy_bytes;
    } else {
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
      set_error("probe_peak: unknown probe '%s'", which);
      return 2;
    }
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaEventRecord(e1, s));
    B200_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    B200_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best_ms) best_ms = ms;
    ctx->launches++;
  }
  if (buf) cudaFree(buf);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *result = work / (best_ms * 1e-3) * 1e-9;
  return 0;
}

}  // namespace b200
