// gemv.cu -- SGEMV / DGEMV for column-major A, hand-written for sm_100a.
//
// Replaces cublasSgemv / cublasDgemv at reference cuBLAS/gemv.hh:130-137,
// 148-155,161-168 (OP_N, incx = incy = 1 there; OP_T and general increments are
// implemented too).
//
// The operation is HBM-bound: A (m*n*s bytes) is streamed exactly once, x and
// y are noise.  Design (OP_N):
//   * consecutive threads own consecutive rows, so a warp reads a contiguous
//     run of one column: 128-bit loads when lda allows (VEC = 16/s), else
//     64-/32-bit, always fully coalesced in column-major -- no transposition.
//   * a CTA is TX row-threads x TY column-lanes (TX*TY = 256); each thread
//     keeps UNROLL independent 128-bit loads in flight (~32 KB per CTA) so a
//     few CTAs per SM cover the HBM latency-bandwidth product.
//   * x for the CTA's column range is staged in shared memory once.
//   * the column range is split across blockIdx.y so short-wide shapes
//     (m = 32, n = 32768) still fill 148 SMs; partial sums go to a workspace
//     and the LAST CTA of each row block (arrival counter) adds them in a fixed
//     order -- one launch, deterministic, no atomics on y.
//   * TY column-lane partials are combined through shared memory; the many
//     independent partial sums keep FP32 rounding error ~1e-7 at n = 32768.
// OP_T: one warp per column, lanes stride down the column with the same
// vector loads, warp-shuffle reduction.
#include "common.cuh"

namespace b200 {
namespace {

constexpr int kThreads = 256;
constexpr int kXChunk = 2048;  // x elements staged in smem per pass
constexpr int kUnroll = 8;

// Streaming (evict-first) vector loads: A is read exactly once.  Plain
// intrinsics rather than inline PTX so the compiler is free to hoist all
// kUnroll loads of an iteration above their uses (memory-level parallelism).
template <typename T, int VEC>
struct VecLoad;
template <>
struct VecLoad<float, 4> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
    const float4 t = __ldcs(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};
template <>
struct VecLoad<float, 2> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[2]) {
    const float2 t = __ldcs(reinterpret_cast<const float2*>(p));
    v[0] = t.x; v[1] = t.y;
  }
};
template <>
struct VecLoad<float, 1> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[1]) { v[0] = __ldcs(p); }
};
template <>
struct VecLoad<double, 2> {
  static __device__ __forceinline__ void ld(const double* p, double (&v)[2]) {
    const double2 t = __ldcs(reinterpret_cast<const double2*>(p));
    v[0] = t.x; v[1] = t.y;
  }
};
template <>
struct VecLoad<double, 1> {
  static __device__ __forceinline__ void ld(const double* p, double (&v)[1]) { v[0] = __ldcs(p); }
};

// ---------------------------------------------------------------- OP_N ----
// grid.x = row blocks, grid.y = column splits.
template <typename T, int VEC, int TX>
__global__ void __launch_bounds__(kThreads)
gemv_n_kernel(int m, int n, T alpha, const T* __restrict__ A, long long lda,
              const T* __restrict__ x, long long incx, T beta, T* __restrict__ y,
              long long incy, int cols_per_split, T* __restrict__ partial,
              unsigned int* __restrict__ counters) {
  constexpr int TY = kThreads / TX;
  constexpr int ROWS = TX * VEC;
  __shared__ T xs[kXChunk];
  __shared__ T red[TY][ROWS + 1];
  __shared__ bool is_last;

  const int tx = threadIdx.x % TX;
  const int ty = threadIdx.x / TX;
  const int row = blockIdx.x * ROWS + tx * VEC;
  const int c0 = blockIdx.y * cols_per_split;
  const int c1 = min(n, c0 + cols_per_split);
  const int nvalid = min(VEC, m - row);  // <= 0: thread has no rows

  T acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; v++) acc[v] = T(0);

  for (int cb = c0; cb < c1; cb += kXChunk) {
    const int clen = min(kXChunk, c1 - cb);
    __syncthreads();
    for (int i = threadIdx.x; i < clen; i += kThreads)
      xs[i] = x[(long long)(cb + i) * incx];
    __syncthreads();

    if (nvalid == VEC) {
      const T* ap = A + (long long)cb * lda + row;
      int j = ty;
      // main loop: kUnroll independent vector loads in flight per thread
      for (; j + (kUnroll - 1) * TY < clen; j += kUnroll * TY) {
        T a[kUnroll][VEC];
#pragma unroll
        for (int u = 0; u < kUnroll; u++)
          VecLoad<T, VEC>::ld(ap + (long long)(j + u * TY) * lda, a[u]);
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
          const T xv = xs[j + u * TY];
#pragma unroll
          for (int v = 0; v < VEC; v++) acc[v] = fma(a[u][v], xv, acc[v]);
        }
      }
      for (; j < clen; j += TY) {
        T a[VEC];
        VecLoad<T, VEC>::ld(ap + (long long)j * lda, a);
        const T xv = xs[j];
#pragma unroll
        for (int v = 0; v < VEC; v++) acc[v] = fma(a[v], xv, acc[v]);
      }
    } else if (nvalid > 0) {
      // ragged last rows of the matrix: scalar loads
      const T* ap = A + (long long)cb * lda + row;
      for (int j = ty; j < clen; j += TY) {
        const T xv = xs[j];
#pragma unroll
        for (int v = 0; v < VEC; v++)
          if (v < nvalid) acc[v] = fma(__ldg(ap + (long long)j * lda + v), xv, acc[v]);
      }
    }
  }

  // combine the TY column lanes
#pragma unroll
  for (int v = 0; v < VEC; v++) red[ty][tx * VEC + v] = acc[v];
  __syncthreads();

  const int r = threadIdx.x;  // one thread per row of the block from here on
  const int grow = blockIdx.x * ROWS + r;
  T sum = T(0);
  if (r < ROWS && grow < m) {
#pragma unroll
    for (int t = 0; t < TY; t++) sum += red[t][r];
  }

  if (gridDim.y == 1) {
    if (r < ROWS && grow < m) {
      T* yp = y + (long long)grow * incy;
      *yp = (beta == T(0)) ? alpha * sum : alpha * sum + beta * *yp;
    }
    return;
  }

  // split-N: publish the partial, last CTA of this row block finishes
  const long long mpad = (long long)gridDim.x * ROWS;
  if (r < ROWS) partial[(long long)blockIdx.y * mpad + blockIdx.x * ROWS + r] = sum;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(&counters[blockIdx.x], 1u);
    is_last = (prev == gridDim.y - 1);
    if (is_last) counters[blockIdx.x] = 0;  // self-reset for the next launch
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // All 256 threads take part: G = 256 / ROWS threads per row, thread g adds splits
  // g, g + G, ... in ascending order with 8 independent loads in flight per step (a plain
  // loop is one dependent L2 round trip per split), then the G partial totals of a row
  // are added in order of g -- a fixed order, so the result is deterministic.
  constexpr int G = kThreads / ROWS;
  const int rr = threadIdx.x % ROWS, g = threadIdx.x / ROWS;
  {
    const T* src = partial + blockIdx.x * ROWS + rr;
    const int nsplit = (int)gridDim.y;
    T total = T(0);
    int s = g;
    for (; s + 7 * G < nsplit; s += 8 * G) {
      T v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) v[u] = __ldcg(src + (long long)(s + u * G) * mpad);
#pragma unroll
      for (int u = 0; u < 8; u++) total += v[u];
    }
    for (; s < nsplit; s += G) total += __ldcg(src + (long long)s * mpad);
    red[g][rr] = total;
  }
  __syncthreads();
  if (g == 0 && blockIdx.x * ROWS + rr < m) {
    T total = T(0);
#pragma unroll
    for (int t = 0; t < G; t++) total += red[t][rr];
    T* yp = y + (long long)(blockIdx.x * ROWS + rr) * incy;
    *yp = (beta == T(0)) ? alpha * total : alpha * total + beta * *yp;
  }
}

// ---------------------------------------------------------------- OP_T ----
// y(n) = alpha * A^T x + beta * y : one warp per column.
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
gemv_t_kernel(int m, int n, T alpha, const T* __restrict__ A, long long lda,
              const T* __restrict__ x, long long incx, T beta, T* __restrict__ y,
              long long incy) {
  const int lane = threadIdx.x & 31;
  const int col = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (col >= n) return;
  const T* ap = A + (long long)col * lda;
  T acc[4] = {T(0), T(0), T(0), T(0)};
  const int mvec = (m / VEC) * VEC;
  int i = lane * VEC;
  for (; i + 3 * 32 * VEC < mvec; i += 4 * 32 * VEC) {
    T a[4][VEC];
#pragma unroll
    for (int u = 0; u < 4; u++) VecLoad<T, VEC>::ld(ap + i + u * 32 * VEC, a[u]);
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int v = 0; v < VEC; v++)
        acc[u] = fma(a[u][v], __ldg(x + (long long)(i + u * 32 * VEC + v) * incx), acc[u]);
  }
  for (; i < mvec; i += 32 * VEC) {
    T a[VEC];
    VecLoad<T, VEC>::ld(ap + i, a);
#pragma unroll
    for (int v = 0; v < VEC; v++)
      acc[0] = fma(a[v], __ldg(x + (long long)(i + v) * incx), acc[0]);
  }
  for (int t = mvec + lane; t < m; t += 32)
    acc[1] = fma(__ldg(ap + t), __ldg(x + (long long)t * incx), acc[1]);
  T s = warp_sum((acc[0] + acc[1]) + (acc[2] + acc[3]));
  if (lane == 0) {
    T* yp = y + (long long)col * incy;
    *yp = (beta == T(0)) ? alpha * s : alpha * s + beta * *yp;
  }
}

template <typename T>
__global__ void scale_vec_kernel(int len, T beta, T* y, long long incy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < len) {
    T* yp = y + (long long)i * incy;
    *yp = (beta == T(0)) ? T(0) : beta * *yp;
  }
}

template <typename T, int VEC, int TX>
int launch_n(b200_ctx* ctx, int m, int n, T alpha, const T* A, int lda, const T* x,
             long long incx, T beta, T* y, long long incy, const char* name) {
  constexpr int ROWS = TX * VEC;
  const int row_blocks = (m + ROWS - 1) / ROWS;
  // aim for ~4 resident CTAs per SM; a split must be wide enough for the unrolled main
  // loop (kUnroll loads in flight per thread) to run at least once
  int splits = 1;
  const int target = ctx->sm_count * 4;
  if (row_blocks < target && row_blocks <= ctx->num_counters) {
    splits = (target + row_blocks - 1) / row_blocks;
    constexpr int min_cols = (kUnroll * (kThreads / TX)) > 64 ? kUnroll * (kThreads / TX) : 64;
    const int max_splits = (n + min_cols - 1) / min_cols;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
  }
  int cols_per_split = (n + splits - 1) / splits;
  splits = (n + cols_per_split - 1) / cols_per_split;
  T* partial = nullptr;
  if (splits > 1) {
    const size_t need = (size_t)splits * row_blocks * ROWS * sizeof(T);
    if (ensure_workspace(ctx, need)) return 1;
    partial = (T*)ctx->workspace;
  }
  dim3 grid(row_blocks, splits);
  gemv_n_kernel<T, VEC, TX><<<grid, kThreads, 0, ctx->compute>>>(
      m, n, alpha, A, (long long)lda, x, incx, beta, y, incy, cols_per_split,
      partial, ctx->counters);
  B200_CUDA(cudaGetLastError());
  note_launch(ctx, name);
  return 0;
}

template <typename T, int VEC>
int launch_n_tx(b200_ctx* ctx, int m, int n, T alpha, const T* A, int lda,
                const T* x, long long incx, T beta, T* y, long long incy,
                const char* n8, const char* n16, const char* n32, const char* n64) {
  if (m <= 8 * VEC)
    return launch_n<T, VEC, 8>(ctx, m, n, alpha, A, lda, x, incx, beta, y, incy, n8);
  if (m <= 16 * VEC)
    return launch_n<T, VEC, 16>(ctx, m, n, alpha, A, lda, x, incx, beta, y, incy, n16);
  if (m <= 32 * VEC || m < 2048)
    return launch_n<T, VEC, 32>(ctx, m, n, alpha, A, lda, x, incx, beta, y, incy, n32);
  return launch_n<T, VEC, 64>(ctx, m, n, alpha, A, lda, x, incx, beta, y, incy, n64);
}

template <typename T>
int max_vec(const T* A, int lda) {
  const int full = 16 / (int)sizeof(T);
  for (int v = full; v > 1; v >>= 1)
    if (((uintptr_t)A % (v * sizeof(T))) == 0 && (lda % v) == 0) return v;
  return 1;
}

}  // namespace

template <typename T>
int gemv_dispatch(b200_ctx* ctx, int trans, int m, int n, T alpha, const T* A,
                  int lda, const T* x, int incx, T beta, T* y, int incy) {
  const int leny = trans ? n : m;
  const int lenx = trans ? m : n;
  if (leny == 0) return 0;
  // BLAS convention for negative increments: start from the far end
  const T* xb = x + (incx < 0 ? (long long)(1 - lenx) * incx : 0);
  T* yb = y + (incy < 0 ? (long long)(1 - leny) * incy : 0);

  if (lenx == 0 || alpha == T(0)) {  // y <- beta * y
    scale_vec_kernel<T><<<(leny + 255) / 256, 256, 0, ctx->compute>>>(leny, beta, yb, incy);
    B200_CUDA(cudaGetLastError());
    note_launch(ctx, "gemv_scale_only");
    return 0;
  }
  const int vec = max_vec(A, lda);
  if (!trans) {
    if constexpr (sizeof(T) == 4) {
      if (vec == 4)
        return launch_n_tx<T, 4>(ctx, m, n, alpha, A, lda, xb, incx, beta, yb, incy,
                                 "sgemv_n_v4_tx8", "sgemv_n_v4_tx16",
                                 "sgemv_n_v4_tx32", "sgemv_n_v4_tx64");
      if (vec == 2)
        return launch_n_tx<T, 2>(ctx, m, n, alpha, A, lda, xb, incx, beta, yb, incy,
                                 "sgemv_n_v2_tx8", "sgemv_n_v2_tx16",
                                 "sgemv_n_v2_tx32", "sgemv_n_v2_tx64");
      return launch_n_tx<T, 1>(ctx, m, n, alpha, A, lda, xb, incx, beta, yb, incy,
                               "sgemv_n_v1_tx8", "sgemv_n_v1_tx16",
                               "sgemv_n_v1_tx32", "sgemv_n_v1_tx64");
    } else {
      if (vec == 2)
        return launch_n_tx<T, 2>(ctx, m, n, alpha, A, lda, xb, incx, beta, yb, incy,
                                 "dgemv_n_v2_tx8", "dgemv_n_v2_tx16",
                                 "dgemv_n_v2_tx32", "dgemv_n_v2_tx64");
      return launch_n_tx<T, 1>(ctx, m, n, alpha, A, lda, xb, incx, beta, yb, incy,
                               "dgemv_n_v1_tx8", "dgemv_n_v1_tx16",
                               "dgemv_n_v1_tx32", "dgemv_n_v1_tx64");
    }
  }
  // OP_T
  const int grid = (n + kThreads / 32 - 1) / (kThreads / 32);
  if constexpr (sizeof(T) == 4) {
    if (vec == 4)
      gemv_t_kernel<T, 4><<<grid, kThreads, 0, ctx->compute>>>(m, n, alpha, A, lda, xb, incx, beta, yb, incy);
    else if (vec == 2)
      gemv_t_kernel<T, 2><<<grid, kThreads, 0, ctx->compute>>>(m, n, alpha, A, lda, xb, incx, beta, yb, incy);
    else
      gemv_t_kernel<T, 1><<<grid, kThreads, 0, ctx->compute>>>(m, n, alpha, A, lda, xb, incx, beta, yb, incy);
    B200_CUDA(cudaGetLastError());
    note_launch(ctx, "sgemv_t_warp_per_col");
  } else {
    if (vec == 2)
      gemv_t_kernel<T, 2><<<grid, kThreads, 0, ctx->compute>>>(m, n, alpha, A, lda, xb, incx, beta, yb, incy);
    else
      gemv_t_kernel<T, 1><<<grid, kThreads, 0, ctx->compute>>>(m, n, alpha, A, lda, xb, incx, beta, yb, incy);
    B200_CUDA(cudaGetLastError());
    note_launch(ctx, "dgemv_t_warp_per_col");
  }
  return 0;
}

template int gemv_dispatch<float>(b200_ctx*, int, int, int, float, const float*, int,
                                  const float*, int, float, float*, int);
template int gemv_dispatch<double>(b200_ctx*, int, int, int, double, const double*,
                                   int, const double*, int, double, double*, int);

}  // namespace b200
