// gemm_tallk.cu -- SGEMM / DGEMM for one small output tile and a long K: the reference's
// "short-wide x tall-thin" problem type M = N = 32, K = d (include/doGemm.hh:171-201; the
// cublasGemmEx calls at cuBLAS/gemm.hh:134-146 with m = n = 32).
//
// The whole problem is one 32 x 32 tile of C and two K x 32 operands (4 MB of FP64 at K = 8192):
// launch latency, one DRAM round trip and the cross-CTA reduction are all there is.  The
// split-K kernels (dgemm_dmma.cu, gemm_simt.cu) spread K over <= 48 CTAs -- too few loads in
// flight to cover the DRAM latency -- and let ONE CTA add the 48 partial tiles in 11 dependent
// L2 round trips (14.7 us at K = 8192).  Here:
//   * K is cut over up to 16 clusters of 8 CTAs (128 SMs); a CTA stages its slice in chunks of
//     64 k (all loads of a chunk in flight at once, the next chunk prefetched into registers);
//   * the 8 partial tiles of a cluster are added through distributed shared memory -- CTA r
//     owns elements [128 r, 128 r + 128) and reads them from its 7 peers, no global traffic;
//   * the per-cluster partial slices go through the workspace; for each of the 8 slices the
//     LAST cluster to arrive (self-resetting counter) adds them in cluster order -- one round
//     trip with <= 16 independent loads per thread -- and writes C.
// Summation order is fixed by indices (CTA rank, then cluster number), never by timing:
// results are deterministic.  FMA pipes, not tensor cores: a CTA's 32 x 32 x 64 slice is 64 K
// FMAs, ~0.5 us on the FP64 pipe, hidden behind the loads.  Any lda / ldb / alignment.
#include <cooperative_groups.h>

#include "common.cuh"

namespace b200 {
namespace {

namespace cg = cooperative_groups;

constexpr int TILE = 32;          // rows and columns of C
constexpr int KC = 64;            // k per chunk
constexpr int NT = 128;           // thread t: row t % 32, columns 8 (t / 32) .. + 8
constexpr int CLUSTER = 8;
constexpr int MAX_CLUSTERS = 16;
constexpr int EPT = TILE * KC / NT;  // 16 elements of each operand per thread per chunk

template <typename T>
__global__ void __cluster_dims__(CLUSTER, 1, 1) __launch_bounds__(NT)
gemm_tallk_kernel(int m, int n, int k, T alpha, const T* __restrict__ A, long long lda, const T* __restrict__ B,
                  long long ldb, T beta, T* __restrict__ C, long long ldc, int k_per_cta, T* __restrict__ partial,
                  unsigned int* __restrict__ counters) {
  // As[kk][i] (row stride 32: lanes read consecutive i), Bs[j][kk] (row stride KC + 4: the 8
  // rows a warp broadcasts from start 16-byte aligned and fall in different bank groups)
  __shared__ __align__(16) T As[KC][TILE];
  __shared__ __align__(16) T Bs[TILE][KC + 4];
  __shared__ __align__(16) T red[TILE * TILE];  // this CTA's partial tile, [j][i]
  __shared__ int s_last;
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x;
  const int rank = (int)cluster.block_rank();
  const int cl = blockIdx.x / CLUSTER, ncl = gridDim.x / CLUSTER;
  const int kbeg = min(k, (int)blockIdx.x * k_per_cta), kend = min(k, kbeg + k_per_cta);

  // element (kk, i) of A's chunk for e = 0..15: kk = (tid >> 5) + 4 e, i = tid & 31 -- a warp
  // reads 32 consecutive rows of one k column; element (kk, j) of B's chunk:
  // kk = tid & 63, j = (tid >> 6) + 2 e -- a warp reads 32 consecutive k of one column
  const int ai = tid & 31, ak = tid >> 5;
  const int bk = tid & 63, bj = tid >> 6;
  T ra[EPT], rb[EPT];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int e = 0; e < EPT; e++) {
      const int kk = k0 + ak + 4 * e;
      ra[e] = (ai < m && kk < kend) ? __ldg(A + ai + (long long)kk * lda) : T(0);
    }
#pragma unroll
    for (int e = 0; e < EPT; e++) {
      const int kk = k0 + bk, j = bj + 2 * e;
      rb[e] = (j < n && kk < kend) ? __ldg(B + kk + (long long)j * ldb) : T(0);
    }
  };
  auto stage = [&]() {
#pragma unroll
    for (int e = 0; e < EPT; e++) As[ak + 4 * e][ai] = ra[e];
#pragma unroll
    for (int e = 0; e < EPT; e++) Bs[bj + 2 * e][bk] = rb[e];
  };

  const int row = tid & 31, j0 = (tid >> 5) * 8;
  T acc[8];
#pragma unroll
  for (int c = 0; c < 8; c++) acc[c] = T(0);

  if (kbeg < kend) fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += KC) {
    stage();
    __syncthreads();
    if (k0 + KC < kend) fetch(k0 + KC);  // in flight while this chunk is multiplied
    constexpr int V = 16 / (int)sizeof(T);  // k per 16-byte broadcast load of B
#pragma unroll 4
    for (int kk = 0; kk < KC; kk += V) {
      T a[V];
#pragma unroll
      for (int v = 0; v < V; v++) a[v] = As[kk + v][row];
#pragma unroll
      for (int c = 0; c < 8; c++) {
        T b[V];
        if constexpr (sizeof(T) == 8) {
          const double2 w = *reinterpret_cast<const double2*>(&Bs[j0 + c][kk]);
          b[0] = w.x; b[1] = w.y;
        } else {
          const float4 w = *reinterpret_cast<const float4*>(&Bs[j0 + c][kk]);
          b[0] = w.x; b[1] = w.y; b[2] = w.z; b[3] = w.w;
        }
#pragma unroll
        for (int v = 0; v < V; v++) acc[c] += a[v] * b[v];
      }
    }
    __syncthreads();
  }

  // ---- cluster reduction through distributed shared memory
#pragma unroll
  for (int c = 0; c < 8; c++) red[(j0 + c) * TILE + row] = acc[c];
  cluster.sync();
  const int e = rank * NT + tid;  // the element of the tile this thread finishes
  T sum = T(0);
#pragma unroll
  for (int q = 0; q < CLUSTER; q++) sum += *cluster.map_shared_rank(&red[e], q);
  cluster.sync();  // nobody leaves while a peer may still read its shared memory

  const int ci = e & 31, cj = e >> 5;
  if (ncl > 1) {
    // ---- across clusters: slices through the workspace, last arriver per slice adds them
    __stcg(partial + (size_t)cl * (TILE * TILE) + e, sum);
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const unsigned int prev = atomicAdd(&counters[rank], 1u);
      const int last = (prev == (unsigned)ncl - 1u);
      if (last) counters[rank] = 0;  // self-reset for the next launch
      s_last = last;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    T v[MAX_CLUSTERS];
#pragma unroll
    for (int z = 0; z < MAX_CLUSTERS; z++) v[z] = z < ncl ? __ldcg(partial + (size_t)z * (TILE * TILE) + e) : T(0);
    sum = T(0);
#pragma unroll
    for (int z = 0; z < MAX_CLUSTERS; z++) sum += v[z];
  }
  if (ci < m && cj < n) {
    T* cp = C + ci + (long long)cj * ldc;
    *cp = (beta == T(0)) ? alpha * sum : alpha * sum + beta * *cp;
  }
}

}  // namespace

// Returns -1 when the shape is not this kernel's (then the split-K paths run).
template <typename T>
int gemm_tallk_dispatch(b200_ctx* ctx, int m, int n, int k, T alpha, const T* A, int lda, const T* B, int ldb, T beta,
                        T* C, int ldc) {
  if (!ctx->opt_tallk || m > TILE || n > TILE || k < 1024) return -1;
  // one cluster per 512 of K (a CTA then takes one 64-wide chunk), at most 16 clusters = 128 SMs
  const int ncl = std::max(1, std::min(MAX_CLUSTERS, (k + CLUSTER * KC - 1) / (CLUSTER * KC)));
  const int ctas = ncl * CLUSTER;
  int k_per_cta = (k + ctas - 1) / ctas;
  k_per_cta = (k_per_cta + 3) / 4 * 4;
  if (ncl > 1 && ensure_workspace(ctx, (size_t)ncl * TILE * TILE * sizeof(T))) return 1;
  if (ctx->num_counters < CLUSTER) return -1;
  gemm_tallk_kernel<T><<<ctas, NT, 0, ctx->compute>>>(m, n, k, alpha, A, (long long)lda, B, (long long)ldb, beta, C,
                                                      (long long)ldc, k_per_cta, (T*)ctx->workspace, ctx->counters);
  B200_CUDA(cudaGetLastError());
  note_launch(ctx, sizeof(T) == 8 ? "dgemm_tallk_32x32_cluster8" : "sgemm_tallk_32x32_cluster8");
  return 0;
}

template int gemm_tallk_dispatch<float>(b200_ctx*, int, int, int, float, const float*, int, const float*, int, float,
                                        float*, int);
template int gemm_tallk_dispatch<double>(b200_ctx*, int, int, int, double, const double*, int, const double*, int,
                                         double, double*, int);

}  // namespace b200
