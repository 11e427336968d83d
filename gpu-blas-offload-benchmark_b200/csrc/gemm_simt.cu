// gemm_simt.cu -- general SGEMM / DGEMM on the FP32 / FP64 FMA pipes.
//
// Replaces cublasGemmEx at reference cuBLAS/gemm.hh:134-146,158-170,177-187 for
// every shape the tensor-core paths (dgemm_dmma.cu, sgemm_tc.cu) do not take:
// dimensions / leading dimensions that break 16-byte alignment (half of a
// stride-1 sweep), the sub-tile sizes of the offload-threshold region, the
// skinny shapes (M=N=32 with K large -> split-K) and transposed operands.
//
// Classic shared-memory tiling, written for column-major operands:
//   * A tile is read along M (contiguous), B tile along K (contiguous), both
//     guarded + zero-filled, staged through registers into a double-buffered
//     smem tile (one __syncthreads per k-tile).
//   * each thread owns a TM x TN micro-tile split into 16-byte chunks that are
//     BM/QM apart so every LDS.128 of a quarter-warp hits 8 distinct 16-byte
//     bank groups (no conflicts) and C stores stay coalesced along M.
//   * split-K over blockIdx.z with a workspace + last-arriving-CTA reduction
//     in fixed order: deterministic, single launch.
#include <stdlib.h>

#include "common.cuh"

namespace b200 {
namespace {

constexpr int kMaxSplits = 48;

template <typename T, int N>
struct alignas(sizeof(T) * N) Pack {
  T v[N];
};

template <typename T, int BM_, int BN_, int BK_, int TM_, int TN_>
struct SimtCfg {
  static constexpr int BM = BM_, BN = BN_, BK = BK_, TM = TM_, TN = TN_;
  static constexpr int TXN = BM / TM;  // threads along M
  static constexpr int TYN = BN / TN;  // threads along N
  static constexpr int NT = TXN * TYN;
  static constexpr int FULLV = 16 / (int)sizeof(T);
  static constexpr int VM = TM < FULLV ? TM : FULLV;  // vector width along M
  static constexpr int VN = TN < FULLV ? TN : FULLV;
  static constexpr int QM = TM / VM;  // chunks along M
  static constexpr int QN = TN / VN;
  static constexpr int PAD = 4;
  static constexpr int LA = (BM * BK) / NT;  // A elements per thread per tile
  static constexpr int LB = (BK * BN) / NT;
  static_assert((BM * BK) % NT == 0 && (BK * BN) % NT == 0, "tile/threads");
};

template <typename T, typename Cfg>
__global__ void __launch_bounds__(Cfg::NT)
gemm_simt_kernel(int m, int n, int k, T alpha, const T* __restrict__ A,
                 long long sa_i, long long sa_p, const T* __restrict__ B,
                 long long sb_p, long long sb_j, T beta, T* __restrict__ C,
                 long long ldc, int k_per_split, T* __restrict__ partial,
                 unsigned int* __restrict__ counters) {
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK;
  constexpr int TM = Cfg::TM, TN = Cfg::TN, NT = Cfg::NT;
  constexpr int VM = Cfg::VM, VN = Cfg::VN, QM = Cfg::QM, QN = Cfg::QN;
  constexpr int PAD = Cfg::PAD, LA = Cfg::LA, LB = Cfg::LB;

  __shared__ __align__(16) T As[2][BK][BM + PAD];
  __shared__ __align__(16) T Bs[2][BK][BN + PAD];
  __shared__ bool is_last;

  const int tid = threadIdx.x;
  const int tx = tid % Cfg::TXN;
  const int ty = tid / Cfg::TXN;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(k, kbeg + k_per_split);
  const int ntiles = (kend - kbeg + BK - 1) / BK;

  T acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = T(0);

  T ra[LA], rb[LB];

  auto gload = [&](int kt) {
    const int kb = kbeg + kt * BK;
#pragma unroll
    for (int l = 0; l < LA; l++) {
      const int e = tid + l * NT;
      const int i = e % BM, p = e / BM;
      const int gi = m0 + i, gp = kb + p;
      ra[l] = (gi < m && gp < kend) ? __ldg(A + gi * sa_i + gp * sa_p) : T(0);
    }
#pragma unroll
    for (int l = 0; l < LB; l++) {
      const int e = tid + l * NT;
      const int p = e % BK, j = e / BK;
      const int gp = kb + p, gj = n0 + j;
      rb[l] = (gp < kend && gj < n) ? __ldg(B + gp * sb_p + gj * sb_j) : T(0);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int l = 0; l < LA; l++) {
      const int e = tid + l * NT;
      As[buf][e / BM][e % BM] = ra[l];
    }
#pragma unroll
    for (int l = 0; l < LB; l++) {
      const int e = tid + l * NT;
      Bs[buf][e % BK][e / BK] = rb[l];
    }
  };

  if (ntiles > 0) {
    gload(0);
    sstore(0);
  }
  __syncthreads();

  for (int kt = 0; kt < ntiles; kt++) {
    const int buf = kt & 1;
    if (kt + 1 < ntiles) gload(kt + 1);
#pragma unroll
    for (int p = 0; p < BK; p++) {
      T a[TM], b[TN];
#pragma unroll
      for (int q = 0; q < QM; q++) {
        const Pack<T, VM> t = *reinterpret_cast<const Pack<T, VM>*>(
            &As[buf][p][q * (BM / QM) + tx * VM]);
#pragma unroll
        for (int r = 0; r < VM; r++) a[q * VM + r] = t.v[r];
      }
#pragma unroll
      for (int q = 0; q < QN; q++) {
        const Pack<T, VN> t = *reinterpret_cast<const Pack<T, VN>*>(
            &Bs[buf][p][q * (BN / QN) + ty * VN]);
#pragma unroll
        for (int r = 0; r < VN; r++) b[q * VN + r] = t.v[r];
      }
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < ntiles) sstore(buf ^ 1);
    __syncthreads();
  }

  // element (i, j) of the micro-tile lives at this row / column of the tile
  auto row_of = [&](int i) { return (i / VM) * (BM / QM) + tx * VM + (i % VM); };
  auto col_of = [&](int j) { return (j / VN) * (BN / QN) + ty * VN + (j % VN); };

  if (gridDim.z == 1) {
#pragma unroll
    for (int j = 0; j < TN; j++) {
      const int gj = n0 + col_of(j);
      if (gj >= n) continue;
#pragma unroll
      for (int i = 0; i < TM; i++) {
        const int gi = m0 + row_of(i);
        if (gi >= m) continue;
        T* cp = C + gi + gj * ldc;
        *cp = (beta == T(0)) ? alpha * acc[i][j] : alpha * acc[i][j] + beta * *cp;
      }
    }
    return;
  }

  // split-K: partial tiles in the workspace, last CTA of the tile reduces
  const int tile_id = blockIdx.y * gridDim.x + blockIdx.x;
  const long long tile_elems = (long long)BM * BN;
  T* mine = partial + ((long long)blockIdx.z * gridDim.x * gridDim.y + tile_id) * tile_elems;
#pragma unroll
  for (int j = 0; j < TN; j++)
#pragma unroll
    for (int i = 0; i < TM; i++) mine[col_of(j) * BM + row_of(i)] = acc[i][j];
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int prev = atomicAdd(&counters[tile_id], 1u);
    is_last = (prev == gridDim.z - 1);
    if (is_last) counters[tile_id] = 0;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const long long zstride = (long long)gridDim.x * gridDim.y * tile_elems;
  const T* base = partial + (long long)tile_id * tile_elems;
  // z outermost so the EPT loads of one split are independent and in flight together,
  // and ZB splits per step: a loop over single splits is one dependent L2 round trip per
  // split (measured: 128 splits of the 32x32 config = 19 of the kernel's 22 us).  The
  // additions stay in ascending z order, so the result does not depend on ZB.
  constexpr int EPT = (BM * BN) / NT;
  constexpr int ZB = EPT >= 32 ? 1 : 32 / EPT;
  T s[EPT];
#pragma unroll
  for (int i = 0; i < EPT; i++) s[i] = T(0);
  const int nz = (int)gridDim.z;
  int z = 0;
  for (; ZB > 1 && z + ZB <= nz; z += ZB) {
    T v[ZB][EPT];
#pragma unroll
    for (int u = 0; u < ZB; u++)
#pragma unroll
      for (int i = 0; i < EPT; i++) v[u][i] = __ldcg(base + (z + u) * zstride + tid + i * NT);
#pragma unroll
    for (int u = 0; u < ZB; u++)
#pragma unroll
      for (int i = 0; i < EPT; i++) s[i] += v[u][i];
  }
  for (; z < nz; z++) {
    const T* src = base + z * zstride + tid;
#pragma unroll
    for (int i = 0; i < EPT; i++) s[i] += __ldcg(src + i * NT);
  }
#pragma unroll
  for (int i = 0; i < EPT; i++) {
    const int e = tid + i * NT;
    const int gi = m0 + e % BM, gj = n0 + e / BM;
    if (gi >= m || gj >= n) continue;
    T* cp = C + gi + gj * ldc;
    *cp = (beta == T(0)) ? alpha * s[i] : alpha * s[i] + beta * *cp;
  }
}

template <typename T>
__global__ void scale_mat_kernel(int m, int n, T beta, T* C, long long ldc) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)m * n) return;
  T* cp = C + (e % m) + (e / m) * ldc;
  *cp = (beta == T(0)) ? T(0) : beta * *cp;
}

template <typename T, typename Cfg>
int launch_simt(b200_ctx* ctx, int m, int n, int k, T alpha, const T* A,
                long long sa_i, long long sa_p, const T* B, long long sb_p,
                long long sb_j, T beta, T* C, int ldc, int want_ctas,
                const char* name) {
  const int gx = (m + Cfg::BM - 1) / Cfg::BM, gy = (n + Cfg::BN - 1) / Cfg::BN;
  const long long tiles = (long long)gx * gy;
  int splits = 1;
  if (tiles < want_ctas && tiles <= ctx->num_counters) {
    splits = (int)((want_ctas + tiles - 1) / tiles);
    const int max_splits = k / (Cfg::BK * 2);  // >= 2 k-tiles per split (latency regime)
    if (splits > max_splits) splits = max_splits;
    // the last CTA of a tile adds the partials of every split: beyond ~48 that serial tail
    // costs more than the extra CTAs save
    if (splits > kMaxSplits) splits = kMaxSplits;
    if (splits < 1) splits = 1;
  }
  int k_per_split = ((k + splits - 1) / splits + Cfg::BK - 1) / Cfg::BK * Cfg::BK;
  if (k_per_split < Cfg::BK) k_per_split = Cfg::BK;
  splits = (k + k_per_split - 1) / k_per_split;
  T* partial = nullptr;
  if (splits > 1) {
    const size_t need = (size_t)splits * tiles * Cfg::BM * Cfg::BN * sizeof(T);
    if (ensure_workspace(ctx, need)) return 1;
    partial = (T*)ctx->workspace;
  }
  if (gy > 65535 || splits > 65535) {
    set_error("gemm_simt: grid too large (gy=%d splits=%d)", gy, splits);
    return 2;
  }
  dim3 grid(gx, gy, splits);
  gemm_simt_kernel<T, Cfg><<<grid, Cfg::NT, 0, ctx->compute>>>(
      m, n, k, alpha, A, sa_i, sa_p, B, sb_p, sb_j, beta, C, (long long)ldc,
      k_per_split, partial, ctx->counters);
  B200_CUDA(cudaGetLastError());
  note_launch(ctx, name);
  return 0;
}

}  // namespace

template <typename T>
int gemm_simt_dispatch(b200_ctx* ctx, int ta, int tb, int m, int n, int k,
                       T alpha, const T* A, int lda, const T* B, int ldb, T beta,
                       T* C, int ldc) {
  if (m == 0 || n == 0) return 0;
  if (k == 0 || alpha == T(0)) {
    const long long total = (long long)m * n;
    scale_mat_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, ctx->compute>>>(
        m, n, beta, C, (long long)ldc);
    B200_CUDA(cudaGetLastError());
    note_launch(ctx, "gemm_scale_only");
    return 0;
  }
  const long long sa_i = ta ? lda : 1, sa_p = ta ? 1 : lda;
  const long long sb_p = tb ? ldb : 1, sb_j = tb ? 1 : ldb;
  constexpr bool F = sizeof(T) == 4;
  using Big = SimtCfg<T, 128, 128, F ? 16 : 8, 8, 8>;
  using Mid = SimtCfg<T, 64, 64, 16, 4, 4>;
  using Small = SimtCfg<T, 32, 32, 32, 2, 2>;
  const int sms = ctx->sm_count;
  const long long t128 = (long long)((m + 127) / 128) * ((n + 127) / 128);
  const long long t64 = (long long)((m + 63) / 64) * ((n + 63) / 64);
  if (m >= 96 && n >= 96 && t128 >= sms)
    return launch_simt<T, Big>(ctx, m, n, k, alpha, A, sa_i, sa_p, B, sb_p, sb_j,
                               beta, C, ldc, sms, F ? "sgemm_ffma_128x128" : "dgemm_dfma_128x128");
  static const char* simt_env = getenv("B200_SIMT_SMALL");  // tuning: t64 limit for the 32x32 config
  const long long small_limit = simt_env ? atoll(simt_env) : 16;  // <= 256^2: measured (ncu durations)
  if (t64 <= small_limit)
    return launch_simt<T, Small>(ctx, m, n, k, alpha, A, sa_i, sa_p, B, sb_p, sb_j,
                                 beta, C, ldc, 4 * sms, F ? "sgemm_ffma_32x32" : "dgemm_dfma_32x32");
  if (m > 32 && n > 32 && t64 * 4 >= sms)
    return launch_simt<T, Mid>(ctx, m, n, k, alpha, A, sa_i, sa_p, B, sb_p, sb_j,
                               beta, C, ldc, 2 * sms, F ? "sgemm_ffma_64x64" : "dgemm_dfma_64x64");
  if (m > 32 && n > 32)
    return launch_simt<T, Mid>(ctx, m, n, k, alpha, A, sa_i, sa_p, B, sb_p, sb_j,
                               beta, C, ldc, 2 * sms, F ? "sgemm_ffma_64x64_splitk" : "dgemm_dfma_64x64_splitk");
  return launch_simt<T, Small>(ctx, m, n, k, alpha, A, sa_i, sa_p, B, sb_p, sb_j,
                               beta, C, ldc, 4 * sms, F ? "sgemm_ffma_32x32" : "dgemm_dfma_32x32");
}

template int gemm_simt_dispatch<float>(b200_ctx*, int, int, int, int, int, float,
                                       const float*, int, const float*, int, float,
                                       float*, int);
template int gemm_simt_dispatch<double>(b200_ctx*, int, int, int, int, int, double,
                                        const double*, int, const double*, int,
                                        double, double*, int);

}  // namespace b200
