// dgemm_dmma.cu -- DGEMM on the FP64 tensor-core pipe (DMMA) for sm_100a.
//
// Replaces cublasGemmEx(..., CUDA_R_64F, ..., CUBLAS_COMPUTE_64F, ...) at
// reference cuBLAS/gemm.hh:141-146,165-170,183-187 for column-major
// NoTrans/NoTrans operands (the only form the harness issues).
//
// FP64 has no tcgen05/TMEM path on Blackwell; the FP64 tensor instruction is
// warp-level mma.sync m8n8k4 (SASS DMMA.8x8x4).  Structure:
//   * CTA tile 128x128x32, 8 warps as 2(M) x 4(N), warp tile 64x32 = 8x4 DMMA
//     tiles -> 64 accumulator doubles per thread, 12 fragment loads per 32 DMMA
//     (plus a 64x64x32 / 4-warp config, 2 CTAs per SM, for mid-size problems).
//   * 3-stage cp.async pipeline (16-byte copies when lda/ldb/pointers allow,
//     8-byte otherwise), zero-filled at the ragged edges via src-size, so any
//     m, n, k works without a separate edge kernel.  The copies of tile kt+2 are
//     issued in 8 slices between the DMMA groups of tile kt (never as one burst
//     after the barrier), and fragments are double-buffered in registers.
//   * A stage is [k][m] with row stride 132 doubles, B stage is [n][k] with row
//     stride 36 doubles: both strides are 8 words mod 32 banks, which makes
//     every fragment LDS.64 of a half-warp (4 k x 4 m|n) conflict-free, and the
//     global->smem copies keep the operands' native column-major order (A is
//     M-contiguous, B is K-contiguous) -- no transposition anywhere.
//   * grouped tile order (8 tile-rows per group) so co-resident CTAs share A/B
//     panels in the 126 MB L2.
//   * split-K with workspace + last-arriving-CTA reduction when the tile count
//     cannot fill 148 SMs.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include <atomic>

#include "common.cuh"

namespace b200 {
namespace {

constexpr int BK = 32;
constexpr int STAGES = 3;
constexpr int B_LD = BK + 4;  // doubles; 36*2 words = 8 mod 32 banks
// A rows are BM + 4 doubles: (BM+4)*2 words = 8 mod 32 for BM = 32, 64 and 128
template <int BM, int BN>
constexpr size_t smem_bytes() {
  return (size_t)STAGES * (BK * (BM + 4) + BN * B_LD) * sizeof(double);
}
static_assert(smem_bytes<128, 128>() <= 227 * 1024, "smem budget");

__device__ __forceinline__ void cp_async16(void* smem, const void* g, int src_bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(g), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* g, int src_bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(g), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// WARPS_M x WARPS_N warps; each owns a (BM/WARPS_M) x (BN/WARPS_N) warp tile of
// MI x NI DMMA tiles.  2x4 (64x32 warp tiles) minimises fragment loads; 4x4 (32x32)
// puts 4 warps on every scheduler, which is what saturates the DMMA pipe: ptxas
// must separate back-to-back DMMAs of one warp (NOPs in the SASS), so 2 warps per
// scheduler top out near 79% tensor-pipe utilisation (profiles/r1_dgemm_ncu.md).
// Tile configs: 128x128 (1 CTA/SM; large unaligned problems -- aligned ones use the
// TMA-fed twin in dgemm_tma.cu) and 32x32 (4 warps, 4 CTAs/SM) for everything below
// about one wave of big tiles, where latency and SM coverage matter more than reuse.
template <bool ALIGNED, int BM, int BN, int WARPS_M, int WARPS_N>
__global__ void __launch_bounds__(32 * WARPS_M * WARPS_N, (BM * BN <= 64 * 64) ? 2 : 1)
dgemm_dmma_kernel(int m, int n, int k, double alpha, const double* __restrict__ A,
                  long long lda, const double* __restrict__ B, long long ldb,
                  double beta, double* __restrict__ C, long long ldc, int tiles_m,
                  int tiles_n, int k_per_split, double* __restrict__ partial,
                  unsigned int* __restrict__ counters) {
  constexpr int A_LD = BM + 4;
  constexpr int A_STAGE = BK * A_LD, B_STAGE = BN * B_LD;  // doubles
  extern __shared__ __align__(16) double smem[];
  double* As = smem;                      // [STAGES][BK][A_LD]
  double* Bs = smem + STAGES * A_STAGE;   // [STAGES][BN][B_LD]
  __shared__ bool is_last;

  constexpr int NTHREADS = 32 * WARPS_M * WARPS_N;
  constexpr int WTM = BM / WARPS_M, WTN = BN / WARPS_N;  // warp tile
  constexpr int MI = WTM / 8, NI = WTN / 8;              // DMMA tiles per warp
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm = warp % WARPS_M, wn = warp / WARPS_M;
  const int g = lane >> 2, t = lane & 3;

  // grouped rasterisation
  constexpr int GROUP = 8;
  const int bid = blockIdx.x;
  const int group_size = GROUP * tiles_n;
  const int first_m = (bid / group_size) * GROUP;
  const int gm = min(tiles_m - first_m, GROUP);
  const int tm = first_m + (bid % group_size) % gm;
  const int tn = (bid % group_size) / gm;
  const int m0 = tm * BM, n0 = tn * BN;

  const int kbeg = blockIdx.y * k_per_split;
  const int kend = min(k, kbeg + k_per_split);
  const int ntiles = (kend - kbeg + BK - 1) / BK;

  // ---- global -> shared copy plan (loop-invariant part hoisted) --------------
  // ALIGNED: thread owns one 16-byte column chunk of A (rows 2c, 2c+1) for k-rows
  // p0, p0+4, ... and one 16-byte k-chunk of B (k = 2cb, 2cb+1) for columns
  // j0, j0+16, ...; per k-tile only the base pointers advance.
  constexpr int A_CHUNKS = (BK * BM / 2) / NTHREADS;  // per thread per tile
  constexpr int B_CHUNKS = (BN * BK / 2) / NTHREADS;
  const int a_c = tid % (BM / 2), a_p0 = tid / (BM / 2);
  const int b_c = tid % (BK / 2), b_j0 = tid / (BK / 2);
  constexpr int A_PSTEP = NTHREADS / (BM / 2);
  constexpr int B_JSTEP = NTHREADS / (BK / 2);
  const int a_bytes = max(0, min(2, m - (m0 + 2 * a_c))) * 8;  // m-edge, invariant
  unsigned b_mask = 0;                                         // n-edge, invariant
#pragma unroll
  for (int l = 0; l < B_CHUNKS; l++)
    if (n0 + b_j0 + l * B_JSTEP < n) b_mask |= 1u << l;
  const double* a_src = A + (m0 + 2 * a_c) + (long long)(kbeg + a_p0) * lda;
  const double* b_src = B + (kbeg + 2 * b_c) + (long long)(n0 + b_j0) * ldb;
  const int a_dst = a_p0 * A_LD + 2 * a_c;
  const int b_dst = b_j0 * B_LD + 2 * b_c;

  // The copies of one k-tile are issued in KSTEPS slices interleaved with the
  // DMMA groups of the previous tile: issuing all 16 LDGSTS per thread in one burst
  // right after the barrier occupies the LSU/MIO pipe for ~1000 cycles and starves
  // the fragment LDS behind it (measured: 29.0 -> 35 TFLOP/s without the burst).
  constexpr int KSTEPS = BK / 4;
  auto in_slice = [](int l, int chunks, int part) {
    return chunks >= KSTEPS ? (l / (chunks / KSTEPS) == part) : (l * (KSTEPS / chunks) == part);
  };
  auto load_slice = [&](int kt, int stage, int part) {
    const int kb = kbeg + kt * BK;
    double* as = As + stage * A_STAGE;
    double* bs = Bs + stage * B_STAGE;
    if (ALIGNED) {
      const bool full_k = kb + BK <= kend;
      const double* ap = a_src + (long long)kt * BK * lda;
      const double* bp = b_src + kt * BK;
#pragma unroll
      for (int l = 0; l < A_CHUNKS; l++) {
        if (!in_slice(l, A_CHUNKS, part)) continue;
        const int bytes = (full_k || kb + a_p0 + l * A_PSTEP < kend) ? a_bytes : 0;
        cp_async16(as + a_dst + l * A_PSTEP * A_LD, bytes ? ap + (long long)l * A_PSTEP * lda : A, bytes);
      }
#pragma unroll
      for (int l = 0; l < B_CHUNKS; l++) {
        if (!in_slice(l, B_CHUNKS, part)) continue;
        int bytes = (b_mask >> l) & 1 ? 16 : 0;
        if (!full_k && bytes) bytes = max(0, min(2, kend - (kb + 2 * b_c))) * 8;
        cp_async16(bs + b_dst + l * B_JSTEP * B_LD, bytes ? bp + (long long)l * B_JSTEP * ldb : B, bytes);
      }
    } else {
      constexpr int A_EL = (BK * BM) / NTHREADS, B_EL = (BN * BK) / NTHREADS;
#pragma unroll
      for (int l = 0; l < A_EL; l++) {
        if (!in_slice(l, A_EL, part)) continue;
        const int e = tid + l * NTHREADS;
        const int i = e % BM, p = e / BM;
        const int gi = m0 + i, gp = kb + p;
        const bool ok = gi < m && gp < kend;
        cp_async8(as + p * A_LD + i, ok ? A + gi + gp * lda : A, ok ? 8 : 0);
      }
#pragma unroll
      for (int l = 0; l < B_EL; l++) {
        if (!in_slice(l, B_EL, part)) continue;
        const int e = tid + l * NTHREADS;
        const int p = e % BK, j = e / BK;
        const int gp = kb + p, gj = n0 + j;
        const bool ok = gj < n && gp < kend;
        cp_async8(bs + j * B_LD + p, ok ? B + gp + gj * ldb : B, ok ? 8 : 0);
      }
    }
  };
  auto load_tile = [&](int kt, int stage) {
#pragma unroll
    for (int part = 0; part < KSTEPS; part++) load_slice(kt, stage, part);
  };

  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; i++)
#pragma unroll
    for (int j = 0; j < NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (s < ntiles) load_tile(s, s);
    cp_async_commit();
  }

  for (int kt = 0; kt < ntiles; kt++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    const int nk = kt + STAGES - 1;
    const bool do_load = nk < ntiles;
    const double* as = As + (kt % STAGES) * A_STAGE + wm * WTM + g;
    const double* bs = Bs + (kt % STAGES) * B_STAGE + (wn * WTN + g) * B_LD;
    // fragments are double-buffered in registers: the 12 LDS.64 of step k4+1 are
    // issued before the 32 DMMAs of step k4, so no DMMA ever waits on shared memory
    double a[2][MI], b[2][NI];
#pragma unroll
    for (int mi = 0; mi < MI; mi++) a[0][mi] = as[t * A_LD + mi * 8];
#pragma unroll
    for (int ni = 0; ni < NI; ni++) b[0][ni] = bs[ni * 8 * B_LD + t];
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; k4++) {
      const int cur = k4 & 1, nxt = cur ^ 1;
      if (do_load) load_slice(nk, nk % STAGES, k4);
      if (k4 + 1 < BK / 4) {
#pragma unroll
        for (int mi = 0; mi < MI; mi++) a[nxt][mi] = as[((k4 + 1) * 4 + t) * A_LD + mi * 8];
#pragma unroll
        for (int ni = 0; ni < NI; ni++) b[nxt][ni] = bs[ni * 8 * B_LD + (k4 + 1) * 4 + t];
      }
#pragma unroll
      for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < NI; ni++) dmma(acc[mi][ni][0], acc[mi][ni][1], a[cur][mi], b[cur][ni]);
    }
    cp_async_commit();
  }
  cp_async_wait<0>();

  if (gridDim.y > 1) {
    // split-K: every CTA publishes its fragment-ordered partial tile
    const long long tile_elems = (long long)BM * BN;
    const long long zstride = (long long)gridDim.x * tile_elems;
    double* mine = partial + (long long)blockIdx.y * zstride + (long long)bid * tile_elems;
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
      for (int ni = 0; ni < NI; ni++)
#pragma unroll
        for (int r = 0; r < 2; r++)
          mine[((mi * NI + ni) * 2 + r) * NTHREADS + tid] = acc[mi][ni][r];
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const unsigned int prev = atomicAdd(&counters[bid], 1u);
      is_last = (prev == gridDim.y - 1);
      if (is_last) counters[bid] = 0;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const double* base = partial + (long long)bid * tile_elems;
    // z outermost: the MI*NI*2 loads of one split are independent and all in flight
    // together (an inner z loop serialises ~64 dependent L2 round trips: +30 us)
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
      for (int ni = 0; ni < NI; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    // ZB splits per step keep >= 32 loads in flight per thread for the small configs (the
    // 32x32 tile has only 8 accumulators per thread: one split per step is one dependent
    // L2 round trip per split).  Ascending z order either way.
    constexpr int EPT = MI * NI * 2;
    constexpr int ZB = EPT >= 32 ? 1 : 32 / EPT;
    const int nz = (int)gridDim.y;
    int z = 0;
    for (; ZB > 1 && z + ZB <= nz; z += ZB) {
      double v[ZB][EPT];
#pragma unroll
      for (int u = 0; u < ZB; u++)
#pragma unroll
        for (int e = 0; e < EPT; e++) v[u][e] = __ldcg(base + (z + u) * zstride + tid + e * NTHREADS);
#pragma unroll
      for (int u = 0; u < ZB; u++)
#pragma unroll
        for (int mi = 0; mi < MI; mi++)
#pragma unroll
          for (int ni = 0; ni < NI; ni++)
#pragma unroll
            for (int r = 0; r < 2; r++) acc[mi][ni][r] += v[u][(mi * NI + ni) * 2 + r];
    }
    for (; z < nz; z++) {
      const double* src = base + z * zstride + tid;
#pragma unroll
      for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < NI; ni++)
#pragma unroll
          for (int r = 0; r < 2; r++) acc[mi][ni][r] += __ldcg(src + ((mi * NI + ni) * 2 + r) * NTHREADS);
    }
  }

  // epilogue: each accumulator register covers 8 consecutive rows x 4 columns
#pragma unroll
  for (int ni = 0; ni < NI; ni++) {
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const int gj = n0 + wn * WTN + ni * 8 + t * 2 + r;
      if (gj >= n) continue;
#pragma unroll
      for (int mi = 0; mi < MI; mi++) {
        const int gi = m0 + wm * WTM + mi * 8 + g;
        if (gi >= m) continue;
        double* cp = C + gi + gj * ldc;
        const double v = alpha * acc[mi][ni][r];
        *cp = (beta == 0.0) ? v : v + beta * *cp;
      }
    }
  }
}

template <bool ALIGNED, int BM, int BN, int WM, int WN>
int launch_dmma(b200_ctx* ctx, int m, int n, int k, double alpha, const double* A, int lda,
                const double* B, int ldb, double beta, double* C, int ldc, int min_ctas,
                int min_ktiles_per_split) {
  const int tiles_m = (m + BM - 1) / BM, tiles_n = (n + BN - 1) / BN;
  const long long tiles = (long long)tiles_m * tiles_n;
  if (tiles > 0x7fffffffLL) return -1;
  int splits = 1;
  if (tiles < min_ctas && tiles <= ctx->num_counters) {
    splits = (int)((min_ctas + tiles - 1) / tiles);
    const int max_splits = k / (BK * min_ktiles_per_split);
    if (splits > max_splits) splits = max_splits;
    if (splits > 48) splits = 48;  // the last CTA's reduction is serial in the split count
    if (splits < 1) splits = 1;
  }
  int k_per_split = ((k + splits - 1) / splits + BK - 1) / BK * BK;
  splits = (k + k_per_split - 1) / k_per_split;
  if (splits > 65535) return -1;
  double* partial = nullptr;
  if (splits > 1) {
    if (ensure_workspace(ctx, (size_t)splits * tiles * BM * BN * sizeof(double))) return 1;
    partial = (double*)ctx->workspace;
  }
  constexpr size_t SMEM = smem_bytes<BM, BN>();
  auto kern = dgemm_dmma_kernel<ALIGNED, BM, BN, WM, WN>;
  static std::atomic<bool> attr_set_dev[kMaxDevices] = {};
  std::atomic<bool>& attr_set = attr_set_dev[ctx->device % kMaxDevices];
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    attr_set = true;
  }
  dim3 grid((unsigned)tiles, splits);
  kern<<<grid, 32 * WM * WN, SMEM, ctx->compute>>>(
      m, n, k, alpha, A, (long long)lda, B, (long long)ldb, beta, C, (long long)ldc, tiles_m, tiles_n,
      k_per_split, partial, ctx->counters);
  B200_CUDA(cudaGetLastError());
  // string literals per (tile, copy width, split-K): b200_last_kernel never points into shared mutable storage
  static_assert(BM == BN && (BM == 32 || BM == 64 || BM == 128), "name table covers the square tile configs");
  static const char* const names[3][2][2] = {
      {{"dgemm_dmma_32x32x32_cp8", "dgemm_dmma_32x32x32_cp8_splitk"}, {"dgemm_dmma_32x32x32_cp16", "dgemm_dmma_32x32x32_cp16_splitk"}},
      {{"dgemm_dmma_64x64x32_cp8", "dgemm_dmma_64x64x32_cp8_splitk"}, {"dgemm_dmma_64x64x32_cp16", "dgemm_dmma_64x64x32_cp16_splitk"}},
      {{"dgemm_dmma_128x128x32_cp8", "dgemm_dmma_128x128x32_cp8_splitk"}, {"dgemm_dmma_128x128x32_cp16", "dgemm_dmma_128x128x32_cp16_splitk"}}};
  note_launch(ctx, names[BM == 32 ? 0 : BM == 64 ? 1 : 2][ALIGNED ? 1 : 0][splits > 1 ? 1 : 0]);
  return 0;
}

}  // namespace

int dgemm_dmma_dispatch(b200_ctx* ctx, int m, int n, int k, double alpha,
                        const double* A, int lda, const double* B, int ldb,
                        double beta, double* C, int ldc) {
  if (m < ctx->opt_dgemm_min || n < ctx->opt_dgemm_min || k < 16) return -1;  // sub-tile problems: SIMT path
  const bool aligned = (lda % 2 == 0) && (ldb % 2 == 0) &&
                       ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0);
  const int sms = ctx->sm_count;
  const long long t128 = (long long)((m + 127) / 128) * ((n + 127) / 128);
  const int force = ctx->opt_dgemm_tile;  // 64 / 128 force a config (tuning)
  // measured crossover (profiles/r1_size_sweep.txt): from ~100 tiles the 128x128 TMA
  // kernel (with split-K 2 below one wave) beats the 64x64 config
  // ... and tall-K shapes (M = N << K, e.g. 512 x 512 x 8192): few 128x128 tiles, but
  // split-K with >= 8 k-tiles per split still fills the machine, and the big tile
  // re-reads 4x less of A and B through L2 than 32x32 tiles do
  const long long tallk_splits = t128 > 0 ? std::min<long long>(sms / t128, k / (8 * 32)) : 0;
  // measured (scripts/gpu_r1k.sh): 512x512x8192 149 us vs 197 us on 32x32 tiles; at 384^2 x 6144
  // and below the 32x32 config is still ahead
  const bool tallk = t128 >= 16 && t128 * tallk_splits * 3 >= sms * 2;
  // ... and anything from ~50 tiles whose k-tiles the stream-K schedule can spread evenly over the SMs
  // (81 tiles at 1088^3: 105 us on 32x32 tiles, 91 us stream-K)
  const bool streamk_fill = ctx->opt_dgemm_streamk && t128 >= 48 && t128 * (long long)(k / 32) / sms >= 8;
  const bool big = force ? force == 128 : (t128 * 3 >= sms * 2 || tallk || streamk_fill);
#define B200_DMMA_ARGS ctx, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc
  if (big && aligned && ctx->opt_dgemm_tma) {
    const int rc = dgemm_tma_dispatch(B200_DMMA_ARGS);  // TMA + mbarrier pipeline (dgemm_tma.cu)
    if (rc >= 0) return rc;
  }
  if (big)
    return aligned ? launch_dmma<true, 128, 128, 2, 4>(B200_DMMA_ARGS, sms, 2)
                   : launch_dmma<false, 128, 128, 2, 4>(B200_DMMA_ARGS, sms, 2);
  // Everything below one wave of 128x128 tiles runs on 32x32 tiles (4 warps of 16x16,
  // up to 4 CTAs/SM, split-K down to one k-tile per CTA): a 64^3 problem spreads over
  // 8 SMs, a 128^3 one over 64, and the kernel is one load, 32 DMMAs per warp and the
  // last-CTA reduction.  Measured GPU-side durations (ncu, profiles/r1_size_sweep.txt):
  // 4.7 us at 48^3, 6.7 us at 128^3, 96 us at 1024^3 -- faster than the 64x64 config
  // everywhere up to the 128x128 crossover, so 64x64 is kept only as a forced option.
  if (force == 64)
    return aligned ? launch_dmma<true, 64, 64, 2, 2>(B200_DMMA_ARGS, 2 * sms, 2)
                   : launch_dmma<false, 64, 64, 2, 2>(B200_DMMA_ARGS, 2 * sms, 2);
  return aligned ? launch_dmma<true, 32, 32, 2, 2>(B200_DMMA_ARGS, 2 * sms, 1)
                 : launch_dmma<false, 32, 32, 2, 2>(B200_DMMA_ARGS, 2 * sms, 1);
#undef B200_DMMA_ARGS
}

}  // namespace b200
