// dgemm_tma.cu -- large-problem DGEMM: FP64 tensor-core tiles (DMMA.8x8x4) fed by a
// warp-specialised TMA producer through a 3-stage full/empty mbarrier pipeline into
// 128-byte-swizzled shared memory.  sm_100a only.
//
// Same math and tile shape as dgemm_dmma.cu (128x128x32 CTA tile, 8 consumer warps as
// 2(M) x 4(N), 64x32 warp tiles); what changes is who moves the data and how work is scheduled:
//   * 384 threads = 3 warpgroups.  Warpgroup 2 is the producer: one lane of its first warp
//     issues cp.async.bulk.tensor.2d for the whole stage (8 boxes of A, 2 of B, 64 KB) and arms
//     the stage's `full` mbarrier with the byte count; the group hands its registers to the
//     consumers (setmaxnreg.dec 40 / .inc 232 -- a 384-thread block gets 168 at launch, the 64
//     accumulator doubles + double-buffered fragments need ~200).  The consumer warps execute
//     no copy instructions at all -- in the cp.async kernel their LDGSTS bursts competed with
//     the fragment LDS for the LSU (87% DMMA-pipe utilisation there, 99% here).
//   * persistent CTAs (one per SM) walk the tiles; the producer runs up to 3 k-tiles ahead
//     across tile boundaries and C stores drain behind the next tile's DMMAs.  The partial last
//     wave is scheduled stream-K (equal k-tile segments per SM, partial tiles combined by the
//     last contributor in contributor order); fewer tiles than SMs -> split-K, grid (tiles, splits).
//   * there is no __syncthreads in the main loop: a consumer warp waits on full[s],
//     runs its 256 DMMAs, and one lane arrives on empty[s] (count 8); warps drift
//     apart freely, so one warp's fragment loads hide behind another's DMMAs.
//   * TMA writes dense boxes, so padding is impossible; bank conflicts are removed by
//     the swizzle plus two index permutations that DMMA is indifferent to:
//       A (M-contiguous): box {16 m, 32 k} -> [k][16 m] rows of 128 B, SWIZZLE_128B
//         keyed on k mod 8.  The four k of one DMMA are taken as {s, 2+s, 4+s, 6+s}
//         (s = 0, 1 for the two DMMA steps of an 8-wide k group), which sends the 16
//         lanes of an LDS.64 half-warp to 16 distinct 8-byte bank pairs.
//       B (K-contiguous): box {16 k, 128 n} -> [n][16 k] rows of 128 B keyed on n mod 8.
//         One LDS.128 fetches k = 2j, 2j+1 (both DMMA steps); fragment lane group g
//         reads column (g&1)*4 + (g>>1) of its 8-column tile so that the 8 lanes of a
//         quarter-warp cover all 8 16-byte chunks.  The matching column permutation is
//         applied when C is stored.
//     (B's k pairing {2j, 2j+1} <-> steps s = 0, 1 is the same k set A uses.)
//   * out-of-range rows / columns / k are zero-filled by TMA, so ragged edges need no
//     special code; the only requirement is TMA's: lda, ldb even and 16-byte aligned
//     bases (otherwise dgemm_dmma.cu's cp.async path runs).
#include <stdio.h>

#include <algorithm>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace b200 {
namespace {

using namespace ptx;

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int STAGES = 3;
constexpr int CONSUMER_WARPS = 8;
// 8 consumer warps (two warpgroups) + one producer warpgroup (1 working warp, 3 idle): whole
// warpgroups so that setmaxnreg can move the producer group's registers to the consumers
constexpr int NTHREADS = 32 * (CONSUMER_WARPS + 4);
constexpr uint32_t A_STAGE_BYTES = (BM / 16) * BK * 128;  // 8 chunks x [32 k][128 B]
constexpr uint32_t B_STAGE_BYTES = (BK / 16) * BN * 128;  // 2 halves x [128 n][128 B]
constexpr uint32_t STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 64;
// "dgemm_smem_pad": launch with (nearly) all of the SM's shared memory, which leaves the L1 its
// minimum size -- the configuration in which the stage-release race described in run_tile showed
// up within a few launches (tests/test_gpu_dgemm_streamk.py keeps it covered)
constexpr uint32_t SMEM_BYTES_PADDED = 226 * 1024;
constexpr int MI = 8, NI = 4;  // DMMA tiles per warp (64 x 32 warp tile)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}
__device__ __forceinline__ double lds64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

__global__ void __launch_bounds__(NTHREADS, 1)
dgemm_tma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 int m, int n, int k, double alpha, double beta, double* __restrict__ C, long long ldc,
                 int tiles_m, int tiles_n, int k_per_split, double* __restrict__ partial,
                 unsigned int* __restrict__ counters, int sk_tiles, int sk_slots, unsigned int zero) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  __shared__ bool is_last;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  // Persistent over tiles when there is no split-K (gridDim.y == 1, gridDim.x <= tiles):
  // CTA b takes tiles b, b + gridDim.x, ...; the producer runs up to STAGES k-tiles ahead
  // ACROSS tile boundaries and the consumers' C stores drain behind the next tile's DMMAs,
  // so a short-K tile no longer pays load latency + compute + store in sequence
  // (M = N = 8192, K = 32: DMMA pipe 51 % active with one CTA per tile).
  // With split-K the grid is (tiles, splits) and every CTA handles exactly one tile.
  const int num_tiles = tiles_m * tiles_n;
  auto tile_coords = [&](int tile, int& m0, int& n0) {
    constexpr int GROUP = 8;
    const int group_size = GROUP * tiles_n;
    const int first_m = (tile / group_size) * GROUP;
    const int gm = min(tiles_m - first_m, GROUP);
    m0 = (first_m + (tile % group_size) % gm) * BM;
    n0 = ((tile % group_size) / gm) * BN;
  };
  const int kbeg = blockIdx.y * k_per_split;
  const int kend = min(k, kbeg + k_per_split);
  const int ntiles = (kend - kbeg + BK - 1) / BK;

  // Stream-K for the partial last wave (persistent mode only).  The first
  // dp_tiles = num_tiles - sk_tiles tiles are whole-tile work, round-robin; the k-tiles of
  // the remaining sk_tiles tiles form one linear space of sk_tiles * ntiles units that is
  // cut into gridDim.x equal contiguous segments, so every SM gets the same amount of the
  // leftover work instead of `sk_tiles` SMs getting a whole tile and the rest nothing
  // (2048^3: 256 tiles on 148 SMs = 2 tile-times -> 1.73).  A segment that does not cover
  // its whole tile publishes a partial; the last contributor to arrive adds the partials in
  // contributor order (fixed by the segment arithmetic, not by timing) and stores C.
  const int dp_tiles = num_tiles - sk_tiles;
  const long long sk_units = (long long)sk_tiles * ntiles;
  auto sk_begin = [&](int b) { return (long long)b * sk_units / (int)gridDim.x; };
  // the CTA whose segment contains unit u: max b with sk_begin(b) <= u
  auto sk_cta_of = [&](long long u) { return (int)(((u + 1) * (int)gridDim.x - 1) / sk_units); };

  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= CONSUMER_WARPS) {
    // ===================== TMA producer =====================
    setmaxnreg_dec<40>();
    if (warp == CONSUMER_WARPS && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto load_ktile = [&](int m0, int n0, int kb) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_STAGE_BYTES;
        mbar_expect_tx(full_bar(stage), STAGE_BYTES);
#pragma unroll
        for (int c = 0; c < BM / 16; c++) tma_load_2d(sa + c * 4096, &map_a, full_bar(stage), m0 + c * 16, kb);
#pragma unroll
        for (int h = 0; h < BK / 16; h++) tma_load_2d(sb + h * 16384, &map_b, full_bar(stage), kb + h * 16, n0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      };
      if (sk_tiles > 0) {
        for (int tile = blockIdx.x; tile < dp_tiles; tile += gridDim.x) {
          int m0, n0;
          tile_coords(tile, m0, n0);
          for (int kt = 0; kt < ntiles; kt++) load_ktile(m0, n0, kt * BK);
        }
        long long u = sk_begin(blockIdx.x);
        const long long uend = sk_begin(blockIdx.x + 1);
        while (u < uend) {
          const int t = (int)(u / ntiles), kt0 = (int)(u % ntiles);
          const int kt1 = (int)min((long long)ntiles, kt0 + (uend - u));
          int m0, n0;
          tile_coords(dp_tiles + t, m0, n0);
          for (int kt = kt0; kt < kt1; kt++) load_ktile(m0, n0, kt * BK);
          u += kt1 - kt0;
        }
        return;
      }
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m0, n0;
        tile_coords(tile, m0, n0);
        for (int kt = 0; kt < ntiles; kt++) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_STAGE_BYTES;
          const int kb = kbeg + kt * BK;
          // A split-K slice must not read past its own k range: the zero fill only covers
          // k >= K, so the last tile of an inner slice is clipped by the tensor map of the
          // full matrix -- k_per_split is a multiple of BK, which makes every slice but the
          // last end on a tile boundary.
          mbar_expect_tx(full_bar(stage), STAGE_BYTES);
#pragma unroll
          for (int c = 0; c < BM / 16; c++) tma_load_2d(sa + c * 4096, &map_a, full_bar(stage), m0 + c * 16, kb);
#pragma unroll
          for (int h = 0; h < BK / 16; h++) tma_load_2d(sb + h * 16384, &map_b, full_bar(stage), kb + h * 16, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    return;  // the split-K reduction below synchronises the consumer warps only
  }
  // 64 accumulator doubles + double-buffered fragments per thread: ~200 registers, more than
  // the 168 a 384-thread block gets at launch
  setmaxnreg_inc<232>();

  const int wm = warp & 1, wn = (warp >> 1) & 3;
  const int g = lane >> 2, t = lane & 3;
  double acc[MI][NI][2];

  // A fragment byte offsets inside a stage for (mi parity p, step s); + (mi/2)*4096 + kk*1024
  uint32_t offA[2][2];
#pragma unroll
  for (int p = 0; p < 2; p++)
#pragma unroll
    for (int s = 0; s < 2; s++) {
      const int x = p * 4 + (g >> 1), kq = 2 * t + s;
      offA[p][s] = (uint32_t)(wm * 4 * 4096 + (g & 1) * 8 + kq * 128 + ((x ^ kq) << 4));
    }
  // B fragment byte offsets for even / odd k groups; + ni*1024 + (kk/2)*16384
  const int rg = (g & 1) * 4 + (g >> 1);  // column of the 8-wide tile this lane group reads
  const uint32_t offB0 = A_STAGE_BYTES + (uint32_t)((wn * 32 + rg) * 128 + ((t ^ rg) << 4));
  const uint32_t offB1 = A_STAGE_BYTES + (uint32_t)((wn * 32 + rg) * 128 + (((4 + t) ^ rg) << 4));

  int stage = 0;
  uint32_t phase = 0;
  // ===================== DMMA consumers: all k-tiles of one output tile =====================
  auto run_tile = [&](int nkt) {
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
      for (int j = 0; j < NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int kt = 0; kt < nkt; kt++) {
      mbar_wait(full_bar(stage), phase);
      const uint32_t sbase = base + stage * STAGE_BYTES;
      // register double buffering over the 8 DMMA steps (kk = 0..3, s = 0..1) of the tile
      double a[2][MI];
      double2 b2[2][NI];
#pragma unroll
      for (int ni = 0; ni < NI; ni++) b2[0][ni] = lds128(sbase + offB0 + ni * 1024);
#pragma unroll
      for (int mi = 0; mi < MI; mi++) a[0][mi] = lds64(sbase + offA[mi & 1][0] + (mi >> 1) * 4096);
#pragma unroll
      for (int step = 0; step < 8; step++) {
        const int kk = step >> 1, s = step & 1;
        const int cur = step & 1, nxt = cur ^ 1;
        if (step + 1 < 8) {
          const int kk2 = (step + 1) >> 1, s2 = (step + 1) & 1;
#pragma unroll
          for (int mi = 0; mi < MI; mi++)
            a[nxt][mi] = lds64(sbase + offA[mi & 1][s2] + (mi >> 1) * 4096 + kk2 * 1024);
          if (s2 == 0) {
#pragma unroll
            for (int ni = 0; ni < NI; ni++)
              b2[kk2 & 1][ni] = lds128(sbase + ((kk2 & 1) ? offB1 : offB0) + ni * 1024 + (kk2 >> 1) * 16384);
          }
        }
#pragma unroll
        for (int mi = 0; mi < MI; mi++)
#pragma unroll
          for (int ni = 0; ni < NI; ni++)
            dmma(acc[mi][ni][0], acc[mi][ni][1], a[cur][mi], s ? b2[kk & 1][ni].y : b2[kk & 1][ni].x);
      }
      // Releasing the stage lets the producer's next TMA overwrite it, so every fragment load
      // must have RETURNED by then.  Shared-memory loads are asynchronous (a register is only
      // waited for by the instruction that reads it) and DMMA is register-only, so ptxas is free
      // to move the mbarrier arrive up between the last loads' issue and the DMMAs that consume
      // them -- it did: LDS, LDS, SYNCS.ARRIVE, DMMA in the SASS -- and when the LSU queue is
      // backed up (beta != 0 epilogues of the other warps streaming C through a minimum-size L1)
      // a load could still be in flight when the refill landed: one DMMA step of a warp computed
      // on the k-tile three ahead (errors of a few products in C(+)= blocks, found by the
      // Always-mode pipeline test once the kernel asked for more shared memory).  Tie the barrier
      // address to one word of every load result: `zero` is a kernel argument, always 0, that the
      // compiler cannot fold, so the arrive cannot issue before the loads have completed.
      uint32_t dep = 0;
#pragma unroll
      for (int mi = 0; mi < MI; mi++) dep ^= (uint32_t)__double2loint(a[0][mi]) ^ (uint32_t)__double2loint(a[1][mi]);
#pragma unroll
      for (int ni = 0; ni < NI; ni++) dep ^= (uint32_t)__double2loint(b2[0][ni].x) ^ (uint32_t)__double2loint(b2[1][ni].x);
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar(stage) + dep * zero);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  };
  // epilogue: DMMA column (2t + r) of tile ni is matrix column r*4 + t (the B permutation)
  auto store_tile = [&](int m0, int n0) {
    if (beta == 0.0 && m0 + BM <= m && n0 + BN <= n) {
      // interior tile, the harness' case: one address per column, 8 stores at immediate
      // offsets -- the general path below spends ~10 instructions per element on guards and
      // 64-bit address arithmetic, which shows when K is short (M = N = 8192, K = 32)
#pragma unroll
      for (int ni = 0; ni < NI; ni++) {
#pragma unroll
        for (int r = 0; r < 2; r++) {
          double* col = C + (m0 + wm * 64 + g) + (long long)(n0 + wn * 32 + ni * 8 + r * 4 + t) * ldc;
#pragma unroll
          for (int mi = 0; mi < MI; mi++) col[mi * 8] = alpha * acc[mi][ni][r];
        }
      }
      return;
    }
    if (m0 + BM <= m && n0 + BN <= n) {
      // interior tile, beta != 0 (the accumulate blocks of the Always-mode pipeline): batch
      // the 16 loads of C per column group, then update and store -- element-by-element
      // load/FMA/store is a chain of 64 dependent DRAM round trips per thread (measured:
      // +0.19 ms on a 1.93 ms 8192 x 2048 x 2048 block)
#pragma unroll
      for (int ni = 0; ni < NI; ni++) {
        double* col[2];
        double old[2][MI];
#pragma unroll
        for (int r = 0; r < 2; r++) {
          col[r] = C + (m0 + wm * 64 + g) + (long long)(n0 + wn * 32 + ni * 8 + r * 4 + t) * ldc;
#pragma unroll
          for (int mi = 0; mi < MI; mi++) old[r][mi] = col[r][mi * 8];
        }
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
          for (int mi = 0; mi < MI; mi++) col[r][mi * 8] = alpha * acc[mi][ni][r] + beta * old[r][mi];
      }
      return;
    }
#pragma unroll
    for (int ni = 0; ni < NI; ni++) {
#pragma unroll
      for (int r = 0; r < 2; r++) {
        const int gj = n0 + wn * 32 + ni * 8 + r * 4 + t;
        if (gj >= n) continue;
#pragma unroll
        for (int mi = 0; mi < MI; mi++) {
          const int gi = m0 + wm * 64 + mi * 8 + g;
          if (gi >= m) continue;
          double* cp = C + gi + gj * ldc;
          const double v = alpha * acc[mi][ni][r];
          *cp = (beta == 0.0) ? v : v + beta * *cp;
        }
      }
    }
  };

  constexpr int CT = 32 * CONSUMER_WARPS;
  constexpr long long tile_elems = (long long)BM * BN;
  if (gridDim.y == 1) {
    // ---- persistent: whole tiles ...
    for (int tile = blockIdx.x; tile < dp_tiles; tile += gridDim.x) {
      int m0, n0;
      tile_coords(tile, m0, n0);
      run_tile(ntiles);
      store_tile(m0, n0);
    }
    if (sk_tiles == 0) return;
    // ---- ... then this CTA's segment of the stream-K space
    long long u = sk_begin(blockIdx.x);
    const long long uend = sk_begin(blockIdx.x + 1);
    while (u < uend) {
      const int t = (int)(u / ntiles), kt0 = (int)(u % ntiles);
      const int kt1 = (int)min((long long)ntiles, kt0 + (uend - u));
      int m0, n0;
      tile_coords(dp_tiles + t, m0, n0);
      run_tile(kt1 - kt0);
      u += kt1 - kt0;
      if (kt0 == 0 && kt1 == ntiles) {  // the segment happens to cover the whole tile
        store_tile(m0, n0);
        continue;
      }
      const int first = sk_cta_of((long long)t * ntiles);
      const int ncontrib = sk_cta_of((long long)(t + 1) * ntiles - 1) - first + 1;
      double* slots = partial + (long long)t * sk_slots * tile_elems;
      {
        double* mine = slots + (long long)((int)blockIdx.x - first) * tile_elems;
#pragma unroll
        for (int mi = 0; mi < MI; mi++)
#pragma unroll
          for (int ni = 0; ni < NI; ni++)
#pragma unroll
            for (int r = 0; r < 2; r++) __stcg(mine + ((mi * NI + ni) * 2 + r) * CT + tid, acc[mi][ni][r]);
      }
      __threadfence();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (tid == 0) {
        const unsigned int prev = atomicAdd(&counters[t], 1u);
        is_last = (prev == (unsigned)ncontrib - 1);
        if (is_last) counters[t] = 0;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const bool last = is_last;
      asm volatile("bar.sync 1, 256;" ::: "memory");  // flag read by all before the next segment rewrites it
      if (!last) continue;
      __threadfence();
#pragma unroll
      for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < NI; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
      for (int z = 0; z < ncontrib; z++) {
        const double* src = slots + (long long)z * tile_elems + tid;
#pragma unroll
        for (int mi = 0; mi < MI; mi++)
#pragma unroll
          for (int ni = 0; ni < NI; ni++)
#pragma unroll
            for (int r = 0; r < 2; r++) acc[mi][ni][r] += __ldcg(src + ((mi * NI + ni) * 2 + r) * CT);
      }
      store_tile(m0, n0);
    }
    return;
  }

  // ---- split-K: one tile per CTA; consumers publish fragment-ordered partials, the last
  // CTA of the tile sums them
  const int bid = blockIdx.x;
  int m0, n0;
  tile_coords(bid, m0, n0);
  run_tile(ntiles);
  {
    const long long zstride = (long long)gridDim.x * tile_elems;
    {
      double* mine = partial + (long long)blockIdx.y * zstride + (long long)bid * tile_elems;
#pragma unroll
      for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < NI; ni++)
#pragma unroll
          for (int r = 0; r < 2; r++) mine[((mi * NI + ni) * 2 + r) * CT + tid] = acc[mi][ni][r];
    }
    __threadfence();
    asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 consumer warps
    if (tid == 0) {
      const unsigned int prev = atomicAdd(&counters[bid], 1u);
      is_last = (prev == gridDim.y - 1);
      if (is_last) counters[bid] = 0;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (!is_last) return;
    __threadfence();
    {
      const double* src0 = partial + (long long)bid * tile_elems + tid;
#pragma unroll
      for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < NI; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
      for (int z = 0; z < (int)gridDim.y; z++) {
        const double* src = src0 + z * zstride;
#pragma unroll
        for (int mi = 0; mi < MI; mi++)
#pragma unroll
          for (int ni = 0; ni < NI; ni++)
#pragma unroll
            for (int r = 0; r < 2; r++) acc[mi][ni][r] += __ldcg(src + ((mi * NI + ni) * 2 + r) * CT);
      }
    }
  }
  store_tile(m0, n0);
}

std::atomic<bool> g_attr_done_dev[kMaxDevices] = {};

}  // namespace

// returns -1 when the shape is not eligible (caller falls back to the cp.async kernel)
int dgemm_tma_dispatch(b200_ctx* ctx, int m, int n, int k, double alpha, const double* A, int lda,
                       const double* B, int ldb, double beta, double* C, int ldc) {
  if ((lda & 1) || (ldb & 1) || ((uintptr_t)A % 16) || ((uintptr_t)B % 16)) return -1;
  const int tiles_m = (m + BM - 1) / BM, tiles_n = (n + BN - 1) / BN;
  const long long tiles = (long long)tiles_m * tiles_n;
  if (tiles > 0x7fffffffLL) return -1;
  int splits = 1;
  if (tiles < ctx->sm_count && tiles <= ctx->num_counters) {
    // floor: tiles x splits CTAs must fit in ONE wave (ceil put 16 tiles x 10 splits = 160
    // CTAs on 148 SMs, i.e. two waves of K/10 instead of one of K/9)
    splits = (int)(ctx->sm_count / tiles);
    const int max_splits = k / (BK * 2);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  // Split-K fills tiles x splits of the SMs; when that leaves > 10 % idle (e.g. 64 tiles x 2 on 148
  // SMs) the stream-K schedule below -- every SM an equal share of ALL k-tiles -- is the better
  // one-wave split, provided its segments are long enough.
  if (splits >= 2 && ctx->opt_dgemm_streamk && tiles * splits * 10 < (long long)ctx->sm_count * 9 &&
      tiles * ((k + BK - 1) / BK) / ctx->sm_count >= 8)
    splits = 1;
  int k_per_split = ((k + splits - 1) / splits + BK - 1) / BK * BK;
  splits = (k + k_per_split - 1) / k_per_split;
  if (splits > 65535) return -1;
  double* partial = nullptr;
  if (splits > 1) {
    if (ensure_workspace(ctx, (size_t)splits * tiles * BM * BN * sizeof(double))) return 1;
    partial = (double*)ctx->workspace;
  }
  // stream-K over the partial last wave (see the kernel): worth it when that wave leaves
  // >= 10 % of the SMs idle, every SM's segment is >= 8 k-tiles (measured: with 2-3 k-tile segments the
  // partial-tile traffic costs more than the idle SMs, 51 vs 43 us at 2048 x 2048 x 128) and the whole job is short
  // enough for the saving to show (<= 8 waves)
  int sk_tiles = 0, sk_slots = 0;
  const int P = ctx->sm_count;
  const int ktiles = (k + BK - 1) / BK;
  if (splits == 1 && ctx->opt_dgemm_streamk) {
    const int rem = (int)(tiles % P);
    const long long waves = (tiles + P - 1) / P;
    const long long seg = (long long)rem * ktiles / P;  // k-tiles per SM in the stream-K space
    if (rem > 0 && rem * 10 <= P * 9 && seg >= 8 && waves <= 8 && rem <= ctx->num_counters) {
      sk_tiles = rem;
      sk_slots = (int)(ktiles / seg) + 2;  // upper bound on contributors per tile
      if (ensure_workspace(ctx, (size_t)sk_tiles * sk_slots * BM * BN * sizeof(double))) return 1;
      partial = (double*)ctx->workspace;
    }
  }
  CUtensorMap map_a, map_b;
  if (make_tensor_map_2d(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, A, m, k, (uint64_t)lda * 8, 16, BK,
                         CU_TENSOR_MAP_SWIZZLE_128B))
    return 2;
  if (make_tensor_map_2d(&map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, B, k, n, (uint64_t)ldb * 8, 16, BN,
                         CU_TENSOR_MAP_SWIZZLE_128B))
    return 2;
  std::atomic<bool>& g_attr_done = g_attr_done_dev[ctx->device % kMaxDevices];
  if (!g_attr_done) {
    B200_CUDA(cudaFuncSetAttribute(dgemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES_PADDED));
    g_attr_done = true;
  }
  // no split-K: persistent CTAs, one per SM, each walks tiles b, b + grid, ...
  dim3 grid((unsigned)(splits > 1 ? tiles : (sk_tiles ? P : std::min<long long>(tiles, P))), splits);
  dgemm_tma_kernel<<<grid, NTHREADS, ctx->opt_dgemm_smem_pad ? SMEM_BYTES_PADDED : SMEM_BYTES, ctx->compute>>>(
      map_a, map_b, m, n, k, alpha, beta, C, (long long)ldc, tiles_m, tiles_n, k_per_split, partial, ctx->counters,
      sk_tiles, sk_slots, 0u);
  B200_CUDA(cudaGetLastError());
  note_launch(ctx, splits > 1 ? "dgemm_dmma_tma_128x128x32_splitk"
                          : (sk_tiles ? "dgemm_dmma_tma_128x128x32_streamk" : "dgemm_dmma_tma_128x128x32"));
  return 0;
}

}  // namespace b200
