// sgemm_tc.cu -- SGEMM on the 5th-generation tensor cores as a 3xTF32 split:
// tcgen05.mma kind::tf32, accumulators in TMEM, operands staged by TMA into
// 128B-swizzled shared memory through an mbarrier pipeline.  sm_100a only.
//
// Replaces cublasGemmEx(..., CUDA_R_32F, ..., CUBLAS_COMPUTE_32F, ...) at
// reference cuBLAS/gemm.hh:134-139,158-163,177-181 (column-major, NoTrans/NoTrans)
// whenever lda, ldb are multiples of 4 floats and the pointers 16-byte aligned
// (TMA's global-stride rule); everything else takes the FFMA path in
// gemm_simt.cu.
//
// Numerics.  TF32 keeps 11 significand bits, FP32 24.  Each operand is split into
// x = hi + lo and the product is formed as
//     A*B ~= A_lo*B_hi + A_hi*B_lo + A_hi*B_hi          (3 tensor-core MMAs)
// Two ways to get hi and lo (chosen per call by plan_sgemm):
//   fused      TMA lands the RAW FP32 tile; the tensor core reads only the top 19 bits of a word,
//              so the raw word is hi = trunc_tf32(x), and two converter warps write
//              lo = RN_tf32(x - hi) (the subtraction is exact) into a second shared-memory ring.
//              One launch, no copies of A and B in HBM.
//   pre-split  one pass (split_tf32_kernel) writes hi = RN_tf32(x), lo = x - hi into the
//              workspace; the GEMM loads both by TMA.
// The dropped lo*lo term and the rounding of lo to 11 bits are each <= 2^-20 relative.  The
// tensor core's FP32 accumulate truncates (measured bias ~7e-8 per K=8 step), so TMEM
// accumulates only 128 of K at a time and the epilogue warps sum those chunks in
// round-to-nearest FP32 registers (bar: 1e-5; measured 1.7e-6 ... 2.7e-6 on the harness data).
//
// Layout.  Column-major operands map onto UMMA without any transposition:
//   A (M x K, M contiguous)  = "MN-major" A: TMA boxes of {32 m, 32 k} land as
//       [k][32 m] rows of 128 bytes, 4 such chunks per 128-row tile.  UMMA
//       accepts exactly one layout for MN-major TF32: SWIZZLE_128B_BASE32B
//       (32-byte chunks XORed with k mod 4; TMA's SWIZZLE_128B_ATOM_32B), so
//       the descriptor is LBO = 4096 between m-chunks, SBO = 512 between groups
//       of 4 k, start += 1024 per K=8 MMA.  (Probed on hardware with
//       tests/cuda/tc_probe.cu: the plain SWIZZLE_128B descriptor yields zeros.)
//   B (K x N, K contiguous)  = "K-major" B: one TMA box {32 k, 128 n} lands as
//       [n][32 k] rows of 128 bytes (SWIZZLE_128B, SBO = 1024 between groups of
//       8 n; successive K=8 MMAs advance the start address by 32 bytes)
//   D (TMEM) has the 128 tile rows on the 128 lanes, so tcgen05.ld 32x32b gives
//       thread t row t: a warp's store of one column is 32 consecutive floats
//       of column-major C -- coalesced from registers, and a conflict-free [column][32 rows]
//       staging layout for the TMA-store epilogue.
//
// Two kernels, three tile shapes, share the layouts above:
//   sgemm_3xtf32_pair_kernel<FUSED, 256>  CTA pairs (cta_group::2), UMMA 256x256x8 -- the large-problem
//                             kernel, described at its definition below
//   sgemm_3xtf32_pair_kernel<FUSED, 128>  the same with UMMA 256x128x8: twice the tiles for mid sizes
//   sgemm_3xtf32_kernel<FUSED>            single CTA, UMMA 128x128x8, persistent over (tile, k-split)
//                             work items, 256 threads:
//     warp 0  TMA producer      warp 1  MMA issuer   (warp-uniform loops, one elected lane issues)
//     warps 2-3  converters (FUSED; warp 2 also allocates TMEM)
//     warps 4-7  epilogue (TMEM -> registers -> shared-memory staging -> TMA store, or -> split-K
//                partial + last-arriver reduction)
//   2 TMEM accumulator stages so the epilogue of chunk i overlaps the MMAs of chunk i+1.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <vector>

#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace b200 {
namespace {

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int ACC_STAGES = 2;
constexpr int UMMA_K = 8;
constexpr int KB_PER_CHUNK = 4;  // k-blocks accumulated in TMEM before promotion (K = 128)
constexpr int NUM_THREADS = 256;
constexpr int SINGLE_EPI_WARPS = 4;  // single-CTA kernel: warps 4-7
constexpr uint32_t TMEM_COLS = ACC_STAGES * BN;  // 256 (power of two >= 32)
constexpr uint32_t A_TILE = BM * BK * 4;         // 16 KB: 4 chunks x [32 k][128 B]
constexpr uint32_t B_TILE = BN * BK * 4;         // 16 KB: [128 n][128 B]
constexpr uint32_t SLOT = A_TILE + B_TILE;       // one ring slot: an A tile and a B tile, 32 KB
constexpr uint32_t STAGING_PER_WARP = 4096;      // epilogue: two half-buffers of [16 columns][32 rows] floats

using namespace ptx;

// Shared-memory rings of one CTA (both kernels: the CTA-pair kernel stages 128 rows of A and 128
// columns of B per CTA too).  The tensor core reads four tiles per k-block: {A_hi, B_hi} from a
// slot of the hi ring, {A_lo, B_lo} from a slot of the lo ring.
//   pre-split (FUSED = false)  TMA fills both slots from the hi/lo copies in the workspace; the
//                              rings advance together (3 + 3 slots = 192 KB).
//   FUSED                      TMA lands the caller's RAW FP32 tiles in the hi ring (4 slots: TMA
//                              latency + conversion + MMA are in flight at once) and two converter
//                              warps write the lo ring (2 slots) from them.
template <bool FUSED, int PN = 256>
struct Rings {
  // PN = columns of the (pair) tile: a CTA stages PN / 2 columns of B.  The narrow pair tile
  // (PN = 128) has 24 KB slots and spends them on a deeper raw ring: a slot is busy from the TMA
  // issue to the end of the MMAs that read it (~2 us of load latency + conversion + MMAs), and
  // with half the MMA time per k-block four slots in flight were not enough (measured: 0.96 us per
  // k-block with 4 slots).
  static constexpr uint32_t SLOT_B = A_TILE + (uint32_t)(PN / 2) * BK * 4;  // 32 KB / 24 KB
  static constexpr int R = FUSED ? (PN == 128 ? 6 : 4) : (PN == 128 ? 4 : 3);
  static constexpr int L = FUSED ? 2 : R;
  // + 4 KB per epilogue warp: the staging buffers of the TMA-store epilogue (store_rows_tma)
  static constexpr uint32_t smem_bytes(int epi_warps) {
    return (uint32_t)(R + L) * SLOT_B + 1024 /*align*/ + 256 /*barriers*/ + (uint32_t)epi_warps * STAGING_PER_WARP;
  }
  static_assert(8 * (2 * R + 2 * L + 2 * ACC_STAGES + 1) <= 256, "barrier area");
  static_assert(SLOT_B % 1024 == 0, "SWIZZLE_128B atoms need 1024-byte aligned tiles");
  uint32_t base, bars;
  __device__ explicit Rings(const uint8_t* smem_raw) {
    base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-byte alignment
    bars = base + (uint32_t)(R + L) * SLOT_B;
  }
  __device__ uint32_t staging(int epi_warp) const { return bars + 256u + (uint32_t)epi_warp * STAGING_PER_WARP; }
  __device__ uint32_t a_hi(int r) const { return base + (uint32_t)r * SLOT_B; }
  __device__ uint32_t b_hi(int r) const { return a_hi(r) + A_TILE; }
  __device__ uint32_t a_lo(int l) const { return base + (uint32_t)(R + l) * SLOT_B; }
  __device__ uint32_t b_lo(int l) const { return a_lo(l) + A_TILE; }
  // hi_full: TMA bytes have landed (pre-split: of all four tiles); hi_empty / lo_empty: the MMAs that
  // read the slot are done; lo_full (FUSED): the converter warps have written the slot
  __device__ uint32_t hi_full(int r) const { return bars + 8u * r; }
  __device__ uint32_t hi_empty(int r) const { return bars + 8u * (R + r); }
  __device__ uint32_t lo_full(int l) const { return bars + 8u * (2 * R + l); }
  __device__ uint32_t lo_empty(int l) const { return bars + 8u * (2 * R + L + l); }
  __device__ uint32_t tfull(int a) const { return bars + 8u * (2 * R + 2 * L + a); }
  __device__ uint32_t tempty(int a) const { return bars + 8u * (2 * R + 2 * L + ACC_STAGES + a); }
  __device__ uint32_t tmem_slot() const { return bars + 8u * (2 * R + 2 * L + 2 * ACC_STAGES); }
};
// position in a ring of N slots + the mbarrier phase that goes with it
struct Pos {
  int i = 0;
  uint32_t ph = 0;
  template <int N>
  __device__ __forceinline__ void next() {
    if (++i == N) { i = 0; ph ^= 1; }
  }
};

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4  [16,30) LBO>>4  [32,46) SBO>>4  [46,48) version=1  [61,64) layout type
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
constexpr uint32_t kSwizzle128B = 2;         // B: K-major, 16-byte swizzle atoms
constexpr uint32_t kSwizzle128B_Base32B = 1; // A: the only layout UMMA accepts for MN-major TF32
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32,
// A MN-major (bit 15), B K-major, N>>3 at [17,23), M>>4 at [24,29).
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (0u << 16) |
                            ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// ------------------------------------------------------------------ kernels

// x = hi + lo with hi = RN_tf32(x) (13 low bits zero), lo = x - hi (exact).  One launch
// splits both operands: blocks [0, blocks_a) take A, the rest take B.  (Pre-split path only.)
__device__ __forceinline__ void split_tf32_range(const float* __restrict__ src, float* __restrict__ hi,
                                                 float* __restrict__ lo, size_t count, int block, int blocks) {
  const size_t n4 = count / 4;
  const size_t stride = (size_t)blocks * blockDim.x;
  for (size_t i = (size_t)block * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(src) + i);
    float4 h, l;
    h.x = __uint_as_float((__float_as_uint(v.x) + 0x1000u) & 0xFFFFE000u); l.x = v.x - h.x;
    h.y = __uint_as_float((__float_as_uint(v.y) + 0x1000u) & 0xFFFFE000u); l.y = v.y - h.y;
    h.z = __uint_as_float((__float_as_uint(v.z) + 0x1000u) & 0xFFFFE000u); l.z = v.z - h.z;
    h.w = __uint_as_float((__float_as_uint(v.w) + 0x1000u) & 0xFFFFE000u); l.w = v.w - h.w;
    reinterpret_cast<float4*>(hi)[i] = h;
    reinterpret_cast<float4*>(lo)[i] = l;
  }
  if (block == 0 && threadIdx.x < (count & 3)) {
    const size_t i = n4 * 4 + threadIdx.x;
    const float x = src[i];
    const float h = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
    hi[i] = h;
    lo[i] = x - h;
  }
}

__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ a, float* __restrict__ a_hi, float* __restrict__ a_lo, size_t count_a,
                  const float* __restrict__ b, float* __restrict__ b_hi, float* __restrict__ b_lo, size_t count_b,
                  int blocks_a) {
  if ((int)blockIdx.x < blocks_a)
    split_tf32_range(a, a_hi, a_lo, count_a, blockIdx.x, blocks_a);
  else
    split_tf32_range(b, b_hi, b_lo, count_b, blockIdx.x - blocks_a, gridDim.x - blocks_a);
}

// ---- fused operand split (FUSED kernels) ----------------------------------------------------
// TMA lands the RAW FP32 tile in the hi ring: tcgen05 kind::tf32 reads only the top 19 bits of
// each word, so the raw word IS hi = trunc_tf32(x).  Two converter warps then write
// lo = RN_tf32(x - trunc_tf32(x)) word for word at the same offset of a lo-ring slot -- the
// subtraction is exact in FP32 (lo < 2^-10 |x|, <= 13 significant bits), the rounding to TF32's
// 11 bits is what the tensor core would otherwise do by truncation (a one-sided error), and the
// same offset means whatever swizzle TMA applied carries over untouched.  No pre-pass over A
// and B, no hi/lo copies in HBM, no workspace.
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128f(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float tf32_lo(float x) {
  const float lo = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  return __uint_as_float((__float_as_uint(lo) + 0x1000u) & 0xFFFFE000u);
}
// The SLOT bytes of raw words at `src` -> their lo parts at `dst`; thread `t` of 64.  Batches of
// four 16-byte words, software-pipelined by hand: the loads of batch b + 1 are in flight while
// batch b is converted and stored (the asm statements are volatile, so the order written here is
// the order issued -- a plain unrolled loop serialised load -> convert -> store on one register set).
template <uint32_t BYTES = SLOT>
__device__ __forceinline__ void convert_slot(uint32_t src, uint32_t dst, int t) {
  constexpr uint32_t STEP = 64u * 16u;  // a warp touches 512 consecutive bytes: conflict-free
  constexpr int U = 4, BATCHES = (int)(BYTES / STEP) / U;
  static_assert(BYTES % (STEP * U) == 0, "whole batches");
  const uint32_t s0 = src + (uint32_t)t * 16u, d0 = dst + (uint32_t)t * 16u;
  float4 cur[U], nxt[U];
#pragma unroll
  for (int u = 0; u < U; u++) cur[u] = lds128f(s0 + u * STEP);
#pragma unroll
  for (int b = 0; b < BATCHES; b++) {
    if (b + 1 < BATCHES) {
#pragma unroll
      for (int u = 0; u < U; u++) nxt[u] = lds128f(s0 + ((b + 1) * U + u) * STEP);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      cur[u].x = tf32_lo(cur[u].x); cur[u].y = tf32_lo(cur[u].y);
      cur[u].z = tf32_lo(cur[u].z); cur[u].w = tf32_lo(cur[u].w);
    }
#pragma unroll
    for (int u = 0; u < U; u++) sts128f(d0 + (b * U + u) * STEP, cur[u]);
#pragma unroll
    for (int u = 0; u < U; u++) cur[u] = nxt[u];
  }
}

// B200_SGEMM_TRACE: per-CTA timestamps (globaltimer, ns) at the phase boundaries of a launch,
// summarised by the host after the launch (diagnostic; `trace` is nullptr otherwise).
//   0 entry  1 set-up done  2 first operands ready (MMA warp)  3 last MMA issued
//   4 last accumulator chunk ready (epilogue warp 4)  5 that warp's stores issued  6 exit
__device__ __forceinline__ void trace_mark(unsigned long long* trace, int slot) {
  if (trace) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    trace[(size_t)blockIdx.x * 8 + slot] = t;
  }
}

// Store one thread's row of `NCOL` column results: C[row, col0 + j] = alpha * sum[j] (+ beta * C).
// The per-column work is kept to a multiply, an address step and the store: guards and the
// beta test are hoisted (ncu on M = N = 8192, K = 32: the epilogue warps were issue-latency
// bound on ~19 instructions per column).
template <int NCOL>
__device__ __forceinline__ void store_row(float* __restrict__ cp, long long ldc, const float (&sum)[NCOL],
                                          int ncols, float alpha, float beta) {
  if (ncols <= 0) return;
  if (beta == 0.0f) {
#pragma unroll
    for (int j = 0; j < NCOL; j++) {
      if (j < ncols) *cp = alpha * sum[j];  // predicated store, no branch
      cp += ldc;
    }
  } else {
#pragma unroll
    for (int j = 0; j < NCOL; j++) {
      if (j < ncols) *cp = alpha * sum[j] + beta * *cp;
      cp += ldc;
    }
  }
}

// The same row block through shared memory and the TMA store engine (beta == 0, C and ldc
// 16-byte aligned): the warp writes alpha * sum as [16 columns][32 rows] into one of its two
// 2 KB half-buffers -- lanes = consecutive rows = consecutive words, conflict-free -- and one
// lane hands the box to cp.async.bulk.tensor (SASS UTMASTG), which clips at C's bounds.  Per
// 16 columns the warp issues 16 shared-memory stores and ONE bulk store instead of 16 global
// stores per lane, and is free for the next tile's tcgen05.ld while the engine drains the box
// (M = N = 8192, K = 32: the epilogue is all there is).
__device__ __forceinline__ void sts32f(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
template <int NCOL>
__device__ __forceinline__ void store_rows_tma(const CUtensorMap* map_c, uint32_t buf, int lane, int row0, int col0,
                                               int m, int n, const float (&sum)[NCOL], float alpha) {
  if (row0 >= m) return;
#pragma unroll
  for (int s = 0; s < NCOL / 16; s++) {
    if (col0 + s * 16 >= n) break;
    const uint32_t half = buf + (uint32_t)(s & 1) * 2048u;
    if (lane == 0) bulk_wait_read<1>();  // the box that used this half two slices ago has been read
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 16; j++) sts32f(half + j * 128 + lane * 4, alpha * sum[s * 16 + j]);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(map_c, half, row0, col0 + s * 16);
      bulk_commit();
    }
  }
}

// FUSED: map_a_hi / map_b_hi describe the caller's A and B themselves; the lo maps are unused.
template <bool FUSED>
__global__ void __launch_bounds__(NUM_THREADS, 1)
sgemm_3xtf32_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                    const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                    const __grid_constant__ CUtensorMap map_c, int tma_c,
                    int m, int n, int k, float alpha, float beta, float* __restrict__ C, long long ldc,
                    int tiles_m, int tiles_n, int splits, int kb_per_split, float* __restrict__ partial,
                    unsigned int* __restrict__ counters, unsigned long long* __restrict__ trace) {
  using RG = Rings<FUSED>;
  constexpr int R = RG::R, L = RG::L;
  extern __shared__ uint8_t smem_raw[];
  __shared__ int s_is_last;
  const RG rg(smem_raw);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (rg.tmem_slot() - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work item = (tile, k split): item / splits is the tile, item % splits the k range
  const int num_items = tiles_m * tiles_n * splits;
  const int total_kb = (k + BK - 1) / BK;
  auto kb_range = [&](int item, int& kb_lo, int& kb_hi) {
    kb_lo = (item % splits) * kb_per_split;
    kb_hi = min(total_kb, kb_lo + kb_per_split);
  };

  if (threadIdx.x == 0) {
    trace_mark(trace, 0);
    for (int s = 0; s < R; s++) {
      mbar_init(rg.hi_full(s), 1);
      mbar_init(rg.hi_empty(s), 1);
    }
    for (int s = 0; s < L; s++) {
      mbar_init(rg.lo_full(s), 2);  // one arrive per converter warp
      mbar_init(rg.lo_empty(s), 1);
    }
    for (int a = 0; a < ACC_STAGES; a++) {
      mbar_init(rg.tfull(a), 1);
      mbar_init(rg.tempty(a), 4);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(rg.tmem_slot(), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (threadIdx.x == 0) trace_mark(trace, 1);

  auto tile_coords = [&](int tile, int& m0, int& n0) {
    constexpr int GROUP = 8;
    const int group_size = GROUP * tiles_n;
    const int first_m = (tile / group_size) * GROUP;
    const int gm = min(tiles_m - first_m, GROUP);
    m0 = (first_m + (tile % group_size) % gm) * BM;
    n0 = ((tile % group_size) / gm) * BN;
  };

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
    const bool elected = elect_one();
    Pos r;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int m0, n0, kb_lo, kb_hi;
      tile_coords(item / splits, m0, n0);
      kb_range(item, kb_lo, kb_hi);
      for (int kb = kb_lo; kb < kb_hi; kb++) {
        mbar_wait(rg.hi_empty(r.i), r.ph ^ 1);
        const uint32_t bar = rg.hi_full(r.i);
        if (elected) {
          mbar_expect_tx(bar, FUSED ? SLOT : 2 * SLOT);
#pragma unroll
          for (int c = 0; c < BM / 32; c++) tma_load_2d(rg.a_hi(r.i) + c * 4096, &map_a_hi, bar, m0 + c * 32, kb * BK);
          tma_load_2d(rg.b_hi(r.i), &map_b_hi, bar, kb * BK, n0);
          if (!FUSED) {  // R == L: the rings advance together
#pragma unroll
            for (int c = 0; c < BM / 32; c++) tma_load_2d(rg.a_lo(r.i) + c * 4096, &map_a_lo, bar, m0 + c * 32, kb * BK);
            tma_load_2d(rg.b_lo(r.i), &map_b_lo, bar, kb * BK, n0);
          }
        }
        __syncwarp();
        r.next<R>();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The tensor core truncates when it adds into the FP32 accumulator (measured:
    // ~7e-8 relative bias per K=8 step, 3.8e-5 at K=4096 on all-positive data), so
    // TMEM only ever accumulates KB_PER_CHUNK k-blocks; the epilogue warps promote
    // each chunk into round-to-nearest FP32 registers while the next chunk runs in
    // the other accumulator stage.
    // The loop is warp-uniform (all lanes wait on the barriers and carry the ring positions, so
    // the descriptors live in uniform registers); one elected lane issues the MMAs and commits.
    {
      const bool elected = elect_one();
      bool first_kb = true;
      Pos r, l, acc;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int kb_lo, kb_hi;
        kb_range(item, kb_lo, kb_hi);
        for (int kb0 = kb_lo; kb0 < kb_hi; kb0 += KB_PER_CHUNK) {
          mbar_wait(rg.tempty(acc.i), acc.ph ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc.i * BN;
          const int kb1 = min(kb_hi, kb0 + KB_PER_CHUNK);
          for (int kb = kb0; kb < kb1; kb++) {
            mbar_wait(rg.hi_full(r.i), r.ph);  // the TMA writes (FUSED: the converters saw them too)
            if (FUSED) mbar_wait(rg.lo_full(l.i), l.ph);
            tc_fence_after();
            if (first_kb) {
              if (elected) trace_mark(trace, 2);
              first_kb = false;
            }
            const int li = FUSED ? l.i : r.i;
            const uint32_t sa_hi = rg.a_hi(r.i), sb_hi = rg.b_hi(r.i), sa_lo = rg.a_lo(li), sb_lo = rg.b_lo(li);
            if (elected) {
#pragma unroll
              for (int ks = 0; ks < BK / UMMA_K; ks++) {
                const uint64_t a_hi = umma_desc(sa_hi + ks * 1024, 4096, 512, kSwizzle128B_Base32B);
                const uint64_t a_lo = umma_desc(sa_lo + ks * 1024, 4096, 512, kSwizzle128B_Base32B);
                const uint64_t b_hi = umma_desc(sb_hi + ks * 32, 16, 1024, kSwizzle128B);
                const uint64_t b_lo = umma_desc(sb_lo + ks * 32, 16, 1024, kSwizzle128B);
                umma_tf32(d_tmem, a_lo, b_hi, kIdesc, (kb != kb0) || (ks != 0));
                umma_tf32(d_tmem, a_hi, b_lo, kIdesc, 1);
                umma_tf32(d_tmem, a_hi, b_hi, kIdesc, 1);
              }
              umma_commit(rg.hi_empty(r.i));  // frees the slot(s) when these MMAs finish
              if (FUSED) umma_commit(rg.lo_empty(l.i));
            }
            __syncwarp();
            if (FUSED) l.next<L>();
            r.next<R>();
          }
          if (elected) umma_commit(rg.tfull(acc.i));  // chunk complete -> epilogue promotes it
          __syncwarp();
          acc.next<ACC_STAGES>();
        }
      }
      if (elected) trace_mark(trace, 3);
    }
  } else if (warp < 4) {
    // ===================== converters (FUSED): lo ring <- hi ring =====================
    if (FUSED) {
      const int ct = threadIdx.x - 64;  // 0..63
      Pos r, l;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int kb_lo, kb_hi;
        kb_range(item, kb_lo, kb_hi);
        for (int kb = kb_lo; kb < kb_hi; kb++) {
          mbar_wait(rg.hi_full(r.i), r.ph);
          mbar_wait(rg.lo_empty(l.i), l.ph ^ 1);
          convert_slot(rg.a_hi(r.i), rg.a_lo(l.i), ct);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(rg.lo_full(l.i));
          r.next<R>();
          l.next<L>();
        }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int et = threadIdx.x - 128;  // 0..127 among the epilogue threads
    Pos acc;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int tile = item / splits;
      int m0, n0, kb_lo, kb_hi;
      tile_coords(tile, m0, n0);
      kb_range(item, kb_lo, kb_hi);
      float sum[BN];
#pragma unroll
      for (int j = 0; j < BN; j++) sum[j] = 0.0f;
      for (int kb0 = kb_lo; kb0 < kb_hi; kb0 += KB_PER_CHUNK) {
        mbar_wait(rg.tfull(acc.i), acc.ph);
        tc_fence_after();
        if (threadIdx.x == 128) trace_mark(trace, 4);
        const uint32_t taddr = tmem_base + acc.i * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
          float v[32];
          tmem_ld32(taddr + c0, v);
#pragma unroll
          for (int j = 0; j < 32; j++) sum[c0 + j] += v[j];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(rg.tempty(acc.i));
        acc.next<ACC_STAGES>();
      }
      if (splits == 1) {
        const int row = m0 + q * 32 + lane;
        if (tma_c)
          store_rows_tma<BN>(&map_c, rg.staging(warp - 4), lane, m0 + q * 32, n0, m, n, sum, alpha);
        else if (row < m)
          store_row<BN>(C + row + (long long)n0 * ldc, ldc, sum, n - n0, alpha, beta);
        continue;
      }
      // ---- split-K: publish this k range's partial tile ([column][row], rows contiguous);
      // the epilogue of whichever CTA arrives last for the tile adds the partials of all
      // splits in ascending split order (deterministic) and writes C.
      {
        float* mine = partial + ((size_t)(item % splits) * (size_t)(tiles_m * tiles_n) + tile) * (BM * BN) +
                      q * 32 + lane;
#pragma unroll
        for (int j = 0; j < BN; j++) __stcg(mine + j * BM, sum[j]);
      }
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) {
        const unsigned int prev = atomicAdd(&counters[tile], 1u);
        const int last = (prev == (unsigned)splits - 1);
        if (last) counters[tile] = 0;  // self-reset for the next launch
        s_is_last = last;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const bool last = s_is_last != 0;
      asm volatile("bar.sync 1, 128;" ::: "memory");  // everyone has read the flag before it is rewritten
      if (!last) continue;
      __threadfence();
      // thread = 4 consecutive rows (one float4) x 32 columns: 32 independent 16-byte loads
      // in flight per split
      const int rg4 = et & 31, cg = et >> 5;
      const float* src = partial + (size_t)tile * (BM * BN) + (size_t)(cg * 32) * BM + rg4 * 4;
      const size_t zstride = (size_t)(tiles_m * tiles_n) * (BM * BN);
      float4 tot[32];
#pragma unroll
      for (int j = 0; j < 32; j++) tot[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int z = 0; z < splits; z++) {
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const float4 v = __ldcg(reinterpret_cast<const float4*>(src + z * zstride + j * BM));
          tot[j].x += v.x; tot[j].y += v.y; tot[j].z += v.z; tot[j].w += v.w;
        }
      }
      const int row0 = m0 + rg4 * 4;
#pragma unroll
      for (int j = 0; j < 32; j++) {
        const int col = n0 + cg * 32 + j;
        if (col >= n) continue;
        float* cp = C + row0 + (long long)col * ldc;
        const float rr[4] = {tot[j].x, tot[j].y, tot[j].z, tot[j].w};
#pragma unroll
        for (int e = 0; e < 4; e++)
          if (row0 + e < m) cp[e] = (beta == 0.0f) ? alpha * rr[e] : alpha * rr[e] + beta * cp[e];
      }
    }
    if (threadIdx.x == 128) trace_mark(trace, 5);
    if (tma_c && lane == 0) bulk_wait_read<0>();  // the staging buffers outlive their last bulk store
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
  if (threadIdx.x == 0) trace_mark(trace, 6);
}

// ------------------------------------------------- CTA-pair kernel (256 x 256)
//
// One UMMA of M = 128, N = 128, K = 8 (TF32) lasts 64 tensor-pipe clocks and reads
// 4 KB of A + 4 KB of B from shared memory: 128 B/clk, the whole shared-memory
// bandwidth of an SM, before TMA has written a byte -- the single-CTA kernel above
// stalls there (ncu: tensor pipe 74 % active).  Pairing the two SMs of a TPC
// (tcgen05 cta_group::2) makes one UMMA M = 256, N = 256: each CTA still feeds its
// own 128 rows of A, but only HALF of B's columns (the hardware shares B between the
// pair), so a 128-clock UMMA reads 4 + 4 KB per SM = 64 B/clk and TMA adds 43 B/clk
// (FUSED: TMA 21 B/clk + the converters' 21 B/clk of reads and 21 B/clk of writes).
//
//   cluster (2,1,1), 384 threads per CTA:
//     warp 0      TMA producer: own rows of A, own column half of B.  Pre-split: hi and lo
//                 tiles, bytes of BOTH CTAs counted on the LEADER's hi_full barrier.  FUSED: raw
//                 tiles, bytes counted on the CTA's own hi_full barrier (its converters wait there)
//     warp 1      MMA issuer (leader CTA only); tcgen05.commit multicasts the
//                 "slot free" / "chunk ready" arrivals to both CTAs
//     warps 2-3   (warp 2: TMEM allocator, all 512 columns = 2 accumulator stages x 256)
//                 FUSED: converters; the four converter warps of the pair arrive on the LEADER's
//                 lo_full barrier (the leader's MMAs read both CTAs' slots)
//     warps 4-11  epilogue: lane quarter = warp % 4, column half = (warp - 4) / 4,
//                 128 running sums per thread (registers moved over with setmaxnreg)
namespace pair {

constexpr int CTA_M = 128;            // rows of A and C per CTA
constexpr int PAIR_M = 2 * CTA_M;     // UMMA M
constexpr int PAIR_N = 256;           // UMMA N of the large-problem configuration (template default)
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 128 + 32 * EPI_WARPS;
static_assert(CTA_M * BK * 4 == A_TILE && (PAIR_N / 2) * BK * 4 == B_TILE, "the pair kernel shares Rings<> with the single-CTA kernel");
// PN = UMMA N = columns of a pair tile.  256: the large-problem tile.  128: the mid-size tile --
// twice as many tiles for the same problem (1536^3: 72 instead of 36 on 74 pairs), each CTA stages
// 64 columns of B (8 KB of the 16 KB B part of a ring slot), 6 KB of operand reads per UMMA
// instead of the single-CTA kernel's 8 KB for half the work.

template <bool FUSED, int PN = PAIR_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
sgemm_3xtf32_pair_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                         const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                         const __grid_constant__ CUtensorMap map_c, int tma_c,
                         int m, int n, int k, float alpha, float beta, float* __restrict__ C, long long ldc,
                         int tiles_m, int tiles_n, int raster_group, int sk_tiles, int sk_slots,
                         float* __restrict__ partial, unsigned int* __restrict__ counters,
                         unsigned long long* __restrict__ trace) {
  using RG = Rings<FUSED, PN>;
  constexpr int R = RG::R, L = RG::L;
  constexpr int PAIR_N = PN;            // shadows the namespace constant: this instantiation's UMMA N
  constexpr int CTA_N = PN / 2;         // columns of B each CTA stages = accumulator columns per epilogue warp
  constexpr uint32_t TMEM_COLS2 = ACC_STAGES * PN;  // 512 / 256
  constexpr uint32_t SLOT_USED = RG::SLOT_B;  // bytes of a ring slot: 16 KB of A + CTA_N columns of B
  constexpr uint32_t kIdesc2 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (0u << 16) |
                               ((uint32_t)(PN >> 3) << 17) | ((uint32_t)(PAIR_M >> 4) << 24);
  extern __shared__ uint8_t smem_raw[];
  __shared__ int s_is_last;
  const RG rg(smem_raw);
  // pre-split: the leader's hi_full is used (both CTAs' bytes); hi_empty / lo_empty / tfull: one per
  // CTA (multicast commits); tempty and (FUSED) lo_full: the leader's are used
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (rg.tmem_slot() - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (k + BK - 1) / BK;
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  // ---- work items.  Whole tiles first (tile = pair_id, pair_id + num_pairs, ... below dp_tiles);
  // then, if the tile count leaves a partial last wave, this pair's segment of the stream-K space:
  // the accumulation chunks (K = 128) of the last sk_tiles tiles, cut into num_pairs equal
  // contiguous runs (same scheme as the DGEMM kernel, csrc/dgemm_tma.cu).  A run that does not
  // cover its whole tile publishes a partial tile; per half of the tile (the 128 rows of one CTA
  // rank) the last CTA of that rank to arrive adds the partials in contributor order and stores C.
  const int chunks = (num_kb + KB_PER_CHUNK - 1) / KB_PER_CHUNK;
  const int dp_tiles = num_tiles - sk_tiles;
  const long long sk_units = (long long)sk_tiles * chunks;
  auto sk_begin = [&](int p) { return (long long)p * sk_units / num_pairs; };
  auto sk_pair_of = [&](long long u) { return (int)(((u + 1) * num_pairs - 1) / sk_units); };
  struct Item { int tile, c0, c1; };
  long long sk_u = 0, sk_end = 0;
  int dp_next = pair_id;
  bool sk_started = false;
  auto next_item = [&](Item& w) -> bool {
    if (dp_next < dp_tiles) {
      w = {dp_next, 0, chunks};
      dp_next += num_pairs;
      return true;
    }
    if (sk_tiles == 0) return false;
    if (!sk_started) {
      sk_u = sk_begin(pair_id);
      sk_end = sk_begin(pair_id + 1);
      sk_started = true;
    }
    if (sk_u >= sk_end) return false;
    const int t = (int)(sk_u / chunks), c0 = (int)(sk_u % chunks);
    const int c1 = (int)min((long long)chunks, c0 + (sk_end - sk_u));
    w = {dp_tiles + t, c0, c1};
    sk_u += c1 - c0;
    return true;
  };

  if (threadIdx.x == 0) {
    trace_mark(trace, 0);
    for (int s = 0; s < R; s++) {
      mbar_init(rg.hi_full(s), 1);
      mbar_init(rg.hi_empty(s), 1);
    }
    for (int s = 0; s < L; s++) {
      mbar_init(rg.lo_full(s), 4);  // two converter warps in each CTA of the pair
      mbar_init(rg.lo_empty(s), 1);
    }
    for (int a = 0; a < ACC_STAGES; a++) {
      mbar_init(rg.tfull(a), 1);
      mbar_init(rg.tempty(a), 2 * EPI_WARPS);  // every epilogue warp of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc_2sm(rg.tmem_slot(), TMEM_COLS2);
  tc_fence_before();
  __syncwarp();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (threadIdx.x == 0) trace_mark(trace, 1);

  auto tile_coords = [&](int tile, int& m0, int& n0) {
    const int GROUP = raster_group;
    const int group_size = GROUP * tiles_n;
    const int first_m = (tile / group_size) * GROUP;
    const int gm = min(tiles_m - first_m, GROUP);
    m0 = (first_m + (tile % group_size) % gm) * PAIR_M;
    n0 = ((tile % group_size) / gm) * PAIR_N;
  };

  if (warp < 4) {
    setmaxnreg_dec<48>();
    if (warp == 0) {
      // ===================== TMA producer (both CTAs; warp-uniform loop, one elected lane issues) =====================
      const bool elected = elect_one();
      Pos r, l;
      Item w;
      while (next_item(w)) {
        int m0, n0;
        tile_coords(w.tile, m0, n0);
        const int my_m = m0 + (int)rank * CTA_M, my_n = n0 + (int)rank * CTA_N;
        const int kb_hi = min(num_kb, w.c1 * KB_PER_CHUNK);
        for (int kb = w.c0 * KB_PER_CHUNK; kb < kb_hi; kb++) {
          mbar_wait(rg.hi_empty(r.i), r.ph ^ 1);
          if (FUSED) {
            const uint32_t bar = rg.hi_full(r.i);
            if (elected) {
              mbar_expect_tx(bar, SLOT_USED);
#pragma unroll
              for (int c = 0; c < CTA_M / 32; c++) tma_load_2d(rg.a_hi(r.i) + c * 4096, &map_a_hi, bar, my_m + c * 32, kb * BK);
              tma_load_2d(rg.b_hi(r.i), &map_b_hi, bar, kb * BK, my_n);
            }
            l.next<L>();
          } else {
            const uint32_t lbar = mapa_shared(rg.hi_full(r.i), 0);
            if (elected) {
              if (rank == 0) mbar_expect_tx(rg.hi_full(r.i), 4 * SLOT_USED);
#pragma unroll
              for (int c = 0; c < CTA_M / 32; c++) {
                tma_load_2d_2sm(rg.a_hi(r.i) + c * 4096, &map_a_hi, lbar, my_m + c * 32, kb * BK);
                tma_load_2d_2sm(rg.a_lo(r.i) + c * 4096, &map_a_lo, lbar, my_m + c * 32, kb * BK);
              }
              tma_load_2d_2sm(rg.b_hi(r.i), &map_b_hi, lbar, kb * BK, my_n);
              tma_load_2d_2sm(rg.b_lo(r.i), &map_b_lo, lbar, kb * BK, my_n);
            }
          }
          __syncwarp();
          r.next<R>();
        }
      }
      // tail: every slot this CTA filled has been released, i.e. no UMMA still reads
      // this CTA's shared memory and no multicast arrival is still in flight to it
      for (int s = 0; s < R; s++) {
        mbar_wait(rg.hi_empty(r.i), r.ph ^ 1);
        r.next<R>();
      }
      if (FUSED)
        for (int s = 0; s < L; s++) {
          mbar_wait(rg.lo_empty(l.i), l.ph ^ 1);
          l.next<L>();
        }
    } else if (warp == 1 && rank == 0) {
      // ===================== MMA issuer (leader CTA; warp-uniform loop, one elected lane issues) =====================
      const bool elected = elect_one();
      bool first_kb = true;
      Pos r, l, acc;
      Item w;
      while (next_item(w)) {
        for (int ch = w.c0; ch < w.c1; ch++) {
          const int kb0 = ch * KB_PER_CHUNK;
          mbar_wait(rg.tempty(acc.i), acc.ph ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc.i * PAIR_N;
          const int kb1 = min(num_kb, kb0 + KB_PER_CHUNK);
          for (int kb = kb0; kb < kb1; kb++) {
            mbar_wait(rg.hi_full(r.i), r.ph);  // FUSED: this CTA's raw tiles; the peer's are behind lo_full
            if (FUSED) mbar_wait(rg.lo_full(l.i), l.ph);
            tc_fence_after();
            if (first_kb) {
              if (elected) trace_mark(trace, 2);
              first_kb = false;
            }
            const int li = FUSED ? l.i : r.i;
            const uint32_t sa_hi = rg.a_hi(r.i), sb_hi = rg.b_hi(r.i), sa_lo = rg.a_lo(li), sb_lo = rg.b_lo(li);
            if (elected) {
#pragma unroll
              for (int ks = 0; ks < BK / UMMA_K; ks++) {
                const uint64_t a_hi = umma_desc(sa_hi + ks * 1024, 4096, 512, kSwizzle128B_Base32B);
                const uint64_t a_lo = umma_desc(sa_lo + ks * 1024, 4096, 512, kSwizzle128B_Base32B);
                const uint64_t b_hi = umma_desc(sb_hi + ks * 32, 16, 1024, kSwizzle128B);
                const uint64_t b_lo = umma_desc(sb_lo + ks * 32, 16, 1024, kSwizzle128B);
                umma_tf32_2sm(d_tmem, a_lo, b_hi, kIdesc2, (kb != kb0) || (ks != 0));
                umma_tf32_2sm(d_tmem, a_hi, b_lo, kIdesc2, 1);
                umma_tf32_2sm(d_tmem, a_hi, b_hi, kIdesc2, 1);
              }
              umma_commit_2sm(rg.hi_empty(r.i), 3);  // frees the slot(s) in both CTAs
              if (FUSED) umma_commit_2sm(rg.lo_empty(l.i), 3);
            }
            __syncwarp();
            if (FUSED) l.next<L>();
            r.next<R>();
          }
          if (elected) umma_commit_2sm(rg.tfull(acc.i), 3);  // chunk complete -> both CTAs' epilogues
          __syncwarp();
          acc.next<ACC_STAGES>();
        }
      }
      if (elected) trace_mark(trace, 3);
    } else if (FUSED && warp >= 2) {
      // ===================== converters (both CTAs): lo ring <- hi ring =====================
      const int ct = threadIdx.x - 64;  // 0..63
      const uint32_t leader_lo_full0 = mapa_shared(rg.lo_full(0), 0);
      Pos r, l;
      Item w;
      while (next_item(w)) {
        const int kb_hi = min(num_kb, w.c1 * KB_PER_CHUNK);
        for (int kb = w.c0 * KB_PER_CHUNK; kb < kb_hi; kb++) {
          mbar_wait(rg.hi_full(r.i), r.ph);
          mbar_wait(rg.lo_empty(l.i), l.ph ^ 1);
          convert_slot<SLOT_USED>(rg.a_hi(r.i), rg.a_lo(l.i), ct);
          fence_proxy_async_smem();
          __syncwarp();
          // what must be ordered before the leader's MMAs are this warp's shared-memory stores, and
          // fence.proxy.async has done that; the arrive itself is only the signal (ptx.cuh:
          // a cluster-scope release would add a MEMBAR.GPU to every k-block)
          if (lane == 0) mbar_arrive_cluster(leader_lo_full0 + 8u * l.i);
          r.next<R>();
          l.next<L>();
        }
      }
    }
  } else {
    // ===================== epilogue (both CTAs) =====================
    // 128 x 48 + 256 x 224 = 63488 of the SM's 65536 registers.  (232 would be an exact fit on
    // paper; on hardware the TRY_ALLOC never succeeded and the kernel hung -- keep slack.)
    setmaxnreg_inc<224>();
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2; // which 128 of the 256 accumulator columns
    const int et = threadIdx.x - 128; // 0..255 among the epilogue threads of this CTA
    const uint32_t leader_tempty0 = mapa_shared(rg.tempty(0), 0);
    Pos acc;
    Item w;
    while (next_item(w)) {
      int m0, n0;
      tile_coords(w.tile, m0, n0);
      float sum[CTA_N];
#pragma unroll
      for (int j = 0; j < CTA_N; j++) sum[j] = 0.0f;
      for (int ch = w.c0; ch < w.c1; ch++) {
        mbar_wait(rg.tfull(acc.i), acc.ph);
        tc_fence_after();
        if (threadIdx.x == 128) trace_mark(trace, 4);
        const uint32_t taddr = tmem_base + acc.i * PAIR_N + half * CTA_N + ((uint32_t)(q * 32) << 16);
#pragma unroll
        for (int c0 = 0; c0 < CTA_N; c0 += 32) {
          float v[32];
          tmem_ld32(taddr + c0, v);
#pragma unroll
          for (int j = 0; j < 32; j++) sum[c0 + j] += v[j];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(leader_tempty0 + 8u * acc.i);
        acc.next<ACC_STAGES>();
      }
      const int trow = (int)rank * CTA_M + q * 32 + lane;  // row inside the 256 x 256 tile
      if (w.c0 == 0 && w.c1 == chunks) {
        const int row = m0 + trow;
        const int col0 = n0 + half * CTA_N;
        if (tma_c)
          store_rows_tma<CTA_N>(&map_c, rg.staging(warp - 4), lane, m0 + (int)rank * CTA_M + q * 32, col0, m, n, sum, alpha);
        else if (row < m)
          store_row<CTA_N>(C + row + (long long)col0 * ldc, ldc, sum, n - col0, alpha, beta);
        continue;
      }
      // ---- stream-K partial: slot layout [column][row] of the whole 256 x 256 tile
      const int t = w.tile - dp_tiles;
      const int first = sk_pair_of((long long)t * chunks);
      const int ncontrib = sk_pair_of((long long)(t + 1) * chunks - 1) - first + 1;
      float* slots = partial + (size_t)t * sk_slots * (PAIR_M * PAIR_N);
      {
        float* mine = slots + (size_t)(pair_id - first) * (PAIR_M * PAIR_N) + (size_t)(half * CTA_N) * PAIR_M + trow;
#pragma unroll
        for (int j = 0; j < CTA_N; j++) __stcg(mine + j * PAIR_M, sum[j]);
      }
      __threadfence();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // Each CTA of the pair owns its 128 rows of the tile to the end: the rows of rank r are complete
      // when the rank-r CTA of every contributing pair has published, and the last of THOSE adds
      // them -- so the two halves of a tile are finished by two CTAs in parallel (one CTA reducing
      // the whole 256 x 256 tile was the long pole of the fix-up: B200_SGEMM_TRACE, 4096^3, ~20 us).
      if (et == 0) {
        const unsigned int prev = atomicAdd(&counters[2 * t + (int)rank], 1u);
        const int last = (prev == (unsigned)ncontrib - 1u);
        if (last) counters[2 * t + (int)rank] = 0;  // self-reset for the next launch
        s_is_last = last;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const bool last = s_is_last != 0;
      asm volatile("bar.sync 1, 256;" ::: "memory");  // flag read by all before it can be rewritten
      if (!last) continue;
      __threadfence();
      // thread = 4 consecutive rows x 32 columns: 32 independent 16-byte loads in flight per contributor
      const int rg4 = et & 31, cg = et >> 5;
      const int hrow = (int)rank * CTA_M + rg4 * 4;  // row inside the tile
      const int row0 = m0 + hrow;
      const int cbase = cg * 32;
      if (cbase >= PAIR_N) continue;  // PN = 128: half of the column groups have nothing to add
      float4 tot[32];
#pragma unroll
      for (int j = 0; j < 32; j++) tot[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int z = 0; z < ncontrib; z++) {
        const float* src = slots + (size_t)z * (PAIR_M * PAIR_N) + (size_t)cbase * PAIR_M + hrow;
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const float4 v = __ldcg(reinterpret_cast<const float4*>(src + j * PAIR_M));
          tot[j].x += v.x; tot[j].y += v.y; tot[j].z += v.z; tot[j].w += v.w;
        }
      }
#pragma unroll
      for (int j = 0; j < 32; j++) {
        const int col = n0 + cbase + j;
        if (col >= n) continue;
        float* cp = C + row0 + (long long)col * ldc;
        const float rr[4] = {tot[j].x, tot[j].y, tot[j].z, tot[j].w};
#pragma unroll
        for (int e = 0; e < 4; e++)
          if (row0 + e < m) cp[e] = (beta == 0.0f) ? alpha * rr[e] : alpha * rr[e] + beta * cp[e];
      }
    }
    if (threadIdx.x == 128) trace_mark(trace, 5);
    if (tma_c && lane == 0) bulk_wait_read<0>();  // the staging buffers outlive their last bulk store
  }

  tc_fence_before();
  __syncwarp();
  cluster_sync();  // both CTAs are done with TMEM, shared memory and each other's barriers
  if (warp == 2) tmem_dealloc_2sm(tmem_base, TMEM_COLS2);
  if (threadIdx.x == 0) trace_mark(trace, 6);
}

}  // namespace pair

// ------------------------------------------------------------------- host

std::atomic<bool> g_attr_done_dev[kMaxDevices] = {};
std::atomic<bool> g_pair_attr_done_dev[kMaxDevices] = {};
std::atomic<int> g_pair_max_clusters_dev[kMaxDevices] = {};

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Tensor maps depend only on the operand addresses and the shape: the harness' ten back-to-back
// calls encode them once; the multi-GPU engine walks a handful of K-chunks of A per step.
struct MapSet {
  const void *pa = nullptr, *pb = nullptr;  // what a_hi / b_hi describe
  int device = -1, m = -1, n = -1, k = -1, lda = -1, ldb = -1, b_box = -1;
  bool fused = false;
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  const void* pc = nullptr;  // what `c` describes (the TMA-store epilogue), nullptr: not encoded
  int ldc = -1;
  CUtensorMap c;
};
constexpr int kMapSets = 8;

// Everything sgemm_tc_dispatch decides before it touches memory -- tile configuration, k-splits,
// stream-K, fused or pre-split operands, scratch -- in one place, so that b200_sgemm_prepare
// reserves exactly the scratch the call will use (a workspace that grows inside the caller's timed
// region costs milliseconds).
struct SgemmPlan {
  long long t1m, t1n, tiles1, t2m, t2n, tiles2, t3n, tiles3;
  int total_kb, pairs, splits, kb_per_split, sk_tiles, sk_slots, pair_n;
  bool use_pair, fused;
  size_t part_bytes, sk_bytes, need, countA, countB, offAlo, offBhi, offBlo;
};

int plan_sgemm(b200_ctx* ctx, int m, int n, int k, int lda, int ldb, SgemmPlan* out) {
  // ---- schedule: CTA pairs on 256 x 256 tiles, or single CTAs on 128 x 128 tiles cut into
  // `splits` k ranges.  Both are persistent, so time = waves x time per work item; the cost
  // model below is in tensor-pipe clocks per k-block of 32 (pair UMMA 256x256x8: 128 clk x 12 =
  // 1536, pipe ~99 % busy; single-CTA 128x128x8: 64 clk x 12 = 768 at ~74 % => ~1040) plus a
  // per-item term for the exposed part of the epilogue / split-K fix-up.  It reproduces the
  // measured crossovers (profiles/r1_size_sweep.txt: pair from 1792^3, single CTA at 1536^3,
  // split-K 2 at 1024^3).
  const long long t1m = (m + BM - 1) / BM, t1n = (n + BN - 1) / BN, tiles1 = t1m * t1n;
  const long long t2m = (m + pair::PAIR_M - 1) / pair::PAIR_M, t2n = (n + pair::PAIR_N - 1) / pair::PAIR_N,
                  tiles2 = t2m * t2n;
  if (tiles1 > 0x7fffffffLL) return -1;
  const int total_kb = (k + BK - 1) / BK;
  const int P = ctx->sm_count;
  auto waves = [](long long items, long long slots) { return (items + slots - 1) / slots; };
  int splits = 1;
  double cost1 = 0.0;
  {
    double best = -1.0;
    const int max_s = (tiles1 <= ctx->num_counters) ? std::min(32, std::max(1, total_kb / (2 * KB_PER_CHUNK))) : 1;
    for (int sp = 1; sp <= max_s; sp++) {
      int kbs = (total_kb + sp - 1) / sp;
      kbs = (kbs + KB_PER_CHUNK - 1) / KB_PER_CHUNK * KB_PER_CHUNK;
      const int real = (total_kb + kbs - 1) / kbs;  // splits after rounding to whole chunks
      if (real != sp && sp != 1) continue;
      // measured: split-K over more than one wave of work items loses (2304^3, 3 splits: 188 vs 154 us)
      if (real > 1 && tiles1 * real > P) continue;
      const double c = (double)waves(tiles1 * real, P) * (kbs * 1040.0 + (real > 1 ? 2500.0 : 1500.0));
      if (best < 0 || c < best * 0.97) {  // prefer fewer splits unless the gain is real
        best = c;
        splits = real;
      }
    }
    cost1 = best;
  }
  int kb_per_split = (total_kb + splits - 1) / splits;
  kb_per_split = (kb_per_split + KB_PER_CHUNK - 1) / KB_PER_CHUNK * KB_PER_CHUNK;
  splits = (total_kb + kb_per_split - 1) / kb_per_split;
  const size_t part_bytes = splits > 1 ? (size_t)splits * tiles1 * BM * BN * sizeof(float) : 0;

  std::atomic<bool>& g_pair_attr_done = g_pair_attr_done_dev[ctx->device % kMaxDevices];
  std::atomic<int>& g_pair_max_clusters = g_pair_max_clusters_dev[ctx->device % kMaxDevices];
  if (!g_pair_attr_done) {
    int pairs_min = ctx->sm_count / 2;
    auto setup = [&](auto kernel, uint32_t smem) -> int {
      B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * ctx->sm_count, 1, 1);
      cfg.blockDim = dim3(pair::THREADS, 1, 1);
      cfg.dynamicSmemBytes = smem;
      int clusters = 0;
      if (cudaOccupancyMaxActiveClusters(&clusters, kernel, &cfg) != cudaSuccess || clusters <= 0) {
        cudaGetLastError();
        clusters = ctx->sm_count / 2;
      }
      pairs_min = std::min(pairs_min, clusters);
      return 0;
    };
    if (setup(pair::sgemm_3xtf32_pair_kernel<false, 256>, Rings<false>::smem_bytes(pair::EPI_WARPS))) return 1;
    if (setup(pair::sgemm_3xtf32_pair_kernel<true, 256>, Rings<true>::smem_bytes(pair::EPI_WARPS))) return 1;
    if (setup(pair::sgemm_3xtf32_pair_kernel<false, 128>, Rings<false, 128>::smem_bytes(pair::EPI_WARPS))) return 1;
    if (setup(pair::sgemm_3xtf32_pair_kernel<true, 128>, Rings<true, 128>::smem_bytes(pair::EPI_WARPS))) return 1;
    g_pair_max_clusters = pairs_min;
    g_pair_attr_done = true;
  }
  const int pairs = g_pair_max_clusters;
  // ---- pair kernel: whole tiles, or stream-K over the partial last wave (see the kernel), and pair
  // vs single CTA.  Times in microseconds from one linear model fitted to this round's size sweeps
  // (profiles/r2_sgemm_schedule.md): a launch costs ~2.7 us before the first MMA result; a pair
  // spends 1.17 us per k-block of 32 (256 x 256 tile; the practical tensor rate under load), a
  // single CTA 0.88 us (128 x 128, shared-memory bound); a split-K fix-up ~11.5 us.  Stream-K: every
  // contributor writes a whole 256 x 256 partial per tile it touches, so the fix-up grows with the
  // number of partials (0.13 us each: ~256 KB out and back at ~4 TB/s when all CTAs do it at
  // once) and with the contributors ONE CTA has to add for its half tile (1.2 us each) --
  // 2304^3 (7 leftover tiles x 12 contributors) 159 -> 118 us, 2816^3 192 -> 175, 3584^3 356 -> 336,
  // 4096^3 523 -> 489, while 2048^3 (64 tiles, 6 us to win) and under-filled grids with short
  // segments (1024^3: 34 us against 29 on split-K single CTAs) stay on whole tiles / single CTAs.
  int sk_tiles = 0, sk_slots = 0;
  size_t sk_bytes = 0;
  const int chunks = (total_kb + KB_PER_CHUNK - 1) / KB_PER_CHUNK;
  constexpr double kLaunchUs = 2.7, kPairKbUs = 1.17, kSingleKbUs = 0.88, kSplitFixUs = 11.5;
  const double t1_us = kLaunchUs + (double)waves(tiles1 * splits, P) * kb_per_split * kSingleKbUs + (splits > 1 ? kSplitFixUs : 0.0);
  double t2_us = kLaunchUs + (double)waves(tiles2, pairs) * total_kb * kPairKbUs;
  if (ctx->opt_sgemm_streamk && tiles2 > 0) {
    const int rem = (int)(tiles2 % pairs);
    const long long units = (long long)rem * chunks;
    const long long seg = units / pairs;
    static const char* sk_env = getenv("B200_SGEMM_SK_MIN_US");  // tuning: smallest predicted gain worth taking
    const double min_gain_us = sk_env ? atof(sk_env) : 3.0;
    if (rem > 0 && seg >= 1 && 2 * rem <= ctx->num_counters) {
      const double contributors = (double)chunks / (double)seg + 1.0;
      const double fix_us = 5.0 + std::max(0.13 * rem * contributors, 1.2 * contributors);
      const double sk_last_kb = (double)((units + pairs - 1) / pairs) * KB_PER_CHUNK;
      const double t2_sk = kLaunchUs + ((double)(tiles2 / pairs) * total_kb + sk_last_kb) * kPairKbUs + fix_us;
      if (t2_us - t2_sk >= min_gain_us && t2_sk <= 0.97 * t2_us) {
        sk_tiles = rem;
        sk_slots = (int)(chunks / seg) + 2;
        sk_bytes = (size_t)rem * (size_t)sk_slots * pair::PAIR_M * pair::PAIR_N * sizeof(float);
        t2_us = t2_sk;
      }
    }
  }
  // ties go to the single-CTA kernel (the model is ~2 us optimistic for all-stream-K grids)
  bool use_pair = m >= pair::PAIR_M && n >= pair::PAIR_N && t2_us <= 0.95 * t1_us;
  (void)cost1;
  // ... and the mid-size pair tile 256 x 128 (whole tiles only): per k-block it is bound by shared
  // memory (144 KB per CTA: 72 KB of operand reads, 24 KB landed by TMA, 48 KB through the
  // converters) at ~0.63 us for HALF the work of the 256 x 256 tile, but it has twice the tiles --
  // what problems between ~1000^3 and ~1700^3 need to fill 74 pairs.
  // It runs on PRE-SPLIT operands: with the converters in the loop the same tile measured 0.82-0.96 us
  // per k-block whatever the ring depth or the polling of the other warps (1536^3: 47.5-48.9 us fused,
  // 41.4 pre-split including the split pass; 1280^3 39.9-41.6 vs 35.2), so the split pass
  // (~2.5 us + 0.3 us per MB of A and B) is part of its price.
  constexpr double kPair128KbUs = 0.64;
  const size_t countA = (size_t)lda * (k - 1) + m, countB = (size_t)ldb * (n - 1) + k;
  const double split_pass_us = 2.5 + 0.3e-6 * 4.0 * (double)(countA + countB);
  const long long t3n = (n + 127) / 128, tiles3 = t2m * t3n;
  const double t3_us = kLaunchUs + (double)waves(tiles3, pairs) * total_kb * kPair128KbUs +
                       (ctx->opt_sgemm_fused == 1 ? 0.3 * total_kb * kPair128KbUs : split_pass_us);
  int pair_n = pair::PAIR_N;
  if (m >= pair::PAIR_M && n >= 128 && t3_us <= 0.95 * std::min(t1_us, use_pair ? t2_us : t1_us)) {
    use_pair = true;
    pair_n = 128;
  }
  if (ctx->opt_sgemm_pair == 1) { use_pair = true; pair_n = pair::PAIR_N; }
  if (ctx->opt_sgemm_pair == 2) use_pair = false;
  if (ctx->opt_sgemm_pair == 3) { use_pair = true; pair_n = 128; }
  if (use_pair && pair_n == 128) { sk_tiles = 0; sk_slots = 0; sk_bytes = 0; }  // whole tiles only

  // ---- operand split: fused into the GEMM (converter warps: one launch, no hi/lo copies in HBM)
  // or a separate pass writing hi/lo copies into the workspace.  Fused wins wherever the launch
  // and the extra pass over A and B matter or the CTA-pair kernel runs (measured, fused vs
  // pre-split: 2048^3 74.9 vs 84.4 us, 4096^3 485 vs 523, 8192x512x512 27.5 vs 35.9,
  // 16384x4096x16384 7.93 vs 8.12 ms, 8192^3 a tie).  The one exception is the single-CTA kernel
  // on a long K range: it is already bound by shared-memory bandwidth (the UMMAs alone read
  // 128 B/clk) and the converters' traffic comes on top (1280^3 41.7 vs 38.3 us, 1536^3 49.6 vs 45.4).
  bool fused = ctx->opt_sgemm_fused != 2;
  if (ctx->opt_sgemm_fused == 0 && !use_pair && splits == 1 && total_kb >= 32) fused = false;
  if (ctx->opt_sgemm_fused == 0 && use_pair && pair_n == 128) fused = false;

  const size_t offAlo = align_up(countA * 4, 1024), offBhi = 2 * offAlo;
  const size_t offBlo = offBhi + align_up(countB * 4, 1024);
  const size_t need = fused ? 0 : offBlo + align_up(countB * 4, 1024);
  *out = SgemmPlan{t1m, t1n, tiles1, t2m, t2n, tiles2, t3n, tiles3, total_kb, pairs, splits, kb_per_split, sk_tiles, sk_slots,
                   pair_n, use_pair, fused, part_bytes, sk_bytes, need, countA, countB, offAlo, offBhi, offBlo};
  return 0;
}

}  // namespace

int sgemm_tc_dispatch(b200_ctx* ctx, int m, int n, int k, float alpha, const float* A, int lda,
                      const float* B, int ldb, float beta, float* C, int ldc) {
  if (ctx->opt_sgemm == 2) return -1;  // "ffma" forces the SIMT path, "tc" forces this one
  const bool forced = ctx->opt_sgemm == 1;
  if ((lda % 4) || (ldb % 4) || ((uintptr_t)A % 16) || ((uintptr_t)B % 16)) return -1;
  // below ~512^3 the launch and the descriptor encodes cost more than the FFMA path's single
  // launch (measured crossover, size_sweep.py)
  if (!forced && ((double)m * n * k < 1.0e8 || m < 64 || n < 64)) return -1;
  SgemmPlan plan;
  if (const int rc = plan_sgemm(ctx, m, n, k, lda, ldb, &plan)) return rc;
  const long long t1m = plan.t1m, t1n = plan.t1n, tiles1 = plan.tiles1, t2m = plan.t2m, t2n = plan.t2n, tiles2 = plan.tiles2,
                  t3n = plan.t3n, tiles3 = plan.tiles3;
  const int total_kb = plan.total_kb, pairs = plan.pairs, splits = plan.splits, kb_per_split = plan.kb_per_split,
            sk_tiles = plan.sk_tiles, sk_slots = plan.sk_slots, pair_n = plan.pair_n;
  const bool use_pair = plan.use_pair, fused = plan.fused;
  const size_t part_bytes = plan.part_bytes, sk_bytes = plan.sk_bytes, need = plan.need, countA = plan.countA,
               countB = plan.countB, offAlo = plan.offAlo, offBhi = plan.offBhi, offBlo = plan.offBlo;
  (void)total_kb; (void)t1n; (void)t2n;
  // room for the pair kernel's stream-K slots: rem tiles x (pairs / rem + 2) contributors x 256 KB
  // = (pairs + 2 rem) x 256 KB <= 56 MB whatever the problem size; the context reserves that much
  // when it is created (and b200_sgemm_prepare the pre-split copies), so that this call does not
  // reallocate inside a timed region
  if (ensure_workspace(ctx, need + std::max(part_bytes, sk_bytes))) return 1;
  char* ws = (char*)ctx->workspace;
  float *a_hi = (float*)ws, *a_lo = (float*)(ws + offAlo), *b_hi = (float*)(ws + offBhi),
        *b_lo = (float*)(ws + offBlo);
  float* sk_partial = sk_tiles ? (float*)(ws + need) : nullptr;
  const float* pa = fused ? A : a_hi;  // what the hi maps describe
  const float* pb = fused ? B : b_hi;

  const int b_box = (use_pair && pair_n == 128) ? 64 : BN;  // columns of B per TMA box = per CTA
  static thread_local MapSet sets[kMapSets];
  static thread_local int next_set = 0;
  MapSet* mc = nullptr;
  for (int i = 0; i < kMapSets; i++) {
    MapSet& s = sets[i];
    if (s.pa == pa && s.pb == pb && s.device == ctx->device && s.fused == fused && s.m == m && s.n == n && s.k == k &&
        s.lda == lda && s.ldb == ldb && s.b_box == b_box) {
      mc = &s;
      break;
    }
  }
  if (!mc) {
    mc = &sets[next_set];
    next_set = (next_set + 1) % kMapSets;
    mc->pa = nullptr;
    mc->pc = nullptr;
    if (make_tensor_map_2d(&mc->a_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, pa, m, k, (uint64_t)lda * 4, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return 2;
    if (make_tensor_map_2d(&mc->b_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, pb, k, n, (uint64_t)ldb * 4, BK, b_box, CU_TENSOR_MAP_SWIZZLE_128B)) return 2;
    if (fused) {
      mc->a_lo = mc->a_hi;
      mc->b_lo = mc->b_hi;
    } else {
      if (make_tensor_map_2d(&mc->a_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, a_lo, m, k, (uint64_t)lda * 4, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return 2;
      if (make_tensor_map_2d(&mc->b_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, b_lo, k, n, (uint64_t)ldb * 4, BK, b_box, CU_TENSOR_MAP_SWIZZLE_128B)) return 2;
    }
    mc->pa = pa; mc->pb = pb; mc->device = ctx->device; mc->fused = fused;
    mc->m = m; mc->n = n; mc->k = k; mc->lda = lda; mc->ldb = ldb; mc->b_box = b_box;
  }
  const CUtensorMap &ma_hi = mc->a_hi, &ma_lo = mc->a_lo, &mb_hi = mc->b_hi, &mb_lo = mc->b_lo;
  // TMA-store epilogue: whole tiles of a beta == 0 call whose C obeys TMA's alignment rules
  // (the harness' case); everything else stores from registers
  const int tma_c = ctx->opt_sgemm_tma_store && beta == 0.0f && (ldc % 4 == 0) && ((uintptr_t)C % 16 == 0);
  if (tma_c && (mc->pc != C || mc->ldc != ldc)) {
    mc->pc = nullptr;
    if (make_tensor_map_2d(&mc->c, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, C, m, n, (uint64_t)ldc * 4, 32, 16, CU_TENSOR_MAP_SWIZZLE_NONE)) return 2;
    mc->pc = C;
    mc->ldc = ldc;
  }

  if (!fused) {
    const int split_blocks = ctx->sm_count * 8;
    int blocks_a = (int)((double)split_blocks * (double)countA / (double)(countA + countB));
    blocks_a = std::max(1, std::min(split_blocks - 1, blocks_a));
    split_tf32_kernel<<<split_blocks, 256, 0, ctx->compute>>>(A, a_hi, a_lo, countA, B, b_hi, b_lo, countB, blocks_a);
    B200_CUDA(cudaGetLastError());
  }

  // B200_SGEMM_TRACE=1: per-CTA phase timestamps of this launch, summarised on stderr (diagnostic:
  // synchronises the stream)
  static const bool trace_on = getenv("B200_SGEMM_TRACE") != nullptr;
  unsigned long long* trace = nullptr;
  const int trace_ctas = 2 * ctx->sm_count;
  if (trace_on) {
    B200_CUDA(cudaMalloc((void**)&trace, sizeof(unsigned long long) * 8 * trace_ctas));
    B200_CUDA(cudaMemsetAsync(trace, 0, sizeof(unsigned long long) * 8 * trace_ctas, ctx->compute));
  }
  auto trace_report = [&](const char* name, int grid) -> int {
    if (!trace) return 0;
    B200_CUDA(cudaStreamSynchronize(ctx->compute));
    std::vector<unsigned long long> h((size_t)8 * grid);
    B200_CUDA(cudaMemcpy(h.data(), trace, sizeof(unsigned long long) * 8 * grid, cudaMemcpyDeviceToHost));
    B200_CUDA(cudaFree(trace));
    unsigned long long t_first = ~0ull, t_last = 0;
    for (int c = 0; c < grid; c++) {
      if (h[c * 8 + 0] && h[c * 8 + 0] < t_first) t_first = h[c * 8 + 0];
      if (h[c * 8 + 6] > t_last) t_last = h[c * 8 + 6];
    }
    static const char* phase[6] = {"setup", "first operands", "MMA issue", "last chunk ready", "epilogue stores", "drain+exit"};
    fprintf(stderr, "[b200 sgemm trace] %s %dx%dx%d grid=%d span=%.2f us; per CTA (min / median / max us):\n", name, m, n, k, grid,
            (double)(t_last - t_first) * 1e-3);
    for (int p = 0; p < 6; p++) {
      std::vector<double> d;
      for (int c = 0; c < grid; c++)
        if (h[c * 8 + p] && h[c * 8 + p + 1]) d.push_back(((double)h[c * 8 + p + 1] - (double)h[c * 8 + p]) * 1e-3);
      if (d.empty()) continue;
      std::sort(d.begin(), d.end());
      fprintf(stderr, "    %-18s %8.2f %8.2f %8.2f   (%zu CTAs)\n", phase[p], d.front(), d[d.size() / 2], d.back(), d.size());
    }
    std::vector<double> e;
    for (int c = 0; c < grid; c++)
      if (h[c * 8]) e.push_back(((double)h[c * 8] - (double)t_first) * 1e-3);
    std::sort(e.begin(), e.end());
    if (!e.empty()) fprintf(stderr, "    CTA start after first CTA: median %.2f max %.2f us\n", e[e.size() / 2], e.back());
    return 0;
  };

  if (use_pair) {
    // Raster: groups of 8 m-tiles walk along n, so the ~74 tiles in flight share A and B panels in
    // L2.  With a short K there is nothing to share and the kernel is bound by the C writes: then
    // whole columns of tiles (all m-tiles of an n-tile first) keep the rows being written
    // contiguous in DRAM (M = N = 8192, K = 32: 58.1 -> 51.0 us; at K = 8192 the same order costs 3 %).
    static const char* rg_env = getenv("B200_SGEMM_RASTER");  // tuning: m-tiles per raster group
    const int raster_group = rg_env ? std::max(1, atoi(rg_env)) : (k <= 256 ? std::max(8, (int)t2m) : 8);
    const bool narrow = pair_n == 128;
    const long long ptiles = narrow ? tiles3 : tiles2;
    const int ptn = (int)(narrow ? t3n : t2n);
    const int grid = 2 * (int)(sk_tiles ? pairs : std::min<long long>(ptiles, pairs));
#define B200_PAIR_LAUNCH(F, PN)                                                                                        \
  pair::sgemm_3xtf32_pair_kernel<F, PN><<<grid, pair::THREADS, Rings<F, PN>::smem_bytes(pair::EPI_WARPS), ctx->compute>>>( \
      ma_hi, ma_lo, mb_hi, mb_lo, mc->c, tma_c, m, n, k, alpha, beta, C, (long long)ldc, (int)t2m, ptn, raster_group,  \
      sk_tiles, sk_slots, sk_partial, ctx->counters, trace)
    if (fused && narrow) B200_PAIR_LAUNCH(true, 128);
    else if (fused) B200_PAIR_LAUNCH(true, 256);
    else if (narrow) B200_PAIR_LAUNCH(false, 128);
    else B200_PAIR_LAUNCH(false, 256);
#undef B200_PAIR_LAUNCH
    B200_CUDA(cudaGetLastError());
    if (narrow)
      note_launch(ctx, fused ? "sgemm_3xtf32_fused_tcgen05_pair_256x128x32" : "sgemm_3xtf32_tcgen05_pair_256x128x32", fused ? 1 : 2);
    else if (fused)
      note_launch(ctx, sk_tiles ? "sgemm_3xtf32_fused_tcgen05_pair_256x256x32_streamk" : "sgemm_3xtf32_fused_tcgen05_pair_256x256x32", 1);
    else
      note_launch(ctx, sk_tiles ? "sgemm_3xtf32_tcgen05_pair_256x256x32_streamk" : "sgemm_3xtf32_tcgen05_pair_256x256x32", 2);
    return trace_report(ctx->last_kernel, grid);
  }
  std::atomic<bool>& g_attr_done = g_attr_done_dev[ctx->device % kMaxDevices];
  if (!g_attr_done) {
    B200_CUDA(cudaFuncSetAttribute(sgemm_3xtf32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)Rings<false>::smem_bytes(SINGLE_EPI_WARPS)));
    B200_CUDA(cudaFuncSetAttribute(sgemm_3xtf32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)Rings<true>::smem_bytes(SINGLE_EPI_WARPS)));
    g_attr_done = true;
  }
  float* partial = splits > 1 ? (float*)(ws + need) : nullptr;
  const int grid = (int)std::min<long long>(tiles1 * splits, ctx->sm_count);
  if (fused)
    sgemm_3xtf32_kernel<true><<<grid, NUM_THREADS, Rings<true>::smem_bytes(SINGLE_EPI_WARPS), ctx->compute>>>(
        ma_hi, ma_lo, mb_hi, mb_lo, mc->c, tma_c, m, n, k, alpha, beta, C, (long long)ldc, (int)t1m, (int)t1n, splits,
        kb_per_split, partial, ctx->counters, trace);
  else
    sgemm_3xtf32_kernel<false><<<grid, NUM_THREADS, Rings<false>::smem_bytes(SINGLE_EPI_WARPS), ctx->compute>>>(
        ma_hi, ma_lo, mb_hi, mb_lo, mc->c, tma_c, m, n, k, alpha, beta, C, (long long)ldc, (int)t1m, (int)t1n, splits,
        kb_per_split, partial, ctx->counters, trace);
  B200_CUDA(cudaGetLastError());
  if (fused)
    note_launch(ctx, splits > 1 ? "sgemm_3xtf32_fused_tcgen05_128x128x32_splitk" : "sgemm_3xtf32_fused_tcgen05_128x128x32", 1);
  else
    note_launch(ctx, splits > 1 ? "sgemm_3xtf32_tcgen05_128x128x32_splitk" : "sgemm_3xtf32_tcgen05_128x128x32", 2);
  return trace_report(ctx->last_kernel, grid);
}

// Scratch one SGEMM of this shape will need (pre-split schedules: the four TF32 split operands;
// always: partial-tile room), reserved outside the caller's timed region.
int sgemm_tc_prepare(b200_ctx* ctx, int m, int n, int k, int lda, int ldb) {
  if (ctx->opt_sgemm == 2 || (lda % 4) || (ldb % 4) || m < 64 || n < 64 || k < 1) return 0;
  if (ctx->opt_sgemm != 1 && (double)m * n * k < 1.0e8) return 0;
  SgemmPlan plan;
  const int rc = plan_sgemm(ctx, m, n, k, lda, ldb, &plan);
  if (rc) return rc < 0 ? 0 : rc;
  return ensure_workspace(ctx, plan.need + std::max(std::max(plan.part_bytes, plan.sk_bytes), (size_t)64 << 20));
}

}  // namespace b200
