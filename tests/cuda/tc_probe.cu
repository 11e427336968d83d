// tc_probe.cu -- bring-up probe for the tcgen05 / TMA building blocks used by
// csrc/sgemm_tc.cu.  One CTA, one 128x128x32 tile, small-integer inputs (exact
// in TF32): dumps what TMA put in shared memory and what the MMA left in TMEM
// for a list of descriptor variants, and reports which variant reproduces
// C = A*B.  Run on the GPU box:  nvcc -arch=sm_100a tc_probe.cu -o tc_probe && ./tc_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int BM = 128, BN = 128, BK = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done, polls = 0;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done && ++polls > (1u << 22)) return false;
  } while (!done);
  return true;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }

struct Variant {
  uint32_t a_lbo, a_sbo, a_kstep;  // bytes
  uint32_t b_lbo, b_sbo, b_kstep;
  uint32_t idesc;
  uint32_t a_layout;  // UMMA layout type of A: 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  uint32_t use_a32;   // load A through the ATOM_32B tensor map
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a32,
             const __grid_constant__ CUtensorMap map_b, Variant v,
             float* __restrict__ smem_dump, float* __restrict__ d_out, int* __restrict__ status) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = raw + (base - smem_u32(raw));
  const uint32_t sa = base, sb = base + 16384, bars = base + 32768;
  const uint32_t full = bars, mma_done = bars + 8, slot = bars + 16;
  volatile uint32_t* slot_ptr = (volatile uint32_t*)(base_ptr + 32768 + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    mbar_init(mma_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot_ptr;
  if (threadIdx.x == 0) {
    status[1] = (int)tmem;
    mbar_expect_tx(full, 32768);
    for (int c = 0; c < 4; c++) tma_load_2d(sa + c * 4096, v.use_a32 ? &map_a32 : &map_a, full, c * 32, 0);
    tma_load_2d(sb, &map_b, full, 0, 0);
  }
  bool ok = mbar_wait(full, 0);
  if (!ok) { if (threadIdx.x == 0) status[0] = 1; }
  __syncthreads();
  // dump raw smem (A then B)
  const float* sp = (const float*)base_ptr;
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) smem_dump[i] = sp[i];
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0 && ok) {
    for (int ks = 0; ks < 4; ks++) {
      auto desc = [](uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
        return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
               ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
      };
      umma_tf32(tmem, desc(sa + ks * v.a_kstep, v.a_lbo, v.a_sbo, v.a_layout), desc(sb + ks * v.b_kstep, v.b_lbo, v.b_sbo, 2),
                v.idesc, ks != 0);
    }
    umma_commit(mma_done);
  }
  bool ok2 = mbar_wait(mma_done, 0);
  if (!ok2 && threadIdx.x == 0) status[0] |= 2;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // read D: thread (warp q, lane) = row q*32+lane, 128 columns
  for (int c0 = 0; c0 < BN; c0 += 32) {
    uint32_t r[32];
    const uint32_t taddr = tmem + c0 + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; j++) d_out[(size_t)(c0 + j) * BM + warp * 32 + lane] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  printf("encode fn %p query %d\n", fp, (int)q);
  const int M = BM, N = BN, K = BK;
  std::vector<float> A(M * K), B(K * N), C(M * N);
  srand(1);
  for (auto& x : A) x = (float)(rand() % 8);
  for (auto& x : B) x = (float)(rand() % 8);
  for (int j = 0; j < N; j++)
    for (int i = 0; i < M; i++) {
      float s = 0;
      for (int p = 0; p < K; p++) s += A[i + p * M] * B[p + j * K];
      C[i + j * M] = s;
    }
  float *dA, *dB, *dump, *dD; int* status;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4));
  CK(cudaMalloc(&dump, 8192 * 4)); CK(cudaMalloc(&dD, M * N * 4)); CK(cudaMalloc(&status, 16));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CUtensorMap ma, ma32, mb;
  {
    cuuint64_t dims[2] = {(cuuint64_t)M, (cuuint64_t)K}; cuuint64_t str[1] = {(cuuint64_t)M * 4};
    cuuint32_t box[2] = {32, 32}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode A -> %d\n", (int)r);
    r = enc(&ma32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode A (atom32) -> %d\n", (int)r);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N}; cuuint64_t str[1] = {(cuuint64_t)K * 4};
    cuuint32_t box[2] = {32, 128}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode B -> %d\n", (int)r);
  }
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000));
  const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  struct Named { const char* name; Variant v; };
  const uint32_t MN = idesc_base | (1u << 15);
  std::vector<Named> variants = {
      {"A sw128 (expect zeros)", {4096, 1024, 1024, 16, 1024, 32, MN, 2, 0}},
      {"A atom32: lbo4096 sbo512 kstep1024", {4096, 512, 1024, 16, 1024, 32, MN, 1, 1}},
      {"A atom32: lbo512 sbo4096 kstep1024", {512, 4096, 1024, 16, 1024, 32, MN, 1, 1}},
      {"A atom32: lbo4096 sbo1024 kstep1024", {4096, 1024, 1024, 16, 1024, 32, MN, 1, 1}},
      {"A atom32 data but sw128 desc", {4096, 512, 1024, 16, 1024, 32, MN, 2, 1}},
  };
  std::vector<float> hdump(8192), hD(M * N);
  for (size_t vi = 0; vi < variants.size(); vi++) {
    CK(cudaMemset(dD, 0xff, M * N * 4)); CK(cudaMemset(status, 0, 16)); CK(cudaMemset(dump, 0, 8192 * 4));
    probe_kernel<<<1, 128, 40000>>>(ma, ma32, mb, variants[vi].v, dump, dD, status);
    cudaError_t e = cudaDeviceSynchronize();
    int hs[4];
    if (e != cudaSuccess) { printf("variant %zu: kernel error %s\n", vi, cudaGetErrorString(e)); return 1; }
    CK(cudaMemcpy(hs, status, 16, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hdump.data(), dump, 8192 * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
    if (vi == 0) {
      // verify the TMA + swizzle picture of A and B
      int badA = 0, badB = 0; double sumA = 0, sumB = 0;
      for (int c = 0; c < 4; c++)
        for (int kk = 0; kk < 32; kk++)
          for (int j = 0; j < 8; j++)
            for (int e2 = 0; e2 < 4; e2++) {
              const int m = c * 32 + 4 * (j ^ (kk & 7)) + e2;
              const float got = hdump[c * 1024 + kk * 32 + j * 4 + e2];
              sumA += got;
              if (got != A[m + kk * M]) badA++;
            }
      for (int nn = 0; nn < 128; nn++)
        for (int j = 0; j < 8; j++)
          for (int e2 = 0; e2 < 4; e2++) {
            const int kk = 4 * (j ^ (nn & 7)) + e2;
            const float got = hdump[4096 + nn * 32 + j * 4 + e2];
            sumB += got;
            if (got != B[kk + nn * K]) badB++;
          }
      printf("TMA check: status=%d tmem_base=0x%x A mismatches=%d (sum %.0f) B mismatches=%d (sum %.0f)\n", hs[0], hs[1], badA, sumA, badB, sumB);
    }
    if (vi == 1) {
      int badA = 0;
      for (int c = 0; c < 4; c++)
        for (int kk = 0; kk < 32; kk++)
          for (int j = 0; j < 4; j++)
            for (int e2 = 0; e2 < 8; e2++) {
              const int m = c * 32 + 8 * (j ^ (kk & 3)) + e2;
              if (hdump[c * 1024 + kk * 32 + j * 8 + e2] != A[m + kk * M]) badA++;
            }
      printf("TMA atom32 check: A mismatches=%d\n", badA);
    }
    int bad = 0; double maxd = 0; double sumD = 0;
    for (int i = 0; i < M * N; i++) { double d = fabs((double)hD[i] - C[i]); if (d > 1e-3) bad++; if (d > maxd) maxd = d; sumD += hD[i]; }
    printf("variant %zu [%s]: status=%d mismatches=%d/%d maxdiff=%.1f sumD=%.0f  D[0]=%.0f C[0]=%.0f D[1]=%.0f C[1]=%.0f D[128]=%.0f C[128]=%.0f\n",
           vi, variants[vi].name, hs[0], bad, M * N, maxd, sumD, hD[0], C[0], hD[1], C[1], hD[128], C[128]);
  }
  return 0;
}
