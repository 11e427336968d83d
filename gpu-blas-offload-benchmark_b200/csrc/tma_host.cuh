// tma_host.cuh -- host-side helper: encode a 2-D TMA tensor map through the driver
// entry point (no link-time dependency on libcuda).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace b200 {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

inline EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D tensor, dim0 contiguous, row pitch ld_bytes; box {box0, box1}; out-of-bounds
// elements read as zero.
inline int make_tensor_map_2d(CUtensorMap* map, CUtensorMapDataType dtype, const void* ptr, uint64_t dim0,
                              uint64_t dim1, uint64_t ld_bytes, uint32_t box0, uint32_t box1,
                              CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = tensor_map_encoder();
  B200_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t dims[2] = {dim0, dim1};
  const cuuint64_t strides[1] = {ld_bytes};
  const cuuint32_t box[2] = {box0, box1};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, dtype, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

}  // namespace b200
