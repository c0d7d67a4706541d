// Host side of tma.cuh: tensor-map encoding through the driver entry point (no link-time dependency on libcuda).
#include <cuda_runtime.h>

#include "gj_common.cuh"
#include "tma.cuh"

void gj_set_error(const char* fmt, ...);

typedef CUresult (*gj_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static gj_encode_tiled_fn encode_tiled() {
  static gj_encode_tiled_fn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<gj_encode_tiled_fn>(p);
  }();
  return fn;
}

int gj_tmap_2d(CUtensorMap* map, const void* base, uint64_t width, uint64_t height, uint64_t row_stride_bytes, uint32_t box_w,
               uint32_t box_h) {
  gj_encode_tiled_fn enc = encode_tiled();
  if (!enc) { gj_set_error("cuTensorMapEncodeTiled is not available from this driver"); return GJ_ERR_CUDA; }
  const cuuint64_t dims[2] = {width, height};
  const cuuint64_t strides[1] = {row_stride_bytes};      // stride of dimension 1; dimension 0 is dense
  const cuuint32_t box[2] = {box_w, box_h};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { gj_set_error("cuTensorMapEncodeTiled failed with %d (base %p, %llu x %llu, stride %llu, box %u x %u)", (int)r, base,
                                        (unsigned long long)width, (unsigned long long)height, (unsigned long long)row_stride_bytes, box_w, box_h);
                           return GJ_ERR_CUDA; }
  return GJ_OK;
}
