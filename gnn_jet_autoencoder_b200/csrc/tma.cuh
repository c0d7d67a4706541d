// Tensor-memory-accelerator (TMA) staging: cp.async.bulk.tensor copies described by a CUtensorMap, completing on an
// mbarrier (expect_tx), issued by ONE thread -- the edge kernels stage their packed parameter image and their P_i chunks
// this way instead of per-thread 16-byte copies.  The tensor map is built on the host (gj_tmap_2d, tma_host.cu: the driver's
// cuTensorMapEncodeTiled resolved through cudaGetDriverEntryPoint, so the library does not link against libcuda) and passed
// to the kernel as a `const __grid_constant__ CUtensorMap` parameter.
#pragma once
#include <cuda.h>
#include <stdint.h>

#include "umma.cuh"

// host: 2-D tensor of 32-bit elements, `width` x `height`, rows `row_stride_bytes` apart; one copy moves a box_w x box_h box
// (box_w, box_h <= 256; the box lands densely, row after row, at the shared-memory destination).  Returns GJ_OK / GJ_ERR_CUDA.
int gj_tmap_2d(CUtensorMap* map, const void* base, uint64_t width, uint64_t height, uint64_t row_stride_bytes, uint32_t box_w,
               uint32_t box_h);

namespace tma {
using umma::smem_u32;

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// box whose first element is (x, y) -> dense rows at dst (128-byte aligned); arrives with its byte count on `bar`
__device__ __forceinline__ void load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
}  // namespace tma
