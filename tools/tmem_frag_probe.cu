// Prints the register <-> (lane, column) mapping of tcgen05.ld.16x256b (matrix-fragment shape): TMEM is filled through
// 32x32b stores with value = lane * 100 + column, then read back with 16x256b.x2 at lane offsets 0 and 16.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_frag_probe.bin tools/tmem_frag_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(float* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)(warp * 32) << 16);
  uint32_t r[16];
  for (int c = 0; c < 16; ++c) r[c] = __float_as_uint((float)((warp * 32 + lane) * 100 + c));
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(base), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                 "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncthreads();
  uint32_t a[8], b[8];
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]) : "r"(base));
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]) : "r"(base + (16u << 16)));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 8; ++i) { out[threadIdx.x * 16 + i] = __uint_as_float(a[i]); out[threadIdx.x * 16 + 8 + i] = __uint_as_float(b[i]); }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(slot) : "memory");
}
int main() {
  float* d; cudaMalloc(&d, 128 * 16 * 4);
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  static float h[128 * 16];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("status: %s\nthread: register i -> lane*100+column (first load: lane offset 0, second: lane offset 16)\n", cudaGetErrorString(e));
  for (int t : {0, 1, 2, 3, 4, 5, 31, 32, 37}) {
    printf("t%3d:", t);
    for (int i = 0; i < 16; ++i) printf(" %5.0f", h[t * 16 + i]);
    printf("\n");
  }
  return 0;
}
