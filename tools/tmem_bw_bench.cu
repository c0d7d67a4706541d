// Tensor-memory load / store throughput per SM (what bounds the epilogues of the fused edge kernels: every accumulator
// element of a hidden layer has to pass through registers once).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_bw_bench.bin tools/tmem_bw_bench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define REPS 64

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD16(MOD) \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16" MOD ".b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), \
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(addr))

// OP 0: ld x16 (16 columns)   1: ld x16.pack::16b (32 columns -> 16 registers)   2: st x16   3: ld x8   4: ld 16x256b.x4 (16 regs)
template <int OP>
__global__ void __launch_bounds__(512, 1) bench(long long* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = threadIdx.x + i;
  // initialise the columns that are read
  for (int c = 0; c < 512; c += 16) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(base + c), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                   "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncthreads();
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int rep = 0; rep < REPS; ++rep) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint32_t addr = base + (uint32_t)(((rep * 8 + u) * 32 + (warp >> 2) * 128) & 511 & ~31);
      if (OP == 0) LD16("");
      if (OP == 1) LD16(".pack::16b");
      if (OP == 2)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                     ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                       "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
      if (OP == 3)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr));
      if (OP == 4)
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(addr));
      if (OP != 2) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) acc += r[i];
      }
    }
    if (OP == 2) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  const long long t1 = clock64();
  if (acc == 0x12345678u) out[1] = 1;
  __syncthreads();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

template <int OP>
void run(const char* name, int nthreads, int bytes_per_lane) {
  long long* d; cudaMalloc(&d, 16);
  bench<OP><<<1, nthreads>>>(d);
  cudaDeviceSynchronize();
  bench<OP><<<1, nthreads>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double n = REPS * 8.0, warps = nthreads / 32;
  printf("%-34s %3d thr (%d warp(s) per lane quadrant): %7.1f cyc per warp-instruction, %6.1f B/cyc/SM  %s\n", name, nthreads, nthreads / 128,
         (double)h / n, n * warps * 32 * bytes_per_lane / (double)h, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int nt : {128, 256, 384, 512}) {
    run<0>("tcgen05.ld 32x32b.x16", nt, 64);
    run<1>("tcgen05.ld 32x32b.x16.pack::16b", nt, 128);
    run<3>("tcgen05.ld 32x32b.x8", nt, 32);
    run<4>("tcgen05.ld 16x256b.x4", nt, 64);
    run<2>("tcgen05.st 32x32b.x16", nt, 64);
  }
  return 0;
}
