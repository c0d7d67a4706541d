// Micro-benchmark of tcgen05.mma issue/execution cost for the operand layouts and shapes of the edge kernels.
// Build here: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gnn_jet_autoencoder_b200/csrc -o tools/umma_bench.bin tools/umma_bench.cu
// Run on the GPU box: tools/umma_bench.bin
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace umma;

// layout: 0 = SWIZZLE_NONE interleaved (K-major both), 1 = SWIZZLE_NONE MN-major both, 2 = SWIZZLE_128B K-major both
__global__ void __launch_bounds__(128, 1) bench(int M, int N, int nk, int layout, int reps, int per_batch_wait, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 16);
  for (int i = threadIdx.x; i < 48 * 1024; i += 128) reinterpret_cast<uint32_t*>(smem + 1024)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(slot, 256);
  fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = *slot;
  const int warp = (int)uni(threadIdx.x >> 5);
  if (warp == 0) {
    const uint32_t sa = smem_u32(smem + 1024), sb = smem_u32(smem + 1024 + 64 * 1024);
    uint64_t ad, bd; uint32_t astep, bstep, idesc;
    if (layout == 0) {
      ad = make_smem_desc(sa, 128 * 16, 128); bd = make_smem_desc(sb, N * 16, 128); astep = 4096u >> 4; bstep = (2u * N * 16) >> 4;
      idesc = make_idesc_bf16(M, N, 0, 0);
    } else if (layout == 1) {
      ad = make_smem_desc(sa, 128, 2048); bd = make_smem_desc(sb, 128, 2048); astep = 256u >> 4; bstep = 256u >> 4;
      idesc = make_idesc_bf16(M, N, 1, 1);
    } else {
      ad = make_smem_desc(sa, 16, 1024) | (2ull << 61); bd = make_smem_desc(sb, 16, 1024) | (2ull << 61); astep = 32u >> 4; bstep = 32u >> 4;
      idesc = make_idesc_bf16(M, N, 0, 0);
    }
    uint32_t phase = 0;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int ks = 0; ks < nk; ++ks) {
        // SW128: 4 k-steps per 128-byte row, then the next 64-column block (16 KB further)
        uint32_t ao = layout == 2 ? (uint32_t)((ks & 3) * 2 + (ks >> 2) * 1024) : astep * ks;
        uint32_t bo = layout == 2 ? (uint32_t)((ks & 3) * 2 + (ks >> 2) * 1024) : bstep * ks;
        mma_bf16_ss_elect(tb, ad + ao, bd + bo, idesc, ks > 0);
      }
      if (per_batch_wait) { mma_commit_elect(bar); mbar_wait(bar, phase); phase ^= 1; tc_fence_after(); }
    }
    if (!per_batch_wait) { mma_commit_elect(bar); mbar_wait(bar, phase); }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 256);
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct { int M, N, nk; const char* name; } shapes[] = {
      {128, 128, 2, "F1  M128 N128 K32"}, {128, 64, 8, "F2  M128 N64 K128"}, {128, 16, 4, "F3  M128 N16 K64"},
      {128, 32, 8, "wg1/dg1 M128 N32 K128"}, {128, 64, 8, "wg2 M128 N64 K128"}, {64, 16, 8, "wg3 M64 N16 K128"},
      {128, 128, 4, "dg2 M128 N128 K64"}, {128, 16, 8, "colsum M128 N16 K128"}, {128, 256, 8, "ref M128 N256 K128"}};
  const char* lay[] = {"none/K-major", "none/MN-major", "sw128/K-major"};
  for (int grid : {1, 148})
    for (auto& s : shapes)
      for (int layout = 0; layout < 3; ++layout)
        for (int pbw = 0; pbw < 2; ++pbw) {
          const int reps = 200;
          bench<<<grid, 128, 200 * 1024>>>(s.M, s.N, s.nk, layout, reps, pbw, d);
          cudaError_t e = cudaDeviceSynchronize();
          long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
          long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
          printf("grid %3d  %-22s %-14s %-16s %8.1f cyc/MMA  %8.1f cyc/batch  %s\n", grid, s.name, lay[layout],
                 pbw ? "commit+wait/batch" : "back-to-back", (double)mx / reps / s.nk, (double)mx / reps,
                 e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
  return 0;
}
