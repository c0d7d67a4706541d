// Execution-rate micro-benchmark of tcgen05.mma (kind::f16, bf16) with fully unrolled issue: cycles per MMA for the operand
// sources / shapes of the edge kernels (SS = both operands from shared memory, TS = A from tensor memory).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gnn_jet_autoencoder_b200/csrc -I include -o tools/umma_bench2.bin tools/umma_bench2.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc2_common.cuh"
using namespace tc2;

// MODE 0: SS K-major A, K-major B    1: SS MN-major A, MN-major B (wgrad)    2: TS, K-major B     3: TS, MN-major B (dgrad)
template <int M, int N, int MODE>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 16);
  for (int i = threadIdx.x; i < 40 * 1024; i += 128) reinterpret_cast<uint32_t*>(smem + 1024)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(slot, 512);
  fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = *slot;
  const int warp = (int)uni(threadIdx.x >> 5);
  if (warp == 0) {
    const uint32_t sa = smem_u32(smem + 1024), sb = smem_u32(smem + 1024 + 64 * 1024);
    uint64_t ad, bd; uint32_t astep, bstep;
    const bool amn = MODE == 1, bmn = MODE == 1 || MODE == 3;
    ad = amn ? make_smem_desc(sa, 128, 2048) : make_smem_desc(sa, 2048, 128);
    astep = amn ? 256u >> 4 : 4096u >> 4;
    bd = bmn ? make_smem_desc(sb, 128, 2048) : make_smem_desc(sb, N * 16, 128);
    bstep = bmn ? 256u >> 4 : (2u * N * 16) >> 4;
    const uint32_t idesc = make_idesc_bf16(M, N, amn, bmn);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        if (MODE < 2) mma_bf16_ss_elect(tb, ad + (uint64_t)(astep * ks), bd + (uint64_t)(bstep * ks), idesc, 1u);
        else mma_ts_elect(tb, tb + 256 + 8 * ks, bd + (uint64_t)(bstep * ks), idesc, 1u);
      }
    }
    mma_commit_elect(bar); mbar_wait(bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

template <int M, int N, int MODE>
void run(const char* name) {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench<M, N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int reps = 200;
  bench<M, N, MODE><<<1, 128, 200 * 1024>>>(d, reps);
  cudaDeviceSynchronize();
  bench<M, N, MODE><<<1, 128, 200 * 1024>>>(d, reps);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const char* modes[] = {"SS K/K", "SS MN/MN", "TS K", "TS MN"};
  printf("%-26s M%-3d N%-3d %-9s %7.1f cyc/MMA (floor %d)  %s\n", name, M, N, modes[MODE], (double)h / (reps * 8), (M > 128 ? M : 128) * N / 256,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<128, 128, 0>("fwd L1 (a0 smem)");
  run<128, 128, 2>("fwd L1 (TS)");
  run<128, 64, 2>("fwd L2 (TS)");
  run<128, 64, 0>("fwd L2 (SS)");
  run<128, 16, 2>("fwd L3 (TS)");
  run<128, 16, 0>("fwd L3 (SS)");
  run<128, 64, 3>("dgrad3 (TS, W MN)");
  run<128, 128, 3>("dgrad2 (TS, W MN)");
  run<128, 128, 0>("dgrad2 (SS)");
  run<128, 32, 3>("dgrad1 (TS, W MN)");
  run<128, 32, 1>("wgrad1 M128 N32");
  run<128, 48, 1>("wgrad1+bias M128 N48");
  run<128, 64, 1>("wgrad2 M128 N64");
  run<64, 128, 1>("wgrad2' M64 N128");
  run<64, 16, 1>("wgrad3 M64 N16");
  run<64, 8, 1>("colsum M64 N8");
  run<128, 16, 1>("colsum M128 N16");
  run<128, 256, 0>("ref M128 N256 SS");
  run<128, 256, 2>("ref M128 N256 TS");
  return 0;
}
