// Validates the kind::tf32 tcgen05.mma operand conventions used by the node-level tensor-core kernels:
//   K-major SS, MN-major SS (A and B), and TS (A from tensor memory, one fp32 per column) -- against a CPU product.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gnn_jet_autoencoder_b200/csrc -I include -o tools/tf32_selftest.bin tools/tf32_selftest.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "tc2_common.cuh"
using namespace tc2;

__host__ __device__ inline uint32_t idesc_tf32(int M, int N, int a_mn, int b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;            // c_format = F32
  d |= 2u << 7;            // a_format = TF32
  d |= 2u << 10;           // b_format = TF32
  d |= (uint32_t)(a_mn & 1) << 15;
  d |= (uint32_t)(b_mn & 1) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// slab layout for 32-bit elements: element (row, c) of a buffer with `nrows` rows at (c / 4) * nrows * 16 + row * 16 + (c % 4) * 4
__host__ __device__ inline int sl_off(int row, int c, int nrows) { return (c >> 2) * nrows * 16 + row * 16 + (c & 3) * 4; }

// mode 0: D[128 x N] = A[128 x K] B[N x K]^T, both K-major SS.   mode 1: A from TMEM (TS), B K-major.
// mode 2: D[M x N] = X^T Y with X [K x M], Y [K x N] row tiles (both MN-major SS, K = 128 rows).   mode 3: TS, B MN-major (dgrad)
__global__ void __launch_bounds__(128, 1) test(int mode, int M, int N, int K, const float* A, const float* B, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 16);
  uint8_t* sa = smem + 1024;
  uint8_t* sb = smem + 1024 + 64 * 1024;
  const int tid = threadIdx.x;
  if (mode == 0 || mode == 1 || mode == 3) {
    for (int idx = tid; idx < 128 * K; idx += 128) { int r = idx / K, k = idx % K; *reinterpret_cast<float*>(sa + sl_off(r, k, 128)) = A[idx]; }
    if (mode == 3) { for (int idx = tid; idx < K * N; idx += 128) { int k = idx / N, n = idx % N; *reinterpret_cast<float*>(sb + sl_off(k, n, K)) = B[idx]; } }   // B given as [K][N] (weights [out=K][in=N])
    else { for (int idx = tid; idx < N * K; idx += 128) { int n = idx / K, k = idx % K; *reinterpret_cast<float*>(sb + sl_off(n, k, N)) = B[idx]; } }
  } else {
    for (int idx = tid; idx < 128 * M; idx += 128) { int r = idx / M, c = idx % M; *reinterpret_cast<float*>(sa + sl_off(r, c, 128)) = A[idx]; }
    for (int idx = tid; idx < 128 * N; idx += 128) { int r = idx / N, c = idx % N; *reinterpret_cast<float*>(sb + sl_off(r, c, 128)) = B[idx]; }
  }
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (tid < 32) tmem_alloc(slot, 512);
  fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = *slot;
  const int warp = tid >> 5, lane = tid & 31;
  const uint32_t lane_base = tb + ((uint32_t)(warp * 32) << 16);
  if (mode == 1 || mode == 3) {      // A into TMEM columns [256, 256 + K): thread = row, one fp32 per column
    for (int k0 = 0; k0 < K; k0 += 8) {
      uint32_t v[8];
      for (int q = 0; q < 8; ++q) v[q] = __float_as_uint(A[(warp * 32 + lane) * K + k0 + q]);
      tmem_st8(lane_base + 256 + k0, v);
    }
    tmem_st_wait();
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 0) {
    const uint32_t a = smem_u32(sa), b = smem_u32(sb);
    if (mode == 0 || mode == 1) {
      const uint32_t idesc = idesc_tf32(128, N, 0, 0);
      for (int s = 0; s < K / 8; ++s) {
        const uint64_t bd = make_smem_desc(b + s * 2 * N * 16, N * 16, 128);
        if (mode == 0) mma_tf32_ss(tb, make_smem_desc(a + s * 4096, 2048, 128), bd, idesc, s > 0);
        else mma_tf32_ts(tb, tb + 256 + 8 * s, bd, idesc, s > 0);
      }
    } else if (mode == 3) {
      // D[128 x N] = A[128 x K] W[K x N]: B = W viewed MN-major (MN = n contiguous, K = k rows of the [K][N] slab buffer)
      const uint32_t idesc = idesc_tf32(128, N, 0, 1);
      for (int s = 0; s < K / 8; ++s) mma_tf32_ts(tb, tb + 256 + 8 * s, make_smem_desc(b + s * 128, 128, K * 16), idesc, s > 0);
    } else {
      const uint32_t idesc = idesc_tf32(M, N, 1, 1);
      for (int s = 0; s < 16; ++s) mma_tf32_ss(tb, make_smem_desc(a + s * 128, 128, 2048), make_smem_desc(b + s * 128, 128, 2048), idesc, s > 0);
    }
    mma_commit_elect(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    tmem_ld8_u(lane_base + c0, v);
    tmem_ld_wait(); tmem_pin8(v);
    for (int q = 0; q < 8; ++q) out[(warp * 32 + lane) * N + c0 + q] = __uint_as_float(v[q]);
  }
  tc_fence_before(); __syncthreads();
  if (tid < 32) tmem_dealloc(tb, 512);
}

static float tf32r(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; float y; memcpy(&y, &u, 4); return y; }

int run(int mode, int M, int N, int K) {
  const int na = (mode == 2) ? 128 * M : 128 * K, nb = (mode == 2) ? 128 * N : N * K;
  std::vector<float> A(na), B(nb), out(128 * N), ref(128 * N, 0.f);
  for (auto& v : A) v = (rand() % 2001 - 1000) / 1000.f;
  for (auto& v : B) v = (rand() % 2001 - 1000) / 1000.f;
  const int Mrows = mode == 2 ? M : 128;
  for (int m = 0; m < Mrows; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      if (mode == 2) for (int r = 0; r < 128; ++r) s += (double)tf32r(A[r * M + m]) * tf32r(B[r * N + n]);
      else if (mode == 3) for (int k = 0; k < K; ++k) s += (double)tf32r(A[m * K + k]) * tf32r(B[k * N + n]);
      else for (int k = 0; k < K; ++k) s += (double)tf32r(A[m * K + k]) * tf32r(B[n * K + k]);
      ref[m * N + n] = (float)s;
    }
  float *dA, *dB, *dO;
  cudaMalloc(&dA, na * 4); cudaMalloc(&dB, nb * 4); cudaMalloc(&dO, 128 * N * 4);
  cudaMemcpy(dA, A.data(), na * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), nb * 4, cudaMemcpyHostToDevice);
  cudaMemset(dO, 0, 128 * N * 4);
  cudaFuncSetAttribute(test, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  test<<<1, 128, 200 * 1024>>>(mode, M, N, K, dA, dB, dO);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(out.data(), dO, 128 * N * 4, cudaMemcpyDeviceToHost);
  double num = 0, den = 0;
  for (int m = 0; m < Mrows; ++m) {
    // M = 64 accumulators: row m lives at lane (m / 16) * 32 + m % 16
    const int lane_row = (mode == 2 && M == 64) ? (m / 16) * 32 + m % 16 : m;
    for (int n = 0; n < N; ++n) { double d = out[lane_row * N + n] - ref[m * N + n]; num += d * d; den += (double)ref[m * N + n] * ref[m * N + n]; }
  }
  printf("mode %d M%-3d N%-3d K%-3d rel err %.3e %s\n", mode, M, N, K, std::sqrt(num / den), e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(dA); cudaFree(dB); cudaFree(dO);
  return 0;
}

int main() {
  run(0, 128, 32, 48); run(0, 128, 16, 32); run(0, 128, 64, 16);
  run(1, 128, 32, 48); run(1, 128, 16, 8);
  run(3, 128, 32, 16); run(3, 128, 48, 32); run(3, 128, 16, 8);
  run(2, 64, 56, 128); run(2, 64, 40, 128); run(2, 128, 32, 128); run(2, 64, 24, 128);
  return 0;
}
