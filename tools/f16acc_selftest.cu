// Probes tcgen05.mma kind::f16 with an FP16 accumulator (c_format = F16): where the results sit in tensor memory, what
// tcgen05.ld returns with and without .pack::16b, and whether a packed fp16 activation tile written back with tcgen05.st
// works as the A operand of a TS-form MMA (a_format = F16, b_format = BF16).  Motivation: DESIGN.md "not yet tried" -- the
// edge-kernel epilogues could drop their cvt.rn.bf16x2 if the accumulator came back as packed halves.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gnn_jet_autoencoder_b200/csrc -I include -o tools/f16acc_selftest.bin tools/f16acc_selftest.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "tc2_common.cuh"
using namespace tc2;

__host__ __device__ inline uint32_t idesc(int M, int N, int cfmt, int afmt, int bfmt) {
  uint32_t d = 0;
  d |= (uint32_t)cfmt << 4;      // 0 = F16, 1 = F32
  d |= (uint32_t)afmt << 7;      // 0 = F16, 1 = BF16
  d |= (uint32_t)bfmt << 10;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}
// 16-bit slab layout: element (row, c) at (c / 8) * nrows * 16 + row * 16 + (c % 8) * 2
__host__ __device__ inline int sl16(int row, int c, int nrows) { return (c >> 3) * nrows * 16 + row * 16 + (c & 7) * 2; }

__device__ __forceinline__ void ld16_pack(uint32_t taddr, uint32_t (&r)[16]) {      // 32 columns of 16-bit data -> 16 registers
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}

// out_raw [128][64]: raw 32-bit TMEM columns 0..63 of every lane after D = A B^T (M = 128, N = 32, K = 32)
// out_pack [128][16]: the same accumulator through .pack::16b (columns 0..31)
// out_ts [128][32] fp32: second GEMM D2 = leaky-free copy: A2 = D (packed fp16, written back to TMEM columns 128..143), B2 = B (N = 32, K = 32)
__global__ void __launch_bounds__(128, 1) probe(int bfmt, int use_pack, int do_ts, int cfmt, const float* A, const float* B, uint32_t* out_raw,
                                                uint32_t* out_pack, float* out_ts) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 16);
  uint8_t* sa = smem + 1024;             // A: fp16 [128][32]
  uint8_t* sb = smem + 1024 + 16384;     // B: bf16 [32][32]
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int idx = tid; idx < 128 * 32; idx += 128) { int r = idx / 32, k = idx % 32; *reinterpret_cast<__half*>(sa + sl16(r, k, 128)) = __float2half_rn(A[idx]); }
  for (int idx = tid; idx < 32 * 32; idx += 128) {
    int n = idx / 32, k = idx % 32;
    if (bfmt) *reinterpret_cast<__nv_bfloat16*>(sb + sl16(n, k, 32)) = __float2bfloat16_rn(B[idx]);
    else *reinterpret_cast<__half*>(sb + sl16(n, k, 32)) = __float2half_rn(B[idx]);
  }
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (tid < 32) tmem_alloc(slot, 256);
  fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = *slot, lane_base = tb + ((uint32_t)(warp * 32) << 16);
  {      // clear columns 0..63 so that untouched halves are visible
    uint32_t z[16];
    for (int c = 0; c < 16; ++c) z[c] = 0xDEAD0000u;
    for (int c0 = 0; c0 < 64; c0 += 16) tmem_st16(lane_base + c0, z);
    tmem_st_wait();
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t sa_u = smem_u32(sa), sb_u = smem_u32(sb);
  if (warp == 0) {
    const uint32_t id = idesc(128, 32, cfmt, 0, bfmt);      // F16 (or F32) accumulator, A fp16, B fp16 / bf16
    for (int s = 0; s < 2; ++s)
      mma_bf16_ss_elect(tb, make_smem_desc(sa_u + s * 4096, 2048, 128), make_smem_desc(sb_u + s * 2 * 32 * 16, 32 * 16, 128), id, s > 0);
    mma_commit_elect(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  uint32_t v[16], pk[16];
  for (int c0 = 0; c0 < 64; c0 += 16) {
    tmem_ld16_u(lane_base + c0, v);
    tmem_ld_wait(); tmem_pin16(v);
    for (int q = 0; q < 16; ++q) out_raw[tid * 64 + c0 + q] = v[q];
  }
  if (cfmt) { for (int q = 0; q < 16; ++q) pk[q] = bf2_as_u32(__floats2bfloat162_rn(0.f, 0.f)), pk[q] = (uint32_t)__half_as_ushort(__float2half_rn(__uint_as_float(out_raw[tid * 64 + 2 * q]))) | ((uint32_t)__half_as_ushort(__float2half_rn(__uint_as_float(out_raw[tid * 64 + 2 * q + 1]))) << 16); }
  else if (use_pack) { ld16_pack(lane_base, pk); tmem_ld_wait(); tmem_pin16(pk); }
  else { for (int q = 0; q < 16; ++q) pk[q] = (out_raw[tid * 64 + 2 * q] & 0xffffu) | (out_raw[tid * 64 + 2 * q + 1] << 16); }
  for (int q = 0; q < 16; ++q) out_pack[tid * 16 + q] = pk[q];
  // packed halves back into TMEM columns 128..143 as the A operand (K = 32) of a TS MMA with B again
  tmem_st16(lane_base + 128, pk);
  tmem_st_wait();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 0 && do_ts) {
    const uint32_t id = idesc(128, 32, 1, 0, bfmt);      // F32 accumulator this time, A fp16 from TMEM, B as before
    for (int s = 0; s < 2; ++s)
      mma_ts_elect(tb + 192, tb + 128 + s * 8, make_smem_desc(sb_u + s * 2 * 32 * 16, 32 * 16, 128), id, s > 0);
    mma_commit_elect(bar);
  }
  if (do_ts) mbar_wait(bar, 1);
  tc_fence_after();
  for (int c0 = 0; c0 < 32; c0 += 16) {
    tmem_ld16_u(lane_base + 192 + c0, v);
    tmem_ld_wait(); tmem_pin16(v);
    for (int q = 0; q < 16; ++q) out_ts[tid * 32 + c0 + q] = __uint_as_float(v[q]);
  }
  tc_fence_before(); __syncthreads();
  if (tid < 32) tmem_dealloc(tb, 256);
}

int main(int argc, char** argv) {
  const int bfmt = argc > 1 ? atoi(argv[1]) : 1, use_pack = argc > 2 ? atoi(argv[2]) : 1, do_ts = argc > 3 ? atoi(argv[3]) : 1, cfmt = argc > 4 ? atoi(argv[4]) : 0;
  printf("B format %s, pack::16b load %d, TS second GEMM %d, first accumulator %s\n", bfmt ? "bf16" : "fp16", use_pack, do_ts, cfmt ? "F32" : "F16");
  std::vector<float> A(128 * 32), B(32 * 32);
  srand(1);
  for (auto& x : A) x = (rand() % 2001 - 1000) / 1000.f;
  for (auto& x : B) x = (rand() % 2001 - 1000) / 1000.f;
  float *dA, *dB, *dts; uint32_t *draw, *dpk;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&draw, 128 * 64 * 4); cudaMalloc(&dpk, 128 * 16 * 4); cudaMalloc(&dts, 128 * 32 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  probe<<<1, 128, 64 * 1024>>>(bfmt, use_pack, do_ts, cfmt, dA, dB, draw, dpk, dts);
  cudaError_t ce = cudaDeviceSynchronize();
  if (ce != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(ce)); return 1; }
  std::vector<uint32_t> raw(128 * 64), pk(128 * 16); std::vector<float> ts(128 * 32);
  cudaMemcpy(raw.data(), draw, raw.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(pk.data(), dpk, pk.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(ts.data(), dts, ts.size() * 4, cudaMemcpyDeviceToHost);
  auto h2f = [](uint16_t h) { __half_raw r; r.x = h; return __half2float(__half(r)); };
  auto bf = [&](float x) { if (!bfmt) return __half2float(__float2half_rn(x)); uint32_t u; memcpy(&u, &x, 4); u = (u + 0x7fff + ((u >> 16) & 1)) & 0xffff0000u; float y; memcpy(&y, &u, 4); return y; };
  auto hf = [&](float x) { return __half2float(__float2half_rn(x)); };
  // reference D[r][n] = sum_k half(A[r][k]) * bf16(B[n][k])
  std::vector<double> D(128 * 32);
  for (int r = 0; r < 128; ++r) for (int n = 0; n < 32; ++n) { double s = 0; for (int k = 0; k < 32; ++k) s += (double)hf(A[r * 32 + k]) * bf(B[n * 32 + k]); D[r * 32 + n] = s; }
  printf("row 0, reference D[0][0..7]:"); for (int n = 0; n < 8; ++n) printf(" %.4f", D[n]); printf("\n");
  printf("row 0, raw TMEM columns 0..11 (hex):"); for (int c = 0; c < 12; ++c) printf(" %08x", raw[c]); printf("\n");
  printf("row 0, raw columns 28..35 (hex):"); for (int c = 28; c < 36; ++c) printf(" %08x", raw[c]); printf("\n");
  printf("row 0, low halves of columns 0..7 as fp16:"); for (int c = 0; c < 8; ++c) printf(" %.4f", h2f(raw[c] & 0xffff)); printf("\n");
  printf("row 0, high halves of columns 0..7 as fp16:"); for (int c = 0; c < 8; ++c) printf(" %.4f", h2f(raw[c] >> 16)); printf("\n");
  printf("row 0, .pack::16b regs 0..3 (lo, hi):"); for (int q = 0; q < 4; ++q) printf(" (%.4f, %.4f)", h2f(pk[q] & 0xffff), h2f(pk[q] >> 16)); printf("\n");
  // hypothesis A: column n holds D[n] in its low half; pack reg q = (D[2q], D[2q+1])
  double errA = 0, errP = 0, errT = 0;
  for (int r = 0; r < 128; ++r) for (int n = 0; n < 32; ++n) {
    errA = fmax(errA, fabs(h2f(raw[r * 64 + n] & 0xffff) - D[r * 32 + n]));
    const uint32_t w = pk[r * 16 + n / 2];
    errP = fmax(errP, fabs(h2f((n & 1) ? (w >> 16) : (w & 0xffff)) - D[r * 32 + n]));
  }
  // TS check: D2[r][n] = sum_k half(D[r][k]) * bf16(B[n][k])
  for (int r = 0; r < 128; ++r) for (int n = 0; n < 32; ++n) {
    double s = 0; for (int k = 0; k < 32; ++k) s += (double)hf((float)D[r * 32 + k]) * bf(B[n * 32 + k]);
    errT = fmax(errT, fabs(ts[r * 32 + n] - s));
  }
  if (cfmt) { errA = errP = 0; for (int r = 0; r < 128; ++r) for (int n = 0; n < 32; ++n) { float f; memcpy(&f, &raw[r * 64 + n], 4); errA = fmax(errA, fabs(f - D[r * 32 + n])); } }
  printf("max |low-half(column n) - D[n]| = %.3e   max |pack::16b pair - D| = %.3e   TS-from-packed-fp16 max err = %.3e\n", errA, errP, errT);
  return 0;
}
