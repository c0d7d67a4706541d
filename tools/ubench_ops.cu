// Issue-rate micro-benchmark of the instructions the tensor-core epilogues are made of (per SM sub-partition).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_ops.bin tools/ubench_ops.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define REPS 256
#define ILP 8

template <int OP>
__global__ void __launch_bounds__(512, 1) bench(long long* out, float seed) {
  float x[2 * ILP];
  uint32_t u[ILP];
#pragma unroll
  for (int i = 0; i < 2 * ILP; ++i) x[i] = seed + threadIdx.x * 0.001f + i;
#pragma unroll
  for (int i = 0; i < ILP; ++i) u[i] = __float_as_uint(x[i]);
  const uint32_t al = 0x3e4d3e4d;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < REPS; ++r) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (OP == 0) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(x[2 * i]), "f"(__uint_as_float(u[i])));
      if (OP == 1) asm volatile("mul.rn.bf16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(al));
      if (OP == 2) asm volatile("max.bf16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(al));
      if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[i]) : "f"(x[i + ILP]));
      if (OP == 4) asm volatile("{.reg .b64 a, b; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %2}; fma.rn.f32x2 a, a, b, b; mov.b64 {%0, %1}, a;}"
                                : "+f"(x[2 * i]), "+f"(x[2 * i + 1]) : "f"(seed));
      if (OP == 5) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(x[i + ILP]));
      if (OP == 6) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(x[i + ILP]));
      if (OP == 7) asm volatile("fma.rn.bf16x2 %0, %0, %1, %1;" : "+r"(u[i]) : "r"(al));
      if (OP == 8) asm volatile("{.reg .b64 a, b; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %2}; add.rn.f32x2 a, a, b; mov.b64 {%0, %1}, a;}"
                                : "+f"(x[2 * i]), "+f"(x[2 * i + 1]) : "f"(seed));
      if (OP == 9) asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(u[i]));
      if (OP == 10) asm volatile("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(x[2 * i]), "f"(__uint_as_float(u[i])));
      if (OP == 11) asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(u[i]) : "r"(al));
      if (OP == 12) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(u[i]) : "r"(al));
      if (OP == 13) asm volatile("{.reg .pred p; setp.gt.f32 p, %0, %1; selp.f32 %0, %0, %1, p;}" : "+f"(x[i]) : "f"(x[i + ILP]));   // FSETP + FSEL
      if (OP == 14) asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(u[i]) : "r"(al));                                              // IMAD
      if (OP == 15) asm volatile("set.gt.bf16x2.bf16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(al));                                        // HSET2
      if (OP == 16) asm volatile("add.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(al));                                                     // IADD3
      if (OP == 17) asm volatile("shf.l.wrap.b32 %0, %0, %1, 3;" : "+r"(u[i]) : "r"(al));                                           // SHF
      if (OP == 18) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(x[i + ILP]));                                          // FADD
      if (OP == 19) asm volatile("{.reg .pred p; setp.gt.s32 p, %0, %1; selp.b32 %0, %0, %1, p;}" : "+r"(u[i]) : "r"(al));          // ISETP + SEL
    }
  }
  long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 2 * ILP; ++i) acc += x[i];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc += __uint_as_float(u[i]);
  if (acc == 12345.678f) out[1] = 1;
  if (threadIdx.x == 0) out[0] = t1 - t0;
}

template <int OP>
void run(const char* name, int nthreads) {
  long long* d; cudaMalloc(&d, 16);
  bench<OP><<<1, nthreads>>>(d, 1.5f);
  cudaDeviceSynchronize();
  bench<OP><<<1, nthreads>>>(d, 1.5f);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double warps_per_smsp = nthreads / 32 / 4.0;
  printf("%-58s %3d thr: %6.2f cyc per warp-instruction per SMSP  %s\n", name, nthreads, (double)h / (REPS * ILP * warps_per_smsp),
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int nt : {128, 512}) {
    run<0>("cvt.rn.bf16x2.f32 (F2FP)", nt);
    run<10>("cvt.rn.relu.bf16x2.f32", nt);
    run<1>("mul.bf16x2 (HMUL2)", nt);
    run<7>("fma.bf16x2 (HFMA2)", nt);
    run<2>("max.bf16x2 (HMNMX2)", nt);
    run<3>("fma.f32 (FFMA)", nt);
    run<4>("fma.f32x2 (FFMA2)", nt);
    run<8>("add.f32x2 (FADD2)", nt);
    run<5>("max.f32 (FMNMX)", nt);
    run<6>("mul.f32 (FMUL)", nt);
    run<9>("shfl.bfly", nt);
    run<11>("lop3", nt);
    run<12>("prmt", nt);
    run<13>("setp.gt.f32 + selp.f32 (FSETP + FSEL, two instructions)", nt);
    run<14>("mad.lo.s32 (IMAD)", nt);
    run<15>("set.gt.bf16x2 (HSET2)", nt);
    run<16>("add.s32 (IADD3)", nt);
    run<17>("shf.l.wrap (SHF)", nt);
    run<18>("add.f32 (FADD)", nt);
    run<19>("setp + selp.b32 (ISETP + SEL, two instructions)", nt);
  }
  return 0;
}
