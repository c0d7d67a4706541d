// Building blocks of the second-generation tensor-core edge kernels (edge_fwd2.cu / edge_bwd2.cu).
//
// What differs from the first-generation kernels in edge_tc.cu:
//   * widths are template parameters, so every MMA batch, descriptor advance and epilogue loop is unrolled;
//   * the A operand of the forward and dgrad GEMMs lives in TENSOR MEMORY (tcgen05.mma "TS" form): an epilogue writes
//     the next layer's bf16 operand straight back into TMEM with tcgen05.st (thread = row = TMEM lane), in place over
//     the accumulator columns it has just read, so forward activations never touch shared memory;
//   * a warp owns a (jet, 32-wide j block) for all i: Q_j lives in registers, P_i / h_i are staged by the warp itself,
//     and the four warps of a tile group (one TMEM lane quadrant each) may belong to four different jets;
//   * the LeakyReLU runs on packed bf16 pairs (cvt.rn.bf16x2.f32 + HMUL2 + HMNMX2: 1.5 instructions per element).
#pragma once
#include "gj_common.cuh"
#include "umma.cuh"

namespace tc2 {
using namespace umma;

// ---- TMEM <-> registers -------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld16_u(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8_u(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_pin8(uint32_t (&r)[8]) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]));
}
__device__ __forceinline__ void tmem_ld32_u(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
// Matrix-fragment load: 16 lanes x 16 columns starting at (lane, column) of taddr.  Thread t receives, for g = t / 4 and q = t % 4,
//   r[0], r[1] = (lane g,     columns 2q, 2q + 1)      r[2], r[3] = (lane g + 8, columns 2q, 2q + 1)
//   r[4], r[5] = (lane g,     columns 8 + 2q, 9 + 2q)  r[6], r[7] = (lane g + 8, columns 8 + 2q, 9 + 2q)
// (measured with tools/tmem_frag_probe.cu).  Every thread holds elements of FOUR lanes after two such loads (lane offsets 0 and 16),
// which turns a sum over the 32 lanes of a quadrant into local adds plus three shuffle rounds instead of a 31-shuffle transpose.
__device__ __forceinline__ void tmem_ld_frag16(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
// None of the TMEM load / store wrappers carries a "memory" clobber: they do not touch C++-visible memory, and being
// volatile they keep their order relative to each other and to the fences / barrier waits (which do clobber memory);
// ordinary shared-memory loads and stores may be scheduled across them.
// Completes ALL outstanding tcgen05.ld of the thread.  The register arrays of those loads must be passed through
// tmem_pin*() right after, which orders every later use behind the wait for the compiler (no instruction is emitted).
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;"); }
__device__ __forceinline__ void tmem_pin16(uint32_t (&r)[16]) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                    "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
}
__device__ __forceinline__ void tmem_pin32(uint32_t (&r)[32]) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                    "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
  asm volatile("" : "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                    "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
         "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- MMA issue (warp-convergent: all 32 lanes execute with warp-uniform operands, one elected lane issues) ----
// D[tmem] (+)= A[tmem] * B[smem]^T : A is 128 lanes x 16 bf16 (8 columns, two K elements per 32-bit column)
__device__ __forceinline__ void mma_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Same instructions with the 64-bit shared-memory descriptors passed as (low, high) words: the high word (strides, version)
// is constant across the k-steps of a GEMM and the low word (address) advances by a small constant, which lets ptxas keep
// each descriptor in one uniform-register pair and advance it with a single 32-bit add per MMA.
__device__ __forceinline__ void mma_ss_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_ts_lh(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t desc_lo(uint64_t d) { return (uint32_t)d; }
__device__ __forceinline__ uint32_t desc_hi(uint64_t d) { return (uint32_t)(d >> 32); }

// ---- packed arithmetic ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bf2_as_u32(__nv_bfloat162 v) { return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ __nv_bfloat162 u32_as_bf2(uint32_t v) { return *reinterpret_cast<__nv_bfloat162*>(&v); }
// bf16 pair -> two floats (low half first)
__device__ __forceinline__ float2 unpack_bf2(uint32_t v) { return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u)); }
// {lo, hi} -> bf16x2 (lo in bits 0..15 = the lower K index), LeakyReLU(x) = max(x, alpha x) for 0 <= alpha <= 1
__device__ __forceinline__ uint32_t leaky_pack(float lo, float hi, __nv_bfloat162 alpha2) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  v = __hmax2(v, __hmul2(v, alpha2));
  return bf2_as_u32(v);
}
// (a.x + b.x, a.y + b.y) and (a * b + c) on the packed fp32 pipe
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rc, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rc;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}

// ---- weights: bf16 K-major B operand (N = out feature rows, K = in features), interleaved SWIZZLE_NONE layout ----
// element (n, k) at (k / 8) * (NOUT * 16) + n * 16 + (k % 8) * 2
template <int NOUT, int KIN>
__device__ __forceinline__ void stage_weight_kmajor(uint8_t* dst, const float* __restrict__ W, int tid, int nthr) {
  for (int idx = tid; idx < NOUT * KIN; idx += nthr) {
    const int n = idx / KIN, k = idx - n * KIN;
    *reinterpret_cast<__nv_bfloat16*>(dst + (k >> 3) * (NOUT * 16) + n * 16 + (k & 7) * 2) = __float2bfloat16_rn(__ldg(W + idx));
  }
}
// bias as one extra k-step of the same GEMM: B rows [NOUT][16], column 0 = bias (split hi + lo over columns 0 and 1 so
// the bias enters the fp32 accumulator to ~16 bits), the matching A chunk is the constant (1, 1, 0, ..., 0)
template <int NOUT>
__device__ __forceinline__ void stage_bias_slab(uint8_t* dst, const float* __restrict__ b, int tid, int nthr) {
  for (int idx = tid; idx < NOUT * 16; idx += nthr) {
    const int n = idx >> 4, k = idx & 15;
    const float v = __ldg(b + n);
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    const __nv_bfloat16 z = __float2bfloat16_rn(0.f);
    *reinterpret_cast<__nv_bfloat16*>(dst + (k >> 3) * (NOUT * 16) + n * 16 + (k & 7) * 2) = k == 0 ? hi : (k == 1 ? lo : z);
  }
}
// ---- packed image of one step's edge-MLP parameters as the kernels hold it in shared memory ----
// [ b1 | b2 | b3 : bias k-step B operands, chunk 0 only: per out feature (b_hi, b_lo, 0, 0, 0, 0, 0, 0) bf16 ]
// [ W1 | W2 | W3 : bf16 K-major B operands (N = out feature), interleaved SWIZZLE_NONE layout ]
// Built once per launch by pack_edge_weights_kernel (one small CTA) and copied by every CTA of the edge kernels with
// 16-byte loads -- converting 13 K weights per CTA with scattered 2-byte shared stores cost ~10 % of the forward kernel.
template <int E0, int E1, int E2, int E3>
struct WImage {
  static constexpr int o_b1 = 0;
  static constexpr int o_b2 = o_b1 + E1 * 16;
  static constexpr int o_b3 = o_b2 + E2 * 16;
  static constexpr int o_w1 = ((o_b3 + E3 * 16 + 1023) / 1024) * 1024;
  static constexpr int o_w2 = o_w1 + E1 * E0 * 2;
  static constexpr int o_w3 = o_w2 + E2 * E1 * 2;
  static constexpr int bytes = o_w3 + E3 * E2 * 2;
};
struct WImageSrc { int pW1, pb1, pW2, pb2, pW3, pb3; };

template <int NOUT>
__device__ __forceinline__ void pack_bias_chunk(uint8_t* dst, const float* __restrict__ b, int tid, int nthr) {
  for (int idx = tid; idx < NOUT * 4; idx += nthr) {      // four 32-bit words per out feature
    const int n = idx >> 2, w = idx & 3;
    const float v = __ldg(b + n);
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    reinterpret_cast<uint32_t*>(dst)[idx] = w == 0 ? ((uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16)) : 0u;
  }
}
template <int NOUT, int KIN>
__device__ __forceinline__ void pack_weight_kmajor(uint8_t* dst, const float* __restrict__ W, int tid, int nthr) {
  // one 16-byte core-matrix row (8 consecutive k of one out feature) per item: coalesced 32-byte reads, 16-byte writes
  for (int idx = tid; idx < NOUT * (KIN / 8); idx += nthr) {
    const int n = idx / (KIN / 8), kc = idx - n * (KIN / 8);
    const float4 a = __ldg(reinterpret_cast<const float4*>(W + n * KIN + kc * 8));
    const float4 c = __ldg(reinterpret_cast<const float4*>(W + n * KIN + kc * 8) + 1);
    uint4 o;
    o.x = bf2_as_u32(__floats2bfloat162_rn(a.x, a.y)); o.y = bf2_as_u32(__floats2bfloat162_rn(a.z, a.w));
    o.z = bf2_as_u32(__floats2bfloat162_rn(c.x, c.y)); o.w = bf2_as_u32(__floats2bfloat162_rn(c.z, c.w));
    *reinterpret_cast<uint4*>(dst + kc * (NOUT * 16) + n * 16) = o;
  }
}
template <int NOUT, int KIN>
__device__ __forceinline__ void pack_weight_kmajor_unaligned(uint8_t* dst, const float* __restrict__ W, int tid, int nthr) {
  for (int idx = tid; idx < NOUT * KIN; idx += nthr) {
    const int n = idx / KIN, k = idx - n * KIN;
    *reinterpret_cast<__nv_bfloat16*>(dst + (k >> 3) * (NOUT * 16) + n * 16 + (k & 7) * 2) = __float2bfloat16_rn(__ldg(W + idx));
  }
}
// fp32 parameters of the dense edge layers -> the packed bf16 image; tid / nthr: this thread's index and the thread count of the
// group of CTAs that packs the image
template <int E0, int E1, int E2, int E3>
__device__ __forceinline__ void pack_edge_weights_body(const float* __restrict__ params, const WImageSrc& P, uint8_t* __restrict__ img,
                                                       int tid, int nthr) {
  using I = WImage<E0, E1, E2, E3>;
  for (int idx = tid; idx < (I::o_w1 - (I::o_b3 + E3 * 16)) / 4; idx += nthr) reinterpret_cast<uint32_t*>(img + I::o_b3 + E3 * 16)[idx] = 0u;
  pack_bias_chunk<E1>(img + I::o_b1, params + P.pb1, tid, nthr);
  pack_bias_chunk<E2>(img + I::o_b2, params + P.pb2, tid, nthr);
  pack_bias_chunk<E3>(img + I::o_b3, params + P.pb3, tid, nthr);
  // the packed parameter block is only 4-byte aligned in general (odd-sized tensors in front of it)
  const bool al = ((reinterpret_cast<uintptr_t>(params) | (uintptr_t)(P.pW1 * 4) | (uintptr_t)(P.pW2 * 4) | (uintptr_t)(P.pW3 * 4)) & 15) == 0;
  if (al) {
    pack_weight_kmajor<E1, E0>(img + I::o_w1, params + P.pW1, tid, nthr);
    pack_weight_kmajor<E2, E1>(img + I::o_w2, params + P.pW2, tid, nthr);
    pack_weight_kmajor<E3, E2>(img + I::o_w3, params + P.pW3, tid, nthr);
  } else {
    pack_weight_kmajor_unaligned<E1, E0>(img + I::o_w1, params + P.pW1, tid, nthr);
    pack_weight_kmajor_unaligned<E2, E1>(img + I::o_w2, params + P.pW2, tid, nthr);
    pack_weight_kmajor_unaligned<E3, E2>(img + I::o_w3, params + P.pW3, tid, nthr);
  }
}
template <int E0, int E1, int E2, int E3>
__global__ void __launch_bounds__(256) pack_edge_weights_kernel(const float* __restrict__ params, WImageSrc P, uint8_t* __restrict__ img) {
  gj_pdl_sync();
  pack_edge_weights_body<E0, E1, E2, E3>(params, P, img, blockIdx.x * 256 + threadIdx.x, gridDim.x * 256);
}
// the images of up to PACK_BATCH_MAX steps in ONE launch (gj_mp_steps_pack): blockIdx.y = step
constexpr int PACK_BATCH_MAX = 16;
struct PackBatch { const float* params[PACK_BATCH_MAX]; WImageSrc P[PACK_BATCH_MAX]; uint8_t* img[PACK_BATCH_MAX]; };
template <int E0, int E1, int E2, int E3>
__global__ void __launch_bounds__(256) pack_edge_weights_batch_kernel(const __grid_constant__ PackBatch Bt) {
  gj_pdl_sync();
  const int s = blockIdx.y;
  pack_edge_weights_body<E0, E1, E2, E3>(Bt.params[s], Bt.P[s], Bt.img[s], blockIdx.x * 256 + threadIdx.x, gridDim.x * 256);
}
// every CTA: image (global, 16-byte aligned) -> shared memory
template <int BYTES>
__device__ __forceinline__ void load_wimage(uint8_t* dst, const uint8_t* __restrict__ img, int tid, int nthr) {
  static_assert(BYTES % 16 == 0, "image is copied in 16-byte pieces");
  for (int idx = tid; idx < BYTES / 16; idx += nthr) reinterpret_cast<uint4*>(dst)[idx] = __ldg(reinterpret_cast<const uint4*>(img) + idx);
}

__device__ __forceinline__ uint64_t wdesc_kmajor(uint32_t saddr, int nout) { return make_smem_desc(saddr, (uint32_t)nout * 16u, 128u); }

// all threads of a tile group wait on the barrier (hardware-suspended try_wait)
__device__ __forceinline__ void mbar_wait_all(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); }

}  // namespace tc2
