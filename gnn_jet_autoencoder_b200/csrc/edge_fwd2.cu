// Forward edge kernel, second generation (GJ_PREC_BF16, four edge layers with compile-time widths).
//
// Replaces the edge part of one iteration of reference models/graphnet.py:154-168 (_getA :186-223, _edge_conv :273-289,
// the sum over j of _concat :243):  e_i = sum_j EdgeNet([h_i | h_j | d_ij]).
//
// Decomposition.  A WARP owns a (jet, j block) pair -- lane = j, so Q_j (registers) and h_j (a per-lane shared-memory
// row) are loaded once -- and walks i = 0..N-1.  A TILE GROUP is four such warps (one TMEM lane quadrant each, possibly
// four different jets) moving in lock step: one tile = 128 edge rows = 4 (jet, i) x 32 j.  A CTA runs NWG tile groups,
// each with a private TMEM slot, so that the tensor pipe works on one group's GEMM while the others run epilogues.
// Per tile:
//   a0 = leaky(P_i + Q_j + wd d_ij)                    CUDA cores (packed fp32), bf16 pairs -> TMEM (tcgen05.st)
//   a_l = leaky(W_l a_{l-1} + b_l), l = 1, 2, 3         tcgen05.mma, A from TMEM, W_l (bf16, K-major) from shared memory,
//                                                      bias as one extra k-step against a constant (1,1,0..) A chunk;
//                                                      epilogue tcgen05.ld -> cvt bf16x2 -> HMUL2/HMNMX2 -> tcgen05.st
//                                                      in place over the accumulator columns just read
//   e_i = sum_j a_3                                    transpose-reduce over the warp's lanes, one 64-byte store
// Nothing but P|Q, h and e touches HBM; activations never touch shared memory.
//
// TMEM slot of a tile group (160 columns):   [0,128) acc1 -> a1 packed in [0,64) -> acc2 in [64,128) -> a2 in [0,32)
//                                            [128,144) a0 (packed)      [144,160) acc3
// column 480..487: the constant bias A chunk shared by all groups.
#include <stdlib.h>

#include "tc2_common.cuh"
#include "tma.cuh"

namespace {
using namespace tc2;

struct Fwd2Args {
  const float* h; const float* pq; const float* params; float* e_out; const uint8_t* wimg;
  float* d_out;          // optional (B, N, NJ32) pair distances, saved for the backward kernel
  int B, N, NJB, cols, ld, mink;
  int Hb, Hs;            // h_i row length (multiple of 4 floats), h_j per-lane row stride (floats, (Hs/4) odd)
  int pW1, pb1, pW2, pb2, pW3, pb3, pWd, K0;
  float alpha;
  int tiles_total;       // ceil(B * NJB / 4) * N
  int NJ32;              // row stride of d_out
};

constexpr int F2_IC = 8;           // i's per staged P_i | h_i chunk (double buffered, cp.async)

template <int E0, int E1, int E2, int E3, int NWG>
struct Fwd2Smem {
  static constexpr int o_bar = 0;                       // NWG * 3 mbarriers
  static constexpr int o_bar_w = 112;                   // parameter image (TMA) has landed
  static constexpr int o_slot = 128;
  static constexpr int o_wd = 256;                      // E0 floats
  static constexpr int o_img = 1024;                    // packed parameter image (WImage): bias chunks, W1, W2, W3
  static constexpr int o_b1 = o_img + WImage<E0, E1, E2, E3>::o_b1;
  static constexpr int o_b2 = o_img + WImage<E0, E1, E2, E3>::o_b2;
  static constexpr int o_b3 = o_img + WImage<E0, E1, E2, E3>::o_b3;
  static constexpr int o_w1 = o_img + WImage<E0, E1, E2, E3>::o_w1;
  static constexpr int o_w2 = o_img + WImage<E0, E1, E2, E3>::o_w2;
  static constexpr int o_w3 = o_img + WImage<E0, E1, E2, E3>::o_w3;
  // [128 rows][16] constant A operand of the bias k-step: chunk 0 = (1, 1, 0, ...) per row, chunk 1 = zeros; chunk 1 doubles
  // as the second k-chunk of every bias B operand
  static constexpr int o_ones = o_img + WImage<E0, E1, E2, E3>::bytes;
  static constexpr int o_a0 = o_ones + 4096;                               // NWG x [128 rows][E0] bf16, interleaved
  static constexpr int o_warp = o_a0 + NWG * 128 * E0 * 2;
  __host__ __device__ static int warp_bytes(int Hb, int Hs) { return (2 * F2_IC * (E0 + Hb) + 32 * Hs) * 4; }
  __host__ __device__ static int total(int Hb, int Hs) { return o_warp + NWG * 4 * warp_bytes(Hb, Hs); }
};

// transpose-reduce of 16 per-lane values over the 32 lanes of a warp: every lane returns the sum over all lanes of
// channel (lane >> 1) & 15
__device__ __forceinline__ float warp_transpose_sum16(const float (&v)[16], int lane) {
  float w8[8], w4[4], w2[2];
  bool up = lane & 16;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float send = up ? v[q] : v[q + 8], keep = up ? v[q + 8] : v[q];
    w8[q] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  up = lane & 8;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float send = up ? w8[q] : w8[q + 4], keep = up ? w8[q + 4] : w8[q];
    w4[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  up = lane & 4;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const float send = up ? w4[q] : w4[q + 2], keep = up ? w4[q + 2] : w4[q];
    w2[q] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  up = lane & 2;
  const float send = up ? w2[0] : w2[1], keep = up ? w2[1] : w2[0];
  const float w1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  return w1 + __shfl_xor_sync(0xffffffffu, w1, 1);
}

// epilogue of a hidden layer: NCOL fp32 accumulator columns at acc -> leaky -> NCOL/2 packed bf16 columns at dst
// (16-column chunks, two register sets: the next chunk's TMEM load flies while the current one is processed)
template <int NCOL>
__device__ __forceinline__ void epilogue_hidden(uint32_t acc, uint32_t dst, __nv_bfloat162 alpha2) {
  static_assert(NCOL % 32 == 0, "hidden widths are multiples of 32");
  constexpr int NCH = NCOL / 16;
  uint32_t va[16], vb[16];
  tmem_ld16_u(acc, va);
#pragma unroll
  for (int ch = 0; ch < NCH; ch += 2) {
    tmem_ld_wait(); tmem_pin16(va);
    tmem_ld16_u(acc + (uint32_t)((ch + 1) * 16), vb);
    {
      uint32_t o[8];
#pragma unroll
      for (int p = 0; p < 8; ++p) o[p] = leaky_pack(__uint_as_float(va[2 * p]), __uint_as_float(va[2 * p + 1]), alpha2);
      tmem_st8(dst + (uint32_t)(ch * 8), o);
    }
    tmem_ld_wait(); tmem_pin16(vb);
    if (ch + 2 < NCH) tmem_ld16_u(acc + (uint32_t)((ch + 2) * 16), va);
    {
      uint32_t o[8];
#pragma unroll
      for (int p = 0; p < 8; ++p) o[p] = leaky_pack(__uint_as_float(vb[2 * p]), __uint_as_float(vb[2 * p + 1]), alpha2);
      tmem_st8(dst + (uint32_t)((ch + 1) * 8), o);
    }
  }
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// optional stage timeline of tile group 0 of CTA 0 (GJ_TRACE=3): 8 clock stamps per tile
__device__ long long g_f2_trace[8 * 256];
#define F2_STAMP(s) do { if (TRACE && blockIdx.x == 0 && tid == 0 && tr_n < 256) g_f2_trace[tr_n * 8 + (s)] = clock64(); } while (0)

template <int E0, int E1, int E2, int E3, int NWG, bool TRACE>
__global__ void __launch_bounds__(NWG * 128, 1) edge_fwd2_kernel(const Fwd2Args A, const __grid_constant__ CUtensorMap tm_w) {
  gj_pdl_wait();      // programmatic dependent launch: the preceding kernels' results (P|Q, parameter image) are visible from here
  static_assert(E0 == 32 && E3 == 16, "tile map assumes a 32-wide first and a 16-wide last edge layer");
  static_assert(E1 % 32 == 0 && E1 <= 128 && E2 % 32 == 0 && E1 / 2 + E2 <= 128 && E2 / 2 + E3 <= E1 / 2, "TMEM slot map");
  static_assert(NWG * 128 <= 512, "one 128-column TMEM slot per tile group");
  using S = Fwd2Smem<E0, E1, E2, E3, NWG>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = (int)uni((uint32_t)(tid >> 5));
  const int wg = warp >> 2, wq = warp & 3;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::o_bar) + wg * 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::o_slot);
  float* s_wd = reinterpret_cast<float*>(smem + S::o_wd);

  // ---- one-time staging: weights / biases as bf16 B operands, wd, the constant bias A chunk, barriers, TMEM ----
  for (int c = tid; c < E0; c += NWG * 128) s_wd[c] = __ldg(A.params + A.pWd + c * A.K0);
  for (int idx = tid; idx < 1024; idx += NWG * 128)     // ones chunk: k = 0, 1 -> 1.0 (bias hi + lo), k = 2..15 -> 0
    reinterpret_cast<uint32_t*>(smem + S::o_ones)[idx] = (idx < 512 && (idx & 3) == 0) ? 0x3F803F80u : 0u;
  {   // the staging rows' padding columns stay zero for the whole kernel
    float* wz = reinterpret_cast<float*>(smem + S::o_warp);
    const int nz = NWG * 4 * S::warp_bytes(A.Hb, A.Hs) / 4;
    for (int idx = tid; idx < nz; idx += NWG * 128) wz[idx] = 0.f;
  }
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + S::o_bar_w);
  if (tid == 0) {
    for (int b = 0; b < NWG * 3; ++b) mbar_init(reinterpret_cast<uint64_t*>(smem + S::o_bar) + b, 1);
    mbar_init(bar_w, 1);
    fence_barrier_init();
    // the packed bf16 parameter image (30 KB, one box) -> shared memory by TMA
    tma::prefetch_map(&tm_w);
    tma::mbar_expect_tx(bar_w, WImage<E0, E1, E2, E3>::bytes);
    tma::load_2d(smem + S::o_img, &tm_w, 0, 0, bar_w);
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  mbar_wait(bar_w, 0u);
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t slot = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(wg * 128);   // this thread's row of the group's slot
  const uint32_t slot0 = tmem_base + (uint32_t)(wg * 128);                                // lane 0 (MMA operand addresses)
  const __nv_bfloat162 alpha2 = __float2bfloat162_rn(A.alpha);
  const float alpha = A.alpha;
  uint8_t* a0 = smem + S::o_a0 + wg * (128 * E0 * 2);
  uint8_t* a0_row = a0 + (wq * 32 + lane) * 16;

  // MMA constants (warp-uniform)
  const uint32_t idesc1 = make_idesc_bf16(128, E1, 0, 0), idesc2 = make_idesc_bf16(128, E2, 0, 0), idesc3 = make_idesc_bf16(128, E3, 0, 0);
  const uint64_t dW1 = wdesc_kmajor(smem_u32(smem + S::o_w1), E1), dW2 = wdesc_kmajor(smem_u32(smem + S::o_w2), E2),
                 dW3 = wdesc_kmajor(smem_u32(smem + S::o_w3), E3);
  const uint32_t zero_a = smem_u32(smem + S::o_ones + 2048);
  const uint64_t dB1 = make_smem_desc(smem_u32(smem + S::o_b1), zero_a - smem_u32(smem + S::o_b1), 128),
                 dB2 = make_smem_desc(smem_u32(smem + S::o_b2), zero_a - smem_u32(smem + S::o_b2), 128),
                 dB3 = make_smem_desc(smem_u32(smem + S::o_b3), zero_a - smem_u32(smem + S::o_b3), 128);
  const uint64_t dA0 = make_smem_desc(smem_u32(a0), 2048, 128), dOnes = make_smem_desc(smem_u32(smem + S::o_ones), 2048, 128);

  // per-warp staging
  const int Hb = A.Hb, Hs = A.Hs, RS = E0 + Hb;
  float* s_pi = reinterpret_cast<float*>(smem + S::o_warp + warp * S::warp_bytes(Hb, Hs));      // [2][F2_IC][E0 + Hb]
  float* s_hj = s_pi + 2 * F2_IC * RS;                                                          // [32][Hs]
  const int H4 = Hb >> 2;
  const bool h_vec = (A.ld & 3) == 0 && (A.cols & 3) == 0 && ((reinterpret_cast<uintptr_t>(A.h) & 15) == 0);

  // tile range of this group: tiles are (group task k, i), k-major
  const int ngroups = gridDim.x * NWG, gidx = blockIdx.x * NWG + wg;
  const long long T = A.tiles_total;
  const int g0 = (int)(T * gidx / ngroups), g1 = (int)(T * (gidx + 1) / ngroups);
  const int N = A.N;
  const int ntasks = A.B * A.NJB;
  // state of the tile being PREPARED (first layer), one tile ahead of the tile in the tensor-core stages
  int k = g0 / N, i = g0 - k * N;
  bool fresh = true;
  float q[E0];             // Q_j rounded to bf16 (the backward kernel recomputes the first layer from the same rounded values), kept
                           // as fp32: the forward kernel has the registers, and unpacking pairs cost two instructions per pair and tile
  bool p_active = false, p_valid = false;
  uint32_t p_vmask = 0u;      // lanes (j) of the warp's j block that are real particles
  size_t node0 = 0;
  float* e_dst = nullptr;
  int e_dst_jb = 0;

  // issues the cp.async copies of chunk c (i in [c * F2_IC, ...)) of the current jet into buffer c & 1
  auto stage_chunk = [&](int c) {
    const int ib = c * F2_IC, n = min(F2_IC, N - ib);
    float* dst = s_pi + (c & 1) * F2_IC * RS;
    for (int idx = lane; idx < n * (E0 / 4); idx += 32) {
      const int r = idx / (E0 / 4), c4 = idx - r * (E0 / 4);
      cp_async16(dst + r * RS + 4 * c4, A.pq + (node0 + ib + r) * (2 * E0) + 4 * c4);
    }
    if (h_vec) {
      const int hq = A.cols >> 2;
      for (int idx = lane; idx < n * hq; idx += 32) {
        const int r = idx / hq, c4 = idx - r * hq;
        cp_async16(dst + r * RS + E0 + 4 * c4, A.h + (node0 + ib + r) * A.ld + 4 * c4);
      }
    } else {
      for (int idx = lane; idx < n * A.cols; idx += 32) {
        const int r = idx / A.cols, kk = idx - r * A.cols;
        cp_async4(dst + r * RS + E0 + kk, A.h + (node0 + ib + r) * A.ld + kk);
      }
    }
    cp_async_commit();
  };

  // first edge layer of the tile (k, i) -> a0 (shared memory, bf16 A operand); advances (k, i)
  auto prepare = [&](bool& t_active, bool& t_valid, float*& t_erow, uint32_t& t_vmask) {
    if (fresh) {
      // ---- new (jet, j block): Q_j -> registers, h_j -> per-lane shared row ----
      const int task = 4 * k + wq;
      p_active = task < ntasks;
      const int tk = p_active ? task : 0;
      const int jet = tk / A.NJB, jb = tk - jet * A.NJB;
      const int j = jb * 32 + lane;
      p_valid = p_active && j < N;
      p_vmask = __ballot_sync(0xffffffffu, p_valid);
      node0 = (size_t)jet * N;
      e_dst = A.e_out + ((size_t)jb * A.B + jet) * N * E3;
      e_dst_jb = jb;
      cp_async_wait<0>();
      __syncwarp();
      stage_chunk(i / F2_IC);
      if ((i / F2_IC + 1) * F2_IC < N) stage_chunk(i / F2_IC + 1);
      if (j < N) {
        const float4* src = reinterpret_cast<const float4*>(A.pq + (node0 + j) * (2 * E0) + E0);
#pragma unroll
        for (int c = 0; c < E0 / 4; ++c) {
          const float4 v = __ldg(src + c);
          q[4 * c] = __bfloat162float(__float2bfloat16_rn(v.x)); q[4 * c + 1] = __bfloat162float(__float2bfloat16_rn(v.y));
          q[4 * c + 2] = __bfloat162float(__float2bfloat16_rn(v.z)); q[4 * c + 3] = __bfloat162float(__float2bfloat16_rn(v.w));
        }
        const float* hsrc = A.h + (node0 + j) * A.ld;
        for (int kk = 0; kk < A.cols; ++kk) s_hj[lane * Hs + kk] = __ldg(hsrc + kk);
      } else {
#pragma unroll
        for (int c = 0; c < E0; ++c) q[c] = 0.f;
        for (int kk = 0; kk < A.cols; ++kk) s_hj[lane * Hs + kk] = 0.f;
      }
      if ((i / F2_IC + 1) * F2_IC < N) cp_async_wait<1>(); else cp_async_wait<0>();
      __syncwarp();
    } else if ((i & (F2_IC - 1)) == 0) {
      // ---- entering the next chunk (prefetched F2_IC tiles ago); prefetch the one after it into the buffer just left ----
      cp_async_wait<0>();
      __syncwarp();
      if ((i / F2_IC + 1) * F2_IC < N) stage_chunk(i / F2_IC + 1);
    }
    fresh = false;
    {
      const float* Pi = s_pi + (((i / F2_IC) & 1) * F2_IC + (i & (F2_IC - 1))) * RS;
      const float4* hi4 = reinterpret_cast<const float4*>(Pi + E0);
      const float4* hj4 = reinterpret_cast<const float4*>(s_hj + lane * Hs);
      float d;
      if (A.mink) {       // width 4 only: x0^2 - x1^2 - x2^2 - x3^2 (graphnet.py:320-323)
        const float4 a = hi4[0], b = hj4[0];
        const float x0 = b.x - a.x, x1 = b.y - a.y, x2 = b.z - a.z, x3 = b.w - a.w;
        d = x0 * x0 - x1 * x1 - x2 * x2 - x3 * x3;
      } else {
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll 2
        for (int k4 = 0; k4 < H4; ++k4) {
          const float4 a = hi4[k4], b = hj4[k4];
          const float2 x = make_float2(b.x - a.x, b.y - a.y), y = make_float2(b.z - a.z, b.w - a.w);
          acc = fma2(x, x, acc); acc = fma2(y, y, acc);
        }
        d = acc.x + acc.y;
      }
      if (A.d_out && p_active) A.d_out[(node0 + i) * A.NJ32 + (e_dst_jb * 32 + lane)] = p_valid ? d : 0.f;
      const float2 d2 = make_float2(d, d);
#pragma unroll
      for (int c = 0; c < E0; c += 8) {
        uint32_t o[4];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int cc = c + 4 * hh;
          const float4 p = *reinterpret_cast<const float4*>(Pi + cc);
          const float4 w = *reinterpret_cast<const float4*>(s_wd + cc);
          const float2 z0 = fma2(make_float2(w.x, w.y), d2, add2(make_float2(p.x, p.y), make_float2(q[cc], q[cc + 1])));
          const float2 z1 = fma2(make_float2(w.z, w.w), d2, add2(make_float2(p.z, p.w), make_float2(q[cc + 2], q[cc + 3])));
          o[2 * hh] = leaky_pack(z0.x, z0.y, alpha2);
          o[2 * hh + 1] = leaky_pack(z1.x, z1.y, alpha2);
        }
        *reinterpret_cast<uint4*>(a0_row + (c >> 3) * 2048) = make_uint4(o[0], o[1], o[2], o[3]);
      }
      fence_proxy_async();
    }
    t_active = p_active; t_valid = p_valid; t_erow = e_dst + (size_t)i * E3; t_vmask = p_vmask;
    if (++i == N) { i = 0; ++k; fresh = true; }
  };

  uint32_t ph = 0;
  int tr_n = 0;
  bool n_active = false, n_valid = false;
  float* n_erow = nullptr;
  uint32_t n_vmask = 0u;
  if (g0 < g1) prepare(n_active, n_valid, n_erow, n_vmask);

  for (int g = g0; g < g1; ++g) {
    F2_STAMP(0);
    const bool active = n_active, valid = n_valid;
    float* const erow = n_erow;
    const uint32_t valid_mask = n_vmask;
    // ---- layer 1: acc1[0,128) = a0 (shared memory) W1^T + b1 ----
    tc_fence_before();
    named_bar_sync(1 + wg, 128);
    if (wq == 0) {
      tc_fence_after();
#pragma unroll
      for (int s = 0; s < E0 / 16; ++s)
        mma_bf16_ss_elect(slot0, dA0 + (uint64_t)(s * (4096 >> 4)), dW1 + (uint64_t)(s * ((2 * E1 * 16) >> 4)), idesc1, s > 0);
      mma_bf16_ss_elect(slot0, dOnes, dB1, idesc1, 1u);
      mma_commit_elect(bars + 0);
    }
    F2_STAMP(1);
    mbar_wait_all(bars + 0, ph);
    tc_fence_after();
    F2_STAMP(2);
    epilogue_hidden<E1>(slot, slot, alpha2);
    tmem_st_wait();
    tc_fence_before();
    F2_STAMP(3);
    named_bar_sync(1 + wg, 128);
    if (wq == 1) {
      tc_fence_after();
#pragma unroll
      for (int s = 0; s < E1 / 16; ++s) mma_ts_elect(slot0 + E1 / 2, slot0 + 8 * s, dW2 + (uint64_t)(s * ((2 * E2 * 16) >> 4)), idesc2, s > 0);
      mma_bf16_ss_elect(slot0 + E1 / 2, dOnes, dB2, idesc2, 1u);
      mma_commit_elect(bars + 1);
    }
    // ---- first layer of the NEXT tile while layer 2 runs on the tensor pipe (a0 is free: layer 1 has completed) ----
    if (g + 1 < g1) prepare(n_active, n_valid, n_erow, n_vmask);
    F2_STAMP(4);
    mbar_wait_all(bars + 1, ph);
    tc_fence_after();
    F2_STAMP(5);
    epilogue_hidden<E2>(slot + E1 / 2, slot, alpha2);
    tmem_st_wait();
    tc_fence_before();
    named_bar_sync(1 + wg, 128);
    if (wq == 2) {
      tc_fence_after();
#pragma unroll
      for (int s = 0; s < E2 / 16; ++s) mma_ts_elect(slot0 + E2 / 2, slot0 + 8 * s, dW3 + (uint64_t)(s * ((2 * E3 * 16) >> 4)), idesc3, s > 0);
      mma_bf16_ss_elect(slot0 + E2 / 2, dOnes, dB3, idesc3, 1u);
      mma_commit_elect(bars + 2);
    }
    F2_STAMP(6);
    mbar_wait_all(bars + 2, ph);
    tc_fence_after();
    F2_STAMP(7);
    ++tr_n;
    {
      // e_i = sum over the quadrant's 32 lanes (j) of leaky(acc3), padded rows masked.  The accumulator is read as matrix
      // fragments (tmem_ld_frag16): this thread holds lanes g, g + 8, g + 16, g + 24 (g = lane / 4) of columns 2q, 2q + 1, 8 + 2q,
      // 9 + 2q (q = lane % 4), so the sum is 12 local adds and three shuffle rounds over g.
      uint32_t fa[8], fb[8];
      tmem_ld_frag16(slot + E2 / 2, fa);
      tmem_ld_frag16(slot + E2 / 2 + (16u << 16), fb);
      const int g4 = lane >> 2;
      const uint32_t vm = valid_mask >> g4;      // bit 0, 8, 16, 24: rows g, g + 8, g + 16, g + 24 are real particles
      const float m0 = (vm & 1u) ? 1.f : 0.f, m1 = (vm & 0x100u) ? 1.f : 0.f, m2 = (vm & 0x10000u) ? 1.f : 0.f, m3 = (vm & 0x1000000u) ? 1.f : 0.f;
      tmem_ld_wait(); tmem_pin8(fa); tmem_pin8(fb);
      float s[4];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float z0 = __uint_as_float(fa[4 * k + c]), z1 = __uint_as_float(fa[4 * k + 2 + c]);
          const float z2 = __uint_as_float(fb[4 * k + c]), z3 = __uint_as_float(fb[4 * k + 2 + c]);
          s[2 * k + c] = fmaf(fmaxf(z0, alpha * z0), m0, fmaxf(z1, alpha * z1) * m1) + fmaf(fmaxf(z2, alpha * z2), m2, fmaxf(z3, alpha * z3) * m3);
        }
      }
#pragma unroll
      for (int off = 4; off < 32; off <<= 1) {
#pragma unroll
        for (int c = 0; c < 4; ++c) s[c] += __shfl_xor_sync(0xffffffffu, s[c], off);
      }
      if (active && lane < 4) {
        *reinterpret_cast<float2*>(erow + 2 * lane) = make_float2(s[0], s[1]);
        *reinterpret_cast<float2*>(erow + 8 + 2 * lane) = make_float2(s[2], s[3]);
      }
    }
    ph ^= 1u;
  }

  gj_pdl_trigger();      // the next kernel of the stream may be launched while the remaining groups finish
  cp_async_wait<0>();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// sums the NJB per-j-block partial aggregates in a fixed order
// e_i = sum over the j blocks' partial aggregates (N > 32): one 16-byte group per thread (n is a multiple of 4: E_last = 16)
__global__ void __launch_bounds__(256) sum_jblocks_kernel(const float4* __restrict__ part, int njb, size_t n4, float4* __restrict__ out) {
  gj_pdl_sync();
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= n4) return;
  float4 s = __ldg(part + idx);
  for (int b = 1; b < njb; ++b) {
    const float4 v = __ldg(part + b * n4 + idx);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  out[idx] = s;
}

}  // namespace

int gj_num_sms();
void gj_set_error(const char* fmt, ...);

// debugging aid (not part of the ABI header): stage timeline of the last traced forward launch
extern "C" int gj_debug_read_fwd2_trace(long long* out) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out, g_f2_trace, sizeof(long long) * 8 * 256) == cudaSuccess ? 0 : 1;
}

// widths covered by the specialised kernels
bool gj_fwd2_supported(const MPLayout& L) {
  return L.Le == 4 && L.E[0] == 32 && L.E[1] == 128 && L.E[2] == 64 && L.E[3] == 16 && L.alpha <= 1.f && L.cols <= 64;
}
// shared-memory bytes / TMEM columns of the forward kernel for this step (gj_mp_plan_info)
void gj_fwd2_plan(const MPLayout& L, int* smem_bytes, int* tmem_cols) {
  const int Hb = (L.cols + 3) & ~3, Hs = 4 * ((Hb >> 2) | 1);
  *smem_bytes = Fwd2Smem<32, 128, 64, 16, 4>::total(Hb, Hs);
  *tmem_cols = 512;
}
size_t gj_fwd2_ws_floats(const MPLayout& L) {
  const int njb = (L.N + 31) / 32;
  return njb > 1 ? (size_t)njb * L.B * L.N * L.E[3] : 0;
}

// wimg: 16-byte aligned buffer of gj_wimage_floats() floats that receives the packed parameter image; d_save: optional
// (B, N, NJ32) buffer that receives the pair distances (both are reused by gj_edge_bwd2 when the caller keeps them)
size_t gj_wimage_floats() { return (WImage<32, 128, 64, 16>::bytes + 255) / 256 * 64; }

// the packed parameter images of n steps (all of them gj_fwd2_supported) in one launch
int gj_pack_batch2(int n, const MPLayout* Ls, const float* const* params, float* const* wimg, cudaStream_t stream) {
  for (int s0 = 0; s0 < n; s0 += PACK_BATCH_MAX) {
    const int m = n - s0 < PACK_BATCH_MAX ? n - s0 : PACK_BATCH_MAX;
    PackBatch Bt;
    for (int s = 0; s < m; ++s) {
      const MPLayout& L = Ls[s0 + s];
      Bt.params[s] = params[s0 + s];
      Bt.P[s] = WImageSrc{L.pW[1], L.pb[1], L.pW[2], L.pb[2], L.pW[3], L.pb[3]};
      Bt.img[s] = reinterpret_cast<uint8_t*>(wimg[s0 + s]);
    }
    for (int s = m; s < PACK_BATCH_MAX; ++s) { Bt.params[s] = nullptr; Bt.P[s] = WImageSrc{0, 0, 0, 0, 0, 0}; Bt.img[s] = nullptr; }
    gj_launch(pack_edge_weights_batch_kernel<32, 128, 64, 16>, dim3(4, m), 256, 0, stream, Bt);
    const cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) { gj_set_error("pack_edge_weights_batch launch: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  }
  return GJ_OK;
}

// skip_pack: the packed parameter image is already in wimg (gj_mp_steps_pack)
int gj_edge_fwd2(const MPLayout& L, const float* h, const float* pq, const float* params, float* e_out, float* ws, float* wimg,
                 float* d_save, cudaStream_t stream, bool kernel_only, bool skip_pack) {
  constexpr int NWG = 4;
  Fwd2Args A;
  A.wimg = reinterpret_cast<const uint8_t*>(wimg);
  A.d_out = d_save;
  if (!kernel_only && !skip_pack) {
    WImageSrc P{L.pW[1], L.pb[1], L.pW[2], L.pb[2], L.pW[3], L.pb[3]};
    gj_launch(pack_edge_weights_kernel<32, 128, 64, 16>, 4, 256, 0, stream, params, P, reinterpret_cast<uint8_t*>(wimg));
  }
  A.h = h; A.pq = pq; A.params = params;
  A.B = L.B; A.N = L.N; A.NJB = (L.N + 31) / 32; A.NJ32 = A.NJB * 32; A.cols = L.cols; A.ld = L.ld; A.mink = L.mink;
  A.e_out = A.NJB > 1 ? ws : e_out;
  A.Hb = (L.cols + 3) & ~3;
  A.Hs = 4 * ((A.Hb >> 2) | 1);
  A.pW1 = L.pW[1]; A.pb1 = L.pb[1]; A.pW2 = L.pW[2]; A.pb2 = L.pb[2]; A.pW3 = L.pW[3]; A.pb3 = L.pb[3];
  A.pWd = L.pW[0] + 2 * L.H; A.K0 = L.K[0];
  A.alpha = L.alpha;
  const long long tasks4 = ((long long)L.B * A.NJB + 3) / 4;
  if (tasks4 * L.N > 0x7fffffffLL) { gj_set_error("gj_mp_step_fwd(bf16): batch * nodes too large"); return GJ_ERR_INVALID; }
  A.tiles_total = (int)(tasks4 * L.N);
  using S = Fwd2Smem<32, 128, 64, 16, NWG>;
  const int smem = S::total(A.Hb, A.Hs);
  if (smem > 227 * 1024) { gj_set_error("gj_mp_step_fwd(bf16): needs %d B shared memory", smem); return GJ_ERR_SMEM; }
  static const int trace_env = getenv("GJ_TRACE") ? atoi(getenv("GJ_TRACE")) : 0;
  auto kern = trace_env == 3 ? edge_fwd2_kernel<32, 128, 64, 16, NWG, true> : edge_fwd2_kernel<32, 128, 64, 16, NWG, false>;
  cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (ce != cudaSuccess) { gj_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  int grid = gj_num_sms();
  const int max_grid = (A.tiles_total + NWG - 1) / NWG;
  if (grid > max_grid) grid = max_grid;
  CUtensorMap tm_w;      // the parameter image as rows of 1 KB
  static_assert(WImage<32, 128, 64, 16>::bytes % 1024 == 0 && WImage<32, 128, 64, 16>::bytes / 1024 <= 256, "one TMA box");
  if (int rc = gj_tmap_2d(&tm_w, wimg, 256, WImage<32, 128, 64, 16>::bytes / 1024, 1024, 256, WImage<32, 128, 64, 16>::bytes / 1024)) return rc;
  gj_launch_edge(kern, grid, NWG * 128, smem, stream, A, tm_w);
  ce = cudaGetLastError();
  if (ce != cudaSuccess) { gj_set_error("edge_fwd2 launch: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  if (kernel_only) return GJ_OK;
  if (A.NJB > 1) {
    const size_t n4 = (size_t)L.B * L.N * (L.E[3] / 4);
    gj_launch(sum_jblocks_kernel, (unsigned)((n4 + 255) / 256), 256, 0, stream, reinterpret_cast<const float4*>(ws), A.NJB, n4,
              reinterpret_cast<float4*>(e_out));
    ce = cudaGetLastError();
    if (ce != cudaSuccess) { gj_set_error("sum_jblocks launch: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  }
  return GJ_OK;
}
