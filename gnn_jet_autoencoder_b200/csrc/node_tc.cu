// Node-level adjoint kernels on tcgen05 (GJ_PREC_BF16 only; the fp32 mode keeps the SIMT kernels of node_kernels.cu).
//
// node_pre_bwd_tc_kernel: adjoint of the first-edge-layer projections P_i = Wa h_i + b0, Q_j = Wb h_j
// (the factorised W0 [h_i | h_j | d] of reference models/graphnet.py:220) over all B*N node rows:
//     dh[row]   += dPQ[row] (64) . [Wa ; Wb] (64 x H)                      one GEMM per 128-row tile, accumulator in TMEM
//     [dWa ; dWb | db0] += dPQ^T [h | 1]    (64 x (H + 16), K = rows)      accumulator resident in TMEM for the whole kernel
// A CTA is one 128-thread tile group (thread = node row = TMEM lane) with a 128-column TMEM allocation, so four CTAs share
// an SM and hide each other's staging / GEMM / epilogue phases.  The row tile is staged as bf16 in the UMMA SWIZZLE_NONE
// slab layout ([8-channel slab][128 rows][8]), which serves as K-major A operand of the dgrad GEMM and as MN-major operand
// (K = tile row) of the weight-gradient GEMM alike.  bf16 operands, fp32 accumulation: the same precision class as the
// edge-MLP weight gradients.
// All global traffic is coalesced (a thread walking its own row touches 32 lines per instruction, which made the first
// version L1-wavefront bound): the dPQ rows are read four whole rows per warp instruction straight into slab entries (and
// prefetched one tile ahead, in registers, behind the GEMMs); the h and dh row tiles are copied with cp.async into padded
// fp32 tiles, converted / accumulated thread-per-row there, and the dh tile is written back with coalesced stores.
#include "tc2_common.cuh"

namespace {
using namespace tc2;

__device__ __forceinline__ void ntc_cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void ntc_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ntc_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int HP>
struct PreBwdSmem {
  static constexpr int TS = HP + 4;                          // fp32 tile row stride (4 * odd floats: per-row 16-byte accesses are conflict free)
  static constexpr int o_bar = 0;
  static constexpr int o_slot = 16;
  static constexpr int o_wt = 1024;                          // [HP][64] bf16 K-major B operand: WabT[n][c] = [Wa ; Wb][c][n]
  static constexpr int o_g = o_wt + HP * 64 * 2;             // dPQ tile: 8 slabs
  static constexpr int o_h = o_g + 8 * 2048;                 // h tile: HP / 8 slabs, then the ones slab and a zero slab
  static constexpr int o_th = o_h + (HP / 8 + 2) * 2048;     // fp32 h tile [128][TS] (coalesced copy of the global rows)
  static constexpr int o_td = o_th + 128 * TS * 4;           // fp32 dh tile [128][TS]: read, accumulated, written back
  static constexpr int total = o_td + 128 * TS * 4;
};

// dPQ rows of one tile as this thread's 8 (row, 8-channel slab) entries: entry it covers row (it * 4 + warp) * 4 + (lane & 3),
// slab lane >> 2, so that a warp instruction reads four whole 256-byte rows
struct DpqRegs { float4 v[16]; };
__device__ __forceinline__ void ntc_load_dpq(DpqRegs& R, const float* __restrict__ dpq, int row0, int rows, int warp, int lane) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int row = row0 + (it * 4 + warp) * 4 + (lane & 3);
    const float4* src = reinterpret_cast<const float4*>(dpq + (size_t)row * 64 + (lane >> 2) * 8);
    if (row < rows) { R.v[2 * it] = __ldg(src); R.v[2 * it + 1] = __ldg(src + 1); }
    else { R.v[2 * it] = make_float4(0.f, 0.f, 0.f, 0.f); R.v[2 * it + 1] = R.v[2 * it]; }
  }
}

template <int HP>
__global__ void __launch_bounds__(128) node_pre_bwd_tc_kernel(int rows, int H, int cols, int ld, int K0, const float* __restrict__ h,
                                                              const float* __restrict__ w0, const float* __restrict__ dpq,
                                                              float* __restrict__ dh, float* __restrict__ part) {
  using S = PreBwdSmem<HP>;
  constexpr int TS = S::TS;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = (int)uni((uint32_t)(tid >> 5));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + S::o_bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::o_slot);
  float* sH = reinterpret_cast<float*>(smem + S::o_th);
  float* sD = reinterpret_cast<float*>(smem + S::o_td);
  const int ntiles = (rows + 127) / 128;
  DpqRegs R;
  if ((int)blockIdx.x < ntiles) ntc_load_dpq(R, dpq, blockIdx.x * 128, rows, warp, lane);
  // ---- one-time staging: transposed projection weights, constant slabs, barrier, TMEM ----
  for (int idx = tid; idx < HP * 64; idx += 128) {
    const int n = idx >> 6, c = idx & 63;
    const float v = n < cols ? __ldg(w0 + (c & 31) * K0 + (c < 32 ? n : H + n)) : 0.f;
    *reinterpret_cast<__nv_bfloat16*>(smem + S::o_wt + (c >> 3) * (HP * 16) + n * 16 + (c & 7) * 2) = __float2bfloat16_rn(v);
  }
  for (int idx = tid; idx < 1024; idx += 128)      // ones slab (channel HP = 1.0, HP + 1 .. HP + 7 = 0), zero slab
    reinterpret_cast<uint32_t*>(smem + S::o_h + (HP / 8) * 2048)[idx] = (idx < 512 && (idx & 3) == 0) ? 0x00003F80u : 0u;
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
  {      // the gradient accumulator [64, 64 + HP + 16) starts at zero: every weight-gradient MMA accumulates
    uint32_t z[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) z[c] = 0u;
#pragma unroll
    for (int c0 = 0; c0 < HP + 16; c0 += 16) tmem_st16(lane_base + 64 + c0, z);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const uint32_t ga = smem_u32(smem + S::o_g), ha = smem_u32(smem + S::o_h), wta = smem_u32(smem + S::o_wt);
  const bool vec = (ld & 3) == 0 && ((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(dh)) & 15) == 0;
  const int q4 = min((ld + 3) >> 2, HP / 4);      // 16-byte pieces fetched per row (they cover the `cols` live columns)
  uint32_t ph = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int row0 = tile * 128, nrows = min(128, rows - row0);
    // ---- coalesced copies of the h and dh row tiles into shared memory ----
    if (vec) {
      for (int idx = tid; idx < nrows * q4; idx += 128) {
        const int r = idx / q4, k = idx - r * q4;
        ntc_cp_async16(sH + r * TS + 4 * k, h + (size_t)(row0 + r) * ld + 4 * k);
        ntc_cp_async16(sD + r * TS + 4 * k, dh + (size_t)(row0 + r) * ld + 4 * k);
      }
    } else {
      for (int idx = tid; idx < nrows * cols; idx += 128) {
        const int r = idx / cols, k = idx - r * cols;
        sH[r * TS + k] = __ldg(h + (size_t)(row0 + r) * ld + k);
        sD[r * TS + k] = dh[(size_t)(row0 + r) * ld + k];
      }
    }
    ntc_cp_async_commit();
    // ---- this tile's dPQ entries (prefetched during the previous tile) -> bf16 slabs ----
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const float4 a = R.v[2 * it], b = R.v[2 * it + 1];
      const int r = (it * 4 + warp) * 4 + (lane & 3);
      *reinterpret_cast<uint4*>(smem + S::o_g + (lane >> 2) * 2048 + r * 16) =
          make_uint4(bf2_as_u32(__floats2bfloat162_rn(a.x, a.y)), bf2_as_u32(__floats2bfloat162_rn(a.z, a.w)),
                     bf2_as_u32(__floats2bfloat162_rn(b.x, b.y)), bf2_as_u32(__floats2bfloat162_rn(b.z, b.w)));
    }
    ntc_cp_async_wait_all();
    __syncthreads();
    // ---- this thread's h row -> bf16 slabs (zero beyond `cols` and beyond the last row) ----
    {
      const bool live = tid < nrows;
#pragma unroll
      for (int s = 0; s < HP / 8; ++s) {
        const float4 a = *reinterpret_cast<const float4*>(sH + tid * TS + 8 * s), b = *reinterpret_cast<const float4*>(sH + tid * TS + 8 * s + 4);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        float w[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) w[q] = (live && 8 * s + q < cols) ? v[q] : 0.f;
        *reinterpret_cast<uint4*>(smem + S::o_h + s * 2048 + tid * 16) =
            make_uint4(bf2_as_u32(__floats2bfloat162_rn(w[0], w[1])), bf2_as_u32(__floats2bfloat162_rn(w[2], w[3])),
                       bf2_as_u32(__floats2bfloat162_rn(w[4], w[5])), bf2_as_u32(__floats2bfloat162_rn(w[6], w[7])));
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      // dgrad: acc[0, HP) = dPQ (A, K-major) . [Wa ; Wb] (B = WabT, K-major), K = 64
      const uint32_t i_d = make_idesc_bf16(128, HP, 0, 0);
#pragma unroll
      for (int s = 0; s < 4; ++s)
        mma_bf16_ss_elect(tmem_base, make_smem_desc(ga + s * 4096, 2048, 128), make_smem_desc(wta + s * 2 * HP * 16, HP * 16, 128), i_d, s > 0);
      // wgrad: acc[64, 64 + HP + 16) += dPQ^T [h | 1 | 0] (both MN-major views, K = tile row)
      const uint32_t i_g = make_idesc_bf16(64, HP + 16, 1, 1);
#pragma unroll
      for (int s = 0; s < 8; ++s)
        mma_bf16_ss_elect(tmem_base + 64, make_smem_desc(ga + s * 256, 128, 2048), make_smem_desc(ha + s * 256, 128, 2048), i_g, 1u);
      mma_commit_elect(bar);
    }
    if (tile + (int)gridDim.x < ntiles) ntc_load_dpq(R, dpq, (tile + gridDim.x) * 128, rows, warp, lane);      // in flight behind the GEMMs
    mbar_wait(bar, ph); ph ^= 1u;
    tc_fence_after();
    // ---- epilogue: dh tile row += acc, then the tile goes back to global memory with coalesced stores ----
#pragma unroll
    for (int c0 = 0; c0 < HP; c0 += 16) {
      uint32_t v[16];
      tmem_ld16_u(lane_base + c0, v);
      tmem_ld_wait(); tmem_pin16(v);
#pragma unroll
      for (int q = 0; q < 16; q += 4) {
        float4 o = *reinterpret_cast<const float4*>(sD + tid * TS + c0 + q);
        o.x += __uint_as_float(v[q]); o.y += __uint_as_float(v[q + 1]); o.z += __uint_as_float(v[q + 2]); o.w += __uint_as_float(v[q + 3]);
        *reinterpret_cast<float4*>(sD + tid * TS + c0 + q) = o;
      }
    }
    tc_fence_before();
    __syncthreads();      // the staging buffers and acc[0, HP) are free again; the dh tile is complete
    if (vec) {
      for (int idx = tid; idx < nrows * q4; idx += 128) {
        const int r = idx / q4, k = idx - r * q4;
        float* dst = dh + (size_t)(row0 + r) * ld + 4 * k;
        const float4 o = *reinterpret_cast<const float4*>(sD + r * TS + 4 * k);
        if (4 * k + 3 < cols) *reinterpret_cast<float4*>(dst) = o;
        else {
          const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) if (4 * k + q < cols) dst[q] = ov[q];
        }
      }
    } else {
      for (int idx = tid; idx < nrows * cols; idx += 128) {
        const int r = idx / cols, k = idx - r * cols;
        dh[(size_t)(row0 + r) * ld + k] = sD[r * TS + k];
      }
    }
    __syncthreads();      // the next tile's copies overwrite the dh tile
  }
  // ---- gradient read-out: M = 64 accumulator, channel m at lane (m / 16) * 32 + m % 16; column k < HP = in feature, HP = bias ----
  tc_fence_after();
  float* out = part + (size_t)blockIdx.x * (64 * H + 32);
  {
    const int m = warp * 16 + (lane & 15);      // 0..31: Wa / b0 rows, 32..63: Wb rows (lanes 16..31 hold nothing)
    float* rowp = out + (m < 32 ? m * H : 32 * H + (m - 32) * H);
#pragma unroll
    for (int c0 = 0; c0 < HP + 16; c0 += 16) {
      uint32_t v[16];
      tmem_ld16_u(lane_base + 64 + c0, v);      // .aligned: every lane of the warp issues the load
      tmem_ld_wait(); tmem_pin16(v);
      if (lane < 16) {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int k = c0 + q;
          if (k < H && k < HP) rowp[k] = k < cols ? __uint_as_float(v[q]) : 0.f;
          if (k == HP && m < 32) out[64 * H + m] = __uint_as_float(v[q]);
        }
      }
    }
    if (lane < 16) for (int k = HP; k < H; ++k) rowp[k] = 0.f;      // columns of W0 that multiply the zero padding of h
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

}  // namespace

int gj_num_sms();
void gj_set_error(const char* fmt, ...);

bool gj_node_pre_bwd_tc_supported(const MPLayout& L) { return L.E[0] == 32 && L.E0p == 32 && L.cols <= 32; }
static int pre_bwd_tc_grid(const MPLayout& L) {      // every CTA gets the same number of tiles (+-1) with 3-4 CTAs per SM resident
  const int tiles = (L.B * L.N + 127) / 128, cap = (L.cols <= 16 ? 4 : 3) * gj_num_sms();
  if (tiles <= cap) return tiles > 0 ? tiles : 1;
  const int rounds = (tiles + cap - 1) / cap;
  return (tiles + rounds - 1) / rounds;
}
// per-CTA partials [Wa grads 32*H][Wb grads 32*H][b0 grads 32], reduced by reduce_pre_partials_kernel
size_t gj_node_pre_bwd_tc_ws_floats(const MPLayout& L) { return (size_t)pre_bwd_tc_grid(L) * (64 * L.H + 32); }

// launches the kernel; *nparts receives the number of per-CTA partials written to `part`
int gj_node_pre_bwd_tc(const MPLayout& L, const float* h, const float* params, const float* dpq, float* dh, float* part, int* nparts,
                       cudaStream_t st) {
  const int grid = pre_bwd_tc_grid(L), rows = L.B * L.N;
  *nparts = grid;
  cudaError_t ce;
  if (L.cols <= 16) {
    using S = PreBwdSmem<16>;
    ce = cudaFuncSetAttribute(node_pre_bwd_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::total);
    if (ce == cudaSuccess) node_pre_bwd_tc_kernel<16><<<grid, 128, S::total, st>>>(rows, L.H, L.cols, L.ld, L.K[0], h, params + L.pW[0], dpq, dh, part);
  } else {
    using S = PreBwdSmem<32>;
    ce = cudaFuncSetAttribute(node_pre_bwd_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::total);
    if (ce == cudaSuccess) node_pre_bwd_tc_kernel<32><<<grid, 128, S::total, st>>>(rows, L.H, L.cols, L.ld, L.K[0], h, params + L.pW[0], dpq, dh, part);
  }
  if (ce == cudaSuccess) ce = cudaGetLastError();
  if (ce != cudaSuccess) { gj_set_error("node_pre_bwd_tc launch: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  return GJ_OK;
}
