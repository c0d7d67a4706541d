// bf16 (GJ_PREC_BF16) edge kernels on tcgen05 tensor cores with TMEM accumulators.
//
// Work decomposition: a CTA holds NWG independent warpgroups (128 threads = 128 TMEM lanes each).  Every
// warpgroup walks whole jets on its own: for a jet the N x N pair set is visited as (32 i's) x (32 j's) blocks,
// each block as tiles of 128 edge rows (4 i's x 32 j's, j on the lanes).  Per tile:
//   a0 = leaky(P_i + Q_j + wd d_ij)            CUDA cores, fp32 -> bf16 A operand in shared memory
//   a_l = leaky(W_l a_{l-1} + b_l), l >= 1     tcgen05.mma (M = 128 rows, N = layer width, K = previous width)
//                                              issued by one elected thread, accumulator in the warpgroup's TMEM
//                                              columns, epilogue tcgen05.ld -> bias + LeakyReLU -> bf16 -> smem
//   e_i = sum_j a_last                         warp shuffles (fp32)
// While one warpgroup runs an epilogue, the other's MMAs occupy the tensor pipe.  Edge activations never leave
// the SM.  Replaces the edge part of reference models/graphnet.py:154-168 and its autograd adjoint.
#include "gj_common.cuh"
#include "umma.cuh"

namespace {

using namespace umma;

struct Carver {
  int off = 0;
  int take(int n) { int o = off; off += (n + 3) & ~3; return o; }
};

int next_pow2_cols(int c) { int p = 32; while (p < c) p <<= 1; return p; }

struct TCPlan {
  int nwg;
  int o_bar, o_tmem_slot;
  int o_wT[GJ_MAX_LAYERS];  // bf16 interleaved edge weights, layers >= 1 (bytes from smem base)
  int o_shared_f32;         // float region shared by all warpgroups (biases, wd)
  int wg_base, wg_stride;   // per-warpgroup region
  int w_act[GJ_MAX_LAYERS]; // bf16 activation buffers inside the warpgroup region (bytes)
  int w_f32;                // float region inside the warpgroup region
  int tmem_cols_per_wg, tmem_cols_total;
  int smem_bytes;
};

// Fills the float offsets of L relative to the two float regions and the byte plan T.
void plan_tc_fwd(MPLayout* L, TCPlan* T, int nwg) {
  L->R = 128; L->Rs = 0;
  T->nwg = nwg;
  Carver c;
  for (int l = 1; l < L->Le; ++l) L->o_bE[l] = c.take(L->Ep[l]);
  L->o_wd = c.take(L->E0p);
  Carver w;
  L->o_h = w.take(GJ_IB * L->Hs);
  L->o_hj = w.take(32 * L->Hs);
  L->o_P = w.take(GJ_IB * L->E0s);
  L->o_Q = w.take(32 * L->E0s);
  L->o_e = w.take(GJ_IB * L->ELs);
  int off = 0;
  T->o_bar = off; off += 64;
  T->o_tmem_slot = off; off += 64;
  for (int l = 1; l < L->Le; ++l) { T->o_wT[l] = off; off += L->Ep[l] * L->Kp[l] * 2; }
  T->o_shared_f32 = off; off += c.off * 4;
  off = gj_round_up(off, 128);
  T->wg_base = off;
  int wo = 0;
  for (int l = 0; l + 1 < L->Le; ++l) { T->w_act[l] = wo; wo += L->Ep[l] * 128 * 2; }
  T->w_f32 = wo; wo += w.off * 4;
  wo = gj_round_up(wo, 128);
  T->wg_stride = wo;
  T->smem_bytes = T->wg_base + nwg * wo;
  int mx = 32;
  for (int l = 1; l < L->Le; ++l) if (L->Ep[l] > mx) mx = L->Ep[l];
  T->tmem_cols_per_wg = gj_round_up(mx, 32);
  T->tmem_cols_total = next_pow2_cols(T->tmem_cols_per_wg * nwg);
}

__device__ void stage_small(const MPLayout& L, const float* __restrict__ params, float* smf, int tid, int nthr) {
  for (int l = 1; l < L.Le; ++l)
    for (int c = tid; c < L.Ep[l]; c += nthr) smf[L.o_bE[l] + c] = c < L.E[l] ? __ldg(params + L.pb[l] + c) : 0.f;
  for (int c = tid; c < L.E0p; c += nthr) smf[L.o_wd + c] = c < L.E[0] ? __ldg(params + L.pW[0] + c * L.K[0] + 2 * L.H) : 0.f;
}

// bf16 interleaved staging of edge weights l >= 1 as the K-major B operand (N = out feature rows, K = in).
__device__ void stage_edge_weights_bf16(const MPLayout& L, const TCPlan& T, const float* __restrict__ params, uint8_t* smem,
                                        int tid, int nthr) {
  for (int l = 1; l < L.Le; ++l) {
    uint8_t* w = smem + T.o_wT[l];
    const int Ep = L.Ep[l], Kp = L.Kp[l], E = L.E[l], K = L.K[l];
    for (int idx = tid; idx < Ep * Kp; idx += nthr) {
      int n = idx / Kp, k = idx - n * Kp;
      float v = (n < E && k < K) ? __ldg(params + L.pW[l] + n * K + k) : 0.f;
      *reinterpret_cast<__nv_bfloat16*>(w + il_off(n, k, Ep)) = __float2bfloat16_rn(v);
    }
  }
}

// rows [r0, r0 + 32) of h (zero padded) and of the P or Q half of PQ, by one warpgroup
__device__ void wg_load_block(const MPLayout& L, const float* __restrict__ hjet, const float* __restrict__ pqjet, int half,
                              int r0, float* sh, float* spq, int t) {
  for (int idx = t; idx < 32 * L.H; idx += 128) {
    int n = idx / L.H, k = idx - n * L.H;
    sh[n * L.Hs + k] = (r0 + n < L.N && k < L.cols) ? __ldg(hjet + (size_t)(r0 + n) * L.ld + k) : 0.f;
  }
  for (int idx = t; idx < 32 * L.E0p; idx += 128) {
    int n = idx / L.E0p, c = idx - n * L.E0p;
    spq[n * L.E0s + c] = (r0 + n < L.N) ? __ldg(pqjet + (size_t)(r0 + n) * 2 * L.E0p + half * L.E0p + c) : 0.f;
  }
}

// D[128 x N](tmem) = A[128 x K](smem, K-major, 128 rows) * W[N x K]^T (smem, K-major, N rows)
__device__ __forceinline__ void issue_layer_mma(uint32_t d_tmem, uint32_t a_saddr, uint32_t w_saddr, int N, int K) {
  const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
  for (int ks = 0; ks < K / 16; ++ks) {
    uint64_t ad = make_smem_desc(a_saddr + ks * (2 * 128 * 16), 128 * 16, 128);
    uint64_t bd = make_smem_desc(w_saddr + ks * (2 * N * 16), N * 16, 128);
    mma_bf16_ss(d_tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
  }
}

// first edge layer of one tile on CUDA cores: thread t owns row (i = it*4 + wq, j = lane); returns d_ij
__device__ __forceinline__ float tc_layer0(const MPLayout& L, const float* sm_h, const float* sm_hj, const float* sm_P,
                                           const float* sm_Q, const float* wd, uint8_t* a0, int il, int lane, int t) {
  const float* hi = sm_h + il * L.Hs;
  const float* hj = sm_hj + lane * L.Hs;
  float d = 0.f;
  for (int k = 0; k < L.H; ++k) {
    float x = hj[k] - hi[k] + GJ_EPS;
    float s = (L.mink && k > 0) ? -1.f : 1.f;
    d = fmaf(s * x, x, d);
  }
  const float* P = sm_P + il * L.E0s;
  const float* Q = sm_Q + lane * L.E0s;
  for (int c0 = 0; c0 < L.E0p; c0 += 8) {
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = gj_leaky(P[c0 + q] + Q[c0 + q] + wd[c0 + q] * d, L.alpha);
    uint4 pk = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    *reinterpret_cast<uint4*>(a0 + (c0 >> 3) * 2048 + t * 16) = pk;
  }
  return d;
}

template <int NWG>
__global__ void __launch_bounds__(NWG * 128, 1)
edge_fwd_tc_kernel(const MPLayout L, const TCPlan T, const float* __restrict__ h, const float* __restrict__ pq,
                   const float* __restrict__ params, float* __restrict__ e_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int NT = NWG * 128;
  const int tid = threadIdx.x, wg = tid >> 7, t = tid & 127, wq = t >> 5, lane = t & 31;
  float* smf = reinterpret_cast<float*>(smem + T.o_shared_f32);
  uint8_t* wgb = smem + T.wg_base + wg * T.wg_stride;
  float* wgf = reinterpret_cast<float*>(wgb + T.w_f32);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + T.o_bar) + wg;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + T.o_tmem_slot);

  stage_small(L, params, smf, tid, NT);
  stage_edge_weights_bf16(L, T, params, smem, tid, NT);
  if (tid == 0) {
    for (int w = 0; w < NWG; ++w) mbar_init(reinterpret_cast<uint64_t*>(smem + T.o_bar) + w, 1);
    fence_barrier_init();
  }
  if (tid < 32) tmem_alloc(tmem_slot, (uint32_t)T.tmem_cols_total);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_d = tmem_base + (uint32_t)(wg * T.tmem_cols_per_wg);
  const uint32_t tmem_row = tmem_d + ((uint32_t)(wq * 32) << 16);
  uint32_t phase = 0;
  const int bar_id = 1 + wg;
  float* sm_h = wgf + L.o_h;
  float* sm_hj = wgf + L.o_hj;
  float* sm_Q = wgf + L.o_Q;
  float* sm_P = wgf + L.o_P;
  float* sm_e = wgf + L.o_e;
  const float* wd = smf + L.o_wd;

  for (int jet = blockIdx.x * NWG + wg; jet < L.B; jet += gridDim.x * NWG) {
    const float* hjet = h + (size_t)jet * L.N * L.ld;
    const float* pqjet = pq + (size_t)jet * L.N * 2 * L.E0p;
    for (int i0 = 0; i0 < L.N; i0 += GJ_IB) {
      const int ni = min(GJ_IB, L.N - i0);
      named_bar_sync(bar_id, 128);
      wg_load_block(L, hjet, pqjet, 0, i0, sm_h, sm_P, t);
      for (int idx = t; idx < GJ_IB * L.ELs; idx += 128) sm_e[idx] = 0.f;
      for (int j0 = 0; j0 < L.N; j0 += 32) {
        const int nj = min(32, L.N - j0);
        named_bar_sync(bar_id, 128);
        wg_load_block(L, hjet, pqjet, 1, j0, sm_hj, sm_Q, t);
        named_bar_sync(bar_id, 128);
        const int nit = (ni + 3) / 4;
        for (int it = 0; it < nit; ++it) {
          const int il = it * 4 + wq;
          const bool valid = il < ni && lane < nj;
          tc_layer0(L, sm_h, sm_hj, sm_P, sm_Q, wd, wgb + T.w_act[0], il, lane, t);
          fence_proxy_async();
          named_bar_sync(bar_id, 128);
          for (int l = 1; l < L.Le; ++l) {
            if (t == 0) {
              tc_fence_after();
              issue_layer_mma(tmem_d, smem_u32(wgb + T.w_act[l - 1]), smem_u32(smem + T.o_wT[l]), L.Ep[l], L.Kp[l]);
              mma_commit(bar);
            }
            __syncwarp();
            mbar_wait(bar, phase);
            phase ^= 1u;
            tc_fence_after();
            const bool last = (l == L.Le - 1);
            const float* bias = smf + L.o_bE[l];
            for (int c0 = 0; c0 < L.Ep[l]; c0 += 16) {
              float v[16];
              tmem_ld16(tmem_row + (uint32_t)c0, v);
#pragma unroll
              for (int q = 0; q < 16; ++q) v[q] = gj_leaky(v[q] + bias[c0 + q], L.alpha);
              if (!last) {
                uint8_t* al = wgb + T.w_act[l];
                uint4 p0 = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                uint4 p1 = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
                *reinterpret_cast<uint4*>(al + (c0 >> 3) * 2048 + t * 16) = p0;
                *reinterpret_cast<uint4*>(al + ((c0 >> 3) + 1) * 2048 + t * 16) = p1;
              } else {
                // e_i += sum_j a_last (padded rows masked AFTER the activation: leaky(b) != 0)
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                  float s = gj_warp_sum(valid ? v[q] : 0.f);
                  if (lane == 0) sm_e[il * L.ELs + c0 + q] += s;
                }
              }
            }
            if (!last) fence_proxy_async();
            tc_fence_before();
            named_bar_sync(bar_id, 128);
          }
        }
      }
      named_bar_sync(bar_id, 128);
      for (int idx = t; idx < ni * L.EL; idx += 128) {
        int n = idx / L.EL, c = idx - n * L.EL;
        e_out[((size_t)jet * L.N + i0 + n) * L.EL + c] = sm_e[n * L.ELs + c];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem_base, (uint32_t)T.tmem_cols_total);
}

// ------------------------------------------------------------------------------------------------
// tcgen05 self-test: one CTA, D = A * B^T with selectable operand major-ness; dumps raw TMEM.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(int M, int N, int K, int a_mn, int b_mn, const float* __restrict__ A, const float* __restrict__ B,
                     float* __restrict__ out, int tmem_cols, int b_off) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 16);
  uint8_t* sa = smem + 1024;
  uint8_t* sb = smem + b_off;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < M * K; idx += 128) {
    int m = idx / K, k = idx - m * K;
    uint32_t off = a_mn ? il_off(k, m, K) : il_off(m, k, M);
    *reinterpret_cast<__nv_bfloat16*>(sa + off) = __float2bfloat16_rn(A[idx]);
  }
  for (int idx = tid; idx < N * K; idx += 128) {
    int n = idx / K, k = idx - n * K;
    uint32_t off = b_mn ? il_off(k, n, K) : il_off(n, k, N);
    *reinterpret_cast<__nv_bfloat16*>(sb + off) = __float2bfloat16_rn(B[idx]);
  }
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (tid < 32) tmem_alloc(slot, (uint32_t)tmem_cols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(M, N, a_mn, b_mn);
    for (int ks = 0; ks < K / 16; ++ks) {
      uint64_t ad = a_mn ? make_smem_desc(smem_u32(sa) + ks * 256, 128, K * 16)
                         : make_smem_desc(smem_u32(sa) + ks * (2 * M * 16), M * 16, 128);
      uint64_t bd = b_mn ? make_smem_desc(smem_u32(sb) + ks * 256, 128, K * 16)
                         : make_smem_desc(smem_u32(sb) + ks * (2 * N * 16), N * 16, 128);
      mma_bf16_ss(tbase, ad, bd, idesc, ks > 0 ? 1u : 0u);
    }
    mma_commit(bar);
  }
  __syncwarp();
  mbar_wait(bar, 0);
  tc_fence_after();
  const int wq = tid >> 5, lane = tid & 31;
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tbase + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0, v);
    for (int q = 0; q < 16 && c0 + q < N; ++q) out[(size_t)(wq * 32 + lane) * N + c0 + q] = v[q];
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tbase, (uint32_t)tmem_cols);
}

}  // namespace

int gj_num_sms();
void gj_set_error(const char* fmt, ...);
int gj_edge_grid(int);
int gj_edge_bwd_simt(MPLayout, const float*, const float*, const float*, const float*, float*, float*, float*, float*,
                     cudaStream_t);

int gj_edge_fwd_tc(MPLayout L, const float* h, const float* pq, const float* params, float* e_out, cudaStream_t stream) {
  for (int l = 1; l < L.Le; ++l)
    if (L.Ep[l] > 256 || L.Kp[l] > 256) { gj_set_error("gj_mp_step_fwd(bf16): edge widths above 256 unsupported"); return GJ_ERR_INVALID; }
  TCPlan T;
  constexpr int NWG = 2;
  plan_tc_fwd(&L, &T, NWG);
  if (T.smem_bytes > 227 * 1024 || T.tmem_cols_total > 512) {
    gj_set_error("gj_mp_step_fwd(bf16): needs %d B shared memory / %d TMEM columns (limits 232448 / 512)", T.smem_bytes, T.tmem_cols_total);
    return GJ_ERR_SMEM;
  }
  auto kern = edge_fwd_tc_kernel<NWG>;
  cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T.smem_bytes);
  if (ce != cudaSuccess) { gj_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  int sms = gj_num_sms();
  int grid = (L.B + NWG - 1) / NWG; if (grid > sms) grid = sms;
  kern<<<grid, NWG * 128, T.smem_bytes, stream>>>(L, T, h, pq, params, e_out);
  ce = cudaGetLastError();
  if (ce != cudaSuccess) { gj_set_error("edge_fwd_tc launch: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  return GJ_OK;
}

size_t gj_edge_bwd_tc_ws_floats(const MPLayout& L) { return (size_t)gj_edge_grid(L.B) * L.pV[0]; }

int gj_edge_bwd_tc(MPLayout L, const float* h, const float* pq, const float* params, const float* de, float* dpq, float* dh,
                   float* dparams, float* part, cudaStream_t stream) {
  // interim: the tensor-core backward is being written; the bf16 mode's gradient runs the fp32 SIMT kernel.
  return gj_edge_bwd_simt(L, h, pq, params, de, dpq, dh, dparams, part, stream);
}

int gj_umma_selftest_launch(int M, int N, int K, int a_mn, int b_mn, const float* a, const float* b, float* out,
                            cudaStream_t stream) {
  if ((M != 64 && M != 128) || N < 8 || N > 256 || (N % 8) || K < 16 || K > 256 || (K % 16)) {
    gj_set_error("gj_umma_selftest: unsupported shape m=%d n=%d k=%d", M, N, K);
    return GJ_ERR_INVALID;
  }
  int a_bytes = gj_round_up(M * K * 2, 1024), b_bytes = gj_round_up(N * K * 2, 1024);
  int b_off = 1024 + a_bytes;
  int smem = b_off + b_bytes;
  int cols = 32; while (cols < N) cols <<= 1;
  cudaError_t ce = cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (ce != cudaSuccess) { gj_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  umma_selftest_kernel<<<1, 128, smem, stream>>>(M, N, K, a_mn, b_mn, a, b, out, cols, b_off);
  ce = cudaGetLastError();
  if (ce != cudaSuccess) { gj_set_error("umma_selftest launch: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  return GJ_OK;
}
