// fp32 (GJ_PREC_FP32) edge kernels: SIMT FFMA, forward and backward.
//
// One CTA walks whole jets.  For a jet the N x N pair set is visited as (32 i's) x (32 j's) blocks, each block as
// tiles of R = TI*32 edge rows (row = (i, j), j on the lanes), and lives only in shared memory:
//   a0_ij = leaky(P_i + Q_j + wd d_ij)           P, Q from node_pre_fwd (node_kernels.cu), d_ij from h
//   a_l   = leaky(W_l a_{l-1} + b_l), l >= 1     register-tiled fp32 GEMMs on the tile
//   e_i   = sum_j a_last                          (forward output; graphnet.py:243)
// The backward kernel recomputes the tile, then runs dgrad/wgrad per layer; nothing N^2-sized reaches HBM.
// Replaces the edge part of reference models/graphnet.py:154-168 (_getA :186-223, _edge_conv :273-289, the sum of
// _concat :243) and its autograd adjoint.
#include "gj_common.cuh"

namespace {

struct Carver {
  int off = 0;
  int take(int n) { int o = off; off += (n + 3) & ~3; return o; }
};

// dpar_global: the backward kernel's parameter-gradient accumulators (one float per edge parameter) do not fit next to the
// staged weights for wide layers; the CTA then accumulates straight into its partial in global memory (L2 resident).
// w_global (backward of wide layers whose weights are already laid out [Ep][Kp] in the parameter block, i.e. E and K need no
// padding): the GEMM helpers read the weights straight from global memory (L1 / L2) instead of a staged copy.
int plan_smem(MPLayout* L, int R, bool backward, bool dpar_global = false, bool w_global = false) {
  L->R = R; L->Rs = R + 4;
  Carver c;
  for (int l = 1; l < L->Le; ++l) L->o_wE[l] = w_global ? -1 : c.take(L->Ep[l] * L->Kp[l]);
  for (int l = 1; l < L->Le; ++l) L->o_bE[l] = c.take(L->Ep[l]);
  L->o_wd = c.take(L->E0p);
  L->o_h = c.take(GJ_IB * L->Hs);      // h rows of the i block
  L->o_hj = c.take(32 * L->Hs);        // h rows of the j block
  L->o_P = c.take(GJ_IB * L->E0s);
  L->o_Q = c.take(32 * L->E0s);
  L->o_e = c.take(GJ_IB * L->ELs);     // forward: e accumulators; backward: de rows
  if (!backward) {
    int w0 = 0, w1 = 0;
    for (int l = 0; l < L->Le; ++l) { if (l & 1) { if (L->Ep[l] > w1) w1 = L->Ep[l]; } else { if (L->Ep[l] > w0) w0 = L->Ep[l]; } }
    int a0 = c.take(w0 * L->Rs), a1 = c.take((w1 ? w1 : 1) * L->Rs);
    for (int l = 0; l < L->Le; ++l) L->o_act[l] = (l & 1) ? a1 : a0;
  } else {
    for (int l = 0; l < L->Le; ++l) L->o_act[l] = c.take(L->Ep[l] * L->Rs);
    L->n_edge_dpar = L->pV[0];           // all edge parameters (only l >= 1 weights/biases and the wd column are used)
    L->o_dpar = dpar_global ? -1 : c.take(L->n_edge_dpar);
    L->o_dh = c.take(GJ_IB * L->Hs);     // i-side distance gradient accumulators
    L->o_dQ = c.take(32 * L->E0s);
    L->o_dP = c.take(GJ_IB * L->E0s);
    L->Gs = 33;
    L->o_G = c.take(GJ_IB * L->Gs);
    L->o_drow = c.take(R);
  }
  L->smem_floats = c.off;
  return c.off * 4;
}

__device__ void stage_weights(const MPLayout& L, const float* __restrict__ params, float* sm) {
  for (int l = 1; l < L.Le; ++l) {
    const int Kp = L.Kp[l], K = L.K[l], E = L.E[l];
    if (L.o_wE[l] >= 0) {
      float* w = sm + L.o_wE[l];
      for (int idx = threadIdx.x; idx < L.Ep[l] * Kp; idx += GJ_THREADS) {
        int c = idx / Kp, k = idx - c * Kp;
        w[idx] = (c < E && k < K) ? __ldg(params + L.pW[l] + c * K + k) : 0.f;
      }
    }
    for (int c = threadIdx.x; c < L.Ep[l]; c += GJ_THREADS) sm[L.o_bE[l] + c] = c < E ? __ldg(params + L.pb[l] + c) : 0.f;
  }
  for (int c = threadIdx.x; c < L.E0p; c += GJ_THREADS)
    sm[L.o_wd + c] = c < L.E[0] ? __ldg(params + L.pW[0] + c * L.K[0] + 2 * L.H) : 0.f;
}

// rows [r0, r0 + 32) of h (zero padded to H columns, zero rows beyond N) and of the P or Q half of PQ
__device__ void load_block(const MPLayout& L, const float* __restrict__ hjet, const float* __restrict__ pqjet, int half,
                           int r0, float* sh, float* spq) {
  for (int idx = threadIdx.x; idx < 32 * L.H; idx += GJ_THREADS) {
    int n = idx / L.H, k = idx - n * L.H;
    sh[n * L.Hs + k] = (r0 + n < L.N && k < L.cols) ? __ldg(hjet + (size_t)(r0 + n) * L.ld + k) : 0.f;
  }
  for (int idx = threadIdx.x; idx < 32 * L.E0p; idx += GJ_THREADS) {
    int n = idx / L.E0p, c = idx - n * L.E0p;
    spq[n * L.E0s + c] = (r0 + n < L.N) ? __ldg(pqjet + (size_t)(r0 + n) * 2 * L.E0p + half * L.E0p + c) : 0.f;
  }
}

// First edge layer for one tile: a0[c][r] = leaky(P_i[c] + Q_j[c] + wd[c] d_ij). Row r = w*32 + lane,
// i = it*TI + w (within the i block), j = lane (within the j block). Optionally records d_ij per row.
template <int R>
__device__ void edge_layer0(const MPLayout& L, float* sm, int it, float* drow) {
  constexpr int TI = R / 32;
  constexpr int NPART = GJ_THREADS / R;
  const int r = threadIdx.x % R, part = threadIdx.x / R;
  const int w = r >> 5, lane = r & 31;
  const int il = it * TI + w;
  const float* hi = sm + L.o_h + il * L.Hs;
  const float* hj = sm + L.o_hj + lane * L.Hs;
  float d = 0.f;
  for (int k = 0; k < L.H; ++k) {
    float x = hj[k] - hi[k] + GJ_EPS;
    float s = (L.mink && k > 0) ? -1.f : 1.f;
    d = fmaf(s * x, x, d);
  }
  if (drow && part == 0) drow[r] = d;
  const float* P = sm + L.o_P + il * L.E0s;
  const float* Q = sm + L.o_Q + lane * L.E0s;
  const float* wd = sm + L.o_wd;
  float* a0 = sm + L.o_act[0];
  const int cn = L.E0p / NPART;
  for (int c = part * cn; c < (part + 1) * cn; ++c)
    a0[c * L.Rs + r] = gj_leaky(P[c] + Q[c] + wd[c] * d, L.alpha);
}

// out[n][r] = leaky(b[n] + sum_k in[k][r] W[n][k]); W natural (out,in) layout, padded [Np][Kp].
template <int R>
__device__ void simt_layer_fwd(const float* __restrict__ in, const float* __restrict__ W, const float* __restrict__ bias,
                               float* __restrict__ out, int Kp, int Np, int Rs, float alpha) {
  constexpr int NRG = R / 4;
  const int items = NRG * (Np / 8);
  for (int item = threadIdx.x; item < items; item += GJ_THREADS) {
    const int rg = item % NRG, cg = item / NRG;
    float acc[8][4];
#pragma unroll
    for (int q = 0; q < 8; ++q) { float b = bias[cg * 8 + q]; acc[q][0] = b; acc[q][1] = b; acc[q][2] = b; acc[q][3] = b; }
    const float* wbase = W + cg * 8 * Kp;
    for (int k = 0; k < Kp; k += 4) {
      float4 a0 = *reinterpret_cast<const float4*>(in + (k + 0) * Rs + rg * 4);
      float4 a1 = *reinterpret_cast<const float4*>(in + (k + 1) * Rs + rg * 4);
      float4 a2 = *reinterpret_cast<const float4*>(in + (k + 2) * Rs + rg * 4);
      float4 a3 = *reinterpret_cast<const float4*>(in + (k + 3) * Rs + rg * 4);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 w = *reinterpret_cast<const float4*>(wbase + q * Kp + k);
        acc[q][0] = fmaf(w.x, a0.x, acc[q][0]); acc[q][1] = fmaf(w.x, a0.y, acc[q][1]);
        acc[q][2] = fmaf(w.x, a0.z, acc[q][2]); acc[q][3] = fmaf(w.x, a0.w, acc[q][3]);
        acc[q][0] = fmaf(w.y, a1.x, acc[q][0]); acc[q][1] = fmaf(w.y, a1.y, acc[q][1]);
        acc[q][2] = fmaf(w.y, a1.z, acc[q][2]); acc[q][3] = fmaf(w.y, a1.w, acc[q][3]);
        acc[q][0] = fmaf(w.z, a2.x, acc[q][0]); acc[q][1] = fmaf(w.z, a2.y, acc[q][1]);
        acc[q][2] = fmaf(w.z, a2.z, acc[q][2]); acc[q][3] = fmaf(w.z, a2.w, acc[q][3]);
        acc[q][0] = fmaf(w.w, a3.x, acc[q][0]); acc[q][1] = fmaf(w.w, a3.y, acc[q][1]);
        acc[q][2] = fmaf(w.w, a3.z, acc[q][2]); acc[q][3] = fmaf(w.w, a3.w, acc[q][3]);
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float4 o = make_float4(gj_leaky(acc[q][0], alpha), gj_leaky(acc[q][1], alpha), gj_leaky(acc[q][2], alpha),
                             gj_leaky(acc[q][3], alpha));
      *reinterpret_cast<float4*>(out + (cg * 8 + q) * Rs + rg * 4) = o;
    }
  }
}

// In place: prev[k][r] <- (sum_c dz[c][r] W[c][k]) * leaky'(prev[k][r]);  W natural [Np][Kp].
template <int R>
__device__ void simt_layer_dgrad(const float* __restrict__ dz, const float* __restrict__ W, float* __restrict__ prev,
                                 int Kp, int Np, int Rs, float alpha) {
  constexpr int NRG = R / 4;
  const int items = NRG * (Kp / 8);
  for (int item = threadIdx.x; item < items; item += GJ_THREADS) {
    const int rg = item % NRG, kg = item / NRG;
    float acc[8][4];
#pragma unroll
    for (int q = 0; q < 8; ++q) { acc[q][0] = 0.f; acc[q][1] = 0.f; acc[q][2] = 0.f; acc[q][3] = 0.f; }
    for (int c = 0; c < Np; ++c) {
      float4 g = *reinterpret_cast<const float4*>(dz + c * Rs + rg * 4);
      float4 w0 = *reinterpret_cast<const float4*>(W + c * Kp + kg * 8);
      float4 w1 = *reinterpret_cast<const float4*>(W + c * Kp + kg * 8 + 4);
      float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        acc[q][0] = fmaf(wv[q], g.x, acc[q][0]); acc[q][1] = fmaf(wv[q], g.y, acc[q][1]);
        acc[q][2] = fmaf(wv[q], g.z, acc[q][2]); acc[q][3] = fmaf(wv[q], g.w, acc[q][3]);
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float4* p = reinterpret_cast<float4*>(prev + (kg * 8 + q) * Rs + rg * 4);
      float4 a = *p;
      *p = make_float4(acc[q][0] * gj_slope(a.x, alpha), acc[q][1] * gj_slope(a.y, alpha),
                       acc[q][2] * gj_slope(a.z, alpha), acc[q][3] * gj_slope(a.w, alpha));
    }
  }
}

// dW[c][k] += sum_r dz[c][r] a[k][r] (c < E, k < K stored unpadded at dW[c*K + k]); db[c] += sum_r dz[c][r].
template <int R>
__device__ void simt_layer_wgrad(const float* __restrict__ dz, const float* __restrict__ a, float* __restrict__ dW,
                                 float* __restrict__ db, int E, int K, int Np, int Kp, int Rs) {
  const int KT = Kp / 4, CT = Np / 4;
  for (int item = threadIdx.x; item < KT * CT; item += GJ_THREADS) {
    const int kt = item % KT, ct = item / KT;
    float acc[4][4];
    float bs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 4; ++q) { acc[q][0] = 0.f; acc[q][1] = 0.f; acc[q][2] = 0.f; acc[q][3] = 0.f; }
    for (int r = 0; r < R; r += 4) {
      float4 g[4], x[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        g[q] = *reinterpret_cast<const float4*>(dz + (ct + q * CT) * Rs + r);
        x[q] = *reinterpret_cast<const float4*>(a + (kt + q * KT) * Rs + r);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        bs[q] += (g[q].x + g[q].y) + (g[q].z + g[q].w);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          acc[q][p] = fmaf(g[q].x, x[p].x, acc[q][p]); acc[q][p] = fmaf(g[q].y, x[p].y, acc[q][p]);
          acc[q][p] = fmaf(g[q].z, x[p].z, acc[q][p]); acc[q][p] = fmaf(g[q].w, x[p].w, acc[q][p]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = ct + q * CT;
      if (c < E) {
#pragma unroll
        for (int p = 0; p < 4; ++p) { const int k = kt + p * KT; if (k < K) dW[c * K + k] += acc[q][p]; }
        if (kt == 0) db[c] += bs[q];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// forward: h, PQ -> e
// ------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(GJ_THREADS, 1)
edge_fwd_simt_kernel(const MPLayout L, const float* __restrict__ h, const float* __restrict__ pq,
                     const float* __restrict__ params, float* __restrict__ e_out) {
  extern __shared__ float4 smem_raw[];
  float* sm = reinterpret_cast<float*>(smem_raw);
  constexpr int TI = R / 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  stage_weights(L, params, sm);
  for (int jet = blockIdx.x; jet < L.B; jet += gridDim.x) {
    const float* hjet = h + (size_t)jet * L.N * L.ld;
    const float* pqjet = pq + (size_t)jet * L.N * 2 * L.E0p;
    for (int i0 = 0; i0 < L.N; i0 += GJ_IB) {
      const int ni = min(GJ_IB, L.N - i0);
      __syncthreads();
      load_block(L, hjet, pqjet, 0, i0, sm + L.o_h, sm + L.o_P);
      for (int idx = tid; idx < GJ_IB * L.ELs; idx += GJ_THREADS) sm[L.o_e + idx] = 0.f;
      for (int j0 = 0; j0 < L.N; j0 += 32) {
        const int nj = min(32, L.N - j0);
        __syncthreads();
        load_block(L, hjet, pqjet, 1, j0, sm + L.o_hj, sm + L.o_Q);
        __syncthreads();
        const int nit = (ni + TI - 1) / TI;
        for (int it = 0; it < nit; ++it) {
          edge_layer0<R>(L, sm, it, nullptr);
          __syncthreads();
          for (int l = 1; l < L.Le; ++l) {
            simt_layer_fwd<R>(sm + L.o_act[l - 1], sm + L.o_wE[l], sm + L.o_bE[l], sm + L.o_act[l], L.Kp[l], L.Ep[l],
                              L.Rs, L.alpha);
            __syncthreads();
          }
          // e_i += sum_j a_last[i,j,:]  (masked AFTER the activation: padded i / j contribute nothing)
          const float* al = sm + L.o_act[L.Le - 1];
          for (int pair = warp; pair < TI * L.ELp; pair += GJ_THREADS / 32) {
            const int w = pair / L.ELp, c = pair - w * L.ELp;
            const int il = it * TI + w;
            const bool valid = il < ni && lane < nj;
            float v = valid ? al[c * L.Rs + w * 32 + lane] : 0.f;
            v = gj_warp_sum(v);
            if (lane == 0) sm[L.o_e + il * L.ELs + c] += v;
          }
          __syncthreads();
        }
      }
      for (int idx = tid; idx < ni * L.EL; idx += GJ_THREADS) {
        int n = idx / L.EL, c = idx - n * L.EL;
        e_out[((size_t)jet * L.N + i0 + n) * L.EL + c] = sm[L.o_e + n * L.ELs + c];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward: h, PQ, de -> dPQ, dh += distance-path gradient, per-CTA partial of the edge parameters
// ------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(GJ_THREADS, 1)
edge_bwd_simt_kernel(const MPLayout L, const float* __restrict__ h, const float* __restrict__ pq,
                     const float* __restrict__ params, const float* __restrict__ de, float* __restrict__ dpq,
                     float* __restrict__ dh, float* __restrict__ part) {
  extern __shared__ float4 smem_raw[];
  float* sm = reinterpret_cast<float*>(smem_raw);
  constexpr int TI = R / 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  stage_weights(L, params, sm);
  float* out = part + (size_t)blockIdx.x * L.n_edge_dpar;
  float* dpar = L.o_dpar >= 0 ? sm + L.o_dpar : out;      // wide layers: accumulate in the CTA's global partial
  for (int idx = tid; idx < L.n_edge_dpar; idx += GJ_THREADS) dpar[idx] = 0.f;
  const int H = L.H, K0 = L.K[0], W2 = 2 * L.E0p;
  for (int jet = blockIdx.x; jet < L.B; jet += gridDim.x) {
    const float* hjet = h + (size_t)jet * L.N * L.ld;
    const float* pqjet = pq + (size_t)jet * L.N * W2;
    float* dpqjet = dpq + (size_t)jet * L.N * W2;
    float* dhjet = dh + (size_t)jet * L.N * L.ld;
    for (int i0 = 0; i0 < L.N; i0 += GJ_IB) {
      const int ni = min(GJ_IB, L.N - i0);
      __syncthreads();
      load_block(L, hjet, pqjet, 0, i0, sm + L.o_h, sm + L.o_P);
      for (int idx = tid; idx < GJ_IB * L.ELs; idx += GJ_THREADS) {
        int n = idx / L.ELs, c = idx - n * L.ELs;
        sm[L.o_e + idx] = (n < ni && c < L.EL) ? __ldg(de + ((size_t)jet * L.N + i0 + n) * L.EL + c) : 0.f;
      }
      for (int idx = tid; idx < GJ_IB * L.E0s; idx += GJ_THREADS) sm[L.o_dP + idx] = 0.f;
      for (int idx = tid; idx < GJ_IB * L.Hs; idx += GJ_THREADS) sm[L.o_dh + idx] = 0.f;
      for (int j0 = 0; j0 < L.N; j0 += 32) {
        const int nj = min(32, L.N - j0);
        __syncthreads();
        load_block(L, hjet, pqjet, 1, j0, sm + L.o_hj, sm + L.o_Q);
        for (int idx = tid; idx < 32 * L.E0s; idx += GJ_THREADS) sm[L.o_dQ + idx] = 0.f;
        for (int idx = tid; idx < GJ_IB * L.Gs; idx += GJ_THREADS) sm[L.o_G + idx] = 0.f;
        __syncthreads();
        const int nit = (ni + TI - 1) / TI;
        float* drow = sm + L.o_drow;
        for (int it = 0; it < nit; ++it) {
          edge_layer0<R>(L, sm, it, drow);
          __syncthreads();
          for (int l = 1; l < L.Le; ++l) {
            simt_layer_fwd<R>(sm + L.o_act[l - 1], L.o_wE[l] >= 0 ? sm + L.o_wE[l] : params + L.pW[l], sm + L.o_bE[l], sm + L.o_act[l], L.Kp[l], L.Ep[l],
                              L.Rs, L.alpha);
            __syncthreads();
          }
          {  // dz_last = de_i * leaky'(a_last), zero on padded rows (this masks everything downstream)
            float* al = sm + L.o_act[L.Le - 1];
            for (int idx = tid; idx < L.ELp * R; idx += GJ_THREADS) {
              int c = idx / R, r = idx - c * R;
              int w = r >> 5, ln = r & 31;
              int il = it * TI + w;
              bool valid = il < ni && ln < nj;
              float a = al[c * L.Rs + r];
              al[c * L.Rs + r] = valid ? sm[L.o_e + il * L.ELs + c] * gj_slope(a, L.alpha) : 0.f;
            }
          }
          __syncthreads();
          for (int l = L.Le - 1; l >= 1; --l) {
            simt_layer_wgrad<R>(sm + L.o_act[l], sm + L.o_act[l - 1], dpar + L.pW[l], dpar + L.pb[l], L.E[l], L.K[l],
                                L.Ep[l], L.Kp[l], L.Rs);
            __syncthreads();
            simt_layer_dgrad<R>(sm + L.o_act[l], L.o_wE[l] >= 0 ? sm + L.o_wE[l] : params + L.pW[l], sm + L.o_act[l - 1], L.Kp[l], L.Ep[l], L.Rs, L.alpha);
            __syncthreads();
          }
          // ---- consume dz0 (in act[0]) ----
          const float* z0 = sm + L.o_act[0];
          // dP_i[c] += sum_j dz0 ; d(wd)[c] += sum_r dz0 d_r
          for (int c = warp; c < L.E0p; c += GJ_THREADS / 32) {
            float sd = 0.f;
#pragma unroll
            for (int w = 0; w < TI; ++w) {
              float v = z0[c * L.Rs + w * 32 + lane];
              float s = gj_warp_sum(v);
              sd += gj_warp_sum(v * drow[w * 32 + lane]);
              if (lane == 0) sm[L.o_dP + (it * TI + w) * L.E0s + c] += s;   // it*TI+w < 32; padded rows add 0
            }
            if (lane == 0 && c < L.E[0]) dpar[L.pW[0] + c * K0 + 2 * H] += sd;
          }
          // dQ_j[c] += sum_i dz0
          for (int idx = tid; idx < 32 * L.E0p; idx += GJ_THREADS) {
            int c = idx >> 5, ln = idx & 31;
            float acc = 0.f;
#pragma unroll
            for (int w = 0; w < TI; ++w) acc += z0[c * L.Rs + w * 32 + ln];
            sm[L.o_dQ + ln * L.E0s + c] += acc;
          }
          // G_ij = d loss / d d_ij = sum_c dz0 wd[c]
          for (int r = tid; r < R; r += GJ_THREADS) {
            float acc = 0.f;
            for (int c = 0; c < L.E0p; ++c) acc = fmaf(z0[c * L.Rs + r], sm[L.o_wd + c], acc);
            int w = r >> 5, ln = r & 31;
            sm[L.o_G + (it * TI + w) * L.Gs + ln] = acc;
          }
          __syncthreads();
        }
        // ---- (i block, j block) epilogue ----
        // dQ rows leave for node_pre_bwd: first i block writes, later ones accumulate (same thread, fixed order)
        for (int idx = tid; idx < nj * L.E0p; idx += GJ_THREADS) {
          int n = idx / L.E0p, c = idx - n * L.E0p;
          float* p = dpqjet + (size_t)(j0 + n) * W2 + L.E0p + c;
          float v = sm[L.o_dQ + n * L.E0s + c];
          *p = (i0 == 0) ? v : (*p + v);
        }
        // distance term: d_ij = sum_k s_k (h_j - h_i + eps)^2  =>  dh_j += 2 s_k G_ij diff ; dh_i -= 2 s_k G_ij diff
        for (int idx = tid; idx < nj * H; idx += GJ_THREADS) {
          int n = idx / H, k = idx - n * H;
          const float sgn = (L.mink && k > 0) ? -2.f : 2.f;
          const float hjk = sm[L.o_hj + n * L.Hs + k];
          float acc = 0.f;
          for (int i = 0; i < ni; ++i) acc = fmaf(sm[L.o_G + i * L.Gs + n], hjk - sm[L.o_h + i * L.Hs + k] + GJ_EPS, acc);
          if (k < L.cols) dhjet[(size_t)(j0 + n) * L.ld + k] += sgn * acc;
        }
        for (int idx = tid; idx < ni * H; idx += GJ_THREADS) {
          int n = idx / H, k = idx - n * H;
          const float sgn = (L.mink && k > 0) ? -2.f : 2.f;
          const float hik = sm[L.o_h + n * L.Hs + k];
          float acc = 0.f;
          for (int j = 0; j < nj; ++j) acc = fmaf(sm[L.o_G + n * L.Gs + j], sm[L.o_hj + j * L.Hs + k] - hik + GJ_EPS, acc);
          sm[L.o_dh + n * L.Hs + k] -= sgn * acc;
        }
      }
      __syncthreads();
      // ---- i block epilogue: dP rows are complete; the i-side distance gradient joins dh ----
      for (int idx = tid; idx < ni * L.E0p; idx += GJ_THREADS) {
        int n = idx / L.E0p, c = idx - n * L.E0p;
        dpqjet[(size_t)(i0 + n) * W2 + c] = sm[L.o_dP + n * L.E0s + c];
      }
      for (int idx = tid; idx < ni * L.cols; idx += GJ_THREADS) {
        int n = idx / L.cols, k = idx - n * L.cols;
        dhjet[(size_t)(i0 + n) * L.ld + k] += sm[L.o_dh + n * L.Hs + k];
      }
    }
  }
  __syncthreads();
  if (dpar != out)
    for (int idx = tid; idx < L.n_edge_dpar; idx += GJ_THREADS) out[idx] = dpar[idx];
}

// dparams[p] = sum_cta part[cta][p] over the edge-parameter block, skipping the slots node_pre_bwd owns
// (the Wa | Wb columns of W0 and b0): of the first layer only the wd column (index 2H of each row) is ours.
__global__ void reduce_edge_partials_kernel(const float* __restrict__ part, int nparts, int n, int E0, int K0,
                                            float* __restrict__ out) {
  __shared__ float red[8][33];
  const int p = blockIdx.x * 32 + threadIdx.x;
  const int first = E0 * K0 + E0;          // W0 then b0
  const bool mine = p < n && !(p < first && !(p < E0 * K0 && (p % K0) == K0 - 1));
  float acc = 0.f;
  if (mine)
    for (int c = threadIdx.y; c < nparts; c += 8) acc += part[(size_t)c * n + p];
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && mine) {
    float tot = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) tot += red[y][threadIdx.x];     // fixed order => deterministic
    out[p] = tot;
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
int gj_num_sms();
void gj_set_error(const char* fmt, ...);

constexpr int kFwdR = 128;
constexpr int kBwdR = 64;

int gj_edge_grid(int batch) {
  int sms = gj_num_sms();
  return batch < sms ? (batch > 0 ? batch : 1) : sms;
}

int gj_edge_fwd_simt(MPLayout L, const float* h, const float* pq, const float* params, float* e_out, cudaStream_t stream) {
  int bytes = plan_smem(&L, kFwdR, false);
  auto kern = edge_fwd_simt_kernel<kFwdR>;
  if (bytes > 227 * 1024) {      // wide layers: half the rows per tile
    bytes = plan_smem(&L, kFwdR / 2, false);
    kern = edge_fwd_simt_kernel<kFwdR / 2>;
  }
  if (bytes > 227 * 1024) { gj_set_error("gj_mp_step_fwd(fp32): needs %d B shared memory (> 227 KB)", bytes); return GJ_ERR_SMEM; }
  cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (ce != cudaSuccess) { gj_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  kern<<<gj_edge_grid(L.B), GJ_THREADS, bytes, stream>>>(L, h, pq, params, e_out);
  ce = cudaGetLastError();
  if (ce != cudaSuccess) { gj_set_error("edge_fwd_simt launch: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  return GJ_OK;
}

// both fp32 edge kernels have a shared-memory plan for these widths (otherwise the step runs the materialised path of dense.cu)
bool gj_edge_simt_fits(MPLayout L, const float* params) {
  MPLayout F = L;
  if (plan_smem(&F, kFwdR, false) > 227 * 1024 && plan_smem(&F, kFwdR / 2, false) > 227 * 1024) return false;
  MPLayout Bk = L;
  if (plan_smem(&Bk, kBwdR, true) <= 227 * 1024 || plan_smem(&Bk, kBwdR, true, true) <= 227 * 1024) return true;
  bool in_place = true;
  for (int l = 1; l < L.Le; ++l)
    in_place = in_place && L.E[l] == L.Ep[l] && L.K[l] == L.Kp[l] && (params == nullptr || (reinterpret_cast<uintptr_t>(params + L.pW[l]) & 15) == 0);
  return in_place && params != nullptr && plan_smem(&Bk, kBwdR, true, true, true) <= 227 * 1024;
}

// number of per-CTA partials and floats per partial of the edge-parameter gradients
void gj_edge_bwd_simt_partials(const MPLayout& L, int* nparts, int* n) { *nparts = gj_edge_grid(L.B); *n = L.pV[0]; }

int gj_reduce_edge_partials(const MPLayout& L, const float* part, int nparts, float* dparams, cudaStream_t stream);

int gj_edge_bwd_simt(MPLayout L, const float* h, const float* pq, const float* params, const float* de, float* dpq, float* dh,
                     float* dparams, float* part, cudaStream_t stream) {
  int bytes = plan_smem(&L, kBwdR, true);
  if (bytes > 227 * 1024) bytes = plan_smem(&L, kBwdR, true, true);      // wide layers: gradient accumulators in global memory
  if (bytes > 227 * 1024) {                                               // ... and the weights read in place
    bool in_place = true;
    for (int l = 1; l < L.Le; ++l)      // same layout as the staged copy, and 16-byte aligned rows
      in_place = in_place && L.E[l] == L.Ep[l] && L.K[l] == L.Kp[l] && (reinterpret_cast<uintptr_t>(params + L.pW[l]) & 15) == 0;
    if (in_place) bytes = plan_smem(&L, kBwdR, true, true, true);
  }
  if (bytes > 227 * 1024) { gj_set_error("gj_mp_step_bwd(fp32): needs %d B shared memory (> 227 KB)", bytes); return GJ_ERR_SMEM; }
  auto kern = edge_bwd_simt_kernel<kBwdR>;
  cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (ce != cudaSuccess) { gj_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  const int grid = gj_edge_grid(L.B);
  kern<<<grid, GJ_THREADS, bytes, stream>>>(L, h, pq, params, de, dpq, dh, part);
  return gj_reduce_edge_partials(L, part, grid, dparams, stream);
}

int gj_reduce_edge_partials(const MPLayout& L, const float* part, int nparts, float* dparams, cudaStream_t stream) {
  const int n = L.pV[0];
  reduce_edge_partials_kernel<<<(n + 31) / 32, dim3(32, 8), 0, stream>>>(part, nparts, n, L.E[0], L.K[0], dparams);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) { gj_set_error("edge_bwd launch: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  return GJ_OK;
}
