// fp32 (GJ_PREC_FP32) message-passing step: SIMT FFMA kernels, forward and backward.
// One CTA walks whole jets; the N x N pair set lives only in shared memory, tile by tile.
// Replaces reference models/graphnet.py:154-168 (_getA, _edge_conv, _concat, _aggregate) and its
// autograd adjoint.  See gj_common.cuh for the per-jet evaluation order.
#include "mp_helpers.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// shared-memory plan
// ------------------------------------------------------------------------------------------------
struct Carver {
  int off = 0;
  int take(int n) { int o = off; off += (n + 3) & ~3; return o; }  // keep every region 16-byte aligned
};

int plan_smem(MPLayout* L, int R, bool backward) {
  L->R = R; L->Rs = R + 4;
  Carver c;
  for (int l = 1; l < L->Le; ++l) L->o_wE[l] = c.take(L->Ep[l] * L->Kp[l]);
  for (int l = 0; l < L->Le; ++l) L->o_bE[l] = c.take(L->Ep[l]);
  L->o_wa = c.take(L->E0p * L->Hs);
  L->o_wb = c.take(L->E0p * L->Hs);
  L->o_wd = c.take(L->E0p);
  for (int m = 0; m < L->Ln; ++m) { L->o_V[m] = c.take(L->O[m] * L->Is[m]); L->o_c[m] = c.take(L->O[m]); }
  L->o_h = c.take(L->N * L->Hs);
  L->o_Q = c.take(L->N * L->E0s);
  L->o_P = c.take(GJ_IB * L->E0s);
  L->o_e = c.take(GJ_IB * L->ELs);
  if (!backward) {
    int w0 = 0, w1 = 0;
    for (int l = 0; l < L->Le; ++l) { if (l & 1) { if (L->Ep[l] > w1) w1 = L->Ep[l]; } else { if (L->Ep[l] > w0) w0 = L->Ep[l]; } }
    int a0 = c.take(w0 * L->Rs), a1 = c.take((w1 ? w1 : 1) * L->Rs);
    for (int l = 0; l < L->Le; ++l) L->o_act[l] = (l & 1) ? a1 : a0;
    L->o_node[0] = c.take(GJ_IB * L->Ws);
    L->o_node[1] = c.take(GJ_IB * L->Ws);
  } else {
    for (int l = 0; l < L->Le; ++l) L->o_act[l] = c.take(L->Ep[l] * L->Rs);
    for (int m = 0; m <= L->Ln + 2; ++m) L->o_node[m] = c.take(GJ_IB * L->Ws);  // Y[0..Ln], two gradient buffers
    L->o_dpar = c.take(L->nparams);
    L->o_dh = c.take(L->N * L->Hs);
    L->o_dQ = c.take(L->N * L->E0s);
    L->o_dP = c.take(GJ_IB * L->E0s);
    L->o_de = c.take(GJ_IB * L->ELs);
    L->Gs = L->Npad + 1;
    L->o_G = c.take(GJ_IB * L->Gs);
    L->o_drow = c.take(R);
  }
  L->smem_floats = c.off;
  return c.off * 4;
}

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
// First edge layer for one tile: a0[c][r] = leaky(P_i[c] + Q_j[c] + wd[c] d_ij). Row r = w*32 + lane,
// i = i0 + it*TI + w, j = jb*32 + lane (clamped when padded). Optionally records d_ij per row.
template <int R>
__device__ void edge_layer0(const MPLayout& L, float* sm, int i0, int ni, int it, int jb, float* drow) {
  constexpr int TI = R / 32;
  constexpr int NPART = GJ_THREADS / R;
  const int r = threadIdx.x % R, part = threadIdx.x / R;
  const int w = r >> 5, lane = r & 31;
  int il = it * TI + w; if (il > ni - 1) il = ni - 1;
  int j = jb * 32 + lane; if (j > L.N - 1) j = L.N - 1;
  const float* hi = sm + L.o_h + (i0 + il) * L.Hs;
  const float* hj = sm + L.o_h + j * L.Hs;
  float d = 0.f;
  for (int k = 0; k < L.H; ++k) {
    float x = hj[k] - hi[k] + GJ_EPS;
    float s = (L.mink && k > 0) ? -1.f : 1.f;
    d = fmaf(s * x, x, d);
  }
  if (drow && part == 0) drow[r] = d;
  const float* P = sm + L.o_P + il * L.E0s;
  const float* Q = sm + L.o_Q + j * L.E0s;
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
// Thread tile 4 x 4 with interleaved rows so consecutive lanes touch consecutive activation rows.
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
// forward kernel
// ------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(GJ_THREADS, 1)
mp_fwd_simt_kernel(const MPLayout L, const float* __restrict__ h, const float* __restrict__ params,
                   float* __restrict__ h_out, float* __restrict__ e_out) {
  extern __shared__ float4 smem_raw[];
  float* sm = reinterpret_cast<float*>(smem_raw);
  constexpr int TI = R / 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  gj_stage_small_weights(L, params, sm, threadIdx.x, GJ_THREADS);
  gj_stage_edge_weights_f32(L, params, sm, threadIdx.x, GJ_THREADS);
  __syncthreads();
  for (int jet = blockIdx.x; jet < L.B; jet += gridDim.x) {
    const float* hg = h + (size_t)jet * L.N * L.ld;
    for (int idx = tid; idx < L.N * L.H; idx += GJ_THREADS) {
      int n = idx / L.H, k = idx - n * L.H;
      sm[L.o_h + n * L.Hs + k] = k < L.cols ? __ldg(hg + n * L.ld + k) : 0.f;
    }
    __syncthreads();
    gj_node_project(L, sm + L.o_wb, nullptr, sm + L.o_h, L.N, sm + L.o_Q, tid, GJ_THREADS);
    for (int i0 = 0; i0 < L.N; i0 += GJ_IB) {
      const int ni = min(GJ_IB, L.N - i0);
      gj_node_project(L, sm + L.o_wa, sm + L.o_bE[0], sm + L.o_h + i0 * L.Hs, ni, sm + L.o_P, tid, GJ_THREADS);
      for (int idx = tid; idx < GJ_IB * L.ELs; idx += GJ_THREADS) sm[L.o_e + idx] = 0.f;
      __syncthreads();
      const int nit = (ni + TI - 1) / TI;
      for (int it = 0; it < nit; ++it) {
        for (int jb = 0; jb < L.Npad / 32; ++jb) {
          edge_layer0<R>(L, sm, i0, ni, it, jb, nullptr);
          __syncthreads();
          for (int l = 1; l < L.Le; ++l) {
            simt_layer_fwd<R>(sm + L.o_act[l - 1], sm + L.o_wE[l], sm + L.o_bE[l], sm + L.o_act[l], L.Kp[l], L.Ep[l],
                              L.Rs, L.alpha);
            __syncthreads();
          }
          // e_i += sum_j a_last[i,j,:]  (masked: padded i / j contribute nothing)
          const float* al = sm + L.o_act[L.Le - 1];
          for (int pair = warp; pair < TI * L.ELp; pair += GJ_THREADS / 32) {
            const int w = pair / L.ELp, c = pair - w * L.ELp;
            const int il = it * TI + w;
            const bool valid = il < ni && (jb * 32 + lane) < L.N;
            float v = valid ? al[c * L.Rs + w * 32 + lane] : 0.f;
            v = gj_warp_sum(v);
            if (lane == 0) sm[L.o_e + il * L.ELs + c] += v;
          }
          __syncthreads();
        }
      }
      // node MLP on [e_i | h_i]
      float* bufA = sm + L.o_node[0];
      float* bufB = sm + L.o_node[1];
      const int I0 = L.EL + L.H;
      for (int idx = tid; idx < ni * I0; idx += GJ_THREADS) {
        int n = idx / I0, k = idx - n * I0;
        float v = k < L.EL ? sm[L.o_e + n * L.ELs + k] : sm[L.o_h + (i0 + n) * L.Hs + (k - L.EL)];
        bufA[n * L.Ws + k] = v;
        if (e_out && k < L.EL) e_out[((size_t)jet * L.N + i0 + n) * L.EL + k] = v;
      }
      __syncthreads();
      for (int m = 0; m < L.Ln; ++m) {
        gj_node_layer_fwd(L, sm, m, bufA, bufB, ni, tid, GJ_THREADS);
        __syncthreads();
        float* t = bufA; bufA = bufB; bufB = t;
      }
      for (int idx = tid; idx < ni * L.Hout; idx += GJ_THREADS) {
        int n = idx / L.Hout, o = idx - n * L.Hout;
        h_out[((size_t)jet * L.N + i0 + n) * L.Hout + o] = bufA[n * L.Ws + o];
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward kernel
// ------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(GJ_THREADS, 1)
mp_bwd_simt_kernel(const MPLayout L, const float* __restrict__ h, const float* __restrict__ e_saved,
                   const float* __restrict__ params, const float* __restrict__ dh_out, float* __restrict__ dh,
                   float* __restrict__ ws) {
  extern __shared__ float4 smem_raw[];
  float* sm = reinterpret_cast<float*>(smem_raw);
  constexpr int TI = R / 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  gj_stage_small_weights(L, params, sm, threadIdx.x, GJ_THREADS);
  gj_stage_edge_weights_f32(L, params, sm, threadIdx.x, GJ_THREADS);
  float* dpar = sm + L.o_dpar;
  for (int idx = tid; idx < L.nparams; idx += GJ_THREADS) dpar[idx] = 0.f;
  __syncthreads();
  const int H = L.H, K0 = L.K[0];
  for (int jet = blockIdx.x; jet < L.B; jet += gridDim.x) {
    const float* hg = h + (size_t)jet * L.N * L.ld;
    for (int idx = tid; idx < L.N * H; idx += GJ_THREADS) {
      int n = idx / H, k = idx - n * H;
      sm[L.o_h + n * L.Hs + k] = k < L.cols ? __ldg(hg + n * L.ld + k) : 0.f;
    }
    for (int idx = tid; idx < L.N * L.Hs; idx += GJ_THREADS) sm[L.o_dh + idx] = 0.f;
    for (int idx = tid; idx < L.N * L.E0s; idx += GJ_THREADS) sm[L.o_dQ + idx] = 0.f;
    __syncthreads();
    gj_node_project(L, sm + L.o_wb, nullptr, sm + L.o_h, L.N, sm + L.o_Q, tid, GJ_THREADS);
    for (int i0 = 0; i0 < L.N; i0 += GJ_IB) {
      const int ni = min(GJ_IB, L.N - i0);
      gj_node_project(L, sm + L.o_wa, sm + L.o_bE[0], sm + L.o_h + i0 * L.Hs, ni, sm + L.o_P, tid, GJ_THREADS);
      for (int idx = tid; idx < GJ_IB * L.E0s; idx += GJ_THREADS) sm[L.o_dP + idx] = 0.f;
      for (int idx = tid; idx < GJ_IB * L.ELs; idx += GJ_THREADS) sm[L.o_de + idx] = 0.f;
      // ---- node MLP: recompute forward from the saved edge aggregate, then its adjoint ----
      {
        float* Y0 = sm + L.o_node[0];
        const int I0 = L.EL + H;
        for (int idx = tid; idx < ni * I0; idx += GJ_THREADS) {
          int n = idx / I0, k = idx - n * I0;
          Y0[n * L.Ws + k] = k < L.EL ? __ldg(e_saved + ((size_t)jet * L.N + i0 + n) * L.EL + k)
                                      : sm[L.o_h + (i0 + n) * L.Hs + (k - L.EL)];
        }
        __syncthreads();
        for (int m = 0; m < L.Ln; ++m) {
          gj_node_layer_fwd(L, sm, m, sm + L.o_node[m], sm + L.o_node[m + 1], ni, tid, GJ_THREADS);
          __syncthreads();
        }
        float* g = sm + L.o_node[L.Ln + 1];
        float* gp = sm + L.o_node[L.Ln + 2];
        // gz = dh_out * leaky'(y_last)
        for (int idx = tid; idx < ni * L.Hout; idx += GJ_THREADS) {
          int n = idx / L.Hout, o = idx - n * L.Hout;
          float y = sm[L.o_node[L.Ln] + n * L.Ws + o];
          g[n * L.Ws + o] = __ldg(dh_out + ((size_t)jet * L.N + i0 + n) * L.Hout + o) * gj_slope(y, L.alpha);
        }
        __syncthreads();
        for (int m = L.Ln - 1; m >= 0; --m) {
          const int O = L.O[m], I = L.I[m], Is = L.Is[m];
          const float* Yin = sm + L.o_node[m];
          // dV[o][k] += sum_n gz[n][o] y[n][k];  dc[o] += sum_n gz[n][o]
          for (int idx = tid; idx < O * (I + 1); idx += GJ_THREADS) {
            int o = idx / (I + 1), k = idx - o * (I + 1);
            float acc = 0.f;
            if (k < I) { for (int n = 0; n < ni; ++n) acc = fmaf(g[n * L.Ws + o], Yin[n * L.Ws + k], acc); dpar[L.pV[m] + o * I + k] += acc; }
            else { for (int n = 0; n < ni; ++n) acc += g[n * L.Ws + o]; dpar[L.pc[m] + o] += acc; }
          }
          // g_prev[n][k] = (sum_o gz[n][o] V[o][k]) * leaky'(y_in[n][k])   (no slope for the MLP input)
          const float* V = sm + L.o_V[m];
          for (int idx = tid; idx < ni * I; idx += GJ_THREADS) {
            int n = idx / I, k = idx - n * I;
            float acc = 0.f;
            for (int o = 0; o < O; ++o) acc = fmaf(g[n * L.Ws + o], V[o * Is + k], acc);
            if (m > 0) acc *= gj_slope(Yin[n * L.Ws + k], L.alpha);
            gp[n * L.Ws + k] = acc;
          }
          __syncthreads();
          float* t = g; g = gp; gp = t;
        }
        // g holds d[e | h]
        for (int idx = tid; idx < ni * I0; idx += GJ_THREADS) {
          int n = idx / I0, k = idx - n * I0;
          float v = g[n * L.Ws + k];
          if (k < L.EL) sm[L.o_de + n * L.ELs + k] = v;
          else sm[L.o_dh + (i0 + n) * L.Hs + (k - L.EL)] += v;
        }
      }
      __syncthreads();
      // ---- edge MLP, tile by tile: recompute, then dgrad / wgrad ----
      const int nit = (ni + TI - 1) / TI;
      float* drow = sm + L.o_drow;
      for (int it = 0; it < nit; ++it) {
        for (int jb = 0; jb < L.Npad / 32; ++jb) {
          edge_layer0<R>(L, sm, i0, ni, it, jb, drow);
          __syncthreads();
          for (int l = 1; l < L.Le; ++l) {
            simt_layer_fwd<R>(sm + L.o_act[l - 1], sm + L.o_wE[l], sm + L.o_bE[l], sm + L.o_act[l], L.Kp[l], L.Ep[l],
                              L.Rs, L.alpha);
            __syncthreads();
          }
          {  // dz_last = de_i * leaky'(a_last), zero on padded rows (this masks everything downstream)
            float* al = sm + L.o_act[L.Le - 1];
            for (int idx = tid; idx < L.ELp * R; idx += GJ_THREADS) {
              int c = idx / R, r = idx - c * R;
              int w = r >> 5, ln = r & 31;
              int il = it * TI + w;
              bool valid = il < ni && (jb * 32 + ln) < L.N;
              float a = al[c * L.Rs + r];
              al[c * L.Rs + r] = valid ? sm[L.o_de + il * L.ELs + c] * gj_slope(a, L.alpha) : 0.f;
            }
          }
          __syncthreads();
          for (int l = L.Le - 1; l >= 1; --l) {
            simt_layer_wgrad<R>(sm + L.o_act[l], sm + L.o_act[l - 1], dpar + L.pW[l], dpar + L.pb[l], L.E[l], L.K[l],
                                L.Ep[l], L.Kp[l], L.Rs);
            __syncthreads();
            simt_layer_dgrad<R>(sm + L.o_act[l], sm + L.o_wE[l], sm + L.o_act[l - 1], L.Kp[l], L.Ep[l], L.Rs, L.alpha);
            __syncthreads();
          }
          if (L.Le == 1) {
            // single edge layer: dz_last above already is dz0 (slope applied, masked)
          }
          // ---- consume dz0 (in act[0]) ----
          const float* z0 = sm + L.o_act[0];
          // dP_i[c] += sum_j dz0 ; d(wd)[c] += sum_r dz0 d_r  (one warp per (w,c) / per c)
          for (int c = warp; c < L.E0p; c += GJ_THREADS / 32) {
            float sd = 0.f;
#pragma unroll
            for (int w = 0; w < TI; ++w) {
              float v = z0[c * L.Rs + w * 32 + lane];
              float s = gj_warp_sum(v);
              sd += gj_warp_sum(v * drow[w * 32 + lane]);
              if (lane == 0) sm[L.o_dP + (it * TI + w) * L.E0s + c] += s;  // it*TI+w < 32; padded rows add 0
            }
            if (lane == 0 && c < L.E[0]) dpar[L.pW[0] + c * K0 + 2 * H] += sd;
          }
          // dQ_j[c] += sum_i dz0
          for (int idx = tid; idx < 32 * L.E0p; idx += GJ_THREADS) {
            int c = idx >> 5, ln = idx & 31;
            int j = jb * 32 + ln;
            if (j < L.N) {
              float acc = 0.f;
#pragma unroll
              for (int w = 0; w < TI; ++w) acc += z0[c * L.Rs + w * 32 + ln];
              sm[L.o_dQ + j * L.E0s + c] += acc;
            }
          }
          // G_ij = d loss / d d_ij = sum_c dz0 wd[c]
          for (int r = tid; r < R; r += GJ_THREADS) {
            float acc = 0.f;
            for (int c = 0; c < L.E0p; ++c) acc = fmaf(z0[c * L.Rs + r], sm[L.o_wd + c], acc);
            int w = r >> 5, ln = r & 31;
            sm[L.o_G + (it * TI + w) * L.Gs + jb * 32 + ln] = acc;
          }
          __syncthreads();
        }
      }
      // ---- i-block epilogue: dP -> dh_i, dWa, db0 ; G -> dh ----
      {
        const float* dP = sm + L.o_dP;
        // dh_i[k] += sum_c dP[i][c] Wa[c][k]
        for (int idx = tid; idx < ni * H; idx += GJ_THREADS) {
          int n = idx / H, k = idx - n * H;
          float acc = 0.f;
          for (int c = 0; c < L.E[0]; ++c) acc = fmaf(dP[n * L.E0s + c], sm[L.o_wa + c * L.Hs + k], acc);
          // distance term, i side: dh_i[k] -= sum_j 2 G_ij s_k (h_j[k] - h_i[k] + eps)
          const float sgn = (L.mink && k > 0) ? -2.f : 2.f;
          const float hik = sm[L.o_h + (i0 + n) * L.Hs + k];
          float accd = 0.f;
          for (int j = 0; j < L.N; ++j)
            accd = fmaf(sm[L.o_G + n * L.Gs + j], sm[L.o_h + j * L.Hs + k] - hik + GJ_EPS, accd);
          sm[L.o_dh + (i0 + n) * L.Hs + k] += acc - sgn * accd;
        }
        // dWa[c][k] += sum_i dP[i][c] h_i[k] ; db0[c] += sum_i dP[i][c]
        for (int idx = tid; idx < L.E[0] * (H + 1); idx += GJ_THREADS) {
          int c = idx / (H + 1), k = idx - c * (H + 1);
          float acc = 0.f;
          if (k < H) { for (int n = 0; n < ni; ++n) acc = fmaf(dP[n * L.E0s + c], sm[L.o_h + (i0 + n) * L.Hs + k], acc); dpar[L.pW[0] + c * K0 + k] += acc; }
          else { for (int n = 0; n < ni; ++n) acc += dP[n * L.E0s + c]; dpar[L.pb[0] + c] += acc; }
        }
        __syncthreads();
        // distance term, j side: dh_j[k] += sum_{i in block} 2 G_ij s_k (h_j[k] - h_i[k] + eps)
        for (int idx = tid; idx < L.N * H; idx += GJ_THREADS) {
          int j = idx / H, k = idx - j * H;
          const float sgn = (L.mink && k > 0) ? -2.f : 2.f;
          const float hjk = sm[L.o_h + j * L.Hs + k];
          float accd = 0.f;
          for (int n = 0; n < ni; ++n)
            accd = fmaf(sm[L.o_G + n * L.Gs + j], hjk - sm[L.o_h + (i0 + n) * L.Hs + k] + GJ_EPS, accd);
          sm[L.o_dh + j * L.Hs + k] += sgn * accd;
        }
        __syncthreads();
      }
    }
    // ---- jet epilogue: dQ -> dh_j, dWb ----
    {
      const float* dQ = sm + L.o_dQ;
      for (int idx = tid; idx < L.N * H; idx += GJ_THREADS) {
        int n = idx / H, k = idx - n * H;
        float acc = 0.f;
        for (int c = 0; c < L.E[0]; ++c) acc = fmaf(dQ[n * L.E0s + c], sm[L.o_wb + c * L.Hs + k], acc);
        if (k < L.cols) dh[((size_t)jet * L.N + n) * L.ld + k] = sm[L.o_dh + n * L.Hs + k] + acc;
      }
      for (int idx = tid; idx < L.E[0] * H; idx += GJ_THREADS) {
        int c = idx / H, k = idx - c * H;
        float acc = 0.f;
        for (int n = 0; n < L.N; ++n) acc = fmaf(dQ[n * L.E0s + c], sm[L.o_h + n * L.Hs + k], acc);
        dpar[L.pW[0] + c * K0 + H + k] += acc;
      }
      __syncthreads();
    }
  }
  float* out = ws + (size_t)blockIdx.x * L.nparams;
  for (int idx = tid; idx < L.nparams; idx += GJ_THREADS) out[idx] = dpar[idx];
}

}  // namespace

// Deterministic reduction of per-CTA parameter-gradient partials: dparams[p] = sum_cta ws[cta][p].
__global__ void gj_reduce_partials_kernel(const float* __restrict__ ws, int nparts, int n, float* __restrict__ out) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  float acc = 0.f;
  for (int c = 0; c < nparts; ++c) acc += ws[(size_t)c * n + p];
  out[p] = acc;
}

// ------------------------------------------------------------------------------------------------
// host launchers (called from api.cu)
// ------------------------------------------------------------------------------------------------
int gj_num_sms();
void gj_set_error(const char* fmt, ...);

constexpr int kFwdR = 128;
constexpr int kBwdR = 64;

int gj_simt_grid(int batch) {
  int sms = gj_num_sms();
  return batch < sms ? (batch > 0 ? batch : 1) : sms;
}

int gj_mp_fwd_simt(const gj_mp_desc* d, const float* h, const float* params, float* h_out, float* e_out,
                   cudaStream_t stream) {
  MPLayout L; const char* why;
  int rc = gj_fill_arch(d, &L, &why);
  if (rc) { gj_set_error("gj_mp_step_fwd: %s", why); return rc; }
  if (L.B == 0) return GJ_OK;
  int bytes = plan_smem(&L, kFwdR, false);
  if (bytes > 227 * 1024) { gj_set_error("gj_mp_step_fwd(fp32): needs %d B shared memory (> 227 KB)", bytes); return GJ_ERR_SMEM; }
  auto kern = mp_fwd_simt_kernel<kFwdR>;
  cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (ce != cudaSuccess) { gj_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  kern<<<gj_simt_grid(L.B), GJ_THREADS, bytes, stream>>>(L, h, params, h_out, e_out);
  ce = cudaGetLastError();
  if (ce != cudaSuccess) { gj_set_error("mp_fwd_simt launch: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  return GJ_OK;
}

size_t gj_mp_bwd_simt_workspace(const gj_mp_desc* d) {
  MPLayout L; const char* why;
  if (gj_fill_arch(d, &L, &why)) return 0;
  return (size_t)gj_simt_grid(L.B) * L.nparams * sizeof(float);
}

int gj_mp_bwd_simt(const gj_mp_desc* d, const float* h, const float* e, const float* params, const float* dh_out,
                   float* dh, float* dparams, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  MPLayout L; const char* why;
  int rc = gj_fill_arch(d, &L, &why);
  if (rc) { gj_set_error("gj_mp_step_bwd: %s", why); return rc; }
  if (L.B == 0) { cudaMemsetAsync(dparams, 0, (size_t)L.nparams * sizeof(float), stream); return GJ_OK; }
  int bytes = plan_smem(&L, kBwdR, true);
  if (bytes > 227 * 1024) { gj_set_error("gj_mp_step_bwd(fp32): needs %d B shared memory (> 227 KB)", bytes); return GJ_ERR_SMEM; }
  int grid = gj_simt_grid(L.B);
  if (ws_bytes < (size_t)grid * L.nparams * sizeof(float)) { gj_set_error("gj_mp_step_bwd: workspace too small"); return GJ_ERR_WORKSPACE; }
  auto kern = mp_bwd_simt_kernel<kBwdR>;
  cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (ce != cudaSuccess) { gj_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  kern<<<grid, GJ_THREADS, bytes, stream>>>(L, h, e, params, dh_out, dh, (float*)workspace);
  gj_reduce_partials_kernel<<<(L.nparams + 255) / 256, 256, 0, stream>>>((const float*)workspace, grid, L.nparams, dparams);
  ce = cudaGetLastError();
  if (ce != cudaSuccess) { gj_set_error("mp_bwd_simt launch: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  return GJ_OK;
}
