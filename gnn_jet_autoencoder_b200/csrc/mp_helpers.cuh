// Device helpers shared by the fp32 (SIMT) and bf16 (tcgen05) message-passing kernels.
// Every helper takes the caller's (tid, nthr) so it can run on a whole CTA or on one warpgroup.
#pragma once
#include "gj_common.cuh"

// fp32 weights that stay on the CUDA cores in both precision modes: factorised first edge layer
// (Wa | Wb | wd of edge_net[t][0], graphnet.py:220 column order), all edge biases, the node MLP.
__device__ inline void gj_stage_small_weights(const MPLayout& L, const float* __restrict__ params, float* sm, int tid, int nthr) {
  for (int l = 0; l < L.Le; ++l)
    for (int c = tid; c < L.Ep[l]; c += nthr) sm[L.o_bE[l] + c] = c < L.E[l] ? __ldg(params + L.pb[l] + c) : 0.f;
  {
    const int H = L.H, K0 = L.K[0], E0 = L.E[0];
    for (int idx = tid; idx < L.E0p * L.Hs; idx += nthr) {
      int c = idx / L.Hs, k = idx - c * L.Hs;
      bool ok = c < E0 && k < H;
      sm[L.o_wa + idx] = ok ? __ldg(params + L.pW[0] + c * K0 + k) : 0.f;
      sm[L.o_wb + idx] = ok ? __ldg(params + L.pW[0] + c * K0 + H + k) : 0.f;
    }
    for (int c = tid; c < L.E0p; c += nthr) sm[L.o_wd + c] = c < E0 ? __ldg(params + L.pW[0] + c * K0 + 2 * H) : 0.f;
  }
  for (int m = 0; m < L.Ln; ++m) {
    const int I = L.I[m], Is = L.Is[m];
    for (int idx = tid; idx < L.O[m] * Is; idx += nthr) {
      int o = idx / Is, k = idx - o * Is;
      sm[L.o_V[m] + idx] = k < I ? __ldg(params + L.pV[m] + o * I + k) : 0.f;
    }
    for (int o = tid; o < L.O[m]; o += nthr) sm[L.o_c[m] + o] = __ldg(params + L.pc[m] + o);
  }
}

// fp32 edge weights of layers >= 1, natural (out,in) layout zero-padded to [Ep][Kp].
__device__ inline void gj_stage_edge_weights_f32(const MPLayout& L, const float* __restrict__ params, float* sm, int tid, int nthr) {
  for (int l = 1; l < L.Le; ++l) {
    float* w = sm + L.o_wE[l];
    const int Kp = L.Kp[l], K = L.K[l], E = L.E[l];
    for (int idx = tid; idx < L.Ep[l] * Kp; idx += nthr) {
      int c = idx / Kp, k = idx - c * Kp;
      w[idx] = (c < E && k < K) ? __ldg(params + L.pW[l] + c * K + k) : 0.f;
    }
  }
}

// out[n][c] = bias[c] + sum_k W[c][k] hrows[n][k]   (Q_j = Wb h_j with bias == nullptr, P_i = Wa h_i + b0)
__device__ inline void gj_node_project(const MPLayout& L, const float* W, const float* bias, const float* hrows, int nrows,
                                       float* out, int tid, int nthr) {
  for (int idx = tid; idx < nrows * L.E0p; idx += nthr) {
    int n = idx / L.E0p, c = idx - n * L.E0p;
    float acc = bias ? bias[c] : 0.f;
    const float* w = W + c * L.Hs;
    const float* x = hrows + n * L.Hs;
    for (int k = 0; k < L.H; ++k) acc = fmaf(w[k], x[k], acc);
    out[n * L.E0s + c] = acc;
  }
}

// Node MLP layer m: out[n][o] = leaky(c[o] + sum_k in[n][k] V[o][k]) for n < nrows (graphnet.py:266-268).
__device__ inline void gj_node_layer_fwd(const MPLayout& L, const float* sm, int m, const float* in, float* out, int nrows,
                                         int tid, int nthr) {
  const int O = L.O[m], I = L.I[m], Is = L.Is[m];
  const float* V = sm + L.o_V[m];
  const float* cb = sm + L.o_c[m];
  for (int idx = tid; idx < nrows * O; idx += nthr) {
    int n = idx / O, o = idx - n * O;
    float acc = cb[o];
    const float* v = V + o * Is;
    const float* x = in + n * L.Ws;
    for (int k = 0; k < I; ++k) acc = fmaf(v[k], x[k], acc);
    out[n * L.Ws + o] = gj_leaky(acc, L.alpha);
  }
}
