// Fused Chamfer / jet-sum loss, forward + gradient w.r.t. the reconstruction, one CTA per jet.
// Replaces reference utils/losses/chamfer_loss/chamfer_loss.py:11-42 and
// utils/losses/chamfer_loss/distance_sq.py:4-77 (which materialise two (B,N,N,D) repeats).
// HBM-bound: 3*N*D*4 bytes per jet; gradients flow through the arg-mins only (torch.min semantics,
// first index on ties).
#include "gj_common.cuh"

namespace {

__global__ void chamfer_kernel(int np_, int nq, int dim, int mink, int mink_jet, float wc, float wj, const float* __restrict__ p,
                               const float* __restrict__ q, float* __restrict__ jet_terms, float* __restrict__ dp) {
  gj_pdl_sync();
  extern __shared__ float sm[];
  float* sp = sm;                       // [np][4]
  float* sq = sp + np_ * 4;             // [nq][4]
  int* jstar = reinterpret_cast<int*>(sq + nq * 4);   // [np]
  int* istar = jstar + np_;                            // [nq]
  float* red = reinterpret_cast<float*>(istar + nq);   // [blockDim/32 * 2] + [8] jet diff
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* pg = p + (size_t)b * np_ * dim;
  const float* qg = q + (size_t)b * nq * dim;
  for (int idx = tid; idx < np_ * 4; idx += blockDim.x) { int n = idx >> 2, c = idx & 3; sp[idx] = c < dim ? __ldg(pg + n * dim + c) : 0.f; }
  for (int idx = tid; idx < nq * 4; idx += blockDim.x) { int n = idx >> 2, c = idx & 3; sq[idx] = c < dim ? __ldg(qg + n * dim + c) : 0.f; }
  __syncthreads();
  const float s1 = mink ? -1.f : 1.f;       // sign of components 1..3 in the pairwise distances
  const float sj = mink_jet ? -1.f : 1.f;   // ... and in the jet term (chamfer_loss.py:40 passes loss_norm_choice unchanged)
  float local = 0.f;
  if (tid < np_) {          // nearest target for each reconstructed particle
    float4 a = *reinterpret_cast<const float4*>(sp + tid * 4);
    float best = INFINITY; int bj = 0;
    for (int j = 0; j < nq; ++j) {
      float4 c = *reinterpret_cast<const float4*>(sq + j * 4);
      float dx = a.x - c.x, dy = a.y - c.y, dz = a.z - c.z, dw = a.w - c.w;
      float d = dx * dx + s1 * (dy * dy + dz * dz + dw * dw);
      if (d < best) { best = d; bj = j; }
    }
    jstar[tid] = bj; local += best;
  }
  if (tid < nq) {           // nearest reconstructed particle for each target
    float4 c = *reinterpret_cast<const float4*>(sq + tid * 4);
    float best = INFINITY; int bi = 0;
    for (int i = 0; i < np_; ++i) {
      float4 a = *reinterpret_cast<const float4*>(sp + i * 4);
      float dx = a.x - c.x, dy = a.y - c.y, dz = a.z - c.z, dw = a.w - c.w;
      float d = dx * dx + s1 * (dy * dy + dz * dz + dw * dw);
      if (d < best) { best = d; bi = i; }
    }
    istar[tid] = bi; local += best;
  }
  // jet feature difference: sum_i p_i - sum_j q_j, one warp, fixed order
  float* jd = red + 64;
  if (tid < 4) {
    float acc = 0.f;
    for (int i = 0; i < np_; ++i) acc += sp[i * 4 + tid];
    float acc2 = 0.f;
    for (int j = 0; j < nq; ++j) acc2 += sq[j * 4 + tid];
    jd[tid] = acc - acc2;
  }
  float ws = gj_warp_sum(local);
  if ((tid & 31) == 0) red[tid >> 5] = ws;
  __syncthreads();
  if (tid == 0) {
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    float jt = jd[0] * jd[0] + sj * (jd[1] * jd[1] + jd[2] * jd[2] + jd[3] * jd[3]);
    jet_terms[b * 2 + 0] = tot;
    jet_terms[b * 2 + 1] = jt;
  }
  if (dp && tid < np_) {
    const float sgn[4] = {2.f, 2.f * s1, 2.f * s1, 2.f * s1};
    const float sgj[4] = {2.f, 2.f * sj, 2.f * sj, 2.f * sj};
    float g[4];
    const float* a = sp + tid * 4;
    const float* c = sq + jstar[tid] * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) g[k] = a[k] - c[k];
    for (int j = 0; j < nq; ++j)
      if (istar[j] == tid) {
#pragma unroll
        for (int k = 0; k < 4; ++k) g[k] += a[k] - sq[j * 4 + k];
      }
    for (int k = 0; k < dim; ++k) dp[((size_t)b * np_ + tid) * dim + k] = sgn[k] * wc * g[k] + sgj[k] * wj * jd[k];
  }
}

// terms[t] = sum_b jet_terms[b][t], single block, fixed tree => deterministic
__global__ void chamfer_reduce_kernel(int batch, float wc, float wj, const float* __restrict__ jet_terms,
                                      float* __restrict__ terms) {
  gj_pdl_sync();
  __shared__ float red[2][32];
  float a0 = 0.f, a1 = 0.f;
  for (int b = threadIdx.x; b < batch; b += blockDim.x) { a0 += jet_terms[b * 2]; a1 += jet_terms[b * 2 + 1]; }
  a0 = gj_warp_sum(a0); a1 = gj_warp_sum(a1);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a0; red[1][threadIdx.x >> 5] = a1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t0 = 0.f, t1 = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { t0 += red[0][w]; t1 += red[1][w]; }
    terms[0] = t0; terms[1] = t1; terms[2] = wc * t0 + wj * t1;
  }
}

// Per-particle nearest-neighbour distances of the anomaly scores (reference utils/jet_analysis/anomaly_detection.py:
// chamfer :459-488 with the Euclidean norm, chamfer_lorentz :491-510 with E^2 - px^2 - py^2 - pz^2), one CTA per jet:
// min_pq[b][i] = min_j dist(p_i, q_j), min_qp[b][j] = min_i dist(p_i, q_j).  No (B,N,N,D) difference tensor is built.
__global__ void pair_min_dist_kernel(int np_, int nq, int dim, int lorentz, const float* __restrict__ p, const float* __restrict__ q,
                                     float* __restrict__ min_pq, float* __restrict__ min_qp) {
  gj_pdl_sync();
  extern __shared__ float sm[];
  float* sp = sm;                       // [np][4]
  float* sq = sp + np_ * 4;             // [nq][4]
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* pg = p + (size_t)b * np_ * dim;
  const float* qg = q + (size_t)b * nq * dim;
  for (int idx = tid; idx < np_ * 4; idx += blockDim.x) { int n = idx >> 2, c = idx & 3; sp[idx] = c < dim ? __ldg(pg + n * dim + c) : 0.f; }
  for (int idx = tid; idx < nq * 4; idx += blockDim.x) { int n = idx >> 2, c = idx & 3; sq[idx] = c < dim ? __ldg(qg + n * dim + c) : 0.f; }
  __syncthreads();
  const float s1 = lorentz ? -1.f : 1.f;
  for (int i = tid; i < np_; i += blockDim.x) {
    const float4 a = *reinterpret_cast<const float4*>(sp + i * 4);
    float best = INFINITY;
    for (int j = 0; j < nq; ++j) {
      const float4 c = *reinterpret_cast<const float4*>(sq + j * 4);
      const float dx = a.x - c.x, dy = a.y - c.y, dz = a.z - c.z, dw = a.w - c.w;
      best = fminf(best, dx * dx + s1 * (dy * dy + dz * dz + dw * dw));
    }
    min_pq[(size_t)b * np_ + i] = lorentz ? best : sqrtf(best);      // the minimum of the norms is the norm at the minimum square
  }
  for (int j = tid; j < nq; j += blockDim.x) {
    const float4 c = *reinterpret_cast<const float4*>(sq + j * 4);
    float best = INFINITY;
    for (int i = 0; i < np_; ++i) {
      const float4 a = *reinterpret_cast<const float4*>(sp + i * 4);
      const float dx = a.x - c.x, dy = a.y - c.y, dz = a.z - c.z, dw = a.w - c.w;
      best = fminf(best, dx * dx + s1 * (dy * dy + dz * dz + dw * dw));
    }
    min_qp[(size_t)b * nq + j] = lorentz ? best : sqrtf(best);
  }
}

// Optimal assignment between the particles of p and q per jet (the matching step of reference
// utils/jet_analysis/anomaly_detection.py: hungarian :513-547, hungarian_lorentz :550-590, and of
// utils/losses/hungarian_mse/hungarian_mse.py:51-56, which call scipy.optimize.linear_sum_assignment jet by jet on the
// host).  One CTA per jet, thread j owns column j: the O(n^3) shortest-augmenting-path Hungarian method with row / column
// potentials (u, v), where the scan over the columns of every path step -- relax, arg-min, potential update -- runs across
// the threads.  Costs are fp32 (|p_i - q_j|_2 or the Lorentz norm squared), potentials and slacks fp64.
// col_for_row[b][i] = column assigned to row i (linear_sum_assignment's second array for a square matrix).
__global__ void assignment_kernel(int n, int dim, int lorentz, const float* __restrict__ p, const float* __restrict__ q,
                                  int* __restrict__ col_for_row, float* __restrict__ total_cost) {
  gj_pdl_sync();
  extern __shared__ float4 asg_smem[];
  float* sp = reinterpret_cast<float*>(asg_smem);             // [n][4]
  float* sq = sp + n * 4;                                     // [n][4]
  double* u = reinterpret_cast<double*>(sq + n * 4);          // [n + 1] row potentials (1-based, 0 = none)
  double* red_v = u + (n + 1);                                // [32] per-warp minima
  int* red_j = reinterpret_cast<int*>(red_v + 32);            // [32]
  int* prow = red_j + 32;                                     // [n + 1] row matched to column j (0 = free)
  int* way = prow + (n + 1);                                  // [n + 1]
  int* ctl = way + (n + 1);                                   // [2]: current column j0, spare
  float* cost = reinterpret_cast<float*>(ctl + 2);            // [n][n]
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const float* pg = p + (size_t)b * n * dim;
  const float* qg = q + (size_t)b * n * dim;
  for (int idx = tid; idx < n * 4; idx += blockDim.x) {
    const int r = idx >> 2, c = idx & 3;
    sp[idx] = c < dim ? __ldg(pg + r * dim + c) : 0.f;
    sq[idx] = c < dim ? __ldg(qg + r * dim + c) : 0.f;
  }
  for (int idx = tid; idx <= n; idx += blockDim.x) { u[idx] = 0.0; prow[idx] = 0; way[idx] = 0; }
  __syncthreads();
  const float s1 = lorentz ? -1.f : 1.f;
  for (int idx = tid; idx < n * n; idx += blockDim.x) {
    const int i = idx / n, j = idx - i * n;
    const float4 a = *reinterpret_cast<const float4*>(sp + i * 4), c = *reinterpret_cast<const float4*>(sq + j * 4);
    const float dx = a.x - c.x, dy = a.y - c.y, dz = a.z - c.z, dw = a.w - c.w;
    const float d2 = dx * dx + s1 * (dy * dy + dz * dz + dw * dw);
    cost[idx] = lorentz ? d2 : sqrtf(d2);
  }
  __syncthreads();
  const int j = tid + 1;                 // this thread's column (1-based); threads beyond n idle along
  const bool has_col = j <= n;
  double v = 0.0, minv = 0.0;
  bool used = false;
  for (int i = 1; i <= n; ++i) {
    if (tid == 0) { prow[0] = i; ctl[0] = 0; }
    minv = INFINITY; used = false;
    __syncthreads();
    while (true) {
      const int j0 = ctl[0], i0 = prow[j0];
      if (has_col && j == j0) used = true;
      double cand = INFINITY;
      if (has_col && !used) {
        const double cur = (double)cost[(size_t)(i0 - 1) * n + (j - 1)] - u[i0] - v;
        if (cur < minv) { minv = cur; way[j] = j0; }
        cand = minv;
      }
      // block arg-min of cand (ties: smallest column)
      int cj = has_col && !used ? j : 0x7fffffff;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, cand, d);
        const int oj = __shfl_xor_sync(0xffffffffu, cj, d);
        if (ov < cand || (ov == cand && oj < cj)) { cand = ov; cj = oj; }
      }
      if (lane == 0) { red_v[warp] = cand; red_j[warp] = cj; }
      __syncthreads();
      double delta = red_v[0]; int j1 = red_j[0];
      for (int w = 1; w < nwarps; ++w) {
        const double ov = red_v[w]; const int oj = red_j[w];
        if (ov < delta || (ov == delta && oj < j1)) { delta = ov; j1 = oj; }
      }
      // potentials: matched (used) columns move with their rows, free columns get closer by delta
      if (has_col) {
        if (used) { u[prow[j]] += delta; v -= delta; }
        else minv -= delta;
      }
      if (tid == 0) { u[prow[0]] += delta; ctl[0] = j1; }
      __syncthreads();
      if (prow[j1] == 0) break;
    }
    if (tid == 0) {      // augment along the recorded path
      int j0 = ctl[0];
      do { const int jn = way[j0]; prow[j0] = prow[jn]; j0 = jn; } while (j0);
    }
    __syncthreads();
  }
  if (has_col) col_for_row[(size_t)b * n + (prow[j] - 1)] = j - 1;
  if (total_cost) {
    float c = has_col ? cost[(size_t)(prow[j] - 1) * n + (j - 1)] : 0.f;
    c = gj_warp_sum(c);
    __syncthreads();
    if (lane == 0) red_v[warp] = (double)c;
    __syncthreads();
    if (tid == 0) { double t = 0.0; for (int w = 0; w < nwarps; ++w) t += red_v[w]; total_cost[b] = (float)t; }
  }
}

}  // namespace

void gj_set_error(const char* fmt, ...);

static size_t assignment_smem(int n) {
  return (size_t)(n + 1 + 32) * sizeof(double) + (size_t)(32 + 2 * (n + 1) + 2) * sizeof(int) + ((size_t)n * n + 8 * (size_t)n) * sizeof(float) + 16;
}

int gj_assignment_launch(int batch, int n, int dim, int lorentz, const float* p, const float* q, int* col_for_row, float* total_cost,
                         cudaStream_t stream) {
  if (batch < 0 || n < 1 || n > 1024 || assignment_smem(n) > 200 * 1024) {
    gj_set_error("gj_assignment: 1 <= particles per jet <= 220 (cost matrix in shared memory), got %d", n); return GJ_ERR_INVALID; }
  if (dim < 1 || dim > 4 || (lorentz && dim != 4)) { gj_set_error("gj_assignment: 1..4 components (4 for the Lorentz norm), got %d", dim); return GJ_ERR_INVALID; }
  if (batch == 0) return GJ_OK;
  const size_t smem = assignment_smem(n);
  cudaError_t ce = cudaSuccess;
  if (smem > 48 * 1024) ce = cudaFuncSetAttribute(assignment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (ce == cudaSuccess) {
    gj_launch(assignment_kernel, batch, gj_round_up(n, 32), smem, stream, n, dim, lorentz, p, q, col_for_row, total_cost);
    ce = cudaGetLastError();
  }
  if (ce != cudaSuccess) { gj_set_error("assignment launch: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  return GJ_OK;
}

int gj_pair_min_dist_launch(int batch, int np_, int nq, int dim, int lorentz, const float* p, const float* q, float* min_pq,
                            float* min_qp, cudaStream_t stream) {
  if (batch < 0 || np_ < 1 || nq < 1 || (size_t)(np_ + nq) * 16 > 200 * 1024) {
    gj_set_error("gj_pair_min_dist: particle counts out of range"); return GJ_ERR_INVALID; }
  if (dim < 1 || dim > 4 || (lorentz && dim != 4)) { gj_set_error("gj_pair_min_dist: 1..4 components (4 for the Lorentz norm), got %d", dim); return GJ_ERR_INVALID; }
  if (batch == 0) return GJ_OK;
  const int nmax = np_ > nq ? np_ : nq;
  int threads = gj_round_up(nmax, 32); if (threads > 256) threads = 256;
  const size_t smem = (size_t)(np_ + nq) * 4 * sizeof(float);
  cudaError_t ce = cudaSuccess;
  if (smem > 48 * 1024) ce = cudaFuncSetAttribute(pair_min_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (ce == cudaSuccess) {
    gj_launch(pair_min_dist_kernel, batch, threads, smem, stream, np_, nq, dim, lorentz, p, q, min_pq, min_qp);
    ce = cudaGetLastError();
  }
  if (ce != cudaSuccess) { gj_set_error("pair_min_dist launch: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  return GJ_OK;
}

int gj_chamfer_launch(int batch, int np_, int nq, int dim, int norm, float wc, float wj, const float* p, const float* q,
                      float* jet_terms, float* terms, float* dp, cudaStream_t stream) {
  if (batch < 0 || np_ < 1 || nq < 1 || np_ > 1024 || nq > 1024) { gj_set_error("gj_chamfer_fwd_bwd: particle counts must be in [1,1024]"); return GJ_ERR_INVALID; }
  if (dim != 3 && dim != 4) { gj_set_error("gj_chamfer_fwd_bwd: p and q must be 3- or 4-vectors (got %d)", dim); return GJ_ERR_INVALID; }
  if (batch == 0) { cudaMemsetAsync(terms, 0, 3 * sizeof(float), stream); return GJ_OK; }
  // distance_sq.py:43-44 forces cartesian for 3-vectors inside pairwise_distance_sq ONLY; the jet term of chamfer_loss.py:40
  // calls normsq with the loss's norm choice directly, so 3-vectors with 'minkowskian' / 'polar' get p0^2 - p1^2 - p2^2 there
  const int mink = (dim != 3 && norm != 0) ? 1 : 0, mink_jet = norm != 0 ? 1 : 0;
  int nmax = np_ > nq ? np_ : nq;
  int threads = gj_round_up(nmax, 32);
  if (threads < 32) threads = 32;
  size_t smem = (size_t)(np_ + nq) * 4 * sizeof(float) + (size_t)(np_ + nq) * sizeof(int) + (64 + 8) * sizeof(float);
  gj_launch(chamfer_kernel, batch, threads, smem, stream, np_, nq, dim, mink, mink_jet, wc, wj, p, q, jet_terms, dp);
  gj_launch(chamfer_reduce_kernel, 1, 1024, 0, stream, batch, wc, wj, jet_terms, terms);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) { gj_set_error("chamfer launch: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  return GJ_OK;
}
