// Flat fused Adam (+L1/L2 regulariser gradients), parameter norms, and the 'mean' latent map.
// Replaces torch.optim.Adam as built in reference utils/initialize.py:152-153 (about 100 tiny foreach
// kernels per step), the regularisers of utils/train.py:376-384 / models/encoder.py:173-179, and
// models/encoder.py:147-149.  All HBM-bound elementwise / reduction work.
#include "gj_common.cuh"

namespace {

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, float lr_over_bc1, float inv_sqrt_bc2, float b1, float b2,
                            float eps, float gscale, float l1, float l2) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float pi = p[i];
    float gi = g[i] * gscale;
    if (l1 != 0.f) gi += l1 * (pi > 0.f ? 1.f : (pi < 0.f ? -1.f : 0.f));
    if (l2 != 0.f) gi += 2.f * l2 * pi;
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] = pi - lr_over_bc1 * (mi / denom);
  }
}

__global__ void norms_stage1(const float* __restrict__ p, size_t n, float* __restrict__ part) {
  __shared__ float red[2][8];
  float a = 0.f, s = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float x = p[i]; a += fabsf(x); s = fmaf(x, x, s);
  }
  a = gj_warp_sum(a); s = gj_warp_sum(s);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = s; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t0 = 0.f, t1 = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { t0 += red[0][w]; t1 += red[1][w]; }
    part[blockIdx.x * 2] = t0; part[blockIdx.x * 2 + 1] = t1;
  }
}

__global__ void norms_stage2(const float* __restrict__ part, int nblk, float* __restrict__ out) {
  if (threadIdx.x == 0) {
    float t0 = 0.f, t1 = 0.f;
    for (int b = 0; b < nblk; ++b) { t0 += part[b * 2]; t1 += part[b * 2 + 1]; }
    out[0] = t0; out[1] = t1;
  }
}

__global__ void latent_mean_fwd_kernel(int N, int W, const float* __restrict__ y, float* __restrict__ z, int total) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int b = idx / W, c = idx - b * W;
  const float* yp = y + (size_t)b * N * W + c;
  float acc = 0.f;
  for (int n = 0; n < N; ++n) acc += yp[(size_t)n * W];
  z[idx] = acc / (float)N;
}

__global__ void latent_mean_bwd_kernel(int N, int W, const float* __restrict__ dz, float* __restrict__ dy, size_t total) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  size_t b = idx / ((size_t)N * W);
  int c = (int)(idx % W);
  dy[idx] = dz[b * W + c] / (float)N;
}

}  // namespace

void gj_set_error(const char* fmt, ...);
static const int kNormBlocks = 64;

int gj_adam_launch(float* param, const float* grad, float* m, float* v, size_t n, float lr, float b1, float b2, float eps,
                   int step, float gscale, float l1, float l2, cudaStream_t stream) {
  if (n == 0) return GJ_OK;
  if (step < 1) { gj_set_error("gj_adam_step_flat: step must be >= 1"); return GJ_ERR_INVALID; }
  double bc1 = 1.0 - pow((double)b1, step), bc2 = 1.0 - pow((double)b2, step);
  int blocks = (int)((n + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
  adam_kernel<<<blocks, 256, 0, stream>>>(param, grad, m, v, n, (float)(lr / bc1), (float)(1.0 / sqrt(bc2)), b1, b2, eps,
                                          gscale, l1, l2);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) { gj_set_error("adam launch: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  return GJ_OK;
}

size_t gj_norms_ws_bytes(size_t) { return kNormBlocks * 2 * sizeof(float); }

int gj_norms_launch(const float* p, size_t n, float* out, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (ws_bytes < gj_norms_ws_bytes(n)) { gj_set_error("gj_param_norms: workspace too small"); return GJ_ERR_WORKSPACE; }
  norms_stage1<<<kNormBlocks, 256, 0, stream>>>(p, n, (float*)ws);
  norms_stage2<<<1, 32, 0, stream>>>((const float*)ws, kNormBlocks, out);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) { gj_set_error("norms launch: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  return GJ_OK;
}

int gj_latent_mean_fwd_launch(int B, int N, int W, const float* y, float* z, cudaStream_t stream) {
  int total = B * W;
  if (total == 0) return GJ_OK;
  latent_mean_fwd_kernel<<<(total + 255) / 256, 256, 0, stream>>>(N, W, y, z, total);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) { gj_set_error("latent_mean_fwd launch: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  return GJ_OK;
}

int gj_latent_mean_bwd_launch(int B, int N, int W, const float* dz, float* dy, cudaStream_t stream) {
  size_t total = (size_t)B * N * W;
  if (total == 0) return GJ_OK;
  latent_mean_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(N, W, dz, dy, total);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) { gj_set_error("latent_mean_bwd launch: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  return GJ_OK;
}

// ------------------------------------------------------------------------------------------------
// small dense layers at node / graph level (decoder.py:127-136, encoder.py:156-161)
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int kLinRowsPerBlock = 128;

// y[r][o] = b[o] + sum_k x[r][k] w[o][k]; one thread per output, o fastest (coalesced store, broadcast x).
__global__ void linear_fwd_kernel(int rows, int K, int O, const float* __restrict__ x, const float* __restrict__ w,
                                  const float* __restrict__ b, float* __restrict__ y) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)rows * O) return;
  int r = (int)(idx / O), o = (int)(idx - (size_t)r * O);
  float acc = b ? __ldg(b + o) : 0.f;
  const float* xr = x + (size_t)r * K;
  const float* wo = w + (size_t)o * K;
  for (int k = 0; k < K; ++k) acc = fmaf(__ldg(xr + k), __ldg(wo + k), acc);
  y[idx] = acc;
}

// dx[r][k] = sum_o dy[r][o] w[o][k]
__global__ void linear_dx_kernel(int rows, int K, int O, const float* __restrict__ dy, const float* __restrict__ w,
                                 float* __restrict__ dx) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)rows * K) return;
  int r = (int)(idx / K), k = (int)(idx - (size_t)r * K);
  const float* g = dy + (size_t)r * O;
  float acc = 0.f;
  for (int o = 0; o < O; ++o) acc = fmaf(__ldg(g + o), __ldg(w + (size_t)o * K + k), acc);
  dx[idx] = acc;
}

// Per row chunk: part[chunk][k*O + o] = sum_{r in chunk} dy[r][o] x[r][k]; slot k == K holds the bias gradient.
__global__ void linear_dw_partial_kernel(int rows, int K, int O, const float* __restrict__ x, const float* __restrict__ dy,
                                         float* __restrict__ part) {
  const int r0 = blockIdx.y * kLinRowsPerBlock;
  const int r1 = min(rows, r0 + kLinRowsPerBlock);
  const int total = (K + 1) * O;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int k = idx / O, o = idx - k * O;
    float acc = 0.f;
    if (k < K) { for (int r = r0; r < r1; ++r) acc = fmaf(__ldg(dy + (size_t)r * O + o), __ldg(x + (size_t)r * K + k), acc); }
    else { for (int r = r0; r < r1; ++r) acc += __ldg(dy + (size_t)r * O + o); }
    part[(size_t)blockIdx.y * total + idx] = acc;
  }
}

__global__ void linear_dw_reduce_kernel(int nchunks, int K, int O, const float* __restrict__ part, float* __restrict__ dw,
                                        float* __restrict__ db) {
  const int total = (K + 1) * O;
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  float acc = 0.f;
  for (int c = 0; c < nchunks; ++c) acc += part[(size_t)c * total + idx];
  int k = idx / O, o = idx - k * O;
  if (k < K) dw[(size_t)o * K + k] = acc;
  else if (db) db[o] = acc;
}

}  // namespace

int gj_linear_fwd_launch(int rows, int K, int O, const float* x, const float* w, const float* b, float* y, cudaStream_t stream) {
  size_t total = (size_t)rows * O;
  if (total == 0) return GJ_OK;
  linear_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(rows, K, O, x, w, b, y);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) { gj_set_error("linear_fwd launch: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  return GJ_OK;
}

static int lin_chunks(int rows) { int c = (rows + kLinRowsPerBlock - 1) / kLinRowsPerBlock; return c < 1 ? 1 : c; }

size_t gj_linear_bwd_ws_bytes(int rows, int K, int O) { return (size_t)lin_chunks(rows) * (K + 1) * O * sizeof(float); }

int gj_linear_bwd_launch(int rows, int K, int O, const float* x, const float* w, const float* dy, float* dx, float* dw,
                         float* db, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (ws_bytes < gj_linear_bwd_ws_bytes(rows, K, O)) { gj_set_error("gj_linear_bwd: workspace too small"); return GJ_ERR_WORKSPACE; }
  const int total = (K + 1) * O;
  const int nch = rows > 0 ? lin_chunks(rows) : 0;
  if (rows > 0) {
    if (dx) linear_dx_kernel<<<(unsigned)(((size_t)rows * K + 255) / 256), 256, 0, stream>>>(rows, K, O, dy, w, dx);
    dim3 grid((total + 255) / 256, nch);
    linear_dw_partial_kernel<<<grid, 256, 0, stream>>>(rows, K, O, x, dy, (float*)ws);
  }
  linear_dw_reduce_kernel<<<(total + 255) / 256, 256, 0, stream>>>(nch, K, O, (const float*)ws, dw, db);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) { gj_set_error("linear_bwd launch: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  return GJ_OK;
}
