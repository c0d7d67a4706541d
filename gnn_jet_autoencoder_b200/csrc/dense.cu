// Node-level dense layers of ANY width (reference models/graphnet.py:249-271 `_aggregate`, and the per-node half of the
// factorised first edge layer, graphnet.py:220): P|Q projections, node MLP and their adjoints as a sequence of generic
// GEMMs over all B*N node rows.  Used where the fused node kernels of node_kernels.cu / node_tc.cu do not cover the
// widths: BASELINE config 5's node_sizes [[H]] / edge_sizes [[H, H]] with H = 64 .. 256 (a 256 x 513 weight matrix does
// not fit shared memory; here weights are streamed through shared memory in k-chunks like any GEMM operand).
//
// One GEMM primitive, two implementations:
//   * GJ_PREC_BF16: tcgen05.mma (bf16 operands converted on the fly from the fp32 tensors in HBM, fp32 accumulation in TMEM),
//     128 x NT output tile per CTA, 64-deep k-chunks double buffered in shared memory in the UMMA SWIZZLE_NONE slab
//     layout, several CTAs per SM;
//   * GJ_PREC_FP32: register-tiled FFMA (64 x 64 x 16 tiles), fp32 everywhere (the <= 1e-5 mode).
// C(m, n) = epilogue( sum_k A(m, k) B(n, k) ), operands addressed by two strides each, so the three products of a dense layer
//   forward  y = x W^T      : A = x (k contiguous), B = W (k contiguous)
//   dgrad    dx = g W       : A = g (k contiguous), B = W (n contiguous)
//   wgrad    dW = g^T x     : A = g (m contiguous), B = x (n contiguous); split over the rows, fixed-order reduction
// are the same kernel.  Epilogue: + bias, + previous C (accumulate), LeakyReLU, or multiplication by leaky'(aux).
#include "tc2_common.cuh"

namespace {
using namespace tc2;

struct GemmP {
  const float* A; long long sAm, sAk; int Am, Ak;      // A(m, k) = A[m sAm + k sAk] for m < Am and k < Ak, else 0
  const float* B; long long sBn, sBk; int Bn, Bk;      // B(n, k) likewise
  int M, N, K;
  const float* bias;      // [N] or null
  int act;                // 0 none, 1 leaky(alpha), 2 multiply by leaky'(aux(m, n)) = (aux > 0 ? 1 : alpha)
  float alpha;
  const float* aux; long long ldaux;
  int accumulate;         // v += C_old before the activation
  float* C; long long ldc; int Cm, Cn;      // only m < Cm, n < Cn are stored
  int kchunk;             // split over k: blockIdx.z covers k in [z kchunk, (z + 1) kchunk) and writes C + z part_stride
  long long part_stride;
};

__device__ __forceinline__ float ldA(const GemmP& P, int m, int k) { return (m < P.Am && k < P.Ak) ? __ldg(P.A + m * P.sAm + k * P.sAk) : 0.f; }
__device__ __forceinline__ float ldB(const GemmP& P, int n, int k) { return (n < P.Bn && k < P.Bk) ? __ldg(P.B + n * P.sBn + k * P.sBk) : 0.f; }

__device__ __forceinline__ void epilogue_store(const GemmP& P, float* __restrict__ C, int m, int n, float v) {
  if (m >= P.Cm || n >= P.Cn) return;
  if (P.bias) v += __ldg(P.bias + n);
  float* dst = C + m * P.ldc + n;
  if (P.accumulate) v += *dst;
  if (P.act == 1) v = v > 0.f ? v : P.alpha * v;
  else if (P.act == 2) v *= (__ldg(P.aux + m * P.ldaux + n) > 0.f ? 1.f : P.alpha);
  *dst = v;
}

// ---------------------------------------------------------------------------------------------------------------------
// fp32: 64 x 64 output tile, 16-deep k steps, 256 threads x (4 x 4) register tile
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmP P) {
  __shared__ __align__(16) float As[16][64 + 4];
  __shared__ __align__(16) float Bs[16][64 + 4];
  const int tid = threadIdx.x, m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const int k_lo = blockIdx.z * P.kchunk, k_hi = min(P.K, k_lo + P.kchunk);
  float* C = P.C + blockIdx.z * P.part_stride;
  const int tm = (tid >> 4) * 4, tn = (tid & 15) * 4;
  // two-level accumulation (256 k per inner sum): keeps the rounding error of the long row contractions of the weight
  // gradients (k = thousands of rows) at the level of a pairwise sum
  float acc[4][4], tot[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; tot[i][j] = 0.f; }
  int flush = 0;
  const bool a_kc = P.sAk == 1, b_kc = P.sBk == 1;      // k contiguous in memory: consecutive threads walk k
  for (int k0 = k_lo; k0 < k_hi; k0 += 16) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = it * 256 + tid;
      const int am = a_kc ? idx >> 4 : idx & 63, ak = a_kc ? idx & 15 : idx >> 6;
      As[ak][am] = (k0 + ak < k_hi) ? ldA(P, m0 + am, k0 + ak) : 0.f;
      const int bn = b_kc ? idx >> 4 : idx & 63, bk = b_kc ? idx & 15 : idx >> 6;
      Bs[bk][bn] = (k0 + bk < k_hi) ? ldB(P, n0 + bn, k0 + bk) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][tm]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tn]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
    if (++flush == 16) {
      flush = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { tot[i][j] += acc[i][j]; acc[i][j] = 0.f; }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) epilogue_store(P, C, m0 + tm + i, n0 + tn + j, tot[i][j] + acc[i][j]);
}

// ---------------------------------------------------------------------------------------------------------------------
// bf16 tensor cores: 128 x NT output tile per CTA (thread = output row = TMEM lane), 64-deep k-chunks, two stages
// ---------------------------------------------------------------------------------------------------------------------
constexpr int TCG_KC = 64;
constexpr int TCG_PAD = 16;      // bytes added to every slab stride: the 8 slabs a warp store touches fall into distinct banks

// Operand tile (T rows of the MN dimension x 64 k) -> bf16 slabs.
//   k contiguous in memory (K-major operand):  element (mn, k) at (k / 8) SK + mn 16 + (k % 8) 2,  SK = T 16 + pad
//   mn contiguous in memory (MN-major operand): element (mn, k) at (mn / 8) SM + k 16 + (mn % 8) 2, SM = 64 16 + pad
// Loads are issued in batches of four units (eight 16-byte loads in flight per thread) before anything is converted: a
// load -> convert -> store loop waits for every load in turn (measured: 14 us per 64 KB chunk).
__device__ __forceinline__ uint4 tcg_pack8(const float4& a, const float4& b) {
  return make_uint4(bf2_as_u32(__floats2bfloat162_rn(a.x, a.y)), bf2_as_u32(__floats2bfloat162_rn(a.z, a.w)),
                    bf2_as_u32(__floats2bfloat162_rn(b.x, b.y)), bf2_as_u32(__floats2bfloat162_rn(b.z, b.w)));
}
__device__ __forceinline__ void stage_tile(uint8_t* dst, const float* __restrict__ p, long long s_mn, long long s_k, int lim_mn, int lim_k,
                                           int mn0, int k0, int k_hi, int T, bool kmajor, int tid) {
  const int klim = min(lim_k, k_hi);
  const int SK = T * 16 + TCG_PAD, SM = TCG_KC * 16 + TCG_PAD, T8 = T >> 3;
  const int total = T * 8;      // units of 8 elements: K-major (mn, k8), MN-major (k, m8)
  // element (i, c) of a unit at p[i * s_row + (c0 + c) * s_col]: i = the strided index, c = the 8 contiguous ones
  const long long s_row = kmajor ? s_mn : s_k, s_col = kmajor ? s_k : s_mn;
  const int lim_row = kmajor ? lim_mn : klim, lim_col = kmajor ? klim : lim_mn;
  const int row0 = kmajor ? mn0 : k0, col0 = kmajor ? k0 : mn0;
  const int cdiv = kmajor ? 8 : T8;      // units per row
  const bool vec = s_col == 1 && (s_row & 3) == 0 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0) && ((col0 & 3) == 0);
  for (int u0 = tid; u0 < total; u0 += 4 * 128) {
    float4 a[4], b[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int u = u0 + q * 128;
      a[q] = make_float4(0.f, 0.f, 0.f, 0.f); b[q] = a[q];
      if (u < total) {
        const int i = u / cdiv, c8 = u - i * cdiv, gi = row0 + i, gc = col0 + 8 * c8;
        if (gi < lim_row) {
          const float* src = p + gi * s_row + gc * s_col;
          if (vec && gc + 7 < lim_col) { a[q] = __ldg(reinterpret_cast<const float4*>(src)); b[q] = __ldg(reinterpret_cast<const float4*>(src) + 1); }
          else {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = gc + e < lim_col ? __ldg(src + e * s_col) : 0.f;
            a[q] = make_float4(v[0], v[1], v[2], v[3]); b[q] = make_float4(v[4], v[5], v[6], v[7]);
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int u = u0 + q * 128;
      if (u < total) {
        const int i = u / cdiv, c8 = u - i * cdiv;
        *reinterpret_cast<uint4*>(dst + (kmajor ? c8 * SK + i * 16 : c8 * SM + i * 16)) = tcg_pack8(a[q], b[q]);
      }
    }
  }
}

__host__ __device__ inline int tcg_stage_bytes(int NT) { return (8 * (128 * 16 + TCG_PAD) > 16 * (TCG_KC * 16 + TCG_PAD) ? 8 * (128 * 16 + TCG_PAD) : 16 * (TCG_KC * 16 + TCG_PAD)) +
                                                                (8 * (NT * 16 + TCG_PAD) > (NT / 8) * (TCG_KC * 16 + TCG_PAD) ? 8 * (NT * 16 + TCG_PAD) : (NT / 8) * (TCG_KC * 16 + TCG_PAD)); }
__host__ __device__ inline int tcg_a_bytes() { return 8 * (128 * 16 + TCG_PAD) > 16 * (TCG_KC * 16 + TCG_PAD) ? 8 * (128 * 16 + TCG_PAD) : 16 * (TCG_KC * 16 + TCG_PAD); }

__global__ void __launch_bounds__(128) gemm_tc_kernel(const GemmP P, int NT, int tmem_cols) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = (int)uni((uint32_t)(tid >> 5));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);      // [0], [1]: stage free; [2]: accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 64);
  const int stage_bytes = (tcg_stage_bytes(NT) + 127) & ~127, a_bytes = (tcg_a_bytes() + 127) & ~127;
  uint8_t* stage0 = smem + 128;
  if (tid == 0) { mbar_init(bars, 1); mbar_init(bars + 1, 1); mbar_init(bars + 2, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, (uint32_t)tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * NT;
  const int k_lo = blockIdx.z * P.kchunk, k_hi = min(P.K, k_lo + P.kchunk);
  float* C = P.C + blockIdx.z * P.part_stride;
  const bool a_km = !(P.sAm == 1 && P.sAk != 1), b_km = !(P.sBn == 1 && P.sBk != 1);
  const uint32_t idesc = make_idesc_bf16(128, NT, a_km ? 0 : 1, b_km ? 0 : 1);
  const int nchunks = (k_hi - k_lo + TCG_KC - 1) / TCG_KC;
  uint32_t ph[2] = {0u, 0u};
  for (int c = 0; c < nchunks; ++c) {
    const int s = c & 1, k0 = k_lo + c * TCG_KC;
    uint8_t* sa = stage0 + s * stage_bytes;
    uint8_t* sb = sa + a_bytes;
    if (c >= 2) { mbar_wait(bars + s, ph[s]); ph[s] ^= 1u; }      // the MMAs that read this stage two chunks ago have completed
    stage_tile(sa, P.A, P.sAm, P.sAk, P.Am, P.Ak, m0, k0, k_hi, 128, a_km, tid);
    stage_tile(sb, P.B, P.sBn, P.sBk, P.Bn, P.Bk, n0, k0, k_hi, NT, b_km, tid);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      const uint32_t aa = smem_u32(sa), ba = smem_u32(sb);
      const uint64_t dA = a_km ? make_smem_desc(aa, 128 * 16 + TCG_PAD, 128) : make_smem_desc(aa, 128, TCG_KC * 16 + TCG_PAD);
      const uint64_t dB = b_km ? make_smem_desc(ba, (uint32_t)NT * 16 + TCG_PAD, 128) : make_smem_desc(ba, 128, TCG_KC * 16 + TCG_PAD);
      const uint32_t stepA = a_km ? (uint32_t)(2 * (128 * 16 + TCG_PAD)) >> 4 : 16u, stepB = b_km ? (uint32_t)(2 * (NT * 16 + TCG_PAD)) >> 4 : 16u;
#pragma unroll
      for (int ks = 0; ks < TCG_KC / 16; ++ks)
        mma_bf16_ss_elect(tmem_base, dA + (uint64_t)(stepA * ks), dB + (uint64_t)(stepB * ks), idesc, (c > 0 || ks > 0) ? 1u : 0u);
      mma_commit_elect(bars + s);
      if (c == nchunks - 1) mma_commit_elect(bars + 2);
    }
  }
  if (nchunks > 0) { mbar_wait(bars + 2, 0u); tc_fence_after(); }
  // ---- epilogue: TMEM (thread = row) -> per-warp shared-memory transpose -> coalesced 16-byte global accesses (a warp instruction
  //      covers eight 64-byte row segments; a thread walking its own row would touch 32 lines per store) ----
  const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
  float* scr = reinterpret_cast<float*>(stage0) + warp * (16 * 33);      // [16 columns][33]; the operand stages are free: every MMA has completed
  const bool vec = (P.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0 && (n0 & 3) == 0 &&
                   (!P.aux || ((P.ldaux & 3) == 0 && (reinterpret_cast<uintptr_t>(P.aux) & 15) == 0)) &&
                   (!P.bias || (reinterpret_cast<uintptr_t>(P.bias) & 15) == 0);
  const int rsub = lane >> 2, cq = (lane & 3) * 4;
  for (int c0 = 0; c0 < NT; c0 += 16) {
    uint32_t v[16];
    if (nchunks > 0) { tmem_ld16_u(lane_base + c0, v); tmem_ld_wait(); tmem_pin16(v); }
    else {
#pragma unroll
      for (int q = 0; q < 16; ++q) v[q] = 0u;
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) scr[q * 33 + lane] = __uint_as_float(v[q]);
    __syncwarp();
    const int n = n0 + c0 + cq;
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec && P.bias && n + 3 < P.Cn) bias4 = __ldg(reinterpret_cast<const float4*>(P.bias + n));
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int r = it * 8 + rsub, m = m0 + warp * 32 + r;
      float4 o = make_float4(scr[cq * 33 + r], scr[(cq + 1) * 33 + r], scr[(cq + 2) * 33 + r], scr[(cq + 3) * 33 + r]);
      if (vec && n + 3 < P.Cn) {
        if (m < P.Cm) {
          float4* dst = reinterpret_cast<float4*>(C + m * P.ldc + n);
          o.x += bias4.x; o.y += bias4.y; o.z += bias4.z; o.w += bias4.w;
          if (P.accumulate) { const float4 c = *dst; o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w; }
          if (P.act == 1) {
            o.x = o.x > 0.f ? o.x : P.alpha * o.x; o.y = o.y > 0.f ? o.y : P.alpha * o.y;
            o.z = o.z > 0.f ? o.z : P.alpha * o.z; o.w = o.w > 0.f ? o.w : P.alpha * o.w;
          } else if (P.act == 2) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(P.aux + m * P.ldaux + n));
            o.x *= a.x > 0.f ? 1.f : P.alpha; o.y *= a.y > 0.f ? 1.f : P.alpha; o.z *= a.z > 0.f ? 1.f : P.alpha; o.w *= a.w > 0.f ? 1.f : P.alpha;
          }
          *dst = o;
        }
      } else {
        epilogue_store(P, C, m, n, o.x); epilogue_store(P, C, m, n + 1, o.y);
        epilogue_store(P, C, m, n + 2, o.z); epilogue_store(P, C, m, n + 3, o.w);
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
}

// out[m ldo + n] = sum_z part[z][m N + n]   (fixed order)
__global__ void __launch_bounds__(256) reduce_split_kernel(const float* __restrict__ part, int nz, int M, int N, float* __restrict__ out, long long ldo,
                                                           int accumulate) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)M * N) return;
  float s = 0.f;
  for (int z = 0; z < nz; ++z) s += part[(long long)z * M * N + idx];
  const int m = (int)(idx / N), n = (int)(idx - (long long)m * N);
  out[m * ldo + n] = accumulate ? out[m * ldo + n] + s : s;
}

// The same for up to RM_MAX reductions in ONE launch (the node-level backward of a step queues its weight-gradient and bias
// reductions and flushes them once: they are latency-bound launches of a few KB each).  blk0[i] = first block of reduction i.
constexpr int RM_MAX = 12;
struct ReduceJob { const float* part; float* out; long long ldo; int nz, M, N, accumulate; };
struct ReduceList { ReduceJob job[RM_MAX]; int blk0[RM_MAX + 1]; int n; };
__global__ void __launch_bounds__(256) reduce_multi_kernel(const ReduceList R) {
  int i = 0;
  while (i + 1 < R.n && (int)blockIdx.x >= R.blk0[i + 1]) ++i;
  const ReduceJob J = R.job[i];
  const long long idx = (long long)(blockIdx.x - R.blk0[i]) * 256 + threadIdx.x;
  if (idx >= (long long)J.M * J.N) return;
  float s = 0.f;
  for (int z = 0; z < J.nz; ++z) s += J.part[(long long)z * J.M * J.N + idx];
  const int m = (int)(idx / J.N), n = (int)(idx - (long long)m * J.N);
  J.out[m * J.ldo + n] = J.accumulate ? J.out[m * J.ldo + n] + s : s;
}

// column sums: part[z][n] = sum over the rows of slice z of g[r ld + n]
// (optionally weighted by weight[r])
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ g, long long ld, int rows, int n, int rows_per,
                                                     const float* __restrict__ weight, float* __restrict__ part) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), ry = threadIdx.x >> 5;
  const int r0 = blockIdx.y * rows_per, r1 = min(rows, r0 + rows_per);
  __shared__ float red[8][33];
  float s = 0.f;
  if (c < n) {
    if (weight) for (int r = r0 + ry; r < r1; r += 8) s = fmaf(__ldg(g + r * ld + c), __ldg(weight + r), s);
    else for (int r = r0 + ry; r < r1; r += 8) s += __ldg(g + r * ld + c);
  }
  red[ry][threadIdx.x & 31] = s;
  __syncthreads();
  if (ry == 0 && c < n) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += red[y][threadIdx.x & 31];
    part[(size_t)blockIdx.y * n + c] = t;
  }
}

// g[r][n] = dy[r][n] * leaky'(y[r][n])
__global__ void __launch_bounds__(256) mask_kernel(const float* __restrict__ dy, const float* __restrict__ y, size_t n, float alpha, float* __restrict__ g) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) g[i] = __ldg(dy + i) * (__ldg(y + i) > 0.f ? 1.f : alpha);
}

}  // namespace

int gj_num_sms();
void gj_set_error(const char* fmt, ...);

#define DN_CHECK(what)                                                                                \
  do { cudaError_t ce_ = cudaGetLastError();                                                          \
       if (ce_ != cudaSuccess) { gj_set_error(what ": %s", cudaGetErrorString(ce_)); return GJ_ERR_CUDA; } } while (0)

static int pow2_cols(int n) { int c = 32; while (c < n) c <<= 1; return c; }

// launches one GEMM (grid.z = nsplit slices over k when nsplit > 1: the caller reduces the partials)
static int launch_gemm(GemmP P, int precision, int nsplit, cudaStream_t st) {
  if (P.M <= 0 || P.N <= 0) return GJ_OK;
  if (nsplit <= 1) { nsplit = 1; P.kchunk = P.K > 0 ? P.K : 1; P.part_stride = 0; }
  if (precision == GJ_PREC_BF16) {
    int NT = (P.N + 15) & ~15; if (NT > 256) NT = 256;
    const int smem = 128 + 2 * ((tcg_stage_bytes(NT) + 127) & ~127) + 1024;
    cudaError_t ce = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 + 2 * ((tcg_stage_bytes(256) + 127) & ~127) + 1024);
    if (ce != cudaSuccess) { gj_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
    dim3 grid((P.M + 127) / 128, (P.N + NT - 1) / NT, nsplit);
    gemm_tc_kernel<<<grid, 128, smem, st>>>(P, NT, pow2_cols(NT));
  } else {
    dim3 grid((P.M + 63) / 64, (P.N + 63) / 64, nsplit);
    gemm_simt_kernel<<<grid, 256, 0, st>>>(P);
  }
  DN_CHECK("dense gemm launch");
  return GJ_OK;
}

static GemmP gemm_base(int M, int N, int K) {
  GemmP P; memset(&P, 0, sizeof(P));
  P.M = M; P.N = N; P.K = K; P.Cm = M; P.Cn = N; P.kchunk = K > 0 ? K : 1;
  return P;
}

// y (rows x N) = epi( x (rows x K; `xk` live columns, row stride ldx) . W^T (W: N x K, row stride ldw) )
static int dense_fwd(int rows, int N, int K, const float* x, long long ldx, int xk, const float* W, long long ldw, const float* bias, int act, float alpha,
                     int accumulate, float* y, long long ldy, int precision, cudaStream_t st) {
  GemmP P = gemm_base(rows, N, K);
  P.A = x; P.sAm = ldx; P.sAk = 1; P.Am = rows; P.Ak = xk;
  P.B = W; P.sBn = ldw; P.sBk = 1; P.Bn = N; P.Bk = K;
  P.bias = bias; P.act = act; P.alpha = alpha; P.accumulate = accumulate; P.C = y; P.ldc = ldy;
  return launch_gemm(P, precision, 1, st);
}
// dx (rows x K; only the first `xk` columns are stored) = epi( g (rows x N) . W (N x K) )
static int dense_dgrad(int rows, int N, int K, const float* g, long long ldg, const float* W, long long ldw, int act, float alpha, const float* aux, long long ldaux,
                       int accumulate, float* dx, long long lddx, int xk, int precision, cudaStream_t st) {
  GemmP P = gemm_base(rows, K, N);
  P.A = g; P.sAm = ldg; P.sAk = 1; P.Am = rows; P.Ak = N;
  P.B = W; P.sBn = 1; P.sBk = ldw; P.Bn = K; P.Bk = N;
  P.act = act; P.alpha = alpha; P.aux = aux; P.ldaux = ldaux; P.accumulate = accumulate; P.C = dx; P.ldc = lddx; P.Cn = xk;
  return launch_gemm(P, precision, 1, st);
}
static int wgrad_splits(int rows, int N, int K) {
  const int tiles = ((N + 127) / 128) * ((K + 255) / 256);
  int s = (2 * gj_num_sms() + tiles - 1) / tiles;
  const int maxs = (rows + 255) / 256;      // at least 256 rows per slice
  if (s > maxs) s = maxs;
  return s < 1 ? 1 : s;
}
static size_t wgrad_part_floats(int rows, int N, int K) { return (size_t)wgrad_splits(rows, N, K) * N * K; }

// Deferred split reductions: while a queue is open (ReduceQueue on the caller's stack) dense_wgrad / dense_colsum take their partial
// buffers from a bump allocator over the queue's region and append the reduction instead of launching it; flush() launches them all.
struct ReduceQueue {
  ReduceList list; float* base; size_t cap, used; cudaStream_t st;
  ReduceQueue(float* b, size_t c, cudaStream_t s) : base(b), cap(c), used(0), st(s) { list.n = 0; list.blk0[0] = 0; }
  float* take(size_t n) { n = (n + 63) & ~(size_t)63; if (used + n > cap || list.n >= RM_MAX) return nullptr; float* p = base + used; used += n; return p; }
  void push(const float* part, int nz, int M, int N, float* out, long long ldo, int accumulate) {
    ReduceJob& J = list.job[list.n];
    J.part = part; J.out = out; J.ldo = ldo; J.nz = nz; J.M = M; J.N = N; J.accumulate = accumulate;
    list.blk0[list.n + 1] = list.blk0[list.n] + (int)(((long long)M * N + 255) / 256);
    ++list.n;
  }
  int flush() {
    if (list.n == 0) return GJ_OK;
    reduce_multi_kernel<<<(unsigned)list.blk0[list.n], 256, 0, st>>>(list);
    list.n = 0; used = 0;
    DN_CHECK("dense reduce launch");
    return GJ_OK;
  }
};
// dW (N x K, row stride lddw; columns >= xk come out as zero) = g^T (rows x N) . x (rows x K, `xk` live columns)
static int dense_wgrad(int rows, int N, int K, const float* g, long long ldg, const float* x, long long ldx, int xk, float* dW, long long lddw, float* part,
                       int precision, cudaStream_t st, int accumulate = 0, ReduceQueue* q = nullptr) {
  const int ns = wgrad_splits(rows, N, K);
  int chunk = (rows + ns - 1) / ns; chunk = (chunk + 63) & ~63;
  const int nz = (rows + chunk - 1) / chunk;
  if (q) {      // deferred: private partial buffer, reduction queued
    float* mine = q->take((size_t)(nz > 1 ? nz : 2) * N * K);
    if (!mine) { if (int rc = q->flush()) return rc; mine = q->take((size_t)(nz > 1 ? nz : 2) * N * K); }
    if (mine) part = mine; else q = nullptr;      // a reduction larger than the whole region: immediate mode on the shared buffer
  }
  GemmP P = gemm_base(N, K, rows);
  P.A = g; P.sAm = 1; P.sAk = ldg; P.Am = N; P.Ak = rows;
  P.B = x; P.sBn = 1; P.sBk = ldx; P.Bn = xk; P.Bk = rows;
  P.C = part; P.ldc = K; P.kchunk = chunk; P.part_stride = (long long)N * K;
  if (int rc = launch_gemm(P, precision, nz > 1 ? nz : 2, st)) return rc;      // (nz == 1 still goes through the partial buffer)
  if (q) { q->push(part, nz > 1 ? nz : 2, N, K, dW, lddw, accumulate); return GJ_OK; }
  const long long n = (long long)N * K;
  reduce_split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(part, nz > 1 ? nz : 2, N, K, dW, lddw, accumulate);
  DN_CHECK("dense wgrad reduce launch");
  return GJ_OK;
}
// db[n] = sum_r g[r][n]
static int dense_colsum(int rows, int N, const float* g, long long ldg, float* db, float* part, cudaStream_t st, int accumulate = 0,
                        const float* weight = nullptr, ReduceQueue* q = nullptr) {
  int slices = (rows + 1023) / 1024; if (slices > 256) slices = 256; if (slices < 1) slices = 1;
  const int per = (rows + slices - 1) / slices;
  if (q) {
    float* mine = q->take((size_t)slices * N);
    if (!mine) { if (int rc = q->flush()) return rc; mine = q->take((size_t)slices * N); }
    if (mine) part = mine; else q = nullptr;
  }
  colsum_kernel<<<dim3((N + 31) / 32, slices), 256, 0, st>>>(g, ldg, rows, N, per, weight, part);
  if (q) { q->push(part, slices, 1, N, db, N, accumulate); DN_CHECK("dense colsum launch"); return GJ_OK; }
  reduce_split_kernel<<<(N + 255) / 256, 256, 0, st>>>(part, slices, 1, N, db, N, accumulate);
  DN_CHECK("dense colsum launch");
  return GJ_OK;
}

// ---- the node-level pieces of one message-passing step -------------------------------------------------------------
// workspace (floats): forward: activations y_0 .. y_{Ln-2}; backward: y_0 .. y_{Ln-1}, two gradient buffers, partials
static size_t ws_align(size_t n) { return (n + 63) & ~(size_t)63; }
struct DenseWs { size_t y[GJ_MAX_LAYERS], g0, g1, part, part_floats, total; };
static DenseWs dense_plan(const MPLayout& L, bool backward) {
  DenseWs w; size_t off = 0;
  const size_t rows = (size_t)L.B * L.N;
  int wmax = 0;
  for (int m = 0; m < L.Ln; ++m) { w.y[m] = off; if (backward || m + 1 < L.Ln) off += ws_align(rows * L.O[m]); if (L.O[m] > wmax) wmax = L.O[m]; }
  w.g0 = w.g1 = w.part = off;
  if (backward) {
    w.g0 = off; off += ws_align(rows * wmax);
    w.g1 = off; off += ws_align(rows * wmax);
    // partial buffers of every weight-gradient / bias reduction of ONE call (gj_dense_post_bwd or gj_dense_pre_bwd): the reductions
    // are queued and run as one launch at the end of the call, so the buffers must not overlap
    size_t p = 0;
    for (int m = 0; m < L.Ln; ++m)
      p += ws_align(wgrad_part_floats((int)rows, L.O[m], L.I[m]) + 2 * (size_t)L.O[m] * L.I[m]) * (m == 0 ? 2 : 1) + ws_align(256 * (size_t)L.O[m]);
    const size_t q = 2 * ws_align(wgrad_part_floats((int)rows, L.E[0], L.H) + 2 * (size_t)L.E[0] * L.H) + ws_align(256 * (size_t)L.E[0]);
    if (q > p) p = q;
    if (p < 256 * 1024) p = 256 * 1024;
    w.part_floats = p;
    w.part = off; off += ws_align(p);
  }
  w.total = off;
  return w;
}
size_t gj_dense_ws_floats(const MPLayout& L, bool backward) { return dense_plan(L, backward).total; }
// floats of the node-MLP activations y_0 .. y_{Ln-1} a forward call can leave for the backward call (same offsets as the plan's)
size_t gj_dense_ysave_floats(const MPLayout& L) { const DenseWs w = dense_plan(L, true); return w.g0; }

// P = Wa h + b0 | Q = Wb h   (pq: rows x 2 E0p)
int gj_dense_pre_fwd(const MPLayout& L, const float* h, const float* params, float* pq, int precision, cudaStream_t st) {
  const int rows = L.B * L.N;
  const float* W0 = params + L.pW[0];
  if (L.E0p != L.E[0]) cudaMemsetAsync(pq, 0, (size_t)rows * 2 * L.E0p * sizeof(float), st);
  if (int rc = dense_fwd(rows, L.E[0], L.H, h, L.ld, L.cols, W0, L.K[0], params + L.pb[0], 0, 0.f, 0, pq, 2 * L.E0p, precision, st)) return rc;
  return dense_fwd(rows, L.E[0], L.H, h, L.ld, L.cols, W0 + L.H, L.K[0], nullptr, 0, 0.f, 0, pq + L.E0p, 2 * L.E0p, precision, st);
}

// node MLP: y_0 = leaky(V0 [e | h] + c0), y_m = leaky(V_m y_{m-1} + c_m); the last layer writes `out`, the others the workspace
static int dense_post_chain(const MPLayout& L, const float* e, const float* h, const float* params, float* out, float* ws, const DenseWs& w,
                            bool keep_last_in_ws, int precision, cudaStream_t st) {
  const int rows = L.B * L.N;
  for (int m = 0; m < L.Ln; ++m) {
    float* y = (m + 1 < L.Ln || keep_last_in_ws) ? ws + w.y[m] : out;
    const float* V = params + L.pV[m];
    const float* c = params + L.pc[m];
    if (m == 0) {
      if (int rc = dense_fwd(rows, L.O[0], L.EL, e, L.EL, L.EL, V, L.I[0], c, 0, 0.f, 0, y, L.O[0], precision, st)) return rc;
      if (int rc = dense_fwd(rows, L.O[0], L.H, h, L.ld, L.cols, V + L.EL, L.I[0], nullptr, 1, L.alpha, 1, y, L.O[0], precision, st)) return rc;
    } else {
      if (int rc = dense_fwd(rows, L.O[m], L.I[m], ws + w.y[m - 1], L.I[m], L.I[m], V, L.I[m], c, 1, L.alpha, 0, y, L.O[m], precision, st)) return rc;
    }
  }
  return GJ_OK;
}
// ysave (optional, gj_dense_ysave_floats): receives every layer's activation for gj_dense_post_bwd
int gj_dense_post_fwd(const MPLayout& L, const float* e, const float* h, const float* params, float* h_out, float* ws, float* ysave, int precision,
                      cudaStream_t st) {
  if (!ysave) return dense_post_chain(L, e, h, params, h_out, ws, dense_plan(L, false), false, precision, st);
  const DenseWs w = dense_plan(L, true);
  if (int rc = dense_post_chain(L, e, h, params, nullptr, ysave, w, true, precision, st)) return rc;
  cudaError_t ce = cudaMemcpyAsync(h_out, ysave + w.y[L.Ln - 1], (size_t)L.B * L.N * L.O[L.Ln - 1] * sizeof(float), cudaMemcpyDeviceToDevice, st);
  if (ce != cudaSuccess) { gj_set_error("cudaMemcpyAsync: %s", cudaGetErrorString(ce)); return GJ_ERR_CUDA; }
  return GJ_OK;
}

// adjoint of the node MLP: de, dh (first `cols` columns OVERWRITTEN), node parameter gradients written to dparams
// ysave (optional): the activations a forward call left (gj_dense_post_fwd); otherwise the chain is recomputed
int gj_dense_post_bwd(const MPLayout& L, const float* e, const float* h, const float* params, const float* dh_out, float* de, float* dh,
                      float* dparams, float* ws, const float* ysave, int precision, cudaStream_t st) {
  const DenseWs w = dense_plan(L, true);
  const int rows = L.B * L.N;
  if (!ysave) { if (int rc = dense_post_chain(L, e, h, params, nullptr, ws, w, true, precision, st)) return rc; }
  const float* yb = ysave ? ysave : ws;      // activations y_m at yb + w.y[m]
  float* g = ws + w.g0;
  float* gn = ws + w.g1;
  float* part = ws + w.part;
  ReduceQueue rq(part, w.part_floats, st);
  {
    const size_t n = (size_t)rows * L.O[L.Ln - 1];
    mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dh_out, yb + w.y[L.Ln - 1], n, L.alpha, g);
    DN_CHECK("dense mask launch");
  }
  for (int m = L.Ln - 1; m >= 0; --m) {
    const int O = L.O[m], I = L.I[m];
    const float* V = params + L.pV[m];
    if (int rc = dense_colsum(rows, O, g, O, dparams + L.pc[m], part, st, 0, nullptr, &rq)) return rc;
    if (m > 0) {
      if (int rc = dense_wgrad(rows, O, I, g, O, yb + w.y[m - 1], I, I, dparams + L.pV[m], I, part, precision, st, 0, &rq)) return rc;
      if (int rc = dense_dgrad(rows, O, I, g, O, V, I, 2, L.alpha, yb + w.y[m - 1], I, 0, gn, I, I, precision, st)) return rc;
      float* t = g; g = gn; gn = t;
    } else {
      if (int rc = dense_wgrad(rows, O, L.EL, g, O, e, L.EL, L.EL, dparams + L.pV[0], I, part, precision, st, 0, &rq)) return rc;
      if (int rc = dense_wgrad(rows, O, L.H, g, O, h, L.ld, L.cols, dparams + L.pV[0] + L.EL, I, part, precision, st, 0, &rq)) return rc;
      if (int rc = dense_dgrad(rows, O, L.EL, g, O, V, I, 0, 0.f, nullptr, 0, 0, de, L.EL, L.EL, precision, st)) return rc;
      if (int rc = dense_dgrad(rows, O, L.H, g, O, V + L.EL, I, 0, 0.f, nullptr, 0, 0, dh, L.ld, L.cols, precision, st)) return rc;
    }
  }
  return rq.flush();
}

// adjoint of the projections: dh += dP Wa + dQ Wb; dW0[:, 0:H] = dP^T h, dW0[:, H:2H] = dQ^T h, db0 = sum dP
int gj_dense_pre_bwd(const MPLayout& L, const float* h, const float* params, const float* dpq, float* dh, float* dparams, float* ws, int precision,
                     cudaStream_t st) {
  const DenseWs w = dense_plan(L, true);
  const int rows = L.B * L.N, E0 = L.E[0], ldp = 2 * L.E0p;
  const float* W0 = params + L.pW[0];
  float* part = ws + w.part;
  if (int rc = dense_dgrad(rows, E0, L.H, dpq, ldp, W0, L.K[0], 0, 0.f, nullptr, 0, 1, dh, L.ld, L.cols, precision, st)) return rc;
  if (int rc = dense_dgrad(rows, E0, L.H, dpq + L.E0p, ldp, W0 + L.H, L.K[0], 0, 0.f, nullptr, 0, 1, dh, L.ld, L.cols, precision, st)) return rc;
  ReduceQueue rq(part, w.part_floats, st);
  if (int rc = dense_wgrad(rows, E0, L.H, dpq, ldp, h, L.ld, L.cols, dparams + L.pW[0], L.K[0], part, precision, st, 0, &rq)) return rc;
  if (int rc = dense_wgrad(rows, E0, L.H, dpq + L.E0p, ldp, h, L.ld, L.cols, dparams + L.pW[0] + L.H, L.K[0], part, precision, st, 0, &rq)) return rc;
  if (int rc = dense_colsum(rows, E0, dpq, ldp, dparams + L.pb[0], part, st, 0, nullptr, &rq)) return rc;
  return rq.flush();
}

// number of kernels the four entry points launch (gj_mp_step_launches)
int gj_dense_launches(const MPLayout& L, bool backward) {
  const int fwd_chain = L.Ln + 1;
  if (!backward) return 2 + fwd_chain;
  // node MLP adjoint: chain again, mask, per layer a column sum, one weight-gradient and one input-gradient GEMM per product (two
  // products in layer 0: [e | h]), ONE launch for all queued reductions; P|Q again (2); projections' adjoint: 2 dgrad + 2 wgrad +
  // column sum + one reduction launch
  return fwd_chain + 1 + (L.Ln + 2 * (L.Ln - 1) + 4) + 1 + 2 + 6;
}

// test hook (declared in the header): C = epi(A B^T-like product) through the same kernels
extern "C" int gj_dense_gemm(int32_t form, int32_t M, int32_t N, int32_t K, const float* A, const float* B, const float* bias, int32_t act, float alpha,
                             const float* aux, int32_t accumulate, float* C, void* workspace, size_t workspace_bytes, int32_t precision, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!A || !B || !C || M < 1 || N < 1 || K < 1) { gj_set_error("gj_dense_gemm: bad argument"); return GJ_ERR_INVALID; }
  if (act == 2 && (form != 1 || !aux)) { gj_set_error("gj_dense_gemm: the slope-mask epilogue belongs to the input gradient (form 1) and needs aux"); return GJ_ERR_INVALID; }
  if (form == 0) return dense_fwd(M, N, K, A, K, K, B, K, bias, act, alpha, accumulate, C, N, precision, st);
  if (form == 1) return dense_dgrad(M, K, N, A, K, B, N, act, alpha, aux, N, accumulate, C, N, N, precision, st);      // C (M x N) = A (M x K) . B (K x N)
  if (form == 2) {                                                                                                       // C (M x N) = A^T (K x M)^T . B (K x N)
    if (!workspace || workspace_bytes < (wgrad_part_floats(K, M, N) + 2 * (size_t)M * N) * sizeof(float)) { gj_set_error("gj_dense_gemm: workspace too small"); return GJ_ERR_WORKSPACE; }
    return dense_wgrad(K, M, N, A, M, B, N, N, C, N, (float*)workspace, precision, st);
  }
  gj_set_error("gj_dense_gemm: unknown form");
  return GJ_ERR_INVALID;
}
extern "C" size_t gj_dense_gemm_workspace(int32_t form, int32_t M, int32_t N, int32_t K) {
  return form == 2 ? (wgrad_part_floats(K, M, N) + 2 * (size_t)M * N) * sizeof(float) : 16;
}

// =====================================================================================================================
// Materialised edge path: the edge MLP of one message-passing step (reference models/graphnet.py:186-223 `_getA`, :273-289
// `_edge_conv`, the sum over j of `_concat` :243) for ANY layer count and width, as the same generic GEMMs over the edge rows of
// a CHUNK of jets at a time (the activations of a chunk live in the workspace, never those of the whole batch).  Used where
// neither the fused tensor-core kernels nor the shared-memory plans of edge_simt.cu / edge_tc.cu cover the widths (e.g.
// edge_sizes [[256, 256]] of BASELINE config 5: W1 alone is 128 KB of bf16).  The first layer stays factorised:
// a0 = leaky(P_i + Q_j + wd d_ij) is one elementwise kernel, its adjoint a few reductions.
// =====================================================================================================================
namespace {

__global__ void __launch_bounds__(256) em_pairdist_kernel(const float* __restrict__ h, int nb, int N, int H, int cols, int ld, int mink,
                                                          float* __restrict__ d) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)nb * N * N) return;
  const int j = (int)(idx % N), i = (int)((idx / N) % N), b = (int)(idx / ((long long)N * N));
  const float* hi = h + ((size_t)b * N + i) * ld;
  const float* hj = h + ((size_t)b * N + j) * ld;
  float acc = 0.f;
  for (int k = 0; k < cols; ++k) {
    const float x = __ldg(hj + k) - __ldg(hi + k) + GJ_EPS;
    acc = (mink && k > 0) ? fmaf(-x, x, acc) : fmaf(x, x, acc);
  }
  for (int k = cols; k < H; ++k) acc += (mink && k > 0) ? -GJ_EPS * GJ_EPS : GJ_EPS * GJ_EPS;      // zero-padded columns
  d[idx] = acc;
}

// a0[(b, i, j)][c] = leaky(P[b, i][c] + Q[b, j][c] + wd[c] d[b, i, j])
__global__ void __launch_bounds__(256) em_first_kernel(const float* __restrict__ pq, const float* __restrict__ d, const float* __restrict__ w0,
                                                       int nb, int N, int E0, int E0p, int K0, float alpha, float* __restrict__ a0) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)nb * N * N * E0) return;
  const int c = (int)(idx % E0);
  const long long r = idx / E0;
  const int j = (int)(r % N), i = (int)((r / N) % N), b = (int)(r / ((long long)N * N));
  const float z = __ldg(pq + ((size_t)b * N + i) * 2 * E0p + c) + __ldg(pq + ((size_t)b * N + j) * 2 * E0p + E0p + c) +
                  __ldg(w0 + (size_t)c * K0 + (K0 - 1)) * __ldg(d + r);
  a0[idx] = z > 0.f ? z : alpha * z;
}

// e[b, i][c] = sum_j a[(b, i, j)][c]
__global__ void __launch_bounds__(256) em_sumj_kernel(const float* __restrict__ a, int nb, int N, int E, float* __restrict__ e) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)nb * N * E) return;
  const int c = (int)(idx % E);
  const long long bi = idx / E;
  const float* src = a + bi * N * E + c;
  float s = 0.f;
  for (int j = 0; j < N; ++j) s += __ldg(src + (size_t)j * E);
  e[idx] = s;
}

// dz[(b, i, j)][c] = de[b, i][c] leaky'(a[(b, i, j)][c])
__global__ void __launch_bounds__(256) em_dzlast_kernel(const float* __restrict__ de, const float* __restrict__ a, int nb, int N, int E, float alpha,
                                                        float* __restrict__ dz) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)nb * N * N * E) return;
  const int c = (int)(idx % E);
  const long long bi = idx / ((long long)N * E);
  dz[idx] = __ldg(de + bi * E + c) * (__ldg(a + idx) > 0.f ? 1.f : alpha);
}

// dP[b, n][c] = sum_j dz0[(b, n, j)][c],  dQ[b, n][c] = sum_i dz0[(b, i, n)][c]
__global__ void __launch_bounds__(256) em_dpq_kernel(const float* __restrict__ dz0, int nb, int N, int E0, int E0p, float* __restrict__ dpq) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)nb * N * E0) return;
  const int c = (int)(idx % E0);
  const long long bn = idx / E0;
  const int n = (int)(bn % N);
  const long long b = bn / N;
  const float* rowi = dz0 + (b * N + n) * N * E0 + c;      // (b, n, j): stride E0 over j
  const float* colj = dz0 + (b * N * N + n) * E0 + c;      // (b, i, n): stride N E0 over i
  float sp = 0.f, sq = 0.f;
  for (int m = 0; m < N; ++m) { sp += __ldg(rowi + (size_t)m * E0); sq += __ldg(colj + (size_t)m * N * E0); }
  dpq[bn * 2 * E0p + c] = sp;
  dpq[bn * 2 * E0p + E0p + c] = sq;
}

// G[r] = sum_c dz0[r][c] wd[c]   (one warp per edge row)
__global__ void __launch_bounds__(256) em_g_kernel(const float* __restrict__ dz0, const float* __restrict__ w0, long long rows, int E0, int K0,
                                                   float* __restrict__ G) {
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int c = lane; c < E0; c += 32) s = fmaf(__ldg(dz0 + r * E0 + c), __ldg(w0 + (size_t)c * K0 + (K0 - 1)), s);
  s = gj_warp_sum(s);
  if (lane == 0) G[r] = s;
}

// dh[b, n][k] += 2 s_k sum_m [ G[b, m, n] (h_n - h_m + eps) - G[b, n, m] (h_m - h_n + eps) ]
__global__ void __launch_bounds__(256) em_dist_bwd_kernel(const float* __restrict__ h, const float* __restrict__ G, int nb, int N, int cols, int ld,
                                                          int mink, float* __restrict__ dh) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)nb * N * cols) return;
  const int k = (int)(idx % cols);
  const long long bn = idx / cols;
  const int n = (int)(bn % N);
  const long long b = bn / N;
  const float hn = __ldg(h + bn * ld + k);
  float acc = 0.f;
  for (int m = 0; m < N; ++m) {
    const float hm = __ldg(h + (b * N + m) * ld + k);
    acc = fmaf(__ldg(G + (b * N + m) * N + n), hn - hm + GJ_EPS, acc);
    acc = fmaf(-__ldg(G + (b * N + n) * N + m), hm - hn + GJ_EPS, acc);
  }
  dh[bn * ld + k] += ((mink && k > 0) ? -2.f : 2.f) * acc;
}

__global__ void em_scatter_wd_kernel(const float* __restrict__ dwd, int E0, int K0, float* __restrict__ dW0) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c < E0) dW0[(size_t)c * K0 + (K0 - 1)] = dwd[c];
}

struct MatWs { size_t d, G, dwd, act[GJ_MAX_LAYERS], dz[2], part, total; int jc; };

MatWs mat_plan(const MPLayout& L, bool backward) {
  MatWs w; memset(&w, 0, sizeof(w));
  int emax = 0, esum = 0;
  for (int l = 0; l < L.Le; ++l) { esum += L.E[l]; if (L.E[l] > emax) emax = L.E[l]; }
  const long long per_jet = (long long)L.N * L.N * (2 + (backward ? esum + 2 * emax : 2 * emax));
  long long jc = (96LL << 20) / (per_jet > 0 ? per_jet : 1);
  if (jc < 1) jc = 1;
  if (jc > L.B) jc = L.B;
  w.jc = (int)jc;
  const size_t rows = (size_t)jc * L.N * L.N;
  size_t off = 0;
  w.d = off; off += ws_align(rows);
  w.G = off; off += ws_align(rows);
  w.dwd = off; off += ws_align((size_t)L.E[0]);
  if (backward) {
    for (int l = 0; l < L.Le; ++l) { w.act[l] = off; off += ws_align(rows * L.E[l]); }
    w.dz[0] = off; off += ws_align(rows * emax);
    w.dz[1] = off; off += ws_align(rows * emax);
    size_t p = 256 * 1024;
    for (int l = 1; l < L.Le; ++l) { const size_t q = wgrad_part_floats((int)rows, L.E[l], L.K[l]) + 2 * (size_t)L.E[l] * L.K[l]; if (q > p) p = q; }
    w.part = off; off += ws_align(p);
  } else {
    w.act[0] = off; off += ws_align(rows * emax);
    w.act[1] = off; off += ws_align(rows * emax);
  }
  w.total = off;
  return w;
}

inline unsigned em_blocks(long long n) { return (unsigned)((n + 255) / 256); }

}  // namespace

size_t gj_edge_mat_ws_floats(const MPLayout& L, bool backward) { return mat_plan(L, backward).total; }

// chunk's forward chain; keep_all: every layer's activation stays in w.act[l] (backward), else two buffers alternate
static int mat_forward_chunk(const MPLayout& L, const MatWs& w, float* ws, const float* h, const float* pq, const float* params, int b0, int nb,
                             bool keep_all, int precision, cudaStream_t st, float** a_last) {
  const long long rows = (long long)nb * L.N * L.N;
  em_pairdist_kernel<<<em_blocks(rows), 256, 0, st>>>(h + (size_t)b0 * L.N * L.ld, nb, L.N, L.H, L.cols, L.ld, L.mink, ws + w.d);
  float* a = ws + w.act[0];
  em_first_kernel<<<em_blocks(rows * L.E[0]), 256, 0, st>>>(pq + (size_t)b0 * L.N * 2 * L.E0p, ws + w.d, params + L.pW[0], nb, L.N, L.E[0], L.E0p, L.K[0],
                                                             L.alpha, a);
  DN_CHECK("edge_mat first layer launch");
  for (int l = 1; l < L.Le; ++l) {
    float* y = ws + w.act[keep_all ? l : (l & 1)];
    if (int rc = dense_fwd((int)rows, L.E[l], L.K[l], a, L.K[l], L.K[l], params + L.pW[l], L.K[l], params + L.pb[l], 1, L.alpha, 0, y, L.E[l], precision, st)) return rc;
    a = y;
  }
  *a_last = a;
  return GJ_OK;
}

int gj_edge_mat_fwd(const MPLayout& L, const float* h, const float* pq, const float* params, float* e_out, float* ws, int precision, cudaStream_t st) {
  const MatWs w = mat_plan(L, false);
  if ((long long)w.jc * L.N * L.N * 256 > 0x7fffffffLL) { gj_set_error("edge_mat: chunk too large"); return GJ_ERR_INVALID; }
  for (int b0 = 0; b0 < L.B; b0 += w.jc) {
    const int nb = L.B - b0 < w.jc ? L.B - b0 : w.jc;
    float* a = nullptr;
    if (int rc = mat_forward_chunk(L, w, ws, h, pq, params, b0, nb, false, precision, st, &a)) return rc;
    em_sumj_kernel<<<em_blocks((long long)nb * L.N * L.EL), 256, 0, st>>>(a, nb, L.N, L.EL, e_out + (size_t)b0 * L.N * L.EL);
    DN_CHECK("edge_mat sum launch");
  }
  return GJ_OK;
}

// edge adjoint: dpq (OVERWRITTEN), dh += distance path, dparams: edge layers >= 1 and the wd column of W0 (the Wa | Wb columns and
// b0 belong to the projections' adjoint)
int gj_edge_mat_bwd(const MPLayout& L, const float* h, const float* pq, const float* params, const float* de, float* dpq, float* dh,
                    float* dparams, float* ws, int precision, cudaStream_t st) {
  const MatWs w = mat_plan(L, true);
  if ((long long)w.jc * L.N * L.N * 256 > 0x7fffffffLL) { gj_set_error("edge_mat: chunk too large"); return GJ_ERR_INVALID; }
  if (L.E0p != L.E[0]) cudaMemsetAsync(dpq, 0, (size_t)L.B * L.N * 2 * L.E0p * sizeof(float), st);
  float* part = ws + w.part;
  for (int b0 = 0; b0 < L.B; b0 += w.jc) {
    const int nb = L.B - b0 < w.jc ? L.B - b0 : w.jc, acc = b0 > 0 ? 1 : 0;
    const long long rows = (long long)nb * L.N * L.N;
    float* a_last = nullptr;
    if (int rc = mat_forward_chunk(L, w, ws, h, pq, params, b0, nb, true, precision, st, &a_last)) return rc;
    float* dz = ws + w.dz[0];
    float* dzn = ws + w.dz[1];
    em_dzlast_kernel<<<em_blocks(rows * L.EL), 256, 0, st>>>(de + (size_t)b0 * L.N * L.EL, a_last, nb, L.N, L.EL, L.alpha, dz);
    DN_CHECK("edge_mat dz launch");
    for (int l = L.Le - 1; l >= 1; --l) {
      const float* a_prev = ws + w.act[l - 1];
      if (int rc = dense_colsum((int)rows, L.E[l], dz, L.E[l], dparams + L.pb[l], part, st, acc)) return rc;
      if (int rc = dense_wgrad((int)rows, L.E[l], L.K[l], dz, L.E[l], a_prev, L.K[l], L.K[l], dparams + L.pW[l], L.K[l], part, precision, st, acc)) return rc;
      if (int rc = dense_dgrad((int)rows, L.E[l], L.K[l], dz, L.E[l], params + L.pW[l], L.K[l], 2, L.alpha, a_prev, L.K[l], 0, dzn, L.K[l], L.K[l], precision, st)) return rc;
      float* t = dz; dz = dzn; dzn = t;
    }
    // dz = dz0 (rows x E0)
    em_dpq_kernel<<<em_blocks((long long)nb * L.N * L.E[0]), 256, 0, st>>>(dz, nb, L.N, L.E[0], L.E0p, dpq + (size_t)b0 * L.N * 2 * L.E0p);
    em_g_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(dz, params + L.pW[0], rows, L.E[0], L.K[0], ws + w.G);
    if (int rc = dense_colsum((int)rows, L.E[0], dz, L.E[0], ws + w.dwd, part, st, acc, ws + w.d)) return rc;
    em_dist_bwd_kernel<<<em_blocks((long long)nb * L.N * L.cols), 256, 0, st>>>(h + (size_t)b0 * L.N * L.ld, ws + w.G, nb, L.N, L.cols, L.ld, L.mink,
                                                                               dh + (size_t)b0 * L.N * L.ld);
    DN_CHECK("edge_mat first-layer adjoint launch");
  }
  em_scatter_wd_kernel<<<(L.E[0] + 127) / 128, 128, 0, st>>>(ws + w.dwd, L.E[0], L.K[0], dparams + L.pW[0]);
  DN_CHECK("edge_mat scatter launch");
  return GJ_OK;
}

int gj_edge_mat_launches(const MPLayout& L, bool backward) {
  const int chunks = (L.B + mat_plan(L, backward).jc - 1) / mat_plan(L, backward).jc;
  if (!backward) return chunks * (2 + (L.Le - 1) + 1);
  return chunks * (2 + (L.Le - 1) + 1 + (L.Le - 1) * (2 + 2 + 1) + 2 + 2 + 1) + 1;
}
