// Small bandwidth-bound pieces around the two GEMM stages:
//   * context vectors + gather from NCHW maps + sum of squares  (preprocessing.py:179-193, 234-256)
//   * normalize_vectors with a global statistic                  (preprocessing.py:219-231)
//   * per-instance concept relevances                            (explainer.py:206-242)
#include <cuda_fp16.h>
#include "common.cuh"

namespace drsa {

namespace {
// grid (N, ceil(L/32), ceil(d/32)), block 32x8.  Reads are coalesced along positions, writes along
// channels (32x32 shared-memory transpose).
__global__ void __launch_bounds__(256) context_gather_kernel(const float* __restrict__ a_map, const float* __restrict__ R_map,
                                                             int d, int HW, const int64_t* __restrict__ idx, int L,
                                                             float* __restrict__ act_out, float* __restrict__ ctx_out,
                                                             double* __restrict__ sumsq) {
  __shared__ float ta[32][33], tcx[32][33];
  const int n = blockIdx.x, l0 = blockIdx.y * 32, c0 = blockIdx.z * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int l = l0 + tx;
  int pos = -1;
  if (l < L) pos = idx != nullptr ? (int)idx[(int64_t)n * L + l] : l;
  float sa = 0.f, sc = 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = c0 + ty + 8 * r;
    float a = 0.f, cx = 0.f;
    if (pos >= 0 && c < d) {
      const int64_t o = ((int64_t)n * d + c) * HW + pos;
      a = __ldg(a_map + o);
      cx = __ldg(R_map + o) / (a + 1e-7f);       // preprocessing.py:193
      sa = fmaf(a, a, sa);
      sc = fmaf(cx, cx, sc);
    }
    ta[ty + 8 * r][tx] = a;
    tcx[ty + 8 * r][tx] = cx;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int ll = l0 + ty + 8 * r, c = c0 + tx;
    if (ll < L && c < d) {
      const int64_t o = ((int64_t)n * L + ll) * d + c;
      act_out[o] = ta[tx][ty + 8 * r];
      ctx_out[o] = tcx[tx][ty + 8 * r];
    }
  }
  if (sumsq != nullptr) {
    // one pair of atomics per CTA: one per warp (65 k double atomics on two addresses for 256 samples) serialised in
    // L2 and made this kernel 7x slower than its memory traffic
    __shared__ double red[2][8];
    const double da = warp_sum((double)sa), dc = warp_sum((double)sc);
    if (tx == 0) { red[0][ty] = da; red[1][ty] = dc; }
    __syncthreads();
    if (threadIdx.x < 2) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
      atomicAdd(&sumsq[threadIdx.x], t);
    }
  }
}

__global__ void normalize_kernel(float* __restrict__ v, int64_t n, const double* __restrict__ sumsq, double inv_count,
                                 float root4_d) {
  const float E = sqrtf((float)(sumsq[0] * inv_count));     // sqrt(mean(v^2)), preprocessing.py:230
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    v[i] = (v[i] / E) / root4_d;                              // preprocessing.py:231
}

__global__ void context_vectors_kernel(const float* __restrict__ a, const float* __restrict__ R, int64_t n,
                                       float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __ldg(R + i) / (__ldg(a + i) + 1e-7f);          // preprocessing.py:193
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ v, int64_t n, double* __restrict__ out) {
  __shared__ double red[8];
  float s = 0.f; double acc = 0.0; int cnt = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = __ldg(v + i);
    s = fmaf(x, x, s);
    if (++cnt == 64) { acc += (double)s; s = 0.f; cnt = 0; }   // bounded fp32 run, fp64 carry
  }
  acc += (double)s;
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out, t);
  }
}

// one warp per (b, k): sum over positions and the concept's columns of HA*HC
__global__ void __launch_bounds__(256) subspace_rel_kernel(const float* __restrict__ HA, const float* __restrict__ HC,
                                                           int64_t B, int64_t P, int m, int K, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= B * K) return;
  const int64_t b = wid / K; const int k = (int)(wid % K); const int d_k = m / K;
  float s = 0.f;
  for (int64_t p = 0; p < P; ++p) {
    const float* ha = HA + (b * P + p) * m + k * d_k;
    const float* hc = HC + (b * P + p) * m + k * d_k;
    for (int j = lane; j < d_k; j += 32) s = fmaf(ha[j], hc[j], s);
  }
  s = warp_sum(s);
  if (lane == 0) out[b * K + k] = s;
}
// The same fused tail of stage 1 straight from the tensor-core stack's NHWC layout: activations as hi + lo fp16 planes
// [N, HW, Cp], relevance fp32 [N, HW, Cp].  A row of the output IS a position of the input, so every access is coalesced along
// the channels and no layout conversion (NHWC -> NCHW -> rows) is needed.  One warp per output row, 8 rows per CTA iteration.
__global__ void __launch_bounds__(256) context_pairs_nhwc_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo,
                                                                 const float* __restrict__ R, int64_t N, int HW, int Cp, int d,
                                                                 const int64_t* __restrict__ idx, int L,
                                                                 float* __restrict__ act_out, float* __restrict__ ctx_out,
                                                                 double* __restrict__ sumsq) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float sa = 0.f, sc = 0.f;
  const int64_t rows = N * L;
  for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < rows; r += (int64_t)gridDim.x * 8) {
    const int64_t n = r / L;
    const int l = (int)(r % L);
    const int pos = idx != nullptr ? (int)idx[n * L + l] : l;
    const int64_t src = (n * HW + pos) * Cp;
    for (int c = 2 * lane; c < d; c += 64) {            // two channels per lane: half2 / float2 accesses (d, Cp even)
      const float2 h = __half22float2(*reinterpret_cast<const __half2*>(hi + src + c));
      const float2 w = __half22float2(*reinterpret_cast<const __half2*>(lo + src + c));
      const float2 rv = *reinterpret_cast<const float2*>(R + src + c);
      const float a0 = h.x + w.x, a1 = h.y + w.y;
      const float c0 = rv.x / (a0 + 1e-7f), c1 = rv.y / (a1 + 1e-7f);      // preprocessing.py:193
      *reinterpret_cast<float2*>(act_out + r * d + c) = make_float2(a0, a1);
      *reinterpret_cast<float2*>(ctx_out + r * d + c) = make_float2(c0, c1);
      sa = fmaf(a0, a0, fmaf(a1, a1, sa));
      sc = fmaf(c0, c0, fmaf(c1, c1, sc));
    }
  }
  if (sumsq != nullptr) {
    __shared__ double red[2][8];
    const double da = warp_sum((double)sa), dc = warp_sum((double)sc);
    if (lane == 0) { red[0][warp] = da; red[1][warp] = dc; }
    __syncthreads();
    if (threadIdx.x < 2) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
      atomicAdd(sumsq + threadIdx.x, t);
    }
  }
}

// Prototype search (prototypes.py:98-119): one warp per (subset, concept) sums relu(s_rk)^2 over the subset's rows
__global__ void __launch_bounds__(256) subset_sumsq_kernel(const float* __restrict__ HA, const float* __restrict__ HC,
                                                           int64_t S, int64_t R, int m, int K, float* __restrict__ sumsq) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= S * K) return;
  const int64_t b = wid / K; const int k = (int)(wid % K); const int d_k = m / K;
  float acc = 0.f;
  for (int64_t r = 0; r < R; ++r) {
    const float* ha = HA + (b * R + r) * m + k * d_k;
    const float* hc = HC + (b * R + r) * m + k * d_k;
    float s = 0.f;
    for (int j = lane; j < d_k; j += 32) s = fmaf(ha[j], hc[j], s);
    s = fmaxf(warp_sum(s), 0.f);                  // ReLU per row (drsa.py:155), then the p = 2 pooling over rows
    acc = fmaf(s, s, acc);
  }
  if (lane == 0) sumsq[b * K + k] = acc;
}
// obj[b] = (mean_k sqrt(sqrt(sumsq[b][k] / R)))^2   (drsa.py:224-238)
__global__ void subset_objective_kernel(const float* __restrict__ sumsq, int64_t S, int64_t R, int K, float* __restrict__ obj) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= S) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc += sqrtf(sqrtf(sumsq[b * K + k] / (float)R));
  acc /= (float)K;
  obj[b] = acc * acc;
}

// out[i] = a[i] + beta * b[i] over the d*m + K row sums (deferred correction of the row rounding, see drsa_sums_combine)
__global__ void __launch_bounds__(256) sums_combine_kernel(const float* __restrict__ a, const float* __restrict__ b, float beta,
                                                           float* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fmaf(beta, b[i], a[i]);
}
}  // namespace

int sums_combine(const float* a, const float* b, float beta, float* out, int64_t n, cudaStream_t stream) {
  sums_combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(a, b, beta, out, n);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

int context_gather(const float* a_map, const float* R_map, int64_t N, int d, int HW, const int64_t* idx, int L,
                   float* act_out, float* ctx_out, double* sumsq, cudaStream_t stream) {
  if (N > 2147483647LL) return DRSA_ERR_SHAPE;
  dim3 grid((unsigned)N, cdiv(L, 32), cdiv(d, 32));
  if (grid.y > 65535 || grid.z > 65535) return DRSA_ERR_SHAPE;
  context_gather_kernel<<<grid, 256, 0, stream>>>(a_map, R_map, d, HW, idx, L, act_out, ctx_out, sumsq);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

int context_pairs_nhwc(const void* hi, const void* lo, const float* R, int64_t N, int HW, int Cp, int d, const int64_t* idx,
                       int L, float* act_out, float* ctx_out, double* sumsq, cudaStream_t stream) {
  if ((d & 1) || (Cp & 1)) return DRSA_ERR_SHAPE;
  int64_t blocks = (N * L + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  context_pairs_nhwc_kernel<<<(int)blocks, 256, 0, stream>>>(static_cast<const __half*>(hi), static_cast<const __half*>(lo), R, N,
                                                              HW, Cp, d, idx, L, act_out, ctx_out, sumsq);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

int normalize(float* v, int64_t rows, int d, const double* sumsq, int64_t count_global, cudaStream_t stream) {
  const int64_t n = rows * d;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  normalize_kernel<<<(int)blocks, 256, 0, stream>>>(v, n, sumsq, 1.0 / (double)count_global,
                                                     (float)pow((double)d, 0.25));
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

int context_vectors(const float* a, const float* R, int64_t count, float* out, cudaStream_t stream) {
  int64_t blocks = (count + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  context_vectors_kernel<<<(int)blocks, 256, 0, stream>>>(a, R, count, out);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

int sumsq(const float* v, int64_t count, double* out, cudaStream_t stream) {
  DRSA_CUDA(cudaMemsetAsync(out, 0, sizeof(double), stream));
  int64_t blocks = (count + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  sumsq_kernel<<<(int)blocks, 256, 0, stream>>>(v, count, out);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

int64_t subspace_relevances_workspace_bytes(int64_t B, int64_t P, int d, int m) {
  (void)d;
  return 2 * align_up(B * P * m * 4, 256);
}

int subspace_relevances(const float* act, const float* ctx, const float* U, int64_t B, int64_t P, int d, int m, int K,
                        float* out, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  if (workspace_bytes < subspace_relevances_workspace_bytes(B, P, d, m)) return DRSA_ERR_WORKSPACE;
  if (B * P > 2147483647LL) return DRSA_ERR_SHAPE;
  float* HA = static_cast<float*>(workspace);
  float* HC = reinterpret_cast<float*>(static_cast<char*>(workspace) + align_up(B * P * m * 4, 256));
  GemmDesc g{};
  g.M = (int)(B * P); g.N = m; g.K = d; g.lda = d; g.ldb = m; g.ldc = m; g.alpha = 1.f; g.splits = 1;
  g.A = act; g.B = U; g.C = HA; DRSA_TRY(sgemm(g, stream));
  g.A = ctx; g.C = HC;          DRSA_TRY(sgemm(g, stream));
  const int64_t warps = B * K;
  subspace_rel_kernel<<<cdiv(warps, 8), 256, 0, stream>>>(HA, HC, B, P, m, K, out);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

// DRSA objective of S subsets of R consecutive rows each, in one pass (get_prototypes_ts, prototypes.py:98-119, evaluates
// obj_val once per subset): two projections of ALL rows, one segmented reduction, S objectives.
int subset_objectives(const float* act, const float* ctx, const float* U, int64_t S, int64_t R, int d, int m, int K,
                      float* obj, float* sumsq, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  if (workspace_bytes < subspace_relevances_workspace_bytes(S, R, d, m)) return DRSA_ERR_WORKSPACE;
  if (S * R > 2147483647LL) return DRSA_ERR_SHAPE;
  float* HA = static_cast<float*>(workspace);
  float* HC = reinterpret_cast<float*>(static_cast<char*>(workspace) + align_up(S * R * m * 4, 256));
  GemmDesc g{};
  g.M = (int)(S * R); g.N = m; g.K = d; g.lda = d; g.ldb = m; g.ldc = m; g.alpha = 1.f; g.splits = 1;
  g.A = act; g.B = U; g.C = HA; DRSA_TRY(sgemm(g, stream));
  g.A = ctx; g.C = HC;          DRSA_TRY(sgemm(g, stream));
  subset_sumsq_kernel<<<cdiv(S * K, 8), 256, 0, stream>>>(HA, HC, S, R, m, K, sumsq);
  DRSA_LAUNCH_CHECK();
  subset_objective_kernel<<<cdiv(S, 128), 128, 0, stream>>>(sumsq, S, R, K, obj);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

}  // namespace drsa
