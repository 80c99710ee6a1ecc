// DRSA row pass, exact fp32 flavour (DRSA_PREC_FP32): CUDA-core GEMMs plus a row
// kernel.  It is the arithmetic twin of the reference (drsa.py:148-155 + autograd
// backward of :100) and the on-device yardstick for the tensor-core kernel; it is also
// the path taken for shapes the tcgen05 kernel does not cover (d < 128, d_k % 32 != 0).
//
//   for each chunk of rows:
//     HA = A_c U ; HC = C_c U                               (2 GEMMs)
//     per row: s_k, g_k = relu(s_k); sumsq_k += g_k^2; HA <- g * HC ; HC <- g * HA
//     X += A_c^T HA + C_c^T HC                              (2 split-K GEMMs, fixed-order reduce)
#include "common.cuh"

namespace drsa {

namespace {
constexpr int ROW_BLOCKS = 592;   // 4 CTAs per SM
constexpr int64_t CHUNK_ROWS = 1 << 18;

// one warp per row; lanes stride over the m projected columns of that row
__global__ void __launch_bounds__(256) row_relevance_kernel(float* __restrict__ HA, float* __restrict__ HC,
                                                            int64_t rows, int m, int K, int d_k,
                                                            double* __restrict__ ss_part) {
  extern __shared__ double sh[];          // [warps][K]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int k = lane; k < K; k += 32) sh[warp * K + k] = 0.0;
  __syncwarp();
  for (int64_t r = (int64_t)blockIdx.x * nwarp + warp; r < rows; r += (int64_t)gridDim.x * nwarp) {
    float* ha = HA + r * m;
    float* hc = HC + r * m;
    for (int k = 0; k < K; ++k) {
      float s = 0.f;
      for (int j = lane; j < d_k; j += 32) s = fmaf(ha[k * d_k + j], hc[k * d_k + j], s);
      s = warp_sum(s);
      const float g = fmaxf(s, 0.f);
      if (lane == 0) sh[warp * K + k] += (double)g * (double)g;
      for (int j = lane; j < d_k; j += 32) {
        const float a = ha[k * d_k + j], c = hc[k * d_k + j];
        ha[k * d_k + j] = g * c;
        hc[k * d_k + j] = g * a;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < K) {
    double t = 0.0;
    for (int w = 0; w < nwarp; ++w) t += sh[w * K + threadIdx.x];
    ss_part[(int64_t)blockIdx.x * K + threadIdx.x] = t;
  }
}

__global__ void sumsq_finalize_kernel(const double* __restrict__ ss_part, int parts, int K,
                                      float* __restrict__ out, int accumulate) {
  const int k = threadIdx.x;
  if (k >= K) return;
  double t = 0.0;
  for (int p = 0; p < parts; ++p) t += ss_part[(int64_t)p * K + k];
  out[k] = (accumulate ? out[k] : 0.f) + (float)t;
}

// ---------------------------------------------------------------------------------------------------------------
// Small split layers (d <= 64: BASELINE cfg 1, the toy CNN of constants.py:40-51): the whole row pass in ONE kernel.
// A CTA takes 64 rows at a time: stages them and U in shared memory, projects (HA, HC), forms s, g = relu(s), P = g*HC,
// Q = g*HA in place, and accumulates its share of X = A^T P + C^T Q in registers over all its row tiles; per-CTA
// partials are summed in a fixed order afterwards (deterministic).  The five GEMM / row kernels of the general path
// above are latency bound at this size (8 launches, 0.11 ms for 16 000 rows; this kernel: see DESIGN 2.2).
constexpr int kSmallRows = 64, kSmallD = 64, kSmallLd = 65, kSmallLdH = 80;
// Thread (row, jq) of the projection owns the columns 16 v + 4 jq + (0..3), v = 0..3 (float4 chunks interleaved over the
// four threads of a row: their shared-memory accesses are 64 contiguous bytes, and the row stride of 80 floats puts the
// two rows of a quarter-warp on different banks).  d_k % 4 == 0, so a chunk lies inside one concept.
__global__ void __launch_bounds__(256) fused_small_step_kernel(const float* __restrict__ A, const float* __restrict__ C,
                                                               const float* __restrict__ U, int64_t M, int d, int m, int K,
                                                               float* __restrict__ part, double* __restrict__ ss_part) {
  extern __shared__ __align__(16) float sm[];
  float* Us = sm;                                   // [64][64]   U[i][j], zero padded
  float* At = Us + kSmallD * kSmallD;               // [64 rows][65]
  float* Ct = At + kSmallRows * kSmallLd;
  float* Qs = Ct + kSmallRows * kSmallLd;           // [64 rows][80]  Q = g * HA
  float* Ps = Qs + kSmallRows * kSmallLdH;          // [64 rows][80]  P = g * HC
  __shared__ double ssq_sh[4][kSmallRows];
  const int tid = threadIdx.x;
  const int d_k = m / K;
  for (int i = tid; i < kSmallD * kSmallD; i += 256) {
    const int r = i / kSmallD, c = i % kSmallD;
    Us[i] = (r < d && c < m) ? __ldg(U + r * m + c) : 0.f;
  }
  const int row = tid >> 2, jq = tid & 3;            // projections: (row, jq); gradient: (channel i = row, jq)
  int kc[4];                                         // concept of this thread's chunk v (K: chunk beyond m, contributes nothing)
#pragma unroll
  for (int v = 0; v < 4; ++v) kc[v] = (16 * v + 4 * jq < m) ? (16 * v + 4 * jq) / d_k : K;
  float X[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) X[e] = 0.f;
  double ssq[4] = {0.0, 0.0, 0.0, 0.0};              // per concept, kept by the jq == 0 thread of a row
  const int64_t tiles = (M + kSmallRows - 1) / kSmallRows;
  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int64_t r0 = t * kSmallRows;
    __syncthreads();                                 // previous tile fully consumed (also orders the U fill)
    for (int i = tid; i < kSmallRows * kSmallD; i += 256) {
      const int r = i / kSmallD, c = i % kSmallD;
      const bool ok = (r0 + r < M) && c < d;
      At[r * kSmallLd + c] = ok ? __ldg(A + (r0 + r) * d + c) : 0.f;
      Ct[r * kSmallLd + c] = ok ? __ldg(C + (r0 + r) * d + c) : 0.f;
    }
    __syncthreads();
    {
      float ha[16], hc[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) { ha[e] = 0.f; hc[e] = 0.f; }
      for (int i = 0; i < d; ++i) {
        const float a = At[row * kSmallLd + i], c = Ct[row * kSmallLd + i];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float4 u = *reinterpret_cast<const float4*>(Us + i * kSmallD + 16 * v + 4 * jq);
          ha[4 * v] = fmaf(a, u.x, ha[4 * v]); ha[4 * v + 1] = fmaf(a, u.y, ha[4 * v + 1]);
          ha[4 * v + 2] = fmaf(a, u.z, ha[4 * v + 2]); ha[4 * v + 3] = fmaf(a, u.w, ha[4 * v + 3]);
          hc[4 * v] = fmaf(c, u.x, hc[4 * v]); hc[4 * v + 1] = fmaf(c, u.y, hc[4 * v + 1]);
          hc[4 * v + 2] = fmaf(c, u.z, hc[4 * v + 2]); hc[4 * v + 3] = fmaf(c, u.w, hc[4 * v + 3]);
        }
      }
      // s_rk: partial dot products of this thread's chunks, summed over the four threads of the row
      float sk[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const float pdot = fmaf(ha[4 * v], hc[4 * v], fmaf(ha[4 * v + 1], hc[4 * v + 1],
                           fmaf(ha[4 * v + 2], hc[4 * v + 2], ha[4 * v + 3] * hc[4 * v + 3])));
#pragma unroll
        for (int k = 0; k < 4; ++k) sk[k] += (kc[v] == k) ? pdot : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        sk[k] += __shfl_xor_sync(0xffffffffu, sk[k], 1);
        sk[k] += __shfl_xor_sync(0xffffffffu, sk[k], 2);
        sk[k] = fmaxf(sk[k], 0.f);                            // g = relu(s)
        if (jq == 0) ssq[k] += (double)sk[k] * (double)sk[k];
      }
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        float g = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) g = (kc[v] == k) ? sk[k] : g;
        *reinterpret_cast<float4*>(Ps + row * kSmallLdH + 16 * v + 4 * jq) =
            make_float4(g * hc[4 * v], g * hc[4 * v + 1], g * hc[4 * v + 2], g * hc[4 * v + 3]);
        *reinterpret_cast<float4*>(Qs + row * kSmallLdH + 16 * v + 4 * jq) =
            make_float4(g * ha[4 * v], g * ha[4 * v + 1], g * ha[4 * v + 2], g * ha[4 * v + 3]);
      }
    }
    __syncthreads();
    // X[i = row][16 v + 4 jq + ..] += sum_r A[r][i] P[r][j] + C[r][i] Q[r][j]
    for (int r = 0; r < kSmallRows; ++r) {
      const float a = At[r * kSmallLd + row], c = Ct[r * kSmallLd + row];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const float4 pp = *reinterpret_cast<const float4*>(Ps + r * kSmallLdH + 16 * v + 4 * jq);
        const float4 qq = *reinterpret_cast<const float4*>(Qs + r * kSmallLdH + 16 * v + 4 * jq);
        X[4 * v] = fmaf(a, pp.x, fmaf(c, qq.x, X[4 * v]));
        X[4 * v + 1] = fmaf(a, pp.y, fmaf(c, qq.y, X[4 * v + 1]));
        X[4 * v + 2] = fmaf(a, pp.z, fmaf(c, qq.z, X[4 * v + 2]));
        X[4 * v + 3] = fmaf(a, pp.w, fmaf(c, qq.w, X[4 * v + 3]));
      }
    }
  }
  // per-CTA partials
  float* dst = part + (int64_t)blockIdx.x * d * m;
  if (row < d) {
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = 16 * v + 4 * jq + e;
        if (j < m) dst[row * m + j] = X[4 * v + e];
      }
  }
  if (jq == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) ssq_sh[k][row] = ssq[k];
  }
  __syncthreads();
  if (tid < K) {
    double tot = 0.0;
    for (int r = 0; r < kSmallRows; ++r) tot += ssq_sh[tid][r];
    ss_part[(int64_t)blockIdx.x * K + tid] = tot;
  }
}

// sums[e] = sum over the per-CTA partials in a fixed order: 64 elements x 4 part groups per CTA, four independent
// accumulators per thread so that the (L2-resident) loads are in flight together; block 0 also folds the sum-of-squares
// partials.  (The generic reduce_partials walks the parts with one dependent accumulator: 0.1 ms for 250 parts.)
__global__ void __launch_bounds__(256) small_reduce_kernel(const float* __restrict__ part, const double* __restrict__ ss_part,
                                                           int parts, int n, int K, float* __restrict__ sums) {
  __shared__ float red[4][64];
  const int e = blockIdx.x * 64 + (threadIdx.x & 63), pg = threadIdx.x >> 6;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (e < n) {
    int p = pg;
    for (; p + 12 < parts; p += 16) {
      a0 += __ldg(part + (int64_t)p * n + e);
      a1 += __ldg(part + (int64_t)(p + 4) * n + e);
      a2 += __ldg(part + (int64_t)(p + 8) * n + e);
      a3 += __ldg(part + (int64_t)(p + 12) * n + e);
    }
    for (; p < parts; p += 4) a0 += __ldg(part + (int64_t)p * n + e);
  }
  red[pg][threadIdx.x & 63] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (pg == 0 && e < n) sums[e] = (red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]);
  if (blockIdx.x == 0 && threadIdx.x < K) {
    double t = 0.0;
    for (int p = 0; p < parts; ++p) t += ss_part[(int64_t)p * K + threadIdx.x];
    sums[n + threadIdx.x] = (float)t;
  }
}

bool fused_small_ok(int d, int m, int K) {
  return d <= kSmallD && m <= kSmallD && K <= 4 && m % K == 0 && (m / K) % 4 == 0;
}
constexpr int kSmallSmemBytes = (kSmallD * kSmallD + 2 * kSmallRows * kSmallLd + 2 * kSmallRows * kSmallLdH) * 4;
int fused_small_grid(int64_t M) {
  const int64_t tiles = (M + kSmallRows - 1) / kSmallRows;
  return (int)(tiles < 2 * 148 ? tiles : 2 * 148);
}

int splits_for(int d, int m) {
  const int tiles = cdiv(d, 64) * cdiv(m, 64);
  int s = (2 * 148 + tiles - 1) / tiles;
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  return s;
}
}  // namespace

int64_t step_fp32_workspace_bytes(int64_t M, int d, int m, int K) {
  const int64_t rows = M < CHUNK_ROWS ? M : CHUNK_ROWS;
  int64_t b = 0;
  if (fused_small_ok(d, m, K))
    b += align_up((int64_t)2 * 148 * d * m * 4, 256) + align_up((int64_t)2 * 148 * K * 8, 256);
  b += align_up(rows * m * 4, 256) * 2;                       // HA, HC
  b += align_up((int64_t)splits_for(d, m) * d * m * 4, 256);  // X partials
  b += align_up((int64_t)ROW_BLOCKS * K * 8, 256);            // sumsq partials
  return b;
}

int step_fp32(const float* A, const float* C, const float* U, int64_t M, int d, int m, int K,
              float* sums, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  if (workspace_bytes < step_fp32_workspace_bytes(M, d, m, K)) return DRSA_ERR_WORKSPACE;
  if (fused_small_ok(d, m, K)) {
    char* w0 = static_cast<char*>(workspace);
    float* part = reinterpret_cast<float*>(w0);
    double* ssp0 = reinterpret_cast<double*>(w0 + align_up((int64_t)2 * 148 * d * m * 4, 256));
    static bool attr_set = false;
    if (!attr_set) {
      DRSA_CUDA(cudaFuncSetAttribute(fused_small_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallSmemBytes));
      attr_set = true;
    }
    const int grid = fused_small_grid(M);
    fused_small_step_kernel<<<grid, 256, kSmallSmemBytes, stream>>>(A, C, U, M, d, m, K, part, ssp0);
    DRSA_LAUNCH_CHECK();
    small_reduce_kernel<<<cdiv((int64_t)d * m, 64), 256, 0, stream>>>(part, ssp0, grid, d * m, K, sums);
    DRSA_LAUNCH_CHECK();
    return DRSA_OK;
  }
  const int d_k = m / K;
  const int64_t rows_max = M < CHUNK_ROWS ? M : CHUNK_ROWS;
  char* w = static_cast<char*>(workspace);
  float* HA = reinterpret_cast<float*>(w); w += align_up(rows_max * m * 4, 256);
  float* HC = reinterpret_cast<float*>(w); w += align_up(rows_max * m * 4, 256);
  const int splits = splits_for(d, m);
  float* Xp = reinterpret_cast<float*>(w); w += align_up((int64_t)splits * d * m * 4, 256);
  double* ssp = reinterpret_cast<double*>(w);
  float* X = sums;
  float* ss = sums + (int64_t)d * m;

  int chunk = 0;
  for (int64_t r0 = 0; r0 < M; r0 += CHUNK_ROWS, ++chunk) {
    const int64_t rows = (M - r0) < CHUNK_ROWS ? (M - r0) : CHUNK_ROWS;
    const float* Ac = A + r0 * d;
    const float* Cc = C + r0 * d;
    GemmDesc g{};
    g.M = (int)rows; g.N = m; g.K = d; g.lda = d; g.ldb = m; g.ldc = m;
    g.transA = 0; g.transB = 0; g.alpha = 1.f; g.beta = 0.f; g.diag = 0.f; g.splits = 1;
    g.A = Ac; g.B = U; g.C = HA; DRSA_TRY(sgemm(g, stream));
    g.A = Cc; g.C = HC;          DRSA_TRY(sgemm(g, stream));
    row_relevance_kernel<<<ROW_BLOCKS, 256, 8 * K * sizeof(double), stream>>>(HA, HC, rows, m, K, d_k, ssp);
    DRSA_LAUNCH_CHECK();
    sumsq_finalize_kernel<<<1, 32 * cdiv(K, 32), 0, stream>>>(ssp, ROW_BLOCKS, K, ss, chunk > 0);
    DRSA_LAUNCH_CHECK();
    // X (+)= A_c^T HA' + C_c^T HC'   (split over the rows of the chunk)
    GemmDesc t{};
    t.M = d; t.N = m; t.K = rows; t.lda = d; t.ldb = m; t.ldc = m;
    t.transA = 1; t.transB = 0; t.alpha = 1.f; t.splits = splits; t.part_stride = (int64_t)d * m;
    t.A = Ac; t.B = HA; t.C = Xp; DRSA_TRY(sgemm(t, stream));
    DRSA_TRY(reduce_partials(Xp, splits, (int64_t)d * m, (int64_t)d * m, X, chunk > 0, stream));
    t.A = Cc; t.B = HC;           DRSA_TRY(sgemm(t, stream));
    DRSA_TRY(reduce_partials(Xp, splits, (int64_t)d * m, (int64_t)d * m, X, 1, stream));
  }
  return DRSA_OK;
}

}  // namespace drsa
