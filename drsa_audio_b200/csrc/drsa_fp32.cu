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
  b += align_up(rows * m * 4, 256) * 2;                       // HA, HC
  b += align_up((int64_t)splits_for(d, m) * d * m * 4, 256);  // X partials
  b += align_up((int64_t)ROW_BLOCKS * K * 8, 256);            // sumsq partials
  return b;
}

int step_fp32(const float* A, const float* C, const float* U, int64_t M, int d, int m, int K,
              float* sums, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  if (workspace_bytes < step_fp32_workspace_bytes(M, d, m, K)) return DRSA_ERR_WORKSPACE;
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
