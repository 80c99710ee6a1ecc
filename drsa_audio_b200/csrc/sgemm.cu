// Generic fp32 CUDA-core GEMM used by the exact (DRSA_PREC_FP32) path, the polar
// retraction and the dense LRP layers.  64x64x16 tiles, 256 threads, 4x4 register
// micro-tiles, operands staged k-major in shared memory so the inner loop reads two
// float4 per k.  Fully bounds-guarded (any M, N, K).
#include "common.cuh"

namespace drsa {

namespace {
constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

template <int TA, int TB>
__global__ void __launch_bounds__(NT) sgemm_kernel(GemmDesc g) {
  if (g.skip_flag != nullptr && *g.skip_flag != 0) return;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int split = blockIdx.z;
  const int64_t kchunk = (g.K + g.splits - 1) / g.splits;
  const int64_t kchunk_al = (kchunk + BK - 1) / BK * BK;
  const int64_t kbeg = (int64_t)split * kchunk_al;
  const int64_t kend = kbeg + kchunk_al < g.K ? kbeg + kchunk_al : g.K;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader coordinates: 64x16 elements = 1024 per operand tile, 4 per thread
  // TA==0: A[m][k], k contiguous -> thread handles (m = tid/4, k = (tid%4)*4 .. +3)
  // TA==1: A[k][m], m contiguous -> thread handles (k = tid/16, m = (tid%16)*4 .. +3)
  float ra[4], rb[4];
  auto load_tiles = [&](int64_t k0) {
    if (TA == 0) {
      const int m = m0 + (tid >> 2); const int64_t k = k0 + (tid & 3) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        ra[i] = (m < g.M && k + i < kend) ? __ldg(g.A + (int64_t)m * g.lda + k + i) : 0.f;
    } else {
      const int64_t k = k0 + (tid >> 4); const int m = m0 + (tid & 15) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        ra[i] = (k < kend && m + i < g.M) ? __ldg(g.A + k * g.lda + m + i) : 0.f;
    }
    if (TB == 0) {
      const int64_t k = k0 + (tid >> 4); const int n = n0 + (tid & 15) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        rb[i] = (k < kend && n + i < g.N) ? __ldg(g.B + k * g.ldb + n + i) : 0.f;
    } else {
      const int n = n0 + (tid >> 2); const int64_t k = k0 + (tid & 3) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        rb[i] = (n < g.N && k + i < kend) ? __ldg(g.B + (int64_t)n * g.ldb + k + i) : 0.f;
    }
  };
  auto store_tiles = [&](int buf) {
    if (TA == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) As[buf][(tid & 3) * 4 + i][tid >> 2] = ra[i];
    } else {
      *reinterpret_cast<float4*>(&As[buf][tid >> 4][(tid & 15) * 4]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
    }
    if (TB == 0) {
      *reinterpret_cast<float4*>(&Bs[buf][tid >> 4][(tid & 15) * 4]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) Bs[buf][(tid & 3) * 4 + i][tid >> 2] = rb[i];
    }
  };

  if (kbeg < kend) {
    load_tiles(kbeg);
    store_tiles(0);
    __syncthreads();
    int buf = 0;
    for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
      const bool more = k0 + BK < kend;
      if (more) load_tiles(k0 + BK);
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      if (more) {
        store_tiles(buf ^ 1);
        __syncthreads();
        buf ^= 1;
      }
    }
  }

  float* Cout = g.C + (g.splits > 1 ? (int64_t)split * g.part_stride : 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = g.alpha * acc[i][j];
      if (g.splits == 1) {
        if (g.beta != 0.f) v += g.beta * Cout[(int64_t)m * g.ldc + n];
        if (m == n) v += g.diag;
      }
      Cout[(int64_t)m * g.ldc + n] = v;
    }
  }
}

__global__ void reduce_partials_kernel(const float* __restrict__ in, int parts, int64_t stride,
                                       int64_t count, float* __restrict__ out, int accumulate) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    float s = accumulate ? out[i] : 0.f;
    for (int p = 0; p < parts; ++p) s += in[(int64_t)p * stride + i];
    out[i] = s;
  }
}
}  // namespace

int sgemm(const GemmDesc& g, cudaStream_t stream) {
  if (g.M <= 0 || g.N <= 0 || g.K < 0 || g.splits < 1) return DRSA_ERR_ARG;
  dim3 grid(cdiv(g.N, BN), cdiv(g.M, BM), g.splits);
  if (g.transA == 0 && g.transB == 0) sgemm_kernel<0, 0><<<grid, NT, 0, stream>>>(g);
  else if (g.transA == 1 && g.transB == 0) sgemm_kernel<1, 0><<<grid, NT, 0, stream>>>(g);
  else if (g.transA == 0 && g.transB == 1) sgemm_kernel<0, 1><<<grid, NT, 0, stream>>>(g);
  else sgemm_kernel<1, 1><<<grid, NT, 0, stream>>>(g);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

int reduce_partials(const float* in, int parts, int64_t stride, int64_t count, float* out,
                    int accumulate, cudaStream_t stream) {
  int blocks = cdiv(count, 256);
  if (blocks > 4 * 148) blocks = 4 * 148;
  reduce_partials_kernel<<<blocks, 256, 0, stream>>>(in, parts, stride, count, out, accumulate);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

}  // namespace drsa
