// Shared helpers for libdrsa_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/drsa_b200.h"

namespace drsa {

extern thread_local int g_last_cuda_error;

inline int cuda_fail(cudaError_t e) {
  g_last_cuda_error = (int)e;
  return DRSA_ERR_CUDA;
}

#define DRSA_CUDA(expr)                                   \
  do {                                                    \
    cudaError_t _e = (expr);                              \
    if (_e != cudaSuccess) return ::drsa::cuda_fail(_e);  \
  } while (0)

#define DRSA_LAUNCH_CHECK() DRSA_CUDA(cudaGetLastError())

#define DRSA_TRY(expr)          \
  do {                          \
    int _s = (expr);            \
    if (_s != DRSA_OK) return _s; \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Device must be sm_100 (B200); cached per device.
int require_sm100();
int sm_count();

// ---------------------------------------------------------------- generic fp32 GEMM (CUDA cores)
// C[M,N] = alpha * op(A)[M,K] op(B)[K,N] + beta * C (+ diag on the diagonal)
//   transA == 0: A is [M,K] row-major (lda);  transA == 1: A is [K,M] row-major (lda)
//   transB == 0: B is [K,N] row-major (ldb);  transB == 1: B is [N,K] row-major (ldb)
// splits > 1: the K range is cut into `splits` slices and slice s writes its own
//   partial product to C + s*part_stride (alpha applied, beta/diag ignored); the caller
//   reduces the partials in a fixed order (deterministic).
// skip_flag: optional device int; the kernel returns immediately if *skip_flag != 0.
struct GemmDesc {
  const float* A; const float* B; float* C;
  int M, N; int64_t K;
  int64_t lda, ldb, ldc;
  int transA, transB;
  float alpha, beta, diag;
  int splits; int64_t part_stride;
  const int* skip_flag;
};
int sgemm(const GemmDesc& g, cudaStream_t stream);

// out[i] = sum_{s<parts} in[s*stride + i] (+ out[i] if accumulate), fixed order.
int reduce_partials(const float* in, int parts, int64_t stride, int64_t count, float* out,
                    int accumulate, cudaStream_t stream);

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// zennit stabilize(): x + ((x == 0) + sign(x)) * eps  -- zero counts as positive
__device__ __forceinline__ float stabilize(float x, float eps) {
  return x + (x >= 0.f ? eps : -eps);
}

}  // namespace drsa
