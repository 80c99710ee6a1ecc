// Concept-conditional relevance at the split layer: the reference inserts three modules after the layer where U was
// optimised (ProjectionModel, cxai/model/modify_model.py:4-123) -- Projection h = a U, SubspaceFilter (identity whose
// backward hook masks the relevance of clone k to concept k, SubspaceHook, cxai/xai/explain/attribute.py:12-67) and
// InvProjection a' = h U^T -- and attributes a batch in which every sample is repeated K+1 times
// (explainer.py:68-123, get_class_composite :186-203: Epsilon on both projections).  All clones of a sample share the
// forward pass and the backward pass above the filter; only the masked relevance below it differs.  These kernels
// compute, for position vectors stored as rows (NHWC view, leading dimension ld >= d, padded columns zero):
//
//   forward :  h = a U                [P, m]          a' = h U^T                      [P, ld]
//   backward:  s = R / stab(a')       (Epsilon on InvProjection)      R_h = h * (s U)
//              v = R_h / stab(h)      (Epsilon on Projection)
//              out_k = a * (v[:, block k] U[:, block k]^T)   k = 1..K      (SubspaceHook: clone k keeps concept k)
//              out_0 = sum_k out_k                                          (clone 0: unmasked = standard relevance)
//
// The contractions are P x d x m GEMMs on the CUDA-core SGEMM (fp32 like the reference; P = batch * positions is
// small compared with the convolutions around it).
#include "common.cuh"

namespace drsa {

namespace {
inline int eblk(int64_t n) {
  int64_t b = (n + 255) / 256;
  return (int)(b > 148 * 8 ? 148 * 8 : (b < 1 ? 1 : b));
}

// s[r, c] = R[r, c] / stabilize(z[r, c], eps) for c < d, 0 in the padded columns
__global__ void sf_ratio_kernel(const float* __restrict__ R, const float* __restrict__ z, int64_t P, int d, int ld, float eps,
                                float* __restrict__ s) {
  const int64_t total = P * ld;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    s[i] = ((int)(i % ld) < d) ? __ldg(R + i) / stabilize(__ldg(z + i), eps) : 0.f;
}

// v = h * t / stabilize(h, eps)   (in place over t)
__global__ void sf_mulratio_kernel(const float* __restrict__ h, float* __restrict__ t, int64_t n, float eps) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float hv = __ldg(h + i);
    t[i] = hv * t[i] / stabilize(hv, eps);
  }
}

// out_k = a * w (in place over w); out_0 (+)= out_k
__global__ void sf_mulacc_kernel(const float* __restrict__ a, float* __restrict__ w, float* __restrict__ acc, int64_t n,
                                 int first) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = __ldg(a + i) * w[i];
    w[i] = v;
    acc[i] = first ? v : acc[i] + v;
  }
}

__global__ void planes_to_f32_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, int64_t n,
                                     float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __half2float(hi[i]) + __half2float(lo[i]);
}
}  // namespace

int64_t subspace_filter_workspace_bytes(int64_t P, int d, int m) {
  (void)d;
  return align_up(P * (int64_t)m * 4, 256) + 256;
}

// h [P, m] = a [P, ld] U [d, m];  a2 [P, ld] = h U^T (padded columns zero)
int subspace_project(const float* a, const float* U, int64_t P, int d, int m, int ld, float* h, float* a2, cudaStream_t s) {
  if (P > 2147483647LL) return DRSA_ERR_SHAPE;
  GemmDesc g{};
  g.A = a; g.B = U; g.C = h; g.M = (int)P; g.N = m; g.K = d; g.lda = ld; g.ldb = m; g.ldc = m;
  g.transA = 0; g.transB = 0; g.alpha = 1.f; g.splits = 1;
  DRSA_TRY(sgemm(g, s));
  if (a2 != nullptr) {
    if (ld > d) DRSA_CUDA(cudaMemsetAsync(a2, 0, (size_t)P * ld * 4, s));
    GemmDesc b{};
    b.A = h; b.B = U; b.C = a2; b.M = (int)P; b.N = d; b.K = m; b.lda = m; b.ldb = m; b.ldc = ld;
    b.transA = 0; b.transB = 1; b.alpha = 1.f; b.splits = 1;
    DRSA_TRY(sgemm(b, s));
  }
  return DRSA_OK;
}

// out [(K+1), P, ld]: slot 0 = standard relevance, slot k = relevance routed through concept k only
int subspace_filter_backward(const float* a, const float* h, const float* a2, const float* R, const float* U, int64_t P, int d,
                             int m, int K, int ld, float eps_inv, float eps_proj, float* out, void* workspace,
                             int64_t workspace_bytes, cudaStream_t s) {
  if (P > 2147483647LL || m % K != 0) return DRSA_ERR_SHAPE;
  if (workspace_bytes < subspace_filter_workspace_bytes(P, d, m)) return DRSA_ERR_WORKSPACE;
  float* t = static_cast<float*>(workspace);                 // [P, m]
  const int64_t plane = P * (int64_t)ld;
  float* sbuf = out + plane;                                 // slot 1 doubles as the buffer of s (consumed before it is written)
  sf_ratio_kernel<<<eblk(plane), 256, 0, s>>>(R, a2, P, d, ld, eps_inv, sbuf);
  DRSA_LAUNCH_CHECK();
  GemmDesc g{};
  g.A = sbuf; g.B = U; g.C = t; g.M = (int)P; g.N = m; g.K = d; g.lda = ld; g.ldb = m; g.ldc = m;
  g.transA = 0; g.transB = 0; g.alpha = 1.f; g.splits = 1;
  DRSA_TRY(sgemm(g, s));                                     // t = s U
  sf_mulratio_kernel<<<eblk(P * (int64_t)m), 256, 0, s>>>(h, t, P * (int64_t)m, eps_proj);   // v = h t / stab(h)
  DRSA_LAUNCH_CHECK();
  const int d_k = m / K;
  for (int k = 1; k <= K; ++k) {
    float* w = out + (int64_t)k * plane;
    if (ld > d) DRSA_CUDA(cudaMemsetAsync(w, 0, (size_t)plane * 4, s));
    GemmDesc b{};
    b.A = t + (int64_t)(k - 1) * d_k; b.B = U + (int64_t)(k - 1) * d_k; b.C = w;
    b.M = (int)P; b.N = d; b.K = d_k; b.lda = m; b.ldb = m; b.ldc = ld;
    b.transA = 0; b.transB = 1; b.alpha = 1.f; b.splits = 1;
    DRSA_TRY(sgemm(b, s));                                   // w = v[:, block k] U[:, block k]^T
    sf_mulacc_kernel<<<eblk(plane), 256, 0, s>>>(a, w, out, plane, k == 1 ? 1 : 0);
    DRSA_LAUNCH_CHECK();
  }
  return DRSA_OK;
}

int planes_to_f32(const void* hi, const void* lo, int64_t count, float* out, cudaStream_t s) {
  planes_to_f32_kernel<<<eblk(count), 256, 0, s>>>(static_cast<const __half*>(hi), static_cast<const __half*>(lo), count, out);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

}  // namespace drsa
