// Everything after the row pass of a DRSA step: pooling scalars, the ascent step and the
// polar (Loewdin) retraction U <- Y (Y^T Y)^(-1/2)  (reference: drsa.py:102, :201-221,
// :224-238).  The reference moves Y^T Y to the host and calls an fp64 eigh every step
// (drsa.py:216); here the polar factor is computed on the device with a scaled
// Newton-Schulz iteration X <- X (1.5 I - 0.5 X^T X), X_0 = Y / sqrt(||Y^T Y||_inf), which
// converges quadratically to the same unique polar factor.  No host synchronisation:
// convergence is tracked in a device flag that turns the remaining sweeps into no-ops.
#include "common.cuh"

namespace drsa {

namespace {

struct PolarWs {
  float* Y; float* X0; float* X1; float* G; float* scal; int* flags;  // flags: done, final_buf, sweeps
};

int64_t polar_ws_bytes(int d, int m) {
  return align_up((int64_t)d * m * 4, 256) * 3 + align_up((int64_t)m * m * 4, 256) + 256 + 256;
}

PolarWs carve(void* workspace, int d, int m) {
  char* w = static_cast<char*>(workspace);
  PolarWs p;
  const int64_t dm = align_up((int64_t)d * m * 4, 256);
  p.Y = reinterpret_cast<float*>(w); w += dm;
  p.X0 = reinterpret_cast<float*>(w); w += dm;
  p.X1 = reinterpret_cast<float*>(w); w += dm;
  p.G = reinterpret_cast<float*>(w); w += align_up((int64_t)m * m * 4, 256);
  p.scal = reinterpret_cast<float*>(w); w += 256;
  p.flags = reinterpret_cast<int*>(w);
  return p;
}

// q_k, obj, coef_k from the K sums of squares; Y = U + coef_k * X_k; obj_log[idx] = obj.
__global__ void __launch_bounds__(256) ascent_kernel(const float* __restrict__ sums, double inv_M,
                                                     const float* __restrict__ U, int d, int m, int K,
                                                     float* __restrict__ Y, float* __restrict__ obj_log,
                                                     int64_t log_index, int* __restrict__ status) {
  extern __shared__ float coef[];   // [K]
  const int d_k = m / K;
  if (threadIdx.x == 0) {
    // fp32 like the reference: q = sqrt(mean r^2) (drsa.py:236), obj = (mean sqrt q)^2 (:237)
    float acc = 0.f; int degenerate = 0;
    for (int k = 0; k < K; ++k) {
      const float q = sqrtf((float)((double)sums[(int64_t)d * m + k] * inv_M));
      if (q == 0.f) ++degenerate;
      acc += sqrtf(q);
    }
    const float root = acc / (float)K;       // sqrt(obj)
    const float obj = root * root;
    for (int k = 0; k < K; ++k) {
      const float q = sqrtf((float)((double)sums[(int64_t)d * m + k] * inv_M));
      // d obj / d s_mk = coef_k * relu(s_mk),  coef_k = sqrt(obj) / (K M q_k^1.5)
      coef[k] = (float)((double)root * inv_M / ((double)K * (double)q * sqrt((double)q)));
    }
    if (blockIdx.x == 0) {
      if (obj_log != nullptr) {
        int64_t idx = log_index;
        if (idx < 0) { idx = status[3]; status[3] = (int)idx + 1; }   // append mode (graph replay)
        obj_log[idx] = obj;
      }
      if (status != nullptr) status[2] = degenerate;
    }
  }
  __syncthreads();
  if (Y == nullptr) return;
  const int64_t total = (int64_t)d * m;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % m);
    Y[i] = U[i] + coef[j / d_k] * sums[i];
  }
}

// scal[0] = 1 / sqrt(max_i sum_j |G_ij|)  (upper bound of sigma_max(Y)); resets the flags.
__global__ void __launch_bounds__(1024) gram_norm_kernel(const float* __restrict__ G, int m,
                                                         float* __restrict__ scal, int* __restrict__ flags) {
  __shared__ float red[32];
  float best = 0.f;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int i = warp; i < m; i += nwarp) {
    float s = 0.f;
    for (int j = lane; j < m; j += 32) s += fabsf(G[(int64_t)i * m + j]);
    s = warp_sum(s);
    best = fmaxf(best, s);
  }
  if (lane == 0) red[warp] = best;
  __syncthreads();
  if (warp == 0) {
    float v = lane < nwarp ? red[lane] : 0.f;
    v = warp_max(v);
    if (lane == 0) {
      scal[0] = rsqrtf(v);
      flags[0] = 0; flags[1] = 0; flags[2] = 0;
    }
  }
}

__global__ void scale_kernel(const float* __restrict__ Y, const float* __restrict__ scal, int64_t n,
                             float* __restrict__ X) {
  const float s = scal[0];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    X[i] = Y[i] * s;
}

// Residual ||G - I||_F of sweep `it`; on convergence latch (done, buffer index, sweeps),
// otherwise turn G into the Newton-Schulz multiplier T = 1.5 I - 0.5 G in place.
__global__ void __launch_bounds__(1024) residual_kernel(float* __restrict__ G, int m, float tol2_m,
                                                        int it, int* __restrict__ flags) {
  if (flags[0] != 0) return;
  __shared__ float red[32];
  __shared__ int converged;
  const int64_t total = (int64_t)m * m;
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
    const int r = (int)(i / m), c = (int)(i % m);
    const float e = G[i] - (r == c ? 1.f : 0.f);
    s = fmaf(e, e, s);
  }
  s = warp_sum(s);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (warp == 0) {
    float v = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
    v = warp_sum(v);
    if (lane == 0) {
      converged = (v < tol2_m) ? 1 : 0;
      if (converged) { flags[1] = it & 1; flags[2] = it; }
    }
  }
  __syncthreads();
  if (converged) {
    __syncthreads();
    if (threadIdx.x == 0) flags[0] = 1;
    return;
  }
  for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
    const int r = (int)(i / m), c = (int)(i % m);
    G[i] = (r == c ? 1.5f : 0.f) - 0.5f * G[i];
  }
}

// U_out = the converged buffer; optional fp16 hi/lo split of U_out^T for the tensor-core step.
__global__ void __launch_bounds__(256) select_kernel(const float* __restrict__ X0, const float* __restrict__ X1,
                                                     int d, int m, int max_iters, const int* __restrict__ flags,
                                                     float* __restrict__ U_out, __half* __restrict__ Ut_hi,
                                                     __half* __restrict__ Ut_lo, int* __restrict__ status) {
  const int done = flags[0];
  const int buf = done ? flags[1] : (max_iters & 1);
  const float* X = buf ? X1 : X0;
  if (blockIdx.x == 0 && threadIdx.x == 0 && status != nullptr) {
    status[0] = done ? flags[2] : max_iters;
    status[1] = done ? 0 : 1;
  }
  const int64_t total = (int64_t)d * m;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float v = X[i];
    U_out[i] = v;
    if (Ut_hi != nullptr) {
      const int r = (int)(i / m), c = (int)(i % m);
      const __half hi = __float2half_rn(v);
      Ut_hi[(int64_t)c * d + r] = hi;
      if (Ut_lo != nullptr) Ut_lo[(int64_t)c * d + r] = __float2half_rn(v - __half2float(hi));
    }
  }
}

__global__ void split_u_kernel(const float* __restrict__ U, int d, int m, __half* __restrict__ Ut_hi,
                               __half* __restrict__ Ut_lo) {
  const int64_t total = (int64_t)d * m;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / m), c = (int)(i % m);
    const float v = U[i];
    const __half hi = __float2half_rn(v);
    Ut_hi[(int64_t)c * d + r] = hi;
    if (Ut_lo != nullptr) Ut_lo[(int64_t)c * d + r] = __float2half_rn(v - __half2float(hi));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// QR retraction (NON-DEFAULT option, BASELINE.json north_star (3)): Q of the thin QR factorisation Y = QR with diag(R) > 0
// (what a Householder QR with sign fix returns; the factorisation with positive diagonal is unique).  The reference's
// retraction is the POLAR factor (drsa.py:201-221) and the two differ by 7.6 % per step: trajectories obtained
// with this option do not follow the reference.  One CTA (a cluster of one), classical Gram-Schmidt with
// re-orthogonalisation (CGS2, as accurate as Householder for this purpose), Q kept transposed [m][d] so that every
// column is contiguous: per column two rounds of { h = Q_<j^T v ; v -= Q_<j h }, then v / ||v||.
__global__ void __launch_bounds__(1024) qr_retract_kernel(const float* __restrict__ Y, int d, int m, float* __restrict__ Qt,
                                                          float* __restrict__ U_out, __half* __restrict__ Ut_hi,
                                                          __half* __restrict__ Ut_lo, int* __restrict__ status) {
  extern __shared__ float sh[];          // v[d] | h[m] | red[32]
  float* v = sh;
  float* h = sh + d;
  float* red = sh + d + m;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  int rank_deficient = 0;
  for (int j = 0; j < m; ++j) {
    for (int t = tid; t < d; t += blockDim.x) v[t] = Y[(int64_t)t * m + j];
    __syncthreads();
    for (int round = 0; round < 2; ++round) {
      for (int i = warp; i < j; i += nwarp) {                    // h_i = q_i . v
        const float* q = Qt + (int64_t)i * d;
        float s = 0.f;
        for (int t = lane; t < d; t += 32) s = fmaf(q[t], v[t], s);
        s = warp_sum(s);
        if (lane == 0) h[i] = s;
      }
      __syncthreads();
      for (int t = tid; t < d; t += blockDim.x) {                // v -= sum_i h_i q_i   (fixed order: deterministic)
        float a = v[t];
        for (int i = 0; i < j; ++i) a = fmaf(-h[i], Qt[(int64_t)i * d + t], a);
        v[t] = a;
      }
      __syncthreads();
    }
    float s = 0.f;
    for (int t = tid; t < d; t += blockDim.x) s = fmaf(v[t], v[t], s);
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    float nrm2 = 0.f;
    for (int w = 0; w < nwarp; ++w) nrm2 += red[w];
    if (!(nrm2 > 1e-30f)) rank_deficient = 1;
    const float inv = nrm2 > 1e-30f ? rsqrtf(nrm2) : 0.f;         // R_jj = ||v|| > 0: the sign fix is built in
    for (int t = tid; t < d; t += blockDim.x) Qt[(int64_t)j * d + t] = v[t] * inv;
    __syncthreads();                                               // also publishes Qt[j] (global) to the whole CTA
  }
  __threadfence_block();
  for (int64_t i = tid; i < (int64_t)d * m; i += blockDim.x) {
    const int r = (int)(i / m), c = (int)(i % m);
    const float q = Qt[(int64_t)c * d + r];
    U_out[i] = q;
    if (Ut_hi != nullptr) {
      const __half hi = __float2half_rn(q);
      Ut_hi[(int64_t)c * d + r] = hi;
      if (Ut_lo != nullptr) Ut_lo[(int64_t)c * d + r] = __float2half_rn(q - __half2float(hi));
    }
  }
  if (tid == 0 && status != nullptr) { status[0] = 0; if (rank_deficient) status[1] += 1; }
}

int qr_from_Y(const float* Y, int d, int m, float* Qt, float* U_out, void* Ut_hi, void* Ut_lo, int* status,
              cudaStream_t stream) {
  const int smem = (d + m + 32) * 4;
  qr_retract_kernel<<<1, 1024, smem, stream>>>(Y, d, m, Qt, U_out, static_cast<__half*>(Ut_hi), static_cast<__half*>(Ut_lo),
                                                status);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

int polar_from_Y(const PolarWs& p, int d, int m, float* U_out, void* Ut_hi, void* Ut_lo, int max_iters,
                 float tol, int* status, cudaStream_t stream) {
  GemmDesc gram{};
  gram.M = m; gram.N = m; gram.K = d; gram.lda = m; gram.ldb = m; gram.ldc = m;
  gram.transA = 1; gram.transB = 0; gram.alpha = 1.f; gram.splits = 1;
  gram.A = p.Y; gram.B = p.Y; gram.C = p.G;
  DRSA_TRY(sgemm(gram, stream));
  gram_norm_kernel<<<1, 1024, 0, stream>>>(p.G, m, p.scal, p.flags);
  DRSA_LAUNCH_CHECK();
  const int64_t n = (int64_t)d * m;
  const int eb = cdiv(n, 256) < 592 ? cdiv(n, 256) : 592;
  scale_kernel<<<eb, 256, 0, stream>>>(p.Y, p.scal, n, p.X0);
  DRSA_LAUNCH_CHECK();
  const float tol2_m = tol * tol * (float)m;
  for (int it = 0; it < max_iters; ++it) {
    float* cur = (it & 1) ? p.X1 : p.X0;
    float* nxt = (it & 1) ? p.X0 : p.X1;
    gram.A = cur; gram.B = cur; gram.C = p.G; gram.skip_flag = p.flags;
    DRSA_TRY(sgemm(gram, stream));
    residual_kernel<<<1, 1024, 0, stream>>>(p.G, m, tol2_m, it, p.flags);
    DRSA_LAUNCH_CHECK();
    GemmDesc mul{};
    mul.M = d; mul.N = m; mul.K = m; mul.lda = m; mul.ldb = m; mul.ldc = m;
    mul.transA = 0; mul.transB = 0; mul.alpha = 1.f; mul.splits = 1;
    mul.A = cur; mul.B = p.G; mul.C = nxt; mul.skip_flag = p.flags;
    DRSA_TRY(sgemm(mul, stream));
  }
  select_kernel<<<eb, 256, 0, stream>>>(p.X0, p.X1, d, m, max_iters, p.flags, U_out,
                                        static_cast<__half*>(Ut_hi), static_cast<__half*>(Ut_lo), status);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}
}  // namespace

bool finish_fused_supported(int d, int m, int K);
int64_t finish_fused_workspace_bytes(int d, int m);
int finish_fused(const float* sums, int64_t M_global, const float* U, int d, int m, int K, float* U_out, void* Ut_hi,
                 void* Ut_lo, float* obj_log, int64_t log_index, int max_iters, float tol, int u_rounded, int* status,
                 void* workspace, int64_t workspace_bytes, const float* Y_in, const drsa_peer_exchange* px,
                 cudaStream_t stream);

int64_t finish_workspace_bytes(int d, int m) {
  const int64_t a = polar_ws_bytes(d, m), b = finish_fused_workspace_bytes(d, m);
  return a > b ? a : b;
}

int finish_step(const float* sums, int64_t M_global, const float* U, int d, int m, int K, float* U_out,
                void* Ut_hi, void* Ut_lo, float* obj_log, int64_t log_index, int max_iters, float tol,
                int u_rounded, int* status, void* workspace, int64_t workspace_bytes, const drsa_peer_exchange* px,
                cudaStream_t stream) {
  if (finish_fused_supported(d, m, K))       // one cooperative kernel instead of ~30 dependent launches
    return finish_fused(sums, M_global, U, d, m, K, U_out, Ut_hi, Ut_lo, obj_log, log_index, max_iters, tol, u_rounded,
                        status, workspace, workspace_bytes, nullptr, px, stream);
  if (px != nullptr && px->world > 1) return DRSA_ERR_SHAPE;      // the peer exchange lives in the fused kernel
  if (u_rounded) return DRSA_ERR_SHAPE;      // the tensor-core modes only exist for shapes the fused kernel covers
  if (workspace_bytes < polar_ws_bytes(d, m)) return DRSA_ERR_WORKSPACE;
  PolarWs p = carve(workspace, d, m);
  const int64_t n = (int64_t)d * m;
  const int eb = U_out == nullptr ? 1 : (cdiv(n, 256) < 592 ? cdiv(n, 256) : 592);
  ascent_kernel<<<eb, 256, K * sizeof(float), stream>>>(sums, 1.0 / (double)M_global, U, d, m, K,
                                                        U_out == nullptr ? nullptr : p.Y, obj_log,
                                                        log_index, status);
  DRSA_LAUNCH_CHECK();
  if (U_out == nullptr) return DRSA_OK;
  return polar_from_Y(p, d, m, U_out, Ut_hi, Ut_lo, max_iters, tol, status, stream);
}

int polar_retract(const float* Y, int d, int m, float* U_out, int max_iters, float tol, int* status,
                  void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  if (finish_fused_supported(d, m, 1))
    return finish_fused(nullptr, 1, nullptr, d, m, 1, U_out, nullptr, nullptr, nullptr, 0, max_iters, tol, 0, status,
                        workspace, workspace_bytes, Y, nullptr, stream);
  if (workspace_bytes < polar_ws_bytes(d, m)) return DRSA_ERR_WORKSPACE;
  PolarWs p = carve(workspace, d, m);
  DRSA_CUDA(cudaMemcpyAsync(p.Y, Y, (int64_t)d * m * 4, cudaMemcpyDeviceToDevice, stream));
  return polar_from_Y(p, d, m, U_out, nullptr, nullptr, max_iters, tol, status, stream);
}

// The finish step with the QR retraction instead of the polar factor (non-default; fp32 arithmetic, single rank).
int finish_step_qr(const float* sums, int64_t M_global, const float* U, int d, int m, int K, float* U_out, float* obj_log,
                   int64_t log_index, int* status, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  if (workspace_bytes < polar_ws_bytes(d, m)) return DRSA_ERR_WORKSPACE;
  if ((d + m + 32) * 4 > 48 * 1024) return DRSA_ERR_SHAPE;
  PolarWs p = carve(workspace, d, m);
  const int64_t n = (int64_t)d * m;
  const int eb = U_out == nullptr ? 1 : (cdiv(n, 256) < 592 ? cdiv(n, 256) : 592);
  ascent_kernel<<<eb, 256, K * sizeof(float), stream>>>(sums, 1.0 / (double)M_global, U, d, m, K,
                                                        U_out == nullptr ? nullptr : p.Y, obj_log, log_index, status);
  DRSA_LAUNCH_CHECK();
  if (U_out == nullptr) return DRSA_OK;
  return qr_from_Y(p.Y, d, m, p.X0, U_out, nullptr, nullptr, status, stream);
}

int qr_retract(const float* Y, int d, int m, float* U_out, int* status, void* workspace, int64_t workspace_bytes,
               cudaStream_t stream) {
  if (workspace_bytes < polar_ws_bytes(d, m)) return DRSA_ERR_WORKSPACE;
  if ((d + m + 32) * 4 > 48 * 1024) return DRSA_ERR_SHAPE;
  PolarWs p = carve(workspace, d, m);
  return qr_from_Y(Y, d, m, p.X0, U_out, nullptr, nullptr, status, stream);
}

int split_u(const float* U, int d, int m, void* Ut_hi, void* Ut_lo, cudaStream_t stream) {
  const int64_t n = (int64_t)d * m;
  const int eb = cdiv(n, 256) < 592 ? cdiv(n, 256) : 592;
  split_u_kernel<<<eb, 256, 0, stream>>>(U, d, m, static_cast<__half*>(Ut_hi), static_cast<__half*>(Ut_lo));
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

}  // namespace drsa
