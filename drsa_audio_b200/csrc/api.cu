// extern "C" surface of libdrsa_b200.so (see include/drsa_b200.h).  Argument checking lives
// here; the kernels live in the other translation units.
#include <string.h>
#include "common.cuh"

namespace drsa {
thread_local int g_last_cuda_error = 0;

int require_sm100() {
  static int cached = 1;   // 1 = unknown
  if (cached == 1) {
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
      cudaGetLastError();
      return DRSA_ERR_ARCH;
    }
    cached = (prop.major == 10) ? DRSA_OK : DRSA_ERR_ARCH;
  }
  return cached;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
      n = v;
    else { cudaGetLastError(); return 148; }
  }
  return n;
}

// defined in the other translation units
int64_t step_fp32_workspace_bytes(int64_t M, int d, int m, int K);
int step_fp32(const float* A, const float* C, const float* U, int64_t M, int d, int m, int K, float* sums,
              void* workspace, int64_t workspace_bytes, cudaStream_t stream);
bool tc_shape_supported(int d, int m, int K);
int64_t step_tc_workspace_bytes(int64_t M, int d, int m, int K);
int step_tc(const void* A16, const void* C16, const void* Ut_hi, const void* Ut_lo, int64_t M, int d, int m, int K,
            float scaleA, float scaleC, float pq_scale, bool split_u, bool split_ac, float* sums, void* workspace,
            int64_t workspace_bytes, cudaStream_t stream);
int rownorm_max(const float* in, int64_t rows, int d, float* out, cudaStream_t stream);
int conv3x3_forward(const float* x, const float* w, const float* b, int64_t N, int Cin, int Cout, int H, int W, int relu,
                    float* y, cudaStream_t s);
int conv3x3_backward(const float* x, const float* w_mod, const float* wt_mod, const float* b_mod, const float* R_out,
                     int64_t N, int Cin, int Cout, int H, int W, float eps, int x_is_ones, float* s_buf, float* R_in,
                     cudaStream_t s);
int flip_weights(const float* w, int Cout, int Cin, float* wt, cudaStream_t s);
int maxpool_forward(const float* x, int64_t NC, int H, int W, int kh, int kw, float* y, int32_t* argmax, cudaStream_t s);
int maxpool_backward(const float* R_out, const int32_t* argmax, int64_t NC, int H, int W, int kh, int kw, float* R_in,
                     cudaStream_t s);
int dense_forward(const float* x, const float* w, const float* b, int64_t N, int In, int Out, int relu, float* y,
                  cudaStream_t s);
int dense_epsilon_backward(const float* x, const float* w, const float* b, const float* R_out, int64_t N, int In,
                           int Out, float eps, float* s_buf, float* R_in, cudaStream_t s);
int relu_mask(const float* a, float* R, int64_t count, cudaStream_t s);
bool conv_tc_supported(int64_t B, int Cin_p, int Cout_p, int H, int W);
int conv_tc_forward(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias, int64_t B,
                    int H, int W, int Cin_p, int Cout_p, int Cout, int relu, void* y_hi, void* y_lo, float* y_nchw,
                    int* err_flag, cudaStream_t stream);
int conv_first_nhwc(const float* x, const float* w, const float* b, int64_t B, int H, int W, int Cout, int Cout_p,
                    int relu, void* y_hi, void* y_lo, cudaStream_t stream);
int conv_tc_run(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias, int64_t B,
                int H, int W, int Cin_p, int Cout_p, int Cout, int relu, int epi, float eps, const float* aux_f32,
                const void* aux_hi, const void* aux_lo, void* y_hi, void* y_lo, float* y_f32, float* y_nchw,
                const float* scale_ref, float* cmax_out, int* err_flag, cudaStream_t stream, int pkh, int pkw,
                void* amax_out);
bool conv_tc_pool_supported(int64_t B, int Cin_p, int Cout_p, int H, int W, int kh, int kw);
int conv_tc_forward_pool(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias, int64_t B,
                         int H, int W, int Cin_p, int Cout_p, int Cout, int relu, int kh, int kw, void* y_hi, void* y_lo,
                         void* argmax_u8, int* err_flag, cudaStream_t stream);
int first_layer_ones_backward(const float* R_out, const float* w_mod, const float* b_mod, int64_t B, int H, int W, int Cout,
                              int Cp, float eps, float* R_in, cudaStream_t stream);
int64_t subspace_filter_workspace_bytes(int64_t P, int d, int m);
int subspace_project(const float* a, const float* U, int64_t P, int d, int m, int ld, float* h, float* a2, cudaStream_t s);
int subspace_filter_backward(const float* a, const float* h, const float* a2, const float* R, const float* U, int64_t P, int d,
                             int m, int K, int ld, float eps_inv, float eps_proj, float* out, void* workspace,
                             int64_t workspace_bytes, cudaStream_t s);
int planes_to_f32(const void* hi, const void* lo, int64_t count, float* out, cudaStream_t s);
int64_t logmel_workspace_bytes(int64_t B, int n_fft, int n_mels, int width);
int logmel_transform(const float* wav, const float* window, const float* basis, const float* fb, int64_t B, int64_t n_samples,
                     int n_fft, int hop, int n_mels, int first_frame, int width, int do_clamp, float clamp_min, float* out,
                     void* workspace, int64_t workspace_bytes, cudaStream_t s);
int sample_absmax_ratio(const float* R, const float* x, int64_t B, int64_t per, float* out, cudaStream_t stream);
int maxpool_nhwc(const void* x_hi, const void* x_lo, int64_t B, int H, int W, int Cp, int kh, int kw, void* y_hi,
                 void* y_lo, void* argmax_u8, cudaStream_t stream);
int maxpool_nhwc_backward(const float* R_out, const void* argmax_u8, int64_t B, int H, int W, int Cp, int kh, int kw,
                          float* R_in, cudaStream_t stream);
int nhwc_f32_to_nchw(const float* x, int64_t B, int H, int W, int Cp, int C, float* y, cudaStream_t stream);
int nchw_to_nhwc_f32(const float* x, int64_t B, int H, int W, int C, int Cp, float* y, cudaStream_t stream);
int nhwc_to_nchw(const void* x_hi, const void* x_lo, int64_t B, int H, int W, int Cp, int C, float* y,
                 cudaStream_t stream);
int split_f16(const float* in, int64_t count, void* hi, void* lo, cudaStream_t stream);
int relu_mask_nhwc(float* R, const void* a_hi, const void* a_lo, int64_t count, cudaStream_t stream);
int pack_f16(const float* in, int64_t count, float scale, void* out, cudaStream_t stream);
int pack_f16_hilo(const float* in, int64_t count, float scale, void* out_hi, void* out_lo, cudaStream_t stream);
int absmax(const float* in, int64_t count, float* out, cudaStream_t stream);
int64_t finish_workspace_bytes(int d, int m);
int finish_step(const float* sums, int64_t M_global, const float* U, int d, int m, int K, float* U_out, void* Ut_hi,
                void* Ut_lo, float* obj_log, int64_t log_index, int max_iters, float tol, int u_rounded, int* status,
                void* workspace, int64_t workspace_bytes, const drsa_peer_exchange* px, cudaStream_t stream);
bool finish_fused_supported(int d, int m, int K);
int64_t exchange_bytes(int d, int m, int K, int world);
int polar_retract(const float* Y, int d, int m, float* U_out, int max_iters, float tol, int* status, void* workspace,
                  int64_t workspace_bytes, cudaStream_t stream);
int split_u(const float* U, int d, int m, void* Ut_hi, void* Ut_lo, cudaStream_t stream);
int finish_step_qr(const float* sums, int64_t M_global, const float* U, int d, int m, int K, float* U_out, float* obj_log,
                   int64_t log_index, int* status, void* workspace, int64_t workspace_bytes, cudaStream_t stream);
int qr_retract(const float* Y, int d, int m, float* U_out, int* status, void* workspace, int64_t workspace_bytes,
               cudaStream_t stream);
int selftest_umma(int variant, float* max_err_host);
void set_tc_profile(long long* p);
int tc_kernel_attrs(int d, int split, int* out5);
void set_tc_variant(int v);
void set_conv_variant(int v);
int64_t subspace_relevances_workspace_bytes(int64_t B, int64_t P, int d, int m);
int subset_objectives(const float* act, const float* ctx, const float* U, int64_t S, int64_t R, int d, int m, int K,
                      float* obj, float* sumsq, void* workspace, int64_t workspace_bytes, cudaStream_t stream);
int subspace_relevances(const float* act, const float* ctx, const float* U, int64_t B, int64_t P, int d, int m, int K,
                        float* out, void* workspace, int64_t workspace_bytes, cudaStream_t stream);
int context_gather(const float* a_map, const float* R_map, int64_t N, int d, int HW, const int64_t* idx, int L,
                   float* act_out, float* ctx_out, double* sumsq, cudaStream_t stream);
int context_vectors(const float* a, const float* R, int64_t count, float* out, cudaStream_t stream);
int sumsq(const float* v, int64_t count, double* out, cudaStream_t stream);
int normalize(float* v, int64_t rows, int d, const double* sumsq, int64_t count_global, cudaStream_t stream);
int context_pairs_nhwc(const void* hi, const void* lo, const float* R, int64_t N, int HW, int Cp, int d, const int64_t* idx,
                       int L, float* act_out, float* ctx_out, double* sumsq, cudaStream_t stream);
int sums_combine(const float* a, const float* b, float beta, float* out, int64_t n, cudaStream_t stream);

static bool shape_ok(int64_t M, int d, int m, int K) {
  return M > 0 && d > 0 && m > 0 && K > 0 && m <= d && m % K == 0;
}
}  // namespace drsa

using namespace drsa;

extern "C" {

const char* drsa_status_string(int status) {
  switch (status) {
    case DRSA_OK: return "ok";
    case DRSA_ERR_ARG: return "invalid argument";
    case DRSA_ERR_SHAPE: return "shape not supported by this precision mode";
    case DRSA_ERR_ARCH: return "device is not sm_100 (no fallback path exists)";
    case DRSA_ERR_ALIGN: return "pointer or leading dimension not 16-byte aligned";
    case DRSA_ERR_WORKSPACE: return "workspace too small";
    case DRSA_ERR_CUDA: return "CUDA call failed (see drsa_last_cuda_error)";
    case DRSA_ERR_RANGE: return "value out of range for this precision mode";
    default: return "unknown status";
  }
}

int drsa_version(void) { return DRSA_B200_VERSION; }
int drsa_last_cuda_error(void) { return g_last_cuda_error; }

int drsa_check_device(int device) {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { cudaGetLastError(); return DRSA_ERR_ARCH; }
  return prop.major == 10 ? DRSA_OK : DRSA_ERR_ARCH;
}

int drsa_pack_f16(const float* in, int64_t count, float scale, void* out_f16, void* stream) {
  if (in == nullptr || out_f16 == nullptr || count <= 0 || !(scale > 0.f)) return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return pack_f16(in, count, scale, out_f16, static_cast<cudaStream_t>(stream));
}

int drsa_pack_f16_hilo(const float* in, int64_t count, float scale, void* out_hi_f16, void* out_lo_f16, void* stream) {
  if (in == nullptr || out_hi_f16 == nullptr || out_lo_f16 == nullptr || count <= 0 || !(scale > 0.f)) return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return pack_f16_hilo(in, count, scale, out_hi_f16, out_lo_f16, static_cast<cudaStream_t>(stream));
}

int drsa_absmax(const float* in, int64_t count, float* out, void* stream) {
  if (in == nullptr || out == nullptr || count <= 0) return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return absmax(in, count, out, static_cast<cudaStream_t>(stream));
}

int drsa_rownorm_max(const float* in, int64_t rows, int d, float* out, void* stream) {
  if (in == nullptr || out == nullptr || rows <= 0 || d <= 0) return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return rownorm_max(in, rows, d, out, static_cast<cudaStream_t>(stream));
}

int64_t drsa_step_workspace_bytes(int64_t M, int d, int m, int K, int precision) {
  if (!shape_ok(M, d, m, K)) return DRSA_ERR_ARG;
  if (precision == DRSA_PREC_FP32) return step_fp32_workspace_bytes(M, d, m, K);
  if (precision == DRSA_PREC_TC_F16X2 || precision == DRSA_PREC_TC_F16 || precision == DRSA_PREC_TC_F16_AC2 ||
      precision == DRSA_PREC_TC_F32C) {
    const bool split_u = precision == DRSA_PREC_TC_F16X2 || precision == DRSA_PREC_TC_F32C;
    if (!tc_shape_supported(d, m, K) || (d == 512 && split_u)) return DRSA_ERR_SHAPE;
    return step_tc_workspace_bytes(M, d, m, K);
  }
  return DRSA_ERR_ARG;
}

int drsa_step(const void* A, const void* C, const float* U, const void* Ut_hi, const void* Ut_lo, int64_t M, int d,
              int m, int K, int precision, float scaleA, float scaleC, float pq_scale, float* sums, void* workspace,
              int64_t workspace_bytes, void* stream) {
  if (A == nullptr || C == nullptr || sums == nullptr || workspace == nullptr || !shape_ok(M, d, m, K))
    return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (precision == DRSA_PREC_FP32) {
    if (U == nullptr) return DRSA_ERR_ARG;
    return step_fp32(static_cast<const float*>(A), static_cast<const float*>(C), U, M, d, m, K, sums, workspace,
                     workspace_bytes, s);
  }
  if (precision == DRSA_PREC_TC_F16X2 || precision == DRSA_PREC_TC_F16 || precision == DRSA_PREC_TC_F16_AC2 ||
      precision == DRSA_PREC_TC_F32C) {
    const bool split = precision == DRSA_PREC_TC_F16X2 || precision == DRSA_PREC_TC_F32C;
    const bool split_ac = precision == DRSA_PREC_TC_F16_AC2 || precision == DRSA_PREC_TC_F32C;
    if (Ut_hi == nullptr || (split && Ut_lo == nullptr) || !(scaleA > 0.f) || !(scaleC > 0.f) || !(pq_scale > 0.f))
      return DRSA_ERR_ARG;
    return step_tc(A, C, Ut_hi, Ut_lo, M, d, m, K, scaleA, scaleC, pq_scale, split, split_ac, sums, workspace,
                   workspace_bytes, s);
  }
  return DRSA_ERR_ARG;
}

int drsa_finish_step_qr(const float* sums, int64_t M_global, const float* U, int d, int m, int K, float* U_out,
                        float* obj_log, int64_t log_index, int* status, void* workspace, int64_t workspace_bytes,
                        void* stream) {
  if (sums == nullptr || U == nullptr || status == nullptr || workspace == nullptr || M_global <= 0 || d <= 0 || m <= 0 ||
      m > d || K <= 0 || m % K != 0)
    return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return finish_step_qr(sums, M_global, U, d, m, K, U_out, obj_log, log_index, status, workspace, workspace_bytes,
                        static_cast<cudaStream_t>(stream));
}

int drsa_qr_retract(const float* Y, int d, int m, float* U_out, int* status, void* workspace, int64_t workspace_bytes,
                    void* stream) {
  if (Y == nullptr || U_out == nullptr || workspace == nullptr || d <= 0 || m <= 0 || m > d) return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return qr_retract(Y, d, m, U_out, status, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int drsa_sums_combine(const float* a, const float* b, float beta, float* out, int64_t n, void* stream) {
  if (a == nullptr || b == nullptr || out == nullptr || n <= 0) return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return sums_combine(a, b, beta, out, n, static_cast<cudaStream_t>(stream));
}

int drsa_split_u(const float* U, int d, int m, void* Ut_hi, void* Ut_lo, void* stream) {
  if (U == nullptr || Ut_hi == nullptr || d <= 0 || m <= 0) return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return split_u(U, d, m, Ut_hi, Ut_lo, static_cast<cudaStream_t>(stream));
}

int64_t drsa_finish_workspace_bytes(int d, int m) {
  if (d <= 0 || m <= 0 || m > d) return DRSA_ERR_ARG;
  return finish_workspace_bytes(d, m);
}

int drsa_finish_step(const float* sums, int64_t M_global, const float* U, int d, int m, int K, float* U_out,
                     void* Ut_hi, void* Ut_lo, float* obj_log, int64_t log_index, int max_iters, float tol,
                     int u_rounded, int* status, void* workspace, int64_t workspace_bytes, void* stream) {
  if (sums == nullptr || !shape_ok(M_global, d, m, K) || workspace == nullptr) return DRSA_ERR_ARG;
  if (U_out != nullptr && (U == nullptr || max_iters < 1 || !(tol > 0.f))) return DRSA_ERR_ARG;
  if (Ut_lo != nullptr && Ut_hi == nullptr) return DRSA_ERR_ARG;
  if (u_rounded && U == nullptr) return DRSA_ERR_ARG;
  if (K > 1024) return DRSA_ERR_SHAPE;
  if (obj_log != nullptr && log_index < 0 && status == nullptr) return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return finish_step(sums, M_global, U, d, m, K, U_out, Ut_hi, Ut_lo, obj_log, log_index, max_iters, tol, u_rounded,
                     status, workspace, workspace_bytes, nullptr, static_cast<cudaStream_t>(stream));
}

int64_t drsa_exchange_bytes(int d, int m, int K, int world) {
  if (d <= 0 || m <= 0 || m > d || K <= 0 || m % K != 0 || world < 2) return DRSA_ERR_ARG;
  if (world > DRSA_MAX_PEERS || !finish_fused_supported(d, m, K)) return DRSA_ERR_SHAPE;
  return exchange_bytes(d, m, K, world);
}

int drsa_finish_step_p2p(const drsa_peer_exchange* px, const float* sums, int64_t M_global, const float* U, int d,
                         int m, int K, float* U_out, void* Ut_hi, void* Ut_lo, float* obj_log, int64_t log_index,
                         int max_iters, float tol, int u_rounded, int* status, void* workspace,
                         int64_t workspace_bytes, void* stream) {
  if (px == nullptr || px->world < 2 || px->world > DRSA_MAX_PEERS || px->rank < 0 || px->rank >= px->world)
    return DRSA_ERR_ARG;
  if (sums == nullptr || !shape_ok(M_global, d, m, K) || workspace == nullptr) return DRSA_ERR_ARG;
  if (U_out != nullptr && (U == nullptr || max_iters < 1 || !(tol > 0.f))) return DRSA_ERR_ARG;
  if (Ut_lo != nullptr && Ut_hi == nullptr) return DRSA_ERR_ARG;
  if (u_rounded && U == nullptr) return DRSA_ERR_ARG;
  if (K > 1024 || !finish_fused_supported(d, m, K)) return DRSA_ERR_SHAPE;
  if (obj_log != nullptr && log_index < 0 && status == nullptr) return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return finish_step(sums, M_global, U, d, m, K, U_out, Ut_hi, Ut_lo, obj_log, log_index, max_iters, tol, u_rounded,
                     status, workspace, workspace_bytes, px, static_cast<cudaStream_t>(stream));
}

/* cudaIpc plumbing for the exchange buffers (used when torch's symmetric memory is not available) */
int drsa_ipc_alloc(int64_t bytes, void** ptr, unsigned char* handle64) {
  if (bytes <= 0 || ptr == nullptr || handle64 == nullptr) return DRSA_ERR_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  void* p = nullptr;
  DRSA_CUDA(cudaMalloc(&p, (size_t)bytes));
  cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e); }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return DRSA_OK;
}

int drsa_ipc_open(const unsigned char* handle64, void** ptr) {
  if (handle64 == nullptr || ptr == nullptr) return DRSA_ERR_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  DRSA_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return DRSA_OK;
}

int drsa_ipc_close(void* ptr) {
  if (ptr == nullptr) return DRSA_ERR_ARG;
  DRSA_CUDA(cudaIpcCloseMemHandle(ptr));
  return DRSA_OK;
}

int drsa_ipc_free(void* ptr) {
  if (ptr == nullptr) return DRSA_ERR_ARG;
  DRSA_CUDA(cudaFree(ptr));
  return DRSA_OK;
}

int drsa_polar_retract(const float* Y, int d, int m, float* U_out, int max_iters, float tol, int* status,
                       void* workspace, int64_t workspace_bytes, void* stream) {
  if (Y == nullptr || U_out == nullptr || workspace == nullptr || d <= 0 || m <= 0 || m > d || max_iters < 1 ||
      !(tol > 0.f))
    return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return polar_retract(Y, d, m, U_out, max_iters, tol, status, workspace, workspace_bytes,
                       static_cast<cudaStream_t>(stream));
}

int64_t drsa_subspace_relevances_workspace_bytes(int64_t B, int64_t P, int d, int m) {
  if (B <= 0 || P <= 0 || d <= 0 || m <= 0) return DRSA_ERR_ARG;
  return subspace_relevances_workspace_bytes(B, P, d, m);
}

int drsa_subspace_relevances(const float* act, const float* ctx, const float* U, int64_t B, int64_t P, int d, int m,
                             int K, float* out, void* workspace, int64_t workspace_bytes, void* stream) {
  if (act == nullptr || ctx == nullptr || U == nullptr || out == nullptr || workspace == nullptr ||
      !shape_ok(B * P, d, m, K) || B <= 0 || P <= 0)
    return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return subspace_relevances(act, ctx, U, B, P, d, m, K, out, workspace, workspace_bytes,
                             static_cast<cudaStream_t>(stream));
}

int drsa_context_pairs_nhwc(const void* a_hi, const void* a_lo, const float* R, int64_t N, int HW, int Cp, int d,
                            const int64_t* idx, int L, float* act_out, float* ctx_out, double* sumsq, void* stream) {
  if (a_hi == nullptr || a_lo == nullptr || R == nullptr || act_out == nullptr || ctx_out == nullptr || N <= 0 || HW <= 0 ||
      d <= 0 || Cp < d || L <= 0 || (idx == nullptr && L != HW))
    return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return context_pairs_nhwc(a_hi, a_lo, R, N, HW, Cp, d, idx, L, act_out, ctx_out, sumsq, static_cast<cudaStream_t>(stream));
}

int drsa_subset_objectives(const float* act, const float* ctx, const float* U, int64_t S, int64_t R, int d, int m, int K,
                           float* obj, float* sumsq, void* workspace, int64_t workspace_bytes, void* stream) {
  if (act == nullptr || ctx == nullptr || U == nullptr || obj == nullptr || sumsq == nullptr || workspace == nullptr ||
      S <= 0 || R <= 0 || !shape_ok(S * R, d, m, K))
    return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return subset_objectives(act, ctx, U, S, R, d, m, K, obj, sumsq, workspace, workspace_bytes,
                           static_cast<cudaStream_t>(stream));
}

int drsa_context_gather(const float* a_map, const float* R_map, int64_t N, int d, int HW, const int64_t* idx, int L,
                        float* act_out, float* ctx_out, double* sumsq, void* stream) {
  if (a_map == nullptr || R_map == nullptr || act_out == nullptr || ctx_out == nullptr || N <= 0 || d <= 0 ||
      HW <= 0 || L <= 0 || (idx == nullptr && L != HW))
    return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return context_gather(a_map, R_map, N, d, HW, idx, L, act_out, ctx_out, sumsq, static_cast<cudaStream_t>(stream));
}

int drsa_context_vectors(const float* a, const float* R, int64_t count, float* out, void* stream) {
  if (a == nullptr || R == nullptr || out == nullptr || count <= 0) return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return context_vectors(a, R, count, out, static_cast<cudaStream_t>(stream));
}

int drsa_sumsq(const float* v, int64_t count, double* out, void* stream) {
  if (v == nullptr || out == nullptr || count <= 0) return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return sumsq(v, count, out, static_cast<cudaStream_t>(stream));
}

int drsa_normalize(float* v, int64_t rows, int d, const double* sumsq, int64_t count_global, void* stream) {
  if (v == nullptr || sumsq == nullptr || rows <= 0 || d <= 0 || count_global <= 0) return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return normalize(v, rows, d, sumsq, count_global, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------- stage 1 (LRP)
static bool conv_ok(int64_t N, int Cin, int Cout, int H, int W) { return N > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0; }

int lrp_conv3x3_forward(const float* x, const float* w, const float* b, int64_t N, int Cin, int Cout, int H, int W,
                        int relu, float* y, void* stream) {
  if (x == nullptr || w == nullptr || y == nullptr || !conv_ok(N, Cin, Cout, H, W)) return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return conv3x3_forward(x, w, b, N, Cin, Cout, H, W, relu, y, static_cast<cudaStream_t>(stream));
}

int lrp_conv3x3_flip_weights(const float* w, int Cout, int Cin, float* wt, void* stream) {
  if (w == nullptr || wt == nullptr || Cout <= 0 || Cin <= 0) return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return flip_weights(w, Cout, Cin, wt, static_cast<cudaStream_t>(stream));
}

int lrp_conv3x3_backward(const float* x, const float* w_mod, const float* wt_mod, const float* b_mod, const float* R_out,
                         int64_t N, int Cin, int Cout, int H, int W, float eps, int x_is_ones, float* s_buf,
                         float* R_in, void* stream) {
  if ((x == nullptr && !x_is_ones) || w_mod == nullptr || wt_mod == nullptr || R_out == nullptr || s_buf == nullptr ||
      R_in == nullptr || !conv_ok(N, Cin, Cout, H, W))
    return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return conv3x3_backward(x, w_mod, wt_mod, b_mod, R_out, N, Cin, Cout, H, W, eps, x_is_ones, s_buf, R_in,
                          static_cast<cudaStream_t>(stream));
}

int lrp_maxpool_forward(const float* x, int64_t NC, int H, int W, int kh, int kw, float* y, int32_t* argmax,
                        void* stream) {
  if (x == nullptr || y == nullptr || argmax == nullptr || NC <= 0 || H <= 0 || W <= 0 || kh <= 0 || kw <= 0 ||
      H / kh == 0 || W / kw == 0)
    return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return maxpool_forward(x, NC, H, W, kh, kw, y, argmax, static_cast<cudaStream_t>(stream));
}

int lrp_maxpool_backward(const float* R_out, const int32_t* argmax, int64_t NC, int H, int W, int kh, int kw,
                         float* R_in, void* stream) {
  if (R_out == nullptr || argmax == nullptr || R_in == nullptr || NC <= 0 || H <= 0 || W <= 0 || kh <= 0 || kw <= 0)
    return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return maxpool_backward(R_out, argmax, NC, H, W, kh, kw, R_in, static_cast<cudaStream_t>(stream));
}

int lrp_dense_forward(const float* x, const float* w, const float* b, int64_t N, int In, int Out, int relu, float* y,
                      void* stream) {
  if (x == nullptr || w == nullptr || y == nullptr || N <= 0 || In <= 0 || Out <= 0 || N > 2147483647LL)
    return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return dense_forward(x, w, b, N, In, Out, relu, y, static_cast<cudaStream_t>(stream));
}

int lrp_dense_epsilon_backward(const float* x, const float* w, const float* b, const float* R_out, int64_t N, int In,
                               int Out, float eps, float* s_buf, float* R_in, void* stream) {
  if (x == nullptr || w == nullptr || R_out == nullptr || s_buf == nullptr || R_in == nullptr || N <= 0 || In <= 0 ||
      Out <= 0 || N > 2147483647LL)
    return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return dense_epsilon_backward(x, w, b, R_out, N, In, Out, eps, s_buf, R_in, static_cast<cudaStream_t>(stream));
}

int lrp_relu_mask(const float* a, float* R, int64_t count, void* stream) {
  if (a == nullptr || R == nullptr || count <= 0) return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return relu_mask(a, R, count, static_cast<cudaStream_t>(stream));
}

int lrp_tc_conv3x3_supported(int64_t B, int Cin_p, int Cout_p, int H, int W) {
  return conv_tc_supported(B, Cin_p, Cout_p, H, W) ? DRSA_OK : DRSA_ERR_SHAPE;
}

int lrp_tc_conv3x3_forward(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias,
                           int64_t B, int H, int W, int Cin_p, int Cout_p, int Cout, int relu, void* y_hi, void* y_lo,
                           float* y_nchw, int* err_flag, void* stream) {
  if (x_hi == nullptr || x_lo == nullptr || w_hi == nullptr || w_lo == nullptr || bias == nullptr || err_flag == nullptr ||
      (y_hi == nullptr) != (y_lo == nullptr) || (y_hi == nullptr && y_nchw == nullptr) || Cout <= 0 || Cout > Cout_p)
    return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return conv_tc_forward(x_hi, x_lo, w_hi, w_lo, bias, B, H, W, Cin_p, Cout_p, Cout, relu, y_hi, y_lo, y_nchw, err_flag,
                         static_cast<cudaStream_t>(stream));
}

int lrp_tc_first_ones_backward(const float* R_out, const float* w_mod, const float* b_mod, int64_t B, int H, int W, int Cout,
                               int Cp, float eps, float* R_in, void* stream) {
  if (R_out == nullptr || w_mod == nullptr || b_mod == nullptr || R_in == nullptr || B <= 0 || H <= 0 || W <= 0 || Cout <= 0)
    return DRSA_ERR_ARG;
  if (!aligned16(R_out)) return DRSA_ERR_ALIGN;
  DRSA_TRY(require_sm100());
  return first_layer_ones_backward(R_out, w_mod, b_mod, B, H, W, Cout, Cp, eps, R_in, static_cast<cudaStream_t>(stream));
}

int lrp_tc_conv3x3_pool_supported(int64_t B, int Cin_p, int Cout_p, int H, int W, int kh, int kw) {
  return conv_tc_pool_supported(B, Cin_p, Cout_p, H, W, kh, kw) ? DRSA_OK : DRSA_ERR_SHAPE;
}

int lrp_tc_conv3x3_forward_pool(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias,
                                int64_t B, int H, int W, int Cin_p, int Cout_p, int Cout, int relu, int kh, int kw,
                                void* y_hi, void* y_lo, void* argmax_u8, int* err_flag, void* stream) {
  if (x_hi == nullptr || x_lo == nullptr || w_hi == nullptr || w_lo == nullptr || bias == nullptr || err_flag == nullptr ||
      y_hi == nullptr || y_lo == nullptr || Cout <= 0 || Cout > Cout_p || kh <= 0 || kw <= 0)
    return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return conv_tc_forward_pool(x_hi, x_lo, w_hi, w_lo, bias, B, H, W, Cin_p, Cout_p, Cout, relu, kh, kw, y_hi, y_lo, argmax_u8,
                              err_flag, static_cast<cudaStream_t>(stream));
}

int lrp_tc_conv3x3_first(const float* x, const float* w, const float* b, int64_t B, int H, int W, int Cout, int Cout_p,
                         int relu, void* y_hi, void* y_lo, void* stream) {
  if (x == nullptr || w == nullptr || b == nullptr || y_hi == nullptr || y_lo == nullptr || B <= 0 || H <= 0 || W <= 0 ||
      Cout <= 0 || Cout_p % 32 != 0 || Cout > Cout_p || Cout_p > 1024)
    return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return conv_first_nhwc(x, w, b, B, H, W, Cout, Cout_p, relu, y_hi, y_lo, static_cast<cudaStream_t>(stream));
}

int lrp_tc_maxpool(const void* x_hi, const void* x_lo, int64_t B, int H, int W, int Cp, int kh, int kw, void* y_hi,
                   void* y_lo, void* argmax_u8, void* stream) {
  if (x_hi == nullptr || x_lo == nullptr || y_hi == nullptr || y_lo == nullptr || B <= 0 || Cp % 8 != 0 || kh <= 0 ||
      kw <= 0 || H / kh == 0 || W / kw == 0 || kh * kw > 255)
    return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return maxpool_nhwc(x_hi, x_lo, B, H, W, Cp, kh, kw, y_hi, y_lo, argmax_u8, static_cast<cudaStream_t>(stream));
}

int lrp_tc_maxpool_backward(const float* R_out, const void* argmax_u8, int64_t B, int H, int W, int Cp, int kh, int kw,
                            float* R_in, void* stream) {
  if (R_out == nullptr || argmax_u8 == nullptr || R_in == nullptr || B <= 0 || Cp % 4 != 0 || kh <= 0 || kw <= 0 ||
      H % kh != 0 || W % kw != 0)
    return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return maxpool_nhwc_backward(R_out, argmax_u8, B, H, W, Cp, kh, kw, R_in, static_cast<cudaStream_t>(stream));
}

int lrp_tc_conv3x3_ratio(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias,
                         const float* R_out, int64_t B, int H, int W, int Cin_p, int Cout_p, float eps,
                         const float* scale_ref, void* s_hi, void* s_lo, int* err_flag, void* stream) {
  if (x_hi == nullptr || x_lo == nullptr || w_hi == nullptr || w_lo == nullptr || bias == nullptr || R_out == nullptr ||
      s_hi == nullptr || s_lo == nullptr || err_flag == nullptr)
    return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return conv_tc_run(x_hi, x_lo, w_hi, w_lo, bias, B, H, W, Cin_p, Cout_p, Cout_p, 0, 1, eps, R_out, nullptr, nullptr, s_hi,
                     s_lo, nullptr, nullptr, scale_ref, nullptr, err_flag, static_cast<cudaStream_t>(stream), 0, 0, nullptr);
}

int lrp_tc_conv3x3_inputmul(const void* s_hi, const void* s_lo, const void* wt_hi, const void* wt_lo, const void* x_hi,
                            const void* x_lo, int64_t B, int H, int W, int Cout_p, int Cin_p, const float* scale_ref,
                            float* cmax_out, float* R_in, int* err_flag, void* stream) {
  if (s_hi == nullptr || s_lo == nullptr || wt_hi == nullptr || wt_lo == nullptr || x_hi == nullptr || x_lo == nullptr ||
      R_in == nullptr || err_flag == nullptr)
    return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return conv_tc_run(s_hi, s_lo, wt_hi, wt_lo, nullptr, B, H, W, Cout_p, Cin_p, Cin_p, 0, 2, 0.f, nullptr, x_hi, x_lo, nullptr,
                     nullptr, R_in, nullptr, scale_ref, cmax_out, err_flag, static_cast<cudaStream_t>(stream), 0, 0, nullptr);
}

int lrp_tc_sample_absmax_ratio(const float* R, const float* x, int64_t B, int64_t per_sample, float* out, void* stream) {
  if (R == nullptr || x == nullptr || out == nullptr || B <= 0 || per_sample <= 0) return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return sample_absmax_ratio(R, x, B, per_sample, out, static_cast<cudaStream_t>(stream));
}

int64_t lrp_subspace_filter_workspace_bytes(int64_t P, int d, int m) {
  if (P <= 0 || d <= 0 || m <= 0) return DRSA_ERR_ARG;
  return subspace_filter_workspace_bytes(P, d, m);
}

int lrp_subspace_project(const float* a, const float* U, int64_t P, int d, int m, int ld, float* h, float* a_rec,
                         void* stream) {
  if (a == nullptr || U == nullptr || h == nullptr || P <= 0 || d <= 0 || m <= 0 || ld < d) return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return subspace_project(a, U, P, d, m, ld, h, a_rec, static_cast<cudaStream_t>(stream));
}

int lrp_subspace_filter(const float* a, const float* h, const float* a_rec, const float* R, const float* U, int64_t P,
                        int d, int m, int K, int ld, float eps_invprojection, float eps_projection, float* out,
                        void* workspace, int64_t workspace_bytes, void* stream) {
  if (a == nullptr || h == nullptr || a_rec == nullptr || R == nullptr || U == nullptr || out == nullptr ||
      workspace == nullptr || P <= 0 || d <= 0 || m <= 0 || K <= 0 || m % K != 0 || ld < d)
    return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return subspace_filter_backward(a, h, a_rec, R, U, P, d, m, K, ld, eps_invprojection, eps_projection, out, workspace,
                                  workspace_bytes, static_cast<cudaStream_t>(stream));
}

int lrp_tc_planes_to_f32(const void* x_hi, const void* x_lo, int64_t count, float* out, void* stream) {
  if (x_hi == nullptr || x_lo == nullptr || out == nullptr || count <= 0) return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return planes_to_f32(x_hi, x_lo, count, out, static_cast<cudaStream_t>(stream));
}

int lrp_tc_relu_mask(float* R, const void* a_hi, const void* a_lo, int64_t count, void* stream) {
  if (R == nullptr || a_hi == nullptr || a_lo == nullptr || count <= 0) return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return relu_mask_nhwc(R, a_hi, a_lo, count, static_cast<cudaStream_t>(stream));
}

int lrp_tc_nhwc_f32_to_nchw(const float* x, int64_t B, int H, int W, int Cp, int C, float* y, void* stream) {
  if (x == nullptr || y == nullptr || B <= 0 || C <= 0 || C > Cp) return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return nhwc_f32_to_nchw(x, B, H, W, Cp, C, y, static_cast<cudaStream_t>(stream));
}

int lrp_tc_nchw_to_nhwc_f32(const float* x, int64_t B, int H, int W, int C, int Cp, float* y, void* stream) {
  if (x == nullptr || y == nullptr || B <= 0 || C <= 0 || C > Cp) return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return nchw_to_nhwc_f32(x, B, H, W, C, Cp, y, static_cast<cudaStream_t>(stream));
}

int lrp_tc_nhwc_to_nchw(const void* x_hi, const void* x_lo, int64_t B, int H, int W, int Cp, int C, float* y,
                        void* stream) {
  if (x_hi == nullptr || x_lo == nullptr || y == nullptr || B <= 0 || C <= 0 || C > Cp) return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return nhwc_to_nchw(x_hi, x_lo, B, H, W, Cp, C, y, static_cast<cudaStream_t>(stream));
}

int lrp_tc_split_f16(const float* in, int64_t count, void* hi, void* lo, void* stream) {
  if (in == nullptr || hi == nullptr || lo == nullptr || count <= 0) return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return split_f16(in, count, hi, lo, static_cast<cudaStream_t>(stream));
}

int64_t logmel_transform_workspace_bytes(int64_t B, int n_fft, int n_mels, int width) {
  if (B <= 0 || n_fft < 2 || n_mels <= 0 || width <= 0) return DRSA_ERR_ARG;
  return logmel_workspace_bytes(B, n_fft, n_mels, width);
}

int logmel_transform_wav(const float* wav, const float* window, const float* dft_basis, const float* mel_fb, int64_t B,
                         int64_t n_samples, int n_fft, int hop_length, int n_mels, int first_frame, int width, int clamp,
                         float clamp_min, float* out, void* workspace, int64_t workspace_bytes, void* stream) {
  if (wav == nullptr || window == nullptr || dft_basis == nullptr || mel_fb == nullptr || out == nullptr ||
      workspace == nullptr || B <= 0 || n_samples <= 0 || n_fft < 2 || hop_length <= 0 || n_mels <= 0 || first_frame < 0 ||
      width <= 0)
    return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  return logmel_transform(wav, window, dft_basis, mel_fb, B, n_samples, n_fft, hop_length, n_mels, first_frame, width, clamp,
                          clamp_min, out, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int drsa_debug_set_tc_variant(int variant) { set_tc_variant(variant); return DRSA_OK; }
int lrp_debug_set_conv_variant(int variant) { set_conv_variant(variant); return DRSA_OK; }

int drsa_debug_tc_kernel_attrs(int d, int split, int* out5) {
  if (out5 == nullptr || (d != 128 && d != 256 && d != 512)) return DRSA_ERR_ARG;
  return tc_kernel_attrs(d, split, out5);
}

int drsa_debug_set_tc_profile(void* device_buf6) { set_tc_profile(static_cast<long long*>(device_buf6)); return DRSA_OK; }

int drsa_selftest_umma(int variant, float* max_err_host) { return selftest_umma(variant, max_err_host); }

}  // extern "C"
