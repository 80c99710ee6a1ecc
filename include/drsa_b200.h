/*
 * drsa_b200.h -- C ABI of libdrsa_b200.so, the sm_100a implementation of the
 * DRSA / LRP explanation hot path of sharckhai/drsa-audio.
 *
 * The reference is pure Python and has no FFI of its own; every entry point below
 * sits UNDER one of the reference's Python functions and names it (file:line are
 * relative to the reference repository).  The Python package `cxai` in this
 * repository keeps the reference's names and signatures and binds these symbols
 * through ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain C, no torch types: device pointers, sizes, a CUDA stream as void*.
 *   - every function returns DRSA_OK (0) or a negative drsa_status; nothing throws.
 *   - all work is asynchronous on `stream`; no function synchronises the host, so a
 *     whole optimisation loop can be enqueued or captured in a CUDA graph.
 *   - the library never allocates or frees device memory.  Callers pass workspaces
 *     sized by the *_workspace_bytes() queries and keep every buffer alive until the
 *     stream work has completed.
 *   - there is no CPU fallback: on a device that is not sm_100 the compute entry
 *     points return DRSA_ERR_ARCH.
 *   - matrices are row-major fp32 unless stated otherwise.
 */
#ifndef DRSA_B200_H_
#define DRSA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRSA_B200_VERSION 100

typedef enum drsa_status {
  DRSA_OK = 0,
  DRSA_ERR_ARG = -1,        /* null pointer, non-positive size, bad enum            */
  DRSA_ERR_SHAPE = -2,      /* shape not supported by the requested precision mode   */
  DRSA_ERR_ARCH = -3,       /* device is not sm_100                                  */
  DRSA_ERR_ALIGN = -4,      /* pointer / leading dimension not 16-byte aligned       */
  DRSA_ERR_WORKSPACE = -5,  /* workspace too small                                   */
  DRSA_ERR_CUDA = -6,       /* a CUDA runtime / driver call failed (see last_cuda)   */
  DRSA_ERR_RANGE = -7       /* value outside the representable range of the mode     */
} drsa_status;

/* Arithmetic of the projection / gradient contractions (drsa.py:148-149 and the
 * autograd backward of them, drsa.py:100). */
typedef enum drsa_precision {
  DRSA_PREC_FP32 = 0,       /* CUDA-core FFMA, fp32 throughout                        */
  DRSA_PREC_TC_F16X2 = 1,   /* tcgen05 kind::f16, fp32 accumulate in TMEM; A and C are
                               stored once as scaled fp16, U is split hi+lo every step */
  DRSA_PREC_TC_F16 = 2,     /* as above with U rounded to fp16 once per step (one MMA per
                               product): the sums are exact for fp16(U); drsa_finish_step
                               (u_rounded = 1) adds the first-order term <grad, U - fp16(U)>
                               to the logged objective, so the objective stays second-order
                               accurate in the rounding                  */
  DRSA_PREC_TC_F16_AC2 = 3, /* as DRSA_PREC_TC_F16 (U rounded once per step, u_rounded = 1)
                               with the rows stored as TWO fp16 planes A = A_hi + A_lo (22
                               bits, drsa_pack_f16_hilo): two MMAs per product with the rows */
  DRSA_PREC_TC_F32C = 4     /* fp32-class operands on the tensor cores: rows hi + lo AND
                               U hi + lo (3 MMAs per forward product, 2 per gradient product,
                               the O(2^-22) lo x lo term dropped); d <= 256.  The mode that
                               follows the reference's fp32 trajectory over its full
                               2 000-step horizon for any row count        */
} drsa_precision;

const char* drsa_status_string(int status);
int drsa_version(void);
/* last cudaError_t seen by this thread inside the library (0 if none). */
int drsa_last_cuda_error(void);
/* DRSA_OK if `device` is an sm_100 part. */
int drsa_check_device(int device);

/* ------------------------------------------------------------------------------------
 * Stage 2: DRSA subspace optimisation
 *   reference: cxai/xai/drsa/drsa.py  SubspaceOptimizer.obj_val :123-155,
 *              objective_fn :224-238, generalized_fmean :171-182, run :76-120,
 *              orthogonalize :201-221
 * ---------------------------------------------------------------------------------- */

/* Data preparation for DRSA_PREC_TC_F16X2: out[i] = fp16(in[i] * scale) over M*d
 * elements (scale is a power of two chosen by the caller so that max|in|*scale stays
 * far below 65504).  Done once per optimiser because A and C are constant over the
 * steps of drsa.py:84.  Returns DRSA_ERR_ALIGN unless both pointers are 16-byte
 * aligned. */
int drsa_pack_f16(const float* in, int64_t count, float scale, void* out_f16, void* stream);

/* Data preparation for DRSA_PREC_TC_F16_AC2 / DRSA_PREC_TC_F32C: out_hi[i] = fp16(in[i]*scale),
 * out_lo[i] = fp16(in[i]*scale - out_hi[i]).  drsa_step expects the two planes of a matrix
 * back to back: out_lo = out_hi + M*d elements. */
int drsa_pack_f16_hilo(const float* in, int64_t count, float scale, void* out_hi_f16, void* out_lo_f16, void* stream);

/* max |x| over `count` floats -> *out (one float).  Used to choose the pack scale. */
int drsa_absmax(const float* in, int64_t count, float* out, void* stream);

/* max over rows of the Euclidean row norm of in [rows, d] -> *out.  With it the caller bounds
 * |g * HC| <= pq_scale * rhoA * rhoC^2 and picks pq_scale so the fp16 P/Q operands of the
 * gradient GEMM can never overflow. */
int drsa_rownorm_max(const float* in, int64_t rows, int d, float* out, void* stream);

/* Bytes of workspace drsa_step needs for the given problem. */
int64_t drsa_step_workspace_bytes(int64_t M, int d, int m, int K, int precision);

/*
 * One pass over the (activation, context) rows: the forward projection, the
 * per-concept ReLU'd relevance and BOTH row-sums the update needs
 *      sums[0 .. d*m)     X[i][j], j in concept k :  sum_r A[r][i] g_rk HC[r][j] + C[r][i] g_rk HA[r][j],
 *                         g_rk = relu(s_rk), s_rk = sum_{j in k} HA[r][j] HC[r][j]      (un-normalised gradient)
 *      sums[d*m .. d*m+K) sum_r relu(s_rk)^2
 * i.e. drsa.py:148-155 fused with the backward of drsa.py:100.  The
 * sums are additive over row shards, so under data parallelism the caller all-reduces
 * `sums` (d*m+K floats) between drsa_step and drsa_finish_step.
 *
 *   A, C      [M, d]   fp32 (DRSA_PREC_FP32) or scaled fp16 from drsa_pack_f16 (TC modes 1, 2), or
 *             [2, M, d] fp16 hi plane + lo plane from drsa_pack_f16_hilo (TC modes 3, 4)
 *   U         [d, m]   fp32, m = K*d_k <= d               (FP32 mode; may be NULL in TC mode)
 *   Ut_hi/lo  [m, d]   fp16 split of U^T written by drsa_finish_step / drsa_split_u
 *                      (TC modes; may be NULL in FP32 mode; Ut_lo is not read by DRSA_PREC_TC_F16)
 *   scaleA, scaleC      the pack scales of A and C (powers of two; 1 in FP32 mode)
 *   pq_scale            power of two applied to g before P = g*HC, Q = g*HA are rounded to
 *                       fp16 for the gradient GEMM (TC mode; undone exactly in the output)
 */
int drsa_step(const void* A, const void* C, const float* U, const void* Ut_hi, const void* Ut_lo,
              int64_t M, int d, int m, int K, int precision, float scaleA, float scaleC,
              float pq_scale, float* sums, void* workspace, int64_t workspace_bytes, void* stream);

/* out[i] = a[i] + beta * b[i] over n floats (out may alias a or b).  Used for the DEFERRED CORRECTION of the row rounding
 * (SubspaceOptimizer precision 'tc_dc'): every few steps the row sums are evaluated twice at the same U, with single-plane
 * rows (DRSA_PREC_TC_F16) and with hi + lo rows (DRSA_PREC_TC_F16_AC2); delta = sums_hilo - sums_hi is kept and added to the
 * cheap single-plane sums of the following steps.  At a fixed point of the iteration the correction is exact, so the
 * optimisation converges to the optimum of the 22-bit data set at (n + 2)/n of the single-plane cost. */
int drsa_sums_combine(const float* a, const float* b, float beta, float* out, int64_t n, void* stream);

/* U [d,m] fp32 -> Ut_hi, Ut_lo [m,d] fp16 with U^T = hi + lo (+ O(2^-22)); Ut_lo may be NULL. */
int drsa_split_u(const float* U, int d, int m, void* Ut_hi, void* Ut_lo, void* stream);

int64_t drsa_finish_workspace_bytes(int d, int m);

/*
 * Everything after the row pass (replicated on every rank):
 *   q_k = sqrt(sums_k / M_global), obj = (mean_k sqrt(q_k))^2       drsa.py:236-237
 *   obj_log[log_index] = obj                                          drsa.py:104
 *   if U_out != NULL:
 *     Y = U + sqrt(obj) / (K M q_k^1.5) * X_k      (ascent, unit step)  drsa.py:102
 *     U_out = Y (Y^T Y)^(-1/2)                      (polar retraction)   drsa.py:201-221
 *     Ut_hi/Ut_lo (optional, Ut_lo alone may be NULL) = fp16 split of U_out^T for the next TC step
 *   u_rounded != 0: `sums` came from DRSA_PREC_TC_F16 / _F16_AC2, i.e. were evaluated at a rounded matrix U^; the logged
 *     objective is f(U^) + <grad f(U^), U - U^>, which equals f(U) up to second order in the rounding.
 *     u_rounded = 1: U^ = fp16(U).  u_rounded = 2: U^ is the matrix stored in Ut_hi (must be passed, it is read even when
 *     U_out == NULL), and the new Ut_hi is written with ERROR FEEDBACK: hi = fp16(U_out + lo_prev), lo = U_out + lo_prev - hi
 *     with the residual carried in Ut_lo from step to step (initialise both with drsa_split_u).  The rounding errors of
 *     successive steps then cancel instead of accumulating along flat directions of the objective.
 * The polar factor is computed on the device by a scaled Newton-Schulz iteration in
 * fp32 (at most `max_iters` sweeps, stops when ||Y^T Y - I||_F < tol*sqrt(m)); the
 * reference uses an fp64 eigendecomposition on the host, both converge to the same
 * (unique) polar factor.  status (4 ints, device): [0] = sweeps used, [1] = number of calls
 * since the caller last zeroed it whose iteration did not converge within max_iters,
 * [2] = number of concepts with q_k == 0 (the reference yields NaN there),
 * [3] = append cursor: when log_index < 0 the objective is stored at obj_log[status[3]++]
 * so that a captured CUDA graph of one step can be replayed without changing arguments.
 */
int drsa_finish_step(const float* sums, int64_t M_global, const float* U, int d, int m, int K,
                     float* U_out, void* Ut_hi, void* Ut_lo, float* obj_log, int64_t log_index,
                     int max_iters, float tol, int u_rounded, int* status, void* workspace,
                     int64_t workspace_bytes, void* stream);

/*
 * Data-parallel variant: the all-reduce of `sums` over the ranks (one process per GPU) fused into the head of the
 * finish kernel over NVLink / NVSwitch peer memory, instead of an NCCL call between drsa_step and drsa_finish_step.
 * (The reference is single-device, drsa.py:84-104; the sums are additive over row shards)
 *
 *   px->buffers[r]   rank r's exchange buffer of drsa_exchange_bytes(d, m, K, world) bytes as mapped into THIS
 *                    process (buffers[px->rank] is this rank's own); zero-filled once, before the first call, with a
 *                    host barrier between the fill and the first call.  Symmetric memory from
 *                    torch.distributed._symmetric_memory or buffers shared with drsa_ipc_* both work.
 *
 * Every CTA pushes its share of this rank's sums into all peers' buffers as 8-byte {value, flag} words (the flag is the
 * exchange number, so the data is its own arrival signal: no fences, no counters), polls the peers' words and adds
 * the shares in rank order, so all ranks obtain bit-identical results.  All ranks must make the same sequence of
 * calls (same shapes, same U_out == NULL pattern).  Asynchronous, capturable in a CUDA graph; a peer that does not
 * arrive within 30 s makes the kernel trap (launch failure) rather than hang.  Shapes: d, m multiples of 32 (the
 * fused finish kernel); drsa_exchange_bytes returns DRSA_ERR_SHAPE otherwise, and the caller keeps NCCL.
 */
#define DRSA_MAX_PEERS 8
typedef struct drsa_peer_exchange {
  int world;                       /* number of ranks, 2..DRSA_MAX_PEERS */
  int rank;                        /* this process's rank */
  void* buffers[DRSA_MAX_PEERS];   /* device addresses valid in this process */
} drsa_peer_exchange;

int64_t drsa_exchange_bytes(int d, int m, int K, int world);

int drsa_finish_step_p2p(const drsa_peer_exchange* px, const float* sums, int64_t M_global, const float* U,
                         int d, int m, int K, float* U_out, void* Ut_hi, void* Ut_lo, float* obj_log,
                         int64_t log_index, int max_iters, float tol, int u_rounded, int* status,
                         void* workspace, int64_t workspace_bytes, void* stream);

/* cudaIpc plumbing for exchange buffers when torch's symmetric memory is unavailable.  These are the only entry
 * points of the library that allocate: drsa_ipc_alloc returns a zero-filled cudaMalloc'ed buffer and its 64-byte
 * cudaIpcMemHandle_t (to be sent to the peers through any host channel); drsa_ipc_open maps a peer's buffer;
 * drsa_ipc_close / drsa_ipc_free undo them (close every mapping before the owner frees). */
int drsa_ipc_alloc(int64_t bytes, void** ptr, unsigned char* handle64);
int drsa_ipc_open(const unsigned char* handle64, void** ptr);
int drsa_ipc_close(void* ptr);
int drsa_ipc_free(void* ptr);

/* orthogonalize(U) of drsa.py:201-221 on its own: U_out = Y (Y^T Y)^(-1/2). */
int drsa_polar_retract(const float* Y, int d, int m, float* U_out, int max_iters, float tol,
                       int* status, void* workspace, int64_t workspace_bytes, void* stream);

/* NON-DEFAULT retraction, BASELINE.json north_star (3): U_out = Q of the thin QR factorisation of Y with diag(R) > 0 (the
 * result of a Householder QR with sign fix; one CTA, Gram-Schmidt with re-orthogonalisation).  The reference retracts with
 * the POLAR factor (drsa.py:201-221); QR yields a different U every step and a different optimisation trajectory
 * (SURVEY F1), so this exists for comparison only.  drsa_finish_step_qr = drsa_finish_step (fp32 sums, single rank,
 * u_rounded = 0) with this retraction; workspace as drsa_finish_workspace_bytes; status[1] counts rank-deficient inputs. */
int drsa_qr_retract(const float* Y, int d, int m, float* U_out, int* status, void* workspace, int64_t workspace_bytes,
                    void* stream);
int drsa_finish_step_qr(const float* sums, int64_t M_global, const float* U, int d, int m, int K, float* U_out,
                        float* obj_log, int64_t log_index, int* status, void* workspace, int64_t workspace_bytes,
                        void* stream);

/* Per-instance concept relevances, no ReLU, summed over positions
 * (cxai/xai/explain/explainer.py:206-242): out[b][k] = sum_p sum_{j in k} (a U)_j (c U)_j
 *   act, ctx [B, P, d] fp32, U [d, m], out [B, K]. */
int drsa_subspace_relevances(const float* act, const float* ctx, const float* U, int64_t B,
                             int64_t P, int d, int m, int K, float* out, void* workspace,
                             int64_t workspace_bytes, void* stream);
int64_t drsa_subspace_relevances_workspace_bytes(int64_t B, int64_t P, int d, int m);

/* DRSA objective (drsa.py:123-155, 224-238) of S subsets of R consecutive rows each in ONE pass -- the prototype search of
 * cxai/xai/drsa/prototypes.py:98-119 evaluates obj_val once per subset of n samples:
 *   sumsq[s][k] = sum_{r in subset s} relu(sum_{j in k} (a_r U)_j (c_r U)_j)^2,  obj[s] = (mean_k sqrt(sqrt(sumsq[s][k] / R)))^2
 *   act, ctx [S*R, d] fp32, U [d, m], obj [S], sumsq [S, K]; workspace as drsa_subspace_relevances_workspace_bytes(S, R, d, m). */
int drsa_subset_objectives(const float* act, const float* ctx, const float* U, int64_t S, int64_t R, int d, int m, int K,
                           float* obj, float* sumsq, void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Stage 1 helpers: context vectors and normalisation
 *   reference: cxai/xai/drsa/preprocessing.py compute_context_vectors :179-193,
 *              normalize_vectors :219-231, get_vectors_from_maps :234-256
 * ---------------------------------------------------------------------------------- */

/* Gather channel vectors from NCHW maps and form the context vector in one pass:
 *   act_out[(n*L+l)][c] = a_map[n][c][pos]          pos = idx[n*L+l] (or l if idx==NULL, L==HW)
 *   ctx_out[(n*L+l)][c] = R_map[n][c][pos] / (a_map[n][c][pos] + 1e-7)
 * and accumulates sum(act^2), sum(ctx^2) into sumsq[0], sumsq[1] (fp64) for
 * normalize_vectors.  (corrected [N*L, d] row layout) */
int drsa_context_gather(const float* a_map, const float* R_map, int64_t N, int d, int HW,
                        const int64_t* idx, int L, float* act_out, float* ctx_out, double* sumsq,
                        void* stream);

/* The same from the NHWC layout of the tensor-core conv stack: activations as hi + lo fp16 planes [N, HW, Cp] (a = hi + lo),
 * relevance fp32 [N, HW, Cp]; rows are positions, so nothing is transposed.  idx [N, L] int64 selects positions (NULL: all,
 * L == HW).  Replaces get_vectors_from_maps + compute_context_vectors + the reduction of normalize_vectors
 * (preprocessing.py:179-256) together with the two NHWC -> NCHW conversions get_intermediate would need first. */
int drsa_context_pairs_nhwc(const void* a_hi, const void* a_lo, const float* R, int64_t N, int HW, int Cp, int d,
                            const int64_t* idx, int L, float* act_out, float* ctx_out, double* sumsq, void* stream);

/* out = R / (a + 1e-7) element-wise over `count` floats (compute_context_vectors, any layout). */
int drsa_context_vectors(const float* a, const float* R, int64_t count, float* out, void* stream);

/* *out = sum v[i]^2 in fp64 (the statistic of normalize_vectors; all-reduce it across ranks). */
int drsa_sumsq(const float* v, int64_t count, double* out, void* stream);

/* v *= 1 / sqrt(sumsq / count_global) / d^0.25  (normalize_vectors with the global statistic). */
int drsa_normalize(float* v, int64_t rows, int d, const double* sumsq, int64_t count_global,
                   void* stream);

/* ------------------------------------------------------------------------------------
 * Stage 1: LRP pass through the log-mel CNN (zennit 0.5.1 rule semantics)
 *   reference call sites: preprocessing.py:106-176 (get_intermediate),
 *   explain/attribute.py:70-108 (compute_relevances); rules per utils/constants.py:27-51.
 *   Activations are NCHW fp32 exactly as the reference stages them
 *   (utils/dataloading.py:176).
 * ---------------------------------------------------------------------------------- */

/* y = relu?(conv3x3_same(x, w) + b); x [N,Cin,H,W], w [Cout,Cin,3,3], y [N,Cout,H,W]
 * (nn.Conv2d of create_model.py:121-130, BatchNorm already folded into w, b). */
int lrp_conv3x3_forward(const float* x, const float* w, const float* b, int64_t N, int Cin,
                        int Cout, int H, int W, int relu, float* y, void* stream);

/* wt [Cin,Cout,3,3] = w [Cout,Cin,3,3] with channels swapped and taps flipped: the weights
 * of the transposed convolution used by lrp_conv3x3_backward. */
int lrp_conv3x3_flip_weights(const float* w, int Cout, int Cin, float* wt, void* stream);

/* Gamma / ZPlus-style rule for inputs x >= 0 (zennit Gamma collapsed):
 *   z' = conv(x, w') + b',  s = R_out / stabilize(z', eps),  R_in = x * conv_transpose(s, w')
 * w', b' are the caller's modified parameters (w + gamma*max(w,0) ...), wt' the flipped copy
 * of w'.  `s_buf` [N,Cout,H,W] is scratch.  If x_is_ones != 0 the input is replaced by ones
 * in the forward and the input factor is dropped (WSquare / Flat, where w' holds w^2 or
 * ones); x may then be NULL. */
int lrp_conv3x3_backward(const float* x, const float* w_mod, const float* wt_mod, const float* b_mod,
                         const float* R_out, int64_t N, int Cin, int Cout, int H, int W, float eps,
                         int x_is_ones, float* s_buf, float* R_in, void* stream);

/* MaxPool2d(kh,kw) (stride = kernel, floor) forward with arg-max capture (first maximum in
 * row-major window order, like PyTorch), and relevance routing to the arg-max (autograd
 * semantics of an un-hooked MaxPool2d). */
int lrp_maxpool_forward(const float* x, int64_t NC, int H, int W, int kh, int kw, float* y,
                        int32_t* argmax, void* stream);
int lrp_maxpool_backward(const float* R_out, const int32_t* argmax, int64_t NC, int H, int W, int kh,
                         int kw, float* R_in, void* stream);

/* Dense layer forward y = relu?(x w^T + b), x [N,In], w [Out,In]. */
int lrp_dense_forward(const float* x, const float* w, const float* b, int64_t N, int In, int Out,
                      int relu, float* y, void* stream);
/* Epsilon rule: R_in = x * ( (R_out / stabilize(z, eps)) w ),  z = x w^T + b (recomputed
 * into s_buf [N,Out]). */
int lrp_dense_epsilon_backward(const float* x, const float* w, const float* b, const float* R_out,
                               int64_t N, int In, int Out, float eps, float* s_buf, float* R_in,
                               void* stream);
/* R *= (a > 0): autograd of an un-hooked ReLU applied to the relevance flow. */
int lrp_relu_mask(const float* a, float* R, int64_t count, void* stream);

/* ---- tensor-core forward pipeline (tcgen05 implicit GEMM, TMA im2col) ----------------------------------
 * Activations are NHWC, split into two fp16 planes hi + lo with channels padded to a multiple of 64
 * ([B,H,W,Cp] each); weights are [9 taps][Cout_p][Cin_p] hi + lo planes (lrp_tc_split_f16 of the padded,
 * tap-major fp32 weights).  hi*hi + lo*hi + hi*lo keeps fp32-class accuracy. */

/* DRSA_OK if lrp_tc_conv3x3_forward supports the shape (H, W must tile into 128-pixel boxes). */
int lrp_tc_conv3x3_supported(int64_t B, int Cin_p, int Cout_p, int H, int W);

/* y = relu?(conv3x3_same(x, w) + b).  Outputs: the hi/lo NHWC planes for the next layer (may be NULL) and/or
 * an fp32 NCHW copy of the first `Cout` channels (may be NULL).  bias has Cout_p entries (zero padded).
 * err_flag: device int set to 1 if the kernel detects a misaligned shared-memory window, 2 if an output value
 * does not fit the fp16 range of the hi plane (|y| >= 60000). */
int lrp_tc_conv3x3_forward(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo,
                           const float* bias, int64_t B, int H, int W, int Cin_p, int Cout_p, int Cout,
                           int relu, void* y_hi, void* y_lo, float* y_nchw, int* err_flag, void* stream);

/* The same convolution with the MaxPool2d(kh, kw) that follows it (stride = kernel, create_model.py:120-127) fused into
 * the epilogue: y_hi / y_lo are the POOLED planes [B, H/kh, W/kw, Cout_p], argmax_u8 (optional) as in lrp_tc_maxpool.
 * The full-resolution map is never written (for the first block of the genre CNN that is 8.4 MB per sample, written
 * and read again by lrp_tc_maxpool).  kh, kw powers of two; tiles of 8 x 16 pixels (H % 8 == 0, W % 16 == 0) or whole
 * 8 x 8 maps; lrp_tc_conv3x3_pool_supported returns DRSA_OK for shapes this covers.  Only valid when nothing needs
 * the un-pooled activation (no split layer, no un-hooked ReLU mask between the convolution and the pooling). */
int lrp_tc_conv3x3_pool_supported(int64_t B, int Cin_p, int Cout_p, int H, int W, int kh, int kw);
int lrp_tc_conv3x3_forward_pool(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo,
                                const float* bias, int64_t B, int H, int W, int Cin_p, int Cout_p, int Cout,
                                int relu, int kh, int kw, void* y_hi, void* y_lo, void* argmax_u8,
                                int* err_flag, void* stream);

/* LRP backward of the first layer (Cin = 1) under WSquare / Flat (rule maps utils/constants.py:27-51; zennit: input
 * replaced by ones, parameters by w^2 / ones, no input factor) in one pass over the relevance R_out, NHWC fp32
 * [B,H,W,Cp] as it leaves the tensor-core stack:  z = b' + sum of w' over the taps inside the image,
 * s = R_out / stabilize(z, eps), R_in [B,1,H,W] = conv_transpose(s, w').  w_mod [Cout,9], b_mod [Cout] are the
 * rule-modified parameters.  Cp must be 64 (DRSA_ERR_SHAPE otherwise: the caller uses lrp_conv3x3_backward). */
int lrp_tc_first_ones_backward(const float* R_out, const float* w_mod, const float* b_mod, int64_t B, int H, int W,
                               int Cout, int Cp, float eps, float* R_in, void* stream);

/* First layer (Cin = 1, bandwidth bound, CUDA cores): x [B,1,H,W] fp32, w [Cout,9] -> NHWC hi/lo planes. */
int lrp_tc_conv3x3_first(const float* x, const float* w, const float* b, int64_t B, int H, int W, int Cout,
                         int Cout_p, int relu, void* y_hi, void* y_lo, void* stream);

/* MaxPool2d(kh,kw), stride = kernel, on NHWC hi/lo planes.  argmax_u8 (optional) [B,Ho,Wo,Cp] receives the
 * window index dy*kw+dx of the first maximum (PyTorch's tie rule) for the relevance routing. */
int lrp_tc_maxpool(const void* x_hi, const void* x_lo, int64_t B, int H, int W, int Cp, int kh, int kw,
                   void* y_hi, void* y_lo, void* argmax_u8, void* stream);
/* R_in [B,H,W,Cp] fp32 = R_out [B,Ho,Wo,Cp] routed to the arg-max of every window, zeros elsewhere. */
int lrp_tc_maxpool_backward(const float* R_out, const void* argmax_u8, int64_t B, int H, int W, int Cp, int kh,
                            int kw, float* R_in, void* stream);

/* Rule-modified forward of a Gamma / ZPlus / Epsilon conv layer on the tensor cores:
 *   s = 2^k(n) * R_out / stabilize(conv(x, w') + b', eps)       x, s NHWC hi/lo; R_out NHWC fp32 [B,H,W,Cout_p].
 * scale_ref (device, [B], may be NULL = no scaling): per-sample bound on |R_out / z'|; the power of two 2^k(n)
 * derived from it keeps s inside the fp16 range of the hi/lo planes however small the relevance has become
 * (it shrinks by orders of magnitude on the way down).  err_flag is set to 2 if a value leaves that range. */
int lrp_tc_conv3x3_ratio(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo,
                         const float* bias, const float* R_out, int64_t B, int H, int W, int Cin_p,
                         int Cout_p, float eps, const float* scale_ref, void* s_hi, void* s_lo, int* err_flag,
                         void* stream);
/* Backward-data step with the input factor: c = 2^-k(n) * conv(s, wt'), R_in = x * c, where wt' = flipped,
 * channel-swapped w' ([9][Cin_p][Cout_p] hi/lo planes); R_in NHWC fp32 [B,H,W,Cin_p].  scale_ref: the same
 * array the ratio pass was given.  cmax_out (device, [B], zeroed by the caller, may be NULL) receives
 * max |c| per sample = the scale_ref of the next layer below. */
int lrp_tc_conv3x3_inputmul(const void* s_hi, const void* s_lo, const void* wt_hi, const void* wt_lo,
                            const void* x_hi, const void* x_lo, int64_t B, int H, int W, int Cout_p,
                            int Cin_p, const float* scale_ref, float* cmax_out, float* R_in, int* err_flag,
                            void* stream);
/* out[n] = max_i |R[n,i]| / x[n,i] over x[n,i] > 0 (R, x: [B, per_sample] fp32): scale_ref of the first
 * tensor-core layer below the dense head, where R = x * c. */
int lrp_tc_sample_absmax_ratio(const float* R, const float* x, int64_t B, int64_t per_sample, float* out,
                               void* stream);
/* R *= (a_hi + a_lo > 0) on NHWC tensors (un-hooked ReLU). */
int lrp_tc_relu_mask(float* R, const void* a_hi, const void* a_lo, int64_t count, void* stream);
/* Layout hand-over of relevance maps between the NHWC tensor-core stack and the NCHW API surface. */
int lrp_tc_nhwc_f32_to_nchw(const float* x, int64_t B, int H, int W, int Cp, int C, float* y, void* stream);
int lrp_tc_nchw_to_nhwc_f32(const float* x, int64_t B, int H, int W, int C, int Cp, float* y, void* stream);

/* NHWC hi/lo planes -> NCHW fp32 with the first C channels (hand-over to the fp32 / LRP-backward path). */
int lrp_tc_nhwc_to_nchw(const void* x_hi, const void* x_lo, int64_t B, int H, int W, int Cp, int C, float* y,
                        void* stream);

/* hi + lo -> fp32 over `count` elements (hand-over from the NHWC planes to the fp32 projection kernels). */
int lrp_tc_planes_to_f32(const void* x_hi, const void* x_lo, int64_t count, float* out, void* stream);

/* ---- concept-conditional relevance at the split layer -----------------------------------------------------
 * The reference inserts Projection (h = a U), SubspaceFilter (+ SubspaceHook) and InvProjection (a' = h U^T)
 * after the layer where U was optimised (cxai/model/modify_model.py:4-123, cxai/xai/explain/attribute.py:12-67)
 * and attributes every sample K+1 times (explainer.py:68-123).  Position vectors are rows [P, ld] (NHWC view,
 * ld >= d, padded columns zero), U is [d, m] row-major.
 *   lrp_subspace_project: h [P, m] = a U and (optional) a_rec [P, ld] = h U^T, the activation the rest of the
 *     network sees.
 *   lrp_subspace_filter: Epsilon rule on InvProjection (R_h = h * ((R / stab(a_rec)) U)), the SubspaceHook mask,
 *     Epsilon rule on Projection (R_a = a * ((R_h / stab(h)) U^T)) for all K+1 clones at once:
 *     out [(K+1), P, ld], slot 0 = unmasked (standard) relevance, slot k = relevance through concept k only. */
int64_t lrp_subspace_filter_workspace_bytes(int64_t P, int d, int m);
int lrp_subspace_project(const float* a, const float* U, int64_t P, int d, int m, int ld, float* h, float* a_rec,
                         void* stream);
int lrp_subspace_filter(const float* a, const float* h, const float* a_rec, const float* R, const float* U,
                        int64_t P, int d, int m, int K, int ld, float eps_invprojection, float eps_projection,
                        float* out, void* workspace, int64_t workspace_bytes, void* stream);

/* hi = fp16(in), lo = fp16(in - hi) over `count` floats. */
int lrp_tc_split_f16(const float* in, int64_t count, void* hi, void* lo, void* stream);

/* ---- log-mel frontend (Loader.transform_wav, cxai/utils/dataloading.py:138-176) -------------------------------
 * wav [B, n_samples] fp32 (peak-normalised by the caller) -> out [B, 1, n_mels, width] fp32:
 *   torchaudio Spectrogram(n_fft, hop_length, power=None) (periodic hann window of n_fft, centred frames, reflect
 *   padding) -> |.| -> MelScale -> log10(. + 1e-7) -> clamp(min = clamp_min) if `clamp` -> frames first_frame ..
 *   first_frame + width - 1 (the reference keeps frames 1 .. width).
 *   window [n_fft]; dft_basis [n_fft, 2F] row-major with F = n_fft/2 + 1, column f = cos(2 pi f n / n_fft), column
 *   F + f = -sin(...); mel_fb [F, n_mels] (torchaudio.functional.melscale_fbanks).  All three are built once by the
 *   host (cxai.utils.dataloading.Loader). */
int64_t logmel_transform_workspace_bytes(int64_t B, int n_fft, int n_mels, int width);
int logmel_transform_wav(const float* wav, const float* window, const float* dft_basis, const float* mel_fb, int64_t B,
                         int64_t n_samples, int n_fft, int hop_length, int n_mels, int first_frame, int width,
                         int clamp, float clamp_min, float* out, void* workspace, int64_t workspace_bytes,
                         void* stream);

/* ------------------------------------------------------------------------------------
 * Self tests of the tcgen05 / TMA building blocks (used by tests/ on the GPU box).
 * Each returns DRSA_OK and writes max |error| against a CUDA-core computation of the
 * same product into *max_err.
 * ---------------------------------------------------------------------------------- */
int drsa_selftest_umma(int variant, float* max_err_host);

/* Diagnostics: when set to a device buffer of 64 int64 (zeroed by the caller; [15] and [16..55] receive the number of and
 * the %globaltimer stamps at the phase boundaries of the fused finish kernel), CTA 0 of the tensor-core row pass stores the SM
 * cycles its MMA thread spent issuing GEMM1 [0], waiting for the epilogue [1], issuing GEMM2 [2] and its
 * first epilogue warp spent waiting for GEMM1 [3] and working [4].  NULL switches it off (default). */
int drsa_debug_set_tc_profile(void* device_buf6);
/* Diagnostics: registers per thread, max threads per block, static shared bytes, local bytes and the configured
 * dynamic shared-memory limit of the tensor-core row-pass kernel for d in {128, 256} (host ints). */
int drsa_debug_tc_kernel_attrs(int d, int split_u, int* out5);
/* Diagnostics: 0 (default) = the shared-memory-operand row-pass kernel, 1 = DRSA_PREC_TC_F16 with d <= 256 runs the
 * experimental kernel that keeps U^T in tensor memory (correct, measured slower; for A/B comparisons). */
int drsa_debug_set_tc_variant(int variant);
/* Diagnostics / experiment: bit e of `variant` makes the tensor-core convolutions with epilogue e (0 forward, 1 ratio,
 * 2 input-multiply) skip the x_lo * w_hi product, i.e. read their activation operand at 11 instead of 22 bits (default 0). */
int lrp_debug_set_conv_variant(int variant);

#ifdef __cplusplus
}
#endif
#endif /* DRSA_B200_H_ */
