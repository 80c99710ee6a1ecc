// Stage 1 of the hot path: forward pass of the log-mel CNN and the LRP backward pass
// (zennit 0.5.1 rule semantics restated in SURVEY appendix B; the reference drives them through
// preprocessing.py:106-176 and explain/attribute.py:70-108).
//
// One direct 3x3 'same' convolution kernel (NCHW fp32, CUDA cores) serves every conv role of the
// pass through its epilogue:
//    EPI_BIAS       y = acc + b                (+ optional ReLU)              forward
//    EPI_RATIO      y = R / stabilize(acc + b) (the "s" tensor of a rule)     modified forward
//    EPI_INPUT_MUL  y = x * acc                (R_in of Gamma / ZPlus / Epsilon, weights pre-flipped)
//    EPI_PLAIN      y = acc                    (R_in of WSquare / Flat: no input factor)
// The transposed convolution of the backward-data step is the same kernel on weights the host
// passes already flipped and channel-transposed.
// Round-1 implementation note: this is the exact-fp32 CUDA-core version (parity yardstick); the
// tcgen05 implicit-GEMM version replaces it behind the same entry points.
#include "common.cuh"

namespace drsa {

namespace {
enum { EPI_BIAS = 0, EPI_RATIO = 1, EPI_INPUT_MUL = 2, EPI_PLAIN = 3 };

constexpr int TW = 32, TH = 8;      // output tile (pixels) per CTA: 32 wide x 8 high, one pixel per thread
constexpr int OCT = 32;             // output channels per CTA (register accumulators per thread)
constexpr int CIT = 8;              // input channels staged per iteration

template <int EPI, bool ONES>
__global__ void __launch_bounds__(256) conv3x3_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ b, const float* __restrict__ aux,
                                                      int Cin, int Cout, int H, int W, int relu, float eps,
                                                      float* __restrict__ y) {
  __shared__ float sx[CIT][TH + 2][TW + 2 + 2];   // input tile with halo (pad to dodge conflicts)
  __shared__ __align__(16) float sw[CIT][9][OCT];  // weights [ci][tap][oc]
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int tiles_x = (W + TW - 1) / TW;
  const int x0 = (blockIdx.x % tiles_x) * TW, y0 = (blockIdx.x / tiles_x) * TH;
  const int oc0 = blockIdx.y * OCT;
  const int64_t n = blockIdx.z;
  const float* xn = x + n * (int64_t)Cin * H * W;

  float acc[OCT];
#pragma unroll
  for (int o = 0; o < OCT; ++o) acc[o] = 0.f;

  for (int c0 = 0; c0 < Cin; c0 += CIT) {
    // stage the input tile (CIT x 10 x 34) and the weight slab (CIT x 9 x 32)
    for (int i = threadIdx.x; i < CIT * (TH + 2) * (TW + 2); i += 256) {
      const int ci = i / ((TH + 2) * (TW + 2));
      const int rem = i % ((TH + 2) * (TW + 2));
      const int yy = rem / (TW + 2), xx = rem % (TW + 2);
      const int gy = y0 + yy - 1, gx = x0 + xx - 1, c = c0 + ci;
      float v = 0.f;
      if (c < Cin && gy >= 0 && gy < H && gx >= 0 && gx < W) v = ONES ? 1.f : __ldg(xn + ((int64_t)c * H + gy) * W + gx);
      sx[ci][yy][xx] = v;
    }
    for (int i = threadIdx.x; i < CIT * 9 * OCT; i += 256) {
      const int oc = i % OCT, tap = (i / OCT) % 9, ci = i / (OCT * 9);
      const int c = c0 + ci, o = oc0 + oc;
      sw[ci][tap][oc] = (c < Cin && o < Cout) ? __ldg(w + ((int64_t)o * Cin + c) * 9 + tap) : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int ci = 0; ci < CIT; ++ci) {
      float v[9];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) v[ky * 3 + kx] = sx[ci][ty + ky][tx + kx];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
        for (int o4 = 0; o4 < OCT / 4; ++o4) {
          const float4 wv = *reinterpret_cast<const float4*>(&sw[ci][tap][4 * o4]);
          acc[4 * o4 + 0] = fmaf(v[tap], wv.x, acc[4 * o4 + 0]);
          acc[4 * o4 + 1] = fmaf(v[tap], wv.y, acc[4 * o4 + 1]);
          acc[4 * o4 + 2] = fmaf(v[tap], wv.z, acc[4 * o4 + 2]);
          acc[4 * o4 + 3] = fmaf(v[tap], wv.w, acc[4 * o4 + 3]);
        }
      }
    }
    __syncthreads();
  }

  const int gy = y0 + ty, gx = x0 + tx;
  if (gy >= H || gx >= W) return;
#pragma unroll
  for (int o = 0; o < OCT; ++o) {
    const int oc = oc0 + o;
    if (oc >= Cout) break;
    const int64_t idx = ((n * Cout + oc) * H + gy) * W + gx;
    float v = acc[o];
    if (EPI == EPI_BIAS) {
      v += (b != nullptr ? __ldg(b + oc) : 0.f);
      if (relu) v = fmaxf(v, 0.f);
    } else if (EPI == EPI_RATIO) {
      v += (b != nullptr ? __ldg(b + oc) : 0.f);
      v = __ldg(aux + idx) / stabilize(v, eps);
    } else if (EPI == EPI_INPUT_MUL) {
      v *= __ldg(aux + idx);
    }
    y[idx] = v;
  }
}

template <int EPI, bool ONES>
int launch_conv(const float* x, const float* w, const float* b, const float* aux, int64_t N, int Cin, int Cout, int H,
                int W, int relu, float eps, float* y, cudaStream_t stream) {
  if (N > 65535) return DRSA_ERR_SHAPE;
  dim3 grid(cdiv(W, TW) * cdiv(H, TH), cdiv(Cout, OCT), (unsigned)N);
  conv3x3_kernel<EPI, ONES><<<grid, 256, 0, stream>>>(x, w, b, aux, Cin, Cout, H, W, relu, eps, y);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

// w [Cout][Cin][3][3] -> wt [Cin][Cout][3][3] with taps flipped: the weights of the transposed conv
__global__ void flip_weights_kernel(const float* __restrict__ w, int Cout, int Cin, float* __restrict__ wt) {
  const int total = Cout * Cin * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % 9, ci = (i / 9) % Cin, co = i / (9 * Cin);
    wt[((int64_t)ci * Cout + co) * 9 + (8 - tap)] = w[i];
  }
}

__global__ void maxpool_fwd_kernel(const float* __restrict__ x, int64_t NC, int H, int W, int kh, int kw, int Ho, int Wo,
                                   float* __restrict__ y, int32_t* __restrict__ argmax) {
  const int64_t total = NC * Ho * Wo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xo = (int)(i % Wo), yo = (int)((i / Wo) % Ho);
    const int64_t nc = i / ((int64_t)Wo * Ho);
    const float* p = x + nc * H * W;
    float best = -INFINITY; int bi = (yo * kh) * W + xo * kw;
    for (int dy = 0; dy < kh; ++dy)
      for (int dx = 0; dx < kw; ++dx) {
        const int idx = (yo * kh + dy) * W + xo * kw + dx;
        const float v = __ldg(p + idx);
        if (v > best || (v != v)) { best = v; bi = idx; }      // first maximum in row-major window order (PyTorch)
      }
    y[i] = best;
    argmax[i] = bi;
  }
}

__global__ void maxpool_bwd_kernel(const float* __restrict__ R_out, const int32_t* __restrict__ argmax, int64_t NC, int HW,
                                   int HoWo, float* __restrict__ R_in) {
  const int64_t total = NC * HoWo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t nc = i / HoWo;
    R_in[nc * HW + argmax[i]] = R_out[i];     // windows do not overlap (stride == kernel): no atomics needed
  }
}

__global__ void bias_act_kernel(float* __restrict__ y, const float* __restrict__ b, int64_t rows, int cols, int relu) {
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float v = y[i] + (b != nullptr ? __ldg(b + (i % cols)) : 0.f);
    y[i] = relu ? fmaxf(v, 0.f) : v;
  }
}

// s = R / stabilize(z + b, eps) in place over z
__global__ void ratio_kernel(float* __restrict__ z, const float* __restrict__ b, const float* __restrict__ R, int64_t rows,
                             int cols, float eps) {
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = z[i] + (b != nullptr ? __ldg(b + (i % cols)) : 0.f);
    z[i] = __ldg(R + i) / stabilize(v, eps);
  }
}

__global__ void mul_kernel(float* __restrict__ y, const float* __restrict__ x, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] *= __ldg(x + i);
}

__global__ void relu_mask_kernel(const float* __restrict__ a, float* __restrict__ R, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (!(__ldg(a + i) > 0.f)) R[i] = 0.f;
}

// y[n][o] = sum_k x[n][k] w[o][k] for few outputs and long rows (the first classifier layer: 64 x 100 outputs,
// K = 4096): one warp per output, both operands read with coalesced float4 loads, fixed reduction order.
__global__ void __launch_bounds__(256) skinny_dense_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           int64_t N, int In, int Out, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= N * Out) return;
  const int64_t n = wid / Out; const int o = (int)(wid % Out);
  const float* xr = x + n * In; const float* wr = w + (int64_t)o * In;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if ((In & 3) == 0) {
    const float4* x4 = reinterpret_cast<const float4*>(xr); const float4* w4 = reinterpret_cast<const float4*>(wr);
    for (int k = lane; k < In / 4; k += 32) {
      const float4 a = __ldg(x4 + k), b = __ldg(w4 + k);
      a0 = fmaf(a.x, b.x, a0); a1 = fmaf(a.y, b.y, a1); a2 = fmaf(a.z, b.z, a2); a3 = fmaf(a.w, b.w, a3);
    }
  } else {
    for (int k = lane; k < In; k += 32) a0 = fmaf(__ldg(xr + k), __ldg(wr + k), a0);
  }
  const float s = warp_sum((a0 + a1) + (a2 + a3));
  if (lane == 0) y[wid] = s;
}

// z = x w^T without bias: skinny shapes go to the warp-per-output kernel, the rest to the tiled SGEMM
int dense_matmul_xwT(const float* x, const float* w, int64_t N, int In, int Out, float* y, cudaStream_t s) {
  if (N * Out <= 65536 && In >= 512) {
    const int64_t warps = N * Out;
    skinny_dense_kernel<<<(int)((warps + 7) / 8), 256, 0, s>>>(x, w, N, In, Out, y);
    DRSA_LAUNCH_CHECK();
    return DRSA_OK;
  }
  GemmDesc g{};
  g.A = x; g.B = w; g.C = y; g.M = (int)N; g.N = Out; g.K = In; g.lda = In; g.ldb = In; g.ldc = Out;
  g.transA = 0; g.transB = 1; g.alpha = 1.f; g.splits = 1;
  return sgemm(g, s);
}

inline int eblocks(int64_t n) {
  int64_t b = (n + 255) / 256;
  return (int)(b > 148 * 8 ? 148 * 8 : (b < 1 ? 1 : b));
}
}  // namespace

int conv3x3_forward(const float* x, const float* w, const float* b, int64_t N, int Cin, int Cout, int H, int W, int relu,
                    float* y, cudaStream_t s) {
  return launch_conv<EPI_BIAS, false>(x, w, b, nullptr, N, Cin, Cout, H, W, relu, 0.f, y, s);
}

int conv3x3_backward(const float* x, const float* w_mod, const float* wt_mod, const float* b_mod, const float* R_out,
                     int64_t N, int Cin, int Cout, int H, int W, float eps, int x_is_ones, float* s_buf, float* R_in,
                     cudaStream_t s) {
  if (x_is_ones) {
    DRSA_TRY((launch_conv<EPI_RATIO, true>(x, w_mod, b_mod, R_out, N, Cin, Cout, H, W, 0, eps, s_buf, s)));
    return launch_conv<EPI_PLAIN, false>(s_buf, wt_mod, nullptr, nullptr, N, Cout, Cin, H, W, 0, 0.f, R_in, s);
  }
  DRSA_TRY((launch_conv<EPI_RATIO, false>(x, w_mod, b_mod, R_out, N, Cin, Cout, H, W, 0, eps, s_buf, s)));
  return launch_conv<EPI_INPUT_MUL, false>(s_buf, wt_mod, nullptr, x, N, Cout, Cin, H, W, 0, 0.f, R_in, s);
}

int flip_weights(const float* w, int Cout, int Cin, float* wt, cudaStream_t s) {
  flip_weights_kernel<<<eblocks((int64_t)Cout * Cin * 9), 256, 0, s>>>(w, Cout, Cin, wt);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

int maxpool_forward(const float* x, int64_t NC, int H, int W, int kh, int kw, float* y, int32_t* argmax, cudaStream_t s) {
  const int Ho = H / kh, Wo = W / kw;
  maxpool_fwd_kernel<<<eblocks(NC * Ho * Wo), 256, 0, s>>>(x, NC, H, W, kh, kw, Ho, Wo, y, argmax);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

int maxpool_backward(const float* R_out, const int32_t* argmax, int64_t NC, int H, int W, int kh, int kw, float* R_in,
                     cudaStream_t s) {
  const int Ho = H / kh, Wo = W / kw;
  DRSA_CUDA(cudaMemsetAsync(R_in, 0, NC * (int64_t)H * W * 4, s));
  maxpool_bwd_kernel<<<eblocks(NC * Ho * Wo), 256, 0, s>>>(R_out, argmax, NC, H * W, Ho * Wo, R_in);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

int dense_forward(const float* x, const float* w, const float* b, int64_t N, int In, int Out, int relu, float* y,
                  cudaStream_t s) {
  DRSA_TRY(dense_matmul_xwT(x, w, N, In, Out, y, s));
  bias_act_kernel<<<eblocks(N * Out), 256, 0, s>>>(y, b, N, Out, relu);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

int dense_epsilon_backward(const float* x, const float* w, const float* b, const float* R_out, int64_t N, int In,
                           int Out, float eps, float* s_buf, float* R_in, cudaStream_t s) {
  DRSA_TRY(dense_matmul_xwT(x, w, N, In, Out, s_buf, s));           // z = x w^T
  ratio_kernel<<<eblocks(N * Out), 256, 0, s>>>(s_buf, b, R_out, N, Out, eps);   // s = R / stabilize(z + b)
  DRSA_LAUNCH_CHECK();
  GemmDesc h{};
  h.A = s_buf; h.B = w; h.C = R_in; h.M = (int)N; h.N = In; h.K = Out; h.lda = Out; h.ldb = In; h.ldc = In;
  h.transA = 0; h.transB = 0; h.alpha = 1.f; h.splits = 1;
  DRSA_TRY(sgemm(h, s));                                            // c = s w
  mul_kernel<<<eblocks(N * In), 256, 0, s>>>(R_in, x, N * In);      // R_in = x * c
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

int relu_mask(const float* a, float* R, int64_t count, cudaStream_t s) {
  relu_mask_kernel<<<eblocks(count), 256, 0, s>>>(a, R, count);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

}  // namespace drsa
