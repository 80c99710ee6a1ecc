// Log-mel frontend on the device (Loader.transform_wav, cxai/utils/dataloading.py:138-176 of the reference):
//   torchaudio Spectrogram(n_fft, hop, power=None: hann window, centred, reflect padding) -> |.| -> MelScale (HTK
//   triangular filters) -> log10(. + 1e-7) -> clamp(min = -4) -> frames 1 .. width -> [B, 1, n_mels, width].
// Only the `width` frames that survive the crop are computed.  The DFT of an n_fft = 800 (or 480) frame is a GEMM
// against a precomputed [n_fft, 2F] basis (cos | -sin), the mel projection a second GEMM; both run on the fp32 SGEMM
// (a 3 s clip is 0.17 GFLOP, nothing next to the CNN), the staging kernels around them use coalesced, vectorised
// accesses: consecutive threads read consecutive samples of the waveform and write consecutive frame elements, the
// final kernel transposes [frame][mel] -> [mel][frame] through shared memory.
#include "common.cuh"

namespace drsa {

namespace {
inline int lblk(int64_t n) {
  int64_t b = (n + 255) / 256;
  return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b));
}

// frames[(b*width + t)][n] = window[n] * wav_b[reflect((t + t0)*hop + n - n_fft/2)]
__global__ void __launch_bounds__(256) logmel_frames_kernel(const float* __restrict__ wav, const float* __restrict__ window,
                                                            int64_t B, int64_t n_samples, int n_fft, int hop, int t0, int width,
                                                            float* __restrict__ frames) {
  const int64_t total = B * width * (int64_t)n_fft;
  const int pad = n_fft / 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % n_fft);
    const int64_t ft = i / n_fft;
    const int t = (int)(ft % width);
    const int64_t b = ft / width;
    int64_t s = (int64_t)(t + t0) * hop + n - pad;
    if (s < 0) s = -s;
    if (s >= n_samples) s = 2 * (n_samples - 1) - s;
    frames[i] = __ldg(window + n) * __ldg(wav + b * n_samples + s);
  }
}

// mag[r][f] = sqrt(re^2 + im^2) with spec[r] = [re_0 .. re_{F-1} | im_0 .. im_{F-1}]
__global__ void __launch_bounds__(256) logmel_mag_kernel(const float* __restrict__ spec, int64_t rows, int F,
                                                         float* __restrict__ mag) {
  const int64_t total = rows * F;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    const int64_t r = i / F;
    const float re = __ldg(spec + r * 2 * F + f), im = __ldg(spec + r * 2 * F + F + f);
    mag[i] = sqrtf(re * re + im * im);
  }
}

// out[b][mel][t] = clamp(log10(mel[b*width + t][mel] + 1e-7), min) -- 32 x 32 shared-memory transpose
__global__ void __launch_bounds__(256) logmel_finish_kernel(const float* __restrict__ mel, int width, int n_mels, int do_clamp,
                                                            float clamp_min, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int t = t0 + ty + 8 * r, m = m0 + tx;
    float v = 0.f;
    if (t < width && m < n_mels) {
      v = log10f(mel[((int64_t)b * width + t) * n_mels + m] + 1e-7f);
      if (do_clamp) v = fmaxf(v, clamp_min);
    }
    tile[ty + 8 * r][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int m = m0 + ty + 8 * r, t = t0 + tx;
    if (m < n_mels && t < width) out[((int64_t)b * n_mels + m) * width + t] = tile[tx][ty + 8 * r];
  }
}
}  // namespace

int64_t logmel_workspace_bytes(int64_t B, int n_fft, int n_mels, int width) {
  const int F = n_fft / 2 + 1;
  const int64_t rows = B * width;
  return align_up(rows * n_fft * 4, 256) + align_up(rows * 2 * F * 4, 256) + align_up(rows * F * 4, 256) +
         align_up(rows * n_mels * 4, 256);
}

int logmel_transform(const float* wav, const float* window, const float* basis, const float* fb, int64_t B, int64_t n_samples,
                     int n_fft, int hop, int n_mels, int first_frame, int width, int do_clamp, float clamp_min, float* out,
                     void* workspace, int64_t workspace_bytes, cudaStream_t s) {
  const int F = n_fft / 2 + 1;
  const int64_t rows = B * width;
  if (rows > 2147483647LL || B > 65535) return DRSA_ERR_SHAPE;
  if (n_samples <= n_fft / 2) return DRSA_ERR_SHAPE;                       // reflect padding needs pad < n_samples
  if ((int64_t)(first_frame + width - 1) * hop > n_samples) return DRSA_ERR_SHAPE;   // frames beyond 1 + n_samples / hop
  if (workspace_bytes < logmel_workspace_bytes(B, n_fft, n_mels, width)) return DRSA_ERR_WORKSPACE;
  char* w = static_cast<char*>(workspace);
  float* frames = reinterpret_cast<float*>(w); w += align_up(rows * n_fft * 4, 256);
  float* spec = reinterpret_cast<float*>(w); w += align_up(rows * 2 * F * 4, 256);
  float* mag = reinterpret_cast<float*>(w); w += align_up(rows * F * 4, 256);
  float* mel = reinterpret_cast<float*>(w);
  logmel_frames_kernel<<<lblk(rows * n_fft), 256, 0, s>>>(wav, window, B, n_samples, n_fft, hop, first_frame, width, frames);
  DRSA_LAUNCH_CHECK();
  GemmDesc g{};
  g.A = frames; g.B = basis; g.C = spec; g.M = (int)rows; g.N = 2 * F; g.K = n_fft; g.lda = n_fft; g.ldb = 2 * F; g.ldc = 2 * F;
  g.transA = 0; g.transB = 0; g.alpha = 1.f; g.splits = 1;
  DRSA_TRY(sgemm(g, s));
  logmel_mag_kernel<<<lblk(rows * F), 256, 0, s>>>(spec, rows, F, mag);
  DRSA_LAUNCH_CHECK();
  GemmDesc h{};
  h.A = mag; h.B = fb; h.C = mel; h.M = (int)rows; h.N = n_mels; h.K = F; h.lda = F; h.ldb = n_mels; h.ldc = n_mels;
  h.transA = 0; h.transB = 0; h.alpha = 1.f; h.splits = 1;
  DRSA_TRY(sgemm(h, s));
  dim3 grid(cdiv(width, 32), cdiv(n_mels, 32), (unsigned)B);
  logmel_finish_kernel<<<grid, 256, 0, s>>>(mel, width, n_mels, do_clamp, clamp_min, out);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

}  // namespace drsa
