// 3x3 'same' convolution as an implicit GEMM on the 5th-generation tensor cores (forward pass of the
// log-mel CNN, create_model.py:100-137 as driven by preprocessing.py:156-162).
//
//   out[pixel][co] = sum_{tap, ci} in[pixel + tap][ci] * w[tap][co][ci]          M = 128 pixels, N = Cout, K = 9*Cin
//
// Layout: activations NHWC, stored as two fp16 planes hi + lo (x = hi + lo up to 2^-22), channels padded to
// a multiple of 64; weights [tap][Cout][Cin] as hi + lo planes.  Three MMAs per k-step
// (hi*hi + lo*hi + hi*lo) keep fp32-class accuracy (the LRP tolerance of 1e-4 does not survive single
// fp16/TF32 operands through ten layers and the R/(z+eps) divisions).
//
// im2col is done by TMA: for tap (ky, kx) the A operand of a tile of 128 output pixels is the 4-D box
// [nb, th, tw, 64 channels] at coordinates shifted by (ky-1, kx-1); out-of-bounds rows/columns are
// zero-filled by the TMA unit, which IS the zero padding of the convolution.  The box lands in shared
// memory as 128 rows of 128 bytes with the 128-byte swizzle, i.e. directly in the K-major operand layout
// of tcgen05.mma.  Accumulators live in TMEM (double buffered when 2*Cout <= 512 columns); the epilogue
// (4 warps, one pixel per thread) adds the bias, applies ReLU and writes the next layer's hi/lo planes.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue (two per TMEM lane quarter, alternating 32-channel chunks); persistent over tiles.
#include <cuda.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace drsa {

namespace {
using namespace tc;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// NHWC fp16 [B, H, W, C] with a box of [nb, th, tw, 64 channels]
int make_tmap_nhwc(CUtensorMap* out, const void* base, uint64_t B, uint64_t H, uint64_t W, uint64_t C, uint32_t nb,
                   uint32_t th, uint32_t tw) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return DRSA_ERR_CUDA;
  cuuint64_t dims[4] = {C, W, H, B};
  cuuint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
  cuuint32_t box[4] = {64, tw, th, nb};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { g_last_cuda_error = (int)r; return DRSA_ERR_CUDA; }
  return DRSA_OK;
}

constexpr int kTilePix = 128;
constexpr int kABytes = kTilePix * 128;        // one [128 pixels x 64 ch] fp16 box
constexpr int kConvThreads = 320;      // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter)
constexpr uint32_t kDescHiK = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint64_t kdesc(uint32_t smem_addr) {
  return ((uint64_t)kDescHiK << 32) | ((smem_addr >> 4) | (1u << 16));
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}

// Power-of-two scale for the fp16 hi/lo planes of s = R / z.  Relevance shrinks by orders of magnitude on the way
// down (biases absorb it), so the unscaled quotient leaves the fp16 range (min subnormal 6e-8) after a few
// layers.  `bound` is a rigorous per-sample bound on |s|: |s| = |R|/z' = a|c|/z' <= |c| because z' >= z >= a for
// Gamma / ZPlus / Epsilon on non-negative input, c being the un-multiplied relevance of the layer above.  The
// scale maps it to [2^11, 2^12), a factor 16 below the fp16 maximum; the pair error is then <= 2^-25 absolute.
__device__ __forceinline__ float pow2_scale(float bound) {
  if (!(bound > 0.f) || !(bound < 3.0e38f)) return 1.f;
  const int e = (int)((__float_as_uint(bound) >> 23) & 0xffu) - 127;
  int k = 11 - e;
  k = k < -100 ? -100 : (k > 100 ? 100 : k);
  return __uint_as_float((uint32_t)(k + 127) << 23);
}

// TMA prefetch of a box into L2 (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

struct ConvGeom {
  int B, H, W, Cin_p, Cout_p, Cout;      // padded channel counts (multiples of 64) and the real Cout
  int nb, th, tw;                        // tile = nb images x th rows x tw columns = 128 pixels
  int tiles_x, tiles_y, tiles_b, num_tiles;
  int relu;
  int epi;      // 0: y = relu?(acc + b); 1: y = aux_f32 / stabilize(acc + b, eps); 2: y = (aux_hi + aux_lo) * acc
  float eps;
  int pkh, pkw;   // > 0: MaxPool2d(pkh, pkw) fused into the epi == 0 epilogue (y planes and arg-max are the POOLED maps)
  int drop_xlo;   // experiment (lrp_debug_set_conv_variant): skip the x_lo * w_hi product (activations at 11 bits)
};

// kHalo (with resident weights, tiles of th = 8 rows x tw = 16 columns of one image): instead of one TMA box per tap
// (9 boxes of 128 pixels per tile and channel chunk, i.e. every activation is fetched from L2 nine times -- the
// kernel was bound by L2 -> SM bandwidth, not by the tensor pipe) a stage holds, for one kx, the box of (th + 2) x tw
// pixels starting one row above the tile.  The three taps (ky, kx) of that stage read it through descriptors whose
// start address is advanced by ky * tw rows = ky * 1024 B, a whole number of 128-byte-swizzle atoms, so the swizzle
// phase of the operand rows is unchanged.  L2 traffic per tile drops from 9 x 32 KB to 3 x 36 KB.
constexpr int kHaloTw = 16, kHaloTh = 8;
constexpr int kHaloBytes = (kHaloTh + 2) * kHaloTw * 128;      // 18 KB: one plane of one kx box

template <int kStages, bool kResW, bool kHalo>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmXh, const __grid_constant__ CUtensorMap tmXl,
                  const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWl, ConvGeom g,
                  const float* __restrict__ bias, const float* __restrict__ aux_f32, const __half* __restrict__ aux_hi,
                  const __half* __restrict__ aux_lo, __half* __restrict__ y_hi, __half* __restrict__ y_lo,
                  float* __restrict__ y_f32, float* __restrict__ y_nchw, const float* __restrict__ scale_ref,
                  float* __restrict__ cmax_out, int* __restrict__ err_flag, uint8_t* __restrict__ amax_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int wbytes = g.Cout_p * 128;                 // one [Cout_p x 64] weight box
  // kResW (Cin_p = Cout_p = 64): all 9 taps of the hi/lo weights (144 KB) stay resident in shared memory and the
  // ring only carries the activation boxes; otherwise the weight boxes travel with the activations.
  const int stage_bytes = kHalo ? 2 * kHaloBytes : (kResW ? 2 * kABytes : 2 * kABytes + 2 * wbytes);
  uint8_t* sW = smem + kStages * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + (kResW ? 18 * wbytes : 0));
  uint64_t* full = bars;                // [kStages]
  uint64_t* empty = bars + 8;           // [kStages]
  uint64_t* acc_full = bars + 16;       // [2]
  uint64_t* acc_empty = bars + 18;      // [2]
  uint64_t* w_full = bars + 20;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // merged: one MMA of N = 2*Cout_p against the stacked [w_hi; w_lo] rows (they are adjacent in shared memory) gives
  // x_hi*w_hi in accumulator columns [0, Cout_p) and x_hi*w_lo in [Cout_p, 2*Cout_p); a second MMA of N = Cout_p adds
  // x_lo*w_hi.  Two MMAs instead of three per k-step and 14 KB instead of 18 KB of operand reads (Cout_p = 64): SS-mode
  // MMAs at N = 64 are bound by shared-memory bandwidth, not by the tensor pipe.  The epilogue adds the two halves.
  const bool merged = 2 * g.Cout_p <= 256;
  const int acc_cols = merged ? 2 * g.Cout_p : g.Cout_p;
  const int acc_stages = (2 * acc_cols <= 512) ? 2 : 1;

  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) atomicExch(err_flag, 1);
    return;
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 256); }
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int kchunks = g.Cin_p / 64;
  const int kiters = 9 * kchunks;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmXh); tma_prefetch_desc(&tmXl); tma_prefetch_desc(&tmWh); tma_prefetch_desc(&tmWl);
      if (kResW) {
        mbar_expect_tx(w_full, 18 * wbytes);
        for (int tap = 0; tap < 9; ++tap) {
          tma_load_2d(sW + (2 * tap) * wbytes, &tmWh, w_full, 0, tap * g.Cout_p);
          tma_load_2d(sW + (2 * tap + 1) * wbytes, &tmWl, w_full, 0, tap * g.Cout_p);
        }
      }
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
        const int txi = tile % g.tiles_x, tyi = (tile / g.tiles_x) % g.tiles_y, tbi = tile / (g.tiles_x * g.tiles_y);
        const int x0 = txi * g.tw, y0 = tyi * g.th, n0 = tbi * g.nb;
        if (kHalo) {
          {   // pull the next tile's rows into L2 while this tile is computed: with two stages of 36 KB the ring cannot
              // hide an HBM round trip, an L2 hit it can (the kx = 1 box covers all but two pixel columns of the three)
            const int nt = tile + gridDim.x;
            if (nt < g.num_tiles) {
              const int ntx = nt % g.tiles_x, nty = (nt / g.tiles_x) % g.tiles_y, ntb = nt / (g.tiles_x * g.tiles_y);
              tma_prefetch_l2_4d(&tmXh, 0, ntx * g.tw, nty * g.th - 1, ntb * g.nb);
              tma_prefetch_l2_4d(&tmXl, 0, ntx * g.tw, nty * g.th - 1, ntb * g.nb);
            }
          }
          for (int kx = 0; kx < 3; ++kx) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], stage_bytes);
            uint8_t* dst = smem + stage * stage_bytes;
            tma_load_4d(dst, &tmXh, &full[stage], 0, x0 + kx - 1, y0 - 1, n0);
            tma_load_4d(dst + kHaloBytes, &tmXl, &full[stage], 0, x0 + kx - 1, y0 - 1, n0);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          continue;
        }
        for (int it = 0; it < kiters; ++it) {
          const int tap = it / kchunks, kc = it % kchunks;
          const int ky = tap / 3, kx = tap % 3;
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], stage_bytes);
          uint8_t* dst = smem + stage * stage_bytes;
          tma_load_4d(dst, &tmXh, &full[stage], 64 * kc, x0 + kx - 1, y0 + ky - 1, n0);
          tma_load_4d(dst + kABytes, &tmXl, &full[stage], 64 * kc, x0 + kx - 1, y0 + ky - 1, n0);
          if (!kResW) {
            tma_load_2d(dst + 2 * kABytes, &tmWh, &full[stage], 64 * kc, tap * g.Cout_p);
            tma_load_2d(dst + 2 * kABytes + wbytes, &tmWl, &full[stage], 64 * kc, tap * g.Cout_p);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(kTilePix, g.Cout_p, 0, 0);
      const uint32_t idesc2 = make_idesc_f16(kTilePix, 2 * g.Cout_p, 0, 0);
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      if (kResW) { mbar_wait(w_full, 0); tc_fence_after(); }
      for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
        mbar_wait(&acc_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * acc_cols;
        if (kHalo) {
          for (int kx = 0; kx < 3; ++kx) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t box_hi = smem_u32(smem + stage * stage_bytes), box_lo = box_hi + kHaloBytes;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const uint32_t a_hi = box_hi + ky * (kHaloTw * 128), a_lo = box_lo + ky * (kHaloTw * 128);
              const uint32_t w_hi = smem_u32(sW + 2 * (ky * 3 + kx) * wbytes);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t dwh = kdesc(w_hi + 32 * kk);
                umma_ss_f16(tacc, kdesc(a_hi + 32 * kk), dwh, idesc2, (kx | ky | kk) ? 1u : 0u);   // [x_hi*w_hi | x_hi*w_lo]
                if (!g.drop_xlo) umma_ss_f16(tacc, kdesc(a_lo + 32 * kk), dwh, idesc, 1u);        // + x_lo*w_hi
              }
            }
            umma_commit(&empty[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          umma_commit(&acc_full[as]);
          if (++as == acc_stages) { as = 0; aphase ^= 1; }
          continue;
        }
        for (int it = 0; it < kiters; ++it) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + stage * stage_bytes), a_lo = a_hi + kABytes;
          const uint32_t w_hi = kResW ? smem_u32(sW + 2 * (it / kchunks) * wbytes) : a_hi + 2 * kABytes, w_lo = w_hi + wbytes;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t dah = kdesc(a_hi + 32 * kk), dal = kdesc(a_lo + 32 * kk);
            const uint64_t dwh = kdesc(w_hi + 32 * kk), dwl = kdesc(w_lo + 32 * kk);
            if (merged) {
              umma_ss_f16(tacc, dah, dwh, idesc2, (it | kk) ? 1u : 0u);      // [x_hi*w_hi | x_hi*w_lo]
              if (!g.drop_xlo) umma_ss_f16(tacc, dal, dwh, idesc, 1u);       // + x_lo*w_hi
            } else {
              umma_ss_f16(tacc, dah, dwh, idesc, (it | kk) ? 1u : 0u);
              if (!g.drop_xlo) umma_ss_f16(tacc, dal, dwh, idesc, 1u);
              umma_ss_f16(tacc, dah, dwl, idesc, 1u);
            }
          }
          umma_commit(&empty[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[as]);
        if (++as == acc_stages) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;                     // the two warps of a lane quarter split the channel chunks
    const int pix = 32 * q + lane;                        // pixel of the tile = TMEM lane
    const uint32_t lane_base = tmem_base + ((uint32_t)(32 * q) << 16);
    int as = 0; uint32_t aphase = 0;
    bool ovf = false;
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
      const int txi = tile % g.tiles_x, tyi = (tile / g.tiles_x) % g.tiles_y, tbi = tile / (g.tiles_x * g.tiles_y);
      const int xx = pix % g.tw, yy = (pix / g.tw) % g.th, bi = pix / (g.tw * g.th);
      const int x = txi * g.tw + xx, y = tyi * g.th + yy, n = tbi * g.nb + bi;
      const bool valid = (x < g.W) && (y < g.H) && (n < g.B);
      // per-sample power-of-two scale of the fp16 planes that carry s = R / z (see pow2_scale)
      const float sc = (scale_ref != nullptr && valid) ? pow2_scale(__ldg(scale_ref + n)) : 1.f;
      const float isc = 1.f / sc;
      float cm = 0.f;
      mbar_wait(&acc_full[as], aphase);
      tc_fence_after();
      const int64_t pbase = (((int64_t)n * g.H + y) * g.W + x) * g.Cout_p;
#pragma unroll 1
      for (int cc = half; cc < g.Cout_p / 32; cc += 2) {
        uint32_t v[32];
        float a[32];
        tmem_ld32(lane_base + as * acc_cols + 32 * cc, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] = __uint_as_float(v[i]);
        if (merged) {
          tmem_ld32(lane_base + as * acc_cols + g.Cout_p + 32 * cc, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) a[i] += __uint_as_float(v[i]);
        }
        if (g.pkh > 0) {
          // Fused MaxPool2d(pkh, pkw), stride = kernel (create_model.py:120): the window of a pooled pixel lies inside
          // this warp's 32 pixels (host picks tw <= 16, so a warp holds 32 / tw >= pkh rows of the tile), i.e. the
          // maximum is a butterfly over the lane bits of x (masks 1 .. pkw/2) and y (masks tw .. tw*pkh/2).  Ties keep
          // the first element in row-major window order like PyTorch.  The full-resolution map is never written.
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + 32 * cc) + i4);
            a[4 * i4] += bb.x; a[4 * i4 + 1] += bb.y; a[4 * i4 + 2] += bb.z; a[4 * i4 + 3] += bb.w;
          }
          if (g.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = fmaxf(a[i], 0.f);
          }
          // Two passes, both with the channel loop innermost so that the 32 shuffles / votes of a stage are independent
          // (a per-channel butterfly that carries the arg-max along serialises 3 dependent shuffles per channel and made
          // the epilogue, not the tensor pipe, the bottleneck): (1) max over the window, (2) arg-max = first lane of the
          // window, in lane order = row-major window order, whose value equals the maximum.
          const int widx = (yy % g.pkh) * g.pkw + (xx % g.pkw);
          const int tw_shift = g.tw == 16 ? 4 : 3;
          const int base = lane - (((lane >> tw_shift) % g.pkh) << tw_shift) - ((lane & (g.tw - 1)) % g.pkw);
          uint32_t wmask = 0u;
          for (int dy = 0; dy < g.pkh; ++dy)
            for (int dx = 0; dx < g.pkw; ++dx) wmask |= 1u << (base + (dy << tw_shift) + dx);
          float vmax[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) vmax[i] = a[i];
          for (int mk = 1; mk < g.pkw; mk <<= 1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) vmax[i] = fmaxf(vmax[i], __shfl_xor_sync(0xffffffffu, vmax[i], mk));
          }
          for (int mk = g.tw; mk < g.tw * g.pkh; mk <<= 1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) vmax[i] = fmaxf(vmax[i], __shfl_xor_sync(0xffffffffu, vmax[i], mk));
          }
          uint32_t packed_idx[8];
          if (amax_out != nullptr) {          // the arg-max is only kept for layers the backward pass will cross
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const uint32_t hit = __ballot_sync(0xffffffffu, a[i] == vmax[i]) & wmask;
              const int rel = hit ? (__ffs(hit) - 1 - base) : 0;
              const uint32_t idx = (uint32_t)((rel >> tw_shift) * g.pkw + (rel & (g.tw - 1)));
              if ((i & 3) == 0) packed_idx[i >> 2] = 0u;
              packed_idx[i >> 2] |= idx << (8 * (i & 3));
            }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) a[i] = vmax[i];
          if (valid && widx == 0) {
            const int Ho = g.H / g.pkh, Wo = g.W / g.pkw;
            const int64_t po = (((int64_t)n * Ho + y / g.pkh) * Wo + x / g.pkw) * g.Cout_p + 32 * cc;
            uint32_t hi[16], lo[16];
            float am = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) am = fmaxf(am, fabsf(a[i]));
            ovf = ovf || !(am < 60000.f);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const __half2 h = __floats2half2_rn(a[2 * i], a[2 * i + 1]);
              const float2 hf = __half22float2(h);
              const __half2 l = __floats2half2_rn(a[2 * i] - hf.x, a[2 * i + 1] - hf.y);
              hi[i] = *reinterpret_cast<const uint32_t*>(&h);
              lo[i] = *reinterpret_cast<const uint32_t*>(&l);
            }
            uint4* ph = reinterpret_cast<uint4*>(y_hi + po);
            uint4* pl = reinterpret_cast<uint4*>(y_lo + po);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              ph[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
              pl[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
            }
            if (amax_out != nullptr) {
              uint4* pa = reinterpret_cast<uint4*>(amax_out + po);
              pa[0] = make_uint4(packed_idx[0], packed_idx[1], packed_idx[2], packed_idx[3]);
              pa[1] = make_uint4(packed_idx[4], packed_idx[5], packed_idx[6], packed_idx[7]);
            }
          }
        } else if (valid) {
          const int64_t o = pbase + 32 * cc;
          if (g.epi != 2) {
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + 32 * cc) + i4);
              a[4 * i4] += bb.x; a[4 * i4 + 1] += bb.y; a[4 * i4 + 2] += bb.z; a[4 * i4 + 3] += bb.w;
            }
          }
          if (g.epi == 0) {
            if (g.relu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) a[i] = fmaxf(a[i], 0.f);
            }
          } else if (g.epi == 1) {            // s = R_out / stabilize(z', eps)
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
              const float4 r = __ldg(reinterpret_cast<const float4*>(aux_f32 + o) + i4);
              a[4 * i4] = r.x / stabilize(a[4 * i4], g.eps) * sc;
              a[4 * i4 + 1] = r.y / stabilize(a[4 * i4 + 1], g.eps) * sc;
              a[4 * i4 + 2] = r.z / stabilize(a[4 * i4 + 2], g.eps) * sc;
              a[4 * i4 + 3] = r.w / stabilize(a[4 * i4 + 3], g.eps) * sc;
            }
          } else {                            // c = conv(s, wt') / scale;  R_in = x * c
#pragma unroll
            for (int i = 0; i < 32; ++i) { a[i] *= isc; cm = fmaxf(cm, fabsf(a[i])); }
#pragma unroll
            for (int i8 = 0; i8 < 4; ++i8) {
              const uint4 h = __ldg(reinterpret_cast<const uint4*>(aux_hi + o) + i8);
              const uint4 l = __ldg(reinterpret_cast<const uint4*>(aux_lo + o) + i8);
              const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[j]));
                const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[j]));
                a[8 * i8 + 2 * j] *= hf.x + lf.x;
                a[8 * i8 + 2 * j + 1] *= hf.y + lf.y;
              }
            }
          }
          if (y_nchw != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int c = 32 * cc + i;
              if (c < g.Cout) y_nchw[(((int64_t)n * g.Cout + c) * g.H + y) * g.W + x] = a[i];
            }
          }
          if (y_f32 != nullptr) {
            float4* pf = reinterpret_cast<float4*>(y_f32 + o);
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) pf[i4] = make_float4(a[4 * i4], a[4 * i4 + 1], a[4 * i4 + 2], a[4 * i4 + 3]);
          }
          if (y_hi != nullptr) {
            uint32_t hi[16], lo[16];
            float am = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) am = fmaxf(am, fabsf(a[i]));
            ovf = ovf || !(am < 60000.f);           // fp16 range of the hi plane
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const __half2 h = __floats2half2_rn(a[2 * i], a[2 * i + 1]);
              const float2 hf = __half22float2(h);
              const __half2 l = __floats2half2_rn(a[2 * i] - hf.x, a[2 * i + 1] - hf.y);
              hi[i] = *reinterpret_cast<const uint32_t*>(&h);
              lo[i] = *reinterpret_cast<const uint32_t*>(&l);
            }
            uint4* ph = reinterpret_cast<uint4*>(y_hi + o);
            uint4* pl = reinterpret_cast<uint4*>(y_lo + o);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              ph[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
              pl[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[as]);
      if (++as == acc_stages) { as = 0; aphase ^= 1; }
      if (cmax_out != nullptr) {              // per-sample max |c|: the bound on |s| of the layer below
        uint32_t bits = valid ? __float_as_uint(cm) : 0u;
        const int n0 = __shfl_sync(0xffffffffu, n, 0);
        if (__all_sync(0xffffffffu, n == n0)) {
          bits = __reduce_max_sync(0xffffffffu, bits);
          if (lane == 0 && bits != 0u && n0 < g.B) atomicMax(reinterpret_cast<unsigned int*>(cmax_out) + n0, bits);
        } else if (bits != 0u) {
          atomicMax(reinterpret_cast<unsigned int*>(cmax_out) + n, bits);
        }
      }
    }
    if (ovf) atomicExch(err_flag, 2);
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// First layer (Cin = 1): bandwidth-bound, CUDA cores.  x [B,1,H,W] fp32 -> y NHWC hi/lo [B,H,W,Cout_p].
// A thread owns one group of 8 output channels for the whole launch (its 72 weights and 8 biases live in registers)
// and walks over pixels; the Cout_p/8 threads of a pixel are adjacent, so a warp stores 512 contiguous bytes per
// plane (Cout_p = 64).  A CTA takes one image row at a time: all index arithmetic inside the loops is 32-bit and
// division-free (the first version of this kernel spent most of its instructions on 64-bit div/mod).
__global__ void __launch_bounds__(256) conv3x3_first_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                 const float* __restrict__ b, int64_t B, int H, int W,
                                                                 int Cout, int Cout_p, int relu, __half* __restrict__ y_hi,
                                                                 __half* __restrict__ y_lo) {
  const int groups = Cout_p / 8;                       // host guarantees groups <= 256 and 256 % groups == 0
  const int gq = threadIdx.x % groups, px = threadIdx.x / groups, ppi = 256 / groups;   // pixels per iteration
  float wr[9][8], br[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int ch = gq * 8 + c;
    br[c] = ch < Cout ? __ldg(b + ch) : 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[t][c] = ch < Cout ? __ldg(w + ch * 9 + t) : 0.f;
  }
  const int64_t rows = B * H;
  // Two adjacent pixels per thread and iteration: the 3 x 4 input window serves both (12 loads instead of 18) and the
  // two independent accumulator sets double the instruction-level parallelism -- with 72 weights in registers only two
  // CTAs fit an SM, and ncu showed the one-pixel version latency bound (issue slots 50 % busy, DRAM 30 %).
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const int yy = (int)(row % H);
    const float* xr = x + row * W;                     // row yy of image n; rows above / below are +-W away
    const bool up = yy > 0, down = yy + 1 < H;
    __half* oh = y_hi + row * W * (int64_t)Cout_p;
    __half* ol = y_lo + row * W * (int64_t)Cout_p;
    for (int x0 = 0; x0 < W; x0 += 2 * ppi) {
      const int xx = x0 + 2 * px;
      if (xx >= W) continue;
      const bool two = xx + 1 < W;
      float v[3][4];                                   // rows yy-1..yy+1, columns xx-1..xx+2
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) {
        const int gx = xx + kx - 1;
        const bool in = gx >= 0 && gx < W;
        v[0][kx] = (in && up) ? __ldg(xr - W + gx) : 0.f;
        v[1][kx] = in ? __ldg(xr + gx) : 0.f;
        v[2][kx] = (in && down) ? __ldg(xr + W + gx) : 0.f;
      }
      float acc[2][8];
#pragma unroll
      for (int c = 0; c < 8; ++c) { acc[0][c] = br[c]; acc[1][c] = br[c]; }
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            acc[0][c] = fmaf(v[ky][kx], wr[ky * 3 + kx][c], acc[0][c]);
            acc[1][c] = fmaf(v[ky][kx + 1], wr[ky * 3 + kx][c], acc[1][c]);
          }
#pragma unroll
      for (int p2 = 0; p2 < 2; ++p2) {
        if (p2 == 1 && !two) break;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a = acc[p2][2 * j], c = acc[p2][2 * j + 1];
          if (relu) { a = fmaxf(a, 0.f); c = fmaxf(c, 0.f); }
          const __half2 h = __floats2half2_rn(a, c);
          const float2 hf = __half22float2(h);
          const __half2 l = __floats2half2_rn(a - hf.x, c - hf.y);
          hi[j] = *reinterpret_cast<const uint32_t*>(&h);
          lo[j] = *reinterpret_cast<const uint32_t*>(&l);
        }
        const int o = (xx + p2) * Cout_p + gq * 8;
        *reinterpret_cast<uint4*>(oh + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(ol + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
  }
}

// LRP backward of the first layer (Cin = 1) under WSquare / Flat (constants.py:27-51: the input is replaced by ones,
// the parameters by w^2 / ones, and there is no input factor), for relevance arriving from the tensor-core stack as NHWC
// fp32 [B,H,W,64]:
//   z[co](y',x')  = b'[co] + sum of w'[co][tap] over the taps that do not fall into the zero padding
//   s             = R_out / stabilize(z, eps)            (as R_out * (1 / stabilize(z, eps)): z has 16 x 64 distinct values)
//   R_in(y,x)     = sum_tap c[tap](y - ky + 1, x - kx + 1),   c[tap](y',x') = sum_co w'[co][tap] s[co](y',x')
// z only depends on the channel and on which of the four image borders the pixel touches (16 classes), so the whole
// rule is ONE pass over R_out: a CTA takes an 8 x 32 tile of R_in, four threads share a source pixel of the
// (8+2) x (32+2) halo (16 channels each: 4 coalesced 16-byte loads, 144 FMAs against weights read from shared memory as
// float4), the nine tap sums go through shared memory and every output pixel gathers its nine neighbours.  Replaces
// two generic NCHW fp32 convolutions (1 -> 64 and 64 -> 1 channels) and a layout conversion: 7.9 ms -> see DESIGN 3.2.
constexpr int kFbTh = 8, kFbTw = 32, kFbSrc = (kFbTh + 2) * (kFbTw + 2);
__global__ void __launch_bounds__(512) first_layer_ones_bwd_kernel(const float* __restrict__ R_out, const float* __restrict__ w_mod,
                                                                   const float* __restrict__ b_mod, int H, int W, int Cout,
                                                                   float eps, float* __restrict__ R_in) {
  __shared__ __align__(16) float ws[9][64];
  __shared__ __align__(16) float zt[16][64];
  __shared__ float contrib[9][kFbSrc];
  const int tid = threadIdx.x;
  const int tiles_x = (W + kFbTw - 1) / kFbTw, tiles_y = (H + kFbTh - 1) / kFbTh;
  const int tx = blockIdx.x % tiles_x, ty = (blockIdx.x / tiles_x) % tiles_y;
  const int64_t n = blockIdx.x / (tiles_x * tiles_y);
  const int x0 = tx * kFbTw, y0 = ty * kFbTh;
  for (int i = tid; i < 9 * 64; i += blockDim.x) {
    const int tap = i / 64, co = i % 64;
    ws[tap][co] = co < Cout ? __ldg(w_mod + co * 9 + tap) : 0.f;
  }
  __syncthreads();
  for (int i = tid; i < 16 * 64; i += blockDim.x) {
    const int cls = i / 64, co = i % 64;
    float z = co < Cout ? __ldg(b_mod + co) : 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int ky = tap / 3, kx = tap % 3;
      const bool pad = (ky == 0 && (cls & 1)) || (ky == 2 && (cls & 2)) || (kx == 0 && (cls & 4)) || (kx == 2 && (cls & 8));
      if (!pad) z += ws[tap][co];
    }
    zt[cls][co] = 1.f / stabilize(z, eps);      // the reciprocal once per class: 16 divisions per source pixel were a third of the kernel
  }
  __syncthreads();
  const int q = tid & 3;                                   // channels 16 q .. 16 q + 15
  for (int it = 0; it < (kFbSrc + 127) / 128; ++it) {      // same trip count for every thread: the shuffles below need full warps
    const int sp = it * 128 + (tid >> 2);
    const int sy = y0 - 1 + sp / (kFbTw + 2), sx = x0 - 1 + sp % (kFbTw + 2);
    float c[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) c[t] = 0.f;
    if (sp < kFbSrc && sy >= 0 && sy < H && sx >= 0 && sx < W) {
      const int cls = (sy == 0 ? 1 : 0) | (sy == H - 1 ? 2 : 0) | (sx == 0 ? 4 : 0) | (sx == W - 1 ? 8 : 0);
      const float4* src = reinterpret_cast<const float4*>(R_out + ((n * H + sy) * W + sx) * 64 + 16 * q);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 r = __ldg(src + i);
        const float4 z = *reinterpret_cast<const float4*>(&zt[cls][16 * q + 4 * i]);
        const float s0 = r.x * z.x, s1 = r.y * z.y, s2 = r.z * z.z, s3 = r.w * z.w;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float4 w4 = *reinterpret_cast<const float4*>(&ws[t][16 * q + 4 * i]);
          c[t] = fmaf(w4.x, s0, fmaf(w4.y, s1, fmaf(w4.z, s2, fmaf(w4.w, s3, c[t]))));
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      c[t] += __shfl_xor_sync(0xffffffffu, c[t], 1);
      c[t] += __shfl_xor_sync(0xffffffffu, c[t], 2);
    }
    if (q == 0 && sp < kFbSrc) {
#pragma unroll
      for (int t = 0; t < 9; ++t) contrib[t][sp] = c[t];
    }
  }
  __syncthreads();
  for (int o = tid; o < kFbTh * kFbTw; o += blockDim.x) {
    const int oy = o / kFbTw, ox = o % kFbTw;
    const int y = y0 + oy, x = x0 + ox;
    if (y >= H || x >= W) continue;
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)          // source pixel (y - ky + 1, x - kx + 1) -> halo-tile index (oy - ky + 2, ox - kx + 2)
        acc += contrib[ky * 3 + kx][(oy - ky + 2) * (kFbTw + 2) + (ox - kx + 2)];
    R_in[(n * H + y) * W + x] = acc;
  }
}

// MaxPool2d(kh, kw), stride = kernel, on NHWC hi/lo planes; one thread per (output pixel, 8 channels).
__global__ void __launch_bounds__(256) maxpool_nhwc_kernel(const __half* __restrict__ x_hi, const __half* __restrict__ x_lo,
                                                           int64_t B, int H, int W, int Cp, int kh, int kw,
                                                           __half* __restrict__ y_hi, __half* __restrict__ y_lo,
                                                           uint8_t* __restrict__ amax) {
  const int Ho = H / kh, Wo = W / kw, groups = Cp / 8;
  const int64_t total = B * Ho * Wo * groups;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int gq = (int)(i % groups);
    const int64_t p = i / groups;
    const int xo = (int)(p % Wo), yo = (int)((p / Wo) % Ho);
    const int64_t n = p / ((int64_t)Wo * Ho);
    float best[8];
    unsigned char bidx[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { best[e] = -INFINITY; bidx[e] = 0; }
    for (int dy = 0; dy < kh; ++dy)
      for (int dx = 0; dx < kw; ++dx) {
        const int64_t o = (((n * H + yo * kh + dy) * W) + xo * kw + dx) * Cp + gq * 8;
        const uint4 h = __ldg(reinterpret_cast<const uint4*>(x_hi + o));
        const uint4 l = __ldg(reinterpret_cast<const uint4*>(x_lo + o));
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[j]));
          const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[j]));
          const float v0 = hf.x + lf.x, v1 = hf.y + lf.y;
          // strict '>' keeps the first maximum in row-major window order (PyTorch's choice)
          if (v0 > best[2 * j]) { best[2 * j] = v0; bidx[2 * j] = (unsigned char)(dy * kw + dx); }
          if (v1 > best[2 * j + 1]) { best[2 * j + 1] = v1; bidx[2 * j + 1] = (unsigned char)(dy * kw + dx); }
        }
      }
    if (amax != nullptr) {
      uint2 packed;
      packed.x = bidx[0] | (bidx[1] << 8) | (bidx[2] << 16) | ((uint32_t)bidx[3] << 24);
      packed.y = bidx[4] | (bidx[5] << 8) | (bidx[6] << 16) | ((uint32_t)bidx[7] << 24);
      *reinterpret_cast<uint2*>(amax + p * Cp + gq * 8) = packed;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __half2 h = __floats2half2_rn(best[2 * j], best[2 * j + 1]);
      const float2 hf = __half22float2(h);
      const __half2 l = __floats2half2_rn(best[2 * j] - hf.x, best[2 * j + 1] - hf.y);
      hi[j] = *reinterpret_cast<const uint32_t*>(&h);
      lo[j] = *reinterpret_cast<const uint32_t*>(&l);
    }
    const int64_t o = p * Cp + gq * 8;
    *reinterpret_cast<uint4*>(y_hi + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(y_lo + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// Relevance routing of an un-hooked MaxPool2d on NHWC fp32: R_in[window arg-max] = R_out, zeros elsewhere.
__global__ void __launch_bounds__(256) maxpool_nhwc_bwd_kernel(const float* __restrict__ R_out, const uint8_t* __restrict__ amax,
                                                               int64_t B, int H, int W, int Cp, int kh, int kw,
                                                               float* __restrict__ R_in) {
  const int Ho = H / kh, Wo = W / kw, groups = Cp / 4;
  const int64_t total = B * Ho * Wo * groups;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int gq = (int)(i % groups);
    const int64_t p = i / groups;
    const int xo = (int)(p % Wo), yo = (int)((p / Wo) % Ho);
    const int64_t n = p / ((int64_t)Wo * Ho);
    const float4 r = __ldg(reinterpret_cast<const float4*>(R_out + p * Cp + gq * 4));
    const uchar4 am = *reinterpret_cast<const uchar4*>(amax + p * Cp + gq * 4);
    const float rv[4] = {r.x, r.y, r.z, r.w};
    const unsigned char av[4] = {am.x, am.y, am.z, am.w};
    for (int dy = 0; dy < kh; ++dy)
      for (int dx = 0; dx < kw; ++dx) {
        const int widx = dy * kw + dx;
        float4 o;
        o.x = av[0] == widx ? rv[0] : 0.f; o.y = av[1] == widx ? rv[1] : 0.f;
        o.z = av[2] == widx ? rv[2] : 0.f; o.w = av[3] == widx ? rv[3] : 0.f;
        *reinterpret_cast<float4*>(R_in + (((n * H + yo * kh + dy) * W) + xo * kw + dx) * Cp + gq * 4) = o;
      }
  }
}

// NHWC fp32 [B,H,W,Cp] <-> NCHW fp32 [B,C,H,W] (32x32 shared-memory transposes; padded channels are zero-filled)
__global__ void __launch_bounds__(256) nhwc_f32_to_nchw_kernel(const float* __restrict__ x, int HW, int Cp, int C,
                                                               float* __restrict__ y) {
  __shared__ float t[32][33];
  const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int p = p0 + ty + 8 * r, c = c0 + tx;
    t[ty + 8 * r][tx] = (p < HW && c < Cp) ? x[((int64_t)n * HW + p) * Cp + c] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = c0 + ty + 8 * r, p = p0 + tx;
    if (c < C && p < HW) y[((int64_t)n * C + c) * HW + p] = t[tx][ty + 8 * r];
  }
}
__global__ void __launch_bounds__(256) nchw_to_nhwc_f32_kernel(const float* __restrict__ x, int HW, int C, int Cp,
                                                               float* __restrict__ y) {
  __shared__ float t[32][33];
  const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = c0 + ty + 8 * r, p = p0 + tx;
    t[ty + 8 * r][tx] = (c < C && p < HW) ? x[((int64_t)n * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int p = p0 + ty + 8 * r, c = c0 + tx;
    if (p < HW && c < Cp) y[((int64_t)n * HW + p) * Cp + c] = t[tx][ty + 8 * r];
  }
}

// NHWC hi/lo -> NCHW fp32 (C real channels); 32x32 shared-memory transpose over (pixel, channel).
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const __half* __restrict__ x_hi, const __half* __restrict__ x_lo,
                                                           int HW, int Cp, int C, float* __restrict__ y) {
  __shared__ float t[32][33];
  const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int p = p0 + ty + 8 * r, c = c0 + tx;
    float v = 0.f;
    if (p < HW && c < Cp) {
      const int64_t o = ((int64_t)n * HW + p) * Cp + c;
      v = __half2float(x_hi[o]) + __half2float(x_lo[o]);
    }
    t[ty + 8 * r][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = c0 + ty + 8 * r, p = p0 + tx;
    if (c < C && p < HW) y[((int64_t)n * C + c) * HW + p] = t[tx][ty + 8 * r];
  }
}

// R = 0 where the ReLU output a = hi + lo is not positive (autograd of an un-hooked ReLU, NHWC)
__global__ void relu_mask_nhwc_kernel(float* __restrict__ R, const __half* __restrict__ a_hi, const __half* __restrict__ a_lo,
                                      int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (!(__half2float(a_hi[i]) + __half2float(a_lo[i]) > 0.f)) R[i] = 0.f;
}

__global__ void split_f16_kernel(const float* __restrict__ in, int64_t n, __half* __restrict__ hi, __half* __restrict__ lo) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = in[i];
    const __half h = __float2half_rn(v);
    hi[i] = h;
    lo[i] = __float2half_rn(v - __half2float(h));
  }
}

// out[n] = max_i |R[n,i]| / x[n,i] over x > 0: the bound on |s| for the first Gamma layer below the dense head
// (R = x * c there, so this is max |c|).  One CTA per sample.
__global__ void __launch_bounds__(256) sample_absmax_ratio_kernel(const float* __restrict__ R, const float* __restrict__ x,
                                                                  int64_t per, float* __restrict__ out) {
  __shared__ float red[8];
  const int64_t base = (int64_t)blockIdx.x * per;
  float m = 0.f;
  for (int64_t i = threadIdx.x; i < per; i += blockDim.x) {
    const float xv = x[base + i];
    if (xv > 0.f) m = fmaxf(m, fabsf(R[base + i]) / xv);
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    m = warp_max(m);
    if (threadIdx.x == 0) out[blockIdx.x] = m;
  }
}

inline int eblocks(int64_t n) {
  int64_t b = (n + 255) / 256;
  return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b));
}

bool pick_tile(int B, int H, int W, int* nb, int* th, int* tw) {
  (void)B;
  // tw = largest power of two <= min(W, 128) that divides W; th, nb fill up to 128 pixels
  int w = 1;
  while (w * 2 <= W && w * 2 <= 128 && W % (w * 2) == 0) w *= 2;
  if (W % w != 0) return false;
  int rest = 128 / w, h = 1;
  while (h * 2 <= rest && h * 2 <= H && H % (h * 2) == 0) h *= 2;
  if (H % h != 0) return false;
  int n = rest / h;
  if (w * h * n != 128 || n > 256 || h > 256 || w > 256) return false;
  *nb = n; *th = h; *tw = w;
  return true;
}
// Tile for a conv with the max-pool fused into its epilogue: at most 16 pixels per tile row so that a warp (32 TMEM
// lanes) holds at least two rows of the tile.
bool pool_tile(int H, int W, int* nb, int* th, int* tw) {
  if (W % 16 == 0 && H % 8 == 0) { *nb = 1; *th = 8; *tw = 16; return true; }
  if (W == 8 && H == 8) { *nb = 2; *th = 8; *tw = 8; return true; }
  return false;
}
}  // namespace

bool conv_tc_pool_supported(int64_t B, int Cin_p, int Cout_p, int H, int W, int kh, int kw);

bool conv_tc_supported(int64_t B, int Cin_p, int Cout_p, int H, int W) {
  int nb, th, tw;
  if (B <= 0 || Cin_p % 64 != 0 || Cout_p % 64 != 0 || Cout_p > 256 || Cin_p > 1024) return false;
  return pick_tile((int)B, H, W, &nb, &th, &tw);
}

bool conv_tc_pool_supported(int64_t B, int Cin_p, int Cout_p, int H, int W, int kh, int kw) {
  int nb, th, tw;
  if (!conv_tc_supported(B, Cin_p, Cout_p, H, W) || !pool_tile(H, W, &nb, &th, &tw)) return false;
  const bool pow2 = kh > 0 && kw > 0 && (kh & (kh - 1)) == 0 && (kw & (kw - 1)) == 0;
  return pow2 && kw <= tw && kh <= 32 / tw && kh * kw <= 255 && H % kh == 0 && W % kw == 0;
}

// Experiment switch (lrp_debug_set_conv_variant): bit e set = the conv launches with epilogue e skip the x_lo * w_hi product.
int g_conv_variant = 0;
void set_conv_variant(int v) { g_conv_variant = v; }

int conv_tc_run(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias, int64_t B,
                int H, int W, int Cin_p, int Cout_p, int Cout, int relu, int epi, float eps, const float* aux_f32,
                const void* aux_hi, const void* aux_lo, void* y_hi, void* y_lo, float* y_f32, float* y_nchw,
                const float* scale_ref, float* cmax_out, int* err_flag, cudaStream_t stream, int pkh, int pkw,
                void* amax_out) {
  ConvGeom g{};
  if (!conv_tc_supported(B, Cin_p, Cout_p, H, W) || B > 2147483647LL / ((int64_t)H * W)) return DRSA_ERR_SHAPE;
  g.B = (int)B; g.H = H; g.W = W; g.Cin_p = Cin_p; g.Cout_p = Cout_p; g.Cout = Cout; g.relu = relu; g.epi = epi; g.eps = eps;
  pick_tile(g.B, H, W, &g.nb, &g.th, &g.tw);
  const bool halo = Cout_p == 64 && Cin_p == 64 && W % kHaloTw == 0 && H % kHaloTh == 0;
  if (halo) { g.nb = 1; g.th = kHaloTh; g.tw = kHaloTw; }
  g.pkh = g.pkw = 0;
  g.drop_xlo = (g_conv_variant & (1 << epi)) ? 1 : 0;      // bit 0: forward, bit 1: ratio, bit 2: input-multiply
  if (pkh > 0 || pkw > 0) {
    if (epi != 0 || y_f32 != nullptr || y_nchw != nullptr || y_hi == nullptr || !conv_tc_pool_supported(B, Cin_p, Cout_p, H, W, pkh, pkw))
      return DRSA_ERR_SHAPE;
    pool_tile(H, W, &g.nb, &g.th, &g.tw);
    g.pkh = pkh; g.pkw = pkw;
  }
  g.tiles_x = W / g.tw; g.tiles_y = H / g.th; g.tiles_b = (g.B + g.nb - 1) / g.nb;
  g.num_tiles = g.tiles_x * g.tiles_y * g.tiles_b;
  CUtensorMap tmXh, tmXl, tmWh, tmWl;
  DRSA_TRY(make_tmap_nhwc(&tmXh, x_hi, (uint64_t)B, H, W, Cin_p, g.nb, halo ? g.th + 2 : g.th, g.tw));
  DRSA_TRY(make_tmap_nhwc(&tmXl, x_lo, (uint64_t)B, H, W, Cin_p, g.nb, halo ? g.th + 2 : g.th, g.tw));
  DRSA_TRY(make_tmap_f16_sw128(&tmWh, w_hi, (uint64_t)9 * Cout_p, (uint64_t)Cin_p, (uint32_t)Cout_p));
  DRSA_TRY(make_tmap_f16_sw128(&tmWl, w_lo, (uint64_t)9 * Cout_p, (uint64_t)Cin_p, (uint32_t)Cout_p));
  int grid = sm_count();
  if (grid > g.num_tiles) grid = g.num_tiles;
  auto launch = [&](auto kernel, int stages, bool resw) -> int {
    const int stage_bytes = halo ? 2 * kHaloBytes : (resw ? 2 * kABytes : 2 * kABytes + 2 * Cout_p * 128);
    const int smem_bytes = stages * stage_bytes + (resw ? 18 * Cout_p * 128 : 0) + 256;
    DRSA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    kernel<<<grid, kConvThreads, smem_bytes, stream>>>(tmXh, tmXl, tmWh, tmWl, g, bias, aux_f32,
                                                        static_cast<const __half*>(aux_hi), static_cast<const __half*>(aux_lo),
                                                        static_cast<__half*>(y_hi), static_cast<__half*>(y_lo), y_f32, y_nchw,
                                                        scale_ref, cmax_out, err_flag, static_cast<uint8_t*>(amax_out));
    DRSA_LAUNCH_CHECK();
    return DRSA_OK;
  };
  if (halo) return launch(conv3x3_tc_kernel<2, true, true>, 2, true);
  if (Cout_p == 64 && Cin_p == 64) return launch(conv3x3_tc_kernel<2, true, false>, 2, true);
  if (Cout_p <= 64) return launch(conv3x3_tc_kernel<4, false, false>, 4, false);
  if (Cout_p <= 128) return launch(conv3x3_tc_kernel<3, false, false>, 3, false);
  return launch(conv3x3_tc_kernel<2, false, false>, 2, false);
}

int conv_tc_forward(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias, int64_t B,
                    int H, int W, int Cin_p, int Cout_p, int Cout, int relu, void* y_hi, void* y_lo, float* y_nchw,
                    int* err_flag, cudaStream_t stream) {
  return conv_tc_run(x_hi, x_lo, w_hi, w_lo, bias, B, H, W, Cin_p, Cout_p, Cout, relu, 0, 0.f, nullptr, nullptr, nullptr, y_hi,
                     y_lo, nullptr, y_nchw, nullptr, nullptr, err_flag, stream, 0, 0, nullptr);
}

// conv + bias (+ ReLU) + MaxPool2d(kh, kw): y planes [B, H/kh, W/kw, Cout_p], argmax_u8 (optional) the window index
int conv_tc_forward_pool(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias, int64_t B,
                         int H, int W, int Cin_p, int Cout_p, int Cout, int relu, int kh, int kw, void* y_hi, void* y_lo,
                         void* argmax_u8, int* err_flag, cudaStream_t stream) {
  return conv_tc_run(x_hi, x_lo, w_hi, w_lo, bias, B, H, W, Cin_p, Cout_p, Cout, relu, 0, 0.f, nullptr, nullptr, nullptr, y_hi,
                     y_lo, nullptr, nullptr, nullptr, nullptr, err_flag, stream, kh, kw, argmax_u8);
}

int conv_first_nhwc(const float* x, const float* w, const float* b, int64_t B, int H, int W, int Cout, int Cout_p,
                    int relu, void* y_hi, void* y_lo, cudaStream_t stream) {
  const int groups = Cout_p / 8;
  if (groups > 256 || 256 % groups != 0) return DRSA_ERR_SHAPE;
  const int64_t rows = B * H;
  conv3x3_first_nhwc_kernel<<<(int)(rows < 148 * 8 ? rows : 148 * 8), 256, 0, stream>>>(
      x, w, b, B, H, W, Cout, Cout_p, relu, static_cast<__half*>(y_hi), static_cast<__half*>(y_lo));
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

int first_layer_ones_backward(const float* R_out, const float* w_mod, const float* b_mod, int64_t B, int H, int W, int Cout,
                              int Cp, float eps, float* R_in, cudaStream_t stream) {
  if (Cp != 64 || Cout > 64) return DRSA_ERR_SHAPE;
  const int64_t tiles = B * ((H + kFbTh - 1) / kFbTh) * ((W + kFbTw - 1) / kFbTw);
  if (tiles > 2147483647LL) return DRSA_ERR_SHAPE;
  first_layer_ones_bwd_kernel<<<(unsigned)tiles, 512, 0, stream>>>(R_out, w_mod, b_mod, H, W, Cout, eps, R_in);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

int maxpool_nhwc(const void* x_hi, const void* x_lo, int64_t B, int H, int W, int Cp, int kh, int kw, void* y_hi,
                 void* y_lo, void* argmax_u8, cudaStream_t stream) {
  maxpool_nhwc_kernel<<<eblocks(B * (H / kh) * (W / kw) * (Cp / 8)), 256, 0, stream>>>(
      static_cast<const __half*>(x_hi), static_cast<const __half*>(x_lo), B, H, W, Cp, kh, kw,
      static_cast<__half*>(y_hi), static_cast<__half*>(y_lo), static_cast<uint8_t*>(argmax_u8));
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

int maxpool_nhwc_backward(const float* R_out, const void* argmax_u8, int64_t B, int H, int W, int Cp, int kh, int kw,
                          float* R_in, cudaStream_t stream) {
  maxpool_nhwc_bwd_kernel<<<eblocks(B * (H / kh) * (W / kw) * (Cp / 4)), 256, 0, stream>>>(
      R_out, static_cast<const uint8_t*>(argmax_u8), B, H, W, Cp, kh, kw, R_in);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

int nhwc_f32_to_nchw(const float* x, int64_t B, int H, int W, int Cp, int C, float* y, cudaStream_t stream) {
  if (B > 65535) return DRSA_ERR_SHAPE;
  dim3 grid(cdiv((int64_t)H * W, 32), cdiv(C, 32), (unsigned)B);
  nhwc_f32_to_nchw_kernel<<<grid, 256, 0, stream>>>(x, H * W, Cp, C, y);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

int nchw_to_nhwc_f32(const float* x, int64_t B, int H, int W, int C, int Cp, float* y, cudaStream_t stream) {
  if (B > 65535) return DRSA_ERR_SHAPE;
  dim3 grid(cdiv((int64_t)H * W, 32), cdiv(Cp, 32), (unsigned)B);
  nchw_to_nhwc_f32_kernel<<<grid, 256, 0, stream>>>(x, H * W, C, Cp, y);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

int nhwc_to_nchw(const void* x_hi, const void* x_lo, int64_t B, int H, int W, int Cp, int C, float* y,
                 cudaStream_t stream) {
  if (B > 65535) return DRSA_ERR_SHAPE;
  dim3 grid(cdiv((int64_t)H * W, 32), cdiv(C, 32), (unsigned)B);
  nhwc_to_nchw_kernel<<<grid, 256, 0, stream>>>(static_cast<const __half*>(x_hi), static_cast<const __half*>(x_lo),
                                                H * W, Cp, C, y);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

int relu_mask_nhwc(float* R, const void* a_hi, const void* a_lo, int64_t count, cudaStream_t stream) {
  relu_mask_nhwc_kernel<<<eblocks(count), 256, 0, stream>>>(R, static_cast<const __half*>(a_hi),
                                                            static_cast<const __half*>(a_lo), count);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

int sample_absmax_ratio(const float* R, const float* x, int64_t B, int64_t per, float* out, cudaStream_t stream) {
  if (B > 2147483647LL) return DRSA_ERR_SHAPE;
  sample_absmax_ratio_kernel<<<(unsigned)B, 256, 0, stream>>>(R, x, per, out);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

int split_f16(const float* in, int64_t count, void* hi, void* lo, cudaStream_t stream) {
  split_f16_kernel<<<eblocks(count), 256, 0, stream>>>(in, count, static_cast<__half*>(hi), static_cast<__half*>(lo));
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

}  // namespace drsa
