// DRSA row pass on the 5th-generation tensor cores (DRSA_PREC_TC_F16 / DRSA_PREC_TC_F16X2).
//
// One kernel does, per subtile of 64 (activation, context) rows and per group of 128 projected
// columns (= 128/d_k whole concepts), everything drsa.py:148-155 and the backward of drsa.py:100
// do, without the projected activations ever leaving the SM:
//
//   GEMM1 (tcgen05.mma SS, fp16 x fp16 -> fp32 in TMEM), transposed so that a TMEM lane is a
//          projected column j and a TMEM column is a row r of the subtile; the C rows are stacked
//          under the A rows, so one MMA of N = 128 gives both projections:
//              [HA^T | HC^T][j][r] = sum_i U^T[j][i] [A; C][r][i]      (U^T = fp16(U), or hi + lo: two MMAs)
//   epilogue (8 warps per subtile, tcgen05.ld): s_rk = sum_{j in k} HA^T[j][r] HC^T[j][r] is a reduction
//          across TMEM lanes -> warp transpose-reduce (31 shuffles per 32 rows) + a 2 KB smem exchange;
//          g = relu(s) ; sumsq_k += g^2 ; P^T = g * HC^T and Q^T = g * HA^T are written back over
//          HA^T / HC^T in TMEM as packed fp16 (tcgen05.st) -- the A operand of GEMM2.
//   GEMM2 (tcgen05.mma TS, A operand from TMEM, B operand = the same rows read MN-major, N = D):
//              X^T[j][i] += sum_r P^T[j][r] A[r][i] + Q^T[j][r] C[r][i]
//          accumulated in TMEM over ALL subtiles of the CTA and written out once.
//
// TMEM: X^T 256 columns + two H buffers of 128 columns, so the epilogue of subtile i runs under GEMM1 of
// subtile i+1 (issue order G1(0) G1(1) G2(0) G1(2) G2(1) ...).  A and C are stored once as scaled fp16
// (drsa_pack_f16); U is rounded to fp16 once per step (or split hi + lo).
// Operand staging: TMA (cp.async.bulk.tensor, 128-byte swizzle) into an mbarrier ring; U^T of the CTA's
// column group stays resident in shared memory.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread) + TMEM owner,
// warps 2..17 = epilogue (TMEM lane quarter = warp % 4; warps 2..9 take the even subtiles, 10..17 the odd ones).
// d = 512 splits X^T over two CTAs per column group.
#include <cuda.h>
#include <vector>
#include <cmath>
#include <cstdlib>
#include "common.cuh"
#include "tc_common.cuh"

namespace drsa {

// ----------------------------------------------------------------------------- tensor maps
namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
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
}  // namespace

int make_tmap_f16_sw128(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return DRSA_ERR_CUDA;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { g_last_cuda_error = (int)r; return DRSA_ERR_CUDA; }
  return DRSA_OK;
}

namespace {
using namespace tc;

constexpr int kSub = 64;               // rows of A / C per subtile (half the MMA N of GEMM1, the K of GEMM2)
constexpr int kNG = 128;               // projected columns per CTA (MMA M)
constexpr int kPanelBytes = 128 * 128; // one [128 x 64] fp16 box, 128-byte rows
constexpr int kThreads = 576;          // TMA warp, MMA warp, 2 sets of 8 epilogue warps (even / odd subtiles)
constexpr int kRedBytes = 2 * 4 * kSub * 4;
constexpr int kBarBytes = 256;

template <int D, bool kSplitU>
struct Cfg {
  static constexpr int kPanels = D / 64;                      // 64-channel panels of the GEMM1 contraction (K = D)
  // channels of X^T a CTA accumulates (GEMM2 N): at most 256 TMEM columns.  d = 512 is split over kHalves CTAs per
  // column group, each of which repeats GEMM1 for the full K = 512 (executed MMA work 1.5x the algorithmic).
  static constexpr int kDN = D > 256 ? 256 : D;
  static constexpr int kHalves = D / kDN;
  static constexpr int kNPanels = kDN / 64;
  static constexpr int kUBytes = (kSplitU ? 2 : 1) * kPanels * kPanelBytes;   // U^T (hi, or hi + lo) of this column group
  static constexpr int kStageBytes = 2 * kPanelBytes;         // GEMM1: two [A64;C64] panels, GEMM2: 32 rows of A and of C
  static constexpr int kStages = (D >= 512) ? 3 : (D >= 256) ? (kSplitU ? 3 : 5) : (kSplitU ? 5 : 6);
  static constexpr int kDataBytes = kUBytes + kStages * kStageBytes;
  static constexpr int kSmemBytes = kDataBytes + kRedBytes + kBarBytes;
};

// tcgen05 shared-memory descriptors split into a constant high word and a cheap low word
// (see tc::make_smem_desc_sw128): hi = SBO 1024 B | version 1 | SWIZZLE_128B.
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t kdesc_lo(uint32_t smem_addr) { return (smem_addr >> 4) | (1u << 16); }           // LBO 16 B
__device__ __forceinline__ uint32_t mndesc_lo(uint32_t smem_addr) { return (smem_addr >> 4) | ((4096u >> 4) << 16); }            // LBO 4 KB: next 64-channel panel
__device__ __forceinline__ uint64_t desc64(uint32_t lo) { return ((uint64_t)kDescHi << 32) | lo; }

__device__ __forceinline__ uint32_t pack_h2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));   // d = {hi : upper, lo : lower}
  return r;
}

// Schedule (per CTA, subtiles i = 0, 1, ... of 64 rows; H has two TMEM buffers of 128 columns):
//     MMA thread :  G1(0) G1(1) G2(0) G1(2) G2(1) G1(3) G2(2) ...          (tensor pipe executes in issue order)
//     epilogue   :        E(0)        E(1)        E(2)   ...               E(i) starts when G1(i) completes
// so the epilogue of subtile i runs under GEMM1 of subtile i+1 and the tensor pipe never waits for it as long as
// E(i) is shorter than G1(i+1) + G2(i-1).  H[i & 1] is free for G1(i) because G2(i-2), issued earlier, has read it.
template <int D, bool kSplitU>
__global__ void __launch_bounds__(kThreads, 1)
drsa_tc_step_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmC,
                    const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmC2,
                    const __grid_constant__ CUtensorMap tmUh, const __grid_constant__ CUtensorMap tmUl,
                    int n_sub, int G, int nRB, int d_k, float inv_scale, float pq_scale, float* __restrict__ part,
                    float* __restrict__ ss_part, int* __restrict__ err_flag, long long* __restrict__ prof) {
  using C = Cfg<D, kSplitU>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sU_hi = smem;
  uint8_t* sU_lo = smem + C::kPanels * kPanelBytes;
  uint8_t* sStage = smem + C::kUBytes;
  float* red = reinterpret_cast<float*>(smem + C::kDataBytes);      // [2 buffers][4 lane quarters][64 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kDataBytes + kRedBytes);
  uint64_t* full = bars;            // [kStages]
  uint64_t* empty = bars + 8;       // [kStages]
  uint64_t* u_full = bars + 16;
  uint64_t* x_full = bars + 17;
  uint64_t* h_full = bars + 18;     // [2]: H buffer written by GEMM1
  uint64_t* p_full = bars + 20;     // [2 buffers][2 chunks]: P^T / Q^T of a 32-row chunk written by its 4 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x % G, ih = (blockIdx.x / G) % C::kHalves, rb = blockIdx.x / (G * C::kHalves);

  if ((smem_u32(smem) & 1023u) != 0) {   // 128-byte swizzle needs 1024-byte aligned panels
    if (threadIdx.x == 0) atomicExch(err_flag, 1);
    return;
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(u_full, 1); mbar_init(x_full, 1); mbar_init(&h_full[0], 1); mbar_init(&h_full[1], 1);
    for (int c = 0; c < 4; ++c) mbar_init(&p_full[c], 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmC); tma_prefetch_desc(&tmUh); tma_prefetch_desc(&tmUl);
      mbar_expect_tx(u_full, C::kUBytes);
      for (int p = 0; p < C::kPanels; ++p) {
        tma_load_2d(sU_hi + p * kPanelBytes, &tmUh, u_full, 64 * p, g * kNG);
        if (kSplitU) tma_load_2d(sU_lo + p * kPanelBytes, &tmUl, u_full, 64 * p, g * kNG);
      }
      tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmC2);
      int stage = 0; uint32_t phase = 0;
      // GEMM1 operand: per 64-channel panel the stacked tile [A rows r0..r0+63 ; C rows r0..r0+63] (K-major), two panels per stage
      auto load_g1 = [&](int sub) {
        for (int hs = 0; hs < C::kPanels / 2; ++hs) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], C::kStageBytes);
          uint8_t* dst = sStage + stage * C::kStageBytes;
#pragma unroll
          for (int pp = 0; pp < 2; ++pp) {
            tma_load_2d(dst + pp * kPanelBytes, &tmA, &full[stage], 64 * (2 * hs + pp), sub * kSub);
            tma_load_2d(dst + pp * kPanelBytes + kPanelBytes / 2, &tmC, &full[stage], 64 * (2 * hs + pp), sub * kSub);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      };
      // GEMM2 operand: row chunks [32 rows x D ch] of A and of C as D/64 panels of 4 KB each (MN-major, N = D)
      auto load_g2 = [&](int sub) {
        for (int c = 0; c < 2; ++c) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], 2 * C::kNPanels * 4096);
          uint8_t* dst = sStage + stage * C::kStageBytes;
          for (int p = 0; p < C::kNPanels; ++p) {
            tma_load_2d(dst + p * 4096, &tmA2, &full[stage], C::kDN * ih + 64 * p, sub * kSub + 32 * c);
            tma_load_2d(dst + kPanelBytes + p * 4096, &tmC2, &full[stage], C::kDN * ih + 64 * p, sub * kSub + 32 * c);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      };
      int prev = -1;
      for (int sub = rb; sub < n_sub; sub += nRB) {
        load_g1(sub);
        if (prev >= 0) load_g2(prev);
        prev = sub;
      }
      if (prev >= 0) load_g2(prev);
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc_f16(kNG, 2 * kSub, 0, 0);  // U^T (K-major) x stacked [A;C] subtile (K-major)
      constexpr uint32_t idesc2 = make_idesc_f16(kNG, C::kDN, 0, 1);    // P^T (TMEM)   x row chunk (MN-major, N = kDN)
      const uint32_t tX = tmem_base;
      mbar_wait(u_full, 0);
      tc_fence_after();
      int stage = 0; uint32_t phase = 0; bool first = true;
      long long pa = 0, pb = 0, pc = 0;
      // ---- GEMM1(i): [HA^T | HC^T] (128 columns of U x (64 + 64) rows) -> H[i & 1], K = D
      auto gemm1 = [&](int i) {
        const long long t0 = prof ? clock64() : 0;
        const uint32_t tH = tmem_base + 256 + 128 * (i & 1);
        for (int hs = 0; hs < C::kPanels / 2; ++hs) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t base = smem_u32(sStage + stage * C::kStageBytes);
#pragma unroll
          for (int pp = 0; pp < 2; ++pp) {
            const int p = 2 * hs + pp;
            // descriptors: the high word is constant, the low word is (address >> 4) | LBO field; advancing
            // along K inside the 128-byte swizzle atom adds 32 B (>> 4 = 2) to the low word
            const uint32_t uh = kdesc_lo(smem_u32(sU_hi + p * kPanelBytes)), ul = kdesc_lo(smem_u32(sU_lo + p * kPanelBytes));
            const uint32_t dAC = kdesc_lo(base + pp * kPanelBytes);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t d_ac = desc64(dAC + 2 * kk);
              umma_ss_f16(tH, desc64(uh + 2 * kk), d_ac, idesc1, (p | kk) ? 1u : 0u);
              if (kSplitU) umma_ss_f16(tH, desc64(ul + 2 * kk), d_ac, idesc1, 1u);
            }
          }
          umma_commit(&empty[stage]);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&h_full[i & 1]);
        if (prof) pa += clock64() - t0;
      };
      // ---- GEMM2(i): X^T[:, 0..D) += P^T rows_A + Q^T rows_C, K = 64 rows in 2 chunks of 32.  One MMA covers all D
      //      channels (N = D) so the TMEM-resident operand P^T is read once per k-step.
      auto gemm2 = [&](int i) {
        const long long t0 = prof ? clock64() : 0;
        long long t1 = t0;
        const uint32_t tH = tmem_base + 256 + 128 * (i & 1);
        const uint32_t par = (uint32_t)(i >> 1) & 1u;
        for (int c = 0; c < 2; ++c) {
          mbar_wait(&p_full[2 * (i & 1) + c], par);
          if (prof && c == 0) t1 = clock64();
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t base = smem_u32(sStage + stage * C::kStageBytes);
          const uint32_t dA = mndesc_lo(base), dC = mndesc_lo(base + kPanelBytes);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            // 16 rows of K = two 8-row swizzle atoms = 2048 B (>> 4 = 128); P^T k-step = 8 TMEM columns
            const uint32_t off = 32 * c + 8 * h;
            umma_ts_f16(tX, tH + off, desc64(dA + 128 * h), idesc2, first ? 0u : 1u);
            umma_ts_f16(tX, tH + kSub + off, desc64(dC + 128 * h), idesc2, 1u);
            first = false;
          }
          umma_commit(&empty[stage]);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        if (prof) { pb += t1 - t0; pc += clock64() - t1; }
      };
      int i = 0;
      for (int sub = rb; sub < n_sub; sub += nRB, ++i) {
        gemm1(i);
        if (i > 0) gemm2(i - 1);
      }
      if (i > 0) gemm2(i - 1);
      umma_commit(x_full);
      if (prof && blockIdx.x == 0) { prof[0] = pa; prof[1] = pb; prof[2] = pc; }
    }
  } else {
    // ======================= epilogue warps =======================
    const int q = warp & 3;                     // TMEM lane quarter this warp may touch
    const int set = (warp - 2) >> 3;            // warps 2..9 take the even subtiles (H buffer 0), warps 10..17 the odd ones
    const int c = ((warp - 2) >> 2) & 1;        // the 32-row chunk of those subtiles this warp works on
    const int j = 32 * q + lane;                // projected column within the group
    const int wpc = d_k >> 5;                   // warps per concept (d_k in {32, 64, 128})
    const int q0 = (q / wpc) * wpc;
    const uint32_t lane_base = tmem_base + ((uint32_t)(32 * q) << 16);
    const bool owner = (j % d_k) == 0;           // one thread per (concept, chunk) accumulates sum g^2
    float ssq = 0.f;
    long long ea = 0, eb = 0, e0 = 0, e1 = 0;
    for (int i = set, sub = rb + set * nRB; sub < n_sub; sub += 2 * nRB, i += 2) {
      const int buf = set;
      if (prof) e0 = clock64();
      mbar_wait(&h_full[buf], (uint32_t)(i >> 1) & 1u);
      tc_fence_after();
      if (prof) e1 = clock64();
      const uint32_t tHA = lane_base + 256 + 128 * buf + 32 * c, tHC = tHA + kSub;
      float* redb = red + buf * (4 * kSub);
      uint32_t ha[32], hc[32];
      tmem_ld32(tHA, ha);
      tmem_ld32(tHC, hc);
      tmem_ld_wait();
      float pr[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) pr[e] = __uint_as_float(ha[e]) * __uint_as_float(hc[e]);
      // transpose-reduce: afterwards lane l holds sum over the warp's 32 lanes of column l
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int e = 0; e < off; ++e) {
          const float send = upper ? pr[e] : pr[e + off];
          const float keep = upper ? pr[e + off] : pr[e];
          pr[e] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      redb[q * kSub + 32 * c + lane] = pr[0];
      asm volatile("bar.sync %0, 128;" ::"r"(1 + 2 * set + c) : "memory");   // the 4 warps working on this chunk
      uint32_t pk[16], qk[16];
#pragma unroll
      for (int i4 = 0; i4 < 8; ++i4) {
        float4 s4 = *reinterpret_cast<const float4*>(&redb[q0 * kSub + 32 * c + 4 * i4]);
#pragma unroll
        for (int w = 1; w < 4; ++w) {
          if (w < wpc) {
            const float4 o = *reinterpret_cast<const float4*>(&redb[(q0 + w) * kSub + 32 * c + 4 * i4]);
            s4.x += o.x; s4.y += o.y; s4.z += o.z; s4.w += o.w;
          }
        }
        // s is in packed scale (sA*sC*s_true): sums of squares are taken in true scale, the fp16
        // operands P, Q of GEMM2 in packed scale times pq_scale (overflow-safe by construction)
        float g0 = fmaxf(s4.x, 0.f), g1 = fmaxf(s4.y, 0.f), g2 = fmaxf(s4.z, 0.f), g3 = fmaxf(s4.w, 0.f);
        if (owner) {
          const float t0 = g0 * inv_scale, t1 = g1 * inv_scale, t2 = g2 * inv_scale, t3 = g3 * inv_scale;
          ssq += t0 * t0 + t1 * t1 + t2 * t2 + t3 * t3;
        }
        g0 *= pq_scale; g1 *= pq_scale; g2 *= pq_scale; g3 *= pq_scale;
        const int e = 4 * i4;
        pk[2 * i4] = pack_h2_sat(g0 * __uint_as_float(hc[e]), g1 * __uint_as_float(hc[e + 1]));
        pk[2 * i4 + 1] = pack_h2_sat(g2 * __uint_as_float(hc[e + 2]), g3 * __uint_as_float(hc[e + 3]));
        qk[2 * i4] = pack_h2_sat(g0 * __uint_as_float(ha[e]), g1 * __uint_as_float(ha[e + 1]));
        qk[2 * i4 + 1] = pack_h2_sat(g2 * __uint_as_float(ha[e + 2]), g3 * __uint_as_float(ha[e + 3]));
      }
      tmem_st16(tHA, pk);   // P^T = g * HC^T (pairs with A rows)
      tmem_st16(tHC, qk);   // Q^T = g * HA^T (pairs with C rows)
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[2 * buf + c]);
      if (prof) { ea += e1 - e0; eb += clock64() - e1; }
    }
    if (prof && blockIdx.x == 0 && warp == 2 && lane == 0) { prof[3] = ea; prof[4] = eb; }
    // ---- final: X^T of this CTA -> partial buffer [cta][j][D]
    mbar_wait(x_full, 0);
    tc_fence_after();
    float* dst = part + (((int64_t)rb * G + g) * kNG + j) * D + C::kDN * ih;
#pragma unroll 1
    for (int cc = (warp - 2) >> 2; cc < C::kDN / 32; cc += 4) {
      uint32_t v[32];
      tmem_ld32(lane_base + 32 * cc, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 8; ++e)
        *reinterpret_cast<float4*>(dst + 32 * cc + 4 * e) =
            make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]),
                        __uint_as_float(v[4 * e + 3]));
    }
    // one slot per (row block, concept, warp group): summed in a fixed order by tc_reduce_kernel (deterministic)
    if (owner && ih == 0) ss_part[((int64_t)rb * (G * (kNG / d_k)) + g * (kNG / d_k) + j / d_k) * 4 + 2 * set + c] = ssq;
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Single-pass variant for d <= 256 (DRSA_PREC_TC_F16): U^T lives in TENSOR MEMORY.
//
// What bounds the kernel above is shared-memory operand bandwidth, not the tensor pipe: an SS-mode MMA reads both
// operands from shared memory (GEMM1 at M = N = 128: 8 KB per 64-cycle MMA = the full 128 B/clk, ~80 B/clk are
// sustained next to the TMA writes), and every row is staged twice (GEMM1 and GEMM2 views).  Here
//   * U^T (fp16, 128 x D) is written once into 128 x D/2 TMEM columns and GEMM1 runs in TS mode (A from TMEM): per
//     MMA only the 2 KB row operand comes from shared memory;
//   * subtiles have 32 rows, so two H buffers need 2 x 64 columns: X^T 256 + U^T 128 + H 128 = 512 columns;
//   * a stage holds, per 64-channel panel, [A rows | C rows] of the subtile (K-major operand of GEMM1 with N = 64)
//     and GEMM2 reads THE SAME bytes as its MN-major operand (N-block stride 8 KB): every row is staged once.
// Per 32-row subtile: GEMM1 16 MMAs (M128 N64 K16), GEMM2 4 MMAs (M128 N256 K16, 128 cycles) for 64 KB of operand reads
// and 32 KB of TMA writes.  Schedule, barriers and epilogue arithmetic as above.  EXPERIMENTAL: see g_tc_variant below
// for the measured outcome (the N = 64 MMAs do not run at 32 cycles).
constexpr int kSub32 = 32;
__device__ __forceinline__ uint32_t mndesc8k_lo(uint32_t smem_addr) { return (smem_addr >> 4) | ((8192u >> 4) << 16); }   // LBO 8 KB

template <int D>
struct Cfg32 {
  static constexpr int kPanels = D / 64;
  static constexpr int kStageBytes = kPanels * 8192;              // per panel: A rows (4 KB) | C rows (4 KB)
  static constexpr int kStages = (D >= 256) ? 6 : 8;
  static constexpr int kRedBytes = 2 * 4 * kSub32 * 4;
  static constexpr int kDataBytes = kStages * kStageBytes;
  static constexpr int kSmemBytes = kDataBytes + kRedBytes + kBarBytes;
  static constexpr int kUCols = D / 2;                            // TMEM columns of U^T (fp16 pairs)
  static constexpr uint32_t tX = 0, tU = D, tH = D + D / 2;       // column offsets: X^T [0,D) | U^T | H0 (64) | H1 (64)
};

template <int D>
__global__ void __launch_bounds__(kThreads, 1)
drsa_tc_step32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmC,
                      const __half* __restrict__ Ut_hi, int n_sub, int G, int nRB, int d_k, float inv_scale, float pq_scale,
                      float* __restrict__ part, float* __restrict__ ss_part, int* __restrict__ err_flag,
                      long long* __restrict__ prof) {
  using C = Cfg32<D>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sStage = smem;
  float* red = reinterpret_cast<float*>(smem + C::kDataBytes);      // [2 buffers][4 lane quarters][32 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kDataBytes + C::kRedBytes);
  uint64_t* full = bars;            // [kStages]
  uint64_t* empty = bars + 8;       // [kStages]
  uint64_t* u_full = bars + 16;
  uint64_t* x_full = bars + 17;
  uint64_t* h_full = bars + 18;     // [2]
  uint64_t* p_full = bars + 20;     // [2 buffers][2 half-chunks of 16 rows]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x % G, rb = blockIdx.x / G;

  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) atomicExch(err_flag, 1);
    return;
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(u_full, 512); mbar_init(x_full, 1); mbar_init(&h_full[0], 1); mbar_init(&h_full[1], 1);
    for (int c = 0; c < 4; ++c) mbar_init(&p_full[c], 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ======================= TMA producer: one stage per subtile =======================
    if (lane == 0) {
      tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmC);
      int stage = 0; uint32_t phase = 0;
      for (int sub = rb; sub < n_sub; sub += nRB) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], C::kStageBytes);
        uint8_t* dst = sStage + stage * C::kStageBytes;
#pragma unroll
        for (int p = 0; p < C::kPanels; ++p) {
          tma_load_2d(dst + p * 8192, &tmA, &full[stage], 64 * p, sub * kSub32);
          tma_load_2d(dst + p * 8192 + 4096, &tmC, &full[stage], 64 * p, sub * kSub32);
        }
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc_f16(kNG, 2 * kSub32, 0, 0);   // U^T (TMEM) x stacked [A;C] subtile (K-major, N = 64)
      constexpr uint32_t idesc2 = make_idesc_f16(kNG, D, 0, 1);            // P^T (TMEM) x the same rows (MN-major, N = D)
      const uint32_t tX = tmem_base + C::tX, tU = tmem_base + C::tU;
      mbar_wait(u_full, 0);                                                // U^T has been written to tensor memory
      tc_fence_after();
      bool first = true;
      long long pa = 0, pb = 0, pc = 0;
      auto gemm1 = [&](int i) {
        const long long t0 = prof ? clock64() : 0;
        const int stage = i % C::kStages;
        const uint32_t tH = tmem_base + C::tH + 64 * (i & 1);
        mbar_wait(&full[stage], (uint32_t)(i / C::kStages) & 1u);
        tc_fence_after();
        const uint32_t base = smem_u32(sStage + stage * C::kStageBytes);
#pragma unroll
        for (int p = 0; p < C::kPanels; ++p) {
          const uint32_t dAC = kdesc_lo(base + p * 8192);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ts_f16(tH, tU + 8 * (4 * p + kk), desc64(dAC + 2 * kk), idesc1, (p | kk) ? 1u : 0u);
        }
        umma_commit(&h_full[i & 1]);
        if (prof) pa += clock64() - t0;
      };
      auto gemm2 = [&](int i) {
        const long long t0 = prof ? clock64() : 0;
        long long t1 = t0;
        const int stage = i % C::kStages;
        const uint32_t tH = tmem_base + C::tH + 64 * (i & 1);
        const uint32_t par = (uint32_t)(i >> 1) & 1u;
        const uint32_t base = smem_u32(sStage + stage * C::kStageBytes);
        const uint32_t dA = mndesc8k_lo(base), dC = mndesc8k_lo(base + 4096);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait(&p_full[2 * (i & 1) + h], par);       // P^T / Q^T of rows 16h .. 16h+15
          if (prof && h == 0) t1 = clock64();
          tc_fence_after();
          umma_ts_f16(tX, tH + 16 * h, desc64(dA + 128 * h), idesc2, first ? 0u : 1u);
          umma_ts_f16(tX, tH + kSub32 + 16 * h, desc64(dC + 128 * h), idesc2, 1u);
          first = false;
        }
        umma_commit(&empty[stage]);                       // the stage served GEMM1 and GEMM2 of this subtile
        if (prof) { pb += t1 - t0; pc += clock64() - t1; }
      };
      int i = 0;
      for (int sub = rb; sub < n_sub; sub += nRB, ++i) {
        gemm1(i);
        if (i > 0) gemm2(i - 1);
      }
      if (i > 0) gemm2(i - 1);
      umma_commit(x_full);
      if (prof && blockIdx.x == 0) { prof[0] = pa; prof[1] = pb; prof[2] = pc; }
    }
  } else {
    // ======================= epilogue warps =======================
    const int q = warp & 3;                     // TMEM lane quarter this warp may touch
    const int e = warp - 2;                     // 0 .. 15
    const int set = e >> 3;                     // even / odd subtiles
    const int ch = (e >> 2) & 1;                // rows 16*ch .. 16*ch+15 of those subtiles
    const int j = 32 * q + lane;                // projected column within the group
    const int wpc = d_k >> 5;
    const int q0 = (q / wpc) * wpc;
    const uint32_t lane_base = tmem_base + ((uint32_t)(32 * q) << 16);
    const bool owner = (j % d_k) == 0;
    // ---- U^T of this column group -> tensor memory: lane j holds U^T[g*128 + j][0 .. D) as fp16 pairs; the four
    //      warps of a lane quarter write D/8 columns each
    {
      const int part_cols = C::kUCols / 4;                                // 32 (D = 256) or 16 (D = 128)
      const int c0 = (e >> 2) * part_cols;
      const uint4* src = reinterpret_cast<const uint4*>(Ut_hi + ((int64_t)g * kNG + j) * D + 2 * c0);
#pragma unroll
      for (int b = 0; b < part_cols / 16; ++b) {
        uint32_t v[16];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint4 w = __ldg(src + 4 * b + t);
          v[4 * t] = w.x; v[4 * t + 1] = w.y; v[4 * t + 2] = w.z; v[4 * t + 3] = w.w;
        }
        tmem_st16(lane_base + C::tU + c0 + 16 * b, v);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(u_full);
    }
    float ssq = 0.f;
    long long ea = 0, eb = 0, e0 = 0, e1 = 0;
    // column of the 16-wide chunk whose sum this lane holds after the transpose-reduce (bit 0 of the lane is a duplicate)
    const int mycol = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
    for (int i = set, sub = rb + set * nRB; sub < n_sub; sub += 2 * nRB, i += 2) {
      const int buf = set;
      if (prof) e0 = clock64();
      mbar_wait(&h_full[buf], (uint32_t)(i >> 1) & 1u);
      tc_fence_after();
      if (prof) e1 = clock64();
      const uint32_t tHA = lane_base + C::tH + 64 * buf + 16 * ch, tHC = tHA + kSub32;
      float* redb = red + buf * (4 * kSub32);
      uint32_t ha[16], hc[16];
      tmem_ld16(tHA, ha);
      tmem_ld16(tHC, hc);
      tmem_ld_wait();
      float pr[16];
#pragma unroll
      for (int t = 0; t < 16; ++t) pr[t] = __uint_as_float(ha[t]) * __uint_as_float(hc[t]);
      // transpose-reduce over the 32 lanes: 8 + 4 + 2 + 1 exchanges, then one add between the duplicate lanes
#pragma unroll
      for (int off = 16, n = 8; n >= 1; off >>= 1, n >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int t = 0; t < n; ++t) {
          const float send = upper ? pr[t] : pr[t + n];
          const float keep = upper ? pr[t + n] : pr[t];
          pr[t] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      pr[0] += __shfl_xor_sync(0xffffffffu, pr[0], 1);
      if ((lane & 1) == 0) redb[q * kSub32 + 16 * ch + mycol] = pr[0];
      asm volatile("bar.sync %0, 128;" ::"r"(1 + 2 * set + ch) : "memory");   // the 4 warps working on these 16 rows
      uint32_t pk[8], qk[8];
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        float4 s4 = *reinterpret_cast<const float4*>(&redb[q0 * kSub32 + 16 * ch + 4 * i4]);
#pragma unroll
        for (int w = 1; w < 4; ++w) {
          if (w < wpc) {
            const float4 o = *reinterpret_cast<const float4*>(&redb[(q0 + w) * kSub32 + 16 * ch + 4 * i4]);
            s4.x += o.x; s4.y += o.y; s4.z += o.z; s4.w += o.w;
          }
        }
        float g0 = fmaxf(s4.x, 0.f), g1 = fmaxf(s4.y, 0.f), g2 = fmaxf(s4.z, 0.f), g3 = fmaxf(s4.w, 0.f);
        if (owner) {
          const float t0 = g0 * inv_scale, t1 = g1 * inv_scale, t2 = g2 * inv_scale, t3 = g3 * inv_scale;
          ssq += t0 * t0 + t1 * t1 + t2 * t2 + t3 * t3;
        }
        g0 *= pq_scale; g1 *= pq_scale; g2 *= pq_scale; g3 *= pq_scale;
        const int t = 4 * i4;
        pk[2 * i4] = pack_h2_sat(g0 * __uint_as_float(hc[t]), g1 * __uint_as_float(hc[t + 1]));
        pk[2 * i4 + 1] = pack_h2_sat(g2 * __uint_as_float(hc[t + 2]), g3 * __uint_as_float(hc[t + 3]));
        qk[2 * i4] = pack_h2_sat(g0 * __uint_as_float(ha[t]), g1 * __uint_as_float(ha[t + 1]));
        qk[2 * i4 + 1] = pack_h2_sat(g2 * __uint_as_float(ha[t + 2]), g3 * __uint_as_float(ha[t + 3]));
      }
      // packed fp16 pairs of rows 16*ch .. 16*ch+15 go to the first 8 of this half-chunk's OWN 16 columns (the other
      // half-chunk's warps may still be reading theirs)
      tmem_st8(tHA, pk);
      tmem_st8(tHC, qk);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[2 * buf + ch]);
      if (prof) { ea += e1 - e0; eb += clock64() - e1; }
    }
    if (prof && blockIdx.x == 0 && warp == 2 && lane == 0) { prof[3] = ea; prof[4] = eb; }
    // ---- final: X^T of this CTA -> partial buffer [rb][g][j][D]
    mbar_wait(x_full, 0);
    tc_fence_after();
    float* dst = part + (((int64_t)rb * G + g) * kNG + j) * D;
#pragma unroll 1
    for (int cc = e >> 2; cc < D / 32; cc += 4) {
      uint32_t v[32];
      tmem_ld32(lane_base + C::tX + 32 * cc, v);
      tmem_ld_wait();
#pragma unroll
      for (int t = 0; t < 8; ++t)
        *reinterpret_cast<float4*>(dst + 32 * cc + 4 * t) =
            make_float4(__uint_as_float(v[4 * t]), __uint_as_float(v[4 * t + 1]), __uint_as_float(v[4 * t + 2]),
                        __uint_as_float(v[4 * t + 3]));
    }
    if (owner) ss_part[((int64_t)rb * (G * (kNG / d_k)) + g * (kNG / d_k) + j / d_k) * 4 + 2 * set + ch] = ssq;
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// sums[i*m + col] = x_scale * sum_rb part[(rb*G + col/128)*128 + col%128][i] ; sums[d*m + k] = sum_rb ss_part.
// A CTA reduces a tile of 16 columns x 32 channels with 512 threads: 128 float4 lanes x 4 groups that take the row
// blocks rb = grp, grp + 4, ... (m/16 x d/32 CTAs = 128 at cfg 2).  The kernel is bound by the latency of its dependent
// load rounds, not by bandwidth (19 MB of partials): with eight independent 16-byte reads per thread and four groups
// a thread needs ceil(nRB / 32) rounds (3 at cfg 2; the first version, one scalar lane per element, needed 10).  The
// groups are combined through shared memory in a fixed order (deterministic; replicas stay bit-identical).
constexpr int kRedGroups = 4;
__global__ void __launch_bounds__(128 * kRedGroups) tc_reduce_kernel(const float* __restrict__ part,
                                                                    const float* __restrict__ ss_part, int nRB, int G, int d,
                                                                    int m, int K, float x_scale, float* __restrict__ sums) {
  __shared__ float tile[kRedGroups][16][33];
  const int lane128 = threadIdx.x & 127, grp = threadIdx.x >> 7;
  const int q = lane128 & 7, cl = lane128 >> 3;             // channels 4q..4q+3 of column cl
  const int i0 = blockIdx.y * 32, c0 = blockIdx.x * 16;
  const int col = c0 + cl;
  const int gg = col >> 7, jj = col & 127;
  const float4* src = reinterpret_cast<const float4*>(part + ((int64_t)gg * 128 + jj) * d + i0) + q;
  const int64_t stride = (int64_t)G * 128 * d / 4;
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
  auto add = [](float4& a, const float4& v) { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; };
  int rb = grp;
  for (; rb + 7 * kRedGroups < nRB; rb += 8 * kRedGroups) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(src + (rb + u * kRedGroups) * stride);
    add(a0, v[0]); add(a1, v[1]); add(a0, v[2]); add(a1, v[3]);
    add(a0, v[4]); add(a1, v[5]); add(a0, v[6]); add(a1, v[7]);
  }
  for (; rb < nRB; rb += kRedGroups) add(a0, __ldg(src + rb * stride));
  tile[grp][cl][4 * q] = a0.x + a1.x;
  tile[grp][cl][4 * q + 1] = a0.y + a1.y;
  tile[grp][cl][4 * q + 2] = a0.z + a1.z;
  tile[grp][cl][4 * q + 3] = a0.w + a1.w;
  __syncthreads();
  {
    const int oc = threadIdx.x & 15, oi = threadIdx.x >> 4;      // 16 consecutive columns of channel row oi (0..31)
    float v = 0.f;
#pragma unroll
    for (int g2 = 0; g2 < kRedGroups; ++g2) v += tile[g2][oc][oi];
    sums[(int64_t)(i0 + oi) * m + c0 + oc] = v * x_scale;
  }
  if (blockIdx.x == 0 && blockIdx.y == 0) {
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      float s = 0.f;
      for (int r = 0; r < nRB; ++r) {
        const float4 v = *reinterpret_cast<const float4*>(ss_part + ((int64_t)r * K + k) * 4);
        s += (v.x + v.y) + (v.z + v.w);
      }
      sums[(int64_t)d * m + k] = s;
    }
  }
}

__global__ void pack_f16_kernel(const float4* __restrict__ in, int64_t n4, float scale, uint2* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(in + i);
    uint2 o;
    o.x = pack_h2_sat(v.x * scale, v.y * scale);
    o.y = pack_h2_sat(v.z * scale, v.w * scale);
    out[i] = o;
  }
}
__global__ void pack_f16_tail_kernel(const float* __restrict__ in, int64_t begin, int64_t count, float scale,
                                     __half* __restrict__ out) {
  const int64_t i = begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < count) out[i] = __float2half_rn(fminf(fmaxf(in[i] * scale, -65504.f), 65504.f));
}

__global__ void __launch_bounds__(1024) absmax_kernel(const float* __restrict__ in, int64_t n, float* __restrict__ out) {
  __shared__ float red[32];
  float best = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    best = fmaxf(best, fabsf(__ldg(in + i)));
  best = warp_max(best);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_max(v);
    if (threadIdx.x == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(v));   // v >= 0: int order == float order
  }
}

// one warp per row: max_r ||in[r,:]||_2
__global__ void __launch_bounds__(256) rownorm_max_kernel(const float* __restrict__ in, int64_t rows, int d,
                                                          float* __restrict__ out) {
  const int lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float best = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * nw + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * nw) {
    float s = 0.f;
    for (int j = lane; j < d; j += 32) { const float v = __ldg(in + r * d + j); s = fmaf(v, v, s); }
    s = warp_sum(s);
    best = fmaxf(best, s);
  }
  if (lane == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(sqrtf(best)));
}

long long* g_tc_prof = nullptr;   // debug: per-phase cycle counters of CTA 0 (drsa_debug_set_tc_profile)

struct TcPlan { int G, nRB, num_tiles; int64_t part_bytes, ss_bytes; };
TcPlan plan_for(int64_t M, int d, int m, int K, int sub_rows = kSub) {
  TcPlan p;
  p.G = m / kNG;
  p.num_tiles = (int)((M + sub_rows - 1) / sub_rows);      // 64-row (or 32-row) subtiles
  int nrb = sm_count() / (p.G * (d > 256 ? d / 256 : 1));
  if (nrb < 1) nrb = 1;
  if (nrb > p.num_tiles) nrb = p.num_tiles;
  p.nRB = nrb;
  p.part_bytes = align_up((int64_t)p.nRB * p.G * kNG * d * 4, 256);
  p.ss_bytes = align_up((int64_t)p.nRB * K * 4 * 4, 256);
  return p;
}
}  // namespace

bool tc_shape_supported(int d, int m, int K) {
  if (K <= 0 || m % K != 0) return false;
  const int d_k = m / K;
  return (d == 128 || d == 256 || d == 512) && m % kNG == 0 && m <= d && (d_k == 32 || d_k == 64 || d_k == 128);
}

int64_t step_tc_workspace_bytes(int64_t M, int d, int m, int K) {
  TcPlan p = plan_for(M, d, m, K, kSub32);       // the 32-row variant never has fewer row blocks: upper bound for both
  return p.part_bytes + p.ss_bytes + 256;
}

// 0 (default): the shared-memory-operand kernel above; 1: U^T in tensor memory for d <= 256 in the single-pass mode.
// Measured at cfg 2: variant 1 is correct but slower (0.368 vs 0.333 ms): its GEMM1 MMAs have N = 64 and take ~76
// cycles each instead of the 32 the M*N/256 rule promises -- tcgen05.mma has a floor of roughly 64 cycles per
// instruction at M = 128, so only N = 256 shapes run at the full rate (GEMM1 at N = 256 in the first version of this
// kernel, GEMM2 everywhere).  Kept selectable for A/B runs (drsa_debug_set_tc_variant).
int g_tc_variant = 0;
void set_tc_variant(int v) { g_tc_variant = v; }

template <int D>
int launch_step32(int grid, cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmC, const void* Ut_hi, int n_sub,
                  int G, int nRB, int d_k, float inv_scale, float pq_scale, float* part, float* ss_part, int* err) {
  static bool attr_set = false;
  if (!attr_set) {
    DRSA_CUDA(cudaFuncSetAttribute(drsa_tc_step32_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   Cfg32<D>::kSmemBytes));
    attr_set = true;
  }
  drsa_tc_step32_kernel<D><<<grid, kThreads, Cfg32<D>::kSmemBytes, stream>>>(
      tmA, tmC, static_cast<const __half*>(Ut_hi), n_sub, G, nRB, d_k, inv_scale, pq_scale, part, ss_part, err, g_tc_prof);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

template <int D, bool kSplitU, typename... Args>
int launch_step(int grid, cudaStream_t stream, Args... args) {
  static bool attr_set = false;      // one process per GPU: a per-process flag is enough
  if (!attr_set) {
    DRSA_CUDA(cudaFuncSetAttribute(drsa_tc_step_kernel<D, kSplitU>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   Cfg<D, kSplitU>::kSmemBytes));
    attr_set = true;
  }
  drsa_tc_step_kernel<D, kSplitU><<<grid, kThreads, Cfg<D, kSplitU>::kSmemBytes, stream>>>(args...);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

// split_u = true: U^T = hi + lo, two MMAs per product (DRSA_PREC_TC_F16X2); false: U^T = fp16(U) only (DRSA_PREC_TC_F16)
int step_tc(const void* A16, const void* C16, const void* Ut_hi, const void* Ut_lo, int64_t M, int d, int m, int K,
            float scaleA, float scaleC, float pq_scale, bool split_u, float* sums, void* workspace, int64_t workspace_bytes,
            cudaStream_t stream) {
  if (!tc_shape_supported(d, m, K)) return DRSA_ERR_SHAPE;
  if (!split_u) Ut_lo = Ut_hi;       // never read; keeps the tensor map valid
  if (!aligned16(A16) || !aligned16(C16) || !aligned16(Ut_hi) || !aligned16(Ut_lo)) return DRSA_ERR_ALIGN;
  if (M >= ((int64_t)1 << 31)) return DRSA_ERR_SHAPE;
  const bool tmem_u = !split_u && d <= 256 && g_tc_variant == 1;
  TcPlan p = plan_for(M, d, m, K, tmem_u ? kSub32 : kSub);
  if (workspace_bytes < p.part_bytes + p.ss_bytes + 256) return DRSA_ERR_WORKSPACE;
  char* w = static_cast<char*>(workspace);
  float* part = reinterpret_cast<float*>(w); w += p.part_bytes;
  float* ss_part = reinterpret_cast<float*>(w); w += p.ss_bytes;
  int* err = reinterpret_cast<int*>(w);

  if (tmem_u) {
    CUtensorMap tA, tC;
    DRSA_TRY(make_tmap_f16_sw128(&tA, A16, (uint64_t)M, (uint64_t)d, kSub32));
    DRSA_TRY(make_tmap_f16_sw128(&tC, C16, (uint64_t)M, (uint64_t)d, kSub32));
    const float inv_s = 1.0f / (scaleA * scaleC);
    const int grid32 = p.nRB * p.G;
    DRSA_TRY(d == 256 ? launch_step32<256>(grid32, stream, tA, tC, Ut_hi, p.num_tiles, p.G, p.nRB, m / K, inv_s, pq_scale, part,
                                           ss_part, err)
                      : launch_step32<128>(grid32, stream, tA, tC, Ut_hi, p.num_tiles, p.G, p.nRB, m / K, inv_s, pq_scale, part,
                                           ss_part, err));
    dim3 rg(m / 16, d / 32);
    tc_reduce_kernel<<<rg, 128 * kRedGroups, 0, stream>>>(part, ss_part, p.nRB, p.G, d, m, K, inv_s * inv_s / pq_scale, sums);
    DRSA_LAUNCH_CHECK();
    return DRSA_OK;
  }
  CUtensorMap tmA, tmC, tmA2, tmC2, tmUh, tmUl;
  DRSA_TRY(make_tmap_f16_sw128(&tmA, A16, (uint64_t)M, (uint64_t)d, kSub));
  DRSA_TRY(make_tmap_f16_sw128(&tmC, C16, (uint64_t)M, (uint64_t)d, kSub));
  DRSA_TRY(make_tmap_f16_sw128(&tmA2, A16, (uint64_t)M, (uint64_t)d, 32));
  DRSA_TRY(make_tmap_f16_sw128(&tmC2, C16, (uint64_t)M, (uint64_t)d, 32));
  DRSA_TRY(make_tmap_f16_sw128(&tmUh, Ut_hi, (uint64_t)m, (uint64_t)d, kNG));
  DRSA_TRY(make_tmap_f16_sw128(&tmUl, Ut_lo, (uint64_t)m, (uint64_t)d, kNG));

  const float inv_scale = 1.0f / (scaleA * scaleC);
  const int d_k = m / K;
  const int grid = p.nRB * p.G * (d > 256 ? d / 256 : 1);
  if (d == 512 && split_u) return DRSA_ERR_SHAPE;      // U^T hi + lo of a column group (256 KB) does not fit in shared memory
  int st;
  if (d == 512)
    st = launch_step<512, false>(grid, stream, tmA, tmC, tmA2, tmC2, tmUh, tmUl, p.num_tiles, p.G, p.nRB, d_k, inv_scale,
                                 pq_scale, part, ss_part, err, g_tc_prof);
  else if (d == 256)
    st = split_u ? launch_step<256, true>(grid, stream, tmA, tmC, tmA2, tmC2, tmUh, tmUl, p.num_tiles, p.G, p.nRB, d_k, inv_scale,
                                          pq_scale, part, ss_part, err, g_tc_prof)
                 : launch_step<256, false>(grid, stream, tmA, tmC, tmA2, tmC2, tmUh, tmUl, p.num_tiles, p.G, p.nRB, d_k, inv_scale,
                                           pq_scale, part, ss_part, err, g_tc_prof);
  else
    st = split_u ? launch_step<128, true>(grid, stream, tmA, tmC, tmA2, tmC2, tmUh, tmUl, p.num_tiles, p.G, p.nRB, d_k, inv_scale,
                                          pq_scale, part, ss_part, err, g_tc_prof)
                 : launch_step<128, false>(grid, stream, tmA, tmC, tmA2, tmC2, tmUh, tmUl, p.num_tiles, p.G, p.nRB, d_k, inv_scale,
                                           pq_scale, part, ss_part, err, g_tc_prof);
  DRSA_TRY(st);
  dim3 rgrid(m / 16, d / 32);
  // X' = pq_scale * (sA sC)^2 * X
  tc_reduce_kernel<<<rgrid, 128 * kRedGroups, 0, stream>>>(part, ss_part, p.nRB, p.G, d, m, K, inv_scale * inv_scale / pq_scale, sums);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

int pack_f16(const float* in, int64_t count, float scale, void* out, cudaStream_t stream) {
  if (!aligned16(in) || (reinterpret_cast<uintptr_t>(out) & 7u) != 0) return DRSA_ERR_ALIGN;
  const int64_t n4 = count / 4;
  if (n4 > 0) {
    int64_t blocks = (n4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    pack_f16_kernel<<<(int)blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(in), n4, scale,
                                                     reinterpret_cast<uint2*>(out));
    DRSA_LAUNCH_CHECK();
  }
  if (count % 4) {
    pack_f16_tail_kernel<<<1, 32, 0, stream>>>(in, n4 * 4, count, scale, static_cast<__half*>(out));
    DRSA_LAUNCH_CHECK();
  }
  return DRSA_OK;
}

int absmax(const float* in, int64_t count, float* out, cudaStream_t stream) {
  DRSA_CUDA(cudaMemsetAsync(out, 0, sizeof(float), stream));
  int64_t blocks = (count + 1023) / 1024;
  if (blocks > 148 * 2) blocks = 148 * 2;
  if (blocks < 1) blocks = 1;
  absmax_kernel<<<(int)blocks, 1024, 0, stream>>>(in, count, out);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

extern long long* g_fused_prof;
void set_tc_profile(long long* p) { g_tc_prof = p; g_fused_prof = p; }

// numRegs, maxThreadsPerBlock, static shared bytes, local bytes, max dynamic shared bytes of the row-pass kernel
int tc_kernel_attrs(int d, int split, int* out5) {
  cudaFuncAttributes a;
  const void* fn = d == 512 ? (const void*)drsa_tc_step_kernel<512, false> : d == 256 ? (split ? (const void*)drsa_tc_step_kernel<256, true> : (const void*)drsa_tc_step_kernel<256, false>)
                            : (split ? (const void*)drsa_tc_step_kernel<128, true> : (const void*)drsa_tc_step_kernel<128, false>);
  DRSA_CUDA(cudaFuncGetAttributes(&a, fn));
  out5[0] = a.numRegs; out5[1] = a.maxThreadsPerBlock; out5[2] = (int)a.sharedSizeBytes; out5[3] = (int)a.localSizeBytes;
  out5[4] = a.maxDynamicSharedSizeBytes;
  return DRSA_OK;
}

int rownorm_max(const float* in, int64_t rows, int d, float* out, cudaStream_t stream) {
  DRSA_CUDA(cudaMemsetAsync(out, 0, sizeof(float), stream));
  int64_t blocks = (rows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  rownorm_max_kernel<<<(int)blocks, 256, 0, stream>>>(in, rows, d, out);
  DRSA_LAUNCH_CHECK();
  return DRSA_OK;
}

// ----------------------------------------------------------------------------- self tests
// Exercise the two descriptor flavours the fused kernel relies on, each in isolation, against
// a host computation of the same fp16 products:
//   variant 0: SS MMA, A and B both K-major 128-byte-swizzled TMA boxes (GEMM1 building block)
//   variant 1: TS MMA, A packed fp16 in TMEM (tcgen05.st), B an MN-major swizzled box (GEMM2)
namespace {
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int variant,
                     const __half* __restrict__ Pglob, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + kPanelBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kPanelBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_base = tmem_base + ((uint32_t)(32 * warp) << 16);

  if (variant == 1) {
    // every thread = one TMEM lane (row i of P); pack P[i][0..127] as fp16 pairs into columns 128..191
    const __half* prow = Pglob + (int64_t)threadIdx.x * 128;
    for (int c = 0; c < 4; ++c) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const __half2 h = __halves2half2(prow[32 * c + 2 * i], prow[32 * c + 2 * i + 1]);
        pk[i] = *reinterpret_cast<const uint32_t*>(&h);
      }
      tmem_st16(lane_base + 128 + 16 * c, pk);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();

  if (threadIdx.x == 0) {
    if (variant == 0) {
      mbar_expect_tx(&bars[0], 2 * kPanelBytes);
      tma_load_2d(sA, &tmA, &bars[0], 0, 0);
      tma_load_2d(sB, &tmB, &bars[0], 0, 0);
      mbar_wait(&bars[0], 0);
      tc_fence_after();
      constexpr uint32_t idesc = make_idesc_f16(128, 128, 0, 0);
      for (int kk = 0; kk < 4; ++kk)
        umma_ss_f16(tmem_base, make_smem_desc_sw128(smem_u32(sA) + kk * 32, 16, 1024),
                    make_smem_desc_sw128(smem_u32(sB) + kk * 32, 16, 1024), idesc, kk > 0);
    } else {
      mbar_expect_tx(&bars[0], kPanelBytes);
      tma_load_2d(sB, &tmB, &bars[0], 0, 0);
      mbar_wait(&bars[0], 0);
      tc_fence_after();
      constexpr uint32_t idesc = make_idesc_f16(128, 64, 0, 1);
      for (int ks = 0; ks < 8; ++ks)
        umma_ts_f16(tmem_base, tmem_base + 128 + 8 * ks, make_smem_desc_sw128(smem_u32(sB) + ks * 2048, kPanelBytes, 1024),
                    idesc, ks > 0);
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const int ncol = variant == 0 ? 128 : 64;
  for (int cc = 0; cc < ncol / 32; ++cc) {
    uint32_t v[32];
    tmem_ld32(lane_base + 32 * cc, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(int64_t)threadIdx.x * ncol + 32 * cc + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<256>(tmem_base); }
}
}  // namespace

int selftest_umma(int variant, float* max_err_host) {
  if (variant < 0 || variant > 1 || max_err_host == nullptr) return DRSA_ERR_ARG;
  DRSA_TRY(require_sm100());
  // host data: small integers / 8 so every product and partial sum is exact in fp32
  std::vector<__half> hA(128 * 128), hB(128 * 64);
  std::vector<float> fA(128 * 128), fB(128 * 64);
  unsigned s = 12345u + variant;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)((int)((s >> 16) % 17) - 8) / 8.0f; };
  for (size_t i = 0; i < hA.size(); ++i) { fA[i] = rnd(); hA[i] = __float2half(fA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { fB[i] = rnd(); hB[i] = __float2half(fB[i]); }
  __half *dA = nullptr, *dB = nullptr; float* dOut = nullptr;
  DRSA_CUDA(cudaMalloc(&dA, hA.size() * 2));
  DRSA_CUDA(cudaMalloc(&dB, hB.size() * 2));
  DRSA_CUDA(cudaMalloc(&dOut, 128 * 128 * 4));
  DRSA_CUDA(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  DRSA_CUDA(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  DRSA_CUDA(cudaMemset(dOut, 0, 128 * 128 * 4));
  CUtensorMap tmA, tmB;
  int ncol;
  std::vector<float> ref;
  if (variant == 0) {
    // A [128 x 64] = first 64 columns of hA viewed as [128 x 128]; B [128 x 64] = hB
    DRSA_TRY(make_tmap_f16_sw128(&tmA, dA, 128, 128, 128));
    DRSA_TRY(make_tmap_f16_sw128(&tmB, dB, 128, 64, 128));
    ncol = 128;
    ref.assign(128 * 128, 0.f);
    for (int i = 0; i < 128; ++i)
      for (int j = 0; j < 128; ++j) {
        float a = 0.f;
        for (int k = 0; k < 64; ++k) a += fA[i * 128 + k] * fB[j * 64 + k];
        ref[i * 128 + j] = a;
      }
  } else {
    // P [128 x 128] = hA (row i = TMEM lane, K = 128); B [K = 128 rows x N = 64] = hB
    DRSA_TRY(make_tmap_f16_sw128(&tmA, dA, 128, 128, 128));
    DRSA_TRY(make_tmap_f16_sw128(&tmB, dB, 128, 64, 128));
    ncol = 64;
    ref.assign(128 * 64, 0.f);
    for (int i = 0; i < 128; ++i)
      for (int n = 0; n < 64; ++n) {
        float a = 0.f;
        for (int k = 0; k < 128; ++k) a += fA[i * 128 + k] * fB[k * 64 + n];
        ref[i * 64 + n] = a;
      }
  }
  const int smem_bytes = 2 * kPanelBytes + 256;
  DRSA_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  umma_selftest_kernel<<<1, 128, smem_bytes>>>(tmA, tmB, variant, dA, dOut);
  DRSA_LAUNCH_CHECK();
  DRSA_CUDA(cudaDeviceSynchronize());
  std::vector<float> got(128 * ncol);
  DRSA_CUDA(cudaMemcpy(got.data(), dOut, got.size() * 4, cudaMemcpyDeviceToHost));
  float worst = 0.f;
  for (size_t i = 0; i < got.size(); ++i) {
    const float e = std::fabs(got[i] - ref[i]);
    if (!(e <= worst)) worst = std::isnan(e) ? 1e30f : e;
  }
  *max_err_host = worst;
  cudaFree(dA); cudaFree(dB); cudaFree(dOut);
  return DRSA_OK;
}

}  // namespace drsa
