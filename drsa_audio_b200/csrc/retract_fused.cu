// The replicated tail of a DRSA step in ONE cooperative kernel: pooling scalars, ascent step and the
// polar retraction (drsa.py:102, :201-221, :224-238).  Same arithmetic as retract.cu (scaled
// Newton-Schulz, fp32) but without ~30 dependent launches: the d x m problem (<= 512 x 512) lives in L2,
// every sweep is two tile-GEMM phases separated by grid-wide barriers, and the convergence decision is
// taken on the device by all CTAs from the same residual, so there is no host round trip and no wasted
// sweep.  Tiles are 32 x 32 per CTA (256 threads, 2 x 2 outputs per thread, k-chunks of 32 double
// buffered through registers); d and m must be multiples of 32.
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace drsa {

long long* g_fused_prof = nullptr;     // debug: see drsa_debug_set_tc_profile

namespace {
constexpr int TS = 32;

struct FusedParams {
  const float* sums; double inv_M; const float* U; int d, m, K;
  float* U_out; __half* Ut_hi; __half* Ut_lo; float* obj_log; long long log_index;
  int max_iters; float tol2_m; int* status;
  float* Y; float* X0; float* X1; float* G;
  float* rowsum;   // [m/32][m] per-tile-column partial row sums of |G| (summed in fixed order: deterministic)
  float* resid;    // [max_iters + 2][gridDim.x] per-CTA partial residuals (summed in fixed order)
  float* fro;      // [(m/32)^2] per-tile partials of ||Y^T Y - I||_F^2 (decides whether the start needs scaling)
  int have_sums;   // 0: Y already holds the matrix to retract (drsa_polar_retract)
  int u_rounded;   // sums were evaluated at a rounded U: log f(U^) + <grad, U - U^> (first-order exact in the rounding);
                   // 1: U^ = fp16(U); 2: U^ = the stored Ut_hi, written with error feedback through Ut_lo
  float* corr;     // [gridDim.x] per-CTA partials of that inner product
  long long* prof; // debug: %globaltimer stamps of CTA 0 at the phase boundaries (drsa_debug_set_tc_profile), or NULL
  // peer exchange (world > 1): xbuf[r] = rank r's exchange buffer as mapped into this process (NVLink peer memory).
  // Layout of a buffer: 64 x u32 header ([0] = exchange counter of the owner), then 8-byte words {value, flag}
  // [parity][source rank][xstride] (peer_push / peer_read).
  int world, rank;
  float* xbuf[DRSA_MAX_PEERS];
  int64_t xstride;
};

constexpr int kXHeaderFloats = 64;
__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// The all-reduce of the row sums, fused into the head of the finish kernel (push model over NVLink peer memory), with the
// FLAG IN THE DATA: every float travels as an 8-byte word {value, flag} with flag = exchange counter + 1, stored into slot
// [parity][rank] of EVERY peer's buffer (16-byte stores of two words: each 8-byte half is self-consistent, which is all the
// protocol needs).  A reader polls the word itself until the flag matches, so there is no fence, no arrival counter and no
// separate signal: the latency of the exchange is one NVLink store.  (The previous version -- 16-byte data stores, one
// system-scope fence per CTA, one release-add per peer, a polling loop on the counter -- spent 33-40 us in the head of the
// kernel on 8 GPUs, profiles/r02_p2p_exchange_profile_n8.log.)  The reduced value of element i is the sum over ranks in rank
// order (own share read from `sums`), so every rank adds the same floats in the same order: replicas of U stay bit-identical.
// (The pushes must be single 16- / 8-byte stores -- value and flag of a word may not be torn apart: STG.E.128 / STG.E.64 in
// the SASS of both finish kernels, checked with cuobjdump; the polls are 8-byte ld.relaxed.sys.v2.)
// Two parities: a peer can run at most one exchange ahead (it needs this rank's words of exchange s + 1 before it can finish
// it, and those are pushed only after this rank has read everything of exchange s), so slot [s & 1] is never overwritten
// while it is still being read; a stale word of exchange s - 2 carries another flag.  Buffer: 64 x u32 header ([0] = exchange
// counter of the owner), then uint2 words[2 parities][world][xstride].
struct PeerView {
  unsigned flag;
  const uint2* inbox;      // words[parity][0] of this rank's own buffer
};

__device__ __forceinline__ PeerView peer_push(const FusedParams& p, int64_t total) {
  unsigned* hdr = reinterpret_cast<unsigned*>(p.xbuf[p.rank]);
  const unsigned s = __ldcg(hdr);
  const unsigned flag = s + 1u;
  const int par = (int)(s & 1u);
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gthreads = (int64_t)gridDim.x * blockDim.x;
  const int64_t slot = ((int64_t)par * p.world + p.rank) * p.xstride;
  const int64_t n2 = total >> 1;
  for (int64_t i = gtid; i < n2; i += gthreads) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(p.sums) + i);
    const uint4 w = make_uint4(__float_as_uint(v.x), flag, __float_as_uint(v.y), flag);
    for (int r = 0; r < p.world; ++r)
      if (r != p.rank) reinterpret_cast<uint4*>(reinterpret_cast<uint2*>(p.xbuf[r] + kXHeaderFloats) + slot)[i] = w;
  }
  if ((total & 1) && gtid == 0) {
    const uint2 w = make_uint2(__float_as_uint(p.sums[total - 1]), flag);
    for (int r = 0; r < p.world; ++r)
      if (r != p.rank) (reinterpret_cast<uint2*>(p.xbuf[r] + kXHeaderFloats) + slot)[total - 1] = w;
  }
  PeerView v;
  v.flag = flag;
  v.inbox = reinterpret_cast<const uint2*>(p.xbuf[p.rank] + kXHeaderFloats) + (int64_t)par * p.world * p.xstride;
  return v;
}

// Slow path of a read: the word had not arrived at the first look; poll it.
__device__ __noinline__ unsigned peer_poll(const uint2* w, unsigned flag) {
  unsigned a, f;
  unsigned polls = 0;
  long long t0 = 0;
  while (true) {
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(f) : "l"(w) : "memory");
    if (f == flag) return a;
    // a peer that never pushes (crashed rank, mismatched call sequence) must not hang the GPU: trap after 30 s
    if ((++polls & 0xfffu) == 0) {
      if (t0 == 0) t0 = global_ns();
      else if (global_ns() - t0 > 30000000000LL) __trap();
    }
  }
}

// Element i of the row sums reduced over the ranks (own share + the peers' words of THIS exchange), added in rank order.
// The first look at the peers' words is a plain (non-volatile) asm so that the compiler can put the loads of all peers --
// and of the neighbouring elements of an unrolled loop -- in flight together; only words that have not arrived yet go
// through the polling loop.
struct ReducedSums {
  const FusedParams& p;
  PeerView pv;
  struct Look { unsigned a[DRSA_MAX_PEERS], f[DRSA_MAX_PEERS]; float own; };
  __device__ __forceinline__ void look(int64_t i, Look& L) const {          // issue the loads of element i
    L.own = p.sums[i];
    if (p.world <= 1) return;
#pragma unroll
    for (int r = 0; r < DRSA_MAX_PEERS; ++r) {
      L.a[r] = 0u; L.f[r] = 0u;
      if (r < p.world && r != p.rank)
        asm("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(L.a[r]), "=r"(L.f[r]) : "l"(pv.inbox + (int64_t)r * p.xstride + i));
    }
  }
  __device__ __forceinline__ float take(int64_t i, Look& L) const {         // ... and reduce them (polling the late ones)
    if (p.world <= 1) return L.own;
    float v = 0.f;
#pragma unroll
    for (int r = 0; r < DRSA_MAX_PEERS; ++r) {
      if (r < p.world) {
        if (r == p.rank) v += L.own;
        else {
          if (L.f[r] != pv.flag) L.a[r] = peer_poll(pv.inbox + (int64_t)r * p.xstride + i, pv.flag);
          v += __uint_as_float(L.a[r]);
        }
      }
    }
    return v;
  }
  __device__ __forceinline__ float operator()(int64_t i) const {
    Look L;
    look(i, L);
    return take(i, L);
  }
};

__device__ __forceinline__ void stamp(const FusedParams& p, int& slot) {
  if (p.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && slot < 40) {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    p.prof[16 + slot] = t;
  }
  ++slot;
}

// C[32x32] = sum_k A_k(i) B_k(j) over the FULL contraction length, both operand panels staged in shared memory by
// cp.async (all loads of a tile are in flight at once; the panels land in up to four groups and the FMA loop starts
// on the first while the rest is still arriving -- the previous version streamed 32-wide chunks with one chunk of
// prefetch and was bound by eight dependent L2 round trips per tile):
//   MODE 0 (gram): A_k(i) = X[k*ld + i0 + i],  B_k(j) = X[k*ld + j0 + j]      panels k-major  [K][LDT]
//   MODE 1 (mul) : A_k(i) = X[(i0 + i)*ldx + k] (panel [32][K + 4]),  B_k(j) = T[k*ldt + j0 + j]
constexpr int LDT = 36;   // padded row stride (floats) of a k-major panel: 16-byte aligned rows, conflict-free float2 reads

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_pending(int n) {      // wait until at most n groups are pending
  if (n <= 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
  else if (n == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
  else if (n == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
  else asm volatile("cp.async.wait_group 3;" ::: "memory");
}

// kFuseY (MODE 0 only): the panels are loaded from U and turned into Y = U + coef_k * sums IN shared memory before the FMA
// loop (`fuse`), so that the ascent step needs no phase and no grid barrier of its own (single rank).
struct FuseY {
  const float* sums; const float* coef; int m, d_k;      // coef[k] in shared memory
  float* Y_out;            // global Y (needed by the first multiply); written for the A panel when write_a, never for B
  bool write_a;
};

// kScaleT (MODE 1 only): the B panel is loaded from the Gram matrix G and turned into the first Newton-Schulz factor
// T_0 = (1.5 I - 0.5 G / c) / sqrt(c) in shared memory (c = 1: unscaled start), so that neither T_0 nor the scaled start
// X_0 = Y / sqrt(c) is ever formed in a phase of its own: X_1 = X_0 (1.5 I - 0.5 X_0^T X_0) = Y T_0.
template <int MODE, bool kFuseY = false, bool kScaleT = false>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ A, int lda, int i0, const float* __restrict__ B,
                                          int ldb, int j0, int Kdim, float (&acc)[2][2], float* sm, FuseY* fuse = nullptr,
                                          float inv_c = 1.f, float inv_s = 1.f) {
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int lda_s = MODE == 0 ? LDT : Kdim + 4;
  float* As = sm;
  float* Bs = sm + (MODE == 0 ? Kdim * LDT : 32 * (Kdim + 4));
  const int groups = (Kdim % 128 == 0) ? 4 : ((Kdim % 64 == 0) ? 2 : 1);
  const int Kg = Kdim / groups;               // multiple of 32
  __syncthreads();                            // the previous tile's panels are no longer read
  for (int g = 0; g < groups; ++g) {
    const int k0 = g * Kg;
    for (int p = 0; p < Kg / 32; ++p) {
      const int idx = tid + 256 * p;
      {   // B panel (and the A panel of MODE 0): 8 segments of 16 bytes per k row
        const int k = k0 + (idx >> 3), seg = idx & 7;
        cp_async16(Bs + k * LDT + 4 * seg, B + (int64_t)k * ldb + j0 + 4 * seg);
        if (MODE == 0) cp_async16(As + k * LDT + 4 * seg, A + (int64_t)k * lda + i0 + 4 * seg);
        if (kFuseY) {      // the matching segments of `sums`, staged behind the two operand panels
          cp_async16(Bs + (2 * Kdim + k) * LDT + 4 * seg, fuse->sums + (int64_t)k * fuse->m + j0 + 4 * seg);
          cp_async16(As + (2 * Kdim + k) * LDT + 4 * seg, fuse->sums + (int64_t)k * fuse->m + i0 + 4 * seg);
        }
      }
      if (MODE == 1) {   // A panel: 32 rows, Kg/4 segments of this group per row
        const int spr = Kg >> 2;
        const int row = idx / spr, seg = idx % spr;
        cp_async16(As + row * lda_s + k0 + 4 * seg, A + (int64_t)(i0 + row) * lda + k0 + 4 * seg);
      }
    }
    cp_async_commit();
  }
  if (kScaleT) {
    cp_async_wait_pending(0);                  // this thread's own segments have landed; the barrier below publishes them
    for (int r = 0; r < Kdim / 32; ++r) {
      const int idx = tid + 256 * r;
      const int k = idx >> 3, seg = idx & 7;
      float* dst = Bs + k * LDT + 4 * seg;
      float4 g4 = *reinterpret_cast<float4*>(dst);
      const int c0 = j0 + 4 * seg;
      g4.x = ((k == c0 ? 1.5f : 0.f) - 0.5f * g4.x * inv_c) * inv_s;
      g4.y = ((k == c0 + 1 ? 1.5f : 0.f) - 0.5f * g4.y * inv_c) * inv_s;
      g4.z = ((k == c0 + 2 ? 1.5f : 0.f) - 0.5f * g4.z * inv_c) * inv_s;
      g4.w = ((k == c0 + 3 ? 1.5f : 0.f) - 0.5f * g4.w * inv_c) * inv_s;
      *reinterpret_cast<float4*>(dst) = g4;
    }
  }
  // Each warp takes 4 consecutive k of every 32 and accumulates a full 32 x 32 partial tile in registers (4 x 8 per
  // lane: 3 LDS.128 per 32 FMA); the eight partials are summed through shared memory at the end.
  const int warp = tid >> 5, lane = tid & 31;
  const int ri = lane >> 2, ci = lane & 3;
  float part[4][8];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) part[r][c] = 0.f;
  for (int g = 0; g < groups; ++g) {
    cp_async_wait_pending(groups - 1 - g);
    if (kFuseY) {
      // Y = U + coef * sums for the segments of group g THIS thread loaded (complete after the wait above; the barrier below
      // publishes them), while the later groups are still in flight: the load / FMA overlap of the plain Gram phase is kept
      FuseY& f = *fuse;
      for (int p = 0; p < Kg / 32; ++p) {
        const int idx = tid + 256 * p;
        const int k = g * Kg + (idx >> 3), seg = idx & 7;
#pragma unroll
        for (int ab = 0; ab < 2; ++ab) {
          const int c0 = (ab == 0 ? i0 : j0) + 4 * seg;
          float* dst = (ab == 0 ? As : Bs) + k * LDT + 4 * seg;
          const float4 sv = *reinterpret_cast<const float4*>(dst + 2 * Kdim * LDT);      // its staged sums segment
          const float cf = f.coef[c0 / f.d_k];                     // 4 consecutive columns share a concept (d_k % 4 == 0)
          float4 uu = *reinterpret_cast<float4*>(dst);
          // Y = fma(coef, sums, U) everywhere (here, objective_correction, finish_small_kernel): one rounding, so that the
          // fused and the separate ascent -- i.e. the NCCL and the peer-memory exchange -- give the same bits
          uu.x = __fmaf_rn(cf, sv.x, uu.x); uu.y = __fmaf_rn(cf, sv.y, uu.y);
          uu.z = __fmaf_rn(cf, sv.z, uu.z); uu.w = __fmaf_rn(cf, sv.w, uu.w);
          *reinterpret_cast<float4*>(dst) = uu;
          if (ab == 0 && f.write_a) *reinterpret_cast<float4*>(f.Y_out + (int64_t)k * f.m + c0) = uu;
        }
      }
    }
    __syncthreads();
    for (int kb = g * Kg; kb < (g + 1) * Kg; kb += 32) {
      const int k4 = kb + 4 * warp;
      float av[4][4];                         // [u = k offset][r = row of this lane]
      if (MODE == 0) {                        // rows 4*ri .. 4*ri+3, contiguous in the k-major panel
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 a = *reinterpret_cast<const float4*>(&As[(k4 + u) * LDT + 4 * ri]);
          av[u][0] = a.x; av[u][1] = a.y; av[u][2] = a.z; av[u][3] = a.w;
        }
      } else {                                // rows ri, ri+8, ri+16, ri+24 (conflict-free with the K+4 row stride)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float4 a = *reinterpret_cast<const float4*>(&As[(ri + 8 * r) * lda_s + k4]);
          av[0][r] = a.x; av[1][r] = a.y; av[2][r] = a.z; av[3][r] = a.w;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[(k4 + u) * LDT + 8 * ci]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[(k4 + u) * LDT + 8 * ci + 4]);
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 8; ++c) part[r][c] = fmaf(av[u][r], bv[c], part[r][c]);
      }
    }
  }
  __syncthreads();                            // every warp is done reading the panels: reuse them for the partial tiles
  float* red_t = sm + warp * (32 * 33);       // [32][33] per warp
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int row = MODE == 0 ? 4 * ri + r : ri + 8 * r;
#pragma unroll
    for (int c = 0; c < 8; ++c) red_t[row * 33 + 8 * ci + c] = part[r][c];
  }
  __syncthreads();
#pragma unroll
  for (int aa = 0; aa < 2; ++aa)
#pragma unroll
    for (int bb = 0; bb < 2; ++bb) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += sm[w * (32 * 33) + (2 * ty + aa) * 33 + 2 * tx + bb];     // fixed order
      acc[aa][bb] = v;
    }
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 8) t = red[threadIdx.x];
  if (threadIdx.x < 32) t = warp_sum(t);
  return t;   // valid in warp 0
}

// Sum of the per-CTA partials of one sweep, same order in every CTA and on every rank (bit-identical replicas).
__device__ __forceinline__ float grid_total(const float* part, float* red) {
  float v = 0.f;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) v += __ldcg(part + i);
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < 8; ++w) t += red[w];
  return t;       // identical in every thread
}

// Sum of `count` partials, same order in every CTA and on every rank.
__device__ __forceinline__ float fixed_total(const float* part, int count, float* red) {
  float v = 0.f;
  for (int i = threadIdx.x; i < count; i += blockDim.x) v += __ldcg(part + i);
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < 8; ++w) t += red[w];
  return t;       // identical in every thread
}

// U_out / Ut_hi / Ut_lo of element (r, cc) of the retracted matrix.
// u_rounded == 2 -- error feedback (first-order sigma-delta) on the per-step rounding of U: the residual the previous
// rounding left behind (kept in Ut_lo) is added before rounding again, so the rounding errors of successive steps cancel
// instead of accumulating along the flat directions of the objective.  Measured over the reference's 2 000 steps: the
// contribution of U's rounding to the final principal angle drops from 3.8e-4 .. 7e-3 rad to 1e-5 .. 1.4e-4.
__device__ __forceinline__ void write_outputs(const FusedParams& p, int r, int cc, float v) {
  p.U_out[(int64_t)r * p.m + cc] = v;
  if (p.Ut_hi != nullptr) {
    const int64_t o = (int64_t)cc * p.d + r;
    const float t = (p.u_rounded == 2 && p.Ut_lo != nullptr) ? v + __half2float(p.Ut_lo[o]) : v;
    const __half hi = __float2half_rn(t);
    p.Ut_hi[o] = hi;
    if (p.Ut_lo != nullptr) p.Ut_lo[o] = __float2half_rn(t - __half2float(hi));
  }
}

// First-order term of the objective for the tensor-core modes: the row sums were evaluated at the rounded matrix U^ the row
// pass read (fp16(U), or with error feedback the stored Ut_hi), and f(U) = f(U^) + <grad f(U^), U - U^> + O(|U - U^|^2).
// U^ is stored TRANSPOSED ([m][d], K-major for the row pass), so the inner product is taken block-wise: a 32 x 32 block
// of U and of the gradient is staged in shared memory (coalesced along the columns), the matching block of U^ is read
// coalesced along its own rows, and every CTA takes its share of the blocks.  (Evaluating it as <grad f, U> - f through
// Euler's identity -- f is homogeneous of degree 2 -- avoids reading U^ but makes the objective inherit the full rounding
// of the gradient accumulation: 2e-4 .. 1e-3 relative at 2 .. 2.5 M rows per GPU; measured and discarded.)
// write_y: the same pass also forms Y = U + grad (the ascent step of drsa.py:102), so that every element of the (possibly
// peer-reduced) row sums is read exactly once.
template <typename SumsFn>
__device__ __forceinline__ float objective_correction(const FusedParams& p, const float* coef, SumsFn S, float (*tu)[33],
                                                       float (*tg)[33], bool write_y) {
  const int d = p.d, m = p.m, d_k = m / p.K;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int tm = m / 32, td = d / 32;
  float corr = 0.f;
  for (int b = blockIdx.x; b < td * tm; b += gridDim.x) {
    const int k0 = 32 * (b / tm), c0 = 32 * (b % tm);
    typename SumsFn::Look look[4];       // the loads of the four rows in flight together
#pragma unroll
    for (int r = 0; r < 4; ++r) S.look((int64_t)(k0 + ty + 8 * r) * m + c0 + tx, look[r]);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int kk = ty + 8 * r;
      const int64_t i = (int64_t)(k0 + kk) * m + c0 + tx;
      const float u = p.U[i];
      const float cf = coef[(c0 + tx) / d_k], sv = S.take(i, look[r]);
      tu[kk][tx] = u;
      tg[kk][tx] = __fmul_rn(cf, sv);
      if (write_y) p.Y[i] = __fmaf_rn(cf, sv, u);
    }
    if (!p.u_rounded) continue;          // (uniform) only the ascent step was wanted: the tiles are not read
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int cc = ty + 8 * r;
      const float u = tu[tx][cc];
      const float uh = p.u_rounded == 2 ? __half2float(p.Ut_hi[(int64_t)(c0 + cc) * d + k0 + tx])
                                        : __half2float(__float2half_rn(u));
      corr = fmaf(tg[tx][cc], u - uh, corr);
    }
    __syncthreads();
  }
  return corr;
}

// Phases of a normal step on one rank (grid barriers in brackets):
//   Gram of Y = U + coef*X, formed on the fly in the operand panels  [1]  multiply by T = 1.5 I - 0.5 G  [2]  Gram  [3] ...
//   last multiply, which also writes U_out and the fp16 planes.
// i.e. 2 GEMM phases and ONE barrier when a single sweep suffices (late in an optimisation), 6 and 5 for three sweeps (its
// first steps); the previous version had separate ascent, scaling and output phases (3 more phases, 3 more barriers).
__global__ void __launch_bounds__(256) finish_fused_kernel(FusedParams p) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) float panels[];      // operand panels of tile_gemm (fused_smem_bytes)
  __shared__ float red[8];
  __shared__ float coef[64];
  __shared__ float bc[2];
  __shared__ float tile_u[32][33], tile_g[32][33];     // objective_correction
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int d = p.d, m = p.m;
  const int64_t n = (int64_t)d * m;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + tid, gthreads = (int64_t)gridDim.x * blockDim.x;
  int slot = 0;
  stamp(p, slot);
  // single rank, full step: the ascent step is fused into the first Gram phase
  const bool fuse0 = p.have_sums && p.world <= 1 && p.U_out != nullptr && p.K <= 64 && (m / p.K) % 4 == 0 &&
                     4 * d * LDT * 4 <= 200 * 1024;

  // ---------------- phase 0: pooling scalars, (Y = U + coef_k X_k), objective log
  if (p.have_sums) {
    const int K = p.K, d_k = m / K;
    PeerView pv{0u, nullptr};
    if (p.world > 1) pv = peer_push(p, n + K);
    const ReducedSums S{p, pv};                 // element i of the row sums reduced over the ranks, fixed order
    if (tid == 0) {
      float acc = 0.f; int degenerate = 0;
      for (int k = 0; k < K; ++k) {
        const float q = sqrtf((float)((double)S(n + k) * p.inv_M));
        if (q == 0.f) ++degenerate;
        acc += sqrtf(q);
      }
      bc[0] = acc / (float)K;
      if (blockIdx.x == 0 && p.status != nullptr) p.status[2] = degenerate;
    }
    __syncthreads();
    const float root = bc[0];
    const bool need_grad = p.U_out != nullptr || p.u_rounded;
    float corr = 0.f;
    if (need_grad) {
      for (int k = tid; k < K; k += blockDim.x) {
        const float q = sqrtf((float)((double)S(n + k) * p.inv_M));
        // K > 64 is exotic: the factor is then recomputed on the fly below
        if (K <= 64) coef[k] = (float)((double)root * p.inv_M / ((double)K * (double)q * sqrt((double)q)));
      }
      __syncthreads();
      if (K <= 64) {
        // ascent step (unless it is fused into the first Gram phase) and first-order objective term in ONE pass over
        // 32 x 32 blocks: every element of the row sums (a sum over the ranks' inbox slots under data parallelism) is read once
        const bool want_y = !fuse0 && p.U_out != nullptr;
        if (want_y || p.u_rounded) corr = objective_correction(p, coef, S, tile_u, tile_g, want_y);
      } else if (!fuse0) {
        for (int64_t i = gtid; i < n; i += gthreads) {     // exotic K > 64: factor recomputed per element, no tensor-core modes
          const int k = (int)(i % m) / d_k;
          const float q = sqrtf((float)((double)S(n + k) * p.inv_M));
          const float c = (float)((double)root * p.inv_M / ((double)K * (double)q * sqrt((double)q)));
          if (p.U_out != nullptr) p.Y[i] = __fmaf_rn(c, S(i), p.U[i]);
        }
      }
    }
    if (p.u_rounded) {
      const float tot = block_sum(corr, red);
      if (tid == 0) p.corr[blockIdx.x] = tot;
    }
    if (p.U_out == nullptr) {                   // objective only: a single CTA (uniform across the grid)
      if (tid == 0 && p.obj_log != nullptr) {
        long long idx = p.log_index;
        if (idx < 0) { idx = p.status[3]; p.status[3] = (int)idx + 1; }
        p.obj_log[idx] = root * root + (p.u_rounded ? p.corr[0] : 0.f);
      }
      __syncthreads();                          // single CTA here: every thread has read its words
      if (p.world > 1 && tid == 0) {            // close this exchange (see below)
        unsigned* hdr = reinterpret_cast<unsigned*>(p.xbuf[p.rank]);
        hdr[0] = __ldcg(hdr) + 1u;
      }
      return;
    }
  }
  if (!fuse0) {
    stamp(p, slot);
    grid.sync();
    stamp(p, slot);
  }
  if (p.have_sums && p.world > 1 && blockIdx.x == 0 && tid == 0) {
    // (after the grid barrier) every CTA has read the exchange counter and its words: advance the counter; the peers' next
    // words on this parity come two exchanges later, after this kernel has ended
    unsigned* hdr = reinterpret_cast<unsigned*>(p.xbuf[p.rank]);
    hdr[0] = __ldcg(hdr) + 1u;
  }
  auto log_objective = [&]() {
    if (p.have_sums && blockIdx.x == 0 && p.obj_log != nullptr) {      // block-uniform condition
      const float extra = p.u_rounded ? fixed_total(p.corr, (int)gridDim.x, red) : 0.f;   // fixed order: bit-identical replicas
      if (tid == 0) {
        long long idx = p.log_index;
        if (idx < 0) { idx = p.status[3]; p.status[3] = (int)idx + 1; }
        p.obj_log[idx] = bc[0] * bc[0] + extra;
      }
    }
  };
  if (!fuse0) log_objective();

  const int tm = m / TS, td = d / TS;
  // ---------------- phase 1: G = Y^T Y, row sums of |G|, ||G - I||_F^2, T for the unscaled start
  {
    for (int t = blockIdx.x; t < tm * tm; t += gridDim.x) {
      const int ti = t / tm, tj = t % tm;
      float acc[2][2];
      if (fuse0) {
        FuseY f{p.sums, coef, m, m / p.K, p.Y, ti == tj};
        tile_gemm<0, true>(p.U, m, ti * TS, p.U, m, tj * TS, d, acc, panels, &f);
      } else {
        tile_gemm<0>(p.Y, m, ti * TS, p.Y, m, tj * TS, d, acc, panels);
      }
      float fr = 0.f, sq = 0.f, tr = 0.f;
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int gi = ti * TS + 2 * ty + a;
        float rs = fabsf(acc[a][0]) + fabsf(acc[a][1]);
        const int gj = tj * TS + 2 * tx;
        *reinterpret_cast<float2*>(&p.G[(int64_t)gi * m + gj]) = make_float2(acc[a][0], acc[a][1]);
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const float e = acc[a][b] - (gi == gj + b ? 1.f : 0.f);
          fr = fmaf(e, e, fr);
          sq = fmaf(acc[a][b], acc[a][b], sq);
          if (gi == gj + b) tr += acc[a][b];
        }
        // the 16 threads of a half-warp share the row gi
        for (int o = 8; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
        if (tx == 0) p.rowsum[(int64_t)tj * m + gi] = rs;     // every (tj, gi) is written by exactly one tile
      }
      const float tot = block_sum(fr, red);
      const float tot_sq = block_sum(sq, red);
      const float tot_tr = block_sum(tr, red);
      if (tid == 0) { p.fro[t] = tot; p.fro[tm * tm + t] = tot_sq; p.fro[2 * tm * tm + t] = tot_tr; }
    }
  }
  stamp(p, slot);
  grid.sync();
  stamp(p, slot);
  if (fuse0) log_objective();
  // ---------------- start of the iteration (no phase, no barrier of its own)
  // Newton-Schulz maps a singular value s of the iterate to s (1.5 - 0.5 s^2) and converges to 1 for s in (0, sqrt 3); the
  // polar factor does not change when Y is scaled.  The objective is homogeneous of degree 2 in U, so the ascent step has
  // a large component along U itself (<grad f, U> = 2 f): Y = U + grad f is NOT close to orthonormal in any step of an
  // optimisation, its singular values sit in a band above 1.  X_0 = Y / sqrt(c) with c = tr(G) / m centres that band on 1
  // (the mean of the squared singular values becomes 1), which takes one sweep less than the one-sided safe scaling
  // c = ||G||_inf (used only when ||G||_inf / c >= 2.9, i.e. when the centred start could leave the convergence interval).
  // The first multiply forms T_0 = (1.5 I - 0.5 G / c) / sqrt(c) from G in its operand panel (tile_gemm<1, false, true>),
  // and the residual of the start, ||G / c - I||_F^2 = sum g^2 / c^2 - 2 tr G / c + m, comes from partials the Gram phase
  // left behind.  The decision is taken from the same partials, summed in the same order, by every CTA on every rank.
  float c, res0;
  {
    float best = 0.f;
    for (int i = tid; i < m; i += blockDim.x) {
      float rs = 0.f;
      for (int tj = 0; tj < tm; ++tj) rs += p.rowsum[(int64_t)tj * m + i];
      best = fmaxf(best, rs);
    }
    best = warp_max(best);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = best;
    __syncthreads();
    if (tid == 0) {
      float v = 0.f;
      for (int w = 0; w < 8; ++w) v = fmaxf(v, red[w]);
      bc[1] = v;
    }
    __syncthreads();
    const float ginf = bc[1];
    const float sum_sq = fixed_total(p.fro + tm * tm, tm * tm, red);
    const float trace = fixed_total(p.fro + 2 * tm * tm, tm * tm, red);
    c = trace / (float)m;
    if (!(c > 0.f) || !(ginf < 2.9f * c)) c = ginf;          // the centred start needs lambda_max(G) / c < 3
    res0 = fmaxf(sum_sq / (c * c) - 2.f * trace / c + (float)m, 0.f);
  }
  const float inv_c0 = 1.f / c, inv_s0 = rsqrtf(c);
  float* cur = p.Y;
  float* nxt = p.X1;

  // ---------------- Newton-Schulz sweeps
  int it = 0, converged = 0;
  bool wrote_out = false;
  while (true) {
    const float res = it == 0 ? res0 : grid_total(p.resid + (int64_t)it * gridDim.x, red);
    if (res < p.tol2_m) { converged = 1; break; }
    if (it >= p.max_iters) break;
    // In the quadratic regime ||G' - I||_F <= 0.75 ||G - I||_F^2 (eigenvalues g -> -0.75 g^2 + O(g^3)).  If that bound
    // is already below half the tolerance, this sweep is the last one: its Gram matrix (one GEMM phase and one grid
    // barrier, only needed to confirm convergence) is not formed and the multiply writes the outputs itself.
    const bool last = 0.5625f * res * res < 0.25f * p.tol2_m;
    // nxt = cur * T
    for (int t = blockIdx.x; t < td * tm; t += gridDim.x) {
      const int tr = t / tm, tj = t % tm;
      float acc[2][2];
      if (it == 0) tile_gemm<1, false, true>(cur, m, tr * TS, p.G, m, tj * TS, m, acc, panels, nullptr, inv_c0, inv_s0);
      else tile_gemm<1>(cur, m, tr * TS, p.G, m, tj * TS, m, acc, panels);
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int gi = tr * TS + 2 * ty + a;
        if (last) {
          write_outputs(p, gi, tj * TS + 2 * tx, acc[a][0]);
          write_outputs(p, gi, tj * TS + 2 * tx + 1, acc[a][1]);
        } else {
          *reinterpret_cast<float2*>(&nxt[(int64_t)gi * m + tj * TS + 2 * tx]) = make_float2(acc[a][0], acc[a][1]);
        }
      }
    }
    if (last) { ++it; converged = 1; wrote_out = true; break; }
    stamp(p, slot);
    grid.sync();
    stamp(p, slot);
    // G = nxt^T nxt -> residual, T
    float r = 0.f;
    for (int t = blockIdx.x; t < tm * tm; t += gridDim.x) {
      const int ti = t / tm, tj = t % tm;
      float acc[2][2];
      tile_gemm<0>(nxt, m, ti * TS, nxt, m, tj * TS, d, acc, panels);
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int gi = ti * TS + 2 * ty + a;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int gj = tj * TS + 2 * tx + b;
          const float e = acc[a][b] - (gi == gj ? 1.f : 0.f);
          r = fmaf(e, e, r);
          acc[a][b] = (gi == gj ? 1.5f : 0.f) - 0.5f * acc[a][b];
        }
        *reinterpret_cast<float2*>(&p.G[(int64_t)gi * m + tj * TS + 2 * tx]) = make_float2(acc[a][0], acc[a][1]);
      }
    }
    {
      const float tot = block_sum(r, red);
      if (tid == 0) p.resid[(int64_t)(it + 1) * gridDim.x + blockIdx.x] = tot;
    }
    stamp(p, slot);
    grid.sync();
    stamp(p, slot);
    // the buffer that held the previous iterate is free now (the unscaled start leaves Y untouched: X0 takes its place)
    float* freebuf = (cur == p.Y) ? p.X0 : cur;
    cur = nxt; nxt = freebuf;
    ++it;
  }
  // ---------------- output (only when the iteration did not end in a multiply that wrote it)
  stamp(p, slot);
  if (blockIdx.x == 0 && tid == 0 && p.status != nullptr) { p.status[0] = it; if (!converged) p.status[1] += 1; }   // [1]: sticky count
  if (!wrote_out)
    for (int64_t i = gtid; i < n; i += gthreads) write_outputs(p, (int)(i / m), (int)(i % m), cur[i]);
  stamp(p, slot);
  if (p.prof != nullptr && blockIdx.x == 0 && tid == 0) p.prof[15] = slot;
}

// ---------------------------------------------------------------------------------------------------------------
// Small problems (d, m <= 64: BASELINE cfg 1, the toy CNN, arch B's first layers): the whole finish step in ONE CTA with
// every matrix in shared memory -- no cooperative launch, no grid barrier, no L2 round trips between the phases.  Same
// mathematics, same order of decisions as finish_fused_kernel.  256 threads, a 4 x 4 output tile per thread (d, m
// multiples of 32 and <= 64, so 16 x 16 threads cover 64 x 64).
constexpr int SLD = 68;      // row stride (floats) of the shared-memory matrices: float4-aligned rows

__device__ __forceinline__ float block_total256(float v, float* red) {     // identical in every thread, fixed order
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += red[w];
  return t;
}

// C[4ti..][4tj..] = sum_k A^T: gram = true  -> sum_k X[k][i] X[k][j]      (X: [rows = Kdim][SLD])
//                              gram = false -> sum_k X[i][k] T[k][j]      (X: [d][SLD], T: [m][SLD])
template <bool kGram>
__device__ __forceinline__ void small_gemm(const float* __restrict__ X, const float* __restrict__ T, int Kdim, int i0, int j0,
                                           float (&acc)[4][4]) {
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  if (kGram) {
    for (int k = 0; k < Kdim; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&X[k * SLD + i0]);
      const float4 bv = *reinterpret_cast<const float4*>(&X[k * SLD + j0]);
      const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(a4[a], b4[b], acc[a][b]);
    }
  } else {
    for (int k = 0; k < Kdim; k += 4) {
      float a4[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const float4 v = *reinterpret_cast<const float4*>(&X[(i0 + a) * SLD + k]);
        a4[a][0] = v.x; a4[a][1] = v.y; a4[a][2] = v.z; a4[a][3] = v.w;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 bv = *reinterpret_cast<const float4*>(&T[(k + u) * SLD + j0]);
        const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(a4[a][u], b4[b], acc[a][b]);
      }
    }
  }
}

__global__ void __launch_bounds__(256) finish_small_kernel(FusedParams p) {
  extern __shared__ __align__(16) float sm[];
  float* Ys = sm;                       // [d][SLD]   Y, then free
  float* Gs = sm + 64 * SLD;            // [m][SLD]   G / T
  float* Xa = sm + 2 * 64 * SLD;        // [d][SLD]
  float* Xb = sm + 3 * 64 * SLD;        // [d][SLD]
  float* rowpart = sm + 4 * 64 * SLD;   // [16][64] partial row sums of |G|
  __shared__ float red[8];
  __shared__ float coef[64];
  __shared__ float bc[2];
  const int tid = threadIdx.x, tj = tid & 15, ti = tid >> 4;
  const int d = p.d, m = p.m;
  const int n = d * m;
  float gu = 0.f;
  if (p.have_sums) {
    const int K = p.K, d_k = m / K;
    PeerView pv{0u, nullptr};
    if (p.world > 1) pv = peer_push(p, n + K);              // gridDim.x == 1
    const ReducedSums S{p, pv};
    if (tid == 0) {
      float acc = 0.f; int degenerate = 0;
      for (int k = 0; k < K; ++k) {
        const float q = sqrtf((float)((double)S(n + k) * p.inv_M));
        if (q == 0.f) ++degenerate;
        acc += sqrtf(q);
      }
      bc[0] = acc / (float)K;
      if (p.status != nullptr) p.status[2] = degenerate;
    }
    __syncthreads();
    const float root = bc[0];
    for (int k = tid; k < K && k < 64; k += blockDim.x) {
      const float q = sqrtf((float)((double)S(n + k) * p.inv_M));
      coef[k] = (float)((double)root * p.inv_M / ((double)K * (double)q * sqrt((double)q)));
    }
    __syncthreads();
    if (p.U_out != nullptr || p.u_rounded) {
      // float4 per thread and iteration, the loads of four iterations in flight together (n is a multiple of 1024)
      if (p.world <= 1 && !p.u_rounded) {
        const float4* U4 = reinterpret_cast<const float4*>(p.U);
        const float4* S4 = reinterpret_cast<const float4*>(p.sums);
        for (int i4 = tid; i4 < n / 4; i4 += 4 * blockDim.x) {
          float4 uv[4], sv[4];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (i4 + q * blockDim.x < n / 4) { uv[q] = U4[i4 + q * blockDim.x]; sv[q] = S4[i4 + q * blockDim.x]; }
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (i4 + q * blockDim.x < n / 4) {
              const int i = 4 * (i4 + q * blockDim.x), r = i / m, c = i % m;
              const float cf = coef[c / d_k];            // d_k % 4 == 0 or the four columns are handled one by one below
              float* dst = &Ys[r * SLD + c];
              if (d_k % 4 == 0) {
                dst[0] = __fmaf_rn(cf, sv[q].x, uv[q].x); dst[1] = __fmaf_rn(cf, sv[q].y, uv[q].y);
                dst[2] = __fmaf_rn(cf, sv[q].z, uv[q].z); dst[3] = __fmaf_rn(cf, sv[q].w, uv[q].w);
              } else {
                dst[0] = __fmaf_rn(coef[c / d_k], sv[q].x, uv[q].x); dst[1] = __fmaf_rn(coef[(c + 1) / d_k], sv[q].y, uv[q].y);
                dst[2] = __fmaf_rn(coef[(c + 2) / d_k], sv[q].z, uv[q].z); dst[3] = __fmaf_rn(coef[(c + 3) / d_k], sv[q].w, uv[q].w);
              }
            }
        }
      } else
#pragma unroll 4
      for (int i = tid; i < n; i += blockDim.x) {
        const int r = i / m, c = i % m;
        const float cf = coef[c / d_k], sv = S(i);
        const float u = p.U[i], gr = __fmul_rn(cf, sv);
        Ys[r * SLD + c] = __fmaf_rn(cf, sv, u);
        if (p.u_rounded)         // first-order objective term, see objective_correction
          gu = fmaf(gr, u - (p.u_rounded == 2 ? __half2float(p.Ut_hi[(int64_t)c * d + r]) : __half2float(__float2half_rn(u))), gu);
      }
    }
    const float gu_tot = p.u_rounded ? block_total256(gu, red) : 0.f;
    __syncthreads();                        // every thread has read its share of the inbox
    if (p.world > 1 && tid == 0) {          // single CTA: close this exchange right away
      unsigned* hdr = reinterpret_cast<unsigned*>(p.xbuf[p.rank]);
      hdr[0] = __ldcg(hdr) + 1u;
    }
    if (tid == 0 && p.obj_log != nullptr) {
      long long idx = p.log_index;
      if (idx < 0) { idx = p.status[3]; p.status[3] = (int)idx + 1; }
      p.obj_log[idx] = root * root + gu_tot;
    }
    if (p.U_out == nullptr) return;
  } else {
    for (int i = tid; i < n; i += blockDim.x) Ys[(i / m) * SLD + i % m] = p.Y[i];
  }
  __syncthreads();
  const int i0 = 4 * ti, j0 = 4 * tj;
  const bool in_g = i0 < m && j0 < m;        // this thread owns a tile of the m x m Gram matrix
  const bool in_x = i0 < d && j0 < m;        // ... of the d x m iterate
  // ---- Gram of Y, statistics of the start
  float acc[4][4];
  float sq = 0.f, tr = 0.f;
  if (in_g) {
    small_gemm<true>(Ys, nullptr, d, i0, j0, acc);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float rs = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const float g = acc[a][b];
        Gs[(i0 + a) * SLD + j0 + b] = g;
        rs += fabsf(g);
        sq = fmaf(g, g, sq);
        if (i0 + a == j0 + b) tr += g;
      }
      rowpart[tj * 64 + i0 + a] = rs;
    }
  }
  const float sum_sq = block_total256(sq, red);
  const float trace = block_total256(tr, red);
  float best = 0.f;
  if (tid < m) {
    float rs = 0.f;
    for (int t = 0; t < m / 4; ++t) rs += rowpart[t * 64 + tid];
    best = rs;
  }
  best = warp_max(best);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = best;
  __syncthreads();
  float ginf = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) ginf = fmaxf(ginf, red[w]);
  float c = trace / (float)m;
  if (!(c > 0.f) || !(ginf < 2.9f * c)) c = ginf;
  float res = fmaxf(sum_sq / (c * c) - 2.f * trace / c + (float)m, 0.f);
  {
    const float inv_c = 1.f / c, inv_s = rsqrtf(c);
    __syncthreads();
    if (in_g) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int o = (i0 + a) * SLD + j0 + b;
          Gs[o] = ((i0 + a == j0 + b ? 1.5f : 0.f) - 0.5f * Gs[o] * inv_c) * inv_s;       // T_0 (own elements only)
        }
    }
    __syncthreads();
  }
  // ---- Newton-Schulz sweeps
  int it = 0, converged = 0;
  bool wrote_out = false;
  float* cur = Ys;
  float* nxt = Xa;
  while (true) {
    if (res < p.tol2_m) { converged = 1; break; }
    if (it >= p.max_iters) break;
    const bool last = 0.5625f * res * res < 0.25f * p.tol2_m;
    if (in_x) {
      small_gemm<false>(cur, Gs, m, i0, j0, acc);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          if (last) write_outputs(p, i0 + a, j0 + b, acc[a][b]);
          else nxt[(i0 + a) * SLD + j0 + b] = acc[a][b];
        }
    }
    if (last) { ++it; converged = 1; wrote_out = true; break; }
    __syncthreads();
    float r = 0.f;
    if (in_g) {
      small_gemm<true>(nxt, nullptr, d, i0, j0, acc);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const float e = acc[a][b] - (i0 + a == j0 + b ? 1.f : 0.f);
          r = fmaf(e, e, r);
          acc[a][b] = (i0 + a == j0 + b ? 1.5f : 0.f) - 0.5f * acc[a][b];
        }
    }
    res = block_total256(r, red);            // (barriers inside: every thread is done reading T)
    if (in_g) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) Gs[(i0 + a) * SLD + j0 + b] = acc[a][b];
    }
    __syncthreads();
    float* freebuf = (cur == Ys) ? Xb : cur;
    cur = nxt; nxt = freebuf;
    ++it;
  }
  if (tid == 0 && p.status != nullptr) { p.status[0] = it; if (!converged) p.status[1] += 1; }
  if (!wrote_out) {
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) write_outputs(p, i / m, i % m, cur[(i / m) * SLD + i % m]);
  }
}

constexpr int kSmallSmemBytes = (4 * 64 * SLD + 16 * 64) * 4;
bool finish_small_supported(int d, int m, int K) { return d <= 64 && m <= 64 && d % 32 == 0 && m % 32 == 0 && K >= 1 && K <= 64 && m % K == 0; }

// the fused ascent stages the `sums` panels behind the operand panels of the first Gram phase: 4 * d * LDT floats (147 KB at
// d = 256); where that does not fit (d = 512) the ascent keeps a phase of its own
bool fuse_ascent_fits(int d) { return 4 * d * LDT * 4 <= 200 * 1024; }
int fused_smem_bytes(int d, int m) {
  const int gram = (fuse_ascent_fits(d) ? 4 : 2) * d * LDT, mul = 32 * (m + 4) + m * LDT, red = 8 * 32 * 33;      // floats
  const int mx = gram > mul ? (gram > red ? gram : red) : (mul > red ? mul : red);
  return mx * 4;
}

int64_t fused_ws_bytes(int d, int m, int max_iters) {
  return align_up((int64_t)d * m * 4, 256) * 3 + align_up((int64_t)m * m * 4, 256) +
         align_up((int64_t)(m / TS) * m * 4, 256) + align_up((int64_t)(max_iters + 3) * 1024 * 4, 256);
}
}  // namespace

bool finish_fused_supported(int d, int m, int K) {
  return d % TS == 0 && m % TS == 0 && d >= TS && m >= TS && K >= 1 && fused_smem_bytes(d, m) <= 200 * 1024;
}

int64_t finish_fused_workspace_bytes(int d, int m) { return fused_ws_bytes(d, m, 64); }

// floats per (parity, source rank) slot of an exchange buffer, and the size of one rank's buffer
int64_t exchange_stride(int d, int m, int K) { return align_up((int64_t)d * m + K, 64); }
int64_t exchange_bytes(int d, int m, int K, int world) {
  return kXHeaderFloats * 4 + 2 * (int64_t)world * exchange_stride(d, m, K) * 8;      // 8-byte {value, flag} words
}

// have_sums = 1: full finish step; have_sums = 0: retract the matrix already stored in Y_in (copied by the caller).
int finish_fused(const float* sums, int64_t M_global, const float* U, int d, int m, int K, float* U_out, void* Ut_hi,
                 void* Ut_lo, float* obj_log, int64_t log_index, int max_iters, float tol, int u_rounded, int* status,
                 void* workspace, int64_t workspace_bytes, const float* Y_in, const drsa_peer_exchange* px,
                 cudaStream_t stream) {
  if (max_iters > 64) max_iters = 64;
  if (workspace_bytes < fused_ws_bytes(d, m, 64)) return DRSA_ERR_WORKSPACE;
  char* w = static_cast<char*>(workspace);
  FusedParams p{};
  const int64_t dm = align_up((int64_t)d * m * 4, 256);
  p.Y = reinterpret_cast<float*>(w); w += dm;
  p.X0 = reinterpret_cast<float*>(w); w += dm;
  p.X1 = reinterpret_cast<float*>(w); w += dm;
  p.G = reinterpret_cast<float*>(w); w += align_up((int64_t)m * m * 4, 256);
  p.rowsum = reinterpret_cast<float*>(w); w += align_up((int64_t)(m / TS) * m * 4, 256);
  p.resid = reinterpret_cast<float*>(w);
  p.sums = sums; p.inv_M = M_global > 0 ? 1.0 / (double)M_global : 0.0; p.U = U; p.d = d; p.m = m; p.K = K;
  p.U_out = U_out; p.Ut_hi = static_cast<__half*>(Ut_hi); p.Ut_lo = static_cast<__half*>(Ut_lo);
  p.obj_log = obj_log; p.log_index = log_index; p.max_iters = max_iters; p.tol2_m = tol * tol * (float)m;
  p.prof = g_fused_prof;
  p.status = status; p.have_sums = (Y_in == nullptr) ? 1 : 0;
  p.u_rounded = (u_rounded && Y_in == nullptr) ? u_rounded : 0;
  if (p.u_rounded == 2 && Ut_hi == nullptr) return DRSA_ERR_ARG;
  p.corr = p.resid + (int64_t)(max_iters + 1) * 1024;       // the row after the last sweep's residual partials
  p.fro = p.resid + (int64_t)(64 + 2) * 1024;                // last row of the 64 + 3 the workspace is sized for
  p.world = 1; p.rank = 0;
  if (px != nullptr && px->world > 1 && Y_in == nullptr) {
    if (px->world > DRSA_MAX_PEERS || px->rank < 0 || px->rank >= px->world) return DRSA_ERR_ARG;
    p.world = px->world; p.rank = px->rank;
    p.xstride = exchange_stride(d, m, K);
    for (int r = 0; r < px->world; ++r) {
      if (px->buffers[r] == nullptr || !aligned16(px->buffers[r])) return DRSA_ERR_ARG;
      p.xbuf[r] = static_cast<float*>(px->buffers[r]);
    }
  }
  if (Y_in != nullptr) DRSA_CUDA(cudaMemcpyAsync(p.Y, Y_in, (int64_t)d * m * 4, cudaMemcpyDeviceToDevice, stream));
  if (finish_small_supported(d, m, K) && g_fused_prof == nullptr) {
    static bool small_attr = false;
    if (!small_attr) {
      DRSA_CUDA(cudaFuncSetAttribute(finish_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallSmemBytes));
      small_attr = true;
    }
    finish_small_kernel<<<1, 256, kSmallSmemBytes, stream>>>(p);
    DRSA_LAUNCH_CHECK();
    return DRSA_OK;
  }
  int tiles = (d / TS) * (m / TS);
  const int t2 = (m / TS) * (m / TS);
  if (t2 > tiles) tiles = t2;
  int grid = tiles;
  const int smem = fused_smem_bytes(d, m);
  static int smem_set = 0;
  if (smem > smem_set) {
    DRSA_CUDA(cudaFuncSetAttribute(finish_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    smem_set = smem;
  }
  static int occ_smem = -1, occ_blocks = 0;          // one process per GPU: per-process cache of the occupancy query
  if (occ_smem != smem) {
    int per_sm = 0;
    DRSA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, finish_fused_kernel, 256, smem));
    occ_smem = smem; occ_blocks = per_sm * sm_count();
  }
  const int max_coresident = occ_blocks;
  if (max_coresident < 1) return DRSA_ERR_CUDA;
  if (grid > max_coresident) grid = max_coresident;
  if (grid > 1024) grid = 1024;            // resid partials are sized for <= 1024 CTAs
  if (U_out == nullptr) grid = 1;
  void* args[] = {&p};
  DRSA_CUDA(cudaLaunchCooperativeKernel((const void*)finish_fused_kernel, dim3(grid), dim3(256), args, smem, stream));
  return DRSA_OK;
}

}  // namespace drsa
