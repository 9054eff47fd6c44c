// gram_h.cu -- K3, default tensor-core path: the TICA Gram matrices as tcgen05 kind::f16 products with
// fp64-grade accuracy.  Same idea as gram_tc.cu (an EXACTLY accumulating leading term plus residual terms),
// re-designed around what bounds the TF32 kernel on B200: that kernel is not MMA-bound but load-bound (every
// producer thread has one stage of global loads in flight), so halving its MMA work changed nothing.  Here
//   * the rows of a stage (16 consecutive frames = 16 KB, contiguous) arrive by ONE bulk copy (TMA engine) into a
//     shared-memory ring kept several stages ahead by a loader warp;
//   * operands are fp16 (kind::f16, K = 16 frames per instruction at the cycle cost of a TF32 K = 8 instruction,
//     half the shared-memory bytes per frame) and the A operand of every product is a VIEW of a B-side tile;
//   * accumulation windows end when an exactness budget is spent, not after a fixed number of frames.
//
//   selected frames g (see `sel`), z = NaN ? 0 : (x - shift) * scale              (mode 1: z_g - z_{g+lag})
//      k  = clamp(rint(8 z + dither_g), -1024, 1024)      a  = k / 8      (leading term: covers |z| <= 128)
//      r  = z - a  (exact in fp32),  r' = 4 r,  rh' = fp16(r'),  rl' = fp16(r' - rh')
//   stored tiles (fp16, MN-major: features contiguous inside a frame, 128-byte swizzle):
//      a' = 8 a = k      rh'      rl'
//   P += a'^T a'                             integers: the fp32 sums in TMEM are EXACT while
//                                            sum_t max_i k_i(t)^2 <= 2^24 over the window (|P_ij| <= max P_ii)
//   Q += a'^T rh' + a'^T rl' + rh'^T rh'     = 32 a r^T + 16 rh rh^T
//      sum z z^T = [P + 2 (Q + Q^T)] / 64        (dropped: rh rl^T + rl rh^T + rl rl^T, ~2^-11 |r|^2 per frame with
//                                                 |r| <= 1/8 inside the range of the leading term)
// The scales satisfy scale(rh')^2 = scale(a') scale(rh') / 2, which is what lets the A operands be views.
//
// Exactness budget.  Each producer warp (= one frame of the stage) publishes m^2, m = max_i |k_i|, with its stage;
// the MMA warp adds the 16 values of a stage to the window's budget and closes the window BEFORE a stage that
// would exceed 2^24 (or after kGhMaxWindow frames): it commits, raises the drain sequence number and waits until
// the 16 producer warps have moved the two accumulators into the fp64 workspace.  Typical data (max |z| ~ 3.3
// over 256 features) gives ~24 k frames per window; a burst of outliers shortens the windows, nothing else.
//
// Weights.  The TICA sums need  sum_g w_g z_g z_g^T  with w in {0, 1, 2} (mode 0) or {0, 1} (mode 1).  A view
// cannot carry a per-frame weight, so the launcher runs the kernel on frame SELECTIONS with a uniform weight:
//      mode 0:  2 * Gram(frames with w >= 1)  -  Gram(frames with w == 1)       (edge frames: 2 lag per shard)
//      mode 1:  Gram(frames with mask & 1)
// A stage without a selected frame is skipped by every role (per-CTA bitmap of non-empty stages in shared
// memory): the second pass of mode 0 costs the mask scan only.
//
// Range.  Beyond |z| = 128 the leading term is clamped and the residual grows; the dropped terms grow with its
// square, so beyond |z| = 132 (a z-score a single outlier among > 17 000 frames can reach) the producers raise a
// device flag and the launcher's follow-up launch of the SIMT kernel (gram.cu, fp32 products folded into fp64: any
// range) recomputes the matrix instead of returning at once -- no host round trip.  So does an infinite input.
//
// Tried and measured slower (kept out of the tree): the two row-block CTAs of a chunk as a cluster that shares the
// producer work, each CTA producing every other stage and storing the tiles into both CTAs' rings with
// st.shared::cluster (multicast commit frees a slot when both CTAs' MMAs have read it).  Correct, same accuracy,
// 8.4 / 9.9 ms per launch against 7.8 / 8.6 ms: 24 KB of remote stores per stage cost more than the arithmetic
// they save -- the same finding as the cta_group::2 variant of gram_tc.cu, from the other side.
//
// One CTA per (128-row block of G, frame chunk): 148 CTAs = 2 row blocks x 74 chunks for d = 256.
//   warps 0-15 producers: two teams of eight warps take alternate stages, a warp converts frames w and w + 8 of its
//              stage; a lane owns features [4 L, 4 L + 4) and [128 + 4 L, ...): two conflict-free 16-byte reads of
//              the raw row, six conflict-free 8-byte stores per frame; all sixteen warps drain
//   warp  16   MMA issuer: 4 tcgen05.mma (M = 128, N = d, K = 16) + 1 commit per stage from one elected block
//   warp  17   loader: bulk copies of raw rows, kGhRawStages ahead
#include <cuda_fp16.h>

#include "tc05.cuh"

namespace pmb {

// Role timing (build with -DPMB_GH_PROF): CTA 0 accumulates the cycles its roles spend waiting; read back with
// pmb_debug_counters_gram.
__device__ long long g_gh_dbg[16];
#ifdef PMB_GH_PROF
#define GH_T(var) const long long var = clock64()
#define GH_ACC(slot, a, b) gh_prof[slot] += (b) - (a)
#define GH_DECL long long gh_prof[4] = {0, 0, 0, 0}
#define GH_FLUSH(base, n) if (blockIdx.x == 0 && lane == 0) { for (int q_ = 0; q_ < (n); ++q_) g_gh_dbg[(base) + q_] = gh_prof[q_]; }
#else
#define GH_T(var)
#define GH_ACC(slot, a, b)
#define GH_DECL
#define GH_FLUSH(base, n)
#endif

constexpr int kGhBK = 16;                 // frames per stage = K of one kind::f16 instruction
constexpr int kGhStages = 4;              // fp16 tile ring
constexpr int kGhRawStages = 3;           // raw fp32 row ring
constexpr int kGhProdWarps = 16;
constexpr int kGhMmaWarp = kGhProdWarps;   // warp kGhProdWarps + 1 is the loader
constexpr int kGhThreads = (kGhProdWarps + 2) * 32;   // 576
constexpr int kGhMaxWindow = 16384;       // frames: bounds the fp32 accumulation length of the residual terms
constexpr int kGhKMax = 1024;             // |k| <= 1024  <=>  |a| <= 128 ; a' = k is an integer <= 2048: exact in fp16
constexpr uint32_t kGhBudget = 1u << 24;
constexpr uint32_t kGhTile = 256 * kGhBK * 2;          // bytes of one tile (256 features x 16 frames, fp16) = 8 KB
constexpr uint32_t kGhStageBytes = 3 * kGhTile;        // a' | rh' | rl'
constexpr uint32_t kGhRawBlock = 256 * kGhBK * 4;      // 16 rows of <= 256 floats
constexpr uint32_t kGhRawStageBytes = 2 * kGhRawBlock + 128; // rows g.., (mode 1) rows g + lag.., the stage's 16 mask bytes
constexpr int kGhBitmapWords = 2048;      // 65536 stages = 1 M frames per CTA
constexpr float kGhZMax = 132.0f;         // |z| above this: the residual of the clamped leading term is too large

struct GramHParams {
  const float* X;
  int64_t n;
  int d;
  int64_t ld;
  const uint8_t* mask;
  int lag;
  int mode;      // 0: z_g ; 1: z_g - z_{g+lag}
  int sel;       // 0: popcount(mask & 3) >= 1 ; 1: popcount(mask & 3) == 1 ; 2: mask & 1
  const float* shift;
  const float* scale;
  double* ws;    // [n_rb * n_chunks][2][d][128]
  int n_chunks;
  int64_t chunk; // frames per chunk, multiple of kGhBK
  unsigned int* flag;   // set to 1 when some |z| does not fit (the SIMT kernel then recomputes)
};

__device__ __forceinline__ bool gh_selected(int mk, int sel) {
  const int pc = __popc(mk & 3);
  return sel == 0 ? (pc >= 1) : (sel == 1 ? (pc == 1) : ((mk & 1) != 0));
}
__device__ __forceinline__ float gh_cond(float x, float sh, float sc4) { return (x == x) ? (x - sh) * sc4 : 0.0f; }

// (feature mn, frame k) -> byte offset of element (mn, k) inside an MN-major 128B-swizzled fp16 tile of 16 frames:
// atoms of 64 features x 8 frames (1024 B), adjacent in K 1024 B apart, adjacent in MN 2048 B apart; inside an
// atom the 16-byte chunk index (8 features) is XOR-ed with the frame row
__device__ __forceinline__ uint32_t gh_off(int mn, int k) {
  return ((uint32_t)mn >> 6) * 2048u + ((uint32_t)k >> 3) * 1024u + ((uint32_t)k & 7u) * 128u +
         (((((uint32_t)mn & 63u) >> 3) ^ ((uint32_t)k & 7u)) << 4) + ((uint32_t)mn & 7u) * 2u;
}
__host__ __device__ constexpr uint32_t gh_idesc(int M, int N) {
  // kind::f16: D = F32, A = B = F16 (format 0), both operands MN-major
  return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct GhBars {
  uint64_t full[kGhStages], empty[kGhStages], raw_full[kGhRawStages], raw_empty[kGhRawStages], wdone, drained;
  uint32_t tmem_slot;
  uint32_t n_active;
  volatile uint32_t drain_seq;    // number of drains requested so far (written by the MMA warp)
  volatile uint32_t finished;     // 0 while running; total number of drains + 1 once the last one is requested
  uint32_t budget[kGhStages][kGhProdWarps / 2];
};

// iterator over the set bits of the stage bitmap (every role walks the same sequence)
struct GhIter {
  const uint32_t* bm;
  int nwords, w;
  uint32_t bits;
  __device__ __forceinline__ void init(const uint32_t* b, int n) { bm = b; nwords = n; w = 0; bits = n > 0 ? b[0] : 0u; }
  __device__ __forceinline__ int next() {   // -1 at the end
    while (bits == 0u) {
      if (++w >= nwords) return -1;
      bits = bm[w];
    }
    const int b = __ffs((int)bits) - 1;
    bits &= bits - 1u;
    return w * 32 + b;
  }
};

// One producer warp's share of a window drain: wait for the window's MMAs, add its TMEM columns to the fp64
// workspace (the first drain overwrites), hand the accumulators back.
__device__ __forceinline__ void gh_drain(uint64_t* wdone, uint64_t* drained, uint32_t tmem, double* ws, int d, int warp,
                                      int lane, uint32_t my_drains) {
  // warp w: lane quarter w % 4, accumulator (w / 4) % 2: 0 = P (columns 0..d), 1 = Q (columns 256..256+d)
  mbar_wait(wdone, my_drains & 1u);
  tc::fence_after_sync();
  const bool first = my_drains == 0;
  const int quarter = warp & 3, which = (warp >> 2) & 1, part = warp >> 3, nparts = kGhProdWarps / 8;
  double* dst = ws + (size_t)which * d * 128 + quarter * 32 + lane;
  for (int cc = part * 32; cc < d; cc += 32 * nparts) {
    float v[32];
    tc::tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(which * 256 + cc), v);
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      double* e = dst + (size_t)(cc + q) * 128;
      *e = first ? (double)v[q] : (*e + (double)v[q]);
    }
  }
  tc::fence_before_sync();
  __syncwarp();
  if (lane == 0) tc::mbar_arrive(drained);
}

template <int MODE>
__global__ void __launch_bounds__(kGhThreads, 1) gram_h_kernel(GramHParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* raw = tiles + (size_t)kGhStages * kGhStageBytes;
  uint32_t* bitmap = reinterpret_cast<uint32_t*>(raw + (size_t)kGhRawStages * kGhRawStageBytes);
  GhBars* B = reinterpret_cast<GhBars*>(bitmap + kGhBitmapWords);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int d = p.d;
  const int rb = blockIdx.x / p.n_chunks, ck = blockIdx.x - rb * p.n_chunks;
  const int64_t g_begin = (int64_t)ck * p.chunk;
  int64_t g_end = g_begin + p.chunk;
  if (g_end > p.n) g_end = p.n;
  const int n_stages = g_begin < g_end ? (int)((g_end - g_begin + kGhBK - 1) / kGhBK) : 0;
  const int n_words = (n_stages + 31) / 32;
  const uint32_t row_bytes = (uint32_t)d * 4u;

  // zero the tiles once (features >= d stay zero) and build the bitmap of stages holding a selected frame
  for (uint32_t i = tid; i < kGhStages * kGhStageBytes / 16; i += kGhThreads)
    reinterpret_cast<uint4*>(tiles)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < n_words; i += kGhThreads) bitmap[i] = 0u;
  if (tid == 0) {
    for (int s = 0; s < kGhStages; ++s) {
      mbar_init(&B->full[s], kGhProdWarps / 2);
      mbar_init(&B->empty[s], 1);
    }
    for (int s = 0; s < kGhRawStages; ++s) {
      mbar_init(&B->raw_full[s], 1);
      mbar_init(&B->raw_empty[s], kGhProdWarps / 2);
    }
    mbar_init(&B->wdone, 1);
    mbar_init(&B->drained, kGhProdWarps);
    B->n_active = 0u;
    B->drain_seq = 0u;
    B->finished = 0u;
    fence_barrier_init();
  }
  if (warp == kGhMmaWarp) tc::tmem_alloc(&B->tmem_slot, 512);
  __syncthreads();
  {
    // one thread per stage: its 16 mask bytes
    int my_active = 0;
    for (int s = tid; s < n_stages; s += kGhThreads) {
      const int64_t g0 = g_begin + (int64_t)s * kGhBK;
      bool any = false;
      if (g0 + kGhBK <= g_end && ((reinterpret_cast<uintptr_t>(p.mask + g0) & 15) == 0)) {
        const uint4 m = *reinterpret_cast<const uint4*>(p.mask + g0);
        const uint32_t w4[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int b = 0; b < 4; ++b) any = any || gh_selected((int)((w4[q] >> (8 * b)) & 0xffu), p.sel);
      } else {
        for (int f = 0; f < kGhBK && g0 + f < g_end; ++f) any = any || gh_selected((int)p.mask[g0 + f], p.sel);
      }
      if (any) {
        atomicOr(&bitmap[s >> 5], 1u << (s & 31));
        ++my_active;
      }
    }
    if (my_active) atomicAdd(&B->n_active, (uint32_t)my_active);
  }
  fence_proxy_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = B->tmem_slot;
  const int n_active = (int)B->n_active;
  double* ws = p.ws + (size_t)blockIdx.x * 2 * d * 128;

  if (warp < kGhProdWarps) {
    // ============================================================ producers
    const int cA = lane * 4, cB = 128 + lane * 4;
    const bool okA = cA < d, okB = cB < d;
    float sh[8], sc[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      sh[q] = okA ? p.shift[cA + q] : 0.f;
      sc[q] = okA ? 4.0f * p.scale[cA + q] : 0.f;        // z4 = 4 z (power-of-two scaling: exact)
      sh[4 + q] = okB ? p.shift[cB + q] : 0.f;
      sc[4 + q] = okB ? 4.0f * p.scale[cB + q] : 0.f;
    }
    uint32_t my_drains = 0;
    auto drain = [&]() {
      gh_drain(&B->wdone, &B->drained, tmem, ws, d, warp, lane, my_drains);
      ++my_drains;
    };
    // wait on a barrier of the tile ring; meanwhile serve drain requests (the MMA warp stops consuming stages
    // while it waits for a drain, so the ring fills up and every producer warp ends up here)
    auto wait_serving = [&](uint64_t* bar, uint32_t parity) {
#pragma unroll 1
      for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
        if (mbar_try_wait(bar, parity)) return;
        if (B->drain_seq != my_drains) drain();
      }
      __trap();
    };

    GH_DECL;
    GhIter it;
    it.init(bitmap, n_words);
    bool bad = false;     // some |z| does not fit the fp16 residual (or is not finite)
    // Two teams of eight warps take alternate stages; a warp converts TWO frames of its stage (wi and wi + 8) one
    // after the other, so the per-stage costs that do not depend on the amount of data (three barrier
    // round trips, the proxy fence, the loop bookkeeping) are paid once per two frames.
    const int team = warp >> 3, wi = warp & 7;
    // one frame: raw row -> the three packed tile rows (8 features per lane); returns max |k|
    auto convert = [&](const unsigned char* R0, int fr, int64_t g, uint2 (&wa)[2], uint2 (&wh)[2], uint2 (&wl)[2]) -> uint32_t {
      const bool sel_now = (g < g_end) && gh_selected((int)R0[2 * kGhRawBlock + fr], p.sel);
      const unsigned char* R = R0 + (size_t)fr * row_bytes;
      float4 xa = make_float4(0.f, 0.f, 0.f, 0.f), xb = xa, ya = xa, yb = xa;
      if (okA) xa = *reinterpret_cast<const float4*>(R + cA * 4);
      if (okB) xb = *reinterpret_cast<const float4*>(R + cB * 4);
      if (MODE == 1) {
        if (okA) ya = *reinterpret_cast<const float4*>(R + kGhRawBlock + cA * 4);
        if (okB) yb = *reinterpret_cast<const float4*>(R + kGhRawBlock + cB * 4);
      }
      // one dither per frame (hash of the frame index): E[r | z] = 0, see gram_tc.cu
      float dith;
      {
        uint32_t h = (uint32_t)g * 0x9E3779B1u;
        h ^= h >> 15;
        h *= 0x85EBCA6Bu;
        h ^= h >> 13;
        dith = __uint_as_float((h >> 9) | 0x3f800000u) - 1.5f;   // uniform in [-0.5, 0.5)
      }
      const float xs[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
      const float ys[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
      // z4 = 4 z.  Lean path: no per-element NaN / range / clamp handling; one sum detects a non-finite value in
      // the lane's eight (NaN inputs are imputed with 0 by the slow path), the warp maximum of |k| detects a value
      // beyond the leading term's range.
      float z4[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        z4[q] = (xs[q] - sh[q]) * sc[q];
        if (MODE == 1) z4[q] -= (ys[q] - sh[q]) * sc[q];
      }
      const float chk = ((z4[0] + z4[1]) + (z4[2] + z4[3])) + ((z4[4] + z4[5]) + (z4[6] + z4[7]));
      if (!(fabsf(chk) < 3.0e38f)) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float v = gh_cond(xs[q], sh[q], sc[q]);
          if (MODE == 1) v -= gh_cond(ys[q], sh[q], sc[q]);
          if (!(fabsf(v) < 3.0e38f)) {   // an infinite input: out of range, the SIMT kernel decides
            bad = bad || sel_now;
            v = 0.f;
          }
          z4[q] = v;
        }
      }
      if (!sel_now) {                    // warp-uniform: the warp is one frame
#pragma unroll
        for (int q = 0; q < 8; ++q) z4[q] = 0.f;
      }
      float kq[8];
      float kabs = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        kq[q] = rintf(fmaf(z4[q], 2.0f, dith));          // k = rint(8 z + dither)
        kabs = fmaxf(kabs, fabsf(kq[q]));
      }
      uint32_t m = __reduce_max_sync(0xffffffffu, (uint32_t)fminf(kabs, 1.0e9f));
      if (m > (uint32_t)kGhKMax) {       // warp-uniform, rare: clamp the leading term, check the residual's range
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          bad = bad || !(fabsf(z4[q]) <= 4.0f * kGhZMax);
          kq[q] = fminf(fmaxf(kq[q], -(float)kGhKMax), (float)kGhKMax);
        }
        m = (uint32_t)kGhKMax;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float r0 = fmaf(kq[2 * q], -0.5f, z4[2 * q]);          // r' = 4 z - k / 2 = 4 (z - k / 8), exact
        const float r1 = fmaf(kq[2 * q + 1], -0.5f, z4[2 * q + 1]);
        const __half2 a2 = __floats2half2_rn(kq[2 * q], kq[2 * q + 1]);      // a' = k: integers <= 1024, exact
        const __half2 h2 = __floats2half2_rn(r0, r1);
        const float2 hf = __half22float2(h2);
        const __half2 l2 = __floats2half2_rn(r0 - hf.x, r1 - hf.y);
        uint32_t* pa = (q & 1) ? &wa[q >> 1].y : &wa[q >> 1].x;
        uint32_t* ph = (q & 1) ? &wh[q >> 1].y : &wh[q >> 1].x;
        uint32_t* pl = (q & 1) ? &wl[q >> 1].y : &wl[q >> 1].x;
        *pa = *reinterpret_cast<const uint32_t*>(&a2);
        *ph = *reinterpret_cast<const uint32_t*>(&h2);
        *pl = *reinterpret_cast<const uint32_t*>(&l2);
      }
      return m;
    };
    auto store = [&](unsigned char* T, int fr, const uint2 (&wa)[2], const uint2 (&wh)[2], const uint2 (&wl)[2]) {
      if (okA) {
        const uint32_t o = gh_off(cA, fr);
        *reinterpret_cast<uint2*>(T + o) = wa[0];
        *reinterpret_cast<uint2*>(T + kGhTile + o) = wh[0];
        *reinterpret_cast<uint2*>(T + 2 * kGhTile + o) = wl[0];
      }
      if (okB) {
        const uint32_t o = gh_off(cB, fr);
        *reinterpret_cast<uint2*>(T + o) = wa[1];
        *reinterpret_cast<uint2*>(T + kGhTile + o) = wh[1];
        *reinterpret_cast<uint2*>(T + 2 * kGhTile + o) = wl[1];
      }
    };
    for (int a = 0; a < n_active; ++a) {
      const int s = it.next();
      if ((a & 1) != team) continue;
      const int slot = a % kGhStages, rslot = a % kGhRawStages;
      const uint32_t use = (uint32_t)(a / kGhStages), ruse = (uint32_t)(a / kGhRawStages);
      const int64_t g0 = g_begin + (int64_t)s * kGhBK;
      // raw rows and the stage's mask bytes (no global access on this path)
      GH_T(t0);
      mbar_wait(&B->raw_full[rslot], ruse & 1u);
      GH_T(t1);
      GH_ACC(0, t0, t1);
      const unsigned char* R0 = raw + (size_t)rslot * kGhRawStageBytes;
      unsigned char* T = tiles + (size_t)slot * kGhStageBytes;
      uint2 wa[2], wh[2], wl[2];
      const uint32_t m0 = convert(R0, wi, g0 + wi, wa, wh, wl);
      GH_T(t2);
      wait_serving(&B->empty[slot], (use & 1u) ^ 1u);
      GH_T(t3);
      GH_ACC(1, t2, t3);
      GH_ACC(2, t1, t2);
      store(T, wi, wa, wh, wl);
      const uint32_t m1 = convert(R0, wi + 8, g0 + wi + 8, wa, wh, wl);
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&B->raw_empty[rslot]);      // both rows are in registers
      store(T, wi + 8, wa, wh, wl);
      if (lane == 0) B->budget[slot][wi] = m0 * m0 + m1 * m1;
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&B->full[slot]);
    }
    if (n_active == 0) {
      for (int i = tid; i < 2 * d * 128; i += kGhProdWarps * 32) ws[i] = 0.0;
    } else {
      // remaining drain requests, the last one included
#pragma unroll 1
      for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
        if (B->drain_seq != my_drains) drain();
        const uint32_t fin = B->finished;
        if (fin != 0u && fin - 1u == my_drains) break;
        if (spin == (1u << 28) - 1) __trap();
      }
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(p.flag, 1u);
    if (warp == 0) { GH_FLUSH(0, 3); }
  } else if (warp == kGhMmaWarp) {
    // ============================================================ MMA issuer
    // Descriptors of a stage = slot base + constants (14-bit address field, 16-byte units); the A operands are
    // views of the B tiles: rows rb * 128 .. of a' and of rh' start rb * 2 atoms (2048 B each) into the tile.
    const uint32_t idesc = gh_idesc(128, d);
    const uint32_t base = smem_u32(tiles);
    const uint64_t d0 = tc::smem_desc(base, 2048u, 1024u, tc::kLayoutSw128);     // tile a', slot 0
    const uint64_t oRh = kGhTile >> 4, oRl = (2 * kGhTile) >> 4, oA = ((uint32_t)rb * 4096u) >> 4,
                   oSlot = kGhStageBytes >> 4;
    const uint32_t bar_empty0 = smem_u32(&B->empty[0]), bar_wdone = smem_u32(&B->wdone);
    const int max_stages = kGhMaxWindow / kGhBK;
    GH_DECL;
    int slot = 0, in_window = 0;
    uint32_t use = 0, n_drains = 0, spent = 0;
    auto close_window = [&]() {
      asm volatile(
          "{\n\t"
          ".reg .pred q;\n\t"
          "elect.sync _|q, 0xffffffff;\n\t"
          "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
          "}" ::"r"(bar_wdone)
          : "memory");
      ++n_drains;
      __syncwarp();
      if (lane == 0) B->drain_seq = n_drains;
    };
    GH_T(m_begin);
    for (int a = 0; a < n_active; ++a) {
      GH_T(m0);
      mbar_wait(&B->full[slot], use & 1u);
      GH_T(m1);
      GH_ACC(0, m0, m1);
      const uint32_t b = __reduce_add_sync(0xffffffffu, lane < kGhProdWarps / 2 ? B->budget[slot][lane] : 0u);
      if (in_window > 0 && (spent + b > kGhBudget || in_window == max_stages)) {
        GH_T(m2);
        close_window();
        mbar_wait(&B->drained, (n_drains - 1u) & 1u);
        GH_T(m3);
        GH_ACC(1, m2, m3);
        in_window = 0;
        spent = 0;
      }
      tc::fence_after_sync();
      const uint64_t dS = d0 + (uint64_t)slot * oSlot;
      const uint32_t acc0 = in_window == 0 ? 0u : 1u;
      asm volatile(
          "{\n\t"
          ".reg .pred q, p;\n\t"
          ".reg .b64 aa, ah, bh, bl;\n\t"
          "elect.sync _|q, 0xffffffff;\n\t"
          "setp.ne.b32 p, %4, 0;\n\t"
          "add.u64 aa, %2, %7;\n\t"        // A view of a'
          "add.u64 bh, %2, %5;\n\t"        // rh'
          "add.u64 bl, %2, %6;\n\t"        // rl'
          "add.u64 ah, bh, %7;\n\t"        // A view of rh'
          "@q tcgen05.mma.cta_group::1.kind::f16 [%0], aa, %2, %3, p;\n\t"
          "@q tcgen05.mma.cta_group::1.kind::f16 [%1], aa, bh, %3, p;\n\t"
          "@q tcgen05.mma.cta_group::1.kind::f16 [%1], aa, bl, %3, 1;\n\t"
          "@q tcgen05.mma.cta_group::1.kind::f16 [%1], ah, bh, %3, 1;\n\t"
          "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
          "}" ::"r"(tmem),
          "r"(tmem + 256u), "l"(dS), "r"(idesc), "r"(acc0), "l"(oRh), "l"(oRl), "l"(oA),
          "r"(bar_empty0 + (uint32_t)slot * 8u)
          : "memory");
      spent += b;
      ++in_window;
      if (++slot == kGhStages) { slot = 0; ++use; }
    }
    if (n_active > 0) {
      close_window();
      __syncwarp();
      if (lane == 0) B->finished = n_drains + 1u;
      mbar_wait(&B->drained, (n_drains - 1u) & 1u);
    }
    GH_T(m_end);
    GH_ACC(2, m_begin, m_end);
#ifdef PMB_GH_PROF
    gh_prof[3] = (long long)n_drains;
#endif
    GH_FLUSH(4, 4);
  } else {
    // ============================================================ loader: raw rows of the active stages
    if (lane == 0) {
      GH_DECL;
      GhIter it;
      it.init(bitmap, n_words);
      int rslot = 0;
      uint32_t ruse = 0;
      for (int a = 0; a < n_active; ++a) {
        const int s = it.next();
        const int64_t g0 = g_begin + (int64_t)s * kGhBK;
        GH_T(l0);
        mbar_wait(&B->raw_empty[rslot], (ruse & 1u) ^ 1u);
        GH_T(l1);
        GH_ACC(0, l0, l1);
        unsigned char* dst = raw + (size_t)rslot * kGhRawStageBytes;
        int64_t rows0 = p.n - g0;
        if (rows0 > kGhBK) rows0 = kGhBK;
        int64_t rows1 = 0;
        if (MODE == 1) {
          rows1 = p.n - (g0 + p.lag);
          if (rows1 > kGhBK) rows1 = kGhBK;
          if (rows1 < 0) rows1 = 0;
        }
        // the stage's 16 mask bytes ride along (the last stage of the array may be shorter: plain stores, made
        // visible by the release of the arrive below)
        const bool mask_bulk = g0 + kGhBK <= p.n && ((reinterpret_cast<uintptr_t>(p.mask + g0) & 15) == 0);
        if (!mask_bulk)
          for (int f = 0; f < (int)rows0; ++f) dst[2 * kGhRawBlock + f] = p.mask[g0 + f];
        mbar_expect_tx(&B->raw_full[rslot], (uint32_t)(rows0 + rows1) * row_bytes + (mask_bulk ? 16u : 0u));
        if (mask_bulk) bulk_g2s(dst + 2 * kGhRawBlock, p.mask + g0, 16u, &B->raw_full[rslot]);
        if (p.ld == d) {
          bulk_g2s(dst, p.X + g0 * p.ld, (uint32_t)rows0 * row_bytes, &B->raw_full[rslot]);
          if (rows1 > 0)
            bulk_g2s(dst + kGhRawBlock, p.X + (g0 + p.lag) * p.ld, (uint32_t)rows1 * row_bytes, &B->raw_full[rslot]);
        } else {
          for (int r = 0; r < (int)rows0; ++r)
            bulk_g2s(dst + (size_t)r * row_bytes, p.X + (g0 + r) * p.ld, row_bytes, &B->raw_full[rslot]);
          for (int r = 0; r < (int)rows1; ++r)
            bulk_g2s(dst + kGhRawBlock + (size_t)r * row_bytes, p.X + (g0 + p.lag + r) * p.ld, row_bytes,
                     &B->raw_full[rslot]);
        }
        if (++rslot == kGhRawStages) { rslot = 0; ++ruse; }
      }
      GH_FLUSH(8, 1);
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == kGhMmaWarp) tc::tmem_dealloc(tmem, 512);
}

// PQ[0] = sum of P partials, PQ[1] = sum of Q partials over the frame chunks, fixed order
__global__ void gram_h_reduce_kernel(const double* __restrict__ ws, int n_chunks, int d, double* __restrict__ PQ) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;   // i fastest
  if (e >= d * d) return;
  const int i = e % d, j = e / d;
  const int rb = i >> 7, il = i & 127;
  double accP = 0.0, accQ = 0.0;
  for (int c = 0; c < n_chunks; ++c) {
    const double* w = ws + (size_t)(rb * n_chunks + c) * 2 * d * 128;
    accP += w[(size_t)j * 128 + il];
    accQ += w[(size_t)d * 128 + (size_t)j * 128 + il];
  }
  PQ[(size_t)i * d + j] = accP;
  PQ[(size_t)d * d + (size_t)i * d + j] = accQ;
}
// G (+)= coef * [P + 2 (Q + Q^T)] / 64, bit-symmetric
__global__ void gram_h_combine_kernel(const double* __restrict__ PQ, int d, double coef, int accumulate,
                                      double* __restrict__ G) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d * d) return;
  const int i = e / d, j = e - i * d;
  const double* P = PQ;
  const double* Q = PQ + (size_t)d * d;
  const int a = i < j ? i : j, b = i < j ? j : i;
  const double v = coef * (1.0 / 64.0) * (P[(size_t)a * d + b] + 2.0 * (Q[(size_t)a * d + b] + Q[(size_t)b * d + a]));
  G[e] = accumulate ? G[e] + v : v;
}

int gram_h_debug_counters(int64_t* out16) {
  long long h[16];
  PMB_CUDA(cudaMemcpyFromSymbol(h, g_gh_dbg, sizeof(h)));
  for (int i = 0; i < 16; ++i) out16[i] = h[i];
  return PMB_OK;
}

static inline int gh_row_blocks(int d) { return (d + 127) / 128; }
static inline int gh_chunks(int d) { return kNumSMs / gh_row_blocks(d); }

bool gram_h_supported(int d, int64_t ld, const float* X, int64_t n) {
  if (!(d >= 32 && d <= 256 && d % 32 == 0 && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0)) return false;
  const int nc = gh_chunks(d);
  const int64_t chunk = (n + nc - 1) / nc;
  return chunk / kGhBK + 1 <= (int64_t)kGhBitmapWords * 32;
}
// workspace: a 256-byte header (the flag), the partials and PQ
size_t gram_h_ws_bytes(int d) {
  if (d < 32 || d > 256 || d % 32 != 0) return 0;
  return 256 + ((size_t)gh_row_blocks(d) * gh_chunks(d) * 2 * d * 128 + (size_t)2 * d * d) * sizeof(double);
}

// gram.cu: the SIMT launches, executed only when *run_if != 0 on the device
int gram_simt(const float* X, int64_t n, int d, int64_t ld, const uint8_t* mask, int lag, int mode, const float* shift,
              const float* scale, double* G, void* ws, cudaStream_t st, const unsigned int* run_if);

int gram_h(const float* X, int64_t n, int d, int64_t ld, const uint8_t* mask, int lag, int mode, const float* shift,
           const float* scale, double* G, void* ws, size_t ws_bytes, cudaStream_t st) {
  PMB_REQUIRE(ws_bytes >= gram_h_ws_bytes(d), "pmb_gram: workspace too small for the fp16 tcgen05 path");
  GramHParams p;
  p.X = X; p.n = n; p.d = d; p.ld = ld; p.mask = mask; p.lag = lag; p.mode = mode; p.shift = shift; p.scale = scale;
  p.flag = static_cast<unsigned int*>(ws);
  void* body = static_cast<unsigned char*>(ws) + 256;
  p.ws = static_cast<double*>(body);
  const int nrb = gh_row_blocks(d), nc = gh_chunks(d);
  int64_t chunk = (n + nc - 1) / nc;
  chunk = ((chunk + kGhBK - 1) / kGhBK) * kGhBK;
  p.n_chunks = nc;
  p.chunk = chunk;
  double* PQ = p.ws + (size_t)nrb * nc * 2 * d * 128;
  PMB_CUDA(cudaMemsetAsync(p.flag, 0, sizeof(unsigned int), st));
  const size_t smem = (size_t)kGhStages * kGhStageBytes + (size_t)kGhRawStages * kGhRawStageBytes +
                      (size_t)kGhBitmapWords * 4 + sizeof(GhBars) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    PMB_CUDA(cudaFuncSetAttribute(gram_h_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PMB_CUDA(cudaFuncSetAttribute(gram_h_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int n_pass = mode == 0 ? 2 : 1;
  for (int pass = 0; pass < n_pass; ++pass) {
    p.sel = mode == 0 ? pass : 2;
    const double coef = mode == 0 ? (pass == 0 ? 2.0 : -1.0) : 1.0;
    if (mode == 0) gram_h_kernel<0><<<nrb * nc, kGhThreads, smem, st>>>(p);
    else           gram_h_kernel<1><<<nrb * nc, kGhThreads, smem, st>>>(p);
    PMB_LAUNCH_CHECK();
    gram_h_reduce_kernel<<<(d * d + 255) / 256, 256, 0, st>>>(p.ws, nc, d, PQ);
    PMB_LAUNCH_CHECK();
    gram_h_combine_kernel<<<(d * d + 255) / 256, 256, 0, st>>>(PQ, d, coef, pass, G);
    PMB_LAUNCH_CHECK();
  }
  // out-of-range data: the SIMT kernels recompute G (they return at once while the flag is clear)
  return gram_simt(X, n, d, ld, mask, lag, mode, shift, scale, G, body, st, p.flag);
}

}  // namespace pmb
