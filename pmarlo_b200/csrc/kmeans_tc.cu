// kmeans_tc.cu -- K6 (tensor-core path): nearest-centre assignment as a tcgen05 score GEMM with the argmin
// fused into the TMEM epilogue, labels bit-exact against the fp64 oracle.
//
//   score(t,k) - |y_t|^2 = |c_k|^2 - 2 y_t.c_k          (one 128 x 128 x KS MMA tile per 128-centre chunk)
//
// Operands are FP16 (kind::f16, K = 16 per instruction: twice the K rate of kind::tf32 for the same 11-bit
// significands), made fp32-accurate by splitting along the MMA K dimension.  Everything is pre-scaled by a
// power of two s (centre coordinates land in [2^9, 2^10)), so that hi/lo pieces stay in fp16's normal range;
// scores scale by s^2, which no comparison below notices.  Every coordinate d contributes three K slots
//      A (frame side):  y_hi  y_lo  y_hi        B (centre side):  c'_hi  c'_hi  c'_lo     (c' = -2 c s)
// and |c|^2 s^2 rides in two more slots (two fp16 pieces of |c|^2 s^2 / 2^13 against the constant 2^13), so the
// accumulator that leaves TMEM is the squared distance minus |y|^2, which does not change the argmin.
// Slot order: 0-1 |c|^2 pieces, then 2 + 3 d + {0,1,2}; KS = 3 D + 2 rounded up to 16 (D = 10: 32 slots =
// two MMAs per chunk).  Products of two fp16 values are exact in the fp32 accumulator, like TF32 x TF32.
//
// The centre operand is laid out once per call by kmeans_tc_prep_kernel in global memory, chunk by chunk in
// the shared-memory image the MMA reads; CTAs bulk-copy (cp.async.bulk) chunks into a ring of shared-memory
// slots.  When all chunks fit (C4: 8 x 8 KB) they are loaded once and stay resident; otherwise (C5: D = 64,
// K = 5000, 40 x 52 KB) they stream through the ring once per frame tile, fed from L2.
//
// Pipeline (one persistent CTA per SM, 26 warps, warp-specialised, mbarrier hand-offs):
//   warps 0-15   epilogue  : warp (g, q) reads columns 32 g .. 32 g + 31 of lane quarter q of EVERY buffer with
//                            one tcgen05.ld, releases the buffer at once, then 3-input-min tree, screening
//                            against the row's threshold, running (best, second) as packed integer keys
//   warp  16     MMA issuer: per chunk KS/16 tcgen05.mma (M 128, N 128) into the next of FOUR TMEM buffers
//   warp  17     loader    : bulk copies of centre chunks into the shared-memory ring
//   warps 18-21  producers : build the A tile (128 frames x KS slots, K-major core-matrix layout) and the
//                            per-row screening threshold from the hinted centre (hint prefetched two tiles
//                            ahead, row and hinted centre one tile ahead)
//   warps 22-25  finalisers: merge the groups, certainty test, labels, fused Lloyd accumulation
//
// Exactness: a frame whose best and second-best scores are closer than a bound on the arithmetic
// error (operand rounding, dropped lo.lo terms, tensor-core accumulation, key truncation) is NOT
// labelled here: it goes to a list, and kmeans_recheck_kernel re-evaluates it against all centres
// in fp64 with the oracle's operation order (oracle/kmeans.py::sqdist_direct), one warp per frame.
// So do frames whose scaled coordinates leave fp16's range (|y_d| > 64 max|c|) and non-finite frames.
#include <cuda_fp16.h>

#include "tc05.cuh"

namespace pmb {

// Role timing (build with -DPMB_KM_PROF): CTA 0 accumulates the cycles each warp role spends waiting on
// its barriers and working; read back with pmb_debug_counters_kmeans.
__device__ long long g_km_dbg[16];
#ifdef PMB_KM_PROF
#define KM_T(var) const long long var = clock64()
#define KM_ACC(slot, a, b) km_prof[slot] += (b) - (a)
#define KM_DECL long long km_prof[4] = {0, 0, 0, 0}
#define KM_FLUSH(base, n) if (blockIdx.x == 0 && lane == 0) { for (int q_ = 0; q_ < (n); ++q_) g_km_dbg[(base) + q_] = km_prof[q_]; }
#else
#define KM_T(var)
#define KM_ACC(slot, a, b)
#define KM_DECL
#define KM_FLUSH(base, n)
#endif

#ifndef PMB_KM_EPI
#define PMB_KM_EPI 1   // 0: every epilogue warp reads one block of every buffer; 1: one warp group per ring slot;
                       // 2: two super-groups of eight warps on two ring slots each (measured slower, DESIGN.md 3.2 g)
#endif
constexpr int kTcTile = 128;     // frames per tile  (MMA M)
constexpr int kTcChunk = 128;    // centres per MMA  (MMA N) = columns of one TMEM buffer
constexpr int kTcRing = 4;       // TMEM buffers (4 x 128 columns = all 512)
constexpr int kTcKpadTo = 256;   // K is padded to a multiple of this (contract of pmb_kmeans_tc_scores)
constexpr int kTcEpiWarps = 16;
constexpr int kTcProdWarps = 4;
constexpr int kTcFinWarps = 4;
constexpr int kTcMmaWarp = kTcEpiWarps;
constexpr int kTcLoadWarp = kTcEpiWarps + 1;
constexpr int kTcProdWarp0 = kTcEpiWarps + 2;
constexpr int kTcFinWarp0 = kTcProdWarp0 + kTcProdWarps;
constexpr int kTcThreads = (kTcEpiWarps + 2 + kTcProdWarps + kTcFinWarps) * 32;   // 832
constexpr int kTcDReg = 16;      // coordinates of a frame held in registers (larger D: re-read from L1/L2)
constexpr int kTcMaxSlots = 16;  // upper bound of the shared-memory ring of centre chunks
constexpr float kTcErrScale = 1.9073486e-6f;   // 2^-19: envelope of the score error in units of (|y| + |c|max)^2
constexpr float kTcKeyTrunc = 1.0f + 7.6293945e-6f;  // 1 + 2^-17 (5 key bits of a 23-bit mantissa dropped)
constexpr float kTcNormUnit = 8192.0f;         // 2^13: the A-side constant the |c|^2 pieces are multiplied with
constexpr float kTcHalfMax = 6.0e4f;           // |scaled coordinate| above this does not fit fp16

struct KmTcParams {
  const float* Y;
  int64_t n;
  int D;
  int64_t ld;
  const double* centers;
  int K;
  int Kpad;   // K rounded up to a multiple of 256; padding rows copy centre K-1 with a large penalty
  int KS;     // K slots per row, multiple of 16
  int nslots; // shared-memory ring slots for centre chunks; resident when nslots == Kpad / 128
  int32_t* labels;
  const int32_t* hints;   // previous labels (may alias labels) or nullptr: only used to skip work
  double* sums;
  int64_t* counts;
  double* inertia;
  int64_t* n_rechecked;
  int* recheck_list;    // n entries
  int* recheck_count;   // zeroed by the host wrapper
  double* ysq;          // sum over frames of |y|^2 (zeroed by the host wrapper)
  // This call's sums / counts: `ncopies` private copies of [K x D sums | K counts] (zeroed by the host wrapper),
  // CTA b accumulates into copy b % ncopies.  All CTAs adding into ONE copy serialise on the L2 atomic unit (every
  // address takes ~10^4 adds per launch); copies at distinct addresses do not, and the commit kernel adds them up
  // in a fixed order.
  double* lsums;        // copy 0: K x D
  unsigned long long* lcounts;   // copy 0: K
  int ncopies;
  size_t copy_stride;   // in 8-byte words
  int accumulate;       // sums / counts / inertia requested
  float* dbg_scores;    // n x Kpad (tests only) or nullptr
  // written by kmeans_tc_prep_kernel
  unsigned char* Bg;    // Kpad / 128 chunks of 128 x KS fp16 in the K-major core-matrix image
  float* cs32;          // K x D: (float)c * s
  float* meta;          // [0] s, [1] max |c| s * 1.0001, [2] 1 / s^2
};

// byte offset of (row, slot) in a K-major fp16 operand tile: 8-row x 16-byte core matrices (8 slots wide),
// adjacent in K 128 B apart, adjacent in M/N `sbo` apart
__device__ __forceinline__ uint32_t off_kmajor_h(int row, int slot, uint32_t sbo) {
  return (uint32_t)(row >> 3) * sbo + (uint32_t)(slot >> 3) * 128u + (uint32_t)(row & 7) * 16u + (uint32_t)(slot & 7) * 2u;
}

// Instruction descriptor for kind::f16 with F16 operands and F32 accumulation (same fields as idesc_tf32 with
// A/B format 0 = F16).
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---------------------------------------------------------------------------------------------------
// Centre operand: scale, split, |c|^2 pieces, laid out chunk by chunk in the shared-memory image.
__global__ void __launch_bounds__(1024) kmeans_tc_prep_kernel(KmTcParams p) {
  __shared__ float s_red[32];
  __shared__ float s_scale;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = p.D, K = p.K, KS = p.KS;
  float amax = 0.f;
  for (int i = tid; i < K * D; i += blockDim.x) amax = fmaxf(amax, fabsf((float)p.centers[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if (lane == 0) s_red[warp] = amax;
  __syncthreads();
  if (tid == 0) {
    float m = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, s_red[w]);
    int e = 0;
    if (m > 0.f && m < 3.0e38f) e = 9 - ilogbf(m);     // m * 2^e in [2^9, 2^10)
    e = e < -100 ? -100 : (e > 100 ? 100 : e);
    s_scale = ldexpf(1.0f, e);
  }
  __syncthreads();
  const float sc = s_scale;
  const uint32_t sbo = (uint32_t)(KS / 8) * 128u;
  const size_t chunk_bytes = (size_t)kTcChunk * KS * 2;
  float cmax2 = 0.f;
  if (D <= kTcDReg) {
    // a centre's row of the operand image is KS / 8 pieces of 16 bytes (8 slots of one core-matrix row): built in
    // registers and stored as vectors, the D coordinate loads issued together (the generic loop below pays D
    // dependent L2 round trips and 2-byte stores: 23 us per launch, 22 launches per Lloyd run)
    for (int k = tid; k < p.Kpad; k += blockDim.x) {
      const int src = k < K ? k : K - 1;
      double cd[kTcDReg];
#pragma unroll
      for (int d = 0; d < kTcDReg; ++d) cd[d] = d < D ? p.centers[(size_t)src * D + d] : 0.0;
      __half hs[64];
#pragma unroll
      for (int sl = 0; sl < 64; ++sl) hs[sl] = __float2half_rn(0.f);
      double n2 = 0.0;
#pragma unroll
      for (int d = 0; d < kTcDReg; ++d) {
        if (d < D) {
          const float c32 = (float)cd[d];
          n2 = fma((double)c32, (double)c32, n2);
          const float cs = c32 * sc;
          if (k < K) p.cs32[(size_t)k * D + d] = cs;
          const float cm = -2.0f * cs;
          const __half hh = __float2half_rn(cm);
          hs[2 + 3 * d + 0] = hh;
          hs[2 + 3 * d + 1] = hh;
          hs[2 + 3 * d + 2] = __float2half_rn(cm - __half2float(hh));
        }
      }
      const double n2s = n2 * (double)sc * (double)sc;
      const double q = n2s / (double)kTcNormUnit;
      const float p1 = __half2float(__float2half_rn((float)q));
      hs[0] = __float2half_rn(p1);
      hs[1] = __float2half_rn(k < K ? (float)(q - (double)p1) : 30000.0f);
      if (k < K) cmax2 = fmaxf(cmax2, (float)n2s * 1.0001f);
      unsigned char* rowp = p.Bg + (size_t)(k / kTcChunk) * chunk_bytes + (uint32_t)((k % kTcChunk) >> 3) * sbo +
                            (uint32_t)(k & 7) * 16u;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        if (8 * g < KS) {
          uint4 w;
          w.x = (uint32_t)__half_as_ushort(hs[8 * g + 0]) | ((uint32_t)__half_as_ushort(hs[8 * g + 1]) << 16);
          w.y = (uint32_t)__half_as_ushort(hs[8 * g + 2]) | ((uint32_t)__half_as_ushort(hs[8 * g + 3]) << 16);
          w.z = (uint32_t)__half_as_ushort(hs[8 * g + 4]) | ((uint32_t)__half_as_ushort(hs[8 * g + 5]) << 16);
          w.w = (uint32_t)__half_as_ushort(hs[8 * g + 6]) | ((uint32_t)__half_as_ushort(hs[8 * g + 7]) << 16);
          *reinterpret_cast<uint4*>(rowp + (uint32_t)g * 128u) = w;
        }
      }
    }
  } else
  for (int k = tid; k < p.Kpad; k += blockDim.x) {
    unsigned char* base = p.Bg + (size_t)(k / kTcChunk) * chunk_bytes;
    const int r = k % kTcChunk;
    auto put = [&](int slot, float v) {
      *reinterpret_cast<__half*>(base + off_kmajor_h(r, slot, sbo)) = __float2half_rn(v);
    };
    const int src = k < K ? k : K - 1;
    double n2 = 0.0;
    for (int d = 0; d < D; ++d) {
      const float c32 = (float)p.centers[(size_t)src * D + d];
      n2 = fma((double)c32, (double)c32, n2);
      const float cs = c32 * sc;
      if (k < K) p.cs32[(size_t)k * D + d] = cs;
      const float cm = -2.0f * cs;
      const float hi = __half2float(__float2half_rn(cm));
      put(2 + 3 * d + 0, hi);
      put(2 + 3 * d + 1, hi);
      put(2 + 3 * d + 2, cm - hi);
    }
    for (int sl = 3 * D + 2; sl < KS; ++sl) put(sl, 0.f);
    const double n2s = n2 * (double)sc * (double)sc;
    const double q = n2s / (double)kTcNormUnit;
    const float p1 = __half2float(__float2half_rn((float)q));
    put(0, p1);
    // padding rows: centre K-1 again with a penalty of 2^13 * 30000 in score units, so they never win
    put(1, k < K ? (float)(q - (double)p1) : 30000.0f);
    if (k < K) cmax2 = fmaxf(cmax2, (float)n2s * 1.0001f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cmax2 = fmaxf(cmax2, __shfl_xor_sync(0xffffffffu, cmax2, o));
  __syncthreads();
  if (lane == 0) s_red[warp] = cmax2;
  __syncthreads();
  if (tid == 0) {
    float m = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, s_red[w]);
    p.meta[0] = sc;
    p.meta[1] = sqrtf(m) * 1.0001f;
    p.meta[2] = 1.0f / (sc * sc);
  }
}

struct TcSmem {
  uint64_t a_full[2], a_empty[2], t_full[kTcRing], t_empty[kTcRing], r_full[2], r_empty[2], thr_ready[4];
  uint64_t b_full[kTcMaxSlots], b_empty[kTcMaxSlots];
  uint32_t tmem_slot;
  float thr[4][kTcTile];   // per-row screening threshold of tile it (slot it & 3), minus |y|^2
  float xn2[4][kTcTile];   // |y|^2 of the row (scaled)
  int res_best[2][4][kTcTile];
  int res_second[2][4][kTcTile];
  int res_block[2][4][kTcTile];
  float res_thr[2][kTcTile];
  float res_xn2[2][kTcTile];
};

// DBG: also write the raw scores (tests).  INREG: D <= kTcDReg, a frame's coordinates live in registers.
// DRP: coordinate registers of a producer thread (12 when D <= 12: the producers hold three such arrays).
template <bool DBG, bool INREG, int DRP>
__global__ void __launch_bounds__(kTcThreads, 1) kmeans_tc_kernel(KmTcParams p) {
  static_assert(DRP <= kTcDReg, "DRP");
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KS = p.KS, D = p.D, K = p.K, Kpad = p.Kpad;
  const uint32_t lbo = 128u, sbo = (uint32_t)(KS / 8) * 128u;
  const uint32_t chunk_bytes = (uint32_t)kTcChunk * (uint32_t)KS * 2u;   // also the bytes of one A tile
  const int nslots = p.nslots;
  unsigned char* sB = smem_raw;                                   // nslots x 128 x KS fp16
  unsigned char* sA = sB + (size_t)nslots * chunk_bytes;          // 2 x 128 x KS fp16
  TcSmem* S = reinterpret_cast<TcSmem*>(sA + (size_t)2 * chunk_bytes);
  float* stage = reinterpret_cast<float*>(S + 1);                 // INREG finalisers: [4][32][kTcDReg + 1]
  const int n_chunks = Kpad / kTcChunk;
  const bool resident = nslots >= n_chunks;
  const int64_t n_tiles = (p.n + kTcTile - 1) / kTcTile;

  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&S->a_full[b], kTcProdWarps * 32);
      mbar_init(&S->a_empty[b], resident ? 2 : 1);
      mbar_init(&S->r_full[b], kTcEpiWarps * 32);
      mbar_init(&S->r_empty[b], kTcFinWarps * 32);
    }
    for (int b = 0; b < kTcRing; ++b) {
      mbar_init(&S->t_full[b], 1);
      mbar_init(&S->t_empty[b], PMB_KM_EPI == 0 ? kTcEpiWarps : (PMB_KM_EPI == 2 ? 8 : 4));
    }
    for (int b = 0; b < 4; ++b) mbar_init(&S->thr_ready[b], 1);
    for (int b = 0; b < kTcMaxSlots; ++b) {
      mbar_init(&S->b_full[b], 1);
      mbar_init(&S->b_empty[b], 1);
    }
    fence_barrier_init();
  }
  if (warp == kTcMmaWarp) tc::tmem_alloc(&S->tmem_slot, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = S->tmem_slot;
  const float scale = p.meta[0], cmax = p.meta[1];

  if (warp < kTcEpiWarps) {
    // =========================================================== epilogue warps
    const int quarter = warp & 3, grp = warp >> 2;
    const int row = quarter * 32 + lane;
    const float inv_s2 = p.meta[2];
    uint32_t it = 0;
    uint32_t uses = 0;   // group-per-buffer variant: how many times this group's TMEM buffer has been filled
    (void)uses;
    KM_DECL;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      int best = 0x7fffffff, second = 0x7fffffff, bblock = 0;
      // screening threshold of this row (written by the producers, handed over by the MMA warp)
      KM_T(e0);
      mbar_wait(&S->thr_ready[it & 3u], (it >> 2) & 1u);
      KM_T(e1);
      KM_ACC(0, e0, e1);
      const float thr = S->thr[it & 3u][row];
      const float xn2 = S->xn2[it & 3u][row];
#if PMB_KM_EPI == 0
      // Every epilogue warp visits every buffer (ring slot j % 4 for chunk j = it * n_chunks + c) and reads ONE
      // 32-column block of it: the round trip MMA -> commit -> epilogue -> release -> next MMA of a slot costs
      // ~1100 cycles of latency whatever the work (tools/ubench_mma.cu, mma_ring), so the time a buffer stays
      // with the epilogue has to be as short as possible -- 16 warps x 1 block, not 4 warps x 4 blocks.
      uint32_t j = it * (uint32_t)n_chunks;
      for (int c = 0; c < n_chunks; ++c, ++j) {
        const uint32_t tb = j & (kTcRing - 1);
        KM_T(e2);
        mbar_wait(&S->t_full[tb], (j >> 2) & 1u);
        KM_T(e3);
        KM_ACC(1, e2, e3);
        tc::fence_after_sync();
        {
          float v[32];
          tc::tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + tb * kTcChunk + (uint32_t)(grp * 32), v);
          tc::fence_before_sync();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&S->t_empty[tb]);     // the values are in registers: release the buffer now
          if constexpr (DBG) {
            const int64_t grow = tile * kTcTile + row;
            if (grow < p.n)
              for (int q = 0; q < 32; ++q)
                p.dbg_scores[grow * Kpad + c * kTcChunk + grp * 32 + q] = (v[q] + xn2) * inv_s2;
          }
          // screen: a block whose minimum is above the row's threshold cannot hold the best centre nor
          // one within the certainty margin of it (the finaliser caps `second` at the threshold)
          // 32 -> 1 with 3-input minima (FMNMX3): 10 + 4 + 1 + 1 instructions
          float t12[12];
#pragma unroll
          for (int q = 0; q < 10; ++q) t12[q] = fminf(fminf(v[3 * q], v[3 * q + 1]), v[3 * q + 2]);
          t12[10] = v[30];
          t12[11] = v[31];
          float t4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) t4[q] = fminf(fminf(t12[3 * q], t12[3 * q + 1]), t12[3 * q + 2]);
          const float bmin = fminf(fminf(fminf(t4[0], t4[1]), t4[2]), t4[3]);
          if (__any_sync(0xffffffffu, bmin <= thr)) {
            // packed keys: score bits with the position inside this 32-column block in the low 5 bits
            const int blk_before = best;
#pragma unroll
            for (int q = 0; q < 32; q += 2) {
              const int k0 = (int)((__float_as_uint(fmaxf(v[q] + xn2, 0.f)) & 0xFFFFFFE0u) | (uint32_t)q);
              const int k1 = (int)((__float_as_uint(fmaxf(v[q + 1] + xn2, 0.f)) & 0xFFFFFFE0u) | (uint32_t)(q + 1));
              const int lo = min(k0, k1), hi = max(k0, k1);
              const int t = max(best, lo);
              second = min(min(second, hi), t);
              best = min(best, lo);
            }
            if (best != blk_before) bblock = c * (kTcChunk / 32) + grp;   // low 5 key bits refer to this block
          }
        }
        KM_T(e4);
        KM_ACC(2, e3, e4);
      }
#elif PMB_KM_EPI == 2
      // Two super-groups of eight warps, each ping-ponging between TWO ring slots (s and s + 2: the chunks issued by
      // MMA warp s).  A warp reads two of the four 32-column blocks of its lane quarter per visit and the eight warps
      // release the slot after their second load, so the round trip release -> MMA -> commit of one slot (~550
      // cycles, fully exposed when a group owns a single slot) overlaps the two blocks read from the other slot.
      const uint32_t j0 = it * (uint32_t)n_chunks;
      const int sg = warp >> 3, half = (warp >> 2) & 1;
      for (int c = sg; c < n_chunks; c += 2) {
        const uint32_t j = j0 + (uint32_t)c, tb = j & (kTcRing - 1);
        KM_T(e2);
        mbar_wait(&S->t_full[tb], (j >> 2) & 1u);
        KM_T(e3);
        KM_ACC(1, e2, e3);
        tc::fence_after_sync();
        const uint32_t tbase = tmem + ((uint32_t)(quarter * 32) << 16) + tb * kTcChunk + (uint32_t)(half * 64);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float v[32];
          tc::tmem_ld32(tbase + (uint32_t)(hh * 32), v);
          if (hh == 1) {
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&S->t_empty[tb]);
          }
          if constexpr (DBG) {
            const int64_t grow = tile * kTcTile + row;
            if (grow < p.n)
              for (int q = 0; q < 32; ++q)
                p.dbg_scores[grow * Kpad + c * kTcChunk + (half * 2 + hh) * 32 + q] = (v[q] + xn2) * inv_s2;
          }
          float t12[12];
#pragma unroll
          for (int q = 0; q < 10; ++q) t12[q] = fminf(fminf(v[3 * q], v[3 * q + 1]), v[3 * q + 2]);
          t12[10] = v[30];
          t12[11] = v[31];
          float t4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) t4[q] = fminf(fminf(t12[3 * q], t12[3 * q + 1]), t12[3 * q + 2]);
          const float bmin = fminf(fminf(fminf(t4[0], t4[1]), t4[2]), t4[3]);
          if (!__any_sync(0xffffffffu, bmin <= thr)) continue;
          const int blk_before = best;
#pragma unroll
          for (int q = 0; q < 32; q += 2) {
            const int k0 = (int)((__float_as_uint(fmaxf(v[q] + xn2, 0.f)) & 0xFFFFFFE0u) | (uint32_t)q);
            const int k1 = (int)((__float_as_uint(fmaxf(v[q + 1] + xn2, 0.f)) & 0xFFFFFFE0u) | (uint32_t)(q + 1));
            const int lo = min(k0, k1), hi = max(k0, k1);
            const int t = max(best, lo);
            second = min(min(second, hi), t);
            best = min(best, lo);
          }
          if (best != blk_before) bblock = c * (kTcChunk / 32) + half * 2 + hh;
        }
        KM_T(e4);
        KM_ACC(2, e3, e4);
      }
#else
      // Group g owns ring slot g: chunk j = it * n_chunks + c lands in slot j % 4, so this group sees the chunks with
      // j % 4 == grp and reads all four 32-column blocks of its lane quarter.
      const uint32_t j0 = it * (uint32_t)n_chunks;
      for (int c = (int)((grp - j0) & 3u); c < n_chunks; c += kTcRing, ++uses) {
        KM_T(e2);
        mbar_wait(&S->t_full[grp], uses & 1u);
        KM_T(e3);
        KM_ACC(1, e2, e3);
        tc::fence_after_sync();
        const uint32_t tbase = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(grp * kTcChunk);
#pragma unroll
        for (int h = 0; h < kTcChunk / 32; ++h) {
          float v[32];
          tc::tmem_ld32(tbase + (uint32_t)(h * 32), v);
          if (h == kTcChunk / 32 - 1) {          // last block in registers: hand the buffer back before the arithmetic
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&S->t_empty[grp]);
          }
          if constexpr (DBG) {
            const int64_t grow = tile * kTcTile + row;
            if (grow < p.n)
              for (int q = 0; q < 32; ++q)
                p.dbg_scores[grow * Kpad + c * kTcChunk + h * 32 + q] = (v[q] + xn2) * inv_s2;
          }
          float t12[12];
#pragma unroll
          for (int q = 0; q < 10; ++q) t12[q] = fminf(fminf(v[3 * q], v[3 * q + 1]), v[3 * q + 2]);
          t12[10] = v[30];
          t12[11] = v[31];
          float t4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) t4[q] = fminf(fminf(t12[3 * q], t12[3 * q + 1]), t12[3 * q + 2]);
          const float bmin = fminf(fminf(fminf(t4[0], t4[1]), t4[2]), t4[3]);
          if (!__any_sync(0xffffffffu, bmin <= thr)) continue;
          const int blk_before = best;
#pragma unroll
          for (int q = 0; q < 32; q += 2) {
            const int k0 = (int)((__float_as_uint(fmaxf(v[q] + xn2, 0.f)) & 0xFFFFFFE0u) | (uint32_t)q);
            const int k1 = (int)((__float_as_uint(fmaxf(v[q + 1] + xn2, 0.f)) & 0xFFFFFFE0u) | (uint32_t)(q + 1));
            const int lo = min(k0, k1), hi = max(k0, k1);
            const int t = max(best, lo);
            second = min(min(second, hi), t);
            best = min(best, lo);
          }
          if (best != blk_before) bblock = c * (kTcChunk / 32) + h;
        }
        KM_T(e4);
        KM_ACC(2, e3, e4);
      }
#endif
      const uint32_t rb = it & 1u;
      KM_T(e5);
      mbar_wait(&S->r_empty[rb], ((it >> 1) & 1u) ^ 1u);
      KM_T(e6);
      KM_ACC(3, e5, e6);
      S->res_best[rb][grp][row] = best;
      S->res_second[rb][grp][row] = second;
      S->res_block[rb][grp][row] = bblock;
      if (grp == 0) {
        S->res_thr[rb][row] = thr + xn2;   // back in score space, like the keys
        S->res_xn2[rb][row] = xn2;
      }
      tc::mbar_arrive(&S->r_full[rb]);
    }
    if (warp == 0) { KM_FLUSH(0, 4); }
  } else if (warp == kTcMmaWarp || (warp == kTcLoadWarp && resident)) {
    // =========================================================== MMA issuer(s)
    // This warp's own instruction stream is the pace-maker of the tensor pipe: a lone warp retires a dependent
    // scalar instruction every ~5 cycles, so descriptor arithmetic, integer divisions or one elect per MMA in
    // this loop cost more than the MMAs themselves (tools/ubench_mma.cu: 708 cycles per 2-MMA chunk with a
    // naive loop, 175 with the loop below).  Everything is precomputed; one elect covers a chunk's MMAs and
    // its commit.
    const uint32_t idesc = idesc_f16(kTcTile, kTcChunk);
    const uint32_t aB = smem_u32(sB), aA = smem_u32(sA);
    const int nks = KS / 16;
    const uint64_t dA0 = tc::smem_desc(aA, lbo, sbo, tc::kLayoutNone);
    const uint64_t dA_other = (uint64_t)(chunk_bytes >> 4);     // address-field distance between the two A tiles
    const uint64_t dB0 = tc::smem_desc(aB, lbo, sbo, tc::kLayoutNone);
    const uint64_t dB_slot = (uint64_t)(chunk_bytes >> 4);      // ... from one ring slot of the centre operand to the next
    const uint64_t kstep = (uint64_t)((2u * lbo) >> 4);         // ... and from one K = 16 step to the next
    const uint32_t t_full0 = smem_u32(&S->t_full[0]), t_empty0 = smem_u32(&S->t_empty[0]);
    uint32_t it = 0;
    KM_DECL;
    KM_T(m_begin);
    // one chunk: nks MMAs into TMEM buffer tb and the commit that hands the buffer to the epilogue
    auto issue_chunk = [&](uint64_t da, uint64_t db, uint32_t tb) {
      const uint32_t d_tmem = tmem + tb * kTcChunk;
      const uint32_t bar = t_full0 + tb * 8u;
      if (nks == 2) {
        asm volatile(
            "{\n\t"
            ".reg .pred q;\n\t"
            "elect.sync _|q, 0xffffffff;\n\t"
            "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %5, 0;\n\t"
            "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %3, %4, %5, 1;\n\t"
            "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n\t"
            "}" ::"r"(d_tmem),
            "l"(da), "l"(db), "l"(da + kstep), "l"(db + kstep), "r"(idesc), "r"(bar)
            : "memory");
      } else if (nks == 1) {
        asm volatile(
            "{\n\t"
            ".reg .pred q;\n\t"
            "elect.sync _|q, 0xffffffff;\n\t"
            "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 0;\n\t"
            "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%4];\n\t"
            "}" ::"r"(d_tmem),
            "l"(da), "l"(db), "r"(idesc), "r"(bar)
            : "memory");
      } else {
        uint64_t a = da, b = db;
#pragma unroll 1
        for (int s = 0; s < nks; ++s, a += kstep, b += kstep) mma_f16_elect(d_tmem, a, b, idesc, s > 0 ? 1u : 0u);
        tc::mma_commit_elect(&S->t_full[tb]);
      }
    };
    // wait until the epilogue has released ring slot tb (phase parity par)
    auto wait_slot = [&](uint32_t tb, uint32_t par) {
      const uint32_t addr = t_empty0 + tb * 8u;
#pragma unroll 1
      for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(addr), "r"(par)
            : "memory");
        if (ok) return;
      }
      __trap();
    };
    if (resident) {
      // TWO issuing warps (this one and the otherwise idle loader warp): the pace of a lone issuer is its own
      // serial instruction stream (~350 cycles per chunk with the waits), twice what the epilogue needs.  Warp m
      // issues the chunks c = m (mod 2); the number of chunks per tile is even, so they live in ring slots
      // m and m + 2.  The second warp first launches the one-time bulk copies of the centre chunks.
      const uint32_t m = (uint32_t)(warp - kTcMmaWarp);
      if (m == 1u && lane == 0) {
        for (int c = 0; c < n_chunks; ++c) {
          mbar_expect_tx(&S->b_full[c], chunk_bytes);
          bulk_g2s(sB + (size_t)c * chunk_bytes, p.Bg + (size_t)c * chunk_bytes, chunk_bytes, &S->b_full[c]);
        }
      }
      __syncwarp();
      for (int c = 0; c < n_chunks; ++c) mbar_wait(&S->b_full[c], 0u);
      uint32_t tb = m, par = 1u;   // ring slot and the parity its t_empty barrier is waited with
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t ab = it & 1u;
        KM_T(m0);
        mbar_wait(&S->a_full[ab], (it >> 1) & 1u);
        KM_T(m1);
        KM_ACC(0, m0, m1);
        tc::fence_after_sync();
        if (m == 0u && lane == 0) tc::mbar_arrive(&S->thr_ready[it & 3u]);   // producers' thr[] -> epilogue warps
        const uint64_t da = dA0 + (ab ? dA_other : 0ull);
        uint64_t db = dB0 + (uint64_t)m * dB_slot;
#pragma unroll 1
        for (int c = (int)m; c < n_chunks; c += 2, db += 2 * dB_slot) {
          KM_T(m2);
          wait_slot(tb, par);
          KM_T(m3);
          KM_ACC(1, m2, m3);
          tc::fence_after_sync();
          issue_chunk(da, db, tb);
          tb ^= 2u;
          par ^= (tb == m) ? 1u : 0u;
        }
        tc::mma_commit_elect(&S->a_empty[ab]);      // one arrival per issuing warp
      }
    } else {
      uint32_t tb = 0, par = 1u;
      uint32_t bs = 0, bfill = 0;   // streamed centre chunks: shared-memory ring slot and fill parity
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t ab = it & 1u;
        KM_T(m0);
        mbar_wait(&S->a_full[ab], (it >> 1) & 1u);
        KM_T(m1);
        KM_ACC(0, m0, m1);
        tc::fence_after_sync();
        if (lane == 0) tc::mbar_arrive(&S->thr_ready[it & 3u]);
        const uint64_t da = dA0 + (ab ? dA_other : 0ull);
#pragma unroll 1
        for (int c = 0; c < n_chunks; ++c) {
          mbar_wait(&S->b_full[bs], bfill);
          KM_T(m2);
          wait_slot(tb, par);
          KM_T(m3);
          KM_ACC(1, m2, m3);
          tc::fence_after_sync();
          issue_chunk(da, dB0 + (uint64_t)bs * dB_slot, tb);
          tc::mma_commit_elect(&S->b_empty[bs]);
          if (++bs == (uint32_t)nslots) { bs = 0; bfill ^= 1u; }
          tb = (tb + 1u) & (kTcRing - 1);
          par ^= (tb == 0u) ? 1u : 0u;
        }
        tc::mma_commit_elect(&S->a_empty[ab]);
      }
    }
    KM_T(m_end);
    KM_ACC(2, m_begin, m_end);
#ifdef PMB_KM_PROF
    km_prof[3] = (long long)it;
#endif
    if (warp == kTcMmaWarp) { KM_FLUSH(4, 4); }
  } else if (warp == kTcLoadWarp) {
    // =========================================================== loader (streamed centre chunks only) -> shared-memory ring
    if (lane == 0) {
      {
        uint32_t bs = 0, bfill = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
          for (int c = 0; c < n_chunks; ++c) {
            mbar_wait(&S->b_empty[bs], bfill ^ 1u);
            mbar_expect_tx(&S->b_full[bs], chunk_bytes);
            bulk_g2s(sB + (size_t)bs * chunk_bytes, p.Bg + (size_t)c * chunk_bytes, chunk_bytes, &S->b_full[bs]);
            if (++bs == (uint32_t)nslots) { bs = 0; bfill ^= 1u; }
          }
        }
      }
    }
  } else if (warp < kTcFinWarp0) {
    // =========================================================== producers: A tile + screening threshold
    const int pt = tid - kTcProdWarp0 * 32;   // 0..127: row of the tile
    constexpr bool in_regs = INREG;
    // Two levels of prefetch: the hint of tile t + 2 is requested while tile t is built, so that the gather of
    // the HINTED centre of tile t + 1 (an address that depends on the hint) is issued a whole tile ahead as
    // well.  With the gather issued at use, every tile paid its L2 round trip on the producers' serial path
    // (1 700 of their 3 400 cycles per tile).
    float yn[DRP];    // coordinates of this thread's row of the NEXT tile (scaled at use)
    float chn[DRP];   // scaled hinted centre of that row (INREG)
    int hint_n = -1, hint_nn = -1;   // hints of the next tile and of the one after
    auto load_hint = [&](int64_t tile) -> int {
      const int64_t row = tile * kTcTile + pt;
      return (tile < n_tiles && row < p.n && p.hints != nullptr) ? ldg_stream_i(p.hints + row) : -1;
    };
    auto prefetch = [&](int64_t tile) {   // hint_n holds this tile's hint
      const int64_t row = tile * kTcTile + pt;
      if (row < p.n) {
        if constexpr (in_regs) {
#pragma unroll
          for (int d = 0; d < DRP; ++d) yn[d] = (d < D) ? ldg_stream_f(p.Y + row * p.ld + d) : 0.f;   // scaled at use
          if (hint_n >= 0 && hint_n < K) {
            const float* ch = p.cs32 + (size_t)hint_n * D;
#pragma unroll
            for (int d = 0; d < DRP; ++d) chn[d] = (d < D) ? __ldg(ch + d) : 0.f;
          }
        }
      } else {
#pragma unroll
        for (int d = 0; d < DRP; ++d) yn[d] = 0.f;
      }
    };
    uint32_t it = 0;
    double ysq_acc = 0.0;
    const double inv_s2d = (double)p.meta[2];
    KM_DECL;
    hint_n = load_hint(blockIdx.x);
    hint_nn = load_hint((int64_t)blockIdx.x + gridDim.x);
    if ((int64_t)blockIdx.x < n_tiles) prefetch(blockIdx.x);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ab = it & 1u;
      const int64_t row = tile * kTcTile + pt;
      const bool valid = row < p.n;
      float y[DRP];
#pragma unroll
      for (int d = 0; d < DRP; ++d) y[d] = yn[d] * scale;
      const int hint = hint_n;
      // fp32 distance to the hinted centre (scaled centres written by the prep kernel)
      float u = __int_as_float(0x7f800000);
      if (valid && hint >= 0 && hint < K) {
        u = 0.f;
        if constexpr (in_regs) {
#pragma unroll
          for (int d = 0; d < DRP; ++d) {
            if (d < D) {
              const float t = y[d] - chn[d];
              u = fmaf(t, t, u);
            }
          }
        } else {
          const float* ch = p.cs32 + (size_t)hint * D;
#pragma unroll 4
          for (int d = 0; d < D; ++d) {
            const float t = p.Y[row * p.ld + d] * scale - ch[d];
            u = fmaf(t, t, u);
          }
        }
      }
      hint_n = hint_nn;
      if (tile + gridDim.x < n_tiles) prefetch(tile + gridDim.x);
      hint_nn = load_hint(tile + 2 * (int64_t)gridDim.x);
      KM_T(p0);
      mbar_wait(&S->a_empty[ab], ((it >> 1) & 1u) ^ 1u);
      KM_T(p1);
      KM_ACC(0, p0, p1);
      unsigned char* A = sA + (size_t)ab * chunk_bytes;
      float xn2 = 0.f;
      bool fits = true;   // every scaled coordinate is finite and inside fp16's range
      if constexpr (in_regs) {
        double rowsq = 0.0;   // |y s|^2 in fp64: the inertia identity cancels, fp32 would not do
#pragma unroll
        for (int d = 0; d < DRP; ++d) {
          const float v = y[d];          // zero beyond D and for rows past the end
          xn2 = fmaf(v, v, xn2);
          rowsq = fma((double)v, (double)v, rowsq);
          fits = fits && (fabsf(v) < kTcHalfMax);
        }
        if (!fits) {
#pragma unroll
          for (int d = 0; d < DRP; ++d) y[d] = 0.f;
        }
        const float one = valid ? kTcNormUnit : 0.0f;
        // slot s: 0-1 -> 2^13, 2 + 3 d + r -> (r == 1 ? y_lo[d] : y_hi[d]); 16-byte stores (8 slots), 8 lanes
        // cover one 128-byte core-matrix row: conflict-free
        auto slot_val = [&](int sl) -> float {
          if (sl < 2) return one;
          const int d = (sl - 2) / 3, r = (sl - 2) % 3;
          if (d >= DRP) return 0.f;
          const float hi = __half2float(__float2half_rn(y[d]));   // recomputed per use: keeps the register count low
          return r == 1 ? (y[d] - hi) : hi;
        };
        unsigned char* arow = A + (uint32_t)(pt >> 3) * sbo + (uint32_t)(pt & 7) * 16u;
#pragma unroll
        for (int j = 0; j < (3 * DRP + 2 + 15) / 16 * 2; ++j) {
          if (8 * j < KS) {
            __half2 h0 = __floats2half2_rn(slot_val(8 * j), slot_val(8 * j + 1));
            __half2 h1 = __floats2half2_rn(slot_val(8 * j + 2), slot_val(8 * j + 3));
            __half2 h2 = __floats2half2_rn(slot_val(8 * j + 4), slot_val(8 * j + 5));
            __half2 h3 = __floats2half2_rn(slot_val(8 * j + 6), slot_val(8 * j + 7));
            uint4 w;
            w.x = *reinterpret_cast<uint32_t*>(&h0);
            w.y = *reinterpret_cast<uint32_t*>(&h1);
            w.z = *reinterpret_cast<uint32_t*>(&h2);
            w.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(arow + (uint32_t)j * lbo) = w;
          }
        }
        if (valid && fits) ysq_acc = fma(rowsq, inv_s2d, ysq_acc);   // s is a power of two: exact
      } else {
        auto put = [&](int slot, float v) {
          *reinterpret_cast<__half*>(A + off_kmajor_h(pt, slot, sbo)) = __float2half_rn(v);
        };
#pragma unroll 4
        for (int d = 0; d < D; ++d) {
          const float v = valid ? p.Y[row * p.ld + d] * scale : 0.f;
          xn2 = fmaf(v, v, xn2);
          fits = fits && (fabsf(v) < kTcHalfMax);
        }
        double acc = 0.0;
#pragma unroll 4
        for (int d = 0; d < D; ++d) {
          const float raw = valid ? p.Y[row * p.ld + d] : 0.f;
          acc = fma((double)raw, (double)raw, acc);
          const float v = fits ? raw * scale : 0.f;
          const float hi = __half2float(__float2half_rn(v));
          put(2 + 3 * d + 0, hi);
          put(2 + 3 * d + 1, v - hi);
          put(2 + 3 * d + 2, hi);
        }
        if (valid && fits) ysq_acc += acc;
        const float one = valid ? kTcNormUnit : 0.0f;
        put(0, one);
        put(1, one);
        for (int sl = 3 * D + 2; sl < KS; ++sl) put(sl, 0.f);
      }
      // A frame that does not fit is decided by the fp64 re-check: NaN |y|^2 fails every certainty test.
      if (!fits) xn2 = __int_as_float(0x7fc00000);
      // screening threshold: the fp32 distance to the hinted centre bounds the best score from above
      // (up to the envelope E); the margin keeps the certainty test decidable.  TMEM holds score - |y|^2, so the
      // epilogue compares against thr - |y|^2 and adds |y|^2 back only on the rare slow path
      const float rr = sqrtf(xn2) * 1.0001f + cmax;
      S->thr[it & 3u][pt] = (u * (1.0f + 1.5258789e-5f) + 4.0f * kTcErrScale * rr * rr) - xn2;
      S->xn2[it & 3u][pt] = xn2;
      fence_proxy_async_smem();
      tc::mbar_arrive(&S->a_full[ab]);
      KM_T(p2);
      KM_ACC(1, p1, p2);
    }
    if (warp == kTcProdWarp0) { KM_FLUSH(8, 2); }
    if (p.accumulate) {
      ysq_acc = warp_sum(ysq_acc);
      if (lane == 0 && ysq_acc != 0.0) atomicAdd(p.ysq, ysq_acc);
    }
  } else {
    // =========================================================== finalisers: merge, certainty test, labels,
    // fused Lloyd accumulation
    const int pt = tid - kTcFinWarp0 * 32;
    constexpr bool in_regs = INREG;
    double* const my_sums = p.lsums + (size_t)(blockIdx.x % p.ncopies) * p.copy_stride;
    unsigned long long* const my_counts = p.lcounts + (size_t)(blockIdx.x % p.ncopies) * p.copy_stride;
    int recheck_acc = 0;
    KM_DECL;
    float yn[kTcDReg];   // UNSCALED coordinates (they feed the Lloyd sums)
    auto prefetch = [&](int64_t tile) {
      const int64_t row = tile * kTcTile + pt;
#pragma unroll
      for (int d = 0; d < kTcDReg; ++d) yn[d] = (in_regs && d < D && row < p.n) ? ldg_stream_f(p.Y + row * p.ld + d) : 0.f;
    };
    uint32_t it = 0;
    double ysq_fix = 0.0;   // |y|^2 of the frames the producers left out of ysq (they did not fit fp16)
    if ((int64_t)blockIdx.x < n_tiles) prefetch(blockIdx.x);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      float y[kTcDReg];
#pragma unroll
      for (int d = 0; d < kTcDReg; ++d) y[d] = yn[d];
      if (tile + gridDim.x < n_tiles) prefetch(tile + gridDim.x);
      const uint32_t rb = it & 1u;
      KM_T(f0);
      mbar_wait(&S->r_full[rb], (it >> 1) & 1u);
      KM_T(f1);
      KM_ACC(0, f0, f1);
      int b = 0x7fffffff, s2 = 0x7fffffff, bb = 0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int gb = S->res_best[rb][g][pt], gs = S->res_second[rb][g][pt], gk = S->res_block[rb][g][pt];
        const int t = max(b, gb);           // merge: new second = min(s2, gs, max(b, gb))
        s2 = min(min(s2, gs), t);
        if (gb < b) { b = gb; bb = gk; }
      }
      const float thr = S->res_thr[rb][pt];
      const float xn2s = S->res_xn2[rb][pt];   // NaN: the frame did not fit fp16
      tc::mbar_arrive(&S->r_empty[rb]);
      const int64_t row = tile * kTcTile + pt;
      const bool valid = row < p.n;
      int lab = -1;
      if (valid) {
        int k = bb * 32 + (b & 31);
        const bool in_range = k < K;   // a padding row can only win when no real centre was scored (b == INT_MAX)
        if (!in_range) k = K - 1;
        const float s1f = __uint_as_float((uint32_t)b & 0xFFFFFFE0u);
        const float s2f = fminf(__uint_as_float((uint32_t)s2 & 0xFFFFFFE0u), thr);
        const float rr = sqrtf(xn2s) * 1.0001f + cmax;
        const float E = kTcErrScale * rr * rr;
        const bool certain = in_range && (b >= 0) && (b != 0x7fffffff) && (K == 1 || (s1f * kTcKeyTrunc + E < s2f - E));
        p.labels[row] = k;
        if (certain) {
          lab = k;
        } else {
          ++recheck_acc;
          const int pos = atomicAdd(p.recheck_count, 1);
          p.recheck_list[pos] = (int)row;
        }
        if (!(xn2s == xn2s) && p.accumulate) {
          if constexpr (in_regs) {
#pragma unroll
            for (int d = 0; d < kTcDReg; ++d) ysq_fix = fma((double)y[d], (double)y[d], ysq_fix);
          } else {
            for (int d = 0; d < D; ++d) {
              const double v = (double)p.Y[row * p.ld + d];
              ysq_fix = fma(v, v, ysq_fix);
            }
          }
        }
      }
      // fused Lloyd accumulation into this call's private sums / counts.  32 consecutive frames of a
      // trajectory share one to three labels (bench data: 67 % of the warps one label, two runs on average), so
      // the warp is split into its label groups: per group one reduction through shared memory (lane (d, half)
      // adds the group's rows of coordinate d in its half in fp64) and D + 1 atomics, instead of one atomic
      // per frame and coordinate whenever the warp was not uniform.
      if (p.accumulate) {
        if constexpr (in_regs) {
          float* stg = stage + (size_t)(warp - kTcFinWarp0) * 32 * (kTcDReg + 1);
#pragma unroll
          for (int d = 0; d < kTcDReg; ++d) stg[lane * (kTcDReg + 1) + d] = y[d];
          __syncwarp();
          const int dd = lane & 15, r0 = (lane >> 4) * 16;
          unsigned todo = __ballot_sync(0xffffffffu, lab >= 0);
          while (todo != 0u) {
            const int lab0 = __shfl_sync(0xffffffffu, lab, __ffs(todo) - 1);
            const unsigned grp = __ballot_sync(0xffffffffu, lab == lab0);
            const unsigned mine = grp >> r0;
            double acc = 0.0;
#pragma unroll
            for (int r = 0; r < 16; ++r)
              if ((mine >> r) & 1u) acc += (double)stg[(r0 + r) * (kTcDReg + 1) + dd];
            acc += __shfl_xor_sync(0xffffffffu, acc, 16);
            if (lane < D) atomicAdd(my_sums + (size_t)lab0 * D + lane, acc);
            if (lane == 0) atomicAdd(my_counts + lab0, (unsigned long long)__popc(grp));
            todo &= ~grp;
          }
          __syncwarp();
        } else {
          const int lab0 = __shfl_sync(0xffffffffu, lab, 0);
          if (__all_sync(0xffffffffu, lab == lab0)) {
            if (lab0 >= 0) {
              for (int d = 0; d < D; ++d) {
                const double v = warp_sum((double)p.Y[row * p.ld + d]);
                if (lane == 0) atomicAdd(my_sums + (size_t)lab0 * D + d, v);
              }
              if (lane == 0) atomicAdd(my_counts + lab0, 32ull);
            }
          } else if (lab >= 0) {
            for (int d = 0; d < D; ++d) atomicAdd(my_sums + (size_t)lab * D + d, (double)p.Y[row * p.ld + d]);
            atomicAdd(my_counts + lab, 1ull);
          }
        }
      }
      KM_T(f2);
      KM_ACC(1, f1, f2);
    }
    if (warp == kTcFinWarp0) { KM_FLUSH(10, 2); }
    if (p.accumulate) {
      ysq_fix = warp_sum(ysq_fix);
      if (lane == 0 && ysq_fix != 0.0) atomicAdd(p.ysq, ysq_fix);
    }
    if (p.n_rechecked != nullptr) {
      int r = recheck_acc;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
      if (lane == 0 && r) atomicAdd(reinterpret_cast<unsigned long long*>(p.n_rechecked), (unsigned long long)r);
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == kTcMmaWarp) tc::tmem_dealloc(tmem, 512);
}

// One warp per listed frame: fp64 direct-difference distances in the oracle's operation order,
// first minimum wins; the frame then joins the Lloyd accumulation.
__global__ void __launch_bounds__(256) kmeans_recheck_kernel(KmTcParams p) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int count = *p.recheck_count;
  const int D = p.D, K = p.K;
  for (int e = gw; e < count; e += nw) {
    const int64_t row = p.recheck_list[e];
    // bk starts at 0 like the SIMT kernel and np.argmin: a NaN / Inf row (every comparison false) is
    // labelled 0 instead of indexing the accumulators out of bounds
    double bd = 1.7976931348623157e308;
    int bk = 0;
    if (D <= kTcDReg) {
      // the row in registers and four centres per lane in flight: the per-centre sum keeps the oracle's order
      // (sub, mul, add, no fma) and the centres are still visited in increasing k, so labels are unchanged; the
      // plain loop below is a chain of ~32 x D dependent operations per lane (45 us per launch)
      double yr[kTcDReg];
#pragma unroll
      for (int d = 0; d < kTcDReg; ++d) yr[d] = d < D ? (double)p.Y[row * p.ld + d] : 0.0;
      for (int k0 = lane; k0 < K; k0 += 128) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int d = 0; d < kTcDReg; ++d) {
          if (d < D) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int k = k0 + 32 * u;
              const double cv = k < K ? p.centers[(size_t)k * D + d] : 0.0;
              const double t = __dsub_rn(yr[d], cv);
              acc[u] = __dadd_rn(acc[u], __dmul_rn(t, t));
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = k0 + 32 * u;
          if (k < K && acc[u] < bd) { bd = acc[u]; bk = k; }
        }
      }
    } else
    for (int k = lane; k < K; k += 32) {
      double acc = 0.0;
      const double* c = p.centers + (size_t)k * D;
      for (int d = 0; d < D; ++d) {
        const double t = __dsub_rn((double)p.Y[row * p.ld + d], c[d]);
        acc = __dadd_rn(acc, __dmul_rn(t, t));
      }
      if (acc < bd) { bd = acc; bk = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, bd, o);
      const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
      if (od < bd || (od == bd && ok < bk)) { bd = od; bk = ok; }
    }
    if (lane == 0) {
      p.labels[row] = bk;
      if (p.accumulate) atomicAdd(p.lcounts + bk, 1ull);
    }
    if (p.accumulate)
      for (int d = lane; d < D; d += 32) atomicAdd(p.lsums + (size_t)bk * D + d, (double)p.Y[row * p.ld + d]);
  }
}

// Fold the private accumulator copies into copy 0, in a fixed order (one thread per accumulator word; adjacent
// threads read adjacent words of every copy).
__global__ void __launch_bounds__(256) kmeans_tc_fold_kernel(KmTcParams p) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n_sum = (size_t)p.K * p.D, n_all = n_sum + (size_t)p.K;
  if (e >= n_all) return;
  // four partial sums in a fixed pattern, sixteen loads in flight: a single running sum walked the copies as a
  // chain of dependent L2 round trips (36 us per launch)
  if (e < n_sum) {
    const double* src = p.lsums + e;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
    int c = 0;
#pragma unroll 4
    for (; c + 4 <= p.ncopies; c += 4) {
      v0 += __ldcg(src + (size_t)c * p.copy_stride);
      v1 += __ldcg(src + (size_t)(c + 1) * p.copy_stride);
      v2 += __ldcg(src + (size_t)(c + 2) * p.copy_stride);
      v3 += __ldcg(src + (size_t)(c + 3) * p.copy_stride);
    }
    for (; c < p.ncopies; ++c) v0 += __ldcg(src + (size_t)c * p.copy_stride);
    p.lsums[e] = (v0 + v1) + (v2 + v3);
  } else {
    const unsigned long long* src = p.lcounts + (e - n_sum);
    unsigned long long v0 = 0, v1 = 0, v2 = 0, v3 = 0;
    int c = 0;
#pragma unroll 4
    for (; c + 4 <= p.ncopies; c += 4) {
      v0 += __ldcg(src + (size_t)c * p.copy_stride);
      v1 += __ldcg(src + (size_t)(c + 1) * p.copy_stride);
      v2 += __ldcg(src + (size_t)(c + 2) * p.copy_stride);
      v3 += __ldcg(src + (size_t)(c + 3) * p.copy_stride);
    }
    for (; c < p.ncopies; ++c) v0 += __ldcg(src + (size_t)c * p.copy_stride);
    p.lcounts[e - n_sum] = (v0 + v1) + (v2 + v3);
  }
}

// Hand this call's sums / counts to the caller and form the inertia from them:
//   sum_t |y_t - c_l(t)|^2 = sum_t |y_t|^2 + sum_k (n_k |c_k|^2 - 2 c_k . S_k)       (all in fp64)
__global__ void __launch_bounds__(256) kmeans_tc_commit_kernel(KmTcParams p) {
  __shared__ double s_red[8];
  const int D = p.D, K = p.K;
  double part = 0.0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) {
    const unsigned long long nk = p.lcounts[k];
    if (nk == 0) continue;
    double n2 = 0.0, cs = 0.0;
    for (int d = 0; d < D; ++d) {
      const double c = p.centers[(size_t)k * D + d], sv = p.lsums[(size_t)k * D + d];
      n2 = fma(c, c, n2);
      cs = fma(c, sv, cs);
      if (p.sums != nullptr) p.sums[(size_t)k * D + d] += sv;
    }
    if (p.counts != nullptr) p.counts[k] += (int64_t)nk;
    part += (double)nk * n2 - 2.0 * cs;
  }
  if (p.inertia != nullptr) {
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += s_red[w];
      if (blockIdx.x == 0) t += *p.ysq;
      atomicAdd(p.inertia, t);
    }
  }
}

int kmeans_tc_debug_counters(int64_t* out16) {
  long long h[16];
  PMB_CUDA(cudaMemcpyFromSymbol(h, g_km_dbg, sizeof(h)));
  for (int i = 0; i < 16; ++i) out16[i] = h[i];
  return PMB_OK;
}

static inline int tc_slots(int D) { return ((3 * D + 2) + 15) / 16 * 16; }
static inline int tc_kpad(int K) { return (K + kTcKpadTo - 1) / kTcKpadTo * kTcKpadTo; }
static inline size_t tc_chunk_bytes(int D) { return (size_t)kTcChunk * tc_slots(D) * 2; }
static inline size_t tc_fixed_smem(int D) {
  // two A tiles + bookkeeping + the finalisers' transpose buffer (register path only)
  return 2 * tc_chunk_bytes(D) + sizeof(TcSmem) + (D <= kTcDReg ? (size_t)kTcFinWarps * 32 * (kTcDReg + 1) * 4 : 0) + 16;
}
// ring slots for centre chunks: all of them when they fit (resident), else as many as fit (streamed, >= 2)
static inline int tc_ring_slots(int D, int K) {
  const size_t budget = 227 * 1024, fixed = tc_fixed_smem(D);
  if (fixed >= budget) return 0;
  const int n_chunks = tc_kpad(K) / kTcChunk;
  int fit = (int)((budget - fixed) / tc_chunk_bytes(D));
  if (fit > kTcMaxSlots) fit = kTcMaxSlots;
  if (n_chunks <= fit) return n_chunks;
  return fit >= 2 ? fit : 0;
}

bool kmeans_tc_supported(int D, int K) {
  return D >= 1 && K >= 1 && tc_slots(D) * 2 / 16 < (1 << 14) && tc_ring_slots(D, K) > 0;
}

static inline size_t tc_align256(size_t x) { return (x + 255) / 256 * 256; }
// private accumulator copies: one per CTA while they stay below 16 MB in total (zeroed every call)
static inline size_t tc_copy_bytes(int D, int K) { return tc_align256(((size_t)K * D + K) * sizeof(double)); }
static inline int tc_ncopies(int D, int K) {
  const size_t fit = ((size_t)16 << 20) / tc_copy_bytes(D, K);
  return fit >= (size_t)kNumSMs ? kNumSMs : (fit < 1 ? 1 : (int)fit);
}

size_t kmeans_tc_ws_bytes(int64_t n, int D, int K) {
  // [count | ysq | meta | pad to 256 B][ncopies x (local sums K x D, local counts K)][scaled centres K x D fp32]
  // [centre operand image Kpad x KS fp16][re-check list n]
  return 256 + (size_t)tc_ncopies(D, K) * tc_copy_bytes(D, K) + tc_align256((size_t)K * D * sizeof(float)) +
         tc_align256((size_t)tc_kpad(K) * tc_slots(D) * 2) + (size_t)n * sizeof(int) + 256;
}

int kmeans_tc_assign(const float* Y, int64_t n, int D, int64_t ld, const double* centers, int K, int32_t* labels,
                     const int32_t* hints, double* sums, int64_t* counts, double* inertia, int64_t* n_rechecked, void* ws,
                     float* dbg_scores, cudaStream_t st) {
  KmTcParams p;
  p.Y = Y; p.n = n; p.D = D; p.ld = ld; p.centers = centers; p.K = K;
  p.Kpad = tc_kpad(K); p.KS = tc_slots(D); p.nslots = tc_ring_slots(D, K);
  p.labels = labels; p.hints = hints; p.sums = sums; p.counts = counts; p.inertia = inertia; p.n_rechecked = n_rechecked;
  PMB_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "pmb_kmeans_assign: workspace must be 256-byte aligned");
  unsigned char* w8 = static_cast<unsigned char*>(ws);
  p.recheck_count = reinterpret_cast<int*>(w8);
  p.ysq = reinterpret_cast<double*>(w8 + 16);
  p.meta = reinterpret_cast<float*>(w8 + 64);
  size_t off = 256;
  p.lsums = reinterpret_cast<double*>(w8 + off);
  p.lcounts = reinterpret_cast<unsigned long long*>(p.lsums + (size_t)K * D);
  p.ncopies = tc_ncopies(D, K);
  p.copy_stride = tc_copy_bytes(D, K) / 8;
  off += (size_t)p.ncopies * tc_copy_bytes(D, K);
  p.cs32 = reinterpret_cast<float*>(w8 + off);
  off += tc_align256((size_t)K * D * sizeof(float));
  p.Bg = w8 + off;
  off += tc_align256((size_t)p.Kpad * p.KS * 2);
  p.recheck_list = reinterpret_cast<int*>(w8 + off);
  p.accumulate = (sums != nullptr || inertia != nullptr) ? 1 : 0;
  p.dbg_scores = dbg_scores;
  PMB_REQUIRE(n < (int64_t)0x7fffffff, "pmb_kmeans_assign: tensor path needs n < 2^31");
  PMB_REQUIRE(p.nslots > 0, "pmb_kmeans_assign: D=%d does not fit the tensor path's shared memory", D);
  const size_t smem = (size_t)p.nslots * tc_chunk_bytes(D) + tc_fixed_smem(D);
  PMB_CUDA(cudaMemsetAsync(ws, 0, 256 + (p.accumulate ? (size_t)p.ncopies * tc_copy_bytes(D, K) : 0), st));
  kmeans_tc_prep_kernel<<<1, 1024, 0, st>>>(p);
  PMB_LAUNCH_CHECK();
  const int64_t n_tiles = (n + kTcTile - 1) / kTcTile;
  const int grid = (int)(n_tiles < kNumSMs ? n_tiles : kNumSMs);
  auto launch = [&](auto kern) -> int {
    PMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kTcThreads, smem, st>>>(p);
    PMB_LAUNCH_CHECK();
    return PMB_OK;
  };
  int rc;
  if (dbg_scores != nullptr)
    rc = (D <= 12) ? launch(kmeans_tc_kernel<true, true, 12>)
                   : (D <= kTcDReg) ? launch(kmeans_tc_kernel<true, true, kTcDReg>) : launch(kmeans_tc_kernel<true, false, kTcDReg>);
  else
    rc = (D <= 12) ? launch(kmeans_tc_kernel<false, true, 12>)
                   : (D <= kTcDReg) ? launch(kmeans_tc_kernel<false, true, kTcDReg>) : launch(kmeans_tc_kernel<false, false, kTcDReg>);
  if (rc != PMB_OK) return rc;
  kmeans_recheck_kernel<<<2 * kNumSMs, 256, 0, st>>>(p);
  PMB_LAUNCH_CHECK();
  if (p.accumulate) {
    if (p.ncopies > 1) {
      const size_t n_all = (size_t)K * D + K;
      kmeans_tc_fold_kernel<<<(unsigned)((n_all + 255) / 256), 256, 0, st>>>(p);
      PMB_LAUNCH_CHECK();
    }
    kmeans_tc_commit_kernel<<<(K + 255) / 256 < 64 ? (K + 255) / 256 : 64, 256, 0, st>>>(p);
    PMB_LAUNCH_CHECK();
  }
  return PMB_OK;
}

}  // namespace pmb
