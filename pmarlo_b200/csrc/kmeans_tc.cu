// kmeans_tc.cu -- K6 (tensor-core path): nearest-centre assignment as a tcgen05 TF32 GEMM with the
// argmin fused into the TMEM epilogue, labels bit-exact against the fp64 oracle.
//
//   score(t,k) - |y_t|^2 = |c_k|^2 - 2 y_t.c_k            (one 128 x 256 x KS MMA tile per chunk)
//
// The product is made fp32-accurate on TF32 tensor cores by splitting along the MMA K dimension
// ("3xTF32 in K"): every coordinate d contributes three K slots
//      A (frame side):  y_hi  y_lo  y_hi        B (centre side):  c'_hi  c'_hi  c'_lo     (c' = -2 c)
// and |c|^2 rides in two more slots (two TF32 pieces against 1.0), so the accumulator that leaves TMEM
// is the squared distance minus |y|^2, which does not change the argmin.  Slot order: 0-1 |c|^2 pieces
// (A side 1.0), then 2 + 3 d + {0,1,2}; KS = 3 D + 2 rounded up to 8 (D = 10: 32 slots, four MMAs).
//
// Pipeline (one persistent CTA per SM, 25 warps, warp-specialised, mbarrier hand-offs):
//   warps 17-20  producers : build the A tile (128 frames x KS slots, K-major core-matrix layout) and
//                            the per-row screening threshold from the hinted centre (rows prefetched)
//   warps 21-24  finalisers: merge the column groups, certainty test, labels, fused Lloyd accumulation
//   warp  16     MMA issuer: per 256-centre chunk KS/8 tcgen05.mma into one of two TMEM buffers
//   warps 0-15   epilogue  : tcgen05.ld 64 columns per warp and chunk, running (best, second) as
//                            packed integer keys (score bits with the column in the low 5 bits)
// The centre operand B (Kpad x KS) is staged once per CTA and stays resident in shared memory.
//
// Exactness: a frame whose best and second-best scores are closer than a bound on the arithmetic
// error (operand rounding, dropped lo.lo terms, tensor-core accumulation, key truncation) is NOT
// labelled here: it goes to a list, and kmeans_recheck_kernel re-evaluates it against all centres
// in fp64 with the oracle's operation order (oracle/kmeans.py::sqdist_direct), one warp per frame.
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

constexpr int kTcTile = 128;     // frames per tile  (MMA M)
constexpr int kTcChunk = 256;    // centres per MMA  (MMA N)
constexpr int kTcEpiWarps = 16;
constexpr int kTcProdWarps = 4;
constexpr int kTcFinWarps = 4;
constexpr int kTcMmaWarp = kTcEpiWarps;
constexpr int kTcThreads = (kTcEpiWarps + 1 + kTcProdWarps + kTcFinWarps) * 32;   // 800
constexpr int kTcDReg = 16;      // coordinates of a frame held in registers (larger D: re-read from L1/L2)
constexpr int kTcColsPerWarp = kTcChunk / (kTcEpiWarps / 4);        // 64
constexpr float kTcErrScale = 1.9073486e-6f;   // 2^-19: envelope of the score error in units of (|y| + |c|max)^2
constexpr float kTcKeyTrunc = 1.0f + 7.6293945e-6f;  // 1 + 2^-17 (5 key bits of a 23-bit mantissa dropped)

struct KmTcParams {
  const float* Y;
  int64_t n;
  int D;
  int64_t ld;
  const double* centers;
  int K;
  int Kpad;   // K rounded up to a multiple of 256 (dummy centres score 2^126)
  int KS;     // K slots per row, multiple of 8
  int32_t* labels;
  const int32_t* hints;   // previous labels (may alias labels) or nullptr: only used to skip work
  double* sums;
  int64_t* counts;
  double* inertia;
  int64_t* n_rechecked;
  int* recheck_list;    // n entries
  int* recheck_count;   // zeroed by the host wrapper
  double* ysq;          // sum over frames of |y|^2 (zeroed by the host wrapper)
  double* lsums;        // K x D, this call's sums (zeroed by the host wrapper)
  unsigned long long* lcounts;   // K
  int accumulate;       // sums / counts / inertia requested
  float* dbg_scores;    // n x Kpad (tests only) or nullptr
};

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

struct TcSmem {
  uint64_t a_full[2], a_empty[2], t_full[2], t_empty[2], r_full[2], r_empty[2], thr_ready[4];
  uint32_t tmem_slot;
  float cmax;
  float red[32];
  float thr[4][kTcTile];   // per-row screening threshold of tile it (slot it & 3), minus |y|^2
  float xn2[4][kTcTile];   // |y|^2 of the row
  int res_best[2][4][kTcTile];
  int res_second[2][4][kTcTile];
  int res_chunk[2][4][kTcTile];
  float res_thr[2][kTcTile];
  float stage[kTcFinWarps][32][kTcDReg + 1];   // finalisers: transposed warp reduction of the coordinates
};

// DBG: also write the raw scores (tests).  INREG: D <= kTcDReg, a frame's coordinates live in registers.
template <bool DBG, bool INREG>
__global__ void __launch_bounds__(kTcThreads, 1) kmeans_tc_kernel(KmTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KS = p.KS, D = p.D, K = p.K, Kpad = p.Kpad;
  const uint32_t lbo = 128u, sbo = (uint32_t)(KS / 4) * 128u;
  unsigned char* sB = smem_raw;                                   // Kpad x KS floats
  unsigned char* sA = sB + (size_t)Kpad * KS * 4;                 // 2 x 128 x KS floats
  TcSmem* S = reinterpret_cast<TcSmem*>(sA + (size_t)2 * kTcTile * KS * 4);
  const int n_chunks = Kpad / kTcChunk;
  const int64_t n_tiles = (p.n + kTcTile - 1) / kTcTile;

  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&S->a_full[b], kTcProdWarps * 32);
      mbar_init(&S->a_empty[b], 1);
      mbar_init(&S->t_full[b], 1);
      mbar_init(&S->t_empty[b], kTcEpiWarps);
      mbar_init(&S->r_full[b], kTcEpiWarps * 32);
      mbar_init(&S->r_empty[b], kTcFinWarps * 32);
    }
    for (int b = 0; b < 4; ++b) mbar_init(&S->thr_ready[b], 1);
    fence_barrier_init();
  }
  if (warp == kTcMmaWarp) tc::tmem_alloc(&S->tmem_slot, 512);

  // ---- stage the centre operand: row k = n2 pieces (2) | [c'_hi c'_hi c'_lo]_d | 0...
  float cmax2 = 0.f;
  for (int k = tid; k < Kpad; k += kTcThreads) {
    auto put = [&](int slot, float v) {
      *reinterpret_cast<float*>(sB + tc::off_kmajor(k, slot, lbo, sbo)) = v;
    };
    for (int s = 0; s < KS; ++s) put(s, 0.f);
    if (k < K) {
      double n2 = 0.0;
      for (int d = 0; d < D; ++d) {
        const float c32 = (float)p.centers[(size_t)k * D + d];
        n2 = fma((double)c32, (double)c32, n2);
        const float cm = -2.0f * c32;
        const float hi = tf32_rna(cm);
        const float lo = tf32_rna(cm - hi);
        put(2 + 3 * d + 0, hi);
        put(2 + 3 * d + 1, hi);
        put(2 + 3 * d + 2, lo);
      }
      const float p1 = tf32_rna((float)n2);
      put(0, p1);
      put(1, tf32_rna((float)(n2 - (double)p1)));
      cmax2 = fmaxf(cmax2, (float)n2 * 1.0001f);
    } else {
      put(0, 8.507059e37f);   // 2^126: a dummy centre never wins
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cmax2 = fmaxf(cmax2, __shfl_xor_sync(0xffffffffu, cmax2, o));
  if (lane == 0) S->red[warp] = cmax2;
  fence_proxy_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (tid == 0) {
    float m = 0.f;
    for (int w = 0; w < kTcThreads / 32; ++w) m = fmaxf(m, S->red[w]);
    S->cmax = sqrtf(m) * 1.0001f;
  }
  __syncthreads();
  const uint32_t tmem = S->tmem_slot;
  const float cmax = S->cmax;

  if (warp < kTcEpiWarps) {
    // =========================================================== epilogue warps
    const int quarter = warp & 3, grp = warp >> 2;
    const int row = quarter * 32 + lane;
    uint32_t j = 0;   // chunk counter across tiles
    uint32_t it = 0;
    KM_DECL;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      int best = 0x7fffffff, second = 0x7fffffff, bchunk = 0;
      // screening threshold of this row (written by the producers, handed over by the MMA warp)
      KM_T(e0);
      mbar_wait(&S->thr_ready[it & 3u], (it >> 2) & 1u);
      KM_T(e1);
      KM_ACC(0, e0, e1);
      const float thr = S->thr[it & 3u][row];
      const float xn2 = S->xn2[it & 3u][row];
      for (int c = 0; c < n_chunks; ++c, ++j) {
        const uint32_t tb = j & 1u;
        KM_T(e2);
        mbar_wait(&S->t_full[tb], (j >> 1) & 1u);
        KM_T(e3);
        KM_ACC(1, e2, e3);
        tc::fence_after_sync();
#pragma unroll
        for (int h = 0; h < kTcColsPerWarp / 32; ++h) {
          const int col0 = grp * kTcColsPerWarp + h * 32;
          float v[32];
          tc::tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + tb * kTcChunk + (uint32_t)col0, v);
          if constexpr (DBG) {
            const int64_t grow = tile * kTcTile + row;
            if (grow < p.n)
              for (int q = 0; q < 32; ++q) p.dbg_scores[grow * Kpad + c * kTcChunk + col0 + q] = v[q] + xn2;
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
          if (!__any_sync(0xffffffffu, bmin <= thr)) continue;
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
          if (best != blk_before) bchunk = (c << 1) | h;   // low 5 key bits refer to this block
        }
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&S->t_empty[tb]);
        KM_T(e4);
        KM_ACC(2, e3, e4);
      }
      const uint32_t rb = it & 1u;
      KM_T(e5);
      mbar_wait(&S->r_empty[rb], ((it >> 1) & 1u) ^ 1u);
      KM_T(e6);
      KM_ACC(3, e5, e6);
      S->res_best[rb][grp][row] = best;
      S->res_second[rb][grp][row] = second;
      S->res_chunk[rb][grp][row] = bchunk;
      if (grp == 0) S->res_thr[rb][row] = thr + xn2;   // back in score space, like the keys
      tc::mbar_arrive(&S->r_full[rb]);
    }
    if (warp == 0) { KM_FLUSH(0, 4); }
  } else if (warp == kTcMmaWarp) {
    // =========================================================== MMA issuer (lane 0 issues, the warp stays converged)
    const uint32_t idesc = tc::idesc_tf32(kTcTile, kTcChunk, 0, 0);
    const uint32_t aB = smem_u32(sB), aA = smem_u32(sA);
    uint32_t j = 0, it = 0;
    KM_DECL;
    KM_T(m_begin);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ab = it & 1u;
      KM_T(m0);
      mbar_wait(&S->a_full[ab], (it >> 1) & 1u);
      KM_T(m1);
      KM_ACC(0, m0, m1);
      tc::fence_after_sync();
      if (lane == 0) tc::mbar_arrive(&S->thr_ready[it & 3u]);   // producers' thr[] -> epilogue warps
      const uint32_t a_base = aA + ab * (uint32_t)(kTcTile * KS * 4);
      for (int c = 0; c < n_chunks; ++c, ++j) {
        const uint32_t tb = j & 1u;
        KM_T(m2);
        mbar_wait(&S->t_empty[tb], ((j >> 1) & 1u) ^ 1u);
        KM_T(m3);
        KM_ACC(1, m2, m3);
        tc::fence_after_sync();
        {
          // descriptor of k-step s = base descriptor + (s * 2 * lbo >> 4) in the 14-bit address field
          const uint64_t da0 = tc::smem_desc(a_base, lbo, sbo, tc::kLayoutNone);
          const uint64_t db0 = tc::smem_desc(aB + (uint32_t)c * (kTcChunk / 8) * sbo, lbo, sbo, tc::kLayoutNone);
          const uint32_t d_tmem = tmem + tb * kTcChunk;
          const int nks = KS / 8;
#pragma unroll 1
          for (int s = 0; s < nks; ++s) {
            const uint64_t step = (uint64_t)((uint32_t)s * ((2u * lbo) >> 4));
            tc::mma_tf32_elect(d_tmem, da0 + step, db0 + step, idesc, s > 0 ? 1u : 0u);
          }
          tc::mma_commit_elect(&S->t_full[tb]);
        }
      }
      tc::mma_commit_elect(&S->a_empty[ab]);
    }
    KM_T(m_end);
    KM_ACC(2, m_begin, m_end);
#ifdef PMB_KM_PROF
    km_prof[3] = (long long)it;
#endif
    KM_FLUSH(4, 4);
  } else if (warp < kTcMmaWarp + 1 + kTcProdWarps) {
    // =========================================================== producers: A tile + screening threshold
    const int pt = tid - (kTcMmaWarp + 1) * 32;   // 0..127: row of the tile
    constexpr bool in_regs = INREG;
    float yn[kTcDReg];   // coordinates of this thread's row of the NEXT tile (prefetched)
    int hint_n = -1;
    auto prefetch = [&](int64_t tile) {
      const int64_t row = tile * kTcTile + pt;
      hint_n = -1;
      if (row < p.n) {
        if (p.hints != nullptr) hint_n = p.hints[row];
        if constexpr (in_regs) {
#pragma unroll
          for (int d = 0; d < kTcDReg; ++d) yn[d] = (d < D) ? p.Y[row * p.ld + d] : 0.f;
        }
      } else {
#pragma unroll
        for (int d = 0; d < kTcDReg; ++d) yn[d] = 0.f;
      }
    };
    uint32_t it = 0;
    double ysq_acc = 0.0;
    KM_DECL;
    if ((int64_t)blockIdx.x < n_tiles) prefetch(blockIdx.x);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ab = it & 1u;
      const int64_t row = tile * kTcTile + pt;
      const bool valid = row < p.n;
      float y[kTcDReg];
#pragma unroll
      for (int d = 0; d < kTcDReg; ++d) y[d] = yn[d];
      const int hint = hint_n;
      if (tile + gridDim.x < n_tiles) prefetch(tile + gridDim.x);
      // distance to the hinted centre from the resident centre operand: c32 = -(c'_hi + c'_lo) / 2 exactly
      float u = __int_as_float(0x7f800000);
      if (valid && hint >= 0 && hint < K) {
        u = 0.f;
        auto cget = [&](int d) {
          const float hi = *reinterpret_cast<const float*>(sB + tc::off_kmajor(hint, 2 + 3 * d, lbo, sbo));
          const float lo = *reinterpret_cast<const float*>(sB + tc::off_kmajor(hint, 2 + 3 * d + 2, lbo, sbo));
          return -0.5f * (hi + lo);
        };
        if constexpr (in_regs) {
#pragma unroll
          for (int d = 0; d < kTcDReg; ++d) {
            if (d < D) {
              const float t = y[d] - cget(d);
              u = fmaf(t, t, u);
            }
          }
        } else {
#pragma unroll 4
          for (int d = 0; d < D; ++d) {
            const float t = p.Y[row * p.ld + d] - cget(d);
            u = fmaf(t, t, u);
          }
        }
      }
      KM_T(p0);
      mbar_wait(&S->a_empty[ab], ((it >> 1) & 1u) ^ 1u);
      KM_T(p1);
      KM_ACC(0, p0, p1);
      unsigned char* A = sA + (size_t)ab * kTcTile * KS * 4;
      float xn2 = 0.f;
      if constexpr (in_regs) {
#pragma unroll
        for (int d = 0; d < kTcDReg; ++d) {
          const float v = y[d];          // zero beyond D and for rows past the end
          xn2 = fmaf(v, v, xn2);
          ysq_acc = fma((double)v, (double)v, ysq_acc);
        }
        const float one = valid ? 1.0f : 0.0f;
        // slot s: 0-1 -> 1, 2 + 3 d + r -> (r == 1 ? y_lo[d] : y_hi[d]); 16-byte stores, 8 lanes
        // cover one 128-byte core-matrix row: conflict-free
        auto slot_val = [&](int sl) -> float {
          if (sl < 2) return one;
          const int d = (sl - 2) / 3, r = (sl - 2) % 3;
          if (d >= kTcDReg) return 0.f;
          const float hi = tf32_rna(y[d]);   // recomputed per use: keeps the producer's register count low
          return r == 1 ? tf32_rna(y[d] - hi) : hi;
        };
        unsigned char* arow = A + (uint32_t)(pt >> 3) * sbo + (uint32_t)(pt & 7) * 16u;
#pragma unroll
        for (int j = 0; j < (3 * kTcDReg + 2 + 7) / 8 * 2; ++j) {
          if (4 * j < KS)
            *reinterpret_cast<float4*>(arow + (uint32_t)j * lbo) =
                make_float4(slot_val(4 * j), slot_val(4 * j + 1), slot_val(4 * j + 2), slot_val(4 * j + 3));
        }
      } else {
        auto put = [&](int slot, float v) {
          *reinterpret_cast<float*>(A + tc::off_kmajor(pt, slot, lbo, sbo)) = v;
        };
#pragma unroll 4
        for (int d = 0; d < D; ++d) {
          const float v = valid ? p.Y[row * p.ld + d] : 0.f;
          xn2 = fmaf(v, v, xn2);
          ysq_acc = fma((double)v, (double)v, ysq_acc);
          const float hi = tf32_rna(v);
          put(2 + 3 * d + 0, hi);
          put(2 + 3 * d + 1, tf32_rna(v - hi));
          put(2 + 3 * d + 2, hi);
        }
        const float one = valid ? 1.0f : 0.0f;
        put(0, one);
        put(1, one);
        for (int sl = 3 * D + 2; sl < KS; ++sl) put(sl, 0.f);
      }
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
    if (warp == kTcMmaWarp + 1) { KM_FLUSH(8, 2); }
    if (p.accumulate) {
      ysq_acc = warp_sum(ysq_acc);
      if (lane == 0 && ysq_acc != 0.0) atomicAdd(p.ysq, ysq_acc);
    }
  } else {
    // =========================================================== finalisers: merge, certainty test, labels,
    // fused Lloyd accumulation
    const int pt = tid - (kTcMmaWarp + 1 + kTcProdWarps) * 32;
    constexpr bool in_regs = INREG;
    int recheck_acc = 0;
    KM_DECL;
    float yn[kTcDReg];
    auto prefetch = [&](int64_t tile) {
      const int64_t row = tile * kTcTile + pt;
#pragma unroll
      for (int d = 0; d < kTcDReg; ++d) yn[d] = (in_regs && d < D && row < p.n) ? p.Y[row * p.ld + d] : 0.f;
    };
    uint32_t it = 0;
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
      int b = 0x7fffffff, s2 = 0x7fffffff, bc = 0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int gb = S->res_best[rb][g][pt], gs = S->res_second[rb][g][pt], gc = S->res_chunk[rb][g][pt];
        const int t = max(b, gb);           // merge: new second = min(s2, gs, max(b, gb))
        s2 = min(min(s2, gs), t);
        if (gb < b) { b = gb; bc = (gc << 2) | g; }
      }
      const float thr = S->res_thr[rb][pt];
      tc::mbar_arrive(&S->r_empty[rb]);
      const int64_t row = tile * kTcTile + pt;
      const bool valid = row < p.n;
      int lab = -1;
      if (valid) {
        // column = chunk * 256 + group * 64 + h * 32 + (key & 31);  bc = ((chunk << 1 | h) << 2) | group
        const int g = bc & 3, h = (bc >> 2) & 1, ch = bc >> 3;
        int k = ch * kTcChunk + g * kTcColsPerWarp + h * 32 + (b & 31);
        if (k >= K) k = K - 1;   // cannot happen for finite data (dummy centres score 2^126)
        const float s1f = __uint_as_float((uint32_t)b & 0xFFFFFFE0u);
        const float s2f = fminf(__uint_as_float((uint32_t)s2 & 0xFFFFFFE0u), thr);
        float xn2 = 0.f;
        if constexpr (in_regs) {
#pragma unroll
          for (int d = 0; d < kTcDReg; ++d) xn2 = fmaf(y[d], y[d], xn2);
        } else {
#pragma unroll 4
          for (int d = 0; d < D; ++d) {
            const float v = p.Y[row * p.ld + d];
            xn2 = fmaf(v, v, xn2);
          }
        }
        const float rr = sqrtf(xn2) * 1.0001f + cmax;
        const float E = kTcErrScale * rr * rr;
        const bool certain = (b >= 0) && (b != 0x7fffffff) && (K == 1 || (s1f * kTcKeyTrunc + E < s2f - E));
        p.labels[row] = k;
        if (certain) {
          lab = k;
        } else {
          ++recheck_acc;
          const int pos = atomicAdd(p.recheck_count, 1);
          p.recheck_list[pos] = (int)row;
        }
      }
      // fused Lloyd accumulation into this call's private sums / counts.  32 consecutive frames of a
      // trajectory mostly share their label: one warp reduction and D atomics; otherwise one atomic per
      // frame and coordinate.
      if (p.accumulate) {
        const int lab0 = __shfl_sync(0xffffffffu, lab, 0);
        if (__all_sync(0xffffffffu, lab == lab0)) {
          if (lab0 >= 0) {
            if constexpr (in_regs) {
              // transpose through shared memory: lane (d, half) adds 16 rows of coordinate d in fp64
              float* stg = &S->stage[warp - (kTcMmaWarp + 1 + kTcProdWarps)][0][0];
#pragma unroll
              for (int d = 0; d < kTcDReg; ++d) stg[lane * (kTcDReg + 1) + d] = y[d];
              __syncwarp();
              const int dd = lane & 15, r0 = (lane >> 4) * 16;
              double acc = 0.0;
#pragma unroll
              for (int r = 0; r < 16; ++r) acc += (double)stg[(r0 + r) * (kTcDReg + 1) + dd];
              acc += __shfl_xor_sync(0xffffffffu, acc, 16);
              if (lane < D) atomicAdd(p.lsums + (size_t)lab0 * D + lane, acc);
              __syncwarp();
            } else {
              for (int d = 0; d < D; ++d) {
                const double v = warp_sum((double)p.Y[row * p.ld + d]);
                if (lane == 0) atomicAdd(p.lsums + (size_t)lab0 * D + d, v);
              }
            }
            if (lane == 0) atomicAdd(p.lcounts + lab0, 32ull);
          }
        } else if (lab >= 0) {
          if constexpr (in_regs) {
#pragma unroll
            for (int d = 0; d < kTcDReg; ++d)
              if (d < D) atomicAdd(p.lsums + (size_t)lab * D + d, (double)y[d]);
          } else {
            for (int d = 0; d < D; ++d) atomicAdd(p.lsums + (size_t)lab * D + d, (double)p.Y[row * p.ld + d]);
          }
          atomicAdd(p.lcounts + lab, 1ull);
        }
      }
      KM_T(f2);
      KM_ACC(1, f1, f2);
    }
    if (warp == kTcMmaWarp + 1 + kTcProdWarps) { KM_FLUSH(10, 2); }
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

static inline int tc_slots(int D) { return ((3 * D + 2) + 7) / 8 * 8; }
static inline int tc_kpad(int K) { return (K + kTcChunk - 1) / kTcChunk * kTcChunk; }
static inline size_t tc_smem_bytes(int D, int K) {
  return (size_t)tc_kpad(K) * tc_slots(D) * 4 + (size_t)2 * kTcTile * tc_slots(D) * 4 + sizeof(TcSmem) + 16;
}

bool kmeans_tc_supported(int D, int K) {
  return D >= 1 && K >= 1 && tc_smem_bytes(D, K) <= 227 * 1024 && tc_kpad(K) / kTcChunk <= (1 << 20);
}

size_t kmeans_tc_ws_bytes(int64_t n, int D, int K) {
  // [count | ysq | pad to 64 B][local sums K x D][local counts K][re-check list n]
  return 64 + ((size_t)K * D + K) * sizeof(double) + (size_t)n * sizeof(int) + 64;
}

int kmeans_tc_assign(const float* Y, int64_t n, int D, int64_t ld, const double* centers, int K, int32_t* labels,
                     const int32_t* hints, double* sums, int64_t* counts, double* inertia, int64_t* n_rechecked, void* ws,
                     float* dbg_scores, cudaStream_t st) {
  KmTcParams p;
  p.Y = Y; p.n = n; p.D = D; p.ld = ld; p.centers = centers; p.K = K;
  p.Kpad = tc_kpad(K); p.KS = tc_slots(D);
  p.labels = labels; p.hints = hints; p.sums = sums; p.counts = counts; p.inertia = inertia; p.n_rechecked = n_rechecked;
  unsigned char* w8 = static_cast<unsigned char*>(ws);
  p.recheck_count = reinterpret_cast<int*>(w8);
  p.ysq = reinterpret_cast<double*>(w8 + 16);
  p.lsums = reinterpret_cast<double*>(w8 + 64);
  p.lcounts = reinterpret_cast<unsigned long long*>(p.lsums + (size_t)K * D);
  p.recheck_list = reinterpret_cast<int*>(p.lcounts + K);
  p.accumulate = (sums != nullptr || inertia != nullptr) ? 1 : 0;
  p.dbg_scores = dbg_scores;
  PMB_REQUIRE(n < (int64_t)0x7fffffff, "pmb_kmeans_assign: tensor path needs n < 2^31");
  const size_t smem = tc_smem_bytes(D, K);
  PMB_CUDA(cudaMemsetAsync(ws, 0, 64 + (p.accumulate ? ((size_t)K * D + K) * sizeof(double) : 0), st));
  const int64_t n_tiles = (n + kTcTile - 1) / kTcTile;
  const int grid = (int)(n_tiles < kNumSMs ? n_tiles : kNumSMs);
  auto launch = [&](auto kern) -> int {
    PMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kTcThreads, smem, st>>>(p);
    PMB_LAUNCH_CHECK();
    return PMB_OK;
  };
  int rc;
  if (dbg_scores != nullptr) rc = (D <= kTcDReg) ? launch(kmeans_tc_kernel<true, true>) : launch(kmeans_tc_kernel<true, false>);
  else rc = (D <= kTcDReg) ? launch(kmeans_tc_kernel<false, true>) : launch(kmeans_tc_kernel<false, false>);
  if (rc != PMB_OK) return rc;
  kmeans_recheck_kernel<<<2 * kNumSMs, 256, 0, st>>>(p);
  PMB_LAUNCH_CHECK();
  if (p.accumulate) {
    kmeans_tc_commit_kernel<<<(K + 255) / 256 < 64 ? (K + 255) / 256 : 64, 256, 0, st>>>(p);
    PMB_LAUNCH_CHECK();
  }
  return PMB_OK;
}

}  // namespace pmb
