// gram_tc.cu -- K3 (tensor-core path): the TICA Gram matrices as tcgen05 TF32 products with fp64-grade
// accuracy, accumulators in TMEM.
//
//   mode 0: G = sum_g w_g z_g z_g^T,  w_g = popcount(mask_g & 3)      (X0^T X0 + Xt^T Xt)
//   mode 1: G = sum_g (mask_g & 1) v_g v_g^T,  v_g = z_g - z_{g+lag}  (-> X0^T Xt + Xt^T X0)
//   z = NaN ? 0 : (x - shift) * scale                                  (same conditioning as gram.cu)
//
// Split ("3xTF32 with an exactly accumulating leading term").  Tensor-core fp32 accumulation of the
// monotone diagonal sums over millions of frames is the accuracy limit of a plain 3xTF32 scheme, so
// each conditioned value is split into THREE TF32-representable parts
//      z = a + r_hi + r_lo,   a = round(8 z + dither_t) / 8 clamped to [-8, 8],  r = z - a (exact),
//      r_hi = tf32(r),  r_lo = tf32(r - r_hi)
// and
//      w z z^T ~= (w a) a^T + [(w a) r^T + r (w a)^T] + (w r_hi) r_hi^T
//   P  += (2 w a)^T a                         every product is a multiple of 2^-5: the fp32 sums in TMEM are
//                                             EXACT while a window's partial sums stay below 2^19
//   Q  += (2 w a)^T r_hi + (2 w a)^T r_lo + (w r_hi)^T r_hi      (|terms| <= 2^-4 |z|: their fp32
//                                             accumulation error is ~1e-9 of G)
//      G = P / 2 + (Q + Q^T) / 2
// Dropped: r_hi r_lo^T + r_lo r_hi^T + r_lo r_lo^T, zero-mean, ~2^-21 |z|^2 per frame.
// Every kWindow frames the two 128 x d accumulators are drained into a per-CTA fp64 workspace.
//
// One CTA per (128-row block of G, frame chunk); 148 CTAs = 2 row blocks x 74 chunks for d = 256.
//   warps 0-15 producers: coalesced float4 loads (next stage prefetched in registers), conditioning,
//              split, 16-byte stores into the MN-major 128B-swizzled operand tiles; drain TMEM at the
//              end of a window
//   warp  16   MMA issuer: 8 tcgen05.mma (M = 128, N = d, K = 8) per 16-frame stage
// 3-stage shared-memory ring, mbarrier full/empty hand-off, tcgen05.commit frees a stage.
#include "tc05.cuh"

namespace pmb {

constexpr int kGtBK = 16;            // frames per stage
constexpr int kGtStages = 3;
constexpr int kGtProdWarps = 16;     // 4 producer warps per scheduler: the producers must keep ~60% of the issue slots busy
constexpr int kGtF4 = 64 / (kGtProdWarps * 32 / kGtBK);   // float4 groups per thread and stage (2)
constexpr int kGtThreads = (kGtProdWarps + 1) * 32;   // 544
constexpr int kGtWindow = 4096;      // frames per TMEM accumulation window (multiple of kGtBK)
constexpr uint32_t kGtTileB = 256 * kGtBK * 4;        // bytes of one N-side tile (256 features x 16 frames)
constexpr uint32_t kGtTileA = 128 * kGtBK * 4;
constexpr uint32_t kGtStageBytes = 3 * kGtTileB + 2 * kGtTileA;   // a, r_hi, r_lo | 2wa, w r_hi  = 64 KB

struct GramTcParams {
  const float* X;
  int64_t n;
  int d;
  int64_t ld;
  const uint8_t* mask;
  int lag;
  int mode;
  const float* shift;
  const float* scale;
  double* ws;        // [n_rb * n_chunks][2][d][128]
  int n_chunks;
  int64_t chunk;     // frames per chunk, multiple of kGtBK
};

__device__ __forceinline__ float gt_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float gt_cond(float x, float sh, float sc) { return (x == x) ? (x - sh) * sc : 0.0f; }

struct GtBars {
  uint64_t full[kGtStages], empty[kGtStages], wdone, drained;
  uint32_t tmem_slot;
};

__global__ void __launch_bounds__(kGtThreads, 1) gram_tc_kernel(GramTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  GtBars* B = reinterpret_cast<GtBars*>(tiles + (size_t)kGtStages * kGtStageBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int d = p.d;
  const int rb = blockIdx.x / p.n_chunks, ck = blockIdx.x - rb * p.n_chunks;
  const int64_t g_begin = (int64_t)ck * p.chunk;
  int64_t g_end = g_begin + p.chunk;
  if (g_end > p.n) g_end = p.n;
  const int n_stages = g_begin < g_end ? (int)((g_end - g_begin + kGtBK - 1) / kGtBK) : 0;
  const int stages_per_window = kGtWindow / kGtBK;
  const uint32_t lbo = kGtBK * 128u, sbo = 512u;

  // zero every tile once (features >= d and the unused half of narrow row blocks stay zero)
  for (uint32_t i = tid; i < kGtStages * kGtStageBytes / 16; i += kGtThreads)
    reinterpret_cast<uint4*>(tiles)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    for (int s = 0; s < kGtStages; ++s) {
      mbar_init(&B->full[s], kGtProdWarps * 32);
      mbar_init(&B->empty[s], 1);
    }
    mbar_init(&B->wdone, 1);
    mbar_init(&B->drained, kGtProdWarps);
    fence_barrier_init();
  }
  if (warp == kGtProdWarps) tc::tmem_alloc(&B->tmem_slot, 512);
  fence_proxy_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = B->tmem_slot;
  double* ws = p.ws + (size_t)blockIdx.x * 2 * d * 128;

  if (warp < kGtProdWarps) {
    // ============================================================ producers
    constexpr int kTpf = kGtProdWarps * 32 / kGtBK;   // threads per frame (32)
    const int fr = tid / kTpf;        // frame within the stage
    const int g4 = tid % kTpf;        // float4 group: handles float4 indices g4 + kTpf j
    const int nf4 = d >> 2;
    float4 sh[kGtF4], sc[kGtF4];
#pragma unroll
    for (int j = 0; j < kGtF4; ++j) {
      const int c = (g4 + kTpf * j) * 4;
      if (c < d) {
        sh[j] = *reinterpret_cast<const float4*>(p.shift + c);
        sc[j] = *reinterpret_cast<const float4*>(p.scale + c);
      } else {
        sh[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        sc[j] = sh[j];
      }
    }
    float4 xa[kGtF4], xb[kGtF4];
    int mk = 0;
    auto prefetch = [&](int s) {
      const int64_t g = g_begin + (int64_t)s * kGtBK + fr;
      mk = 0;
      if (g < g_end) mk = p.mask[g];
      const int w = (p.mode == 0) ? __popc(mk & 3) : (mk & 1);
#pragma unroll
      for (int j = 0; j < kGtF4; ++j) {
        xa[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        xb[j] = xa[j];
        const int idx = g4 + kTpf * j;
        if (w && idx < nf4) {
          xa[j] = ldg_stream_f4(reinterpret_cast<const float4*>(p.X + g * p.ld) + idx);
          if (p.mode == 1) xb[j] = ldg_stream_f4(reinterpret_cast<const float4*>(p.X + (g + p.lag) * p.ld) + idx);
        }
      }
      mk = w;
    };
    // element offset of (feature mn, frame k) inside an MN-major swizzled tile
    auto toff = [&](int mn, int k) -> uint32_t {
      return tc::off_mnmajor_sw128b32(mn, k, lbo, sbo);
    };
    auto drain = [&](bool first) {
      // warp w: lane quarter w % 4, accumulator (w / 4): 0 = P (columns 0..d), 1 = Q (columns 256..256+d)
      const int quarter = warp & 3, which = (warp >> 2) & 1, part = warp >> 3, nparts = kGtProdWarps / 8;
      double* dst = ws + (size_t)which * d * 128 + quarter * 32 + lane;
      for (int c0 = part * 32; c0 < d; c0 += 32 * nparts) {
        float v[32];
        tc::tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(which * 256 + c0), v);
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          double* e = dst + (size_t)(c0 + q) * 128;
          *e = first ? (double)v[q] : (*e + (double)v[q]);
        }
      }
    };

    if (n_stages > 0) prefetch(0);
    int n_windows_done = 0;
    for (int s = 0; s < n_stages; ++s) {
      const int slot = s % kGtStages;
      const uint32_t use = (uint32_t)(s / kGtStages);
      // current stage data -> registers, then prefetch the next stage
      float4 ca[kGtF4], cb[kGtF4];
#pragma unroll
      for (int j = 0; j < kGtF4; ++j) { ca[j] = xa[j]; cb[j] = xb[j]; }
      const int w = mk;
      if (s + 1 < n_stages) prefetch(s + 1);
      mbar_wait(&B->empty[slot], (use & 1u) ^ 1u);
      unsigned char* T = tiles + (size_t)slot * kGtStageBytes;
      unsigned char* Ta = T, *Th = T + kGtTileB, *Tl = T + 2 * kGtTileB;
      unsigned char* T2a = T + 3 * kGtTileB, *Twh = T2a + kGtTileA;
      const float wf = (float)w;
      // stochastic rounding of the leading term: one dither per frame (hash of the frame index).  With
      // round-to-nearest a tightly clustered feature gives residuals r of one sign, and the a r^T sums
      // then grow monotonically in the fp32 accumulator (systematic truncation error, measured 3e-7 of
      // scale); with a dithered grid E[r | z] = 0 and the residual correlation is <= 2^-8 per frame.
      float dith;
      {
        uint32_t h = (uint32_t)(g_begin + (int64_t)s * kGtBK + fr) * 0x9E3779B1u;
        h ^= h >> 15;
        h *= 0x85EBCA6Bu;
        h ^= h >> 13;
        dith = __uint_as_float((h >> 9) | 0x3f800000u) - 1.5f;   // uniform in [-0.5, 0.5)
      }
#pragma unroll
      for (int j = 0; j < kGtF4; ++j) {
        const int idx = g4 + kTpf * j;
        if (idx >= nf4) continue;
        float z[4];
        z[0] = gt_cond(ca[j].x, sh[j].x, sc[j].x);
        z[1] = gt_cond(ca[j].y, sh[j].y, sc[j].y);
        z[2] = gt_cond(ca[j].z, sh[j].z, sc[j].z);
        z[3] = gt_cond(ca[j].w, sh[j].w, sc[j].w);
        if (p.mode == 1) {
          z[0] -= gt_cond(cb[j].x, sh[j].x, sc[j].x);
          z[1] -= gt_cond(cb[j].y, sh[j].y, sc[j].y);
          z[2] -= gt_cond(cb[j].z, sh[j].z, sc[j].z);
          z[3] -= gt_cond(cb[j].w, sh[j].w, sc[j].w);
        }
        float a[4], rh[4], rl[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float zz = w ? z[q] : 0.f;
          a[q] = fminf(fmaxf(rintf(fmaf(zz, 8.0f, dith)), -64.0f), 64.0f) * 0.125f;
          const float r = zz - a[q];
          rh[q] = gt_tf32(r);
          rl[q] = gt_tf32(r - rh[q]);
        }
        const int mn = idx * 4;
        const uint32_t o = toff(mn, fr);
        *reinterpret_cast<float4*>(Ta + o) = make_float4(a[0], a[1], a[2], a[3]);
        *reinterpret_cast<float4*>(Th + o) = make_float4(rh[0], rh[1], rh[2], rh[3]);
        *reinterpret_cast<float4*>(Tl + o) = make_float4(rl[0], rl[1], rl[2], rl[3]);
        const int mloc = mn - rb * 128;
        if (mloc >= 0 && mloc < 128) {
          const uint32_t oa = toff(mloc, fr);
          const float w2 = 2.0f * wf;
          *reinterpret_cast<float4*>(T2a + oa) = make_float4(w2 * a[0], w2 * a[1], w2 * a[2], w2 * a[3]);
          *reinterpret_cast<float4*>(Twh + oa) = make_float4(wf * rh[0], wf * rh[1], wf * rh[2], wf * rh[3]);
        }
      }
      fence_proxy_async_smem();
      tc::mbar_arrive(&B->full[slot]);

      const bool window_end = ((s + 1) % stages_per_window == 0) || (s + 1 == n_stages);
      if (window_end) {
        mbar_wait(&B->wdone, (uint32_t)(n_windows_done & 1));
        tc::fence_after_sync();
        drain(n_windows_done == 0);
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&B->drained);
        ++n_windows_done;
      }
    }
    if (n_stages == 0) {
      // empty chunk: the reduce kernel still reads this CTA's slice
      for (int i = tid; i < 2 * d * 128; i += kGtProdWarps * 32) ws[i] = 0.0;
    }
  } else {
    // ============================================================ MMA issuer
    // The issuing warp's own instruction stream sets the pace of the tensor pipe (tools/ubench_mma.cu): the
    // descriptors of a ring slot differ from those of slot 0 by a constant in the 14-bit address field, so the
    // ten descriptors of a stage (5 tiles x 2 K steps) are slot base + precomputed offsets, and ONE elected
    // lane issues the stage's eight MMAs and its commit(s) from a single instruction block.
    const uint32_t idesc = tc::idesc_tf32(128, d, 1, 1);
    const uint32_t base = smem_u32(tiles);
    const uint64_t d0 = tc::smem_desc(base, lbo, sbo, tc::kLayoutSw128Base32);     // tile `a`, K step 0, slot 0
    const uint64_t oTh = kGtTileB >> 4, oTl = (2 * kGtTileB) >> 4, oT2a = (3 * kGtTileB) >> 4,
                   oTwh = (3 * kGtTileB + kGtTileA) >> 4, oK = (2u * sbo) >> 4, oSlot = kGtStageBytes >> 4;
    const uint32_t bar_empty0 = smem_u32(&B->empty[0]), bar_wdone = smem_u32(&B->wdone);
    int n_windows_done = 0;
    int slot = 0;
    uint32_t use = 0;
    int in_window = 0;   // stages issued in the current window
    for (int s = 0; s < n_stages; ++s) {
      const bool window_start = in_window == 0;
      if (window_start && n_windows_done > 0) {
        mbar_wait(&B->drained, (uint32_t)((n_windows_done - 1) & 1));
        tc::fence_after_sync();
      }
      mbar_wait(&B->full[slot], use & 1u);
      tc::fence_after_sync();
      const uint64_t dS = d0 + (uint64_t)slot * oSlot;
      const bool window_end = (in_window + 1 == stages_per_window) || (s + 1 == n_stages);
      const uint32_t acc0 = window_start ? 0u : 1u;
      // K step 0 then K step 1: P += (2wa)^T a ; Q += (2wa)^T r_hi + (2wa)^T r_lo + (w r_hi)^T r_hi
      asm volatile(
          "{\n\t"
          ".reg .pred q, p;\n\t"
          ".reg .b64 a2, awh, ba, bh, bl;\n\t"
          "elect.sync _|q, 0xffffffff;\n\t"
          "setp.ne.b32 p, %4, 0;\n\t"
          "add.u64 a2, %2, %7;\n\t"        // 2wa
          "add.u64 awh, %2, %8;\n\t"       // w r_hi
          "add.u64 bh, %2, %5;\n\t"        // r_hi
          "add.u64 bl, %2, %6;\n\t"        // r_lo
          "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], a2, %2, %3, p;\n\t"
          "@q tcgen05.mma.cta_group::1.kind::tf32 [%1], a2, bh, %3, p;\n\t"
          "@q tcgen05.mma.cta_group::1.kind::tf32 [%1], a2, bl, %3, 1;\n\t"
          "@q tcgen05.mma.cta_group::1.kind::tf32 [%1], awh, bh, %3, 1;\n\t"
          "add.u64 a2, a2, %9;\n\t"
          "add.u64 awh, awh, %9;\n\t"
          "add.u64 bh, bh, %9;\n\t"
          "add.u64 bl, bl, %9;\n\t"
          "add.u64 ba, %2, %9;\n\t"
          "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], a2, ba, %3, 1;\n\t"
          "@q tcgen05.mma.cta_group::1.kind::tf32 [%1], a2, bh, %3, 1;\n\t"
          "@q tcgen05.mma.cta_group::1.kind::tf32 [%1], a2, bl, %3, 1;\n\t"
          "@q tcgen05.mma.cta_group::1.kind::tf32 [%1], awh, bh, %3, 1;\n\t"
          "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%10];\n\t"
          "}" ::"r"(tmem),
          "r"(tmem + 256u), "l"(dS), "r"(idesc), "r"(acc0), "l"(oTh), "l"(oTl), "l"(oT2a), "l"(oTwh), "l"(oK),
          "r"(bar_empty0 + (uint32_t)slot * 8u)
          : "memory");
      if (window_end) {
        asm volatile(
            "{\n\t"
            ".reg .pred q;\n\t"
            "elect.sync _|q, 0xffffffff;\n\t"
            "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
            "}" ::"r"(bar_wdone)
            : "memory");
        ++n_windows_done;
        in_window = 0;
      } else {
        ++in_window;
      }
      if (++slot == kGtStages) { slot = 0; ++use; }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == kGtProdWarps) tc::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant for d = 256 (cta_group::2), selected with impl = 4.  NOT the default: it is correct
// (same parity tests) but measured 1.3x SLOWER than the kernel above on B200.
// Idea: in the kernel above the two row blocks of a frame chunk run on two unrelated SMs and each of them
// loads, conditions and splits ALL 256 features of every frame for its N side.  Here the two row blocks are
// the two CTAs of a cluster and issue ONE M = 256, N = 256 MMA: each CTA stages only its own 128 features --
// they are both its 128 rows of A (2wa, w r_hi) and its half of B (a, r_hi, r_lo) -- so the producer work per
// frame is halved.  Stage = 5 tiles of 128 features x 16 frames = 40 KB, 5 stages.
//   producers (16 warps per CTA, one frame each per stage) -> arrive on the LEADER's full barrier
//   MMA warp of the leader CTA: 8 tcgen05.mma.cta_group::2 per stage, commit multicast to both CTAs'
//   empty / window barriers; each CTA drains its own TMEM (its 128 rows of P and Q).
// Measurement (2.5 M frames, both Gram launches): single-CTA 4.92 ms, of which 3.50 ms remain when the
// producers store nothing -- 204 cycles per M128 N256 K8 MMA, the TF32 rate cuBLAS sustains -- so that kernel
// is MMA-bound with 1.4 ms of imperfect overlap.  The pair kernel: 6.45 ms, 5.75 ms with idle producers,
// 335 cycles per MMA with all 74 clusters resident: with 4-byte operands the half of B that every MMA pulls
// from the peer SM's shared memory is the bound, not the staging work the pairing saves.
constexpr int kGpStages = 5;
constexpr uint32_t kGpTile = 128 * kGtBK * 4;        // 8 KB
constexpr uint32_t kGpStageBytes = 5 * kGpTile;      // a, r_hi, r_lo | 2wa, w r_hi
static_assert(kGtProdWarps == kGtBK, "one producer warp per frame of a stage");

struct GpBars {
  uint64_t full[kGpStages], empty[kGpStages], wdone, drained;
  uint32_t tmem_slot;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGtThreads, 1) gram_tc_pair_kernel(GramTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  GpBars* B = reinterpret_cast<GpBars*>(tiles + (size_t)kGpStages * kGpStageBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int d = 256;
  const uint32_t rank = tc::cluster_ctarank();   // = row block
  const int ck = (int)(blockIdx.x >> 1);
  const int64_t g_begin = (int64_t)ck * p.chunk;
  int64_t g_end = g_begin + p.chunk;
  if (g_end > p.n) g_end = p.n;
  const int n_stages = g_begin < g_end ? (int)((g_end - g_begin + kGtBK - 1) / kGtBK) : 0;
  const int stages_per_window = kGtWindow / kGtBK;
  const uint32_t lbo = kGtBK * 128u, sbo = 512u;

  for (uint32_t i = tid; i < kGpStages * kGpStageBytes / 16; i += kGtThreads)
    reinterpret_cast<uint4*>(tiles)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    for (int s = 0; s < kGpStages; ++s) {
      mbar_init(&B->full[s], 2 * kGtProdWarps);   // one arrival per producer warp of BOTH CTAs (leader's copy is used)
      mbar_init(&B->empty[s], 1);
    }
    mbar_init(&B->wdone, 1);
    mbar_init(&B->drained, 2 * kGtProdWarps);
    fence_barrier_init();
  }
  if (warp == kGtProdWarps) tc::tmem_alloc_pair(&B->tmem_slot, 512);
  fence_proxy_async_smem();
  tc::fence_before_sync();
  tc::cluster_sync();
  tc::fence_after_sync();
  const uint32_t tmem = B->tmem_slot;
  double* ws = p.ws + ((size_t)rank * p.n_chunks + ck) * 2 * d * 128;

  if (warp < kGtProdWarps) {
    // ============================================================ producers: warp = frame of the stage
    const int fr = warp;
    const int c = (int)rank * 128 + lane * 4;   // first of this lane's four features
    const float4 sh = *reinterpret_cast<const float4*>(p.shift + c);
    const float4 sc = *reinterpret_cast<const float4*>(p.scale + c);
    float4 xa, xb;
    int mk = 0;
    auto prefetch = [&](int s) {
      const int64_t g = g_begin + (int64_t)s * kGtBK + fr;
      int m = 0;
      if (g < g_end) m = p.mask[g];
      const int w = (p.mode == 0) ? __popc(m & 3) : (m & 1);
      xa = make_float4(0.f, 0.f, 0.f, 0.f);
      xb = xa;
      if (w) {
        xa = ldg_stream_f4(reinterpret_cast<const float4*>(p.X + g * p.ld + c));
        if (p.mode == 1) xb = ldg_stream_f4(reinterpret_cast<const float4*>(p.X + (g + p.lag) * p.ld + c));
      }
      mk = w;
    };
    auto drain = [&](bool first) {
      const int quarter = warp & 3, which = (warp >> 2) & 1, part = warp >> 3, nparts = kGtProdWarps / 8;
      double* dst = ws + (size_t)which * d * 128 + quarter * 32 + lane;
      for (int c0 = part * 32; c0 < d; c0 += 32 * nparts) {
        float v[32];
        tc::tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(which * 256 + c0), v);
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          double* e = dst + (size_t)(c0 + q) * 128;
          *e = first ? (double)v[q] : (*e + (double)v[q]);
        }
      }
    };
    const uint32_t o = tc::off_mnmajor_sw128b32(lane * 4, fr, lbo, sbo);

    if (n_stages > 0) prefetch(0);
    int n_windows_done = 0;
    for (int s = 0; s < n_stages; ++s) {
      const int slot = s % kGpStages;
      const uint32_t use = (uint32_t)(s / kGpStages);
      const float4 ca = xa, cb = xb;
      const int w = mk;
      if (s + 1 < n_stages) prefetch(s + 1);
      float dith;   // one dither per frame, see gram_tc_kernel
      {
        uint32_t h = (uint32_t)(g_begin + (int64_t)s * kGtBK + fr) * 0x9E3779B1u;
        h ^= h >> 15;
        h *= 0x85EBCA6Bu;
        h ^= h >> 13;
        dith = __uint_as_float((h >> 9) | 0x3f800000u) - 1.5f;
      }
      float z[4];
      z[0] = gt_cond(ca.x, sh.x, sc.x);
      z[1] = gt_cond(ca.y, sh.y, sc.y);
      z[2] = gt_cond(ca.z, sh.z, sc.z);
      z[3] = gt_cond(ca.w, sh.w, sc.w);
      if (p.mode == 1) {
        z[0] -= gt_cond(cb.x, sh.x, sc.x);
        z[1] -= gt_cond(cb.y, sh.y, sc.y);
        z[2] -= gt_cond(cb.z, sh.z, sc.z);
        z[3] -= gt_cond(cb.w, sh.w, sc.w);
      }
      float a[4], rh[4], rl[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float zz = w ? z[q] : 0.f;
        a[q] = fminf(fmaxf(rintf(fmaf(zz, 8.0f, dith)), -64.0f), 64.0f) * 0.125f;
        const float r = zz - a[q];
        rh[q] = gt_tf32(r);
        rl[q] = gt_tf32(r - rh[q]);
      }
      const float wf = (float)w, w2 = 2.0f * wf;
      mbar_wait(&B->empty[slot], (use & 1u) ^ 1u);
      unsigned char* T = tiles + (size_t)slot * kGpStageBytes + o;
      *reinterpret_cast<float4*>(T) = make_float4(a[0], a[1], a[2], a[3]);
      *reinterpret_cast<float4*>(T + kGpTile) = make_float4(rh[0], rh[1], rh[2], rh[3]);
      *reinterpret_cast<float4*>(T + 2 * kGpTile) = make_float4(rl[0], rl[1], rl[2], rl[3]);
      *reinterpret_cast<float4*>(T + 3 * kGpTile) = make_float4(w2 * a[0], w2 * a[1], w2 * a[2], w2 * a[3]);
      *reinterpret_cast<float4*>(T + 4 * kGpTile) = make_float4(wf * rh[0], wf * rh[1], wf * rh[2], wf * rh[3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_cluster(&B->full[slot], 0u);

      const bool window_end = ((s + 1) % stages_per_window == 0) || (s + 1 == n_stages);
      if (window_end) {
        mbar_wait(&B->wdone, (uint32_t)(n_windows_done & 1));
        tc::fence_after_sync();
        drain(n_windows_done == 0);
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive_cluster(&B->drained, 0u);
        ++n_windows_done;
      }
    }
    if (n_stages == 0) {
      for (int i = tid; i < 2 * d * 128; i += kGtProdWarps * 32) ws[i] = 0.0;
    }
  } else if (rank == 0) {
    // ============================================================ MMA issuer of the pair (leader CTA)
    const uint32_t idesc = tc::idesc_tf32(256, d, 1, 1);
    const uint32_t base = smem_u32(tiles);
    int n_windows_done = 0;
    for (int s = 0; s < n_stages; ++s) {
      const int slot = s % kGpStages;
      const uint32_t use = (uint32_t)(s / kGpStages);
      const bool window_start = (s % stages_per_window) == 0;
      if (window_start && n_windows_done > 0) {
        tc::mbar_wait_cluster(&B->drained, (uint32_t)((n_windows_done - 1) & 1));
        tc::fence_after_sync();
      }
      tc::mbar_wait_cluster(&B->full[slot], use & 1u);
      tc::fence_after_sync();
      const uint32_t T = base + (uint32_t)slot * kGpStageBytes;
#pragma unroll
      for (int ks = 0; ks < kGtBK / 8; ++ks) {
        const uint32_t off = (uint32_t)ks * 2u * sbo;
        const uint64_t dBa = tc::smem_desc(T + off, lbo, sbo, tc::kLayoutSw128Base32);
        const uint64_t dBh = tc::smem_desc(T + kGpTile + off, lbo, sbo, tc::kLayoutSw128Base32);
        const uint64_t dBl = tc::smem_desc(T + 2 * kGpTile + off, lbo, sbo, tc::kLayoutSw128Base32);
        const uint64_t dA2a = tc::smem_desc(T + 3 * kGpTile + off, lbo, sbo, tc::kLayoutSw128Base32);
        const uint64_t dAwh = tc::smem_desc(T + 4 * kGpTile + off, lbo, sbo, tc::kLayoutSw128Base32);
        const uint32_t acc = (window_start && ks == 0) ? 0u : 1u;
        tc::mma_tf32_pair_elect(tmem, dA2a, dBa, idesc, acc);
        tc::mma_tf32_pair_elect(tmem + 256u, dA2a, dBh, idesc, acc);
        tc::mma_tf32_pair_elect(tmem + 256u, dA2a, dBl, idesc, 1u);
        tc::mma_tf32_pair_elect(tmem + 256u, dAwh, dBh, idesc, 1u);
      }
      tc::mma_commit_pair_elect(&B->empty[slot], 3u);
      const bool window_end = ((s + 1) % stages_per_window == 0) || (s + 1 == n_stages);
      if (window_end) {
        tc::mma_commit_pair_elect(&B->wdone, 3u);
        ++n_windows_done;
      }
      __syncwarp();
    }
  }

  tc::fence_before_sync();
  tc::cluster_sync();
  if (warp == kGtProdWarps) tc::tmem_dealloc_pair(tmem, 512);
}

// Pf[i][j], Qf[i][j] = sum over frame chunks (fixed order) of the per-CTA fp64 partials
__global__ void gram_tc_reduce_kernel(const double* __restrict__ ws, int n_chunks, int d, double* __restrict__ PQ) {
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
__global__ void gram_tc_combine_kernel(const double* __restrict__ PQ, int d, double* __restrict__ G) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d * d) return;
  const int i = e / d, j = e - i * d;
  const double* P = PQ;
  const double* Q = PQ + (size_t)d * d;
  // P is symmetric by construction (exact integer-valued sums); use the (min, max) entry so that G is
  // bit-symmetric as well
  const int a = i < j ? i : j, b = i < j ? j : i;
  G[e] = 0.5 * P[(size_t)a * d + b] + 0.5 * (Q[(size_t)a * d + b] + Q[(size_t)b * d + a]);
}

static inline int gt_row_blocks(int d) { return (d + 127) / 128; }
static inline int gt_chunks(int d) { return kNumSMs / gt_row_blocks(d); }

bool gram_tcgen05_supported(int d, int64_t ld, const float* X) {
  return d >= 32 && d <= 256 && d % 32 == 0 && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0;
}
size_t gram_tcgen05_ws_bytes(int d) {
  if (d < 32 || d > 256 || d % 32 != 0) return 0;
  return ((size_t)gt_row_blocks(d) * gt_chunks(d) * 2 * d * 128 + (size_t)2 * d * d) * sizeof(double);
}

int gram_tcgen05(const float* X, int64_t n, int d, int64_t ld, const uint8_t* mask, int lag, int mode,
                 const float* shift, const float* scale, double* G, void* ws, size_t ws_bytes, cudaStream_t st,
                 bool cta_pairs) {
  PMB_REQUIRE(ws_bytes >= gram_tcgen05_ws_bytes(d), "pmb_gram: workspace too small for the tcgen05 path");
  PMB_REQUIRE((reinterpret_cast<uintptr_t>(shift) & 15) == 0 && (reinterpret_cast<uintptr_t>(scale) & 15) == 0,
              "pmb_gram: shift/scale must be 16-byte aligned for the tcgen05 path");
  GramTcParams p;
  p.X = X; p.n = n; p.d = d; p.ld = ld; p.mask = mask; p.lag = lag; p.mode = mode; p.shift = shift; p.scale = scale;
  p.ws = static_cast<double*>(ws);
  const int nrb = gt_row_blocks(d);
  int nc = gt_chunks(d);
  const bool pair = d == 256 && cta_pairs;
  const size_t smem_pair = (size_t)kGpStages * kGpStageBytes + sizeof(GpBars) + 1024;
  if (pair) {
    // One cluster (CTA pair) per frame chunk, all of them resident at once: a pair needs both SMs of a TPC,
    // and a part with disabled SMs has fewer usable pairs than SMs / 2.
    static int max_clusters = -1;
    if (max_clusters < 0) {
      PMB_CUDA(cudaFuncSetAttribute(gram_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pair));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * nc);
      cfg.blockDim = dim3(kGtThreads);
      cfg.dynamicSmemBytes = smem_pair;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int ncl = 0;
      PMB_CUDA(cudaOccupancyMaxActiveClusters(&ncl, gram_tc_pair_kernel, &cfg));
      max_clusters = ncl;
    }
    PMB_REQUIRE(max_clusters >= 1, "pmb_gram: the CTA-pair kernel does not fit on this device");
    if (nc > max_clusters) nc = max_clusters;
  }
  int64_t chunk = (n + nc - 1) / nc;
  chunk = ((chunk + kGtBK - 1) / kGtBK) * kGtBK;
  p.n_chunks = nc;
  p.chunk = chunk;
  if (pair) {
    gram_tc_pair_kernel<<<2 * nc, kGtThreads, smem_pair, st>>>(p);   // clusters of 2: (row block 0, row block 1) of a chunk
  } else {
    const size_t smem = (size_t)kGtStages * kGtStageBytes + sizeof(GtBars) + 1024;
    PMB_CUDA(cudaFuncSetAttribute(gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gram_tc_kernel<<<nrb * nc, kGtThreads, smem, st>>>(p);
  }
  PMB_LAUNCH_CHECK();
  double* PQ = p.ws + (size_t)nrb * nc * 2 * d * 128;
  gram_tc_reduce_kernel<<<(d * d + 255) / 256, 256, 0, st>>>(p.ws, nc, d, PQ);
  PMB_LAUNCH_CHECK();
  gram_tc_combine_kernel<<<(d * d + 255) / 256, 256, 0, st>>>(PQ, d, G);
  PMB_LAUNCH_CHECK();
  return PMB_OK;
}

}  // namespace pmb
