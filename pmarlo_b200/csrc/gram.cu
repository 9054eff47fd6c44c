// gram.cu -- K3 (SIMT path): weighted symmetric Gram matrices for TICA.
//
//   mode 0: G = sum_g w_g z_g z_g^T,  w_g = popcount(mask_g & 3)      (X0^T X0 + Xt^T Xt)
//   mode 1: G = sum_g (mask_g & 1) v_g v_g^T,  v_g = z_g - z_{g+lag}  (-> X0^T Xt + Xt^T X0)
//   z = NaN ? 0 : (x - shift) * scale        (fp32 conditioning, undone exactly in fp64 later)
//
// Precision plan (north star: covariances within 1e-6 of the fp64 oracle, and
// TICA eigenvectors amplify covariance error by 1/gap): products are fp32 FMAs
// of conditioned (|z| ~ 1) data, accumulated in fp32 registers over only
// kFoldFrames = 32 frames starting from zero, then folded into fp64 REGISTER
// accumulators (one F2F + DADD per 32 FFMAs); each CTA writes its fp64 tile once
// and a second kernel adds the CTA tiles in a fixed order (deterministic).
//
// Tiling: 128x128 output tile per CTA (only tiles on/above the diagonal), 8x8
// micro-tile per thread, 8-frame shared-memory stages with register prefetch,
// frames split into chunks along grid.y (split-K) so that the grid fills 148 SMs.
// This is the path for small/odd d and the numerical reference of the tcgen05 path.
#include "common.cuh"

namespace pmb {

constexpr int kGT = 128;            // tile edge
constexpr int kGBK = 8;             // frames per stage
constexpr int kGThreads = 256;
constexpr int kFoldFrames = 32;     // fp32 accumulation extent before the fp64 fold
constexpr int kGramTargetCtas = 2 * kNumSMs;

struct GramParams {
  const float* X;
  int64_t n;
  int d;
  int64_t ld;
  const uint8_t* mask;
  int lag;
  const float* shift;
  const float* scale;
  double* part;   // [n_tiles][n_chunks][128*128]
  int n_chunks;
  int64_t chunk;  // frames per chunk (multiple of kGBK)
  int vec_ok;     // float4 loads legal
  const unsigned int* run_if;   // nullptr, or a device flag: the kernels return at once while it is 0 (gram_h.cu's
                                // out-of-range fallback)
};

__device__ __forceinline__ float4 load4(const float* __restrict__ X, int64_t row, int64_t ld, int c,
                                        int d, int vec_ok) {
  if (vec_ok && c + 3 < d) return ldg_stream_f4(reinterpret_cast<const float4*>(X + row * ld + c));
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* p = X + row * ld;
  if (c + 0 < d) v.x = p[c + 0];
  if (c + 1 < d) v.y = p[c + 1];
  if (c + 2 < d) v.z = p[c + 2];
  if (c + 3 < d) v.w = p[c + 3];
  return v;
}
__device__ __forceinline__ float cond1(float x, float sh, float sc) {
  return (x == x) ? (x - sh) * sc : 0.0f;
}
__device__ __forceinline__ float4 cond4(float4 x, float4 sh, float4 sc) {
  return make_float4(cond1(x.x, sh.x, sc.x), cond1(x.y, sh.y, sc.y), cond1(x.z, sh.z, sc.z),
                     cond1(x.w, sh.w, sc.w));
}

template <int MODE>
__global__ void __launch_bounds__(kGThreads, 1) gram_simt_kernel(GramParams p) {
  __shared__ __align__(16) float As[2][kGBK][kGT];
  __shared__ __align__(16) float Bs[2][kGBK][kGT];
  if (p.run_if != nullptr && *p.run_if == 0u) return;

  // upper-triangular tile enumeration
  const int nb = (p.d + kGT - 1) / kGT;
  int ti = 0, rem = blockIdx.x;
  while (rem >= nb - ti) { rem -= nb - ti; ++ti; }
  const int tj = ti + rem;
  const bool diag = (ti == tj);
  const int i0 = ti * kGT, j0 = tj * kGT;

  const int tid = threadIdx.x;
  const int fr = tid >> 5;             // frame within stage
  const int c4 = (tid & 31) * 4;       // column within tile (x4)
  const int tx = tid & 15, ty = tid >> 4;

  // per-thread conditioning constants for its 4 load columns (A block / B block)
  float4 shA, scA, shB, scB;
  {
    float s[8], c[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ca = i0 + c4 + q, cb = j0 + c4 + q;
      s[q] = ca < p.d ? p.shift[ca] : 0.f;  c[q] = ca < p.d ? p.scale[ca] : 0.f;
      s[4 + q] = cb < p.d ? p.shift[cb] : 0.f;  c[4 + q] = cb < p.d ? p.scale[cb] : 0.f;
    }
    shA = make_float4(s[0], s[1], s[2], s[3]);  scA = make_float4(c[0], c[1], c[2], c[3]);
    shB = make_float4(s[4], s[5], s[6], s[7]);  scB = make_float4(c[4], c[5], c[6], c[7]);
  }

  const int64_t g_begin = (int64_t)blockIdx.y * p.chunk;
  int64_t g_end = g_begin + p.chunk;
  if (g_end > p.n) g_end = p.n;
  const int n_stages = g_begin < g_end ? (int)((g_end - g_begin + kGBK - 1) / kGBK) : 0;

  float acc[8][8];
  double dacc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[i][j] = 0.f; dacc[i][j] = 0.0; }

  double* P = p.part + ((size_t)blockIdx.x * p.n_chunks + blockIdx.y) * (kGT * kGT);

  float4 pa, pb;
  float pw;
  auto prefetch = [&](int s) {
    const int64_t g = g_begin + (int64_t)s * kGBK + fr;
    pa = make_float4(0.f, 0.f, 0.f, 0.f);
    pb = pa;
    pw = 0.f;
    if (g < g_end) {
      const int m = p.mask[g];
      const int w = (MODE == 0) ? __popc(m & 3) : (m & 1);
      if (w) {
        pw = (float)w;
        pa = cond4(load4(p.X, g, p.ld, i0 + c4, p.d, p.vec_ok), shA, scA);
        if (!diag) pb = cond4(load4(p.X, g, p.ld, j0 + c4, p.d, p.vec_ok), shB, scB);
        if (MODE == 1) {
          float4 qa = cond4(load4(p.X, g + p.lag, p.ld, i0 + c4, p.d, p.vec_ok), shA, scA);
          pa = make_float4(pa.x - qa.x, pa.y - qa.y, pa.z - qa.z, pa.w - qa.w);
          if (!diag) {
            float4 qb = cond4(load4(p.X, g + p.lag, p.ld, j0 + c4, p.d, p.vec_ok), shB, scB);
            pb = make_float4(pb.x - qb.x, pb.y - qb.y, pb.z - qb.z, pb.w - qb.w);
          }
        }
        if (diag) pb = pa;
      }
    }
  };
  auto stash = [&](int buf) {
    *reinterpret_cast<float4*>(&As[buf][fr][c4]) =
        make_float4(pa.x * pw, pa.y * pw, pa.z * pw, pa.w * pw);
    *reinterpret_cast<float4*>(&Bs[buf][fr][c4]) = pb;
  };
  auto fold = [&]() {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dacc[i][j] += (double)acc[i][j];
        acc[i][j] = 0.f;
      }
  };

  int buf = 0;
  if (n_stages > 0) {
    prefetch(0);
    stash(0);
  }
  __syncthreads();
  for (int s = 0; s < n_stages; ++s) {
    const bool has_next = (s + 1 < n_stages);
    if (has_next) prefetch(s + 1);
#pragma unroll
    for (int k = 0; k < kGBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (has_next) stash(buf ^ 1);
    __syncthreads();
    buf ^= 1;
    if (((s + 1) % (kFoldFrames / kGBK)) == 0 || !has_next) fold();
  }
  // one write of the CTA's fp64 tile (zeros for an empty chunk)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4);
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      double2* dst = reinterpret_cast<double2*>(P + row * kGT + jh * 64 + tx * 4);
      dst[0] = make_double2(dacc[i][jh * 4 + 0], dacc[i][jh * 4 + 1]);
      dst[1] = make_double2(dacc[i][jh * 4 + 2], dacc[i][jh * 4 + 3]);
    }
  }
}

// G[i][j] = sum over chunks (fixed order); lower triangle mirrored.
__global__ void gram_reduce_kernel(const double* __restrict__ part, int n_chunks, int d,
                                   double* __restrict__ G, const unsigned int* run_if) {
  if (run_if != nullptr && *run_if == 0u) return;
  const int nb = (d + kGT - 1) / kGT;
  int ti = 0, rem = blockIdx.x;
  while (rem >= nb - ti) { rem -= nb - ti; ++ti; }
  const int tj = ti + rem;
  const double* P = part + (size_t)blockIdx.x * n_chunks * (kGT * kGT);
  for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < kGT * kGT; e += gridDim.y * blockDim.x) {
    const int r = e / kGT, c = e - r * kGT;
    const int gi = ti * kGT + r, gj = tj * kGT + c;
    if (gi >= d || gj >= d) continue;
    double acc = 0.0;
    for (int q = 0; q < n_chunks; ++q) acc += P[(size_t)q * (kGT * kGT) + e];
    if (ti == tj) {
      G[(size_t)gi * d + gj] = acc;
    } else {
      G[(size_t)gi * d + gj] = acc;
      G[(size_t)gj * d + gi] = acc;
    }
  }
}

static inline int gram_tiles(int d) {
  const int nb = (d + kGT - 1) / kGT;
  return nb * (nb + 1) / 2;
}
static inline int gram_chunks(int d) {
  int c = kGramTargetCtas / gram_tiles(d);
  return c < 1 ? 1 : c;
}

int gram_tcgen05(const float* X, int64_t n, int d, int64_t ld, const uint8_t* mask, int lag,
                 int mode, const float* shift, const float* scale, double* G, void* ws,
                 size_t ws_bytes, cudaStream_t stream, bool cta_pairs);  // gram_tc.cu
bool gram_tcgen05_supported(int d, int64_t ld, const float* X);
size_t gram_tcgen05_ws_bytes(int d);
// gram_h.cu: the fp16-kind path (default on tensor cores)
bool gram_h_supported(int d, int64_t ld, const float* X, int64_t n);
size_t gram_h_ws_bytes(int d);
int gram_h(const float* X, int64_t n, int d, int64_t ld, const uint8_t* mask, int lag, int mode, const float* shift,
           const float* scale, double* G, void* ws, size_t ws_bytes, cudaStream_t st);
int gram_h_debug_counters(int64_t* out16);
constexpr int64_t kGramTcMinFrames = 65536;   // auto dispatch: below this the SIMT kernel is as fast

// The SIMT launches; with run_if != nullptr they execute only when *run_if != 0 on the device.
int gram_simt(const float* X, int64_t n, int d, int64_t ld, const uint8_t* mask, int lag, int mode, const float* shift,
              const float* scale, double* G, void* ws, cudaStream_t st, const unsigned int* run_if) {
  GramParams p;
  p.X = X; p.n = n; p.d = d; p.ld = ld; p.mask = mask; p.lag = lag; p.shift = shift; p.scale = scale;
  p.part = static_cast<double*>(ws);
  p.run_if = run_if;
  const int nt = gram_tiles(d);
  int nc = gram_chunks(d);
  int64_t chunk = (n + nc - 1) / nc;
  chunk = ((chunk + kFoldFrames - 1) / kFoldFrames) * kFoldFrames;
  nc = (int)((n + chunk - 1) / chunk);
  p.n_chunks = nc;
  p.chunk = chunk;
  p.vec_ok = ((reinterpret_cast<uintptr_t>(X) & 15) == 0) && (ld % 4 == 0);
  dim3 grid(nt, nc);
  if (mode == 0) gram_simt_kernel<0><<<grid, kGThreads, 0, st>>>(p);
  else           gram_simt_kernel<1><<<grid, kGThreads, 0, st>>>(p);
  PMB_LAUNCH_CHECK();
  dim3 rgrid(nt, 16);
  gram_reduce_kernel<<<rgrid, 256, 0, st>>>(p.part, nc, d, G, run_if);
  PMB_LAUNCH_CHECK();
  return PMB_OK;
}

}  // namespace pmb

extern "C" int pmb_debug_counters_gram(int64_t* out16) {
  using namespace pmb;
  PMB_REQUIRE(out16 != nullptr, "pmb_debug_counters_gram: null pointer");
  return gram_h_debug_counters(out16);
}

extern "C" size_t pmb_gram_ws_bytes(int d) {
  if (d <= 0) return 0;
  const size_t simt = (size_t)pmb::gram_tiles(d) * pmb::gram_chunks(d) * pmb::kGT * pmb::kGT * sizeof(double);
  size_t tc = pmb::gram_tcgen05_ws_bytes(d);
  if (pmb::gram_h_ws_bytes(d) > tc) tc = pmb::gram_h_ws_bytes(d);
  // the fp16 path keeps a 256-byte header in front of the region its SIMT fallback would use
  return (simt + 256 > tc ? simt + 256 : tc);
}

extern "C" int pmb_gram(const float* X, int64_t n, int d, int64_t ld, const uint8_t* mask, int lag,
                        int mode, const float* shift, const float* scale, double* G, void* ws,
                        size_t ws_bytes, int impl, pmb_stream_t stream) {
  using namespace pmb;
  PMB_REQUIRE(n > 0 && d > 0 && ld >= d && lag >= 0, "pmb_gram: bad sizes");
  PMB_REQUIRE(mode == 0 || mode == 1, "pmb_gram: mode must be 0 or 1");
  PMB_REQUIRE(X && mask && shift && scale && G && ws, "pmb_gram: null pointer");
  if (ws_bytes < pmb_gram_ws_bytes(d)) {
    set_error("pmb_gram: workspace too small (%zu < %zu)", ws_bytes, pmb_gram_ws_bytes(d));
    return PMB_EWORKSPACE;
  }
  // impl: 0 auto, 1 SIMT, 2 tensor cores (kind::tf32, gram_tc.cu), 4 the same with CTA pairs (d = 256; measured
  // slower), 5 tensor cores (kind::f16, gram_h.cu: what auto selects)
  if (impl == 5 || (impl == 0 && n >= kGramTcMinFrames && gram_h_supported(d, ld, X, n))) {
    if (!gram_h_supported(d, ld, X, n)) {
      set_error("pmb_gram: the fp16 tcgen05 path needs d %% 32 == 0, 32 <= d <= 256, 16B-aligned rows");
      return PMB_EUNSUPPORTED;
    }
    return gram_h(X, n, d, ld, mask, lag, mode, shift, scale, G, ws, ws_bytes, as_stream(stream));
  }
  if (impl == 2 || impl == 4 || (impl == 0 && n >= kGramTcMinFrames && gram_tcgen05_supported(d, ld, X))) {
    if (!gram_tcgen05_supported(d, ld, X)) {
      set_error("pmb_gram: tcgen05 path needs d %% 32 == 0, 32 <= d <= 256, 16B-aligned rows");
      return PMB_EUNSUPPORTED;
    }
    return gram_tcgen05(X, n, d, ld, mask, lag, mode, shift, scale, G, ws, ws_bytes, as_stream(stream), impl == 4);
  }
  return gram_simt(X, n, d, ld, mask, lag, mode, shift, scale, G, ws, as_stream(stream), nullptr);
}
