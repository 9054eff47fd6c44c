// moments.cu -- K2: NaN-aware column moments in fp64 + the lag pair mask.
//
// One streaming pass over X (4*d bytes/frame), HBM-bound.  Threads are laid
// out (rows x columns) so that a warp always reads consecutive floats; each
// thread keeps fp64 partial sums for ONE column, CTA partials go to a
// workspace and a second tiny kernel folds them in a fixed order, so the
// result is bit-reproducible run to run.
#include "common.cuh"

namespace pmb {

constexpr int kMomThreads = 256;
constexpr int kMomGrid = kNumSMs * 4;

__global__ void __launch_bounds__(kMomThreads) col_moments_kernel(
    const float* __restrict__ X, int64_t n, int d, int64_t ld, const uint8_t* __restrict__ mask,
    const double* __restrict__ shift_in, int cw, double* __restrict__ part /* [grid.y][grid.x][5][cw] */) {
  __shared__ double red[kMomThreads];
  const int tid = threadIdx.x;
  const int rpp = kMomThreads / cw;  // rows per pass
  const int r = tid / cw, cl = tid - r * cw;
  const int c = blockIdx.y * cw + cl;
  const bool col_ok = c < d;
  double shift = 0.0;
  if (col_ok) {
    if (shift_in != nullptr) {
      shift = shift_in[c];
    } else {
      float x0 = X[c];
      shift = (x0 == x0) ? (double)x0 : 0.0;
    }
  }
  double cnt = 0, s1 = 0, s2 = 0, es1 = 0, ecnt = 0;
  constexpr int U = 8;
  const int64_t stride = (int64_t)gridDim.x * rpp * U;
  for (int64_t row0 = (int64_t)blockIdx.x * rpp * U + r; row0 < n; row0 += stride) {
    float v[U];
    int e[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = row0 + (int64_t)u * rpp;
      const bool ok = col_ok && row < n;
      v[u] = ok ? ldg_stream_f(X + row * ld + c) : __int_as_float(0x7fc00000);
      e[u] = (ok && mask) ? 2 - __popc(mask[row] & 3) : 0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (v[u] == v[u]) {
        const double dx = (double)v[u] - shift;
        cnt += 1.0;
        s1 += dx;
        s2 = fma(dx, dx, s2);
        es1 = fma((double)e[u], dx, es1);
        ecnt += (double)e[u];
      }
    }
  }
  // fold the row lanes of this CTA (fixed order)
  double vals[5] = {cnt, s1, s2, es1, ecnt};
  double* dst = part + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 5 * cw;
  for (int q = 0; q < 5; ++q) {
    red[tid] = vals[q];
    __syncthreads();
    if (r == 0) {
      double acc = 0.0;
      for (int rr = 0; rr < rpp; ++rr) acc += red[rr * cw + cl];
      dst[q * cw + cl] = acc;
    }
    __syncthreads();
  }
}

// float4 path (d % 4 == 0, 16-byte aligned rows): a thread owns FOUR adjacent columns, so a frame costs
// one 16-byte load per thread instead of four 4-byte ones and the mask byte is shared by four columns;
// counts are integers and the edge sums are only touched on the (rare) edge frames.
constexpr int kMomU = 4;   // rows in flight per thread

__global__ void __launch_bounds__(kMomThreads, 3) col_moments_v4_kernel(
    const float* __restrict__ X, int64_t n, int d, int64_t ld, const uint8_t* __restrict__ mask,
    const double* __restrict__ shift_in, int cw, double* __restrict__ part /* [grid.x][5][cw] */) {
  extern __shared__ double red[];   // [rpp][5][d]
  const int tid = threadIdx.x;
  const int tpr = d >> 2;                 // threads per row
  const int rpp = kMomThreads / tpr;      // rows per pass
  const int r = tid / tpr, cg = tid - r * tpr, c = cg * 4;
  const bool active = r < rpp;
  double shift[4] = {0.0, 0.0, 0.0, 0.0};
  if (active) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (shift_in != nullptr) {
        shift[q] = shift_in[c + q];
      } else {
        const float x0 = X[c + q];
        shift[q] = (x0 == x0) ? (double)x0 : 0.0;
      }
    }
  }
  int cnt[4] = {0, 0, 0, 0}, ecnt[4] = {0, 0, 0, 0};
  double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0}, es1[4] = {0, 0, 0, 0};
  const int64_t stride = (int64_t)gridDim.x * rpp * kMomU;
  if (active) {
    for (int64_t row0 = (int64_t)blockIdx.x * rpp * kMomU + r; row0 < n; row0 += stride) {
      float4 v[kMomU];
      int e[kMomU];
#pragma unroll
      for (int u = 0; u < kMomU; ++u) {
        const int64_t row = row0 + (int64_t)u * rpp;
        const bool ok = row < n;
        const float qnan = __int_as_float(0x7fc00000);
        v[u] = ok ? ldg_stream_f4(reinterpret_cast<const float4*>(X + row * ld + c)) : make_float4(qnan, qnan, qnan, qnan);
        e[u] = (ok && mask) ? 2 - __popc(mask[row] & 3) : 0;
      }
#pragma unroll
      for (int u = 0; u < kMomU; ++u) {
        const float x[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (x[q] == x[q]) {
            const double dx = (double)x[q] - shift[q];
            cnt[q] += 1;
            s1[q] += dx;
            s2[q] = fma(dx, dx, s2[q]);
            if (e[u]) {
              es1[q] = fma((double)e[u], dx, es1[q]);
              ecnt[q] += e[u];
            }
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double* dst = red + (size_t)r * 5 * d + c + q;
      dst[0 * d] = (double)cnt[q];
      dst[1 * d] = s1[q];
      dst[2 * d] = s2[q];
      dst[3 * d] = es1[q];
      dst[4 * d] = (double)ecnt[q];
    }
  }
  __syncthreads();
  double* out = part + (size_t)blockIdx.x * 5 * cw;
  for (int i = tid; i < 5 * d; i += kMomThreads) {
    const int q = i / d, col = i - q * d;
    double acc = 0.0;
    for (int rr = 0; rr < rpp; ++rr) acc += red[(size_t)rr * 5 * d + i];   // fixed order
    out[q * cw + col] = acc;
  }
}

// One WARP per column: the lanes split the CTA partials (fixed assignment, fixed shuffle tree: the result
// is bit-reproducible), so the dependent chain is gx / 32 L2 loads instead of gx.
__global__ void col_moments_reduce_kernel(const float* __restrict__ X, const double* __restrict__ shift_in,
                                          int d, int cw, int gx, const double* __restrict__ part,
                                          double* __restrict__ out) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= d) return;
  const int by = c / cw, cl = c - by * cw;
  double acc[5] = {0, 0, 0, 0, 0};
  for (int b = lane; b < gx; b += 32) {
    const double* src = part + ((size_t)by * gx + b) * 5 * cw;
#pragma unroll
    for (int q = 0; q < 5; ++q) acc[q] += src[q * cw + cl];
  }
#pragma unroll
  for (int q = 0; q < 5; ++q) acc[q] = warp_sum(acc[q]);
  if (lane != 0) return;
  double sh;
  if (shift_in != nullptr) {
    sh = shift_in[c];
  } else {
    float x0 = X[c];
    sh = (x0 == x0) ? (double)x0 : 0.0;
  }
  out[0 * d + c] = acc[0];
  out[1 * d + c] = sh;
  out[2 * d + c] = acc[1];
  out[3 * d + c] = acc[2];
  out[4 * d + c] = acc[3];
  out[5 * d + c] = acc[4];
}

__device__ __forceinline__ int find_segment(const int64_t* __restrict__ off, int n_seg, int64_t g) {
  // largest s with off[s] <= g, or -1
  int lo = 0, hi = n_seg;  // off has n_seg+1 entries
  if (g < off[0] || g >= off[n_seg]) return -1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= g) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void pair_mask_kernel(const int64_t* __restrict__ off, int n_seg, int64_t n, int lag,
                                 uint8_t* __restrict__ mask) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  uint8_t m = 0;
  const int s = find_segment(off, n_seg, g);
  if (s >= 0) {
    const int64_t t = g - off[s], L = off[s + 1] - off[s];
    if (t + lag < L) m |= 1;
    if (t >= lag) m |= 2;
  }
  mask[g] = m;
}

static inline int col_width(int d) {
  int cw = 1;
  while (cw < d && cw < kMomThreads) cw <<= 1;
  return cw;
}

}  // namespace pmb

extern "C" size_t pmb_col_moments_ws_bytes(int d) {
  if (d <= 0) return 0;
  const int cw = pmb::col_width(d);
  const int gy = (d + cw - 1) / cw;
  return (size_t)gy * (pmb::kMomGrid / gy + 1) * 5 * cw * sizeof(double);
}

extern "C" int pmb_col_moments(const float* X, int64_t n, int d, int64_t ld, const uint8_t* mask,
                               const double* shift_in, double* out, void* ws, size_t ws_bytes,
                               pmb_stream_t stream) {
  using namespace pmb;
  PMB_REQUIRE(n > 0 && d > 0 && ld >= d, "pmb_col_moments: bad sizes n=%lld d=%d ld=%lld",
              (long long)n, d, (long long)ld);
  PMB_REQUIRE(X && out && ws, "pmb_col_moments: null pointer");
  if (ws_bytes < pmb_col_moments_ws_bytes(d)) {
    set_error("pmb_col_moments: workspace too small (%zu < %zu)", ws_bytes, pmb_col_moments_ws_bytes(d));
    return PMB_EWORKSPACE;
  }
  const int cw = col_width(d);
  const int gy = (d + cw - 1) / cw;
  int gx = kMomGrid / gy;
  if (gx < 1) gx = 1;
  const int rpp = kMomThreads / cw;
  const int64_t groups = (n + (int64_t)rpp * 8 - 1) / ((int64_t)rpp * 8);
  if (gx > groups) gx = (int)groups;
  const bool v4 = (d % 4 == 0) && d >= 16 && d <= 4 * kMomThreads && (ld % 4 == 0) &&
                  ((reinterpret_cast<uintptr_t>(X) & 15) == 0) && gy == 1;
  if (v4) {
    const int rpp4 = kMomThreads / (d / 4);
    const int64_t groups4 = (n + (int64_t)rpp4 * kMomU - 1) / ((int64_t)rpp4 * kMomU);
    gx = 3 * kNumSMs;   // 80 registers x 256 threads: three resident CTAs per SM, one wave
    if (gx > groups4) gx = (int)groups4;
    const size_t smem = (size_t)rpp4 * 5 * d * sizeof(double);
    PMB_CUDA(cudaFuncSetAttribute(col_moments_v4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    col_moments_v4_kernel<<<gx, kMomThreads, smem, as_stream(stream)>>>(X, n, d, ld, mask, shift_in, cw,
                                                                        static_cast<double*>(ws));
    PMB_LAUNCH_CHECK();
  } else {
    dim3 grid(gx, gy);
    col_moments_kernel<<<grid, kMomThreads, 0, as_stream(stream)>>>(X, n, d, ld, mask, shift_in, cw,
                                                                   static_cast<double*>(ws));
    PMB_LAUNCH_CHECK();
  }
  col_moments_reduce_kernel<<<(d * 32 + 255) / 256, 256, 0, as_stream(stream)>>>(
      X, shift_in, d, cw, gx, static_cast<const double*>(ws), out);
  PMB_LAUNCH_CHECK();
  return PMB_OK;
}

extern "C" int pmb_pair_mask(const int64_t* seg_offsets, int n_seg, int64_t n, int lag,
                             uint8_t* mask, pmb_stream_t stream) {
  using namespace pmb;
  PMB_REQUIRE(n >= 0 && n_seg >= 1 && lag >= 0, "pmb_pair_mask: bad sizes");
  if (n == 0) return PMB_OK;
  PMB_REQUIRE(seg_offsets && mask, "pmb_pair_mask: null pointer");
  const int64_t blocks = (n + 255) / 256;
  pair_mask_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(seg_offsets, n_seg, n, lag, mask);
  PMB_LAUNCH_CHECK();
  return PMB_OK;
}
