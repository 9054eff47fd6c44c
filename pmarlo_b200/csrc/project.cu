// project.cu -- K5: TICA projection  y = (x_imputed - a) W  (skinny GEMM, HBM-bound).
//
// Reads 4*d bytes and writes 4*m (or 8*m) bytes per frame.  Three kernels:
//  * project_warp_kernel (the pipeline's path: d <= 256, m <= 16, fp32 output): a warp per frame, the lane's
//    slice of W in registers, packed fp32 FMAs, recursive-halving reduction over the lanes;
//  * project_fast_kernel (wider d, fp32 output): a thread per frame, W in shared memory as fp32;
//  * project_kernel (fp64 output or m > 16): a CTA owns 256 consecutive frames; the feature axis is streamed
//    through shared memory in chunks of 32 columns (every warp load is one 128-byte line), stored with a
//    padded stride so that the thread-per-frame reads are bank-conflict free; arithmetic is fp64 (W, a in
//    shared memory, broadcast reads) so that the projection itself adds no error beyond the fp32 input.
#include "common.cuh"

namespace pmb {

constexpr int kPrjFrames = 256;
constexpr int kPrjChunk = 32;

template <int MP>
__global__ void __launch_bounds__(kPrjFrames) project_kernel(
    const float* __restrict__ X, int64_t n, int d, int64_t ld, const double* __restrict__ a,
    const double* __restrict__ nanfill, const double* __restrict__ W, int m, int ldw, int c_off,
    void* __restrict__ Y, int64_t ldy, int out_f64) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sW = reinterpret_cast<double*>(smem_raw);            // d x MP
  double* sa = sW + (size_t)d * MP;                            // d
  double* sfill = sa + d;                                      // d
  float* xs = reinterpret_cast<float*>(sfill + d);             // 256 x 33
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < d * MP; i += kPrjFrames) {
    const int j = i / MP, c = i - j * MP;
    sW[i] = (c_off + c < m) ? W[(size_t)j * ldw + c_off + c] : 0.0;
  }
  for (int i = tid; i < d; i += kPrjFrames) {
    sa[i] = a[i];
    sfill[i] = nanfill[i];
  }
  const int64_t f0 = (int64_t)blockIdx.x * kPrjFrames;
  double acc[MP];
#pragma unroll
  for (int c = 0; c < MP; ++c) acc[c] = 0.0;

  for (int j0 = 0; j0 < d; j0 += kPrjChunk) {
    __syncthreads();
    // warp w loads rows w, w+8, ...: 32 consecutive floats each
    for (int r = warp; r < kPrjFrames; r += kPrjFrames / 32) {
      const int64_t row = f0 + r;
      const int j = j0 + lane;
      float v = 0.f;
      if (row < n && j < d) v = ldg_stream_f(X + row * ld + j);
      xs[r * (kPrjChunk + 1) + lane] = v;
    }
    __syncthreads();
    const int jn = (d - j0) < kPrjChunk ? (d - j0) : kPrjChunk;
    for (int jj = 0; jj < jn; ++jj) {
      const float xv = xs[tid * (kPrjChunk + 1) + jj];
      const int j = j0 + jj;
      const double xd = ((xv == xv) ? (double)xv : sfill[j]) - sa[j];
      const double* w = sW + (size_t)j * MP;
#pragma unroll
      for (int c = 0; c < MP; ++c) acc[c] = fma(xd, w[c], acc[c]);
    }
  }
  const int64_t row = f0 + tid;
  if (row < n) {
    if (out_f64) {
      double* y = static_cast<double*>(Y) + row * ldy + c_off;
#pragma unroll
      for (int c = 0; c < MP; ++c)
        if (c_off + c < m) y[c] = acc[c];
    } else {
      float* y = static_cast<float*>(Y) + row * ldy + c_off;
#pragma unroll
      for (int c = 0; c < MP; ++c)
        if (c_off + c < m) y[c] = (float)acc[c];
    }
  }
}

// fp32-output fast path (the pipeline's Y): thread per frame, the frame's row is read straight from
// global memory as 32-byte pieces (every sector is consumed completely, so no shared-memory transpose and
// no barriers), W lives in shared memory as fp32 and is read with 16-byte broadcasts.  Precision: the
// row is centred in fp32 against a32 = float(a) (exact for values near the mean), products are
// accumulated in fp32 over 32 columns and then folded into fp64, and the fp64 remainder
// - sum_j (a_j - a32_j) W_jc is added once per frame.
constexpr int kPfThreads = 128;

template <int MP>   // outputs padded to a multiple of 4
__global__ void __launch_bounds__(kPfThreads) project_fast_kernel(
    const float* __restrict__ X, int64_t n, int d, int64_t ld, const double* __restrict__ a,
    const double* __restrict__ nanfill, const double* __restrict__ W, int m, int ldw,
    float* __restrict__ Y, int64_t ldy) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sW = reinterpret_cast<float*>(smem_raw);              // d x MP
  float* sa = sW + (size_t)d * MP;                             // d
  float* sfz = sa + d;                                         // d: imputed value, centred
  double* scorr = reinterpret_cast<double*>(sfz + d);          // MP
  const int tid = threadIdx.x;
  for (int i = tid; i < d * MP; i += kPfThreads) {
    const int j = i / MP, c = i - j * MP;
    sW[i] = (c < m) ? (float)W[(size_t)j * ldw + c] : 0.f;
  }
  for (int i = tid; i < d; i += kPfThreads) {
    const float a32 = (float)a[i];
    sa[i] = a32;
    sfz[i] = (float)(nanfill[i] - a[i]);
  }
  if (tid < MP) {
    double acc = 0.0;
    if (tid < m)
      for (int j = 0; j < d; ++j) acc = fma(a[j] - (double)(float)a[j], W[(size_t)j * ldw + tid], acc);
    scorr[tid] = -acc;
  }
  __syncthreads();
  for (int64_t row = (int64_t)blockIdx.x * kPfThreads + tid; row < n; row += (int64_t)gridDim.x * kPfThreads) {
    const float4* xr = reinterpret_cast<const float4*>(X + row * ld);
    double acc64[MP];
#pragma unroll
    for (int c = 0; c < MP; ++c) acc64[c] = scorr[c];
    for (int j0 = 0; j0 < d; j0 += 32) {
      float acc[MP];
#pragma unroll
      for (int c = 0; c < MP; ++c) acc[c] = 0.f;
      float4 xv[8];
#pragma unroll
      // default caching: a lane's eight 16-byte loads walk one 128-byte line, the second half of every
      // 32-byte sector must hit L1 instead of going back to L2
      for (int q = 0; q < 8; ++q) xv[q] = (j0 + 4 * q < d) ? __ldg(xr + (j0 >> 2) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int j = j0 + 4 * q;
        if (j >= d) break;
        const float4 av = *reinterpret_cast<const float4*>(sa + j);
        const float x[4] = {xv[q].x, xv[q].y, xv[q].z, xv[q].w};
        const float aa[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float z = (x[e] == x[e]) ? x[e] - aa[e] : sfz[j + e];
          const float4* w4 = reinterpret_cast<const float4*>(sW + (size_t)(j + e) * MP);
#pragma unroll
          for (int c4 = 0; c4 < MP / 4; ++c4) {
            const float4 w = w4[c4];
            acc[4 * c4 + 0] = fmaf(z, w.x, acc[4 * c4 + 0]);
            acc[4 * c4 + 1] = fmaf(z, w.y, acc[4 * c4 + 1]);
            acc[4 * c4 + 2] = fmaf(z, w.z, acc[4 * c4 + 2]);
            acc[4 * c4 + 3] = fmaf(z, w.w, acc[4 * c4 + 3]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < MP; ++c) acc64[c] += (double)acc[c];
    }
    float* y = Y + row * ldy;
#pragma unroll
    for (int c = 0; c < MP; ++c)
      if (c < m) y[c] = (float)acc64[c];
  }
}

template <int MP>
static int launch_project_fast(const float* X, int64_t n, int d, int64_t ld, const double* a, const double* nanfill,
                               const double* W, int m, float* Y, int64_t ldy, cudaStream_t st) {
  const size_t smem = ((size_t)d * MP + 2 * (size_t)d) * sizeof(float) + MP * sizeof(double) + 16;
  PMB_CUDA(cudaFuncSetAttribute(project_fast_kernel<MP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t blocks = (n + kPfThreads - 1) / kPfThreads;
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  project_fast_kernel<MP><<<(unsigned)blocks, kPfThreads, smem, st>>>(X, n, d, ld, a, nanfill, W, m, m, Y, ldy);
  PMB_LAUNCH_CHECK();
  return PMB_OK;
}

// Warp-per-frame path (d <= 256, fp32 output): a lane owns the same 4 J columns of every frame, so its
// slice of W (4 J x MPF fp32) and of the centring vector live in REGISTERS -- no shared memory, and a frame
// is two fully coalesced 512-byte loads instead of 64 strided ones (the thread-per-frame kernel above is
// bound by the L1 tag stage: 32 different lines per load instruction).  The kernel is bound by instruction
// issue, not by the FMA pipe, so the products run as packed fma.rn.f32x2 (FFMA2: two outputs per issue slot,
// each half rounded exactly like fmaf) over MPF = m rounded up to even outputs instead of a fixed 16.  The
// partial sums of a frame are reduced over the 32 lanes by recursive halving on a virtual width of 16 (15
// shuffles instead of 5 per output); slots known to be zero at compile time are skipped, and where only the
// upper half of a pair is zero both lanes simply add (no selects; the duplicate is never stored).
// Precision as in project_fast_kernel (fp32 products, pairwise fp32 sums, fp64 remainder of the centring
// added once).

__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

template <int J, int MPF, int kPwWarps, int kPwMinBlocks, int kPwFrames>   // kPwFrames: frames in flight per warp
__global__ void __launch_bounds__(kPwWarps * 32, kPwMinBlocks) project_warp_kernel(
    const float* __restrict__ X, int64_t n, int d, int64_t ld, const double* __restrict__ a,
    const double* __restrict__ nanfill, const double* __restrict__ W, int m, int ldw,
    float* __restrict__ Y, int64_t ldy) {
  static_assert(MPF % 2 == 0 && MPF >= 2 && MPF <= 16, "MPF: even, at most 16");
  const int lane = threadIdx.x & 31;
  const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  unsigned long long w2[J][4][MPF / 2];
  float a32[J][4], fz[J][4];
#pragma unroll
  for (int j = 0; j < J; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int col = (lane + 32 * j) * 4 + e;
      const bool ok = col < d;
      a32[j][e] = ok ? (float)a[col] : 0.f;
      fz[j][e] = ok ? (float)(nanfill[col] - a[col]) : 0.f;
#pragma unroll
      for (int c = 0; c < MPF; c += 2)
        w2[j][e][c / 2] = pack_f32x2((ok && c < m) ? (float)W[(size_t)col * ldw + c] : 0.f,
                                     (ok && c + 1 < m) ? (float)W[(size_t)col * ldw + c + 1] : 0.f);
    }
  // after the halving steps lane L holds output index oidx(L); its fp64 remainder -sum_j (a_j - a32_j) W_jc
  const int oidx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
  double corr = 0.0;
  if (oidx < m)
    for (int col = 0; col < d; ++col) corr = fma(a[col] - (double)(float)a[col], W[(size_t)col * ldw + oidx], corr);
  corr = -corr;

  for (int64_t f0 = gwarp * kPwFrames; f0 < n; f0 += nwarps * kPwFrames) {
    float4 xv[kPwFrames][J];
#pragma unroll
    for (int u = 0; u < kPwFrames; ++u)
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int f4 = lane + 32 * j;
        xv[u][j] = (f0 + u < n && f4 * 4 < d) ? ldg_stream_f4(reinterpret_cast<const float4*>(X + (f0 + u) * ld) + f4)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
    for (int u = 0; u < kPwFrames; ++u) {
      unsigned long long acc2[MPF / 2];
#pragma unroll
      for (int c = 0; c < MPF / 2; ++c) acc2[c] = 0ull;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const float x[4] = {xv[u][j].x, xv[u][j].y, xv[u][j].z, xv[u][j].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float z = (x[e] == x[e]) ? x[e] - a32[j][e] : fz[j][e];
          const unsigned long long z2 = pack_f32x2(z, z);
#pragma unroll
          for (int c = 0; c < MPF / 2; ++c) acc2[c] = ffma2(z2, w2[j][e][c], acc2[c]);
        }
      }
      float acc[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[c] = 0.f;
#pragma unroll
      for (int c = 0; c < MPF / 2; ++c) unpack_f32x2(acc2[c], acc[2 * c], acc[2 * c + 1]);
      // recursive halving: 16 -> 8 -> 4 -> 2 -> 1 slots per lane, then the two lanes of a pair add up
#pragma unroll
      for (int step = 0; step < 4; ++step) {
        const int off = 16 >> step, half = 8 >> step;
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
          if (i + half < MPF) {
            const float keep = up ? acc[i + half] : acc[i];
            const float give = up ? acc[i] : acc[i + half];
            acc[i] = keep + __shfl_xor_sync(0xffffffffu, give, off);
          } else if (i < MPF) {
            acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);   // slot i + half is identically zero
          }
        }
      }
      float r = acc[0] + __shfl_xor_sync(0xffffffffu, acc[0], 1);
      if (f0 + u < n && (lane & 1) == 0 && oidx < m) Y[(f0 + u) * ldy + oidx] = (float)((double)r + corr);
    }
  }
}

template <int J, int MPF, int kPwWarps, int kPwMinBlocks, int kPwFrames>
static int launch_project_warp_cfg(const float* X, int64_t n, int d, int64_t ld, const double* a, const double* nanfill,
                                   const double* W, int m, float* Y, int64_t ldy, cudaStream_t st) {
  int64_t blocks = (n + (int64_t)kPwWarps * kPwFrames - 1) / ((int64_t)kPwWarps * kPwFrames);
  if (blocks > (int64_t)kPwMinBlocks * kNumSMs) blocks = (int64_t)kPwMinBlocks * kNumSMs;
  project_warp_kernel<J, MPF, kPwWarps, kPwMinBlocks, kPwFrames>
      <<<(unsigned)blocks, kPwWarps * 32, 0, st>>>(X, n, d, ld, a, nanfill, W, m, m, Y, ldy);
  PMB_LAUNCH_CHECK();
  return PMB_OK;
}

template <int J, int MPF>
static int launch_project_warp_mpf(const float* X, int64_t n, int d, int64_t ld, const double* a, const double* nanfill,
                                   const double* W, int m, float* Y, int64_t ldy, cudaStream_t st) {
  // Measured at 10 M x 256 -> 10 (tools/prj_bench.py): 8 warps x 4 frames in flight at 226 registers 2.86 ms,
  // 3 CTAs of 4 warps at 168 registers 2.70 ms, 2 frames in flight 3.8 ms (bytes in flight bound it); staging
  // the rows through a bulk-copy ring in shared memory gave 2.77 ms for 2..6 stages (issue-bound by then).
  // The 168-register form spills for the wide slices, which stay on one CTA of 8 warps.
  if constexpr (J * MPF <= 20) return launch_project_warp_cfg<J, MPF, 4, 3, 4>(X, n, d, ld, a, nanfill, W, m, Y, ldy, st);
  else
  return launch_project_warp_cfg<J, MPF, 8, 1, 4>(X, n, d, ld, a, nanfill, W, m, Y, ldy, st);
}

template <int J>
static int launch_project_warp(const float* X, int64_t n, int d, int64_t ld, const double* a, const double* nanfill,
                               const double* W, int m, float* Y, int64_t ldy, cudaStream_t st) {
  switch ((m + 1) / 2) {
    case 1: return launch_project_warp_mpf<J, 2>(X, n, d, ld, a, nanfill, W, m, Y, ldy, st);
    case 2: return launch_project_warp_mpf<J, 4>(X, n, d, ld, a, nanfill, W, m, Y, ldy, st);
    case 3: return launch_project_warp_mpf<J, 6>(X, n, d, ld, a, nanfill, W, m, Y, ldy, st);
    case 4: return launch_project_warp_mpf<J, 8>(X, n, d, ld, a, nanfill, W, m, Y, ldy, st);
    case 5: return launch_project_warp_mpf<J, 10>(X, n, d, ld, a, nanfill, W, m, Y, ldy, st);
    case 6: return launch_project_warp_mpf<J, 12>(X, n, d, ld, a, nanfill, W, m, Y, ldy, st);
    case 7: return launch_project_warp_mpf<J, 14>(X, n, d, ld, a, nanfill, W, m, Y, ldy, st);
    default: return launch_project_warp_mpf<J, 16>(X, n, d, ld, a, nanfill, W, m, Y, ldy, st);
  }
}

template <int MP>
static int launch_project(const float* X, int64_t n, int d, int64_t ld, const double* a,
                          const double* nanfill, const double* W, int m, int c_off, void* Y,
                          int64_t ldy, int out_f64, cudaStream_t st) {
  const size_t smem = ((size_t)d * MP + 2 * (size_t)d) * sizeof(double) +
                      (size_t)kPrjFrames * (kPrjChunk + 1) * sizeof(float);
  if (smem > 220 * 1024) {
    set_error("pmb_project: d=%d too large for shared-memory staging", d);
    return PMB_EUNSUPPORTED;
  }
  PMB_CUDA(cudaFuncSetAttribute(project_kernel<MP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = (n + kPrjFrames - 1) / kPrjFrames;
  project_kernel<MP><<<(unsigned)blocks, kPrjFrames, smem, st>>>(X, n, d, ld, a, nanfill, W, m, m,
                                                               c_off, Y, ldy, out_f64);
  PMB_LAUNCH_CHECK();
  return PMB_OK;
}

}  // namespace pmb

extern "C" int pmb_project(const float* X, int64_t n, int d, int64_t ld, const double* a,
                           const double* nanfill, const double* W, int m, void* Y, int64_t ldy,
                           int out_f64, pmb_stream_t stream) {
  using namespace pmb;
  PMB_REQUIRE(n >= 0 && d > 0 && m > 0 && ld >= d && ldy >= m, "pmb_project: bad sizes");
  if (n == 0) return PMB_OK;
  PMB_REQUIRE(X && a && nanfill && W && Y, "pmb_project: null pointer");
  cudaStream_t st = as_stream(stream);
  if (!out_f64 && m <= 16 && d % 4 == 0 && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 &&
      (size_t)d * 16 * 4 <= 160 * 1024) {
    float* Yf = static_cast<float*>(Y);
    if (d <= 128) return launch_project_warp<1>(X, n, d, ld, a, nanfill, W, m, Yf, ldy, st);
    if (d <= 256) return launch_project_warp<2>(X, n, d, ld, a, nanfill, W, m, Yf, ldy, st);
    if (m <= 4) return launch_project_fast<4>(X, n, d, ld, a, nanfill, W, m, Yf, ldy, st);
    if (m <= 8) return launch_project_fast<8>(X, n, d, ld, a, nanfill, W, m, Yf, ldy, st);
    if (m <= 12) return launch_project_fast<12>(X, n, d, ld, a, nanfill, W, m, Yf, ldy, st);
    return launch_project_fast<16>(X, n, d, ld, a, nanfill, W, m, Yf, ldy, st);
  }
  for (int c_off = 0; c_off < m; c_off += 16) {
    const int rem = m - c_off;
    int rc;
    if (rem <= 2) rc = launch_project<2>(X, n, d, ld, a, nanfill, W, m, c_off, Y, ldy, out_f64, st);
    else if (rem <= 4) rc = launch_project<4>(X, n, d, ld, a, nanfill, W, m, c_off, Y, ldy, out_f64, st);
    else if (rem <= 8) rc = launch_project<8>(X, n, d, ld, a, nanfill, W, m, c_off, Y, ldy, out_f64, st);
    else if (rem <= 12) rc = launch_project<12>(X, n, d, ld, a, nanfill, W, m, c_off, Y, ldy, out_f64, st);
    else rc = launch_project<16>(X, n, d, ld, a, nanfill, W, m, c_off, Y, ldy, out_f64, st);
    if (rc != PMB_OK) return rc;
  }
  return PMB_OK;
}
