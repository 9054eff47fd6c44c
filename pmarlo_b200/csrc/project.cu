// project.cu -- K5: TICA projection  y = (x_imputed - a) W  (skinny GEMM, HBM-bound).
//
// Reads 4*d bytes and writes 4*m (or 8*m) bytes per frame.  A CTA owns 256
// consecutive frames; the feature axis is streamed through shared memory in
// chunks of 32 columns (every warp load is one 128-byte line), stored with a
// padded stride so that the thread-per-frame reads are bank-conflict free.
// Arithmetic is fp64 (W, a in shared memory, broadcast reads) so that the
// projection itself adds no error beyond the fp32 input.
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
