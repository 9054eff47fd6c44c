// hist2d.cu -- weighted 2-D histogram of two collective variables: the O(N) part of the free-energy surface
// (`generate_2d_fes`, src/pmarlo/markov_state_model/free_energy.py:417-865, which calls np.histogram2d).
//
// numpy semantics: bins [e_k, e_{k+1}) except the last, which is closed on the right; samples outside the range
// (or NaN) are dropped; bin index floor((x - lo) / (hi - lo) * nb) corrected against the edges lo + k * width so
// that a sample exactly on an edge lands where np.histogram2d puts it.
// HBM-bound: 16 (+ 8) bytes read per frame; per-CTA shared-memory histogram (bx * by <= 16384 cells of fp64),
// merged into the global histogram with one atomic per non-empty cell.
#include "common.cuh"

namespace pmb {

constexpr int kHistThreads = 256;

__device__ __forceinline__ int hist_bin(double x, double lo, double hi, int nb) {
  if (!(x >= lo && x <= hi)) return -1;         // also drops NaN
  if (x == hi) return nb - 1;
  const double w = (hi - lo) / (double)nb;
  int k = (int)floor((x - lo) / (hi - lo) * (double)nb);
  if (k >= nb) k = nb - 1;
  if (k < 0) k = 0;
  // edges as numpy builds them: linspace(lo, hi, nb + 1)[k] = lo + k * w (to rounding); one correction step
  if (x < lo + (double)k * w && k > 0) --k;
  else if (k + 1 < nb && x >= lo + (double)(k + 1) * w) ++k;
  return k;
}

__global__ void __launch_bounds__(kHistThreads) hist2d_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                              const double* __restrict__ w, int64_t n, double xlo, double xhi,
                                                              int bx, double ylo, double yhi, int by, double* __restrict__ H) {
  extern __shared__ double s_h[];
  const int cells = bx * by;
  for (int c = threadIdx.x; c < cells; c += kHistThreads) s_h[c] = 0.0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * kHistThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kHistThreads) {
    const int kx = hist_bin(x[i], xlo, xhi, bx), ky = hist_bin(y[i], ylo, yhi, by);
    if (kx >= 0 && ky >= 0) atomicAdd(&s_h[kx * by + ky], w != nullptr ? w[i] : 1.0);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cells; c += kHistThreads)
    if (s_h[c] != 0.0) atomicAdd(H + c, s_h[c]);
}

}  // namespace pmb

extern "C" int pmb_hist2d(const double* x, const double* y, const double* w, int64_t n, double xlo, double xhi, int bx,
                          double ylo, double yhi, int by, double* H, pmb_stream_t stream) {
  using namespace pmb;
  PMB_REQUIRE(n >= 0 && bx > 0 && by > 0 && (int64_t)bx * by <= 16384, "pmb_hist2d: bad sizes (bx * by <= 16384)");
  PMB_REQUIRE(xhi > xlo && yhi > ylo, "pmb_hist2d: empty range");
  PMB_REQUIRE(x && y && H, "pmb_hist2d: null pointer");
  if (n == 0) return PMB_OK;
  const size_t smem = (size_t)bx * by * sizeof(double);
  PMB_CUDA(cudaFuncSetAttribute(hist2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t grid = (n + kHistThreads - 1) / kHistThreads;
  if (grid > 2 * kNumSMs) grid = 2 * kNumSMs;
  hist2d_kernel<<<(int)grid, kHistThreads, smem, as_stream(stream)>>>(x, y, w, n, xlo, xhi, bx, ylo, yhi, by, H);
  PMB_LAUNCH_CHECK();
  return PMB_OK;
}
