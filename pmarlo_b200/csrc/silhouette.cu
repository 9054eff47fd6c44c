// silhouette.cu -- per-sample silhouette coefficients for the automatic choice of the number of microstates
// (`_auto_select_n_states`, src/pmarlo/markov_state_model/clustering.py:155-250, which calls
// sklearn.metrics.silhouette_score on the (sampled) frames for k = 4 .. 20).
//
//   a_i = mean distance from frame i to the other members of its cluster          (0 for a singleton)
//   b_i = min over the other clusters of the mean distance to their members
//   s_i = (b_i - a_i) / max(a_i, b_i)      (0 for singletons, 0 when max(a, b) = 0 like sklearn's nan_to_num)
//
// O(n^2 D) Euclidean distances in fp64: one thread per frame i, the frames j staged through shared memory in
// tiles, per-thread distance sums per cluster in shared memory (K <= 64 clusters).  Each tile of Y is read
// from L2 once per CTA; the n x n distance matrix sklearn materialises in chunks never exists.
#include "common.cuh"

namespace pmb {

constexpr int kSilThreads = 128;
constexpr int kSilTile = 128;
constexpr int kSilMaxK = 64;

__global__ void __launch_bounds__(kSilThreads) silhouette_kernel(const double* __restrict__ Y, int64_t n, int D,
                                                                 const int32_t* __restrict__ labels, int K,
                                                                 const int64_t* __restrict__ sizes,
                                                                 double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_y = reinterpret_cast<double*>(smem_raw);                  // kSilTile x D
  int* s_lab = reinterpret_cast<int*>(s_y + (size_t)kSilTile * D);    // kSilTile
  double* s_acc = reinterpret_cast<double*>(s_lab + kSilTile);        // kSilThreads x K, thread-major with stride K + 1... (see below)
  const int tid = threadIdx.x;
  const int64_t i = (int64_t)blockIdx.x * kSilThreads + tid;
  const bool valid = i < n;
  double* acc = s_acc + (size_t)tid * (K | 1);                        // odd stride: no systematic bank conflicts
  for (int k = 0; k < K; ++k) acc[k] = 0.0;
  const double* yi = Y + (valid ? i : 0) * D;
  for (int64_t j0 = 0; j0 < n; j0 += kSilTile) {
    __syncthreads();
    const int nj = (int)((n - j0) < kSilTile ? (n - j0) : kSilTile);
    for (int e = tid; e < nj * D; e += kSilThreads) s_y[e] = Y[j0 * D + e];
    for (int e = tid; e < nj; e += kSilThreads) s_lab[e] = labels[j0 + e];
    __syncthreads();
    if (valid) {
      for (int j = 0; j < nj; ++j) {
        double d2 = 0.0;
        for (int q = 0; q < D; ++q) {
          const double t = yi[q] - s_y[j * D + q];
          d2 = fma(t, t, d2);
        }
        acc[s_lab[j]] += sqrt(d2);
      }
    }
  }
  if (!valid) return;
  const int li = labels[i];
  const int64_t ni = sizes[li];
  double s = 0.0;
  if (ni > 1) {
    const double a = acc[li] / (double)(ni - 1);
    double b = 1.7976931348623157e308;
    for (int k = 0; k < K; ++k)
      if (k != li && sizes[k] > 0) b = fmin(b, acc[k] / (double)sizes[k]);
    const double m = fmax(a, b);
    s = (m > 0.0 && b < 1.7976931348623157e308) ? (b - a) / m : 0.0;
  }
  out[i] = s;
}

}  // namespace pmb

extern "C" int pmb_silhouette_samples(const double* Y, int64_t n, int D, const int32_t* labels, int K,
                                      const int64_t* sizes, double* out, pmb_stream_t stream) {
  using namespace pmb;
  PMB_REQUIRE(n > 0 && D > 0 && K > 0 && K <= kSilMaxK, "pmb_silhouette_samples: bad sizes (K <= 64)");
  PMB_REQUIRE(Y && labels && sizes && out, "pmb_silhouette_samples: null pointer");
  const size_t smem = (size_t)kSilTile * D * sizeof(double) + kSilTile * sizeof(int) + (size_t)kSilThreads * (K | 1) * sizeof(double);
  PMB_REQUIRE(smem <= 200 * 1024, "pmb_silhouette_samples: D = %d does not fit shared memory", D);
  PMB_CUDA(cudaFuncSetAttribute(silhouette_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)((n + kSilThreads - 1) / kSilThreads);
  silhouette_kernel<<<grid, kSilThreads, smem, as_stream(stream)>>>(Y, n, D, labels, K, sizes, out);
  PMB_LAUNCH_CHECK();
  return PMB_OK;
}
