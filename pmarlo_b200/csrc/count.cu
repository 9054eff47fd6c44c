// count.cu -- K7: lagged transition histogram over discrete trajectories.
//
// 4 bytes/frame algorithmic traffic (the lag-shifted re-read hits L1/L2), one
// atomic per pair.  Bit-exact integer result.
//  * K <= kCountSmemK: CTA-private int32 histogram in shared memory, flushed
//    (non-zero bins only) to the global int64 matrix at the end;
//  * larger K: global 64-bit atomics (the K*K matrix stays L2-resident up to
//    K ~ 3900) with warp aggregation -- consecutive frames of a metastable
//    trajectory mostly hit the same (i,j) bin, so __match_any_sync collapses
//    them to one atomic per distinct bin per warp.
// Segments ("shards"): a pair never crosses a trajectory boundary; the segment
// of a frame is found by a warp-uniform binary search in seg_offsets.
#include "common.cuh"

namespace pmb {

constexpr int kCountThreads = 256;
constexpr int kCountSmemK = 104;  // 104*104*4 B = 43 KB

__device__ __forceinline__ int seg_find(const int64_t* __restrict__ off, int n_seg, int64_t g) {
  if (g < off[0] || g >= off[n_seg]) return -1;
  int lo = 0, hi = n_seg;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= g) lo = mid; else hi = mid;
  }
  return lo;
}

// returns the bin (i*K+j) of the pair starting at frame g, or -1
__device__ __forceinline__ int64_t pair_bin(const int32_t* __restrict__ labels, int64_t n,
                                            const int64_t* __restrict__ off, int n_seg, int K,
                                            int lag, int step, int64_t g, int& seg_cache) {
  if (g >= n) return -1;
  int s = seg_cache;
  if (s < 0 || g < off[s] || g >= off[s + 1]) s = seg_find(off, n_seg, g);
  seg_cache = s;
  if (s < 0) return -1;
  const int64_t t = g - off[s];
  if (g + lag >= off[s + 1]) return -1;
  if (step > 1 && (t % step) != 0) return -1;
  const int a = labels[g], b = labels[g + lag];
  if (a < 0 || b < 0 || a >= K || b >= K) return -1;
  return (int64_t)a * K + b;
}

template <bool WEIGHTED>
__global__ void __launch_bounds__(kCountThreads) count_global_kernel(
    const int32_t* __restrict__ labels, const double* __restrict__ weights, int64_t n,
    const int64_t* __restrict__ off, int n_seg, int K, int lag, int step, void* __restrict__ Cout) {
  int seg_cache = -1;
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n_round = ((n + 31) / 32) * 32;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_round; g += stride) {
    const int64_t bin = pair_bin(labels, n, off, n_seg, K, lag, step, g, seg_cache);
    if constexpr (WEIGHTED) {
      if (bin >= 0) atomicAdd(static_cast<double*>(Cout) + bin, weights[g]);
    } else {
      // all 32 lanes of the warp reach this point together (n_round is a multiple of 32)
      const unsigned peers = __match_any_sync(0xffffffffu, bin);
      if (bin >= 0 && lane == (__ffs(peers) - 1))
        atomicAdd(static_cast<unsigned long long*>(Cout) + bin, (unsigned long long)__popc(peers));
    }
  }
}

__global__ void __launch_bounds__(kCountThreads) count_smem_kernel(
    const int32_t* __restrict__ labels, int64_t n, const int64_t* __restrict__ off, int n_seg, int K,
    int lag, int step, unsigned long long* __restrict__ C) {
  extern __shared__ unsigned int hist[];
  const int KK = K * K;
  for (int i = threadIdx.x; i < KK; i += blockDim.x) hist[i] = 0u;
  __syncthreads();
  int seg_cache = -1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += stride) {
    const int64_t bin = pair_bin(labels, n, off, n_seg, K, lag, step, g, seg_cache);
    if (bin >= 0) atomicAdd(&hist[(int)bin], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < KK; i += blockDim.x) {
    const unsigned int v = hist[i];
    if (v) atomicAdd(C + i, (unsigned long long)v);
  }
}

// one warp per state: row + column totals, fp64 copy of the row
__global__ void counts_active_kernel(const long long* __restrict__ C, int K, double eps,
                                     double* __restrict__ Cf, uint8_t* __restrict__ active) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= K) return;
  long long tot = 0;
  for (int j = lane; j < K; j += 32) {
    const long long r = C[(size_t)i * K + j];
    tot += r + C[(size_t)j * K + i];
    Cf[(size_t)i * K + j] = (double)r;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
  if (lane == 0) active[i] = ((double)tot > eps) ? 1 : 0;
}

}  // namespace pmb

extern "C" int pmb_counts_active(const int64_t* C, int K, double eps, double* Cf, uint8_t* active,
                                 pmb_stream_t stream) {
  using namespace pmb;
  PMB_REQUIRE(K > 0, "pmb_counts_active: bad K");
  PMB_REQUIRE(C && Cf && active, "pmb_counts_active: null pointer");
  counts_active_kernel<<<(K * 32 + 255) / 256, 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const long long*>(C), K, eps, Cf, active);
  PMB_LAUNCH_CHECK();
  return PMB_OK;
}

extern "C" int pmb_count_lagged(const int32_t* labels, int64_t n, const int64_t* seg_offsets,
                                int n_seg, int K, int lag, int step, int64_t* C,
                                pmb_stream_t stream) {
  using namespace pmb;
  PMB_REQUIRE(n >= 0 && n_seg >= 1 && K > 0 && lag >= 1 && step >= 1, "pmb_count_lagged: bad sizes");
  PMB_REQUIRE((int64_t)K * K < (int64_t)1 << 31, "pmb_count_lagged: K too large");
  if (n == 0) return PMB_OK;
  PMB_REQUIRE(labels && seg_offsets && C, "pmb_count_lagged: null pointer");
  const int64_t blocks_needed = (n + kCountThreads - 1) / kCountThreads;
  // shared-memory privatisation only pays when every CTA sees many more frames than bins
  const int smem_grid = 2 * kNumSMs;
  const bool use_smem = K <= kCountSmemK && n >= (int64_t)smem_grid * K * K * 2;
  if (use_smem) {
    const size_t smem = (size_t)K * K * sizeof(unsigned int);
    count_smem_kernel<<<smem_grid, kCountThreads, smem, as_stream(stream)>>>(
        labels, n, seg_offsets, n_seg, K, lag, step, reinterpret_cast<unsigned long long*>(C));
  } else {
    int grid = (int)(blocks_needed < 8 * kNumSMs ? blocks_needed : 8 * kNumSMs);
    count_global_kernel<false><<<grid, kCountThreads, 0, as_stream(stream)>>>(
        labels, nullptr, n, seg_offsets, n_seg, K, lag, step, C);
  }
  PMB_LAUNCH_CHECK();
  return PMB_OK;
}

extern "C" int pmb_count_lagged_weighted(const int32_t* labels, const double* weights, int64_t n,
                                         const int64_t* seg_offsets, int n_seg, int K, int lag,
                                         int step, double* C, pmb_stream_t stream) {
  using namespace pmb;
  PMB_REQUIRE(n >= 0 && n_seg >= 1 && K > 0 && lag >= 1 && step >= 1, "pmb_count_lagged_weighted: bad sizes");
  if (n == 0) return PMB_OK;
  PMB_REQUIRE(labels && weights && seg_offsets && C, "pmb_count_lagged_weighted: null pointer");
  const int64_t blocks_needed = (n + kCountThreads - 1) / kCountThreads;
  int grid = (int)(blocks_needed < 8 * kNumSMs ? blocks_needed : 8 * kNumSMs);
  count_global_kernel<true><<<grid, kCountThreads, 0, as_stream(stream)>>>(
      labels, weights, n, seg_offsets, n_seg, K, lag, step, C);
  PMB_LAUNCH_CHECK();
  return PMB_OK;
}
