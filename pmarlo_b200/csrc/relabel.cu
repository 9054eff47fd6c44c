// relabel.cu -- state relabelling with frame removal, the O(N) step of the Chapman-Kolmogorov test.
//
// The reference rebuilds every discrete trajectory as
//   [state_map[s] for s in traj if s in state_map]
// (ck_runner.py:150-153, :235-238, _ck.py:86-90): frames whose state was dropped are REMOVED, the
// survivors keep their order, and lagged pairs are then counted on the shortened trajectories.  On the
// device this is an order-preserving stream compaction of the label shard plus new shard offsets:
//   1. kept frames per 4096-frame tile,
//   2. exclusive scan of the tile totals (one CTA),
//   3. scatter: ballot ranks inside a warp, warp totals through shared memory, tile base from (2),
//   4. new offset of every shard boundary = kept frames before it.
// HBM-bound: 8 B read + <= 4 B written per frame; the lookup table (<= a few thousand int32) stays in L1.
// A table without -1 entries relabels in place of compaction (macrostate lumping, ck_runner.py:204).
#include "common.cuh"

namespace pmb {

constexpr int kRlThreads = 256;
constexpr int kRlItems = 16;
constexpr int kRlTile = kRlThreads * kRlItems;

__device__ __forceinline__ int rl_map(int s, const int32_t* __restrict__ lut, int n_lut) {
  return (s >= 0 && s < n_lut) ? __ldg(lut + s) : -1;
}

__global__ void __launch_bounds__(kRlThreads) relabel_count_kernel(
    const int32_t* __restrict__ labels, int64_t n, const int32_t* __restrict__ lut, int n_lut,
    int64_t n_tiles, long long* __restrict__ tile_count) {
  __shared__ int s_w[kRlThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t base = tile * kRlTile;
    int c = 0;
#pragma unroll
    for (int j = 0; j < kRlItems; ++j) {
      const int64_t g = base + (int64_t)j * kRlThreads + tid;
      if (g < n) c += rl_map(labels[g], lut, n_lut) >= 0;
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) s_w[warp] = c;
    __syncthreads();
    if (tid == 0) {
      int t = 0;
#pragma unroll
      for (int w = 0; w < kRlThreads / 32; ++w) t += s_w[w];
      tile_count[tile] = t;
    }
    __syncthreads();
  }
}

// exclusive scan in place; prefix[n_tiles] = total
__global__ void __launch_bounds__(1024) relabel_scan_kernel(long long* __restrict__ prefix, int64_t n_tiles) {
  __shared__ long long s_part[1024];
  const int tid = threadIdx.x;
  const int64_t per = (n_tiles + 1023) / 1024;
  const int64_t lo = (int64_t)tid * per, hi = (lo + per < n_tiles) ? lo + per : n_tiles;
  long long sum = 0;
  for (int64_t i = lo; i < hi; ++i) sum += prefix[i];
  s_part[tid] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {   // Hillis-Steele inclusive scan of the per-thread sums
    const long long v = (tid >= o) ? s_part[tid - o] : 0;
    __syncthreads();
    s_part[tid] += v;
    __syncthreads();
  }
  long long run = s_part[tid] - sum;
  for (int64_t i = lo; i < hi; ++i) {
    const long long c = prefix[i];
    prefix[i] = run;
    run += c;
  }
  if (tid == 1023) prefix[n_tiles] = s_part[1023];
}

__global__ void __launch_bounds__(kRlThreads) relabel_scatter_kernel(
    const int32_t* __restrict__ labels, int64_t n, const int32_t* __restrict__ lut, int n_lut,
    int64_t n_tiles, const long long* __restrict__ prefix, int32_t* __restrict__ out) {
  __shared__ int s_w[2][kRlThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t base = tile * kRlTile;
    long long run = prefix[tile];
#pragma unroll 4
    for (int j = 0; j < kRlItems; ++j) {
      const int64_t g = base + (int64_t)j * kRlThreads + tid;
      const int m = (g < n) ? rl_map(labels[g], lut, n_lut) : -1;
      const unsigned keep = __ballot_sync(0xffffffffu, m >= 0);
      if (lane == 0) s_w[j & 1][warp] = __popc(keep);
      __syncthreads();
      int before = 0, total = 0;
#pragma unroll
      for (int w = 0; w < kRlThreads / 32; ++w) {
        const int t = s_w[j & 1][w];
        before += (w < warp) ? t : 0;
        total += t;
      }
      if (m >= 0) out[run + before + __popc(keep & ((1u << lane) - 1u))] = m;
      run += total;
    }
    __syncthreads();   // both halves of s_w are free again before the next tile
  }
}

// one warp per shard boundary: new_off[s] = kept frames in [0, off[s])
__global__ void __launch_bounds__(kRlThreads) relabel_offsets_kernel(
    const int32_t* __restrict__ labels, int64_t n, const int32_t* __restrict__ lut, int n_lut,
    const int64_t* __restrict__ off, int n_seg, int64_t n_tiles, const long long* __restrict__ prefix,
    int64_t* __restrict__ new_off) {
  const int lane = threadIdx.x & 31;
  const int s = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  if (s > n_seg) return;
  int64_t g = off[s];
  g = g < 0 ? 0 : (g > n ? n : g);
  const int64_t tile = g / kRlTile;
  int c = 0;
  for (int64_t i = tile * kRlTile + lane; i < g; i += 32) c += rl_map(labels[i], lut, n_lut) >= 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) new_off[s] = (int64_t)prefix[tile < n_tiles ? tile : n_tiles] + c;
}

}  // namespace pmb

extern "C" size_t pmb_relabel_compact_ws_bytes(int64_t n) {
  const int64_t n_tiles = (n + pmb::kRlTile - 1) / pmb::kRlTile;
  return (size_t)(n_tiles + 1) * sizeof(long long);
}

extern "C" int pmb_relabel_compact(const int32_t* labels, int64_t n, const int64_t* seg_offsets, int n_seg,
                                   const int32_t* lut, int n_lut, int32_t* labels_out,
                                   int64_t* new_seg_offsets, void* workspace, size_t workspace_bytes,
                                   pmb_stream_t stream) {
  using namespace pmb;
  PMB_REQUIRE(n >= 0 && n_seg >= 1 && n_lut >= 1, "pmb_relabel_compact: bad sizes");
  PMB_REQUIRE(seg_offsets && lut && new_seg_offsets && workspace, "pmb_relabel_compact: null pointer");
  PMB_REQUIRE(n == 0 || (labels && labels_out), "pmb_relabel_compact: null pointer");
  if (workspace_bytes < pmb_relabel_compact_ws_bytes(n)) return PMB_EWORKSPACE;
  const int64_t n_tiles = (n + kRlTile - 1) / kRlTile;
  long long* prefix = static_cast<long long*>(workspace);
  cudaStream_t st = as_stream(stream);
  if (n_tiles > 0) {
    const int grid = (int)(n_tiles < 8 * (int64_t)kNumSMs ? n_tiles : 8 * (int64_t)kNumSMs);
    relabel_count_kernel<<<grid, kRlThreads, 0, st>>>(labels, n, lut, n_lut, n_tiles, prefix);
    relabel_scan_kernel<<<1, 1024, 0, st>>>(prefix, n_tiles);
    relabel_scatter_kernel<<<grid, kRlThreads, 0, st>>>(labels, n, lut, n_lut, n_tiles, prefix, labels_out);
  } else {
    cudaMemsetAsync(prefix, 0, sizeof(long long), st);
  }
  const int64_t warps = (int64_t)n_seg + 1;
  relabel_offsets_kernel<<<(unsigned)((warps * 32 + kRlThreads - 1) / kRlThreads), kRlThreads, 0, st>>>(
      labels, n, lut, n_lut, seg_offsets, n_seg, n_tiles, prefix, new_seg_offsets);
  PMB_LAUNCH_CHECK();
  return PMB_OK;
}
