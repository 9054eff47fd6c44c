// mle.cu -- K8: reversible maximum-likelihood transition matrix (fixed point on
// the row-sum vector, matrix-free).
//
//   S = C + C^T, c_i = sum_j C_ij, x^0 = rowsum(S)/sum(S)
//   q_i = c_i / x_i
//   u_i = sum_j S_ij / (q_i + q_j);  x' = u / sum(u)
//   err = max_i |x_i - x'_i| / (0.5 (x_i + x'_i));  stop err <= maxerr or maxiter
//   T_ij = (S_ij / (q_i + q_j)) / rs_i,  pi = rs / sum(rs)   with rs from the final x
//
// Only the K-vector x feeds back, so one iteration is ONE streaming pass over S
// (8 K^2 bytes, fp64).  Two kernels:
//  * one CTA per problem (batched; K <= kMleCtaMaxK): x, q, u live in shared
//    memory, S is re-read from L2, barriers are __syncthreads -- this is the
//    ITS sweep path (one CTA per lag time);
//  * one cooperative grid per problem (large K): rows are spread over all SMs
//    (a CTA's rows of S stay L1-resident across iterations), u goes through a
//    global double buffer and a flag-based all-gather instead of a grid barrier:
//    every CTA publishes "generation g written" in its own flag word
//    (st.release), the consumers poll the 148 flags in parallel (ld.acquire) and
//    then read u from L2 -- one L2 round trip after the slowest producer.  Every
//    CTA recomputes the normalisation and the error in the same order, so all
//    CTAs take the same branch.  The launch is cooperative only to guarantee
//    co-residency of the CTAs that poll one another.
// The K^2 divisions per iteration use an fp32-seeded reciprocal with two fp64
// Newton steps (<= 1.5 ulp from the IEEE quotient; the fixed point is a
// contraction, so iterates track the reference's to ~1e-13).
// All reductions have a fixed order: results are bit-reproducible run to run.
// A state with active[i] == 0 is excluded (T_ii = 1, pi_i = 0), which is how the
// caller restricts the estimate to the largest connected set without
// re-packing the count matrix.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace pmb {

constexpr int kMleThreads = 1024;
constexpr int kMleCtaMaxK = 2048;
constexpr int kMleGridThreads = 256;

struct MleParams {
  const double* C;        // batch x K x K
  const uint8_t* active;  // batch x K or nullptr
  int K;
  double alpha;           // pseudocount added to every cell of the active block
  double maxerr;
  long long maxiter;
  double* T;              // batch x K x K
  double* pi;             // batch x K
  long long* info;        // batch x 2
  double* S;              // batch x K x K workspace
  double* cvec;           // batch x K workspace
  double* ubuf;           // 2 x K (grid kernel only)
  unsigned long long* flags;  // gridDim.x generation counters, 16 words apart (grid kernel only)
};

constexpr int kFlagStride = 16;  // 128 B between the flag words of different CTAs

// 1/d for d > 0: fp32 seed + two Newton steps in fp64 (falls back to the IEEE
// quotient outside the fp32 exponent range).
__device__ __forceinline__ double fast_rcp(double d) {
  if (!(d > 1e-30 && d < 1e30)) return 1.0 / d;
  double r = (double)__frcp_rn(__double2float_rn(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// deterministic CTA-wide sum / max; every thread gets the result
__device__ __forceinline__ double cta_sum(double v, double* s_red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  double r = 0.0;
  for (int w = 0; w < nw; ++w) r += s_red[w];
  return r;
}
__device__ __forceinline__ double cta_max(double v, double* s_red) {
  v = warp_max(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  double r = 0.0;
  for (int w = 0; w < nw; ++w) r = fmax(r, s_red[w]);
  return r;
}

// S = mask(C + C^T), c = masked row sums of C.  grid: (ceil(K/32), ceil(K/32), batch)
__global__ void mle_prepare_kernel(MleParams p) {
  __shared__ double tile[32][33];
  const int K = p.K;
  const size_t base = (size_t)blockIdx.z * K * K;
  const uint8_t* act = p.active ? p.active + (size_t)blockIdx.z * K : nullptr;
  const int bi = blockIdx.y * 32, bj = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int i = bj + r, j = bi + tx;  // transposed block
    tile[r][tx] = (i < K && j < K) ? p.C[base + (size_t)i * K + j] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = bi + r, j = bj + tx;
    if (i < K && j < K) {
      const bool on = !act || (act[i] && act[j]);
      p.S[base + (size_t)i * K + j] =
          on ? (p.C[base + (size_t)i * K + j] + p.alpha) + (tile[tx][r] + p.alpha) : 0.0;
    }
  }
}
__global__ void mle_rowsum_kernel(MleParams p) {
  const int K = p.K;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= K) return;
  const size_t base = (size_t)blockIdx.y * K * K;
  const uint8_t* act = p.active ? p.active + (size_t)blockIdx.y * K : nullptr;
  double acc = 0.0;
  if (!act || act[warp])
    for (int j = lane; j < K; j += 32)
      if (!act || act[j]) acc += p.C[base + (size_t)warp * K + j] + p.alpha;
  acc = warp_sum(acc);
  if (lane == 0) p.cvec[(size_t)blockIdx.y * K + warp] = acc;
}

__device__ __forceinline__ double row_apply(const double* __restrict__ Srow, const double* q, double qi,
                                            int K, int lane) {
  double acc0 = 0.0, acc1 = 0.0;
  int j = lane;
  for (; j + 32 < K; j += 64) {
    const double s0 = Srow[j], s1 = Srow[j + 32];
    const double r0 = fast_rcp(qi + q[j]), r1 = fast_rcp(qi + q[j + 32]);
    acc0 = fma(s0, r0, acc0);
    acc1 = fma(s1, r1, acc1);
  }
  if (j < K) acc0 = fma(Srow[j], fast_rcp(qi + q[j]), acc0);
  return warp_sum(acc0 + acc1);
}

// ---------------------------------------------------------------- one CTA per problem
__global__ void __launch_bounds__(kMleThreads) mle_cta_kernel(MleParams p) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double s_red[32];
  const int K = p.K;
  double* x = sm;
  double* q = sm + K;
  double* u = sm + 2 * K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const size_t b = blockIdx.x;
  const double* S = p.S + b * K * K;
  const double* c = p.cvec + b * K;
  const uint8_t* act = p.active ? p.active + b * K : nullptr;

  // x0 = rowsum(S) / sum(S); validity
  int bad = 0;
  for (int i = warp; i < K; i += nw) {
    double acc = 0.0;
    for (int j = lane; j < K; j += 32) acc += S[(size_t)i * K + j];
    acc = warp_sum(acc);
    if (lane == 0) u[i] = acc;
  }
  __syncthreads();
  double part = 0.0;
  for (int i = tid; i < K; i += blockDim.x) {
    part += u[i];
    const bool on = !act || act[i];
    if (on && !(c[i] > 0.0)) bad = 1;
  }
  const double tot = cta_sum(part, s_red);
  bad = __syncthreads_or(bad);
  if (bad || !(tot > 0.0)) {
    if (tid == 0) { p.info[2 * b] = 0; p.info[2 * b + 1] = -1; }
    return;
  }
  for (int i = tid; i < K; i += blockDim.x) x[i] = u[i] / tot;
  __syncthreads();

  long long it = 0;
  double err = 1.7976931348623157e308;
  while (it < p.maxiter && err > p.maxerr) {
    for (int i = tid; i < K; i += blockDim.x) q[i] = (x[i] > 0.0) ? c[i] / x[i] : 0.0;
    __syncthreads();
    for (int i = warp; i < K; i += nw) {
      double r = 0.0;
      if (x[i] > 0.0) r = row_apply(S + (size_t)i * K, q, q[i], K, lane);
      if (lane == 0) u[i] = r;
    }
    __syncthreads();
    part = 0.0;
    for (int i = tid; i < K; i += blockDim.x) part += u[i];
    const double norm = cta_sum(part, s_red);
    double e = 0.0;
    for (int i = tid; i < K; i += blockDim.x) {
      const double xo = x[i], xn = u[i] / norm;
      if (xo > 0.0 || xn > 0.0) e = fmax(e, fabs(xo - xn) / (0.5 * (xo + xn)));
      x[i] = xn;
    }
    err = cta_max(e, s_red);
    ++it;
  }
  // final application with the converged x
  for (int i = tid; i < K; i += blockDim.x) q[i] = (x[i] > 0.0) ? c[i] / x[i] : 0.0;
  __syncthreads();
  double* T = p.T + b * K * K;
  for (int i = warp; i < K; i += nw) {
    const bool on = x[i] > 0.0;
    double rs = 0.0;
    if (on) rs = row_apply(S + (size_t)i * K, q, q[i], K, lane);
    if (lane == 0) u[i] = rs;
    const double qi = q[i];
    for (int j = lane; j < K; j += 32) {
      double t;
      if (on && rs > 0.0) {
        const double s = S[(size_t)i * K + j];
        t = (s != 0.0) ? (s / (qi + q[j])) / rs : 0.0;
      } else {
        t = (i == j) ? 1.0 : 0.0;
      }
      T[(size_t)i * K + j] = t;
    }
  }
  __syncthreads();
  part = 0.0;
  for (int i = tid; i < K; i += blockDim.x) part += u[i];
  const double rtot = cta_sum(part, s_red);
  for (int i = tid; i < K; i += blockDim.x) p.pi[b * K + i] = u[i] / rtot;
  if (tid == 0) { p.info[2 * b] = it; p.info[2 * b + 1] = (err <= p.maxerr) ? 1 : 0; }
}

// ---------------------------------------------------------------- cooperative grid, one problem
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// publish generation g of this CTA's rows, then wait until every CTA has published it.
// Only warp 0 polls (relaxed loads; one fence after the loop): polling with acquire loads from
// every thread costs a memory barrier per poll and floods L2.
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void gather_sync(unsigned long long* flags, unsigned long long gen) {
  __syncthreads();  // all row results of this CTA are written
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) {
      __threadfence();
      st_release_u64(flags + (size_t)blockIdx.x * kFlagStride, gen);
    }
    for (int c = threadIdx.x; c < (int)gridDim.x; c += 32) {
      const unsigned long long* f = flags + (size_t)c * kFlagStride;
      unsigned int spins = 0;
      while (ld_relaxed_u64(f) < gen) {
        if (++spins > (1u << 24)) __trap();  // a lost CTA must not hang the GPU
      }
    }
    __threadfence();  // acquire side: order the u reads below after the flag observations
  }
  __syncthreads();
}

__device__ long long g_dbg[8];  // phase cycle counters of the last grid-kernel run (CTA 0, thread 0)

// REG: one row per warp with K <= 1024: the warp keeps its row of S in registers (32 doubles per
// lane) for the whole iteration, so an iteration touches no memory but the K-vector exchange.
template <bool REG>
__global__ void __launch_bounds__(kMleGridThreads) mle_grid_kernel(MleParams p) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double s_red[32];
  const int K = p.K;
  double* x = sm;
  double* q = sm + K;
  double* cs = sm + 2 * K;   // row sums of the (regularised) counts, staged once
  const int tid = threadIdx.x, lane = tid & 31;
  const int gwarp = (blockIdx.x * blockDim.x + tid) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int i = tid; i < K; i += blockDim.x) cs[i] = p.cvec[i];
  double sreg[32];
  if constexpr (REG) {
#pragma unroll
    for (int t = 0; t < 32; ++t) {
      const int j = lane + 32 * t;
      sreg[t] = (gwarp < K && j < K) ? p.S[(size_t)gwarp * K + j] : 0.0;
    }
  }
  auto apply_row = [&](int i, const double* qv, double qi) -> double {
    if constexpr (REG) {
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int t = 0; t < 32; t += 2) {
        const int j0 = lane + 32 * t, j1 = j0 + 32;
        if (j0 < K) a0 = fma(sreg[t], fast_rcp(qi + qv[j0]), a0);
        if (j1 < K) a1 = fma(sreg[t + 1], fast_rcp(qi + qv[j1]), a1);
      }
      return warp_sum(a0 + a1);
    } else {
      return row_apply(p.S + (size_t)i * K, qv, qi, K, lane);
    }
  };
  const double* __restrict__ S = p.S;
  const double* __restrict__ c = p.cvec;
  const uint8_t* act = p.active;
  unsigned long long gen = 0;

  // generation 1: row sums of S
  {
    double* ub = p.ubuf + (size_t)((gen + 1) & 1) * K;
    for (int i = gwarp; i < K; i += nwarps) {
      double acc = 0.0;
      for (int j = lane; j < K; j += 32) acc += S[(size_t)i * K + j];
      acc = warp_sum(acc);
      if (lane == 0) ub[i] = acc;
    }
    ++gen;
    gather_sync(p.flags, gen);
  }
  const double* ucur = p.ubuf + (size_t)(gen & 1) * K;
  double part = 0.0;
  int bad = 0;
  for (int i = tid; i < K; i += blockDim.x) {
    part += __ldcg(ucur + i);
    const bool on = !act || act[i];
    if (on && !(c[i] > 0.0)) bad = 1;
  }
  const double tot = cta_sum(part, s_red);
  bad = __syncthreads_or(bad);
  if (bad || !(tot > 0.0)) {  // uniform across the grid: every CTA sees the same data
    if (blockIdx.x == 0 && tid == 0) { p.info[0] = 0; p.info[1] = -1; }
    return;
  }
  for (int i = tid; i < K; i += blockDim.x) x[i] = __ldcg(ucur + i) / tot;
  __syncthreads();

  long long it = 0;
  double err = 1.7976931348623157e308;
  long long cyc[4] = {0, 0, 0, 0};
  while (it < p.maxiter && err > p.maxerr) {
    const long long t0 = clock64();
    for (int i = tid; i < K; i += blockDim.x) q[i] = (x[i] > 0.0) ? cs[i] * fast_rcp(x[i]) : 0.0;
    __syncthreads();
    const long long t1 = clock64();
    double* ub = p.ubuf + (size_t)((gen + 1) & 1) * K;
    for (int i = gwarp; i < K; i += nwarps) {
      double r = 0.0;
      if (x[i] > 0.0) r = apply_row(i, q, q[i]);
      if (lane == 0) ub[i] = r;
    }
    const long long t2 = clock64();
    ++gen;
    gather_sync(p.flags, gen);
    const long long t3 = clock64();
    ucur = ub;
    part = 0.0;
    for (int i = tid; i < K; i += blockDim.x) part += __ldcg(ucur + i);
    const double rnorm = 1.0 / cta_sum(part, s_red);
    double e = 0.0;
    for (int i = tid; i < K; i += blockDim.x) {
      const double xo = x[i], xn = __ldcg(ucur + i) * rnorm;
      if (xo > 0.0 || xn > 0.0) e = fmax(e, fabs(xo - xn) * fast_rcp(0.5 * (xo + xn)));
      x[i] = xn;
    }
    err = cta_max(e, s_red);
    ++it;
    const long long t4 = clock64();
    cyc[0] += t1 - t0; cyc[1] += t2 - t1; cyc[2] += t3 - t2; cyc[3] += t4 - t3;
  }
  if (blockIdx.x == 0 && tid == 0) {
    g_dbg[0] = it; g_dbg[1] = cyc[0]; g_dbg[2] = cyc[1]; g_dbg[3] = cyc[2]; g_dbg[4] = cyc[3];
    g_dbg[5] = gridDim.x; g_dbg[6] = REG ? 1 : 0;
  }
  for (int i = tid; i < K; i += blockDim.x) q[i] = (x[i] > 0.0) ? cs[i] / x[i] : 0.0;
  __syncthreads();
  double* ub = p.ubuf + (size_t)((gen + 1) & 1) * K;
  for (int i = gwarp; i < K; i += nwarps) {
    const bool on = x[i] > 0.0;
    double rs = 0.0;
    if (on) rs = apply_row(i, q, q[i]);
    if (lane == 0) ub[i] = rs;
    const double qi = q[i];
    for (int j = lane; j < K; j += 32) {
      double t;
      if (on && rs > 0.0) {
        const double s = S[(size_t)i * K + j];
        t = (s != 0.0) ? (s * fast_rcp(qi + q[j])) / rs : 0.0;
      } else {
        t = (i == j) ? 1.0 : 0.0;
      }
      p.T[(size_t)i * K + j] = t;
    }
  }
  ++gen;
  gather_sync(p.flags, gen);
  part = 0.0;
  for (int i = tid; i < K; i += blockDim.x) part += __ldcg(ub + i);
  const double rtot = cta_sum(part, s_red);
  for (int i = blockIdx.x * blockDim.x + tid; i < K; i += gridDim.x * blockDim.x)
    p.pi[i] = __ldcg(ub + i) / rtot;
  if (blockIdx.x == 0 && tid == 0) { p.info[0] = it; p.info[1] = (err <= p.maxerr) ? 1 : 0; }
}

}  // namespace pmb

extern "C" size_t pmb_mle_rev_ws_bytes(int K, int batch) {
  if (K <= 0 || batch <= 0) return 0;
  return ((size_t)batch * K * K + (size_t)batch * K + 2 * (size_t)K) * sizeof(double) + 64 +
         (size_t)256 * pmb::kFlagStride * sizeof(unsigned long long);
}

extern "C" int pmb_mle_rev(const double* C, const uint8_t* active, int K, int batch, double alpha, double maxerr,
                           int64_t maxiter, double* T, double* pi, int64_t* info, void* ws,
                           size_t ws_bytes, pmb_stream_t stream) {
  using namespace pmb;
  PMB_REQUIRE(K > 0 && batch > 0 && maxiter >= 0 && alpha >= 0.0, "pmb_mle_rev: bad sizes");
  PMB_REQUIRE(C && T && pi && info && ws, "pmb_mle_rev: null pointer");
  if (ws_bytes < pmb_mle_rev_ws_bytes(K, batch)) {
    set_error("pmb_mle_rev: workspace too small (%zu < %zu)", ws_bytes, pmb_mle_rev_ws_bytes(K, batch));
    return PMB_EWORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  MleParams p;
  p.C = C; p.active = active; p.K = K; p.alpha = alpha; p.maxerr = maxerr; p.maxiter = maxiter;
  p.T = T; p.pi = pi; p.info = reinterpret_cast<long long*>(info);
  p.S = static_cast<double*>(ws);
  p.cvec = p.S + (size_t)batch * K * K;
  p.ubuf = p.cvec + (size_t)batch * K;
  p.flags = reinterpret_cast<unsigned long long*>(p.ubuf + 2 * (size_t)K);
  {
    dim3 g((K + 31) / 32, (K + 31) / 32, batch), b(32, 8);
    mle_prepare_kernel<<<g, b, 0, st>>>(p);
    PMB_LAUNCH_CHECK();
    dim3 g2((K * 32 + 255) / 256, batch);
    mle_rowsum_kernel<<<g2, 256, 0, st>>>(p);
    PMB_LAUNCH_CHECK();
  }
  const bool use_cta = (K <= kMleCtaMaxK) && (batch > 1 || K <= 384);
  if (use_cta) {
    const size_t smem = (size_t)3 * K * sizeof(double);
    PMB_CUDA(cudaFuncSetAttribute(mle_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mle_cta_kernel<<<batch, kMleThreads, smem, st>>>(p);
    PMB_LAUNCH_CHECK();
    return PMB_OK;
  }
  const size_t smem = (size_t)3 * K * sizeof(double);
  PMB_REQUIRE(smem <= 200 * 1024, "pmb_mle_rev: K=%d too large", K);
  PMB_CUDA(cudaFuncSetAttribute(mle_grid_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PMB_CUDA(cudaFuncSetAttribute(mle_grid_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  PMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mle_grid_kernel<true>, kMleGridThreads, smem));
  PMB_REQUIRE(per_sm >= 1, "pmb_mle_rev: kernel does not fit on an SM");
  int dev = 0, sms = 0;
  PMB_CUDA(cudaGetDevice(&dev));
  PMB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int warps_per_cta = kMleGridThreads / 32;
  int grid = (K + warps_per_cta - 1) / warps_per_cta;
  if (grid > sms * per_sm) grid = sms * per_sm;
  if (grid > sms) grid = sms;  // one CTA per SM
  if (grid > 256) grid = 256;  // flag words reserved in the workspace
  for (int b = 0; b < batch; ++b) {
    MleParams pb = p;
    pb.active = active ? active + (size_t)b * K : nullptr;
    pb.T = T + (size_t)b * K * K;
    pb.pi = pi + (size_t)b * K;
    pb.info = p.info + 2 * b;
    pb.S = p.S + (size_t)b * K * K;
    pb.cvec = p.cvec + (size_t)b * K;
    PMB_CUDA(cudaMemsetAsync(p.flags, 0, (size_t)256 * kFlagStride * sizeof(unsigned long long), st));
    void* args[] = {&pb};
    const bool reg = (K <= 1024) && ((long long)grid * warps_per_cta >= K);
    PMB_CUDA(cudaLaunchCooperativeKernel(reg ? (void*)mle_grid_kernel<true> : (void*)mle_grid_kernel<false>,
                                         dim3(grid), dim3(kMleGridThreads), args,
                                         smem, st));
    count_launch();
  }
  return PMB_OK;
}

extern "C" int pmb_debug_counters(int64_t* out8) {
  using namespace pmb;
  PMB_REQUIRE(out8 != nullptr, "pmb_debug_counters: null pointer");
  long long h[8];
  PMB_CUDA(cudaMemcpyFromSymbol(h, g_dbg, sizeof(h)));
  for (int i = 0; i < 8; ++i) out8[i] = h[i];
  return PMB_OK;
}
