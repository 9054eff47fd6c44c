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
// Exchange of the K-vector u between CTAs without fences (NCCL "LL" style): every double is
// published as two 8-byte words, each carrying 32 data bits and a 32-bit generation tag; an aligned
// 8-byte store is single-copy atomic, so a consumer that sees the expected tag in both words has the
// whole value.  Consumers spin on the data itself: one L2 round trip after the producer's store.
__device__ __forceinline__ void ll_store(unsigned long long* slot, double v, unsigned int tag) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  const unsigned long long w0 = (bits & 0xffffffffull) | ((unsigned long long)tag << 32);
  const unsigned long long w1 = (bits >> 32) | ((unsigned long long)tag << 32);
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(w0), "l"(w1) : "memory");   // gpu scope: ld/st.volatile compile to .STRONG.SYS
}
__device__ __forceinline__ bool ll_try_load(const unsigned long long* slot, unsigned int tag, double& out) {
  unsigned long long w0, w1;
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot) : "memory");
  out = __longlong_as_double((long long)(((w1 & 0xffffffffull) << 32) | (w0 & 0xffffffffull)));
  return (unsigned int)(w0 >> 32) == tag && (unsigned int)(w1 >> 32) == tag;
}
// gather entries i = tid, tid + nthreads, ... of generation `tag` into dst[]: four independent
// polls in flight per thread (a dependent chain of L2 round trips would cost ~1000 cycles each)
__device__ __forceinline__ void ll_gather(const unsigned long long* base, unsigned int tag, double* dst, int K) {
  const int nt = blockDim.x;
  for (int i0 = threadIdx.x; i0 < K; i0 += 4 * nt) {
    double v[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) ok[u] = (i0 + u * nt >= K);
    unsigned int spins = 0;
    while (!(ok[0] && ok[1] && ok[2] && ok[3])) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (!ok[u]) ok[u] = ll_try_load(base + (size_t)(i0 + u * nt) * 2, tag, v[u]);
      if (++spins > (1u << 24)) __trap();  // a lost producer must not hang the GPU
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i0 + u * nt < K) dst[i0 + u * nt] = v[u];
  }
}

// reciprocal of d > 0 from an fp32 seed r0 ~ 1/d (rel. error <= 2^-21): ONE Newton step in fp64
// (rel. error <= 2^-42); the seed is widened to fp64 with integer ops, not an fp64-pipe convert.
__device__ __forceinline__ double widen_pos(float f) {
  const unsigned int fb = __float_as_uint(f);
  return __hiloint2double((int)((fb >> 3) + 0x38000000u), (int)(fb << 29));
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ double rcp_newton1(double d, float df) {
  const double r0 = widen_pos(rcp_approx(df));
  const double e = fma(-d, r0, 1.0);
  return fma(r0, e, r0);
}

__device__ long long g_dbg[8];  // phase cycle counters of the last grid-kernel run (CTA 0, thread 0)

// REG: one row per warp with K <= 1024: the warp keeps its row of S in registers (32 doubles per
// lane) for the whole iteration, so an iteration touches no memory but the K-vector exchange.
// The kernel keeps the fp64 work per matrix element at 4 instructions (add, 2 Newton FMAs, accumulate
// FMA; B200 issues 57 fp64 FMA/clk/SM, profiles/r01a_ubench.log); an iteration is bound by the
// latency of the K-vector exchange between the CTAs.
// The streamed variant (REG = false: K > 1024, the rows of S come from L2 / HBM every iteration) runs up to 1024
// threads per CTA with eight independent loads in flight per lane: with 256 threads and two loads it reached
// 0.96 TB/s at K = 5000 (latency-bound: 0.6 MB in flight on the whole GPU), the 200 MB of S per iteration want HBM
// speed.  The launcher sizes the CTAs so that every warp owns the same number of rows (K = 5000: 17 warps per CTA,
// two rows each): a last round with a few busy warps is latency-bound again and costs as much as a full one.
constexpr int kMleStreamThreads = 1024;
template <bool REG>
__global__ void __launch_bounds__(REG ? kMleGridThreads : kMleStreamThreads) mle_grid_kernel(MleParams p) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double s_red[32];
  const int K = p.K;
  double* x = sm;
  double* q = sm + K;
  double* cs = sm + 2 * K;   // row sums of the (regularised) counts, staged once
  double* us = sm + 3 * K;   // gathered u of the current generation
  float* qf = reinterpret_cast<float*>(sm + 4 * K);   // fp32 image of q (reciprocal seeds)
  const int tid = threadIdx.x, lane = tid & 31;
  const int gwarp = (blockIdx.x * blockDim.x + tid) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint8_t* act = p.active;
  unsigned long long* xb = reinterpret_cast<unsigned long long*>(p.ubuf);   // 2 generations x K x 2 words
  for (int i = tid; i < K; i += blockDim.x) cs[i] = p.cvec[i];
  double sreg[32];
  if constexpr (REG) {
#pragma unroll
    for (int t = 0; t < 32; ++t) {
      const int j = lane + 32 * t;
      sreg[t] = (gwarp < K && j < K) ? p.S[(size_t)gwarp * K + j] : 0.0;
    }
  }
  // sum_j S_ij / (q_i + q_j) for row i (warp-wide)
  auto apply_row = [&](int i, double qi, float qfi) -> double {
    double a0 = 0.0, a1 = 0.0;
    if constexpr (REG) {
      // No per-element branch: entries beyond K have sreg = 0 and read a clamped (valid, positive) q, so they
      // add exactly 0.  With `if (j < K)` around every element ptxas emitted a branch per element and the 32
      // reciprocal chains of a lane ran one after the other (measured: 4 266 cycles per row, 133 per
      // element = the length of one LDS -> MUFU -> Newton -> FMA chain).  Batches of four independent chains;
      // the accumulation order (even entries -> a0, odd -> a1) is unchanged, so the result is bit-identical.
#pragma unroll
      for (int t = 0; t < 32; t += 4) {
        double den[4];
        float denf[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          int j = lane + 32 * (t + u);
          j = j < K ? j : K - 1;
          den[u] = qi + q[j];
          denf[u] = qfi + qf[j];
        }
        double r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) r[u] = rcp_newton1(den[u], denf[u]);
        a0 = fma(sreg[t], r[0], a0);
        a1 = fma(sreg[t + 1], r[1], a1);
        a0 = fma(sreg[t + 2], r[2], a0);
        a1 = fma(sreg[t + 3], r[3], a1);
      }
    } else {
      const double* __restrict__ Srow = p.S + (size_t)i * K;
      int j = lane;
      // eight loads in flight; the accumulation order (even entries -> a0, odd -> a1) is that of the two-at-a-time
      // loop it replaces, so the iterates are bit-identical
      for (; j + 224 < K; j += 256) {
        double sv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) sv[u] = __ldcs(Srow + j + 32 * u);
#pragma unroll
        for (int u = 0; u < 8; u += 2) {
          a0 = fma(sv[u], rcp_newton1(qi + q[j + 32 * u], qfi + qf[j + 32 * u]), a0);
          a1 = fma(sv[u + 1], rcp_newton1(qi + q[j + 32 * u + 32], qfi + qf[j + 32 * u + 32]), a1);
        }
      }
      for (; j + 32 < K; j += 64) {
        const double s0 = Srow[j], s1 = Srow[j + 32];
        a0 = fma(s0, rcp_newton1(qi + q[j], qfi + qf[j]), a0);
        a1 = fma(s1, rcp_newton1(qi + q[j + 32], qfi + qf[j + 32]), a1);
      }
      if (j < K) a0 = fma(Srow[j], rcp_newton1(qi + q[j], qfi + qf[j]), a0);
    }
    return warp_sum(a0 + a1);
  };
  // publish row results of generation `gen`, gather the whole vector into us[]
  auto exchange_store = [&](unsigned int gen, int i, double v) {
    ll_store(xb + ((size_t)(gen & 1u) * K + i) * 2, v, gen);
  };
  auto exchange_gather = [&](unsigned int gen) {
    ll_gather(xb + (size_t)(gen & 1u) * K * 2, gen, us, K);
  };

  unsigned int gen = 1;
  // generation 1: row sums of S -> x0
  for (int i = gwarp; i < K; i += nwarps) {
    double acc = 0.0;
    if constexpr (REG) {
#pragma unroll
      for (int t = 0; t < 32; ++t) acc += sreg[t];
    } else {
      for (int j = lane; j < K; j += 32) acc += p.S[(size_t)i * K + j];
    }
    acc = warp_sum(acc);
    if (lane == 0) exchange_store(gen, i, acc);
  }
  exchange_gather(gen);
  double part = 0.0;
  int bad = 0;
  for (int i = tid; i < K; i += blockDim.x) {
    part += us[i];
    const bool on = !act || act[i];
    if (on && !(cs[i] > 0.0)) bad = 1;
  }
  const double tot = cta_sum(part, s_red);
  bad = __syncthreads_or(bad);
  if (bad || !(tot > 0.0)) {  // uniform across the grid: every CTA sees the same data
    if (blockIdx.x == 0 && tid == 0) { p.info[0] = 0; p.info[1] = -1; }
    return;
  }
  for (int i = tid; i < K; i += blockDim.x) {
    const double xn = us[i] / tot;
    x[i] = xn;
    const double qn = (xn > 0.0) ? cs[i] / xn : 0.0;
    q[i] = qn;
    qf[i] = (float)qn;
  }
  __syncthreads();

  long long it = 0;
  int notconv = 1;
  long long cyc[4] = {0, 0, 0, 0};
#ifdef PMB_MLE_PROF
  long long prof_apply = 0;
#endif
  while (it < p.maxiter && notconv) {
    const long long t0 = clock64();
    ++gen;
#ifdef PMB_MLE_PROF
    long long t0a = t0;
#endif
    for (int i = gwarp; i < K; i += nwarps) {
      double r = 0.0;
      if (x[i] > 0.0) r = apply_row(i, q[i], qf[i]);
#ifdef PMB_MLE_PROF
      t0a = clock64();
#endif
      if (lane == 0) exchange_store(gen, i, r);
    }
    const long long t1 = clock64();
    exchange_gather(gen);
    const long long t2 = clock64();
    part = 0.0;
    for (int i = tid; i < K; i += blockDim.x) part += us[i];
    const double norm = cta_sum(part, s_red);      // includes the barrier that orders us[] writes
#ifdef PMB_MLE_PROF
    cyc[3] += clock64() - t2;      // own entries gathered -> every thread of the CTA has its entries (stragglers)
    prof_apply += t0a - t0;        // the row product alone
#endif
    double rnorm = rcp_newton1(norm, (float)norm);  // two Newton steps: full fp64 accuracy
    rnorm = fma(rnorm, fma(-norm, rnorm, 1.0), rnorm);
    int flag = 0;
    for (int i = tid; i < K; i += blockDim.x) {
      const double xo = x[i], xn = us[i] * rnorm;
      // |xo - xn| / (0.5 (xo + xn)) > maxerr  without the division
      if (fabs(xo - xn) > p.maxerr * 0.5 * (xo + xn)) flag = 1;
      x[i] = xn;
      double qn = 0.0;
      if (xn > 0.0) qn = cs[i] * rcp_newton1(xn, (float)xn);
      q[i] = qn;
      qf[i] = (float)qn;
    }
    notconv = __syncthreads_or(flag);
    ++it;
    const long long t3 = clock64();
    cyc[0] += t1 - t0; cyc[1] += t2 - t1; cyc[2] += t3 - t2;
  }
  if (blockIdx.x == 0 && tid == 0) {
    g_dbg[0] = it; g_dbg[1] = cyc[0]; g_dbg[2] = cyc[1]; g_dbg[3] = cyc[2]; g_dbg[4] = cyc[3];
    g_dbg[5] = gridDim.x; g_dbg[6] = REG ? 1 : 0; g_dbg[7] = 0;
#ifdef PMB_MLE_PROF
    g_dbg[7] = prof_apply;
#endif
  }
  // final application with the last x: T rows and pi
  ++gen;
  for (int i = gwarp; i < K; i += nwarps) {
    const bool on = x[i] > 0.0;
    double rs = 0.0;
    if (on) rs = apply_row(i, q[i], qf[i]);
    if (lane == 0) exchange_store(gen, i, rs);
    const double qi = q[i];
    for (int j = lane; j < K; j += 32) {
      double t;
      if (on && rs > 0.0) {
        const double s = p.S[(size_t)i * K + j];
        t = (s != 0.0) ? (s / (qi + q[j])) / rs : 0.0;
      } else {
        t = (i == j) ? 1.0 : 0.0;
      }
      p.T[(size_t)i * K + j] = t;
    }
  }
  exchange_gather(gen);
  part = 0.0;
  for (int i = tid; i < K; i += blockDim.x) part += us[i];
  const double rtot = cta_sum(part, s_red);
  for (int i = blockIdx.x * blockDim.x + tid; i < K; i += gridDim.x * blockDim.x) p.pi[i] = us[i] / rtot;
  if (blockIdx.x == 0 && tid == 0) { p.info[0] = it; p.info[1] = notconv ? 0 : 1; }
}

}  // namespace pmb

extern "C" size_t pmb_mle_rev_ws_bytes(int K, int batch) {
  if (K <= 0 || batch <= 0) return 0;
  return ((size_t)batch * K * K + (size_t)batch * K + 4 * (size_t)K + 8) * sizeof(double) + 64;
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
  {
    // exchange buffer: 2 generations x K x 2 words, 16-byte aligned
    uintptr_t a = reinterpret_cast<uintptr_t>(p.cvec + (size_t)batch * K);
    a = (a + 15) & ~(uintptr_t)15;
    p.ubuf = reinterpret_cast<double*>(a);
    p.flags = nullptr;
  }
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
  const size_t smem = (size_t)4 * K * sizeof(double) + (size_t)K * sizeof(float);
  PMB_REQUIRE(smem <= 200 * 1024, "pmb_mle_rev: K=%d too large", K);
  PMB_CUDA(cudaFuncSetAttribute(mle_grid_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PMB_CUDA(cudaFuncSetAttribute(mle_grid_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  PMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mle_grid_kernel<true>, kMleGridThreads, smem));
  PMB_REQUIRE(per_sm >= 1, "pmb_mle_rev: kernel does not fit on an SM");
  int dev = 0, sms = 0;
  PMB_CUDA(cudaGetDevice(&dev));
  PMB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const bool streamed = K > 1024;
  int threads = kMleGridThreads;
  if (streamed) {
    const int max_warps = sms * (kMleStreamThreads / 32);
    const int rounds = (K + max_warps - 1) / max_warps;                  // rows per warp
    int wpc = (K + rounds * sms - 1) / (rounds * sms);                   // warps per CTA that make the rounds even
    wpc = wpc < 4 ? 4 : (wpc > kMleStreamThreads / 32 ? kMleStreamThreads / 32 : wpc);
    threads = wpc * 32;
  }
  if (streamed) {
    PMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mle_grid_kernel<false>, threads, smem));
    PMB_REQUIRE(per_sm >= 1, "pmb_mle_rev: the streamed kernel does not fit on an SM");
  }
  const int warps_per_cta = threads / 32;
  int grid = (K + warps_per_cta - 1) / warps_per_cta;
  if (grid > sms * per_sm) grid = sms * per_sm;
  if (grid > sms) grid = sms;  // one CTA per SM
  for (int b = 0; b < batch; ++b) {
    MleParams pb = p;
    pb.active = active ? active + (size_t)b * K : nullptr;
    pb.T = T + (size_t)b * K * K;
    pb.pi = pi + (size_t)b * K;
    pb.info = p.info + 2 * b;
    pb.S = p.S + (size_t)b * K * K;
    pb.cvec = p.cvec + (size_t)b * K;
    PMB_CUDA(cudaMemsetAsync(p.ubuf, 0, (size_t)4 * K * sizeof(double), st));   // tags start at 1
    void* args[] = {&pb};
    const bool reg = !streamed && ((long long)grid * warps_per_cta >= K);
    PMB_CUDA(cudaLaunchCooperativeKernel(reg ? (void*)mle_grid_kernel<true> : (void*)mle_grid_kernel<false>,
                                         dim3(grid), dim3(threads), args,
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
