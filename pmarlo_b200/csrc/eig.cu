// eig.cu -- K9: leading eigenvalues of a reversible transition matrix.
//
// A = D^{1/2} T D^{-1/2} (D = diag(pi)) is symmetric for a reversible T, so its
// spectrum is real; states with pi_i = 0 (outside the active set) contribute a
// zero row/column and therefore only zero eigenvalues, which sort last.
//  * K <= kEigJacobiMaxK: A is formed explicitly and handed to the batched
//    one-sided Jacobi solver of tica.cu (one CTA per matrix) -- the ITS sweep
//    path, one CTA per lag time;
//  * larger K: Lanczos with partial re-orthogonalisation (pairs of full
//    two-pass Gram-Schmidt steps whenever a scalar bound on the lost
//    orthogonality nears sqrt(eps), the three-term recurrence in between) in
//    ONE cooperative kernel: the mat-vec streams T once per step
//    (8 K^2 bytes, L2-resident for K = 1000), a step costs 3-5 grid barriers and
//    is latency-bound; the Ritz values wanted are the two ends of the spectrum
//    of the m x m tridiagonal matrix: one warp per eigenvalue, 32-way
//    multisection with a division-free Sturm count (sturm_count_minors); the
//    one-thread-per-eigenvalue quotient-form bisection remains for k > 32.
//    Convergence flag: the k leading Ritz values of T_m and of T_{m - m/8}
//    agree to 1e-10.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace pmb {

constexpr int kEigJacobiMaxK = 256;
constexpr double kLanOmegaReset = 1e-15;  // orthogonality level right after a pair of full steps
constexpr double kLanOmegaMax = 1e-8;     // semi-orthogonality bound (sqrt(eps)) that local steps must keep
constexpr int kLanThreads = 256;

int sym_eigvals_launch(double* A, int n, int batch, double* evals, double* scratch, int* order,
                       cudaStream_t st);  // tica.cu

__global__ void eig_build_sym_kernel(const double* __restrict__ T, const double* __restrict__ pi, int K,
                                     double* __restrict__ A) {
  const size_t b = blockIdx.z;
  const int i = blockIdx.y * blockDim.y + threadIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K || j >= K) return;
  const double pi_i = pi[b * K + i], pi_j = pi[b * K + j];
  double v = 0.0;
  if (pi_i > 0.0 && pi_j > 0.0) {
    const double si = sqrt(pi_i), sj = sqrt(pi_j);
    const double a = si * T[b * K * K + (size_t)i * K + j] / sj;
    const double at = sj * T[b * K * K + (size_t)j * K + i] / si;
    v = 0.5 * (a + at);
  }
  A[b * K * K + (size_t)i * K + j] = v;
}

__global__ void eig_take_topk_kernel(const double* __restrict__ full, int K, int k, double* __restrict__ out,
                                     long long* __restrict__ info) {
  const size_t b = blockIdx.x;
  for (int t = threadIdx.x; t < k; t += blockDim.x) out[b * k + t] = (t < K) ? full[b * K + t] : 0.0;
  if (threadIdx.x == 0) { info[2 * b] = 0; info[2 * b + 1] = 1; }
}

// number of eigenvalues of the tridiagonal (a, b) of size m that are < x
__device__ __forceinline__ int sturm_count(const double* a, const double* b, int m, double x, double tiny) {
  int cnt = 0;
  double d = a[0] - x;
  if (d == 0.0) d = -tiny;
  cnt += d < 0.0;
  for (int i = 1; i < m; ++i) {
    d = a[i] - x - (b[i - 1] * b[i - 1]) / d;
    if (d == 0.0) d = -tiny;
    cnt += d < 0.0;
  }
  return cnt;
}

// all eigenvalues (ascending) of the m x m tridiagonal; thread t -> eigenvalue t
__device__ void tridiag_bisect(const double* a, const double* b, int m, double* out) {
  double lo = 1e300, hi = -1e300;
  for (int i = 0; i < m; ++i) {
    const double r = (i > 0 ? fabs(b[i - 1]) : 0.0) + (i + 1 < m ? fabs(b[i]) : 0.0);
    lo = fmin(lo, a[i] - r);
    hi = fmax(hi, a[i] + r);
  }
  const double span = fmax(fabs(lo), fabs(hi));
  const double tiny = 2.3e-308 + 1e-30 * span;
  lo -= 1e-12 * span + 1e-300;
  hi += 1e-12 * span + 1e-300;
  for (int t = threadIdx.x; t < m; t += blockDim.x) {
    double l = lo, h = hi;
    for (int it = 0; it < 200; ++it) {
      const double mid = 0.5 * (l + h);
      if (mid <= l || mid >= h) break;
      if (sturm_count(a, b, m, mid, tiny) > t) h = mid; else l = mid;
    }
    out[t] = 0.5 * (l + h);
  }
}

// Number of eigenvalues of the tridiagonal (a, b) of size m below x, from the sign changes of the leading
// principal minors p_i(x) = (a_i - x) p_{i-1} - b_{i-1}^2 p_{i-2}: ONE dependent fma per row.  The quotient form
// d_i = p_i / p_{i-1} of sturm_count above puts a double-precision division (~110 cycles of latency) on the
// chain; with it the Ritz values took half of the Lanczos kernel.  The problem is scaled by `inv` = 1 / (its
// Gershgorin span) so that a minor changes by at most 2^24 in eight rows, and the pair is renormalised every
// eight rows.  A minor that is exactly zero takes the sign opposite to its predecessor (the d = -tiny rule).
// The scaled rows are prepared once per call in shared memory, eight rows are unrolled between renormalisations.
__device__ __forceinline__ int sturm_count_minors(const double* as, const double* bb, int m, double xs) {
  // as[i] = a_i inv, bb[i] = (b_{i-1} inv)^2 (bb[0] unused), xs = x inv
  double p0 = 1.0, p1 = as[0] - xs;
  bool neg = p1 < 0.0 || p1 == 0.0;   // sign of p_1 (p_0 > 0)
  int cnt = neg ? 1 : 0;
  auto row = [&](int i) {
    const double pn = fma(as[i] - xs, p1, -bb[i] * p0);
    p0 = p1;
    p1 = pn;
    const bool ng = pn < 0.0 || (pn == 0.0 && !neg);
    cnt += (ng != neg) ? 1 : 0;
    neg = ng;
  };
  int i = 1;
  for (; i + 8 <= m; i += 8) {
#pragma unroll
    for (int u = 0; u < 8; ++u) row(i + u);
    const double mx = fmax(fabs(p0), fabs(p1));
    if (mx > 0x1p256) { p0 *= 0x1p-256; p1 *= 0x1p-256; }
    else if (mx < 0x1p-256) { p0 *= 0x1p256; p1 *= 0x1p256; }
  }
  for (; i < m; ++i) row(i);
  return cnt;
}

// the kk smallest and kk largest eigenvalues of the m x m tridiagonal: out[0..kk) ascending from the bottom,
// out[kk..2kk) descending from the top.  One WARP per eigenvalue: the 32 lanes evaluate the Sturm count at 32
// points of the bracket at once (a factor 33 per pass, 11-12 passes to the last bit instead of ~60 bisections).
// Every lane of every warp computes the same brackets from the same numbers: the result is uniform.
__device__ void tridiag_bisect_extremes(const double* a, const double* b, int m, int kk, double* out,
                                        double* as, double* bb /* shared scratch, m entries each */) {
  double lo = 1e300, hi = -1e300;
  for (int i = 0; i < m; ++i) {
    const double r = (i > 0 ? fabs(b[i - 1]) : 0.0) + (i + 1 < m ? fabs(b[i]) : 0.0);
    lo = fmin(lo, a[i] - r);
    hi = fmax(hi, a[i] + r);
  }
  const double span = fmax(fabs(lo), fabs(hi));
  const double inv = span > 0.0 ? 1.0 / span : 1.0;
  lo -= 1e-12 * span + 1e-300;
  hi += 1e-12 * span + 1e-300;
  __syncthreads();   // scratch free (previous call)
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    as[i] = a[i] * inv;
    const double bi = i > 0 ? b[i - 1] * inv : 0.0;
    bb[i] = bi * bi;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int t = threadIdx.x >> 5; t < 2 * kk; t += nw) {
    const int idx = t < kk ? t : m - 1 - (t - kk);
    double v = 0.0;
    if (idx >= 0 && idx < m) {
      double l = lo, h = hi;
      for (int pass = 0; pass < 16; ++pass) {
        const double w = (h - l) * (1.0 / 33.0);
        const double x_first = l + w, x_last = l + 32.0 * w;
        if (!(x_first > l) || !(x_last < h)) break;   // the grid no longer separates the bracket
        const int c = sturm_count_minors(as, bb, m, (l + (double)(lane + 1) * w) * inv);
        const unsigned above = __ballot_sync(0xffffffffu, c > idx);   // points with more than idx eigenvalues below
        if (above == 0u) {
          l = x_last;
        } else {
          const int f = __ffs(above) - 1;
          h = l + (double)(f + 1) * w;
          if (f > 0) l = l + (double)f * w;
        }
      }
      v = 0.5 * (l + h);
    }
    if (lane == 0) out[t] = v;
  }
}

// the k leading eigenvalues by magnitude out of the two sorted ends produced above (needs 2 k <= m)
__device__ __forceinline__ void merge_extremes_by_magnitude(const double* ext, int k, double* top) {
  int lo_i = 0, hi_i = 0;
  for (int t = 0; t < k; ++t) {
    const double a_lo = ext[lo_i], a_hi = ext[k + hi_i];
    if (fabs(a_hi) >= fabs(a_lo)) { top[t] = a_hi; ++hi_i; } else { top[t] = a_lo; ++lo_i; }
  }
}

constexpr int kLanCheckFrom = 48;    // first convergence checkpoint (Lanczos steps done)
constexpr int kLanCheckEvery = 16;   // ... and the distance between checkpoints
constexpr int kLanMaxTop = 32;       // early stopping watches at most this many leading Ritz values

struct LanParams {
  const double* T;
  const double* pi;
  int K, k, m;
  double* evals;       // k
  long long* info;     // 2
  double* V;           // (m+1) x K
  double* w;           // K
  double* z;           // K (second w buffer)
  double* rs;          // 2 x K: 1/sqrt(pi), sqrt(pi)
  double* h;           // 2 x (m+1)
  double* alpha;       // m
  double* beta;        // m
  double* part;        // gridDim.x
  double* ritz;        // 2 x m
};

// Sum of the per-CTA partials, evaluated by every warp with the loads spread over its lanes (fixed lane
// assignment and shuffle tree: the same value in every warp of the grid).  A per-thread sequential loop over
// the 125 partials was a chain of 125 dependent L2 round trips and took a quarter of the kernel (ncu).
__device__ __forceinline__ double grid_sum_partials(const double* part, int n) {
  double r = 0.0;
  for (int i = threadIdx.x & 31; i < n; i += 32) r += __ldcg(part + i);
  return warp_sum(r);
}

__global__ void __launch_bounds__(kLanThreads, 1) lanczos_kernel(LanParams p) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double s_red[32];
  __shared__ int s_order[1024];
  // every CTA keeps its own copy of the recurrence (a_j, b_j are the same numbers in every thread of the grid),
  // so that the convergence checkpoints need no extra grid barrier
  __shared__ double s_alpha[1024], s_beta[1024];
  __shared__ double s_as[1024], s_bb[1024];   // scaled copies for the Sturm counts
  __shared__ double s_ext[2 * kLanMaxTop], s_top[kLanMaxTop], s_top_prev[kLanMaxTop];
  __shared__ int s_stop;
  const int K = p.K, tid = threadIdx.x, lane = tid & 31;
  const int gtid = blockIdx.x * blockDim.x + tid, gthreads = gridDim.x * blockDim.x;
  const int gwarp = gtid >> 5, nwarps = gthreads >> 5;

  auto block_partial = [&](double v) {  // deterministic per-CTA sum -> part[blockIdx.x]
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) s_red[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
      double r = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += s_red[w];
      p.part[blockIdx.x] = r;
    }
  };

  // Buffers: w is double buffered (p.w, p.z); rs = 1/sqrt(pi), sq = sqrt(pi) on the active states (0 elsewhere).
  double* w_cur = p.w;
  double* w_new = p.z;
  double* rs = p.rs;
  double* sq = p.rs + K;
  // v0: deterministic, supported on the active states
  double acc = 0.0;
  for (int e = gtid; e < K; e += gthreads) {
    const double pe = p.pi[e];
    const double v = (pe > 0.0) ? 1.0 + 0.5 * sin(0.7548776662466927 * (double)(e + 1)) : 0.0;
    w_cur[e] = v;
    rs[e] = pe > 0.0 ? 1.0 / sqrt(pe) : 0.0;
    sq[e] = pe > 0.0 ? sqrt(pe) : 0.0;
    acc = fma(v, v, acc);
  }
  block_partial(acc);
  grid.sync();
  double nrm = sqrt(grid_sum_partials(p.part, gridDim.x));
  if (!(nrm > 0.0)) {
    if (gtid == 0) { p.info[0] = 0; p.info[1] = -1; }
    for (int t = gtid; t < p.k; t += gthreads) p.evals[t] = 0.0;
    return;
  }
  // The Lanczos vector v_j is never materialised on its own pass: w_cur holds the un-normalised vector and
  // `binv` its scale, and phase A of step j both stores V[j] = binv w_cur and applies the operator to it.
  double binv = 1.0 / nrm;

  // Re-orthogonalisation schedule (partial, after Simon): a pair of consecutive steps orthogonalises w against
  // ALL previous vectors, twice (CGS2, "twice is enough"); the steps in between only against v_{j-1} and v_j
  // (the three-term recurrence).  In a local step the orthogonality lost to a converged Ritz vector grows by at
  // most g_j = (max_k |a_k| + |a_j| + 2 max_k b_k) / b_j (the omega recurrence with every term at its bound), so
  // a scalar `omega` carries the bound: reset to ~eps by a full pair, multiplied by g_j per local step, and the
  // next full pair starts when one more local step could pass sqrt(eps) -- semi-orthogonality, which keeps the
  // Ritz values of T_m at full accuracy.  g_j is 1 / (the relative width of the bulk spectrum): ~40 for the
  // K = 1000 bench matrix (4 local steps per pair), ~55 for K = 2000..5000, where a fixed period of 8 lost
  // orthogonality completely after ~100 steps (55^6 eps = 3e-6 per period) and the projections diverged.
  // Every thread evaluates the rule on the same numbers: uniform without a barrier.
  // A full step costs 5 grid barriers, a local one 3.
#ifdef PMB_LAN_PROF
  long long lp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define LAN_T(v) const long long v = clock64()
#define LAN_ACC(i, a, b) lp[i] += (b) - (a)
#else
#define LAN_T(v)
#define LAN_ACC(i, a, b)
#endif
  int m_eff = 0;
  double omega = kLanOmegaReset, g_last = 1.0, a_max = 0.0, b_max = 0.0;
  int full_left = 2;
  for (int j = 0; j < p.m; ++j) {
    const double* wc = w_cur;
    double* wn = w_new;
    LAN_T(t0);
    // A: V[j] = binv w_cur;  w_new = D^{1/2} T D^{-1/2} V[j]
    {
      double* vj = p.V + (size_t)j * K;
      for (int e = gtid; e < K; e += gthreads) vj[e] = __ldcg(wc + e) * binv;
    }
    for (int i = gwarp; i < K; i += nwarps) {
      const double sqi = sq[i];
      double s = 0.0;
      if (sqi > 0.0) {
        const double* Trow = p.T + (size_t)i * K;
        // four independent accumulators: four L2 round trips in flight instead of one
        double s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int c = lane;
#pragma unroll 4
        for (; c + 96 < K; c += 128) {
          s = fma(Trow[c], __ldcg(wc + c) * rs[c], s);
          s1 = fma(Trow[c + 32], __ldcg(wc + c + 32) * rs[c + 32], s1);
          s2 = fma(Trow[c + 64], __ldcg(wc + c + 64) * rs[c + 64], s2);
          s3 = fma(Trow[c + 96], __ldcg(wc + c + 96) * rs[c + 96], s3);
        }
        for (; c < K; c += 32) s = fma(Trow[c], __ldcg(wc + c) * rs[c], s);
        s = warp_sum((s + s1) + (s2 + s3)) * sqi * binv;
      }
      if (lane == 0) wn[i] = s;
    }
    LAN_T(t1);
    grid.sync();
    LAN_T(t2);
    LAN_ACC(0, t0, t1);
    LAN_ACC(1, t1, t2);
    if (full_left == 0 && omega * 1.5 * g_last > kLanOmegaMax) full_left = 2;
    const bool full = full_left > 0;
    const int i_lo = full ? 0 : (j > 0 ? j - 1 : 0);
    const int npass = full ? 2 : 1;
    double a_j = 0.0;
    for (int pass = 0; pass < npass; ++pass) {
      double* h = p.h + (size_t)pass * (p.m + 1);
      LAN_T(t3);
      // B: h_i = v_i . w
      for (int i = i_lo + gwarp; i <= j; i += nwarps) {
        const double* vi = p.V + (size_t)i * K;
        double s = 0.0;
        double s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int c = lane;
#pragma unroll 4
        for (; c + 96 < K; c += 128) {
          s = fma(__ldcg(vi + c), __ldcg(wn + c), s);
          s1 = fma(__ldcg(vi + c + 32), __ldcg(wn + c + 32), s1);
          s2 = fma(__ldcg(vi + c + 64), __ldcg(wn + c + 64), s2);
          s3 = fma(__ldcg(vi + c + 96), __ldcg(wn + c + 96), s3);
        }
        for (; c < K; c += 32) s = fma(__ldcg(vi + c), __ldcg(wn + c), s);
        s = warp_sum((s + s1) + (s2 + s3));
        if (lane == 0) h[i] = s;
      }
      LAN_T(t4);
      grid.sync();
      LAN_T(t5);
      LAN_ACC(2, t3, t4);
      LAN_ACC(3, t4, t5);
      // C: w -= sum_i h_i v_i  (+ partial |w|^2 on the last pass)
      double nn = 0.0;
      if (full) {
        // one warp per element, the previous vectors spread over its lanes: a thread per element walked the
        // j + 1 vectors as a chain of (j + 1) / 4 dependent L2 round trips on 1000 of the grid's 32 000
        // threads (9 us per pass at j = 100); this is one round trip and a shuffle tree.  Four elements per
        // warp are in flight for K > the number of warps.
        constexpr int EU = 4;
        for (int e0 = gwarp; e0 < K; e0 += EU * nwarps) {
          double part[EU];
#pragma unroll
          for (int u = 0; u < EU; ++u) {
            const int e = e0 + u * nwarps;
            double s0 = 0.0, s1 = 0.0;
            if (e < K) {
              int i = lane;
              for (; i + 32 <= j; i += 64) {
                s0 = fma(__ldcg(h + i), __ldcg(p.V + (size_t)i * K + e), s0);
                s1 = fma(__ldcg(h + i + 32), __ldcg(p.V + (size_t)(i + 32) * K + e), s1);
              }
              if (i <= j) s0 = fma(__ldcg(h + i), __ldcg(p.V + (size_t)i * K + e), s0);
            }
            part[u] = s0 + s1;
          }
#pragma unroll
          for (int u = 0; u < EU; ++u) {
            const int e = e0 + u * nwarps;
            const double sub = warp_sum(part[u]);
            if (e < K && lane == 0) {
              const double v = __ldcg(wn + e) - sub;
              wn[e] = v;
              nn = fma(v, v, nn);
            }
          }
        }
      } else {
        for (int e = gtid; e < K; e += gthreads) {
          double v = __ldcg(wn + e);
          for (int i = i_lo; i <= j; ++i) v = fma(-__ldcg(h + i), __ldcg(p.V + (size_t)i * K + e), v);
          wn[e] = v;
          nn = fma(v, v, nn);
        }
      }
      a_j += __ldcg(h + j);
      if (pass == npass - 1) block_partial(nn);
      LAN_T(t6);
      grid.sync();
      LAN_T(t7);
      LAN_ACC(4, t5, t6);
      LAN_ACC(5, t6, t7);
    }
    LAN_T(t8);
    const double b_j = sqrt(grid_sum_partials(p.part, gridDim.x));
    if (gtid == 0) { p.alpha[j] = a_j; p.beta[j] = b_j; }
    if (tid == 0) { s_alpha[j] = a_j; s_beta[j] = b_j; }
    m_eff = j + 1;
    a_max = fmax(a_max, fabs(a_j));
    b_max = fmax(b_max, b_j);
    g_last = fmax(1.0, (a_max + fabs(a_j) + 2.0 * b_max) / b_j);
    if (full) { if (--full_left == 0) omega = kLanOmegaReset; } else omega = fmin(omega * g_last, 1.0);
    if (!(b_j > 1e-13) || j + 1 == p.m) break;  // invariant subspace / done (uniform across the grid)
    // Convergence checkpoint: the k leading (by magnitude) Ritz values of T_{j+1} against those of the previous
    // checkpoint, 16 steps earlier.  Every CTA evaluates it on its own copy of the recurrence with the same
    // arithmetic, so the decision is uniform across the grid without a barrier.  (The fixed 200 steps this
    // replaces cost 5.4 ms for K = 1000; the leading values settle after 60-100.)
    if (p.k <= kLanMaxTop && 2 * p.k <= m_eff && m_eff >= kLanCheckFrom && (m_eff - kLanCheckFrom) % kLanCheckEvery == 0) {
      __syncthreads();
      tridiag_bisect_extremes(s_alpha, s_beta, m_eff, p.k, s_ext, s_as, s_bb);
      __syncthreads();
      if (tid == 0) {
        merge_extremes_by_magnitude(s_ext, p.k, s_top);
        int stop = (m_eff > kLanCheckFrom) ? 1 : 0;
        for (int t = 0; t < p.k; ++t) {
          if (stop && fabs(s_top[t] - s_top_prev[t]) > 1e-12 * fmax(1.0, fabs(s_top[t]))) stop = 0;
          s_top_prev[t] = s_top[t];
        }
        s_stop = stop;
      }
      __syncthreads();
      if (s_stop) break;
    }
    binv = 1.0 / b_j;
    { double* t = w_cur; w_cur = w_new; w_new = t; }
    LAN_T(t9);
    LAN_ACC(6, t8, t9);
  }
  grid.sync();
  if (blockIdx.x != 0) return;
#ifdef PMB_LAN_PROF
  const long long t_tail0 = clock64();
#endif
  // Ritz values of T_m and of a shorter recurrence, CTA 0 only
  const int m2 = m_eff - (m_eff / 8 > 1 ? m_eff / 8 : 1);
  int ok = 1;
  if (p.k <= kLanMaxTop && 2 * p.k <= m2) {
    // only the two ends of the spectrum are needed: k values from each, merged by magnitude
    __syncthreads();
    tridiag_bisect_extremes(s_alpha, s_beta, m_eff, p.k, s_ext, s_as, s_bb);
    __syncthreads();
    if (tid == 0) merge_extremes_by_magnitude(s_ext, p.k, s_top);
    __syncthreads();
    tridiag_bisect_extremes(s_alpha, s_beta, m2, p.k, s_ext, s_as, s_bb);
    __syncthreads();
    if (tid == 0) merge_extremes_by_magnitude(s_ext, p.k, s_top_prev);
    __syncthreads();
    for (int t = tid; t < p.k; t += blockDim.x) {
      const double a = s_top[t], b = s_top_prev[t];
      p.evals[t] = a;
      if (fabs(a - b) > 1e-10 * fmax(1.0, fabs(a))) ok = 0;
    }
  } else {
    double* r1 = p.ritz;
    double* r2 = p.ritz + p.m;
    for (int i = tid; i < m_eff; i += blockDim.x) { r1[i] = 0.0; r2[i] = 0.0; }
    __syncthreads();
    tridiag_bisect(p.alpha, p.beta, m_eff, r1);
    if (m2 >= 1) tridiag_bisect(p.alpha, p.beta, m2, r2);
    __syncthreads();
    // top-k by magnitude of each set (m_eff <= 1024 guaranteed by the launcher)
    auto rank = [&](const double* vals, int n, int* order) {
      for (int jj = tid; jj < n; jj += blockDim.x) {
        const double aj = fabs(vals[jj]);
        int rk = 0;
        for (int i = 0; i < n; ++i) {
          const double ai = fabs(vals[i]);
          rk += (ai > aj) || (ai == aj && i < jj);
        }
        order[rk] = jj;
      }
      __syncthreads();
    };
    rank(r1, m_eff, s_order);
    for (int t = tid; t < p.k; t += blockDim.x) p.evals[t] = (t < m_eff) ? r1[s_order[t]] : 0.0;
    __syncthreads();
    if (m2 >= 1) {
      rank(r2, m2, s_order);
      for (int t = tid; t < p.k && t < m2; t += blockDim.x) {
        const double a = p.evals[t], b = r2[s_order[t]];
        if (fabs(a - b) > 1e-10 * fmax(1.0, fabs(a))) ok = 0;
      }
    }
  }
  ok = __syncthreads_and(ok);
  if (m_eff >= K || m_eff < p.m) ok = 1;  // exhausted the (active) space: Ritz values are exact
  if (tid == 0) { p.info[0] = m_eff; p.info[1] = ok; }
#ifdef PMB_LAN_PROF
  if (tid == 0) {
    lp[7] = clock64() - t_tail0;
    for (int i = 0; i < 8; ++i) p.part[256 + i] = (double)lp[i];
  }
#endif
}

static int lanczos_steps(int K, int k, int max_steps) {
  int m = max_steps > 0 ? max_steps : (10 * k > 200 ? 10 * k : 200);
  if (m > K) m = K;
  if (m > 1024) m = 1024;
  return m;
}

}  // namespace pmb

extern "C" size_t pmb_eig_rev_topk_ws_bytes(int K, int k, int batch, int max_steps) {
  using namespace pmb;
  if (K <= 0 || k <= 0 || batch <= 0) return 0;
  if (K <= kEigJacobiMaxK)
    return ((size_t)batch * K * K + 2 * (size_t)batch * K) * sizeof(double) + (size_t)batch * K * sizeof(int) + 64;
  const int m = lanczos_steps(K, k, max_steps);
  return ((size_t)(m + 1) * K + 4 * (size_t)K + 2 * (size_t)(m + 1) + 4 * (size_t)m + 1024) * sizeof(double);
}

extern "C" int pmb_eig_rev_topk(const double* T, const double* pi, int K, int k, int batch, int max_steps,
                                double* evals, int64_t* info, void* ws, size_t ws_bytes,
                                pmb_stream_t stream) {
  using namespace pmb;
  PMB_REQUIRE(K > 0 && k > 0 && batch > 0, "pmb_eig_rev_topk: bad sizes");
  PMB_REQUIRE(T && pi && evals && info && ws, "pmb_eig_rev_topk: null pointer");
  if (ws_bytes < pmb_eig_rev_topk_ws_bytes(K, k, batch, max_steps)) {
    set_error("pmb_eig_rev_topk: workspace too small");
    return PMB_EWORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  if (K <= kEigJacobiMaxK) {
    double* A = static_cast<double*>(ws);
    double* full = A + (size_t)batch * K * K;
    double* scratch = full + (size_t)batch * K;
    int* order = reinterpret_cast<int*>(scratch + (size_t)batch * K);
    dim3 b(16, 16), g((K + 15) / 16, (K + 15) / 16, batch);
    eig_build_sym_kernel<<<g, b, 0, st>>>(T, pi, K, A);
    PMB_LAUNCH_CHECK();
    int rc = sym_eigvals_launch(A, K, batch, full, scratch, order, st);
    if (rc != PMB_OK) return rc;
    eig_take_topk_kernel<<<batch, 128, 0, st>>>(full, K, k, evals, reinterpret_cast<long long*>(info));
    PMB_LAUNCH_CHECK();
    return PMB_OK;
  }
  const int m = lanczos_steps(K, k, max_steps);
  int dev = 0, sms = 0, per_sm = 0;
  PMB_CUDA(cudaGetDevice(&dev));
  PMB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  PMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lanczos_kernel, kLanThreads, 0));
  PMB_REQUIRE(per_sm >= 1, "pmb_eig_rev_topk: kernel does not fit on an SM");
  int grid = (K + (kLanThreads / 32) - 1) / (kLanThreads / 32);
  if (grid > sms) grid = sms;
  for (int b = 0; b < batch; ++b) {
    LanParams p;
    p.T = T + (size_t)b * K * K;
    p.pi = pi + (size_t)b * K;
    p.K = K; p.k = k; p.m = m;
    p.evals = evals + (size_t)b * k;
    p.info = reinterpret_cast<long long*>(info) + 2 * b;
    double* base = static_cast<double*>(ws);
    p.V = base;
    p.w = p.V + (size_t)(m + 1) * K;
    p.z = p.w + K;
    p.rs = p.z + K;
    p.h = p.rs + 2 * (size_t)K;
    p.alpha = p.h + 2 * (size_t)(m + 1);
    p.beta = p.alpha + m;
    p.ritz = p.beta + m;
    p.part = p.ritz + 2 * (size_t)m;
    void* args[] = {&p};
    PMB_CUDA(cudaLaunchCooperativeKernel((void*)lanczos_kernel, dim3(grid), dim3(kLanThreads), args, 0, st));
    count_launch();
  }
  return PMB_OK;
}
