// tica_grid.cu -- K4 on several SMs: the same algorithm as tica_solve_kernel (tica.cu) as ONE
// cooperative kernel.  A d = 256 Jacobi eigen-solve is latency-bound (about 3e9 fp64 operations, but
// thousands of dependent rounds): the single-CTA kernel needs 171 ms, a grid-wide cyclic ordering
// with one grid barrier per round 56 ms, the kernel below 7.8 ms (measured on B200).  The rows are
// processed in BLOCKS of 8: a CTA orthogonalises the rows of its block pair against each other
// completely in shared memory / registers before the next grid barrier, so a sweep has d / 8 grid
// barriers instead of d - 1 (see jacobi_grid).  Data that other CTAs wrote is always read with
// ld.global.cg (L2), never through L1.  The C00 solve starts from the transposed factor of a pivoted
// Cholesky factorisation instead of from C00 itself (chol_pivoted_cta: 19 -> 11 sweeps); for d <= 128 the
// kernel runs as ONE 16-CTA thread-block cluster with hardware cluster barriers (TgBar).
#include <cooperative_groups.h>
#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace pmb {

constexpr int kTgThreads = 512;   // 16 warps: one local row pair each (b of them busy in a block round)
constexpr int kTgMaxSweeps = 40;
constexpr double kTgTol = 4.5e-16;  // x sqrt(n), as in tica.cu

struct TicaGridWs {
  double *W, *V, *L, *TMP, *M, *s;
  int* order;
  int* ctrl;  // [0] rank m, [1..2] sweep flags, [4..5] big-rotation flags
};

__device__ __forceinline__ double ldg_cg(const double* p) { return __ldcg(p); }

// phase cycle counters of the last solve (CTA 0, thread 0): load, local sweep, store, grid barrier, total, block rounds
__device__ long long g_tg_dbg[8];
// trace of the last solve: [0..15] %smid of CTAs 0..15; then per Jacobi sweep (both solves, up to 32) the pair
// (clock64 cycles, globaltimer ns) spent in it on CTA 0 -- separates a slow SM clock from time outside the SM
__device__ long long g_tg_trace[16 + 64];
__device__ __forceinline__ long long tg_globaltimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// 1/sqrt(x) for x in [2^-8, 2^8]: fp32 seed (MUFU) and ONE third-order step, y <- y (1 + e/2 + 3 e^2 / 8) with
// e = 1 - x y^2 ~ 2^-21, error O(e^3).  The IEEE fp64 divide / sqrt sequences cost ~300 dependent cycles
// each and the rotation sits on the critical path of every Jacobi round (ncu: the round is bound by the
// dependent fp64 chain -- 'wait' and 'short_scoreboard' stalls -- not by fp64 throughput); this is 7
// dependent instructions and accurate to a few ulp.
__device__ __forceinline__ double tg_rsqrt(double x) {
  const double y = (double)rsqrtf((float)x);
  const double e = fma(-x * y, y, 1.0);
  return fma(y * e, fma(0.375, e, 0.5), y);
}
// Rotation (c, s), t = s / c = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = (b - a) / (2 g), that
// orthogonalises two rows with squared norms a, b and inner product g != 0.  With alpha = (b - a) / 2 and
// r = sqrt(alpha^2 + g^2) the double angle has cos = |alpha| / r and sin = sign(alpha) g / r, so
//   c^2 = (1 + |alpha| / r) / 2,   s = sign(alpha) g / (2 r c):
// two reciprocal square roots, no division, no cancellation.  alpha and g are first scaled by a power of
// two so that the fp32 seeds cannot under- or overflow.
__device__ __forceinline__ void tg_rotation(double a, double b, double g, double& c, double& s, double& t_out) {
  const double alpha = 0.5 * (b - a);
  const double mx = fmax(fabs(alpha), fabs(g));
  const int ex = (__double2hiint(mx) >> 20) & 0x7ff;
  if (ex < 64 || ex > 1980) {   // denormal-range or huge operands: IEEE sequence
    const double zeta = (b - a) / (2.0 * g);
    const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    c = 1.0 / sqrt(1.0 + t * t);
    s = c * t;
    t_out = t;
    return;
  }
  const double sc = __hiloint2double((2046 - ex) << 20, 0);   // 2^(1023 - ex): mx * sc in [1, 2)
  const double as = alpha * sc, gs = g * sc;
  const double y = tg_rsqrt(fma(as, as, gs * gs));            // 1 / r, r^2 in [1, 8)
  const double c2 = fma(0.5 * fabs(as), y, 0.5);              // [0.5, 1]
  const double z = tg_rsqrt(c2);
  c = c2 * z;
  const double h = 0.5 * gs * y * z;
  s = (__double2hiint(alpha) < 0) ? -h : h;
  t_out = s * z;   // z = 1 / c
}

// One-sided BLOCK Jacobi on the rows of W (n x n, symmetric input); the rotations are not accumulated.
// The rows are cut into blocks of `b`.  A sweep visits every row pair exactly once:
//   A. every block on its own CTA: the b (b - 1) / 2 pairs inside the block (b - 1 local rounds of a
//      round-robin tournament, b / 2 warps busy), then one grid barrier;
//   B. nb - 1 block rounds (round-robin tournament over the blocks), a CTA per block PAIR: it loads its
//      2 b rows into shared memory and orthogonalises every row of one block against every row of the
//      other -- b local rounds, in round lr row i meets row (i + lr) mod b, one warp per pair, separated
//      by __syncthreads -- writes them back, grid barrier.
// The sequential depth of a sweep is (nb - 1) b + (b - 1) = n - 1 local rounds, the minimum for n rows,
// with nb grid barriers instead of n - 1, and every rotation works on shared memory instead of L2.
// (Sweeping ALL pairs of the 2 b resident rows in every block round, 2 b - 1 local rounds, repeats the
// in-block pairs nb - 1 times per sweep: 1.8x the depth for the same number of sweeps.)
// NPL = elements of a row per lane (n <= 32 NPL): the loops are fully unrolled, a warp holds its two
// rows in registers for the duration of a rotation and the global loads are batched.
// The barrier between rounds.  Only d / (2 b) <= 16 CTAs own a block pair, so the whole solve runs on ONE
// 16-CTA thread-block cluster when the device can place it: barrier.cluster (hardware, release / acquire at
// cluster scope, which orders the global-memory traffic between the CTAs of the cluster) instead of the
// cooperative-groups grid barrier over 148 CTAs (an atomic counter in L2 and a polling loop, ~1.5 us per
// barrier, 960 barriers per solve).  cluster == false is the cooperative launch (fallback).
struct TgBar {
  bool cluster;
  __device__ __forceinline__ void sync() const {
    if (cluster) {
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
      cg::this_grid().sync();
    }
  }
};

template <int NPL>
__device__ int jacobi_grid(TgBar& grid, double* W, int n, int b, int* flags, double* sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  int nb = (n + b - 1) / b;
  if (nb & 1) ++nb;
  if (nb < 2) nb = 2;
  const int npairs = nb >> 1;
  double* sW = sm;
  const double tol2 = kTgTol * kTgTol * (double)n;
  // flags[0..1]: "some pair was rotated" per sweep parity; flags[3..4]: "some rotated pair had a relative
  // off-diagonal |g| / sqrt(a b) above 1e-10", also per sweep parity
  int* bigflags = flags + 3;
  if (blockIdx.x == 0 && threadIdx.x == 0) { flags[0] = 0; flags[1] = 0; bigflags[0] = 0; bigflags[1] = 0; }
  grid.sync();
  int sweep = 0;
  long long cyc[4] = {0, 0, 0, 0}, n_rounds = 0;
  int rotated = 0, big = 0;

  // global row of local row li when blocks bi (local rows 0..b-1) and bj (b..2b-1) are resident
  auto grow = [&](int bi, int bj, int li) { return li < b ? bi * b + li : bj * b + (li - b); };
  auto load_rows = [&](int bi, int bj, int rows) {
    for (int li = warp; li < rows; li += nwarps) {
      const int gr = grow(bi, bj, li);
      if (gr < n) {
        double tw[NPL];
#pragma unroll
        for (int q = 0; q < NPL; ++q) {
          const int e = lane + 32 * q;
          tw[q] = e < n ? ldg_cg(W + (size_t)gr * n + e) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < NPL; ++q) {
          const int e = lane + 32 * q;
          if (e < n) sW[(size_t)li * n + e] = tw[q];
        }
      }
    }
  };
  auto store_rows = [&](int bi, int bj, int rows) {
    for (int li = warp; li < rows; li += nwarps) {
      const int gr = grow(bi, bj, li);
      if (gr < n) {
#pragma unroll
        for (int q = 0; q < NPL; ++q) {
          const int e = lane + 32 * q;
          if (e < n) __stcg(W + (size_t)gr * n + e, sW[(size_t)li * n + e]);
        }
      }
    }
  };
  // orthogonalise local rows lp, lq (global rows gp, gq, both < n)
  auto rotate_pair = [&](int lp, int lq, int gp, int gq) {
    if (gp > gq) { const int t = lp; lp = lq; lq = t; }
    double* wp = sW + (size_t)lp * n;
    double* wq = sW + (size_t)lq * n;
    double x[NPL], y[NPL];
    double a = 0.0, bb = 0.0, g = 0.0;
#pragma unroll
    for (int q = 0; q < NPL; ++q) {
      const int e = lane + 32 * q;
      x[q] = e < n ? wp[e] : 0.0;
      y[q] = e < n ? wq[e] : 0.0;
    }
    {
      // two accumulators per sum: half the dependent fp64 chain
      double a1 = 0.0, b1 = 0.0, g1 = 0.0;
#pragma unroll
      for (int q = 0; q + 1 < NPL; q += 2) {
        a = fma(x[q], x[q], a);
        bb = fma(y[q], y[q], bb);
        g = fma(x[q], y[q], g);
        a1 = fma(x[q + 1], x[q + 1], a1);
        b1 = fma(y[q + 1], y[q + 1], b1);
        g1 = fma(x[q + 1], y[q + 1], g1);
      }
      if (NPL & 1) {
        a = fma(x[NPL - 1], x[NPL - 1], a);
        bb = fma(y[NPL - 1], y[NPL - 1], bb);
        g = fma(x[NPL - 1], y[NPL - 1], g);
      }
      a += a1;
      bb += b1;
      g += g1;
    }
    a = warp_sum(a);
    bb = warp_sum(bb);
    g = warp_sum(g);
    if (g * g <= tol2 * (a * bb) || g == 0.0) return;
    rotated = 1;
    if (g * g > 1e-20 * (a * bb)) big = 1;
    double c, s, t_unused;
    tg_rotation(a, bb, g, c, s, t_unused);
    // de Rijk's ordering: the row with the larger norm goes first (rotation followed by a swap when
    // a < b); helps on graded matrices such as a covariance with eigenvalues over 8 decades.
    const bool sw = a < bb;
    const double c0 = sw ? s : c, s0 = sw ? c : -s;    // row p <- c0 x + s0 y
    const double c1 = sw ? c : s, s1 = sw ? -s : c;    // row q <- c1 x + s1 y
#pragma unroll
    for (int q = 0; q < NPL; ++q) {
      const int e = lane + 32 * q;
      if (e < n) {
        wp[e] = c0 * x[q] + s0 * y[q];
        wq[e] = c1 * x[q] + s1 * y[q];
      }
    }
  };

  for (; sweep < kTgMaxSweeps; ++sweep) {
    rotated = 0;
    big = 0;
    const long long sw_c0 = clock64(), sw_t0 = tg_globaltimer();
    // A. pairs inside a block
    long long t0 = clock64();
    if (b >= 2) {
      for (int k = blockIdx.x; k < nb; k += gridDim.x) {
        load_rows(k, k, b);
        __syncthreads();
        const long long t1 = clock64();
        cyc[0] += t1 - t0;
        const int lm = b - 1;   // tournament over b players (b is a power of two)
        for (int lr = 0; lr < lm; ++lr) {
          for (int kk = warp; kk < (b >> 1); kk += nwarps) {
            int lp, lq;
            if (kk == 0) { lp = lm; lq = lr; }
            else { lp = (lr + kk) % lm; lq = (lr - kk + lm) % lm; }
            const int gp = k * b + lp, gq = k * b + lq;
            if (gp < n && gq < n) rotate_pair(lp, lq, gp, gq);
          }
          __syncthreads();
        }
        const long long t2 = clock64();
        cyc[1] += t2 - t1;
        store_rows(k, k, b);
        __syncthreads();
        t0 = clock64();
        cyc[2] += t0 - t2;
      }
      grid.sync();
      cyc[3] += clock64() - t0;
      ++n_rounds;
    }
    // B. pairs across two blocks
    for (int r = 0; r < nb - 1; ++r) {
      t0 = clock64();
      for (int k = blockIdx.x; k < npairs; k += gridDim.x) {
        int bi, bj;
        if (k == 0) { bi = nb - 1; bj = r; }
        else { bi = (r + k) % (nb - 1); bj = (r - k + (nb - 1)) % (nb - 1); }
        if (bi > bj) { const int t = bi; bi = bj; bj = t; }
        load_rows(bi, bj, 2 * b);
        __syncthreads();
        const long long t1 = clock64();
        cyc[0] += t1 - t0;
        if (b <= nwarps) {
          // Warp kk keeps row kk of block bi in REGISTERS for all b local rounds (its partner changes, it
          // does not), and the squared norms travel with the rows -- exact at the start of the block round,
          // then updated by a' = a - t g, b' = b + t g -- so a local round moves one row instead of two
          // through shared memory and reduces one inner product instead of three.  (ncu: the round was
          // bound by the LDS / STS / SHFL issue rate of the SM, not by fp64 throughput.)
          double* sN = sW + (size_t)2 * b * n;   // b squared norms of the rows of bj
          const int kk = warp;
          const int gp = bi * b + kk;
          const bool have_p = kk < b && gp < n;
          double x[NPL];
          double a = 0.0;
          if (kk < b) {
            double bn = 0.0;
            const double* wp = sW + (size_t)kk * n;
            const double* wq = sW + (size_t)(b + kk) * n;
#pragma unroll
            for (int q = 0; q < NPL; ++q) {
              const int e = lane + 32 * q;
              x[q] = e < n ? wp[e] : 0.0;
              const double yv = e < n ? wq[e] : 0.0;
              a = fma(x[q], x[q], a);
              bn = fma(yv, yv, bn);
            }
            a = warp_sum(a);
            bn = warp_sum(bn);
            if (lane == 0) sN[kk] = bn;
          }
          __syncthreads();
          for (int lr = 0; lr < b; ++lr) {
            int lq = kk + lr;
            if (lq >= b) lq -= b;
            if (have_p && bj * b + lq < n) {
              double* wq = sW + (size_t)(b + lq) * n;
              const double bb = sN[lq];
              double y[NPL];
              double g = 0.0, g1 = 0.0;
#pragma unroll
              for (int q = 0; q < NPL; ++q) {
                const int e = lane + 32 * q;
                y[q] = e < n ? wq[e] : 0.0;
              }
#pragma unroll
              for (int q = 0; q + 1 < NPL; q += 2) {
                g = fma(x[q], y[q], g);
                g1 = fma(x[q + 1], y[q + 1], g1);
              }
              if (NPL & 1) g = fma(x[NPL - 1], y[NPL - 1], g);
              g = warp_sum(g + g1);
              if (!(g * g <= tol2 * (a * bb) || g == 0.0)) {
                rotated = 1;
                if (g * g > 1e-20 * (a * bb)) big = 1;
                double c, s, t;
                tg_rotation(a, bb, g, c, s, t);
                const bool sw = a < bb;                            // de Rijk: the larger norm goes to row p
                const double c0 = sw ? s : c, s0 = sw ? c : -s;    // row p <- c0 x + s0 y
                const double c1 = sw ? c : s, s1 = sw ? -s : c;    // row q <- c1 x + s1 y
#pragma unroll
                for (int q = 0; q < NPL; ++q) {
                  const int e = lane + 32 * q;
                  const double xn = c0 * x[q] + s0 * y[q];
                  if (e < n) wq[e] = c1 * x[q] + s1 * y[q];
                  x[q] = xn;
                }
                const double lo = fmax(a - t * g, 0.0), hi = fmax(bb + t * g, 0.0);   // |c x - s y|^2, |s x + c y|^2
                a = sw ? hi : lo;
                if (lane == 0) sN[lq] = sw ? lo : hi;
              }
            }
            __syncthreads();
          }
          if (have_p) {
            double* wp = sW + (size_t)kk * n;
#pragma unroll
            for (int q = 0; q < NPL; ++q) {
              const int e = lane + 32 * q;
              if (e < n) wp[e] = x[q];
            }
          }
          __syncthreads();
        } else {
          for (int lr = 0; lr < b; ++lr) {
            for (int kk = warp; kk < b; kk += nwarps) {
              int lq = kk + lr;
              if (lq >= b) lq -= b;
              const int gp = bi * b + kk, gq = bj * b + lq;
              if (gp < n && gq < n) rotate_pair(kk, b + lq, gp, gq);
            }
            __syncthreads();
          }
        }
        const long long t2 = clock64();
        cyc[1] += t2 - t1;
        store_rows(bi, bj, 2 * b);
        __syncthreads();
        t0 = clock64();
        cyc[2] += t0 - t2;
      }
      grid.sync();
      cyc[3] += clock64() - t0;
      ++n_rounds;
    }
    if (rotated && lane == 0) {
      atomicOr(&flags[sweep & 1], 1);
      if (big) atomicOr(&bigflags[sweep & 1], 1);
    }
    grid.sync();
    const int any = __ldcg(&flags[sweep & 1]);
    const int anybig = __ldcg(&bigflags[sweep & 1]);
    if (blockIdx.x == 0 && threadIdx.x == 0) { flags[(sweep + 1) & 1] = 0; bigflags[(sweep + 1) & 1] = 0; }
    grid.sync();
    // Converged when nothing was rotated, or when every rotation of this sweep started from a relative
    // off-diagonal below 1e-10: cyclic Jacobi converges quadratically, so the sweep leaves ~1e-20, far
    // below the 7e-15 threshold, and the confirmation sweep (dot products only, ~60% of a sweep) is skipped.
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      const int slot = (int)g_tg_dbg[7];
      if (slot < 32) {
        g_tg_trace[16 + 2 * slot] = clock64() - sw_c0;
        g_tg_trace[16 + 2 * slot + 1] = tg_globaltimer() - sw_t0;
        g_tg_dbg[7] = slot + 1;
      }
    }
    if (!any || !anybig) { ++sweep; break; }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int q = 0; q < 4; ++q) g_tg_dbg[q] += cyc[q];
    g_tg_dbg[4] += n_rounds;
  }
  return sweep;
}

// Pivoted Cholesky A = G G^T of a symmetric positive semidefinite n x n matrix, on ONE CTA, without
// interchanges: a step takes a large residual diagonal entry p, g = (A[:, p] - sum_{j<k} g_j g_j[p]) /
// sqrt(residual_pp), entries of already chosen indices are exactly zero.  W[k][:] = g_k^T for k < r (the numerical
// rank: the point at which the largest residual diagonal falls below n eps max diag), zero rows after.
// Why: the one-sided Jacobi sweeps orthogonalise the ROWS of W.  Started from W = A their Gram matrix is A^2 (the
// condition number squared: 18-19 sweeps for the bench's C00, cond ~1e8); started from W = G^T the Gram matrix
// G^T G is similar to A itself and pivoting has already ordered and nearly orthogonalised the rows: 11 sweeps
// (Drmac / Veselic preconditioning).  At convergence row j is sqrt(lambda_j) v_j^T instead of lambda_j v_j^T: the
// caller takes v_j = w_j / |w_j| and the Rayleigh quotient as before.
// Blocked by 8: the pivots of a block are the 8 largest residual diagonals at its start, their 8 columns come
// from ONE pass over the k rows already in W (a column-at-a-time version paid one chain of L2 round trips per
// column: 2.3 ms for n = 256, more than the sweeps it saved), then the block is factorised in registers / shared
// memory; a candidate whose residual has collapsed inside the block (< 1/4 of its value at the start: nearly
// dependent on an earlier pivot of the block) is left for a later block.  Any order of positive pivots gives a
// valid factorisation; the order only matters for the preconditioning.
// NQ threads per column index i (n <= 512 / NQ), each owning 8 / NQ of the block's columns.
// Shared memory: diag[n], gcol[n], gp[8][n] doubles, done[n], sel[n] ints (11 n doubles).
template <int NQ>
__device__ int chol_pivoted_cta(const double* __restrict__ A, int n, double* __restrict__ W, double* sm) {
  constexpr int B = 8, BQ = B / NQ;
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
  __shared__ double s_val[kTgThreads / 32];
  __shared__ int s_idx[kTgThreads / 32];
  __shared__ int s_cand[B];
  __shared__ double s_d0[B], s_pivot[B];
  __shared__ int s_accept[B];
  __shared__ int s_nb;
  __shared__ double s_thresh;
  double* diag = sm;
  double* gcol = sm + n;
  double* gp = sm + 2 * (size_t)n;
  int* done = reinterpret_cast<int*>(gp + (size_t)B * n);
  int* sel = done + n;
  const int i = tid % n, q = tid / n;
  const bool act = q < NQ;
  for (int e = tid; e < n; e += T) { diag[e] = ldg_cg(A + (size_t)e * n + e); done[e] = 0; }
  if (tid == 0) s_thresh = -1.0;
  __syncthreads();
  int kk = 0;
  while (kk < n) {
    // 1. up to B candidates: the largest residual diagonals among the indices not chosen yet
    for (int e = tid; e < n; e += T) sel[e] = 0;
    if (tid == 0) s_nb = 0;
    __syncthreads();
    for (int t = 0; t < B; ++t) {
      double bv = -1.0;
      int bi = 0x7fffffff;
      for (int e = tid; e < n; e += T) {
        const double v = (done[e] || sel[e]) ? -1.0 : diag[e];
        if (v > bv) { bv = v; bi = e; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) { s_val[warp] = bv; s_idx[warp] = bi; }
      __syncthreads();
      if (tid == 0) {
        double v = s_val[0];
        int ix = s_idx[0];
        for (int w = 1; w < (T >> 5); ++w)
          if (s_val[w] > v || (s_val[w] == v && s_idx[w] < ix)) { v = s_val[w]; ix = s_idx[w]; }
        if (s_thresh < 0.0) s_thresh = fmax(v, 0.0) * (double)n * 2.220446049250313e-16;
        if (v > s_thresh && ix < n) { s_cand[t] = ix; s_d0[t] = v; sel[ix] = 1; s_nb = t + 1; }
      }
      __syncthreads();
      if (s_nb != t + 1) break;   // uniform: nothing left above the threshold
    }
    const int nb = s_nb;
    if (nb == 0) break;
    const double thresh = s_thresh;
    // 2. the pivots' rows of G (gp[t][j] = W[j][p_t], j < kk) and their columns of A
    for (int e = tid; e < nb * kk; e += T) {
      const int t = e / kk, j = e - t * kk;
      gp[(size_t)t * n + j] = ldg_cg(W + (size_t)j * n + s_cand[t]);
    }
    double c[BQ];
#pragma unroll
    for (int u = 0; u < BQ; ++u) {
      const int t = q * BQ + u;
      c[u] = 0.0;
      if (act && t < nb) {
        const int pt = s_cand[t];
        c[u] = 0.5 * (ldg_cg(A + (size_t)i * n + pt) + ldg_cg(A + (size_t)pt * n + i));
      }
    }
    __syncthreads();
    // 3. one pass over the rows already in W for all columns of the block
    if (act) {
      const double* gq = gp + (size_t)q * BQ * n;
#pragma unroll 16
      for (int j = 0; j < kk; ++j) {
        const double w = ldg_cg(W + (size_t)j * n + i);
#pragma unroll
        for (int u = 0; u < BQ; ++u) c[u] = fma(-w, gq[(size_t)u * n + j], c[u]);
      }
    }
    // 4. factorise the block
    for (int t = 0; t < nb; ++t) {
      const int pt = s_cand[t], tq = t / BQ, tu = t - tq * BQ;
      double ct = 0.0;
#pragma unroll
      for (int u = 0; u < BQ; ++u) ct = (u == tu) ? c[u] : ct;
      if (act && q == tq && i == pt) {
        s_pivot[t] = ct;
        s_accept[t] = (ct > thresh && ct >= 0.25 * s_d0[t]) ? 1 : 0;
      }
      __syncthreads();
      if (!s_accept[t]) continue;   // uniform; the index stays available for a later block
      const double root = sqrt(s_pivot[t]), rinv = 1.0 / root;
      if (act && q == tq) {
        double g = done[i] ? 0.0 : ct * rinv;
        if (i == pt) g = root;
        W[(size_t)kk * n + i] = g;
        gcol[i] = g;
        diag[i] -= g * g;
      }
      __syncthreads();
      if (act) {
        const double gi = gcol[i];
#pragma unroll
        for (int u = 0; u < BQ; ++u) {
          const int tt = q * BQ + u;
          if (tt > t && tt < nb) c[u] = fma(-gi, gcol[s_cand[tt]], c[u]);
        }
      }
      if (tid == 0) done[pt] = 1;
      ++kk;
      __syncthreads();
    }
  }
  for (size_t idx = (size_t)kk * n + tid; idx < (size_t)n * n; idx += T) W[idx] = 0.0;
  __syncthreads();
  return kk;
}

// order[rk] = index of the rk-th largest |vals| (stable); one value per thread over the grid, then a grid barrier
__device__ void rank_grid(TgBar& grid, const double* vals, int n, int* order) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const double aj = fabs(ldg_cg(vals + j));
    int rk = 0;
#pragma unroll 8
    for (int i = 0; i < n; ++i) {
      const double ai = fabs(ldg_cg(vals + i));
      rk += (ai > aj) || (ai == aj && i < j);
    }
    order[rk] = j;
  }
  grid.sync();
}

template <int NPL>
__global__ void __launch_bounds__(kTgThreads, 1) tica_solve_grid_kernel(
    const double* __restrict__ C00, const double* __restrict__ C0t, int d, double eps,
    double* __restrict__ evals, double* __restrict__ evecs, int32_t* __restrict__ rank_out, TicaGridWs ws,
    int blk, int cluster_mode, int use_chol) {
  extern __shared__ __align__(16) double tg_sm[];
  TgBar grid{cluster_mode != 0};
  const long long t_begin = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int q = 0; q < 8; ++q) g_tg_dbg[q] = 0;
    for (int q = 16; q < 80; ++q) g_tg_trace[q] = 0;
  }
  if (blockIdx.x < 16 && threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g_tg_trace[blockIdx.x] = smid;
  }
  const long long gt_begin = tg_globaltimer();
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gnt = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31, gwarp = gtid >> 5, nwarps = gnt >> 5;
  const size_t dd = (size_t)d * d;

  // 1. eig(C00)
  for (size_t i = gtid; i < dd; i += gnt) {
    const int r = (int)(i / d), c = (int)(i - (size_t)r * d);
    const double a = 0.5 * (C00[i] + C00[(size_t)c * d + r]);
    ws.W[i] = a;
    if (use_chol) ws.TMP[i] = a;
  }
  grid.sync();
  // Rotations are NOT accumulated: at convergence row j of W is lambda_j v_j^T, so v_j = w_j / |w_j| and the
  // signed eigenvalue is the Rayleigh quotient v_j^T C00 v_j.  (Rows whose norm is at rounding level carry no
  // direction; they get eigenvalue 0 and are removed by the rank cut.)  This halves the shared-memory
  // traffic that bounds a local Jacobi round.
  if (use_chol) {
    // W <- G^T of the pivoted Cholesky factorisation of sym(C00) (one CTA; see chol_pivoted_cta).  ws.W holds
    // sym(C00) at this point and is the output: the input is read from the TMP copy made above.
    if (blockIdx.x == 0) {
      if (d <= kTgThreads / 2) chol_pivoted_cta<2>(ws.TMP, d, ws.W, tg_sm);
      else chol_pivoted_cta<1>(ws.TMP, d, ws.W, tg_sm);
    }
    grid.sync();
  }
  const int sweeps1 = jacobi_grid<NPL>(grid, ws.W, d, blk, ws.ctrl + 1, tg_sm);
  for (int j = gwarp; j < d; j += nwarps) {
    double acc = 0.0;
    for (int e = lane; e < d; e += 32) { const double w = ldg_cg(ws.W + (size_t)j * d + e); acc = fma(w, w, acc); }
    acc = warp_sum(acc);
    if (lane == 0) ws.s[j] = sqrt(acc);
  }
  grid.sync();
  {
    double mx = 0.0;
    for (int j = lane; j < d; j += 32) mx = fmax(mx, ldg_cg(ws.s + j));
    mx = warp_max(mx);
    const double floor_norm = 16.0 * 2.220446049250313e-16 * (double)d * mx;
    grid.sync();   // every warp has read the norms before they are replaced by eigenvalues
    for (int j = gwarp; j < d; j += nwarps) {
      double nrm2 = 0.0;
      double v[NPL];
#pragma unroll
      for (int q = 0; q < NPL; ++q) {
        const int e = lane + 32 * q;
        v[q] = e < d ? ldg_cg(ws.W + (size_t)j * d + e) : 0.0;
        nrm2 = fma(v[q], v[q], nrm2);
      }
      const double nrm = sqrt(warp_sum(nrm2));
      const bool live = nrm > floor_norm && nrm > 0.0;
      const double inv = live ? 1.0 / nrm : 0.0;
#pragma unroll
      for (int q = 0; q < NPL; ++q) {
        const int e = lane + 32 * q;
        v[q] *= inv;
        if (e < d) ws.V[(size_t)j * d + e] = v[q];
      }
      double acc = 0.0;
      if (live) {
#pragma unroll
        for (int q2 = 0; q2 < NPL; ++q2) {
          for (int l = 0; l < 32; ++l) {
            const int i = l + 32 * q2;
            const double vi = __shfl_sync(0xffffffffu, v[q2], l);
            if (i >= d) continue;
            const double* row = C00 + (size_t)i * d;
            double part = 0.0;
#pragma unroll
            for (int q = 0; q < NPL; ++q) {
              const int e = lane + 32 * q;
              if (e < d) part = fma(row[e], v[q], part);
            }
            acc = fma(vi, part, acc);
          }
        }
      }
      acc = warp_sum(acc);
      if (lane == 0) ws.s[j] = acc;
    }
  }
  grid.sync();
  rank_grid(grid, ws.s, d, ws.order);
  if (gtid == 0) {
    double evmin = ldg_cg(ws.s);
    for (int j = 1; j < d; ++j) evmin = fmin(evmin, ldg_cg(ws.s + j));
    double e = eps;
    if (evmin < 0.0) e = fmax(e, -evmin + 1e-16);
    int m = 0;
    for (int k = 0; k < d; ++k)
      if (fabs(ldg_cg(ws.s + __ldcg(ws.order + k))) >= e) ++m; else break;
    ws.ctrl[0] = m;
  }
  grid.sync();
  const int m = __ldcg(ws.ctrl);
  if (m == 0) {
    if (gtid == 0) { rank_out[0] = 0; rank_out[1] = sweeps1; rank_out[2] = 0; }
    for (size_t i = gtid; i < dd; i += gnt) evecs[i] = 0.0;
    for (int j = gtid; j < d; j += gnt) evals[j] = 0.0;
    return;
  }
  // 2. L[:,k] = sign * v_k / sqrt(s_k)  (d x m row-major)
  for (int k = gwarp; k < m; k += nwarps) {
    const int j = __ldcg(ws.order + k);
    const double* v = ws.V + (size_t)j * d;
    double best = -1.0, bval = 0.0;
    int bidx = 0x7fffffff;
    for (int e = lane; e < d; e += 32) {
      const double ve = ldg_cg(v + e), av = fabs(ve);
      if (av > best) { best = av; bval = ve; bidx = e; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const double ov = __shfl_xor_sync(0xffffffffu, bval, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ob > best || (ob == best && oi < bidx)) { best = ob; bval = ov; bidx = oi; }
    }
    const double sc = ((bval < 0.0) ? -1.0 : 1.0) / sqrt(ldg_cg(ws.s + j));
    for (int e = lane; e < d; e += 32) ws.L[(size_t)e * m + k] = ldg_cg(v + e) * sc;
  }
  // W <- sym(C0t)
  for (size_t i = gtid; i < dd; i += gnt) {
    const int r = (int)(i / d), c = (int)(i - (size_t)r * d);
    ws.W[i] = 0.5 * (C0t[i] + C0t[(size_t)c * d + r]);
  }
  grid.sync();
  // 3. TMP = C0t_sym L (d x m), M = L^T TMP (m x m)
  for (size_t i = gtid; i < (size_t)d * m; i += gnt) {
    const int r = (int)(i / m), c = (int)(i - (size_t)r * m);
    double acc = 0.0;
    for (int e = 0; e < d; ++e) acc = fma(ldg_cg(ws.W + (size_t)r * d + e), ldg_cg(ws.L + (size_t)e * m + c), acc);
    ws.TMP[i] = acc;
  }
  grid.sync();
  for (size_t i = gtid; i < (size_t)m * m; i += gnt) {
    const int r = (int)(i / m), c = (int)(i - (size_t)r * m);
    double acc = 0.0;
    for (int e = 0; e < d; ++e) acc = fma(ldg_cg(ws.L + (size_t)e * m + r), ldg_cg(ws.TMP + (size_t)e * m + c), acc);
    ws.M[i] = acc;
  }
  grid.sync();
  // sigma = 1.0625 * ||sym(M)||_inf, computed by every thread's CTA redundantly via one warp-sum each
  double rowmax = 0.0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    double s = 0.0;
    for (int j = 0; j < m; ++j) s += fabs(0.5 * (ldg_cg(ws.M + (size_t)i * m + j) + ldg_cg(ws.M + (size_t)j * m + i)));
    rowmax = fmax(rowmax, s);
  }
  __shared__ double s_red[kTgThreads / 32];
  rowmax = warp_max(rowmax);
  if (lane == 0) s_red[threadIdx.x >> 5] = rowmax;
  __syncthreads();
  double nrm = 0.0;
  for (int w = 0; w < kTgThreads / 32; ++w) nrm = fmax(nrm, s_red[w]);
  const double sigma = 1.0625 * nrm + 1e-300;
  grid.sync();   // everyone has read M before W/V are overwritten
  for (size_t i = gtid; i < (size_t)m * m; i += gnt) {
    const int r = (int)(i / m), c = (int)(i - (size_t)r * m);
    const double a = 0.5 * (ldg_cg(ws.M + i) + ldg_cg(ws.M + (size_t)c * m + r)) + ((r == c) ? sigma : 0.0);
    ws.W[i] = a;
  }
  grid.sync();
  // (the Cholesky preconditioner does not pay here: sym(M) + sigma I has condition ~30, 12 sweeps either way)
  // W = sym(M) + sigma I is positive definite: eigenvalue_j = |w_j| - sigma, v_j = w_j / |w_j|
  const int sweeps2 = jacobi_grid<NPL>(grid, ws.W, m, blk, ws.ctrl + 1, tg_sm);
  for (int j = gwarp; j < m; j += nwarps) {
    double acc = 0.0;
    for (int e = lane; e < m; e += 32) { const double w = ldg_cg(ws.W + (size_t)j * m + e); acc = fma(w, w, acc); }
    const double nrm = sqrt(warp_sum(acc));
    const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
    for (int e = lane; e < m; e += 32) ws.V[(size_t)j * m + e] = ldg_cg(ws.W + (size_t)j * m + e) * inv;
    if (lane == 0) ws.s[j] = nrm - sigma;
  }
  grid.sync();
  rank_grid(grid, ws.s, m, ws.order);
  // 4. R[:,k] = L V2[:,k] with canonical signs; evecs is d x d row-major
  for (size_t i = gtid; i < dd; i += gnt) evecs[i] = 0.0;
  grid.sync();
  for (size_t i = gtid; i < (size_t)d * m; i += gnt) {
    const int r = (int)(i / m), k = (int)(i - (size_t)r * m);
    const double* v2 = ws.V + (size_t)__ldcg(ws.order + k) * m;
    double acc = 0.0;
    for (int e = 0; e < m; ++e) acc = fma(ldg_cg(ws.L + (size_t)r * m + e), ldg_cg(v2 + e), acc);
    evecs[(size_t)r * d + k] = acc;
  }
  grid.sync();
  for (int k = gwarp; k < m; k += nwarps) {
    double best = -1.0, bval = 0.0;
    int bidx = 0x7fffffff;
    for (int r = lane; r < d; r += 32) {
      const double v = ldg_cg(evecs + (size_t)r * d + k), av = fabs(v);
      if (av > best) { best = av; bval = v; bidx = r; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const double ov = __shfl_xor_sync(0xffffffffu, bval, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ob > best || (ob == best && oi < bidx)) { best = ob; bval = ov; bidx = oi; }
    }
    if (bval < 0.0)
      for (int r = lane; r < d; r += 32) evecs[(size_t)r * d + k] = -ldg_cg(evecs + (size_t)r * d + k);
    if (lane == 0) evals[k] = ldg_cg(ws.s + __ldcg(ws.order + k));
  }
  for (int k = m + gtid; k < d; k += gnt) evals[k] = 0.0;
  if (gtid == 0) { rank_out[0] = m; rank_out[1] = sweeps1; rank_out[2] = sweeps2; g_tg_dbg[5] = clock64() - t_begin; g_tg_dbg[6] = tg_globaltimer() - gt_begin; }
}

int tica_grid_debug_trace(int64_t* out80) {
  long long h[80];
  PMB_CUDA(cudaMemcpyFromSymbol(h, g_tg_trace, sizeof(h)));
  for (int i = 0; i < 80; ++i) out80[i] = h[i];
  return PMB_OK;
}

int tica_grid_debug_counters(int64_t* out8) {
  long long h[8];
  PMB_CUDA(cudaMemcpyFromSymbol(h, g_tg_dbg, sizeof(h)));
  for (int i = 0; i < 8; ++i) out8[i] = h[i];
  return PMB_OK;
}

int tica_solve_grid_launch(const double* C00, const double* C0t, int d, double eps, double* evals, double* evecs,
                           int32_t* rank, void* ws, cudaStream_t st) {
  TicaGridWs w;
  double* base = static_cast<double*>(ws);
  const size_t dd = (size_t)d * d;
  w.W = base; w.V = base + dd; w.L = base + 2 * dd; w.TMP = base + 3 * dd; w.M = base + 4 * dd;
  w.s = base + 5 * dd;
  w.order = reinterpret_cast<int*>(base + 5 * dd + d);
  w.ctrl = w.order + d;
  int dev = 0, sms = 0, per_sm = 0;
  PMB_CUDA(cudaGetDevice(&dev));
  PMB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // rows per block: the 2 b rows of a block pair (d doubles each) must fit in shared memory
  // b = 8: d / 16 CTAs own a block pair.  Measured on B200 for d = 256 (cycles of the whole solve):
  // b = 16 -> 16.8 M, b = 8 -> 15.6 M (half the warps per SM contend for the fp64 pipe, twice the grid
  // barriers), b = 4 -> 19.7 M.
  int blk = 8;
  while (blk > 1 && (size_t)2 * blk * d * sizeof(double) > (size_t)200 * 1024) blk >>= 1;
  const size_t smem = (size_t)2 * blk * (d + 1) * sizeof(double);   // rows + their squared norms
  PMB_REQUIRE(smem <= (size_t)227 * 1024, "pmb_tica_solve: d=%d too large for the block-Jacobi kernel", d);
  void* kern = nullptr;
  if (d <= 64) kern = (void*)tica_solve_grid_kernel<2>;
  else if (d <= 128) kern = (void*)tica_solve_grid_kernel<4>;
  else if (d <= 256) kern = (void*)tica_solve_grid_kernel<8>;
  else if (d <= 512) kern = (void*)tica_solve_grid_kernel<16>;
  else if (d <= 1024) kern = (void*)tica_solve_grid_kernel<32>;
  else {
    set_error("pmb_tica_solve: d=%d > 1024 is not supported by the block-Jacobi kernel", d);
    return PMB_EUNSUPPORTED;
  }
  PMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTgThreads, smem));
  PMB_REQUIRE(per_sm >= 1, "pmb_tica_solve: kernel does not fit on an SM");
  int nb = (d + blk - 1) / blk;
  if (nb & 1) ++nb;
  // One CTA per SM even though only nb / 2 of them own a block pair: the others join the grid barriers and
  // the dense phases.  (The in-kernel trace -- clock64 against %globaltimer per sweep, pmb_debug_trace_tica --
  // shows the kernel at the full SM clock in every run; stage-time outliers seen in round 1 were host-side
  // allocator stalls in front of the launch, not the kernel.)
  int cluster_mode = 0;
  static const int chol_env = [] { const char* e = getenv("PMB_TICA_CHOL"); return e ? atoi(e) : 1; }();
  int use_chol = (chol_env && d <= kTgThreads) ? 1 : 0;
  void* args[] = {(void*)&C00, (void*)&C0t, (void*)&d, (void*)&eps, (void*)&evals, (void*)&evecs, (void*)&rank, (void*)&w,
                  (void*)&blk, (void*)&cluster_mode, (void*)&use_chol};
  // Preferred: one cluster of 16 CTAs (non-portable size), hardware cluster barriers.
  static const int allow_cluster = [] { const char* e = getenv("PMB_TICA_CLUSTER"); return e ? atoi(e) : 1; }();
  if (allow_cluster && nb / 2 <= 16 && d <= 128) {   // d = 256: the dense phases on 16 SMs cost what the barriers save (8.3 vs 7.9 ms in the bench)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(16);
    cfg.blockDim = dim3(kTgThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 16;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n_clusters = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
        cudaOccupancyMaxActiveClusters(&n_clusters, kern, &cfg) == cudaSuccess && n_clusters >= 1) {
      cluster_mode = 1;
      PMB_CUDA(cudaLaunchKernelExC(&cfg, kern, args));
      count_launch();
      return PMB_OK;
    }
    (void)cudaGetLastError();   // no room for a 16-CTA cluster: cooperative launch below
  }
  // One CTA per SM even though only nb / 2 of them own a block pair: the others join the grid barriers and
  // the dense phases.
  int grid = sms;
  if (grid < nb / 2) grid = nb / 2;
  if (grid > sms * per_sm) grid = sms * per_sm;
  PMB_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kTgThreads), args, smem, st));
  count_launch();
  return PMB_OK;
}

}  // namespace pmb
