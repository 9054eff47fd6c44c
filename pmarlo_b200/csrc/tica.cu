// tica.cu -- K4: scaler statistics, covariance assembly and the generalized
// symmetric eigenproblem of TICA, all in fp64 on d x d data (latency-bound,
// not reported against a roofline).
//
// Eigen-solver: one-sided (Hestenes) cyclic Jacobi, one CTA per matrix.  The
// matrix columns live as ROWS of W (A is symmetric), V accumulates the
// rotations, so every access is a contiguous row; one warp owns one pair per
// round of a round-robin tournament ordering, rounds are separated by
// __syncthreads().  Indefinite matrices are shifted by their inf-norm so that
// singular values and eigenvalues coincide; eigenvalues come back as Rayleigh
// quotients v.w - sigma.
//
// Conventions follow deeptime's spd_inv_split / eig_corr (oracle/tica.py):
// sort by magnitude descending, keep |s| >= epsilon (epsilon raised to
// -min(s)+1e-16 if C00 has negative eigenvalues), canonical signs.
#include "common.cuh"

namespace pmb {

constexpr int kJacThreads = 1024;
constexpr int kJacMaxSweeps = 40;
constexpr double kJacTol = 4.5e-16;  // x sqrt(n): rounding noise of an n-term dot product (cf. LAPACK dgesvj)

// ------------------------------------------------------------ Jacobi (device)
// W: n x n (rows = vectors), V: n x n or nullptr.  Returns number of sweeps.
__device__ int jacobi_onesided(double* __restrict__ W, double* __restrict__ V, int n, int* s_flag) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int np = (n + 1) & ~1;  // players (even)
  const int half = np >> 1;
  int sweep = 0;
  for (; sweep < kJacMaxSweeps; ++sweep) {
    if (tid == 0) *s_flag = 0;
    __syncthreads();
    int rotated = 0;
    for (int r = 0; r < np - 1; ++r) {
      for (int k = warp; k < half; k += nwarps) {
        int p, q;
        if (k == 0) {
          p = np - 1;
          q = r;
        } else {
          p = (r + k) % (np - 1);
          q = (r - k + (np - 1)) % (np - 1);
        }
        if (p >= n || q >= n) continue;
        if (p > q) { int t = p; p = q; q = t; }
        double* wp = W + (size_t)p * n;
        double* wq = W + (size_t)q * n;
        double a = 0.0, b = 0.0, g = 0.0;
        for (int e = lane; e < n; e += 32) {
          const double x = wp[e], y = wq[e];
          a = fma(x, x, a);
          b = fma(y, y, b);
          g = fma(x, y, g);
        }
        a = warp_sum(a);
        b = warp_sum(b);
        g = warp_sum(g);
        if (fabs(g) <= kJacTol * sqrt((double)n) * sqrt(a * b) || g == 0.0) continue;
        rotated = 1;
        const double zeta = (b - a) / (2.0 * g);
        const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int e = lane; e < n; e += 32) {
          const double x = wp[e], y = wq[e];
          wp[e] = c * x - s * y;
          wq[e] = s * x + c * y;
        }
        if (V) {
          double* vp = V + (size_t)p * n;
          double* vq = V + (size_t)q * n;
          for (int e = lane; e < n; e += 32) {
            const double x = vp[e], y = vq[e];
            vp[e] = c * x - s * y;
            vq[e] = s * x + c * y;
          }
        }
      }
      __syncthreads();
    }
    if (rotated && lane == 0) atomicOr(s_flag, 1);
    __syncthreads();
    const int any = *s_flag;
    __syncthreads();
    if (!any) { ++sweep; break; }
  }
  return sweep;
}

// rank[j] = position of j when sorting |vals| descending (stable)
__device__ void rank_by_magnitude(const double* vals, int n, int* order) {
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const double aj = fabs(vals[j]);
    int rk = 0;
    for (int i = 0; i < n; ++i) {
      const double ai = fabs(vals[i]);
      rk += (ai > aj) || (ai == aj && i < j);
    }
    order[rk] = j;
  }
  __syncthreads();
}

__device__ double inf_norm(const double* A, int n, double* s_red) {
  double m = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double s = 0.0;
    for (int j = 0; j < n; ++j) s += fabs(A[(size_t)i * n + j]);
    m = fmax(m, s);
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r = fmax(r, s_red[w]);
    s_red[0] = r;
  }
  __syncthreads();
  const double r = s_red[0];
  __syncthreads();
  return r;
}

// ------------------------------------------------------------ tica_solve
struct TicaWs {
  double *W, *V, *L, *TMP, *M, *s;
  int* order;
};

__global__ void __launch_bounds__(kJacThreads) tica_solve_kernel(
    const double* __restrict__ C00, const double* __restrict__ C0t, int d, double eps,
    double* __restrict__ evals, double* __restrict__ evecs, int32_t* __restrict__ rank_out, TicaWs ws) {
  __shared__ int s_flag;
  __shared__ int s_m;
  __shared__ double s_red[32];
  const int tid = threadIdx.x, nt = blockDim.x;
  const size_t dd = (size_t)d * d;

  // 1. eig(C00)
  for (size_t i = tid; i < dd; i += nt) {
    const int r = (int)(i / d), c = (int)(i - (size_t)r * d);
    ws.W[i] = 0.5 * (C00[i] + C00[(size_t)c * d + r]);
    ws.V[i] = (r == c) ? 1.0 : 0.0;
  }
  __syncthreads();
  const int sweeps1 = jacobi_onesided(ws.W, ws.V, d, &s_flag);
  for (int j = tid; j < d; j += nt) {
    double acc = 0.0;
    for (int e = 0; e < d; ++e) acc = fma(ws.V[(size_t)j * d + e], ws.W[(size_t)j * d + e], acc);
    ws.s[j] = acc;
  }
  __syncthreads();
  rank_by_magnitude(ws.s, d, ws.order);
  if (tid == 0) {
    double evmin = ws.s[0];
    for (int j = 1; j < d; ++j) evmin = fmin(evmin, ws.s[j]);
    double e = eps;
    if (evmin < 0.0) e = fmax(e, -evmin + 1e-16);
    int m = 0;
    for (int k = 0; k < d; ++k)
      if (fabs(ws.s[ws.order[k]]) >= e) ++m; else break;
    s_m = m;
  }
  __syncthreads();
  const int m = s_m;
  if (m == 0) {
    if (tid == 0) { rank_out[0] = 0; rank_out[1] = sweeps1; rank_out[2] = 0; }
    for (size_t i = tid; i < dd; i += nt) evecs[i] = 0.0;
    for (int j = tid; j < d; j += nt) evals[j] = 0.0;
    return;
  }
  // 2. L[:,k] = sign * v_k / sqrt(s_k)   (L stored d x m, row-major)
  for (int k = tid >> 5; k < m; k += nt >> 5) {
    const int lane = tid & 31;
    const int j = ws.order[k];
    const double* v = ws.V + (size_t)j * d;
    double best = -1.0, bval = 0.0;
    int bidx = 0x7fffffff;
    for (int e = lane; e < d; e += 32) {
      const double av = fabs(v[e]);
      if (av > best) { best = av; bval = v[e]; bidx = e; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const double ov = __shfl_xor_sync(0xffffffffu, bval, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ob > best || (ob == best && oi < bidx)) { best = ob; bval = ov; bidx = oi; }
    }
    const double sc = ((bval < 0.0) ? -1.0 : 1.0) / sqrt(ws.s[j]);
    for (int e = lane; e < d; e += 32) ws.L[(size_t)e * m + k] = v[e] * sc;
  }
  __syncthreads();
  // 3. TMP = C0t_sym L (d x m);  M = L^T TMP (m x m)
  for (size_t i = tid; i < dd; i += nt) {  // W <- sym(C0t) (W is free here)
    const int r = (int)(i / d), c = (int)(i - (size_t)r * d);
    ws.W[i] = 0.5 * (C0t[i] + C0t[(size_t)c * d + r]);
  }
  __syncthreads();
  for (size_t i = tid; i < (size_t)d * m; i += nt) {
    const int r = (int)(i / m), c = (int)(i - (size_t)r * m);
    double acc = 0.0;
    for (int e = 0; e < d; ++e) acc = fma(ws.W[(size_t)r * d + e], ws.L[(size_t)e * m + c], acc);
    ws.TMP[i] = acc;
  }
  __syncthreads();
  for (size_t i = tid; i < (size_t)m * m; i += nt) {
    const int r = (int)(i / m), c = (int)(i - (size_t)r * m);
    double acc = 0.0;
    for (int e = 0; e < d; ++e) acc = fma(ws.L[(size_t)e * m + r], ws.TMP[(size_t)e * m + c], acc);
    ws.M[i] = acc;
  }
  __syncthreads();
  for (size_t i = tid; i < (size_t)m * m; i += nt) {
    const int r = (int)(i / m), c = (int)(i - (size_t)r * m);
    ws.W[i] = 0.5 * (ws.M[i] + ws.M[(size_t)c * m + r]);
    ws.V[i] = (r == c) ? 1.0 : 0.0;
  }
  __syncthreads();
  const double sigma = 1.0625 * inf_norm(ws.W, m, s_red) + 1e-300;
  for (int j = tid; j < m; j += nt) ws.W[(size_t)j * m + j] += sigma;
  __syncthreads();
  const int sweeps2 = jacobi_onesided(ws.W, ws.V, m, &s_flag);
  for (int j = tid; j < m; j += nt) {
    double acc = 0.0;
    for (int e = 0; e < m; ++e) acc = fma(ws.V[(size_t)j * m + e], ws.W[(size_t)j * m + e], acc);
    ws.s[j] = acc - sigma;
  }
  __syncthreads();
  rank_by_magnitude(ws.s, m, ws.order);
  // 4. R[:,k] = L V2[:,k], canonical signs; evecs is d x d row-major
  for (size_t i = tid; i < dd; i += nt) evecs[i] = 0.0;
  __syncthreads();
  for (size_t i = tid; i < (size_t)d * m; i += nt) {
    const int r = (int)(i / m), k = (int)(i - (size_t)r * m);
    const double* v2 = ws.V + (size_t)ws.order[k] * m;
    double acc = 0.0;
    for (int e = 0; e < m; ++e) acc = fma(ws.L[(size_t)r * m + e], v2[e], acc);
    evecs[(size_t)r * d + k] = acc;
  }
  __syncthreads();
  for (int k = tid; k < m; k += nt) {
    double best = -1.0, bval = 0.0;
    for (int r = 0; r < d; ++r) {
      const double v = evecs[(size_t)r * d + k];
      if (fabs(v) > best) { best = fabs(v); bval = v; }
    }
    if (bval < 0.0)
      for (int r = 0; r < d; ++r) evecs[(size_t)r * d + k] = -evecs[(size_t)r * d + k];
    evals[k] = ws.s[ws.order[k]];
  }
  for (int k = m + tid; k < d; k += nt) evals[k] = 0.0;
  if (tid == 0) { rank_out[0] = m; rank_out[1] = sweeps1; rank_out[2] = sweeps2; }
}

// ------------------------------------------------------------ batched eigenvalues
__global__ void __launch_bounds__(kJacThreads) sym_eigvals_kernel(double* __restrict__ A, int n,
                                                                   double* __restrict__ evals,
                                                                   double* __restrict__ scratch,
                                                                   int* __restrict__ order_ws) {
  __shared__ int s_flag;
  __shared__ double s_red[32];
  double* W = A + (size_t)blockIdx.x * n * n;
  double* s = scratch + (size_t)blockIdx.x * n;
  int* order = order_ws + (size_t)blockIdx.x * n;
  const int tid = threadIdx.x, nt = blockDim.x;
  // symmetrise in place (upper <- average), then mirror
  for (size_t i = tid; i < (size_t)n * n; i += nt) {
    const int r = (int)(i / n), c = (int)(i - (size_t)r * n);
    if (r < c) {
      const double v = 0.5 * (W[i] + W[(size_t)c * n + r]);
      W[i] = v;
      W[(size_t)c * n + r] = v;
    }
  }
  __syncthreads();
  const double sigma = 1.0625 * inf_norm(W, n, s_red) + 1e-300;
  for (int j = tid; j < n; j += nt) W[(size_t)j * n + j] += sigma;
  __syncthreads();
  jacobi_onesided(W, nullptr, n, &s_flag);
  for (int j = tid >> 5; j < n; j += nt >> 5) {
    double acc = 0.0;
    for (int e = tid & 31; e < n; e += 32) acc = fma(W[(size_t)j * n + e], W[(size_t)j * n + e], acc);
    acc = warp_sum(acc);
    if ((tid & 31) == 0) s[j] = sqrt(acc) - sigma;
  }
  __syncthreads();
  rank_by_magnitude(s, n, order);
  for (int k = tid; k < n; k += nt) evals[(size_t)blockIdx.x * n + k] = s[order[k]];
}

// ------------------------------------------------------------ scaler / covariances
__global__ void scaler_kernel(const double* __restrict__ mom, int64_t n, int d, int semantic,
                              int with_std, double* __restrict__ stats, float* __restrict__ cond) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  const double nv = mom[0 * d + c], sh = mom[1 * d + c], s1 = mom[2 * d + c], s2 = mom[3 * d + c];
  const double m = nv > 0 ? sh + s1 / nv : 0.0;
  double ss = nv > 0 ? s2 - s1 * s1 / nv : 0.0;
  if (ss < 0) ss = 0;
  const double var = ss / (double)n;
  const double eps = 2.220446049250313e-16;
  const double upper = (double)n * eps * var + ((double)n * m * eps) * ((double)n * m * eps);
  double sd = sqrt(var);
  const bool constant = (var <= upper) || (sd < 10 * eps);
  const double sd_safe = constant ? 1.0 : sd;
  stats[0 * d + c] = semantic ? m : 0.0;
  stats[1 * d + c] = (semantic && with_std) ? sd_safe : 1.0;
  stats[2 * d + c] = n > 1 ? sqrt(ss / (double)(n - 1)) : 0.0;
  cond[0 * d + c] = (float)m;
  cond[1 * d + c] = (float)(1.0 / sd_safe);
}

__global__ void tica_cov_kernel(const double* __restrict__ G0, const double* __restrict__ G1,
                                const double* __restrict__ mom, const double* __restrict__ stats,
                                const float* __restrict__ cond, int64_t n, int64_t n_pairs, int d,
                                double* __restrict__ C00, double* __restrict__ C0t,
                                double* __restrict__ mu_out) {
  // per-column quantities (recomputed per thread; d is small)
  auto col = [&](int j, double& alpha, double& beta, double& mu) {
    const double nv = mom[0 * d + j], sh = mom[1 * d + j], s1 = mom[2 * d + j];
    const double es1 = mom[4 * d + j], ecnt = mom[5 * d + j];
    const double m_imp = nv > 0 ? sh + s1 / nv : 0.0;  // imputation mean
    const double m = stats[0 * d + j], sc = stats[1 * d + j];
    const double sh32 = (double)cond[0 * d + j], sc32 = (double)cond[1 * d + j];
    const double twoT = 2.0 * (double)n_pairs;
    const double e_total = 2.0 * (double)n - twoT;
    // sum_g e_g x_imp  and  sum_g w_g x_imp
    const double E = sh * ecnt + es1 + m_imp * (e_total - ecnt);
    const double Sw = 2.0 * (double)n * m_imp - E;
    alpha = 1.0 / (sc32 * sc);
    beta = (sh32 - m) / sc;
    mu = (Sw / twoT - m) / sc;
  };
  const int i = blockIdx.y * blockDim.y + threadIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d || j >= d) return;
  double ai, bi, mi, aj, bj, mj;
  col(i, ai, bi, mi);
  col(j, aj, bj, mj);
  const double twoT = 2.0 * (double)n_pairs;
  const double Si = twoT * (mi - bi) / ai, Sj = twoT * (mj - bj) / aj;  // sum_g w z~
  const size_t ij = (size_t)i * d + j;
  const double g0z = ai * aj * G0[ij] + ai * bj * Si + bi * aj * Sj + bi * bj * twoT;
  const double g1z = ai * aj * G1[ij];
  C00[ij] = g0z / twoT - mi * mj;
  C0t[ij] = (g0z - g1z) / twoT - mi * mj;
  if (i == 0) mu_out[j] = mj;
}

// projection operands: y = (impute(x) - a) W  with  a = m + sc*mu,  W = diag(1/sc) R[:, :m] (x lambda)
__global__ void tica_finalize_kernel(const double* __restrict__ evals, const double* __restrict__ evecs,
                                     const double* __restrict__ mom, const double* __restrict__ stats,
                                     const double* __restrict__ mu, int d, int m, int kinetic_map,
                                     double* __restrict__ a, double* __restrict__ nanfill,
                                     double* __restrict__ W) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= d) return;
  const double mj = stats[0 * d + j], sc = stats[1 * d + j];
  const double nv = mom[0 * d + j], sh = mom[1 * d + j], s1 = mom[2 * d + j];
  a[j] = mj + sc * mu[j];
  nanfill[j] = nv > 0 ? sh + s1 / nv : 0.0;
  for (int c = 0; c < m; ++c) {
    const double lam = kinetic_map ? evals[c < d ? c : 0] : 1.0;
    W[(size_t)j * m + c] = (c < d) ? evecs[(size_t)j * d + c] * lam / sc : 0.0;
  }
}

int tica_solve_grid_launch(const double* C00, const double* C0t, int d, double eps, double* evals, double* evecs,
                           int32_t* rank, void* ws, cudaStream_t st);  // tica_grid.cu

int sym_eigvals_launch(double* A, int n, int batch, double* evals, double* scratch, int* order,
                       cudaStream_t st) {
  sym_eigvals_kernel<<<batch, kJacThreads, 0, st>>>(A, n, evals, scratch, order);
  PMB_LAUNCH_CHECK();
  return PMB_OK;
}

}  // namespace pmb

namespace pmb { int tica_grid_debug_counters(int64_t* out8); }
extern "C" int pmb_debug_counters_tica(int64_t* out8) {
  using namespace pmb;
  PMB_REQUIRE(out8 != nullptr, "pmb_debug_counters_tica: null pointer");
  return tica_grid_debug_counters(out8);
}

namespace pmb { int tica_grid_debug_trace(int64_t* out80); }
extern "C" int pmb_debug_trace_tica(int64_t* out80) {
  using namespace pmb;
  PMB_REQUIRE(out80 != nullptr, "pmb_debug_trace_tica: null pointer");
  return tica_grid_debug_trace(out80);
}

extern "C" size_t pmb_tica_solve_ws_bytes(int d) {
  if (d <= 0) return 0;
  return ((size_t)5 * d * d + d) * sizeof(double) + (size_t)d * sizeof(int) + 64;
}

extern "C" int pmb_tica_solve(const double* C00, const double* C0t, int d, double eps, double* evals,
                              double* evecs, int32_t* rank, void* ws, size_t ws_bytes,
                              pmb_stream_t stream) {
  using namespace pmb;
  PMB_REQUIRE(d > 0 && d <= 1024, "pmb_tica_solve: bad d=%d (supported: 1..1024)", d);
  PMB_REQUIRE(C00 && C0t && evals && evecs && rank && ws, "pmb_tica_solve: null pointer");
  if (ws_bytes < pmb_tica_solve_ws_bytes(d)) {
    set_error("pmb_tica_solve: workspace too small");
    return PMB_EWORKSPACE;
  }
  TicaWs w;
  double* base = static_cast<double*>(ws);
  const size_t dd = (size_t)d * d;
  w.W = base; w.V = base + dd; w.L = base + 2 * dd; w.TMP = base + 3 * dd; w.M = base + 4 * dd;
  w.s = base + 5 * dd;
  w.order = reinterpret_cast<int*>(base + 5 * dd + d);
  if (d >= 48) return tica_solve_grid_launch(C00, C0t, d, eps, evals, evecs, rank, ws, as_stream(stream));
  tica_solve_kernel<<<1, kJacThreads, 0, as_stream(stream)>>>(C00, C0t, d, eps, evals, evecs, rank, w);
  PMB_LAUNCH_CHECK();
  return PMB_OK;
}

extern "C" int pmb_tica_finalize(const double* evals, const double* evecs, const double* moments,
                                 const double* stats, const double* mu, int d, int m, int kinetic_map,
                                 double* a, double* nanfill, double* W, pmb_stream_t stream) {
  using namespace pmb;
  PMB_REQUIRE(d > 0 && m > 0, "pmb_tica_finalize: bad sizes");
  PMB_REQUIRE(evals && evecs && moments && stats && mu && a && nanfill && W, "pmb_tica_finalize: null pointer");
  tica_finalize_kernel<<<(d + 127) / 128, 128, 0, as_stream(stream)>>>(evals, evecs, moments, stats, mu, d, m,
                                                                     kinetic_map, a, nanfill, W);
  PMB_LAUNCH_CHECK();
  return PMB_OK;
}

extern "C" size_t pmb_sym_eigvals_ws_bytes(int n, int batch) {
  if (n <= 0 || batch <= 0) return 0;
  return (size_t)n * batch * (sizeof(double) + sizeof(int)) + 64;
}

extern "C" int pmb_sym_eigvals_batched(double* A, int n, int batch, double* evals, void* ws,
                                       size_t ws_bytes, pmb_stream_t stream) {
  using namespace pmb;
  PMB_REQUIRE(n > 0 && batch > 0 && n <= 4096, "pmb_sym_eigvals_batched: bad sizes");
  PMB_REQUIRE(A && evals && ws, "pmb_sym_eigvals_batched: null pointer");
  if (ws_bytes < pmb_sym_eigvals_ws_bytes(n, batch)) {
    set_error("pmb_sym_eigvals_batched: workspace too small");
    return PMB_EWORKSPACE;
  }
  double* scratch = static_cast<double*>(ws);
  int* order = reinterpret_cast<int*>(scratch + (size_t)n * batch);
  return sym_eigvals_launch(A, n, batch, evals, scratch, order, as_stream(stream));
}

extern "C" int pmb_scaler_from_moments(const double* moments, int64_t n, int d, int semantic,
                                       int with_std, double* stats, float* cond,
                                       pmb_stream_t stream) {
  using namespace pmb;
  PMB_REQUIRE(n > 0 && d > 0, "pmb_scaler_from_moments: bad sizes");
  PMB_REQUIRE(moments && stats && cond, "pmb_scaler_from_moments: null pointer");
  scaler_kernel<<<(d + 127) / 128, 128, 0, as_stream(stream)>>>(moments, n, d, semantic, with_std, stats, cond);
  PMB_LAUNCH_CHECK();
  return PMB_OK;
}

extern "C" int pmb_tica_covariances(const double* G0, const double* G1, const double* moments,
                                    const double* stats, const float* cond, int64_t n,
                                    int64_t n_pairs, int d, int semantic, double* C00, double* C0t,
                                    double* mu, pmb_stream_t stream) {
  using namespace pmb;
  (void)semantic;
  PMB_REQUIRE(n > 0 && d > 0 && n_pairs > 0, "pmb_tica_covariances: bad sizes (n_pairs=%lld)",
              (long long)n_pairs);
  PMB_REQUIRE(G0 && G1 && moments && stats && cond && C00 && C0t && mu, "pmb_tica_covariances: null pointer");
  dim3 block(16, 16), grid((d + 15) / 16, (d + 15) / 16);
  tica_cov_kernel<<<grid, block, 0, as_stream(stream)>>>(G0, G1, moments, stats, cond, n, n_pairs, d, C00, C0t, mu);
  PMB_LAUNCH_CHECK();
  return PMB_OK;
}
