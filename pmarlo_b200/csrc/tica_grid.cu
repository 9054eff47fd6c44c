// tica_grid.cu -- K4 on many SMs: the same algorithm as tica_solve_kernel (tica.cu) as ONE
// cooperative kernel.  SIMT fp64 on B200 runs at ~1/16 of the fp32 rate, so a d = 256 Jacobi
// eigen-solve (3e9 fp64 operations over ~30 sweeps) is throughput-bound on a single SM
// (measured 171 ms); here every round of the round-robin ordering gives one column pair to each
// warp of the grid and the rounds are separated by grid barriers.  Data that other CTAs wrote is
// always read with ld.global.cg (L2), never through L1.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace pmb {

constexpr int kTgThreads = 128;
constexpr int kTgMaxSweeps = 40;
constexpr double kTgTol = 4.5e-16;  // x sqrt(n), as in tica.cu

struct TicaGridWs {
  double *W, *V, *L, *TMP, *M, *s;
  int* order;
  int* ctrl;  // [0] rank m, [1..2] sweep flags, [3] spare
};

__device__ __forceinline__ double ldg_cg(const double* p) { return __ldcg(p); }

// one-sided cyclic Jacobi on the rows of W (n x n, symmetric input), rotations accumulated in V
__device__ int jacobi_grid(cg::grid_group& grid, double* W, double* V, int n, int* flags) {
  const int lane = threadIdx.x & 31;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int np = (n + 1) & ~1, half = np >> 1;
  const double tol = kTgTol * sqrt((double)n);
  if (blockIdx.x == 0 && threadIdx.x == 0) { flags[0] = 0; flags[1] = 0; }
  grid.sync();
  int sweep = 0;
  for (; sweep < kTgMaxSweeps; ++sweep) {
    int rotated = 0;
    for (int r = 0; r < np - 1; ++r) {
      for (int k = gwarp; k < half; k += nwarps) {
        int p, q;
        if (k == 0) { p = np - 1; q = r; }
        else { p = (r + k) % (np - 1); q = (r - k + (np - 1)) % (np - 1); }
        if (p >= n || q >= n) continue;
        if (p > q) { const int t = p; p = q; q = t; }
        double* wp = W + (size_t)p * n;
        double* wq = W + (size_t)q * n;
        double a = 0.0, b = 0.0, g = 0.0;
        for (int e = lane; e < n; e += 32) {
          const double x = ldg_cg(wp + e), y = ldg_cg(wq + e);
          a = fma(x, x, a);
          b = fma(y, y, b);
          g = fma(x, y, g);
        }
        a = warp_sum(a);
        b = warp_sum(b);
        g = warp_sum(g);
        if (fabs(g) <= tol * sqrt(a * b) || g == 0.0) continue;
        rotated = 1;
        const double zeta = (b - a) / (2.0 * g);
        const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int e = lane; e < n; e += 32) {
          const double x = ldg_cg(wp + e), y = ldg_cg(wq + e);
          __stcg(wp + e, c * x - s * y);
          __stcg(wq + e, s * x + c * y);
        }
        if (V != nullptr) {
          double* vp = V + (size_t)p * n;
          double* vq = V + (size_t)q * n;
          for (int e = lane; e < n; e += 32) {
            const double x = ldg_cg(vp + e), y = ldg_cg(vq + e);
            __stcg(vp + e, c * x - s * y);
            __stcg(vq + e, s * x + c * y);
          }
        }
      }
      grid.sync();
    }
    if (rotated && lane == 0) atomicOr(&flags[sweep & 1], 1);
    grid.sync();
    const int any = __ldcg(&flags[sweep & 1]);
    if (blockIdx.x == 0 && threadIdx.x == 0) flags[(sweep + 1) & 1] = 0;
    grid.sync();
    if (!any) { ++sweep; break; }
  }
  return sweep;
}

// order[rk] = index of the rk-th largest |vals| (stable); CTA 0 only, followed by a grid barrier
__device__ void rank_grid(cg::grid_group& grid, const double* vals, int n, int* order) {
  if (blockIdx.x == 0) {
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      const double aj = fabs(ldg_cg(vals + j));
      int rk = 0;
      for (int i = 0; i < n; ++i) {
        const double ai = fabs(ldg_cg(vals + i));
        rk += (ai > aj) || (ai == aj && i < j);
      }
      order[rk] = j;
    }
  }
  grid.sync();
}

__global__ void __launch_bounds__(kTgThreads) tica_solve_grid_kernel(
    const double* __restrict__ C00, const double* __restrict__ C0t, int d, double eps,
    double* __restrict__ evals, double* __restrict__ evecs, int32_t* __restrict__ rank_out, TicaGridWs ws) {
  cg::grid_group grid = cg::this_grid();
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gnt = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31, gwarp = gtid >> 5, nwarps = gnt >> 5;
  const size_t dd = (size_t)d * d;

  // 1. eig(C00)
  for (size_t i = gtid; i < dd; i += gnt) {
    const int r = (int)(i / d), c = (int)(i - (size_t)r * d);
    ws.W[i] = 0.5 * (C00[i] + C00[(size_t)c * d + r]);
    ws.V[i] = (r == c) ? 1.0 : 0.0;
  }
  grid.sync();
  const int sweeps1 = jacobi_grid(grid, ws.W, ws.V, d, ws.ctrl + 1);
  for (int j = gwarp; j < d; j += nwarps) {   // Rayleigh quotients v_j . w_j
    double acc = 0.0;
    for (int e = lane; e < d; e += 32) acc = fma(ldg_cg(ws.V + (size_t)j * d + e), ldg_cg(ws.W + (size_t)j * d + e), acc);
    acc = warp_sum(acc);
    if (lane == 0) ws.s[j] = acc;
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
    ws.W[i] = 0.5 * (ldg_cg(ws.M + i) + ldg_cg(ws.M + (size_t)c * m + r)) + ((r == c) ? sigma : 0.0);
    ws.V[i] = (r == c) ? 1.0 : 0.0;
  }
  grid.sync();
  const int sweeps2 = jacobi_grid(grid, ws.W, ws.V, m, ws.ctrl + 1);
  for (int j = gwarp; j < m; j += nwarps) {
    double acc = 0.0;
    for (int e = lane; e < m; e += 32) acc = fma(ldg_cg(ws.V + (size_t)j * m + e), ldg_cg(ws.W + (size_t)j * m + e), acc);
    acc = warp_sum(acc);
    if (lane == 0) ws.s[j] = acc - sigma;
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
  if (gtid == 0) { rank_out[0] = m; rank_out[1] = sweeps1; rank_out[2] = sweeps2; }
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
  PMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tica_solve_grid_kernel, kTgThreads, 0));
  PMB_REQUIRE(per_sm >= 1, "pmb_tica_solve: kernel does not fit on an SM");
  const int half = ((d + 1) & ~1) / 2, wpc = kTgThreads / 32;
  int grid = (half + wpc - 1) / wpc;
  if (grid > sms) grid = sms;
  if (grid < 1) grid = 1;
  void* args[] = {(void*)&C00, (void*)&C0t, (void*)&d, (void*)&eps, (void*)&evals, (void*)&evecs, (void*)&rank, (void*)&w};
  PMB_CUDA(cudaLaunchCooperativeKernel((void*)tica_solve_grid_kernel, dim3(grid), dim3(kTgThreads), args, 0, st));
  count_launch();
  return PMB_OK;
}

}  // namespace pmb
