// bayes.cu -- Gibbs sampler for reversible transition matrices (Bayesian MSM confidence intervals).
//
// Replaces deeptime `BayesianMSM(lagtime, n_samples).fit(dtrajs)` -> `TransitionMatrixSampler(reversible=True)`
// (C++ `SamplerRev`), which `ITSMixin.compute_implied_timescales` runs for every lag
// (src/pmarlo/markov_state_model/_its.py:272-357).  Published algorithm: Trendelkamp-Schroer, Wu, Paul, Noe,
// "Estimation and uncertainty of reversible Markov models", J. Chem. Phys. 143, 174101 (2015).
//
// State of a chain: a symmetric non-negative matrix X (x_i = sum_j X_ij, T_ij = X_ij / x_i, pi_i ~ x_i), started
// at X = diag(pi) T of the reversible maximum-likelihood estimate.  With c0 = C_ij + C_ji, c_i = sum_j C_ij and
// the "-1" prior the full conditionals are
//   diagonal      : X_ii = t / (1 - t) (x_i - X_ii),  t ~ Beta(C_ii, c_i - C_ii)           (when both are > 0)
//   off-diagonal  : p(x) ~ x^(c0 - 1) (x + v1)^(-c_i) (x + v2)^(-c_j),  v1 = x_i - X_ij, v2 = x_j - X_ij,
//                   updated by two Metropolis steps: a Gamma(k, theta) proposal fitted to the mode and curvature of
//                   log p, then a log-normal random walk of unit step (target x^c0 ... in log space).
// Only elements with c0 > 0 are touched.  One sample = n_steps sweeps (deeptime default: sqrt(K)), then X is
// normalised and T = X / x is written out.
//
// Parallelisation: one CTA per chain (lag time).  Inside a sweep the K diagonal updates are mutually independent,
// and so are the off-diagonal updates of DISJOINT index pairs (the conditional of X_ij involves rows i and j only),
// so a sweep is 1 + (K - 1) rounds of a round-robin tournament with K / 2 simultaneous pair updates each -- the
// same Gibbs kernel as the sequential scan, in a different (still systematic) order.  X lives in global memory
// (L2-resident), the row sums in shared memory.  RNG: one xoshiro256** stream per thread, seeded by
// splitmix64(seed, chain, thread); normals by Box-Muller, gammas by Marsaglia-Tsang.
#include "common.cuh"

namespace pmb {

struct Rng {
  uint64_t s[4];
  __device__ static uint64_t splitmix(uint64_t& x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  __device__ void seed(uint64_t seed, uint64_t a, uint64_t b) {
    uint64_t x = seed ^ (a * 0xD1342543DE82EF95ull) ^ (b * 0xA0761D6478BD642Full);
    for (int i = 0; i < 4; ++i) s[i] = splitmix(x);
  }
  __device__ static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  __device__ uint64_t next() {
    const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  __device__ double uniform() { return ((double)(next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }   // (0, 1)
  __device__ double normal() {
    const double u1 = uniform(), u2 = uniform();
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
  }
  __device__ double gamma(double a) {          // shape a > 0, scale 1 (Marsaglia & Tsang 2000)
    double boost = 1.0;
    if (a < 1.0) {
      boost = pow(uniform(), 1.0 / a);
      a += 1.0;
    }
    const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (int it = 0; it < 1000; ++it) {
      double x, v;
      do {
        x = normal();
        v = 1.0 + c * x;
      } while (v <= 0.0);
      v = v * v * v;
      const double u = uniform();
      if (u < 1.0 - 0.0331 * (x * x) * (x * x) || log(u) < 0.5 * x * x + d * (1.0 - v + log(v))) return boost * d * v;
    }
    return boost * d;
  }
  __device__ double beta(double a, double b) {
    const double x = gamma(a), y = gamma(b);
    return x / (x + y);
  }
};

__device__ __forceinline__ bool bayes_pos(double x) { return x > 2.220446049250313e-16; }

// v0 ~ v0^(c0-1) (v0+v1)^(-c1) (v0+v2)^(-c2): Gamma-proposal Metropolis step, then a log-normal random walk
__device__ double bayes_update_step(double v0, double v1, double v2, double c0, double c1, double c2, Rng& g) {
  const double a = c1 + c2 - c0;
  const double b = (c1 - c0) * v2 + (c2 - c0) * v1;
  const double c = -c0 * v1 * v2;
  const double v_bar = 0.5 * (-b + sqrt(b * b - 4.0 * a * c)) / a;
  const double h = c1 / ((v_bar + v1) * (v_bar + v1)) + c2 / ((v_bar + v2) * (v_bar + v2)) - c0 / (v_bar * v_bar);
  const double k = -h * v_bar * v_bar;
  const double theta = -1.0 / (h * v_bar);
  double log_v0 = log(v0);
  if (bayes_pos(k) && bayes_pos(theta)) {
    const double v_new = g.gamma(k) * theta;
    const double log_new = log(v_new);
    if (bayes_pos(v0) && bayes_pos(v_new)) {
      double lp_new = (c0 - 1.0) * log_new - c1 * log(v_new + v1) - c2 * log(v_new + v2);
      lp_new -= (k - 1.0) * log_new - v_new / theta;
      double lp_old = (c0 - 1.0) * log_v0 - c1 * log(v0 + v1) - c2 * log(v0 + v2);
      lp_old -= (k - 1.0) * log_v0 - v0 / theta;
      if (g.uniform() < exp(fmin(lp_new - lp_old, 0.0))) {
        v0 = v_new;
        log_v0 = log_new;
      }
    }
  }
  const double log_new = log_v0 + g.normal();
  const double v_new = exp(log_new);
  if (bayes_pos(v_new)) {
    if (!bayes_pos(v0)) return v_new;
    const double lp_new = c0 * log_new - c1 * log(v_new + v1) - c2 * log(v_new + v2);
    const double lp_old = c0 * log_v0 - c1 * log(v0 + v1) - c2 * log(v0 + v2);
    if (g.uniform() < exp(fmin(lp_new - lp_old, 0.0))) v0 = v_new;
  }
  return v0;
}

constexpr int kBayesThreads = 256;

// C: B x K x K counts (fp64), X: B x K x K (in: diag(pi) T_mle, out: last state), Tout: B x n_samples x K x K
__global__ void __launch_bounds__(kBayesThreads) bayes_rev_sampler_kernel(const double* __restrict__ C, double* __restrict__ X,
                                                                          int K, int n_samples, int n_steps,
                                                                          unsigned long long seed, double* __restrict__ Tout,
                                                                          double* __restrict__ pi_out) {
  extern __shared__ double s_d[];
  double* sumX = s_d;          // K
  double* sumC = s_d + K;      // K
  __shared__ double s_red[kBayesThreads / 32];
  __shared__ double s_total;
  const int tid = threadIdx.x, lane = tid & 31;
  const size_t chain = blockIdx.x;
  const double* Cc = C + chain * (size_t)K * K;
  double* Xc = X + chain * (size_t)K * K;
  Rng g;
  g.seed(seed, chain, (uint64_t)tid);
  for (int i = tid; i < K; i += kBayesThreads) {
    double sx = 0.0, sc = 0.0;
    for (int j = 0; j < K; ++j) {
      sx += Xc[(size_t)i * K + j];
      sc += Cc[(size_t)i * K + j];
    }
    sumX[i] = sx;
    sumC[i] = sc;
  }
  __syncthreads();
  const int Kp = (K + 1) & ~1;          // even number of players; player K (if any) is a bye
  for (int smp = 0; smp < n_samples; ++smp) {
    for (int step = 0; step < n_steps; ++step) {
      // Row sums from X itself at the start of EVERY sweep, like the reference sampler.  Carrying them
      // incrementally is algebraically exact but numerically unstable: the absolute rounding error of x_i is
      // invariant under the updates while the normalisation rescales it by 1 / total, and E[log(1 / total)] > 0
      // (Jensen) -- the relative error performs a multiplicative random walk with positive drift.
      for (int i = tid >> 5; i < K; i += kBayesThreads / 32) {
        double sx = 0.0;
        for (int j = lane; j < K; j += 32) sx += Xc[(size_t)i * K + j];
        sx = warp_sum(sx);
        if (lane == 0) sumX[i] = sx;
      }
      __syncthreads();
      // ---- diagonal elements: independent of each other
      for (int i = tid; i < K; i += kBayesThreads) {
        const double cii = Cc[(size_t)i * K + i], rest_c = sumC[i] - cii;
        if (cii > 0.0 && bayes_pos(cii) && bayes_pos(rest_c)) {
          const double xii = Xc[(size_t)i * K + i];
          const double t = g.beta(cii, rest_c);
          const double xn = t / (1.0 - t) * (sumX[i] - xii);
          if (bayes_pos(xn)) {
            sumX[i] += xn - xii;
            Xc[(size_t)i * K + i] = xn;
          }
        }
      }
      __syncthreads();
      // ---- off-diagonal elements: round-robin tournament, K / 2 disjoint pairs per round
      for (int r = 0; r < Kp - 1; ++r) {
        for (int t = tid; t < Kp / 2; t += kBayesThreads) {
          int a, b;
          if (t == 0) { a = r; b = Kp - 1; }
          else { a = (r + t) % (Kp - 1); b = (r - t + (Kp - 1)) % (Kp - 1); }
          if (a >= K || b >= K) continue;
          const int i = a > b ? a : b, j = a > b ? b : a;
          const double c0 = Cc[(size_t)i * K + j] + Cc[(size_t)j * K + i];
          if (!(c0 > 0.0)) continue;
          const double xij = Xc[(size_t)i * K + j];
          const double v1 = sumX[i] - xij, v2 = sumX[j] - xij;
          const double xn = bayes_update_step(xij, v1, v2, c0, sumC[i], sumC[j], g);
          Xc[(size_t)i * K + j] = xn;
          Xc[(size_t)j * K + i] = xn;
          sumX[i] = v1 + xn;
          sumX[j] = v2 + xn;
        }
        __syncthreads();
      }
    }
    // ---- normalise X (sum = 1) and emit T = X / x
    double part = 0.0;
    for (int i = tid; i < K; i += kBayesThreads) part += sumX[i];
    part = warp_sum(part);
    if (lane == 0) s_red[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int w = 0; w < kBayesThreads / 32; ++w) tot += s_red[w];
      s_total = tot;
    }
    __syncthreads();
    const double inv = s_total > 0.0 ? 1.0 / s_total : 1.0;
    double* To = Tout + (chain * (size_t)n_samples + smp) * (size_t)K * K;
    for (size_t e = tid; e < (size_t)K * K; e += kBayesThreads) {
      const int i = (int)(e / K), j = (int)(e - (size_t)i * K);
      const double x = Xc[e];
      const double sx = sumX[i];
      To[e] = sx > 0.0 ? x / sx : (i == j ? 1.0 : 0.0);
      Xc[e] = x * inv;
    }
    __syncthreads();
    for (int i = tid; i < K; i += kBayesThreads) {
      sumX[i] *= inv;
      if (pi_out != nullptr) pi_out[(chain * (size_t)n_samples + smp) * (size_t)K + i] = sumX[i];
    }
    __syncthreads();
  }
}

}  // namespace pmb

extern "C" int pmb_bayes_rev_sample(const double* C, double* X, int K, int batch, int n_samples, int n_steps,
                                    uint64_t seed, double* Tout, double* pi_out, pmb_stream_t stream) {
  using namespace pmb;
  PMB_REQUIRE(K > 0 && batch > 0 && n_samples > 0 && n_steps > 0, "pmb_bayes_rev_sample: bad sizes");
  PMB_REQUIRE(C && X && Tout, "pmb_bayes_rev_sample: null pointer");
  const size_t smem = 2 * (size_t)K * sizeof(double);
  PMB_REQUIRE(smem <= 200 * 1024, "pmb_bayes_rev_sample: K = %d too large", K);
  PMB_CUDA(cudaFuncSetAttribute(bayes_rev_sampler_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  bayes_rev_sampler_kernel<<<batch, kBayesThreads, smem, as_stream(stream)>>>(C, X, K, n_samples, n_steps,
                                                                              (unsigned long long)seed, Tout, pi_out);
  PMB_LAUNCH_CHECK();
  return PMB_OK;
}
