/* oracle_c.c -- plain-C restatement of the K^2-sized CPU stages of the oracle (TEST INFRASTRUCTURE).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this.  It restates, in compiled
 * code like the reference's own dependency, what oracle/msm.py::mle_rev does in numpy:
 *
 *   deeptime 0.4.5 MaximumLikelihoodMSM(reversible=True) -> tools.estimation.dense.mle.mle_trev
 *   (msmtools _mle_trev.c lineage), called at src/pmarlo/markov_state_model/_msm_utils.py:255-261 and
 *   ck_its_selector.py:397-399.  Matrix-free form (only the row-sum vector x feeds back):
 *     c_i = sum_j C_ij;  x_i <- sum_j (C_ij + C_ji) / (c_i / x_i + c_j / x_j), normalised;
 *     err = max_i |x_i - x'_i| / (0.5 (x_i + x'_i));  stop at err <= maxerr or maxiter.
 *   T_ij = X_ij / sum_j X_ij,  pi_i = sum_j X_ij / sum X.
 *
 * deeptime is absent from this image: PARITY UNPINNED against its binary; pinned against oracle/msm.py
 * (tests/test_oracle_golden.py) which is itself validated by the reference assertions that run here.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Worker team for the fixed point: every thread owns a contiguous block of rows; the iteration is
 * three phases separated by barriers (q = c / x; row sums; normalise + error on thread 0). */
typedef struct {
  int K, nthreads;
  const double* S;
  const double* c;
  double *x, *xn, *q, *part;
  double maxerr, err;
  int64_t maxiter, it;
  pthread_barrier_t bar;
} mle_team_t;

typedef struct { mle_team_t* team; int tid; } mle_arg_t;

static void* mle_worker(void* argp) {
  mle_arg_t* a = (mle_arg_t*)argp;
  mle_team_t* t = a->team;
  const size_t n = (size_t)t->K;
  const size_t lo = n * (size_t)a->tid / (size_t)t->nthreads, hi = n * (size_t)(a->tid + 1) / (size_t)t->nthreads;
  for (;;) {
    pthread_barrier_wait(&t->bar);                 /* x, it, err of the previous iteration are visible */
    if (!(t->it < t->maxiter && t->err > t->maxerr)) break;
    for (size_t i = lo; i < hi; ++i) t->q[i] = t->c[i] / t->x[i];
    pthread_barrier_wait(&t->bar);
    double sum = 0.0;
    for (size_t i = lo; i < hi; ++i) {
      const double qi = t->q[i];
      const double* Si = t->S + i * n;
      double acc = 0.0;
      for (size_t j = 0; j < n; ++j) acc += Si[j] / (qi + t->q[j]);
      t->xn[i] = acc;
      sum += acc;
    }
    t->part[a->tid] = sum;
    pthread_barrier_wait(&t->bar);
    if (a->tid == 0) {
      double tot = 0.0, err = 0.0;
      for (int w = 0; w < t->nthreads; ++w) tot += t->part[w];
      for (size_t i = 0; i < n; ++i) {
        const double v = t->xn[i] / tot;
        const double e = fabs(t->x[i] - v) / (0.5 * (t->x[i] + v));
        if (e > err) err = e;
        t->x[i] = v;
      }
      t->err = err;
      t->it += 1;
    }
  }
  return NULL;
}

/* Returns the number of iterations, or -1 when a state has no outgoing counts.  nthreads <= 1: serial. */
int64_t oracle_mle_rev(const double* C, int K, double maxerr, int64_t maxiter, double* T, double* pi, int nthreads) {
  const size_t n = (size_t)K;
  double* S = (double*)malloc(n * n * sizeof(double));
  double* c = (double*)malloc(n * sizeof(double));
  double* x = (double*)malloc(n * sizeof(double));
  double* xn = (double*)malloc(n * sizeof(double));
  double* q = (double*)malloc(n * sizeof(double));
  int64_t it = 0;
  double tot = 0.0;
  for (size_t i = 0; i < n; ++i) {
    double ci = 0.0, si = 0.0;
    for (size_t j = 0; j < n; ++j) {
      ci += C[i * n + j];
      S[i * n + j] = C[i * n + j] + C[j * n + i];
      si += S[i * n + j];
    }
    c[i] = ci;
    x[i] = si;
    tot += si;
  }
  for (size_t i = 0; i < n; ++i) {
    if (!(c[i] > 0.0)) { it = -1; goto done; }
    x[i] /= tot;
  }
  {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    if ((size_t)nthreads > n) nthreads = (int)n;
    mle_team_t team;
    team.K = K; team.nthreads = nthreads; team.S = S; team.c = c; team.x = x; team.xn = xn; team.q = q;
    team.part = (double*)malloc((size_t)nthreads * sizeof(double));
    team.maxerr = maxerr; team.err = INFINITY; team.maxiter = maxiter; team.it = 0;
    pthread_barrier_init(&team.bar, NULL, (unsigned)nthreads);
    pthread_t th[64];
    mle_arg_t args[64];
    for (int w = 0; w < nthreads; ++w) { args[w].team = &team; args[w].tid = w; }
    for (int w = 1; w < nthreads; ++w) pthread_create(&th[w], NULL, mle_worker, &args[w]);
    mle_worker(&args[0]);
    for (int w = 1; w < nthreads; ++w) pthread_join(th[w], NULL);
    pthread_barrier_destroy(&team.bar);
    free(team.part);
    it = team.it;
  }
  {
    for (size_t i = 0; i < n; ++i) q[i] = c[i] / x[i];
    double tot2 = 0.0;
    for (size_t i = 0; i < n; ++i) {
      double rs = 0.0;
      for (size_t j = 0; j < n; ++j) {
        const double v = S[i * n + j] / (q[i] + q[j]);
        T[i * n + j] = v;
        rs += v;
      }
      for (size_t j = 0; j < n; ++j) T[i * n + j] /= rs;
      pi[i] = rs;
      tot2 += rs;
    }
    for (size_t i = 0; i < n; ++i) pi[i] /= tot2;
  }
done:
  free(S); free(c); free(x); free(xn); free(q);
  return it;
}

/* Sliding-window lagged counts of one label trajectory, accumulated into C (K x K int64): the dense
 * equivalent of deeptime's TransitionCountEstimator(count_mode="sliding") COO sum
 * (_msm_utils.py:238-246); pairs with an endpoint outside [0, K) are skipped (ck_runner.py:70-83). */
void oracle_count_lagged(const int64_t* labels, int64_t n, int K, int lag, int64_t* C) {
  for (int64_t t = 0; t + lag < n; ++t) {
    const int64_t a = labels[t], b = labels[t + lag];
    if (a >= 0 && a < K && b >= 0 && b < K) C[a * (int64_t)K + b] += 1;
  }
}
