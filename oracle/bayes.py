"""Oracle: Gibbs sampler of reversible transition matrices and the ITS summary statistics
(TEST INFRASTRUCTURE).

* ``sample_reversible`` restates deeptime 0.4.5 ``TransitionMatrixSampler(reversible=True)`` (C++ ``SamplerRev``,
  the algorithm of Trendelkamp-Schroer, Wu, Paul, Noe, J. Chem. Phys. 143, 174101 (2015)) as driven by
  ``BayesianMSM(lagtime, n_samples).fit`` at src/pmarlo/markov_state_model/_its.py:289-310: sequential scan over
  the lower triangle, Beta update of the diagonal, Gamma-proposal + log-normal random-walk Metropolis updates of
  the off-diagonal elements, ``-1`` prior, row sums recomputed and X normalised at every sweep, n_steps = sqrt(K)
  sweeps per sample.
  deeptime is absent from this image: PARITY UNPINNED against its binary (and a sampler can only be compared
  in distribution anyway).  The device sampler visits the same conditionals in a round-robin order; the two are
  compared statistically (tests/test_gpu_bayes.py).
* ``summarize_its_stats`` restates ``ITSMixin._summarize_its_stats`` (_its.py:543-668); pinned by
  tests/golden/its_stats.npz, generated from the reference file itself.
"""

from __future__ import annotations

import numpy as np

from . import msm

EPS = np.finfo(float).eps


def _update_step(v0, v1, v2, c0, c1, c2, rng):
    a = c1 + c2 - c0
    b = (c1 - c0) * v2 + (c2 - c0) * v1
    c = -c0 * v1 * v2
    v_bar = 0.5 * (-b + np.sqrt(b * b - 4.0 * a * c)) / a
    h = c1 / (v_bar + v1) ** 2 + c2 / (v_bar + v2) ** 2 - c0 / v_bar ** 2
    k = -h * v_bar * v_bar
    theta = -1.0 / (h * v_bar)
    log_v0 = np.log(v0)
    if k > EPS and theta > EPS:
        v_new = rng.gamma(k, theta)
        log_new = np.log(v_new)
        if v0 > EPS and v_new > EPS:
            lp_new = (c0 - 1.0) * log_new - c1 * np.log(v_new + v1) - c2 * np.log(v_new + v2)
            lp_new -= (k - 1.0) * log_new - v_new / theta
            lp_old = (c0 - 1.0) * log_v0 - c1 * np.log(v0 + v1) - c2 * np.log(v0 + v2)
            lp_old -= (k - 1.0) * log_v0 - v0 / theta
            if rng.random() < np.exp(min(lp_new - lp_old, 0.0)):
                v0, log_v0 = v_new, log_new
    log_new = log_v0 + rng.normal()
    v_new = np.exp(log_new)
    if v_new > EPS:
        if not v0 > EPS:
            return v_new
        lp_new = c0 * log_new - c1 * np.log(v_new + v1) - c2 * np.log(v_new + v2)
        lp_old = c0 * log_v0 - c1 * np.log(v0 + v1) - c2 * np.log(v0 + v2)
        if rng.random() < np.exp(min(lp_new - lp_old, 0.0)):
            v0 = v_new
    return v0


def sample_reversible(C: np.ndarray, T0: np.ndarray, pi0: np.ndarray, n_samples: int, n_steps: int | None = None,
                      seed: int = 0):
    """Returns (T samples (n_samples, K, K), pi samples (n_samples, K))."""
    rng = np.random.default_rng(seed)
    C = np.asarray(C, dtype=float)
    K = C.shape[0]
    X = pi0[:, None] * np.asarray(T0, dtype=float)
    X = 0.5 * (X + X.T)
    sumC = C.sum(axis=1)
    n_steps = max(1, int(np.sqrt(K))) if n_steps is None else int(n_steps)
    Ts, pis = np.empty((n_samples, K, K)), np.empty((n_samples, K))
    for s in range(n_samples):
        for _ in range(n_steps):
            sumX = X.sum(axis=1)
            for i in range(K):
                for j in range(i + 1):
                    if not C[i, j] + C[j, i] > 0:
                        continue
                    if i == j:
                        if C[i, i] > EPS and sumC[i] - C[i, i] > EPS:
                            t = rng.beta(C[i, i], sumC[i] - C[i, i])
                            xn = t / (1.0 - t) * (sumX[i] - X[i, i])
                            if xn > EPS:
                                sumX[i] += xn - X[i, i]
                                X[i, i] = xn
                    else:
                        v1, v2 = sumX[i] - X[i, j], sumX[j] - X[j, i]
                        xn = _update_step(X[i, j], v1, v2, C[i, j] + C[j, i], sumC[i], sumC[j], rng)
                        X[i, j] = X[j, i] = xn
                        sumX[i], sumX[j] = v1 + xn, v2 + xn
            X /= X.sum()
        rs = X.sum(axis=1)
        Ts[s] = X / rs[:, None]
        pis[s] = rs / rs.sum()
    return Ts, pis


def summarize_its_stats(lag: int, matrices: np.ndarray, n_timescales: int, q_low: float, q_high: float):
    """_its.py:543-668: per sample the eigenvalues sorted by real part, lambda_2 .. lambda_{n+1} clipped to
    [1e-12, 1 - 1e-12], timescales by ``safe_timescales``; nanmedian / nanpercentile over the samples, NaN-padded
    to ``n_timescales``.  Returns the reference's 9-tuple."""
    import warnings

    eig_s, ts_s = [], []
    for T in np.asarray(matrices, dtype=float):
        n = int(max(0, n_timescales))
        k = n + 1 if n > 0 else 1
        ev = np.linalg.eigvals(T)
        ev = ev[np.argsort(-np.abs(ev), kind="stable")]
        if k < T.shape[0]:
            ev = ev[:k]
        ev = ev[np.argsort(-np.real(ev), kind="stable")]
        e = np.clip(np.abs(np.real(ev[1:1 + n])), 1e-12, 1.0 - 1e-12) if n > 0 else np.empty((0,))
        eig_s.append(e.astype(float))
        ts_s.append(msm.safe_timescales(int(max(1, lag)), e) if n > 0 else np.empty((0,)))
    eig_arr, ts_arr = np.asarray(eig_s, dtype=float), np.asarray(ts_s, dtype=float)
    with warnings.catch_warnings():
        warnings.filterwarnings("ignore", category=RuntimeWarning)
        rate_arr = np.reciprocal(ts_arr, where=np.isfinite(ts_arr), out=np.full_like(ts_arr, np.nan))
        out = []
        for arr in (eig_arr, ts_arr, rate_arr):
            out += [np.nanmedian(arr, axis=0), np.nanpercentile(arr, q_low, axis=0), np.nanpercentile(arr, q_high, axis=0)]

    def pad(v):
        o = np.full((int(n_timescales),), np.nan)
        kk = min(v.shape[0], int(n_timescales))
        o[:kk] = v[:kk]
        return o

    return tuple(pad(np.asarray(v, dtype=float)) for v in out)
