"""Oracle: the whole hot path on the host cores (TEST INFRASTRUCTURE / CPU baseline).

featurize -> z-score -> TICA -> k-means -> lagged counts -> reversible MLE ->
eigenvalues -> implied timescales, restating the chain pmarlo runs at
src/pmarlo/api/conformations.py:192-200 with the libraries that exist in this
image (numpy/scipy BLAS, scikit-learn), as laid out in BASELINE.md section 3:

* k-means fit uses ``sklearn.cluster.KMeans(init=centres, n_init=1, max_iter=it,
  tol=0, algorithm="lloyd")`` as the multi-threaded stand-in for deeptime's
  OpenMP Lloyd loop (clustering.py:349-355); final labels are sklearn's
  ``predict`` (fp64).  ``exact=True`` switches to the fp64 direct-difference
  Lloyd of ``oracle.kmeans`` (slow, parity-grade).

Only ``bench.py`` (``cpu_baseline`` leg and ``--impl reference``) and tests call this.
"""

from __future__ import annotations

import time
from dataclasses import dataclass

import numpy as np

from . import counts, featurize, kmeans, msm, tica


@dataclass
class CpuResult:
    timescales: np.ndarray
    eigenvalues: np.ndarray
    stage_seconds: dict
    n_frames: int


def block_features(xyz: np.ndarray, phi_q: np.ndarray, psi_q: np.ndarray, pairs: np.ndarray) -> np.ndarray:
    """[cos phi | sin phi | cos psi | sin psi | distances] float32 (mdtraj is float32)."""
    x = np.asarray(xyz, dtype=np.float32)
    phi = featurize.compute_dihedrals(x, phi_q)
    psi = featurize.compute_dihedrals(x, psi_q)
    dist = featurize.compute_distances(x, pairs)
    return np.hstack([np.cos(phi), np.sin(phi), np.cos(psi), np.sin(psi), dist]).astype(np.float32)


def run(trajs_xyz, phi_q, psi_q, pairs, *, tica_lag, tica_dim, n_states, kmeans_iters, msm_lag,
        n_timescales, seed=0, mle_maxerr=1e-8, mle_maxiter=1_000_000, exact=False, chunk=20000,
        threads=1) -> CpuResult:
    st = {}
    t0 = time.perf_counter()
    feats = []
    for x in trajs_xyz:
        parts = [block_features(x[s:s + chunk], phi_q, psi_q, pairs) for s in range(0, x.shape[0], chunk)]
        feats.append(np.concatenate(parts, axis=0))
    st["featurize"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    lengths = [f.shape[0] for f in feats]
    flat = np.concatenate(feats, axis=0).astype(np.float64)
    Z = tica.preprocess(flat, scale=True)
    off = np.concatenate([[0], np.cumsum(lengths)])
    prepped = [Z[off[i]:off[i + 1]] for i in range(len(lengths))]
    model = tica.tica_fit(prepped, tica_lag)
    st["tica_fit"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    Y = tica.tica_transform(model, Z, tica_dim).astype(np.float32).astype(np.float64)
    st["project"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    rng = np.random.default_rng(seed)
    c0 = Y[np.sort(rng.choice(lengths[0], size=n_states, replace=False))]
    if exact:
        c = c0.copy()
        for _ in range(kmeans_iters):
            lab, _ = kmeans.assign(Y, c)
            c, _ = kmeans._update(Y, lab, c)
        labels, _ = kmeans.assign(Y, c)
    else:
        from sklearn.cluster import KMeans

        km = KMeans(n_clusters=n_states, init=c0, n_init=1, max_iter=kmeans_iters, tol=0.0, algorithm="lloyd")
        km.fit(Y)
        labels = km.predict(Y)
    st["kmeans"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    dtrajs = [labels[off[i]:off[i + 1]] for i in range(len(lengths))]
    C = counts.count_lagged(dtrajs, n_states, msm_lag).astype(float)
    st["count"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    Ca, active = msm.ensure_connected_counts(C)
    from . import cext

    if cext.available():   # same fixed point, compiled and row-parallel (pinned to msm.mle_rev in tests/test_oracle_c.py)
        T, pi, mle_iters = cext.mle_rev(Ca, maxerr=mle_maxerr, maxiter=mle_maxiter, threads=threads)
    else:
        T, pi, mle_iters = msm.mle_rev(Ca, maxerr=mle_maxerr, maxiter=mle_maxiter)
    st["mle"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    ev = msm.eigenvalues_rev(T, pi, min(n_timescales + 1, T.shape[0]))
    ts = msm.safe_timescales(msm_lag, ev[1:])
    st["eig"] = time.perf_counter() - t0
    st["mle_iters"] = float(mle_iters)
    return CpuResult(ts, ev, st, int(sum(lengths)))


@dataclass
class ChainResult:
    """Every intermediate of the chain, for the chained parity tests of the BASELINE configs."""
    tica: object | None
    Y: np.ndarray            # fp64 view of the fp32 projected coordinates the reference would cluster
    centers: np.ndarray
    labels: np.ndarray
    kmeans_iters: int
    counts: np.ndarray
    T: np.ndarray
    pi: np.ndarray
    eigenvalues: np.ndarray
    timescales: np.ndarray


def run_chain(feats, *, preprocess="standard", tica_lag=10, tica_dim=2, n_states=100, init_rows=None,
              kmeans_iters=20, kmeans_tolerance=None, msm_lag=10, n_timescales=5, alpha=1e-3,
              mle_maxerr=1e-8, mle_maxiter=1_000_000) -> ChainResult:
    """features (list of per-trajectory fp32 arrays) -> z-score -> TICA -> exact fp64 Lloyd from the centres
    at frames ``init_rows`` -> sliding counts -> +alpha -> reversible MLE -> eigenvalues -> timescales
    (src/pmarlo/api/conformations.py:192-200 with reduce_features / cluster_microstates(initial_centers=) /
    build_msm_from_labels).  ``tica_dim <= 0`` clusters the features directly (config C2).
    ``kmeans_tolerance=None`` runs exactly ``kmeans_iters`` Lloyd updates (the benchmark's fixed-iteration mode)."""
    lengths = [int(f.shape[0]) for f in feats]
    off = np.concatenate([[0], np.cumsum(lengths)])
    flat = np.concatenate([np.asarray(f, dtype=np.float64) for f in feats], axis=0)
    model = None
    if tica_dim > 0:
        Z = flat if preprocess is None else tica.preprocess(flat, scale=(preprocess == "standard"))
        prepped = [Z[off[i]:off[i + 1]] for i in range(len(lengths))]
        model = tica.tica_fit(prepped, tica_lag)
        Y = tica.tica_transform(model, Z, tica_dim).astype(np.float32).astype(np.float64)
    else:
        Y = flat.astype(np.float32).astype(np.float64)
    c = Y[np.asarray(init_rows, dtype=np.int64)].copy()
    if kmeans_tolerance is None:
        for _ in range(kmeans_iters):
            lab, _ = kmeans.assign(Y, c)
            c, _ = kmeans._update(Y, lab, c)
        n_it = kmeans_iters
    else:
        c, n_it, _, _ = kmeans.lloyd(Y, c, kmeans_iters, kmeans_tolerance)
    labels, _ = kmeans.assign(Y, c)
    dtrajs = [labels[off[i]:off[i + 1]] for i in range(len(lengths))]
    C = counts.count_lagged(dtrajs, n_states, msm_lag)
    Ca, active = msm.ensure_connected_counts(C.astype(float), alpha=alpha)
    from . import cext

    mle = cext.mle_rev if cext.available() else msm.mle_rev   # same algorithm, compiled (pinned to each other)
    Ta, pia, _ = mle(Ca, maxerr=mle_maxerr, maxiter=mle_maxiter)
    T, pi = msm.expand_results(n_states, active, Ta, pia)
    ev = msm.eigenvalues_rev(Ta, pia, min(n_timescales + 1, Ta.shape[0]))
    return ChainResult(model, Y, c, labels, n_it, C, T, pi, ev, msm.safe_timescales(msm_lag, ev[1:]))
