"""CPU suite: the plain-C oracle (oracle/csrc/oracle_c.c) is pinned to the numpy oracle, which is itself
pinned to the reference-generated golden vectors (test_oracle_golden.py); the synthetic generators of the
BASELINE configs do what their docstrings say; the oracle chain is self-consistent."""

from __future__ import annotations

import numpy as np
import pytest

import oracle
from oracle import cext
from tests import synth


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not cext.available():
        import subprocess, pathlib

        subprocess.run(["make", "-C", str(pathlib.Path(oracle.__file__).resolve().parent)], check=True)
    assert cext.available()


def test_c_counts_equal_numpy_counts_and_golden(golden):
    z = golden("counts")
    d = synth.metastable_dtrajs(5, 30000, 60, 11)
    d[2][100:110] = -1
    for lag in (1, 7, 50):
        np.testing.assert_array_equal(cext.count_lagged(d, 60, lag),
                                      oracle.counts.count_lagged(d, 60, lag, mode="endpoint"))
    assert len(z.files) > 0


@pytest.mark.parametrize("K,threads", [(3, 1), (40, 1), (300, 4), (1100, 8)])
def test_c_mle_equals_numpy_mle(K, threads):
    rng = np.random.default_rng(K)
    C = rng.poisson(2.0, size=(K, K)).astype(float) * (rng.random((K, K)) < 0.3) + 1e-3
    T, pi, it = oracle.msm.mle_rev(C, maxerr=1e-10)
    T2, pi2, it2 = cext.mle_rev(C, maxerr=1e-10, threads=threads)
    assert it == it2
    np.testing.assert_allclose(T2, T, rtol=1e-11, atol=1e-15)
    np.testing.assert_allclose(pi2, pi, rtol=1e-11)
    # fixed iteration count: same iterate
    Ta, pa, ita = oracle.msm.mle_rev(C, maxerr=0.0, maxiter=7)
    Tb, pb, itb = cext.mle_rev(C, maxerr=0.0, maxiter=7, threads=threads)
    assert ita == itb == 7
    np.testing.assert_allclose(pb, pa, rtol=1e-12)
    with pytest.raises(ValueError):
        cext.mle_rev(np.zeros((3, 3)))


def test_c_its_sweep_equals_numpy_sweep():
    d = synth.metastable_dtrajs(4, 8000, 12, 5, stay=0.8)
    lags = [1, 2, 5, 10]
    ref = oracle.msm.its_rev_mle(d, 12, lags, 4)
    got, counts, iters = cext.its_rev_mle(d, 12, lags, 4, threads=3)
    np.testing.assert_allclose(got, ref, rtol=1e-9, equal_nan=True)
    assert all(int(c.sum()) == sum(max(0, x.size - lag) for x in d) for c, lag in zip(counts, lags))


def test_ala2_generator_drives_phi_psi(topologies):
    t = topologies["ala2"]
    trajs, walks = synth.ala2_trajectories(t, n_traj=3, n_frames=[50, 51, 52], seed=2, jitter=0.0)
    assert [x.shape for x in trajs] == [(50, 22, 3), (51, 22, 3), (52, 22, 3)] and trajs[0].dtype == np.float32
    for x, w in zip(trajs, walks):
        ang = oracle.featurize.featurize_trajectory(x, t["names"], t["resid"], t["chain"], "phi_psi")
        d = np.abs(ang - w)
        assert float(np.max(np.minimum(d, 2 * np.pi - d))) < 1e-5
    # bond lengths are untouched by the rotations
    b0 = np.linalg.norm(t["xyz"][8] - t["xyz"][6])
    np.testing.assert_allclose(np.linalg.norm(trajs[0][:, 8] - trajs[0][:, 6], axis=1), b0, rtol=1e-5)
    full, _ = synth.ala2_trajectories(t, seed=1)
    assert len(full) == 35 and sum(x.shape[0] for x in full) == 13_000


def test_structure_generator_is_ar1(topologies):
    t = topologies["chig"]
    tr = synth.structure_trajectories(t["xyz"], 2, 20000, seed=3, rho=0.99, sigma=0.03)
    x = tr[0][:, 5, 0].astype(np.float64) - float(t["xyz"][5, 0])
    assert abs(np.std(x) - 0.03) < 0.006
    r1 = float(np.corrcoef(x[:-1], x[1:])[0, 1])
    assert abs(r1 - 0.99) < 0.01


def test_oracle_chain_runs_and_is_consistent():
    feats = synth.ar1_features(3, 3000, 6, 2)
    rows = np.arange(0, 3000, 100)
    r = oracle.pipeline.run_chain(feats, tica_lag=5, tica_dim=3, n_states=30, init_rows=rows, kmeans_iters=4,
                                  kmeans_tolerance=None, msm_lag=5, n_timescales=3)
    assert r.labels.shape == (9000,) and r.counts.sum() == 3 * (3000 - 5)
    np.testing.assert_allclose(r.T.sum(axis=1), 1.0, atol=1e-12)
    np.testing.assert_allclose(r.pi @ r.T, r.pi, atol=1e-10)
    assert abs(r.eigenvalues[0] - 1.0) < 1e-9
    # no-TICA branch (config C2)
    r2 = oracle.pipeline.run_chain([f[:, :2] for f in feats], tica_dim=0, n_states=10, init_rows=np.arange(10),
                                   kmeans_iters=3, kmeans_tolerance=None, msm_lag=2, n_timescales=2)
    assert r2.tica is None and r2.Y.shape == (9000, 2)


def test_oracle_its_summary_matches_reference_golden(golden):
    """oracle.bayes.summarize_its_stats against _its.py:543-668 run from the reference file (its_stats.npz)."""
    z = golden("its_stats")
    for i, (K, n_ts, lag) in enumerate(z["cases"]):
        st = oracle.bayes.summarize_its_stats(int(lag), z[f"T_{i}"], int(n_ts), 2.5, 97.5)
        for j in range(9):
            np.testing.assert_allclose(st[j], z[f"stat_{i}_{j}"], rtol=1e-12, equal_nan=True)


def test_oracle_sampler_invariants():
    C = np.array([[50, 5, 1], [6, 80, 4], [2, 3, 30]], float)
    T0, pi0, _ = oracle.msm.mle_rev(C)
    Ts, pis = oracle.bayes.sample_reversible(C, T0, pi0, 300, seed=1)
    F = pis[:, :, None] * Ts
    assert np.max(np.abs(F - F.transpose(0, 2, 1))) < 1e-15
    np.testing.assert_allclose(Ts.sum(axis=2), 1.0, atol=1e-12)
    assert np.max(np.abs(Ts.mean(axis=0) - T0)) < 0.02
