"""Bayesian MSM on the device: Gibbs sampler of reversible transition matrices, the ITS
summary statistics and ``EnhancedMSM.compute_implied_timescales`` with confidence intervals.

A sampler is compared in distribution: against the oracle's sequential restatement of deeptime's sampler
(different RNG, different scan order, same conditionals), against the closed-form posterior of a two-state
chain, and through invariants every sample must satisfy exactly."""

from __future__ import annotations

import numpy as np
import pytest
import torch

import oracle
from tests import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    assert torch.cuda.is_available()
    from pmarlo_b200 import load_library

    load_library()
    torch.cuda.set_device(0)
    yield
    torch.cuda.synchronize()


def _sample(C, n_samples, seed=1, n_steps=None):
    from pmarlo_b200.bayes import sample_reversible_matrices

    Ts, pis, T0, pi0 = sample_reversible_matrices(torch.from_numpy(np.asarray(C, dtype=np.float64)).cuda(), None,
                                                  n_samples, n_steps=n_steps, seed=seed)
    return Ts[0].cpu().numpy(), pis[0].cpu().numpy(), T0[0].cpu().numpy(), pi0[0].cpu().numpy()


def test_samples_are_reversible_stochastic_and_centred_on_the_mle():
    rng = np.random.default_rng(3)
    K = 7
    C = rng.poisson(5.0, size=(K, K)).astype(float) + np.diag(rng.poisson(80.0, size=K).astype(float))
    Ts, pis, T0, pi0 = _sample(C, 1500)
    assert Ts.shape == (1500, K, K) and pis.shape == (1500, K)
    np.testing.assert_allclose(Ts.sum(axis=2), 1.0, atol=1e-12)
    np.testing.assert_allclose(pis.sum(axis=1), 1.0, atol=1e-12)
    F = pis[:, :, None] * Ts
    assert float(np.max(np.abs(F - F.transpose(0, 2, 1)))) <= 1e-15          # detailed balance, every sample
    assert float(np.max(np.abs(np.einsum("si,sij->sj", pis, Ts) - pis))) <= 1e-12
    assert np.all(Ts >= 0)
    To, pio, _ = oracle.msm.mle_rev(C)
    np.testing.assert_allclose(T0, To, rtol=1e-6, atol=1e-12)
    # posterior mean close to the MLE, spread ~ sqrt(p (1 - p) / c_i)
    sd = Ts.std(axis=0)
    assert float(np.max(np.abs(Ts.mean(axis=0) - To) / (sd + 1e-3))) < 0.5
    c = C.sum(axis=1)[:, None]
    np.testing.assert_allclose(sd, np.sqrt(To * (1 - To) / c), rtol=0.35, atol=2e-3)


def test_two_state_posterior_is_the_closed_form():
    """Every 2-state chain is reversible and X has three free elements: with the -1 prior the sampler's stationary
    law makes T_01 ~ Beta(c01 + c10 ... ) only jointly; the marginal moments are compared with the oracle sampler,
    and T_01 / T_10 = pi_1 / pi_0 holds for every sample."""
    C = np.array([[90.0, 10.0], [14.0, 60.0]])
    Ts, pis, _, _ = _sample(C, 4000, seed=5, n_steps=3)
    To, pio = oracle.bayes.sample_reversible(C, *oracle.msm.mle_rev(C)[:2], 4000, n_steps=3, seed=9)
    for a, b in ((Ts[:, 0, 1], To[:, 0, 1]), (Ts[:, 1, 0], To[:, 1, 0]), (pis[:, 0], pio[:, 0])):
        se = np.sqrt(a.var() / 400 + b.var() / 400)          # effective sample size >= n / 10
        assert abs(a.mean() - b.mean()) < 4 * se, (a.mean(), b.mean(), se)
        assert abs(a.std() - b.std()) < 0.15 * b.std()
        np.testing.assert_allclose(np.percentile(a, [5, 50, 95]), np.percentile(b, [5, 50, 95]), rtol=0.08)
    np.testing.assert_allclose(Ts[:, 0, 1] * pis[:, 0], Ts[:, 1, 0] * pis[:, 1], rtol=1e-12)


def test_sampler_matches_the_oracle_sampler_in_distribution():
    rng = np.random.default_rng(8)
    K = 5
    C = rng.poisson(3.0, size=(K, K)).astype(float) + np.diag(rng.poisson(40.0, size=K).astype(float))
    C[0, 3] = C[3, 0] = 0.0                     # an element without counts stays zero in every sample
    Ts, pis, _, _ = _sample(C, 3000, seed=2)
    To, pio = oracle.bayes.sample_reversible(C, *oracle.msm.mle_rev(C)[:2], 3000, seed=4)
    assert np.all(Ts[:, 0, 3] == 0) and np.all(Ts[:, 3, 0] == 0)
    m1, m2, s1, s2 = Ts.mean(0), To.mean(0), Ts.std(0), To.std(0)
    se = np.sqrt(s1 ** 2 / 300 + s2 ** 2 / 300) + 1e-12
    assert float(np.max(np.abs(m1 - m2) / se)) < 4.5, np.abs(m1 - m2) / se
    nz = s2 > 1e-6
    np.testing.assert_allclose(s1[nz], s2[nz], rtol=0.2)
    # slowest timescale distribution
    def t2(T):
        ev = np.sort(np.real(np.linalg.eigvals(T)))[::-1]
        return -1.0 / np.log(ev[1])
    a = np.array([t2(T) for T in Ts[::3]])
    b = np.array([t2(T) for T in To[::3]])
    np.testing.assert_allclose(np.percentile(a, [10, 50, 90]), np.percentile(b, [10, 50, 90]), rtol=0.1)


def test_summarize_its_stats_matches_reference_golden(golden):
    """_its.py:543-668 run from the reference file (make_golden.py::make_its_stats) on stored samples: the device
    eigenvalues of every sample + the host percentiles reproduce its nine outputs."""
    from pmarlo_b200 import kernels
    from pmarlo_b200.bayes import summarize_its_stats

    z = golden("its_stats")
    for i, (K, n_ts, lag) in enumerate(z["cases"]):
        Ts = torch.from_numpy(z[f"T_{i}"]).cuda()
        pis = torch.from_numpy(z[f"pi_{i}"]).cuda()
        k = min(int(n_ts) + 1, int(K))
        ev, _ = kernels.eig_rev_topk(Ts, pis, k)
        st = summarize_its_stats(int(lag), ev.cpu().numpy(), int(n_ts), 2.5, 97.5)
        for j in range(9):
            np.testing.assert_allclose(st[j], z[f"stat_{i}_{j}"], rtol=1e-7, equal_nan=True, err_msg=f"case {i} output {j}")


def test_enhanced_msm_bayesian_its_with_confidence_intervals():
    from pmarlo_b200 import EnhancedMSM

    dtrajs = synth.metastable_dtrajs(6, 6000, 8, seed=21, stay=0.85)
    m = EnhancedMSM(dtrajs, n_states=8)
    m.random_state = 3
    m.build_msm(lag_time=5)
    lags = [1, 2, 5, 10]
    m.compute_implied_timescales(lags, n_timescales=3, n_samples=200, ci=0.9)
    its = m.implied_timescales
    assert its.timescales.shape == (4, 3) and its.timescales_ci.shape == (4, 3, 2)
    assert np.all(np.isfinite(its.timescales)) and np.all(np.isfinite(its.timescales_ci))
    assert np.all(its.timescales_ci[:, :, 0] <= its.timescales) and np.all(its.timescales <= its.timescales_ci[:, :, 1])
    assert np.all(its.eigenvalues_ci[:, :, 0] <= its.eigenvalues) and np.all(its.eigenvalues <= its.eigenvalues_ci[:, :, 1])
    # the posterior median brackets the deterministic maximum-likelihood timescales
    ref = oracle.msm.its_rev_mle(dtrajs, 8, lags, 3)
    assert np.all(ref >= its.timescales_ci[:, :, 0] * 0.9) and np.all(ref <= its.timescales_ci[:, :, 1] * 1.1)
    np.testing.assert_allclose(its.timescales, ref, rtol=0.15)
    np.testing.assert_allclose(its.rates, 1.0 / its.timescales, rtol=0.05)
    # same seed, same samples
    m.compute_implied_timescales(lags, n_timescales=3, n_samples=200, ci=0.9)
    np.testing.assert_array_equal(m.implied_timescales.timescales, its.timescales)
    # estimator="mle": the deterministic sweep, NaN intervals
    m.compute_implied_timescales(lags, n_timescales=3, estimator="mle")
    np.testing.assert_allclose(m.implied_timescales.timescales, ref, rtol=1e-6)
    assert np.all(np.isnan(m.implied_timescales.timescales_ci))
    out = m.sample_bayesian_timescales(n_samples=50)
    assert out["timescales_samples"].shape[0] == 50 and out["population_samples"].shape == (50, 8)
    np.testing.assert_allclose(out["population_samples"].sum(axis=1), 1.0, atol=1e-12)
