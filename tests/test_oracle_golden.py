"""The oracle against every vector the importable reference functions produced
(tests/golden/make_golden.py) and against analytic known answers.  CPU only."""

from __future__ import annotations

import numpy as np
import pytest

from oracle import counts as ocounts
from oracle import featurize as ofeat
from oracle import kmeans as okm
from oracle import msm as omsm
from oracle import tica as otica


# ------------------------------------------------------------------ counts
def test_counts_match_reference_functions(golden):
    z = golden("counts")
    for ci in range(int(z["n_cases"])):
        g = lambda k: z[f"c{ci}_{k}"]
        labels, bounds = g("labels"), g("bounds")
        K, lag, stride = int(g("K")), int(g("lag")), int(g("stride"))
        segs = [(int(bounds[i]), int(bounds[i + 1])) for i in range(len(bounds) - 1)]
        C, tp = ocounts.weighted_counts(labels, n_states=K, lag_time=lag, segments=segs, stride=stride)
        assert np.array_equal(C, g("C_u")) and tp == int(g("tp_u"))
        C, tp = ocounts.weighted_counts(labels, n_states=K, lag_time=lag, weights=g("weights"),
                                        segments=segs, stride=stride)
        np.testing.assert_allclose(C, g("C_w"), rtol=1e-13, atol=0)
        assert tp == int(g("tp_w"))
        C, tp = ocounts.weighted_counts(labels, n_states=K, lag_time=lag)
        assert np.array_equal(C, g("C_all")) and tp == int(g("tp_all"))
        dtrajs = [labels[s:e] for s, e in segs]
        assert np.array_equal(ocounts.count_lagged(dtrajs, K, lag), g("C_sl").astype(np.int64))
        assert np.array_equal(ocounts.count_lagged(dtrajs, K, lag, step=lag), g("C_st").astype(np.int64))
        assert int(ocounts.count_lagged(dtrajs, K, lag).sum()) == int(g("n_sl"))
        assert ocounts.expected_pairs([e - s for s, e in segs], lag, stride) == int(g("ep"))


def test_expected_pairs_bruteforce():
    # tests/unit/analysis/test_counting.py:23-66 restated without hypothesis
    rng = np.random.default_rng(0)
    for _ in range(200):
        lengths = rng.integers(0, 40, size=rng.integers(1, 6)).tolist()
        tau = int(rng.integers(0, 12))
        stride = int(rng.integers(1, 5))
        brute = sum(len(range(0, L - tau, stride)) for L in lengths if L - tau > 0)
        assert ocounts.expected_pairs(lengths, tau, stride) == brute
    with pytest.raises(ValueError):
        ocounts.expected_pairs([3], -1)
    with pytest.raises(ValueError):
        ocounts.expected_pairs([3], 1, 0)


def test_split_mode_drops_spanning_pairs():
    d = np.array([0, 1, -1, 1, 0, 0, 1])
    Ce = ocounts.count_lagged([d], 2, 2, mode="endpoint")
    Cs = ocounts.count_lagged([d], 2, 2, mode="split")
    # endpoint mode keeps (1 -> 1) spanning the invalid frame, split mode drops it
    assert Ce.sum() == 3 and Cs.sum() == 2
    assert Ce[1, 1] == 1 and Cs[1, 1] == 0


# -------------------------------------------------------------- timescales
def test_safe_timescales_matches_reference(golden):
    z = golden("timescales")
    for lag in (1, 10, 400):
        np.testing.assert_array_equal(omsm.safe_timescales(lag, z["ev_real"]), z[f"ts_real_{lag}"])
        np.testing.assert_array_equal(omsm.safe_timescales(lag, z["ev_cplx"]), z[f"ts_cplx_{lag}"])
    assert omsm.safe_timescales(5, np.array([])).shape == (0,)


# -------------------------------------------------------------- preprocess
def test_preprocess_matches_reference(golden):
    z = golden("preprocess")
    np.testing.assert_allclose(otica.preprocess(z["X"], True), z["P_scale"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(otica.preprocess(z["X"], False), z["P_noscale"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(otica.preprocess(z["Xn"], True), z["Pn_scale"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(otica.preprocess(z["Xn"], False), z["Pn_noscale"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(otica.preprocess(z["X"][:, 1], True), z["P_1d"], rtol=1e-12, atol=1e-12)


# -------------------------------------------------------------------- TICA
def test_tica_against_reference_numpy_crosscheck(golden):
    z = golden("tica_xcheck")
    X, lag = z["X"].astype(np.float64), int(z["lag"])
    model = otica.tica_fit([X], lag)
    # _estimate_top_eigenvalues is non-symmetrised with Bessel: agrees to O(1/sqrt(N))
    np.testing.assert_allclose(model.eigenvalues, z["ref_eigs"], atol=1e-2)
    np.testing.assert_allclose(model.eigenvalues, z["rho"] ** lag, atol=2e-2)


def test_tica_invariants_and_conventions():
    rng = np.random.default_rng(5)
    X = np.cumsum(rng.standard_normal((5000, 6)), axis=0) * 0.01 + rng.standard_normal((5000, 6))
    X[:, 5] = X[:, 0] + X[:, 1]  # rank deficient -> epsilon cut
    trajs = [X[:3000], X[3000:]]
    m = otica.tica_fit(trajs, 3)
    assert m.rank == 5
    R = m.eigenvectors / m.eigenvalues[None, :]
    np.testing.assert_allclose(R.T @ m.C00 @ R, np.eye(5), atol=1e-8)
    np.testing.assert_allclose(R.T @ m.C0t @ R, np.diag(m.eigenvalues), atol=1e-8)
    assert np.all(np.diff(np.abs(m.eigenvalues)) <= 1e-15)
    for j in range(R.shape[1]):
        assert R[np.argmax(np.abs(R[:, j])), j] > 0
    Y, nc = otica.maybe_apply_tica(X, [3000, 2000], 9, 3)
    assert nc == 5 and Y.shape == (5000 - 2 * 3, 5)


# ----------------------------------------------------------------- k-means
def test_assignment_matches_sklearn_predict(golden):
    z = golden("assign")
    std = np.where(z["std"] > 1e-10, z["std"], 1.0)
    for X, lab in ((z["X"], z["lab_train"]), (z["Xt"], z["lab_test"])):
        got, _ = okm.assign((X - z["mean"]) / std, z["centers"])
        assert np.array_equal(got, lab.astype(np.int64))


def test_lloyd_matches_sklearn_lloyd():
    from sklearn.cluster import KMeans

    rng = np.random.default_rng(1)
    c0 = rng.standard_normal((8, 3)) * 5
    Y = c0[rng.integers(0, 8, 4000)] + rng.standard_normal((4000, 3))
    init = Y[rng.choice(4000, 8, replace=False)]
    centers, n_iter, cost, conv = okm.lloyd(Y, init, max_iter=500, tolerance=1e-12)
    km = KMeans(n_clusters=8, init=init, n_init=1, max_iter=500, tol=0.0, algorithm="lloyd").fit(Y)
    np.testing.assert_allclose(centers, km.cluster_centers_, atol=1e-8)
    labels, new_centers, n_unique, inertia, _ = okm.cluster_microstates(Y, 8, init, tolerance=1e-12)
    assert n_unique == 8 and np.array_equal(labels, km.labels_)
    np.testing.assert_allclose(inertia, km.inertia_, rtol=1e-9)


def test_blobs_give_requested_unique_labels():
    # tests/unit/markov_state_model/test_cluster_micro.py:69-78
    rng = np.random.default_rng(42)
    centers = np.array([[i * 10.0, j * 10.0] for i in range(4) for j in range(2)])
    Y = np.vstack([c + 0.1 * rng.standard_normal((50, 2)) for c in centers])
    labels, _, n_unique, _, _ = okm.cluster_microstates(Y, 8, centers + 0.5)
    assert n_unique == 8 and len(np.unique(labels)) == 8


# --------------------------------------------------------------------- MSM
def test_two_state_chain_timescale():
    # tests/unit/markov_state_model/test_two_state_msm.py:6-22
    rng = np.random.default_rng(0)
    p, n = 0.1, 200000
    flips = rng.random(n) < p
    d = np.cumsum(flips) % 2
    T, pi = omsm.build_simple_msm([d], n_states=2, lag=1)
    ev = omsm.eigenvalues_rev(T, pi)
    t2 = omsm.safe_timescales(1, ev[1:])[0]
    assert abs(t2 - (-1.0 / np.log(0.8))) / (-1.0 / np.log(0.8)) < 0.1


def test_mle_invariants_and_fullmatrix_agreement():
    rng = np.random.default_rng(2)
    C = rng.integers(0, 40, (9, 9)).astype(float)
    Ca, active = omsm.ensure_connected_counts(C)
    T, pi, it = omsm.mle_rev(Ca, 1e-14)
    T2, pi2, it2 = omsm.mle_rev_fullmatrix(Ca, 1e-14)
    assert it == it2
    np.testing.assert_allclose(T, T2, atol=1e-14)
    np.testing.assert_allclose(T.sum(axis=1), 1.0, atol=1e-14)
    F = pi[:, None] * T
    np.testing.assert_allclose(F, F.T, atol=1e-15)
    np.testing.assert_allclose(pi @ T, pi, atol=1e-15)
    omsm.check_transition_matrix(T, pi)
    np.testing.assert_allclose(omsm.stationary_distribution(T), pi, atol=1e-10)


def test_reversible_chain_is_its_own_mle_limit():
    # a reversible T sampled heavily: MLE(T) -> T
    rng = np.random.default_rng(3)
    A = rng.random((5, 5)); A = A + A.T
    T0 = A / A.sum(axis=1, keepdims=True)
    pi0 = A.sum(axis=1) / A.sum()
    C = 1e9 * pi0[:, None] * T0
    T, pi, _ = omsm.mle_rev(C, 1e-14)
    np.testing.assert_allclose(T, T0, atol=1e-12)
    np.testing.assert_allclose(pi, pi0, atol=1e-12)


def test_unvisited_state_is_identity_row():
    # tests/unit/markov_state_model/test_markov_state_model.py:18-82
    d = np.array([0, 1, 0, 1, 1, 0, -1, 0, 1])
    T, pi = omsm.build_simple_msm([d], n_states=3, lag=1)
    assert T.shape == (3, 3) and T[2, 2] == 1.0 and pi[2] == 0.0
    np.testing.assert_allclose(T.sum(axis=1), 1.0, atol=1e-12)
    np.testing.assert_allclose(pi @ T, pi, atol=1e-10)
    assert omsm.build_simple_msm([], None, 1)[0].shape == (0, 0)


def test_its_sweep_two_state():
    rng = np.random.default_rng(4)
    d = np.cumsum(rng.random(100000) < 0.05) % 2
    ts = omsm.its_rev_mle([d], 2, [1, 2, 5], 3)
    assert ts.shape == (3, 3) and np.all(np.isnan(ts[:, 1:]))
    np.testing.assert_allclose(ts[:, 0], -1.0 / np.log(0.9), rtol=0.1)


# --------------------------------------------------------------- featurize
def test_dihedral_known_answers():
    # planar cis = 0, planar trans = pi, +-90 degrees (IUPAC sign convention)
    base = np.array([[1.0, 0, 0], [0, 0, 0], [0, 0, 1.0]])
    def frame(p4):
        return np.vstack([base, p4]).astype(np.float32)[None]
    q = np.array([[0, 1, 2, 3]])
    assert abs(ofeat.compute_dihedrals(frame([1, 0, 1]), q)[0, 0]) < 1e-7
    assert abs(abs(ofeat.compute_dihedrals(frame([-1, 0, 1]), q)[0, 0]) - np.pi) < 1e-7
    a = ofeat.compute_dihedrals(frame([0, 1, 1]), q)[0, 0]
    b = ofeat.compute_dihedrals(frame([0, -1, 1]), q)[0, 0]
    assert abs(abs(a) - np.pi / 2) < 1e-7 and abs(a + b) < 1e-7
    # rotation about the central bond by +60 degrees
    ang = np.deg2rad(60.0)
    p4 = [np.cos(ang), np.sin(ang), 1.0]
    got = ofeat.compute_dihedrals(frame(p4), q)[0, 0]
    assert abs(abs(got) - ang) < 1e-6


def test_topology_indices(topologies):
    a = topologies["ala2"]
    phi = ofeat.dihedral_quads(a["names"], a["resid"], a["chain"], "phi")
    psi = ofeat.dihedral_quads(a["names"], a["resid"], a["chain"], "psi")
    assert phi.tolist() == [[4, 6, 8, 14]] and psi.tolist() == [[6, 8, 14, 16]]
    c = topologies["chig"]
    assert ofeat.dihedral_quads(c["names"], c["resid"], c["chain"], "phi").shape == (9, 4)
    assert ofeat.dihedral_quads(c["names"], c["resid"], c["chain"], "psi").shape == (9, 4)
    ca = ofeat.ca_indices(c["names"])
    assert len(ca) == 10 and ofeat.ca_pairs_all(ca).shape == (45, 2)
    X = ofeat.featurize_trajectory(c["xyz"][None], c["names"], c["resid"], c["chain"], "ca_distances")
    assert X.shape == (1, 45) and np.all(X > 0.3) and np.all(X < 3.0)
    # the PDB's ala2 is built planar: phi = psi = 180 degrees
    P = ofeat.featurize_trajectory(a["xyz"][None], a["names"], a["resid"], a["chain"], "phi_psi")
    np.testing.assert_allclose(np.abs(P), np.pi, atol=2e-3)


def test_trig_expand_mapping():
    # tests/unit/features/test_trig_expand_mapping.py:6-17
    X = np.array([[0.0, 1.0, np.pi / 2], [np.pi, 2.0, 0.0]])
    Xe, mapping = ofeat.trig_expand_periodic(X, np.array([True, False, True]))
    assert mapping.tolist() == [0, 0, 1, 2, 2]
    np.testing.assert_allclose(Xe[:, 0], np.cos(X[:, 0]))
    np.testing.assert_allclose(Xe[:, 1], np.sin(X[:, 0]))
    np.testing.assert_allclose(Xe[:, 2], X[:, 1])
    with pytest.raises(ValueError):
        ofeat.trig_expand_periodic(X, np.array([True]))
    w = ofeat.wrap_to_minus_pi_pi(np.array([-np.pi, np.pi, 3 * np.pi, 0.5]))
    np.testing.assert_allclose(w, [np.pi, np.pi, np.pi, 0.5])


# ------------------------------------------------------------------ Chapman-Kolmogorov test (8f-1)
def test_ck_oracle_matches_reference(golden):
    from oracle import ck as ock
    from tests import parity

    z = golden("ck")
    n = 0
    for name, dtrajs, kw in parity.ck_cases(z):
        parity.check_ck_case(z, name, dtrajs, kw, ock.run_ck, ock.compute_ck_test_micro, ock.select_lag_time_ck)
        n += 1
    assert n == 8


def test_ck_oracle_reference_test_cases():
    """tests/unit/markov_state_model/test_ck_fallback.py:14-38 and test_ck_tau_selection.py:14-19."""
    from oracle import ck as ock

    cyc = np.array([0, 1, 2] * 1000, dtype=int)
    r = ock.run_ck([cyc], lag_time=1, macro_k=3, min_trans=5, top_n_micro=3)
    assert sorted(r.mse) == [2, 3, 4, 5] and r.mse[5] <= r.mse[2] + 1e-12
    r = ock.run_ck([np.array([0, 1, 0, 1])], lag_time=1, macro_k=2, min_trans=50, top_n_micro=2)
    assert not r.mse and set(r.insufficient_k) == {2, 3, 4, 5} and r.max_error == float("inf")
    sel, *_ = ock.select_lag_time_ck([np.array([0, 0, 1, 1] * 500)], 2, [1, 2, 3])
    assert sel == 2
    m = ock.compute_ck_test_micro([cyc], 3, 1)
    assert sorted(m.mse) == [2, 3, 4, 5] and m.mse[2] < 1e-6 and not m.insufficient_data
    with pytest.raises(ValueError):
        ock.run_ck([], lag_time=1)
    with pytest.raises(ValueError):
        ock.run_ck([cyc], lag_time=0)
    # a macro lumping supplied by the caller takes the macro branch (ck_runner.py:182-219)
    rng = np.random.default_rng(0)
    blocks = np.repeat(np.arange(3), 2)
    s = np.empty(30000, dtype=int); s[0] = 0
    for t in range(1, s.size):
        u = rng.random()
        s[t] = s[t - 1] if u < 0.6 else (rng.choice(np.flatnonzero(blocks == blocks[s[t - 1]])) if u < 0.995
                                          else rng.integers(0, 6))
    r = ock.run_ck([s], lag_time=5, macro_k=3, min_trans=20, macro_lumper=lambda T, k: blocks)
    assert r.mode == "macro" and sorted(r.mse) == [2, 3, 4, 5]


def test_ck_selector_oracle_matches_reference(golden):
    """oracle.ck.select_optimal_lag_ck_its against the reference's ck_its_selector.py (control flow, CK errors,
    coverage / median-count / diagonal-mass guardrails, selection rule; see make_golden._load_reference_selector)."""
    from oracle import ck as ock
    from tests import parity

    z = golden("ck_selector")
    n = 0
    for name, dtrajs, kw, lumper in parity.selector_cases(z):
        parity.check_selector_case(z, name, dtrajs, kw, lumper, ock.select_optimal_lag_ck_its, mle_rtol=1e-12)
        n += 1
    assert n == 7
    with pytest.raises(ValueError, match="No discrete trajectories"):
        ock.select_optimal_lag_ck_its([])
    with pytest.raises(ValueError, match="exceed the available trajectory length"):
        ock.select_optimal_lag_ck_its([np.array([0, 1, 0, 1])], tau_candidates=[10])


def test_oracle_pcca_known_answers():
    """oracle.pcca (PARITY UNPINNED against deeptime: restated from Roeblitz & Weber 2013 / the msmtools form):
    nearly uncoupled blocks give indicator-like memberships; for m = 2 the membership is an affine function of
    the second eigenvector reaching 0 and 1 at its extremes; rows sum to one."""
    from oracle import pcca

    rng = np.random.default_rng(2)
    K, nb = 15, 3
    C = np.zeros((K, K))
    for b in range(nb):
        C[5 * b:5 * b + 5, 5 * b:5 * b + 5] = rng.integers(30, 90, size=(5, 5))
    C = C + C.T
    C[4, 5] = C[5, 4] = 1.0
    C[9, 10] = C[10, 9] = 1.0
    C[0, 14] = C[14, 0] = 1.0
    T = C / C.sum(axis=1, keepdims=True)
    pi = C.sum(axis=1) / C.sum()
    chi = pcca.pcca_memberships(T, 3, pi)
    np.testing.assert_allclose(chi.sum(axis=1), 1.0, atol=1e-12)
    hard = chi.argmax(axis=1)
    assert [len(set(hard[5 * b:5 * b + 5])) for b in range(nb)] == [1, 1, 1] and len(set(hard)) == 3
    assert chi.max(axis=1).min() > 0.95
    # m = 2 on a birth-death chain: chi[:, j] = (r2 - min r2) / (max r2 - min r2) or its complement
    n = 9
    P = np.zeros((n, n))
    for i in range(n):
        up = 0.3 if i < n - 1 else 0.0
        dn = 0.3 if i > 0 else 0.0
        if i == 4:
            up, dn = 0.02, 0.02
        if i < n - 1:
            P[i, i + 1] = up
        if i > 0:
            P[i, i - 1] = dn
        P[i, i] = 1.0 - up - dn
    # make it reversible: birth-death chains are; stationary distribution from detailed balance
    pi2 = np.ones(n)
    for i in range(1, n):
        pi2[i] = pi2[i - 1] * P[i - 1, i] / P[i, i - 1]
    pi2 /= pi2.sum()
    chi2 = pcca.pcca_memberships(P, 2, pi2)
    R = pcca.right_eigenvectors(P, pi2, 2)
    r2 = R[:, 1]
    aff = (r2 - r2.min()) / (r2.max() - r2.min())
    err = min(np.abs(chi2[:, 0] - aff).max(), np.abs(chi2[:, 1] - aff).max())
    assert err < 5e-3, err
