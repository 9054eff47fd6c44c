"""Host-side logic of the product (no GPU): shard bookkeeping, atom index plans,
regularisation and timescale formulas against the oracle and the golden vectors."""

from __future__ import annotations

import numpy as np
import pytest

import oracle
from pmarlo_b200 import features as F
from pmarlo_b200 import msm
from pmarlo_b200.shards import Segments, partition_trajectories
from pmarlo_b200.topology import Topology, load_pdb
from tests import synth


def test_segments_pairs_match_expected_pairs():
    rng = np.random.default_rng(0)
    for _ in range(50):
        lengths = rng.integers(0, 60, size=rng.integers(1, 6)).tolist()
        lag, step = int(rng.integers(0, 30)), int(rng.integers(1, 5))
        s = Segments.from_lengths(lengths)
        assert s.n_pairs(lag, step) == oracle.counts.expected_pairs(lengths, lag, step)
    s = Segments.from_lengths([5, 0, 7])
    assert s.n_frames == 12 and s.lengths.tolist() == [5, 0, 7]
    idx, s2 = s.drop_tail(3)
    assert idx.tolist() == [0, 1, 5, 6, 7, 8] and s2.lengths.tolist() == [2, 0, 4]
    with pytest.raises(ValueError):
        Segments.from_lengths([3, -1])


def test_partition_is_balanced_and_complete():
    lengths = [125_000] * 80
    parts = partition_trajectories(lengths, 8)
    assert sorted(i for p in parts for i in p) == list(range(80))
    assert all(len(p) == 10 for p in parts)
    ragged = [371, 372] * 17 + [371]
    parts = partition_trajectories(ragged, 4)
    loads = [sum(ragged[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= 372
    assert sorted(i for p in parts for i in p) == list(range(35))
    assert partition_trajectories([5], 3) == [[0], [], []]


def test_plans_match_oracle_indices(topologies):
    for key in ("ala2", "chig"):
        t = topologies[key]
        top = Topology(t["names"], t["resid"], t["chain"])
        for kind in ("phi", "psi"):
            np.testing.assert_array_equal(F.dihedral_quads(top, kind),
                                          oracle.featurize.dihedral_quads(t["names"], t["resid"], t["chain"], kind))
        ca = top.select_name("CA")
        np.testing.assert_array_equal(ca, oracle.featurize.ca_indices(t["names"]))
        if len(ca) >= 2:
            np.testing.assert_array_equal(F.ca_pairs_all(ca), oracle.featurize.ca_pairs_all(ca))
            np.testing.assert_array_equal(F.ca_pairs_stride3(ca, None), oracle.featurize.ca_pairs_stride3(ca, None))
            np.testing.assert_array_equal(F.ca_pairs_stride3(ca, 2), oracle.featurize.ca_pairs_stride3(ca, 2))


def test_plan_layouts():
    top = synth.backbone_topology(5)
    ang = F.plan_phi_psi(top)
    assert ang.n_cols == 8 and ang.periodic.all() and ang.columns[0] == "phi:res1" and ang.columns[4] == "psi:res0"
    blk = F.plan_phi_psi_block(top)
    # [cos phi | sin phi | cos psi | sin psi]
    assert blk.n_cols == 16
    assert blk.units[0, 6] == 0 and blk.units[0, 7] == 4 and blk.units[4, 6] == 8 and blk.units[4, 7] == 12
    inter = F.plan_phi_psi_interleaved(top)
    assert inter.units[1, 6] == 2 and inter.units[1, 7] == 3
    both = F.plan_concat([blk, F.plan_distances(np.array([[1, 4], [1, 7]]))])
    assert both.n_cols == 18 and both.units[-1, 5] == 17 and both.units[-1, 6] == -1
    assert F.parse_feature_spec("phi_psi") == ("phi_psi", {})
    assert F.parse_feature_spec("distance([3, 9])") == ("distance", {"indices": [3, 9]})
    assert F.parse_feature_spec("dist:atompair(2,5)") == ("distance_pair", {"i": 2, "j": 5})


def test_load_pdb_roundtrip(tmp_path):
    lines = []
    top = synth.backbone_topology(3)
    xyz = synth.backbone_trajectories(3, 1, 2, seed=0)[0]
    for m in range(2):
        lines.append(f"MODEL     {m + 1}")
        for i, n in enumerate(top.names):
            x, y, z = xyz[m, i] * 10.0
            lines.append(f"ATOM  {i + 1:5d} {n:<4s} ALA A{top.resid[i] + 1:4d}    {x:8.3f}{y:8.3f}{z:8.3f}  1.00  0.00")
        lines.append("ENDMDL")
    p = tmp_path / "t.pdb"
    p.write_text("\n".join(lines) + "\n")
    traj = load_pdb(str(p))
    assert traj.n_frames == 2 and traj.n_atoms == 9
    np.testing.assert_allclose(traj.xyz, xyz, atol=1e-4)
    np.testing.assert_array_equal(traj.topology.resid, top.resid)


def test_ensure_connected_counts_and_expand():
    C = np.array([[5, 1, 0, 0], [2, 7, 0, 0], [0, 0, 0, 0], [0, 1, 0, 3]], dtype=float)
    res = msm.ensure_connected_counts(C)
    Co, act = oracle.msm.ensure_connected_counts(C)
    np.testing.assert_array_equal(res.active, act)
    np.testing.assert_array_equal(res.counts, Co)
    assert res.active.tolist() == [0, 1, 3] and res.counts[0, 2] == pytest.approx(1e-3)
    empty = msm.ensure_connected_counts(np.zeros((3, 3)))
    assert empty.counts.shape == (0, 0) and empty.active.size == 0
    with pytest.raises(ValueError):
        msm.ensure_connected_counts(np.zeros((2, 3)))


def test_safe_timescales_golden(golden):
    z = golden("timescales")
    for lag in (1, 10, 400):
        np.testing.assert_array_equal(msm.safe_timescales(lag, z["ev_real"]), z[f"ts_real_{lag}"])
        np.testing.assert_array_equal(msm.safe_timescales(lag, z["ev_cplx"]), z[f"ts_cplx_{lag}"])
    assert msm.safe_timescales(5, np.array([])).shape == (0,)


def test_check_transition_matrix_and_split():
    T = np.array([[0.9, 0.1], [0.2, 0.8]])
    pi = np.array([2 / 3, 1 / 3])
    msm.check_transition_matrix(T, pi)
    with pytest.raises(ValueError):
        msm.check_transition_matrix(np.array([[1.0 + 1e-13, -1e-13], [0.2, 0.8]]), pi)
    with pytest.raises(ValueError):
        msm.check_transition_matrix(T, np.array([0.5, 0.5]))
    d = [np.array([0, 1, -1, 2, 2, 7, 1]), np.array([-1, -1]), np.array([], dtype=int)]
    got = msm.split_at_invalid(d, 3)
    ref = oracle.counts.split_at_invalid(d, 3)
    assert len(got) == len(ref) == 3
    for a, b in zip(got, ref):
        np.testing.assert_array_equal(a, b)
    assert msm.infer_n_states(d) == 8 and msm.infer_n_states(d, 4) == 4 and msm.infer_n_states([np.array([-1])]) == 0
    lcs = msm.largest_connected_set(np.array([[1, 1, 0], [1, 1, 0], [0, 1, 1]]))
    assert lcs.tolist() == [0, 1]


def test_lloyd_host_loop_matches_oracle_semantics(monkeypatch):
    """The stopping rule of lloyd_device (deeptime's: the inertia of the UPDATED centres under the labels that
    produced them, evaluated from the iteration's sums and counts without a second assignment) reproduces
    oracle.kmeans.lloyd, which sums the squared differences directly; kernels are replaced by numpy stand-ins.
    The identity rounds at 1e-16 of the cost, so the exact-zero tolerance of the last case is 1e-13 here."""
    import torch

    from pmarlo_b200 import clustering, kernels

    rng = np.random.default_rng(1)
    Y = np.concatenate([rng.normal(size=(200, 2)) + c for c in ([0, 0], [6, 0], [0, 6])])
    c0 = Y[[0, 250, 500]].copy()

    def fake_assign(Yt, centers, labels=None, sums=None, counts=None, inertia=None, n_rechecked=None, **_ignored):
        lab, dmin = oracle.kmeans.assign(Yt.numpy(), centers.numpy())
        labels.copy_(torch.from_numpy(lab.astype(np.int32)))
        if sums is not None:
            s = np.zeros(tuple(sums.shape))
            np.add.at(s, lab, Yt.numpy())
            sums += torch.from_numpy(s)
            counts += torch.from_numpy(np.bincount(lab, minlength=counts.numel()))
            inertia += float(dmin.sum())
        return labels

    def fake_update(centers, sums, counts, shift2=None):
        nz = counts > 0
        centers[nz] = sums[nz] / counts[nz].to(torch.float64)[:, None]

    monkeypatch.setattr(kernels, "kmeans_assign", fake_assign)
    monkeypatch.setattr(kernels, "kmeans_update", fake_update)
    for max_iter, tol in ((500, 1e-5), (1, 1e-5), (3, 1e-13)):
        res = clustering.lloyd_device(torch.from_numpy(Y), torch.from_numpy(c0), max_iter=max_iter, tolerance=tol)
        co, it, cost, conv = oracle.kmeans.lloyd(Y, c0, max_iter=max_iter, tolerance=tol)
        assert res.n_iter == it and res.converged == conv
        np.testing.assert_allclose(res.centers.numpy(), co, rtol=1e-12)
        assert res.cost == pytest.approx(cost, rel=1e-12)
    res = clustering.lloyd_device(torch.from_numpy(Y), torch.from_numpy(c0), max_iter=4, tolerance=None)
    assert res.n_iter == 4 and res.cost is None


# ----------------------------------------------------------------------------- discretize host logic (CPU)
def test_expected_pairs_matches_simulation_like_the_reference_tests():
    """Restates tests/unit/analysis/test_counting.py:23-66 of the reference (brute-force simulation of the
    (t, t+tau) pairs per segment, single and per-segment strides, error behaviour)."""
    from pmarlo_b200.discretize import expected_pairs

    def simulate(lengths, tau, strides):
        total = 0
        for length, stride in zip(lengths, strides):
            for idx in range(0, max(0, length - tau), max(1, stride)):
                total += idx + tau < length
        return total

    rng = np.random.default_rng(0)
    for _ in range(300):
        lengths = [int(v) for v in rng.integers(0, 51, size=rng.integers(1, 6))]
        tau = int(rng.integers(0, 11))
        stride = int(rng.integers(1, 11))
        assert expected_pairs(lengths, tau, stride) == simulate(lengths, tau, [stride] * len(lengths))
        strides = [int(v) for v in rng.integers(1, 11, size=rng.integers(1, 6))]
        full = (strides + [strides[-1]] * len(lengths))[: len(lengths)]
        assert expected_pairs(lengths, tau, strides) == simulate(lengths, tau, full)
    with pytest.raises(ValueError, match="lengths must be non-negative"):
        expected_pairs([5, -1], tau=1)
    with pytest.raises(ValueError, match="stride values must be positive"):
        expected_pairs([5], tau=1, stride=0)
    with pytest.raises(ValueError, match="stride iterable must not be empty"):
        expected_pairs([5], tau=1, stride=[])


def test_discretize_split_and_segment_bookkeeping():
    from pmarlo_b200 import discretize as dz

    X = np.random.default_rng(1).normal(size=(50, 3))
    ds = {"splits": {"train": {"X": X, "segments": [{"length": 20}, {"start": 20, "stop": 45, "stride": 2}, 30]},
                     "bad": {"X": np.full((4, 3), np.nan)}, "val": X[:10]}}
    sp = dz._find_splits(ds)
    assert list(sp) == ["train", "val"]                       # non-finite splits are ignored
    assert dz._segments_of(sp["train"], 50) == ([20, 25, 5], [1, 2, 1])      # truncated to the frames present
    assert dz._segments_of(sp["val"], 10) == ([10], [1])
    assert dz._segments_of({"X": X, "segment_lengths": [7, 8]}, 50) == ([7, 8], [1, 1])
    assert list(dz._find_splits({"a": X, "__meta__": X})) == ["a"]
    assert list(dz._find_splits(X)) == ["all"]
    with pytest.raises(ValueError):
        dz._find_splits({"splits": {"t": np.zeros((0, 3))}})
    assert dz._schema_of({"X": X, "cv_names": ["p", "q", "r"]}, 3)["names"] == ["p", "q", "r"]
    assert dz._schema_of(X, 3)["names"] == ["feature_0", "feature_1", "feature_2"]
    with pytest.raises(ValueError):
        dz._check_schema({"names": ["a"], "n_features": 1}, {"names": ["b"], "n_features": 1}, "s")


def test_ck_host_bookkeeping_on_the_kernel_contracts(monkeypatch, golden, tmp_path):
    """pmarlo_b200.ck's control flow (state filtering, top-n selection, insufficient-pair bookkeeping,
    lag selection, side outputs) against the reference's golden vectors, with the two device kernels
    replaced by numpy statements of their contracts; the kernels themselves are covered by -m gpu."""
    import json

    import torch

    from pmarlo_b200 import ck, kernels
    from tests import fake_kernels, parity

    fake_kernels.install(monkeypatch)
    monkeypatch.setattr(kernels, "require_cuda", lambda: torch.device("cpu"))
    z = golden("ck")
    for name, dtrajs, kw in parity.ck_cases(z):
        parity.check_ck_case(z, name, dtrajs, kw, ck.run_ck, ck.compute_ck_test_micro, ck.select_lag_time_ck)
    cyc = [np.array([0, 1, 2] * 1000)]
    r = ck.run_ck(cyc, 1, tmp_path, macro_k=3, min_trans=5, top_n_micro=3)
    saved = json.loads((tmp_path / "ck_mse.json").read_text())
    assert saved["mode"] == "micro" and sorted(saved["mse"]) == ["2", "3", "4", "5"] and saved["insufficient_k"] == []
    assert (tmp_path / "ck_mse.csv").read_text().splitlines()[0] == "k,mse"
    assert ck.ck_rms_error(r) == r.max_error and r.has_valid_tests
    for bad in (dict(dtrajs=[], lag_time=1), dict(dtrajs=cyc, lag_time=0), dict(dtrajs=cyc, lag_time=1, factors=(1,))):
        with pytest.raises(ValueError):
            ck.run_ck(bad.pop("dtrajs"), bad.pop("lag_time"), None, **bad)
    # a state that is only ever the last frame has an empty row: deeptime's row normalisation raises
    with pytest.raises(ValueError, match="strictly positive row sums"):
        ck.run_ck([np.array([0, 1] * 200 + [2])], 1, None, min_trans=1)
    # macro branch with an injected lumping
    blocks = np.repeat(np.arange(3), 2)
    rng = np.random.default_rng(0)
    s = np.empty(30000, dtype=int)
    s[0] = 0
    for t in range(1, s.size):
        u = rng.random()
        s[t] = s[t - 1] if u < 0.6 else (rng.choice(np.flatnonzero(blocks == blocks[s[t - 1]])) if u < 0.995
                                          else rng.integers(0, 6))
    got = ck.run_ck([s], 5, None, macro_k=3, min_trans=20, macro_lumper=lambda T, k: blocks)
    ref = oracle.ck.run_ck([s], 5, macro_k=3, min_trans=20, macro_lumper=lambda T, k: blocks)
    assert got.mode == ref.mode == "macro"
    np.testing.assert_allclose([got.mse[k] for k in (2, 3, 4, 5)], [ref.mse[k] for k in (2, 3, 4, 5)], rtol=1e-9)


def test_ck_lag_selector_host_logic_on_the_kernel_contracts(monkeypatch, golden):
    """pmarlo_b200.ck.select_optimal_lag_ck_its against the reference's golden vectors with the device
    kernels replaced by numpy statements of their contracts (guardrails, macro / micro CK error, diagonal
    mass, selection rule and its fallbacks, error messages)."""
    import torch

    from pmarlo_b200 import ck, kernels
    from tests import fake_kernels, parity

    fake_kernels.install(monkeypatch)
    monkeypatch.setattr(kernels, "require_cuda", lambda: torch.device("cpu"))
    z = golden("ck_selector")
    for name, dtrajs, kw, lumper in parity.selector_cases(z):
        parity.check_selector_case(z, name, dtrajs, kw, lumper, ck.select_optimal_lag_ck_its, mle_rtol=1e-9)
    with pytest.raises(ValueError, match="No discrete trajectories"):
        ck.select_optimal_lag_ck_its([])
    with pytest.raises(ValueError, match="contain no frames"):
        ck.select_optimal_lag_ck_its([np.array([], dtype=int)])
    with pytest.raises(ValueError, match="exceed the available trajectory length"):
        ck.select_optimal_lag_ck_its([np.array([0, 1, 0, 1])], tau_candidates=[10])


def test_macro_helpers_match_reference_golden(monkeypatch, golden):
    """pmarlo_b200.macro against the reference's own numpy functions (tests/golden/macro.npz); the flux
    aggregation runs through torch, here on the CPU device."""
    import torch

    from pmarlo_b200 import kernels, macro

    monkeypatch.setattr(kernels, "require_cuda", lambda: torch.device("cpu"))
    z = golden("macro")
    np.testing.assert_allclose(macro.compute_macro_populations(z["pi"], z["lab"]), z["pops"], rtol=1e-13)
    Tm = macro.lump_micro_to_macro_T(z["T"], z["pi"], z["lab"])
    np.testing.assert_allclose(Tm, z["Tm"], rtol=1e-12)
    np.testing.assert_allclose(macro.compute_macro_mfpt(Tm), z["mfpt"], rtol=1e-10)
    assert macro.compute_macro_populations(np.zeros(0), np.zeros(0, dtype=int)).shape == (0,)
    assert macro.lump_micro_to_macro_T(np.zeros((0, 0)), np.zeros(0), np.zeros(0, dtype=int)).shape == (0, 0)


def test_ck_host_logic_fuzz_against_oracle(monkeypatch):
    """Seeded random label trajectories (rare states, unassigned frames, unused ids, short shards) through
    run_ck, compute_ck_test_micro, select_lag_time_ck and select_optimal_lag_ck_its: the device-side host
    logic (kernels replaced by their numpy contracts) must agree with the oracle restatement of the
    reference, outcome by outcome (mode, factors, errors raised, selected lags)."""
    import torch

    from oracle import ck as ock
    from pmarlo_b200 import ck, kernels
    from tests import fake_kernels

    fake_kernels.install(monkeypatch)
    monkeypatch.setattr(kernels, "require_cuda", lambda: torch.device("cpu"))
    rng = np.random.default_rng(321)

    def chain(K, n, stay, rare):
        P = rng.random((K, K)) ** 3 + 1e-3
        P[:, rare] *= 0.01
        P /= P.sum(1, keepdims=True)
        cum = np.cumsum(stay * np.eye(K) + (1 - stay) * P, 1)
        s = np.empty(n, dtype=np.int64)
        s[0] = rng.integers(0, K)
        u = rng.random(n)
        for t in range(1, n):
            s[t] = min(int(np.searchsorted(cum[s[t - 1]], u[t])), K - 1)
        return s

    def outcome(fn):
        try:
            return ("ok", fn())
        except Exception as e:      # the reference raises (deeptime row normalisation, eigenvalue count): compare the type
            return ("err", type(e).__name__)

    for trial in range(24):
        K = int(rng.integers(3, 20))
        rare = list(rng.choice(K, size=int(rng.integers(0, 3)), replace=False))
        dtr = [chain(K, int(rng.integers(20, 1200)), float(rng.uniform(0.3, 0.9)), rare)
               for _ in range(int(rng.integers(1, 4)))]
        if trial % 3 == 0:
            dtr[0] = dtr[0].copy()
            dtr[0][rng.random(dtr[0].size) < 0.05] = -1
        if trial % 4 == 1:
            dtr = [d + 2 for d in dtr]
        lag = int(rng.integers(1, 5))
        kw = dict(min_trans=int(rng.integers(1, 30)), top_n_micro=int(rng.integers(2, 25)), macro_k=int(rng.integers(2, 5)))
        a = outcome(lambda: ock.run_ck(dtr, lag, **kw))
        b = outcome(lambda: ck.run_ck(dtr, lag, None, **kw))
        assert a[0] == b[0], (trial, a, b)
        if a[0] == "ok":
            ra, rb = a[1], b[1]
            assert (ra.mode, sorted(ra.mse), list(ra.insufficient_k)) == (rb.mode, sorted(rb.mse), list(rb.insufficient_k)), trial
            np.testing.assert_allclose([rb.mse[k] for k in sorted(rb.mse)], [ra.mse[k] for k in sorted(ra.mse)], rtol=1e-9, atol=1e-15)
        else:
            assert a[1] == b[1], (trial, a, b)
        n_states = int(max(int(d.max()) for d in dtr)) + 1
        ma = ock.compute_ck_test_micro(dtr, n_states, lag, max_states=kw["top_n_micro"])
        mb = ck.compute_ck_test_micro(dtr, n_states, lag, max_states=kw["top_n_micro"])
        assert (ma.insufficient_data, sorted(ma.mse)) == (mb.insufficient_data, sorted(mb.mse)), trial
        np.testing.assert_allclose([mb.mse[k] for k in sorted(mb.mse)], [ma.mse[k] for k in sorted(ma.mse)], rtol=1e-9, atol=1e-15)
        sa = outcome(lambda: ock.select_lag_time_ck(dtr, n_states, [1, 2, 3])[0])
        sb = outcome(lambda: ck.select_lag_time_ck(dtr, n_states, [1, 2, 3])[0])
        assert sa == sb, (trial, sa, sb)
        if all(int(d.min()) >= 0 for d in dtr):
            oa = ock.select_optimal_lag_ck_its(dtr, [1, 2, 4], min_median_count=5, coverage_threshold=0.7, n_timescales=3)
            ob = ck.select_optimal_lag_ck_its(dtr, [1, 2, 4], min_median_count=5, coverage_threshold=0.7, n_timescales=3)
            assert oa[0] == ob[0], trial
            for ea, eb in zip(oa[1], ob[1]):
                assert (ea.lag, ea.passed_sanity, ea.median_count, ea.n_macrostates) == \
                       (eb.lag, eb.passed_sanity, eb.median_count, eb.n_macrostates), trial
                assert (ea.failure_reason or "")[:9] == (eb.failure_reason or "")[:9], (trial, ea.failure_reason, eb.failure_reason)
                if np.isfinite(ea.ck_error):
                    np.testing.assert_allclose(eb.ck_error, ea.ck_error, rtol=1e-8)


def test_numpy_models_of_two_kernel_algorithms():
    """CPU models of two algorithmic choices made inside kernels, so that the claims in DESIGN.md are checked
    without a GPU:
    (1) eig.cu: Lanczos with partial re-orthogonalisation (pairs of full CGS2 steps placed by the scalar omega
        bound of lanczos_kernel, three-term recurrence in between) gives the Ritz values of full
        re-orthogonalisation on a clustered spectrum and on a narrow-bulk matrix (K = 2000, relative bulk width
        ~1/50) where a fixed period of 8 loses orthogonality completely; without any re-orthogonalisation a
        ghost copy of the top eigenvalue appears;
    (2) tica_grid.cu: the division-free Jacobi rotation c^2 = (1 + |alpha|/r)/2, s = sign(alpha) g / (2 r c)
        orthogonalises two rows, and the carried squared norms follow a' = a - t g, b' = b + t g."""
    rng = np.random.default_rng(0)
    K = 400
    lam = np.concatenate([[1.0, 0.98994, 0.97752, 0.97577, 0.97517, 0.97441], np.sort(rng.uniform(-0.3, 0.95, K - 6))[::-1]])
    Q, _ = np.linalg.qr(rng.standard_normal((K, K)))
    S = (Q * lam) @ Q.T
    S = 0.5 * (S + S.T)

    def lanczos(S, m, mode):
        # mode: "full" | "partial" (the kernel's rule) | "period8" (the rule it replaced) | "none"
        K = S.shape[0]
        V = np.zeros((m + 1, K))
        w = 1.0 + 0.5 * np.sin(0.7548776662466927 * np.arange(1, K + 1))
        binv = 1.0 / np.linalg.norm(w)
        alpha, beta = np.zeros(m), np.zeros(m)
        omega, g_last, a_max, b_max, full_left, n_full = 1e-15, 1.0, 0.0, 0.0, 2, 0
        for j in range(m):
            V[j] = w * binv
            wn = S @ V[j]
            if mode == "partial":
                if full_left == 0 and omega * 1.5 * g_last > 1e-8:
                    full_left = 2
                full = full_left > 0
            else:
                full = mode == "full" or (mode == "period8" and (j < 2 or j % 8 >= 6))
            n_full += full
            lo = 0 if full else max(0, j - 1)
            a = 0.0
            for _ in range(2 if full else 1):
                h = V[lo:j + 1] @ wn
                wn = wn - h @ V[lo:j + 1]
                a += h[-1]
            alpha[j], beta[j] = a, np.linalg.norm(wn)
            a_max, b_max = max(a_max, abs(a)), max(b_max, beta[j])
            g_last = max(1.0, (a_max + abs(a) + 2.0 * b_max) / beta[j])
            if full:
                full_left -= 1
                if full_left == 0:
                    omega = 1e-15
            else:
                omega = min(omega * g_last, 1.0)
            w, binv = wn, 1.0 / beta[j]
        Tm = np.diag(alpha) + np.diag(beta[:-1], 1) + np.diag(beta[:-1], -1)
        ev = np.linalg.eigvalsh(Tm)
        return ev[np.argsort(-np.abs(ev))], np.abs(V[:m] @ V[:m].T - np.eye(m)).max(), n_full

    ev_full, orth_full, _ = lanczos(S, 150, "full")
    ev_par, orth_par, n_full = lanczos(S, 150, "partial")
    ev_none, orth_none, _ = lanczos(S, 150, "none")
    np.testing.assert_allclose(ev_par[:6], ev_full[:6], rtol=0, atol=1e-13)
    assert orth_full < 1e-13 and orth_par < 1e-9 and n_full < 100
    assert orth_none > 1e-3 and abs(ev_none[1] - 1.0) < 1e-6       # ghost of the Perron eigenvalue

    # narrow bulk: five metastable blocks, K = 2000 (the matrix of tools/eig_bench.py); growth per local step ~50
    Kb = 2000
    Cb = rng.random((Kb, Kb)) * 0.02
    for i in range(5):
        Cb[i * 400:(i + 1) * 400, i * 400:(i + 1) * 400] += rng.random((400, 400))
    Cb = Cb + Cb.T
    sq = np.sqrt(Cb.sum(1))
    Sb = Cb / sq[:, None] / sq[None, :]
    ref = np.linalg.eigvalsh(Sb)
    ref = ref[np.argsort(-np.abs(ref))][:10]
    ev_par, orth_par, n_full = lanczos(Sb, 200, "partial")
    np.testing.assert_allclose(ev_par[:10], ref, rtol=0, atol=1e-12)
    assert orth_par < 1e-9 and n_full <= 100
    _, orth_p8, _ = lanczos(Sb, 200, "period8")
    assert orth_p8 > 0.5                                           # what the fixed period did here

    for _ in range(200):
        x, y = rng.standard_normal(64) * 10 ** rng.uniform(-3, 3), rng.standard_normal(64) * 10 ** rng.uniform(-3, 3)
        a, b, g = x @ x, y @ y, x @ y
        alpha = 0.5 * (b - a)
        r = np.hypot(alpha, g)
        c = np.sqrt(0.5 * (1.0 + abs(alpha) / r))
        s = (1.0 if alpha >= 0 else -1.0) * g / (2.0 * r * c)
        t = s / c
        xn, yn = c * x - s * y, s * x + c * y
        assert abs(xn @ yn) <= 1e-12 * np.sqrt(a * b)
        assert abs(c * c + s * s - 1.0) < 1e-15
        np.testing.assert_allclose([xn @ xn, yn @ yn], [a - t * g, b + t * g], rtol=1e-9)


def test_numpy_model_of_the_ritz_value_multisection():
    """eig.cu: the Ritz values of the Lanczos tridiagonal come from a 32-way multisection per eigenvalue (one
    warp, 32 Sturm counts per pass) with the DIVISION-FREE Sturm count on the scaled leading principal minors
    (renormalised every eight rows).  The model restates sturm_count_minors / tridiag_bisect_extremes and is
    compared with eigvalsh on tridiagonals shaped like the kernel's (a few large leading entries, a narrow
    bulk, one nearly decoupled row), at three scales."""
    def sturm_minors(a, b, x, inv):
        xs = x * inv
        p0, p1 = 1.0, a[0] * inv - xs
        neg = p1 < 0 or p1 == 0
        cnt = int(neg)
        for i in range(1, len(a)):
            bi = b[i - 1] * inv
            pn = (a[i] * inv - xs) * p1 - (bi * bi) * p0
            p0, p1 = p1, pn
            ng = pn < 0 or (pn == 0 and not neg)
            cnt += ng != neg
            neg = ng
            if i % 8 == 0:
                mx = max(abs(p0), abs(p1))
                if mx > 2.0 ** 256:
                    p0, p1 = p0 * 2.0 ** -256, p1 * 2.0 ** -256
                elif mx < 2.0 ** -256:
                    p0, p1 = p0 * 2.0 ** 256, p1 * 2.0 ** 256
        return cnt

    def extremes(a, b, kk):
        m = len(a)
        r = np.zeros(m)
        r[1:] += np.abs(b[:m - 1])
        r[:-1] += np.abs(b[:m - 1])
        lo, hi = (a - r).min(), (a + r).max()
        span = max(abs(lo), abs(hi))
        inv = 1.0 / span
        lo -= 1e-12 * span + 1e-300
        hi += 1e-12 * span + 1e-300
        out, passes = [], 0
        for t in range(2 * kk):
            idx = t if t < kk else m - 1 - (t - kk)
            lb, hb = lo, hi
            for _ in range(16):
                w = (hb - lb) * (1.0 / 33.0)
                if not (lb + w > lb) or not (lb + 32.0 * w < hb):
                    break
                passes += 1
                above = [sturm_minors(a, b, lb + (ln + 1) * w, inv) > idx for ln in range(32)]
                if not any(above):
                    lb = lb + 32.0 * w
                else:
                    f = above.index(True)
                    hb, lb = lb + (f + 1) * w, (lb + f * w if f > 0 else lb)
            out.append(0.5 * (lb + hb))
        return np.array(out), passes / (2 * kk)

    rng = np.random.default_rng(1)
    for m, scale in ((48, 1.0), (96, 1e6), (64, 1e-9)):
        a = rng.normal(0, 0.01, m) * scale
        a[:5] = np.array([0.9, 0.8, 0.1, 0.5, -0.3]) * scale
        b = np.abs(rng.normal(0.02, 0.005, m)) * scale
        b[:4] = np.array([0.3, 0.2, 0.1, 0.05]) * scale
        b[10] = 1e-13 * scale
        ev = np.linalg.eigvalsh(np.diag(a) + np.diag(b[:m - 1], 1) + np.diag(b[:m - 1], -1))
        got, passes = extremes(a, b, 4)
        np.testing.assert_allclose(got, np.concatenate([ev[:4], ev[::-1][:4]]), rtol=0, atol=4e-15 * scale)
        assert passes <= 12


def test_numpy_model_of_the_cholesky_preconditioned_jacobi():
    """tica_grid.cu: the one-sided Jacobi sweeps for eig(C00) start from W = G^T of a pivoted Cholesky
    factorisation C00 = G G^T (chol_pivoted_cta: no interchanges, blocks of 8 pivots chosen by residual
    diagonal, a collapsed candidate deferred) instead of from W = C00.  The model restates the blocked
    factorisation and the row sweeps with the kernel's stopping rule on an ill-conditioned correlation matrix:
    G G^T = C00 to rounding, the eigenvalues |w_j|^2 match eigvalsh, and the sweep count drops."""
    rng = np.random.default_rng(0)
    d, n = 64, 4000
    x = rng.standard_normal((n, d)).cumsum(axis=0) * 0.02 + rng.standard_normal((n, d))
    x[:, d // 2:] = 0.9 * x[:, :d - d // 2] + 0.1 * x[:, d // 2:]      # nearly dependent pairs of columns
    x -= x.mean(0)
    x /= x.std(0)
    C = x.T @ x / n

    def chol_blocked(A, B=8):
        n_ = A.shape[0]
        diag = np.diag(A).copy()
        done = np.zeros(n_, dtype=bool)
        rows = []
        thresh = diag.max() * n_ * 2.220446049250313e-16
        while len(rows) < n_:
            cand_val = np.where(done, -1.0, diag)
            cand = [int(i) for i in np.argsort(-cand_val, kind="stable")[:B] if cand_val[i] > thresh]
            if not cand:
                break
            G = np.array(rows) if rows else np.zeros((0, n_))
            cols = {p: A[:, p] - G.T @ G[:, p] for p in cand}
            d0 = {p: diag[p] for p in cand}
            for t, p in enumerate(cand):
                piv = cols[p][p]
                if not (piv > thresh and piv >= 0.25 * d0[p]):
                    continue
                g = np.where(done, 0.0, cols[p] / np.sqrt(piv))
                g[p] = np.sqrt(piv)
                rows.append(g)
                diag -= g * g
                done[p] = True
                for p2 in cand[t + 1:]:
                    cols[p2] = cols[p2] - g * g[p2]
        return np.array(rows)

    def sweeps(W, tol=4.5e-16):
        W = W.copy()
        m = W.shape[0]
        tol2 = tol * tol * W.shape[1]
        for s_ in range(60):
            rot = big = 0
            for i in range(m - 1):
                for j in range(i + 1, m):
                    a, b, g = W[i] @ W[i], W[j] @ W[j], W[i] @ W[j]
                    if g * g <= tol2 * a * b:
                        continue
                    rot += 1
                    big += g * g > 1e-20 * a * b
                    zeta = (b - a) / (2 * g)
                    t = np.sign(zeta) / (abs(zeta) + np.sqrt(1 + zeta * zeta)) if zeta != 0 else 1.0
                    c = 1 / np.sqrt(1 + t * t)
                    W[i], W[j] = c * W[i] - c * t * W[j], c * t * W[i] + c * W[j]
            if rot == 0 or big == 0:
                return s_ + 1, W
        return 60, W

    G = chol_blocked(C)
    assert G.shape[0] == d
    np.testing.assert_allclose(G.T @ G, C, rtol=0, atol=1e-13)
    ref = np.sort(np.linalg.eigvalsh(C))[::-1]
    s_plain, W1 = sweeps(C)
    s_chol, W2 = sweeps(G)
    np.testing.assert_allclose(np.sort(np.linalg.norm(W1, axis=1))[::-1], ref, rtol=0, atol=1e-12)
    np.testing.assert_allclose(np.sort(np.linalg.norm(W2, axis=1) ** 2)[::-1], ref, rtol=0, atol=1e-12)
    assert s_chol + 2 <= s_plain, (s_chol, s_plain)


def test_candidate_lag_ladder_matches_reference_golden(golden):
    """utils/msm_utils.py:21-105 -- outputs of the genuine reference function (tests/golden/make_golden.py::make_ladders)."""
    from pmarlo_b200 import candidate_lag_ladder

    z = golden("ladders")
    for i, (a, b, c) in enumerate(z["ladder_args"]):
        got = candidate_lag_ladder(int(a), int(b), None if c < 0 else int(c))
        assert got == [int(v) for v in z[f"ladder_{i}"]], (a, b, c)
    assert candidate_lag_ladder() == [1, 2, 3, 5, 8, 10, 15, 20, 30, 40, 50, 75, 80, 100, 150, 160, 200]
    for bad in ((0, 10, None), (5, 4, None), (1, 10, 0), (2001, 3000, None)):
        with pytest.raises(ValueError):
            candidate_lag_ladder(*bad)
