"""Chained parity on the shapes of BASELINE.json's configs C1, C2, C3 and C5.

Unlike ``test_gpu_parity.py`` (stage by stage, each stage fed the oracle-checked output of the previous
one), these tests feed COORDINATES (or raw 2-D positions) in and compare what comes out of the whole
device chain with ``oracle.pipeline.run_chain`` run on the same inputs, i.e. across the fp32 featurize
boundary.  Bars (north_star): counts bit-exact given the labels; T, pi, eigenvalues within 1e-6 of the
oracle given the same counts; across the whole chain (where a frame on a Voronoi boundary may flip
because Y is fp32 on the device and the features differ in the last fp32 bit) label mismatch <= 1e-4 of
the frames and implied timescales within 1e-3.
"""

from __future__ import annotations

import numpy as np
import pytest
import torch

import oracle
from oracle import cext
from pmarlo_b200 import kernels
from pmarlo_b200.features import (ca_pairs_all, featurize_device, plan_concat, plan_distances, plan_phi_psi_block,
                                  plan_phi_psi_interleaved)
from pmarlo_b200.msm import implied_timescales, msm_from_counts_device
from pmarlo_b200.pipeline import PipelineConfig, run_pipeline
from pmarlo_b200.shards import Segments
from pmarlo_b200.topology import Topology
from tests import synth
from tests.parity import REL, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from pmarlo_b200 import load_library

    load_library()
    torch.cuda.set_device(0)
    yield
    torch.cuda.synchronize()


def _topology(t: dict) -> Topology:
    return Topology(list(t["names"]), np.asarray(t["resid"], dtype=np.int64), np.asarray(t["chain"], dtype=np.int64))


def _chain_checks(res, ref, segs: Segments, cfg: PipelineConfig, *, label_frac=1e-4, ts_rel=1e-3, tica=True,
                  lloyd_on_device_Y=None):
    """GPU pipeline result vs oracle chain on the same inputs."""
    rep = {}
    n = segs.n_frames
    K = cfg.n_states
    if tica:
        m = res.tica
        rep["C00_rel"] = rel_err(m.C00.cpu().numpy(), ref.tica.C00)
        rep["C0t_rel"] = rel_err(m.C0t.cpu().numpy(), ref.tica.C0t)
        k = min(cfg.tica_dim, ref.tica.rank)
        rep["tica_eval_rel"] = float(np.max(np.abs(m.eigenvalues.cpu().numpy()[:k] - ref.tica.eigenvalues[:k]))
                                     / np.max(np.abs(ref.tica.eigenvalues)))
        # across the featurize boundary: fp32 features differ from the oracle's in the last bit or two
        assert rep["C00_rel"] <= 5e-6 and rep["C0t_rel"] <= 5e-6 and rep["tica_eval_rel"] <= 5e-6, rep
    lab = res.labels.cpu().numpy().astype(np.int64)
    # (1) labels are the exact fp64 argmin of the device's own Y and final centres
    Yd = res.Y.cpu().numpy().astype(np.float64)
    lab_o, _ = oracle.kmeans.assign(Yd, res.centers.cpu().numpy())
    assert np.array_equal(lab, lab_o), f"{np.count_nonzero(lab != lab_o)} labels differ from the fp64 argmin"
    # (2) counts bit-exact given those labels
    C = res.counts.cpu().numpy()
    Co = oracle.counts.count_lagged(segs.split(lab), K, cfg.msm_lag)
    assert np.array_equal(C, Co), "count matrix differs"
    assert int(C.sum()) == segs.n_pairs(cfg.msm_lag)
    # (3) T, pi, eigenvalues within 1e-6 given the same counts
    Ca, active = oracle.msm.ensure_connected_counts(Co.astype(float), alpha=cfg.dirichlet_alpha)
    Ta, pia, it_o = cext.mle_rev(Ca, maxerr=cfg.mle_maxerr)
    To, pio = oracle.msm.expand_results(K, active, Ta, pia)
    rep["T_rel"] = rel_err(res.T.cpu().numpy(), To)
    rep["pi_rel"] = rel_err(res.pi.cpu().numpy(), pio)
    rep["mle_iters"] = (int(res.mle_info[0].item()), it_o)
    assert rep["T_rel"] <= REL and rep["pi_rel"] <= REL, rep
    kk = min(cfg.n_timescales + 1, Ta.shape[0])
    evo = oracle.msm.eigenvalues_rev(Ta, pia, kk)
    ev = res.eigenvalues.cpu().numpy()
    if active.size == K:
        rep["eig_rel"] = float(np.max(np.abs(ev[:kk] - evo) / np.maximum(np.abs(evo), 1e-3)))
        assert rep["eig_rel"] <= REL, (rep, ev, evo)
    # (4) the oracle's exact fp64 Lloyd run on the DEVICE's projected coordinates from the same initial
    # frames: identical Y => the same assignment at every iteration => identical final labels
    if lloyd_on_device_Y is not None:
        rows, n_it, tol = lloyd_on_device_Y
        c = Yd[np.asarray(rows)].copy()
        if tol is None:
            for _ in range(n_it):
                lo, _ = oracle.kmeans.assign(Yd, c)
                c, _ = oracle.kmeans._update(Yd, lo, c)
        else:
            c, _, _, _ = oracle.kmeans.lloyd(Yd, c, n_it, tol)
        lo, _ = oracle.kmeans.assign(Yd, c)
        rep["lloyd_same_Y_label_mismatch"] = int(np.count_nonzero(lo != lab))
        rep["lloyd_same_Y_centers_rel"] = rel_err(res.centers.cpu().numpy(), c)
        assert rep["lloyd_same_Y_label_mismatch"] == 0 and rep["lloyd_same_Y_centers_rel"] <= 1e-9, rep
    # (5) the whole chain against the oracle's own chain (its own fp32 features, fp64 TICA, exact Lloyd).
    # An unconverged Lloyd run amplifies the last-bit differences of the fp32 features (a frame that changes
    # side moves two centres, which moves more frames at the next iteration), so this bound is statistical.
    rep["label_mismatch_frac"] = float(np.count_nonzero(lab != ref.labels)) / n
    rep["centers_rel"] = rel_err(res.centers.cpu().numpy(), ref.centers)
    ts = np.asarray(res.timescales, dtype=float)
    ok = np.isfinite(ref.timescales)
    rep["ts_rel"] = float(np.max(np.abs(ts[: ok.size][ok] - ref.timescales[ok]) / np.abs(ref.timescales[ok]))) if ok.any() else 0.0
    assert rep["label_mismatch_frac"] <= label_frac, rep
    assert np.array_equal(np.isfinite(ts[: ok.size]), ok), (ts, ref.timescales)
    if ts_rel is not None:
        assert rep["ts_rel"] <= ts_rel, rep
    return rep


# ----------------------------------------------------------------------------- C1: alanine dipeptide
def test_config_c1_alanine_dipeptide_chain(topologies):
    """C1: 35 trajectories x 371-372 frames (13 000) of the 22-atom alanine dipeptide, interleaved cos/sin
    of (phi, psi) -> F = 4, z-score, TICA lag 10 -> 2, K = 100 from fixed initial frames, MSM lag 10."""
    t = topologies["ala2"]
    top = _topology(t)
    trajs, _ = synth.ala2_trajectories(t, n_traj=35, seed=1)
    lengths = [x.shape[0] for x in trajs]
    assert len(trajs) == 35 and sum(lengths) == 13_000 and set(lengths) == {371, 372}
    plan = plan_phi_psi_interleaved(top)
    assert plan.n_cols == 4
    segs = Segments.from_lengths(lengths)
    xyz = torch.from_numpy(np.concatenate(trajs, axis=0)).cuda()
    rows = np.sort(np.random.default_rng(11).choice(lengths[0] * 8, size=100, replace=False))
    cfg = PipelineConfig(tica_lag=10, tica_dim=2, preprocess="standard", n_states=100, kmeans_max_iter=500,
                         kmeans_tolerance=1e-5, msm_lag=10, n_timescales=5)
    res = run_pipeline(xyz, segs, plan, cfg, initial_center_rows=rows)
    # oracle chain from the same coordinates: fp32 features like mdtraj, interleaved expansion (api/features.py:138-180)
    feats = []
    for x in trajs:
        ang = oracle.featurize.featurize_trajectory(x, t["names"], t["resid"], t["chain"], "phi_psi")
        Xe, _ = oracle.featurize.trig_expand_periodic(ang.astype(np.float64), np.ones(ang.shape[1], dtype=bool))
        feats.append(Xe.astype(np.float32))
    got = res.features.cpu().numpy()
    assert float(np.max(np.abs(got - np.concatenate(feats)))) <= 1e-4
    ref = oracle.pipeline.run_chain(feats, preprocess="standard", tica_lag=10, tica_dim=2, n_states=100, init_rows=rows,
                                    kmeans_iters=500, kmeans_tolerance=1e-5, msm_lag=10, n_timescales=5)
    assert res.kmeans_iters == ref.kmeans_iters, (res.kmeans_iters, ref.kmeans_iters)
    rep = _chain_checks(res, ref, segs, cfg, label_frac=2.0 / 13_000 + 1e-4, lloyd_on_device_Y=(rows, 500, 1e-5))
    print("C1", rep)


# ----------------------------------------------------------------------------- C2: Mueller-Brown, lags 1..100
def test_config_c2_muller_brown_its_sweep():
    """C2: 8 x 125 000 frames of the 2-D Mueller-Brown walk (no TICA), K = 200, reversible MSM at every lag
    1..100: all 100 count matrices bit-exact; timescales <= 1e-6 of the oracle on a ladder of lags; every
    lag's (T, pi) is a fixed point of the reversible-MLE update and satisfies detailed balance."""
    trajs = synth.muller_brown_trajectories(8, 125_000, 20260518)
    lengths = [x.shape[0] for x in trajs]
    segs = Segments.from_lengths(lengths)
    X = torch.from_numpy(np.concatenate(trajs, axis=0)).cuda()
    K = 200
    rows = np.sort(np.random.default_rng(2).choice(125_000, size=K, replace=False))
    cfg = PipelineConfig(tica_dim=0, n_states=K, kmeans_max_iter=10, kmeans_tolerance=None, msm_lag=10, n_timescales=5)
    res = run_pipeline(None, segs, None, cfg, features=X, initial_center_rows=rows)
    ref = oracle.pipeline.run_chain(trajs, tica_dim=0, n_states=K, init_rows=rows, kmeans_iters=10,
                                    kmeans_tolerance=None, msm_lag=10, n_timescales=5)
    rep = _chain_checks(res, ref, segs, cfg, tica=False, lloyd_on_device_Y=None)   # C2 has no fp32 boundary: (5) is exact
    print("C2 chain", rep)
    # ---- the ITS sweep on the device labels
    lab = res.labels.cpu().numpy().astype(np.int64)
    dtrajs = segs.split(lab)
    lags = list(range(1, 101))
    its = implied_timescales(dtrajs, lags, n_states=K, n_timescales=5)
    assert np.all(its.active_sizes >= 2)
    # every count matrix, bit-exact (C oracle == numpy oracle is pinned in the CPU suite)
    dev = torch.device("cuda")
    labels_d, off = torch.from_numpy(lab.astype(np.int32)).to(dev), segs.device(dev)
    for lag in lags:
        Cd = kernels.count_lagged(labels_d, off, K, lag).cpu().numpy()
        assert np.array_equal(Cd, cext.count_lagged(dtrajs, K, lag)), f"counts differ at lag {lag}"
    # oracle timescales on a ladder of lags (the reversible MLE needs ~1e5 iterations per lag here)
    ladder = [1, 2, 3, 5, 8, 10, 15, 20, 30, 40, 50, 60, 70, 80, 90, 100]
    ts_o, _, it_o = cext.its_rev_mle(dtrajs, K, ladder, 5)
    sel = np.array(ladder) - 1
    ok = np.isfinite(ts_o)
    assert np.array_equal(np.isfinite(its.timescales[sel]), ok)
    err = float(np.max(np.abs(its.timescales[sel][ok] - ts_o[ok]) / np.abs(ts_o[ok])))
    assert err <= REL, (err, its.iterations[sel], it_o)
    assert np.array_equal(its.iterations[sel], it_o), (its.iterations[sel], it_o)
    # timescales are positive and ordered at every lag
    t = its.timescales
    assert np.all(t[np.isfinite(t)] > 0)
    for row in t:                      # eigenvalues come sorted by magnitude: the finite timescales descend
        f = row[np.isfinite(row)]
        assert np.all(np.diff(f) <= 1e-9 * f[:-1]) if f.size > 1 else True


# ----------------------------------------------------------------------------- C3: chignolin
def test_config_c3_chignolin_chain(topologies):
    """C3 shape: chignolin (138 atoms) -> 45 CA distances + block cos/sin of 9 phi + 9 psi (36) = 81
    features -> z-score -> TICA lag 10 -> 10 dims -> K = 500 -> MSM lag 10; 16 trajectories x 2 500 frames
    here (the 2 M-frame size is covered by test_config_c3_full_size_properties)."""
    t = topologies["chig"]
    top = _topology(t)
    trajs = synth.structure_trajectories(t["xyz"], 16, 2500, seed=3, rho=0.999, sigma=0.03)
    lengths = [x.shape[0] for x in trajs]
    segs = Segments.from_lengths(lengths)
    ca = top.select_name("CA")
    pairs = ca_pairs_all(ca)
    plan = plan_concat([plan_distances(pairs), plan_phi_psi_block(top)])
    assert pairs.shape[0] == 45 and plan.n_cols == 81, (pairs.shape, plan.n_cols)
    xyz = torch.from_numpy(np.concatenate(trajs, axis=0)).cuda()
    K = 500
    rows = np.sort(np.random.default_rng(3).choice(2500 * 4, size=K, replace=False))
    cfg = PipelineConfig(tica_lag=10, tica_dim=10, preprocess="standard", n_states=K, kmeans_max_iter=6,
                         kmeans_tolerance=None, msm_lag=10, n_timescales=5)
    res = run_pipeline(xyz, segs, plan, cfg, initial_center_rows=rows)
    names, resid, chain = t["names"], t["resid"], t["chain"]
    feats = []
    for x in trajs:
        dist = oracle.featurize.compute_distances(x, pairs)
        blk = oracle.featurize.phi_psi_block_features(x, names, resid, chain)
        feats.append(np.hstack([dist, blk]).astype(np.float32))
    got = res.features.cpu().numpy()
    ref_f = np.concatenate(feats)
    assert rel_err(got[:, :45], ref_f[:, :45]) <= 2e-6 and float(np.max(np.abs(got[:, 45:] - ref_f[:, 45:]))) <= 1e-4
    ref = oracle.pipeline.run_chain(feats, preprocess="standard", tica_lag=10, tica_dim=10, n_states=K, init_rows=rows,
                                    kmeans_iters=6, kmeans_tolerance=None, msm_lag=10, n_timescales=5)
    # 40 000 frames over 500 states (80 frames per state) with an unconverged Lloyd run: the timescales of the two
    # chains are dominated by which frames changed state, so only the label statistics are bounded here; the
    # timescales are compared where the labels are identical (item 3 of _chain_checks: same counts -> 1e-6)
    rep = _chain_checks(res, ref, segs, cfg, label_frac=5e-2, ts_rel=None, lloyd_on_device_Y=(rows, 6, None))
    print("C3", rep)


def test_config_c3_full_size_properties(topologies):
    """C3 at BASELINE size (2 M frames x 138 atoms = 3.3 GB of coordinates): size-independent properties."""
    import bench

    t = topologies["chig"]
    top = _topology(t)
    ca = top.select_name("CA")
    plan = plan_concat([plan_distances(ca_pairs_all(ca)), plan_phi_psi_block(top)])
    n_traj, fpt, K = 16, 125_000, 500
    xyz = bench.synth_xyz_device(n_traj, fpt, torch.device("cuda"), seed=3, rho=0.999, sigma=0.03,
                                 base=torch.from_numpy(np.asarray(t["xyz"], dtype=np.float32)))
    segs = Segments.from_lengths([fpt] * n_traj)
    cfg = PipelineConfig(tica_lag=10, tica_dim=10, preprocess="standard", n_states=K, kmeans_max_iter=10,
                         kmeans_tolerance=None, msm_lag=10, n_timescales=5, seed=3)
    res = run_pipeline(xyz, segs, plan, cfg)
    n = n_traj * fpt
    C = res.counts.cpu().numpy()
    lab = res.labels.cpu().numpy()
    assert int(C.sum()) == segs.n_pairs(10) and lab.min() >= 0 and lab.max() < K
    starts = np.ones(n, dtype=bool)
    starts.reshape(n_traj, fpt)[:, fpt - 10:] = False
    np.testing.assert_array_equal(C.sum(axis=1), np.bincount(lab[starts], minlength=K))
    T, pi = res.T.cpu().numpy(), res.pi.cpu().numpy()
    np.testing.assert_allclose(T.sum(axis=1), 1.0, atol=1e-12)
    np.testing.assert_allclose(pi @ T, pi, atol=1e-10)
    F = pi[:, None] * T
    np.testing.assert_allclose(F, F.T, atol=1e-12)
    for sl in (slice(0, 20000), slice(n - 20000, n)):
        lo, _ = oracle.kmeans.assign(res.Y[sl].cpu().numpy().astype(np.float64), res.centers.cpu().numpy())
        np.testing.assert_array_equal(lab[sl], lo)
    # TICA model: L^T C00 L = I on the retained subspace, i.e. the projected coordinates are decorrelated
    Yc = res.Y.to(torch.float64)
    Yc = Yc - Yc.mean(dim=0, keepdim=True)
    cov = (Yc.T @ Yc / n).cpu().numpy()
    lam = res.tica.eigenvalues.cpu().numpy()[:10]
    offd = cov - np.diag(np.diag(cov))
    assert float(np.max(np.abs(offd))) <= 1e-3 * float(np.max(np.diag(cov)))
    np.testing.assert_allclose(np.diag(cov), lam ** 2, rtol=5e-3)      # kinetic map: var(y_i) = lambda_i^2


# ----------------------------------------------------------------------------- C5-shaped: K = 5000
def test_config_c5_assignment_d64_k5000():
    """D = 64, K = 5000 nearest-centre assignment bit-exact against the fp64 direct-difference argmin,
    with the fused accumulation (sums / counts / inertia) against the oracle's."""
    rng = np.random.default_rng(5)
    K, D, n = 5000, 64, 6000
    cen = rng.normal(scale=5.0, size=(K, D))
    Y = (cen[rng.integers(0, K, size=n)] + rng.normal(size=(n, D))).astype(np.float32)
    Yd, cd = torch.from_numpy(Y).cuda(), torch.from_numpy(cen).cuda()
    sums = torch.zeros((K, D), dtype=torch.float64, device="cuda")
    cnt = torch.zeros((K,), dtype=torch.int64, device="cuda")
    inertia = torch.zeros((1,), dtype=torch.float64, device="cuda")
    lab = kernels.kmeans_assign(Yd, cd, sums=sums, counts=cnt, inertia=inertia)
    lo, dmin = oracle.kmeans.assign(Y.astype(np.float64), cen, chunk=512)
    assert np.array_equal(lab.cpu().numpy().astype(np.int64), lo)
    np.testing.assert_array_equal(cnt.cpu().numpy(), np.bincount(lo, minlength=K))
    so = np.zeros((K, D))
    np.add.at(so, lo, Y.astype(np.float64))
    np.testing.assert_allclose(sums.cpu().numpy(), so, rtol=1e-12, atol=1e-9)
    assert abs(float(inertia.item()) - float(dmin.sum())) <= 1e-6 * float(dmin.sum())
    # hints (previous labels) never change the result
    again = kernels.kmeans_assign(Yd, cd, hints=lab.clone())
    assert torch.equal(again, lab)


def _c5_counts(K=5000, n_traj=16, n_frames=250_000, seed=5):
    """Label chains with K = 5000 states: a jump process over nearby states (band +-40) with dwell, so the
    5000 x 5000 count matrix is banded like a k-means discretisation of a continuous trajectory."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_traj):
        steps = rng.integers(-40, 41, size=n_frames)
        steps[rng.random(n_frames) < 0.5] = 0
        s = np.mod(rng.integers(0, K) + np.cumsum(steps), K)
        out.append(s.astype(np.int64))
    return out


def test_config_c5_counts_mle_eig_k5000():
    """K = 5000: 5000 x 5000 counts bit-exact; reversible MLE on the non-register path follows the oracle
    iteration by iteration (same iterate after a fixed number of updates, <= 1e-6); top-20 eigenvalues of
    the reversible T against LAPACK ``eigvalsh`` of the symmetrised matrix."""
    K, lag = 5000, 20
    dtrajs = _c5_counts(K)
    dev = torch.device("cuda")
    lab = torch.from_numpy(np.concatenate(dtrajs).astype(np.int32)).to(dev)
    segs = Segments.from_lengths([d.size for d in dtrajs])
    C = kernels.count_lagged(lab, segs.device(dev), K, lag)
    Co = cext.count_lagged(dtrajs, K, lag)
    assert np.array_equal(C.cpu().numpy(), Co)
    assert int(Co.sum()) == segs.n_pairs(lag)
    # fixed number of MLE updates on both sides: the same iterate
    n_it = 25
    T, pi, info, act = msm_from_counts_device(C, maxerr=0.0, maxiter=n_it)
    Ca, active = oracle.msm.ensure_connected_counts(Co.astype(float))
    assert active.size == K
    To, pio, it_o = cext.mle_rev(Ca, maxerr=0.0, maxiter=n_it, threads=8)
    assert int(info[0].item()) == it_o == n_it
    assert rel_err(pi.cpu().numpy(), pio) <= REL
    Th = T.cpu().numpy()
    assert rel_err(Th, To) <= REL
    np.testing.assert_allclose(Th.sum(axis=1), 1.0, atol=1e-12)
    # converged run: fixed point + detailed balance + stationarity (size-independent properties)
    T2, pi2, info2, _ = msm_from_counts_device(C, maxerr=1e-10)
    assert int(info2[1].item()) == 1
    T2h, pi2h = T2.cpu().numpy(), pi2.cpu().numpy()
    F = pi2h[:, None] * T2h
    assert float(np.max(np.abs(F - F.T))) <= 1e-14
    assert float(np.max(np.abs(pi2h @ T2h - pi2h))) <= 1e-12 * float(pi2h.max()) + 1e-16
    # one oracle update applied to the device's pi moves it by no more than the tolerance
    q = Ca.sum(axis=1) / pi2h
    xn = ((Ca + Ca.T) / (q[:, None] + q[None, :])).sum(axis=1)
    xn /= xn.sum()
    assert float(np.max(np.abs(xn - pi2h) / (0.5 * (xn + pi2h)))) <= 2e-10
    # top-20 eigenvalues of the reversible T
    ev, einfo = kernels.eig_rev_topk(T2, pi2, 20)
    d = np.sqrt(pi2h)
    Ssym = (d[:, None] * T2h) / d[None, :]
    Ssym = 0.5 * (Ssym + Ssym.T)
    w = np.linalg.eigvalsh(Ssym)
    w = w[np.argsort(-np.abs(w))][:20]
    evh = ev.cpu().numpy()
    assert float(np.max(np.abs(evh - w) / np.maximum(np.abs(w), 1e-3))) <= REL, (evh, w, einfo)


# ----------------------------------------------------------------------------- non-finite rows
@pytest.mark.parametrize("D,K", [(10, 1000), (3, 17)])
def test_kmeans_nan_and_inf_rows_are_labelled_like_numpy(D, K):
    """A NaN or Inf row has no smallest distance: np.argmin returns 0, and so must both assignment paths
    (the tensor path's re-check used to start from INT_MAX and indexed the accumulators out of bounds)."""
    rng = np.random.default_rng(9)
    n = 4096
    Y = rng.normal(size=(n, D)).astype(np.float32)
    Y[5, 0] = np.nan
    Y[77, D - 1] = np.inf
    Y[300] = -np.inf
    Y[4000, 1] = np.nan
    cen = rng.normal(size=(K, D))
    for impl in (1, 2):
        sums = torch.zeros((K, D), dtype=torch.float64, device="cuda")
        cnt = torch.zeros((K,), dtype=torch.int64, device="cuda")
        inertia = torch.zeros((1,), dtype=torch.float64, device="cuda")
        lab = kernels.kmeans_assign(torch.from_numpy(Y).cuda(), torch.from_numpy(cen).cuda(), sums=sums, counts=cnt,
                                    inertia=inertia, impl=impl)
        torch.cuda.synchronize()
        with np.errstate(invalid="ignore"):
            lo, _ = oracle.kmeans.assign(Y.astype(np.float64), cen)
        got = lab.cpu().numpy().astype(np.int64)
        assert np.array_equal(got, lo), (impl, np.flatnonzero(got != lo))
        assert got[5] == 0 and got[77] == 0 and got[300] == 0 and got[4000] == 0
        assert int(cnt.sum().item()) == n
