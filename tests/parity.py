"""Stage-by-stage parity of the CUDA path (through the C ABI) against the CPU oracle.

Used by the ``-m gpu`` tests and by ``__graft_entry__.smoke()``.  Tolerances (the
bar of BASELINE.json's north_star): labels and counts bit-exact given identical
centroids; covariances, TICA eigenvalues, T, pi, eigenvalues and timescales within
1e-6 relative of the fp64 oracle; fp32 features (mdtraj computes in fp32) within
1e-4 rad / 2e-6 relative.
"""

from __future__ import annotations

import ast

import numpy as np
import torch

import oracle
from pmarlo_b200 import kernels
from pmarlo_b200.clustering import lloyd_device
from pmarlo_b200.features import featurize_device, plan_concat, plan_distances, plan_phi_psi, plan_phi_psi_block
from pmarlo_b200.features import ca_pairs_all
from pmarlo_b200.msm import msm_from_counts_device, safe_timescales
from pmarlo_b200.reduction import TICA
from pmarlo_b200.shards import Segments, concat_to_device
from tests import synth

REL = 1e-6          # north_star tolerance for fp64 quantities
ANGLE_ABS = 1e-4    # fp32 atan2 of fp32 coordinates vs fp64 oracle, radians
DIST_REL = 2e-6     # fp32 sqrt of fp32 differences


def rel_err(a, b) -> float:
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(float(np.max(np.abs(b))), 1e-300)
    return float(np.max(np.abs(a - b))) / scale


def angle_err(a, b) -> float:
    d = np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))
    return float(np.max(np.minimum(d, 2 * np.pi - d))) if d.size else 0.0


def check_featurize(trajs, top) -> dict:
    dev = torch.device("cuda")
    xyz = np.concatenate(trajs, axis=0)
    xd = torch.from_numpy(xyz).to(dev)
    ca = top.select_name("CA")
    pairs = ca_pairs_all(ca)[:40]
    plan = plan_concat([plan_phi_psi(top), plan_phi_psi_block(top), plan_distances(pairs)])
    got = featurize_device(xd, plan).cpu().numpy()
    names, resid, chain = top.names, top.resid, top.chainid
    phi = oracle.featurize.compute_dihedrals(xyz, oracle.featurize.dihedral_quads(names, resid, chain, "phi"))
    psi = oracle.featurize.compute_dihedrals(xyz, oracle.featurize.dihedral_quads(names, resid, chain, "psi"))
    ang = np.concatenate([phi, psi], axis=1)
    na = ang.shape[1]
    blk = oracle.featurize.phi_psi_block_features(xyz, names, resid, chain)
    dist = oracle.featurize.compute_distances(xyz, pairs)
    e_ang = angle_err(got[:, :na], ang)
    e_trig = float(np.max(np.abs(got[:, na:na + blk.shape[1]] - blk)))
    e_dist = rel_err(got[:, na + blk.shape[1]:], dist)
    assert e_ang <= ANGLE_ABS, f"dihedral angles off by {e_ang}"
    assert e_trig <= ANGLE_ABS, f"cos/sin off by {e_trig}"
    assert e_dist <= DIST_REL, f"distances off by {e_dist}"
    assert np.all(got[:, :na] > -np.pi - 1e-6) and np.all(got[:, :na] <= np.pi + 1e-6)
    return {"angle_abs": e_ang, "trig_abs": e_trig, "dist_rel": e_dist}


def check_tica(feats, lag: int, dim: int, preprocess, gram_impl: int = 0) -> tuple[dict, torch.Tensor, Segments]:
    dev = torch.device("cuda")
    X, segs = concat_to_device(feats, dev)
    est = TICA(lag, dim, preprocess=preprocess, gram_impl=gram_impl)
    model = est.fit_device(X, segs)
    Y = est.transform_device(model, X, out_f64=True)
    # oracle on the same float32 inputs
    flat = np.concatenate([np.asarray(f, dtype=np.float64) for f in feats], axis=0)
    if preprocess is None:
        prepped = [np.asarray(f, dtype=np.float64) for f in feats]
    else:
        Z = oracle.tica.preprocess(flat, scale=(preprocess == "standard"))
        prepped = segs.split(Z)
    om = oracle.tica.tica_fit(prepped, lag)
    rep = {
        "C00_rel": rel_err(model.C00.cpu().numpy(), om.C00),
        "C0t_rel": rel_err(model.C0t.cpu().numpy(), om.C0t),
        "mu_abs": float(np.max(np.abs(model.mu.cpu().numpy() - om.mean))),
        "n_pairs": model.n_pairs,
    }
    assert model.n_pairs == om.n_pairs
    assert model.rank == om.rank, (model.rank, om.rank)
    r = om.rank
    m = min(dim, r)
    ev = model.eigenvalues.cpu().numpy()[:r]
    # the `dim` leading eigenvalues (the ones the reference keeps), relative to the spectrum's scale;
    # the full spectrum is compared as a sorted set (near-degenerate +-pairs may swap order by magnitude)
    rep["eval_rel"] = float(np.max(np.abs(ev[:m] - om.eigenvalues[:m])) / np.max(np.abs(om.eigenvalues)))
    rep["spectrum_abs"] = float(np.max(np.abs(np.sort(ev) - np.sort(om.eigenvalues[:r]))))
    Yo = np.concatenate([oracle.tica.tica_transform(om, p, m) for p in prepped], axis=0)
    rep["Y_rel"] = rel_err(Y.cpu().numpy()[:, :m], Yo)
    assert rep["C00_rel"] <= REL and rep["C0t_rel"] <= REL, rep
    assert rep["mu_abs"] <= REL, rep
    assert rep["eval_rel"] <= REL, rep
    # Projected coordinates inherit the covariance error amplified by the eigenvector condition number
    # ~ cond(C00) / gap (first-order perturbation theory); the bound below is that amplification applied
    # to the measured covariance error, with a floor for well-separated spectra.
    lam = np.asarray(om.eigenvalues[: min(m + 1, r)])
    gap = float(np.min(np.abs(np.diff(lam)))) if lam.size > 1 else 1.0
    s00 = np.linalg.eigvalsh(om.C00)
    cond = float(s00.max() / max(s00[s00 > 1e-6].min(), 1e-6))
    rep["Y_tol"] = max(5e-6, 4.0 * max(rep["C00_rel"], rep["C0t_rel"]) * cond / max(gap, 1e-12))
    assert rep["Y_rel"] <= rep["Y_tol"], rep
    return rep, Y[:, :m].to(torch.float32).contiguous(), segs


def check_kmeans(Y: torch.Tensor, K: int, n_iter: int, seed: int) -> tuple[dict, torch.Tensor]:
    Yh = Y.cpu().numpy().astype(np.float64)
    rng = np.random.default_rng(seed)
    c0 = Yh[np.sort(rng.choice(Yh.shape[0], size=K, replace=False))]
    # (a) labels bit-exact given identical centroids
    cd = torch.from_numpy(c0).to(Y.device)
    nre = torch.zeros((1,), dtype=torch.int64, device=Y.device)
    lab = kernels.kmeans_assign(Y, cd, n_rechecked=nre)
    lab_o, dmin_o = oracle.kmeans.assign(Yh, c0)
    n_bad = int(np.count_nonzero(lab.cpu().numpy().astype(np.int64) != lab_o))
    assert n_bad == 0, f"{n_bad} labels differ from the fp64 argmin"
    # (b) fixed number of Lloyd iterations: centres track the oracle
    res = lloyd_device(Y, cd, max_iter=n_iter, tolerance=None)
    co = c0.copy()
    for _ in range(n_iter):
        lo, _ = oracle.kmeans.assign(Yh, co)
        co, _ = oracle.kmeans._update(Yh, lo, co)
    c_rel = rel_err(res.centers.cpu().numpy(), co)
    final = kernels.kmeans_assign(Y, res.centers)
    lab_f, _ = oracle.kmeans.assign(Yh, res.centers.cpu().numpy())
    n_bad_f = int(np.count_nonzero(final.cpu().numpy().astype(np.int64) != lab_f))
    assert n_bad_f == 0, f"{n_bad_f} final labels differ"
    assert c_rel <= REL, f"Lloyd centres off by {c_rel}"
    return {"labels_mismatch": n_bad + n_bad_f, "centers_rel": c_rel,
            "rechecked_frac": float(nre.item()) / max(1, Y.shape[0])}, final


def check_counts_msm(labels: torch.Tensor, segs: Segments, K: int, lag: int, n_ts: int) -> dict:
    dev = labels.device
    C = kernels.count_lagged(labels, segs.device(dev), K, lag)
    dtrajs = segs.split(labels.cpu().numpy())
    Co = oracle.counts.count_lagged(dtrajs, K, lag)
    assert np.array_equal(C.cpu().numpy(), Co), "count matrix differs"
    assert int(Co.sum()) == segs.n_pairs(lag)
    T, pi, info, act = msm_from_counts_device(C, maxerr=1e-12)
    Ca, active = oracle.msm.ensure_connected_counts(Co.astype(float))
    To, pio, it_o = oracle.msm.mle_rev(Ca, maxerr=1e-12)
    Tf, pif = oracle.msm.expand_results(K, active, To, pio)
    rep = {"T_rel": rel_err(T.cpu().numpy(), Tf), "pi_rel": rel_err(pi.cpu().numpy(), pif),
           "mle_iters": int(info[0].item()), "mle_iters_oracle": int(it_o)}
    assert int(info[1].item()) == 1, "MLE did not converge"
    assert rep["T_rel"] <= REL and rep["pi_rel"] <= REL, rep
    oracle.msm.check_transition_matrix(T.cpu().numpy(), pi.cpu().numpy())
    k = min(n_ts + 1, K)
    ev, _ = kernels.eig_rev_topk(T, pi, k)
    evo = oracle.msm.eigenvalues_rev(To, pio, k)
    evh = ev.cpu().numpy()
    rep["eig_rel"] = float(np.max(np.abs(evh - evo) / np.maximum(np.abs(evo), 1e-3)))
    assert rep["eig_rel"] <= REL, (rep, evh, evo)
    ts, tso = safe_timescales(lag, evh[1:]), oracle.msm.safe_timescales(lag, evo[1:])
    ok = np.isfinite(tso)
    assert np.array_equal(np.isfinite(ts), ok)
    # timescales amplify eigenvalue error by 1/((1-lambda)) near 1; compare through the eigenvalues' bar
    rep["ts_rel"] = float(np.max(np.abs(ts[ok] - tso[ok]) / np.abs(tso[ok]))) if ok.any() else 0.0
    assert rep["ts_rel"] <= 1e-5, rep
    return rep


def run_and_check_small(seed: int = 1, n_traj: int = 4, n_frames: int = 600, n_res: int = 8,
                        K: int = 12, lag: int = 5) -> dict:
    """One tiny end-to-end pass, every stage checked against the oracle."""
    top = synth.backbone_topology(n_res)
    trajs = synth.backbone_trajectories(n_res, n_traj, n_frames, seed)
    report = {"featurize": check_featurize(trajs, top)}
    dev = torch.device("cuda")
    plan = plan_phi_psi_block(top)
    feats_d = featurize_device(torch.from_numpy(np.concatenate(trajs, axis=0)).to(dev), plan)
    segs = Segments.from_lengths([t.shape[0] for t in trajs])
    feats = segs.split(feats_d.cpu().numpy())
    rep, Y, segs = check_tica(feats, lag, 3, "standard")
    report["tica"] = rep
    rep, labels = check_kmeans(Y, K, 5, seed)
    report["kmeans"] = rep
    report["msm"] = check_counts_msm(labels, segs, K, lag, 4)
    return report


# ----------------------------------------------------------------------------- CK golden cases
def ck_cases(z):
    """Yield (name, dtrajs, run_ck kwargs) from tests/golden/ck.npz."""
    for name in [str(s) for s in z["case_names"]]:
        lens = [int(v) for v in z[f"{name}_lens"]]
        flat = z[f"{name}_labels"].astype(np.int64)
        offs = np.concatenate([[0], np.cumsum(lens)])
        dtrajs = [flat[offs[i]:offs[i + 1]] for i in range(len(lens))]
        kw = ast.literal_eval(str(z[f"{name}_kw"]))  # repr of a plain dict
        yield name, dtrajs, kw


def check_ck_case(z, name, dtrajs, kw, run_ck, micro, select, rtol=1e-9):
    """``run_ck`` / ``micro`` / ``select`` are the implementation under test (oracle or device)."""
    r = run_ck(dtrajs, **kw)
    assert r.mode == str(z[f"{name}_mode"]), name
    assert sorted(r.mse) == [int(k) for k in z[f"{name}_ks"]], name
    np.testing.assert_allclose([r.mse[k] for k in sorted(r.mse)], z[f"{name}_mse"], rtol=rtol, atol=1e-15)
    assert list(r.insufficient_k) == [int(k) for k in z[f"{name}_insufficient"]], name
    K = int(max(int(t.max()) for t in dtrajs)) + 1
    m = micro(dtrajs, K, int(kw["lag_time"]), factors=[2, 3, 4], max_states=int(kw["top_n_micro"]),
              min_transitions=5)
    assert bool(m.insufficient_data) == bool(z[f"{name}_mixin_insufficient"]), name
    assert sorted(m.mse) == [int(k) for k in z[f"{name}_mixin_ks"]], name
    np.testing.assert_allclose([m.mse[k] for k in sorted(m.mse)], z[f"{name}_mixin_mse"], rtol=rtol, atol=1e-15)
    if f"{name}_sel" in z.files:
        sel, taus, mses, its = select(dtrajs, K, [int(t) for t in z[f"{name}_sel_taus"]], factor=2)
        assert sel == int(z[f"{name}_sel"]), name
        np.testing.assert_allclose(mses, z[f"{name}_sel_mses"], rtol=rtol, atol=1e-15)
        np.testing.assert_allclose(its, z[f"{name}_sel_its"], rtol=1e-8)


def block_lumper(T, n_macro):
    """Stand-in for PCCA+ used by the golden generator: contiguous blocks of states."""
    n = T.shape[0]
    return (np.arange(n) * int(n_macro)) // n


def selector_cases(z):
    """Yield (name, dtrajs, kwargs, lumper) from tests/golden/ck_selector.npz."""
    for name in [str(s) for s in z["case_names"]]:
        lens = [int(v) for v in z[f"{name}_lens"]]
        flat = z[f"{name}_labels"].astype(np.int64)
        offs = np.concatenate([[0], np.cumsum(lens)])
        dtrajs = [flat[offs[i]:offs[i + 1]] for i in range(len(lens))]
        kw = ast.literal_eval(str(z[f"{name}_kw"]))
        yield name, dtrajs, kw, (block_lumper if bool(z[f"{name}_blocks"]) else None)


def check_selector_case(z, name, dtrajs, kw, lumper, select, mle_rtol=1e-6):
    sel, evs = select(dtrajs, macro_lumper=lumper, **kw)
    assert sel == int(z[f"{name}_selected"]), name
    assert [e.lag for e in evs] == [int(v) for v in z[f"{name}_lags"]], name
    np.testing.assert_allclose([e.ck_error for e in evs], z[f"{name}_ck_error"], rtol=1e-8, err_msg=name)
    np.testing.assert_allclose([e.coverage_fraction for e in evs], z[f"{name}_coverage"], rtol=1e-12, err_msg=name)
    assert [e.median_count for e in evs] == [int(v) for v in z[f"{name}_median"]], name
    assert [e.n_macrostates for e in evs] == [int(v) for v in z[f"{name}_n_macro"]], name
    assert [bool(e.passed_sanity) for e in evs] == [bool(v) for v in z[f"{name}_passed"]], name
    assert ["" if e.failure_reason is None else e.failure_reason for e in evs] == [str(v) for v in z[f"{name}_reason"]], name
    dm = np.array([np.nan if e.diag_mass is None else e.diag_mass for e in evs], dtype=float)
    np.testing.assert_allclose(dm, z[f"{name}_diag_mass"], rtol=mle_rtol, equal_nan=True, err_msg=name)
    gap = np.array([np.nan if e.eigenvalue_gap is None else e.eigenvalue_gap for e in evs], dtype=float)
    np.testing.assert_allclose(gap, z[f"{name}_gap"], rtol=1e-8, atol=1e-12, equal_nan=True, err_msg=name)
    ts = np.array([[np.nan] * 3 if e.timescales is None else list(np.asarray(e.timescales)[:3]) for e in evs], dtype=float)
    np.testing.assert_allclose(ts, z[f"{name}_ts3"], rtol=mle_rtol, equal_nan=True, err_msg=name)
