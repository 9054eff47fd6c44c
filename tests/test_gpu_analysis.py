"""GPU tests for SURVEY.md 8f-3 / 8f-4: histogram free-energy surface, transition-path theory, streaming ingest."""

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


@pytest.mark.parametrize("n,bins", [(1000, (7, 5)), (300000, (100, 100)), (50000, (128, 128))])
def test_hist2d_matches_numpy(n, bins):
    from pmarlo_b200 import kernels

    rng = np.random.default_rng(n)
    x, y = rng.normal(size=n), rng.normal(size=n) * 2 + 1
    x[::97] = np.nan
    x[5], y[6] = -2.0, 7.0                         # exactly on the outer edges: last bin is closed
    x[7] = -2.0 + 4.0 * 3 / bins[0]                 # exactly on an inner edge
    ranges = ((-2.0, 2.0), (-5.0, 7.0))
    w = rng.random(n)
    for weights in (None, w):
        ok = np.isfinite(x)
        ref, _, _ = np.histogram2d(x[ok], y[ok], bins=bins, range=ranges, weights=None if weights is None else weights[ok])
        got = kernels.hist2d(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), bins, ranges,
                             None if weights is None else torch.from_numpy(weights).cuda()).cpu().numpy()
        if weights is None:
            np.testing.assert_array_equal(got, ref)
        else:
            np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12)


def test_generate_2d_fes_matches_numpy_restatement():
    from pmarlo_b200 import free_energy_from_density, generate_2d_fes
    from pmarlo_b200.analysis import kT_kJ_per_mol

    rng = np.random.default_rng(2)
    phi = np.concatenate([rng.normal(-60, 15, 40000), rng.normal(60, 20, 20000)])
    psi = np.concatenate([rng.normal(-45, 15, 40000), rng.normal(150, 25, 20000)])
    res = generate_2d_fes(phi, psi, bins=(36, 36), temperature=300.0, periodic=(True, True), min_count=1)
    xw = (phi + 180) % 360 - 180
    yw = (psi + 180) % 360 - 180
    H, xe, ye = np.histogram2d(xw, yw, bins=(36, 36), range=((-180, 180), (-180, 180)))
    dens = H / (H.sum() * np.diff(xe)[:, None] * np.diff(ye)[None, :])
    F = free_energy_from_density(dens, 300.0, mask=H < 1)
    np.testing.assert_array_equal(res.metadata["counts"], H)
    np.testing.assert_allclose(res.F, F, rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(res.xedges, xe)
    assert np.nanmin(res.F) == 0.0 and np.isnan(res.F[H < 1]).all()
    # two wells with populations 2 : 1 differ by about kT ln 2 (different widths: compare basin minima loosely)
    assert abs(kT_kJ_per_mol(300.0) - 2.494) < 0.01
    ad = generate_2d_fes(phi, psi, bins=(20, 20), grid_strategy="adaptive")
    assert ad.xedges[0] > phi.min() and ad.xedges[-1] < phi.max()
    with pytest.raises(ValueError):
        free_energy_from_density(dens, -1.0)


def test_tpt_matches_oracle_and_conserves_flux():
    from pmarlo_b200 import TPTAnalysis

    rng = np.random.default_rng(3)
    K = 40
    Cm = rng.random((K, K)) ** 3
    Cm = Cm + Cm.T + np.diag(5 * rng.random(K))
    T = Cm / Cm.sum(axis=1, keepdims=True)
    pi = Cm.sum(axis=1) / Cm.sum()
    A, B = [0, 1, 2], [37, 38, 39]
    r = TPTAnalysis(T, pi).analyze(A, B, n_paths=5, pathway_fraction=0.9)
    o = oracle.tpt.reactive_flux(T, pi, A, B)
    np.testing.assert_allclose(r.forward_committor, o["qf"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(r.backward_committor, o["qb"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(r.flux_matrix, o["gross"], rtol=1e-10, atol=1e-16)
    np.testing.assert_allclose(r.net_flux, o["net"], rtol=1e-9, atol=1e-16)
    assert abs(r.total_flux - o["total_flux"]) <= 1e-12 * o["total_flux"] and abs(r.rate - o["rate"]) <= 1e-10 * o["rate"]
    np.testing.assert_allclose(r.backward_committor, 1.0 - r.forward_committor, atol=1e-10)   # reversible chain
    inter = np.setdiff1d(np.arange(K), A + B)
    np.testing.assert_allclose(r.net_flux[inter].sum(axis=1), r.net_flux[:, inter].sum(axis=0), rtol=1e-9)   # conservation
    assert len(r.pathways) >= 1 and all(p[0] in A and p[-1] in B for p in r.pathways)
    assert np.all(np.diff(r.pathway_fluxes) <= 1e-15) and r.pathway_fluxes.sum() <= r.total_flux * (1 + 1e-9)
    with pytest.raises(ValueError):
        TPTAnalysis(T, pi).analyze([0, 1], [1, 2])


def test_featurize_stream_from_dcd_matches_in_memory(tmp_path, topologies):
    from pmarlo_b200.features import featurize_device, plan_phi_psi_block
    from pmarlo_b200.io import featurize_stream, write_dcd
    from pmarlo_b200.topology import Topology

    t = topologies["ala2"]
    top = Topology(list(t["names"]), np.asarray(t["resid"], dtype=np.int64), np.asarray(t["chain"], dtype=np.int64))
    trajs, _ = synth.ala2_trajectories(t, n_traj=2, n_frames=[700, 650], seed=3)
    xyz = np.concatenate(trajs)
    path = tmp_path / "ala2.dcd"
    write_dcd(path, xyz)
    plan = plan_phi_psi_block(top)
    got = featurize_stream(path, top, plan, chunk=256)
    from pmarlo_b200.io import DCDReader

    ref = featurize_device(torch.from_numpy(DCDReader(path).read(0, 1350)).cuda(), plan)
    assert got.shape == (1350, plan.n_cols)
    assert torch.equal(got, ref)


def _metastable_T(rng, n_blocks, size, eps):
    K = n_blocks * size
    C = np.zeros((K, K))
    for b in range(n_blocks):
        C[b * size:(b + 1) * size, b * size:(b + 1) * size] = rng.integers(20, 100, size=(size, size))
    C = C + C.T
    for b in range(n_blocks):                       # weak symmetric links between neighbouring blocks
        i, j = b * size + size - 1, ((b + 1) % n_blocks) * size
        C[i, j] += eps
        C[j, i] += eps
    pi = C.sum(axis=1) / C.sum()
    return C / C.sum(axis=1, keepdims=True), pi


@pytest.mark.gpu
@pytest.mark.parametrize("n_blocks,size", [(3, 5), (4, 40), (6, 25)])
def test_pcca_memberships_match_oracle_and_recover_blocks(n_blocks, size):
    """PCCA+ (macro.pcca_memberships: device eigenvectors + device inner simplex + host Nelder-Mead) against the
    oracle's loop restatement of the published algorithm, and against the known answer of a nearly uncoupled
    block matrix.  The optimiser's tolerances (fmin defaults, xtol = ftol = 1e-4) bound the agreement."""
    from oracle import pcca as opcca
    from pmarlo_b200 import macro

    rng = np.random.default_rng(n_blocks * 100 + size)
    T, pi = _metastable_T(rng, n_blocks, size, eps=1.0)
    chi = macro.pcca_memberships(T, n_blocks, pi)
    ref = opcca.pcca_memberships(T, n_blocks, pi)
    assert chi.shape == (n_blocks * size, n_blocks)
    np.testing.assert_allclose(chi.sum(axis=1), 1.0, atol=1e-12)
    assert chi.min() >= 0.0
    # columns may come out permuted relative to the oracle only if the simplex vertices differ: they must not
    np.testing.assert_allclose(chi, ref, atol=2e-3)
    hard = chi.argmax(axis=1)
    for b in range(n_blocks):
        assert len(set(hard[b * size:(b + 1) * size])) == 1          # one macrostate per block
    assert len(set(hard)) == n_blocks
    assert chi.max(axis=1).min() > 0.9
    # pi omitted: computed on the device
    np.testing.assert_allclose(macro.pcca_memberships(T, n_blocks), chi, atol=1e-6)


@pytest.mark.gpu
def test_pcca_like_macrostates_contract_and_ck_macro_branch():
    """Reference contract of pcca_like_macrostates (_msm_utils.py:284-299): labels numbered by descending
    population, None for a matrix with <= n_macrostates states or one PCCA+ rejects (irreversible / disconnected);
    and the macrostate branch of run_ck with macro_lumper="pcca"."""
    from pmarlo_b200 import ck, macro

    rng = np.random.default_rng(5)
    T, pi = _metastable_T(rng, 3, 6, eps=0.5)
    lab = macro.pcca_like_macrostates(T, 3)
    assert lab.shape == (18,) and set(lab) == {0, 1, 2}
    pops = macro.compute_macro_populations(pi, lab)
    assert np.all(np.diff(pops) <= 1e-12)                                # descending
    assert macro.pcca_like_macrostates(T[:3, :3] / T[:3, :3].sum(1, keepdims=True), 4) is None
    cyc = np.roll(np.eye(6), 1, axis=1) * 0.9 + np.eye(6) * 0.1          # irreversible cycle
    assert macro.pcca_like_macrostates(cyc, 2) is None
    two = np.kron(np.eye(2), np.full((3, 3), 1.0 / 3.0))                 # disconnected
    assert macro.pcca_like_macrostates(two, 2) is None
    # a three-well trajectory: the macro CK test runs on PCCA+ macrostates
    K = 18
    P = T
    s = np.zeros(60000, dtype=np.int64)
    cum = np.cumsum(P, axis=1)
    u = rng.random(s.size)
    for t in range(1, s.size):
        s[t] = min(int(np.searchsorted(cum[s[t - 1]], u[t])), K - 1)
    # The row-normalised count matrix run_ck lumps is not reversible, and PCCA+ (deeptime's as well) insists on
    # detailed balance: the reference's wrapper turns that ValueError into None and the microstate test runs.
    res = ck.run_ck([s], 5, None, macro_k=3, min_trans=20, macro_lumper="pcca")
    assert res.mode == "micro" and len(res.mse) >= 1

    def reversibilised_pcca(Tm, k):
        w, V = np.linalg.eig(Tm.T)
        p = np.abs(np.real(V[:, np.argmax(np.real(w))]))
        F = (p / p.sum())[:, None] * Tm
        Fs = 0.5 * (F + F.T)
        return macro.pcca_like_macrostates(Fs / Fs.sum(axis=1, keepdims=True), k)

    res = ck.run_ck([s], 5, None, macro_k=3, min_trans=20, macro_lumper=reversibilised_pcca)
    assert res.mode == "macro" and len(res.mse) >= 1
    assert all(np.isfinite(v) for v in res.mse.values())
