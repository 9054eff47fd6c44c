"""Parity of the CUDA path (through the C ABI of libpmb200.so) against the CPU
oracle and the committed golden vectors.  Run on the B200 box: ``pytest -m gpu``."""

from __future__ import annotations

import numpy as np
import pytest
import torch

import oracle
from tests import parity, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from pmarlo_b200 import load_library

    load_library()
    torch.cuda.set_device(0)
    yield
    torch.cuda.synchronize()


def dev():
    return torch.device("cuda")


# ----------------------------------------------------------------------------- featurize
@pytest.mark.parametrize("n_res,n_traj,n_frames", [(4, 1, 7), (8, 3, 257), (33, 2, 1500)])
def test_featurize_vs_oracle(n_res, n_traj, n_frames):
    top = synth.backbone_topology(n_res)
    trajs = synth.backbone_trajectories(n_res, n_traj, n_frames, seed=n_res)
    parity.check_featurize(trajs, top)


def test_featurize_analytic_geometries():
    """Planar cis/trans = 0 / pi, +-90 degree constructions."""
    from pmarlo_b200.features import FeaturePlan, featurize_device

    def frame(angle):
        # atoms 0,1,2 in the xy plane; atom 3 rotated by `angle` about the 1->2 axis
        p0, p1, p2 = [1.0, 0.0, 0.0], [0.0, 0.0, 0.0], [0.0, 0.0, 1.0]
        p3 = [np.cos(angle), np.sin(angle), 1.0]
        return [p0, p1, p2, p3]

    angles = np.array([0.0, np.pi, np.pi / 2, -np.pi / 2, 0.3, -2.5, 3.0])
    xyz = np.asarray([frame(a) for a in angles], dtype=np.float32) * 0.15
    plan = FeaturePlan(np.asarray([[0, 0, 1, 2, 3, 0, 1, 2]], dtype=np.int32), 3)
    got = featurize_device(torch.from_numpy(xyz).to(dev()), plan).cpu().numpy()
    ref = oracle.featurize.compute_dihedrals(xyz, np.asarray([[0, 1, 2, 3]]))[:, 0]
    assert parity.angle_err(got[:, 0], ref) < 1e-6
    # sign convention (IUPAC, as mdtraj): the construction above yields angle = +a
    assert parity.angle_err(got[:, 0], angles) < 1e-6
    np.testing.assert_allclose(got[:, 1], np.cos(ref), atol=1e-6)
    np.testing.assert_allclose(got[:, 2], np.sin(ref), atol=1e-6)
    assert np.all(got[:, 0] > -np.pi) and np.all(got[:, 0] <= np.float32(np.pi))


def test_featurize_trajectory_api(topologies):
    """Drop-in signatures on the two reference topologies (ala2: 1 phi + 1 psi; chignolin: 45 CA pairs)."""
    from pmarlo_b200 import Trajectory, Topology, compute_features, featurize_trajectory, trig_expand_periodic

    for key, n_ang, n_ca in (("ala2", 2, 1), ("chig", 18, 10)):
        t = topologies[key]
        top = Topology(t["names"], t["resid"], t["chain"])
        rng = np.random.default_rng(5)
        xyz = (t["xyz"][None] + rng.normal(scale=0.01, size=(50,) + t["xyz"].shape)).astype(np.float32)
        traj = Trajectory(xyz, top)
        X = featurize_trajectory(traj, "phi_psi")
        ref = oracle.featurize.featurize_trajectory(xyz, t["names"], t["resid"], t["chain"], "phi_psi")
        assert X.shape == ref.shape == (50, n_ang) and X.dtype == np.float32
        assert parity.angle_err(X, ref) < parity.ANGLE_ABS
        if n_ca >= 2:
            D = featurize_trajectory(traj, "ca_distances")
            Dref = oracle.featurize.featurize_trajectory(xyz, t["names"], t["resid"], t["chain"], "ca_distances")
            assert D.shape == (50, n_ca * (n_ca - 1) // 2)
            assert parity.rel_err(D, Dref) < parity.DIST_REL
        else:
            with pytest.raises(ValueError):
                featurize_trajectory(traj, "ca_distances")
        Xc, cols, per = compute_features(traj, ["phi_psi"])
        assert Xc.shape == X.shape and len(cols) == n_ang and per.all()
        assert cols[0].startswith("phi:res") and cols[-1].startswith("psi:res")
        Xe, mapping = trig_expand_periodic(Xc, per)
        Xeo, mo = oracle.featurize.trig_expand_periodic(Xc.astype(np.float64), per)
        np.testing.assert_array_equal(mapping, mo)
        np.testing.assert_allclose(Xe, Xeo, atol=1e-12)
    with pytest.raises(ValueError):
        featurize_trajectory(traj, "nope")
    with pytest.raises(ValueError):
        trig_expand_periodic(np.zeros((3, 2)), np.array([True]))


# ----------------------------------------------------------------------------- TICA
@pytest.mark.parametrize("d,lag,dim,pre", [(4, 10, 2, "standard"), (12, 3, 5, None), (37, 7, 10, "center"),
                                            (81, 10, 10, "standard"), (256, 20, 10, "standard")])
def test_tica_vs_oracle(d, lag, dim, pre):
    n_traj = 3 if d < 200 else 2
    feats = synth.ar1_features(n_traj, 4000 if d < 200 else 6000, d, seed=d, offset=3.0)
    feats[1] = feats[1][: 4000 - 333]          # ragged shards
    if d == 12:
        feats.append(feats[0][:lag])           # a shard no longer than the lag contributes no pair
    parity.check_tica(feats, lag, dim, pre)


@pytest.mark.parametrize("d,n_frames,lag", [(256, 9000, 20), (64, 5000, 3), (96, 4100, 7), (160, 6000, 11)])
def test_gram_tensor_path_matches_fp64(d, n_frames, lag):
    """tcgen05 Gram (impl=2: exact leading term + TF32 remainders, fp64 window drain) against an fp64
    evaluation of the same conditioned fp32 data, and against the SIMT kernel.  Tolerance 1e-7 of the
    matrix scale (north star: covariances within 1e-6)."""
    from pmarlo_b200 import kernels
    from pmarlo_b200.shards import Segments

    feats = synth.ar1_features(3, n_frames, d, seed=d + 1, offset=2.0)
    feats[1] = feats[1][: n_frames - 777]
    feats.append(feats[0][: lag])                      # contributes no pair
    feats[0][5, 3] = np.nan
    X = np.concatenate(feats).astype(np.float32)
    segs = Segments.from_lengths([f.shape[0] for f in feats])
    Xd = torch.from_numpy(X).to(dev())
    mask = kernels.pair_mask(segs.device(dev()), X.shape[0], lag)
    shift = np.nanmean(X, axis=0).astype(np.float32)
    scale = (1.0 / np.nanstd(X, axis=0)).astype(np.float32)
    cond = torch.from_numpy(np.stack([shift, scale])).to(dev())
    m = mask.cpu().numpy()
    z = np.where(np.isnan(X), np.float32(0), (X - shift[None]) * scale[None]).astype(np.float32)
    w0 = ((m & 1) + ((m >> 1) & 1)).astype(np.float64)
    G0 = (z.astype(np.float64) * w0[:, None]).T @ z.astype(np.float64)
    idx = np.flatnonzero(m & 1)
    v = (z[idx] - z[idx + lag]).astype(np.float32).astype(np.float64)
    G1 = v.T @ v
    for mode, ref in ((0, G0), (1, G1)):
        simt = kernels.gram(Xd, mask, lag, mode, cond, impl=1).cpu().numpy()
        scale_ref = np.max(np.abs(ref))
        e_simt = np.max(np.abs(simt - ref)) / scale_ref
        # impl 5: kind::f16 kernel (gram_h.cu, what impl=0 selects for large n); impl 2: kind::tf32 kernel (its
        # out-of-range fallback)
        for impl in (5, 2):
            tc = kernels.gram(Xd, mask, lag, mode, cond, impl=impl).cpu().numpy()
            e_tc = np.max(np.abs(tc - ref)) / scale_ref
            e_diag = np.max(np.abs(np.diag(tc) - np.diag(ref)) / np.abs(np.diag(ref)))
            print(f"gram d={d} mode={mode} impl={impl}: tcgen05 err {e_tc:.2e} (diag rel {e_diag:.2e}), SIMT err {e_simt:.2e}")
            # diagonal entries of mode 1 are sums of small squared differences: relative to themselves the
            # stochastic-rounding split is good to a few 1e-7, relative to the matrix scale to ~3e-8
            assert e_tc <= 1e-7 and e_diag <= 5e-7
            np.testing.assert_array_equal(tc, tc.T)
        if d == 256:
            # the CTA-pair kernel (cta_group::2, impl=4) must agree with the default single-CTA kernel
            one = kernels.gram(Xd, mask, lag, mode, cond, impl=4).cpu().numpy()
            assert np.max(np.abs(one - ref)) / scale_ref <= 1e-7
            assert np.max(np.abs(one - tc)) / scale_ref <= 2e-8


def test_gram_f16_path_edge_weights_and_out_of_range_fallback():
    """gram_h.cu specifics.  (i) Many short shards: a third of the frames carry weight 1 in mode 0, i.e. the
    second (edge-frame) pass matters.  (ii) A z-score beyond the accepted range of the clamped leading term
    (|z| > 132), or an infinite input: the device flag makes the SIMT kernel recompute the matrix, no host round
    trip; (iii) |z| = 120 stays on the fp16 path (inside the leading term's range; the exactness budget shortens that window)."""
    from pmarlo_b200 import kernels
    from pmarlo_b200.shards import Segments

    d, lag = 128, 5
    rng = np.random.default_rng(11)
    lens = [int(v) for v in rng.integers(12, 40, size=700)] + [3, 5, 1]
    X = rng.normal(size=(sum(lens), d)).astype(np.float32) + 1.5
    segs = Segments.from_lengths(lens)
    shift = X.mean(axis=0).astype(np.float32)
    scale = (1.0 / X.std(axis=0)).astype(np.float32)
    for case, big in (("edges", 0.0), ("large_z", 120.0), ("out_of_range", 5000.0)):
        Xc = X.copy()
        if big:
            Xc[1234, 7] = shift[7] + big / scale[7]
            Xc[4321, 9] = np.nan                      # imputed with 0 on every path
        Xd = torch.from_numpy(Xc).to(dev())
        mask = kernels.pair_mask(segs.device(dev()), Xc.shape[0], lag)
        cond = torch.from_numpy(np.stack([shift, scale])).to(dev())
        m = mask.cpu().numpy()
        z = np.where(np.isnan(Xc), np.float32(0), (Xc - shift[None]) * scale[None]).astype(np.float32)
        w0 = ((m & 1) + ((m >> 1) & 1)).astype(np.float64)
        assert 0.2 < np.mean(w0 == 1) < 0.6
        G0 = (z.astype(np.float64) * w0[:, None]).T @ z.astype(np.float64)
        idx = np.flatnonzero(m & 1)
        v = (z[idx] - z[idx + lag]).astype(np.float32).astype(np.float64)
        G1 = v.T @ v
        for mode, ref in ((0, G0), (1, G1)):
            got = kernels.gram(Xd, mask, lag, mode, cond, impl=5).cpu().numpy()
            simt = kernels.gram(Xd, mask, lag, mode, cond, impl=1).cpu().numpy()
            err = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
            print(f"gram f16 {case} mode={mode}: err {err:.2e}")
            assert err <= 1e-7, (case, mode, err)
            if case == "out_of_range":
                np.testing.assert_array_equal(got, simt)    # the fallback IS the SIMT kernel


@pytest.mark.parametrize("case", ["stationary", "drifting", "nan"])
def test_tica_streaming_accumulator_matches_one_shot_fit(case):
    """Chunk-wise fit (conditioning from the first chunk, or the fallback second pass when the first
    chunk is not representative / NaNs are present) against the one-shot fit: covariances within 1e-9."""
    from pmarlo_b200.reduction import TICA
    from pmarlo_b200.shards import Segments, concat_to_device

    d, lag = 64, 7
    feats = synth.ar1_features(5, 30000, d, seed=11, offset=1.5)
    if case == "drifting":
        for i in range(2, 5):
            feats[i] = (feats[i] * 40.0 + 300.0).astype(np.float32)   # first chunk says nothing about these
    if case == "nan":
        feats[3][100, 5] = np.nan
    X, segs = concat_to_device(feats, dev())
    est = TICA(lag, 5, preprocess="standard")
    ref = est.fit_device(X, segs)
    acc = est.accumulator(d, dev())
    offs = segs.offsets
    for lo, hi in ((0, 2), (2, 3), (3, 5)):
        a, b = int(offs[lo]), int(offs[hi])
        acc.add(X[a:b], Segments(offs[lo:hi + 1] - a))
    got = acc.finish()
    assert acc.fell_back == (case != "stationary")
    flat = np.concatenate(feats).astype(np.float64)
    Z = oracle.tica.preprocess(flat, scale=True)
    om = oracle.tica.tica_fit(segs.split(Z), lag)
    print(case, "vs oracle: streaming", parity.rel_err(got.C00.cpu().numpy(), om.C00), "one-shot", parity.rel_err(ref.C00.cpu().numpy(), om.C00))
    assert got.n_pairs == ref.n_pairs and got.rank == ref.rank
    for name in ("C00", "C0t"):
        e = parity.rel_err(getattr(got, name).cpu().numpy(), getattr(ref, name).cpu().numpy())
        assert e < 1e-6, (case, name, e)   # two conditionings of the same fp32 data
    m = min(5, ref.rank)
    # the drifting data set is nearly rank one (condition number ~1e6): eigenvalues amplify the 1e-8 covariance error
    np.testing.assert_allclose(got.eigenvalues.cpu().numpy()[:m], ref.eigenvalues.cpu().numpy()[:m],
                               rtol=2e-5 if case == "drifting" else 1e-6)


def test_host_buffer_pipeline_matches_device_resident_pipeline():
    """estimate_msm_from_host (chunked copies overlapped with featurize / moments / Gram) gives the same
    MSM as run_pipeline on resident coordinates."""
    import bench
    from pmarlo_b200.pipeline import estimate_msm_from_host, run_pipeline
    from pmarlo_b200.shards import Segments

    wl = bench.make_workload(6, 20000, dev(), seed=5)
    cfg = bench.bench_config(n_states=200, kmeans_iters=4)
    res = run_pipeline(wl.xyz, wl.segs, wl.plan, cfg)
    host = wl.xyz.cpu().pin_memory()
    out = estimate_msm_from_host(host, [20000] * 6, wl.plan, cfg, chunk_frames=40000)
    np.testing.assert_allclose(out["eigenvalues"], res.eigenvalues.cpu().numpy(), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(out["stationary_distribution"], res.pi.cpu().numpy(), rtol=1e-5, atol=1e-12)


def test_tica_nan_imputation_and_constant_column(golden):
    from pmarlo_b200.reduction import preprocess, tica_reduce

    z = golden("preprocess")
    for key_in, key_out, scale in (("X", "P_scale", True), ("X", "P_noscale", False),
                                   ("Xn", "Pn_scale", True), ("Xn", "Pn_noscale", False)):
        X32 = z[key_in].astype(np.float32)
        got = preprocess(X32, scale=scale)
        ref = oracle.tica.preprocess(X32.astype(np.float64), scale=scale)
        np.testing.assert_allclose(got, ref, rtol=0, atol=2e-6 * max(1.0, np.abs(ref).max()))
        # the golden vectors were made from float64 inputs: float32 rounding of the inputs is the only gap
        np.testing.assert_allclose(got, z[key_out], rtol=0, atol=5e-6 * max(1.0, np.abs(z[key_out]).max()))
    Xn = z["Xn"].astype(np.float32)
    Y = tica_reduce(Xn, lag=2, n_components=2, scale=True)
    Yo = oracle.tica.tica_reduce(Xn.astype(np.float64), lag=2, n_components=2, scale=True)
    assert Y.shape == Yo.shape and Y.dtype == np.float64
    assert parity.rel_err(Y, Yo) < 5e-6


def test_tica_ar1_spectrum(golden):
    """Analytic pin: TICA eigenvalues of independent AR(1) processes are rho^lag; the reference's own
    numpy estimator (golden ref_eigs) agrees to O(1/N)."""
    from pmarlo_b200.reduction import TICA

    z = golden("tica_xcheck")
    X, lag, rho = z["X"], int(z["lag"]), z["rho"]
    model = TICA(lag, 4, preprocess=None).fit([X])
    ev = model.eigenvalues.cpu().numpy()[:4]
    np.testing.assert_allclose(np.sort(ev)[::-1], np.sort(rho ** lag)[::-1], atol=0.03)
    np.testing.assert_allclose(np.sort(ev)[::-1], np.sort(z["ref_eigs"])[::-1], atol=5e-3)
    om = oracle.tica.tica_fit([X.astype(np.float64)], lag)
    assert np.max(np.abs(ev - om.eigenvalues[:4])) / np.max(np.abs(om.eigenvalues)) < parity.REL


def test_maybe_apply_tica_drops_lag_frames():
    from pmarlo_b200.reduction import maybe_apply_tica

    feats = synth.ar1_features(3, 500, 6, seed=2)
    lengths = [f.shape[0] for f in feats]
    flat = np.concatenate(feats)
    Y, ncomp, kept = maybe_apply_tica(flat, lengths, 9, 7)
    Yo, nco = oracle.tica.maybe_apply_tica(flat.astype(np.float64), lengths, 9, 7)
    assert ncomp == nco == 5 and kept == [493] * 3
    assert Y.shape == Yo.shape and parity.rel_err(Y, Yo) < 5e-6


@pytest.mark.parametrize("d,m,n", [(256, 10, 5003), (256, 16, 4100), (256, 1, 777), (256, 13, 3000), (128, 4, 6001),
                                   (64, 3, 2500), (200, 7, 3333), (84, 3, 1000), (320, 10, 2048)])
def test_project_fp32_paths_vs_fp64(d, m, n):
    """K5 with fp32 output: the warp-per-frame kernel (packed fp32 FMAs over m rounded up to even, the
    lane-halving reduction with compile-time-zero slots skipped) for d <= 256 and the thread-per-frame kernel
    beyond, against (impute(X) - a) W in fp64; NaN entries take the imputation value.  Tolerance: fp32
    products summed pairwise, 1e-6 of the scale of the output."""
    from pmarlo_b200 import kernels

    g = torch.Generator(device=dev()).manual_seed(d * 1000 + m)
    X = torch.randn((n, d), generator=g, device=dev(), dtype=torch.float32) * 0.7 + 0.3
    X[5, min(7, d - 1)] = float("nan")
    X[n - 1, 0] = float("nan")
    a = 0.3 + 1e-3 * torch.randn(d, generator=g, device=dev(), dtype=torch.float64)
    fill = a + 0.01
    W = torch.randn((d, m), generator=g, device=dev(), dtype=torch.float64) / d ** 0.5
    Y = kernels.project(X, a, fill, W)
    assert Y.dtype == torch.float32 and tuple(Y.shape) == (n, m)
    Xd = X.double()
    Xd = torch.where(torch.isnan(Xd), fill.expand_as(Xd), Xd)
    ref = (Xd - a) @ W
    err = float((Y.double() - ref).abs().max().item()) / float(ref.abs().max().item())
    assert err < 1e-6, err
    Y64 = kernels.project(X, a, fill, W, out_f64=True)
    assert float((Y64 - ref).abs().max().item()) / float(ref.abs().max().item()) < 1e-12


def test_subsampled_hints_never_change_labels():
    """clustering.subsampled_hints: the labels of a 1-in-16 subsample repeated over the frames in between are
    hints for the first (otherwise cold) Lloyd assignment.  Labels with the hints equal the cold labels on a
    slow trajectory (hints kept) and on the same frames shuffled (hints withdrawn on the device: all -1)."""
    from pmarlo_b200 import kernels
    from pmarlo_b200.clustering import subsampled_hints

    n, D, K = 300_000, 10, 300
    rng = np.random.default_rng(3)
    e = rng.standard_normal((n, D)).astype(np.float32)
    y = np.empty((n, D), dtype=np.float32)
    y[0] = e[0]
    rho = np.float32(0.999)
    s = np.float32(np.sqrt(1 - 0.999 ** 2))
    for t in range(1, n):
        y[t] = rho * y[t - 1] + s * e[t]
    for Yh in (y, y[rng.permutation(n)]):
        Y = torch.from_numpy(np.ascontiguousarray(Yh)).to(dev())
        C = Y[torch.from_numpy(np.sort(rng.choice(n, K, replace=False))).to(dev())].double().contiguous()
        hints = subsampled_hints(Y, C)
        cold = kernels.kmeans_assign(Y, C, impl=2)
        warm = kernels.kmeans_assign(Y, C, impl=2, hints=hints)
        assert torch.equal(cold, warm)
        kept = bool((hints >= 0).all().item())
        assert kept == (Yh is y), "hints are kept for the trajectory and withdrawn for the shuffled frames"
        if kept:
            assert torch.equal(hints[::16], cold[::16])


# ----------------------------------------------------------------------------- k-means
@pytest.mark.parametrize("D,K,n", [(2, 200, 20000), (3, 17, 5000), (10, 1000, 30000), (10, 3000, 8000),
                                   (16, 64, 4000), (64, 300, 3000)])
def test_kmeans_vs_oracle(D, K, n):
    feats = synth.ar1_features(1, n, D, seed=D * 7 + K)[0]
    Y = torch.from_numpy(feats).to(dev())
    parity.check_kmeans(Y, K, 3, seed=K)


@pytest.mark.parametrize("D,K,n", [(10, 1000, 40000), (3, 17, 5000), (10, 500, 1000), (2, 200, 70000),
                                   (16, 400, 20011), (1, 5, 300)])
def test_kmeans_tensor_path_labels_and_accumulation(D, K, n):
    """tcgen05 score GEMM + fused argmin (impl=2): labels bit-exact against the fp64 argmin, the fused
    sums / counts / inertia equal to the oracle's, re-check fraction small."""
    from pmarlo_b200 import kernels

    feats = synth.ar1_features(1, n, D, seed=D * 11 + K)[0]
    Y = torch.from_numpy(feats).to(dev())
    Yh = feats.astype(np.float64)
    rng = np.random.default_rng(K)
    c0 = Yh[np.sort(rng.choice(n, size=K, replace=False))] + 1e-3 * rng.normal(size=(K, D))
    cd = torch.from_numpy(c0).to(dev())
    sums = torch.zeros((K, D), dtype=torch.float64, device=dev())
    counts = torch.zeros((K,), dtype=torch.int64, device=dev())
    inertia = torch.zeros((1,), dtype=torch.float64, device=dev())
    nre = torch.zeros((1,), dtype=torch.int64, device=dev())
    lab = kernels.kmeans_assign(Y, cd, sums=sums, counts=counts, inertia=inertia, n_rechecked=nre, impl=2)
    lab_o, dmin_o = oracle.kmeans.assign(Yh, c0)
    np.testing.assert_array_equal(lab.cpu().numpy().astype(np.int64), lab_o)
    so = np.zeros((K, D))
    np.add.at(so, lab_o, Yh)
    np.testing.assert_array_equal(counts.cpu().numpy(), np.bincount(lab_o, minlength=K))
    assert parity.rel_err(sums.cpu().numpy(), so) < 1e-12
    assert abs(float(inertia.item()) - float(dmin_o.sum())) <= 1e-10 * float(dmin_o.sum())   # identity-based (kmeans_tc_commit_kernel)
    frac = float(nre.item()) / n
    print(f"tensor path D={D} K={K}: re-checked fraction {frac:.2e}")
    assert frac < 0.05
    # the SIMT path gives the same labels
    np.testing.assert_array_equal(kernels.kmeans_assign(Y, cd, impl=1).cpu().numpy(), lab.cpu().numpy())


@pytest.mark.parametrize("kind", ["exact", "random", "garbage", "stale"])
def test_kmeans_tensor_path_hints_never_change_labels(kind):
    """Hints only let the epilogue skip 32-centre blocks; labels, sums, counts must be identical for
    correct, random, out-of-range and stale hints."""
    from pmarlo_b200 import kernels

    n, D, K = 50000, 10, 1000
    feats = synth.ar1_features(1, n, D, seed=5)[0]
    Y = torch.from_numpy(feats).to(dev())
    Yh = feats.astype(np.float64)
    rng = np.random.default_rng(3)
    c0 = Yh[np.sort(rng.choice(n, size=K, replace=False))]
    lab_o, _ = oracle.kmeans.assign(Yh, c0)
    if kind == "exact":
        h = lab_o.astype(np.int32)
    elif kind == "random":
        h = rng.integers(0, K, size=n).astype(np.int32)
    elif kind == "garbage":
        h = rng.integers(-2**31, 2**31 - 1, size=n).astype(np.int32)
    else:
        h = np.roll(lab_o, 977).astype(np.int32)
    hints = torch.from_numpy(h).to(dev())
    sums = torch.zeros((K, D), dtype=torch.float64, device=dev())
    counts = torch.zeros((K,), dtype=torch.int64, device=dev())
    nre = torch.zeros((1,), dtype=torch.int64, device=dev())
    lab = kernels.kmeans_assign(Y, torch.from_numpy(c0).to(dev()), labels=hints, sums=sums, counts=counts,
                                n_rechecked=nre, impl=2, hints=hints)      # aliasing labels and hints
    np.testing.assert_array_equal(lab.cpu().numpy().astype(np.int64), lab_o)
    np.testing.assert_array_equal(counts.cpu().numpy(), np.bincount(lab_o, minlength=K))
    so = np.zeros((K, D))
    np.add.at(so, lab_o, Yh)
    assert parity.rel_err(sums.cpu().numpy(), so) < 1e-12
    print(f"hints={kind}: re-checked fraction {float(nre.item()) / n:.2e}")


def test_kmeans_tensor_path_ties_and_duplicates():
    from pmarlo_b200 import kernels

    rng = np.random.default_rng(0)
    c = rng.normal(size=(300, 6))
    c[5] = c[2]
    c[299] = c[0]
    Y = np.concatenate([c, 0.5 * (c[0] + c[1])[None], 0.5 * (c[10] + c[200])[None],
                        rng.normal(size=(3000, 6))]).astype(np.float32)
    cd = torch.from_numpy(c.astype(np.float32).astype(np.float64)).to(dev())
    lab = kernels.kmeans_assign(torch.from_numpy(Y).to(dev()), cd, impl=2).cpu().numpy()
    ref, _ = oracle.kmeans.assign(Y.astype(np.float64), c.astype(np.float32).astype(np.float64))
    np.testing.assert_array_equal(lab, ref)
    assert lab[5] == 2 and lab[299] == 0


@pytest.mark.parametrize("scale,offset", [(1.0, 0.0), (1e-3, 0.0), (50.0, 0.0), (1.0, 30.0)])
def test_kmeans_tensor_score_error_envelope(scale, offset):
    """The certainty test of kmeans_tc.cu assumes |score - exact d^2| <= 2^-19 (|y| + max|c|)^2.
    Measure the scores as they leave TMEM against fp64 (tolerance written here: a quarter of the envelope)."""
    from pmarlo_b200 import kernels

    rng = np.random.default_rng(7)
    n, D, K = 4096, 10, 1000
    Y = (scale * rng.normal(size=(n, D)) + offset).astype(np.float32)
    c = scale * rng.normal(size=(K, D)) + offset
    lab, sc = kernels.kmeans_tc_scores(torch.from_numpy(Y).to(dev()), torch.from_numpy(c).to(dev()))
    sc = sc.cpu().numpy().astype(np.float64)[:, :K]
    exact = oracle.kmeans.sqdist_direct(Y.astype(np.float64), c)
    yn = np.linalg.norm(Y.astype(np.float64), axis=1)[:, None]
    cmax = np.linalg.norm(c, axis=1).max()
    ratio = np.abs(sc - exact) / (2.0 ** -19 * (yn + cmax) ** 2)
    print(f"score error / envelope: max {ratio.max():.3f}, mean {ratio.mean():.4f}; "
          f"signed mean error / envelope {((sc - exact) / (2.0 ** -19 * (yn + cmax) ** 2)).mean():+.4f}")
    assert ratio.max() <= 0.25
    ref, _ = oracle.kmeans.assign(Y.astype(np.float64), c)
    np.testing.assert_array_equal(lab.cpu().numpy(), ref)


def test_assign_ties_duplicates_and_float64():
    """Duplicate centres and exact ties: the FIRST minimum wins, as np.argmin does."""
    from pmarlo_b200 import kernels

    rng = np.random.default_rng(0)
    c = rng.normal(size=(9, 4))
    c[5] = c[2]                                   # duplicate centre
    Y = np.concatenate([c, 0.5 * (c[0] + c[1])[None], rng.normal(size=(500, 4))])   # midpoint tie
    for dt in (np.float32, np.float64):
        Yd = torch.from_numpy(Y.astype(dt)).to(dev())
        cd = torch.from_numpy(c.astype(dt).astype(np.float64)).to(dev())
        lab = kernels.kmeans_assign(Yd, cd).cpu().numpy()
        ref, _ = oracle.kmeans.assign(Y.astype(dt).astype(np.float64), c.astype(dt).astype(np.float64))
        np.testing.assert_array_equal(lab, ref)
        assert lab[5] == 2 and lab[2] == 2


def test_assign_golden_matches_reference_discretizer(golden):
    """Labels of pmarlo's own _KMeansDiscretizer (sklearn predict, analysis/discretize.py:493) on its
    whitened inputs, reproduced from the committed centres."""
    from pmarlo_b200 import kernels

    z = golden("assign")
    for X, lab in ((z["X"], z["lab_train"]), (z["Xt"], z["lab_test"])):
        Xs = ((X - z["mean"]) / z["std"])
        got = kernels.kmeans_assign(torch.from_numpy(Xs).to(dev()), torch.from_numpy(z["centers"]).to(dev()))
        np.testing.assert_array_equal(got.cpu().numpy(), lab)


def test_cluster_microstates_api():
    from pmarlo_b200.clustering import cluster_microstates, cluster_microstates_labels

    rng = np.random.default_rng(4)
    blobs = rng.normal(size=(8, 3)) * 10
    Y = (blobs[rng.integers(0, 8, 4000)] + rng.normal(size=(4000, 3))).astype(np.float64)
    res = cluster_microstates(Y, method="kmeans", n_states=8, random_state=42)
    assert res.n_states == 8 and len(np.unique(res.labels)) == 8 and res.centers.shape == (8, 3)
    # reproducible under a fixed seed; explicit initial centres track the oracle bit-for-bit
    res2 = cluster_microstates(Y, method="kmeans", n_states=8, random_state=42)
    np.testing.assert_array_equal(res.labels, res2.labels)
    c0 = Y[:8].copy()
    r3 = cluster_microstates(Y, method="kmeans", n_states=8, initial_centers=c0)
    lo, co, nu, inertia, _ = oracle.kmeans.cluster_microstates(Y, 8, c0)
    np.testing.assert_array_equal(r3.labels, lo)
    np.testing.assert_allclose(r3.centers, co, rtol=1e-9, atol=1e-9)
    # n_init restarts keep the lowest inertia; labels-only wrapper
    r4 = cluster_microstates(Y, method="kmeans", n_states=8, random_state=1, n_init=3)
    assert len(np.unique(r4.labels)) == 8
    assert cluster_microstates_labels(Y, "kmeans", 8).shape == (4000,)
    # error behaviour of the reference
    assert cluster_microstates(np.empty((0, 3)), n_states=4).n_states == 0
    with pytest.raises(ValueError):
        cluster_microstates(Y[:, 0], n_states=3)
    with pytest.raises(ValueError):
        cluster_microstates(Y, n_states=0)
    with pytest.raises(TypeError):
        cluster_microstates(Y, n_states=3, bogus=1)
    with pytest.raises(TypeError):
        cluster_microstates(Y, method="kmeans", n_states=3, batch_size=10)
    with pytest.raises(ValueError):
        cluster_microstates(Y, n_states=3, n_init=2, fixed_seed=1)


# ----------------------------------------------------------------------------- counts
def test_counts_golden(golden):
    """Bit-exact against pmarlo's own _weighted_counts / _build_transition_counts outputs."""
    from pmarlo_b200 import kernels

    z = golden("counts")
    for ci in range(int(z["n_cases"])):
        g = lambda k: z[f"c{ci}_{k}"]  # noqa: E731
        labels = torch.from_numpy(g("labels")).to(dev())
        K, lag, stride = int(g("K")), int(g("lag")), int(g("stride"))
        off = torch.from_numpy(g("bounds")).to(dev())
        C = kernels.count_lagged(labels, off, K, lag, stride).cpu().numpy()
        np.testing.assert_array_equal(C, g("C_u").astype(np.int64))
        assert int(C.sum()) == int(g("tp_u"))
        whole = torch.tensor([0, labels.numel()], dtype=torch.int64, device=dev())
        np.testing.assert_array_equal(kernels.count_lagged(labels, whole, K, lag, 1).cpu().numpy(),
                                      g("C_all").astype(np.int64))
        np.testing.assert_array_equal(kernels.count_lagged(labels, off, K, lag, 1).cpu().numpy(),
                                      g("C_sl").astype(np.int64))
        np.testing.assert_array_equal(kernels.count_lagged(labels, off, K, lag, lag).cpu().numpy(),
                                      g("C_st").astype(np.int64))
        w = torch.from_numpy(g("weights")).to(dev())
        Cw = kernels.count_lagged_weighted(labels, w, off, K, lag, stride).cpu().numpy()
        np.testing.assert_allclose(Cw, g("C_w"), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("K,n,lag", [(5, 300000, 1), (100, 6_000_000, 10), (200, 50000, 100), (1000, 400000, 20),
                                     (5000, 300000, 20)])
def test_counts_vs_oracle(K, n, lag):
    from pmarlo_b200 import kernels
    from pmarlo_b200.msm import dtrajs_to_device

    dtrajs = synth.metastable_dtrajs(5, n // 5, K, seed=K) if n <= 400000 else None
    if dtrajs is None:   # big case: vectorised labels (python chain generation would take minutes)
        rng = np.random.default_rng(K)
        dtrajs = [rng.integers(0, K, size=n // 5).astype(np.int32) for _ in range(5)]
    dtrajs[1][::97] = -1
    dtrajs[2][5::101] = K + 3
    labels, segs = dtrajs_to_device(dtrajs, dev())
    C = kernels.count_lagged(labels, segs.device(dev()), K, lag).cpu().numpy()
    np.testing.assert_array_equal(C, oracle.counts.count_lagged(dtrajs, K, lag))


def test_count_transitions_api_split_mode():
    from pmarlo_b200.msm import count_transitions

    dtrajs = synth.metastable_dtrajs(3, 800, 6, seed=3)
    dtrajs[0][100] = -1
    dtrajs[2][::50] = 9
    C = count_transitions(dtrajs, n_states=6, lag=4, split_invalid=True)
    # labels >= n_states enlarge the matrix exactly as deeptime's estimator does (max label + 1)
    Co = oracle.counts.count_lagged(dtrajs, 10, 4, mode="split")
    np.testing.assert_array_equal(C, Co.astype(float))
    assert count_transitions([], lag=3).shape == (0, 0)


# ----------------------------------------------------------------------------- MSM
@pytest.mark.parametrize("K,stay", [(2, 0.9), (3, 0.5), (40, 0.9), (200, 0.95), (600, 0.9)])
def test_msm_vs_oracle(K, stay):
    from pmarlo_b200.msm import dtrajs_to_device

    dtrajs = synth.metastable_dtrajs(4, 20000 if K < 300 else 60000, K, seed=K + 1, stay=stay)
    labels, segs = dtrajs_to_device(dtrajs, dev())
    parity.check_counts_msm(labels, segs, K, 3, 5)


def test_two_state_chain_known_answer():
    """tests/unit/markov_state_model/test_two_state_msm.py:6-22: p=0.1 -> t2 = -1/ln 0.8 within 10 %."""
    from pmarlo_b200 import build_msm_from_labels
    from pmarlo_b200.msm import eigenvalues_rev, safe_timescales

    rng = np.random.default_rng(0)
    s = np.zeros(200000, dtype=np.int32)
    flips = rng.random(s.size) < 0.1
    for t in range(1, s.size):
        s[t] = s[t - 1] ^ int(flips[t])
    T, pi = build_msm_from_labels([s], n_states=2, lag=1)
    np.testing.assert_allclose(T.sum(axis=1), 1.0, atol=1e-12)
    np.testing.assert_allclose(pi @ T, pi, atol=1e-10)
    ev = eigenvalues_rev(T, pi, 2)
    t2 = safe_timescales(1, ev[1:])[0]
    assert abs(t2 - (-1.0 / np.log(0.8))) / (-1.0 / np.log(0.8)) < 0.1


def test_build_simple_msm_api():
    from pmarlo_b200 import build_simple_msm

    dtrajs = synth.metastable_dtrajs(3, 5000, 7, seed=9)
    dtrajs[0][::40] = -1                         # unassigned frames are ignored
    T, pi = build_simple_msm(dtrajs, n_states=9, lag=2)   # states 7, 8 never visited
    To, pio = oracle.msm.build_simple_msm(dtrajs, n_states=9, lag=2)
    assert T.shape == (9, 9)
    assert parity.rel_err(T, To) < 1e-6 and parity.rel_err(pi, pio) < 1e-6
    assert T[7, 7] == 1.0 and T[8, 8] == 1.0 and pi[7] == 0.0 and pi[8] == 0.0
    # detailed balance of the reversible estimate
    F = pi[:, None] * T
    np.testing.assert_allclose(F, F.T, atol=1e-12)
    Te, pe = build_simple_msm([], lag=3)
    assert Te.shape == (0, 0) and pe.shape == (0,)


def test_mle_batched_with_active_masks():
    """One CTA per problem: different lags, each restricted to its own connected set."""
    from pmarlo_b200 import kernels
    from pmarlo_b200.msm import dtrajs_to_device, largest_connected_set

    K = 30
    dtrajs = synth.metastable_dtrajs(3, 4000, K, seed=12, stay=0.97)
    labels, segs = dtrajs_to_device(dtrajs, dev())
    lags = [1, 5, 25, 100]
    Cs, acts = [], []
    for lag in lags:
        C = kernels.count_lagged(labels, segs.device(dev()), K, lag)
        Cs.append(C.to(torch.float64))
        Ch = C.cpu().numpy().copy()
        if lag == 25:                            # force a non-trivial mask: drop states 0 and 3
            Ch[[0, 3], :] = 0
            Ch[:, [0, 3]] = 0
        a = np.zeros(K, dtype=np.uint8)
        a[largest_connected_set(Ch)] = 1
        acts.append(a)
    assert acts[2][0] == 0 and acts[2].sum() >= 2
    act = torch.from_numpy(np.stack(acts)).to(dev())
    T, pi, info = kernels.mle_rev(torch.stack(Cs), act, alpha=0.0, maxerr=1e-12)
    for b in range(len(lags)):
        idx = np.flatnonzero(acts[b])
        Cb = Cs[b].cpu().numpy()[np.ix_(idx, idx)]
        if np.any(Cb.sum(axis=1) <= 0):
            assert int(info[b, 1].item()) == -1
            continue
        To, pio, _ = oracle.msm.mle_rev(Cb, maxerr=1e-12)
        Tf, pif = oracle.msm.expand_results(K, idx, To, pio)
        assert parity.rel_err(T[b].cpu().numpy(), Tf) < 1e-6
        assert parity.rel_err(pi[b].cpu().numpy(), pif) < 1e-6


def test_eigenvalues_lanczos_large_K():
    """K > 256 takes the Lanczos path; compare with eigvalsh of the oracle."""
    from pmarlo_b200 import kernels

    K = 700
    dtrajs = synth.metastable_dtrajs(4, 80000, K, seed=77, stay=0.9)
    C = oracle.counts.count_lagged(dtrajs, K, 2).astype(float)
    Ca, active = oracle.msm.ensure_connected_counts(C)
    To, pio, _ = oracle.msm.mle_rev(Ca, maxerr=1e-12)
    T = torch.from_numpy(To).to(dev())
    pi = torch.from_numpy(pio).to(dev())
    ev, info = kernels.eig_rev_topk(T, pi, 10)
    evo = oracle.msm.eigenvalues_rev(To, pio, 10)
    assert int(info[0].item()) > 0, "expected the Lanczos path"
    np.testing.assert_allclose(ev.cpu().numpy(), evo, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("K,k", [(2000, 10), (1500, 6)])
def test_eigenvalues_lanczos_narrow_bulk(K, k):
    """A reversible chain of five weakly coupled dense blocks: the bulk of the spectrum narrows like
    1 / sqrt(K), so orthogonality lost to the converged Perron vector grows ~50x per local Lanczos step.  The
    fixed re-orthogonalisation period this kernel once used returned eigenvalues of the order of 1e3 for
    K = 2000, k = 10 (tools/eig_dump.py); the partial re-orthogonalisation must agree with eigvalsh."""
    from pmarlo_b200 import kernels

    rng = np.random.default_rng(0)
    C = rng.random((K, K)) * 0.02
    b = K // 5
    for i in range(5):
        C[i * b:(i + 1) * b, i * b:(i + 1) * b] += rng.random((b, b))
    C = C + C.T
    T, pi = C / C.sum(1, keepdims=True), C.sum(1) / C.sum()
    sq = np.sqrt(pi)
    ref = np.linalg.eigvalsh(sq[:, None] * T / sq[None, :])
    ref = ref[np.argsort(-np.abs(ref))][:k]
    ev, info = kernels.eig_rev_topk(torch.from_numpy(T).to(dev()), torch.from_numpy(pi).to(dev()), k)
    assert int(info[0].item()) > 0 and int(info[1].item()) == 1, info
    np.testing.assert_allclose(ev.cpu().numpy(), ref, rtol=0, atol=1e-12)


def test_implied_timescales_sweep():
    from pmarlo_b200 import implied_timescales

    K = 20
    dtrajs = synth.metastable_dtrajs(4, 6000, K, seed=5, stay=0.95)
    lags = [1, 2, 5, 10, 20, 50]
    res = implied_timescales(dtrajs, lags, n_states=K, n_timescales=4, maxerr=1e-12)
    ref = oracle.msm.its_rev_mle(dtrajs, K, lags, 4, maxerr=1e-12)
    assert res.timescales.shape == ref.shape == (6, 4)
    ok = np.isfinite(ref)
    assert np.array_equal(np.isfinite(res.timescales), ok)
    np.testing.assert_allclose(res.timescales[ok], ref[ok], rtol=1e-5)


def test_safe_timescales_golden(golden):
    from pmarlo_b200 import safe_timescales

    z = golden("timescales")
    for lag in (1, 10, 400):
        np.testing.assert_array_equal(safe_timescales(lag, z["ev_real"]), z[f"ts_real_{lag}"])
        np.testing.assert_array_equal(safe_timescales(lag, z["ev_cplx"]), z[f"ts_cplx_{lag}"])


# ----------------------------------------------------------------------------- end to end
def test_pipeline_small_end_to_end():
    parity.run_and_check_small(seed=3, n_traj=5, n_frames=900, n_res=10, K=20, lag=4)


def test_pipeline_properties_at_scale():
    """Size-independent properties on a bench-sized shard (1.25 M frames x 99 atoms -> 256 features):
    pair bookkeeping, row-stochastic T, pi T = pi, detailed balance, labels in range, counts total."""
    from pmarlo_b200.pipeline import PipelineConfig, run_pipeline
    from pmarlo_b200.shards import Segments
    import bench

    wl = bench.make_workload(n_traj=4, frames_per_traj=50_000, device=dev(), seed=4)
    cfg = bench.bench_config(n_states=300, kmeans_iters=3)
    res = run_pipeline(wl.xyz, wl.segs, wl.plan, cfg)
    K = cfg.n_states
    C = res.counts.cpu().numpy()
    assert int(C.sum()) == wl.segs.n_pairs(cfg.msm_lag)
    lab = res.labels.cpu().numpy()
    assert lab.min() >= 0 and lab.max() < K
    T, pi = res.T.cpu().numpy(), res.pi.cpu().numpy()
    np.testing.assert_allclose(T.sum(axis=1), 1.0, atol=1e-12)
    np.testing.assert_allclose(pi @ T, pi, atol=1e-10)
    F = pi[:, None] * T
    np.testing.assert_allclose(F, F.T, atol=1e-12)
    ev = res.eigenvalues.cpu().numpy()
    assert abs(ev[0] - 1.0) < 1e-9 and np.all(np.abs(ev) <= 1 + 1e-9)
    # labels are the exact argmin for the final centres (checked on a slice with the oracle)
    sl = slice(0, 20000)
    lo, _ = oracle.kmeans.assign(res.Y[sl].cpu().numpy().astype(np.float64), res.centers.cpu().numpy())
    np.testing.assert_array_equal(lab[sl], lo)


def test_pipeline_properties_at_full_size():
    """The bench configuration itself (C4: 10 M frames x 99 atoms -> 256 features -> 10 TICA dims, K = 1000,
    20 Lloyd iterations): size-independent properties of every stage's result, plus oracle spot checks on
    slices the CPU can afford."""
    from pmarlo_b200.pipeline import run_pipeline
    import bench

    free, _ = torch.cuda.mem_get_info()
    if free < 60e9:
        pytest.skip("needs ~40 GB of HBM")
    n_traj, fpt = 80, 125_000
    wl = bench.make_workload(n_traj=n_traj, frames_per_traj=fpt, device=dev(), seed=4000)
    cfg = bench.bench_config()
    res = run_pipeline(wl.xyz, wl.segs, wl.plan, cfg)
    K, n = cfg.n_states, n_traj * fpt
    # counts: every (t, t + lag) pair inside a trajectory is counted exactly once
    C = res.counts.cpu().numpy()
    assert int(C.sum()) == n_traj * (fpt - cfg.msm_lag) == wl.segs.n_pairs(cfg.msm_lag)
    lab = res.labels.cpu().numpy()
    assert lab.shape == (n,) and lab.min() >= 0 and lab.max() < K
    # row sums of the count matrix = population of each state over the frames that start a pair
    starts = np.ones(n, dtype=bool)
    starts.reshape(n_traj, fpt)[:, fpt - cfg.msm_lag:] = False
    np.testing.assert_array_equal(C.sum(axis=1), np.bincount(lab[starts], minlength=K))
    ends = np.ones(n, dtype=bool)
    ends.reshape(n_traj, fpt)[:, :cfg.msm_lag] = False
    np.testing.assert_array_equal(C.sum(axis=0), np.bincount(lab[ends], minlength=K))
    # reversible MSM: stochastic, stationary, detailed balance; spectrum inside the unit disc
    T, pi = res.T.cpu().numpy(), res.pi.cpu().numpy()
    np.testing.assert_allclose(T.sum(axis=1), 1.0, atol=1e-12)
    np.testing.assert_allclose(pi.sum(), 1.0, atol=1e-12)
    np.testing.assert_allclose(pi @ T, pi, atol=1e-10)
    F = pi[:, None] * T
    np.testing.assert_allclose(F, F.T, atol=1e-12)
    ev = res.eigenvalues.cpu().numpy()
    assert abs(ev[0] - 1.0) < 1e-9 and np.all(np.abs(ev) <= 1 + 1e-9) and np.all(np.diff(np.abs(ev)) <= 1e-12)
    # labels are the exact fp64 argmin for the final centres; the assignment is idempotent
    centers = res.centers.cpu().numpy()
    for sl in (slice(0, 20000), slice(n - 20000, n), slice(5_000_000, 5_020_000)):
        lo, _ = oracle.kmeans.assign(res.Y[sl].cpu().numpy().astype(np.float64), centers)
        np.testing.assert_array_equal(lab[sl], lo)
    from pmarlo_b200 import kernels
    again = kernels.kmeans_assign(res.Y, res.centers, hints=res.labels.clone())
    assert torch.equal(again, res.labels)
    ts = res.timescales
    assert ts is None or np.all(np.asarray(ts)[:-1] >= np.asarray(ts)[1:])


# ----------------------------------------------------------------------------- discretize_dataset (a7)
def _discretize_dataset_inputs(z):
    seg = [int(v) for v in z["seg"]]
    names = {"names": ["a", "b", "c", "d"], "n_features": 4}
    return {"splits": {
        "train": {"X": z["Xtr"], "feature_schema": names, "segments": [{"length": L, "stride": 1} for L in seg]},
        "test": {"X": z["Xte"], "feature_schema": names, "segments": [{"length": 500}]},
    }}


@pytest.mark.parametrize("tag", ["plain", "weighted"])
def test_discretize_dataset_matches_reference_given_its_centres(golden, tag):
    """pmarlo.analysis.discretize.discretize_dataset run in the reference (tests/golden/make_golden.py):
    with the reference's fitted centres injected, assignments, pruning, counts, T and the pair bookkeeping
    are reproduced (labels and unweighted counts bit-exactly)."""
    from pmarlo_b200.discretize import discretize_dataset

    z = golden("discretize")
    kw = dict(frame_weights={"train": z["w"]}, min_out_count=2) if tag == "weighted" else {}
    r = discretize_dataset(_discretize_dataset_inputs(z), cluster_mode="kmeans", n_microstates=9, lag_time=3,
                           random_state=0, centers=z[f"{tag}_centers"], **kw)
    np.testing.assert_array_equal(r.assignments["train"], z[f"{tag}_lab_train"])
    np.testing.assert_array_equal(r.assignments["test"], z[f"{tag}_lab_test"])
    np.testing.assert_array_equal(r.pruned_state_indices, z[f"{tag}_pruned"])
    assert r.counted_pairs["train"] == int(z[f"{tag}_counted_pairs"])
    assert r.expected_pairs["train"] == int(z[f"{tag}_expected_pairs"])
    if tag == "plain":
        np.testing.assert_array_equal(r.counts, z["plain_counts"])
        np.testing.assert_array_equal(r.counts_before_prune, z["plain_counts_before_prune"])
        np.testing.assert_array_equal(r.transition_matrix, z["plain_T"])
    else:
        np.testing.assert_allclose(r.counts, z["weighted_counts"], rtol=1e-13, atol=0)
        np.testing.assert_allclose(r.counts_before_prune, z["weighted_counts_before_prune"], rtol=1e-13, atol=0)
        np.testing.assert_allclose(r.transition_matrix, z["weighted_T"], rtol=1e-13, atol=1e-16)
    np.testing.assert_allclose(r.state_counts, z[f"{tag}_state_counts"], rtol=1e-13)
    assert abs(r.diag_mass - float(z[f"{tag}_diag_mass"])) < 1e-13
    assert r.assignments["train"].dtype == np.int32 and r.counts.dtype == np.float64


def test_discretize_dataset_fit_path_and_errors(golden):
    from pmarlo_b200.discretize import discretize_dataset

    z = golden("discretize")
    ds = _discretize_dataset_inputs(z)
    r = discretize_dataset(ds, n_microstates=9, lag_time=3, random_state=0, n_init=3)
    # own fit: different centres than sklearn's, same invariants
    n = r.counts.shape[0]
    assert r.centers.shape == (9, 4) and 1 <= n <= 9
    assert r.counted_pairs["train"] == int(r.counts.sum()) <= r.expected_pairs["train"]
    rows = r.transition_matrix.sum(axis=1)
    assert np.all((np.abs(rows - 1.0) < 1e-12) | (rows == 0.0))
    assert set(r.assignments) == {"train", "test"} and r.assignments["test"].max() < n
    r2 = discretize_dataset(ds, n_microstates=9, lag_time=3, random_state=0, n_init=3)
    np.testing.assert_array_equal(r.assignments["train"], r2.assignments["train"])     # seeded -> reproducible
    with pytest.raises(ValueError):
        discretize_dataset(ds, lag_time=0)
    with pytest.raises(ValueError):
        discretize_dataset({"splits": {}})
    with pytest.raises(NotImplementedError):
        discretize_dataset(ds, cluster_mode="grid")
    bad = {"splits": {"train": ds["splits"]["train"], "test": {"X": z["Xte"][:, :3]}}}
    with pytest.raises(ValueError):
        discretize_dataset(bad, n_microstates=9, centers=z["plain_centers"])
    # a plain (n, d) array is a dataset with one split called "all"
    r3 = discretize_dataset(z["Xtr"], n_microstates=5, lag_time=2, random_state=1, n_init=1)
    assert list(r3.assignments) == ["all"] and r3.segment_lengths["all"] == [z["Xtr"].shape[0]]


# ----------------------------------------------------------------------------- EnhancedMSM interface (8b)
def test_enhanced_msm_build_and_its_match_oracle():
    from pmarlo_b200 import EnhancedMSM

    rng = np.random.default_rng(7)
    P = np.array([[0.92, 0.06, 0.02, 0.0], [0.05, 0.9, 0.05, 0.0], [0.02, 0.08, 0.9, 0.0], [0.0, 0.0, 0.0, 1.0]])
    dtrajs = []
    for n in (4000, 2500, 3000):
        s = np.empty(n, dtype=np.int64)
        s[0] = rng.integers(0, 3)
        for t in range(1, n):
            s[t] = rng.choice(4, p=P[s[t - 1]])
        dtrajs.append(s)
    dtrajs[1][100:103] = -1                       # unassigned frames split the trajectory
    m = EnhancedMSM(dtrajs, n_states=5)           # state 3, 4 never visited
    m.build_msm(lag_time=4)
    # oracle: split at invalid labels, sliding counts, +alpha on the active block, row normalisation
    parts = []
    for d in dtrajs:
        ok = (d >= 0) & (d < 5)
        edges = np.flatnonzero(np.diff(np.concatenate([[0], ok.astype(np.int8), [0]])))
        parts += [d[a:b] for a, b in zip(edges[::2], edges[1::2])]
    C = oracle.counts.count_lagged(parts, 5, 4).astype(float)
    Ca, act = oracle.msm.ensure_connected_counts(C)
    Ta = oracle.msm.transition_matrix_nonrev(Ca)
    T = np.eye(5)
    T[np.ix_(act, act)] = Ta
    pi = np.zeros(5)
    pi[act] = oracle.msm.stationary_distribution(Ta)
    np.testing.assert_allclose(m.transition_matrix, T, rtol=0, atol=1e-14)
    np.testing.assert_allclose(m.stationary_distribution, pi, rtol=1e-9, atol=1e-14)
    cm = np.zeros((5, 5))
    cm[np.ix_(act, act)] = Ca
    np.testing.assert_array_equal(m.count_matrix, cm)
    assert m.stationary_distribution[3] == 0.0 and m.transition_matrix[4, 4] == 1.0
    assert m.free_energies.min() == 0.0
    # ITS: per-lag reversible MLE on the largest connected set (ck_its_selector.py:397-399)
    m.compute_implied_timescales([1, 2, 4, 8, 100000], n_timescales=2, plateau_m=2, plateau_epsilon=0.5, estimator="mle")
    its = m.implied_timescales
    np.testing.assert_array_equal(its.lag_times, [1, 2, 4, 8])
    ref = oracle.msm.its_rev_mle(dtrajs, 5, [1, 2, 4, 8], 2)
    np.testing.assert_allclose(its.timescales, ref, rtol=1e-6, equal_nan=True)
    assert its.timescales_ci.shape == (4, 2, 2) and np.all(np.isnan(its.timescales_ci))
    np.testing.assert_allclose(its.rates, 1.0 / ref, rtol=1e-6, equal_nan=True)
    # reference error / empty behaviour
    with pytest.raises(ValueError):
        EnhancedMSM([], n_states=0).build_msm()
    with pytest.raises(ValueError):
        m.build_msm(method="tram")
    e = EnhancedMSM([], n_states=3)
    e.compute_implied_timescales()
    assert e.implied_timescales.lag_times.size == 0 and e.implied_timescales.timescales.shape == (0, 5)
    short = EnhancedMSM([np.array([0, 1])], n_states=2)
    short.build_msm(lag_time=20)                   # lag capped at 1
    assert short.transition_matrix.shape == (2, 2)


# ----------------------------------------------------------------------------- Chapman-Kolmogorov test (8f-1)
def test_relabel_compact_kernel_is_exact():
    """pmb_relabel_compact against the reference's list comprehension, shard by shard (bit-exact)."""
    from pmarlo_b200 import kernels

    rng = np.random.default_rng(7)
    cases = [
        ([5000, 0, 4097, 1, 8192, 3], 40, 0.3),      # empty shard, tile-boundary lengths
        ([100], 7, 0.0),                              # nothing dropped
        ([4096 * 3], 12, 1.0),                        # everything dropped
        ([70001, 129999, 50000], 1000, 0.5),
    ]
    for lens, K, p_drop in cases:
        lab = rng.integers(-1, K, size=sum(lens)).astype(np.int32)     # -1 = unassigned frame
        lut = np.where(rng.random(K) < p_drop, -1, rng.permutation(K)).astype(np.int32)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        out, new_off = kernels.relabel_compact(torch.from_numpy(lab).to(dev()), torch.from_numpy(off).to(dev()),
                                               torch.from_numpy(lut).to(dev()))
        new_off = new_off.cpu().numpy()
        out = out.cpu().numpy()
        for i in range(len(lens)):
            traj = lab[off[i]:off[i + 1]]
            want = np.array([lut[s] for s in traj if s >= 0 and lut[s] >= 0], dtype=np.int32)
            got = out[new_off[i]:new_off[i + 1]]
            assert np.array_equal(got, want), (lens, i)
    # a label beyond the table is dropped, not read out of bounds
    out, new_off = kernels.relabel_compact(torch.tensor([0, 9, 1, 2], dtype=torch.int32, device=dev()),
                                           torch.tensor([0, 4], dtype=torch.int64, device=dev()),
                                           torch.tensor([3, 4, -1], dtype=torch.int32, device=dev()))
    assert new_off.cpu().tolist() == [0, 2] and out[:2].cpu().tolist() == [3, 4]


def test_ck_matches_reference_golden(golden, tmp_path):
    from pmarlo_b200 import ck

    z = golden("ck")
    for name, dtrajs, kw in parity.ck_cases(z):
        parity.check_ck_case(z, name, dtrajs, kw, ck.run_ck, ck.compute_ck_test_micro, ck.select_lag_time_ck)
    r = ck.run_ck([np.array([0, 1, 2] * 1000)], 1, tmp_path, macro_k=3, min_trans=5, top_n_micro=3)
    assert sorted(r.mse) == [2, 3, 4, 5] and (tmp_path / "ck_mse.json").exists()
    with pytest.raises(ValueError, match="strictly positive row sums"):
        ck.run_ck([np.array([0, 1] * 200 + [2])], 1, None, min_trans=1)


def test_ck_enhanced_msm_methods(tmp_path, capsys):
    """tests/unit/markov_state_model/test_ck.py:16-33 and test_ck_tau_selection.py:14-28."""
    from pmarlo_b200 import EnhancedMSM

    m = EnhancedMSM([np.array([0, 1, 2] * 1000)], n_states=3, output_dir=str(tmp_path))
    m.lag_time = 1
    res = m.compute_ck_test_micro()
    assert sorted(res.mse) == [2, 3, 4, 5] and all(v >= 0.0 for v in res.mse.values())
    assert res.mse[2] < 1e-6 and not res.insufficient_data
    m2 = EnhancedMSM([np.array([0, 0, 1, 1] * 500)], n_states=2, output_dir=str(tmp_path))
    assert m2.select_lag_time_ck([1, 2, 3]) == 2 and m2.lag_time == 2
    assert (tmp_path / "ck_mse.csv").exists() and "Selected τ =" in capsys.readouterr().out


def test_ck_at_scale_against_oracle():
    """2 M labels, 300 states (some rare, some never visited): the device path and the numpy oracle agree on
    every MSE, and a label shard already in HBM gives the same answer as host arrays."""
    from pmarlo_b200 import ck

    rng = np.random.default_rng(12)
    K, n_traj, L = 300, 8, 250_000
    P = rng.random((K, K)) ** 6 + 1e-6
    P[:, 250:] *= 1e-4                         # rare states
    P[:, 290:] = 0                             # never visited
    P /= P.sum(axis=1, keepdims=True)
    cum = np.cumsum(0.5 * np.eye(K) + 0.5 * P, axis=1)
    dtrajs = []
    for _ in range(n_traj):                    # vectorised over time is impossible; over blocks of chains is
        s = np.empty(L, dtype=np.int64)
        s[0] = rng.integers(0, 250)
        u = rng.random(L)
        for t in range(1, L):
            s[t] = min(int(np.searchsorted(cum[s[t - 1]], u[t])), K - 1)
        dtrajs.append(s)
    kw = dict(lag_time=5, min_trans=20, top_n_micro=40, factors=(2, 3, 4))
    ref = oracle.ck.run_ck(dtrajs, **kw)
    got = ck.run_ck(dtrajs, output_dir=None, **kw)
    assert got.mode == ref.mode and sorted(got.mse) == sorted(ref.mse) and got.mse
    assert got.insufficient_k == ref.insufficient_k
    np.testing.assert_allclose([got.mse[k] for k in sorted(got.mse)], [ref.mse[k] for k in sorted(ref.mse)], rtol=1e-9)
    shard = ck.LabelShard.from_dtrajs(dtrajs)
    again = ck.run_ck(shard, output_dir=None, **kw)
    assert again.mse == got.mse


def test_ck_lag_selector_matches_reference_golden(golden):
    """select_optimal_lag_ck_its on the device kernels against the vectors produced by the reference's
    ck_its_selector.py (reversible-MLE quantities to 1e-6, everything else to rounding)."""
    from pmarlo_b200 import ck

    z = golden("ck_selector")
    for name, dtrajs, kw, lumper in parity.selector_cases(z):
        parity.check_selector_case(z, name, dtrajs, kw, lumper, ck.select_optimal_lag_ck_its, mle_rtol=1e-6)


def test_macro_helpers_on_device(golden):
    from pmarlo_b200 import macro

    z = golden("macro")
    np.testing.assert_allclose(macro.lump_micro_to_macro_T(z["T"], z["pi"], z["lab"]), z["Tm"], rtol=1e-12)


# ----------------------------------------------------------------------------- deterministic ITS fallback (a12)
def test_deterministic_its_from_counts_matches_reference_golden(golden):
    """_its.py:742-801 run from the reference file itself (deeptime's two analysis functions stood in for by
    their numpy definitions, see make_golden.py::make_ladders): eigenvalues, timescales, rates."""
    from pmarlo_b200 import deterministic_its_from_counts

    z = golden("ladders")
    for i, (K, n_ts) in enumerate(z["det_cases"]):
        ev, ts, rates = deterministic_its_from_counts(7, z[f"det_C_{i}"], int(n_ts))
        assert ev.shape == ts.shape == rates.shape == (int(n_ts),)
        np.testing.assert_allclose(ev, z[f"det_ev_{i}"], rtol=1e-9, atol=1e-12, err_msg=f"case {i}")
        np.testing.assert_allclose(ts, z[f"det_ts_{i}"], rtol=1e-7, equal_nan=True, err_msg=f"case {i}")
        np.testing.assert_allclose(rates, z[f"det_rate_{i}"], rtol=1e-7, equal_nan=True, err_msg=f"case {i}")


# ----------------------------------------------------------------------------- VAMP reduction (boundary: reduce_features("vamp"))
@pytest.mark.parametrize("d,lag,m,scale", [(6, 10, 3, True), (12, 3, 2, False), (40, 5, 8, True)])
def test_vamp_reduce_vs_oracle(d, lag, m, scale):
    """reduce_features(X, "vamp", lag, n_components) -- the call api/conformations.py:195 makes -- against the
    oracle's restatement of deeptime VAMP on the same fp32-representable input."""
    from pmarlo_b200 import reduce_features, vamp_reduce

    X = np.concatenate(synth.ar1_features(1, 6000, d, seed=d + lag), axis=0).astype(np.float32).astype(np.float64)
    if d == 12:
        X[100, 3] = np.nan
        X[2500, 0] = np.nan
    got = vamp_reduce(X, lag=lag, n_components=m, scale=scale)
    ref = oracle.tica.vamp_reduce(X, lag=lag, n_components=m, scale=scale)
    assert got.shape == ref.shape == (6000, m) and got.dtype == np.float64
    # singular functions inherit the covariance error (fp32 products with fp64 folds, ~1e-7) amplified by
    # cond(C00) / singular-value gap, like the TICA projection in parity.check_tica
    Xp = oracle.tica.preprocess(X, scale=scale)
    _, _, sv, _, _, (C00, _, _) = oracle.tica.vamp_fit([Xp], lag, m, 1e-6)
    gap = float(np.min(np.abs(np.diff(sv[: m + 1])))) if sv.size > m else float(np.min(np.abs(np.diff(sv[:m]))))
    s00 = np.linalg.eigvalsh(C00)
    cond = float(s00.max() / max(s00[s00 > 1e-6].min(), 1e-6))
    tol = min(2e-3, max(2e-5, 4.0 * 2e-7 * cond / max(gap, 1e-12)))
    assert parity.rel_err(got, ref) <= tol, (parity.rel_err(got, ref), tol, cond, gap)
    # basis-independent check: the projected coordinates are whitened exactly like the oracle's
    cg = np.cov(got.T, bias=True) if m > 1 else np.array([[np.var(got)]])
    cr = np.cov(ref.T, bias=True) if m > 1 else np.array([[np.var(ref)]])
    assert float(np.max(np.abs(cg - cr))) <= 1e-4
    if scale and lag == 10:
        np.testing.assert_allclose(reduce_features(X, "vamp", lag=10, n_components=m), got, rtol=0, atol=0)
    with pytest.raises(ValueError):
        vamp_reduce(X[:3], lag=5)


# ----------------------------------------------------------------------------- silhouette / n_states="auto" (a6)
@pytest.mark.parametrize("n,D,K", [(700, 2, 4), (3000, 5, 11), (1500, 10, 20), (130, 3, 7)])
def test_silhouette_samples_match_sklearn(n, D, K):
    from sklearn.metrics import silhouette_samples, silhouette_score

    from pmarlo_b200 import kernels
    from pmarlo_b200.clustering import silhouette_score_device

    rng = np.random.default_rng(n + D)
    cen = rng.normal(scale=3.0, size=(K, D))
    lab = rng.integers(0, K, size=n)
    lab[:K] = np.arange(K)
    if n == 130:
        lab[lab == 3] = 2
        lab[0] = 3                     # a singleton cluster: coefficient 0
    Y = cen[lab] + rng.normal(size=(n, D))
    Yd = torch.from_numpy(Y).cuda()
    ld = torch.from_numpy(lab.astype(np.int32)).cuda()
    got = kernels.silhouette_samples(Yd, ld, K).cpu().numpy()
    ref = silhouette_samples(Y, lab)
    np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-12)
    assert abs(silhouette_score_device(Yd, ld, K) - silhouette_score(Y, lab)) <= 1e-12


def test_cluster_microstates_auto_n_states():
    """n_states="auto": the silhouette scan picks the planted number of well-separated clusters; the scores it
    used are sklearn's for the labellings it produced."""
    from pmarlo_b200 import cluster_microstates
    from pmarlo_b200.clustering import auto_select_n_states

    rng = np.random.default_rng(4)
    cen = np.array([[0, 0], [10, 0], [0, 10], [10, 10], [5, 20], [20, 5]], dtype=float)
    Y = np.concatenate([c + 0.4 * rng.normal(size=(400, 2)) for c in cen])
    res = cluster_microstates(Y, method="kmeans", n_states="auto", random_state=7)
    assert res.n_states == 6 and res.rationale.startswith("silhouette=")
    assert np.unique(res.labels).size == 6
    chosen, rationale, scores = auto_select_n_states(torch.from_numpy(Y).cuda(), 7, return_scores=True)
    assert chosen == 6 and [k for k, _ in scores] == list(range(4, 21))
    assert max(scores, key=lambda x: x[1])[0] == 6
    sub = cluster_microstates(Y, method="kmeans", n_states="auto", random_state=7, silhouette_sample_size=500)
    assert sub.rationale.endswith("sample=500") and sub.n_states == 6
    over = cluster_microstates(Y, n_states="auto", auto_n_states_override=9)
    assert over.n_states == 9 and over.rationale == "auto-override=9"
    with pytest.raises(ValueError):
        cluster_microstates(Y, n_states="auto", silhouette_sample_size=1)
