"""Generate the committed golden fixtures from the importable parts of pmarlo.

Run once in the build container (needs /root/reference, which does NOT exist
on the GPU box):

    python tests/golden/make_golden.py

It imports the reference functions that run without deeptime/mdtraj
 and stores seeded inputs together with the reference outputs:

* counts.npz      -- analysis.discretize._weighted_counts (discretize.py:609-645),
                     analysis.debug_export._build_transition_counts
                     (debug_export.py:385-409), analysis.counting.expected_pairs
* timescales.npz  -- markov_state_model.utils.safe_timescales (utils.py:17-57)
* preprocess.npz  -- markov_state_model.reduction._preprocess (reduction.py:13-40)
* tica_xcheck.npz -- features.deeptica.core.trainer_api._estimate_top_eigenvalues
                     (trainer_api.py:632-656), a non-symmetrised numpy TICA
                     used only as a loose cross-check of the oracle
* assign.npz      -- analysis.discretize._KMeansDiscretizer fit/transform
                     (discretize.py:406-514; sklearn KMeans.predict labels)
* ck.npz          -- markov_state_model.ck_runner.run_ck (ck_runner.py:293-332) and
                     CKMixin.compute_ck_test_micro / select_lag_time_ck (_ck.py:61-228)
* ck_selector.npz -- markov_state_model.ck_its_selector.select_optimal_lag_ck_its (ck_its_selector.py:279-599)
* macro.npz       -- _msm_utils.compute_macro_populations / lump_micro_to_macro_T / compute_macro_mfpt
* topologies.npz  -- atom names / residue ids / coordinates (nm) parsed from
                     data/alanine-dipeptide.pdb and data/chignolin.pdb (model 1)
"""

from __future__ import annotations

import pathlib
import sys
import types

import numpy as np

REF = pathlib.Path("/root/reference")
sys.path.insert(0, str(REF / "src"))
OUT = pathlib.Path(__file__).resolve().parent


def _parse_pdb(path: pathlib.Path, first_model_only: bool = True):
    names, resn, resid, chain, xyz = [], [], [], [], []
    last_key, ridx = None, -1
    for line in path.read_text().splitlines():
        if line.startswith("ENDMDL") and first_model_only:
            break
        if not line.startswith(("ATOM", "HETATM")):
            continue
        name = line[12:16].strip()
        rn = line[17:20].strip()
        ch = line[21].strip() or "A"
        rs = line[22:27].strip()
        key = (ch, rs, rn)
        if key != last_key:
            ridx += 1
            last_key = key
        names.append(name)
        resn.append(rn)
        resid.append(ridx)
        chain.append(0)
        xyz.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
    return (np.array(names), np.array(resn), np.array(resid, dtype=np.int32),
            np.array(chain, dtype=np.int32), (np.array(xyz, dtype=np.float64) / 10.0).astype(np.float32))


def make_topologies():
    a = _parse_pdb(REF / "data" / "alanine-dipeptide.pdb")
    c = _parse_pdb(REF / "data" / "chignolin.pdb")
    np.savez_compressed(
        OUT / "topologies.npz",
        ala2_names=a[0], ala2_resn=a[1], ala2_resid=a[2], ala2_chain=a[3], ala2_xyz=a[4],
        chig_names=c[0], chig_resn=c[1], chig_resid=c[2], chig_chain=c[3], chig_xyz=c[4],
    )
    print("topologies:", a[4].shape, c[4].shape)


def make_counts():
    from pmarlo.analysis.counting import expected_pairs
    from pmarlo.analysis.debug_export import _build_transition_counts
    from pmarlo.analysis.discretize import _weighted_counts

    rng = np.random.default_rng(20260518)
    out = {}
    cases = []
    for ci, (n, K, lag, stride, pneg) in enumerate(
        [(500, 7, 1, 1, 0.0), (2000, 13, 5, 1, 0.05), (3000, 50, 17, 3, 0.1),
         (64, 3, 70, 1, 0.0), (1500, 9, 2, 2, 0.3), (1, 2, 1, 1, 0.0)]):
        labels = rng.integers(0, K, size=n).astype(np.int32)
        labels[rng.random(n) < pneg] = -1
        cuts = np.sort(rng.choice(np.arange(1, max(2, n)), size=min(4, max(0, n - 1)), replace=False)) if n > 1 else np.array([], dtype=int)
        bounds = np.concatenate([[0], cuts, [n]]).astype(np.int64)
        segments = [(int(bounds[i]), int(bounds[i + 1])) for i in range(len(bounds) - 1)]
        weights = rng.random(n)
        C_u, tp_u = _weighted_counts(labels, n_states=K, lag_time=lag, segments=segments, stride=stride)
        C_w, tp_w = _weighted_counts(labels, n_states=K, lag_time=lag, weights=weights, segments=segments, stride=stride)
        C_all, tp_all = _weighted_counts(labels, n_states=K, lag_time=lag)
        dtrajs = [labels[s:e] for s, e in segments]
        C_sl, n_sl = _build_transition_counts(dtrajs, K, lag, "sliding")
        C_st, n_st = _build_transition_counts(dtrajs, K, lag, "strided")
        ep = expected_pairs([e - s for s, e in segments], lag, stride)
        for k, v in dict(labels=labels, bounds=bounds, weights=weights, K=K, lag=lag, stride=stride,
                         C_u=C_u, tp_u=tp_u, C_w=C_w, tp_w=tp_w, C_all=C_all, tp_all=tp_all,
                         C_sl=C_sl, n_sl=n_sl, C_st=C_st, n_st=n_st, ep=ep).items():
            out[f"c{ci}_{k}"] = np.asarray(v)
        cases.append(ci)
    out["n_cases"] = np.asarray(len(cases))
    np.savez_compressed(OUT / "counts.npz", **out)
    print("counts: cases", len(cases))


def make_timescales():
    from pmarlo.markov_state_model.utils import safe_timescales

    ev_real = np.array([1.0, 0.999999999999, 0.99, 0.8, 0.5, 1e-3, 1e-13, 0.0, -0.2, 1.2, np.nan, np.inf])
    ev_cplx = np.array([0.9 + 0.0j, 0.5 + 0.3j, -0.4 + 0.0j, 0.0 + 0.0j, 0.7 - 0.7j, 1.0 + 0.0j])
    out = {"ev_real": ev_real, "ev_cplx": ev_cplx}
    for lag in (1, 10, 400):
        out[f"ts_real_{lag}"] = safe_timescales(lag, ev_real)
        out[f"ts_cplx_{lag}"] = safe_timescales(lag, ev_cplx)
    out["ts_empty"] = safe_timescales(5, np.array([]))
    np.savez_compressed(OUT / "timescales.npz", **out)
    print("timescales ok")


def make_preprocess():
    from pmarlo.markov_state_model.reduction import _preprocess

    rng = np.random.default_rng(7)
    X = rng.standard_normal((400, 6)) * np.array([1.0, 5.0, 0.01, 100.0, 1.0, 1.0]) + np.array([0, 3, -2, 50, 0, 1.0])
    X[:, 4] = 2.5  # constant column
    Xn = X.copy()
    Xn[rng.random(X.shape) < 0.03] = np.nan
    Xn[:, 4] = 2.5
    np.savez_compressed(
        OUT / "preprocess.npz", X=X, Xn=Xn,
        P_scale=_preprocess(X, scale=True), P_noscale=_preprocess(X, scale=False),
        Pn_scale=_preprocess(Xn, scale=True), Pn_noscale=_preprocess(Xn, scale=False),
        P_1d=_preprocess(X[:, 1], scale=True),
    )
    print("preprocess ok")


def make_tica_xcheck():
    from pmarlo.features.deeptica.core.trainer_api import _estimate_top_eigenvalues

    rng = np.random.default_rng(11)
    N, rho, lag = 60000, np.array([0.97, 0.9, 0.6, 0.2]), 4
    x = np.zeros((N, 4))
    e = rng.standard_normal((N, 4))
    for t in range(1, N):
        x[t] = rho * x[t - 1] + np.sqrt(1 - rho ** 2) * e[t]
    A = rng.standard_normal((4, 4))
    X = (x @ A).astype(np.float64)
    idx_t = np.arange(0, N - lag)
    idx_tau = idx_t + lag
    cfg = types.SimpleNamespace(n_out=4)
    ev = np.asarray(_estimate_top_eigenvalues(X, idx_t, idx_tau, cfg))
    np.savez_compressed(OUT / "tica_xcheck.npz", X=X.astype(np.float32), lag=lag, rho=rho, ref_eigs=ev)
    print("tica xcheck eigs", ev, rho ** lag)


def make_assign():
    from pmarlo.analysis.discretize import _KMeansDiscretizer

    rng = np.random.default_rng(3)
    centers0 = rng.standard_normal((12, 5)) * 4
    X = centers0[rng.integers(0, 12, 3000)] + rng.standard_normal((3000, 5))
    Xt = centers0[rng.integers(0, 12, 1000)] + rng.standard_normal((1000, 5))
    disc = _KMeansDiscretizer(12, random_state=0, apply_whitening=True)
    disc.fit(X)
    lab_train = disc.transform(X)
    lab_test = disc.transform(Xt)
    np.savez_compressed(
        OUT / "assign.npz", X=X, Xt=Xt, mean=disc.scaler_mean_, std=disc.scaler_std_,
        centers=disc.centers, lab_train=lab_train, lab_test=lab_test,
    )
    print("assign ok", np.bincount(lab_train))


def make_discretize():
    """analysis.discretize.discretize_dataset (discretize.py:901-1120) end to end: whitening, k-means fit
    (sklearn, not reproducible elsewhere -> the fitted centres are stored), per-split assignment,
    weighted lagged counts over segments, zero-row pruning, row-normalised T."""
    from pmarlo.analysis.discretize import discretize_dataset

    rng = np.random.default_rng(11)
    K = 9
    centers0 = rng.standard_normal((K, 4)) * 3
    centers0[K - 1] = 40.0            # far away: k-means gives the sink its own cluster
    # a jump process between the blobs; blob K-1 is only ever entered at the very end of a segment,
    # so its row of the count matrix is empty and the reference prunes it
    def traj(n, last_to_sink):
        s = np.empty(n, dtype=int)
        s[0] = rng.integers(0, K - 1)
        for t in range(1, n):
            s[t] = s[t - 1] if rng.random() < 0.8 else rng.integers(0, K - 1)
        if last_to_sink:
            s[-3:] = K - 1
        return centers0[s] + 0.3 * rng.standard_normal((n, 4))
    seg = [400, 250, 350]
    Xtr = np.concatenate([traj(seg[0], True), traj(seg[1], False), traj(seg[2], False)])
    Xte = traj(500, False)
    w = rng.uniform(0.5, 1.5, size=Xtr.shape[0])
    dataset = {"splits": {
        "train": {"X": Xtr, "feature_schema": {"names": ["a", "b", "c", "d"], "n_features": 4},
                  "segments": [{"length": L, "stride": 1} for L in seg]},
        "test": {"X": Xte, "feature_schema": {"names": ["a", "b", "c", "d"], "n_features": 4},
                 "segments": [{"length": 500}]},
    }}
    out = {}
    for tag, kw in (("plain", dict()), ("weighted", dict(frame_weights={"train": w}, min_out_count=2))):
        r = discretize_dataset(dataset, cluster_mode="kmeans", n_microstates=K, lag_time=3, random_state=0, **kw)
        out[f"{tag}_centers"] = r.centers
        out[f"{tag}_lab_train"] = r.assignments["train"]
        out[f"{tag}_lab_test"] = r.assignments["test"]
        out[f"{tag}_counts"] = r.counts
        out[f"{tag}_T"] = r.transition_matrix
        out[f"{tag}_diag_mass"] = r.diag_mass
        out[f"{tag}_counts_before_prune"] = r.counts_before_prune
        out[f"{tag}_state_counts"] = r.state_counts
        out[f"{tag}_pruned"] = (np.asarray([], dtype=np.int32) if r.pruned_state_indices is None
                                 else r.pruned_state_indices)
        out[f"{tag}_counted_pairs"] = r.counted_pairs["train"]
        out[f"{tag}_expected_pairs"] = r.expected_pairs["train"]
        print("discretize", tag, r.counts.shape, out[f"{tag}_pruned"], r.counted_pairs, r.expected_pairs, r.diag_mass)
    np.savez_compressed(OUT / "discretize.npz", Xtr=Xtr, Xte=Xte, seg=np.asarray(seg), w=w, **out)


def _load_reference_ck():
    """ck_runner.py and _ck.py import matplotlib, mdtraj (via _base.py) and deeptime (via _msm_utils.py),
    none of which is installed here.  The two files are loaded straight from the reference tree under a
    throw-away package name with those imports stubbed: matplotlib / mdtraj by inert mocks (plots are not
    part of the fixture), and ``_msm_utils`` by a two-function module -- ``_row_normalize`` restating
    deeptime 0.4.5 ``transition_matrix_non_reversible`` (strictly positive row sums or ValueError) and
    ``pcca_like_macrostates`` returning None (PCCA+ unavailable -> the micro branch, ck_runner.py:201-203).
    Everything else that runs is the reference's own code."""
    import importlib.util
    from unittest import mock

    for name in ("matplotlib", "matplotlib.pyplot", "mdtraj"):
        sys.modules.setdefault(name, mock.MagicMock(name=name))
    pkg = types.ModuleType("_refmsm")
    pkg.__path__ = [str(REF / "src" / "pmarlo" / "markov_state_model")]
    sys.modules["_refmsm"] = pkg
    utils = types.ModuleType("_refmsm._msm_utils")

    def _row_normalize(C):
        arr = np.asarray(C, dtype=float)
        if arr.size == 0:
            return arr.copy()
        rows = 1.0 * np.sum(arr, axis=1)
        if np.min(rows) <= 0:
            raise ValueError("Transition matrix has row sum of " + str(np.min(rows))
                             + ". Must have strictly positive row sums.")
        return np.divide(arr, rows[:, np.newaxis])

    utils._row_normalize = _row_normalize
    utils.pcca_like_macrostates = lambda T, n_macrostates=4, random_state=42: None
    sys.modules["_refmsm._msm_utils"] = utils
    mods = {}
    for name in ("_base", "_ck", "ck_runner"):
        spec = importlib.util.spec_from_file_location(f"_refmsm.{name}", pkg.__path__[0] + f"/{name}.py")
        m = importlib.util.module_from_spec(spec)
        sys.modules[f"_refmsm.{name}"] = m
        spec.loader.exec_module(m)
        mods[name] = m
    return mods


def _markov_dtrajs(rng, K, lengths, stay=0.7, rare=()):
    """Seeded jump process on K states; states in ``rare`` are only entered with small probability."""
    P = rng.random((K, K)) ** 3 + 1e-3
    for r in rare:
        P[:, r] *= 0.01
    P = P / P.sum(axis=1, keepdims=True)
    P = stay * np.eye(K) + (1 - stay) * P
    cum = np.cumsum(P, axis=1)
    out = []
    for n in lengths:
        s = np.empty(n, dtype=np.int64)
        s[0] = rng.integers(0, K)
        u = rng.random(n)
        for t in range(1, n):
            s[t] = min(int(np.searchsorted(cum[s[t - 1]], u[t])), K - 1)
        out.append(s)
    return out


def make_ck():
    """ck_runner.run_ck (ck_runner.py:293-332) and CKMixin.compute_ck_test_micro / select_lag_time_ck
    (_ck.py:61-110,159-175) on seeded label trajectories."""
    import tempfile

    mods = _load_reference_ck()
    run_ck = mods["ck_runner"].run_ck
    CKMixin = mods["_ck"].CKMixin

    class Host(CKMixin):
        def __init__(self, dtrajs, n_states, lag, out):
            self.dtrajs, self.n_states, self.lag_time = dtrajs, n_states, lag
            self.transition_matrix, self.output_dir = None, out

    rng = np.random.default_rng(20260519)
    cases = {
        # name: (dtrajs, kwargs for run_ck)
        "dense12": (_markov_dtrajs(rng, 12, [4000, 2500, 3500]), dict(lag_time=2, min_trans=20, top_n_micro=50)),
        "top8of20": (_markov_dtrajs(rng, 20, [6000, 5000], rare=(3, 11, 17)),
                     dict(lag_time=3, min_trans=10, top_n_micro=8, factors=(2, 3, 5))),
        "gaps": (None, dict(lag_time=1, min_trans=5, top_n_micro=6)),
        "scarce": (_markov_dtrajs(rng, 15, [300, 200], rare=(1, 2)), dict(lag_time=4, min_trans=40, top_n_micro=10)),
        "cycle": ([np.array([0, 1, 2] * 1000, dtype=int)], dict(lag_time=1, macro_k=3, min_trans=5, top_n_micro=3)),
        "tiny": ([np.array([0, 1, 0, 1], dtype=int)], dict(lag_time=1, macro_k=2, min_trans=50, top_n_micro=2)),
        # the longest lag multiples run out of pairs: some factors are reported, others flagged
        "partial": (_markov_dtrajs(rng, 4, [60] * 40, stay=0.5), dict(lag_time=10, min_trans=30, top_n_micro=4)),
        # test_ck_tau_selection.py: tau=2 replaces tau=1 on an MSE tie
        "pattern": ([np.array([0, 0, 1, 1] * 500, dtype=int)], dict(lag_time=1, min_trans=5, top_n_micro=2)),
    }
    # "gaps": labels with unused ids (never visited) and negative (unassigned) frames
    g = _markov_dtrajs(rng, 9, [3000, 3000])
    remap = np.array([0, 2, 3, 5, 8, 9, 12, 13, 14])
    g = [remap[t] for t in g]
    g[0][100:130] = -1
    cases["gaps"] = (g, cases["gaps"][1])

    out = {"case_names": np.array(sorted(cases))}
    with tempfile.TemporaryDirectory() as tmp:
        for name, (dtrajs, kw) in cases.items():
            r = run_ck(dtrajs, output_dir=tmp, **kw)
            ks = sorted(r.mse)
            out[f"{name}_lens"] = np.array([len(t) for t in dtrajs])
            out[f"{name}_labels"] = np.concatenate(dtrajs).astype(np.int32)
            out[f"{name}_kw"] = np.array(repr(kw))
            out[f"{name}_ks"] = np.array(ks, dtype=np.int64)
            out[f"{name}_mse"] = np.array([r.mse[k] for k in ks], dtype=np.float64)
            out[f"{name}_insufficient"] = np.array(r.insufficient_k, dtype=np.int64)
            out[f"{name}_mode"] = np.array(r.mode)
            print("ck", name, r.mode, {k: float(f"{v:.3e}") for k, v in r.mse.items()}, r.insufficient_k)
            # the mixin's micro test and lag selection on the same labels
            K = int(max(int(t.max()) for t in dtrajs)) + 1
            h = Host([np.asarray(t) for t in dtrajs], K, int(kw["lag_time"]), tmp)
            m = h.compute_ck_test_micro(factors=[2, 3, 4], max_states=int(kw["top_n_micro"]), min_transitions=5)
            mk = sorted(m.mse)
            out[f"{name}_mixin_ks"] = np.array(mk, dtype=np.int64)
            out[f"{name}_mixin_mse"] = np.array([m.mse[k] for k in mk], dtype=np.float64)
            out[f"{name}_mixin_insufficient"] = np.array(bool(m.insufficient_data))
            if name in ("dense12", "top8of20", "gaps", "pattern"):
                import contextlib, io
                with contextlib.redirect_stdout(io.StringIO()):
                    cand = [1, 2, 3] if name == "pattern" else [1, 2, 3, 5, 8]
                    taus, mses, its = h._evaluate_candidates(tau_candidates=cand, factor=2)
                    sel = h.select_lag_time_ck(cand, factor=2)
                out[f"{name}_sel_taus"] = np.array(cand)
                out[f"{name}_sel"] = np.array(sel)
                out[f"{name}_sel_mses"] = np.array(mses)
                out[f"{name}_sel_its"] = np.array(its)
                print("   select", sel, np.round(mses, 6))
    np.savez_compressed(OUT / "ck.npz", **out)


def _load_reference_selector():
    """ck_its_selector.py imports deeptime's TransitionCountEstimator / MaximumLikelihoodMSM and three helpers
    of _msm_utils (deeptime again).  It is loaded from the reference tree with
    * ``MaximumLikelihoodMSM`` replaced by a minimal estimator built on the ORACLE's restatement of
      deeptime (sliding counts, largest connected set, reversible MLE, timescales) -- so this fixture pins the
      reference's control flow, thresholds, CK errors, coverage / median-count / diagonal-mass logic and the
      selection rule, NOT deeptime's MLE arithmetic (that part stays "parity unpinned");
    * ``_row_normalize`` as in _load_reference_ck, ``_stationary_from_T`` by the Perron vector (numpy),
      ``pcca_like_macrostates`` by a switchable stand-in (None, or contiguous blocks of states)."""
    import importlib.util
    from unittest import mock

    sys.path.insert(0, str(OUT.parent.parent))
    import oracle.ck as ock
    import oracle.msm as omsm

    mods = _load_reference_ck()          # installs the _refmsm package and its _msm_utils stub
    utils = sys.modules["_refmsm._msm_utils"]
    utils._stationary_from_T = lambda T: omsm.stationary_distribution(np.asarray(T, dtype=float))
    state = {"blocks": False}

    def pcca(T, n_macrostates=4, random_state=42):
        if not state["blocks"]:
            return None
        n = T.shape[0]
        return (np.arange(n) * int(n_macrostates)) // n
    utils.pcca_like_macrostates = pcca

    class _Model:
        def __init__(self, ts, T):
            self._ts, self.transition_matrix = ts, T

        def timescales(self):
            return self._ts

    class MaximumLikelihoodMSM:
        def __init__(self, lagtime, reversible=True):
            self.lag = int(lagtime)

        def fit(self, dtrajs):
            self._m = _Model(*ock.rev_msm_summary(dtrajs, self.lag, None))
            return self

        def fetch_model(self):
            return self._m

    dt = types.ModuleType("deeptime"); dtm = types.ModuleType("deeptime.markov"); dtmm = types.ModuleType("deeptime.markov.msm")
    dtm.TransitionCountEstimator = mock.MagicMock()
    dtmm.MaximumLikelihoodMSM = MaximumLikelihoodMSM
    sys.modules.update({"deeptime": dt, "deeptime.markov": dtm, "deeptime.markov.msm": dtmm})
    spec = importlib.util.spec_from_file_location("_refmsm.ck_its_selector",
                                                  str(REF / "src/pmarlo/markov_state_model/ck_its_selector.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules["_refmsm.ck_its_selector"] = m
    spec.loader.exec_module(m)
    return m, state


def make_ck_selector():
    """ck_its_selector.select_optimal_lag_ck_its (ck_its_selector.py:462-599) with _evaluate_single_lag :279-459."""
    import logging
    logging.disable(logging.CRITICAL)
    sel_mod, state = _load_reference_selector()
    rng = np.random.default_rng(20260520)
    base = _markov_dtrajs(rng, 12, [6000, 5000, 30], stay=0.9)
    split = [np.concatenate([t % 6, 6 + (t[::-1] % 6)]) for t in _markov_dtrajs(rng, 12, [4000], stay=0.8)]
    split = [split[0][:4000], split[0][4000:]]            # two disconnected halves of the state space
    sparse = _markov_dtrajs(rng, 40, [900, 700], stay=0.5)
    tail = [np.concatenate([t, [12]]) for t in _markov_dtrajs(rng, 12, [3000], stay=0.8)]   # state 12: empty row
    meta = _markov_dtrajs(rng, 12, [8000, 8000], stay=0.95)
    cases = {
        "base": (base, dict(tau_candidates=[1, 2, 5, 10, 20000], min_median_count=50), False),
        "strict": (base, dict(tau_candidates=[2, 5, 10], min_median_count=50, ck_threshold=0.01), False),
        "split": (split, dict(tau_candidates=[1, 3], min_median_count=10), False),
        "sparse": (sparse, dict(tau_candidates=[1, 2, 4], horizons=[1, 2]), False),
        "tail": (tail, dict(tau_candidates=[1, 2], min_median_count=10, coverage_threshold=0.9), False),
        "lowdiag": (base, dict(tau_candidates=[10, 20], min_median_count=50, diag_mass_threshold=0.9), False),
        "macro": (meta, dict(tau_candidates=[2, 4, 8], min_median_count=50, horizons=[1, 2, 3]), True),
    }
    out = {"case_names": np.array(sorted(cases))}
    for name, (dtrajs, kw, blocks) in cases.items():
        state["blocks"] = blocks
        sel, evs = sel_mod.select_optimal_lag_ck_its([np.asarray(t) for t in dtrajs], **kw)
        out[f"{name}_lens"] = np.array([len(t) for t in dtrajs])
        out[f"{name}_labels"] = np.concatenate(dtrajs).astype(np.int32)
        out[f"{name}_kw"] = np.array(repr(kw))
        out[f"{name}_blocks"] = np.array(blocks)
        out[f"{name}_selected"] = np.array(sel)
        out[f"{name}_lags"] = np.array([e.lag for e in evs])
        out[f"{name}_ck_error"] = np.array([e.ck_error for e in evs], dtype=float)
        out[f"{name}_coverage"] = np.array([e.coverage_fraction for e in evs], dtype=float)
        out[f"{name}_median"] = np.array([e.median_count for e in evs])
        out[f"{name}_n_macro"] = np.array([e.n_macrostates for e in evs])
        out[f"{name}_passed"] = np.array([e.passed_sanity for e in evs])
        out[f"{name}_reason"] = np.array(["" if e.failure_reason is None else e.failure_reason for e in evs])
        out[f"{name}_diag_mass"] = np.array([np.nan if e.diag_mass is None else e.diag_mass for e in evs], dtype=float)
        out[f"{name}_gap"] = np.array([np.nan if e.eigenvalue_gap is None else e.eigenvalue_gap for e in evs], dtype=float)
        out[f"{name}_ts3"] = np.array([[np.nan] * 3 if e.timescales is None else list(e.timescales[:3]) for e in evs], dtype=float)
        print("selector", name, "->", sel, [(e.lag, float(f"{e.ck_error:.4g}"), e.passed_sanity, e.n_macrostates, e.failure_reason) for e in evs])
    np.savez_compressed(OUT / "ck_selector.npz", **out)


def make_macro():
    """_msm_utils.compute_macro_populations / lump_micro_to_macro_T / compute_macro_mfpt (_msm_utils.py:103-162):
    pure numpy functions of a module that imports deeptime at the top, loaded with deeptime mocked."""
    import importlib.util
    from unittest import mock

    for name in ("deeptime", "deeptime.markov", "deeptime.markov.msm", "deeptime.markov.tools",
                 "deeptime.markov.tools.analysis", "deeptime.markov.tools.estimation",
                 "deeptime.markov.tools.estimation.dense", "deeptime.markov.tools.estimation.dense.transition_matrix"):
        sys.modules[name] = mock.MagicMock(name=name)
    spec = importlib.util.spec_from_file_location("_ref_msm_utils_real",
                                                  str(REF / "src/pmarlo/markov_state_model/_msm_utils.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    rng = np.random.default_rng(77)
    K, n_macro = 60, 5
    T = rng.random((K, K)) ** 4 + 1e-4
    T /= T.sum(axis=1, keepdims=True)
    w, v = np.linalg.eig(T.T)
    pi = np.abs(np.real(v[:, np.argmax(np.real(w))])); pi /= pi.sum()
    lab = rng.integers(0, n_macro, size=K)
    lab[:n_macro] = np.arange(n_macro)
    pops = m.compute_macro_populations(pi, lab)
    Tm = m.lump_micro_to_macro_T(T, pi, lab)
    mf = m.compute_macro_mfpt(Tm)
    np.savez_compressed(OUT / "macro.npz", T=T, pi=pi, lab=lab, pops=pops, Tm=Tm, mfpt=mf)
    print("macro:", pops.round(3), Tm.shape, float(mf.max()))


def make_ladders():
    """utils.msm_utils.candidate_lag_ladder (utils/msm_utils.py:21-105), the genuine function, and
    ITSMixin._deterministic_its_from_counts (_its.py:742-801) run with deeptime's two analysis functions
    stood in for by their documented numpy definitions (eigenvalues(T, k, reversible=True, mu): eigvalsh of
    sqrt(mu) T / sqrt(mu), sorted by magnitude; timescales: -tau / ln|ev|, inf where |ev| = 1): everything
    around those two calls -- symmetrisation, the weights passed as mu, sorting, clipping, padding -- is the
    reference's own arithmetic."""
    import importlib.util
    from unittest import mock

    import scipy.linalg

    def dt_eigenvalues(T, k=None, reversible=False, mu=None, **kw):
        smu = np.sqrt(mu)
        S = smu[:, None] * np.asarray(T) / smu
        ev = scipy.linalg.eigvalsh(S)
        ev = ev[np.argsort(np.abs(ev))[::-1]]
        return ev if k is None else ev[:k]

    def dt_timescales(T, tau=1, k=None, reversible=False, mu=None, **kw):
        ev = dt_eigenvalues(T, k=k, reversible=reversible, mu=mu)
        ts = np.zeros(len(ev))
        one = np.isclose(np.abs(ev), 1.0, rtol=0.0, atol=1e-14)
        ts[one] = np.inf
        ts[~one] = -1.0 * tau / np.log(np.abs(ev[~one]))
        return ts

    for name in ("deeptime", "deeptime.markov", "deeptime.markov.msm", "deeptime.markov.tools"):
        sys.modules[name] = mock.MagicMock(name=name)
    ana = mock.MagicMock(name="deeptime.markov.tools.analysis")
    ana.eigenvalues, ana.timescales = dt_eigenvalues, dt_timescales
    sys.modules["deeptime.markov.tools.analysis"] = ana
    spec = importlib.util.spec_from_file_location("_ref_utils_msm", str(REF / "src/pmarlo/utils/msm_utils.py"))
    mu_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mu_mod)
    out = {}
    args = [(1, 200, None), (1, 100, None), (1, 2000, 7), (5, 500, 4), (1, 2000, 1), (1, 2000, 2), (10, 10, None),
            (3, 1280, 12), (1, 50, 30)]
    out["ladder_args"] = np.array([[a, b, -1 if c is None else c] for a, b, c in args], dtype=np.int64)
    for i, (a, b, c) in enumerate(args):
        out[f"ladder_{i}"] = np.asarray(mu_mod.candidate_lag_ladder(a, b, c), dtype=np.int64)
    spec = importlib.util.spec_from_file_location("pmarlo.markov_state_model._its_real",
                                                  str(REF / "src/pmarlo/markov_state_model/_its.py"))
    its = importlib.util.module_from_spec(spec)
    its.__package__ = "pmarlo.markov_state_model"
    spec.loader.exec_module(its)
    rng = np.random.default_rng(5)
    cases = []
    for K, n_ts in ((6, 3), (40, 5), (3, 5), (25, 0), (12, 4)):
        C = rng.poisson(3.0, size=(K, K)).astype(float) * (rng.random((K, K)) < 0.5)
        C += np.diag(rng.poisson(30.0, size=K).astype(float))
        if K == 12:
            C[4, :] = 0.0
            C[:, 4] = 0.0              # an empty state: zero row of T
        ev, ts, rt = its.ITSMixin._deterministic_its_from_counts(None, 7, C, n_ts)
        i = len(cases)
        out[f"det_C_{i}"], out[f"det_ev_{i}"], out[f"det_ts_{i}"], out[f"det_rate_{i}"] = C, ev, ts, rt
        cases.append((K, n_ts))
    out["det_cases"] = np.asarray(cases, dtype=np.int64)
    np.savez_compressed(OUT / "ladders.npz", **out)
    print("ladders:", [out[f"ladder_{i}"].tolist() for i in (0, 2, 3)], out["det_ts_0"])


def make_its_stats():
    """ITSMixin._summarize_its_stats (_its.py:543-668) run from the reference file on seeded reversible
    transition-matrix samples; deeptime's ``eigenvalues(T, k)`` / ``stationary_distribution`` are stood in for
    by their numpy definitions (eigvals sorted by magnitude; the left Perron vector)."""
    import importlib.util
    from unittest import mock

    sys.path.insert(0, str(OUT.parents[1]))
    import oracle

    def dt_eigenvalues(T, k=None, **kw):
        ev = np.linalg.eigvals(np.asarray(T))
        ev = ev[np.argsort(np.abs(ev))[::-1]]
        return ev if k is None else ev[:k]

    def dt_stationary(T, check_inputs=True):
        w, v = np.linalg.eig(np.asarray(T).T)
        p = np.abs(np.real(v[:, np.argmax(np.real(w))]))
        return p / p.sum()

    for name in ("deeptime", "deeptime.markov", "deeptime.markov.msm", "deeptime.markov.tools"):
        sys.modules[name] = mock.MagicMock(name=name)
    ana = mock.MagicMock(name="deeptime.markov.tools.analysis")
    ana.eigenvalues, ana.stationary_distribution = dt_eigenvalues, dt_stationary
    sys.modules["deeptime.markov.tools.analysis"] = ana
    spec = importlib.util.spec_from_file_location("pmarlo.markov_state_model._its_real2",
                                                  str(REF / "src/pmarlo/markov_state_model/_its.py"))
    its = importlib.util.module_from_spec(spec)
    its.__package__ = "pmarlo.markov_state_model"
    spec.loader.exec_module(its)
    rng = np.random.default_rng(11)
    out, cases = {}, []
    for K, n_ts, lag in ((5, 3, 4), (4, 6, 10), (8, 5, 1)):
        C = rng.poisson(4.0, size=(K, K)).astype(float) + np.diag(rng.poisson(60.0, size=K).astype(float))
        T0, pi0, _ = oracle.msm.mle_rev(C)
        Ts, pis = oracle.bayes.sample_reversible(C, T0, pi0, 40, seed=K)
        st = its.ITSMixin._summarize_its_stats(None, lag, Ts, n_ts, 2.5, 97.5)
        i = len(cases)
        out[f"T_{i}"], out[f"pi_{i}"] = Ts, pis
        for j, v in enumerate(st):
            out[f"stat_{i}_{j}"] = np.asarray(v, dtype=float)
        cases.append((K, n_ts, lag))
    out["cases"] = np.asarray(cases, dtype=np.int64)
    np.savez_compressed(OUT / "its_stats.npz", **out)
    print("its_stats:", out["stat_0_3"], out["stat_1_3"])


if __name__ == "__main__":
    make_its_stats()
    make_ladders()
    make_macro()
    make_ck_selector()
    make_ck()
    make_discretize()
    make_topologies()
    make_counts()
    make_timescales()
    make_preprocess()
    make_tica_xcheck()
    make_assign()
