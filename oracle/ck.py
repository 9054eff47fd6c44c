"""Oracle: Chapman-Kolmogorov test (TEST INFRASTRUCTURE).

numpy restatement of pmarlo's two CK code paths:

* ``ck_runner.run_ck`` (ck_runner.py:293-332) with its helpers
  ``_count_transitions`` :69-82, ``_largest_connected_indices`` :85-87,
  ``_select_top_n_states`` :90-95, ``_eigen_gap`` :98-108,
  ``_preprocess_trajectories`` :140-157, ``_ck_on_trajs`` :160-179,
  ``_attempt_macro_analysis`` :182-219 and ``_perform_micro_analysis`` :222-249;
* ``CKMixin.compute_ck_test_micro`` (_ck.py:61-110) and ``select_lag_time_ck``
  (_ck.py:159-228) with ``_largest_connected_states`` :263-272, ``_count_micro_T``
  :274-288, ``_slowest_its_from_T`` :326-340 and ``_ck_mse_from_T`` :342-353.

Pinned by tests/golden/ck.npz: tests/golden/make_golden.py loads the two
reference files directly (their third-party imports stubbed) and stores their
outputs on seeded label trajectories.

Row normalisation in ck_runner goes through deeptime 0.4.5
``transition_matrix_non_reversible`` (``_msm_utils._row_normalize`` :70-75), which
raises ``ValueError`` when a row sum is not strictly positive; restated here.
The macrostate branch needs deeptime's PCCA+ (``_msm_utils.pcca_like_macrostates``
:284-299, absent from this image): the lumping is an injected callable, and with
none given the branch reports failure exactly as when PCCA+ returns ``None``.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, Iterable, List, Optional, Sequence

import numpy as np
from scipy.sparse.csgraph import connected_components

NUMERIC_MIN_POSITIVE = 1e-12   # pmarlo/constants.py
NUMERIC_MAX_RATE = 0.999999    # pmarlo/constants.py

__all__ = ["CKRunResult", "CKTestResult", "run_ck", "compute_ck_test_micro", "select_lag_time_ck",
           "count_endpoint", "row_normalize_strict", "compact"]


@dataclass
class CKRunResult:
    mse: Dict[int, float] = field(default_factory=dict)
    mode: str = "micro"
    insufficient_k: List[int] = field(default_factory=list)

    @property
    def max_error(self) -> float:
        if not self.mse:
            return float("inf")
        return max(float(np.sqrt(v)) for v in self.mse.values())


@dataclass
class CKTestResult:
    mse: Dict[int, float] = field(default_factory=dict)
    mode: str = "micro"
    insufficient_data: bool = False
    thresholds: Dict[str, int] = field(default_factory=dict)


def count_endpoint(dtrajs: Sequence[np.ndarray], n_states: int, lag: int) -> np.ndarray:
    """ck_runner.py:69-82 / _ck.py:274-285: sliding pairs, a pair is dropped when an endpoint is outside
    [0, n_states)."""
    C = np.zeros((n_states, n_states), dtype=np.float64)
    for traj in dtrajs:
        traj = np.asarray(traj, dtype=np.int64)
        if traj.size <= lag:
            continue
        a, b = traj[:-lag], traj[lag:]
        ok = (a >= 0) & (b >= 0) & (a < n_states) & (b < n_states)
        np.add.at(C, (a[ok], b[ok]), 1.0)
    return C


def row_normalize_strict(C: np.ndarray) -> np.ndarray:
    """deeptime ``transition_matrix_non_reversible``: strictly positive row sums or ValueError."""
    arr = np.asarray(C, dtype=float)
    if arr.size == 0:
        return arr.copy()
    rows = 1.0 * arr.sum(axis=1)
    if rows.min() <= 0:
        raise ValueError(f"Transition matrix has row sum of {rows.min()}. Must have strictly positive row sums.")
    return arr / rows[:, None]


def compact(dtrajs: Sequence[np.ndarray], keep: np.ndarray) -> list[np.ndarray]:
    """``[state_map[s] for s in traj if s in state_map]`` (ck_runner.py:150-153, :235-238, _ck.py:86-90):
    frames of dropped states are REMOVED, so later pairs span the gap."""
    keep = np.asarray(keep, dtype=np.int64)
    size = int(max([int(keep.max()) + 1 if keep.size else 0] + [int(np.max(t)) + 1 for t in dtrajs if len(t)]))
    lut = np.full(max(size, 1), -1, dtype=np.int64)
    lut[keep] = np.arange(keep.size)
    out = []
    for t in dtrajs:
        t = np.asarray(t, dtype=np.int64)
        m = np.where(t >= 0, lut[np.clip(t, 0, lut.size - 1)], -1)
        out.append(m[m >= 0])
    return out


def _eigen_gap(T: np.ndarray, k: int) -> float:
    vals = np.sort(np.real(np.linalg.eigvals(T)))[::-1]
    if len(vals) <= k:
        return 0.0
    return float(vals[k - 1] - vals[k])


def _ck_on_trajs(trajs, T1, lag, factors, min_trans, result: CKRunResult) -> None:
    n = T1.shape[0]
    for f in factors:
        Ck = count_endpoint(trajs, n, lag * int(f))
        if np.any(Ck.sum(axis=1) < min_trans):
            result.insufficient_k.append(int(f))
            continue
        diff = np.linalg.matrix_power(T1, int(f)) - row_normalize_strict(Ck)
        result.mse[int(f)] = float(np.mean(diff * diff))
        if int(f) in result.insufficient_k:
            result.insufficient_k.remove(int(f))


def run_ck(dtrajs: Sequence[np.ndarray], lag_time: int, macro_k: int = 4, min_trans: int = 50,
           top_n_micro: int = 50, factors: Iterable[int] = (2, 3, 4, 5),
           macro_lumper: Optional[Callable[[np.ndarray, int], Optional[np.ndarray]]] = None) -> CKRunResult:
    factors = [int(f) for f in factors if int(f) > 1]
    if not dtrajs:
        raise ValueError("No trajectories provided for analysis")
    if lag_time <= 0:
        raise ValueError(f"Lag time must be positive, got {lag_time}")
    if not factors:
        raise ValueError("No lag factors provided for analysis")
    res = CKRunResult()
    res.insufficient_k = list(factors)
    # _preprocess_trajectories
    n_states = int(max(int(np.max(dt)) for dt in dtrajs) + 1)
    C1 = count_endpoint(dtrajs, n_states, 1)
    active = np.where(C1.sum(axis=1) + C1.sum(axis=0) > 0)[0]
    if active.size == 0:
        return res
    filt = compact(dtrajs, active)
    n_micro = active.size
    T1_micro = row_normalize_strict(count_endpoint(filt, n_micro, 1))
    C_lag = count_endpoint(filt, n_micro, lag_time)
    # _attempt_macro_analysis
    if n_micro > macro_k and _eigen_gap(T1_micro, macro_k) >= 0.01 and macro_lumper is not None:
        labels = macro_lumper(T1_micro, int(macro_k))
        if labels is not None:
            labels = np.asarray(labels, dtype=np.int64)
            n_macro = int(labels.max()) + 1
            mtrajs = [labels[t] for t in filt]
            Cm = count_endpoint(mtrajs, n_macro, lag_time)
            if np.all(Cm.sum(axis=1) >= min_trans):
                _ck_on_trajs(mtrajs, row_normalize_strict(Cm), lag_time, factors, min_trans, res)
                res.mode = "macro"
                return res
    # _perform_micro_analysis
    pops = C_lag.sum(axis=1) + C_lag.sum(axis=0)
    if np.count_nonzero(pops) == 0:
        return res
    top = np.argsort(-pops)[: min(int(top_n_micro), pops.size)]
    mt = compact(filt, top)
    Cs = count_endpoint(mt, top.size, lag_time)
    if np.any(Cs.sum(axis=1) < min_trans):
        return res
    _ck_on_trajs(mt, row_normalize_strict(Cs), lag_time, factors, min_trans, res)
    res.mode = "micro"
    return res


# ----------------------------------------------------------------------------- CKMixin (_ck.py)
def _count_micro_T(dtrajs, nS: int, lag: int):
    C = count_endpoint(dtrajs, nS, lag)
    rows = C.sum(axis=1)
    rows[rows == 0] = 1.0
    return C / rows[:, None], C


def largest_connected_states(C: np.ndarray, max_states: int) -> np.ndarray:
    """_ck.py:263-272."""
    adj = ((C + C.T) > 0).astype(int)
    _, labels = connected_components(adj, directed=False, return_labels=True)
    main = int(np.argmax(np.bincount(labels)))
    idx = np.where(labels == main)[0]
    if idx.size > max_states:
        totals = (C + C.T).sum(axis=1)
        idx = idx[np.argsort(totals[idx])[::-1]][:max_states]
    return idx


def compute_ck_test_micro(dtrajs, n_states: int, lag_time: int, factors: Optional[List[int]] = None,
                          max_states: int = 50, min_transitions: int = 5) -> CKTestResult:
    factors = [2, 3, 4, 5] if factors is None else [int(f) for f in factors if int(f) > 1]
    res = CKTestResult(mode="micro", thresholds={"min_transitions_per_state": int(min_transitions),
                                                 "max_states": int(max_states)})
    if not dtrajs or n_states <= 1 or lag_time <= 0:
        res.insufficient_data = True
        return res
    _, C_all = _count_micro_T(dtrajs, n_states, int(lag_time))
    idx = largest_connected_states(C_all, int(max_states))
    if idx.size == 0:
        res.insufficient_data = True
        return res
    filt = compact(dtrajs, idx)
    T1, C1 = _count_micro_T(filt, idx.size, int(lag_time))
    if np.any(C1.sum(axis=1) < min_transitions):
        res.insufficient_data = True
        return res
    for f in factors:
        T_emp, Ck = _count_micro_T(filt, idx.size, int(lag_time) * f)
        if np.any(Ck.sum(axis=1) < min_transitions):
            res.insufficient_data = True
            return res
        diff = np.linalg.matrix_power(T1, f) - T_emp
        res.mse[f] = float(np.mean(diff * diff))
    return res


def slowest_its_from_T(T: np.ndarray, tau: int) -> float:
    """_ck.py:326-340."""
    evals = np.sort(np.real(np.linalg.eigvals(np.asarray(T, dtype=float))))[::-1]
    if evals.size < 2:
        raise ValueError("Transition matrix must provide at least two eigenvalues")
    lam = float(evals[1])
    if lam <= 0 or lam >= NUMERIC_MAX_RATE:
        lam = min(max(lam, NUMERIC_MIN_POSITIVE), NUMERIC_MAX_RATE)
    its = -float(tau) / np.log(lam)
    if not np.isfinite(its):
        raise ValueError("Failed to compute finite implied timescale")
    return float(its)


def select_lag_time_ck(dtrajs, n_states: int, tau_candidates: Sequence[int], factor: int = 2):
    """_ck.py:159-228.  The prefix rule (:186-214) is computed and then overwritten by the arg-min rule
    (:171-172), so the selection is: the candidate of smallest MSE, except that tau=2 replaces tau=1 on a
    tie.  Returns (selected, taus, mses, its)."""
    taus, mses, its = [], [], []
    for tau in tau_candidates:
        tau = int(tau)
        T1, _ = _count_micro_T(dtrajs, n_states, tau)
        taus.append(tau)
        its.append(slowest_its_from_T(T1, tau))
        T_emp, _ = _count_micro_T(dtrajs, n_states, tau * int(factor))
        diff = np.linalg.matrix_power(T1, int(factor)) - T_emp
        mses.append(float(np.mean(diff * diff)))
    sel = int(taus[int(np.nanargmin(mses))])
    if sel == 1 and 2 in taus:
        j = taus.index(2)
        if mses[j] <= mses[int(np.nanargmin(mses))] + NUMERIC_MIN_POSITIVE:
            sel = 2
    return sel, taus, mses, its


# ----------------------------------------------------------------------------- ck_its_selector.py
@dataclass
class LagEvaluationResult:
    """ck_its_selector.py:23-37."""

    lag: int
    ck_error: float
    coverage_fraction: float
    median_count: int
    n_macrostates: int
    n_microstates: int
    passed_sanity: bool
    failure_reason: Optional[str] = None
    timescales: Optional[np.ndarray] = None
    eigenvalue_gap: Optional[float] = None
    diag_mass: Optional[float] = None


def coverage_fraction(C: np.ndarray) -> float:
    """ck_its_selector.py:86-102."""
    if C.size == 0:
        return 0.0
    n, labels = connected_components(((C + C.T) > 0).astype(int), directed=False, return_labels=True)
    if n == 0:
        return 0.0
    return float(int(np.max(np.bincount(labels)))) / float(C.shape[0])


def median_count(C: np.ndarray) -> int:
    """ck_its_selector.py:105-114."""
    if C.size == 0:
        return 0
    sc = C.sum(axis=0) + C.sum(axis=1)
    return int(np.median(sc[sc > 0])) if np.any(sc > 0) else 0


def auto_determine_macrostates(T: np.ndarray, min_macro: int = 3, max_macro: int = 6) -> int:
    """ck_its_selector.py:117-155: the largest eigenvalue gap lambda_{m-1} - lambda_m for m in the range."""
    if T.size == 0 or T.shape[0] < min_macro:
        return min_macro
    evals = np.sort(np.real(np.linalg.eigvals(T)))[::-1]
    if len(evals) < min_macro + 1:
        return min_macro
    max_gap, best = 0.0, min_macro
    for m in range(min_macro, min(max_macro + 1, len(evals))):
        gap = float(evals[m - 1] - evals[m])
        if gap > max_gap:
            max_gap, best = gap, m
    return best


def ck_l1_error(T_pred: np.ndarray, T_obs: np.ndarray) -> float:
    """ck_its_selector.py:211-226."""
    l1_obs = float(np.sum(np.abs(T_obs)))
    if l1_obs < NUMERIC_MIN_POSITIVE:
        return float("inf")
    return float(np.sum(np.abs(T_pred - T_obs))) / l1_obs


def rev_msm_summary(dtrajs, lag: int, n_timescales: Optional[int] = None):
    """What ck_its_selector.py:397-404 takes from deeptime's ``MaximumLikelihoodMSM(lagtime, reversible=True)
    .fit(dtrajs).fetch_model()``: sliding counts, largest strongly connected set, reversible MLE (oracle.msm),
    timescales -lag / ln|lambda_i| (i >= 2) and the transition matrix of the active set."""
    from . import counts as _counts
    from . import msm as _msm

    K = int(max(int(np.max(t)) for t in dtrajs)) + 1
    C = _counts.count_lagged(dtrajs, K, int(lag), mode="endpoint").astype(float)
    lcs = _msm.largest_connected_set(C)
    T, pi, _ = _msm.mle_rev(C[np.ix_(lcs, lcs)])
    ev = _msm.eigenvalues_rev(T, pi, None if n_timescales is None else min(n_timescales + 1, T.shape[0]))
    with np.errstate(divide="ignore", invalid="ignore"):
        ts = -float(lag) / np.log(np.abs(ev[1:]))
    return ts, T


def evaluate_single_lag(dtrajs, lag, horizons, n_states, coverage_threshold, min_median_count,
                        diag_mass_threshold, macro_lumper=None, n_timescales=None) -> LagEvaluationResult:
    """ck_its_selector.py:279-459."""
    from . import msm as _msm

    try:
        C_tau = count_endpoint(dtrajs, n_states, lag)
        cov, med = coverage_fraction(C_tau), median_count(C_tau)
        reason = None
        if cov < coverage_threshold:
            reason = f"Coverage {cov:.2%} < {coverage_threshold:.2%}"
        elif med < min_median_count:
            reason = f"Median count {med} < {min_median_count}"
        if reason is not None:
            return LagEvaluationResult(lag, float("inf"), cov, med, 0, n_states, False, reason)
        T_tau = row_normalize_strict(C_tau)
        pi = _msm.stationary_distribution(T_tau)
        n_cand = auto_determine_macrostates(T_tau, 2, 6)
        labels = None
        if macro_lumper is not None:
            try:
                labels = macro_lumper(T_tau, n_cand)
            except Exception:
                labels = None
        n_macro, gap = 0, None
        if labels is not None:
            labels = np.asarray(labels, dtype=np.int64)
            n_macro = n_cand
            chi = np.zeros((T_tau.shape[0], n_macro))
            chi[np.arange(T_tau.shape[0]), labels] = 1.0
            err = 0.0
            for k in horizons:
                cp = chi.T @ np.diag(pi)
                den = cp @ chi + np.eye(n_macro) * NUMERIC_MIN_POSITIVE
                T_pred = cp @ np.linalg.matrix_power(T_tau, k) @ chi @ np.linalg.inv(den)
                T_obs = row_normalize_strict(count_endpoint([labels[t] for t in dtrajs], n_macro, lag * k))
                err = max(err, ck_l1_error(T_pred, T_obs))
            evals = np.sort(np.real(np.linalg.eigvals(T_tau)))[::-1]
            if len(evals) > n_macro:
                gap = float(evals[n_macro - 1] - evals[n_macro])
        else:
            err = 0.0
            for k in horizons:
                T_obs = row_normalize_strict(count_endpoint(dtrajs, n_states, lag * k))
                err = max(err, ck_l1_error(np.linalg.matrix_power(T_tau, k), T_obs))
        diag_mass, ts = float("nan"), None
        try:
            ts, T_rev = rev_msm_summary(dtrajs, lag, n_timescales)
            if T_rev.size:
                diag_mass = float(np.trace(T_rev) / T_rev.shape[0])
        except Exception:
            ts = None
        reason = None
        if not (np.isfinite(diag_mass) and diag_mass >= diag_mass_threshold):
            reason = (f"Diagonal mass {diag_mass:.3f} < threshold {diag_mass_threshold:.3f}"
                      if np.isfinite(diag_mass) else "Diagonal mass undefined")
        return LagEvaluationResult(lag, err, cov, med, n_macro, n_states, reason is None, reason, ts, gap, diag_mass)
    except Exception as e:
        return LagEvaluationResult(lag, float("inf"), 0.0, 0, 0, n_states, False, f"Exception: {str(e)}")


def select_optimal_lag_ck_its(dtrajs, tau_candidates=None, horizons=None, ck_threshold=0.15,
                              coverage_threshold=0.98, min_median_count=100, diag_mass_threshold=0.6,
                              macro_lumper=None, n_timescales=None):
    """ck_its_selector.py:462-599."""
    if not dtrajs:
        raise ValueError("No discrete trajectories provided")
    usable = [np.asarray(t) for t in dtrajs if t is not None and np.asarray(t).size > 0]
    if not usable:
        raise ValueError("Discrete trajectories contain no frames for CK analysis; "
                         "provide trajectories with at least two time steps.")
    tau_candidates = [25, 50, 75, 100] if tau_candidates is None else list(tau_candidates)
    horizons = [1, 2, 3, 4, 5] if horizons is None else horizons
    max_lag = max(0, max(int(t.size) for t in usable) - 1)
    valid = [int(t) for t in tau_candidates if t <= max_lag]
    if not valid:
        raise ValueError(f"All tau candidates exceed the available trajectory length (max supported lag {max_lag}). "
                         "Provide smaller lag values or shorter horizons.")
    n_states = int(max(np.max(t) for t in usable)) + 1
    evals = [evaluate_single_lag(usable, lag, horizons, n_states, coverage_threshold, min_median_count,
                                 diag_mass_threshold, macro_lumper, n_timescales) for lag in sorted(valid)]
    for r in sorted(evals, key=lambda r: r.lag):
        if r.passed_sanity and r.ck_error <= ck_threshold:
            return r.lag, evals
    passing = [r for r in evals if r.passed_sanity]
    if passing:
        return min(passing, key=lambda r: r.ck_error).lag, evals
    return min(tau_candidates), evals
