"""Chapman-Kolmogorov test on the device kernels.

Mirrors
* ``ck_its_selector.select_optimal_lag_ck_its`` ck_its_selector.py:462-599 (``_evaluate_single_lag`` :279-459,
  ``LagEvaluationResult`` :23-37): per candidate lag the counts at tau and k*tau, coverage / median-count
  guardrails, macro- or microstate CK error, reversible MLE timescales and the diagonal-mass guardrail;
* ``ck_runner.run_ck(dtrajs, lag_time, output_dir, macro_k=4, min_trans=50, top_n_micro=50,
  factors=(2,3,4,5))`` ck_runner.py:293-332 with ``CKRunResult`` :32-48 and ``ck_rms_error`` :51-66;
* ``CKMixin.compute_ck_test_micro`` _ck.py:61-110 and ``CKMixin.select_lag_time_ck`` _ck.py:159-175
  (as functions here, as methods on ``EnhancedMSM``), result container ``CKTestResult`` _base.py:18-32.

Where the work goes: every per-frame Python loop of the reference -- the lagged pair counts at tau and
k*tau (``_count_transitions`` ck_runner.py:69-82, ``_count_micro_T`` _ck.py:274-288) and the rebuilding of
the trajectories after states are dropped or lumped (ck_runner.py:150-153, :204, :235-238, _ck.py:86-90)
-- runs in libpmb200 (``pmb_count_lagged``, ``pmb_relabel_compact``) on a label shard that is copied to
the device once.  T^k and the mean squared difference are K x K fp64 matrix products on the device
(torch.matmul, a plain library GEMM).  What stays on the host is the K-sized bookkeeping whose result
depends on library tie-breaking the reference gets from numpy / scipy and which must therefore be the same
call: ``np.argsort`` of the state populations (ck_runner.py:90-95), ``connected_components`` of the K x K
adjacency (_ck.py:263-272) and ``np.linalg.eigvals`` of a non-symmetric K x K matrix for the spectral-gap
test and the slowest implied timescale (ck_runner.py:98-108, _ck.py:326-340).

The macrostate branch of ``run_ck`` needs PCCA+ memberships, which the reference takes from deeptime
(``_msm_utils.pcca_like_macrostates`` :284-299).  That lumping is the ``macro_lumper`` argument here:
``"pcca"`` selects this package's PCCA+ (``macro.pcca_like_macrostates``), a callable
``macro_lumper(T1_micro, macro_k) -> labels | None`` is used as given (e.g. the reference's own function, see
INTEGRATION.md), and None reports "not feasible" so that the microstate branch runs, exactly as when PCCA+
returns ``None`` (ck_runner.py:201-203).  ``ck.png`` is written only when matplotlib
is importable.
"""

from __future__ import annotations

import csv
import json
import logging
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Callable, Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch
from scipy.sparse.csgraph import connected_components

from . import kernels
from .msm import NUMERIC_MIN_POSITIVE, dtrajs_to_device

logger = logging.getLogger("pmarlo")

NUMERIC_MAX_RATE = 0.999999   # pmarlo/constants.py

__all__ = ["CKRunResult", "CKTestResult", "ck_rms_error", "run_ck", "compute_ck_test_micro",
           "select_lag_time_ck", "LabelShard", "LagEvaluationResult", "select_optimal_lag_ck_its"]


@dataclass
class CKRunResult:
    """ck_runner.py:32-48."""

    mse: Dict[int, float] = field(default_factory=dict)
    mode: str = "micro"
    insufficient_k: List[int] = field(default_factory=list)

    @property
    def max_error(self) -> float:
        if not self.mse:
            return float("inf")
        return max(float(np.sqrt(v)) for v in self.mse.values())

    @property
    def has_valid_tests(self) -> bool:
        return bool(self.mse)


def ck_rms_error(result: CKRunResult) -> float:
    """ck_runner.py:51-66."""
    return result.max_error


@dataclass
class CKTestResult:
    """_base.py:18-32."""

    mse: Dict[int, float] = field(default_factory=dict)
    mode: str = "micro"
    insufficient_data: bool = False
    thresholds: Dict[str, int] = field(default_factory=dict)

    def to_dict(self) -> Dict[str, Any]:
        return {"mse": {int(k): float(v) for k, v in self.mse.items()}, "mode": self.mode,
                "insufficient_data": self.insufficient_data, "thresholds": self.thresholds}


# ----------------------------------------------------------------------------- the label shard on the device
class LabelShard:
    """Back-to-back discrete trajectories in HBM: int32 labels plus int64 shard offsets."""

    def __init__(self, labels: torch.Tensor, offsets: torch.Tensor):
        self.labels, self.offsets = labels, offsets

    @classmethod
    def from_dtrajs(cls, dtrajs: Sequence[np.ndarray]) -> "LabelShard":
        dev = kernels.require_cuda()
        labels, segs = dtrajs_to_device(dtrajs, dev)
        return cls(labels, segs.device(dev))

    @property
    def n_frames(self) -> int:
        return int(self.labels.numel())

    def max_label(self) -> int:
        return int(self.labels.max().item()) if self.n_frames else -1

    def counts(self, n_states: int, lag: int) -> torch.Tensor:
        """(K,K) fp64 counts of (t, t+lag) pairs with both endpoints in [0, K) -- ck_runner.py:69-82."""
        C = kernels.count_lagged(self.labels, self.offsets, int(n_states), int(lag))
        return C.to(torch.float64)

    def relabel(self, lut: np.ndarray) -> "LabelShard":
        """New shard with label s replaced by lut[s]; frames with lut[s] < 0 (or s outside the table) removed."""
        lut_d = torch.from_numpy(np.ascontiguousarray(lut, dtype=np.int32)).to(self.labels.device)
        out, off = kernels.relabel_compact(self.labels, self.offsets, lut_d)
        kept = int(off[-1].item())
        return LabelShard(out[:kept], off)

    def keep_states(self, keep: np.ndarray, n_states: int) -> "LabelShard":
        lut = np.full((max(int(n_states), 1),), -1, dtype=np.int32)
        keep = np.asarray(keep, dtype=np.int64)
        lut[keep] = np.arange(keep.size, dtype=np.int32)
        return self.relabel(lut)


def _as_shard(dtrajs) -> LabelShard:
    return dtrajs if isinstance(dtrajs, LabelShard) else LabelShard.from_dtrajs(dtrajs)


def _row_normalize_strict(C: torch.Tensor) -> torch.Tensor:
    """deeptime ``transition_matrix_non_reversible`` behind ``_msm_utils._row_normalize`` :70-75."""
    if C.numel() == 0:
        return C.clone()
    rows = C.sum(dim=1)
    lo = float(rows.min().item())
    if lo <= 0:
        raise ValueError(f"Transition matrix has row sum of {lo}. Must have strictly positive row sums.")
    return C / rows[:, None]


def _row_normalize_lenient(C: torch.Tensor) -> torch.Tensor:
    """_ck.py:286-288: empty rows stay zero."""
    rows = C.sum(dim=1)
    rows = torch.where(rows == 0, torch.ones_like(rows), rows)
    return C / rows[:, None]


def _mse(T1: torch.Tensor, k: int, T_emp: torch.Tensor) -> float:
    d = torch.linalg.matrix_power(T1, int(k)) - T_emp
    return float((d * d).mean().item())


def _eigvals_desc(T: torch.Tensor | np.ndarray) -> np.ndarray:
    A = T.cpu().numpy() if isinstance(T, torch.Tensor) else np.asarray(T, dtype=float)
    return np.sort(np.real(np.linalg.eigvals(A)))[::-1]


# ----------------------------------------------------------------------------- run_ck
def _validate(n_traj: int, total_frames: int, lag_time: int, factors: Sequence[int]) -> None:
    """ck_runner.py:111-137."""
    if n_traj == 0:
        raise ValueError("No trajectories provided for analysis")
    if lag_time <= 0:
        raise ValueError(f"Lag time must be positive, got {lag_time}")
    if not factors:
        raise ValueError("No lag factors provided for analysis")
    bad = [f for f in factors if f <= 1]
    if bad:
        raise ValueError(f"All lag factors must be > 1, got {bad}")
    if total_frames < 100:
        logger.warning("Very few total frames (%d) may lead to unreliable results", total_frames)
    if lag_time * max(factors) >= total_frames:
        logger.warning("Lag time * max factor (%d) exceeds total frames (%d)", lag_time * max(factors), total_frames)


def _ck_on_shard(shard: LabelShard, T1: torch.Tensor, lag: int, factors, min_trans: int, result: CKRunResult) -> None:
    """ck_runner.py:160-179 (a factor without enough pairs is appended a second time, as there)."""
    n = int(T1.shape[0])
    for f in factors:
        Ck = shard.counts(n, lag * int(f))
        if bool((Ck.sum(dim=1) < min_trans).any().item()):
            result.insufficient_k.append(int(f))
            continue
        result.mse[int(f)] = _mse(T1, int(f), _row_normalize_strict(Ck))
        if int(f) in result.insufficient_k:
            result.insufficient_k.remove(int(f))


def _save_outputs(result: CKRunResult, out: Path) -> None:
    """ck_runner.py:252-290."""
    out.mkdir(parents=True, exist_ok=True)
    with (out / "ck_mse.csv").open("w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["k", "mse"])
        for k, v in sorted(result.mse.items()):
            w.writerow([k, v])
    with (out / "ck_mse.json").open("w", encoding="utf-8") as fh:
        json.dump({"mode": result.mode, "mse": {str(k): v for k, v in result.mse.items()},
                   "insufficient_k": result.insufficient_k}, fh, indent=2)
    try:
        import matplotlib
        matplotlib.use("Agg", force=True)
        import matplotlib.pyplot as plt
    except ImportError:
        return
    plt.figure()
    if result.mse:
        ks = sorted(result.mse)
        plt.plot(ks, [result.mse[k] for k in ks], marker="o", linestyle="-", label="MSE")
        plt.xlabel("k (lag multiple)")
        plt.ylabel("MSE")
        plt.legend()
    if result.insufficient_k:
        msg = "insufficient transitions for CK at k=" + ",".join(str(k) for k in result.insufficient_k)
        plt.text(0.5, 0.5, msg, ha="center", va="center", transform=plt.gca().transAxes)
    plt.tight_layout()
    try:
        plt.savefig(out / "ck.png")
    finally:
        plt.close()


def _resolve_lumper(macro_lumper):
    """``macro_lumper="pcca"`` selects this package's PCCA+ (`macro.pcca_like_macrostates`, what the reference
    calls at ck_runner.py:196-203 / ck_its_selector.py:300-330); a callable is used as given; None disables the
    macrostate branch (the reference's behaviour when PCCA+ returns None).  Note that PCCA+ -- deeptime's as well --
    rejects a transition matrix without detailed balance, which the row-normalised count matrices of this module
    are unless the counts happen to be symmetric: with "pcca" the macrostate branch then reports "not feasible",
    exactly as in the reference."""
    if isinstance(macro_lumper, str):
        if macro_lumper != "pcca":
            raise ValueError(f"unknown macro_lumper {macro_lumper!r}")
        from .macro import pcca_like_macrostates

        return lambda T, k: pcca_like_macrostates(T, n_macrostates=int(k))
    return macro_lumper


def run_ck(dtrajs, lag_time: int, output_dir: str | Path | None = None, macro_k: int = 4, min_trans: int = 50,
           top_n_micro: int = 50, factors: Iterable[int] = (2, 3, 4, 5), *,
           macro_lumper: Optional[Callable[[np.ndarray, int], Optional[np.ndarray]]] = None) -> CKRunResult:
    """``ck_runner.run_ck``.  ``dtrajs``: sequence of integer arrays, or a ``LabelShard`` already in HBM.
    ``output_dir=None`` skips the csv/json/png side outputs."""
    macro_lumper = _resolve_lumper(macro_lumper)
    factors_list = [int(f) for f in factors if int(f) > 1]
    n_traj = int(dtrajs.offsets.numel()) - 1 if isinstance(dtrajs, LabelShard) else len(dtrajs)
    if n_traj == 0:
        raise ValueError("No trajectories provided for analysis")
    shard = _as_shard(dtrajs)
    _validate(n_traj, shard.n_frames, int(lag_time), factors_list)
    result = CKRunResult()
    result.insufficient_k = factors_list.copy()

    def done() -> CKRunResult:
        if output_dir is not None:
            _save_outputs(result, Path(output_dir))
        return result

    # _preprocess_trajectories (ck_runner.py:140-157)
    n_states = shard.max_label() + 1
    if n_states <= 0:
        return done()
    C1 = shard.counts(n_states, 1)
    pops = (C1.sum(dim=1) + C1.sum(dim=0)).cpu().numpy()
    active = np.where(pops > 0)[0]
    if active.size == 0:
        logger.warning("No connected states found in trajectories")
        return done()
    filt = shard.keep_states(active, n_states)
    n_micro = int(active.size)
    T1_micro = _row_normalize_strict(filt.counts(n_micro, 1))
    C_lag = filt.counts(n_micro, int(lag_time))

    # _attempt_macro_analysis (ck_runner.py:182-219)
    if _attempt_macro(filt, T1_micro, int(lag_time), int(macro_k), int(min_trans), factors_list, result, macro_lumper):
        return done()

    # _perform_micro_analysis (ck_runner.py:222-249)
    pops = (C_lag.sum(dim=1) + C_lag.sum(dim=0)).cpu().numpy()
    if np.count_nonzero(pops) == 0:
        logger.warning("No populated states found for micro analysis")
        return done()
    top = np.argsort(-pops)[: min(int(top_n_micro), pops.size)]
    sel = filt.keep_states(top, n_micro)
    Cs = sel.counts(int(top.size), int(lag_time))
    if bool((Cs.sum(dim=1) < min_trans).any().item()):
        logger.info("Insufficient transitions in selected microstates (min_trans=%d)", min_trans)
        return done()
    _ck_on_shard(sel, _row_normalize_strict(Cs), int(lag_time), factors_list, int(min_trans), result)
    result.mode = "micro"
    return done()


def _attempt_macro(filt: LabelShard, T1_micro: torch.Tensor, lag_time: int, macro_k: int, min_trans: int,
                   factors, result: CKRunResult, macro_lumper) -> bool:
    n_micro = int(T1_micro.shape[0])
    if n_micro <= macro_k:
        return False
    vals = _eigvals_desc(T1_micro)
    gap = float(vals[macro_k - 1] - vals[macro_k]) if len(vals) > macro_k else 0.0
    if gap < 0.01 or macro_lumper is None:
        return False
    try:
        labels = macro_lumper(T1_micro.cpu().numpy(), int(macro_k))
    except Exception as exc:                                    # ck_runner.py:196-199
        logger.warning("PCCA+ decomposition failed: %s", exc)
        return False
    if labels is None:
        return False
    labels = np.asarray(labels, dtype=np.int64)
    n_macro = int(labels.max()) + 1
    macro = filt.relabel(labels.astype(np.int32))
    Cm = macro.counts(n_macro, lag_time)
    if not bool((Cm.sum(dim=1) >= min_trans).all().item()):
        return False
    _ck_on_shard(macro, _row_normalize_strict(Cm), lag_time, factors, min_trans, result)
    result.mode = "macro"
    return True


# ----------------------------------------------------------------------------- CKMixin
def _largest_connected_states(C: np.ndarray, max_states: int) -> np.ndarray:
    """_ck.py:263-272."""
    S = C + C.T
    _, labels = connected_components((S > 0).astype(int), directed=False, return_labels=True)
    main = int(np.argmax(np.bincount(labels)))
    idx = np.where(labels == main)[0]
    if idx.size > max_states:
        totals = S.sum(axis=1)
        idx = idx[np.argsort(totals[idx])[::-1]][:max_states]
    return idx


def compute_ck_test_micro(dtrajs, n_states: int, lag_time: int, factors: Optional[List[int]] = None,
                          max_states: int = 50, min_transitions: int = 5) -> CKTestResult:
    """``CKMixin.compute_ck_test_micro`` (_ck.py:61-110)."""
    factors = [2, 3, 4, 5] if factors is None else [int(f) for f in factors if int(f) > 1]
    res = CKTestResult(mode="micro", thresholds={"min_transitions_per_state": int(min_transitions),
                                                 "max_states": int(max_states)})
    empty = (dtrajs.n_frames == 0) if isinstance(dtrajs, LabelShard) else (not dtrajs)
    if empty or int(n_states) <= 1 or int(lag_time) <= 0:
        res.insufficient_data = True
        return res
    shard = _as_shard(dtrajs)
    idx = _largest_connected_states(shard.counts(int(n_states), int(lag_time)).cpu().numpy(), int(max_states))
    if idx.size == 0:
        res.insufficient_data = True
        return res
    filt = shard.keep_states(idx, int(n_states))
    n_sel = int(idx.size)
    C1 = filt.counts(n_sel, int(lag_time))
    if bool((C1.sum(dim=1) < min_transitions).any().item()):
        res.insufficient_data = True
        return res
    T1 = _row_normalize_lenient(C1)
    for f in factors:
        Ck = filt.counts(n_sel, int(lag_time) * f)
        if bool((Ck.sum(dim=1) < min_transitions).any().item()):
            res.insufficient_data = True
            return res
        res.mse[int(f)] = _mse(T1, f, _row_normalize_lenient(Ck))
    return res


def _slowest_its(T: torch.Tensor, tau: int) -> float:
    """_ck.py:326-340."""
    evals = _eigvals_desc(T)
    if evals.size < 2:
        raise ValueError("Transition matrix must provide at least two eigenvalues")
    lam = float(evals[1])
    if lam <= 0 or lam >= NUMERIC_MAX_RATE:
        lam = min(max(lam, NUMERIC_MIN_POSITIVE), NUMERIC_MAX_RATE)
    its = -float(tau) / np.log(lam)
    if not np.isfinite(its):
        raise ValueError("Failed to compute finite implied timescale")
    return float(its)


def select_lag_time_ck(dtrajs, n_states: int, tau_candidates: Sequence[int], factor: int = 2,
                       mse_epsilon: float = 0.05, output_dir: str | Path | None = None):
    """``CKMixin.select_lag_time_ck`` (_ck.py:159-175).  The prefix rule of :186-214 is evaluated by the
    reference and then overwritten by the smallest-MSE rule (:171), so ``mse_epsilon`` has no effect on
    the selection; tau = 2 replaces tau = 1 on a tie (:220-228).  Returns (selected, taus, mses, its)."""
    del mse_epsilon
    shard = _as_shard(dtrajs)
    taus, mses, its = [], [], []
    for tau in tau_candidates:
        tau = int(tau)
        T1 = _row_normalize_lenient(shard.counts(int(n_states), tau))
        taus.append(tau)
        its.append(_slowest_its(T1, tau))
        mses.append(_mse(T1, int(factor), _row_normalize_lenient(shard.counts(int(n_states), tau * int(factor)))))
    selected = int(taus[int(np.nanargmin(mses))])
    if selected == 1 and 2 in taus:
        j = taus.index(2)
        if mses[j] <= mses[int(np.nanargmin(mses))] + NUMERIC_MIN_POSITIVE:
            selected = 2
    if output_dir is not None:
        out = Path(output_dir)
        out.mkdir(parents=True, exist_ok=True)
        with (out / "ck_mse.csv").open("w", newline="") as fh:     # _ck.py:230-243
            w = csv.writer(fh)
            w.writerow(["tau", "mse"])
            for t, m in zip(taus, mses):
                w.writerow([int(t), float(m)])
    return int(selected), taus, mses, its


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


def _coverage_and_median(C: np.ndarray) -> tuple[float, int]:
    """ck_its_selector.py:86-114 on the K x K count matrix (host: graph bookkeeping, a K-sized median)."""
    if C.size == 0:
        return 0.0, 0
    n, labels = connected_components(((C + C.T) > 0).astype(int), directed=False, return_labels=True)
    cov = 0.0 if n == 0 else float(int(np.max(np.bincount(labels)))) / float(C.shape[0])
    sc = C.sum(axis=0) + C.sum(axis=1)
    med = int(np.median(sc[sc > 0])) if np.any(sc > 0) else 0
    return cov, med


def _auto_determine_macrostates(evals_desc: np.ndarray, n: int, min_macro: int, max_macro: int) -> int:
    """ck_its_selector.py:117-155 given the sorted real parts of the spectrum of T."""
    if n < min_macro or len(evals_desc) < min_macro + 1:
        return min_macro
    max_gap, best = 0.0, min_macro
    for m in range(min_macro, min(max_macro + 1, len(evals_desc))):
        gap = float(evals_desc[m - 1] - evals_desc[m])
        if gap > max_gap:
            max_gap, best = gap, m
    return best


def _l1_error(T_pred: torch.Tensor, T_obs: torch.Tensor) -> float:
    """ck_its_selector.py:211-226."""
    l1_obs = float(T_obs.abs().sum().item())
    if l1_obs < NUMERIC_MIN_POSITIVE:
        return float("inf")
    return float((T_pred - T_obs).abs().sum().item()) / l1_obs


def _stationary_device(T: torch.Tensor) -> torch.Tensor:
    """Left Perron vector of a row-stochastic matrix (``_stationary_from_T``, _msm_utils.py:78-88): one
    K x K linear solve on the device, the redundant balance equation replaced by sum(pi) = 1."""
    n = int(T.shape[0])
    A = T.t() - torch.eye(n, dtype=T.dtype, device=T.device)
    A[n - 1, :] = 1.0
    b = torch.zeros((n,), dtype=T.dtype, device=T.device)
    b[n - 1] = 1.0
    pi = torch.linalg.solve(A, b).abs()
    return pi / pi.sum()


def _reversible_summary(shard: LabelShard, n_states: int, lag: int, n_timescales: int):
    """What ck_its_selector.py:397-404 takes from ``MaximumLikelihoodMSM(lagtime, reversible=True).fit(dtrajs)``:
    sliding counts (K7) -> largest strongly connected set (host graph bookkeeping) -> reversible MLE (K8) ->
    leading timescales (K9) and trace(T) / n over the active set."""
    from .msm import largest_connected_set

    C = kernels.count_lagged(shard.labels, shard.offsets, int(n_states), int(lag))
    lcs = largest_connected_set(C.cpu().numpy())
    act = torch.zeros((int(n_states),), dtype=torch.uint8, device=C.device)
    act[torch.from_numpy(lcs).to(C.device)] = 1
    T, pi, info = kernels.mle_rev(C.to(torch.float64), act, alpha=0.0)
    if int(info[1].item()) < 0:
        raise ValueError("reversible MLE failed: a state of the active set has no outgoing counts")
    idx = torch.from_numpy(lcs).to(C.device)
    diag_mass = float(T.diagonal()[idx].sum().item()) / float(lcs.size) if lcs.size else float("nan")
    k = int(min(int(n_timescales) + 1, lcs.size))
    ev, _ = kernels.eig_rev_topk(T, pi, k)
    ev = ev.cpu().numpy()
    with np.errstate(divide="ignore", invalid="ignore"):
        ts = -float(lag) / np.log(np.abs(ev[1:]))
    return ts, diag_mass


def _evaluate_single_lag(shard: LabelShard, lag: int, horizons: Sequence[int], n_states: int,
                         coverage_threshold: float, min_median_count: int, diag_mass_threshold: float,
                         macro_lumper, n_timescales: int) -> LagEvaluationResult:
    """ck_its_selector.py:279-459."""
    logger.info("[CK-ITS] Evaluating lag=%d", lag)
    try:
        C_tau = shard.counts(n_states, lag)
        cov, med = _coverage_and_median(C_tau.cpu().numpy())
        reason = None
        if cov < coverage_threshold:
            reason = f"Coverage {cov:.2%} < {coverage_threshold:.2%}"
        elif med < min_median_count:
            reason = f"Median count {med} < {min_median_count}"
        if reason is not None:
            logger.warning("[CK-ITS] Lag %d failed sanity: %s", lag, reason)
            return LagEvaluationResult(lag, float("inf"), cov, med, 0, n_states, False, reason)
        T_tau = _row_normalize_strict(C_tau)
        evals = _eigvals_desc(T_tau)
        n_cand = _auto_determine_macrostates(evals, int(T_tau.shape[0]), 2, 6)
        labels = None
        if macro_lumper is not None:
            try:
                labels = macro_lumper(T_tau.cpu().numpy(), n_cand)
            except Exception as exc:                                      # ck_its_selector.py:335-338
                logger.warning("[CK-ITS] PCCA+ exception for lag %d: %s", lag, exc)
        n_macro, gap, err = 0, None, 0.0
        if labels is not None:
            labels = np.asarray(labels, dtype=np.int64)
            n_macro = n_cand
            pi = _stationary_device(T_tau)
            chi = torch.zeros((int(T_tau.shape[0]), n_macro), dtype=torch.float64, device=T_tau.device)
            chi[torch.arange(int(T_tau.shape[0]), device=T_tau.device), torch.from_numpy(labels).to(T_tau.device)] = 1.0
            cp = chi.t() * pi[None, :]                                    # chi^T diag(pi)
            den = cp @ chi + torch.eye(n_macro, dtype=torch.float64, device=T_tau.device) * NUMERIC_MIN_POSITIVE
            macro = shard.relabel(labels.astype(np.int32))
            for k in horizons:
                T_pred = cp @ torch.linalg.matrix_power(T_tau, int(k)) @ chi @ torch.linalg.inv(den)
                T_obs = _row_normalize_strict(macro.counts(n_macro, lag * int(k)))
                err = max(err, _l1_error(T_pred, T_obs))
            if len(evals) > n_macro:
                gap = float(evals[n_macro - 1] - evals[n_macro])
        else:
            if macro_lumper is not None:
                logger.warning("[CK-ITS] PCCA+ failed for lag %d, using microstate CK test fallback", lag)
            for k in horizons:
                T_obs = _row_normalize_strict(shard.counts(n_states, lag * int(k)))
                err = max(err, _l1_error(torch.linalg.matrix_power(T_tau, int(k)), T_obs))
        diag_mass, ts = float("nan"), None
        try:
            ts, diag_mass = _reversible_summary(shard, n_states, lag, n_timescales)
        except Exception as exc:
            logger.warning("[CK-ITS] Failed to compute timescales for lag %d: %s", lag, exc)
            ts = None
        reason = None
        if not (np.isfinite(diag_mass) and diag_mass >= diag_mass_threshold):
            reason = (f"Diagonal mass {diag_mass:.3f} < threshold {diag_mass_threshold:.3f}"
                      if np.isfinite(diag_mass) else "Diagonal mass undefined")
            logger.warning("[CK-ITS] Lag %d failed diagonal-mass guardrail: %s", lag, reason)
        return LagEvaluationResult(lag, err, cov, med, n_macro, n_states, reason is None, reason, ts, gap, diag_mass)
    except Exception as e:
        logger.error("[CK-ITS] Failed to evaluate lag %d: %s", lag, e)
        return LagEvaluationResult(lag, float("inf"), 0.0, 0, 0, n_states, False, f"Exception: {str(e)}")


def select_optimal_lag_ck_its(dtrajs: Sequence[np.ndarray], tau_candidates: Optional[List[int]] = None,
                              horizons: Optional[List[int]] = None, ck_threshold: float = 0.15,
                              coverage_threshold: float = 0.98, min_median_count: int = 100,
                              diag_mass_threshold: float = 0.6, *, macro_lumper=None,
                              n_timescales: int = 10):
    """``ck_its_selector.select_optimal_lag_ck_its`` (ck_its_selector.py:462-599): the smallest lag whose CK
    error is below the threshold among those that pass the coverage, median-count and diagonal-mass
    guardrails; fallbacks as in the reference.  Returns (selected_lag, [LagEvaluationResult]).

    Differences, both documented in INTEGRATION.md: PCCA+ is the injected ``macro_lumper(T, n_macro)`` (None:
    microstate CK test, the reference's fallback); ``LagEvaluationResult.timescales`` holds the leading
    ``n_timescales`` values (the reference stores all n - 1; they do not enter the selection)."""
    macro_lumper = _resolve_lumper(macro_lumper)
    if dtrajs is None or len(dtrajs) == 0:
        raise ValueError("No discrete trajectories provided")
    usable = [np.asarray(t) for t in dtrajs if t is not None and np.asarray(t).size > 0]
    if not usable:
        raise ValueError("Discrete trajectories contain no frames for CK analysis; "
                         "provide trajectories with at least two time steps.")
    tau_candidates = [25, 50, 75, 100] if tau_candidates is None else list(tau_candidates)
    horizons = [1, 2, 3, 4, 5] if horizons is None else list(horizons)
    max_lag = max(0, max(int(t.size) for t in usable) - 1)                # ck_its_selector.py:40-67
    valid = [int(t) for t in tau_candidates if t <= max_lag]
    ignored = [int(t) for t in tau_candidates if t > max_lag]
    if ignored:
        logger.warning("[CK-ITS] Ignoring %d tau candidates that exceed available length (max supported lag=%d): %s",
                       len(ignored), max_lag, ignored)
    if not valid:
        raise ValueError(f"All tau candidates exceed the available trajectory length (max supported lag {max_lag}). "
                         "Provide smaller lag values or shorter horizons.")
    n_states = int(max(int(np.max(t)) for t in usable)) + 1
    shard = LabelShard.from_dtrajs(usable)
    evaluations = [_evaluate_single_lag(shard, lag, horizons, n_states, coverage_threshold, min_median_count,
                                        diag_mass_threshold, macro_lumper, n_timescales) for lag in sorted(valid)]
    for r in sorted(evaluations, key=lambda r: r.lag):
        if r.passed_sanity and r.ck_error <= ck_threshold:
            return r.lag, evaluations
    passing = [r for r in evaluations if r.passed_sanity]
    if passing:
        return min(passing, key=lambda r: r.ck_error).lag, evaluations
    return min(tau_candidates), evaluations
