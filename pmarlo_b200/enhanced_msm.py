"""``EnhancedMSM.build_msm`` / ``.compute_implied_timescales`` behind pmarlo's mixin interface.

Mirrors:
* ``EstimationMixin.build_msm(lag_time=20, method="standard")``  markov_state_model/_estimation.py:50-83
  -> ``_validate_and_cap_lag`` :99-114, the split of every dtraj at labels outside [0, n_states) :121-145,
  sliding counts :150-156, ``ensure_connected_counts`` + NON-reversible maximum likelihood (row
  normalisation, stationary vector of the active block) embedded with T = I, pi = 0 on inactive states
  :158-188, free energies -kT ln pi :211-220.
* ``ITSMixin.compute_implied_timescales(lag_times=None, n_timescales=5, *, n_samples, ci, dirichlet_alpha,
  plateau_m, plateau_epsilon)``  _its.py:137-192 with its input validation :453-524 and default lag ladder
  :425-451; result container ``results.ITSResult`` :136-147.
* attributes of ``_base.py:59-99``: dtrajs, n_states, count_matrix, transition_matrix,
  stationary_distribution, implied_timescales, count_mode, lag_time.

The reference estimates the ITS with deeptime's ``BayesianMSM`` (100 reversible samples per lag); the Gibbs
sampler is SURVEY.md 8f item 2 ("next").  Here every lag gets the deterministic reversible maximum-
likelihood estimate of ck_its_selector.py:397-399 (one batched device sweep, ``msm.implied_timescales``);
the ``*_ci`` fields are NaN.  Counting (K7), the reversible MLE (K8) and the eigenvalues (K9) run in
libpmb200; the stationary vector of the non-reversible ``build_msm`` is one K x K linear solve on the
device (torch.linalg, a plain library call on a cold path).
"""

from __future__ import annotations

import logging
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import kernels
from .msm import (NUMERIC_DIRICHLET_ALPHA, NUMERIC_MIN_POSITIVE, count_transitions, ensure_connected_counts,
                  implied_timescales, infer_n_states)
from .msm import safe_timescales as safe_timescales_np

logger = logging.getLogger("pmarlo")

__all__ = ["EnhancedMSM", "ITSResultCI", "DEFAULT_ITS_LAGS"]

DEFAULT_ITS_LAGS = [1, 2, 3, 5, 8, 10, 15, 20, 30, 40, 50, 75, 80, 100, 150, 160, 200, 320, 640, 1280]   # _its.py:425-451


@dataclass
class ITSResultCI:
    """``results.ITSResult`` (results.py:136-147)."""

    lag_times: np.ndarray
    eigenvalues: np.ndarray
    eigenvalues_ci: np.ndarray
    timescales: np.ndarray
    timescales_ci: np.ndarray
    rates: np.ndarray
    rates_ci: np.ndarray
    recommended_lag_window: Optional[tuple] = None


def _empty_its(n_timescales: int) -> ITSResultCI:
    z2 = np.empty((0, n_timescales), dtype=float)
    z3 = np.empty((0, n_timescales, 2), dtype=float)
    return ITSResultCI(np.array([], dtype=int), z2.copy(), z3.copy(), z2.copy(), z3.copy(), z2.copy(), z3.copy())


def stationary_nonreversible(T: np.ndarray) -> np.ndarray:
    """Left Perron vector of a row-stochastic matrix: solve [(T^T - I); 1^T] pi = e_last on the device."""
    n = int(T.shape[0])
    if n == 0:
        return np.empty((0,))
    dev = kernels.require_cuda()
    Td = torch.from_numpy(np.ascontiguousarray(T, dtype=np.float64)).to(dev)
    A = Td.t() - torch.eye(n, dtype=torch.float64, device=dev)
    A[n - 1, :] = 1.0                       # one balance equation is redundant: replace it by sum(pi) = 1
    b = torch.zeros((n,), dtype=torch.float64, device=dev)
    b[n - 1] = 1.0
    pi = torch.linalg.solve(A, b).abs()
    return (pi / pi.sum()).cpu().numpy()


class EnhancedMSM:
    """The estimation / ITS slice of pmarlo's ``EnhancedMSM`` (markov_state_model/enhanced_msm.py)."""

    def __init__(self, dtrajs: Sequence[np.ndarray] | None = None, n_states: int | None = None,
                 count_mode: str = "sliding", random_state: int | None = None, temperature: float = 300.0,
                 output_dir: str | None = None):
        self.dtrajs: List[np.ndarray] = [np.asarray(d) for d in (dtrajs or [])]
        self.n_states = infer_n_states(self.dtrajs, n_states)
        self.count_mode = str(count_mode)
        self.random_state = random_state
        self.temperature = float(temperature)
        self.lag_time = 20
        self.features = None
        self.count_matrix = None
        self.transition_matrix = None
        self.stationary_distribution = None
        self.free_energies = None
        self.implied_timescales = None
        self.estimator_backend = "b200"
        self.output_dir = output_dir

    # ---------------------------------------------------------------- build_msm
    def build_msm(self, lag_time: int = 20, method: str = "standard") -> None:
        lag_time = int(max(1, lag_time))
        logger.info("Building MSM with lag time %d using %s method...", lag_time, method)
        if int(self.n_states) <= 0:
            raise ValueError("Cannot build a Markov state model without defined microstates; "
                             "ensure clustering produced at least one state.")
        self.lag_time = lag_time
        if method != "standard":
            raise ValueError(f"Unknown MSM method: {method}")
        eff = getattr(self, "effective_frames", 0)
        if eff and lag_time >= eff:
            raise ValueError(f"lag_time {lag_time} exceeds available effective frames {eff}")
        K = int(self.n_states)
        max_valid = min(len(d) for d in self.dtrajs) - 1 if self.dtrajs else 0
        lag = lag_time
        if lag > max_valid > 0:
            logger.warning("Lag %s exceeds max feasible %s; capping", lag, max_valid)
            lag = max_valid
        if max_valid < 1:
            self.count_matrix = np.zeros((K, K))
            self.transition_matrix = np.eye(K)
            self.stationary_distribution = np.zeros((K,))
        else:
            mode = "sliding" if self.count_mode == "strided" else self.count_mode    # _estimation.py:152
            inside = [np.where((np.asarray(d) >= 0) & (np.asarray(d) < K), np.asarray(d), -1) for d in self.dtrajs]
            C = count_transitions(inside, K, lag, count_mode=mode, split_invalid=True)[:K, :K]
            self._finalize_transition_and_stationary(C)
        self._compute_free_energies(self.temperature)
        logger.info("MSM construction completed")

    def _finalize_transition_and_stationary(self, counts: np.ndarray) -> None:
        K = int(self.n_states)
        res = ensure_connected_counts(counts)
        cm = np.zeros((K, K))
        T_full, pi_full = np.eye(K), np.zeros((K,))
        if res.counts.size:
            act = np.asarray(res.active)
            cm[np.ix_(act, act)] = res.counts
            rs = res.counts.sum(axis=1, keepdims=True)
            T_act = res.counts / rs                       # every cell carries +alpha: no empty row
            T_full[np.ix_(act, act)] = T_act
            pi_full[act] = stationary_nonreversible(T_act)
        self.count_matrix, self.transition_matrix, self.stationary_distribution = cm, T_full, pi_full

    def _compute_free_energies(self, temperature: float = 300.0) -> None:
        if self.stationary_distribution is None:
            raise ValueError("Stationary distribution must be computed first")
        kT = 1.380649e-23 * temperature * 6.02214076e23 / 1000.0        # kJ/mol
        fe = -kT * np.log(np.maximum(self.stationary_distribution, NUMERIC_MIN_POSITIVE))
        self.free_energies = fe - np.min(fe) if fe.size else fe

    # ---------------------------------------------------------------- implied timescales
    def compute_implied_timescales(self, lag_times: Optional[List[int]] = None, n_timescales: int = 5, *,
                                   n_samples: int = 100, ci: float = 0.95,
                                   dirichlet_alpha: float = NUMERIC_DIRICHLET_ALPHA,
                                   plateau_m: int | None = None, plateau_epsilon: float = 0.1,
                                   estimator: str = "bayesian") -> None:
        """``ITSMixin.compute_implied_timescales`` (_its.py:137-192).  ``estimator="bayesian"`` (the reference's
        behaviour): per lag ``n_samples`` reversible transition matrices from the device Gibbs sampler, medians and
        ``ci`` percentiles of eigenvalues / timescales / rates, lags without any finite timescale filled by
        ``_deterministic_its_from_counts`` (:403-423).  ``estimator="mle"`` (extension, also used when
        ``n_samples <= 0``): the deterministic reversible maximum-likelihood sweep of ck_its_selector.py:397-399,
        CI fields NaN."""
        if estimator not in ("bayesian", "mle"):
            raise ValueError(f"unknown estimator {estimator!r}")
        lags = list(DEFAULT_ITS_LAGS) if lag_times is None else [int(max(1, v)) for v in lag_times]
        if not self.dtrajs:
            logger.warning("No trajectories available for implied timescales")
            self.implied_timescales = _empty_its(n_timescales)
            return
        max_valid = min(len(d) for d in self.dtrajs) - 1
        if max_valid < 1:
            logger.warning("Trajectories too short for implied timescales")
            self.implied_timescales = _empty_its(n_timescales)
            return
        if any(v > max_valid for v in lags):
            logger.warning("Capping lag times above max_valid_lag=%s", max_valid)
        lags = [v for v in lags if 1 <= v <= max_valid]
        if not lags:
            logger.warning("No valid lag times after capping")
            self.implied_timescales = _empty_its(n_timescales)
            return
        eff = getattr(self, "effective_frames", None)
        if eff is not None and eff > 0 and max(lags) >= eff:
            raise ValueError(f"Maximum lag {max(lags)} exceeds available effective frames {eff}")
        if estimator == "bayesian" and int(n_samples) > 0:
            from .bayes import bayesian_implied_timescales
            from .msm import deterministic_its_from_counts

            logger.info("Computing implied timescales with Bayesian estimation")
            if "effective" not in str(self.count_mode).lower():
                logger.warning("Bayesian MSM confidence sampling expects effective counts; continuing with "
                               "count_mode='%s' may yield correlated samples.", self.count_mode)
            b = bayesian_implied_timescales(self.dtrajs, lags, n_states=self.n_states, n_timescales=n_timescales,
                                            n_samples=int(n_samples), ci=float(ci),
                                            seed=getattr(self, "random_state", None))
            result = ITSResultCI(np.asarray(lags, dtype=int), b.eigenvalues, b.eigenvalues_ci, b.timescales,
                                 b.timescales_ci, b.rates, b.rates_ci)
            for i, lag in enumerate(lags):                   # _its_fill_missing_timescales
                if np.any(np.isfinite(result.timescales[i])):
                    continue
                C = count_transitions(self.dtrajs, self.n_states, lag)
                cc = ensure_connected_counts(C, alpha=dirichlet_alpha)
                if cc.counts.size == 0:
                    raise RuntimeError(f"Transition counts empty for lag {lag}")
                ev_d, ts_d, rate_d = deterministic_its_from_counts(int(lag), cc.counts, n_timescales)
                result.eigenvalues[i], result.timescales[i], result.rates[i] = ev_d, ts_d, rate_d
            ts = result.timescales
        else:
            sweep = implied_timescales(self.dtrajs, lags, n_states=self.n_states, n_timescales=n_timescales)
            ts = sweep.timescales
            ev = np.full((len(lags), n_timescales), np.nan)
            ev[:, : max(0, sweep.eigenvalues.shape[1] - 1)] = sweep.eigenvalues[:, 1 : 1 + n_timescales]
            ev = np.where(np.isfinite(ev), np.clip(np.abs(ev), NUMERIC_MIN_POSITIVE, 1.0 - NUMERIC_MIN_POSITIVE), ev)
            with np.errstate(divide="ignore", invalid="ignore"):
                rates = np.where(np.isfinite(ts), 1.0 / ts, np.nan)
            nan_ci = np.full((len(lags), n_timescales, 2), np.nan)
            result = ITSResultCI(np.asarray(lags, dtype=int), ev, nan_ci.copy(), ts, nan_ci.copy(), rates, nan_ci.copy())
        if plateau_m is not None:
            result.recommended_lag_window = _plateau_window(lags, ts, int(plateau_m), float(plateau_epsilon))
        self.implied_timescales = result

    def sample_bayesian_timescales(self, n_samples: int = 200, count_mode: str = "effective"):
        """``ITSMixin.sample_bayesian_timescales`` (_its.py:670-740): timescale and population samples of the
        reversible posterior at ``self.lag_time``.  ``count_mode="effective"`` (deeptime's statistically
        uncorrelated counts) is not offered: sliding counts are used and the samples are narrower than deeptime's,
        which is logged."""
        from .bayes import sample_reversible_matrices
        from .msm import largest_connected_set

        if not self.dtrajs:
            return None
        if "effective" in str(count_mode).lower():
            logger.warning("count_mode='effective' is not available on the device path; using sliding counts")
        lag = int(max(1, self.lag_time))
        C = count_transitions(self.dtrajs, self.n_states, lag)
        lcs = largest_connected_set(C)
        if lcs.size < 2:
            return None
        dev = kernels.require_cuda()
        act = np.zeros((1, C.shape[0]), dtype=np.uint8)
        act[0, lcs] = 1
        Ts, pis, _, _ = sample_reversible_matrices(torch.from_numpy(np.asarray(C, dtype=np.float64)).to(dev),
                                                   torch.from_numpy(act).to(dev), int(max(1, n_samples)),
                                                   seed=int(getattr(self, "random_state", None) or 0))
        K = int(C.shape[0])
        k = min(6, int(lcs.size))
        ev, _ = kernels.eig_rev_topk(Ts.reshape(-1, K, K), pis.reshape(-1, K), k)
        evh = -np.sort(-ev.cpu().numpy(), axis=1)
        ts_list = []
        for row in evh:
            ts = safe_timescales_np(lag, row[1:k])
            ts = ts[np.isfinite(ts)]
            if ts.size:
                ts_list.append(ts)
        out = {}
        if ts_list:
            m = max(a.shape[0] for a in ts_list)
            out["timescales_samples"] = np.vstack([np.pad(a, (0, m - a.shape[0]), constant_values=np.nan) for a in ts_list])
        out["population_samples"] = pis[0].cpu().numpy()
        return out


def _ck_methods():
    """``CKMixin`` (_ck.py:61-110, 159-175) on the device kernels; see pmarlo_b200/ck.py."""
    from . import ck

    def compute_ck_test_micro(self, factors: Optional[List[int]] = None, max_states: int = 50,
                              min_transitions: int = 5) -> "ck.CKTestResult":
        return ck.compute_ck_test_micro(self.dtrajs, int(self.n_states), int(self.lag_time), factors,
                                        max_states, min_transitions)

    def select_lag_time_ck(self, tau_candidates: List[int], factor: int = 2, mse_epsilon: float = 0.05) -> int:
        selected, _, _, _ = ck.select_lag_time_ck(self.dtrajs, int(self.n_states), tau_candidates, factor,
                                                  mse_epsilon, output_dir=self.output_dir)
        self.lag_time = int(selected)
        print(f"Selected τ = {int(selected)}")       # _ck.py:261
        return int(selected)

    EnhancedMSM.compute_ck_test_micro = compute_ck_test_micro
    EnhancedMSM.select_lag_time_ck = select_lag_time_ck


_ck_methods()


def _plateau_window(lags: Sequence[int], ts: np.ndarray, m: int, eps: float):
    """First run of ``m`` consecutive lags over which the slowest timescale changes by less than ``eps``
    (relative): the window the reference attaches as ``recommended_lag_window``."""
    slow = ts[:, 0] if ts.size else np.empty((0,))
    for s in range(0, len(lags) - m + 1):
        w = slow[s : s + m]
        if np.all(np.isfinite(w)) and np.max(w) > 0 and (np.max(w) - np.min(w)) / np.max(w) <= eps:
            return (float(lags[s]), float(lags[s + m - 1]))
    return None
