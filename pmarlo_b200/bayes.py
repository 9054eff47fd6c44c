"""Bayesian MSM confidence intervals on the device.

Mirrors what ``ITSMixin.compute_implied_timescales`` does for every lag
(src/pmarlo/markov_state_model/_its.py:272-357): deeptime ``BayesianMSM(lagtime, n_samples)`` on the sliding
counts of the largest connected set -- reversible maximum-likelihood estimate as the starting point, then
``n_samples`` reversible transition matrices from the Gibbs sampler -- and ``_summarize_its_stats``
(:543-668) over the samples: eigenvalues lambda_2 .. lambda_{n+1}, timescales by ``safe_timescales``, rates;
median and the (1 - ci) / 2 percentiles.  Also ``sample_bayesian_timescales`` (:670-740).

Device work: K7 counts for all lags, K8 batched reversible MLE, ``pmb_bayes_rev_sample`` (one CTA per lag,
round-robin Gibbs sweeps), K9 batched eigenvalues of the lags x samples matrices.  Host work: the connected
set of each K x K count graph (scipy, as the reference) and the percentiles of the
(lags, samples, n_timescales) array.
"""

from __future__ import annotations

import warnings
from dataclasses import dataclass
from typing import Sequence

import numpy as np
import torch

from . import kernels
from .msm import (NUMERIC_MIN_POSITIVE, dtrajs_to_device, infer_n_states, largest_connected_set, safe_timescales)

__all__ = ["BayesianITS", "sample_reversible_matrices", "bayesian_implied_timescales", "summarize_its_stats"]


@dataclass
class BayesianITS:
    lag_times: np.ndarray
    eigenvalues: np.ndarray        # (n_lags, n_ts) medians
    eigenvalues_ci: np.ndarray     # (n_lags, n_ts, 2)
    timescales: np.ndarray
    timescales_ci: np.ndarray
    rates: np.ndarray
    rates_ci: np.ndarray
    active_sizes: np.ndarray
    timescale_samples: np.ndarray  # (n_lags, n_samples, n_ts)


def sample_reversible_matrices(C: torch.Tensor, active: torch.Tensor | None, n_samples: int, n_steps: int | None = None,
                               seed: int = 0, maxerr: float = 1e-8, maxiter: int = 1_000_000):
    """counts (B,K,K) fp64 [+ active masks (B,K) uint8] -> (T samples (B,S,K,K), pi samples (B,S,K), T_mle, pi_mle).
    Counts outside the active set are ignored (zeroed), like deeptime's restriction to the connected submodel."""
    Cb = (C.unsqueeze(0) if C.dim() == 2 else C).to(torch.float64).contiguous()
    B, K = int(Cb.shape[0]), int(Cb.shape[1])
    if active is not None:
        a = active.reshape(B, K).to(torch.float64)
        Cb = (Cb * a[:, :, None] * a[:, None, :]).contiguous()
        act = active.reshape(B, K).to(torch.uint8).contiguous()
    else:
        act = None
    T0, pi0, _ = kernels.mle_rev(Cb, act, alpha=0.0, maxerr=maxerr, maxiter=maxiter)
    Ts, pis = kernels.bayes_rev_sample(Cb, T0, pi0, n_samples, n_steps, seed)
    return Ts, pis, T0, pi0


def summarize_its_stats(lag: int, ev: np.ndarray, n_timescales: int, q_low: float, q_high: float):
    """``_summarize_its_stats`` (_its.py:543-668) from per-sample leading eigenvalues ``ev`` (S, >= n + 1) sorted by
    magnitude: re-sorted by value, lambda_2 .. lambda_{n+1} clipped to [1e-12, 1 - 1e-12]; medians and
    percentiles NaN-padded to n_timescales.  Returns the 9-tuple of the reference plus the (S, n) timescales."""
    n = int(max(0, n_timescales))
    S = ev.shape[0]
    e = np.full((S, n), np.nan)
    srt = -np.sort(-ev, axis=1)
    kk = min(n, srt.shape[1] - 1)
    if kk > 0:
        e[:, :kk] = np.clip(np.abs(srt[:, 1:1 + kk]), NUMERIC_MIN_POSITIVE, 1.0 - NUMERIC_MIN_POSITIVE)
    ts = np.full((S, n), np.nan)
    for s in range(S):
        if kk > 0:
            ts[s, :kk] = safe_timescales(int(max(1, lag)), e[s, :kk])
    with warnings.catch_warnings():
        warnings.filterwarnings("ignore", category=RuntimeWarning)
        rate = np.reciprocal(ts, where=np.isfinite(ts), out=np.full_like(ts, np.nan))
        out = []
        for arr in (e, ts, rate):
            out += [np.nanmedian(arr, axis=0), np.nanpercentile(arr, q_low, axis=0), np.nanpercentile(arr, q_high, axis=0)]
    return tuple(np.asarray(v, dtype=float) for v in out) + (ts,)


def bayesian_implied_timescales(dtrajs: Sequence[np.ndarray], lag_times: Sequence[int], n_states: int | None = None,
                                n_timescales: int = 5, n_samples: int = 100, ci: float = 0.95, seed: int | None = None,
                                n_steps: int | None = None, max_bytes: int = 4 << 30) -> BayesianITS:
    lags = [int(max(1, v)) for v in lag_times]
    K = infer_n_states(dtrajs, n_states)
    n_ts, S, L = int(n_timescales), int(max(1, n_samples)), len(lags)
    nan2 = np.full((L, n_ts), np.nan)
    nan3 = np.full((L, n_ts, 2), np.nan)
    res = BayesianITS(np.asarray(lags, dtype=int), nan2.copy(), nan3.copy(), nan2.copy(), nan3.copy(), nan2.copy(),
                      nan3.copy(), np.zeros(L, dtype=np.int64), np.full((L, S, n_ts), np.nan))
    if K == 0 or L == 0:
        return res
    dev = kernels.require_cuda()
    labels, segs = dtrajs_to_device(dtrajs, dev)
    off = segs.device(dev)
    alpha_tail = 50.0 * (1.0 - float(ci))
    q_low, q_high = alpha_tail, 100.0 - alpha_tail
    seed = int(np.random.SeedSequence().entropy % (2**63)) if seed is None else int(seed)
    per_lag = S * K * K * 8
    group = max(1, int(max_bytes // max(per_lag, 1)))
    k = min(n_ts + 1, K)
    for g0 in range(0, L, group):
        gl = lags[g0:g0 + group]
        B = len(gl)
        C = torch.zeros((B, K, K), dtype=torch.int64, device=dev)
        for b, lag in enumerate(gl):
            kernels.count_lagged(labels, off, K, lag, 1, out=C[b])
        Ch = C.cpu().numpy()
        active = np.zeros((B, K), dtype=np.uint8)
        for b in range(B):
            lcs = largest_connected_set(Ch[b])
            res.active_sizes[g0 + b] = lcs.size
            if lcs.size >= 2:
                active[b, lcs] = 1
        good = np.flatnonzero(active.sum(axis=1) >= 2)
        if good.size == 0:
            continue
        gi = torch.from_numpy(good).to(dev)
        Ts, pis, _, _ = sample_reversible_matrices(C[gi].to(torch.float64), torch.from_numpy(active[good]).to(dev), S,
                                                   n_steps, seed + g0)
        Bg = int(good.size)
        ev, _ = kernels.eig_rev_topk(Ts.reshape(Bg * S, K, K), pis.reshape(Bg * S, K), k)
        evh = ev.reshape(Bg, S, k).cpu().numpy()
        for row, b in enumerate(good):
            i = g0 + int(b)
            kk = int(min(k, res.active_sizes[i]))
            st = summarize_its_stats(gl[int(b)], evh[row, :, :kk], n_ts, q_low, q_high)
            res.eigenvalues[i], res.eigenvalues_ci[i, :, 0], res.eigenvalues_ci[i, :, 1] = st[0], st[1], st[2]
            res.timescales[i], res.timescales_ci[i, :, 0], res.timescales_ci[i, :, 1] = st[3], st[4], st[5]
            res.rates[i], res.rates_ci[i, :, 0], res.rates_ci[i, :, 1] = st[6], st[7], st[8]
            res.timescale_samples[i] = st[9]
    return res
