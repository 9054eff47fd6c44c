"""Lagged counts, reversible MSM, eigenvalues and implied timescales (K7-K10).

Mirrors
* ``build_simple_msm`` / ``_fit_msm_deeptime`` / ``_expand_results``
      src/pmarlo/markov_state_model/_msm_utils.py:163-281
* ``build_msm_from_labels``                       src/pmarlo/api/msm.py:455-488
* ``ensure_connected_counts``                     src/pmarlo/utils/msm_utils.py:129-167
* ``check_transition_matrix``                     src/pmarlo/utils/msm_utils.py:272-299
* ``safe_timescales``                             src/pmarlo/markov_state_model/utils.py:17-57
* the deterministic ITS sweep ``MaximumLikelihoodMSM(lagtime=lag, reversible=True)
  .fit(dtrajs).timescales()``                     src/pmarlo/markov_state_model/ck_its_selector.py:397-399
* the count pre-filter that splits trajectories at invalid labels
      src/pmarlo/markov_state_model/_estimation.py:121-145
"""

from __future__ import annotations

import logging
from dataclasses import dataclass
from typing import Sequence

import numpy as np
import torch

from . import kernels
from .distributed import Comm
from .shards import Segments

logger = logging.getLogger("pmarlo")

NUMERIC_DIRICHLET_ALPHA = 1e-3   # src/pmarlo/constants.py:44
NUMERIC_MIN_POSITIVE = 1e-12     # src/pmarlo/constants.py:29
NUMERIC_RELATIVE_TOLERANCE = 1e-8

__all__ = [
    "infer_n_states", "dtrajs_to_device", "split_at_invalid", "count_transitions_device",
    "count_transitions", "ensure_connected_counts", "ConnectedCountResult", "msm_from_counts_device",
    "build_simple_msm", "build_msm_from_labels", "check_transition_matrix", "safe_timescales",
    "largest_connected_set", "implied_timescales", "ITSResult", "eigenvalues_rev",
]


# --------------------------------------------------------------------------- host bookkeeping
def infer_n_states(dtrajs: Sequence[np.ndarray], n_states: int | None = None) -> int:
    """``_infer_n_states`` (_msm_utils.py:190-207)."""
    if n_states is not None:
        return int(n_states)
    mx = -1
    for dt in dtrajs:
        dt = np.asarray(dt)
        if dt.size:
            m = int(dt.max())
            if m >= 0:
                mx = max(mx, m)
    return mx + 1 if mx >= 0 else 0


def split_at_invalid(dtrajs: Sequence[np.ndarray], n_states: int) -> list[np.ndarray]:
    """``_estimation.py:121-145``: cut every trajectory at labels outside [0, n_states)."""
    out = []
    for d in dtrajs:
        arr = np.asarray(d, dtype=np.int64)
        if arr.size == 0:
            continue
        valid = (arr >= 0) & (arr < n_states)
        if not valid.any():
            continue
        edges = np.flatnonzero(np.diff(np.concatenate([[0], valid.view(np.int8), [0]])))
        for s, e in zip(edges[::2], edges[1::2]):
            out.append(arr[s:e])
    return out


def dtrajs_to_device(dtrajs: Sequence[np.ndarray], device) -> tuple[torch.Tensor, Segments]:
    arrs = [np.asarray(d).astype(np.int32, copy=False).reshape(-1) for d in dtrajs]
    segs = Segments.from_lengths([a.size for a in arrs])
    flat = np.concatenate(arrs) if arrs else np.zeros((0,), dtype=np.int32)
    return torch.from_numpy(np.ascontiguousarray(flat)).to(device), segs


# --------------------------------------------------------------------------- K7
def count_transitions_device(labels: torch.Tensor, seg_offsets: torch.Tensor, n_states: int, lag: int,
                             step: int = 1, comm: Comm | None = None,
                             out: torch.Tensor | None = None) -> torch.Tensor:
    """(K,K) int64 sliding-window count matrix of a label shard, summed over ranks."""
    C = kernels.count_lagged(labels, seg_offsets, n_states, lag, step, out=out)
    (comm if comm is not None else Comm()).allreduce_sum(C)
    return C


def count_transitions(dtrajs: Sequence[np.ndarray], n_states: int | None = None, lag: int = 1,
                      count_mode: str = "sliding", split_invalid: bool = False) -> np.ndarray:
    """``TransitionCountEstimator(lagtime, count_mode).fit(dtrajs).count_matrix`` as a
    dense float64 array.  ``count_mode="strided"`` steps by ``lag`` inside every shard.
    deeptime sizes the matrix by ``max label + 1``."""
    if count_mode not in ("sliding", "strided"):
        raise ValueError(f"unsupported count_mode {count_mode!r}")
    K = max(infer_n_states(dtrajs, n_states), infer_n_states(dtrajs))
    if K == 0:
        return np.zeros((0, 0), dtype=float)
    if split_invalid:
        dtrajs = split_at_invalid(dtrajs, K)
    if not dtrajs:
        return np.zeros((K, K), dtype=float)
    dev = kernels.require_cuda()
    labels, segs = dtrajs_to_device(dtrajs, dev)
    lag = int(max(1, lag))
    C = kernels.count_lagged(labels, segs.device(dev), K, lag, lag if count_mode == "strided" else 1)
    return C.cpu().numpy().astype(float)


# --------------------------------------------------------------------------- regularisation
@dataclass
class ConnectedCountResult:
    counts: np.ndarray
    active: np.ndarray

    def to_dict(self) -> dict:
        return {"counts": self.counts.tolist(), "active": self.active.tolist()}


def ensure_connected_counts(C, alpha: float = NUMERIC_DIRICHLET_ALPHA,
                            epsilon: float = NUMERIC_MIN_POSITIVE) -> ConnectedCountResult:
    """``utils/msm_utils.py:129-167`` (host view of what ``msm_from_counts_device`` applies
    on the device through the ``active`` mask and ``alpha``)."""
    C = np.asarray(C)
    if C.ndim != 2 or C.shape[0] != C.shape[1]:
        raise ValueError("count matrix must be square")
    totals = C.sum(axis=1) + C.sum(axis=0)
    active = np.where(totals > epsilon)[0]
    if active.size == 0:
        return ConnectedCountResult(np.empty((0, 0), dtype=float), active)
    Ca = C[np.ix_(active, active)].astype(float)
    Ca += float(alpha)
    return ConnectedCountResult(Ca, active)


# --------------------------------------------------------------------------- K8
def msm_from_counts_device(C: torch.Tensor, alpha: float = NUMERIC_DIRICHLET_ALPHA,
                           epsilon: float = NUMERIC_MIN_POSITIVE, maxerr: float = 1e-8,
                           maxiter: int = 1_000_000, active: torch.Tensor | None = None):
    """int64 counts (K,K) -> (T, pi, info, active): trim states without counts, add the
    Dirichlet pseudocount to the active block, reversible MLE, T = I / pi = 0 on
    the inactive states (``_expand_results``) -- all on the device."""
    Cf, act = kernels.counts_active(C, epsilon)
    if active is not None:
        act = active
    T, pi, info = kernels.mle_rev(Cf, act, alpha=alpha, maxerr=maxerr, maxiter=maxiter)
    return T, pi, info, act


def check_transition_matrix(T, pi, *, row_tol: float = NUMERIC_MIN_POSITIVE,
                            stat_tol: float = NUMERIC_RELATIVE_TOLERANCE) -> None:
    """``utils/msm_utils.py:272-299``."""
    T = np.asarray(T, dtype=float)
    pi = np.asarray(pi, dtype=float)
    if T.ndim != 2 or T.shape[0] != T.shape[1]:
        raise ValueError("transition matrix must be square")
    if pi.shape != (T.shape[0],):
        raise ValueError("stationary distribution size mismatch")
    if T.size == 0:
        return
    if np.any(T < 0.0):
        raise ValueError("Negative probabilities in transition matrix")
    if np.max(np.abs(T.sum(axis=1) - 1.0)) > row_tol:
        raise ValueError("transition matrix fails stochasticity checks")
    s = float(pi.sum())
    if not np.isfinite(s) or s <= 0:
        raise ValueError("stationary distribution must be normalisable")
    pn = pi / s
    if float(np.max(np.abs(pn @ T - pn))) > stat_tol:
        raise ValueError("provided stationary distribution fails invariance check")


def build_simple_msm(dtrajs: list[np.ndarray], n_states: int | None = None, lag: int = 20,
                     count_mode: str = "sliding") -> tuple[np.ndarray, np.ndarray]:
    """Drop-in for ``pmarlo.markov_state_model._msm_utils.build_simple_msm``."""
    if not dtrajs:
        logger.error("build_simple_msm: No dtrajs provided")
        return np.empty((0, 0), dtype=float), np.empty((0,), dtype=float)
    if count_mode not in ("sliding", "strided"):
        raise ValueError(f"unsupported count_mode {count_mode!r}")
    n_states = infer_n_states(dtrajs, n_states)
    logger.info(f"build_simple_msm: Using {n_states} states")
    lag = int(max(1, lag))
    K = max(n_states, infer_n_states(dtrajs))
    if K == 0:
        return np.eye(0), np.zeros((0,))
    dev = kernels.require_cuda()
    labels, segs = dtrajs_to_device(dtrajs, dev)
    C = kernels.count_lagged(labels, segs.device(dev), K, lag, lag if count_mode == "strided" else 1)
    T, pi, info, act = msm_from_counts_device(C)
    if int(act.sum().item()) == 0:
        return np.eye(K), np.zeros((K,))
    if int(info[1].item()) < 0:
        raise ValueError("reversible MLE failed: a state of the active set has no outgoing counts")
    Tn, pin = T.cpu().numpy(), pi.cpu().numpy()
    check_transition_matrix(Tn, pin)
    return Tn, pin


def build_msm_from_labels(dtrajs: list[np.ndarray], n_states: int | None = None,
                          lag: int = 20) -> tuple[np.ndarray, np.ndarray]:
    """Drop-in for ``pmarlo.api.msm.build_msm_from_labels``."""
    logger.info("[msm] Building MSM from labels: n_trajectories=%d, lag=%d", len(dtrajs), lag)
    T, pi = build_simple_msm(dtrajs, n_states=n_states, lag=lag)
    logger.info("[msm] MSM built: transition_matrix_shape=%s, stationary_dist_shape=%s", T.shape, pi.shape)
    return T, pi


# --------------------------------------------------------------------------- K9 / K10
def safe_timescales(lag: float, eigvals, eps: float = NUMERIC_MIN_POSITIVE) -> np.ndarray:
    """``markov_state_model/utils.py:17-57``: t = -lag / ln|lambda| with |lambda| clipped to
    [eps, 1-eps]; NaN for eigenvalues outside (0, 1)."""
    eig = np.asarray(eigvals)
    if eig.size == 0:
        return np.empty_like(eig, dtype=np.float64)
    ec = eig.astype(np.complex128)
    mag = np.abs(ec)
    with np.errstate(divide="ignore", invalid="ignore"):
        ts = -float(lag) / np.log(np.clip(mag, eps, 1 - eps))
    ts = np.asarray(ts, dtype=np.float64)
    invalid = ~np.isfinite(mag) | (mag <= 0) | (mag >= 1)
    real = np.isclose(ec.imag, 0.0)
    invalid |= real & ((ec.real <= 0.0) | (ec.real >= 1.0))
    ts[invalid] = np.nan
    return ts


def eigenvalues_rev(T: np.ndarray, pi: np.ndarray, k: int | None = None) -> np.ndarray:
    """deeptime ``eigenvalues(T, k, reversible=True, mu=pi)``: spectrum of D^1/2 T D^-1/2 by magnitude."""
    dev = kernels.require_cuda()
    Td = torch.from_numpy(np.ascontiguousarray(T, dtype=np.float64)).to(dev)
    pid = torch.from_numpy(np.ascontiguousarray(pi, dtype=np.float64)).to(dev)
    K = int(Td.shape[0])
    kk = K if k is None else min(int(k), K)
    ev, _ = kernels.eig_rev_topk(Td, pid, kk)
    return ev.cpu().numpy()


def largest_connected_set(C: np.ndarray) -> np.ndarray:
    """Largest strongly connected set of the count graph (graph bookkeeping on the
    K x K matrix, scipy like src/pmarlo/utils/scc.py:69-130)."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import connected_components

    n, comp = connected_components(csr_matrix(np.asarray(C) > 0), directed=True, connection="strong")
    sizes = np.bincount(comp, minlength=n)
    return np.flatnonzero(comp == int(np.argmax(sizes)))


@dataclass
class ITSResult:
    """Fields of ``results.ITSResult`` that the deterministic sweep fills."""

    lag_times: np.ndarray
    timescales: np.ndarray      # (n_lags, n_timescales), NaN padded
    eigenvalues: np.ndarray     # (n_lags, n_timescales + 1)
    iterations: np.ndarray      # MLE iterations per lag
    active_sizes: np.ndarray


def implied_timescales(dtrajs: Sequence[np.ndarray], lag_times: Sequence[int], n_states: int | None = None,
                       n_timescales: int = 5, maxerr: float = 1e-8, maxiter: int = 1_000_000) -> ITSResult:
    """Deterministic ITS sweep: per lag raw sliding counts -> largest connected set ->
    reversible MLE -> leading eigenvalues -> ``safe_timescales``.  All lags are
    counted, estimated and diagonalised as ONE batch (one CTA per lag for K <= 2048)."""
    lags = [int(v) for v in lag_times]
    K = infer_n_states(dtrajs, n_states)
    out_ts = np.full((len(lags), n_timescales), np.nan)
    out_ev = np.full((len(lags), n_timescales + 1), np.nan)
    iters = np.zeros(len(lags), dtype=np.int64)
    sizes = np.zeros(len(lags), dtype=np.int64)
    if K == 0 or not lags:
        return ITSResult(np.asarray(lags), out_ts, out_ev, iters, sizes)
    dev = kernels.require_cuda()
    labels, segs = dtrajs_to_device(dtrajs, dev)
    off = segs.device(dev)
    B = len(lags)
    C = torch.zeros((B, K, K), dtype=torch.int64, device=dev)
    for b, lag in enumerate(lags):
        kernels.count_lagged(labels, off, K, lag, 1, out=C[b])
    Ch = C.cpu().numpy()
    active = np.zeros((B, K), dtype=np.uint8)
    for b in range(B):
        lcs = largest_connected_set(Ch[b])
        if lcs.size >= 2:
            active[b, lcs] = 1
        sizes[b] = lcs.size
    good = np.flatnonzero(sizes >= 2)
    if good.size == 0:
        return ITSResult(np.asarray(lags), out_ts, out_ev, iters, sizes)
    Cg = C[torch.from_numpy(good).to(dev)].to(torch.float64)
    act = torch.from_numpy(active[good]).to(dev)
    T, pi, info = kernels.mle_rev(Cg, act, alpha=0.0, maxerr=maxerr, maxiter=maxiter)
    k = min(n_timescales + 1, K)
    ev, _ = kernels.eig_rev_topk(T, pi, k)
    evh, infoh = ev.cpu().numpy(), info.cpu().numpy()
    for row, b in enumerate(good):
        kk = int(min(k, sizes[b]))
        out_ev[b, :kk] = evh[row, :kk]
        ts = safe_timescales(lags[b], evh[row, 1:kk])
        out_ts[b, : ts.size] = ts
        iters[b] = infoh[row, 0]
    return ITSResult(np.asarray(lags), out_ts, out_ev, iters, sizes)


# ----------------------------------------------------------------------------- lag ladders / deterministic ITS
_LAG_LADDER = (1, 2, 3, 5, 8, 10, 15, 20, 30, 40, 50, 75, 80, 100, 150, 160, 200, 300, 320, 500, 640, 750, 1000,
               1280, 1500, 2000)


def candidate_lag_ladder(min_lag: int = 1, max_lag: int = 200, n_candidates: int | None = None) -> list[int]:
    """``pmarlo.utils.msm_utils.candidate_lag_ladder`` (utils/msm_utils.py:21-105): the curated ladder of
    "nice" lags inside [min_lag, max_lag], optionally thinned to ``n_candidates`` roughly evenly spaced
    entries with both endpoints kept.  Same exceptions and messages."""
    lo, hi = int(min_lag), int(max_lag)
    if lo < 1:
        raise ValueError("min_lag must be >= 1")
    if hi < lo:
        raise ValueError("max_lag must be >= min_lag")
    if n_candidates is not None and n_candidates < 1:
        raise ValueError("n_candidates must be positive")
    filtered = [x for x in _LAG_LADDER if lo <= x <= hi]
    if not filtered:
        raise ValueError(f"No predefined lag values available in range [{lo}, {hi}]")
    if n_candidates is None or n_candidates >= len(filtered):
        return filtered
    if n_candidates == 1:
        return [filtered[0]]
    if n_candidates == 2:
        return [filtered[0], filtered[-1]]
    step = (len(filtered) - 1) / (n_candidates - 1)
    picks = sorted({int(round(i * step)) for i in range(n_candidates)})
    picks[0] = 0
    picks[-1] = len(filtered) - 1
    return [filtered[i] for i in picks]


def deterministic_its_from_counts(lag: int, counts, n_timescales: int):
    """``ITSMixin._deterministic_its_from_counts`` (_its.py:742-801), the fallback that fills lags whose
    Bayesian estimate has no finite timescale: T = row-normalised (C + C^T) / 2, the weights
    ``mu = max(rowsum(T), 1e-15)`` normalised (uniform over non-empty rows -- what the reference passes as
    ``mu``), eigenvalues of deeptime's reversible branch = LAPACK ``eigvalsh`` of
    S = sqrt(mu)_i T_ij / sqrt(mu)_j, which reads the LOWER triangle only; sorted by real part, clipped to
    [1e-12, 1 - 1e-12]; timescales -lag / ln|lambda| from the magnitude-sorted spectrum (infinite where
    |lambda| = 1 to 1e-14).  Returns (eigenvalues, timescales, rates), each of length ``n_timescales``.
    The K x K symmetric eigenproblem runs in libpmb200 (``pmb_sym_eigvals_batched``, cyclic Jacobi)."""
    dev = kernels.require_cuda()
    C = torch.as_tensor(np.asarray(counts, dtype=np.float64)).to(dev)
    K = int(C.shape[0])
    n_ts = int(n_timescales)
    C_rev = 0.5 * (C + C.T)
    row = C_rev.sum(dim=1, keepdim=True)
    T = C_rev / torch.where(row == 0, torch.ones_like(row), row)
    mu = torch.clamp(T.sum(dim=1), min=1e-15)
    mu = mu / mu.sum()
    smu = torch.sqrt(mu)
    S = smu[:, None] * T / smu[None, :]
    L = torch.tril(S)
    sym = L + torch.tril(S, -1).T
    lam = kernels.sym_eigvals_batched(sym)[0].cpu().numpy()          # by magnitude, descending
    k_request = n_ts + 1 if n_ts > 0 else None
    k_eval = None if (k_request is not None and k_request > K) else k_request
    ev = lam if k_eval is None else lam[:k_eval]
    ev_sorted = ev[np.argsort(-ev, kind="stable")]
    slow = ev_sorted[1:1 + n_ts] if n_ts > 0 else np.empty((0,))
    slow = np.clip(np.abs(slow), NUMERIC_MIN_POSITIVE, 1.0 - NUMERIC_MIN_POSITIVE)
    evals = np.zeros((n_ts,), dtype=float)
    evals[: slow.shape[0]] = slow
    if n_ts > 0:
        k_times = None if k_request is None else min(K, n_ts + 1)
        evt = lam if k_times is None else lam[:k_times]
        ts_raw = np.zeros(evt.shape[0])
        one = np.isclose(np.abs(evt), 1.0, rtol=0.0, atol=1e-14)
        ts_raw[one] = np.inf
        with np.errstate(divide="ignore", invalid="ignore"):
            ts_raw[~one] = -float(max(1, int(lag))) / np.log(np.abs(evt[~one]))
        ts_arr = ts_raw[1:1 + n_ts]
    else:
        ts_arr = np.empty((0,))
    if ts_arr.shape[0] < n_ts:
        ts_arr = np.pad(ts_arr, (0, n_ts - ts_arr.shape[0]), mode="constant", constant_values=np.nan)
    rates = np.reciprocal(ts_arr, where=np.isfinite(ts_arr), out=np.full_like(ts_arr, np.nan))
    return evals, ts_arr, rates
