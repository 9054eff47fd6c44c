"""Oracle: preprocessing + TICA (TEST INFRASTRUCTURE).

Restates in numpy/scipy fp64:

* ``_preprocess`` -- src/pmarlo/markov_state_model/reduction.py:13-40
  (sklearn SimpleImputer(mean) -> StandardScaler(with_std=scale) -> nan_to_num).
  Pinned against the importable reference function in tests/golden.
* deeptime 0.4.5 ``TICA(lagtime, dim)`` with defaults epsilon=1e-6,
  scaling="kinetic_map" as called at reduction.py:103-110,
  _features.py:181-231 and cv/__init__.py:42-50.  deeptime is absent from
  this image; the algorithm below follows its published implementation
  (``covariance.Covariance(reversible=True, remove_data_mean=True,
  bessels_correction=False)`` + ``numeric.eigen.spd_inv_split`` /
  ``eig_corr(canonical_signs=True)``).  PARITY UNPINNED against deeptime
  itself; checked against analytic AR(1) spectra, the invariants
  L^T C00 L = I and R^T C00 R = I, and the reference's own numpy
  cross-check ``_estimate_top_eigenvalues``
  (src/pmarlo/features/deeptica/core/trainer_api.py:632-656).
"""

from __future__ import annotations

from typing import Sequence

import numpy as np
import scipy.linalg

__all__ = [
    "preprocess",
    "scaler_stats",
    "lagged_covariances",
    "spd_inv_split",
    "eig_corr",
    "tica_fit",
    "tica_transform",
    "tica_reduce",
    "maybe_apply_tica",
    "TicaModel",
]


def scaler_stats(X: np.ndarray, scale: bool = True):
    """Mean (NaN-aware), population std with sklearn's constant-feature rule.

    Returns (mean, scale_) such that preprocess(X) == (impute(X) - mean) / scale_.
    sklearn ``StandardScaler``: var ddof=0, features whose variance is within
    rounding of zero get scale 1 (``_is_constant_feature``).
    """
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    nan = np.isnan(X)
    cnt = (~nan).sum(axis=0)
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = np.where(cnt > 0, np.nansum(X, axis=0) / np.maximum(cnt, 1), 0.0)
    Xi = np.where(nan, mean[None, :], X)
    m = Xi.mean(axis=0)
    var = ((Xi - m) ** 2).mean(axis=0)
    if scale:
        eps = np.finfo(np.float64).eps
        upper = n * eps * var + (n * m * eps) ** 2
        const_mask = var <= upper
        sc = np.sqrt(var)
        sc = np.where(const_mask | (sc < 10 * eps), 1.0, sc)
    else:
        sc = np.ones_like(var)
    return mean, m, sc


def preprocess(X: np.ndarray, scale: bool = True) -> np.ndarray:
    """reduction.py:13-40."""
    Xp = np.asarray(X, dtype=float)
    if Xp.size == 0:
        return np.zeros_like(Xp, dtype=float)
    squeeze = False
    if Xp.ndim == 1:
        Xp = Xp.reshape(-1, 1)
        squeeze = True
    imp_mean, m, sc = scaler_stats(Xp, scale)
    Xi = np.where(np.isnan(Xp), imp_mean[None, :], Xp)
    out = np.nan_to_num((Xi - m) / sc, nan=0.0)
    return out.reshape(-1) if squeeze else out


def lagged_covariances(trajs: Sequence[np.ndarray], lag: int):
    """Symmetrised mean-free C00, C0t, mean and pair count (deeptime reversible
    estimator, no Bessel): mu=(sum x_t + sum x_{t+lag})/2T,
    C00=(X0c^T X0c + Xtc^T Xtc)/2T, C0t=(X0c^T Xtc + Xtc^T X0c)/2T."""
    lag = int(lag)
    d = trajs[0].shape[1]
    s0 = np.zeros(d)
    st = np.zeros(d)
    T = 0
    for x in trajs:
        x = np.asarray(x, dtype=np.float64)
        if x.shape[0] <= lag:
            continue
        s0 += x[: x.shape[0] - lag].sum(axis=0)
        st += x[lag:].sum(axis=0)
        T += x.shape[0] - lag
    if T == 0:
        raise ValueError("no trajectory longer than the lag time")
    mu = (s0 + st) / (2.0 * T)
    C00 = np.zeros((d, d))
    C0t = np.zeros((d, d))
    for x in trajs:
        x = np.asarray(x, dtype=np.float64)
        if x.shape[0] <= lag:
            continue
        a = x[: x.shape[0] - lag] - mu
        b = x[lag:] - mu
        C00 += a.T @ a + b.T @ b
        cx = a.T @ b
        C0t += cx + cx.T
    C00 /= 2.0 * T
    C0t /= 2.0 * T
    return C00, C0t, mu, T


def _sort_abs_desc(s, V):
    order = np.argsort(-np.abs(s), kind="stable")
    return s[order], V[:, order]


def _canonical_signs(V):
    V = V.copy()
    for j in range(V.shape[1]):
        jj = int(np.argmax(np.abs(V[:, j])))
        sgn = np.sign(V[jj, j])
        if sgn != 0:
            V[:, j] *= sgn
    return V


def spd_inv_split(C0: np.ndarray, epsilon: float = 1e-6):
    """deeptime ``spd_inv_split``: L with L^T C0 L = I on the retained subspace.
    Eigenvalues sorted by magnitude descending, those with |s| <= epsilon
    dropped (epsilon raised to -min(s) if C0 has negative eigenvalues),
    canonical signs (largest-|entry| of each vector positive)."""
    s, V = scipy.linalg.eigh(C0)
    s, V = _sort_abs_desc(s, V)
    evmin = float(np.min(s))
    if evmin < 0:
        epsilon = max(epsilon, -evmin + 1e-16)
    evnorms = np.abs(s)
    n = evnorms.shape[0]
    m = n - int(np.searchsorted(evnorms[::-1], epsilon))
    if m == 0:
        raise ValueError("all eigenvalues below epsilon (zero rank)")
    Vm = _canonical_signs(V[:, :m])
    sm = s[:m]
    return Vm @ np.diag(1.0 / np.sqrt(sm)), sm


def eig_corr(C0: np.ndarray, Ct: np.ndarray, epsilon: float = 1e-6):
    """deeptime ``eig_corr(..., canonical_signs=True)`` for symmetric Ct."""
    L, _ = spd_inv_split(C0, epsilon)
    M = L.T @ Ct @ L
    lam, Rt = scipy.linalg.eigh(0.5 * (M + M.T))
    lam, Rt = _sort_abs_desc(lam, Rt)
    R = _canonical_signs(L @ Rt)
    return lam, R, L.shape[1]


class TicaModel:
    def __init__(self, mean, eigenvalues, eigenvectors, rank, C00, C0t, n_pairs):
        self.mean = mean
        self.eigenvalues = eigenvalues
        self.eigenvectors = eigenvectors  # already kinetic-map scaled
        self.rank = rank
        self.C00 = C00
        self.C0t = C0t
        self.n_pairs = n_pairs


def tica_fit(trajs: Sequence[np.ndarray], lag: int, epsilon: float = 1e-6,
             scaling: str | None = "kinetic_map") -> TicaModel:
    C00, C0t, mu, T = lagged_covariances(trajs, lag)
    lam, R, rank = eig_corr(C00, C0t, epsilon)
    if scaling in ("km", "kinetic_map"):
        R = R * lam[None, :]
    elif scaling is not None:
        raise ValueError(f"unsupported scaling {scaling!r}")
    return TicaModel(mu, lam, R, rank, C00, C0t, T)


def tica_transform(model: TicaModel, X: np.ndarray, dim: int | None) -> np.ndarray:
    m = model.eigenvectors.shape[1] if dim is None else min(int(dim), model.eigenvectors.shape[1])
    return (np.asarray(X, dtype=np.float64) - model.mean) @ model.eigenvectors[:, :m]


def tica_reduce(X: np.ndarray, lag: int = 1, n_components: int = 2, scale: bool = True):
    """reduction.py:77-110."""
    Xp = preprocess(X, scale=scale)
    model = tica_fit([Xp], lag)
    return np.asarray(tica_transform(model, Xp, n_components), dtype=float)


def maybe_apply_tica(features: np.ndarray, lengths: Sequence[int], n_components_hint: int, lag: int):
    """_features.py:181-231: per-trajectory fit, dim clamped to [2,5], the last
    ``lag`` frames of each trajectory dropped after projection."""
    n_components = int(max(2, min(5, n_components_hint)))
    lag_eff = int(max(1, lag or 1))
    Xs, start = [], 0
    for n in lengths:
        Xs.append(np.asarray(features[start:start + n]))
        start += n
    model = tica_fit(Xs, lag_eff)
    Ys = [tica_transform(model, x, n_components) for x in Xs]
    drop = int(max(0, lag))
    if drop > 0:
        Ys = [y[:-drop] if y.shape[0] > drop else np.empty((0, y.shape[1])) for y in Ys]
    return np.vstack(Ys), n_components


def vamp_fit(trajs: Sequence[np.ndarray], lag: int, dim: int | None = None, epsilon: float = 1e-6):
    """deeptime 0.4.5 ``VAMP(lagtime, dim, epsilon)`` (scaling=None) as called by
    src/pmarlo/markov_state_model/reduction.py:113-148.  PARITY UNPINNED against the deeptime binary.
    Published algorithm (Wu & Noe, VAMP; deeptime ``decomposition/_vamp.py::_decomposition``):
    non-reversible covariances with the data mean removed separately for the two time windows, no Bessel
    correction: mu0 = mean x_t, mut = mean x_{t+lag}, C00 = X0c^T X0c / T, C0t = X0c^T Xtc / T,
    Ctt = Xtc^T Xtc / T;  L0 = spd_inv_split(C00, eps), Lt = spd_inv_split(Ctt, eps);
    W = L0^T C0t Lt = A diag(s) B^T (LAPACK gesvd);  left singular functions U = L0 A[:, :m],
    m = min(rank0, rankt, dim);  transform(x) = (x - mu0) U.  The column signs of a singular vector pair
    are LAPACK's choice; here every column of U gets the sign that makes its largest-magnitude entry positive.
    Returns (mu0, U, singular values, rank0, rankt, (C00, C0t, Ctt))."""
    import scipy.linalg

    trajs = [np.asarray(t, dtype=np.float64) for t in trajs]
    X0 = np.concatenate([t[:-lag] for t in trajs if t.shape[0] > lag], axis=0)
    Xt = np.concatenate([t[lag:] for t in trajs if t.shape[0] > lag], axis=0)
    T = X0.shape[0]
    mu0, mut = X0.mean(axis=0), Xt.mean(axis=0)
    X0c, Xtc = X0 - mu0, Xt - mut
    C00, C0t, Ctt = X0c.T @ X0c / T, X0c.T @ Xtc / T, Xtc.T @ Xtc / T
    L0, _ = spd_inv_split(C00, epsilon)
    Lt, _ = spd_inv_split(Ctt, epsilon)
    W = L0.T @ C0t @ Lt
    A, sv, BT = scipy.linalg.svd(W, lapack_driver="gesvd")
    m = min(L0.shape[1], Lt.shape[1]) if dim is None else min(L0.shape[1], Lt.shape[1], int(dim))
    U = L0 @ A[:, :m]
    piv = np.argmax(np.abs(U), axis=0)
    U = U * np.where(U[piv, np.arange(m)] < 0, -1.0, 1.0)[None, :]
    return mu0, U, sv, L0.shape[1], Lt.shape[1], (C00, C0t, Ctt)


def vamp_reduce(X: np.ndarray, lag: int = 1, n_components: int = 2, scale: bool = True, epsilon: float = 1e-6):
    """reduction.py:113-148: _preprocess, VAMP fit on [X_prep], transform(X_prep)."""
    Xp = preprocess(X, scale=scale)
    mu0, U, *_ = vamp_fit([Xp], lag, n_components, epsilon)
    return np.asarray((Xp - mu0) @ U, dtype=float)
