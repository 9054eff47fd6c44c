"""Preprocessing + TICA behind pmarlo's signatures (K2-K5 on the device).

Mirrors
* ``_preprocess`` / ``tica_reduce(X, lag, n_components, scale)``
      src/pmarlo/markov_state_model/reduction.py:13-40,77-110
* ``reduce_features(X, method, lag, n_components)``   src/pmarlo/api/features.py:468-481
* ``FeaturesMixin._maybe_apply_tica(n_components_hint, lag)``
      src/pmarlo/markov_state_model/_features.py:181-231
* ``train_cv_model(features, lag_time, n_components, method="tica")``  src/pmarlo/cv/__init__.py:42-50

deeptime ``TICA(lagtime, dim)`` semantics (defaults epsilon=1e-6,
scaling="kinetic_map", reversible symmetrised covariances, no Bessel): one pass
of column moments, two tensor/SIMT Gram passes over the conditioned data, an
fp64 assembly that undoes the conditioning exactly, a Jacobi eigen-solve on the
device, and a bandwidth-bound projection.  The scaler (mean-impute + z-score)
is folded into the projection operands, so the features are read three times
and never rewritten.
"""

from __future__ import annotations

import logging
from dataclasses import dataclass
from typing import Sequence

import numpy as np
import torch

from . import kernels
from .distributed import Comm
from .shards import Segments, concat_to_device
from .timing import NULL_TIMER

logger = logging.getLogger("pmarlo")

__all__ = ["TicaModel", "TICA", "TicaAccumulator", "tica_reduce", "reduce_features", "maybe_apply_tica", "train_cv_model",
           "preprocess"]


@dataclass
class TicaModel:
    """Fitted model; every tensor lives on the device."""

    lag: int
    dim: int
    n_frames: int
    n_pairs: int
    moments: torch.Tensor      # (6,d)
    stats: torch.Tensor        # (3,d): mean, scale, sample std
    C00: torch.Tensor          # (d,d) in preprocessed units
    C0t: torch.Tensor
    mu: torch.Tensor           # pair mean (preprocessed units)
    eigenvalues: torch.Tensor  # (d,), sorted by magnitude; entries >= rank are 0
    eigenvectors: torch.Tensor  # (d,d) unscaled generalized eigenvectors R (columns)
    rank_dev: torch.Tensor     # int32[1]
    a: torch.Tensor            # projection offset (raw units)
    nanfill: torch.Tensor      # imputation value per column
    W: torch.Tensor            # (d, dim) projection matrix (raw units)

    @property
    def rank(self) -> int:
        return int(self.rank_dev[0].item())

    @property
    def output_dim(self) -> int:
        return min(self.dim, self.rank)


class TICA:
    """``preprocess``: "standard" (mean-impute + z-score, reduction.py:13-40 with
    scale=True), "center" (scale=False) or None (raw features, _features.py:181-231)."""

    def __init__(self, lagtime: int, dim: int | None = None, epsilon: float = 1e-6,
                 scaling: str | None = "kinetic_map", preprocess: str | None = None,
                 comm: Comm | None = None, gram_impl: int = 0):
        if int(lagtime) < 1:
            raise ValueError("lagtime must be >= 1")
        if scaling not in (None, "kinetic_map", "km"):
            raise ValueError(f"unsupported scaling {scaling!r}")
        if preprocess not in (None, "standard", "center"):
            raise ValueError(f"unsupported preprocess {preprocess!r}")
        self.lagtime, self.dim, self.epsilon = int(lagtime), dim, float(epsilon)
        self.scaling, self.preprocess = scaling, preprocess
        self.comm = comm if comm is not None else Comm()
        self.gram_impl = int(gram_impl)

    def fit_device(self, X: torch.Tensor, segs: Segments, seg_offsets_dev: torch.Tensor | None = None,
                   timer=NULL_TIMER) -> TicaModel:
        n_local, d = int(X.shape[0]), int(X.shape[1])
        comm, lag = self.comm, self.lagtime
        dev = X.device
        off = seg_offsets_dev if seg_offsets_dev is not None else segs.device(dev)
        n = comm.sum_int(n_local, dev)
        n_pairs = comm.sum_int(segs.n_pairs(lag), dev)
        if n_pairs <= 0:
            raise ValueError("no trajectory longer than the lag time")
        mask = kernels.pair_mask(off, n_local, lag)
        shift = None
        if comm.size > 1:
            # rank 0's first frame; an empty shard elsewhere still takes part in the broadcast, and an empty
            # rank 0 is reported on every rank
            n0 = comm.sum_int(n_local if comm.rank == 0 else 0, dev)
            if n0 <= 0:
                raise ValueError("rank 0 holds no frames")
            shift = (torch.nan_to_num(X[0].to(torch.float64), nan=0.0).contiguous() if comm.rank == 0
                     else torch.empty((int(X.shape[1]),), dtype=torch.float64, device=dev))
            comm.broadcast(shift, src=0)
        with timer.stage("col_moments"):
            moments = kernels.col_moments(X, mask, shift)
        if comm.size > 1:
            comm.allreduce_sum(moments)
            moments[1].copy_(shift)
        semantic = 0 if self.preprocess is None else 1
        with_std = 1 if self.preprocess == "standard" else 0
        stats, cond = kernels.scaler_from_moments(moments, n, semantic, with_std)
        G = torch.empty((2, d, d), dtype=torch.float64, device=dev)
        with timer.stage("gram"):
            kernels.gram(X, mask, lag, 0, cond, self.gram_impl, out=G[0])
        with timer.stage("gram"):
            kernels.gram(X, mask, lag, 1, cond, self.gram_impl, out=G[1])
        comm.allreduce_sum(G)
        C00, C0t, mu = kernels.tica_covariances(G[0], G[1], moments, stats, cond, n, n_pairs, semantic)
        with timer.stage("tica_solve"):
            evals, evecs, rank = kernels.tica_solve(C00, C0t, self.epsilon)
        dim = d if self.dim is None else int(self.dim)
        dim = max(1, min(dim, d))
        a, nanfill, W = kernels.tica_finalize(evals, evecs, moments, stats, mu, dim,
                                              self.scaling in ("kinetic_map", "km"))
        return TicaModel(lag, dim, n, n_pairs, moments, stats, C00, C0t, mu, evals, evecs, rank, a,
                         nanfill, W)

    def accumulator(self, d: int, device) -> "TicaAccumulator":
        """Streaming fit: feed whole-trajectory chunks with ``add`` while later chunks are still being
        copied or featurized, then ``finish``."""
        return TicaAccumulator(self, d, device)

    @staticmethod
    def transform_device(model: TicaModel, X: torch.Tensor, out_f64: bool = False,
                         out: torch.Tensor | None = None) -> torch.Tensor:
        return kernels.project(X, model.a, model.nanfill, model.W, out_f64=out_f64, out=out)

    # numpy convenience (list of per-trajectory arrays, like deeptime's fit(list))
    def fit(self, trajs: Sequence[np.ndarray]) -> TicaModel:
        dev = kernels.require_cuda()
        X, segs = concat_to_device(trajs, dev)
        return self.fit_device(X, segs)

    def fit_transform(self, trajs: Sequence[np.ndarray]) -> tuple[list[np.ndarray], TicaModel]:
        dev = kernels.require_cuda()
        X, segs = concat_to_device(trajs, dev)
        model = self.fit_device(X, segs)
        Y = self.transform_device(model, X, out_f64=True)[:, : model.output_dim]
        return segs.split(Y.cpu().numpy()), model


class TicaAccumulator:
    """Chunk-wise TICA fit (same result as :meth:`TICA.fit_device` up to fp64 summation order).

    The column moments and the two Gram matrices are sums over frames, so they can be accumulated
    chunk by chunk (chunks = whole trajectories: lagged pairs never cross a chunk) while the next chunk
    is still in flight from the host.  The fp32 conditioning (shift, scale) of the Gram kernel only
    serves numerics and is undone exactly by ``pmb_tica_covariances``, so it is taken from the FIRST
    chunk (rank 0's, broadcast) instead of the global statistics.  Two cases fall back to a second pass
    over the resident features with the global conditioning: NaNs (the kernel imputes at the conditioning
    shift) and a first chunk whose statistics are far from the global ones.
    """

    def __init__(self, est: TICA, d: int, device):
        self.est, self.d, self.device = est, int(d), device
        self.moments = None
        self.shift = None
        self.cond = None
        self.G = torch.zeros((2, self.d, self.d), dtype=torch.float64, device=device)
        self._tmp = torch.empty((self.d, self.d), dtype=torch.float64, device=device)
        self.n_local = 0
        self.n_pairs_local = 0
        self.chunks: list[tuple[torch.Tensor, torch.Tensor]] = []   # (X view, mask) for a possible second pass

    def _start(self, X: torch.Tensor) -> None:
        comm = self.est.comm
        if comm.size > 1:
            n0 = comm.sum_int(int(X.shape[0]) if comm.rank == 0 else 0, X.device)
            if n0 <= 0:
                raise ValueError("rank 0's first chunk holds no frames")
        if comm.size > 1 and comm.rank != 0:
            self.shift = torch.empty((int(X.shape[1]),), dtype=torch.float64, device=X.device)
        else:
            self.shift = torch.nan_to_num(X[0].to(torch.float64), nan=0.0).contiguous()
        if comm.size > 1:
            comm.broadcast(self.shift, src=0)

    def add(self, X: torch.Tensor, segs: Segments, timer=NULL_TIMER) -> None:
        est, comm, lag = self.est, self.est.comm, self.est.lagtime
        if self.shift is None:
            self._start(X)
        n = int(X.shape[0])
        if n == 0:
            # an empty shard still takes part in the first chunk's broadcasts (shift above, conditioning here)
            if self.moments is None:
                self.moments = torch.zeros((6, self.d), dtype=torch.float64, device=self.device)
                self.moments[1].copy_(self.shift)
                self.cond = torch.empty((2, self.d), dtype=torch.float32, device=self.device)
                if comm.size > 1:
                    comm.broadcast(self.cond, src=0)
            return
        mask = kernels.pair_mask(segs.device(self.device), n, lag)
        with timer.stage("col_moments"):
            mom = kernels.col_moments(X, mask, self.shift)
        if self.moments is None:
            self.moments = mom
            semantic = 0 if est.preprocess is None else 1
            _, self.cond = kernels.scaler_from_moments(mom, n, semantic, 1 if est.preprocess == "standard" else 0)
            if comm.size > 1:
                comm.broadcast(self.cond, src=0)
        else:
            self.moments[0] += mom[0]
            self.moments[2:] += mom[2:]
        with timer.stage("gram"):
            kernels.gram(X, mask, lag, 0, self.cond, est.gram_impl, out=self._tmp)
            self.G[0] += self._tmp
            kernels.gram(X, mask, lag, 1, self.cond, est.gram_impl, out=self._tmp)
            self.G[1] += self._tmp
        self.n_local += n
        self.n_pairs_local += segs.n_pairs(lag)
        self.chunks.append((X, mask))

    def finish(self, timer=NULL_TIMER) -> TicaModel:
        est, comm, lag, d, dev = self.est, self.est.comm, self.est.lagtime, self.d, self.device
        if self.moments is None:
            raise ValueError("no frames given")
        n = comm.sum_int(self.n_local, dev)
        n_pairs = comm.sum_int(self.n_pairs_local, dev)
        if n_pairs <= 0:
            raise ValueError("no trajectory longer than the lag time")
        moments, G = self.moments, self.G
        if comm.size > 1:
            comm.allreduce_sum(moments)
            moments[1].copy_(self.shift)
            comm.allreduce_sum(G)
        semantic = 0 if est.preprocess is None else 1
        with_std = 1 if est.preprocess == "standard" else 0
        stats, cond_global = kernels.scaler_from_moments(moments, n, semantic, with_std)
        # is the first chunk's conditioning good enough?  |z| must stay O(1): scales within 4x, means within
        # 4 sigma, and no NaNs (one device->host read of three booleans)
        ratio = self.cond[1].to(torch.float64) / cond_global[1].to(torch.float64)
        drift = (self.cond[0].to(torch.float64) - cond_global[0].to(torch.float64)).abs() * cond_global[1].to(torch.float64)
        ok = torch.stack([(ratio.max() <= 4.0) & (ratio.min() >= 0.25), drift.max() <= 4.0,
                          (moments[0] == float(n)).all()]).all()
        cond = self.cond
        self.fell_back = not bool(ok.item())
        if self.fell_back:
            cond = cond_global
            G.zero_()
            for X, mask in self.chunks:
                with timer.stage("gram"):
                    kernels.gram(X, mask, lag, 0, cond, est.gram_impl, out=self._tmp)
                    G[0] += self._tmp
                    kernels.gram(X, mask, lag, 1, cond, est.gram_impl, out=self._tmp)
                    G[1] += self._tmp
            if comm.size > 1:
                comm.allreduce_sum(G)
        self.chunks = []
        C00, C0t, mu = kernels.tica_covariances(G[0], G[1], moments, stats, cond, n, n_pairs, semantic)
        with timer.stage("tica_solve"):
            evals, evecs, rank = kernels.tica_solve(C00, C0t, est.epsilon)
        dim = d if est.dim is None else int(est.dim)
        dim = max(1, min(dim, d))
        a, nanfill, W = kernels.tica_finalize(evals, evecs, moments, stats, mu, dim,
                                              est.scaling in ("kinetic_map", "km"))
        return TicaModel(lag, dim, n, n_pairs, moments, stats, C00, C0t, mu, evals, evecs, rank, a,
                         nanfill, W)


def _as_2d(X) -> tuple[np.ndarray, bool]:
    Xp = np.asarray(X, dtype=float)
    if Xp.ndim == 1:
        return Xp.reshape(-1, 1), True
    return Xp, False


def preprocess(X: np.ndarray, scale: bool = True) -> np.ndarray:
    """``_preprocess`` (reduction.py:13-40): mean-impute NaNs, centre, optional z-score."""
    Xp, squeeze = _as_2d(X)
    if Xp.size == 0:
        return np.zeros_like(np.asarray(X, dtype=float))
    dev = kernels.require_cuda()
    Xd = torch.from_numpy(np.ascontiguousarray(Xp, dtype=np.float32)).to(dev)
    moments = kernels.col_moments(Xd)
    stats, _ = kernels.scaler_from_moments(moments, Xd.shape[0], 1, 1 if scale else 0)
    d = Xd.shape[1]
    eye = torch.eye(d, dtype=torch.float64, device=dev)
    nv, sh, s1 = moments[0], moments[1], moments[2]
    nanfill = torch.where(nv > 0, sh + s1 / torch.clamp(nv, min=1.0), torch.zeros_like(sh))
    W = (eye / stats[1][:, None]).contiguous()
    out = kernels.project(Xd, stats[0].contiguous(), nanfill.contiguous(), W, out_f64=True)
    res = out.cpu().numpy()
    return res.reshape(-1) if squeeze else res


def tica_reduce(X: np.ndarray, lag: int = 1, n_components: int = 2, scale: bool = True) -> np.ndarray:
    """Drop-in for ``pmarlo.markov_state_model.reduction.tica_reduce``."""
    Xp, _ = _as_2d(X)
    est = TICA(lagtime=lag, dim=n_components, preprocess="standard" if scale else "center")
    dev = kernels.require_cuda()
    Xd = torch.from_numpy(np.ascontiguousarray(Xp, dtype=np.float32)).to(dev)
    segs = Segments.from_lengths([Xd.shape[0]])
    model = est.fit_device(Xd, segs)
    Y = est.transform_device(model, Xd, out_f64=True)[:, : model.output_dim]
    return np.asarray(Y.cpu().numpy(), dtype=float)


def _spd_inv_split_device(C: torch.Tensor, epsilon: float) -> torch.Tensor:
    """deeptime ``spd_inv_split`` on a d x d fp64 device matrix: L with L^T C L = I on the retained subspace
    (eigenvalues sorted by magnitude, |s| <= epsilon dropped, epsilon raised to -min(s) for an indefinite C,
    largest-magnitude entry of every eigenvector positive)."""
    s, V = torch.linalg.eigh(C)                       # d x d, cold path: library call
    order = torch.argsort(s.abs(), descending=True, stable=True)
    s, V = s[order], V[:, order]
    eps = float(epsilon)
    smin = float(s.min().item())
    if smin < 0:
        eps = max(eps, -smin + 1e-16)
    m = int((s.abs() > eps).sum().item())
    if m == 0:
        raise ValueError("all eigenvalues below epsilon (zero rank)")
    V, s = V[:, :m], s[:m]
    piv = V.abs().argmax(dim=0)
    sign = torch.where(V[piv, torch.arange(m, device=V.device)] < 0, -1.0, 1.0).to(V.dtype)
    return (V * sign[None, :]) / torch.sqrt(s)[None, :]


def vamp_reduce(X: np.ndarray, lag: int = 1, n_components: int = 2, scale: bool = True,
                epsilon: float = 1e-6) -> np.ndarray:
    """Drop-in for ``pmarlo.markov_state_model.reduction.vamp_reduce`` (reduction.py:113-148): ``_preprocess``
    then deeptime ``VAMP(lagtime, dim, epsilon)`` fit on the one trajectory and ``transform`` of it (the call
    src/pmarlo/api/conformations.py:195 makes).

    The N d^2 work -- C00, C0t (not symmetrised), Ctt about the separate window means -- is ONE pass of the K3
    Gram kernel over the stacked pairs Z = [x_t | x_{t+lag}] (N - lag, 2 d): G = sum z z^T with the kernel's
    conditioning z = (Z - block mean) / std holds all three blocks; the fp32 rounding of the conditioning
    constants is undone exactly in fp64.  The d x d whitening and the SVD of L0^T C0t Lt are library calls on
    the device (torch.linalg, cold path); the projection is K5.  Singular-vector signs are LAPACK's choice in
    the reference; here the largest-magnitude entry of every left singular function is positive."""
    Xp, _ = _as_2d(X)
    n, d = int(Xp.shape[0]), int(Xp.shape[1])
    lag = int(lag)
    if lag < 1:
        raise ValueError("lagtime must be >= 1")
    if n <= lag:
        raise ValueError("no trajectory longer than the lag time")
    dev = kernels.require_cuda()
    Xd = torch.from_numpy(np.ascontiguousarray(Xp, dtype=np.float32)).to(dev)
    moments = kernels.col_moments(Xd)
    stats, _ = kernels.scaler_from_moments(moments, n, 1, 1 if scale else 0)
    mean, std = stats[0].contiguous(), stats[1].contiguous()
    if bool((moments[0] < n).any().item()):           # mean-impute NaNs exactly like SimpleImputer
        Xd = torch.where(torch.isnan(Xd), mean.to(torch.float32)[None, :], Xd).contiguous()
    T = n - lag
    Z = torch.cat([Xd[:T], Xd[lag:]], dim=1).contiguous()            # the lagged pairs, side by side
    mz = kernels.col_moments(Z)
    zmean = mz[1] + mz[2] / torch.clamp(mz[0], min=1.0)              # window means mu_0 | mu_t (raw units)
    inv_std = (1.0 / std).repeat(2)
    cond = torch.stack([zmean.to(torch.float32), inv_std.to(torch.float32)]).contiguous()
    ones = torch.ones((T,), dtype=torch.uint8, device=dev)            # weight 1 for every pair
    G = kernels.gram(Z, ones, 0, 0, cond)
    # undo the fp32 rounding of the conditioning constants: z_exact = (z32 - delta) * r
    r = inv_std / cond[1].to(torch.float64)
    delta = (zmean - cond[0].to(torch.float64)) * cond[1].to(torch.float64)
    C = (G - float(T) * torch.outer(delta, delta)) * torch.outer(r, r) / float(T)
    C00, C0t, Ctt = C[:d, :d], C[:d, d:], C[d:, d:]
    C00, Ctt = 0.5 * (C00 + C00.T), 0.5 * (Ctt + Ctt.T)
    L0 = _spd_inv_split_device(C00, epsilon)
    Lt = _spd_inv_split_device(Ctt, epsilon)
    W = L0.T @ C0t @ Lt
    A, sv, _ = torch.linalg.svd(W, full_matrices=False)
    m = min(int(L0.shape[1]), int(Lt.shape[1]), int(n_components))
    U = L0 @ A[:, :m]
    piv = U.abs().argmax(dim=0)
    U = U * torch.where(U[piv, torch.arange(m, device=dev)] < 0, -1.0, 1.0).to(U.dtype)[None, :]
    Wp = (U / std[:, None]).contiguous()
    Y = kernels.project(Xd, zmean[:d].contiguous(), mean, Wp, out_f64=True)
    return np.asarray(Y.cpu().numpy(), dtype=float)


def reduce_features(X: np.ndarray, method: str = "tica", lag: int = 10, n_components: int = 2) -> np.ndarray:
    """Drop-in for ``pmarlo.api.features.reduce_features`` (api/features.py:468-481): TICA and VAMP on the
    device; PCA (sklearn) is outside the north-star path and is not silently emulated."""
    if method == "tica":
        return tica_reduce(X, lag=lag, n_components=n_components)
    if method == "vamp":
        return vamp_reduce(X, lag=lag, n_components=n_components)
    if method == "pca":
        raise NotImplementedError("reduce_features(method='pca') is outside the B200 hot path")
    raise ValueError(f"Unknown reduction method: {method}")


def maybe_apply_tica(features: np.ndarray, lengths: Sequence[int], n_components_hint: int | None, lag: int):
    """``FeaturesMixin._maybe_apply_tica``: fit on the list of trajectories, dim
    clamped to [2,5], the last ``lag`` frames of every trajectory dropped.
    Returns ``(Y, n_components, kept_lengths)``."""
    if features is None or n_components_hint is None:
        return features, None, list(lengths)
    n_components = int(max(2, min(5, n_components_hint)))
    lag_eff = int(max(1, lag or 1))
    dev = kernels.require_cuda()
    Xd = torch.from_numpy(np.ascontiguousarray(features, dtype=np.float32)).to(dev)
    segs = Segments.from_lengths(lengths)
    if segs.n_frames != Xd.shape[0]:
        raise ValueError("lengths do not add up to the number of feature rows")
    est = TICA(lagtime=lag_eff, dim=n_components, preprocess=None)
    model = est.fit_device(Xd, segs)
    Y = est.transform_device(model, Xd, out_f64=True)[:, : model.output_dim].cpu().numpy()
    keep, new_segs = segs.drop_tail(int(max(0, lag)))
    logger.info("Total projected arrays stored: %d", len(lengths))
    return Y[keep], n_components, [int(v) for v in new_segs.lengths]


def train_cv_model(features: Sequence[np.ndarray], lag_time: int, n_components: int = 2, method: str = "tica"):
    """TICA branch of ``pmarlo.cv.train_cv_model`` (cv/__init__.py:42-50)."""
    if method != "tica":
        raise NotImplementedError("only method='tica' is on the B200 hot path")
    return TICA(lagtime=lag_time, dim=n_components).fit(list(features))
