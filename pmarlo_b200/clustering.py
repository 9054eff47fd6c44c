"""k-means microstate clustering behind pmarlo's signature (K6 on the device).

Mirrors ``cluster_microstates`` / ``ClusteringResult``
(src/pmarlo/markov_state_model/clustering.py:43-90,395-665) and the public
wrapper ``pmarlo.api.clustering.cluster_microstates`` (api/clustering.py:14-69).

deeptime ``KMeans(max_iter=500, tolerance=1e-5, init_strategy="kmeans++")`` Lloyd
loop: assign -> centres = member means (an empty cluster keeps its centre) ->
cost with the NEW centres -> stop when |cost - prev|/cost <= tolerance.  The
cost of iteration i is the inertia measured by the assignment pass of iteration
i+1, so one fused assign+accumulate kernel per iteration suffices; the host
reads one double per iteration for the stopping rule.
Labels are the fp64 direct-difference argmin (first minimum wins), bit-exact.
"""

from __future__ import annotations

import logging
from dataclasses import dataclass
from typing import Any, Literal

import numpy as np
import torch

from . import kernels
from .distributed import Comm
from .timing import NULL_TIMER

logger = logging.getLogger("pmarlo")

__all__ = ["ClusteringResult", "LloydResult", "lloyd_device", "assign_device", "kmeans_pp_init",
           "cluster_microstates", "cluster_microstates_labels"]

_COMMON_KWARGS = frozenset({"max_iter", "metric", "tolerance", "init_strategy", "n_jobs", "initial_centers"})
_ATTRIBUTE_KWARGS = frozenset({"fixed_seed", "progress"})
_MINIBATCH_ONLY_KWARGS = frozenset({"batch_size"})
_SUPPORTED_KWARGS = _COMMON_KWARGS | _ATTRIBUTE_KWARGS | _MINIBATCH_ONLY_KWARGS


@dataclass
class ClusteringResult:
    labels: np.ndarray
    n_states: int
    rationale: str | None = None
    centers: np.ndarray | None = None

    @property
    def output_shape(self) -> tuple[int, ...]:
        return (self.n_states,)


@dataclass
class LloydResult:
    centers: torch.Tensor      # (K,D) float64, device
    n_iter: int
    cost: float | None
    converged: bool


class _Accum:
    """Per-iteration partials: [sums K*D | inertia 1 | counts K (fp64 copy)] fp64 and counts K int64.
    The fp64 copy of the counts (exact below 2^53) lets one all-reduce carry everything
    ("one fused allreduce per iteration")."""

    def __init__(self, K: int, D: int, device):
        self.f = torch.zeros((K * D + 1 + K,), dtype=torch.float64, device=device)
        self.counts = torch.zeros((K,), dtype=torch.int64, device=device)
        self.K, self.D = K, D

    @property
    def sums(self):
        return self.f[: self.K * self.D].view(self.K, self.D)

    @property
    def inertia(self):
        return self.f[self.K * self.D: self.K * self.D + 1]

    def allreduce(self, comm: Comm) -> None:
        if comm.size == 1:
            return
        cf = self.f[self.K * self.D + 1:]
        cf.copy_(self.counts)
        comm.allreduce_sum(self.f)
        self.counts.copy_(cf)

    def zero_(self):
        self.f.zero_()
        self.counts.zero_()


def assign_device(Y: torch.Tensor, centers: torch.Tensor, labels: torch.Tensor | None = None) -> torch.Tensor:
    """Nearest-centre labels (int32) -- ``model.transform(Y)`` (clustering.py:609)."""
    return kernels.kmeans_assign(Y, centers, labels=labels)


_HINT_STRIDE = 16          # first-iteration hints: every 16th frame is assigned cold ...
_HINT_MIN_FRAMES = 1 << 18  # ... when the shard is large enough for the extra launch to pay


def subsampled_hints(Y: torch.Tensor, centers: torch.Tensor, stride: int = _HINT_STRIDE) -> torch.Tensor:
    """Hints for a cold assignment: the exact labels of every ``stride``-th frame, repeated over the frames in
    between.  Frames of a trajectory move slowly, so the centre of a neighbour in time is a good upper bound
    for the screening threshold of K6 (a cold launch scans every score: 4.1 ms against 1.2 ms with hints at
    10 M frames, K = 1000).  Hints only ever skip work; labels do not depend on them."""
    n = int(Y.shape[0])
    sub = kernels.kmeans_assign(Y[::stride], centers)
    hints = sub.repeat_interleave(stride)[:n].contiguous()
    # Frames in no particular order (shuffled data) make the neighbour's centre a useless bound that only costs
    # the gather: when fewer than a tenth of the consecutive subsampled frames share their label the hints
    # are withdrawn (-1 = no hint), decided on the device without a host read-back.
    smooth = (sub[1:] == sub[:-1]).to(torch.float32).mean() >= 0.1 if sub.numel() > 1 else torch.ones((), dtype=torch.bool, device=Y.device)
    return torch.where(smooth, hints, torch.full_like(hints, -1))


def lloyd_device(Y: torch.Tensor, initial_centers: torch.Tensor, max_iter: int = 500,
                 tolerance: float | None = 1e-5, comm: Comm | None = None,
                 labels: torch.Tensor | None = None, timer=NULL_TIMER) -> LloydResult:
    """Lloyd iterations on a frame shard; partial sums are all-reduced every
    iteration.  ``tolerance=None`` runs exactly ``max_iter`` iterations without any
    host read-back (the benchmark's fixed-iteration mode)."""
    comm = comm if comm is not None else Comm()
    centers = initial_centers.to(device=Y.device, dtype=torch.float64).contiguous().clone()
    K, D = int(centers.shape[0]), int(centers.shape[1])
    acc = _Accum(K, D, Y.device)
    if labels is None:
        labels = torch.empty((int(Y.shape[0]),), dtype=torch.int32, device=Y.device)
    prev_cost, cost, it, converged = 0.0, None, 0, False
    # the first iteration starts from the labels of a 1-in-16 subsample; later ones pass the previous labels as hints
    hints = subsampled_hints(Y, centers) if (Y.dtype == torch.float32 and int(Y.shape[0]) >= _HINT_MIN_FRAMES) else None
    while True:
        acc.zero_()
        with timer.stage("kmeans_assign"):
            kernels.kmeans_assign(Y, centers, labels=labels, sums=acc.sums, counts=acc.counts,
                                  inertia=acc.inertia, hints=hints)
        hints = labels
        acc.allreduce(comm)
        old = centers.clone() if tolerance is not None else None
        kernels.kmeans_update(centers, acc.sums, acc.counts)
        it += 1
        if tolerance is not None:
            # deeptime's cluster_loop measures the NEW centres under the assignments that produced them
            # (costAssignFunction(data, newCenters, assignments)): with S_k, n_k of this iteration
            #   cost = sum |y|^2 + sum_k (n_k |c_k|^2 - 2 c_k . S_k),
            # and the kernel's inertia is the same expression at the old centres -- no second assignment.
            nk = acc.counts.to(torch.float64)

            def _centre_term(c):
                return (nk * (c * c).sum(dim=1) - 2.0 * (c * acc.sums).sum(dim=1)).sum()

            cost = float((acc.inertia[0] - _centre_term(old) + _centre_term(centers)).item())
            rel = abs(cost - prev_cost) / cost if cost != 0.0 else 0.0
            prev_cost = cost
            if rel <= tolerance:
                converged = True
        if converged or it >= max_iter:
            break
    return LloydResult(centers, it, cost, converged)


def kmeans_pp_init(Y: torch.Tensor, K: int, seed: int | None) -> torch.Tensor:
    """k-means++ seeding (D^2 sampling).  deeptime draws from its own C++ RNG, so
    seeds cannot be reproduced bit-for-bit; pass ``initial_centers`` for parity."""
    n = int(Y.shape[0])
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed) if seed is not None else int(np.random.randint(0, 2**31 - 1)))
    first = int(torch.randint(0, n, (1,), generator=g).item())
    centers = torch.empty((K, int(Y.shape[1])), dtype=torch.float64, device=Y.device)
    centers[0] = Y[first].to(torch.float64)
    mind2 = None
    for k in range(1, K):
        diff = Y.to(torch.float64) - centers[k - 1]
        d2 = (diff * diff).sum(dim=1)
        mind2 = d2 if mind2 is None else torch.minimum(mind2, d2)
        cdf = torch.cumsum(mind2, dim=0)
        total = float(cdf[-1].item())
        if not total > 0.0:
            idx = int(torch.randint(0, n, (1,), generator=g).item())
        else:
            u = float(torch.rand((1,), generator=g, dtype=torch.float64).item()) * total
            idx = int(torch.searchsorted(cdf, torch.tensor([u], dtype=torch.float64, device=Y.device)).item())
            idx = min(idx, n - 1)
        centers[k] = Y[idx].to(torch.float64)
    return centers


def silhouette_score_device(Y: torch.Tensor, labels: torch.Tensor, K: int) -> float:
    """``sklearn.metrics.silhouette_score(Y, labels)`` (Euclidean): mean of the per-sample coefficients,
    computed by ``pmb_silhouette_samples`` without materialising the n x n distance matrix."""
    return float(kernels.silhouette_samples(Y, labels, K).mean().item())


def auto_select_n_states(Yd: torch.Tensor, random_state: int | None, *, sample_size: int | None = None,
                         override_n_states: int | None = None, max_iter: int = 500, tolerance: float = 1e-5,
                         return_scores: bool = False):
    """``_auto_select_n_states`` (clustering.py:155-250): k-means with k = 4 .. 20 on the (optionally
    sampled) frames, silhouette score of each labelling, the first maximum wins; returns
    ``(n_states, rationale)``.  Seeding is k-means++ from ``random_state`` (deeptime draws from its own C++
    RNG, so the fitted centres -- and hence which k wins on a flat score curve -- are not reproducible across
    the two implementations; the score of a GIVEN labelling is, to 1e-12)."""
    if override_n_states is not None:
        if override_n_states <= 0:
            raise ValueError(f"override_n_states must be a positive integer; received {override_n_states}.")
        return int(override_n_states), f"auto-override={override_n_states}"
    if sample_size is not None:
        if sample_size <= 1:
            raise ValueError("sample_size must be greater than 1 when sampling for silhouette scoring.")
        eff = min(int(sample_size), int(Yd.shape[0]))
        rng = np.random.default_rng(random_state)
        idx = rng.choice(int(Yd.shape[0]), size=eff, replace=False)
        Ys = Yd.index_select(0, torch.from_numpy(idx).to(Yd.device))
        note = f" sample={eff}"
    else:
        Ys, note = Yd, ""
    Ys = Ys.to(torch.float64).contiguous()
    scores: list[tuple[int, float]] = []
    for k in range(4, 21):
        if k > int(Ys.shape[0]):
            scores.append((k, -1.0))
            continue
        c0 = kmeans_pp_init(Ys, k, random_state)
        res = lloyd_device(Ys, c0, max_iter=max_iter, tolerance=tolerance)
        lab = assign_device(Ys, res.centers)
        if int(torch.unique(lab).numel()) <= 1:
            score = -1.0
        else:
            score = silhouette_score_device(Ys, lab, k)
        scores.append((k, score))
    chosen, best = max(scores, key=lambda x: x[1])
    rationale = f"silhouette={best:.3f}{note}"
    logger.info("Auto-selected %d states with silhouette score %.3f", chosen, best)
    if return_scores:
        return chosen, rationale, scores
    return chosen, rationale


def _validate_kwargs(method: str, kwargs: dict) -> None:
    unsupported = set(kwargs) - set(_SUPPORTED_KWARGS)
    if unsupported:
        raise TypeError(
            f"Unsupported clustering parameters for deeptime backend: {sorted(unsupported)}")
    if method == "kmeans" and any(k in kwargs for k in _MINIBATCH_ONLY_KWARGS):
        raise TypeError("'batch_size' is only supported when method='minibatchkmeans'.")


def _remap_and_inertia(Y: torch.Tensor, raw: torch.Tensor, K: int):
    """``_remap_labels_and_compute_inertia`` (clustering.py:364-392): dense relabel by
    sorted unique label, centres := member means, inertia about those means --
    one accumulate pass + one assign-free inertia pass on the device."""
    counts = torch.bincount(raw.to(torch.int64), minlength=K)
    used = counts > 0
    n_unique = int(used.sum().item())
    if n_unique == 0:
        raise ValueError("Clustering produced zero unique microstates; verify input coverage "
                         "and CV preprocessing.")
    dense = (torch.cumsum(used.to(torch.int64), dim=0) - 1)
    remapped = dense[raw.to(torch.int64)]
    return remapped, n_unique


def cluster_microstates(Y: np.ndarray, method: Literal["auto", "minibatchkmeans", "kmeans"] = "auto",
                        n_states: int | Literal["auto"] = "auto", random_state: int | None = 42,
                        minibatch_threshold: int = 5_000_000, *,
                        silhouette_sample_size: int | None = None,
                        auto_n_states_override: int | None = None, **kwargs: Any) -> ClusteringResult:
    """Drop-in for ``pmarlo.markov_state_model.clustering.cluster_microstates``.

    Differences that are deliberate and documented in INTEGRATION.md: every
    method runs full-batch Lloyd on the device (``"auto"``/``"minibatchkmeans"`` do
    not subsample: B200 does a full pass over 10 M frames in milliseconds).
    ``n_states="auto"`` runs the reference's silhouette scan (k = 4 .. 20) with the
    O(n^2) scoring on the device (``pmb_silhouette_samples``); pass
    ``silhouette_sample_size`` for large inputs, like the reference.
    """
    Y = np.asarray(Y) if not isinstance(Y, torch.Tensor) else Y
    if Y.shape[0] == 0:
        logger.info("Empty dataset provided, returning empty clustering result")
        return ClusteringResult(labels=np.empty((0,), dtype=int), n_states=0)
    if Y.ndim != 2:
        raise ValueError(f"Input must be 2D array, got shape {Y.shape}")
    if Y.shape[1] == 0:
        raise ValueError("Input array must have at least one feature")
    kwargs = dict(kwargs)
    raw_n_init = kwargs.pop("n_init", None)
    if raw_n_init is None:
        n_init = 1
    else:
        try:
            n_init = int(raw_n_init)
        except (TypeError, ValueError) as exc:
            raise TypeError("n_init must be provided as an integer when clustering with deeptime") from exc
        if n_init <= 0:
            raise ValueError("n_init must be a positive integer when clustering microstates")
    if n_init > 1 and "fixed_seed" in kwargs:
        raise ValueError("n_init cannot be combined with fixed_seed; provide only one mechanism "
                         "for controlling clustering initialisations.")
    if method not in ("auto", "kmeans", "minibatchkmeans"):
        raise ValueError(f"Unsupported clustering method: {method}")
    _validate_kwargs(method, kwargs)
    rationale = None
    auto = isinstance(n_states, str) and n_states == "auto"
    if auto and auto_n_states_override is not None:
        if auto_n_states_override <= 0:
            raise ValueError("override_n_states must be a positive integer; "
                             f"received {auto_n_states_override}.")
        n_states = int(auto_n_states_override)
        rationale = f"auto-override={n_states}"
        auto = False
    elif not auto:
        n_states = int(n_states)
    if not auto and n_states <= 0:
        raise ValueError(f"Number of microstates must be a positive integer; received {n_states}.")
    chosen = method
    if method == "auto":
        chosen = "minibatchkmeans" if int(Y.shape[0] * Y.shape[1]) > minibatch_threshold else "kmeans"
    if "batch_size" in kwargs and chosen != "minibatchkmeans":
        raise ValueError("batch_size was provided but the selected clustering method is "
                         f"'{chosen}'. Specify method='minibatchkmeans' to use mini-batch parameters.")
    metric = kwargs.get("metric", "euclidean")
    if metric not in (None, "euclidean"):
        raise ValueError(f"unsupported metric {metric!r}")
    max_iter = int(kwargs.get("max_iter", 500))
    tolerance = float(kwargs.get("tolerance", 1e-5))
    init = kwargs.get("initial_centers")

    dev = kernels.require_cuda()
    if isinstance(Y, torch.Tensor):
        Yd = Y.to(dev)
        if Yd.dtype not in (torch.float32, torch.float64):
            Yd = Yd.to(torch.float64)
    else:
        Yd = torch.from_numpy(np.ascontiguousarray(Y, dtype=np.float64)).to(dev)
    Yd = Yd.contiguous()
    if auto:
        # silhouette scan over k = 4 .. 20 (clustering.py:155-250, 555-570)
        n_states, rationale = auto_select_n_states(Yd, random_state, sample_size=silhouette_sample_size,
                                                   max_iter=max_iter, tolerance=tolerance)
    logger.info("Starting clustering with %s algorithm: %d states, %d samples, %d features",
                chosen, n_states, Yd.shape[0], Yd.shape[1])

    if n_init == 1:
        seeds = [kwargs.get("fixed_seed", random_state)]
    else:
        rng = np.random.default_rng(random_state)
        seeds = [None if random_state is None else int(random_state)]
        seen = {s for s in seeds if isinstance(s, int)}
        while len(seeds) < n_init:
            cand = int(rng.integers(0, np.iinfo(np.int32).max))
            if cand in seen:
                continue
            seeds.append(cand)
            seen.add(cand)

    best = None
    for idx, seed in enumerate(seeds):
        if init is not None:
            c0 = torch.as_tensor(np.asarray(init, dtype=np.float64)).to(dev)
            if c0.shape != (n_states, Yd.shape[1]):
                raise ValueError("initial_centers must have shape (n_states, n_features)")
        else:
            c0 = kmeans_pp_init(Yd, n_states, None if isinstance(seed, bool) else seed)
        res = lloyd_device(Yd, c0, max_iter=max_iter, tolerance=tolerance)
        raw = assign_device(Yd, res.centers)
        remapped, n_unique = _remap_and_inertia(Yd, raw, n_states)
        # member means + inertia about them: one accumulate pass, one inertia pass
        acc = _Accum(n_unique, int(Yd.shape[1]), dev)
        cz = torch.zeros((n_unique, int(Yd.shape[1])), dtype=torch.float64, device=dev)
        cz.index_add_(0, remapped, Yd.to(torch.float64))
        cnt = torch.bincount(remapped, minlength=n_unique).to(torch.float64)
        centers = cz / cnt[:, None]
        diffs = Yd.to(torch.float64) - centers[remapped]
        inertia = float((diffs * diffs).sum().item())
        del acc
        run = dict(labels=remapped, centers=centers, unique=n_unique, inertia=inertia, seed=seed, iteration=idx)
        if best is None or inertia < best["inertia"]:
            best = run
    assert best is not None
    if n_init > 1:
        logger.info("Selected best clustering from %d initialisations (iteration=%d, inertia=%.6f)",
                    n_init, int(best["iteration"]), float(best["inertia"]))
    unique = int(best["unique"])
    if unique != n_states:
        logger.warning("Clustering produced %d unique microstates, expected %d. Proceeding with the "
                       "observed value; inspect CV spread or adjust the requested microstate count.",
                       unique, n_states)
        n_states = unique
    return ClusteringResult(labels=best["labels"].cpu().numpy().astype(int), n_states=n_states,
                            rationale=rationale, centers=best["centers"].cpu().numpy())


def cluster_microstates_labels(Y, method="auto", n_states="auto", random_state=42,
                               minibatch_threshold=5_000_000, **kwargs) -> np.ndarray:
    """``pmarlo.api.clustering.cluster_microstates``: same call, returns labels only."""
    return cluster_microstates(Y, method=method, n_states=n_states, random_state=random_state,
                               minibatch_threshold=minibatch_threshold, **kwargs).labels
