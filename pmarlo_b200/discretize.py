"""MSM discretisation of continuous CVs behind pmarlo's ``discretize_dataset`` signature.

Mirrors ``pmarlo.analysis.discretize`` (src/pmarlo/analysis/discretize.py):

* ``discretize_dataset(dataset, *, cluster_mode, n_microstates, lag_time, frame_weights,
  min_out_count, random_state, apply_whitening)``                         :901-1120
* ``_KMeansDiscretizer`` (z-score with the SAMPLE standard deviation, k-means fit on the train split,
  nearest-centre prediction for every split, int32 labels)                :406-514
* ``_weighted_counts`` (per segment, pairs (t, t+lag) with both labels >= 0, optional frame weights)
                                                                          :609-645
* zero-row pruning ``_prune_zero_rows_if_needed``                         :819-898
* ``MSMDiscretizationResult``                                             :21-43

What runs where: the k-means fit (Lloyd iterations), every nearest-centre assignment and the lagged
(weighted) counting are libpmb200 kernels (K6, K7); the z-score and the bookkeeping of splits, segments,
schemas and pruning are host logic (torch / numpy on a handful of vectors).  sklearn's k-means++ draws
cannot be reproduced, so the fitted centres differ from the reference's for the same ``random_state``;
everything downstream of the centres is reproduced exactly (``tests/golden/discretize.npz``, keyword
``centers=`` injects the reference's centres).  ``cluster_mode="grid"`` is outside the accelerated path.
"""

from __future__ import annotations

import logging
from dataclasses import dataclass, field
from typing import Any, Dict, Iterable, List, Mapping, MutableMapping, Sequence

import numpy as np
import torch

from . import kernels
from .clustering import kmeans_pp_init, lloyd_device
from .shards import Segments

logger = logging.getLogger("pmarlo")

__all__ = ["MSMDiscretizationResult", "discretize_dataset", "expected_pairs", "PruningFailedError"]


class PruningFailedError(ValueError):
    """Pruning removed every microstate or left empty rows (reference: analysis/errors)."""


@dataclass
class MSMDiscretizationResult:
    """Same fields as the reference container (discretize.py:21-43)."""

    assignments: Dict[str, np.ndarray]
    centers: np.ndarray | None
    counts: np.ndarray
    transition_matrix: np.ndarray
    lag_time: int
    diag_mass: float
    cluster_mode: str
    assignment_masks: Dict[str, np.ndarray] = field(default_factory=dict)
    segment_lengths: Dict[str, List[int]] = field(default_factory=dict)
    segment_strides: Dict[str, List[int]] = field(default_factory=dict)
    counted_pairs: Dict[str, int] = field(default_factory=dict)
    expected_pairs: Dict[str, int] = field(default_factory=dict)
    feature_schema: Dict[str, Any] = field(default_factory=dict)
    fingerprint: Dict[str, Any] = field(default_factory=dict)
    feature_stats: Dict[str, Any] = field(default_factory=dict)
    counts_before_prune: np.ndarray | None = None
    state_counts_before_prune: np.ndarray | None = None
    state_counts: np.ndarray | None = None
    pruned_state_indices: np.ndarray | None = None


def expected_pairs(lengths: Iterable[int], tau: int, stride: int | Iterable[int] = 1) -> int:
    """``1 + (L - tau - 1) // stride`` pairs per segment (analysis/counting.py:10-68)."""
    if tau < 0:
        raise ValueError("tau must be non-negative")
    ls = [int(v) for v in lengths]
    if any(v < 0 for v in ls):
        raise ValueError("lengths must be non-negative")
    if isinstance(stride, (str, bytes)):
        raise TypeError("stride must be an integer or iterable of integers")
    st = [int(v) for v in stride] if isinstance(stride, Iterable) else [int(stride)]
    if not st:
        raise ValueError("stride iterable must not be empty")
    if any(v <= 0 for v in st):
        raise ValueError("stride values must be positive")
    total = 0
    for i, L in enumerate(ls):
        eff = L - tau
        if L > 0 and eff > 0:
            total += 1 + (eff - 1) // (st[i] if i < len(st) else st[-1])
    return total


# ----------------------------------------------------------------------------- dataset plumbing (host)
def _split_array(value: Any) -> np.ndarray | None:
    raw = value.get("X") if isinstance(value, (Mapping, MutableMapping)) else getattr(value, "X", value)
    if raw is None:
        return None
    try:
        arr = np.asarray(raw)
    except Exception:
        return None
    if arr.ndim != 2 or arr.shape[0] == 0 or arr.shape[1] == 0 or arr.dtype.kind not in "fiu":
        return None
    return arr if np.isfinite(arr).all() else None


def _find_splits(dataset) -> Dict[str, Any]:
    """``splits`` mapping, else top-level entries that look like CV matrices, else the dataset itself."""
    found: Dict[str, Any] = {}
    inner = dataset.get("splits") if isinstance(dataset, Mapping) else None
    if isinstance(inner, Mapping):
        found = {str(k): v for k, v in inner.items() if _split_array(v) is not None}
    if not found and isinstance(dataset, Mapping):
        found = {str(k): v for k, v in dataset.items()
                 if not str(k).startswith("__") and _split_array(v) is not None}
    if not found and _split_array(dataset) is not None:
        found = {"all": dataset}
    if not found:
        raise ValueError("No continuous CV splits found in dataset")
    return found


def _schema_of(split: Any, n_features: int) -> Dict[str, Any]:
    names: list[str] = []
    getter = split.get if isinstance(split, (Mapping, MutableMapping)) else (lambda k: getattr(split, k, None))
    for key in ("feature_schema", "cv_names", "feature_names", "columns"):
        raw = getter(key)
        if isinstance(raw, Mapping):
            raw = raw.get("names")
        if raw is None:
            continue
        names = [str(raw)] if isinstance(raw, (str, bytes)) else [str(v) for v in raw if v is not None]
        if names:
            break
    if not names:
        names = [f"feature_{i}" for i in range(n_features)]
    return {"names": names, "n_features": int(n_features)}


def _check_schema(ref: Mapping[str, Any], got: Mapping[str, Any], split_name: str) -> None:
    problems = []
    if int(got["n_features"]) != int(ref["n_features"]):
        problems.append(f"n_features mismatch: expected {ref['n_features']}, got {got['n_features']}")
    elif list(ref["names"]) != list(got["names"]):
        problems.append(f"feature names differ: expected {list(ref['names'])}, got {list(got['names'])}")
    if problems:
        raise ValueError(f"Feature schema mismatch for split '{split_name}': " + "; ".join(problems))


def _segments_of(split: Any, n_frames: int) -> tuple[list[int], list[int]]:
    """(lengths, strides) from ``segments`` / ``__segments__`` / ``segment_lengths`` metadata, truncated
    to the frames present; one segment covering everything when there is no metadata."""
    lengths: list[int] = []
    strides: list[int] = []
    if isinstance(split, Mapping):
        meta = split.get("segments") or split.get("__segments__")
        if isinstance(meta, Iterable):
            for entry in meta:
                if isinstance(entry, Mapping):
                    L = entry.get("length")
                    if L is None and entry.get("start") is not None and entry.get("stop") is not None:
                        L = int(entry["stop"]) - int(entry["start"])
                    st = entry.get("stride") or entry.get("effective_frame_stride")
                else:
                    L, st = entry, None
                try:
                    L = int(L)
                except Exception:
                    continue
                if L > 0:
                    lengths.append(L)
                    strides.append(max(1, int(st)) if st is not None else 1)
        if not lengths and isinstance(split.get("segment_lengths"), Iterable):
            for L in split["segment_lengths"]:
                if int(L) > 0:
                    lengths.append(int(L))
                    strides.append(1)
    out_l, out_s, used = [], [], 0
    for L, st in zip(lengths, strides):
        take = min(L, n_frames - used)
        if take <= 0:
            break
        out_l.append(take)
        out_s.append(st)
        used += take
    if not out_l and n_frames > 0:
        return [n_frames], [1]
    return out_l, out_s


def _column_stats(X: np.ndarray, names: Sequence[str]) -> Dict[str, Any]:
    return {"feature_names": list(names), "n_features": int(X.shape[1]), "n_frames": int(X.shape[0]),
            "means": X.mean(axis=0).tolist(), "stds": X.std(axis=0).tolist(),
            "mins": X.min(axis=0).tolist(), "maxs": X.max(axis=0).tolist()}


# ----------------------------------------------------------------------------- device pieces
def _counts_device(labels: torch.Tensor, lengths: Sequence[int], n_states: int, lag: int,
                   weights: torch.Tensor | None) -> tuple[np.ndarray, int]:
    """``_weighted_counts`` (discretize.py:609-645) on the device: K7 over the segment list; frames past
    the last segment belong to no segment.  Returns (float64 counts, number of counted pairs)."""
    n = int(labels.numel())
    if n == 0 or lag <= 0 or n_states <= 0:
        return np.zeros((max(n_states, 0), max(n_states, 0))), 0
    lens = [int(v) for v in lengths]
    covered = sum(lens)
    segs = Segments.from_lengths(lens + ([n - covered] if covered < n else []))
    off = segs.device(labels.device)
    if covered < n:                       # the tail is not a segment: mark it unassigned for the count
        labels = labels.clone()
        labels[covered:] = -1
    Ci = kernels.count_lagged(labels, off, n_states, int(lag), 1)
    pairs = int(Ci.sum().item())
    if weights is None:
        return Ci.to(torch.float64).cpu().numpy(), pairs
    Cw = kernels.count_lagged_weighted(labels, weights, off, n_states, int(lag), 1)
    return Cw.cpu().numpy(), pairs


def discretize_dataset(dataset, *, cluster_mode: str = "kmeans", n_microstates: int = 150, lag_time: int = 1,
                       frame_weights=None, min_out_count: int = 0, random_state: int | None = None,
                       apply_whitening: bool = True, centers: np.ndarray | None = None,
                       n_init: int = 10, max_iter: int = 300, tolerance: float = 1e-4) -> MSMDiscretizationResult:
    """Drop-in for ``pmarlo.analysis.discretize.discretize_dataset`` (k-means mode).

    Extensions (keyword-only, not in the reference): ``centers`` -- cluster centres in the whitened space,
    skips the fit (used to reproduce a reference run exactly); ``n_init`` / ``max_iter`` / ``tolerance`` of
    the Lloyd fit (sklearn's defaults 10 / 300 / 1e-4)."""
    if lag_time < 1:
        raise ValueError("lag_time must be >= 1")
    if cluster_mode == "grid":
        raise NotImplementedError("cluster_mode='grid' is outside the B200 hot path")
    if cluster_mode != "kmeans":
        raise ValueError("cluster_mode must be 'kmeans' or 'grid'")
    dev = kernels.require_cuda()
    splits = _find_splits(dataset)
    train_key = "train" if "train" in splits else next(iter(splits))
    Xtr = np.array(_split_array(splits[train_key]), dtype=np.float64)
    d = Xtr.shape[1]
    schema = _schema_of(splits[train_key], d)
    if len(schema["names"]) != d:
        raise ValueError(f"Feature schema names length {len(schema['names'])} does not match n_features {d}")
    stats: Dict[str, Dict[str, Any]] = {train_key: _column_stats(Xtr, schema["names"])}
    K = int(n_microstates)

    # whitening: mean and SAMPLE standard deviation of the train split; std <= 1e-10 -> 1 (discretize.py:423-448)
    Xtr_d = torch.from_numpy(Xtr).to(dev)
    if apply_whitening:
        mean_d = Xtr_d.mean(dim=0)
        std_d = Xtr_d.std(dim=0, unbiased=True) if Xtr.shape[0] > 1 else torch.full_like(mean_d, float("nan"))
        std_safe = torch.where(std_d > 1e-10, std_d, torch.ones_like(std_d))
        scaler = {"mean": mean_d.cpu().numpy().tolist(), "std": std_d.cpu().numpy().tolist(), "enabled": True}
    else:
        mean_d = std_safe = None
        scaler = {}

    def whiten(Xd: torch.Tensor) -> torch.Tensor:
        return ((Xd - mean_d) / std_safe).contiguous() if apply_whitening else Xd.contiguous()

    Ztr = whiten(Xtr_d)
    if centers is not None:
        C = torch.from_numpy(np.ascontiguousarray(centers, dtype=np.float64)).to(dev)
        if C.dim() != 2 or int(C.shape[1]) != d:
            raise ValueError("centers must have shape (n_states, n_features)")
    else:
        if Xtr.shape[0] < K:
            raise ValueError(f"n_samples={Xtr.shape[0]} should be >= n_clusters={K}.")
        best = None
        for r in range(max(1, int(n_init))):
            seed = None if random_state is None else int(random_state) + r
            res = lloyd_device(Ztr, kmeans_pp_init(Ztr, K, seed), max_iter=int(max_iter), tolerance=float(tolerance))
            cost = res.cost if res.cost is not None else float("inf")
            if best is None or cost < best[0]:
                best = (cost, res.centers)
        C = best[1]
    K = int(C.shape[0])

    assignments: Dict[str, np.ndarray] = {}
    masks: Dict[str, np.ndarray] = {}
    seg_len: Dict[str, List[int]] = {}
    seg_str: Dict[str, List[int]] = {}
    labels_dev: Dict[str, torch.Tensor] = {}
    max_state = -1
    for name, split in splits.items():
        X = np.array(_split_array(split), dtype=np.float64)
        sch = _schema_of(split, X.shape[1])
        _check_schema(schema, sch, name)
        st = stats.setdefault(name, _column_stats(X, sch["names"]))
        lengths, strides = _segments_of(split, X.shape[0])
        st["segment_lengths"], st["segment_strides"] = list(lengths), list(strides)
        st["expected_pairs"] = expected_pairs(lengths, lag_time, strides if strides else 1)
        Zd = Ztr if name == train_key else whiten(torch.from_numpy(X).to(dev))
        lab = kernels.kmeans_assign(Zd, C)                       # K6: fp64 nearest centre, first minimum wins
        labels_dev[name] = lab
        lab_h = lab.cpu().numpy().astype(np.int32, copy=False)
        if lab_h.size == 0 or not np.any(lab_h >= 0):
            raise ValueError(f"No valid assignments found for split '{name}'")
        assignments[name] = lab_h
        masks[name] = lab_h >= 0
        seg_len[name], seg_str[name] = list(lengths), list(strides)
        max_state = max(max_state, int(lab_h.max()))
    n_states = max_state + 1

    train_lab = labels_dev[train_key]
    n_train = int(train_lab.numel())
    w_d = None
    if frame_weights is not None:
        cand = frame_weights.get(train_key) if isinstance(frame_weights, Mapping) else frame_weights
        if cand is not None:
            w = np.asarray(cand, dtype=np.float64).reshape(-1)
            if w.shape[0] != n_train:
                raise ValueError(f"Frame weights for split '{train_key}' have length {w.shape[0]}, expected {n_train}")
            w_d = torch.from_numpy(w).to(dev)
    train_lengths = seg_len.get(train_key) or [n_train]
    train_strides = seg_str.get(train_key) or []

    counts, counted = _counts_device(train_lab, train_lengths, n_states, lag_time, w_d)
    counts_before = counts.copy()
    counted_before = counted
    exp_pairs = expected_pairs(train_lengths, lag_time, train_strides if train_strides else 1)

    def state_counts(lab_h: np.ndarray, k: int) -> np.ndarray:
        ok = (lab_h >= 0) & (lab_h < k)
        wts = None if w_d is None else w_d.cpu().numpy()[ok]
        return np.bincount(lab_h[ok], weights=wts, minlength=k).astype(np.float64)

    state_counts_before = state_counts(assignments[train_key], n_states)
    row_sums = counts_before.sum(axis=1)
    zero_before = int(np.count_nonzero(row_sums == 0))
    min_out = max(0, int(min_out_count))
    pruned = None
    zero_after = zero_before
    if zero_before > 0:                                           # discretize.py:819-898
        drop = row_sums == 0
        if min_out > 0:
            drop |= row_sums < float(min_out)
        keep = ~drop
        if not np.any(keep):
            raise PruningFailedError(f"Pruning removed all microstates (zero_rows={zero_before}, min_out_count={min_out})")
        pruned = np.where(drop)[0].astype(np.int32)
        mapping = np.full(n_states, -1, dtype=np.int32)
        mapping[keep] = np.arange(int(np.count_nonzero(keep)), dtype=np.int32)
        map_d = torch.from_numpy(mapping).to(dev)
        for name in list(assignments):
            old = labels_dev[name]
            ok = (old >= 0) & (old < n_states)
            new = torch.where(ok, map_d[old.clamp(0, n_states - 1).long()], torch.full_like(old, -1)).to(torch.int32)
            labels_dev[name] = new.contiguous()
            assignments[name] = new.cpu().numpy()
            masks[name] = masks[name] & (assignments[name] >= 0)
        n_states = int(np.count_nonzero(keep))
        counts, counted = _counts_device(labels_dev[train_key], train_lengths, n_states, lag_time, w_d)
        zero_after = int(np.count_nonzero(counts.sum(axis=1) == 0))
        if zero_after > 0:
            raise PruningFailedError(f"Pruning left {zero_after} zero-row microstates (min_out_count={min_out})")
    state_counts_final = state_counts(assignments[train_key], n_states)
    if exp_pairs > 0 and counted == 0:
        raise ValueError(f"No transition pairs counted for split '{train_key}' despite expected {exp_pairs} pairs")

    st = stats[train_key]
    st.update(expected_pairs=int(exp_pairs), counted_pairs_before_prune=int(counted_before), counted_pairs=int(counted),
              zero_rows_before_prune=int(zero_before), zero_rows_after_prune=int(zero_after))
    if pruned is not None and pruned.size:
        st["pruned_state_indices"] = pruned.astype(int).tolist()
        st["prune_min_out_count"] = int(min_out)

    rs = counts.sum(axis=1, keepdims=True)
    T = np.divide(counts, rs, out=np.zeros_like(counts), where=rs > 0)
    diag_mass = float(np.trace(T) / n_states) if n_states else float("nan")
    if np.isfinite(diag_mass) and diag_mass > 0.95:
        logger.warning("MSM diagonal mass high (%.3f)", diag_mass)
    empty = int(np.count_nonzero(counts.sum(axis=1) == 0))
    if counts.shape[0] and empty / counts.shape[0] > 0.3:
        logger.warning("More than 30%% of the microstates are empty (%d/%d)", empty, counts.shape[0])

    fingerprint = {
        "mode": "kmeans", "n_states": int(max(n_states, 0)),
        "seed": None if random_state is None else int(random_state),
        "feature_schema": {"names": list(schema["names"]), "n_features": int(schema["n_features"])},
        "expected_pairs": int(exp_pairs), "counted_pairs": int(counted),
        "segment_lengths": {train_key: train_lengths}, "segment_strides": {train_key: seg_str.get(train_key, [])},
        "zero_rows_before_prune": int(zero_before), "zero_rows_after_prune": int(zero_after),
        "pruned_state_count": int(pruned.size) if pruned is not None else 0, "min_out_count": int(min_out),
        "scaler": scaler,
    }
    return MSMDiscretizationResult(
        assignments=assignments, assignment_masks=masks, segment_lengths=seg_len, segment_strides=seg_str,
        counted_pairs={train_key: int(counted)}, expected_pairs={train_key: int(exp_pairs)},
        centers=C.cpu().numpy(), counts=counts, transition_matrix=T, lag_time=int(lag_time), diag_mass=diag_mass,
        cluster_mode="kmeans", feature_schema=schema, fingerprint=fingerprint, feature_stats=stats,
        counts_before_prune=counts_before, state_counts_before_prune=state_counts_before,
        state_counts=state_counts_final, pruned_state_indices=pruned)
