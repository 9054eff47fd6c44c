"""Oracle: lagged transition counting (TEST INFRASTRUCTURE).

Bit-exact integer stage; pinned against pmarlo's own importable functions
(``analysis.discretize._weighted_counts`` discretize.py:609-645,
``analysis.debug_export._build_transition_counts`` debug_export.py:385-409,
``analysis.counting.expected_pairs`` counting.py:10-68) by
tests/golden/make_golden.py.

Two validity modes exist in the reference:
* "endpoint": a pair (t, t+lag) is dropped only when one of its endpoints is
  invalid (label < 0 or >= n_states) -- debug_export.py:399-408,
  ck_its_selector.py:70-83, discretize.py:631-640;
* "split": the trajectory is split at every invalid frame and pairs never
  span a split -- _estimation.py:121-145 feeding deeptime's
  TransitionCountEstimator(count_mode="sliding").
"""

from __future__ import annotations

from typing import Iterable, Sequence

import numpy as np

__all__ = ["count_lagged", "split_at_invalid", "weighted_counts", "expected_pairs", "infer_n_states"]


def infer_n_states(dtrajs: Sequence[np.ndarray]) -> int:
    """_msm_utils.py:190-207."""
    mx = -1
    for dt in dtrajs:
        dt = np.asarray(dt)
        if dt.size:
            m = int(dt.max())
            if m >= 0:
                mx = max(mx, m)
    return mx + 1 if mx >= 0 else 0


def split_at_invalid(dtrajs: Sequence[np.ndarray], n_states: int) -> list[np.ndarray]:
    """_estimation.py:121-145."""
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


def count_lagged(dtrajs: Sequence[np.ndarray], n_states: int, lag: int,
                 mode: str = "endpoint", step: int = 1) -> np.ndarray:
    """C[i,j] = #{t : s_t=i, s_{t+lag}=j}, never across a trajectory boundary."""
    C = np.zeros((n_states, n_states), dtype=np.int64)
    if n_states == 0 or lag <= 0:
        return C
    if mode == "split":
        dtrajs = split_at_invalid(dtrajs, n_states)
    elif mode != "endpoint":
        raise ValueError(mode)
    for d in dtrajs:
        d = np.asarray(d, dtype=np.int64)
        if d.size <= lag:
            continue
        a = d[: d.size - lag: step]
        b = d[lag:: step]
        ok = (a >= 0) & (b >= 0) & (a < n_states) & (b < n_states)
        np.add.at(C, (a[ok], b[ok]), 1)
    return C


def weighted_counts(labels: np.ndarray, *, n_states: int, lag_time: int,
                    weights: np.ndarray | None = None,
                    segments: Iterable[tuple[int, int]] | None = None,
                    stride: int = 1):
    """discretize.py:609-645 (weight of a pair = weight of its starting frame)."""
    counts = np.zeros((n_states, n_states), dtype=np.float64)
    labels = np.asarray(labels)
    if labels.size == 0 or lag_time <= 0:
        return counts, 0
    step = max(1, int(stride))
    segs = [(0, labels.size)] if segments is None else [
        (max(0, int(s)), min(labels.size, int(e))) for s, e in segments]
    total = 0
    for start, stop in segs:
        if stop - start <= lag_time:
            continue
        src = labels[start: stop - lag_time: step]
        dst = labels[start + lag_time: stop: step]
        valid = (src >= 0) & (dst >= 0)
        if not valid.any():
            continue
        w = 1.0 if weights is None else weights[start: stop - lag_time: step][valid]
        np.add.at(counts, (src[valid], dst[valid]), w)
        total += int(np.count_nonzero(valid))
    return counts, total


def expected_pairs(lengths, tau: int, stride=1) -> int:
    """counting.py:10-68."""
    if tau < 0:
        raise ValueError("tau must be non-negative")
    lengths = [int(x) for x in lengths]
    if any(x < 0 for x in lengths):
        raise ValueError("lengths must be non-negative")
    if not lengths or not any(lengths):
        return 0
    if isinstance(stride, (str, bytes)):
        raise TypeError("stride must be an integer or iterable of integers")
    strides = [int(v) for v in stride] if isinstance(stride, Iterable) else [int(stride)]
    if not strides:
        raise ValueError("stride iterable must not be empty")
    if any(v <= 0 for v in strides):
        raise ValueError("stride values must be positive")
    total = 0
    for i, L in enumerate(lengths):
        eff = L - tau
        if L <= 0 or eff <= 0:
            continue
        st = strides[i] if i < len(strides) else strides[-1]
        total += 1 + (eff - 1) // st
    return total
