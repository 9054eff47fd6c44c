"""Shard semantics: a dataset is a list of independent trajectories.

pmarlo feeds lists of per-trajectory arrays into deeptime (``TICA.fit(list)``,
``TransitionCountEstimator.fit(dtrajs)``; src/pmarlo/markov_state_model/
_features.py:182-202, _msm_utils.py:238-246): lagged pairs never cross a
trajectory boundary.  On the device the shards are stored back to back in one
buffer plus an offsets vector; whole trajectories are dealt to ranks
(SURVEY.md section 8e), so no halo exchange is needed.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import numpy as np
import torch

__all__ = ["Segments", "partition_trajectories", "concat_to_device"]


@dataclass
class Segments:
    """Offsets of back-to-back trajectories: ``offsets[s]:offsets[s+1]`` is shard s."""

    offsets: np.ndarray  # (n_seg + 1,) int64, host

    @classmethod
    def from_lengths(cls, lengths: Sequence[int]) -> "Segments":
        lengths = [int(v) for v in lengths]
        if any(v < 0 for v in lengths):
            raise ValueError("trajectory lengths must be non-negative")
        if not lengths:
            lengths = [0]
        return cls(np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64))

    @property
    def n_frames(self) -> int:
        return int(self.offsets[-1])

    @property
    def lengths(self) -> np.ndarray:
        return np.diff(self.offsets)

    def n_pairs(self, lag: int, step: int = 1) -> int:
        """Number of (t, t+lag) pairs, ``1 + (L - lag - 1)//step`` per shard
        (src/pmarlo/analysis/counting.py:10-68)."""
        eff = self.lengths - int(lag)
        eff = eff[eff > 0]
        return int(np.sum(1 + (eff - 1) // int(step))) if eff.size else 0

    def device(self, device) -> torch.Tensor:
        return torch.from_numpy(self.offsets).to(device)

    def split(self, flat: np.ndarray) -> list[np.ndarray]:
        return [flat[self.offsets[i]:self.offsets[i + 1]] for i in range(len(self.offsets) - 1)]

    def drop_tail(self, lag: int) -> tuple[np.ndarray, "Segments"]:
        """Row index that keeps all but the last ``lag`` frames of every shard
        (``_maybe_apply_tica``, _features.py:216-229) and the new segmentation."""
        keep, lens = [], []
        for s, e in zip(self.offsets[:-1], self.offsets[1:]):
            n = max(0, int(e - s) - int(lag)) if lag > 0 else int(e - s)
            keep.append(np.arange(s, s + n, dtype=np.int64))
            lens.append(n)
        idx = np.concatenate(keep) if keep else np.zeros((0,), dtype=np.int64)
        return idx, Segments.from_lengths(lens)


def partition_trajectories(lengths: Sequence[int], world_size: int) -> list[list[int]]:
    """Deal whole trajectories to ranks, longest first onto the least loaded rank
    (deterministic; ties -> lower rank).  Returns the trajectory indices per rank,
    each list in ascending order so that the per-rank concatenation is stable."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world_size
    parts: list[list[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        parts[r].append(i)
        load[r] += int(lengths[i])
    return [sorted(p) for p in parts]


def concat_to_device(arrays: Sequence, device, dtype=torch.float32) -> tuple[torch.Tensor, Segments]:
    """Concatenate per-trajectory arrays (numpy or torch) into one device buffer."""
    tensors = []
    for a in arrays:
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
        tensors.append(t.to(device=device, dtype=dtype, non_blocking=True))
    segs = Segments.from_lengths([int(t.shape[0]) for t in tensors])
    if not tensors:
        raise ValueError("no trajectories given")
    flat = tensors[0] if len(tensors) == 1 else torch.cat(tensors, dim=0)
    return flat.contiguous(), segs
