"""Macrostate bookkeeping downstream of (T, pi, labels) -- SURVEY.md 8f item 4, the small part.

Mirrors ``markov_state_model/_msm_utils.py``:
* ``compute_macro_populations(pi_micro, micro_to_macro)``  :103-117
* ``lump_micro_to_macro_T(T_micro, pi_micro, micro_to_macro)``  :120-137 -- the reference walks all K^2
  entries in a Python double loop; here the stationary flux F = M^T diag(pi) T M (M one-hot) is two
  index-adds on the device
* ``compute_macro_mfpt(T_macro)``  :140-162 -- n_macro <= a handful: the same n linear solves on the host
PCCA+ itself (deeptime) stays outside: pass its labels in.
"""

from __future__ import annotations

import numpy as np
import torch

from . import kernels

__all__ = ["compute_macro_populations", "lump_micro_to_macro_T", "compute_macro_mfpt"]


def compute_macro_populations(pi_micro: np.ndarray, micro_to_macro: np.ndarray) -> np.ndarray:
    micro_to_macro = np.asarray(micro_to_macro)
    n_macro = int(np.max(micro_to_macro)) + 1 if micro_to_macro.size else 0
    pi_macro = np.bincount(micro_to_macro.astype(np.int64), weights=np.asarray(pi_micro, dtype=float),
                           minlength=n_macro).astype(float) if n_macro else np.zeros((0,), dtype=float)
    s = float(np.sum(pi_macro))
    if s > 0:
        pi_macro /= s
    return pi_macro


def lump_micro_to_macro_T(T_micro: np.ndarray, pi_micro: np.ndarray, micro_to_macro: np.ndarray) -> np.ndarray:
    micro_to_macro = np.asarray(micro_to_macro)
    n_macro = int(np.max(micro_to_macro)) + 1 if micro_to_macro.size else 0
    if n_macro == 0:
        return np.zeros((0, 0), dtype=float)
    dev = kernels.require_cuda()
    T = torch.from_numpy(np.ascontiguousarray(T_micro, dtype=np.float64)).to(dev)
    pi = torch.from_numpy(np.ascontiguousarray(pi_micro, dtype=np.float64)).to(dev)
    lab = torch.from_numpy(micro_to_macro.astype(np.int64)).to(dev)
    flux = pi[:, None] * T                                                            # pi_i T_ij
    rows = torch.zeros((n_macro, T.shape[1]), dtype=torch.float64, device=dev).index_add_(0, lab[: T.shape[0]], flux)
    F = torch.zeros((n_macro, n_macro), dtype=torch.float64, device=dev).index_add_(1, lab[: T.shape[1]], rows)
    rs = F.sum(dim=1)
    rs = torch.where(rs == 0, torch.ones_like(rs), rs)
    return (F / rs[:, None]).cpu().numpy()


def compute_macro_mfpt(T_macro: np.ndarray) -> np.ndarray:
    T_macro = np.asarray(T_macro, dtype=float)
    n = T_macro.shape[0]
    mfpt = np.zeros((n, n), dtype=float)
    for j in range(n):
        mask = np.ones((n,), dtype=bool)
        mask[j] = False
        A = np.eye(n - 1) - T_macro[np.ix_(mask, mask)]
        try:
            t = np.linalg.solve(A, np.ones((n - 1,)))
        except np.linalg.LinAlgError:
            t = np.full((n - 1,), np.nan)
        mfpt[mask, j] = t
    return mfpt
