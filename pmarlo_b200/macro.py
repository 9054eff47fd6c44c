"""Macrostate bookkeeping downstream of (T, pi, labels) -- SURVEY.md 8f item 4, the small part.

Mirrors ``markov_state_model/_msm_utils.py``:
* ``compute_macro_populations(pi_micro, micro_to_macro)``  :103-117
* ``lump_micro_to_macro_T(T_micro, pi_micro, micro_to_macro)``  :120-137 -- the reference walks all K^2
  entries in a Python double loop; here the stationary flux F = M^T diag(pi) T M (M one-hot) is two
  index-adds on the device
* ``compute_macro_mfpt(T_macro)``  :140-162 -- n_macro <= a handful: the same n linear solves on the host
* ``pcca_like_macrostates(T, n_macrostates, random_state)``  :284-299 and ``_canonicalize_macro_labels``  :91-100 --
  PCCA+ as deeptime's ``pcca`` implements it for a connected reversible matrix (inner simplex algorithm, then the
  Nelder-Mead refinement of the rotation matrix).  The K x K work (symmetrised eigen-decomposition, the simplex
  search over K rows) runs on the device; the optimiser walks an (m-1)^2-dimensional simplex on the host like the
  reference's dependency does (scipy ``fmin``), its objective being a K x m product.  A matrix that is not
  reversible or not connected raises ValueError inside, which the wrapper turns into ``None`` exactly like
  the reference (``except ValueError``).
"""

from __future__ import annotations

import numpy as np
import torch

from . import kernels

__all__ = ["compute_macro_populations", "lump_micro_to_macro_T", "compute_macro_mfpt", "pcca_memberships",
           "pcca_like_macrostates"]


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


# ------------------------------------------------------------------------------------------------ PCCA+
def _stationary_reversible(T: torch.Tensor) -> torch.Tensor:
    """pi of a connected reversible chain from detailed balance along a spanning tree would need the graph; the
    leading left eigenvector of the K x K matrix is one symmetric-free solve here: (T^T - I) pi = 0 with the
    normalisation row replacing the last equation (the same system `_stationary_from_T` hands to deeptime)."""
    K = T.shape[0]
    A = T.T - torch.eye(K, dtype=T.dtype, device=T.device)
    A[-1, :] = 1.0
    b = torch.zeros((K,), dtype=T.dtype, device=T.device)
    b[-1] = 1.0
    return torch.linalg.solve(A, b)


def _inner_simplex(evecs: torch.Tensor):
    """Inner simplex algorithm on the device: the same vertex sequence as the per-row loops of the oracle
    (first maximum wins), each Gram-Schmidt step one rank-1 update of the K x (m-1) coordinates."""
    m = evecs.shape[1]
    c = evecs[:, 1:]
    ortho = c.clone()
    ind = [int(torch.argmax(torch.linalg.vector_norm(c, dim=1)).item())]
    ortho -= c[ind[0]].clone()
    for k in range(1, m):
        temp = ortho[ind[k - 1]].clone()
        ortho -= torch.outer(ortho @ temp, temp)
        dist = torch.linalg.vector_norm(ortho, dim=1)
        dist[torch.tensor(ind, device=dist.device)] = -1.0
        nxt = int(torch.argmax(dist).item())
        ortho /= dist[nxt]
        ind.append(nxt)
    idx = torch.tensor(ind, device=evecs.device)
    rot = torch.linalg.inv(evecs[idx])
    return rot, ind


def _fill_matrix(crop: np.ndarray, evecs: np.ndarray) -> np.ndarray:
    x, y = crop.shape
    crop = np.concatenate((-np.sum(crop, axis=1).reshape(x, 1), crop), axis=1)
    col_max = np.max(-(evecs[:, 1:] @ crop), axis=0).reshape(1, y + 1)
    rot = np.concatenate((col_max, crop), axis=0)
    return rot / np.sum(col_max)


def pcca_memberships(T: np.ndarray, m: int, pi: np.ndarray | None = None) -> np.ndarray:
    """Membership matrix chi (K x m) of PCCA+ for a connected, reversible transition matrix
    (deeptime ``pcca(T, m).memberships``).  Raises ValueError for anything else, like deeptime."""
    from scipy.optimize import fmin
    from scipy.sparse.csgraph import connected_components

    T_np = np.asarray(T, dtype=np.float64)
    K = T_np.shape[0]
    if T_np.ndim != 2 or T_np.shape[1] != K:
        raise ValueError("transition matrix must be square")
    if m <= 0 or m > K:
        raise ValueError(f"number of metastable sets must be in [1, {K}], got {m}")
    n_comp, _ = connected_components(T_np > 0, directed=True, connection="strong")
    if n_comp != 1:
        raise ValueError("PCCA+ here needs a connected transition matrix")
    dev = kernels.require_cuda()
    Td = torch.from_numpy(np.ascontiguousarray(T_np)).to(dev)
    pid = _stationary_reversible(Td) if pi is None else torch.from_numpy(np.asarray(pi, dtype=np.float64)).to(dev)
    if not bool(torch.all(pid > 0)):
        raise ValueError("stationary distribution must be positive")
    flux = pid[:, None] * Td
    if float((flux - flux.T).abs().max().item()) > 1e-8 * float(flux.abs().max().item()) + 1e-15:
        raise ValueError("PCCA+ needs a reversible transition matrix")
    # right eigenvectors of largest |eigenvalue| through the symmetrised matrix, pi-normalised
    dsq = torch.sqrt(pid)
    S = (dsq[:, None] * Td) / dsq[None, :]
    w, V = torch.linalg.eigh(0.5 * (S + S.T))
    order = torch.argsort(-w.abs(), stable=True)[:m]
    R = V[:, order] / dsq[:, None]
    R = R / torch.sqrt((R * R * pid[:, None]).sum(dim=0))[None, :]
    R[:, 0] = R[:, 0].abs()
    if m == 1:
        return np.ones((K, 1), dtype=float)
    rot, _ = _inner_simplex(R)
    R_np, rot_np = R.cpu().numpy(), rot.cpu().numpy()
    crop = rot_np[1:, 1:]
    x, y = crop.shape

    def objective(vec):
        A = _fill_matrix(vec.reshape(x, y), R_np)
        return -float(np.sum(A * A / A[0][None, :]))

    best = fmin(objective, crop.reshape(x * y), disp=False)
    A = _fill_matrix(best.reshape(x, y), R_np)
    chi = np.clip(R_np @ A, 0.0, 1.0)
    return chi / chi.sum(axis=1, keepdims=True)


def _canonicalize_macro_labels(labels: np.ndarray, T: np.ndarray) -> np.ndarray:
    """Macrostate ids renumbered by descending population (reference :91-100)."""
    labels = np.asarray(labels)
    if labels.size == 0:
        return labels.astype(int)
    dev = kernels.require_cuda()
    pi_micro = _stationary_reversible(torch.from_numpy(np.ascontiguousarray(T, dtype=np.float64)).to(dev)).cpu().numpy()
    pi_micro = np.abs(pi_micro) / np.sum(np.abs(pi_micro))
    pops = compute_macro_populations(pi_micro, labels)
    unique = np.unique(labels)
    order = np.argsort(-pops[unique])
    mapping = {int(unique[idx]): int(i) for i, idx in enumerate(order)}
    return np.asarray([mapping[int(lbl)] for lbl in labels], dtype=int)


def pcca_like_macrostates(T: np.ndarray, n_macrostates: int = 4, random_state: int | None = 42) -> np.ndarray | None:
    """Metastable sets by PCCA+ (reference ``pcca_like_macrostates`` :284-299): hard labels = argmax of the
    memberships, renumbered by descending population; ``None`` when the matrix is too small or PCCA+ rejects it."""
    T = np.asarray(T, dtype=float)
    if T.size == 0 or T.shape[0] <= n_macrostates:
        return None
    _ = random_state   # kept for API stability, like the reference
    try:
        chi = pcca_memberships(T, int(n_macrostates))
    except ValueError:
        return None
    labels = np.argmax(chi, axis=1)
    return _canonicalize_macro_labels(labels.astype(int), T)
