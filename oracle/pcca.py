"""Oracle: PCCA+ memberships (TEST INFRASTRUCTURE).

Restates what deeptime ``deeptime.markov.pcca(T, m).memberships`` returns to
src/pmarlo/markov_state_model/_msm_utils.py:284-299 (``pcca_like_macrostates``), from the published algorithm:
Deuflhard & Weber, Lin. Alg. Appl. 398 (2005) (inner simplex algorithm); Roeblitz & Weber, Adv. Data Anal. Classif.
7 (2013) (the soft optimisation of the rotation matrix, "PCCA++"), in the form msmtools / deeptime implement it
for a connected reversible transition matrix:
  1. the m right eigenvectors of largest |eigenvalue|, normalised to sum_i pi_i r_i^2 = 1, first one made positive;
  2. inner simplex algorithm: the row of largest norm (first column dropped), then m - 1 Gram-Schmidt steps each
     picking the farthest remaining row; A = inverse of the eigenvector rows at those indices; chi = R A;
  3. Nelder-Mead (scipy.optimize.fmin, default tolerances) on the (m-1) x (m-1) lower-right block of A, the first
     row / column being fixed by the row-sum and positivity constraints, maximising sum_ij A_ji^2 / A_0i;
  4. memberships clipped to [0, 1] and rows renormalised.
Written with the same per-row Python loops as the source it restates.  PARITY UNPINNED against deeptime (absent
from the image); pinned by known answers in tests (block matrices, 3-state analytic case) and invariants."""

from __future__ import annotations

import math

import numpy as np
from scipy.optimize import fmin


def right_eigenvectors(T: np.ndarray, pi: np.ndarray, m: int) -> np.ndarray:
    """m right eigenvectors of largest |eigenvalue| of a reversible T, pi-normalised."""
    d = np.sqrt(pi)
    S = (d[:, None] * T) / d[None, :]
    S = 0.5 * (S + S.T)
    w, V = np.linalg.eigh(S)
    order = np.argsort(-np.abs(w), kind="stable")[:m]
    R = V[:, order] / d[:, None]
    for i in range(m):
        R[:, i] /= math.sqrt(float(np.dot(R[:, i] * pi, R[:, i])))
    R[:, 0] = np.abs(R[:, 0])
    return R


def inner_simplex(evecs: np.ndarray):
    n, m = evecs.shape
    c = evecs[:, 1:].copy()
    ortho = c.copy()
    ind = np.zeros(m, dtype=np.int64)
    max_dist = 0.0
    for i, row in enumerate(c):
        dist = float(np.linalg.norm(row, 2))
        if dist > max_dist:
            max_dist = dist
            ind[0] = i
    ortho -= c[ind[0]]
    for k in range(1, m):
        max_dist = 0.0
        temp = ortho[ind[k - 1]].copy()
        for i, row in enumerate(ortho):
            row -= np.dot(np.dot(temp, row), temp)
            dist = float(np.linalg.norm(row, 2))
            if dist > max_dist and i not in ind[0:k]:
                max_dist = dist
                ind[k] = i
        ortho /= max_dist
    rot = np.linalg.inv(evecs[ind])
    return evecs @ rot, rot, ind


def fill_matrix(crop: np.ndarray, evecs: np.ndarray) -> np.ndarray:
    x, y = crop.shape
    row_sums = np.sum(crop, axis=1).reshape(x, 1)
    crop = np.concatenate((-row_sums, crop), axis=1)
    tmp = -np.dot(evecs[:, 1:], crop)
    col_max = np.max(tmp, axis=0).reshape(1, y + 1)
    rot = np.concatenate((col_max, crop), axis=0)
    rot /= np.sum(col_max)
    return rot


def opt_soft(evecs: np.ndarray, rot: np.ndarray, m: int) -> np.ndarray:
    evecs = evecs[:, :m]
    crop = rot[1:, 1:]
    x, y = crop.shape

    def objective(vec):
        A = fill_matrix(vec.reshape(x, y), evecs)
        result = 0.0
        for i in range(m):
            for j in range(m):
                result += A[j, i] ** 2 / A[0, i]
        return -result

    best = fmin(objective, crop.reshape(x * y), disp=False)
    return fill_matrix(best.reshape(x, y), evecs)


def pcca_memberships(T: np.ndarray, m: int, pi: np.ndarray) -> np.ndarray:
    R = right_eigenvectors(np.asarray(T, dtype=float), np.asarray(pi, dtype=float), m)
    _, rot, _ = inner_simplex(R)
    rot = opt_soft(R, rot, m)
    chi = np.clip(R @ rot, 0.0, 1.0)
    for i in range(chi.shape[0]):
        chi[i] /= np.sum(chi[i])
    return chi
