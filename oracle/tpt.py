"""Oracle: transition-path theory quantities (TEST INFRASTRUCTURE).

Restates what deeptime ``MarkovStateModel(T, stationary_distribution=pi).reactive_flux(A, B)`` returns to
src/pmarlo/conformations/tpt_analysis.py:95-140 (published: Metzner, Schuette, Vanden-Eijnden 2009; Noe et al.
PNAS 2009): forward committor q+ ((I - T) q = 0 off A u B, q = 0 on A, 1 on B), backward committor q- (same for
the time-reversed chain, 1 on A), gross flux f_ij = pi_i q-_i T_ij q+_j (i != j), net flux max(0, f_ij - f_ji),
total flux sum_{i in A, j not in A} f_ij, rate F / sum_i pi_i q-_i, MFPT 1 / rate.  PARITY UNPINNED against
deeptime; pinned by textbook cases in tests (two-state and linear chains with known committors)."""

from __future__ import annotations

import numpy as np


def committor(P, A, B):
    n = P.shape[0]
    q = np.zeros(n)
    q[B] = 1.0
    inter = np.setdiff1d(np.arange(n), np.concatenate([A, B]))
    if inter.size:
        L = np.eye(inter.size) - P[np.ix_(inter, inter)]
        q[inter] = np.linalg.solve(L, P[np.ix_(inter, B)].sum(axis=1))
    return q


def reactive_flux(T, pi, A, B):
    T, pi = np.asarray(T, dtype=float), np.asarray(pi, dtype=float)
    A, B = np.unique(A).astype(int), np.unique(B).astype(int)
    qf = committor(T, A, B)
    Trev = (pi[None, :] * T.T) / pi[:, None]
    qb = committor(Trev, B, A)
    gross = pi[:, None] * qb[:, None] * T * qf[None, :]
    np.fill_diagonal(gross, 0.0)
    net = np.maximum(gross - gross.T, 0.0)
    notA = np.setdiff1d(np.arange(T.shape[0]), A)
    F = float(gross[np.ix_(A, notA)].sum())
    rate = F / float((pi * qb).sum())
    return {"qf": qf, "qb": qb, "gross": gross, "net": net, "total_flux": F, "rate": rate, "mfpt": 1.0 / rate}
