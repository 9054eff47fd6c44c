"""Oracle: backbone dihedrals, C-alpha distances, trig expansion (TEST INFRASTRUCTURE).

Restates, in numpy fp64, what pmarlo obtains from mdtraj (absent from this
image, version pinned at /root/reference/poetry.lock:1754-1790):

* ``mdtraj.geometry.dihedral._dihedral``: b1=x1-x0, b2=x2-x1, b3=x3-x2,
  c1=b2 x b3, c2=b1 x b2, angle = atan2((b1.c1)|b2|, c1.c2)
  -- called through src/pmarlo/features/featurize.py:41-44 and
  src/pmarlo/features/builtins.py:17-24.
* ``mdtraj.geometry.dihedral._atom_sequence``: phi = (C[i-1], N[i], CA[i], C[i]),
  psi = (N[i], CA[i], C[i], N[i+1]) for consecutive residues of one chain.
* ``mdtraj.compute_distances`` without unit cell: |x_j - x_i|
  -- src/pmarlo/features/featurize.py:46-53.

Parity note: mdtraj evaluates in float32; this oracle evaluates the same
formula in float64 from the float32 coordinates.  Angle/distance values are
PARITY UNPINNED against mdtraj itself (no reference test pins a value, see
SURVEY.md section 8c); they are pinned analytically in tests/test_oracle_featurize.py.
"""

from __future__ import annotations

from typing import Sequence

import numpy as np

__all__ = [
    "dihedral_quads",
    "ca_indices",
    "ca_pairs_all",
    "ca_pairs_stride3",
    "compute_dihedrals",
    "compute_distances",
    "wrap_to_minus_pi_pi",
    "trig_expand_periodic",
    "featurize_trajectory",
    "phi_psi_block_features",
]


def _atom_dict(names: Sequence[str], resid: Sequence[int], chainid: Sequence[int]):
    d: dict[int, dict[int, dict[str, int]]] = {}
    for i, (n, r, c) in enumerate(zip(names, resid, chainid)):
        d.setdefault(int(c), {}).setdefault(int(r), {})[str(n)] = i
    return d


def dihedral_quads(names, resid, chainid, kind: str) -> np.ndarray:
    """Atom quadruples for ``kind`` in {"phi","psi"} (mdtraj ``_atom_sequence``)."""
    if kind == "phi":
        atoms, offs = ("C", "N", "CA", "C"), (-1, 0, 0, 0)
    elif kind == "psi":
        atoms, offs = ("N", "CA", "C", "N"), (0, 0, 0, 1)
    else:
        raise ValueError(kind)
    d = _atom_dict(names, resid, chainid)
    out = []
    for c in sorted(d):
        for r in sorted(d[c]):
            if all((r + o) in d[c] for o in offs) and all(
                a in d[c][r + o] for a, o in zip(atoms, offs)
            ):
                out.append([d[c][r + o][a] for a, o in zip(atoms, offs)])
    return np.asarray(out, dtype=np.int32).reshape(-1, 4)


def ca_indices(names) -> np.ndarray:
    return np.asarray([i for i, n in enumerate(names) if n == "CA"], dtype=np.int32)


def ca_pairs_all(ca: np.ndarray) -> np.ndarray:
    """Row-major i<j enumeration, src/pmarlo/features/featurize.py:50-52."""
    return np.asarray(
        [(ca[i], ca[j]) for i in range(len(ca)) for j in range(i + 1, len(ca))],
        dtype=np.int32,
    ).reshape(-1, 2)


def ca_pairs_stride3(ca: np.ndarray, n_features: int | None) -> np.ndarray:
    """Stride-3 enumeration capped at ``n_features or 200``,
    src/pmarlo/markov_state_model/_features.py:155-171."""
    total = len(ca) * (len(ca) - 1) // 2
    n_pairs = min(n_features or 200, total)
    pairs = []
    for i in range(0, len(ca), 3):
        for j in range(i + 3, len(ca), 3):
            pairs.append((int(ca[i]), int(ca[j])))
            if len(pairs) >= n_pairs:
                break
        if len(pairs) >= n_pairs:
            break
    return np.asarray(pairs, dtype=np.int32).reshape(-1, 2)


def compute_dihedrals(xyz: np.ndarray, quads: np.ndarray) -> np.ndarray:
    """(N,A,3) float32, (nq,4) -> (N,nq) float64 radians in (-pi, pi]."""
    x = np.asarray(xyz, dtype=np.float64)
    q = np.asarray(quads, dtype=np.int64).reshape(-1, 4)
    if q.shape[0] == 0:
        return np.zeros((x.shape[0], 0))
    b1 = x[:, q[:, 1]] - x[:, q[:, 0]]
    b2 = x[:, q[:, 2]] - x[:, q[:, 1]]
    b3 = x[:, q[:, 3]] - x[:, q[:, 2]]
    c1 = np.cross(b2, b3)
    c2 = np.cross(b1, b2)
    p1 = np.sum(b1 * c1, axis=-1) * np.sqrt(np.sum(b2 * b2, axis=-1))
    p2 = np.sum(c1 * c2, axis=-1)
    return np.arctan2(p1, p2)


def compute_distances(xyz: np.ndarray, pairs: np.ndarray) -> np.ndarray:
    x = np.asarray(xyz, dtype=np.float64)
    p = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    if p.shape[0] == 0:
        return np.zeros((x.shape[0], 0))
    d = x[:, p[:, 1]] - x[:, p[:, 0]]
    return np.sqrt(np.sum(d * d, axis=-1))


def wrap_to_minus_pi_pi(angles: np.ndarray) -> np.ndarray:
    """src/pmarlo/features/builtins.py:11-14."""
    wrapped = ((angles + np.pi) % (2 * np.pi)) - np.pi
    return np.where(wrapped <= -np.pi, wrapped + 2 * np.pi, wrapped)


def trig_expand_periodic(X: np.ndarray, periodic: np.ndarray):
    """src/pmarlo/api/features.py:138-180 (interleaved cos/sin + mapping)."""
    if X.size == 0:
        return X, np.array([], dtype=int)
    if periodic.size != X.shape[1]:
        raise ValueError(
            f"periodic array size ({periodic.size}) must match number of features ({X.shape[1]})"
        )
    cols, mapping = [], []
    for j in range(X.shape[1]):
        col = X[:, j]
        if bool(periodic[j]):
            cols += [np.cos(col), np.sin(col)]
            mapping += [j, j]
        else:
            cols.append(col)
            mapping.append(j)
    return np.vstack(cols).T, np.asarray(mapping, dtype=int)


def featurize_trajectory(xyz, names, resid, chainid, feature_type="phi_psi"):
    """src/pmarlo/features/featurize.py:17-66 (phi_psi / ca_distances)."""
    if feature_type == "phi_psi":
        phi = compute_dihedrals(xyz, dihedral_quads(names, resid, chainid, "phi"))
        psi = compute_dihedrals(xyz, dihedral_quads(names, resid, chainid, "psi"))
        return np.concatenate([phi, psi], axis=1)
    if feature_type == "ca_distances":
        ca = ca_indices(names)
        if len(ca) < 2:
            raise ValueError("Topology has fewer than 2 Cα atoms.")
        return compute_distances(xyz, ca_pairs_all(ca))
    raise ValueError(f"Unknown feature_type {feature_type!r}.")


def phi_psi_block_features(xyz, names, resid, chainid):
    """[cos phi | sin phi | cos psi | sin psi] block layout,
    src/pmarlo/markov_state_model/_features.py:131-142."""
    phi = compute_dihedrals(xyz, dihedral_quads(names, resid, chainid, "phi"))
    psi = compute_dihedrals(xyz, dihedral_quads(names, resid, chainid, "psi"))
    blocks = []
    if phi.shape[1]:
        blocks += [np.cos(phi), np.sin(phi)]
    if psi.shape[1]:
        blocks += [np.cos(psi), np.sin(psi)]
    return np.hstack(blocks)
