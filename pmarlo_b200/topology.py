"""Minimal topology / trajectory containers for the featurization boundary.

pmarlo passes ``mdtraj.Trajectory`` objects into ``featurize_trajectory`` /
``compute_features`` (src/pmarlo/features/featurize.py:17-20).  mdtraj is not a
dependency of this package; anything that exposes ``.xyz`` (n_frames, n_atoms, 3)
float32 nm and ``.topology`` with mdtraj-style ``atoms`` (``.name``,
``.residue.index``, ``.residue.chain.index``) is accepted, and so are the two
light-weight classes below.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import numpy as np

__all__ = ["Topology", "Trajectory", "load_pdb", "as_topology"]


@dataclass
class Topology:
    names: list[str]          # atom names ("N", "CA", "C", ...)
    resid: np.ndarray         # residue index per atom (0-based, consecutive along a chain)
    chainid: np.ndarray       # chain index per atom
    resnames: list[str] | None = None

    @property
    def n_atoms(self) -> int:
        return len(self.names)

    def select_name(self, name: str) -> np.ndarray:
        return np.asarray([i for i, n in enumerate(self.names) if n == name], dtype=np.int32)


@dataclass
class Trajectory:
    xyz: np.ndarray           # (n_frames, n_atoms, 3) float32, nanometres
    topology: Topology

    def __post_init__(self):
        self.xyz = np.ascontiguousarray(self.xyz, dtype=np.float32)
        if self.xyz.ndim != 3 or self.xyz.shape[2] != 3:
            raise ValueError("xyz must have shape (n_frames, n_atoms, 3)")
        if self.xyz.shape[1] != self.topology.n_atoms:
            raise ValueError("xyz and topology disagree on the number of atoms")

    @property
    def n_frames(self) -> int:
        return int(self.xyz.shape[0])

    @property
    def n_atoms(self) -> int:
        return int(self.xyz.shape[1])


def as_topology(top) -> Topology:
    """Accept our Topology or an mdtraj-like topology object."""
    if isinstance(top, Topology):
        return top
    atoms = list(top.atoms)
    names = [str(a.name) for a in atoms]
    resid = np.asarray([int(a.residue.index) for a in atoms], dtype=np.int64)
    chain = np.asarray([int(a.residue.chain.index) for a in atoms], dtype=np.int64)
    return Topology(names, resid, chain)


def load_pdb(path: str, models: Sequence[int] | None = None) -> Trajectory:
    """Tiny PDB reader (ATOM/HETATM records, MODEL blocks; Angstrom -> nm)."""
    names: list[str] = []
    resnames: list[str] = []
    reskeys: list[tuple] = []
    frames: list[list[tuple[float, float, float]]] = []
    cur: list[tuple[float, float, float]] = []
    first = True
    with open(path, "r", encoding="utf-8", errors="replace") as fh:
        for line in fh:
            rec = line[:6]
            if rec in ("ATOM  ", "HETATM"):
                cur.append((float(line[30:38]), float(line[38:46]), float(line[46:54])))
                if first:
                    names.append(line[12:16].strip())
                    resnames.append(line[17:20].strip())
                    reskeys.append((line[21], line[22:27]))
            elif rec.startswith("ENDMDL"):
                frames.append(cur)
                cur = []
                first = False
    if cur:
        frames.append(cur)
    if not frames:
        raise ValueError(f"{path}: no atoms found")
    resid = np.zeros(len(names), dtype=np.int64)
    chain = np.zeros(len(names), dtype=np.int64)
    chain_ids: dict[str, int] = {}
    r = -1
    prev = None
    for i, key in enumerate(reskeys):
        if key != prev:
            r += 1
            prev = key
        resid[i] = r
        chain[i] = chain_ids.setdefault(key[0], len(chain_ids))
    xyz = np.asarray(frames, dtype=np.float64) * 0.1
    if models is not None:
        xyz = xyz[list(models)]
    return Trajectory(xyz.astype(np.float32), Topology(names, resid, chain, resnames))
