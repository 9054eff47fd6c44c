"""Trajectory featurization behind pmarlo's call signatures (K1 on the device).

Mirrors
* ``featurize_trajectory(traj, feature_type)``      src/pmarlo/features/featurize.py:17-66
* ``compute_features(traj, specs, cache_path)``     src/pmarlo/api/features.py:192-208
* ``trig_expand_periodic(X, periodic)``             src/pmarlo/api/features.py:138-180
* ``FeaturesMixin._compute_phi_psi_features`` / ``_compute_distance_features``
                                                    src/pmarlo/markov_state_model/_features.py:131-171
The atom bookkeeping (which quadruples / pairs) is host logic; every angle,
distance, cosine and sine is evaluated by ``pmb_featurize`` / ``pmb_trig_expand``.
"""

from __future__ import annotations

import hashlib
import logging
import pathlib
import re
from dataclasses import dataclass, field
from typing import Sequence

import numpy as np
import torch

from . import kernels
from .topology import Topology, as_topology

logger = logging.getLogger("pmarlo")

__all__ = [
    "FeaturePlan", "dihedral_quads", "ca_pairs_all", "ca_pairs_stride3", "plan_phi_psi",
    "plan_ca_distances", "plan_phi_psi_block", "plan_concat", "featurize_device",
    "featurize_trajectory", "compute_features", "trig_expand_periodic", "parse_feature_spec",
]

_SUPPORTED = ("phi_psi", "ca_distances", "backbone_torsions")
KIND_DIHEDRAL, KIND_DISTANCE = 0, 1


# --------------------------------------------------------------------------- atom bookkeeping
def dihedral_quads(top: Topology, kind: str) -> np.ndarray:
    """phi = (C[i-1], N[i], CA[i], C[i]); psi = (N[i], CA[i], C[i], N[i+1]) for
    consecutive residues of one chain (mdtraj ``_atom_sequence`` as used by
    ``md.compute_phi`` / ``compute_psi``, features/featurize.py:42-43)."""
    if kind == "phi":
        atoms, offs = ("C", "N", "CA", "C"), (-1, 0, 0, 0)
    elif kind == "psi":
        atoms, offs = ("N", "CA", "C", "N"), (0, 0, 0, 1)
    else:
        raise ValueError(f"unknown dihedral kind {kind!r}")
    table: dict[int, dict[int, dict[str, int]]] = {}
    for i, (n, r, c) in enumerate(zip(top.names, top.resid, top.chainid)):
        table.setdefault(int(c), {}).setdefault(int(r), {})[str(n)] = i
    quads = []
    for c in sorted(table):
        res = table[c]
        for r in sorted(res):
            ok = all((r + o) in res and a in res[r + o] for a, o in zip(atoms, offs))
            if ok:
                quads.append([res[r + o][a] for a, o in zip(atoms, offs)])
    return np.asarray(quads, dtype=np.int32).reshape(-1, 4)


def ca_pairs_all(ca: np.ndarray) -> np.ndarray:
    """All i<j pairs in row-major order (features/featurize.py:50-52)."""
    n = len(ca)
    iu, ju = np.triu_indices(n, k=1)
    return np.stack([ca[iu], ca[ju]], axis=1).astype(np.int32).reshape(-1, 2)


def ca_pairs_stride3(ca: np.ndarray, n_features: int | None) -> np.ndarray:
    """Stride-3 pair enumeration capped at ``n_features or 200`` (_features.py:155-171)."""
    total = len(ca) * (len(ca) - 1) // 2
    n_pairs = min(n_features or 200, total)
    pairs: list[tuple[int, int]] = []
    for i in range(0, len(ca), 3):
        for j in range(i + 3, len(ca), 3):
            pairs.append((int(ca[i]), int(ca[j])))
            if len(pairs) >= n_pairs:
                break
        if len(pairs) >= n_pairs:
            break
    return np.asarray(pairs, dtype=np.int32).reshape(-1, 2)


# --------------------------------------------------------------------------- plans
@dataclass
class FeaturePlan:
    """Work list of the featurize kernel: one unit per dihedral / distance with
    the output columns of its value, cosine and sine (-1 = not written)."""

    units: np.ndarray                      # (n_units, 8) int32
    n_cols: int
    columns: list[str] = field(default_factory=list)
    periodic: np.ndarray = field(default_factory=lambda: np.zeros((0,), dtype=bool))

    def device_units(self, device) -> torch.Tensor:
        """Unit list on the device (uploaded once per device and cached)."""
        cache = self.__dict__.setdefault("_dev_units", {})
        key = str(device)
        if key not in cache:
            cache[key] = torch.from_numpy(np.ascontiguousarray(self.units, dtype=np.int32)).to(device)
        return cache[key]


def _unit(kind, atoms, col_value=-1, col_cos=-1, col_sin=-1):
    a = list(atoms) + [0] * (4 - len(atoms))
    return [kind, a[0], a[1], a[2], a[3], col_value, col_cos, col_sin]


def _labels(top: Topology, quads: np.ndarray, kind: str, atom_pos: int) -> list[str]:
    return [f"{kind}:res{int(top.resid[int(q[atom_pos])])}" for q in quads]


def plan_phi_psi(top: Topology) -> FeaturePlan:
    """Angles [phi..., psi...] in radians, (-pi, pi] (features/builtins.py:42-86)."""
    phi, psi = dihedral_quads(top, "phi"), dihedral_quads(top, "psi")
    units = [_unit(KIND_DIHEDRAL, q, col_value=i) for i, q in enumerate(phi)]
    units += [_unit(KIND_DIHEDRAL, q, col_value=len(phi) + i) for i, q in enumerate(psi)]
    n = len(phi) + len(psi)
    cols = _labels(top, phi, "phi", 1) + _labels(top, psi, "psi", 2)
    return FeaturePlan(np.asarray(units, dtype=np.int32).reshape(-1, 8), n, cols, np.ones(n, dtype=bool))


def plan_phi_psi_interleaved(top: Topology) -> FeaturePlan:
    """[cos x0, sin x0, cos x1, sin x1, ...] over [phi..., psi...]: compute_features("phi_psi")
    followed by trig_expand_periodic, fused into one pass over the coordinates."""
    phi, psi = dihedral_quads(top, "phi"), dihedral_quads(top, "psi")
    quads = list(phi) + list(psi)
    units = [_unit(KIND_DIHEDRAL, q, col_cos=2 * i, col_sin=2 * i + 1) for i, q in enumerate(quads)]
    base = _labels(top, phi, "phi", 1) + _labels(top, psi, "psi", 2)
    cols = [f"{f}({b})" for b in base for f in ("cos", "sin")]
    n = 2 * len(quads)
    return FeaturePlan(np.asarray(units, dtype=np.int32).reshape(-1, 8), n, cols, np.zeros(n, dtype=bool))


def plan_phi_psi_block(top: Topology) -> FeaturePlan:
    """[cos phi | sin phi | cos psi | sin psi] (_features.py:131-142)."""
    phi, psi = dihedral_quads(top, "phi"), dihedral_quads(top, "psi")
    nphi, npsi = len(phi), len(psi)
    units = [_unit(KIND_DIHEDRAL, q, col_cos=i, col_sin=nphi + i) for i, q in enumerate(phi)]
    units += [_unit(KIND_DIHEDRAL, q, col_cos=2 * nphi + i, col_sin=2 * nphi + npsi + i)
              for i, q in enumerate(psi)]
    n = 2 * (nphi + npsi)
    cols = ([f"cos(phi{i})" for i in range(nphi)] + [f"sin(phi{i})" for i in range(nphi)]
            + [f"cos(psi{i})" for i in range(npsi)] + [f"sin(psi{i})" for i in range(npsi)])
    return FeaturePlan(np.asarray(units, dtype=np.int32).reshape(-1, 8), n, cols, np.zeros(n, dtype=bool))


def plan_distances(pairs: np.ndarray) -> FeaturePlan:
    pairs = np.asarray(pairs, dtype=np.int32).reshape(-1, 2)
    units = [_unit(KIND_DISTANCE, p, col_value=i) for i, p in enumerate(pairs)]
    cols = [f"distance([{int(p[0])}, {int(p[1])}])" for p in pairs]
    n = len(pairs)
    return FeaturePlan(np.asarray(units, dtype=np.int32).reshape(-1, 8), n, cols, np.zeros(n, dtype=bool))


def plan_ca_distances(top: Topology) -> FeaturePlan:
    ca = top.select_name("CA")
    if len(ca) < 2:
        raise ValueError("Topology has fewer than 2 Cα atoms.")
    return plan_distances(ca_pairs_all(ca))


def plan_concat(plans: Sequence[FeaturePlan]) -> FeaturePlan:
    """Column-wise concatenation (one kernel pass computes all blocks)."""
    units, cols, per, off = [], [], [], 0
    for p in plans:
        u = np.array(p.units, dtype=np.int32, copy=True).reshape(-1, 8)
        for c in (5, 6, 7):
            u[:, c] = np.where(u[:, c] >= 0, u[:, c] + off, -1)
        units.append(u)
        cols += list(p.columns)
        per.append(np.asarray(p.periodic, dtype=bool))
        off += p.n_cols
    return FeaturePlan(np.concatenate(units, axis=0) if units else np.zeros((0, 8), np.int32), off, cols,
                       np.concatenate(per) if per else np.zeros((0,), dtype=bool))


# --------------------------------------------------------------------------- device entry
def featurize_device(xyz: torch.Tensor, plan: FeaturePlan, out: torch.Tensor | None = None) -> torch.Tensor:
    """xyz (N,A,3) float32 CUDA tensor -> (N, plan.n_cols) float32 on the device."""
    if plan.n_cols == 0 or len(plan.units) == 0:
        return torch.zeros((int(xyz.shape[0]), plan.n_cols), dtype=torch.float32, device=xyz.device)
    return kernels.featurize(xyz, plan.device_units(xyz.device), plan.n_cols, out=out)


def _xyz_to_device(traj) -> torch.Tensor:
    dev = kernels.require_cuda()
    xyz = traj.xyz
    if isinstance(xyz, torch.Tensor):
        return xyz.to(device=dev, dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(xyz, dtype=np.float32)).to(dev)


# --------------------------------------------------------------------------- pmarlo signatures
def featurize_trajectory(traj, feature_type: str = "phi_psi") -> np.ndarray:
    """Drop-in for ``pmarlo.features.featurize.featurize_trajectory`` (float32 out,
    like mdtraj).  ``"backbone_torsions"`` needs chi1 side-chain torsions, which
    are outside the accelerated path."""
    top = as_topology(traj.topology)
    if feature_type == "phi_psi":
        plan = plan_phi_psi(top)
    elif feature_type == "ca_distances":
        plan = plan_ca_distances(top)
    elif feature_type == "backbone_torsions":
        raise NotImplementedError("backbone_torsions (chi1) is outside the B200 hot path")
    else:
        raise ValueError(f"Unknown feature_type {feature_type!r}. Choose one of {_SUPPORTED}.")
    return featurize_device(_xyz_to_device(traj), plan).cpu().numpy()


_LIST_SPEC = re.compile(r"^\s*(distance|dihedral)\s*\(\s*\[([^\]]*)\]\s*\)\s*$", re.IGNORECASE)
_PAIR_SPEC = re.compile(r"^\s*(dist|distance|distances)\s*:\s*atompair\s*\(\s*(\d+)\s*,\s*(\d+)\s*\)\s*$",
                        re.IGNORECASE)


def parse_feature_spec(spec: str):
    """The part of ``features/base.py:129-174`` that names features on this path:
    ``phi_psi``, ``distance([i, j])``, ``dist:atompair(i,j)``, ``dihedral([i,j,k,l])``."""
    s = str(spec).strip()
    m = _LIST_SPEC.match(s)
    if m:
        idx = [int(v) for v in m.group(2).replace(" ", "").split(",") if v != ""]
        return m.group(1).lower(), {"indices": idx}
    m = _PAIR_SPEC.match(s)
    if m:
        return "distance_pair", {"i": int(m.group(2)), "j": int(m.group(3))}
    return s.lower() if ":" not in s else s, {}


def _plan_for_spec(top: Topology, n_atoms: int, name: str, kwargs: dict) -> tuple[FeaturePlan, type]:
    if name == "phi_psi":
        return plan_phi_psi(top), np.float32
    if name in ("distance", "distance_pair"):
        idx = kwargs.get("indices") if name == "distance" else [kwargs["i"], kwargs["j"]]
        if idx is None or len(idx) != 2:
            raise ValueError(
                f"Distance feature requires 'indices' with exactly 2 atom indices, got {idx}")
        i, j = int(idx[0]), int(idx[1])
        if not (0 <= i < n_atoms) or not (0 <= j < n_atoms):
            raise ValueError(f"Atom indices {i}, {j} out of range [0, {n_atoms})")
        p = plan_distances(np.asarray([[i, j]]))
        p.columns = [f"distance({list(idx)})"] if name == "distance" else [f"dist:atompair({i},{j})"]
        return p, np.float64
    if name == "dihedral":
        idx = kwargs.get("indices")
        if idx is None or len(idx) != 4:
            raise ValueError(f"Dihedral feature requires 'indices' with exactly 4 atom indices, got {idx}")
        if any(not (0 <= int(a) < n_atoms) for a in idx):
            raise ValueError(f"Atom indices {idx} out of range [0, {n_atoms})")
        p = FeaturePlan(np.asarray([_unit(KIND_DIHEDRAL, [int(a) for a in idx], col_value=0)], np.int32),
                        1, [f"dihedral({list(idx)})"], np.ones(1, dtype=bool))
        return p, np.float64
    raise KeyError(f"Feature {name!r} is not available on the B200 hot path "
                   "(supported: phi_psi, distance([i, j]), dist:atompair(i,j), dihedral([i,j,k,l]))")


def _cache_file(traj, specs, cache_path) -> pathlib.Path | None:
    if not cache_path:
        return None
    xyz = np.asarray(traj.xyz)
    h = hashlib.sha1()
    h.update(repr((xyz.shape, tuple(str(s) for s in specs))).encode())
    top = as_topology(traj.topology)
    h.update(",".join(top.names).encode())
    if xyz.size:
        sample = xyz[:: max(1, xyz.shape[0] // 16)]
        h.update(np.round(sample * 1000.0).astype(np.int64).tobytes())
    d = pathlib.Path(cache_path)
    d.mkdir(parents=True, exist_ok=True)
    return d / f"features_{h.hexdigest()}.npz"


def compute_features(traj, feature_specs: Sequence[str], cache_path: str | None = None):
    """Drop-in for ``pmarlo.api.features.compute_features``: returns ``(X, columns, periodic)``."""
    cf = _cache_file(traj, feature_specs, cache_path)
    if cf is not None and cf.exists():
        with np.load(cf, allow_pickle=False) as z:
            return z["X"], [str(c) for c in z["columns"]], z["periodic"].astype(bool)
    top = as_topology(traj.topology)
    n_frames = int(np.asarray(traj.xyz).shape[0]) if not isinstance(traj.xyz, torch.Tensor) else int(traj.xyz.shape[0])
    n_atoms = top.n_atoms
    plans, dtypes = [], []
    for spec in feature_specs:
        name, kwargs = parse_feature_spec(spec)
        p, dt = _plan_for_spec(top, n_atoms, name, kwargs)
        plans.append(p)
        dtypes.append(dt)
    if not plans:
        return np.empty((n_frames, 0), dtype=float), [], np.empty((0,), dtype=bool)
    plan = plan_concat(plans)
    X = featurize_device(_xyz_to_device(traj), plan).cpu().numpy()
    out_dtype = np.result_type(*dtypes) if dtypes else np.float64
    X = X.astype(out_dtype, copy=False)
    # distances: NaN / inf -> 0 like features/builtins.py:308
    dist_cols = np.flatnonzero(plan.units[:, 0] == KIND_DISTANCE)
    if dist_cols.size:
        cols = plan.units[dist_cols, 5]
        X[:, cols] = np.nan_to_num(X[:, cols], nan=0.0, posinf=0.0, neginf=0.0)
    columns, periodic = list(plan.columns), np.asarray(plan.periodic, dtype=bool)
    if cf is not None:
        np.savez(cf, X=X, columns=np.asarray(columns, dtype=str), periodic=periodic)
    return X, columns, periodic


def trig_expand_periodic(X: np.ndarray, periodic: np.ndarray):
    """Drop-in for ``pmarlo.api.features.trig_expand_periodic``: periodic columns
    become interleaved ``[cos, sin]`` pairs; returns ``(Xe, mapping)``."""
    X = np.asarray(X)
    periodic = np.asarray(periodic)
    if X.size == 0:
        return X, np.array([], dtype=int)
    if periodic.size != X.shape[1]:
        raise ValueError(
            f"periodic array size ({periodic.size}) must match number of features ({X.shape[1]})")
    per = periodic.astype(bool)
    width = np.where(per, 2, 1)
    out_col = np.concatenate([[0], np.cumsum(width)[:-1]]).astype(np.int32)
    Fe = int(width.sum())
    mapping = np.repeat(np.arange(X.shape[1]), width).astype(int)
    dev = kernels.require_cuda()
    Xd = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64)).to(dev)
    Xe = kernels.trig_expand(Xd, torch.from_numpy(per.astype(np.uint8)).to(dev),
                             torch.from_numpy(out_col).to(dev), Fe)
    return Xe.cpu().numpy(), mapping
