"""Downstream consumers of (T, pi, dtrajs) -- SURVEY.md 8f-4 -- on the device where there is per-frame or
K x K work: the 2-D free-energy surface and transition-path theory.

* ``generate_2d_fes`` / ``free_energy_from_density`` / ``FESResult``: markov_state_model/free_energy.py:43-313,417-865
  (histogram branch, smoothing "never" -- the reference default; the "adaptive" grid crops to the [1, 99]
  percentiles).  The O(N) part is ``pmb_hist2d``.
* ``TPTAnalysis`` / ``TPTResult``: conformations/tpt_analysis.py:35-260, which drives deeptime's
  ``MarkovStateModel.reactive_flux``: committors from two K x K linear solves (torch.linalg on the device, a
  library call on a cold path), gross / net flux, total flux, rate, MFPT, and the dominant pathways by
  bottleneck decomposition of the net flux (graph search on the K x K matrix, host).
PCCA+ (deeptime ``pcca``) lives in macro.py (``pcca_memberships`` / ``pcca_like_macrostates``).
"""

from __future__ import annotations

import logging
from dataclasses import dataclass, field
from typing import Any, Optional, Sequence, Tuple

import numpy as np
import torch

from . import kernels

logger = logging.getLogger("pmarlo")

__all__ = ["kT_kJ_per_mol", "free_energy_from_density", "FESResult", "generate_2d_fes", "TPTResult", "TPTAnalysis"]

_K_B = 1.380649e-23          # J / K           (scipy.constants.k)
_N_A = 6.02214076e23         # 1 / mol         (scipy.constants.Avogadro)


def kT_kJ_per_mol(temperature_kelvin: float) -> float:
    """utils/thermodynamics.py:8-25."""
    return float(_K_B * float(temperature_kelvin) * _N_A / 1000.0)


def free_energy_from_density(density, temperature: float, *, mask=None, inpaint: bool = False, tiny: float | None = None):
    """free_energy.py:257-310: F = -kT ln(density) shifted to min 0; +inf where the density is zero, NaN on
    masked (empty) bins unless they were inpainted."""
    if temperature <= 0:
        raise ValueError("temperature must be positive when computing free energy")
    d = np.array(density, dtype=np.float64)
    tiny_val = float(tiny if tiny is not None else np.finfo(np.float64).tiny)
    kT = kT_kJ_per_mol(float(temperature))
    with np.errstate(divide="ignore", invalid="ignore"):
        fe = -kT * np.log(np.clip(d, tiny_val, None))
    result = np.where(d > tiny_val, fe, np.inf)
    if mask is not None and not inpaint:
        result = np.where(np.array(mask, dtype=bool), np.nan, result)
    if np.any(np.isfinite(result)):
        result = result - np.nanmin(result)
    return result


@dataclass
class FESResult:
    """Fields of free_energy.py:43-130 that the histogram branch fills."""

    F: np.ndarray
    xedges: np.ndarray
    yedges: np.ndarray
    levels_kJmol: np.ndarray | None = None
    metadata: dict = field(default_factory=dict)

    @property
    def free_energy(self) -> np.ndarray:
        return self.F

    @property
    def output_shape(self) -> Tuple[int, int]:
        return (int(self.F.shape[0]), int(self.F.shape[1]))


def generate_2d_fes(cv1, cv2, bins: Tuple[int, int] = (100, 100), temperature: float = 300.0,
                    periodic: Tuple[bool, bool] = (False, False),
                    ranges: Optional[Tuple[Tuple[float, float], Tuple[float, float]]] = None, min_count: int = 1,
                    weights=None, grid_strategy: str = "fixed") -> FESResult:
    """Histogram free-energy surface: H = histogram2d(cv1, cv2) on the device, density = H / (sum H dx dy),
    bins with fewer than ``min_count`` samples masked (NaN), F = -kT ln(density) - min.  Periodic axes are
    wrapped to [-180, 180) degrees like the reference and span exactly that range."""
    if grid_strategy not in ("fixed", "adaptive"):
        raise ValueError(f"unknown grid_strategy {grid_strategy!r}")
    dev = kernels.require_cuda()
    x = torch.as_tensor(np.asarray(cv1, dtype=np.float64).reshape(-1)).to(dev)
    y = torch.as_tensor(np.asarray(cv2, dtype=np.float64).reshape(-1)).to(dev)
    if x.numel() != y.numel():
        raise ValueError("cv1 and cv2 must have the same length")
    if x.numel() == 0:
        raise ValueError("empty collective variables")
    w = None if weights is None else torch.as_tensor(np.asarray(weights, dtype=np.float64).reshape(-1)).to(dev)
    if periodic[0]:
        x = torch.remainder(x + 180.0, 360.0) - 180.0
    if periodic[1]:
        y = torch.remainder(y + 180.0, 360.0) - 180.0

    def axis_range(v, per, given):
        if given is not None:
            return float(given[0]), float(given[1])
        if per:
            return -180.0, 180.0
        fin = v[torch.isfinite(v)]
        if grid_strategy == "adaptive" and fin.numel() >= 100:
            lo = float(torch.quantile(fin, 0.01).item())
            hi = float(torch.quantile(fin, 0.99).item())
        else:
            lo, hi = float(fin.min().item()), float(fin.max().item())
        if not hi > lo:
            lo, hi = lo - 0.5, hi + 0.5
        return lo, hi

    rx = axis_range(x, periodic[0], None if ranges is None else ranges[0])
    ry = axis_range(y, periodic[1], None if ranges is None else ranges[1])
    bx, by = int(bins[0]), int(bins[1])
    H = kernels.hist2d(x.contiguous(), y.contiguous(), (bx, by), (rx, ry), w).cpu().numpy()
    xedges, yedges = np.linspace(rx[0], rx[1], bx + 1), np.linspace(ry[0], ry[1], by + 1)
    total = float(H.sum())
    if not total > 0:
        raise ValueError("no samples inside the histogram range")
    dx, dy = np.diff(xedges)[:, None], np.diff(yedges)[None, :]
    density = H / (total * dx * dy)
    mask = H < float(min_count)
    F = free_energy_from_density(density, temperature, mask=mask)
    return FESResult(F=F, xedges=xedges, yedges=yedges,
                     metadata={"counts": H, "temperature": float(temperature), "periodic": tuple(bool(p) for p in periodic),
                               "method": "histogram", "grid_strategy": grid_strategy, "min_count": int(min_count)})


# ----------------------------------------------------------------------------- transition path theory
@dataclass
class TPTResult:
    """conformations/results.py TPTResult fields filled by ``TPTAnalysis.analyze``."""

    source_states: np.ndarray
    sink_states: np.ndarray
    forward_committor: np.ndarray
    backward_committor: np.ndarray
    flux_matrix: np.ndarray
    net_flux: np.ndarray
    total_flux: float
    rate: float
    mfpt: float
    pathways: list
    pathway_fluxes: np.ndarray
    pathway_fractions: np.ndarray = field(default_factory=lambda: np.empty(0))


class TPTAnalysis:
    """``TPTAnalysis(T, pi).analyze(source, sink, n_paths, pathway_fraction)`` (tpt_analysis.py:35-260)."""

    def __init__(self, T: np.ndarray, pi: np.ndarray) -> None:
        self.T = np.asarray(T, dtype=float)
        self.pi = np.asarray(pi, dtype=float)
        if self.T.ndim != 2 or self.T.shape[0] != self.T.shape[1]:
            raise ValueError("Transition matrix must be square")
        self.n_states = int(self.T.shape[0])
        if len(self.pi) != self.n_states:
            raise ValueError("Stationary distribution length must match T dimensions")

    def _committor(self, P: torch.Tensor, A: np.ndarray, B: np.ndarray) -> torch.Tensor:
        """q = 0 on A, 1 on B, (I - P) q = 0 elsewhere: one dense solve on the intermediate states."""
        n = self.n_states
        dev = P.device
        q = torch.zeros((n,), dtype=torch.float64, device=dev)
        q[torch.from_numpy(B).to(dev)] = 1.0
        inter = np.setdiff1d(np.arange(n), np.concatenate([A, B]))
        if inter.size:
            I = torch.from_numpy(inter).to(dev)
            Bd = torch.from_numpy(B).to(dev)
            L = torch.eye(inter.size, dtype=torch.float64, device=dev) - P[I][:, I]
            rhs = P[I][:, Bd].sum(dim=1)
            q[I] = torch.linalg.solve(L, rhs)
        return q

    def analyze(self, source_states, sink_states, n_paths: int = 5, pathway_fraction: float = 0.99) -> TPTResult:
        A = np.unique(np.asarray(source_states)).astype(int)
        B = np.unique(np.asarray(sink_states)).astype(int)
        if np.intersect1d(A, B).size > 0:
            raise ValueError("Source and sink states must not overlap")
        if A.size == 0 or B.size == 0:
            raise ValueError("Source and sink states must not be empty")
        dev = kernels.require_cuda()
        T = torch.from_numpy(self.T).to(dev)
        pi = torch.from_numpy(self.pi).to(dev)
        qf = self._committor(T, A, B)
        # backward committor: forward committor of the time-reversed chain towards A
        pis = torch.clamp(pi, min=1e-300)
        Trev = (pi[None, :] * T.T) / pis[:, None]
        qb = self._committor(Trev, B, A)
        gross = pi[:, None] * qb[:, None] * T * qf[None, :]
        gross.fill_diagonal_(0.0)
        net = torch.clamp(gross - gross.T, min=0.0)
        Ad = torch.from_numpy(A).to(dev)
        notA = torch.ones((self.n_states,), dtype=torch.bool, device=dev)
        notA[Ad] = False
        total_flux = float(gross[Ad][:, notA].sum().item())
        pi_back = float((pi * qb).sum().item())
        rate = total_flux / pi_back if pi_back > 0 else float("nan")
        mfpt = 1.0 / rate if rate > 0 else float("inf")
        net_h = net.cpu().numpy()
        paths, caps = pathways(net_h, A, B, fraction=float(pathway_fraction), maxiter=10_000)
        paths, caps = paths[: int(n_paths)], np.asarray(caps[: int(n_paths)], dtype=float)
        fr = caps / total_flux if total_flux > 0 else np.zeros_like(caps)
        return TPTResult(A, B, qf.cpu().numpy(), qb.cpu().numpy(), gross.cpu().numpy(), net_h, total_flux, float(rate),
                         float(mfpt), [np.asarray(p, dtype=int) for p in paths], caps, fr)


def _widest_path(F: np.ndarray, A: Sequence[int], B: Sequence[int]):
    """Path from A to B maximising the minimal edge capacity (modified Dijkstra on the K x K net flux)."""
    import heapq

    n = F.shape[0]
    best = np.zeros(n)
    prev = -np.ones(n, dtype=int)
    heap = []
    for a in A:
        best[a] = np.inf
        heapq.heappush(heap, (-np.inf, int(a)))
    inB = np.zeros(n, dtype=bool)
    inB[list(B)] = True
    done = np.zeros(n, dtype=bool)
    while heap:
        negc, u = heapq.heappop(heap)
        if done[u]:
            continue
        done[u] = True
        if inB[u]:
            path = [u]
            while prev[path[-1]] >= 0:
                path.append(int(prev[path[-1]]))
            return path[::-1], -negc
        for v in np.flatnonzero(F[u] > 0):
            c = min(-negc, F[u, v])
            if c > best[v]:
                best[v] = c
                prev[v] = u
                heapq.heappush(heap, (-c, int(v)))
    return None, 0.0


def pathways(net_flux: np.ndarray, A, B, fraction: float = 1.0, maxiter: int = 1000):
    """Dominant reaction pathways: repeatedly take the widest A -> B path of the net flux, subtract its
    bottleneck capacity along the path, until ``fraction`` of the total flux is accounted for
    (deeptime ``ReactiveFlux.pathways``)."""
    F = np.array(net_flux, dtype=float)
    total = float(F[list(A)].sum() - F[np.ix_(list(A), list(A))].sum())
    paths, caps, acc = [], [], 0.0
    for _ in range(int(maxiter)):
        if total <= 0 or acc / total >= fraction:
            break
        p, c = _widest_path(F, A, B)
        if p is None or not c > 0:
            break
        for u, v in zip(p[:-1], p[1:]):
            F[u, v] -= c
        paths.append(p)
        caps.append(c)
        acc += c
    return paths, caps
