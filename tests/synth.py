"""Seeded synthetic inputs shared by the parity tests, smoke() and bench.py.

Nothing here reads /root/reference: the generators restate the recipes of
SURVEY.md section 8d (backbone chains with AR(1) internal motion; a 2-D
Mueller-Brown-like Langevin walk; label chains with metastable dwell times).
"""

from __future__ import annotations

import numpy as np

from pmarlo_b200.topology import Topology


def backbone_topology(n_res: int) -> Topology:
    names, resid = [], []
    for r in range(n_res):
        names += ["N", "CA", "C"]
        resid += [r, r, r]
    return Topology(names, np.asarray(resid, dtype=np.int64), np.zeros(len(names), dtype=np.int64))


def backbone_trajectories(n_res: int, n_traj: int, n_frames: int, seed: int, rho: float = 0.98,
                          sigma: float = 0.03) -> list[np.ndarray]:
    """n_traj arrays (n_frames, 3*n_res, 3) float32 nm: a random-coil base chain
    (0.15 nm bonds) plus an AR(1) perturbation per coordinate."""
    rng = np.random.default_rng(seed)
    A = 3 * n_res
    steps = rng.normal(size=(A, 3))
    steps /= np.linalg.norm(steps, axis=1, keepdims=True)
    base = np.cumsum(0.15 * steps, axis=0)
    out = []
    for _ in range(n_traj):
        noise = rng.normal(scale=sigma * np.sqrt(1 - rho * rho), size=(n_frames, A, 3))
        x = np.empty_like(noise)
        x[0] = rng.normal(scale=sigma, size=(A, 3))
        for t in range(1, n_frames):
            x[t] = rho * x[t - 1] + noise[t]
        out.append((base[None] + x).astype(np.float32))
    return out


def ar1_features(n_traj: int, n_frames: int, d: int, seed: int, mix: bool = True,
                 offset: float = 0.0) -> list[np.ndarray]:
    """Mixed AR(1) processes with per-dimension correlation times (float32)."""
    rng = np.random.default_rng(seed)
    rhos = 1.0 - np.geomspace(0.002, 0.5, d)
    M = rng.normal(size=(d, d)) / np.sqrt(d) + np.eye(d) if mix else np.eye(d)
    out = []
    for _ in range(n_traj):
        e = rng.normal(size=(n_frames, d)) * np.sqrt(1 - rhos ** 2)
        z = np.empty((n_frames, d))
        z[0] = rng.normal(size=d)
        for t in range(1, n_frames):
            z[t] = rhos * z[t - 1] + e[t]
        out.append((z @ M + offset).astype(np.float32))
    return out


def metastable_dtrajs(n_traj: int, n_frames: int, K: int, seed: int, stay: float = 0.9,
                      ragged: bool = True) -> list[np.ndarray]:
    """Label chains from a random reversible-ish K-state chain with dwell probability ``stay``."""
    rng = np.random.default_rng(seed)
    P = rng.random((K, K)) ** 3
    P = P + P.T
    np.fill_diagonal(P, 0.0)
    P = (1 - stay) * P / P.sum(axis=1, keepdims=True) + stay * np.eye(K)
    cdf = np.cumsum(P, axis=1)
    out = []
    for i in range(n_traj):
        n = n_frames - (i * 37 % max(1, n_frames // 3)) if ragged else n_frames
        s = np.empty(n, dtype=np.int32)
        s[0] = rng.integers(K)
        u = rng.random(n)
        for t in range(1, n):
            s[t] = min(K - 1, int(np.searchsorted(cdf[s[t - 1]], u[t])))
        out.append(s)
    return out


def muller_brown_trajectories(n_traj: int, n_frames: int, seed: int, dt: float = 1e-3, gamma: float = 5.0,
                              kT: float = 15.0, stride: int = 1) -> list[np.ndarray]:
    """Overdamped-ish Langevin walk on the Mueller-Brown surface (parameters of
    SURVEY.md section 8d), float32 (n_frames, 2)."""
    A = np.array([-200.0, -100.0, -170.0, 15.0])
    a = np.array([-1.0, -1.0, -6.5, 0.7])
    b = np.array([0.0, 0.0, 11.0, 0.6])
    c = np.array([-10.0, -10.0, -6.5, 0.7])
    x0 = np.array([1.0, 0.0, -0.5, -1.0])
    y0 = np.array([0.0, 0.5, 1.5, 1.0])

    def force(p):
        dx, dy = p[:, 0:1] - x0, p[:, 1:2] - y0
        e = A * np.exp(a * dx * dx + b * dx * dy + c * dy * dy)
        fx = -(e * (2 * a * dx + b * dy)).sum(axis=1)
        fy = -(e * (b * dx + 2 * c * dy)).sum(axis=1)
        f = np.stack([fx, fy], axis=1)
        lo, hi = np.array([-1.5, -0.5]), np.array([1.5, 2.5])
        f -= 1000.0 * np.minimum(p - lo, 0.0) + 1000.0 * np.maximum(p - hi, 0.0)
        return f

    rng = np.random.default_rng(seed)
    p = np.tile(np.array([[-0.55, 1.45]]), (n_traj, 1))
    v = np.zeros_like(p)
    c1 = np.exp(-gamma * dt)
    c2 = np.sqrt(kT * (1 - c1 * c1))
    out = np.empty((n_traj, n_frames, 2), dtype=np.float32)
    f = force(p)
    for t in range(n_frames * stride):
        v += 0.5 * dt * f
        p += 0.5 * dt * v
        v = c1 * v + c2 * rng.normal(size=p.shape)
        p += 0.5 * dt * v
        f = force(p)
        v += 0.5 * dt * f
        if t % stride == 0:
            out[:, t // stride] = p
    return [out[i] for i in range(n_traj)]


def _ar1(noise: np.ndarray, x0: np.ndarray, rho: float) -> np.ndarray:
    """x[t] = rho x[t-1] + noise[t], x[0] = x0 (vectorised over the trailing axes)."""
    from scipy.signal import lfilter

    e = noise.copy()
    e[0] = x0
    return lfilter([1.0], [1.0, -rho], e, axis=0)


def structure_trajectories(base_xyz: np.ndarray, n_traj: int, n_frames, seed: int, rho: float = 0.999,
                           sigma: float = 0.03) -> list[np.ndarray]:
    """Config C3 recipe: AR(1) perturbation (rho, sigma nm) of every Cartesian
    coordinate around a mean structure ``base_xyz`` (A,3) nm; float32 (n_frames, A, 3).
    ``n_frames`` may be a list (ragged trajectories)."""
    rng = np.random.default_rng(seed)
    base = np.asarray(base_xyz, dtype=np.float64)
    lens = [int(n_frames)] * n_traj if np.isscalar(n_frames) else [int(v) for v in n_frames]
    out = []
    for n in lens:
        noise = rng.normal(scale=sigma * np.sqrt(1 - rho * rho), size=(n,) + base.shape)
        x = _ar1(noise, rng.normal(scale=sigma, size=base.shape), rho)
        out.append((base[None] + x).astype(np.float32))
    return out


def _rotate_about(points: np.ndarray, origin: np.ndarray, axis: np.ndarray, angle: np.ndarray) -> np.ndarray:
    """Rodrigues rotation of points (n, m, 3) about per-frame axes (n, 3) through origins (n, 3)."""
    k = axis / np.linalg.norm(axis, axis=1, keepdims=True)
    v = points - origin[:, None, :]
    c, s = np.cos(angle)[:, None, None], np.sin(angle)[:, None, None]
    kx = np.cross(np.broadcast_to(k[:, None, :], v.shape), v)
    kd = np.sum(v * k[:, None, :], axis=2, keepdims=True)
    return origin[:, None, :] + v * c + kx * s + k[:, None, :] * kd * (1 - c)


def torus_walk(n_traj: int, n_frames, seed: int, dt: float = 0.05, kT: float = 1.0, barrier: float = 2.5):
    """Overdamped Langevin walk of (phi, psi) on the 4-well torus potential
    V = barrier * (cos 2 phi + cos 2 psi) / 2 (wells at +-pi/2); returns a list of (n, 2) arrays."""
    rng = np.random.default_rng(seed)
    lens = [int(n_frames)] * n_traj if np.isscalar(n_frames) else [int(v) for v in n_frames]
    out = []
    for n in lens:
        th = np.empty((n, 2))
        cur = rng.choice([-np.pi / 2, np.pi / 2], size=2) + rng.normal(scale=0.2, size=2)
        xi = rng.normal(size=(n, 2)) * np.sqrt(2 * kT * dt)
        for t in range(n):
            cur = cur + barrier * np.sin(2 * cur) * dt + xi[t]
            th[t] = cur
        out.append((th + np.pi) % (2 * np.pi) - np.pi)
    return out


def ala2_trajectories(top: dict, n_traj: int = 35, n_frames=None, seed: int = 1, jitter: float = 0.005):
    """Config C1 recipe: the alanine-dipeptide structure (22 atoms, nm) with its
    phi / psi driven by a seeded walk on a 4-well torus potential -- the atoms beyond the N-CA bond are
    rotated about it by (phi - phi0), those beyond CA-C about that axis by (psi - psi0) -- plus Gaussian
    jitter.  Default lengths: 35 trajectories of 371-372 frames = 13 000 frames.
    Returns (list of (n, 22, 3) float32, list of the driving (phi, psi) arrays)."""
    names = list(top["names"])
    resid = np.asarray(top["resid"])
    base = np.asarray(top["xyz"], dtype=np.float64)
    if n_frames is None:
        n_frames = [372 if i < 13000 - 371 * n_traj else 371 for i in range(n_traj)]

    def idx(name, res):
        return [i for i, (a, r) in enumerate(zip(names, resid)) if a == name and r == res][0]

    c0, n1, ca, c1, n2 = idx("C", 0), idx("N", 1), idx("CA", 1), idx("C", 1), idx("N", 2)

    def dihedral(p0, p1, p2, p3):
        b1, b2, b3 = p1 - p0, p2 - p1, p3 - p2
        cc1, cc2 = np.cross(b2, b3), np.cross(b1, b2)
        return np.arctan2(np.dot(b1, cc1) * np.linalg.norm(b2), np.dot(cc1, cc2))

    phi0 = dihedral(base[c0], base[n1], base[ca], base[c1])
    psi0 = dihedral(base[n1], base[ca], base[c1], base[n2])
    # atoms that move with phi: everything bonded beyond CA (HA, CB, HB*, C, O, NME); with psi: O(ALA) + NME
    s_phi = np.array([i for i in range(len(names)) if (resid[i] == 1 and names[i] not in ("N", "H", "CA")) or resid[i] == 2])
    s_psi = np.array([i for i in range(len(names)) if (resid[i] == 1 and names[i] == "O") or resid[i] == 2])
    walks = torus_walk(n_traj, n_frames, seed)
    rng = np.random.default_rng(seed + 7919)
    out = []
    for th in walks:
        n = th.shape[0]
        x = np.broadcast_to(base, (n,) + base.shape).copy()
        x[:, s_phi] = _rotate_about(x[:, s_phi], x[:, ca], x[:, ca] - x[:, n1], th[:, 0] - phi0)
        x[:, s_psi] = _rotate_about(x[:, s_psi], x[:, c1], x[:, c1] - x[:, ca], th[:, 1] - psi0)
        x += rng.normal(scale=jitter, size=x.shape)
        out.append(x.astype(np.float32))
    return out, walks
