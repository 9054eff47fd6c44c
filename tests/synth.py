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
