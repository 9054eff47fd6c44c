"""numpy statements of the libpmb200 kernel CONTRACTS (test infrastructure).

They let the host-side orchestration (TICA moment algebra with a shared shift,
Lloyd loop, count/MLE chain, all-reduce of partials across ranks) run on CPU
tensors under ``gloo`` with world_size 2.  Each function mirrors the signature of
the wrapper of the same name in ``pmarlo_b200/kernels.py`` and the semantics
documented in ``include/pmb200.h``; the heavy lifting is delegated to the oracle.
"""

from __future__ import annotations

import numpy as np
import torch

import oracle


def _np(t):
    return t.detach().cpu().numpy()


def pair_mask(seg_offsets, n, lag):
    off = _np(seg_offsets)
    m = np.zeros(int(n), dtype=np.uint8)
    for s, e in zip(off[:-1], off[1:]):
        L = int(e - s)
        t = np.arange(L)
        m[s:e] |= (t + lag < L).astype(np.uint8)
        m[s:e] |= ((t >= lag).astype(np.uint8) << 1)
    return torch.from_numpy(m)


def col_moments(X, mask=None, shift=None):
    x = _np(X).astype(np.float64)
    n, d = x.shape
    sh = _np(shift).astype(np.float64) if shift is not None else np.nan_to_num(x[0], nan=0.0)
    ok = ~np.isnan(x)
    dx = np.where(ok, x - sh, 0.0)
    e = np.zeros(n) if mask is None else 2.0 - np.array([bin(int(v) & 3).count("1") for v in _np(mask)])
    out = np.stack([ok.sum(0).astype(np.float64), sh, dx.sum(0), (dx * dx).sum(0),
                    (e[:, None] * dx).sum(0), (e[:, None] * ok).sum(0)])
    return torch.from_numpy(out)


def scaler_from_moments(moments, n, semantic, with_std):
    mom = _np(moments)
    nv, sh, s1, s2 = mom[0], mom[1], mom[2], mom[3]
    m = np.where(nv > 0, sh + s1 / np.maximum(nv, 1), 0.0)
    ss = np.maximum(np.where(nv > 0, s2 - s1 * s1 / np.maximum(nv, 1), 0.0), 0.0)
    var = ss / n
    eps = np.finfo(np.float64).eps
    sd = np.sqrt(var)
    const = (var <= n * eps * var + (n * m * eps) ** 2) | (sd < 10 * eps)
    sd_safe = np.where(const, 1.0, sd)
    stats = np.stack([m if semantic else np.zeros_like(m),
                      sd_safe if (semantic and with_std) else np.ones_like(m),
                      np.sqrt(ss / max(n - 1, 1))])
    cond = np.stack([m.astype(np.float32), (1.0 / sd_safe).astype(np.float32)])
    return torch.from_numpy(stats), torch.from_numpy(cond)


def gram(X, mask, lag, mode, cond, impl=0, out=None):
    x = _np(X).astype(np.float32)
    c = _np(cond)
    z = np.where(np.isnan(x), np.float32(0), (x - c[0]) * c[1]).astype(np.float64)
    mk = _np(mask)
    if mode == 0:
        w = np.array([bin(int(v) & 3).count("1") for v in mk], dtype=np.float64)
        G = (z * w[:, None]).T @ z
    else:
        idx = np.flatnonzero(mk & 1)
        v = z[idx] - z[idx + lag]
        G = v.T @ v
    G = torch.from_numpy(G)
    if out is not None:
        out.copy_(G)
        return out
    return G


def tica_covariances(G0, G1, moments, stats, cond, n, n_pairs, semantic=1):
    mom, st, c = _np(moments), _np(stats), _np(cond).astype(np.float64)
    nv, sh, s1, es1, ecnt = mom[0], mom[1], mom[2], mom[4], mom[5]
    m_imp = np.where(nv > 0, sh + s1 / np.maximum(nv, 1), 0.0)
    twoT = 2.0 * n_pairs
    e_total = 2.0 * n - twoT
    E = sh * ecnt + es1 + m_imp * (e_total - ecnt)
    Sw = 2.0 * n * m_imp - E
    alpha = 1.0 / (c[1] * st[1])
    beta = (c[0] - st[0]) / st[1]
    mu = (Sw / twoT - st[0]) / st[1]
    S = twoT * (mu - beta) / alpha
    g0 = (np.outer(alpha, alpha) * _np(G0) + np.outer(alpha * S, beta) + np.outer(beta, alpha * S)
          + np.outer(beta, beta) * twoT)
    g1 = np.outer(alpha, alpha) * _np(G1)
    C00 = g0 / twoT - np.outer(mu, mu)
    C0t = (g0 - g1) / twoT - np.outer(mu, mu)
    return torch.from_numpy(C00), torch.from_numpy(C0t), torch.from_numpy(mu)


def tica_solve(C00, C0t, eps=1e-6):
    lam, R, rank = oracle.tica.eig_corr(_np(C00), _np(C0t), eps)
    d = C00.shape[0]
    ev = np.zeros(d)
    ev[:rank] = lam
    Rf = np.zeros((d, d))
    Rf[:, :rank] = R
    return torch.from_numpy(ev), torch.from_numpy(Rf), torch.tensor([rank], dtype=torch.int32)


def tica_finalize(evals, evecs, moments, stats, mu, m, kinetic_map=True):
    ev, R, mom, st = _np(evals), _np(evecs), _np(moments), _np(stats)
    a = st[0] + st[1] * _np(mu)
    nanfill = np.where(mom[0] > 0, mom[1] + mom[2] / np.maximum(mom[0], 1), 0.0)
    W = R[:, :m] * (ev[:m] if kinetic_map else 1.0) / st[1][:, None]
    return torch.from_numpy(a), torch.from_numpy(nanfill), torch.from_numpy(np.ascontiguousarray(W))


def project(X, a, nanfill, W, out_f64=False, out=None):
    x = _np(X).astype(np.float64)
    x = np.where(np.isnan(x), _np(nanfill)[None, :], x)
    y = (x - _np(a)) @ _np(W)
    return torch.from_numpy(y if out_f64 else y.astype(np.float32))


def kmeans_assign(Y, centers, labels=None, sums=None, counts=None, inertia=None, n_rechecked=None, impl=0, hints=None):
    y = _np(Y).astype(np.float64)
    lab, dmin = oracle.kmeans.assign(y, _np(centers))
    if labels is None:
        labels = torch.empty((y.shape[0],), dtype=torch.int32)
    labels.copy_(torch.from_numpy(lab.astype(np.int32)))
    if sums is not None:
        s = np.zeros(tuple(sums.shape))
        np.add.at(s, lab, y)
        sums += torch.from_numpy(s)
        counts += torch.from_numpy(np.bincount(lab, minlength=counts.numel()))
    if inertia is not None:
        inertia += float(dmin.sum())
    return labels


def kmeans_update(centers, sums, counts, shift2=None):
    nz = counts > 0
    centers[nz] = sums[nz] / counts[nz].to(torch.float64)[:, None]


def count_lagged(labels, seg_offsets, K, lag, step=1, out=None):
    off = _np(seg_offsets)
    lab = _np(labels)
    C = oracle.counts.count_lagged([lab[s:e] for s, e in zip(off[:-1], off[1:])], K, lag, step=step)
    if out is None:
        return torch.from_numpy(C)
    out += torch.from_numpy(C)
    return out


def relabel_compact(labels, seg_offsets, lut):
    off, lab, table = _np(seg_offsets), _np(labels).astype(np.int64), _np(lut).astype(np.int64)
    ok = (lab >= 0) & (lab < table.size)
    m = np.where(ok, table[np.clip(lab, 0, table.size - 1)], -1)
    keep = m >= 0
    before = np.concatenate([[0], np.cumsum(keep)])
    out = np.zeros(max(lab.size, 1), dtype=np.int32)
    out[: int(keep.sum())] = m[keep]
    return torch.from_numpy(out), torch.from_numpy(before[off].astype(np.int64))


def counts_active(C, eps=1e-12):
    c = _np(C).astype(np.float64)
    act = ((c.sum(0) + c.sum(1)) > eps).astype(np.uint8)
    return torch.from_numpy(c), torch.from_numpy(act)


def mle_rev(C, active=None, alpha=0.0, maxerr=1e-8, maxiter=1_000_000):
    c = _np(C)
    K = c.shape[0]
    idx = np.arange(K) if active is None else np.flatnonzero(_np(active))
    T, pi, it = oracle.msm.mle_rev(c[np.ix_(idx, idx)] + alpha, maxerr, maxiter)
    Tf, pif = oracle.msm.expand_results(K, idx, T, pi)
    return torch.from_numpy(Tf), torch.from_numpy(pif), torch.tensor([it, 1], dtype=torch.int64)


def eig_rev_topk(T, pi, k, max_steps=0):
    t, p = _np(T), _np(pi)
    idx = np.flatnonzero(p > 0)
    ev = oracle.msm.eigenvalues_rev(t[np.ix_(idx, idx)], p[idx], None)
    out = np.zeros(k)
    out[: min(k, ev.size)] = ev[:k]
    return torch.from_numpy(out), torch.tensor([0, 1], dtype=torch.int64)


def install(monkeypatch_or_module):
    """Replace every wrapper of pmarlo_b200.kernels by its numpy contract."""
    from pmarlo_b200 import kernels

    names = ["pair_mask", "col_moments", "scaler_from_moments", "gram", "tica_covariances", "tica_solve",
             "tica_finalize", "project", "kmeans_assign", "kmeans_update", "count_lagged", "relabel_compact",
             "counts_active",
             "mle_rev", "eig_rev_topk"]
    for n in names:
        if hasattr(monkeypatch_or_module, "setattr"):
            monkeypatch_or_module.setattr(kernels, n, globals()[n])
        else:
            setattr(kernels, n, globals()[n])
