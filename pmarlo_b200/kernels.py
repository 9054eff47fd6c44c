"""Typed host-side wrappers of the libpmb200 C ABI (one function per entry point).

Every argument that is a tensor must live on the CUDA device; outputs are
allocated here with torch (device memory plumbing only) and the call is
enqueued on torch's current stream.  No arithmetic happens in this module.
"""

from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_handle

__all__ = [
    "require_cuda", "featurize", "pair_mask", "col_moments", "scaler_from_moments", "gram",
    "tica_covariances", "tica_solve", "tica_finalize", "sym_eigvals_batched", "project", "kmeans_assign", "kmeans_tc_scores",
    "kmeans_update", "count_lagged", "count_lagged_weighted", "counts_active", "trig_expand", "mle_rev", "eig_rev_topk",
]


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.Pmb200Error(
            "pmarlo_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback."
        )
    return torch.device("cuda", torch.cuda.current_device())


def _dev(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor")
    if t.dtype != dtype:
        raise TypeError(f"{name} must have dtype {dtype}, got {t.dtype}")
    return t


def _flat(t, dtype, name: str, numel: int | None = None, shape: tuple | None = None) -> torch.Tensor:
    """A contiguous device tensor of the stated dtype (and size): the C ABI sees raw pointers, so a
    strided view or a short output would read or write the wrong memory instead of failing."""
    _dev(t, dtype, name)
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    if numel is not None and int(t.numel()) != int(numel):
        raise ValueError(f"{name} must have {numel} elements, got {t.numel()}")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(t.shape)}")
    return t


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _rowmajor(t: torch.Tensor, name: str) -> int:
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name} must be a row-major 2-D tensor")
    return int(t.stride(0)) if t.shape[0] > 1 else max(int(t.shape[1]), int(t.stride(0)))


def featurize(xyz: torch.Tensor, units: torch.Tensor, n_cols: int, out: torch.Tensor | None = None):
    """K1.  xyz (N,A,3) float32, units (n_units,8) int32 -> out (N,n_cols) float32."""
    _dev(xyz, torch.float32, "xyz")
    _dev(units, torch.int32, "units")
    if xyz.dim() != 3 or xyz.shape[2] != 3 or not xyz.is_contiguous():
        raise ValueError("xyz must be a contiguous (n_frames, n_atoms, 3) tensor")
    n, a = int(xyz.shape[0]), int(xyz.shape[1])
    units = units.contiguous()
    if out is None:
        out = torch.empty((n, n_cols), dtype=torch.float32, device=xyz.device)
    ld = _rowmajor(out, "out") if n_cols else n_cols
    check(_lib.lib().pmb_featurize(ptr(xyz), n, a, ptr(units), int(units.shape[0]), int(n_cols),
                                   ptr(out), ld, stream_handle(xyz.device)), "pmb_featurize")
    return out


def pair_mask(seg_offsets: torch.Tensor, n: int, lag: int) -> torch.Tensor:
    _dev(seg_offsets, torch.int64, "seg_offsets")
    mask = torch.empty((int(n),), dtype=torch.uint8, device=seg_offsets.device)
    check(_lib.lib().pmb_pair_mask(ptr(seg_offsets), int(seg_offsets.numel()) - 1, int(n), int(lag),
                                   ptr(mask), stream_handle(mask.device)), "pmb_pair_mask")
    return mask


def col_moments(X: torch.Tensor, mask: torch.Tensor | None = None,
                shift: torch.Tensor | None = None) -> torch.Tensor:
    """K2.  Returns (6,d) float64: n_valid, shift, S1, S2, edge S1, edge count."""
    _dev(X, torch.float32, "X")
    ld = _rowmajor(X, "X")
    n, d = int(X.shape[0]), int(X.shape[1])
    if mask is not None:
        _dev(mask, torch.uint8, "mask")
    if shift is not None:
        _dev(shift, torch.float64, "shift")
    out = torch.empty((6, d), dtype=torch.float64, device=X.device)
    L = _lib.lib()
    ws = _ws(L.pmb_col_moments_ws_bytes(d), X.device)
    check(L.pmb_col_moments(ptr(X), n, d, ld, ptr(mask), ptr(shift), ptr(out), ptr(ws), ws.numel(),
                            stream_handle(X.device)), "pmb_col_moments")
    return out


def scaler_from_moments(moments: torch.Tensor, n: int, semantic: int, with_std: int):
    _dev(moments, torch.float64, "moments")
    d = int(moments.shape[1])
    stats = torch.empty((3, d), dtype=torch.float64, device=moments.device)
    cond = torch.empty((2, d), dtype=torch.float32, device=moments.device)
    check(_lib.lib().pmb_scaler_from_moments(ptr(moments), int(n), d, int(semantic), int(with_std),
                                             ptr(stats), ptr(cond), stream_handle(moments.device)),
          "pmb_scaler_from_moments")
    return stats, cond


def gram(X: torch.Tensor, mask: torch.Tensor, lag: int, mode: int, cond: torch.Tensor,
         impl: int = 0, out: torch.Tensor | None = None) -> torch.Tensor:
    """K3.  (d,d) float64 weighted Gram matrix, mode 0 or 1 (see pmb200.h)."""
    _dev(X, torch.float32, "X")
    _dev(mask, torch.uint8, "mask")
    _dev(cond, torch.float32, "cond")
    ld = _rowmajor(X, "X")
    n, d = int(X.shape[0]), int(X.shape[1])
    if out is None:
        out = torch.empty((d, d), dtype=torch.float64, device=X.device)
    L = _lib.lib()
    ws = _ws(L.pmb_gram_ws_bytes(d), X.device)
    check(L.pmb_gram(ptr(X), n, d, ld, ptr(mask), int(lag), int(mode), cond[0].data_ptr(),
                     cond[1].data_ptr(), ptr(out), ptr(ws), ws.numel(), int(impl),
                     stream_handle(X.device)), "pmb_gram")
    return out


def tica_covariances(G0, G1, moments, stats, cond, n: int, n_pairs: int, semantic: int = 1):
    d = int(G0.shape[0])
    dev = G0.device
    C00 = torch.empty((d, d), dtype=torch.float64, device=dev)
    C0t = torch.empty((d, d), dtype=torch.float64, device=dev)
    mu = torch.empty((d,), dtype=torch.float64, device=dev)
    check(_lib.lib().pmb_tica_covariances(ptr(G0), ptr(G1), ptr(moments), ptr(stats), ptr(cond),
                                          int(n), int(n_pairs), d, int(semantic), ptr(C00), ptr(C0t),
                                          ptr(mu), stream_handle(dev)), "pmb_tica_covariances")
    return C00, C0t, mu


def tica_solve(C00: torch.Tensor, C0t: torch.Tensor, eps: float = 1e-6):
    """K4.  Returns (evals (d,), evecs (d,d) unscaled, rank int32[1]) on the device."""
    _dev(C00, torch.float64, "C00")
    _dev(C0t, torch.float64, "C0t")
    d = int(C00.shape[0])
    dev = C00.device
    evals = torch.empty((d,), dtype=torch.float64, device=dev)
    evecs = torch.empty((d, d), dtype=torch.float64, device=dev)
    rank = torch.zeros((4,), dtype=torch.int32, device=dev)
    L = _lib.lib()
    ws = _ws(L.pmb_tica_solve_ws_bytes(d), dev)
    check(L.pmb_tica_solve(ptr(C00.contiguous()), ptr(C0t.contiguous()), d, float(eps), ptr(evals),
                           ptr(evecs), ptr(rank), ptr(ws), ws.numel(), stream_handle(dev)),
          "pmb_tica_solve")
    return evals, evecs, rank


def tica_finalize(evals, evecs, moments, stats, mu, m: int, kinetic_map: bool = True):
    """Projection operands (a, nanfill, W) of the fitted TICA model, all fp64 on the device."""
    d = int(evecs.shape[0])
    dev = evecs.device
    a = torch.empty((d,), dtype=torch.float64, device=dev)
    nanfill = torch.empty((d,), dtype=torch.float64, device=dev)
    W = torch.empty((d, int(m)), dtype=torch.float64, device=dev)
    check(_lib.lib().pmb_tica_finalize(ptr(evals), ptr(evecs), ptr(moments), ptr(stats), ptr(mu), d,
                                       int(m), 1 if kinetic_map else 0, ptr(a), ptr(nanfill), ptr(W),
                                       stream_handle(dev)), "pmb_tica_finalize")
    return a, nanfill, W


def sym_eigvals_batched(A: torch.Tensor) -> torch.Tensor:
    """Eigenvalues (sorted by magnitude, descending) of a batch of symmetric matrices."""
    _dev(A, torch.float64, "A")
    if A.dim() == 2:
        A = A.unsqueeze(0)
    A = A.contiguous().clone()
    b, n = int(A.shape[0]), int(A.shape[1])
    evals = torch.empty((b, n), dtype=torch.float64, device=A.device)
    L = _lib.lib()
    ws = _ws(L.pmb_sym_eigvals_ws_bytes(n, b), A.device)
    check(L.pmb_sym_eigvals_batched(ptr(A), n, b, ptr(evals), ptr(ws), ws.numel(),
                                    stream_handle(A.device)), "pmb_sym_eigvals_batched")
    return evals


def project(X: torch.Tensor, a: torch.Tensor, nanfill: torch.Tensor, W: torch.Tensor,
            out_f64: bool = False, out: torch.Tensor | None = None) -> torch.Tensor:
    """K5.  Y = (impute(X) - a) @ W with fp64 accumulation."""
    _dev(X, torch.float32, "X")
    _dev(a, torch.float64, "a")
    _dev(nanfill, torch.float64, "nanfill")
    _dev(W, torch.float64, "W")
    ld = _rowmajor(X, "X")
    n, d = int(X.shape[0]), int(X.shape[1])
    W = W.contiguous()
    m = int(W.shape[1])
    if out is None:
        out = torch.empty((n, m), dtype=torch.float64 if out_f64 else torch.float32, device=X.device)
    ldy = _rowmajor(out, "out")
    check(_lib.lib().pmb_project(ptr(X), n, d, ld, ptr(a), ptr(nanfill), ptr(W), m, ptr(out), ldy,
                                 1 if out.dtype == torch.float64 else 0, stream_handle(X.device)),
          "pmb_project")
    return out


def kmeans_assign(Y: torch.Tensor, centers: torch.Tensor, labels: torch.Tensor | None = None,
                  sums: torch.Tensor | None = None, counts: torch.Tensor | None = None,
                  inertia: torch.Tensor | None = None, n_rechecked: torch.Tensor | None = None,
                  impl: int = 0, hints: torch.Tensor | None = None):
    """K6.  labels int32 (n,), optional accumulation into sums/counts/inertia.
    impl: 0 auto, 1 SIMT kernel, 2 tcgen05 score GEMM with fused argmin.
    hints: int32 labels of a previous assignment (may be ``labels`` itself); results never depend on them."""
    if Y.dtype not in (torch.float32, torch.float64) or not Y.is_cuda:
        raise TypeError("Y must be a float32/float64 CUDA tensor")
    _dev(centers, torch.float64, "centers")
    ld = _rowmajor(Y, "Y")
    n, D = int(Y.shape[0]), int(Y.shape[1])
    centers = centers.contiguous()
    K = int(centers.shape[0])
    if int(centers.shape[1]) != D:
        raise ValueError("centers and Y disagree on the feature dimension")
    if labels is None:
        labels = torch.empty((n,), dtype=torch.int32, device=Y.device)
    _flat(labels, torch.int32, "labels", numel=n)
    if sums is not None:
        _flat(sums, torch.float64, "sums", numel=K * D)
    if counts is not None:
        _flat(counts, torch.int64, "counts", numel=K)
    if inertia is not None:
        _flat(inertia, torch.float64, "inertia", numel=1)
    if n_rechecked is not None:
        _flat(n_rechecked, torch.int64, "n_rechecked", numel=1)
    if hints is not None:
        _dev(hints, torch.int32, "hints")
        if hints.numel() != n or not hints.is_contiguous():
            raise ValueError("hints must be a contiguous int32 tensor with one entry per frame")
    L = _lib.lib()
    ws = _ws(L.pmb_kmeans_assign_ws_bytes(n, D, K), Y.device) if Y.dtype == torch.float32 else None
    check(L.pmb_kmeans_assign(ptr(Y), 1 if Y.dtype == torch.float64 else 0, n, D, ld,
                              ptr(centers), K, ptr(labels), ptr(sums), ptr(counts),
                              ptr(inertia), ptr(n_rechecked), ptr(hints), ptr(ws),
                              ws.numel() if ws is not None else 0,
                              int(impl), stream_handle(Y.device)),
          "pmb_kmeans_assign")
    return labels


def kmeans_tc_scores(Y: torch.Tensor, centers: torch.Tensor):
    """Test hook: (labels, scores (n, Kpad) float32) of the tcgen05 assignment path."""
    _dev(Y, torch.float32, "Y")
    _dev(centers, torch.float64, "centers")
    ld = _rowmajor(Y, "Y")
    n, D = int(Y.shape[0]), int(Y.shape[1])
    centers = centers.contiguous()
    K = int(centers.shape[0])
    kpad = (K + 255) // 256 * 256
    labels = torch.empty((n,), dtype=torch.int32, device=Y.device)
    scores = torch.empty((n, kpad), dtype=torch.float32, device=Y.device)
    L = _lib.lib()
    ws = _ws(L.pmb_kmeans_assign_ws_bytes(n, D, K), Y.device)
    check(L.pmb_kmeans_tc_scores(ptr(Y), n, D, ld, ptr(centers), K, ptr(labels), ptr(scores), ptr(ws),
                                 ws.numel(), stream_handle(Y.device)), "pmb_kmeans_tc_scores")
    return labels, scores


def kmeans_update(centers: torch.Tensor, sums: torch.Tensor, counts: torch.Tensor,
                  shift2: torch.Tensor | None = None) -> None:
    _dev(centers, torch.float64, "centers")
    _dev(sums, torch.float64, "sums")
    _dev(counts, torch.int64, "counts")
    K, D = int(centers.shape[0]), int(centers.shape[1])
    check(_lib.lib().pmb_kmeans_update(ptr(centers), ptr(sums), ptr(counts), K, D, ptr(shift2),
                                       stream_handle(centers.device)), "pmb_kmeans_update")


def hist2d(x: torch.Tensor, y: torch.Tensor, bins: tuple[int, int], ranges, weights: torch.Tensor | None = None) -> torch.Tensor:
    """np.histogram2d(x, y, bins, range, weights=...)[0] on the device (fp64)."""
    _flat(x, torch.float64, "x")
    _flat(y, torch.float64, "y", numel=x.numel())
    if weights is not None:
        _flat(weights, torch.float64, "weights", numel=x.numel())
    bx, by = int(bins[0]), int(bins[1])
    H = torch.zeros((bx, by), dtype=torch.float64, device=x.device)
    (xlo, xhi), (ylo, yhi) = ranges
    check(_lib.lib().pmb_hist2d(ptr(x), ptr(y), ptr(weights), int(x.numel()), float(xlo), float(xhi), bx, float(ylo),
                                float(yhi), by, ptr(H), stream_handle(x.device)), "pmb_hist2d")
    return H


def silhouette_samples(Y: torch.Tensor, labels: torch.Tensor, K: int) -> torch.Tensor:
    """Per-sample silhouette coefficients (fp64) of a labelling with K <= 64 clusters."""
    _dev(Y, torch.float64, "Y")
    if Y.dim() != 2 or not Y.is_contiguous():
        raise ValueError("Y must be a contiguous (n, D) tensor")
    n, D = int(Y.shape[0]), int(Y.shape[1])
    _flat(labels, torch.int32, "labels", numel=n)
    sizes = torch.bincount(labels.to(torch.int64), minlength=int(K)).to(torch.int64).contiguous()
    out = torch.empty((n,), dtype=torch.float64, device=Y.device)
    check(_lib.lib().pmb_silhouette_samples(ptr(Y), n, D, ptr(labels), int(K), ptr(sizes), ptr(out),
                                            stream_handle(Y.device)), "pmb_silhouette_samples")
    return out


def count_lagged(labels: torch.Tensor, seg_offsets: torch.Tensor, K: int, lag: int, step: int = 1,
                 out: torch.Tensor | None = None) -> torch.Tensor:
    """K7.  Accumulates into ``out`` (K,K) int64 (zero-initialised when None)."""
    _flat(labels, torch.int32, "labels")
    _flat(seg_offsets, torch.int64, "seg_offsets")
    if out is None:
        out = torch.zeros((K, K), dtype=torch.int64, device=labels.device)
    _flat(out, torch.int64, "out", shape=(K, K))
    check(_lib.lib().pmb_count_lagged(ptr(labels), int(labels.numel()), ptr(seg_offsets),
                                      int(seg_offsets.numel()) - 1, int(K), int(lag), int(step),
                                      ptr(out), stream_handle(labels.device)), "pmb_count_lagged")
    return out


def count_lagged_weighted(labels, weights, seg_offsets, K: int, lag: int, step: int = 1, out=None):
    _flat(labels, torch.int32, "labels")
    _flat(weights, torch.float64, "weights", numel=labels.numel())
    _flat(seg_offsets, torch.int64, "seg_offsets")
    if out is None:
        out = torch.zeros((K, K), dtype=torch.float64, device=labels.device)
    _flat(out, torch.float64, "out", shape=(K, K))
    check(_lib.lib().pmb_count_lagged_weighted(ptr(labels), ptr(weights), int(labels.numel()),
                                               ptr(seg_offsets), int(seg_offsets.numel()) - 1, int(K),
                                               int(lag), int(step), ptr(out),
                                               stream_handle(labels.device)), "pmb_count_lagged_weighted")
    return out


def relabel_compact(labels: torch.Tensor, seg_offsets: torch.Tensor, lut: torch.Tensor):
    """``[lut[s] for s in traj if lut[s] >= 0]`` for every shard (ck_runner.py:150-153): returns the
    shortened label shard (int32, device) and its offsets (int64, device, length n_seg + 1)."""
    _flat(labels, torch.int32, "labels")
    _flat(seg_offsets, torch.int64, "seg_offsets")
    _flat(lut, torch.int32, "lut")
    n, n_seg = int(labels.numel()), int(seg_offsets.numel()) - 1
    out = torch.empty((max(n, 1),), dtype=torch.int32, device=labels.device)
    new_off = torch.empty((n_seg + 1,), dtype=torch.int64, device=labels.device)
    nbytes = int(_lib.lib().pmb_relabel_compact_ws_bytes(n))
    ws = _ws(nbytes, labels.device)
    check(_lib.lib().pmb_relabel_compact(ptr(labels), n, ptr(seg_offsets), n_seg, ptr(lut), int(lut.numel()),
                                         ptr(out), ptr(new_off), ptr(ws), nbytes,
                                         stream_handle(labels.device)), "pmb_relabel_compact")
    return out, new_off


def counts_active(C: torch.Tensor, eps: float = 1e-12):
    """int64 counts -> (fp64 counts, active mask uint8)."""
    _dev(C, torch.int64, "C")
    C = C.contiguous()
    K = int(C.shape[0])
    Cf = torch.empty((K, K), dtype=torch.float64, device=C.device)
    active = torch.empty((K,), dtype=torch.uint8, device=C.device)
    check(_lib.lib().pmb_counts_active(ptr(C), K, float(eps), ptr(Cf), ptr(active),
                                       stream_handle(C.device)), "pmb_counts_active")
    return Cf, active


def trig_expand(X: torch.Tensor, periodic: torch.Tensor, out_col: torch.Tensor, Fe: int) -> torch.Tensor:
    _dev(X, torch.float64, "X")
    _dev(periodic, torch.uint8, "periodic")
    _dev(out_col, torch.int32, "out_col")
    X = X.contiguous()
    n, F = int(X.shape[0]), int(X.shape[1])
    Xe = torch.empty((n, int(Fe)), dtype=torch.float64, device=X.device)
    check(_lib.lib().pmb_trig_expand(ptr(X), n, F, ptr(periodic), ptr(out_col), ptr(Xe), int(Fe),
                                     stream_handle(X.device)), "pmb_trig_expand")
    return Xe


def mle_rev(C: torch.Tensor, active: torch.Tensor | None = None, alpha: float = 0.0,
            maxerr: float = 1e-8, maxiter: int = 1_000_000):
    """K8.  C (K,K) or (B,K,K) float64 -> (T, pi, info[B,2])."""
    _dev(C, torch.float64, "C")
    squeeze = C.dim() == 2
    Cb = (C.unsqueeze(0) if squeeze else C).contiguous()
    B, K = int(Cb.shape[0]), int(Cb.shape[1])
    if active is not None:
        _dev(active, torch.uint8, "active")
        active = active.reshape(B, K).contiguous()
    dev = C.device
    T = torch.empty((B, K, K), dtype=torch.float64, device=dev)
    pi = torch.empty((B, K), dtype=torch.float64, device=dev)
    info = torch.zeros((B, 2), dtype=torch.int64, device=dev)
    L = _lib.lib()
    ws = _ws(L.pmb_mle_rev_ws_bytes(K, B), dev)
    check(L.pmb_mle_rev(ptr(Cb), ptr(active), K, B, float(alpha), float(maxerr), int(maxiter), ptr(T), ptr(pi),
                        ptr(info), ptr(ws), ws.numel(), stream_handle(dev)), "pmb_mle_rev")
    if squeeze:
        return T[0], pi[0], info[0]
    return T, pi, info


def bayes_rev_sample(C: torch.Tensor, T_mle: torch.Tensor, pi: torch.Tensor, n_samples: int, n_steps: int | None = None,
                     seed: int = 0) -> torch.Tensor:
    """Reversible transition-matrix samples (B, n_samples, K, K) and their stationary vectors (B, n_samples, K)
    from counts C (B,K,K) fp64, started at the
    reversible MLE (T_mle, pi).  n_steps Gibbs sweeps per sample (default sqrt(K), like deeptime)."""
    _dev(C, torch.float64, "C")
    _dev(T_mle, torch.float64, "T_mle")
    _dev(pi, torch.float64, "pi")
    Cb = (C.unsqueeze(0) if C.dim() == 2 else C).contiguous()
    B, K = int(Cb.shape[0]), int(Cb.shape[1])
    X = (pi.reshape(B, K, 1) * T_mle.reshape(B, K, K)).contiguous()
    if n_steps is None:
        n_steps = max(1, int(np.sqrt(K)))
    out = torch.empty((B, int(n_samples), K, K), dtype=torch.float64, device=C.device)
    pis = torch.empty((B, int(n_samples), K), dtype=torch.float64, device=C.device)
    check(_lib.lib().pmb_bayes_rev_sample(ptr(Cb), ptr(X), K, B, int(n_samples), int(n_steps), int(seed) & (2**64 - 1),
                                          ptr(out), ptr(pis), stream_handle(C.device)), "pmb_bayes_rev_sample")
    return out, pis


def eig_rev_topk(T: torch.Tensor, pi: torch.Tensor, k: int, max_steps: int = 0):
    """K9.  Leading k eigenvalues (by magnitude) of reversible T; batched like mle_rev."""
    _dev(T, torch.float64, "T")
    _dev(pi, torch.float64, "pi")
    squeeze = T.dim() == 2
    Tb = (T.unsqueeze(0) if squeeze else T).contiguous()
    B, K = int(Tb.shape[0]), int(Tb.shape[1])
    pib = pi.reshape(B, K).contiguous()
    dev = T.device
    evals = torch.empty((B, int(k)), dtype=torch.float64, device=dev)
    info = torch.zeros((B, 2), dtype=torch.int64, device=dev)
    L = _lib.lib()
    ws = _ws(L.pmb_eig_rev_topk_ws_bytes(K, int(k), B, int(max_steps)), dev)
    check(L.pmb_eig_rev_topk(ptr(Tb), ptr(pib), K, int(k), B, int(max_steps), ptr(evals), ptr(info),
                             ptr(ws), ws.numel(), stream_handle(dev)), "pmb_eig_rev_topk")
    if squeeze:
        return evals[0], info[0]
    return evals, info


def tc_selftest(A: torch.Tensor, B: torch.Tensor, mode: int) -> torch.Tensor:
    """One 128 x N x Kdim TF32 tile product through tcgen05 (see pmb200.h)."""
    _dev(A, torch.float32, "A")
    _dev(B, torch.float32, "B")
    A, B = A.contiguous(), B.contiguous()
    if mode == 0:
        kdim, n = int(A.shape[1]), int(B.shape[0])
    else:
        kdim, n = int(A.shape[0]), int(B.shape[1])
    D = torch.empty((128, n), dtype=torch.float32, device=A.device)
    check(_lib.lib().pmb_tc_selftest(ptr(A), ptr(B), n, kdim, int(mode), ptr(D), stream_handle(A.device)),
          "pmb_tc_selftest")
    return D


def tc_selftest_raw(a_img: torch.Tensor, b_img: torch.Tensor, n: int, nsteps: int, lbo: int, sbo: int, layout: int,
                    step_a: int, step_b: int, idesc: int, kind: int) -> torch.Tensor:
    """tcgen05 product from caller-built shared-memory operand images (uint8 tensors); see pmb200.h."""
    _flat(a_img, torch.uint8, "a_img")
    _flat(b_img, torch.uint8, "b_img")
    D = torch.empty((128, int(n)), dtype=torch.float32, device=a_img.device)
    check(_lib.lib().pmb_tc_selftest_raw(ptr(a_img), int(a_img.numel()), ptr(b_img), int(b_img.numel()), int(n),
                                         int(nsteps), int(lbo), int(sbo), int(layout), int(step_a), int(step_b),
                                         int(idesc), int(kind), ptr(D), stream_handle(a_img.device)),
          "pmb_tc_selftest_raw")
    return D


def as_numpy(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()
