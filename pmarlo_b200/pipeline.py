"""End-to-end MSM estimation on a frame shard, device resident.

featurize -> (scaler) -> TICA fit + project -> k-means (Lloyd) -> lagged counts
-> reversible MLE -> leading eigenvalues -> implied timescales: the chain
``compute_features -> reduce_features -> cluster_microstates ->
build_msm_from_labels`` of src/pmarlo/api/conformations.py:192-200 (call stack
(A) of SURVEY.md section 3), with the trajectories ("shards") dealt to one
process per GPU and the small partials summed over ranks.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from . import kernels
from .clustering import lloyd_device
from .distributed import Comm
from .features import FeaturePlan, featurize_device
from .msm import msm_from_counts_device, safe_timescales
from .reduction import TICA, TicaModel
from .shards import Segments
from .timing import NULL_TIMER, StageTimer

__all__ = ["PipelineConfig", "PipelineResult", "StageTimer", "run_pipeline", "seeded_initial_centers",
           "estimate_msm_from_host"]


@dataclass
class PipelineConfig:
    tica_lag: int = 10
    tica_dim: int = 2
    preprocess: str | None = "standard"
    n_states: int = 100
    kmeans_max_iter: int = 500
    kmeans_tolerance: float | None = 1e-5     # None: run exactly kmeans_max_iter iterations
    msm_lag: int = 10
    n_timescales: int = 5
    dirichlet_alpha: float = 1e-3             # src/pmarlo/constants.py:44
    mle_maxerr: float = 1e-8
    mle_maxiter: int = 1_000_000
    gram_impl: int = 0
    seed: int = 0


@dataclass
class PipelineResult:
    features: torch.Tensor | None
    tica: TicaModel | None
    Y: torch.Tensor
    centers: torch.Tensor
    labels: torch.Tensor
    counts: torch.Tensor
    T: torch.Tensor
    pi: torch.Tensor
    active: torch.Tensor
    mle_info: torch.Tensor
    eigenvalues: torch.Tensor
    timescales: np.ndarray | None = None
    kmeans_iters: int = 0
    extra: dict = field(default_factory=dict)


def seeded_initial_centers(Y: torch.Tensor, K: int, seed: int, comm: Comm) -> torch.Tensor:
    """K distinct frames of rank 0's shard chosen by a seeded permutation, broadcast
    to every rank (the ``initial_centers=`` kwarg of clustering.py:243-250)."""
    centers = torch.empty((K, int(Y.shape[1])), dtype=torch.float64, device=Y.device)
    # validated collectively: a raise on rank 0 alone would leave the other ranks inside the broadcast
    n0 = torch.tensor([int(Y.shape[0]) if comm.rank == 0 else 0], dtype=torch.int64, device=Y.device)
    comm.allreduce_sum(n0)
    if int(n0.item()) < K:
        raise ValueError(f"rank 0 holds {int(n0.item())} frames, fewer than n_states={K}")
    if comm.rank == 0:
        n = int(Y.shape[0])
        # O(K) draw (numpy's Generator.choice without replacement does not permute all n frames)
        pick = np.sort(np.random.default_rng(int(seed)).choice(n, size=K, replace=False))
        idx = torch.from_numpy(pick).to(Y.device)
        centers.copy_(Y.index_select(0, idx).to(torch.float64))
    comm.broadcast(centers, src=0)
    return centers


def run_pipeline(xyz: torch.Tensor | None, segs: Segments, plan: FeaturePlan | None, cfg: PipelineConfig,
                 comm: Comm | None = None, features: torch.Tensor | None = None,
                 initial_centers: torch.Tensor | None = None, timer: StageTimer | None = None,
                 read_back: bool = True, buffers: dict | None = None,
                 tica_model: TicaModel | None = None,
                 initial_center_rows: np.ndarray | None = None) -> PipelineResult:
    """Run the whole path on this rank's shard.  Either ``xyz`` (+ ``plan``) or
    precomputed ``features`` (N,d) float32 must be given; with ``cfg.tica_dim <= 0``
    the features are clustered directly (config C2: 2-D Mueller-Brown data).
    ``buffers``: a dict that keeps the large per-frame tensors (features, Y, labels) alive between
    calls so that repeated runs on same-sized shards do not go through the allocator.
    ``initial_center_rows``: frame indices into RANK 0's shard whose projected coordinates seed Lloyd
    (``initial_centers=`` given as frames, so that runs with different rank counts can share them)."""
    comm = comm if comm is not None else Comm()
    timer = timer if timer is not None else NULL_TIMER
    dev = (xyz if xyz is not None else features).device
    off = segs.device(dev)

    def _buf(name, shape, dtype):
        if buffers is None:
            return None
        t = buffers.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype or t.device != dev:
            t = torch.empty(shape, dtype=dtype, device=dev)
            buffers[name] = t
        return t

    if features is None:
        out = _buf("features", (int(xyz.shape[0]), plan.n_cols), torch.float32)
        with timer.stage("featurize"):
            features = featurize_device(xyz, plan, out=out)
    X = features

    if cfg.tica_dim > 0:
        est = TICA(cfg.tica_lag, cfg.tica_dim, preprocess=cfg.preprocess, comm=comm, gram_impl=cfg.gram_impl)
        if tica_model is None:
            with timer.stage("tica_fit"):
                tica_model = est.fit_device(X, segs, off, timer=timer)
        with timer.stage("project"):
            Y = est.transform_device(tica_model, X, out_f64=False,
                                     out=_buf("Y", (int(X.shape[0]), tica_model.dim), torch.float32))
    else:
        Y = X

    with timer.stage("kmeans"):
        if initial_centers is None and initial_center_rows is not None:
            initial_centers = torch.empty((cfg.n_states, int(Y.shape[1])), dtype=torch.float64, device=dev)
            if comm.rank == 0:
                idx = torch.from_numpy(np.asarray(initial_center_rows, dtype=np.int64)).to(dev)
                initial_centers.copy_(Y.index_select(0, idx).to(torch.float64))
            comm.broadcast(initial_centers, src=0)
        if initial_centers is None:
            initial_centers = seeded_initial_centers(Y, cfg.n_states, cfg.seed, comm)
        labels = _buf("labels", (int(Y.shape[0]),), torch.int32)
        if labels is None:
            labels = torch.empty((int(Y.shape[0]),), dtype=torch.int32, device=dev)
        res = lloyd_device(Y, initial_centers, cfg.kmeans_max_iter, cfg.kmeans_tolerance, comm, labels=labels,
                           timer=timer)
        # final labels against the final centres (model.transform, clustering.py:609)
        with timer.stage("kmeans_assign"):
            kernels.kmeans_assign(Y, res.centers, labels=labels, hints=labels)

    with timer.stage("count"):
        C = torch.zeros((cfg.n_states, cfg.n_states), dtype=torch.int64, device=dev)
        kernels.count_lagged(labels, off, cfg.n_states, cfg.msm_lag, 1, out=C)
        comm.allreduce_sum(C)

    with timer.stage("mle"):
        T, pi, info, active = msm_from_counts_device(C, alpha=cfg.dirichlet_alpha, maxerr=cfg.mle_maxerr,
                                                     maxiter=cfg.mle_maxiter)
    with timer.stage("eig"):
        k = min(cfg.n_timescales + 1, cfg.n_states)
        ev, ev_info = kernels.eig_rev_topk(T, pi, k)

    ts = None
    if read_back:
        ts = safe_timescales(cfg.msm_lag, ev.cpu().numpy()[1:])
    return PipelineResult(features=X, tica=tica_model, Y=Y, centers=res.centers, labels=labels, counts=C,
                          T=T, pi=pi, active=active, mle_info=info, eigenvalues=ev, timescales=ts,
                          kmeans_iters=res.n_iter, extra={"eig_info": ev_info})


def estimate_msm_from_host(xyz_host, lengths, plan: FeaturePlan, cfg: PipelineConfig, comm: Comm | None = None,
                           device=None, buffers: dict | None = None, chunk_frames: int = 1_000_000) -> dict:
    """The call a user of the host-buffer API makes: coordinates of this rank's
    trajectories in (pinned) host memory -> timescales, eigenvalues, stationary
    vector and MLE diagnostics as numpy arrays.  The host->device copy of the
    coordinates and the device->host read of the results happen inside."""
    device = device if device is not None else kernels.require_cuda()
    comm = comm if comm is not None else Comm()
    if isinstance(xyz_host, np.ndarray):
        xyz_host = torch.from_numpy(np.ascontiguousarray(xyz_host, dtype=np.float32))
    segs = Segments.from_lengths(lengths)
    n = segs.n_frames
    if n != int(xyz_host.shape[0]):
        raise ValueError("lengths do not add up to the number of frames")
    bufs = buffers if buffers is not None else {}

    def _buf(name, shape, dtype):
        t = bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype or t.device != device:
            t = torch.empty(shape, dtype=dtype, device=device)
            bufs[name] = t
        return t

    xyz = _buf("xyz", tuple(xyz_host.shape), torch.float32)
    X = _buf("features", (n, plan.n_cols), torch.float32)
    # chunks of whole trajectories, about chunk_frames each: the copy of chunk c+1 (side stream) overlaps
    # featurization and the moment / Gram accumulation of chunk c (compute stream)
    bounds = [0]
    for end in segs.offsets[1:]:
        if int(end) - bounds[-1] >= chunk_frames:
            bounds.append(int(end))
    if bounds[-1] != n:
        bounds.append(n)
    main = torch.cuda.current_stream(device)
    copy_stream = bufs.get("copy_stream")
    if copy_stream is None:
        copy_stream = torch.cuda.Stream(device=device)
        bufs["copy_stream"] = copy_stream
    copy_stream.wait_stream(main)          # the previous call may still be reading xyz
    events = []
    with torch.cuda.stream(copy_stream):
        for a, b in zip(bounds[:-1], bounds[1:]):
            xyz[a:b].copy_(xyz_host[a:b], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
            events.append(ev)
    streaming = cfg.tica_dim > 0
    acc = None
    if streaming:
        est = TICA(cfg.tica_lag, cfg.tica_dim, preprocess=cfg.preprocess, comm=comm, gram_impl=cfg.gram_impl)
        acc = est.accumulator(plan.n_cols, device)
    offs = segs.offsets
    for (a, b), ev in zip(zip(bounds[:-1], bounds[1:]), events):
        main.wait_event(ev)
        featurize_device(xyz[a:b], plan, out=X[a:b])
        if acc is not None:
            lo, hi = int(np.searchsorted(offs, a)), int(np.searchsorted(offs, b))
            acc.add(X[a:b], Segments(offs[lo:hi + 1] - a))
    if acc is not None and acc.moments is None:
        # this rank holds no frames: it still joins the first chunk's broadcasts of the other ranks
        acc.add(X[0:0], Segments(np.zeros((1,), dtype=np.int64)))
    model = acc.finish() if acc is not None else None
    res = run_pipeline(None, segs, plan, cfg, comm, features=X, read_back=False, buffers=buffers, tica_model=model)
    # what the reference's API hands back: labels (cluster_microstates), T and pi (build_msm_from_labels),
    # eigenvalues / timescales; read into pinned host buffers kept between calls
    def _pinned(name, t):
        h = bufs.get(name)
        if h is None or tuple(h.shape) != tuple(t.shape) or h.dtype != t.dtype:
            h = torch.empty(tuple(t.shape), dtype=t.dtype, pin_memory=True)
            bufs[name] = h
        h.copy_(t, non_blocking=True)
        return h

    h_lab = _pinned("h_labels", res.labels)
    h_T = _pinned("h_T", res.T)
    h_pi = _pinned("h_pi", res.pi)
    h_ev = _pinned("h_ev", res.eigenvalues)
    h_info = _pinned("h_info", res.mle_info)
    torch.cuda.current_stream(device).synchronize()
    ev = h_ev.numpy()
    ts = safe_timescales(cfg.msm_lag, ev[1:])
    return {"timescales": ts, "eigenvalues": ev, "stationary_distribution": h_pi.numpy(), "mle_info": h_info.numpy(),
            "labels": h_lab.numpy(), "transition_matrix": h_T.numpy(),
            "d2h_bytes": int(sum(h.numel() * h.element_size() for h in (h_lab, h_T, h_pi, h_ev, h_info)))}
