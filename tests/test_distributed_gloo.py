"""Host-side orchestration on CPU tensors: the kernels are replaced by numpy
statements of their contracts (tests/fake_kernels.py), the collectives run over
``gloo`` with world_size 2.  Checks that frame-sharded partials (moments with a
shared shift, Gram matrices, centroid sums, count matrices) combine to the
single-process / oracle result."""

from __future__ import annotations

import os
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch

import oracle
from tests import fake_kernels, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cfg():
    from pmarlo_b200.pipeline import PipelineConfig

    return PipelineConfig(tica_lag=5, tica_dim=3, preprocess="standard", n_states=9, kmeans_max_iter=4,
                          kmeans_tolerance=None, msm_lag=3, n_timescales=3, mle_maxerr=1e-12)


def _feats():
    f = synth.ar1_features(5, 700, 6, seed=21, offset=2.0)
    f[1] = f[1][:431]
    f[3] = f[3][:9]
    f[0][17, 2] = np.nan            # exercises imputation in the shared-shift moment algebra
    return f


def _run(feats, comm, c0):
    from pmarlo_b200.pipeline import run_pipeline
    from pmarlo_b200.shards import Segments

    X = torch.from_numpy(np.concatenate(feats, axis=0))
    segs = Segments.from_lengths([f.shape[0] for f in feats])
    return run_pipeline(None, segs, None, _cfg(), comm, features=X, initial_centers=torch.from_numpy(c0))


def _oracle(feats, c0):
    cfg = _cfg()
    flat = np.concatenate(feats, axis=0).astype(np.float64)
    Z = oracle.tica.preprocess(flat, scale=True)
    off = np.concatenate([[0], np.cumsum([f.shape[0] for f in feats])])
    prepped = [Z[off[i]:off[i + 1]] for i in range(len(feats))]
    model = oracle.tica.tica_fit(prepped, cfg.tica_lag)
    Y = oracle.tica.tica_transform(model, Z, cfg.tica_dim).astype(np.float32).astype(np.float64)
    c = c0.copy()
    for _ in range(cfg.kmeans_max_iter):
        lab, _ = oracle.kmeans.assign(Y, c)
        c, _ = oracle.kmeans._update(Y, lab, c)
    lab, _ = oracle.kmeans.assign(Y, c)
    C = oracle.counts.count_lagged([lab[off[i]:off[i + 1]] for i in range(len(feats))], cfg.n_states, cfg.msm_lag)
    T, pi = oracle.msm.build_simple_msm([lab[off[i]:off[i + 1]] for i in range(len(feats))], cfg.n_states,
                                        cfg.msm_lag, maxerr=1e-12)
    return model, c, C, T, pi


def _c0(feats):
    # initial centres in TICA space: a few projected frames of the oracle model
    cfg = _cfg()
    flat = np.concatenate(feats, axis=0).astype(np.float64)
    Z = oracle.tica.preprocess(flat, scale=True)
    off = np.concatenate([[0], np.cumsum([f.shape[0] for f in feats])])
    model = oracle.tica.tica_fit([Z[off[i]:off[i + 1]] for i in range(len(feats))], cfg.tica_lag)
    Y = oracle.tica.tica_transform(model, Z, cfg.tica_dim)
    return np.ascontiguousarray(Y[:: Y.shape[0] // cfg.n_states][: cfg.n_states])


def test_single_process_fake_pipeline_matches_oracle(monkeypatch):
    from pmarlo_b200.distributed import Comm

    fake_kernels.install(monkeypatch)
    feats = _feats()
    c0 = _c0(feats)
    res = _run(feats, Comm(), c0)
    model, c, C, T, pi = _oracle(feats, c0)
    np.testing.assert_allclose(res.tica.C00.numpy(), model.C00, rtol=0, atol=1e-6 * np.abs(model.C00).max())
    np.testing.assert_allclose(res.tica.C0t.numpy(), model.C0t, rtol=0, atol=1e-6 * np.abs(model.C0t).max())
    np.testing.assert_allclose(res.tica.eigenvalues.numpy()[:3], model.eigenvalues[:3], atol=1e-6)
    np.testing.assert_allclose(res.centers.numpy(), c, atol=1e-5)
    np.testing.assert_array_equal(res.counts.numpy(), C)
    np.testing.assert_allclose(res.T.numpy(), T, atol=1e-8)
    np.testing.assert_allclose(res.pi.numpy(), pi, atol=1e-8)


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pmarlo_b200.distributed import Comm
        from pmarlo_b200.shards import partition_trajectories

        fake_kernels.install(None)
        feats = _feats()
        c0 = _c0(feats)
        parts = partition_trajectories([f.shape[0] for f in feats], world)
        mine = [feats[i] for i in parts[rank]]
        comm = Comm()
        assert comm.size == world and comm.rank == rank
        assert comm.sum_int(rank + 1) == world * (world + 1) // 2
        t = torch.tensor([float(rank)])
        comm.broadcast(t, src=1)
        assert float(t.item()) == 1.0
        res = _run(mine, comm, c0)
        np.savez(f"{out_path}.{rank}.npz", C00=res.tica.C00.numpy(), C0t=res.tica.C0t.numpy(),
                 ev=res.tica.eigenvalues.numpy(), centers=res.centers.numpy(), counts=res.counts.numpy(),
                 T=res.T.numpy(), pi=res.pi.numpy(), n_pairs=res.tica.n_pairs, n_frames=res.tica.n_frames)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_pipeline_matches_oracle():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "res")
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        r0, r1 = np.load(out + ".0.npz"), np.load(out + ".1.npz")
    feats = _feats()
    c0 = _c0(feats)
    model, c, C, T, pi = _oracle(feats, c0)
    for r in (r0, r1):   # every rank holds the replicated result
        assert int(r["n_frames"]) == sum(f.shape[0] for f in feats)
        assert int(r["n_pairs"]) == model.n_pairs
        np.testing.assert_allclose(r["C00"], model.C00, rtol=0, atol=1e-6 * np.abs(model.C00).max())
        np.testing.assert_allclose(r["C0t"], model.C0t, rtol=0, atol=1e-6 * np.abs(model.C0t).max())
        np.testing.assert_allclose(r["ev"][:3], model.eigenvalues[:3], atol=1e-6)
        np.testing.assert_allclose(r["centers"], c, atol=1e-5)
        np.testing.assert_array_equal(r["counts"], C)          # integer sums: bit-exact for any rank count
        np.testing.assert_allclose(r["T"], T, atol=1e-8)
        np.testing.assert_allclose(r["pi"], pi, atol=1e-8)
    np.testing.assert_array_equal(r0["counts"], r1["counts"])
    np.testing.assert_allclose(r0["T"], r1["T"], atol=1e-14)


# ----------------------------------------------------------------------------- streaming TICA accumulator
def _stream_fit(feats, comm, chunks):
    """Chunk-wise fit as estimate_msm_from_host does it: ``chunks`` = list of lists of trajectory indices."""
    from pmarlo_b200.reduction import TICA
    from pmarlo_b200.shards import Segments

    cfg = _cfg()
    est = TICA(cfg.tica_lag, cfg.tica_dim, preprocess=cfg.preprocess, comm=comm)
    acc = est.accumulator(feats[0].shape[1], torch.device("cpu"))
    for idx in chunks:
        X = torch.from_numpy(np.concatenate([feats[i] for i in idx], axis=0))
        acc.add(X, Segments.from_lengths([feats[i].shape[0] for i in idx]))
    return acc.finish(), acc


def _check_model(m, om):
    assert m.n_pairs == om.n_pairs
    np.testing.assert_allclose(m.C00.numpy(), om.C00, rtol=0, atol=1e-6 * np.abs(om.C00).max())
    np.testing.assert_allclose(m.C0t.numpy(), om.C0t, rtol=0, atol=1e-6 * np.abs(om.C0t).max())
    np.testing.assert_allclose(m.eigenvalues.numpy()[:3], om.eigenvalues[:3], atol=1e-6)


def test_streaming_accumulator_matches_oracle(monkeypatch):
    """Conditioning from the first chunk (clean data) and the second-pass fallback (NaN in the data)."""
    from pmarlo_b200.distributed import Comm

    fake_kernels.install(monkeypatch)
    feats = _feats()
    c0 = _c0(feats)
    om = _oracle(feats, c0)[0]
    m, acc = _stream_fit(feats, Comm(), [[0, 1], [2], [3, 4]])
    assert acc.fell_back            # feats[0] holds a NaN
    _check_model(m, om)
    clean = [f.copy() for f in feats]
    clean[0][17, 2] = 0.25
    om2 = _oracle(clean, _c0(clean))[0]
    m2, acc2 = _stream_fit(clean, Comm(), [[0], [1, 2, 3], [4]])
    assert not acc2.fell_back
    _check_model(m2, om2)


def _stream_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pmarlo_b200.distributed import Comm
        from pmarlo_b200.shards import partition_trajectories

        fake_kernels.install(None)
        feats = [f.copy() for f in _feats()]
        feats[0][17, 2] = 0.25
        parts = partition_trajectories([f.shape[0] for f in feats], world)
        mine = [feats[i] for i in parts[rank]]
        # ranks use different chunkings; the broadcasts of shift / conditioning come from rank 0's first chunk
        chunks = [[i] for i in range(len(mine))] if rank == 0 else [list(range(len(mine)))]
        m, acc = _stream_fit(mine, Comm(), chunks)
        np.savez(f"{out_path}.{rank}.npz", C00=m.C00.numpy(), C0t=m.C0t.numpy(), ev=m.eigenvalues.numpy(),
                 n_pairs=m.n_pairs, fell_back=int(acc.fell_back))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_streaming_accumulator_matches_oracle():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "res")
        mp.spawn(_stream_worker, args=(2, port, out), nprocs=2, join=True)
        r0, r1 = np.load(out + ".0.npz"), np.load(out + ".1.npz")
    feats = [f.copy() for f in _feats()]
    feats[0][17, 2] = 0.25
    om = _oracle(feats, _c0(feats))[0]
    for r in (r0, r1):
        assert int(r["n_pairs"]) == om.n_pairs and int(r["fell_back"]) == 0
        np.testing.assert_allclose(r["C00"], om.C00, rtol=0, atol=1e-6 * np.abs(om.C00).max())
        np.testing.assert_allclose(r["C0t"], om.C0t, rtol=0, atol=1e-6 * np.abs(om.C0t).max())
        np.testing.assert_allclose(r["ev"][:3], om.eigenvalues[:3], atol=1e-6)
    np.testing.assert_allclose(r0["C00"], r1["C00"], rtol=0, atol=1e-14)


# ----------------------------------------------------------------------------- empty shards
def _empty_rank_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pmarlo_b200.distributed import Comm
        from pmarlo_b200.pipeline import seeded_initial_centers
        from pmarlo_b200.reduction import TICA
        from pmarlo_b200.shards import Segments

        fake_kernels.install(None)
        feats = [f.copy() for f in _feats()]
        feats[0][17, 2] = 0.25
        cfg = _cfg()
        comm = Comm()
        # rank 1 holds NO frames: it must still take part in every broadcast of the first chunk
        est = TICA(cfg.tica_lag, cfg.tica_dim, preprocess=cfg.preprocess, comm=comm)
        acc = est.accumulator(feats[0].shape[1], torch.device("cpu"))
        if rank == 0:
            for f in feats:
                acc.add(torch.from_numpy(f), Segments.from_lengths([f.shape[0]]))
        else:
            acc.add(torch.zeros((0, feats[0].shape[1]), dtype=torch.float32), Segments(np.zeros((1,), dtype=np.int64)))
        m = acc.finish()
        # a rank-0 shard that is too small for the seeding raises on EVERY rank instead of deadlocking the others
        Y = torch.zeros((3 if rank == 0 else 50, 2), dtype=torch.float32)
        try:
            seeded_initial_centers(Y, 9, 0, comm)
            raised = 0
        except ValueError:
            raised = 1
        np.savez(f"{out_path}.{rank}.npz", C00=m.C00.numpy(), ev=m.eigenvalues.numpy(), n_pairs=m.n_pairs, raised=raised)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_empty_shard_and_collective_validation():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "res")
        mp.spawn(_empty_rank_worker, args=(2, port, out), nprocs=2, join=True)
        r0, r1 = np.load(out + ".0.npz"), np.load(out + ".1.npz")
    feats = [f.copy() for f in _feats()]
    feats[0][17, 2] = 0.25
    om = _oracle(feats, _c0(feats))[0]
    for r in (r0, r1):
        assert int(r["n_pairs"]) == om.n_pairs and int(r["raised"]) == 1
        np.testing.assert_allclose(r["C00"], om.C00, rtol=0, atol=1e-6 * np.abs(om.C00).max())
        np.testing.assert_allclose(r["ev"][:3], om.eigenvalues[:3], atol=1e-6)
