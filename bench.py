#!/usr/bin/env python
"""bench.py -- frames/s of the end-to-end MSM pipeline on B200 (BASELINE.json metric).

One "step" = one pass of the whole hot path (featurize -> z-score -> TICA fit +
project -> k-means Lloyd -> lagged counts -> reversible MLE -> leading
eigenvalues -> implied timescales) over this rank's shard of synthetic
protein-scale trajectories (config C4 of BASELINE.json / SURVEY.md section 8d:
33-residue backbone = 99 atoms, 256 features = cos/sin of 32 phi + 32 psi + the
first 128 C-alpha pair distances, z-score, TICA lag 20 -> 10 dims, k-means 1000
states with 20 Lloyd iterations from seeded initial centres, MSM lag 20).

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    # CPU reference arm (oracle port)

Prints ONE JSON line (rank 0).  `value` has the coordinates resident in HBM when
the timed region starts; `e2e` goes through the host-buffer API with the
host->device copy of the coordinates and the device->host read of the result
inside the timed region.
"""

from __future__ import annotations

import argparse
import json
import os
import pathlib
import subprocess
import sys
import tempfile
import time
from dataclasses import dataclass

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

N_RES = 33
N_DIST = 128
FRAMES_PER_TRAJ = 125_000
TICA_LAG, TICA_DIM, N_STATES, KMEANS_ITERS, MSM_LAG, N_TIMESCALES = 20, 10, 1000, 20, 20, 5
METRIC = "frames/sec end-to-end MSM pipeline"
# CPU arm: the reversible-MLE fixed point costs K^2 = 10^6 cells per iteration on the host whatever the number
# of frames, and a small sample's sparse count matrix needs > 50 000 iterations.  The full C4 count matrix
# (10 M frames) converges in ~1.3 x 10^3 iterations (bench.py prints mle_iters), so the CPU sample is capped there.
CPU_MLE_ITER_CAP = 1324
CPU_SAMPLE_FRAMES = 250_000     # ~10 s of host work per step: K^2-sized stages stay below 15 % of the CPU step
UNIT = "frames/s"


# ----------------------------------------------------------------------------- workload
@dataclass
class Workload:
    xyz: "object"            # torch (N, 99, 3) float32 on the device
    segs: "object"           # pmarlo_b200.shards.Segments
    plan: "object"           # FeaturePlan (256 columns)
    top: "object"


def c4_plan():
    from pmarlo_b200.features import ca_pairs_all, plan_concat, plan_distances, plan_phi_psi_block
    from tests.synth import backbone_topology

    top = backbone_topology(N_RES)
    pairs = ca_pairs_all(top.select_name("CA"))[:N_DIST]
    plan = plan_concat([plan_phi_psi_block(top), plan_distances(pairs)])
    assert plan.n_cols == 256, plan.n_cols
    return top, plan, pairs


def bench_config(n_states=N_STATES, kmeans_iters=KMEANS_ITERS, gram_impl=0):
    from pmarlo_b200.pipeline import PipelineConfig

    return PipelineConfig(tica_lag=TICA_LAG, tica_dim=TICA_DIM, preprocess="standard", n_states=n_states,
                          kmeans_max_iter=kmeans_iters, kmeans_tolerance=None, msm_lag=MSM_LAG,
                          n_timescales=N_TIMESCALES, mle_maxerr=1e-8, mle_maxiter=1_000_000,
                          gram_impl=gram_impl, seed=4)


def synth_xyz_device(n_traj, frames_per_traj, device, seed, rho=0.9995, sigma=0.03, chunk=2048, base=None):
    """AR(1) internal motion (rho) around a random-coil 33-residue backbone (or the mean structure
    ``base`` (A,3) nm), generated on the device chunk by chunk: x_t = rho^t (x_0 + sum_{s<=t} rho^-s e_s)
    inside a chunk."""
    import torch

    # the molecule (mean structure) is the same on every rank; only the dynamics are seeded per rank
    g0 = torch.Generator(device=device)
    g0.manual_seed(4)
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    if base is None:
        A = 3 * N_RES
        steps = torch.randn((A, 3), generator=g0, device=device, dtype=torch.float64)
        steps /= steps.norm(dim=1, keepdim=True)
        base = torch.cumsum(0.15 * steps, dim=0).to(torch.float32)
    else:
        base = base.to(device=device, dtype=torch.float32)
        A = int(base.shape[0])
    out = torch.empty((n_traj, frames_per_traj, A, 3), dtype=torch.float32, device=device)
    state = sigma * torch.randn((n_traj, 1, A, 3), generator=g, device=device, dtype=torch.float32)
    amp = sigma * (1 - rho * rho) ** 0.5
    for s in range(0, frames_per_traj, chunk):
        n = min(chunk, frames_per_traj - s)
        t = torch.arange(1, n + 1, device=device, dtype=torch.float32).view(1, n, 1, 1)
        e = amp * torch.randn((n_traj, n, A, 3), generator=g, device=device, dtype=torch.float32)
        x = (rho ** t) * (state + torch.cumsum(e * (rho ** (-t)), dim=1))
        out[:, s:s + n] = x
        state = x[:, -1:].clone()
    out += base
    return out.view(n_traj * frames_per_traj, A, 3)


def make_workload(n_traj, frames_per_traj, device, seed) -> Workload:
    from pmarlo_b200.shards import Segments

    top, plan, _ = c4_plan()
    xyz = synth_xyz_device(n_traj, frames_per_traj, device, seed)
    return Workload(xyz, Segments.from_lengths([frames_per_traj] * n_traj), plan, top)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampler running DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                 "-i", str(self.gpu_index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_sample_inputs(n_traj, frames_per_traj, seed):
    """Host copy of a bounded sample of the same synthetic workload (numpy AR(1))."""
    from tests.synth import backbone_trajectories

    return backbone_trajectories(N_RES, n_traj, frames_per_traj, seed, rho=0.9995, sigma=0.03)


def run_cpu_pipeline(trajs, n_states=N_STATES, kmeans_iters=KMEANS_ITERS):
    from oracle import featurize as ofeat
    from oracle import pipeline as opipe
    from tests.synth import backbone_topology

    top = backbone_topology(N_RES)
    phi = ofeat.dihedral_quads(top.names, top.resid, top.chainid, "phi")
    psi = ofeat.dihedral_quads(top.names, top.resid, top.chainid, "psi")
    pairs = ofeat.ca_pairs_all(ofeat.ca_indices(top.names))[:N_DIST]
    return opipe.run(trajs, phi, psi, pairs, tica_lag=TICA_LAG, tica_dim=TICA_DIM, n_states=n_states,
                     kmeans_iters=kmeans_iters, msm_lag=MSM_LAG, n_timescales=N_TIMESCALES, seed=4,
                     mle_maxiter=CPU_MLE_ITER_CAP, threads=cpu_threads())


def cpu_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def time_cpu(sample_traj, sample_frames, repeats=1):
    trajs = cpu_sample_inputs(sample_traj, sample_frames, seed=4)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        res = run_cpu_pipeline(trajs)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return res.n_frames / best, best, res


CPU_IMPL_NOTE = ("numpy/scipy BLAS + scikit-learn Lloyd + plain-C (pthread) reversible MLE port of the algorithm; "
                 "NOT deeptime's C++/OpenMP nor mdtraj's C (absent from the image)")


def cpu_cost_split(stage_seconds, n_frames):
    """Per-frame and fixed (K^2-sized, frame-independent) parts of the CPU step, so that the extrapolation
    to the GPU arm's frame count is explicit: t(n) ~ fixed_s + n * per_frame_us."""
    fixed = sum(float(stage_seconds.get(k, 0.0)) for k in ("mle", "eig"))
    per_frame = sum(float(stage_seconds.get(k, 0.0)) for k in ("featurize", "tica_fit", "project", "kmeans", "count"))
    return {"per_frame_us": 1e6 * per_frame / max(1, n_frames), "fixed_s": fixed,
            "asymptotic_frames_per_s": max(1, n_frames) / per_frame if per_frame > 0 else None}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = cpu_threads()
    n_traj, n_frames = 2, args.cpu_sample_frames // 2
    trajs = cpu_sample_inputs(n_traj, n_frames, seed=4)
    for _ in range(args.warmup):
        run_cpu_pipeline(trajs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = run_cpu_pipeline(trajs)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = res.n_frames / dt
    sample = (f"{n_traj} trajectories x {n_frames} frames of the C4 workload per step (same shapes, K={N_STATES}, "
              f"{KMEANS_ITERS} Lloyd iterations, reversible MLE capped at {CPU_MLE_ITER_CAP} iterations = what the full "
              f"10 M-frame count matrix needs) on {cores} threads; {CPU_IMPL_NOTE}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, n_traj * n_frames, cpu=True),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         **cpu_cost_split(res.stage_seconds, res.n_frames)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "stages_s": res.stage_seconds,
        "fixed_share_of_step": (float(res.stage_seconds.get("mle", 0.0)) + float(res.stage_seconds.get("eig", 0.0))) / dt,
        "parity": PARITY_NOTE,
    }
    print(json.dumps(line))
    return 0


def workload_config(args, frames_per_gpu, cpu=False):
    return {
        "workload": ("C4 synthetic protein-scale: 33-residue backbone (99 atoms) -> 256 features "
                     "(cos/sin of 32 phi + 32 psi, 128 CA distances), z-score, TICA lag 20 -> 10 dims, "
                     "k-means K=1000 x 20 Lloyd iterations, counts + reversible MLE + top-6 eigenvalues at lag 20"),
        "frames_per_gpu": int(frames_per_gpu), "frames_per_trajectory": FRAMES_PER_TRAJ if not cpu else None,
        "n_features": 256, "n_states": N_STATES, "kmeans_iters": KMEANS_ITERS, "tica_lag": TICA_LAG,
        "msm_lag": MSM_LAG, "parallelism": f"frame shards x{args.gpus}, allreduce of partial sums",
        "l2_policy": "inputs (1188 B of coordinates per frame, >= 1.4 GB per step) larger than the 126 MB L2",
    }


# ----------------------------------------------------------------------------- extra measurements
def h2d_ceiling(host, dst, comm, device, repeats=3):
    """Pinned host -> device copy rate with every rank copying AT THE SAME TIME and no kernels running:
    the platform ceiling the e2e number (11.9 GB of coordinates per step and GPU) is bounded by."""
    import torch

    n = min(int(host.shape[0]), int(dst.shape[0]))
    nbytes = n * int(host[0].numel()) * 4
    best = None
    for _ in range(repeats + 1):
        torch.cuda.synchronize(device)
        comm.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dst[:n].copy_(host[:n], non_blocking=True)
        b.record()
        torch.cuda.synchronize(device)
        ms = a.elapsed_time(b)
        best = ms if best is None else min(best, ms)
    t = torch.tensor([best], dtype=torch.float64, device=device)
    comm.allreduce_max(t)                      # slowest rank: all ranks copy concurrently
    gbs = nbytes / (float(t.item()) * 1e-3) / 1e9
    return {"per_gpu_gbs": gbs, "aggregate_gbs": gbs * comm.size, "bytes": nbytes, "concurrent_ranks": comm.size}


def strong_scaling_record(args, comm, device, rank, world, cfg):
    """The north-star configuration itself: C4's 10 M frames IN TOTAL dealt over the ranks (80 trajectories,
    80 / N per GPU), device resident, same pipeline.  Reported next to the weak-scaling `value`."""
    import torch

    from pmarlo_b200.pipeline import StageTimer, run_pipeline

    total_traj = 80
    per = total_traj // world
    wl = make_workload(per, FRAMES_PER_TRAJ, device, seed=4000 + rank)
    bufs: dict = {}
    res = None
    for _ in range(max(3, min(args.warmup, 5))):
        res = run_pipeline(wl.xyz, wl.segs, wl.plan, cfg, comm, read_back=True, buffers=bufs)
    torch.cuda.synchronize(device)
    comm.barrier()
    torch.cuda.synchronize(device)
    steps = max(3, min(args.steps, 10))
    timer = StageTimer(True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        res = run_pipeline(wl.xyz, wl.segs, wl.plan, cfg, comm, timer=timer, read_back=True, buffers=bufs)
    b.record()
    torch.cuda.synchronize(device)
    comm.barrier()
    ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=device)
    comm.allreduce_max(ms)
    ms_step = float(ms.item()) / steps
    frames_total = per * world * FRAMES_PER_TRAJ
    return {"frames_total": frames_total, "frames_per_gpu": per * FRAMES_PER_TRAJ, "steps": steps,
            "ms_per_step": ms_step, "value": frames_total / (ms_step * 1e-3), "unit": UNIT,
            "stages_ms": {k: v / steps for k, v in timer.totals_ms().items() if k in TOP_LEVEL_STAGES},
            "mle_iters": int(res.mle_info[0].item())}


def multi_rank_parity(comm, device, rank, world, plan):
    """N-rank results against a 1-rank run of the same pipeline over the concatenated shards (rank 0 runs
    it with a solo communicator).  Counts given identical labels must be bit-exact for any N (integer
    sums); along the whole chain the fp64 partial sums are added in a different order, so covariances agree
    to ~1e-15 and a frame sitting on a Voronoi boundary may change its label."""
    import torch
    import torch.distributed as dist

    from pmarlo_b200 import kernels
    from pmarlo_b200.distributed import Comm
    from pmarlo_b200.pipeline import PipelineConfig, run_pipeline
    from pmarlo_b200.shards import Segments

    n_traj, fpt, K = 2, 20_000, 200
    xyz = synth_xyz_device(n_traj, fpt, device, seed=9000 + rank, rho=0.995)
    segs = Segments.from_lengths([fpt] * n_traj)
    cfg = PipelineConfig(tica_lag=TICA_LAG, tica_dim=TICA_DIM, preprocess="standard", n_states=K, kmeans_max_iter=5,
                         kmeans_tolerance=None, msm_lag=MSM_LAG, n_timescales=N_TIMESCALES, seed=4)
    rows = np.sort(np.random.default_rng(12).choice(n_traj * fpt, size=K, replace=False))
    res = run_pipeline(xyz, segs, plan, cfg, comm, initial_center_rows=rows)
    all_xyz = [torch.empty_like(xyz) for _ in range(world)]
    all_lab = [torch.empty_like(res.labels) for _ in range(world)]
    dist.all_gather(all_xyz, xyz.contiguous())
    dist.all_gather(all_lab, res.labels.contiguous())
    rec = None
    if rank == 0:
        solo = Comm(solo=True)
        X1 = torch.cat(all_xyz, dim=0)
        segs1 = Segments.from_lengths([fpt] * (n_traj * world))
        one = run_pipeline(X1, segs1, plan, cfg, solo, initial_center_rows=rows)
        labN = torch.cat(all_lab, dim=0)
        C_given = kernels.count_lagged(labN, segs1.device(device), K, cfg.msm_lag)

        def rel(a, b):
            a, b = a.double(), b.double()
            return float((a - b).abs().max().item() / max(float(b.abs().max().item()), 1e-300))

        # fp64 reference of the symmetrised covariances on the concatenated shards (torch.matmul, test arithmetic)
        F = one.features.to(torch.float64)
        mu_all, sd_all = F.mean(dim=0), F.std(dim=0, unbiased=False)
        Zs = (F - mu_all) / torch.where(sd_all > 0, sd_all, torch.ones_like(sd_all))
        Zt = Zs.view(n_traj * world, fpt, -1)
        X0, Xt = Zt[:, :-cfg.tica_lag].reshape(-1, Zs.shape[1]), Zt[:, cfg.tica_lag:].reshape(-1, Zs.shape[1])
        mu = 0.5 * (X0.mean(dim=0) + Xt.mean(dim=0))
        X0c, Xtc = X0 - mu, Xt - mu
        Tn = float(X0.shape[0])
        C00_ref = (X0c.T @ X0c + Xtc.T @ Xtc) / (2 * Tn)
        C0t_ref = (X0c.T @ Xtc + Xtc.T @ X0c) / (2 * Tn)
        rec = {
            "ranks": world, "frames": int(X1.shape[0]), "n_states": K,
            "C00_rel_vs_fp64": {"n_rank": rel(res.tica.C00, C00_ref), "one_rank": rel(one.tica.C00, C00_ref)},
            "C0t_rel_vs_fp64": {"n_rank": rel(res.tica.C0t, C0t_ref), "one_rank": rel(one.tica.C0t, C0t_ref)},
            "counts_equal_given_labels": bool(torch.equal(C_given, res.counts)),
            "labels_mismatch": int((labN != one.labels).sum().item()),
            "counts_equal": bool(torch.equal(one.counts, res.counts)),
            "C00_rel": rel(res.tica.C00, one.tica.C00), "C0t_rel": rel(res.tica.C0t, one.tica.C0t),
            "T_rel": rel(res.T, one.T), "pi_rel": rel(res.pi, one.pi),
            "eig_rel": rel(res.eigenvalues, one.eigenvalues),
        }
    comm.barrier()
    return rec


def end_to_end_roofline(stages, frames, cfg, peaks, world):
    """T_min = sum over stages of max(bytes / HBM peak, flops / tensor-or-fp32 peak)
    for this rank's frames (dense-equivalent algorithmic work, no 3xTF32 split factor) + all-reduce volume /
    NVLink rate; `frac` = T_min / measured step."""
    hbm = float(peaks.get("hbm_gbs", 6650.0)) * 1e9
    tf32 = float(peaks.get("bf16_tflops_sustained", 1400.0)) / 2.0 * 1e12
    d, m, K = 256, cfg.tica_dim, cfg.n_states
    per_frame = {
        "featurize": (12 * 3 * N_RES + 4 * d, 0.0),
        "col_moments": (4 * d, 0.0),
        "gram": (2 * 4 * d, 2 * 2 * d * d),                                   # two launches
        "project": (4 * d + 4 * m, 2 * d * m),
        "kmeans": ((4 * m + 4) * (cfg.kmeans_max_iter + 1), 2 * m * K * (cfg.kmeans_max_iter + 1)),
        "count": (4, 0.0),
    }
    t_min = 0.0
    parts = {}
    for name, (b, f) in per_frame.items():
        t = max(b * frames / hbm, f * frames / tf32)
        parts[name] = t * 1e3
        t_min += t
    nv = 770e9
    allreduce_bytes = (2 * d * d * 8 + 6 * d * 8 + (K * m + K + 1) * 8 * (cfg.kmeans_max_iter + 1) + K * K * 8) if world > 1 else 0
    t_min += 2.0 * allreduce_bytes / nv
    measured = sum(v for k, v in stages.items() if k in TOP_LEVEL_STAGES)
    return {"t_min_ms": t_min * 1e3, "measured_ms": measured, "frac": (t_min * 1e3) / measured if measured else None,
            "parts_ms": parts, "note": "dense-equivalent algorithmic work per SURVEY 8(d); the latency-bound d x d and "
                                       "K x K solves (K4, K8, K9) have no roofline term and count fully against the fraction"}


# ----------------------------------------------------------------------------- GPU arm
def gpu_arm(args):
    import torch
    import torch.distributed as dist

    from pmarlo_b200 import _lib
    from pmarlo_b200.distributed import Comm
    from pmarlo_b200.pipeline import StageTimer, estimate_msm_from_host, run_pipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    comm = Comm()
    _lib.load()

    n_traj = max(1, args.frames_per_gpu // FRAMES_PER_TRAJ)
    fpt = FRAMES_PER_TRAJ if args.frames_per_gpu >= FRAMES_PER_TRAJ else args.frames_per_gpu
    frames = n_traj * fpt
    wl = make_workload(n_traj, fpt, device, seed=4000 + rank)
    cfg = bench_config(gram_impl=args.gram_impl)

    def sync():
        torch.cuda.synchronize(device)
        comm.barrier()
        torch.cuda.synchronize(device)

    # ---- device-resident timing -------------------------------------------------
    timer = StageTimer(True)
    bufs: dict = {}
    # Warm-up runs exactly what the timed loop runs, INCLUDING keeping the previous step's result alive
    # while the next one is computed: torch's caching allocator then reaches its steady state here.  (With
    # the result discarded during warm-up the second timed step needed new blocks, and the cudaMalloc inside
    # a small torch.empty stalled the host for 15-140 ms while the GPU sat idle -- measured, round 1.)
    res = None
    for _ in range(args.warmup):
        res = run_pipeline(wl.xyz, wl.segs, wl.plan, cfg, comm, timer=StageTimer(True), read_back=True, buffers=bufs)
    sync()
    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("PMB_BENCH_NO_SAMPLER"):
        sampler.start()
    l0 = _lib.launch_count()
    if args.profile_range:
        torch.cuda.cudart().cudaProfilerStart()   # ncu --profile-from-start off: only the timed steps
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        res = run_pipeline(wl.xyz, wl.segs, wl.plan, cfg, comm, timer=timer, read_back=True, buffers=bufs)
    ev1.record()
    sync()
    if args.profile_range:
        torch.cuda.cudart().cudaProfilerStop()
    launches = _lib.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=device)
    comm.allreduce_max(ms)
    ms_per_step = float(ms.item()) / args.steps
    value = frames * world / (ms_per_step * 1e-3)
    stages = {k: v / args.steps for k, v in timer.totals_ms().items()}
    counts = timer.counts()

    # ---- end to end through the host-buffer API ---------------------------------
    host = torch.empty(wl.xyz.shape, dtype=torch.float32, pin_memory=True)
    host.copy_(wl.xyz)
    del wl.xyz
    torch.cuda.empty_cache()
    lengths = [fpt] * n_traj
    for _ in range(max(1, min(3, args.warmup))):
        out = estimate_msm_from_host(host, lengths, wl.plan, cfg, comm, buffers=bufs)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.e2e_steps):
        out = estimate_msm_from_host(host, lengths, wl.plan, cfg, comm, buffers=bufs)
    e1.record()
    sync()
    ems = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    comm.allreduce_max(ems)
    e2e_ms = float(ems.item()) / args.e2e_steps
    e2e_value = frames * world / (e2e_ms * 1e-3)
    h2d = int(host.numel() * 4)
    d2h = int(out["d2h_bytes"])
    ceiling = h2d_ceiling(host, bufs["xyz"], comm, device)
    del host
    bufs.clear()
    torch.cuda.empty_cache()

    strong = verify = None
    if world == 1 and frames == 10_000_000:
        strong = {"frames_total": frames, "frames_per_gpu": frames, "steps": args.steps, "ms_per_step": ms_per_step,
                  "value": value, "unit": UNIT, "stages_ms": {k: v for k, v in stages.items() if k in TOP_LEVEL_STAGES},
                  "mle_iters": int(res.mle_info[0].item())}
    if world > 1 and not args.no_extras:
        strong = strong_scaling_record(args, comm, device, rank, world, cfg)
        torch.cuda.empty_cache()
        verify = multi_rank_parity(comm, device, rank, world, wl.plan)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    roof = roofline(stages, counts, frames, cfg, peaks, args)

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, secs, cres = time_cpu(2, args.cpu_sample_frames // 2)
        cpu = {"value": v, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
               "sample": f"2 trajectories x {args.cpu_sample_frames // 2} frames of the same workload "
                         f"(K={N_STATES}, {KMEANS_ITERS} Lloyd iterations, reversible MLE capped at {CPU_MLE_ITER_CAP} "
                         f"iterations), {secs:.1f} s of oracle work on {cpu_threads()} threads; {CPU_IMPL_NOTE}",
               "stages_s": {k: round(v2, 3) for k, v2 in cres.stage_seconds.items()},
               **cpu_cost_split(cres.stage_seconds, cres.n_frames)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 features / 3xTF32-or-fp32 Gram with fp64 flush / f64 MSM",
        "data": "synthetic", "config": workload_config(args, frames),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms, "h2d_ceiling_gbs": ceiling,
                "h2d_bound_ms": h2d / (ceiling["per_gpu_gbs"] * 1e9) * 1e3,
                "returns": "labels, T, pi, eigenvalues, timescales, MLE info (numpy)"},
        "e2e_roofline": end_to_end_roofline(stages, frames, cfg, peaks, world),
        "strong": strong, "multi_rank_parity": verify, "parity": PARITY_NOTE,
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        "stages_ms": stages, "kmeans_iters_run": int(res.kmeans_iters),
        "tica_rank_sweeps": [int(v) for v in res.tica.rank_dev.tolist()],
        "mle_phase_cycles": debug_counters(), "kmeans_role_cycles": kmeans_debug_counters(), "tica_phase_cycles": tica_debug_counters(),
        "mle_iters": int(res.mle_info[0].item()), "lanczos_steps": int(res.extra["eig_info"].reshape(-1)[0].item()),
        "timescales": [None if not np.isfinite(t) else float(t)
                                                                for t in (res.timescales if res.timescales is not None else [])],
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def debug_counters():
    import ctypes

    from pmarlo_b200 import _lib

    buf = (ctypes.c_int64 * 8)()
    _lib.check(_lib.lib().pmb_debug_counters(ctypes.cast(buf, ctypes.c_void_p)), "pmb_debug_counters")
    return list(buf)


def tica_debug_counters():
    import ctypes

    from pmarlo_b200 import _lib

    buf = (ctypes.c_int64 * 8)()
    _lib.check(_lib.lib().pmb_debug_counters_tica(ctypes.cast(buf, ctypes.c_void_p)), "pmb_debug_counters_tica")
    return list(buf)


def kmeans_debug_counters():
    import ctypes

    from pmarlo_b200 import _lib

    buf = (ctypes.c_int64 * 16)()
    _lib.check(_lib.lib().pmb_debug_counters_kmeans(ctypes.cast(buf, ctypes.c_void_p)), "pmb_debug_counters_kmeans")
    return list(buf)


def roofline(stages, counts, frames, cfg, peaks, args):
    """Roofline of the dominant kernel (largest share of the step), from the CUDA-event time of its
    launches inside the timed region, plus the same figures for every other per-frame kernel under
    ``all``.  Algorithmic work per frame: DESIGN.md section 3 / SURVEY.md 8d.  ``traffic`` is the DRAM
    traffic per launch measured by ncu (profiles/ncu_traffic.json: bytes per frame of one ``--set full``
    capture, scaled to this run's frames)."""
    kernels_alg = {
        # name: (bound, bytes or flops per frame per launch)
        "featurize": ("hbm", 12 * 3 * N_RES + 4 * 256),
        "col_moments": ("hbm", 4 * 256),
        "gram": ("tensor", 2 * 256 * 256),            # one d x d rank-1 update per frame per launch
        "project": ("hbm", 4 * 256 + 4 * cfg.tica_dim),
        "kmeans_assign": ("tensor", 2 * cfg.tica_dim * cfg.n_states),
        "count": ("hbm", 4),
    }
    traffic_pf = {}
    tp = ROOT / "profiles" / "ncu_traffic.json"
    if tp.exists():
        try:
            traffic_pf = json.loads(tp.read_text()).get("dram_bytes_per_frame", {})
        except Exception:
            traffic_pf = {}
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tf32_peak = float(peaks.get("bf16_tflops_sustained", 1400.0)) / 2.0
    step_ms = max(1e-9, sum(v for k, v in stages.items() if k in TOP_LEVEL_STAGES))

    def one(name):
        bound, per_frame = kernels_alg[name]
        n_launch = max(1, counts.get(name, 1))
        ms_per_launch = stages.get(name, 0.0) * args.steps / n_launch
        work = per_frame * frames
        if bound == "hbm":
            peak, unit = hbm_peak, "GB/s"
            achieved = work / (ms_per_launch * 1e-3) / 1e9 if ms_per_launch else 0.0
            src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6.65 TB/s (of fallback)"
        else:
            # fp32-exact tensor work is issued as TF32: peak = measured bf16 / 2
            peak, unit = tf32_peak, "TFLOP/s"
            achieved = work / (ms_per_launch * 1e-3) / 1e12 if ms_per_launch else 0.0
            src = ("MEASURED_PEAKS.json bf16_tflops_sustained / 2 (TF32 rate, of measured; kernel timed inside a "
                   "long step)" if peaks else "fallback 1.4 PFLOP/s / 2 (of fallback)")
        tr = traffic_pf.get(name)
        return {"kernel": name, "bound": bound, "achieved": achieved, "peak": peak, "unit": unit,
                "frac": achieved / peak if peak else None,
                "traffic": (float(tr) * frames) if tr is not None else None,
                "ms_per_launch": ms_per_launch, "launches_per_step": n_launch / args.steps,
                "algorithmic_per_frame": per_frame, "peak_source": src,
                "share_of_step": stages.get(name, 0.0) / step_ms}

    top = max(kernels_alg, key=lambda k: stages.get(k, 0.0))
    out = one(top)
    out["all"] = {k: {kk: vv for kk, vv in one(k).items() if kk in ("bound", "achieved", "unit", "frac", "ms_per_launch",
                                                                     "share_of_step", "traffic")}
                  for k in kernels_alg}
    return out


TOP_LEVEL_STAGES = ("featurize", "tica_fit", "project", "kmeans", "count", "mle", "eig")

PARITY_NOTE = ("partial: integer stages (labels given centres, counts) bit-exact against reference-generated goldens; "
               "floating stages (dihedrals, TICA, reversible MLE, eigenvalues) within 1e-6 of the oracle restatement, "
               "which is unpinned against deeptime/mdtraj binaries (absent from the image)")


# ----------------------------------------------------------------------------- config C5 (large-state stress)
def synth_c5_device(n_traj, frames_per_traj, D, K, device, seed, hop=1e-3, sigma_c=5.0, chunk_traj=40):
    """Y = 64-dim mixture around K seeded centres (sigma 5), unit noise, a trajectory hops to a
    new random centre with probability 1e-3 per frame.  Returns (Y (N, D) float32, true centres (K, D))."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    cen = sigma_c * torch.randn((K, D), generator=g, device=device, dtype=torch.float32)
    Y = torch.empty((n_traj * frames_per_traj, D), dtype=torch.float32, device=device)
    for t0 in range(0, n_traj, chunk_traj):
        nt = min(chunk_traj, n_traj - t0)
        hops = torch.rand((nt, frames_per_traj), generator=g, device=device) < hop
        hops[:, 0] = True
        seg = torch.cumsum(hops.to(torch.int32), dim=1) - 1                      # segment number inside the trajectory
        seg_cen = torch.randint(0, K, (nt, int(seg.max().item()) + 1), generator=g, device=device)
        idx = torch.gather(seg_cen, 1, seg.to(torch.int64)).reshape(-1)
        out = Y[t0 * frames_per_traj:(t0 + nt) * frames_per_traj]
        torch.index_select(cen, 0, idx, out=out)
        out += torch.randn(out.shape, generator=g, device=device, dtype=torch.float32)
        del hops, seg, seg_cen, idx
    return Y, cen


def gpu_arm_c5(args):
    """BASELINE configs[4]: 50 M frames x 64 dims, K = 5000, 10 Lloyd iterations, 5000 x 5000 counts at lag 20,
    reversible MLE, top-20 eigenvalues -- one GPU (12.8 GB of Y).  Prints one JSON line in the bench contract;
    the roofline entries are K6 on the tensor roofline (2 D K flop per frame and iteration; issued as fp16-kind
    MMAs) and K8 against HBM (8 K^2 bytes per iteration: 200 MB, no longer L2-resident)."""
    import torch

    from pmarlo_b200 import _lib
    from pmarlo_b200.distributed import Comm
    from pmarlo_b200.pipeline import PipelineConfig, StageTimer, run_pipeline
    from pmarlo_b200.shards import Segments

    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(device)
    _lib.load()
    D, K, lag = 64, 5000, 20
    n_traj = max(1, args.frames_per_gpu // FRAMES_PER_TRAJ) if args.frames_per_gpu != 10_000_000 else 400
    frames = n_traj * FRAMES_PER_TRAJ
    Y, cen = synth_c5_device(n_traj, FRAMES_PER_TRAJ, D, K, device, seed=5)
    segs = Segments.from_lengths([FRAMES_PER_TRAJ] * n_traj)
    cfg = PipelineConfig(tica_dim=0, n_states=K, kmeans_max_iter=10, kmeans_tolerance=None, msm_lag=lag, n_timescales=19,
                         mle_maxerr=1e-12, mle_maxiter=args.c5_mle_maxiter, seed=5)
    init = (cen + 0.5 * torch.randn(cen.shape, device=device, generator=torch.Generator(device=device).manual_seed(6))).to(torch.float64)
    bufs: dict = {}
    res = None
    for _ in range(max(1, min(args.warmup, 3))):
        res = run_pipeline(None, segs, None, cfg, Comm(), features=Y, initial_centers=init, buffers=bufs)
    torch.cuda.synchronize()
    sampler = ClockSampler(device.index or 0)
    sampler.start()
    timer = StageTimer(True)
    l0 = _lib.launch_count()
    steps = max(1, min(args.steps, 3))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        res = run_pipeline(None, segs, None, cfg, Comm(), features=Y, initial_centers=init, timer=timer, buffers=bufs)
    b.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = a.elapsed_time(b) / steps
    stages = {k: v / steps for k, v in timer.totals_ms().items()}
    counts = timer.counts()
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    bf16 = float(peaks.get("bf16_tflops_sustained", 1400.0))
    n_assign = max(1, counts.get("kmeans_assign", 1)) / steps
    ms_assign = stages.get("kmeans_assign", 0.0) / n_assign
    tf = 2.0 * D * K * frames / (ms_assign * 1e-3) / 1e12 if ms_assign else 0.0
    mle_iters = int(res.mle_info[0].item())
    ms_mle_iter = stages.get("mle", 0.0) / max(1, mle_iters)
    mle_gbs = 8.0 * K * K / (ms_mle_iter * 1e-3) / 1e9 if ms_mle_iter else 0.0
    # properties that hold at any size
    C = res.counts
    props = {"counts_total_ok": bool(int(C.sum().item()) == segs.n_pairs(lag)),
             "rows_stochastic": float((res.T.sum(dim=1) - 1.0).abs().max().item()),
             "stationarity": float((res.pi @ res.T - res.pi).abs().max().item()),
             "detailed_balance": float((res.pi[:, None] * res.T - (res.pi[:, None] * res.T).T).abs().max().item()),
             "mle_converged": bool(int(res.mle_info[1].item()) == 1), "mle_iters": mle_iters,
             "lanczos_steps": int(res.extra["eig_info"].reshape(-1)[0].item())}
    line = {
        "metric": METRIC, "value": frames / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 coordinates / fp16-kind split-K score GEMM with fp64 re-check / f64 MSM", "data": "synthetic",
        "config": {"workload": "C5 large-state stress: 50 M frames x 64 dims, k-means K=5000 x 10 Lloyd iterations, 5000x5000 counts "
                               "at lag 20, reversible MLE (1e-12), top-20 eigenvalues", "frames_per_gpu": frames, "n_states": K,
                   "n_dims": D, "kmeans_iters": 10, "msm_lag": lag,
                   "l2_policy": "Y (256 B per frame, 12.8 GB) and the 200 MB count / transition matrices exceed the 126 MB L2"},
        "gpu_launches": int(_lib.launch_count() - l0), "clocks": clocks, "stages_ms": stages,
        "roofline": {"kernel": "kmeans_assign", "bound": "tensor", "achieved": tf, "peak": bf16 / 2.0, "unit": "TFLOP/s",
                     "frac": tf / (bf16 / 2.0), "frac_of_fp16_peak": tf / bf16, "traffic": None, "ms_per_launch": ms_assign,
                     "algorithmic_per_frame": 2 * D * K,
                     "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (same TF32-rate denominator as C4; the kernel "
                                    "issues 3 fp16 MMAs per algorithmic product, so 1/3 of the fp16 peak is its ceiling)",
                     "all": {"mle": {"bound": "hbm", "achieved": mle_gbs, "peak": hbm, "unit": "GB/s", "frac": mle_gbs / hbm,
                                     "ms_per_iteration": ms_mle_iter, "algorithmic_bytes_per_iteration": 8 * K * K}}},
        "properties": props, "e2e": None, "cpu_baseline": None, "parity": PARITY_NOTE,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames-per-gpu", type=int, default=10_000_000,
                    help="frames of this rank's shard (default: the whole of C4, 10 M frames = 80 trajectories "
                         "x 125 000, on every GPU; weak scaling)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample-frames", type=int, default=CPU_SAMPLE_FRAMES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the strong-scaling record and the multi-rank parity check (N > 1)")
    ap.add_argument("--gram-impl", type=int, default=0)
    ap.add_argument("--config", default="C4", choices=["C4", "C5"],
                    help="C4 (default, the metric's configuration) or the large-state stress configuration C5 (N = 1)")
    ap.add_argument("--c5-mle-maxiter", type=int, default=200_000)
    ap.add_argument("--profile-range", action="store_true",
                    help="bracket the timed steps with cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        print("note: fewer than 3 warm-up steps", file=sys.stderr)
    if args.impl == "reference":
        return reference_arm(args)
    if args.config == "C5":
        return gpu_arm_c5(args)
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
