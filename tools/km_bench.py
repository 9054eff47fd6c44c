"""Per-launch time of the tensor k-means assignment (K6) at the C4 and C5 shapes, with and without hints and
with the fused Lloyd accumulation; prints the role cycle counters of a -DPMB_KM_PROF build when present.
    python tools/km_bench.py [c4|c5|all]
"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pmarlo_b200 import _lib, kernels  # noqa: E402

import os  # noqa: E402

if os.environ.get("PMB_LIB"):   # experiment builds (e.g. -DPMB_KM_PROF) live outside the package
    _lib.load(os.environ["PMB_LIB"])


def counters():
    buf = (ctypes.c_int64 * 16)()
    _lib.check(_lib.lib().pmb_debug_counters_kmeans(ctypes.cast(buf, ctypes.c_void_p)), "dbg")
    return list(buf)


def run(name, n, D, K, reps=6, sigma_c=1.0):
    dev = kernels.require_cuda()
    g = torch.Generator(device=dev).manual_seed(1)
    if name == "c5":
        cen = sigma_c * 5.0 * torch.randn((K, D), generator=g, device=dev, dtype=torch.float32)
        idx = torch.randint(0, K, (n,), generator=g, device=dev)
        Y = cen[idx] + torch.randn((n, D), generator=g, device=dev, dtype=torch.float32)
        C = cen.double().contiguous()
    elif name == "c4":
        # trajectory-like rows (AR(1), rho = 0.9995, like the bench's TICA coordinates): neighbouring frames
        # share their candidate centres, which is what the block screening of the epilogue exploits
        import bench
        A = (D + 2) // 3
        xyz = bench.synth_xyz_device(n // 125_000, 125_000, dev, seed=7, rho=0.9995, sigma=1.0,
                                     base=torch.zeros((A, 3), device=dev))
        Y = xyz.reshape(xyz.shape[0], -1)[:, :D].contiguous()
        del xyz
        n = int(Y.shape[0])
        C = Y[torch.randperm(n, generator=g, device=dev)[:K]].double().contiguous()
    else:
        Y = torch.randn((n, D), generator=g, device=dev, dtype=torch.float32)
        C = Y[torch.randperm(n, generator=g, device=dev)[:K]].double().contiguous()
    labels = torch.empty((n,), dtype=torch.int32, device=dev)
    sums = torch.zeros((K, D), dtype=torch.float64, device=dev)
    cnt = torch.zeros((K,), dtype=torch.int64, device=dev)
    inertia = torch.zeros((1,), dtype=torch.float64, device=dev)
    nre = torch.zeros((1,), dtype=torch.int64, device=dev)
    for mode in ("cold", "cold-subsampled-hints", "hints", "hints+accumulate"):
        ts = []
        for it in range(reps):
            nre.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if mode == "cold":
                kernels.kmeans_assign(Y, C, labels=labels, impl=2, n_rechecked=nre)
            elif mode == "cold-subsampled-hints":
                from pmarlo_b200.clustering import subsampled_hints
                kernels.kmeans_assign(Y, C, labels=labels, impl=2, hints=subsampled_hints(Y, C), n_rechecked=nre)
            elif mode == "hints":
                kernels.kmeans_assign(Y, C, labels=labels, impl=2, hints=labels, n_rechecked=nre)
            else:
                sums.zero_(); cnt.zero_(); inertia.zero_()
                kernels.kmeans_assign(Y, C, labels=labels, impl=2, hints=labels, sums=sums, counts=cnt, inertia=inertia,
                                      n_rechecked=nre)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.median(ts[1:]))
        flops = 2.0 * D * K * n
        print(f"{name} n={n} D={D} K={K} {mode:18s} ms/launch={ms:.3f} cycles/tile@1.965GHz={ms*1e-3*1.965e9/(n/128/148):.0f} "
              f"algorithmic TFLOP/s={flops/ms/1e9:.1f} rechecked={int(nre.item())/n:.2e} counters={counters()[:12]}", flush=True)
    # spot check against the fp64 argmin
    sl = slice(0, 4096)
    d = ((Y[sl].double()[:, None, :] - C[None, :, :]) ** 2).sum(-1) if K * D <= 20000 else None
    if d is not None:
        ref = d.argmin(dim=1).to(torch.int32)
        print(f"  spot check: {(ref != labels[sl]).sum().item()} of 4096 labels differ", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("c4", "all"):
        run("c4", 10_000_000, 10, 1000)
    if which in ("rand", "all"):
        run("rand", 10_000_000, 10, 1000)
    if which in ("c5", "all"):
        run("c5", 2_000_000, 64, 5000)
