"""Accuracy of the two Gram paths (SIMT fp32-with-fp64-folds, tcgen05 split) on the bench's own features against
an fp64 torch reference: symmetrised TICA covariances C00 / C0t, max |error| / max |entry|."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from pmarlo_b200.features import featurize_device  # noqa: E402
from pmarlo_b200.reduction import TICA  # noqa: E402
from pmarlo_b200.shards import Segments  # noqa: E402

dev = torch.device("cuda")
top, plan, _ = bench.c4_plan()
for n_traj, fpt, rho in ((4, 20000, 0.995), (8, 125000, 0.9995), (2, 40000, 0.98)):
    xyz = bench.synth_xyz_device(n_traj, fpt, dev, seed=9000, rho=rho)
    X = featurize_device(xyz, plan)
    segs = Segments.from_lengths([fpt] * n_traj)
    F = X.to(torch.float64)
    mu, sd = F.mean(dim=0), F.std(dim=0, unbiased=False)
    Z = ((F - mu) / torch.where(sd > 0, sd, torch.ones_like(sd))).view(n_traj, fpt, -1)
    lag = 20
    X0, Xt = Z[:, :-lag].reshape(-1, Z.shape[2]), Z[:, lag:].reshape(-1, Z.shape[2])
    m = 0.5 * (X0.mean(dim=0) + Xt.mean(dim=0))
    X0c, Xtc = X0 - m, Xt - m
    T = float(X0.shape[0])
    C00 = (X0c.T @ X0c + Xtc.T @ Xtc) / (2 * T)
    C0t = (X0c.T @ Xtc + Xtc.T @ X0c) / (2 * T)
    zmax = float(Z.abs().max().item())
    for impl in (1, 2, 5):
        model = TICA(lag, 10, preprocess="standard", gram_impl=impl).fit_device(X, segs)
        e00 = float((model.C00 - C00).abs().max().item() / C00.abs().max().item())
        e0t = float((model.C0t - C0t).abs().max().item() / C0t.abs().max().item())
        # where is the worst element?
        d = (model.C00 - C00).abs()
        ij = int(d.argmax().item())
        print(f"n={n_traj*fpt} rho={rho} impl={impl}: C00_rel={e00:.2e} C0t_rel={e0t:.2e} worst=({ij // 256},{ij % 256}) max|z|={zmax:.1f}", flush=True)
    del xyz, X, F, Z
