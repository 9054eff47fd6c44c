"""How many frames would an exact bound-based Lloyd iteration (Hamerly 2010) have to re-score on the bench
workload?  Probe only (torch), prints the active fraction per iteration."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from pmarlo_b200.distributed import Comm  # noqa: E402
from pmarlo_b200.pipeline import run_pipeline, seeded_initial_centers  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda", 0)
n_traj = max(1, frames // bench.FRAMES_PER_TRAJ)
wl = bench.make_workload(n_traj, frames // n_traj, dev, seed=4000)
cfg = bench.bench_config()
res = run_pipeline(wl.xyz, wl.segs, wl.plan, cfg)
Y = res.Y.float()
del wl
n, D = Y.shape
K = cfg.n_states
C = seeded_initial_centers(res.Y, K, cfg.seed, Comm()).float()


def two_nearest(Y, C, chunk=500_000):
    d1 = torch.empty(n, device=dev); d2 = torch.empty(n, device=dev); a = torch.empty(n, dtype=torch.long, device=dev)
    for s in range(0, n, chunk):
        d = torch.cdist(Y[s:s + chunk], C)
        v, i = torch.topk(d, 2, dim=1, largest=False)
        d1[s:s + chunk], d2[s:s + chunk], a[s:s + chunk] = v[:, 0], v[:, 1], i[:, 0]
    return d1, d2, a


u = l = lab = None
for it in range(cfg.kmeans_max_iter + 1):
    d1, d2, a = two_nearest(Y, C)
    if it == 0:
        active = torch.ones(n, dtype=torch.bool, device=dev)
        stateless = torch.zeros(n, dtype=torch.bool, device=dev)
    else:
        # stateless test: inside half the distance to the nearest other centre
        cc = torch.cdist(C, C); cc.fill_diagonal_(float("inf"))
        s_half = 0.5 * cc.min(dim=1).values
        d_own = (Y - C[lab]).norm(dim=1)
        stateless = d_own <= s_half[lab]
        top2 = torch.topk(delta, 2).values
        arg = int(torch.argmax(delta))
        u = u + delta[lab]
        l = l - torch.where(lab == arg, top2[1], top2[0])
        need = u > l
        u = torch.where(need, d_own, u)          # tighten the upper bound
        active = need & (u > l) & ~stateless
        changed = int((a != lab).sum())
        assert int((a[~active] != lab[~active]).sum()) == 0, "a skipped frame changed its label"
    lab = torch.where(active, a, lab) if lab is not None else a
    u = torch.where(active, d1, u) if u is not None else d1
    l = torch.where(active, d2, l) if l is not None else d2
    sums = torch.zeros((K, D), device=dev, dtype=torch.float64).index_add_(0, lab, Y.double())
    cnt = torch.bincount(lab, minlength=K).double()
    Cn = torch.where(cnt[:, None] > 0, (sums / cnt.clamp(min=1)[:, None]).float(), C)
    delta = (Cn - C).norm(dim=1)
    print(f"iter {it:2d}: active {float(active.float().mean()):.4f}  stateless-skip {float(stateless.float().mean()):.4f}"
          f"  max move {float(delta.max()):.3e}  mean move {float(delta.mean()):.3e}"
          + (f"  changed {changed}" if it else ""), flush=True)
    C = Cn
