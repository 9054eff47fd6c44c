"""Would lower bounds at (32-frame group x 32-centre block) granularity let the K6 epilogue skip its TMEM reads?
Probe only (torch): Lloyd on the bench's TICA coordinates; per iteration the fraction of (group, block) pairs whose
carried lower bound  l[g][b] (distance of the group's nearest frame-centre pair in the block minus the distance to
the frame's own centre, decremented by the block's and the hinted centres' drift)  stays positive, i.e. whose 32x32
scores need not be read; and the fraction of (128-frame tile, 128-centre chunk) pairs with all 16 of theirs skipped
(whose MMAs need not be issued).  `sort` = centres ordered along a Morton curve of their first coordinates first."""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from pmarlo_b200.distributed import Comm  # noqa: E402
from pmarlo_b200.pipeline import run_pipeline, seeded_initial_centers  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
order = sys.argv[2] if len(sys.argv) > 2 else "asis"
BS = int(sys.argv[3]) if len(sys.argv) > 3 else 32     # centres per block (32: a TMEM load, 128: a whole chunk)
dev = torch.device("cuda", 0)
n_traj = max(1, frames // bench.FRAMES_PER_TRAJ)
wl = bench.make_workload(n_traj, frames // n_traj, dev, seed=4000)
cfg = bench.bench_config()
res = run_pipeline(wl.xyz, wl.segs, wl.plan, cfg)
Y = res.Y.float()
del wl
n, D = Y.shape
n = n // 128 * 128
Y = Y[:n]
K = cfg.n_states
C = seeded_initial_centers(res.Y, K, cfg.seed, Comm()).float()
if order == "sort":
    # crude spatial order: sort by the first principal coordinates' interleaved bits
    q = ((C[:, :3] - C[:, :3].min(0).values) / (C[:, :3].max(0).values - C[:, :3].min(0).values) * 15.99).long()
    key = torch.zeros(K, dtype=torch.long, device=dev)
    for bit in range(4):
        for d in range(3):
            key |= ((q[:, d] >> bit) & 1) << (3 * bit + d)
    C = C[torch.argsort(key)]
Kp = (K + 127) // 128 * 128
nb, ng = Kp // BS, n // 32
margin = 1e-3


def block_min_dist(Y, C, chunk=262144):
    """(n, nb) distance to the nearest centre of each block, and labels."""
    out = torch.empty((Y.shape[0], nb), device=dev)
    lab = torch.empty((Y.shape[0],), dtype=torch.long, device=dev)
    Cp = torch.cat([C, torch.full((Kp - K, D), 1e6, device=dev)])
    for s in range(0, Y.shape[0], chunk):
        d = torch.cdist(Y[s:s + chunk], Cp)
        out[s:s + chunk] = d.view(-1, nb, BS).min(dim=2).values
        lab[s:s + chunk] = d.argmin(dim=1)
    return out, lab


l = None
for it in range(cfg.kmeans_max_iter + 1):
    bd, lab_new = block_min_dist(Y, C)
    if l is not None:
        # carried bounds after this iteration's centre update (delta known from the previous update)
        delta_b = torch.cat([delta, torch.zeros(Kp - K, device=dev)]).view(nb, BS).max(dim=1).values      # (nb,)
        d_hint = delta[lab].view(ng, 32).max(dim=1).values                                                 # (ng,)
        l = l - delta_b[None, :] - d_hint[:, None]
        skip = l > margin
        # soundness: a skipped block must not hold the new label of any frame of the group
        lab_blk = (lab_new // BS).view(ng, 32)
        bad = skip.gather(1, lab_blk).any()
        frac = float(skip.float().mean())
        chunk_skip = skip.view(ng // 4, 4, nb * BS // 128, 128 // BS).all(dim=3).all(dim=1)          # (tiles, chunks)
        print(f"iter {it:2d}: blocks skipped {frac:.3f}  chunks (MMA) skipped {float(chunk_skip.float().mean()):.3f}  "
              f"mean drift {float(delta.mean()):.4f} max {float(delta.max()):.3f}  unsound={bool(bad)}", flush=True)
        # read blocks get their exact slack, skipped ones keep the bound
        d_own = (Y - C[lab]).norm(dim=1)                     # distance to the HINTED centre under the new centres
        exact = (bd - d_own[:, None]).view(ng, 32, nb).min(dim=1).values
        l = torch.where(skip, l, exact)
    else:
        d_best = bd.min(dim=1).values
        l = (bd - d_best[:, None]).view(ng, 32, nb).min(dim=1).values
    lab = lab_new
    sums = torch.zeros((K, D), device=dev, dtype=torch.float64).index_add_(0, lab, Y.double())
    cnt = torch.bincount(lab, minlength=K).double()
    Cn = torch.where(cnt[:, None] > 0, sums / cnt.clamp(min=1)[:, None], C.double()).float()
    delta = (Cn - C).norm(dim=1)
    C = Cn
