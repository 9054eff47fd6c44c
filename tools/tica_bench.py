"""Time of the TICA eigen-solve (K4) on a d x d problem shaped like the bench's (AR(1) features).
python tools/tica_bench.py [d] [reps];  PMB_TICA_CLUSTER=0 forces the cooperative (148-CTA) launch."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pmarlo_b200 import kernels  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = kernels.require_cuda()
rng = np.random.default_rng(0)
n, lag = 20000, 10
rho = rng.uniform(0.5, 0.9995, d)
x = np.zeros((n, d))
e = rng.standard_normal((n, d))
for t in range(1, n):
    x[t] = rho * x[t - 1] + np.sqrt(1 - rho ** 2) * e[t]
x = x @ (np.eye(d) + 0.1 * rng.standard_normal((d, d)))
x -= x.mean(0)
C00 = torch.from_numpy(x.T @ x / n).to(dev)
C0t = torch.from_numpy(0.5 * (x[:-lag].T @ x[lag:] + x[lag:].T @ x[:-lag]) / (n - lag)).to(dev)
ts = []
for it in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ev, V, rank = kernels.tica_solve(C00, C0t)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
import scipy.linalg as sl
ref = np.sort(sl.eigh(C0t.cpu().numpy(), C00.cpu().numpy(), eigvals_only=True))[::-1]
got = np.sort(ev.cpu().numpy())[::-1]
print(f"cluster={os.environ.get('PMB_TICA_CLUSTER', '1')} chol={os.environ.get('PMB_TICA_CHOL', '1')} d={d}: {np.median(ts[1:]):.3f} ms  rank={int(rank[0])} sweeps={int(rank[1])}+{int(rank[2])}  "
      f"max |ev - scipy| = {np.abs(got - ref).max():.2e}  checksum={float(ev.sum()):.17g} {float(V.abs().sum()):.17g}")
