"""Time and accuracy of the TICA projection (K5) on the C4 shape.  python tools/prj_bench.py [n] [d] [m] [reps]
PMB_PRJ_VARIANT selects the launch configuration of project_warp_kernel (see project.cu)."""
import os
import sys

import torch

sys.path.insert(0, ".")
from pmarlo_b200 import kernels  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 256
m = int(sys.argv[3]) if len(sys.argv) > 3 else 10
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 6
dev = kernels.require_cuda()
g = torch.Generator(device=dev).manual_seed(0)
X = torch.randn((n, d), generator=g, device=dev, dtype=torch.float32) * 0.7 + 0.3
X[5, 7] = float("nan")
a = torch.full((d,), 0.3, dtype=torch.float64, device=dev) + torch.randn(d, generator=g, device=dev, dtype=torch.float64) * 1e-3
fill = a + 0.01
W = torch.randn((d, m), generator=g, device=dev, dtype=torch.float64) / d ** 0.5
Y = torch.empty((n, m), dtype=torch.float32, device=dev)
ts = []
for it in range(reps):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    kernels.project(X, a, fill, W, out=Y)
    e.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
ns = min(n, 200_000)
Xs = X[:ns].double()
Xs = torch.where(torch.isnan(Xs), fill.expand_as(Xs), Xs)
ref = (Xs - a) @ W
err = (Y[:ns].double() - ref).abs().max().item()
t = sorted(ts[1:])[len(ts[1:]) // 2]
print(f"variant={os.environ.get('PMB_PRJ_VARIANT', '0')} n={n} d={d} m={m}: {t:.3f} ms  {n * (d + m) * 4 / t / 1e6:.0f} GB/s  max err vs fp64 {err:.2e}")
