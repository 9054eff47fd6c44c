"""Per-launch time of the tensor k-means assignment as a function of D (number of MMA k-steps)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pmarlo_b200 import kernels  # noqa: E402

n, K = 5_000_000, 1000
dev = kernels.require_cuda()
g = torch.Generator(device=dev).manual_seed(1)
for D in (2, 4, 5, 7, 10):
    Y = torch.randn((n, D), generator=g, device=dev, dtype=torch.float32)
    C = Y[torch.randperm(n, generator=g, device=dev)[:K]].double().contiguous()
    labels = torch.empty((n,), dtype=torch.int32, device=dev)
    out = kernels.kmeans_assign(Y, C, labels=labels, impl=2)
    ts = []
    for it in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = kernels.kmeans_assign(Y, C, labels=labels, impl=2, hints=labels)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ks = ((3 * D + 2) + 7) // 8
    print(f"D={D:2d} k-steps={ks} ms/launch={np.median(ts):.3f}  cycles/tile={np.median(ts)*1e-3*1.965e9/(n/128/148):.0f}", flush=True)
