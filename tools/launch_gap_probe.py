"""Per-call CUDA-event time of pmb_tica_solve (cooperative kernel) in a tight loop, to look for launch gaps."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pmarlo_b200 import kernels  # noqa: E402

dev = kernels.require_cuda()
g = torch.Generator(device=dev).manual_seed(0)
d = 256
A = torch.randn((4 * d, d), generator=g, device=dev, dtype=torch.float64)
B = torch.roll(A, 1, 0)
C00 = (A.T @ A) / A.shape[0]
C0t = 0.5 * (A.T @ B + B.T @ A) / A.shape[0] * 0.9
big = torch.randn((8192, 8192), device=dev)


def run(label, pre, n=150):
    ts, wall = [], []
    for i in range(n):
        pre()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        kernels.tica_solve(C00, C0t)
        b.record()
        torch.cuda.synchronize()
        wall.append((time.perf_counter() - t0) * 1e3)
        ts.append(a.elapsed_time(b))
    ts = np.array(ts)
    print(f"{label}: median {np.median(ts):.2f} ms, p90 {np.percentile(ts, 90):.2f}, max {ts.max():.2f}, "
          f"n>1.5x median {(ts > 1.5 * np.median(ts)).sum()}/{n}; outliers at {np.flatnonzero(ts > 1.5 * np.median(ts))[:20].tolist()} "
          f"values {np.round(ts[ts > 1.5 * np.median(ts)][:10], 1).tolist()}", flush=True)


run("back-to-back", lambda: None)
run("after an idle gap of 20 ms", lambda: time.sleep(0.02))
run("after a 10 ms tensor-heavy matmul", lambda: torch.matmul(big, big))
run("after matmul + sync", lambda: (torch.matmul(big, big), torch.cuda.synchronize()))
