"""Time and accuracy of the top-k eigenvalue solve (K9) on a reversible K x K matrix.
python tools/eig_bench.py [K] [k] [reps]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pmarlo_b200 import kernels  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 6
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
dev = kernels.require_cuda()
rng = np.random.default_rng(0)
# metastable reversible chain: 5 blocks + weak coupling
C = rng.random((K, K)) * 0.02
b = K // 5
for i in range(5):
    C[i * b:(i + 1) * b, i * b:(i + 1) * b] += rng.random((b, b))
C = C + C.T
T = torch.from_numpy(C / C.sum(1, keepdims=True)).to(dev)
pi = torch.from_numpy(C.sum(1) / C.sum()).to(dev)
ref = np.sort(np.abs(np.linalg.eigvals(T.cpu().numpy())))[::-1][:k]
ts = []
for it in range(reps):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ev, info = kernels.eig_rev_topk(T, pi, k)
    e.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(e))
    got = np.sort(np.abs(ev.cpu().numpy()))[::-1]
    print(f"  rep {it}: steps={int(info[0].item())} ok={int(info[1].item())} err={np.max(np.abs(got - ref)):.2e} top={got[:4]}")
print(f"K={K} k={k}: {np.median(ts[1:]):.3f} ms; ref top = {ref[:4]}")
