"""Phase cycle counters of lanczos_kernel (a -DPMB_LAN_PROF build, PMB_LIB=...): python tools/eig_prof.py K k"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pmarlo_b200 import _lib, kernels  # noqa: E402
from pmarlo_b200._lib import check, ptr, stream_handle  # noqa: E402

if os.environ.get("PMB_LIB"):
    _lib.load(os.environ["PMB_LIB"])
K = int(sys.argv[1]); k = int(sys.argv[2])
dev = kernels.require_cuda()
rng = np.random.default_rng(0)
C = rng.random((K, K)) * 0.02
b = K // 5
for i in range(5):
    C[i * b:(i + 1) * b, i * b:(i + 1) * b] += rng.random((b, b))
C = C + C.T
T = torch.from_numpy(C / C.sum(1, keepdims=True)).to(dev)
pi = torch.from_numpy(C.sum(1) / C.sum()).to(dev)
L = _lib.lib()
nb = L.pmb_eig_rev_topk_ws_bytes(K, k, 1, 0)
ws = torch.zeros(nb // 8 + 8, dtype=torch.float64, device=dev)
ev = torch.empty(k, dtype=torch.float64, device=dev)
info = torch.zeros(2, dtype=torch.int64, device=dev)
for rep in range(3):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    check(L.pmb_eig_rev_topk(ptr(T), ptr(pi), K, k, 1, 0, ptr(ev), ptr(info), ptr(ws), ws.numel() * 8, stream_handle(dev)), "eig")
    e.record()
    torch.cuda.synchronize()
m = min(max(10 * k, 200), K, 1024)
off = (m + 1) * K + 4 * K + 2 * (m + 1) + 2 * m + 2 * m
prof = ws[off + 256: off + 264].cpu().numpy()
steps = int(info[0].item())
names = ["A matvec", "sync1", "B dots", "sync2", "C update", "sync3", "norm+check", "tail (Ritz values)"]
print(f"K={K} k={k} steps={steps} {a.elapsed_time(e):.3f} ms")
for n_, v in zip(names, prof):
    print(f"  {n_:20s} {v / 1.965e3:9.1f} us total  {v / 1.965e3 / steps:7.2f} us/step")
