"""Debug: run pmb_eig_rev_topk with an inspectable workspace and print the recurrence / orthogonality per step."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pmarlo_b200 import _lib, kernels  # noqa: E402
from pmarlo_b200._lib import check, ptr, stream_handle  # noqa: E402

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
check(L.pmb_eig_rev_topk(ptr(T), ptr(pi), K, k, 1, 0, ptr(ev), ptr(info), ptr(ws), ws.numel() * 8, stream_handle(dev)), "eig")
torch.cuda.synchronize()
m = max(10 * k, 200); m = min(m, K, 1024)
w = ws.cpu().numpy()
V = w[: (m + 1) * K].reshape(m + 1, K)
off = (m + 1) * K + 4 * K
h = w[off: off + 2 * (m + 1)]; off += 2 * (m + 1)
alpha = w[off: off + m]; off += m
beta = w[off: off + m]; off += m
steps = int(info[0].item())
print("steps", steps, "ok", int(info[1].item()), "ev", ev.cpu().numpy()[:4])
G = V[:steps] @ V[:steps].T - np.eye(steps)
for j in range(0, steps, 8):
    print(j, "alpha", alpha[j:j + 8].round(4), "beta", beta[j:j + 8].round(4), "orth(row max)", np.abs(G[j:j + 8, :]).max(axis=1).max())
