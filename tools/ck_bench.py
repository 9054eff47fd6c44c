"""Times the Chapman-Kolmogorov test (pmarlo_b200.ck.run_ck) and its compaction kernel on a label shard
already in HBM.  python tools/ck_bench.py [n_frames] -> one JSON line."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pmarlo_b200 import ck, kernels  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
K, n_traj = 1000, 80
dev = kernels.require_cuda()
g = torch.Generator(device=dev).manual_seed(5)
# a lazy walk on K states: stay w.p. 0.8, else jump to a near neighbour; states >= 950 are rare
step = torch.randint(-3, 4, (n,), generator=g, device=dev)
stay = torch.rand((n,), generator=g, device=dev) < 0.8
walk = torch.cumsum(torch.where(stay, torch.zeros_like(step), step), 0)
labels = (walk.remainder(950)).to(torch.int32)
rare = torch.rand((n,), generator=g, device=dev) < 1e-5
labels = torch.where(rare, torch.randint(950, K, (n,), generator=g, device=dev, dtype=torch.int32), labels)
off = torch.arange(0, n + 1, n // n_traj, device=dev, dtype=torch.int64)
off[-1] = n
shard = ck.LabelShard(labels.contiguous(), off)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


lut = torch.arange(K, dtype=torch.int32, device=dev)
lut[::3] = -1
ms_compact = timed(lambda: kernels.relabel_compact(shard.labels, shard.offsets, lut))
kept = int(kernels.relabel_compact(shard.labels, shard.offsets, lut)[1][-1].item())
ms_count = timed(lambda: kernels.count_lagged(shard.labels, shard.offsets, K, 20))
t0 = time.perf_counter()
res = ck.run_ck(shard, 20, None, min_trans=50, top_n_micro=50)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
print(json.dumps({
    "n_frames": n, "K": K,
    "relabel_compact_ms": ms_compact, "relabel_compact_GBps": (8 * n + 4 * kept) / ms_compact / 1e6,
    "count_lagged_ms": ms_count,
    "run_ck_wall_s": wall, "run_ck_frames_per_s": n / wall, "mode": res.mode,
    "mse": {str(k): v for k, v in res.mse.items()}, "insufficient_k": res.insufficient_k,
}))
