"""Per-launch time of the Gram kernels on C4-shaped features (10 M x 256 by default), both modes, with the role
counters of a -DPMB_GH_PROF build when present.   python tools/gram_bench.py [n_frames] [impls...]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pmarlo_b200 import _lib, kernels  # noqa: E402

if os.environ.get("PMB_LIB"):
    _lib.load(os.environ["PMB_LIB"])
from pmarlo_b200.shards import Segments  # noqa: E402


def counters():
    buf = (ctypes.c_int64 * 16)()
    _lib.check(_lib.lib().pmb_debug_counters_gram(ctypes.cast(buf, ctypes.c_void_p)), "dbg")
    return list(buf)


n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
impls = [int(a) for a in sys.argv[2:]] or [5, 2]
dev = kernels.require_cuda()
d, lag, fpt = 256, 20, 125_000
g = torch.Generator(device=dev).manual_seed(3)
X = torch.randn((n, d), generator=g, device=dev, dtype=torch.float32)
X[:, 93] = torch.where(torch.rand((n,), generator=g, device=dev) < 0.002, 12.0, 0.0) + 0.05 * X[:, 93]   # a heavy tail
segs = Segments.from_lengths([fpt] * (n // fpt))
mask = kernels.pair_mask(segs.device(dev), n, lag)
cond = torch.stack([X[:100000].mean(dim=0), 1.0 / X[:100000].std(dim=0)]).contiguous()
for impl in impls:
    for mode in (0, 1):
        ts = []
        for it in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            G = kernels.gram(X, mask, lag, mode, cond, impl=impl)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.median(ts[1:]))
        stages = n / 74 / 16
        print(f"impl={impl} mode={mode} n={n}: {ms:.3f} ms/launch, {ms * 1e-3 * 1.965e9 / stages:.0f} cycles per 16-frame stage and CTA, "
              f"{2.0 * d * d * n / ms / 1e9:.1f} algorithmic TFLOP/s, counters={counters()[:9]}", flush=True)
