"""tcgen05 plumbing in isolation: descriptor encodings, TMEM round trip, TF32 operand handling."""

from __future__ import annotations

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def tf32_trunc(x: np.ndarray) -> np.ndarray:
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("N,Kdim", [(32, 8), (128, 16), (256, 16), (256, 64), (96, 32)])
def test_tile_product_exact_on_tf32_inputs(mode, N, Kdim):
    """Inputs representable in TF32 -> products exact in fp32; only the fp32 accumulation order differs."""
    from pmarlo_b200 import kernels

    rng = np.random.default_rng(N * 100 + Kdim + mode)
    A = tf32_trunc(rng.normal(size=(128, Kdim)))
    B = tf32_trunc(rng.normal(size=(N, Kdim)))
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    dev = torch.device("cuda")
    if mode == 0:
        D = kernels.tc_selftest(torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev), 0)
    else:
        D = kernels.tc_selftest(torch.from_numpy(np.ascontiguousarray(A.T)).to(dev),
                                torch.from_numpy(np.ascontiguousarray(B.T)).to(dev), 1)
    got = D.cpu().numpy().astype(np.float64)
    bound = 2.0 ** -21 * (np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T)
    assert np.all(np.abs(got - ref) <= bound + 1e-30), float(np.max(np.abs(got - ref) / (bound + 1e-30)))


def test_operand_rounding_mode_is_truncation_or_rounding(tmp_path):
    """Documents how the tensor core narrows fp32 operands to TF32 (drives the 3xTF32 split design)."""
    from pmarlo_b200 import kernels

    rng = np.random.default_rng(1)
    A = rng.normal(size=(128, 16)).astype(np.float32)
    B = rng.normal(size=(64, 16)).astype(np.float32)
    dev = torch.device("cuda")
    got = kernels.tc_selftest(torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev), 0).cpu().numpy().astype(np.float64)
    ref_trunc = tf32_trunc(A).astype(np.float64) @ tf32_trunc(B).astype(np.float64).T
    full = A.astype(np.float64) @ B.astype(np.float64).T
    e_trunc = float(np.max(np.abs(got - ref_trunc)))
    e_full = float(np.max(np.abs(got - full)))
    print(f"tf32 operand handling: |got - truncated-input product| = {e_trunc:.3e}, |got - fp32-input product| = {e_full:.3e}")
    # either way the result must be within TF32 accuracy of the full-precision product
    assert e_full < 16 * 2.0 ** -10 * 4.0
