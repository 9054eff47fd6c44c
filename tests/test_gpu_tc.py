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


# ----------------------------------------------------------------------------- 16-bit operand layouts
def _idesc_f16(M, N, a_fmt, b_fmt, a_mn, b_mn):
    """kind::f16 instruction descriptor: D = F32, A/B format 0 = F16, 1 = BF16, major bits 15 / 16."""
    return (1 << 4) | (a_fmt << 7) | (b_fmt << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def _img_kmajor16(M16: np.ndarray) -> tuple[np.ndarray, int, int]:
    """(rows, K) 16-bit matrix -> K-major no-swizzle image: 8-row x 16-byte core matrices, adjacent in K 128 B
    apart (lbo), adjacent in M/N sbo = (K / 8) * 128 apart."""
    rows, K = M16.shape
    lbo, sbo = 128, (K // 8) * 128
    img = np.zeros(rows * K, dtype=np.uint16)
    r, k = np.meshgrid(np.arange(rows), np.arange(K), indexing="ij")
    off = (r >> 3) * sbo + (k >> 3) * lbo + (r & 7) * 16 + (k & 7) * 2
    img[off.ravel() // 2] = M16.view(np.uint16).ravel()
    return img.view(np.uint8), lbo, sbo


def _img_mnmajor16_sw128(Mt16: np.ndarray, k_atoms_contiguous: bool) -> tuple[np.ndarray, int, int]:
    """(K, MN) 16-bit matrix (MN contiguous, like a frame-major feature block) -> MN-major 128-byte-swizzle
    image: an atom is 8 k-rows of 128 B (64 elements along MN), the 16-byte chunk index is XOR-ed with the
    k-row index.  Returns (image, stride between atoms along MN, stride between atoms along K)."""
    K, MN = Mt16.shape
    n_mn, n_k = MN // 64, K // 8
    if k_atoms_contiguous:
        s_k, s_mn = 1024, n_k * 1024
    else:
        s_mn, s_k = 1024, n_mn * 1024
    img = np.zeros(K * MN, dtype=np.uint16)
    k, mn = np.meshgrid(np.arange(K), np.arange(MN), indexing="ij")
    chunk = ((mn & 63) >> 3) ^ (k & 7)
    off = (mn >> 6) * s_mn + (k >> 3) * s_k + (k & 7) * 128 + chunk * 16 + (mn & 7) * 2
    img[off.ravel() // 2] = Mt16.view(np.uint16).ravel()
    return img.view(np.uint8), s_mn, s_k


def _to16(x: np.ndarray, fmt: str) -> np.ndarray:
    if fmt == "f16":
        return x.astype(np.float16)
    t = torch.from_numpy(x.astype(np.float32)).to(torch.bfloat16)
    return t.view(torch.int16).numpy().view(np.uint16)


def _from16(x16: np.ndarray, fmt: str) -> np.ndarray:
    if fmt == "f16":
        return x16.astype(np.float64)
    return torch.from_numpy(x16.view(np.int16).copy()).view(torch.bfloat16).to(torch.float64).numpy()


@pytest.mark.parametrize("N,Kdim", [(128, 32), (256, 16), (64, 208)])
def test_f16_kmajor_tile_product(N, Kdim):
    """The K6 operand layout: fp16, K-major, no swizzle, K = 16 per instruction."""
    from pmarlo_b200 import kernels

    rng = np.random.default_rng(N + Kdim)
    A = rng.normal(size=(128, Kdim)).astype(np.float16)
    B = rng.normal(size=(N, Kdim)).astype(np.float16)
    a_img, lbo, sbo = _img_kmajor16(A)
    b_img, _, _ = _img_kmajor16(B)
    dev = torch.device("cuda")
    D = kernels.tc_selftest_raw(torch.from_numpy(a_img.copy()).to(dev), torch.from_numpy(b_img.copy()).to(dev), N, Kdim // 16,
                                lbo, sbo, 0, 2 * lbo, 2 * lbo, _idesc_f16(128, N, 0, 0, 0, 0), 1)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    bound = 2.0 ** -21 * (np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T)
    assert np.all(np.abs(D.cpu().numpy() - ref) <= bound + 1e-30)


@pytest.mark.parametrize("fmt", ["bf16", "f16"])
def test_16bit_mnmajor_sw128_tile_product(fmt):
    """The layout the Gram kernel wants for its exactly representable leading term: 16-bit operands, MN-major
    (features contiguous inside a frame), 128-byte swizzle.  Both arrangements of the atoms are tried; the
    test states which one the hardware accepts (gram_tc.cu uses the K-contiguous one)."""
    from pmarlo_b200 import kernels

    rng = np.random.default_rng(3)
    N, Kdim = 256, 32
    At = np.round(rng.normal(size=(Kdim, 128)) * 8) / 8        # exactly representable in bf16 and fp16
    Bt = np.round(rng.normal(size=(Kdim, N)) * 8) / 8
    a16, b16 = _to16(At, fmt), _to16(Bt, fmt)
    ref = _from16(a16, fmt).T @ _from16(b16, fmt)
    f = 0 if fmt == "f16" else 1
    dev = torch.device("cuda")
    ok = {}
    for kc in (True, False):
        a_img, a_smn, a_sk = _img_mnmajor16_sw128(a16, kc)
        b_img, b_smn, b_sk = _img_mnmajor16_sw128(b16, kc)
        for swapped in (False, True):
            # descriptor semantics: lbo = stride between atoms along MN, sbo = along K (or the other way round);
            # with K-contiguous atoms both operands share s_k = 1024 and only the MN stride differs with the extent
            if not kc and a_sk != b_sk:
                continue   # one (lbo, sbo) pair serves both operands in the raw entry point
            lbo, sbo = (a_sk, a_smn) if swapped else (a_smn, a_sk)
            if kc and a_smn != b_smn:
                pass       # K-contiguous: s_mn = n_k * 1024 is the same for both operands
            D = kernels.tc_selftest_raw(torch.from_numpy(a_img.copy()).to(dev), torch.from_numpy(b_img.copy()).to(dev), N,
                                        Kdim // 16, lbo, sbo, 2, 2 * a_sk, 2 * b_sk, _idesc_f16(128, N, f, f, 1, 1), 1)
            err = float(np.max(np.abs(D.cpu().numpy() - ref)))
            ok[(kc, swapped)] = err
    print("16-bit MN-major SW128 hypotheses (k_contiguous, lbo/sbo swapped) -> max error:", ok)
    assert ok[(True, False)] == 0.0, ok
