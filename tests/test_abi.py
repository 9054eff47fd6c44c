"""The C-ABI library loads on a CPU-only box and exports every symbol that
include/pmb200.h declares (no compute calls without a GPU)."""

from __future__ import annotations

import pathlib
import re
import subprocess

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "pmb200.h"


def declared_symbols() -> list[str]:
    text = HEADER.read_text()
    return sorted(set(re.findall(r"PMB_API\s+[\w\s\*]+?\b(pmb_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def built_lib():
    from pmarlo_b200 import LIB_PATH

    if not LIB_PATH.exists():
        subprocess.run(["make", "-C", str(ROOT / "pmarlo_b200" / "csrc"), "-j8"], check=True)
    return LIB_PATH


def test_header_declares_the_whole_path():
    syms = declared_symbols()
    for needed in ("pmb_featurize", "pmb_col_moments", "pmb_gram", "pmb_tica_solve", "pmb_project",
                   "pmb_kmeans_assign", "pmb_kmeans_update", "pmb_count_lagged", "pmb_mle_rev",
                   "pmb_eig_rev_topk", "pmb_last_error"):
        assert needed in syms


def test_library_exports_every_declared_symbol(built_lib):
    out = subprocess.run(["nm", "-D", "--defined-only", str(built_lib)], check=True, capture_output=True,
                         text=True).stdout
    exported = set(re.findall(r"\sT\s+(pmb_\w+)", out))
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, f"declared but not exported: {missing}"
    extra = sorted(exported - set(declared_symbols()))
    assert not extra, f"exported but not declared in include/pmb200.h: {extra}"


def test_ctypes_binding_covers_the_header(built_lib):
    from pmarlo_b200 import _lib

    assert sorted(_lib.PROTOTYPES) == declared_symbols()
    handle = _lib.load()
    assert handle.pmb_version() >= 100
    assert handle.pmb_launch_count() == 0 or handle.pmb_launch_count() > 0
    assert handle.pmb_last_error() is not None


def test_sass_is_sm100a_only(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", str(built_lib)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under pmarlo_b200/ may import or call it."""
    for path in (ROOT / "pmarlo_b200").rglob("*.py"):
        text = path.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), path
    for path in (ROOT / "pmarlo_b200" / "csrc").glob("*.cu*"):
        text = path.read_text(errors="ignore")
        assert not re.search(r"#include\s+[\"<][^\">]*oracle", text), path
    assert "oracle" not in (ROOT / "pmarlo_b200" / "csrc" / "Makefile").read_text()


def test_no_cpu_fallback_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import numpy as np

    import pmarlo_b200 as pm

    with pytest.raises(pm.Pmb200Error):
        pm.tica_reduce(np.zeros((50, 3)), lag=2)
    with pytest.raises(pm.Pmb200Error):
        pm.build_msm_from_labels([np.zeros(10, dtype=int)], lag=1)
    with pytest.raises(pm.Pmb200Error):
        pm.cluster_microstates(np.zeros((20, 2)), method="kmeans", n_states=2)
