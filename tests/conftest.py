"""pytest configuration: registers the ``gpu`` marker and shared fixtures."""

from __future__ import annotations

import pathlib
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    def load(name: str):
        return np.load(GOLDEN / f"{name}.npz", allow_pickle=False)
    return load


@pytest.fixture(scope="session")
def topologies(golden):
    z = golden("topologies")
    out = {}
    for key in ("ala2", "chig"):
        out[key] = dict(
            names=[str(s) for s in z[f"{key}_names"]],
            resn=[str(s) for s in z[f"{key}_resn"]],
            resid=z[f"{key}_resid"].astype(int),
            chain=z[f"{key}_chain"].astype(int),
            xyz=z[f"{key}_xyz"].astype(np.float32),
        )
    return out
