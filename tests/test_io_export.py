"""CPU suite for SURVEY.md 8f-3: DCD ingest (round trip through the minimal writer, chunking like
mdtraj.iterload) and the export file set of _export.py:24-130 / _estimation.py:257-282."""

from __future__ import annotations

import json
import pickle

import numpy as np
import pytest

from pmarlo_b200.io import DCDReader, iterload, save_analysis_results, save_matrix_intelligent, write_dcd
from pmarlo_b200.topology import Topology


@pytest.mark.parametrize("with_cell", [False, True])
def test_dcd_round_trip_and_chunking(tmp_path, with_cell):
    rng = np.random.default_rng(1)
    xyz = rng.normal(size=(257, 11, 3)).astype(np.float32)
    path = tmp_path / "t.dcd"
    write_dcd(path, xyz, with_cell=with_cell)
    rd = DCDReader(path)
    assert rd.n_frames == 257 and rd.n_atoms == 11 and rd.has_cell == with_cell
    np.testing.assert_allclose(rd.read(0, 257), xyz, rtol=2e-7, atol=1e-7)      # A <-> nm scaling in fp32
    np.testing.assert_allclose(rd.read(10, 20, atom_indices=[0, 5]), xyz[10:20][:, [0, 5]], rtol=2e-7, atol=1e-7)
    top = Topology(["CA"] * 11, np.arange(11), np.zeros(11, dtype=int))
    chunks = list(iterload(path, top=top, chunk=100))
    assert [c.n_frames for c in chunks] == [100, 100, 57]
    np.testing.assert_allclose(np.concatenate([c.xyz for c in chunks]), xyz, rtol=2e-7, atol=1e-7)
    strided = list(iterload(path, top=top, chunk=50, stride=3))
    np.testing.assert_allclose(np.concatenate([c.xyz for c in strided]), xyz[::3], rtol=2e-7, atol=1e-7)
    with pytest.raises(ValueError):
        (tmp_path / "bad.dcd").write_bytes(b"\x00" * 200)
        DCDReader(tmp_path / "bad.dcd")


def test_save_matrix_rule_and_result_files(tmp_path):
    class M:
        pass

    m = M()
    K = 120
    T = np.eye(K)
    T[0, 1], T[0, 0] = 0.25, 0.75
    m.transition_matrix, m.count_matrix = T, np.full((K, K), 3.0)
    m.free_energies, m.stationary_distribution = np.arange(K, dtype=float), np.full(K, 1.0 / K)
    m.dtrajs = [np.array([0, 1, 2]), np.array([3, 4])]
    m.implied_timescales = None
    files = save_analysis_results(m, tmp_path, prefix="msm_analysis")
    names = sorted(f.name for f in files)
    assert "msm_analysis_transition_matrix.npy" in names and "msm_analysis_transition_matrix.npz" in names   # 14400 cells, < 5 % non-zero
    assert "msm_analysis_count_matrix.npy" in names and "msm_analysis_count_matrix.npz" not in names        # dense
    for n in ("msm_analysis_free_energies.npy", "msm_analysis_stationary_distribution.npy", "msm_analysis_dtrajs.npy",
              "analysis_results.pkl", "analysis_results.json"):
        assert n in names
    from scipy.sparse import load_npz

    np.testing.assert_array_equal(load_npz(tmp_path / "msm_analysis_transition_matrix.npz").toarray(), T)
    np.testing.assert_array_equal(np.load(tmp_path / "msm_analysis_transition_matrix.npy"), T)
    d = np.load(tmp_path / "msm_analysis_dtrajs.npy", allow_pickle=True)
    assert len(d) == 2 and list(d[1]) == [3, 4]
    res = pickle.load(open(tmp_path / "analysis_results.pkl", "rb"))
    np.testing.assert_array_equal(res["msm"]["transition_matrix"], T)
    meta = json.load(open(tmp_path / "analysis_results.json"))
    assert meta["msm"]["count_matrix"]["shape"] == [K, K]
    small = save_matrix_intelligent(np.zeros((10, 10)), "x", tmp_path)
    assert [f.name for f in small] == ["msm_analysis_x.npy"]


def test_oracle_tpt_textbook_chain():
    """Symmetric nearest-neighbour chain: the committor between the ends is linear in the state index."""
    import oracle

    n = 7
    T = np.zeros((n, n))
    for i in range(n):
        for j in (i - 1, i + 1):
            if 0 <= j < n:
                T[i, j] = 0.25
        T[i, i] = 1.0 - T[i].sum()
    pi = np.full(n, 1.0 / n)
    r = oracle.tpt.reactive_flux(T, pi, [0], [n - 1])
    np.testing.assert_allclose(r["qf"], np.arange(n) / (n - 1), atol=1e-12)
    np.testing.assert_allclose(r["qb"], 1.0 - np.arange(n) / (n - 1), atol=1e-12)
    # flux is conserved along the chain
    f = np.array([r["net"][i, i + 1] for i in range(n - 1)])
    np.testing.assert_allclose(f, f[0], rtol=1e-12)
    assert abs(r["total_flux"] - f[0]) < 1e-15
