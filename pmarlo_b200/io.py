"""Trajectory ingest and result export around the hot path.

* ``iterload`` / ``DCDReader``: the chunked streaming of io/trajectory.py:136-177 and
  markov_state_model/_loading.py:167-228 (``mdtraj.iterload(dcd, top=pdb, chunk=1000)``) without mdtraj: a DCD
  file (CHARMM / NAMD / OpenMM little-endian layout, optional unit-cell records) is memory-mapped and handed out
  as (frames, atoms, 3) float32 chunks in NANOMETRES like mdtraj does.
* ``featurize_stream``: chunks go through a pinned staging buffer to the device on a side stream while K1
  featurizes the previous chunk -- the path ``compute_features`` takes for a file on disk.
* ``save_analysis_results``: the files of _export.py:24-130 / _estimation.py:257-282 --
  ``{prefix}_{transition_matrix,count_matrix,free_energies,stationary_distribution,dtrajs,fes}.npy``, a
  ``.npz`` (CSC) next to a matrix with more than 10 000 cells of which fewer than 5 % are non-zero, and
  ``analysis_results.{pkl,json}``.
"""

from __future__ import annotations

import json
import pathlib
import pickle
import struct
from typing import Iterator, Sequence

import numpy as np
import torch

from . import kernels
from .features import FeaturePlan, featurize_device
from .topology import Topology, Trajectory, load_pdb

__all__ = ["DCDReader", "iterload", "featurize_stream", "save_matrix_intelligent", "save_analysis_results", "write_dcd"]


class DCDReader:
    """Random access to the frames of a DCD file (coordinates in Angstrom on disk)."""

    def __init__(self, path):
        self.path = pathlib.Path(path)
        raw = np.memmap(self.path, dtype=np.uint8, mode="r")
        if raw.size < 100 or struct.unpack("<i", raw[0:4].tobytes())[0] != 84 or raw[4:8].tobytes() != b"CORD":
            raise ValueError(f"{path}: not a little-endian CORD DCD file")
        icntrl = np.frombuffer(raw[8:88].tobytes(), dtype="<i4")
        self.n_frames_header = int(icntrl[0])
        self.has_cell = bool(icntrl[10])
        if icntrl[11]:
            raise ValueError("4-dimensional DCD files are not supported")
        pos = 92
        title_len = struct.unpack("<i", raw[pos:pos + 4].tobytes())[0]
        pos += 4 + title_len + 4
        if struct.unpack("<i", raw[pos:pos + 4].tobytes())[0] != 4:
            raise ValueError("malformed DCD header (atom count record)")
        self.n_atoms = int(struct.unpack("<i", raw[pos + 4:pos + 8].tobytes())[0])
        pos += 12
        self._first = pos
        self._cell_bytes = 56 if self.has_cell else 0
        self._frame_bytes = self._cell_bytes + 3 * (4 * self.n_atoms + 8)
        self.n_frames = int((raw.size - pos) // self._frame_bytes)
        self._raw = raw

    def read(self, start: int, stop: int, atom_indices: Sequence[int] | None = None, stride: int = 1) -> np.ndarray:
        """Frames [start, stop) with the stride applied -> (n, atoms, 3) float32 in nm."""
        idx = np.arange(start, min(stop, self.n_frames), stride)
        na = self.n_atoms
        out = np.empty((idx.size, na, 3), dtype=np.float32)
        for k, f in enumerate(idx):
            base = self._first + int(f) * self._frame_bytes + self._cell_bytes
            for c in range(3):
                o = base + c * (4 * na + 8) + 4
                out[k, :, c] = np.frombuffer(self._raw[o:o + 4 * na], dtype="<f4")
        out *= 0.1                                          # Angstrom -> nm, as mdtraj reports
        return out if atom_indices is None else np.ascontiguousarray(out[:, np.asarray(atom_indices, dtype=int)])


def iterload(filename, *, top=None, stride: int = 1, atom_indices: Sequence[int] | None = None,
             chunk: int = 1000) -> Iterator[Trajectory]:
    """io/trajectory.py:136-177: yields ``Trajectory`` chunks of ``chunk`` frames (after striding)."""
    topo = load_pdb(top).topology if isinstance(top, (str, pathlib.Path)) else (top.topology if hasattr(top, "topology") else top)
    if topo is None:
        raise ValueError("a topology (PDB path, Trajectory or Topology) is required")
    rd = DCDReader(filename)
    if atom_indices is not None:
        ai = np.asarray(atom_indices, dtype=int)
        topo = Topology([topo.names[i] for i in ai], np.asarray(topo.resid)[ai], np.asarray(topo.chainid)[ai])
    step = int(max(1, chunk)) * int(max(1, stride))
    for s in range(0, rd.n_frames, step):
        xyz = rd.read(s, s + step, atom_indices, int(max(1, stride)))
        if xyz.shape[0]:
            yield Trajectory(xyz, topo)


def featurize_stream(filename, top, plan: FeaturePlan, chunk: int = 100_000, stride: int = 1) -> torch.Tensor:
    """Features (N, n_cols) float32 on the device of a DCD file: chunks are staged in two pinned buffers and copied
    on a side stream while the previous chunk is featurized (K1)."""
    dev = kernels.require_cuda()
    rd = DCDReader(filename)
    n = len(range(0, rd.n_frames, int(max(1, stride))))
    out = torch.empty((n, plan.n_cols), dtype=torch.float32, device=dev)
    step = int(max(1, chunk)) * int(max(1, stride))
    pinned = [torch.empty((int(max(1, chunk)), rd.n_atoms, 3), dtype=torch.float32, pin_memory=True) for _ in range(2)]
    staged = [torch.empty((int(max(1, chunk)), rd.n_atoms, 3), dtype=torch.float32, device=dev) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    done = [None, None]       # event: featurize of the chunk that last used buffer b has been enqueued and finished
    row = 0
    for k, s in enumerate(range(0, rd.n_frames, step)):
        b = k & 1
        xyz = rd.read(s, s + step, None, int(max(1, stride)))
        m = xyz.shape[0]
        if m == 0:
            continue
        if done[b] is not None:
            done[b].synchronize()                       # the pinned buffer is free again
        pinned[b][:m].copy_(torch.from_numpy(xyz))
        with torch.cuda.stream(copy_stream):
            staged[b][:m].copy_(pinned[b][:m], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        main.wait_event(ev)
        featurize_device(staged[b][:m], plan, out=out[row:row + m])
        fin = torch.cuda.Event()
        fin.record(main)
        copy_stream.wait_event(fin)
        done[b] = fin
        row += m
    return out[:row]


def write_dcd(path, xyz_nm: np.ndarray, with_cell: bool = False) -> None:
    """Minimal CORD DCD writer (tests and round trips): (n, atoms, 3) nm -> Angstrom on disk."""
    xyz = np.asarray(xyz_nm, dtype=np.float32) * 10.0
    n, na = xyz.shape[0], xyz.shape[1]
    icntrl = np.zeros(20, dtype="<i4")
    icntrl[0], icntrl[1], icntrl[2], icntrl[3] = n, 0, 1, n
    icntrl[10] = 1 if with_cell else 0
    icntrl[19] = 24
    with open(path, "wb") as f:
        f.write(struct.pack("<i", 84) + b"CORD" + icntrl.tobytes() + struct.pack("<i", 84))
        title = b"pmarlo_b200".ljust(80)
        f.write(struct.pack("<i", 84) + struct.pack("<i", 1) + title + struct.pack("<i", 84))
        f.write(struct.pack("<iii", 4, na, 4))
        for k in range(n):
            if with_cell:
                f.write(struct.pack("<i", 48) + np.array([10.0, 90.0, 10.0, 90.0, 90.0, 10.0], dtype="<f8").tobytes() + struct.pack("<i", 48))
            for c in range(3):
                f.write(struct.pack("<i", 4 * na) + xyz[k, :, c].astype("<f4").tobytes() + struct.pack("<i", 4 * na))


def save_matrix_intelligent(matrix, filename_base: str, output_dir, prefix: str = "msm_analysis") -> list[pathlib.Path]:
    """_estimation.py:257-282."""
    from scipy.sparse import csc_matrix, issparse, save_npz

    out_dir = pathlib.Path(output_dir)
    if matrix is None:
        return []
    written = [out_dir / f"{prefix}_{filename_base}.npy"]
    np.save(written[0], matrix.toarray() if issparse(matrix) else matrix)
    if matrix.size > 10000:
        if issparse(matrix):
            written.append(out_dir / f"{prefix}_{filename_base}.npz")
            save_npz(written[-1], matrix)
        elif np.count_nonzero(matrix) / matrix.size < 0.05:
            written.append(out_dir / f"{prefix}_{filename_base}.npz")
            save_npz(written[-1], csc_matrix(matrix))
    return written


def save_analysis_results(msm, output_dir, prefix: str = "msm_analysis", fes=None) -> list[pathlib.Path]:
    """_export.py:24-130 for an ``EnhancedMSM``-like object (attributes transition_matrix, count_matrix,
    free_energies, stationary_distribution, dtrajs, implied_timescales) and an optional ``FESResult``."""
    out_dir = pathlib.Path(output_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    written: list[pathlib.Path] = []
    written += save_matrix_intelligent(getattr(msm, "transition_matrix", None), "transition_matrix", out_dir, prefix)
    written += save_matrix_intelligent(getattr(msm, "count_matrix", None), "count_matrix", out_dir, prefix)
    for attr in ("free_energies", "stationary_distribution"):
        v = getattr(msm, attr, None)
        if v is not None:
            written.append(out_dir / f"{prefix}_{attr}.npy")
            np.save(written[-1], v)
    dtrajs = getattr(msm, "dtrajs", None)
    if dtrajs:
        written.append(out_dir / f"{prefix}_dtrajs.npy")
        arr = np.empty(len(dtrajs), dtype=object)
        for i, d in enumerate(dtrajs):
            arr[i] = np.asarray(d)
        np.save(written[-1], arr, allow_pickle=True)
    if fes is not None:
        written.append(out_dir / f"{prefix}_fes.npy")
        np.save(written[-1], fes.F)
    results: dict = {}
    if getattr(msm, "transition_matrix", None) is not None and getattr(msm, "count_matrix", None) is not None:
        results["msm"] = {"transition_matrix": msm.transition_matrix, "count_matrix": msm.count_matrix,
                          "free_energies": getattr(msm, "free_energies", None),
                          "stationary_distribution": getattr(msm, "stationary_distribution", None)}
    if fes is not None:
        results["fes"] = {"free_energy": fes.F, "xedges": fes.xedges, "yedges": fes.yedges,
                          "temperature": fes.metadata.get("temperature")}
    its = getattr(msm, "implied_timescales", None)
    if its is not None:
        results["its"] = {k: getattr(its, k) for k in ("lag_times", "eigenvalues", "eigenvalues_ci", "timescales",
                                                       "timescales_ci", "rates", "rates_ci") if hasattr(its, k)}
    with (out_dir / "analysis_results.pkl").open("wb") as f:
        pickle.dump(results, f)

    def meta(v):
        if isinstance(v, dict):
            return {k: meta(x) for k, x in v.items()}
        if isinstance(v, np.ndarray):
            return {"shape": list(v.shape), "dtype": str(v.dtype)}
        return v

    with (out_dir / "analysis_results.json").open("w") as f:
        json.dump(meta(results), f)
    written += [out_dir / "analysis_results.pkl", out_dir / "analysis_results.json"]
    return written
