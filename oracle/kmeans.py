"""Oracle: k-means assignment, Lloyd iterations, pmarlo post-processing
(TEST INFRASTRUCTURE).

* ``assign`` -- the definition the reference pins in
  tests/perf/test_discretize_assignment_perf.py:73-151:
  labels == argmin_k ||x - c_k||_2 by direct difference in fp64, first
  minimum wins.  Bit-exact target for the CUDA path.
* ``lloyd`` -- deeptime 0.4.5 ``KMeans.fit`` loop (C++ ``cluster_loop``) as
  called at src/pmarlo/markov_state_model/clustering.py:349-355,605-609:
  assign -> centres = member means (an empty cluster keeps its old centre)
  -> cost with the NEW centres -> stop when |cost-prev|/cost <= tolerance or
  max_iter.  deeptime is absent here: PARITY UNPINNED against its binaries;
  cross-checked against sklearn ``KMeans(init=c0, n_init=1, algorithm="lloyd")``.
  kmeans++ seeding uses deeptime's C++ RNG and is not reproducible, so the
  parity harness always passes ``initial_centers`` (a supported kwarg,
  clustering.py:243-250).
* ``remap_labels_and_inertia`` -- clustering.py:364-392 (dense relabel by
  sorted unique label, centres := member means, inertia).
"""

from __future__ import annotations

import numpy as np

__all__ = ["sqdist_direct", "assign", "lloyd", "remap_labels_and_inertia", "cluster_microstates"]


def sqdist_direct(Y: np.ndarray, centers: np.ndarray) -> np.ndarray:
    """(n,K) matrix of sum_d (y_d - c_kd)^2, summed sequentially over d in fp64
    (the same order the CUDA fp64 re-check uses)."""
    Y = np.asarray(Y, dtype=np.float64)
    C = np.asarray(centers, dtype=np.float64)
    out = np.zeros((Y.shape[0], C.shape[0]))
    for d in range(Y.shape[1]):
        diff = Y[:, d][:, None] - C[:, d][None, :]
        out += diff * diff
    return out


def assign(Y: np.ndarray, centers: np.ndarray, chunk: int = 8192):
    """Returns (labels int64, min squared distance fp64)."""
    Y = np.asarray(Y, dtype=np.float64)
    n = Y.shape[0]
    labels = np.empty(n, dtype=np.int64)
    dmin = np.empty(n, dtype=np.float64)
    for s in range(0, n, chunk):
        D = sqdist_direct(Y[s:s + chunk], centers)
        lab = np.argmin(D, axis=1)
        labels[s:s + chunk] = lab
        dmin[s:s + chunk] = D[np.arange(D.shape[0]), lab]
    return labels, dmin


def _update(Y, labels, centers):
    K, D = centers.shape
    sums = np.zeros((K, D))
    np.add.at(sums, labels, Y)
    cnt = np.bincount(labels, minlength=K).astype(np.float64)
    new = centers.copy()
    nz = cnt > 0
    new[nz] = sums[nz] / cnt[nz, None]
    return new, cnt


def lloyd(Y: np.ndarray, initial_centers: np.ndarray, max_iter: int = 500,
          tolerance: float = 1e-5):
    """Returns (centers, n_iter, cost, converged)."""
    Y = np.asarray(Y, dtype=np.float64)
    centers = np.asarray(initial_centers, dtype=np.float64).copy()
    prev_cost = 0.0
    it = 0
    converged = False
    cost = 0.0
    while True:
        labels, _ = assign(Y, centers)
        centers, _ = _update(Y, labels, centers)
        # deeptime kmeans.h cluster_loop: costAssignFunction(data, NEW centres, the assignments of the cluster
        # step), i.e. the inertia of the updated centres under the labels that produced them -- not under a
        # re-assignment (restated from the published source; deeptime is absent from the image: unpinned)
        diff = Y - centers[labels]
        cost = float(np.sum(diff * diff))
        rel = abs(cost - prev_cost) / cost if cost != 0.0 else 0.0
        prev_cost = cost
        it += 1
        if rel <= tolerance:
            converged = True
        if converged or it >= max_iter:
            break
    return centers, it, cost, converged


def remap_labels_and_inertia(Y: np.ndarray, raw_labels: np.ndarray):
    """clustering.py:364-392, vectorised (same results as the Python loops)."""
    Y = np.asarray(Y, dtype=np.float64)
    unique, remapped = np.unique(raw_labels, return_inverse=True)
    n_unique = int(unique.size)
    if n_unique == 0:
        raise ValueError("Clustering produced zero unique microstates")
    centers = np.zeros((n_unique, Y.shape[1]))
    np.add.at(centers, remapped, Y)
    cnt = np.bincount(remapped, minlength=n_unique).astype(np.float64)
    centers /= cnt[:, None]
    diffs = Y - centers[remapped]
    inertia = float(np.sum(diffs * diffs))
    return remapped.astype(np.int64), centers, n_unique, inertia


def cluster_microstates(Y, n_states, initial_centers, max_iter=500, tolerance=1e-5):
    """clustering.py:395-665 for method="kmeans", integer n_states, explicit
    initial centres, n_init=1."""
    centers, n_iter, cost, conv = lloyd(Y, initial_centers, max_iter, tolerance)
    raw, _ = assign(Y, centers)
    labels, new_centers, n_unique, inertia = remap_labels_and_inertia(Y, raw)
    return labels, new_centers, n_unique, inertia, n_iter
