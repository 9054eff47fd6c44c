"""ctypes binding of the plain-C oracle (``oracle/csrc/oracle_c.c`` -> ``liboracle_c.so``, built by
``make -C oracle`` / ``__graft_entry__.build()``).  TEST INFRASTRUCTURE: same access rule as the rest of
``oracle/``.  The numpy functions of ``oracle.msm`` / ``oracle.counts`` remain the definition; these are the
same algorithms compiled, used where the numpy loop would take minutes (ITS sweeps, K = 5000)."""

from __future__ import annotations

import ctypes as C
import pathlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_PATH = pathlib.Path(__file__).resolve().parent / "liboracle_c.so"
_lib = None


def available() -> bool:
    return _PATH.exists()


def _load():
    global _lib
    if _lib is None:
        if not _PATH.exists():
            raise RuntimeError(f"{_PATH} not built: run `make -C oracle`")
        h = C.CDLL(str(_PATH))
        h.oracle_mle_rev.restype = C.c_int64
        h.oracle_mle_rev.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        h.oracle_count_lagged.restype = None
        h.oracle_count_lagged.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]
        _lib = h
    return _lib


def mle_rev(Cm: np.ndarray, maxerr: float = 1e-8, maxiter: int = 1_000_000, threads: int = 1):
    """Same contract as ``oracle.msm.mle_rev``: (T, pi, n_iter).  ``threads`` > 1: rows dealt to a pthread team."""
    Cm = np.ascontiguousarray(Cm, dtype=np.float64)
    K = Cm.shape[0]
    T = np.empty((K, K))
    pi = np.empty((K,))
    it = _load().oracle_mle_rev(Cm.ctypes.data, K, float(maxerr), int(maxiter), T.ctypes.data, pi.ctypes.data,
                               int(threads))
    if it < 0:
        raise ValueError("count matrix has a state without outgoing counts")
    return T, pi, int(it)


def count_lagged(dtrajs, n_states: int, lag: int) -> np.ndarray:
    Cm = np.zeros((n_states, n_states), dtype=np.int64)
    h = _load()
    for d in dtrajs:
        d = np.ascontiguousarray(d, dtype=np.int64)
        h.oracle_count_lagged(d.ctypes.data, d.size, int(n_states), int(lag), Cm.ctypes.data)
    return Cm


def its_rev_mle(dtrajs, n_states: int, lags, n_timescales: int, maxerr: float = 1e-8,
                maxiter: int = 1_000_000, threads: int = 8):
    """``oracle.msm.its_rev_mle`` with the counting and the MLE in C, lags in parallel threads.
    Returns (timescales (n_lags, n_timescales), counts list, iterations)."""
    from . import msm

    lags = [int(v) for v in lags]
    out = np.full((len(lags), n_timescales), np.nan)
    iters = np.zeros(len(lags), dtype=np.int64)
    counts = [None] * len(lags)

    def one(a):
        Cm = count_lagged(dtrajs, n_states, lags[a])
        counts[a] = Cm
        lcs = msm.largest_connected_set(Cm)
        if lcs.size < 2:
            return
        T, pi, it = mle_rev(Cm[np.ix_(lcs, lcs)].astype(float), maxerr, maxiter)
        ev = msm.eigenvalues_rev(T, pi, min(n_timescales + 1, T.shape[0]))
        ts = msm.safe_timescales(lags[a], ev[1:])
        out[a, : ts.size] = ts
        iters[a] = it

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(one, range(len(lags))))
    return out, counts, iters
