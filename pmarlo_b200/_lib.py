"""ctypes binding of ``libpmb200.so`` (the C ABI declared in ``include/pmb200.h``).

The library holds every hand-written sm_100a kernel of the hot path.  There is
NO fallback: if the shared object is missing or a call fails, a
:class:`Pmb200Error` is raised.  PyTorch is used only for device memory,
streams and ``torch.distributed``.
"""

from __future__ import annotations

import ctypes as C
import os
import pathlib
import threading

__all__ = ["Pmb200Error", "lib", "load", "check", "ptr", "stream_handle", "LIB_PATH", "launch_count"]

LIB_PATH = pathlib.Path(__file__).resolve().parent / "libpmb200.so"

_i32, _i64, _f64 = C.c_int, C.c_int64, C.c_double
_p, _sz = C.c_void_p, C.c_size_t

# name -> (restype, argtypes); must list every symbol of include/pmb200.h
PROTOTYPES = {
    "pmb_last_error": (C.c_char_p, []),
    "pmb_version": (_i32, []),
    "pmb_launch_count": (_i64, []),
    "pmb_featurize": (_i32, [_p, _i64, _i32, _p, _i32, _i32, _p, _i64, _p]),
    "pmb_col_moments_ws_bytes": (_sz, [_i32]),
    "pmb_col_moments": (_i32, [_p, _i64, _i32, _i64, _p, _p, _p, _p, _sz, _p]),
    "pmb_pair_mask": (_i32, [_p, _i32, _i64, _i32, _p, _p]),
    "pmb_gram_ws_bytes": (_sz, [_i32]),
    "pmb_gram": (_i32, [_p, _i64, _i32, _i64, _p, _i32, _i32, _p, _p, _p, _p, _sz, _i32, _p]),
    "pmb_scaler_from_moments": (_i32, [_p, _i64, _i32, _i32, _i32, _p, _p, _p]),
    "pmb_tica_covariances": (_i32, [_p, _p, _p, _p, _p, _i64, _i64, _i32, _i32, _p, _p, _p, _p]),
    "pmb_tica_solve_ws_bytes": (_sz, [_i32]),
    "pmb_tica_solve": (_i32, [_p, _p, _i32, _f64, _p, _p, _p, _p, _sz, _p]),
    "pmb_tica_finalize": (_i32, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _p, _p, _p, _p]),
    "pmb_sym_eigvals_ws_bytes": (_sz, [_i32, _i32]),
    "pmb_sym_eigvals_batched": (_i32, [_p, _i32, _i32, _p, _p, _sz, _p]),
    "pmb_project": (_i32, [_p, _i64, _i32, _i64, _p, _p, _p, _i32, _p, _i64, _i32, _p]),
    "pmb_kmeans_assign_ws_bytes": (_sz, [_i64, _i32, _i32]),
    "pmb_kmeans_assign": (_i32, [_p, _i32, _i64, _i32, _i64, _p, _i32, _p, _p, _p, _p, _p, _p, _p, _sz, _i32, _p]),
    "pmb_debug_counters_kmeans": (_i32, [_p]),
    "pmb_debug_counters_gram": (_i32, [_p]),
    "pmb_debug_counters_tica": (_i32, [_p]),
    "pmb_debug_trace_tica": (_i32, [_p]),
    "pmb_kmeans_tc_scores": (_i32, [_p, _i64, _i32, _i64, _p, _i32, _p, _p, _p, _sz, _p]),
    "pmb_kmeans_update": (_i32, [_p, _p, _p, _i32, _i32, _p, _p]),
    "pmb_silhouette_samples": (_i32, [_p, _i64, _i32, _p, _i32, _p, _p, _p]),
    "pmb_count_lagged": (_i32, [_p, _i64, _p, _i32, _i32, _i32, _i32, _p, _p]),
    "pmb_count_lagged_weighted": (_i32, [_p, _p, _i64, _p, _i32, _i32, _i32, _i32, _p, _p]),
    "pmb_relabel_compact_ws_bytes": (_sz, [_i64]),
    "pmb_relabel_compact": (_i32, [_p, _i64, _p, _i32, _p, _i32, _p, _p, _p, _sz, _p]),
    "pmb_tc_selftest": (_i32, [_p, _p, _i32, _i32, _i32, _p, _p]),
    "pmb_tc_selftest_raw": (_i32, [_p, C.c_uint32, _p, C.c_uint32, _i32, _i32, C.c_uint32, C.c_uint32, C.c_uint32,
                                   C.c_uint32, C.c_uint32, C.c_uint32, _i32, _p, _p]),
    "pmb_mle_rev_ws_bytes": (_sz, [_i32, _i32]),
    "pmb_mle_rev": (_i32, [_p, _p, _i32, _i32, _f64, _f64, _i64, _p, _p, _p, _p, _sz, _p]),
    "pmb_hist2d": (_i32, [_p, _p, _p, _i64, _f64, _f64, _i32, _f64, _f64, _i32, _p, _p]),
    "pmb_bayes_rev_sample": (_i32, [_p, _p, _i32, _i32, _i32, _i32, C.c_uint64, _p, _p, _p]),
    "pmb_counts_active": (_i32, [_p, _i32, _f64, _p, _p, _p]),
    "pmb_trig_expand": (_i32, [_p, _i64, _i32, _p, _p, _p, _i32, _p]),
    "pmb_debug_counters": (_i32, [_p]),
    "pmb_eig_rev_topk_ws_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "pmb_eig_rev_topk": (_i32, [_p, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _sz, _p]),
}


class Pmb200Error(RuntimeError):
    """Raised when libpmb200 is missing or one of its entry points fails."""


_lock = threading.Lock()
_lib = None


def load(path: os.PathLike | str | None = None):
    """Load the shared library (idempotent) and bind every prototype."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        p = pathlib.Path(path) if path is not None else LIB_PATH
        if not p.exists():
            raise Pmb200Error(
                f"{p} not found: build it with `make -C pmarlo_b200/csrc` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
                "pmarlo_b200 has no CPU fallback."
            )
        try:
            handle = C.CDLL(str(p))
        except OSError as exc:  # e.g. libcudart missing
            raise Pmb200Error(f"cannot load {p}: {exc}") from exc
        for name, (res, args) in PROTOTYPES.items():
            try:
                fn = getattr(handle, name)
            except AttributeError as exc:
                raise Pmb200Error(f"{p} does not export {name}") from exc
            fn.restype = res
            fn.argtypes = args
        _lib = handle
        return _lib


def lib():
    return load()


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().pmb_last_error().decode("utf-8", "replace")
        raise Pmb200Error(f"{what or 'libpmb200'} failed (code {rc}): {msg}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None passes NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream_handle(device=None) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream


def launch_count() -> int:
    return int(lib().pmb_launch_count())
