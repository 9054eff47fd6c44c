"""pmarlo_b200 -- B200-native (sm_100a) MSM-estimation hot path behind pmarlo's
Python call signatures: featurize -> TICA -> k-means -> lagged counts ->
reversible MSM -> implied timescales.  Hand-written CUDA kernels in
``libpmb200.so`` (C ABI: include/pmb200.h); no CPU fallback.
"""

from __future__ import annotations

from ._lib import LIB_PATH, Pmb200Error, launch_count, load as load_library  # noqa: F401
from .clustering import ClusteringResult, cluster_microstates  # noqa: F401
from .discretize import MSMDiscretizationResult, discretize_dataset  # noqa: F401
from .ck import (CKRunResult, CKTestResult, LagEvaluationResult, ck_rms_error, run_ck,  # noqa: F401
                 select_optimal_lag_ck_its)
from .enhanced_msm import EnhancedMSM  # noqa: F401
from .features import compute_features, featurize_trajectory, trig_expand_periodic  # noqa: F401
from .msm import (  # noqa: F401
    build_msm_from_labels,
    build_simple_msm,
    candidate_lag_ladder,
    deterministic_its_from_counts,
    check_transition_matrix,
    count_transitions,
    ensure_connected_counts,
    implied_timescales,
    safe_timescales,
)
from .reduction import TICA, maybe_apply_tica, reduce_features, tica_reduce, vamp_reduce  # noqa: F401
from .topology import Topology, Trajectory, load_pdb  # noqa: F401
from .analysis import FESResult, TPTAnalysis, TPTResult, free_energy_from_density, generate_2d_fes  # noqa: F401
from .io import DCDReader, featurize_stream, iterload, save_analysis_results  # noqa: F401

__version__ = "0.1.0"
