"""CPU oracle for the pmarlo MSM-estimation hot path.  TEST INFRASTRUCTURE ONLY.

This package is a numpy/scipy fp64 restatement of what pmarlo computes on its
featurize -> TICA -> k-means -> lagged count -> reversible MSM -> ITS path
(SURVEY.md section 8a/8c).  pmarlo itself is pure Python and delegates the
arithmetic to mdtraj 1.10/1.11, deeptime 0.4.5 and scikit-learn 1.7, of which
only scikit-learn is installed in this image.  Therefore:

* integer stages (labels given centres, lagged counts, pair bookkeeping) are
  pinned bit-exactly against pmarlo's own importable functions
  (``tests/golden/make_golden.py`` runs them from /root/reference and commits
  the vectors);
* floating stages whose arithmetic lives in mdtraj / deeptime (dihedrals,
  symmetrised TICA covariances + spd_inv_split/eig_corr, the reversible MLE
  fixed point, eigenvalues_rev) are restated from the published algorithms and
  checked against analytic known answers, invariants and the assertions of the
  reference's own tests -- PARITY UNPINNED in the sense of literal golden
  vectors produced by deeptime/mdtraj (see DESIGN.md section "Oracle").

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  Nothing under
``pmarlo_b200/`` imports it; the product path fails loudly when the CUDA
library is missing instead of falling back to this code.
"""

from . import bayes, cext, ck, tpt, counts, featurize, kmeans, msm, pipeline, tica  # noqa: F401

__all__ = ["featurize", "tica", "kmeans", "counts", "msm", "ck", "pipeline", "cext", "bayes", "tpt"]
