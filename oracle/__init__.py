"""CPU oracle for the ORB-matching + RANSAC-E hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (NumPy + a small plain-C library) of the
reference's algorithm for the path named in BASELINE.json `north_star`.  It is
the checker the CUDA path is compared against, never the product:

* only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
  ``--impl reference`` legs of ``bench.py`` may import it;
* nothing under ``monocular-visual-slam_b200/`` or ``integration/`` imports it,
  and the product fails loudly when the CUDA library is missing.

Pinning (how we know the oracle is right): ``tests/golden/make_golden.py`` ran
the UNMODIFIED reference (``/root/reference/homography.py``,
``feature_pipeline.py.bak``) and the third-party library it calls
(``cv2.BFMatcher`` 4.13.0) in the build container and froze inputs + outputs
under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every
oracle function against those vectors.  The reference itself ships no golden
vectors for this path, so those generated fixtures are the pin.
"""

from .hamming_oracle import (  # noqa: F401
    NONE_KEY,
    IDX_BITS,
    IDX_MASK,
    hamming_matrix,
    knn2,
    cross_check_match,
    pipeline_match,
    match_orb_descriptors,
    packed_keys,
    ratio_lut,
    match_stats,
    adaptive_ransac_threshold,
)
from .ransac_oracle import (  # noqa: F401
    eight_point_E,
    eight_point_E_batch,
    sampson_sq_err,
    score_hypotheses,
    select_hypothesis,
    ransac_essential,
    draw_samples,
    decompose_essential,
)
