"""``integration`` — the package the reference's ``feature_pipeline.py`` shim imports
(/root/reference/feature_pipeline.py:1) and that does not exist in the reference tree.

Putting this repository's root on ``sys.path`` / ``PYTHONPATH`` makes
``from integration.feature_pipeline_bridge import ...`` resolve, which un-breaks every
reference module that imports ``feature_pipeline`` (slam_api, robust_pose_estimator,
feature_control_plane, slam_runner, ...).  Importing this package touches neither CUDA
nor torch: worker threads and forked children of the reference's control planes import
and construct pipelines freely.
"""
import sys
from pathlib import Path

_HOST = Path(__file__).resolve().parents[1] / "monocular-visual-slam_b200"
if str(_HOST) not in sys.path:
    sys.path.insert(0, str(_HOST))
