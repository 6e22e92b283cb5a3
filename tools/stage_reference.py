#!/usr/bin/env python
"""Stage the UNMODIFIED reference tree into baseline/_ref/ (git-ignored; it travels to the GPU box with
gpurun like the built .so files).  Nothing is edited: Python sources, the .bak, configs/ and tests/ are
copied byte for byte, so that on the GPU box

  * the reference's own hot-path tests (tests/test_feature_pipeline.py, test_robust_pose_estimator.py,
    test_keyframe_manager.py, ...) run unmodified against the device path (tests/test_gpu_reference_suite.py),
  * bench.py --impl reference times the reference's own homography.ransac_essential and
    ORBFeaturePipeline.match (oracle/reference_path.py, kind "reference").

The live feature_pipeline.py shim stays as it is (it imports integration.feature_pipeline_bridge, i.e. this
repo); the original CPU implementation is read from feature_pipeline.py.bak under its own name.
Run by __graft_entry__.build() whenever /root/reference exists; a no-op elsewhere."""
from __future__ import annotations

import shutil
import sys
from pathlib import Path

SRC = Path("/root/reference")
DST = Path(__file__).resolve().parents[1] / "baseline" / "_ref"
KEEP_DIRS = ("configs", "tests")


def stage(verbose: bool = False) -> bool:
    if not SRC.exists():
        return False
    DST.mkdir(parents=True, exist_ok=True)
    n = 0
    for f in SRC.iterdir():
        if f.is_file() and (f.suffix in (".py", ".bak", ".txt") or f.name == "requirements.txt"):
            shutil.copy2(f, DST / f.name)
            n += 1
    for d in KEEP_DIRS:
        if (SRC / d).is_dir():
            shutil.copytree(SRC / d, DST / d, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    if verbose:
        print(f"staged {n} files + {KEEP_DIRS} into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage(verbose=True) else 1)
