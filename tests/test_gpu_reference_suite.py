"""GPU: the reference's OWN hot-path tests, unmodified, against the device path.

``baseline/_ref/`` is the byte-for-byte staged reference tree (tools/stage_reference.py; git-ignored, it travels to the
GPU box with gpurun).  Its live ``feature_pipeline.py`` shim imports ``integration.feature_pipeline_bridge`` — this
repo — and the ``integration.pytest_dropin`` plugin calls ``install()`` before collection, so ``homography.*``,
``robust_pose_estimator.*``, ``persistent_map.*`` and ``keyframe_manager.KeyframeManager`` run on libb2s.  Skipped where
the staged tree is absent."""
import os
import re
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]
REF = ROOT / "baseline" / "_ref"

needs_ref = pytest.mark.skipif(not (REF / "tests" / "test_feature_pipeline.py").exists(), reason="baseline/_ref not staged")


def _run(test_file, tmp_path):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(ROOT), str(ROOT / "monocular-visual-slam_b200"), str(REF)]),
               PYTHONDONTWRITEBYTECODE="1")
    return subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "-p", "integration.pytest_dropin",
                           str(REF / "tests" / test_file), "--rootdir", str(tmp_path)], cwd=str(tmp_path), env=env,
                          capture_output=True, text=True, timeout=900)


@needs_ref
@pytest.mark.parametrize("test_file,min_passed,needs_launches", [
    ("test_feature_pipeline.py", 1, True),            # ORB on noise images -> pipeline.match on the device
    ("test_robust_pose_estimator.py", 3, True),       # essential + homography RANSAC, gates: the reference's estimator over device RANSAC
    ("test_keyframe_manager.py", 1, True),            # identical descriptors -> match ratio 1.0 through the cross-check matcher
    ("test_persistent_map.py", 1, False),             # float32 / L2 branch stays on OpenCV (persistent_map.py:326-331)
    ("test_slam_api.py", 2, False),                   # blank frames: match(None / empty) -> []
])
def test_reference_test_file_passes_unmodified_on_the_device(test_file, min_passed, needs_launches, tmp_path):
    if not (REF / "tests" / test_file).exists():
        pytest.skip(f"{test_file} not in the staged tree")
    r = _run(test_file, tmp_path)
    tail = r.stdout[-3000:] + r.stderr[-2000:]
    assert r.returncode == 0, tail
    m = re.search(r"(\d+) passed", r.stdout)
    assert m and int(m.group(1)) >= min_passed, tail
    assert " failed" not in r.stdout, tail
    n = re.search(r"b2s kernel launches: (\d+)", r.stdout)
    assert n, tail
    if needs_launches:
        assert int(n.group(1)) > 0, "the reference test passed without a single libb2s kernel: " + tail
        assert "homography.ransac_essential" in r.stdout


@needs_ref
def test_staged_tree_is_the_unmodified_reference():
    """The staged files are what the stage script copies: the dangling shim is still the one-line import of this
    repo's bridge and the original implementation is still next to it as .bak."""
    shim = (REF / "feature_pipeline.py").read_text().strip()
    assert shim.startswith("from integration.feature_pipeline_bridge import") and len(shim.splitlines()) == 1
    assert "class ORBFeaturePipeline" in (REF / "feature_pipeline.py.bak").read_text()
    assert "def ransac_essential" in (REF / "homography.py").read_text()
