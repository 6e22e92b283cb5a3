"""pytest plugin: run the REFERENCE's own test files, unmodified, against the device path.

    PYTHONPATH=<repo>:<repo>/monocular-visual-slam_b200:<reference tree> \
        python -m pytest -p integration.pytest_dropin <reference tree>/tests/test_robust_pose_estimator.py

``pytest_configure`` calls ``integration.pose_bridge.install()`` before any test module is collected, so the names the
reference binds with ``from homography import ...`` / ``from keyframe_manager import ...`` already are the device
drop-ins; the summary line ``b2s kernel launches: N`` proves the CUDA library did the work (tests/test_gpu_reference_suite.py).
"""
from __future__ import annotations


def pytest_configure(config):
    from integration.pose_bridge import install

    config._b2s_patched = install()


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    try:
        from b200slam import _capi

        n = int(_capi.load_library().b2s_launch_count())
    except Exception as exc:                      # library not built: say so, the calling test fails on the missing count
        terminalreporter.write_line(f"b2s kernel launches: unavailable ({exc})")
        return
    terminalreporter.write_line(f"b2s kernel launches: {n}")
    terminalreporter.write_line("b2s patched: " + ", ".join(sorted(getattr(config, "_b2s_patched", []))))
