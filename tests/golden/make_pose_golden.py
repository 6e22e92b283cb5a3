#!/usr/bin/env python
"""Generate tests/golden/pose_golden.npz: PoseEstimate / PoseEstimationFailure of the UNMODIFIED reference
``RobustPoseEstimator.estimate_pose`` (/root/reference/robust_pose_estimator.py:89-251) on seeded two-view scenes.

Run in the build container only (needs /root/reference and cv2):

    python tests/golden/make_pose_golden.py

Nothing is copied.  ``feature_pipeline`` is bound to the reference's own ``feature_pipeline.py.bak`` (the live shim
would resolve to this repo) before ``robust_pose_estimator`` is imported, so every line that runs is the reference's.
The reference's RANSACs are unseeded (``np.random.default_rng()`` inside ``ransac_essential`` / ``ransac_homography``,
homography.py:191-192, 315-316; SURVEY finding 4): the generator replaces ``np.random.default_rng`` for the duration of
a call by a factory that hands out ``default_rng(seed0 + k)`` for the k-th generator requested, and records that seed —
the GPU test installs the same factory through ``integration.pose_bridge.RNG_FACTORY``.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import sys
from pathlib import Path

import cv2
import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def _load(name: str, path: Path):
    loader = importlib.machinery.SourceFileLoader(name, str(path))
    spec = importlib.util.spec_from_loader(name, loader)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    loader.exec_module(mod)
    return mod


_load("feature_pipeline", REF / "feature_pipeline.py.bak")
sys.path.insert(0, str(REF))
import robust_pose_estimator as rpe  # noqa: E402  (the unmodified reference module)


class SeededFactory:
    def __init__(self, seed0):
        self.seed0, self.k, self._orig = seed0, 0, np.random.default_rng

    def __call__(self, *args):
        if args and args[0] is not None:
            return self._orig(*args)
        g = self._orig(self.seed0 + self.k)
        self.k += 1
        return g

    def __enter__(self):
        np.random.default_rng = self
        return self

    def __exit__(self, *exc):
        np.random.default_rng = self._orig


def project(P, R, t):
    c = P @ R.T + t
    return (c[:, :2] / c[:, 2:3]).astype(np.float32)


def rot_y(a):
    return np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])


def scenes():
    """name -> (pts1, pts2, config kwargs).  K = I throughout (SURVEY finding 3)."""
    out = {}
    rng = np.random.default_rng(0)                                         # tests/test_robust_pose_estimator.py:32-51
    P = rng.uniform(-1, 1, (50, 3)) + np.array([0, 0, 3.0])
    out["ref_test_selects_model"] = (project(P, np.eye(3), np.zeros(3)), project(P, np.eye(3), np.array([0.1, 0, 0])),
                                     dict(min_matches=20, min_parallax=0.0))
    rng = np.random.default_rng(1)                                         # :54-74
    P = rng.uniform(-1, 1, (40, 3)) + np.array([0, 0, 3.0])
    out["ref_test_low_parallax"] = (project(P, np.eye(3), np.zeros(3)), project(P, np.eye(3), np.array([0.01, 0, 0])),
                                    dict(min_matches=20, min_parallax=10.0))
    rng = np.random.default_rng(2)                                         # :77-97
    P = rng.uniform(-1, 1, (50, 3)) + np.array([0, 0, 3.0])
    out["ref_test_cheirality_ratio"] = (project(P, np.eye(3), np.zeros(3)), project(P, np.eye(3), np.array([0.2, 0, 0])),
                                        dict(min_matches=20, min_cheirality_ratio=1.1, min_parallax=0.0))
    rng = np.random.default_rng(3)                                         # general motion, noise, 25 % outliers
    n = 300
    P = np.stack([rng.uniform(-4, 4, n), rng.uniform(-2, 2, n), rng.uniform(4, 20, n)], axis=1)
    p1 = project(P, np.eye(3), np.zeros(3)) + rng.normal(0, 5e-4, (n, 2)).astype(np.float32)
    p2 = project(P, rot_y(0.04), np.array([0.3, 0.02, -0.5])) + rng.normal(0, 5e-4, (n, 2)).astype(np.float32)
    bad = rng.permutation(n)[: n // 4]
    p2[bad] += rng.uniform(-0.3, 0.3, (len(bad), 2)).astype(np.float32)
    out["noisy_outliers_300"] = (p1, p2.astype(np.float32), dict(min_matches=20, min_parallax=0.0))
    rng = np.random.default_rng(4)                                         # planar scene: the homography branch is competitive
    n = 200
    xy = rng.uniform(-2, 2, (n, 2))
    P = np.stack([xy[:, 0], xy[:, 1], 6.0 + 0.3 * xy[:, 0]], axis=1)
    p1 = project(P, np.eye(3), np.zeros(3)) + rng.normal(0, 3e-4, (n, 2)).astype(np.float32)
    p2 = project(P, rot_y(-0.03), np.array([-0.4, 0.0, 0.1])) + rng.normal(0, 3e-4, (n, 2)).astype(np.float32)
    out["planar_200"] = (p1.astype(np.float32), p2.astype(np.float32), dict(min_matches=20, min_parallax=0.0))
    rng = np.random.default_rng(5)                                         # too few inliers for the default gates
    n = 60
    P = rng.uniform(-1, 1, (n, 3)) + np.array([0, 0, 3.0])
    p1 = project(P, np.eye(3), np.zeros(3))
    p2 = project(P, np.eye(3), np.array([0.15, 0, 0]))
    p2[20:] += rng.uniform(-0.4, 0.4, (n - 20, 2)).astype(np.float32)
    out["low_inlier_count"] = (p1, p2.astype(np.float32), dict(min_matches=20, min_parallax=0.0))
    rng = np.random.default_rng(6)                                         # default configuration end to end
    n = 400
    P = np.stack([rng.uniform(-5, 5, n), rng.uniform(-2, 2, n), rng.uniform(5, 30, n)], axis=1)
    p1 = project(P, np.eye(3), np.zeros(3)) * 500.0                        # pixel-like scale, identity intrinsics: parallax in "pixels"
    p2 = project(P, rot_y(0.02), np.array([0.05, 0.0, -1.0])) * 500.0
    out["default_config_pixel_scale"] = (p1.astype(np.float32), p2.astype(np.float32), {})
    return out


def main():
    data = {"names": []}
    for i, (name, (p1, p2, kw)) in enumerate(scenes().items()):
        kp1 = [cv2.KeyPoint(float(x), float(y), 1.0) for x, y in p1]
        kp2 = [cv2.KeyPoint(float(x), float(y), 1.0) for x, y in p2]
        matches = [cv2.DMatch(_queryIdx=j, _trainIdx=j, _distance=0.0) for j in range(len(kp1))]
        cfg = rpe.RobustPoseEstimatorConfig(**kw)
        seed0 = 1000 + 10 * i
        data["names"].append(name)
        data[f"{name}/pts1"], data[f"{name}/pts2"] = p1, p2
        data[f"{name}/seed0"] = np.int64(seed0)
        data[f"{name}/cfg_keys"] = np.array(sorted(kw), dtype="U32")
        data[f"{name}/cfg_vals"] = np.array([float(kw[k]) for k in sorted(kw)], np.float64)
        with SeededFactory(seed0) as fac:
            try:
                est = rpe.RobustPoseEstimator(cfg).estimate_pose(kp1, kp2, matches, np.eye(3))
                d = est.diagnostics
                data[f"{name}/outcome"] = np.array("estimate")
                data[f"{name}/method"] = np.array(d.method)
                data[f"{name}/R"], data[f"{name}/t"] = np.asarray(est.rotation, np.float64), np.asarray(est.translation, np.float64)
                data[f"{name}/inlier_indices"] = np.asarray(est.inlier_indices, np.int64)
                data[f"{name}/diag"] = np.array([d.match_count, d.inliers, d.inlier_ratio, d.median_parallax, d.cheirality_inliers,
                                                 d.cheirality_ratio, d.score], np.float64)
                print(f"{name}: {d.method} inliers {d.inliers}/{d.match_count} parallax {d.median_parallax:.4g} score {d.score:.4g} "
                      f"cheirality {d.cheirality_inliers} ({fac.k} generators)")
            except rpe.PoseEstimationFailure as f:
                data[f"{name}/outcome"] = np.array("failure")
                data[f"{name}/reason"] = np.array(f.reason)
                keys = sorted(f.metrics)
                data[f"{name}/metric_keys"] = np.array(keys, dtype="U32")
                data[f"{name}/metric_vals"] = np.array([f.metrics[k] for k in keys], np.float64)
                print(f"{name}: FAILURE {f.reason} {f.metrics} ({fac.k} generators)")
            except Exception as e:                                  # the reference raising something else is part of the contract too
                data[f"{name}/outcome"] = np.array("error")
                data[f"{name}/error_type"] = np.array(type(e).__name__)
                data[f"{name}/error_text"] = np.array(str(e))
                print(f"{name}: ERROR {type(e).__name__}: {e}")
    data["names"] = np.array(data["names"], dtype="U64")
    np.savez_compressed(OUT / "pose_golden.npz", **data)
    print("wrote", OUT / "pose_golden.npz")


if __name__ == "__main__":
    main()
