"""Drop-in for the reference's feature pipeline module.

Exports exactly the six names ``feature_pipeline.py:1`` imports.  Behavioural spec:
``/root/reference/feature_pipeline.py.bak`` (129 lines).  Detection stays on
``cv2.ORB`` (CPU, out of scope, §8f #4); ``match`` — the hot path — runs on the B200
through libb2s (Hamming kNN-2 + selection kernels) and only materialises
``cv2.DMatch`` objects at the very end as a compatibility veneer.

No CUDA work happens at import or in constructors; the device is initialised lazily on
the first non-empty ``match`` in the calling process.  There is no CPU fallback: without
libb2s.so or a CUDA device ``match`` raises.
"""
from __future__ import annotations

import os
import threading
from dataclasses import dataclass
from typing import Callable, Sequence

import cv2
import numpy as np


@dataclass(frozen=True)
class FeaturePipelineConfig:
    """Same fields, defaults and validation as feature_pipeline.py.bak:12-31."""
    name: str = "orb"
    nfeatures: int = 2000
    ratio_test: float = 0.8
    cross_check: bool = True
    max_matches: int | None = 500
    deterministic_seed: int | None = 1337

    def __post_init__(self) -> None:
        if not self.name:
            raise ValueError("Feature pipeline name must be non-empty")
        if self.nfeatures <= 0:
            raise ValueError("nfeatures must be positive")
        if not 0 < self.ratio_test <= 1.0:
            raise ValueError("ratio_test must be in (0, 1]")
        if self.max_matches is not None and self.max_matches <= 0:
            raise ValueError("max_matches must be positive when provided")
        if self.deterministic_seed is not None and self.deterministic_seed < 0:
            raise ValueError("deterministic_seed must be non-negative")


@dataclass(frozen=True)
class MatchStats:
    """feature_pipeline.py.bak:34-38."""
    match_count: int
    mean_distance: float
    median_distance: float


MatcherCallable = Callable[[np.ndarray, np.ndarray], list]


class FeaturePipeline:
    """Interface of feature_pipeline.py.bak:44-61."""

    def detect_and_describe(self, image: np.ndarray):
        raise NotImplementedError

    def match(self, desc1, desc2) -> list:
        raise NotImplementedError

    def match_stats(self, matches: Sequence) -> MatchStats:
        if not matches:
            return MatchStats(match_count=0, mean_distance=0.0, median_distance=0.0)
        distances = np.array([m.distance for m in matches], dtype=np.float32)
        return MatchStats(match_count=len(matches), mean_distance=float(distances.mean()),
                          median_distance=float(np.median(distances)))


# ---- lazily created, per-process device matcher ---------------------------------------
_matcher = None
_matcher_pid = None
_matcher_lock = threading.Lock()


def _device_matcher():
    global _matcher, _matcher_pid
    with _matcher_lock:
        if _matcher is None or _matcher_pid != os.getpid():
            from b200slam.frontend import HammingMatcher  # imports torch; first use only

            _matcher = HammingMatcher()
            _matcher_pid = os.getpid()
        return _matcher


def to_dmatches(qi, ti, d) -> list:
    """(queryIdx, trainIdx, distance) arrays -> list[cv2.DMatch] (imgIdx 0, float distance)."""
    return [cv2.DMatch(int(a), int(b), 0, float(c)) for a, b, c in zip(qi.tolist(), ti.tolist(), d.tolist())]


def hamming_match_arrays(desc1, desc2, *, cross_check: bool, ratio_test: float = 0.8,
                         max_matches: int | None = None, sort_by_distance: bool = True,
                         combined: bool = False):
    """Array form of the matcher: -> (queryIdx, trainIdx, distance) int32 arrays.

    cross_check=True  -> BFMatcher(crossCheck=True).match rule (.bak:81-82), no ratio test
    cross_check=False -> knnMatch(k=2) + Lowe ratio (.bak:84-91)
    combined=True     -> ratio AND symmetry (homography.match_orb_descriptors, :9-26)
    """
    m = _device_matcher()
    return m.match_pairs([desc1], [desc2], use_ratio=combined or not cross_check,
                         use_cross=combined or cross_check, ratio=ratio_test,
                         sort_by_distance=sort_by_distance, max_matches=max_matches)[0]


class ORBFeaturePipeline(FeaturePipeline):
    """cv2.ORB detection + B200 Hamming matching (feature_pipeline.py.bak:64-95)."""

    def __init__(self, config: FeaturePipelineConfig) -> None:
        self.config = config
        self.detector = cv2.ORB_create(nfeatures=config.nfeatures)

    def detect_and_describe(self, image: np.ndarray):
        if self.config.deterministic_seed is not None:
            cv2.setRNGSeed(self.config.deterministic_seed)
        keypoints, descriptors = self.detector.detectAndCompute(image, None)
        return keypoints, descriptors

    def match_arrays(self, desc1, desc2):
        """The hot path without the DMatch veneer."""
        if desc1 is None or desc2 is None or len(desc1) == 0 or len(desc2) == 0:   # .bak:79-80
            z = np.zeros(0, np.int32)
            return z, z.copy(), z.copy()
        c = self.config
        return hamming_match_arrays(desc1, desc2, cross_check=c.cross_check, ratio_test=c.ratio_test,
                                    max_matches=c.max_matches, sort_by_distance=True)

    def match(self, desc1, desc2) -> list:
        return to_dmatches(*self.match_arrays(desc1, desc2))


def build_feature_pipeline(config: FeaturePipelineConfig) -> FeaturePipeline:
    """feature_pipeline.py.bak:98-101."""
    if config.name.lower() == "orb":
        return ORBFeaturePipeline(config)
    raise ValueError(f"Unsupported feature pipeline: {config.name}")


def matches_to_points(kp1, kp2, matches):
    """feature_pipeline.py.bak:104-111 (shape (0,) when empty, like the reference)."""
    pts1 = np.array([kp1[m.queryIdx].pt for m in matches], dtype=np.float32)
    pts2 = np.array([kp2[m.trainIdx].pt for m in matches], dtype=np.float32)
    return pts1, pts2


def adaptive_ransac_threshold(pts1, pts2, base_threshold: float, min_threshold: float,
                              max_threshold: float) -> float:
    """feature_pipeline.py.bak:114-129."""
    if pts1.size == 0 or pts2.size == 0:
        return float(np.clip(base_threshold, min_threshold, max_threshold))
    displacements = np.linalg.norm(pts2 - pts1, axis=1)
    if displacements.size == 0:
        return float(np.clip(base_threshold, min_threshold, max_threshold))
    scale = float(np.clip(float(np.median(displacements)) / 25.0, 0.5, 2.0))
    return float(np.clip(base_threshold * scale, min_threshold, max_threshold))
