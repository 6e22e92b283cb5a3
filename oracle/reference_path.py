"""CPU baseline of the hot path.  TEST / BENCH INFRASTRUCTURE ONLY (bench.py's
``cpu_baseline`` leg and ``--impl reference``).

The reference is Python and cannot travel to the GPU box, so this is a *port* that makes
the same library calls in the same order as the reference's CPU path:
``cv2.BFMatcher(NORM_HAMMING)`` knnMatch(k=2) + Lowe ratio (feature_pipeline.py.bak:84-91),
``cv2.BFMatcher(crossCheck=True).match`` (.bak:82), sort + top-500 (.bak:92-94), then the
Python RANSAC loop of ``ransac_essential`` (homography.py:324-339: one 8-point SVD solve
and one NumPy Sampson pass per iteration, early exit above 0.8 n).
"""
from __future__ import annotations

import os
import time

import numpy as np

from . import ransac_oracle as ro


def match_pair_cv2(q, t, ratio=0.8, max_matches=500):
    """kNN-2 + ratio AND cross-check with OpenCV, sorted by distance, truncated."""
    import cv2

    knn = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
    ratio_ok = {p[0].queryIdx for p in knn if len(p) == 2 and p[0].distance < ratio * p[1].distance}
    cc = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(q, t)
    ms = [m for m in cc if m.queryIdx in ratio_ok]
    ms.sort(key=lambda m: m.distance)
    ms = ms[:max_matches] if max_matches else ms
    return (np.array([m.queryIdx for m in ms], np.int64), np.array([m.trainIdx for m in ms], np.int64),
            np.array([m.distance for m in ms], np.int64))


def cpu_pair(q, t, kq, kt, ratio=0.8, max_matches=500, th=0.01, max_iter=2000, seed=0, full_budget=False):
    """One frame pair on the CPU -> (n_matches, best_h, n_inliers)."""
    qi, ti, _ = match_pair_cv2(q, t, ratio, max_matches)
    if len(qi) < 8:
        return len(qi), -1, 0
    src, dst = kq[qi], kt[ti]
    rng = np.random.default_rng(seed)
    if not full_budget:                       # the reference's own control flow (early exit)
        try:
            _, inl, trace = ro.ransac_essential(src, dst, np.eye(3), th, max_iter, rng, return_trace=True)
            return len(qi), trace[3], len(inl)
        except RuntimeError:
            return len(qi), -1, 0
    samples = ro.draw_samples(rng, len(src), max_iter)      # same work as the GPU unit: all H hypotheses
    Es = ro.eight_point_E_batch(src, dst, np.eye(3), samples)
    _, counts = ro.score_hypotheses(Es, src, dst, th)
    h = ro.select_hypothesis(counts, len(src))
    return len(qi), h, int(counts[h]) if h >= 0 else 0


def _worker(args):
    import cv2

    cv2.setNumThreads(1)
    return cpu_pair(*args[0], **args[1])


def run_pairs(pairs, workers=None, **kw):
    """Process `pairs` (list of (q, t, kq, kt)) on `workers` processes (one OpenCV thread
    each).  -> (results, seconds, workers)."""
    import multiprocessing as mp

    workers = workers or os.cpu_count() or 1
    jobs = [(p, dict(kw, seed=i)) for i, p in enumerate(pairs)]
    t0 = time.perf_counter()
    if workers == 1:
        out = [_worker(j) for j in jobs]
    else:
        with mp.get_context("fork").Pool(workers) as pool:
            out = pool.map(_worker, jobs, chunksize=1)
    return out, time.perf_counter() - t0, workers
