"""CPU baseline of the hot path.  TEST / BENCH INFRASTRUCTURE ONLY (bench.py's
``cpu_baseline`` leg and ``--impl reference``).

Two arms, same library calls in the same order as the reference's CPU path:

* ``kind = "reference"`` — when ``baseline/_ref/`` holds the staged, UNMODIFIED reference tree
  (``tools/stage_reference.py``; git-ignored, travels to the GPU box with gpurun): matching through
  the reference's ``ORBFeaturePipeline`` matchers (``feature_pipeline.py.bak:64-95``: the very
  ``cv2.BFMatcher`` objects it builds) and RANSAC through the reference's own
  ``homography.ransac_essential`` (``homography.py:302-345``).
* ``kind = "port"`` — otherwise: the same cv2 calls + ``oracle.ransac_oracle.ransac_essential``, the
  restatement pinned to the reference by ``tests/golden``.

Per pair: ``cv2.BFMatcher(NORM_HAMMING)`` knnMatch(k=2) + Lowe ratio (.bak:84-91),
``cv2.BFMatcher(crossCheck=True).match`` (.bak:82), sort + top-500 (.bak:92-94), then the Python
RANSAC loop (one 8-point SVD solve and one NumPy Sampson pass per iteration, early exit above 0.8 n).

Threading: mode (ii) = one worker process per core, every worker pinned to ONE OpenCV
thread and ONE BLAS/OpenMP thread (an unpinned LAPACK team per worker oversubscribes the box about
10x — round 1's mistake); mode (i) = one process, OpenCV and BLAS free to use every core.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import time
from pathlib import Path

import numpy as np

from . import ransac_oracle as ro

REF_DIR = Path(__file__).resolve().parents[1] / "baseline" / "_ref"
_THREAD_VARS = ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS", "VECLIB_MAXIMUM_THREADS")

_ref_mods = None
_ref_pipes = {}


def reference_modules():
    """(homography, feature_pipeline_bak) of the staged reference, or None.  The .bak is imported
    under its own name (the live ``feature_pipeline.py`` shim resolves to this repo's bridge)."""
    global _ref_mods
    if _ref_mods is None:
        _ref_mods = False
        hp, fp = REF_DIR / "homography.py", REF_DIR / "feature_pipeline.py.bak"
        if hp.exists() and fp.exists():
            try:
                spec = importlib.util.spec_from_file_location("b2s_ref_homography", hp)
                hom = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(hom)
                from importlib.machinery import SourceFileLoader
                loader = SourceFileLoader("b2s_ref_feature_pipeline_bak", str(fp))
                spec = importlib.util.spec_from_loader(loader.name, loader)
                bak = importlib.util.module_from_spec(spec)
                sys.modules[loader.name] = bak            # dataclasses look their module up while the class is built
                loader.exec_module(bak)
                _ref_mods = (hom, bak)
            except Exception:                              # a broken copy must not take the baseline down
                _ref_mods = False
    return _ref_mods or None


def kind() -> str:
    return "reference" if reference_modules() else "port"


def cpu_model() -> str:
    try:
        for line in Path("/proc/cpuinfo").read_text().splitlines():
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def match_pair_cv2(q, t, ratio=0.8, max_matches=500):
    """kNN-2 + ratio AND cross-check with OpenCV, sorted by distance, truncated.  ratio=None: cross-check only
    (the relocalizer's matcher, persistent_map.py:266-270)."""
    import cv2

    ref = reference_modules()
    if ratio is None:
        if ref:
            if "cc" not in _ref_pipes:
                _ref_pipes["cc"] = ref[1].ORBFeaturePipeline(ref[1].FeaturePipelineConfig(cross_check=True, max_matches=None))
            ms = _ref_pipes["cc"].match(q, t)
        else:
            ms = sorted(cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(q, t), key=lambda m: m.distance)
        ms = ms[:max_matches] if max_matches else ms
        return (np.array([m.queryIdx for m in ms], np.int64), np.array([m.trainIdx for m in ms], np.int64),
                np.array([m.distance for m in ms], np.int64))
    if ref:   # the reference's own ORBFeaturePipeline.match (.bak:78-95), in both of its modes
        key = float(ratio)
        if key not in _ref_pipes:
            bak = ref[1]
            _ref_pipes[key] = (bak.ORBFeaturePipeline(bak.FeaturePipelineConfig(cross_check=False, ratio_test=ratio, max_matches=None)),
                               bak.ORBFeaturePipeline(bak.FeaturePipelineConfig(cross_check=True, max_matches=None)))
        knn_pipe, cc_pipe = _ref_pipes[key]
        ratio_ok = {m.queryIdx for m in knn_pipe.match(q, t)}
        ms = [m for m in cc_pipe.match(q, t) if m.queryIdx in ratio_ok]     # already stably sorted by distance
    else:
        knn = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
        ratio_ok = {p[0].queryIdx for p in knn if len(p) == 2 and p[0].distance < ratio * p[1].distance}
        cc = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(q, t)
        ms = [m for m in cc if m.queryIdx in ratio_ok]
        ms.sort(key=lambda m: m.distance)
    ms = ms[:max_matches] if max_matches else ms
    return (np.array([m.queryIdx for m in ms], np.int64), np.array([m.trainIdx for m in ms], np.int64),
            np.array([m.distance for m in ms], np.int64))


def cpu_pair(q, t, kq, kt, ratio=0.8, max_matches=500, th=0.01, max_iter=2000, seed=0, full_budget=False, ransac=True):
    """One frame pair on the CPU -> (n_matches, best_h, n_inliers)."""
    qi, ti, _ = match_pair_cv2(q, t, ratio, max_matches)
    if not ransac:
        return len(qi), -1, 0
    if len(qi) < 8:
        return len(qi), -1, 0
    src, dst = kq[qi], kt[ti]
    rng = np.random.default_rng(seed)
    if not full_budget:                       # the reference's own control flow (early exit)
        ref = reference_modules()
        try:
            if ref:
                _, inl = ref[0].ransac_essential(src, dst, np.eye(3), th, max_iter, rng)
                return len(qi), -2, len(inl)          # the reference does not report the winning iteration
            _, inl, trace = ro.ransac_essential(src, dst, np.eye(3), th, max_iter, rng, return_trace=True)
            return len(qi), trace[3], len(inl)
        except RuntimeError:
            return len(qi), -1, 0
    samples = ro.draw_samples(rng, len(src), max_iter)      # same work as the GPU unit: all H hypotheses
    Es = ro.eight_point_E_batch(src, dst, np.eye(3), samples)
    _, counts = ro.score_hypotheses(Es, src, dst, th)
    h = ro.select_hypothesis(counts, len(src))
    return len(qi), h, int(counts[h]) if h >= 0 else 0


def _init_worker():
    """One OpenCV thread and one BLAS/OpenMP thread per worker process."""
    for v in _THREAD_VARS:
        os.environ[v] = "1"
    import cv2

    cv2.setNumThreads(1)
    try:
        from threadpoolctl import threadpool_limits

        global _blas_limit
        _blas_limit = threadpool_limits(limits=1)      # kept alive for the life of the worker
    except ImportError:
        pass
    reference_modules()


def _worker(args):
    return cpu_pair(*args[0], **args[1])


def _ready(delay):
    time.sleep(delay)          # holds the worker long enough that every OTHER worker has to take one of these too
    return os.getpid()


class PairPool:
    """Mode (ii): a pool of `workers` processes, one OpenCV and one BLAS thread each, created ONCE and reused for every
    timed step.  `__enter__` returns only when every worker has finished its initializer (imports of cv2 / NumPy / the
    staged reference): a pool whose stragglers are still importing under the clock reads up to 2x slow."""

    def __init__(self, workers=None):
        self.workers = workers or os.cpu_count() or 1
        self.pool = None

    def __enter__(self):
        import multiprocessing as mp

        if self.workers > 1:
            # spawn, not fork: the parent may already own CUDA / NCCL state or OpenCV / BLAS thread pools (mode (i)
            # starts them), and a forked child of a threaded parent deadlocks on their locks
            self.pool = mp.get_context("spawn").Pool(self.workers, initializer=_init_worker)
            for _ in range(3):
                pids = set(self.pool.map(_ready, [0.25] * self.workers, chunksize=1))
                if len(pids) == self.workers:
                    break
        else:
            self._saved = {v: os.environ.get(v) for v in _THREAD_VARS}
            _init_worker()
        return self

    def __exit__(self, *exc):
        if self.pool is not None:
            self.pool.terminate()
            self.pool.join()
        else:
            for v, old in self._saved.items():
                if old is None:
                    os.environ.pop(v, None)
                else:
                    os.environ[v] = old

    def run(self, pairs, **kw):
        """-> (results, seconds) for one pass over `pairs`."""
        jobs = [(p, dict(kw, seed=i)) for i, p in enumerate(pairs)]
        t0 = time.perf_counter()
        out = self.pool.map(_worker, jobs, chunksize=1) if self.pool is not None else [_worker(j) for j in jobs]
        return out, time.perf_counter() - t0


def run_pairs(pairs, workers=None, **kw):
    """One pass over `pairs` (list of (q, t, kq, kt)) on a fresh PairPool.  -> (results, seconds, workers)."""
    with PairPool(workers) as pool:
        out, sec = pool.run(pairs, **kw)
        return out, sec, pool.workers


def run_pairs_single_process(pairs, **kw):
    """Mode (i) of SURVEY §8d: ONE process, ``cv2.setNumThreads(all cores)`` and the BLAS pool free —
    how the reference runs when ``slam_api`` calls it frame by frame.  -> (results, seconds, cv2 threads)."""
    import cv2

    cv2.setNumThreads(os.cpu_count() or 1)
    reference_modules()
    t0 = time.perf_counter()
    out = [cpu_pair(*p, **dict(kw, seed=i)) for i, p in enumerate(pairs)]
    return out, time.perf_counter() - t0, cv2.getNumThreads()
