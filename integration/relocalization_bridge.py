"""Relocalizer and keyframe-overlap call sites on the batched matcher.

The reference bypasses the pipeline interface here and builds ``cv2.BFMatcher(NORM_HAMMING,
crossCheck=True)`` objects directly:

  * ``MapRelocalizer.relocalize`` (/root/reference/persistent_map.py:226-319) matches the query
    frame against up to ``max_candidates`` BoW-ranked keyframes ONE AT A TIME and runs
    ``estimate_pose_from_matches`` on each;
  * ``KeyframeManager._match_ratio`` / ``_build_window_observations``
    (/root/reference/keyframe_manager.py:123-155) match consecutive keyframes pairwise.

``BatchedMapRelocalizer`` keeps the reference's constructor, validation, candidate ranking,
rejection order and best-candidate rule, but issues ONE Hamming launch for all candidates
(every keyframe block against the single device copy of the query block — the layout of
BASELINE config #5) and ONE set of RANSAC launches for all candidates that passed the match
gate.  ``sweep`` is config #5 itself: the query against EVERY keyframe of the map, the map's
descriptors resident on the device.  Float32 descriptors (the L2 branch of
``persistent_map._build_matcher``, :326-331) stay on OpenCV exactly as in the reference.
"""
from __future__ import annotations

import logging
from dataclasses import dataclass
from typing import Callable, Sequence

import cv2
import numpy as np

LOGGER = logging.getLogger(__name__)


@dataclass(frozen=True)
class RelocalizationResult:
    """Field-for-field persistent_map.RelocalizationResult (:57-64)."""
    frame_id: int
    score: float
    match_count: int
    inliers: int
    rotation: np.ndarray
    translation: np.ndarray


def _validate_bow(descriptors, vocab):
    """The argument checks of persistent_map.compute_bow_histogram (:85-90)."""
    if descriptors.ndim != 2:
        raise ValueError("Descriptors must be a 2D array")
    if vocab.ndim != 2:
        raise ValueError("Vocabulary must be a 2D array")
    if descriptors.shape[1] != vocab.shape[1]:
        raise ValueError("Descriptor dimensionality must match vocabulary")


def _is_orb(descriptors, vocab) -> bool:
    return descriptors.dtype == np.uint8 and descriptors.shape[1] == 32 and vocab.shape[1] == 32


def host_bow_histogram(descriptors: np.ndarray, vocab: np.ndarray) -> np.ndarray:
    """The reference's arithmetic on the host: what non-ORB descriptors (float32 / L2 maps, the
    branch persistent_map._build_matcher keeps on OpenCV, :326-331) use, and what the CPU
    control-flow tests inject.  Never a stand-in for the device path on ORB descriptors."""
    if descriptors is None or len(descriptors) == 0:
        return np.zeros(vocab.shape[0], dtype=np.float32)
    _validate_bow(descriptors, vocab)
    desc = descriptors.astype(np.float64)
    voc = vocab.astype(np.float32, copy=False).astype(np.float64)
    d2 = (voc * voc).sum(1)[None] - 2.0 * desc @ voc.T
    words = np.argmin(d2, axis=1)
    hist = np.bincount(words, minlength=vocab.shape[0]).astype(np.float32)
    if hist.sum() > 0:
        hist /= hist.sum()
    return hist


def host_bow_scores(descriptors: np.ndarray, vocab: np.ndarray, hists: np.ndarray) -> np.ndarray:
    """host_bow_histogram + sklearn's cosine_similarity, as persistent_map.py:234-235."""
    return _cosine_row(host_bow_histogram(descriptors, vocab), hists)


_BOW = {}


def _bow_index(vocab: np.ndarray):
    """One device copy of a vocabulary (keyed by its bytes)."""
    from b200slam.frontend import BowIndex

    v = np.ascontiguousarray(vocab, dtype=np.float32)
    key = (v.shape, hash(v.tobytes()))
    if key not in _BOW:
        if len(_BOW) > 8:
            _BOW.clear()
        _BOW[key] = BowIndex(v)
    return _BOW[key]


def compute_bow_histogram(descriptors: np.ndarray, vocab: np.ndarray) -> np.ndarray:
    """Drop-in for persistent_map.compute_bow_histogram (:82-96): nearest-centroid word counts,
    L1-normalised, float32.  ORB descriptors ((N, 32) uint8) go through K9 on the device
    (csrc/bow.cu); anything else keeps the reference's host arithmetic."""
    if descriptors is None or len(descriptors) == 0:
        return np.zeros(vocab.shape[0], dtype=np.float32)
    _validate_bow(descriptors, vocab)
    if not _is_orb(descriptors, vocab):
        return host_bow_histogram(descriptors, vocab)
    return _bow_index(vocab).histograms_host([descriptors])[0]


def bow_histograms_batch(descriptor_blocks: Sequence[np.ndarray], vocab: np.ndarray) -> np.ndarray:
    """The ``np.vstack([compute_bow_histogram(kf.descriptors, vocab) ...])`` of build_snapshot
    (persistent_map.py:110-112) for a whole map in ONE launch -> float32 [n, k]."""
    vocab = np.asarray(vocab)
    blocks = [np.zeros((0, 32), np.uint8) if d is None or len(d) == 0 else np.asarray(d) for d in descriptor_blocks]
    for d in blocks:
        if len(d):
            _validate_bow(d, vocab)
    if not all(_is_orb(d, vocab) for d in blocks):
        return np.vstack([host_bow_histogram(d, vocab) for d in blocks]) if blocks else np.zeros((0, vocab.shape[0]), np.float32)
    return _bow_index(vocab).histograms_host(blocks)


def _cosine_row(hist: np.ndarray, hists: np.ndarray) -> np.ndarray:
    """sklearn.metrics.pairwise.cosine_similarity([hist], hists)[0] (persistent_map.py:235), in
    the input precision (float32 histograms) so that near-ties rank as in the reference."""
    try:
        from sklearn.metrics.pairwise import cosine_similarity
        return cosine_similarity([hist], hists)[0]
    except ImportError:
        a = np.asarray(hist, dtype=np.float32)[None].copy()
        B = np.asarray(hists, dtype=np.float32).copy()
        na = np.sqrt(np.einsum("ij,ij->i", a, a))
        nb = np.sqrt(np.einsum("ij,ij->i", B, B))
        a /= np.where(na == 0, 1, na)[:, None]
        B /= np.where(nb == 0, 1, nb)[:, None]
        return (a @ B.T)[0]


def cross_check_batch(query: np.ndarray, blocks: Sequence[np.ndarray], query_is_train: bool = True):
    """``BFMatcher(NORM_HAMMING, crossCheck=True).match(block, query)`` for every block, one
    launch; -> list of (queryIdx, trainIdx, distance) int32 arrays in ascending queryIdx order.
    The query descriptors exist once on the device (q_src / t_src row indirection)."""
    import torch
    from b200slam.frontend import HammingMatcher, PairBatch

    blocks = [np.ascontiguousarray(b, dtype=np.uint8) for b in blocks]
    if not blocks:
        return []
    m = _matcher()
    n = len(blocks)
    if query_is_train:
        b = PairBatch.from_host(blocks, [query] * n)
        b.t_desc = torch.from_numpy(np.ascontiguousarray(query, dtype=np.uint8)).cuda()
        b.t_src = torch.zeros(n, dtype=torch.int32, device="cuda")
    else:
        b = PairBatch.from_host([query] * n, blocks)
        b.q_desc = torch.from_numpy(np.ascontiguousarray(query, dtype=np.uint8)).cuda()
        b.q_src = torch.zeros(n, dtype=torch.int32, device="cuda")
    keys = m.knn2(b, need_second=False)
    sel = m.select(b, keys, use_ratio=False, use_cross=True, sort_by_distance=False, max_matches=None)
    packed = torch.stack([sel.out_q, sel.out_t, sel.out_d]).cpu().numpy()
    cnt = sel.count.cpu().numpy()
    out = []
    for p in range(n):
        o, c = int(b.q_off_host[p]), int(cnt[p])
        out.append((packed[0, o:o + c].copy(), packed[1, o:o + c].copy(), packed[2, o:o + c].copy()))
    return out


_MATCHER = None


def _matcher():
    global _MATCHER
    if _MATCHER is None:
        from b200slam.frontend import HammingMatcher
        _MATCHER = HammingMatcher()
    return _MATCHER


def keyframe_matcher() -> Callable[[np.ndarray, np.ndarray], list]:
    """The callable ``KeyframeManager(matcher=...)`` expects (keyframe_manager.py:39): the
    cross-check matcher, DMatch list in ascending queryIdx order."""
    from .pose_bridge import CrossCheckMatcher
    return CrossCheckMatcher().match


class BatchedMapRelocalizer:
    """Drop-in for persistent_map.MapRelocalizer (:196-319)."""

    def __init__(self, snapshot, intrinsics: np.ndarray | None, *, min_matches: int = 60, min_inliers: int = 30,
                 max_candidates: int = 5, score_threshold: float = 0.75, ransac_threshold: float = 0.01,
                 verify_geometry: bool = True, batch_matcher=None, pose_solver=None, bow_scorer=None) -> None:
        if snapshot.bow_hists.size == 0:
            raise ValueError("Persistent map has no BoW histograms")
        if verify_geometry and intrinsics is None:
            raise ValueError("Intrinsics are required for geometric verification")
        self.snapshot = snapshot
        self.intrinsics = intrinsics
        self.min_matches = min_matches
        self.min_inliers = min_inliers
        self.max_candidates = max_candidates
        self.score_threshold = score_threshold
        self.ransac_threshold = ransac_threshold
        self.verify_geometry = verify_geometry
        self._frame_lookup = {kf.frame_id: kf for kf in snapshot.keyframes}
        # injection points (tests run the control flow on the CPU with reference matchers)
        self._batch_matcher = batch_matcher or cross_check_batch
        self._pose_solver = pose_solver
        self._bow_scorer = bow_scorer
        self._map_dev = None
        self._bow_dev = None
        self._map_hists_dev = None

    # ---- persistent_map.py:226-319 ----------------------------------------------------------
    def relocalize(self, keypoints, descriptors: np.ndarray):
        if descriptors is None or len(descriptors) == 0:
            raise ValueError("Descriptors are required for relocalization")
        scores = self._bow_scores(descriptors)
        ranked = sorted(range(len(scores)), key=lambda idx: (-float(scores[idx]), int(self.snapshot.bow_frame_ids[idx])))
        cands = []
        for idx in ranked[: self.max_candidates]:
            score = float(scores[idx])
            if score < self.score_threshold:
                continue
            frame_id = int(self.snapshot.bow_frame_ids[idx])
            kf = self._frame_lookup.get(frame_id)
            if kf is None:
                LOGGER.warning("BoW frame id %d missing from keyframes", frame_id)
                continue
            if not self.verify_geometry:                          # first eligible candidate wins (:250-258)
                return RelocalizationResult(frame_id, score, 0, 0, np.eye(3), np.zeros(3))
            if keypoints is None:
                raise ValueError("Keypoints required for geometric verification")
            cands.append((frame_id, score, kf))
        if not cands:
            LOGGER.info("Relocalization failed: no candidates passed thresholds")
            return None
        # one launch: every candidate keyframe block (query side) against the current frame (train side)
        hamming = [c for c in cands if c[2].descriptors.dtype == np.uint8 and np.asarray(descriptors).dtype == np.uint8]
        matched = {}
        if hamming:
            for (fid, _, _), arrs in zip(hamming, self._batch_matcher(descriptors, [c[2].descriptors for c in hamming])):
                matched[fid] = arrs
        for fid, _, kf in cands:
            if fid not in matched:                                # float32 / L2 branch stays on OpenCV (:326-331)
                ms = cv2.BFMatcher(cv2.NORM_L2, crossCheck=True).match(kf.descriptors, descriptors)
                matched[fid] = (np.array([m.queryIdx for m in ms], np.int32), np.array([m.trainIdx for m in ms], np.int32),
                                np.array([m.distance for m in ms], np.float32))
        survivors = []
        for fid, score, kf in cands:
            qi, ti, d = matched[fid]
            if len(qi) < self.min_matches:
                LOGGER.debug("Candidate %d rejected: only %d matches", fid, len(qi))
                continue
            order = np.argsort(d, kind="stable")                  # sorted(matches, key=distance) (:266)
            survivors.append((fid, score, kf, qi[order], ti[order]))
        kp_query = np.array([k.pt for k in keypoints], dtype=np.float32).reshape(-1, 2) if len(survivors) else None
        poses = self._solve_poses([(np.asarray(kf.keypoints, np.float32)[qi], kp_query[ti]) for _, _, kf, qi, ti in survivors])
        best = None
        for (fid, score, kf, qi, ti), pose in zip(survivors, poses):
            if pose is None:
                LOGGER.debug("Candidate %d rejected: pose estimation failed", fid)
                continue
            rotation, translation, inliers = pose
            if len(inliers) < self.min_inliers:
                LOGGER.debug("Candidate %d rejected: %d inliers < %d", fid, len(inliers), self.min_inliers)
                continue
            result = RelocalizationResult(fid, score, len(qi), int(len(inliers)), rotation, translation)
            if best is None or (result.inliers, result.score, -result.frame_id) > (best.inliers, best.score, -best.frame_id):
                best = result
        if best:
            LOGGER.info("Relocalized against frame %d (score=%.3f inliers=%d)", best.frame_id, best.score, best.inliers)
        else:
            LOGGER.info("Relocalization failed: no candidates passed thresholds")
        return best

    def _bow_scores(self, descriptors) -> np.ndarray:
        """persistent_map.py:234-235: the query's BoW histogram and its cosine similarity to every
        map histogram.  ORB descriptors: K9 on the device (vocabulary and map histograms resident,
        one histogram launch + one cosine launch); other descriptors: the reference's host arithmetic."""
        vocab, hists = self.snapshot.bow_vocab, self.snapshot.bow_hists
        if self._bow_scorer is not None:
            return np.asarray(self._bow_scorer(descriptors, vocab, hists))
        descriptors = np.asarray(descriptors)
        _validate_bow(descriptors, vocab)
        if not _is_orb(descriptors, vocab):
            return host_bow_scores(descriptors, vocab, hists)
        if self._bow_dev is None:
            # the vocabulary object is shared process-wide (keyed by its bytes); THIS map's histograms stay
            # with this relocalizer, so two relocalizers over different snapshots never see each other's map
            self._bow_dev = _bow_index(vocab)
            self._map_hists_dev = self._bow_dev.upload_map(hists)
        import torch
        idx = self._bow_dev
        d = torch.from_numpy(np.ascontiguousarray(descriptors)).to(idx.dev)
        off = torch.tensor([0, len(descriptors)], dtype=torch.int32, device=idx.dev)
        return idx.scores(idx.histograms(d, off, 1, len(descriptors))[0], self._map_hists_dev).cpu().numpy()

    def _solve_poses(self, pairs):
        """estimate_pose_from_matches (homography.py:423-438) for every surviving candidate:
        one batched RANSAC (hypotheses + scoring + winner), the refit per candidate on the host and
        one batched decomposition / cheirality vote on the device (K7).
        -> list of (R, t, inlier indices) or None where the reference raises."""
        if not pairs:
            return []
        if self._pose_solver is not None:
            return [self._pose_solver(s, d) for s, d in pairs]
        from .pose_bridge import estimate_poses_batch

        return estimate_poses_batch([s for s, _ in pairs], [d for _, d in pairs], self.intrinsics, th=self.ransac_threshold)

    # ---- BASELINE config #5: the whole map -------------------------------------------------
    def sweep(self, descriptors: np.ndarray, top: int = 5):
        """Cross-check match of the query frame against EVERY keyframe of the map in one launch
        (map descriptors uploaded once and kept on the device).  -> (match_counts per keyframe in
        snapshot order, indices of the `top` keyframes by match count, ties to the lower frame id)."""
        import torch
        from b200slam.frontend import PairBatch, SharedBlocks

        kfs = self.snapshot.keyframes
        q = np.ascontiguousarray(descriptors, dtype=np.uint8)
        if self._map_dev is None or self._map_dev[3] < len(q):
            # map descriptors once, followed by a slot for the query: ONE buffer, so that every block
            # (each keyframe, and the query once instead of once per keyframe) is expanded a single time
            sizes = np.array([len(kf.descriptors) for kf in kfs], np.int64)
            off = np.zeros(len(kfs) + 1, np.int32)
            np.cumsum(sizes, out=off[1:])
            room = max(4096, len(q))
            cat = np.concatenate([np.ascontiguousarray(kf.descriptors, dtype=np.uint8) for kf in kfs] + [np.zeros((room, 32), np.uint8)], axis=0)
            self._map_dev = (torch.from_numpy(cat).cuda(), off, torch.from_numpy(off).cuda(), room)
        kmap, off, off_d, _ = self._map_dev
        n = len(kfs)
        kmap[int(off[-1]):int(off[-1]) + len(q)].copy_(torch.from_numpy(q))
        q_off = (np.arange(n + 1, dtype=np.int64) * len(q)).astype(np.int32)
        shared = SharedBlocks.build(np.concatenate([off[:-1], off[-1:]]), np.concatenate([np.diff(off), [len(q)]]),
                                    np.full(n, n), np.arange(n), kmap.device)
        batch = PairBatch(q_desc=kmap, t_desc=kmap, q_off=torch.from_numpy(q_off).cuda(), t_off=off_d,
                          q_off_host=q_off, t_off_host=off, q_src=torch.full((n,), int(off[-1]), dtype=torch.int32, device="cuda"),
                          t_src=off_d[:n].contiguous(), shared=shared)
        m = _matcher()
        keys = m.knn2(batch, need_second=False)
        sel = m.select(batch, keys, use_ratio=False, use_cross=True, sort_by_distance=False, max_matches=None)
        counts = sel.count.cpu().numpy().astype(np.int64)
        ids = np.array([int(kf.frame_id) for kf in kfs])
        order = sorted(range(n), key=lambda i: (-int(counts[i]), int(ids[i])))
        return counts, order[:top]
