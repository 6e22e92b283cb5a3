"""Control flow of integration.relocalization_bridge.BatchedMapRelocalizer against the
reference's MapRelocalizer (persistent_map.py:196-319), on the CPU: the batched matcher is
replaced by cv2 (the reference's own matcher) and the pose solver by a deterministic stub on
BOTH sides, so every gate, the candidate order and the best-candidate rule are compared
exactly.  Needs /root/reference (build container only)."""
import sys
from pathlib import Path

import numpy as np
import pytest

REF = Path("/root/reference")
pytestmark = pytest.mark.skipif(not (REF / "persistent_map.py").exists(), reason="reference tree not present")


@pytest.fixture(scope="module")
def pm():
    sys.path.insert(0, str(REF))
    try:
        import persistent_map
        yield persistent_map
    finally:
        sys.path.remove(str(REF))


def _cv2_batch(query, blocks):
    import cv2
    out = []
    for b in blocks:
        ms = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(b, query)
        out.append((np.array([m.queryIdx for m in ms], np.int32), np.array([m.trainIdx for m in ms], np.int32),
                    np.array([m.distance for m in ms], np.int32)))
    return out


def _keyframes(pm, rng, n_kf, n_desc, dtype):
    kfs = []
    for f in range(n_kf):
        d = rng.integers(0, 256, (n_desc, 32)).astype(dtype)
        pts = rng.uniform(0, 500, (n_desc, 2)).astype(np.float32)
        pose = np.eye(4)
        pose[0, 3] = f
        kfs.append(pm.MapKeyframe(frame_id=10 * f + 3, pose=pose, keypoints=pts, descriptors=d))
    return kfs


def test_bow_histogram_and_ranking_match_reference(pm):
    from integration.relocalization_bridge import BatchedMapRelocalizer, host_bow_histogram, host_bow_scores
    rng = np.random.default_rng(11)
    vocab = rng.normal(size=(24, 32)).astype(np.float32) * 60 + 128
    for dtype in (np.uint8, np.float32):
        kfs = _keyframes(pm, rng, 9, 80, dtype)
        snap = pm.build_snapshot(kfs, vocab)
        for kf in kfs:
            np.testing.assert_allclose(host_bow_histogram(kf.descriptors, vocab), pm.compute_bow_histogram(kf.descriptors, vocab))
        for q in (kfs[4].descriptors, rng.integers(0, 256, (70, 32)).astype(dtype)):
            for thr in (0.0, 0.9, 0.999):
                a = pm.MapRelocalizer(snap, None, verify_geometry=False, score_threshold=thr).relocalize(None, q)
                b = BatchedMapRelocalizer(snap, None, verify_geometry=False, score_threshold=thr, bow_scorer=host_bow_scores).relocalize(None, q)
                assert (a is None) == (b is None)
                if a is not None:
                    assert (a.frame_id, a.match_count, a.inliers) == (b.frame_id, b.match_count, b.inliers)
                    assert abs(a.score - b.score) < 1e-9
    with pytest.raises(ValueError):
        BatchedMapRelocalizer(snap, None, verify_geometry=True)
    with pytest.raises(ValueError):
        BatchedMapRelocalizer(snap, None, verify_geometry=False, bow_scorer=host_bow_scores).relocalize(None, np.zeros((0, 32), np.uint8))


def test_geometry_gates_and_best_rule_match_reference(pm, monkeypatch):
    import cv2
    from integration.relocalization_bridge import BatchedMapRelocalizer, host_bow_scores
    rng = np.random.default_rng(5)
    vocab = rng.normal(size=(16, 32)).astype(np.float32) * 60 + 128
    kfs = _keyframes(pm, rng, 8, 120, np.uint8)
    query = kfs[2].descriptors.copy()
    query[::3] ^= rng.integers(0, 4, (len(query[::3]), 32)).astype(np.uint8)       # noisy re-observation of keyframe 2
    kfs[5] = pm.MapKeyframe(frame_id=kfs[5].frame_id, pose=kfs[5].pose, keypoints=kfs[5].keypoints, descriptors=query[::-1].copy())
    snap = pm.build_snapshot(kfs, vocab)
    kps = [cv2.KeyPoint(float(x), float(y), 1.0) for x, y in rng.uniform(0, 500, (len(query), 2))]

    def stub_solver(src, dst):                       # deterministic "pose": inliers = matches whose points are far enough apart
        inl = np.flatnonzero(np.linalg.norm(src - dst, axis=1) > 150.0)
        if len(inl) < 8:
            return None
        return np.eye(3), np.array([0.0, 0.0, 1.0]), inl

    def stub_estimate(kp1, kp2, matches, K, ransac_threshold=0.01, min_matches=15):
        if len(matches) < min_matches:
            raise RuntimeError("too few matches")
        src = np.float32([kp1[m.queryIdx].pt for m in matches])
        dst = np.float32([kp2[m.trainIdx].pt for m in matches])
        r = stub_solver(src, dst)
        if r is None:
            raise RuntimeError("RANSAC essential matrix failed")
        return r[0], r[1], r[2], len(matches)

    monkeypatch.setattr(pm, "estimate_pose_from_matches", stub_estimate)
    K = np.eye(3)
    for kw in (dict(min_matches=20, min_inliers=5, score_threshold=0.0, max_candidates=8),
               dict(min_matches=60, min_inliers=30, score_threshold=0.0, max_candidates=8),
               dict(min_matches=20, min_inliers=5, score_threshold=0.0, max_candidates=2),
               dict(min_matches=500, min_inliers=5, score_threshold=0.0, max_candidates=8)):
        a = pm.MapRelocalizer(snap, K, **kw).relocalize(kps, query)
        b = BatchedMapRelocalizer(snap, K, batch_matcher=_cv2_batch, pose_solver=stub_solver, bow_scorer=host_bow_scores, **kw).relocalize(kps, query)
        assert (a is None) == (b is None), kw
        if a is not None:
            assert (a.frame_id, a.match_count, a.inliers) == (b.frame_id, b.match_count, b.inliers), kw
            assert abs(a.score - b.score) < 1e-9
    with pytest.raises(ValueError):
        BatchedMapRelocalizer(snap, K, batch_matcher=_cv2_batch, pose_solver=stub_solver, bow_scorer=host_bow_scores, score_threshold=0.0).relocalize(None, query)
