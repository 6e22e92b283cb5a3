"""integration.relocalization_bridge on the device: the batched
cross-check matcher against the oracle, the whole-map sweep (BASELINE config #5 layout) and an
end-to-end relocalization on a synthetic map whose true keyframe is known."""
from dataclasses import dataclass

import numpy as np
import pytest

from oracle import hamming_oracle as ho

pytestmark = pytest.mark.gpu


@dataclass
class _KF:
    frame_id: int
    pose: np.ndarray
    keypoints: np.ndarray
    descriptors: np.ndarray


@dataclass
class _Snap:
    keyframes: tuple
    bow_vocab: np.ndarray
    bow_hists: np.ndarray
    bow_frame_ids: np.ndarray


def _noisy(rng, d, flip=0.06):
    bits = np.unpackbits(d, axis=1)
    bits ^= (rng.random(bits.shape) < flip).astype(np.uint8)
    return np.packbits(bits, axis=1)


def test_cross_check_batch_equals_oracle():
    from integration.relocalization_bridge import cross_check_batch
    rng = np.random.default_rng(8)
    query = rng.integers(0, 4, (300, 32), dtype=np.uint8)                      # tie-heavy alphabet
    blocks = [rng.integers(0, 4, (n, 32), dtype=np.uint8) for n in (1, 77, 300, 513)] + [query[::-1].copy()]
    got = cross_check_batch(query, blocks)
    for blk, (qi, ti, d) in zip(blocks, got):
        b, s, bw = ho.packed_keys(blk, query)
        eq, et, ed = ho.select_matches(b, s, bw, use_ratio=False, use_cross=True, ratio=1.0, max_matches=None, sort_by_distance=False)
        np.testing.assert_array_equal(qi, eq)
        np.testing.assert_array_equal(ti, et)
        np.testing.assert_array_equal(d, ed)


def _synthetic_map(rng, n_kf=40, n_desc=400):
    from integration.relocalization_bridge import compute_bow_histogram
    vocab = (rng.normal(size=(32, 32)) * 60 + 128).astype(np.float32)
    P = np.stack([rng.uniform(-6, 6, n_desc), rng.uniform(-2, 2, n_desc), rng.uniform(6, 30, n_desc)], axis=1)
    kfs = []
    for f in range(n_kf):
        d = rng.integers(0, 256, (n_desc, 32), dtype=np.uint8)
        pts = rng.uniform(-0.5, 0.5, (n_desc, 2)).astype(np.float32)
        kfs.append(_KF(frame_id=7 * f + 1, pose=np.eye(4), keypoints=pts, descriptors=d))
    true_kf = 23
    kfs[true_kf].keypoints = (P[:, :2] / P[:, 2:]).astype(np.float32)
    yaw = 0.03
    R = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
    P2 = P @ R.T + np.array([0.2, 0.0, -0.8])
    perm = rng.permutation(n_desc)
    q_desc = _noisy(rng, kfs[true_kf].descriptors)[perm]
    q_pts = (P2[:, :2] / P2[:, 2:] + rng.normal(0, 0.0007, (n_desc, 2)))[perm].astype(np.float32)
    hists = np.vstack([compute_bow_histogram(k.descriptors, vocab) for k in kfs])
    snap = _Snap(tuple(kfs), vocab, hists, np.array([k.frame_id for k in kfs], np.int64))
    return snap, true_kf, q_desc, q_pts


def test_sweep_counts_equal_oracle_and_rank_the_true_keyframe_first():
    from integration.relocalization_bridge import BatchedMapRelocalizer
    rng = np.random.default_rng(15)
    snap, true_kf, q_desc, _ = _synthetic_map(rng)
    rel = BatchedMapRelocalizer(snap, np.eye(3))
    counts, top = rel.sweep(q_desc, top=3)
    assert top[0] == true_kf
    for i in (0, true_kf, len(snap.keyframes) - 1):
        b, s, bw = ho.packed_keys(q_desc, snap.keyframes[i].descriptors)
        eq, _, _ = ho.select_matches(b, s, bw, use_ratio=False, use_cross=True, ratio=1.0, max_matches=None, sort_by_distance=False)
        assert counts[i] == len(eq)
    counts2, _ = rel.sweep(q_desc)                                             # second call reuses the device-resident map
    np.testing.assert_array_equal(counts, counts2)


def test_relocalize_end_to_end_finds_the_planted_keyframe():
    import cv2
    from integration.relocalization_bridge import BatchedMapRelocalizer
    rng = np.random.default_rng(16)
    snap, true_kf, q_desc, q_pts = _synthetic_map(rng)
    kps = [cv2.KeyPoint(float(x), float(y), 1.0) for x, y in q_pts]
    rel = BatchedMapRelocalizer(snap, np.eye(3), score_threshold=0.0, max_candidates=len(snap.keyframes), min_matches=60, min_inliers=30)
    res = rel.relocalize(kps, q_desc)
    assert res is not None and res.frame_id == snap.keyframes[true_kf].frame_id
    assert res.match_count > 300 and res.inliers > 200
    assert abs(np.linalg.det(res.rotation) - 1.0) < 1e-6 and abs(np.linalg.norm(res.translation) - 1.0) < 1e-6
    t_true = np.array([0.2, 0.0, -0.8]) / np.linalg.norm([0.2, 0.0, -0.8])
    assert abs(float(res.translation @ t_true)) > 0.98                         # direction up to the usual sign / noise
    # nothing passes a match gate no keyframe can meet
    assert BatchedMapRelocalizer(snap, np.eye(3), score_threshold=0.0, max_candidates=5, min_matches=10_000).relocalize(kps, q_desc) is None


def test_two_relocalizers_share_a_vocabulary_but_not_a_map():
    """Two snapshots built on the SAME vocabulary (old and new map): each relocalizer scores against its
    own histograms (the reference recomputes cosine_similarity against self.snapshot.bow_hists on every
    call, persistent_map.py:234-235) even when their first queries interleave."""
    from integration.relocalization_bridge import BatchedMapRelocalizer, host_bow_scores
    rng = np.random.default_rng(31)
    snap_a, _, q_desc, _ = _synthetic_map(rng, n_kf=30, n_desc=200)
    snap_b, _, _, _ = _synthetic_map(rng, n_kf=26, n_desc=200)
    snap_b = _Snap(snap_b.keyframes, snap_a.bow_vocab,
                   np.vstack([host_bow_scores.__globals__["host_bow_histogram"](k.descriptors, snap_a.bow_vocab) for k in snap_b.keyframes]),
                   snap_b.bow_frame_ids)
    ra, rb = BatchedMapRelocalizer(snap_a, np.eye(3)), BatchedMapRelocalizer(snap_b, np.eye(3))
    sa1 = ra._bow_scores(q_desc)
    sb1 = rb._bow_scores(q_desc)
    sa2 = ra._bow_scores(q_desc)          # after rb's first query: must still be map A
    assert sa1.shape == (30,) and sb1.shape == (26,)
    np.testing.assert_array_equal(sa1, sa2)
    np.testing.assert_allclose(sa1, host_bow_scores(q_desc, snap_a.bow_vocab, snap_a.bow_hists), atol=2e-6)
    np.testing.assert_allclose(sb1, host_bow_scores(q_desc, snap_b.bow_vocab, snap_b.bow_hists), atol=2e-6)
