"""GPU drop-in tests: the reference-shaped interfaces (integration.*) behave like the
reference's own — scenarios restated from /root/reference/tests/test_feature_pipeline.py,
test_robust_pose_estimator.py, test_loop_closure_verification.py, test_keyframe_manager.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

cv2 = pytest.importorskip("cv2")


def _project(P, R, t):
    c = (R @ P.T).T + t
    return (c[:, :2] / c[:, 2:3]).astype(np.float32)


def _kps(p):
    return [cv2.KeyPoint(float(x), float(y), 1.0) for x, y in p]


def _matches(n):
    return [cv2.DMatch(_queryIdx=i, _trainIdx=i, _distance=0.0) for i in range(n)]


def test_pipeline_on_orb_images_equals_cv2(golden_dir):
    from integration.feature_pipeline_bridge import (FeaturePipelineConfig, adaptive_ransac_threshold,
                                                     build_feature_pipeline, matches_to_points)
    rng = np.random.default_rng(0)
    a = rng.integers(0, 255, size=(240, 320), dtype=np.uint8)
    b = np.roll(a, 3, axis=1)
    for cross in (True, False):
        cfg = FeaturePipelineConfig(name="orb", nfeatures=500, cross_check=cross)
        pipe = build_feature_pipeline(cfg)
        ka, da = pipe.detect_and_describe(a)
        kb, db = pipe.detect_and_describe(b)
        got = pipe.match(da, db)
        assert isinstance(got, list)
        bf = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=cross)
        if cross:
            want = list(bf.match(da, db))
        else:
            want = [p[0] for p in bf.knnMatch(da, db, k=2) if len(p) == 2 and p[0].distance < cfg.ratio_test * p[1].distance]
        want.sort(key=lambda m: m.distance)
        want = want[:cfg.max_matches]
        assert [(m.queryIdx, m.trainIdx, m.distance, m.imgIdx) for m in got] == \
               [(m.queryIdx, m.trainIdx, m.distance, m.imgIdx) for m in want]
        pa, pb = matches_to_points(ka, kb, got)
        th = adaptive_ransac_threshold(pa, pb, 0.01, 0.005, 0.03)
        assert 0.005 <= th <= 0.03
        st = pipe.match_stats(got)
        assert st.match_count == len(got)
    assert pipe.match(None, db) == [] and pipe.match(da, np.zeros((0, 32), np.uint8)) == []


def test_identical_descriptors_match_ratio_one():
    """tests/test_keyframe_manager.py:30-37 — identical 50x32 descriptors all cross-match."""
    from integration.pose_bridge import CrossCheckMatcher
    d = np.random.default_rng(0).integers(0, 256, (50, 32), dtype=np.uint8)
    ms = CrossCheckMatcher().match(d, d)
    assert [(m.queryIdx, m.trainIdx, m.distance) for m in ms] == [(i, i, 0.0) for i in range(50)]


def test_robust_pose_estimator_scenarios():
    from integration.pose_bridge import PoseEstimationFailure, RobustPoseEstimator, RobustPoseEstimatorConfig
    K = np.eye(3)

    def scene(seed, n, t):
        P = np.random.default_rng(seed).uniform(-1, 1, (n, 3)) + np.array([0, 0, 3.0])
        return _kps(_project(P, np.eye(3), np.zeros(3))), _kps(_project(P, np.eye(3), np.array(t)))

    k1, k2 = scene(0, 50, [0.1, 0, 0])
    est = RobustPoseEstimator(RobustPoseEstimatorConfig(min_matches=20, min_parallax=0.0)).estimate_pose(k1, k2, _matches(50), K)
    assert est.diagnostics.inliers > 0 and est.diagnostics.method in {"essential", "homography"}
    assert abs(np.linalg.norm(est.translation) - 1.0) < 1e-9

    k1, k2 = scene(1, 40, [0.01, 0, 0])
    with pytest.raises(PoseEstimationFailure) as e:
        RobustPoseEstimator(RobustPoseEstimatorConfig(min_matches=20, min_parallax=10.0)).estimate_pose(k1, k2, _matches(40), K)
    assert e.value.reason == "low_parallax"

    k1, k2 = scene(2, 50, [0.2, 0, 0])
    with pytest.raises(PoseEstimationFailure) as e:
        RobustPoseEstimator(RobustPoseEstimatorConfig(min_matches=20, min_cheirality_ratio=1.1, min_parallax=0.0)).estimate_pose(k1, k2, _matches(50), K)
    assert e.value.reason == "cheirality_ratio"

    with pytest.raises(ValueError):
        RobustPoseEstimator(RobustPoseEstimatorConfig()).estimate_pose(k1, k2, _matches(5), K)


def test_pose_from_orb_with_inliers_identity_intrinsics():
    """tests/test_loop_closure_verification.py restated in the K = I regime (the reference's
    own fx=500 variant fails on the reference because of its intrinsics quirk, SURVEY finding 3):
    identical descriptor sets -> all 40 matched, >= 80 % inliers, pose recovered."""
    from integration.pose_bridge import estimate_pose_from_orb_with_inliers
    rng = np.random.default_rng(42)
    P = rng.uniform(-1, 1, (40, 3))
    P[:, 2] += 4.0
    yaw = 0.05
    R = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
    t = np.array([0.2, 0.0, 0.0])
    desc = rng.integers(0, 256, (40, 32), dtype=np.uint8)
    Re, te, inl, mc = estimate_pose_from_orb_with_inliers(_kps(_project(P, np.eye(3), np.zeros(3))), desc,
                                                          _kps(_project(P, R, t)), desc, np.eye(3),
                                                          ransac_threshold=0.01, min_matches=20)
    assert mc == 40 and len(inl) >= 32
    assert Re.shape == (3, 3) and te.shape == (3,)
    np.testing.assert_allclose(Re, R, atol=1e-3)
    np.testing.assert_allclose(te / np.linalg.norm(te), t / np.linalg.norm(t), atol=1e-2)


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()


def test_sequence_tracker_equals_independent_pairs():
    """Shared-frame sequence path (frames uploaded once, q_src/t_src indirection, compact
    outputs, 3-stream overlap) gives the same matches as independent pairs through the oracle."""
    import torch
    from b200slam import _capi
    from b200slam.frontend import FrontendConfig, SequenceTracker
    from b200slam.synthetic import tracking_sequence
    from oracle import hamming_oracle as ho
    F, N = 6, 700
    desc, kp = tracking_sequence(F, N, seed=7)
    counts = np.array([700, 650, 700, 512, 700, 699], np.int32)
    cfg = FrontendConfig(hypotheses=128, max_matches=300)
    for variant in (_capi.VARIANT_POPC, _capi.VARIANT_I8MMA, _capi.VARIANT_I8MMA1):
        tr = SequenceTracker(F, N, cfg, variant=variant, chunks=3)
        out = tr.run(torch.from_numpy(desc.reshape(-1, 32)).pin_memory(), torch.from_numpy(kp.reshape(-1, 2)).pin_memory(), counts)
        torch.cuda.synchronize()
        for p in range(F - 1):
            q, t = desc[p][:counts[p]], desc[p + 1][:counts[p + 1]]
            qi, ti, d = ho.select_matches(*ho.packed_keys(q, t), use_ratio=True, use_cross=True, ratio=0.8, max_matches=300)
            c = int(out["count"][p])
            assert c == len(qi)
            np.testing.assert_array_equal(out["out_q"][p * 300:p * 300 + c].numpy(), qi)
            np.testing.assert_array_equal(out["out_t"][p * 300:p * 300 + c].numpy(), ti)
            np.testing.assert_array_equal(out["out_d"][p * 300:p * 300 + c].numpy(), d)
            assert int(out["best_count"][p]) == int(out["mask"][p * 300:p * 300 + c].sum())
            assert int(out["best_count"][p]) > 0.5 * c          # the planted motion is found


def test_frontend_with_pose_recovers_the_planted_motion():
    """Frontend(with_pose=True): match -> select -> RANSAC -> refit -> decomposition, nothing on
    the host; the recovered rotation / translation direction match the synthetic camera motion
    (yaw 0.02 rad, t = (0.05, 0, -1) per frame) and the host pipeline (oracle decomposition)."""
    import torch
    from b200slam.frontend import Frontend, FrontendConfig, sequence_batch
    from b200slam.synthetic import tracking_sequence
    from oracle import ransac_oracle as ro
    F, N = 5, 1200
    desc, kp = tracking_sequence(F, N, seed=3)
    counts = np.full(F, N, np.int32)
    fe = Frontend(FrontendConfig(hypotheses=512, max_matches=400, with_pose=True))
    b = sequence_batch(torch.from_numpy(desc.reshape(-1, 32)).cuda(), torch.from_numpy(kp.reshape(-1, 2)).cuda(), counts, 0, F - 1, N)
    res = fe.run(b)
    torch.cuda.synchronize()
    R, t = res.R.cpu().numpy().reshape(-1, 3, 3), res.t.cpu().numpy()
    yaw = 0.02
    Rt = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
    tt = np.array([0.05, 0.0, -1.0]) / np.linalg.norm([0.05, 0.0, -1.0])
    corr, cnt, mask = res.sel.corr.cpu().numpy(), res.sel.count.cpu().numpy(), res.inlier_mask.cpu().numpy()
    for p in range(F - 1):
        assert np.abs(R[p] - Rt).max() < 2e-2, p            # estimation noise (0.5 px, forward motion); exactness is checked against the oracle below
        assert float(t[p] @ tt) > 0.99, p
        sl = slice(p * 400, p * 400 + int(cnt[p]))
        inl = mask[sl] > 0
        Ro, to = ro.decompose_essential(res.E_refit[p].cpu().numpy().reshape(3, 3), corr[sl][inl][:, :2], corr[sl][inl][:, 2:], np.eye(3))
        np.testing.assert_allclose(R[p], Ro, atol=1e-8)
        np.testing.assert_allclose(t[p], to, atol=1e-8)


def test_sequence_pipeline_equals_single_shot_frontend():
    """SequencePipeline (depth 2; both schedules: whole-step graphs on per-slot streams, and three streams with
    the kernels serialised) returns, for every submitted sequence, exactly what one Frontend.run on the same frames
    returns — through the ONE packed record download per step — also when the slots are reused with different
    data / host buffers, and eagerly (use_graph=False)."""
    import torch
    from b200slam.frontend import Frontend, FrontendConfig, SequencePipeline, sequence_batch
    from b200slam.synthetic import tracking_sequence
    F, N, S = 6, 700, 300
    cfg = FrontendConfig(hypotheses=256, max_matches=S)
    counts = np.array([700, 650, 700, 1, 700, 699], np.int32)
    seqs = [tracking_sequence(F, N, seed=s) for s in (11, 12, 13)]
    ref = []
    fe = Frontend(cfg)
    for desc, kp in seqs:
        b = sequence_batch(torch.from_numpy(desc.reshape(-1, 32)).cuda(), torch.from_numpy(kp.reshape(-1, 2)).cuda(), counts, 0, F - 1, N)
        r = fe.run(b)
        torch.cuda.synchronize()
        ref.append({k: v.cpu().numpy().copy() for k, v in (("count", r.sel.count), ("best_h", r.best_h), ("best_count", r.best_count),
                                                           ("out_q", r.sel.out_q), ("out_t", r.sel.out_t), ("out_d", r.sel.out_d),
                                                           ("mask", r.inlier_mask))})
    hosts = [(torch.from_numpy(d.reshape(-1, 32)).pin_memory(), torch.from_numpy(k.reshape(-1, 2)).pin_memory()) for d, k in seqs]
    for use_graph, schedule in ((True, "interleaved"), (True, "serial"), (False, "serial"), (False, "interleaved")):
        pipe = SequencePipeline(F, N, cfg, depth=2, use_graph=use_graph, schedule=schedule)
        order = [0, 1, 2, 1, 0, 2, 2]
        slots = []
        for i in order:                                       # submit everything first: slots are reused while in flight
            slots.append(pipe.submit(hosts[i][0], hosts[i][1], counts))
            if len(slots) >= 2:                               # read the result of the step before the one just submitted
                j = len(slots) - 2
                out = pipe.result(slots[j])
                for k, v in ref[order[j]].items():
                    n = len(v)
                    got = out[k].numpy()[:n]
                    if k in ("out_q", "out_t", "out_d", "mask"):   # only the first count[p] entries of a pair's stride are defined
                        for p in range(F - 1):
                            c = int(ref[order[j]]["count"][p])
                            assert np.array_equal(got[p * S:p * S + c], v[p * S:p * S + c]), (use_graph, j, k, p)
                    else:
                        assert np.array_equal(got, v), (use_graph, j, k)
        out = pipe.result(slots[-1])
        assert np.array_equal(out["best_h"].numpy(), ref[order[-1]]["best_h"])
        torch.cuda.synchronize()
