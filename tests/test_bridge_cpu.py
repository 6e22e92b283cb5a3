"""CPU: host logic of the drop-in boundary (config dataclasses, empty inputs, DMatch
veneer, host geometry vs the reference's golden outputs)."""
import dataclasses
import pickle

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from integration import feature_pipeline_bridge as fpb
from integration import pose_bridge as pb


def test_exports_exactly_the_shim_names():
    for n in ("FeaturePipelineConfig", "MatchStats", "FeaturePipeline", "build_feature_pipeline",
              "matches_to_points", "adaptive_ransac_threshold"):
        assert hasattr(fpb, n)


def test_config_fields_defaults_validation():
    cfg = fpb.FeaturePipelineConfig()
    assert [f.name for f in dataclasses.fields(cfg)] == ["name", "nfeatures", "ratio_test", "cross_check", "max_matches", "deterministic_seed"]
    assert dataclasses.asdict(cfg) == dict(name="orb", nfeatures=2000, ratio_test=0.8, cross_check=True, max_matches=500, deterministic_seed=1337)
    assert pickle.loads(pickle.dumps(cfg)) == cfg
    assert dataclasses.replace(cfg, deterministic_seed=5).deterministic_seed == 5
    with pytest.raises(dataclasses.FrozenInstanceError):
        cfg.nfeatures = 3
    for bad in (dict(name=""), dict(nfeatures=0), dict(ratio_test=0.0), dict(ratio_test=1.5), dict(max_matches=0), dict(deterministic_seed=-1)):
        with pytest.raises(ValueError):
            fpb.FeaturePipelineConfig(**bad)
    with pytest.raises(ValueError):
        fpb.build_feature_pipeline(fpb.FeaturePipelineConfig(name="sift"))


def test_empty_inputs_need_no_device():
    pipe = fpb.build_feature_pipeline(fpb.FeaturePipelineConfig(nfeatures=100))
    d = np.zeros((3, 32), np.uint8)
    assert pipe.match(None, d) == [] and pipe.match(d, None) == [] and pipe.match(d[:0], d) == []
    assert pipe.match_stats([]) == fpb.MatchStats(0, 0.0, 0.0)
    assert pb.CrossCheckMatcher().match(None, d) == []
    img = np.random.default_rng(0).integers(0, 255, (120, 160), dtype=np.uint8)
    kp, desc = pipe.detect_and_describe(img)
    assert desc is None or (desc.dtype == np.uint8 and desc.shape[1] == 32)


def test_dmatch_veneer_and_stats():
    ms = fpb.to_dmatches(np.array([3, 1]), np.array([7, 2]), np.array([10, 40]))
    assert [(m.queryIdx, m.trainIdx, m.imgIdx, m.distance) for m in ms] == [(3, 7, 0, 10.0), (1, 2, 0, 40.0)]
    st = fpb.FeaturePipeline().match_stats(ms)
    assert st == fpb.MatchStats(2, 25.0, 25.0)
    kp = [cv2.KeyPoint(float(i), float(2 * i), 1.0) for i in range(10)]
    p1, p2 = fpb.matches_to_points(kp, kp, ms)
    assert p1.dtype == np.float32 and p1.tolist() == [[3.0, 6.0], [1.0, 2.0]] and p2.tolist() == [[7.0, 14.0], [2.0, 4.0]]
    e1, _ = fpb.matches_to_points(kp, kp, [])
    assert e1.shape == (0,)


def test_adaptive_threshold_golden(golden_dir):
    rg = np.load(golden_dir / "ransac_golden.npz")
    p1 = rg["thr/p1"]
    for scale, want in zip((0.1, 5.0, 25.0, 80.0), rg["thr/values"]):
        assert fpb.adaptive_ransac_threshold(p1, rg[f"thr/p2_{scale}"], 0.01, 0.005, 0.02) == want
    z = np.zeros((0,), np.float32)
    assert fpb.adaptive_ransac_threshold(z, z, 0.01, 0.005, 0.02) == float(rg["thr/empty"])


def test_host_refit_and_decompose_match_reference(golden_dir):
    from b200slam.geometry import decompose_essential, eight_point_refit
    rg = np.load(golden_dir / "ransac_golden.npz")
    for name in rg["names"]:
        src, dst, K = rg[f"{name}/src"], rg[f"{name}/dst"], rg[f"{name}/K"]
        key = f"{name}/run_s7_i2000"
        inl = rg[key + "_inl"]
        np.testing.assert_allclose(eight_point_refit(src[inl], dst[inl], K), rg[key + "_E"], rtol=0, atol=1e-12)
    for name in ("clean_50", "noisy_200_o30"):
        src, dst, K = rg[f"{name}/src"], rg[f"{name}/dst"], rg[f"{name}/K"]
        E, inl = rg[f"{name}/run_s7_i2000_E"], rg[f"{name}/run_s7_i2000_inl"]
        R, t = decompose_essential(E, src[inl], dst[inl], K)
        np.testing.assert_allclose(R, rg[f"{name}/dec_R"], atol=1e-9)
        np.testing.assert_allclose(t, rg[f"{name}/dec_t"], atol=1e-9)


def test_host_homography_refit_and_decomposition_recover_plane():
    """The host pieces that follow the device homography RANSAC (n-point DLT refit, decompose_homography) behind the
    oracle's CPU loop (the product has no CPU RANSAC: geometry.ransac_homography moved to oracle/)."""
    from b200slam.geometry import decompose_homography, dlt_homography_batch
    from oracle.homography_oracle import ransac_homography
    rng = np.random.default_rng(3)
    Ht = np.array([[1.02, 0.01, 5.0], [-0.02, 0.98, -3.0], [1e-5, 2e-5, 1.0]])
    src = rng.uniform(0, 640, (200, 2))
    d = np.hstack([src, np.ones((200, 1))]) @ Ht.T
    dst = d[:, :2] / d[:, 2:]
    dst[:50] += rng.uniform(-80, 80, (50, 2))
    H, inl = ransac_homography(src.astype(np.float32), dst.astype(np.float32), rng=np.random.default_rng(1))
    assert set(range(50, 200)) <= set(inl.tolist())
    np.testing.assert_allclose(H, Ht, atol=1e-3, rtol=1e-3)
    np.testing.assert_allclose(dlt_homography_batch(src.astype(np.float32).astype(np.float64)[inl][None], dst.astype(np.float32).astype(np.float64)[inl][None])[0], H, atol=1e-8, rtol=1e-8)
    R, t = decompose_homography(H)
    np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-9)
    with pytest.raises(ValueError):
        ransac_homography(src[:3], dst[:3])


def test_pose_estimator_config_and_gate_order():
    cfg = pb.RobustPoseEstimatorConfig()
    assert (cfg.min_matches, cfg.min_inliers, cfg.base_ransac_threshold, cfg.min_cheirality_inliers) == (20, 30, 0.01, 12)
    for bad in (dict(min_matches=0), dict(min_inliers=0), dict(min_inlier_ratio=0), dict(min_parallax=-1),
                dict(min_cheirality_ratio=0), dict(min_cheirality_inliers=0)):
        with pytest.raises(ValueError):
            pb.RobustPoseEstimatorConfig(**bad)
    est = pb.RobustPoseEstimator(cfg)
    mk = lambda **kw: pb.PoseEstimate(np.eye(3), np.array([1.0, 0, 0]), np.arange(3), pb.PoseEstimationDiagnostics(**{
        **dict(method="essential", match_count=100, inliers=50, inlier_ratio=0.5, median_parallax=5.0,
               cheirality_inliers=40, cheirality_ratio=0.9, score=1.0), **kw}))
    est._apply_stability_gates(mk())
    for kw, reason in ((dict(inliers=10, inlier_ratio=0.1), "low_inlier_count"), (dict(inlier_ratio=0.1), "low_inlier_ratio"),
                       (dict(median_parallax=0.5), "low_parallax"), (dict(cheirality_inliers=5), "cheirality_inliers"),
                       (dict(cheirality_ratio=0.2), "cheirality_ratio")):
        with pytest.raises(pb.PoseEstimationFailure) as e:
            est._apply_stability_gates(mk(**kw))
        assert e.value.reason == reason and e.value.recovery_action == "relocalize"
    est._apply_stability_gates(mk(method="homography", cheirality_ratio=0.0, cheirality_inliers=0))
    with pytest.raises(ValueError):
        est.estimate_pose([1], [1], [], np.eye(4))
