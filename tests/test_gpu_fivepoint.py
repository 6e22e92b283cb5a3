"""K8 (5-point minimal solver) against cv2.findEssentialMat's solution sets on exactly five
points (tests/golden/fivepoint_golden.npz) and the NumPy oracle; end-to-end as a hypothesis
generator for the scoring kernel."""
import numpy as np
import pytest

from oracle import fivepoint_oracle as fp
from oracle import ransac_oracle as ro

pytestmark = pytest.mark.gpu


def test_solutions_match_cv2_and_oracle(golden_dir):
    import torch
    from b200slam.frontend import EssentialRansac
    g = np.load(golden_dir / "fivepoint_golden.npz")
    n = int(g["n"])
    R = EssentialRansac()
    corr = torch.from_numpy(np.concatenate([np.hstack([g[f"c{k}/src"], g[f"c{k}/dst"]]) for k in range(n)]).astype(np.float32)).cuda()
    c_off = (torch.arange(n + 1, dtype=torch.int32, device="cuda") * 5).contiguous()
    c_cnt = torch.full((n,), 5, dtype=torch.int32, device="cuda")
    samples = torch.arange(5, dtype=torch.int32, device="cuda").repeat(n, 1, 1).contiguous()
    E, ns = R.hypotheses_5pt(corr, c_off, c_cnt, n, 1, samples=samples, return_counts=True)
    E, ns = E.cpu().numpy().reshape(n, 10, 3, 3), ns.cpu().numpy()[:, 0]
    total = hit = extra = 0
    for k in range(n):
        ref = g[f"c{k}/E"]
        got = E[k, :ns[k]]
        assert (E[k, ns[k]:] == 0).all()
        total += len(ref)
        hit += fp.match_solution_sets(ref, got, 1e-6)
        extra += max(0, len(got) - len(ref))
        for e in got:                                  # every returned matrix is a solution of the minimal problem
            src, dst = g[f"c{k}/src"].astype(np.float64), g[f"c{k}/dst"].astype(np.float64)
            res = np.abs(np.einsum("ni,ij,nj->n", np.hstack([dst, np.ones((5, 1))]), e, np.hstack([src, np.ones((5, 1))])))
            assert res.max() < 1e-9
            assert abs(np.linalg.det(e)) < 1e-6
            assert np.abs(2 * e @ e.T @ e - np.trace(e @ e.T) * e).max() < 1e-6
    assert total > 200 and hit >= total - 3 and extra <= 3, (total, hit, extra)


def test_five_point_ransac_finds_the_motion():
    import torch
    from b200slam.frontend import EssentialRansac
    rng = np.random.default_rng(55)
    M, S = 400, 256
    P = np.stack([rng.uniform(-6, 6, M), rng.uniform(-2, 2, M), rng.uniform(5, 30, M)], axis=1)
    yaw = 0.05
    Rt = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
    P2 = P @ Rt.T + np.array([0.4, 0.05, -0.6])
    src = (P[:, :2] / P[:, 2:]).astype(np.float32)
    dst = (P2[:, :2] / P2[:, 2:] + rng.normal(0, 5e-4, (M, 2))).astype(np.float32)
    out = rng.permutation(M)[: M // 3]
    dst[out] = rng.uniform(-0.5, 0.5, (len(out), 2)).astype(np.float32)
    R = EssentialRansac()
    corr = torch.from_numpy(np.hstack([src, dst])).cuda()
    c_off = torch.tensor([0, M], dtype=torch.int32, device="cuda")
    c_cnt = torch.tensor([M], dtype=torch.int32, device="cuda")
    E, ns = R.hypotheses_5pt(corr, c_off, c_cnt, 1, S, seed=5, return_counts=True)
    E2 = R.hypotheses_5pt(corr, c_off, c_cnt, 1, S, seed=5)
    assert torch.equal(E, E2)                                   # deterministic in the seed
    assert 1.0 < float(ns.float().mean()) < 9.0
    counts = R.score(corr, c_off, c_cnt, 1, E, 0.005 ** 2)
    best_h, best_c, mask = R.select(counts, corr, c_off, c_cnt, 1, E, 0.005 ** 2)
    cd = counts[0].cpu().numpy()
    _, oc = ro.score_hypotheses(E[0].cpu().numpy().reshape(-1, 3, 3)[:400], src, dst, 0.005)
    np.testing.assert_array_equal(cd[:400], oc)
    assert int(best_c[0]) > 0.6 * M
    inl = mask.cpu().numpy().astype(bool)
    assert inl[np.setdiff1d(np.arange(M), out)].mean() > 0.9 and inl[out].mean() < 0.1


def test_find_essential_mat_agrees_with_cv2_on_inliers():
    """pose_bridge.find_essential_mat (device 5-point RANSAC) against cv2.findEssentialMat on a
    synthetic pixel-coordinate scene: same inlier set up to threshold-boundary points, same E up
    to scale and sign on the calibrated points."""
    import cv2
    from integration.pose_bridge import find_essential_mat
    rng = np.random.default_rng(19)
    M = 300
    K = np.array([[718.856, 0, 607.1928], [0, 718.856, 185.2157], [0, 0, 1.0]])
    P = np.stack([rng.uniform(-8, 8, M), rng.uniform(-2, 2, M), rng.uniform(6, 40, M)], axis=1)
    yaw = 0.04
    Rt = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
    P2 = P @ Rt.T + np.array([0.3, 0.0, -0.9])
    px1 = (P / P[:, 2:]) @ K.T
    px2 = (P2 / P2[:, 2:]) @ K.T
    p1 = px1[:, :2]
    p2 = px2[:, :2] + rng.normal(0, 0.3, (M, 2))
    out = rng.permutation(M)[: M // 4]
    p2[out] = np.stack([rng.uniform(0, 1241, len(out)), rng.uniform(0, 376, len(out))], axis=1)
    E, mask = find_essential_mat(p1, p2, K, threshold=1.0, samples=512, seed=3)
    Ec, mc = cv2.findEssentialMat(p1, p2, K, method=cv2.RANSAC, prob=0.999, threshold=1.0)
    assert E is not None and mask.shape == (M, 1)
    agree = (mask.ravel() > 0) == (mc.ravel() > 0)
    assert agree.mean() > 0.9          # both keep an unrefined minimal-sample E: points near the 1 px boundary differ
    assert mask[out].mean() < 0.05 and mask[np.setdiff1d(np.arange(M), out)].mean() > 0.9
    a, b = E / np.linalg.norm(E), Ec[:3] / np.linalg.norm(Ec[:3])
    assert min(np.abs(a - b).max(), np.abs(a + b).max()) < 0.05
