"""GPU parity — 8-point hypotheses, float64 Sampson scoring and winner selection through
the C ABI, against fixtures produced by the unmodified reference (homography.py) and the
CPU oracle.  Tolerance (north star): <= 0.1 % of correspondences may flip at the Sampson
threshold boundary per hypothesis; float64 scoring is expected to flip none."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import ransac_oracle as ro

FLIP_TOL = 1e-3


@pytest.fixture(scope="module")
def rg(golden_dir):
    return np.load(golden_dir / "ransac_golden.npz")


@pytest.fixture(scope="module")
def R():
    from b200slam.frontend import EssentialRansac
    return EssentialRansac()


def _dev(src, dst):
    import torch
    corr = torch.from_numpy(np.hstack([src, dst]).astype(np.float32)).cuda()
    off = torch.tensor([0, len(src)], dtype=torch.int32, device="cuda")
    cnt = torch.tensor([len(src)], dtype=torch.int32, device="cuda")
    return corr, off, cnt


def _align(E, ref):
    """E is defined up to sign/scale: normalise both and align the sign."""
    E, ref = E / np.linalg.norm(E), ref / np.linalg.norm(ref)
    return (E if np.sum(E * ref) >= 0 else -E), ref


def test_scoring_reference_hypotheses_float64_exact(rg, R):
    """Identical hypotheses (the reference's own E matrices) -> identical inlier counts."""
    import torch
    for name in rg["names"]:
        src, dst, th = rg[f"{name}/src"], rg[f"{name}/dst"], float(rg[f"{name}/th"])
        Es, masks, valid = rg[f"{name}/E"], rg[f"{name}/masks"], rg[f"{name}/valid"]
        corr, off, cnt = _dev(src, dst)
        E = torch.from_numpy(Es.reshape(1, -1, 9).copy()).cuda()
        counts = R.score(corr, off, cnt, 1, E, th ** 2, precision=64).cpu().numpy()[0]
        np.testing.assert_array_equal(counts[valid], masks[valid].sum(1), err_msg=name)
        _, oc = ro.score_hypotheses(Es, src, dst, th)
        np.testing.assert_array_equal(counts, oc, err_msg=name)
        c32 = R.score(corr, off, cnt, 1, E, th ** 2, precision=32).cpu().numpy()[0]
        assert np.abs(c32 - oc).max() <= max(1, FLIP_TOL * len(src) * 5), name   # fp32 variant: reported, looser
        best_h, best_c, mask = R.select(torch.from_numpy(counts[None].astype(np.int32)).cuda(), corr, off, cnt, 1, E, th ** 2)
        want = ro.select_hypothesis(oc, len(src))
        assert int(best_h[0]) == want
        if want >= 0:
            om, _ = ro.score_hypotheses(Es[want:want + 1], src, dst, th)
            np.testing.assert_array_equal(mask.cpu().numpy().astype(bool), om[0])
            assert int(best_c[0]) == oc[want]


def test_device_eight_point_matches_reference(rg, R):
    """Device 8-point solve on the reference's sample sets: E equal up to sign/scale, and
    scoring those hypotheses flips <= 0.1 % of correspondences vs the reference's sets."""
    import torch
    for name in rg["names"]:
        src, dst, K, th = rg[f"{name}/src"], rg[f"{name}/dst"], rg[f"{name}/K"], float(rg[f"{name}/th"])
        samples, Es, masks, valid = rg[f"{name}/samples"], rg[f"{name}/E"], rg[f"{name}/masks"], rg[f"{name}/valid"]
        corr, off, cnt = _dev(src, dst)
        smp = torch.from_numpy(samples.astype(np.int32)[None].copy()).cuda()
        E = R.hypotheses(corr, off, cnt, 1, len(samples), samples=smp, K=K)
        Eh = E.cpu().numpy()[0].reshape(-1, 3, 3)
        for h in range(len(samples)):
            a, b = _align(Eh[h], Es[h])
            np.testing.assert_allclose(a, b, atol=1e-7, err_msg=f"{name} h={h}")
        counts = R.score(corr, off, cnt, 1, E, th ** 2).cpu().numpy()[0]
        om, oc = ro.score_hypotheses(Eh, src, dst, th)
        np.testing.assert_array_equal(counts, oc)                              # same E -> same counts
        flips = np.abs(counts[valid] - masks[valid].sum(1))
        assert (flips <= max(1, int(FLIP_TOL * len(src)))).all(), (name, flips.max())


def test_ransac_essential_dropin_matches_reference_runs(rg):
    """integration.pose_bridge.ransac_essential with the reference's seeded rng lands on the
    same inlier set as the reference run (golden), early-exit and full-budget cases."""
    from integration.pose_bridge import ransac_essential
    agree = total = 0
    for name in rg["names"]:
        src, dst, K, th = rg[f"{name}/src"], rg[f"{name}/dst"], rg[f"{name}/K"], float(rg[f"{name}/th"])
        for seed in (7, 8, 9):
            for max_iter in (2000, 25):
                key = f"{name}/run_s{seed}_i{max_iter}"
                if not bool(rg[key + "_ok"]):
                    with pytest.raises(RuntimeError):
                        ransac_essential(src, dst, K, th, max_iter, np.random.default_rng(seed))
                    continue
                E, inl = ransac_essential(src, dst, K, th, max_iter, np.random.default_rng(seed))
                want = rg[key + "_inl"]
                sym = len(np.setxor1d(inl, want))
                assert sym <= max(1, int(FLIP_TOL * len(src))), (key, sym)
                total += 1
                if sym == 0:
                    agree += 1
                    a, b = _align(E, rg[key + "_E"])
                    np.testing.assert_allclose(a, b, atol=1e-8, err_msg=key)
    assert total >= 20 and agree >= total - 2


def test_division_free_test_equals_literal_on_boundary(R):
    """num^2 < th^2 * den  vs the reference's literal num^2 / den < th^2, on values placed
    at and around the boundary and on den == 0 (NaN -> outlier)."""
    import torch
    rng = np.random.default_rng(5)
    src = rng.uniform(-1, 1, (4000, 2)).astype(np.float32)
    dst = rng.uniform(-1, 1, (4000, 2)).astype(np.float32)
    src[:8] = 0
    dst[:8] = 0
    Es = rng.normal(size=(16, 3, 3))
    Es[0] = 0                                   # den == 0 everywhere
    Es[1] = np.array([[0, 0, 0], [0, 0, 0], [0, 0, 1.0]])   # den == 0, num != 0
    corr, off, cnt = _dev(src, dst)
    E = torch.from_numpy(Es.reshape(1, -1, 9).copy()).cuda()
    for th in (0.5, 0.05, 1.0):
        counts = R.score(corr, off, cnt, 1, E, th ** 2).cpu().numpy()[0]
        _, oc = ro.score_hypotheses(Es, src, dst, th)
        assert counts[0] == 0 and counts[1] == 0
        assert np.abs(counts - oc).max() <= 1


def test_device_sampler_is_valid_and_deterministic(R):
    import torch
    rng = np.random.default_rng(1)
    src = rng.uniform(-1, 1, (37, 2)).astype(np.float32)
    dst = rng.uniform(-1, 1, (37, 2)).astype(np.float32)
    corr, off, cnt = _dev(src, dst)
    E1, s1 = R.hypotheses(corr, off, cnt, 1, 512, seed=42, return_samples=True)
    E2, s2 = R.hypotheses(corr, off, cnt, 1, 512, seed=42, return_samples=True)
    E3, s3 = R.hypotheses(corr, off, cnt, 1, 512, seed=43, return_samples=True)
    s1, s2, s3 = s1.cpu().numpy()[0], s2.cpu().numpy()[0], s3.cpu().numpy()[0]
    assert np.array_equal(s1, s2) and not np.array_equal(s1, s3)
    assert s1.min() >= 0 and s1.max() < 37
    assert all(len(set(row)) == 8 for row in s1)
    assert abs(np.bincount(s1.ravel(), minlength=37).std() / (512 * 8 / 37)) < 0.2   # roughly uniform
    assert torch.equal(E1, E2)


def test_too_few_correspondences(R):
    import torch
    src = np.zeros((5, 2), np.float32)
    corr, off, cnt = _dev(src, src)
    E = R.hypotheses(corr, off, cnt, 1, 32, seed=1)
    assert float(E.abs().max()) == 0.0
    counts = R.score(corr, off, cnt, 1, E, 1e-4)
    best_h, best_c, _ = R.select(counts, corr, off, cnt, 1, E, 1e-4)
    assert int(best_h[0]) == -1 and int(best_c[0]) == 0
    from integration.pose_bridge import ransac_essential
    with pytest.raises(ValueError):
        ransac_essential(src, src, np.eye(3))


def test_hybrid_scoring_equals_float64_everywhere(rg, R):
    """precision=64 (float32 screening + float64 decisions inside the rounding band) must give
    exactly the counts of precision=6464 (every evaluation in float64): golden scenes, random
    matrices with thresholds placed ON individual residuals (forced band hits), pixel-scale
    coordinates, tiny and huge E scales, degenerate E."""
    import torch
    rng = np.random.default_rng(64)
    cases = []
    for name in rg["names"]:
        cases.append((rg[f"{name}/src"], rg[f"{name}/dst"], rg[f"{name}/E"].reshape(-1, 9), float(rg[f"{name}/th"])))
    for scale, coord in ((1.0, 1.0), (1e-6, 1.0), (1e5, 1.0), (1.0, 700.0)):
        M = 3000
        src = (rng.uniform(-1, 1, (M, 2)) * coord).astype(np.float32)
        dst = (src + rng.normal(0, 0.01 * coord, (M, 2))).astype(np.float32)
        Es = rng.normal(size=(256, 9)) * scale
        Es[0] = 0
        Es[1] = np.array([0, 0, 0, 0, 0, 0, 0, 0, 1.0]) * scale
        # thresholds equal to the exact residual of some correspondence under some hypothesis
        sh, dh = np.hstack([src, np.ones((M, 1))]).astype(np.float64), np.hstack([dst, np.ones((M, 1))]).astype(np.float64)
        with np.errstate(all="ignore"):
            errs = ro.sampson_sq_err(Es[7].reshape(3, 3), sh, dh)
        for th2 in (float(np.nanmedian(errs)), float(np.nanquantile(errs, 0.1)), 1e-4 * coord ** 2):
            cases.append((src, dst, Es, float(np.sqrt(th2))))
    total_band = 0
    for src, dst, Es, th in cases:
        corr, off, cnt = _dev(src, dst)
        E = torch.from_numpy(np.ascontiguousarray(Es.reshape(1, -1, 9))).cuda()
        c_h = R.score(corr, off, cnt, 1, E, th ** 2, precision=64).cpu().numpy()[0]
        c_d = R.score(corr, off, cnt, 1, E, th ** 2, precision=6464).cpu().numpy()[0]
        np.testing.assert_array_equal(c_h, c_d)
        total_band += 1
    assert total_band == len(cases)


def test_hybrid_scoring_ragged_counts_equal_float64_and_numpy(R):
    """The packed (two correspondences per instruction) loop on every count around its group sizes (32-evaluation groups,
    64-evaluation spans, 512-correspondence chunks), in ONE ragged batch with an empty pair, thresholds on residuals of
    the last correspondences (band hits in the padded tail group), whole and sliced over the correspondences (max_m)."""
    import torch
    rng = np.random.default_rng(6464)
    Ms = [0, 1, 2, 8, 31, 32, 33, 63, 64, 65, 95, 96, 97, 127, 128, 129, 500, 511, 512, 513, 575, 576, 577, 1024 + 33, 2000]
    H = 130                                                        # two CTAs of hypotheses, the second with dead lanes
    src = [rng.uniform(-1, 1, (m, 2)).astype(np.float32) for m in Ms]
    dst = [(s_ + rng.normal(0, 0.01, s_.shape)).astype(np.float32) for s_ in src]
    Es = rng.normal(size=(len(Ms), H, 9))
    corr = torch.from_numpy(np.vstack([np.hstack([a, b]) for a, b in zip(src, dst)]).astype(np.float32)).cuda()
    off = torch.from_numpy(np.concatenate([[0], np.cumsum(Ms)]).astype(np.int32)).cuda()
    cnt = torch.tensor(Ms, dtype=torch.int32, device="cuda")
    th2 = np.full(len(Ms), 1e-4)
    for p, m in enumerate(Ms):                                     # the exact residual of the LAST correspondence under hypothesis 3
        if m:
            h = lambda a: np.hstack([a, np.ones((len(a), 1))]).astype(np.float64)
            with np.errstate(all="ignore"):
                e = ro.sampson_sq_err(Es[p, 3].reshape(3, 3), h(src[p][-1:]), h(dst[p][-1:]))
            if np.isfinite(e[0]) and e[0] > 0:
                th2[p] = e[0]
    E = torch.from_numpy(np.ascontiguousarray(Es)).cuda()
    tpp = torch.from_numpy(th2).cuda()
    ref = R.score(corr, off, cnt, len(Ms), E, 1e-4, th2_per_pair=tpp, precision=6464).cpu().numpy()
    for max_m in (0, max(Ms)):
        got = R.score(corr, off, cnt, len(Ms), E, 1e-4, th2_per_pair=tpp, precision=64, max_m=max_m).cpu().numpy()
        np.testing.assert_array_equal(got, ref, err_msg=f"max_m={max_m}")
    assert not ref[0].any()
    for p in (3, 9, 16, 19):                                       # NumPy float64 on a few pairs (boundary evaluations may flip: tolerance)
        h = lambda a: np.hstack([a, np.ones((len(a), 1))]).astype(np.float64)
        for hyp in (0, 3, 129):
            with np.errstate(all="ignore"):
                want = int(np.sum(ro.sampson_sq_err(Es[p, hyp].reshape(3, 3), h(src[p]), h(dst[p])) < th2[p]))
            assert abs(int(ref[p, hyp]) - want) <= max(1, int(FLIP_TOL * Ms[p])), (p, hyp)


def test_tensor_core_scoring_equals_float64(rg, R):
    """K3t (tcgen05 kind::tf32, hi/lo-split operands): (1) its raw accumulators stay inside the
    error bound the kernel assumes (kappa * ||coefficients|| * ||monomials||, measured here with a
    4x safety factor), (2) its counts equal the all-float64 kernel's on the golden scenes, on
    random matrices with thresholds placed on individual residuals, on pixel-scale coordinates
    and on ragged multi-pair batches."""
    import torch
    rng = np.random.default_rng(65)
    kappa = 2.0 ** -18

    def exact_forms(Es, src, dst):
        x, y, u, v = (src[:, 0].astype(np.float64), src[:, 1].astype(np.float64), dst[:, 0].astype(np.float64), dst[:, 1].astype(np.float64))
        one = np.ones_like(x)
        phi = np.stack([u * x, u * y, u, v * x, v * y, v, x, y, one], 1)
        E3 = Es.reshape(-1, 3, 3)
        num = Es.reshape(-1, 9) @ phi.T
        x1, x2 = np.stack([x, y, one], 1), np.stack([u, v, one], 1)
        a = np.einsum("hij,mj->hmi", E3, x1)
        b = np.einsum("hji,mj->hmi", E3, x2)
        den = a[..., 0] ** 2 + a[..., 1] ** 2 + b[..., 0] ** 2 + b[..., 1] ** 2
        psi = np.stack([x * x, x * y, y * y, x, y, u * u, u * v, v * v, u, v, one], 1)
        return num, den, np.linalg.norm(phi, axis=1).max(), np.linalg.norm(psi, axis=1).max()

    cases = []
    for name in rg["names"]:
        cases.append((rg[f"{name}/src"], rg[f"{name}/dst"], rg[f"{name}/E"].reshape(-1, 9), float(rg[f"{name}/th"])))
    for scale, coord, M, H in ((1.0, 1.0, 700, 300), (1e-3, 1.0, 129, 128), (1.0, 700.0, 500, 257), (50.0, 1.0, 8, 5)):
        src = (rng.uniform(-1, 1, (M, 2)) * coord).astype(np.float32)
        dst = (src + rng.normal(0, 0.01 * coord, (M, 2))).astype(np.float32)
        Es = rng.normal(size=(H, 9)) * scale
        Es[0] = 0
        sh, dh = np.hstack([src, np.ones((M, 1))]).astype(np.float64), np.hstack([dst, np.ones((M, 1))]).astype(np.float64)
        with np.errstate(all="ignore"):
            errs = ro.sampson_sq_err(Es[min(7, H - 1)].reshape(3, 3), sh, dh)
        for th2 in (float(np.nanmedian(errs)), 1e-4 * coord ** 2):
            cases.append((src, dst, Es, float(np.sqrt(th2))))
    worst = 0.0
    for src, dst, Es, th in cases:
        corr, off, cnt = _dev(src, dst)
        E = torch.from_numpy(np.ascontiguousarray(Es.reshape(1, -1, 9))).cuda()
        c_t, num, den, band = R.score_tc(corr, off, cnt, 1, E, th ** 2, max_m=len(src), debug=True)
        c_d = R.score(corr, off, cnt, 1, E, th ** 2, precision=6464).cpu().numpy()[0]
        np.testing.assert_array_equal(c_t.cpu().numpy()[0], c_d)
        en, ed, Phi, Psi = exact_forms(Es, src, dst)
        nE = np.linalg.norm(Es.reshape(-1, 9), axis=1)
        E3 = Es.reshape(-1, 3, 3)
        G = np.einsum("hka,hkb->hab", E3[:, :2, :], E3[:, :2, :])
        Gp = np.einsum("hak,hbk->hab", E3[:, :, :2], E3[:, :, :2])
        g = np.stack([G[:, 0, 0], 2 * G[:, 0, 1], G[:, 1, 1], 2 * G[:, 0, 2], 2 * G[:, 1, 2],
                      Gp[:, 0, 0], 2 * Gp[:, 0, 1], Gp[:, 1, 1], 2 * Gp[:, 0, 2], 2 * Gp[:, 1, 2], G[:, 2, 2] + Gp[:, 2, 2]], 1)
        nG = np.linalg.norm(g, axis=1)
        with np.errstate(all="ignore"):
            r1 = np.abs(num.cpu().numpy()[0] - en) / (kappa * nE[:, None] * Phi)
            r2 = np.abs(den.cpu().numpy()[0] - ed) / (kappa * nG[:, None] * Psi)
        worst = max(worst, float(np.nanmax(r1[nE > 0])), float(np.nanmax(r2[nG > 0])))
    assert worst < 0.5, worst           # the assumed bound (kappa = 2^-18) holds with a 2x margin
    # ragged multi-pair batch through the compact layout
    Ms = [0, 5, 8, 130, 500, 257]
    H = 200
    srcs = [rng.uniform(-1, 1, (m, 2)).astype(np.float32) for m in Ms]
    dsts = [(s + rng.normal(0, 0.02, s.shape)).astype(np.float32) for s in srcs]
    stride = 512
    corr = torch.zeros((len(Ms) * stride, 4), dtype=torch.float32, device="cuda")
    for p, (s, d) in enumerate(zip(srcs, dsts)):
        if len(s):
            corr[p * stride:p * stride + len(s)] = torch.from_numpy(np.hstack([s, d])).cuda()
    c_off = (torch.arange(len(Ms) + 1, dtype=torch.int32, device="cuda") * stride).contiguous()
    c_cnt = torch.tensor(Ms, dtype=torch.int32, device="cuda")
    E = torch.from_numpy(rng.normal(size=(len(Ms), H, 9))).cuda()
    c_t = R.score_tc(corr, c_off, c_cnt, len(Ms), E, 0.05 ** 2, max_m=stride).cpu().numpy()
    c_d = R.score(corr, c_off, c_cnt, len(Ms), E, 0.05 ** 2, precision=6464).cpu().numpy()
    for p, m in enumerate(Ms):
        np.testing.assert_array_equal(c_t[p], c_d[p], err_msg=str(m))


def test_device_decompose_essential_matches_reference(rg):
    """K7: (R, t) of the device's decompose_essential against the reference's golden result and
    the oracle's, and the per-candidate cheirality votes against a NumPy restatement of the vote."""
    import torch
    from b200slam.frontend import PoseRecovery
    from b200slam.geometry import eight_point_refit
    P = PoseRecovery()
    rng = np.random.default_rng(77)
    cases = []
    for name in rg["names"]:
        src, dst, K = rg[f"{name}/src"], rg[f"{name}/dst"], rg[f"{name}/K"]
        inl = rg[f"{name}/run_s7_i2000_inl"]
        if f"{name}/dec_R" in rg.files and len(inl) >= 8:
            cases.append((name, src[inl], dst[inl], K, rg[f"{name}/run_s7_i2000_E"], rg[f"{name}/dec_R"], rg[f"{name}/dec_t"]))
    assert cases
    n_pairs = len(cases)
    Ms = [len(c[1]) for c in cases]
    stride = max(Ms)
    corr = torch.zeros((n_pairs * stride, 4), dtype=torch.float32, device="cuda")
    for p, c in enumerate(cases):
        corr[p * stride:p * stride + Ms[p]] = torch.from_numpy(np.hstack([c[1], c[2]]).astype(np.float32)).cuda()
    c_off = (torch.arange(n_pairs + 1, dtype=torch.int32, device="cuda") * stride).contiguous()
    c_cnt = torch.tensor(Ms, dtype=torch.int32, device="cuda")
    for p, (name, src, dst, K, E, Rw, tw) in enumerate(cases):
        # one pair at a time (K differs between the scenes)
        R, t, votes = P.decompose(torch.from_numpy(np.ascontiguousarray(E.reshape(1, 9))).cuda(), corr[p * stride:], c_off[:2] * 0 + torch.tensor([0, stride], dtype=torch.int32, device="cuda"),
                                  c_cnt[p:p + 1], 1, stride, K=K)
        Ro, to = ro.decompose_essential(E, src, dst, K)
        np.testing.assert_allclose(R[0], Ro, atol=1e-8, err_msg=name)
        np.testing.assert_allclose(t[0], to, atol=1e-8, err_msg=name)
        if Rw.shape == (3, 3):
            np.testing.assert_allclose(R[0], Rw, atol=1e-8, err_msg=name)
            np.testing.assert_allclose(t[0], tw, atol=1e-8, err_msg=name)
        assert votes[0].max() > 0 and votes[0].sum() <= 4 * len(src), (name, votes)
    # batched, with inlier masks, identity intrinsics: votes equal a NumPy DLT vote on every candidate
    n_pairs, M = 5, 200
    Es, srcs, dsts, masks = [], [], [], []
    for p in range(n_pairs):
        Pw = np.stack([rng.uniform(-3, 3, M), rng.uniform(-2, 2, M), rng.uniform(4, 20, M)], axis=1)
        yaw = rng.uniform(-0.1, 0.1)
        Rt = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
        tt = rng.normal(size=3)
        tt /= np.linalg.norm(tt)
        P2 = Pw @ Rt.T + tt
        s = (Pw[:, :2] / Pw[:, 2:]).astype(np.float32)
        d = (P2[:, :2] / P2[:, 2:] + rng.normal(0, 1e-3, (M, 2))).astype(np.float32)
        mk = (rng.random(M) < 0.7).astype(np.uint8)
        Es.append(eight_point_refit(s[mk > 0], d[mk > 0], np.eye(3)).reshape(9))
        srcs.append(s), dsts.append(d), masks.append(mk)
    corr = torch.from_numpy(np.hstack([np.concatenate(srcs), np.concatenate(dsts)])).cuda()
    c_off = (torch.arange(n_pairs + 1, dtype=torch.int32, device="cuda") * M).contiguous()
    c_cnt = torch.full((n_pairs,), M, dtype=torch.int32, device="cuda")
    R, t, votes = P.decompose(torch.from_numpy(np.stack(Es)).cuda(), corr, c_off, c_cnt, n_pairs, M, mask=torch.from_numpy(np.concatenate(masks)).cuda())
    for p in range(n_pairs):
        mk = masks[p] > 0
        Ro, to = ro.decompose_essential(Es[p].reshape(3, 3), srcs[p][mk], dsts[p][mk], np.eye(3))
        np.testing.assert_allclose(R[p], Ro, atol=1e-8)
        np.testing.assert_allclose(t[p], to, atol=1e-8)
        assert votes[p].max() >= 0.95 * mk.sum() and np.sort(votes[p])[-2] < 0.5 * mk.sum()


def test_device_refit_matches_reference_runs(rg):
    """Refit of E on the winner's inliers (homography.py:344): the device's Gram-matrix + Jacobi
    solution against the refined E the unmodified reference returned for its seeded runs."""
    import torch
    from b200slam.frontend import PoseRecovery
    P = PoseRecovery()
    checked = 0
    for name in rg["names"]:
        src, dst, K = rg[f"{name}/src"], rg[f"{name}/dst"], rg[f"{name}/K"]
        for run in ("run_s7_i2000", "run_s8_i2000", "run_s9_i25"):
            if not bool(rg[f"{name}/{run}_ok"]):
                continue
            inl, Eref = rg[f"{name}/{run}_inl"], rg[f"{name}/{run}_E"]
            mask = np.zeros(len(src), np.uint8)
            mask[inl] = 1
            corr, off, cnt = _dev(src, dst)
            E, used = P.refit(corr, off, cnt, 1, mask=torch.from_numpy(mask).cuda(), K=K)
            assert int(used[0]) == len(inl)
            a, b = _align(E[0].cpu().numpy().reshape(3, 3), Eref)
            np.testing.assert_allclose(a, b, atol=1e-8, err_msg=f"{name} {run}")
            checked += 1
    assert checked >= 8
    # fewer than 8 inliers -> zeros, like the guard of eight_point_E (:224-225) would stop the reference
    corr, off, cnt = _dev(src[:5], dst[:5])
    E, used = P.refit(corr, off, cnt, 1)
    assert int(used[0]) == 5 and (E.cpu().numpy() == 0).all()
