"""GPU: winner-only RANSAC scoring (b2s_ransac_winner_batched) returns EXACTLY the winner, inlier count and inlier
mask of the full evaluation (b2s_ransac_score_batched precision 64 + b2s_ransac_select) — the reference's rule
(first hypothesis above 0.8 n, else the lowest index among the maximum, homography.py:335-339) — while finishing only
the hypotheses that can still win."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _both(corr, c_off, c_cnt, n, E, th2, th2_pp=None):
    from b200slam.frontend import EssentialRansac
    R = EssentialRansac()
    counts = R.score(corr, c_off, c_cnt, n, E, th2, th2_per_pair=th2_pp, precision=64)
    full = R.select(counts, corr, c_off, c_cnt, n, E, th2, th2_per_pair=th2_pp)
    win = R.winner(corr, c_off, c_cnt, n, E, th2, th2_per_pair=th2_pp, return_counts=True)
    return counts, full, win


def _check(counts, full, win, c_cnt_host, H):
    for a, b, name in zip(full, win[:3], ("best_h", "best_count", "mask")):
        np.testing.assert_array_equal(a.cpu().numpy(), b.cpu().numpy(), err_msg=name)
    c_full, c_win, nfin = counts.cpu().numpy(), win[3].cpu().numpy(), win[4].cpu().numpy()
    assert (c_win <= c_full).all()                         # abandoned hypotheses keep a lower bound ...
    bh = full[0].cpu().numpy()
    for p in range(len(bh)):
        if bh[p] >= 0:
            assert c_win[p, bh[p]] == c_full[p, bh[p]]     # ... the winner's count is complete
    return nfin


def test_winner_only_equals_full_on_tracking_batch():
    import torch
    from b200slam.frontend import Frontend, FrontendConfig, sequence_batch
    from b200slam.synthetic import tracking_sequence
    F, N = 13, 2000
    desc, kp = tracking_sequence(F, N, seed=1234)
    b = sequence_batch(torch.from_numpy(desc.reshape(-1, 32)).cuda(), torch.from_numpy(kp.reshape(-1, 2)).cuda(), np.full(F, N, np.int32), 0, F - 1, N)
    fe = Frontend(FrontendConfig(hypotheses=2000, max_matches=500))
    res = fe.run(b)
    sel, E = res.sel, res.E
    for th in (0.01, 0.003, 0.0005, 0.05):
        counts, full, win = _both(sel.corr, sel.c_off, sel.count, b.n_pairs, E, th * th)
        nfin = _check(counts, full, win, None, 2000)
        assert (nfin < 2000).all()
    # Frontend(winner_only=True) == Frontend()
    few = Frontend(FrontendConfig(hypotheses=2000, max_matches=500, winner_only=True)).run(b)
    np.testing.assert_array_equal(few.best_h.cpu().numpy(), res.best_h.cpu().numpy())
    np.testing.assert_array_equal(few.best_count.cpu().numpy(), res.best_count.cpu().numpy())
    np.testing.assert_array_equal(few.inlier_mask.cpu().numpy(), res.inlier_mask.cpu().numpy())
    assert few.counts is None


def test_winner_only_on_golden_scenes_and_adversarial_batches(golden_dir):
    """Ragged batch: golden two-view scenes (early exit present and absent), M < 8, M <= 64 (no second pass), M = 65,
    all-zero hypotheses, one perfect hypothesis at the END (every prefix bound must keep it), ties at the maximum."""
    import torch
    rg = np.load(golden_dir / "ransac_golden.npz")
    rng = np.random.default_rng(3)
    H = 300
    scenes = sorted({k.split("/")[0] for k in rg.files if k.endswith("/src")})
    srcs, dsts, Es = [], [], []
    for name in scenes:                                     # the reference's own hypotheses (eight_point_E on its seeded samples)
        src, dst = rg[f"{name}/src"].astype(np.float32), rg[f"{name}/dst"].astype(np.float32)
        E = np.nan_to_num(rg[f"{name}/E"].reshape(-1, 9))
        E = np.concatenate([E, rng.normal(size=(max(0, H - len(E)), 9))])[:H]
        srcs.append(src), dsts.append(dst), Es.append(E)
    assert len(srcs) >= 3
    # synthetic pairs with a planted essential matrix
    def planted(M, frac_in, where):
        P = np.stack([rng.uniform(-3, 3, M), rng.uniform(-2, 2, M), rng.uniform(4, 20, M)], axis=1)
        t = np.array([0.3, 0.05, -0.6])
        p1 = (P[:, :2] / P[:, 2:]).astype(np.float32)
        P2 = P + t
        p2 = (P2[:, :2] / P2[:, 2:]).astype(np.float32)
        bad = rng.permutation(M)[: int(M * (1 - frac_in))]
        p2[bad] += rng.uniform(-0.3, 0.3, (len(bad), 2)).astype(np.float32)
        tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
        E = rng.normal(size=(H, 9))
        for w in where:
            E[w] = tx.reshape(-1) * rng.uniform(0.5, 2.0)
        return p1, p2, E
    for M, frac, where in ((500, 0.9, [H - 1]), (500, 0.6, [7, 200]), (500, 0.6, [150, 151]), (65, 0.7, [3]), (64, 0.7, [10]),
                           (7, 1.0, [0]), (1200, 0.85, [H - 2]), (33, 0.5, []), (400, 0.3, [299, 0])):
        a, b_, E = planted(M, frac, where)
        srcs.append(a), dsts.append(b_), Es.append(E)
    a, b_, E = planted(300, 0.8, [])
    srcs.append(a), dsts.append(b_), Es.append(np.zeros((H, 9)))            # nothing scores: best_h = -1
    n = len(srcs)
    cnt = np.array([len(s) for s in srcs], np.int32)
    off = np.zeros(n + 1, np.int32)
    np.cumsum(cnt, out=off[1:])
    corr = np.concatenate([np.hstack([s, d]) for s, d in zip(srcs, dsts)]).astype(np.float32)
    corr_d, off_d, cnt_d = torch.from_numpy(corr).cuda(), torch.from_numpy(off).cuda(), torch.from_numpy(cnt).cuda()
    E_d = torch.from_numpy(np.stack(Es)).cuda()
    for th in (0.01, 0.002):
        counts, full, win = _both(corr_d, off_d, cnt_d, n, E_d, th * th)
        _check(counts, full, win, cnt, H)
    th_pp = torch.from_numpy(rng.uniform(0.003, 0.03, n) ** 2).cuda()
    counts, full, win = _both(corr_d, off_d, cnt_d, n, E_d, 0.0, th_pp)
    _check(counts, full, win, cnt, H)
    assert int(full[0].cpu().numpy()[-1]) == -1
