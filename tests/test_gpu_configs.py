"""BASELINE.json configs #3-#5 at their full sizes: the tensor-core Hamming
variants against K1 (itself pinned to cv2 / the oracle in test_gpu_hamming.py) on the whole
batch, the oracle on a sample, and size-independent properties; RANSAC at H = 4096 with
thousands of correspondences against the float64 oracle."""
import numpy as np
import pytest

from oracle import hamming_oracle as ho
from oracle import ransac_oracle as ro

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def matchers():
    from b200slam import _capi
    from b200slam.frontend import HammingMatcher
    return {"popc": HammingMatcher(variant=_capi.VARIANT_POPC), "i8": HammingMatcher(variant=_capi.VARIANT_I8MMA),
            "i8s": HammingMatcher(variant=_capi.VARIANT_I8MMA1)}


def _keys(m, batch):
    k = m.knn2(batch)
    return (k.fwd_best.cpu().numpy().view(np.uint32), k.fwd_second.cpu().numpy().view(np.uint32),
            k.bwd_best.cpu().numpy().view(np.uint32))


def test_config3_loop_closure_batch_256_pairs(matchers):
    """256 candidate pairs x (2000 vs 2000), 40 % true matches: one launch per variant."""
    import torch
    from b200slam.frontend import PairBatch
    rng = np.random.default_rng(256)
    n, pairs = 2000, 256
    q = rng.integers(0, 256, (pairs, n, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (pairs, n, 32), dtype=np.uint8)
    for p in range(pairs):
        k = int(0.4 * n)
        rows = rng.permutation(n)[:k]
        bits = np.unpackbits(q[p, rng.permutation(n)[:k]], axis=1)
        bits ^= (rng.random(bits.shape) < 0.08).astype(np.uint8)
        t[p, rows] = np.packbits(bits, axis=1)
    batch = PairBatch.from_host(list(q), list(t))
    want = _keys(matchers["popc"], batch)
    for name in ("i8", "i8s"):
        got = _keys(matchers[name], batch)
        for g, w, what in zip(got, want, ("fwd_best", "fwd_second", "bwd_best")):
            np.testing.assert_array_equal(g, w, err_msg=f"{name} {what}")
    for p in (0, 17, 255):
        e = ho.packed_keys(q[p], t[p])
        for g, w in zip(want, e):
            sl = slice(p * n, (p + 1) * n)
            np.testing.assert_array_equal(g[sl], w)
    torch.cuda.synchronize()


def test_config4_dense_10k_all_variants(matchers):
    """10k x 10k per pair: variants agree bit for bit; role swap swaps forward / backward minima."""
    from b200slam.frontend import PairBatch
    rng = np.random.default_rng(4096)
    n = 10_000
    q = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    perm = rng.permutation(n)
    bits = np.unpackbits(q[perm], axis=1)
    bits ^= (rng.random(bits.shape) < 0.08).astype(np.uint8)
    t = np.packbits(bits, axis=1)
    b1, b2 = PairBatch.from_host([q], [t]), PairBatch.from_host([t], [q])
    want = _keys(matchers["popc"], b1)
    for name in ("i8", "i8s"):
        got = _keys(matchers[name], b1)
        for g, w, what in zip(got, want, ("fwd_best", "fwd_second", "bwd_best")):
            np.testing.assert_array_equal(g, w, err_msg=f"{name} {what}")
        sw = _keys(matchers[name], b2)
        np.testing.assert_array_equal(sw[0], got[2])
        np.testing.assert_array_equal(sw[2], got[0])
    rows = rng.choice(n, 100, replace=False)
    D = ho.hamming_matrix(q[rows], t)
    kf = np.sort((D.astype(np.uint32) << ho.IDX_BITS) | np.arange(n, dtype=np.uint32)[None], axis=1)
    np.testing.assert_array_equal(want[0][rows], kf[:, 0])
    np.testing.assert_array_equal(want[1][rows], kf[:, 1])


def test_config5_relocalization_sweep_4541_keyframes(matchers):
    """One query frame (2000 descriptors, ONE device copy) against 4541 keyframes x 2000
    (KITTI 00 size, 290 MB of map descriptors): per keyframe the cross-check matcher."""
    import torch
    from b200slam.frontend import PairBatch
    rng = np.random.default_rng(4541)
    n, kfs = 2000, 4541
    query = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    gen = torch.Generator(device="cuda").manual_seed(4541)
    kmap = torch.randint(0, 256, (kfs * n, 32), dtype=torch.uint8, device="cuda", generator=gen)
    # keyframes 7, 1234 and 4540 really see the query scene (noisy copies, permuted rows)
    planted = {}
    for kf in (7, 1234, 4540):
        perm = rng.permutation(n)
        bits = np.unpackbits(query[perm], axis=1)
        bits ^= (rng.random(bits.shape) < 0.08).astype(np.uint8)
        blk = np.packbits(bits, axis=1)
        kmap[kf * n:(kf + 1) * n] = torch.from_numpy(blk).cuda()
        planted[kf] = blk
    off = (np.arange(kfs + 1, dtype=np.int64) * n).astype(np.int32)
    dev_off = torch.from_numpy(off).cuda()
    batch = PairBatch(q_desc=torch.from_numpy(query).cuda(), t_desc=kmap, q_off=dev_off, t_off=dev_off,
                      q_off_host=off, t_off_host=off,
                      q_src=torch.zeros(kfs, dtype=torch.int32, device="cuda"), t_src=dev_off[:kfs].clone())
    want = _keys(matchers["popc"], batch)
    got = _keys(matchers["i8s"], batch)
    for g, w, what in zip(got, want, ("fwd_best", "fwd_second", "bwd_best")):
        np.testing.assert_array_equal(g, w, err_msg=what)
    for kf, blk in planted.items():
        e = ho.packed_keys(query, blk)
        sl = slice(kf * n, (kf + 1) * n)
        np.testing.assert_array_equal(got[0][sl], e[0])
        np.testing.assert_array_equal(got[1][sl], e[1])
        np.testing.assert_array_equal(got[2][sl], e[2])
    # per-keyframe mutual-match counts: the planted keyframes stand out, as the relocalizer expects
    fb, bb = got[0].reshape(kfs, n), got[2].reshape(kfs, n)
    j = (fb & ho.IDX_MASK).astype(np.int64)
    mutual = (np.take_along_axis(bb & ho.IDX_MASK, j, axis=1) == np.arange(n)[None]).sum(axis=1)
    top = np.argsort(-mutual)[:3]
    assert set(top.tolist()) == set(planted)
    assert mutual[top].min() > 1.5 * np.delete(mutual, top).max()      # random sets still agree mutually on ~half their rows


def test_config4_ransac_4096_hypotheses_thousands_of_correspondences():
    """H = 4096 hypotheses drawn like np.random.default_rng(4096) would, M = 6000 correspondences:
    device 8-point + float64 Sampson counts against the oracle on a hypothesis sample, winner
    selection against the oracle's sequential rule on the device's own counts."""
    import torch
    from b200slam.frontend import EssentialRansac
    rng = np.random.default_rng(4096)
    M, H = 6000, 4096
    P = np.stack([rng.uniform(-10, 10, M), rng.uniform(-2, 2, M), rng.uniform(5, 40, M)], axis=1)
    yaw = 0.03
    R = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
    P2 = P @ R.T + np.array([0.1, 0.0, -1.0])
    src = (P[:, :2] / P[:, 2:]).astype(np.float32)
    dst = (P2[:, :2] / P2[:, 2:] + rng.normal(0, 0.5 / 718.0, (M, 2))).astype(np.float32)
    out = rng.permutation(M)[: M // 2]
    dst[out] = rng.uniform(-0.8, 0.8, (len(out), 2)).astype(np.float32)
    samples = np.stack([rng.choice(M, 8, replace=False) for _ in range(H)]).astype(np.int32)
    er = EssentialRansac()
    corr = torch.from_numpy(np.hstack([src, dst])).cuda()
    c_off = torch.zeros(2, dtype=torch.int32, device="cuda")
    c_off[1] = M
    c_cnt = torch.tensor([M], dtype=torch.int32, device="cuda")
    E = er.hypotheses(corr, c_off, c_cnt, 1, H, samples=torch.from_numpy(samples).cuda()[None])
    counts = er.score(corr, c_off, c_cnt, 1, E, 0.01 ** 2)
    best_h, best_c, mask = er.select(counts, corr, c_off, c_cnt, 1, E, 0.01 ** 2)
    torch.cuda.synchronize()
    Ed, cd = E[0].cpu().numpy().reshape(H, 3, 3), counts[0].cpu().numpy()
    pick = rng.choice(H, 96, replace=False)
    # same hypotheses -> same counts (float64 on both sides)
    _, c_or = ro.score_hypotheses(Ed[pick], src, dst, 0.01)
    np.testing.assert_array_equal(cd[pick], c_or)
    # device 8-point vs the reference's SVD formulation on the same samples: inlier sets within 0.1 % of M
    E_ref = ro.eight_point_E_batch(src, dst, np.eye(3), samples[pick[:24]])
    _, c_ref = ro.score_hypotheses(E_ref, src, dst, 0.01)
    assert np.abs(c_ref - cd[pick[:24]]).max() <= 0.001 * M
    assert int(best_h[0]) == ro.select_hypothesis(cd, M)
    assert int(best_c[0]) == int(mask.sum()) == int(cd[int(best_h[0])])
    assert int(best_c[0]) > 0.4 * M


def test_config4_selection_10k_rows_equals_oracle():
    """The selection kernel's per-warp counting sort at BASELINE config #4 size: 10 000 x 10 000 (planted matches,
    8 % bit noise: hundreds of distance ties), every mode — cross-check only, ratio only, both; sorted by distance
    (stable: ties keep ascending queryIdx) and unsorted; truncated and complete — against the oracle's selection on
    the oracle's keys, plus a ragged batch around the segment boundaries of the 32 warps."""
    from b200slam import _capi
    from b200slam.frontend import HammingMatcher
    from oracle import hamming_oracle as ho
    rng = np.random.default_rng(40)
    n = 10_000
    q = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    perm = rng.permutation(n)
    bits = np.unpackbits(q[perm], axis=1)
    bits ^= (rng.random(bits.shape) < 0.08).astype(np.uint8)
    t = np.packbits(bits, axis=1)
    t[rng.permutation(n)[:3000]] = rng.integers(0, 256, (3000, 32), dtype=np.uint8)      # 30 % of the train rows are strangers
    m = HammingMatcher(variant=_capi.VARIANT_I8MMA1)
    (fb, fs, bb), = m.knn2_pairs([q], [t])
    for use_ratio, use_cross in ((False, True), (True, False), (True, True)):
        for sort, mm in ((True, None), (True, 500), (False, None), (True, 10_000)):
            (qi, ti, d), = m.match_pairs([q], [t], use_ratio=use_ratio, use_cross=use_cross, ratio=0.8, sort_by_distance=sort, max_matches=mm)
            e = ho.select_matches(fb, fs, bb, use_ratio=use_ratio, use_cross=use_cross, ratio=0.8, max_matches=mm, sort_by_distance=sort)
            np.testing.assert_array_equal(qi, e[0], err_msg=str((use_ratio, use_cross, sort, mm)))
            np.testing.assert_array_equal(ti, e[1])
            np.testing.assert_array_equal(d, e[2])
            assert len(qi) > 400
    # ragged sizes around the warps' segment boundaries (32 warps x multiples of 32 rows), tie-heavy alphabet
    sizes = [1, 31, 32, 33, 1023, 1024, 1025, 2047, 2048, 2049, 3000]
    qs = [rng.integers(0, 4, (a, 32), dtype=np.uint8) for a in sizes]
    ts = [rng.integers(0, 4, (max(2, a // 2), 32), dtype=np.uint8) for a in sizes]
    got = m.match_pairs(qs, ts, use_ratio=False, use_cross=True, sort_by_distance=True, max_matches=None)
    for a, b, (qi, ti, d) in zip(qs, ts, got):
        e = ho.select_matches(*ho.packed_keys(a, b), use_ratio=False, use_cross=True, ratio=1.0, max_matches=None, sort_by_distance=True)
        np.testing.assert_array_equal(qi, e[0], err_msg=str(len(a)))
        np.testing.assert_array_equal(ti, e[1])
        np.testing.assert_array_equal(d, e[2])
