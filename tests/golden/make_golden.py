#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference and cv2):

    python tests/golden/make_golden.py

What is executed (nothing is copied):
  * ``cv2.BFMatcher`` (OpenCV, the third-party library the reference calls at
    feature_pipeline.py.bak:68,82,84) — knnMatch(k=2) and crossCheck match;
  * ``/root/reference/feature_pipeline.py.bak`` ``ORBFeaturePipeline.match``,
    ``match_stats``, ``adaptive_ransac_threshold`` (loaded with importlib under a
    private module name; the live ``feature_pipeline.py`` is a dangling shim);
  * ``/root/reference/homography.py`` ``match_orb_descriptors``,
    ``eight_point_E``, ``ransac_essential``, ``decompose_essential``.

Per-hypothesis inlier sets are obtained from the unmodified ``ransac_essential``
by calling it with ``max_iter=1`` and a stub generator whose ``choice`` returns a
preset 8-sample — the function then returns exactly that hypothesis's inlier set.

The fixtures pin the oracle (tests/test_oracle_golden.py) and, through the
oracle and directly, the CUDA path (tests/test_gpu_*.py).  The reference ships
no golden vectors of its own for this path.
"""

from __future__ import annotations

import importlib.machinery
import importlib.util
import sys
from pathlib import Path

import cv2
import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def _load(name: str, path: Path):
    loader = importlib.machinery.SourceFileLoader(name, str(path))
    spec = importlib.util.spec_from_loader(name, loader)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    loader.exec_module(mod)
    return mod


ref_fp = _load("_ref_feature_pipeline_bak", REF / "feature_pipeline.py.bak")
ref_h = _load("_ref_homography", REF / "homography.py")


# --------------------------------------------------------------------------- #
# descriptor generators
# --------------------------------------------------------------------------- #

def tie_heavy(rng, nq, nt, width, alphabet=4):
    return (rng.integers(0, alphabet, (nq, width), dtype=np.uint8),
            rng.integers(0, alphabet, (nt, width), dtype=np.uint8))


def noisy_copy(rng, nq, nt, width=32, keep=0.7, flip=0.08):
    """SURVEY §8d config-2 style: train = permuted noisy copy of the query set."""
    q = rng.integers(0, 256, (nq, width), dtype=np.uint8)
    t = rng.integers(0, 256, (nt, width), dtype=np.uint8)
    m = min(nq, nt)
    src = rng.permutation(nq)[:m]
    dstpos = rng.permutation(nt)[:m]
    copy = rng.random(m) < keep
    bits = np.unpackbits(q[src[copy]], axis=1)
    bits ^= (rng.random(bits.shape) < flip).astype(np.uint8)
    t[dstpos[copy]] = np.packbits(bits, axis=1)
    return q, t


def real_orb(seed, nfeatures=500):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 255, (376, 620), dtype=np.uint8)
    img = cv2.GaussianBlur(img, (0, 0), 2.0)
    img = cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX)
    img2 = np.roll(np.roll(img, 3, axis=1), 1, axis=0)
    orb = cv2.ORB_create(nfeatures=nfeatures)
    cv2.setRNGSeed(1337)
    _, d1 = orb.detectAndCompute(img, None)
    cv2.setRNGSeed(1337)
    _, d2 = orb.detectAndCompute(img2, None)
    return d1, d2


def hamming_cases():
    rng = np.random.default_rng(20261018)
    cases = {}
    for (nq, nt) in [(1, 1), (1, 5), (5, 1), (2, 2), (7, 3), (127, 129), (128, 128), (129, 127), (300, 257)]:
        for w in (1, 2, 4, 32):
            cases[f"tie_w{w}_{nq}x{nt}"] = tie_heavy(rng, nq, nt, w)
    cases["noisy_500x500"] = noisy_copy(rng, 500, 500)
    cases["noisy_1944x2000"] = noisy_copy(rng, 1944, 2000)
    cases["noisy_2000x2000"] = noisy_copy(rng, 2000, 2000)
    cases["noisy_640x33"] = noisy_copy(rng, 640, 33)
    d = rng.integers(0, 256, (50, 32), dtype=np.uint8)
    cases["identical_50"] = (d, d.copy())                       # tests/test_keyframe_manager.py:30-37
    dup = np.repeat(rng.integers(0, 256, (10, 32), dtype=np.uint8), 7, axis=0)
    cases["duplicates_70x70"] = (dup, dup[rng.permutation(70)])
    z = np.zeros((40, 32), np.uint8)
    z[::3] = 255
    cases["zeros_ones_40x64"] = (z, np.vstack([z, rng.integers(0, 256, (24, 32), dtype=np.uint8)]))
    for s in (0, 1):
        d1, d2 = real_orb(s)
        cases[f"orb_real_{s}"] = (d1, d2)
    return cases


def dm_arrays(ms):
    return (np.array([m.queryIdx for m in ms], np.int32), np.array([m.trainIdx for m in ms], np.int32),
            np.array([m.distance for m in ms], np.float32))


def gen_hamming():
    out = {}
    names = []
    for name, (q, t) in hamming_cases().items():
        names.append(name)
        out[f"{name}/q"], out[f"{name}/t"] = q, t
        knn = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
        idx = np.full((len(q), 2), -1, np.int32)
        dist = np.full((len(q), 2), -1, np.int32)
        for i, pair in enumerate(knn):
            for k, m in enumerate(pair):
                assert m.queryIdx == i
                idx[i, k], dist[i, k] = m.trainIdx, int(m.distance)
        out[f"{name}/knn_idx"], out[f"{name}/knn_dist"] = idx, dist
        cc = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(q, t)
        out[f"{name}/cc_q"], out[f"{name}/cc_t"], out[f"{name}/cc_d"] = dm_arrays(cc)
        for cross in (True, False):
            for ratio in (0.6, 0.75, 0.8, 1.0):
                for mm in (None, 1, 500):
                    if cross and ratio != 0.8:
                        continue                                  # ratio unused when cross_check (.bak:81-82)
                    cfg = ref_fp.FeaturePipelineConfig(cross_check=cross, ratio_test=ratio, max_matches=mm)
                    ms = ref_fp.ORBFeaturePipeline(cfg).match(q, t)
                    key = f"{name}/pipe_c{int(cross)}_r{ratio}_m{mm}"
                    out[key + "_q"], out[key + "_t"], out[key + "_d"] = dm_arrays(ms)
                    st = ref_fp.ORBFeaturePipeline(cfg).match_stats(ms)
                    out[key + "_stats"] = np.array([st.match_count, st.mean_distance, st.median_distance], np.float64)
        if len(t) >= 2 and len(q) <= 700:
            for ratio in (0.8, 0.6, 1.0):
                pairs = ref_h.match_orb_descriptors(q, t, ratio=ratio)
                out[f"{name}/mod_r{ratio}"] = np.array([(int(i), int(j)) for i, j in pairs], np.int32).reshape(-1, 2)
    out["names"] = np.array(names)
    np.savez_compressed(OUT / "hamming_golden.npz", **out)
    print("hamming_golden.npz:", len(names), "cases")


# --------------------------------------------------------------------------- #
# geometry scenes
# --------------------------------------------------------------------------- #

def scene(seed, n, outlier_frac, K, noise=0.0, t=(0.2, 0.0, 0.05), yaw=0.03, normalised=True):
    """Two-view scene in the style of tests/test_robust_pose_estimator.py:18-45 and
    tests/test_loop_closure_verification.py:19-40 (random 3-D points in front of
    the camera, small yaw + translation)."""
    rng = np.random.default_rng(seed)
    P = rng.uniform(-1.0, 1.0, (n, 3)) + np.array([0.0, 0.0, 4.0])
    R = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
    t = np.asarray(t, float)

    def proj(X, R, t):
        c = (R @ X.T).T + t
        if normalised:
            return c[:, :2] / c[:, 2:3]
        p = (K @ c.T).T
        return p[:, :2] / p[:, 2:3]

    p1, p2 = proj(P, np.eye(3), np.zeros(3)), proj(P, R, t)
    p2 = p2 + rng.normal(0, noise, p2.shape)
    nout = int(outlier_frac * n)
    if nout:
        o = rng.permutation(n)[:nout]
        lo, hi = p2.min(0), p2.max(0)
        p2[o] = rng.uniform(lo, hi, (nout, 2))
    return p1.astype(np.float32), p2.astype(np.float32)


class _FixedRng:
    def __init__(self, idx):
        self.idx = np.asarray(idx)

    def choice(self, n, k, replace=False):
        return self.idx


def gen_ransac():
    out, names = {}, []
    I3 = np.eye(3)
    K500 = np.array([[500.0, 0, 320.0], [0, 500.0, 240.0], [0, 0, 1.0]])
    scenes = {
        "clean_50": dict(seed=0, n=50, outlier_frac=0.0, K=I3, th=0.01),
        "noisy_200_o30": dict(seed=1, n=200, outlier_frac=0.3, K=I3, noise=0.002, th=0.01),
        "noisy_500_o40": dict(seed=2, n=500, outlier_frac=0.4, K=I3, noise=0.001, th=0.01),
        "tight_300_o50": dict(seed=3, n=300, outlier_frac=0.5, K=I3, noise=0.0005, th=0.005),
        # the K quirk replayed: pixel coordinates, K = K500, th in pixels (SURVEY finding 3)
        "quirk_k500_120": dict(seed=4, n=120, outlier_frac=0.2, K=K500, noise=0.3, th=1.0, normalised=False),
    }
    H = 48
    for name, sc in scenes.items():
        names.append(name)
        th, K = sc.pop("th"), sc["K"]
        src, dst = scene(**sc)
        n = len(src)
        out[f"{name}/src"], out[f"{name}/dst"], out[f"{name}/K"], out[f"{name}/th"] = src, dst, K, np.float64(th)
        rng = np.random.default_rng(1000 + sc["seed"])
        samples = np.stack([rng.choice(n, 8, replace=False) for _ in range(H)])
        Es = np.stack([ref_h.eight_point_E(src[i], dst[i], K) for i in samples])
        masks = np.zeros((H, n), bool)
        valid = np.zeros(H, bool)
        for h in range(H):
            try:
                _, inl = ref_h.ransac_essential(src, dst, K, th=th, max_iter=1, rng=_FixedRng(samples[h]))
                masks[h, inl] = True
                valid[h] = True
            except RuntimeError:
                pass                                              # fewer than 8 inliers -> set not observable
        out[f"{name}/samples"], out[f"{name}/E"], out[f"{name}/masks"], out[f"{name}/valid"] = samples, Es, masks, valid
        for seed in (7, 8, 9):
            for max_iter in (2000, 25):
                key = f"{name}/run_s{seed}_i{max_iter}"
                try:
                    E, inl = ref_h.ransac_essential(src, dst, K, th=th, max_iter=max_iter,
                                                    rng=np.random.default_rng(seed))
                    out[key + "_E"], out[key + "_inl"], out[key + "_ok"] = E, inl.astype(np.int64), np.bool_(True)
                except RuntimeError:
                    out[key + "_ok"] = np.bool_(False)
    out["names"] = np.array(names)

    # decompose_essential on the clean and the noisy scene
    for name in ("clean_50", "noisy_200_o30"):
        src, dst, K = out[f"{name}/src"], out[f"{name}/dst"], out[f"{name}/K"]
        E, inl = out[f"{name}/run_s7_i2000_E"], out[f"{name}/run_s7_i2000_inl"]
        R, t = ref_h.decompose_essential(E, src[inl], dst[inl], K)
        out[f"{name}/dec_R"], out[f"{name}/dec_t"] = R, t

    # adaptive threshold known answers (feature_pipeline.py.bak:114-129)
    rng = np.random.default_rng(5)
    p1 = rng.uniform(0, 1000, (64, 2)).astype(np.float32)
    ths = []
    for scale in (0.1, 5.0, 25.0, 80.0):
        p2 = (p1 + rng.normal(0, scale, p1.shape)).astype(np.float32)
        out[f"thr/p2_{scale}"] = p2
        ths.append(ref_fp.adaptive_ransac_threshold(p1, p2, 0.01, 0.005, 0.02))
    out["thr/p1"], out["thr/values"] = p1, np.array(ths)
    out["thr/empty"] = np.float64(ref_fp.adaptive_ransac_threshold(np.zeros((0,), np.float32), np.zeros((0,), np.float32), 0.01, 0.005, 0.02))
    np.savez_compressed(OUT / "ransac_golden.npz", **out)
    print("ransac_golden.npz:", names)


if __name__ == "__main__":
    print("cv2", cv2.__version__, "numpy", np.__version__)
    gen_hamming()
    gen_ransac()
