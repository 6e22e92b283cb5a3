"""CPU oracle — RANSAC homography hypothesis generation / scoring.  TEST INFRASTRUCTURE ONLY.

NumPy restatement of /root/reference/homography.py:118-216 (Hartley normalisation, 4-point /
n-point DLT, symmetric transfer error, sequential best / early-exit selection).  float64
throughout, like the reference.  Nothing in the product imports this module.
"""
from __future__ import annotations

import numpy as np


def normalise_points(pts):
    """homography.py:118-125."""
    pts = np.asarray(pts, dtype=np.float64)
    c = pts.mean(axis=0)
    diffs = pts - c
    with np.errstate(divide="ignore", invalid="ignore"):
        rms = np.sqrt((diffs ** 2).sum(axis=1).mean())
        s = np.sqrt(2) / rms
    T = np.array([[s, 0, -s * c[0]], [0, s, -s * c[1]], [0, 0, 1]])
    pts_h = np.hstack([pts, np.ones((len(pts), 1))])
    return (T @ pts_h.T).T[:, :2], T


def dlt_homography(src, dst) -> np.ndarray:
    """homography.py:131-142: normalised DLT, null vector = last right-singular vector,
    de-normalised, scaled so that H[2, 2] = 1."""
    src_n, T_src = normalise_points(src)
    dst_n, T_dst = normalise_points(dst)
    x, y, u, v = src_n[:, 0], src_n[:, 1], dst_n[:, 0], dst_n[:, 1]
    z, o = np.zeros_like(x), np.ones_like(x)
    r1 = np.stack([-x, -y, -o, z, z, z, u * x, u * y, u], axis=1)
    r2 = np.stack([z, z, z, -x, -y, -o, v * x, v * y, v], axis=1)
    A = np.stack([r1, r2], axis=1).reshape(-1, 9)
    _, _, Vt = np.linalg.svd(A)
    Hn = Vt[-1].reshape(3, 3)
    H = np.linalg.inv(T_dst) @ Hn @ T_src
    return H / H[2, 2]


def transfer_error(H, src, dst) -> np.ndarray:
    """Symmetric transfer error of homography.py:199-205 (sum of the two Euclidean norms)."""
    src, dst = np.asarray(src, dtype=np.float64), np.asarray(dst, dtype=np.float64)
    n = len(src)
    src_h, dst_h = np.hstack([src, np.ones((n, 1))]), np.hstack([dst, np.ones((n, 1))])
    with np.errstate(all="ignore"):
        pf = (H @ src_h.T).T
        pf = pf[:, :2] / pf[:, 2, None]
        pb = (np.linalg.inv(H) @ dst_h.T).T
        pb = pb[:, :2] / pb[:, 2, None]
        return np.linalg.norm(pf - dst, axis=1) + np.linalg.norm(pb - src, axis=1)


def score_hypotheses(Hs, src, dst, th: float):
    """Inlier masks (n_hyp, M) and counts for every homography (NaN error -> outlier)."""
    Hs = np.asarray(Hs, dtype=np.float64).reshape(-1, 3, 3)
    masks = np.zeros((len(Hs), len(src)), dtype=bool)
    for h, H in enumerate(Hs):
        try:
            with np.errstate(all="ignore"):
                masks[h] = transfer_error(H, src, dst) < th
        except np.linalg.LinAlgError:
            pass                                   # singular H: the reference would raise here; the device scores it as no inliers
    return masks, masks.sum(axis=1).astype(np.int64)


def select_hypothesis(counts, n: int) -> int:
    """homography.py:207-211: strictly-more-inliers update, break above 0.8 n; -1 if nothing scored."""
    best, best_count = -1, 0
    for h, c in enumerate(np.asarray(counts)):
        if c > best_count:
            best, best_count = h, int(c)
            if c > 0.8 * n:
                break
    return best


def draw_samples(rng: np.random.Generator, n: int, max_iter: int) -> np.ndarray:
    """The index stream of ransac_homography (homography.py:193)."""
    return np.stack([rng.choice(n, 4, replace=False) for _ in range(max_iter)])


def ransac_homography(src, dst, th: float = 3.0, max_iter: int = 2000, rng=None):
    """homography.py:148-216 -> (refined H, inlier indices)."""
    src, dst = np.asarray(src, dtype=np.float64), np.asarray(dst, dtype=np.float64)
    n = len(src)
    if n < 4:
        raise ValueError("At least four correspondences are required")
    if rng is None:
        rng = np.random.default_rng()
    best_H, best_inliers = None, np.array([], dtype=int)
    for _ in range(max_iter):
        idx = rng.choice(n, 4, replace=False)
        H = dlt_homography(src[idx], dst[idx])
        inliers = np.flatnonzero(transfer_error(H, src, dst) < th)
        if inliers.size > best_inliers.size:
            best_H, best_inliers = H, inliers
            if inliers.size > 0.8 * n:
                break
    if best_H is None or best_inliers.size < 4:
        raise RuntimeError("RANSAC failed — too few inliers")
    return dlt_homography(src[best_inliers], dst[best_inliers]), best_inliers
