"""CPU oracle — RANSAC essential-matrix hypothesis scoring.  TEST INFRASTRUCTURE ONLY.

NumPy restatement of the geometry half of the hot path
(/root/reference/homography.py:222-345).  float64 throughout, like the reference.
Nothing in the product imports this module (see oracle/__init__.py).
"""

from __future__ import annotations

import numpy as np


def _homog(p: np.ndarray) -> np.ndarray:
    p = np.asarray(p)
    return np.hstack([p, np.ones((len(p), 1))])


# --------------------------------------------------------------------------- #
# minimal / n-point solver
# --------------------------------------------------------------------------- #

def eight_point_E(src, dst, K) -> np.ndarray:
    """``eight_point_E`` (homography.py:222-248), quirk included.

    Points are normalised by ``K^-1`` (:228-232), each design-matrix row is
    ``[u x, u y, u, v x, v y, v, x, y, 1]`` (:235-238), the null vector is the
    last right-singular vector (:241-242), the rank-2 projection only zeroes
    the third singular value (:244-246, values NOT equalised) and the function
    returns ``K^T F K`` (:248) although F was fitted on normalised points — the
    latent intrinsics quirk of SURVEY.md finding 3, identity when K = I.
    """
    src, dst, K = np.asarray(src), np.asarray(dst), np.asarray(K, dtype=float)
    n = len(src)
    if n < 8:
        raise ValueError("Eight correspondences required")
    Kinv = np.linalg.inv(K)
    x1 = (Kinv @ _homog(src).T).T
    x2 = (Kinv @ _homog(dst).T).T
    p1 = x1 / x1[:, 2:3]
    p2 = x2 / x2[:, 2:3]
    x, y = p1[:, 0], p1[:, 1]
    u, v = p2[:, 0], p2[:, 1]
    A = np.stack([u * x, u * y, u, v * x, v * y, v, x, y, np.ones(n)], axis=1)
    _, _, Vt = np.linalg.svd(A)
    F = Vt[-1].reshape(3, 3)
    U, S, Vt = np.linalg.svd(F)
    S[2] = 0.0
    F = U @ np.diag(S) @ Vt
    return K.T @ F @ K


def eight_point_E_batch(src, dst, K, samples) -> np.ndarray:
    """E for every row of ``samples`` ((H, 8) index sets) -> (H, 3, 3)."""
    src, dst = np.asarray(src), np.asarray(dst)
    return np.stack([eight_point_E(src[idx], dst[idx], K) for idx in np.asarray(samples)])


def draw_samples(rng: np.random.Generator, n: int, max_iter: int) -> np.ndarray:
    """The index stream of ``ransac_essential`` (homography.py:325): one
    ``rng.choice(n, 8, replace=False)`` per iteration, in order."""
    return np.stack([rng.choice(n, 8, replace=False) for _ in range(max_iter)])


# --------------------------------------------------------------------------- #
# Sampson scoring + sequential selection
# --------------------------------------------------------------------------- #

def sampson_sq_err(E, src_h, dst_h) -> np.ndarray:
    """Per-correspondence squared Sampson error, literally as homography.py:328-332."""
    Ex1 = (E @ src_h.T).T
    Etx2 = (E.T @ dst_h.T).T
    err = np.abs(np.sum(dst_h * (E @ src_h.T).T, axis=1))
    denom = Ex1[:, 0] ** 2 + Ex1[:, 1] ** 2 + Etx2[:, 0] ** 2 + Etx2[:, 1] ** 2
    with np.errstate(divide="ignore", invalid="ignore"):
        return err ** 2 / denom


def score_hypotheses(Es, src, dst, th: float):
    """Inlier masks (H, M) bool and counts (H,) for every hypothesis.

    ``err < th**2`` with NaN (0/0) counting as outlier (homography.py:333).
    """
    src_h, dst_h = _homog(np.asarray(src, dtype=np.float64)), _homog(np.asarray(dst, dtype=np.float64))
    Es = np.asarray(Es, dtype=np.float64).reshape(-1, 3, 3)
    masks = np.zeros((len(Es), len(src_h)), dtype=bool)
    for h, E in enumerate(Es):
        with np.errstate(invalid="ignore"):
            masks[h] = sampson_sq_err(E, src_h, dst_h) < th ** 2
    return masks, masks.sum(axis=1).astype(np.int64)


def select_hypothesis(counts, n: int) -> int:
    """Index the reference's sequential loop ends up holding as best
    (homography.py:335-339): strictly-more-inliers update, and the loop breaks
    at the first update whose count exceeds ``0.8 * n``.  -1 if nothing ever
    beat the initial empty set (all counts 0)."""
    best, best_count = -1, 0
    for h, c in enumerate(np.asarray(counts)):
        if c > best_count:
            best, best_count = h, int(c)
            if c > 0.8 * n:
                break
    return best


def ransac_essential(src, dst, K, th: float = 0.01, max_iter: int = 2000, rng=None,
                     return_trace: bool = False):
    """``ransac_essential`` (homography.py:302-345), same loop, same rng use.

    With ``return_trace`` also returns the per-iteration
    ``(samples, hypotheses, counts, best_h)`` actually visited.
    """
    src, dst = np.asarray(src), np.asarray(dst)
    n = len(src)
    if n < 8:
        raise ValueError("At least eight correspondences are required")
    if rng is None:
        rng = np.random.default_rng()
    best_E, best_inl, best_h = None, np.array([], dtype=int), -1
    src_h, dst_h = _homog(src), _homog(dst)
    samples, hyps, counts = [], [], []
    for it in range(max_iter):
        idx = rng.choice(n, 8, replace=False)
        E = eight_point_E(src[idx], dst[idx], K)
        with np.errstate(invalid="ignore"):
            inl = np.flatnonzero(sampson_sq_err(E, src_h, dst_h) < th ** 2)
        samples.append(idx), hyps.append(E), counts.append(inl.size)
        if inl.size > best_inl.size:
            best_E, best_inl, best_h = E, inl, it
            if inl.size > 0.8 * n:
                break
    if best_E is None or best_inl.size < 8:
        raise RuntimeError("RANSAC essential matrix failed")
    refined = eight_point_E(src[best_inl], dst[best_inl], K)
    if return_trace:
        return refined, best_inl, (np.array(samples), np.array(hyps), np.array(counts), best_h)
    return refined, best_inl


# --------------------------------------------------------------------------- #
# pose recovery (next-row #2; used by the drop-in tests)
# --------------------------------------------------------------------------- #

def decompose_essential(E, src, dst, K):
    """``decompose_essential`` (homography.py:251-299): SVD, four (R, t)
    candidates, DLT triangulation per point, first candidate with the most
    points in front of both cameras."""
    E, K = np.asarray(E, dtype=float), np.asarray(K, dtype=float)
    U, _, Vt = np.linalg.svd(E)
    if np.linalg.det(U) < 0:
        U *= -1
    if np.linalg.det(Vt) < 0:
        Vt *= -1
    W = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]])
    cands = [(U @ W @ Vt, U[:, 2]), (U @ W @ Vt, -U[:, 2]),
             (U @ W.T @ Vt, U[:, 2]), (U @ W.T @ Vt, -U[:, 2])]
    src_h, dst_h = _homog(src), _homog(dst)
    P1 = K @ np.hstack([np.eye(3), np.zeros((3, 1))])
    best, best_count = None, -1
    for R, t in cands:
        P2 = K @ np.hstack([R, t.reshape(3, 1)])
        count = 0
        for a, b in zip(src_h, dst_h):
            A = np.vstack([a[0] * P1[2] - P1[0], a[1] * P1[2] - P1[1],
                           b[0] * P2[2] - P2[0], b[1] * P2[2] - P2[1]])
            X = np.linalg.svd(A)[2][-1]
            X = X[:3] / X[3]
            if X[2] > 0 and (R @ X + t)[2] > 0:
                count += 1
        if count > best_count:
            best, best_count = (R, t), count
    if best is None:
        raise RuntimeError("Essential matrix decomposition failed")
    return best
