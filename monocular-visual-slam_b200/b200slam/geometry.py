"""Host-side two-view geometry that follows the device kernels in the pose path.

These are the steps of the reference that run ONCE per frame pair after the batched
kernels have picked the winning hypothesis (SURVEY.md §8 a8 tail, a11, a12) — the refit
on all inliers, the cheirality vote and the parallax statistics.  Vectorised NumPy
(batched LAPACK), float64 like the reference.  They are next-row #2 candidates for a
device kernel; they are not a fallback for anything that has one.
"""
from __future__ import annotations

import numpy as np


def _homog(p: np.ndarray) -> np.ndarray:
    return np.hstack([p, np.ones((len(p), 1))])


def eight_point_refit(src: np.ndarray, dst: np.ndarray, K: np.ndarray) -> np.ndarray:
    """n-point least-squares E on all inliers == the reference's final
    ``eight_point_E(src[best_inliers], dst[best_inliers], K)`` (homography.py:344, :222-248),
    ``K^T F K`` return included."""
    n = len(src)
    if n < 8:
        raise ValueError("Eight correspondences required")
    K = np.asarray(K, dtype=np.float64)
    Kinv = np.linalg.inv(K)
    a = (Kinv @ _homog(np.asarray(src)).T).T
    b = (Kinv @ _homog(np.asarray(dst)).T).T
    a = a / a[:, 2:3]
    b = b / b[:, 2:3]
    x, y, u, v = a[:, 0], a[:, 1], b[:, 0], b[:, 1]
    A = np.stack([u * x, u * y, u, v * x, v * y, v, x, y, np.ones(n)], axis=1)
    F = np.linalg.svd(A, full_matrices=False)[2][-1].reshape(3, 3)    # economy form: the reference's full U is n x n and unused
    U, S, Vt = np.linalg.svd(F)
    S[2] = 0.0
    return K.T @ (U @ np.diag(S) @ Vt) @ K


def decompose_essential(E: np.ndarray, src: np.ndarray, dst: np.ndarray, K: np.ndarray):
    """(R, t) with the most triangulated points in front of both cameras; first candidate
    wins ties (homography.py:251-299).  The per-point 4x4 DLT SVDs of the reference's
    Python double loop are issued as one batched SVD per candidate."""
    E, K = np.asarray(E, dtype=np.float64), np.asarray(K, dtype=np.float64)
    U, _, Vt = np.linalg.svd(E)
    if np.linalg.det(U) < 0:
        U = -U
    if np.linalg.det(Vt) < 0:
        Vt = -Vt
    W = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    cands = [(U @ W @ Vt, U[:, 2]), (U @ W @ Vt, -U[:, 2]), (U @ W.T @ Vt, U[:, 2]), (U @ W.T @ Vt, -U[:, 2])]
    s, d = np.asarray(src, dtype=np.float64), np.asarray(dst, dtype=np.float64)
    P1 = K @ np.hstack([np.eye(3), np.zeros((3, 1))])
    best, best_count = None, -1
    for R, t in cands:
        P2 = K @ np.hstack([R, t.reshape(3, 1)])
        A = np.stack([s[:, 0:1] * P1[2] - P1[0], s[:, 1:2] * P1[2] - P1[1],
                      d[:, 0:1] * P2[2] - P2[0], d[:, 1:2] * P2[2] - P2[1]], axis=1)      # (M,4,4)
        if len(A):
            X = np.linalg.svd(A)[2][:, -1, :]
            with np.errstate(divide="ignore", invalid="ignore"):
                X3 = X[:, :3] / X[:, 3:4]
                z2 = (X3 @ R.T + t)[:, 2]
                count = int(np.sum((X3[:, 2] > 0) & (z2 > 0)))
        else:
            count = 0
        if count > best_count:
            best, best_count = (R, t), count
    if best is None:
        raise RuntimeError("Essential matrix decomposition failed")
    return best


# ---- homography branch: host pieces that follow the device RANSAC (K5 / K6) -------------------------

def _normalise_batch(p: np.ndarray):
    """Hartley normalisation per sample set; p: (H, n, 2) (homography.py:118-125)."""
    c = p.mean(axis=1, keepdims=True)
    rms = np.sqrt(((p - c) ** 2).sum(axis=2).mean(axis=1))
    with np.errstate(divide="ignore"):
        s = np.sqrt(2) / rms
    T = np.zeros((len(p), 3, 3))
    T[:, 0, 0] = s
    T[:, 1, 1] = s
    T[:, 0, 2] = -s * c[:, 0, 0]
    T[:, 1, 2] = -s * c[:, 0, 1]
    T[:, 2, 2] = 1.0
    return (p - c) * s[:, None, None], T


def dlt_homography_batch(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    """Batched normalised DLT (homography.py:131-142); src, dst: (H, n, 2) -> (H, 3, 3)."""
    sn, Ts = _normalise_batch(src)
    dn, Td = _normalise_batch(dst)
    x, y, u, v = sn[..., 0], sn[..., 1], dn[..., 0], dn[..., 1]
    z, o = np.zeros_like(x), np.ones_like(x)
    r1 = np.stack([-x, -y, -o, z, z, z, u * x, u * y, u], axis=-1)
    r2 = np.stack([z, z, z, -x, -y, -o, v * x, v * y, v], axis=-1)
    A = np.stack([r1, r2], axis=2).reshape(len(src), -1, 9)
    A = np.nan_to_num(A, nan=0.0, posinf=0.0, neginf=0.0)
    Hn = np.linalg.svd(A)[2][:, -1, :].reshape(-1, 3, 3)
    with np.errstate(all="ignore"):
        Hm = np.linalg.solve(Td, Hn @ Ts)
        return Hm / Hm[:, 2:3, 2:3]


def decompose_homography(H, K=np.eye(3)):
    """homography.py:59-78."""
    Kinv = np.linalg.inv(np.asarray(K, dtype=np.float64))
    h1, h2, h3 = H[:, 0], H[:, 1], H[:, 2]
    norm = np.linalg.norm(Kinv @ h1)
    r1, r2, t = Kinv @ h1 / norm, Kinv @ h2 / norm, Kinv @ h3 / norm
    R = np.stack([r1, r2, np.cross(r1, r2)], axis=1)
    U, _, Vt = np.linalg.svd(R)
    return U @ Vt, t
