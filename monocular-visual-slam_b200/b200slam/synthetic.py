"""Synthetic KITTI-shaped workloads (BASELINE.json configs #2-#5).

No dataset ships with the box, so descriptors and keypoints are generated:
frame k+1 keeps `keep` of frame k's points (descriptor copied with per-bit flip noise
`flip`), replaces the rest with fresh random ones, and permutes the rows.  Keypoints are
3-D points projected with the KITTI-00 intrinsics under a forward motion with a small
yaw, plus Gaussian pixel noise; they are handed out K^-1-normalised (the reference's
essential-matrix path is only self-consistent for K = I, SURVEY finding 3).
"""
from __future__ import annotations

import numpy as np

KITTI_K = np.array([[718.856, 0.0, 607.1928], [0.0, 718.856, 185.2157], [0.0, 0.0, 1.0]])


def _project(P, R, t, K):
    c = P @ R.T + t
    uv = c @ K.T
    return uv[:, :2] / uv[:, 2:3]


def tracking_pairs(n_pairs: int, n_desc: int = 2000, seed: int = 1234, keep: float = 0.7,
                   flip: float = 0.08, ragged: bool = False, pixel_noise: float = 0.5):
    """Consecutive-frame pairs -> (q_list, t_list, kpq_list, kpt_list); keypoints are
    normalised image coordinates (float32)."""
    rng = np.random.default_rng(seed)
    Kinv = np.linalg.inv(KITTI_K)
    yaw = 0.02
    R = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
    t = np.array([0.05, 0.0, -1.0])
    qs, ts, kq, kt = [], [], [], []
    for _ in range(n_pairs):
        nq = int(rng.integers(int(0.9 * n_desc), n_desc + 1)) if ragged else n_desc
        nt = int(rng.integers(int(0.9 * n_desc), n_desc + 1)) if ragged else n_desc
        q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
        P = np.stack([rng.uniform(-10, 10, nq), rng.uniform(-2, 2, nq), rng.uniform(5, 40, nq)], axis=1)
        tdesc = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
        Pt = np.stack([rng.uniform(-10, 10, nt), rng.uniform(-2, 2, nt), rng.uniform(5, 40, nt)], axis=1)
        m = min(nq, nt)
        src = rng.permutation(nq)[:m]
        dst = rng.permutation(nt)[:m]
        kept = rng.random(m) < keep
        bits = np.unpackbits(q[src[kept]], axis=1)
        bits ^= (rng.random(bits.shape) < flip).astype(np.uint8)
        tdesc[dst[kept]] = np.packbits(bits, axis=1)
        Pt[dst[kept]] = P[src[kept]]
        p1 = _project(P, np.eye(3), np.zeros(3), KITTI_K) + rng.normal(0, pixel_noise, (nq, 2))
        p2 = _project(Pt, R, t, KITTI_K) + rng.normal(0, pixel_noise, (nt, 2))
        n1 = np.hstack([p1, np.ones((nq, 1))]) @ Kinv.T
        n2 = np.hstack([p2, np.ones((nt, 1))]) @ Kinv.T
        qs.append(q), ts.append(tdesc)
        kq.append(n1[:, :2].astype(np.float32)), kt.append(n2[:, :2].astype(np.float32))
    return qs, ts, kq, kt


def tracking_sequence(n_frames: int, n_desc: int = 2000, seed: int = 1234, keep: float = 0.7,
                      flip: float = 0.08, pixel_noise: float = 0.5):
    """A true frame SEQUENCE (BASELINE configs[1]): frame k+1 re-observes `keep` of frame k's
    3-D points (descriptor copied with bit noise, rows permuted) under the camera motion
    and replaces the rest.  -> (desc uint8 [F, N, 32], kp float32 [F, N, 2] normalised)."""
    rng = np.random.default_rng(seed)
    Kinv = np.linalg.inv(KITTI_K)
    yaw = 0.02
    R = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
    t = np.array([0.05, 0.0, -1.0])

    def fresh(n):
        return (rng.integers(0, 256, (n, 32), dtype=np.uint8),
                np.stack([rng.uniform(-10, 10, n), rng.uniform(-2, 2, n), rng.uniform(5, 40, n)], axis=1))

    def observe(P):
        uv = _project(P, np.eye(3), np.zeros(3), KITTI_K) + rng.normal(0, pixel_noise, (len(P), 2))
        return (np.hstack([uv, np.ones((len(P), 1))]) @ Kinv.T)[:, :2].astype(np.float32)

    desc = np.empty((n_frames, n_desc, 32), np.uint8)
    kp = np.empty((n_frames, n_desc, 2), np.float32)
    d, P = fresh(n_desc)
    for f in range(n_frames):
        desc[f], kp[f] = d, observe(P)
        nd, nP = fresh(n_desc)
        kept = rng.random(n_desc) < keep
        Pm = P @ R.T + t                                   # same points in the next camera frame
        kept &= Pm[:, 2] > 1.0                             # points that passed the camera are dropped
        bits = np.unpackbits(d[kept], axis=1)
        bits ^= (rng.random(bits.shape) < flip).astype(np.uint8)
        slots = rng.permutation(n_desc)[: int(kept.sum())]
        nd[slots], nP[slots] = np.packbits(bits, axis=1), Pm[kept]
        d, P = nd, nP
    return desc, kp
