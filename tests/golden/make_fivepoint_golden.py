#!/usr/bin/env python
"""Generate tests/golden/fivepoint_golden.npz: cv2.findEssentialMat on exactly five
correspondences returns every real solution of the minimal problem (stacked 3k x 3) — the
third-party arithmetic behind the reference's 5-point call sites (slam_viewer.py:195,
web_dashboard_server.py:145, visual_slam_offline_entry_point.py:51).  Build container only."""
from pathlib import Path

import cv2
import numpy as np

OUT = Path(__file__).resolve().parent


def main():
    rng = np.random.default_rng(5)
    out, n = {}, 0
    for case in range(60):
        P = np.stack([rng.uniform(-3, 3, 5), rng.uniform(-2, 2, 5), rng.uniform(3, 15, 5)], axis=1)
        yaw, pitch = rng.uniform(-0.3, 0.3, 2)
        Ry = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
        Rx = np.array([[1, 0, 0], [0, np.cos(pitch), -np.sin(pitch)], [0, np.sin(pitch), np.cos(pitch)]])
        t = rng.normal(size=3)
        P2 = P @ (Ry @ Rx).T + t
        if (P2[:, 2] < 0.5).any():
            continue
        src = (P[:, :2] / P[:, 2:] + rng.normal(0, 1e-3 * (case % 2), (5, 2))).astype(np.float32)
        dst = (P2[:, :2] / P2[:, 2:] + rng.normal(0, 1e-3 * (case % 2), (5, 2))).astype(np.float32)
        E, _ = cv2.findEssentialMat(src.astype(np.float64), dst.astype(np.float64), np.eye(3), method=0)
        if E is None:
            continue
        out[f"c{n}/src"], out[f"c{n}/dst"], out[f"c{n}/E"] = src, dst, E.reshape(-1, 3, 3)
        n += 1
    out["n"] = np.int64(n)
    np.savez_compressed(OUT / "fivepoint_golden.npz", **out)
    print("fivepoint_golden.npz:", n, "cases,", sum(len(out[f"c{k}/E"]) for k in range(n)), "solutions")


if __name__ == "__main__":
    main()
