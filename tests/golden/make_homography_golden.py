#!/usr/bin/env python
"""Generate tests/golden/homography_golden.npz by running the UNMODIFIED reference
(/root/reference/homography.py: dlt_homography, ransac_homography).  Build container only.

Per scene: the correspondences, the 4-samples the seeded generator draws, the reference's H for
every sample (dlt_homography), every hypothesis' inlier mask (obtained from the unmodified
ransac_homography with max_iter=1 and a stub generator whose ``choice`` returns the preset
sample), and the result of the seeded full run (early exit included)."""
import importlib.machinery
import importlib.util
import sys
import warnings
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
loader = importlib.machinery.SourceFileLoader("_ref_homography", str(REF / "homography.py"))
spec = importlib.util.spec_from_loader("_ref_homography", loader)
ref_h = importlib.util.module_from_spec(spec)
sys.modules["_ref_homography"] = ref_h
loader.exec_module(ref_h)


class _Stub:
    def __init__(self, idx):
        self.idx = np.asarray(idx)

    def choice(self, n, k, replace=False):
        return self.idx


def planar_scene(seed, n, outlier_frac, noise, pixel=True):
    rng = np.random.default_rng(seed)
    src = rng.uniform(0, 1200, (n, 2)) if pixel else rng.uniform(-1, 1, (n, 2))
    a = 0.05
    H = np.array([[np.cos(a), -np.sin(a), 12.0], [np.sin(a), np.cos(a), -7.0], [1e-5, -2e-5, 1.0]]) if pixel else \
        np.array([[1.02, 0.03, 0.05], [-0.02, 0.98, -0.03], [0.01, -0.02, 1.0]])
    p = (H @ np.hstack([src, np.ones((n, 1))]).T).T
    dst = p[:, :2] / p[:, 2:] + rng.normal(0, noise, (n, 2))
    out = rng.permutation(n)[: int(outlier_frac * n)]
    dst[out] = rng.uniform(0, 1200, (len(out), 2)) if pixel else rng.uniform(-1, 1, (len(out), 2))
    return src.astype(np.float32), dst.astype(np.float32)


def main():
    out, names = {}, []
    specs = [("planar_px_clean", 1, 200, 0.0, 0.3, True, 3.0, 40), ("planar_px_outliers", 2, 400, 0.5, 0.5, True, 3.0, 64),
             ("planar_norm", 3, 150, 0.3, 0.002, False, 0.02, 48), ("planar_tiny", 4, 9, 0.2, 0.2, True, 3.0, 16),
             ("planar_dup", 5, 60, 0.2, 0.3, True, 3.0, 24)]
    for name, seed, n, of, noise, pixel, th, hyp in specs:
        src, dst = planar_scene(seed, n, of, noise, pixel)
        if name == "planar_dup":
            src[:6] = src[0]                                  # repeated points: degenerate samples occur
            dst[:6] = dst[0]
        rng = np.random.default_rng(100 + seed)
        samples = np.stack([rng.choice(n, 4, replace=False) for _ in range(hyp)])
        Hs = np.full((hyp, 3, 3), np.nan)
        masks = np.zeros((hyp, n), bool)
        valid = np.zeros(hyp, bool)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for h, idx in enumerate(samples):
                try:
                    Hs[h] = ref_h.dlt_homography(src[idx], dst[idx])
                    _, inl = ref_h.ransac_homography(src, dst, th=th, max_iter=1, rng=_Stub(idx))
                    masks[h, inl] = True
                    valid[h] = True
                except (RuntimeError, np.linalg.LinAlgError):
                    pass                                      # < 4 inliers or singular: ransac_homography raises
            Hf, inl = ref_h.ransac_homography(src, dst, th=th, max_iter=hyp, rng=np.random.default_rng(100 + seed))
        names.append(name)
        out.update({f"{name}/src": src, f"{name}/dst": dst, f"{name}/th": np.float64(th), f"{name}/samples": samples,
                    f"{name}/H": Hs, f"{name}/masks": masks, f"{name}/valid": valid, f"{name}/run_H": Hf, f"{name}/run_inliers": inl})
    out["names"] = np.array(names)
    np.savez_compressed(OUT / "homography_golden.npz", **out)
    print("homography_golden.npz:", names, {n: int(out[f"{n}/valid"].sum()) for n in names})


if __name__ == "__main__":
    main()
