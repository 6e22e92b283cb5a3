"""Runs the i8 Hamming kernel once on the bench workload (for ncu).  K2S=1 selects the single-product variant."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "monocular-visual-slam_b200")]
import numpy as np, torch
from b200slam import _capi
from b200slam.frontend import HammingMatcher, PairBatch
pairs, n = 296, 2000
rng = np.random.default_rng(0)
qs = [rng.integers(0, 256, (n, 32), dtype=np.uint8) for _ in range(pairs)]
ts = [rng.integers(0, 256, (n, 32), dtype=np.uint8) for _ in range(pairs)]
batch = PairBatch.from_host(qs, ts)
m = HammingMatcher(variant=_capi.VARIANT_I8MMA1 if os.environ.get("K2S") else _capi.VARIANT_I8MMA)
for _ in range(3):
    m.knn2(batch)
torch.cuda.synchronize()
