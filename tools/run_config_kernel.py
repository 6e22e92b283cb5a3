"""ncu target: the Hamming call of a BASELINE config shape, three times.  CONFIG=3 (256 independent 2000^2 pairs),
4 (16 x 10 000^2), 41 (one lone 10 000^2 pair: train-axis split), 5 (4541-keyframe sweep, best-neighbour-only kernel)."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "monocular-visual-slam_b200")]
import numpy as np
import torch

from b200slam.frontend import FrontendConfig, HammingMatcher, MapSweep, PairBatch

cfg = int(os.environ.get("CONFIG", 4))
rng = np.random.default_rng(0)
m = HammingMatcher()
if cfg == 5:
    n_kf, N = 4541, 2000
    kf = [rng.integers(0, 256, (N, 32), dtype=np.uint8) for _ in range(n_kf)]
    kp = [np.zeros((N, 2), np.float32)] * n_kf
    sw = MapSweep(kf, kp, np.arange(n_kf), FrontendConfig(max_matches=500), top=5)
    sw.set_query(torch.from_numpy(kf[7]).cuda(), torch.zeros((N, 2), device="cuda"))
    b = sw._batch(N)
    for _ in range(3):
        sw.matcher.knn2(b, need_second=False)
else:
    P, N = {3: (256, 2000), 4: (16, 10000), 41: (1, 10000)}[cfg]
    qs = [rng.integers(0, 256, (N, 32), dtype=np.uint8) for _ in range(P)]
    ts = [rng.integers(0, 256, (N, 32), dtype=np.uint8) for _ in range(P)]
    b = PairBatch.from_host(qs, ts)
    for _ in range(3):
        m.knn2(b)
torch.cuda.synchronize()
print("done")
