"""Diagnostics (GPU): K3t on the bench workload — time, float64 re-evaluation rate, equality with K3."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "monocular-visual-slam_b200")]
import numpy as np, torch
from b200slam.frontend import Frontend, FrontendConfig, sequence_batch
from b200slam.synthetic import tracking_sequence
pairs, n = 296, 2000
desc, kp = tracking_sequence(pairs + 1, n, seed=1234)
counts = np.full(pairs + 1, n, np.int32)
fe = Frontend(FrontendConfig(hypotheses=2000, max_matches=500))
b = sequence_batch(torch.from_numpy(desc.reshape(-1, 32)).cuda(), torch.from_numpy(kp.reshape(-1, 2)).cuda(), counts, 0, pairs, n)
keys = fe.matcher.knn2(b)
c = fe.cfg
sel = fe.matcher.select(b, keys, use_ratio=True, use_cross=True, ratio=0.8, sort_by_distance=True, max_matches=500, with_corr=True, compact=True)
E = fe.ransac.hypotheses(sel.corr, sel.c_off, sel.count, pairs, 2000, seed=1337)
ref = fe.ransac.score(sel.corr, sel.c_off, sel.count, pairs, E, 1e-4, precision=6464)
def tm(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("K3 fp64      ms", tm(lambda: fe.ransac.score(sel.corr, sel.c_off, sel.count, pairs, E, 1e-4, precision=6464)))
print("K3h hybrid   ms", tm(lambda: fe.ransac.score(sel.corr, sel.c_off, sel.count, pairs, E, 1e-4, precision=64)))
print("K3t tensor   ms", tm(lambda: fe.ransac.score_tc(sel.corr, sel.c_off, sel.count, pairs, E, 1e-4, max_m=500)))
ct, num, den, band = fe.ransac.score_tc(sel.corr, sel.c_off, sel.count, 8, E[:8].contiguous(), 1e-4, max_m=500, debug=True)
tot = int(sel.count[:8].sum()) * 2000
print("band re-evaluations", int(band[0]), "overflow", int(band[1]), "of", tot, "=", int(band[0]) / tot)
full = fe.ransac.score_tc(sel.corr, sel.c_off, sel.count, pairs, E, 1e-4, max_m=500)
print("counts equal fp64:", bool((full == ref).all()), "mean matches", float(sel.count.float().mean()))
