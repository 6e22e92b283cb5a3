"""K3h against K3 (all float64) on the whole bench batch + adversarial near-threshold sets."""
import sys
sys.path[:0] = ["/root/repo", "/root/repo/monocular-visual-slam_b200"]
import numpy as np, torch
from b200slam.frontend import Frontend, FrontendConfig, sequence_batch
from b200slam.synthetic import tracking_sequence
pairs, n = 296, 2000
desc, kp = tracking_sequence(pairs + 1, n, seed=1234)
counts = np.full(pairs + 1, n, np.int32)
fe = Frontend(FrontendConfig(hypotheses=2000, max_matches=500))
b = sequence_batch(torch.from_numpy(desc.reshape(-1, 32)).cuda(), torch.from_numpy(kp.reshape(-1, 2)).cuda(), counts, 0, pairs, n)
keys = fe.matcher.knn2(b)
sel = fe.matcher.select(b, keys, use_ratio=True, use_cross=True, ratio=0.8, sort_by_distance=True, max_matches=500, with_corr=True, compact=True)
bad = 0
for seed in (1337, 7, 99):
    E = fe.ransac.hypotheses(sel.corr, sel.c_off, sel.count, pairs, 2000, seed=seed)
    for th2 in (1e-4, 2.5e-5, 9e-4, 1e-6):
        ref = fe.ransac.score(sel.corr, sel.c_off, sel.count, pairs, E, th2, precision=6464)
        got = fe.ransac.score(sel.corr, sel.c_off, sel.count, pairs, E, th2, precision=64)
        nb = int((ref != got).sum()); bad += nb
        print(seed, th2, "mismatching counts:", nb, "of", ref.numel())
    # scaled hypotheses (scale of E must not matter)
    for sc in (1e-6, 1e-3, 1e3, 1e5):
        ref = fe.ransac.score(sel.corr, sel.c_off, sel.count, pairs, E * sc, 1e-4, precision=6464)
        got = fe.ransac.score(sel.corr, sel.c_off, sel.count, pairs, E * sc, 1e-4, precision=64)
        nb = int((ref != got).sum()); bad += nb
        print(seed, "scale", sc, "mismatching counts:", nb)
print("TOTAL MISMATCHES", bad)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
fe.ransac.score(sel.corr, sel.c_off, sel.count, pairs, E, 1e-4, precision=64); torch.cuda.synchronize()
e0.record()
for _ in range(10): fe.ransac.score(sel.corr, sel.c_off, sel.count, pairs, E, 1e-4, precision=64)
e1.record(); torch.cuda.synchronize()
print("K3h ms", e0.elapsed_time(e1) / 10)
