"""A/B of diagnostic builds of the scoring kernel (K3h): for every library given on the command line (B2S_LIB is set
per child process) the float64-equality of the counts on the bench batch and the kernel time."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
if os.environ.get("K3H_AB_CHILD"):
    sys.path[:0] = [str(ROOT), str(ROOT / "monocular-visual-slam_b200")]
    import numpy as np
    import torch
    from b200slam.frontend import Frontend, FrontendConfig, sequence_batch
    from b200slam.synthetic import tracking_sequence
    pairs, n = 296, 2000
    desc, kp = tracking_sequence(pairs + 1, n, seed=1234)
    fe = Frontend(FrontendConfig(hypotheses=2000, max_matches=500))
    b = sequence_batch(torch.from_numpy(desc.reshape(-1, 32)).cuda(), torch.from_numpy(kp.reshape(-1, 2)).cuda(), np.full(pairs + 1, n, np.int32), 0, pairs, n)
    keys = fe.matcher.knn2(b)
    sel = fe.matcher.select(b, keys, use_ratio=True, use_cross=True, ratio=0.8, sort_by_distance=True, max_matches=500, with_corr=True, compact=True)
    E = fe.ransac.hypotheses(sel.corr, sel.c_off, sel.count, pairs, 2000, seed=1337)
    bad = 0
    for th2 in (1e-4, 1e-6):
        ref = fe.ransac.score(sel.corr, sel.c_off, sel.count, pairs, E, th2, precision=6464)
        got = fe.ransac.score(sel.corr, sel.c_off, sel.count, pairs, E, th2, precision=64)
        bad += int((ref != got).sum())
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            fe.ransac.score(sel.corr, sel.c_off, sel.count, pairs, E, 1e-4, precision=64)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    print(f"{os.environ.get('B2S_LIB', 'shipped'):50s} mismatches {bad}  K3h {best:.4f} ms", flush=True)
else:
    for lib in ["shipped"] + sys.argv[1:]:
        env = dict(os.environ, K3H_AB_CHILD="1")
        if lib != "shipped":
            env["B2S_LIB"] = str(Path(lib).resolve())
        r = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True, timeout=120)
        print(r.stdout.strip() or r.stderr[-500:], flush=True)
