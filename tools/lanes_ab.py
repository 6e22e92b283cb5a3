"""A/B (GPU): the metric's step (16 launch sets of 296 pairs, one CUDA graph) with the launch sets of a batch
alternating between 1, 2 or 3 streams (ShardedFrontend(lanes=...)); records must be identical."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "monocular-visual-slam_b200")]
import dataclasses

import numpy as np
import torch

import bench
from b200slam.frontend import FrontendConfig, sequence_batch
from b200slam.sharding import ShardedFrontend
from b200slam.synthetic import tracking_sequence

a = bench.parse(["--extras", "none"])
env = bench.Env(a)
P, N, W, S = 296, 2000, 4, 16
desc, kp = tracking_sequence(W * P + 1, N, seed=1234)
d, k = torch.from_numpy(desc.reshape(-1, 32)).cuda(), torch.from_numpy(kp.reshape(-1, 2)).cuda()
counts = np.full(W * P + 1, N, np.int32)
batches = [sequence_batch(d, k, counts, w * P, P, N) for w in range(W)]
bl = [batches[i % W] for i in range(S)]
out, ref = {}, None
for name, cfg in (("full", FrontendConfig(hypotheses=2000, max_matches=500, threshold=0.01, seed=1337)),
                  ("winner_only", FrontendConfig(hypotheses=2000, max_matches=500, threshold=0.01, seed=1337, winner_only=True)),
                  ("with_pose", FrontendConfig(hypotheses=2000, max_matches=500, threshold=0.01, seed=1337, with_pose=True))):
    ref = None
    for lanes in (1, 3, 4):
        sf = ShardedFrontend(cfg, P, sets_per_gather=S, lanes=lanes)
        step, launches, graphed = bench.sharded_step(env, sf, bl)
        ms = env.timed(step, 10, 3)
        torch.cuda.synchronize()
        rec = sf.records().clone()
        same = True if ref is None else bool(torch.equal(rec, ref))
        ref = rec if ref is None else ref
        out[f"{name}/lanes{lanes}"] = {"ms_per_step": round(float(np.mean(ms)), 4), "pairs_per_s": round(P * S / (float(np.mean(ms)) * 1e-3)), "identical_records": same, "launches": launches}
        print(name, lanes, out[f"{name}/lanes{lanes}"], flush=True)
        sf.close()
        del sf
print(json.dumps(out))
