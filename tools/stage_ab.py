"""A/B of per-stage device times for the metric's launch set with the library given by B2S_LIB (diagnostic builds)."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "monocular-visual-slam_b200")]
import numpy as np
import torch

import bench
from b200slam.frontend import Frontend, FrontendConfig, PoseRecovery, sequence_batch
from b200slam.synthetic import tracking_sequence

a = bench.parse(["--extras", "none"])
env = bench.Env(a)
P, N = 296, 2000
desc, kp = tracking_sequence(P + 1, N, seed=1234)
b = sequence_batch(torch.from_numpy(desc.reshape(-1, 32)).cuda(), torch.from_numpy(kp.reshape(-1, 2)).cuda(), np.full(P + 1, N, np.int32), 0, P, N)
fe = Frontend(FrontendConfig(hypotheses=2000, max_matches=500))
st = bench.stage_times(env, fe, b, PoseRecovery())
print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in st.items() if k != "how"}))
