"""One launch set of the metric's configuration (296 consecutive pairs, 2000 descriptors / frame, 2000 hypotheses), three
times, eagerly — the target of the ncu captures (profiles/).  WINNER=1 uses winner-only scoring, POSE=1 adds K7."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "monocular-visual-slam_b200")]
import numpy as np
import torch

from b200slam.frontend import Frontend, FrontendConfig, record_bytes, sequence_batch
from b200slam.synthetic import tracking_sequence

P, N = int(os.environ.get("PAIRS", 296)), 2000
desc, kp = tracking_sequence(P + 1, N, seed=1234)
b = sequence_batch(torch.from_numpy(desc.reshape(-1, 32)).cuda(), torch.from_numpy(kp.reshape(-1, 2)).cuda(), np.full(P + 1, N, np.int32), 0, P, N)
cfg = FrontendConfig(hypotheses=2000, max_matches=500, winner_only=bool(os.environ.get("WINNER")), with_pose=bool(os.environ.get("POSE")))
fe = Frontend(cfg)
rec = torch.empty((P, record_bytes(500)), dtype=torch.uint8, device="cuda")
for _ in range(3):
    fe.run(b, records=rec)
torch.cuda.synchronize()
print("done", int(fe.run(b, records=rec).best_count.float().mean()))
