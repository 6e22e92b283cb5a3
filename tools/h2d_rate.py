"""Diagnostics (GPU): pinned host -> device and device -> host copy rate, per rank and AGGREGATE when every rank of a
box copies at the same time — the ceiling of bench.py's end-to-end number (80 264 bytes per frame pair go up, 3 584 come
down).  Plain `python tools/h2d_rate.py` = one GPU; under torchrun (`--nproc-per-node N`) all N ranks start together
after a barrier and rank 0 prints the per-rank and the summed rates as one JSON line."""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "monocular-visual-slam_b200")]
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:                                       # same affinity as bench.py
    import pynvml
    pynvml.nvmlInit()
    pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
except Exception:
    pass
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
out = {}
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True)),
                 ("both", lambda: None)):
    if name == "both":                      # upload and download at once on two streams (the e2e pattern)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
        d2 = torch.empty(n, dtype=torch.uint8, device="cuda")

        def fn():
            with torch.cuda.stream(s1):
                d.copy_(h, non_blocking=True)
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    gbs = 10 * n * (2 if name == "both" else 1) / sec / 1e9
    t = torch.tensor([gbs], dtype=torch.float64, device="cuda")
    if world > 1:
        allr = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
        rates = [float(x.item()) for x in allr]
    else:
        rates = [gbs]
    out[name] = {"per_rank_gbs": [round(r, 1) for r in rates], "aggregate_gbs": round(sum(rates), 1)}
if rank == 0:
    out["ranks"] = world
    out["cpus"] = os.cpu_count()
    out["e2e_ceiling_pairs_per_s_aggregate"] = round(out["h2d"]["aggregate_gbs"] * 1e9 / 80264)
    print(json.dumps(out))
if world > 1:
    from b200slam.sharding import shutdown_process_group
    shutdown_process_group()
