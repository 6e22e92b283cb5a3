#!/usr/bin/env python
"""bench.py — frame-pairs/s of the ORB front-end hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One *step* = one pass of the whole hot path over one batch of synthetic frame pairs:
Hamming kNN-2 (+ratio LUT + cross-check + top-500 select) on 2000x2000 ORB descriptors,
then 2000 8-point essential-matrix hypotheses per pair solved on device, float64 Sampson
scoring against the selected correspondences, winner + inlier mask.
Workload = BASELINE.json configs[1] (KITTI-shaped synthetic sequence, consecutive-pair
tracking, 2000 descriptors / frame).  Weak scaling: every rank owns its own batch; for
N > 1 each step ends with the path's only collective, an NCCL all-gather of per-pair
result records.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
for p in (str(ROOT), str(ROOT / "monocular-visual-slam_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "frame-pairs/sec (2k ORB kNN+ratio+RANSAC E)"
UNIT = "frame-pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=296, help="frame pairs per GPU per step (2 per SM)")
    ap.add_argument("--nfeat", type=int, default=2000)
    ap.add_argument("--hyps", type=int, default=2000)
    ap.add_argument("--cpu-pairs", type=int, default=0, help="CPU baseline sample size (0 = one per core, min 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", default="i8s", choices=["popc", "i8", "i8s"],
                    help="Hamming kernel: K1 POPC, K2 tcgen05 kind::i8 with two products, K2s single product (shipped default)")
    ap.add_argument("--chunks", type=int, default=1, help="e2e: copy/compute overlap chunks")
    ap.add_argument("--scoring", default="cuda", choices=["tc", "cuda"], help="RANSAC scoring: K3t tensor cores or K3h CUDA cores (same counts)")
    ap.add_argument("--e2e-depth", type=int, default=3, help="e2e: steps in flight (device/pinned buffer sets)")
    ap.add_argument("--e2e-mode", default="graphs", choices=["single", "graphs"],
                    help="e2e: 'graphs' = one whole-step graph per tracker and stream (default, 437k pairs/s); "
                         "'single' = SequencePipeline, one kernel stream with overlapping copies (427k, interleave-proof)")
    ap.add_argument("--no-graph", action="store_true", help="e2e: launch eagerly instead of replaying one CUDA graph per step")
    ap.add_argument("--sweep", action="store_true", help="also time every POPC-kernel configuration (extra key)")
    return ap.parse_args()


def workload_config(a, world):
    return {"workload": "BASELINE configs[1]: KITTI-shaped synthetic 1241x376 sequence, consecutive-pair tracking, "
                        f"{a.nfeat} ORB descriptors/frame, kNN-2 + ratio 0.8 + cross-check + top-500, "
                        f"{a.hyps} 8-point E hypotheses/pair, float64 Sampson th=0.01",
            "pairs_per_gpu_per_step": a.pairs, "descriptors_per_frame": a.nfeat, "hypotheses": a.hyps,
            "max_matches": 500, "parallelism": f"pair-sharded x{world}",
            "l2": "flushed between timed steps (256 MiB memset outside the event pairs)"}


# ------------------------------------------------------------------------------------ #
# CPU arm (reference path port; see oracle/reference_path.py)
# ------------------------------------------------------------------------------------ #

def cpu_arm(a, n_pairs, steps=1, warmup=0):
    from b200slam.synthetic import tracking_sequence
    from oracle import reference_path as rp

    cores = os.cpu_count() or 1
    n_pairs = n_pairs or max(8, cores)
    desc, kp = tracking_sequence(n_pairs + 1, a.nfeat, seed=1234)          # the same sequence the GPU arm tracks
    pairs = [(desc[k], desc[k + 1], kp[k], kp[k + 1]) for k in range(n_pairs)]
    for _ in range(warmup):
        rp.run_pairs(pairs[: max(1, min(len(pairs), cores))], workers=cores, max_iter=a.hyps)
    times, res = [], None
    for _ in range(steps):
        res, sec, workers = rp.run_pairs(pairs, workers=cores, max_iter=a.hyps)
        times.append(sec)
    sec = float(np.mean(times))
    # same work as the GPU unit (every hypothesis scored) on a smaller sample
    nfull = max(1, min(len(pairs), cores))
    _, sec_full, _ = rp.run_pairs(pairs[:nfull], workers=cores, max_iter=a.hyps, full_budget=True)
    return {"value": n_pairs / sec, "unit": UNIT, "cores": workers, "kind": "port",
            "sample": f"{n_pairs} pairs/step x {steps} step(s), one process per core, cv2.BFMatcher knnMatch+crossCheck "
                      f"(1 thread each) + the reference's Python RANSAC loop with its early exit (8-point SVD + NumPy Sampson "
                      f"per iteration, max_iter={a.hyps})",
            "value_full_budget": nfull / sec_full,
            "sample_full_budget": f"{nfull} pairs, all {a.hyps} hypotheses scored (no early exit) = the GPU unit's work",
            "mean_matches": float(np.mean([r[0] for r in res])), "sec_per_step": sec}


def reference_main(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_arm(a, a.cpu_pairs, steps=max(1, a.steps), warmup=min(a.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": cb["sec_per_step"] * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8 popcount + f64 Sampson", "data": "synthetic",
            "config": workload_config(a, 1), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ #
# clocks sampler
# ------------------------------------------------------------------------------------ #

class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ #
# B200 arm
# ------------------------------------------------------------------------------------ #

def b200_main(a):
    import torch
    import torch.distributed as dist

    from b200slam import _capi
    from b200slam.frontend import Frontend, FrontendConfig, PairBatch, mma_microbench, pipe_microbench
    from b200slam.synthetic import tracking_sequence

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa = _bind_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ["NCCL_DEBUG"] = os.environ.get("B2S_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    lib = _capi.load_library()

    # ---- synthetic SEQUENCE: pairs+1 frames, consecutive pairs share frames (uploaded once) ----
    from b200slam.frontend import SequenceTracker, sequence_batch
    desc_np, kp_np = tracking_sequence(a.pairs + 1, a.nfeat, seed=1234 + rank)
    counts = np.full(a.pairs + 1, a.nfeat, np.int32)
    cfg = FrontendConfig(hypotheses=a.hyps, max_matches=500, threshold=0.01, precision=64, seed=1337 + rank, scoring=a.scoring)
    VARIANTS = {"popc": _capi.VARIANT_POPC, "i8": _capi.VARIANT_I8MMA, "i8s": _capi.VARIANT_I8MMA1}
    variant = VARIANTS[a.variant]
    fe = Frontend(cfg, variant=variant)
    desc_host = torch.from_numpy(desc_np.reshape(-1, 32)).pin_memory()
    kp_host = torch.from_numpy(kp_np.reshape(-1, 2)).pin_memory()
    desc_dev, kp_dev = desc_host.to(dev), kp_host.to(dev)
    batch = sequence_batch(desc_dev, kp_dev, counts, 0, a.pairs, a.nfeat)     # inputs resident in HBM
    qs = ts = [None] * a.pairs
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    pair_ids = torch.arange(rank * a.pairs, (rank + 1) * a.pairs, dtype=torch.int32, device=dev)

    # One step = the library calls of Frontend.run on the resident batch.  They are captured ONCE in a
    # CUDA graph and replayed per step (same kernels, same work; --no-graph launches them eagerly): with
    # eight ranks on one 32-vCPU host the eager Python path could no longer keep a 0.7 ms step fed.
    l_eager0 = lib.b2s_launch_count()
    res_static = fe.run(batch)                                   # eager once: lazy init, workspace allocation
    torch.cuda.synchronize()
    launches_per_step = int(lib.b2s_launch_count() - l_eager0)
    step_graph = None
    if not a.no_graph:
        step_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(step_graph):
            res_static = fe.run(batch)
        torch.cuda.synchronize()

    def step():
        if step_graph is not None:
            step_graph.replay()
            res = res_static
        else:
            res = fe.run(batch)
        if world > 1:
            rec = torch.stack([res.sel.count, res.best_h, res.best_count, pair_ids], dim=1).contiguous()
            _allgather(rec)
        return res

    gather_buf = torch.empty((world * a.pairs, 4), dtype=torch.int32, device=dev) if world > 1 else None

    def _allgather(rec):
        dist.all_gather_into_tensor(gather_buf, rec)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for s, e in ev:
            flush.fill_(0)                      # evict L2 between timed steps (outside the event pair)
            s.record()
            fn()
            e.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        return [s.elapsed_time(e) for s, e in ev]

    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    ms = timed(step, a.steps, a.warmup)
    warm_launches_per_step = launches_per_step                  # replayed graph nodes = the eager step's launches
    total_ms = float(np.sum(ms))
    if world > 1:
        tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    value = world * a.pairs * a.steps / (total_ms * 1e-3)

    # ---- dominant kernel alone: Hamming kNN-2 (CUDA events on the launching stream) ----
    def k1():
        fe.matcher.knn2(batch)
    k1_ms = timed(k1, a.steps, a.warmup)
    k1_avg = float(np.mean(k1_ms)) * 1e-3
    # the tensor-core kernel proper (CUDA events inside the library, around that one launch)
    import ctypes as C
    kern_ms = []
    if a.variant != "popc":
        lib.b2s_hamming_kernel_timing(1, None)
        for _ in range(a.steps):
            flush.fill_(0)
            fe.matcher.knn2(batch)
            v = C.c_float(0.0)
            lib.b2s_hamming_kernel_timing(-1, C.byref(v))
            kern_ms.append(float(v.value))
        lib.b2s_hamming_kernel_timing(0, None)
    # the other variants, for the K1 / K2 / K2s decision record
    from b200slam.frontend import HammingMatcher
    variants_ms = {a.variant: k1_avg * 1e3}
    for name, vid in VARIANTS.items():
        if name != a.variant:
            other = HammingMatcher(variant=vid)
            variants_ms[name] = float(np.mean(timed(lambda: other.knn2(batch), max(3, a.steps // 2), 2)))
            del other
    popc_ops = 8.0 * float(a.pairs) * a.nfeat * a.nfeat                    # algorithmic POPC32 per launch
    alg_bytes = float(batch.total_nq + batch.total_nt) * 32 + 4.0 * (2 * batch.total_nq + batch.total_nt)

    # ---- per-stage device times (CUDA events around each C-ABI call, same inputs) ----
    from b200slam.frontend import PoseRecovery
    pose_rec = PoseRecovery()

    def stage_times(reps=5):
        c = fe.cfg
        acc = {}
        def tm(name, fn):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            e1.synchronize()
            acc.setdefault(name, []).append(e0.elapsed_time(e1))
            return out
        for _ in range(reps + 1):
            keys = tm("hamming", lambda: fe.matcher.knn2(batch))
            sel = tm("select", lambda: fe.matcher.select(batch, keys, use_ratio=c.use_ratio, use_cross=c.use_cross, ratio=c.ratio,
                                                         sort_by_distance=True, max_matches=c.max_matches, with_corr=True, compact=True))
            E = tm("eight_point", lambda: fe.ransac.hypotheses(sel.corr, sel.c_off, sel.count, batch.n_pairs, c.hypotheses, seed=c.seed))
            cnts = tm("score", lambda: fe.score(sel, batch, E))
            w = tm("winner", lambda: fe.ransac.select(cnts, sel.corr, sel.c_off, sel.count, batch.n_pairs, E, c.threshold ** 2))
            if pose_rec is not None:     # not part of the metric's unit: reported beside it
                tm("pose_recovery_extra", lambda: pose_rec.recover(sel.corr, sel.c_off, sel.count, batch.n_pairs, sel.stride, mask=w[2]))
        return {k: float(np.mean(v[1:])) for k, v in acc.items()}
    stages = stage_times()

    # ---- end to end through host buffers: pinned host frames -> device -> kernels -> pinned host ----
    # EVERY step uploads its frames (23.8 MB) and downloads its results (1.9 MB) inside the timed region;
    # one event pair brackets all K steps (with the steps overlapping there is no per-step time to add up).
    # Default ("graphs"): one whole-step CUDA graph (uploads, kernels, downloads on three streams) per
    # tracker, `depth` trackers on their own streams: the upload of step s+1 overlaps the kernels of step
    # s, and so may the first / last kernels of neighbouring steps (which hides the graph launch gaps).
    # "single": SequencePipeline — copies overlap, but all kernels run on ONE stream and never interleave
    # across steps: 2 % slower here (the 20 us between back-to-back graph launches show), immune to the
    # interleaving lottery (a different memset scheme once cost the graphs mode 8 %).
    depth = max(1, a.e2e_depth)
    from b200slam.frontend import SequencePipeline
    if a.e2e_mode == "single":
        pipe = SequencePipeline(a.pairs + 1, a.nfeat, cfg, variant=variant, depth=depth, device=dev, use_graph=not a.no_graph)
        tracker, e2e_streams, n_chunks = pipe, list(pipe.streams()), 1
        gather_small = torch.empty((world * a.pairs, 4), dtype=torch.int32, device=dev) if world > 1 else None

        def _gather(res):                                          # runs on the kernel stream, after the step's graph
            rec = torch.stack([res.sel.count, res.best_h, res.best_count, pair_ids], dim=1).contiguous()
            dist.all_gather_into_tensor(gather_small, rec)

        def e2e_step(i):
            pipe.submit(desc_host, kp_host, counts, after_compute=_gather if world > 1 else None)
    else:
        trackers = [SequenceTracker(a.pairs + 1, a.nfeat, cfg, variant=variant, chunks=a.chunks, device=dev, use_graph=not a.no_graph)
                    for _ in range(depth)]
        e2e_streams = [torch.cuda.Stream(dev) for _ in range(depth)]
        tracker, n_chunks = trackers[0], len(trackers[0].bounds)
        gather_small = None
        if world > 1:
            lo, hi = tracker.bounds[-1]
            gather_small = [torch.empty((world * (hi - lo), 4), dtype=torch.int32, device=dev) for _ in range(depth)]

        def e2e_step(i):
            tr = trackers[i % depth]
            with torch.cuda.stream(e2e_streams[i % depth]):
                tr.run(desc_host, kp_host, counts)
                if world > 1:
                    last = tr._keep[-1]
                    rec = torch.stack([last.sel.count, last.best_h, last.best_count, pair_ids[: last.best_h.numel()]], dim=1).contiguous()
                    dist.all_gather_into_tensor(gather_small[i % depth], rec)

    def e2e_timed(steps, warmup):
        for i in range(max(warmup, 2 * depth)):                # eager pass + graph capture for every buffer set
            e2e_step(i)
            torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        main = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for st in e2e_streams:
            st.wait_event(e0)
        for i in range(steps):
            e2e_step(i)
        for st in e2e_streams:
            main.wait_stream(st)
        e1.record(main)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    e2e_total = e2e_timed(a.steps, a.warmup)
    if world > 1:
        tt = torch.tensor([e2e_total], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_total = float(tt.item())
    e2e_value = world * a.pairs * a.steps / (e2e_total * 1e-3)
    clk = clocks.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline denominators measured live ----
    peaks = {}
    mp_path = ROOT / "MEASURED_PEAKS.json"
    mp = json.loads(mp_path.read_text()) if mp_path.exists() else {}
    hbm_peak, hbm_src = float(mp.get("hbm_gbs", 6650.0)), ("MEASURED_PEAKS.json" if mp else "fallback (B200_PROFILING.md)")
    for name in ("popc", "lop3", "imnmx", "imad", "dfma", "ffma", "shfl"):
        peaks[name] = pipe_microbench(name)
    i8_peak = mma_microbench()                                  # dense tcgen05 kind::i8, int8 op/s
    sms, _, _, clock_khz = _devinfo(lib)
    popc_peak = peaks["popc"]
    bf16_peak = float(mp.get("bf16_tflops", 1590.0))
    i8_ops = 64.0 * popc_ops                                    # 2*256 int8 ops per descriptor pair = 64 per POPC32
    shipped = "i8s" if a.variant == "popc" else a.variant
    t_stage = variants_ms[shipped] * 1e-3                       # memsets + 2 expand launches + kernel
    t_i8 = (float(np.mean(kern_ms)) * 1e-3) if kern_ms else t_stage
    # DRAM traffic of one launch of the shipped kernel, from the ncu --set full capture in profiles/
    traffic = {"i8s": 185.3e6, "i8": 399.7e6}.get(shipped)
    roof = {"bound": "tensor", "achieved": i8_ops / t_i8 / 1e12, "peak": i8_peak / 1e12, "unit": "TOP/s (int8)",
            "frac": i8_ops / t_i8 / i8_peak, "traffic": traffic,
            "kernel": {"i8s": "hamming_knn2_i8s_kernel", "i8": "hamming_knn2_i8_kernel"}[shipped],
            "kernel_ms": t_i8 * 1e3, "algorithmic_int8_ops_per_launch": i8_ops,
            "stage": {"what": "whole Hamming call (b2s_hamming_knn2_shared for the shipped variant: 3 memsets + ONE expand_pm8_kernel launch over the frames + the kernel; the other variants expand per pair and side)",
                      "ms": t_stage * 1e3, "frac": i8_ops / t_stage / i8_peak},
            "peak_source": "b2s_mma_microbench measured in this run (dense tcgen05.mma kind::i8 M128.N128.K32 from shared memory); "
                           "2 x MEASURED_PEAKS bf16 would be %.0f TOP/s" % (2.0 * bf16_peak),
            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, profiles/r01_final5_ncu.md",
            "note": "algorithmic = ONE 2*256*Nq*Nt int8 contraction per frame pair (SURVEY 8d); the kernel issues 9 K-steps per 8 of "
                    "data (the 9th adds the row/column index), and the two-product variant i8 issues the contraction twice",
            "hamming_variants_ms": variants_ms,
            "popc_view": {"bound": "int-popc-pipe", "achieved": popc_ops / (variants_ms["popc"] * 1e-3) / 1e12, "peak": popc_peak / 1e12,
                          "unit": "TPOPC/s", "frac": popc_ops / (variants_ms["popc"] * 1e-3) / popc_peak,
                          "kernel": "hamming_knn2_popc_kernel", "kernel_ms": variants_ms["popc"],
                          "note": "K1 on the integer pipe: 8 POPC32 per descriptor pair algorithmic; the carry-save tree issues 5, "
                                  "hence > 1.  Kept as the integer-pipe baseline of the K1-vs-K2 decision."},
            "hbm": {"bound": "hbm", "achieved": alg_bytes / t_i8 / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": alg_bytes / t_i8 / 1e9 / hbm_peak, "peak_source": hbm_src,
                    "note": "not the binding roofline: 210 POPC per byte (SURVEY 8d)"},
            "ransac_score": {"bound": "fp32-fma-pipe", "achieved": 22.0 * a.pairs * a.hyps * 500.0 / (stages["score"] * 1e-3) / 1e12,
                             "peak": peaks["ffma"] / 1e12, "unit": "T FFMA-class instr/s",
                             "frac": 22.0 * a.pairs * a.hyps * 500.0 / (stages["score"] * 1e-3) / peaks["ffma"],
                             "kernel": "ransac_score_hybrid_kernel", "kernel_ms": stages["score"],
                             "note": "K3h (shipped): 22 FFMA/FMUL + 2 FSETP + 2 predicated integer ops + 1 LDS.128 per (hypothesis, "
                                     "correspondence) in float32, M = 500; evaluations inside the rounding band (~2e-3) are redone in "
                                     "float64 once per group of 32, so the counts are the float64 kernel's"},
            "pipe_rates_per_clk_per_sm": {k: v / sms / (clock_khz * 1e3) for k, v in peaks.items()}}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8 (+-8 contraction of u8 bit vectors, exact) + f64 Sampson", "data": "synthetic", "config": workload_config(a, world),
            "clocks": clk, "roofline": roof,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": tracker.h2d_bytes, "d2h_bytes_per_step": tracker.d2h_bytes, "chunks": n_chunks, "cuda_graph": not a.no_graph, "steps_in_flight": depth, "pipeline": a.e2e_mode,
                    "timing": "one CUDA-event pair around all K steps (steps overlap); inputs arrive over PCIe every step",
                    "cpu_affinity_first_count": numa,
                    "ms_per_step": e2e_total / a.steps},
            "stage_ms": stages, "gpu_launches": warm_launches_per_step * a.steps, "gpu_launches_per_step": warm_launches_per_step,
            "value_cuda_graph": step_graph is not None,
            "hamming_variant": a.variant,
            "popc_kernel_config": dict(zip(("csa_level", "rows_per_thread", "warps"), _getcfg(lib)))}
    if a.sweep:
        line["popc_kernel_sweep_ms"] = sweep(lib, fe, batch, timed, a)
    if world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_arm(a, a.cpu_pairs)
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _devinfo(lib):
    import ctypes as C
    v = [C.c_int(0) for _ in range(4)]
    lib.b2s_device_info(*[C.byref(x) for x in v])
    return tuple(x.value for x in v)


def _getcfg(lib):
    import ctypes as C
    v = [C.c_int(0) for _ in range(3)]
    lib.b2s_hamming_get_config(*[C.byref(x) for x in v])
    return tuple(x.value for x in v)


def sweep(lib, fe, batch, timed, a):
    out = {}
    keep = _getcfg(lib)
    for rows in (2, 4):
        for warps in (4, 8):
            for csa in (0, 1, 2, 3):
                lib.b2s_hamming_set_config(csa, rows, warps)
                out[f"csa{csa}_r{rows}_w{warps}"] = float(np.mean(timed(lambda: fe.matcher.knn2(batch), 5, 2)))
    lib.b2s_hamming_set_config(*keep)
    return out


def _bind_to_gpu_numa_node(local: int):
    """Pin this rank to the CPUs next to its GPU (NVML's ideal affinity) BEFORE the pinned host
    buffers are allocated, so that with several ranks on one box every rank's uploads come from its
    own NUMA node instead of all of them crossing the socket interconnect.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return sorted(os.sched_getaffinity(0))[:1] + [len(os.sched_getaffinity(0))]
    except Exception:
        return None


def main():
    a = parse()
    if a.impl == "reference":
        reference_main(a)
    else:
        b200_main(a)


if __name__ == "__main__":
    main()
