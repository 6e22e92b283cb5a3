#!/usr/bin/env python
"""bench.py — frame-pairs/s of the ORB front-end hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 1..5]

One *step* = one pass of the whole hot path over one batch of synthetic frame pairs: Hamming kNN-2 (+ratio LUT +
cross-check + top-N select) on ORB descriptors, then 8-point essential-matrix hypotheses per pair solved on the
device, float64 Sampson scoring against the selected correspondences, winner + inlier mask, the per-pair result
records — and, for N > 1, the path's only collective: ONE NCCL all-gather of those records,
captured in the step's CUDA graph.

--config 2 (default, the metric's configuration = BASELINE.json configs[1]): KITTI-shaped synthetic sequence,
  consecutive-pair tracking, 2000 descriptors / frame; a step tracks `--sub-batches` windows of `--pairs` pairs.
  Weak scaling: every rank owns its own sequence.
--config 3: loop-closure verification batch, 256 independent candidate pairs (40 % true matches), pair-sharded
  over the ranks (strong scaling).
--config 4: high-density matching, 10 000 x 10 000 descriptors per pair, 4096 hypotheses, every mutual match kept.
--config 5: relocalization sweep, one query frame against 4541 map keyframes, keyframes sharded over the ranks
  (strong scaling), geometric verification of the top 5.
--config 1: the CPU reference path on the synthetic stand-in for sharp_curve.mp4 (the clip is not in the tree),
  with the drop-in's single-call latency beside it.
The default run adds the other configs as compact sub-records under "configs" (N = 1 only).
Prints ONE JSON line on rank 0.
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
DISTINCT_WINDOWS = 4      # different data windows a step cycles through


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5], help="BASELINE.json configs[config-1]")
    ap.add_argument("--pairs", type=int, default=0, help="frame pairs per GPU per launch set (default: 296 = 2 per SM for config 2, "
                                                         "256 in all for config 3, 16 for config 4, 4541 keyframes for config 5)")
    ap.add_argument("--sub-batches", type=int, default=0, help="launch sets per step (default 16 for configs 2/3, 4 for config 4, 1 for config 5): "
                                                               "keeps a timed step >= ~10 ms")
    ap.add_argument("--nfeat", type=int, default=0)
    ap.add_argument("--hyps", type=int, default=0)
    ap.add_argument("--cpu-pairs", type=int, default=0, help="CPU baseline sample size (0 = one per core, min 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", default="i8s", choices=["popc", "i8", "i8s"],
                    help="Hamming kernel: K1 POPC, K2 tcgen05 kind::i8 with two products, K2s single product (shipped default)")
    ap.add_argument("--scoring", default="cuda", choices=["tc", "cuda"], help="RANSAC scoring: K3t tensor cores or K3h CUDA cores (same counts)")
    ap.add_argument("--e2e-depth", type=int, default=4, help="e2e: steps in flight (device/pinned buffer sets)")
    ap.add_argument("--e2e-schedule", default="interleaved", choices=["interleaved", "serial"], help="SequencePipeline schedule")
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying CUDA graphs")
    ap.add_argument("--extras", default="auto", choices=["auto", "none", "all"],
                    help="sub-records for the other configs + single-call latency in the default line (auto: N = 1 only)")
    ap.add_argument("--ragged", action="store_true", help="config 2: N ~ U[1800, 2000] descriptors per frame (masking path)")
    ap.add_argument("--sweep", action="store_true", help="also time every POPC-kernel configuration (extra key)")
    a = ap.parse_args(argv)
    d = {2: (296, 16, 2000, 2000), 3: (256, 16, 2000, 2000), 4: (16, 4, 10000, 4096), 5: (4541, 1, 2000, 2000), 1: (29, 1, 2000, 2000)}[a.config]
    a.pairs, a.sub_batches, a.nfeat, a.hyps = a.pairs or d[0], a.sub_batches or d[1], a.nfeat or d[2], a.hyps or d[3]
    return a


WORKLOADS = {
    1: "BASELINE configs[0]: consecutive-frame ORB (2000 kp) matching + RANSAC essential matrix on the CPU reference path; sharp_curve.mp4 is "
       "not in the tree (.MISSING_LARGE_BLOBS), substituted by the 30-frame synthetic clip of the reference's tests/test_visual_slam.py:13-36",
    2: "BASELINE configs[1]: KITTI-shaped synthetic 1241x376 sequence, consecutive-pair tracking",
    3: "BASELINE configs[2]: loop-closure verification batch, 256 BoW candidate pairs x 2000 descriptors (40 % true matches), pair-sharded",
    4: "BASELINE configs[3]: high-density matching, 10k x 10k ORB descriptors per pair, 4096 seeded RANSAC hypotheses, every mutual match kept",
    5: "BASELINE configs[4]: relocalization sweep, one query frame vs 4541 persistent-map keyframes, all-pairs Hamming kNN, top-5 verified",
}


def workload_config(a, world):
    mm = {2: 500, 3: 500, 4: 10000, 5: 500, 1: 500}[a.config]
    sel = "kNN-2 + ratio 0.8 + cross-check" if a.config != 5 else "cross-check (persistent_map.py:266)"
    return {"workload": f"{WORKLOADS[a.config]}, {a.nfeat} ORB descriptors/frame, {sel} + top-{mm}, "
                        f"{a.hyps} 8-point E hypotheses/pair, float64 Sampson th=0.01",
            "baseline_config": a.config, "pairs_per_launch_set": a.pairs, "launch_sets_per_step": a.sub_batches,
            "pairs_per_step_all_gpus": a.pairs * a.sub_batches * (world if a.config in (2, 4) else 1),
            "descriptors_per_frame": a.nfeat, "hypotheses": a.hyps, "max_matches": mm,
            "parallelism": (f"pair-sharded x{world}" if a.config != 5 else f"keyframe-sharded x{world}"),
            "ragged": bool(getattr(a, "ragged", False)),
            "l2": "flushed between timed steps (256 MiB memset outside the event pairs); a step's working set (expanded operand tiles, "
                  "> 170 MB per launch set) exceeds the 126 MB L2"}


# ------------------------------------------------------------------------------------ #
# CPU arm (reference path; see oracle/reference_path.py)
# ------------------------------------------------------------------------------------ #

def _cpu_pairs_for(a, n_pairs, seed=1234):
    """The pairs the GPU arm of this config processes, as host arrays (q, t, kq, kt)."""
    from b200slam.synthetic import tracking_pairs, tracking_sequence

    if a.config in (2, 1):
        desc, kp = tracking_sequence(n_pairs + 1, a.nfeat, seed=seed)          # the same sequence the GPU arm tracks
        return [(desc[k], desc[k + 1], kp[k], kp[k + 1]) for k in range(n_pairs)]
    if a.config == 3:
        qs, ts, kq, kt = tracking_pairs(n_pairs, a.nfeat, seed=seed, keep=0.4)
    elif a.config == 4:
        qs, ts, kq, kt = tracking_pairs(n_pairs, a.nfeat, seed=seed)
    else:                                                                     # config 5: the query against keyframes
        qs, ts, kq, kt = tracking_pairs(n_pairs, a.nfeat, seed=seed, keep=0.3)
    return list(zip(qs, ts, kq, kt))


def cpu_arm(a, n_pairs, steps=1, warmup=0):
    from oracle import reference_path as rp

    cores = os.cpu_count() or 1
    mm = {4: None}.get(a.config, 500)
    if a.config == 4:
        n_pairs = n_pairs or max(2, min(cores, 8))                            # ~2-4 s per 10k x 10k pair
    n_pairs = n_pairs or max(8, cores)
    pairs = _cpu_pairs_for(a, n_pairs)
    kw = dict(max_iter=a.hyps, max_matches=mm)
    if a.config == 5:      # the sweep's unit: cross-check match of the query against ONE keyframe; only the top 5 of 4541 are verified
        kw.update(ratio=None, ransac=False)
    times, res, sec_full = [], None, None
    with rp.PairPool(cores) as pool:                                         # ONE pool, every worker initialised before the clock
        workers = pool.workers
        for _ in range(warmup):
            pool.run(pairs[: max(1, min(len(pairs), cores))], **kw)
        for _ in range(steps):
            res, sec = pool.run(pairs, **kw)
            times.append(sec)
        if a.config in (2, 3):   # same work as the GPU unit (every hypothesis scored) on a smaller sample
            nfull = max(1, min(len(pairs), cores))
            _, sec_full = pool.run(pairs[:nfull], full_budget=True, **kw)
    sec = float(np.median(times))
    out = {"value": n_pairs / sec, "unit": UNIT, "cores": workers, "kind": rp.kind(), "cpu_model": rp.cpu_model(),
           "sample": f"{n_pairs} pairs/step x {steps} step(s), mode (ii) of SURVEY 8d: one process per core, each ONE OpenCV thread and ONE BLAS thread; "
                     f"per pair cv2.BFMatcher knnMatch + ratio and crossCheck match, sort, top-{mm}, then the reference's Python RANSAC loop with its "
                     f"early exit (8-point SVD + NumPy Sampson per iteration, max_iter={a.hyps}); "
                     + ("the UNMODIFIED reference code staged in baseline/_ref (ORBFeaturePipeline.match, homography.ransac_essential)" if rp.kind() == "reference"
                        else "oracle port (baseline/_ref absent)"),
           "mean_matches": float(np.mean([r[0] for r in res])), "sec_per_step": sec}
    if a.config == 5:
        out["sample"] = (f"{n_pairs} (keyframe, query) pairs, one process per core (ONE OpenCV thread each): the relocalizer's "
                         "cv2.BFMatcher(crossCheck=True).match + sort (persistent_map.py:266-270); the RANSAC verification of the top 5 keyframes "
                         "(5 x ~0.3 s per query of 4541 keyframes, < 1 %) is not in this figure; " + ("reference code from baseline/_ref" if rp.kind() == "reference" else "oracle port"))
    out["sec_per_step_all"] = [round(t, 4) for t in times]
    if a.config in (2, 3):
        out["value_full_budget"] = nfull / sec_full
        out["sample_full_budget"] = f"{nfull} pairs, all {a.hyps} hypotheses scored (no early exit) = the GPU unit's work (oracle port, vectorised Sampson)"
        # mode (i): one process, OpenCV + BLAS free to use every core (how slam_api calls the path, frame by frame)
        n1 = min(len(pairs), 6)
        _, sec1, thr = rp.run_pairs_single_process(pairs[:n1], **kw)
        out["value_single_process"] = n1 / sec1
        out["sample_single_process"] = f"mode (i): {n1} pairs sequentially in ONE process, cv2.setNumThreads({thr}), BLAS pool unrestricted"
    return out


def reference_main(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if a.config == 1:
        print(json.dumps(config1_line(a, impl="reference")))
        return
    cb = cpu_arm(a, a.cpu_pairs, steps=max(1, a.steps), warmup=min(a.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": cb["sec_per_step"] * 1e3, "higher_is_better": True,
            "scaling": "weak" if a.config in (2, 4) else "strong", "vs_baseline": None, "dtype": "u8 popcount + f64 Sampson", "data": "synthetic",
            "config": workload_config(a, 1), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ #
# config 1: the CPU reference path on the synthetic clip (+ the drop-in's single-call latency)
# ------------------------------------------------------------------------------------ #

def synthetic_clip(n_frames=30, shape=(1080, 1920)):
    """tests/test_visual_slam.py:13-36 of the reference: a seed-0 noise image translated by (2, 1) px per frame."""
    import cv2

    rng = np.random.default_rng(0)
    base = rng.integers(0, 255, size=shape, dtype=np.uint8)
    frames = []
    for i in range(n_frames):
        M = np.float32([[1, 0, 2 * i], [0, 1, i]])
        frames.append(cv2.warpAffine(base, M, (shape[1], shape[0])))
    return frames


def config1_line(a, impl="b200", with_gpu=False):
    import cv2

    from oracle import reference_path as rp

    K = np.array([[1000.0, 0, 960.0], [0, 1000.0, 540.0], [0, 0, 1.0]])
    Kinv = np.linalg.inv(K)
    frames = synthetic_clip(a.pairs + 1)
    orb = cv2.ORB_create(nfeatures=a.nfeat)
    t0 = time.perf_counter()
    feats = []
    for f in frames:
        cv2.setRNGSeed(1337)
        kps, des = orb.detectAndCompute(f, None)
        pts = np.array([k.pt for k in kps], np.float64)
        nrm = (np.hstack([pts, np.ones((len(pts), 1))]) @ Kinv.T)[:, :2].astype(np.float32)      # K = I on K^-1-normalised points (SURVEY finding 3)
        feats.append((des, nrm))
    t_orb = time.perf_counter() - t0
    pairs = [(feats[i][0], feats[i + 1][0], feats[i][1], feats[i + 1][1]) for i in range(len(frames) - 1)]
    res, sec, thr = rp.run_pairs_single_process(pairs, max_iter=a.hyps, max_matches=500)
    cb = {"value": len(pairs) / sec, "unit": UNIT, "cores": thr, "kind": rp.kind(), "cpu_model": rp.cpu_model(),
          "sample": f"{len(pairs)} consecutive pairs of the synthetic clip (1080x1920 noise image, (2, 1) px/frame), real cv2.ORB_create({a.nfeat}) "
                    f"descriptors ({np.mean([len(f[0]) for f in feats]):.0f} per frame), ONE process as the reference runs it (slam_api.py:248-288), "
                    f"cv2 threads = {thr}; ORB detection itself ({t_orb / len(frames) * 1e3:.0f} ms/frame) is outside the path and the timing",
          "mean_matches": float(np.mean([r[0] for r in res])), "mean_inliers": float(np.mean([r[2] for r in res])), "sec_per_step": sec}
    line = {"impl": impl, "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": 1, "warmup": 0,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8 popcount + f64 Sampson", "data": "synthetic", "config": workload_config(a, 1), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if with_gpu:
        line["b200_drop_in"] = single_call_latency(pairs[: min(len(pairs), 12)])
    return line


def single_call_latency(pairs, reps=3):
    """The way the reference calls the path (slam_api.py:254 one `match`, :283 one pose estimate per frame): wall
    time of ONE 2000 x 2000 drop-in call through integration.* — H2D, kernels (K2s cut along the train axis so that
    a lone pair fills the machine), D2H, the Python DMatch list — next to the reference's own calls on the same arrays."""
    import cv2
    import torch

    from integration.feature_pipeline_bridge import FeaturePipelineConfig, build_feature_pipeline
    from integration.pose_bridge import ransac_essential
    from oracle import reference_path as rp

    out = {}
    for cross in (True, False):
        pipe = build_feature_pipeline(FeaturePipelineConfig(cross_check=cross))
        ref = rp.reference_modules()
        if ref:
            rpipe = ref[1].ORBFeaturePipeline(ref[1].FeaturePipelineConfig(cross_check=cross))
            ref_match = rpipe.match
        else:
            bf = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=cross)

            def ref_match(d1, d2, bf=bf, cross=cross):
                if cross:
                    ms = list(bf.match(d1, d2))
                else:
                    ms = [p[0] for p in bf.knnMatch(d1, d2, k=2) if len(p) == 2 and p[0].distance < 0.8 * p[1].distance]
                ms.sort(key=lambda m: m.distance)
                return ms[:500]
        pipe.match(pairs[0][0], pairs[0][1])                    # lazy init
        torch.cuda.synchronize()
        tg, tc, same = [], [], True
        for _ in range(reps):
            for q, t, _, _ in pairs:
                t0 = time.perf_counter()
                got = pipe.match(q, t)
                tg.append(time.perf_counter() - t0)
                t0 = time.perf_counter()
                want = ref_match(q, t)
                tc.append(time.perf_counter() - t0)
                same &= [(m.queryIdx, m.trainIdx, m.distance) for m in got] == [(m.queryIdx, m.trainIdx, m.distance) for m in want]
        key = "match_cross_check" if cross else "match_knn_ratio"
        out[key] = {"b200_ms_median": float(np.median(tg) * 1e3), "cpu_ms_median": float(np.median(tc) * 1e3),
                    "identical_to_cpu": bool(same), "cv2_threads": cv2.getNumThreads()}
    # pose: ransac_essential on the matched points (seeded reference on the CPU side; all 2000 hypotheses on the device)
    pipe = build_feature_pipeline(FeaturePipelineConfig(cross_check=True))
    tg, tc = [], []
    ref = rp.reference_modules()
    for q, t, kq, kt in pairs[:6]:
        ms = pipe.match(q, t)
        src = np.float32([kq[m.queryIdx] for m in ms])
        dst = np.float32([kt[m.trainIdx] for m in ms])
        if len(src) < 8:
            continue
        t0 = time.perf_counter()
        try:
            ransac_essential(src, dst, np.eye(3), th=0.01)
        except RuntimeError:
            pass
        tg.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        try:
            if ref:
                ref[0].ransac_essential(src, dst, np.eye(3), 0.01, 2000, np.random.default_rng(0))
            else:
                from oracle import ransac_oracle as ro
                ro.ransac_essential(src, dst, np.eye(3), 0.01, 2000, np.random.default_rng(0))
        except RuntimeError:
            pass
        tc.append(time.perf_counter() - t0)
    if tg:
        out["ransac_essential"] = {"b200_ms_median": float(np.median(tg) * 1e3), "cpu_ms_median": float(np.median(tc) * 1e3),
                                   "note": "device: all 2000 hypotheses scored + host refit; CPU: the reference's loop with its early exit"}
    out["note"] = ("wall clock of single calls through the reference-shaped API (host NumPy in, Python objects out); "
                   "the batched array API is what the throughput numbers use")
    return out


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
        ok = [r for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        sm = [float(r[1]) for r in ok]
        mx = [float(r[2]) for r in ok if r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in ok if r[3].replace(".", "").isdigit()]
        reasons = set()
        for r in ok:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None}


# ------------------------------------------------------------------------------------ #
# B200 arm
# ------------------------------------------------------------------------------------ #

class Env:
    """Process-wide state of the B200 arm: device, process group, library, timing helpers."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist

        from b200slam import _capi

        self.a, self.torch, self.dist = a, torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.numa = _bind_to_gpu_numa_node(self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))   # NCCL_DEBUG etc. stay as the launcher set them
        self.dev = torch.device("cuda", self.local)
        self.lib = _capi.load_library()
        self.capi = _capi
        self.variants = {"popc": _capi.VARIANT_POPC, "i8": _capi.VARIANT_I8MMA, "i8s": _capi.VARIANT_I8MMA1}
        self.variant = self.variants[a.variant]
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, warmup):
        """Per-step CUDA-event times (ms), L2 flushed between steps, barrier + synchronize on both sides."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for s, e in ev:
            self.flush.fill_(0)                      # evict L2 between timed steps (outside the event pair)
            s.record()
            fn()
            e.record()
        self.barrier()
        return [s.elapsed_time(e) for s, e in ev]

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        tt = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(tt, op=self.dist.ReduceOp.MAX)
        return float(tt.item())

    def finish(self):
        if self.world > 1:      # every run_config* has returned: its graphs (which captured the collective) are gone
            from b200slam.sharding import shutdown_process_group
            shutdown_process_group()


def sharded_step(env, sf, batch_list, use_graph=True):
    """One step = the launch sets of `batch_list` through the library's ShardedFrontend: one eager pass (lazy init;
    counts the library's launches), then the whole batch — kernels, record kernels and the collective — as ONE CUDA
    graph.  -> (step callable, launches per step, graphed)."""
    torch = env.torch
    l0 = env.lib.b2s_launch_count()
    for b in batch_list:
        sf.step(b)
    sf.flush()
    torch.cuda.synchronize()
    launches = int(env.lib.b2s_launch_count() - l0)
    if not use_graph:
        def eager():
            for b in batch_list:
                sf.step(b)
        return eager, launches, False
    sf.capture(batch_list)
    return sf.replay, launches, True


def _capture(env, fn, use_graph=True):
    """Run fn once eagerly (lazy init; counts the library's launches), then capture it into ONE CUDA graph.
    -> (replay callable, launches per call)."""
    torch = env.torch
    l0 = env.lib.b2s_launch_count()
    fn()
    torch.cuda.synchronize()
    launches = int(env.lib.b2s_launch_count() - l0)
    if not use_graph:
        return fn, launches, False
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            fn()
    except Exception as exc:          # a collective the NCCL build cannot capture: kernels stay eager
        torch.cuda.synchronize()
        sys.stderr.write(f"[bench] CUDA-graph capture failed ({type(exc).__name__}: {exc}); running eagerly\n")
        return fn, launches, False
    torch.cuda.synchronize()
    return g.replay, launches, True


def hamming_roofline(env, matcher, batch, n_desc_pairs, steps, peaks=None, need_second=True):
    """Kernel-only time of the tensor-core Hamming kernel on `batch` (CUDA events inside the library around that
    one launch; the GPU is kept busy by the call's own pre-pass so no host gap is inside) and its roofline view."""
    import ctypes as C

    torch, lib = env.torch, env.lib
    for _ in range(2):
        matcher.knn2(batch, need_second=need_second)
    torch.cuda.synchronize()
    lib.b2s_hamming_kernel_timing(1, None)
    kern, call = [], []
    for _ in range(max(3, steps)):
        env.flush.fill_(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        matcher.knn2(batch, need_second=need_second)
        e1.record()
        torch.cuda.synchronize()
        v = C.c_float(0.0)
        lib.b2s_hamming_kernel_timing(-1, C.byref(v))
        kern.append(float(v.value))
        call.append(e0.elapsed_time(e1))
    lib.b2s_hamming_kernel_timing(0, None)
    plan = [C.c_int(0) for _ in range(3)]
    lib.b2s_hamming_last_plan(*[C.byref(x) for x in plan])
    t_k, t_call = float(np.mean(kern)) * 1e-3, float(np.mean(call)) * 1e-3
    ops = 2.0 * 256.0 * n_desc_pairs
    out = {"kernel_ms": t_k * 1e3, "call_ms": t_call * 1e3, "achieved_top_s": ops / t_k / 1e12, "int8_ops": ops, "second_neighbour": bool(need_second),
           "plan": {"query_subtiles_per_item": plan[0].value, "train_split": plan[1].value, "ctas": plan[2].value}}
    if peaks:
        out["frac_of_int8_peak"] = ops / t_k / peaks["i8"]
        out["frac_of_2x_measured_bf16"] = ops / t_k / (2.0 * peaks["bf16"] * 1e12)
    return out


def stage_times(env, fe, b, pose_rec=None, reps=10):
    """Device time of every stage of one launch set on batch `b`: each stage is captured into its OWN CUDA graph
    (inputs = the previous stage's outputs, kept alive) and replayed `reps` times back to back between two events,
    so neither host launch gaps nor allocator calls sit inside the figure (they dominate an eager event pair around
    a 10-40 us kernel)."""
    torch, c, out = env.torch, fe.cfg, {}

    def tm(name, fn):
        res = fn()                                             # eager once (lazy init, also the value the next stage consumes)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                keep = fn()
        g.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1) / reps)
        out[name] = best
        del keep
        return res
    keys = tm("hamming", lambda: fe.matcher.knn2(b))
    sel = tm("select", lambda: fe.matcher.select(b, keys, use_ratio=c.use_ratio, use_cross=c.use_cross, ratio=c.ratio, sort_by_distance=True,
                                                 max_matches=c.max_matches, with_corr=True, compact=True))
    E = tm("eight_point", lambda: fe.ransac.hypotheses(sel.corr, sel.c_off, sel.count, b.n_pairs, c.hypotheses, seed=c.seed))
    cnts = tm("score", lambda: fe.score(sel, b, E))
    w = tm("winner", lambda: fe.ransac.select(cnts, sel.corr, sel.c_off, sel.count, b.n_pairs, E, c.threshold ** 2))
    if pose_rec is not None:
        tm("pose_recovery_extra", lambda: pose_rec.recover(sel.corr, sel.c_off, sel.count, b.n_pairs, sel.stride or b.max_nq, mask=w[2]))
    out["mean_matches"] = float(sel.count.float().mean())
    out["how"] = f"per stage: one CUDA graph of {reps} back-to-back calls, best of 3 replays / {reps}"
    return out


def hamming_variants(env, batch, steps=5):
    """Whole-call time (ms) of the three Hamming kernels on `batch` — the K1-vs-K2 decision record at this shape."""
    from b200slam.frontend import HammingMatcher
    out = {}
    for name, vid in env.variants.items():
        m = HammingMatcher(variant=vid)
        out[name] = float(np.mean(env.timed(lambda: m.knn2(batch), steps, 2)))
        del m
    return out


def measured_peaks(env):
    from b200slam.frontend import mma_microbench, pipe_microbench

    mp_path = ROOT / "MEASURED_PEAKS.json"
    mp = json.loads(mp_path.read_text()) if mp_path.exists() else {}
    peaks = {"i8": mma_microbench(), "bf16": float(mp.get("bf16_tflops", 1590.0)), "hbm": float(mp.get("hbm_gbs", 6650.0)),
             "src": "MEASURED_PEAKS.json" if mp else "fallback (B200_PROFILING.md)"}
    for name in ("popc", "lop3", "imnmx", "imad", "dfma", "ffma", "shfl", "ffma2"):
        peaks[name] = pipe_microbench(name)
    return peaks


def ncu_traffic(kernel_substr="hamming_knn2_i8s_kernel"):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the kernel, read from the committed ncu raw-page CSV
    (profiles/*_ncu_raw.csv, newest round first).  None when no capture is committed."""
    import csv

    for f in sorted((ROOT / "profiles").glob("r*_ncu_raw.csv"), reverse=True):
        try:
            rows = list(csv.reader(f.read_text().splitlines()))
            hdr = rows[0]
            ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            units = rows[1]
            scale = lambda u: {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
            for r in rows[2:]:
                if kernel_substr in r[ki]:
                    val = float(r[ri].replace(",", "")) * scale(units[ri]) + float(r[wi].replace(",", "")) * scale(units[wi])
                    return val, f"profiles/{f.name}"
        except (ValueError, IndexError, OSError):
            continue
    return None, None


def b200_main(a):
    # CPU baseline FIRST (N = 1): before CUDA / NCCL exist in this process, so the worker pool inherits nothing
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cpu = None
    if world == 1 and not a.no_cpu_baseline and a.config != 1:
        try:
            cpu = cpu_arm(a, a.cpu_pairs)
        except Exception as exc:                                   # the baseline must never take the measurement down
            cpu = {"error": f"{type(exc).__name__}: {exc}"}
    env = Env(a)
    line = {1: run_config1, 2: run_config2, 3: run_config3, 4: run_config4, 5: run_config5}[a.config](env, a)
    if env.rank == 0:
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    env.finish()


def _base_line(env, a, value, total_ms, scaling, clk, roof, e2e, stages, launches, extra=None):
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": env.world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "int8 (+-8 contraction of u8 bit vectors, exact) + f64 Sampson", "data": "synthetic", "config": workload_config(a, env.world),
            "clocks": clk, "roofline": roof, "e2e": e2e, "stage_ms": stages,
            "gpu_launches": launches * a.steps, "gpu_launches_per_step": launches, "hamming_variant": a.variant}
    if extra:
        line.update(extra)
    return line


def _roof_from(hr, peaks, kernel, traffic=None, traffic_src=None, extra=None):
    roof = {"bound": "tensor", "achieved": hr["achieved_top_s"], "peak": peaks["i8"] / 1e12, "unit": "TOP/s (int8)",
            "frac": hr["frac_of_int8_peak"], "traffic": traffic, "kernel": kernel, "kernel_ms": hr["kernel_ms"],
            "algorithmic_int8_ops_per_launch": hr["int8_ops"], "frac_of_2x_measured_bf16": hr["frac_of_2x_measured_bf16"],
            "plan": hr["plan"], "second_neighbour_computed": hr["second_neighbour"],
            "stage": {"what": "whole Hamming call: 3 memsets + operand expansion + the kernel (+ the merge of a train-axis split)",
                      "ms": hr["call_ms"], "frac": hr["int8_ops"] / (hr["call_ms"] * 1e-3) / peaks["i8"]},
            "peak_source": "b2s_mma_microbench measured in this run (dense tcgen05.mma kind::i8 M128.N128.K32 from shared memory); MEASURED_PEAKS.json "
                           "has no int8 entry: frac_of_2x_measured_bf16 quotes the same kernel against 2 x its bf16_tflops "
                           f"({2.0 * peaks['bf16']:.0f} TOP/s, {peaks['src']})",
            "traffic_source": traffic_src or "no ncu --set full capture committed for this shape",
            "note": "algorithmic = ONE 2*256*Nq*Nt int8 contraction per frame pair (SURVEY 8d); the kernel issues 9 K-steps per 8 of data "
                    "(the 9th adds the row/column index that makes the accumulator a sort key)"}
    if extra:
        roof.update(extra)
    return roof


# ---- config 2: consecutive-pair tracking (the metric's configuration) ------------------------------------

def run_config2(env, a):
    torch, dist, lib = env.torch, env.dist, env.lib
    from b200slam.frontend import (Frontend, FrontendConfig, HammingMatcher, PoseRecovery, SequencePipeline, sequence_batch)
    from b200slam.sharding import ShardedFrontend
    from b200slam.synthetic import tracking_sequence

    rank, world, dev = env.rank, env.world, env.dev
    P, W = a.pairs, min(DISTINCT_WINDOWS, a.sub_batches)
    F = W * P + 1
    desc_np, kp_np = tracking_sequence(F, a.nfeat, seed=1234 + rank)
    counts = np.full(F, a.nfeat, np.int32)
    if a.ragged:
        counts = np.random.default_rng(77 + rank).integers(int(0.9 * a.nfeat), a.nfeat + 1, F).astype(np.int32)
    cfg = FrontendConfig(hypotheses=a.hyps, max_matches=500, threshold=0.01, precision=64, seed=1337, scoring=a.scoring)
    desc_host = torch.from_numpy(desc_np.reshape(-1, 32)).pin_memory()
    kp_host = torch.from_numpy(kp_np.reshape(-1, 2)).pin_memory()
    desc_dev, kp_dev = desc_host.to(dev), kp_host.to(dev)                       # inputs resident in HBM
    batches = [sequence_batch(desc_dev, kp_dev, counts, w * P, P, a.nfeat) for w in range(W)]
    torch.cuda.synchronize()

    # One step = `sub_batches` launch sets (windows of P consecutive pairs, W distinct windows cycled) through the
    # library's pair-sharded entry: kernels + record kernel + (N > 1) the in-place all-gather, ONE CUDA graph.
    # weak scaling: P pairs per rank per launch set; the records of a step's launch sets leave in ONE all-gather
    sf = ShardedFrontend(cfg, world * P, variant=env.variant, sets_per_gather=a.sub_batches)

    batch_list = [batches[i % W] for i in range(a.sub_batches)]
    # steady state: the graph of a step holds the all-gather of the PREVIOUS step's records (it runs under this step's
    # first RANSAC kernels), so every timed step contains exactly one collective
    step, launches, graphed = sharded_step(env, sf, batch_list, use_graph=not a.no_graph)
    clocks = Clocks(env.local)
    if rank == 0:
        clocks.start()
    ms = env.timed(step, a.steps, a.warmup)
    total_ms = env.max_over_ranks(float(np.sum(ms)))
    pairs_per_step = world * P * a.sub_batches
    value = pairs_per_step * a.steps / (total_ms * 1e-3)

    # ---- sustained behaviour: the step replayed back to back for >= 2 s (clocks / power sampled over it) ----
    import dataclasses
    soak = None
    if world == 1 or rank == 0:
        n_soak = max(10, int(2.2e3 / max(total_ms / a.steps, 1e-3)))
    if world > 1:
        t_n = torch.tensor([n_soak if rank == 0 else 0], dtype=torch.int64, device=dev)
        dist.broadcast(t_n, src=0)
        n_soak = int(t_n.item())
    soak_clk = Clocks(env.local)
    env.barrier()
    if rank == 0:
        soak_clk.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_soak):
        step()
    e1.record()
    env.barrier()
    soak_ms = env.max_over_ranks(e0.elapsed_time(e1))
    if rank == 0:
        sc = soak_clk.stop()
        soak = {"seconds": soak_ms * 1e-3, "steps": n_soak, "value": pairs_per_step * n_soak / (soak_ms * 1e-3), "unit": UNIT,
                "sm_mhz": sc["sm_mhz"], "power_w_max": sc["power_w_max"], "reasons": sc["reasons"], "clock_samples": sc["samples"],
                "note": "the timed step replayed back to back without L2 flushes; sustained clocks and power next to the 20-step figure"}

    # ---- the same step with the reference's DEFAULT matcher (cross_check=True, no ratio test: feature_pipeline.py.bak:17,81-82,
    #      configs/pipeline/kitti_default.json): the Hamming kernel then skips the second neighbour ----
    sfc = ShardedFrontend(dataclasses.replace(cfg, use_ratio=False), world * P, variant=env.variant, sets_per_gather=a.sub_batches)
    step_c, _, _ = sharded_step(env, sfc, batch_list, use_graph=not a.no_graph)
    ms_c = env.timed(step_c, max(3, a.steps // 2), 2)
    sfc.flush()
    value_cc = pairs_per_step * len(ms_c) / (env.max_over_ranks(float(np.sum(ms_c))) * 1e-3)
    del sfc

    # ---- the same step, additionally refitting E on the inliers and recovering (R, t) (K7) into the records ----
    sfp = ShardedFrontend(dataclasses.replace(cfg, with_pose=True), world * P, variant=env.variant, sets_per_gather=a.sub_batches)

    step_p, _, _ = sharded_step(env, sfp, batch_list, use_graph=not a.no_graph)
    ms_p = env.timed(step_p, max(3, a.steps // 2), 2)
    sfp.flush()
    value_pose = pairs_per_step * len(ms_p) / (env.max_over_ranks(float(np.sum(ms_p))) * 1e-3)
    del sfp
    # ---- the same step with winner-only scoring: identical winner / inlier mask, hypotheses that cannot win abandoned early ----
    sfw = ShardedFrontend(dataclasses.replace(cfg, winner_only=True), world * P, variant=env.variant, sets_per_gather=a.sub_batches)

    step_w, launches_w, _ = sharded_step(env, sfw, batch_list, use_graph=not a.no_graph)
    ms_w = env.timed(step_w, max(3, a.steps // 2), 2)
    sfw.flush()
    sf.flush()
    value_winner = pairs_per_step * len(ms_w) / (env.max_over_ranks(float(np.sum(ms_w))) * 1e-3)
    same_winner = bool(torch.equal(sfw.res.best_h, sf.res.best_h) and torch.equal(sfw.res.best_count, sf.res.best_count)
                       and torch.equal(sfw.res.inlier_mask, sf.res.inlier_mask))
    del sfw

    fe = sf.fe
    # ---- dominant kernel alone + the decision record of the three Hamming kernels ----
    peaks = measured_peaks(env) if rank == 0 else None
    hr = hamming_roofline(env, fe.matcher, batches[0], float(P) * a.nfeat * a.nfeat, a.steps, peaks) if a.variant != "popc" else None
    variants_ms = {}
    for name, vid in env.variants.items():
        m = fe.matcher if vid == env.variant else HammingMatcher(variant=vid)
        variants_ms[name] = float(np.mean(env.timed(lambda: m.knn2(batches[0]), max(3, a.steps // 2), 2)))

    # ---- per-stage device times ----
    stages = stage_times(env, fe, batches[0], PoseRecovery())

    # ---- end to end through the library's host-buffer front door: SequencePipeline ----
    # EVERY launch set uploads its 297 frames (23.8 MB) from pinned host memory and downloads its records (1.06 MB, ONE
    # copy) inside the timed region; `depth` steps in flight; one event pair brackets all K steps.
    depth = max(1, a.e2e_depth)
    gathered = None
    hook = None
    pipe = SequencePipeline(P + 1, a.nfeat, cfg, variant=env.variant, depth=depth, device=dev, use_graph=not a.no_graph,
                            schedule=a.e2e_schedule)
    if world > 1:      # the path's collective inside the step: every slot's records are all-gathered in place
        gathered = [torch.empty((world, P, pipe.rec_bytes), dtype=torch.uint8, device=dev) for _ in range(depth)]
        for sl, g in zip(pipe.slots, gathered):
            sl["rec_dev"] = g[rank]
            sl["gathered"] = g
        pipe.after_compute = lambda sl: dist.all_gather_into_tensor(sl["gathered"].view(-1), sl["rec_dev"].reshape(-1))
    win_counts = [np.ascontiguousarray(counts[w * P:w * P + P + 1]) for w in range(W)]
    if a.ragged:
        win_counts = [win_counts[0]] * W        # one frame-size vector (changing it rebuilds the slot graphs); data still differs per window
    host_win = [(desc_host[w * P * a.nfeat:(w * P + P + 1) * a.nfeat], kp_host[w * P * a.nfeat:(w * P + P + 1) * a.nfeat]) for w in range(W)]

    def e2e_step():
        for i in range(a.sub_batches):
            d, k = host_win[i % W]
            pipe.submit(d, k, win_counts[i % W])

    def e2e_timed(steps, warmup):
        for _ in range(max(warmup, 2)):
            e2e_step()
            torch.cuda.synchronize()
        env.barrier()
        main = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for st in pipe.streams():
            st.wait_event(e0)
        for _ in range(steps):
            e2e_step()
        for st in pipe.streams():
            main.wait_stream(st)
        e1.record(main)
        env.barrier()
        return e0.elapsed_time(e1)

    e2e_total = env.max_over_ranks(e2e_timed(a.steps, a.warmup))
    e2e_value = pairs_per_step * a.steps / (e2e_total * 1e-3)
    out0 = pipe.result(0)
    e2e_check = {"pairs_with_matches": int((out0["count"] > 0).sum()), "mean_inliers": float(out0["best_count"].float().mean())}
    clk = clocks.stop() if rank == 0 else None
    if rank != 0:
        return None

    traffic, traffic_src = ncu_traffic("hamming_knn2_i8s_kernel" if a.variant == "i8s" else "hamming_knn2_i8_kernel")
    popc_ops = 8.0 * float(P) * a.nfeat * a.nfeat
    alg_bytes = float(batches[0].total_nq + batches[0].total_nt) * 32 + 4.0 * (2 * batches[0].total_nq + batches[0].total_nt)
    sms, _, _, clock_khz = _devinfo(lib)
    if hr is None:
        t = variants_ms["popc"] * 1e-3
        roof = {"bound": "int-popc-pipe", "achieved": popc_ops / t / 1e12, "peak": peaks["popc"] / 1e12, "unit": "TPOPC/s",
                "frac": popc_ops / t / peaks["popc"], "traffic": None, "kernel": "hamming_knn2_popc_kernel", "kernel_ms": variants_ms["popc"]}
    else:
        kname = {"i8s": "hamming_knn2_i8s_kernel", "i8": "hamming_knn2_i8_kernel"}[a.variant]
        t_i8 = hr["kernel_ms"] * 1e-3
        roof = _roof_from(hr, peaks, kname, traffic, traffic_src, extra={
            "hamming_variants_ms": variants_ms,
            "popc_view": {"bound": "int-popc-pipe", "achieved": popc_ops / (variants_ms["popc"] * 1e-3) / 1e12, "peak": peaks["popc"] / 1e12,
                          "unit": "TPOPC/s", "frac": popc_ops / (variants_ms["popc"] * 1e-3) / peaks["popc"], "kernel": "hamming_knn2_popc_kernel",
                          "kernel_ms": variants_ms["popc"],
                          "note": "K1 on the integer pipe: 8 POPC32 per descriptor pair algorithmic; the carry-save tree issues 5, hence > 1.  "
                                  "Kept as the integer-pipe baseline of the K1-vs-K2 decision."},
            "hbm": {"bound": "hbm", "achieved": alg_bytes / t_i8 / 1e9, "peak": peaks["hbm"], "unit": "GB/s", "frac": alg_bytes / t_i8 / 1e9 / peaks["hbm"],
                    "peak_source": peaks["src"], "note": "not the binding roofline: 210 POPC per byte (SURVEY 8d)"},
            "ransac_score": {"bound": "fp32-fma-pipe (packed FFMA2: register-file reads of three-operand instructions)",
                             "achieved": 19.0 * P * a.hyps * 500.0 / (stages["score"] * 1e-3) / 1e12,
                             "peak": 2.0 * peaks["ffma2"] / 1e12, "unit": "T float32 FMA/s",
                             "frac": 19.0 * P * a.hyps * 500.0 / (stages["score"] * 1e-3) / (2.0 * peaks["ffma2"]), "kernel": "ransac_score_hybrid_kernel",
                             "kernel_ms": stages["score"], "peak_scalar_ffma": peaks["ffma"] / 1e12,
                             "note": "K3h: 19 float32 FMA-class operations per (hypothesis, correspondence), issued as fma.rn.f32x2 on two "
                                     "correspondences at once (9.5 FFMA2 + 2 FSETP + 2 predicated integer ops + 0.5 LDS.128 per evaluation), M = 500; "
                                     "peak = 2 x the measured FFMA2 issue rate with two distinct register-pair sources (an FFMA2 with three distinct "
                                     "pairs issues at 2/3 of it: tools/pipe_rates.py); evaluations inside the rounding band (~2e-3) are redone in "
                                     "float64, so the counts are the float64 kernel's"},
            "pipe_rates_per_clk_per_sm": {k: peaks[k] / sms / (clock_khz * 1e3) for k in ("popc", "lop3", "imnmx", "imad", "dfma", "ffma", "ffma2", "shfl")}})
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes * a.sub_batches, "d2h_bytes_per_step": pipe.d2h_bytes * a.sub_batches,
           "api": "b200slam.frontend.SequencePipeline.submit / result (one library call per launch set)", "schedule": a.e2e_schedule,
           "cuda_graph": pipe.use_graph, "steps_in_flight": depth, "d2h_copies_per_launch_set": 1,
           "timing": "one CUDA-event pair around all K steps (steps overlap); inputs arrive over PCIe from pinned host memory every launch set",
           "cpu_affinity_first_count": env.numa, "ms_per_step": e2e_total / a.steps, "result_check": e2e_check}
    extra = {"value_cuda_graph": graphed, "collective_in_graph": bool(sf.world > 1 and graphed), "soak": soak,
             "value_winner_only": value_winner,
             "value_winner_only_note": "same step through b2s_ransac_winner_batched: the winner, its inlier count and mask are those of the full "
                                       "evaluation (checked on the last launch set: %s), but only the hypotheses that can still exceed 0.8 M or reach the "
                                       "largest complete count are scored to the end; reported beside, not instead of, `value` (all %d hypotheses scored)"
                                       % (same_winner, a.hyps),
             "winner_only_identical": same_winner, "gpu_launches_per_step_winner_only": launches_w,
             "value_cross_check_matcher": value_cc,
             "value_cross_check_matcher_note": "same step with the reference's default matcher (cross_check=True, no ratio test): the tensor-core kernel "
                                               "skips the per-row second neighbour (B2S_HAMMING_BEST_ONLY), which BFMatcher(crossCheck=True).match never reads",
             "value_with_pose": value_pose,
             "value_with_pose_note": "same step + n-point refit of E on the winner's inliers + decomposition / cheirality vote (K7), R | t in the records; "
                                     "outside the metric's unit (SURVEY 8d), reported beside it",
             "stream_lanes": sf.lanes,
             "stream_lanes_note": "the launch sets of a step alternate between this many streams inside the step's CUDA graph (own workspaces "
                                  "per lane): tails, one-CTA-per-pair kernels and launch gaps of one set are filled by the next",
             "record_bytes_per_pair": sf.rec_bytes, "collectives_per_step": 1 if world > 1 else 0,
             "collective_schedule": ("the all-gather of step s is a node of step s + 1's graph, started after its first selection kernel and joined before "
                                     "its first record kernel (under the 8-point / scoring kernels)") if (world > 1 and sf.pipelined) else ("after the step's last launch set" if world > 1 else None),
             "gather_bytes_per_step_per_rank": int(sf.rec_bytes * P * a.sub_batches),
             "popc_kernel_config": dict(zip(("csa_level", "rows_per_thread", "warps"), _getcfg(lib)))}
    line = _base_line(env, a, value, total_ms, "weak", clk, roof, e2e, stages, launches, extra)
    if a.sweep:
        line["popc_kernel_sweep_ms"] = sweep(lib, fe, batches[0], env.timed, a)
    if a.extras == "all" or (a.extras == "auto" and world == 1):
        line["configs"] = extras(env, a, peaks)
    return line


def extras(env, a, peaks):
    """Compact sub-records of the other BASELINE configs + the drop-in's single-call latency (N = 1)."""
    out = {}
    for cfgno in (3, 4, 5):
        t0 = time.perf_counter()
        try:
            b = parse(["--config", str(cfgno), "--steps", str(max(5, a.steps // 2)), "--warmup", "2", "--variant", a.variant, "--extras", "none"])
            line = {3: run_config3, 4: run_config4, 5: run_config5}[cfgno](env, b, peaks=peaks, brief=True)
            out[f"config{cfgno}"] = {k: line[k] for k in ("value", "unit", "ms_per_step", "scaling", "config", "stage_ms", "e2e", "roofline", "gpu_launches_per_step")
                                     if k in line}
            out[f"config{cfgno}"]["roofline"] = {k: v for k, v in line["roofline"].items() if k in ("frac", "achieved", "peak", "unit", "kernel_ms", "plan", "stage", "frac_of_2x_measured_bf16")}
            out[f"config{cfgno}"]["wall_s"] = time.perf_counter() - t0
        except Exception as exc:
            out[f"config{cfgno}"] = {"error": f"{type(exc).__name__}: {exc}"}
            env.torch.cuda.synchronize()
    try:
        pairs = _cpu_pairs_for(parse(["--config", "2"]), 6)
        out["single_call"] = single_call_latency(pairs, reps=2)
    except Exception as exc:
        out["single_call"] = {"error": f"{type(exc).__name__}: {exc}"}
    return out


# ---- config 3: loop-closure verification batch ------------------------------------------------------------

def run_config3(env, a, peaks=None, brief=False):
    torch = env.torch
    from b200slam.frontend import FrontendConfig, PairBatch, PairPipeline
    from b200slam.sharding import ShardedFrontend
    from b200slam.synthetic import tracking_pairs

    rank, world, dev = env.rank, env.world, env.dev
    n_glob, W = a.pairs, min(2 if brief else DISTINCT_WINDOWS, a.sub_batches)
    cfg = FrontendConfig(hypotheses=a.hyps, max_matches=500, threshold=0.01, precision=64, seed=1337, scoring=a.scoring)
    sf = ShardedFrontend(cfg, n_glob, variant=env.variant, sets_per_gather=a.sub_batches)   # strong scaling: the 256 pairs are split over the ranks
    lo, hi = sf.lo, sf.hi
    host = []
    for w in range(W):                                                            # every rank generates the same batch and keeps its block
        qs, ts, kq, kt = tracking_pairs(n_glob, a.nfeat, seed=300 + w, keep=0.4)
        host.append((qs[lo:hi], ts[lo:hi], kq[lo:hi], kt[lo:hi]))
    batches = [PairBatch.from_host(*h) for h in host]
    torch.cuda.synchronize()

    step, launches, graphed = sharded_step(env, sf, [batches[i % W] for i in range(a.sub_batches)], use_graph=not a.no_graph)
    clocks = Clocks(env.local)
    if rank == 0 and not brief:
        clocks.start()
    ms = env.timed(step, a.steps, a.warmup)
    sf.flush()
    total_ms = env.max_over_ranks(float(np.sum(ms)))
    value = n_glob * a.sub_batches * a.steps / (total_ms * 1e-3)
    if peaks is None and rank == 0:
        peaks = measured_peaks(env)
    hr = hamming_roofline(env, sf.fe.matcher, batches[0], float(hi - lo) * a.nfeat * a.nfeat, a.steps, peaks)

    # e2e: host descriptor / keypoint blocks of the candidate pairs -> records, through PairPipeline (pinned buffers, depth in flight)
    depth = max(1, a.e2e_depth)
    pipe = PairPipeline(hi - lo, a.nfeat, a.nfeat, cfg, variant=env.variant, depth=depth, device=dev, use_graph=not a.no_graph)
    if world > 1:
        gathered = [torch.empty((world, sf.cap, pipe.rec_bytes), dtype=torch.uint8, device=dev) for _ in range(depth)]
        for sl, g in zip(pipe.slots, gathered):
            sl["rec_dev"] = g[rank][: hi - lo]
            sl["gathered"] = g
        pipe.after_compute = lambda sl: env.dist.all_gather_into_tensor(sl["gathered"].view(-1), sl["gathered"][rank].reshape(-1))
    staged = [pipe.stage_host(*h) for h in host]

    def e2e_step():
        for i in range(a.sub_batches):
            pipe.submit(staged[i % W])

    for _ in range(2):
        e2e_step()
        torch.cuda.synchronize()
    env.barrier()
    main = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for st in pipe.streams():
        st.wait_event(e0)
    n_e2e = max(3, a.steps // 2)
    for _ in range(n_e2e):
        e2e_step()
    for st in pipe.streams():
        main.wait_stream(st)
    e1.record(main)
    env.barrier()
    e2e_total = env.max_over_ranks(e0.elapsed_time(e1))
    e2e_value = n_glob * a.sub_batches * n_e2e / (e2e_total * 1e-3)
    clk = clocks.stop() if (rank == 0 and not brief) else None
    if rank != 0:
        return None
    roof = _roof_from(hr, peaks, "hamming_knn2_i8s_kernel")
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes * a.sub_batches, "d2h_bytes_per_step": pipe.d2h_bytes * a.sub_batches,
           "api": "b200slam.frontend.PairPipeline.submit / result", "steps_in_flight": depth, "ms_per_step": e2e_total / n_e2e}
    return _base_line(env, a, value, total_ms, "strong", clk, roof, e2e, {}, launches,
                      {"value_cuda_graph": graphed, "collective_in_graph": bool(world > 1 and graphed), "pairs_this_rank": hi - lo})


# ---- config 4: high-density pairs -------------------------------------------------------------------------

def run_config4(env, a, peaks=None, brief=False):
    torch = env.torch
    from b200slam.frontend import Frontend, FrontendConfig, PairBatch, PairPipeline
    from b200slam.synthetic import tracking_pairs

    rank, world, dev = env.rank, env.world, env.dev
    P = a.pairs if not brief else min(a.pairs, 8)
    cfg = FrontendConfig(hypotheses=a.hyps, max_matches=10000, threshold=0.01, precision=64, seed=4096, scoring="cuda")
    fe = Frontend(cfg, variant=env.variant)
    qs, ts, kq, kt = tracking_pairs(P, a.nfeat, seed=4096 + rank)
    batch = PairBatch.from_host(qs, ts, kq, kt)
    one = PairBatch.from_host(qs[:1], ts[:1], kq[:1], kt[:1])
    torch.cuda.synchronize()

    def step_body():
        for _ in range(a.sub_batches):
            fe.run(batch)
    step, launches, graphed = _capture(env, step_body, use_graph=not a.no_graph)
    clocks = Clocks(env.local)
    if rank == 0 and not brief:
        clocks.start()
    ms = env.timed(step, a.steps, a.warmup)
    total_ms = env.max_over_ranks(float(np.sum(ms)))
    value = world * P * a.sub_batches * a.steps / (total_ms * 1e-3)
    # one lone pair (latency): the train-axis split and the correspondence slicing fill the machine
    lone, _, _ = _capture(env, lambda: fe.run(one), use_graph=not a.no_graph)
    lone_ms = float(np.median(env.timed(lone, max(5, a.steps), 2)))
    if peaks is None and rank == 0:
        peaks = measured_peaks(env)
    hr = hamming_roofline(env, fe.matcher, batch, float(P) * a.nfeat * a.nfeat, a.steps, peaks)
    hr1 = hamming_roofline(env, fe.matcher, one, float(a.nfeat) * a.nfeat, a.steps, peaks)
    variants_ms = hamming_variants(env, batch, 3) if not brief else None

    stages = stage_times(env, fe, batch)
    stages["lone_pair"] = stage_times(env, fe, one)

    depth = 2
    pipe = PairPipeline(P, a.nfeat, a.nfeat, cfg, variant=env.variant, depth=depth, device=dev, use_graph=not a.no_graph)
    staged = pipe.stage_host(qs, ts, kq, kt)
    for _ in range(2):
        pipe.submit(staged)
        torch.cuda.synchronize()
    env.barrier()
    main = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for st in pipe.streams():
        st.wait_event(e0)
    n_e2e = max(3, a.steps // 2) * a.sub_batches
    for _ in range(n_e2e):
        pipe.submit(staged)
    for st in pipe.streams():
        main.wait_stream(st)
    e1.record(main)
    env.barrier()
    e2e_total = env.max_over_ranks(e0.elapsed_time(e1))
    clk = clocks.stop() if (rank == 0 and not brief) else None
    if rank != 0:
        return None
    roof = _roof_from(hr, peaks, "hamming_knn2_i8s_kernel", extra={"lone_pair": {k: hr1[k] for k in ("kernel_ms", "call_ms", "achieved_top_s", "frac_of_int8_peak", "plan")},
                                                                    "hamming_variants_ms": variants_ms})
    e2e = {"value": world * P * n_e2e / (e2e_total * 1e-3), "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes * a.sub_batches,
           "d2h_bytes_per_step": pipe.d2h_bytes * a.sub_batches, "api": "b200slam.frontend.PairPipeline.submit / result", "steps_in_flight": depth,
           "ms_per_step": e2e_total / n_e2e * a.sub_batches}
    return _base_line(env, a, value, total_ms, "weak", clk, roof, e2e, stages, launches,
                      {"value_cuda_graph": graphed, "lone_pair_ms": lone_ms, "lone_pair_pairs_per_s": 1e3 / lone_ms})


# ---- config 5: relocalization sweep -----------------------------------------------------------------------

def run_config5(env, a, peaks=None, brief=False):
    torch = env.torch
    from b200slam.frontend import FrontendConfig
    from b200slam.sharding import ShardedSweep, shard_bounds

    rank, world, dev = env.rank, env.world, env.dev
    n_kf, N = a.pairs, a.nfeat
    lo, hi = shard_bounds(n_kf, rank, world)
    rng = np.random.default_rng(5000)                                             # every rank draws the same map and keeps its block
    ids = np.arange(n_kf, dtype=np.int32)
    planted = {int(k): int(r) for k, r in zip(rng.choice(n_kf, 6, replace=False), (1400, 900, 600, 400, 300, 200))}
    q_desc = rng.integers(0, 256, (N, 32), dtype=np.uint8)
    q_kp = rng.uniform(-0.6, 0.6, (N, 2)).astype(np.float32)
    kf_desc, kf_kp = [], []
    blk = 256
    for b0 in range(0, n_kf, blk):                                                # generated block-wise: only this rank's keyframes are kept
        b1 = min(n_kf, b0 + blk)
        d = rng.integers(0, 256, (b1 - b0, N, 32), dtype=np.uint8)
        k = rng.uniform(-0.6, 0.6, (b1 - b0, N, 2)).astype(np.float32)
        for i in range(b0, b1):
            if i in planted:                                                       # the query re-observes part of this keyframe
                m = planted[i]
                Pw = np.stack([rng.uniform(-6, 6, m), rng.uniform(-2, 2, m), rng.uniform(6, 30, m)], axis=1)
                k[i - b0, :m] = (Pw[:, :2] / Pw[:, 2:]).astype(np.float32)
                if i == max(planted, key=planted.get):
                    P2 = Pw + np.array([0.2, 0.0, -0.8])
                    bits = np.unpackbits(d[i - b0, :m], axis=1)
                    bits ^= (rng.random(bits.shape) < 0.06).astype(np.uint8)
                    q_desc[:m] = np.packbits(bits, axis=1)
                    q_kp[:m] = (P2[:, :2] / P2[:, 2:]).astype(np.float32)
            if lo <= i < hi:
                kf_desc.append(d[i - b0])
                kf_kp.append(k[i - b0])
    cfg = FrontendConfig(hypotheses=a.hyps, max_matches=500, threshold=0.01, precision=64, seed=1337)
    sw = ShardedSweep(kf_desc, kf_kp, ids[lo:hi], cfg, top=5, max_query_rows=2048, n_keyframes_global=n_kf)
    del kf_desc, kf_kp
    qd_host, qk_host = torch.from_numpy(q_desc).pin_memory(), torch.from_numpy(q_kp).pin_memory()
    qd_dev, qk_dev = qd_host.to(dev), qk_host.to(dev)
    sw.sweep.set_query_rows(N)
    sw.query(qd_dev if rank == 0 else None, qk_dev if rank == 0 else None, n_rows=N)
    torch.cuda.synchronize()
    l0 = env.lib.b2s_launch_count()
    sw.query(n_rows=N)
    torch.cuda.synchronize()
    launches = int(env.lib.b2s_launch_count() - l0)
    graphed = sw.capture() if not a.no_graph else False
    clocks = Clocks(env.local)
    if rank == 0 and not brief:
        clocks.start()

    def step():
        for _ in range(a.sub_batches):
            sw.replay()
    ms = env.timed(step, a.steps, a.warmup)
    total_ms = env.max_over_ranks(float(np.sum(ms)))
    value = n_kf * a.sub_batches * a.steps / (total_ms * 1e-3)
    if peaks is None and rank == 0:
        peaks = measured_peaks(env)
    m = sw.sweep
    hr = hamming_roofline(env, m.matcher, m._batch(N), float(hi - lo) * N * N, a.steps, peaks, need_second=False)   # the sweep's kernel: cross-check only
    hr_full = hamming_roofline(env, m.matcher, m._batch(N), float(hi - lo) * N * N, max(3, a.steps // 2), peaks)
    variants_ms = hamming_variants(env, m._batch(N), 3) if not brief else None

    # e2e: the query frame arrives from pinned host memory (80 KB), the gathered candidate records + per-keyframe counts go back
    out_host = torch.empty(sw.gather.buf.shape, dtype=torch.uint8).pin_memory()

    def e2e_step():
        if rank == 0:
            m.desc[m.q_row0:m.q_row0 + N].copy_(qd_host, non_blocking=True)
            m.kp[m.q_row0:m.q_row0 + N].copy_(qk_host, non_blocking=True)
        sw.replay()
        out_host.copy_(sw.gather.buf, non_blocking=True)
    e2e_ms = env.timed(e2e_step, max(3, a.steps // 2), 2)
    e2e_total = env.max_over_ranks(float(np.sum(e2e_ms)))
    counts, cand = sw.result_host()
    clk = clocks.stop() if (rank == 0 and not brief) else None
    if rank != 0:
        return None
    best = max(planted, key=planted.get)
    check = {"top_frame_ids": cand["pair_id"].tolist(), "top_match_counts": cand["n_matches"].tolist(), "top_inliers": cand["inliers"].tolist(),
             "planted_keyframe": best, "planted_found_first": bool(len(cand["pair_id"]) and int(cand["pair_id"][0]) == best),
             "mean_matches_per_keyframe": float(np.mean(counts))}
    roof = _roof_from(hr, peaks, "hamming_knn2_i8s_kernel<1, false, 2, false> (best neighbour only: BFMatcher(crossCheck=True).match never reads the second)",
                      extra={"with_second_neighbour": {k: hr_full[k] for k in ("kernel_ms", "achieved_top_s", "frac_of_int8_peak")},
                             "hamming_variants_ms": variants_ms})
    e2e = {"value": n_kf * len(e2e_ms) / (e2e_total * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(N * 40), "d2h_bytes_per_step": int(out_host.numel()),
           "api": "b200slam.sharding.ShardedSweep.query / result_host", "ms_per_step": e2e_total / len(e2e_ms)}
    return _base_line(env, a, value, total_ms, "strong", clk, roof, e2e, {}, launches,
                      {"value_cuda_graph": bool(graphed), "collective_in_graph": bool(world > 1 and graphed), "keyframes_this_rank": hi - lo,
                       "map_bytes_resident_this_rank": int(m.desc.numel() + m.kp.numel() * 4), "result_check": check})


def run_config1(env, a):
    if env.rank != 0:
        return None
    return config1_line(a, impl="b200", with_gpu=True)


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
    from b200slam.frontend import HammingMatcher
    out = {}
    keep = _getcfg(lib)
    m = HammingMatcher(variant=0)
    for rows in (2, 4):
        for warps in (4, 8):
            for csa in (0, 1, 2, 3):
                lib.b2s_hamming_set_config(csa, rows, warps)
                out[f"csa{csa}_r{rows}_w{warps}"] = float(np.mean(timed(lambda: m.knn2(batch), 5, 2)))
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
    import faulthandler
    # a stuck collective or kernel must end with a traceback, never by hanging the box until the harness kills it
    faulthandler.dump_traceback_later(int(os.environ.get("B2S_WATCHDOG_S", "840")), exit=True)
    a = parse()
    if a.impl == "reference":
        reference_main(a)
    else:
        b200_main(a)


if __name__ == "__main__":
    main()
