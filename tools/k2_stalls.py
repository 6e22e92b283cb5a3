"""Diagnostics (GPU): where the i8 Hamming kernel's MMA thread / producer / epilogue wait.
Usage: python tools/k2_stalls.py [pairs] [nfeat]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "monocular-visual-slam_b200")]
import ctypes as C

import numpy as np
import torch

from b200slam import _capi
from b200slam.frontend import HammingMatcher, PairBatch

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 296
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
lib = _capi.load_library()
rng = np.random.default_rng(0)
qs = [rng.integers(0, 256, (n, 32), dtype=np.uint8) for _ in range(pairs)]
ts = [rng.integers(0, 256, (n, 32), dtype=np.uint8) for _ in range(pairs)]
batch = PairBatch.from_host(qs, ts)
import os
m = HammingMatcher(variant=_capi.VARIANT_I8MMA1 if os.environ.get('K2S') else _capi.VARIANT_I8MMA)
for _ in range(3):
    m.knn2(batch)
torch.cuda.synchronize()
# MMA issue-rate microbenchmark, N = 128 and 256
for nd, var in ((128, 0), (256, 0), (128, 1), (128, 2), (128, 3), (128, 4)):
    macs = C.c_double()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    arg = nd | (var << 16)
    lib.b2s_mma_microbench(2000, arg, C.byref(macs), None)
    torch.cuda.synchronize()
    e0.record(); lib.b2s_mma_microbench(2000, arg, C.byref(macs), None); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    per = (18 if var else 8)
    print(f"mma N={nd} variant {var}: {2*macs.value/ms/1e15:.2f} POP/s, {ms*1e-3*1.965e9/(2000*per):.1f} clk/instr")
names = ["mma_total", "mma_wait_tempty", "mma_wait_full", "prod_wait_empty", "epi_total", "epi_wait_tfull"]
for mode in ((0, 16, 1, 4, 8, 12, 20, 24) if os.environ.get('K2S') else (0, 1, 2, 3)):
    dbg = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
    lib.b2s_hamming_i8_debug(C.c_void_p(dbg.data_ptr()), mode)
    m.knn2(batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m.knn2(batch); e1.record(); torch.cuda.synchronize()
    lib.b2s_hamming_i8_debug(None, 0)
    d = dbg.cpu().numpy().reshape(148, 8).astype(np.float64)
    tp = d[:, 6]
    print(f"mode {mode} (bit0: no epilogue work, bit1: no ring reloads, bit2: no column minima, bit3: no top-2, bit4: 32x32b epilogue): knn2 call {e0.elapsed_time(e1):.3f} ms; tile pairs/CTA {tp.mean():.0f}")
    print("   " + "  ".join(f"{nm} {d[:, i].sum() / tp.sum():.0f}" for i, nm in enumerate(names)))
