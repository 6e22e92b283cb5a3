"""Diagnostics (GPU): per-SM issue rates of the instructions the kernels lean on."""
import ctypes as C
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "monocular-visual-slam_b200")]
import torch
from b200slam import _capi
lib = _capi.load_library()
sink = torch.zeros(4, dtype=torch.int32, device="cuda")
for name, which in _capi.PIPE_IDS.items():
    ops = C.c_double()
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); lib.b2s_pipe_microbench(which, 2000, 8, C.byref(ops), C.c_void_p(sink.data_ptr()), None); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{name:16s} {ops.value / (ms * 1e-3) / 148 / 1.965e9:7.1f} thread-instr/clk/SM")
