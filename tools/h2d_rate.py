"""Diagnostics (GPU): pinned host -> device and device -> host copy rate of this box (the ceiling of
bench.py's end-to-end number: 80 264 bytes per frame pair go up, 6 512 come down)."""
import torch
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    gbs = 10 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
    print(f"{name}: {gbs:.1f} GB/s" + (f"  -> end-to-end ceiling {gbs * 1e9 / 80264:.0f} frame-pairs/s" if name == "H2D" else ""))
