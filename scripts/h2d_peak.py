"""Bare pinned host-to-device copy rate of the box (what the end-to-end number of bench.py is bounded by).
Run on a GPU box: python scripts/h2d_peak.py"""
import torch, time
for mb in (64, 256, 1024):
    h = torch.empty(mb << 20, dtype=torch.uint8, pin_memory=True); d = torch.empty_like(h, device="cuda:0")
    d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = max(4, 4096 // mb)
    e0.record()
    for _ in range(n): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print("H2D %4d MiB x %d: %.2f GB/s" % (mb, n, n * (mb << 20) / e0.elapsed_time(e1) / 1e6))
