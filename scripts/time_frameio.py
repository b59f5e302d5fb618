#!/usr/bin/env python
"""CUDA-event timing of the frame ingest / egress kernels and of a pinned H2D / D2H copy at 1080p (debug aid)."""
import sys
import torch
sys.path.insert(0, ".")
from drba_b200.tools import frame_egress_u8, frame_ingest_u8  # noqa: E402

h, w, H, W = 1080, 1920, 1088, 1920
u8 = torch.randint(0, 256, (h, w, 3), dtype=torch.uint8).pin_memory()
d8 = u8.cuda()
f = torch.empty((1, 3, H, W), device="cuda")
o8 = torch.empty((h, w, 3), dtype=torch.uint8, device="cuda")
host_out = torch.empty((h, w, 3), dtype=torch.uint8).pin_memory()


def t(fn, n=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


print("ingest kernel us", round(t(lambda: frame_ingest_u8(d8, (H, W), out=f)), 1))
print("egress kernel us", round(t(lambda: frame_egress_u8(f, (h, w), out=o8)), 1))
print("H2D 6.2 MB us", round(t(lambda: d8.copy_(u8, non_blocking=True)), 1))
print("D2H 6.2 MB us", round(t(lambda: host_out.copy_(o8, non_blocking=True)), 1))
print("clone 25 MB us", round(t(lambda: f.clone()), 1))
