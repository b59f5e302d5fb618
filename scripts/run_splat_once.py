#!/usr/bin/env python
"""One softsplat call per case (for ncu captures): python scripts/run_splat_once.py C H W kind variant"""
import sys
import torch
sys.path.insert(0, ".")
from drba_b200.softsplat import softsplat  # noqa: E402
from scripts.bench_splat import smooth_flow  # noqa: E402

c, h, w, kind, variant = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5])
x = torch.randn((1, c, h, w), device="cuda")
flow = (smooth_flow(h, w, 8.0, 1) if kind == "smooth" else smooth_flow(h, w, 2.0, 1) + 6.5 if kind == "gentle"
        else 8 * torch.randn((1, 2, h, w), device="cuda"))
metric = torch.randn((1, 1, h, w), device="cuda")
for _ in range(2):
    softsplat(x, flow, metric, "soft", _variant=variant)
torch.cuda.synchronize()
